"""jax.ops.index_update / index of the jax 0.2.x era (removed upstream in 0.3.2)."""
import torch


class _Index:
    def __getitem__(self, idx):
        return idx


index = _Index()


def index_update(x, idx, y):
    out = x.clone()
    if isinstance(idx, list):
        idx = tuple(torch.as_tensor(i) if isinstance(i, (list, tuple)) else i for i in idx)
    out[idx] = y
    return out
