import torch
from torch.utils._pytree import tree_flatten


def ravel_pytree(tree):
    leaves, _ = tree_flatten(tree)
    return torch.cat([l.reshape(-1) for l in leaves]), None
