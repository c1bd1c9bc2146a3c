"""Deterministic stand-in for jax.random: keys are ints, samples come from a torch Generator."""
import torch


def PRNGKey(seed):
    return int(seed)


def split(key, num=2):
    return tuple(int(key) * 1000003 + 7919 * (i + 1) for i in range(num))


def _gen(key):
    g = torch.Generator()
    g.manual_seed(int(key) % (2**63 - 1))
    return g


def normal(key, shape=(), dtype=None):
    return torch.randn(tuple(shape), generator=_gen(key), dtype=torch.get_default_dtype())


def uniform(key, shape=(), dtype=None, minval=0., maxval=1.):
    return torch.rand(tuple(shape), generator=_gen(key), dtype=torch.get_default_dtype()) * (maxval - minval) + minval


def permutation(key, x):
    n = x if isinstance(x, int) else x.shape[0]
    p = torch.randperm(n, generator=_gen(key))
    return p if isinstance(x, int) else x[p]
