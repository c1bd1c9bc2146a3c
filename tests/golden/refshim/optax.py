"""Stand-in for optax==0.0.1 (pinned in the reference's setup.py:17-20): the gradient transformations
`Particles` uses, restated from optax's published update rules.
sgd(lr): u = -lr * g.   adam(lr, b1, b2, eps): m,v EMAs with bias correction, u = -lr * m_hat/(sqrt(v_hat)+eps).
apply_updates(p, u) = p + u.   global_norm(g) = sqrt(sum g^2)."""
from typing import NamedTuple, Callable

import torch
from torch.utils._pytree import tree_map, tree_flatten


class GradientTransformation(NamedTuple):
    init: Callable
    update: Callable


def constant_schedule(value):
    return lambda count: value


def scale(step_size):
    return GradientTransformation(lambda p: (), lambda u, s, p=None: (tree_map(lambda g: step_size * g, u), s))


def scale_by_schedule(fn):
    def update(u, s, p=None):
        return tree_map(lambda g: fn(s) * g, u), s + 1
    return GradientTransformation(lambda p: 0, update)


def chain(*ts):
    def init(p):
        return tuple(t.init(p) for t in ts)

    def update(u, s, p=None):
        new = []
        for t, st in zip(ts, s):
            u, st = t.update(u, st, p)
            new.append(st)
        return u, tuple(new)
    return GradientTransformation(init, update)


def sgd(learning_rate, momentum=None, nesterov=False):
    assert not momentum
    return scale(-learning_rate)


class AdamState(NamedTuple):
    count: int
    mu: object
    nu: object


def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    def init(p):
        return AdamState(0, tree_map(torch.zeros_like, p), tree_map(torch.zeros_like, p))

    def update(u, s, p=None):
        mu = tree_map(lambda g, m: (1 - b1) * g + b1 * m, u, s.mu)
        nu = tree_map(lambda g, v: (1 - b2) * (g * g) + b2 * v, u, s.nu)
        count = s.count + 1
        mu_hat = tree_map(lambda m: m / (1 - b1 ** count), mu)
        nu_hat = tree_map(lambda v: v / (1 - b2 ** count), nu)
        upd = tree_map(lambda m, v: -learning_rate * m / (torch.sqrt(v + eps_root) + eps), mu_hat, nu_hat)
        return upd, AdamState(count, mu, nu)
    return GradientTransformation(init, update)


def apply_updates(params, updates):
    return tree_map(lambda p, u: p + u, params, updates)


def global_norm(updates):
    leaves, _ = tree_flatten(updates)
    return torch.sqrt(sum(torch.sum(l * l) for l in leaves))


def add_noise(*a, **k):
    raise NotImplementedError("optax stand-in: SGLD noise is outside the SVGD hot path")
