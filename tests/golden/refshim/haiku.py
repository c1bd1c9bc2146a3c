"""Import-time stand-in for dm-haiku: the hot path (KernelGradient / Particles) never touches haiku, but
`models.py` / `nets.py` import it and define hk.Module subclasses at module level."""


class Module:
    def __init__(self, name=None):
        self.name = name


class _Init:
    def __init__(self, *a, **k):
        pass


class initializers:
    RandomNormal = _Init
    VarianceScaling = _Init
    Constant = _Init


class Transformed:
    def __init__(self, f):
        self.f = f


def transform(f):
    return Transformed(f)


def without_apply_rng(t):
    return t


def __getattr__(name):
    raise AttributeError(f"haiku stand-in: '{name}' is outside the SVGD hot path and not provided")
