"""Import-time stand-in for POT (`import ot` in metrics.py); exact OT is outside the hot path."""


def emd2(*a, **k):
    raise NotImplementedError("POT stand-in: wasserstein_distance is outside the SVGD hot path")
