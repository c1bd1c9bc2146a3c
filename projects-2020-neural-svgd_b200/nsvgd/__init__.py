"""nsvgd — B200-native drop-in for the SVGD particle-interaction hot path of
ratschlab/projects-2020-Neural-SVGD (nvgd.src.{kernels,stein,models,metrics,utils,flows}).

Same function names and argument meaning as the reference; arrays are CUDA float32 torch
tensors, compute is hand-written CUDA (libnsvgd.so, C ABI in include/nsvgd.h).  No fallback.
"""
from . import _lib, engine, kernels, stein, utils, metrics, models, flows, distributions, nets, autodiff  # noqa: F401
from .utils import vmap  # noqa: F401

__all__ = ["kernels", "stein", "utils", "metrics", "models", "flows", "distributions", "nets", "autodiff", "engine", "vmap"]
