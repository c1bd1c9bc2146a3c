"""Drop-in for the MMD part of nvgd/src/metrics.py (metrics.py:88-127) + append_to_log."""
from __future__ import annotations

from . import engine, kernels


def append_to_log(dct, update_dict):
    """metrics.py:9-15"""
    for key, newvalue in update_dict.items():
        dct.setdefault(key, []).append(newvalue)
    return dct


def get_mmd(kernel=None):
    """metrics.py:88-105: mmd(xs, ys) = mean offdiag Kxx + mean offdiag Kyy - 2 mean Kxy."""
    kernel = kernels.get_rbf_kernel(1.) if kernel is None else kernel
    if not getattr(kernel, "is_rbf", False):
        raise NotImplementedError("get_mmd: only RBF / funnel kernels dispatch to the CUDA engine")
    if getattr(kernel, "is_full", False):
        def mmd_full(xs, ys):                     # full-matrix Gaussian kernel: unit-bandwidth engine call on x Sigma^T
            sig_t = kernel.sigma.to(device=xs.device, dtype=xs.dtype).T.contiguous()
            return engine.mmd(engine.matmul(xs, sig_t), engine.matmul(ys, sig_t), 1.0).to(xs.dtype)
        return mmd_full
    if not kernel.funnel and kernel.is_ard():
        def mmd_ard(xs, ys):                      # vector bandwidth: unit-bandwidth engine call on x / h
            sc = kernel.ard_scale(xs.device)
            return engine.mmd(engine.scale_cols(xs, sc, True), engine.scale_cols(ys, sc, True), 1.0).to(xs.dtype)
        return mmd_ard
    h = float(kernel.h())

    def mmd(xs, ys):
        return engine.mmd(xs, ys, h, funnel=kernel.funnel).to(xs.dtype)
    return mmd


def get_mmd_tracer(target_samples, kernel=None):
    """metrics.py:112-117"""
    mmd = get_mmd(kernel)

    def compute_mmd(particles):
        return {"mmd": mmd(particles, target_samples)}
    k = kernels.get_rbf_kernel(1.) if kernel is None else kernel
    if not getattr(k, "funnel", False) and not getattr(k, "is_full", False) and not k.is_ard():
        compute_mmd.tracer_spec = (1, target_samples, float(k.h()), ("mmd",))          # device-side cadence (small_loop.cu)
    return compute_mmd


def get_funnel_tracer(target_samples):
    """metrics.py:120-127"""
    rbf_mmd = get_mmd(kernels.get_rbf_kernel(1.))
    funnel_mmd = get_mmd(kernels.get_funnel_kernel(1.))

    def compute_mmd(particles):
        return {"rbf_mmd": rbf_mmd(particles, target_samples),
                "funnel_mmd": funnel_mmd(particles, target_samples)}
    compute_mmd.tracer_spec = (2, target_samples, 1.0, ("rbf_mmd", "funnel_mmd"))
    return compute_mmd
