"""Drop-in for flows.svgd_flow (flows.py:70-110): the host loop that wires target/proposal ->
KernelGradient -> Particles.  neural_svgd_flow / sgld_flow are other methods (out of scope)."""
from __future__ import annotations

import warnings

import torch

from . import kernels, metrics, models

default_num_particles = 50
default_num_steps = 100


def svgd_flow(key,
              setup,
              n_particles=default_num_particles,
              n_steps=default_num_steps,
              particle_lr=1e-1,
              lambda_reg=1 / 2,
              particle_optimizer="sgd",
              scaled=True,
              bandwidth=1.,
              compute_metrics=None,
              catch_exceptions: bool = True,
              device="cuda",
              cuda_graph: bool = False):
    """Same arguments, defaults and return value (kernel_gradient, particles, err) as flows.py:70-110.
    cuda_graph=True (extension): the steps are replayed from one captured CUDA graph (Particles.step_many)."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(int(key))
    target, proposal = setup.get()

    kernel_gradient = models.KernelGradient(target_logp=target.logpdf,
                                            kernel=kernels.get_rbf_kernel,
                                            bandwidth=bandwidth,
                                            lambda_reg=lambda_reg,
                                            scaled=scaled)

    if compute_metrics is None:
        tgt = torch.as_tensor(target.sample(500, gen), dtype=torch.float32).to(device)
        compute_metrics = metrics.get_mmd_tracer(tgt)
    svgd_particles = models.Particles(key=gen,
                                      gradient=kernel_gradient.gradient,
                                      init_samples=proposal.sample,
                                      n_particles=n_particles,
                                      learning_rate=particle_lr,
                                      optimizer=particle_optimizer,
                                      compute_metrics=compute_metrics,
                                      device=device)
    if cuda_graph:
        try:
            svgd_particles.step_many(n_steps, None)
        except Exception as err:
            if catch_exceptions:
                warnings.warn("caught error!")
                return kernel_gradient, svgd_particles, err
            raise
        svgd_particles.done()
        return kernel_gradient, svgd_particles, None
    for _ in range(n_steps):
        try:
            svgd_particles.step(None)
        except Exception as err:
            if catch_exceptions:
                warnings.warn("caught error!")
                return kernel_gradient, svgd_particles, err
            else:
                raise err
    svgd_particles.done()
    return kernel_gradient, svgd_particles, None
