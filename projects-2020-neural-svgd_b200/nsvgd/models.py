"""Drop-in for the SVGD plug-in of nvgd/src/models.py: `KernelGradient` (models.py:622-689) and the
`Particles` container (models.py:61-211).  Same constructor arguments and aux-dict contract; arrays
are CUDA float32 tensors; `key` is an int seed or a torch.Generator (JAX PRNG keys do not exist here)."""
from __future__ import annotations

import torch

from . import engine, kernels, metrics, stein, utils
from .utils import vmap


class KernelGradient:
    """Computes the SVGD approximation to grad(KL), ie
    phi*(y) = E[grad(log p)(y) k(x, y) + div(k)(x, y)]                    models.py:622-652"""

    def __init__(self,
                 target_logp: callable = None,
                 get_target_logp: callable = None,
                 kernel=kernels.get_rbf_kernel,
                 bandwidth=None,
                 scaled=False,
                 lambda_reg=1 / 2,
                 use_hutchinson: bool = False):
        if target_logp:
            assert not get_target_logp
            self.get_target_logp = lambda *args: target_logp
        elif get_target_logp:
            self.get_target_logp = get_target_logp
        else:
            # the reference *returns* the ValueError here (models.py:645-646, a bug); we raise it
            raise ValueError("One of target_logp and get_target_logp must be given.")
        if use_hutchinson:
            raise NotImplementedError("use_hutchinson (random estimator) is outside the hot path (SURVEY.md 2)")
        self.bandwidth = bandwidth
        self.kernel = kernel
        self.lambda_reg = lambda_reg
        self.rundata = {}
        self.scaled = scaled
        self.use_hutchinson = use_hutchinson

    def _bandwidth(self, inducing_particles):
        # truthiness test as in models.py:657/680: None or 0 -> median heuristic
        if isinstance(self.bandwidth, torch.Tensor) or self._ard():
            return self.bandwidth
        if self.bandwidth:
            return self.bandwidth
        # the plug-in is called once per SVGD step on slowly moving particles: keep the warm-start state of the filtered
        # select between calls (engine.median_heuristic; bit-identical bandwidth, sampling + bracket select skipped on a hit)
        if inducing_particles.is_cuda:
            key = (inducing_particles.device, tuple(inducing_particles.shape))
            if getattr(self, "_median_hint_key", None) != key:
                self._median_hint, self._median_hint_key = engine.new_median_hint(inducing_particles.device), key
            return engine.median_heuristic(inducing_particles, None, hint=self._median_hint).reshape(())
        return kernels.median_heuristic(inducing_particles)

    def _ard(self):
        """a (d,) bandwidth vector: served by stein.get_phistar (unscaled field); the fused scaled step needs a scalar"""
        b = self.bandwidth
        if b is None or isinstance(b, (int, float)):
            return False
        return torch.as_tensor(b).squeeze().ndim == 1 and torch.as_tensor(b).numel() > 1

    def get_field(self, inducing_particles, batch=None):
        """return -phistar                                                 models.py:654-660"""
        target_logp = self.get_target_logp(batch)
        bandwidth = self._bandwidth(inducing_particles)
        kernel = self.kernel(bandwidth)
        phi = stein.get_phistar(kernel, target_logp, inducing_particles)
        return utils.negative(phi), bandwidth

    def get_field_scaled(self, inducing_particles, batch=None):
        """-alpha phi*, alpha = ksd / (2 lambda l2)                        models.py:677-689
        One fused engine pass yields phi, sum_i(s_i.phi_i + div phi_i) and sum_i|phi_i|^2, replacing
        the reference's n extra phi evaluations (l2_norm_squared) and its jacfwd (stein_discrepancy)."""
        target_logp = self.get_target_logp(batch)
        bandwidth = self._bandwidth(inducing_particles)
        kernel = self.kernel(bandwidth)
        phi = stein.get_phistar(kernel, target_logp, inducing_particles)
        if phi.ard:
            raise NotImplementedError("scaled=True needs the closed-form KSD, which is built for a scalar bandwidth")
        vals, stats = engine.phistar(inducing_particles, inducing_particles, phi.scores, phi.h,
                                     Sq=phi.scores, stats=True)
        ksd = stats[0] / stats[2]
        l2_phi_squared = stats[1] / stats[2]
        alpha = (ksd / (2 * self.lambda_reg * l2_phi_squared)).to(torch.float32)
        field = utils.mul(phi, -alpha)
        field._cached = (inducing_particles, vals)   # reuse when evaluated at the inducing particles
        self._last_scaled = {"ksd": ksd, "l2": l2_phi_squared, "alpha": alpha}
        return field, bandwidth

    def gradient(self, batch, particles, aux=False):
        """Compute approximate KL gradient.                                models.py:662-675
        args:
            batch: minibatch data used to estimate logp (can be None)
            particles: array of shape (n, d)
        """
        target_logp = self.get_target_logp(batch)
        v, h = self.get_field_scaled(particles, batch) if self.scaled \
            else self.get_field(particles, batch)
        cached = getattr(v, "_cached", None)
        if cached is not None and cached[0] is particles:
            grads = cached[1] * v.scale
        else:
            grads = vmap(v)(particles)
        if aux:
            h_out = h if isinstance(h, torch.Tensor) else torch.tensor(float(h), device=particles.device)
            auxd = {"bandwidth": h_out}
            if hasattr(target_logp, "batch_logp"):
                auxd["logp"] = target_logp.batch_logp(particles)
            else:
                auxd["logp"] = torch.func.vmap(target_logp)(particles)
            return grads, auxd
        else:
            return grads

    # ---- fused fast path used by Particles when the optimizer is plain sgd ---------------------------
    def step_sgd(self, particles, batch, lr):
        """median (if needed) -> phi* (+ KSD, l2) -> x += lr alpha phi, all enqueued on the stream by
        ONE composite C call (nsvgd_step_sgd_f32).  Mutates `particles`; returns (grads, aux)."""
        target_logp = self.get_target_logp(batch)
        h = self._bandwidth(particles)
        _ = stein._require_rbf(self.kernel(h), "KernelGradient")
        s = stein.score(target_logp, particles)
        if hasattr(target_logp, "batch_logp"):
            logp_vals = target_logp.batch_logp(particles)
        else:
            logp_vals = torch.func.vmap(target_logp)(particles)
        step_aux, grads = engine.svgd_step_sgd(particles, s, h, lr, self.scaled, self.lambda_reg, want_grads=True)
        h_out = h if isinstance(h, torch.Tensor) else torch.full((), float(h), device=particles.device)   # capture-safe
        return grads, {"bandwidth": h_out, "logp": logp_vals,
                       "global_grad_norm": torch.sqrt(step_aux[3]).to(torch.float32),
                       "global_grad_norm_post_update": torch.sqrt(step_aux[4]).to(torch.float32)}


class Particles:
    """Container class for particles, particle optimizer, particle update step method, and
    particle metrics.                                                       models.py:61-211"""

    def __init__(self,
                 key,
                 gradient: callable,
                 init_samples,
                 learning_rate=1e-2,
                 optimizer="sgd",
                 custom_optimizer=None,
                 n_particles: int = 50,
                 compute_metrics=None,
                 device="cuda"):
        self.gradient = gradient
        self.n_particles = n_particles
        self.device = torch.device(device)
        if isinstance(key, torch.Generator):
            self.generator = key
        else:
            self.generator = torch.Generator(device="cpu")
            self.generator.manual_seed(int(key))
        self.init_samples = init_samples
        self.particles = self.init_particles()

        if custom_optimizer:
            self.optimizer_str = "custom"
            self.learning_rate = None
            self.opt = custom_optimizer     # callable(grads, state, particles) -> (updates, state)
            self.optimizer_state = None
        else:
            if optimizer not in ("sgd", "adam"):
                raise KeyError(optimizer)   # utils.optimizer_mapping (utils.py:711-715)
            self.optimizer_str = optimizer
            self.learning_rate = learning_rate
            self.opt = None
            self.optimizer_state = None
            if optimizer == "adam":
                self.optimizer_state = (torch.zeros_like(self.particles), torch.zeros_like(self.particles), 0)
        self.step_counter = 0
        self.rundata = {}
        self.donedone = False
        self.compute_metrics = compute_metrics

    def init_particles(self, key=None):
        """Returns a tensor of shape (n, d) containing particles.          models.py:107-117"""
        if callable(self.init_samples):
            particles = self.init_samples(self.n_particles, self.generator if key is None else key)
        else:
            particles = self.init_samples
            self.n_particles = len(particles)
        particles = torch.as_tensor(particles, dtype=torch.float32).to(self.device).contiguous().clone()
        self.d = particles.shape[1]
        return particles

    def get_params(self):
        return self.particles

    def _step(self, particles, optimizer_state, params):
        """Updates particles in the direction given by self.gradient      models.py:143-167
        Returns: particles (updated), optimizer_state (updated), grad_aux dict."""
        owner = getattr(self.gradient, "__self__", None)
        if isinstance(owner, KernelGradient) and self.optimizer_str == "sgd" and not owner._ard():
            particles = particles.clone()
            grads, grad_aux = owner.step_sgd(particles, params, self.learning_rate)
            return particles, optimizer_state, grad_aux
        grads, grad_aux = self.gradient(params, particles, aux=True)
        particles = particles.clone()
        if self.optimizer_str == "sgd":
            updated_grads = -self.learning_rate * grads
            particles += updated_grads
        elif self.optimizer_str == "adam":
            m1, m2, t = optimizer_state
            before = particles.clone()
            engine.adam_update(particles, grads.contiguous(), m1, m2, self.learning_rate, t + 1)
            optimizer_state = (m1, m2, t + 1)
            updated_grads = particles - before
        else:
            updated_grads, optimizer_state = self.opt(grads, optimizer_state, particles)
            particles += updated_grads
        grad_aux.update({
            "global_grad_norm": torch.linalg.vector_norm(grads),
            "global_grad_norm_post_update": torch.linalg.vector_norm(updated_grads),
        })
        return particles, optimizer_state, grad_aux

    def step(self, params):
        """Log rundata, take step. Mutates state                            models.py:169-176"""
        updated_particles, self.optimizer_state, auxdata = self._step(
            self.particles, self.optimizer_state, params)
        self.log(auxdata)
        self.particles = updated_particles
        self.step_counter += 1
        return None

    def step_many(self, n_steps: int, params=None):
        """`n_steps` calls of `step(params)` with ONE CUDA-graph replay per step.

        At n=100, d=2 a step is ~40 kernels of a few microseconds: the loop is bound by Python, ctypes and allocator
        work, not by the GPU.  The step (scores -> median bandwidth -> phi* + KSD/l2 -> scaled sgd update, plus the
        per-step log entries of `_log`) is captured once and replayed; the log lands in preallocated device arrays
        indexed by a device-side counter, and `rundata` ends up with the same keys and values as the eager loop.
        The tracer (`compute_metrics`, every 10th step) runs eagerly between replays, as in `log()`.
        Falls back to the eager loop unless the plug-in is `KernelGradient` with the sgd optimizer on CUDA."""
        owner = getattr(self.gradient, "__self__", None)
        if not (isinstance(owner, KernelGradient) and self.optimizer_str == "sgd" and self.particles.is_cuda and n_steps > 0
                and not owner._ard()):
            for _ in range(n_steps):
                self.step(params)
            return None
        X = self.particles
        n, d = X.shape
        dev = X.device
        log_particles = self.d < 400                                   # models.py:189
        if self._step_many_persistent(owner, n_steps, params, log_particles):
            return None
        bufs = {"bandwidth": torch.empty(n_steps, device=dev), "logp": torch.empty((n_steps, n), device=dev),
                "global_grad_norm": torch.empty(n_steps, device=dev),
                "global_grad_norm_post_update": torch.empty(n_steps, device=dev)}
        if log_particles:
            bufs.update({"particles": torch.empty((n_steps, n, d), device=dev), "mean": torch.empty((n_steps, d), device=dev),
                         "std": torch.empty((n_steps, d), device=dev)})
        idx = torch.zeros(1, dtype=torch.int64, device=dev)

        def body():
            if log_particles:
                bufs["particles"].index_copy_(0, idx, X.unsqueeze(0))
                bufs["mean"].index_copy_(0, idx, torch.mean(X, dim=0, keepdim=True))
                bufs["std"].index_copy_(0, idx, torch.std(X, dim=0, unbiased=False, keepdim=True))
            _, aux = owner.step_sgd(X, params, self.learning_rate)
            bufs["bandwidth"].index_copy_(0, idx, aux["bandwidth"].reshape(1).to(torch.float32))
            bufs["logp"].index_copy_(0, idx, aux["logp"].reshape(1, n).to(torch.float32))
            bufs["global_grad_norm"].index_copy_(0, idx, aux["global_grad_norm"].reshape(1))
            bufs["global_grad_norm_post_update"].index_copy_(0, idx, aux["global_grad_norm_post_update"].reshape(1))
            idx.add_(1)

        # warm-up on a side stream (library state, allocator pools), then restore the particles and capture
        keep = X.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream(dev).wait_stream(side)
        X.copy_(keep)
        idx.zero_()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            body()
        X.copy_(keep)
        idx.zero_()
        first = self.step_counter
        for k in range(n_steps):
            if self.step_counter % 10 == 0 and self.compute_metrics:
                aux_metrics = self.compute_metrics(self.particles)
                metrics.append_to_log(self.rundata, {key: (self.step_counter, v.clone() if isinstance(v, torch.Tensor) else v)
                                                     for key, v in aux_metrics.items()})
            graph.replay()
            self.step_counter += 1
        for k in range(n_steps):                                       # same keys / order as log() would have produced
            entry = {}
            if log_particles:
                entry.update({"step": first + k, "particles": bufs["particles"][k], "mean": bufs["mean"][k], "std": bufs["std"][k]})
            entry.update({"bandwidth": bufs["bandwidth"][k], "logp": bufs["logp"][k],
                          "global_grad_norm": bufs["global_grad_norm"][k],
                          "global_grad_norm_post_update": bufs["global_grad_norm_post_update"][k]})
            metrics.append_to_log(self.rundata, entry)
        return None

    def _step_many_persistent(self, owner, n_steps, params, log_particles):
        """All `n_steps` steps in ONE kernel launch (csrc/small_loop.cu) when the problem is latency-bound
        and everything the step needs lives on the device: n * d small, a target with device-side scores (funnel,
        Gaussian), the default RBF kernel with a scalar or median bandwidth, tracers that are MMDs against fixed target
        samples.  Returns False (nothing done) otherwise; `rundata` gets the same keys and order as the eager loop."""
        import ctypes as C
        import os
        from . import _lib
        if os.environ.get("NSVGD_SMALL_LOOP", "1") == "0":
            return False
        X = self.particles
        n, d = X.shape
        lib = _lib.load()
        logp = owner.get_target_logp(params)
        spec = getattr(logp, "device_spec", None)
        tr = getattr(self.compute_metrics, "tracer_spec", None) if self.compute_metrics else None
        bw = owner.bandwidth
        if (spec is None or owner.kernel is not kernels.get_rbf_kernel or not lib.nsvgd_loop_small_supported(n, d)
                or (self.compute_metrics and tr is None) or isinstance(bw, torch.Tensor) or (spec[0] == 0 and d < 2)):
            return False
        dev = X.device
        K = int(n_steps)
        f32 = dict(dtype=torch.float32, device=dev)
        bufs = {"bandwidth": torch.empty(K, **f32), "logp": torch.empty((K, n), **f32), "global_grad_norm": torch.empty(K, **f32),
                "global_grad_norm_post_update": torch.empty(K, **f32)}
        if log_particles:
            bufs.update({"particles": torch.empty((K, n, d), **f32), "mean": torch.empty((K, d), **f32), "std": torch.empty((K, d), **f32)})
        tparams = spec[1].to(dev) if spec[1] is not None else None
        first = self.step_counter
        mmd_out = mmd_step = mmd_count = Y = None
        if tr is not None:
            slots = K // 10 + 2
            Y = tr[1].to(device=dev, dtype=torch.float32).contiguous()
            mmd_out = torch.zeros((slots, 2), **f32)
            mmd_step = torch.zeros(slots, dtype=torch.int64, device=dev)
            mmd_count = torch.zeros(1, dtype=torch.int32, device=dev)
        mode = _lib.MEDIAN_MODES[engine.DEFAULT_MEDIAN_MODE]
        p = _lib.ptr
        _lib.check(lib.nsvgd_svgd_loop_small_f32(
            _lib.stream_ptr(dev), p(X), n, d, int(spec[0]), p(tparams), int(spec[2]), float(spec[3]), mode,
            float(bw) if bw else 0.0, float(self.learning_rate), int(bool(owner.scaled)), float(owner.lambda_reg), K, first,
            p(bufs.get("particles")), p(bufs.get("mean")), p(bufs.get("std")), p(bufs["bandwidth"]), p(bufs["logp"]),
            p(bufs["global_grad_norm"]), p(bufs["global_grad_norm_post_update"]), 0, p(Y), 0 if Y is None else Y.shape[0], 10,
            0 if tr is None else int(tr[0]), 1.0 if tr is None else float(tr[2]), p(mmd_out), p(mmd_step), p(mmd_count)))
        self.step_counter += K
        if tr is not None:
            cnt = int(mmd_count.item())
            steps = mmd_step[:cnt].tolist()
            for i in range(cnt):
                entry = {name: (steps[i], mmd_out[i, j].clone()) for j, name in enumerate(tr[3])}
                metrics.append_to_log(self.rundata, entry)
        for k in range(K):
            entry = {}
            if log_particles:
                entry.update({"step": first + k, "particles": bufs["particles"][k], "mean": bufs["mean"][k], "std": bufs["std"][k]})
            entry.update({"bandwidth": bufs["bandwidth"][k], "logp": bufs["logp"][k],
                          "global_grad_norm": bufs["global_grad_norm"][k],
                          "global_grad_norm_post_update": bufs["global_grad_norm_post_update"][k]})
            metrics.append_to_log(self.rundata, entry)
        return True

    def log(self, grad_aux=None):
        """models.py:178-185 (tracer cadence: every 10 steps)"""
        metrics.append_to_log(self.rundata, self._log(self.particles, self.step_counter))
        if self.step_counter % 10 == 0 and self.compute_metrics:
            aux_metrics = self.compute_metrics(self.particles)
            metrics.append_to_log(self.rundata,
                                  {k: (self.step_counter, v) for k, v in aux_metrics.items()})
        if grad_aux is not None:
            metrics.append_to_log(self.rundata, grad_aux)

    def _log(self, particles, step):
        """models.py:187-197.  Device tensors are logged as-is (no forced host copy per step)."""
        auxdata = {}
        if self.d < 400:
            auxdata.update({
                "step": step,
                "particles": particles,
                "mean": torch.mean(particles, dim=0),
                "std": torch.std(particles, dim=0, unbiased=False),
            })
        return auxdata

    def done(self):
        """converts rundata into arrays                                     models.py:199-211"""
        if self.donedone:
            print("already done.")
            return

        def to_array(v):
            if len(v) and isinstance(v[0], torch.Tensor):
                return torch.stack([t.detach() for t in v]).cpu()
            if len(v) and isinstance(v[0], tuple) and isinstance(v[0][1], torch.Tensor):
                return torch.tensor([[float(s), float(t)] for s, t in v])
            return torch.as_tensor(v)
        self.rundata = {k: to_array(v) for k, v in self.rundata.items()}
        self.donedone = True
