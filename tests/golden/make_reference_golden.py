"""Generate tests/golden/ref_*.npz by executing the UNMODIFIED reference sources in /root/reference.

    python tests/golden/make_reference_golden.py          # writes ref_<case>_f32.npz and ref_<case>_f64.npz

JAX is not installable here, so the reference's third-party dependencies are replaced by the stand-ins in
tests/golden/refshim/ (jax.numpy -> CPU torch, vmap/grad/jacfwd -> torch.func, optax/haiku/ot stubs; see
refshim/README.md).  The reference modules themselves are imported from where they lie and every number stored
below comes out of the reference's own functions:
    kernels.median_heuristic (kernels.py:160-172), kernels.get_rbf_kernel (:52-65), kernels.get_funnel_kernel (:118-123),
    stein.phistar (stein.py:206-222), stein.get_phistar (:200-203), stein.stein_discrepancy (:99-106),
    stein.ksd_squared / _u / _v (:273-349), utils.l2_norm_squared (utils.py:483-487),
    metrics.get_mmd / get_funnel_tracer (metrics.py:88-127),
    models.KernelGradient.gradient (models.py:662-689), models.Particles.step (:143-176), flows.svgd_flow (flows.py:70-110).
Two passes: `f32` (faithful: float32 everywhere as in JAX with x64 off) and `f64` (REFSHIM_WIDEN=1: the shim aliases
float32 to float64; the source is still unmodified).  This script cannot run on the GPU box (no /root/reference);
the fixtures it writes are committed.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"

CASES = [  # name, n, d, m, seed, logp kind, fixed bandwidth   (same inputs as make_golden.py)
    ("funnel_n100_d2", 100, 2, 40, 1, "funnel", 1.0),
    ("gauss_n64_d5", 64, 5, 33, 2, "gauss", 0.7),
    ("gauss_n30_d50", 30, 50, 17, 3, "gauss", 3.0),
    ("gauss_n72_d16", 72, 16, 25, 4, "gauss", 2.0),
]
MEDIAN_SHAPES = [(100, 2), (64, 5), (30, 50), (50, 50), (51, 50), (52, 50), (72, 16), (200, 50), (2, 3), (3, 1), (17, 17), (18, 17)]


def inputs(n, d, m, seed, kind):
    import numpy as np
    rng = np.random.default_rng(seed)
    if kind == "funnel":
        X = (rng.standard_normal((n, d)) * np.array([1.5] * (d - 1) + [1.0])).astype(np.float32)
        Y = (rng.standard_normal((m, d)) * 2.0).astype(np.float32)
        return X, Y, None, None
    A = rng.standard_normal((d, d))
    prec = A @ A.T / d + np.eye(d) * 0.5
    mean = rng.standard_normal(d) * 0.3
    X = (rng.standard_normal((n, d)) * 0.8 + 0.4).astype(np.float32)
    Y = (rng.standard_normal((m, d)) + 0.1).astype(np.float32)
    return X, Y, mean, prec


def worker(tag):
    import collections
    import collections.abc
    import warnings
    warnings.simplefilter("ignore", SyntaxWarning)
    collections.Mapping = collections.abc.Mapping          # nets.py:6 (`from collections import Mapping`, pre-3.10)
    sys.path[:0] = [os.path.join(HERE, "refshim"), REF]
    sys.dont_write_bytecode = True
    import numpy as onp
    import torch
    import jax
    import jax.numpy as jnp
    from jax import vmap, grad
    from nvgd.src import kernels, stein, utils, metrics, models, flows   # the reference, unmodified

    fdt = jnp.float32
    T = lambda a: torch.as_tensor(onp.asarray(a)).to(fdt)
    N = lambda t: t.detach().numpy() if isinstance(t, torch.Tensor) else onp.asarray(t)

    def funnel_logp(x):                       # distributions.py:311-319 (Funnel.logpdf up to a constant)
        xs, y = x[:-1], x[-1]
        dm1 = xs.shape[0]
        return -y ** 2 / 18.0 - 0.5 * dm1 * y - 0.5 * jnp.exp(-y) * jnp.sum(xs ** 2)

    def gaussian_logp(mean, prec):
        mean, prec = T(mean), T(prec)
        return lambda x: -0.5 * jnp.dot(x - mean, jnp.dot(prec, x - mean))

    # ---- A2: the literal median heuristic, all regimes (n>d+1, n=d, n=d+1, n<d, tiny) ----------------------------
    med = {}
    for (n, d) in MEDIAN_SHAPES:
        rng = onp.random.default_rng(1000 + 7 * n + d)
        X = (rng.standard_normal((n, d)) * 1.3 + 0.25).astype(onp.float32)
        h = kernels.median_heuristic(T(X))
        med[f"X_{n}_{d}"] = X
        med[f"h_{n}_{d}"] = N(h)
    X1 = onp.random.default_rng(5).standard_normal(40).astype(onp.float32)     # (n,) input: documented as allowed
    try:
        med["h_1d_40"] = N(kernels.median_heuristic(T(X1)))
        med["X_1d_40"] = X1
    except Exception as e:                                                     # record how the reference fails
        med["h_1d_40_error"] = onp.array(type(e).__name__)
    onp.savez_compressed(os.path.join(HERE, f"ref_median_{tag}.npz"), **med)
    print(tag, "median", {k: float(v) for k, v in med.items() if k.startswith("h_") and v.dtype.kind == "f"})

    # ---- A1, A3-A9 on the four golden clouds ---------------------------------------------------------------------
    for (name, n, d, m, seed, kind, hfix) in CASES:
        X, Y, mean, prec = inputs(n, d, m, seed, kind)
        logp = funnel_logp if kind == "funnel" else gaussian_logp(mean, prec)
        Xt, Yt = T(X), T(Y)
        out = {"X": X, "Y": Y}
        out["Sx"] = N(vmap(grad(logp))(Xt))
        out["Sy"] = N(vmap(grad(logp))(Yt))
        out["h_median"] = N(kernels.median_heuristic(Xt))
        for htag, h in (("hfix", hfix), ("hmed", float(out["h_median"]))):
            k = kernels.get_rbf_kernel(h)
            out[f"h_{htag}"] = onp.float64(h)
            out[f"K_{htag}"] = N(vmap(vmap(k, (0, None)), (None, 0))(Xt, Yt))          # [m, n]: K[b,a] = k(x_a, y_b)
            phi, aux = stein.phistar(Xt, Xt, logp, k)
            out[f"phi_{htag}"], out[f"aux_{htag}"] = N(phi), N(aux)
            phi_r, aux_r = stein.phistar(Yt, Xt, logp, k)                               # queries Y, particles X
            out[f"phi_rect_{htag}"], out[f"aux_rect_{htag}"] = N(phi_r), N(aux_r)
            f = stein.get_phistar(k, logp, Xt)
            out[f"phi_closure_{htag}"] = N(vmap(f)(Xt))
            out[f"sd_phistar_{htag}"] = N(stein.stein_discrepancy(Xt, logp, f))
            out[f"l2_{htag}"] = N(utils.l2_norm_squared(Xt, f))
            out[f"ksd_u_{htag}"] = N(stein.ksd_squared_u(Xt, logp, k))
            out[f"ksd_v_{htag}"] = N(stein.ksd_squared_v(Xt, logp, k, None, None))
            out[f"ksd_rect_{htag}"] = N(stein.ksd_squared(Xt, Yt, logp, k))
        out["mmd_rbf"] = N(metrics.get_mmd(kernels.get_rbf_kernel(1.))(Xt, Yt))
        out["mmd_default_tracer"] = N(metrics.get_mmd_tracer(Yt)(Xt)["mmd"])
        ft = metrics.get_funnel_tracer(Yt)(Xt)
        out["mmd_tracer_rbf"], out["mmd_funnel"] = N(ft["rbf_mmd"]), N(ft["funnel_mmd"])

        # A7: KernelGradient.gradient, the four (bandwidth, scaled) corners the experiments use
        for gtag, bw, scaled in (("med_scaled", None, True), ("med_plain", None, False),
                                 ("fix_scaled", hfix, True), ("fix_plain", hfix, False)):
            kg = models.KernelGradient(target_logp=logp, bandwidth=bw, scaled=scaled, lambda_reg=0.5)
            g, aux = kg.gradient(None, Xt, aux=True)
            out[f"grad_{gtag}"], out[f"grad_{gtag}_h"], out[f"grad_{gtag}_logp"] = N(g), N(aux["bandwidth"]), N(aux["logp"])

        # A8: Particles.step x3 (sgd and adam), median bandwidth, scaled — the funnel experiment's configuration
        for opt, lr in (("sgd", 5e-2), ("adam", 1e-2)):
            kg = models.KernelGradient(target_logp=logp, bandwidth=None, scaled=True, lambda_reg=0.5)
            p = models.Particles(key=jax.random.PRNGKey(0), gradient=kg.gradient, init_samples=Xt,
                                 learning_rate=lr, optimizer=opt, compute_metrics=metrics.get_mmd_tracer(Yt))
            for _ in range(3):
                p.step(None)
            out[f"step3_{opt}_X"] = N(p.particles)
            out[f"step3_{opt}_lr"] = onp.float64(lr)
            out[f"step3_{opt}_h"] = onp.asarray([N(b) for b in p.rundata["bandwidth"]])
            out[f"step3_{opt}_gnorm"] = onp.asarray([N(b) for b in p.rundata["global_grad_norm"]])
            out[f"step3_{opt}_mmd0"] = onp.float64(N(p.rundata["mmd"][0][1]))

        # flows.svgd_flow end to end (defaults: bandwidth=1., scaled=True, lr=1e-1, lambda_reg=1/2; flows.py:70-80)
        class _Target:
            logpdf = staticmethod(logp)
            @staticmethod
            def sample(k_, key):
                return Yt[:k_]
        class _Proposal:
            @staticmethod
            def sample(k_, key):
                return Xt[:k_]
        class _Setup:
            @staticmethod
            def get():
                return _Target, _Proposal
        _, parts, err = flows.svgd_flow(jax.random.PRNGKey(0), _Setup, n_particles=n, n_steps=4, catch_exceptions=False)
        assert err is None
        out["flow4_X"] = N(parts.particles)
        out["flow4_mmd0"] = onp.float64(N(parts.rundata["mmd"][0][1]))
        onp.savez_compressed(os.path.join(HERE, f"ref_{name}_{tag}.npz"), **out)
        print(tag, name, "h_median", float(out["h_median"]), "ksd_u", float(out["ksd_u_hfix"]), "mmd", float(out["mmd_rbf"]))


def worker_ard(tag):
    """vector (ARD) bandwidth: kernels.get_rbf_kernel(bandwidth of shape (d,)) through stein.phistar / get_phistar /
    KernelGradient (unscaled) / metrics.get_mmd -> tests/golden/ref_ard_<tag>.npz"""
    import collections
    import collections.abc
    import warnings
    warnings.simplefilter("ignore", SyntaxWarning)
    collections.Mapping = collections.abc.Mapping
    sys.path[:0] = [os.path.join(HERE, "refshim"), REF]
    sys.dont_write_bytecode = True
    import numpy as onp
    import torch
    import jax.numpy as jnp
    from jax import vmap, grad
    from nvgd.src import kernels, stein, metrics, models
    fdt = jnp.float32
    T = lambda a: torch.as_tensor(onp.asarray(a)).to(fdt)
    N = lambda t: t.detach().numpy() if isinstance(t, torch.Tensor) else onp.asarray(t)
    name, n, d, m, seed, kind, hfix = CASES[1]                      # gauss_n64_d5
    X, Y, mean, prec = inputs(n, d, m, seed, kind)
    mean_t, prec_t = T(mean), T(prec)
    logp = lambda x: -0.5 * jnp.dot(x - mean_t, jnp.dot(prec_t, x - mean_t))
    bw = onp.array([0.6, 1.1, 0.8, 1.7, 0.9], dtype=onp.float64)
    Xt, Yt = T(X), T(Y)
    k = kernels.get_rbf_kernel(T(bw))
    out = {"X": X, "Y": Y, "bandwidth": bw, "Sx": N(vmap(grad(logp))(Xt))}
    out["K"] = N(vmap(vmap(k, (0, None)), (None, 0))(Xt, Yt))
    phi, aux = stein.phistar(Xt, Xt, logp, k)
    out["phi"], out["aux"] = N(phi), N(aux)
    phi_r, aux_r = stein.phistar(Yt, Xt, logp, k)
    out["phi_rect"], out["aux_rect"] = N(phi_r), N(aux_r)
    out["phi_closure"] = N(vmap(stein.get_phistar(k, logp, Xt))(Xt))
    out["mmd"] = N(metrics.get_mmd(k)(Xt, Yt))
    # (models.KernelGradient cannot take the vector: `self.bandwidth if self.bandwidth else ...` (models.py:657) raises on
    #  an array with more than one element, in JAX as under the stand-in)
    # leader / follower U-statistic (stein.phistar_u, stein.py:225-270), scalar bandwidth
    l_ = 24
    ku = kernels.get_rbf_kernel(hfix)
    phi_u, aux_u = stein.phistar_u(Xt[l_:], Xt[:l_], logp, ku)       # followers, leaders
    out["u_leaders"] = onp.int64(l_)
    out["u_h"] = onp.float64(hfix)
    out["phi_u"], out["aux_u"] = N(phi_u), N(aux_u)
    onp.savez_compressed(os.path.join(HERE, f"ref_ard_{tag}.npz"), **out)
    print(tag, "ard", float(onp.abs(out["phi"]).max()), float(out["mmd"]), "phistar_u", out["phi_u"].shape)


def worker_cfg1(tag, n_steps=1000):
    """BASELINE configs[0]: SVGD on the 2-D funnel, n = 100, median-heuristic bandwidth, scaled field, sgd lr = 5e-2,
    MMD tracers every 10 steps (funnel_mmd_comparison.py:33,79) -- a LONG run of the reference's own Particles /
    KernelGradient / tracer code -> tests/golden/ref_cfg1_<tag>.npz (snapshots every 50 steps, bandwidth and MMD traces)."""
    import collections
    import collections.abc
    import warnings
    warnings.simplefilter("ignore", SyntaxWarning)
    collections.Mapping = collections.abc.Mapping
    sys.path[:0] = [os.path.join(HERE, "refshim"), REF]
    sys.dont_write_bytecode = True
    import numpy as onp
    import torch
    import jax
    import jax.numpy as jnp
    from nvgd.src import metrics, models
    fdt = jnp.float32
    T = lambda a: torch.as_tensor(onp.asarray(a)).to(fdt)
    N = lambda t: t.detach().numpy() if isinstance(t, torch.Tensor) else onp.asarray(t)

    def funnel_logp(x):
        xs, y = x[:-1], x[-1]
        return -y ** 2 / 18.0 - 0.5 * xs.shape[0] * y - 0.5 * jnp.exp(-y) * jnp.sum(xs ** 2)
    rng = onp.random.default_rng(2020)
    X0 = rng.standard_normal((100, 2)).astype(onp.float32)
    yy = rng.standard_normal((100, 1)) * 3.0
    target = onp.concatenate([rng.standard_normal((100, 1)) * onp.exp(yy / 2), yy], axis=1).astype(onp.float32)
    kg = models.KernelGradient(target_logp=funnel_logp, bandwidth=None, scaled=True, lambda_reg=0.5)
    p = models.Particles(key=jax.random.PRNGKey(0), gradient=kg.gradient, init_samples=T(X0), learning_rate=5e-2,
                         optimizer="sgd", compute_metrics=metrics.get_funnel_tracer(T(target)))
    snaps = {0: N(p.particles).copy()}
    for k in range(1, n_steps + 1):
        p.step(None)
        if k % 50 == 0:
            snaps[k] = N(p.particles).copy()
    out = {"X0": X0, "target": target, "lr": onp.float64(5e-2), "n_steps": onp.int64(n_steps),
           "snap_steps": onp.asarray(sorted(snaps)), "snaps": onp.stack([snaps[k] for k in sorted(snaps)]),
           "h": onp.asarray([N(b) for b in p.rundata["bandwidth"]]),
           "gnorm": onp.asarray([N(b) for b in p.rundata["global_grad_norm"]]),
           "mmd_steps": onp.asarray([st for st, _ in p.rundata["rbf_mmd"]]),
           "rbf_mmd": onp.asarray([N(v) for _, v in p.rundata["rbf_mmd"]]),
           "funnel_mmd": onp.asarray([N(v) for _, v in p.rundata["funnel_mmd"]])}
    onp.savez_compressed(os.path.join(HERE, f"ref_cfg1_{tag}.npz"), **out)
    print(tag, "cfg1", n_steps, "steps; h", float(out["h"][0]), "->", float(out["h"][-1]), "funnel mmd", float(out["funnel_mmd"][0]),
          "->", float(out["funnel_mmd"][-1]))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "cfg1":
        worker_cfg1(sys.argv[1])
    elif len(sys.argv) > 2 and sys.argv[2] == "ard":
        worker_ard(sys.argv[1])
    elif len(sys.argv) > 1:
        worker(sys.argv[1])
    else:
        for tag, widen in (("f32", "0"), ("f64", "1")):
            env = dict(os.environ, REFSHIM_WIDEN=widen, PYTHONDONTWRITEBYTECODE="1")
            subprocess.run([sys.executable, os.path.abspath(__file__), tag], check=True, env=env)
            subprocess.run([sys.executable, os.path.abspath(__file__), tag, "ard"], check=True, env=env)
            subprocess.run([sys.executable, os.path.abspath(__file__), tag, "cfg1"], check=True, env=env)
