#!/usr/bin/env python
"""bench.py — SVGD pairwise interactions/s (n^2 * d per step) at n=65536, d=128 (BASELINE.json configs[3]).

A "step" is one pass of the hot path over synthetic particles:
    [all-gather X,S when row-sharded] -> median-heuristic bandwidth (exact select, kernels.py:160-172)
    -> phi* + KSD/l2 statistics (stein.py:180-203, 99-106) -> scaled sgd update (models.py:143-167, 677-689)
i.e. what models.Particles.step does for KernelGradient(bandwidth=None, scaled=True).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (torchrun for N > 1)
    python bench.py --impl reference ...                      # CPU arm (the oracle port: JAX is not installable)

Prints ONE JSON line on rank 0.  Timing: CUDA events on the launching stream, barrier + synchronize on both sides,
max over ranks; particles move every step and an L2-sized buffer is rewritten between steps (config.l2_flush).
Both arms print the same `config` object.  The headline run uses the literal median mode (what kernels.py:169
executes); `offdiag_mode` holds a second, shorter run with the docstring's off-diagonal median (a 13x tighter bandwidth:
three bf16-split Gram passes instead of one fp16 pass).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "projects-2020-neural-svgd_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "svgd_pairwise_interactions_per_s"
UNIT = "n^2*d/s"
LR, LAMBDA_REG = 1e-3, 0.5

# Exact median-heuristic bandwidths of workload(n, d, seed=0), as produced by the CUDA select (bit-identical between
# the tensor-core filtered select and the 3-pass SIMT select: tests/test_gpu_fullsize.py).  The CPU arm uses the SAME
# bandwidth as the GPU arm; computing it on the CPU (2.1e9 pair distances) is out of reach and is stated as excluded.
KNOWN_BANDWIDTH = {(65536, 128, "literal"): 51.85246658325195, (65536, 128, "offdiag"): 3.3723738193511963}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=65536)
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--path", default="auto")
    ap.add_argument("--median-mode", default="literal", choices=["literal", "offdiag", "full"])
    ap.add_argument("--bandwidth", type=float, default=0.0, help="> 0: fixed bandwidth (skip the median heuristic)")
    ap.add_argument("--cpu-rows", type=int, default=8192, help="query rows of the bounded CPU sample (per step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-offdiag", action="store_true", help="skip the secondary off-diagonal-median run")
    ap.add_argument("--no-secondary", action="store_true", help="skip the timings of the other BASELINE configs")
    ap.add_argument("--no-hint", action="store_true",
                    help="do not warm-start the median bracket from the previous step (engine.median_heuristic(hint=))")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step as one CUDA graph (auto = on)")
    return ap.parse_args()


def workload(n, d, seed=0):
    """X ~ N(0, I) + 0.5, S = -X / diag(logspace(-2, 0, d)) scaled to O(1).
    Generated with NumPy on the host (identical on every rank / arm), then copied."""
    from oracle.svgd_oracle import gaussian_diag_workload
    X, S = gaussian_diag_workload(n, d, seed=seed, shift=0.5)
    return X, (S * np.float32(0.02)).astype(np.float32)


def config_of(args, world):
    """identical in both arms (ours / reference): the driver compares it"""
    return {"workload": f"svgd_step n={args.n} d={args.d} (BASELINE configs[3])", "n": args.n, "d": args.d,
            "bandwidth": (f"median heuristic ({args.median_mode}) every step" if args.bandwidth <= 0 else f"fixed {args.bandwidth}"),
            "median_warm_start": bool(args.bandwidth <= 0 and not args.no_hint),
            "scaled": True, "optimizer": "sgd", "lr": LR, "lambda_reg": LAMBDA_REG,
            "parallelism": f"row-sharded x{world}", "l2_flush": "192 MiB fill between steps (GPU arm)"}


class ClockSampler:
    """SM clock, power and clock-event (throttle) reasons DURING the timed region, sampled through NVML every 5 ms
    (nvidia-smi -lms cannot go below ~100 ms: the row-sharded timed regions are 30-70 ms long)."""
    BITS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, device_index=0, period_s=0.005):
        self.samples, self.period, self.stop_flag, self.thread, self.h, self.err = [], period_s, False, None, None, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            uuid = None
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            except Exception:
                uuid = None
            if uuid:
                try:
                    self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
                except Exception:
                    self.h = None
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:          # NVML unavailable: say so in the record
            self.err = str(e)[:120]

    def _reasons(self):
        nv = self.nv
        for name in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
            fn = getattr(nv, name, None)
            if fn is not None:
                try:
                    return int(fn(self.h))
                except Exception:
                    continue
        return 0

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                     nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0, self._reasons()))
            except Exception as e:
                self.err = str(e)[:120]
                return
            time.sleep(self.period)

    def start(self):
        if self.h is None:
            return
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples" + (f": {self.err}" if self.err else "")]}
        sm = [s[0] for s in self.samples]
        bits = 0
        for s in self.samples:
            bits |= s[2]
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": self.max_mhz,
                "power_w_max": float(max(s[1] for s in self.samples)), "samples": len(sm), "period_ms": 1e3 * self.period,
                "source": "NVML", "reasons": sorted(v for k, v in self.BITS.items() if bits & k)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"burst": p.get("bf16_tflops"), "sustained": p.get("bf16_tflops_sustained", p.get("bf16_tflops")),
                "hbm_gbs": p.get("hbm_gbs"), "source": "MEASURED_PEAKS.json"}
    return {"burst": 1650.0, "sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(n, d, m_local):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu --set full capture of
    the SHIPPED kernel (profiles/phi_tc_traffic.json); only valid for the configuration it was captured on."""
    path = os.path.join(ROOT, "profiles", "phi_tc_traffic.json")
    if n == 65536 and d == 128 and m_local == n and os.path.exists(path):
        t = json.load(open(path))
        return t["dram_read_bytes"] + t["dram_write_bytes"]
    return None


def host_gram_passes(Xh, h):
    """the kernel's device-side rule (phi_tc.cu::tcw_onepass), re-evaluated on the host for the report: ONE fp16 Gram pass
    iff the rigorous bound (2 N R + R^2) / h^2 on |delta ln k| is <= gram_eps, N = max|xc|, R = max|xc - fp16(xc)|."""
    try:
        eps = float(os.environ.get("NSVGD_TC_GRAM_EPS", "5e-5"))
        if not (eps > 0 and np.isfinite(h) and h > 0):
            return 3, None
        xc = (Xh - Xh.astype(np.float64).mean(0).astype(np.float32)).astype(np.float32)
        with np.errstate(over="ignore"):
            f = xc.astype(np.float16).astype(np.float32)
        f[np.abs(f) < 6.103515625e-05] = 0.0
        nmax = float(np.sqrt((xc.astype(np.float64) ** 2).sum(1).max()))
        r2 = float((((xc - f).astype(np.float64)) ** 2).sum(1).max())
        bound = (2.0 * nmax * np.sqrt(r2) + r2) / (h * h)
        return (1 if bound <= eps else 3), bound
    except Exception:
        return 3, None


def cpu_port_rate(X, S, h, rows, threads=None):
    """The oracle's fp32 blocked port (sgemm Gram -> exp -> sgemm) on torch CPU with every host core, for `rows` query
    rows against all n particles.  Returns (n^2*d-units per second, seconds, threads)."""
    from oracle.svgd_oracle import phistar_f32_blocked_torch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)          # torchrun exports OMP_NUM_THREADS=1: undo it for the CPU arm
    n, d = X.shape
    rows = min(rows, n)
    Xt, St = torch.as_tensor(X), torch.as_tensor(S)
    t0 = time.perf_counter()
    phistar_f32_blocked_torch(Xt[:rows], Xt, St, h, block=1024)
    dt = time.perf_counter() - t0
    return rows * n * d / dt, dt, torch.get_num_threads()


def cpu_bandwidth(args, X):
    """the bandwidth the CPU arm uses: the GPU arm's exact value for this workload (median selection itself is excluded
    from the CPU arm: 2.1e9 pair distances)"""
    if args.bandwidth > 0:
        return args.bandwidth, "fixed"
    key = (args.n, args.d, args.median_mode)
    if key in KNOWN_BANDWIDTH:
        return KNOWN_BANDWIDTH[key], "exact value of the GPU arm's median select for this workload (bench.KNOWN_BANDWIDTH)"
    from oracle.svgd_oracle import median_heuristic
    sub = X[: min(args.n, 2048)]
    return float(median_heuristic(sub, args.median_mode)), "oracle median of the first 2048 particles"


def secondary_workloads(dev):
    """The other BASELINE configs as end-to-end steps on one GPU (parity for each is in tests/; these are timings only,
    CUDA events, 3 warm-up + 10 timed steps each).  Reported under "secondary"; not part of `value`."""
    from nsvgd import distributions, engine, kernels, metrics, models, stein
    from oracle import svgd_oracle as so
    out = {}

    def timed(fn, steps=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps

    try:    # configs[1]: 50-D correlated Gaussian, n = 2048, KSD_U + MMD tracked every step (device-side scores)
        rng = np.random.default_rng(3)
        d, n, m = 50, 2048, 500
        Q, _ = np.linalg.qr(rng.standard_normal((d, d)))
        cov = Q.T @ np.diag(np.logspace(-2, 0, d)) @ Q
        tgt = distributions.Gaussian(np.zeros(d), cov)
        X = torch.tensor(rng.standard_normal((n, d)).astype(np.float32), device=dev)
        Y = torch.tensor((rng.standard_normal((m, d)) @ np.linalg.cholesky(cov).T).astype(np.float32), device=dev)
        kg = models.KernelGradient(target_logp=tgt.logpdf, bandwidth=None, scaled=True, lambda_reg=0.5)
        mmd = metrics.get_mmd(kernels.get_rbf_kernel(1.0))

        def step2():
            _, aux = kg.step_sgd(X, None, 1e-3)
            stein.ksd_squared_u(X, tgt.logpdf, kernels.get_rbf_kernel(aux["bandwidth"]))
            mmd(X, Y)
        ms = timed(step2)
        out["cfg2_gaussian_n2048_d50_ksdU_mmd_every_step"] = {"ms_per_step": ms, "n2d_per_s": n * n * d / (ms * 1e-3)}
    except Exception as e:
        out["cfg2_gaussian_n2048_d50_ksdU_mmd_every_step"] = {"error": str(e)[:200]}
    for F in (54, 55):      # configs[2]: logistic regression posterior, d = 55 (BASELINE) / 56 (reference), n = 8192
        try:
            P, Xb, yb, N = so.logreg_workload(8192, F, B=128, seed=F)
            tgt = distributions.LogisticRegressionPosterior(F, N)
            kg = models.KernelGradient(get_target_logp=lambda batch: tgt.get_minibatch_logp(*batch), scaled=True)
            Pd = torch.tensor(P, device=dev)
            batch = (torch.tensor(Xb, device=dev), torch.tensor(yb, device=dev))
            ms = timed(lambda: kg.step_sgd(Pd, batch, 1e-9))
            out[f"cfg3_logreg_n8192_d{F + 1}"] = {"ms_per_step": ms, "n2d_per_s": 8192.0 * 8192 * (F + 1) / (ms * 1e-3)}
        except Exception as e:
            out[f"cfg3_logreg_n8192_d{F + 1}"] = {"error": str(e)[:200]}
    try:    # configs[0]: 2-D funnel, n = 100, median bandwidth, scaled sgd, tracer every 10th step (graph-replayed loop)
        tgt, prop = distributions.funnel.get()
        g = torch.Generator().manual_seed(0)
        kg = models.KernelGradient(target_logp=tgt.logpdf, bandwidth=None, scaled=True, lambda_reg=0.5)
        parts = models.Particles(0, kg.gradient, prop.sample(100, g), learning_rate=5e-2, optimizer="sgd",
                                 compute_metrics=metrics.get_funnel_tracer(tgt.sample(100, g).to(dev)))
        parts.step_many(50)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        parts.step_many(1000)
        torch.cuda.synchronize()
        us = (time.perf_counter() - t0) * 1e3
        out["cfg1_funnel_n100_d2_1000_steps"] = {"us_per_step": us, "steps_per_s": 1e6 / us}
    except Exception as e:
        out["cfg1_funnel_n100_d2_1000_steps"] = {"error": str(e)[:200]}
    try:    # configs[4]: BNN weight space, d = 4096, n = 16384 (phi* + KSD/l2 statistics at a fixed bandwidth)
        rng = np.random.default_rng(5)
        n, d = 16384, 4096
        X = torch.tensor((rng.standard_normal((n, d)) * 0.15).astype(np.float32), device=dev)
        S = torch.tensor((rng.standard_normal((n, d)) * 0.5).astype(np.float32), device=dev)
        h = float(np.sqrt(0.5 * 2 * d * 0.15 ** 2 / np.log(n + 1)))
        ms = timed(lambda: engine.phistar(X, X, S, h, Sq=S, stats=True), steps=5, warm=2)
        ms_med = timed(lambda: engine.median_heuristic(X), steps=3, warm=1)      # exact median bandwidth, literal mode (large-d filter)
        out["cfg5_bnn_n16384_d4096_phi"] = {"ms_per_call": ms, "n2d_per_s": float(n) * n * d / (ms * 1e-3),
                                            "algorithmic_tflops": 6.0 * n * n * d / (ms * 1e-3) / 1e12,
                                            "median_bandwidth_ms": ms_med, "step_ms": ms + ms_med,
                                            "step_n2d_per_s": float(n) * n * d / ((ms + ms_med) * 1e-3)}
    except Exception as e:
        out["cfg5_bnn_n16384_d4096_phi"] = {"error": str(e)[:200]}
    return out


def run_reference(args):
    """--impl reference: the reference's own implementation cannot be installed (JAX/jaxlib/haiku/optax absent, no
    network), so this arm times the CPU restatement (oracle port) on ALL host cores, each step a bounded sample
    (cpu-rows query rows x all n particles) of the same workload at the same bandwidth as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    X, S = workload(args.n, args.d)
    h, h_src = cpu_bandwidth(args, X)
    threads = os.cpu_count() or 1
    for _ in range(max(args.warmup, 1)):
        cpu_port_rate(X, S, h, min(args.cpu_rows, 1024), threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_rate(X, S, h, args.cpu_rows, threads)
    dt = time.perf_counter() - t0
    rows = min(args.cpu_rows, args.n)
    value = args.steps * rows * args.n * args.d / dt
    # what the reference's own algorithmic STRUCTURE costs (nested vmap / grad / jacfwd, O(n^2 d^2) for the scaled
    # field): one scaled-field evaluation of the line-for-line torch.func mirror at n=400, d=50 -- the
    # size of the reference's 1.02 s/it notebook anecdote
    structural = None
    try:
        from oracle import structural as st
        gq = torch.Generator().manual_seed(0)
        Xs = torch.randn(400, 50, generator=gq)
        logp = st.make_gaussian_logp(torch.zeros(50), torch.eye(50))
        t1 = time.perf_counter()
        st.get_field_scaled(Xs, logp, 3.0)
        ts = time.perf_counter() - t1
        structural = {"n": 400, "d": 50, "s_per_step": ts, "n2d_per_s": 400.0 * 400 * 50 / ts,
                      "what": "oracle/structural.py get_field_scaled (vmap/grad/jacfwd mirror of stein.py + models.py:677-689), torch CPU"}
    except Exception as e:      # informational only
        structural = {"error": str(e)[:200]}
    sample = (f"{rows} query rows x {args.n} particles per step, phi* only at h={h:.6g} ({h_src}); the median selection is "
              f"EXCLUDED from the CPU arm; torch CPU fp32 (MKL sgemm + vectorised exp), {threads} threads")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(args, world),
        "arm": {"note": "reference JAX unavailable in image (jax/jaxlib/haiku/optax absent, no network): CPU restatement "
                        "(oracle port) on all host cores, bounded sample per step", "bandwidth_used": h},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "structural_restatement": structural,
    }
    print(json.dumps(out))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from nsvgd import _lib, engine
    from nsvgd.parallel import GraphedStep, ShardedSVGD, row_partition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs CUDA; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, d, K, W = args.n, args.d, args.steps, args.warmup
    lr, lam = LR, LAMBDA_REG
    fixed_h = args.bandwidth if args.bandwidth > 0 else None

    Xh, Sh = workload(n, d)
    r0, r1 = row_partition(n, world, rank)
    m_local = r1 - r0
    X0 = torch.tensor(Xh[r0:r1], device=dev)
    S_loc = torch.tensor(Sh[r0:r1], device=dev)
    X_loc = X0.clone()
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    # auto: the step is replayed as ONE CUDA graph at every N (round 2: also at N = 1, where it removes the launch gaps between
    # the ~15 short kernels of the bandwidth selection: 6.28 -> 6.17 ms on the same box)
    use_graph = args.graph in ("on", "auto")
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_mode(mode, K, W, want_clocks):
        """W warm-up + K timed steps of the full SVGD step with the median in `mode`.  Returns a dict."""
        sharded = ShardedSVGD(n, d, median_mode=mode) if world > 1 else None
        X_loc.copy_(X0)
        state = {"graphed": None}
        # the SVGD loop keeps the warm-start state of the bandwidth selection between steps (bit-identical h; the sampling and
        # bracket-select phases are skipped when the previous step's order statistics still bracket the target ranks)
        if args.no_hint:
            os.environ["NSVGD_MEDIAN_HINT"] = "0"
        hint = None if (args.no_hint or world > 1) else engine.new_median_hint(dev)
        state["hint"] = hint

        def one_step():
            if state["graphed"] is not None:
                return state["graphed"].replay()
            if world > 1:
                h, aux, _ = sharded.step(X_loc, S_loc, lr, scaled=True, lambda_reg=lam, bandwidth=fixed_h)
            else:
                h = fixed_h if fixed_h else engine.median_heuristic(X_loc, mode, hint=hint)
                aux, _ = engine.svgd_step_sgd(X_loc, S_loc, h, lr, scaled=True, lambda_reg=lam, path=args.path)
            return h, aux

        launches_per_replay = 0
        if use_graph:
            l0 = _lib.launch_count()
            if world > 1:
                state["graphed"] = sharded.capture(X_loc, S_loc, lr, scaled=True, lambda_reg=lam, bandwidth=fixed_h)
            else:
                one_step(); torch.cuda.synchronize()
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_):
                    h_, aux_ = one_step()
                state["graphed"] = GraphedStep(g_, h_, aux_)
            launches_per_replay = (_lib.launch_count() - l0) // 2       # warm-up pass + captured pass
            X_loc.copy_(X0)
        for _ in range(W):
            flush.fill_(1)
            one_step()
        barrier()
        sampler = ClockSampler(local_rank) if (want_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        graphed = state["graphed"]
        _lib.profile_enable(graphed is None)      # library-side events cannot be recorded inside a replayed graph
        launches0 = _lib.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        t_wall0 = time.perf_counter()
        for i in range(K):
            flush.fill_(i & 1)
            ev[i][0].record()
            h, aux = one_step()
            ev[i][1].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
        launches = _lib.launch_count() - launches0 if graphed is None else launches_per_replay * K
        clocks = sampler.stop() if sampler else None
        step_ms = sum(a.elapsed_time(b) for a, b in ev)
        h_val = float(h) if not isinstance(h, torch.Tensor) else float(h.item())
        aux_host = aux.cpu().numpy()
        hint_dev = hint if hint is not None else (getattr(sharded, "_hint", None) if sharded is not None else None)
        hint_stats = engine.median_hint_stats(hint_dev) if hint_dev is not None else None     # after warm-up + the timed steps
        if graphed is not None:
            # per-kernel breakdown / roofline: a second, un-graphed pass of min(K, 5) steps with the library-side
            # CUDA events on (not part of `value`; says so in the notes)
            state["graphed"] = None
            _lib.profile_enable(True)
            KP = max(1, min(K, 5))
            for i in range(KP):
                flush.fill_(i & 1)
                one_step()
            barrier()
            state["graphed"] = graphed
        else:
            KP = K
        prof = {tag: _lib.profile_read(tag) for tag in range(15)}
        _lib.profile_enable(False)
        t = torch.tensor([step_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        return {"one_step": one_step, "state": state, "sharded": sharded, "total_ms": total_ms, "K": K, "KP": KP, "prof": prof,
                "h": h_val, "aux": aux_host, "launches": int(launches), "clocks": clocks, "t_wall": t_wall, "graphed": graphed,
                "hint_stats": hint_stats}

    def kernel_roofline(res, mode):
        ms, cnt = res["prof"][0]
        if not cnt:
            return None
        avg_ms = ms / cnt
        gram_passes, bound = host_gram_passes(Xh, res["h"]) if d > 32 else (0, None)
        flop = 6.0 * m_local * n * d            # algorithmic: Gram 2 m n d + K [S|X] 4 m n d   
        achieved = flop / (avg_ms * 1e-3) / 1e12
        timed_s = res["total_ms"] * 1e-3
        use_burst = timed_s < 1.0               # a kernel timed alone / in a short region holds burst clocks
        peak = pk["burst"] if use_burst else pk["sustained"]
        exec_units = (2 * gram_passes + 6 * (d + 16.0) / d) if d > 32 else 6.0      # N = d + 16 columns in the second GEMM
        return {"bound": "tensor", "kernel": "phi_tcw_kernel" if d > 32 else "phi_smalld_kernel", "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "frac_burst": achieved / pk["burst"],
                "frac_sustained": achieved / pk["sustained"], "peak_kind": "burst" if use_burst else "sustained",
                "traffic": ncu_traffic(n, d, m_local) if mode == "literal" else None, "avg_kernel_ms": avg_ms, "launches": cnt,
                "gram_passes": gram_passes, "gram_error_bound": bound,
                "executed_mma_frac_burst": exec_units / 6.0 * achieved / pk["burst"],
                "note": "achieved = algorithmic 6*m*n*d flop (SURVEY 8d) / CUDA-event time of the kernel; the kernel executes "
                        "(2*gram_passes + 6*(d+16)/d)*m*n*d MMA flop (second GEMM against [S - Xc/h^2 | 1 | c], bf16 2-piece split, "
                        "3 passes; Gram: ONE fp16 pass when the rigorous bound (2NR+R^2)/h^2 <= 5e-5, else three bf16-split passes); "
                        "peak = " + pk["source"] + (" burst (timed region %.2f s < 1 s)" % timed_s if use_burst else " sustained")
                        + ("" if res["graphed"] is None else "; kernel events from an un-graphed pass after the graph-replayed timed region")}

    def step_roofline(res):
        ms = res["total_ms"] / res["K"]
        flop = 7.0 * float(n) * n * d / world          # per GPU: phi* 6 + one upper-triangular Gram for the median
        ach = flop / (ms * 1e-3) / 1e12
        return {"achieved": ach, "unit": "TFLOP/s per GPU", "peak": pk["burst"], "frac_burst": ach / pk["burst"],
                "frac_sustained": ach / pk["sustained"], "what": "algorithmic 7*n^2*d/N flop per GPU / ms_per_step"}

    res = run_mode(args.median_mode, K, W, want_clocks=True)
    total_ms = res["total_ms"]
    value = K * float(n) * n * d / (total_ms * 1e-3)
    one_step, graphed, sharded = res["one_step"], res["graphed"], res["sharded"]

    # ---- e2e: the reference-facing call with HOST buffers, copies inside the timed region -------------
    e2e = None
    if not args.no_e2e:
        Xp = torch.tensor(Xh[r0:r1]).pin_memory(); Sp = torch.tensor(Sh[r0:r1]).pin_memory()
        out_host = torch.empty_like(Xp).pin_memory()
        aux_hostbuf = torch.empty(8, dtype=torch.float64).pin_memory()

        up = torch.cuda.Stream(device=dev)
        d2h = torch.cuda.Stream(device=dev)
        snap_ready, d2h_done = torch.cuda.Event(), torch.cuda.Event()
        out_dev = torch.empty_like(X_loc)
        aux_dev = torch.empty(8, dtype=torch.float64, device=dev)
        d2h_done.record()
        # Inputs are double-buffered: the host->device copy of step t + 1 (X and S, pinned host memory -> staging buffers, its own
        # stream) runs under the compute of step t, like the device->host copy of the logged particles; every copy is inside the
        # timed region (K uploads, K downloads for K steps) and each step consumes exactly the bytes uploaded for it.
        stage = [(torch.empty_like(X_loc), torch.empty_like(S_loc)) for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]
        for ev in freed:
            ev.record()
        tick = [0]

        def upload(i):
            with torch.cuda.stream(up):
                up.wait_event(freed[i])
                stage[i][0].copy_(Xp, non_blocking=True)
                stage[i][1].copy_(Sp, non_blocking=True)
                ready[i].record(up)
        # Row-sharded runs keep the upload in front of the step: at 8 ranks the prefetch stream next to the GRAPH-replayed step stalls
        # for seconds (measured e2e 3.4e11, wherever the prefetch is issued; fine at N <= 4 and fine with eager launches: 2.65e14
        # at N = 8 against 2.50e14 for graph + serial uploads) -- a graph-replay / copy-stream interaction that is not understood yet.
        # NSVGD_BENCH_PREFETCH=1|late re-enables it for experiments.
        prefetch = world == 1 or os.environ.get("NSVGD_BENCH_PREFETCH", "0") != "0"
        late = os.environ.get("NSVGD_BENCH_PREFETCH", "0") == "late"      # experiment: issue the prefetch after the step's launch
        if prefetch:
            upload(0)

        def e2e_step():
            cur = torch.cuda.current_stream()
            if prefetch:
                i = tick[0] & 1
                tick[0] += 1
                if not late:
                    upload(i ^ 1)                                # prefetch the next step's inputs
                cur.wait_event(ready[i])
                X_loc.copy_(stage[i][0], non_blocking=True)      # device-to-device into the step's buffers
                S_loc.copy_(stage[i][1], non_blocking=True)
                freed[i].record(cur)
            else:
                X_loc.copy_(Xp, non_blocking=True)
                S_loc.copy_(Sp, non_blocking=True)
            if world == 1 and graphed is None and not fixed_h:
                h_ = engine.median_heuristic(X_loc, args.median_mode, hint=res["state"].get("hint"))
                a, _ = engine.svgd_step_sgd(X_loc, S_loc, h_, lr, scaled=True, lambda_reg=lam, path=args.path)
            else:
                _, a = one_step()
            if prefetch and late:
                upload(i ^ 1)
            # the reference logs the particle array every step (models.py:187-197): snapshot on the device, then the
            # device->host copy rides on its own stream under the next step's work (PCIe is full duplex); the timed
            # region ends only after the last copy has landed
            cur.wait_event(d2h_done)                         # the previous snapshot has left the device
            out_dev.copy_(X_loc, non_blocking=True)
            aux_dev.copy_(a, non_blocking=True)
            snap_ready.record(cur)
            with torch.cuda.stream(d2h):
                d2h.wait_event(snap_ready)
                out_host.copy_(out_dev, non_blocking=True)
                aux_hostbuf.copy_(aux_dev, non_blocking=True)
                d2h_done.record(d2h)
        e2e_step()
        barrier()
        KE = max(2, min(K, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(KE):
            e2e_step()
        torch.cuda.current_stream().wait_stream(d2h)         # every device->host copy is inside the timed region
        torch.cuda.current_stream().wait_stream(up)          # ... and so is the (prefetched) upload issued by the last step
        e1.record()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": KE * float(n) * n * d / (float(te.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(2 * m_local * d * 4), "d2h_bytes_per_step": int(m_local * d * 4 + 64),
               "steps": KE, "api": "nsvgd.engine.median_heuristic + svgd_step_sgd (C ABI nsvgd_median_h_f32 / nsvgd_step_sgd_f32), pinned host buffers; "
                      + ("uploads double-buffered (the copy for step t+1 overlaps step t)" if world == 1 else "uploads in front of each step")
                      + ", downloads on their own stream"}

    # ---- secondary: the docstring's off-diagonal median (north-star intent; three Gram passes) ---------------
    offdiag = None
    if not args.no_offdiag and not fixed_h and args.median_mode != "offdiag":
        if res["state"]["graphed"] is not None:
            res["state"]["graphed"] = None
            torch.cuda.synchronize()
        K2 = max(3, min(K, 10))
        r2 = run_mode("offdiag", K2, 3, want_clocks=False)
        rf = kernel_roofline(r2, "offdiag") if rank == 0 else None
        offdiag = {"median_mode": "offdiag", "steps": K2, "ms_per_step": r2["total_ms"] / K2,
                   "value": K2 * float(n) * n * d / (r2["total_ms"] * 1e-3), "unit": UNIT,
                   "result": {"h": r2["h"], "ksd": float(r2["aux"][0]), "l2": float(r2["aux"][1]), "alpha": float(r2["aux"][2])},
                   "phi_main_kernel_ms": (r2["prof"][0][0] / r2["prof"][0][1]) if r2["prof"][0][1] else None,
                   "median_op_ms": r2["prof"][3][0] / r2["KP"],
                   "roofline": rf, "roofline_step": step_roofline(r2)}
        if r2["state"]["graphed"] is not None:
            r2["state"]["graphed"] = None
        res_graph_live = False
    else:
        res_graph_live = graphed is not None

    def shutdown():
        """A CUDA graph that captured NCCL kernels must be released before the communicator goes away, and
        destroy_process_group() was observed to hang after graph replays (torch 2.11 / NCCL 2.28): synchronise,
        drop the graph, barrier, and leave without tearing the communicator down."""
        if world == 1:
            return
        torch.cuda.synchronize()
        if use_graph:
            res["state"]["graphed"] = None
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush(); sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()

    if rank != 0:
        shutdown()
        return

    roofline = kernel_roofline(res, args.median_mode)
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        h_cpu = res["h"] if np.isfinite(res["h"]) and res["h"] > 0 else 6.0
        rate, secs, cores = cpu_port_rate(Xh, Sh, h_cpu, args.cpu_rows)      # warm-up (thread pool, MKL)
        reps, tot_rows, tot_s = 0, 0, 0.0
        while tot_s < 8.0 and reps < 8:                                        # bounded: ~10 s of CPU work
            rate, secs, cores = cpu_port_rate(Xh, Sh, h_cpu, args.cpu_rows)
            reps += 1; tot_rows += min(args.cpu_rows, n); tot_s += secs
        rate = tot_rows * float(n) * d / tot_s
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{reps} x {min(args.cpu_rows, n)} query rows x {n} particles at h={h_cpu:.6g} (the GPU arm's), phi* only "
                                  f"(median selection EXCLUDED), torch CPU fp32 (MKL sgemm + vectorised exp), {cores} threads, {tot_s:.2f} s; "
                                  "reference JAX not installable in this image"}
    secondary = None
    if world == 1 and not args.no_secondary and n == 65536 and d == 128:
        try:
            secondary = secondary_workloads(dev)
        except Exception as e:      # informational only: never lose the headline line
            secondary = {"error": str(e)[:200]}
    KP = res["KP"]
    prof = res["prof"]
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 (tensor-core MMAs on bf16x2-split / fp16 operands, fp32 accumulate)" if d > 32 else "f32",
        "data": "synthetic", "config": config_of(args, world),
        "arm": {"rows_per_gpu": m_local, "path": args.path,
                "small_allreduces": ("peer memory over NVLink, fused into the consumer kernels (csrc/p2p.cuh)"
                                     if (sharded is not None and sharded.comm is not None) else ("nccl" if world > 1 else "none")),
                "launch": ("one CUDA graph replay per step" + (" (collectives captured)" if world > 1 else "")) if graphed is not None
                          else "eager (C-ABI calls per op)"},
        "breakdown_ms_per_step": {"median_op": prof[3][0] / KP, "median_tc_filter_kernel": prof[6][0] / KP,
                                  "median_sample_kernel": prof[4][0] / KP, "median_exact_list_kernel": prof[5][0] / KP,
                                  "median_prepare": prof[12][0] / KP, "median_prep_kernels": prof[8][0] / KP,
                                  "median_bracket_select": prof[9][0] / KP, "median_resolve_candidate_select": prof[10][0] / KP,
                                  "median_fallback_passes_(early_exit)": prof[11][0] / KP,
                                  "phi_prep": prof[13][0] / KP, "phi_epilogue": prof[14][0] / KP, "phi_op": prof[2][0] / KP,
                                  "phi_main_kernel": prof[0][0] / KP,
                                  "source": "library-side CUDA events, " + ("the timed region" if graphed is None else
                                            f"a separate un-graphed pass of {KP} steps after the timed region")},
        "phi_only_value": (KP * float(m_local) * n * d / (prof[2][0] * 1e-3)) if prof[2][0] else None,
        "result": {"h": res["h"], "ksd": float(res["aux"][0]), "l2": float(res["aux"][1]), "alpha": float(res["aux"][2])},
        "median_hint": (dict(res["hint_stats"], note="warm-started bracket of the exact bandwidth selection: hits / misses over warm-up + timed "
                             "steps (a miss runs the exact fallback); h is bit-identical with --no-hint") if res.get("hint_stats") else None),
        "wall_s": res["t_wall"], "gpu_launches": res["launches"], "clocks": res["clocks"], "roofline": roofline,
        "roofline_step": step_roofline(res), "offdiag_mode": offdiag, "secondary": secondary, "cpu_baseline": cpu_baseline, "e2e": e2e,
    }
    print(json.dumps(out), flush=True)
    shutdown()


if __name__ == "__main__":
    main()
