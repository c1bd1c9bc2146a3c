"""Generate tests/golden/*.npz.  Run from the repo root:  python tests/golden/make_golden.py

The reference cannot be imported in this image (no JAX), so the golden vectors
come from oracle/structural.py — a torch.func vmap/grad/jacfwd mirror written
line-for-line after nvgd/src/stein.py — in fp64, plus the literal (O(n^3)-style)
median-heuristic emulation.  (The fixtures produced by executing the reference's own sources are
ref_*.npz, see make_reference_golden.py; these are the second, independent witness.)
Inputs are float32 (what the CUDA path consumes); all outputs float64 except h
(float32, the bit-exact contract, from oracle/svgd_oracle.py's defined-order path).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import structural as st      # noqa: E402
from oracle import svgd_oracle as so     # noqa: E402

torch.set_default_dtype(torch.float64)
OUT = os.path.dirname(os.path.abspath(__file__))


def case(name, n, d, m, seed, logp_kind, h_fixed):
    rng = np.random.default_rng(seed)
    if logp_kind == "funnel":
        X = (rng.standard_normal((n, d)) * np.array([1.5] * (d - 1) + [1.0])).astype(np.float32)
        Y = (rng.standard_normal((m, d)) * 2.0).astype(np.float32)
        logp = st.funnel_logp
    else:
        A = rng.standard_normal((d, d))
        prec = A @ A.T / d + np.eye(d) * 0.5
        mean = rng.standard_normal(d) * 0.3
        X = (rng.standard_normal((n, d)) * 0.8 + 0.4).astype(np.float32)
        Y = (rng.standard_normal((m, d)) + 0.1).astype(np.float32)
        logp = st.make_gaussian_logp(torch.tensor(mean), torch.tensor(prec))
    Xt, Yt = torch.tensor(X, dtype=torch.float64), torch.tensor(Y, dtype=torch.float64)
    score = torch.func.vmap(torch.func.grad(logp))
    Sx, Sy = score(Xt).numpy(), score(Yt).numpy()

    out = {"X": X, "Y": Y, "Sx": Sx.astype(np.float32), "Sy": Sy.astype(np.float32)}
    # scores as the CUDA path will see them (fp32-rounded) drive everything below
    Sx32 = torch.tensor(out["Sx"], dtype=torch.float64)

    # bandwidths: bit-exact contract values (fp32) for all three modes + brute-force literal (fp64)
    for mode in so.MEDIAN_MODES:
        out[f"h_{mode}"] = np.float32(so.median_heuristic(X, mode))
    out["h_literal_bruteforce"] = np.float64(so.median_heuristic_bruteforce_literal(X))

    for tag, h in (("hfix", h_fixed), ("hmed", float(out["h_offdiag"]))):
        k = st.get_rbf_kernel(h)
        # structural mirror needs logp; the fp32-rounded scores differ from grad(logp) by
        # <= 6e-8 relative, far below the 1e-4 parity tolerance.
        phi, aux = st.phistar(Xt, Xt, logp, k)
        out[f"phi_{tag}"] = phi.numpy()
        out[f"aux_{tag}"] = aux.numpy()
        phi_rect, _ = st.phistar(Yt, Xt, logp, k)
        out[f"phi_rect_{tag}"] = phi_rect.numpy()
        out[f"ksd_u_{tag}"] = st.ksd_squared_u(Xt, logp, k).item()
        out[f"ksd_v_{tag}"] = st.ksd_squared_v(Xt, logp, k).item()
        out[f"ksd_rect_{tag}"] = st.ksd_squared(Xt, Yt, logp, k).item()
        alpha, phivals, ksd, l2 = st.get_field_scaled(Xt, logp, h, lambda_reg=0.5)
        out[f"alpha_{tag}"] = alpha.item()
        out[f"sd_phistar_{tag}"] = ksd.item()
        out[f"l2_{tag}"] = l2.item()
        out[f"h_{tag}"] = np.float64(h)
    out["mmd_rbf"] = st.get_mmd(st.get_rbf_kernel(1.0))(Xt, Yt).item()
    out["mmd_funnel"] = st.get_mmd(st.get_funnel_kernel(1.0))(Xt, Yt).item()
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)
    print(name, {k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in out.items()
                 if not hasattr(v, "shape") or v.shape == ()})


if __name__ == "__main__":
    case("funnel_n100_d2", n=100, d=2, m=40, seed=1, logp_kind="funnel", h_fixed=1.0)
    case("gauss_n64_d5", n=64, d=5, m=33, seed=2, logp_kind="gauss", h_fixed=0.7)
    case("gauss_n30_d50", n=30, d=50, m=17, seed=3, logp_kind="gauss", h_fixed=3.0)
    case("gauss_n72_d16", n=72, d=16, m=25, seed=4, logp_kind="gauss", h_fixed=2.0)
