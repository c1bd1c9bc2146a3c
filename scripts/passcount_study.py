"""CPU study for DESIGN.md section 8 item 1: how many Gram passes does phi* need?
Emulates the tensor-core arithmetic in NumPy (bf16 pieces by bit truncation-with-rounding, fp32 accumulation) for the
COMBINED form on a subsample of the cfg4 workload and reports the phi* error vs float64 for 1, 2 and 3 Gram passes, at the
literal-mode bandwidth (h^2 >> |xc|^2) and at the off-diagonal-mode bandwidth (h^2 ~ |xc|^2 / 20).
    python scripts/passcount_study.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import svgd_oracle as so


def bf16(x):
    """round-to-nearest-even float32 -> bfloat16, returned as float32"""
    u = np.asarray(x, np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) & 0xFFFF0000).astype(np.uint32).view(np.float32)


def split(x):
    b0 = bf16(x)
    return b0, bf16(np.asarray(x, np.float32) - b0)


def phi_emulated(X, S, h, gram_passes):
    n, d = X.shape
    mu = X.mean(0, dtype=np.float64).astype(np.float32)
    Xc = (X - mu).astype(np.float32)
    W = (S - Xc / np.float32(h * h)).astype(np.float32)
    x0, x1 = split(Xc)
    w0, w1 = split(W)
    G = x0 @ x0.T
    if gram_passes >= 2:
        G = G + x0 @ x1.T
    if gram_passes >= 3:
        G = G + x1 @ x0.T
    xn = (Xc * Xc).sum(1)
    cs = np.float32(0.72134752044448170368 / (h * h))
    ex = 2 * cs * G - cs * xn[:, None] - cs * xn[None, :]
    K = np.exp2(ex.astype(np.float32))
    p0, p1 = split(K)
    Kq = p0 + p1
    M = p0 @ w0 + p0 @ w1 + p1 @ w0
    return (M + Kq.sum(1, keepdims=True) * Xc / np.float32(h * h)) / n


if __name__ == "__main__":
    n, d = 4096, 128
    X, S = so.gaussian_diag_workload(n, d, seed=0)
    S = (S * 0.02).astype(np.float32)
    h_off = float(so.median_heuristic(X[:1500], "offdiag")) * (np.log(1501) / np.log(n + 1)) ** 0.5
    xc2 = ((X - X.mean(0)) ** 2).sum(1)
    for name, h in (("cfg4 literal-mode bandwidth at n = 65536 (bench.py: h = 51.85)", 51.85), ("intermediate", 9.5),
                    ("offdiag-mode bandwidth", h_off)):
        ref = so.phistar(X, X, S, h)
        print(f"{name}: h^2 = {h * h:.1f}, max|xc|^2 / h^2 = {xc2.max() / (h * h):.3f}, rigorous 1-pass kernel bound 2^-8 |x||y|/h^2 = "
              f"{2 ** -8 * xc2.max() / (h * h):.1e}")
        for gp in (1, 2, 3):
            got = phi_emulated(X, S, h, gp)
            err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
            row = np.max(np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1))
            print(f"   {gp} Gram pass(es): rel Frobenius error {err:.2e}, max row error {row:.2e}")


def phi_emulated_pw(X, S, h, variant):
    """second-GEMM variants with ONE Gram pass: 'full' p0w0+p0w1+p1w0, 'no_p1' p0w0+p0w1, 'no_w1' p0w0+p1w0, 'single' p0w0"""
    n, d = X.shape
    mu = X.mean(0, dtype=np.float64).astype(np.float32)
    Xc = (X - mu).astype(np.float32)
    W = (S - Xc / np.float32(h * h)).astype(np.float32)
    x0, _ = split(Xc)
    w0, w1 = split(W)
    G = x0 @ x0.T
    xn = (Xc * Xc).sum(1)
    cs = np.float32(0.72134752044448170368 / (h * h))
    K = np.exp2((2 * cs * G - cs * xn[:, None] - cs * xn[None, :]).astype(np.float32))
    p0, p1 = split(K)
    if variant == "full":
        M, Kq = p0 @ w0 + p0 @ w1 + p1 @ w0, p0 + p1
    elif variant == "no_p1":
        M, Kq = p0 @ w0 + p0 @ w1, p0          # rowsum must use the same quantised k
    elif variant == "no_w1":
        M, Kq = p0 @ w0 + p1 @ w0, p0 + p1
    else:
        M, Kq = p0 @ w0, p0
    return (M + Kq.sum(1, keepdims=True) * Xc / np.float32(h * h)) / n


def study_pw():
    n, d = 4096, 128
    X, S = so.gaussian_diag_workload(n, d, seed=0)
    S = (S * 0.02).astype(np.float32)
    S_coh = np.tile(S[:1] * 5 + 0.3, (n, 1)).astype(np.float32)          # coherent scores: every particle pushes the same way
    print("\nsecond GEMM (P.W) variants, one Gram pass:")
    for sname, SS in (("cfg4 scores", S), ("coherent scores", S_coh)):
        for h in (51.85, 9.5):
            ref = so.phistar(X, X, SS, h)
            out = []
            for v in ("full", "no_p1", "no_w1", "single"):
                got = phi_emulated_pw(X, SS, h, v)
                out.append(f"{v} {np.linalg.norm(got - ref) / np.linalg.norm(ref):.1e}")
            print(f"   {sname}, h = {h}: " + ", ".join(out))


if __name__ == "__main__":
    study_pw()
