"""N1: the differentiable forms of the operators.  The CUDA backward rules (csrc/vjp.cu, the d/dh accumulator of
the KSD pair kernel), bound through torch.autograd.Function (nsvgd/autodiff.py), against torch autograd in float64 of
(a) the line-for-line structural mirror of stein.py (oracle/structural.py: nested vmap / grad / jacfwd) and (b) the
closed form written in plain differentiable torch.  The reference differentiates exactly these quantities
(stein.py:99-106, 180-203; experiments/sd_maxing.py:71-75; the kernel-learning notebooks)."""
import numpy as np
import pytest
import torch

from oracle import structural as st

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def phi_torch(Q, X, S, h):
    """plain differentiable closed form (float64): phi_i = (1/n) sum_j k_ij [s_j + (q_i - x_j)/h^2]"""
    r2 = ((Q[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    K = torch.exp(-r2 / (2 * h * h))
    return (K @ S + (K.sum(1, keepdim=True) * Q - K @ X) / (h * h)) / X.shape[0]


@pytest.mark.parametrize("variant", ["u", "v", "rect"])
def test_ksd_gradient_wrt_bandwidth_vs_structural_autodiff(variant):
    from nsvgd import autodiff
    g = torch.Generator().manual_seed(0)
    n, m, d = 48, 31, 3
    X = torch.randn(n, d, generator=g, dtype=torch.float64) * 1.3 + 0.2
    Y = torch.randn(m, d, generator=g, dtype=torch.float64)
    prec = torch.tensor([0.5, 2.0, 1.0], dtype=torch.float64)
    logp = lambda x: -0.5 * torch.sum(prec * x * x)
    score = lambda Z: -Z * prec.to(Z.device, Z.dtype)

    def ksd_of_h(h):
        k = st.get_rbf_kernel(h)
        if variant == "u":
            return st.ksd_squared_u(X, logp, k)
        if variant == "v":
            return st.ksd_squared_v(X, logp, k)
        return st.ksd_squared(X, Y, logp, k)
    h0 = torch.tensor(1.7, dtype=torch.float64)
    val_ref = ksd_of_h(h0).item()
    dh_ref = torch.func.grad(ksd_of_h)(h0).item()

    h = torch.tensor(1.7, device="cuda", requires_grad=True)
    Xd, Yd = X.float().cuda(), Y.float().cuda()
    val = autodiff.ksd_squared(Xd, score(Xd), h, variant, Yd if variant == "rect" else None, score(Yd) if variant == "rect" else None)
    val.backward()
    assert val.item() == pytest.approx(val_ref, rel=2e-5, abs=2e-6)
    assert h.grad.item() == pytest.approx(dh_ref, rel=1e-4, abs=1e-6)


def test_stein_api_ksd_is_differentiable_in_the_bandwidth():
    """the drop-in entry point itself: stein.ksd_squared_u(xs, logp, get_rbf_kernel(h)) with h.requires_grad"""
    from nsvgd import kernels, stein
    g = torch.Generator().manual_seed(1)
    X = torch.randn(200, 5, generator=g)
    logp = lambda x: -0.5 * torch.sum(x * x)
    h = torch.tensor(0.9, device="cuda", requires_grad=True)
    val = stein.ksd_squared_u(X.cuda(), logp, kernels.get_rbf_kernel(h))
    (gh,) = torch.autograd.grad(val, h)
    X64 = X.double()
    f = lambda hh: st.ksd_squared_u(X64, lambda x: -0.5 * torch.sum(x * x), st.get_rbf_kernel(hh))
    assert gh.item() == pytest.approx(torch.func.grad(f)(torch.tensor(0.9, dtype=torch.float64)).item(), rel=2e-4)
    # a step of gradient ASCENT on the KSD in h (kernel learning): the value must increase
    with torch.no_grad():
        h2 = h + 0.05 * gh / gh.abs()
    assert stein.ksd_squared_u(X.cuda(), logp, kernels.get_rbf_kernel(h2.detach())).item() > val.item()


@pytest.mark.parametrize("m,n,d,same", [(50, 70, 3, False), (150, 150, 7, True), (200, 300, 50, False), (333, 257, 130, False)])
def test_phistar_vjp_vs_autograd(m, n, d, same):
    from nsvgd import autodiff
    g = torch.Generator().manual_seed(m + n + d)
    X = torch.randn(n, d, generator=g, dtype=torch.float64) * 0.8 + 0.1
    Q = X if same else torch.randn(m, d, generator=g, dtype=torch.float64) * 0.9
    S = torch.randn(n, d, generator=g, dtype=torch.float64) * 0.5
    G = torch.randn(Q.shape[0], d, generator=g, dtype=torch.float64)
    hv = float(np.sqrt(d)) * 0.8
    # float64 autograd reference
    Xr = X.clone().requires_grad_(True); Sr = S.clone().requires_grad_(True)
    hr = torch.tensor(hv, dtype=torch.float64, requires_grad=True)
    Qr = Xr if same else Q.clone().requires_grad_(True)
    (phi_torch(Qr, Xr, Sr, hr) * G).sum().backward()
    # CUDA rules
    Xd = X.float().cuda().requires_grad_(True); Sd = S.float().cuda().requires_grad_(True)
    hd = torch.tensor(hv, device="cuda", requires_grad=True)
    Qd = Xd if same else Q.float().cuda().requires_grad_(True)
    phi = autodiff.phistar(Qd, Xd, Sd, hd)
    assert rel(phi.detach().cpu(), phi_torch(Q, X, S, torch.tensor(hv, dtype=torch.float64))) < 1e-4
    (phi * G.float().cuda()).sum().backward()
    assert rel(Xd.grad.cpu(), Xr.grad) < 2e-4, rel(Xd.grad.cpu(), Xr.grad)
    assert rel(Sd.grad.cpu(), Sr.grad) < 2e-4
    if not same:
        assert rel(Qd.grad.cpu(), Qr.grad) < 2e-4
    assert hd.grad.item() == pytest.approx(hr.grad.item(), rel=5e-4, abs=1e-6 * abs(hr.grad.item()) + 1e-7)


def test_phistar_vjp_matches_structural_mirror_and_chunks_rows():
    """(a) against autodiff THROUGH the structural mirror's get_phistar (vmap / grad nest of stein.py:180-203);
    (b) more query rows than one K/C chunk holds (m * n > 2^24): the transposed products accumulate over chunks"""
    from nsvgd import autodiff
    g = torch.Generator().manual_seed(5)
    n, d = 40, 3
    X = torch.randn(n, d, generator=g, dtype=torch.float64)
    prec = torch.tensor([1.0, 0.3, 2.0], dtype=torch.float64)
    logp = lambda x: -0.5 * torch.sum(prec * x * x)
    G = torch.randn(n, d, generator=g, dtype=torch.float64)

    def loss(Xv, h):
        phi = st.get_phistar(st.get_rbf_kernel(h), logp, Xv)
        return (torch.func.vmap(phi)(Xv) * G).sum()
    h0 = torch.tensor(1.1, dtype=torch.float64)
    gx_ref, gh_ref = torch.func.grad(loss, argnums=(0, 1))(X, h0)
    Xd = X.float().cuda().requires_grad_(True)
    hd = torch.tensor(1.1, device="cuda", requires_grad=True)
    phi = autodiff.phistar(Xd, Xd, -Xd * prec.float().cuda(), hd)     # the scores depend on X: autograd chains through them
    (phi * G.float().cuda()).sum().backward()
    assert rel(Xd.grad.cpu(), gx_ref) < 2e-4
    assert hd.grad.item() == pytest.approx(gh_ref.item(), rel=5e-4)
    # (b)
    m, n, d = 5000, 4096, 16
    Q = torch.randn(m, d, generator=g); X = torch.randn(n, d, generator=g); S = torch.randn(n, d, generator=g) * 0.3
    Gm = torch.randn(m, d, generator=g)
    Xr = X.double().requires_grad_(True)
    (phi_torch(Q.double(), Xr, S.double(), torch.tensor(3.0, dtype=torch.float64)) * Gm.double()).sum().backward()
    Xd = X.cuda().requires_grad_(True)
    (autodiff.phistar(Q.cuda(), Xd, S.cuda(), 3.0) * Gm.cuda()).sum().backward()
    assert rel(Xd.grad.cpu(), Xr.grad) < 2e-4


def test_ksd_squared_l_kernel_vs_oracle():
    from nsvgd import kernels, stein
    from oracle import svgd_oracle as so
    rng = np.random.default_rng(0)
    X = rng.standard_normal((1001, 6)).astype(np.float32)
    logp = lambda x: -0.5 * torch.sum(x * x)
    val, sd = stein.ksd_squared_l(torch.tensor(X).cuda(), logp, kernels.get_rbf_kernel(1.3), return_stddev=True)
    xs, ys = X[:500].astype(np.float64), X[500:1000].astype(np.float64)
    M = so.ksd_matrix(xs, -xs, ys, -ys, 1.3)
    outs = np.diag(M)
    assert val.item() == pytest.approx(outs.mean(), rel=1e-4, abs=1e-6)
    assert sd.item() == pytest.approx(outs.std(ddof=1) / 500, rel=1e-3)
