"""N3: vector (ARD) and full-matrix Gaussian kernels and the learnable nets.RBFKernel.  The CUDA path against
the line-for-line structural mirror of stein.py (oracle/structural.py: vmap / grad / jacfwd over the kernel CALLABLE --
it shares no algebra with the linear-map / per-coordinate closed forms used on the device).
Reference: kernels.py:52-65 (vector bandwidth), kernels.py:83-88 (get_multivariate_gaussian_kernel), nets.py:24-49."""
import numpy as np
import pytest
import torch

from oracle import structural as st

pytestmark = pytest.mark.gpu
TOL = 1e-4


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def _setup(n=60, d=4, seed=0):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(n, d, generator=g, dtype=torch.float64) * 1.2 + 0.3
    Y = torch.randn(n // 2, d, generator=g, dtype=torch.float64)
    prec = torch.linspace(0.5, 2.0, d, dtype=torch.float64)
    logp64 = lambda x: -0.5 * torch.sum(prec * x * x)
    logp32 = lambda x: -0.5 * torch.sum(prec.to(x.device, x.dtype) * x * x)
    return X, Y, logp64, logp32


def full_kernel64(sigma):
    def k(x, y):
        s = sigma @ (x - y)
        return torch.exp(-torch.sum(s * s) / 2)
    return k


def test_full_matrix_gaussian_kernel_phistar_and_mmd():
    from nsvgd import kernels, metrics, stein, utils
    X, Y, logp64, logp32 = _setup()
    g = torch.Generator().manual_seed(3)
    sigma = torch.eye(4, dtype=torch.float64) * 0.7 + 0.2 * torch.randn(4, 4, generator=g, dtype=torch.float64)
    k64 = full_kernel64(sigma)
    k = kernels.get_multivariate_gaussian_kernel(sigma.float())
    assert float(k(X[0].float(), X[1].float())) == pytest.approx(float(k64(X[0], X[1])), rel=1e-5)
    ref = torch.func.vmap(st.get_phistar(k64, logp64, X))(Y)
    got = utils.vmap(stein.get_phistar(k, logp32, X.float().cuda()))(Y.float().cuda())
    assert rel(got.cpu(), ref) < TOL
    ref_phi, ref_aux = st.phistar(Y, X, logp64, k64)
    phi, aux = stein.phistar(Y.float().cuda(), X.float().cuda(), logp32, k)
    assert rel(phi.cpu(), ref_phi) < TOL and rel(aux.cpu(), ref_aux) < TOL
    mmd_ref = st.get_mmd(k64)(X, Y)
    assert float(metrics.get_mmd(k)(X.float().cuda(), Y.float().cuda())) == pytest.approx(float(mmd_ref), abs=2e-6)
    with pytest.raises(ValueError):
        kernels.get_multivariate_gaussian_kernel(torch.ones(3))


@pytest.mark.parametrize("variant", ["u", "v", "rect"])
def test_ksd_under_an_ard_bandwidth(variant):
    from nsvgd import kernels, stein
    X, Y, logp64, logp32 = _setup(n=50, d=4, seed=1)
    hvec = torch.tensor([0.8, 1.5, 2.2, 1.1], dtype=torch.float64)
    k64 = st.get_rbf_kernel(hvec)
    k = kernels.get_rbf_kernel(hvec.float())
    Xd, Yd = X.float().cuda(), Y.float().cuda()
    if variant == "u":
        ref, got = st.ksd_squared_u(X, logp64, k64), stein.ksd_squared_u(Xd, logp32, k)
    elif variant == "v":
        ref, got = st.ksd_squared_v(X, logp64, k64), stein.ksd_squared_v(Xd, logp32, k)
    else:
        ref, got = st.ksd_squared(X, Y, logp64, k64), stein.ksd_squared(Xd, Yd, logp32, k)
    scale = abs(float(st.ksd_squared_v(X, logp64, k64)))
    assert float(got) == pytest.approx(float(ref), abs=TOL * scale)


def test_ard_phistar_through_device_maps_matches_structural():
    from nsvgd import kernels, stein, utils
    X, Y, logp64, logp32 = _setup(n=70, d=5, seed=2)
    hvec = torch.tensor([0.8, 1.5, 2.2, 1.1, 0.6], dtype=torch.float64)
    ref = torch.func.vmap(st.get_phistar(st.get_rbf_kernel(hvec), logp64, X))(Y)
    got = utils.vmap(stein.get_phistar(kernels.get_rbf_kernel(hvec.float()), logp32, X.float().cuda()))(Y.float().cuda())
    assert rel(got.cpu(), ref) < TOL


@pytest.mark.parametrize("param", ["diagonal", "log_diagonal", "full"])
def test_learnable_rbf_kernel_module_dispatches_to_the_engine(param):
    """nets.RBFKernel (nets.py:24-49): the three parametrisations coincide at initialisation; after a parameter change the
    module's own evaluation and the engine's phi* under module.kernel() agree with the structural mirror"""
    from nsvgd import nets, stein, utils
    X, Y, logp64, logp32 = _setup(n=40, d=3, seed=4)
    mod = nets.RBFKernel(parametrization=param, dim=3)
    xy = torch.stack([X[0], X[1]]).float()
    assert float(mod(xy)) == pytest.approx(float(st.get_rbf_kernel(torch.tensor(1.0))(X[0], X[1])), rel=1e-5)
    with torch.no_grad():
        for p in mod.parameters():
            p.add_(0.3 * torch.randn(p.shape, generator=torch.Generator().manual_seed(7)))
    k = mod.kernel()
    # float64 callable with the module's parameter values
    if param == "diagonal":
        b = mod.bandwidth.detach().double(); kc = lambda x, y: torch.exp(-torch.sum((x - y) ** 2 / b ** 2) / 2)
    elif param == "log_diagonal":
        b = torch.exp(torch.clamp(mod.log_bandwidth.detach().double(), -5, 5)); kc = lambda x, y: torch.exp(-torch.sum((x - y) ** 2 / b) / 2)
    else:
        kc = full_kernel64(mod.sigma.detach().double())
    assert float(mod(xy)) == pytest.approx(float(kc(X[0], X[1])), rel=1e-5)
    ref = torch.func.vmap(st.get_phistar(kc, logp64, X))(Y)
    got = utils.vmap(stein.get_phistar(k, logp32, X.float().cuda()))(Y.float().cuda())
    assert rel(got.cpu(), ref) < TOL
