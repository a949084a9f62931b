"""GPU: the pykeops.torch-compatible front-end (Genred / LazyTensor / KernelSolve) on the sm_100a engine.

Shapes, seeds and tolerances mirror the reference's own tests (all "KeOps vs dense torch"):
pykeops/pykeops/torch/test_install.py:5-25, test/test_lazytensor_gaussian_cpu.py:5-57,
test/test_torch.py:104-160,241-323,373-415,478-530,585-630, test/test_numpy.py:260-285."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

# The test BODIES below also run unchanged against the real reference on CPU tensors, which pins their expectations:
#   KT_MODULE=pykeops.torch KB200_TEST_DEVICE=cpu CXX=/usr/bin/g++ PYTHONPATH=/root/reference/keopscore:/root/reference/pykeops \
#       python -m pytest tests/test_torch_api_gpu.py -m gpu        (23 passed, 4 skipped in this container)
# Only the documented deviations are switched off there: this engine has no CPU backend and refuses host tensors.
import os
AGAINST_REFERENCE = os.environ.get("KT_MODULE", "").startswith("pykeops")


@pytest.fixture(scope="module")
def api(cuda_device):
    import importlib
    import os
    return importlib.import_module(os.environ.get("KT_MODULE", "keops_b200.torch"))


def test_install_known_answer(api, cuda_device):
    x = torch.arange(1, 10, dtype=torch.float32).view(-1, 3).to(cuda_device)
    y = torch.arange(3, 9, dtype=torch.float32).view(-1, 3).to(cuda_device)
    out = api.Genred("SqNorm2(x-y)", ["x=Vi(3)", "y=Vj(3)"], reduction_op="Sum", axis=0)(x, y)
    assert out.view(-1).tolist() == [63.0, 90.0]


@pytest.mark.skipif(AGAINST_REFERENCE, reason="the shim is this package's own entry point")
def test_pykeops_shim_runs_reference_style_user_code(cuda_device):
    import keops_b200
    keops_b200.install_as_pykeops()
    from pykeops.torch import LazyTensor  # noqa: the user's unchanged import line
    torch.manual_seed(0)
    x, y = torch.rand(50, 3, device=cuda_device), torch.rand(60, 3, device=cuda_device)
    D = ((LazyTensor(x[:, None, :]) - LazyTensor(y[None, :, :])) ** 2).sum(-1)
    idx = D.argKmin(3, dim=1)
    ref = ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1).argsort(dim=1)[:, :3]
    assert torch.equal(idx, ref)


@pytest.mark.parametrize("dtype,rtol", [(torch.float32, 1e-5), (torch.float64, 1e-12)])
def test_lazytensor_gaussian_conv_and_gradients(api, cuda_device, dtype, rtol):
    # test_lazytensor_gaussian_cpu.py: M=N=1000, D=3, seed 0, scaled inputs, forward + 1st + 2nd order grads
    torch.manual_seed(0)
    M, N, D, DV = 1000, 1000, 3, 1
    x = (torch.randn(M, D, dtype=dtype) / math.sqrt(D)).to(cuda_device)
    y = (torch.randn(N, D, dtype=dtype) / math.sqrt(D)).to(cuda_device)
    b = torch.randn(N, DV, dtype=dtype).to(cuda_device).requires_grad_(True)
    x.requires_grad_(True)

    def fun(x, y, b, lazy):
        if lazy:
            Dxy = ((api.LazyTensor(x[:, None, :]) - api.LazyTensor(y[None, :, :])) ** 2).sum(dim=2)
            return (-Dxy).exp() @ b
        Dxy = ((x[:, None, :] - y[None, :, :]) ** 2).sum(dim=2)
        return (-Dxy).exp() @ b

    outs = []
    for lazy in (True, False):
        out = fun(x, y, b, lazy)
        g_b, g_x = torch.autograd.grad((out ** 2).sum(), [b, x], create_graph=True)
        gg_b, = torch.autograd.grad((g_b ** 2).sum() + (g_x ** 2).sum(), [b], retain_graph=True)
        outs.append((out, g_b, g_x, gg_b))
    for got, ref in zip(*outs):
        scale = ref.abs().max()
        assert (got - ref).abs().max() <= 20 * rtol * scale, (got - ref).abs().max() / scale


@pytest.mark.parametrize("backend", ["auto", "GPU", "GPU_1D", "GPU_2D"])
def test_genred_sum_backends(api, cuda_device, backend):
    if AGAINST_REFERENCE and backend != "auto":
        pytest.skip("the reference's GPU backends need a GPU")
    torch.manual_seed(1)
    M, N, D = 700, 1900, 3
    x, y, b = torch.rand(M, D, device=cuda_device), torch.rand(N, D, device=cuda_device), torch.rand(N, 2, device=cuda_device)
    p = torch.tensor([0.7], device=cuda_device)
    conv = api.Genred("Exp(-p*SqDist(x,y))*b", ["x=Vi(3)", "y=Vj(3)", "b=Vj(2)", "p=Pm(1)"], reduction_op="Sum", axis=1)
    got = conv(x, y, b, p, backend=backend)
    ref = torch.exp(-0.7 * ((x[:, None] - y[None]) ** 2).sum(-1)) @ b
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5 * ref.abs().max().item())
    if not AGAINST_REFERENCE:                   # documented deviation: no CPU backend here
        with pytest.raises(ValueError):
            conv(x, y, b, p, backend="CPU")


def test_heterogeneous_aliases_and_out_inplace(api, cuda_device):
    torch.manual_seed(2)
    M, N = 300, 500
    x, y, b = torch.rand(M, 3, device=cuda_device), torch.rand(N, 3, device=cuda_device), torch.rand(N, 1, device=cuda_device)
    conv = api.Genred("Exp(-SqDist(x,y))*b", ["x=Vi(0,3)", "b=Vj(2,1)", "y=Vj(1,3)"], reduction_op="Sum", axis=1)
    out = torch.empty(M, 1, device=cuda_device)
    ret = conv(x, y, b, out=out)
    assert ret.data_ptr() == out.data_ptr()
    ref = torch.exp(-((x[:, None] - y[None]) ** 2).sum(-1)) @ b
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.float64, 1e-11)])
def test_logsumexp_and_gradients(api, cuda_device, dtype, tol):
    torch.manual_seed(3)
    M, N = 400, 600
    x = torch.rand(M, 3, dtype=dtype, device=cuda_device, requires_grad=True)
    y = torch.rand(N, 3, dtype=dtype, device=cuda_device)
    f = torch.rand(N, 1, dtype=dtype, device=cuda_device, requires_grad=True)
    eps = 0.05

    def softmin(lazy):
        if lazy:
            C = ((api.Vi(x) - api.Vj(y)) ** 2).sum(-1) / 2
            return (api.Vj(f) - C / eps).logsumexp(dim=1)
        C = ((x[:, None] - y[None]) ** 2).sum(-1) / 2
        return (f.view(1, -1) - C / eps).logsumexp(dim=1, keepdim=True)

    res = []
    for lazy in (True, False):
        s = softmin(lazy)
        gx, gf = torch.autograd.grad((s ** 2).sum(), [x, f])
        res.append((s, gx, gf))
    for got, ref in zip(*res):
        assert (got - ref).abs().max() <= tol * ref.abs().max(), (got - ref).abs().max() / ref.abs().max()


def test_sumsoftmaxweight_and_weighted_lse(api, cuda_device):
    torch.manual_seed(4)
    M, N = 200, 350
    x, y = torch.randn(M, 3, device=cuda_device), torch.randn(N, 3, device=cuda_device)
    w = torch.rand(N, 2, device=cuda_device) + 0.5
    S = (api.Vi(x) | api.Vj(y))
    got = S.sumsoftmaxweight(api.Vj(w), dim=1)
    ref = torch.softmax(x @ y.T, dim=1) @ w
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5)
    got2 = S.logsumexp(dim=1, weight=api.Vj(w))
    ref2 = torch.log(torch.exp(x @ y.T) @ w)
    assert torch.allclose(got2, ref2, rtol=1e-4, atol=1e-4)


def test_argkmin_exact_indices(api, cuda_device):
    # test_numpy.py:260-285 (ArgKMin K=3 vs argsort, exact)
    torch.manual_seed(5)
    M, N, D, K = 1500, 4000, 3, 10
    x, y = torch.randn(M, D, device=cuda_device), torch.randn(N, D, device=cuda_device)
    D_ij = ((api.Vi(x) - api.Vj(y)) ** 2).sum(-1)
    idx = D_ij.argKmin(K, dim=1)
    assert idx.dtype == torch.int64 and idx.shape == (M, K)
    d = ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)
    ref = d.topk(K, dim=1, largest=False).indices
    mism = idx != ref
    if mism.any():  # only genuine near-ties may differ
        dv = d.gather(1, idx)[mism] - d.gather(1, ref)[mism]
        assert dv.abs().max() <= 1e-6 and mism.float().mean() < 1e-3
    vals, idx2 = D_ij.Kmin_argKmin(K, dim=1)
    assert torch.equal(idx2, idx)
    assert torch.allclose(vals, d.gather(1, idx), rtol=1e-5, atol=1e-6)


def test_min_argmin_high_dim(api, cuda_device):
    # test_torch.py:585-630 uses D=990 to force the reference's chunked mode; here D=200 keeps the
    # fully unrolled kernel small
    torch.manual_seed(6)
    M, N, D = 100, 250, 200
    x, y = torch.rand(M, D, device=cuda_device), torch.rand(N, D, device=cuda_device)
    Dij = ((api.Vi(x) - api.Vj(y)) ** 2).sum(-1)
    vals, idx = Dij.min_argmin(dim=1)
    d = torch.cdist(x.double(), y.double()) ** 2
    rv, ri = d.min(dim=1)
    assert torch.equal(idx.view(-1), ri)
    assert torch.allclose(vals.view(-1).double(), rv, rtol=1e-5)
    with pytest.raises(NotImplementedError):
        xr = x.clone().requires_grad_(True)
        v = ((api.Vi(xr) - api.Vj(y)) ** 2).sum(-1).min(dim=1)
        v.sum().backward()


def test_kernel_solve_matches_dense_solve_and_grad(api, cuda_device):
    # test_torch.py:373-415
    torch.manual_seed(7)
    N, D, Dv = 300, 3, 2
    x = torch.rand(N, D, device=cuda_device, requires_grad=True)
    b = torch.rand(N, Dv, device=cuda_device)
    sigma = torch.tensor([0.5], device=cuda_device)
    alpha = 0.1
    Kinv = api.KernelSolve("Exp(-oos2*SqDist(x,y))*b", ["x=Vi(3)", "y=Vj(3)", "b=Vj(2)", "oos2=Pm(1)"], "b", axis=1)
    oos2 = 1 / sigma ** 2
    c = Kinv(x, x, b, oos2, alpha=alpha, eps=1e-7)
    Kxx = torch.exp(-oos2 * ((x[:, None] - x[None]) ** 2).sum(-1)) + alpha * torch.eye(N, device=cuda_device)
    cref = torch.linalg.solve(Kxx, b)
    assert torch.allclose(c, cref, atol=1e-4)
    u = torch.rand_like(c)
    g, = torch.autograd.grad(c, x, u)
    gref, = torch.autograd.grad(cref, x, u)
    assert torch.allclose(g, gref, atol=1e-3, rtol=1e-3)
    # LazyTensor.solve
    K_xx = (-(api.Vi(x.detach()) - api.Vj(x.detach())).sqnorm2() * oos2.item()).exp()
    c2 = K_xx.solve(api.Vi(b), alpha=alpha, eps=1e-7)
    assert torch.allclose(c2, cref, atol=1e-4)


def test_batch_dimensions_with_broadcast_and_grads(api, cuda_device):
    # test_torch.py:478-530 / test_lazytensor_gaussian_batch.py
    torch.manual_seed(8)
    B1, B2, M, N, D = 3, 2, 40, 70, 3
    x = torch.randn(B1, B2, M, 1, D, device=cuda_device, requires_grad=True)
    y = torch.randn(B1, 1, 1, N, D, device=cuda_device)
    b = torch.randn(1, B2, 1, N, 2, device=cuda_device, requires_grad=True)

    def fun(lazy):
        if lazy:
            K = (-((api.LazyTensor(x) - api.LazyTensor(y)) ** 2).sum(-1)).exp()
            return (K * api.LazyTensor(b)).sum(dim=3)
        K = (-((x - y) ** 2).sum(-1)).exp()
        return (K[..., None] * b).sum(dim=3)

    res = []
    for lazy in (True, False):
        o = fun(lazy)
        gx, gb = torch.autograd.grad((o ** 2).sum(), [x, b])
        res.append((o, gx, gb))
    for got, ref in zip(*res):
        assert got.shape == ref.shape
        assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4 * ref.abs().max().item())


def test_error_behaviour(api, cuda_device):
    x = torch.rand(10, 3, device=cuda_device)
    with pytest.raises(ValueError):
        api.Genred("SqDist(x,y)", ["x=Vi(3)", "y=Vj(3)"], axis=1)(x)            # wrong number of args
    if not AGAINST_REFERENCE:
        with pytest.raises(ValueError):
            api.Genred("SqDist(x,y)", ["x=Vi(3)", "y=Vj(3)"], axis=1)(x, x.cpu())   # mixed devices / no CPU path
    with pytest.raises(ValueError):
        api.Genred("SqDist(x,y)", ["x=Vi(3)", "y=Vj(4)"], axis=1)(x, x)         # wrong vector size
    with pytest.raises(ValueError):
        api.Genred("SqDist(x,y)", ["x=Vi(3)", "y=Vj(3)"], axis=1, sum_scheme="nope")


def test_lazytensor_transpose_matvec_rmatvec_sinc(api, cuda_device):
    """``K.t()``, ``K.T``, ``matvec`` / ``rmatvec`` (scipy LinearOperator protocol) and ``sinc``
    (pykeops/pykeops/common/lazy_tensor.py:2529-2639, 1267-1274; docstring example of ``t``)."""
    torch.manual_seed(3)
    dev = cuda_device
    x, y = torch.randn(300, 3, device=dev), torch.randn(200, 3, device=dev)
    x_i, y_j = api.LazyTensor(x[:, None, :]), api.LazyTensor(y[None, :, :])
    K = (-((x_i - y_j) ** 2).sum(2)).exp()
    Kd = (-((x[:, None, :] - y[None, :, :]) ** 2).sum(2)).exp()
    assert K.shape == (300, 200) and K.t().shape == (200, 300) and K.T.shape == (200, 300)
    w, v = torch.rand(300, 2, device=dev), torch.rand(200, 2, device=dev)
    torch.testing.assert_close(K.t() @ w, Kd.t() @ w, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(K.rmatvec(w), Kd.t() @ w, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(K.matvec(v), Kd @ v, rtol=1e-5, atol=1e-5)
    # K and K.t() in one formula: the same tensors appear as i- and j-variables
    xx_i, xx_j = api.LazyTensor(x[:, None, :]), api.LazyTensor(x[None, :, :])
    S = (-((xx_i - xx_j) ** 2).sum(2)).exp() * (xx_i | xx_j)
    Sd = (-((x[:, None, :] - x[None, :, :]) ** 2).sum(2)).exp() * (x @ x.t())
    torch.testing.assert_close((S + S.t()).sum(1).view(-1), (Sd + Sd.t()).sum(1), rtol=1e-4, atol=1e-4)
    s = ((x_i - y_j) ** 2).sum(2).sinc().sum(1)
    d2 = ((x[:, None, :] - y[None, :, :]) ** 2).sum(2)
    torch.testing.assert_close(s.view(-1), torch.sinc(d2).sum(1), rtol=1e-4, atol=1e-4)


def test_lazytensor_symbolic_grad(api, cuda_device):
    """``F.grad(v, e)``: formula-level ``Grad`` (lazy_tensor.py:1918-1936) against torch autograd."""
    torch.manual_seed(4)
    dev = cuda_device
    x = torch.randn(120, 3, device=dev, requires_grad=True)
    y, b = torch.randn(90, 3, device=dev), torch.randn(90, 1, device=dev)
    e = torch.randn(120, 1, device=dev)
    x_i, y_j, b_j, e_i = api.LazyTensor(x.detach()[:, None, :]), api.LazyTensor(y[None, :, :]), \
        api.LazyTensor(b[None, :, :]), api.LazyTensor(e[:, None, :])
    F = (-((x_i - y_j) ** 2).sum(2)).exp() * b_j
    g = F.grad(x_i, e_i).sum(1)                     # sum_j dF_ij/dx_i^T e_i
    dense = ((-((x[:, None, :] - y[None, :, :]) ** 2).sum(2)).exp() * b.view(1, -1)).sum(1, keepdim=True)
    (ref,) = torch.autograd.grad(dense, x, e)
    torch.testing.assert_close(g, ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("D,DV", [(300, 1), (3, 300)])
def test_high_dimensional_x_or_b_like_reference_chunk_tests(api, cuda_device, D, DV):
    """Mirrors pykeops/pykeops/test/test_chunks.py (D = 300: the reference's chunked mode) and test_finalchunks.py
    (dim b = 300: its final-chunk mode): same shapes, seed, scaling and tolerance."""
    M, N = 200, 300
    torch.manual_seed(0)
    x = torch.rand(M, 1, D, device=cuda_device) / math.sqrt(D)
    y = torch.rand(1, N, D, device=cuda_device) / math.sqrt(D)
    b = torch.randn(N, DV, device=cuda_device)
    Kxy = (-((api.LazyTensor(x) - api.LazyTensor(y)).square()).sum(dim=2)).exp()
    out = Kxy.__matmul__(b, sum_scheme="block_sum")
    ref = (-((x - y) ** 2).sum(dim=2)).exp() @ b
    assert torch.allclose(out, ref, atol=1e-4)


# ---- torch.func transforms (pykeops/pykeops/test/test_torch_func.py, test_vmap.py) ---------------------------------
def _func_inputs(dev):
    torch.manual_seed(0)
    B1, B2, M, N, D = 5, 4, 10, 20, 2
    mk = lambda *s: torch.randn(*s, device=dev)
    prim = (mk(B1, B2, M, 1, D), mk(B1, 1, 1, N, D), mk(B1, B2, 1, N, 1), torch.rand(B1, 1, 1, 1, 1, device=dev))
    tang = (mk(B1, B2, M, 1, D), mk(B1, 1, 1, N, D), mk(B1, B2, 1, N, 1), mk(B1, 1, 1, 1, 1))
    return prim, tang


def _fn_torch(x_i, y_j, b_j, p):
    K_ij = (-p[..., 0] * ((x_i - y_j) ** 2).sum(-1)).exp()[..., None]
    return ((K_ij * b_j).sum(2)).norm()


def _fn_keops(api):
    def fn(x_i, y_j, b_j, p):
        x_i, y_j, b_j, p = api.LazyTensor(x_i), api.LazyTensor(y_j), api.LazyTensor(b_j), api.LazyTensor(p)
        K_ij = (-p * ((x_i - y_j) ** 2).sum(-1)).exp()
        return ((K_ij * b_j).sum(2)).norm()
    return fn


def test_torch_func_jvp_grad_vjp(api, cuda_device):
    prim, tang = _func_inputs(cuda_device)
    fk = _fn_keops(api)
    v1, t1 = torch.func.jvp(_fn_torch, prim, tang)
    v2, t2 = torch.func.jvp(fk, prim, tang)
    assert torch.allclose(v1, v2, atol=1e-4) and torch.allclose(t1, t2, atol=1e-4)
    g1 = torch.func.grad(_fn_torch, (0, 1, 2, 3))(*prim)
    g2 = torch.func.grad(fk, (0, 1, 2, 3))(*prim)
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, atol=1e-4)
    r1 = torch.func.vjp(_fn_torch, *prim)[1](torch.tensor(1.0, device=cuda_device))[0]
    r2 = torch.func.vjp(fk, *prim)[1](torch.tensor(1.0, device=cuda_device))[0]
    assert torch.allclose(r1, r2, atol=1e-4)


def test_torch_func_vmap(api, cuda_device):
    prim, _ = _func_inputs(cuda_device)
    res1 = torch.func.vmap(_fn_torch)(*prim)
    res2 = torch.func.vmap(_fn_keops(api))(*prim)
    assert torch.allclose(res1, res2, atol=1e-4)

    def fn(x_i, x_j, y_j, use_keops):
        if use_keops:
            x_i, x_j, y_j = api.LazyTensor(x_i), api.LazyTensor(x_j), api.LazyTensor(y_j)
        K_ij = (-((x_i - x_j) ** 2).sum(-1)).exp()
        if not use_keops:
            K_ij = K_ij[..., None]
        return (K_ij * y_j).sum(1)

    torch.manual_seed(0)
    for in_dim, shapes in ((0, ((5, 10, 1, 2), (5, 1, 20, 2), (5, 1, 20, 1))), (1, ((10, 5, 1, 2), (1, 5, 20, 2), (1, 5, 20, 1)))):
        x_i = torch.randn(*shapes[0], device=cuda_device, dtype=torch.float64, requires_grad=True)
        x_j = torch.randn(*shapes[1], device=cuda_device, dtype=torch.float64)
        y_j = torch.randn(*shapes[2], device=cuda_device, dtype=torch.float64)
        rk = torch.vmap(lambda a, b, c: fn(a, b, c, True), in_dims=in_dim, out_dims=in_dim)(x_i, x_j, y_j)
        rt = torch.vmap(lambda a, b, c: fn(a, b, c, False), in_dims=in_dim, out_dims=in_dim)(x_i, x_j, y_j)
        assert torch.allclose(rk, rt)
        u = torch.rand(rt.shape, device=cuda_device, dtype=torch.float64)
        (gt,) = torch.autograd.grad(rt, x_i, u)
        (gk,) = torch.autograd.grad(rk, x_i, u)
        assert torch.allclose(gk, gt)


def test_lazytensor_diff_divergence_laplacian(api, cuda_device):
    """``x.diff(v, e)``, ``x.divergence(v)``, ``x.laplacian(v)``, ``x.factorize(y)`` (lazy_tensor.py:1938-2068) on the GPU."""
    torch.manual_seed(6)
    dev = cuda_device
    x, y = torch.randn(80, 3, device=dev, dtype=torch.float64), torch.randn(70, 3, device=dev, dtype=torch.float64)
    e = torch.randn(80, 3, device=dev, dtype=torch.float64)
    x_i, y_j, e_i = api.LazyTensor(x[:, None, :]), api.LazyTensor(y[None, :, :]), api.LazyTensor(e[:, None, :])
    K = (-((x_i - y_j) ** 2).sum(2)).exp()
    d2 = ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)
    lap = K.laplacian(x_i).sum(1).view(-1)
    torch.testing.assert_close(lap, ((4 * d2 - 6) * torch.exp(-d2)).sum(1), rtol=1e-10, atol=1e-10)
    dK = K.diff(x_i, e_i).sum(1).view(-1)                     # directional derivative of K along e
    ref = (-2 * ((x[:, None, :] - y[None, :, :]) * e[:, None, :]).sum(-1) * torch.exp(-d2)).sum(1)
    torch.testing.assert_close(dK, ref, rtol=1e-10, atol=1e-10)
    div = (x_i * K).divergence(x_i).sum(1).view(-1)
    xd = (x[:, None, :] * (x[:, None, :] - y[None, :, :])).sum(-1)
    torch.testing.assert_close(div, ((3 - 2 * xd) * torch.exp(-d2)).sum(1), rtol=1e-10, atol=1e-10)
    fac = (K + K).factorize(K).sum(1).view(-1)
    torch.testing.assert_close(fac, 2 * torch.exp(-d2).sum(1), rtol=1e-10, atol=1e-10)


def test_non_contiguous_inputs_and_lazytensor_inplace_out(api, cuda_device):
    """pykeops/pykeops/test/test_contiguous_torch.py (a permuted view of a batch variable gives the same result as its
    contiguous copy) and test_lazytensor_gaussian_inplace.py (``K.__matmul__(b, sum_scheme=..., out=a)`` writes into ``a``)."""
    torch.manual_seed(0)
    a1 = torch.rand(2, 1000, 5, device=cuda_device)
    a2 = a1.permute(2, 0, 1).contiguous().permute(1, 2, 0)
    assert not a2.is_contiguous() and torch.equal(a1, a2)
    b, c = torch.rand(2, 1000, 5, device=cuda_device), torch.rand(2, 1000, 5, device=cuda_device)
    b_j = api.LazyTensor(b[:, None])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        d1 = (-api.LazyTensor(a1[:, :, None]).sqdist(b_j)).exp() @ c
        d2 = (-api.LazyTensor(a2[:, :, None]).sqdist(b_j)).exp() @ c
    ref = torch.einsum("bmn,bnd->bmd", torch.exp(-((a1[:, :, None] - b[:, None]) ** 2).sum(-1)), c)
    assert d1.shape == (2, 1000, 5) and torch.equal(d1, d2)
    torch.testing.assert_close(d1, ref, rtol=1e-5, atol=1e-5 * ref.abs().max().item())

    M, N, D = 20, 30, 3
    x = torch.rand(M, 1, D, device=cuda_device) / math.sqrt(D)
    y = torch.rand(1, N, D, device=cuda_device) / math.sqrt(D)
    w = torch.randn(N, 1, device=cuda_device)
    a = torch.empty(M, 1, device=cuda_device)
    K = (-((api.LazyTensor(x) - api.LazyTensor(y)).square()).sum(dim=2)).exp()
    ret = K.__matmul__(w, sum_scheme="block_sum", out=a)
    assert ret.data_ptr() == a.data_ptr()
    torch.testing.assert_close(a, torch.exp(-((x - y) ** 2).sum(-1)) @ w, rtol=1e-5, atol=1e-5)


def test_lazytensor_clamp_and_broadcast_grads(api, cuda_device):
    """pykeops/pykeops/test/test_lazytensor_clamp.py (forward + gradient through ``clamp``) and test_lazytensor_grad.py
    (a dimension-1 j-variable broadcast against ``cos(x_i)``, first and second order gradients)."""
    torch.manual_seed(0)
    x = torch.randn(1000, 1, 3, requires_grad=True, device=cuda_device)
    y = torch.randn(1, 1000, 3, device=cuda_device)

    def fun(x, y, lazy):
        if lazy:
            x, y = api.LazyTensor(x), api.LazyTensor(y)
        Dxy = ((x * y).clamp(-1.23, 1.54)).sum(dim=2)
        return (-(Dxy ** 2)).exp().sum(dim=1)

    out = [fun(x, y, lazy).squeeze() for lazy in (False, True)]
    g = [torch.autograd.grad((o ** 2).sum(), [x])[0] for o in out]
    assert torch.allclose(out[0], out[1])
    assert torch.allclose(g[0], g[1], atol=0.01, rtol=0.001)

    x = torch.rand(100, 1, 3, requires_grad=True, device=cuda_device)
    y = torch.rand(1, 100, 1, device=cuda_device)
    b = torch.randn(100, 1, device=cuda_device)

    def fun2(x, y, b, lazy):
        if lazy:
            x, y = api.LazyTensor(x), api.LazyTensor(y)
        return ((x.cos() - y).sum(dim=2) @ b).squeeze()

    out = [fun2(x, y, b, lazy) for lazy in (True, False)]
    g1 = [torch.autograd.grad((o ** 2).sum(), [x], create_graph=True)[0] for o in out]
    g2 = [torch.autograd.grad((g ** 2).sum(), [x])[0] for g in g1]
    assert torch.allclose(out[0], out[1], rtol=1e-4)
    assert torch.allclose(g1[0], g1[1], rtol=1e-4)
    assert torch.allclose(g2[0], g2[1], rtol=1e-3)


def test_accumulator_options(api, cuda_device):
    """``dtype_acc="float64"`` / ``use_double_acc=True`` (fp32 data, fp64 accumulators) and ``sum_scheme="kahan_scheme"``
    (pykeops/pykeops/torch/generic/generic_red.py:437-520, test_torch.py:417-445): all agree with a float64 dense sum,
    the wide accumulators more closely than plain fp32 block sums."""
    torch.manual_seed(5)
    M, N = 200, 20000
    x, y = torch.rand(M, 3, device=cuda_device), torch.rand(N, 3, device=cuda_device)
    b = torch.rand(N, 1, device=cuda_device) + 1.0
    ref = (torch.exp(-((x.double()[:, None] - y.double()[None]) ** 2).sum(-1)) @ b.double())
    al = ["x=Vi(3)", "y=Vj(3)", "b=Vj(1)"]
    err = {}
    for name, kw in (("block", {}), ("direct", dict(sum_scheme="direct_sum")), ("kahan", dict(sum_scheme="kahan_scheme")),
                     ("acc64", dict(dtype_acc="float64")), ("double_acc", dict(use_double_acc=True))):
        out = api.Genred("Exp(-SqDist(x,y))*b", al, reduction_op="Sum", axis=1, **kw)(x, y, b)
        assert out.dtype == torch.float32 and out.shape == (M, 1)
        err[name] = ((out.double() - ref).abs() / ref.abs()).max().item()
    assert max(err.values()) < 1e-5, err
    assert err["acc64"] < 1e-6 and err["double_acc"] < 1e-6 and err["kahan"] < 1e-6, err
    with pytest.raises(ValueError):
        api.Genred("Exp(-SqDist(x,y))*b", al, reduction_op="Sum", axis=1, dtype_acc="float64", use_double_acc=True)
    with pytest.raises(ValueError):
        api.Genred("SqDist(x,y)", al[:2], reduction_op="Min", axis=1, dtype_acc="float64")
    # the same options act on the (m, s) accumulator of the LogSumExp reductions (reference: acc is declared in dtypeacc
    # for every reduction, MapReduce.py:74; Kahan: Max_SumShiftExpWeight_Reduction.py:65-85) -- plus the combination
    # kahan_scheme + dtype_acc="float64", which get_optional_flags accepts
    D2 = ((x.double()[:, None] - y.double()[None]) ** 2).sum(-1)
    ref_lse = torch.logsumexp(-D2, dim=1, keepdim=True)
    ref_w = (torch.softmax(-D2, dim=1) @ b.double())
    for name, kw in (("block", {}), ("direct", dict(sum_scheme="direct_sum")), ("kahan", dict(sum_scheme="kahan_scheme")),
                     ("acc64", dict(dtype_acc="float64")), ("kahan64", dict(sum_scheme="kahan_scheme", dtype_acc="float64")),
                     ("direct64", dict(sum_scheme="direct_sum", dtype_acc="float64"))):
        out = api.Genred("-SqDist(x,y)", al[:2], reduction_op="LogSumExp", axis=1, **kw)(x, y)
        e1 = ((out.double() - ref_lse).abs() / ref_lse.abs()).max().item()
        out = api.Genred("-SqDist(x,y)", al, reduction_op="SumSoftMaxWeight", axis=1, formula2="b", **kw)(x, y, b)
        e2 = ((out.double() - ref_w).abs() / ref_w.abs()).max().item()
        assert e1 < 1e-5 and e2 < 1e-5, (name, e1, e2)
        if name in ("acc64", "kahan64", "direct64", "kahan"):
            assert e2 < 2e-6, (name, e2)
    out = api.Genred("Exp(-SqDist(x,y))*b", al, reduction_op="Sum", axis=1, sum_scheme="kahan_scheme", dtype_acc="float64")(x, y, b)
    assert ((out.double() - ref).abs() / ref.abs()).max().item() < 1e-6


def test_tensordot_with_permute_and_gradients(api, cuda_device):
    # test_torch.py:680-730 (contraction over three paired axes, permuted kept axes, both gradients) and
    # test_lazytensor_tensordot.py (matrix-vector product as a contraction)
    import numpy as np
    torch.manual_seed(0)
    M, N = 40, 50
    x = torch.randn(M, 2, 3, 2, 2, 4, dtype=torch.float64, device=cuda_device, requires_grad=True)
    y = torch.randn(N, 2, 4, 2, 3, 2, 3, dtype=torch.float64, device=cuda_device, requires_grad=True)
    dimfa, dimfb = tuple(x.shape[1:]), tuple(y.shape[1:])
    contfa, contfb = [5, 1, 3], [2, 5, 3]
    perm = [4, 3, 2, 0, 1]
    perm_torch = (0,) + tuple(int(i) + 1 for i in np.argsort(perm))
    ref = torch.tensordot(x, y, dims=(contfa, contfb)).sum(3).permute(perm_torch)
    f = api.LazyTensor(x.reshape(M, 1, -1)).keops_tensordot(
        api.LazyTensor(y.reshape(1, N, -1)), dimfa, dimfb, tuple(c - 1 for c in contfa), tuple(c - 1 for c in contfb),
        tuple(perm))
    out = f.sum_reduction(dim=1)
    assert out.shape == (M, 108)
    assert torch.allclose(out.flatten(), ref.flatten(), rtol=1e-10, atol=1e-10)
    e = torch.randn_like(ref)
    for v in (x, y):
        gk = torch.autograd.grad(out, v, e.reshape(M, -1), retain_graph=True)[0]
        gt = torch.autograd.grad(ref, v, e, retain_graph=True)[0]
        assert torch.allclose(gk.flatten(), gt.flatten(), rtol=1e-9, atol=1e-9)
    a = torch.randn(4 * 7, dtype=torch.float64, device=cuda_device)
    b = torch.randn(7, dtype=torch.float64, device=cuda_device)
    C = api.LazyTensor(a[None, None, :]).keops_tensordot(api.LazyTensor(b[None, None, :]), (4, 7), (7,), (1,), (0,)).sum_reduction(dim=1)
    assert torch.allclose(C.flatten(), a.reshape(4, 7) @ b)


def test_kron(api, cuda_device):
    # test_kron.py: numpy.kron of the internal dimensions, through the LazyTensor method and through both Genred spellings
    torch.manual_seed(1)
    M, N = 11, 150
    for dimfa, dimfb in (([2, 3], [2, 2]), ([3, 1, 3, 2], [3, 3, 2, 1])):
        da, db = math.prod(dimfa), math.prod(dimfb)
        x = torch.rand(M, da, device=cuda_device)
        y = torch.rand(N, db, device=cuda_device)
        ref = torch.stack([sum(torch.kron(x[i].reshape(dimfa), y[j].reshape(dimfb)).flatten() for j in range(N))
                           for i in range(M)])
        out = api.LazyTensor(x[:, None, :]).keops_kron(api.LazyTensor(y[None, :, :]), dimfa, dimfb).sum(axis=1)
        assert out.shape == (M, da * db)
        assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5)
    al = ["x=Vi(0,6)", "y=Vj(1,4)"]
    x, y = torch.rand(M, 6, device=cuda_device), torch.rand(N, 4, device=cuda_device)
    o1 = api.Genred("Kron(x, y, [2, 3], [2, 2])", al, reduction_op="Sum", axis=1)(x, y)
    o2 = api.Genred("TensorDot(x, y, [2, 3], [2, 2], [], [], [0, 2, 1, 3])", al, reduction_op="Sum", axis=1)(x, y)
    assert torch.equal(o1, o2)


def test_host_tensors_and_the_multi_gpu_thread_pattern(api, cuda_device):
    # examples/pytorch/plot_multi_gpu.py:40-96 -- the reference's only multi-GPU code: host tensors, ``device_id=gpuid``, one
    # thread per GPU, blocks of the reduction index summed on the host / blocks of the output index concatenated.
    # With one visible GPU both workers land on device 0 (still two threads launching concurrently).
    from concurrent.futures import ThreadPoolExecutor, wait
    if AGAINST_REFERENCE and not torch.cuda.is_available():
        pytest.skip("device_id needs the reference's CUDA backend")
    torch.manual_seed(0)
    ngpus = max(1, torch.cuda.device_count())
    workers = max(2, ngpus)
    M, N = 6000, 9000
    x, y = torch.randn(M, 3), torch.randn(N, 3)

    def gauss_kernel_sum(x, y, gpuid=-1):
        x_i, y_j = api.LazyTensor(x[:, None, :]), api.LazyTensor(y[None, :, :])
        return (-((x_i - y_j) ** 2).sum(dim=2)).exp().sum(dim=1, device_id=gpuid)

    ref = torch.exp(-torch.cdist(x.double(), y.double()) ** 2).sum(1, keepdim=True)
    out = gauss_kernel_sum(x, y)                     # host in, host out, computed on the default GPU
    assert out.device.type == "cpu" and out.shape == (M, 1)
    assert torch.allclose(out.double(), ref, rtol=1e-5)
    for gpuid in range(ngpus):
        gauss_kernel_sum(x[:100], y[:100], gpuid)
    pool = ThreadPoolExecutor(workers)
    step = -(-N // workers)
    futs = [pool.submit(gauss_kernel_sum, x, y[k * step:(k + 1) * step], k % ngpus) for k in range(workers)]
    wait(futs)
    assert torch.allclose(sum(f.result().cpu() for f in futs).double(), ref, rtol=1e-5)
    step = -(-M // workers)
    futs = [pool.submit(gauss_kernel_sum, x[k * step:(k + 1) * step], y, k % ngpus) for k in range(workers)]
    wait(futs)
    assert torch.allclose(torch.cat([f.result().cpu() for f in futs]).double(), ref, rtol=1e-5)
    # gradients flow through the staged call as well
    xg = x[:500].clone().requires_grad_(True)
    gauss_kernel_sum(xg, y).sum().backward()
    xd = x[:500].double().requires_grad_(True)
    torch.exp(-torch.cdist(xd, y.double()) ** 2).sum().backward()
    assert xg.grad.device.type == "cpu" and torch.allclose(xg.grad.double(), xd.grad, rtol=1e-4, atol=1e-4 * xd.grad.abs().max().item())
    if not AGAINST_REFERENCE:
        with pytest.raises(ValueError):
            gauss_kernel_sum(x.to(cuda_device), y)                       # mixed memory locations


def test_complex_lazytensors_nonuniform_fourier_sum(api, cuda_device):
    # test_complex.py: X_k = sum_n x_n exp(-2 i pi p_n f_k) with complex x -- complex tensors are carried as interleaved
    # (re, im) pairs, python complex scalars become complex parameters; plus real / imag / abs / conj and the gradient
    # with respect to the complex signal
    torch.manual_seed(0)
    M, N = 300, 400
    x = torch.rand(1, N, 1, dtype=torch.complex64, device=cuda_device, requires_grad=True)
    p = torch.rand(1, N, 1, dtype=torch.float32, device=cuda_device)
    f = torch.rand(M, 1, 1, dtype=torch.float32, device=cuda_device)

    def fun(x, p, f, lazy):
        if lazy:
            x, p, f = api.LazyTensor(x), api.LazyTensor(p), api.LazyTensor(f)
        return (x * (-2 * math.pi * 1j * p * f).exp()).sum(dim=1)

    got, ref = fun(x, p, f, True).squeeze(), fun(x, p, f, False).squeeze()
    assert got.dtype == torch.complex64 and got.shape == ref.shape == (M,)
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4)
    X, P, F = api.LazyTensor(x), api.LazyTensor(p), api.LazyTensor(f)
    K = X * (P * F).exp1j()
    assert torch.allclose((K.real * K.imag + abs(K)).sum(dim=1).squeeze(),
                          ((x * torch.exp(1j * p * f)).real * (x * torch.exp(1j * p * f)).imag + (x * torch.exp(1j * p * f)).abs()).sum(1).squeeze(),
                          rtol=1e-4, atol=1e-4)
    assert torch.allclose((K.conj() * X).sum(dim=1).squeeze(), ((x * torch.exp(1j * p * f)).conj() * x).sum(1).squeeze(), rtol=1e-4, atol=1e-4)
    if not AGAINST_REFERENCE:       # the reference's own backward test is commented out (test_complex.py:45-58)
        g1 = torch.autograd.grad((got.real ** 2 + got.imag ** 2).sum(), [x])[0]
        g2 = torch.autograd.grad((ref.real ** 2 + ref.imag ** 2).sum(), [x])[0]
        assert torch.allclose(g1, g2, rtol=1e-3, atol=1e-3 * g2.abs().max().item())
