"""GPU: the pykeops.numpy-compatible front-end (keops_b200/numpy) -- host arrays in, host arrays out, kernels on the B200.

Shapes and checks follow the reference's numpy tests ("KeOps vs dense numpy"): pykeops/pykeops/numpy/test_install.py:5-25,
pykeops/pykeops/test/test_numpy.py:60-120,260-285,300-340.  The very same test bodies were run once against the REAL
reference (``pykeops.numpy``, CPU backend; ``KN_MODULE=pykeops.numpy python -m pytest tests/test_numpy_api_gpu.py``
without the gpu marker) to pin the expected shapes, dtypes and values."""
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kn(request):
    name = os.environ.get("KN_MODULE", "keops_b200.numpy")
    if name == "keops_b200.numpy":
        request.getfixturevalue("cuda_device")
    return importlib.import_module(name)


def gauss(x, y, g=1.0):
    return np.exp(-g * ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1))


def test_install_known_answer(kn):
    x = np.arange(1, 10).reshape(-1, 3).astype("float32")
    y = np.arange(3, 9).reshape(-1, 3).astype("float32")
    out = kn.Genred("SqNorm2(x-y)", ["x=Vi(3)", "y=Vj(3)"], reduction_op="Sum", axis=0)(x, y)
    assert isinstance(out, np.ndarray) and out.dtype == np.float32
    assert out.reshape(-1).tolist() == [63.0, 90.0]


@pytest.mark.parametrize("dtype,tol", [("float32", 1e-5), ("float64", 1e-12)])
def test_genred_sum_and_out(kn, dtype, tol):
    rng = np.random.default_rng(0)
    M, N = 700, 900
    x, y, b = rng.random((M, 3)).astype(dtype), rng.random((N, 3)).astype(dtype), rng.standard_normal((N, 2)).astype(dtype)
    g = np.array([0.7], dtype=dtype)
    conv = kn.Genred("Exp(-g*SqDist(x,y))*b", ["x=Vi(3)", "y=Vj(3)", "b=Vj(2)", "g=Pm(1)"], reduction_op="Sum", axis=1)
    ref = gauss(x, y, 0.7) @ b
    res = conv(x, y, b, g)
    assert res.shape == (M, 2) and res.dtype == np.dtype(dtype)
    assert np.abs(res - ref).max() <= tol * np.abs(ref).max()
    out = np.zeros((M, 2), dtype=dtype)
    res2 = conv(x, y, b, g, out=out)
    assert res2 is out and np.array_equal(out, res)
    res0 = kn.Genred("Exp(-g*SqDist(x,y))*a", ["x=Vi(3)", "y=Vj(3)", "a=Vi(2)", "g=Pm(1)"], reduction_op="Sum", axis=0)(
        x, y, rng.standard_normal((M, 2)).astype(dtype) * 0 + 1, g)
    assert res0.shape == (N, 2)
    assert np.abs(res0[:, 0] - gauss(x, y, 0.7).sum(0)).max() <= tol * N


def test_argkmin_and_generic_helpers(kn):
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal((500, 3)).astype("float32"), rng.standard_normal((2000, 3)).astype("float32")
    d = ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)
    ref = np.argsort(d, axis=1, kind="stable")[:, :5]
    ind = kn.Genred("SqDist(x,y)", ["x=Vi(3)", "y=Vj(3)"], reduction_op="ArgKMin", axis=1, opt_arg=5)(x, y)
    assert ind.shape == (500, 5) and ind.dtype == np.int64
    same = (ind == ref).mean()
    assert same > 0.999                                    # dense numpy rounds the distances differently on near-ties
    ind2 = kn.generic_argkmin("SqDist(x,y)", "a = Vi(5)", "x = Vi(3)", "y = Vj(3)")(x, y)
    assert np.array_equal(ind2, ind)
    am = kn.generic_argmin("SqDist(x,y)", "a = Vi(1)", "x = Vi(3)", "y = Vj(3)")(x, y)
    assert am.shape == (500, 1) and np.array_equal(am[:, 0], ind[:, 0])
    s = kn.generic_sum("Exp(-SqDist(x,y))", "a = Vi(1)", "x = Vi(3)", "y = Vj(3)")(x, y)
    assert np.abs(s[:, 0] - np.exp(-d).sum(1)).max() <= 1e-5 * np.exp(-d).sum(1).max()
    lse = kn.generic_logsumexp("-SqDist(x,y)", "a = Vi(1)", "x = Vi(3)", "y = Vj(3)")(x, y)
    assert np.abs(lse[:, 0] - np.log(np.exp(-d.astype("float64")).sum(1))).max() <= 1e-4


def test_lazytensor_reductions(kn):
    rng = np.random.default_rng(2)
    M, N = 600, 800
    x, y, b = rng.random((M, 3)).astype("float32"), rng.random((N, 3)).astype("float32"), rng.standard_normal((N, 2)).astype("float32")
    x_i, y_j = kn.LazyTensor(x[:, None, :]), kn.LazyTensor(y[None, :, :])
    D_ij = ((x_i - y_j) ** 2).sum(-1)
    K_ij = (-D_ij * 2.0).exp()
    assert K_ij.shape == (M, N)
    ref = gauss(x, y, 2.0)
    out = K_ij @ b
    assert isinstance(out, np.ndarray) and out.shape == (M, 2)
    assert np.abs(out - ref @ b).max() <= 1e-5 * np.abs(ref @ b).max()
    v = rng.standard_normal(N).astype("float32")
    assert (K_ij @ v).shape == (M,)
    assert np.abs(K_ij @ v - ref @ v).max() <= 1e-5 * np.abs(ref).sum(1).max()
    outT = K_ij.T @ rng.random((M, 1)).astype("float32")
    assert outT.shape == (N, 1)
    s0 = (K_ij * kn.Vi(x)).sum(0)
    assert s0.shape == (N, 3) and np.abs(s0 - ref.T @ x).max() <= 1e-5 * np.abs(ref.T @ x).max()
    nn = D_ij.argmin(1)
    d = ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)
    assert nn.shape == (M, 1) and nn.dtype == np.int64 and (nn[:, 0] == d.argmin(1)).mean() > 0.995
    knn = D_ij.argKmin(3, dim=1)
    assert knn.shape == (M, 3) and np.array_equal(knn[:, 0], nn[:, 0])
    mn, amn = D_ij.min_argmin(dim=1)
    assert np.array_equal(amn, nn) and np.abs(mn[:, 0] - d.min(1)).max() <= 1e-6
    lse = (-D_ij).logsumexp(1, weight=kn.Vj(np.abs(b[:, :1]) + 0.1))
    lref = np.log((np.exp(-d.astype("float64")) * (np.abs(b[:, 0]) + 0.1)[None, :]).sum(1))
    assert lse.shape == (M, 1) and np.abs(lse[:, 0] - lref).max() <= 1e-4
    p = kn.Pm(np.array([0.5], dtype="float32"))
    assert np.abs(((-D_ij * p).exp() @ b) - gauss(x, y, 0.5) @ b).max() <= 1e-4


def test_batch_dimensions(kn):
    rng = np.random.default_rng(3)
    B, M, N = 4, 50, 70
    x, y, b = rng.random((B, M, 1, 3)).astype("float32"), rng.random((1, 1, N, 3)).astype("float32"), rng.random((B, 1, N, 2)).astype("float32")
    K = (-((kn.LazyTensor(x) - kn.LazyTensor(y)) ** 2).sum(-1)).exp()
    assert K.shape == (B, M, N)
    out = (K * kn.LazyTensor(b)).sum(2)
    ref = np.einsum("bmn,bnd->bmd", np.exp(-((x - y) ** 2).sum(-1)), b[:, 0])
    assert out.shape == (B, M, 2) and np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max()


def test_kernel_solve(kn):
    rng = np.random.default_rng(4)
    N = 400
    x, b = rng.random((N, 3)).astype("float64"), rng.random((N, 2)).astype("float64")
    g = np.array([4.0])
    Kinv = kn.KernelSolve("Exp(-g*SqDist(x,y))*b", ["x=Vi(3)", "y=Vj(3)", "b=Vj(2)", "g=Pm(1)"], "b", axis=1)
    a = Kinv(x, x, b, g, alpha=0.1, eps=1e-9)
    ref = np.linalg.solve(gauss(x, x, 4.0) + 0.1 * np.eye(N), b)
    assert isinstance(a, np.ndarray) and a.shape == (N, 2) and np.abs(a - ref).max() <= 1e-6
    K = (-kn.Vi(x).sqdist(kn.Vj(x)) * 4.0).exp()
    a2 = K.solve(kn.Vi(b), alpha=0.1, eps=1e-9)
    assert np.abs(a2 - ref).max() <= 1e-6


def test_block_sparse_ranges(kn):
    rng = np.random.default_rng(5)
    x, y = np.sort(rng.random((300, 1)), 0).astype("float32"), np.sort(rng.random((400, 1)), 0).astype("float32")
    b = rng.random((400, 1)).astype("float32")
    ri = np.array([[0, 100], [100, 300]], dtype=np.int32)
    rj = np.array([[0, 150], [150, 250], [250, 400]], dtype=np.int32)
    # i-cluster 0 sees j-clusters 0 and 1, i-cluster 1 sees j-cluster 2
    ranges = (ri, np.array([2, 3], dtype=np.int32), np.array([[0, 150], [150, 250], [250, 400]], dtype=np.int32),
              rj, np.array([1, 2, 3], dtype=np.int32), np.array([[0, 100], [0, 100], [100, 300]], dtype=np.int32))
    out = kn.Genred("Exp(-SqDist(x,y))*b", ["x=Vi(1)", "y=Vj(1)", "b=Vj(1)"], reduction_op="Sum", axis=1)(x, y, b, ranges=ranges)
    mask = np.zeros((300, 400), dtype=bool)
    mask[:100, :250] = True
    mask[100:, 250:] = True
    ref = (gauss(x, y) * mask) @ b
    assert np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max()


def test_non_contiguous_inputs_and_inplace_out(kn):
    """pykeops/pykeops/test/test_contiguous_numpy.py and test_lazytensor_gaussian_numpy_inplace.py."""
    rng = np.random.default_rng(6)
    a1 = rng.random((2, 500, 5)).astype("float32")
    a2 = np.ascontiguousarray(a1.transpose(2, 0, 1)).transpose(1, 2, 0)
    assert not a2.flags["C_CONTIGUOUS"] and np.array_equal(a1, a2)
    b, c = rng.random((2, 500, 5)).astype("float32"), rng.random((2, 500, 5)).astype("float32")
    b_j = kn.LazyTensor(b[:, None])
    d1 = (-kn.LazyTensor(a1[:, :, None]).sqdist(b_j)).exp() @ c
    d2 = (-kn.LazyTensor(a2[:, :, None]).sqdist(b_j)).exp() @ c
    ref = np.einsum("bmn,bnd->bmd", np.exp(-((a1[:, :, None] - b[:, None]) ** 2).sum(-1)), c)
    assert d1.shape == (2, 500, 5) and np.array_equal(d1, d2)
    assert np.abs(d1 - ref).max() <= 1e-5 * np.abs(ref).max()

    M, N = 300, 200
    x = rng.random((M, 1, 3)).astype("float32") / np.sqrt(3).astype("float32")
    y = rng.random((1, N, 3)).astype("float32") / np.sqrt(3).astype("float32")
    w = rng.standard_normal((N, 1)).astype("float32")
    a = np.empty((M, 1), dtype="float32")
    K = (-((kn.LazyTensor(x) - kn.LazyTensor(y)) ** 2).sum(axis=2)).exp()
    ret = K.__matmul__(w, out=a)
    assert ret is a
    assert np.abs(a - np.exp(-((x - y) ** 2).sum(-1)) @ w).max() <= 1e-5


def test_kron_tensordot_and_complex_on_numpy_arrays(kn):
    # test_kron.py / test_complex_numpy.py: the contraction operators and complex LazyTensors through the numpy front-end
    rng = np.random.default_rng(0)
    M, N = 11, 150
    dimfa, dimfb = [2, 3], [2, 2]
    x, y = rng.random((M, 6)), rng.random((N, 4))
    ref = np.stack([sum(np.kron(x[i].reshape(dimfa), y[j].reshape(dimfb)).ravel() for j in range(N)) for i in range(M)])
    out = kn.LazyTensor(x[:, None, :]).keops_kron(kn.LazyTensor(y[None, :, :]), dimfa, dimfb).sum(axis=1)
    assert isinstance(out, np.ndarray) and out.shape == (M, 24) and np.allclose(out, ref, atol=1e-10)
    al = ["x=Vi(0,6)", "y=Vj(1,4)"]
    o2 = kn.Genred("TensorDot(x, y, [2, 3], [2, 2], [], [], [0, 2, 1, 3])", al, reduction_op="Sum", axis=1)(x, y)
    assert np.allclose(o2, ref, atol=1e-10)
    a = rng.random((4, 7)); b = rng.random(7)
    C = kn.LazyTensor(a.reshape(1, 1, 28)).keops_tensordot(kn.LazyTensor(b.reshape(1, 1, 7)), (4, 7), (7,), (1,), (0,)).sum_reduction(axis=1)
    assert np.allclose(C.ravel(), a @ b)
    # non-uniform Fourier sum with a complex signal
    z = (rng.random((1, N, 1)) + 1j * rng.random((1, N, 1))).astype("complex64")
    p, f = rng.random((1, N, 1)).astype("float32"), rng.random((M, 1, 1)).astype("float32")
    Z, P, F = kn.LazyTensor(z), kn.LazyTensor(p), kn.LazyTensor(f)
    got = (Z * (-2 * np.pi * 1j * P * F).exp()).sum(axis=1)
    want = (z * np.exp(-2 * np.pi * 1j * p * f)).sum(axis=1)
    assert got.dtype == np.complex64 and np.allclose(got.ravel(), want.ravel(), rtol=1e-4, atol=1e-4)
