"""Block-sparse ``ranges`` mode: the oracle against golden vectors from the real reference
(``tests/golden/make_golden_ranges.py``), and the CUDA kernel ``kb200_reduce_ranges`` against both."""
import json
import os

import numpy as np
import pytest

from oracle import keops_oracle as ko

HERE = os.path.dirname(os.path.abspath(__file__))
RDIR = os.path.join(HERE, "golden", "ranges")
with open(os.path.join(RDIR, "index.json")) as f:
    INDEX = json.load(f)


def load(name):
    z = np.load(os.path.join(RDIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    args = [z["arg%d" % k] for k in range(meta["nargs"])]
    ranges = tuple(z["range%d" % k] for k in range(6))
    return meta, args, ranges, z["out"]


def check(meta, got, ref):
    tol = 1e-5 if meta["dtype"] == "float32" else 1e-12
    head = meta["formula"].split("(")[0]
    if head.startswith("Arg"):
        assert np.array_equal(got, ref)
    elif head.startswith("Min_ArgMin"):
        d = ref.shape[1] // 2
        assert np.array_equal(got[:, d:], ref[:, d:])
        np.testing.assert_allclose(got[:, :d], ref[:, :d], rtol=tol, atol=tol * np.abs(ref[:, :d]).max())
    else:
        np.testing.assert_allclose(got, ref, rtol=tol, atol=tol * np.abs(ref).max())


@pytest.mark.parametrize("name", [m["name"] for m in INDEX])
def test_oracle_ranges_vs_reference(name):
    meta, args, ranges, ref = load(name)
    got = ko.genred_ranges(meta["formula"], [], args, ranges, meta["dtype"])
    check(meta, got, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("name", [m["name"] for m in INDEX])
def test_cuda_ranges_vs_reference(name, cuda_device):
    import torch
    from keops_b200.binder import LoadKeOpsB200
    meta, args, ranges, ref = load(name)
    flags = dict(dtype_acc="auto", sum_scheme="auto", use_fast_math=1, enable_chunks=1, multVar_highdim=0)
    b = LoadKeOpsB200(1, 0, 1, True, -1, meta["formula"], [], meta["nargs"], meta["dtype"], "torch", flags)
    targs = [torch.from_numpy(a).to(cuda_device) for a in args]
    tr = tuple(torch.from_numpy(r).to(cuda_device) for r in ranges)
    out = b.import_module().genred_pytorch(cuda_device, tr, meta["nx"], meta["ny"], 0, None, *targs)
    torch.cuda.synchronize()
    assert b.dimout == meta["dimout"] and b.tagIJ == meta["tagIJ"]
    check(meta, out.cpu().numpy(), ref)


@pytest.mark.gpu
def test_cuda_ranges_large_vs_dense_blocks(cuda_device):
    """Block-diagonal ranges on 20k points == independent dense reductions per block (size-independent property);
    row ranges longer than one CTA's row chunk exercise the (range, chunk) bisection."""
    import torch
    from keops_b200.binder import get_binder
    g = "Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),0)"
    torch.manual_seed(0)
    M, N = 20000, 15000
    x, y, b = torch.rand(M, 3, device=cuda_device), torch.rand(N, 3, device=cuda_device), torch.rand(N, 1, device=cuda_device)
    ex = torch.tensor([0, 3000, 3001, 9000, 20000])
    ey = torch.tensor([0, 100, 7000, 7300, 15000])
    ri = torch.stack([ex[:-1], ex[1:]], 1)
    rj = torch.stack([ey[:-1], ey[1:]], 1)
    sl = torch.arange(1, 5)
    ranges = tuple(t.to(cuda_device) for t in (ri, sl, rj, rj, sl, ri))
    bd = get_binder(g, [], 3, "float32")
    out = bd.genred_pytorch(cuda_device, ranges, M, N, 0, None, x, y, b)
    for k in range(4):
        r0, r1, c0, c1 = int(ex[k]), int(ex[k + 1]), int(ey[k]), int(ey[k + 1])
        ref = bd.genred_pytorch(cuda_device, None, r1 - r0, c1 - c0, 0, None, x[r0:r1].contiguous(),
                                y[c0:c1].contiguous(), b[c0:c1].contiguous())
        torch.testing.assert_close(out[r0:r1], ref, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        bd.genred_pytorch(cuda_device, ranges, M, N, 1, None, x[None], y[None], b[None])


def test_cluster_helpers_cpu():
    """``from_matrix`` covers exactly the kept cluster pairs, on both sides (host-side logic, no GPU)."""
    import torch
    from keops_b200.torch.cluster import grid_cluster, cluster_ranges_centroids, sort_clusters, from_matrix, swap_axes
    torch.manual_seed(1)
    x, y = torch.rand(400, 2), torch.rand(300, 2)
    lx, ly = grid_cluster(x, 0.25), grid_cluster(y, 0.25)
    assert lx.dtype == torch.int32 and int(lx.min()) == 0 and int(lx.max()) + 1 == len(torch.unique(lx))
    rx, cx, wx = cluster_ranges_centroids(x, lx)
    ry, cy, wy = cluster_ranges_centroids(y, ly)
    xs, lxs = sort_clusters(x, lx)
    ys, lys = sort_clusters(y, ly)
    assert torch.all(lxs[1:] >= lxs[:-1]) and int(wx.sum()) == 400
    for k in range(rx.shape[0]):
        assert torch.all(lxs[int(rx[k, 0]):int(rx[k, 1])] == k)
        torch.testing.assert_close(cx[k], xs[int(rx[k, 0]):int(rx[k, 1])].mean(0))
    keep = ((cx[:, None] - cy[None]) ** 2).sum(-1) < 0.35 ** 2
    rg = from_matrix(rx, ry, keep)
    want = keep[lxs.long()][:, lys.long()]

    def cover(r_out, sl, r_red, transpose):
        m = torch.zeros(400, 300, dtype=torch.bool)
        for k in range(r_out.shape[0]):
            for l in range(0 if k == 0 else int(sl[k - 1]), int(sl[k])):
                a, b, c, d = int(r_out[k, 0]), int(r_out[k, 1]), int(r_red[l, 0]), int(r_red[l, 1])
                if transpose:
                    m[c:d, a:b] = True
                else:
                    m[a:b, c:d] = True
        return m
    assert torch.equal(cover(rg[0], rg[1], rg[2], False), want)
    assert torch.equal(cover(rg[3], rg[4], rg[5], True), want)
    sw = swap_axes(rg)
    assert sw[0] is rg[3] and sw[5] is rg[2]


@pytest.mark.gpu
def test_lazytensor_block_sparse_gpu(cuda_device):
    """Same scenario as the reference's ``pykeops/pykeops/test/test_block_sparse_reduction.py:16-91``:
    a truncated Gaussian conv through cluster ranges agrees with the dense one within 0.2 %."""
    import torch
    from keops_b200.torch import LazyTensor
    from keops_b200.torch.cluster import grid_cluster, cluster_ranges_centroids, sort_clusters, from_matrix
    torch.manual_seed(0)
    M, N = 5000, 5000
    t = torch.linspace(0, 2 * np.pi, M + 1)[:-1]
    x = torch.stack((0.4 + 0.4 * torch.cos(t), 0.5 + 0.3 * torch.sin(t)), 1) + 0.01 * torch.randn(M, 2)
    x = x.to(cuda_device)
    y = (torch.randn(N, 2) / 10 + 0.6).to(cuda_device)
    eps, sigma = 0.05, 0.05
    xl, yl = grid_cluster(x, eps), grid_cluster(y, eps)
    xr, xc, _ = cluster_ranges_centroids(x, xl)
    yr, yc, _ = cluster_ranges_centroids(y, yl)
    x, xl = sort_clusters(x, xl)
    y, yl = sort_clusters(y, yl)
    keep = ((xc[:, None, :] - yc[None, :, :]) ** 2).sum(2) < (4 * sigma) ** 2
    ranges = from_matrix(xr, yr, keep)
    xi, yj = LazyTensor((x / sigma)[:, None, :]), LazyTensor((y / sigma)[None, :, :])
    K = (-((xi - yj) ** 2).sum(2) / 2).exp()
    b = torch.randn(N, 1, device=cuda_device)
    dense = K @ b
    K.ranges = ranges
    sparse = K @ b
    rel = float((dense - sparse).abs().sum() / dense.abs().sum())
    assert rel < 2e-3, rel
    # reduction over i uses the last three arrays
    a = torch.randn(M, 1, device=cuda_device)
    K.ranges = None
    dense_t = (K * LazyTensor(a[:, None, :])).sum(0)
    K.ranges = ranges
    sparse_t = (K * LazyTensor(a[:, None, :])).sum(0)
    rel = float((dense_t - sparse_t).abs().sum() / dense_t.abs().sum())
    assert rel < 3e-3, rel          # truncation error of the 4-sigma cut-off itself: the real reference gives 2.0e-3 here
