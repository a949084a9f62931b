"""Block-sparse helpers (keops_b200/torch/cluster.py) against recordings of the REAL reference's
``pykeops.torch.cluster`` functions (tests/golden/make_cluster.py -> tests/golden/cluster.npz).

Exact where the reference's result is determined (labels, cluster ranges, sorted labels, the output-side range lists);
semantic where it leaves freedom: ``sort_clusters`` in the reference uses an unstable sort (utils.py:4-47), so the order
of the points INSIDE a cluster is compared as a multiset, and ``from_matrix`` here fuses adjacent kept clusters into
longer ranges (fewer tiles for the ranges kernel), so the reduction side is compared as the set of (i, j) pairs covered.
"""
import os

import numpy as np
import pytest
import torch

from keops_b200.torch.cluster import (grid_cluster, cluster_ranges_centroids, cluster_ranges, cluster_centroids,
                                      sort_clusters, from_matrix, swap_axes)

REC = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cluster.npz"))
TAGS = sorted({k.split(".")[0] for k in REC.files if not k.startswith(("doc.", "np_"))})


def t(tag, k):
    return torch.from_numpy(REC["%s.%s" % (tag, k)])


def covered(ranges_out, slices, redranges, n_out, n_red):
    """dense (n_out, n_red) boolean mask of the pairs a (ranges, slices, redranges) triple schedules
    (semantics of keopscore/keopscore/binders/nvrtc/keops_nvrtc.cpp:75-325 / include/Ranges.h:40-86)"""
    m = np.zeros((n_out, n_red), dtype=np.int32)
    start = 0
    for k in range(ranges_out.shape[0]):
        a, b = int(ranges_out[k, 0]), int(ranges_out[k, 1])
        for s in range(start, int(slices[k])):
            m[a:b, int(redranges[s, 0]):int(redranges[s, 1])] += 1
        start = int(slices[k])
    return m


@pytest.mark.parametrize("tag", TAGS)
def test_grid_cluster_and_ranges(tag):
    x, y, w = t(tag, "x"), t(tag, "y"), t(tag, "w")
    lx, ly = grid_cluster(x, float(tag.split("_s")[1].replace("p", "."))), grid_cluster(y, float(tag.split("_s")[1].replace("p", ".")))
    assert lx.dtype == t(tag, "lx").dtype and torch.equal(lx, t(tag, "lx")) and torch.equal(ly, t(tag, "ly"))
    rx, cx, wx = cluster_ranges_centroids(x, lx)
    assert rx.dtype == t(tag, "rx").dtype and torch.equal(rx, t(tag, "rx"))
    torch.testing.assert_close(cx, t(tag, "cx"), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(wx.to(t(tag, "wx").dtype), t(tag, "wx"))
    rxw, cxw, wxw = cluster_ranges_centroids(x, lx, weights=w)
    assert torch.equal(rxw, t(tag, "rxw"))
    torch.testing.assert_close(cxw, t(tag, "cxw"), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(wxw.to(t(tag, "wxw").dtype), t(tag, "wxw"), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(cluster_centroids(x, lx.long()), t(tag, "cc"), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("tag", TAGS)
def test_sort_clusters(tag):
    x, w, lx = t(tag, "x"), t(tag, "w"), t(tag, "lx")
    (xs, ws), lxs = sort_clusters((x, w), lx)
    assert torch.equal(lxs, t(tag, "lxs"))
    assert torch.equal(cluster_ranges(lxs), t(tag, "cr"))
    ref_xs, ref_ws = t(tag, "xs"), t(tag, "ws")
    for a, b in t(tag, "cr").tolist():                     # same points in every cluster, any order inside it
        mine = torch.cat((xs[a:b], ws[a:b, None]), 1).tolist()
        ref = torch.cat((ref_xs[a:b], ref_ws[a:b, None]), 1).tolist()
        assert sorted(mine) == sorted(ref)
    ys, lys = sort_clusters(t(tag, "y"), t(tag, "ly"))     # single-tensor form
    assert torch.equal(lys, t(tag, "lys")) and ys.shape == t(tag, "ys").shape


@pytest.mark.parametrize("tag", TAGS + ["doc"])
def test_from_matrix(tag):
    if tag == "doc":
        ri, rj = t(tag, "ri"), t(tag, "rj")
        n_i, n_j = 12, 30
    else:
        ri, rj = t(tag, "rx"), t(tag, "ry")
        n_i, n_j = t(tag, "x").shape[0], t(tag, "y").shape[0]
    keep = t(tag, "keep")
    mine = from_matrix(ri, rj, keep)
    ref = tuple(t(tag, "r%d" % k) for k in range(6))
    assert all(m.dtype == r.dtype for m, r in zip(mine, ref))
    assert torch.equal(mine[0], ref[0]) and torch.equal(mine[3], ref[3])
    for side, (n_out, n_red) in ((0, (n_i, n_j)), (3, (n_j, n_i))):
        a = covered(mine[side], mine[side + 1], mine[side + 2], n_out, n_red)
        b = covered(ref[side], ref[side + 1], ref[side + 2], n_out, n_red)
        assert a.max() <= 1                                # no pair scheduled twice
        assert np.array_equal(a, b)
        assert mine[side + 2].shape[0] <= ref[side + 2].shape[0]    # fused runs: never more tiles than the reference
        assert mine[side + 1].shape == ref[side + 1].shape
    assert np.array_equal(covered(mine[0], mine[1], mine[2], n_i, n_j), covered(mine[3], mine[4], mine[5], n_j, n_i).T)
    sw = swap_axes(mine)
    assert all(torch.equal(u, v) for u, v in zip(sw, mine[3:] + mine[:3])) and swap_axes(None) is None


@pytest.mark.parametrize("tag", TAGS)
def test_numpy_flavour(tag):
    """keops_b200.numpy.cluster against recordings of ``pykeops.numpy.cluster`` (grid anchored at the origin, weights not
    clamped, int64 everywhere): same exact / semantic split as the torch flavour above."""
    from keops_b200.numpy import cluster as npc
    x, y = REC[tag + ".x"], REC[tag + ".y"]
    r = lambda k: REC["np_%s.%s" % (tag, k)]
    size = float(tag.split("_s")[1].replace("p", "."))
    w = r("w")
    lx, ly = npc.grid_cluster(x, size), npc.grid_cluster(y, size)
    assert lx.dtype == r("lx").dtype and np.array_equal(lx, r("lx")) and np.array_equal(ly, r("ly"))
    rx, cx, wx = npc.cluster_ranges_centroids(x, lx)
    assert rx.dtype == r("rx").dtype and np.array_equal(rx, r("rx")) and cx.dtype == r("cx").dtype
    np.testing.assert_allclose(cx, r("cx"), rtol=1e-6, atol=1e-6)
    assert wx.dtype == r("wx").dtype and np.array_equal(wx, r("wx"))
    rxw, cxw, wxw = npc.cluster_ranges_centroids(x, lx, weights=w)
    assert np.array_equal(rxw, r("rxw"))
    np.testing.assert_allclose(cxw, r("cxw"), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(wxw, r("wxw"), rtol=1e-6)
    np.testing.assert_allclose(npc.cluster_centroids(x, lx), r("cc"), rtol=1e-6, atol=1e-6)
    (xs, ws), lxs = npc.sort_clusters((x, w), lx)
    assert np.array_equal(lxs, r("lxs")) and np.array_equal(npc.cluster_ranges(lxs), r("cr"))
    for a, b in r("cr").tolist():
        assert sorted(np.concatenate((xs[a:b], ws[a:b, None]), 1).tolist()) == \
               sorted(np.concatenate((r("xs")[a:b], r("ws")[a:b, None]), 1).tolist())
    xl, _ = npc.sort_clusters([x, w], lx)
    assert isinstance(xl, list) and np.array_equal(xl[0], xs)
    ry = npc.cluster_ranges_centroids(y, ly)[0]
    mine = npc.from_matrix(rx, ry, r("keep"))
    ref = tuple(r("r%d" % k) for k in range(6))
    assert all(m.dtype == q.dtype for m, q in zip(mine, ref))
    assert np.array_equal(mine[0], ref[0]) and np.array_equal(mine[3], ref[3])
    for side, (n_out, n_red) in ((0, (x.shape[0], y.shape[0])), (3, (y.shape[0], x.shape[0]))):
        a = covered(mine[side], mine[side + 1], mine[side + 2], n_out, n_red)
        assert a.max() <= 1 and np.array_equal(a, covered(ref[side], ref[side + 1], ref[side + 2], n_out, n_red))
    assert all(u is v for u, v in zip(npc.swap_axes(mine), mine[3:] + mine[:3])) and npc.swap_axes(None) is None
