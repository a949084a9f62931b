"""Record what the REAL reference's block-sparse helpers return (``pykeops.torch.cluster``:
pykeops/pykeops/torch/cluster/grid_cluster.py:4, utils.py:4,50,101,152,235, matrix.py:5) on fixed CPU inputs, for
tests/test_cluster_golden.py.  Same environment as make_golden.py (only torch code runs, nothing is compiled).

    PYTHONPATH=/root/reference/keopscore:/root/reference/pykeops python tests/golden/make_cluster.py
"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    from pykeops.torch.cluster import (grid_cluster, cluster_ranges_centroids, cluster_ranges, cluster_centroids,
                                       sort_clusters, from_matrix)
    rec = {}
    g = torch.Generator().manual_seed(20261017)
    for D, size, thr in ((1, 0.25, 0.3), (2, 0.25, 0.35), (3, 0.25, 0.4), (3, 0.5, 0.7), (2, 0.1, 0.15)):
        tag = "d%d_s%s" % (D, str(size).replace(".", "p"))
        x, y = torch.rand(400, D, generator=g), torch.rand(300, D, generator=g) * 1.3 - 0.1
        w = torch.rand(400, generator=g)
        w[::7] = 0.0                                       # exercises min_weight clamping
        lx, ly = grid_cluster(x, size), grid_cluster(y, size)
        rx, cx, wx = cluster_ranges_centroids(x, lx)
        ry, cy, wy = cluster_ranges_centroids(y, ly)
        rxw, cxw, wxw = cluster_ranges_centroids(x, lx, weights=w)
        (xs, ws), lxs = sort_clusters((x, w), lx)
        ys, lys = sort_clusters(y, ly)
        keep = ((cx[:, None] - cy[None]) ** 2).sum(-1) < thr ** 2
        rng = from_matrix(rx, ry, keep)
        for k, v in dict(x=x, y=y, w=w, lx=lx, ly=ly, rx=rx, cx=cx, wx=wx, ry=ry, cy=cy, wy=wy, rxw=rxw, cxw=cxw, wxw=wxw,
                         xs=xs, ws=ws, lxs=lxs, ys=ys, lys=lys, keep=keep, cr=cluster_ranges(lxs),
                         cc=cluster_centroids(x, lx.long()),
                         r0=rng[0], r1=rng[1], r2=rng[2], r3=rng[3], r4=rng[4], r5=rng[5]).items():
            rec["%s.%s" % (tag, k)] = v.numpy()
    # the numpy flavour of the same helpers (pykeops/pykeops/numpy/cluster/*.py) on the same points: it anchors the grid
    # at the origin and does not clamp weights, so its outputs are recorded separately under "np_<tag>"
    from pykeops.numpy import cluster as npc
    for tag in sorted({k.split(".")[0] for k in rec}):
        x, y, w = rec[tag + ".x"], rec[tag + ".y"], rec[tag + ".w"] + 0.05
        size = float(tag.split("_s")[1].replace("p", "."))
        lx, ly = npc.grid_cluster(x, size), npc.grid_cluster(y, size)
        rx, cx, wx = npc.cluster_ranges_centroids(x, lx)
        ry, cy, wy = npc.cluster_ranges_centroids(y, ly)
        rxw, cxw, wxw = npc.cluster_ranges_centroids(x, lx, weights=w)
        (xs, ws), lxs = npc.sort_clusters((x, w), lx)
        keep = ((cx[:, None] - cy[None]) ** 2).sum(-1) < (1.4 * size) ** 2
        rng = npc.from_matrix(rx, ry, keep)
        for k, v in dict(w=w, lx=lx, ly=ly, rx=rx, cx=cx, wx=wx, ry=ry, cy=cy, rxw=rxw, cxw=cxw, wxw=wxw, xs=xs, ws=ws, lxs=lxs,
                         keep=keep, cr=npc.cluster_ranges(lxs), cc=npc.cluster_centroids(x, lx),
                         r0=rng[0], r1=rng[1], r2=rng[2], r3=rng[3], r4=rng[4], r5=rng[5]).items():
            rec["np_%s.%s" % (tag, k)] = np.asarray(v)
    # the worked example of the reference's from_matrix docstring (matrix.py:59-110): gaps between clusters
    r_i = torch.IntTensor([[2, 5], [7, 12]])
    r_j = torch.IntTensor([[1, 4], [4, 9], [20, 30]])
    keep = torch.tensor([[1, 1, 0], [0, 1, 0]], dtype=torch.bool)
    rng = from_matrix(r_i, r_j, keep)
    for k, v in dict(ri=r_i, rj=r_j, keep=keep, r0=rng[0], r1=rng[1], r2=rng[2], r3=rng[3], r4=rng[4], r5=rng[5]).items():
        rec["doc.%s" % k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "cluster.npz"), **rec)
    print("recorded %d arrays" % len(rec))


if __name__ == "__main__":
    main()
