"""Block-sparse helpers on numpy arrays -- API of ``pykeops.numpy.cluster``
(``pykeops/pykeops/numpy/cluster/{grid_cluster,utils,matrix}.py``).  Host-side index bookkeeping only; the 6-tuple built
by :func:`from_matrix` is what ``keops_b200.numpy.Genred(..., ranges=...)`` stages on the GPU.

Two behaviours of the reference's numpy helpers differ from its torch ones and are followed here: ``grid_cluster`` anchors
the cells at the origin (not at the bounding box corner, ``numpy/cluster/grid_cluster.py:41-45`` vs
``torch/cluster/grid_cluster.py:41``), and ``cluster_ranges_centroids`` does not clamp small weights.
"""
from __future__ import annotations

import numpy as np
import torch

from ..torch import cluster as _tc

__all__ = ["grid_cluster", "sort_clusters", "cluster_ranges", "cluster_centroids", "cluster_ranges_centroids",
           "from_matrix", "swap_axes"]


def grid_cluster(x, size):
    """Labels in a compact range [0, C) of the cubic cell (edge ``size``, anchored at the origin) that contains each
    point of ``x`` (M, D), D <= 3; cells are numbered in lexicographic order of their integer coordinates."""
    if x.ndim != 2 or not 1 <= x.shape[1] <= 3:
        raise NotImplementedError("grid_cluster handles (M, D) point clouds with D in {1, 2, 3}")
    cell = np.floor(x / size).astype(np.int64)
    key = cell[:, 0]
    for k in range(1, x.shape[1]):
        key = key * 1024 + cell[:, k]
    return np.unique(key, return_inverse=True)[1].reshape(-1).astype(np.int64)


def sort_clusters(x, lab):
    """Sort points (an array, or a tuple / list of arrays) by label so that clusters are contiguous."""
    perm = np.argsort(lab.ravel(), kind="stable")
    if isinstance(x, tuple):
        return tuple(a[perm] for a in x), lab[perm]
    if isinstance(x, list):
        return [a[perm] for a in x], lab[perm]
    return x[perm], lab[perm]


def cluster_ranges(lab, Nlab=None):
    """(C, 2) int64 array of [start, end) positions of each cluster in the label-sorted point list."""
    counts = np.bincount(lab) if Nlab is None else np.asarray(Nlab)
    ends = np.cumsum(counts, axis=0)
    return np.stack((ends - counts, ends), axis=1).astype(np.int64)


def cluster_centroids(x, lab, Nlab=None, weights=None, weights_c=None):
    """(C, D) weighted means of the points of each cluster, in the dtype of ``x``."""
    C = int(lab.max()) + 1 if Nlab is None else len(Nlab)
    w = np.ones(x.shape[0]) if weights is None else np.asarray(weights, dtype=np.float64).ravel()
    if weights is None:
        tot = np.bincount(lab, minlength=C).astype(np.float64) if Nlab is None else np.asarray(Nlab, dtype=np.float64)
    else:
        tot = np.bincount(lab, weights=w, minlength=C) if weights_c is None else np.asarray(weights_c, dtype=np.float64).ravel()
    acc = np.zeros((C, x.shape[1]))
    np.add.at(acc, lab, x * w[:, None])
    return (acc / tot[:, None]).astype(x.dtype)


def cluster_ranges_centroids(x, lab, weights=None):
    """``(ranges, centroids, cluster weights)`` for points ``x`` with integer labels ``lab``."""
    Nlab = np.bincount(lab).astype(np.float64)
    if weights is None:
        return cluster_ranges(lab, Nlab), cluster_centroids(x, lab, Nlab), Nlab
    w_c = np.bincount(lab, weights=np.asarray(weights).ravel())
    return cluster_ranges(lab, Nlab), cluster_centroids(x, lab, Nlab, weights=weights, weights_c=w_c), w_c


def from_matrix(ranges_i, ranges_j, keep):
    """Boolean (Ci, Cj) mask between i-clusters and j-clusters -> the 6-tuple of ``ranges`` (int64 arrays); adjacent kept
    clusters are fused into longer ranges exactly as in :func:`keops_b200.torch.cluster.from_matrix`."""
    out = _tc.from_matrix(torch.from_numpy(np.ascontiguousarray(ranges_i, dtype=np.int64)),
                          torch.from_numpy(np.ascontiguousarray(ranges_j, dtype=np.int64)),
                          torch.from_numpy(np.ascontiguousarray(keep).astype(bool)))
    return tuple(t.numpy() for t in out)


def swap_axes(ranges):
    """Ranges of the transposed problem (i <-> j)."""
    if ranges is None:
        return None
    return (*ranges[3:6], *ranges[0:3])
