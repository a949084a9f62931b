"""Helpers that build the block-sparse ``ranges`` argument -- API of ``pykeops.torch.cluster``
(``pykeops/pykeops/torch/cluster/{grid_cluster,utils,matrix}.py``): bin points into cells, sort them so
that clusters are contiguous, and turn a boolean cluster-vs-cluster mask into the 6-tuple
``(ranges_i, slices_i, redranges_j, ranges_j, slices_j, redranges_i)`` consumed by
``Genred(..., ranges=...)`` (``pykeops/pykeops/torch/generic/generic_red.py:581-639``).

Pure torch tensor code (runs on whatever device the inputs live on); no kernel of the hot path here.
"""
from __future__ import annotations

import torch

__all__ = ["grid_cluster", "sort_clusters", "cluster_ranges", "cluster_centroids", "cluster_ranges_centroids",
           "from_matrix", "swap_axes"]


def grid_cluster(x: torch.Tensor, size) -> torch.Tensor:
    """Labels in a compact range [0, C) of the cubic cell (edge ``size``) that contains each point of
    ``x`` (M, D), D <= 3; cells are numbered in lexicographic order of their integer coordinates."""
    if x.dim() != 2 or not 1 <= x.shape[1] <= 3:
        raise NotImplementedError("grid_cluster handles (M, D) point clouds with D in {1, 2, 3}")
    with torch.no_grad():
        cell = torch.floor((x - x.min(dim=0, keepdim=True).values) / size).to(torch.int64)
        key = cell[:, 0]
        for k in range(1, x.shape[1]):
            key = key * 1024 + cell[:, k]                 # same packing radix as the reference (2**10 per axis)
        _, lab = torch.unique(key, sorted=True, return_inverse=True)
    return lab.to(torch.int32)


def sort_clusters(x, lab):
    """Sort points (a tensor or a tuple of tensors) by label so that clusters are contiguous."""
    lab_sorted, perm = torch.sort(lab.view(-1), stable=True)
    if isinstance(x, tuple):
        return tuple(a[perm] for a in x), lab_sorted
    return x[perm], lab_sorted


def cluster_ranges(lab, Nlab=None):
    """(C, 2) int64 array of [start, end) positions of each cluster in the label-sorted point list."""
    if Nlab is None:
        Nlab = torch.bincount(lab.to(torch.int64))
    ends = torch.cumsum(Nlab.to(torch.int64), 0)
    starts = ends - Nlab.to(torch.int64)
    return torch.stack((starts, ends), dim=1).to(torch.int64)


def cluster_centroids(x, lab, Nlab=None, weights=None, weights_c=None):
    """(C, D) weighted means of the points of each cluster."""
    lab = lab.to(torch.int64)
    C = int(lab.max().item()) + 1 if Nlab is None else int(Nlab.shape[0])
    if weights is None:
        w = torch.ones(x.shape[0], dtype=x.dtype, device=x.device)
    else:
        w = weights.view(-1).to(x.dtype)
    tot = torch.zeros(C, dtype=x.dtype, device=x.device).index_add_(0, lab, w) if weights_c is None else weights_c.to(x.dtype)
    acc = torch.zeros(C, x.shape[1], dtype=x.dtype, device=x.device).index_add_(0, lab, x * w[:, None])
    return acc / tot[:, None]


def cluster_ranges_centroids(x, lab, weights=None, min_weight=1e-9):
    """``(ranges, centroids, cluster weights)`` for points ``x`` with integer labels ``lab``."""
    lab64 = lab.to(torch.int64)
    Nlab = torch.bincount(lab64).to(torch.float32)
    if weights is None:
        return cluster_ranges(lab64, Nlab), cluster_centroids(x, lab64, Nlab), Nlab
    w = weights.clone().view(-1)
    w[w <= min_weight] = min_weight
    w_c = torch.bincount(lab64, weights=w).view(-1)
    return cluster_ranges(lab64, Nlab), cluster_centroids(x, lab64, Nlab, weights=w, weights_c=w_c), w_c


def _one_side(ranges_out, ranges_red, keep):
    """For every output cluster, the list of kept reduction clusters -> (slices, redranges), with runs of
    adjacent kept clusters whose ranges touch fused into one range (fewer, longer tiles)."""
    C_out, C_red = keep.shape
    if C_red == 0 or C_out == 0:
        return (torch.zeros(C_out, dtype=torch.int64, device=keep.device),
                torch.zeros(0, 2, dtype=ranges_red.dtype, device=keep.device))
    k = keep.to(torch.bool)
    prev = torch.zeros_like(k)
    prev[:, 1:] = k[:, :-1]
    touch = torch.ones(C_red, dtype=torch.bool, device=keep.device)
    touch[1:] = ranges_red[1:, 0] == ranges_red[:-1, 1]
    starts = k & ~(prev & touch[None, :])                 # a run begins here
    nxt = torch.zeros_like(k)
    nxt[:, :-1] = k[:, 1:]
    touch_next = torch.ones(C_red, dtype=torch.bool, device=keep.device)
    touch_next[:-1] = touch[1:]
    ends = k & ~(nxt & touch_next[None, :])               # a run ends here
    rs, cs = torch.nonzero(starts, as_tuple=True)          # row-major order: sorted by output cluster, then column
    re_, ce = torch.nonzero(ends, as_tuple=True)
    red = torch.stack((ranges_red[cs, 0], ranges_red[ce, 1]), dim=1)    # keeps the caller's integer dtype, like matrix.py:118-130
    slices = torch.cumsum(starts.sum(dim=1), 0).to(torch.int64)
    return slices, red


def from_matrix(ranges_i, ranges_j, keep):
    """Boolean (Ci, Cj) mask between i-clusters and j-clusters -> the 6-tuple of ``ranges``."""
    keep = keep.to(ranges_i.device)
    slices_i, redranges_j = _one_side(ranges_i, ranges_j, keep)
    slices_j, redranges_i = _one_side(ranges_j, ranges_i, keep.t())
    return ranges_i, slices_i, redranges_j, ranges_j, slices_j, redranges_i


def swap_axes(ranges):
    """Ranges of the transposed problem (i <-> j)."""
    if ranges is None:
        return None
    return (*ranges[3:6], *ranges[0:3])
