"""GPU box: the reference's own block-sparse test body (pykeops/test/test_block_sparse_reduction.py:16-91) on the same seeded
data through (1) the reference's own nvrtc binder and (2) keops_b200's binder at the same seam; the truncation error of
the 4-sigma cutoff is a property of the data, so both must report the same relative error."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("KEOPS_CACHE_FOLDER", "/tmp/kb200_refsuite_keops_cache")
os.makedirs(os.environ["KEOPS_CACHE_FOLDER"], exist_ok=True)
os.environ["CXX"] = "/usr/bin/g++"
os.environ.setdefault("CUDA_HOME", "/usr/local/cuda")
sys.path[:0] = [os.path.join(ROOT, "baseline", "_ref"), ROOT]
import numpy as np
import torch
import pykeops
from pykeops.torch import LazyTensor
from pykeops.torch.cluster import grid_cluster, cluster_ranges_centroids, sort_clusters, from_matrix
from pykeops.common import keops_io
from keops_b200.binder import LoadKeOpsB200

ref_binder = keops_io.keops_binder.get("nvrtc")
print("reference detected CUDA:", pykeops.config.gpu_available, " reference binder:", ref_binder)
device, dtype = torch.device("cuda"), torch.float32
for seed in range(5):
    torch.manual_seed(seed)
    M = N = 5000
    t = torch.linspace(0, 2 * np.pi, M + 1)[:-1]
    x = torch.stack((0.4 + 0.4 * torch.cos(t), 0.5 + 0.3 * torch.sin(t)), dim=1)
    x = (x + 0.01 * torch.randn(x.shape)).to(device=device, dtype=dtype)
    y = torch.randn(N, 2, device=device, dtype=dtype) / 10 + torch.tensor([0.6, 0.6], device=device, dtype=dtype)
    eps = 0.05
    x_labels, y_labels = grid_cluster(x.cpu(), eps).to(device), grid_cluster(y.cpu(), eps).to(device)
    x_ranges, x_centroids, _ = cluster_ranges_centroids(x, x_labels)
    y_ranges, y_centroids, _ = cluster_ranges_centroids(y, y_labels)
    x, x_labels = sort_clusters(x, x_labels)
    y, y_labels = sort_clusters(y, y_labels)
    D = ((x_centroids[:, None, :] - y_centroids[None, :, :]) ** 2).sum(2)
    sigma = 0.05
    keep = D < (4 * sigma) ** 2
    ranges_ij = from_matrix(x_ranges, y_ranges, keep)
    b = torch.randn(N, 1, device=device, dtype=dtype)
    res = {}
    for name, binder in (("reference", ref_binder), ("keops_b200", LoadKeOpsB200)):
        if binder is None:
            continue
        keops_io.keops_binder["nvrtc"] = binder
        x_i, y_j = LazyTensor((x / sigma)[:, None, :]), LazyTensor((y / sigma)[None, :, :])
        K = (-((x_i - y_j) ** 2).sum(dim=2) / 2).exp()
        a_full = K @ b
        K.ranges = ranges_ij
        a_sparse = K @ b
        res[name] = (a_full, a_sparse, (100 * (a_sparse - a_full).abs().sum() / a_full.abs().sum()).item())
    line = "seed %d: " % seed + "  ".join("%s rel_error %.4f%%" % (k, v[2]) for k, v in res.items())
    if len(res) == 2:
        line += "  | sparse results differ by %.2e, dense by %.2e (max abs / max |a|)" % (
            ((res["reference"][1] - res["keops_b200"][1]).abs().max() / res["reference"][1].abs().max()).item(),
            ((res["reference"][0] - res["keops_b200"][0]).abs().max() / res["reference"][0].abs().max()).item())
    print(line)
