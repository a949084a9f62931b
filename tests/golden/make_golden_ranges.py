"""Golden vectors for the block-sparse ``ranges`` mode, produced by the REAL reference (CPU backend,
``CpuReduc_ranges``, ``keopscore/keopscore/mapreduce/cpu/CpuReduc_ranges.py:173-228``).

Run where /root/reference is mounted::

    CXX=/usr/bin/g++ KEOPS_CACHE_FOLDER=/tmp/keops_cache \
    PYTHONPATH=/root/reference/keopscore:/root/reference/pykeops python tests/golden/make_golden_ranges.py

Each case stores the reduction string, inputs, the 6 range arrays
``(ranges_i, slices_i, redranges_j, ranges_j, slices_j, redranges_i)`` (``generic_red.py:581-639``) and
the raw binder output.  Row ranges cover every output row (uncovered rows are left uninitialised by
the reference).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ranges")
sys.path.insert(0, HERE)
from make_golden import flags  # noqa: E402

G = "Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),%d)"
CASES = [
    ("gauss_sum", G % 0, [(3, 0), (3, 1), (1, 1)]),
    ("gauss_sum_axis0", G % 1, [(3, 0), (3, 1), (1, 1)]),
    ("argkmin3", "ArgKMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),3,0)", [(3, 0), (3, 1)]),
    ("min_argmin_axis0", "Min_ArgMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),1)", [(3, 0), (3, 1)]),
    ("lse", "Max_SumShiftExp_Reduction((Var(0,1,1) - Sum(Square((Var(1,3,0) - Var(2,3,1))))),0)", [(1, 1), (3, 0), (3, 1)]),
]


def random_partition(rng, n, nparts):
    cuts = np.sort(rng.choice(np.arange(1, n), size=nparts - 1, replace=False))
    edges = np.concatenate([[0], cuts, [n]])
    return np.stack([edges[:-1], edges[1:]], axis=1).astype(np.int64)


def random_ranges(rng, nx, ny, ncx, ncy):
    """Block-sparse pattern between ``ncx`` row clusters and ``ncy`` column clusters (every row cluster
    sees at least one column cluster; adjacent clusters are often both kept -> exercises merging)."""
    rx, ry = random_partition(rng, nx, ncx), random_partition(rng, ny, ncy)
    keep = rng.random((ncx, ncy)) < 0.45
    keep[np.arange(ncx), rng.integers(0, ncy, ncx)] = True
    keep[rng.integers(0, ncx, ncy), np.arange(ncy)] = True

    def one_side(r_out, r_red, mask):
        slices, red = [], []
        for k in range(len(r_out)):
            red.extend(r_red[l] for l in np.nonzero(mask[k])[0])
            slices.append(len(red))
        return r_out, np.asarray(slices, np.int64), np.asarray(red, np.int64).reshape(-1, 2)

    ri, si, rrj = one_side(rx, ry, keep)
    rj, sj, rri = one_side(ry, rx, keep.T)
    return ri, si, rrj, rj, sj, rri


def main():
    import torch
    from pykeops.common.keops_io import keops_binder
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20260202)
    index = []
    for name, formula, spec in CASES:
        for dtype in ("float32", "float64"):
            nx, ny = 211, 173
            args = [rng.standard_normal((nx if cat == 0 else ny, dim)).astype(dtype) for dim, cat in spec]
            ranges = random_ranges(rng, nx, ny, 7, 9)
            targs = [torch.from_numpy(a) for a in args]
            tr = tuple(torch.from_numpy(r) for r in ranges)
            b = keops_binder["cpp"](0, 0, 0, True, -1, formula, [], len(targs), dtype, "torch", flags(formula))
            out = b.import_module().genred_pytorch(targs[0].device, tr, nx, ny, 0, None, *targs).numpy().copy()
            meta = dict(name="%s_%s" % (name, dtype), formula=formula, dtype=dtype, nx=nx, ny=ny, nargs=len(args),
                        dimout=int(b.dimout), tagIJ=int(b.tagIJ))
            np.savez_compressed(os.path.join(OUT, meta["name"] + ".npz"), out=out, meta=json.dumps(meta),
                                **{"arg%d" % k: a for k, a in enumerate(args)},
                                **{"range%d" % k: r for k, r in enumerate(ranges)})
            index.append(meta)
            print("%-32s out %s" % (meta["name"], out.shape))
    with open(os.path.join(OUT, "index.json"), "w") as f:
        json.dump(index, f, indent=1)


if __name__ == "__main__":
    main()
