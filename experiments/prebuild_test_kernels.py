"""Cross-compile (NVRTC, sm_100a, no GPU needed) the kernels the GPU test-suite uses into the in-tree cubin cache, so that a
fresh GPU box does not spend its test time in the JIT.  python experiments/prebuild_test_kernels.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keops_b200.binder import get_binder
from keops_b200.codegen import conv_tc, knn_tc
from keops_b200.runtime import plan
from tests import util

t0 = time.time()
n = 0
for name in util.golden_names():
    meta, args, _ = util.load_case(name)
    for fm in (1, 0):
        flags = util.default_flags(meta["formula"])
        flags["use_fast_math"] = fm
        try:
            b = get_binder(meta["formula"], meta.get("aliases") or [], meta["nargs"], meta["dtype"], flags)
            if not b.ks.is_zero:
                plan.get_cubin(b.ks, bool(fm) and meta["dtype"] == "float32")
                n += 1
        except Exception as e:                      # a case the engine rejects is reported by the tests themselves
            print("skip", name, str(e)[:80])
for D, DB, lse in ((16, 1, False), (8, 3, False), (5, 1, False), (6, 2, False), (13, 2, False), (20, 1, False), (16, 2, False),
                   (21, 4, False), (22, 1, False), (64, 2, False), (95, 1, False), (128, 2, False), (300, 1, False),
                   (64, 1, False), (16, 1, True), (8, 2, True), (64, 1, True), (128, 3, True), (16, 2, True)):
    conv_tc.ConvTensorCore(D, DB, -1, lse=lse)
    n += 1
for kind, D, K in (("ArgKMin", 64, 10), ("KMin_ArgKMin", 100, 5), ("KMin", 16, 3), ("ArgKMin", 32, 1), ("ArgKMin", 128, 10),
                   ("ArgKMin", 61, 4), ("KMin_ArgKMin", 32, 5), ("KMin", 100, 3)):
    knn_tc.KnnTensorCore(D, K, kind, True, device=-1)
    n += 1
print("prebuilt %d kernels in %.0f s" % (n, time.time() - t0))
