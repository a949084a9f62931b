"""Writes profiles/r02_sass_excerpts.txt: static instruction histograms + one example per key mnemonic of the kernels the
bench runs (cross-compiled for sm_100a here, no GPU needed) -- the evidence that the generated kernels use the packed FP32
instructions of sm_100 and the hand-written ones tcgen05 / TMA / TMEM."""
import collections
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keops_b200.binder import get_binder
from keops_b200.codegen import conv_tc, knn_tc
from keops_b200.codegen.emit import generate
from keops_b200.runtime import plan

G = "Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),0)"
KNN3 = "ArgKMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),10,0)"


def sass(cubin: bytes, fun: str) -> str:
    with tempfile.NamedTemporaryFile(suffix=".cubin", delete=False) as f:
        f.write(cubin)
    try:
        return subprocess.run(["cuobjdump", "-sass", "-fun", fun, f.name], capture_output=True, text=True).stdout
    finally:
        os.unlink(f.name)


def report(out, title, text, keys):
    ops = collections.Counter()
    for line in text.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            ops[m.group(1)] += 1
    out.write("== %s, sm_100a\n" % title)
    out.write("instruction histogram (static): " + ", ".join("%s x%d" % kv for kv in ops.most_common(28)) + "\n")
    for k in keys:
        hits = [l for l in text.splitlines() if re.search(r"\b%s" % re.escape(k), l)]
        out.write("  %-10s %d occurrences; e.g. %s\n" % (k, len(hits), hits[0].strip()[:110] if hits else "-"))


def main():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_sass_excerpts.txt")
    with open(path, "w") as out:
        bd = get_binder(G, [], 3, "float32")
        ks = generate(bd.red, "float32", 3, bd.ks.cfg, use_fast_math=True, sum_scheme="block_sum", dtype_acc="float32", prescale=True)
        report(out, "generated kb200_reduce, Gaussian conv D=3 fp32, pre-scaled packed variant (%s)" % ks.cfg.key(),
               sass(plan.get_cubin(ks, True), "kb200_reduce"), ("FFMA2", "FADD2", "FMUL2", "MUFU.EX2", "LDS.128", "LDGSTS", "LDL", "STL"))
        kf = get_binder(KNN3, [], 2, "float32")._kfilter_source()
        report(out, "generated kb200_reduce, filtered ArgKMin K=10 D=3 (%s)" % kf.cfg.key(),
               sass(plan.get_cubin(kf, True), "kb200_reduce"), ("FFMA2", "FMNMX3", "FSETP", "LDS.128", "LDGSTS", "LDL", "STL"))
        eng = knn_tc.KnnTensorCore(64, 10, "ArgKMin", True, device=-1)
        report(out, "hand-written kb200_knn_filter (K-min filter, D=64, resident row operand)", sass(eng.cubin, "kb200_knn_filter"),
               ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "FMNMX", "ATOMS", "HMMA."))
        eng = knn_tc.KnnTensorCore(784, 10, "ArgKMin", True, device=-1)
        report(out, "hand-written kb200_knn_filter (K-min filter, D=784, streamed k-blocks)", sass(eng.cubin, "kb200_knn_filter"),
               ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "HMMA."))
        eng = conv_tc.ConvTensorCore(16, 1, device=-1)
        report(out, "hand-written kb200_conv_tc (Gaussian conv D=16, TMEM-resident A operand)", sass(eng.cubin, "kb200_conv_tc"),
               ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "MUFU.EX2", "FFMA", "HMMA."))
    print(open(path).read())


if __name__ == "__main__":
    main()
