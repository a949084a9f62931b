"""Run the reference's OWN test files (``baseline/_ref/pykeops/test``, staged by ``baseline/stage_ref.py`` from the
unmodified getkeops/keops v2.3) on the GPU with this engine registered at the reference's plugin seam:
``pykeops.common.keops_io.keops_binder["nvrtc"] = LoadKeOpsB200`` (``pykeops/pykeops/common/keops_io/__init__.py:6-13``).

Everything above the seam -- ``pykeops.torch.Genred`` / ``LazyTensor`` / ``KernelSolve``, their autograd functions, pre- and
post-processing, the ``cluster`` helpers, and the test bodies with their own tolerances -- is the reference's code, unmodified;
everything below it (formula -> kernel -> launch) is ``keops_b200``.  Run in its own process (``tests/
test_reference_frontend_gpu.py`` does) because other tests install ``keops_b200`` itself under the name ``pykeops``.

    python tests/reference_suite.py [--json out.json] [pytest args...]

Prints one JSON object: per-test outcomes, how many launches went through ``LoadKeOpsB200.genred_pytorch``, and whether the
reference detected CUDA on its own (it then also builds its own nvrtc binder, which is simply never called).
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

# reference test files whose bodies exercise the hot path on torch CUDA tensors
FILES = [
    "pykeops/torch/test_install.py",
    "pykeops/test/test_torch.py",
    "pykeops/test/test_lazytensor_gaussian_batch.py",
    "pykeops/test/test_block_sparse_reduction.py",
    "pykeops/test/test_lazytensor_gaussian_inplace.py",
    "pykeops/test/test_lazytensor_grad.py",
    "pykeops/test/test_lazytensor_clamp.py",
    "pykeops/test/test_contiguous_torch.py",
    "pykeops/test/test_transpose.py",
    "pykeops/test/test_sinc.py",
    "pykeops/test/test_torch_func.py",
    "pykeops/test/test_torch_func_logsumexp.py",
    "pykeops/test/test_vmap.py",
    "pykeops/test/test_chunks.py",
    "pykeops/test/test_chunks_ranges.py",
    "pykeops/test/test_finalchunks.py",
    "pykeops/test/test_finalchunks_ranges.py",
    "pykeops/test/test_lazytensor_tensordot.py",
    "pykeops/test/test_kron.py",
    # the reference's numpy front-end and host-tensor calls: same seam, ``genred_numpy`` / ``tagHostDevice == 0``
    "pykeops/test/test_numpy.py",
    "pykeops/test/test_contiguous_numpy.py",
    "pykeops/test/test_lazytensor_gaussian_fromhost.py",
    "pykeops/test/test_lazytensor_gaussian_numpy_inplace.py",
    "pykeops/test/test_gpu_cpu.py",
    "pykeops/numpy/test_install.py",
    # complex LazyTensors (the reference's ComplexLazyTensor front-end over this engine's interleaved (re, im) operations)
    "pykeops/test/test_complex.py",
    "pykeops/test/test_complex_numpy.py",
    # the reference's per-operation harness: every operation of keopscore.formulas.maths (forward and, where it has one,
    # gradient) as a Genred Sum reduction against the operation's torch equivalent
    "keopscore/test/test_op.py",
]
# nothing is deselected any more (TensorDot / Kron are lowered like every other operation); every test in FILES must pass
DESELECT = []


def prebuild_kernels(argv):
    """GPU-less container only (``--prebuild``): run the same reference test files on the reference's own CPU backend while a
    recording shim at ``keops_binder["cpp"]`` hands every (formula, aliases, dtype, flags) the front-end asks for to
    ``LoadKeOpsB200`` as the GPU request it will be on the box -- which generates the kernel and cross-compiles it (NVRTC,
    sm_100a) into the in-tree cubin cache.  The GPU box then runs the suite without spending its minutes in the JIT."""
    import pykeops
    from pykeops.common import keops_io
    from keops_b200 import binder as kb
    from keops_b200.runtime import plan
    real_cpp = keops_io.keops_binder["cpp"]
    seen, failed = set(), []

    def shim(tagCPUGPU, tag1D2D, tagHostDevice, use_ranges, device_id, formula, aliases, nargs, dtype, lang, flags):
        key = (formula, tuple(aliases), nargs, dtype, tuple(sorted(flags.items())), bool(use_ranges))
        if key not in seen:
            seen.add(key)
            try:
                for t2d in (0, 1):
                    b = kb.LoadKeOpsB200(1, t2d, 1, use_ranges, -1, formula, list(aliases), nargs, dtype, lang, dict(flags))
                fast = bool(b.params.use_fast_math) and dtype == "float32"
                if not b.ks.is_zero:
                    plan.get_cubin(b.ks, fast)
                    if use_ranges:
                        from keops_b200.codegen.emit import generate
                        p = b.params
                        ksr = generate(b.red, p.dtype, p.nargs, b.ks.cfg, use_fast_math=p.use_fast_math,
                                       sum_scheme=p.sum_scheme, dtype_acc=p.dtype_acc, ranges=True)
                        plan.get_cubin(ksr, fast)
            except Exception as e:             # reported; the GPU run will show the same failure as a test failure
                failed.append((formula, repr(e)[:200]))
                if os.environ.get("KB200_PREBUILD_TRACE"):
                    import traceback
                    traceback.print_exc()
        return real_cpp(tagCPUGPU, tag1D2D, tagHostDevice, use_ranges, device_id, formula, aliases, nargs, dtype, lang, flags)

    shim.reset = getattr(real_cpp, "reset", lambda: None)
    keops_io.keops_binder["cpp"] = shim
    import pytest
    files = [os.path.join(REF, f) for f in FILES if os.path.exists(os.path.join(REF, f))]
    rc = pytest.main(files + ["-q", "-p", "no:cacheprovider", "--no-header"] + (["-k", " and ".join("not " + d for d in DESELECT)] if DESELECT else []) + argv)
    print("REFERENCE_SUITE_PREBUILD " + json.dumps({"formulas": len(seen), "engine_rejected": failed, "pytest_rc": int(rc)}))
    return 0 if not failed else 1


def main(argv):
    out_json = None
    if "--json" in argv:
        k = argv.index("--json")
        out_json = argv[k + 1]
        argv = argv[:k] + argv[k + 2:]
    prebuild = "--prebuild" in argv
    if prebuild:
        argv = [a for a in argv if a != "--prebuild"]
    if not os.path.isdir(os.path.join(REF, "pykeops")):
        print(json.dumps({"unavailable": "baseline/_ref is not staged (python baseline/stage_ref.py)"}))
        return 3
    os.environ.setdefault("KEOPS_CACHE_FOLDER", os.path.join(tempfile.gettempdir(), "kb200_refsuite_keops_cache"))
    os.makedirs(os.environ["KEOPS_CACHE_FOLDER"], exist_ok=True)
    if os.path.exists("/usr/bin/g++"):
        os.environ["CXX"] = "/usr/bin/g++"      # the image's default CXX wrapper lacks libgomp.spec
    os.environ.setdefault("CUDA_HOME", "/usr/local/cuda")
    sys.path[:0] = [REF, ROOT]
    sys.dont_write_bytecode = True
    import torch
    if prebuild:
        return prebuild_kernels(argv)
    if not torch.cuda.is_available():
        print(json.dumps({"unavailable": "no CUDA device (this engine has no CPU fallback)"}))
        return 3
    import pykeops                                      # the real reference
    assert os.path.realpath(pykeops.__file__).startswith(os.path.realpath(REF)), pykeops.__file__
    ref_detected_cuda = bool(pykeops.config.gpu_available)
    from keops_b200 import binder as kb
    kb.register_into_pykeops()
    from pykeops.common import keops_io
    assert keops_io.keops_binder["nvrtc"] is kb.LoadKeOpsB200

    calls = {"launches": 0, "engines": {}}
    inner = kb.LoadKeOpsB200.genred

    def counted(self, *a, **k):
        res = inner(self, *a, **k)
        calls["launches"] += 1
        eng = (getattr(self, "last_launch", None) or {}).get("engine") or "cuda-core"
        calls["engines"][eng] = calls["engines"].get(eng, 0) + 1
        return res

    kb.LoadKeOpsB200.genred = counted
    kb.LoadKeOpsB200.genred_pytorch = counted

    import pytest

    class Collect:
        def __init__(self):
            self.outcomes = {}

        def pytest_runtest_setup(self, item):
            # several reference tests draw unseeded data against fixed tolerances (the block-sparse cutoff test sits at
            # 0.15-0.25 % of its 0.2 % bar depending on the draw): seed every test so that a run is reproducible
            import numpy as np
            torch.manual_seed(0)
            np.random.seed(0)

        def pytest_runtest_logreport(self, report):
            if report.when == "call" or (report.when == "setup" and report.outcome != "passed"):
                self.outcomes[report.nodeid.split("baseline/_ref/")[-1]] = report.outcome

    col = Collect()
    files = [os.path.join(REF, f) for f in FILES if os.path.exists(os.path.join(REF, f))]
    args = files + ["-q", "-p", "no:cacheprovider", "--no-header", "--continue-on-collection-errors"] + \
        (["-k", " and ".join("not " + d for d in DESELECT)] if DESELECT else []) + argv
    rc = pytest.main(args, plugins=[col])
    summary = {
        "reference_version": pykeops.__version__, "reference_path": os.path.relpath(os.path.dirname(pykeops.__file__), ROOT),
        "reference_detected_cuda_itself": ref_detected_cuda, "binder": "keops_b200.binder.LoadKeOpsB200",
        "device": torch.cuda.get_device_name(0), "pytest_rc": int(rc),
        "passed": sum(v == "passed" for v in col.outcomes.values()),
        "failed": sorted(k for k, v in col.outcomes.items() if v == "failed"),
        "skipped": sorted(k for k, v in col.outcomes.items() if v == "skipped"),
        "launches_through_LoadKeOpsB200": calls["launches"], "engines": calls["engines"], "deselected": DESELECT,
        "outcomes": col.outcomes,
    }
    line = json.dumps(summary)
    if out_json:
        with open(out_json, "w") as f:
            f.write(line + "\n")
    print("REFERENCE_SUITE_JSON " + line)
    return int(rc)


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
