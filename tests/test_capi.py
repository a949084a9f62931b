"""CPU: the C-ABI library loads without a GPU and exports every symbol include/kb200.h declares;
NVRTC cross-compiles generated kernels to sm_100a cubins here (no compute calls)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "kb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(kb200_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from keops_b200.runtime import capi
    lib = ctypes.CDLL(capi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), "libkb200.so does not export %s" % s
    assert set(syms) == set(capi.EXPORTED_SYMBOLS)
    assert lib.kb200_abi_version() == 4


def test_compute_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from keops_b200.runtime import capi
    with pytest.raises(capi.KB200Error) as e:
        capi.device_info(0)
    assert e.value.status == capi.KB200_ENOCUDA


def test_nvrtc_reports_the_compiler_log():
    from keops_b200.runtime import capi
    with pytest.raises(capi.KB200CompileError) as e:
        capi.compile_to_cubin("__global__ void k() { this is not cuda; }")
    assert "error" in str(e.value)


def test_generated_kernel_compiles_to_sm100a_with_packed_fp32():
    from keops_b200.formulas.parser import parse_reduction
    from keops_b200.codegen.emit import generate
    from keops_b200.runtime import capi, plan
    red = parse_reduction("Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),0)")
    ks = generate(red, "float32", 3, plan.default_config(red, "float32"))
    cubin = capi.compile_to_cubin(ks.source, plan.nvrtc_options(True))
    assert cubin[:4] == b"\x7fELF"
    path = "/tmp/kb200_test_%d.cubin" % os.getpid()
    open(path, "wb").write(cubin)
    try:
        sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    except FileNotFoundError:
        pytest.skip("cuobjdump not available")
    finally:
        pass
    os.remove(path)
    assert "sm_100a" in sass or "SM100a" in sass or "EF_CUDA_SM100" in sass
    for mnemonic in ("FFMA2", "FADD2", "FMUL2", "MUFU.EX2", "LDGSTS", "LDS.128"):
        assert mnemonic in sass, mnemonic
    assert "LDL" not in sass and "STL" not in sass  # no local-memory spills


def test_binder_refuses_cpu_backend():
    from keops_b200.binder import LoadKeOpsB200
    flags = dict(dtype_acc="auto", sum_scheme="auto", use_fast_math=1, enable_chunks=1, multVar_highdim=0)
    with pytest.raises(ValueError):
        LoadKeOpsB200(0, 0, 0, False, -1, "Sum_Reduction(Var(0,1,0)*Var(1,1,1),0)", [], 2, "float32", "torch", flags)


def test_double_accumulator_and_kahan_kernels_compile():
    """``dtype_acc="float64"`` (fp32 data, fp64 accumulators: pykeops/pykeops/common/parse_type.py:195-207,
    keopscore/keopscore/formulas/reductions/Sum_Reduction.py:10-50) and ``sum_scheme="kahan_scheme"``
    (keopscore/keopscore/formulas/reductions/sum_schemes.py:74-86) generate valid sm_100a kernels with fp64 adds
    (respectively the compensation subtractions) in the reduction."""
    from keops_b200.formulas.parser import parse_reduction
    from keops_b200.codegen.emit import generate
    from keops_b200.runtime import capi, plan
    red = parse_reduction("Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,2,1)),0)")
    cfg = plan.default_config(red, "float32")
    plain = generate(red, "float32", 3, cfg)
    acc64 = generate(red, "float32", 3, cfg, dtype_acc="float64")
    kahan = generate(red, "float32", 3, cfg, sum_scheme="kahan_scheme")
    assert "double" in acc64.source and acc64.source != plain.source and kahan.source != plain.source
    for ks in (acc64, kahan):
        assert capi.compile_to_cubin(ks.source, plan.nvrtc_options(True))[:4] == b"\x7fELF"
    with pytest.raises(ValueError):
        generate(red, "float32", 3, cfg, sum_scheme="nope")


@pytest.mark.parametrize("head", ["Sum", "LSE", "LSEW"])
def test_sum_scheme_x_dtype_acc_matrix_compiles(head):
    """Every (sum_scheme, dtype_acc) pair that ``get_optional_flags`` accepts
    (pykeops/pykeops/common/parse_type.py:195-207) yields a kernel that compiles, for the sum-like reductions whose
    accumulator the options act on (reference: ``acc`` is declared in dtypeacc for every reduction,
    keopscore/keopscore/mapreduce/MapReduce.py:74; Kahan for LSE: Max_SumShiftExpWeight_Reduction.py:65-85)."""
    from keops_b200.formulas.parser import parse_reduction
    from keops_b200.codegen.emit import generate
    from keops_b200.runtime import capi, plan
    f = {"Sum": "Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,2,1)),0)",
         "LSE": "Max_SumShiftExp_Reduction(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1))))),0)",
         "LSEW": "Max_SumShiftExpWeight_Reduction(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1))))),0,Var(2,2,1))"}[head]
    red = parse_reduction(f)
    cfg = plan.default_config(red, "float32")
    seen = set()
    for scheme in ("block_sum", "direct_sum", "kahan_scheme"):
        for acc in ("auto", "float32", "float64"):
            ks = generate(red, "float32", 3, cfg, sum_scheme=scheme, dtype_acc=acc)
            assert capi.compile_to_cubin(ks.source, plan.nvrtc_options(True))[:4] == b"\x7fELF", (scheme, acc)
            seen.add(ks.hash)
            if acc == "float64":
                assert "double accd" in ks.source, (scheme, acc)
    assert len(seen) == 6        # auto == float32; every other combination is a different kernel


def test_tensor_core_kernels_are_blackwell_native_sass():
    """The hand-written tensor-core kernels really are tcgen05 / TMA / TMEM code: their sm_100a SASS carries the 5th-generation
    MMA (UTCHMMA), TMA tensor loads (UTMALDG), tensor-memory loads / stores (LDTM / STTM) and the MMA-completion barrier
    (UTCBAR) -- not mma.sync (HMMA) recompiled for a newer architecture.  (Mnemonics: /opt/skills/guides/B200_PROFILING.md.)"""
    from keops_b200.codegen import conv_tc, knn_tc
    blobs = {"kb200_knn_tc": knn_tc.KnnTensorCore(64, 10, "ArgKMin", True, device=-1).cubin,
             "kb200_conv_tc": conv_tc.ConvTensorCore(16, 1, device=-1).cubin}
    for name, cubin in blobs.items():
        path = "/tmp/kb200_tc_%d.cubin" % os.getpid()
        open(path, "wb").write(cubin)
        try:
            sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
        except FileNotFoundError:
            pytest.skip("cuobjdump not available")
        finally:
            os.remove(path)
        for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS"):
            assert mnemonic in sass, (name, mnemonic)
        assert "HMMA." not in sass.replace("UTCHMMA", ""), name            # no legacy mma.sync tensor-core path
        if name == "kb200_conv_tc":
            assert "STTM" in sass, name                                     # the row block is written to tensor memory once
