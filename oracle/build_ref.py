"""Build the *reference's own* implementations of the hot path into ``oracle/_ref/``.
TEST / BENCHMARK INFRASTRUCTURE ONLY -- never imported by the product.

Run in the build container (needs /root/reference; outputs are git-ignored but travel to the GPU
box with the repository snapshot)::

    python oracle/build_ref.py

For every formula in ``REF_FORMULAS`` it asks the unmodified reference code generator
(``keopscore`` imported from /root/reference -- no reference source is copied into the repo) for

* the CPU code of ``CpuReduc`` (``keopscore/keopscore/mapreduce/cpu/CpuReduc.py:22-103``), compiled
  with the reference's own flags ``g++ -shared -fPIC -O3 -std=c++11 -fopenmp -flto=auto``
  (``keopscore/keopscore/config/base_config.py:266-291``) behind a one-line ``extern "C"`` shim into
  ``oracle/_ref/libkeops_ref_cpu.so``  -> ``cpu_baseline.kind = "reference"`` and
  ``bench.py --impl reference``;
* the CUDA code of ``GpuReduc1D`` (``keopscore/keopscore/mapreduce/gpu/GpuReduc1D.py:47-95``, kernel
  ``GpuConv1DOnDevice``), compiled with NVRTC + ``-use_fast_math`` like the reference does at run time
  (``keopscore/keopscore/binders/nvrtc/nvrtc_jit.cpp:86-97``) into ``oracle/_ref/ref_cuda_<key>.cubin``
  -> the "reference's existing CUDA path on the same GPU" comparator of ``bench.py``.
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_ref")
REFERENCE = "/root/reference"

GAUSS = "Sum_Reduction((Exp(Minus(Sum(Square((Var(0,3,0) - Var(1,3,1)))))) * Var(2,1,1)),0)"
LSE = "Max_SumShiftExp_Reduction((Var(0,1,1) - ((Sum(Square((Var(1,3,0) - Var(2,3,1)))) / IntCst(2)) / Var(3,1,2))),0)"
# key -> (formula, nargs, c dtype, sum scheme)
REF_FORMULAS = {
    "gauss_fwd_f32": (GAUSS, 3, "float", "block_sum"),
    "gauss_dx_f32": ("Grad_WithSavedForward(%s,Var(0,3,0),Var(3,1,0),Var(4,1,0))" % GAUSS, 5, "float", "block_sum"),
    "gauss_dy_f32": ("Grad_WithSavedForward(%s,Var(1,3,1),Var(3,1,0),Var(4,1,0))" % GAUSS, 5, "float", "block_sum"),
    "gauss_db_f32": ("Grad_WithSavedForward(%s,Var(2,1,1),Var(3,1,0),Var(4,1,0))" % GAUSS, 5, "float", "block_sum"),
    "argkmin10_d3_f32": ("ArgKMin_Reduction(Sum(Square((Var(0,3,0) - Var(1,3,1)))),10,0)", 2, "float", "direct_sum"),
    "argkmin10_d64_f32": ("ArgKMin_Reduction(Sum(Square((Var(0,64,0) - Var(1,64,1)))),10,0)", 2, "float", "direct_sum"),
    "lse_f32": (LSE, 4, "float", "block_sum"),
    "lse_f64": (LSE, 4, "double", "block_sum"),
    "gauss16_pm_f32": ("Sum_Reduction(Exp(-Var(3,1,2)*SqDist(Var(0,16,0),Var(1,16,1)))*Var(2,1,1),0)", 4, "float", "block_sum"),
}


def main():
    if not os.path.isdir(REFERENCE):
        print("build_ref: %s is not mounted; keeping existing oracle/_ref" % REFERENCE)
        return 0
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("KEOPS_CACHE_FOLDER", "/tmp/keops_cache_buildref")
    os.environ["CXX"] = "/usr/bin/g++"   # the container's default CXX wrapper lacks libgomp.spec
    sys.path[:0] = [os.path.join(REFERENCE, "keopscore"), ROOT]
    sys.dont_write_bytecode = True
    from keopscore.mapreduce.MapReduce import MapReduce
    from keopscore.mapreduce.cpu.CpuReduc import CpuReduc
    # There is no GPU (nor libcuda.so.1) in the build container: stub the reference's run-time GPU
    # probing so that its *code generator* for GpuReduc1D can be imported.  Only text is produced here.
    import types
    from keopscore.config import cuda_config as _ref_cuda_config
    stub = types.ModuleType("keopscore.utils.gpu_utils")
    stub.custom_cuda_include_fp16_path = lambda: ""
    sys.modules["keopscore.utils.gpu_utils"] = stub
    if _ref_cuda_config.get_cuda_version() is None:
        type(_ref_cuda_config).get_cuda_version = lambda self, out_type="single_value": 12090
    from keopscore.mapreduce.gpu.GpuReduc1D import GpuReduc1D

    manifest = {}
    shim = ['#include <cstddef>']
    for key, (formula, nargs, ctype, scheme) in REF_FORMULAS.items():
        # ---- CPU ----
        cpu = object.__new__(CpuReduc)
        MapReduce.__init__(cpu, formula, [], nargs, ctype, ctype, scheme, 0, 0, 0, False, 1, -1)
        cpu.dimy = cpu.varloader.dimy
        cpu.gencode_filename = key
        cpu.get_code()
        src = os.path.join(OUT, "ref_cpu_%s.cpp" % key)
        with open(src, "w") as f:
            f.write(cpu.code)
        shim.append('#include "ref_cpu_%s.cpp"' % key)
        shim.append('extern "C" int kref_%s(long nx, long ny, int tagI, %s* out, %s** args) { return launch_keops_%s< %s >(nx, ny, tagI, out, args); }'
                    % (key, ctype, ctype, key, ctype))
        rf = cpu.red_formula
        manifest[key] = dict(formula=formula, nargs=nargs, ctype=ctype, tagI=int(rf.tagI), dim=int(rf.dim),
                             dimy=int(cpu.varloader.dimy))
        # ---- CUDA ----
        gpu = object.__new__(GpuReduc1D)
        MapReduce.__init__(gpu, formula, [], nargs, ctype, ctype, scheme, 1, 1, 0, False, 1, 0)
        gpu.dimy = gpu.varloader.dimy
        gpu.get_code()
        cu = os.path.join(OUT, "ref_cuda_%s.cu" % key)
        with open(cu, "w") as f:
            f.write(gpu.code)
        manifest[key]["cuda_block"] = 192   # cuda_block_size, keopscore/keopscore/config/cuda.py:160-162
        manifest[key]["cuda_kernel"] = "GpuConv1DOnDevice"

    shim_path = os.path.join(OUT, "ref_cpu_shim.cpp")
    with open(shim_path, "w") as f:
        f.write("\n".join(shim) + "\n")
    so = os.path.join(OUT, "libkeops_ref_cpu.so")
    cmd = ["/usr/bin/g++", "-shared", "-fPIC", "-O3", "-std=c++11", "-fopenmp", "-flto=auto", "-I", OUT, shim_path, "-o", so]
    print(" ".join(cmd))
    subprocess.check_call(cmd)

    # reference CUDA kernels -> cubins (NVRTC, fast math, like nvrtc_jit.cpp; arch is the B200's)
    if not os.path.exists(os.path.join(ROOT, "keops_b200", "libkb200.so")):
        import __graft_entry__ as ge
        ge.build_runtime()
    from keops_b200.runtime import capi
    for key in REF_FORMULAS:
        code = open(os.path.join(OUT, "ref_cuda_%s.cu" % key)).read()
        cubin = capi.compile_to_cubin(code, ["--use_fast_math", "-lineinfo"], name="ref_cuda_%s.cu" % key)
        with open(os.path.join(OUT, "ref_cuda_%s.cubin" % key), "wb") as f:
            f.write(cubin)
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("build_ref: wrote %s and %d reference CUDA cubins" % (so, len(REF_FORMULAS)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
