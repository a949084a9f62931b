// kb200_runtime.cpp -- implementation of include/kb200.h (libkb200.so).
//
// Thin, allocation-free launch path over the CUDA driver API plus an NVRTC front door that
// always targets sm_100a.  libcuda / libnvrtc are resolved lazily with dlopen so the library
// (and its exported symbols) can be loaded on a GPU-less build machine.
//
// Build: g++ -O2 -std=c++17 -shared -fPIC -I include -I /usr/local/cuda/include \
//            keops_b200/csrc/kb200_runtime.cpp -o keops_b200/libkb200.so -ldl
#include "kb200.h"

#include <cuda.h>
#include <nvrtc.h>
#include <dlfcn.h>
#include <unistd.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

// ---------------------------------------------------------------------------------------------
// lazily bound driver / NVRTC entry points
// ---------------------------------------------------------------------------------------------
struct Driver {
  void* lib = nullptr;
  CUresult (*Init)(unsigned) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*DeviceTotalMem)(size_t*, CUdevice) = nullptr;
  CUresult (*PrimaryCtxRetain)(CUcontext*, CUdevice) = nullptr;
  CUresult (*PrimaryCtxRelease)(CUdevice) = nullptr;
  CUresult (*CtxPushCurrent)(CUcontext) = nullptr;
  CUresult (*CtxPopCurrent)(CUcontext*) = nullptr;
  CUresult (*CtxGetCurrent)(CUcontext*) = nullptr;
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
  CUresult (*OccupancyMaxActiveBlocks)(int*, CUfunction, int, size_t) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                           unsigned, CUstream, void**, void**) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
  int status = -1;  // -1 not tried, 0 ok, >0 error code
};

struct Nvrtc {
  void* lib = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*,
                               const char* const*) = nullptr;
  nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  const char* (*GetErrorString)(nvrtcResult) = nullptr;
  int status = -1;
};

Driver g_drv;
Nvrtc g_rtc;
std::mutex g_mu;

template <class F>
bool bind(void* lib, const char* name, F& out) {
  out = reinterpret_cast<F>(dlsym(lib, name));
  return out != nullptr;
}

int driver() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_drv.status >= 0) return g_drv.status ? fail(g_drv.status, "CUDA driver unavailable") : 0;
  void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    g_drv.status = KB200_ENOCUDA;
    return fail(KB200_ENOCUDA, "cannot load libcuda.so.1: %s", dlerror());
  }
  Driver& d = g_drv;
  d.lib = lib;
  bool ok = bind(lib, "cuInit", d.Init) && bind(lib, "cuDeviceGet", d.DeviceGet) &&
            bind(lib, "cuDeviceGetAttribute", d.DeviceGetAttribute) &&
            bind(lib, "cuDeviceTotalMem_v2", d.DeviceTotalMem) &&
            bind(lib, "cuDevicePrimaryCtxRetain", d.PrimaryCtxRetain) &&
            bind(lib, "cuDevicePrimaryCtxRelease_v2", d.PrimaryCtxRelease) &&
            bind(lib, "cuCtxPushCurrent_v2", d.CtxPushCurrent) &&
            bind(lib, "cuCtxPopCurrent_v2", d.CtxPopCurrent) &&
            bind(lib, "cuCtxGetCurrent", d.CtxGetCurrent) &&
            bind(lib, "cuModuleLoadData", d.ModuleLoadData) &&
            bind(lib, "cuModuleUnload", d.ModuleUnload) &&
            bind(lib, "cuModuleGetFunction", d.ModuleGetFunction) &&
            bind(lib, "cuFuncGetAttribute", d.FuncGetAttribute) &&
            bind(lib, "cuOccupancyMaxActiveBlocksPerMultiprocessor", d.OccupancyMaxActiveBlocks) &&
            bind(lib, "cuLaunchKernel", d.LaunchKernel) &&
            bind(lib, "cuGetErrorString", d.GetErrorString) &&
            bind(lib, "cuFuncSetAttribute", d.FuncSetAttribute) &&
            bind(lib, "cuTensorMapEncodeTiled", d.TensorMapEncodeTiled);
  if (!ok) {
    g_drv.status = KB200_ENOCUDA;
    return fail(KB200_ENOCUDA, "libcuda.so.1 lacks a required symbol");
  }
  CUresult r = d.Init(0);
  if (r != CUDA_SUCCESS) {
    g_drv.status = KB200_ENOCUDA;
    return fail(KB200_ENOCUDA, "cuInit failed with CUresult %d", (int)r);
  }
  g_drv.status = 0;
  return 0;
}

int nvrtc() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_rtc.status >= 0) return g_rtc.status ? fail(g_rtc.status, "NVRTC unavailable") : 0;
  // The toolkit's own NVRTC first: a process that imported torch already holds torch's bundled
  // (older) libnvrtc.so.12, which a bare soname lookup would silently return.
  const char* names[] = {"/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so"};
  void* lib = nullptr;
  const char* env = getenv("KB200_NVRTC_PATH");
  if (env) lib = dlopen(env, RTLD_NOW);
  for (size_t i = 0; !lib && i < sizeof names / sizeof *names; ++i) lib = dlopen(names[i], RTLD_NOW);
  if (!lib) {
    g_rtc.status = KB200_ENONVRTC;
    return fail(KB200_ENONVRTC, "cannot load libnvrtc: %s", dlerror());
  }
  Nvrtc& n = g_rtc;
  n.lib = lib;
  bool ok = bind(lib, "nvrtcCreateProgram", n.CreateProgram) &&
            bind(lib, "nvrtcDestroyProgram", n.DestroyProgram) &&
            bind(lib, "nvrtcCompileProgram", n.CompileProgram) &&
            bind(lib, "nvrtcGetCUBINSize", n.GetCUBINSize) && bind(lib, "nvrtcGetCUBIN", n.GetCUBIN) &&
            bind(lib, "nvrtcGetProgramLogSize", n.GetProgramLogSize) &&
            bind(lib, "nvrtcGetProgramLog", n.GetProgramLog) &&
            bind(lib, "nvrtcGetErrorString", n.GetErrorString);
  if (!ok) {
    g_rtc.status = KB200_ENONVRTC;
    return fail(KB200_ENONVRTC, "libnvrtc lacks a required symbol");
  }
  g_rtc.status = 0;
  return 0;
}

int cu_fail(CUresult r, const char* what) {
  const char* s = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
  return fail((int)r, "[KeOps-B200] CUDA driver error in %s: %s (CUresult %d)", what, s ? s : "?", (int)r);
}

#define CU_TRY(call, what)                         \
  do {                                             \
    CUresult r_ = (call);                          \
    if (r_ != CUDA_SUCCESS) return cu_fail(r_, what); \
  } while (0)

}  // namespace

struct kb200_module {
  CUmodule mod = nullptr;
  CUcontext ctx = nullptr;
  CUdevice dev = 0;
};

// Kernel handle -> context that owns it.  Driver calls that take a CUfunction (attribute queries, occupancy, launches) need
// that context current on the CALLING thread; PyTorch's autograd worker threads -- where every backward reduction is
// issued -- may have no current driver context at all (seen on a 2-GPU NCCL job: cuOccupancyMaxActiveBlocksPerMultiprocessor
// -> CUDA_ERROR_INVALID_CONTEXT).  CtxScope makes the owner current for the duration of one entry point when it is not.
static std::mutex g_fn_mutex;
static std::unordered_map<CUfunction, CUcontext> g_fn_ctx;

struct CtxScope {
  bool pushed = false;
  explicit CtxScope(kb200_function fn) {
    CUcontext want = nullptr;
    {
      std::lock_guard<std::mutex> lock(g_fn_mutex);
      auto it = g_fn_ctx.find(reinterpret_cast<CUfunction>(fn));
      if (it != g_fn_ctx.end()) want = it->second;
    }
    if (!want) return;
    CUcontext cur = nullptr;
    if (g_drv.CtxGetCurrent(&cur) == CUDA_SUCCESS && cur == want) return;
    pushed = g_drv.CtxPushCurrent(want) == CUDA_SUCCESS;
  }
  ~CtxScope() {
    if (pushed) {
      CUcontext dummy;
      g_drv.CtxPopCurrent(&dummy);
    }
  }
  CtxScope(const CtxScope&) = delete;
  CtxScope& operator=(const CtxScope&) = delete;
};


extern "C" {

int kb200_abi_version(void) { return KB200_ABI_VERSION; }

const char* kb200_last_error(void) { return g_err.c_str(); }

void kb200_free(void* p) { free(p); }

int kb200_compile_to_memory(const char* cu_source, const char* program_name, const char* const* options,
                            int num_options, void** cubin, size_t* cubin_size, char* log,
                            size_t log_capacity) {
  if (!cu_source || !cubin || !cubin_size) return fail(KB200_EINVAL, "kb200_compile: null argument");
  if (log && log_capacity) log[0] = 0;
  int st = nvrtc();
  if (st) return st;
  nvrtcProgram prog;
  nvrtcResult r = g_rtc.CreateProgram(&prog, cu_source, program_name ? program_name : "kb200_kernel.cu", 0,
                                      nullptr, nullptr);
  if (r != NVRTC_SUCCESS) return fail(KB200_NVRTC_BASE + (int)r, "nvrtcCreateProgram: %s", g_rtc.GetErrorString(r));
  std::vector<const char*> opts;
  opts.push_back("--gpu-architecture=sm_100a");
  opts.push_back("--std=c++17");
  opts.push_back("-default-device");
  for (int i = 0; i < num_options; ++i) opts.push_back(options[i]);
  r = g_rtc.CompileProgram(prog, (int)opts.size(), opts.data());
  size_t logsize = 0;
  g_rtc.GetProgramLogSize(prog, &logsize);
  std::string full_log(logsize, '\0');
  if (logsize > 1) g_rtc.GetProgramLog(prog, full_log.data());
  if (log && log_capacity) {
    snprintf(log, log_capacity, "%s", full_log.c_str());
  }
  if (r != NVRTC_SUCCESS) {
    g_rtc.DestroyProgram(&prog);
    return fail(KB200_NVRTC_BASE + (int)r, "NVRTC compilation failed (%s):\n%.800s", g_rtc.GetErrorString(r),
                full_log.c_str());
  }
  size_t n = 0;
  r = g_rtc.GetCUBINSize(prog, &n);
  if (r != NVRTC_SUCCESS || n == 0) {
    g_rtc.DestroyProgram(&prog);
    return fail(KB200_NVRTC_BASE + (int)r, "nvrtcGetCUBINSize failed");
  }
  char* buf = static_cast<char*>(malloc(n));
  if (!buf) {
    g_rtc.DestroyProgram(&prog);
    return fail(KB200_EINVAL, "out of host memory");
  }
  r = g_rtc.GetCUBIN(prog, buf);
  g_rtc.DestroyProgram(&prog);
  if (r != NVRTC_SUCCESS) {
    free(buf);
    return fail(KB200_NVRTC_BASE + (int)r, "nvrtcGetCUBIN failed");
  }
  *cubin = buf;
  *cubin_size = n;
  return 0;
}

int kb200_compile(const char* cu_source, const char* program_name, const char* const* options, int num_options,
                  const char* cubin_path, char* log, size_t log_capacity) {
  if (!cubin_path) return fail(KB200_EINVAL, "kb200_compile: null output path");
  void* cubin = nullptr;
  size_t n = 0;
  int st = kb200_compile_to_memory(cu_source, program_name, options, num_options, &cubin, &n, log, log_capacity);
  if (st) return st;
  std::string tmp = std::string(cubin_path) + ".tmp" + std::to_string((long)getpid());
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) {
    free(cubin);
    return fail(KB200_EIO, "cannot open %s for writing", tmp.c_str());
  }
  size_t w = fwrite(cubin, 1, n, f);
  fclose(f);
  free(cubin);
  if (w != n || rename(tmp.c_str(), cubin_path) != 0) {
    remove(tmp.c_str());
    return fail(KB200_EIO, "cannot write %s", cubin_path);
  }
  return 0;
}

int kb200_module_load(const void* cubin, size_t cubin_size, int device, kb200_module** module) {
  if (!cubin || !cubin_size || !module) return fail(KB200_EINVAL, "kb200_module_load: null argument");
  int st = driver();
  if (st) return st;
  kb200_module* m = new kb200_module();
  CUresult r = g_drv.DeviceGet(&m->dev, device);
  if (r == CUDA_SUCCESS) r = g_drv.PrimaryCtxRetain(&m->ctx, m->dev);
  if (r != CUDA_SUCCESS) {
    delete m;
    return cu_fail(r, "cuDevicePrimaryCtxRetain");
  }
  g_drv.CtxPushCurrent(m->ctx);
  r = g_drv.ModuleLoadData(&m->mod, cubin);
  CUcontext dummy;
  g_drv.CtxPopCurrent(&dummy);
  if (r != CUDA_SUCCESS) {
    g_drv.PrimaryCtxRelease(m->dev);
    delete m;
    return cu_fail(r, "cuModuleLoadData");
  }
  *module = m;
  return 0;
}

int kb200_module_load_file(const char* cubin_path, int device, kb200_module** module) {
  if (!cubin_path) return fail(KB200_EINVAL, "kb200_module_load_file: null path");
  FILE* f = fopen(cubin_path, "rb");
  if (!f) return fail(KB200_EIO, "cannot open %s", cubin_path);
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> buf(n > 0 ? n : 0);
  size_t rd = n > 0 ? fread(buf.data(), 1, n, f) : 0;
  fclose(f);
  if (n <= 0 || rd != (size_t)n) return fail(KB200_EIO, "cannot read %s", cubin_path);
  return kb200_module_load(buf.data(), buf.size(), device, module);
}

int kb200_module_unload(kb200_module* m) {
  if (!m) return 0;
  if (g_drv.status == 0) {
    g_drv.CtxPushCurrent(m->ctx);
    g_drv.ModuleUnload(m->mod);
    CUcontext dummy;
    g_drv.CtxPopCurrent(&dummy);
    g_drv.PrimaryCtxRelease(m->dev);
  }
  delete m;
  return 0;
}

int kb200_module_get_function(kb200_module* m, const char* name, kb200_function* fn) {
  if (!m || !name || !fn) return fail(KB200_EINVAL, "kb200_module_get_function: null argument");
  CUfunction f;
  CU_TRY(g_drv.ModuleGetFunction(&f, m->mod, name), "cuModuleGetFunction");
  {
    std::lock_guard<std::mutex> lock(g_fn_mutex);
    g_fn_ctx[f] = m->ctx;
  }
  *fn = reinterpret_cast<kb200_function>(f);
  return 0;
}

int kb200_function_info(kb200_function fn, int* num_regs, int* static_smem_bytes, int* max_threads_per_block) {
  int st = driver();
  if (st) return st;
  CUfunction f = reinterpret_cast<CUfunction>(fn);
  CtxScope scope(fn);
  if (num_regs) CU_TRY(g_drv.FuncGetAttribute(num_regs, CU_FUNC_ATTRIBUTE_NUM_REGS, f), "cuFuncGetAttribute");
  if (static_smem_bytes)
    CU_TRY(g_drv.FuncGetAttribute(static_smem_bytes, CU_FUNC_ATTRIBUTE_SHARED_SIZE_BYTES, f), "cuFuncGetAttribute");
  if (max_threads_per_block)
    CU_TRY(g_drv.FuncGetAttribute(max_threads_per_block, CU_FUNC_ATTRIBUTE_MAX_THREADS_PER_BLOCK, f),
           "cuFuncGetAttribute");
  return 0;
}

int kb200_max_active_blocks(kb200_function fn, int block_size, size_t dynamic_smem_bytes, int* blocks_per_sm) {
  int st = driver();
  if (st) return st;
  if (!blocks_per_sm) return fail(KB200_EINVAL, "null argument");
  CtxScope scope(fn);
  CU_TRY(g_drv.OccupancyMaxActiveBlocks(blocks_per_sm, reinterpret_cast<CUfunction>(fn), block_size,
                                        dynamic_smem_bytes),
         "cuOccupancyMaxActiveBlocksPerMultiprocessor");
  return 0;
}

int kb200_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes,
                      int64_t* total_mem_bytes) {
  int st = driver();
  if (st) return st;
  CUdevice dev;
  CU_TRY(g_drv.DeviceGet(&dev, device), "cuDeviceGet");
  int v = 0;
  if (sm_count) {
    CU_TRY(g_drv.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, dev), "cuDeviceGetAttribute");
    *sm_count = v;
  }
  if (cc_major) {
    CU_TRY(g_drv.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MAJOR, dev), "cuDeviceGetAttribute");
    *cc_major = v;
  }
  if (cc_minor) {
    CU_TRY(g_drv.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_COMPUTE_CAPABILITY_MINOR, dev), "cuDeviceGetAttribute");
    *cc_minor = v;
  }
  if (l2_bytes) {
    CU_TRY(g_drv.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_L2_CACHE_SIZE, dev), "cuDeviceGetAttribute");
    *l2_bytes = v;
  }
  if (total_mem_bytes) {
    size_t b = 0;
    CU_TRY(g_drv.DeviceTotalMem(&b, dev), "cuDeviceTotalMem");
    *total_mem_bytes = (int64_t)b;
  }
  return 0;
}

int kb200_launch(kb200_function fn, void* stream, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by,
                 unsigned bz, unsigned dynamic_smem_bytes, void** params) {
  int st = driver();
  if (st) return st;
  if (!fn) return fail(KB200_EINVAL, "kb200_launch: null function");
  CtxScope scope(fn);
  CU_TRY(g_drv.LaunchKernel(reinterpret_cast<CUfunction>(fn), gx, gy, gz, bx, by, bz, dynamic_smem_bytes,
                            reinterpret_cast<CUstream>(stream), params, nullptr),
         "cuLaunchKernel");
  return 0;
}

int kb200_genred_launch(kb200_function reduce_fn, kb200_function combine_fn, void* stream, int64_t nrow,
                        int64_t ncol, int rows_per_block, int block, int nsplit, int64_t col_chunk, void* out,
                        void* workspace, int nargs, const void* const* args, int nbatch,
                        const int64_t* batch_strides) {
  return kb200_genred_launch_smem(reduce_fn, combine_fn, stream, nrow, ncol, rows_per_block, block, nsplit, col_chunk, out,
                                  workspace, nargs, args, nbatch, batch_strides, 0u);
}

int kb200_genred_launch_smem(kb200_function reduce_fn, kb200_function combine_fn, void* stream, int64_t nrow,
                             int64_t ncol, int rows_per_block, int block, int nsplit, int64_t col_chunk, void* out,
                             void* workspace, int nargs, const void* const* args, int nbatch,
                             const int64_t* batch_strides, unsigned dynamic_smem_bytes) {
  int st = driver();
  if (st) return st;
  if (!reduce_fn || !out || rows_per_block <= 0 || block <= 0 || nsplit <= 0 || nargs < 0 || nargs > 60 ||
      nbatch < 1 || nbatch > 65535)
    return fail(KB200_EINVAL, "kb200_genred_launch: invalid argument");
  if (nsplit > 1 && (!workspace || !combine_fn))
    return fail(KB200_EINVAL, "kb200_genred_launch: split reduction needs a workspace and a combine kernel");
  if (nrow <= 0) return 0;
  CtxScope scope(reduce_fn);
  long long nrow_ = nrow, ncol_ = ncol, chunk_ = col_chunk;
  void* out_ = out;
  void* ws_ = workspace;
  const void* argv[64];
  long long strides[64];
  void* params[5 + 128];
  params[0] = &nrow_;
  params[1] = &ncol_;
  params[2] = &chunk_;
  params[3] = &out_;
  params[4] = &ws_;
  for (int i = 0; i < nargs; ++i) {
    argv[i] = args[i];
    params[5 + i] = &argv[i];
    strides[i] = batch_strides ? (long long)batch_strides[i] : 0LL;
    params[5 + nargs + i] = &strides[i];
  }
  unsigned gx = (unsigned)((nrow + rows_per_block - 1) / rows_per_block);
  CU_TRY(g_drv.LaunchKernel(reinterpret_cast<CUfunction>(reduce_fn), gx, (unsigned)nsplit, (unsigned)nbatch, (unsigned)block,
                            1, 1, dynamic_smem_bytes, reinterpret_cast<CUstream>(stream), params, nullptr),
         "cuLaunchKernel(kb200_reduce)");
  if (nsplit > 1) {
    int ns = nsplit;
    const void* ws_c = workspace;
    void* cparams[4] = {&nrow_, &ns, &ws_c, &out_};
    unsigned cg = (unsigned)((nrow + 255) / 256);
    CU_TRY(g_drv.LaunchKernel(reinterpret_cast<CUfunction>(combine_fn), cg, 1, (unsigned)nbatch, 256, 1, 1, 0,
                              reinterpret_cast<CUstream>(stream), cparams, nullptr),
           "cuLaunchKernel(kb200_combine)");
  }
  return 0;
}

int kb200_function_set_max_dynamic_smem(kb200_function fn, int bytes) {
  int st = driver();
  if (st) return st;
  if (!fn || bytes < 0) return fail(KB200_EINVAL, "kb200_function_set_max_dynamic_smem: invalid argument");
  CtxScope scope(fn);
  CU_TRY(g_drv.FuncSetAttribute(reinterpret_cast<CUfunction>(fn), CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, bytes),
         "cuFuncSetAttribute(MAX_DYNAMIC_SHARED_SIZE_BYTES)");
  return 0;
}

int kb200_tensor_map_2d(void* tensor_map, int elem_bytes, const void* base, uint64_t inner, uint64_t outer,
                        uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  int st = driver();
  if (st) return st;
  if (!tensor_map || !base || (reinterpret_cast<uintptr_t>(tensor_map) & 63) || (reinterpret_cast<uintptr_t>(base) & 15) ||
      (row_stride_bytes & 15) || box_inner == 0 || box_outer == 0 || box_inner > 256 || box_outer > 256)
    return fail(KB200_EINVAL, "kb200_tensor_map_2d: invalid argument / alignment");
  CUtensorMapDataType dt;
  switch (elem_bytes) {
    case 2: dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; break;
    case 4: dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; break;
    default: return fail(KB200_EINVAL, "kb200_tensor_map_2d: elem_bytes must be 2 (bf16) or 4 (f32)");
  }
  CUtensorMapSwizzle sw;
  switch (swizzle_bytes) {
    case 0: sw = CU_TENSOR_MAP_SWIZZLE_NONE; break;
    case 32: sw = CU_TENSOR_MAP_SWIZZLE_32B; break;
    case 64: sw = CU_TENSOR_MAP_SWIZZLE_64B; break;
    case 128: sw = CU_TENSOR_MAP_SWIZZLE_128B; break;
    default: return fail(KB200_EINVAL, "kb200_tensor_map_2d: swizzle_bytes must be 0, 32, 64 or 128");
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CU_TRY(g_drv.TensorMapEncodeTiled(reinterpret_cast<CUtensorMap*>(tensor_map), dt, 2, const_cast<void*>(base), dims,
                                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE),
         "cuTensorMapEncodeTiled");
  return 0;
}

int kb200_genred_launch_ranges(kb200_function reduce_fn, void* stream, int64_t nrow, int64_t ncol, int block,
                               int64_t grid_x, void* out, int nargs, const void* const* args,
                               const int64_t* ranges_x, const int64_t* slices_x, const int64_t* ranges_y,
                               const int64_t* block_ends, int nranges) {
  int st = driver();
  if (st) return st;
  if (!reduce_fn || !out || block <= 0 || nargs < 0 || nargs > 60 || nranges < 0 || grid_x < 0 ||
      grid_x > 2147483647LL || (nranges > 0 && (!ranges_x || !slices_x || !ranges_y || !block_ends)))
    return fail(KB200_EINVAL, "kb200_genred_launch_ranges: invalid argument");
  if (nrow <= 0 || nranges == 0 || grid_x == 0) return 0;
  CtxScope scope(reduce_fn);
  long long nrow_ = nrow, ncol_ = ncol;
  void* out_ = out;
  const void* argv[64];
  const void *rx = ranges_x, *sx = slices_x, *ry = ranges_y, *be = block_ends;
  int nr = nranges;
  void* params[3 + 64 + 5];
  params[0] = &nrow_;
  params[1] = &ncol_;
  params[2] = &out_;
  for (int i = 0; i < nargs; ++i) {
    argv[i] = args[i];
    params[3 + i] = &argv[i];
  }
  params[3 + nargs] = &rx;
  params[4 + nargs] = &sx;
  params[5 + nargs] = &ry;
  params[6 + nargs] = &be;
  params[7 + nargs] = &nr;
  CU_TRY(g_drv.LaunchKernel(reinterpret_cast<CUfunction>(reduce_fn), (unsigned)grid_x, 1, 1, (unsigned)block, 1, 1, 0,
                            reinterpret_cast<CUstream>(stream), params, nullptr),
         "cuLaunchKernel(kb200_reduce_ranges)");
  return 0;
}

}  // extern "C"
