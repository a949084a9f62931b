/* kb200.h -- C ABI of the B200-native KeOps map-reduce runtime (libkb200.so).
 *
 * This is the drop-in boundary for the reference's native layer.  Every entry point is plain C
 * (pointers and sizes only, no torch / pybind types) and replaces one piece of the reference's
 * C++ binder; the "replaces" notes cite /root/reference paths (getkeops/keops v2.3).
 *
 * Conventions
 *   - every function returns 0 on success, otherwise a non-zero status: a CUresult for driver
 *     failures, an nvrtcResult + 10000 for compiler failures, or a KB200_E* code below.  A
 *     human-readable message for the last failure on the calling thread is available through
 *     kb200_last_error() (the reference prints to stderr and throws "[KeOps] Cuda error.",
 *     keopscore/keopscore/include/utils_pe.h:27-37, and drops the NVRTC log,
 *     keopscore/keopscore/binders/nvrtc/nvrtc_jit.cpp:110-112).
 *   - all sizes are int64_t (the reference uses `signed long int` throughout).
 *   - the caller owns every buffer, including `out` and the split-reduction workspace
 *     (reference: pykeops/pykeops/common/keops_io/LoadKeOps.py:227-229 allocates `out` in Python).
 *   - launches are asynchronous on the caller's CUstream; nothing here synchronises the host
 *     (reference: NULL stream + cuCtxSynchronize, keops_nvrtc.cpp:609-616).
 *   - the library loads libcuda / libnvrtc lazily, so it can be dlopen'ed (and its symbols
 *     inspected) on a machine without a GPU; compute entry points then fail with KB200_ENOCUDA.
 */
#ifndef KB200_H
#define KB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KB200_ABI_VERSION 4

#define KB200_OK 0
#define KB200_ENOCUDA 20001  /* libcuda.so.1 could not be loaded / cuInit failed        */
#define KB200_ENONVRTC 20002 /* libnvrtc could not be loaded                            */
#define KB200_EINVAL 20003   /* bad argument                                            */
#define KB200_EIO 20004      /* file read / write failure                               */
#define KB200_NVRTC_BASE 10000

typedef struct kb200_module kb200_module; /* a loaded cubin (CUmodule) on one device   */
typedef void* kb200_function;             /* a CUfunction inside a module               */

int kb200_abi_version(void);

/* Message describing the last failure on this thread ("" if none). */
const char* kb200_last_error(void);

/* ---- compilation ------------------------------------------------------------------------
 * replaces: extern "C" int Compile(const char* target_file, const char* cu_code, int use_half,
 *           int use_fast_math, int device_id, const char* cuda_include_path)
 *           (keopscore/keopscore/binders/nvrtc/nvrtc_jit.cpp:38-136).
 * Differences: the target is always sm_100a (no device query, so it also works without a GPU),
 * the output is a raw cubin file (no size_t prefix), and the compiler log is returned.
 * `options` are extra NVRTC flags (e.g. "--use_fast_math", "-lineinfo"); `log` may be NULL.      */
int kb200_compile(const char* cu_source, const char* program_name, const char* const* options,
                  int num_options, const char* cubin_path, char* log, size_t log_capacity);

/* Same, but returns the cubin in a malloc'ed buffer that the caller releases with kb200_free. */
int kb200_compile_to_memory(const char* cu_source, const char* program_name,
                            const char* const* options, int num_options, void** cubin,
                            size_t* cubin_size, char* log, size_t log_capacity);
void kb200_free(void* p);

/* ---- modules ----------------------------------------------------------------------------
 * replaces: KeOps_module<TYPE>::KeOps_module(device_id, nargs, target_file_name) and its
 *           destructor (keopscore/keopscore/binders/nvrtc/keops_nvrtc.cpp:328-388): retains the
 *           device's primary context and loads the cubin.  No per-module device pointer table is
 *           kept: kernel arguments are passed by value.                                          */
int kb200_module_load(const void* cubin, size_t cubin_size, int device, kb200_module** module);
int kb200_module_load_file(const char* cubin_path, int device, kb200_module** module);
int kb200_module_unload(kb200_module* module);
int kb200_module_get_function(kb200_module* module, const char* name, kb200_function* fn);

/* Function / device introspection used by the launch-shape tuner (the reference hard-codes
 * block size 192 and a 48 KB shared memory assumption, keopscore/include/CudaSizes.h:11-31). */
int kb200_function_info(kb200_function fn, int* num_regs, int* static_smem_bytes,
                        int* max_threads_per_block);
int kb200_max_active_blocks(kb200_function fn, int block_size, size_t dynamic_smem_bytes,
                            int* blocks_per_sm);
int kb200_device_info(int device, int* sm_count, int* cc_major, int* cc_minor,
                      int64_t* l2_bytes, int64_t* total_mem_bytes);

/* Opt a kernel in to more than 48 KB of dynamic shared memory (cuFuncSetAttribute); the tensor-core
 * K-nearest-neighbour kernel stages its operand tiles in ~140 KB.  The reference assumes 48 KB
 * (keopscore/keopscore/include/CudaSizes.h:11-31).                                                     */
int kb200_function_set_max_dynamic_smem(kb200_function fn, int bytes);

/* Encode a TMA descriptor (CUtensorMap, 128 bytes, 64-byte aligned host buffer) for a row-major 2-D
 * array `base[outer][inner]` of 2-byte (bf16) or 4-byte (f32) elements with a row pitch of
 * `row_stride_bytes`; the kernel receives it as a __grid_constant__ parameter and issues
 * cp.async.bulk.tensor loads of (box_outer x box_inner) tiles, optionally 32/64/128-byte swizzled
 * (the layout tcgen05.mma shared-memory descriptors expect).  No reference counterpart: the
 * reference's kernels only use plain global loads.                                     */
int kb200_tensor_map_2d(void* tensor_map, int elem_bytes, const void* base, uint64_t inner, uint64_t outer,
                        uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes);

/* ---- launch -----------------------------------------------------------------------------
 * Generic asynchronous launch (cuLaunchKernel) on `stream` (a CUstream / cudaStream_t; NULL =
 * legacy default stream).  `params[i]` points to the i-th kernel argument.                      */
int kb200_launch(kb200_function fn, void* stream, unsigned grid_x, unsigned grid_y,
                 unsigned grid_z, unsigned block_x, unsigned block_y, unsigned block_z,
                 unsigned dynamic_smem_bytes, void** params);

/* The map-reduce launch: out[r, :] = Reduce_{c in [0, ncol)} F(row r, column c).
 * replaces: KeOps_module<TYPE>::launch_kernel(...) for the 1D and 2D schemes
 *           (keopscore/keopscore/binders/nvrtc/keops_nvrtc.cpp:390-643, launches at :596-614 and
 *           :509-566) and the pybind wrapper KeOps_module_python<T>::operator()
 *           (pykeops/pykeops/common/keops_io/pykeops_nvrtc.cpp:17-83).
 *
 *   reduce_fn / combine_fn  the two kernels of a generated module ("kb200_reduce",
 *                           "kb200_combine"); combine_fn may be NULL when nsplit == 1
 *   nrow, ncol              output rows / streamed (reduced) rows: (nx, ny) when tagI == 0 and
 *                           (ny, nx) when tagI == 1 -- the caller has already resolved tagI, unlike
 *                           the reference which swaps twice (LoadKeOps.py:151-153,
 *                           keops_nvrtc.cpp:431-439)
 *   rows_per_block, block   rows handled by one CTA (threads * rows-per-thread) and CTA size
 *   nsplit, col_chunk       number of column splits (gridDim.y) and columns per split; the split
 *                           scheme needs `workspace` of nsplit*nrow*dimpart elements when nsplit > 1
 *   out                     device pointer, nrow * dim elements
 *   args / nargs            device pointers of the formula's variables, in Var index order
 *   nbatch, batch_strides   number of independent problems (gridDim.z, 1..65535) and, per argument,
 *                           the element stride between two problems (0 = argument shared by all;
 *                           NULL = all zero).  `out` / `workspace` are batch-major.  This replaces
 *                           the reference's emulation of batch dimensions through block-diagonal
 *                           ranges (keopscore/keopscore/include/Ranges.h:40-86).
 * All pointers are device pointers of the formula dtype (float or double).                       */
int kb200_genred_launch(kb200_function reduce_fn, kb200_function combine_fn, void* stream,
                        int64_t nrow, int64_t ncol, int rows_per_block, int block, int nsplit,
                        int64_t col_chunk, void* out, void* workspace, int nargs,
                        const void* const* args, int nbatch, const int64_t* batch_strides);

/* Same launch with `dynamic_smem_bytes` of dynamic shared memory for kb200_reduce: kernels whose staged tile exceeds the
 * 48 KB static limit (streamed rows of several hundred reals) declare it `extern __shared__` and are opted in with
 * kb200_function_set_max_dynamic_smem.  The reference sizes its tile as blockDim * dimY * sizeof(T) of dynamic shared
 * memory under a 48 KB assumption and shrinks the block size to fit (keopscore/keopscore/binders/nvrtc/keops_nvrtc.cpp:441-458,
 * keopscore/keopscore/include/CudaSizes.h:11-31).                                                                     */
int kb200_genred_launch_smem(kb200_function reduce_fn, kb200_function combine_fn, void* stream,
                             int64_t nrow, int64_t ncol, int rows_per_block, int block, int nsplit,
                             int64_t col_chunk, void* out, void* workspace, int nargs,
                             const void* const* args, int nbatch, const int64_t* batch_strides,
                             unsigned dynamic_smem_bytes);

/* Block-sparse map-reduce launch: row-range k = rows [ranges_x[2k], ranges_x[2k+1]) is reduced over the
 * union of the column ranges ranges_y[2l .. 2l+1], l in [slices_x[k-1], slices_x[k]) (slices_x[-1] = 0).
 * replaces: the `ranges` branch of KeOps_module<TYPE>::launch_kernel
 *           (keopscore/keopscore/binders/nvrtc/keops_nvrtc.cpp:571-595), its host-side lookup-table
 *           construction range_preprocess_from_device (:152-237) and kernel GpuConv1DOnDevice_ranges
 *           (keopscore/keopscore/mapreduce/gpu/GpuReduc1D_ranges.py:77-172).
 * Differences: no host copy of the ranges and no per-block lookup table -- every CTA locates its
 * (range, chunk) by bisection in `block_ends` (device, int64[nranges]: inclusive prefix sum of
 * ceil(range length / rows_per_block)); `grid_x` >= block_ends[nranges-1] (surplus CTAs exit).
 *
 *   reduce_fn   the "kb200_reduce_ranges" kernel of a module generated in ranges mode
 *   ranges_x, slices_x, ranges_y, block_ends   device pointers, int64, contiguous
 * Rows covered by no range are not written (as in the reference).                                     */
int kb200_genred_launch_ranges(kb200_function reduce_fn, void* stream, int64_t nrow, int64_t ncol,
                               int block, int64_t grid_x, void* out, int nargs, const void* const* args,
                               const int64_t* ranges_x, const int64_t* slices_x, const int64_t* ranges_y,
                               const int64_t* block_ends, int nranges);

#ifdef __cplusplus
}
#endif
#endif /* KB200_H */
