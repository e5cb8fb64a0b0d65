// Host-side helpers shared by the C-ABI translation units: error reporting, device queries and
// CUtensorMap construction (driver entry point resolved at run time through the CUDA runtime, so
// the library has no link-time dependency on libcuda and loads on a machine without a GPU).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace u2 {

int set_error(int code, const char* fmt, ...);
int num_sms();

// 4-D bf16 tensor map: dims (inner -> outer) {k, rows, zi, zo}; strides in ELEMENTS for rows/zi/zo
// (a stride of 0 is allowed for size-1 dims); box = {box_k, box_rows, 1, 1}; 128-byte swizzle,
// out-of-bounds elements read as zero.
int make_tmap_bf16_4d(CUtensorMap* out, const void* base, int64_t k, int64_t rows, int64_t zi,
                      int64_t zo, int64_t ld, int64_t stride_zi, int64_t stride_zo, int box_k,
                      int box_rows);

// 4-D tensor map of a GEMM OUTPUT (bf16: 64-byte swizzle, fp32: 128-byte swizzle; the box row is 32 elements): dims
// (inner -> outer) {n, rows, zi, zo}, strides in ELEMENTS; box elements outside the extents are not written by a store.
int make_tmap_store_4d(CUtensorMap* out, void* base, int elem_bytes, int64_t n, int64_t rows, int64_t zi, int64_t zo,
                       int64_t ld, int64_t stride_zi, int64_t stride_zo, int box_n, int box_rows);

// 3-D fp32 tensor map (no swizzle) used by the patch-embed brick gather.
int make_tmap_f32_3d(CUtensorMap* out, const void* base, int64_t d0, int64_t d1, int64_t d2,
                     int64_t stride1_elems, int64_t stride2_elems, int box0, int box1, int box2);

// n-D fp32 tensor map (no swizzle): dims / box inner -> outer, strides in BYTES for dims 1..n-1.
int make_tmap_f32_nd(CUtensorMap* out, const void* base, int n, const int64_t* dims, const int64_t* strides_bytes,
                     const int* box);

}  // namespace u2

#define U2_CHECK_LAUNCH(what)                                                          \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess)                                                            \
      return ::u2::set_error(U2_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e__));   \
  } while (0)
