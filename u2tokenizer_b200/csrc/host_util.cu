#include "host_util.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "u2b200.h"

namespace u2 {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

const char* last_error() { return g_err; }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (dev < 0 || dev >= 64) return -1;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_bf16_4d(CUtensorMap* out, const void* base, int64_t k, int64_t rows, int64_t zi,
                      int64_t zo, int64_t ld, int64_t stride_zi, int64_t stride_zo, int box_k,
                      int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(U2_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {(cuuint64_t)k, (cuuint64_t)rows, (cuuint64_t)(zi > 0 ? zi : 1),
                        (cuuint64_t)(zo > 0 ? zo : 1)};
  // strides for dims 1..3 in bytes; size-1 dims still need a legal (multiple of 16) stride
  int64_t s1 = ld * 2;
  int64_t s2 = (dims[2] > 1 ? stride_zi * 2 : s1 * (int64_t)rows);
  int64_t s3 = (dims[3] > 1 ? stride_zo * 2 : s2 * (int64_t)dims[2]);
  if (dims[2] == 1) s2 = (s2 + 15) / 16 * 16;
  if (dims[3] == 1) s3 = (s3 + 15) / 16 * 16;
  if (s2 <= 0) s2 = 16;
  if (s3 <= 0) s3 = 16;
  cuuint64_t strides[3] = {(cuuint64_t)s1, (cuuint64_t)s2, (cuuint64_t)s3};
  cuuint32_t box[4] = {(cuuint32_t)box_k, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(U2_ERR_CUDA,
                     "cuTensorMapEncodeTiled(bf16 4d) failed: %d (k=%lld rows=%lld zi=%lld zo=%lld "
                     "ld=%lld szi=%lld szo=%lld box=%dx%d base=%p)",
                     (int)r, (long long)k, (long long)rows, (long long)zi, (long long)zo,
                     (long long)ld, (long long)stride_zi, (long long)stride_zo, box_k, box_rows, base);
  return U2_OK;
}

int make_tmap_store_4d(CUtensorMap* out, void* base, int elem_bytes, int64_t n, int64_t rows, int64_t zi, int64_t zo,
                       int64_t ld, int64_t stride_zi, int64_t stride_zo, int box_n, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(U2_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (elem_bytes != 2 && elem_bytes != 4) return set_error(U2_ERR_ARG, "store tensor map: bf16 or fp32 elements");
  cuuint64_t dims[4] = {(cuuint64_t)n, (cuuint64_t)rows, (cuuint64_t)(zi > 0 ? zi : 1), (cuuint64_t)(zo > 0 ? zo : 1)};
  int64_t s1 = ld * elem_bytes;
  int64_t s2 = (dims[2] > 1 ? stride_zi * elem_bytes : s1 * (int64_t)rows);
  int64_t s3 = (dims[3] > 1 ? stride_zo * elem_bytes : s2 * (int64_t)dims[2]);
  if (dims[2] == 1) s2 = (s2 + 15) / 16 * 16;
  if (dims[3] == 1) s3 = (s3 + 15) / 16 * 16;
  if (s2 <= 0) s2 = 16;
  if (s3 <= 0) s3 = 16;
  cuuint64_t strides[3] = {(cuuint64_t)s1, (cuuint64_t)s2, (cuuint64_t)s3};
  cuuint32_t box[4] = {(cuuint32_t)box_n, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  // the inner box extent is 64 B (32 bf16) or 128 B (32 fp32): the swizzle span equals the box row
  CUresult r = enc(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   elem_bytes == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(U2_ERR_CUDA,
                     "cuTensorMapEncodeTiled(store 4d) failed: %d (n=%lld rows=%lld zi=%lld zo=%lld ld=%lld szi=%lld szo=%lld "
                     "box=%dx%d base=%p)",
                     (int)r, (long long)n, (long long)rows, (long long)zi, (long long)zo, (long long)ld, (long long)stride_zi,
                     (long long)stride_zo, box_n, box_rows, base);
  return U2_OK;
}

int make_tmap_f32_3d(CUtensorMap* out, const void* base, int64_t d0, int64_t d1, int64_t d2,
                     int64_t stride1_elems, int64_t stride2_elems, int box0, int box1, int box2) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(U2_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)(stride1_elems * 4), (cuuint64_t)(stride2_elems * 4)};
  cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, (cuuint32_t)box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(U2_ERR_CUDA, "cuTensorMapEncodeTiled(f32 3d) failed: %d", (int)r);
  return U2_OK;
}

int make_tmap_f32_nd(CUtensorMap* out, const void* base, int n, const int64_t* dims, const int64_t* strides_bytes,
                     const int* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(U2_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (n < 1 || n > 5) return set_error(U2_ERR_ARG, "tensor map rank %d", n);
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < n; ++i) { d[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < n; ++i) st[i] = (cuuint64_t)strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)n, const_cast<void*>(base), d, st, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(U2_ERR_CUDA, "cuTensorMapEncodeTiled(f32 %d-d) failed: %d", n, (int)r);
  return U2_OK;
}

}  // namespace u2

namespace u2 { const char* last_error(); }

extern "C" U2_API int u2_version(void) { return 1; }
extern "C" U2_API const char* u2_last_error(void) { return u2::last_error(); }
extern "C" U2_API int u2_device_sm_count(void) { return u2::num_sms(); }
