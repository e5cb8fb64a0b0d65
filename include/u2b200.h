/*
 * u2b200.h - C ABI of libu2b200.so: the sm_100a kernels behind the mu2-LLM
 * "visual-tokenize-then-decode" hot path (CT volume -> 3D patch embed -> ViT3D -> spatial-pooling
 * projector -> mu2-Tokenizer -> splice -> Qwen3/Llama decoder forward / greedy decode).
 *
 * The reference (Siyou-Li/u2Tokenizer) is pure Python: it has no FFI of its own. The boundary it
 * exposes is the HuggingFace module surface (forward()/generate(), reference
 * src/model/language_model/u2llama.py:41-127); that surface is mirrored in Python by
 * u2tokenizer_b200/modeling.py and everything underneath it calls the entry points declared
 * here. Each entry point cites the reference call site whose arithmetic it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch allocates), no allocation inside;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, re-entrant;
 *   - return value: 0 (U2_OK) or a negative U2_ERR_* code; u2_last_error() gives the message;
 *   - bf16 = __nv_bfloat16 bits, "f32" = float; row-major unless stated otherwise.
 */
#ifndef U2B200_H_
#define U2B200_H_

#include <stdint.h>

#if defined(__GNUC__)
#define U2_API __attribute__((visibility("default")))
#else
#define U2_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define U2_OK 0
#define U2_ERR_ARG (-1)
#define U2_ERR_CUDA (-2)
#define U2_ERR_UNSUPPORTED (-3)

#define U2_DT_BF16 0
#define U2_DT_F32 1

#define U2_ACT_NONE 0
#define U2_ACT_GELU 1 /* exact erf GELU (torch.nn.GELU default) */
#define U2_ACT_SILU 2

/* Library / device info ----------------------------------------------------------------------- */
U2_API int u2_version(void);                 /* ABI version, currently 1 */
U2_API const char* u2_last_error(void);      /* message of the last failing call on this thread */
U2_API int u2_device_sm_count(void);         /* SMs of the current device (148 on B200), <0 on error */

/* GEMM --------------------------------------------------------------------------------------------
 * For every batch z = (zo, zi):
 *     C[z] (M x N) = act( alpha * A[z] (M x K) * B[z'] (N x K)^T + bias[n] ) + residual
 * A, B are bf16, K-major (row stride lda/ldb elements, multiples of 8); fp32 accumulation in TMEM.
 * Batch offsets (elements): A: zi*a_stride_zi + zo*a_stride_zo; B: (zi / b_zi_div)*b_stride_zi +
 * zo*b_stride_zo (b_zi_div > 1 shares one B among consecutive inner batches: GQA);
 * C: zi*c_stride_zi + zo*c_stride_zo.
 * Output row remap (row_div > 0): out_row = (r / row_div) * row_stride + row_off + r % row_div.
 * Residual (bf16, row stride ldr): row = r % res_row_mod when res_row_mod > 0 (broadcast table,
 * e.g. the ViT position embedding) else the output row (same batch offsets as C).
 * Replaces: every nn.Linear / torch.matmul on the path (reference src/model/u2tokenizer/rma.py:52-73,
 * tta.py:42-69, svr.py:107, spatial_pooling_projector.py:48-50; MONAI blocks; HF decoder Linears).
 */
typedef struct u2_gemm_desc {
  int32_t M, N, K;
  int32_t zi, zo;     /* inner / outer batch counts (<=0 means 1) */
  int32_t b_zi_div;   /* <=0 means 1 */
  int64_t lda, a_stride_zi, a_stride_zo;
  int64_t ldb, b_stride_zi, b_stride_zo;
  int64_t ldc, c_stride_zi, c_stride_zo;
  int32_t c_dtype;    /* U2_DT_BF16 or U2_DT_F32 */
  float alpha;
  const float* bias;  /* [N] fp32 or NULL */
  int32_t act;        /* U2_ACT_* */
  const void* residual; /* bf16 or NULL */
  int64_t ldr;
  int32_t res_row_mod;
  int32_t row_div, row_stride, row_off;
  int32_t block_n;    /* 0 = auto, else 64/128/256 */
} u2_gemm_desc;

U2_API int u2_gemm_bf16(const void* A, const void* B, void* C, const u2_gemm_desc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* U2B200_H_ */
