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

#define U2_EPI_NONE 0
#define U2_EPI_EXP_ROW 1
#define U2_EPI_DS_ROW 2

/* Library / device info ----------------------------------------------------------------------- */
U2_API int u2_version(void);                 /* ABI version, currently 1 */
U2_API const char* u2_last_error(void);      /* message of the last failing call on this thread */
U2_API int u2_device_sm_count(void);         /* SMs of the current device (148 on B200), <0 on error */

/* GEMM --------------------------------------------------------------------------------------------
 * For every batch z = (zo, zi):
 *     C[z] (M x N) = act( alpha * A[z] (M x K) * B[z'] (N x K)^T + bias[n] ) + residual
 * A, B are bf16, K-major (row stride lda/ldb elements, multiples of 8) unless a_mn / b_mn say otherwise;
 * fp32 accumulation in TMEM. residual == C (same ld) accumulates into a bf16 C (gradient accumulation).
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
  /* transposed operands (training: dgrad = dY * W, wgrad = dY^T * X, P^T dO ...): a_mn != 0 -> A is stored
   * [K][M] (element (m, k) at A[k * lda + m]); b_mn != 0 -> B is stored [K][N]. lda / ldb are then the strides
   * between consecutive contraction indices. No transposed copy is made: the tile is loaded MN-major. */
  int32_t a_mn, b_mn;
  /* fused epilogue of the attention backward (applied after alpha, before bias / act / residual; rowvec is indexed
   * rowvec[zo * rv_stride_zo + zi * rv_stride_zi + row]):
   *   U2_EPI_EXP_ROW:  v = exp(v - rowvec[row])             P = exp(scale * q.k - lse) straight out of the score GEMM
   *   U2_EPI_DS_ROW:   v = mul[row, col] * (v - rowvec[row]) dS = P * (dO.V^T - rowsum(dO * O)); mul (bf16) has C's
   *                                                          layout (ldc, c_stride_*) and may BE C (in place) */
  int32_t epi_op;
  const float* rowvec;
  int64_t rv_stride_zi, rv_stride_zo;
  const void* mul;
} u2_gemm_desc;

U2_API int u2_gemm_bf16(const void* A, const void* B, void* C, const u2_gemm_desc* desc, void* stream);

/* Row-wise normalisation ----------------------------------------------------------------------------
 * y = LayerNorm(x [+ residual]) * gamma + beta   (fp32 statistics, eps inside the sqrt)
 * y = RMSNorm (x [+ residual]) * gamma
 * x, residual, y, sum_out: bf16 rows of E elements (row strides ldx/ldr/ldy, multiples of 8);
 * gamma/beta fp32 [E]. When residual and sum_out are given, sum_out receives x + residual (the new
 * residual stream, row stride ldy). Replaces nn.LayerNorm (MONAI TransformerBlock norm1/norm2, ViT
 * final norm vit.py:123; tta.py:95,99,103) and HF Qwen3RMSNorm / LlamaRMSNorm.
 */
U2_API int u2_layernorm_bf16(const void* x, const void* residual, const float* gamma, const float* beta,
                             void* y, void* sum_out, int64_t rows, int32_t E, int64_t ldx, int64_t ldr,
                             int64_t ldy, float eps, void* stream);
U2_API int u2_rmsnorm_bf16(const void* x, const void* residual, const float* gamma, void* y, void* sum_out,
                           int64_t rows, int32_t E, int64_t ldx, int64_t ldr, int64_t ldy, float eps,
                           void* stream);

/* Softmax over fp32 score rows -> bf16 probabilities ---------------------------------------------------
 * Rows are indexed (i0 batch, i1 head in [0,H), i2 query in [0,S)); element strides given for input
 * and output. p[j] = softmax_j(in[j] * scale + rel_bias[(j - i2 + rel_max - 1) * H + i1]) over the
 * visible keys j < n (and j <= i2 + causal_off when causal). Columns [n, zero_pad_to) are written 0.
 * Replaces F.softmax in rma.py:60-73 (relative bias gather included), tta.py:55-57, the MONAI
 * SABlock softmax and the HF eager attention softmax + causal mask.
 */
typedef struct u2_softmax_desc {
  int64_t in_s0, in_s1, in_s2;
  int64_t out_s0, out_s1, out_s2;
  int32_t n0, H, S, n;
  float scale;
  const float* rel_bias;
  int32_t rel_max;
  int32_t causal, causal_off;
  int32_t zero_pad_to;
} u2_softmax_desc;
U2_API int u2_softmax_f32_bf16(const float* in, void* out, const u2_softmax_desc* desc, void* stream);

/* out[r, i] = silu(g) * u with (g, u) = gate_up[r, i], gate_up[r, I + i]  (interleaved == 0) or
 * gate_up[r, 2i], gate_up[r, 2i + 1] (interleaved != 0). Replaces `act_fn(gate_proj(x)) * up_proj(x)` of the HF decoder MLP
 * (transformers models/qwen3/modeling_qwen3.py:81-83, reached from the reference through super().forward,
 * src/model/language_model/u2llama.py:76-87). */
U2_API int u2_silu_mul_bf16(const void* gate_up, void* out, int64_t rows, int32_t I, int64_t ldg,
                            int64_t ldo, int32_t interleaved, void* stream);

/* Vision-front data movement ----------------------------------------------------------------------------
 * patchify: fp32 volume [frames, d0, d1, d2] -> bf16 patch rows [frames * n_patches, p0*p1*p2] in the
 * MONAI "perceptron" order "b c (h p1) (w p2) (d p3) -> b (h w d) (p1 p2 p3 c)", c == 1
 * (reference vit.py:90-99 -> MONAI PatchEmbeddingBlock).
 */
U2_API int u2_patchify_f32_bf16(const float* vol, void* rows, int64_t frames, int32_t d0, int32_t d1,
                                int32_t d2, int32_t p0, int32_t p1, int32_t p2, void* stream);
/* Fused 3-D patch embedding: out[f, 1 + t, :] = bf16(patch(f, t) . W^T + bias + pos[t]) for the fp32 volume
 * vol [frames, d0, d1, d2] (single channel), W [N, p0*p1*p2] bf16 in MONAI's (p1 p2 p3 c) feature order, bias fp32 [N],
 * pos bf16 [tokens, N], out bf16 [frames, out_frame_rows, N] (row 0 = cls and the rows behind the tokens are left to
 * u2_vit_frame_rows_bf16). 5-D TMA slabs of the volume are converted to the swizzled bf16 A operand in shared memory: the
 * einops gather of MONAI PatchEmbeddingBlock + Linear + position add (reference vit.py:90-99,115) in one kernel, the volume
 * is read once. Covers patch (p0, 4k, 16) on a (g0, 8m, 16) token grid (the canonical 4 x 16 x 16 patches of
 * 32 x 256 x 256 frames); other geometries return U2_ERR_UNSUPPORTED (use u2_patchify_f32_bf16 + u2_gemm_bf16). */
U2_API int u2_patch_embed_f32_bf16(const float* vol, const void* W, const float* bias, const void* pos, void* out,
                                   int64_t frames, int32_t d0, int32_t d1, int32_t d2, int32_t p0, int32_t p1, int32_t p2,
                                   int32_t N, int64_t out_frame_rows, void* stream);
/* dst[(r * row_stride + row_off), :] = vec for r in [0, n_rows)   (cls token rows, vit.py:116-118) */
U2_API int u2_set_rows_bf16(void* dst, const void* vec, int64_t n_rows, int64_t row_stride, int64_t row_off,
                            int32_t E, void* stream);
/* ViT sequence buffer dst [frames][Sp][E]: row 0 of every frame = cls, rows [S, Sp) = 0 (the padding that keeps the
 * frame stride a multiple of 16 bytes); rows 1..S-1 come from the patch-embed GEMM (vit.py:116-118 cls concat). */
U2_API int u2_vit_frame_rows_bf16(void* dst, const void* cls, int64_t frames, int32_t Sp, int32_t S, int32_t E,
                                  void* stream);
/* in[b][s][h][d] (element strides in_sb, in_ss, in_sh; d contiguous) -> out[b][h][d][s] with the s axis
 * padded to ld_out (zeros): a K-major V^T / X^T operand. Replaces the `split_heads` + `transpose(-2, -1)` copies of
 * src/model/u2tokenizer/rma.py:41-44,60 and tta.py:24-26,52. The path itself no longer calls it: V and the DiffTS token
 * matrix are consumed in place as MN-major operands (u2_gemm_desc.b_mn, u2_flash_attention_d64_bf16). */
U2_API int u2_transpose_heads_bf16(const void* in, void* out, int32_t B, int32_t S, int32_t H, int32_t Dh,
                                   int64_t in_sb, int64_t in_ss, int64_t in_sh, int64_t out_sb,
                                   int64_t out_sh, int64_t ld_out, void* stream);
/* SpatialPoolingProjector pooling (spatial_pooling_projector.py:38-46): token (a0,a1,a2) of frame f lives
 * at row f*in_frame_stride + in_off + (a0*g1+a1)*g2+a2 (row stride ldx); out [frames, n_out, E] dense.
 * sequence != 0 selects the avg_pool1d(ps^3) variant. */
U2_API int u2_spp_pool_bf16(const void* x, void* out, int64_t frames, int32_t g0, int32_t g1, int32_t g2,
                            int32_t ps, int32_t E, int64_t in_frame_stride, int64_t in_off, int64_t ldx,
                            int32_t sequence, void* stream);
/* Multi-scale token pooling, scales (1,2,4) over the token dim of x [B, K, E] -> out [B, K+K/2+K/4, E];
 * dynamic != 0: DynamicMultiScalePooling gate (svr.py:126-151) with gate_w [E] fp32, gate_bias scalar and a
 * [B,3] fp32 workspace; dynamic == 0: the plain concat (svr.py:175-184). */
U2_API int u2_multiscale_pool_bf16(const void* x, void* out, const float* gate_w, float gate_bias,
                                   float* logits_ws, int32_t B, int32_t K, int32_t E, int32_t dynamic,
                                   void* stream);
/* out[b][l] = vis[b][l-1] for 1 <= l <= n_vis (when vis != NULL) else table[ids[b][l]]
 * (u2_arch.py:118-121: embed_tokens gather + cat splice). ids int64. */
U2_API int u2_embed_splice_bf16(const int64_t* ids, const void* table, const void* vis, void* out, int32_t B,
                                int32_t L, int32_t E, int32_t n_vis, int64_t vocab, void* stream);

/* Small attention pieces ----------------------------------------------------------------------------------
 * temporal attention of the SVR layer (svr.py:33-36 with rma.py:60-73): qkv rows ordered (b, c, n),
 * columns [q|k|v] (E each); attends across the C <= 32 frames of each (b, n, head). rel_bias as above. */
U2_API int u2_temporal_attention_bf16(const void* qkv, void* out, int32_t B, int32_t C, int32_t N, int32_t H,
                                      int32_t dh, int64_t ld_qkv, int64_t ld_out, float scale,
                                      const float* rel_bias, int32_t rel_max, void* stream);

/* In-place rotate-half RoPE on the first n_q_heads + n_k_heads heads of every row (optionally preceded by
 * the per-head RMSNorm of Qwen3, HF modeling_qwen3.py:263-268), and optional KV-cache append
 * (caches [B, n_k_heads, Tmax, dh]). position(row) = pos0 + (row / pos_div) % pos_mod; pos0 may be read
 * from the device (pos0_dev). Also used for attn_type == "rope" of the tokenizer (rope.py:77-80). */
typedef struct u2_rope_desc {
  int64_t rows, ld;
  int32_t dh, n_q_heads, n_k_heads, n_v_heads;
  const float* q_norm_w;
  const float* k_norm_w;
  float eps;
  const float* inv_freq; /* [dh/2] fp32 */
  int32_t pos0, pos_div, pos_mod;
  const int32_t* pos0_dev;
  void* k_cache;
  void* v_cache;
  int32_t Tmax, rows_per_batch;
} u2_rope_desc;
U2_API int u2_rope_bf16(void* x, const u2_rope_desc* desc, void* stream);

/* One query token per sequence against the KV cache (GQA). q [B, Hq*dh] (row stride ldq), caches
 * [B, Hkv, Tmax, dh]; T valid keys (or *T_dev when T_dev != NULL). Replaces the HF eager attention at q_len == 1
 * (transformers models/qwen3/modeling_qwen3.py:252-291) inside generate() (src/model/language_model/u2llama.py:123-126);
 * unfused variant used by the CUDA-core decode path. */
U2_API int u2_decode_attention_bf16(const void* q, const void* k_cache, const void* v_cache, void* out,
                                    int32_t B, int32_t Hq, int32_t Hkv, int32_t dh, int32_t Tmax, int32_t T,
                                    const int32_t* T_dev, int64_t ldq, int64_t ldo, float scale, void* stream);

/* Decode-step linear (weight streaming, HBM-bound): y[b, n] = sum_k norm(x)[b, k] * w[n, k] (+ residual).
 * CUDA-core variant of the HF decoder Linears (+ Qwen3RMSNorm, modeling_qwen3.py:50-67) at q_len == 1 inside generate()
 * (src/model/language_model/u2llama.py:123-126) for shapes the tcgen05 path does not take (K % 64 != 0).
 * B <= 8. norm_gamma != NULL fuses the input RMSNorm; silu_pair != 0 treats rows (2j, 2j+1) of w as
 * (gate_j, up_j) and writes silu(gate) * up (N/2 outputs). */
typedef struct u2_gemv_desc {
  int32_t B, N, K;
  int64_t ldx, ldw, ldy, ldr;
  int32_t y_dtype;
  const void* residual;
  const float* norm_gamma;
  float norm_eps;
  int32_t silu_pair;
} u2_gemv_desc;
U2_API int u2_gemv_bf16(const void* x, const void* w, void* y, const u2_gemv_desc* desc, void* stream);
/* ids[b] = argmax_v logits[b, v] (first index on ties). scratch: uint64 [B], zero on entry and zero again on exit.
 * Replaces `torch.argmax(next_token_scores, dim=-1)` of HF GenerationMixin._sample with do_sample=False
 * (transformers generation/utils.py:2793; reference call src/model/language_model/u2llama.py:123-126). */
U2_API int u2_argmax_f32(const float* logits, int64_t* out, uint64_t* scratch, int32_t B, int32_t V, int64_t ld,
                         void* stream);

/* Decode-step linear on the tensor cores (swap-AB, stream-K over all SMs, TMA weight stream; HBM-bound):
 *   acc[b, n] = sum_k x[b, k] * w[n, k]                       B <= 16, K % 64 == 0
 *   v = acc * rsqrt(ssq_in[b] / K + eps)   (when ssq_in != NULL: fused RMSNorm, x must already carry gamma)
 *   silu_pair: rows (2j, 2j+1) of w are (gate_j, up_j):  y[b, j] = silu(v_2j) * v_2j+1
 *   else:      y[b, n] = v + residual[b, n];  optionally xg[b, n] = bf16(y * gamma_next[n]) and
 *              ssq_out[b] += sum_n y^2 (prepares the next fused norm); ssq_zero[0..15] is reset to 0.
 * ws: fp32 partial-sum slots (ws_elems floats; u2_dlinear_ws_elems(N, K) gives the size needed by the stream-K
 * schedule): every 32-bit word must hold 0xffffffff ("empty") on entry and does so again on exit;
 * counters: int32 [ceil(N/64)], zero on entry and on exit.
 * Replaces the HF decoder Linears at q_len == 1 (reference u2llama.py:123-126 -> GenerationMixin._sample). */
#define U2_DLIN_STREAMK128 0
#define U2_DLIN_TILES64 1
typedef struct u2_dlinear_desc {
  int32_t B, N, K;
  int64_t ldx, ldw, ldy, ldr, ldxg;
  int32_t y_dtype;
  float* ws;
  int32_t* counters;
  const float* ssq_in;
  float eps;
  const void* residual;
  int32_t silu_pair;
  const float* gamma_next;
  void* xg;
  float* ssq_out;
  float* ssq_zero;
  int32_t pdl; /* != 0: launch with programmatic stream serialization (weight prefetch overlaps the previous kernel) */
  void* dbg;   /* optional uint64 [grid][4][8] globaltimer stamps (tuning aid), normally NULL */
  int64_t ws_elems; /* capacity of ws in floats */
  /* multi-op launches only - fine-grained dataflow instead of a grid-wide wait before the first MMA of an op:
   * out_flags: int32 [ceil(N/128)] set to the step counter when a tile of THIS op is final;
   * dep_flags/dep_shift: flags of the op producing our x; k-block kb needs producer tile kb >> dep_shift
   * (1: 128 producer rows = 2 k-blocks; 0: a silu_pair producer, 128 rows = 64 activations = 1 k-block).
   * Ops linked this way must use disjoint ws / counters / x buffers (see engine.py). */
  const int32_t* dep_flags;
  int32_t dep_shift;
  int32_t* out_flags;
  int32_t sched; /* U2_DLIN_STREAMK128 (128-row tiles, stream-K + workspace reduction) or
                    U2_DLIN_TILES64 (whole 64-row tiles per CTA, no inter-CTA reduction) */
} u2_dlinear_desc;
U2_API int u2_dlinear_bf16(const void* x, const void* w, void* y, const u2_dlinear_desc* desc, void* stream);
U2_API int64_t u2_dlinear_ws_elems(int32_t N, int32_t K); /* fp32 elements of workspace for an N x K linear */
/* Up to four DEPENDENT decode linears in one launch (o_proj -> gate|up -> down -> next qkv; the Linears of
 * Qwen3DecoderLayer.forward, transformers models/qwen3/modeling_qwen3.py:305-336, at q_len == 1): software grid
 * barriers between them (gridbar: uint32[4], monotonically increasing; target = *step_dev * #SMs, step_dev is
 * the per-step counter u2_decode_embed_bf16 bumps), the weight stream of op i+1 is prefetched while op i
 * drains. x[i], w[i], y[i], descs[i] as for u2_dlinear_bf16. */
/* Optional L2 look-ahead for u2_dlinear_multi_bf16: lookahead_units = per-CTA number of 16 KB weight tiles
 * prefetched into L2 beyond the shared-memory ring while an in-launch dependency is pending; w[j] (N[j] x K[j],
 * row stride ldw[j]) = weights the NEXT launch streams first, units[j] leading tiles per CTA are prefetched when
 * this launch has issued all of its own loads (covers the launch gap / the attention kernel in between). */
typedef struct u2_dlinear_next {
  int32_t pre_stages;   /* ring stages of the next op's weights requested before its dependency resolves (0 = all) */
  int32_t lookahead_units;
  int32_t n;            /* 0..2 */
  const void* w[2];
  int32_t N[2], K[2];
  int64_t ldw[2];
  int32_t units[2];
} u2_dlinear_next;
U2_API int u2_dlinear_multi_bf16(const void* const* x, const void* const* w, void* const* y,
                                 const u2_dlinear_desc* descs, int32_t n_ops, uint32_t* gridbar,
                                 const int32_t* step_dev, int32_t pdl, const u2_dlinear_next* next, void* stream);
/* x[b] = table[ids[b]]; xg[b] = bf16(x * gamma); ssq[b] = sum x^2; ssq_zero[b] = 0; *step_counter += 1
 * (start of a decode step; step_counter may be NULL). Replaces `embed_tokens(input_ids)` of the cached decode step
 * (transformers models/qwen3/modeling_qwen3.py:392) + the first half of the first layer's input RMSNorm. */
U2_API int u2_decode_embed_bf16(const int64_t* ids, const void* table, const float* gamma, void* x, void* xg,
                                float* ssq, float* ssq_zero, int32_t* step_counter, int32_t B, int32_t E,
                                int64_t vocab, void* stream);

/* Fused decode-step attention (one launch per layer): per-head RMSNorm (optional) + RoPE of the new q/k,
 * KV-cache append at position pos (or *pos_dev) and GQA attention over the pos + 1 cached keys.
 * qkv [B, (Hq + 2 Hkv) * dh] raw projections; caches [B, Hkv, Tmax, dh]; out [B, Hq * dh].
 * Replaces HF modeling_qwen3.py:263-288 at q_len == 1. */
typedef struct u2_fused_decode_desc {
  int32_t B, Hq, Hkv, dh, Tmax, pos;
  const int32_t* pos_dev;
  int64_t ldq, ldo;
  const float* q_norm_w;
  const float* k_norm_w;
  float eps;
  const float* inv_freq;
  float scale;
  int32_t kv_splits;   /* 0/1: one CTA per (sequence, KV head); 2/4/8: a cluster of that many CTAs splits the cached
                          keys and merges over distributed shared memory (fills the SMs when B * Hkv is small) */
  int32_t pdl;         /* != 0 (split-KV variant only): launch with programmatic stream serialisation - position read
                          and K/V prefetch overlap the tail of the preceding kernel, which must not write the cache */
} u2_fused_decode_desc;
U2_API int u2_decode_attention_fused_bf16(const void* qkv, void* k_cache, void* v_cache, void* out,
                                          const u2_fused_decode_desc* desc, void* stream);

/* Row-wise top-k of fp32 scores, sorted descending (ties: lower index first), as torch.topk in the hard
 * TokenSelection (reference svr.py:75-91). out_idx[r, i] = index + r * idx_offset_per_row (int64). T <= 16384. */
U2_API int u2_topk_rows_f32(const float* scores, int64_t* out_idx, int32_t rows, int32_t T, int32_t K, int64_t ld,
                            int64_t idx_offset_per_row, void* stream);

/* Fused attention forward, head_dim 64, non-causal (ViT3D): out = softmax(q k^T * scale) v, scores never leave
 * the SM (tcgen05: S and PV partials in TMEM, P through swizzled shared memory).
 * q [B, Sq, H, 64], k [B, Sk, H, 64], v [B, Sk, H, 64] as strided views (element strides *_sb batch, *_ss token,
 * *_sh head; d contiguous) - typically the three slices of one fused QKV activation; v is consumed as stored (MN-major
 * B operand of the PV product, no transposed copy); out [B, Sq, H*64] (strides out_sb, out_ss). All strides multiples
 * of 8 elements.
 * Replaces MONAI SABlock einsum/softmax/einsum (reference vit.py:100-105,120-122). */
typedef struct u2_fa_desc {
  int32_t B, H, Sq, Sk, dh;
  float scale;
  int64_t q_sb, q_ss, q_sh;
  int64_t k_sb, k_ss, k_sh;
  int64_t v_sb, v_ss, v_sh;
  int64_t out_sb, out_ss;
  float* lse; /* optional fp32 [B, H, Sq]: log-sum-exp of the scaled score rows (training: the backward rebuilds the
                 probabilities as exp(scale * q.k - lse) in the score GEMM's epilogue); NULL at inference */
} u2_fa_desc;
U2_API int u2_flash_attention_d64_bf16(const void* q, const void* k, const void* v, void* out, const u2_fa_desc* desc,
                                       void* stream);

/* Sampled decoding head: ids[b] ~ multinomial(top_p(top_k(softmax(logits[b] / temperature)))) - the HF warper chain
 * behind generate(do_sample=True, temperature, top_k, top_p) used by the reference's eval scripts
 * (eval/mrg.py:74-75). top_k <= 0 disables top-k, top_p = 1 disables nucleus filtering. Counter-based RNG
 * keyed by (seed, step or *step_dev, row). */
U2_API int u2_sample_f32(const float* logits, int64_t* out, int32_t B, int32_t V, int64_t ld, float temperature,
                         int32_t top_k, float top_p, uint64_t seed, const int32_t* step_dev, int32_t step,
                         void* stream);

/* Same head (HF _sample with do_sample=True, generation/utils.py:2791; reference eval/mrg.py:74-75,
 * src/train/dpo_u2trainer.py:71-79) with its parameters in DEVICE memory (24 bytes): a captured decode step reads them at
 * replay time, so a new seed / temperature / top-k / top-p per request (HF generate kwargs) needs one small
 * host-to-device copy, not a new capture. The caller validates temperature > 0 and 0 < top_p <= 1. */
typedef struct u2_sample_params {
  float temperature;
  int32_t top_k;
  float top_p;
  int32_t reserved;
  uint64_t seed;
} u2_sample_params;
U2_API int u2_sample_dev_f32(const float* logits, int64_t* out, int32_t B, int32_t V, int64_t ld,
                             const u2_sample_params* params_dev, const int32_t* step_dev, int32_t step, void* stream);

/* Fused lm_head + selective log-softmax (the DPO / SFT log-probability head) -----------------------------------
 * logp[r] = log_softmax(hidden[r] . W^T)[labels[r]]  (0 where labels[r] < 0) without materialising the [R, V] logits:
 * the GEMM's epilogue reduces every 128-column half tile to (max, sum exp, sum) per row, a second small kernel merges
 * them. Replaces lm_head + `selective_log_softmax(logits, labels)` in u2DPOTrainer.concatenated_forward
 * (src/train/dpo_u2trainer.py:267-300; [2B, 1024, 151936] logits = 622 MB per pair in bf16) and HF's
 * ForCausalLMLoss in forward(labels=...) (src/model/language_model/u2llama.py:76-87).
 * hidden [R, E] bf16 (row stride ldh), W [V, E] bf16 (row stride ldw), labels int64 [R]; optional outputs: lse [R]
 * (log-sum-exp), logit_sum [R] (sum of the row's V logits: the trainer's mean_*_logits statistics,
 * dpo_u2trainer.py:343-350), nll_acc [2] (+= sum of -logp and the number of labelled rows; the caller zeroes it).
 * ws: workspace of u2_logprob_ws_bytes(R, V) bytes, 16-byte aligned. */
typedef struct u2_logprob_desc {
  int32_t R, V, E;
  int64_t ldh, ldw;
  const int64_t* labels;
  void* ws;
  int64_t ws_bytes;
  float* lse;
  float* logit_sum;
  float* nll_acc;
} u2_logprob_desc;
U2_API int64_t u2_logprob_ws_bytes(int32_t R, int32_t V);
U2_API int u2_lmhead_logprob_bf16(const void* hidden, const void* W, float* logp, const u2_logprob_desc* desc,
                                  void* stream);

/* Volume preprocessing in front of the path ------------------------------------------------
 * The reference's u2Transform.adaptive_resize (src/utils/u2Transform.py:62-122, validation pipeline :47-56) on a volume
 * that is already on the device: ScaleIntensityRangePercentiles(lower, upper -> [0, 1], clip) -> CropForeground (> 0) ->
 * anti-aliased trilinear resize (align_corners) so that the larger in-plane side becomes `target` (depth kept when it is
 * <= pad_depth, resized to pad_depth otherwise) -> zero pad to [pad_depth, target, target].
 * vol: fp32 [D, H, W] (the reference's data[0] after get_fdata().transpose(2, 0, 1)); out: fp32 [pad_depth, target,
 * target] (viewed as [pad_depth / 32, 32, target, target] it is the `images` tensor of one study). info (device memory)
 * receives the data-dependent quantities; nothing is synchronised with the host. status != 0 flags the inputs on which
 * the reference itself fails or degenerates (the output is then all zeros, except U2_PP_FLAT_INTENSITY). */
#define U2_PP_OK 0
#define U2_PP_EMPTY_FOREGROUND 1  /* no voxel above the lower percentile */
#define U2_PP_DEGENERATE_SHAPE 2  /* a resized extent of 0, or an anti-aliasing kernel wider than 129 taps */
#define U2_PP_FLAT_INTENSITY 3    /* a_min == a_max: MONAI returns img - a_min unscaled */
typedef struct u2_preprocess_info {
  double a_min, a_max;      /* the two percentiles (np.percentile, linear interpolation, float64) */
  int32_t lo[3], hi[3];     /* foreground box [lo, hi) on (D, H, W) */
  int32_t out[3];           /* resized extents on (D, H, W) before padding */
  float sigma[3];           /* anti-aliasing sigma per axis (0: none) */
  int32_t tail[3];          /* Gaussian half width in taps */
  int32_t status;           /* U2_PP_* */
} u2_preprocess_info;
typedef struct u2_preprocess_desc {
  int32_t D, H, W;
  int32_t target, pad_depth;
  double lower_pct, upper_pct;
  void* ws;                 /* u2_preprocess_ws_bytes(D, H, W) bytes, 256-byte aligned */
  int64_t ws_bytes;
} u2_preprocess_desc;
U2_API int64_t u2_preprocess_ws_bytes(int32_t D, int32_t H, int32_t W);
U2_API int u2_preprocess_volume_f32(const float* vol, float* out, u2_preprocess_info* info,
                                    const u2_preprocess_desc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* U2B200_H_ */
