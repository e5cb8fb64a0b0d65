// Fused (flash-style) attention forward on tcgen05 for head_dim 64, non-causal - the ViT3D tower's attention
// (12 heads, S = 2049 tokens per frame): O = softmax(Q K^T * scale) V without materialising the S x S scores.
//
// One CTA owns a 128-row query tile of one (frame, head) and walks the keys in tiles of 128:
//   warp 0   : TMA producer   Q once, then (K_j, V_j) into a 2-deep ring
//   warp 1   : MMA issuer     S_j = Q K_j^T  (128 x 128 x 64, fp32 in TMEM, double buffered) and
//                             PV_j = P_j V_j (128 x 64 x 128); S_{j+1} is issued BEFORE waiting for P_j, so the
//                             tensor pipe works on the next scores while the CUDA cores do the softmax of tile j
//   warp 2   : TMEM allocator (512 columns: 2 x 128 for S, 2 x 64 for the PV partials)
//   warps 4-7: softmax        thread == query row: tcgen05.ld the S row, online max/sum in the log2 domain,
//                             P (bf16) written to shared memory in the swizzled K-major UMMA layout, running O kept
//                             in registers and rescaled there (no TMEM read-modify-write)
// Q / K / V are 4-D TMA views of the fused QKV activation ([frame, token, 3, head, d]). V is consumed AS STORED: a
// {64 d, 128 keys} box with the 128-byte swizzle is the canonical MN-major UMMA operand layout (one 64-wide N chunk,
// groups of 8 key rows 1024 B apart), so the PV product takes it through an MN-major B descriptor - no transposed
// copy of V (round 1 ran a transpose kernel per block: 86 us x 12 at the bench configuration).
//
// Replaces MONAI SABlock's einsum / softmax / einsum (reference call site src/model/multimodal_encoder/vit.py:
// 100-105,120-122), which materialises a [frames*12, 2049, 2049] fp32 score tensor per block.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"
#include "u2b200.h"

namespace u2 {

constexpr int kFaDh = 64;
constexpr int kFaBM = 128;   // query rows per CTA
constexpr int kFaBN = 128;   // keys per tile
constexpr int kFaThreads = 256;
constexpr int kFaQBytes = kFaBM * kFaDh * 2;          // 16 KB
constexpr int kFaKBytes = kFaBN * kFaDh * 2;          // 16 KB
constexpr int kFaVBytes = kFaBN * kFaDh * 2;          // 16 KB (128 key rows of 128 B)
constexpr int kFaPBytes = kFaBM * kFaBN * 2;          // 32 KB (two 64-key slabs of 16 KB)
constexpr int kFaSmem = kFaQBytes + 2 * kFaKBytes + 2 * kFaVBytes + 2 * kFaPBytes + 1024 + 256;
constexpr int kFaTmemCols = 512;

struct FaArgs {
  int Sq, Sk;              // valid query rows / keys per (batch, head)
  float scale_log2e;       // softmax scale * log2(e)
  __nv_bfloat16* out;      // [b][s][h*64 + d]
  long long out_sb, out_ss;  // element strides of batch and token
  float* lse;              // optional [b][h][Sq]: log-sum-exp of the scaled scores (the attention backward rebuilds P from it)
  int H;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kFaThreads, 1)
fa_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                      const __grid_constant__ CUtensorMap tmap_v, const FaArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kFaQBytes;            // [2]
  uint8_t* sV = sK + 2 * kFaKBytes;        // [2]
  uint8_t* sP = sV + 2 * kFaVBytes;        // [2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kFaPBytes);
  uint64_t* q_full = bars;          // 1
  uint64_t* kv_full = bars + 1;     // 2
  uint64_t* kv_empty = bars + 3;    // 2
  uint64_t* s_full = bars + 5;      // 2
  uint64_t* s_empty = bars + 7;     // 2
  uint64_t* p_full = bars + 9;      // 2
  uint64_t* p_empty = bars + 11;    // 2
  uint64_t* o_full = bars + 13;     // 2
  uint64_t* o_empty = bars + 15;    // 2
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const int q_tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int q0 = q_tile * kFaBM;
  const int J = (p.Sk + kFaBN - 1) / kFaBN;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp_idx == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);   // one arrival per softmax WARP: 128 per-thread arrivals on one mbarrier serialise (~10 ns each)
      mbar_init(&p_full[i], 4);
      mbar_init(&p_empty[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) tmem_alloc<kFaTmemCols>(tmem_base_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const uint32_t tS = tmem_base;         // S[buf] at columns buf * 128
  const uint32_t tO = tmem_base + 256;   // O[buf] at columns 256 + buf * 64

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kFaQBytes);
      tma_load_4d(sQ, &tmap_q, q_full, 0, q0, h, b);
      for (int j = 0; j < J; ++j) {
        const int buf = j & 1, n = j >> 1;
        mbar_wait(&kv_empty[buf], (n & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[buf], kFaKBytes + kFaVBytes);
        tma_load_4d(sK + buf * kFaKBytes, &tmap_k, &kv_full[buf], 0, j * kFaBN, h, b);
        tma_load_4d(sV + buf * kFaVBytes, &tmap_v, &kv_full[buf], 0, j * kFaBN, h, b);
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(kFaBM, kFaBN);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kFaBM, kFaDh) | (1u << 16);  // B (= V) is MN-major
      const uint64_t q_desc = umma_desc_kmajor_sw128(smem_u32(sQ));
      auto issue_qk = [&](int j) {
        const int buf = j & 1, n = j >> 1;
        mbar_wait(&kv_full[buf], n & 1);
        mbar_wait(&s_empty[buf], (n & 1) ^ 1);
        tc_fence_after();
        const uint64_t k_desc = umma_desc_kmajor_sw128(smem_u32(sK + buf * kFaKBytes));
#pragma unroll
        for (int k = 0; k < kFaDh / 16; ++k) umma_f16(tS + buf * kFaBN, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k != 0);
        umma_commit(&s_full[buf]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < J; ++j) {
        if (j + 1 < J) issue_qk(j + 1);  // next scores first: overlaps the softmax of tile j
        const int buf = j & 1, n = j >> 1;
        mbar_wait(&p_full[buf], n & 1);
        mbar_wait(&o_empty[buf], (n & 1) ^ 1);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(sP + buf * kFaPBytes);
        const uint32_t v_addr = smem_u32(sV + buf * kFaVBytes);
#pragma unroll
        for (int k = 0; k < kFaBN / 16; ++k) {
          const uint64_t a_desc = umma_desc_kmajor_sw128(p_addr + (k >> 2) * (kFaPBytes / 2)) + 2 * (k & 3);
          // 16 keys per UMMA_K step = two 1024-byte groups of 8 key rows (descriptor start address in 16-byte units)
          const uint64_t b_desc = umma_desc_mnmajor_sw128(v_addr, kFaVBytes) + 128 * k;
          umma_f16(tO + buf * kFaDh, a_desc, b_desc, idesc_pv, k != 0);
        }
        umma_commit(&o_full[buf]);
        umma_commit(&kv_empty[buf]);
        umma_commit(&p_empty[buf]);
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== softmax / output (thread == query row) =====================
    const int q = warp_idx - 4;
    const int r = threadIdx.x - 128;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    float m = -INFINITY, l = 0.f;
    float o[kFaDh];
#pragma unroll
    for (int d = 0; d < kFaDh; ++d) o[d] = 0.f;

    for (int j = 0; j < J; ++j) {
      const int buf = j & 1, n = j >> 1;
      const int nk = min(kFaBN, p.Sk - j * kFaBN);  // valid keys in this tile
      mbar_wait(&s_full[buf], n & 1);
      tc_fence_after();
      // ---- pass 1: row maximum (scores stay in TMEM, they are read again in pass 2)
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < kFaBN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tS + buf * kFaBN + lane_off + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c + i < nk) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m, mx * p.scale_log2e);
      const float alpha = ex2_approx(m - m_new);  // m == -inf on the first tile -> 0
      // ---- retire the previous tile's PV partial before rescaling the running output
      if (j > 0) {
        const int pb = (j - 1) & 1, pn = (j - 1) >> 1;
        mbar_wait(&o_full[pb], pn & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kFaDh; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tO + pb * kFaDh + lane_off + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c + i] += __uint_as_float(v[i]);
        }
        tc_fence_before();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&o_empty[pb]);
      }
      // ---- pass 2: probabilities -> shared memory (bf16, 128-byte swizzled K-major rows)
      mbar_wait(&p_empty[buf], (n & 1) ^ 1);
      float rowsum = 0.f;
      uint8_t* prow = sP + buf * kFaPBytes + r * 128;
#pragma unroll 1
      for (int c = 0; c < kFaBN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tS + buf * kFaBN + lane_off + c, v);
        tmem_ld_wait();
        float pr[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float e = (c + i < nk) ? ex2_approx(__uint_as_float(v[i]) * p.scale_log2e - m_new) : 0.f;
          pr[i] = e;
          rowsum += e;
        }
        uint8_t* slab = prow + (c >> 6) * (kFaPBytes / 2);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(pr[g * 8 + 2 * e], pr[g * 8 + 2 * e + 1]);
          const int chunk = ((c & 63) >> 3) + g;           // 16-byte chunk index inside the 128-byte row
          *reinterpret_cast<uint4*>(slab + ((chunk ^ (r & 7)) << 4)) = pk;
        }
      }
      tc_fence_before();
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&s_empty[buf]);  // S[buf] may be overwritten by the scores of tile j + 2
      l = l * alpha + rowsum;
#pragma unroll
      for (int d = 0; d < kFaDh; ++d) o[d] *= alpha;
      m = m_new;
      fence_proxy_async_smem();    // generic-proxy P stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&p_full[buf]);
    }
    // ---- last PV partial, normalise, store
    {
      const int pb = (J - 1) & 1, pn = (J - 1) >> 1;
      mbar_wait(&o_full[pb], pn & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < kFaDh; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tO + pb * kFaDh + lane_off + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c + i] += __uint_as_float(v[i]);
      }
    }
    const int row = q0 + r;
    if (row < p.Sq) {
      const float inv = 1.f / l;
      if (p.lse) p.lse[((long long)b * p.H + h) * p.Sq + row] = (m + log2f(l)) * 0.6931471805599453f;
      __nv_bfloat16* dst = p.out + (long long)b * p.out_sb + (long long)row * p.out_ss + h * kFaDh;
#pragma unroll
      for (int d = 0; d < kFaDh; d += 8) {
        uint4 pk;
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(o[d + 2 * e] * inv, o[d + 2 * e + 1] * inv);
        *reinterpret_cast<uint4*>(dst + d) = pk;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc<kFaTmemCols>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------------
// Two query tiles per CTA (256 rows), two softmax warp-groups: the tile-A group exponentiates while the tensor
// pipe runs PV / the next QK^T of tile B and vice versa (ping-pong). The softmax is issue-bound with one warp per
// scheduler; the second group doubles the CUDA-core side and both tiles share every K / V^T tile brought in by TMA.
// ------------------------------------------------------------------------------------------------
constexpr int kFa2Threads = 384;
constexpr int kFa2Smem = 2 * kFaQBytes + 2 * kFaKBytes + 2 * kFaVBytes + 2 * kFaPBytes + 1024 + 256;

__global__ void __launch_bounds__(kFa2Threads, 1)
fa_fwd2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                       const __grid_constant__ CUtensorMap tmap_v, const FaArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                       // [2 tiles]
  uint8_t* sK = sQ + 2 * kFaQBytes;         // [2 stages]
  uint8_t* sV = sK + 2 * kFaKBytes;         // [2 stages]
  uint8_t* sP = sV + 2 * kFaVBytes;         // [2 tiles]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kFaPBytes);
  uint64_t* q_full = bars;          // 1
  uint64_t* kv_full = bars + 1;     // 2 (stage)
  uint64_t* kv_empty = bars + 3;    // 2
  uint64_t* s_full = bars + 5;      // 2 (tile)
  uint64_t* s_empty = bars + 7;     // 2
  uint64_t* p_full = bars + 9;      // 2
  uint64_t* p_empty = bars + 11;    // 2
  uint64_t* o_full = bars + 13;     // 2
  uint64_t* o_empty = bars + 15;    // 2
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp_idx = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * 2 * kFaBM;
  const int J = (p.Sk + kFaBN - 1) / kFaBN;
  const bool has_b = (q0 + kFaBM) < p.Sq;  // second query tile holds at least one valid row

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp_idx == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 4);   // one arrival per softmax WARP: 128 per-thread arrivals on one mbarrier serialise (~10 ns each)
      mbar_init(&p_full[i], 4);
      mbar_init(&p_empty[i], 1);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) tmem_alloc<kFaTmemCols>(tmem_base_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const uint32_t tS = tmem_base;         // S[tile] at columns tile * 128
  const uint32_t tO = tmem_base + 256;   // O[tile] at columns 256 + tile * 64

  if (warp_idx == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, has_b ? 2 * kFaQBytes : kFaQBytes);
      tma_load_4d(sQ, &tmap_q, q_full, 0, q0, h, b);
      if (has_b) tma_load_4d(sQ + kFaQBytes, &tmap_q, q_full, 0, q0 + kFaBM, h, b);
      for (int j = 0; j < J; ++j) {
        const int st = j & 1, n = j >> 1;
        mbar_wait(&kv_empty[st], (n & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], kFaKBytes + kFaVBytes);
        tma_load_4d(sK + st * kFaKBytes, &tmap_k, &kv_full[st], 0, j * kFaBN, h, b);
        tma_load_4d(sV + st * kFaVBytes, &tmap_v, &kv_full[st], 0, j * kFaBN, h, b);
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(kFaBM, kFaBN);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kFaBM, kFaDh) | (1u << 16);  // B (= V) is MN-major
      const int ntile = has_b ? 2 : 1;
      auto issue_qk = [&](int t, int j) {
        const int st = j & 1;
        mbar_wait(&s_empty[t], (j & 1) ^ 1);  // softmax group t finished reading S[t] of key tile j - 1
        tc_fence_after();
        const uint64_t q_desc = umma_desc_kmajor_sw128(smem_u32(sQ + t * kFaQBytes));
        const uint64_t k_desc = umma_desc_kmajor_sw128(smem_u32(sK + st * kFaKBytes));
#pragma unroll
        for (int k = 0; k < kFaDh / 16; ++k) umma_f16(tS + t * kFaBN, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k != 0);
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int j) {
        const int st = j & 1;
        mbar_wait(&p_full[t], j & 1);
        mbar_wait(&o_empty[t], (j & 1) ^ 1);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(sP + t * kFaPBytes);
        const uint32_t v_addr = smem_u32(sV + st * kFaVBytes);
#pragma unroll
        for (int k = 0; k < kFaBN / 16; ++k) {
          const uint64_t a_desc = umma_desc_kmajor_sw128(p_addr + (k >> 2) * (kFaPBytes / 2)) + 2 * (k & 3);
          // 16 keys per UMMA_K step = two 1024-byte groups of 8 key rows (descriptor start address in 16-byte units)
          const uint64_t b_desc = umma_desc_mnmajor_sw128(v_addr, kFaVBytes) + 128 * k;
          umma_f16(tO + t * kFaDh, a_desc, b_desc, idesc_pv, k != 0);
        }
        umma_commit(&o_full[t]);
        umma_commit(&p_empty[t]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      for (int t = 0; t < ntile; ++t) issue_qk(t, 0);
      for (int j = 0; j < J; ++j) {
        const bool more = (j + 1 < J);
        if (more) mbar_wait(&kv_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
        for (int t = 0; t < ntile; ++t) {
          issue_pv(t, j);
          if (t == ntile - 1) umma_commit(&kv_empty[j & 1]);  // every MMA that reads stage j & 1 has been issued
          if (more) issue_qk(t, j + 1);  // tile t's next scores: overlaps the other group's softmax
        }
      }
    }
  } else if (warp_idx >= 4) {
    const int t = (warp_idx - 4) >> 2;          // softmax group == query tile
    if (t == 0 || has_b) {
      const int q = (warp_idx - 4) & 3;
      const int r = (threadIdx.x - 128) & 127;  // row inside the tile
      const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
      float m = -INFINITY, l = 0.f;
      float o[kFaDh];
#pragma unroll
      for (int d = 0; d < kFaDh; ++d) o[d] = 0.f;
      uint8_t* prow = sP + t * kFaPBytes + r * 128;

      for (int j = 0; j < J; ++j) {
        const int nk = min(kFaBN, p.Sk - j * kFaBN);
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < kFaBN; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tS + t * kFaBN + lane_off + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i < nk) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        const float m_new = fmaxf(m, mx * p.scale_log2e);
        const float alpha = ex2_approx(m - m_new);
        if (j > 0) {
          mbar_wait(&o_full[t], (j - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < kFaDh; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tO + t * kFaDh + lane_off + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[c + i] += __uint_as_float(v[i]);
          }
          tc_fence_before();
          __syncwarp();
          if ((threadIdx.x & 31) == 0) mbar_arrive(&o_empty[t]);
        }
        mbar_wait(&p_empty[t], (j & 1) ^ 1);
        float rowsum = 0.f;
#pragma unroll 1
        for (int c = 0; c < kFaBN; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tS + t * kFaBN + lane_off + c, v);
          tmem_ld_wait();
          float pr[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = (c + i < nk) ? ex2_approx(__uint_as_float(v[i]) * p.scale_log2e - m_new) : 0.f;
            pr[i] = e;
            rowsum += e;
          }
          uint8_t* slab = prow + (c >> 6) * (kFaPBytes / 2);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 pk;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(pr[g * 8 + 2 * e], pr[g * 8 + 2 * e + 1]);
            const int chunk = ((c & 63) >> 3) + g;
            *reinterpret_cast<uint4*>(slab + ((chunk ^ (r & 7)) << 4)) = pk;
          }
        }
        tc_fence_before();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&s_empty[t]);
        l = l * alpha + rowsum;
#pragma unroll
        for (int d = 0; d < kFaDh; ++d) o[d] *= alpha;
        m = m_new;
        fence_proxy_async_smem();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&p_full[t]);
      }
      mbar_wait(&o_full[t], (J - 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < kFaDh; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tO + t * kFaDh + lane_off + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c + i] += __uint_as_float(v[i]);
      }
      const int row = q0 + t * kFaBM + r;
      if (row < p.Sq) {
        const float inv = 1.f / l;
        if (p.lse) p.lse[((long long)b * p.H + h) * p.Sq + row] = (m + log2f(l)) * 0.6931471805599453f;
        __nv_bfloat16* dst = p.out + (long long)b * p.out_sb + (long long)row * p.out_ss + h * kFaDh;
#pragma unroll
        for (int d = 0; d < kFaDh; d += 8) {
          uint4 pk;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(o[d + 2 * e] * inv, o[d + 2 * e + 1] * inv);
          *reinterpret_cast<uint4*>(dst + d) = pk;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc<kFaTmemCols>(tmem_base);
  }
}

}  // namespace u2

extern "C" U2_API int u2_flash_attention_d64_bf16(const void* q, const void* k, const void* v, void* out,
                                                  const u2_fa_desc* d, void* stream) {
  using namespace u2;
  if (!q || !k || !v || !out || !d) return set_error(U2_ERR_ARG, "flash_attention: null pointer");
  if (d->dh != kFaDh) return set_error(U2_ERR_UNSUPPORTED, "flash_attention: head_dim %d (this kernel: 64)", d->dh);
  if (d->B <= 0 || d->H <= 0 || d->Sq <= 0 || d->Sk <= 0) return set_error(U2_ERR_ARG, "flash_attention: bad extents");
  if (d->B > 65535 || d->H > 65535) return set_error(U2_ERR_ARG, "flash_attention: B, H must be <= 65535");
  if ((d->q_ss & 7) || (d->q_sh & 7) || (d->q_sb & 7) || (d->k_ss & 7) || (d->k_sh & 7) || (d->k_sb & 7) || (d->v_ss & 7) ||
      (d->v_sh & 7) || (d->v_sb & 7) || (d->out_ss & 7))
    return set_error(U2_ERR_ARG, "flash_attention: strides must be multiples of 8 elements");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fa_fwd_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaSmem);
    if (e != cudaSuccess) return set_error(U2_ERR_CUDA, "flash_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  CUtensorMap tq, tk, tv;
  int rc = make_tmap_bf16_4d(&tq, q, kFaDh, d->Sq, d->H, d->B, d->q_ss, d->q_sh, d->q_sb, kFaDh, kFaBM);
  if (rc) return rc;
  rc = make_tmap_bf16_4d(&tk, k, kFaDh, d->Sk, d->H, d->B, d->k_ss, d->k_sh, d->k_sb, kFaDh, kFaBN);
  if (rc) return rc;
  // V as stored: dims {d, token, head, batch}, box {64 d, 128 keys} = the MN-major B operand of the PV product
  rc = make_tmap_bf16_4d(&tv, v, kFaDh, d->Sk, d->H, d->B, d->v_ss, d->v_sh, d->v_sb, kFaDh, kFaBN);
  if (rc) return rc;
  FaArgs a;
  a.Sq = d->Sq; a.Sk = d->Sk;
  a.scale_log2e = d->scale * 1.4426950408889634f;
  a.out = reinterpret_cast<__nv_bfloat16*>(out);
  a.out_sb = d->out_sb; a.out_ss = d->out_ss;
  a.lse = d->lse; a.H = d->H;
  const bool two_tiles = d->Sq > kFaBM && !getenv("U2_FA_V1");
  if (two_tiles) {
    static bool configured2 = false;
    if (!configured2) {
      cudaError_t e = cudaFuncSetAttribute(fa_fwd2_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFa2Smem);
      if (e != cudaSuccess) return set_error(U2_ERR_CUDA, "flash_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      configured2 = true;
    }
    dim3 grid2((unsigned)((d->Sq + 2 * kFaBM - 1) / (2 * kFaBM)), (unsigned)d->H, (unsigned)d->B);
    fa_fwd2_tcgen05_kernel<<<grid2, kFa2Threads, kFa2Smem, reinterpret_cast<cudaStream_t>(stream)>>>(tq, tk, tv, a);
    U2_CHECK_LAUNCH("flash_attention (2 query tiles)");
    return U2_OK;
  }
  dim3 grid((unsigned)((d->Sq + kFaBM - 1) / kFaBM), (unsigned)d->H, (unsigned)d->B);
  fa_fwd_tcgen05_kernel<<<grid, kFaThreads, kFaSmem, reinterpret_cast<cudaStream_t>(stream)>>>(tq, tk, tv, a);
  U2_CHECK_LAUNCH("flash_attention");
  return U2_OK;
}
