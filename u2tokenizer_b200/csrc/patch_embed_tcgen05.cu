// Fused 3-D patch embedding (K1 / K2 of SURVEY.md section 2d): fp32 CT volume -> bf16 ViT tokens in ONE kernel.
//
//   x[f, 1 + t, :] = bf16( patch(f, t) [1024 fp32 voxels] . W^T + bias + pos[t] )      t = (h, w, d) over the 8 x 16 x 16 grid
//
// The einops gather "b c (h p1) (w p2) (d p3) -> b (h w d) (p1 p2 p3 c)" of MONAI's PatchEmbeddingBlock (reference
// src/model/multimodal_encoder/vit.py:90-99,115) never materialises: a 128-token M tile is the 8 (w) x 16 (d) tokens of one
// (frame, h) slab, and k-block kb = (p1, 4 consecutive p2, 16 p3) of those tokens is ONE 5-D TMA box
// {256 (D2), 4 (p2), 8 (w), 1 (D0 = 4 h + p1), 1 (frame)} = 32 KB of fp32 that lands in shared memory as
// [w][p2][256]. A converter warp-group (thread = token) rewrites it as the 128 x 64 bf16 K-major SWIZZLE_128B A tile the
// tensor core wants (bank-conflict-free: the four 16-byte pieces of a thread's 64-byte segment are read in an order
// rotated by the token index), tcgen05.mma accumulates in TMEM, the epilogue adds bias + position embedding and stores the
// rows behind the cls row through a swizzled staging tile (whole 64-byte segments per store). The unfused path wrote and
// re-read 134 MB of bf16 im2col rows per 4 volumes; here the volume is read once (the three N tiles of an M tile run on
// neighbouring CTAs in the same time window, so the second and third read of a slab hit L2).
//
// Warp roles (480 threads, 1 CTA / SM, persistent over (m tile, n tile)):
//   warp 0      TMA producer: fp32 slabs -> 3-stage staging ring (96 KB in flight), W tiles (256 x 64 bf16) -> 2-stage ring
//   warp 1      MMA issuer (one thread), 128 x 256 x 16 UMMAs, fp32 accumulators double-buffered in TMEM
//   warp 2      TMEM allocator
//   warps 3-6   converter warp-group (fp32 staging -> swizzled bf16 A ring, 2 stages)
//   warps 7-14  epilogue: two warps per TMEM lane quarter (quarter = warp % 4), each drains half of the 256 columns; the
//               position-table segment of a chunk is requested before the TMEM load is waited for (with K = 1024 a tile's
//               mainloop is only ~7 us: the first version, 4 epilogue warps with serial wait -> load -> wait chunks, was
//               epilogue-bound at 419 us per 4 volumes)
#include <cuda_bf16.h>
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"
#include "u2b200.h"

namespace u2 {

constexpr int kPeM = 128, kPeN = 256, kPeK = 64;
constexpr int kPeStg = 3, kPeA = 2, kPeB = 2;
constexpr int kPeStgBytes = 128 * 64 * 4;   // 32 KB fp32 slab
constexpr int kPeABytes = kPeM * kPeK * 2;  // 16 KB
constexpr int kPeBBytes = kPeN * kPeK * 2;  // 32 KB
constexpr int kPeEpiBytes = 8 * 2048;   // one 32 x 32 bf16 staging tile per epilogue warp
constexpr int kPeSmem = kPeA * kPeABytes + kPeB * kPeBBytes + kPeStg * kPeStgBytes + kPeEpiBytes + 1024 + 256;
constexpr int kPeThreads = 480;
constexpr int kPeConvWarp0 = 3, kPeEpiWarp0 = 7;

struct PeArgs {
  int frames, g0, g1, g2;   // token grid per frame (8, 16, 16)
  int p0, p1;               // patch extents along D0, D1 (4, 16); the extent along D2 is 16
  int N;                    // hidden size (768)
  int P, Sp;                // tokens per frame, padded rows per frame of the output buffer
  const float* bias;        // [N]
  const __nv_bfloat16* pos; // [P, N]
  __nv_bfloat16* out;       // [frames, Sp, N]; token t goes to row 1 + t
  int dbg;                  // timing experiments only (U2_PE_DBG): 1 no proxy fence, 2 no staging reads, 4 no A-tile stores
};

// 5-D tiled load (fp32 slab of the volume)
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(c4)
      : "memory");
}

__global__ void __launch_bounds__(kPeThreads, 1)
patch_embed_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_vol, const __grid_constant__ CUtensorMap tmap_w,
                           const PeArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + kPeA * kPeABytes;
  uint8_t* sStg = sB + kPeB * kPeBBytes;
  uint8_t* sEpi = sStg + kPeStg * kPeStgBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + kPeEpiBytes);
  uint64_t* stg_full = bars;                 // [kPeStg]
  uint64_t* stg_empty = stg_full + kPeStg;   // [kPeStg]
  uint64_t* a_full = stg_empty + kPeStg;     // [kPeA]
  uint64_t* a_empty = a_full + kPeA;         // [kPeA]
  uint64_t* b_full = a_empty + kPeA;         // [kPeB]
  uint64_t* b_empty = b_full + kPeB;         // [kPeB]
  uint64_t* t_full = b_empty + kPeB;         // [2]
  uint64_t* t_empty = t_full + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w_tiles = p.g1 / 8;                                // M tiles per (frame, h)
  const int num_m = p.frames * p.g0 * w_tiles;
  const int num_n = (p.N + kPeN - 1) / kPeN;
  const int num_tiles = num_m * num_n;
  const int kb_per_p0 = p.p1 / 4;
  const int num_kb = p.p0 * kb_per_p0;                          // 16

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_vol);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp_idx == 1 && lane == 0) {
    // consumer barriers count WARPS, not threads: 128 mbarrier.arrive on one barrier serialise (~10 ns each) - with per-thread
    // arrivals on a_full and stg_empty every k-block cost 2.5 us whatever the converter did (first version: 419 us per pass)
    for (int s = 0; s < kPeStg; ++s) { mbar_init(&stg_full[s], 1); mbar_init(&stg_empty[s], 4); }
    for (int s = 0; s < kPeA; ++s) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kPeB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 8); }
    fence_barrier_init();
  }
  if (warp_idx == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ===================== TMA producer 1: fp32 volume slabs =====================
    // (separate from the weight-tile producer: with one thread feeding both rings the slab prefetch depth was tied to the
    //  2-stage weight ring and every other k-block paid a full HBM latency)
    if (lane == 0) {
      int ss = 0;
      uint32_t sph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n;
        const int wt = m_blk % w_tiles;
        const int fh = m_blk / w_tiles;
        const int h = fh % p.g0, f = fh / p.g0;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int q0 = kb / kb_per_p0, q1 = (kb - q0 * kb_per_p0) * 4;
          mbar_wait(&stg_empty[ss], sph ^ 1);
          mbar_arrive_expect_tx(&stg_full[ss], kPeStgBytes);
          tma_load_5d(sStg + ss * kPeStgBytes, &tmap_vol, &stg_full[ss], 0, q1, wt * 8, h * p.p0 + q0, f);
          if (++ss == kPeStg) { ss = 0; sph ^= 1; }
        }
      }
    }
  } else if (warp_idx == 2) {
    // ===================== TMA producer 2: weight tiles (this warp allocated TMEM above) =====================
    if (lane == 0) {
      int bs = 0;
      uint32_t bph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_blk = tile % num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&b_empty[bs], bph ^ 1);
          mbar_arrive_expect_tx(&b_full[bs], kPeBBytes);
          tma_load_4d(sB + bs * kPeBBytes, &tmap_w, &b_full[bs], kb * kPeK, n_blk * kPeN, 0, 0);
          if (++bs == kPeB) { bs = 0; bph ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kPeM, kPeN);
      int as = 0, bs = 0, acc = 0;
      uint32_t aph = 0, bph = 0, acc_ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&t_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kPeN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&a_full[as], aph);
          mbar_wait(&b_full[bs], bph);
          tc_fence_after();
          const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(sA + as * kPeABytes));
          const uint64_t b_desc = umma_desc_kmajor_sw128(smem_u32(sB + bs * kPeBBytes));
#pragma unroll
          for (int k = 0; k < kPeK / 16; ++k) umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&a_empty[as]);
          umma_commit(&b_empty[bs]);
          if (++as == kPeA) { as = 0; aph ^= 1; }
          if (++bs == kPeB) { bs = 0; bph ^= 1; }
        }
        umma_commit(&t_full[acc]);
        if (++acc == 2) { acc = 0; acc_ph ^= 1; }
      }
    }
  } else if (warp_idx >= kPeConvWarp0 && warp_idx < kPeEpiWarp0) {
    // ===================== converter: fp32 [w][p2][256] slab -> bf16 K-major SW128 A tile =====================
    // A staging row (w, p2) holds 16 tokens (d) x 16 fp32 (p3) = 1 KB: a warp reads it with two fully coalesced 16-byte
    // loads per lane (lane l -> token d = l / 4 (+ 8), p3 quarter l % 4), converts and writes 8 bytes into row
    // r = 16 w + d of the A tile at chunk (2 p2 + quarter / 2) ^ (r & 7): 32 lanes x 8 B cover every bank exactly twice.
    // All register indices are static (the first version routed pieces with data-dependent selects, which the compiler
    // turned into divergent branch trees: 819 instructions per warp and k-block, 2.5 us per k-block).
    const int cw = warp_idx - kPeConvWarp0;   // this warp converts w = 2 cw, 2 cw + 1
    const int dq = lane >> 2, quarter = lane & 3;
    int ss = 0, as = 0;
    uint32_t sph = 0, aph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&stg_full[ss], sph);
        mbar_wait(&a_empty[as], aph ^ 1);
        const uint32_t src0 = smem_u32(sStg) + ss * kPeStgBytes + lane * 16;
        const uint32_t dst0 = smem_u32(sA) + as * kPeABytes + (quarter & 1) * 8;
#pragma unroll
        for (int wi = 0; wi < 2; ++wi) {
          const int w = 2 * cw + wi;
#pragma unroll
          for (int q = 0; q < 4; ++q) {      // p2 offset inside the k-block
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const float4 v = (p.dbg & 2) ? make_float4(1.f, 2.f, 3.f, 4.f)
                                           : lds128_f32(src0 + (w * 4 + q) * 1024 + half * 512);
              const int r = w * 16 + half * 8 + dq;
              const int chunk = 2 * q + (quarter >> 1);
              __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
              if (!(p.dbg & 4))
                sts64(dst0 + r * 128 + ((chunk ^ (r & 7)) << 4), *reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
            }
          }
        }
        if (!(p.dbg & 1)) fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&a_full[as]);
          mbar_arrive(&stg_empty[ss]);
        }
        if (++ss == kPeStg) { ss = 0; sph ^= 1; }
        if (++as == kPeA) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp_idx >= kPeEpiWarp0) {
    // ===================== epilogue =====================
    const int q = warp_idx & 3;                       // TMEM lane quarter this warp may read
    const int half = (warp_idx - kPeEpiWarp0) >> 2;   // which 128 of the tile's 256 columns
    const uint32_t st = smem_u32(sEpi) + (warp_idx - kPeEpiWarp0) * 2048;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
      const long long row = (long long)m_blk * kPeM + q * 32 + lane;   // global token index (frame * P + t)
      const int t = (int)(row % p.P);
      const uint32_t taddr = tmem_base + acc * kPeN + (static_cast<uint32_t>(q * 32) << 16);
      const __nv_bfloat16* prow = p.pos + (long long)t * p.N;
      mbar_wait(&t_full[acc], acc_ph);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = half * (kPeN / 2); c0 < (half + 1) * (kPeN / 2); c0 += 32) {
        const int col0 = n_blk * kPeN + c0;
        if (col0 >= p.N) break;
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c0, v);
        uint4 rr[4];                                   // position-table segment: in flight while tcgen05.ld completes
#pragma unroll
        for (int j = 0; j < 4; ++j) rr[j] = *reinterpret_cast<const uint4*>(prow + col0 + 8 * j);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
          f[j] = __uint_as_float(v[j]) + b4.x;
          f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
          f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
          f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rr[j]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 rf = __bfloat1622float2(r2[e]);
            f[8 * j + 2 * e] += rf.x;
            f[8 * j + 2 * e + 1] += rf.y;
          }
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 o;
          __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
          for (int e = 0; e < 4; ++e) o2[e] = __floats2bfloat162_rn(f[8 * c + 2 * e], f[8 * c + 2 * e + 1]);
          sts128(st + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4), o);
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rw = it * 8 + (lane >> 2), ch = lane & 3;
          const uint4 o = lds128(st + rw * 64 + ((ch ^ ((rw >> 1) & 3)) << 4));
          const long long grow = (long long)m_blk * kPeM + q * 32 + rw;
          const long long orow = (grow / p.P) * p.Sp + 1 + grow % p.P;
          *reinterpret_cast<uint4*>(p.out + orow * p.N + col0 + ch * 8) = o;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
      if (++acc == 2) { acc = 0; acc_ph ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace u2

extern "C" U2_API int u2_patch_embed_f32_bf16(const float* vol, const void* W, const float* bias, const void* pos, void* out,
                                              int64_t frames, int32_t d0, int32_t d1, int32_t d2, int32_t p0, int32_t p1,
                                              int32_t p2, int32_t N, int64_t out_frame_rows, void* stream) {
  using namespace u2;
  if (!vol || !W || !bias || !pos || !out) return set_error(U2_ERR_ARG, "patch_embed: null pointer");
  if (p0 <= 0 || p1 <= 0 || p2 <= 0 || d0 % p0 || d1 % p1 || d2 % p2) return set_error(U2_ERR_ARG, "patch_embed: image not divisible by the patch");
  const int g0 = d0 / p0, g1 = d1 / p1, g2 = d2 / p2;
  // the fused tiling: 16-voxel patch rows along D2, 16 tokens along D2 per (h, w) and groups of 8 w per M tile
  if (p2 != 16 || g2 != 16 || (p1 & 3) || (g1 & 7) || d2 > 256 || (N & 31))
    return set_error(U2_ERR_UNSUPPORTED, "patch_embed: fused kernel covers patch (*, 4k, 16) on a (*, 8m, 16) token grid with D2 <= 256 "
                                         "(got patch %d x %d x %d, grid %d x %d x %d); use u2_patchify_f32_bf16 + u2_gemm_bf16", p0, p1, p2, g0, g1, g2);
  const long long P = (long long)g0 * g1 * g2;
  if (out_frame_rows < P + 1) return set_error(U2_ERR_ARG, "patch_embed: out_frame_rows must be >= tokens + 1 (cls row)");
  if ((reinterpret_cast<uintptr_t>(vol) & 15) || (reinterpret_cast<uintptr_t>(W) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(pos) & 15) || (reinterpret_cast<uintptr_t>(bias) & 15))
    return set_error(U2_ERR_ARG, "patch_embed: pointers must be 16-byte aligned");
  if (frames <= 0) return U2_OK;
  const int K = p0 * p1 * p2;
  CUtensorMap tv, tw;
  {
    // volume [frames][D0][D1 = (w, p2)][D2] fp32 as a 5-D map {D2, p2, w, D0, frame}; box {D2, 4, 8, 1, 1}
    const int64_t dims[5] = {d2, p1, g1, d0, frames};
    const int64_t strides[4] = {(int64_t)d2 * 4, (int64_t)p1 * d2 * 4, (int64_t)d1 * d2 * 4, (int64_t)d0 * d1 * d2 * 4};
    const int box[5] = {d2, 4, 8, 1, 1};
    int rc = make_tmap_f32_nd(&tv, vol, 5, dims, strides, box);
    if (rc) return rc;
  }
  int rc = make_tmap_bf16_4d(&tw, W, K, N, 1, 1, K, 0, 0, kPeK, kPeN);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(patch_embed_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeSmem);
    if (e != cudaSuccess) return set_error(U2_ERR_CUDA, "patch_embed: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  PeArgs a;
  a.frames = (int)frames; a.g0 = g0; a.g1 = g1; a.g2 = g2; a.p0 = p0; a.p1 = p1; a.N = N;
  a.P = (int)P; a.Sp = (int)out_frame_rows;
  a.dbg = getenv("U2_PE_DBG") ? atoi(getenv("U2_PE_DBG")) : 0;
  a.bias = bias; a.pos = reinterpret_cast<const __nv_bfloat16*>(pos); a.out = reinterpret_cast<__nv_bfloat16*>(out);
  const long long tiles = (long long)frames * g0 * (g1 / 8) * ((N + kPeN - 1) / kPeN);
  const int sms = num_sms();
  const int grid = (int)(tiles < sms ? tiles : sms);
  patch_embed_tcgen05_kernel<<<grid, kPeThreads, kPeSmem, reinterpret_cast<cudaStream_t>(stream)>>>(tv, tw, a);
  U2_CHECK_LAUNCH("patch_embed");
  return U2_OK;
}
