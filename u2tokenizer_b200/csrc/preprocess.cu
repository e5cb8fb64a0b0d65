// Volume preprocessing in front of the hot path: the reference's `u2Transform.adaptive_resize`
// (src/utils/u2Transform.py:62-122; MONAI ScaleIntensityRangePercentiles -> CropForeground -> anti-aliased trilinear
// resize -> zero pad -> 32-slice chunks) as a short sequence of HBM-bound kernels on a volume that is already on the
// device. Everything data dependent (percentiles, foreground box, output extents, smoothing widths) stays on the device:
// no host synchronisation between the steps.
//
//   1. exact percentiles: 3-pass radix select (11 + 11 + 10 key bits) of the 4 order statistics np.percentile's linear
//      interpolation needs (one histogram pass over the volume per radix digit, shared by the 4 ranks)
//   2. intensity scaling + clip of every voxel and, in the same pass, the foreground box (min / max coordinates of the
//      voxels above the lower percentile)
//   3. plan: crop extents -> output extents, per-axis anti-aliasing sigma (MONAI resize), one thread
//   4. separable Gaussian (zero padded at the crop faces), H then W then D like MONAI's separable_filtering
//   5. trilinear resample (align_corners) + zero pad, written in the (D, H, W) order the model consumes
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "host_util.h"
#include "u2b200.h"

namespace u2 {

constexpr int kPpMaxTail = 64;

struct PpState {                 // device-side working state (lives in the workspace)
  unsigned int prefix[4];        // radix select: matched high key bits per rank
  unsigned int rank[4];          // remaining rank inside the matched prefix
  float order[4];                // the order statistics v[lo0], v[lo0 + 1], v[lo1], v[lo1 + 1]
};

__device__ __forceinline__ unsigned int pp_key(float x) {
  const unsigned int u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float pp_unkey(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// warp-aggregated shared-memory histogram increment (CT volumes hold huge runs of identical values)
__device__ __forceinline__ void pp_hist_add(unsigned int* hist, unsigned int bin, bool pred) {
  const unsigned int act = __ballot_sync(0xffffffffu, pred);
  if (act == 0) return;  // warp-uniform
  const int leader = __ffs(act) - 1;
  const unsigned int b0 = __shfl_sync(0xffffffffu, bin, leader);
  if (__all_sync(0xffffffffu, !pred || bin == b0)) {  // the common case in air / saturated regions: one bin per warp
    if ((int)(threadIdx.x & 31) == leader) atomicAdd(hist + b0, (unsigned int)__popc(act));
    return;
  }
  if (!pred) return;
  const unsigned int peers = __match_any_sync(act, bin);
  if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(hist + bin, (unsigned int)__popc(peers));
}

// kPass 0: digit = key[31:21], every element, one histogram.   kPass 1: digit = key[20:10] of the elements whose
// key[31:21] matches rank r's prefix, 4 histograms.   kPass 2: digit = key[9:0], prefix key[31:10].
template <int kPass>
__global__ void __launch_bounds__(256) pp_hist_kernel(const float* __restrict__ vol, long long n,
                                                      const PpState* __restrict__ st, unsigned int* __restrict__ hist) {
  constexpr int kBins = (kPass == 2) ? 1024 : 2048;
  constexpr int kHists = (kPass == 0) ? 1 : 4;
  __shared__ unsigned int sh[kHists * kBins];
  for (int i = threadIdx.x; i < kHists * kBins; i += 256) sh[i] = 0;
  constexpr int kPreShift = (kPass == 1) ? 21 : 10;
  unsigned int pre[4] = {0, 0, 0, 0};
  bool dup[4] = {false, false, false, false};  // same prefix as a lower rank: that rank's histogram serves both
  if (kPass > 0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) pre[r] = st->prefix[r] >> kPreShift;
#pragma unroll
    for (int r = 1; r < 4; ++r)
      for (int q = 0; q < r; ++q) dup[r] = dup[r] || pre[q] == pre[r];
  }
  __syncthreads();
  const long long stride = (long long)gridDim.x * 256;
  const long long n_round = (n + 31) / 32 * 32;  // whole warps stay in the loop: the ballots need every lane
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n_round; i += stride) {
    const bool in = i < n;
    const unsigned int k = in ? pp_key(__ldg(vol + i)) : 0u;
    if (kPass == 0) {
      pp_hist_add(sh, k >> 21, in);
    } else if (kPass == 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (!dup[r]) pp_hist_add(sh + r * kBins, (k >> 10) & 0x7ffu, in && (k >> 21) == pre[r]);
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (!dup[r]) pp_hist_add(sh + r * kBins, k & 0x3ffu, in && (k >> 10) == pre[r]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kHists * kBins; i += 256)
    if (sh[i]) atomicAdd(hist + i, sh[i]);
}

// One block: per rank, find the digit whose cumulative count crosses the remaining rank.
template <int kPass>
__global__ void __launch_bounds__(128) pp_pick_kernel(PpState* st, const unsigned int* __restrict__ hist) {
  constexpr int kBins = (kPass == 2) ? 1024 : 2048;
  constexpr int kShift = (kPass == 0) ? 21 : (kPass == 1 ? 10 : 0);
  const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;  // warp r handles rank r
  constexpr int kPreShift = (kPass == 1) ? 21 : 10;
  int src = r;  // histogram slot: the lowest rank with the same prefix (pp_hist_kernel filled only that one)
  unsigned int my_prefix = 0;
  if (kPass > 0) {
    my_prefix = st->prefix[r];
    for (int q = r - 1; q >= 0; --q)
      if ((st->prefix[q] >> kPreShift) == (my_prefix >> kPreShift)) src = q;
  }
  __syncthreads();  // every warp has read the prefixes before any of them is updated
  const unsigned int* h = hist + (kPass == 0 ? 0 : src * kBins);
  unsigned int want = st->rank[r];
  // lane l owns the contiguous segment [l * kSeg, (l + 1) * kSeg): independent loads, then one warp scan of the segment
  // totals, then the owning lane walks its segment
  constexpr int kSeg = kBins / 32;
  unsigned int seg = 0;
#pragma unroll 8
  for (int i = 0; i < kSeg; ++i) seg += h[lane * kSeg + i];
  unsigned int inc = seg;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  const unsigned int m = __ballot_sync(0xffffffffu, want < inc);
  const int owner = m ? __ffs(m) - 1 : 31;
  int found = -1;
  if (lane == owner) {
    unsigned int before = inc - seg;
    for (int i = 0; i < kSeg; ++i) {
      const unsigned int c = h[lane * kSeg + i];
      if (want < before + c) {
        found = lane * kSeg + i;
        break;
      }
      before += c;
    }
    if (found < 0) {  // cannot happen for ranks < n
      found = lane * kSeg + kSeg - 1;
      before = 0;
    }
    want -= before;
  }
  found = __shfl_sync(0xffffffffu, found, owner);
  want = __shfl_sync(0xffffffffu, want, owner);
  if (lane == 0) {
    if (found < 0) found = kBins - 1;  // cannot happen for ranks < n
    const unsigned int p = my_prefix | ((unsigned int)found << kShift);
    st->prefix[r] = p;
    st->rank[r] = want;
    if (kPass == 2) st->order[r] = pp_unkey(p);
  }
}

__global__ void pp_init_kernel(PpState* st, u2_preprocess_info* info, unsigned int r0, unsigned int r1, unsigned int r2,
                               unsigned int r3, int D, int H, int W) {
  st->rank[0] = r0; st->rank[1] = r1; st->rank[2] = r2; st->rank[3] = r3;
#pragma unroll
  for (int r = 0; r < 4; ++r) st->prefix[r] = 0;
  info->lo[0] = D; info->lo[1] = H; info->lo[2] = W;
  info->hi[0] = 0; info->hi[1] = 0; info->hi[2] = 0;
  info->status = 0;
}

// numpy's _lerp (np.percentile, method "linear"), float64
__device__ __forceinline__ double pp_lerp(double a, double b, double t) {
  const double diff = b - a;
  double v = a + diff * t;
  if (t >= 0.5) v = b - diff * (1.0 - t);
  return v;
}

__global__ void pp_percentile_kernel(const PpState* st, u2_preprocess_info* info, double g_lo, double g_hi) {
  info->a_min = pp_lerp((double)st->order[0], (double)st->order[1], g_lo);
  info->a_max = pp_lerp((double)st->order[2], (double)st->order[3], g_hi);
}

// Intensity scaling + clip of every voxel (MONAI ScaleIntensityRange: float64 arithmetic, float32 result; the division by
// the range is a multiplication with its reciprocal, at most one float64 ulp away before the rounding to float32) and, in
// the same pass, the foreground box: voxels whose scaled intensity is positive  <=>  x > a_min.
// One block per (d, h) row at a time.
__global__ void __launch_bounds__(256) pp_scale_bbox_kernel(const float* __restrict__ vol, float* __restrict__ scaled, int D,
                                                            int H, int W, u2_preprocess_info* info) {
  const double a_min = info->a_min;
  const double range = info->a_max - info->a_min;
  const bool flat = range == 0.0;
  const double inv_range = flat ? 0.0 : 1.0 / range;
  int lo[3] = {D, H, W}, hi[3] = {0, 0, 0};
  for (unsigned int row = blockIdx.x; row < (unsigned int)(D * H); row += gridDim.x) {
    const int d = (int)(row / (unsigned int)H), h = (int)(row - (unsigned int)d * H);
    const float* p = vol + (size_t)row * W;
    float* q = scaled + (size_t)row * W;
    for (int w = threadIdx.x; w < W; w += 256) {
      const double v = (double)__ldg(p + w);
      q[w] = (float)(flat ? v - a_min : fmin(fmax((v - a_min) * inv_range, 0.0), 1.0));
      if (v > a_min) {
        lo[0] = min(lo[0], d); hi[0] = max(hi[0], d + 1);
        lo[1] = min(lo[1], h); hi[1] = max(hi[1], h + 1);
        lo[2] = min(lo[2], w); hi[2] = max(hi[2], w + 1);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if ((threadIdx.x & 31) == 0 && hi[a] > 0) {
      atomicMin(&info->lo[a], lo[a]);
      atomicMax(&info->hi[a], hi[a]);
    }
  }
}

// Output extents and smoothing widths (u2Transform.py:72-80,96-97 and MONAI resize's anti-aliasing rule).
__global__ void pp_plan_kernel(u2_preprocess_info* info, int target, int pad_depth) {
  const int Dc = info->hi[0] - info->lo[0], Hc = info->hi[1] - info->lo[1], Wc = info->hi[2] - info->lo[2];
  if (Dc <= 0 || Hc <= 0 || Wc <= 0) {
    info->status = U2_PP_EMPTY_FOREGROUND;
    info->out[0] = info->out[1] = info->out[2] = 0;
    return;
  }
  const double ratio = fmin((double)target / (double)Hc, (double)target / (double)Wc);
  const int oh = (int)((double)Hc * ratio), ow = (int)((double)Wc * ratio);
  const int od = (pad_depth >= Dc) ? Dc : pad_depth;
  info->out[0] = od; info->out[1] = oh; info->out[2] = ow;
  if (oh <= 0 || ow <= 0) {
    info->status = U2_PP_DEGENERATE_SHAPE;
    return;
  }
  const int in_sz[3] = {Dc, Hc, Wc}, out_sz[3] = {od, oh, ow};
  const bool shrink = od < Dc || oh < Hc || ow < Wc;
  for (int a = 0; a < 3; ++a) {
    const float f = (float)in_sz[a] / (float)out_sz[a];
    const float s = shrink ? fmaxf(0.f, (f - 1.f) / 2.f) : 0.f;
    const int tail = (int)(fmax((double)s * 4.0, 0.5) + 0.5);
    info->sigma[a] = s;
    info->tail[a] = tail;
    if (tail > kPpMaxTail) info->status = U2_PP_DEGENERATE_SHAPE;
  }
  if (info->a_max - info->a_min == 0.0 && info->status == 0) info->status = U2_PP_FLAT_INTENSITY;
}

// One separable pass along kAxis (0 = D, 1 = H, 2 = W in storage order) over the crop box; zero outside the box.
template <int kAxis>
__global__ void __launch_bounds__(256) pp_smooth_kernel(const float* __restrict__ src, float* __restrict__ dst, int D, int H,
                                                        int W, const u2_preprocess_info* __restrict__ info) {
  __shared__ float taps[2 * kPpMaxTail + 1];
  if (info->status == U2_PP_EMPTY_FOREGROUND || info->status == U2_PP_DEGENERATE_SHAPE) return;
  const int tail = info->tail[kAxis];
  if (threadIdx.x <= 2 * tail) {
    // MONAI gaussian_1d(sigma, truncated=4, approx="erf", normalize=False); sigma == 0 gives the identity (0, 1, 0)
    const float s = info->sigma[kAxis];
    const float t = 0.70710678f / fabsf(s);
    const float x = (float)((int)threadIdx.x - tail);
    taps[threadIdx.x] = fmaxf(0.f, 0.5f * (erff(t * (x + 0.5f)) - erff(t * (x - 0.5f))));
  }
  __syncthreads();
  const int lo0 = info->lo[0], lo1 = info->lo[1], lo2 = info->lo[2];
  const int Dc = info->hi[0] - lo0, Hc = info->hi[1] - lo1, Wc = info->hi[2] - lo2;
  const int len = kAxis == 0 ? Dc : (kAxis == 1 ? Hc : Wc);
  const long long step = kAxis == 0 ? (long long)H * W : (kAxis == 1 ? W : 1);
  for (unsigned int row = blockIdx.x; row < (unsigned int)(Dc * Hc); row += gridDim.x) {  // one (d, h) row of the crop
    const int d = (int)(row / (unsigned int)Hc), h = (int)(row - (unsigned int)d * Hc);
    const long long base = ((long long)(lo0 + d) * H + (lo1 + h)) * W + lo2;
    for (int w = threadIdx.x; w < Wc; w += 256) {
      const long long at = base + w;
      const int pos = kAxis == 0 ? d : (kAxis == 1 ? h : w);
      const int k0 = max(-tail, -pos), k1 = min(tail, len - 1 - pos);
      float acc = 0.f;
      for (int k = k0; k <= k1; ++k) acc += taps[k + tail] * __ldg(src + at + k * step);
      dst[at] = acc;
    }
  }
}

// out[z][y][x] = trilinear(align_corners=True) sample of the smoothed crop, zero outside the resized extent.
__global__ void __launch_bounds__(256) pp_resize_kernel(const float* __restrict__ src, float* __restrict__ out, int H, int W,
                                                        int target, int pad_depth,
                                                        const u2_preprocess_info* __restrict__ info) {
  const bool ok = info->status == 0 || info->status == U2_PP_FLAT_INTENSITY;
  const int od = info->out[0], oh = info->out[1], ow = info->out[2];
  const int lo0 = info->lo[0], lo1 = info->lo[1], lo2 = info->lo[2];
  const int Dc = info->hi[0] - lo0, Hc = info->hi[1] - lo1, Wc = info->hi[2] - lo2;
  // torch area_pixel_compute_scale<float>(align_corners=true)
  const float sd = od > 1 ? (float)(Dc - 1) / (float)(od - 1) : 0.f;
  const float sh = oh > 1 ? (float)(Hc - 1) / (float)(oh - 1) : 0.f;
  const float sw = ow > 1 ? (float)(Wc - 1) / (float)(ow - 1) : 0.f;
  for (unsigned int row = blockIdx.x; row < (unsigned int)(pad_depth * target); row += gridDim.x) {  // one (z, y) row
    const int z = (int)(row / (unsigned int)target), y = (int)(row - (unsigned int)z * target);
    float* orow = out + (size_t)row * target;
    const bool row_in = ok && z < od && y < oh;
    const float fh = sh * y, fd = sd * z;
    const int h0 = (int)fh, d0 = (int)fd;
    const int h1 = h0 + (h0 < Hc - 1), d1 = d0 + (d0 < Dc - 1);
    const float lh1 = fh - h0, ld1 = fd - d0;
    const float lh0 = 1.f - lh1, ld0 = 1.f - ld1;
    const float* r00 = src + ((long long)(lo0 + d0) * H + (lo1 + h0)) * W + lo2;
    const float* r10 = src + ((long long)(lo0 + d1) * H + (lo1 + h0)) * W + lo2;
    const float* r01 = src + ((long long)(lo0 + d0) * H + (lo1 + h1)) * W + lo2;
    const float* r11 = src + ((long long)(lo0 + d1) * H + (lo1 + h1)) * W + lo2;
    for (int x = threadIdx.x; x < target; x += 256) {
      float v = 0.f;
      if (row_in && x < ow) {
        const float fw = sw * x;
        const int w0 = (int)fw;
        const int w1 = w0 + (w0 < Wc - 1);
        const float lw1 = fw - w0, lw0 = 1.f - lw1;
        // nesting of torch's upsample_trilinear3d on the reference's (H, W, D) tensor: t = H, h = W, w = D
        v = lh0 * (lw0 * (ld0 * __ldg(r00 + w0) + ld1 * __ldg(r10 + w0)) + lw1 * (ld0 * __ldg(r00 + w1) + ld1 * __ldg(r10 + w1))) +
            lh1 * (lw0 * (ld0 * __ldg(r01 + w0) + ld1 * __ldg(r11 + w0)) + lw1 * (ld0 * __ldg(r01 + w1) + ld1 * __ldg(r11 + w1)));
      }
      orow[x] = v;
    }
  }
}

static inline size_t pp_align(size_t x) { return (x + 255) / 256 * 256; }
constexpr size_t kPpHistBytes = 4 * 2048 * sizeof(unsigned int);

}  // namespace u2

extern "C" U2_API int64_t u2_preprocess_ws_bytes(int32_t D, int32_t H, int32_t W) {
  using namespace u2;
  if (D <= 0 || H <= 0 || W <= 0) return 0;
  const size_t n = (size_t)D * H * W;
  return (int64_t)(pp_align(sizeof(PpState)) + 3 * pp_align(kPpHistBytes) + 2 * pp_align(n * sizeof(float)));
}

extern "C" U2_API int u2_preprocess_volume_f32(const float* vol, float* out, u2_preprocess_info* info,
                                               const u2_preprocess_desc* d, void* stream) {
  using namespace u2;
  if (!vol || !out || !info || !d || !d->ws) return set_error(U2_ERR_ARG, "preprocess: null pointer");
  if (d->D <= 0 || d->H <= 0 || d->W <= 0 || d->target <= 0 || d->pad_depth <= 0)
    return set_error(U2_ERR_ARG, "preprocess: extents must be > 0");
  if (!(d->lower_pct >= 0.0 && d->lower_pct <= d->upper_pct && d->upper_pct <= 100.0))
    return set_error(U2_ERR_ARG, "preprocess: need 0 <= lower_pct <= upper_pct <= 100");
  if (d->ws_bytes < u2_preprocess_ws_bytes(d->D, d->H, d->W))
    return set_error(U2_ERR_ARG, "preprocess: workspace too small (%lld < %lld bytes)", (long long)d->ws_bytes,
                     (long long)u2_preprocess_ws_bytes(d->D, d->H, d->W));
  if (reinterpret_cast<uintptr_t>(d->ws) & 255) return set_error(U2_ERR_ARG, "preprocess: workspace must be 256-byte aligned");
  const long long n = (long long)d->D * d->H * d->W;
  if (n >= (1LL << 32)) return set_error(U2_ERR_UNSUPPORTED, "preprocess: volumes of 2^32 voxels or more");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* w = reinterpret_cast<char*>(d->ws);
  PpState* state = reinterpret_cast<PpState*>(w);
  w += pp_align(sizeof(PpState));
  unsigned int* hist[3];
  for (int i = 0; i < 3; ++i) {
    hist[i] = reinterpret_cast<unsigned int*>(w);
    w += pp_align(kPpHistBytes);
  }
  float* buf_a = reinterpret_cast<float*>(w);
  float* buf_b = reinterpret_cast<float*>(w + pp_align((size_t)n * sizeof(float)));

  // np.percentile(method="linear"): virtual index (n - 1) * q, neighbours floor / floor + 1, weight = fraction
  double vi[2], gam[2];
  unsigned int ranks[4];
  const double qs[2] = {d->lower_pct / 100.0, d->upper_pct / 100.0};
  for (int i = 0; i < 2; ++i) {
    vi[i] = (double)(n - 1) * qs[i];
    double fl = floor(vi[i]);
    if (fl > (double)(n - 1)) fl = (double)(n - 1);
    gam[i] = vi[i] - fl;
    ranks[2 * i] = (unsigned int)fl;
    ranks[2 * i + 1] = (unsigned int)((fl + 1 <= (double)(n - 1)) ? fl + 1 : fl);
  }
  cudaMemsetAsync(hist[0], 0, 3 * pp_align(kPpHistBytes), st);
  pp_init_kernel<<<1, 1, 0, st>>>(state, info, ranks[0], ranks[1], ranks[2], ranks[3], d->D, d->H, d->W);
  const int blocks = num_sms() * 8;
  pp_hist_kernel<0><<<blocks, 256, 0, st>>>(vol, n, state, hist[0]);
  pp_pick_kernel<0><<<1, 128, 0, st>>>(state, hist[0]);
  pp_hist_kernel<1><<<blocks, 256, 0, st>>>(vol, n, state, hist[1]);
  pp_pick_kernel<1><<<1, 128, 0, st>>>(state, hist[1]);
  pp_hist_kernel<2><<<blocks, 256, 0, st>>>(vol, n, state, hist[2]);
  pp_pick_kernel<2><<<1, 128, 0, st>>>(state, hist[2]);
  pp_percentile_kernel<<<1, 1, 0, st>>>(state, info, gam[0], gam[1]);
  pp_scale_bbox_kernel<<<blocks, 256, 0, st>>>(vol, buf_b, d->D, d->H, d->W, info);
  pp_plan_kernel<<<1, 1, 0, st>>>(info, d->target, d->pad_depth);
  pp_smooth_kernel<1><<<blocks, 256, 0, st>>>(buf_b, buf_a, d->D, d->H, d->W, info);
  pp_smooth_kernel<2><<<blocks, 256, 0, st>>>(buf_a, buf_b, d->D, d->H, d->W, info);
  pp_smooth_kernel<0><<<blocks, 256, 0, st>>>(buf_b, buf_a, d->D, d->H, d->W, info);
  pp_resize_kernel<<<blocks, 256, 0, st>>>(buf_a, out, d->H, d->W, d->target, d->pad_depth, info);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(U2_ERR_CUDA, "preprocess launch: %s", cudaGetErrorString(e));
  return U2_OK;
}
