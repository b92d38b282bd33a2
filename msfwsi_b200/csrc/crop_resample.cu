// A2: crop the low-magnification feature map to each high-magnification tile's footprint and
// bilinearly resample it.  Integer boxes whose size equals (oh, ow) reproduce the reference's only
// feature-map crop, src/models/hooknet.py:29-32 (`x[:, :, 12:20, 12:20]`), bit-exactly; everything else
// is the north_star extension (F.interpolate bilinear / align_corners=False on the crop).
//
// HBM-bound: output rows are written with 128-bit stores (VEC consecutive ox per thread); the four
// taps of neighbouring outputs overlap, so input traffic is the footprint once (L1/L2 hits after).
// Algorithmic bytes: B*C*(union of footprints)*e read + B*K*C*oh*ow*e written + 16 B per box.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace msf {
namespace {

struct Geo {
  int64_t B;
  int C, H, W, K, oh, ow;
};

template <int DT>
__device__ __forceinline__ float ld1(const void* base, int64_t i) {
  if constexpr (DT == MSF_F32) return __ldg(static_cast<const float*>(base) + i);
  else if constexpr (DT == MSF_BF16) return __bfloat162float(__ldg(static_cast<const __nv_bfloat16*>(base) + i));
  else return __half2float(__ldg(static_cast<const __half*>(base) + i));
}
template <int DT>
__device__ __forceinline__ void st1(void* base, int64_t i, float v) {
  if constexpr (DT == MSF_F32) static_cast<float*>(base)[i] = v;
  else if constexpr (DT == MSF_BF16) static_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v);
  else static_cast<__half*>(base)[i] = __float2half_rn(v);
}

// One axis of F.interpolate(bilinear, align_corners=False) on a crop [lo, lo+len): taps i0,i1 (absolute,
// clamped to the map) and the weight of i1.
__device__ __forceinline__ void axis_taps(float lo, float len, int o, int osz, int limit, int& i0, int& i1, float& w) {
  const float s = fmaxf((o + 0.5f) * (len / osz) - 0.5f, 0.f);
  const int imax = max(static_cast<int>(ceilf(len)) - 1, 0);
  const int f = min(static_cast<int>(floorf(s)), imax);
  const int f1 = min(f + 1, imax);
  w = f < imax ? s - floorf(s) : 0.f;
  const int off = static_cast<int>(floorf(lo));
  i0 = min(max(off + f, 0), limit - 1);
  i1 = min(max(off + f1, 0), limit - 1);
}

__device__ __forceinline__ float lerp_exact(float a, float b, float w) {
  return w == 0.f ? a : fmaf(b, w, a * (1.f - w));  // w==0 keeps the integer case a bit-exact copy
}

template <int DT, int VEC>
__global__ void __launch_bounds__(256) crop_fwd_kernel(const void* __restrict__ feat, const float* __restrict__ boxes,
                                                       void* __restrict__ out, Geo g, int64_t total) {
  const int owv = g.ow / VEC;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t r = t;
    const int oxv = static_cast<int>(r % owv); r /= owv;
    const int oy = static_cast<int>(r % g.oh); r /= g.oh;
    const int c = static_cast<int>(r % g.C); r /= g.C;
    const int k = static_cast<int>(r % g.K);
    const int64_t b = r / g.K;
    const float4 bx = __ldg(reinterpret_cast<const float4*>(boxes) + b * g.K + k);  // y0,x0,y1,x1
    int y0, y1;
    float wy;
    axis_taps(bx.x, bx.z - bx.x, oy, g.oh, g.H, y0, y1, wy);
    const int64_t plane = (b * g.C + c) * static_cast<int64_t>(g.H) * g.W;
    const int64_t r0 = plane + static_cast<int64_t>(y0) * g.W, r1 = plane + static_cast<int64_t>(y1) * g.W;
    float res[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      int x0, x1;
      float wx;
      axis_taps(bx.y, bx.w - bx.y, oxv * VEC + v, g.ow, g.W, x0, x1, wx);
      const float top = lerp_exact(ld1<DT>(feat, r0 + x0), wx == 0.f ? 0.f : ld1<DT>(feat, r0 + x1), wx);
      float bot = 0.f;
      if (wy != 0.f) bot = lerp_exact(ld1<DT>(feat, r1 + x0), wx == 0.f ? 0.f : ld1<DT>(feat, r1 + x1), wx);
      res[v] = lerp_exact(top, bot, wy);
    }
    const int64_t o = (((b * g.K + k) * g.C + c) * g.oh + oy) * static_cast<int64_t>(g.ow) + static_cast<int64_t>(oxv) * VEC;
    if constexpr (VEC == 1) {
      st1<DT>(out, o, res[0]);
    } else {
      static_assert(VEC == Elem<DT>::VEC, "vector path moves one 16-byte chunk");
      stg_stream(static_cast<char*>(out) + o * (16 / VEC), Elem<DT>::pack(res));
    }
  }
}

// ---- fast path: one warp per (box, channel) plane, horizontal pass once per source row -----------------------
// CTA = one box (b,k) x 8 consecutive channels (one per warp).  A warp walks down its output plane:
//   horizontal  every source row is interpolated along x ONCE (lane owns ox = lane + 32k; neighbouring lanes read
//               the same or adjacent source pixels, so a warp load is 1-2 sectors) into a 4-row rolling buffer in
//               shared memory -- its cost is shared by the `zoom` output rows that reuse the row;
//   vertical    two output rows per iteration: each lane reads 2 x (two 128-bit shared loads) of the two source
//               rows, blends with (1-wy, wy), and stores 16 bytes (VEC outputs).
// Weights equal to zero short-circuit, so an integer box of the output size is a bit-exact copy.
constexpr int kFastThreads = 256, kMaxOxPerLane = 8;

template <int DT>
__global__ void __launch_bounds__(kFastThreads) crop_fwd_fast_kernel(const void* __restrict__ feat, const float* __restrict__ boxes,
                                                                     void* __restrict__ out, Geo g, int kRoll) {
  constexpr int VEC = Elem<DT>::VEC;
  extern __shared__ __align__(16) float sm[];
  int* y_i0 = reinterpret_cast<int*>(sm);
  int* y_i1 = y_i0 + g.oh;
  float* y_w = reinterpret_cast<float*>(y_i1 + g.oh);
  float* roll = sm + ((3 * g.oh + 3) & ~3);  // [8 warps][kRoll][ow]
  const int groups_c = (g.C + 7) / 8;
  const int64_t bk = blockIdx.x / groups_c;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = static_cast<int>(blockIdx.x % groups_c) * 8 + warp;
  const int64_t b = bk / g.K;
  const float4 bx = __ldg(reinterpret_cast<const float4*>(boxes) + bk);  // y0,x0,y1,x1
  for (int o = threadIdx.x; o < g.oh; o += kFastThreads) axis_taps(bx.x, bx.z - bx.x, o, g.oh, g.H, y_i0[o], y_i1[o], y_w[o]);
  // x taps of the output columns this lane interpolates in the horizontal pass
  const int nk = g.ow >> 5;  // ow is a multiple of 32 on this path
  int xa[kMaxOxPerLane], xb[kMaxOxPerLane];
  float xw[kMaxOxPerLane];
#pragma unroll
  for (int k = 0; k < kMaxOxPerLane; ++k) {
    xa[k] = xb[k] = 0;
    xw[k] = 0.f;
    if (k < nk) axis_taps(bx.y, bx.w - bx.y, lane + 32 * k, g.ow, g.W, xa[k], xb[k], xw[k]);
  }
  __syncthreads();
  const int owv = g.ow / VEC;                 // lanes per output row (power of two <= 32)
  const int rows_per_iter = 32 / owv;
  // the source rows one iteration touches must fit the rolling buffer (always true unless the box is strongly
  // down-sampled); otherwise this CTA takes the direct four-tap path
  int fits = 1;
  for (int o = threadIdx.x * rows_per_iter; o < g.oh; o += kFastThreads * rows_per_iter)
    if (y_i1[min(o + rows_per_iter, g.oh) - 1] - y_i0[o] + 1 > kRoll) fits = 0;
  fits = __syncthreads_and(fits);
  if (c >= g.C) return;
  const int64_t plane = (b * g.C + c) * static_cast<int64_t>(g.H) * g.W;
  const int64_t oplane = (bk * g.C + c) * static_cast<int64_t>(g.oh) * g.ow;
  // integer box of exactly the output size (the reference's case, hooknet.py:29-32): plain strided copy with the
  // widest loads the source alignment allows; every lane moves one 16-byte output chunk per step
  const bool is_copy = bx.x == floorf(bx.x) && bx.y == floorf(bx.y) && bx.z - bx.x == static_cast<float>(g.oh) &&
                       bx.w - bx.y == static_cast<float>(g.ow) && bx.x >= 0.f && bx.y >= 0.f &&
                       bx.z <= static_cast<float>(g.H) && bx.w <= static_cast<float>(g.W);
  if (is_copy) {
    constexpr int ES = 16 / VEC;  // element size in bytes
    const char* src = static_cast<const char*>(feat) + (plane + static_cast<int64_t>(bx.x) * g.W + static_cast<int64_t>(bx.y)) * ES;
    char* dst = static_cast<char*>(out) + oplane * ES;
    const int64_t src_pitch = static_cast<int64_t>(g.W) * ES;
    const int unit = static_cast<int>((reinterpret_cast<uintptr_t>(src) | static_cast<uintptr_t>(src_pitch) | 16u) &
                                      (~(reinterpret_cast<uintptr_t>(src) | static_cast<uintptr_t>(src_pitch) | 16u) + 1));  // lowest set bit
    const int chunks = g.oh * owv;
    for (int e = lane; e < chunks; e += 32) {
      const int oy = e / owv, xc = e - oy * owv;
      const char* sp = src + oy * src_pitch + xc * 16;
      uint4 v;
      if (unit >= 16) {
        v = ldg_stream(sp);
      } else if (unit == 8) {
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(sp)), b2 = __ldg(reinterpret_cast<const uint2*>(sp + 8));
        v = make_uint4(a.x, a.y, b2.x, b2.y);
      } else if (unit == 4) {
        v = make_uint4(__ldg(reinterpret_cast<const uint32_t*>(sp)), __ldg(reinterpret_cast<const uint32_t*>(sp + 4)),
                       __ldg(reinterpret_cast<const uint32_t*>(sp + 8)), __ldg(reinterpret_cast<const uint32_t*>(sp + 12)));
      } else {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          w[i] = static_cast<uint32_t>(__ldg(reinterpret_cast<const uint16_t*>(sp + 4 * i))) |
                 (static_cast<uint32_t>(__ldg(reinterpret_cast<const uint16_t*>(sp + 4 * i + 2))) << 16);
        v = make_uint4(w[0], w[1], w[2], w[3]);
      }
      stg_stream(dst + static_cast<int64_t>(e) * 16, v);
    }
    return;
  }
  if (!fits) {
    for (int e = lane; e < g.oh * g.ow; e += 32) {
      const int oy = e / g.ow, ox = e - oy * g.ow;
      int x0, x1;
      float wx;
      axis_taps(bx.y, bx.w - bx.y, ox, g.ow, g.W, x0, x1, wx);
      const int64_t r0 = plane + static_cast<int64_t>(y_i0[oy]) * g.W, r1 = plane + static_cast<int64_t>(y_i1[oy]) * g.W;
      const float top = lerp_exact(ld1<DT>(feat, r0 + x0), wx == 0.f ? 0.f : ld1<DT>(feat, r0 + x1), wx);
      float bot = 0.f;
      if (y_w[oy] != 0.f) bot = lerp_exact(ld1<DT>(feat, r1 + x0), wx == 0.f ? 0.f : ld1<DT>(feat, r1 + x1), wx);
      st1<DT>(out, oplane + e, lerp_exact(top, bot, y_w[oy]));
    }
    return;
  }
  float* myroll = roll + warp * kRoll * g.ow;
  const int r_in = lane / owv, xv = lane - r_in * owv;
  int y_have = y_i0[0] - 1;                   // highest source row already in the rolling buffer
  for (int oy0 = 0; oy0 < g.oh; oy0 += rows_per_iter) {
    const int oy_last = min(oy0 + rows_per_iter, g.oh) - 1;
    const int y_need = y_i1[oy_last];
    int y_from = max(y_have + 1, y_i0[oy0]);  // rows below the window are never read again
    for (int y = y_from; y <= y_need; ++y) {
      const int64_t row = plane + static_cast<int64_t>(y) * g.W;
      float* dst = myroll + (y & (kRoll - 1)) * g.ow + lane;
#pragma unroll
      for (int k = 0; k < kMaxOxPerLane; ++k)
        if (k < nk) {
          const float a = ld1<DT>(feat, row + xa[k]);
          dst[32 * k] = xw[k] == 0.f ? a : fmaf(ld1<DT>(feat, row + xb[k]), xw[k], a * (1.f - xw[k]));
        }
    }
    y_have = max(y_have, y_need);
    __syncwarp();
    const int oy = oy0 + r_in;
    if (oy < g.oh) {
      const float wy = y_w[oy], om = 1.f - wy;
      const float* top = myroll + (y_i0[oy] & (kRoll - 1)) * g.ow + xv * VEC;
      const float* bot = myroll + (y_i1[oy] & (kRoll - 1)) * g.ow + xv * VEC;
      float res[VEC];
#pragma unroll
      for (int v = 0; v < VEC; v += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(top + v);
        res[v] = t4.x; res[v + 1] = t4.y; res[v + 2] = t4.z; res[v + 3] = t4.w;
      }
      if (wy != 0.f) {
#pragma unroll
        for (int v = 0; v < VEC; v += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bot + v);
          res[v] = fmaf(b4.x, wy, res[v] * om); res[v + 1] = fmaf(b4.y, wy, res[v + 1] * om);
          res[v + 2] = fmaf(b4.z, wy, res[v + 2] * om); res[v + 3] = fmaf(b4.w, wy, res[v + 3] * om);
        }
      }
      stg_stream(static_cast<char*>(out) + (oplane + static_cast<int64_t>(oy) * g.ow + xv * VEC) * (16 / VEC), Elem<DT>::pack(res));
    }
    __syncwarp();
  }
}

// ---- strip path: one thread per (plane, 16-byte output strip), walking down the output rows ----------------------
// CTA = one box (b,k) x (256 / strips) consecutive channels, strips = ow / VEC (power of two).  A thread owns the VEC
// consecutive output columns of its strip and keeps the horizontally interpolated values of the two source rows the
// current output row blends (h_top, h_bot) in registers: a source row is interpolated along x once and reused by all the
// output rows that read it (`zoom` of them), every output row costs one vertical blend + one 128-bit store.  No shared
// memory traffic on the row path, no warp synchronisation; the row taps (shared by the whole box) sit in shared memory.
// Weights equal to zero short-circuit, so an integer box of the output size is a bit-exact copy (and takes a pure
// 128-bit copy loop when the source is 16-byte aligned).
template <int DT, int MINB>
__global__ void __launch_bounds__(256, MINB) crop_fwd_strip_kernel(const void* __restrict__ feat, const float* __restrict__ boxes,
                                                             void* __restrict__ out, Geo g, int strips) {
  constexpr int VEC = Elem<DT>::VEC;
  constexpr int ES = 16 / VEC;
  extern __shared__ __align__(16) float sm[];
  int* y_i0 = reinterpret_cast<int*>(sm);
  int* y_i1 = y_i0 + g.oh;
  float* y_w = reinterpret_cast<float*>(y_i1 + g.oh);
  const int planes_cta = 256 / strips;
  const int groups_c = (g.C + planes_cta - 1) / planes_cta;
  const int64_t bk = blockIdx.x / groups_c;
  const int c = static_cast<int>(blockIdx.x % groups_c) * planes_cta + static_cast<int>(threadIdx.x) / strips;
  const int sx = static_cast<int>(threadIdx.x) % strips;
  const int64_t b = bk / g.K;
  const float4 bx = __ldg(reinterpret_cast<const float4*>(boxes) + bk);  // y0,x0,y1,x1
  int* seg = reinterpret_cast<int*>(y_w + g.oh);  // [nseg + 1] first output row of every segment, then oh
  __shared__ int nseg_s;
  for (int o = threadIdx.x; o < g.oh; o += 256) axis_taps(bx.x, bx.z - bx.x, o, g.oh, g.H, y_i0[o], y_i1[o], y_w[o]);
  __syncthreads();
  if (threadIdx.x < 32) {  // ordered compaction of the segment starts by one warp
    int count = 0;
    for (int base = 0; base < g.oh; base += 32) {
      const int o = base + static_cast<int>(threadIdx.x);
      const bool st = o < g.oh && (o == 0 || y_i0[o] != y_i0[o - 1] || y_i1[o] != y_i1[o - 1] || (y_w[o] != 0.f) != (y_w[o - 1] != 0.f));
      const unsigned m = __ballot_sync(0xffffffffu, st);
      if (st) seg[count + __popc(m & ((1u << threadIdx.x) - 1u))] = o;
      count += __popc(m);
    }
    if (threadIdx.x == 0) {
      seg[count] = g.oh;
      nseg_s = count;
    }
  }
  __syncthreads();
  const int nseg = nseg_s;
  if (c >= g.C) return;
  const char* src = static_cast<const char*>(feat) + (b * g.C + c) * static_cast<int64_t>(g.H) * g.W * ES;
  char* dst = static_cast<char*>(out) + ((bk * g.C + c) * static_cast<int64_t>(g.oh) * g.ow + sx * VEC) * ES;
  const int64_t src_pitch = static_cast<int64_t>(g.W) * ES, dst_pitch = static_cast<int64_t>(g.ow) * ES;

  const bool is_copy = bx.x == floorf(bx.x) && bx.y == floorf(bx.y) && bx.z - bx.x == static_cast<float>(g.oh) &&
                       bx.w - bx.y == static_cast<float>(g.ow) && bx.x >= 0.f && bx.y >= 0.f &&
                       bx.z <= static_cast<float>(g.H) && bx.w <= static_cast<float>(g.W);
  if (is_copy) {
    const char* sp = src + (static_cast<int64_t>(bx.x) * g.W + static_cast<int64_t>(bx.y) + sx * VEC) * ES;
    if (((reinterpret_cast<uintptr_t>(sp) | static_cast<uintptr_t>(src_pitch)) & 15u) == 0) {
      int oy = 0;
#pragma unroll 1
      for (; oy + 4 <= g.oh; oy += 4) {  // four independent 128-bit loads in flight
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ldg_stream(sp + (oy + u) * src_pitch);
#pragma unroll
        for (int u = 0; u < 4; ++u) stg_stream(dst + (oy + u) * dst_pitch, v[u]);
      }
      for (; oy < g.oh; ++oy) stg_stream(dst + oy * dst_pitch, ldg_stream(sp + oy * src_pitch));
      return;
    }
  }

  // column taps of this strip
  int xa[VEC], xb[VEC];
  float xw[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    axis_taps(bx.y, bx.w - bx.y, sx * VEC + v, g.ow, g.W, xa[v], xb[v], xw[v]);
  }
  auto hpass = [&](int y, float* h) {
    const char* row = src + static_cast<int64_t>(y) * src_pitch;
    float a[VEC], b2[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {  // all 2*VEC loads in flight before the first use
      a[v] = ld1<DT>(row, xa[v]);
      b2[v] = ld1<DT>(row, xb[v]);
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) h[v] = xw[v] == 0.f ? a[v] : fmaf(b2[v], xw[v], a[v] * (1.f - xw[v]));
  };
  // Walk the SEGMENTS: maximal runs of output rows that blend the same two source rows (and agree on wy != 0) -- at
  // zoom z a segment has ~z rows.  Per segment at most one horizontal pass (the bottom row of one segment is the top row
  // of the next); per output row only the vertical blend, the pack and one 128-bit store remain.
  float h_top[VEC], h_bot[VEC];
  int cur0 = -1, cur1 = -1;  // source rows held in h_top / h_bot
  for (int s = 0; s < nseg; ++s) {
    const int ra = seg[s], rb = seg[s + 1];
    const int i0 = y_i0[ra], i1 = y_i1[ra];
    if (i0 != cur0) {
      if (i0 == cur1) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) h_top[v] = h_bot[v];
      } else {
        hpass(i0, h_top);
      }
      cur0 = i0;
    }
    char* drow = dst + ra * dst_pitch;
    if (y_w[ra] != 0.f) {
      if (i1 != cur1) {
        hpass(i1, h_bot);
        cur1 = i1;
      }
#pragma unroll 2
      for (int oy = ra; oy < rb; ++oy, drow += dst_pitch) {
        const float wy = y_w[oy];
        // packed fp32x2 multiply / fma: same roundings as the scalar fmaf(h_bot, wy, h_top * om), half the issue slots
        const float2 wy2 = make_float2(wy, wy), om2 = make_float2(1.f - wy, 1.f - wy);
        float res[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v += 2) {
          const float2 t = __fmul2_rn(make_float2(h_top[v], h_top[v + 1]), om2);
          const float2 r = __ffma2_rn(make_float2(h_bot[v], h_bot[v + 1]), wy2, t);
          res[v] = r.x;
          res[v + 1] = r.y;
        }
        stg_stream(drow, Elem<DT>::pack(res));
      }
    } else {
      const uint4 packed = Elem<DT>::pack(h_top);
      for (int oy = ra; oy < rb; ++oy, drow += dst_pitch) stg_stream(drow, packed);
    }
  }
}

// ---- backward, gather form: deterministic, no atomics -----------------------------------------------------------
// grad_feat[b,c] = sum_k Wy_k^T * G_k * Wx_k with two taps per output row / column.  CTA = (sample, band of R source
// rows, channel range); nobody else writes that part of grad_feat, and every sum runs in a fixed order (box, tap row,
// tap column), so the result is bit-reproducible and overlapping boxes need no atomics.
//   tables   the taps of one axis are non-decreasing in the output coordinate, so the outputs that reach a given source
//            row / column through tap 0 (or tap 1) form ONE contiguous range: per box the CTA tabulates taps + weights
//            (ow + oh evaluations), then finds the range ends by comparing neighbouring outputs (one writer per entry);
//   vertical T[r][c][ox] = sum over the output rows that tap source row r, read with 128-bit loads along ox;
//   horizon. acc[r][c][x] += sum over the output columns that tap x, from T in shared memory.
// grad_out is read from HBM once (+ the rows shared by two bands), grad_feat written once.
constexpr int kBwdThreads = 256, kBwdMaxGroup = 4;  // boxes per table group (<= 4: a 16-bit row x box mask)

struct BwdPlan {
  int R, cchunk, KL, csplit, bands, owp;
  size_t off_wx, off_wy, off_ax, off_ay, off_xr, off_yr, off_list, off_flag, smem;
};

// column of T: one pad float per 16, so that lanes reading every 4th .. 16th column (1x .. 4x zoom, 4 pixels per lane) hit distinct banks
__device__ __forceinline__ int tcol(int ox) { return ox + (ox >> 4); }

__device__ __forceinline__ bool empty2(int lo0, int hi0, int lo1, int hi1) { return hi0 < lo0 && hi1 < lo1; }

template <int DT, int VEC>
__global__ void __launch_bounds__(kBwdThreads) crop_bwd_gather_kernel(const void* __restrict__ gout, const float* __restrict__ boxes,
                                                                      float* __restrict__ gfeat, Geo g, BwdPlan p) {
  extern __shared__ __align__(16) unsigned char bsm[];
  float* T = reinterpret_cast<float*>(bsm);                        // [KL][R*cchunk][owp]
  float* wx = reinterpret_cast<float*>(bsm + p.off_wx);            // [KL][ow]   weight of tap 1 along x
  float* wy = reinterpret_cast<float*>(bsm + p.off_wy);            // [KL][oh]
  short2* ax = reinterpret_cast<short2*>(bsm + p.off_ax);          // [KL][ow]   {tap 0, tap 1} source columns
  short2* ay = reinterpret_cast<short2*>(bsm + p.off_ay);          // [KL][oh]
  short4* xr = reinterpret_cast<short4*>(bsm + p.off_xr);          // [KL][W]    {lo0, hi0, lo1, hi1} output columns per source column
  int4* yr = reinterpret_cast<int4*>(bsm + p.off_yr);              // [KL][R]    idem, output rows per source row of the band
  int* list = reinterpret_cast<int*>(bsm + p.off_list);            // [K]        boxes that touch the band, ascending
  int* flag = reinterpret_cast<int*>(bsm + p.off_flag);            // [K] + count
  const int tid = threadIdx.x;
  const int band = blockIdx.x % p.bands;
  const int64_t b = blockIdx.x / p.bands;
  const int y0 = band * p.R, rows = min(p.R, g.H - y0);
  const int cbeg = blockIdx.y * p.csplit, cend = min(cbeg + p.csplit, g.C);
  const float4* bxs = reinterpret_cast<const float4*>(boxes) + b * g.K;
  constexpr int E = DT == MSF_F32 ? 4 : 2;

  // boxes whose source rows [tap 0 of the first output row, tap 1 of the last] meet the band
  for (int k = tid; k < g.K; k += kBwdThreads) {
    const float4 bx = __ldg(bxs + k);
    int lo, hi, u;
    float w;
    axis_taps(bx.x, bx.z - bx.x, 0, g.oh, g.H, lo, u, w);
    axis_taps(bx.x, bx.z - bx.x, g.oh - 1, g.oh, g.H, u, hi, w);
    flag[k] = lo <= y0 + rows - 1 && hi >= y0;
  }
  __syncthreads();
  if (tid == 0) {
    int n = 0;
    for (int k = 0; k < g.K; ++k)
      if (flag[k]) list[n++] = k;
    flag[g.K] = n;
  }
  __syncthreads();
  const int nlist = flag[g.K];

  int g0 = 0;
  do {
    const int gn = min(p.KL, nlist - g0);
    // ---- tables of this group of boxes
    for (int i = tid; i < gn * g.ow; i += kBwdThreads) {
      const int j = i / g.ow, o = i - j * g.ow;
      const float4 bx = __ldg(bxs + list[g0 + j]);
      int a0, a1;
      float w;
      axis_taps(bx.y, bx.w - bx.y, o, g.ow, g.W, a0, a1, w);
      ax[i] = make_short2(static_cast<short>(a0), static_cast<short>(a1));
      wx[i] = w;
    }
    for (int i = tid; i < gn * g.oh; i += kBwdThreads) {
      const int j = i / g.oh, o = i - j * g.oh;
      const float4 bx = __ldg(bxs + list[g0 + j]);
      int a0, a1;
      float w;
      axis_taps(bx.x, bx.z - bx.x, o, g.oh, g.H, a0, a1, w);
      ay[i] = make_short2(static_cast<short>(a0), static_cast<short>(a1));
      wy[i] = w;
    }
    for (int i = tid; i < gn * g.W; i += kBwdThreads) xr[i] = make_short4(1, 0, 1, 0);
    for (int i = tid; i < gn * p.R; i += kBwdThreads) yr[i] = make_int4(1, 0, 1, 0);
    __syncthreads();
    // range ends: an output owns the start (end) of a range when its predecessor (successor) taps another source index.
    // Tap-1 ranges ignore the weight (so they stay contiguous); zero-weight outputs are skipped when the range is walked.
    for (int i = tid; i < gn * g.ow; i += kBwdThreads) {
      const int j = i / g.ow, o = i - j * g.ow;
      const short2 a = ax[i];
      const short2 pr = o ? ax[i - 1] : make_short2(-1, -1), nx = o + 1 < g.ow ? ax[i + 1] : make_short2(-1, -1);
      short* e0 = reinterpret_cast<short*>(xr + j * g.W + a.x);
      short* e1 = reinterpret_cast<short*>(xr + j * g.W + a.y);
      if (pr.x != a.x) e0[0] = static_cast<short>(o);
      if (nx.x != a.x) e0[1] = static_cast<short>(o);
      if (pr.y != a.y) e1[2] = static_cast<short>(o);
      if (nx.y != a.y) e1[3] = static_cast<short>(o);
    }
    for (int i = tid; i < gn * g.oh; i += kBwdThreads) {
      const int j = i / g.oh, o = i - j * g.oh;
      const short2 a = ay[i];
      const short2 pr = o ? ay[i - 1] : make_short2(-1, -1), nx = o + 1 < g.oh ? ay[i + 1] : make_short2(-1, -1);
      const int r0 = a.x - y0, r1 = a.y - y0;
      if (r0 >= 0 && r0 < rows) {
        int* e = reinterpret_cast<int*>(yr + j * p.R + r0);
        if (pr.x != a.x) e[0] = o;
        if (nx.x != a.x) e[1] = o;
      }
      if (r1 >= 0 && r1 < rows) {
        int* e = reinterpret_cast<int*>(yr + j * p.R + r1);
        if (pr.y != a.y) e[2] = o;
        if (nx.y != a.y) e[3] = o;
      }
    }
    __syncthreads();

    // per thread: which boxes of the group reach which band row (bit r*4+j), and the boxes' source-column extents
    unsigned rowmask = 0;
    int xlo[kBwdMaxGroup], xhi[kBwdMaxGroup];
#pragma unroll
    for (int j = 0; j < kBwdMaxGroup; ++j) {
      xlo[j] = 1;
      xhi[j] = 0;
      if (j < gn) {
        xlo[j] = ax[j * g.ow].x;
        xhi[j] = ax[j * g.ow + g.ow - 1].y;
        for (int r = 0; r < rows; ++r) {
          const int4 yy = yr[j * p.R + r];
          if (!empty2(yy.x, yy.y, yy.z, yy.w)) rowmask |= 1u << (r * kBwdMaxGroup + j);
        }
      }
    }
    for (int c0 = cbeg; c0 < cend; c0 += p.cchunk) {
      const int cn = min(p.cchunk, cend - c0);
      const int rc_n = rows * cn;
      const int tstride = p.R * p.cchunk * p.owp;  // floats of T per box
      // ---- vertical: T[j][r][c][ox] for every box of the group
      const int owv = g.ow / VEC;
      const float inv_owv = 1.f / owv;
      for (int j = 0; j < gn; ++j) {
        const int64_t plane0 = ((b * g.K + list[g0 + j]) * g.C + c0) * static_cast<int64_t>(g.oh) * g.ow;
        const float* wyj = wy + j * g.oh;
        for (int i = tid; i < rc_n * owv; i += kBwdThreads) {
          const int rc = __float2int_rz((i + 0.5f) * inv_owv), v = i - rc * owv;
          const int r = (rc >= cn) + (rc >= 2 * cn) + (rc >= 3 * cn), c = rc - r * cn;
          if (!((rowmask >> (r * kBwdMaxGroup + j)) & 1u)) continue;
          const int4 yy = yr[j * p.R + r];
          const char* src = static_cast<const char*>(gout) + (plane0 + static_cast<int64_t>(c) * g.oh * g.ow + static_cast<int64_t>(v) * VEC) * E;
          float t[VEC];
#pragma unroll
          for (int q = 0; q < VEC; ++q) t[q] = 0.f;
#pragma unroll
          for (int tap = 0; tap < 2; ++tap) {
            const int lo = tap ? yy.z : yy.x, hi = tap ? yy.w : yy.y;
            if constexpr (VEC == 1) {
#pragma unroll 1
              for (int oy = lo; oy <= hi; ++oy) {
                const float w1 = wyj[oy];
                if (tap && w1 == 0.f) continue;
                t[0] = fmaf(ld1<DT>(src + static_cast<int64_t>(oy) * g.ow * E, 0), tap ? w1 : 1.f - w1, t[0]);
              }
            } else {
              // four rows per trip: the loads are issued together (the walk is latency-bound otherwise)
#pragma unroll 1
              for (int oy = lo; oy <= hi; oy += 4) {
                uint4 raw[4];
                float w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const bool on = oy + u <= hi;
                  const float w1 = on ? wyj[oy + u] : 0.f;
                  w[u] = on ? (tap ? w1 : 1.f - w1) : 0.f;
                  // a zero weight (tap 1 of an output that sits exactly on a source row, or past the range) reads nothing
                  raw[u] = (on && !(tap && w1 == 0.f)) ? ldg_stream(src + static_cast<int64_t>(oy + u) * g.ow * E) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  float f[VEC];
                  Elem<DT>::unpack(raw[u], f);
#pragma unroll
                  for (int q = 0; q < VEC; ++q) t[q] = fmaf(f[q], w[u], t[q]);
                }
              }
            }
          }
          float* dst = T + j * tstride + (r * p.cchunk + c) * p.owp;
#pragma unroll
          for (int q = 0; q < VEC; ++q) dst[tcol(v * VEC + q)] = t[q];
        }
      }
      __syncthreads();
      // ---- horizontal: one thread per source pixel (r, c, x) of the band walks the boxes in order; consecutive lanes own
      // consecutive columns, so a warp diverges only where a box ends
      const float inv_w = 1.f / g.W;
      for (int i = tid; i < rc_n * g.W; i += kBwdThreads) {
        const int rc = __float2int_rz((i + 0.5f) * inv_w), x = i - rc * g.W;
        const int r = (rc >= cn) + (rc >= 2 * cn) + (rc >= 3 * cn), c = rc - r * cn;
        float* out = gfeat + ((b * g.C + c0 + c) * static_cast<int64_t>(g.H) + y0 + r) * g.W + x;
        float a = g0 ? *out : 0.f;  // what the earlier groups of boxes left (same owner thread every time)
        const unsigned live = (rowmask >> (r * kBwdMaxGroup)) & ((1u << kBwdMaxGroup) - 1u);
#pragma unroll
        for (int j = 0; j < kBwdMaxGroup; ++j) {
          if (!((live >> j) & 1u) || x > xhi[j] || x < xlo[j]) continue;
          const short4 xx = xr[j * g.W + x];
          const float* trow = T + j * tstride + (r * p.cchunk + c) * p.owp;
          const float* wxj = wx + j * g.ow;
          float s = 0.f;
#pragma unroll 1
          for (int ox = xx.x; ox <= xx.y; ++ox) s = fmaf(trow[tcol(ox)], 1.f - wxj[ox], s);
#pragma unroll 1
          for (int ox = xx.z; ox <= xx.w; ++ox) {
            const float w1 = wxj[ox];
            if (w1 != 0.f) s = fmaf(trow[tcol(ox)], w1, s);
          }
          a += s;
        }
        *out = a;
      }
      __syncthreads();  // T is rewritten by the next channel pass
    }
    g0 += p.KL;
  } while (g0 < nlist);
}

// Shared-memory plan of the backward: band height R (<= 4), channels per pass, boxes per table group.
bool plan_bwd_in(const Geo& g, BwdPlan& p, size_t kBudget) {
  auto up16 = [](size_t v) { return (v + 15) & ~static_cast<size_t>(15); };
  p.owp = g.ow + (g.ow >> 4) + 1;
  const size_t fixed = up16(static_cast<size_t>(g.K) * 4) + up16(static_cast<size_t>(g.K) * 4 + 4);
  for (int R = 4; R >= 1; R >>= 1) {
    // per box: tap weights + tap indices of both axes, column ranges, row ranges; plus its slab of T per channel
    const size_t per_box = 2 * up16(static_cast<size_t>(g.ow) * 4) + 2 * up16(static_cast<size_t>(g.oh) * 4) + up16(static_cast<size_t>(g.W) * 8) +
                           static_cast<size_t>(R) * 16;
    const size_t t_chan = static_cast<size_t>(R) * p.owp * 4;  // T of one box and one channel
    if (fixed + per_box + t_chan > kBudget) continue;
    // boxes per group: the boxes that meet one band (4 for a 4x4 grid of footprints); more run as further groups
    int KL = std::min(g.K, kBwdMaxGroup);
    while (KL > 1 && fixed + KL * (per_box + t_chan) > kBudget) --KL;
    const size_t room = kBudget - fixed - KL * per_box;
    const int cchunk = static_cast<int>(std::max<size_t>(1, std::min<size_t>(static_cast<size_t>(g.C), room / (KL * t_chan))));
    p.R = R;
    p.cchunk = cchunk;
    p.KL = KL;
    size_t o = up16(static_cast<size_t>(KL) * R * cchunk * p.owp * 4);
    p.off_wx = o; o += static_cast<size_t>(KL) * up16(static_cast<size_t>(g.ow) * 4);
    p.off_wy = o; o += static_cast<size_t>(KL) * up16(static_cast<size_t>(g.oh) * 4);
    p.off_ax = o; o += static_cast<size_t>(KL) * up16(static_cast<size_t>(g.ow) * 4);
    p.off_ay = o; o += static_cast<size_t>(KL) * up16(static_cast<size_t>(g.oh) * 4);
    p.off_xr = o; o += static_cast<size_t>(KL) * up16(static_cast<size_t>(g.W) * 8);
    p.off_yr = o; o += static_cast<size_t>(KL) * R * 16;
    p.off_list = o; o += up16(static_cast<size_t>(g.K) * 4);
    p.off_flag = o; o += up16(static_cast<size_t>(g.K) * 4 + 4);
    p.smem = o;
    return o <= kBudget;
  }
  return false;
}

// 48 KB keeps 4 CTAs per SM (the two phases of different CTAs overlap); very wide maps need more room for their tables
bool plan_bwd(const Geo& g, BwdPlan& p) {
  for (size_t budget : {static_cast<size_t>(48) << 10, static_cast<size_t>(96) << 10, static_cast<size_t>(200) << 10})
    if (plan_bwd_in(g, p, budget)) return true;
  return false;
}

int check(const void* a, const float* boxes, const void* o, int64_t B, int C, int H, int W, int K, int oh, int ow, int dtype) {
  MSF_REQUIRE(dtype_ok(dtype), MSF_ERR_INVALID, "bad dtype %d", dtype);
  MSF_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0 && K > 0 && oh > 0 && ow > 0, MSF_ERR_INVALID, "bad shape");
  MSF_REQUIRE(B == 0 || (a && boxes && o), MSF_ERR_INVALID, "NULL pointer");
  MSF_REQUIRE(aligned16(boxes), MSF_ERR_INVALID, "boxes must be 16-byte aligned");
  return MSF_OK;
}

unsigned grid_for(int64_t total) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 16;
  return static_cast<unsigned>(blocks < cap ? blocks : cap);
}

}  // namespace
}  // namespace msf

using namespace msf;

extern "C" int msf_crop_resample_fwd(const void* feat, int64_t B, int C, int H, int W, const float* boxes, int K, int oh,
                                     int ow, int dtype, void* out, void* stream) {
  if (int rc = check(feat, boxes, out, B, C, H, W, K, oh, ow, dtype)) return rc;
  if (B == 0) return MSF_OK;
  const Geo g{B, C, H, W, K, oh, ow};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vec = 16 / static_cast<int>(dtype_size(dtype));
  const bool vec_ok = (ow % vec == 0) && aligned16(out);
  const int64_t rows = B * K * static_cast<int64_t>(C) * oh;
  // fast path: output rows of 32..256 pixels (multiple of 32); rolling buffer of kRoll source rows per warp
  const int owv = vec_ok ? ow / vec : 0;
  int kroll = 4;
  while (kroll < 32 && kroll * 2 * ow <= 1024) kroll *= 2;
  const size_t fast_smem = (static_cast<size_t>((3 * oh + 3) & ~3) + static_cast<size_t>(8) * kroll * ow) * 4;
  ProfScope prof(stream, MSF_K_CROP_FWD, (static_cast<double>(B) * C * H * W + static_cast<double>(rows) * ow) * dtype_size(dtype) + 16.0 * B * K);
  const bool fast_ok = vec_ok && ow % 32 == 0 && ow <= 32 * kMaxOxPerLane && owv <= 32 && (owv & (owv - 1)) == 0 &&
                       fast_smem <= 48 * 1024 && B * K < (1ll << 24);
  // strip path: ow/vec strips per plane (power of two <= 256); the row taps + segment list of one box ((4 * oh + 1) * 4 bytes) fit the
  // 48 KB of dynamic shared memory a kernel gets without opting in
  const bool strip_ok = vec_ok && owv >= 1 && owv <= 256 && (owv & (owv - 1)) == 0 && oh <= 3000 &&
                        B * K * static_cast<int64_t>((C + 256 / owv - 1) / (256 / owv)) < (1ll << 31);
  if (strip_ok) {
    const int planes_cta = 256 / owv;
    const int64_t ctas = B * K * ((C + planes_cta - 1) / planes_cta);
    const size_t smem = (static_cast<size_t>(4) * oh + 1) * 4;  // row taps + weights + segment starts
    // 16-bit elements: 64 registers (4 CTAs per SM) against the natural 76 (3 CTAs); MSF_CROP_STRIP_MINB picks for measurements
    static const int minb = [] {
      const char* e = getenv("MSF_CROP_STRIP_MINB");
      return e && atoi(e) == 3 ? 3 : 4;
    }();
    if (minb == 4) {
      MSF_DISPATCH_DTYPE(dtype, (crop_fwd_strip_kernel<DT, 4><<<static_cast<unsigned>(ctas), 256, smem, st>>>(feat, boxes, out, g, owv)));
    } else {
      MSF_DISPATCH_DTYPE(dtype, (crop_fwd_strip_kernel<DT, 3><<<static_cast<unsigned>(ctas), 256, smem, st>>>(feat, boxes, out, g, owv)));
    }
  } else if (fast_ok) {
    const int64_t ctas = B * K * ((C + 7) / 8);
    MSF_DISPATCH_DTYPE(dtype, (crop_fwd_fast_kernel<DT><<<static_cast<unsigned>(ctas), kFastThreads, fast_smem, st>>>(feat, boxes, out, g, kroll)));
  } else if (vec_ok) {
    const int64_t total = rows * (ow / vec);
    MSF_DISPATCH_DTYPE(dtype, (crop_fwd_kernel<DT, Elem<DT>::VEC><<<grid_for(total), 256, 0, st>>>(feat, boxes, out, g, total)));
  } else {
    const int64_t total = rows * ow;
    MSF_DISPATCH_DTYPE(dtype, (crop_fwd_kernel<DT, 1><<<grid_for(total), 256, 0, st>>>(feat, boxes, out, g, total)));
  }
  MSF_LAUNCH_OK("crop_fwd_kernel");
  return MSF_OK;
}

extern "C" int msf_crop_resample_bwd(const void* grad_out, int64_t B, int C, int H, int W, const float* boxes, int K,
                                     int oh, int ow, int dtype, float* grad_feat, void* stream) {
  if (int rc = check(grad_out, boxes, grad_feat, B, C, H, W, K, oh, ow, dtype)) return rc;
  if (B == 0) return MSF_OK;
  MSF_REQUIRE(oh < 32768 && ow < 32768 && H < 32768 && W < 32768, MSF_ERR_UNSUPPORTED, "crop backward: sides must be < 32768");
  const Geo g{B, C, H, W, K, oh, ow};
  const int vec = 16 / static_cast<int>(dtype_size(dtype));
  const bool vec_ok = ow % vec == 0 && aligned16(grad_out);
  BwdPlan p{};
  MSF_REQUIRE(plan_bwd(g, p), MSF_ERR_UNSUPPORTED, "crop backward: the tap tables of one box (W=%d, oh=%d, ow=%d, K=%d) exceed shared memory", W, oh, ow, K);
  p.bands = (H + p.R - 1) / p.R;
  MSF_REQUIRE(B * p.bands < (1ll << 31), MSF_ERR_UNSUPPORTED, "crop backward: B * H too large");
  // channel ranges per CTA: as few as fill the machine ~8 CTAs deep (the tables are rebuilt per CTA)
  int64_t splits = (static_cast<int64_t>(kNumSMs) * 8 + B * p.bands - 1) / (B * p.bands);
  const int64_t max_splits = (C + p.cchunk - 1) / p.cchunk;
  splits = std::max<int64_t>(1, std::min<int64_t>({splits, max_splits, 65535}));
  p.csplit = static_cast<int>(((C + splits - 1) / splits + p.cchunk - 1) / p.cchunk) * p.cchunk;
  const dim3 grid(static_cast<unsigned>(B * p.bands), static_cast<unsigned>((C + p.csplit - 1) / p.csplit));
  const int64_t total = B * K * static_cast<int64_t>(C) * oh * ow;
  ProfScope prof(stream, MSF_K_CROP_BWD, static_cast<double>(total) * dtype_size(dtype) + 4.0 * B * C * H * W + 16.0 * B * K);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the packed tables assume 4-byte rows of shorts pairs; the transposed accumulator layouts need no alignment beyond 16 B slabs
  MSF_DISPATCH_DTYPE(dtype, {
    if (vec_ok) {
      auto kern = crop_bwd_gather_kernel<DT, Elem<DT>::VEC>;
      if (p.smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem));
      kern<<<grid, kBwdThreads, p.smem, st>>>(grad_out, boxes, grad_feat, g, p);
    } else {
      auto kern = crop_bwd_gather_kernel<DT, 1>;
      if (p.smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem));
      kern<<<grid, kBwdThreads, p.smem, st>>>(grad_out, boxes, grad_feat, g, p);
    }
  });
  MSF_LAUNCH_OK("crop_bwd_gather_kernel");
  return MSF_OK;
}
