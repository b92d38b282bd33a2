// A2: crop the low-magnification feature map to each high-magnification tile's footprint and
// bilinearly resample it.  Integer boxes whose size equals (oh, ow) reproduce the reference's only
// feature-map crop, src/models/hooknet.py:29-32 (`x[:, :, 12:20, 12:20]`), bit-exactly; everything else
// is the north_star extension (F.interpolate bilinear / align_corners=False on the crop).
//
// HBM-bound: output rows are written with 128-bit stores (VEC consecutive ox per thread); the four
// taps of neighbouring outputs overlap, so input traffic is the footprint once (L1/L2 hits after).
// Algorithmic bytes: B*C*(union of footprints)*e read + B*K*C*oh*ow*e written + 16 B per box.
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

// ---- fast path: warp-cooperative rows, 128-bit stores ---------------------------------------------------
// CTA = one box (b,k) x a group of channels; a lane group of ow/VEC lanes owns one output row at a time.
//   1. the group loads the source span of the row from the two source rows with CONSECUTIVE lanes on consecutive
//      pixels (coalesced sectors, L1-resident for the `zoom` output rows that share them) and interpolates
//      vertically;
//   2. every source pixel is parked in the warp's shared row as float2 {v[s], v[s+1]-v[s]} (the neighbour comes
//      from a shuffle);
//   3. each lane forms its VEC outputs with ONE 64-bit shared load + one FMA each (out = v + wx*dv; wx == 0 returns
//      v exactly, so an integer box of the output size is a bit-exact copy) and stores 16 bytes.
// x taps live in registers (a lane always owns the same output columns); y taps in shared memory.
constexpr int kFastThreads = 256;

template <int DT>
__global__ void __launch_bounds__(kFastThreads) crop_fwd_fast_kernel(const void* __restrict__ feat, const float* __restrict__ boxes,
                                                                     void* __restrict__ out, Geo g, int ch_per_cta, int park_stride) {
  constexpr int VEC = Elem<DT>::VEC;
  extern __shared__ __align__(16) float sm[];
  const int tpr = g.ow / VEC;            // lanes per output row (power of two <= 32)
  const int rows_per_warp = 32 / tpr;
  int* y_i0 = reinterpret_cast<int*>(sm);
  int* y_i1 = y_i0 + g.oh;
  float* y_w = reinterpret_cast<float*>(y_i1 + g.oh);
  float2* parks = reinterpret_cast<float2*>(sm + ((3 * g.oh + 3) & ~3));  // [warps * rows_per_warp][park_stride]
  const int groups_c = (g.C + ch_per_cta - 1) / ch_per_cta;
  const int64_t bk = blockIdx.x / groups_c;
  const int c_begin = static_cast<int>(blockIdx.x % groups_c) * ch_per_cta;
  const int c_end = min(c_begin + ch_per_cta, g.C);
  const int64_t b = bk / g.K;
  const float4 bx = __ldg(reinterpret_cast<const float4*>(boxes) + bk);  // y0,x0,y1,x1
  for (int o = threadIdx.x; o < g.oh; o += kFastThreads) axis_taps(bx.x, bx.z - bx.x, o, g.oh, g.H, y_i0[o], y_i1[o], y_w[o]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / tpr, lig = lane % tpr;
  int xs, xl0, xl1;
  float wtmp;
  axis_taps(bx.y, bx.w - bx.y, 0, g.ow, g.W, xs, xl0, wtmp);            // first source column of the box
  axis_taps(bx.y, bx.w - bx.y, g.ow - 1, g.ow, g.W, xl0, xl1, wtmp);    // last one
  const int fw = xl1 - xs + 1;                                           // <= park_stride
  int xi0[VEC];
  float xw[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    int a0, a1;
    axis_taps(bx.y, bx.w - bx.y, lig * VEC + v, g.ow, g.W, a0, a1, xw[v]);
    if (a1 == a0) xw[v] = 0.f;  // clamped second tap
    xi0[v] = a0 - xs;
  }
  __syncthreads();
  float2* park = parks + (warp * rows_per_warp + grp) * park_stride;
  const int rows_per_iter = (kFastThreads / 32) * rows_per_warp;
  const int total_rows = (c_end - c_begin) * g.oh;
  const int iters = (total_rows + rows_per_iter - 1) / rows_per_iter;
  const int64_t plane_sz = static_cast<int64_t>(g.H) * g.W;
  for (int itn = 0; itn < iters; ++itn) {
    const int r = itn * rows_per_iter + warp * rows_per_warp + grp;
    const bool valid = r < total_rows;
    const int cl = valid ? r / g.oh : 0, oy = valid ? r - cl * g.oh : 0;
    const int c = c_begin + cl;
    const float wy = y_w[oy], om = 1.f - wy;
    const int64_t r0 = (b * g.C + c) * plane_sz + static_cast<int64_t>(y_i0[oy]) * g.W + xs;
    const int64_t r1 = (b * g.C + c) * plane_sz + static_cast<int64_t>(y_i1[oy]) * g.W + xs;
    for (int j0 = 0; j0 < fw; j0 += tpr) {  // group-uniform trip count
      const int j = j0 + lig;
      float vj = 0.f;
      if (valid && j < fw) {
        const float t = ld1<DT>(feat, r0 + j);
        vj = wy == 0.f ? t : fmaf(ld1<DT>(feat, r1 + j), wy, t * om);
      }
      float vn = __shfl_down_sync(0xffffffffu, vj, 1, tpr);     // v[j+1] from the next lane of the group (warp-uniform loop)
      if (lig == tpr - 1) {                                      // ... or the first pixel of the next batch
        vn = vj;
        if (valid && j + 1 < fw) {
          const float t = ld1<DT>(feat, r0 + j + 1);
          vn = wy == 0.f ? t : fmaf(ld1<DT>(feat, r1 + j + 1), wy, t * om);
        }
      }
      if (valid && j < fw) park[j] = make_float2(vj, j + 1 < fw ? vn - vj : 0.f);
    }
    __syncwarp();
    if (valid) {
      float res[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float2 p2 = park[xi0[v]];
        res[v] = fmaf(xw[v], p2.y, p2.x);
      }
      const int64_t o = (((bk * g.C + c) * g.oh + oy) * static_cast<int64_t>(g.ow)) + lig * VEC;
      stg_stream(static_cast<char*>(out) + o * (16 / VEC), Elem<DT>::pack(res));
    }
    __syncwarp();
  }
}

template <int DT>
__global__ void __launch_bounds__(256) crop_bwd_kernel(const void* __restrict__ gout, const float* __restrict__ boxes,
                                                       float* __restrict__ gfeat, Geo g, int64_t total) {
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t r = t;
    const int ox = static_cast<int>(r % g.ow); r /= g.ow;
    const int oy = static_cast<int>(r % g.oh); r /= g.oh;
    const int c = static_cast<int>(r % g.C); r /= g.C;
    const int k = static_cast<int>(r % g.K);
    const int64_t b = r / g.K;
    const float4 bx = __ldg(reinterpret_cast<const float4*>(boxes) + b * g.K + k);
    int y0, y1, x0, x1;
    float wy, wx;
    axis_taps(bx.x, bx.z - bx.x, oy, g.oh, g.H, y0, y1, wy);
    axis_taps(bx.y, bx.w - bx.y, ox, g.ow, g.W, x0, x1, wx);
    const float go = ld1<DT>(gout, t);
    float* plane = gfeat + (b * g.C + c) * static_cast<int64_t>(g.H) * g.W;
    atomicAdd(plane + static_cast<int64_t>(y0) * g.W + x0, go * (1.f - wy) * (1.f - wx));
    if (wx != 0.f) atomicAdd(plane + static_cast<int64_t>(y0) * g.W + x1, go * (1.f - wy) * wx);
    if (wy != 0.f) {
      atomicAdd(plane + static_cast<int64_t>(y1) * g.W + x0, go * wy * (1.f - wx));
      if (wx != 0.f) atomicAdd(plane + static_cast<int64_t>(y1) * g.W + x1, go * wy * wx);
    }
  }
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
  const int tpr = vec_ok ? ow / vec : 0;
  const int park_stride = (W + 2 + 1) & ~1;
  const size_t fast_smem = static_cast<size_t>((3 * oh + 3) & ~3) * 4 +
                           static_cast<size_t>(kFastThreads / 32) * (tpr > 0 && tpr <= 32 ? 32 / tpr : 1) * park_stride * 8;
  const bool fast_ok = vec_ok && tpr >= 1 && tpr <= 32 && (tpr & (tpr - 1)) == 0 && fast_smem <= 48 * 1024 && B * K < (1ll << 24);
  if (fast_ok) {
    // enough CTAs for several waves of 148 SMs x 8 CTAs while keeping >= 256 output rows per CTA
    int ch_per_cta = C;
    while (ch_per_cta > 1 && (static_cast<int64_t>(ch_per_cta / 2) * oh >= 256) &&
           B * K * ((C + ch_per_cta - 1) / ch_per_cta) < static_cast<int64_t>(kNumSMs) * 64)
      ch_per_cta = (ch_per_cta + 1) / 2;
    const int64_t ctas = B * K * ((C + ch_per_cta - 1) / ch_per_cta);
    MSF_DISPATCH_DTYPE(dtype, (crop_fwd_fast_kernel<DT><<<static_cast<unsigned>(ctas), kFastThreads, fast_smem, st>>>(
                                  feat, boxes, out, g, ch_per_cta, park_stride)));
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
  const Geo g{B, C, H, W, K, oh, ow};
  const int64_t total = B * K * static_cast<int64_t>(C) * oh * ow;
  MSF_DISPATCH_DTYPE(dtype, (crop_bwd_kernel<DT><<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(grad_out, boxes, grad_feat, g, total)));
  MSF_LAUNCH_OK("crop_bwd_kernel");
  return MSF_OK;
}
