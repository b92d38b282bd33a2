// D1: on-device data path of the target / context views: what the dataset does per sample on CPU
// workers -- src/utils/data/bcss.py:171-177 `blockshaped(img, 256, 256)[jigsaw_idx]`, then per tile Resize(224) +
// Normalize + ToTensor (transforms[2]) -- as one kernel from the uint8 source image:
//   out[b*K + j, :, oy, ox] = normalise(bilinear(tile perm[b, j] of src[b]))          (NHWC output, K = grid*grid tiles)
// tile t of an (H, W) image with a g x g grid covers rows [th*(t/g), +th), cols [tw*(t%g), +tw) (raster order, exactly
// blockshaped); bilinear = F.interpolate(align_corners=False) on the cropped tile (taps clamp at the tile border);
// normalise = (v - 255*mean[c]) / (255*std[c]) (albumentations.Normalize with max_pixel_value 255).  grid = 1 gives
// the context view (whole image -> oh x ow).  The random photometric / geometric augmentations of the reference's
// albumentations pipelines are not reproduced (and cv2's fixed-point uint8 resize rounds differently): the tiling and
// the permutation are exact, the resampling is the float bilinear formula.
// HBM-bound: B*H*W*3 bytes read + B*K*oh*ow*3*e written; one CTA per output row, one thread per output pixel (12 byte
// loads from L1, 3 stores).
#include "common.cuh"

namespace msf {
namespace {

struct TileGeo {
  int H, W, g, th, tw, oh, ow;
};

// taps of output index o: source coordinate (o + 0.5) * scale - 0.5 with scale = len / osz formed in double on the host (in
// fp32 the scale's rounding error times the output index reaches 1e-5, which a 255-level difference between neighbouring
// pixels turns into a visible 1e-3; a double division per pixel on the device would dominate the kernel)
__device__ __forceinline__ void taps(int o, double scale, int len, int& i0, int& i1, float& w) {
  const double s = fmax(fma(o + 0.5, scale, -0.5), 0.0);
  const int f = min(static_cast<int>(s), len - 1);
  i0 = f;
  i1 = min(f + 1, len - 1);
  w = f < len - 1 ? static_cast<float>(s - static_cast<double>(f)) : 0.f;
}

template <int ODT>
__device__ __forceinline__ void store1(void* base, int64_t i, float v) {
  if constexpr (ODT == MSF_F32) static_cast<float*>(base)[i] = v;
  else if constexpr (ODT == MSF_BF16) static_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v);
  else static_cast<__half*>(base)[i] = __float2half_rn(v);
}

// One CTA per (tile image t = b*K + j, block of kRowsPerCta output rows): the tile lookup is CTA-uniform, each thread
// walks the block's pixels (consecutive threads = consecutive pixels of a row) -- no per-pixel division.
constexpr int kRowsPerCta = 16;
template <int ODT>
__global__ void __launch_bounds__(256) jigsaw_tiles_kernel(const unsigned char* __restrict__ src, const int64_t* __restrict__ perm,
                                                           void* __restrict__ out, TileGeo g, double scale_y, double scale_x, float a0, float a1,
                                                           float a2, float b0, float b1, float b2, int* __restrict__ status) {
  const int K = g.g * g.g;
  const float sc[3] = {a0, a1, a2}, sh[3] = {b0, b1, b2};
  const int row_blocks = (g.oh + kRowsPerCta - 1) / kRowsPerCta;
  const int rb = static_cast<int>(blockIdx.x % row_blocks);
  const int64_t t = blockIdx.x / row_blocks;    // b*K + j
  const int j = static_cast<int>(t % K);
  const int64_t b = t / K;
  int64_t tile = perm ? perm[b * K + j] : j;
  if (tile < 0) tile += K;                      // Python-style negative index
  if (tile < 0 || tile >= K) {                  // the reference would raise IndexError; flag it and clamp
    if (status && threadIdx.x == 0) atomicOr(status, 1);
    tile = min(max(tile, int64_t{0}), static_cast<int64_t>(K - 1));
  }
  const int ty = static_cast<int>(tile) / g.g, tx = static_cast<int>(tile) % g.g;
  const unsigned char* base = src + ((b * g.H + ty * g.th) * static_cast<int64_t>(g.W) + tx * g.tw) * 3;
  const int oy0 = rb * kRowsPerCta, rows = min(kRowsPerCta, g.oh - oy0);
  const int64_t src_pitch = static_cast<int64_t>(g.W) * 3;
  for (int p = threadIdx.x; p < rows * g.ow; p += 256) {
    const int r = p / g.ow, ox = p - r * g.ow, oy = oy0 + r;
    int y0, y1, x0, x1;
    float wy, wx;
    taps(oy, scale_y, g.th, y0, y1, wy);
    taps(ox, scale_x, g.tw, x0, x1, wx);
    const unsigned char* r0 = base + y0 * src_pitch;
    const unsigned char* r1 = base + y1 * src_pitch;
    const float omx = 1.f - wx, omy = 1.f - wy;
    const int64_t e = (t * g.oh + oy) * g.ow + ox;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float p00 = __ldg(r0 + x0 * 3 + c), p01 = __ldg(r0 + x1 * 3 + c), p10 = __ldg(r1 + x0 * 3 + c), p11 = __ldg(r1 + x1 * 3 + c);
      const float top = wx == 0.f ? p00 : fmaf(p01, wx, p00 * omx);
      const float bot = wx == 0.f ? p10 : fmaf(p11, wx, p10 * omx);
      const float v = wy == 0.f ? top : fmaf(bot, wy, top * omy);
      store1<ODT>(out, e * 3 + c, fmaf(v, sc[c], sh[c]));
    }
  }
}

}  // namespace
}  // namespace msf

using namespace msf;

extern "C" int msf_jigsaw_tiles(const uint8_t* src, int64_t B, int H, int W, int grid, const int64_t* perm, int oh, int ow,
                                const float* mean3, const float* std3, void* out, int out_dtype, int32_t* status_flag, void* stream) {
  MSF_REQUIRE(dtype_ok(out_dtype), MSF_ERR_INVALID, "bad dtype");
  MSF_REQUIRE(B >= 0 && H > 0 && W > 0 && grid >= 1 && oh > 0 && ow > 0, MSF_ERR_INVALID, "bad sizes");
  MSF_REQUIRE(H % grid == 0 && W % grid == 0, MSF_ERR_INVALID, "%d x %d is not evenly divisible into a %d x %d grid", H, W, grid, grid);  // bcss.py:212-213
  MSF_REQUIRE(mean3 && std3 && std3[0] > 0.f && std3[1] > 0.f && std3[2] > 0.f, MSF_ERR_INVALID, "mean / std (host, 3 values) invalid");
  if (B == 0) return MSF_OK;
  MSF_REQUIRE(src && out, MSF_ERR_INVALID, "NULL pointer");
  const TileGeo g{H, W, grid, H / grid, W / grid, oh, ow};
  const int64_t total = B * grid * grid * static_cast<int64_t>(oh) * ow;
  const int64_t blocks = B * grid * grid * static_cast<int64_t>((oh + kRowsPerCta - 1) / kRowsPerCta);  // one CTA per block of output rows
  MSF_REQUIRE(blocks < (int64_t{1} << 31), MSF_ERR_UNSUPPORTED, "%lld row blocks must be < 2^31", static_cast<long long>(blocks));
  const double scale_y = static_cast<double>(H / grid) / oh, scale_x = static_cast<double>(W / grid) / ow;
  float a[3], b[3];
  for (int c = 0; c < 3; ++c) { a[c] = 1.f / (255.f * std3[c]); b[c] = -mean3[c] / std3[c]; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(stream, MSF_K_JIGSAW_TILES, static_cast<double>(B) * H * W * 3 + static_cast<double>(total) * 3 * dtype_size(out_dtype));
  MSF_DISPATCH_DTYPE(out_dtype, (jigsaw_tiles_kernel<DT><<<static_cast<unsigned>(blocks), 256, 0, st>>>(src, perm, out, g, scale_y, scale_x, a[0], a[1], a[2], b[0], b[1],
                                                                                                      b[2], status_flag)));
  MSF_LAUNCH_OK("jigsaw_tiles_kernel");
  return MSF_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// D1b: every view of a step from the uint8 source tiles in ONE launch, written directly in the layout the stem
// convolution consumes (S1, stem_s2d.cu: zero-padded by 3, 2x2 pixel-unshuffled, 16 channels, NHWC).
//   view i = normalise(hflip?(resize_bilinear(src[sample_i][y0:y1, x0:x1], oh x ow)))
// which is what the reference's geometric augmentations do to a sample once their random numbers are drawn
// (albumentations RandomResizedCrop(224, scale=(0.5, 1)) -> integer crop box, HorizontalFlip, Normalize:
// tools/ssl_train.py:175-217), applied to the whole tile (context views) or to tile jigsaw_idx[j] of
// blockshaped(img, 256, 256) (target views, src/utils/data/bcss.py:171-177: the box is then the tile origin + the crop
// inside the tile -- the caller builds it with the exact integer tiling of blockshaped).  The photometric augmentations
// (ColorJitter / ToGray / blur / sharpen) are not reproduced.  Host -> device traffic drops from the 34 bf16 views per tile
// (10.2 MB) to the uint8 source (3.1 MB) + 24 bytes per view, and the separate layout pass (msf_stem_s2d) disappears.
// HBM-bound, write-dominated: bytes = cropped source regions read once + n * ((oh+6)/2) * ((ow+6)/2) * 32 written.
// One thread per output s2d pixel = 2 x 2 image pixels x 3 channels: the four taps of a pixel come from two unaligned
// 8-byte windows per source row (2 x LDG.64 + funnel shift instead of 6 byte loads), 32 contiguous bytes are stored.
// ------------------------------------------------------------------------------------------------------------------
namespace msf {
namespace {

__device__ __forceinline__ uint64_t window8(const unsigned char* p, const unsigned char* end) {
  // 8 bytes starting at p (any alignment), never touching memory at or beyond `end` + 16 (the tail falls back to bytes)
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  if (p + 16 > end) {
    uint64_t v = 0;
    for (int i = 0; i < 8 && p + i < end; ++i) v |= static_cast<uint64_t>(__ldg(p + i)) << (8 * i);
    return v;
  }
  const uint64_t* w = reinterpret_cast<const uint64_t*>(a & ~static_cast<uintptr_t>(7));
  const uint32_t sh = static_cast<uint32_t>(a & 7) * 8;
  const uint64_t lo = __ldg(w), hi = __ldg(w + 1);
  return sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
}

struct CropGeo {
  int H, W, oh, ow, sh, sw;  // source size, view size, s2d grid size
};

constexpr int kS2dRowsPerCta = 2;

template <int ODT>
__global__ void __launch_bounds__(256) view_crops_s2d_kernel(const unsigned char* __restrict__ src, int64_t src_bytes,
                                                             const msf_view_crop* __restrict__ crops, void* __restrict__ out, CropGeo g, int64_t B,
                                                             float a0, float a1, float a2, float b0, float b1, float b2, int* __restrict__ status) {
  const int row_blocks = (g.sh + kS2dRowsPerCta - 1) / kS2dRowsPerCta;
  const int64_t view = blockIdx.x / row_blocks;
  const int rb = static_cast<int>(blockIdx.x % row_blocks);
  msf_view_crop c = crops[view];
  // the reference would raise on an out-of-range crop; flag it and clamp (CTA-uniform)
  if (c.sample < 0 || c.sample >= B || c.y0 < 0 || c.x0 < 0 || c.y1 > g.H || c.x1 > g.W || c.y1 <= c.y0 || c.x1 <= c.x0) {
    if (status && threadIdx.x == 0) atomicOr(status, 1);
    c.sample = min(max(c.sample, 0), static_cast<int>(B - 1));
    c.y0 = min(max(c.y0, 0), g.H - 1); c.x0 = min(max(c.x0, 0), g.W - 1);
    c.y1 = min(max(c.y1, c.y0 + 1), g.H); c.x1 = min(max(c.x1, c.x0 + 1), g.W);
  }
  const int ch = c.y1 - c.y0, cw = c.x1 - c.x0;
  // source coordinates in double: in fp32 the scale's rounding error times the output index reaches 2e-5 pixels, which a
  // 255-level step between neighbouring pixels turns into 1e-4 of a normalised value (one DFMA per coordinate is noise here)
  const double sy = static_cast<double>(ch) / g.oh, sx = static_cast<double>(cw) / g.ow;
  const unsigned char* base = src + ((static_cast<int64_t>(c.sample) * g.H + c.y0) * g.W + c.x0) * 3;
  const unsigned char* end = src + src_bytes;
  const int64_t pitch = static_cast<int64_t>(g.W) * 3;
  const float sc[3] = {a0, a1, a2}, sf[3] = {b0, b1, b2};
  for (int p = threadIdx.x; p < kS2dRowsPerCta * g.sw; p += 256) {
    const int r = p / g.sw, ox = p - r * g.sw, oy = rb * kS2dRowsPerCta + r;
    if (oy >= g.sh) break;
    float v[3][4];  // [channel][dy*2 + dx]
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int y = 2 * oy + (q >> 1) - 3;
      int x = 2 * ox + (q & 1) - 3;
      if (y < 0 || y >= g.oh || x < 0 || x >= g.ow) {  // the stem convolution's zero padding
        v[0][q] = v[1][q] = v[2][q] = 0.f;
        continue;
      }
      if (c.flip) x = g.ow - 1 - x;
      // F.interpolate(align_corners=False) on the crop: source = (o + 0.5) * scale - 0.5, taps clamped at the crop border
      const double fy = fmax(fma(y + 0.5, sy, -0.5), 0.0), fx = fmax(fma(x + 0.5, sx, -0.5), 0.0);
      const int y0 = min(static_cast<int>(fy), ch - 1), x0 = min(static_cast<int>(fx), cw - 1);
      const int y1 = min(y0 + 1, ch - 1);
      const float wy = y0 < ch - 1 ? static_cast<float>(fy - y0) : 0.f, wx = x0 < cw - 1 ? static_cast<float>(fx - x0) : 0.f;
      const bool has_x1 = x0 < cw - 1;
      const uint64_t t0 = window8(base + y0 * pitch + x0 * 3, end), t1 = window8(base + y1 * pitch + x0 * 3, end);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float p00 = static_cast<float>((t0 >> (8 * k)) & 0xff), p10 = static_cast<float>((t1 >> (8 * k)) & 0xff);
        const float p01 = has_x1 ? static_cast<float>((t0 >> (8 * (k + 3))) & 0xff) : p00;
        const float p11 = has_x1 ? static_cast<float>((t1 >> (8 * (k + 3))) & 0xff) : p10;
        const float top = fmaf(p01 - p00, wx, p00), bot = fmaf(p11 - p10, wx, p10);
        v[k][q] = fmaf(fmaf(bot - top, wy, top), sc[k], sf[k]);
      }
    }
    const int64_t e = (view * g.sh + oy) * g.sw + ox;  // s2d pixel index: 16 channels = c*4 + dy*2 + dx (channels 12..15 are zero)
    if constexpr (ODT == MSF_F32) {
      float4* dst = reinterpret_cast<float4*>(static_cast<float*>(out) + e * 16);
      dst[0] = make_float4(v[0][0], v[0][1], v[0][2], v[0][3]);
      dst[1] = make_float4(v[1][0], v[1][1], v[1][2], v[1][3]);
      dst[2] = make_float4(v[2][0], v[2][1], v[2][2], v[2][3]);
      dst[3] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      float f0[8] = {v[0][0], v[0][1], v[0][2], v[0][3], v[1][0], v[1][1], v[1][2], v[1][3]};
      float f1[8] = {v[2][0], v[2][1], v[2][2], v[2][3], 0.f, 0.f, 0.f, 0.f};
      char* dst = static_cast<char*>(out) + e * 32;
      stg_stream(dst, Elem<ODT>::pack(f0));
      stg_stream(dst + 16, Elem<ODT>::pack(f1));
    }
  }
}


// ---- D1b, staged form --------------------------------------------------------------------------------------------
// CTA = one view x kS2dBlockRows rows of the space-to-depth grid.  The source rows those output rows tap are copied
// into shared memory with aligned 128-bit loads (every source byte crosses L1 once per CTA, coalesced), then each thread
// produces whole s2d pixels (2x2 view pixels x 3 channels -> 32 bytes of bf16) from byte reads of shared memory.
// The resampling is exact integer arithmetic: the source coordinate of output o is N/d with N = max((2o+1)*crop - out, 0),
// d = 2*out, so tap = N div d and the weight is the integer remainder rx (ry) out of dx (dy); one output value is
//   ((b00*(dx-rx) + b01*rx)*(dy-ry) + (b10*(dx-rx) + b11*rx)*ry) / (dx*dy)
// -- six integer multiply-adds on the raw bytes, ONE conversion and one FMA (normalise); no double precision, no per-byte
// int->float conversions.  Column taps are tabulated once per CTA, row taps once per staged sub-block.
constexpr int kS2dBlockRows = 8;            // s2d rows per CTA (16 view rows)
constexpr int kS2dStageBytes = 43 * 1024;   // shared-memory budget of the staged source rows (+ 4 KB column table + row table <= 48 KB)
constexpr int kS2dMaxCols = 1024;           // padded view columns (ow + 6) the column table holds (4 KB after the stage)

struct AxisTapI {
  int i0;    // crop-relative first tap (the second is i0 + 1, or i0 again at the crop's last pixel, where rem = 0)
  int rem;   // weight of the second tap, out of 2*out
  bool last;
};

__device__ __forceinline__ AxisTapI axis_tap_int(int o, int out, int crop, float inv_d) {
  const int d = 2 * out;
  const int n = max((2 * o + 1) * crop - out, 0);
  int q = __float2int_rz(__int2float_rz(n) * inv_d);  // estimate of n / d, off by at most one for n < 2^24
  int rem = n - q * d;
  if (rem < 0) { --q; rem += d; } else if (rem >= d) { ++q; rem -= d; }
  AxisTapI t;
  t.last = q >= crop - 1;
  t.i0 = min(q, crop - 1);
  t.rem = t.last ? 0 : rem;
  return t;
}

template <int ODT>
__global__ void __launch_bounds__(256) view_crops_s2d_staged_kernel(const unsigned char* __restrict__ src, int64_t src_bytes,
                                                                    const msf_view_crop* __restrict__ crops, void* __restrict__ out, CropGeo g, int64_t B,
                                                                    float a0, float a1, float a2, float b0, float b1, float b2, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char stage[];
  uint32_t* xtab = reinterpret_cast<uint32_t*>(stage + kS2dStageBytes);  // [ow + 6]  byte offset of tap 0 | rx << 16   (0xffffffff: zero padding)
  __shared__ __align__(16) uint2 ytab[2 * kS2dBlockRows];                           // {stage offset of tap row 0 | tap row 1 << 16, ry}  (ry = 0xffffffff: padding)
  const int row_blocks = (g.sh + kS2dBlockRows - 1) / kS2dBlockRows;
  const int64_t view = blockIdx.x / row_blocks;
  const int rb = static_cast<int>(blockIdx.x % row_blocks);
  msf_view_crop c = crops[view];
  if (c.sample < 0 || c.sample >= B || c.y0 < 0 || c.x0 < 0 || c.y1 > g.H || c.x1 > g.W || c.y1 <= c.y0 || c.x1 <= c.x0) {
    if (status && threadIdx.x == 0) atomicOr(status, 1);
    c.sample = min(max(c.sample, 0), static_cast<int>(B - 1));
    c.y0 = min(max(c.y0, 0), g.H - 1); c.x0 = min(max(c.x0, 0), g.W - 1);
    c.y1 = min(max(c.y1, c.y0 + 1), g.H); c.x1 = min(max(c.x1, c.x0 + 1), g.W);
  }
  const int ch = c.y1 - c.y0, cw = c.x1 - c.x0;
  const int dy = 2 * g.oh, dx = 2 * g.ow;
  const float inv_dy = 1.f / dy, inv_dx = 1.f / dx;
  const int64_t pitch = static_cast<int64_t>(g.W) * 3;
  const unsigned char* base = src + ((static_cast<int64_t>(c.sample) * g.H + c.y0) * g.W + c.x0) * 3;  // crop origin
  const unsigned char* end = src + src_bytes;
  const int row_bytes = cw * 3;
  const int sp = ((row_bytes + 15 + 15) & ~15) + 16;  // staged row pitch: the row + its misalignment + the 3 bytes read past the last pixel
  const int fit = kS2dStageBytes / sp;                // source rows that fit
  const float norm = 1.f / (static_cast<float>(dx) * static_cast<float>(dy));
  const float sc[3] = {a0 * norm, a1 * norm, a2 * norm}, sf[3] = {b0, b1, b2};
  const int sr_end = min((rb + 1) * kS2dBlockRows, g.sh);

  // column taps of the whole view (padded column xp = x + 3)
  for (int xp = threadIdx.x; xp < g.ow + 6; xp += 256) {
    int x = xp - 3;
    uint32_t e = 0xffffffffu;
    if (x >= 0 && x < g.ow) {
      if (c.flip) x = g.ow - 1 - x;
      const AxisTapI t = axis_tap_int(x, g.ow, cw, inv_dx);
      e = static_cast<uint32_t>(t.i0 * 3) | (static_cast<uint32_t>(t.rem) << 16);
    }
    xtab[xp] = e;
  }

  // s2d rows per staged sub-block: view rows [ya, yb] tap at most floor((yb - ya) * ch / oh) + 3 source rows
  const int sub = max(1, min(kS2dBlockRows, ((fit - 3) * g.oh / ch + 1) / 2));
  const int64_t view_row0 = view * g.sh;

  for (int sr0 = rb * kS2dBlockRows; sr0 < sr_end;) {
    const int sr1 = min(sr0 + sub, sr_end);
    const int first = axis_tap_int(min(max(2 * sr0 - 3, 0), g.oh - 1), g.oh, ch, inv_dy).i0;
    const AxisTapI tl = axis_tap_int(min(max(2 * sr1 - 4, 0), g.oh - 1), g.oh, ch, inv_dy);  // last view row of the sub-block
    const int nrows = (tl.last ? tl.i0 : tl.i0 + 1) - first + 1;
    // row taps of the sub-block (padded row yp = y + 3 = 2*sr + dy)
    if (threadIdx.x < 2 * (sr1 - sr0)) {
      const int y = 2 * sr0 + static_cast<int>(threadIdx.x) - 3;
      uint2 e = make_uint2(0u, 0xffffffffu);
      if (y >= 0 && y < g.oh) {
        const AxisTapI t = axis_tap_int(y, g.oh, ch, inv_dy);
        const int r0 = t.i0 - first, r1 = (t.last ? t.i0 : t.i0 + 1) - first;
        const uint32_t m0 = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(base + t.i0 * pitch) & 15);
        const uint32_t m1 = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(base + (first + r1) * pitch) & 15);
        e = make_uint2((r0 * sp + m0) | ((r1 * sp + m1) << 16), static_cast<uint32_t>(t.rem));
      }
      ytab[threadIdx.x] = e;
    }
    // ---- stage source rows [first, last] of the crop: aligned 16-byte chunks, misalignment kept (byte j of row i sits at i*sp + m_i + j)
    {
      const int chunks_max = sp >> 4, total = nrows * chunks_max;
      const float inv_cm = 1.f / chunks_max;
      for (int i0 = threadIdx.x; i0 < total; i0 += 4 * 256) {  // four loads in flight per thread
        uint4 v[4];
        int dst[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * 256;
          dst[u] = -1;
          if (i >= total) continue;
          const int r = __float2int_rz((i + 0.5f) * inv_cm), k = i - r * chunks_max;
          const unsigned char* row = base + (first + r) * pitch;
          const int m = static_cast<int>(reinterpret_cast<uintptr_t>(row) & 15);
          if (k * 16 >= m + row_bytes) continue;
          const unsigned char* gp = row - m + k * 16;
          dst[u] = r * sp + k * 16;
          if (gp + 16 <= end) {
            v[u] = ldg_stream(gp);
          } else {  // the last chunk of the buffer: bytes
            unsigned char tmp[16];
            for (int q = 0; q < 16; ++q) tmp[q] = gp + q < end ? __ldg(gp + q) : 0;
            v[u] = *reinterpret_cast<uint4*>(tmp);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (dst[u] >= 0) *reinterpret_cast<uint4*>(stage + dst[u]) = v[u];
      }
    }
    __syncthreads();
    // ---- produce the s2d pixels of rows [sr0, sr1)
    const int items = (sr1 - sr0) * g.sw;
    const float inv_sw = 1.f / g.sw;
    for (int p = threadIdx.x; p < items; p += 256) {
      const int r = __float2int_rz((p + 0.5f) * inv_sw), ox = p - r * g.sw;
      const uint2 xe = *reinterpret_cast<const uint2*>(xtab + 2 * ox);  // the two columns of this s2d pixel
      const uint4 ye = *reinterpret_cast<const uint4*>(ytab + 2 * r);   // the two rows
      float v[3][4];  // [channel][dy*2 + dx]
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t xw = (q & 1) ? xe.y : xe.x;
        const uint32_t yo = (q >> 1) ? ye.z : ye.x, yr = (q >> 1) ? ye.w : ye.y;
        if (xw == 0xffffffffu || yr == 0xffffffffu) {  // the stem convolution's zero padding
          v[0][q] = v[1][q] = v[2][q] = 0.f;
          continue;
        }
        const int rx = static_cast<int>(xw >> 16), cx = dx - rx, ry = static_cast<int>(yr), cy = dy - ry;
        const unsigned char* s0 = stage + (yo & 0xffffu) + (xw & 0xffffu);
        const unsigned char* s1 = stage + (yo >> 16) + (xw & 0xffffu);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          // the second column tap is read at +3 even at the crop's last pixel, where rx = 0 cancels whatever byte sits there
          const int top = static_cast<int>(s0[k]) * cx + static_cast<int>(s0[k + 3]) * rx;
          const int bot = static_cast<int>(s1[k]) * cx + static_cast<int>(s1[k + 3]) * rx;
          v[k][q] = fmaf(__int2float_rn(top * cy + bot * ry), sc[k], sf[k]);
        }
      }
      const int64_t e = static_cast<int64_t>(view_row0 + sr0 + r) * g.sw + ox;
      if constexpr (ODT == MSF_F32) {
        float4* dst = reinterpret_cast<float4*>(static_cast<float*>(out) + e * 16);
        dst[0] = make_float4(v[0][0], v[0][1], v[0][2], v[0][3]);
        dst[1] = make_float4(v[1][0], v[1][1], v[1][2], v[1][3]);
        dst[2] = make_float4(v[2][0], v[2][1], v[2][2], v[2][3]);
        dst[3] = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        float f0[8] = {v[0][0], v[0][1], v[0][2], v[0][3], v[1][0], v[1][1], v[1][2], v[1][3]};
        float f1[8] = {v[2][0], v[2][1], v[2][2], v[2][3], 0.f, 0.f, 0.f, 0.f};
        char* dst = static_cast<char*>(out) + e * 32;
        stg_stream(dst, Elem<ODT>::pack(f0));
        stg_stream(dst + 16, Elem<ODT>::pack(f1));
      }
    }
    sr0 = sr1;
    if (sr0 < sr_end) __syncthreads();  // the stage and the row table are rewritten
  }
}

}  // namespace
}  // namespace msf

extern "C" int msf_view_crops_s2d(const uint8_t* src, int64_t B, int H, int W, const msf_view_crop* crops, int64_t n_views, int oh, int ow,
                                  const float* mean3, const float* std3, void* out, int out_dtype, int32_t* status_flag, void* stream) {
  MSF_REQUIRE(dtype_ok(out_dtype), MSF_ERR_INVALID, "bad dtype");
  MSF_REQUIRE(B > 0 && H > 0 && W > 0 && oh > 0 && ow > 0 && oh % 2 == 0 && ow % 2 == 0, MSF_ERR_INVALID, "bad sizes (the view size must be even)");
  MSF_REQUIRE(mean3 && std3 && std3[0] > 0.f && std3[1] > 0.f && std3[2] > 0.f, MSF_ERR_INVALID, "mean / std (host, 3 values) invalid");
  if (n_views == 0) return MSF_OK;
  MSF_REQUIRE(src && crops && out && aligned16(out) && n_views > 0, MSF_ERR_INVALID, "NULL or misaligned pointer");
  const CropGeo g{H, W, oh, ow, (oh + 6) / 2, (ow + 6) / 2};
  float a[3], b[3];
  for (int c = 0; c < 3; ++c) { a[c] = 1.f / (255.f * std3[c]); b[c] = -mean3[c] / std3[c]; }
  // algorithmic bytes: ~0.75 of each crop's maximal source area is not known here; count the output (the dominant term) + 3 source bytes
  // per output pixel (a 1:1 resample reads what it writes; the context views read ~4.6x that, 6 % of the launch)
  ProfScope prof(stream, MSF_K_JIGSAW_TILES, static_cast<double>(n_views) * (static_cast<double>(g.sh) * g.sw * 16 * dtype_size(out_dtype) + 3.0 * oh * ow));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // staged form: the source rows of ONE s2d row of the largest possible crop (the whole tile) must fit the stage, and the exact integer
  // coordinates need (2*out + 1) * crop < 2^24
  const int64_t worst_rows = (2 * static_cast<int64_t>(H) + oh - 1) / oh + 3;
  const int64_t worst_pitch = ((static_cast<int64_t>(W) * 3 + 30) & ~static_cast<int64_t>(15)) + 16;
  const bool staged = aligned16(src) && worst_rows * worst_pitch <= kS2dStageBytes && (2 * static_cast<int64_t>(oh) + 1) * H < (1 << 24) &&
                      (2 * static_cast<int64_t>(ow) + 1) * W < (1 << 24) && ow + 6 <= kS2dMaxCols && static_cast<int64_t>(W) * 3 + 3 < 65536 &&
                      255ll * 4 * oh * ow < (1ll << 31);  // the integer numerator of one output value
  if (staged) {
    const int64_t blocks = n_views * ((g.sh + kS2dBlockRows - 1) / kS2dBlockRows);
    MSF_REQUIRE(blocks < (int64_t{1} << 31), MSF_ERR_UNSUPPORTED, "%lld CTAs must be < 2^31", static_cast<long long>(blocks));
    MSF_DISPATCH_DTYPE(out_dtype, (view_crops_s2d_staged_kernel<DT><<<static_cast<unsigned>(blocks), 256, kS2dStageBytes + kS2dMaxCols * 4, st>>>(
                                      src, B * static_cast<int64_t>(H) * W * 3, crops, out, g, B, a[0], a[1], a[2], b[0], b[1], b[2], status_flag)));
    MSF_LAUNCH_OK("view_crops_s2d_staged_kernel");
    return MSF_OK;
  }
  const int64_t blocks = n_views * ((g.sh + kS2dRowsPerCta - 1) / kS2dRowsPerCta);
  MSF_REQUIRE(blocks < (int64_t{1} << 31), MSF_ERR_UNSUPPORTED, "%lld CTAs must be < 2^31", static_cast<long long>(blocks));
  MSF_DISPATCH_DTYPE(out_dtype, (view_crops_s2d_kernel<DT><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
                                    src, B * static_cast<int64_t>(H) * W * 3, crops, out, g, B, a[0], a[1], a[2], b[0], b[1], b[2], status_flag)));
  MSF_LAUNCH_OK("view_crops_s2d_kernel");
  return MSF_OK;
}
