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
  const int64_t blocks = n_views * ((g.sh + kS2dRowsPerCta - 1) / kS2dRowsPerCta);
  MSF_REQUIRE(blocks < (int64_t{1} << 31), MSF_ERR_UNSUPPORTED, "%lld CTAs must be < 2^31", static_cast<long long>(blocks));
  float a[3], b[3];
  for (int c = 0; c < 3; ++c) { a[c] = 1.f / (255.f * std3[c]); b[c] = -mean3[c] / std3[c]; }
  // algorithmic bytes: ~0.75 of each crop's maximal source area is not known here; count the output (the dominant term) + 3 source bytes
  // per output pixel (a 1:1 resample reads what it writes; the context views read ~4.6x that, 6 % of the launch)
  ProfScope prof(stream, MSF_K_JIGSAW_TILES, static_cast<double>(n_views) * (static_cast<double>(g.sh) * g.sw * 16 * dtype_size(out_dtype) + 3.0 * oh * ow));
  MSF_DISPATCH_DTYPE(out_dtype, (view_crops_s2d_kernel<DT><<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
                                    src, B * static_cast<int64_t>(H) * W * 3, crops, out, g, B, a[0], a[1], a[2], b[0], b[1], b[2], status_flag)));
  MSF_LAUNCH_OK("view_crops_s2d_kernel");
  return MSF_OK;
}
