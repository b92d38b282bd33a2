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
