// S1: space-to-depth re-layout of the encoder input for the ResNet stem convolution (src/models/resnet.py:155, 244:
// Conv2d(3, 64, kernel 7, stride 2, padding 3)).  A 7x7 stride-2 convolution over C_in channels equals a 4x4 stride-1
// convolution over 4*C_in channels of the zero-padded, 2x2 pixel-unshuffled input (the 7x7 kernel zero-padded to 8x8 and
// rearranged the same way).  cuDNN runs that form ~3x faster (3 input channels give its tensor-core kernels nothing to
// tile over; 16 do), so the convolution stays on cuDNN and this kernel only produces the layout it likes:
//   out[n, oy, ox, c*4 + dy*2 + dx] = x[n, c, 2*oy + dy - 3, 2*ox + dx - 3]   (0 outside the image; channels >= 4*C_in are 0)
// out is NHWC with 16 channels, (H+6)/2 x (W+6)/2 pixels; x is addressed through element strides (NCHW or NHWC, any
// float dtype), the cast to the convolution's dtype (bf16 under autocast) is folded in.
// HBM-bound: N*C_in*H*W*e_in read + N*(H+6)/2*(W+6)/2*16*e_out written; one thread per output pixel, 2 x 128-bit stores.
#include "common.cuh"

namespace msf {
namespace {

template <int IDT>
__device__ __forceinline__ float ld_elem(const void* p, int64_t i) {
  if constexpr (IDT == MSF_F32) return __ldg(static_cast<const float*>(p) + i);
  else if constexpr (IDT == MSF_BF16) return __bfloat162float(__ldg(static_cast<const __nv_bfloat16*>(p) + i));
  else return __half2float(__ldg(static_cast<const __half*>(p) + i));
}

// One CTA per output row (n, oy): a single division per CTA, threads walk the row's pixels.
template <int IDT, int ODT>
__global__ void __launch_bounds__(128) stem_s2d_kernel(const void* __restrict__ x, char* __restrict__ out, int cin, int H, int W, int OH,
                                                       int OW, int64_t sn, int64_t sc, int64_t sy, int64_t sx) {
  constexpr int OV = Elem<ODT>::VEC;      // output elements per 16-byte chunk
  constexpr int CHUNKS = 16 / OV;         // 16 channels = 2 chunks (16-bit) or 4 chunks (fp32)
  const int64_t row = blockIdx.x;         // n * OH + oy
  const int oy = static_cast<int>(row % OH);
  const int64_t n = row / OH;
  const int y0 = 2 * oy - 3;
  const bool oky[2] = {y0 >= 0 && y0 < H, y0 + 1 >= 0 && y0 + 1 < H};
  const int64_t base_row = n * sn + y0 * sy;
  char* orow = out + row * OW * (16 * (16 / OV));
  for (int ox = threadIdx.x; ox < OW; ox += 128) {
    float v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0.f;
    const int x0 = 2 * ox - 3;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      if (!oky[dy]) continue;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int xx = x0 + dx;
        if (xx < 0 || xx >= W) continue;
        const int64_t base = base_row + dy * sy + xx * sx;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < cin) v[c * 4 + dy * 2 + dx] = ld_elem<IDT>(x, base + c * sc);
      }
    }
    char* o = orow + static_cast<int64_t>(ox) * (16 * (16 / OV));
#pragma unroll
    for (int k = 0; k < CHUNKS; ++k) stg_stream(o + k * 16, Elem<ODT>::pack(v + k * OV));
  }
}

}  // namespace
}  // namespace msf

using namespace msf;

extern "C" int msf_stem_s2d(const void* x, int64_t N, int C_in, int H, int W, int64_t stride_n, int64_t stride_c, int64_t stride_y,
                            int64_t stride_x, int in_dtype, void* out, int out_dtype, void* stream) {
  MSF_REQUIRE(dtype_ok(in_dtype) && dtype_ok(out_dtype), MSF_ERR_INVALID, "bad dtype");
  MSF_REQUIRE(N >= 0 && C_in >= 1 && C_in <= 4, MSF_ERR_UNSUPPORTED, "C_in=%d: the 16-channel space-to-depth layout holds at most 4 input channels", C_in);
  MSF_REQUIRE(H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, MSF_ERR_UNSUPPORTED, "H=%d, W=%d must be positive and even", H, W);
  if (N == 0) return MSF_OK;
  MSF_REQUIRE(x && out && aligned16(out), MSF_ERR_INVALID, "NULL or misaligned pointer");
  const int OH = (H + 6) / 2, OW = (W + 6) / 2;
  const int64_t total = N * OH * OW;
  const int64_t blocks = N * OH;  // one CTA per output row
  MSF_REQUIRE(blocks < (int64_t{1} << 31), MSF_ERR_UNSUPPORTED, "N*(H+6)/2 = %lld must be < 2^31", static_cast<long long>(blocks));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(stream, MSF_K_STEM_S2D, static_cast<double>(N) * C_in * H * W * dtype_size(in_dtype) + static_cast<double>(total) * 16 * dtype_size(out_dtype));
#define MSF_S2D(I, O)                                                                                                               \
  if (in_dtype == I && out_dtype == O) {                                                                                            \
    stem_s2d_kernel<I, O><<<static_cast<unsigned>(blocks), 128, 0, st>>>(x, static_cast<char*>(out), C_in, H, W, OH, OW,            \
                                                                         stride_n, stride_c, stride_y, stride_x);                 \
    MSF_LAUNCH_OK("stem_s2d_kernel");                                                                                               \
    return MSF_OK;                                                                                                                  \
  }
  MSF_S2D(MSF_F32, MSF_F32) MSF_S2D(MSF_F32, MSF_BF16) MSF_S2D(MSF_F32, MSF_F16)
  MSF_S2D(MSF_BF16, MSF_BF16) MSF_S2D(MSF_BF16, MSF_F32) MSF_S2D(MSF_F16, MSF_F16) MSF_S2D(MSF_F16, MSF_F32)
#undef MSF_S2D
  set_error("unsupported dtype pair in=%d out=%d", in_dtype, out_dtype);
  return MSF_ERR_UNSUPPORTED;
}
