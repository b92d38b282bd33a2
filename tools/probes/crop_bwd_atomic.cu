// Development probe: the scatter-form crop backward this repo used in round 1 (one thread per output pixel, four float
// atomicAdds into grad_feat) timed at the two bench shapes, as the baseline the deterministic gather-form kernel
// (csrc/crop_resample.cu) is compared with.  Not part of the library.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o crop_bwd_atomic crop_bwd_atomic.cu && ./crop_bwd_atomic
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

struct Geo { int64_t B; int C, H, W, K, oh, ow; };

__device__ __forceinline__ void taps(float lo, float len, int o, int osz, int limit, int& i0, int& i1, float& w) {
  const float s = fmaxf((o + 0.5f) * (len / osz) - 0.5f, 0.f);
  const int imax = max(static_cast<int>(ceilf(len)) - 1, 0);
  const int f = min(static_cast<int>(floorf(s)), imax);
  const int f1 = min(f + 1, imax);
  w = f < imax ? s - floorf(s) : 0.f;
  const int off = static_cast<int>(floorf(lo));
  i0 = min(max(off + f, 0), limit - 1);
  i1 = min(max(off + f1, 0), limit - 1);
}

__global__ void __launch_bounds__(256) scatter_kernel(const __nv_bfloat16* __restrict__ gout, const float4* __restrict__ boxes, float* __restrict__ gfeat, Geo g, int64_t total) {
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t r = t;
    const int ox = static_cast<int>(r % g.ow); r /= g.ow;
    const int oy = static_cast<int>(r % g.oh); r /= g.oh;
    const int c = static_cast<int>(r % g.C); r /= g.C;
    const int k = static_cast<int>(r % g.K);
    const int64_t b = r / g.K;
    const float4 bx = boxes[b * g.K + k];
    int y0, y1, x0, x1;
    float wy, wx;
    taps(bx.x, bx.z - bx.x, oy, g.oh, g.H, y0, y1, wy);
    taps(bx.y, bx.w - bx.y, ox, g.ow, g.W, x0, x1, wx);
    const float go = __bfloat162float(gout[t]);
    float* plane = gfeat + (b * g.C + c) * static_cast<int64_t>(g.H) * g.W;
    atomicAdd(plane + static_cast<int64_t>(y0) * g.W + x0, go * (1.f - wy) * (1.f - wx));
    if (wx != 0.f) atomicAdd(plane + static_cast<int64_t>(y0) * g.W + x1, go * (1.f - wy) * wx);
    if (wy != 0.f) {
      atomicAdd(plane + static_cast<int64_t>(y1) * g.W + x0, go * wy * (1.f - wx));
      if (wx != 0.f) atomicAdd(plane + static_cast<int64_t>(y1) * g.W + x1, go * wy * wx);
    }
  }
}

int main() {
  const int shapes[2][6] = {{64, 128, 128, 128, 128, 128}, {256, 128, 128, 128, 32, 32}};
  for (auto& s : shapes) {
    Geo g{s[0], s[1], s[2], s[3], 16, s[4], s[5]};
    const int64_t total = g.B * g.K * static_cast<int64_t>(g.C) * g.oh * g.ow, src = g.B * static_cast<int64_t>(g.C) * g.H * g.W;
    __nv_bfloat16* go;
    float* gf;
    float4* bx;
    CK(cudaMalloc(&go, total * 2));
    CK(cudaMalloc(&gf, src * 4));
    CK(cudaMalloc(&bx, g.B * g.K * 16));
    CK(cudaMemset(go, 0x3c, total * 2));
    float4* h = new float4[g.B * g.K];
    for (int64_t b = 0; b < g.B; ++b)
      for (int k = 0; k < 16; ++k) h[b * 16 + k] = make_float4((k / 4) * g.H / 4.f, (k % 4) * g.W / 4.f, (k / 4 + 1) * g.H / 4.f, (k % 4 + 1) * g.W / 4.f);
    CK(cudaMemcpy(bx, h, g.B * g.K * 16, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 6; ++it) {
      cudaEventRecord(e0);
      CK(cudaMemsetAsync(gf, 0, src * 4));  // the scatter form needs a zero-filled destination: part of its cost
      scatter_kernel<<<148 * 16, 256>>>(go, bx, gf, g, total);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (it) best = fminf(best, ms);
    }
    printf("scatter + memset, oh=%d: %.3f ms  (%.0f GB/s of the algorithmic bytes)\n", g.oh, best, (total * 2.0 + src * 4.0) / best / 1e6);
    cudaFree(go); cudaFree(gf); cudaFree(bx);
    delete[] h;
  }
  return 0;
}
