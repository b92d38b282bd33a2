// O1: multi-tensor Adam step of the reference's optimizer (tools/ssl_train.py:303-309: torch.optim.Adam over three
// learning-rate groups context_/target_/inter_, 264 tensors, 123.6 M parameters) with the GradScaler work
// (tools/ssl_train.py:472-474: unscale + non-finite check + skipped step) and an optional EMA teacher update folded in.
//
//   msf_grad_check_multi   found_inf |= any(!isfinite(grad))                         4 B/parameter read
//   msf_adam_multi         g = grad * inv_scale; m = lerp(m, g, 1-b1); v = b2*v + (1-b2)*g*g;
//                          p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps);  [teacher = mom*teacher + (1-mom)*p]
//                          skipped entirely when *found_inf != 0 (what GradScaler.step does)
// HBM-bound: one CTA per 4096-element chunk of a device-resident tensor table (same plan as E1), 16 x 128-bit loads in
// flight per thread.  Algorithmic bytes per parameter: 16 read (p, g, m, v fp32) + 12 written = 28 B (+8 with EMA,
// +2 where a 16-bit shadow of the parameter -- the GEMM operand of the head Linears -- is rewritten in the same pass).
#include "common.cuh"

namespace msf {
namespace {

constexpr int kThreads = 256;
constexpr int kPer = MSF_ADAM_CHUNK / kThreads;  // 16 elements = 4 x float4 per array
static_assert(kPer % 4 == 0, "chunk must split into float4 groups");

__device__ __forceinline__ int find_tensor(const int32_t* __restrict__ prefix, int n_tensors, int chunk) {
  int lo = 0, hi = n_tensors;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(prefix + mid) <= chunk) lo = mid; else hi = mid;
  }
  return lo;
}

template <int GDT>
__device__ __forceinline__ void load_grad4(const void* g, int64_t i, int64_t n, bool vec, float* f) {
  if (vec) {
    if constexpr (GDT == MSF_F32) {
      Elem<MSF_F32>::unpack(ldg_stream(static_cast<const float*>(g) + i), f);
    } else {  // 4 x 16-bit = 8 bytes
      const uint2 w = __ldg(reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(g) + i));
      if constexpr (GDT == MSF_BF16) {
        f[0] = __uint_as_float(w.x << 16); f[1] = __uint_as_float(w.x & 0xffff0000u);
        f[2] = __uint_as_float(w.y << 16); f[3] = __uint_as_float(w.y & 0xffff0000u);
      } else {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&w.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
        f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v = 0.f;
      if (i + k < n) {
        if constexpr (GDT == MSF_F32) v = static_cast<const float*>(g)[i + k];
        else if constexpr (GDT == MSF_BF16) v = __bfloat162float(static_cast<const __nv_bfloat16*>(g)[i + k]);
        else v = __half2float(static_cast<const __half*>(g)[i + k]);
      }
      f[k] = v;
    }
  }
}
__device__ __forceinline__ void load4(const float* p, int64_t i, int64_t n, bool vec, float* f) {
  if (vec) {
    Elem<MSF_F32>::unpack(ldg_stream(p + i), f);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) f[k] = i + k < n ? p[i + k] : 0.f;
  }
}
__device__ __forceinline__ void store4(float* p, int64_t i, int64_t n, bool vec, const float* f) {
  if (vec) {
    stg_stream(p + i, Elem<MSF_F32>::pack(f));
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (i + k < n) p[i + k] = f[k];
  }
}

template <int GDT>
__global__ void __launch_bounds__(kThreads) grad_check_kernel(const msf_adam_entry* __restrict__ entries, const int32_t* __restrict__ prefix,
                                                              int n_tensors, float* __restrict__ found_inf) {
  const int chunk = blockIdx.x;
  const int t = find_tensor(prefix, n_tensors, chunk);
  const msf_adam_entry e = entries[t];
  const int64_t base = static_cast<int64_t>(chunk - __ldg(prefix + t)) * MSF_ADAM_CHUNK;
  const bool aligned = (reinterpret_cast<uintptr_t>(e.grad) & 15u) == 0;
  bool bad = false;
#pragma unroll
  for (int g = 0; g < kPer / 4; ++g) {
    const int64_t i = base + (static_cast<int64_t>(g) * kThreads + threadIdx.x) * 4;
    if (i >= e.numel) continue;
    float f[4];
    load_grad4<GDT>(e.grad, i, e.numel, aligned && i + 4 <= e.numel, f);
#pragma unroll
    for (int k = 0; k < 4; ++k) bad |= !isfinite(f[k]);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) *found_inf = 1.f;
}

template <int GDT, bool EMA>
__global__ void __launch_bounds__(kThreads) adam_kernel(const msf_adam_entry* __restrict__ entries, const int32_t* __restrict__ prefix,
                                                        int n_tensors, const float* __restrict__ lr, double beta1d, double beta2d, float eps,
                                                        float weight_decay, const float* __restrict__ step, const float* __restrict__ inv_scale,
                                                        const float* __restrict__ found_inf, float ema_m, float ema_one_minus_m) {
  // 1 - beta is formed in double and then narrowed, as torch does with its Python-float hyper-parameters
  // (1.f - 0.999f differs from float(1 - 0.999) by 5e-5 relative)
  const float beta2 = static_cast<float>(beta2d);
  const float omb1 = static_cast<float>(1.0 - beta1d), omb2 = static_cast<float>(1.0 - beta2d);
  if (found_inf && *found_inf != 0.f) return;  // GradScaler: skip the whole step
  const int chunk = blockIdx.x;
  const int t = find_tensor(prefix, n_tensors, chunk);
  const msf_adam_entry e = entries[t];
  const int64_t n = e.numel;
  const int64_t base = static_cast<int64_t>(chunk - __ldg(prefix + t)) * MSF_ADAM_CHUNK;
  const bool aligned = ((reinterpret_cast<uintptr_t>(e.param) | reinterpret_cast<uintptr_t>(e.grad) | reinterpret_cast<uintptr_t>(e.exp_avg) |
                         reinterpret_cast<uintptr_t>(e.exp_avg_sq) | reinterpret_cast<uintptr_t>(e.ema)) & 15u) == 0;
  // bias corrections in double, like the Python side of torch.optim.Adam
  const double st = static_cast<double>(*step);
  const double bc1 = 1.0 - pow(beta1d, st), bc2 = 1.0 - pow(beta2d, st);
  const float step_size = static_cast<float>(static_cast<double>(__ldg(lr + e.group)) / bc1);
  const float bc2_sqrt = static_cast<float>(sqrt(bc2));
  const float gscale = inv_scale ? *inv_scale : 1.f;
  float p[kPer / 4][4], g[kPer / 4][4], m[kPer / 4][4], v[kPer / 4][4], tch[kPer / 4][4];
  bool vec[kPer / 4];
#pragma unroll
  for (int q = 0; q < kPer / 4; ++q) {  // all loads first
    const int64_t i = base + (static_cast<int64_t>(q) * kThreads + threadIdx.x) * 4;
    vec[q] = aligned && i + 4 <= n;
    if (i < n) {
      load4(static_cast<const float*>(e.param), i, n, vec[q], p[q]);
      load_grad4<GDT>(e.grad, i, n, vec[q], g[q]);
      load4(static_cast<const float*>(e.exp_avg), i, n, vec[q], m[q]);
      load4(static_cast<const float*>(e.exp_avg_sq), i, n, vec[q], v[q]);
      if (EMA) load4(static_cast<const float*>(e.ema), i, n, vec[q], tch[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < kPer / 4; ++q) {
    const int64_t i = base + (static_cast<int64_t>(q) * kThreads + threadIdx.x) * 4;
    if (i >= n) continue;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gr = g[q][k] * gscale;
      if (weight_decay != 0.f) gr = fmaf(weight_decay, p[q][k], gr);
      m[q][k] = fmaf(omb1, gr - m[q][k], m[q][k]);                          // lerp(m, g, 1 - beta1)
      v[q][k] = fmaf(omb2, gr * gr, beta2 * v[q][k]);
      const float denom = sqrtf(v[q][k]) / bc2_sqrt + eps;
      p[q][k] -= step_size * (m[q][k] / denom);
      if (EMA) tch[q][k] = __fmaf_rn(ema_one_minus_m, p[q][k], __fmul_rn(tch[q][k], ema_m));
    }
    store4(static_cast<float*>(e.param), i, n, vec[q], p[q]);
    store4(static_cast<float*>(e.exp_avg), i, n, vec[q], m[q]);
    store4(static_cast<float*>(e.exp_avg_sq), i, n, vec[q], v[q]);
    if (EMA) store4(static_cast<float*>(e.ema), i, n, vec[q], tch[q]);
    if (e.shadow) {  // 16-bit operand copy of the stepped parameter (tensor-uniform branch)
      uint16_t h[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (e.shadow_dtype == MSF_F16) { const __half t = __float2half_rn(p[q][k]); h[k] = *reinterpret_cast<const uint16_t*>(&t); }
        else { const __nv_bfloat16 t = __float2bfloat16_rn(p[q][k]); h[k] = *reinterpret_cast<const uint16_t*>(&t); }
      }
      uint16_t* dst = static_cast<uint16_t*>(e.shadow) + i;
      if (vec[q] && (reinterpret_cast<uintptr_t>(dst) & 7u) == 0) {
        *reinterpret_cast<uint2*>(dst) = make_uint2(h[0] | (static_cast<uint32_t>(h[1]) << 16), h[2] | (static_cast<uint32_t>(h[3]) << 16));
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (i + k < n) dst[k] = h[k];
      }
    }
  }
}

}  // namespace
}  // namespace msf

using namespace msf;

extern "C" int msf_adam_plan(const int64_t* numels, int n_tensors, int32_t* chunk_prefix) {
  MSF_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || (numels && chunk_prefix)), MSF_ERR_INVALID, "bad arguments");
  int64_t acc = 0;
  if (chunk_prefix) chunk_prefix[0] = 0;
  for (int i = 0; i < n_tensors; ++i) {
    MSF_REQUIRE(numels[i] >= 0, MSF_ERR_INVALID, "numels[%d] < 0", i);
    acc += (numels[i] + MSF_ADAM_CHUNK - 1) / MSF_ADAM_CHUNK;
    MSF_REQUIRE(acc < (1ll << 31), MSF_ERR_UNSUPPORTED, "too many chunks");
    chunk_prefix[i + 1] = static_cast<int32_t>(acc);
  }
  return MSF_OK;
}

extern "C" int msf_grad_check_multi(const msf_adam_entry* entries, const int32_t* chunk_prefix, int n_tensors, int total_chunks,
                                    int grad_dtype, float* found_inf, void* stream) {
  MSF_REQUIRE(n_tensors >= 0 && total_chunks >= 0, MSF_ERR_INVALID, "negative sizes");
  if (n_tensors == 0 || total_chunks == 0) return MSF_OK;
  MSF_REQUIRE(entries && chunk_prefix && found_inf && dtype_ok(grad_dtype), MSF_ERR_INVALID, "bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(stream, MSF_K_GRAD_CHECK, static_cast<double>(total_chunks) * MSF_ADAM_CHUNK * dtype_size(grad_dtype));
  MSF_DISPATCH_DTYPE(grad_dtype, (grad_check_kernel<DT><<<total_chunks, kThreads, 0, st>>>(entries, chunk_prefix, n_tensors, found_inf)));
  MSF_LAUNCH_OK("grad_check_kernel");
  return MSF_OK;
}

extern "C" int msf_adam_multi(const msf_adam_entry* entries, const int32_t* chunk_prefix, int n_tensors, int total_chunks, int grad_dtype,
                              const float* lr, double beta1, double beta2, double eps, double weight_decay, const float* step,
                              const float* inv_scale, const float* found_inf, int with_ema, float ema_momentum,
                              float ema_one_minus_momentum, void* stream) {
  MSF_REQUIRE(n_tensors >= 0 && total_chunks >= 0, MSF_ERR_INVALID, "negative sizes");
  if (n_tensors == 0 || total_chunks == 0) return MSF_OK;
  MSF_REQUIRE(entries && chunk_prefix && lr && step && dtype_ok(grad_dtype), MSF_ERR_INVALID, "bad arguments");
  MSF_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, MSF_ERR_INVALID, "bad hyper-parameters");
  const float epsf = static_cast<float>(eps), wdf = static_cast<float>(weight_decay);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // upper bound on the bytes (last chunk of a tensor may be partial): p, m, v read + written, grad read (+ teacher)
  ProfScope prof(stream, MSF_K_ADAM, static_cast<double>(total_chunks) * MSF_ADAM_CHUNK * (24.0 + dtype_size(grad_dtype) + (with_ema ? 8.0 : 0.0)));
  if (with_ema) {
    MSF_DISPATCH_DTYPE(grad_dtype, (adam_kernel<DT, true><<<total_chunks, kThreads, 0, st>>>(entries, chunk_prefix, n_tensors, lr, beta1, beta2, epsf, wdf,
                                                                                           step, inv_scale, found_inf, ema_momentum, ema_one_minus_momentum)));
  } else {
    MSF_DISPATCH_DTYPE(grad_dtype, (adam_kernel<DT, false><<<total_chunks, kThreads, 0, st>>>(entries, chunk_prefix, n_tensors, lr, beta1, beta2, epsf, wdf,
                                                                                            step, inv_scale, found_inf, 0.f, 0.f)));
  }
  MSF_LAUNCH_OK("adam_kernel");
  return MSF_OK;
}
