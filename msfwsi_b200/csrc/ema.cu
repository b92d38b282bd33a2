// E1: multi-tensor EMA  teacher <- m*teacher + (1-m)*student  in ONE launch over a device-resident
// tensor table (extension: the reference keeps no momentum encoder, backbone.py:58-65).
//
// HBM-bound: one CTA per 8192-element chunk, 128-bit loads/stores, L1 bypassed.  Algorithmic bytes
// per parameter: e_t (read teacher) + e_s (read student) + e_t (write teacher); 12 B for fp32/fp32.
// Arithmetic matches torch's `t.mul_(m).add_(s, alpha=1-m)` on CUDA: fma(1-m, s, rn(m*t)).
#include "common.cuh"

namespace msf {
namespace {

constexpr int kThreads = 256;
constexpr int kPerThread = MSF_EMA_CHUNK / kThreads;  // 32 elements
static_assert(kPerThread % 8 == 0, "chunk must split into 8-element groups per thread");

template <int DT>
__device__ __forceinline__ void load8(const void* base, int64_t i, bool vec, int64_t n, float* f) {
  if (vec) {
    if constexpr (DT == MSF_F32) {
      const uint4 a = ldg_stream(static_cast<const float*>(base) + i), b = ldg_stream(static_cast<const float*>(base) + i + 4);
      Elem<DT>::unpack(a, f);
      Elem<DT>::unpack(b, f + 4);
    } else {
      Elem<DT>::unpack(ldg_stream(static_cast<const uint16_t*>(base) + i), f);
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = 0.f;
      if (i + k < n) {
        if constexpr (DT == MSF_F32) v = static_cast<const float*>(base)[i + k];
        else if constexpr (DT == MSF_BF16) v = __bfloat162float(static_cast<const __nv_bfloat16*>(base)[i + k]);
        else v = __half2float(static_cast<const __half*>(base)[i + k]);
      }
      f[k] = v;
    }
  }
}

template <int DT>
__device__ __forceinline__ void store8(void* base, int64_t i, bool vec, int64_t n, const float* f) {
  if (vec) {
    if constexpr (DT == MSF_F32) {
      stg_stream(static_cast<float*>(base) + i, Elem<DT>::pack(f));
      stg_stream(static_cast<float*>(base) + i + 4, Elem<DT>::pack(f + 4));
    } else {
      stg_stream(static_cast<uint16_t*>(base) + i, Elem<DT>::pack(f));
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (i + k < n) {
        if constexpr (DT == MSF_F32) static_cast<float*>(base)[i + k] = f[k];
        else if constexpr (DT == MSF_BF16) static_cast<__nv_bfloat16*>(base)[i + k] = __float2bfloat16_rn(f[k]);
        else static_cast<__half*>(base)[i + k] = __float2half_rn(f[k]);
      }
    }
  }
}

template <int TDT, int SDT>
__global__ void __launch_bounds__(kThreads) ema_kernel(const msf_ema_entry* __restrict__ entries,
                                                       const int32_t* __restrict__ prefix, int n_tensors, float m,
                                                       float one_minus_m) {
  // CTA-uniform binary search: tensor t with prefix[t] <= chunk < prefix[t+1]
  const int chunk = blockIdx.x;
  int lo = 0, hi = n_tensors;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(prefix + mid) <= chunk) lo = mid; else hi = mid;
  }
  const msf_ema_entry e = entries[lo];
  const int64_t n = e.numel;
  const int64_t base = static_cast<int64_t>(chunk - __ldg(prefix + lo)) * MSF_EMA_CHUNK;
  const bool aligned = ((reinterpret_cast<uintptr_t>(e.teacher) | reinterpret_cast<uintptr_t>(e.student)) & 15u) == 0;
  float t[kPerThread / 8][8], s[kPerThread / 8][8];
  bool vec[kPerThread / 8];
#pragma unroll
  for (int g = 0; g < kPerThread / 8; ++g) {  // all loads first: 8 x 128-bit requests in flight per thread
    const int64_t i = base + (static_cast<int64_t>(g) * kThreads + threadIdx.x) * 8;
    vec[g] = aligned && (i + 8 <= n);
    if (i < n) {
      load8<TDT>(e.teacher, i, vec[g], n, t[g]);
      load8<SDT>(e.student, i, vec[g], n, s[g]);
    }
  }
#pragma unroll
  for (int g = 0; g < kPerThread / 8; ++g) {
    const int64_t i = base + (static_cast<int64_t>(g) * kThreads + threadIdx.x) * 8;
    if (i < n) {
#pragma unroll
      for (int k = 0; k < 8; ++k) t[g][k] = __fmaf_rn(one_minus_m, s[g][k], __fmul_rn(t[g][k], m));
      store8<TDT>(e.teacher, i, vec[g], n, t[g]);
    }
  }
}

}  // namespace
}  // namespace msf

using namespace msf;

extern "C" int msf_ema_plan(const int64_t* numels, int n_tensors, int32_t* chunk_prefix) {
  MSF_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || (numels && chunk_prefix)), MSF_ERR_INVALID, "bad arguments");
  int64_t acc = 0;
  if (chunk_prefix) chunk_prefix[0] = 0;
  for (int i = 0; i < n_tensors; ++i) {
    MSF_REQUIRE(numels[i] >= 0, MSF_ERR_INVALID, "numels[%d] < 0", i);
    acc += (numels[i] + MSF_EMA_CHUNK - 1) / MSF_EMA_CHUNK;
    MSF_REQUIRE(acc < (1ll << 31), MSF_ERR_UNSUPPORTED, "too many chunks");
    chunk_prefix[i + 1] = static_cast<int32_t>(acc);
  }
  return MSF_OK;
}

extern "C" int msf_ema_multi(const msf_ema_entry* entries, const int32_t* chunk_prefix, int n_tensors, int total_chunks,
                             int teacher_dtype, int student_dtype, float momentum, float one_minus_momentum, void* stream) {
  MSF_REQUIRE(n_tensors >= 0 && total_chunks >= 0, MSF_ERR_INVALID, "negative sizes");
  if (n_tensors == 0 || total_chunks == 0) return MSF_OK;
  MSF_REQUIRE(entries && chunk_prefix, MSF_ERR_INVALID, "NULL table");
  MSF_REQUIRE(dtype_ok(teacher_dtype) && dtype_ok(student_dtype), MSF_ERR_INVALID, "bad dtype");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float m = momentum, om = one_minus_momentum;
  // upper bound on the bytes (the last chunk of every tensor may be partial; numels live on the device)
  ProfScope prof(stream, MSF_K_EMA, static_cast<double>(total_chunks) * MSF_EMA_CHUNK * (2.0 * dtype_size(teacher_dtype) + dtype_size(student_dtype)));
#define MSF_EMA_CASE(T, S)                                                                                   \
  if (teacher_dtype == T && student_dtype == S) {                                                            \
    ema_kernel<T, S><<<total_chunks, kThreads, 0, st>>>(entries, chunk_prefix, n_tensors, m, om);            \
    MSF_LAUNCH_OK("ema_kernel");                                                                             \
    return MSF_OK;                                                                                           \
  }
  MSF_EMA_CASE(MSF_F32, MSF_F32)
  MSF_EMA_CASE(MSF_F32, MSF_BF16)
  MSF_EMA_CASE(MSF_F32, MSF_F16)
  MSF_EMA_CASE(MSF_BF16, MSF_BF16)
  MSF_EMA_CASE(MSF_F16, MSF_F16)
#undef MSF_EMA_CASE
  set_error("unsupported EMA dtype pair teacher=%d student=%d", teacher_dtype, student_dtype);
  return MSF_ERR_UNSUPPORTED;
}
