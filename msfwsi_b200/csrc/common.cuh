// Shared helpers for the msfwsi_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/msfwsi_b200.h"

namespace msf {

void set_error(const char* fmt, ...);  // thread-local message behind msf_last_error()

// Opt-in launch profiler (msf_prof_begin / msf_prof_end): records a CUDA event pair on `stream` around the scope when
// profiling is on, otherwise costs one branch.  `work` = algorithmic bytes (HBM-bound kernels) or FLOP (tensor-bound).
class ProfScope {
 public:
  ProfScope(void* stream, int kernel, double work);
  ~ProfScope();
  ProfScope(const ProfScope&) = delete;
  ProfScope& operator=(const ProfScope&) = delete;

 private:
  int slot_;
  void* stream_;
};

#define MSF_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::msf::set_error(__VA_ARGS__);   \
      return (code);                   \
    }                                  \
  } while (0)

#define MSF_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      ::msf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return MSF_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define MSF_LAUNCH_OK(name)                                                        \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      ::msf::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__)); \
      return MSF_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t dtype_size(int dt) { return dt == MSF_F32 ? 4 : 2; }
inline bool dtype_ok(int dt) { return dt == MSF_F32 || dt == MSF_BF16 || dt == MSF_F16; }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- 128-bit global access --------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream(const void* p) {  // read-once data: bypass L1
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_keep(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// ---- dtype traits: a 16-byte chunk holds VEC elements -------------------------------------
template <int DT>
struct Elem;
template <>
struct Elem<MSF_F32> {
  static constexpr int VEC = 4;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
  }
};
template <>
struct Elem<MSF_BF16> {
  static constexpr int VEC = 8;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bf16 -> fp32 is a 16-bit shift
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <>
struct Elem<MSF_F16> {
  static constexpr int VEC = 8;
  __device__ static __forceinline__ void unpack(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  }
  __device__ static __forceinline__ uint4 pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

#define MSF_DISPATCH_DTYPE(dt, ...)                     \
  switch (dt) {                                         \
    case MSF_F32: { constexpr int DT = MSF_F32; __VA_ARGS__; } break;   \
    case MSF_BF16: { constexpr int DT = MSF_BF16; __VA_ARGS__; } break; \
    default: { constexpr int DT = MSF_F16; __VA_ARGS__; } break;        \
  }

// sub-warp sum over `lanes` (power of two <= 32) consecutive lanes
__device__ __forceinline__ float group_sum(float v, int lanes) {
  for (int o = lanes >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace msf
