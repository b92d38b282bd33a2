// Write-bandwidth probe (development tool): which store flavour fills HBM fastest on B200?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw store_bw.cu && ./store_bw
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__device__ __forceinline__ void st16(void* p, uint4 v) {
  if (MODE == 0) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  if (MODE == 1) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  if (MODE == 2) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  if (MODE == 3) asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// grid-stride, U stores per thread per iteration
template <int MODE, int U>
__global__ void __launch_bounds__(256) fill_stride(char* out, int64_t chunks, uint32_t val) {
  const int64_t stride = (int64_t)gridDim.x * 256;
  const uint4 v = make_uint4(val, val, val, val);
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < chunks; e += U * stride) {
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (e + u * stride < chunks) st16<MODE>(out + (e + u * stride) * 16, v);
  }
}

// 256-bit stores
__global__ void __launch_bounds__(256) fill_v8(char* out, int64_t chunks32, uint32_t val) {
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < chunks32; e += stride) {
    asm volatile("st.global.v8.u32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(out + e * 32), "r"(val) : "memory");
  }
}

// TMA bulk stores: each CTA fills a 16 KB shared buffer once, then issues bulk copies shared -> global
__global__ void __launch_bounds__(128) fill_bulk(char* out, int64_t blocks16k, uint32_t val) {
  extern __shared__ __align__(128) char sm[];
  for (int i = threadIdx.x; i < 16384 / 16; i += 128) reinterpret_cast<uint4*>(sm)[i] = make_uint4(val, val, val, val);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm);
    int pending = 0;
    for (int64_t b = blockIdx.x; b < blocks16k; b += gridDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + b * 16384), "r"(saddr), "r"(16384) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (++pending >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); pending = 4; }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// read-only and copy for reference
__global__ void __launch_bounds__(256) read_only(const char* in, int64_t chunks, uint32_t* sink) {
  const int64_t stride = (int64_t)gridDim.x * 256;
  uint32_t acc = 0;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < chunks; e += 4 * stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (e + u * stride < chunks)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(in + (e + u * stride) * 16));
      else v[u] = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < 4; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

// read-only variants: U 128-bit loads in flight per thread, or 256-bit loads
template <int U>
__global__ void __launch_bounds__(256) read_u(const char* in, int64_t chunks, uint32_t* sink) {
  const int64_t stride = (int64_t)gridDim.x * 256;
  uint32_t acc = 0;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < chunks; e += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = make_uint4(0, 0, 0, 0);
      if (e + u * stride < chunks)
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(in + (e + u * stride) * 16));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}
template <int U>
__global__ void __launch_bounds__(256) read_v8(const char* in, int64_t chunks32, uint32_t* sink) {
  const int64_t stride = (int64_t)gridDim.x * 256;
  uint32_t acc = 0;
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < chunks32; e += U * stride) {
    uint32_t r[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int k = 0; k < 8; ++k) r[u][k] = 0;
      if (e + u * stride < chunks32)
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]), "=r"(r[u][4]), "=r"(r[u][5]), "=r"(r[u][6]), "=r"(r[u][7])
                     : "l"(in + (e + u * stride) * 32));
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc ^= r[u][k];
  }
  if (acc == 0x12345678u) *sink = acc;
}

int main() {
  const int64_t bytes = 8ll << 30;
  char *a, *b;
  uint32_t* sink;
  CK(cudaMalloc(&a, bytes));
  CK(cudaMalloc(&b, bytes));
  CK(cudaMalloc(&sink, 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto report = [&](const char* name, float ms, double nbytes) { printf("%-44s %8.3f ms  %7.0f GB/s\n", name, ms, nbytes / ms / 1e6); fflush(stdout); };
#define TIME(name, nbytes, ...)                                \
  {                                                            \
    for (int w = 0; w < 2; ++w) { __VA_ARGS__; }               \
    float best = 1e30f;                                        \
    for (int it = 0; it < 5; ++it) {                           \
      cudaEventRecord(e0); __VA_ARGS__; cudaEventRecord(e1);   \
      cudaEventSynchronize(e1);                                \
      float ms; cudaEventElapsedTime(&ms, e0, e1);             \
      if (ms < best) best = ms;                                \
    }                                                          \
    CK(cudaGetLastError());                                    \
    report(name, best, nbytes);                                \
  }
  const int64_t chunks = bytes / 16;
  for (int g : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
    char nm[96];
    snprintf(nm, 96, "st.v4 default      U=4 grid=%d", g); TIME(nm, (double)bytes, (fill_stride<0, 4><<<g, 256>>>(a, chunks, 1)));
    snprintf(nm, 96, "st.v4 L1::no_alloc U=4 grid=%d", g); TIME(nm, (double)bytes, (fill_stride<1, 4><<<g, 256>>>(a, chunks, 2)));
    snprintf(nm, 96, "st.v4 .cs          U=4 grid=%d", g); TIME(nm, (double)bytes, (fill_stride<2, 4><<<g, 256>>>(a, chunks, 3)));
    snprintf(nm, 96, "st.v4 .wt          U=4 grid=%d", g); TIME(nm, (double)bytes, (fill_stride<3, 4><<<g, 256>>>(a, chunks, 4)));
    snprintf(nm, 96, "st.v4 default      U=1 grid=%d", g); TIME(nm, (double)bytes, (fill_stride<0, 1><<<g, 256>>>(a, chunks, 5)));
    snprintf(nm, 96, "st.v8 (256-bit)        grid=%d", g); TIME(nm, (double)bytes, (fill_v8<<<g, 256>>>(a, bytes / 32, 6)));
  }
  cudaFuncSetAttribute(fill_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  for (int g : {148, 148 * 2, 148 * 4, 148 * 8}) {
    char nm[96];
    snprintf(nm, 96, "TMA bulk store 16 KB   grid=%d", g); TIME(nm, (double)bytes, (fill_bulk<<<g, 128, 16384>>>(a, bytes / 16384, 7)));
  }
  TIME("cudaMemsetAsync", (double)bytes, cudaMemsetAsync(a, 0, bytes));
  TIME("cudaMemcpyAsync D2D (read+write bytes)", 2.0 * bytes, cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice));
  TIME("read-only ld.nc.v4 U=4 grid=148*16", (double)bytes, (read_only<<<148 * 16, 256>>>(a, chunks, sink)));
  TIME("read-only ld.nc.v4 U=4 grid=148*8", (double)bytes, (read_only<<<148 * 8, 256>>>(a, chunks, sink)));
  for (int g : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
    char nm[96];
    snprintf(nm, 96, "read ld.nc.v4 U=4  grid=%d", g); TIME(nm, (double)bytes, (read_u<4><<<g, 256>>>(a, chunks, sink)));
    snprintf(nm, 96, "read ld.nc.v4 U=8  grid=%d", g); TIME(nm, (double)bytes, (read_u<8><<<g, 256>>>(a, chunks, sink)));
    snprintf(nm, 96, "read ld.nc.v4 U=16 grid=%d", g); TIME(nm, (double)bytes, (read_u<16><<<g, 256>>>(a, chunks, sink)));
    snprintf(nm, 96, "read ld.nc.v8 U=2  grid=%d", g); TIME(nm, (double)bytes, (read_v8<2><<<g, 256>>>(a, bytes / 32, sink)));
    snprintf(nm, 96, "read ld.nc.v8 U=4  grid=%d", g); TIME(nm, (double)bytes, (read_v8<4><<<g, 256>>>(a, bytes / 32, sink)));
    snprintf(nm, 96, "read ld.nc.v8 U=8  grid=%d", g); TIME(nm, (double)bytes, (read_v8<8><<<g, 256>>>(a, bytes / 32, sink)));
  }
  return 0;
}
