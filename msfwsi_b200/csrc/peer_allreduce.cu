// C1: latency-bound all-reduce of a small fp64 vector over NVLink peer memory (one kernel, no NCCL launch).
//
// Every batch-norm layer exchanges a {sum, sum of squares, count} vector of 2C+1 doubles per direction
// (SyncBatchNorm, tools/ssl_train.py:160): 352 tiny collectives per step, each a ~30 us NCCL ring kernel plus two
// cross-stream hops.  Here each rank owns a SYMMETRIC workspace (same layout on every GPU, every rank's copy mapped into
// every process: torch.distributed._symmetric_memory) and one single-CTA kernel does the whole exchange:
//   1. PUSH the local vector into slot [parity][my rank] of EVERY rank's workspace (NVLink stores: fire and forget),
//   2. __threadfence_system(), then store `seq` into flags[my_rank] of every rank's workspace,
//   3. spin (bounded by the caller's timeout, ld.acquire.sys) until my flags[r] >= seq for all r,
//   4. sum the slots of all ranks from LOCAL memory in rank order -- fixed order, so all ranks get bit-identical results.
// One NVLink one-way trip (data, then flag) per exchange; the first version pulled the peers' slots after the flag, i.e. a
// second round trip per peer, issued one after the other.
// Two parities suffice: a rank enters call k+2 only after passing the wait of call k+1, which needs every peer's flag
// k+1, which a peer publishes only after it has finished reading the slots of call k.
// Workspace layout: flags[MSF_PEER_MAX_WORLD] (uint64) | pad to 256 B | slots [2 parities][MSF_PEER_MAX_WORLD][capacity] doubles.
#include "common.cuh"

namespace msf {
namespace {

constexpr int kThreads = 1024;
constexpr size_t kFlagBytes = 256;

__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(kThreads) peer_allreduce_kernel(double* __restrict__ vec, int n, char* const* __restrict__ peers, int world,
                                                                  int rank, uint64_t seq, size_t capacity, uint64_t timeout_ns) {
  char* mine = peers[rank];
  const size_t par_off = kFlagBytes + (seq & 1) * MSF_PEER_MAX_WORLD * capacity * sizeof(double);
  for (int i = threadIdx.x; i < n; i += kThreads) {
    const double v = vec[i];
    for (int r = 0; r < world; ++r) reinterpret_cast<double*>(peers[r] + par_off)[static_cast<size_t>(rank) * capacity + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world)  // publish: my flag in every rank's workspace
    st_release_sys(reinterpret_cast<uint64_t*>(peers[threadIdx.x]) + rank, seq);
  if (threadIdx.x < world) {  // wait for everybody (flags only grow)
    const uint64_t* flag = reinterpret_cast<const uint64_t*>(mine) + threadIdx.x;
    const uint64_t t0 = globaltimer_ns();
    while (ld_acquire_sys(flag) < seq) {
      if (globaltimer_ns() - t0 > timeout_ns) {
        printf("msfwsi_b200: peer all-reduce timed out (rank %d waits for rank %d, seq %llu)\n", rank, static_cast<int>(threadIdx.x),
               static_cast<unsigned long long>(seq));
        __trap();
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
  const double* slots = reinterpret_cast<const double*>(mine + par_off);
  for (int i = threadIdx.x; i < n; i += kThreads) {
    double acc = 0.0;
    for (int r0 = 0; r0 < world; r0 += 8) {  // eight loads in flight, added in rank order
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t[u] = r0 + u < world ? ld_volatile_f64(slots + static_cast<size_t>(r0 + u) * capacity + i) : 0.0;
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (r0 + u < world) acc += t[u];
    }
    vec[i] = acc;
  }
}

}  // namespace
}  // namespace msf

using namespace msf;

extern "C" size_t msf_peer_workspace_bytes(int64_t capacity_doubles) {
  if (capacity_doubles <= 0) return 0;
  return kFlagBytes + 2 * static_cast<size_t>(MSF_PEER_MAX_WORLD) * static_cast<size_t>(capacity_doubles) * sizeof(double);
}

extern "C" int msf_peer_allreduce_f64(double* vec, int n, void* const* peers, int world, int rank, uint64_t seq, int64_t capacity_doubles,
                                      int timeout_ms, void* stream) {
  MSF_REQUIRE(vec && peers && n > 0, MSF_ERR_INVALID, "bad arguments");
  MSF_REQUIRE(world >= 1 && world <= MSF_PEER_MAX_WORLD && rank >= 0 && rank < world, MSF_ERR_INVALID, "world=%d rank=%d out of range", world, rank);
  MSF_REQUIRE(n <= capacity_doubles, MSF_ERR_WORKSPACE, "vector of %d doubles exceeds the workspace capacity %lld", n,
              static_cast<long long>(capacity_doubles));
  MSF_REQUIRE(seq > 0, MSF_ERR_INVALID, "sequence numbers start at 1");
  MSF_REQUIRE(timeout_ms > 0, MSF_ERR_INVALID, "timeout_ms must be positive");
  ProfScope prof(stream, MSF_K_PEER_ALLREDUCE, static_cast<double>(n) * 8.0 * (world + 1));
  peer_allreduce_kernel<<<1, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(vec, n, reinterpret_cast<char* const*>(peers), world, rank, seq,
                                                                               static_cast<size_t>(capacity_doubles), static_cast<uint64_t>(timeout_ms) * 1000000ull);
  MSF_LAUNCH_OK("peer_allreduce_kernel");
  return MSF_OK;
}
