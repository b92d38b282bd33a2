// L1 (infonce mode), bf16 tensor-core main loop for sm_100a: TMA -> shared memory -> tcgen05.mma with
// accumulators in TMEM, flash-style, so the Nq x N logit matrix never exists in memory.
//
// Per CTA: one 128-row query tile x one split of the keys.  For every 128-key tile
//   GEMM1  S  = Q_hat K_hat^T          (SS form, both operands K-major SWIZZLE_128B slabs, fp32 in TMEM)
//   soft   P  = exp2(a*S - a)          (two softmax warpgroups ping-pong on two S buffers; P is written
//                                       back over S in TMEM as packed bf16; row sums stay in registers)
//   GEMM2  O += P K_hat                (TS form: A = P from TMEM, B = the SAME smem key tile read MN-major)
// L2-normalised operands bound every logit by 1/tau, so the softmax needs no running max, no rescale and
// no correction pass: partial (rowsum, O) of different key splits simply add (done deterministically by the
// finalize / backward kernels in infonce.cu).
//
// Warp roles (384 threads): w0 TMA producer, w1 MMA issuer, w2 TMEM allocator, w3 idle, w4-7 softmax WG0,
// w8-11 softmax WG1.  TMEM map (512 columns): O [0,D), S0/P0 [256,384), S1/P1 [384,512).
// Algorithmic work: 4*Nq*N*D FLOP per launch (2 GEMMs); bytes 2*(Nq+N)*D read + 4*splits*Nq*(D+1) written.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "infonce_plan.cuh"

namespace msf {
namespace {

constexpr int BM = 128, BN = 128;
constexpr int kThreads = 384;
constexpr uint32_t kSlabBytes = 128 * 128;  // 128 rows x 64 bf16 (one 128-byte swizzle span per row)
constexpr uint32_t kTmemCols = 512, kColS0 = 256, kColS1 = 384;

template <int D>
struct Cfg {
  static constexpr int kSlabs = D / 64;
  static constexpr uint32_t kTileBytes = kSlabs * kSlabBytes;
  static constexpr int kStages = D == 64 ? 6 : (D == 128 ? 4 : 2);
  static constexpr uint32_t kBarBytes = 1024;
  static constexpr uint32_t kSmem = 1024 /*alignment slack*/ + kTileBytes * (1 + kStages) + kBarBytes;
};

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug traps (reported as a CUDA error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint64_t* bar) {  // arrives on `bar` when all prior MMAs of this thread retire
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {  // 32 lanes x 32 consecutive columns
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// UMMA shared-memory descriptor, SWIZZLE_128B, version 1 (sm_100).  Byte offsets are encoded >> 4.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
         (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: fp32 accumulate, bf16 x bf16, M=128
__host__ __device__ constexpr uint32_t umma_idesc(int n, bool b_mn_major) {
  return (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ | (b_mn_major ? (1u << 16) : 0u) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
}

template <int D>
__global__ void __launch_bounds__(kThreads, 1)
infonce_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k, int64_t n_keys,
                  int64_t k_tiles, int64_t tiles_per_split, int64_t nq_pad, float a, float* __restrict__ rowsum,
                  float* __restrict__ o_part) {
  using C = Cfg<D>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + C::kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sK + C::kStages * C::kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + C::kStages;
  uint64_t* s_full = k_empty + C::kStages;  // [2]
  uint64_t* p_full = s_full + 2;            // [2]
  uint64_t* o_full = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  float* rs_xchg = reinterpret_cast<float*>(bars + 32);  // 128 floats

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q0 = static_cast<int64_t>(blockIdx.x) * BM;
  const int split = blockIdx.y;
  const int64_t kt0 = split * tiles_per_split;
  int64_t kt1 = kt0 + tiles_per_split;
  if (kt1 > k_tiles) kt1 = k_tiles;
  const int T = kt1 > kt0 ? static_cast<int>(kt1 - kt0) : 0;

  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < C::kStages; ++s) { mbar_init(k_full + s, 1); mbar_init(k_empty + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(s_full + b, 1); mbar_init(p_full + b, 128); }
    mbar_init(o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0 && T > 0) {
      tma_prefetch_desc(&tm_q);
      tma_prefetch_desc(&tm_k);
      mbar_expect_tx(q_full, C::kTileBytes);
      for (int s = 0; s < C::kSlabs; ++s) tma_load_2d(sQ + s * kSlabBytes, &tm_q, s * 64, static_cast<int>(q0), q_full);
      for (int t = 0; t < T; ++t) {
        const int stage = t % C::kStages;
        mbar_wait(k_empty + stage, ((t / C::kStages) & 1) ^ 1);
        mbar_expect_tx(k_full + stage, C::kTileBytes);
        uint8_t* dst = sK + stage * C::kTileBytes;
        const int row = static_cast<int>((kt0 + t) * BN);
        for (int s = 0; s < C::kSlabs; ++s) tma_load_2d(dst + s * kSlabBytes, &tm_k, s * 64, row, k_full + stage);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0 && T > 0) {
      constexpr uint32_t idesc1 = umma_idesc(BN, false);  // S = Q K^T : N = 128 keys, B K-major
      constexpr uint32_t idesc2 = umma_idesc(D, true);    // O += P K  : N = D, B MN-major (same smem tile)
      const uint32_t q_addr = smem_u32(sQ);
      auto gemm1 = [&](int t) {
        const int stage = t % C::kStages;
        mbar_wait(k_full + stage, (t / C::kStages) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + stage * C::kTileBytes);
        const uint32_t d_tmem = tmem + ((t & 1) ? kColS1 : kColS0);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k >> 2) * kSlabBytes + (k & 3) * 32;  // 16 bf16 = 32 B inside the 128 B swizzle span
          mma_ss(d_tmem, umma_desc(q_addr + off, 16, 1024), umma_desc(k_addr + off, 16, 1024), idesc1, k > 0);
        }
        tc_commit(s_full + (t & 1));
      };
      auto gemm2 = [&](int t) {
        const int stage = t % C::kStages;
        mbar_wait(p_full + (t & 1), (t >> 1) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + stage * C::kTileBytes);
        const uint32_t p_tmem = tmem + ((t & 1) ? kColS1 : kColS0);
#pragma unroll
        for (int k = 0; k < BN / 16; ++k)  // 16 keys per MMA: 8 packed columns of P, 16 smem rows (2048 B) of K_hat
          mma_ts(tmem, p_tmem + k * 8, umma_desc(k_addr + k * 2048, kSlabBytes, 1024), idesc2, (t > 0 || k > 0));
        tc_commit(k_empty + stage);
      };
      mbar_wait(q_full, 0);
      gemm1(0);
      for (int t = 0; t < T; ++t) {
        if (t + 1 < T) gemm1(t + 1);  // keep the tensor pipe busy while tile t is in the softmax
        gemm2(t);
      }
      tc_commit(o_full);
    }
  } else if (warp >= 4) {
    // ===================== softmax warpgroups =====================
    const int wg = (warp - 4) >> 2;             // 0 or 1 -> S buffer
    const int row = ((warp & 3) << 5) + lane;   // TMEM lane == query row inside the tile
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) << 5) << 16);
    const uint32_t s_addr = lane_base + (wg ? kColS1 : kColS0);
    float rs = 0.f;
    const float neg_a = -a;
    for (int t = wg; t < T; t += 2) {
      mbar_wait(s_full + wg, (t >> 1) & 1);
      tc_fence_after();
      const int64_t key0 = (kt0 + t) * BN;
      const int valid = (n_keys - key0) < BN ? static_cast<int>(n_keys - key0) : BN;
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32], u[16];
        tmem_ld32(s_addr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float e0 = ex2_approx(fmaf(__uint_as_float(v[i]), a, neg_a));
          float e1 = ex2_approx(fmaf(__uint_as_float(v[i + 1]), a, neg_a));
          if (valid < BN) {  // zero-filled (out-of-range) keys of the last tile contribute nothing
            if (c * 32 + i >= valid) e0 = 0.f;
            if (c * 32 + i + 1 >= valid) e1 = 0.f;
          }
          rs += e0 + e1;
          __nv_bfloat162 h = __floats2bfloat162_rn(e0, e1);  // key 2j in the low half, 2j+1 in the high half
          u[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
        tmem_st16(s_addr + c * 16, u);  // P overlays the S columns already consumed
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full + wg);
    }
    // ---- epilogue: row sums (WG1 -> smem -> WG0 -> global), then O from TMEM to the split's partial ----
    if (wg == 1) rs_xchg[row] = rs;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    float* rowsum_dst = rowsum + static_cast<int64_t>(split) * nq_pad + q0;
    if (wg == 0) rowsum_dst[row] = rs + rs_xchg[row];
    float* o_dst = o_part + ((static_cast<int64_t>(split) * nq_pad + q0 + row) * D);
    constexpr int kHalf = D / 2;  // WG0 drains columns [0, D/2), WG1 [D/2, D)
    if (T > 0) {
      mbar_wait(o_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = wg * kHalf; c0 < (wg + 1) * kHalf; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(lane_base + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<uint4*>(o_dst + c0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    } else {
      for (int c0 = wg * kHalf; c0 < (wg + 1) * kHalf; c0 += 4) *reinterpret_cast<uint4*>(o_dst + c0) = make_uint4(0, 0, 0, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
  }
}

// ---- host side ---------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}

int make_map(CUtensorMap* map, const void* base, int64_t rows, int dim) {
  auto fn = encode_fn();
  MSF_REQUIRE(fn, MSF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(dim), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(dim) * 2};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MSF_REQUIRE(r == CUDA_SUCCESS, MSF_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
  return MSF_OK;
}

template <int D>
int launch(const CUtensorMap& tq, const CUtensorMap& tk, int64_t n_keys, const NcePlan& plan, float a, float* rowsum,
           float* o_part, cudaStream_t st) {
  MSF_CUDA_OK(cudaFuncSetAttribute(infonce_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<D>::kSmem));
  dim3 grid(static_cast<unsigned>(plan.q_tiles), static_cast<unsigned>(plan.splits));
  infonce_tc_kernel<D><<<grid, kThreads, Cfg<D>::kSmem, st>>>(tq, tk, n_keys, plan.k_tiles, plan.tiles_per_split, plan.nq_pad, a,
                                                              rowsum, o_part);
  MSF_LAUNCH_OK("infonce_tc_kernel");
  return MSF_OK;
}

}  // namespace

int launch_infonce_tc(const void* q_hat, const void* k_hat, int64_t nq, int64_t n_keys, int dim, float tau, const NcePlan& plan,
                      float* rowsum, float* o_part, cudaStream_t st) {
  MSF_REQUIRE(tc_dim_ok(dim), MSF_ERR_UNSUPPORTED, "tcgen05 path covers dim in {64,128,256}; got %d", dim);
  MSF_REQUIRE(n_keys < (1ll << 31) && nq < (1ll << 31), MSF_ERR_UNSUPPORTED, "row counts must fit 31 bits");
  CUtensorMap tq, tk;
  if (int rc = make_map(&tq, q_hat, nq, dim)) return rc;
  if (int rc = make_map(&tk, k_hat, n_keys, dim)) return rc;
  const float a = 1.4426950408889634f / tau;
  switch (dim) {
    case 64: return launch<64>(tq, tk, n_keys, plan, a, rowsum, o_part, st);
    case 128: return launch<128>(tq, tk, n_keys, plan, a, rowsum, o_part, st);
    default: return launch<256>(tq, tk, n_keys, plan, a, rowsum, o_part, st);
  }
}

}  // namespace msf
