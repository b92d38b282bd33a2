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
#include <stdlib.h>

#include "infonce_plan.cuh"
#include "tc_common.cuh"

namespace msf {
namespace {

using namespace tc;
constexpr int BN = 128;
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

template <int D>
__global__ void __launch_bounds__(kThreads, 1)
infonce_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k, int64_t n_keys,
                  int64_t k_tiles, int64_t tiles_per_split, int64_t nq_pad, float a, float* __restrict__ rowsum,
                  float* __restrict__ o_part, int issue_policy) {
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
      auto gemm1 = [&](int t) {  // operands are known to be ready
        const int stage = t % C::kStages;
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
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + stage * C::kTileBytes);
        const uint32_t p_tmem = tmem + ((t & 1) ? kColS1 : kColS0);
#pragma unroll
        for (int k = 0; k < BN / 16; ++k)  // 16 keys per MMA: 8 packed columns of P, 16 smem rows (2048 B) of K_hat
          mma_ts(tmem, p_tmem + k * 8, umma_desc(k_addr + k * 2048, kSlabBytes, 1024), idesc2, (t > 0 || k > 0));
        tc_commit(k_empty + stage);
      };
      mbar_wait(q_full, 0);
      // Out-of-order issue over two queues: GEMM1(a) needs K tile a in smem and its S buffer free (GEMM2(a-2)
      // issued, i.e. a <= b+1); GEMM2(b) needs P(b) from the softmax.  GEMM1 goes first whenever it can, so the
      // softmax warpgroups always have a tile and a late TMA load is hidden behind the other queue.
      int a_next = 0, b_next = 0;
      if (issue_policy == 0) {
        // in-order issue: GEMM1(t+1) is queued before GEMM2(t) so the tensor pipe works while tile t is in the softmax
        mbar_wait(k_full, 0);
        gemm1(0);
        for (int t = 0; t < T; ++t) {
          if (t + 1 < T) {
            mbar_wait(k_full + ((t + 1) % C::kStages), ((t + 1) / C::kStages) & 1);
            gemm1(t + 1);
          }
          mbar_wait(p_full + (t & 1), (t >> 1) & 1);
          gemm2(t);
        }
      } else {
        long long t_progress = clock64();
        while (b_next < T) {
          const bool g1_idx = a_next < T && a_next <= b_next + 1;  // GEMM1 allowed by the S-buffer rule
          const bool g2_idx = b_next < a_next;                      // GEMM2 has an S tile in flight
          uint64_t* kbar = k_full + (a_next % C::kStages);
          const uint32_t kpar = (a_next / C::kStages) & 1;
          uint64_t* pbar = p_full + (b_next & 1);
          const uint32_t ppar = (b_next >> 1) & 1;
          if (g1_idx && !g2_idx) {
            mbar_wait(kbar, kpar);  // only one candidate: sleep on its barrier (fast hardware wake-up)
            gemm1(a_next++);
          } else if (g2_idx && !g1_idx) {
            mbar_wait(pbar, ppar);
            gemm2(b_next++);
          } else if (mbar_test_wait(kbar, kpar)) {  // both possible: GEMM1 first so the softmax always has a tile
            gemm1(a_next++);
          } else if (mbar_test_wait(pbar, ppar)) {
            gemm2(b_next++);
          } else {
            if (clock64() - t_progress > 4000000000ll) __trap();  // watchdog: a pipeline bug is a CUDA error, not a hung box
            continue;
          }
          t_progress = clock64();
        }
      }
      tc_commit(o_full);
    }
  } else if (warp >= 4) {
    // ===================== softmax warpgroups =====================
    const int wg = (warp - 4) >> 2;             // 0 or 1 -> S buffer
    const int row = ((warp & 3) << 5) + lane;   // TMEM lane == query row inside the tile
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) << 5) << 16);
    const uint32_t s_addr = lane_base + (wg ? kColS1 : kColS0);
    float2 rs2 = make_float2(0.f, 0.f);
    const float2 a2 = make_float2(a, a), na2 = make_float2(-a, -a);
    // one 32-column chunk of S -> 16 packed bf16x2 of P; packed fp32x2 FMA/ADD halve the issue slots per logit
    auto chunk = [&](const uint32_t* v, uint32_t* u, int col0, int valid, bool mask) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float2 y = __ffma2_rn(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), a2, na2);
        float2 e = make_float2(ex2_approx(y.x), ex2_approx(y.y));
        if (mask) {  // zero-filled (out-of-range) keys of the last tile contribute nothing
          if (col0 + i >= valid) e.x = 0.f;
          if (col0 + i + 1 >= valid) e.y = 0.f;
        }
        rs2 = __fadd2_rn(rs2, e);
        __nv_bfloat162 h = __floats2bfloat162_rn(e.x, e.y);  // key 2j in the low half, 2j+1 in the high half
        u[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
      }
    };
    for (int t = wg; t < T; t += 2) {
      mbar_wait(s_full + wg, (t >> 1) & 1);
      tc_fence_after();
      const int64_t key0 = (kt0 + t) * BN;
      const int valid = (n_keys - key0) < BN ? static_cast<int>(n_keys - key0) : BN;
      uint32_t v0[32], v1[32], u[16];
      tmem_ld32(s_addr, v0);
      if (valid == BN) {  // hot path: no per-element masking; the next chunk's TMEM load overlaps the math
        tmem_ld_wait();
        tmem_ld32(s_addr + 32, v1);
        chunk(v0, u, 0, BN, false);
        tmem_st16(s_addr, u);  // P overlays the S columns already consumed
        tmem_ld_wait();
        tmem_ld32(s_addr + 64, v0);
        chunk(v1, u, 32, BN, false);
        tmem_st16(s_addr + 16, u);
        tmem_ld_wait();
        tmem_ld32(s_addr + 96, v1);
        chunk(v0, u, 64, BN, false);
        tmem_st16(s_addr + 32, u);
        tmem_ld_wait();
        chunk(v1, u, 96, BN, false);
        tmem_st16(s_addr + 48, u);
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          if (c > 0) tmem_ld32(s_addr + c * 32, v0);
          tmem_ld_wait();
          chunk(v0, u, c * 32, valid, true);
          tmem_st16(s_addr + c * 16, u);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full + wg);
    }
    const float rs = rs2.x + rs2.y;
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
template <int D>
int launch(const CUtensorMap& tq, const CUtensorMap& tk, int64_t n_keys, const NcePlan& plan, float a, float* rowsum,
           float* o_part, cudaStream_t st) {
  MSF_CUDA_OK(cudaFuncSetAttribute(infonce_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<D>::kSmem));
  dim3 grid(static_cast<unsigned>(plan.q_tiles), static_cast<unsigned>(plan.splits));
  // D = 256 has only two 64 KB key stages: out-of-order issue hides the TMA latency there; with >= 4 stages the
  // in-order schedule is already load-latency free and keeps the issuing warp asleep between tiles
  static const int forced = getenv("MSF_TC_ISSUE_POLICY") ? atoi(getenv("MSF_TC_ISSUE_POLICY")) : -1;
  const int policy = forced >= 0 ? forced : (Cfg<D>::kStages <= 2 ? 1 : 0);
  infonce_tc_kernel<D><<<grid, kThreads, Cfg<D>::kSmem, st>>>(tq, tk, n_keys, plan.k_tiles, plan.tiles_per_split, plan.nq_pad, a,
                                                              rowsum, o_part, policy);
  MSF_LAUNCH_OK("infonce_tc_kernel");
  return MSF_OK;
}

}  // namespace

int launch_infonce_tc(const void* q_hat, const void* k_hat, int64_t nq, int64_t n_keys, int dim, float tau, const NcePlan& plan,
                      float* rowsum, float* o_part, cudaStream_t st) {
  MSF_REQUIRE(tc_dim_ok(dim), MSF_ERR_UNSUPPORTED, "tcgen05 path covers dim in {64,128,256}; got %d", dim);
  MSF_REQUIRE(n_keys < (1ll << 31) && nq < (1ll << 31), MSF_ERR_UNSUPPORTED, "row counts must fit 31 bits");
  CUtensorMap tq, tk;
  if (int rc = make_map_bf16(&tq, q_hat, nq, dim, dim, 64, 128)) return rc;
  if (int rc = make_map_bf16(&tk, k_hat, n_keys, dim, dim, 64, 128)) return rc;
  const float a = 1.4426950408889634f / tau;
  switch (dim) {
    case 64: return launch<64>(tq, tk, n_keys, plan, a, rowsum, o_part, st);
    case 128: return launch<128>(tq, tk, n_keys, plan, a, rowsum, o_part, st);
    default: return launch<256>(tq, tk, n_keys, plan, a, rowsum, o_part, st);
  }
}

}  // namespace msf
