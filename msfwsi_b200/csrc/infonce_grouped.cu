// L1 (infonce mode), grouped: the InfoNCE terms of ALL (branch, level, direction) pairs of the loss block
// (tools/ssl_train.py:448-466 pairs, positives on the diagonal; extension named by BASELINE.json:north_star) in a
// handful of launches, fed straight from the head stage:
//   * queries are the RAW predictor outputs p (bf16 / fp16); their L2 norms come from the row sum-of-squares partials the
//     predictor-tail GEMM left in its epilogue (gemm_grouped.cu) -- the softmax warps own one query row per thread, so
//     1/||p_i|| simply scales that row's logits: no normalise pass over p;
//   * keys are the L2-normalised projector outputs written by the head stage's batch-norm apply (head_bn.cu), laid out as
//     rank-major blocks (this rank's block + the all-gathered blocks of the other ranks) and read through a 3-D TMA map
//     {dim, row, rank}: no repacking of the gathered buffer;
//   * one flash launch per width class D in {64, 128, 256} over all problems of that width (persistent problem table in
//     kernel parameters), the same TMA -> tcgen05 -> TMEM pipeline as infonce_tc.cu (S = QK^T, P = exp2(a_i*S - a),
//     O += PK, no running max: |cos| <= 1 bounds every logit by 1/tau);
//   * widths above 256 (512 and the fuser widths 576..4608) take two grouped GEMM launches per rank block: P = exp2(.) in
//     16 bit with row sums (EXP epilogue), then O_r = P K_r in fp32 -- the workspace is O(Nq * rows_per_rank), not
//     O(Nq * N): the 4.3 GB P matrix of a 16384 x 131072 problem never exists;
//   * one finalize launch over all pairs (positive logit, log-sum, weighted loss partials) + the fixed-order final sum, and
//     ONE backward launch over all pairs.
// Tensor-bound (flash / GEMM launches): 4 * Nq * N * D FLOP per pair.  Keys are detached (backbone.py:188-191): the
// forward already yields dq_hat = (g / tau) (O / sum - k_pos); the backward only applies the normalise Jacobian.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "infonce_plan.cuh"
#include "tc_common.cuh"

namespace msf {

// implemented in gemm_grouped.cu
int gemm_grouped_launch(const msf_gemm_problem* problems, int n_problems, int op_dtype, void* workspace, size_t workspace_bytes, int32_t* counters,
                        void* stream);

namespace {

using namespace tc;
constexpr float kLog2e = 1.4426950408889634f;
constexpr int BN = 128;
constexpr int kThreads = 384;
constexpr uint32_t kSlabBytes = 128 * 128;  // 128 rows x 64 16-bit columns (one 128-byte swizzle span per row)
constexpr uint32_t kTmemCols = 512, kColS0 = 256, kColS1 = 384;
constexpr int kMaxFlash = 8;  // problems per flash launch (ctx / tgt x 2 directions = 4 per width in the reference)

template <int D>
struct Cfg {
  static constexpr int kSlabs = D / 64;
  static constexpr uint32_t kTileBytes = kSlabs * kSlabBytes;
  static constexpr int kStages = D == 64 ? 6 : (D == 128 ? 4 : 2);
  static constexpr uint32_t kBarBytes = 1024;
  static constexpr uint32_t kSmem = 1024 /*alignment slack*/ + kTileBytes * (1 + kStages) + kBarBytes;
};

struct alignas(64) FlashProblem {
  CUtensorMap tq;           // queries (nq, D) 2-D
  CUtensorMap tk;           // keys {D, rows_per_rank, world} 3-D
  const float* q_rowsq;     // [D/64][nq] or null (queries already normalised)
  float* rowsum;            // [splits][nq_pad]
  float* o_part;            // [splits][nq_pad][D]
  int32_t nq, rows_per_rank, world, tiles_per_rank, k_tiles, tiles_per_split, splits, nq_pad, cta_start, cta_end;
  const float* col_bias;    // COLB kernels: [k_tiles * 128] additive exponent term per KEY column (see msf_infonce_dk)
};
struct alignas(64) FlashParams {
  FlashProblem p[kMaxFlash];
  int32_t n;
  float a, eps;
  int32_t issue_policy;
  int32_t exp_mode;  // 0: one MUFU ex2 per logit; 1 / 2: every 4th / 2nd pair of logits through a cubic on the FMA pipes instead
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
constexpr uint32_t kColBiasBytes = BN * sizeof(float);

// COLB = true is the transposed pass behind msf_infonce_dk (gradient of the KEYS, north_star (4)): the exponent of logit
// (row, col) is a * s + col_bias[col] instead of a_row * s - a, the bias travelling with each key tile (one 512-byte bulk copy
// on the tile's own mbarrier).  Everything else -- pipeline, TMEM map, issue order -- is the same kernel.
template <int D, bool COLB>
__global__ void __launch_bounds__(kThreads, 1) infonce_grouped_kernel(const __grid_constant__ FlashParams P) {
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
  float* sCB = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + C::kBarBytes);  // COLB: [kStages][128] column biases

  int pi = 0;
#pragma unroll 1
  while (pi + 1 < P.n && static_cast<int>(blockIdx.x) >= P.p[pi].cta_end) ++pi;
  const FlashProblem& q = P.p[pi];
  // split-major inside a pair: the ~148 co-resident CTAs then hold consecutive query tiles of the SAME key split and walk the
  // same key tiles in near lockstep, so L2 serves each key tile once to all of them (with the splits interleaved two key
  // streams compete for the L2 -> SM feed, which this kernel uses to ~75 %: measured 0.76 vs 0.92 of peak at D = 256)
  const int local = blockIdx.x - q.cta_start;
  const int q_tiles = (q.nq + BM - 1) / BM;
  const int split = local / q_tiles;
  const int64_t q0 = static_cast<int64_t>(local % q_tiles) * BM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt0 = split * q.tiles_per_split;
  const int kt1 = min(q.k_tiles, kt0 + q.tiles_per_split);
  const int T = kt1 > kt0 ? kt1 - kt0 : 0;
  const float a = P.a;

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
      tma_prefetch_desc(&q.tq);
      tma_prefetch_desc(&q.tk);
      mbar_expect_tx(q_full, C::kTileBytes);
      for (int s = 0; s < C::kSlabs; ++s) tma_load_2d(sQ + s * kSlabBytes, &q.tq, s * 64, static_cast<int>(q0), q_full);
      // a key tile never straddles two rank blocks; (rank, row) advance incrementally -- a division per tile would sit between a
      // stage becoming free and its refill, which the two 64 KB stages of D = 256 cannot hide
      const int tpr = q.tiles_per_rank;
      int rank = kt0 / tpr, tile_in_rank = kt0 - rank * tpr;
      for (int t = 0; t < T; ++t) {
        const int stage = t % C::kStages;
        mbar_wait(k_empty + stage, ((t / C::kStages) & 1) ^ 1);
        mbar_expect_tx(k_full + stage, C::kTileBytes + (COLB ? kColBiasBytes : 0u));
        uint8_t* dst = sK + stage * C::kTileBytes;
        const int row = tile_in_rank * BN;
        if constexpr (COLB) bulk_load_1d(sCB + stage * BN, q.col_bias + static_cast<size_t>(kt0 + t) * BN, kColBiasBytes, k_full + stage);
        for (int s = 0; s < C::kSlabs; ++s) tma_load_3d(dst + s * kSlabBytes, &q.tk, s * 64, row, rank, k_full + stage);
        if (++tile_in_rank == tpr) { tile_in_rank = 0; ++rank; }
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
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + stage * C::kTileBytes);
        const uint32_t d_tmem = tmem + ((t & 1) ? kColS1 : kColS0);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k >> 2) * kSlabBytes + (k & 3) * 32;
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
        for (int k = 0; k < BN / 16; ++k)
          mma_ts(tmem, p_tmem + k * 8, umma_desc(k_addr + k * 2048, kSlabBytes, 1024), idesc2, (t > 0 || k > 0));
        tc_commit(k_empty + stage);
      };
      mbar_wait(q_full, 0);
      int a_next = 0, b_next = 0;
      if (P.issue_policy == 0) {
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
          const bool g1_idx = a_next < T && a_next <= b_next + 1;
          const bool g2_idx = b_next < a_next;
          uint64_t* kbar = k_full + (a_next % C::kStages);
          const uint32_t kpar = (a_next / C::kStages) & 1;
          uint64_t* pbar = p_full + (b_next & 1);
          const uint32_t ppar = (b_next >> 1) & 1;
          if (g1_idx && !g2_idx) {
            mbar_wait(kbar, kpar);
            gemm1(a_next++);
          } else if (g2_idx && !g1_idx) {
            mbar_wait(pbar, ppar);
            gemm2(b_next++);
          } else if (mbar_test_wait(kbar, kpar)) {
            gemm1(a_next++);
          } else if (mbar_test_wait(pbar, ppar)) {
            gemm2(b_next++);
          } else {
            if (clock64() - t_progress > 4000000000ll) __trap();
            continue;
          }
          t_progress = clock64();
        }
      }
      tc_commit(o_full);
    }
  } else if (warp >= 4) {
    // ===================== softmax warpgroups =====================
    const int wg = (warp - 4) >> 2;
    const int row = ((warp & 3) << 5) + lane;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) << 5) << 16);
    const uint32_t s_addr = lane_base + (wg ? kColS1 : kColS0);
    // per-row scale: logits = (p_i . k_hat_j) / (||p_i|| tau); the norm comes from the predictor-tail GEMM's epilogue
    float a_row = a;
    if (q.q_rowsq) {
      float ss = 0.f;
      if (q0 + row < q.nq)
        for (int b = 0; b < D / 64; ++b) ss += __ldg(q.q_rowsq + static_cast<size_t>(b) * q.nq + q0 + row);
      a_row = a / fmaxf(sqrtf(ss), P.eps);
    }
    float2 rs2 = make_float2(0.f, 0.f);
    const float2 a2 = make_float2(a_row, a_row), na2 = make_float2(-a, -a);
    auto chunk = [&](const uint32_t* v, uint32_t* u, int col0, int valid, bool mask, const float* cb) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        float2 nb = na2;
        if constexpr (COLB) nb = *reinterpret_cast<const float2*>(cb + col0 + i);  // same address in every lane: a broadcast
        const float2 y = __ffma2_rn(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), a2, nb);
        float2 e = make_float2(ex2_approx(y.x), ex2_approx(y.y));
        if (mask) {
          if (col0 + i >= valid) e.x = 0.f;
          if (col0 + i + 1 >= valid) e.y = 0.f;
        }
        rs2 = __fadd2_rn(rs2, e);
        __nv_bfloat162 h = __floats2bfloat162_rn(e.x, e.y);
        u[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
      }
    };
    // MUFU-lean variant.  At D <= 128 one ex2 per logit needs as many cycles on the SM's 16 MUFU lanes as the two GEMMs need on
    // the tensor pipe (1024 cycles per 128 x 128 tile each), so any hiccup of either stalls the other.  Every 4th (exp_mode 1)
    // or every 2nd (exp_mode 2) PAIR of logits therefore takes the FMA pipes instead: 2^y = 2^round(y) * p(y - round(y)) with a
    // cubic p on [-0.5, 0.5] (max relative error 1.0e-4, far below the bf16 rounding of P) -- magic-number rounding, three packed
    // FFMA2 and an integer add into the exponent field per pair.
    auto chunk_mix = [&](const uint32_t* v, uint32_t* u, const int mask) {
      const float2 magic = make_float2(12582912.f, 12582912.f), nmagic = make_float2(-12582912.f, -12582912.f), none = make_float2(-1.f, -1.f);
      const float2 c1 = make_float2(0.6932829022407532f, 0.6932829022407532f), c2 = make_float2(0.24221095442771912f, 0.24221095442771912f),
                   c3 = make_float2(0.05500892922282219f, 0.05500892922282219f), one = make_float2(1.f, 1.f);
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float2 y = __ffma2_rn(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), a2, na2);
        float2 e;
        if (((i >> 1) & mask) == mask) {
          const float2 r = __fadd2_rn(y, magic);                  // low mantissa bits of r = round(y) as an integer
          const float2 f = __ffma2_rn(__fadd2_rn(r, nmagic), none, y);  // y - round(y) in [-0.5, 0.5]
          float2 pl = __ffma2_rn(c3, f, c2);
          pl = __ffma2_rn(pl, f, c1);
          pl = __ffma2_rn(pl, f, one);
          e.x = __uint_as_float(__float_as_uint(pl.x) + (__float_as_uint(r.x) << 23));  // * 2^round(y): the shift drops the magic's bits
          e.y = __uint_as_float(__float_as_uint(pl.y) + (__float_as_uint(r.y) << 23));
        } else {
          e = make_float2(ex2_approx(y.x), ex2_approx(y.y));
        }
        rs2 = __fadd2_rn(rs2, e);
        __nv_bfloat162 h = __floats2bfloat162_rn(e.x, e.y);
        u[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
      }
    };
    const int tpr_s = q.tiles_per_rank, rows_s = q.rows_per_rank;
    // COLB: the bias of the padding columns is -1e30 (their exponential is 0), and the polynomial exponential is off -- its
    // exponent-field arithmetic assumes y >= -126, which log2(softmax) does not guarantee
    const bool ragged = !COLB && (rows_s % BN) != 0;
    const bool fast = !COLB && P.exp_mode != 0;
    const int mix_mask = P.exp_mode == 2 ? 1 : 3;
    for (int t = wg; t < T; t += 2) {
      mbar_wait(s_full + wg, (t >> 1) & 1);
      tc_fence_after();
      // rows past a rank block are zero-filled by TMA and must be masked out: only the last tile of a rank block can be short
      int valid = BN;
      if (ragged) {
        const int krow0 = ((kt0 + t) % tpr_s) * BN;
        valid = min(BN, rows_s - krow0);
      }
      uint32_t v0[32], v1[32], u[16];
      const float* cb = sCB + (t % C::kStages) * BN;  // stays valid until gemm2(t), which waits for this warpgroup's p_full
      tmem_ld32(s_addr, v0);
      if (valid == BN && fast) {
        tmem_ld_wait();
        tmem_ld32(s_addr + 32, v1);
        chunk_mix(v0, u, mix_mask);
        tmem_st16(s_addr, u);
        tmem_ld_wait();
        tmem_ld32(s_addr + 64, v0);
        chunk_mix(v1, u, mix_mask);
        tmem_st16(s_addr + 16, u);
        tmem_ld_wait();
        tmem_ld32(s_addr + 96, v1);
        chunk_mix(v0, u, mix_mask);
        tmem_st16(s_addr + 32, u);
        tmem_ld_wait();
        chunk_mix(v1, u, mix_mask);
        tmem_st16(s_addr + 48, u);
      } else if (valid == BN) {
        tmem_ld_wait();
        tmem_ld32(s_addr + 32, v1);
        chunk(v0, u, 0, BN, false, cb);
        tmem_st16(s_addr, u);
        tmem_ld_wait();
        tmem_ld32(s_addr + 64, v0);
        chunk(v1, u, 32, BN, false, cb);
        tmem_st16(s_addr + 16, u);
        tmem_ld_wait();
        tmem_ld32(s_addr + 96, v1);
        chunk(v0, u, 64, BN, false, cb);
        tmem_st16(s_addr + 32, u);
        tmem_ld_wait();
        chunk(v1, u, 96, BN, false, cb);
        tmem_st16(s_addr + 48, u);
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          if (c > 0) tmem_ld32(s_addr + c * 32, v0);
          tmem_ld_wait();
          chunk(v0, u, c * 32, valid, true, cb);
          tmem_st16(s_addr + c * 16, u);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full + wg);
    }
    const float rs = rs2.x + rs2.y;
    if (wg == 1) rs_xchg[row] = rs;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    float* rowsum_dst = q.rowsum + static_cast<int64_t>(split) * q.nq_pad + q0;
    if (wg == 0) rowsum_dst[row] = rs + rs_xchg[row];
    float* o_dst = q.o_part + ((static_cast<int64_t>(split) * q.nq_pad + q0 + row) * D);
    constexpr int kHalf = D / 2;
    if (T > 0) {
      mbar_wait(o_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = wg * kHalf; c0 < (wg + 1) * kHalf; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(lane_base + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<uint4*>(o_dst + c0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
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

// ---- finalize / backward over all pairs ---------------------------------------------------------------------------
struct PairDev {
  const void* q;          // (nq, D) raw or normalised queries
  const float* q_rowsq;   // [D/64][nq] or null
  const void* k_local;    // (rows_per_rank, D) this rank's normalised keys: row i is the positive of query i
  const float* rowsum;    // [splits][nq_pad]
  const float* o_part;    // [splits][nq_pad][D]
  float* row_stat;        // [nq][2] {total row sum, 1/max(||q||, eps)}
  void* grad_q;           // (nq, D) (backward)
  int32_t nq, D, splits, nq_pad, blk_start, blk_end;
  float coef_over_rows;   // coef / nq
  int32_t pad_;
};
struct PairParams {
  PairDev p[MSF_NCE_MAX_PAIRS];
  int32_t n;
  float inv_tau, eps;
};

template <int DT>
__global__ void __launch_bounds__(256) nce_grouped_final_kernel(const __grid_constant__ PairParams P, float* __restrict__ partials) {
  constexpr int V = Elem<DT>::VEC;
  int pi = 0;
#pragma unroll 1
  while (pi + 1 < P.n && static_cast<int>(blockIdx.x) >= P.p[pi].blk_end) ++pi;
  const PairDev& q = P.p[pi];
  const uint32_t cpr = q.D / V;
  uint32_t lanes = 1;
  while (lanes < 32 && lanes * 4 < cpr) lanes <<= 1;
  const uint32_t lane = threadIdx.x & (lanes - 1), grp = threadIdx.x / lanes, groups = 256 / lanes;
  const int64_t row = static_cast<int64_t>(blockIdx.x - q.blk_start) * groups + grp;
  const bool valid = row < q.nq;
  const size_t row_bytes = static_cast<size_t>(cpr) * 16;
  float dot = 0.f;
  if (valid) {
    const char* a = static_cast<const char*>(q.q) + row * row_bytes;
    const char* b = static_cast<const char*>(q.k_local) + row * row_bytes;
    for (uint32_t c = lane; c < cpr; c += lanes) {
      float fa[V], fb[V];
      Elem<DT>::unpack(ldg_keep(a + static_cast<size_t>(c) * 16), fa);
      Elem<DT>::unpack(ldg_keep(b + static_cast<size_t>(c) * 16), fb);
#pragma unroll
      for (int i = 0; i < V; ++i) dot = fmaf(fa[i], fb[i], dot);
    }
  }
  dot = group_sum(dot, lanes);
  float loss = 0.f;
  if (valid && lane == 0) {
    float inv = 1.f;
    if (q.q_rowsq) {
      float ss = 0.f;
      for (int b = 0; b < (q.D + 63) / 64; ++b) ss += __ldg(q.q_rowsq + static_cast<size_t>(b) * q.nq + row);
      inv = 1.f / fmaxf(sqrtf(ss), P.eps);
    }
    float sum = 0.f;
    for (int s = 0; s < q.splits; ++s) sum += q.rowsum[static_cast<int64_t>(s) * q.nq_pad + row];  // fixed order
    const float s_ii = dot * inv;  // cosine of the positive pair
    loss = (logf(sum) + P.inv_tau - s_ii * P.inv_tau) * q.coef_over_rows;
    q.row_stat[2 * row] = sum;
    q.row_stat[2 * row + 1] = inv;
  }
  __shared__ float sm[8];
  const float w = group_sum(loss, 32);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i];
    partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) nce_grouped_sum_kernel(const float* __restrict__ partials, uint32_t n, float* out) {
  __shared__ double sm[256];
  double t = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += 256) t += static_cast<double>(partials[i]);
  sm[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = static_cast<float>(sm[0]);
}

__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  const uint4 r = ldg_stream(p);
  return make_float4(__uint_as_float(r.x), __uint_as_float(r.y), __uint_as_float(r.z), __uint_as_float(r.w));
}

// dq = g * coef/(nq tau) * J^T (O/sum - k_pos),  J = d(q/||q||)/dq:  dq = (dq_hat - q_hat (q_hat . dq_hat)) / ||q||
template <int DT>
__global__ void __launch_bounds__(256) nce_grouped_bwd_kernel(const __grid_constant__ PairParams P, const float* __restrict__ grad_out) {
  constexpr int V = Elem<DT>::VEC;
  int pi = 0;
#pragma unroll 1
  while (pi + 1 < P.n && static_cast<int>(blockIdx.x) >= P.p[pi].blk_end) ++pi;
  const PairDev& q = P.p[pi];
  const uint32_t cpr = q.D / V;
  uint32_t lanes = 1;
  while (lanes < 32 && lanes * 4 < cpr) lanes <<= 1;
  const uint32_t lane = threadIdx.x & (lanes - 1), grp = threadIdx.x / lanes, groups = 256 / lanes;
  const int64_t row = static_cast<int64_t>(blockIdx.x - q.blk_start) * groups + grp;
  const bool valid = row < q.nq;
  const size_t row_bytes = static_cast<size_t>(cpr) * 16;
  const float gs = __ldg(grad_out) * q.coef_over_rows * P.inv_tau;
  const float inv_sum = valid ? 1.f / q.row_stat[2 * row] : 0.f;
  const float inn = valid ? q.row_stat[2 * row + 1] : 0.f;
  const char* qrow = static_cast<const char*>(q.q) + row * row_bytes;
  const char* krow = static_cast<const char*>(q.k_local) + row * row_bytes;
  // d[i] = gs * (O_i / sum - k_pos,i): the O partials of the key splits are summed in fixed order, read as 128-bit vectors
  auto load_d = [&](uint32_t c, float* fq, float* d) {
    float fk[V];
    Elem<DT>::unpack(ldg_keep(qrow + static_cast<size_t>(c) * 16), fq);
    Elem<DT>::unpack(ldg_keep(krow + static_cast<size_t>(c) * 16), fk);
    float o[V];
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] = 0.f;
    for (int s = 0; s < q.splits; ++s) {
      const float4* src = reinterpret_cast<const float4*>(q.o_part + (static_cast<int64_t>(s) * q.nq_pad + row) * q.D + c * V);
#pragma unroll
      for (int i = 0; i < V; i += 4) {
        const float4 v = ldg_stream_f4(src + i / 4);
        o[i] += v.x; o[i + 1] += v.y; o[i + 2] += v.z; o[i + 3] += v.w;
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) d[i] = gs * (o[i] * inv_sum - fk[i]);
  };
  constexpr int kKeep = 4;  // chunks per lane kept in registers between the two passes (every flash width: D <= 256)
  if (cpr <= lanes * kKeep) {
    float fq[kKeep][V], d[kKeep][V];
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < kKeep; ++j) {
      const uint32_t c = lane + j * lanes;
      if (valid && c < cpr) {
        load_d(c, fq[j], d[j]);
#pragma unroll
        for (int i = 0; i < V; ++i) t = fmaf(fq[j][i] * inn, d[j][i], t);
      }
    }
    t = group_sum(t, lanes);
    if (!valid) return;
#pragma unroll
    for (int j = 0; j < kKeep; ++j) {
      const uint32_t c = lane + j * lanes;
      if (c < cpr) {
        float g[V];
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] = (d[j][i] - fq[j][i] * inn * t) * inn;
        stg_stream(static_cast<char*>(q.grad_q) + row * row_bytes + static_cast<size_t>(c) * 16, Elem<DT>::pack(g));
      }
    }
    return;
  }
  // wide rows (the fuser widths): two passes over the row, the second one served by L2
  float t = 0.f;
  if (valid)
    for (uint32_t c = lane; c < cpr; c += lanes) {
      float fq[V], d[V];
      load_d(c, fq, d);
#pragma unroll
      for (int i = 0; i < V; ++i) t = fmaf(fq[i] * inn, d[i], t);
    }
  t = group_sum(t, lanes);
  if (!valid) return;
  for (uint32_t c = lane; c < cpr; c += lanes) {
    float fq[V], d[V], g[V];
    load_d(c, fq, d);
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = (d[i] - fq[i] * inn * t) * inn;
    stg_stream(static_cast<char*>(q.grad_q) + row * row_bytes + static_cast<size_t>(c) * 16, Elem<DT>::pack(g));
  }
}

// ---- host: plan shared by forward and backward (pure function of the problem list) ---------------------------------
struct PairPlan {
  int mode;        // 0 flash, 2 two-pass GEMMs per rank block
  int splits;      // O partials: flash key splits, or world (one per rank block)
  int rs_splits;   // row-sum partials
  int q_tiles, k_tiles, tiles_per_rank, tiles_per_split, nq_pad;
  int p_ranks;     // two-pass: rank blocks whose P matrices are resident at once (the rest reuse the slots in later rounds)
  size_t off_rowsum, off_o, off_stat, off_p, off_inv;
  int64_t ld_p;
};
struct Plan {
  std::vector<PairPlan> pp;
  size_t off_partials, partial_cap, off_gemm_ws, gemm_ws_bytes, total;
};
inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

bool flash_dim(int d) { return d == 64 || d == 128 || d == 256; }

Plan make_plan(const msf_nce_pair* pr, int n) {
  Plan pl;
  pl.pp.resize(n);
  // The pairs of a width class share ONE flash launch: choose the key tiles per CTA, T, that minimises
  // waves x (T + 3) over the whole class (3 tile-times stand for a CTA's prologue / epilogue), waves = ceil(CTAs / 148) with
  // CTAs = sum_pairs q_tiles x ceil(k_tiles / T).  Ties go to the larger T (fewer partials to write and re-read).
  auto cls = [](int d) { return d == 64 ? 0 : (d == 128 ? 1 : 2); };
  int best_T[3] = {1, 1, 1};
  for (int c = 0; c < 3; ++c) {
    int max_kt = 0;
    for (int i = 0; i < n; ++i)
      if (flash_dim(pr[i].D) && cls(pr[i].D) == c) max_kt = std::max(max_kt, ((pr[i].rows_per_rank + BN - 1) / BN) * pr[i].world);
    double best = 1e300;
    for (int T = max_kt; T >= 1; --T) {
      int64_t ctas = 0;
      bool ok = true;
      for (int i = 0; i < n; ++i) {
        if (!flash_dim(pr[i].D) || cls(pr[i].D) != c) continue;
        const int64_t kt = static_cast<int64_t>((pr[i].rows_per_rank + BN - 1) / BN) * pr[i].world;
        const int64_t sp = (kt + T - 1) / T;
        if (sp > 32) ok = false;  // at most 32 O partials per pair
        ctas += static_cast<int64_t>((pr[i].nq + BM - 1) / BM) * sp;
      }
      if (!ok) break;
      const double cost = static_cast<double>((ctas + kNumSMs - 1) / kNumSMs) * (T + 3.0);
      if (cost < best * 0.999) { best = cost; best_T[c] = T; }
    }
  }
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    const msf_nce_pair& g = pr[i];
    PairPlan& p = pl.pp[i];
    p.q_tiles = (g.nq + BM - 1) / BM;
    p.nq_pad = p.q_tiles * BM;
    if (flash_dim(g.D)) {
      p.mode = 0;
      p.tiles_per_rank = (g.rows_per_rank + BN - 1) / BN;
      p.k_tiles = p.tiles_per_rank * g.world;
      p.tiles_per_split = std::min(best_T[cls(g.D)], p.k_tiles);
      p.splits = (p.k_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
      p.rs_splits = p.splits;
      p.ld_p = 0;
    } else {
      p.mode = 2;
      p.splits = g.world;  // one O partial per rank block
      p.tiles_per_rank = (g.rows_per_rank + 63) / 64;
      p.rs_splits = g.world * p.tiles_per_rank;  // row-sum partials: one per 64-key block
      p.k_tiles = p.tiles_per_split = 0;
      p.ld_p = (g.rows_per_rank + 7) & ~7;
    }
    p.off_rowsum = off;
    off = align256(off + static_cast<size_t>(p.rs_splits) * p.nq_pad * sizeof(float));
    p.off_o = off;
    off = align256(off + static_cast<size_t>(p.splits) * p.nq_pad * g.D * sizeof(float));
    p.off_stat = off;
    off = align256(off + static_cast<size_t>(g.nq) * 2 * sizeof(float));
    p.off_inv = off;
    if (p.mode == 2) off = align256(off + static_cast<size_t>(g.nq) * sizeof(float));
    p.off_p = off;
    p.p_ranks = 1;
    if (p.mode == 2) {
      // P = exp2(.) of one rank block is nq x rows_per_rank 16-bit values; keep at most ~256 MB of them resident per pair and
      // walk the remaining rank blocks in later rounds (a 16384 x 131072 problem needs 512 MB instead of 4.3 GB)
      const size_t per_rank = static_cast<size_t>(g.nq) * p.ld_p * 2;
      const size_t fit = per_rank ? (static_cast<size_t>(256) << 20) / per_rank : 1;
      p.p_ranks = static_cast<int>(std::max<size_t>(1, std::min<size_t>(fit, g.world)));
      off = align256(off + static_cast<size_t>(p.p_ranks) * per_rank);
    }
  }
  size_t blocks = 0;
  for (int i = 0; i < n; ++i) blocks += static_cast<size_t>(pr[i].nq) / 8 + 2;
  pl.off_partials = off;
  pl.partial_cap = blocks;
  off = align256(off + blocks * sizeof(float));
  pl.off_gemm_ws = off;
  pl.gemm_ws_bytes = 0;
  pl.total = off;
  return pl;
}

int check_pairs(const msf_nce_pair* pr, int n, int dtype, float tau) {
  MSF_REQUIRE(pr && n > 0 && n <= MSF_NCE_MAX_PAIRS, MSF_ERR_INVALID, "n_pairs %d outside [1, %d]", n, MSF_NCE_MAX_PAIRS);
  MSF_REQUIRE(dtype == MSF_BF16, MSF_ERR_UNSUPPORTED, "the grouped InfoNCE path computes in bf16 (fp32 inputs take msf_infonce_fwd)");
  MSF_REQUIRE(tau > 0.f && 2.f * kLog2e / tau <= 120.f, MSF_ERR_UNSUPPORTED,
              "tau=%g outside the supported range (tau >= 0.0241): fixed-bound softmax would underflow fp32", tau);
  for (int i = 0; i < n; ++i) {
    const msf_nce_pair& g = pr[i];
    MSF_REQUIRE(g.nq > 0 && g.rows_per_rank > 0 && g.world >= 1 && g.D > 0 && g.D % 64 == 0, MSF_ERR_INVALID,
                "pair %d: bad shape (nq %d, rows_per_rank %d, world %d, D %d: widths are multiples of 64)", i, g.nq, g.rows_per_rank, g.world, g.D);
    MSF_REQUIRE(g.q && g.keys && aligned16(g.q) && aligned16(g.keys), MSF_ERR_INVALID, "pair %d: NULL or misaligned pointer", i);
    MSF_REQUIRE(g.pos_rank >= 0 && g.pos_rank < g.world && g.nq <= g.rows_per_rank, MSF_ERR_INVALID, "pair %d: positives fall outside the keys", i);
    MSF_REQUIRE(g.world == 1 || (g.rank_stride * 2) % 16 == 0, MSF_ERR_INVALID, "pair %d: rank stride must be a multiple of 8 elements", i);
  }
  return MSF_OK;
}

inline int make_map_keys(CUtensorMap* map, const void* base, int D, int rows, int world, int64_t rank_stride) {
  auto fn = encode_fn();
  MSF_REQUIRE(fn, MSF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(world)};
  const cuuint64_t gstride[2] = {static_cast<cuuint64_t>(D) * 2, static_cast<cuuint64_t>(world > 1 ? rank_stride : static_cast<int64_t>(rows) * D) * 2};
  const cuuint32_t box[3] = {64, 128, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MSF_REQUIRE(r == CUDA_SUCCESS, MSF_ERR_CUDA, "cuTensorMapEncodeTiled (3-D keys) failed with CUresult %d", static_cast<int>(r));
  return MSF_OK;
}

// which exponential the softmax warps use per width class (see chunk_fast): decided by measurement, profiles/README.md
// Measured at N = Nq = 65536 (profiles/README.md): D = 128 forward 0.745 -> 0.785 of the bf16 peak with every 4th pair on the
// FMA pipes (MUFU and tensor pipe both need 1024 cycles per tile there); D = 256 is tensor-bound (2048 cycles per tile) and
// loses 3 % to the extra issue slots, so it keeps one MUFU op per logit.
inline int kDefaultExpMode(int D) { return D <= 128 ? 1 : 0; }

template <int D, bool COLB = false>
int launch_flash(const FlashParams& FP, int ctas, cudaStream_t st) {
  constexpr uint32_t smem = Cfg<D>::kSmem + (COLB ? Cfg<D>::kStages * kColBiasBytes : 0u);
  static const cudaError_t attr = cudaFuncSetAttribute(infonce_grouped_kernel<D, COLB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  MSF_REQUIRE(attr == cudaSuccess, MSF_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr));
  infonce_grouped_kernel<D, COLB><<<ctas, kThreads, smem, st>>>(FP);
  MSF_LAUNCH_OK("infonce_grouped_kernel");
  return MSF_OK;
}

// q_hat-free inverse norms for the two-pass path's EXP epilogue: inv[i] = 1 / max(sqrt(sum_b rowsq[b][i]), eps)
__global__ void __launch_bounds__(256) nce_inv_norm_kernel(const float* __restrict__ rowsq, int nq, int blocks, float eps, float* __restrict__ inv) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= nq) return;
  float ss = 0.f;
  for (int b = 0; b < blocks; ++b) ss += rowsq[static_cast<size_t>(b) * nq + i];
  inv[i] = 1.f / fmaxf(sqrtf(ss), eps);
}

int fill_pairs(PairParams& PP, const msf_nce_pair* pr, int n, const Plan& pl, char* ws, float tau, float eps, int vec, int* blocks_out) {
  int blk = 0;
  for (int i = 0; i < n; ++i) {
    const msf_nce_pair& g = pr[i];
    const PairPlan& p = pl.pp[i];
    PairDev& d = PP.p[i];
    d.q = g.q;
    d.q_rowsq = g.q_rowsq;
    d.k_local = static_cast<const char*>(g.keys) + static_cast<size_t>(g.pos_rank) * (g.world > 1 ? g.rank_stride : 0) * 2;
    d.rowsum = reinterpret_cast<const float*>(ws + p.off_rowsum);
    d.o_part = reinterpret_cast<const float*>(ws + p.off_o);
    d.row_stat = reinterpret_cast<float*>(ws + p.off_stat);
    d.grad_q = g.grad_q;
    d.nq = g.nq; d.D = g.D; d.splits = p.splits; d.nq_pad = p.nq_pad;
    d.coef_over_rows = g.coef / static_cast<float>(g.nq);
    const uint32_t cpr = g.D / vec;
    uint32_t lanes = 1;
    while (lanes < 32 && lanes * 4 < cpr) lanes <<= 1;
    const int groups = 256 / static_cast<int>(lanes);
    d.blk_start = blk;
    blk += (g.nq + groups - 1) / groups;
    d.blk_end = blk;
  }
  PP.n = n;
  PP.inv_tau = 1.f / tau;
  PP.eps = eps;
  *blocks_out = blk;
  return MSF_OK;
}

}  // namespace

// The transposed pass of msf_infonce_dk (infonce.cu): one "pair" whose query rows are ALL keys (n_keys, contiguous) and whose
// key columns are this rank's normalised queries (one block, world = 1), exponent a * s + col_bias[column].
int launch_infonce_dk_flash(const void* k_all, const void* q_hat, int64_t n_keys, int64_t nq, int dim, float tau, const NcePlan& planT,
                            const float* col_bias, float* rowsum, float* o_part, cudaStream_t st) {
  MSF_REQUIRE(flash_dim(dim) && n_keys < (1ll << 31) - 128 && nq < (1ll << 31) - 128, MSF_ERR_UNSUPPORTED, "key-gradient flash pass: bad shape");
  MSF_REQUIRE(planT.q_tiles * planT.splits < (1ll << 31), MSF_ERR_UNSUPPORTED, "key-gradient flash pass: too many CTAs");
  static thread_local FlashParams FP;
  FlashProblem& f = FP.p[0];
  f = FlashProblem{};
  if (int rc = make_map_bf16(&f.tq, k_all, n_keys, dim, dim, 64, 128)) return rc;
  if (int rc = make_map_keys(&f.tk, q_hat, dim, static_cast<int>(nq), 1, 0)) return rc;
  f.q_rowsq = nullptr;
  f.rowsum = rowsum;
  f.o_part = o_part;
  f.nq = static_cast<int32_t>(n_keys);
  f.rows_per_rank = static_cast<int32_t>(nq);
  f.world = 1;
  f.tiles_per_rank = f.k_tiles = static_cast<int32_t>(planT.k_tiles);
  f.tiles_per_split = static_cast<int32_t>(planT.tiles_per_split);
  f.splits = planT.splits;
  f.nq_pad = static_cast<int32_t>(planT.nq_pad);
  f.cta_start = 0;
  f.cta_end = static_cast<int32_t>(planT.q_tiles * planT.splits);
  f.col_bias = col_bias;
  FP.n = 1;
  FP.a = kLog2e / tau;
  FP.eps = 1e-8f;
  FP.issue_policy = dim == 256 ? 1 : 0;
  FP.exp_mode = 0;
  return dim == 64 ? launch_flash<64, true>(FP, f.cta_end, st) : dim == 128 ? launch_flash<128, true>(FP, f.cta_end, st)
                                                                          : launch_flash<256, true>(FP, f.cta_end, st);
}

}  // namespace msf

using namespace msf;

extern "C" size_t msf_nce_grouped_workspace_bytes(const msf_nce_pair* pairs, int n_pairs) {
  if (!pairs || n_pairs <= 0 || n_pairs > MSF_NCE_MAX_PAIRS) return 0;
  return make_plan(pairs, n_pairs).total;
}

extern "C" int msf_nce_grouped_fwd(const msf_nce_pair* pairs, int n_pairs, int dtype, float tau, float eps, float* loss_out, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (int rc = check_pairs(pairs, n_pairs, dtype, tau)) return rc;
  MSF_REQUIRE(loss_out, MSF_ERR_INVALID, "loss_out is NULL");
  const Plan pl = make_plan(pairs, n_pairs);
  MSF_REQUIRE(workspace && aligned16(workspace) && workspace_bytes >= pl.total, MSF_ERR_WORKSPACE, "workspace of %zu bytes < %zu required", workspace_bytes,
              pl.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  const float a = kLog2e / tau;

  // ---- flash launches, one per width class ----
  static thread_local FlashParams FP;
  static const int exp_env = getenv("MSF_NCE_EXP_MODE") ? atoi(getenv("MSF_NCE_EXP_MODE")) : -1;
  for (int D : {64, 128, 256}) {
    int nf = 0, ctas = 0;
    double flops = 0.0;
    for (int i = 0; i < n_pairs; ++i) {
      const msf_nce_pair& g = pairs[i];
      if (g.D != D) continue;
      const PairPlan& p = pl.pp[i];
      if (nf == kMaxFlash) {  // table full: flush
        FP.n = nf; FP.a = a; FP.eps = eps; FP.issue_policy = D == 256 ? 1 : 0;
        FP.exp_mode = exp_env >= 0 ? exp_env : kDefaultExpMode(D);
        ProfScope prof(stream, MSF_K_NCE_FLASH, flops);
        if (int rc = (D == 64 ? launch_flash<64>(FP, ctas, st) : D == 128 ? launch_flash<128>(FP, ctas, st) : launch_flash<256>(FP, ctas, st))) return rc;
        nf = 0; ctas = 0; flops = 0.0;
      }
      FlashProblem& f = FP.p[nf++];
      f = FlashProblem{};
      if (int rc = make_map_bf16(&f.tq, g.q, g.nq, g.D, g.D, 64, 128)) return rc;
      if (int rc = make_map_keys(&f.tk, g.keys, g.D, g.rows_per_rank, g.world, g.rank_stride)) return rc;
      f.q_rowsq = g.q_rowsq;
      f.rowsum = reinterpret_cast<float*>(ws + p.off_rowsum);
      f.o_part = reinterpret_cast<float*>(ws + p.off_o);
      f.nq = g.nq; f.rows_per_rank = g.rows_per_rank; f.world = g.world; f.tiles_per_rank = p.tiles_per_rank; f.k_tiles = p.k_tiles;
      f.tiles_per_split = p.tiles_per_split; f.splits = p.splits; f.nq_pad = p.nq_pad;
      f.cta_start = ctas;
      ctas += p.q_tiles * p.splits;
      f.cta_end = ctas;
      flops += 4.0 * g.nq * static_cast<double>(g.rows_per_rank) * g.world * g.D;
    }
    if (nf > 0) {
      FP.n = nf; FP.a = a; FP.eps = eps; FP.issue_policy = D == 256 ? 1 : 0;
      FP.exp_mode = exp_env >= 0 ? exp_env : kDefaultExpMode(D);
      ProfScope prof(stream, MSF_K_NCE_FLASH, flops);
      if (int rc = (D == 64 ? launch_flash<64>(FP, ctas, st) : D == 128 ? launch_flash<128>(FP, ctas, st) : launch_flash<256>(FP, ctas, st))) return rc;
    }
  }

  // ---- widths above 256: per rank block, P = exp2(a_i q.k - a) (16 bit, + row sums) then O_r = P K_r; rank blocks beyond
  // the resident P budget are walked in rounds that reuse the P slots (stream order makes the reuse safe) ----
  {
    int rounds = 0;
    for (int i = 0; i < n_pairs; ++i) {
      const PairPlan& p = pl.pp[i];
      if (p.mode != 2) continue;
      rounds = std::max(rounds, (pairs[i].world + p.p_ranks - 1) / p.p_ranks);
      if (pairs[i].q_rowsq) {
        nce_inv_norm_kernel<<<(pairs[i].nq + 255) / 256, 256, 0, st>>>(pairs[i].q_rowsq, pairs[i].nq, (pairs[i].D + 63) / 64, eps,
                                                                       reinterpret_cast<float*>(ws + p.off_inv));
        MSF_LAUNCH_OK("nce_inv_norm_kernel");
      }
    }
    for (int round = 0; round < rounds; ++round) {
      std::vector<msf_gemm_problem> g1, g2;
      double flops = 0.0;
      for (int i = 0; i < n_pairs; ++i) {
        const msf_nce_pair& g = pairs[i];
        const PairPlan& p = pl.pp[i];
        if (p.mode != 2) continue;
        float* inv = reinterpret_cast<float*>(ws + p.off_inv);
        for (int r = round * p.p_ranks; r < std::min(g.world, (round + 1) * p.p_ranks); ++r) {
          const char* kr = static_cast<const char*>(g.keys) + static_cast<size_t>(r) * (g.world > 1 ? g.rank_stride : 0) * 2;
          char* Pm = ws + p.off_p + static_cast<size_t>(r - round * p.p_ranks) * g.nq * p.ld_p * 2;
          msf_gemm_problem a1{};
          a1.A = g.q; a1.lda = g.D; a1.B = kr; a1.ldb = g.D; a1.C = Pm; a1.ldc = p.ld_p;
          a1.M = g.nq; a1.N = g.rows_per_rank; a1.K = g.D; a1.out_dtype = MSF_BF16; a1.alpha = 1.f; a1.split_k = -1;
          a1.exp_a = a; a1.row_scale = g.q_rowsq ? inv : nullptr;
          a1.row_sumsq = reinterpret_cast<float*>(ws + p.off_rowsum) + static_cast<size_t>(r) * p.tiles_per_rank * p.nq_pad;
          a1.row_sum_ld = p.nq_pad;
          g1.push_back(a1);
          msf_gemm_problem a2{};
          a2.A = Pm; a2.lda = p.ld_p; a2.B = kr; a2.ldb = g.D; a2.b_is_kn = 1;
          a2.C = reinterpret_cast<float*>(ws + p.off_o) + static_cast<size_t>(r) * p.nq_pad * g.D; a2.ldc = g.D;
          a2.M = g.nq; a2.N = g.D; a2.K = g.rows_per_rank; a2.out_dtype = MSF_F32; a2.alpha = 1.f; a2.split_k = -1;
          g2.push_back(a2);
          flops += 4.0 * g.nq * static_cast<double>(g.rows_per_rank) * g.D;
        }
      }
      if (g1.empty()) continue;
      ProfScope prof(stream, MSF_K_NCE_TWOPASS, flops);
      for (size_t lo = 0; lo < g1.size(); lo += MSF_GEMM_MAX_PROBLEMS) {
        const int cnt = static_cast<int>(std::min<size_t>(MSF_GEMM_MAX_PROBLEMS, g1.size() - lo));
        if (int rc = gemm_grouped_launch(g1.data() + lo, cnt, MSF_BF16, nullptr, 0, nullptr, stream)) return rc;
        // the matching O_r = P K_r problems right behind their P producers: a later chunk of the same round may not reuse a slot
        if (int rc = gemm_grouped_launch(g2.data() + lo, cnt, MSF_BF16, nullptr, 0, nullptr, stream)) return rc;
      }
    }
  }

  // ---- finalize over all pairs + fixed-order final sum ----
  static thread_local PairParams PP;
  int blocks = 0;
  if (int rc = fill_pairs(PP, pairs, n_pairs, pl, ws, tau, eps, 8, &blocks)) return rc;
  for (int i = 0; i < n_pairs; ++i)
    if (pl.pp[i].mode == 2) PP.p[i].splits = pl.pp[i].rs_splits;  // finalize sums the ROW-SUM partials (one per 64-key block)
  MSF_REQUIRE(static_cast<size_t>(blocks) <= pl.partial_cap, MSF_ERR_WORKSPACE, "internal: loss partial capacity");
  float* partials = reinterpret_cast<float*>(ws + pl.off_partials);
  nce_grouped_final_kernel<MSF_BF16><<<blocks, 256, 0, st>>>(PP, partials);
  MSF_LAUNCH_OK("nce_grouped_final_kernel");
  nce_grouped_sum_kernel<<<1, 256, 0, st>>>(partials, static_cast<uint32_t>(blocks), loss_out);
  MSF_LAUNCH_OK("nce_grouped_sum_kernel");
  return MSF_OK;
}

extern "C" int msf_nce_grouped_bwd(const msf_nce_pair* pairs, int n_pairs, int dtype, float tau, float eps, const float* grad_out, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (int rc = check_pairs(pairs, n_pairs, dtype, tau)) return rc;
  MSF_REQUIRE(grad_out, MSF_ERR_INVALID, "grad_out is NULL");
  for (int i = 0; i < n_pairs; ++i) MSF_REQUIRE(pairs[i].grad_q && aligned16(pairs[i].grad_q), MSF_ERR_INVALID, "pair %d: grad_q NULL or misaligned", i);
  const Plan pl = make_plan(pairs, n_pairs);
  MSF_REQUIRE(workspace && workspace_bytes >= pl.total, MSF_ERR_WORKSPACE, "workspace of %zu bytes < %zu required", workspace_bytes, pl.total);
  static thread_local PairParams PP;
  int blocks = 0;
  if (int rc = fill_pairs(PP, pairs, n_pairs, pl, static_cast<char*>(workspace), tau, eps, 8, &blocks)) return rc;
  double bytes = 0.0;
  for (int i = 0; i < n_pairs; ++i) bytes += static_cast<double>(pairs[i].nq) * pairs[i].D * (4.0 * pl.pp[i].splits + 6.0);
  ProfScope prof(stream, MSF_K_NCE_BWD, bytes);
  nce_grouped_bwd_kernel<MSF_BF16><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(PP, grad_out);
  MSF_LAUNCH_OK("nce_grouped_bwd_kernel");
  return MSF_OK;
}
