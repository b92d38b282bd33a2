// G1 (grouped): ONE persistent tcgen05 launch over a table of GEMM problems -- the head stage of
// src/models/backbone.py:12-31, 161-186, 205-212 (12 projectors + 12 predictors x 2 views = up to 24 Linear layers of
// the same depth per launch, forward; their dX and dW GEMMs per launch, backward).
//
//   C_p[M,N] = epi( alpha * pro(A_p)[M,K] * op(B_p) )        for every problem p of the table
//
//   pro (A-operand prologue, optional)   a' = relu?(a * scale[k] + shift[k])  -- the PREVIOUS layer's batch-norm apply
//       + ReLU, done on the shared-memory A tile between the TMA load and the MMA (backbone.py:15-16, 18-19, 28-29), so the
//       normalised activation is never written to / re-read from HBM in the forward;
//   epi (epilogue, all optional)         + bias[n]; round to the output dtype; per-column sum / sum of squares of the
//       ROUNDED outputs per 32-row group (the batch-norm statistics of THIS layer: the reference's BatchNorm1d sees the
//       bf16 Linear output, backbone.py:15); per-row sum of squares per 64-column block (the L2 norms the loss needs,
//       tools/ssl_train.py:422); 16-bit outputs leave through shared memory and a TMA store (coalesced, clipped at the
//       tensor edge), fp32 outputs through 128-byte row segments.
//
// Work unit = (problem, 128-row M tile, bn-column N tile, K split); bn in {64,128,256} per problem so that small-M
// problems (the 4608-wide fuser heads have 256 rows) still spread over the 148 SMs; split-K (deterministic: fp32
// partials + last-arriver reduction in fixed order, no float atomics) for the weight gradients whose K is the row count.
// Persistent CTAs walk the unit list round-robin; the host sorts problems by cost per unit, so the walk is an LPT-like
// schedule, and units of one problem are m-fastest so a weight panel is shared through L2.
//
// Warp roles (384 threads): w0 TMA producer, w1 MMA issuer (one thread), w2 TMEM allocator, w4-7 epilogue (one
// accumulator row per thread), w8-11 A-prologue transform.  4-stage TMA/mbarrier ring (48 KB per stage), accumulators
// double-buffered in TMEM (2 x 256 columns): the epilogue of unit i overlaps the main loop of unit i+1.
// Tensor-bound for the target heads at large batch, weight-streaming (HBM) bound for the fuser heads at small batch:
// 2*M*N*K FLOP; bytes (M*K + N*K)*2 read (re-reads served by L2) + M*N*e written.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "tc_common.cuh"

namespace msf {
int gemm_grouped_launch(const msf_gemm_problem* problems, int n_problems, int op_dtype, void* workspace, size_t workspace_bytes, int32_t* counters,
                        void* stream);
namespace {

using namespace tc;
constexpr int GBK = 64, kStages = 4, kThreads = 384;
constexpr uint32_t kABytes = BM * GBK * 2, kBBytesMax = 256 * GBK * 2, kStageBytes = kABytes + kBBytesMax;
constexpr uint32_t kStoreSlabBytes = BM * 128;  // 128 rows x 64 16-bit columns, SWIZZLE_128B
constexpr uint32_t kBarBytes = 2048;  // barriers (first 256 bytes) + the column-statistics hand-over of the epilogue groups
constexpr uint32_t kSmem = 1024 + kStages * kStageBytes + 2 * kStoreSlabBytes + kBarBytes;
static_assert(kSmem <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");

enum : uint32_t { F_A_MN = 1, F_B_MN = 2, F_OUT_F32 = 4, F_A_RELU = 8, F_TMA_STORE = 16 };

struct alignas(64) DevProblem {
  CUtensorMap ta, tb, tc;   // A and B loads, C store (16-bit outputs)
  void* C;
  const float* bias;        // [N] or null
  float* col_stats;         // [ceil(M/32)][2][N] or null
  float* row_sumsq;         // [ceil(N/64)][M] or null
  const float* a_scale;     // [K] or null (no prologue)
  const float* a_shift;     // [K]
  const float* row_scale;   // EXP epilogue: per-row factor of the exponent (1 / ||q_i||) or null
  float* ws;                // split-K partials [tiles][k_splits][bn/4][128 rows][4] fp32 (a warp's 128-bit accesses are contiguous)
  int32_t* counters;        // [tiles], zero between launches (self-cleaning)
  int64_t ldc;
  int32_t M, N, K, bn, tiles_m, tiles_n, k_splits, kb_per_split, num_kb, unit_start, unit_end;
  uint32_t flags;
  float alpha;
  float exp_a;              // != 0: EXP epilogue  y = exp2(exp_a * row_scale[m] * acc - exp_a), row partials = sums of y (not squares)
  int32_t row_sum_ld;       // leading dimension of row_sumsq (>= M)
  int32_t pad_[3];
};

// The table travels as a kernel parameter (no upload, no fence); three sizes so that a small launch does not push 24 KB
// of parameters through the command buffer.
template <int NP>
struct alignas(64) GroupParams {
  DevProblem p[NP];
  int32_t n_problems, total_units, is_f16;
  int32_t epi2;  // no problem of the launch has an A prologue: warps 8-11 are a SECOND epilogue warpgroup (odd 64-column slabs)
};
static_assert(sizeof(GroupParams<MSF_GEMM_MAX_PROBLEMS>) <= 32000, "kernel parameter space (32764 bytes)");

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// named barriers of the epilogue: 2 / 3 = the 128 threads of epilogue group 0 / 1, 4 = both groups (256 threads, epi2 launches)
__device__ __forceinline__ void epi_bar(int grp) {
  if (grp == 0) asm volatile("bar.sync 2, 128;" ::: "memory");
  else asm volatile("bar.sync 3, 128;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_all(bool both, int grp) {
  if (both) asm volatile("bar.sync 4, 256;" ::: "memory");
  else epi_bar(grp);
}

__device__ __forceinline__ uint32_t idesc_of(int n, bool b_mn, bool a_mn, bool f16) {
  const uint32_t fmt = f16 ? 0u : 1u;  // kind::f16 operand formats: 0 = f16, 1 = bf16
  return (1u << 4) /*D = f32*/ | (fmt << 7) | (fmt << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
}

template <bool F16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (F16) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  }
}
template <bool F16>
__device__ __forceinline__ float2 unpack2(uint32_t w) {
  if constexpr (F16) {
    return __half22float2(*reinterpret_cast<const __half2*>(&w));
  } else {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
  }
}

struct Unit {
  int p, m_tile, n_tile, ks, kb0, kb1;
};
template <int NP>
__device__ __forceinline__ Unit decode(const GroupParams<NP>& P, int u, int& p) {
  while (u >= P.p[p].unit_end) ++p;
  const DevProblem& q = P.p[p];
  const int local = u - q.unit_start;
  Unit r;
  r.p = p;
  r.ks = local % q.k_splits;
  const int t = local / q.k_splits;
  r.m_tile = t % q.tiles_m;
  r.n_tile = t / q.tiles_m;
  r.kb0 = r.ks * q.kb_per_split;
  r.kb1 = min(q.num_kb, r.kb0 + q.kb_per_split);
  return r;
}

// column sums over the 32 rows a warp holds: butterfly transpose-reduce, 31 shuffles for 32 columns; lane L ends up
// with the sum of column L.  `v` is destroyed.
__device__ __forceinline__ float warp_colsum32(float* v, int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <bool F16, int NP>
__global__ void __launch_bounds__(kThreads, 1) gemm_grouped_kernel(const __grid_constant__ GroupParams<NP> P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sC = smem + kStages * kStageBytes;  // 2 store slabs
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + 2 * kStoreSlabBytes);
  // Loads of prologue units complete on `pfull`, all others on `full`: a barrier that is only used by SOME iterations is
  // waited on with the parity of its own use count (bit s of fpar / ppar / xpar), tracked identically by every role
  // because all roles walk the same unit list.  (Skipping waits on a shared barrier would let a role fall a whole phase
  // behind, where the parity test aliases.)  `empty` is used by every iteration: parity from the iteration count.
  uint64_t* full = bars;                 // [kStages] TMA landed (units without an A prologue)
  uint64_t* pfull = full + kStages;      // [kStages] TMA landed (units with an A prologue)
  uint64_t* empty = pfull + kStages;     // [kStages] MMAs that read the stage retired
  uint64_t* xformed = empty + kStages;   // [kStages] A-prologue applied (128 arrivals)
  uint64_t* acc_full = xformed + kStages;  // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  volatile int* last_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
  float* stat_sm = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2 groups][3 warps][2][32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(pfull + s, 1); mbar_init(empty + s, 1); mbar_init(xformed + s, 128); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full + b, 1); mbar_init(acc_empty + b, P.epi2 ? 256 : 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      int p = 0;
      for (int u = blockIdx.x; u < P.total_units; u += gridDim.x) {
        const Unit w = decode(P, u, p);
        const DevProblem& q = P.p[w.p];
        const int m0 = w.m_tile * BM, n0 = w.n_tile * q.bn;
        const uint32_t bytes = kABytes + static_cast<uint32_t>(q.bn) * GBK * 2;
        uint64_t* landed = q.a_scale ? pfull : full;
        for (int kb = w.kb0; kb < w.kb1; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(empty + s, ((it / kStages) & 1) ^ 1);
          mbar_expect_tx(landed + s, bytes);
          uint8_t* sa = smem + s * kStageBytes;
          uint8_t* sb = sa + kABytes;
          if (!(q.flags & F_A_MN)) {
            tma_load_2d(sa, &q.ta, kb * GBK, m0, landed + s);           // one 64(k) x 128(m) box, rows = m
          } else {
            for (int j = 0; j < BM / 64; ++j)                            // two 64(m) x 64(k) boxes, rows = k
              tma_load_2d(sa + j * (GBK * 128), &q.ta, m0 + j * 64, kb * GBK, landed + s);
          }
          if (!(q.flags & F_B_MN)) {
            tma_load_2d(sb, &q.tb, kb * GBK, n0, landed + s);           // one 64(k) x bn(n) box, rows = n
          } else {
            for (int j = 0; j < q.bn / 64; ++j)                          // bn/64 boxes of 64(n) x 64(k), rows = k
              tma_load_2d(sb + j * (GBK * 128), &q.tb, n0 + j * 64, kb * GBK, landed + s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      uint32_t it = 0, acc_it = 0, xpar = 0, fpar = 0;  // phase parities of xformed[s] / full[s] in bit s
      int p = 0;
      for (int u = blockIdx.x; u < P.total_units; u += gridDim.x, ++acc_it) {
        const Unit w = decode(P, u, p);
        const DevProblem& q = P.p[w.p];
        const bool a_mn = (q.flags & F_A_MN) != 0, b_mn = (q.flags & F_B_MN) != 0, pro = q.a_scale != nullptr;
        const uint32_t idesc = idesc_of(q.bn, b_mn, a_mn, F16);
        const uint32_t buf = acc_it & 1;
        mbar_wait(acc_empty + buf, ((acc_it >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem + buf * 256;
        for (int kb = w.kb0; kb < w.kb1; ++kb, ++it) {
          const int s = it % kStages;
          if (pro) {
            mbar_wait(xformed + s, (xpar >> s) & 1);
            xpar ^= 1u << s;
          } else {
            mbar_wait(full + s, (fpar >> s) & 1);
            fpar ^= 1u << s;
          }
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * kStageBytes), b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < GBK / 16; ++k) {
            const uint64_t ad = a_mn ? umma_desc(a_addr + k * 2048, GBK * 128, 1024) : umma_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bd = b_mn ? umma_desc(b_addr + k * 2048, GBK * 128, 1024) : umma_desc(b_addr + k * 32, 16, 1024);
            mma_ss(d_tmem, ad, bd, idesc, (kb > w.kb0 || k > 0));
          }
          tc_commit(empty + s);
        }
        tc_commit(acc_full + buf);
      }
    }
  } else if (warp >= 8 && !P.epi2) {
    // ===================== A-prologue: a' = relu?(a * scale[k] + shift[k]) on the landed A tile =====================
    const int r = ((warp - 8) << 5) + lane;  // row of the 128 x 64 K-major tile: 128 bytes, 16-byte chunks XOR-swizzled by (r & 7)
    uint32_t it = 0, ppar = 0;
    int p = 0;
    for (int u = blockIdx.x; u < P.total_units; u += gridDim.x) {
      const Unit w = decode(P, u, p);
      const DevProblem& q = P.p[w.p];
      if (q.a_scale == nullptr) { it += static_cast<uint32_t>(w.kb1 - w.kb0); continue; }
      const bool relu = (q.flags & F_A_RELU) != 0;
      const float* const a_scale = q.a_scale;
      const float* const a_shift = q.a_shift;
      const int kdim = q.K;
      for (int kb = w.kb0; kb < w.kb1; ++kb, ++it) {
        const int s = it % kStages;
        mbar_wait(pfull + s, (ppar >> s) & 1);
        ppar ^= 1u << s;
        uint8_t* rowp = smem + s * kStageBytes + r * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k0 = kb * GBK + j * 8;
          if (k0 >= kdim) break;  // K is a multiple of 8: whole chunks; columns past K stay zero (TMA fill)
          uint4* cp = reinterpret_cast<uint4*>(rowp + ((j ^ (r & 7)) << 4));
          const uint4 v = *cp;
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(a_scale + k0)), s1 = __ldg(reinterpret_cast<const float4*>(a_scale + k0 + 4));
          const float4 h0 = __ldg(reinterpret_cast<const float4*>(a_shift + k0)), h1 = __ldg(reinterpret_cast<const float4*>(a_shift + k0 + 4));
          const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
          const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack2<F16>(wv[e]);
            // the reference's BatchNorm1d output is a 16-bit tensor and its ReLU acts on that: round first, clamp after
            const float2 g = unpack2<F16>(pack2<F16>(fmaf(f.x, sc[2 * e], sh[2 * e]), fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1])));
            o[e] = relu ? pack2<F16>(fmaxf(g.x, 0.f), fmaxf(g.y, 0.f)) : pack2<F16>(g.x, g.y);
          }
          *cp = make_uint4(o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        mbar_arrive(xformed + s);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (one accumulator row per thread) =====================
    // Launches without an A prologue (every backward launch: dX and dW, whose short K = rows makes the epilogue of a tile
    // outlast its MMAs) run TWO epilogue warpgroups: group g = warps 4+4g .. 7+4g takes the 64-column slabs with (slab & 1) == g
    // -- a warp may only read the TMEM lanes 32 * (warp % 4) .., which both groups' warps cover -- and stores through its own
    // shared-memory slab.  Split-K bookkeeping synchronises both groups.
    const bool epi2 = P.epi2 != 0;
    const int grp = warp >= 8 ? 1 : 0;
    const int wq = warp & 3;
    const int row_in_tile = (wq << 5) + lane;
    const int etid = row_in_tile;  // 0..127
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(wq << 5) << 16);
    uint32_t acc_it = 0, slab_it = 0;
    int p = 0;
    for (int u = blockIdx.x; u < P.total_units; u += gridDim.x, ++acc_it) {
      const Unit w = decode(P, u, p);
      const DevProblem& q = P.p[w.p];
      const uint32_t buf = acc_it & 1;
      const int m0 = w.m_tile * BM, n0 = w.n_tile * q.bn;
      const int64_t row = static_cast<int64_t>(m0) + row_in_tile;
      const bool row_ok = row < q.M;
      const int n_chunks = q.bn / 32;
      const int tile_idx = w.n_tile * q.tiles_m + w.m_tile;
      mbar_wait(acc_full + buf, (acc_it >> 1) & 1);
      tc_fence_after();

      bool from_ws = false;
      if (q.k_splits > 1) {
        // ---- split-K: park the fp32 partial; the last CTA to arrive for this tile reduces all of them in split order ----
        // layout [column quad][row][4]: the 32 rows of a warp write 512 contiguous bytes per store instruction
        float* part = q.ws + (static_cast<size_t>(tile_idx) * q.k_splits + w.ks) * BM * q.bn + row_in_tile * 4;
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c) {
          if (epi2 && ((c >> 1) & 1) != grp) continue;
          uint32_t v[32];
          tmem_ld32(lane_base + buf * 256 + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4) __stcg(reinterpret_cast<uint4*>(part + (c * 8 + (i >> 2)) * (BM * 4)), make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]));
        }
        tc_fence_before();
        mbar_arrive(acc_empty + buf);  // TMEM buffer is free again
        __threadfence();
        epi_bar_all(epi2, grp);
        if (etid == 0 && grp == 0) {
          const int old = atomicAdd(q.counters + tile_idx, 1);
          const int last = old == q.k_splits - 1;
          if (last) q.counters[tile_idx] = 0;  // self-cleaning: the next launch finds zeros
          *last_flag = last;
        }
        epi_bar_all(epi2, grp);
        const bool last = *last_flag != 0;
        epi_bar_all(epi2, grp);  // everyone has read the flag before a later unit overwrites it
        if (!last) continue;
        __threadfence();
        from_ws = true;
      }

      const bool tma_store = (q.flags & F_TMA_STORE) != 0;
      const bool out_f32 = (q.flags & F_OUT_F32) != 0;
      float rq = 0.f;  // row sum of squares of the current 64-column block
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        if (epi2 && ((c >> 1) & 1) != grp) continue;
        float x[32];
        if (!from_ws) {
          uint32_t v[32];
          tmem_ld32(lane_base + buf * 256 + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(v[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = 0.f;
          const int ks_n = q.k_splits, bn_w = q.bn;
          const float* const ws0 = q.ws;
          for (int s = 0; s < ks_n; ++s) {  // fixed order: deterministic
            const float* part = ws0 + (static_cast<size_t>(tile_idx) * ks_n + s) * BM * bn_w + row_in_tile * 4 + c * 8 * (BM * 4);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 t = __ldcg(reinterpret_cast<const float4*>(part + (i >> 2) * (BM * 4)));
              x[i] += t.x; x[i + 1] += t.y; x[i + 2] += t.z; x[i + 3] += t.w;
            }
          }
        }
        const int col = n0 + c * 32;  // first global column of this chunk
        // ---- alpha, bias (or the InfoNCE exponential), rounding to the output dtype ----
        if (q.exp_a != 0.f) {
          const float exp_a = q.exp_a;
          const float ea = exp_a * (q.row_scale ? (row_ok ? __ldg(q.row_scale + row) : 0.f) : 1.f);
          const int ncol = q.N - col;
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] = i < ncol ? ex2_approx(fmaf(x[i], ea, -exp_a)) : 0.f;  // |cos| <= 1: never overflows
        } else {
          // q.* are indexed loads from the kernel-parameter space: read once per chunk, not once per element (with the
          // per-element form the compiler kept 32 predicates + addresses live: 167 registers and an 11 us epilogue per
          // 128x256 tile -- 3.5x the tile's MMA time at K = 512; now 133 registers and the epilogue hides behind the MMAs)
          const float alpha = q.alpha;
          const float* const bias = q.bias;
#pragma unroll
          for (int i = 0; i < 32; ++i) x[i] *= alpha;
          if (bias) {
            const int ncol = q.N - col;  // valid columns of this chunk
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < ncol) x[i] += __ldg(bias + col + i);
          }
        }
        uint32_t u16[16];
        if (!out_f32) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            u16[i >> 1] = pack2<F16>(x[i], x[i + 1]);
            const float2 rr = unpack2<F16>(u16[i >> 1]);  // statistics are taken on what the next layer will read
            x[i] = rr.x;
            x[i + 1] = rr.y;
          }
        }
        // ---- row sum of squares per 64-column block (columns >= N hold exact zeros: zero-filled B rows, no bias) ----
        if (q.row_sumsq) {
          if (q.exp_a != 0.f) {
#pragma unroll
            for (int i = 0; i < 32; ++i) rq += x[i];
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) rq = fmaf(x[i], x[i], rq);
          }
          if ((c & 1) == 1 || c == n_chunks - 1) {
            const int blk = (n0 >> 6) + (c >> 1);
            if (row_ok && blk * 64 < q.N) q.row_sumsq[static_cast<size_t>(blk) * q.row_sum_ld + row] = rq;
            rq = 0.f;
          }
        }
        // ---- store ----
        if (tma_store) {
          const int slab = c >> 1, half = c & 1;
          // one group: the two buffers alternate; two groups: each owns one buffer and waits for its previous store
          uint8_t* sbuf = sC + (epi2 ? grp : ((slab_it + slab) & 1)) * kStoreSlabBytes;
          if (half == 0) {
            if (etid == 0) {  // the store that last read this buffer is done with it
              if (epi2) bulk_wait_read<0>();
              else bulk_wait_read<1>();
            }
            epi_bar(grp);
          }
          uint8_t* rowp = sbuf + row_in_tile * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(rowp + (((half * 4 + j) ^ (row_in_tile & 7)) << 4)) = make_uint4(u16[4 * j], u16[4 * j + 1], u16[4 * j + 2], u16[4 * j + 3]);
          if (half == 1 || c == n_chunks - 1) {
            fence_proxy_async_smem();
            epi_bar(grp);
            if (etid == 0 && n0 + slab * 64 < q.N) {
              tma_store_2d(&q.tc, sbuf, n0 + slab * 64, m0);  // rows >= M and columns >= N are clipped by the tensor map
              bulk_commit();
            }
          }
        } else if (row_ok && col < q.N) {
          const bool full_chunk = col + 32 <= q.N;
          if (out_f32) {
            float* dst = static_cast<float*>(q.C) + row * q.ldc + col;
            if (full_chunk && (q.ldc & 3) == 0) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
            } else {
              for (int i = 0; i < 32 && col + i < q.N; ++i) dst[i] = x[i];
            }
          } else {
            uint16_t* dst = static_cast<uint16_t*>(q.C) + row * q.ldc + col;
            if (full_chunk && (q.ldc & 7) == 0) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) *reinterpret_cast<uint4*>(dst + 2 * i) = make_uint4(u16[i], u16[i + 1], u16[i + 2], u16[i + 3]);
            } else {
              for (int i = 0; i < 32 && col + i < q.N; ++i) dst[i] = reinterpret_cast<const uint16_t*>(u16)[i];
            }
          }
        }
        // ---- batch-norm statistics of this layer: column sum / sum of squares over the warp's 32 rows ----
        if (q.col_stats) {
          float s1[32], s2[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float t = row_ok ? x[i] : 0.f;
            s1[i] = t;
            s2[i] = t * t;
          }
          const float cs = warp_colsum32(s1, lane), cq = warp_colsum32(s2, lane);
          // one entry per 128-row tile: warps 1..3 hand their 32-row sums to warp 0 of the group (added in warp order), so the
          // finalize kernel merges rows / 128 groups per column instead of rows / 32 (its per-column walk is a latency chain)
          float* sx = stat_sm + grp * (3 * 2 * 32);
          if (wq > 0) {
            sx[((wq - 1) * 2) * 32 + lane] = cs;
            sx[((wq - 1) * 2 + 1) * 32 + lane] = cq;
          }
          epi_bar(grp);
          if (wq == 0) {
            float a = cs, b = cq;
#pragma unroll
            for (int w2 = 0; w2 < 3; ++w2) {
              a += sx[(w2 * 2) * 32 + lane];
              b += sx[(w2 * 2 + 1) * 32 + lane];
            }
            if (col + lane < q.N) {
              float* dst = q.col_stats + (static_cast<size_t>(w.m_tile) * 2) * q.N + col + lane;
              dst[0] = a;
              dst[q.N] = b;
            }
          }
          epi_bar(grp);  // the hand-over slots are reused by the next chunk
        }
      }
      if (tma_store) slab_it += static_cast<uint32_t>((n_chunks + 1) >> 1);
      if (!from_ws) {
        tc_fence_before();
        mbar_arrive(acc_empty + buf);
      }
    }
    if (etid == 0) bulk_wait_read<0>();  // shared memory must outlive the last TMA store's reads
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

// ---- host: tensor-map cache ------------------------------------------------------------------------------------
// Encoding a map costs a driver call; the head weights never move and the caching allocator hands the activations the same
// addresses every step, so maps are memoised on everything that defines them.  (Immutable once built; guarded by a mutex.)
struct MapKey {
  const void* base;
  int64_t rows, cols, ld;
  uint32_t box_cols, box_rows, f16;
  bool operator==(const MapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_cols == o.box_cols && box_rows == o.box_rows && f16 == o.f16;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.base) * 0x9E3779B97F4A7C15ull;
    for (uint64_t v : {static_cast<uint64_t>(k.rows), static_cast<uint64_t>(k.cols), static_cast<uint64_t>(k.ld),
                       (static_cast<uint64_t>(k.box_cols) << 32) | (static_cast<uint64_t>(k.box_rows) << 1) | k.f16})
      h = (h ^ v) * 0x9E3779B97F4A7C15ull + (h >> 29);
    return static_cast<size_t>(h);
  }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

int cached_map(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, uint32_t box_cols, uint32_t box_rows, bool f16) {
  const MapKey key{base, rows, cols, ld, box_cols, box_rows, f16 ? 1u : 0u};
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return MSF_OK; }
  }
  if (int rc = make_map_16(out, base, rows, cols, ld, box_cols, box_rows, f16)) return rc;
  std::lock_guard<std::mutex> lk(g_map_mu);
  if (g_maps.size() > 8192) g_maps.clear();
  g_maps.emplace(key, *out);
  return MSF_OK;
}

struct Plan {
  int bn, tiles_m, tiles_n, num_kb, k_splits, kb_per_split, tiles;
  size_t ws_floats;
  double cost;  // per unit
};

Plan plan_of(const msf_gemm_problem& g) {
  Plan pl{};
  pl.tiles_m = (g.M + BM - 1) / BM;
  pl.bn = (g.tile_n == 64 || g.tile_n == 128 || g.tile_n == 256) ? g.tile_n : 64;  // plan_launch chooses; 64 only as a fallback
  pl.tiles_n = (g.N + pl.bn - 1) / pl.bn;
  pl.tiles = pl.tiles_m * pl.tiles_n;
  pl.num_kb = (g.K + GBK - 1) / GBK;
  pl.k_splits = 1;
  if (g.split_k > 0) pl.k_splits = g.split_k;
  if (pl.k_splits > 32) pl.k_splits = 32;
  if (pl.k_splits > pl.num_kb) pl.k_splits = pl.num_kb;
  pl.kb_per_split = (pl.num_kb + pl.k_splits - 1) / pl.k_splits;
  pl.k_splits = (pl.num_kb + pl.kb_per_split - 1) / pl.kb_per_split;  // no empty split
  pl.ws_floats = pl.k_splits > 1 ? static_cast<size_t>(pl.tiles) * pl.k_splits * BM * pl.bn : 0;
  pl.cost = static_cast<double>(pl.kb_per_split) * (BM + pl.bn) + 2.0 * pl.bn;  // bytes staged per unit + epilogue
  return pl;
}

// Launch-wide plan.  Work is counted in shared-memory rows of 128 bytes, which is what bounds this kernel (L2 -> SM feed and
// its latency): a k-block stages 128 + bn rows, a split-K partial is 4 * bn rows written by every unit and ks * 4 * bn rows
// read back by the last arriver, and every unit pays ~3 k-blocks of pipeline fill / epilogue.  A WIDE tile stages the fewest
// bytes per MAC (a 256-row fuser problem with bn = 64 re-reads its A panel 72 times), so the plan prefers splitting K over
// narrowing the tile:
//   one problem    every (bn, ks) is scored by the makespan  ceil(units / 148) * unit_cost + reduction  and the best wins;
//   many problems  the units share the 148 SMs, so the launch takes about max(total / 148, longest unit): per problem the
//                  (bn, ks) with the least total cost among those whose unit (+ reduction) fits under that balanced share.
struct Choice {
  int bn, ks;
};
inline double fill_rows() { return 3.0 * (BM + 256); }
inline int kb_per_split_of(int num_kb, int ks) { return (num_kb + ks - 1) / ks; }
inline double unit_rows(int num_kb, int bn, int ks) {
  return static_cast<double>(kb_per_split_of(num_kb, ks)) * (BM + bn) + fill_rows() + (ks > 1 ? 4.0 * bn : 0.0);
}
// (A fixed ~10 us penalty per split -- what the warm-L2 sweep of tools/diag/gemm_plan_sweep.py suggests -- was tried and made
// the training step slower: in the step the weights come from HBM, where more concurrent units win: 0.953 vs 0.899 ms of GEMM
// time per 256-tile step.)
inline double reduce_rows(int bn, int ks) { return ks > 1 ? 4.0 * bn * ks : 0.0; }
inline int max_splits(int num_kb) { return std::max(1, std::min(32, num_kb / 4)); }  // at least 4 k-blocks per split

void plan_launch(const msf_gemm_problem* problems, int n, Plan* plans) {
  double work = 0.0;
  for (int i = 0; i < n; ++i) {
    const msf_gemm_problem& g = problems[i];
    const int tiles_m = (g.M + BM - 1) / BM, num_kb = (g.K + GBK - 1) / GBK;
    const int bn = g.N <= 64 ? 64 : (g.N <= 128 ? 128 : 256);
    work += static_cast<double>(tiles_m) * ((g.N + bn - 1) / bn) * unit_rows(num_kb, bn, 1);
  }
  const double cap = std::max(work / kNumSMs, unit_rows(24, 256, 1));
  for (int i = 0; i < n; ++i) {
    msf_gemm_problem g = problems[i];
    const int num_kb = (g.K + GBK - 1) / GBK, tiles_m = (g.M + BM - 1) / BM;
    const bool bn_forced = g.tile_n == 64 || g.tile_n == 128 || g.tile_n == 256;
    const bool ks_forced = g.split_k != 0;
    Choice best{bn_forced ? g.tile_n : 64, ks_forced ? std::max(1, g.split_k) : 1};
    double best_score = 1e300;
    bool best_fits = false;
    for (int bn : {256, 128, 64}) {
      if (bn_forced ? bn != g.tile_n : (bn > 64 && bn / 2 >= g.N)) continue;  // a tile twice as wide as the problem only stages zeros
      const int64_t tiles = static_cast<int64_t>(tiles_m) * ((g.N + bn - 1) / bn);
      const int ks_hi = ks_forced ? std::max(1, g.split_k) : max_splits(num_kb);
      for (int ks = ks_forced ? ks_hi : 1; ks <= ks_hi; ++ks) {
        const int kbps = kb_per_split_of(num_kb, ks);
        const int ks_eff = (num_kb + kbps - 1) / kbps;  // no empty split
        if (!ks_forced && ks_eff != ks) continue;
        const double unit = unit_rows(num_kb, bn, ks_eff), red = reduce_rows(bn, ks_eff);
        const double units = static_cast<double>(tiles) * ks_eff;
        double score;
        bool fits = true;
        if (n == 1) {
          score = std::ceil(units / kNumSMs) * unit + red;  // makespan of the launch
        } else {
          fits = unit + red <= cap;
          score = fits ? units * unit + red * tiles : unit + red;  // least total work among the fitting plans, else the shortest unit
        }
        if ((fits && !best_fits) || (fits == best_fits && score < best_score * 0.999)) {
          best = Choice{bn, ks_eff};
          best_score = score;
          best_fits = fits;
        }
      }
    }
    g.tile_n = best.bn;
    g.split_k = best.ks > 1 ? best.ks : -1;
    plans[i] = plan_of(g);
  }
}

int check_problem(const msf_gemm_problem& g, int i, int op_dtype) {
  MSF_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, MSF_ERR_INVALID, "problem %d: empty GEMM %d x %d x %d", i, g.M, g.N, g.K);
  MSF_REQUIRE(g.A && g.B && g.C && aligned16(g.A) && aligned16(g.B) && aligned16(g.C), MSF_ERR_INVALID, "problem %d: NULL or misaligned operand", i);
  MSF_REQUIRE(g.lda % 8 == 0 && g.ldb % 8 == 0, MSF_ERR_INVALID, "problem %d: lda / ldb must be multiples of 8 elements (TMA row pitch)", i);
  MSF_REQUIRE(!g.a_scale || g.K % 8 == 0, MSF_ERR_INVALID, "problem %d: K = %d must be a multiple of 8 for the A prologue", i, g.K);
  MSF_REQUIRE(g.exp_a == 0.f || (!g.bias && g.out_dtype != MSF_F32), MSF_ERR_UNSUPPORTED, "problem %d: the EXP epilogue has 16-bit outputs and no bias", i);
  MSF_REQUIRE(g.row_sum_ld == 0 || g.row_sum_ld >= g.M, MSF_ERR_INVALID, "problem %d: row_sum_ld < M", i);
  MSF_REQUIRE(g.out_dtype == MSF_F32 || g.out_dtype == op_dtype, MSF_ERR_INVALID, "problem %d: output dtype must be MSF_F32 or the operand dtype", i);
  MSF_REQUIRE(!(g.a_scale && g.a_is_km), MSF_ERR_UNSUPPORTED, "problem %d: the A prologue needs A stored [M][K]", i);
  MSF_REQUIRE(!g.a_scale || (g.a_shift && aligned16(g.a_scale) && aligned16(g.a_shift)), MSF_ERR_INVALID, "problem %d: a_scale / a_shift must both be set, 16-byte aligned", i);
  return MSF_OK;
}

}  // namespace
}  // namespace msf

using namespace msf;

namespace msf {
namespace {
template <bool F16, int NP>
int launch_one(const GroupParams<NP>& Q, unsigned grid, cudaStream_t st) {
  static const cudaError_t attr = cudaFuncSetAttribute(gemm_grouped_kernel<F16, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  MSF_REQUIRE(attr == cudaSuccess, MSF_ERR_CUDA, "cudaFuncSetAttribute(max dynamic shared memory) failed: %s", cudaGetErrorString(attr));
  gemm_grouped_kernel<F16, NP><<<grid, kThreads, kSmem, st>>>(Q);
  MSF_LAUNCH_OK("gemm_grouped_kernel");
  return MSF_OK;
}
template <int NP>
int launch_np(const GroupParams<MSF_GEMM_MAX_PROBLEMS>& P, bool f16, unsigned grid, cudaStream_t st) {
  if constexpr (NP == MSF_GEMM_MAX_PROBLEMS) {
    return f16 ? launch_one<true, NP>(P, grid, st) : launch_one<false, NP>(P, grid, st);
  } else {
    static thread_local GroupParams<NP> Q;
    for (int i = 0; i < P.n_problems; ++i) Q.p[i] = P.p[i];
    Q.n_problems = P.n_problems; Q.total_units = P.total_units; Q.is_f16 = P.is_f16; Q.epi2 = P.epi2;
    return f16 ? launch_one<true, NP>(Q, grid, st) : launch_one<false, NP>(Q, grid, st);
  }
}
}  // namespace
}  // namespace msf

extern "C" size_t msf_gemm_grouped_workspace_bytes(const msf_gemm_problem* problems, int n_problems) {
  if (!problems || n_problems <= 0) return 0;
  if (n_problems > MSF_GEMM_MAX_PROBLEMS) return 0;
  Plan plans[MSF_GEMM_MAX_PROBLEMS];
  plan_launch(problems, n_problems, plans);
  size_t fl = 0;
  for (int i = 0; i < n_problems; ++i) fl += plans[i].ws_floats;
  return fl * sizeof(float) + 256;
}

extern "C" int msf_gemm_grouped(const msf_gemm_problem* problems, int n_problems, int op_dtype, void* workspace, size_t workspace_bytes,
                                int32_t* counters, void* stream) {
  return gemm_grouped_launch(problems, n_problems, op_dtype, workspace, workspace_bytes, counters, stream);
}

int msf::gemm_grouped_launch(const msf_gemm_problem* problems, int n_problems, int op_dtype, void* workspace, size_t workspace_bytes,
                             int32_t* counters, void* stream) {
  MSF_REQUIRE(problems && n_problems > 0 && n_problems <= MSF_GEMM_MAX_PROBLEMS, MSF_ERR_INVALID, "n_problems %d outside [1, %d]", n_problems,
              MSF_GEMM_MAX_PROBLEMS);
  MSF_REQUIRE(op_dtype == MSF_BF16 || op_dtype == MSF_F16, MSF_ERR_INVALID, "operand dtype must be MSF_BF16 or MSF_F16");
  const bool f16 = op_dtype == MSF_F16;
  std::vector<Plan> plans(n_problems);
  std::vector<int> order(n_problems);
  for (int i = 0; i < n_problems; ++i) {
    if (int rc = check_problem(problems[i], i, op_dtype)) return rc;
    order[i] = i;
  }
  plan_launch(problems, n_problems, plans.data());
  // most expensive units first (stable: problems sharing a weight panel stay adjacent), walked round-robin by the CTAs
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return plans[a].cost > plans[b].cost; });
  static thread_local GroupParams<MSF_GEMM_MAX_PROBLEMS> P;  // 24 KB: not on the stack of the autograd thread
  size_t ws_off = 0;
  int ctr_off = 0, unit = 0;
  double flops = 0.0;
  for (int oi = 0; oi < n_problems; ++oi) {
    const msf_gemm_problem& g = problems[order[oi]];
    const Plan& pl = plans[order[oi]];
    DevProblem& q = P.p[oi];
    q = DevProblem{};
    if (g.a_is_km) {
      if (int rc = cached_map(&q.ta, g.A, g.K, g.M, g.lda, 64, GBK, f16)) return rc;
    } else {
      if (int rc = cached_map(&q.ta, g.A, g.M, g.K, g.lda, GBK, BM, f16)) return rc;
    }
    if (g.b_is_kn) {
      if (int rc = cached_map(&q.tb, g.B, g.K, g.N, g.ldb, 64, GBK, f16)) return rc;
    } else {
      if (int rc = cached_map(&q.tb, g.B, g.N, g.K, g.ldb, GBK, static_cast<uint32_t>(pl.bn), f16)) return rc;
    }
    q.flags = (g.a_is_km ? F_A_MN : 0u) | (g.b_is_kn ? F_B_MN : 0u) | (g.out_dtype == MSF_F32 ? F_OUT_F32 : 0u) | (g.a_relu ? F_A_RELU : 0u);
    const bool can_tma_store = g.out_dtype != MSF_F32 && g.ldc % 8 == 0 && !g.no_tma_store;
    if (can_tma_store) {
      if (int rc = cached_map(&q.tc, g.C, g.M, g.N, g.ldc, 64, BM, f16)) return rc;
      q.flags |= F_TMA_STORE;
    }
    q.C = g.C; q.bias = g.bias; q.col_stats = g.col_stats; q.row_sumsq = g.row_sumsq; q.a_scale = g.a_scale; q.a_shift = g.a_shift;
    q.ldc = g.ldc; q.M = g.M; q.N = g.N; q.K = g.K; q.bn = pl.bn; q.tiles_m = pl.tiles_m; q.tiles_n = pl.tiles_n;
    q.k_splits = pl.k_splits; q.kb_per_split = pl.kb_per_split; q.num_kb = pl.num_kb; q.alpha = g.alpha;
    q.row_scale = g.row_scale; q.exp_a = g.exp_a; q.row_sum_ld = g.row_sum_ld > 0 ? g.row_sum_ld : g.M;
    if (pl.k_splits > 1) {
      MSF_REQUIRE(workspace && aligned16(workspace) && counters, MSF_ERR_WORKSPACE, "split-K needs a workspace and the counter array");
      MSF_REQUIRE((ws_off + pl.ws_floats) * sizeof(float) <= workspace_bytes, MSF_ERR_WORKSPACE, "workspace of %zu bytes too small", workspace_bytes);
      MSF_REQUIRE(ctr_off + pl.tiles <= MSF_GEMM_MAX_COUNTERS, MSF_ERR_UNSUPPORTED, "more than %d split-K tiles in one launch", MSF_GEMM_MAX_COUNTERS);
      q.ws = static_cast<float*>(workspace) + ws_off;
      q.counters = counters + ctr_off;
      ws_off += pl.ws_floats;
      ctr_off += pl.tiles;
    }
    q.unit_start = unit;
    unit += pl.tiles * pl.k_splits;
    q.unit_end = unit;
    flops += 2.0 * g.M * g.N * static_cast<double>(g.K);
  }
  P.n_problems = n_problems;
  P.total_units = unit;
  P.is_f16 = f16 ? 1 : 0;
  static const bool epi2_off = getenv("MSF_GEMM_NO_EPI2") != nullptr;  // measurements
  P.epi2 = 1;
  for (int i = 0; i < n_problems; ++i)
    if (problems[i].a_scale) P.epi2 = 0;
  if (epi2_off) P.epi2 = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>(unit < kNumSMs ? unit : kNumSMs);
  ProfScope prof(stream, MSF_K_GEMM_GROUPED, flops);
  if (n_problems <= 4) return launch_np<4>(P, f16, grid, st);
  if (n_problems <= 16) return launch_np<16>(P, f16, grid, st);
  return launch_np<MSF_GEMM_MAX_PROBLEMS>(P, f16, grid, st);
}

extern "C" int msf_gemm_grouped_plan_info(const msf_gemm_problem* problem, int32_t* info) {
  MSF_REQUIRE(problem && info, MSF_ERR_INVALID, "NULL argument");
  Plan pl;
  plan_launch(problem, 1, &pl);
  info[0] = pl.bn; info[1] = pl.tiles_m; info[2] = pl.tiles_n; info[3] = pl.k_splits; info[4] = pl.kb_per_split; info[5] = pl.num_kb;
  return MSF_OK;
}

// y = x W^T with the batch-norm statistics of y in the epilogue (SURVEY 8b `linear_bnstat`): the single-problem form.
extern "C" int msf_linear_bnstat(const void* x, const void* w, void* y, int64_t rows, int in_features, int out_features, int dtype,
                                 float* col_stats, const float* a_scale, const float* a_shift, int a_relu, void* stream) {
  MSF_REQUIRE(rows > 0 && rows < (1ll << 31), MSF_ERR_INVALID, "rows out of range");
  msf_gemm_problem g{};
  g.A = x; g.lda = in_features; g.B = w; g.ldb = in_features; g.C = y; g.ldc = out_features;
  g.M = static_cast<int32_t>(rows); g.N = out_features; g.K = in_features; g.out_dtype = dtype; g.alpha = 1.f;
  g.col_stats = col_stats; g.a_scale = a_scale; g.a_shift = a_shift; g.a_relu = a_relu; g.split_k = -1;
  return msf_gemm_grouped(&g, 1, dtype, nullptr, 0, nullptr, stream);
}

// ------------------------------------------------------------------------------------------------------------------
// fp32 SIMT grouped GEMM: the exact-arithmetic path of the heads (fp32 parameters and activations outside autocast; the
// <= 1e-5 parity cases).  Same problem table, fp32 operands, plain FMA accumulation in k order (deterministic), no tensor
// cores -- kind::tf32 would round the operands to 10 mantissa bits.  64 x 64 tile, 256 threads, 4 x 4 outputs per thread.
// Not a performance path (the training recipes run under 16-bit autocast).
// ------------------------------------------------------------------------------------------------------------------
namespace msf {
namespace {

struct SimtProblem {
  const float* A; const float* B; float* C; const float* bias;
  int64_t sam, sak, sbn, sbk, ldc;  // element strides: A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk]
  int32_t M, N, K, tiles_m, tile_start;
  float alpha;
};
struct SimtParams {
  SimtProblem p[MSF_GEMM_MAX_PROBLEMS];
  int32_t n, total;
};

__global__ void __launch_bounds__(256) gemm_simt_grouped_kernel(const __grid_constant__ SimtParams P) {
  __shared__ float As[16][64 + 4], Bs[16][64 + 4];
  int pi = 0;
#pragma unroll 1
  while (pi + 1 < P.n && static_cast<int>(blockIdx.x) >= P.p[pi + 1].tile_start) ++pi;
  const SimtProblem& q = P.p[pi];
  const int t = blockIdx.x - q.tile_start;
  const int m0 = (t % q.tiles_m) * 64, n0 = (t / q.tiles_m) * 64;
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < q.K; k0 += 16) {
    for (int i = tid; i < 64 * 16; i += 256) {
      // consecutive threads walk the contiguous dimension of each operand
      const int ra = q.sak == 1 ? i / 16 : i % 64, ka = q.sak == 1 ? i % 16 : i / 64;
      As[ka][ra] = (m0 + ra < q.M && k0 + ka < q.K) ? q.A[(m0 + ra) * q.sam + (k0 + ka) * q.sak] : 0.f;
      const int rb = q.sbk == 1 ? i / 16 : i % 64, kb = q.sbk == 1 ? i % 16 : i / 64;
      Bs[kb][rb] = (n0 + rb < q.N && k0 + kb < q.K) ? q.B[(n0 + rb) * q.sbn + (k0 + kb) * q.sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= q.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < q.N) q.C[m * q.ldc + n] = acc[i][j] * q.alpha + (q.bias ? q.bias[n] : 0.f);
    }
  }
}

}  // namespace
}  // namespace msf

extern "C" int msf_gemm_grouped_f32(const msf_gemm_problem* problems, int n_problems, void* stream) {
  MSF_REQUIRE(problems && n_problems > 0 && n_problems <= MSF_GEMM_MAX_PROBLEMS, MSF_ERR_INVALID, "n_problems %d outside [1, %d]", n_problems,
              MSF_GEMM_MAX_PROBLEMS);
  static thread_local SimtParams P;
  int total = 0;
  double flops = 0.0;
  for (int i = 0; i < n_problems; ++i) {
    const msf_gemm_problem& g = problems[i];
    MSF_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0 && g.A && g.B && g.C, MSF_ERR_INVALID, "problem %d: empty or NULL", i);
    MSF_REQUIRE(g.out_dtype == MSF_F32 && !g.a_scale && !g.col_stats && !g.row_sumsq, MSF_ERR_UNSUPPORTED,
                "problem %d: the fp32 path has fp32 outputs and no fused prologue / statistics (use msf_head_bn_stats / msf_head_bn_apply)", i);
    SimtProblem& q = P.p[i];
    q.A = static_cast<const float*>(g.A); q.B = static_cast<const float*>(g.B); q.C = static_cast<float*>(g.C); q.bias = g.bias;
    q.sam = g.a_is_km ? 1 : g.lda; q.sak = g.a_is_km ? g.lda : 1;
    q.sbn = g.b_is_kn ? 1 : g.ldb; q.sbk = g.b_is_kn ? g.ldb : 1;
    q.ldc = g.ldc; q.M = g.M; q.N = g.N; q.K = g.K; q.alpha = g.alpha;
    q.tiles_m = (g.M + 63) / 64;
    q.tile_start = total;
    total += q.tiles_m * ((g.N + 63) / 64);
    flops += 2.0 * g.M * g.N * static_cast<double>(g.K);
  }
  P.n = n_problems;
  P.total = total;
  ProfScope prof(stream, MSF_K_GEMM_F32, flops);
  gemm_simt_grouped_kernel<<<total, 256, 0, static_cast<cudaStream_t>(stream)>>>(P);
  MSF_LAUNCH_OK("gemm_simt_grouped_kernel");
  return MSF_OK;
}
