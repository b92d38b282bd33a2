// N2: the batch norms of the 12 projectors + 12 predictors (src/models/backbone.py:15,18,21,28), grouped: every kernel
// below covers ALL heads and both views of one depth of the head stack in one launch ("multi-tensor" batch norm).
//
// Forward of a depth:   grouped GEMM (gemm_grouped.cu) leaves per-32-row-group column sums of the rounded Linear outputs
//   head_bn_finalize    merges them in fp64, exchanges {sum, sum of squares} with the other ranks over NVLink peer
//                       memory INSIDE the same kernel (SyncBatchNorm, tools/ssl_train.py:160: one exchange per depth
//                       instead of one per layer), and writes scale / shift (consumed by the next GEMM's A prologue),
//                       mean / invstd (saved for the backward) and the running statistics (view 1 then view 2, the
//                       order in which the reference calls the module)
//   head_bn_apply       y = relu?(x * scale + shift) (+ the L2-normalised rows: the InfoNCE keys) -- for the projector
//                       output z, and to rebuild the ReLU activations in the backward (they are never stored)
//   head_bn_stats       column sums straight from an activation matrix (fp64): the fp32 path, where the GEMM is the
//                       SIMT kernel, and a cross-check of the GEMM epilogue
// Backward of a depth:
//   head_bn_bwd_reduce  partial sum(dy'), sum(dy' * xhat) per 256-row block, dy' = g * (bn(y) > 0); also plain column
//                       sums (the bias gradient of the predictor tail)
//   head_bn_bwd_finalize merge (fp64) -> d gamma, d beta (LOCAL sums over both views: DDP averages parameter gradients,
//                       as with SyncBatchNorm), cross-rank exchange -> the two means the input gradient needs
//   head_bn_bwd_elemt   dy = scale * (dy' - mean(dy') - xhat * mean(dy' * xhat))
// All tensors are tiny next to HBM bandwidth (the whole head stack moves < 100 MB per step at batch 256): what matters is
// the launch count (10 forward + 18 backward launches for 60 Linear + 48 BatchNorm layers x 2 views).
// Cross-rank protocol: see head_exchange() -- per-CTA flags in a symmetric workspace, two parities, bounded wait.
#include "common.cuh"

namespace msf {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint64_t ld_acquire_sys_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(uint64_t* p, uint64_t v) {
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

// Symmetric workspace (one per rank, every rank's copy mapped into every process):
//   flags [2 parities][MSF_HEAD_SYNC_MAX_CTAS][MSF_PEER_MAX_WORLD] uint64 | slots [2 parities][world][capacity] doubles
constexpr size_t kFlagBytes = 2ull * MSF_HEAD_SYNC_MAX_CTAS * MSF_PEER_MAX_WORLD * sizeof(uint64_t);

struct Sync {
  char* const* peers;  // device array of `world` workspace pointers (peers[rank] is the local one), or null
  int world, rank;
  uint64_t seq;        // 1, 2, 3, ... identical on all ranks (like any collective)
  size_t capacity;     // doubles per parity and source rank
  uint64_t timeout_ns;
};

// All-reduce (sum, rank order: bit-identical everywhere) of 4 doubles per thread; slot index = blockIdx.x * kThreads +
// threadIdx.x.  PUSH protocol: every thread stores its 4 doubles into slot [parity][my rank] of EVERY rank's workspace
// (NVLink stores), the CTA then raises its flag everywhere, waits for the peers' flags and sums the slots from LOCAL
// memory -- one NVLink one-way trip per exchange.  CTA b of every rank pairs with CTA b of the other ranks only (its own
// flags), so no grid-wide step is needed.  Two parities suffice: a rank enters exchange k+2 only after every peer
// published k+1, which a peer does after its kernel of exchange k has finished reading.  The wait is bounded: a lost peer
// traps (CUDA error), not a hang.
__device__ __forceinline__ void head_exchange(const Sync& sy, double v[4]) {
  if (sy.world <= 1) return;
  const size_t par = sy.seq & 1;
  const size_t par_off = kFlagBytes + par * sy.world * sy.capacity * sizeof(double);
  const size_t idx = (static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x) * 4;
  for (int r = 0; r < sy.world; ++r) {
    double2* dst = reinterpret_cast<double2*>(reinterpret_cast<double*>(sy.peers[r] + par_off) + static_cast<size_t>(sy.rank) * sy.capacity + idx);
    dst[0] = make_double2(v[0], v[1]);
    dst[1] = make_double2(v[2], v[3]);
  }
  __threadfence_system();
  __syncthreads();
  const size_t flag_base = (par * MSF_HEAD_SYNC_MAX_CTAS + blockIdx.x) * MSF_PEER_MAX_WORLD;
  if (threadIdx.x < sy.world)  // publish: my flag for this CTA in every rank's workspace (NVLink stores)
    st_release_sys_u64(reinterpret_cast<uint64_t*>(sy.peers[threadIdx.x]) + flag_base + sy.rank, sy.seq);
  if (threadIdx.x < sy.world) {
    const uint64_t* flag = reinterpret_cast<const uint64_t*>(sy.peers[sy.rank]) + flag_base + threadIdx.x;
    const uint64_t t0 = globaltimer_ns();
    while (ld_acquire_sys_u64(flag) != sy.seq) {
      if (globaltimer_ns() - t0 > sy.timeout_ns) {
        printf("msfwsi_b200: head batch-norm exchange timed out (rank %d, CTA %d waits for rank %d, seq %llu)\n", sy.rank,
               static_cast<int>(blockIdx.x), static_cast<int>(threadIdx.x), static_cast<unsigned long long>(sy.seq));
        __trap();
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
  const double* slots = reinterpret_cast<const double*>(sy.peers[sy.rank] + par_off) + idx;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int r = 0; r < sy.world; ++r) {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] += ld_volatile_f64(slots + static_cast<size_t>(r) * sy.capacity + i);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = acc[i];
}

// ---- block -> (item, block inside the item) through a prefix over the items (kernel-parameter table) ----
template <typename Table>
__device__ __forceinline__ int find_item(const Table& T, int n, int blk) {
  int i = 0;
#pragma unroll 1
  while (i + 1 < n && blk >= T.prefix[i + 1]) ++i;
  return i;
}


// V consecutive floats of a per-column vector as 128-bit loads (16-byte aligned: checked by the entry points)
template <int V>
__device__ __forceinline__ void load_cols(const float* p, float* out) {
#pragma unroll
  for (int e = 0; e < V; e += 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p + e));
    out[e] = v.x; out[e + 1] = v.y; out[e + 2] = v.z; out[e + 3] = v.w;
  }
}

struct FinTable {
  msf_head_bn_item it[MSF_HEAD_MAX_ITEMS];
  int prefix[MSF_HEAD_MAX_ITEMS + 1];  // CTAs: ceil(C / 256) per item
  int n;
};

// ---- forward finalize -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) head_bn_finalize_kernel(const __grid_constant__ FinTable T, const Sync sy, float eps, float momentum,
                                                                    int training) {
  const int i = find_item(T, T.n, blockIdx.x);
  const msf_head_bn_item& q = T.it[i];
  const int c = (blockIdx.x - T.prefix[i]) * kThreads + threadIdx.x;
  const bool ok = c < q.C;
  double v[4] = {0.0, 0.0, 0.0, 0.0};  // {sum, sum of squares} of view 0, view 1
  if (training && ok) {
    const int grows = q.group_rows > 0 ? q.group_rows : 32;
    const int groups = (q.rows + grows - 1) / grows;
    for (int view = 0; view < q.n_views; ++view) {
      const float* cs = q.col_stats[view];
      double s1 = 0.0, s2 = 0.0;
      // fixed order; eight groups per trip with all 16 loads issued before the first add (one thread owns a column and walks
      // rows / 32 groups per view: a chain of dependent round trips otherwise).  Interleaving both views in one trip (32 loads)
      // measured SLOWER (63 vs 35 us per launch at 4096 rows), so the views stay sequential.
#pragma unroll 1
      for (int g = 0; g < groups; g += 8) {
        float a[8], b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool on = g + u < groups;
          a[u] = on ? __ldg(cs + (static_cast<size_t>(g + u) * 2) * q.C + c) : 0.f;
          b[u] = on ? __ldg(cs + (static_cast<size_t>(g + u) * 2 + 1) * q.C + c) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {  // + 0.0 past the end is exact
          s1 += static_cast<double>(a[u]);
          s2 += static_cast<double>(b[u]);
        }
      }
      if (q.centered) {
        // entry 1 of a group is its M2 about the GROUP mean (msf_head_bn_stats): Chan's merge, then back to the
        // sum-reducible {sum, sum of squares} in fp64 -- no E[x^2] - E[x]^2 cancellation in fp32 anywhere
        const double m = s1 / q.rows;
        double m2 = s2;
        for (int g = 0; g < groups; ++g) {
          const int ng = min(32, q.rows - g * 32);
          const double mg = static_cast<double>(__ldg(cs + (static_cast<size_t>(g) * 2) * q.C + c)) / ng;
          m2 += ng * (mg - m) * (mg - m);
        }
        s2 = m2 + s1 * m;
      }
      v[2 * view] = s1;
      v[2 * view + 1] = s2;
    }
  }
  if (training) head_exchange(sy, v);  // every thread of every CTA takes part (uniform control flow)
  if (!ok) return;
  const double n = static_cast<double>(q.rows) * (sy.world > 1 ? sy.world : 1);
  float rm = q.running_mean ? q.running_mean[c] : 0.f, rv = q.running_var ? q.running_var[c] : 1.f;
  const float gam = q.gamma ? q.gamma[c] : 1.f, bet = q.beta ? q.beta[c] : 0.f;
  for (int view = 0; view < q.n_views; ++view) {
    float mean, var;
    if (training) {
      const double m = v[2 * view] / n;
      double vr = v[2 * view + 1] / n - m * m;  // biased variance (what the batch is normalised with)
      if (vr < 0.0) vr = 0.0;
      mean = static_cast<float>(m);
      var = static_cast<float>(vr);
      // running statistics: momentum update with the UNBIASED variance, view 0 first (the reference's call order)
      rm = (1.f - momentum) * rm + momentum * mean;
      rv = (1.f - momentum) * rv + momentum * static_cast<float>(vr * (n / (n - 1.0)));
    } else {
      mean = rm;
      var = rv;
    }
    const float invstd = rsqrtf(var + eps);
    const float sc = gam * invstd;
    q.scale[view][c] = sc;
    // centered (exact fp32 path): consumers evaluate (x - mean) * scale + beta, torch's own association -- near a ReLU zero
    // crossing x * scale and mean * scale are large and nearly cancel, so the folded shift would cost a digit there
    q.shift[view][c] = q.centered ? bet : fmaf(-mean, sc, bet);
    if (q.mean[view]) q.mean[view][c] = mean;
    if (q.invstd[view]) q.invstd[view][c] = invstd;
  }
  if (training) {
    if (q.running_mean) q.running_mean[c] = rm;
    if (q.running_var) q.running_var[c] = rv;
  }
}

// ---- column statistics straight from an activation matrix (the exact fp32 path): col_stats [ceil(rows/32)][2][C] with,
// per 32-row group, entry 0 = sum and entry 1 = M2 = sum (x - group mean)^2 (CENTERED: set msf_head_bn_item.centered) ----
struct StatsTable {
  msf_head_mat it[MSF_HEAD_MAX_MATS];
  int prefix[MSF_HEAD_MAX_MATS + 1];  // CTAs: groups(32 rows) x ceil(C/256)
  int n;
};
template <int DT>
__global__ void __launch_bounds__(kThreads) head_bn_stats_kernel(const __grid_constant__ StatsTable T) {
  const int i = find_item(T, T.n, blockIdx.x);
  const msf_head_mat& q = T.it[i];
  const int cblocks = (q.C + kThreads - 1) / kThreads;
  const int local = blockIdx.x - T.prefix[i];
  const int g = local / cblocks, c = (local % cblocks) * kThreads + threadIdx.x;
  if (c >= q.C) return;
  auto at = [&](int r) -> float {
    if constexpr (DT == MSF_F32) return static_cast<const float*>(q.x)[static_cast<size_t>(r) * q.C + c];
    else if constexpr (DT == MSF_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(q.x)[static_cast<size_t>(r) * q.C + c]);
    else return __half2float(static_cast<const __half*>(q.x)[static_cast<size_t>(r) * q.C + c]);
  };
  const int r0 = g * 32, r1 = min(q.rows, (g + 1) * 32);
  double s1 = 0.0;
  for (int r = r0; r < r1; ++r) s1 += at(r);
  const float s1f = static_cast<float>(s1);
  const double mg = static_cast<double>(s1f) / (r1 - r0);  // the mean the finalize kernel will reconstruct from the stored sum
  double m2 = 0.0;
  for (int r = r0; r < r1; ++r) {  // second read hits L1
    const double d = at(r) - mg;
    m2 += d * d;
  }
  float* dst = q.col_stats + (static_cast<size_t>(g) * 2) * q.C + c;
  dst[0] = s1f;
  dst[q.C] = static_cast<float>(m2);
}

// ---- apply: y = relu?(round(x * scale + shift)) (+ y_hat = y / max(||y||, eps) per row) ----------------------------
struct ApplyTable {
  msf_head_apply_item it[MSF_HEAD_MAX_MATS];
  int prefix[MSF_HEAD_MAX_MATS + 1];  // CTAs: ceil(rows / rows_per_cta)
  int n;
};

template <int DT>
__device__ __forceinline__ float round_to(float v) {
  if constexpr (DT == MSF_F32) return v;
  else if constexpr (DT == MSF_BF16) return __bfloat162float(__float2bfloat16_rn(v));
  else return __half2float(__float2half_rn(v));
}

template <int DT>
__global__ void __launch_bounds__(kThreads) head_bn_apply_kernel(const __grid_constant__ ApplyTable T, float norm_eps) {
  constexpr int V = Elem<DT>::VEC;
  const int i = find_item(T, T.n, blockIdx.x);
  const msf_head_apply_item& q = T.it[i];
  const uint32_t cpr = q.C / V;
  uint32_t lanes = 1;
  while (lanes < 32 && lanes * 4 < cpr) lanes <<= 1;
  const uint32_t lane = threadIdx.x & (lanes - 1), grp = threadIdx.x / lanes, groups = kThreads / lanes;
  const int64_t row = static_cast<int64_t>(blockIdx.x - T.prefix[i]) * groups + grp;
  const bool valid = row < q.rows;
  const char* src = static_cast<const char*>(q.x) + row * static_cast<int64_t>(q.C) * (16 / V);
  char* dst = static_cast<char*>(q.y) + row * static_cast<int64_t>(q.C) * (16 / V);
  float ss = 0.f;
  if (valid)
    for (uint32_t c = lane; c < cpr; c += lanes) {
      float f[V], sc[V], sh[V], mu[V];
      Elem<DT>::unpack(ldg_keep(src + static_cast<size_t>(c) * 16), f);
      load_cols<V>(q.scale + c * V, sc);  // per-column vectors as 128-bit loads (C is a multiple of the chunk width)
      load_cols<V>(q.shift + c * V, sh);
      if (q.mean) load_cols<V>(q.mean + c * V, mu);
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const float xc = q.mean ? f[e] - mu[e] : f[e];
        float t = round_to<DT>(fmaf(xc, sc[e], sh[e]));  // the 16-bit BN output ...
        if (q.relu) t = fmaxf(t, 0.f);                   // ... then the ReLU on it
        f[e] = t;
        ss = fmaf(t, t, ss);
      }
#pragma unroll
      for (int e = 0; e < V; e += (DT == MSF_F32 ? 4 : 8)) stg_stream(dst + static_cast<size_t>(c) * 16, Elem<DT>::pack(f));
    }
  if (q.y_hat == nullptr) return;  // item-uniform
  ss = group_sum(ss, lanes);
  if (!valid) return;
  const float inv = 1.f / fmaxf(sqrtf(ss), norm_eps);
  if (lane == 0 && q.inv_norm) q.inv_norm[row] = inv;
  char* hat = static_cast<char*>(q.y_hat) + row * static_cast<int64_t>(q.C) * (16 / V);
  for (uint32_t c = lane; c < cpr; c += lanes) {  // second read hits L1
    float f[V], sc[V], sh[V], mu[V];
    Elem<DT>::unpack(ldg_keep(src + static_cast<size_t>(c) * 16), f);
    load_cols<V>(q.scale + c * V, sc);
    load_cols<V>(q.shift + c * V, sh);
    if (q.mean) load_cols<V>(q.mean + c * V, mu);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float xc = q.mean ? f[e] - mu[e] : f[e];
      float t = round_to<DT>(fmaf(xc, sc[e], sh[e]));
      if (q.relu) t = fmaxf(t, 0.f);
      f[e] = t * inv;
    }
    stg_stream(hat + static_cast<size_t>(c) * 16, Elem<DT>::pack(f));
  }
}

// ---- backward reduce: per (matrix, 256-row block, 256-column block): partial[rb][2][C] fp32 -------------------------
struct BwdTable {
  msf_head_bwd_item it[MSF_HEAD_MAX_MATS];
  int prefix[MSF_HEAD_MAX_MATS + 1];  // reduce: row blocks x column blocks; elemt: ceil(rows*C/V / (256*4))
  int n;
};
constexpr int kRedRows = 256, kRedCols = 256;  // a CTA: 32 chunk-columns (8 columns each at 16 bit) x 8 row lanes

template <int DT>
__global__ void __launch_bounds__(kThreads) head_bn_bwd_reduce_kernel(const __grid_constant__ BwdTable T) {
  constexpr int V = Elem<DT>::VEC;
  constexpr int kChunkCols = kRedCols / V;           // chunk-columns per CTA (32 at 16 bit, 64 at fp32)
  constexpr int kRowLanes = kThreads / kChunkCols;   // 8 / 4
  __shared__ float sm[kRowLanes][2][kRedCols];
  const int i = find_item(T, T.n, blockIdx.x);
  const msf_head_bwd_item& q = T.it[i];
  const int cblocks = (q.C + kRedCols - 1) / kRedCols;
  const int local = blockIdx.x - T.prefix[i];
  const int rb = local / cblocks, cb = local % cblocks;
  const int cc = threadIdx.x % kChunkCols, rl = threadIdx.x / kChunkCols;
  const int col0 = cb * kRedCols + cc * V;
  float s1[V], s2[V];
#pragma unroll
  for (int e = 0; e < V; ++e) s1[e] = s2[e] = 0.f;
  if (col0 < q.C) {
    float sc[V], sh[V], mu[V], is[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      sc[e] = q.scale ? __ldg(q.scale + col0 + e) : 1.f;
      sh[e] = q.shift ? __ldg(q.shift + col0 + e) : 0.f;
      mu[e] = q.mean ? __ldg(q.mean + col0 + e) : 0.f;
      is[e] = q.invstd ? __ldg(q.invstd + col0 + e) : 0.f;
    }
    const int r1 = min(q.rows, (rb + 1) * kRedRows);
    const bool has_y = q.y != nullptr;
    // four rows per trip, all loads issued before the first use: a thread walks up to 32 rows, and with one row per trip the
    // walk was a chain of 32 DRAM round trips (the launch took ~55 us for ~25 MB).  Same rows in the same order per thread.
#pragma unroll 1
    for (int r = rb * kRedRows + rl; r < r1; r += 4 * kRowLanes) {
      uint4 graw[4], yraw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = r + u * kRowLanes;
        const size_t off = (static_cast<size_t>(rr) * q.C + col0) * (16 / V);
        graw[u] = rr < r1 ? ldg_stream(static_cast<const char*>(q.g) + off) : make_uint4(0, 0, 0, 0);
        yraw[u] = (has_y && rr < r1) ? ldg_stream(static_cast<const char*>(q.y) + off) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r + u * kRowLanes >= r1) break;
        float g[V], y[V];
        Elem<DT>::unpack(graw[u], g);
        Elem<DT>::unpack(yraw[u], y);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          float d = g[e];
          if (has_y) {
            if (q.relu && !(round_to<DT>(fmaf(q.centered ? y[e] - mu[e] : y[e], sc[e], sh[e])) > 0.f)) d = 0.f;  // the mask of the forward's ReLU
            s2[e] = fmaf(d, (y[e] - mu[e]) * is[e], s2[e]);
          }
          s1[e] += d;
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < V; ++e) {
    sm[rl][0][cc * V + e] = s1[e];
    sm[rl][1][cc * V + e] = s2[e];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * kRedCols; t += kThreads) {
    const int k = t / kRedCols, c = t % kRedCols;
    if (cb * kRedCols + c >= q.C) continue;
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < kRowLanes; ++j) a += sm[j][k][c];  // fixed order
    q.partial[(static_cast<size_t>(rb) * 2 + k) * q.C + cb * kRedCols + c] = a;
  }
}

// ---- backward finalize: (head, layer) items, both views ------------------------------------------------------------
struct BwdFinTable {
  msf_head_bwd_fin_item it[MSF_HEAD_MAX_ITEMS];
  int prefix[MSF_HEAD_MAX_ITEMS + 1];
  int n;
};
__global__ void __launch_bounds__(kThreads) head_bn_bwd_finalize_kernel(const __grid_constant__ BwdFinTable T, const Sync sy, int training) {
  const int i = find_item(T, T.n, blockIdx.x);
  const msf_head_bwd_fin_item& q = T.it[i];
  const int c = (blockIdx.x - T.prefix[i]) * kThreads + threadIdx.x;
  const bool ok = c < q.C;
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  if (ok) {
    const int rbs = (q.rows + kRedRows - 1) / kRedRows;
    for (int view = 0; view < q.n_views; ++view) {
      const float* pp = q.partial[view];
#pragma unroll 1
      for (int rb = 0; rb < rbs; rb += 8) {  // eight row blocks per trip, loads first (same order of the adds)
        float a[8], b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const bool on = rb + u < rbs;
          a[u] = on ? __ldg(pp + (static_cast<size_t>(rb + u) * 2) * q.C + c) : 0.f;
          b[u] = on ? __ldg(pp + (static_cast<size_t>(rb + u) * 2 + 1) * q.C + c) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          v[2 * view] += static_cast<double>(a[u]);
          v[2 * view + 1] += static_cast<double>(b[u]);
        }
      }
    }
    // parameter gradients: LOCAL sums over both applications of the module (DDP averages them over ranks)
    if (q.d_beta) q.d_beta[c] = static_cast<float>(v[0] + v[2]);
    if (q.d_gamma) q.d_gamma[c] = static_cast<float>(v[1] + v[3]);
  }
  if (q.plain) {  // item-uniform, but the exchange below must be entered by every thread of the grid: no early return
    v[0] = v[1] = v[2] = v[3] = 0.0;
  }
  if (training) head_exchange(sy, v);
  if (!ok || q.plain) return;
  const double n = static_cast<double>(q.rows) * (sy.world > 1 ? sy.world : 1);
  for (int view = 0; view < q.n_views; ++view) {
    // eval mode: batch norm is an affine map with constant statistics -> no mean terms
    q.c1[view][c] = training ? static_cast<float>(v[2 * view] / n) : 0.f;
    q.c2[view][c] = training ? static_cast<float>(v[2 * view + 1] / n) : 0.f;
  }
}

// ---- backward element-wise: dy = scale * (dy' - c1 - xhat * c2) ----------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(kThreads) head_bn_bwd_elemt_kernel(const __grid_constant__ BwdTable T) {
  constexpr int V = Elem<DT>::VEC;
  const int i = find_item(T, T.n, blockIdx.x);
  const msf_head_bwd_item& q = T.it[i];
  const uint32_t cpr = q.C / V;
  const int64_t chunks = static_cast<int64_t>(q.rows) * cpr;
  const int64_t base = static_cast<int64_t>(blockIdx.x - T.prefix[i]) * (kThreads * 4);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t ch = base + u * kThreads + threadIdx.x;
    if (ch >= chunks) continue;
    const uint32_t c = static_cast<uint32_t>(ch % cpr);
    float g[V], y[V], o[V];
    Elem<DT>::unpack(ldg_stream(static_cast<const char*>(q.g) + ch * 16), g);
    Elem<DT>::unpack(ldg_stream(static_cast<const char*>(q.y) + ch * 16), y);
    // the six per-column vectors as 128-bit loads (C is a multiple of the chunk width, the arrays are 16-byte aligned)
    float sc[V], sh[V], mu[V], is[V], c1[V], c2[V];
    const int col0 = c * V;
#pragma unroll
    for (int e = 0; e < V; e += 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(q.scale + col0 + e)), b = __ldg(reinterpret_cast<const float4*>(q.shift + col0 + e));
      const float4 m = __ldg(reinterpret_cast<const float4*>(q.mean + col0 + e)), i4 = __ldg(reinterpret_cast<const float4*>(q.invstd + col0 + e));
      const float4 k1 = __ldg(reinterpret_cast<const float4*>(q.c1 + col0 + e)), k2 = __ldg(reinterpret_cast<const float4*>(q.c2 + col0 + e));
      sc[e] = a.x; sc[e + 1] = a.y; sc[e + 2] = a.z; sc[e + 3] = a.w;
      sh[e] = b.x; sh[e + 1] = b.y; sh[e + 2] = b.z; sh[e + 3] = b.w;
      mu[e] = m.x; mu[e + 1] = m.y; mu[e + 2] = m.z; mu[e + 3] = m.w;
      is[e] = i4.x; is[e + 1] = i4.y; is[e + 2] = i4.z; is[e + 3] = i4.w;
      c1[e] = k1.x; c1[e + 1] = k1.y; c1[e + 2] = k1.z; c1[e + 3] = k1.w;
      c2[e] = k2.x; c2[e + 1] = k2.y; c2[e + 2] = k2.z; c2[e + 3] = k2.w;
    }
#pragma unroll
    for (int e = 0; e < V; ++e) {
      float d = g[e];
      if (q.relu && !(round_to<DT>(fmaf(q.centered ? y[e] - mu[e] : y[e], sc[e], sh[e])) > 0.f)) d = 0.f;
      o[e] = sc[e] * (d - c1[e] - (y[e] - mu[e]) * is[e] * c2[e]);
    }
    stg_stream(static_cast<char*>(q.dy) + ch * 16, Elem<DT>::pack(o));
  }
}

int fill_sync(Sync& sy, void* const* peers, int world, int rank, uint64_t seq, int64_t capacity, int timeout_ms, int grid) {
  sy = Sync{};
  sy.world = world > 1 ? world : 1;
  if (world <= 1) return MSF_OK;
  MSF_REQUIRE(peers && world <= MSF_PEER_MAX_WORLD && rank >= 0 && rank < world, MSF_ERR_INVALID, "bad peer arguments (world %d rank %d)", world, rank);
  MSF_REQUIRE(seq > 0 && timeout_ms > 0, MSF_ERR_INVALID, "sequence numbers start at 1; timeout must be positive");
  MSF_REQUIRE(grid <= MSF_HEAD_SYNC_MAX_CTAS, MSF_ERR_UNSUPPORTED, "%d CTAs exceed the %d flag slots of the exchange", grid, MSF_HEAD_SYNC_MAX_CTAS);
  MSF_REQUIRE(static_cast<int64_t>(grid) * kThreads * 4 <= capacity, MSF_ERR_WORKSPACE, "exchange of %lld doubles exceeds the capacity %lld",
              static_cast<long long>(grid) * kThreads * 4, static_cast<long long>(capacity));
  sy.peers = reinterpret_cast<char* const*>(peers);
  sy.rank = rank;
  sy.seq = seq;
  sy.capacity = static_cast<size_t>(capacity);
  sy.timeout_ns = static_cast<uint64_t>(timeout_ms) * 1000000ull;
  return MSF_OK;
}

}  // namespace
}  // namespace msf

using namespace msf;

extern "C" size_t msf_head_sync_workspace_bytes(int64_t capacity_doubles) {
  if (capacity_doubles <= 0) return 0;
  return kFlagBytes + 2 * static_cast<size_t>(MSF_PEER_MAX_WORLD) * static_cast<size_t>(capacity_doubles) * sizeof(double);
}

extern "C" int msf_head_bn_finalize(const msf_head_bn_item* items, int n_items, float eps, float momentum, int training, void* const* peers,
                                    int world, int rank, uint64_t seq, int64_t capacity_doubles, int timeout_ms, void* stream) {
  MSF_REQUIRE(items && n_items > 0 && n_items <= MSF_HEAD_MAX_ITEMS, MSF_ERR_INVALID, "n_items %d outside [1, %d]", n_items, MSF_HEAD_MAX_ITEMS);
  static thread_local FinTable T;
  int blocks = 0;
  for (int i = 0; i < n_items; ++i) {
    const msf_head_bn_item& q = items[i];
    MSF_REQUIRE(q.C > 0 && q.rows > 0 && q.n_views >= 1 && q.n_views <= 2, MSF_ERR_INVALID, "item %d: bad shape", i);
    for (int v = 0; v < q.n_views; ++v)
      MSF_REQUIRE((q.col_stats[v] || !training) && q.scale[v] && q.shift[v], MSF_ERR_INVALID, "item %d view %d: NULL pointer", i, v);
    MSF_REQUIRE(!training || static_cast<int64_t>(q.rows) * (world > 1 ? world : 1) > 1, MSF_ERR_INVALID,
                "item %d: Expected more than 1 value per channel when training", i);
    MSF_REQUIRE(training || (q.running_mean && q.running_var), MSF_ERR_INVALID, "item %d: eval mode needs running statistics", i);
    MSF_REQUIRE(q.group_rows >= 0 && (!q.centered || q.group_rows == 0 || q.group_rows == 32), MSF_ERR_INVALID,
                "item %d: group_rows %d (the centered form merges 32-row groups)", i, q.group_rows);
    T.it[i] = q;
    T.prefix[i] = blocks;
    blocks += (q.C + kThreads - 1) / kThreads;
  }
  T.prefix[n_items] = blocks;
  T.n = n_items;
  Sync sy;
  if (int rc = fill_sync(sy, training ? peers : nullptr, training ? world : 1, rank, seq, capacity_doubles, timeout_ms, blocks)) return rc;
  ProfScope prof(stream, MSF_K_HEAD_BN_FINALIZE, 0.0);
  head_bn_finalize_kernel<<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(T, sy, eps, momentum, training);
  MSF_LAUNCH_OK("head_bn_finalize_kernel");
  return MSF_OK;
}

extern "C" int msf_head_bn_stats(const msf_head_mat* mats, int n, int dtype, void* stream) {
  MSF_REQUIRE(mats && n > 0 && n <= MSF_HEAD_MAX_MATS && dtype_ok(dtype), MSF_ERR_INVALID, "bad arguments");
  static thread_local StatsTable T;
  int blocks = 0;
  double bytes = 0.0;
  for (int i = 0; i < n; ++i) {
    MSF_REQUIRE(mats[i].x && mats[i].col_stats && mats[i].rows > 0 && mats[i].C > 0, MSF_ERR_INVALID, "matrix %d: bad arguments", i);
    T.it[i] = mats[i];
    T.prefix[i] = blocks;
    blocks += ((mats[i].rows + 31) / 32) * ((mats[i].C + kThreads - 1) / kThreads);
    bytes += static_cast<double>(mats[i].rows) * mats[i].C * dtype_size(dtype);
  }
  T.prefix[n] = blocks;
  T.n = n;
  ProfScope prof(stream, MSF_K_HEAD_BN_ELEMWISE, bytes);
  MSF_DISPATCH_DTYPE(dtype, (head_bn_stats_kernel<DT><<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(T)));
  MSF_LAUNCH_OK("head_bn_stats_kernel");
  return MSF_OK;
}

extern "C" int msf_head_bn_apply(const msf_head_apply_item* items, int n, int dtype, float norm_eps, void* stream) {
  MSF_REQUIRE(items && n > 0 && n <= MSF_HEAD_MAX_MATS && dtype_ok(dtype), MSF_ERR_INVALID, "bad arguments");
  static thread_local ApplyTable T;
  const int vec = 16 / static_cast<int>(dtype_size(dtype));
  int blocks = 0;
  double bytes = 0.0;
  for (int i = 0; i < n; ++i) {
    const msf_head_apply_item& q = items[i];
    MSF_REQUIRE(q.x && q.y && q.scale && q.shift && q.rows > 0 && q.C > 0 && q.C % vec == 0, MSF_ERR_INVALID, "item %d: bad arguments (C %% %d)", i, vec);
    MSF_REQUIRE(aligned16(q.x) && aligned16(q.y) && aligned16(q.y_hat) && aligned16(q.scale) && aligned16(q.shift) && aligned16(q.mean), MSF_ERR_INVALID,
                "item %d: pointers must be 16-byte aligned", i);
    T.it[i] = q;
    T.prefix[i] = blocks;
    const uint32_t cpr = q.C / vec;
    uint32_t lanes = 1;
    while (lanes < 32 && lanes * 4 < cpr) lanes <<= 1;
    const int groups = kThreads / static_cast<int>(lanes);
    blocks += (q.rows + groups - 1) / groups;
    bytes += static_cast<double>(q.rows) * q.C * dtype_size(dtype) * (q.y_hat ? 3.0 : 2.0);
  }
  T.prefix[n] = blocks;
  T.n = n;
  ProfScope prof(stream, MSF_K_HEAD_BN_ELEMWISE, bytes);
  MSF_DISPATCH_DTYPE(dtype, (head_bn_apply_kernel<DT><<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(T, norm_eps)));
  MSF_LAUNCH_OK("head_bn_apply_kernel");
  return MSF_OK;
}

extern "C" int msf_head_bn_bwd_reduce(const msf_head_bwd_item* items, int n, int dtype, void* stream) {
  MSF_REQUIRE(items && n > 0 && n <= MSF_HEAD_MAX_MATS && dtype_ok(dtype), MSF_ERR_INVALID, "bad arguments");
  static thread_local BwdTable T;
  const int vec = 16 / static_cast<int>(dtype_size(dtype));
  int blocks = 0;
  double bytes = 0.0;
  for (int i = 0; i < n; ++i) {
    const msf_head_bwd_item& q = items[i];
    MSF_REQUIRE(q.g && q.partial && q.rows > 0 && q.C > 0 && q.C % vec == 0, MSF_ERR_INVALID, "item %d: bad arguments", i);
    MSF_REQUIRE(!q.y || (q.mean && q.invstd && (!q.relu || (q.scale && q.shift))), MSF_ERR_INVALID, "item %d: statistics missing", i);
    T.it[i] = q;
    T.prefix[i] = blocks;
    blocks += ((q.rows + kRedRows - 1) / kRedRows) * ((q.C + kRedCols - 1) / kRedCols);
    bytes += static_cast<double>(q.rows) * q.C * dtype_size(dtype) * (q.y ? 2.0 : 1.0);
  }
  T.prefix[n] = blocks;
  T.n = n;
  ProfScope prof(stream, MSF_K_HEAD_BN_ELEMWISE, bytes);
  MSF_DISPATCH_DTYPE(dtype, (head_bn_bwd_reduce_kernel<DT><<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(T)));
  MSF_LAUNCH_OK("head_bn_bwd_reduce_kernel");
  return MSF_OK;
}

extern "C" int msf_head_bn_bwd_finalize(const msf_head_bwd_fin_item* items, int n_items, int training, void* const* peers, int world, int rank,
                                        uint64_t seq, int64_t capacity_doubles, int timeout_ms, void* stream) {
  MSF_REQUIRE(items && n_items > 0 && n_items <= MSF_HEAD_MAX_ITEMS, MSF_ERR_INVALID, "n_items %d outside [1, %d]", n_items, MSF_HEAD_MAX_ITEMS);
  static thread_local BwdFinTable T;
  int blocks = 0;
  for (int i = 0; i < n_items; ++i) {
    const msf_head_bwd_fin_item& q = items[i];
    MSF_REQUIRE(q.C > 0 && q.rows > 0 && q.n_views >= 1 && q.n_views <= 2, MSF_ERR_INVALID, "item %d: bad shape", i);
    for (int v = 0; v < q.n_views; ++v)
      MSF_REQUIRE(q.partial[v] && (q.plain || (q.c1[v] && q.c2[v])), MSF_ERR_INVALID, "item %d view %d: NULL pointer", i, v);
    T.it[i] = q;
    T.prefix[i] = blocks;
    blocks += (q.C + kThreads - 1) / kThreads;
  }
  T.prefix[n_items] = blocks;
  T.n = n_items;
  Sync sy;
  if (int rc = fill_sync(sy, training ? peers : nullptr, training ? world : 1, rank, seq, capacity_doubles, timeout_ms, blocks)) return rc;
  ProfScope prof(stream, MSF_K_HEAD_BN_FINALIZE, 0.0);
  head_bn_bwd_finalize_kernel<<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(T, sy, training);
  MSF_LAUNCH_OK("head_bn_bwd_finalize_kernel");
  return MSF_OK;
}

extern "C" int msf_head_bn_bwd_elemt(const msf_head_bwd_item* items, int n, int dtype, void* stream) {
  MSF_REQUIRE(items && n > 0 && n <= MSF_HEAD_MAX_MATS && dtype_ok(dtype), MSF_ERR_INVALID, "bad arguments");
  static thread_local BwdTable T;
  const int vec = 16 / static_cast<int>(dtype_size(dtype));
  int blocks = 0;
  double bytes = 0.0;
  for (int i = 0; i < n; ++i) {
    const msf_head_bwd_item& q = items[i];
    MSF_REQUIRE(q.g && q.y && q.dy && q.scale && q.shift && q.mean && q.invstd && q.c1 && q.c2 && q.rows > 0 && q.C > 0 && q.C % vec == 0,
                MSF_ERR_INVALID, "item %d: bad arguments", i);
    MSF_REQUIRE(aligned16(q.scale) && aligned16(q.shift) && aligned16(q.mean) && aligned16(q.invstd) && aligned16(q.c1) && aligned16(q.c2), MSF_ERR_INVALID,
                "item %d: the per-column vectors must be 16-byte aligned", i);
    T.it[i] = q;
    T.prefix[i] = blocks;
    const int64_t chunks = static_cast<int64_t>(q.rows) * (q.C / vec);
    blocks += static_cast<int>((chunks + kThreads * 4 - 1) / (kThreads * 4));
    bytes += static_cast<double>(q.rows) * q.C * dtype_size(dtype) * 3.0;
  }
  T.prefix[n] = blocks;
  T.n = n;
  ProfScope prof(stream, MSF_K_HEAD_BN_ELEMWISE, bytes);
  MSF_DISPATCH_DTYPE(dtype, (head_bn_bwd_elemt_kernel<DT><<<blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(T)));
  MSF_LAUNCH_OK("head_bn_bwd_elemt_kernel");
  return MSF_OK;
}
