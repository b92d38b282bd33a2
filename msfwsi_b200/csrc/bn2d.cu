// Channels-last (NHWC) train-mode BatchNorm2d for the encoders that feed the hot path, with the element-wise work
// around it fused in: ReLU, the residual add of a BasicBlock, and the stem's 3x3/2 max-pool.  (Caller-side helper:
// the convolutions stay on cuDNN.  ATen's channels-last batch-norm / max-pool / add / relu kernels reach 0.5-1 TB/s on
// the 64..512-channel ResNet-18 activations and together cost 2.7x the convolutions at batch 4096; every kernel here
// is a streaming pass bounded by HBM bandwidth.)
//
// The activation is viewed as a [rows = N*H*W][C] matrix with C contiguous.
//   bn_stats          partial per-channel sum / sum of squares per CTA (fp32 in registers, fixed-order shared reduce)
//   bn_combine        partials -> fp64 sums[2C] (+ the element count at sums[2C], so a cross-rank all-reduce of the
//                     whole vector yields global statistics even with unequal per-rank batches)
//   bn_finalize       sums -> mean, invstd, running-stat update
//   bn_apply          y = act(x * sc + sh (+ res)),  sc = gamma*invstd, sh = beta - mean*sc
//   bn_bwd_reduce     partial sum(dy'), sum(dy' * xhat), dy' = dy * (y > 0); y is recomputed from x (no mask stored)
//                     or, with a residual, read from the saved output
//   bn_bwd_elemt      dx = sc * (dy' - mean(dy') - xhat * mean(dy' * xhat));  dres = dy'
//   bn_apply_pool     stem: y = maxpool3x3/2/pad1(relu(bn(x))) + one byte per output with the arg-max tap (255 = dead)
//   bn_pool_bwd_*     the two backward passes through pool + relu + bn from the pooled gradient and the tap bytes
// Bytes per element: forward 2 reads + 1 write, backward 4 reads + 1 write (+1 read / +1 write with a residual).
// 64-bit indexing throughout (the stem activation of a 4096-image batch has 3.3e9 elements).
#include <math_constants.h>

#include "common.cuh"

namespace msf {
namespace {

constexpr int kThreads = 256;

// thread layout inside a CTA of the reduce kernels: `ct` consecutive threads cover `ct` 16-byte chunks of a row
// (a "column group"), kThreads/ct row lanes walk the rows.  CTA b handles the row groups b, b+G, b+2G, ... (G = grid
// size = SMs x resident CTAs, so there is exactly one wave and all CTAs sweep memory together).
constexpr int kMaxRowBlocks = kNumSMs * 8;
constexpr int kStatsMaxBlocks = 640;  // row blocks of the forward statistics kernel (one wave at 4 CTAs/SM is 592)
struct Layout {
  int cvec;     // 16-byte chunks per row
  int ct;       // chunks per column group handled by one CTA (power of two <= 32)
  int cgroups;  // column groups
  int rlanes;   // row lanes per CTA
};
inline Layout make_layout(int C, int vec) {
  Layout l;
  l.cvec = C / vec;
  l.ct = 1;
  while (l.ct < 32 && l.ct * 2 <= l.cvec && l.cvec % (l.ct * 2) == 0) l.ct *= 2;
  l.cgroups = l.cvec / l.ct;
  l.rlanes = kThreads / l.ct;
  return l;
}
// number of row blocks (grid.x) for a reduce kernel: one wave of resident CTAs, never more than the rows need
template <typename Kern>
inline int reduce_grid(Kern kern, size_t smem, const Layout& l, int64_t rows, int rows_per_iter) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem) != cudaSuccess || occ < 1) occ = 1;
  int64_t g = static_cast<int64_t>(kNumSMs) * occ / l.cgroups;
  const int64_t need = (rows + rows_per_iter - 1) / rows_per_iter;
  if (g > need) g = need;
  if (g > kMaxRowBlocks) g = kMaxRowBlocks;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// fixed-order reduction of the per-thread accumulators over the row lanes of a CTA -> partial[blk][2][C]
template <int V>
__device__ __forceinline__ void cta_reduce_store(const float* s, const float* q, int ct, int cvec, int cl, int rl,
                                                 float* __restrict__ partial) {
  extern __shared__ float sm[];
  const int rlanes = kThreads / ct;
  float* mine = sm + (static_cast<size_t>(rl) * ct + cl) * (2 * V);
#pragma unroll
  for (int i = 0; i < V; ++i) { mine[i] = s[i]; mine[V + i] = q[i]; }
  __syncthreads();
  for (int e = threadIdx.x; e < ct * 2 * V; e += kThreads) {
    const int c2 = e / (2 * V), k = e % (2 * V);
    float t = 0.f;
    for (int j = 0; j < rlanes; ++j) t += sm[(static_cast<size_t>(j) * ct + c2) * (2 * V) + k];
    const int C = cvec * V;
    const int ch = (blockIdx.y * ct + c2) * V + (k % V);
    partial[(static_cast<size_t>(blockIdx.x) * 2 + (k / V)) * C + ch] = t;
  }
}

constexpr int kStatsRows = 8;  // rows in flight per thread
// Shifted sums: every CTA accumulates d = x - K with K = its first row (per channel), so sum d^2 - (sum d)^2/n does not
// cancel when |mean| >> std.  partial layout [blk][3][C] = {sum d, sum d^2, K}, followed by cnt[blk] (rows of the block).
template <int DT>
__global__ void __launch_bounds__(kThreads) bn_stats_kernel(const char* __restrict__ x, int64_t rows, int cvec, int ct,
                                                            float* __restrict__ partial, float* __restrict__ cnt) {
  constexpr int V = Elem<DT>::VEC;
  constexpr int U = kStatsRows;
  extern __shared__ float sm[];
  const int cl = threadIdx.x % ct, rl = threadIdx.x / ct, rlanes = kThreads / ct;
  const int chunk = blockIdx.y * ct + cl;
  const int C = cvec * V;
  const int64_t first = static_cast<int64_t>(blockIdx.x) * rlanes;  // first row of this CTA
  float kshift[V], s[V], q[V];
#pragma unroll
  for (int i = 0; i < V; ++i) kshift[i] = s[i] = q[i] = 0.f;
  if (first < rows) Elem<DT>::unpack(ldg_keep(x + (first * cvec + chunk) * 16), kshift);
  // the U rows a thread has in flight are a whole grid sweep apart (gridDim.x * rlanes rows): every sweep is one
  // contiguous window of memory read by all CTAs together
  const int64_t sweep = static_cast<int64_t>(gridDim.x) * rlanes;
  auto accumulate = [&](const uint4& v) {
    float f[V];
    Elem<DT>::unpack(v, f);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float d = f[i] - kshift[i];
      s[i] += d;
      q[i] = fmaf(d, d, q[i]);
    }
  };
  int64_t r = first + rl;
  const int64_t mine = r < rows ? (rows - r + sweep - 1) / sweep : 0;  // rows this thread accumulates
  for (; r + (U - 1) * sweep < rows; r += U * sweep) {  // full groups: U unconditional loads in flight
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ldg_stream(x + ((r + u * sweep) * cvec + chunk) * 16);
#pragma unroll
    for (int u = 0; u < U; ++u) accumulate(v[u]);
  }
  for (; r < rows; r += sweep) accumulate(ldg_stream(x + (r * cvec + chunk) * 16));
  // fixed-order reduction over the row lanes -> partial[blk][0..1][C]; K -> partial[blk][2][C]; row count -> cnt[blk]
  float* slot = sm + (static_cast<size_t>(rl) * ct + cl) * (2 * V);
#pragma unroll
  for (int i = 0; i < V; ++i) { slot[i] = s[i]; slot[V + i] = q[i]; }
  __syncthreads();
  for (int e = threadIdx.x; e < ct * 2 * V; e += kThreads) {
    const int c2 = e / (2 * V), k = e % (2 * V);
    float t = 0.f;
    for (int j = 0; j < rlanes; ++j) t += sm[(static_cast<size_t>(j) * ct + c2) * (2 * V) + k];
    const int ch = (blockIdx.y * ct + c2) * V + (k % V);
    partial[(static_cast<size_t>(blockIdx.x) * 3 + (k / V)) * C + ch] = t;
  }
  if (rl == 0) {
#pragma unroll
    for (int i = 0; i < V; ++i) partial[(static_cast<size_t>(blockIdx.x) * 3 + 2) * C + chunk * V + i] = kshift[i];
  }
  // rows of this block = sum over its row lanes (every column group sees the same rows)
  __shared__ int lane_rows[kThreads];
  if (cl == 0) lane_rows[rl] = static_cast<int>(mine);
  __syncthreads();
  if (blockIdx.y == 0 && threadIdx.x == 0) {
    int t = 0;
    for (int j = 0; j < rlanes; ++j) t += lane_rows[j];
    cnt[blockIdx.x] = static_cast<float>(t);
  }
}

// Moments of one channel over all row blocks of the forward partials, fp64, division-free per block:
//   pass 1   S1 = sum_b (n_b * K_b + sd_b)                      -> mean = S1 / N
//   pass 2   M2 = sum_b (sq_b + 2*delta_b*sd_b + n_b*delta_b^2),  delta_b = K_b - mean   (= sum over rows of (x - mean)^2,
//            since x = K_b + d; every term is a deviation from a nearby value, so nothing cancels catastrophically)
// One warp per channel (lane l takes the blocks l, l+32, ...; fixed butterfly => deterministic); result in every lane.
struct Moments {
  double n, mean, m2, s1;
};
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ Moments channel_moments(const float* __restrict__ partial, const float* __restrict__ cnt, int nblk, int C,
                                                   int ch, int lane) {
  // the partials are few (<= 1184 blocks) but sit in L2: the loads of four blocks are issued together and the values kept
  // in registers for the second pass (the statistics kernel runs at most kStatsMaxBlocks = 640 row blocks: 20 per lane)
  constexpr int kMine = kStatsMaxBlocks / 32;
  float nb[kMine], sd[kMine], sq[kMine], kk[kMine];
#pragma unroll
  for (int j = 0; j < kMine; ++j) {
    const int b = lane + 32 * j;
    const bool in = b < nblk;
    nb[j] = in ? __ldg(cnt + b) : 0.f;
    sd[j] = in ? __ldg(partial + (static_cast<size_t>(b) * 3 + 0) * C + ch) : 0.f;
    sq[j] = in ? __ldg(partial + (static_cast<size_t>(b) * 3 + 1) * C + ch) : 0.f;
    kk[j] = in ? __ldg(partial + (static_cast<size_t>(b) * 3 + 2) * C + ch) : 0.f;
  }
  double n = 0.0, s1 = 0.0;
#pragma unroll
  for (int j = 0; j < kMine; ++j) {
    n += static_cast<double>(nb[j]);
    s1 += static_cast<double>(nb[j]) * static_cast<double>(kk[j]) + static_cast<double>(sd[j]);
  }
  n = warp_sum(n);
  s1 = warp_sum(s1);
  const double mean = s1 / n;
  double m2 = 0.0;
#pragma unroll
  for (int j = 0; j < kMine; ++j) {
    const double delta = static_cast<double>(kk[j]) - mean;
    m2 += static_cast<double>(sq[j]) + delta * (2.0 * static_cast<double>(sd[j]) + static_cast<double>(nb[j]) * delta);  // empty blocks: all zero
  }
  m2 = fmax(warp_sum(m2), 0.0);
  return Moments{n, mean, m2, s1};
}

// forward partials -> the sum-reducible fp64 vector {sum x, sum x^2 per channel, count} (exact to fp64 rounding)
constexpr int kCombineWarps = 8;
__global__ void __launch_bounds__(kCombineWarps * 32) bn_stats_combine_kernel(const float* __restrict__ partial, const float* __restrict__ cnt,
                                                                              int nblk, int C, double* __restrict__ sums, double count) {
  const int lane = threadIdx.x & 31;
  const int ch = blockIdx.x * kCombineWarps + (threadIdx.x >> 5);
  if (blockIdx.x == 0 && threadIdx.x == 0) sums[2 * C] = count;
  if (ch >= C) return;
  const Moments m = channel_moments(partial, cnt, nblk, C, ch, lane);
  if (lane == 0) {
    sums[ch] = m.s1;
    sums[C + ch] = m.m2 + m.n * m.mean * m.mean;
  }
}

// partial [nblk][2][C] -> sums[2][C] (fp64); sums[2C] = count when count >= 0.  One warp per (which, channel) entry:
// lane l sums the row blocks l, l+32, ... in fp64, then a fixed butterfly over the lanes (deterministic).
__global__ void __launch_bounds__(kCombineWarps * 32) bn_combine_kernel(const float* __restrict__ partial, int nblk, int C,
                                                                        double* __restrict__ sums, double count) {
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * kCombineWarps + (threadIdx.x >> 5);
  if (blockIdx.x == 0 && threadIdx.x == 0 && count >= 0.0) sums[2 * C] = count;
  if (e >= 2 * C) return;
  constexpr int kMine = (kMaxRowBlocks + 31) / 32;
  float v[kMine];
#pragma unroll
  for (int j = 0; j < kMine; ++j) v[j] = (lane + 32 * j < nblk) ? __ldg(partial + static_cast<size_t>(lane + 32 * j) * 2 * C + e) : 0.f;  // loads in flight together
  double t = 0.0;
#pragma unroll
  for (int j = 0; j < kMine; ++j) t += static_cast<double>(v[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) sums[e] = t;
}

// single-process forward: combine + finalize in one launch (one warp per channel)
__global__ void __launch_bounds__(kCombineWarps * 32) bn_combine_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ cnt,
                                                                                 int nblk, int C, double* __restrict__ sums, double count,
                                                                                 float eps, float momentum, float* __restrict__ mean,
                                                                                 float* __restrict__ invstd, float* __restrict__ running_mean,
                                                                                 float* __restrict__ running_var) {
  const int lane = threadIdx.x & 31;
  const int ch = blockIdx.x * kCombineWarps + (threadIdx.x >> 5);
  if (blockIdx.x == 0 && threadIdx.x == 0) sums[2 * C] = count;
  if (ch >= C) return;
  const Moments mo = channel_moments(partial, cnt, nblk, C, ch, lane);
  if (lane != 0) return;
  sums[ch] = mo.s1;
  sums[C + ch] = mo.m2 + mo.n * mo.mean * mo.mean;
  const double var = mo.m2 / mo.n;
  mean[ch] = static_cast<float>(mo.mean);
  invstd[ch] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  if (running_mean) {
    const double unbiased = mo.n > 1.0 ? mo.m2 / (mo.n - 1.0) : var;
    running_mean[ch] = static_cast<float>((1.0 - momentum) * running_mean[ch] + momentum * mo.mean);
    running_var[ch] = static_cast<float>((1.0 - momentum) * running_var[ch] + momentum * unbiased);
  }
}

// sums[2C+1] (possibly all-reduced over ranks) -> mean, invstd, running stats (momentum update, unbiased variance)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int C, float eps, float momentum, float* __restrict__ mean,
                                   float* __restrict__ invstd, float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= C) return;
  const double count = sums[2 * C];
  const double m = sums[ch] / count;
  double var = sums[C + ch] / count - m * m;
  if (var < 0.0) var = 0.0;
  mean[ch] = static_cast<float>(m);
  invstd[ch] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * (count / (count - 1.0)) : var;
    running_mean[ch] = static_cast<float>((1.0 - momentum) * running_mean[ch] + momentum * m);
    running_var[ch] = static_cast<float>((1.0 - momentum) * running_var[ch] + momentum * unbiased);
  }
}

// per-channel constants of one 16-byte chunk
template <int V>
struct ChanConst {
  float sc[V], sh[V];
  __device__ __forceinline__ void load(int ch0, const float* __restrict__ mean, const float* __restrict__ invstd,
                                       const float* __restrict__ gamma, const float* __restrict__ beta) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      sc[i] = __ldg(invstd + ch0 + i) * (gamma ? __ldg(gamma + ch0 + i) : 1.f);
      sh[i] = fmaf(-__ldg(mean + ch0 + i), sc[i], beta ? __ldg(beta + ch0 + i) : 0.f);  // explicit: fwd and bwd kernels must agree bit for bit
    }
  }
};

// `bits` (may be NULL): one byte per 16-byte chunk, bit i = (pre-ReLU value of element i > 0) -- the ReLU mask the
// backward needs when a residual was added, 16x smaller than re-reading the bf16 output.
template <int DT, bool RES>
__global__ void __launch_bounds__(kThreads) bn_apply_kernel(const char* __restrict__ x, const char* __restrict__ res,
                                                            char* __restrict__ y, unsigned char* __restrict__ bits, int64_t chunks, int cvec,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta, int relu) {
  constexpr int V = Elem<DT>::VEC;
  constexpr int U = RES ? 2 : 4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  const bool fixed = (kThreads % cvec) == 0;  // then a thread always sees the same channel chunk: constants in registers
  ChanConst<V> k;
  if (fixed) k.load((threadIdx.x % cvec) * V, mean, invstd, gamma, beta);
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; e < chunks; e += U * stride) {
    uint4 v[U], w[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (e + u * stride < chunks) {
        v[u] = ldg_stream(x + (e + u * stride) * 16);
        if (RES) w[u] = ldg_stream(res + (e + u * stride) * 16);
      }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t idx = e + u * stride;
      if (idx >= chunks) break;
      if (!fixed) k.load(static_cast<int>(idx % cvec) * V, mean, invstd, gamma, beta);
      float f[V], g[V];
      Elem<DT>::unpack(v[u], f);
      if (RES) Elem<DT>::unpack(w[u], g);
      uint32_t alive = 0;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float o = fmaf(f[i], k.sc[i], k.sh[i]);
        if (RES) o += g[i];
        alive |= (o > 0.f ? 1u : 0u) << i;
        if (relu) o = fmaxf(o, 0.f);
        f[i] = o;
      }
      stg_stream(y + idx * 16, Elem<DT>::pack(f));
      if (bits) bits[idx] = static_cast<unsigned char>(alive);
    }
  }
}

// MASK: 0 none, 1 relu mask recomputed from x, 2 relu mask from the saved output y, 3 relu mask from one bit per element
constexpr int kReduceRows = 4;  // rows in flight per thread (x 2 or 3 tensors)
// GP: a per-(image, channel) gradient gp[n][c] * gp_scale is added to every dy of that image before the mask (the
// backward of a global average pool over the same output, folded in instead of being materialised and added).
template <int DT, int MASK, bool GP>
__global__ void __launch_bounds__(kThreads) bn_bwd_reduce_kernel(const char* __restrict__ x, const char* __restrict__ dy,
                                                                 const char* __restrict__ ymask, int64_t rows, int cvec, int ct,
                                                                 const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 const char* __restrict__ gp, unsigned hw, float gp_scale,
                                                                 float* __restrict__ partial) {
  constexpr int V = Elem<DT>::VEC;
  constexpr int U = kReduceRows;
  const int cl = threadIdx.x % ct, rl = threadIdx.x / ct, rlanes = kThreads / ct;
  const int chunk = blockIdx.y * ct + cl;
  float mu[V], is[V], s[V], q[V];
  ChanConst<V> k;
  if (MASK == 1) k.load(chunk * V, mean, invstd, gamma, beta);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    mu[i] = mean[chunk * V + i];
    is[i] = invstd[chunk * V + i];
    s[i] = q[i] = 0.f;
  }
  // (two or three streams already spread the requests; adjacent row groups measured faster here than the grid-sweep
  // spacing bn_stats_kernel uses)
  const int64_t step = static_cast<int64_t>(gridDim.x) * U * rlanes;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * U * rlanes + rl; r < rows; r += step) {
    uint4 a[U], b[U], m[U], p[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + u * rlanes;
      a[u] = b[u] = m[u] = p[u] = make_uint4(0, 0, 0, 0);  // dy = 0 (and y = 0 for the mask) contributes nothing
      if (rr < rows) {
        a[u] = ldg_stream(x + (rr * cvec + chunk) * 16);
        b[u] = ldg_stream(dy + (rr * cvec + chunk) * 16);
        if (MASK == 2) m[u] = ldg_stream(ymask + (rr * cvec + chunk) * 16);
        if (MASK == 3) m[u].x = __ldg(reinterpret_cast<const unsigned char*>(ymask) + rr * cvec + chunk);
        if (GP) p[u] = ldg_keep(gp + (static_cast<int64_t>(static_cast<unsigned>(rr) / hw) * cvec + chunk) * 16);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float fx[V], fd[V], fm[V], fp[V];
      Elem<DT>::unpack(a[u], fx);
      Elem<DT>::unpack(b[u], fd);
      if (MASK == 2) Elem<DT>::unpack(m[u], fm);
      if (GP) Elem<DT>::unpack(p[u], fp);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float d = GP ? fmaf(fp[i], gp_scale, fd[i]) : fd[i];
        if (MASK == 1 && fmaf(fx[i], k.sc[i], k.sh[i]) <= 0.f) d = 0.f;
        if (MASK == 2 && fm[i] <= 0.f) d = 0.f;
        if (MASK == 3 && !((m[u].x >> i) & 1u)) d = 0.f;
        s[i] += d;
        q[i] = fmaf(d, (fx[i] - mu[i]) * is[i], q[i]);
      }
    }
  }
  cta_reduce_store<V>(s, q, ct, cvec, cl, rl, partial);
}

// per-channel constants of the input-gradient pass:  dx = sc*(d - m1 - xhat*m2) = sc*d + (ca*x + cb),
// m1 = sum(dy')/n, m2 = sum(dy'*xhat)/n, ca = -sc*m2*invstd, cb = sc*(m2*invstd*mean - m1)
template <int V>
struct GradConst {
  float sc[V], ca[V], cb[V];
  __device__ __forceinline__ void load(int ch0, int C, const float* __restrict__ mean, const float* __restrict__ invstd,
                                       const float* __restrict__ gamma, const double* __restrict__ sums, float inv_n) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int ch = ch0 + i;
      const float is = __ldg(invstd + ch), mu = __ldg(mean + ch);
      sc[i] = is * (gamma ? __ldg(gamma + ch) : 1.f);
      const float m1 = static_cast<float>(sums[ch]) * inv_n, m2 = static_cast<float>(sums[C + ch]) * inv_n;
      ca[i] = -sc[i] * m2 * is;
      cb[i] = sc[i] * (m2 * is * mu - m1);
    }
  }
};

template <int DT, int MASK, bool DRES, bool GP>
__global__ void __launch_bounds__(kThreads) bn_bwd_elemt_kernel(const char* __restrict__ x, const char* __restrict__ dy,
                                                                const char* __restrict__ ymask, char* __restrict__ dx,
                                                                char* __restrict__ dres, int64_t chunks, int cvec,
                                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                const double* __restrict__ sums /*[2][C]*/,
                                                                const double* __restrict__ count, const char* __restrict__ gp,
                                                                unsigned hw_chunks /* H*W*cvec */, float gp_scale) {
  constexpr int V = Elem<DT>::VEC;
  const int C = cvec * V;
  const float inv_n = static_cast<float>(1.0 / *count);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  const bool fixed = (kThreads % cvec) == 0;
  GradConst<V> g;
  float sh[V];  // relu mask recomputation: y = x*sc + sh
  auto load_consts = [&](int ch0) {
    g.load(ch0, C, mean, invstd, gamma, sums, inv_n);
    if (MASK == 1) {
#pragma unroll
      for (int i = 0; i < V; ++i) sh[i] = fmaf(-__ldg(mean + ch0 + i), g.sc[i], beta ? __ldg(beta + ch0 + i) : 0.f);
    }
  };
  if (fixed) load_consts((threadIdx.x % cvec) * V);
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; e < chunks; e += 2 * stride) {
    uint4 vx[2], vd[2], vm[2], vp[2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (e + u * stride < chunks) {
        const int64_t idx = e + u * stride;
        vx[u] = ldg_stream(x + idx * 16);
        vd[u] = ldg_stream(dy + idx * 16);
        if (MASK == 2) vm[u] = ldg_stream(ymask + idx * 16);
        if (MASK == 3) vm[u].x = __ldg(reinterpret_cast<const unsigned char*>(ymask) + idx);
        if (GP) {  // chunk (n, c) of the pooled gradient; 32-bit index math (the ABI requires rows*cvec < 2^32 here)
          const unsigned i32 = static_cast<unsigned>(idx);
          vp[u] = ldg_keep(gp + (static_cast<int64_t>(i32 / hw_chunks) * cvec + i32 % static_cast<unsigned>(cvec)) * 16);
        }
      }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t idx = e + u * stride;
      if (idx >= chunks) break;
      if (!fixed) load_consts(static_cast<int>(idx % cvec) * V);
      float fx[V], fd[V], fm[V], fp[V];
      Elem<DT>::unpack(vx[u], fx);
      Elem<DT>::unpack(vd[u], fd);
      if (MASK == 2) Elem<DT>::unpack(vm[u], fm);
      if (GP) Elem<DT>::unpack(vp[u], fp);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float d = GP ? fmaf(fp[i], gp_scale, fd[i]) : fd[i];
        if (MASK == 1 && fmaf(fx[i], g.sc[i], sh[i]) <= 0.f) d = 0.f;
        if (MASK == 2 && fm[i] <= 0.f) d = 0.f;
        if (MASK == 3 && !((vm[u].x >> i) & 1u)) d = 0.f;
        fd[i] = d;
        fx[i] = fmaf(g.sc[i], d, fmaf(g.ca[i], fx[i], g.cb[i]));
      }
      stg_stream(dx + idx * 16, Elem<DT>::pack(fx));
      if (DRES) stg_stream(dres + idx * 16, Elem<DT>::pack(fd));
    }
  }
}

// ---- stem: bn + relu + 3x3 stride-2 pad-1 max-pool ------------------------------------------------------------
// Window (ph, pw) covers input rows 2ph-1..2ph+1 and columns 2pw-1..2pw+1; tap t = dr*3 + dc.  The arg-max is the
// first maximum in (row, column) scan order (ATen's max_pool2d rule) of the fp32 batch-norm output.  Tap byte 255 = the
// pooled value is 0 (ReLU-dead): no gradient.
// Work items are (window, 16-byte chunk) pairs: `cpad` (power of two >= cvec) consecutive threads cover the chunks of
// one window (threads with chunk >= cvec idle), kThreads/cpad windows per CTA step; every CTA walks a contiguous range
// of window groups, so the input row shared by vertically adjacent windows is still in its L1.  32-bit index math
// (the ABI requires N*PH*PW < 2^31); byte offsets are 64-bit.
struct PoolGeom {
  int H, W, PH, PW, cvec, cpad;
  unsigned windows;  // N * PH * PW
};

struct WinPos {
  int n, ph, pw;
};
__device__ __forceinline__ WinPos decode_window(unsigned w, const PoolGeom& g) {
  WinPos p;
  p.pw = static_cast<int>(w % static_cast<unsigned>(g.PW));
  const unsigned t = w / static_cast<unsigned>(g.PW);
  p.ph = static_cast<int>(t % static_cast<unsigned>(g.PH));
  p.n = static_cast<int>(t / static_cast<unsigned>(g.PH));
  return p;
}
__device__ __forceinline__ int64_t in_chunk(const PoolGeom& g, int n, int r, int c, int chunk) {
  return ((static_cast<int64_t>(n) * g.H + r) * g.W + c) * g.cvec + chunk;
}
// contiguous range of window groups of this CTA
__device__ __forceinline__ void cta_group_range(const PoolGeom& g, unsigned& g0, unsigned& g1, unsigned& wpi) {
  wpi = kThreads / g.cpad;
  const unsigned groups = (g.windows + wpi - 1) / wpi;
  const unsigned per = (groups + gridDim.x - 1) / gridDim.x;
  g0 = min(blockIdx.x * per, groups);
  g1 = min(g0 + per, groups);
}

template <int V>
__device__ __forceinline__ void store_taps(unsigned char* __restrict__ tap, int64_t e, const uint32_t* b) {
  if (V == 8) {
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { lo |= b[i] << (8 * i); hi |= b[4 + i] << (8 * i); }
    *reinterpret_cast<uint2*>(tap + e * 8) = make_uint2(lo, hi);
  } else {
    uint32_t lo = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) lo |= b[i] << (8 * i);
    *reinterpret_cast<uint32_t*>(tap + e * 4) = lo;
  }
}
template <int V>
__device__ __forceinline__ uint2 load_tap_words(const unsigned char* __restrict__ tap, int64_t e) {
  if (V == 8) return __ldg(reinterpret_cast<const uint2*>(tap + e * 8));
  return make_uint2(__ldg(reinterpret_cast<const uint32_t*>(tap + e * 4)), 0xffffffffu);
}
__device__ __forceinline__ uint32_t tap_byte(const uint2& t, int k) { return ((k < 4 ? t.x : t.y) >> (8 * (k & 3))) & 255u; }

// packed pair arithmetic on the 16-bit dtypes
template <int DT>
struct Pk;
template <>
struct Pk<MSF_BF16> {
  static constexpr uint32_t kNegInf2 = 0xff80ff80u;
  __device__ static __forceinline__ uint32_t gt_mask(uint32_t a, uint32_t b) {
    return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  }
  __device__ static __forceinline__ uint32_t add(uint32_t a, uint32_t b) {
    const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
};
template <>
struct Pk<MSF_F16> {
  static constexpr uint32_t kNegInf2 = 0xfc00fc00u;
  __device__ static __forceinline__ uint32_t gt_mask(uint32_t a, uint32_t b) {
    return __hgt2_mask(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  }
  __device__ static __forceinline__ uint32_t add(uint32_t a, uint32_t b) {
    const __half2 r = __hadd2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
};

// Forward.  y = relu(max over the window of bn(x)); because bn is monotone in x per channel (increasing for sc > 0,
// decreasing for sc < 0, constant for sc == 0) the arg-max is found on the stored x values themselves: key = x with
// the sign flipped where sc < 0 and zeroed where sc == 0, compared two channels at a time for the 16-bit dtypes.
// Writes the pooled output, the arg-max tap byte (255 = ReLU-dead) and the x value at the arg-max (x_arg), which
// turns the backward reduction into a plain streaming pass over pooled-size tensors.
template <int DT>
__global__ void __launch_bounds__(kThreads) bn_apply_pool_kernel(const char* __restrict__ x, char* __restrict__ y,
                                                                 unsigned char* __restrict__ tap, char* __restrict__ xarg,
                                                                 PoolGeom g, const float* __restrict__ mean,
                                                                 const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta) {
  constexpr int V = Elem<DT>::VEC;
  const int chunk = threadIdx.x % g.cpad;
  if (chunk >= g.cvec) return;
  const unsigned lane_w = threadIdx.x / g.cpad;
  ChanConst<V> k;
  k.load(chunk * V, mean, invstd, gamma, beta);
  unsigned g0, g1, wpi;
  cta_group_range(g, g0, g1, wpi);
  if constexpr (DT == MSF_F32) {
    for (unsigned grp = g0; grp < g1; ++grp) {
      const unsigned w = grp * wpi + lane_w;
      if (w >= g.windows) break;
      const WinPos p = decode_window(w, g);
      uint4 v[9];
      bool ok[9];
      // tap (dr, dc) lives (dr*W + dc)*cvec chunks after tap (0, 0); only in-range taps are dereferenced
      const char* p0 = x + in_chunk(g, p.n, 2 * p.ph - 1, 2 * p.pw - 1, chunk) * 16;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int r = 2 * p.ph - 1 + t / 3, c = 2 * p.pw - 1 + t % 3;
        ok[t] = r >= 0 && r < g.H && c >= 0 && c < g.W;
        if (ok[t]) v[t] = ldg_keep(p0 + static_cast<int64_t>(((t / 3) * g.W + (t % 3)) * g.cvec) * 16);
      }
      float best[V], bx[V];
      uint32_t bi[V];
#pragma unroll
      for (int i = 0; i < V; ++i) { best[i] = -CUDART_INF_F; bx[i] = 0.f; bi[i] = 0; }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        if (!ok[t]) continue;
        float f[V];
        Elem<DT>::unpack(v[t], f);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float key = k.sc[i] > 0.f ? f[i] : (k.sc[i] < 0.f ? -f[i] : 0.f);
          if (key > best[i]) { best[i] = key; bx[i] = f[i]; bi[i] = t; }
        }
      }
      float out[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float o = fmaf(bx[i], k.sc[i], k.sh[i]);
        if (!(o > 0.f)) bi[i] = 255u;
        out[i] = fmaxf(o, 0.f);
      }
      const int64_t e = static_cast<int64_t>(w) * g.cvec + chunk;
      stg_stream(y + e * 16, Elem<DT>::pack(out));
      stg_stream(xarg + e * 16, Elem<DT>::pack(bx));
      store_taps<V>(tap, e, bi);
    }
  } else {
    uint32_t flip[4], keep[4], some_zero = 0u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float lo = k.sc[2 * i], hi = k.sc[2 * i + 1];
      flip[i] = (lo < 0.f ? 0x8000u : 0u) | (hi < 0.f ? 0x80000000u : 0u);
      keep[i] = (lo != 0.f ? 0xffffu : 0u) | (hi != 0.f ? 0xffff0000u : 0u);
      some_zero |= ~keep[i];
      // opaque to the optimiser: otherwise it re-derives the masks from sc (8 FSETP + 8 SEL) for every tap
      asm volatile("" : "+r"(flip[i]), "+r"(keep[i]));
    }
    asm volatile("" : "+r"(some_zero));
    const bool has_zero = some_zero != 0u;  // a channel with sc == 0 (gamma == 0): x_arg cannot be recovered from the key
    // window coordinates advance incrementally (w grows by wpi per step): no division in the loop
    unsigned w = g0 * wpi + lane_w;
    WinPos p = decode_window(min(w, g.windows - 1), g);
    const unsigned step_pw = wpi % static_cast<unsigned>(g.PW), step_ph = wpi / static_cast<unsigned>(g.PW);  // wpi <= 256
    const int64_t row_chunks = static_cast<int64_t>(g.W) * g.cvec;
    auto one_tap = [&](const uint4& v, int t, uint32_t* best, uint32_t* bi, uint32_t* bx, bool track_x) {
      const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
      const uint32_t tc = static_cast<uint32_t>(t) * 0x00010001u;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t key = (wv[i] ^ flip[i]) & keep[i];
        const uint32_t m = Pk<DT>::gt_mask(key, best[i]);
        best[i] = (key & m) | (best[i] & ~m);
        bi[i] = (tc & m) | (bi[i] & ~m);
        if (track_x) bx[i] = (wv[i] & m) | (bx[i] & ~m);
      }
    };
    for (unsigned grp = g0; grp < g1; ++grp, w += wpi) {
      if (w >= g.windows) break;
      uint32_t best[4], bx[4], bi[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { best[i] = Pk<DT>::kNegInf2; bx[i] = 0u; bi[i] = 0u; }
      // tap (dr, dc) lives (dr*W + dc)*cvec chunks after tap (0, 0); only in-range taps are dereferenced
      const char* p0 = x + in_chunk(g, p.n, 2 * p.ph - 1, 2 * p.pw - 1, chunk) * 16;
      const bool interior = p.ph > 0 && p.pw > 0 && 2 * p.ph + 1 < g.H && 2 * p.pw + 1 < g.W;
      if (interior && !has_zero) {  // the common case: nine unconditional loads, three logic ops + one compare per pair
        uint4 v[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) v[t] = ldg_keep(p0 + ((t / 3) * row_chunks + (t % 3) * g.cvec) * 16);
#pragma unroll
        for (int t = 0; t < 9; ++t) one_tap(v[t], t, best, bi, bx, false);
      } else {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int r = 2 * p.ph - 1 + t / 3, c = 2 * p.pw - 1 + t % 3;
          if (r >= 0 && r < g.H && c >= 0 && c < g.W) one_tap(ldg_keep(p0 + ((t / 3) * row_chunks + (t % 3) * g.cvec) * 16), t, best, bi, bx, true);
        }
      }
      if (interior && !has_zero) {
#pragma unroll
        for (int i = 0; i < 4; ++i) bx[i] = best[i] ^ flip[i];
      }
      float f[V], out[V];
      Elem<DT>::unpack(make_uint4(bx[0], bx[1], bx[2], bx[3]), f);
      uint32_t tb[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float o = fmaf(f[i], k.sc[i], k.sh[i]);
        tb[i] = o > 0.f ? ((bi[i >> 1] >> (16 * (i & 1))) & 255u) : 255u;
        out[i] = fmaxf(o, 0.f);
      }
      const int64_t e = static_cast<int64_t>(w) * g.cvec + chunk;
      stg_stream(y + e * 16, Elem<DT>::pack(out));
      stg_stream(xarg + e * 16, make_uint4(bx[0], bx[1], bx[2], bx[3]));
      store_taps<V>(tap, e, tb);
      // advance (n, ph, pw) by wpi windows
      p.pw += static_cast<int>(step_pw);
      p.ph += static_cast<int>(step_ph);
      if (p.pw >= g.PW) { p.pw -= g.PW; ++p.ph; }
      while (p.ph >= g.PH) { p.ph -= g.PH; ++p.n; }
    }
  }
}

// Input gradient through pool + relu + bn.  One thread owns the 2x2 input positions (2a+i, 2b+j) of block (a, b) for
// one chunk.  They are covered by the windows (a+da, b+db), da, db in {0,1}: position (i, j) is tap
// (i-2da+1)*3 + (j-2db+1) of window (da, db) when i >= da and j >= db.  dy' of a position = sum of the pooled gradients
// of the windows whose arg-max it is (for the 16-bit dtypes selected with byte compares and summed as packed pairs).
template <int DT>
__global__ void __launch_bounds__(kThreads, 2) bn_pool_bwd_elemt_kernel(const char* __restrict__ x, const char* __restrict__ dpool,
                                                                        const unsigned char* __restrict__ tap, char* __restrict__ dx,
                                                                        PoolGeom g, const float* __restrict__ mean,
                                                                        const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                                        const double* __restrict__ sums, const double* __restrict__ count) {
  constexpr int V = Elem<DT>::VEC;
  const int chunk = threadIdx.x % g.cpad;
  if (chunk >= g.cvec) return;
  const unsigned lane_w = threadIdx.x / g.cpad;
  GradConst<V> gc;
  gc.load(chunk * V, g.cvec * V, mean, invstd, gamma, sums, static_cast<float>(1.0 / *count));
  unsigned g0, g1, wpi;
  cta_group_range(g, g0, g1, wpi);
  unsigned w = g0 * wpi + lane_w;
  WinPos p = decode_window(min(w, g.windows - 1), g);  // advanced incrementally below: no division in the loop
  const unsigned step_pw = wpi % static_cast<unsigned>(g.PW), step_ph = wpi / static_cast<unsigned>(g.PW);
  for (unsigned grp = g0; grp < g1; ++grp, w += wpi) {
    if (w >= g.windows) break;
    uint4 vx[2][2], vd[2][2];
    uint2 tb[2][2];
    bool okx[2][2];
    const int64_t xe = in_chunk(g, p.n, 2 * p.ph, 2 * p.pw, chunk);   // chunk index of position (0, 0) of the block
    const int64_t we0 = static_cast<int64_t>(w) * g.cvec + chunk;     // chunk index of window (a, b)
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        okx[i][j] = 2 * p.ph + i < g.H && 2 * p.pw + j < g.W;
        if (okx[i][j]) vx[i][j] = ldg_stream(x + (xe + (i * g.W + j) * g.cvec) * 16);
        if (p.ph + i < g.PH && p.pw + j < g.PW) {
          const int64_t we = we0 + (i * g.PW + j) * g.cvec;
          vd[i][j] = ldg_keep(dpool + we * 16);
          tb[i][j] = load_tap_words<V>(tap, we);
        } else {
          vd[i][j] = make_uint4(0, 0, 0, 0);
          tb[i][j] = make_uint2(0xffffffffu, 0xffffffffu);
        }
      }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (!okx[i][j]) continue;
        float d[V];
        if constexpr (DT == MSF_F32) {
#pragma unroll
          for (int k = 0; k < V; ++k) d[k] = 0.f;
#pragma unroll
          for (int da = 0; da <= i; ++da)
#pragma unroll
            for (int db = 0; db <= j; ++db) {
              const uint32_t me = static_cast<uint32_t>((i - 2 * da + 1) * 3 + (j - 2 * db + 1));
              float fd[V];
              Elem<DT>::unpack(vd[da][db], fd);
#pragma unroll
              for (int k = 0; k < V; ++k)
                if (tap_byte(tb[da][db], k) == me) d[k] += fd[k];
            }
        } else {
          uint32_t acc[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int da = 0; da <= i; ++da)
#pragma unroll
            for (int db = 0; db <= j; ++db) {
              const uint32_t me = static_cast<uint32_t>((i - 2 * da + 1) * 3 + (j - 2 * db + 1)) * 0x01010101u;
              const uint32_t mlo = __vcmpeq4(tb[da][db].x, me), mhi = __vcmpeq4(tb[da][db].y, me);
              const uint32_t wd[4] = {vd[da][db].x, vd[da][db].y, vd[da][db].z, vd[da][db].w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {  // byte mask of channels 2q, 2q+1 -> half-word mask
                const uint32_t hm = __byte_perm(q < 2 ? mlo : mhi, 0u, (q & 1) ? 0x3322u : 0x1100u);
                acc[q] = Pk<DT>::add(acc[q], wd[q] & hm);
              }
            }
          Elem<DT>::unpack(make_uint4(acc[0], acc[1], acc[2], acc[3]), d);
        }
        float fx[V];
        Elem<DT>::unpack(vx[i][j], fx);
#pragma unroll
        for (int k = 0; k < V; ++k) fx[k] = fmaf(gc.sc[k], d[k], fmaf(gc.ca[k], fx[k], gc.cb[k]));
        stg_stream(dx + (xe + (i * g.W + j) * g.cvec) * 16, Elem<DT>::pack(fx));
      }
    p.pw += static_cast<int>(step_pw);
    p.ph += static_cast<int>(step_ph);
    if (p.pw >= g.PW) { p.pw -= g.PW; ++p.ph; }
    while (p.ph >= g.PH) { p.ph -= g.PH; ++p.n; }
  }
}

int check_bn(const void* x, int64_t rows, int C, int dtype) {
  MSF_REQUIRE(dtype_ok(dtype), MSF_ERR_INVALID, "bad dtype %d", dtype);
  const int vec = 16 / static_cast<int>(dtype_size(dtype));
  MSF_REQUIRE(rows > 0 && C > 0 && C % vec == 0, MSF_ERR_INVALID, "C=%d must be a positive multiple of %d", C, vec);
  MSF_REQUIRE(x && aligned16(x), MSF_ERR_INVALID, "NULL or misaligned activation pointer");
  return MSF_OK;
}

// grid of a grid-stride streaming kernel: exactly one wave of resident CTAs (no wave-quantisation tail), never more than
// the work needs.  Any grid size keeps the per-thread channel constants valid: a thread's chunk index advances by
// gridDim.x * kThreads, a multiple of cvec whenever kThreads is.
template <typename Kern>
inline unsigned wave_grid(Kern kern, int64_t per_thread_items, int cvec) {
  (void)cvec;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, 0) != cudaSuccess || occ < 1) occ = 1;
  int64_t g = static_cast<int64_t>(kNumSMs) * occ;
  const int64_t need = (per_thread_items + kThreads - 1) / kThreads;
  if (g > need) g = need;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

template <typename Kern>
inline unsigned resident_grid(Kern kern, int64_t max_useful) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, 0) != cudaSuccess || occ < 1) occ = 1;
  int64_t g = static_cast<int64_t>(kNumSMs) * occ;
  if (g > max_useful) g = max_useful;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

}  // namespace
}  // namespace msf

using namespace msf;

extern "C" size_t msf_bn2d_workspace_bytes(int64_t rows, int C) {
  if (rows <= 0 || C <= 0) return 0;
  return static_cast<size_t>(kMaxRowBlocks) * (3 * C + 1) * sizeof(float);  // partial[row blocks][3][C] + cnt[row blocks], one wave of CTAs at most
}

namespace {
int bn_stats_impl(const void* x, int64_t rows, int C, int dtype, double* sums_out, void* workspace, size_t workspace_bytes,
                  void* stream, bool finalize, float eps, float momentum, float* mean, float* invstd, float* running_mean,
                  float* running_var);
}

extern "C" int msf_bn2d_stats(const void* x, int64_t rows, int C, int dtype, double* sums_out, void* workspace,
                              size_t workspace_bytes, void* stream) {
  return bn_stats_impl(x, rows, C, dtype, sums_out, workspace, workspace_bytes, stream, false, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int msf_bn2d_stats_finalize(const void* x, int64_t rows, int C, int dtype, float eps, float momentum, double* sums_out,
                                       float* mean, float* invstd, float* running_mean, float* running_var, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  MSF_REQUIRE(mean && invstd, MSF_ERR_INVALID, "mean / invstd are NULL");
  MSF_REQUIRE((running_mean == nullptr) == (running_var == nullptr), MSF_ERR_INVALID, "running_mean / running_var must both be given or both be NULL");
  return bn_stats_impl(x, rows, C, dtype, sums_out, workspace, workspace_bytes, stream, true, eps, momentum, mean, invstd, running_mean,
                       running_var);
}

namespace {
int bn_stats_impl(const void* x, int64_t rows, int C, int dtype, double* sums_out, void* workspace, size_t workspace_bytes,
                  void* stream, bool finalize, float eps, float momentum, float* mean, float* invstd, float* running_mean,
                  float* running_var) {
  if (int rc = check_bn(x, rows, C, dtype)) return rc;
  MSF_REQUIRE(sums_out && workspace && workspace_bytes >= msf_bn2d_workspace_bytes(rows, C), MSF_ERR_WORKSPACE, "workspace too small");
  const int vec = 16 / static_cast<int>(dtype_size(dtype));
  const Layout l = make_layout(C, vec);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(kThreads) * 2 * vec * sizeof(float);
  float* partial = static_cast<float*>(workspace);
  float* cnt = partial + static_cast<size_t>(kMaxRowBlocks) * 3 * C;
  int nblk = 1;
  ProfScope prof(stream, MSF_K_BN_STATS, static_cast<double>(rows) * C * dtype_size(dtype));
  MSF_DISPATCH_DTYPE(dtype, {
    nblk = reduce_grid(bn_stats_kernel<DT>, smem, l, rows, l.rlanes);
    if (nblk > kStatsMaxBlocks) nblk = kStatsMaxBlocks;
    dim3 grid(static_cast<unsigned>(nblk), static_cast<unsigned>(l.cgroups));
    bn_stats_kernel<DT><<<grid, kThreads, smem, st>>>(static_cast<const char*>(x), rows, l.cvec, l.ct, partial, cnt);
  });
  MSF_LAUNCH_OK("bn_stats_kernel");
  if (finalize) {
    bn_combine_finalize_kernel<<<(C + kCombineWarps - 1) / kCombineWarps, kCombineWarps * 32, 0, st>>>(
        partial, cnt, nblk, C, sums_out, static_cast<double>(rows), eps, momentum, mean, invstd, running_mean, running_var);
  } else {
    bn_stats_combine_kernel<<<(C + kCombineWarps - 1) / kCombineWarps, kCombineWarps * 32, 0, st>>>(partial, cnt, nblk, C, sums_out,
                                                                                                   static_cast<double>(rows));
  }
  MSF_LAUNCH_OK("bn_combine_kernel");
  return MSF_OK;
}
}  // namespace

extern "C" int msf_bn2d_finalize(const double* sums, int C, float eps, float momentum, float* mean, float* invstd,
                                 float* running_mean, float* running_var, void* stream) {
  MSF_REQUIRE(sums && mean && invstd && C > 0, MSF_ERR_INVALID, "bad arguments");
  MSF_REQUIRE((running_mean == nullptr) == (running_var == nullptr), MSF_ERR_INVALID, "running_mean / running_var must both be given or both be NULL");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(sums, C, eps, momentum, mean, invstd,
                                                                                     running_mean, running_var);
  MSF_LAUNCH_OK("bn_finalize_kernel");
  return MSF_OK;
}

extern "C" int msf_bn2d_apply(const void* x, const void* res, void* y, uint8_t* relu_bits, int64_t rows, int C, int dtype,
                              const float* mean, const float* invstd, const float* gamma, const float* beta, int relu, void* stream) {
  if (int rc = check_bn(x, rows, C, dtype)) return rc;
  MSF_REQUIRE(y && aligned16(y) && aligned16(res) && mean && invstd, MSF_ERR_INVALID, "bad arguments");
  const int vec = 16 / static_cast<int>(dtype_size(dtype));
  const int cvec = C / vec;
  const int64_t chunks = rows * cvec;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const char* xp = static_cast<const char*>(x);
  const char* rp = static_cast<const char*>(res);
  char* yp = static_cast<char*>(y);
  ProfScope prof(stream, res ? MSF_K_BN_APPLY_RES : MSF_K_BN_APPLY, static_cast<double>(rows) * C * dtype_size(dtype) * (res ? 3 : 2) + (relu_bits ? static_cast<double>(chunks) : 0.0));
  if (res) {
    MSF_DISPATCH_DTYPE(dtype, (bn_apply_kernel<DT, true><<<wave_grid(bn_apply_kernel<DT, true>, (chunks + 1) / 2, cvec), kThreads, 0, st>>>(xp, rp, yp, relu_bits, chunks, cvec, mean, invstd, gamma, beta, relu)));
  } else {
    MSF_DISPATCH_DTYPE(dtype, (bn_apply_kernel<DT, false><<<wave_grid(bn_apply_kernel<DT, false>, (chunks + 3) / 4, cvec), kThreads, 0, st>>>(xp, rp, yp, relu_bits, chunks, cvec, mean, invstd, gamma, beta, relu)));
  }
  MSF_LAUNCH_OK("bn_apply_kernel");
  return MSF_OK;
}

namespace {
// the optional pooled-gradient term: gp (N, C) in the activation dtype, hw = rows per image
int check_gp(const void* gp, int64_t hw, int64_t rows, int C, int dtype, const void* y_mask, int relu) {
  if (!gp) return MSF_OK;
  MSF_REQUIRE(aligned16(gp) && hw > 0 && rows % hw == 0, MSF_ERR_INVALID, "pooled gradient: rows=%lld is not a multiple of hw=%lld",
              static_cast<long long>(rows), static_cast<long long>(hw));
  MSF_REQUIRE(relu && y_mask, MSF_ERR_UNSUPPORTED, "the pooled-gradient term is implemented for the relu + saved-output-mask variant");
  const int64_t chunks = rows * (C / (16 / static_cast<int>(dtype_size(dtype))));
  MSF_REQUIRE(chunks < (int64_t{1} << 32), MSF_ERR_UNSUPPORTED, "pooled-gradient variant needs rows*C/vec < 2^32");
  return MSF_OK;
}
}  // namespace

extern "C" int msf_bn2d_bwd_reduce(const void* x, const void* dy, const void* y_mask, int mask_is_bits, int64_t rows, int C, int dtype,
                                   const float* mean, const float* invstd, const float* gamma, const float* beta, int relu,
                                   const void* gpool, int64_t hw, double* sums_out, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  if (int rc = check_bn(x, rows, C, dtype)) return rc;
  MSF_REQUIRE(dy && aligned16(dy) && (mask_is_bits || aligned16(y_mask)) && mean && invstd && sums_out, MSF_ERR_INVALID, "bad arguments");
  MSF_REQUIRE(!mask_is_bits || (y_mask && relu), MSF_ERR_INVALID, "mask_is_bits needs relu and a mask pointer");
  if (int rc = check_gp(gpool, hw, rows, C, dtype, y_mask, relu)) return rc;
  const char* gpp = static_cast<const char*>(gpool);
  const unsigned hw32 = gpool ? static_cast<unsigned>(hw) : 1u;
  const float gp_scale = gpool ? 1.f / static_cast<float>(hw) : 0.f;
  MSF_REQUIRE(workspace && workspace_bytes >= msf_bn2d_workspace_bytes(rows, C), MSF_ERR_WORKSPACE, "workspace too small");
  const int vec = 16 / static_cast<int>(dtype_size(dtype));
  const Layout l = make_layout(C, vec);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(kThreads) * 2 * vec * sizeof(float);
  float* partial = static_cast<float*>(workspace);
  const char* xp = static_cast<const char*>(x);
  const char* dp = static_cast<const char*>(dy);
  const char* mp = static_cast<const char*>(y_mask);
  int nblk = 1;
  ProfScope prof(stream, MSF_K_BN_BWD_REDUCE, static_cast<double>(rows) * C * (dtype_size(dtype) * ((relu && y_mask && !mask_is_bits) ? 3 : 2) +
                                                                             (mask_is_bits ? 1.0 / vec : 0.0)));
#define MSF_BWD_REDUCE(MASK, GP)                                                                                      \
  MSF_DISPATCH_DTYPE(dtype, {                                                                                         \
    nblk = reduce_grid(bn_bwd_reduce_kernel<DT, MASK, GP>, smem, l, rows, kReduceRows * l.rlanes);                    \
    dim3 grid(static_cast<unsigned>(nblk), static_cast<unsigned>(l.cgroups));                                         \
    bn_bwd_reduce_kernel<DT, MASK, GP><<<grid, kThreads, smem, st>>>(xp, dp, mp, rows, l.cvec, l.ct, mean, invstd, gamma, beta, gpp, \
                                                                     hw32, gp_scale, partial);                       \
  })
  if (!relu) { MSF_BWD_REDUCE(0, false); }
  else if (!y_mask) { MSF_BWD_REDUCE(1, false); }
  else if (!mask_is_bits && !gpool) { MSF_BWD_REDUCE(2, false); }
  else if (!mask_is_bits) { MSF_BWD_REDUCE(2, true); }
  else if (!gpool) { MSF_BWD_REDUCE(3, false); }
  else { MSF_BWD_REDUCE(3, true); }
#undef MSF_BWD_REDUCE
  MSF_LAUNCH_OK("bn_bwd_reduce_kernel");
  bn_combine_kernel<<<(2 * C + kCombineWarps - 1) / kCombineWarps, kCombineWarps * 32, 0, st>>>(partial, nblk, C, sums_out, -1.0);
  MSF_LAUNCH_OK("bn_combine_kernel");
  return MSF_OK;
}

extern "C" int msf_bn2d_bwd_elemt(const void* x, const void* dy, const void* y_mask, int mask_is_bits, void* dx, void* dres, int64_t rows,
                                  int C, int dtype, const float* mean, const float* invstd, const float* gamma, const float* beta,
                                  int relu, const void* gpool, int64_t hw, const double* sums, const double* count, void* stream) {
  if (int rc = check_bn(x, rows, C, dtype)) return rc;
  if (int rc = check_gp(gpool, hw, rows, C, dtype, y_mask, relu)) return rc;
  const char* gpp = static_cast<const char*>(gpool);
  const float gp_scale = gpool ? 1.f / static_cast<float>(hw) : 0.f;
  MSF_REQUIRE(dy && dx && aligned16(dy) && aligned16(dx) && (mask_is_bits || aligned16(y_mask)) && aligned16(dres) && mean && invstd && sums && count,
              MSF_ERR_INVALID, "bad arguments");
  MSF_REQUIRE(!mask_is_bits || (y_mask && relu), MSF_ERR_INVALID, "mask_is_bits needs relu and a mask pointer");
  const int vec = 16 / static_cast<int>(dtype_size(dtype));
  const int cvec = C / vec;
  const int64_t chunks = rows * cvec;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t per_thread_total = (chunks + 1) / 2;
  const char* xp = static_cast<const char*>(x);
  const char* dp = static_cast<const char*>(dy);
  const char* mp = static_cast<const char*>(y_mask);
  char* dxp = static_cast<char*>(dx);
  char* drp = static_cast<char*>(dres);
  ProfScope prof(stream, MSF_K_BN_BWD_ELEMT, static_cast<double>(rows) * C * (dtype_size(dtype) * (3 + ((relu && y_mask && !mask_is_bits) ? 1 : 0) + (dres ? 1 : 0)) +
                                                                            (mask_is_bits ? 1.0 / vec : 0.0)));
  const unsigned hwc = gpool ? static_cast<unsigned>(hw * cvec) : 1u;
#define MSF_BWD_ELEMT(MASK, DRES, GP) \
  MSF_DISPATCH_DTYPE(dtype, (bn_bwd_elemt_kernel<DT, MASK, DRES, GP><<<wave_grid(bn_bwd_elemt_kernel<DT, MASK, DRES, GP>, per_thread_total, cvec), kThreads, 0, st>>>(xp, dp, mp, dxp, drp, chunks, cvec, mean, invstd, gamma, beta, sums, count, gpp, hwc, gp_scale)))
  const int mask = !relu ? 0 : (y_mask ? (mask_is_bits ? 3 : 2) : 1);
  if (mask == 0 && !dres) { MSF_BWD_ELEMT(0, false, false); }
  else if (mask == 0) { MSF_BWD_ELEMT(0, true, false); }
  else if (mask == 1 && !dres) { MSF_BWD_ELEMT(1, false, false); }
  else if (mask == 1) { MSF_BWD_ELEMT(1, true, false); }
  else if (mask == 2 && !dres && !gpool) { MSF_BWD_ELEMT(2, false, false); }
  else if (mask == 2 && !gpool) { MSF_BWD_ELEMT(2, true, false); }
  else if (mask == 2 && !dres) { MSF_BWD_ELEMT(2, false, true); }
  else if (mask == 2) { MSF_BWD_ELEMT(2, true, true); }
  else if (!dres && !gpool) { MSF_BWD_ELEMT(3, false, false); }
  else if (!gpool) { MSF_BWD_ELEMT(3, true, false); }
  else if (!dres) { MSF_BWD_ELEMT(3, false, true); }
  else { MSF_BWD_ELEMT(3, true, true); }
#undef MSF_BWD_ELEMT
  MSF_LAUNCH_OK("bn_bwd_elemt_kernel");
  return MSF_OK;
}

namespace {
int check_pool(const void* x, int64_t N, int H, int W, int C, int dtype, PoolGeom* g) {
  if (int rc = check_bn(x, N, C, dtype)) return rc;
  MSF_REQUIRE(H >= 2 && W >= 2, MSF_ERR_INVALID, "pooling needs H, W >= 2 (got %d x %d)", H, W);
  const int vec = 16 / static_cast<int>(dtype_size(dtype));
  g->H = H; g->W = W; g->PH = (H - 1) / 2 + 1; g->PW = (W - 1) / 2 + 1; g->cvec = C / vec;
  g->cpad = 1;
  while (g->cpad < g->cvec) g->cpad *= 2;
  MSF_REQUIRE(g->cpad <= kThreads, MSF_ERR_UNSUPPORTED, "C=%d is too wide for the pooled kernels", C);
  const int64_t windows = N * g->PH * g->PW;
  MSF_REQUIRE(windows < (int64_t{1} << 31), MSF_ERR_UNSUPPORTED, "N*PH*PW = %lld must be < 2^31", static_cast<long long>(windows));
  g->windows = static_cast<unsigned>(windows);
  return MSF_OK;
}
}  // namespace

extern "C" int msf_bn2d_apply_pool(const void* x, void* y, uint8_t* tap, void* x_arg, int64_t N, int H, int W, int C, int dtype,
                                   const float* mean, const float* invstd, const float* gamma, const float* beta, void* stream) {
  PoolGeom g;
  if (int rc = check_pool(x, N, H, W, C, dtype, &g)) return rc;
  MSF_REQUIRE(y && tap && x_arg && aligned16(y) && aligned16(tap) && aligned16(x_arg) && mean && invstd, MSF_ERR_INVALID, "bad arguments");
  const int64_t groups = (static_cast<int64_t>(g.windows) + kThreads / g.cpad - 1) / (kThreads / g.cpad);
  // x read once; y, x_arg and one tap byte written per pooled element
  ProfScope prof(stream, MSF_K_BN_APPLY_POOL, (static_cast<double>(N) * H * W + 2.0 * g.windows) * C * dtype_size(dtype) + static_cast<double>(g.windows) * C);
  MSF_DISPATCH_DTYPE(dtype, (bn_apply_pool_kernel<DT><<<resident_grid(bn_apply_pool_kernel<DT>, groups), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
                                static_cast<const char*>(x), static_cast<char*>(y), tap, static_cast<char*>(x_arg), g, mean, invstd, gamma, beta)));
  MSF_LAUNCH_OK("bn_apply_pool_kernel");
  return MSF_OK;
}

extern "C" int msf_bn2d_pool_bwd_elemt(const void* x, const void* dpool, const uint8_t* tap, void* dx, int64_t N, int H, int W,
                                       int C, int dtype, const float* mean, const float* invstd, const float* gamma,
                                       const double* sums, const double* count, void* stream) {
  PoolGeom g;
  if (int rc = check_pool(x, N, H, W, C, dtype, &g)) return rc;
  MSF_REQUIRE(dpool && tap && dx && aligned16(dpool) && aligned16(dx) && mean && invstd && sums && count, MSF_ERR_INVALID, "bad arguments");
  const int64_t groups = (static_cast<int64_t>(g.windows) + kThreads / g.cpad - 1) / (kThreads / g.cpad);
  ProfScope prof(stream, MSF_K_BN_POOL_BWD_ELEMT, (2.0 * N * H * W + static_cast<double>(g.windows)) * C * dtype_size(dtype) + static_cast<double>(g.windows) * C);
  MSF_DISPATCH_DTYPE(dtype, (bn_pool_bwd_elemt_kernel<DT><<<resident_grid(bn_pool_bwd_elemt_kernel<DT>, groups), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
                                static_cast<const char*>(x), static_cast<const char*>(dpool), tap, static_cast<char*>(dx), g, mean, invstd,
                                gamma, sums, count)));
  MSF_LAUNCH_OK("bn_pool_bwd_elemt_kernel");
  return MSF_OK;
}
