// L1 (cosine mode): the 24 nn.CosineSimilarity(dim=1)+mean() calls of tools/ssl_train.py:422,448-466
// as one forward launch (+ one 1-CTA deterministic final sum) and one backward launch.
//
// HBM-bound streaming kernel: a row is owned by a sub-warp group of 8/16/32 lanes, every lane
// issues 128-bit loads, all arithmetic in fp32 (CUDA autocast runs cosine_similarity in fp32).
// Algorithmic bytes: forward rows*dim*(e_p+e_z) read + 16 B/row stats; backward the same read
// + rows*dim*e_p written.
#include "common.cuh"

namespace msf {
namespace {

struct Pair {
  const char* p;
  const char* z;
  float4* stats;
  char* grad_p;
  int64_t rows;
  uint32_t cpr;            // 16-byte chunks per row
  uint32_t lanes;          // lanes per row (8, 16 or 32)
  uint32_t rows_per_block;
  uint32_t block_prefix;
  float scale;             // coef / rows
};

struct Params {
  Pair pair[MSF_COS_MAX_PAIRS];
  int n_pairs;
  uint32_t total_blocks;
  float eps;
};

constexpr int kThreads = 256;

template <int DT>
__device__ __forceinline__ void row_dots(const char* p, const char* z, uint32_t cpr, uint32_t lane, uint32_t lanes,
                                         bool valid, float& dot, float& pp, float& zz) {
  constexpr int VEC = Elem<DT>::VEC;
  dot = pp = zz = 0.f;
  uint32_t c = valid ? lane : cpr;  // invalid rows load nothing but still take part in the shuffles
  for (; c + 3 * lanes < cpr; c += 4 * lanes) {  // four chunks of each operand in flight
    uint4 pa[4], za[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      pa[u] = ldg_stream(p + static_cast<size_t>(c + u * lanes) * 16);
      za[u] = ldg_stream(z + static_cast<size_t>(c + u * lanes) * 16);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float fa[VEC], fb[VEC];
      Elem<DT>::unpack(pa[u], fa);
      Elem<DT>::unpack(za[u], fb);
#pragma unroll
      for (int i = 0; i < VEC; ++i) { dot = fmaf(fa[i], fb[i], dot); pp = fmaf(fa[i], fa[i], pp); zz = fmaf(fb[i], fb[i], zz); }
    }
  }
  for (; c + lanes < cpr; c += 2 * lanes) {  // two chunks of each operand in flight
    const uint4 a0 = ldg_stream(p + static_cast<size_t>(c) * 16), b0 = ldg_stream(z + static_cast<size_t>(c) * 16);
    const uint4 a1 = ldg_stream(p + static_cast<size_t>(c + lanes) * 16), b1 = ldg_stream(z + static_cast<size_t>(c + lanes) * 16);
    float fa[VEC], fb[VEC];
    Elem<DT>::unpack(a0, fa);
    Elem<DT>::unpack(b0, fb);
#pragma unroll
    for (int i = 0; i < VEC; ++i) { dot = fmaf(fa[i], fb[i], dot); pp = fmaf(fa[i], fa[i], pp); zz = fmaf(fb[i], fb[i], zz); }
    Elem<DT>::unpack(a1, fa);
    Elem<DT>::unpack(b1, fb);
#pragma unroll
    for (int i = 0; i < VEC; ++i) { dot = fmaf(fa[i], fb[i], dot); pp = fmaf(fa[i], fa[i], pp); zz = fmaf(fb[i], fb[i], zz); }
  }
  if (c < cpr) {
    const uint4 a0 = ldg_stream(p + static_cast<size_t>(c) * 16), b0 = ldg_stream(z + static_cast<size_t>(c) * 16);
    float fa[VEC], fb[VEC];
    Elem<DT>::unpack(a0, fa);
    Elem<DT>::unpack(b0, fb);
#pragma unroll
    for (int i = 0; i < VEC; ++i) { dot = fmaf(fa[i], fb[i], dot); pp = fmaf(fa[i], fa[i], pp); zz = fmaf(fb[i], fb[i], zz); }
  }
  dot = group_sum(dot, lanes);
  pp = group_sum(pp, lanes);
  zz = group_sum(zz, lanes);
}

__device__ __forceinline__ const Pair& find_pair(const Params& P) {
  int s = 0;
#pragma unroll 1
  while (s + 1 < P.n_pairs && blockIdx.x >= P.pair[s + 1].block_prefix) ++s;
  return P.pair[s];
}

template <int DT>
__global__ void __launch_bounds__(kThreads) cos_fwd_kernel(const __grid_constant__ Params P, float* partials) {
  const Pair& g = find_pair(P);
  const uint32_t lanes = g.lanes, lane = threadIdx.x & (lanes - 1);
  const uint32_t groups = kThreads / lanes, grp = threadIdx.x / lanes;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x - g.block_prefix) * g.rows_per_block;
  const size_t row_bytes = static_cast<size_t>(g.cpr) * 16;
  float acc = 0.f;
  for (uint32_t r = grp; r < g.rows_per_block; r += groups) {
    const int64_t row = row0 + r;
    const bool valid = row < g.rows;  // trip count is CTA-uniform (rows_per_block % groups == 0)
    float dot, pp, zz;
    row_dots<DT>(g.p + row * row_bytes, g.z + row * row_bytes, g.cpr, lane, lanes, valid, dot, pp, zz);
    // ATen cosine_similarity: (p / max(||p||,eps)) . (z / max(||z||,eps))
    const float pn = sqrtf(pp), zn = sqrtf(zz);
    const float ipn = 1.f / fmaxf(pn, P.eps), izn = 1.f / fmaxf(zn, P.eps);
    const float cosv = dot * ipn * izn;
    if (lane == 0 && valid) {
      g.stats[row] = make_float4(cosv, ipn, izn, pn >= P.eps ? 1.f : 0.f);
      acc += cosv;
    }
  }
  // deterministic block reduction (fixed tree), one partial per CTA
  __shared__ float sm[kThreads / 32];
  float w = group_sum(acc, 32);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) t += sm[i];
    partials[blockIdx.x] = t * g.scale;
  }
}

__global__ void __launch_bounds__(256) cos_final_kernel(const float* __restrict__ partials, uint32_t n, float* loss_out) {
  __shared__ double sm[256];
  double t = 0.0;
  for (uint32_t i = threadIdx.x; i < n; i += 256) t += static_cast<double>(partials[i]);
  sm[threadIdx.x] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_out = static_cast<float>(sm[0]);
}

template <int DT>
__global__ void __launch_bounds__(kThreads) cos_bwd_kernel(const __grid_constant__ Params P, const float* __restrict__ grad_out) {
  constexpr int VEC = Elem<DT>::VEC;
  const Pair& g = find_pair(P);
  const uint32_t lanes = g.lanes, lane = threadIdx.x & (lanes - 1);
  const uint32_t groups = kThreads / lanes, grp = threadIdx.x / lanes;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x - g.block_prefix) * g.rows_per_block;
  const size_t row_bytes = static_cast<size_t>(g.cpr) * 16;
  const float s = g.scale * __ldg(grad_out);
  for (uint32_t r = grp; r < g.rows_per_block; r += groups) {
    const int64_t row = row0 + r;
    if (row >= g.rows) break;
    const float4 st = __ldg(g.stats + row);  // {cos, 1/|p|c, 1/|z|c, active}
    const float a = s * st.y * st.z;           // multiplies z
    const float b = s * st.x * st.y * st.y * st.w;  // multiplies p (0 where the clamp is active)
    const char* p = g.p + row * row_bytes;
    const char* z = g.z + row * row_bytes;
    char* o = g.grad_p + row * row_bytes;
    uint32_t c = lane;
    for (; c + 3 * lanes < g.cpr; c += 4 * lanes) {  // 8 x 128-bit loads in flight per thread
      uint4 pv[4], zv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        pv[u] = ldg_stream(p + static_cast<size_t>(c + u * lanes) * 16);
        zv[u] = ldg_stream(z + static_cast<size_t>(c + u * lanes) * 16);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float fp[VEC], fz[VEC];
        Elem<DT>::unpack(pv[u], fp);
        Elem<DT>::unpack(zv[u], fz);
#pragma unroll
        for (int i = 0; i < VEC; ++i) fp[i] = a * fz[i] - b * fp[i];
        stg_stream(o + static_cast<size_t>(c + u * lanes) * 16, Elem<DT>::pack(fp));
      }
    }
    for (; c < g.cpr; c += lanes) {
      const uint4 pv = ldg_stream(p + static_cast<size_t>(c) * 16), zv = ldg_stream(z + static_cast<size_t>(c) * 16);
      float fp[VEC], fz[VEC];
      Elem<DT>::unpack(pv, fp);
      Elem<DT>::unpack(zv, fz);
#pragma unroll
      for (int i = 0; i < VEC; ++i) fp[i] = a * fz[i] - b * fp[i];
      stg_stream(o + static_cast<size_t>(c) * 16, Elem<DT>::pack(fp));
    }
  }
}

int build(Params& P, const msf_cos_pair* pairs, int n_pairs, int dtype, bool need_grad) {
  MSF_REQUIRE(n_pairs >= 0 && n_pairs <= MSF_COS_MAX_PAIRS, MSF_ERR_INVALID, "n_pairs %d outside [0,%d]", n_pairs,
              MSF_COS_MAX_PAIRS);
  MSF_REQUIRE(pairs || n_pairs == 0, MSF_ERR_INVALID, "pairs is NULL");
  MSF_REQUIRE(dtype_ok(dtype), MSF_ERR_INVALID, "bad dtype %d", dtype);
  const uint32_t vec = 16 / dtype_size(dtype);
  P.n_pairs = 0;
  P.total_blocks = 0;
  for (int i = 0; i < n_pairs; ++i) {
    const msf_cos_pair& q = pairs[i];
    MSF_REQUIRE(q.rows >= 0 && q.dim > 0 && q.dim % vec == 0, MSF_ERR_INVALID,
                "pair %d: dim=%d must be a positive multiple of %u", i, q.dim, vec);
    if (q.rows == 0) continue;  // mean over zero rows: the reference yields nan; we contribute 0 and say so in DESIGN.md
    MSF_REQUIRE(q.p && q.z && q.row_stats, MSF_ERR_INVALID, "pair %d: NULL pointer", i);
    MSF_REQUIRE(!need_grad || q.grad_p, MSF_ERR_INVALID, "pair %d: grad_p is NULL", i);
    MSF_REQUIRE(aligned16(q.p) && aligned16(q.z) && aligned16(q.row_stats) && aligned16(q.grad_p), MSF_ERR_INVALID,
                "pair %d: pointers must be 16-byte aligned", i);
    Pair& g = P.pair[P.n_pairs++];
    g.p = static_cast<const char*>(q.p);
    g.z = static_cast<const char*>(q.z);
    g.stats = reinterpret_cast<float4*>(q.row_stats);
    g.grad_p = static_cast<char*>(q.grad_p);
    g.rows = q.rows;
    g.cpr = q.dim / vec;
    // ~4 chunks of each operand per lane: few shuffle rounds per row, 8 x 128-bit loads in flight per thread
    g.lanes = 1;
    while (g.lanes < 32 && g.lanes * 4 < g.cpr) g.lanes <<= 1;
    const uint32_t groups = kThreads / g.lanes;
    const size_t row_bytes = static_cast<size_t>(g.cpr) * 16 * 2;
    uint32_t rpb = static_cast<uint32_t>((96 * 1024 + row_bytes - 1) / row_bytes);  // ~96 KB of reads per CTA
    rpb = ((rpb + groups - 1) / groups) * groups;
    g.rows_per_block = rpb;
    g.block_prefix = P.total_blocks;
    g.scale = q.coef / static_cast<float>(q.rows);
    P.total_blocks += static_cast<uint32_t>((q.rows + rpb - 1) / rpb);
  }
  return MSF_OK;
}

}  // namespace
}  // namespace msf

using namespace msf;

extern "C" size_t msf_cosine_loss_workspace_bytes(const msf_cos_pair* pairs, int n_pairs) {
  // one fp32 partial per CTA; a CTA never covers fewer than 8 rows, so rows/8+2 per pair bounds it
  size_t blocks = 0;
  for (int i = 0; pairs && i < n_pairs; ++i) blocks += static_cast<size_t>(pairs[i].rows > 0 ? pairs[i].rows / 8 + 2 : 0);
  return (blocks + 64) * sizeof(float);
}

extern "C" int msf_cosine_loss_fwd(const msf_cos_pair* pairs, int n_pairs, int dtype, float eps, float* loss_out,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  Params P{};
  if (int rc = build(P, pairs, n_pairs, dtype, false)) return rc;
  MSF_REQUIRE(loss_out, MSF_ERR_INVALID, "loss_out is NULL");
  MSF_REQUIRE(eps > 0.f, MSF_ERR_INVALID, "eps must be positive");
  P.eps = eps;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (P.total_blocks == 0) {
    MSF_CUDA_OK(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
    return MSF_OK;
  }
  MSF_REQUIRE(workspace && workspace_bytes >= P.total_blocks * sizeof(float), MSF_ERR_WORKSPACE,
              "workspace of %zu bytes < %zu required", workspace_bytes, P.total_blocks * sizeof(float));
  float* partials = static_cast<float*>(workspace);
  double bytes = 0.0;  // p and z read, 16 B of row statistics written per row
  for (int i = 0; i < n_pairs; ++i) bytes += 2.0 * pairs[i].rows * pairs[i].dim * dtype_size(dtype) + 16.0 * pairs[i].rows;
  ProfScope prof(stream, MSF_K_COS_FWD, bytes);
  MSF_DISPATCH_DTYPE(dtype, (cos_fwd_kernel<DT><<<P.total_blocks, kThreads, 0, st>>>(P, partials)));
  MSF_LAUNCH_OK("cos_fwd_kernel");
  cos_final_kernel<<<1, 256, 0, st>>>(partials, P.total_blocks, loss_out);
  MSF_LAUNCH_OK("cos_final_kernel");
  return MSF_OK;
}

extern "C" int msf_cosine_loss_bwd(const msf_cos_pair* pairs, int n_pairs, int dtype, const float* grad_out,
                                   void* stream) {
  Params P{};
  if (int rc = build(P, pairs, n_pairs, dtype, true)) return rc;
  MSF_REQUIRE(grad_out, MSF_ERR_INVALID, "grad_out is NULL");
  if (P.total_blocks == 0) return MSF_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double bytes = 0.0;  // p, z and the row statistics read, grad_p written
  for (int i = 0; i < n_pairs; ++i) bytes += 3.0 * pairs[i].rows * pairs[i].dim * dtype_size(dtype) + 16.0 * pairs[i].rows;
  ProfScope prof(stream, MSF_K_COS_BWD, bytes);
  MSF_DISPATCH_DTYPE(dtype, (cos_bwd_kernel<DT><<<P.total_blocks, kThreads, 0, st>>>(P, grad_out)));
  MSF_LAUNCH_OK("cos_bwd_kernel");
  return MSF_OK;
}
