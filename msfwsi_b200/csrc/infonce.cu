// L1 (infonce mode): entry points, row normalisation, fp32 SIMT main loop, finalize and backward.
// (The bf16 tensor-core main loop lives in infonce_tc.cu.)
//
// Extension named by BASELINE.json:north_star -- the reference computes only the positives
// (tools/ssl_train.py:448-466).  Anchor: the positive logit times tau IS the reference's cosine.
//
// Math (keys detached, backbone.py:188-191), a = log2(e)/tau, c = a (|q_hat.k_hat| <= 1):
//   e_ij   = exp2(a * s_ij - c)            never overflows, no running max / rescale needed
//   sum_i  = sum_j e_ij ,   O_i = sum_j e_ij k_hat_j      (one pass over the keys)
//   loss_i = ln(sum_i) + 1/tau - s_ii/tau
//   dq_hat = g/(tau) * (O_i / sum_i - k_hat_pos(i)) ,  dq = (dq_hat - q_hat (q_hat . dq_hat)) / max(||q||,eps)
#include <math.h>

#include "infonce_plan.cuh"

namespace msf {
namespace {

constexpr float kLog2e = 1.4426950408889634f;

// --------------------------------------------------------------------------------------------
// row normalisation: x -> x / max(||x||, eps), out bf16 or fp32, plus 1/max(||x||,eps)
// --------------------------------------------------------------------------------------------
template <int IDT, int ODT>
__global__ void __launch_bounds__(256) rownorm_kernel(const char* __restrict__ x, int64_t rows, uint32_t dim, float eps,
                                                      char* __restrict__ xh, float* __restrict__ inv_norm, uint32_t lanes) {
  constexpr int VI = Elem<IDT>::VEC;
  const uint32_t lane = threadIdx.x & (lanes - 1), grp = threadIdx.x / lanes, groups = 256 / lanes;
  const uint32_t cpr = dim / VI;
  const size_t in_row = static_cast<size_t>(dim) * (16 / VI), out_row = static_cast<size_t>(dim) * (ODT == MSF_F32 ? 4 : 2);
  for (int64_t base = static_cast<int64_t>(blockIdx.x) * groups; base < rows; base += static_cast<int64_t>(gridDim.x) * groups) {
    const int64_t row = base + grp;  // CTA-uniform trip count; invalid groups still join the shuffles
    const bool valid = row < rows;
    const char* src = x + row * in_row;
    float ss = 0.f;
    if (valid)
      for (uint32_t c = lane; c < cpr; c += lanes) {
        float f[VI];
        Elem<IDT>::unpack(ldg_keep(src + static_cast<size_t>(c) * 16), f);
#pragma unroll
        for (int i = 0; i < VI; ++i) ss = fmaf(f[i], f[i], ss);
      }
    ss = group_sum(ss, lanes);
    const float inv = 1.f / fmaxf(sqrtf(ss), eps);
    if (!valid) continue;
    if (lane == 0 && inv_norm) inv_norm[row] = inv;
    char* dst = xh + row * out_row;
    for (uint32_t c = lane; c < cpr; c += lanes) {  // second read hits L1/L2
      float f[VI];
      Elem<IDT>::unpack(ldg_keep(src + static_cast<size_t>(c) * 16), f);
#pragma unroll
      for (int i = 0; i < VI; ++i) f[i] *= inv;
      if constexpr (ODT == MSF_F32) {
#pragma unroll
        for (int i = 0; i < VI; i += 4) stg_stream(dst + (static_cast<size_t>(c) * VI + i) * 4, Elem<MSF_F32>::pack(f + i));
      } else {
        if constexpr (VI == 8) {
          stg_stream(dst + static_cast<size_t>(c) * 16, Elem<MSF_BF16>::pack(f));
        } else {  // 4 fp32 in -> 4 bf16 out (8 bytes)
          __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
          uint2 v = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
          *reinterpret_cast<uint2*>(dst + static_cast<size_t>(c) * 8) = v;
        }
      }
    }
  }
}

// --------------------------------------------------------------------------------------------
// fp32 SIMT main loop (exact-arithmetic path for <=1e-5 parity and for widths the tcgen05 kernel
// does not cover).  CTA = 64 query rows x one key split; key tiles of 64; D streamed in chunks of 32.
// --------------------------------------------------------------------------------------------
constexpr int SM_ = 64, SN_ = 64, SK_ = 32;

__global__ void __launch_bounds__(256) infonce_simt_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                           int64_t nq, int64_t n_keys, int dim, float a, float c,
                                                           int64_t tiles_per_split, int64_t nq_pad,
                                                           float* __restrict__ rowsum, float* __restrict__ o_part,
                                                           const float* __restrict__ col_scale) {
  __shared__ float Qs[SM_][SK_ + 1];
  __shared__ float Ks[SN_][SK_ + 1];
  __shared__ float Ps[SM_][SN_ + 1];
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const int64_t q0 = static_cast<int64_t>(blockIdx.x) * SM_;
  const int split = blockIdx.y;
  const int64_t kt0 = split * tiles_per_split;
  int64_t kt1 = kt0 + tiles_per_split;
  const int64_t k_tiles = (n_keys + SN_ - 1) / SN_;
  if (kt1 > k_tiles) kt1 = k_tiles;
  float* o_base = o_part + (static_cast<int64_t>(split) * nq_pad + q0) * dim;
  float rs[4] = {0.f, 0.f, 0.f, 0.f};
  const int orow = tid / 4, ocol = (tid % 4) * 8;  // phase-2 ownership: row, 8 columns of a 32-wide chunk

  for (int64_t kt = kt0; kt < kt1; ++kt) {
    const int64_t key0 = kt * SN_;
    float s[4][4] = {};
    for (int d0 = 0; d0 < dim; d0 += SK_) {
      __syncthreads();
      for (int i = tid; i < SM_ * SK_; i += 256) {
        const int r = i / SK_, cc = i % SK_;
        const bool okd = d0 + cc < dim;
        Qs[r][cc] = (q0 + r < nq && okd) ? q[(q0 + r) * dim + d0 + cc] : 0.f;
        Ks[r][cc] = (key0 + r < n_keys && okd) ? k[(key0 + r) * dim + d0 + cc] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < SK_; ++kk) {
        float qa[4], kb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { qa[i] = Qs[ty * 4 + i][kk]; kb[i] = Ks[tx * 4 + i][kk]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qa[i], kb[j], s[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = key0 + tx * 4 + j < n_keys;
        float e = ok ? exp2f(fmaf(s[i][j], a, -c)) : 0.f;
        if (col_scale && ok) e *= col_scale[key0 + tx * 4 + j];  // transposed pass of msf_infonce_dk: 1 / row sum of that query
        rs[i] += e;
        Ps[ty * 4 + i][tx * 4 + j] = e;
      }
    // O[64 x dim] += P[64 x 64] . K[64 x dim], chunk by chunk; this CTA owns its O rows for this split
    for (int d0 = 0; d0 < dim; d0 += SK_) {
      __syncthreads();
      for (int i = tid; i < SN_ * SK_; i += 256) {
        const int r = i / SK_, cc = i % SK_;
        Ks[r][cc] = (key0 + r < n_keys && d0 + cc < dim) ? k[(key0 + r) * dim + d0 + cc] : 0.f;
      }
      __syncthreads();
      float o[8] = {};
#pragma unroll 8
      for (int j = 0; j < SN_; ++j) {
        const float pv = Ps[orow][j];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(pv, Ks[j][ocol + e], o[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int col = d0 + ocol + e;
        if (col < dim) {
          float* dst = o_base + static_cast<int64_t>(orow) * dim + col;
          *dst = (kt == kt0 ? 0.f : *dst) + o[e];
        }
      }
    }
  }
  // row sums: combine the 16 tx-threads of each row group (16 consecutive lanes)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v = rs[i];
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (tx == 0) rowsum[static_cast<int64_t>(split) * nq_pad + q0 + ty * 4 + i] = v;
  }
  if (kt0 >= kt1) {  // empty split: contribute zeros
    for (int i = tid; i < SM_ * dim; i += 256) o_base[i] = 0.f;
  }
}

// --------------------------------------------------------------------------------------------
// forward finalize: combine split partials, positive logit, per-row loss, deterministic sum
// --------------------------------------------------------------------------------------------
template <int DT>
__device__ __forceinline__ float row_dot(const char* a, const char* b, uint32_t cpr, uint32_t lane, uint32_t lanes, bool valid) {
  constexpr int V = Elem<DT>::VEC;
  float d = 0.f;
  if (valid)
    for (uint32_t c = lane; c < cpr; c += lanes) {
      float fa[V], fb[V];
      Elem<DT>::unpack(ldg_keep(a + static_cast<size_t>(c) * 16), fa);
      Elem<DT>::unpack(ldg_keep(b + static_cast<size_t>(c) * 16), fb);
#pragma unroll
      for (int i = 0; i < V; ++i) d = fmaf(fa[i], fb[i], d);
    }
  return group_sum(d, lanes);
}

template <int DT>
__global__ void __launch_bounds__(256) nce_fwd_final_kernel(const char* __restrict__ qh, const char* __restrict__ kh,
                                                            int64_t nq, uint32_t dim, int64_t pos_offset, float inv_tau,
                                                            int splits, int64_t nq_pad, const float* __restrict__ rowsum,
                                                            float* __restrict__ pos, float* __restrict__ sum_tot,
                                                            float* __restrict__ row_lse, float* __restrict__ partials,
                                                            uint32_t lanes) {
  const uint32_t lane = threadIdx.x & (lanes - 1), grp = threadIdx.x / lanes, groups = 256 / lanes;
  const uint32_t cpr = dim / Elem<DT>::VEC;
  const size_t row_bytes = static_cast<size_t>(cpr) * 16;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * groups + grp;
  const bool valid = row < nq;
  const float s_ii = row_dot<DT>(qh + row * row_bytes, kh + (row + pos_offset) * row_bytes, cpr, lane, lanes, valid);
  float loss = 0.f;
  if (valid && lane == 0) {
    float sum = 0.f;
    for (int s = 0; s < splits; ++s) sum += rowsum[static_cast<int64_t>(s) * nq_pad + row];  // fixed order
    const float lse = logf(sum) + inv_tau;  // ln(sum_j exp(s_ij/tau)) with the fixed bound folded back
    loss = lse - s_ii * inv_tau;
    pos[row] = s_ii;
    sum_tot[row] = sum;
    if (row_lse) row_lse[row] = lse;
  }
  __shared__ float sm[8];
  float w = group_sum(loss, 32);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i];
    partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) final_sum_kernel(const float* __restrict__ partials, uint32_t n, float* out) {
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

// --------------------------------------------------------------------------------------------
// backward: dq from the saved O partials (the keys are NOT re-read except the positive row)
// --------------------------------------------------------------------------------------------
template <int DT, int GDT>
__global__ void __launch_bounds__(256) nce_bwd_kernel(const char* __restrict__ qh, const char* __restrict__ kh,
                                                      const float* __restrict__ inv_norm, int64_t nq, uint32_t dim,
                                                      int64_t pos_offset, float inv_tau, int splits, int64_t nq_pad,
                                                      const float* __restrict__ o_part, const float* __restrict__ sum_tot,
                                                      const float* __restrict__ grad_out, float scale,
                                                      char* __restrict__ grad_q, uint32_t lanes) {
  constexpr int V = Elem<DT>::VEC;
  const uint32_t lane = threadIdx.x & (lanes - 1), grp = threadIdx.x / lanes, groups = 256 / lanes;
  const uint32_t cpr = dim / V;
  const size_t row_bytes = static_cast<size_t>(cpr) * 16;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * groups + grp;
  const bool valid = row < nq;
  const float gs = __ldg(grad_out) * scale * inv_tau;
  const float inv_sum = valid ? 1.f / sum_tot[row] : 0.f;
  const char* qrow = qh + row * row_bytes;
  const char* krow = kh + (row + pos_offset) * row_bytes;
  // d[i] = gs * (O_i / sum - k_pos,i): the O partials of the key splits are summed in fixed order, read as 128-bit vectors
  auto load_d = [&](uint32_t c, float* fq, float* d) {
    float fk[V];
    Elem<DT>::unpack(ldg_keep(qrow + static_cast<size_t>(c) * 16), fq);
    Elem<DT>::unpack(ldg_keep(krow + static_cast<size_t>(c) * 16), fk);
    float o[V];
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] = 0.f;
    for (int s = 0; s < splits; ++s) {
      const float4* src = reinterpret_cast<const float4*>(o_part + (static_cast<int64_t>(s) * nq_pad + row) * dim + c * V);
#pragma unroll
      for (int i = 0; i < V; i += 4) {
        const float4 v = __ldg(src + i / 4);
        o[i] += v.x; o[i + 1] += v.y; o[i + 2] += v.z; o[i + 3] += v.w;
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) d[i] = gs * (o[i] * inv_sum - fk[i]);
  };
  auto store_g = [&](uint32_t c, const float* g) {
    constexpr int GV = Elem<GDT>::VEC;
    char* dst = grad_q + (row * dim + static_cast<int64_t>(c) * V) * (16 / GV);
    if constexpr (GV == V) {
      stg_stream(dst, Elem<GDT>::pack(g));
    } else if constexpr (GV < V) {  // fp32 gradient from bf16 operands: two chunks
      stg_stream(dst, Elem<GDT>::pack(g));
      stg_stream(dst + 16, Elem<GDT>::pack(g + 4));
    } else {  // 16-bit gradient from fp32 operands: half a chunk (8 bytes)
      float tmp[8] = {g[0], g[1], g[2], g[3], 0.f, 0.f, 0.f, 0.f};
      const uint4 pk = Elem<GDT>::pack(tmp);
      *reinterpret_cast<uint2*>(dst) = make_uint2(pk.x, pk.y);
    }
  };
  const float inn = (valid && inv_norm) ? inv_norm[row] : 1.f;
  constexpr int kKeep = 4;  // chunks per lane kept in registers between the two passes (every flash width)
  if (cpr <= lanes * kKeep) {
    float fq[kKeep][V], d[kKeep][V];
    float t = 0.f;  // q_hat . d
#pragma unroll
    for (int j = 0; j < kKeep; ++j) {
      const uint32_t c = lane + j * lanes;
      if (valid && c < cpr) {
        load_d(c, fq[j], d[j]);
#pragma unroll
        for (int i = 0; i < V; ++i) t = fmaf(fq[j][i], d[j][i], t);
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
        for (int i = 0; i < V; ++i) g[i] = (d[j][i] - fq[j][i] * t) * inn;
        store_g(c, g);
      }
    }
    return;
  }
  // wide rows: two passes over the row, the second one served by L2
  float t = 0.f;
  if (valid)
    for (uint32_t c = lane; c < cpr; c += lanes) {
      float fq[V], d[V];
      load_d(c, fq, d);
#pragma unroll
      for (int i = 0; i < V; ++i) t = fmaf(fq[i], d[i], t);
    }
  t = group_sum(t, lanes);
  if (!valid) return;
  for (uint32_t c = lane; c < cpr; c += lanes) {
    float fq[V], d[V], g[V];
    load_d(c, fq, d);
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = (d[i] - fq[i] * t) * inn;
    store_g(c, g);
  }
}

// --------------------------------------------------------------------------------------------
// key gradient (keys NOT detached -- north_star (4); the reference detaches every key, backbone.py:188-191)
//   dk_hat_j = g/tau * ( sum_i softmax_ij q_hat_i  -  q_hat_{i : pos(i) = j} )
// The sum over the LOCAL queries is the same flash pass with the roles swapped (rows = all keys, columns = local queries);
// softmax_ij = e_ij / sum_i turns the per-query 1/sum_i into a per-COLUMN term of that pass.  The partial over this rank's
// queries covers every global key; the caller reduce-scatters it to the keys' owners (NCCL), then msf_infonce_dk_finish
// subtracts the positives (local by construction) and applies the normalise Jacobian of z.
// --------------------------------------------------------------------------------------------
// mode 0: exponent bias -(a + log2 sum_i) for the tcgen05 pass (padding -1e30 -> exp2 = 0); mode 1: 1 / sum_i (fp32 SIMT pass)
__global__ void __launch_bounds__(256) nce_col_terms_kernel(const float* __restrict__ sum_tot, int64_t nq, int64_t n_pad, float a, int mode,
                                                            float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n_pad) return;
  if (mode == 0) out[i] = i < nq ? -(a + log2f(sum_tot[i])) : -1e30f;
  else out[i] = i < nq ? 1.f / sum_tot[i] : 0.f;
}

// two-pass widths: Qs_i = q_hat_i / sum_i in bf16 (the B operand of dK = P^T Qs)
__global__ void __launch_bounds__(256) nce_scale_rows_kernel(const char* __restrict__ qh, const float* __restrict__ sum_tot, int64_t nq,
                                                             uint32_t cpr, char* __restrict__ out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= nq * cpr) return;
  const int64_t row = idx / cpr;
  float f[8];
  Elem<MSF_BF16>::unpack(ldg_keep(qh + idx * 16), f);
  const float inv = 1.f / sum_tot[row];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] *= inv;
  *reinterpret_cast<uint4*>(out + idx * 16) = Elem<MSF_BF16>::pack(f);
}

// dk_out[j] = g * scale / tau * sum_splits O^T partials (fixed order)
__global__ void __launch_bounds__(256) nce_dk_combine_kernel(const float* __restrict__ o_part, int splits, int64_t n_pad, int64_t n_keys,
                                                             uint32_t dim4, const float* __restrict__ grad_out, float scale_inv_tau,
                                                             float* __restrict__ dk_out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= n_keys * dim4) return;
  const float gs = __ldg(grad_out) * scale_inv_tau;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(o_part + (static_cast<int64_t>(s) * n_pad * dim4 + idx) * 4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(dk_out + idx * 4) = make_float4(acc.x * gs, acc.y * gs, acc.z * gs, acc.w * gs);
}

// grad_z_j = J_normalize(z_j)^T (dk_j - [j < nq] g scale/tau q_hat_j);  J^T d = (d - k_hat (k_hat . d)) / max(||z||, eps)
template <int DT, int GDT>
__global__ void __launch_bounds__(256) nce_dk_finish_kernel(const float* __restrict__ dk, const char* __restrict__ qh, const char* __restrict__ kh,
                                                            const float* __restrict__ inv_norm, int64_t rows, int64_t nq, uint32_t dim,
                                                            const float* __restrict__ grad_out, float scale_inv_tau, char* __restrict__ grad_z,
                                                            uint32_t lanes) {
  constexpr int V = Elem<DT>::VEC;
  const uint32_t lane = threadIdx.x & (lanes - 1), grp = threadIdx.x / lanes, groups = 256 / lanes;
  const uint32_t cpr = dim / V;
  const size_t row_bytes = static_cast<size_t>(cpr) * 16;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * groups + grp;
  const bool valid = row < rows;
  const float gs = (valid && row < nq) ? __ldg(grad_out) * scale_inv_tau : 0.f;
  auto load_d = [&](uint32_t c, float* fk, float* d) {
    float fq[V];
    Elem<DT>::unpack(ldg_keep(kh + row * row_bytes + static_cast<size_t>(c) * 16), fk);
    if (row < nq) {
      Elem<DT>::unpack(ldg_keep(qh + row * row_bytes + static_cast<size_t>(c) * 16), fq);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) fq[i] = 0.f;
    }
    const float* src = dk + row * dim + static_cast<int64_t>(c) * V;
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      const float4 v = *reinterpret_cast<const float4*>(src + i);
      d[i] = v.x - gs * fq[i]; d[i + 1] = v.y - gs * fq[i + 1]; d[i + 2] = v.z - gs * fq[i + 2]; d[i + 3] = v.w - gs * fq[i + 3];
    }
  };
  float t = 0.f;
  if (valid)
    for (uint32_t c = lane; c < cpr; c += lanes) {
      float fk[V], d[V];
      load_d(c, fk, d);
#pragma unroll
      for (int i = 0; i < V; ++i) t = fmaf(fk[i], d[i], t);
    }
  t = group_sum(t, lanes);
  if (!valid) return;
  const float inn = inv_norm ? inv_norm[row] : 1.f;
  for (uint32_t c = lane; c < cpr; c += lanes) {
    float fk[V], d[V], g[V];
    load_d(c, fk, d);
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = (d[i] - fk[i] * t) * inn;
    constexpr int GV = Elem<GDT>::VEC;
    char* dst = grad_z + (row * dim + static_cast<int64_t>(c) * V) * (16 / GV);
    if constexpr (GV == V) {
      stg_stream(dst, Elem<GDT>::pack(g));
    } else if constexpr (GV < V) {
      stg_stream(dst, Elem<GDT>::pack(g));
      stg_stream(dst + 16, Elem<GDT>::pack(g + 4));
    } else {
      float tmp[8] = {g[0], g[1], g[2], g[3], 0.f, 0.f, 0.f, 0.f};
      const uint4 pk = Elem<GDT>::pack(tmp);
      *reinterpret_cast<uint2*>(dst) = make_uint2(pk.x, pk.y);
    }
  }
}

// workspace of msf_infonce_dk: [column terms | the transposed pass's own plan (flash / SIMT) or {Qs, O^T} (two-pass)]
struct DkPlan {
  int mode;
  NcePlan t;        // transposed problem: "queries" = all keys, "keys" = local queries (modes 0 / 1)
  int64_t col_pad;  // floats of column terms
  size_t off_col, off_t, off_qs, off_o, total;
};
inline DkPlan make_dk_plan(int64_t nq, int64_t n_keys, int dim, int precision) {
  DkPlan d{};
  auto align = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  d.t = make_nce_plan(n_keys, nq, dim, precision);
  d.mode = d.t.mode;
  d.col_pad = (nq + 127) / 128 * 128;
  size_t off = 0;
  d.off_col = off;
  off = align(off + static_cast<size_t>(d.col_pad) * sizeof(float));
  d.off_t = d.off_qs = d.off_o = off;
  if (d.mode == 2) {
    off = align(off + static_cast<size_t>(nq) * dim * 2);
    d.off_o = off;
    off = align(off + static_cast<size_t>(n_keys) * dim * sizeof(float));
  } else {
    off += d.t.total;
  }
  d.total = off;
  return d;
}

inline uint32_t lanes_for(uint32_t cpr) {  // ~4 chunks per lane: few shuffle rounds per row, several loads in flight
  uint32_t lanes = 1;
  while (lanes < 32 && lanes * 4 < cpr) lanes <<= 1;
  return lanes;
}

int check_nce(const void* qh, const void* kh, int64_t nq, int64_t n_keys, int dim, int64_t pos_offset, float tau,
              int precision) {
  MSF_REQUIRE(precision == MSF_F32 || precision == MSF_BF16, MSF_ERR_INVALID, "precision must be MSF_F32 or MSF_BF16");
  MSF_REQUIRE(nq > 0 && n_keys > 0 && dim > 0, MSF_ERR_INVALID, "empty problem nq=%lld n_keys=%lld dim=%d",
              static_cast<long long>(nq), static_cast<long long>(n_keys), dim);
  MSF_REQUIRE(qh && kh && aligned16(qh) && aligned16(kh), MSF_ERR_INVALID, "q_hat/k_hat NULL or not 16-byte aligned");
  MSF_REQUIRE(pos_offset >= 0 && pos_offset + nq <= n_keys, MSF_ERR_INVALID,
              "positives [%lld, %lld) fall outside the %lld keys", static_cast<long long>(pos_offset),
              static_cast<long long>(pos_offset + nq), static_cast<long long>(n_keys));
  MSF_REQUIRE(tau > 0.f && 2.f * kLog2e / tau <= 120.f, MSF_ERR_UNSUPPORTED,
              "tau=%g outside the supported range (tau >= 0.0241): fixed-bound softmax would underflow fp32", tau);
  if (precision == MSF_BF16)
    MSF_REQUIRE(tc_dim_ok(dim) || twopass_dim_ok(dim), MSF_ERR_UNSUPPORTED,
                "tcgen05 paths cover dim in {64,128,256} (flash) and multiples of 64 above 256 (two-pass); got %d (use MSF_F32)", dim);
  else
    MSF_REQUIRE(dim % 4 == 0, MSF_ERR_INVALID, "dim must be a multiple of 4");
  return MSF_OK;
}

}  // namespace
}  // namespace msf

using namespace msf;

extern "C" int msf_rownorm(const void* x, int64_t rows, int dim, int in_dtype, float eps, void* x_hat, int out_dtype,
                           float* inv_norm, void* stream) {
  MSF_REQUIRE(dtype_ok(in_dtype) && (out_dtype == MSF_F32 || out_dtype == MSF_BF16), MSF_ERR_INVALID, "bad dtype");
  MSF_REQUIRE(rows >= 0 && dim > 0 && dim % 8 == 0, MSF_ERR_INVALID, "dim=%d must be a positive multiple of 8", dim);
  if (rows == 0) return MSF_OK;
  MSF_REQUIRE(x && x_hat && aligned16(x) && aligned16(x_hat), MSF_ERR_INVALID, "NULL or misaligned pointer");
  MSF_REQUIRE(eps > 0.f, MSF_ERR_INVALID, "eps must be positive");
  const uint32_t cpr = dim / (16 / dtype_size(in_dtype));
  const uint32_t lanes = lanes_for(cpr), groups = 256 / lanes;
  int64_t blocks = (rows + groups - 1) / groups;
  const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const char* xi = static_cast<const char*>(x);
  char* xo = static_cast<char*>(x_hat);
  ProfScope prof(stream, MSF_K_ROWNORM, static_cast<double>(rows) * dim * (dtype_size(in_dtype) + dtype_size(out_dtype)) + 4.0 * rows);
#define MSF_RN(I, O)                                                                                                 \
  if (in_dtype == I && out_dtype == O) {                                                                             \
    rownorm_kernel<I, O><<<static_cast<unsigned>(blocks), 256, 0, st>>>(xi, rows, dim, eps, xo, inv_norm, lanes);    \
    MSF_LAUNCH_OK("rownorm_kernel");                                                                                 \
    return MSF_OK;                                                                                                   \
  }
  MSF_RN(MSF_F32, MSF_F32) MSF_RN(MSF_F32, MSF_BF16) MSF_RN(MSF_BF16, MSF_F32) MSF_RN(MSF_BF16, MSF_BF16)
  MSF_RN(MSF_F16, MSF_F32) MSF_RN(MSF_F16, MSF_BF16)
#undef MSF_RN
  return MSF_ERR_UNSUPPORTED;
}

extern "C" size_t msf_infonce_workspace_bytes(int64_t nq, int64_t n_keys, int dim, int precision) {
  if (nq <= 0 || n_keys <= 0 || dim <= 0) return 0;
  return make_nce_plan(nq, n_keys, dim, precision).total;
}

extern "C" int msf_infonce_plan_info(int64_t nq, int64_t n_keys, int dim, int precision, int64_t* info) {
  MSF_REQUIRE(info && nq > 0 && n_keys > 0 && dim > 0, MSF_ERR_INVALID, "bad arguments");
  const NcePlan p = make_nce_plan(nq, n_keys, dim, precision);
  info[0] = p.splits; info[1] = p.nq_pad; info[2] = static_cast<int64_t>(p.off_rowsum); info[3] = static_cast<int64_t>(p.off_o);
  info[4] = static_cast<int64_t>(p.off_pos); info[5] = static_cast<int64_t>(p.off_sum); info[6] = p.tile_m; info[7] = p.tile_n;
  return MSF_OK;
}

extern "C" int msf_infonce_fwd(const void* q_hat, const void* k_hat, int64_t nq, int64_t n_keys, int dim,
                               int64_t pos_offset, float tau, int precision, float* loss_sum_out, float* row_lse,
                               void* workspace, size_t workspace_bytes, void* stream) {
  return msf_infonce_fwd_timed(q_hat, k_hat, nq, n_keys, dim, pos_offset, tau, precision, loss_sum_out, row_lse, workspace,
                               workspace_bytes, stream, nullptr, nullptr);
}

extern "C" int msf_infonce_fwd_timed(const void* q_hat, const void* k_hat, int64_t nq, int64_t n_keys, int dim,
                                     int64_t pos_offset, float tau, int precision, float* loss_sum_out, float* row_lse,
                                     void* workspace, size_t workspace_bytes, void* stream, void* ev_main_start,
                                     void* ev_main_stop) {
  if (int rc = check_nce(q_hat, k_hat, nq, n_keys, dim, pos_offset, tau, precision)) return rc;
  MSF_REQUIRE(loss_sum_out, MSF_ERR_INVALID, "loss_sum_out is NULL");
  const NcePlan plan = make_nce_plan(nq, n_keys, dim, precision);
  MSF_REQUIRE(workspace && aligned16(workspace) && workspace_bytes >= plan.total, MSF_ERR_WORKSPACE,
              "workspace of %zu bytes < %zu required", workspace_bytes, plan.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  float* rowsum = reinterpret_cast<float*>(ws + plan.off_rowsum);
  float* o_part = reinterpret_cast<float*>(ws + plan.off_o);
  float* pos = reinterpret_cast<float*>(ws + plan.off_pos);
  float* sum_tot = reinterpret_cast<float*>(ws + plan.off_sum);
  float* partials = reinterpret_cast<float*>(ws + plan.off_part);
  const float a = kLog2e / tau;
  if (ev_main_start) MSF_CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(ev_main_start), st));
  {
  ProfScope prof(stream, plan.mode == 0 ? MSF_K_NCE_FLASH : (plan.mode == 2 ? MSF_K_NCE_TWOPASS : MSF_K_NCE_SIMT),
                 4.0 * static_cast<double>(nq) * static_cast<double>(n_keys) * dim);  // S = QK^T and O = PK
  if (plan.mode == 0) {
    if (int rc = launch_infonce_tc(q_hat, k_hat, nq, n_keys, dim, tau, plan, rowsum, o_part, st)) return rc;
  } else if (plan.mode == 2) {
    // pass 1: P = exp2(a * Q K^T - a) in bf16 (+ row-sum partials per 256-key tile column); pass 2: O = P K
    void* P = ws + plan.off_p;
    if (int rc = launch_gemm_tc(q_hat, dim, k_hat, dim, P, plan.ld_p, nq, n_keys, dim, 0, 2 /*EPI_EXP*/, a, nullptr, rowsum,
                                plan.nq_pad, st)) return rc;
    if (int rc = launch_gemm_tc(P, plan.ld_p, k_hat, dim, o_part, dim, nq, dim, n_keys, 1, 0 /*EPI_F32*/, 1.f, nullptr, nullptr, 0, st))
      return rc;
  } else {
    dim3 grid(static_cast<unsigned>(plan.q_tiles), static_cast<unsigned>(plan.splits));
    infonce_simt_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(q_hat), static_cast<const float*>(k_hat), nq,
                                              n_keys, dim, a, a, plan.tiles_per_split, plan.nq_pad, rowsum, o_part, nullptr);
    MSF_LAUNCH_OK("infonce_simt_kernel");
  }
  }
  if (ev_main_stop) MSF_CUDA_OK(cudaEventRecord(static_cast<cudaEvent_t>(ev_main_stop), st));
  const uint32_t cpr = dim / (precision == MSF_BF16 ? 8 : 4);
  const uint32_t lanes = lanes_for(cpr), groups = 256 / lanes;
  const unsigned blocks = static_cast<unsigned>((nq + groups - 1) / groups);
  MSF_REQUIRE(blocks <= plan.part_cap, MSF_ERR_WORKSPACE, "internal: loss partial capacity");
  const char* qh = static_cast<const char*>(q_hat);
  const char* kh = static_cast<const char*>(k_hat);
  if (precision == MSF_BF16)
    nce_fwd_final_kernel<MSF_BF16><<<blocks, 256, 0, st>>>(qh, kh, nq, dim, pos_offset, 1.f / tau, plan.rs_splits, plan.nq_pad,
                                                           rowsum, pos, sum_tot, row_lse, partials, lanes);
  else
    nce_fwd_final_kernel<MSF_F32><<<blocks, 256, 0, st>>>(qh, kh, nq, dim, pos_offset, 1.f / tau, plan.rs_splits, plan.nq_pad,
                                                          rowsum, pos, sum_tot, row_lse, partials, lanes);
  MSF_LAUNCH_OK("nce_fwd_final_kernel");
  final_sum_kernel<<<1, 256, 0, st>>>(partials, blocks, loss_sum_out);
  MSF_LAUNCH_OK("final_sum_kernel");
  return MSF_OK;
}

extern "C" int msf_infonce_bwd(const void* q_hat, const void* k_hat, const float* q_inv_norm, int64_t nq, int64_t n_keys,
                               int dim, int64_t pos_offset, float tau, int precision, const float* grad_out, float scale,
                               const void* workspace, size_t workspace_bytes, void* grad_q, int grad_dtype, void* stream) {
  if (int rc = check_nce(q_hat, k_hat, nq, n_keys, dim, pos_offset, tau, precision)) return rc;
  MSF_REQUIRE(grad_out && grad_q && aligned16(grad_q) && dtype_ok(grad_dtype), MSF_ERR_INVALID, "bad gradient arguments");
  const NcePlan plan = make_nce_plan(nq, n_keys, dim, precision);
  MSF_REQUIRE(workspace && workspace_bytes >= plan.total, MSF_ERR_WORKSPACE, "workspace of %zu bytes < %zu required",
              workspace_bytes, plan.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const char* ws = static_cast<const char*>(workspace);
  const float* o_part = reinterpret_cast<const float*>(ws + plan.off_o);
  const float* sum_tot = reinterpret_cast<const float*>(ws + plan.off_sum);
  const uint32_t cpr = dim / (precision == MSF_BF16 ? 8 : 4);
  const uint32_t lanes = lanes_for(cpr), groups = 256 / lanes;
  const unsigned blocks = static_cast<unsigned>((nq + groups - 1) / groups);
  const char* qh = static_cast<const char*>(q_hat);
  const char* kh = static_cast<const char*>(k_hat);
  char* gq = static_cast<char*>(grad_q);
  // O partials + q_hat + the positive key rows read, grad_q written
  ProfScope prof(stream, MSF_K_NCE_BWD, static_cast<double>(nq) * dim * (4.0 * plan.splits + 2.0 * (precision == MSF_BF16 ? 2 : 4) + dtype_size(grad_dtype)));
#define MSF_NB(P, G)                                                                                              \
  if (precision == P && grad_dtype == G) {                                                                        \
    nce_bwd_kernel<P, G><<<blocks, 256, 0, st>>>(qh, kh, q_inv_norm, nq, dim, pos_offset, 1.f / tau, plan.splits, \
                                                 plan.nq_pad, o_part, sum_tot, grad_out, scale, gq, lanes);       \
    MSF_LAUNCH_OK("nce_bwd_kernel");                                                                              \
    return MSF_OK;                                                                                                \
  }
  MSF_NB(MSF_F32, MSF_F32) MSF_NB(MSF_F32, MSF_BF16) MSF_NB(MSF_F32, MSF_F16)
  MSF_NB(MSF_BF16, MSF_F32) MSF_NB(MSF_BF16, MSF_BF16) MSF_NB(MSF_BF16, MSF_F16)
#undef MSF_NB
  return MSF_ERR_UNSUPPORTED;
}


extern "C" size_t msf_infonce_dk_workspace_bytes(int64_t nq, int64_t n_keys, int dim, int precision) {
  if (nq <= 0 || n_keys <= 0 || dim <= 0) return 0;
  return make_dk_plan(nq, n_keys, dim, precision).total;
}

extern "C" int msf_infonce_dk(const void* q_hat, const void* k_hat, int64_t nq, int64_t n_keys, int dim, int64_t pos_offset, float tau,
                              int precision, const float* grad_out, float scale, const void* fwd_workspace, size_t fwd_workspace_bytes,
                              float* dk_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_nce(q_hat, k_hat, nq, n_keys, dim, pos_offset, tau, precision)) return rc;
  MSF_REQUIRE(grad_out && dk_out && aligned16(dk_out), MSF_ERR_INVALID, "grad_out / dk_out NULL or misaligned");
  const NcePlan fwd = make_nce_plan(nq, n_keys, dim, precision);
  MSF_REQUIRE(fwd_workspace && fwd_workspace_bytes >= fwd.total, MSF_ERR_WORKSPACE, "forward workspace of %zu bytes < %zu required",
              fwd_workspace_bytes, fwd.total);
  const DkPlan dp = make_dk_plan(nq, n_keys, dim, precision);
  MSF_REQUIRE(workspace && aligned16(workspace) && workspace_bytes >= dp.total, MSF_ERR_WORKSPACE, "workspace of %zu bytes < %zu required",
              workspace_bytes, dp.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const char* fws = static_cast<const char*>(fwd_workspace);
  const float* sum_tot = reinterpret_cast<const float*>(fws + fwd.off_sum);
  char* ws = static_cast<char*>(workspace);
  float* col = reinterpret_cast<float*>(ws + dp.off_col);
  const float a = kLog2e / tau;
  const float* o_part = nullptr;
  int splits = 1;
  int64_t n_pad = n_keys;
  ProfScope prof(stream, MSF_K_NCE_DK, 4.0 * static_cast<double>(nq) * static_cast<double>(n_keys) * dim);
  if (dp.mode == 2) {
    // dK = P^T Qs with the forward's 16-bit P = exp2(a s - a) still in its workspace and Qs_i = q_hat_i / sum_i
    char* qs = ws + dp.off_qs;
    const uint32_t cpr = dim / 8;
    nce_scale_rows_kernel<<<static_cast<unsigned>((nq * cpr + 255) / 256), 256, 0, st>>>(static_cast<const char*>(q_hat), sum_tot, nq, cpr, qs);
    MSF_LAUNCH_OK("nce_scale_rows_kernel");
    float* o = reinterpret_cast<float*>(ws + dp.off_o);
    if (int rc = launch_gemm_tc(fws + fwd.off_p, fwd.ld_p, qs, dim, o, dim, n_keys, dim, nq, 1, 0 /*EPI_F32*/, 1.f, nullptr, nullptr, 0, st, 1))
      return rc;
    o_part = o;
  } else {
    nce_col_terms_kernel<<<static_cast<unsigned>((dp.col_pad + 255) / 256), 256, 0, st>>>(sum_tot, nq, dp.col_pad, a, dp.mode == 0 ? 0 : 1, col);
    MSF_LAUNCH_OK("nce_col_terms_kernel");
    char* tws = ws + dp.off_t;
    float* rowsum = reinterpret_cast<float*>(tws + dp.t.off_rowsum);
    float* o = reinterpret_cast<float*>(tws + dp.t.off_o);
    if (dp.mode == 0) {
      if (int rc = launch_infonce_dk_flash(k_hat, q_hat, n_keys, nq, dim, tau, dp.t, col, rowsum, o, st)) return rc;
    } else {
      dim3 grid(static_cast<unsigned>(dp.t.q_tiles), static_cast<unsigned>(dp.t.splits));
      infonce_simt_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(k_hat), static_cast<const float*>(q_hat), n_keys, nq, dim, a, a,
                                                dp.t.tiles_per_split, dp.t.nq_pad, rowsum, o, col);
      MSF_LAUNCH_OK("infonce_simt_kernel");
    }
    o_part = o;
    splits = dp.t.splits;
    n_pad = dp.t.nq_pad;
  }
  const uint32_t dim4 = dim / 4;
  nce_dk_combine_kernel<<<static_cast<unsigned>((n_keys * dim4 + 255) / 256), 256, 0, st>>>(o_part, splits, n_pad, n_keys, dim4, grad_out, scale / tau,
                                                                                          dk_out);
  MSF_LAUNCH_OK("nce_dk_combine_kernel");
  return MSF_OK;
}

extern "C" int msf_infonce_dk_finish(const float* dk_local, const void* q_hat, const void* k_hat_local, const float* k_inv_norm, int64_t rows,
                                     int64_t nq, int dim, float tau, int precision, const float* grad_out, float scale, void* grad_z,
                                     int grad_dtype, void* stream) {
  MSF_REQUIRE(precision == MSF_F32 || precision == MSF_BF16, MSF_ERR_INVALID, "precision must be MSF_F32 or MSF_BF16");
  MSF_REQUIRE(rows > 0 && nq >= 0 && nq <= rows && dim > 0 && dim % (precision == MSF_BF16 ? 8 : 4) == 0, MSF_ERR_INVALID,
              "bad shape rows=%lld nq=%lld dim=%d", static_cast<long long>(rows), static_cast<long long>(nq), dim);
  MSF_REQUIRE(dk_local && q_hat && k_hat_local && grad_out && grad_z && aligned16(dk_local) && aligned16(q_hat) && aligned16(k_hat_local) &&
                  aligned16(grad_z) && dtype_ok(grad_dtype) && tau > 0.f,
              MSF_ERR_INVALID, "NULL / misaligned pointer or bad dtype");
  const uint32_t cpr = dim / (precision == MSF_BF16 ? 8 : 4);
  const uint32_t lanes = lanes_for(cpr), groups = 256 / lanes;
  const unsigned blocks = static_cast<unsigned>((rows + groups - 1) / groups);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const char* qh = static_cast<const char*>(q_hat);
  const char* kh = static_cast<const char*>(k_hat_local);
  char* gz = static_cast<char*>(grad_z);
#define MSF_NF(P, G)                                                                                                                       \
  if (precision == P && grad_dtype == G) {                                                                                                 \
    nce_dk_finish_kernel<P, G><<<blocks, 256, 0, st>>>(dk_local, qh, kh, k_inv_norm, rows, nq, dim, grad_out, scale / tau, gz, lanes);     \
    MSF_LAUNCH_OK("nce_dk_finish_kernel");                                                                                                 \
    return MSF_OK;                                                                                                                         \
  }
  MSF_NF(MSF_F32, MSF_F32) MSF_NF(MSF_F32, MSF_BF16) MSF_NF(MSF_F32, MSF_F16)
  MSF_NF(MSF_BF16, MSF_F32) MSF_NF(MSF_BF16, MSF_BF16) MSF_NF(MSF_BF16, MSF_F16)
#undef MSF_NF
  return MSF_ERR_UNSUPPORTED;
}
