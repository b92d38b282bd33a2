// Work decomposition + workspace layout shared by the InfoNCE kernels (host side).
#pragma once
#include "common.cuh"

namespace msf {

struct NcePlan {
  int tile_m;            // query rows per CTA
  int tile_n;            // keys per inner tile
  int64_t q_tiles;       // CTAs along the query rows
  int64_t k_tiles;       // key tiles in total
  int mode;              // 0 flash tcgen05, 1 fp32 SIMT, 2 two-pass tcgen05 GEMMs (P materialised in bf16)
  int splits;            // key-range splits of the main loop = number of O partials
  int rs_splits;         // number of row-sum partials (two-pass: one per 256-key GEMM tile column)
  int64_t tiles_per_split;
  int64_t nq_pad;        // q_tiles * tile_m
  // workspace offsets in bytes
  size_t off_rowsum;     // [splits][nq_pad] fp32   partial sum_j exp2(a*s_ij - c)
  size_t off_o;          // [splits][nq_pad][dim] fp32 partial sum_j p_ij * k_hat_j
  size_t off_pos;        // [nq] fp32 positive cosine s_ii
  size_t off_sum;        // [nq] fp32 total row sum
  size_t off_part;       // [part_cap] fp32 loss partials (one per finalize CTA)
  size_t part_cap;
  size_t off_p;          // two-pass only: P [nq][ld_p] bf16
  int64_t ld_p;
  size_t total;
};

inline bool tc_dim_ok(int dim) { return dim == 64 || dim == 128 || dim == 256; }
inline bool twopass_dim_ok(int dim) { return dim % 64 == 0 && dim > 256; }  // 512 and the fuser widths 576..4608

inline NcePlan make_nce_plan(int64_t nq, int64_t n_keys, int dim, int precision) {
  NcePlan p{};
  const bool tc = precision == MSF_BF16;
  p.mode = !tc ? 1 : (tc_dim_ok(dim) ? 0 : 2);
  p.tile_m = tc ? 128 : 64;
  p.tile_n = p.mode == 2 ? 256 : (tc ? 128 : 64);
  p.q_tiles = (nq + p.tile_m - 1) / p.tile_m;
  p.k_tiles = (n_keys + p.tile_n - 1) / p.tile_n;
  // choose the split count that minimises (waves x per-CTA tiles), the per-CTA prologue/epilogue
  // counted as ~3 tile-times; ties go to fewer splits (less partial traffic)
  int best = 1;
  double best_cost = 1e300;
  const int max_s = static_cast<int>(p.k_tiles < 32 ? (p.k_tiles > 0 ? p.k_tiles : 1) : 32);
  for (int s = 1; s <= max_s; ++s) {
    const int64_t per = (p.k_tiles + s - 1) / s;
    const int64_t ctas = p.q_tiles * s;
    const int64_t waves = (ctas + kNumSMs - 1) / kNumSMs;
    const double cost = static_cast<double>(waves) * (static_cast<double>(per) + 3.0);
    if (cost < best_cost * 0.999) {
      best_cost = cost;
      best = s;
    }
  }
  if (p.mode == 2) best = 1;  // the GEMMs own their tiling; O is written once
  p.splits = best;
  p.rs_splits = p.mode == 2 ? static_cast<int>(p.k_tiles) : best;
  p.tiles_per_split = (p.k_tiles + best - 1) / best;
  p.nq_pad = p.q_tiles * p.tile_m;
  auto align = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  size_t off = 0;
  p.off_rowsum = off;
  off = align(off + static_cast<size_t>(p.rs_splits) * p.nq_pad * sizeof(float));
  p.off_o = off;
  off = align(off + static_cast<size_t>(p.splits) * p.nq_pad * dim * sizeof(float));
  p.off_pos = off;
  off = align(off + static_cast<size_t>(nq) * sizeof(float));
  p.off_sum = off;
  off = align(off + static_cast<size_t>(nq) * sizeof(float));
  p.off_part = off;
  p.part_cap = static_cast<size_t>(nq / 8 + 2);
  off = align(off + p.part_cap * sizeof(float));
  p.off_p = off;
  p.ld_p = (n_keys + 7) & ~static_cast<int64_t>(7);
  if (p.mode == 2) off = align(off + static_cast<size_t>(nq) * p.ld_p * 2);
  p.total = off;
  return p;
}

// implemented in infonce_tc.cu: TMA + tcgen05/TMEM main loop (bf16 operands, dim in {64,128,256})
int launch_infonce_tc(const void* q_hat, const void* k_hat, int64_t nq, int64_t n_keys, int dim, float tau,
                      const NcePlan& plan, float* rowsum, float* o_part, cudaStream_t st);

// implemented in infonce_grouped.cu: the transposed flash pass of msf_infonce_dk -- rows = all keys, columns = the local
// queries, exponent a * s + col_bias[column]; planT = make_nce_plan(n_keys, nq, dim, MSF_BF16)
int launch_infonce_dk_flash(const void* k_all, const void* q_hat, int64_t n_keys, int64_t nq, int dim, float tau, const NcePlan& planT,
                            const float* col_bias, float* rowsum, float* o_part, cudaStream_t st);

// implemented in gemm_tc.cu
int launch_gemm_tc(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                   int b_mn_major, int epi, float alpha, const float* bias, float* rowsum_part, int64_t m_pad, cudaStream_t st,
                   int a_mn_major = 0);

}  // namespace msf
