// G1: persistent bf16 tcgen05 GEMM with fused epilogues, C[M,N] = A[M,K] * op(B).
//
// Used for (a) the head Linears of src/models/backbone.py:12-31 (y = x W^T, no bias / + bias), and (b) the
// two-pass InfoNCE path for widths the flash kernel cannot hold in TMEM (512 and the fuser widths 576..4608):
//   pass 1  P = exp2(a * Q_hat K_hat^T - a)  -> bf16 + per-row partial sums      (EPI_EXP, B K-major)
//   pass 2  O = P K_hat                        -> fp32                            (EPI_F32, B MN-major)
//
// CTA tile 128 x 256, BK = 64, 4-stage TMA/mbarrier ring (48 KB per stage), accumulators double-buffered in
// TMEM (2 x 256 columns) so the epilogue of tile i overlaps the main loop of tile i+1; persistent CTAs walk
// the tile list m-fastest (one B panel is shared by a whole wave through L2).
// Warp roles (256 threads): w0 TMA producer, w1 MMA issuer, w2 TMEM allocator, w4-7 epilogue (one row/thread).
// Tensor-bound: 2*M*N*K FLOP; bytes (M*K + N*K)*2 read (+ re-reads served by L2) + M*N*e written.
#include "tc_common.cuh"

namespace msf {
namespace {

using namespace tc;
constexpr int GBN = 256, GBK = 64, kStages = 4, kThreads = 256;
constexpr uint32_t kABytes = BM * GBK * 2, kBBytes = GBN * GBK * 2, kStageBytes = kABytes + kBBytes;
constexpr uint32_t kSmem = 1024 + kStages * kStageBytes + 1024;

enum { EPI_F32 = 0, EPI_BF16 = 1, EPI_EXP = 2 };

struct GemmArgs {
  int64_t M, N, K;
  int64_t ldc;          // elements
  void* C;
  const float* bias;    // per output column or null (EPI_F32 / EPI_BF16)
  float alpha;          // plain: scale; EPI_EXP: a = log2(e)/tau
  float* rowsum_part;   // EPI_EXP: [n_tiles][m_pad]
  int64_t m_pad;
  int b_mn_major;       // 0: B is [N,K] row-major (C = A B^T); 1: B is [K,N] row-major (C = A B)
  int a_mn_major;       // 0: A is [M,K] row-major; 1: A is stored transposed, [K,M] row-major (C = A^T-stored x B)
};

template <int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* acc_full = empty + kStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tiles_m = (g.M + BM - 1) / BM, tiles_n = (g.N + GBN - 1) / GBN;
  const int64_t n_tiles = tiles_m * tiles_n;
  const int num_kb = static_cast<int>((g.K + GBK - 1) / GBK);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full + b, 1); mbar_init(acc_empty + b, 128); }
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
    if (lane == 0) {
      tma_prefetch_desc(&tm_a);
      tma_prefetch_desc(&tm_b);
      uint32_t kb_total = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = static_cast<int>((tile % tiles_m) * BM), n0 = static_cast<int>((tile / tiles_m) * GBN);
        for (int kb = 0; kb < num_kb; ++kb, ++kb_total) {
          const int s = kb_total % kStages;
          mbar_wait(empty + s, ((kb_total / kStages) & 1) ^ 1);
          mbar_expect_tx(full + s, kStageBytes);
          uint8_t* sa = smem + s * kStageBytes;
          uint8_t* sb = sa + kABytes;
          if (!g.a_mn_major) {
            tma_load_2d(sa, &tm_a, kb * GBK, m0, full + s);  // one 64(k) x 128(m) box: rows = m, 128 B of k per row
          } else {
            for (int j = 0; j < BM / 64; ++j)                // two 64(m) x 64(k) boxes: rows = k, 128 B of m per row
              tma_load_2d(sa + j * (GBK * 128), &tm_a, m0 + j * 64, kb * GBK, full + s);
          }
          if (!g.b_mn_major) {
            tma_load_2d(sb, &tm_b, kb * GBK, n0, full + s);  // one 64 x 256 box: rows = n, 128 B of k per row
          } else {
            for (int j = 0; j < GBN / 64; ++j)               // four 64(n) x 64(k) boxes: rows = k, 128 B of n per row
              tma_load_2d(sb + j * (GBK * 128), &tm_b, n0 + j * 64, kb * GBK, full + s);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(GBN, g.b_mn_major != 0, g.a_mn_major != 0);
      uint32_t kb_total = 0, it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(acc_empty + buf, ((it >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem + buf * GBN;
        for (int kb = 0; kb < num_kb; ++kb, ++kb_total) {
          const int s = kb_total % kStages;
          mbar_wait(full + s, (kb_total / kStages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * kStageBytes), b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < GBK / 16; ++k) {
            const uint64_t ad = g.a_mn_major ? umma_desc(a_addr + k * 2048, GBK * 128, 1024) : umma_desc(a_addr + k * 32, 16, 1024);
            const uint64_t bd = g.b_mn_major ? umma_desc(b_addr + k * 2048, GBK * 128, 1024) : umma_desc(b_addr + k * 32, 16, 1024);
            mma_ss(d_tmem, ad, bd, idesc, (kb > 0 || k > 0));
          }
          tc_commit(empty + s);
        }
        tc_commit(acc_full + buf);
      }
    }
  } else if (warp >= 4) {
    const int row_in_tile = ((warp & 3) << 5) + lane;
    const uint32_t lane_base = tmem + (static_cast<uint32_t>((warp & 3) << 5) << 16);
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1;
      const int64_t tm = tile % tiles_m, tn = tile / tiles_m;
      const int64_t row = tm * BM + row_in_tile, col0 = tn * GBN;
      mbar_wait(acc_full + buf, (it >> 1) & 1);
      tc_fence_after();
      float rs = 0.f;
#pragma unroll 1
      for (int c = 0; c < GBN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_base + buf * GBN + c * 32, v);
        tmem_ld_wait();
        const int64_t col = col0 + c * 32;
        if (col >= g.N) continue;
        const bool full_chunk = col + 32 <= g.N;
        if constexpr (EPI == EPI_F32) {
          float* dst = static_cast<float*>(g.C) + row * g.ldc + col;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = __uint_as_float(v[i]) * g.alpha;
            if (g.bias && col + i < g.N) x += __ldg(g.bias + col + i);
            v[i] = __float_as_uint(x);
          }
          if (row < g.M) {
            if (full_chunk && (g.ldc & 3) == 0) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) *reinterpret_cast<uint4*>(dst + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
              for (int i = 0; i < 32 && col + i < g.N; ++i) dst[i] = __uint_as_float(v[i]);
            }
          }
        } else {
          uint32_t u[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float x0, x1;
            if constexpr (EPI == EPI_EXP) {
              x0 = col + i < g.N ? ex2_approx(fmaf(__uint_as_float(v[i]), g.alpha, -g.alpha)) : 0.f;
              x1 = col + i + 1 < g.N ? ex2_approx(fmaf(__uint_as_float(v[i + 1]), g.alpha, -g.alpha)) : 0.f;
              rs += x0 + x1;
            } else {
              x0 = __uint_as_float(v[i]) * g.alpha;
              x1 = __uint_as_float(v[i + 1]) * g.alpha;
              if (g.bias) {
                if (col + i < g.N) x0 += __ldg(g.bias + col + i);
                if (col + i + 1 < g.N) x1 += __ldg(g.bias + col + i + 1);
              }
            }
            __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
            u[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
          }
          if (row < g.M) {
            __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(g.C) + row * g.ldc + col;
            // EPI_EXP writes whole 32-column chunks inside the padded leading dimension (zeros past N)
            const bool can_vec = (g.ldc & 7) == 0 && (full_chunk || (EPI == EPI_EXP && col + 32 <= g.ldc));
            if (can_vec) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) *reinterpret_cast<uint4*>(dst + 2 * i) = make_uint4(u[i], u[i + 1], u[i + 2], u[i + 3]);
            } else {
              const int64_t lim = EPI == EPI_EXP ? g.ldc : g.N;
              for (int i = 0; i < 32 && col + i < lim; ++i)
                dst[i] = reinterpret_cast<const __nv_bfloat16*>(u)[i];
            }
          }
        }
      }
      if constexpr (EPI == EPI_EXP) {
        if (g.rowsum_part && row < g.m_pad) g.rowsum_part[tn * g.m_pad + row] = rs;
      }
      tc_fence_before();
      mbar_arrive(acc_empty + buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

template <int EPI>
int launch_epi(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& g, cudaStream_t st) {
  MSF_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  const int64_t tiles = ((g.M + BM - 1) / BM) * ((g.N + GBN - 1) / GBN);
  const unsigned grid = static_cast<unsigned>(tiles < kNumSMs ? tiles : kNumSMs);
  gemm_tc_kernel<EPI><<<grid, kThreads, kSmem, st>>>(ta, tb, g);
  MSF_LAUNCH_OK("gemm_tc_kernel");
  return MSF_OK;
}

}  // namespace

// Internal entry used by the InfoNCE two-pass path and by msf_gemm_bf16.
int launch_gemm_tc(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                   int b_mn_major, int epi, float alpha, const float* bias, float* rowsum_part, int64_t m_pad, cudaStream_t st,
                   int a_mn_major) {
  MSF_REQUIRE(M > 0 && N > 0 && K > 0, MSF_ERR_INVALID, "empty GEMM %lld x %lld x %lld", (long long)M, (long long)N, (long long)K);
  MSF_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), MSF_ERR_UNSUPPORTED, "GEMM extents must fit 31 bits");
  MSF_REQUIRE(A && B && C, MSF_ERR_INVALID, "NULL operand");
  CUtensorMap ta, tb;
  if (a_mn_major) {
    if (int rc = make_map_bf16(&ta, A, K, M, lda, 64, GBK)) return rc;
  } else {
    if (int rc = make_map_bf16(&ta, A, M, K, lda, GBK, BM)) return rc;
  }
  if (b_mn_major) {
    if (int rc = make_map_bf16(&tb, B, K, N, ldb, 64, GBK)) return rc;
  } else {
    if (int rc = make_map_bf16(&tb, B, N, K, ldb, GBK, GBN)) return rc;
  }
  GemmArgs g{M, N, K, ldc, C, bias, alpha, rowsum_part, m_pad, b_mn_major, a_mn_major};
  switch (epi) {
    case EPI_F32: return launch_epi<EPI_F32>(ta, tb, g, st);
    case EPI_BF16: return launch_epi<EPI_BF16>(ta, tb, g, st);
    case EPI_EXP: return launch_epi<EPI_EXP>(ta, tb, g, st);
    default: set_error("bad epilogue %d", epi); return MSF_ERR_INVALID;
  }
}

}  // namespace msf

using namespace msf;

extern "C" int msf_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int64_t M, int64_t N,
                             int64_t K, int a_is_km, int b_is_kn, int out_dtype, float alpha, const float* bias, void* stream) {
  MSF_REQUIRE(out_dtype == MSF_F32 || out_dtype == MSF_BF16, MSF_ERR_INVALID, "out_dtype must be MSF_F32 or MSF_BF16");
  ProfScope prof(stream, MSF_K_GEMM, 2.0 * static_cast<double>(M) * static_cast<double>(N) * static_cast<double>(K));
  return launch_gemm_tc(A, lda, B, ldb, C, ldc, M, N, K, b_is_kn, out_dtype == MSF_F32 ? 0 : 1, alpha, bias, nullptr, 0,
                        static_cast<cudaStream_t>(stream), a_is_km);
}
