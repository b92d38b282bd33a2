// A1: inverse-jigsaw gather + fuser concat (and their backward), every level and both views in
// one launch.  Replaces src/models/backbone.py:147-158 and :195-202 of the reference.
//
// HBM-bound copy kernel.  Each item is split into "row-copy segments"; a thread moves 16-byte
// chunks (128-bit loads/stores, L1 bypassed, 4 independent chunks in flight per thread).
// Algorithmic bytes per item (e = element size): read 16B*d*e (+B*d*e ctx) + 16B*8 index bytes,
// write 16B*d*e (sorted) + 9B*d*e (ms).
#include "common.cuh"

namespace msf {
namespace {

struct Segment {
  const char* src;     // base of the source rows
  const char* src2;    // optional second source added to the first (backward only)
  char* dst;
  char* dst2;          // forward gather only: second destination (fuser rows) for source rows < n_keep
  const int64_t* idx;  // optional (B,K) permutation
  int64_t rows;        // rows in this segment
  uint32_t block_prefix;  // first CTA of this segment
  uint32_t cpr;        // 16-byte chunks per row
  int32_t shift;       // log2(cpr) if cpr is a power of two, else -1
  uint32_t src_stride, src2_stride, dst_stride, dst2_stride;  // row strides in chunks
  int32_t mode;        // 0 plain copy; 1 gather src row via idx; 2 scatter dst row via idx (+src2 if dstrow%K<n_keep)
};

constexpr int kMaxSeg = 3 * MSF_GATHER_MAX_ITEMS;
struct Params {
  Segment seg[kMaxSeg];
  int n_seg;
  int K, n_keep;
  uint32_t total_blocks;
};

constexpr int kIlp = 4, kIter = 2, kThreads = 256;
constexpr int kChunksPerBlock = kThreads * kIlp * kIter;  // 32 KB moved per CTA

template <int DT>
__global__ void __launch_bounds__(kThreads) row_copy_kernel(const __grid_constant__ Params P, int32_t* status) {
  int s = 0;  // block-uniform: which segment this CTA works on
#pragma unroll 1
  while (s + 1 < P.n_seg && blockIdx.x >= P.seg[s + 1].block_prefix) ++s;
  const Segment& g = P.seg[s];
  // 32-bit index arithmetic inside a segment (the host rejects segments with >= 2^31 chunks)
  const uint32_t seg_chunks = static_cast<uint32_t>(g.rows) * g.cpr;
  const uint32_t chunk0 = (blockIdx.x - g.block_prefix) * kChunksPerBlock + threadIdx.x;
  const uint32_t K = static_cast<uint32_t>(P.K);
  const bool k_pow2 = (K & (K - 1)) == 0;
#pragma unroll 1
  for (int it = 0; it < kIter; ++it) {
    uint4 v[kIlp], v2[kIlp];
    char* dst[kIlp];
    char* dstb[kIlp];
    bool has2[kIlp];
#pragma unroll
    for (int u = 0; u < kIlp; ++u) {
      const uint32_t local = chunk0 + static_cast<uint32_t>(it * kIlp + u) * kThreads;
      dst[u] = nullptr;
      dstb[u] = nullptr;
      has2[u] = false;
      if (local >= seg_chunks) continue;
      uint32_t row, c;
      if (g.shift >= 0) {
        row = local >> g.shift;
        c = local & (g.cpr - 1);
      } else {
        row = local / g.cpr;
        c = local - row * g.cpr;
      }
      uint32_t srow = row, drow = row;
      if (g.mode != 0) {
        const uint32_t bk = k_pow2 ? (row & ~(K - 1)) : (row / K) * K;  // first row of this sample
        long long r = g.idx[row];
        if (r < -static_cast<long long>(K) || r >= static_cast<long long>(K)) {  // the reference raises IndexError
          if (status) atomicOr(status, 1);
          r = 0;
        }
        if (r < 0) r += K;
        const uint32_t ru = static_cast<uint32_t>(r);
        if (g.mode == 1) {
          srow = bk + ru;
          if (g.dst2 && ru < static_cast<uint32_t>(P.n_keep)) {  // the same 16 bytes also feed ms[b, (1+r)*d + c]
            const uint32_t b = k_pow2 ? (row >> (31 - __clz(K))) : row / K;
            dstb[u] = g.dst2 + (static_cast<size_t>(b) * g.dst2_stride + ru * g.cpr + c) * 16;
          }
        } else {
          drow = bk + ru;
          has2[u] = g.src2 != nullptr && ru < static_cast<uint32_t>(P.n_keep);
          if (has2[u]) {  // g_ms[b, (1+r)*d + c]  -- src2 points at g_ms + d (first target slot)
            const uint32_t b = k_pow2 ? (row >> (31 - __clz(K))) : row / K;
            v2[u] = ldg_stream(g.src2 + (static_cast<size_t>(b) * g.src2_stride + ru * g.cpr + c) * 16);
          }
        }
      }
      dst[u] = g.dst + (static_cast<size_t>(drow) * g.dst_stride + c) * 16;
      if (g.src) v[u] = ldg_stream(g.src + (static_cast<size_t>(srow) * g.src_stride + c) * 16);
      else v[u] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < kIlp; ++u) {
      if (!dst[u]) continue;
      if (has2[u]) {
        float a[Elem<DT>::VEC], b2[Elem<DT>::VEC];
        Elem<DT>::unpack(v[u], a);
        Elem<DT>::unpack(v2[u], b2);
#pragma unroll
        for (int i = 0; i < Elem<DT>::VEC; ++i) a[i] += b2[i];
        v[u] = Elem<DT>::pack(a);
      }
      stg_stream(dst[u], v[u]);
      if (dstb[u]) stg_stream(dstb[u], v[u]);
    }
  }
}

int push(Params& P, const void* src, const void* src2, void* dst, const int64_t* idx, int64_t rows, uint32_t cpr,
         uint32_t ss, uint32_t s2s, uint32_t ds, int mode, void* dst2 = nullptr, uint32_t d2s = 0) {
  if (rows == 0 || cpr == 0) return 0;
  if (rows * static_cast<int64_t>(cpr) >= (1ll << 31)) return -1;
  Segment& g = P.seg[P.n_seg++];
  g.src = static_cast<const char*>(src);
  g.src2 = static_cast<const char*>(src2);
  g.dst = static_cast<char*>(dst);
  g.dst2 = static_cast<char*>(dst2);
  g.dst2_stride = d2s;
  g.idx = idx;
  g.rows = rows;
  g.block_prefix = P.total_blocks;
  g.cpr = cpr;
  g.shift = -1;
  for (int s = 0; s < 31; ++s)
    if ((1u << s) == cpr) g.shift = s;
  g.src_stride = ss;
  g.src2_stride = s2s;
  g.dst_stride = ds;
  g.mode = mode;
  P.total_blocks += static_cast<uint32_t>((rows * static_cast<int64_t>(cpr) + kChunksPerBlock - 1) / kChunksPerBlock);
  return 0;
}

int launch(const Params& P, int dtype, int32_t* status, cudaStream_t st) {
  if (P.total_blocks == 0) return MSF_OK;
  MSF_DISPATCH_DTYPE(dtype, (row_copy_kernel<DT><<<P.total_blocks, kThreads, 0, st>>>(P, status)));
  MSF_LAUNCH_OK("row_copy_kernel");
  return MSF_OK;
}

int check_common(int n_items, int64_t B, int K, int n_keep, int dtype) {
  MSF_REQUIRE(n_items >= 0 && n_items <= MSF_GATHER_MAX_ITEMS, MSF_ERR_INVALID, "n_items %d outside [0,%d]", n_items,
              MSF_GATHER_MAX_ITEMS);
  MSF_REQUIRE(B >= 0 && K > 0 && n_keep >= 0 && n_keep <= K, MSF_ERR_INVALID, "bad B=%lld K=%d n_keep=%d",
              static_cast<long long>(B), K, n_keep);
  MSF_REQUIRE(dtype_ok(dtype), MSF_ERR_INVALID, "bad dtype %d", dtype);
  return MSF_OK;
}

}  // namespace
}  // namespace msf

using namespace msf;

extern "C" int msf_gather_concat_fwd(const msf_gather_item* items, int n_items, int64_t B, int K, int n_keep,
                                     int dtype, int32_t* status_flag, void* stream) {
  if (int rc = check_common(n_items, B, K, n_keep, dtype)) return rc;
  MSF_REQUIRE(items || n_items == 0, MSF_ERR_INVALID, "items is NULL");
  Params P{};
  const uint32_t vec = 16 / dtype_size(dtype);
  for (int i = 0; i < n_items; ++i) {
    const msf_gather_item& it = items[i];
    MSF_REQUIRE(it.d > 0 && it.d % vec == 0, MSF_ERR_INVALID, "item %d: d=%d must be a positive multiple of %u", i, it.d, vec);
    MSF_REQUIRE(it.tgt_f && it.ctx_f && it.rev && it.tgt_sorted && it.ms_f, MSF_ERR_INVALID, "item %d: NULL pointer", i);
    MSF_REQUIRE(aligned16(it.tgt_f) && aligned16(it.ctx_f) && aligned16(it.tgt_sorted) && aligned16(it.ms_f),
                MSF_ERR_INVALID, "item %d: pointers must be 16-byte aligned", i);
    const uint32_t cpr = it.d / vec;
    const uint32_t ms_stride = (n_keep + 1) * cpr;
    // sorted[b*K+j] = tgt_f[b*K + rev[b,j]]
    int bad = push(P, it.tgt_f, nullptr, it.tgt_sorted, it.rev, B * K, cpr, cpr, 0, cpr, 1);
    // ms[b, 0:d] = ctx_f[b]
    bad |= push(P, it.ctx_f, nullptr, it.ms_f, nullptr, B, cpr, cpr, 0, ms_stride, 0);
    // ms[b, d:(1+n_keep)*d] = tgt_f[b*K : b*K + n_keep].flatten()  -- the first n_keep SHUFFLED vectors of the sample are
    // adjacent rows, so this is one contiguous n_keep*d copy per sample, independent of `rev` (backbone.py:195-202 builds
    // ms_f from target_f_split[:, :n_keep] whatever the index tensor holds); the re-read is served by L2
    bad |= push(P, it.tgt_f, nullptr, static_cast<char*>(it.ms_f) + static_cast<size_t>(cpr) * 16, nullptr, B,
                static_cast<uint32_t>(n_keep) * cpr, static_cast<uint32_t>(K) * cpr, 0, ms_stride, 0);
    MSF_REQUIRE(!bad, MSF_ERR_UNSUPPORTED, "item %d: more than 2^31 16-byte chunks in one tensor", i);
  }
  P.K = K;
  P.n_keep = n_keep;
  double bytes = 0.0;  // per item: tgt read + sorted written, ctx read, ms written, rev read
  for (int i = 0; i < n_items; ++i) bytes += (2.0 * B * K + B + (n_keep + 1.0) * B) * items[i].d * dtype_size(dtype) + 8.0 * B * K;
  ProfScope prof(stream, MSF_K_GATHER_FWD, bytes);
  return launch(P, dtype, status_flag, static_cast<cudaStream_t>(stream));
}

extern "C" int msf_gather_concat_bwd(const msf_gather_grad_item* items, int n_items, int64_t B, int K, int n_keep,
                                     int dtype, void* stream) {
  if (int rc = check_common(n_items, B, K, n_keep, dtype)) return rc;
  MSF_REQUIRE(items || n_items == 0, MSF_ERR_INVALID, "items is NULL");
  Params P{};
  const uint32_t vec = 16 / dtype_size(dtype);
  for (int i = 0; i < n_items; ++i) {
    const msf_gather_grad_item& it = items[i];
    MSF_REQUIRE(it.d > 0 && it.d % vec == 0, MSF_ERR_INVALID, "item %d: d=%d must be a positive multiple of %u", i, it.d, vec);
    MSF_REQUIRE(it.rev && it.g_tgt_f && it.g_ctx_f, MSF_ERR_INVALID, "item %d: NULL pointer", i);
    MSF_REQUIRE(aligned16(it.g_sorted) && aligned16(it.g_ms) && aligned16(it.g_tgt_f) && aligned16(it.g_ctx_f),
                MSF_ERR_INVALID, "item %d: pointers must be 16-byte aligned", i);
    const uint32_t cpr = it.d / vec;
    const uint32_t ms_stride = (n_keep + 1) * cpr;
    const char* ms_tgt = it.g_ms ? static_cast<const char*>(it.g_ms) + static_cast<size_t>(cpr) * 16 : nullptr;
    // g_tgt_f[b*K + rev[b,j]] = g_sorted[b*K+j] (+ g_ms[b, (1+rev)*d ...] when rev < n_keep)
    int bad = push(P, it.g_sorted, ms_tgt, it.g_tgt_f, it.rev, B * K, cpr, cpr, ms_stride, cpr, 2);
    // g_ctx_f[b] = g_ms[b, 0:d]
    bad |= push(P, it.g_ms, nullptr, it.g_ctx_f, nullptr, B, cpr, ms_stride, 0, cpr, 0);
    MSF_REQUIRE(!bad, MSF_ERR_UNSUPPORTED, "item %d: more than 2^31 16-byte chunks in one tensor", i);
  }
  P.K = K;
  P.n_keep = n_keep;
  double bytes = 0.0;
  for (int i = 0; i < n_items; ++i) bytes += (2.0 * B * K + B + (n_keep + 1.0) * B) * items[i].d * dtype_size(dtype) + 8.0 * B * K;
  ProfScope prof(stream, MSF_K_GATHER_BWD, bytes);
  return launch(P, dtype, nullptr, static_cast<cudaStream_t>(stream));
}
