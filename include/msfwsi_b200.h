/*
 * msfwsi_b200 -- C ABI of the B200-native MSF-WSI SSL head + loss hot path.
 *
 * The reference (Dylan-H-Wang/msf-wsi) has no FFI / plugin boundary: its hot path is plain
 * PyTorch inside `MSFWSI.forward` (src/models/backbone.py:129-222) and the loss block of
 * `train` (tools/ssl_train.py:448-466).  This header is the boundary a maintainer would bind
 * (ctypes stub in INTEGRATION.md); every entry point cites the reference lines it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless a parameter
 *     is documented as host memory; `stream` is a cudaStream_t passed as void*.
 *   - the library never allocates, frees or synchronises the device: the caller owns every
 *     buffer including workspaces (`msf_*_workspace_bytes`), and every call is asynchronous on
 *     `stream`.
 *   - return value: MSF_OK (0) or a negative msf_status; the message of the last failure on
 *     the calling thread is available from msf_last_error().
 *   - re-entrant: no global mutable state; callable from the autograd engine thread.
 *   - row-major, contiguous tensors; base pointers 16-byte aligned; feature widths must be
 *     multiples of 8 elements (the reference widths are 64..4608).
 */
#ifndef MSFWSI_B200_H_
#define MSFWSI_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSF_ABI_VERSION 2

typedef enum { MSF_F32 = 0, MSF_BF16 = 1, MSF_F16 = 2 } msf_dtype;

typedef enum {
  MSF_OK = 0,
  MSF_ERR_INVALID = -1,     /* bad argument (null pointer, misalignment, bad shape) */
  MSF_ERR_UNSUPPORTED = -2, /* shape / dtype / device outside what the kernels cover */
  MSF_ERR_WORKSPACE = -3,   /* workspace too small */
  MSF_ERR_CUDA = -4         /* a CUDA runtime / driver call failed */
} msf_status;

int msf_abi_version(void);
/* Thread-local, never NULL. */
const char* msf_last_error(void);
/* Fills sm_count / compute-capability of the current device; MSF_ERR_UNSUPPORTED unless cc == 10.0. */
int msf_device_check(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * A1  inverse-jigsaw gather + fuser concat, all levels and both views in one launch.
 * Replaces src/models/backbone.py:147-158 (reshape + advanced-index gather) and :195-202
 * (torch.cat of the context vector with the first n_keep shuffled target vectors).
 *   tgt_sorted[b*K + j, :] = tgt_f[b*K + rev[b, j], :]
 *   ms_f[b, :]             = [ ctx_f[b, :] | tgt_f[b*K + 0, :] | ... | tgt_f[b*K + n_keep-1, :] ]
 * Exact copies (bit-exact for every dtype).  rev entries outside [-K, K) set bit 0 of
 * *status_flag (device int32, may be NULL) and are clamped; the reference raises IndexError.
 * ms_f is filled from tgt_f independently of rev, as in the reference, so the forward is defined for ANY in-range index
 * tensor; the backward (a scatter by the same indices) requires every rev row to be a permutation, which is what the
 * datasets produce (argsort(randperm(K)), src/utils/data/bcss.py:171-177).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const void* tgt_f;    /* (B*K, d)  target vectors in the shuffled order the encoder saw */
  const void* ctx_f;    /* (B, d) */
  const int64_t* rev;   /* (B, K)    jigsaw_idx of this view */
  void* tgt_sorted;     /* (B*K, d)  out */
  void* ms_f;           /* (B, (n_keep+1)*d) out */
  int32_t d;
  int32_t reserved;
} msf_gather_item;

#define MSF_GATHER_MAX_ITEMS 16
int msf_gather_concat_fwd(const msf_gather_item* items /*host*/, int n_items, int64_t B, int K, int n_keep,
                          int dtype, int32_t* status_flag, void* stream);

/* Backward of the above:
 *   g_tgt_f[b*K + k, :] = g_sorted[b*K + inv(b)[k], :] + (k < n_keep ? g_ms[b, (1+k)*d : (2+k)*d] : 0)
 *   g_ctx_f[b, :]       = g_ms[b, 0:d]
 * with inv(b) the inverse permutation of rev[b, :].  g_sorted or g_ms may be NULL (treated as 0). */
typedef struct {
  const void* g_sorted; /* (B*K, d) or NULL */
  const void* g_ms;     /* (B, (n_keep+1)*d) or NULL */
  const int64_t* rev;   /* (B, K) */
  void* g_tgt_f;        /* (B*K, d) out */
  void* g_ctx_f;        /* (B, d)   out */
  int32_t d;
  int32_t reserved;
} msf_gather_grad_item;
int msf_gather_concat_bwd(const msf_gather_grad_item* items /*host*/, int n_items, int64_t B, int K, int n_keep,
                          int dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * L1 (cosine mode)  SimSiam negative-cosine loss over many (p, z) pairs in one launch.
 * Replaces the 24 nn.CosineSimilarity(dim=1) + .mean() calls of tools/ssl_train.py:422,448-466:
 *   *loss_out = sum_pairs coef_pair * mean_i cos(p_i, z_i),  cos with ATen's per-vector
 *   clamp_min(||v||, eps); coef_pair = -0.5 * fuser_weight[level].
 * Arithmetic is fp32 for every input dtype (CUDA autocast runs cosine_similarity in fp32);
 * the reduction order is fixed (deterministic).  row_stats receives 4 floats per row
 * {cos, 1/max(||p||,eps), 1/max(||z||,eps), ||p||>=eps} consumed by the backward.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const void* p;       /* (rows, dim) predictor output (receives gradient) */
  const void* z;       /* (rows, dim) detached projector output */
  float* row_stats;    /* (rows, 4) fp32 */
  void* grad_p;        /* (rows, dim) out of the backward, dtype of p; unused by the forward */
  int64_t rows;
  int32_t dim;
  float coef;
} msf_cos_pair;

#define MSF_COS_MAX_PAIRS 32
size_t msf_cosine_loss_workspace_bytes(const msf_cos_pair* pairs /*host*/, int n_pairs);
int msf_cosine_loss_fwd(const msf_cos_pair* pairs /*host*/, int n_pairs, int dtype, float eps, float* loss_out,
                        void* workspace, size_t workspace_bytes, void* stream);
/* grad_p[i,:] = *grad_out * coef/rows * d cos(p_i,z_i)/d p_i ; grad_out is a DEVICE fp32 scalar
 * (GradScaler's non-unit upstream gradient, tools/ssl_train.py:472, without a host sync). */
int msf_cosine_loss_bwd(const msf_cos_pair* pairs /*host*/, int n_pairs, int dtype, const float* grad_out,
                        void* stream);

/* ------------------------------------------------------------------------------------------
 * L1 (infonce mode)  flash-style fused InfoNCE (extension named by BASELINE.json:north_star; the
 * reference has no such loss -- parity unpinned, oracle = oracle/msf_oracle.py:infonce_loss).
 *   row_loss_i = logsumexp_j(q_hat_i . k_hat_j / tau) - q_hat_i . k_hat_{pos_offset+i} / tau
 * Keys are detached (backbone.py:188-191): one pass over the keys yields the loss AND
 * O_i = sum_j exp(.)_ij k_hat_j, from which the backward forms dq without touching the keys again.
 * The N x N logits are never materialised.
 *
 * Step 1  msf_rownorm: x -> x_hat = x / max(||x||, eps) (bf16 or fp32 out) and 1/max(||x||,eps).
 * Step 2  (multi-GPU) the caller all-gathers k_hat over NCCL; rank-major order.
 * Step 3  msf_infonce_fwd: loss pieces + O partials into the workspace.
 * Step 4  msf_infonce_bwd: grad_q (dtype of q) from the saved workspace.
 *
 * precision: MSF_F32  -> fp32 SIMT kernel (q_hat/k_hat fp32), for <=1e-5 parity.
 *            MSF_BF16 -> TMA + tcgen05/TMEM kernels (q_hat/k_hat bf16): flash-style single pass for dim in
 *                        {64,128,256}; two tcgen05 GEMM passes with P materialised in bf16 for larger multiples
 *                        of 64 (512 and the fuser widths 576..4608), where TMEM cannot hold O and S together.
 * tau must satisfy 2*log2(e)/tau <= 120 (tau >= 0.0241): the softmax uses the fixed bound
 * max_j s_ij <= 1/tau that L2-normalised operands guarantee, so no running max is kept.
 * ---------------------------------------------------------------------------------------- */
int msf_rownorm(const void* x, int64_t rows, int dim, int in_dtype, float eps, void* x_hat, int out_dtype,
                float* inv_norm, void* stream);

size_t msf_infonce_workspace_bytes(int64_t nq, int64_t n_keys, int dim, int precision);
/* Workspace layout for inspection / tests: info[0]=splits, [1]=nq_pad, [2]=byte offset of the fp32 row-sum
 * partials [splits][nq_pad], [3]=byte offset of the fp32 O partials [splits][nq_pad][dim], [4]=offset of the
 * positive cosines [nq], [5]=offset of the total row sums [nq], [6]=query rows per CTA, [7]=keys per tile. */
int msf_infonce_plan_info(int64_t nq, int64_t n_keys, int dim, int precision, int64_t* info /*host, 8*/);
/* loss_sum_out: device fp32 scalar that receives sum_i row_loss_i (NOT divided; the caller
 * divides by the global row count).  row_lse (nq) may be NULL. */
int msf_infonce_fwd(const void* q_hat, const void* k_hat, int64_t nq, int64_t n_keys, int dim,
                    int64_t pos_offset, float tau, int precision, float* loss_sum_out, float* row_lse,
                    void* workspace, size_t workspace_bytes, void* stream);
/* Same as msf_infonce_fwd, additionally recording two caller-owned cudaEvent_t (passed as void*, may be NULL) on
 * `stream` immediately before and after the MAIN kernel (flash tcgen05 kernel, or the two GEMM passes, or the SIMT
 * kernel) -- used by bench.py to time the dominant kernel alone, without the small finalize launches. */
int msf_infonce_fwd_timed(const void* q_hat, const void* k_hat, int64_t nq, int64_t n_keys, int dim, int64_t pos_offset,
                          float tau, int precision, float* loss_sum_out, float* row_lse, void* workspace,
                          size_t workspace_bytes, void* stream, void* ev_main_start, void* ev_main_stop);
/* grad_q = *grad_out * scale * J_normalize(q)^T [ (O_i / rowsum_i - k_hat_pos(i)) / tau ],
 * scale = 1 / n_rows_global supplied by the caller. */
int msf_infonce_bwd(const void* q_hat, const void* k_hat, const float* q_inv_norm, int64_t nq, int64_t n_keys,
                    int dim, int64_t pos_offset, float tau, int precision, const float* grad_out, float scale,
                    const void* workspace, size_t workspace_bytes, void* grad_q, int grad_dtype, void* stream);

/* Key gradient -- the variant of BASELINE.json:north_star (4) in which the keys are NOT detached (the reference detaches
 * every key, backbone.py:188-191, so MSFWSI itself never calls this; ops.infonce_loss(..., detach_keys=False) does):
 *   dk_hat_j = *grad_out * scale / tau * ( sum_{i local} softmax_ij q_hat_i  -  q_hat_{i : pos(i) = j} ).
 * Step 5  msf_infonce_dk: dk_out (n_keys, dim) fp32 = *grad_out * scale / tau * sum_{i local} softmax_ij q_hat_i for EVERY
 *         global key j -- the flash pass with the roles swapped (rows = keys, columns = this rank's queries; 1/sum_i from
 *         the forward workspace becomes a per-column exponent term), 4 * nq * n_keys * dim FLOP.  fwd_workspace is the
 *         workspace msf_infonce_fwd filled for the same (q_hat, k_hat).
 * Step 6  (multi-GPU) the caller reduce-scatters dk_out over the ranks (NCCL, rank-major like the all-gather): every rank
 *         receives the summed rows of its own keys.
 * Step 7  msf_infonce_dk_finish: subtracts the positive term (local: key j's query is row j of this rank's q_hat, j < nq)
 *         and applies the normalise Jacobian: grad_z_j = (d - k_hat_j (k_hat_j . d)) * k_inv_norm[j], in grad_dtype. */
size_t msf_infonce_dk_workspace_bytes(int64_t nq, int64_t n_keys, int dim, int precision);
int msf_infonce_dk(const void* q_hat, const void* k_hat, int64_t nq, int64_t n_keys, int dim, int64_t pos_offset, float tau,
                   int precision, const float* grad_out, float scale, const void* fwd_workspace, size_t fwd_workspace_bytes,
                   float* dk_out, void* workspace, size_t workspace_bytes, void* stream);
int msf_infonce_dk_finish(const float* dk_local, const void* q_hat, const void* k_hat_local, const float* k_inv_norm,
                          int64_t rows, int64_t nq, int dim, float tau, int precision, const float* grad_out, float scale,
                          void* grad_z, int grad_dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * L1 (infonce mode, grouped)  the InfoNCE terms of ALL (branch, level, direction) pairs of the loss block in a handful of
 * launches: one flash launch per width class D in {64, 128, 256}, two grouped GEMM launches per chunk of rank blocks for
 * the wider pairs, one finalize (+ fixed-order sum) and ONE backward launch.
 *   loss = sum_pairs coef_pair * mean_i [ logsumexp_j(q_hat_i . k_hat_j / tau) - q_hat_i . k_hat_pos(i) / tau ]
 * q: (nq, D) bf16 queries.  q_rowsq != NULL: q are RAW predictor outputs and q_rowsq [D/64][nq] their per-64-column sums of
 * squares (the predictor-tail GEMM's epilogue): the kernels divide each row's logits by max(||q_i||, eps) themselves.
 * keys: L2-normalised bf16 keys as rank-major blocks: block r = keys + r * rank_stride elements, (rows_per_rank, D); the
 * positive of local query i is row i of block pos_rank.  world = 1: a single block.  Keys carry no gradient.
 * msf_nce_grouped_bwd writes grad_q (nq, D) bf16 = *grad_out * d loss / d q (grad_out: DEVICE fp32 scalar).
 * The workspace (msf_nce_grouped_workspace_bytes) must be the same memory in forward and backward.
 * ---------------------------------------------------------------------------------------- */
#define MSF_NCE_MAX_PAIRS 32
typedef struct {
  const void* q;
  const float* q_rowsq;
  const void* keys;
  void* grad_q;          /* backward only */
  int64_t rank_stride;   /* elements */
  int32_t nq, rows_per_rank, world, D, pos_rank;
  float coef;
} msf_nce_pair;
size_t msf_nce_grouped_workspace_bytes(const msf_nce_pair* pairs /*host*/, int n_pairs);
int msf_nce_grouped_fwd(const msf_nce_pair* pairs /*host*/, int n_pairs, int dtype, float tau, float eps, float* loss_out,
                        void* workspace, size_t workspace_bytes, void* stream);
int msf_nce_grouped_bwd(const msf_nce_pair* pairs /*host*/, int n_pairs, int dtype, float tau, float eps, const float* grad_out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * G1  bf16 tcgen05 GEMM  C[M,N] = alpha * A[M,K] * op(B) (+ bias[N]), fp32 accumulate in TMEM.
 * The Linear layers of the heads (src/models/backbone.py:14,17,20,27,30): y = x W^T is b_is_kn = 0 with
 * B = W [N=out, K=in]; the input gradient dX = dY W is b_is_kn = 1 with B = W [K=out, N=in].
 * the weight gradient dW = dY^T X is a_is_km = 1 (A = dY stored [K=rows, M=out]) with b_is_kn = 1 (B = X [K=rows, N=in]).
 * A [M,K] (a_is_km=0) or [K,M] (a_is_km=1) row-major (lda), B [N,K] (b_is_kn=0) or [K,N] (b_is_kn=1) row-major (ldb),
 * both bf16; C row-major (ldc), MSF_F32 or MSF_BF16.  lda/ldb multiples of 8, bases 16-byte aligned.
 * ---------------------------------------------------------------------------------------- */
int msf_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int64_t M, int64_t N,
                  int64_t K, int a_is_km, int b_is_kn, int out_dtype, float alpha, const float* bias, void* stream);

/* ------------------------------------------------------------------------------------------
 * G1 (grouped)  ONE persistent tcgen05 launch over a table of GEMM problems: the head stage of
 * src/models/backbone.py:12-31, 161-186, 205-212 -- the Linear layers of the same depth of all 12 projectors / predictors
 * and both views in one launch (forward), their dX and dW GEMMs in one launch (backward).  Per problem
 *   C[M,N] = epi( alpha * pro(A)[M,K] * op(B) ),   operands bf16 or fp16 (op_dtype), fp32 accumulate in TMEM
 *   pro: a' = relu?(a * a_scale[k] + a_shift[k]) applied to the A tile in shared memory (the PREVIOUS layer's BatchNorm1d
 *        apply + ReLU, backbone.py:15-16,18-19,28-29; rounded to the operand dtype before the ReLU like the reference's
 *        16-bit BatchNorm output); needs A stored [M][K]
 *   epi: + bias[n]; rounding to out_dtype; optional statistics of the ROUNDED outputs:
 *        col_stats [ceil(M/128)][2][N] fp32 = per 128-row M tile g: {sum_r y[r,n], sum_r y[r,n]^2} (rows past M excluded) --
 *        the batch-norm statistics of THIS layer, merged in fp64 by msf_head_bn_finalize;
 *        row_sumsq [ceil(N/64)][M] fp32 = per 64-column block: sum_n y[m,n]^2 -- the row norms the loss needs.
 * A [M,K] (a_is_km=0) or [K,M] (a_is_km=1) row-major (lda), B [N,K] (b_is_kn=0: C = A B^T) or [K,N] (b_is_kn=1: C = A B)
 * row-major (ldb); lda, ldb multiples of 8 (K too when the prologue is used); bases 16-byte aligned; C row-major (ldc),
 * out_dtype = MSF_F32 or op_dtype.
 * y = x W^T: b_is_kn=0, B = W; dX = dY W: b_is_kn=1, B = W; dW = dY^T X: a_is_km=1 (A = dY [rows,out]), b_is_kn=1 (B = X).
 * tile_n: 0 = chosen by the library (64 / 128 / 256 so that small-M problems still cover the SMs), else forced.
 * split_k: 0 = chosen by the library (deterministic split-K for few-tile, long-K problems such as the target heads' dW with
 * K = rows), > 0 forced, < 0 never.  Split partials live in `workspace` (msf_gemm_grouped_workspace_bytes) and are reduced in
 * split order by the last CTA to arrive (no float atomics: bit-reproducible); `counters` = MSF_GEMM_MAX_COUNTERS int32, zero
 * before the first call, left zero by every call (one array per stream).
 * ---------------------------------------------------------------------------------------- */
#define MSF_GEMM_MAX_PROBLEMS 48
#define MSF_GEMM_MAX_COUNTERS 8192
typedef struct {
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* C; int64_t ldc;
  int32_t M, N, K;
  int32_t a_is_km, b_is_kn;
  int32_t out_dtype;
  float alpha;
  const float* bias;      /* [N] or NULL */
  float* col_stats;       /* NULL or [ceil(M/32)][2][N] */
  float* row_sumsq;       /* NULL or [ceil(N/64)][M] */
  const float* a_scale;   /* NULL (no prologue) or [K] */
  const float* a_shift;   /* [K] */
  int32_t a_relu;
  int32_t tile_n;
  int32_t split_k;
  int32_t no_tma_store;   /* 1: 16-bit outputs leave through per-thread row segments instead of shared memory + TMA store */
  /* EXP epilogue (the two-pass InfoNCE path): exp_a != 0 -> C = exp2(exp_a * row_scale[m] * acc - exp_a) in the 16-bit output
   * dtype (0 past column N), and row_sumsq receives per-64-column-block SUMS of C instead of sums of squares. */
  float exp_a;
  int32_t row_sum_ld;     /* leading dimension of row_sumsq; 0 = M */
  const float* row_scale; /* [M] or NULL (= 1) */
} msf_gemm_problem;
size_t msf_gemm_grouped_workspace_bytes(const msf_gemm_problem* problems /*host*/, int n_problems);
int msf_gemm_grouped(const msf_gemm_problem* problems /*host*/, int n_problems, int op_dtype, void* workspace,
                     size_t workspace_bytes, int32_t* counters, void* stream);
/* The same table with fp32 operands and outputs on a plain-FMA SIMT kernel (exact fp32 accumulation in k order): the path
 * of fp32 parameters / activations outside autocast (<= 1e-5 parity cases).  No prologue / statistics epilogues here
 * (msf_head_bn_stats / msf_head_bn_apply do that work); bias and alpha are honoured; lda/ldb/ldc in elements, any value. */
int msf_gemm_grouped_f32(const msf_gemm_problem* problems /*host*/, int n_problems, void* stream);
/* info[0..5] = {tile_n, tiles_m, tiles_n, k_splits, k_blocks_per_split, k_blocks} the library would use for one problem. */
int msf_gemm_grouped_plan_info(const msf_gemm_problem* problem /*host*/, int32_t* info);
/* y (rows,out) = pro(x) (rows,in) W^T with W (out,in) and the column statistics of y in the epilogue -- the single-problem
 * form (SURVEY 8b `linear_bnstat`).  col_stats / a_scale / a_shift as above (may be NULL). */
int msf_linear_bnstat(const void* x, const void* w, void* y, int64_t rows, int in_features, int out_features, int dtype,
                      float* col_stats, const float* a_scale, const float* a_shift, int a_relu, void* stream);

/* ------------------------------------------------------------------------------------------
 * N2  the batch norms of the heads (src/models/backbone.py:15,18,21,28: train-mode BatchNorm1d, converted to
 * SyncBatchNorm by tools/ssl_train.py:160), grouped: every call covers ALL heads and both views of one depth of the head
 * stack in one launch.  Table entries are HOST structs holding DEVICE pointers (copied into kernel parameters).
 *
 * Forward:  msf_gemm_grouped leaves col_stats; msf_head_bn_finalize merges them (fp64), exchanges {sum, sum of squares}
 * with the other ranks INSIDE the kernel (NVLink peer loads/stores on a symmetric workspace: `peers` = DEVICE array of
 * `world` pointers to the ranks' workspaces of msf_head_sync_workspace_bytes(capacity) bytes, zeroed once; `seq` = 1, 2, 3 ...
 * identical on all ranks; world <= 1: no exchange) and writes, per view,
 *   scale = gamma * invstd, shift = beta - mean * scale   (y_norm = y * scale + shift: the next GEMM's A prologue)
 *   mean, invstd (for the backward), running_mean / running_var (momentum update with the unbiased variance; view 0 then
 *   view 1, the order in which the reference calls the module on the two views).
 * training = 0 (eval): scale / shift from the running statistics, nothing else written.
 * ---------------------------------------------------------------------------------------- */
#define MSF_HEAD_MAX_ITEMS 48      /* (head, layer) entries per finalize call */
#define MSF_HEAD_MAX_MATS 96       /* matrices per element-wise / reduce call */
#define MSF_HEAD_SYNC_MAX_CTAS 256
typedef struct {
  const float* col_stats[2]; /* per view: [ceil(rows/group_rows)][2][C] fp32 (msf_gemm_grouped epilogue: 128-row groups; msf_head_bn_stats: 32) */
  float* scale[2];           /* out [C] */
  float* shift[2];           /* out [C] */
  float* mean[2];            /* out [C] or NULL */
  float* invstd[2];          /* out [C] or NULL */
  const float* gamma;        /* [C] fp32 or NULL (affine = False) */
  const float* beta;
  float* running_mean;       /* [C] fp32, updated in place, or NULL */
  float* running_var;
  int32_t rows;              /* rows per view on this rank */
  int32_t C;
  int32_t n_views;           /* 1 or 2 */
  int32_t centered;          /* 1 (exact fp32 path): entry 1 of each col_stats group is M2 about the group mean (msf_head_bn_stats) and
                              * shift receives beta (consumers evaluate (x - mean) * scale + shift); 0: sums of squares, shift = beta - mean * scale */
  int32_t group_rows;        /* rows per col_stats group: 128 for the msf_gemm_grouped epilogue (one group per M tile), 0 or 32 for
                              * msf_head_bn_stats (the centered form needs 32) */
} msf_head_bn_item;
size_t msf_head_sync_workspace_bytes(int64_t capacity_doubles);
int msf_head_bn_finalize(const msf_head_bn_item* items /*host*/, int n_items, float eps, float momentum, int training,
                         void* const* peers /*device*/, int world, int rank, uint64_t seq, int64_t capacity_doubles,
                         int timeout_ms, void* stream);
/* col_stats [ceil(rows/32)][2][C] of x (rows, C) straight from memory, per 32-row group {sum, M2 about the group mean}
 * accumulated in fp64 (CENTERED form: set msf_head_bn_item.centered = 1): the exact fp32 path. */
typedef struct {
  const void* x;
  float* col_stats;
  int32_t rows, C;
} msf_head_mat;
int msf_head_bn_stats(const msf_head_mat* mats /*host*/, int n, int dtype, void* stream);
/* y = relu?(round_dtype(x * scale + shift)); optional y_hat = y / max(||y_row||, norm_eps) (+ inv_norm per row): the
 * projector output z and the L2-normalised InfoNCE keys in one pass; also rebuilds ReLU activations for the backward. */
typedef struct {
  const void* x;        /* (rows, C) */
  void* y;              /* (rows, C) */
  void* y_hat;          /* (rows, C) or NULL */
  float* inv_norm;      /* (rows) or NULL */
  const float* scale;   /* [C] */
  const float* shift;   /* [C] */
  const float* mean;    /* NULL, or [C]: centered form y = (x - mean) * scale + shift (shift = beta; finalize items with centered = 1) */
  int32_t rows, C, relu, reserved;
} msf_head_apply_item;
int msf_head_bn_apply(const msf_head_apply_item* items /*host*/, int n, int dtype, float norm_eps, void* stream);
/* Backward of y_norm = relu?(bn(y)) given g = dL/d(y_norm):  dy' = g * (bn(y) > 0);
 *   reduce:   partial [ceil(rows/256)][2][C] fp32 = per 256-row block {sum dy', sum dy' * xhat}, xhat = (y - mean) * invstd
 *             (y = NULL: plain column sums of g, e.g. the bias gradient of the predictor's last Linear)
 *   finalize: merges the partials of both views (fp64); d_gamma / d_beta = LOCAL sums over both views (DDP averages
 *             parameter gradients, as with SyncBatchNorm); exchanges the sums across ranks and writes c1 = mean(dy'),
 *             c2 = mean(dy' * xhat) per view (plain = 1: only d_beta, no exchange partner data)
 *   elemt:    dy = scale * (dy' - c1 - xhat * c2) */
typedef struct {
  const void* g;        /* (rows, C) */
  const void* y;        /* (rows, C) raw Linear output, or NULL (plain column sums) */
  void* dy;             /* (rows, C) out (elemt) */
  float* partial;       /* [ceil(rows/256)][2][C] (reduce out) */
  const float* scale;
  const float* shift;
  const float* mean;
  const float* invstd;
  const float* c1;      /* [C] (elemt in) */
  const float* c2;
  int32_t rows, C, relu;
  int32_t centered;     /* 1: the forward evaluated (y - mean) * scale + shift (shift = beta): the ReLU mask is rebuilt the same way */
} msf_head_bwd_item;
typedef struct {
  const float* partial[2];
  float* c1[2];
  float* c2[2];
  float* d_gamma;       /* [C] fp32 or NULL */
  float* d_beta;        /* [C] fp32 or NULL */
  int32_t rows, C, n_views, plain;
} msf_head_bwd_fin_item;
int msf_head_bn_bwd_reduce(const msf_head_bwd_item* items /*host*/, int n, int dtype, void* stream);
int msf_head_bn_bwd_finalize(const msf_head_bwd_fin_item* items /*host*/, int n_items, int training, void* const* peers /*device*/,
                             int world, int rank, uint64_t seq, int64_t capacity_doubles, int timeout_ms, void* stream);
int msf_head_bn_bwd_elemt(const msf_head_bwd_item* items /*host*/, int n, int dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * A2  feature-map crop to each tile's footprint + bilinear resample.
 * Integer case reproduces src/models/hooknet.py:29-32 (`x[:, :, 12:20, 12:20]`) bit-exactly;
 * the fractional / resampling case is an extension (F.interpolate bilinear, align_corners=False
 * semantics on the crop; oracle/msf_oracle.py:crop_resample).
 *   feat (B,C,H,W) NCHW ; boxes (B,K,4) fp32 [y0,x0,y1,x1) ; out (B,K,C,oh,ow)
 * ---------------------------------------------------------------------------------------- */
int msf_crop_resample_fwd(const void* feat, int64_t B, int C, int H, int W, const float* boxes, int K, int oh,
                          int ow, int dtype, void* out, void* stream);
/* grad_feat (B,C,H,W) fp32, overwritten.  Gather form: one thread owns each source pixel and sums the output gradients
 * that tap it in a fixed order -- no atomics, bit-reproducible run to run (also with overlapping boxes). */
int msf_crop_resample_bwd(const void* grad_out, int64_t B, int C, int H, int W, const float* boxes, int K, int oh,
                          int ow, int dtype, float* grad_feat, void* stream);

/* ------------------------------------------------------------------------------------------
 * E1  multi-tensor EMA  teacher <- m * teacher + (1 - m) * student   (extension: the reference
 * keeps no momentum encoder, src/models/backbone.py:58-65; sized to its 60-tensor ResNet-18).
 * `entries` and `chunk_prefix` are DEVICE arrays built once per parameter list:
 * chunk_prefix[t] = number of MSF_EMA_CHUNK-element chunks in tensors [0, t), length n+1.
 * msf_ema_plan fills a HOST chunk_prefix from HOST numels.
 * ---------------------------------------------------------------------------------------- */
#define MSF_EMA_CHUNK 8192
typedef struct {
  void* teacher;
  const void* student;
  int64_t numel;
} msf_ema_entry;
int msf_ema_plan(const int64_t* numels /*host*/, int n_tensors, int32_t* chunk_prefix /*host, n+1*/);
/* `one_minus_momentum` is passed separately so the caller controls its rounding (torch evaluates 1-m in
 * double before narrowing); the kernel computes fma(one_minus_momentum, student, rn(momentum*teacher)). */
int msf_ema_multi(const msf_ema_entry* entries /*device*/, const int32_t* chunk_prefix /*device*/, int n_tensors,
                  int total_chunks, int teacher_dtype, int student_dtype, float momentum, float one_minus_momentum,
                  void* stream);

/* ------------------------------------------------------------------------------------------
 * N1  train-mode batch normalisation over [rows][C] matrices with the element-wise work around it fused in:
 * the BatchNorm1d + ReLU of the projector / predictor heads (src/models/backbone.py:15-16,18-19,21,28-29; rows = batch rows)
 * and the channels-last BatchNorm2d of the encoders (caller side of the hot path: src/models/resnet.py:59-82 BasicBlock
 * `bn -> relu`, `bn -> (+identity) -> relu`, and the stem `bn1 -> relu -> maxpool` of src/models/resnet.py:244-247;
 * rows = N*H*W).  Statistics follow torch.nn.BatchNorm{1,2}d / SyncBatchNorm (tools/ssl_train.py:160): one sum-reducible
 * fp64 vector per layer and direction instead of SyncBatchNorm's all_gather + per-rank recombination.  The convolutions
 * and Linears stay on cuDNN / the tcgen05 GEMM.
 * The activation is a [rows = N*H*W][C] matrix, C contiguous (NHWC), C a multiple of 8 (16-bit) or 4 (fp32).
 *
 * forward : msf_bn2d_stats -> (all-reduce `sums` over ranks for SyncBN) -> msf_bn2d_finalize -> msf_bn2d_apply[_pool]
 * backward: msf_bn2d_bwd_reduce -> (all-reduce) -> msf_bn2d[_pool]_bwd_elemt
 *   sums (forward)  : 2C+1 doubles {sum x, sum x^2 per channel, element count} -- sum-reducible over ranks
 *   sums (backward) : 2C doubles {sum dy', sum dy' * xhat}; grad_beta = sums[0:C], grad_gamma = sums[C:2C] (local part)
 *   dy' = dy * (y > 0) when relu != 0; y is recomputed from x, or read from `y_mask` (the saved output) when the
 *   forward added a residual.  gamma / beta may be NULL (1 / 0).
 * ---------------------------------------------------------------------------------------- */
size_t msf_bn2d_workspace_bytes(int64_t rows, int C);
int msf_bn2d_stats(const void* x, int64_t rows, int C, int dtype, double* sums_out /*2C+1*/, void* workspace,
                   size_t workspace_bytes, void* stream);
/* Single-process forward: msf_bn2d_stats and msf_bn2d_finalize in two launches instead of three (no cross-rank reduction
 * in between); sums_out still receives the 2C+1 doubles the backward needs. */
int msf_bn2d_stats_finalize(const void* x, int64_t rows, int C, int dtype, float eps, float momentum, double* sums_out /*2C+1*/,
                            float* mean, float* invstd, float* running_mean, float* running_var, void* workspace,
                            size_t workspace_bytes, void* stream);
/* mean / invstd (fp32, C each) from sums; running_mean / running_var (both or neither) get the momentum update with
 * the unbiased variance, like torch.nn.BatchNorm2d. */
int msf_bn2d_finalize(const double* sums /*2C+1*/, int C, float eps, float momentum, float* mean, float* invstd,
                      float* running_mean, float* running_var, void* stream);
/* y = act(gamma * (x - mean) * invstd + beta (+ res)); res may be NULL; relu != 0 applies max(., 0).
 * relu_bits (may be NULL): rows*C/vec bytes, one per 16-byte chunk (vec = 8 elements for 16-bit dtypes, 4 for fp32), bit i
 * = (pre-ReLU value of element i > 0): the ReLU mask for the backward of the residual form, 16x smaller than the output. */
int msf_bn2d_apply(const void* x, const void* res, void* y, uint8_t* relu_bits, int64_t rows, int C, int dtype,
                   const float* mean, const float* invstd, const float* gamma, const float* beta, int relu, void* stream);
/* gpool (may be NULL): gradient of the global average pool of the same output y (src/models/resnet.py:250-254 pools the
 * layer outputs), (N, C) in the activation dtype with hw = H*W rows per image: dy of every row of image n is taken as
 * dy + gpool[n, :] / hw, i.e. the pooled branch's gradient is folded in instead of being expanded and added by ATen.
 * Only with relu != 0 and y_mask given (the BasicBlock output); rows*C/vec must be < 2^32. */
/* y_mask: NULL (mask recomputed from x), the saved output y (mask_is_bits = 0), or the relu_bits of msf_bn2d_apply
 * (mask_is_bits = 1). */
int msf_bn2d_bwd_reduce(const void* x, const void* dy, const void* y_mask, int mask_is_bits, int64_t rows, int C, int dtype,
                        const float* mean, const float* invstd, const float* gamma, const float* beta, int relu,
                        const void* gpool, int64_t hw, double* sums_out /*2C*/, void* workspace, size_t workspace_bytes,
                        void* stream);
/* dx = gamma*invstd * (dy' - sums[c]/count - xhat * sums[C+c]/count); dres (may be NULL) = dy'.
 * `count` is a DEVICE double (element 2C of the forward sums). */
int msf_bn2d_bwd_elemt(const void* x, const void* dy, const void* y_mask, int mask_is_bits, void* dx, void* dres, int64_t rows,
                       int C, int dtype, const float* mean, const float* invstd, const float* gamma, const float* beta,
                       int relu, const void* gpool, int64_t hw, const double* sums /*2C*/, const double* count,
                       void* stream);
/* Stem: y (N,PH,PW,C) = maxpool3x3/stride2/pad1(relu(bn(x))), x (N,H,W,C), PH = (H-1)/2+1, PW = (W-1)/2+1, N*PH*PW < 2^31.
 * tap (N,PH,PW,C) uint8 = arg-max tap dr*3+dc of each output (first maximum in scan order, ATen's max_pool2d rule),
 * 255 where the output is 0 (no gradient); x_arg (N,PH,PW,C) = the x value at the arg-max.
 * Backward: dy' lives on the pooled grid, so the reduction is msf_bn2d_bwd_reduce(x_arg, dpool, y_mask = y, mask_is_bits = 0,
 * rows = N*PH*PW, relu = 1) -- a streaming pass over pooled-size tensors -- followed by msf_bn2d_pool_bwd_elemt, which scatters the pooled
 * gradient to the arg-max positions and applies the batch-norm input gradient in one pass over x. */
int msf_bn2d_apply_pool(const void* x, void* y, uint8_t* tap, void* x_arg, int64_t N, int H, int W, int C, int dtype,
                        const float* mean, const float* invstd, const float* gamma, const float* beta, void* stream);
int msf_bn2d_pool_bwd_elemt(const void* x, const void* dpool, const uint8_t* tap, void* dx, int64_t N, int H, int W,
                            int C, int dtype, const float* mean, const float* invstd, const float* gamma,
                            const double* sums /*2C*/, const double* count, void* stream);

/* ------------------------------------------------------------------------------------------
 * O1  multi-tensor Adam step with the GradScaler work and an optional EMA teacher update folded in.
 * Replaces `optimizer.step()` of tools/ssl_train.py:303-309 (torch.optim.Adam over the three learning-rate groups
 * context_/target_/inter_) and `scaler.step(optimizer)` of :472-474 (unscale, non-finite check, skipped step).
 *   g = grad * *inv_scale (+ weight_decay * p);  m = m + (1-beta1)*(g - m);  v = beta2*v + (1-beta2)*g*g;
 *   p -= lr[group] / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps);   t = *step (already incremented)
 *   teacher = ema_momentum * teacher + (1 - ema_momentum) * p   (entries with ema != NULL, when with_ema != 0)
 *   shadow  = (bf16 | fp16)(p)   (entries with shadow != NULL): the 16-bit operand copy the head GEMMs read, so no
 *             separate fp32 -> bf16 cast pass runs per step and the copy can never go stale behind a raw-pointer update
 * Nothing is written when *found_inf != 0.  params / exp_avg / exp_avg_sq / ema are fp32; grads fp32, bf16 or fp16.
 * `entries` and `chunk_prefix` are DEVICE arrays (msf_adam_plan fills a HOST prefix from HOST numels; chunks of
 * MSF_ADAM_CHUNK elements); lr is a DEVICE array indexed by entry.group; step / inv_scale / found_inf are DEVICE scalars
 * (inv_scale and found_inf may be NULL), so the whole optimizer step is free of host synchronisation.  The scalar
 * hyper-parameters are doubles because torch forms 1 - beta in double before narrowing (1.f - 0.999f != float(0.001)).
 * ---------------------------------------------------------------------------------------- */
#define MSF_ADAM_CHUNK 4096
typedef struct {
  void* param;
  const void* grad;
  void* exp_avg;
  void* exp_avg_sq;
  void* ema;      /* teacher copy of this parameter or NULL */
  int64_t numel;
  int32_t group;  /* index into lr[] */
  int32_t shadow_dtype; /* MSF_BF16 / MSF_F16 when `shadow` is set, else 0 */
  void* shadow;   /* 16-bit copy of this parameter (the GEMM operand of the head Linears) rewritten in the same pass, or NULL */
} msf_adam_entry;
int msf_adam_plan(const int64_t* numels /*host*/, int n_tensors, int32_t* chunk_prefix /*host, n+1*/);
/* *found_inf = 1 if any gradient element is NaN/Inf (never cleared here: the caller zeroes it once per step). */
int msf_grad_check_multi(const msf_adam_entry* entries /*device*/, const int32_t* chunk_prefix /*device*/, int n_tensors,
                         int total_chunks, int grad_dtype, float* found_inf, void* stream);
int msf_adam_multi(const msf_adam_entry* entries /*device*/, const int32_t* chunk_prefix /*device*/, int n_tensors,
                   int total_chunks, int grad_dtype, const float* lr, double beta1, double beta2, double eps,
                   double weight_decay, const float* step, const float* inv_scale, const float* found_inf, int with_ema,
                   float ema_momentum, float ema_one_minus_momentum, void* stream);

/* ------------------------------------------------------------------------------------------
 * S1  space-to-depth re-layout of the encoder input for the stem convolution (src/models/resnet.py:155, 244:
 * Conv2d(3, 64, 7, stride 2, padding 3)).  conv7x7/2(x, w) == conv4x4/1(s2d(x), w') with
 *   s2d(x)[n, oy, ox, c*4 + dy*2 + dx] = x[n, c, 2*oy + dy - 3, 2*ox + dx - 3]   (0 outside; channels >= 4*C_in are 0)
 * out (N, (H+6)/2, (W+6)/2, 16) NHWC in out_dtype; x addressed through ELEMENT strides (NCHW or NHWC), C_in <= 4, H and W
 * even.  The convolution itself stays on cuDNN (3x faster in this form); w' is the 7x7 kernel zero-padded to 8x8 and
 * rearranged the same way (a tiny differentiable torch expression on the host side, msfwsi_b200/resnet.py).
 * ---------------------------------------------------------------------------------------- */
int msf_stem_s2d(const void* x, int64_t N, int C_in, int H, int W, int64_t stride_n, int64_t stride_c, int64_t stride_y,
                 int64_t stride_x, int in_dtype, void* out, int out_dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * C1  all-reduce (sum) of a small fp64 vector over NVLink peer memory, one single-CTA kernel per call: the cross-rank
 * exchange of the batch-norm statistics (SyncBatchNorm, tools/ssl_train.py:160; 2C+1 doubles per layer and direction)
 * without an NCCL launch.  `peers` is a DEVICE array of `world` pointers to the ranks' SYMMETRIC workspaces of
 * msf_peer_workspace_bytes(capacity) bytes each, every one mapped into this process (peers[rank] is the local one), zeroed
 * once before the first call.  All ranks must issue the same sequence of calls with seq = 1, 2, 3, ... (like any
 * collective).  Result: vec[i] = sum over ranks, added in rank order (bit-identical on every rank).  The wait for the
 * peers is bounded by timeout_ms; a timeout traps, i.e. surfaces as a CUDA error instead of a hang.
 * ---------------------------------------------------------------------------------------- */
#define MSF_PEER_MAX_WORLD 32
size_t msf_peer_workspace_bytes(int64_t capacity_doubles);
int msf_peer_allreduce_f64(double* vec, int n, void* const* peers /*device*/, int world, int rank, uint64_t seq,
                           int64_t capacity_doubles, int timeout_ms, void* stream);

/* ------------------------------------------------------------------------------------------
 * D1  on-device data path of the views: blockshaped tiling + jigsaw shuffle + resize + normalise from the uint8 source
 * image in one kernel.  Replaces, per sample, src/utils/data/bcss.py:171-177 `blockshaped(img, 256, 256)[jigsaw_idx]`
 * and the per-tile Resize / Normalize / ToTensor of transforms[2] (the random augmentations are not reproduced).
 *   src (B, H, W, 3) uint8 HWC; perm (B, grid*grid) int64 = jigsaw_idx (the FORWARD shuffle; NULL = identity);
 *   out (B*grid*grid, oh, ow, 3) NHWC in out_dtype: tile j of sample b = source tile perm[b, j] (raster order of
 *   blockshaped), bilinearly resampled to oh x ow (align_corners = False on the cropped tile) and normalised with
 *   (v - 255*mean[c]) / (255*std[c]).  mean3 / std3 are HOST arrays.  grid = 1 yields the context view.
 * perm values outside [-K, K) set bit 0 of *status_flag (device, may be NULL) and are clamped.
 * ---------------------------------------------------------------------------------------- */
int msf_jigsaw_tiles(const uint8_t* src, int64_t B, int H, int W, int grid, const int64_t* perm, int oh, int ow,
                     const float* mean3 /*host*/, const float* std3 /*host*/, void* out, int out_dtype,
                     int32_t* status_flag, void* stream);

/* D1b  every view of a step from the uint8 source tiles in ONE launch, written in the stem convolution's input layout
 * (msf_stem_s2d's output: (n, (oh+6)/2, (ow+6)/2, 16) NHWC, zero padding and zero channels included).
 *   view i = normalise(hflip?(resize_bilinear(src[sample][y0:y1, x0:x1] -> oh x ow)))
 * = the reference's geometric augmentations once their random numbers are drawn (albumentations RandomResizedCrop -> an
 * integer crop box, HorizontalFlip, Normalize; tools/ssl_train.py:175-217) on the whole tile (context views) or on tile
 * jigsaw_idx[j] of blockshaped(img, 256, 256) (target views, src/utils/data/bcss.py:171-177: box = tile origin + crop).
 * crops is a DEVICE array; boxes are integer source-pixel coordinates [y0, y1) x [x0, x1) inside the (H, W) image;
 * out-of-range entries set bit 0 of *status_flag and are clamped.  oh and ow must be even. */
typedef struct {
  int32_t sample;         /* index into src's first dimension */
  int32_t y0, x0, y1, x1;
  int32_t flip;           /* 1: horizontal flip of the resized crop */
} msf_view_crop;
int msf_view_crops_s2d(const uint8_t* src, int64_t B, int H, int W, const msf_view_crop* crops /*device*/, int64_t n_views,
                       int oh, int ow, const float* mean3 /*host*/, const float* std3 /*host*/, void* out, int out_dtype,
                       int32_t* status_flag, void* stream);

/* ------------------------------------------------------------------------------------------
 * Measurement hook (bench.py): when switched on, every compute entry point records a CUDA event pair on its stream
 * immediately around its main kernel(s); msf_prof_end synchronises those events and returns, per kernel id, the number
 * of calls, the summed algorithmic work (bytes for HBM-bound kernels, FLOP for tensor-bound ones, as defined in
 * DESIGN.md section 4) and the summed device time.  Off by default; the only global state of the library.
 * ---------------------------------------------------------------------------------------- */
typedef enum {
  MSF_K_GATHER_FWD = 0, MSF_K_GATHER_BWD, MSF_K_COS_FWD, MSF_K_COS_BWD, MSF_K_ROWNORM, MSF_K_NCE_FLASH, MSF_K_NCE_TWOPASS,
  MSF_K_NCE_SIMT, MSF_K_NCE_BWD, MSF_K_GEMM, MSF_K_CROP_FWD, MSF_K_CROP_BWD, MSF_K_EMA, MSF_K_BN_STATS, MSF_K_BN_APPLY,
  MSF_K_BN_APPLY_RES, MSF_K_BN_BWD_REDUCE, MSF_K_BN_BWD_ELEMT, MSF_K_BN_APPLY_POOL, MSF_K_BN_POOL_BWD_ELEMT,
  MSF_K_ADAM, MSF_K_GRAD_CHECK, MSF_K_STEM_S2D, MSF_K_PEER_ALLREDUCE, MSF_K_JIGSAW_TILES, MSF_K_GEMM_GROUPED, MSF_K_HEAD_BN_FINALIZE,
  MSF_K_HEAD_BN_ELEMWISE, MSF_K_GEMM_F32, MSF_K_NCE_DK, MSF_K_COUNT
} msf_kernel_id;
typedef struct {
  int32_t kernel;   /* msf_kernel_id */
  int32_t launches; /* profiled calls */
  double work;      /* summed algorithmic bytes or FLOP */
  double ms;        /* summed device time between the event pairs */
} msf_prof_record;
int msf_prof_begin(int capacity /* event pairs to pre-create; calls beyond it are dropped and counted */);
int msf_prof_end(msf_prof_record* out /* host, MSF_K_COUNT entries */, int* dropped /* host, may be NULL */);
const char* msf_prof_kernel_name(int kernel);
int msf_prof_kernel_bound(int kernel); /* 'h' HBM, 't' tensor pipe, 'f' fp32 FMA, 'l' latency (NVLink round trip) */

#ifdef __cplusplus
}
#endif
#endif /* MSFWSI_B200_H_ */
