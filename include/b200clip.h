/*
 * b200clip — C ABI of the B200-native contrastive head (libb200clip.so).
 *
 * Drop-in boundary for the ONE hot path of HeartWise-AI/DeepCORO_CLIP named in BASELINE.json:
 * the CLIP / SigLIP losses (utils/loss), the streaming retrieval metrics
 * (utils/retrieval_metrics_streaming.py), Rope3D (models/rope_3d.py), AttentionPool
 * (models/attention_pool.py) and the multi-view query pool (models/video_aggregator.py:131-158); widened to the
 * dense multi-label retrieval metrics (utils/retrieval_metrics.py).
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer unless its name ends in `_host`.
 *  - `stream` is a cudaStream_t passed as void*. No entry point synchronises, allocates or touches host
 *    memory; the caller owns every buffer (workspace sizes are documented per call).
 *  - Return value: 0 on success, negative errno-style code otherwise (b200clip_strerror()).
 *  - dtype codes: 0 = float32, 1 = bfloat16, 2 = float16.
 *  - "operand" buffers are the bf16 L2-normalised MMA operands produced by b200clip_l2norm_fwd():
 *    row-major [rows, ld] with the first Kp (= dim rounded up to 64) columns valid and zero padded; in
 *    bf16x3 mode three such panels are K-concatenated (ld >= 3*Kp).
 *
 * There is no CPU fallback anywhere behind this header: on a machine without an sm_100a GPU the calls
 * fail with a CUDA error code.
 */
#ifndef B200CLIP_H_
#define B200CLIP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200CLIP_ABI_VERSION 1

int b200clip_abi_version(void);
const char* b200clip_strerror(int code);
/* Number of SMs of the current device (grid sizing / tests). */
int b200clip_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * K1  L2 normalise + operand packing.   Replaces F.normalize(x.float(), dim=-1)
 *     (reference utils/loss/contrastive.py:146-147, 259-260; utils/loss/losses.py:46-47, 138-139, 191-192;
 *      utils/retrieval_metrics_streaming.py:130-131).
 *   x [rows, dim] (row pitch ldx elements, dtype code) -> operand [rows, ld_out] bf16, inv_norm [rows] fp32
 *   (= 1 / max(||x||, 1e-12)), optional xhat_f32 [rows, ld_hat] fp32 (may be NULL).
 *   normalize = 0 packs the raw features instead (compute_recall_at_k_streaming does not normalise:
 *   retrieval_metrics_streaming.py:35-41) and writes ||x|| to inv_norm.
 *   split3_role: -1 plain bf16 operand; 0 / 1 = A-side [lo|hi|hi] / B-side [hi|lo|hi] panels of the bf16x3
 *   compensated product (small terms first: the tensor core truncates its fp32 accumulator each K step).
 * ------------------------------------------------------------------------------------------------ */
int b200clip_l2norm_fwd(const void* x, int dtype, int64_t ldx, int rows, int dim, void* operand, int ld_out,
                        int Kp, int split3_role, float* inv_norm, float* xhat_f32, int ld_hat, int normalize,
                        void* stream);

/* K1m Normalise + all-gather in ONE kernel (multi-GPU row-slab path; replaces F.normalize followed by gather_with_gradient's
 *     all_gather, utils/loss/contrastive.py:75-101, 142-147): every bf16 operand row r is stored at row row_offset + r of
 *     EACH of the n_operands [N, ld_out] buffers listed in the HOST array operands_host — the same symmetric-memory
 *     allocation on every rank of the job, remote ones written with plain stores over NVLink / NVSwitch. The caller runs a
 *     cross-rank barrier after the launch. Plain bf16 operands only; dim % 8 == 0, Kp <= 1024, 16-byte aligned rows;
 *     n_operands <= 8. */
int b200clip_l2norm_fwd_multi(const void* x, int dtype, int64_t ldx, int rows, int dim, void* const* operands_host,
                              int n_operands, int64_t row_offset, int ld_out, int Kp, float* inv_norm, int normalize,
                              void* stream);

/* l2norm_fwd_multi through the MULTICAST mapping of the symmetric operand buffer (NVLink SHARP / NVLS: one multimem.st per
 * 16 bytes, replicated by the switch into the buffer of every rank of the group, the caller's included): mc_operand = the
 * multicast address of the [N, ld_out] operand buffer. Same constraints and the same cross-rank barrier afterwards. */
int b200clip_l2norm_fwd_mc(const void* x, int dtype, int64_t ldx, int rows, int dim, void* mc_operand, int64_t row_offset,
                           int ld_out, int Kp, float* inv_norm, int normalize, void* stream);

/* K4  normalise backward (autograd of F.normalize) fused with the rank-sparse gradient corrections:
 *   g  = gmul * ( gscale * dxhat[r] + omul * (res_r * yh_r + gb_r * (yh_r - yhi_r)) + (ucoef * omul) * usum )
 *   dx = (g - (g . xhat) xhat) * inv_norm,   xhat = x * inv_norm (recomputed in fp32 from the caller's input)
 *   yh_r = other_x[r] * other_inv_norm[r] is the exact fp32 partner (target) row, yhi_r = other_hi[r] its bf16 hi
 *   panel, diag_corr[r] = {res_r, gb_r} as written by b200clip_logits_bwd: the tensor-core product contributed
 *   bf16(g_rr) * yhi_r for the target pair, this restores g_rr * yh_r in fp32. usum: column sum of the partner
 *   operand (label smoothing). other_x / diag_corr / usum may be NULL. dev_omul / dev_gmul: optional DEVICE scalars
 *   (1/tau inside dyn, upstream grad_output). dx fp32 [rows, lddx]. */
int b200clip_l2norm_bwd(const float* dxhat, int ldg, const void* x, int dtype, int64_t ldx, const float* inv_norm,
                        const void* other_x, int other_dtype, int64_t ld_other_x, const float* other_inv_norm,
                        const void* other_hi, int ld_other_hi, const float* diag_corr, const float* usum, float gscale,
                        float ucoef, const float* dev_omul, const float* dev_gmul, int rows, int dim, float* dx,
                        int64_t lddx, void* stream);

/* out[c] += sum_r operand[r, c]  (label-smoothing helper; out must be zeroed by the caller) */
int b200clip_colsum_bf16(const void* operand, int ld, int rows, int dim, float* out, void* stream);
/* out[r] = (a[r, :dim] . b[r, :dim]) * a_inv_norm[r] * b_inv_norm[r]: cosine of pair r from the RAW features in fp32
 * (the target logit S_ii of utils/loss/contrastive.py:150-162 at fp32 accuracy; dtype codes as in l2norm_fwd). */
int b200clip_rowdot_raw(const void* a, int a_dtype, int64_t lda, const float* a_inv_norm, const void* b, int b_dtype,
                        int64_t ldb, const float* b_inv_norm, int rows, int dim, float* out, void* stream);
/* out[r] = a[r, :K] . b[idx ? idx[r] : r, :K]  (diagonal / ground-truth logits; idx int64 or NULL) */
int b200clip_rowdot_bf16(const void* a, int lda, const void* b, int ldb, const int64_t* idx, int rows,
                         int b_rows, int K, float* out, void* stream);

/* dst[r, :K] = src[idx[r], :K]  (bf16 operand rows; K, lds, ldd multiples of 8; out-of-range idx -> zero row) */
int b200clip_gather_rows_bf16(const void* src, int lds, const int64_t* idx, int rows, int src_rows, int K, void* dst,
                              int ldd, void* stream);
/* out[r] = a[r, :Kp] . b[r, :Kp] computed by the tcgen05 tile engine (diagonal tiles only): BIT-IDENTICAL to the value
 * every similarity tile of retrieval_sweep / logits_* produces for that pair, which a CUDA-core dot product is not.
 * The ground-truth similarity of the rank counts (retrieval_metrics_streaming.py:151-166 compares s_ij with s_i,gt)
 * must come from here so that exact duplicates of the ground-truth text tie exactly. */
int b200clip_rowdot_tc(const void* a, int lda, const void* b, int ldb, int rows, int Kp, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2  Fused logits forward for the softmax-CE losses. Replaces matmul -> /temp -> 2x cross_entropy
 *     (utils/loss/contrastive.py:150-162; losses.py:50-62, 143-156; gated: losses.py:195-210, 258-274).
 *   A [Ma, >=Kp], B [Nb, >=Kp] operands. P_ij = 2^(f(S_ij) * scale2 - shift2), S = A B^T,
 *   f(s) = s (gated = 0) or s * sigmoid(s) (gated = 1).
 *   rowsum[i] += sum_j P_ij, colsum[j] += sum_i P_ij (fp32, caller zeroes them). S is never stored.
 *   dyn (may be NULL): device float[16] written by b200clip_dyn_prep; when given, scale2/shift2 are read from
 *   it on the device, so a learnable temperature never forces a host synchronisation.
 *   diag (may be NULL): diag[i] = S[i, i + diag_off] as produced by the tensor core (the target logit of row i;
 *   diag_off = rank * B_local under DDP), so target and LSE share one rounding and cancel in the loss.
 *   skip_if_stable != 0 (with dyn): the launch returns at once when dyn[11] != 0, i.e. when the temperature left the
 *   window the fixed shift covers and b200clip_logits_rowlse computes the statistics instead (both are enqueued; the
 *   choice is made on the device, the host never reads tau).
 * ------------------------------------------------------------------------------------------------ */
int b200clip_logits_lse_fwd(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, float scale2,
                            float shift2, int gated, const float* dyn, int skip_if_stable, float* rowsum, float* colsum,
                            float* diag, int diag_off, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2s Stable row log-sum-exp: lse2[i] = log2 sum_j 2^(f(S_ij) * log2(e) / tau) with a running per-row maximum (online
 *     softmax, tile order outer = A tile so the epilogue thread owns its row), for temperatures down to the reference's
 *     clamp floor tau = 1e-4 (utils/loss/contrastive.py:153: logits up to +-1e4) and the unclamped legacy classes
 *     (utils/loss/losses.py:53, 146), where log_softmax's own max subtraction keeps the reference finite. The column
 *     statistics of the symmetric loss are the row statistics of the role-swapped call (A = text, B = video).
 *   part  : float[Ma][slots][2] scratch, slots = b200clip_rowlse_slots(Ma, Nb, Kp); ticket: int32[Ma], ZERO on entry
 *           (left zero on exit); the last partial of a row merges them, so lse2 is complete when the launch ends.
 *   only_if_stable != 0: the launch returns at once unless dyn[11] != 0 (see b200clip_dyn_prep).
 *   diag (may be NULL): diag[i] = S[i, i + diag_off] with the tensor core's rounding, as in logits_lse_fwd.
 *   gap (may be NULL; needs diag): gap[i] = lse2[i] - L2_ii formed as (max - L2_ii) + log2(sum) — no cancellation between
 *           two numbers of the size of the logits; this is the loss term of row i in the log2 domain.
 * ------------------------------------------------------------------------------------------------ */
int b200clip_rowlse_slots(int Ma, int Nb, int Kp);
int b200clip_logits_rowlse(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, int gated,
                           const float* dyn, int only_if_stable, float* part, int slots, int32_t* ticket, float* lse2,
                           float* diag, int diag_off, float* gap, void* stream);

/* Validation hook: out[i, j] = S_ij (fp32, row pitch ldo) computed by the same tcgen05 tile engine.
 * max_ctas > 0 limits the grid (exercises the multi-tile-per-CTA schedule). Used by the tests only. */
int b200clip_logits_dump(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, float* out,
                         int ldo, int max_ctas, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K3  Fused logits backward (tile recompute). Replaces autograd through matmul / cross_entropy /
 *     binary_cross_entropy_with_logits (utils/loss/contrastive.py:150-162, 263-303; losses.py:195-210).
 *       dX[i, :D] += out_scale * sum_j G_ij * Y[j, :D],   S = X Y^T recomputed per 128x128 tile.
 *     mode 0 (CLIP)   : G = 2^(S*scale2 - shift2) * (rowscale[i] + colscale[j])
 *     mode 1 (gated)  : same with f(S) = S*sigmoid(S) inside the exponent and G *= f'(S)
 *     modes 0/1 in the stable mode (dyn given and dyn[11] != 0, see b200clip_dyn_prep / b200clip_clip_finalize):
 *                       G = 2^(f(S)*scale2 - rowscale[i]) + 2^(f(S)*scale2 - colscale[j]) — rowscale / colscale then
 *                       hold log2-domain log-sum-exps, each exponential is a softmax probability times c <= c
 *                       (utils/loss/contrastive.py:153-162 at tau down to 1e-4; the choice is made on the device)
 *     mode 2 (SigLIP) : R = S*inv_tau + bias, G = wneg_c * (sigmoid(clamp(R,+-lc)) - yneg) * [|R| <= lc]; lc = dyn[8]
 *                       (30), yneg = dyn[9] (label-smoothing target of the non-positive pairs, default 0)
 *     mode 3 (SigLIP + entropy regulariser, utils/loss/contrastive.py:19-68, 306-313): mode 2 plus
 *                       G_ij += dyn[10] * p_ij (h_ij - m_v) [|R| <= lc], p_ij = exp(L_ij - 30) / Z_v,
 *                       h = -ln(p + 1e-10) - p / (p + 1e-10); v = the video of the pair. The pairs {1/Z_v, m_v}
 *                       (b200clip_siglip_entropy_rows) are passed as rowscale [Nx][2] when X holds the videos or
 *                       as colscale [Ny][2] when Y does (exactly one of the two non-NULL); dyn is required.
 *     modes 0/1: G_ij -= ydiag where i + diag_off == j (the (1-eps)/N diagonal target, subtracted in fp32 before G
 *     is rounded to bf16; diag_off = rank * B_local under DDP); diag_corr (may be NULL) receives per row
 *     {g_ii - rounded(g_ii), rounded(g_ii)} for b200clip_l2norm_bwd. SigLIP positives are rank-sparse corrections applied
 *     by b200clip_siglip_pos.
 *   X [Nx, >=Kp], Y [Ny, >=Kp] operands; Dp = padded width of the hi panel (columns Y[:, hi_off:hi_off+Dp] feed
 *   the output product; hi_off = 0 for plain bf16 operands, 2*Dp in bf16x3 mode), D = valid output columns; dX fp32 [Nx, ldd] accumulated atomically (caller zeroes).
 *   scal (may be NULL): [0] += sum G*f(S)  [1] += sum softplus(L) (mode 2)  [2] += sum G (mode 2).
 *   gnorm: G is formed, rounded to bf16 and fed to the tensor core as G*gnorm (choose gnorm so that G*gnorm = O(1):
 *   2N for CLIP, 1/wneg_c for SigLIP); sums and dX are scaled back.
 *   hp = 1: G is split into bf16 hi + lo and the output product issues two TS-MMAs per K step (rounding error of
 *   the gradient operand 2^-17 instead of 2^-9); meant for the small, latency-bound problems that also use bf16x3.
 *   nseg_hint <= 0 lets the library pick the split of the Y sweep.
 * ------------------------------------------------------------------------------------------------ */
int b200clip_logits_bwd(int mode, const void* X, const void* Y, int Nx, int Ny, int Kp, int Dp, int D, int hi_off,
                        int ldx, int ldy, float scale2, float shift2, float inv_tau, float bias, float wneg_c,
                        const float* rowscale, const float* colscale, float out_scale, float gnorm, int hp,
                        const float* dyn, float ydiag, int diag_off, float* diag_corr, float* dX, int ldd, double* scal, int nseg_hint,
                        void* stream);

/* K3b  Both gradients of the softmax (mode 0) / gated (mode 1) / sigmoid (mode 2, rowscale = colscale = NULL) contrastive step
 *      from ONE recompute of the logits
 *      (utils/loss/contrastive.py:150-164 backward: dV̂ = G T̂ / tau, dT̂ = G^T V̂ / tau). logits_bwd as above for dX with every
 *      G tile also stored through TMA (bf16, scaled by gnorm), then dY[Ny, D] += dyn[2] / gnorm * G^T X as a tcgen05 product
 *      with both operands MN-major (csrc/gt_gemm.cu). Executed work 8 instead of 10 Nx Ny D per step.
 *      G: caller-owned bf16 buffer of gstore_elems(Nx, Ny) elements, 128-byte aligned, in the BLOCKED layout the two kernels
 *      share: [64 x 64] blocks of 8 KB, block (ib, jb) = rows 64 ib .., columns 64 jb .. at element offset
 *      (ib * nJB + jb) * 4096, nJB = 4 ceil(Ny / 256), ib < 2 ceil(Nx / 128); inside a block row r holds its 64 columns as eight
 *      16-byte units, unit u at position u ^ (r & 7) (the SWIZZLE_128B operand image); blocks past the edges hold zeros.
 *      Plain bf16 operands, Kp == Dp in {256, 512, 768}, dyn required; B2_ENOSYS otherwise (the caller then launches logits_bwd
 *      twice), B2_ENOMEM if g_elems is too small. dX and dY are ACCUMULATED (caller zeroes); diag_corr of the Y side equals
 *      the one written here when the problem is square with diag_off = 0.
 *      gt_gemm alone: the product for any G in that layout and bf16 X [Nx, ldx] of padded width Dp in {256, 512, 768} (the
 *      accumulator of a CTA holds 512 columns: Dp = 768 sweeps G twice). */
int b200clip_gstore_elems(int Nx, int Ny, int64_t* elems);
int b200clip_gt_gemm(const void* G, int64_t g_elems, int Nx, int Ny, const void* X, int ldx, int Dp, int D, const float* dyn,
                     float gnorm, float* dY, int ldd, void* stream);
int b200clip_logits_bwd_both(int mode, const void* X, const void* Y, int Nx, int Ny, int Kp, int Dp, int D, int ldx, int ldy,
                             float wneg_c, const float* rowscale, const float* colscale, float gnorm, const float* dyn,
                             float ydiag, int diag_off, float* diag_corr, float* dX, int ldd, float* dY, int lddy,
                             double* scal, void* G, int64_t g_elems, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Device-side scalar plumbing (no host sync on log_temp / bias).
 *   dyn_prep     : tau = exp(log_temp) [clamped at clamp_min if > 0: contrastive.py:153, 266]; bound = max of
 *                  f(S) (1 for plain, 0.7311 for gated). dyn = {log2e/tau, shift2, 1/tau, tau, clamped, bias,
 *                  ln2*shift2, 1-clamped, logit clamp = 30, negative target = 0, entropy coefficient = 0,
 *                  stable softmax mode = [2 * bound * log2e / tau > 224 bits], ...} (float[16]).
 *   dyn_set_siglip : overrides dyn[8] (logit clamp; 3e38 = the SigLIP2 BCE variants that do not clamp,
 *                  utils/loss/siglip2_bce.py:88-90) and dyn[9] (label smoothing eps/2, siglip2_bce.py:98-99).
 *   lse_finalize : acc[0] += sum_r (ln sums[r] + ln2*shift2)  (double) ; scale_out[r] = c / sums[r]
 *   diag_sum     : acc[0] += sum_r f(a[r,:K] . b[r,:K]) (double), f = identity / s*sigmoid(s); optional dots[r].
 * ------------------------------------------------------------------------------------------------ */
int b200clip_dyn_prep(const float* log_temp, const float* bias, float clamp_min, float bound, float* dyn,
                      void* stream);
int b200clip_dyn_set_siglip(float* dyn, float logit_clamp, float neg_target, void* stream);
/* Overrides the softmax mode dyn_prep chose from tau (dyn[11]): stable = 1 is valid for every tau, stable = 0 only inside
 * the fixed-shift window. For A/B measurements and the parity tests of the two modes against each other. */
int b200clip_dyn_set_stable(float* dyn, int stable, void* stream);
int b200clip_lse_finalize(const float* sums, int n, const float* dyn, float c, float* scale_out, double* acc,
                          void* stream);
/* acc[0] += sum_i f(v[i]) in double, f = identity (gated = 0) or s*sigmoid(s) (gated = 1) */
int b200clip_vec_fsum(const float* v, int n, int gated, double* acc, void* stream);
int b200clip_diag_sum(const void* a, int lda, const void* b, int ldb, int rows, int K, int gated, float* dots,
                      double* acc, void* stream);
/* The whole scalar tail of the softmax-CE forward in ONE launch (replaces 2x cross_entropy's reductions and the
 * scalar arithmetic of utils/loss/contrastive.py:155-164; losses.py:56-62). sums = nvec vectors of n floats, all-reduced
 * across ranks by the caller:
 *   [0] colsum  [1] rowsum  [2] target dots S_ii with the tensor core's rounding         (logits_lse_fwd; nvec = 3 or 7)
 *   [3] lse2 rows  [4] lse2 columns (logits_rowlse)  [5] S_ii from the raw features in fp32 (rowdot_raw)
 *   [6] S_ii as the column sweep's tensor core produced it                                               (nvec = 7)
 *   Stable mode (nvec = 7 and dyn[11] != 0): [0] / [1] hold the column / row gaps of logits_rowlse instead of the sums.
 * Fixed shift: rowscale[i] = c / rowsum[i], colscale[j] = c / colsum[j], c = 0.5 / n; row term t_i = ln rowsum_i +
 *   ln2 shift2 - L_ii. Stable: rowscale[i] = lse2_row[i] - log2 c, colscale likewise (b200clip_logits_bwd then forms
 *   c * softmax as 2^(L2 - rowscale) + 2^(L2 - colscale), every exponential <= c); t_i = ln2 gap_i.
 * nvec = 7: the fp32 target logit enters as t_i -= (1 - exp(-t_i)) (L_ii^fp32 - L_ii^tensor-core).
 *   loss_out[0] = c sum_i (t_row_i + t_col_i + 2 eps L_ii) - unif[0] / n   (fp64 inside);
 *   unif (may be NULL): label-smoothing uniform-target term; acc_out (may be NULL): the three fp64 sums.
 * clip_dlogtemp: out[0] = (unif / n - scal0[0] / tau) * [tau not clamped] * gmul[0]  (d loss / d log_temp). */
int b200clip_clip_finalize(const float* sums, int n, int nvec, const float* dyn, float eps, int gated, const double* unif,
                           float* rowscale, float* colscale, float* loss_out, double* acc_out, void* stream);
/* clip_finalize over the PEERS' statistics blocks (one-shot exchange instead of an all-reduce): peer_sums_host is a HOST
 * array of `world` device pointers (<= 8), block r = the nvec * n floats rank r accumulated (symmetric memory; the caller
 * runs a cross-rank barrier first). Vector 0 is summed over the peers in rank order, vectors 1.. are read from the rank
 * that owns the row (n / world rows per rank). Everything else as b200clip_clip_finalize. */
int b200clip_clip_finalize_peers(const void* const* peer_sums_host, int world, int n, int nvec, const float* dyn, float eps,
                                 int gated, const double* unif, float* rowscale, float* colscale, float* loss_out,
                                 double* acc_out, void* stream);
int b200clip_clip_dlogtemp(const double* scal0, const float* dyn, const float* gmul, const double* unif, int n,
                           float* out, void* stream);
/* Multi-GPU plumbing over symmetric memory (deepcoro_clip_b200/symm.py; one process per GPU, <= 8 ranks of one NVLink domain):
 *   symm_barrier        : cross-rank barrier on `channel` (0..3). flags_host: HOST array of `world` device pointers, entry r =
 *                         rank r's flag block (48 x int32, zeroed once at set-up: [4 channels][8 sources] arrival epochs + 4
 *                         local epoch counters) as mapped into THIS process. Epochs are counted on the device, so the call is
 *                         CUDA-graph safe; every rank must enqueue the same sequence of barriers per channel.
 *   clip_dlogtemp_peers : clip_dlogtemp with scal0 = sum over ranks of scal_host[r][0] (each rank's fp64 partial of sum G L in
 *                         its symmetric block; call after a barrier that follows the video-side backward of every rank). */
int b200clip_symm_barrier(void* const* flags_host, int world, int rank, int channel, void* stream);
/*   symm_allreduce_f32  : in-place sum-all-reduce of an fp32 buffer of n elements (n % 4 == 0, 16-byte aligned) that every rank
 *                         holds in symmetric memory (bufs_host[r] = rank r's copy): rank r reduces slice r over the peers and
 *                         writes it back into every copy; bracket with symm_barrier (before: partial sums complete, after:
 *                         every slice broadcast). Replaces dist.all_reduce of the replicated SigLIP text gradient.
 *   symm_sum_f64        : out[i] = sum_r peers_host[r][i], i < n <= 32 (fp64 scalar tails, rank order, after a barrier). */
int b200clip_symm_allreduce_f32(void* const* bufs_host, int64_t n, int world, int rank, void* stream);
int b200clip_symm_sum_f64(const void* const* peers_host, int n, int world, double* out, void* stream);
int b200clip_clip_dlogtemp_peers(const void* const* scal_host, int world, const float* dyn, const float* gmul,
                                 const double* unif, int n, float* out, void* stream);
/* Alignment diagnostics of a batch from the same forward statistics, instead of the dense [B, B] similarity +
 * log_softmax the runner recomputes after every step (runners/video_constrative_learning_runner.py:1323-1335):
 *   sums = [colsum (n) | rowsum (n) | S_ii (n)] from logits_lse_fwd of the LOCAL batch, dyn from dyn_prep;
 *   out[0] = mean S_ii (alignment_cosine), out[1] = mean (f(S_ii) / tau - lse_row_i) (alignment_logprob),
 *   out[2] = exp(out[1]) (alignment_prob); f = identity (gated = 0) or s * sigmoid(s) (gated = 1). fp64 inside. */
int b200clip_alignment_diag(const float* sums, int n, const float* dyn, int gated, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SigLIP multi-positive loss pieces (utils/loss/contrastive.py:230-315). The dense term treats every pair as a
 * negative (b200clip_logits_bwd mode 2 when gradients are needed, siglip_dense_fwd otherwise); the positives are
 * compacted from the dense fp32 pos_mask / pos_weights in ONE streaming pass and applied as exact corrections.
 *   siglip_dense_fwd : acc[0] += sum_ij [softplus(L_ij) - dyn[9] L_ij], L = clamp(S_ij/tau + bias, +-dyn[8])
 *   siglip_compact   : per video row, entries with clamp(pos_mask,0,1) > 0 -> col/y/w [B][cap], cnt[B], ysum[B]
 *                      (= sum_j y_ij, for auto_balance); *overflow = 1 if a row holds more than cap positives.
 *                      pos_mask == NULL: diagonal targets (:274-278). pos_weights may be NULL.
 *   siglip_pos       : for every entry adds  w(sp - L y) - wn sp  to acc[0] (scaled by c = 1/(B_global*T)), the
 *                      dbias / dlog_temp corrections to acc[1] / acc[2], and (dV, dT non-NULL) the gradient
 *                      corrections to dVhat[row] and (atomically) dThat[col]. Weight rule :283-298.
 *                      use_pos_weights: bit 0 = multiply by the per-pair pos_weights; bit 1 = weight rule
 *                      "pos_mask > 0" (utils/loss/siglip_pairwise.py:352) instead of "target > 0.5". Targets are
 *                      smoothed with dyn[9]: y (1 - 2 yneg) + yneg.
 *                      video_raw / text_raw (may be NULL; dtype codes as in l2norm_fwd) with their inv_norm vectors: the
 *                      gradient of a positive pair is formed with the fp32 normalised partner row instead of its bf16
 *                      operand (a row's few positives carry most of its gradient; their operand rounding is not
 *                      averaged away like that of the negatives).
 *   Entropy regulariser (contrastive.py:19-68; compute_entropy_regularization), three passes over L = clamp(R):
 *   siglip_entropy_rowsum : Z[i] += sum_j exp(L_ij - 30)                                  (caller zeroes Z)
 *   siglip_entropy_stats  : p = exp(L - 30) / Z_i ; H[i] += -sum_j p ln(p + 1e-10) ; Q[i] += sum_j p^2 / (p + 1e-10)
 *   siglip_entropy_rows   : rowvec[i] = {1 / Z_i, m_i = H_i - Q_i}; stats = {sum_i H_i, min_i H_i, max_i H_i} (double[3])
 *   siglip_entropy_coef   : stats_all [W][3] (one triple per rank) -> out = {mean, min, max, mean / ln T,
 *                           deficit = relu(threshold - mean), weight * deficit} and dyn[10] = deficit > 0 ?
 *                           -weight / B_global : 0 (read by b200clip_logits_bwd mode 3).
 * ------------------------------------------------------------------------------------------------ */
int b200clip_siglip_dense_fwd(const void* video, const void* text, int B, int T, int Kp, int ldv, int ldt,
                              const float* dyn, double* acc, void* stream);
int b200clip_siglip_entropy_rowsum(const void* video, const void* text, int B, int T, int Kp, int ldv, int ldt,
                                   const float* dyn, float* Z, void* stream);
int b200clip_siglip_entropy_stats(const void* video, const void* text, int B, int T, int Kp, int ldv, int ldt,
                                  const float* dyn, const float* Z, float* H, float* Q, void* stream);
int b200clip_siglip_entropy_rows(const float* Z, const float* H, const float* Q, int B, float* rowvec, double* stats,
                                 void* stream);
int b200clip_siglip_entropy_coef(const double* stats_all, int W, int B_global, int T, float weight, float threshold,
                                 float* dyn, float* out, void* stream);
/* Scalar tails of the SigLIP loss (one launch each instead of ~20 single-element framework launches per step):
 *   siglip_combine      : red[5]: red[0..2] = {wn_c * acc[1] + acc[4], acc[2] + acc[5], acc[0] + acc[6]} = {loss, dbias,
 *                         sum G*s} (acc[0..2] dense sums of logits_bwd / siglip_dense_fwd, acc[4..6] corrections of
 *                         siglip_pos); red[3] = c, red[4] = c^2, c = a position-weighted fp64 checksum of text_inv_norm
 *                         [T] (0 when NULL); all-reduced (SUM) across ranks by the caller in the sharded fast path;
 *   siglip_loss_out     : loss_out[0] = red[0], NaN if overflow[0] > 0 or (world > 1 and world * red[4] != red[3]^2, i.e.
 *                         the ranks of a "text replicated" job held different texts), + ent[5] with the entropy block;
 *                         diag (may be NULL) = {ent[0..5], bce loss} for get_entropy_diagnostics();
 *   siglip_scalar_grads : dlog_temp[0] = -red[2] / tau * [tau not clamped] * grad_out[0], dbias[0] = red[1] * grad_out[0]
 *                         (either output may be NULL). */
int b200clip_siglip_combine(const double* acc, double wn_c, const float* text_inv_norm, int T, double* red,
                            void* stream);
int b200clip_siglip_loss_out(const double* red, const int32_t* overflow, const float* ent, int world, float* loss_out,
                             float* diag, void* stream);
int b200clip_siglip_scalar_grads(const double* red, const float* dyn, const float* grad_out, float* dlog_temp,
                                 float* dbias, void* stream);
int b200clip_siglip_compact(const float* pos_mask, int64_t ld_mask, const float* pos_weights, int64_t ld_weights, int B,
                            int T, int cap, int32_t* col, float* y, float* w, int32_t* cnt, float* ysum,
                            int32_t* overflow, void* stream);
int b200clip_siglip_pos(const void* video, int ldv, const void* text, int ldt, int K, int Dp, int D, int hi_off, int B,
                        int T, int cap, const int32_t* col, const float* y, const float* w, const int32_t* cnt,
                        const float* ysum, const float* dyn, float positive_weight, float negative_weight, float c,
                        float gnorm, int hp, int use_pos_weights, int auto_balance, float* dV, int lddv, float* dT, int lddt, double* acc,
                        const void* video_raw, int video_dtype, int64_t ld_video_raw, const float* video_inv_norm,
                        const void* text_raw, int text_dtype, int64_t ld_text_raw, const float* text_inv_norm,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-positive softmax cross-entropy over MATERIALISED fp32 logits [n_rows, n_cols]. Replaces
 * WeightedSigLIPLoss.forward (utils/loss/weighted_siglip.py:18-51; mode 0, eps) and MultiPositiveInfoNCELoss.forward
 * (utils/loss/multi_positive_infonce.py:30-100; mode 1, reduce_sum = 0 mean / 1 sum; mode 2 = the same with
 * use_importance_weighting=True, :57-93: selected rows / columns weighted by their summed raw pos_weights / pos_mask).
 *   w_ij = max(0, pos_weights_ij [* pos_mask_ij]) (either may be NULL, not both; same row stride ld_w).
 *   multipos_fwd : rstat [n_rows][4] / cstat [n_cols][4] = {logsumexp, sum w L, sum w, #mask > 0} per row / column (one
 *                  read of logits and weights per direction), coef [n_rows + n_cols] gradient coefficients, loss_out[0];
 *                  workspace of b200clip_multipos_workspace_bytes(n_rows, n_cols) bytes.
 *   multipos_bwd : dlogits_ij = grad_out[0] * (coef_i (exp(L_ij - lse_i) P_i - w_ij) + coef_j (exp(L_ij - lse_j) Q_j - w_ij)).
 * ------------------------------------------------------------------------------------------------ */
int b200clip_multipos_workspace_bytes(int n_rows, int n_cols);
int b200clip_multipos_fwd(const float* logits, int64_t ld_logits, const float* pos_weights, const float* pos_mask,
                          int64_t ld_w, int n_rows, int n_cols, int mode, float eps, int reduce_sum, float* rstat,
                          float* cstat, float* coef, float* loss_out, void* workspace, void* stream);
int b200clip_multipos_bwd(const float* logits, int64_t ld_logits, const float* pos_weights, const float* pos_mask,
                          int64_t ld_w, int n_rows, int n_cols, const float* rstat, const float* cstat, const float* coef,
                          const float* grad_out, float* dlogits, int64_t ld_d, void* stream);

/* flag[0] = 1 (never cleared: caller zeroes) if any element of the row-major [rows, dim] fp32 (dtype 0) / fp16 (dtype 2)
 * matrix is not exactly representable in bf16 — one streaming read, no temporaries. Decides precision="auto" of the
 * streaming metrics (exact-grid evaluation embeddings run on plain bf16 operands and stay bit-exact). */
int b200clip_inexact_bf16(const void* x, int dtype, int64_t ld, int rows, int dim, int32_t* flag, void* stream);

/* MRR numerator: out[0] = sum_i 1 / (counts[i] + 1) in fp64 through the rank histogram (hist [n_bins] int32, zeroed by
 * the caller; n_bins > max count, i.e. the number of texts): sum_r hist[r] / (r + 1) in increasing-rank order with a fixed
 * reduction tree — deterministic, independent of row order and sharding. Replaces the per-row Python loop of
 * compute_metrics_streaming (utils/retrieval_metrics_streaming.py:162-172). */
int b200clip_mrr_from_counts(const int32_t* counts, int rows, int n_bins, int32_t* hist, double* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense multi-label retrieval metrics over a materialised similarity matrix. Replaces the argsort + Python loops of
 * compute_recall_at_k / compute_mrr / compute_ndcg_at_k / compute_median_rank / compute_map
 * (utils/retrieval_metrics.py:65-324): every metric is a function of the ranks of the ground-truth items,
 *     rank(i, g) = 1 + #{j : s_ij > s_ig} + #{j < g : s_ij == s_ig}     (score descending, lowest index first).
 *   dense_gt_ranks    : sim [n_rows, n_cols] (dtype code, row stride ld elements) read ONCE; gt [n_rows, G] int32,
 *                       G <= 16, entries < 0 or >= n_cols absent; ranks [n_rows, G] (0 = absent). sanitize = 1
 *                       applies nan_to_num(nan=0, posinf=1e4, neginf=-1e4) first (compute_mrr :118-120).
 *   dense_rank_metrics: per-row terms in double, in the reference's operation order: best (smallest rank, n_cols
 *                       when none), rr = 1/best | 0, ap, hit [n_rows, n_recall_k] (best <= min(k, n_cols)), ndcg
 *                       [n_rows, n_ndcg_k]; gsize [n_rows] = size of the row's ground-truth set (ideal DCG).
 * ------------------------------------------------------------------------------------------------ */
int b200clip_dense_gt_ranks(const void* sim, int dtype, int64_t ld, int n_rows, int n_cols, const int32_t* gt, int G,
                            int sanitize, int32_t* ranks, void* stream);
int b200clip_dense_rank_metrics(const int32_t* ranks, const int32_t* gsize, int n_rows, int G, int n_cols,
                                const int32_t* recall_k, int n_recall_k, const int32_t* ndcg_k, int n_ndcg_k,
                                int32_t* best, double* rr, double* ap, uint8_t* hit, double* ndcg, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K5 / K6  Streaming retrieval. Replaces compute_recall_at_k_streaming / compute_metrics_streaming
 *          (utils/retrieval_metrics_streaming.py:10-101, 104-197): chunked matmul + topk + merge + argsort rank.
 *   retrieval_sweep : one pass of similarity tiles video [n_video, >=Kp] x text [n_text, >=Kp] (operands).
 *       s_gt/gt/counts (all or none): counts[i] += #{j : s_ij > s_gt[i]  or  (s_ij == s_gt[i] and col(j) < gt[i])},
 *       col(j) = col_offset + j, the ground-truth column itself excluded  => rank_i = 1 + counts[i]
 *       (lowest-index tie rule, BASELINE.json north_star).
 *       k > 0 (<= 64): partial top-k lists part_score/part_idx [n_video][2*segs][k] (global column indices,
 *       unused entries idx = INT32_MAX), to be merged by topk_merge.  segs = retrieval_segments(...) or any >= 1.
 *   topk_merge      : out[row] = best k of `candidates` (<= 512) entries by (score desc, index asc); idx -1 = none.
 *   recall_hits     : hits[j] += #{i : counts[i] < k_values[j]}  (recall@k numerators, exact integers).
 * ------------------------------------------------------------------------------------------------ */
int b200clip_retrieval_segments(int n_video, int n_text);
int b200clip_retrieval_sweep(const void* video, const void* text, int n_video, int n_text, int Kp, int ldv, int ldt,
                             const float* s_gt, const int64_t* gt, int col_offset, int32_t* counts, int k, int segs,
                             float* part_score, int32_t* part_idx, void* stream);
int b200clip_topk_merge(const float* part_score, const int32_t* part_idx, int rows, int candidates, int k,
                        float* out_score, int64_t* out_idx, void* stream);
int b200clip_recall_hits(const int32_t* counts, int rows, const int32_t* k_values, int nk, uint64_t* hits,
                         void* stream);
/* Two-sweep top-k (replaces the same chunked topk / cat / topk merges, retrieval_metrics_streaming.py:61-82, for k <= 16;
 * opt-in on the host side with B200CLIP_TOPK2=1 until it has been measured): threshold first, then collect.
 *   retrieval_colmax  : part_max[i][slot][e] = max of s_ij over the columns j of slot (segment x column half) whose
 *                       position inside its 32-column chunk is e; part_max [n_video][2*segs][32] (16-byte aligned).
 *                       The subsets are disjoint, so the k-th largest of row i's 64*segs values bounds its k-th best
 *                       score from below.
 *   kth_largest       : thr[row] = k-th largest (with multiplicity) of vals[row][0..cand); -inf when cand < k.
 *   retrieval_collect : appends every (s_ij, col_offset + j) with s_ij >= thr[i] to row i's candidate buffer
 *                       buf_s/buf_i [n_video][cap] (cnt [n_video] zeroed, buf_i pre-filled with INT32_MAX by the
 *                       caller); *overflow = 1 if some row had more than cap candidates (the caller then falls back to
 *                       retrieval_sweep's register lists). topk_merge(buf_s, buf_i, n_video, cap, k) gives the exact
 *                       top-k by (score desc, index asc): every element tied with the k-th score is a candidate. */
int b200clip_retrieval_colmax(const void* video, const void* text, int n_video, int n_text, int Kp, int ldv, int ldt,
                              int segs, float* part_max, void* stream);
int b200clip_kth_largest(const float* vals, int rows, int cand, int k, float* thr, void* stream);
int b200clip_retrieval_collect(const void* video, const void* text, int n_video, int n_text, int Kp, int ldv, int ldt,
                               const float* thr, int col_offset, int segs, int32_t* cnt, float* buf_s, int32_t* buf_i,
                               int cap, int32_t* overflow, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K7  3D RoPE apply. Replaces Rope3D.forward's split / rotate_half / mul / add / cat graph and its autograd
 *     (models/rope_3d.py:13-17, 232-250; free function apply_rope_qk :255-282).
 *   q, k [B, heads, N, head_dim]: element strides (sb, sh, sn), head_dim contiguous; outputs contiguous, same dtype.
 *   sin_table / cos_table [N, head_dim] in the tensor dtype (the reference's cached tables, special-token rows
 *   cos = 1 / sin = 0). backward = 1 applies the adjoint (inverse rotation) to upstream gradients.
 *   Every product and the sum are rounded to the tensor dtype like the reference ops => bit-identical results.
 *   k / k_out may be NULL (single tensor).
 * ------------------------------------------------------------------------------------------------ */
int b200clip_rope3d_apply(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sn, void* q_out, const void* k,
                          int64_t k_sb, int64_t k_sh, int64_t k_sn, void* k_out, const void* sin_table,
                          const void* cos_table, int dtype, int B, int heads, int N, int head_dim, int backward,
                          void* stream);

/* ------------------------------------------------------------------------------------------------
 * K8  Attention pooling with one learnable query (models/attention_pool.py:77-93), folded form:
 *     scores s_hn = x_n . qt_h, a_h = softmax_n(s_h), xbar_h = sum_n a'_hn x_n. x [B, N, D] (D contiguous,
 *     D*sizeof(dtype) a multiple of 512 bytes), mask [B, N] bytes (non-zero = ignore) or NULL, heads <= 16.
 *     16-bit x with heads <= 8 and D % 128 == 0 runs on mma.sync tiles fed by TMA, fp32 x on CUDA cores.
 *     Attention dropout (nn.MultiheadAttention(dropout=p) in training mode, attention_pool.py:45-50): a' = a * keep /
 *     (1 - p) with a counter-based keep mask (splitmix64 of drop_seed, b*heads + h, n) that the backward regenerates;
 *     drop_p = 0 disables it. sa_h = sum_n a'_hn (1 without dropout) multiplies the value bias b_v.
 *   attnpool_fwd    : per (batch row, token split) partial (m, l, acc[heads][D]) (+ l2 = sum of kept weights when
 *                     drop_p > 0); with `weights` [B, heads, N] given instead of qt it computes plain weighted sums
 *                     (used for dqt in the backward).
 *   attnpool_merge  : merges the splits -> out [B, heads, D] (+ m, l [B, heads], sa [B, heads] if part_l2 / out_sa are
 *                     given); weighted-sum mode: part_m NULL, sum_over_b = 1 accumulates over the batch into
 *                     out [heads, D] (caller zeroes).
 *   attnpool_bwd_dx : dx_n = sum_h a'_hn dxbar_h + ds_hn qt_h, ds_hn = a_hn (kappa_hn (dxbar_h . x_n + dsa_h) - c_h),
 *                     c_h = dxbar_h . xbar_h + dsa_h sa_h; writes dx [B, N, D] (input dtype, contiguous) and
 *                     ds [B, heads, N] fp32. sa / dsa [B, heads] may be NULL when drop_p = 0. dlse [B, heads] (or NULL)
 *                     is the upstream gradient of lse_h = m_h + log l_h (d lse_h / d s_hn = a_hn, i.e. c_h -= dlse_h):
 *                     AttentionPoolWithCLS (attention_pool.py:104-197) merges the CLS key with the streamed keys
 *                     through lse.
 * ------------------------------------------------------------------------------------------------ */
int b200clip_attnpool_splits(int B, int N);
int b200clip_attnpool_fwd(const void* x, int dtype, int64_t x_sb, int64_t x_sn, const uint8_t* mask, int64_t mask_sb,
                          const float* qt, const float* weights, int64_t w_sb, int64_t w_sh, int B, int N, int D,
                          int heads, int splits, float* part_m, float* part_l, float* part_acc, float drop_p,
                          int64_t drop_seed, float* part_l2, void* stream);
int b200clip_attnpool_merge(const float* part_m, const float* part_l, const float* part_acc, int B, int splits,
                            int heads, int D, float* out, float* out_m, float* out_l, int sum_over_b,
                            const float* part_l2, float* out_sa, void* stream);
int b200clip_attnpool_bwd_dx(const void* x, int dtype, int64_t x_sb, int64_t x_sn, const uint8_t* mask, int64_t mask_sb,
                             const float* qt, const float* dxbar, const float* xbar, const float* m, const float* l,
                             int B, int N, int D, int heads, void* dx, float* ds, const float* sa, const float* dsa,
                             float drop_p, int64_t drop_seed, const float* dlse, void* stream);
/* attnpool_bwd_dx with the query gradient accumulated in the SAME pass over x (16-bit x, heads <= 8, D % 128 == 0: the
 * MMA kernels; -38 otherwise and the caller uses the separate weighted-sum launch): part_dq [B, attnpool_bwd_splits(B, N),
 * heads, D] fp32, zeroed by the caller; dqt[h, :] = sum over (b, split) = attnpool_merge(NULL, NULL, part_dq, ...,
 * sum_over_b = 1). Opt-in on the host side (B200CLIP_POOL_FUSED_DQ=1) until it has been timed on hardware. */
int b200clip_attnpool_bwd_splits(int B, int N);
int b200clip_attnpool_bwd_dx_dq(const void* x, int dtype, int64_t x_sb, int64_t x_sn, const uint8_t* mask,
                                int64_t mask_sb, const float* qt, const float* dxbar, const float* xbar, const float* m,
                                const float* l, int B, int N, int D, int heads, void* dx, float* ds, const float* sa,
                                const float* dsa, float drop_p, int64_t drop_seed, const float* dlse, float* part_dq,
                                void* stream);

/* K8 on tcgen05 tiles (csrc/attnpool_tc.cu): contiguous 16-bit x [B, N, D], heads <= 8, D % 128 == 0, D <= 512 (the C3
 * shape). 64-token tiles staged once by TMA and read by the tensor core K-major (scores) and MN-major (weighted sums),
 * accumulators in TMEM, softmax arithmetic by four epilogue warps; the backward also accumulates the query gradient in
 * the same pass (part_dq), so x is read once forward and once backward and no [B, heads, N] tensor is written.
 *   attnpool_tc_splits : token splits per batch row for this problem, or 0 when the tcgen05 path does not apply (the
 *                        caller then uses attnpool_fwd / attnpool_bwd_dx above).
 *   attnpool_tc_fwd    : partials in the format of attnpool_fwd (merged by attnpool_merge), splits = attnpool_tc_splits.
 *   attnpool_tc_bwd    : dx [B, N, D] (input dtype) and, when part_dq [B, splits, heads, D] is given, the per-split
 *                        partials of dqt = sum_n ds_hn x_n (every slot is written; summed by attnpool_merge(NULL, NULL,
 *                        part_dq, ..., sum_over_b = 1)). Same c_h / dsa / dlse conventions as attnpool_bwd_dx. */
int b200clip_attnpool_tc_splits(const void* x, int dtype, int64_t x_sb, int64_t x_sn, int B, int N, int D, int heads);
int b200clip_attnpool_tc_fwd(const void* x, int dtype, const uint8_t* mask, int64_t mask_sb, const float* qt,
                             const void* qt_img, int B, int N, int D, int heads, int splits, float* part_m, float* part_l,
                             float* part_acc, float drop_p, int64_t drop_seed, float* part_l2, void* stream);
int b200clip_attnpool_tc_bwd(const void* x, int dtype, const uint8_t* mask, int64_t mask_sb, const float* qt,
                             const float* dxbar, const float* xbar, const void* w_img, const float* cdot, const float* m,
                             const float* l, int B, int N, int D, int heads, int splits, void* dx, const float* sa,
                             const float* dsa, float drop_p, int64_t drop_seed, const float* dlse, float* part_dq,
                             void* stream);

/* The [B, D]-vector work of AttentionPool around the streaming kernels (csrc/pooltail.cu; reference
 * models/attention_pool.py:77-99 = nn.MultiheadAttention with one query + LayerNorm + optional Linear), fp32 parameters,
 * D % 128 == 0, D <= 512, heads in {1, 2, 4, 8}, out_dim % 8 == 0 (pooltail_ok = 1), forward and backward in 7 launches:
 *   pool_prep        : q0 = W_q query + b_q [D], qt_h = W_k,h^T q0_h / sqrt(Dh) [heads, D], and (qt_img != NULL) the 16-bit
 *                      hi / lo operand image of qt, D/64 x [16 x 64] with the 128-byte swizzle, which attnpool_tc_fwd
 *                      bulk-copies instead of converting qt in every CTA (img_fp16 = 1 for fp16 x, 0 for bf16).
 *   pool_tail_fwd    : merge of the partials of attnpool(_tc)_fwd -> xbar, m, l, sa; o = W_v,h xbar_h + b_v sa; y = W_o o + b_o;
 *                      LayerNorm -> yhat (normalised), rstd, yln = yhat ln_w + ln_b; out = yln or W_p yln + b_p, stored as
 *                      out_dtype (0 fp32 / 1 bf16 / 2 fp16). Clusters of 8 CTAs x 4 batch rows, rows exchanged over DSMEM.
 *   pool_tail_bwd    : from dout: dyln, dy (before the LayerNorm), d_o = W_o^T dy, dxbar_h = W_v,h^T d_o_h, dsa (NULL unless
 *                      attention dropout), cdot[b, h] = dxbar_h . xbar_h and (w_img != NULL) the per-row operand images
 *                      [B][D/64][32 x 64] that attnpool_tc_bwd bulk-copies (with cdot).
 *   pool_param_grads : dW_o = dy^T o, db_o, dW_v (per head: d_o_h^T xbar_h), db_v, dln_w, dln_b (dW_p, db_p when given).
 *   pool_qgrads      : dqt = sum of the nparts = B * splits partials of attnpool_tc_bwd, then rows [0, 2D) of the
 *                      in_proj_weight / in_proj_bias gradients (W_q, W_k; b_k gets 0) and dquery. */
int b200clip_pooltail_ok(int D, int heads, int out_dim);
int b200clip_pool_prep(const float* query, const float* in_proj_weight, const float* in_proj_bias, int D, int heads,
                       float* q0, float* qt, void* qt_img, int img_fp16, void* stream);
int b200clip_pool_tail_fwd(const float* part_m, const float* part_l, const float* part_l2, const float* part_acc, int B,
                           int splits, int heads, int D, const float* w_v, const float* b_v, const float* w_o,
                           const float* b_o, const float* ln_w, const float* ln_b, float eps, const float* w_p,
                           const float* b_p, int out_dim, float* xbar, float* m, float* l, float* sa, float* o, float* yhat,
                           float* rstd, float* yln, void* out, int out_dtype, void* stream);
int b200clip_pool_tail_bwd(const void* dout, int dout_dtype, const float* yhat, const float* rstd, const float* xbar,
                           const float* sa, const float* w_v, const float* b_v, const float* w_o, const float* ln_w,
                           const float* w_p, int out_dim, const float* qt, int B, int heads, int D, float* dyln, float* dy,
                           float* d_o, float* dxbar, float* dsa, float* cdot, void* w_img, int img_fp16, void* stream);
int b200clip_pool_param_grads(const float* dy, const float* o, const float* d_o, const float* xbar, const float* sa,
                              int use_sa, const float* dyln, const float* yhat, const void* dout, int dout_dtype,
                              const float* yln, int out_dim, int B, int heads, int D, float* dw_o, float* db_o, float* dw_v,
                              float* db_v, float* dln_w, float* dln_b, float* dw_p, float* db_p, void* stream);
int b200clip_pool_qgrads(const float* part_dq, int nparts, const float* q0, const float* query, const float* in_proj_weight,
                         int heads, int D, float* dqt, float* dw_in, float* db_in, float* dquery, void* stream);

/* The runner's inline multi-positive branch FROM FEATURES (runners/video_constrative_learning_runner.py:1256-1322, validation
 * twin :1585-1641): video [B, D], text [M, D] raw fp32 rows (D <= 1024) -> L2-normalised, gated logits
 * L_ij = s sigmoid(s) / exp(log_temp) + margin * abnormal_j, then
 *   mode 0: WeightedSigLIPLoss (utils/loss/weighted_siglip.py:38-51) with pos = clamp_min(targets [* pos_weights], 0)
 *           (the product only when pos_weights is given and not all zero: `flag`, set on the device, no host read);
 *   mode 1: sum_ij w_ij BCEWithLogits(L_ij, targets_ij) / max(1, sum targets), w = where(targets > 0, pos_weights, neg_weight)
 *           (neg_weight everywhere without pos_weights);
 * plus the alignment scalars the runner logs (:1296-1311). No [B, M] matrix is written: one CTA per row recomputes its
 * similarity row in fp32 per sweep (the branch is rank-local, B and M are batch sized).
 *   inline_mp_fwd : row_stat [B, 8], col_stat [M, 8], scalars[8] = {loss, sum targets, alignment_logprob, alignment_prob,
 *                   alignment_cosine, BCE denominator, valid rows, -}; flag: one int of device scratch.
 *   inline_mp_bwd : dvideo [B, D], dtext [M, D] (either may be NULL ... dvideo must be given when dlog_temp_acc is),
 *                   dlog_temp_acc: one fp64 (zeroed here) receiving d loss / d log_temp; grad_out: device scalar or NULL (= 1). */
int b200clip_inline_mp_fwd(const float* video, int64_t ldv, const float* text, int64_t ldt, const float* targets,
                           const float* pos_weights, int64_t ldm, const float* abnormal, float margin,
                           const float* log_temp, int B, int M, int D, int mode, float eps, float neg_weight, float* row_stat,
                           float* col_stat, float* scalars, int* flag, void* stream);
int b200clip_inline_mp_bwd(const float* video, int64_t ldv, const float* text, int64_t ldt, const float* targets,
                           const float* pos_weights, int64_t ldm, const float* abnormal, float margin,
                           const float* log_temp, int B, int M, int D, int mode, float eps, float neg_weight,
                           const float* row_stat, const float* col_stat, const float* scalars, const int* flag,
                           const float* grad_out, float* dvideo, float* dtext, double* dlog_temp_acc, void* stream);

/* Pre-LN transformer block over the N <= 16 views of a study (models/video_aggregator.py:7-54: the blocks of
 * EnhancedVideoAggregator), fp32, D % 128 == 0, D <= 512, F = hidden width (4 D) <= 2048, heads | 8
 * (xfblock_ok = 1): one cluster of 8 CTAs per study, column slices per CTA, rows exchanged over distributed shared memory.
 *   xfblock(backward = 0): out = x + drop(attn(LN1 x)); out += drop(W2 drop(gelu(W1 LN2(.) + b1)) + b2), saving what the
 *                          backward reads;  xfblock(backward = 1): dx and the row-level gradients.
 *   ptrs: HOST array of 35 device pointers —
 *     [0] x [B,N,D]  [1] out [B,N,D]  [2] key_padding_mask [B,N] bytes, non-zero = ignore (or NULL; row pitch mask_sb)
 *     [3] ln1.weight [4] ln1.bias [5] in_proj_weight [3D,D] [6] in_proj_bias [7] out_proj.weight [8] out_proj.bias
 *     [9] ln2.weight [10] ln2.bias [11] mlp.0.weight [F,D] [12] mlp.0.bias [13] mlp.3.weight [D,F] [14] mlp.3.bias
 *     saved: [15] xhat1 [R,D] [16] rstd1 [R] [17] h1 [R,D] [18] qkv [R,3D] [19] attn [B,heads,N,N] [20] o [R,D] [21] x1 [R,D]
 *            [22] xhat2 [R,D] [23] rstd2 [R] [24] h2 [R,D] [25] z [R,F] [26] u [R,F]                    (R = B N)
 *     backward: [27] dout [B,N,D] [28] dx [B,N,D] [29] d_f2 [R,D] [30] d_z [R,F] [31] d_ao [R,D] [32] d_qkv [R,3D]
 *               [33] d_h2 [R,D] [34] d_h1 [R,D]
 *   drop_p / seed: dropout of the four sites with the counter-based keep mask (0 = none).
 *   xfblock_wgrad : dw[j, i] = sum_r a[r, j] b[r, i] (J x I, I <= 2048), db[j] = sum_r a[r, j]; with a2 / xhat / dgamma /
 *                   dbeta also the LayerNorm sums dgamma[i] = sum_r a2[r, i] xhat[r, i], dbeta[i] = sum_r a2[r, i] (width D2).
 *                   Four calls give every parameter gradient of the block: (d_f2, u) -> mlp.3; (d_z, h2; d_h2, xhat2) ->
 *                   mlp.0 + ln2; (d_ao, o) -> out_proj; (d_qkv, h1; d_h1, xhat1) -> in_proj + ln1. */
int b200clip_xfblock_ok(int N, int D, int heads, int F);
int b200clip_xfblock(int backward, const void* const* ptrs, int B, int N, int D, int heads, int F, float eps1, float eps2,
                     float drop_p, int64_t seed, int64_t mask_sb, void* stream);
int b200clip_xfblock_wgrad(const float* a, int64_t lda, const float* b, int64_t ldb, float* dw, float* db, int J, int I,
                           int R, const float* a2, const float* xhat, float* dgamma, float* dbeta, int D2, void* stream);

/* The whole EnhancedVideoAggregator.forward with depth >= 1 blocks (models/video_aggregator.py:128-158) as ONE host call per
 * direction: positional add, `depth` x xfblock, final LayerNorm + query pool (K9); backward = the reverse with every parameter
 * gradient. Same kernels as the separate entry points; what this removes is host work between the launches.
 *   aggregator_sizes: sizes[0] = floats the forward saves per block, [1] = floats of backward work space, [2] = parameter-
 *                     gradient floats per block, packed in the order of ptrs[15 ..].
 *   ptrs: HOST array of 15 + 12 depth device pointers —
 *     [0] x [B,N,D] (strides x_sb, x_sn; forward)   [1] pos_encoding [pos_rows, D] or NULL   [2] mask [B,N] bytes or NULL
 *     [3] acts [depth + 1, B, N, D] (acts[0] = x + pos, acts[i + 1] = output of block i; kept for the backward)
 *     [4] saved [depth, sizes[0]]   [5] out [B, D]   [6] final_ln.weight [7] final_ln.bias [8] attn_query [D]
 *     backward: [9] dout [B,D]  [10] dact [2, B, N, D] (dx = dact[0] on return)  [11] work [sizes[1]]
 *               [12] d(final_ln.weight | final_ln.bias | attn_query) [3 D]  [13] d pos_encoding [pos_rows, D] (NULL iff [1] is)
 *               [14] block gradients [depth, sizes[2]]
 *     [15 + 12 i ..] the parameters of block i in the order of xfblock ptrs[3 .. 14].
 *   eps_host: HOST array {eps1, eps2} per block, then the final LayerNorm's; seeds_host: HOST array of `depth` dropout seeds. */
int b200clip_aggregator_sizes(int B, int N, int D, int heads, int F, int64_t* sizes);
int b200clip_aggregator(int backward, const void* const* ptrs, int depth, int B, int N, int D, int heads, int F,
                        const float* eps_host, float drop_p, const int64_t* seeds_host, int64_t mask_sb, int64_t x_sb,
                        int64_t x_sn, int pos_rows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K11  Gated-attention MIL pooling (models/multi_instance_linear_probing.py:493-507 `_attention_pooling`, and each of the
 *      two levels of :509-536 `_hierarchical_attention_pooling`), fp32:
 *        a_l = w . (tanh(V x_l + bV) * sigmoid(U x_l + bU)) + bw ; A = softmax_l(a, invalid -> -inf) ; out = sum_l drop(A_l) x_l
 *   x: S sequences of L instances of width D (element strides x_sseq / x_stok, rows contiguous, 16-byte aligned);
 *   valid [S, L] bytes, non-zero = instance present (row pitch valid_sseq) or NULL; V, U [Hd, D]; bV, bU, w [Hd]; bw [1].
 *   milpool_ok: D % 16 == 0, Hd % 8 == 0, L <= 49152.  milpool_plan fills plan[4] = {P, Z, chunks, unit_tiles}, the sizes of
 *   the caller-allocated work buffers (R = S L):
 *     forward : tg [R, 2 Hd] (tanh | sigmoid, kept for the backward), spart [unit_tiles, R], attn [R] (softmax before
 *               dropout, kept), opart [S, P, D] (only read when P > 1), out [S, D].
 *     backward: dout [S, D] -> ds [R], dpre [R, 2 Hd], wpart [Z, 2 Hd, D], fpart [chunks, 3 Hd + 4] (work), dx [R, D] (written),
 *               dW [2 Hd, D] = [dV; dU], dsmall [3 Hd + 1] = [dbV | dbU | dw | dbw] (written; sums in a fixed order).
 *   drop_p / seed: dropout of A with the counter-based keep mask (0 = none). A sequence without a valid instance gives NaN
 *   (softmax of an all -inf row), as in the reference.
 * ------------------------------------------------------------------------------------------------ */
int b200clip_milpool_ok(int L, int D, int Hd);
int b200clip_milpool_plan(int S, int L, int D, int Hd, int* plan);
int b200clip_milpool_fwd(const float* x, int64_t x_sseq, int64_t x_stok, const uint8_t* valid, int64_t valid_sseq,
                         const float* V, const float* bV, const float* U, const float* bU, const float* w, const float* bw,
                         int S, int L, int D, int Hd, float drop_p, int64_t seed, float* tg, float* spart, float* attn,
                         float* opart, float* out, void* stream);
int b200clip_milpool_bwd(const float* x, int64_t x_sseq, int64_t x_stok, const float* V, const float* U, const float* w,
                         int S, int L, int D, int Hd, float drop_p, int64_t seed, const float* tg, const float* attn,
                         const float* dout, float* ds, float* dx, float* dpre, float* wpart, float* fpart, float* dW,
                         float* dsmall, void* stream);

/* K11b The same pooling with its three products on tcgen05 (R = S L >= 1024 rows, D in {256, 512, 768}; B2_ENOSYS otherwise:
 *      use K11). Split-precision operands: v = hi + lo in bf16, a b = lo_a hi_b + hi_a lo_b + hi_a hi_b as ONE bf16 product over
 *      a three times longer K (A rows [lo | hi | hi], B rows [hi | lo | hi]), fp32 accumulation — 2^-17 relative per product.
 *   milpool_tc_plan: plan[5] = {P, chunks, slots, Hp, g_elems}. Work buffers (R = S L, bf16 unless noted):
 *     forward : x3 [R, 3 D], w3 [2 Hd, 3 D], wt3 [D, 3 Hp] (kept for the backward), tg [R, 2 Hd] fp32, spart [slots, R] fp32,
 *               attn [R] fp32, opart [S, P, D] fp32 (P > 1), out [S, D] fp32.
 *     backward: ds [R] fp32, dx [R, D] fp32 (written), dpre3 [R, 3 Hp], ghi / glo [g_elems] (128-byte aligned, K3b layout),
 *               ad [R] fp32, fpart [chunks, 3 Hd + 4] fp32, dW [2 Hd, D] fp32 with INTERLEAVED rows (2u = dV_u, 2u + 1 = dU_u),
 *               dsmall [3 Hd + 1] fp32 as in K11; one3: three device floats with one3[2] = 1. */
int b200clip_milpool_tc_plan(int S, int L, int D, int Hd, int64_t* plan);
int b200clip_milpool_tc_fwd(const float* x, int64_t x_sseq, int64_t x_stok, const uint8_t* valid, int64_t valid_sseq,
                            const float* V, const float* bV, const float* U, const float* bU, const float* w, const float* bw,
                            int S, int L, int D, int Hd, float drop_p, int64_t seed, void* x3, void* w3, void* wt3, float* tg,
                            float* spart, float* attn, float* opart, float* out, void* stream);
int b200clip_milpool_tc_bwd(const float* x, int64_t x_sseq, int64_t x_stok, const float* w, int S, int L, int D, int Hd,
                            float drop_p, int64_t seed, const void* x3, const void* wt3, const float* tg, const float* attn,
                            const float* dout, float* ds, float* dx, void* dpre3, void* ghi, void* glo, float* ad,
                            float* fpart, float* dW, float* dsmall, const float* one3, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K9  Multi-view query pool: tail of EnhancedVideoAggregator.forward (models/video_aggregator.py:119-123, 128-158).
 *   x [B, N, D] fp32 (strides sb, sn), pos [>=N, D] or NULL, final LayerNorm (ln_w, ln_b, eps), attn_query [D],
 *   mask [B, N] bytes (non-zero = masked view) or NULL. backward = 0: out [B, D]. backward = 1: dx [B, N, D] written,
 *   dpos [N, D] / dln_w / dln_b / dquery [D] ACCUMULATED atomically (caller zeroes). N*D*4 + 20 N <= 200 KB.
 * ------------------------------------------------------------------------------------------------ */
int b200clip_querypool(int backward, const float* x, int64_t x_sb, int64_t x_sn, const float* pos, const float* ln_w,
                       const float* ln_b, const float* query, const uint8_t* mask, int64_t mask_sb, int B, int N, int D,
                       float eps, float* out, const float* dout, float* dx, float* dpos, float* dln_w, float* dln_b,
                       float* dquery, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200CLIP_H_ */
