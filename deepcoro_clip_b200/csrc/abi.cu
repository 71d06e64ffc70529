// extern "C" ABI of libb200clip.so — thin argument checking + dispatch to the kernel launchers.
#include "../../include/b200clip.h"
#include "common.cuh"
#include "host_api.h"

using namespace b2;
using namespace b2host;

#define S(stream) reinterpret_cast<cudaStream_t>(stream)

extern "C" {

int b200clip_abi_version(void) { return B200CLIP_ABI_VERSION; }

const char* b200clip_strerror(int code) {
  switch (code) {
    case B2_OK: return "ok";
    case B2_EINVAL: return "invalid argument (shape, alignment or null pointer)";
    case B2_ENOSYS: return "driver entry point cuTensorMapEncodeTiled unavailable";
    case B2_ECUDA: return "CUDA launch error (is this an sm_100a device?)";
    case B2_ENOMEM: return "workspace too small";
    default: return "unknown b200clip error";
  }
}

int b200clip_sm_count(void) { return sm_count(); }

int b200clip_l2norm_fwd(const void* x, int dtype, int64_t ldx, int rows, int dim, void* operand, int ld_out, int Kp,
                        int split3_role, float* inv_norm, float* xhat_f32, int ld_hat, int normalize, void* stream) {
  if (!x || !operand) return B2_EINVAL;
  if (ld_out < (split3_role >= 0 ? 3 * Kp : Kp)) return B2_EINVAL;
  return l2norm_fwd(x, dtype, (long)ldx, rows, dim, operand, ld_out, Kp, split3_role, inv_norm, xhat_f32, ld_hat,
                    normalize, S(stream));
}

int b200clip_l2norm_bwd(const float* dxhat, int ldg, const void* x, int dtype, int64_t ldx, const float* inv_norm,
                        const void* other_x, int other_dtype, int64_t ld_other_x, const float* other_inv_norm,
                        const void* other_hi, int ld_other_hi, const float* diag_corr, const float* usum, float gscale,
                        float ucoef, const float* dev_omul, const float* dev_gmul, int rows, int dim, float* dx,
                        int64_t lddx, void* stream) {
  if (!dxhat || !x || !inv_norm || !dx) return B2_EINVAL;
  return l2norm_bwd(dxhat, ldg, x, dtype, (long)ldx, inv_norm, other_x, other_dtype, (long)ld_other_x, other_inv_norm,
                    other_hi, ld_other_hi, diag_corr, usum, gscale, ucoef, dev_omul, dev_gmul, rows, dim, dx,
                    (long)lddx, S(stream));
}

int b200clip_colsum_bf16(const void* operand, int ld, int rows, int dim, float* out, void* stream) {
  if (!operand || !out) return B2_EINVAL;
  return colsum_bf16(operand, ld, rows, dim, out, S(stream));
}

int b200clip_rowdot_bf16(const void* a, int lda, const void* b, int ldb, const int64_t* idx, int rows, int b_rows,
                         int K, float* out, void* stream) {
  if (!a || !b || !out) return B2_EINVAL;
  return rowdot_bf16(a, lda, b, ldb, reinterpret_cast<const long long*>(idx), rows, b_rows, K, out, S(stream));
}

int b200clip_rowdot_raw(const void* a, int a_dtype, int64_t lda, const float* a_inv_norm, const void* b, int b_dtype,
                        int64_t ldb, const float* b_inv_norm, int rows, int dim, float* out, void* stream) {
  return rowdot_raw(a, a_dtype, lda, a_inv_norm, b, b_dtype, ldb, b_inv_norm, rows, dim, out, S(stream));
}

int b200clip_clip_finalize(const float* sums, int n, int nvec, const float* dyn, float eps, int gated, const double* unif,
                           float* rowscale, float* colscale, float* loss_out, double* acc_out, void* stream) {
  return clip_finalize(sums, n, nvec, dyn, eps, gated, unif, rowscale, colscale, loss_out, acc_out, S(stream));
}

int b200clip_clip_finalize_peers(const void* const* peer_sums_host, int world, int n, int nvec, const float* dyn, float eps,
                                 int gated, const double* unif, float* rowscale, float* colscale, float* loss_out,
                                 double* acc_out, void* stream) {
  return clip_finalize_peers(reinterpret_cast<const float* const*>(peer_sums_host), world, n, nvec, dyn, eps, gated, unif,
                             rowscale, colscale, loss_out, acc_out, S(stream));
}

int b200clip_l2norm_fwd_multi(const void* x, int dtype, int64_t ldx, int rows, int dim, void* const* operands_host,
                              int n_operands, int64_t row_offset, int ld_out, int Kp, float* inv_norm, int normalize,
                              void* stream) {
  if (!x) return B2_EINVAL;
  return l2norm_fwd_multi(x, dtype, (long)ldx, rows, dim, operands_host, n_operands, row_offset, ld_out, Kp, inv_norm,
                          normalize, S(stream));
}

int b200clip_clip_dlogtemp(const double* scal0, const float* dyn, const float* gmul, const double* unif, int n,
                           float* out, void* stream) {
  return clip_dlogtemp(scal0, dyn, gmul, unif, n, out, S(stream));
}

int b200clip_alignment_diag(const float* sums, int n, const float* dyn, int gated, float* out, void* stream) {
  return alignment_diag(sums, n, dyn, gated, out, S(stream));
}

int b200clip_gather_rows_bf16(const void* src, int lds, const int64_t* idx, int rows, int src_rows, int K, void* dst,
                              int ldd, void* stream) {
  if (!src || !idx || !dst) return B2_EINVAL;
  return gather_rows_bf16(src, lds, reinterpret_cast<const long long*>(idx), rows, src_rows, K, dst, ldd, S(stream));
}

int b200clip_rowdot_tc(const void* a, int lda, const void* b, int ldb, int rows, int Kp, float* out, void* stream) {
  if (!a || !b || !out || rows <= 0) return B2_EINVAL;
  return rowdot_tc(a, b, rows, Kp, lda, ldb, out, S(stream));
}

int b200clip_logits_lse_fwd(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, float scale2,
                            float shift2, int gated, const float* dyn, int skip_if_stable, float* rowsum, float* colsum,
                            float* diag, int diag_off, void* stream) {
  if (!A || !B || !rowsum || !colsum) return B2_EINVAL;
  return logits_lse_fwd(A, B, Ma, Nb, Kp, lda, ldb, scale2, shift2, gated, dyn, skip_if_stable, rowsum, colsum, diag,
                        diag_off, S(stream));
}

int b200clip_rowlse_slots(int Ma, int Nb, int Kp) { return rowlse_slots(Ma, Nb, Kp); }

int b200clip_logits_rowlse(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, int gated,
                           const float* dyn, int only_if_stable, float* part, int slots, int32_t* ticket, float* lse2,
                           float* diag, int diag_off, float* gap, void* stream) {
  if (!A || !B) return B2_EINVAL;
  return logits_rowlse(A, B, Ma, Nb, Kp, lda, ldb, gated, dyn, only_if_stable, part, slots, ticket, lse2, diag, diag_off,
                       gap, S(stream));
}

int b200clip_logits_dump(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, float* out, int ldo,
                         int max_ctas, void* stream) {
  if (!A || !B || !out) return B2_EINVAL;
  return logits_dump(A, B, Ma, Nb, Kp, lda, ldb, out, ldo, max_ctas, S(stream));
}

int b200clip_logits_bwd(int mode, const void* X, const void* Y, int Nx, int Ny, int Kp, int Dp, int D, int hi_off,
                        int ldx, int ldy,
                        float scale2, float shift2, float inv_tau, float bias, float wneg_c, const float* rowscale,
                        const float* colscale, float out_scale, float gnorm, int hp, const float* dyn,
                        float ydiag, int diag_off, float* diag_corr, float* dX, int ldd, double* scal, int nseg_hint, void* stream) {
  if (!X || !Y || !dX) return B2_EINVAL;
  return logits_bwd(mode, X, Y, Nx, Ny, Kp, Dp, D, hi_off, ldx, ldy, scale2, shift2, inv_tau, bias, wneg_c, rowscale,
                    colscale, out_scale, gnorm, hp, dyn, ydiag, diag_off, diag_corr, dX, ldd, scal, nseg_hint, S(stream));
}

int b200clip_logits_bwd_both(int mode, const void* X, const void* Y, int Nx, int Ny, int Kp, int Dp, int D, int ldx, int ldy,
                             float wneg_c, const float* rowscale, const float* colscale, float gnorm, const float* dyn,
                             float ydiag, int diag_off, float* diag_corr, float* dX, int ldd, float* dY, int lddy,
                             double* scal, void* G, int64_t g_elems, void* stream) {
  if (!X || !Y || !dX || !dY || !G || !dyn) return B2_EINVAL;
  if (mode != 2 && (!rowscale || !colscale)) return B2_EINVAL;      // the sigmoid mode (2) has no row / column statistics
  if (Nx <= 0 || Ny <= 0 || Kp != Dp || Dp <= 0 || Dp % 64 || D > Dp || D <= 0) return B2_EINVAL;
  return logits_bwd_both(mode, X, Y, Nx, Ny, Kp, Dp, D, ldx, ldy, wneg_c, rowscale, colscale, gnorm, dyn, ydiag, diag_off,
                         diag_corr, dX, ldd, dY, lddy, scal, G, g_elems, S(stream));
}

int b200clip_gstore_elems(int Nx, int Ny, int64_t* elems) {
  if (!elems || Nx <= 0 || Ny <= 0) return B2_EINVAL;
  *elems = gstore_elems(Nx, Ny);
  return B2_OK;
}

int b200clip_gt_gemm(const void* G, int64_t g_elems, int Nx, int Ny, const void* X, int ldx, int Dp, int D, const float* dyn,
                     float gnorm, float* dY, int ldd, void* stream) {
  return gt_gemm(G, g_elems, Nx, Ny, X, ldx, Dp, D, dyn, gnorm, dY, ldd, S(stream));
}

int b200clip_dyn_prep(const float* log_temp, const float* bias, float clamp_min, float bound, float* dyn,
                      void* stream) {
  if (!log_temp || !dyn) return B2_EINVAL;
  return dyn_prep(log_temp, bias, clamp_min, bound, dyn, S(stream));
}

int b200clip_dyn_set_siglip(float* dyn, float logit_clamp, float neg_target, void* stream) {
  if (!dyn || !(logit_clamp > 0.f)) return B2_EINVAL;
  return dyn_set_siglip(dyn, logit_clamp, neg_target, S(stream));
}

int b200clip_dyn_set_stable(float* dyn, int stable, void* stream) {
  if (!dyn) return B2_EINVAL;
  return dyn_set_stable(dyn, stable, S(stream));
}

int b200clip_lse_finalize(const float* sums, int n, const float* dyn, float c, float* scale_out, double* acc,
                          void* stream) {
  if (!sums || !dyn) return B2_EINVAL;
  return lse_finalize(sums, n, dyn, c, scale_out, acc, S(stream));
}

int b200clip_vec_fsum(const float* v, int n, int gated, double* acc, void* stream) {
  if (!v || !acc) return B2_EINVAL;
  return vec_fsum(v, n, gated, acc, S(stream));
}

int b200clip_diag_sum(const void* a, int lda, const void* b, int ldb, int rows, int K, int gated, float* dots,
                      double* acc, void* stream) {
  if (!a || !b || !acc) return B2_EINVAL;
  return diag_sum(a, lda, b, ldb, rows, K, gated, dots, acc, S(stream));
}

int b200clip_siglip_dense_fwd(const void* video, const void* text, int B, int T, int Kp, int ldv, int ldt,
                              const float* dyn, double* acc, void* stream) {
  if (!video || !text || !dyn || !acc) return B2_EINVAL;
  return siglip_dense_fwd(video, text, B, T, Kp, ldv, ldt, dyn, acc, S(stream));
}

int b200clip_siglip_entropy_rowsum(const void* video, const void* text, int B, int T, int Kp, int ldv, int ldt,
                                   const float* dyn, float* Z, void* stream) {
  if (!video || !text || !dyn || !Z) return B2_EINVAL;
  return siglip_entropy_rowsum(video, text, B, T, Kp, ldv, ldt, dyn, Z, S(stream));
}

int b200clip_siglip_entropy_stats(const void* video, const void* text, int B, int T, int Kp, int ldv, int ldt,
                                  const float* dyn, const float* Z, float* H, float* Q, void* stream) {
  if (!video || !text || !dyn || !Z || !H || !Q) return B2_EINVAL;
  return siglip_entropy_stats(video, text, B, T, Kp, ldv, ldt, dyn, Z, H, Q, S(stream));
}

int b200clip_siglip_entropy_rows(const float* Z, const float* H, const float* Q, int B, float* rowvec, double* stats,
                                 void* stream) {
  if (!Z || !H || !Q || !rowvec || !stats) return B2_EINVAL;
  return siglip_entropy_rows(Z, H, Q, B, rowvec, stats, S(stream));
}

int b200clip_siglip_entropy_coef(const double* stats_all, int W, int B_global, int T, float weight, float threshold,
                                 float* dyn, float* out, void* stream) {
  if (!stats_all || !dyn || !out) return B2_EINVAL;
  return siglip_entropy_coef(stats_all, W, B_global, T, weight, threshold, dyn, out, S(stream));
}

int b200clip_siglip_combine(const double* acc, double wn_c, const float* text_inv_norm, int T, double* red,
                            void* stream) {
  if (!acc || !red || (text_inv_norm && T <= 0)) return B2_EINVAL;
  return siglip_combine(acc, wn_c, text_inv_norm, T, red, S(stream));
}

int b200clip_siglip_loss_out(const double* red, const int32_t* overflow, const float* ent, int world, float* loss_out,
                             float* diag, void* stream) {
  if (!red || !loss_out) return B2_EINVAL;
  return siglip_loss_out(red, overflow, ent, world, loss_out, diag, S(stream));
}

int b200clip_siglip_scalar_grads(const double* red, const float* dyn, const float* grad_out, float* dlog_temp,
                                 float* dbias, void* stream) {
  if (!red || !dyn || !grad_out) return B2_EINVAL;
  return siglip_scalar_grads(red, dyn, grad_out, dlog_temp, dbias, S(stream));
}

int b200clip_siglip_compact(const float* pos_mask, int64_t ld_mask, const float* pos_weights, int64_t ld_weights, int B,
                            int T, int cap, int32_t* col, float* y, float* w, int32_t* cnt, float* ysum,
                            int32_t* overflow, void* stream) {
  if (!col || !y || !w || !cnt || !ysum || !overflow) return B2_EINVAL;
  return siglip_compact(pos_mask, (long)ld_mask, pos_weights, (long)ld_weights, B, T, cap, col, y, w, cnt, ysum,
                        overflow, S(stream));
}

int b200clip_siglip_pos(const void* video, int ldv, const void* text, int ldt, int K, int Dp, int D, int hi_off, int B,
                        int T, int cap, const int32_t* col, const float* y, const float* w, const int32_t* cnt,
                        const float* ysum, const float* dyn, float positive_weight, float negative_weight, float c,
                        float gnorm, int hp, int use_pos_weights, int auto_balance, float* dV, int lddv, float* dT, int lddt, double* acc,
                        const void* video_raw, int video_dtype, int64_t ld_video_raw, const float* video_inv_norm,
                        const void* text_raw, int text_dtype, int64_t ld_text_raw, const float* text_inv_norm,
                        void* stream) {
  if (!video || !text || !col || !y || !w || !cnt || !ysum || !dyn || !acc) return B2_EINVAL;
  return siglip_pos(video, ldv, text, ldt, K, Dp, D, hi_off, B, T, cap, col, y, w, cnt, ysum, dyn, positive_weight,
                    negative_weight, c, gnorm, hp, use_pos_weights, auto_balance, dV, lddv, dT, lddt, acc, video_raw,
                    video_dtype, ld_video_raw, video_inv_norm, text_raw, text_dtype, ld_text_raw, text_inv_norm,
                    S(stream));
}

int b200clip_multipos_workspace_bytes(int n_rows, int n_cols) { return multipos_workspace_bytes(n_rows, n_cols); }

int b200clip_multipos_fwd(const float* logits, int64_t ld_logits, const float* pos_weights, const float* pos_mask,
                          int64_t ld_w, int n_rows, int n_cols, int mode, float eps, int reduce_sum, float* rstat,
                          float* cstat, float* coef, float* loss_out, void* workspace, void* stream) {
  if (!logits || !rstat || !cstat || !coef || !loss_out || !workspace) return B2_EINVAL;
  return multipos_fwd(logits, (long long)ld_logits, pos_weights, pos_mask, (long long)ld_w, n_rows, n_cols, mode, eps,
                      reduce_sum, rstat, cstat, coef, loss_out, workspace, S(stream));
}

int b200clip_multipos_bwd(const float* logits, int64_t ld_logits, const float* pos_weights, const float* pos_mask,
                          int64_t ld_w, int n_rows, int n_cols, const float* rstat, const float* cstat, const float* coef,
                          const float* grad_out, float* dlogits, int64_t ld_d, void* stream) {
  if (!logits || !rstat || !cstat || !coef || !grad_out || !dlogits) return B2_EINVAL;
  return multipos_bwd(logits, (long long)ld_logits, pos_weights, pos_mask, (long long)ld_w, n_rows, n_cols, rstat, cstat,
                      coef, grad_out, dlogits, (long long)ld_d, S(stream));
}

int b200clip_inexact_bf16(const void* x, int dtype, int64_t ld, int rows, int dim, int32_t* flag, void* stream) {
  if (!x || !flag) return B2_EINVAL;
  return inexact_bf16(x, dtype, (long long)ld, rows, dim, flag, S(stream));
}

int b200clip_mrr_from_counts(const int32_t* counts, int rows, int n_bins, int32_t* hist, double* out, void* stream) {
  if (!counts || !hist || !out) return B2_EINVAL;
  return mrr_from_counts(counts, rows, n_bins, hist, out, S(stream));
}

int b200clip_dense_gt_ranks(const void* sim, int dtype, int64_t ld, int n_rows, int n_cols, const int32_t* gt, int G,
                            int sanitize, int32_t* ranks, void* stream) {
  if (!sim || !gt || !ranks) return B2_EINVAL;
  return dense_gt_ranks(sim, dtype, (long long)ld, n_rows, n_cols, gt, G, sanitize, ranks, S(stream));
}

int b200clip_dense_rank_metrics(const int32_t* ranks, const int32_t* gsize, int n_rows, int G, int n_cols,
                                const int32_t* recall_k, int n_recall_k, const int32_t* ndcg_k, int n_ndcg_k,
                                int32_t* best, double* rr, double* ap, uint8_t* hit, double* ndcg, void* stream) {
  if (!ranks || !gsize || !best || !rr || !ap) return B2_EINVAL;
  if ((n_recall_k > 0 && (!recall_k || !hit)) || (n_ndcg_k > 0 && (!ndcg_k || !ndcg))) return B2_EINVAL;
  return dense_rank_metrics(ranks, gsize, n_rows, G, n_cols, recall_k, n_recall_k, ndcg_k, n_ndcg_k, best, rr, ap, hit,
                            ndcg, S(stream));
}

int b200clip_retrieval_segments(int n_video, int n_text) { return retrieval_segments(n_video, n_text); }

int b200clip_retrieval_sweep(const void* video, const void* text, int n_video, int n_text, int Kp, int ldv, int ldt,
                             const float* s_gt, const int64_t* gt, int col_offset, int32_t* counts, int k, int segs,
                             float* part_score, int32_t* part_idx, void* stream) {
  if (!video || !text) return B2_EINVAL;
  return retrieval_sweep(video, text, n_video, n_text, Kp, ldv, ldt, s_gt, reinterpret_cast<const long long*>(gt),
                         col_offset, counts, k, segs, part_score, part_idx, S(stream));
}

int b200clip_topk_merge(const float* part_score, const int32_t* part_idx, int rows, int candidates, int k,
                        float* out_score, int64_t* out_idx, void* stream) {
  if (!part_score || !part_idx || !out_score || !out_idx) return B2_EINVAL;
  return topk_merge(part_score, part_idx, rows, candidates, k, out_score, reinterpret_cast<long long*>(out_idx),
                    S(stream));
}

int b200clip_retrieval_colmax(const void* video, const void* text, int n_video, int n_text, int Kp, int ldv, int ldt,
                              int segs, float* part_max, void* stream) {
  if (!video || !text) return B2_EINVAL;
  return retrieval_colmax(video, text, n_video, n_text, Kp, ldv, ldt, segs, part_max, S(stream));
}

int b200clip_kth_largest(const float* vals, int rows, int cand, int k, float* thr, void* stream) {
  return kth_largest(vals, rows, cand, k, thr, S(stream));
}

int b200clip_retrieval_collect(const void* video, const void* text, int n_video, int n_text, int Kp, int ldv, int ldt,
                               const float* thr, int col_offset, int segs, int32_t* cnt, float* buf_s, int32_t* buf_i,
                               int cap, int32_t* overflow, void* stream) {
  if (!video || !text) return B2_EINVAL;
  return retrieval_collect(video, text, n_video, n_text, Kp, ldv, ldt, thr, col_offset, segs, cnt, buf_s, buf_i, cap,
                           overflow, S(stream));
}

int b200clip_recall_hits(const int32_t* counts, int rows, const int32_t* k_values, int nk, uint64_t* hits,
                         void* stream) {
  if (!counts || !k_values || !hits) return B2_EINVAL;
  return recall_hits(counts, rows, k_values, nk, reinterpret_cast<unsigned long long*>(hits), S(stream));
}

int b200clip_rope3d_apply(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sn, void* q_out, const void* k,
                          int64_t k_sb, int64_t k_sh, int64_t k_sn, void* k_out, const void* sin_table,
                          const void* cos_table, int dtype, int B, int heads, int N, int head_dim, int backward,
                          void* stream) {
  return rope3d_apply(q, q_sb, q_sh, q_sn, q_out, k, k_sb, k_sh, k_sn, k_out, sin_table, cos_table, dtype, B, heads, N,
                      head_dim, backward, S(stream));
}

int b200clip_attnpool_splits(int B, int N) { return attnpool_splits(B, N); }

int b200clip_attnpool_fwd(const void* x, int dtype, int64_t x_sb, int64_t x_sn, const uint8_t* mask, int64_t mask_sb,
                          const float* qt, const float* weights, int64_t w_sb, int64_t w_sh, int B, int N, int D,
                          int heads, int splits, float* part_m, float* part_l, float* part_acc, float drop_p,
                          int64_t drop_seed, float* part_l2, void* stream) {
  if (drop_p < 0.f || drop_p >= 1.f || (drop_p > 0.f && !part_l2)) return B2_EINVAL;
  return attnpool_fwd(x, dtype, x_sb, x_sn, mask, mask_sb, qt, weights, w_sb, w_sh, B, N, D, heads, splits, part_m,
                      part_l, part_acc, drop_p, (unsigned long long)drop_seed, part_l2, S(stream));
}

int b200clip_attnpool_merge(const float* part_m, const float* part_l, const float* part_acc, int B, int splits,
                            int heads, int D, float* out, float* out_m, float* out_l, int sum_over_b,
                            const float* part_l2, float* out_sa, void* stream) {
  return attnpool_merge(part_m, part_l, part_acc, B, splits, heads, D, out, out_m, out_l, sum_over_b, part_l2, out_sa,
                        S(stream));
}

int b200clip_attnpool_bwd_dx(const void* x, int dtype, int64_t x_sb, int64_t x_sn, const uint8_t* mask, int64_t mask_sb,
                             const float* qt, const float* dxbar, const float* xbar, const float* m, const float* l,
                             int B, int N, int D, int heads, void* dx, float* ds, const float* sa, const float* dsa,
                             float drop_p, int64_t drop_seed, const float* dlse, void* stream) {
  return attnpool_bwd_dx(x, dtype, x_sb, x_sn, mask, mask_sb, qt, dxbar, xbar, m, l, B, N, D, heads, dx, ds, sa, dsa,
                         drop_p, (unsigned long long)drop_seed, dlse, S(stream));
}

int b200clip_attnpool_bwd_splits(int B, int N) { return attnpool_bwd_splits(B, N); }

int b200clip_attnpool_bwd_dx_dq(const void* x, int dtype, int64_t x_sb, int64_t x_sn, const uint8_t* mask,
                                int64_t mask_sb, const float* qt, const float* dxbar, const float* xbar, const float* m,
                                const float* l, int B, int N, int D, int heads, void* dx, float* ds, const float* sa,
                                const float* dsa, float drop_p, int64_t drop_seed, const float* dlse, float* part_dq,
                                void* stream) {
  return attnpool_bwd_dx_dq(x, dtype, x_sb, x_sn, mask, mask_sb, qt, dxbar, xbar, m, l, B, N, D, heads, dx, ds, sa, dsa,
                            drop_p, (unsigned long long)drop_seed, dlse, part_dq, S(stream));
}

int b200clip_attnpool_tc_splits(const void* x, int dtype, int64_t x_sb, int64_t x_sn, int B, int N, int D, int heads) {
  return attnpool_tc_ok(x, dtype, x_sb, x_sn, N, D, heads) ? attnpool_tc_splits(B, N) : 0;
}

int b200clip_attnpool_tc_fwd(const void* x, int dtype, const uint8_t* mask, int64_t mask_sb, const float* qt,
                             const void* qt_img, int B, int N, int D, int heads, int splits, float* part_m, float* part_l,
                             float* part_acc, float drop_p, int64_t drop_seed, float* part_l2, void* stream) {
  if (drop_p < 0.f || drop_p >= 1.f || (drop_p > 0.f && !part_l2)) return B2_EINVAL;
  if (!attnpool_tc_ok(x, dtype, (int64_t)N * D, D, N, D, heads)) return B2_ENOSYS;
  return attnpool_tc_fwd(x, dtype, mask, mask_sb, qt, qt_img, B, N, D, heads, splits, part_m, part_l, part_acc, drop_p,
                         (unsigned long long)drop_seed, part_l2, S(stream));
}

int b200clip_attnpool_tc_bwd(const void* x, int dtype, const uint8_t* mask, int64_t mask_sb, const float* qt,
                             const float* dxbar, const float* xbar, const void* w_img, const float* cdot, const float* m,
                             const float* l, int B, int N, int D, int heads, int splits, void* dx, const float* sa,
                             const float* dsa, float drop_p, int64_t drop_seed, const float* dlse, float* part_dq,
                             void* stream) {
  if (!attnpool_tc_ok(x, dtype, (int64_t)N * D, D, N, D, heads)) return B2_ENOSYS;
  return attnpool_tc_bwd(x, dtype, mask, mask_sb, qt, dxbar, xbar, w_img, cdot, m, l, B, N, D, heads, splits, dx, sa, dsa,
                         drop_p, (unsigned long long)drop_seed, dlse, part_dq, S(stream));
}

int b200clip_pooltail_ok(int D, int heads, int out_dim) { return pooltail_ok(D, heads, out_dim) ? 1 : 0; }

int b200clip_pool_prep(const float* query, const float* in_proj_weight, const float* in_proj_bias, int D, int heads,
                       float* q0, float* qt, void* qt_img, int img_fp16, void* stream) {
  return pool_prep(query, in_proj_weight, in_proj_bias, D, heads, q0, qt, qt_img, img_fp16, S(stream));
}

int b200clip_pool_tail_fwd(const float* part_m, const float* part_l, const float* part_l2, const float* part_acc, int B,
                           int splits, int heads, int D, const float* w_v, const float* b_v, const float* w_o,
                           const float* b_o, const float* ln_w, const float* ln_b, float eps, const float* w_p,
                           const float* b_p, int out_dim, float* xbar, float* m, float* l, float* sa, float* o, float* yhat,
                           float* rstd, float* yln, void* out, int out_dtype, void* stream) {
  return pool_tail_fwd(part_m, part_l, part_l2, part_acc, B, splits, heads, D, w_v, b_v, w_o, b_o, ln_w, ln_b, eps, w_p, b_p,
                       out_dim, xbar, m, l, sa, o, yhat, rstd, yln, out, out_dtype, S(stream));
}

int b200clip_pool_tail_bwd(const void* dout, int dout_dtype, const float* yhat, const float* rstd, const float* xbar,
                           const float* sa, const float* w_v, const float* b_v, const float* w_o, const float* ln_w,
                           const float* w_p, int out_dim, const float* qt, int B, int heads, int D, float* dyln, float* dy,
                           float* d_o, float* dxbar, float* dsa, float* cdot, void* w_img, int img_fp16, void* stream) {
  return pool_tail_bwd(dout, dout_dtype, yhat, rstd, xbar, sa, w_v, b_v, w_o, ln_w, w_p, out_dim, qt, B, heads, D, dyln, dy,
                       d_o, dxbar, dsa, cdot, w_img, img_fp16, S(stream));
}

int b200clip_pool_param_grads(const float* dy, const float* o, const float* d_o, const float* xbar, const float* sa,
                              int use_sa, const float* dyln, const float* yhat, const void* dout, int dout_dtype,
                              const float* yln, int out_dim, int B, int heads, int D, float* dw_o, float* db_o, float* dw_v,
                              float* db_v, float* dln_w, float* dln_b, float* dw_p, float* db_p, void* stream) {
  return pool_param_grads(dy, o, d_o, xbar, sa, use_sa, dyln, yhat, dout, dout_dtype, yln, out_dim, B, heads, D, dw_o, db_o,
                          dw_v, db_v, dln_w, dln_b, dw_p, db_p, S(stream));
}

int b200clip_pool_qgrads(const float* part_dq, int nparts, const float* q0, const float* query, const float* in_proj_weight,
                         int heads, int D, float* dqt, float* dw_in, float* db_in, float* dquery, void* stream) {
  return pool_qgrads(part_dq, nparts, q0, query, in_proj_weight, heads, D, dqt, dw_in, db_in, dquery, S(stream));
}

int b200clip_inline_mp_fwd(const float* video, int64_t ldv, const float* text, int64_t ldt, const float* targets,
                           const float* pos_weights, int64_t ldm, const float* abnormal, float margin,
                           const float* log_temp, int B, int M, int D, int mode, float eps, float neg_weight, float* row_stat,
                           float* col_stat, float* scalars, int* flag, void* stream) {
  return inline_mp_fwd(video, ldv, text, ldt, targets, pos_weights, ldm, abnormal, margin, log_temp, B, M, D, mode, eps,
                       neg_weight, row_stat, col_stat, scalars, flag, S(stream));
}

int b200clip_inline_mp_bwd(const float* video, int64_t ldv, const float* text, int64_t ldt, const float* targets,
                           const float* pos_weights, int64_t ldm, const float* abnormal, float margin,
                           const float* log_temp, int B, int M, int D, int mode, float eps, float neg_weight,
                           const float* row_stat, const float* col_stat, const float* scalars, const int* flag,
                           const float* grad_out, float* dvideo, float* dtext, double* dlog_temp_acc, void* stream) {
  return inline_mp_bwd(video, ldv, text, ldt, targets, pos_weights, ldm, abnormal, margin, log_temp, B, M, D, mode, eps,
                       neg_weight, row_stat, col_stat, scalars, flag, grad_out, dvideo, dtext, dlog_temp_acc, S(stream));
}

int b200clip_xfblock_ok(int N, int D, int heads, int F) { return xfblock_ok(N, D, heads, F) ? 1 : 0; }

int b200clip_xfblock(int backward, const void* const* ptrs, int B, int N, int D, int heads, int F, float eps1, float eps2,
                     float drop_p, int64_t seed, int64_t mask_sb, void* stream) {
  if (drop_p < 0.f || drop_p >= 1.f) return B2_EINVAL;
  return xfblock(backward, ptrs, B, N, D, heads, F, eps1, eps2, drop_p, (unsigned long long)seed, mask_sb, S(stream));
}

int b200clip_xfblock_wgrad(const float* a, int64_t lda, const float* b, int64_t ldb, float* dw, float* db, int J, int I,
                           int R, const float* a2, const float* xhat, float* dgamma, float* dbeta, int D2, void* stream) {
  return xfblock_wgrad(a, lda, b, ldb, dw, db, J, I, R, a2, xhat, dgamma, dbeta, D2, S(stream));
}

int b200clip_aggregator_sizes(int B, int N, int D, int heads, int F, int64_t* sizes) {
  if (!sizes || B < 1 || !xfblock_ok(N, D, heads, F)) return B2_EINVAL;
  long long t[3];
  aggregator_sizes(B, N, D, heads, F, t);
  for (int i = 0; i < 3; ++i) sizes[i] = t[i];
  return B2_OK;
}

int b200clip_aggregator(int backward, const void* const* ptrs, int depth, int B, int N, int D, int heads, int F,
                        const float* eps_host, float drop_p, const int64_t* seeds_host, int64_t mask_sb, int64_t x_sb,
                        int64_t x_sn, int pos_rows, void* stream) {
  if (drop_p < 0.f || drop_p >= 1.f || depth < 1 || depth > 64) return B2_EINVAL;
  long long seeds[64];
  for (int i = 0; i < depth; ++i) seeds[i] = seeds_host ? seeds_host[i] : 0;
  return aggregator(backward, ptrs, depth, B, N, D, heads, F, eps_host, drop_p, seeds, mask_sb, x_sb, x_sn, pos_rows, S(stream));
}

int b200clip_milpool_ok(int L, int D, int Hd) { return milpool_ok(L, D, Hd) ? 1 : 0; }

int b200clip_milpool_plan(int S, int L, int D, int Hd, int* plan) {
  if (!plan || S < 1 || !milpool_ok(L, D, Hd)) return B2_EINVAL;
  milpool_plan(S, L, D, Hd, plan);
  return B2_OK;
}

int b200clip_milpool_fwd(const float* x, int64_t x_sseq, int64_t x_stok, const uint8_t* valid, int64_t valid_sseq,
                         const float* V, const float* bV, const float* U, const float* bU, const float* w, const float* bw,
                         int S, int L, int D, int Hd, float drop_p, int64_t seed, float* tg, float* spart, float* attn,
                         float* opart, float* out, void* stream) {
  return milpool_fwd(x, x_sseq, x_stok, valid, valid_sseq, V, bV, U, bU, w, bw, S, L, D, Hd, drop_p,
                     (unsigned long long)seed, tg, spart, attn, opart, out, S(stream));
}

int b200clip_milpool_bwd(const float* x, int64_t x_sseq, int64_t x_stok, const float* V, const float* U, const float* w,
                         int S, int L, int D, int Hd, float drop_p, int64_t seed, const float* tg, const float* attn,
                         const float* dout, float* ds, float* dx, float* dpre, float* wpart, float* fpart, float* dW,
                         float* dsmall, void* stream) {
  return milpool_bwd(x, x_sseq, x_stok, V, U, w, S, L, D, Hd, drop_p, (unsigned long long)seed, tg, attn, dout, ds, dx,
                     dpre, wpart, fpart, dW, dsmall, S(stream));
}

int b200clip_milpool_tc_plan(int S, int L, int D, int Hd, int64_t* plan) {
  if (!plan || S < 1 || !milpool_tc_ok((long long)S * L, L, D, Hd)) return B2_ENOSYS;
  long long t[5];
  milpool_tc_plan(S, L, D, Hd, t);
  for (int i = 0; i < 5; ++i) plan[i] = t[i];
  return B2_OK;
}

int b200clip_milpool_tc_fwd(const float* x, int64_t x_sseq, int64_t x_stok, const uint8_t* valid, int64_t valid_sseq,
                            const float* V, const float* bV, const float* U, const float* bU, const float* w, const float* bw,
                            int S, int L, int D, int Hd, float drop_p, int64_t seed, void* x3, void* w3, void* wt3, float* tg,
                            float* spart, float* attn, float* opart, float* out, void* stream) {
  return milpool_tc_fwd(x, x_sseq, x_stok, valid, valid_sseq, V, bV, U, bU, w, bw, S, L, D, Hd, drop_p,
                        (unsigned long long)seed, x3, w3, wt3, tg, spart, attn, opart, out, S(stream));
}

int b200clip_milpool_tc_bwd(const float* x, int64_t x_sseq, int64_t x_stok, const float* w, int S, int L, int D, int Hd,
                            float drop_p, int64_t seed, const void* x3, const void* wt3, const float* tg, const float* attn,
                            const float* dout, float* ds, float* dx, void* dpre3, void* ghi, void* glo, float* ad,
                            float* fpart, float* dW, float* dsmall, const float* one3, void* stream) {
  return milpool_tc_bwd(x, x_sseq, x_stok, w, S, L, D, Hd, drop_p, (unsigned long long)seed, x3, wt3, tg, attn, dout, ds, dx,
                        dpre3, ghi, glo, ad, fpart, dW, dsmall, one3, S(stream));
}

int b200clip_symm_barrier(void* const* flags_host, int world, int rank, int channel, void* stream) {
  return symm_barrier(flags_host, world, rank, channel, S(stream));
}

int b200clip_symm_allreduce_f32(void* const* bufs_host, int64_t n, int world, int rank, void* stream) {
  return symm_allreduce_f32(bufs_host, n, world, rank, S(stream));
}

int b200clip_symm_sum_f64(const void* const* peers_host, int n, int world, double* out, void* stream) {
  return symm_sum_f64(peers_host, n, world, out, S(stream));
}

int b200clip_clip_dlogtemp_peers(const void* const* scal_host, int world, const float* dyn, const float* gmul,
                                 const double* unif, int n, float* out, void* stream) {
  return clip_dlogtemp_peers(scal_host, world, dyn, gmul, unif, n, out, S(stream));
}

int b200clip_l2norm_fwd_mc(const void* x, int dtype, int64_t ldx, int rows, int dim, void* mc_operand, int64_t row_offset,
                           int ld_out, int Kp, float* inv_norm, int normalize, void* stream) {
  return l2norm_fwd_mc(x, dtype, (long)ldx, rows, dim, mc_operand, row_offset, ld_out, Kp, inv_norm, normalize, S(stream));
}

int b200clip_querypool(int backward, const float* x, int64_t x_sb, int64_t x_sn, const float* pos, const float* ln_w,
                       const float* ln_b, const float* query, const uint8_t* mask, int64_t mask_sb, int B, int N, int D,
                       float eps, float* out, const float* dout, float* dx, float* dpos, float* dln_w, float* dln_b,
                       float* dquery, void* stream) {
  return querypool(backward, x, x_sb, x_sn, pos, ln_w, ln_b, query, mask, mask_sb, B, N, D, eps, out, dout, dx, dpos,
                   dln_w, dln_b, dquery, S(stream));
}

}  // extern "C"
