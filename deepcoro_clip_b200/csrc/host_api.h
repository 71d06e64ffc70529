// Internal C++ host entry points (one per kernel family); abi.cu wraps them in the extern "C" ABI.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2host {

// l2norm.cu
int l2norm_fwd(const void* x, int dtype, long ldx, int rows, int dim, void* out, int ldo, int Kp, int split3_role,
               float* inv_norm, float* xhat_f32, int ldh, int normalize, cudaStream_t s);
int l2norm_fwd_multi(const void* x, int dtype, long ldx, int rows, int dim, void* const* outs_host, int n_out,
                     long long row_offset, int ldo, int Kp, float* inv_norm, int normalize, cudaStream_t s);
int l2norm_bwd(const float* dxh, int ldg, const void* x, int dtype, long ldx, const float* inv_norm, const void* ox,
               int odtype, long ldox, const float* oinv, const void* ohi, int ldohi, const float* dc,
               const float* usum, float gscale, float ucoef, const float* dev_omul, const float* dev_gmul, int rows,
               int dim, float* dx, long lddx, cudaStream_t s);
int inexact_bf16(const void* x, int dtype, long long ld, int rows, int dim, int* flag, cudaStream_t s);
int colsum_bf16(const void* xh, int ld, int rows, int dim, float* out, cudaStream_t s);
int gather_rows_bf16(const void* src, int lds, const long long* idx, int rows, int src_rows, int K, void* dst, int ldd,
                     cudaStream_t s);
int rowdot_bf16(const void* a, int lda, const void* b, int ldb, const long long* idx, int rows, int b_rows, int K,
                float* out, cudaStream_t s);
int rowdot_raw(const void* a, int adtype, long long lda, const float* ainv, const void* b, int bdtype, long long ldb,
               const float* binv, int rows, int dim, float* out, cudaStream_t s);

// logits_fwd.cu
int logits_lse_fwd(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, float scale2,
                   float shift2, int gated, const float* dyn, int skip_if_stable, float* rowsum, float* colsum,
                   float* diag, int diag_off, cudaStream_t stream);
int rowlse_slots(int Ma, int Nb, int Kp);
int logits_rowlse(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, int gated, const float* dyn,
                  int only_if_stable, float* part, int slots, int* ticket, float* lse2, float* diag, int diag_off,
                  float* gap, cudaStream_t stream);
int rowdot_tc(const void* A, const void* B, int rows, int Kp, int lda, int ldb, float* out, cudaStream_t stream);
int logits_dump(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, float* out, int ldo,
                int max_ctas, cudaStream_t stream);

// logits_bwd.cu
int logits_bwd(int mode, const void* X, const void* Y, int Nx, int Ny, int Kp, int Dp, int D, int hi_off, int ldx,
               int ldy,
               float scale2, float shift2, float inv_tau, float bias, float wneg_c, const float* rowscale,
               const float* colscale, float out_scale, float gnorm, int hp, const float* dyn, float ydiag,
               int diag_off, float* diag_corr,
               float* dX, int ldd, double* scal, int nseg_hint, cudaStream_t stream);

// logits_bwd2.cu: the same contract on CTA pairs (tcgen05 cta_group::2); Kp <= 512, Dp % 128 == 0, hp == 0 only
int logits_bwd_pair(int mode, const void* X, const void* Y, int Nx, int Ny, int Kp, int Dp, int D, int hi_off, int ldx,
                    int ldy, float scale2, float shift2, float inv_tau, float bias, float wneg_c,
                    const float* rowscale, const float* colscale, float out_scale, float gnorm, int hp,
                    const float* dyn, float ydiag, int diag_off, float* diag_corr, float* dX, int ldd, double* scal,
                    int nseg_hint, cudaStream_t stream);

// logits_bwd3.cu: 64-row CTA pairs, whole output width in TMEM (Kp == Dp in {256, 512, 768}, hp == 0); B2_ENOSYS otherwise
int logits_bwd_pair64(int mode, const void* X, const void* Y, int Nx, int Ny, int Kp, int Dp, int D, int hi_off, int ldx,
                      int ldy, float scale2, float shift2, float inv_tau, float bias, float wneg_c,
                      const float* rowscale, const float* colscale, float out_scale, float gnorm, int hp,
                      const float* dyn, float ydiag, int diag_off, float* diag_corr, float* dX, int ldd, double* scal,
                      int nseg_hint, cudaStream_t stream, void* gstore = nullptr, long long g_elems = 0);
int logits_bwd_both(int mode, const void* X, const void* Y, int Nx, int Ny, int Kp, int Dp, int D, int ldx, int ldy,
                    float wneg_c, const float* rowscale, const float* colscale, float gnorm, const float* dyn, float ydiag,
                    int diag_off, float* diag_corr, float* dX, int ldd, float* dY, int lddy, double* scal, void* G,
                    long long g_elems, cudaStream_t stream);
// gt_gemm.cu: dY += dyn[2] / gnorm * G^T X from the stored bf16 gradient tiles (blocked layout, gstore_elems() elements)
long long gstore_elems(int Nx, int Ny);
int gt_gemm(const void* G, long long g_elems, int Nx, int Ny, const void* X, int ldx, int Dp, int D, const float* dyn,
            float gnorm, float* dY, int ldd, cudaStream_t stream);
int gt_gemm_multi(const void* G, long long g_elems_total, int nsrc, const int* g_blk_off, int Nx, int Ny, const void* X,
                  int ldx, int x_cols, const int* x_col_off, int Dp, int D, const float* dyn, float gnorm, float* dY, int ldd,
                  cudaStream_t stream);

// attnpool_mma.cu: 16-bit inputs on mma.sync (heads <= 8, D % 128 == 0, D <= 1024, 16-byte aligned rows)
bool attnpool_mma_ok(const void* x, int dtype, long long sb, long long sn, int D, int H);
int attnpool_fwd_mma(const void* x, int dtype, long long sb, long long sn, const unsigned char* mask, long long mb,
                     const float* qt, const float* w, long long wb, long long wh, int B, int N, int D, int H, int S,
                     float* part_m, float* part_l, float* part_acc, float drop_p, unsigned long long drop_seed,
                     float* part_l2, cudaStream_t s);

int attnpool_bwd_splits(int B, int N);
int attnpool_bwd_dx_mma(const void* x, int dtype, long long sb, long long sn, const unsigned char* mask, long long mb,
                        const float* qt, const float* dxbar, const float* xbar, const float* m, const float* l, int B,
                        int N, int D, int H, void* dx, float* ds, const float* sa, const float* dsa, float drop_p,
                        unsigned long long drop_seed, const float* dlse, cudaStream_t s, float* part_dq = nullptr);

// scalars.cu
int dyn_prep(const float* log_temp, const float* bias, float clamp_min, float bound, float* dyn, cudaStream_t s);
int dyn_set_siglip(float* dyn, float lclamp, float yneg, cudaStream_t s);
int dyn_set_stable(float* dyn, int stable, cudaStream_t s);
int lse_finalize(const float* sums, int n, const float* dyn, float c, float* scale_out, double* acc, cudaStream_t s);
int vec_fsum(const float* v, int n, int gated, double* acc, cudaStream_t s);
int clip_finalize(const float* sums, int n, int nvec, const float* dyn, float eps, int gated, const double* unif,
                  float* rowscale, float* colscale, float* loss_out, double* acc_out, cudaStream_t s);
int clip_finalize_peers(const float* const* peer_sums_host, int world, int n, int nvec, const float* dyn, float eps, int gated,
                        const double* unif, float* rowscale, float* colscale, float* loss_out, double* acc_out,
                        cudaStream_t s);
int clip_dlogtemp(const double* scal0, const float* dyn, const float* gmul, const double* unif, int n, float* out,
                  cudaStream_t s);
int diag_sum(const void* a, int lda, const void* b, int ldb, int rows, int K, int gated, float* dots, double* acc,
             cudaStream_t s);
int alignment_diag(const float* sums, int n, const float* dyn, int gated, float* out, cudaStream_t s);

// siglip.cu
int siglip_dense_fwd(const void* V, const void* T, int B, int Tn, int Kp, int ldv, int ldt, const float* dyn,
                     double* acc, cudaStream_t stream);
int siglip_entropy_rowsum(const void* V, const void* T, int B, int Tn, int Kp, int ldv, int ldt, const float* dyn,
                          float* Z, cudaStream_t stream);
int siglip_entropy_stats(const void* V, const void* T, int B, int Tn, int Kp, int ldv, int ldt, const float* dyn,
                         const float* Z, float* H, float* Q, cudaStream_t stream);
int siglip_entropy_rows(const float* Z, const float* H, const float* Q, int B, float* rowvec, double* stats,
                        cudaStream_t s);
int siglip_entropy_coef(const double* stats_all, int W, int Bg, int T, float weight, float thr, float* dyn, float* out,
                        cudaStream_t s);
int siglip_combine(const double* acc, double wn_c, const float* tinv, int T, double* red, cudaStream_t s);
int siglip_loss_out(const double* red, const int* overflow, const float* ent, int world, float* loss_out, float* diag,
                    cudaStream_t s);
int siglip_scalar_grads(const double* red, const float* dyn, const float* gmul, float* dlt, float* dbias,
                        cudaStream_t s);
int siglip_compact(const float* mask, long ldm, const float* pw, long ldw, int B, int T, int cap, int* col, float* y,
                   float* w, int* cnt, float* ysum, int* overflow, cudaStream_t s);
int siglip_pos(const void* V, int ldv, const void* T, int ldt, int K, int Dp, int D, int hi_off, int B, int Tn, int cap,
               const int* col, const float* y, const float* w, const int* cnt, const float* ysum, const float* dyn,
               float positive_weight, float negative_weight, float c, float gnorm, int hp, int use_pw, int auto_balance,
               float* dV, int lddv, float* dT, int lddt, double* acc, const void* Vraw, int v_dtype, long long ld_vraw,
               const float* vinv, const void* Traw, int t_dtype, long long ld_traw, const float* tinv, cudaStream_t s);

// retrieval.cu
int retrieval_segments(int Ma, int Nb);
int retrieval_sweep(const void* V, const void* T, int Nv, int Mt, int Kp, int ldv, int ldt, const float* sgt,
                    const long long* gt, int col_offset, int* counts, int k, int segs, float* part_score,
                    int* part_idx, cudaStream_t stream);
int retrieval_colmax(const void* V, const void* T, int Nv, int Mt, int Kp, int ldv, int ldt, int segs, float* part_max,
                     cudaStream_t stream);
int kth_largest(const float* vals, int rows, int cand, int k, float* thr, cudaStream_t s);
int retrieval_collect(const void* V, const void* T, int Nv, int Mt, int Kp, int ldv, int ldt, const float* thr,
                      int col_offset, int segs, int* cnt, float* buf_s, int* buf_i, int cap, int* overflow,
                      cudaStream_t stream);
int topk_merge(const float* ps, const int* pi, int rows, int cand, int k, float* out_s, long long* out_i,
               cudaStream_t s);
int recall_hits(const int* counts, int rows, const int* kvals, int nk, unsigned long long* hits, cudaStream_t s);
int mrr_from_counts(const int* counts, int rows, int n_bins, int* hist, double* out, cudaStream_t s);

// dense_metrics.cu
int dense_gt_ranks(const void* sim, int dtype, long long ld, int N, int M, const int* gt, int G, int sanitize,
                   int* ranks, cudaStream_t s);
int dense_rank_metrics(const int* ranks, const int* gsize, int N, int G, int M, const int* recall_k, int nrk,
                       const int* ndcg_k, int nnk, int* best, double* rr, double* ap, unsigned char* hit, double* ndcg,
                       cudaStream_t s);

// multipos.cu
int multipos_workspace_bytes(int N, int M);
int multipos_fwd(const float* L, long long ldl, const float* pw, const float* mk, long long ldw, int N, int M, int mode,
                 float eps, int reduce_sum, float* rstat, float* cstat, float* coef, float* loss_out, void* workspace,
                 cudaStream_t s);
int multipos_bwd(const float* L, long long ldl, const float* pw, const float* mk, long long ldw, int N, int M,
                 const float* rstat, const float* cstat, const float* coef, const float* gmul, float* dL, long long ldd,
                 cudaStream_t s);

// rope3d.cu
int rope3d_apply(const void* q, long long qsb, long long qsh, long long qsn, void* q_out, const void* k, long long ksb,
                 long long ksh, long long ksn, void* k_out, const void* sin_t, const void* cos_t, int dtype, int B,
                 int Hh, int N, int Dh, int backward, cudaStream_t s);

// attnpool.cu
int attnpool_splits(int B, int N);
int attnpool_fwd(const void* x, int dtype, long long sb, long long sn, const unsigned char* mask, long long mb,
                 const float* qt, const float* w, long long wb, long long wh, int B, int N, int D, int H, int S,
                 float* part_m, float* part_l, float* part_acc, float drop_p, unsigned long long drop_seed,
                 float* part_l2, cudaStream_t s);
int attnpool_merge(const float* part_m, const float* part_l, const float* part_acc, int B, int S, int H, int D,
                   float* out, float* out_m, float* out_l, int sum_over_b, const float* part_l2, float* out_sa,
                   cudaStream_t s);
int attnpool_bwd_dx(const void* x, int dtype, long long sb, long long sn, const unsigned char* mask, long long mb,
                    const float* qt, const float* dxbar, const float* xbar, const float* m, const float* l, int B, int N,
                    int D, int H, void* dx, float* ds, const float* sa, const float* dsa, float drop_p,
                    unsigned long long drop_seed, const float* dlse, cudaStream_t s);
int attnpool_bwd_dx_dq(const void* x, int dtype, long long sb, long long sn, const unsigned char* mask, long long mb,
                       const float* qt, const float* dxbar, const float* xbar, const float* m, const float* l, int B, int N,
                       int D, int H, void* dx, float* ds, const float* sa, const float* dsa, float drop_p,
                       unsigned long long drop_seed, const float* dlse, float* part_dq, cudaStream_t s);

// attnpool_tc.cu: 16-bit contiguous x, heads <= 8, D % 128 == 0, D <= 512 on tcgen05 tiles (TMEM accumulators)
bool attnpool_tc_ok(const void* x, int dtype, long long sb, long long sn, int N, int D, int H);
int attnpool_tc_splits(int B, int N);
int attnpool_tc_fwd(const void* x, int dtype, const unsigned char* mask, long long mb, const float* qt, const void* qt_img,
                    int B, int N, int D, int H, int S, float* part_m, float* part_l, float* part_acc, float drop_p,
                    unsigned long long drop_seed, float* part_l2, cudaStream_t s);
int attnpool_tc_bwd(const void* x, int dtype, const unsigned char* mask, long long mb, const float* qt, const float* dxbar,
                    const float* xbar, const void* w_img, const float* cdot, const float* m, const float* l, int B, int N,
                    int D, int H, int S, void* dx, const float* sa, const float* dsa, float drop_p,
                    unsigned long long drop_seed, const float* dlse, float* part_dq, cudaStream_t s);

// pooltail.cu: the [B, D]-vector work around the streaming pool kernels (fp32 parameters, D % 128 == 0, D <= 512, heads | 8)
bool pooltail_ok(int D, int H, int Do);
int pool_prep(const float* query, const float* w_in, const float* b_in, int D, int H, float* q0, float* qt, void* qt_img,
              int fp16, cudaStream_t s);
int pool_tail_fwd(const float* pm, const float* pl, const float* pl2, const float* pa, int B, int S, int H, int D,
                  const float* w_v, const float* b_v, const float* w_o, const float* b_o, const float* gamma,
                  const float* beta, float eps, const float* w_p, const float* b_p, int Do, float* xbar, float* m, float* l,
                  float* sa, float* o, float* yhat, float* rstd, float* yln, void* out, int out_dtype, cudaStream_t s);
int pool_tail_bwd(const void* dout, int dout_dtype, const float* yhat, const float* rstd, const float* xbar, const float* sa,
                  const float* w_v, const float* b_v, const float* w_o, const float* gamma, const float* w_p, int Do,
                  const float* qt, int B, int H, int D, float* dyln, float* dy, float* do_, float* dxbar, float* dsa,
                  float* cdot, void* w_img, int fp16, cudaStream_t s);
int pool_param_grads(const float* dy, const float* o, const float* do_, const float* xbar, const float* sa, int use_sa,
                     const float* dyln, const float* yhat, const void* dout, int dout_dtype, const float* yln, int Do, int B,
                     int H, int D, float* dw_o, float* db_o, float* dw_v, float* db_v, float* dgamma, float* dbeta, float* dw_p,
                     float* db_p, cudaStream_t s);
int pool_qgrads(const float* part_dq, int nparts, const float* q0, const float* query, const float* w_in, int H, int D,
                float* dqt, float* dw_in, float* db_in, float* dquery, cudaStream_t s);

// inline_mp.cu: the runner's inline multi-positive branch from features (fp32 rows, D <= 1024)
int inline_mp_fwd(const float* v, long long ldv, const float* t, long long ldt, const float* targets, const float* pw,
                  long long ldm, const float* abn, float margin, const float* log_temp, int B, int M, int D, int mode,
                  float eps, float neg_w, float* rstat, float* cstat, float* scal, int* flag, cudaStream_t s);
int inline_mp_bwd(const float* v, long long ldv, const float* t, long long ldt, const float* targets, const float* pw,
                  long long ldm, const float* abn, float margin, const float* log_temp, int B, int M, int D, int mode,
                  float eps, float neg_w, const float* rstat, const float* cstat, const float* scal, const int* flag,
                  const float* gout, float* dv, float* dt, double* dlt_acc, cudaStream_t s);

// xfblock.cu: pre-LN transformer block over the views of a study (fp32, N <= 16, D <= 512, F = 4D <= 2048)
bool xfblock_ok(int N, int D, int H, int F);
int xfblock(int backward, const void* const* ptrs, int B, int N, int D, int H, int F, float eps1, float eps2, float drop_p,
            unsigned long long seed, long long mask_sb, cudaStream_t s);
int xfblock_wgrad(const float* a, long long lda, const float* bm, long long ldb, float* dw, float* db, int J, int I, int R,
                  const float* a2, const float* xh, float* dg, float* dbeta, int D2, cudaStream_t s);

void aggregator_sizes(int B, int N, int D, int H, int F, long long* sizes);
int aggregator(int backward, const void* const* ptrs, int depth, int B, int N, int D, int H, int F, const float* eps,
               float drop_p, const long long* seeds, long long mask_sb, long long x_sb, long long x_sn, int pos_rows,
               cudaStream_t s);

// milpool.cu: gated-attention pooling of the multi-instance probing head (fp32)
bool milpool_ok(int L, int D, int Hd);
void milpool_plan(int S, int L, int D, int Hd, int* plan);
int milpool_fwd(const float* x, long long sx_seq, long long sx_tok, const uint8_t* mask, long long smask, const float* V,
                const float* bV, const float* U, const float* bU, const float* w, const float* bw, int S, int L, int D,
                int Hd, float drop_p, unsigned long long seed, float* tg, float* spart, float* attn, float* opart, float* out,
                cudaStream_t s);
int milpool_bwd(const float* x, long long sx_seq, long long sx_tok, const float* V, const float* U, const float* w, int S,
                int L, int D, int Hd, float drop_p, unsigned long long seed, const float* tg, const float* attn,
                const float* dout, float* ds, float* dx, float* dpre, float* wpart, float* fpart, float* dW, float* dsmall,
                cudaStream_t s);

bool milpool_tc_ok(long long R, int L, int D, int Hd);
void milpool_tc_plan(int S, int L, int D, int Hd, long long* plan);
int milpool_tc_fwd(const float* x, long long sx_seq, long long sx_tok, const uint8_t* mask, long long smask, const float* V,
                   const float* bV, const float* U, const float* bU, const float* w, const float* bw, int S, int L, int D,
                   int Hd, float drop_p, unsigned long long seed, void* x3, void* w3, void* wt3, float* tg, float* spart,
                   float* attn, float* opart, float* out, cudaStream_t s);
int milpool_tc_bwd(const float* x, long long sx_seq, long long sx_tok, const float* w, int S, int L, int D, int Hd,
                   float drop_p, unsigned long long seed, const void* x3, const void* wt3, const float* tg, const float* attn,
                   const float* dout, float* ds, float* dx, void* dpre3, void* ghi, void* glo, float* ad, float* fpart,
                   float* dW, float* dsmall, const float* one3, cudaStream_t s);

// scalars.cu: symmetric-memory plumbing of the multi-GPU CLIP path
int symm_barrier(void* const* flags_host, int world, int rank, int channel, cudaStream_t s);
int symm_allreduce_f32(void* const* bufs_host, long long n, int world, int rank, cudaStream_t s);
int symm_sum_f64(const void* const* peers_host, int n, int world, double* out, cudaStream_t s);
int clip_dlogtemp_peers(const void* const* scal_host, int world, const float* dyn, const float* gmul, const double* unif, int n,
                        float* out, cudaStream_t s);

int l2norm_fwd_mc(const void* x, int dtype, long ldx, int rows, int dim, void* mc_out, long long row_offset, int ldo, int Kp,
                  float* inv_norm, int normalize, cudaStream_t s);

// querypool.cu
int querypool(int backward, const float* x, long long sb, long long sn, const float* pos, const float* lnw,
              const float* lnb, const float* q, const unsigned char* mask, long long mb, int B, int N, int D, float eps,
              float* out, const float* dout, float* dx, float* dpos, float* dlnw, float* dlnb, float* dq,
              cudaStream_t s);

// tmap.cu
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                      uint32_t box_rows);
int make_tmap_bf16_rows64(CUtensorMap* out, const void* base, uint64_t rows, uint32_t box_rows);
int sm_count();
int current_device();   // tmap.cu

}  // namespace b2host
