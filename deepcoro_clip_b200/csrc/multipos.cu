// Multi-positive softmax cross-entropy over MATERIALISED logits: the two losses the reference applies to an
// already computed [N, M] logits matrix —
//   WeightedSigLIPLoss       utils/loss/weighted_siglip.py:18-51   (runner's inline multi-positive branch, :1275-1283)
//   MultiPositiveInfoNCELoss utils/loss/multi_positive_infonce.py:8-100 (registry key "multi_positive_infonce")
// Both are  -sum_j w_ij log_softmax(L)_ij / denom_i  in the two directions, i.e. functions of six vectors only:
//   row:    lse_i = logsumexp_j L_ij,  A_i = sum_j w_ij L_ij,  P_i = sum_j w_ij,  cnt_i = #{j : mask_ij > 0}
//   column: lse_j, B_j, Q_j, cnt_j likewise
//   sum_j w_ij log_softmax(L)_ij = A_i - lse_i P_i.
// The reference materialises log_softmax twice, the transposes, and the weighted products (6 x [N, M] temporaries); here
// the statistics take ONE read of L / w per direction and the backward is one elementwise pass:
//   dL_ij = gr_i (exp(L_ij - lse_i) P_i - w_ij) + gc_j (exp(L_ij - lse_j) Q_j - w_ij).
// HBM-bound: 2 reads of (L, w) forward, 1 read of (L, w) + 1 write backward.
#include "common.cuh"
#include "host_api.h"
#include "multipos_kernels.cuh"

namespace b2host {
using namespace b2;

int multipos_chunks(int N, int M) {
  const int col_blocks = (M + 127) / 128;
  int ch = (4 * sm_count() + col_blocks - 1) / col_blocks;
  if (ch > N / 32) ch = N / 32;
  if (ch < 1) ch = 1;
  if (ch > 256) ch = 256;
  return ch;
}

int multipos_workspace_bytes(int N, int M) { return multipos_chunks(N, M) * M * (int)sizeof(MpAcc); }

int multipos_fwd(const float* L, long long ldl, const float* pw, const float* mk, long long ldw, int N, int M, int mode,
                 float eps, int reduce_sum, float* rstat, float* cstat, float* coef, float* loss_out, void* workspace,
                 cudaStream_t s) {
  if (N <= 0 || M <= 0 || mode < 0 || mode > 2 || (!pw && !mk)) return B2_EINVAL;
  const int ch = multipos_chunks(N, M);
  float* imp_r = mode == 2 ? coef : nullptr;          // importance sums travel in coef (finalize reads, then overwrites)
  float* imp_c = mode == 2 ? coef + N : nullptr;
  mp_row_stats_kernel<<<N, 256, 0, s>>>(L, ldl, pw, mk, ldw, N, M, reinterpret_cast<float4*>(rstat), imp_r);
  mp_col_partial_kernel<<<dim3((M + 127) / 128, ch), 128, 0, s>>>(L, ldl, pw, mk, ldw, N, M, ch,
                                                                 reinterpret_cast<MpAcc*>(workspace), mode == 2);
  mp_col_merge_kernel<<<(M + 127) / 128, 128, 0, s>>>(reinterpret_cast<const MpAcc*>(workspace), M, ch,
                                                      reinterpret_cast<float4*>(cstat), imp_c);
  mp_finalize_kernel<<<1, 1024, 0, s>>>(reinterpret_cast<const float4*>(rstat), reinterpret_cast<const float4*>(cstat), N,
                                        M, mode, eps, reduce_sum, coef, loss_out);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int multipos_bwd(const float* L, long long ldl, const float* pw, const float* mk, long long ldw, int N, int M,
                 const float* rstat, const float* cstat, const float* coef, const float* gmul, float* dL, long long ldd,
                 cudaStream_t s) {
  if (N <= 0 || M <= 0) return B2_EINVAL;
  mp_backward_kernel<<<N, 256, 0, s>>>(L, ldl, pw, mk, ldw, N, M, reinterpret_cast<const float4*>(rstat),
                                       reinterpret_cast<const float4*>(cstat), coef, gmul, dL, ldd);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
