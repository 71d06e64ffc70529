// Dense multi-label retrieval metrics (utils/retrieval_metrics.py:65-324): recall@k with ground-truth
// SETS, MRR, MAP, NDCG@k and median rank over a materialised similarity matrix [N, M].
//
// The reference argsorts every row (O(M log M)) and then walks Python loops over rows and ground-truth items. Every one
// of its metrics is a function of the RANKS of the row's ground-truth items only, and a rank needs no sort:
//     rank(i, g) = 1 + #{j : s_ij > s_ig} + #{j < g : s_ij == s_ig}        (score descending, lowest index first)
// so one streaming pass over the matrix (HBM-bound, each element read exactly once, up to 16 thresholds per row kept in
// registers) replaces the argsort, and a per-row kernel turns the <= 16 ranks into the per-row metric terms.
#include "common.cuh"
#include "host_api.h"

#include "dense_metrics_kernels.cuh"

namespace b2host {
using namespace b2;

template <typename T>
static int launch_ranks(const void* sim, long long ld, int N, int M, const int* gt, int G, int sanitize, int* ranks,
                        cudaStream_t s) {
  const T* p = reinterpret_cast<const T*>(sim);
  if (G <= 1) dense_gt_ranks_kernel<T, 1><<<N, 256, 0, s>>>(p, ld, N, M, gt, G, sanitize, ranks);
  else if (G <= 4) dense_gt_ranks_kernel<T, 4><<<N, 256, 0, s>>>(p, ld, N, M, gt, G, sanitize, ranks);
  else if (G <= 8) dense_gt_ranks_kernel<T, 8><<<N, 256, 0, s>>>(p, ld, N, M, gt, G, sanitize, ranks);
  else dense_gt_ranks_kernel<T, 16><<<N, 256, 0, s>>>(p, ld, N, M, gt, G, sanitize, ranks);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int dense_gt_ranks(const void* sim, int dtype, long long ld, int N, int M, const int* gt, int G, int sanitize,
                   int* ranks, cudaStream_t s) {
  if (N <= 0 || M <= 0 || G <= 0 || G > 16) return B2_EINVAL;
  if (dtype == 0) return launch_ranks<float>(sim, ld, N, M, gt, G, sanitize, ranks, s);
  if (dtype == 1) return launch_ranks<__nv_bfloat16>(sim, ld, N, M, gt, G, sanitize, ranks, s);
  if (dtype == 2) return launch_ranks<__half>(sim, ld, N, M, gt, G, sanitize, ranks, s);
  return B2_EINVAL;
}

int dense_rank_metrics(const int* ranks, const int* gsize, int N, int G, int M, const int* recall_k, int nrk,
                       const int* ndcg_k, int nnk, int* best, double* rr, double* ap, unsigned char* hit, double* ndcg,
                       cudaStream_t s) {
  if (N <= 0 || G <= 0 || G > 16 || M <= 0) return B2_EINVAL;
  dense_rank_metrics_kernel<<<(N + 255) / 256, 256, 0, s>>>(ranks, gsize, N, G, M, recall_k, nrk, ndcg_k, nnk, best, rr,
                                                           ap, hit, ndcg);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
