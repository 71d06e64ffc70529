// K5 / K6: streaming retrieval — similarity tiles on tcgen05 with a rank-count + top-k epilogue.
// Replaces the chunked matmul / topk / cat / topk / gather / argsort loops of
// utils/retrieval_metrics_streaming.py:47-101, 143-169. The [N, M] similarity matrix never exists.
//
//   tile order: outer = 128-row video tile (one TMEM lane = one video), inner = 256-wide text blocks.
//   per element (thread = row, 128 columns per tile per epilogue warpgroup):
//     * rank count: cnt += [s > s_gt] (+ [s == s_gt] for columns before the ground truth): lowest-index tie rule,
//       rank_i = 1 + cnt_i    -> recall@k and MRR need only these integer counts;
//     * top-k (optional, K <= 16 or <= 64): sorted list in REGISTERS per (row, warpgroup, sweep segment); a chunk is
//       scanned only if its maximum beats the current k-th score; columns arrive in increasing index, so a strict
//       ">" keeps the lowest index among equal scores.
//   Partial lists go to global memory; topk_merge (one warp per row) merges the slots (and, across GPUs, the
//   text shards) by (score desc, index asc).
#include "tile_engine2.cuh"
#include "host_api.h"
#include "retrieval_epi.cuh"

namespace b2host {
using namespace b2;

int make_shape(TeShape& g, int Ma, int Nb, int Kp);   // logits_fwd.cu
bool te_pair_enabled(int Kp);                          // logits_fwd.cu

template <class Epi>
static int launch_epi(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb,
                      const typename Epi::Params& p, int segs, cudaStream_t stream) {
  TeShape g;
  int rc = make_shape(g, Ma, Nb, Kp);
  if (rc) return rc;
  g.segs = segs;
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap_bf16_2d(&tmA, A, Ma, Kp, lda, TE_BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmB, B, Nb, Kp, ldb, TE_BN))) return rc;
  if (te_pair_enabled(Kp)) {
    // CTA pairs: A tile (video rows) resident per CTA, text blocks streamed as halves (tile_engine2.cuh)
    CUtensorMap tmB2;
    if ((rc = make_tmap_bf16_2d(&tmB2, B, Nb, Kp, ldb, 128))) return rc;
    auto kern2 = te2_kernel<Epi, false>;
    static bool attr2_done_dev[64] = {};
    bool& attr2_done = attr2_done_dev[current_device() & 63];   // cudaFuncSetAttribute is per device
    if (!attr2_done) {
      if (cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, TE2_SMEM_BYTES) != cudaSuccess)
        return B2_ECUDA;
      attr2_done = true;
    }
    const long long items2 = (long long)((g.m_tiles + 1) / 2) * segs;
    long long clusters = sm_count() / 2;
    if (items2 < clusters) clusters = items2;
    kern2<<<(int)(2 * clusters), TE_THREADS, TE2_SMEM_BYTES, stream>>>(tmA, tmB2, g, p);
    return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
  }
  auto kern = te_kernel<Epi, false>;
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device() & 63];   // cudaFuncSetAttribute is per device
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TE_SMEM_BYTES) != cudaSuccess)
      return B2_ECUDA;
    attr_done = true;
  }
  const long long items = (long long)g.m_tiles * segs;
  int grid = sm_count();
  if (items < grid) grid = (int)items;
  kern<<<grid, TE_THREADS, TE_SMEM_BYTES, stream>>>(tmA, tmB, g, p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

template <int kMaxK>
static int launch_retr(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, const RetrParams& p,
                       int segs, cudaStream_t stream) {
  return launch_epi<RetrEpi<kMaxK>>(A, B, Ma, Nb, Kp, lda, ldb, p, segs, stream);
}

int retrieval_segments(int Ma, int Nb) {
  const int m_tiles = (Ma + TE_BM - 1) / TE_BM, n_blocks = (Nb + TE_BN - 1) / TE_BN;
  const int sms = sm_count();
  int segs = 1;
  if (m_tiles < 4 * sms) {          // too few row tiles to fill the machine: split the text sweep
    segs = (4 * sms + m_tiles - 1) / m_tiles;
    if (segs > n_blocks) segs = n_blocks;
    if (segs > 64) segs = 64;
    if (segs < 1) segs = 1;
  }
  return segs;
}

int retrieval_sweep(const void* V, const void* T, int Nv, int Mt, int Kp, int ldv, int ldt, const float* sgt,
                    const long long* gt, int col_offset, int* counts, int k, int segs, float* part_score,
                    int* part_idx, cudaStream_t stream) {
  if (Nv <= 0 || Mt <= 0 || k < 0 || k > 64 || segs < 1) return B2_EINVAL;
  if (k > 0 && (!part_score || !part_idx)) return B2_EINVAL;
  if (sgt && (!gt || !counts)) return B2_EINVAL;
  RetrParams p{sgt, gt, col_offset, counts, part_score, part_idx, 2 * segs, k};
  if (k == 0) return launch_retr<0>(V, T, Nv, Mt, Kp, ldv, ldt, p, segs, stream);
  if (k <= 16) return launch_retr<16>(V, T, Nv, Mt, Kp, ldv, ldt, p, segs, stream);
  return launch_retr<64>(V, T, Nv, Mt, Kp, ldv, ldt, p, segs, stream);
}

int retrieval_colmax(const void* V, const void* T, int Nv, int Mt, int Kp, int ldv, int ldt, int segs, float* part_max,
                     cudaStream_t stream) {
  if (Nv <= 0 || Mt <= 0 || segs < 1 || !part_max || (reinterpret_cast<uintptr_t>(part_max) & 15)) return B2_EINVAL;
  ColMaxParams p{part_max, 2 * segs};
  return launch_epi<ColMaxEpi>(V, T, Nv, Mt, Kp, ldv, ldt, p, segs, stream);
}

int kth_largest(const float* vals, int rows, int cand, int k, float* thr, cudaStream_t s) {
  if (rows <= 0 || cand <= 0 || cand > (1 << 20) || k <= 0 || !vals || !thr) return B2_EINVAL;
  kth_largest_kernel<<<(rows + 7) / 8, 256, 0, s>>>(vals, rows, cand, k, thr);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int retrieval_collect(const void* V, const void* T, int Nv, int Mt, int Kp, int ldv, int ldt, const float* thr,
                      int col_offset, int segs, int* cnt, float* buf_s, int* buf_i, int cap, int* overflow,
                      cudaStream_t stream) {
  if (Nv <= 0 || Mt <= 0 || segs < 1 || cap < 1 || !thr || !cnt || !buf_s || !buf_i || !overflow) return B2_EINVAL;
  CollectParams p{thr, col_offset, cnt, buf_s, buf_i, cap, overflow};
  return launch_epi<CollectEpi>(V, T, Nv, Mt, Kp, ldv, ldt, p, segs, stream);
}

int topk_merge(const float* ps, const int* pi, int rows, int cand, int k, float* out_s, long long* out_i,
               cudaStream_t s) {
  if (rows <= 0 || cand <= 0 || cand > (1 << 20) || k <= 0) return B2_EINVAL;
  topk_merge_kernel<<<(rows + 7) / 8, 256, 0, s>>>(ps, pi, rows, cand, k, out_s, out_i);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int recall_hits(const int* counts, int rows, const int* kvals, int nk, unsigned long long* hits, cudaStream_t s) {
  if (rows <= 0 || nk <= 0) return B2_EINVAL;
  recall_hits_kernel<<<(rows + 255) / 256, 256, 0, s>>>(counts, rows, kvals, nk, hits);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int mrr_from_counts(const int* counts, int rows, int n_bins, int* hist, double* out, cudaStream_t s) {
  if (rows <= 0 || n_bins <= 0) return B2_EINVAL;
  rank_hist_kernel<<<(rows + 255) / 256, 256, 0, s>>>(counts, rows, n_bins, hist);
  rank_hist_mrr_kernel<<<1, 1024, 0, s>>>(hist, n_bins, out);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
