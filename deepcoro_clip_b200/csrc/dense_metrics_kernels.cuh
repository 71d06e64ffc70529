// Kernels of dense_metrics.cu (see there). No inline PTX and no include: CUDA types / intrinsics come from the including
// translation unit or from the host emulation (tests/emul/).
#pragma once

namespace b2 {

template <typename T> __device__ __forceinline__ float dm_val(T v);
template <> __device__ __forceinline__ float dm_val<float>(float v) { return v; }
template <> __device__ __forceinline__ float dm_val<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float dm_val<__half>(__half v) { return __half2float(v); }

// torch.nan_to_num(x, nan=0, posinf=1e4, neginf=-1e4) (compute_mrr, retrieval_metrics.py:118-120)
__device__ __forceinline__ float dm_sanitize(float v, int on) {
  if (!on) return v;
  if (v != v) return 0.f;
  if (v == INFINITY) return 1e4f;
  if (v == -INFINITY) return -1e4f;
  return v;
}

// (score descending, index ascending) as ONE unsigned 64-bit order: monotone map of the float bits in the high word
// (-0 canonicalised to +0 first), complemented column index in the low word. "column j ranks before ground truth g"
// <=> key(s_j, j) > key(s_g, g): two instructions per (element, threshold) instead of a float compare chain.
__device__ __forceinline__ unsigned long long dm_key(float v, int j) {
  const uint32_t u = __float_as_uint(v + 0.f);
  const uint32_t o = u ^ (uint32_t(int32_t(u) >> 31) | 0x80000000u);
  return (static_cast<unsigned long long>(o) << 32) | (0xFFFFFFFFu - (uint32_t)j);
}

// One CTA per row. gt [N, G] int32 (entries < 0 or >= M: absent). ranks [N, G] int32: 1-based rank, 0 = absent.
template <typename T, int GMAX>
__global__ void __launch_bounds__(256)
dense_gt_ranks_kernel(const T* __restrict__ sim, long long ld, int N, int M, const int* __restrict__ gt, int G,
                      int sanitize, int* __restrict__ ranks) {
  const int row = blockIdx.x;
  if (row >= N) return;
  const T* sr = sim + (size_t)row * ld;
  unsigned long long thr[GMAX];
#pragma unroll
  for (int g = 0; g < GMAX; ++g) {
    int c = g < G ? gt[(size_t)row * G + g] : -1;
    if (c >= M) c = -1;
    // absent entries: the maximal key, nothing ranks before it, the count stays 0
    thr[g] = c >= 0 ? dm_key(dm_sanitize(dm_val<T>(sr[c]), sanitize), c) : ~0ull;
  }
  int cnt[GMAX];
#pragma unroll
  for (int g = 0; g < GMAX; ++g) cnt[g] = 0;
  auto consider = [&](float v, int j) {
    const unsigned long long k = dm_key(dm_sanitize(v, sanitize), j);
#pragma unroll
    for (int g = 0; g < GMAX; ++g) cnt[g] += k > thr[g] ? 1 : 0;
  };
  // 16-byte vector loads over the aligned middle of the row, scalar head / tail
  constexpr int VE = 16 / (int)sizeof(T);
  const uintptr_t addr = reinterpret_cast<uintptr_t>(sr);
  int head = (int)(((16 - (addr & 15)) & 15) / sizeof(T));
  if (head > M) head = M;
  const int nvec = (M - head) / VE;
  for (int j = threadIdx.x; j < head; j += blockDim.x) consider(dm_val<T>(sr[j]), j);
  const uint4* vp = reinterpret_cast<const uint4*>(sr + head);
#pragma unroll 2
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const uint4 raw = __ldg(vp + v);
    const T* e = reinterpret_cast<const T*>(&raw);
    const int j0 = head + v * VE;
#pragma unroll
    for (int k = 0; k < VE; ++k) consider(dm_val<T>(e[k]), j0 + k);
  }
  for (int j = head + nvec * VE + threadIdx.x; j < M; j += blockDim.x) consider(dm_val<T>(sr[j]), j);
  __shared__ int sh[8][GMAX];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int g = 0; g < GMAX; ++g) {
    int v = cnt[g];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp][g] = v;
  }
  __syncthreads();
  if (threadIdx.x < G) {
    int v = 0;
    for (int w = 0; w < 8; ++w) v += sh[w][threadIdx.x];
    int c = gt[(size_t)row * G + threadIdx.x];
    ranks[(size_t)row * G + threadIdx.x] = (c >= 0 && c < M) ? v + 1 : 0;
  }
}

// Per-row metric terms from the ranks (one thread per row), arithmetic in double in the reference's operation order:
//   best[row]  = smallest rank, or M when the row has no ground truth in range (compute_median_rank :266-281)
//   rr[row]    = 1 / best or 0                                                  (compute_mrr :127-146)
//   ap[row]    = (1/H) sum_h h / r_(h), ranks ascending, H found items          (compute_map :305-322)
//   hit[row,k] = best <= min(k, M)                                              (compute_recall_at_k :77-99)
//   ndcg[row,k]= sum_{r <= k_eff} 1/log2(r + 1) / sum_{r < min(|set|, k_eff)} 1/log2(r + 2)   (compute_ndcg_at_k :214-242)
// gsize [N]: size of the row's ground-truth set (items >= M still count towards the ideal DCG, as in the reference).
__global__ void __launch_bounds__(256)
dense_rank_metrics_kernel(const int* __restrict__ ranks, const int* __restrict__ gsize, int N, int G, int M,
                          const int* __restrict__ recall_k, int nrk, const int* __restrict__ ndcg_k, int nnk,
                          int* __restrict__ best, double* __restrict__ rr, double* __restrict__ ap,
                          unsigned char* __restrict__ hit, double* __restrict__ ndcg) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= N) return;
  int r[16];
  int H = 0;
  for (int g = 0; g < G && g < 16; ++g) {
    const int v = ranks[(size_t)row * G + g];
    if (v > 0) {
      int p = H++;
      while (p > 0 && r[p - 1] > v) { r[p] = r[p - 1]; --p; }     // insertion sort, ascending
      r[p] = v;
    }
  }
  const int b = H > 0 ? r[0] : 0;
  best[row] = b > 0 ? b : M;
  rr[row] = b > 0 ? 1.0 / (double)b : 0.0;
  double psum = 0.0;
  for (int h = 0; h < H; ++h) psum += (double)(h + 1) / (double)r[h];
  ap[row] = H > 0 ? psum / (double)H : 0.0;
  for (int k = 0; k < nrk; ++k) {
    const int ku = recall_k[k] < M ? recall_k[k] : M;
    hit[(size_t)row * nrk + k] = (b > 0 && b <= ku) ? 1 : 0;
  }
  const int gs = gsize[row];
  for (int k = 0; k < nnk; ++k) {
    const int ke = ndcg_k[k] < M ? ndcg_k[k] : M;
    double dcg = 0.0;
    for (int h = 0; h < H; ++h)
      if (r[h] <= ke) dcg += 1.0 / log2((double)(r[h] + 1));
    const int ideal = gs < ke ? gs : ke;
    double idcg = 0.0;
    for (int q = 0; q < ideal; ++q) idcg += 1.0 / log2((double)(q + 2));
    ndcg[(size_t)row * nnk + k] = (gs > 0 && ideal > 0 && idcg > 0.0) ? dcg / idcg : 0.0;
  }
}

}  // namespace b2
