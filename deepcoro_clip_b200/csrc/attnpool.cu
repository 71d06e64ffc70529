// K8: attention pooling with ONE learnable query (models/attention_pool.py:77-93), streaming form.
//
// nn.MultiheadAttention(query[B,1,D], x, x) materialises K = x Wk^T and V = x Wv^T ([B*N, D] x [D, D] GEMMs, 2x the
// bytes of x written and re-read). With a single query token the projections fold (exact):
//     s_hn = x_n . qt_h          qt_h = Wk_h^T q0_h / sqrt(Dh)   (the b_k term is constant over n and cancels)
//     a_h  = softmax_n(s_h)      xbar_h = sum_n a_hn x_n  in R^D ;   o_h = Wv_h xbar_h + b_v,h
// so the hot path is ONE pass over x producing xbar [B, H, D]; the O(B D^2) projections / LayerNorm around it stay
// ordinary dense ops on [B, D] vectors (host side, autograd). Kernels here:
//   pool_fwd   : warp = head; CTA = (batch row, token split); x tiles staged in smem by cp.async (double buffered),
//                online softmax per head, partial (m, l, acc[D]) per split.           reads x once.
//                The same kernel with given weights computes dqt_h = sum_n ds_hn x_n for the backward.
//   pool_merge : merges the splits -> xbar, m, l.
//   pool_bwd_dx: warp = token; recomputes a_hn, ds_hn = a_hn (dxbar_h . x_n - dxbar_h . xbar_h),
//                dx_n = sum_h a_hn dxbar_h + ds_hn qt_h; writes dx and ds.              reads x once, writes dx.
// key_padding_mask (True = ignore) -> score -inf. Dropout is not supported (eval / p = 0), as in the folded algebra.
#include "common.cuh"
#include "host_api.h"

#include <cuda_pipeline.h>

#define B2_DYN_SMEM16(name) extern __shared__ __align__(16) unsigned char name[]
#include "attnpool_kernels.cuh"

namespace b2host {
using namespace b2;

static int pick_splits(int B, int N) {
  const int sms = sm_count();
  int S = (2 * sms) / B;                 // B * S <= 2 CTAs per SM: one full wave of the 16-bit MMA kernel
  const int maxS = (N + 4 * AP_TOK - 1) / (4 * AP_TOK);
  if (S > maxS) S = maxS;
  if (S < 1) S = 1;
  if (S > 64) S = 64;
  return S;
}
int attnpool_splits(int B, int N) { return pick_splits(B, N); }

template <typename T>
static int pool_fwd_t(const PoolFwdParams& p, cudaStream_t s) {
  const int VE = 16 / (int)sizeof(T);
  if (p.D % (32 * VE)) return B2_EINVAL;
  const int nch = p.D / (32 * VE);
  const size_t smem = 2 * (size_t)AP_TOK * p.D * sizeof(T);
  dim3 grid(p.B, p.S), block(32 * p.H);
#define POOL_CASE(NC)                                                                                     \
  case NC: {                                                                                              \
    auto k = pool_fwd_kernel<T, NC>;                                                                      \
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)  \
      return B2_ECUDA;                                                                                    \
    k<<<grid, block, smem, s>>>(p);                                                                       \
    break;                                                                                                \
  }
  switch (nch) {
    POOL_CASE(1) POOL_CASE(2) POOL_CASE(3) POOL_CASE(4) POOL_CASE(6) POOL_CASE(8)
    default: return B2_EINVAL;
  }
#undef POOL_CASE
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int attnpool_fwd(const void* x, int dtype, long long sb, long long sn, const unsigned char* mask, long long mb,
                 const float* qt, const float* w, long long wb, long long wh, int B, int N, int D, int H, int S,
                 float* part_m, float* part_l, float* part_acc, float drop_p, unsigned long long drop_seed,
                 float* part_l2, cudaStream_t s) {
  if (!x || !part_acc || B <= 0 || N <= 0 || H <= 0 || H > 16 || S < 1 || (!qt && !w)) return B2_EINVAL;
  if (attnpool_mma_ok(x, dtype, sb, sn, D, H) && sb == (long long)N * sn)      // 16-bit inputs: tensor-core kernel (attnpool_mma.cu)
    return attnpool_fwd_mma(x, dtype, sb, sn, mask, mb, qt, w, wb, wh, B, N, D, H, S, part_m, part_l, part_acc, drop_p,
                            drop_seed, part_l2, s);
  PoolFwdParams p{x, sb, sn, mask, mb, qt, w, wb, wh, part_m, part_l, part_acc, B, N, D, H, S, drop_p, drop_seed, part_l2};
  switch (dtype) {
    case 0: return pool_fwd_t<float>(p, s);
    case 1: return pool_fwd_t<__nv_bfloat16>(p, s);
    case 2: return pool_fwd_t<__half>(p, s);
    default: return B2_EINVAL;
  }
}

int attnpool_merge(const float* part_m, const float* part_l, const float* part_acc, int B, int S, int H, int D,
                   float* out, float* out_m, float* out_l, int sum_over_b, const float* part_l2, float* out_sa,
                   cudaStream_t s) {
  if (!part_acc || !out) return B2_EINVAL;
  pool_merge_kernel<<<B * H, 256, 0, s>>>(part_m, part_l, part_acc, B, S, H, D, out, out_m, out_l, sum_over_b, part_l2,
                                          out_sa);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

template <typename T>
static int pool_bwd_t(const PoolBwdParams& p, cudaStream_t s) {
  const int VE = 16 / (int)sizeof(T);
  if (p.D % (32 * VE)) return B2_EINVAL;
  const int nch = p.D / (32 * VE);
  const size_t smem = (2 * (size_t)p.H * p.D + 3 * p.H) * sizeof(float);
  int gy = (2 * sm_count() + p.B - 1) / p.B;
  const int maxy = (p.N + 7) / 8;
  if (gy > maxy) gy = maxy;
  if (gy < 1) gy = 1;
  dim3 grid(p.B, gy);
#define POOL_CASE(NC)                                                                                     \
  case NC: {                                                                                              \
    auto k = pool_bwd_dx_kernel<T, NC>;                                                                   \
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)  \
      return B2_ECUDA;                                                                                    \
    k<<<grid, 256, smem, s>>>(p);                                                                         \
    break;                                                                                                \
  }
  switch (nch) {
    POOL_CASE(1) POOL_CASE(2) POOL_CASE(3) POOL_CASE(4) POOL_CASE(6) POOL_CASE(8)
    default: return B2_EINVAL;
  }
#undef POOL_CASE
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int attnpool_bwd_dx(const void* x, int dtype, long long sb, long long sn, const unsigned char* mask, long long mb,
                    const float* qt, const float* dxbar, const float* xbar, const float* m, const float* l, int B, int N,
                    int D, int H, void* dx, float* ds, const float* sa, const float* dsa, float drop_p,
                    unsigned long long drop_seed, const float* dlse, cudaStream_t s) {
  if (!x || !qt || !dxbar || !xbar || !m || !l || !dx || !ds || H > 16) return B2_EINVAL;
  if (drop_p < 0.f || drop_p >= 1.f || (drop_p > 0.f && (!sa || !dsa))) return B2_EINVAL;
  if (attnpool_mma_ok(x, dtype, sb, sn, D, H) && sb == (long long)N * sn && (reinterpret_cast<uintptr_t>(dx) % 16) == 0)
    return attnpool_bwd_dx_mma(x, dtype, sb, sn, mask, mb, qt, dxbar, xbar, m, l, B, N, D, H, dx, ds, sa, dsa, drop_p,
                               drop_seed, dlse, s);
  PoolBwdParams p{x, sb, sn, mask, mb, qt, dxbar, xbar, m, l, dx, ds, B, N, D, H, sa, dsa, drop_p, drop_seed, dlse};
  switch (dtype) {
    case 0: return pool_bwd_t<float>(p, s);
    case 1: return pool_bwd_t<__nv_bfloat16>(p, s);
    case 2: return pool_bwd_t<__half>(p, s);
    default: return B2_EINVAL;
  }
}

// attnpool_bwd_dx with the query gradient fused into the same pass over x (16-bit x on the MMA kernels only): part_dq
// [B, attnpool_bwd_splits(B, N), H, D] zeroed by the caller, summed afterwards by attnpool_merge(sum_over_b = 1).
int attnpool_bwd_dx_dq(const void* x, int dtype, long long sb, long long sn, const unsigned char* mask, long long mb,
                       const float* qt, const float* dxbar, const float* xbar, const float* m, const float* l, int B, int N,
                       int D, int H, void* dx, float* ds, const float* sa, const float* dsa, float drop_p,
                       unsigned long long drop_seed, const float* dlse, float* part_dq, cudaStream_t s) {
  if (!x || !qt || !dxbar || !xbar || !m || !l || !dx || !ds || !part_dq || H > 16) return B2_EINVAL;
  if (drop_p < 0.f || drop_p >= 1.f || (drop_p > 0.f && (!sa || !dsa))) return B2_EINVAL;
  if (!(attnpool_mma_ok(x, dtype, sb, sn, D, H) && sb == (long long)N * sn && (reinterpret_cast<uintptr_t>(dx) % 16) == 0))
    return B2_ENOSYS;
  return attnpool_bwd_dx_mma(x, dtype, sb, sn, mask, mb, qt, dxbar, xbar, m, l, B, N, D, H, dx, ds, sa, dsa, drop_p,
                             drop_seed, dlse, s, part_dq);
}

}  // namespace b2host
