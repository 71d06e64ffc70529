// K8 for 16-bit inputs: the single-query attention pool (csrc/attnpool.cu) with the two skinny
// contractions on mma.sync m16n8k16 (N = 8 = the heads), so the kernel streams x at HBM speed instead of spending
// ~70 CUDA-core instructions and a 5-step shuffle reduction per (token, head).
//
//   forward, per 32-token tile staged in shared memory by cp.async (2 stages, XOR-swizzled 16-byte chunks):
//     phase 1  S[tok, h]   = x[tok, :] . qt[h, :]      A = x tile (ldmatrix), B = qt (bf16 hi + lo), K split 4 ways
//     softmax  warp h: online max / sum over the tile -> P[h, tok] (bf16 hi + lo), rescale factor per head
//     phase 2  acc[d, h]  += sum_tok x[tok, d] P[h, tok]   A = x tile TRANSPOSED (ldmatrix.trans), B = P
//   each warp owns D/8 output channels of the [D, 8] accumulator (fp32 registers) across the whole token split.
//   The "given weights" mode (dqt = sum_n ds_hn x_n of the backward) skips phase 1 and the softmax.
//
//   backward dx, per tile:  [S | T][tok, 16] = x . [qt ; dxbar_b]^T  (phase 1, two n8 MMAs per k-step),
//     a = exp(S - m)/l, ds = a (T - c)  ->  C[tok, 16] = [a | ds] (bf16 hi + lo)  and ds[B, H, N] (fp32, for dqt),
//     dx[tok, d] = C . [dxbar_b ; qt]  (phase 2, K = 16 = one MMA step), staged through shared memory, 16-byte stores.
//
// hi + lo bf16 splits of qt / dxbar / P / C keep the contractions at fp32-level accuracy for 16-bit x (the products
// x * hi and x * lo are exact in fp32). Legacy tensor path (HMMA) on purpose: N = 8 is below anything worth a TMEM
// round trip and the kernel is HBM-bound (~4 of ~11 us per tile are MMA issue).
#include "common.cuh"
#include "host_api.h"

#include "attnpool_mma_prims.cuh"
#include "attnpool_mma_kernels.cuh"

namespace b2host {
using namespace b2;

static int pm_stages(int D, size_t fixed, bool wide = false) {
  // latency-bound per CTA (3 barriers and two dependent MMA chains per tile): two resident CTAs per SM with a
  // 2-tile pipeline beat one CTA with a deep pipeline (measured 41 vs 57 us at C3 sizing); deepen only when a single
  // CTA fits anyway
  if (!wide && (size_t)2 * PM_TT * D * 2 + fixed <= 112 * 1024) return 2;   // 16-warp CTAs are alone on their SM (registers)
  int st = 4;
  while (st > 2 && (size_t)st * PM_TT * D * 2 + fixed > 220 * 1024) --st;
  return st;
}
static size_t pm_fwd_smem(int D, int stages) {
  return (size_t)stages * PM_TT * D * 2 + 2 * 8 * (D + 8) * 2 + 2 * 8 * (PM_TT + 8) * 2 + 8 * PM_TT * 8 * 4 + 8 * 4 + 64 + 2048;
}

bool attnpool_mma_ok(const void* x, int dtype, long long sb, long long sn, int D, int H) {
  return (dtype == 1 || dtype == 2) && H <= 8 && D % 128 == 0 && D <= 1024 &&
         (reinterpret_cast<uintptr_t>(x) % 16) == 0 && (sn % 8) == 0;   // + sb == N * sn, checked by the callers
}

int attnpool_fwd_mma(const void* x, int dtype, long long sb, long long sn, const unsigned char* mask, long long mb,
                     const float* qt, const float* w, long long wb, long long wh, int B, int N, int D, int H, int S,
                     float* part_m, float* part_l, float* part_acc, float drop_p, unsigned long long drop_seed,
                     float* part_l2, cudaStream_t s) {
  // forward: 8-warp CTAs, two per SM when the 2-tile pipeline fits (measured best at D = 512); 16 warps for wide rows
  const bool wide = D % 256 == 0 && (size_t)2 * PM_TT * D * 2 + pm_fwd_smem(D, 0) > 112 * 1024;
  const int stages = pm_stages(D, pm_fwd_smem(D, 0), wide);
  PmFwdParams p{x, sb, sn, mask, mb, qt, w, wb, wh, part_m, part_l, part_acc, B, N, D, H, S, drop_p, drop_seed, part_l2, stages};
  const size_t smem = pm_fwd_smem(D, stages);
  CUtensorMap tmx;
  {
    int rc = make_tmap_bf16_2d(&tmx, x, (uint64_t)B * N, D, sn, PM_TT);
    if (rc) return rc;
  }
  dim3 grid(B, S);
#define PM_LAUNCH(TT_, NW_)                                                                                          \
  {                                                                                                                  \
    auto k = pool_fwd_mma_kernel<TT_, NW_>;                                                                          \
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return B2_ECUDA; \
    k<<<grid, NW_ * 32, smem, s>>>(tmx, p);                                                                          \
  }
  if (dtype == 1) { if (wide) PM_LAUNCH(__nv_bfloat16, 16) else PM_LAUNCH(__nv_bfloat16, 8) }
  else { if (wide) PM_LAUNCH(__half, 16) else PM_LAUNCH(__half, 8) }
#undef PM_LAUNCH
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

static size_t pm_bwd_smem(int D, int stages) {
  return (size_t)stages * PM_TT * D * 2 + 4 * 8 * (D + 8) * 2 + (size_t)D * 48 + 2 * PM_TT * 16 * 2 + 8 * PM_TT * 16 * 4 + 3 * 8 * 4 +
         64 + 2048;
}

// token splits of the backward grid (B x S CTAs); part_dq of the fused-dq variant is [B, S, H, D]
int attnpool_bwd_splits(int B, int N) {
  if (B <= 0 || N <= 0) return 1;
  int S = sm_count() / B;
  const int maxS = (N + 2 * PM_TT - 1) / (2 * PM_TT);
  if (S > maxS) S = maxS;
  if (S < 1) S = 1;
  return S;
}

int attnpool_bwd_dx_mma(const void* x, int dtype, long long sb, long long sn, const unsigned char* mask, long long mb,
                        const float* qt, const float* dxbar, const float* xbar, const float* m, const float* l, int B,
                        int N, int D, int H, void* dx, float* ds, const float* sa, const float* dsa, float drop_p,
                        unsigned long long drop_seed, const float* dlse, cudaStream_t s, float* part_dq) {
  const int S = attnpool_bwd_splits(B, N);
  const size_t dq_smem = part_dq ? 2 * 8 * (PM_TT + 8) * 2 : 0;
  const int stages = pm_stages(D, pm_bwd_smem(D, 0) + dq_smem, D % 256 == 0);
  PmBwdParams p{x, sb, sn, mask, mb, qt, dxbar, xbar, m, l, dx, ds, B, N, D, H, S, sa, dsa, drop_p, drop_seed, stages, dlse,
                part_dq};
  const size_t smem = pm_bwd_smem(D, stages) + dq_smem;
  CUtensorMap tmx;
  {
    int rc = make_tmap_bf16_2d(&tmx, x, (uint64_t)B * N, D, sn, PM_TT);
    if (rc) return rc;
  }
  dim3 grid(B, S);
#define PM_LAUNCH(TT_, NW_)                                                                                          \
  {                                                                                                                  \
    auto k = pool_bwd_mma_kernel<TT_, NW_>;                                                                          \
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return B2_ECUDA; \
    k<<<grid, NW_ * 32, smem, s>>>(tmx, p);                                                                          \
  }
#define PM_LAUNCH_DQ(TT_, NW_)                                                                                       \
  {                                                                                                                  \
    auto k = pool_bwd_mma_kernel<TT_, NW_, true>;                                                                    \
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return B2_ECUDA; \
    k<<<grid, NW_ * 32, smem, s>>>(tmx, p);                                                                          \
  }
  const bool wide = D % 256 == 0;
  if (part_dq) {
    if (dtype == 1) { if (wide) PM_LAUNCH_DQ(__nv_bfloat16, 16) else PM_LAUNCH_DQ(__nv_bfloat16, 8) }
    else { if (wide) PM_LAUNCH_DQ(__half, 16) else PM_LAUNCH_DQ(__half, 8) }
  } else {
    if (dtype == 1) { if (wide) PM_LAUNCH(__nv_bfloat16, 16) else PM_LAUNCH(__nv_bfloat16, 8) }
    else { if (wide) PM_LAUNCH(__half, 16) else PM_LAUNCH(__half, 8) }
  }
#undef PM_LAUNCH
#undef PM_LAUNCH_DQ
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
