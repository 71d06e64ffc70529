// K9: multi-view query pooling — the tail of EnhancedVideoAggregator.forward (models/video_aggregator.py:119-123,
// 128-158): x + pos_encoding[:N] -> final LayerNorm -> masked rows := 0 -> scores = attn_query . x_n (no 1/sqrt(D))
// -> masked softmax over the N views (all-masked study -> uniform-over-valid fallback = zeros) -> weighted sum.
// N is tiny (<= 64 views): the reference launches 8 kernels over a few KB; here one CTA per study does everything
// out of shared memory, forward and backward (parameter gradients by atomics). Latency-bound by construction.
#include "common.cuh"
#include "host_api.h"

#define B2_DYN_SMEM_F32(name) extern __shared__ __align__(16) float name[]
#include "querypool_kernels.cuh"

namespace b2host {
using namespace b2;

int querypool(int backward, const float* x, long long sb, long long sn, const float* pos, const float* lnw,
              const float* lnb, const float* q, const unsigned char* mask, long long mb, int B, int N, int D, float eps,
              float* out, const float* dout, float* dx, float* dpos, float* dlnw, float* dlnb, float* dq,
              cudaStream_t s) {
  if (!x || !lnw || !lnb || !q || B <= 0 || N <= 0 || D <= 0) return B2_EINVAL;
  const size_t smem = ((size_t)N * D + 5 * (size_t)N) * sizeof(float);
  if (smem > 200 * 1024) return B2_EINVAL;
  QpParams p{x, sb, sn, pos, lnw, lnb, q, mask, mb, B, N, D, eps, out, dout, dx, dpos, dlnw, dlnb, dq};
  if (backward) {
    if (!dout || !dx || !dlnw || !dlnb || !dq) return B2_EINVAL;
    if (cudaFuncSetAttribute(querypool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return B2_ECUDA;
    querypool_kernel<true><<<B, 256, smem, s>>>(p);
  } else {
    if (!out) return B2_EINVAL;
    if (cudaFuncSetAttribute(querypool_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return B2_ECUDA;
    querypool_kernel<false><<<B, 256, smem, s>>>(p);
  }
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
