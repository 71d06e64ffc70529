// Scalar tail of the per-step alignment diagnostics (scalars.cu / b200clip_alignment_diag). Plain CUDA C++ with no other
// include, so that tests/emul/ can compile this very file for the host under the emulation shim.
#pragma once

namespace b2 {

// Alignment diagnostics of one batch from the forward statistics (the runner recomputes a dense [B, B] similarity and its
// log-softmax after every step just to log them: runners/video_constrative_learning_runner.py:1323-1335):
//   sums = [colsum (n) | rowsum (n) | S_ii (n)] as produced by logits_lse_fwd;
//   out[0] = mean S_ii                                     (alignment_cosine: diag(similarity).mean())
//   out[1] = mean (f(S_ii) / tau - ln rowsum_i - ln2 shift2)  (alignment_logprob: diag(log_softmax(logits, 1)).mean())
//   out[2] = exp(out[1])                                   (alignment_prob)
// Stable mode (dyn[11] != 0, temperatures below the fixed-shift window): the rowsum slot holds the log2-domain row
// log-sum-exp written by logits_rowlse instead, and ln rowsum_i + ln2 shift2 is replaced by ln2 * lse2_i.
// One CTA, fp64 accumulation with a fixed reduction tree (deterministic).
__global__ void __launch_bounds__(1024)
alignment_diag_kernel(const float* __restrict__ sums, int n, const float* __restrict__ dyn, int gated,
                      float* __restrict__ out) {
  const double shift = (double)dyn[6], inv_tau = (double)dyn[2];
  const bool stable = dyn[11] != 0.f;
  double a_cos = 0.0, a_lp = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = (double)sums[2 * n + i];
    const double f = gated ? d / (1.0 + exp(-d)) : d;
    a_cos += d;
    a_lp += f * inv_tau - (stable ? 0.6931471805599453 * (double)sums[n + i] : (double)logf(sums[n + i]) + shift);
  }
  __shared__ double sh[2][1024];
  sh[0][threadIdx.x] = a_cos;
  sh[1][threadIdx.x] = a_lp;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double lp = sh[1][0] / n;
    out[0] = (float)(sh[0][0] / n);
    out[1] = (float)lp;
    out[2] = (float)exp(lp);
  }
}

}  // namespace b2
