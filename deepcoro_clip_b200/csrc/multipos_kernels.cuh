// Kernels of the logits-level multi-positive softmax losses (multipos.cu). Plain CUDA C++ without inline PTX and without
// any project include, so that tests/emul/ can compile this very file for the host under the emulation shim.
#pragma once
#include <math.h>
#include <stddef.h>

namespace b2 {

// w_ij = max(0, pw_ij) [* mask_ij]; either pointer may be null (pw null: w = mask)
__device__ __forceinline__ float mp_weight(const float* pw, const float* mk, size_t off) {
  float w = pw ? pw[off] : 1.f;
  if (mk) w *= mk[off];
  return fmaxf(w, 0.f);
}

// RAW importance base of multi_positive_infonce.py:57-61: pos_weights as given (not masked, not clamped) or pos_mask
__device__ __forceinline__ float mp_base(const float* pw, const float* mk, size_t off) { return pw ? pw[off] : mk[off]; }

struct MpAcc {
  float m, s, a, p;
  int c;
  float q;          // sum of the raw importance base (use_importance_weighting)
  __device__ __forceinline__ void init() { m = -INFINITY; s = 0.f; a = 0.f; p = 0.f; c = 0; q = 0.f; }
  __device__ __forceinline__ void add(float l, float w, bool pos, float base = 0.f) {
    if (l > m) { s = s * __expf(m - l) + 1.f; m = l; }          // exp(-inf) = 0 on the first element
    else s += __expf(l - m);
    a = fmaf(w, l, a);
    p += w;
    c += pos ? 1 : 0;
    q += base;
  }
  __device__ __forceinline__ void merge(const MpAcc& o) {
    const float mm = fmaxf(m, o.m);
    if (mm != -INFINITY) s = s * __expf(m - mm) + o.s * __expf(o.m - mm);
    m = mm;
    a += o.a;
    p += o.p;
    c += o.c;
    q += o.q;
  }
};

__device__ __forceinline__ MpAcc mp_shfl_xor(const MpAcc& v, int o) {
  MpAcc r;
  r.m = __shfl_xor_sync(0xffffffffu, v.m, o);
  r.s = __shfl_xor_sync(0xffffffffu, v.s, o);
  r.a = __shfl_xor_sync(0xffffffffu, v.a, o);
  r.p = __shfl_xor_sync(0xffffffffu, v.p, o);
  r.c = __shfl_xor_sync(0xffffffffu, v.c, o);
  r.q = __shfl_xor_sync(0xffffffffu, v.q, o);
  return r;
}

// stat[i] = {lse, A, P, cnt}. One CTA per row.
__global__ void __launch_bounds__(256)
mp_row_stats_kernel(const float* __restrict__ L, long long ldl, const float* __restrict__ pw, const float* __restrict__ mk,
                    long long ldw, int N, int M, float4* __restrict__ stat, float* __restrict__ imp) {
  const int i = blockIdx.x;
  if (i >= N) return;
  MpAcc acc;
  acc.init();
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    const size_t ow = (size_t)i * ldw + j;
    acc.add(L[(size_t)i * ldl + j], mp_weight(pw, mk, ow), mk ? mk[ow] > 0.f : (pw ? pw[ow] > 0.f : false),
            imp ? mp_base(pw, mk, ow) : 0.f);
  }
  for (int o = 16; o > 0; o >>= 1) acc.merge(mp_shfl_xor(acc, o));
  __shared__ MpAcc sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) acc.merge(sh[w]);
    stat[i] = make_float4(acc.m + __logf(acc.s), acc.a, acc.p, (float)acc.c);
    if (imp) imp[i] = acc.q;
  }
}

// partial[chunk][j] over the rows of the chunk; thread = column (consecutive threads read consecutive addresses)
__global__ void __launch_bounds__(128)
mp_col_partial_kernel(const float* __restrict__ L, long long ldl, const float* __restrict__ pw,
                      const float* __restrict__ mk, long long ldw, int N, int M, int chunks, MpAcc* __restrict__ partial,
                      int want_imp) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int ch = blockIdx.y;
  if (j >= M) return;
  const int i0 = (int)((long long)N * ch / chunks), i1 = (int)((long long)N * (ch + 1) / chunks);
  MpAcc acc;
  acc.init();
  for (int i = i0; i < i1; ++i) {
    const size_t ow = (size_t)i * ldw + j;
    acc.add(L[(size_t)i * ldl + j], mp_weight(pw, mk, ow), mk ? mk[ow] > 0.f : (pw ? pw[ow] > 0.f : false),
            want_imp ? mp_base(pw, mk, ow) : 0.f);
  }
  partial[(size_t)ch * M + j] = acc;
}
__global__ void __launch_bounds__(128)
mp_col_merge_kernel(const MpAcc* __restrict__ partial, int M, int chunks, float4* __restrict__ stat,
                    float* __restrict__ imp) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  MpAcc acc = partial[j];
  for (int ch = 1; ch < chunks; ++ch) acc.merge(partial[(size_t)ch * M + j]);
  stat[j] = make_float4(acc.m + __logf(acc.s), acc.a, acc.p, (float)acc.c);
  if (imp) imp[j] = acc.q;
}

// One CTA: loss and the per-row / per-column gradient coefficients.
//   mode 0 (WeightedSigLIPLoss, eps): loss = 0.5 (mean_i l_i + mean_j l_j), l = -(A - lse P) / max(P, eps)
//   mode 1 (MultiPositiveInfoNCELoss): l = -(A - lse P) / max(P, 1) over rows / columns with cnt > 0, stacked;
//           reduction mean (reduce_sum = 0) or sum (1); no positives at all -> loss 0, zero gradient.
//   mode 2 (MultiPositiveInfoNCELoss(use_importance_weighting=True), multi_positive_infonce.py:57-93): the selected rows /
//           columns are weighted by r = max(sum of the raw importance base, FLT_EPSILON): loss = sum l r / max(sum r,
//           FLT_EPSILON) (mean) or sum l r (sum). On entry coef holds the importance sums written by the statistics kernels.
// coef[0..N) = gr_i, coef[N..N+M) = gc_j (multiply (softmax * P - w)).
__global__ void __launch_bounds__(1024)
mp_finalize_kernel(const float4* __restrict__ rstat, const float4* __restrict__ cstat, int N, int M, int mode, float eps,
                   int reduce_sum, float* __restrict__ coef, float* __restrict__ loss_out) {
  __shared__ double s_sum[32];
  __shared__ int s_cnt[32];
  __shared__ double tot_sum;
  __shared__ int tot_cnt;
  __shared__ double s_wsum[32];
  __shared__ double tot_w;
  double acc = 0.0, wacc = 0.0;
  int sel = 0;
  for (int t = threadIdx.x; t < N + M; t += blockDim.x) {
    const float4 st = t < N ? rstat[t] : cstat[t - N];
    const float den = mode == 0 ? fmaxf(st.z, eps) : fmaxf(st.z, 1.f);
    const bool on = mode == 0 ? true : st.w > 0.f;
    if (on) {
      double l = -((double)st.y - (double)st.x * (double)st.z) / (double)den;
      if (mode == 0) l *= t < N ? 0.5 / N : 0.5 / M;
      if (mode == 2) {
        const double r = (double)fmaxf(coef[t], 1.1920929e-7f);
        l *= r;
        wacc += r;
      }
      acc += l;
      ++sel;
    }
  }
  for (int o = 16; o > 0; o >>= 1) wacc += __shfl_xor_sync(0xffffffffu, wacc, o);
  if ((threadIdx.x & 31) == 0) s_wsum[threadIdx.x >> 5] = wacc;
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    sel += __shfl_xor_sync(0xffffffffu, sel, o);
  }
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = acc; s_cnt[threadIdx.x >> 5] = sel; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    int c = 0;
    double wsum = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s_sum[w]; c += s_cnt[w]; wsum += s_wsum[w]; }
    tot_sum = a;
    tot_cnt = c;
    tot_w = fmax(wsum, 1.1920929e-7);
    double loss = a;
    if (mode == 1) loss = c == 0 ? 0.0 : (reduce_sum ? a : a / c);
    if (mode == 2) loss = c == 0 ? 0.0 : (reduce_sum ? a : a / tot_w);
    loss_out[0] = (float)loss;
  }
  __syncthreads();
  const float scale1 = (mode == 1 && tot_cnt > 0) ? (reduce_sum ? 1.f : 1.f / (float)tot_cnt) : 0.f;
  const float scale2 = (mode == 2 && tot_cnt > 0) ? (reduce_sum ? 1.f : (float)(1.0 / tot_w)) : 0.f;
  for (int t = threadIdx.x; t < N + M; t += blockDim.x) {
    const float4 st = t < N ? rstat[t] : cstat[t - N];
    float g;
    if (mode == 0) g = (t < N ? 0.5f / N : 0.5f / M) / fmaxf(st.z, eps);
    else if (mode == 1) g = st.w > 0.f ? scale1 / fmaxf(st.z, 1.f) : 0.f;
    else g = st.w > 0.f ? scale2 * fmaxf(coef[t], 1.1920929e-7f) / fmaxf(st.z, 1.f) : 0.f;
    coef[t] = g;
  }
}

// dL_ij = gmul (gr_i (exp(L_ij - lse_i) P_i - w_ij) + gc_j (exp(L_ij - lse_j) Q_j - w_ij)); one CTA per row
__global__ void __launch_bounds__(256)
mp_backward_kernel(const float* __restrict__ L, long long ldl, const float* __restrict__ pw, const float* __restrict__ mk,
                   long long ldw, int N, int M, const float4* __restrict__ rstat, const float4* __restrict__ cstat,
                   const float* __restrict__ coef, const float* __restrict__ gmul, float* __restrict__ dL, long long ldd) {
  const int i = blockIdx.x;
  if (i >= N) return;
  const float4 rs = rstat[i];
  const float gr = coef[i] * gmul[0];
  const float gm = gmul[0];
  for (int j = threadIdx.x; j < M; j += blockDim.x) {
    const float l = L[(size_t)i * ldl + j];
    const float w = mp_weight(pw, mk, (size_t)i * ldw + j);
    const float4 cs = cstat[j];
    const float gc = coef[N + j] * gm;
    dL[(size_t)i * ldd + j] = gr * (__expf(l - rs.x) * rs.z - w) + gc * (__expf(l - cs.x) * cs.z - w);
  }
}

}  // namespace b2
