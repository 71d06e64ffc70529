// K1 / K4: row L2-normalise forward (+ bf16 operand packing) and backward.
//   forward : xhat = x / max(||x||, 1e-12)   (F.normalize, utils/loss/contrastive.py:146-147)
//             writes the bf16 MMA operand row [Kp] (zero padded to a multiple of 64) and 1/max(||x||,eps);
//             split3 != 0 additionally writes the error-compensated K-concatenated operand
//             [hi | second | third] so that  A3 . B3 = hi.hi + hi.lo + lo.hi  (bf16x3, ~fp32 accuracy):
//             role 0 (A side): [hi | hi | lo], role 1 (B side): [hi | lo | hi].
//   backward: dx = (g - (g . xhat) xhat) * inv_norm, with g = scale * dxhat + diag_coef * other_hat
//             (the analytic diagonal / label-smoothing terms of the CLIP gradient are folded in here so the
//             tile kernels never special-case the diagonal). SURVEY Appendix A.1.
// One warp per row, 16-byte vector loads, fp32 math, warp-shuffle reductions. HBM-bound.
#include "common.cuh"

namespace b2 {

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T>
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const T* __restrict__ x, long ldx, int rows, int dim, __nv_bfloat16* __restrict__ out, int ldo,
                  int Kp, int split3_role, float* __restrict__ inv_norm, float* __restrict__ xhat_f32, int ldh) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const T* xr = x + (size_t)warp * ldx;
  float ss = 0.f;
  for (int c = lane; c < dim; c += 32) {
    const float v = to_f32<T>(xr[c]);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  if (lane == 0 && inv_norm) inv_norm[warp] = inv;
  __nv_bfloat16* o = out + (size_t)warp * ldo;
  for (int c = lane; c < Kp; c += 32) {
    const float v = c < dim ? to_f32<T>(xr[c]) * inv : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    if (xhat_f32 && c < dim) xhat_f32[(size_t)warp * ldh + c] = v;
    if (split3_role < 0) {
      o[c] = hi;
    } else {
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      o[c] = hi;
      o[Kp + c] = split3_role == 0 ? hi : lo;
      o[2 * Kp + c] = split3_role == 0 ? lo : hi;
    }
  }
}

// dx[r, :] = (g - (g . xh) xh) * inv_norm[r],  xh = x[r, :] * inv_norm[r] (exact fp32 from the caller's input),
//   g = gmul * ( gscale * dxh[r, :] + (ocoef * omul * fp_r) * oth[r, :] + (ucoef * omul) * usum[:] )
// omul / gmul are optional DEVICE scalars (1/tau from dyn_prep, upstream grad_output); fp_r = f'(dots[r]) for the
// gated variant (f(s) = s sigmoid(s)), 1 otherwise. oth is the bf16 hi panel of the partner operand.
template <typename T>
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ dxh, int ldg, const T* __restrict__ x, long ldx,
                  const float* __restrict__ inv_norm, const __nv_bfloat16* __restrict__ oth, int ldoth, int oth_rows,
                  const float* __restrict__ usum, const float* __restrict__ dots, int gated, float gscale, float ocoef,
                  float ucoef, const float* __restrict__ dev_omul, const float* __restrict__ dev_gmul, int rows, int dim,
                  float* __restrict__ dx, long lddx) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float omul = dev_omul ? dev_omul[0] : 1.f;
  const float gmul = dev_gmul ? dev_gmul[0] : 1.f;
  float fpr = 1.f;
  if (gated && dots) {
    const float sdot = dots[warp];
    const float sig = 1.f / (1.f + __expf(-sdot));
    fpr = sig * (1.f + sdot * (1.f - sig));
  }
  const bool has_oth = oth != nullptr && warp < oth_rows && ocoef != 0.f;
  const float oc = ocoef * omul * fpr, uc = ucoef * omul;
  const float* g = dxh + (size_t)warp * ldg;
  const T* xr = x + (size_t)warp * ldx;
  const float inv = inv_norm[warp];
  float dot = 0.f;
  for (int c = lane; c < dim; c += 32) {
    float gv = gscale * g[c];
    if (has_oth) gv = fmaf(oc, __bfloat162float(oth[(size_t)warp * ldoth + c]), gv);
    if (usum) gv = fmaf(uc, usum[c], gv);
    dot = fmaf(gv, to_f32<T>(xr[c]) * inv, dot);
  }
  dot = warp_sum(dot);
  for (int c = lane; c < dim; c += 32) {
    float gv = gscale * g[c];
    if (has_oth) gv = fmaf(oc, __bfloat162float(oth[(size_t)warp * ldoth + c]), gv);
    if (usum) gv = fmaf(uc, usum[c], gv);
    const float xv = to_f32<T>(xr[c]) * inv;
    dx[(size_t)warp * lddx + c] = gmul * (gv - dot * xv) * inv;
  }
}

// out[c] = sum_r xh[r, c]  (column sum of the bf16 operand; used by label smoothing) — tiny.
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const __nv_bfloat16* __restrict__ xh, int ld, int rows, int dim, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= dim) return;
  float s = 0.f;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) s += __bfloat162float(xh[(size_t)r * ld + c]);
  atomicAdd(out + c, s);
}

// out[r] = a[r, :] . b[idx ? idx[r] : r, :]   (bf16 operands, fp32 accumulate; K columns)
__global__ void __launch_bounds__(256)
rowdot_bf16_kernel(const __nv_bfloat16* __restrict__ a, int lda, const __nv_bfloat16* __restrict__ b, int ldb,
                   const long long* __restrict__ idx, int rows, int b_rows, int K, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  long long br = idx ? idx[warp] : warp;
  float s = 0.f;
  if (br >= 0 && br < b_rows) {
    const __nv_bfloat162* ar = reinterpret_cast<const __nv_bfloat162*>(a + (size_t)warp * lda);
    const __nv_bfloat162* bp = reinterpret_cast<const __nv_bfloat162*>(b + (size_t)br * ldb);
    for (int c = lane; c < K / 2; c += 32) {
      const float2 av = __bfloat1622float2(ar[c]);
      const float2 bv = __bfloat1622float2(bp[c]);
      s = fmaf(av.x, bv.x, s);
      s = fmaf(av.y, bv.y, s);
    }
  }
  s = warp_sum(s);
  if (lane == 0) out[warp] = s;
}

}  // namespace b2

namespace b2host {
using namespace b2;

int l2norm_fwd(const void* x, int dtype, long ldx, int rows, int dim, void* out, int ldo, int Kp, int split3_role,
               float* inv_norm, float* xhat_f32, int ldh, cudaStream_t s) {
  if (rows <= 0 || dim <= 0 || Kp < dim || Kp % 64) return B2_EINVAL;
  const int blocks = (rows + 7) / 8;
  auto o = reinterpret_cast<__nv_bfloat16*>(out);
  switch (dtype) {
    case 0: l2norm_fwd_kernel<float><<<blocks, 256, 0, s>>>((const float*)x, ldx, rows, dim, o, ldo, Kp, split3_role, inv_norm, xhat_f32, ldh); break;
    case 1: l2norm_fwd_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)x, ldx, rows, dim, o, ldo, Kp, split3_role, inv_norm, xhat_f32, ldh); break;
    case 2: l2norm_fwd_kernel<__half><<<blocks, 256, 0, s>>>((const __half*)x, ldx, rows, dim, o, ldo, Kp, split3_role, inv_norm, xhat_f32, ldh); break;
    default: return B2_EINVAL;
  }
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int l2norm_bwd(const float* dxh, int ldg, const void* x, int dtype, long ldx, const float* inv_norm, const void* oth,
               int ldoth, int oth_rows, const float* usum, const float* dots, int gated, float gscale, float ocoef,
               float ucoef, const float* dev_omul, const float* dev_gmul, int rows, int dim, float* dx, long lddx,
               cudaStream_t s) {
  if (rows <= 0 || dim <= 0) return B2_EINVAL;
  const int blocks = (rows + 7) / 8;
  auto o = (const __nv_bfloat16*)oth;
  switch (dtype) {
    case 0: l2norm_bwd_kernel<float><<<blocks, 256, 0, s>>>(dxh, ldg, (const float*)x, ldx, inv_norm, o, ldoth, oth_rows, usum, dots, gated, gscale, ocoef, ucoef, dev_omul, dev_gmul, rows, dim, dx, lddx); break;
    case 1: l2norm_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(dxh, ldg, (const __nv_bfloat16*)x, ldx, inv_norm, o, ldoth, oth_rows, usum, dots, gated, gscale, ocoef, ucoef, dev_omul, dev_gmul, rows, dim, dx, lddx); break;
    case 2: l2norm_bwd_kernel<__half><<<blocks, 256, 0, s>>>(dxh, ldg, (const __half*)x, ldx, inv_norm, o, ldoth, oth_rows, usum, dots, gated, gscale, ocoef, ucoef, dev_omul, dev_gmul, rows, dim, dx, lddx); break;
    default: return B2_EINVAL;
  }
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int colsum_bf16(const void* xh, int ld, int rows, int dim, float* out, cudaStream_t s) {
  if (rows <= 0 || dim <= 0) return B2_EINVAL;
  dim3 grid((dim + 255) / 256, rows < 256 ? 1 : 64);
  colsum_bf16_kernel<<<grid, 256, 0, s>>>((const __nv_bfloat16*)xh, ld, rows, dim, out);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int rowdot_bf16(const void* a, int lda, const void* b, int ldb, const long long* idx, int rows, int b_rows, int K,
                float* out, cudaStream_t s) {
  if (rows <= 0 || K <= 0 || (K & 1)) return B2_EINVAL;
  rowdot_bf16_kernel<<<(rows + 7) / 8, 256, 0, s>>>((const __nv_bfloat16*)a, lda, (const __nv_bfloat16*)b, ldb, idx,
                                                    rows, b_rows, K, out);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
