// K1 / K4: row L2-normalise forward (+ bf16 operand packing) and backward.
//   forward : xhat = x / max(||x||, 1e-12)   (F.normalize, utils/loss/contrastive.py:146-147)
//             writes the bf16 MMA operand row [Kp] (zero padded to a multiple of 64) and 1/max(||x||,eps);
//             split3 != 0 additionally writes the error-compensated K-concatenated operand
//             so that  A3 . B3 = lo.hi + hi.lo + hi.hi  (bf16x3, ~fp32 accuracy):
//             role 0 (A side): [lo | hi | hi], role 1 (B side): [hi | lo | hi].  The hi panel is LAST in both
//             (offset 2*Kp): the tensor core truncates its fp32 accumulator every K=16 step (measured bias
//             ~ -0.25 ulp per step), so the small correction terms are accumulated first, while the
//             accumulator is still small, and the dominant hi.hi term last.
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
                  int Kp, int split3_role, float* __restrict__ inv_norm, float* __restrict__ xhat_f32, int ldh,
                  int normalize) {
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
  // normalize == 0: pack the raw features (retrieval_metrics_streaming.py:35-41 does not normalise); inv_norm then
  // receives ||x|| itself (used for the *_norm metrics)
  const float nrm = sqrtf(ss);
  const float inv = normalize ? 1.f / fmaxf(nrm, 1e-12f) : 1.f;
  if (lane == 0 && inv_norm) inv_norm[warp] = normalize ? inv : nrm;
  __nv_bfloat16* o = out + (size_t)warp * ldo;
  for (int c = lane; c < Kp; c += 32) {
    const float v = c < dim ? to_f32<T>(xr[c]) * inv : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    if (xhat_f32 && c < dim) xhat_f32[(size_t)warp * ldh + c] = v;
    if (split3_role < 0) {
      o[c] = hi;
    } else {
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      o[c] = split3_role == 0 ? lo : hi;
      o[Kp + c] = split3_role == 0 ? hi : lo;
      o[2 * Kp + c] = hi;
    }
  }
}

// dx[r, :] = (g - (g . xh) xh) * inv_norm[r],  xh = x[r, :] * inv_norm[r] (exact fp32 from the caller's input),
//   g = gmul * ( gscale * dxh[r, :] + omul * (dc[r].res * yh + dc[r].gb * (yh - yhi)) + (ucoef * omul) * usum[:] )
// where yh = ox[r, :] * oinv[r] is the exact fp32 partner row, yhi its bf16 hi panel and dc = {res, gb} the
// diagonal correction written by logits_bwd: the tensor-core product used bf16(g_ii) * yhi for the target pair;
// res = g_ii - bf16(g_ii) and gb = bf16(g_ii) restore g_ii * yh exactly (the dominant, cancellation-prone term).
// omul / gmul are optional DEVICE scalars (1/tau from dyn_prep, upstream grad_output).
template <typename T, typename TO>
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ dxh, int ldg, const T* __restrict__ x, long ldx,
                  const float* __restrict__ inv_norm, const TO* __restrict__ ox, long ldox,
                  const float* __restrict__ oinv, const __nv_bfloat16* __restrict__ ohi, int ldohi,
                  const float2* __restrict__ dc, const float* __restrict__ usum, float gscale, float ucoef,
                  const float* __restrict__ dev_omul, const float* __restrict__ dev_gmul, int rows, int dim,
                  float* __restrict__ dx, long lddx) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float omul = dev_omul ? dev_omul[0] : 1.f;
  const float gmul = dev_gmul ? dev_gmul[0] : 1.f;
  const bool has_dc = dc != nullptr && ox != nullptr;
  float res = 0.f, gb = 0.f, oi = 0.f;
  if (has_dc) {
    const float2 d = dc[warp];
    res = d.x * omul;
    gb = d.y * omul;
    oi = oinv[warp];
  }
  const float uc = ucoef * omul;
  const float* g = dxh + (size_t)warp * ldg;
  const T* xr = x + (size_t)warp * ldx;
  const float inv = inv_norm[warp];
  auto gval = [&](int c) {
    float gv = gscale * g[c];
    if (has_dc) {
      const float yh = to_f32<TO>(ox[(size_t)warp * ldox + c]) * oi;
      const float yhi = __bfloat162float(ohi[(size_t)warp * ldohi + c]);
      gv = fmaf(res, yh, gv);
      gv = fmaf(gb, yh - yhi, gv);
    }
    if (usum) gv = fmaf(uc, usum[c], gv);
    return gv;
  };
  float dot = 0.f;
  for (int c = lane; c < dim; c += 32) dot = fmaf(gval(c), to_f32<T>(xr[c]) * inv, dot);
  dot = warp_sum(dot);
  for (int c = lane; c < dim; c += 32) {
    const float xv = to_f32<T>(xr[c]) * inv;
    dx[(size_t)warp * lddx + c] = gmul * (gval(c) - dot * xv) * inv;
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

// dst[r, :K] = src[idx[r], :K] (bf16 operand rows, 16-byte vectors; rows with an out-of-range index are zeroed)
__global__ void __launch_bounds__(256)
gather_rows_bf16_kernel(const uint4* __restrict__ src, int lds16, const long long* __restrict__ idx, int rows,
                        int src_rows, int k16, uint4* __restrict__ dst, int ldd16) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const long long r = idx[warp];
  const bool ok = r >= 0 && r < src_rows;
  for (int c = lane; c < k16; c += 32)
    dst[(size_t)warp * ldd16 + c] = ok ? src[(size_t)r * lds16 + c] : make_uint4(0u, 0u, 0u, 0u);
}

// flag[0] = 1 if any element of the row-major [rows, dim] matrix is not exactly representable in bf16 (one streaming
// read, grid-stride, no temporaries): the test behind precision="auto" of the streaming metrics. fp16 / fp32 inputs;
// a bf16 input is exact by construction (the host does not launch this for it).
template <typename T>
__global__ void __launch_bounds__(256)
inexact_bf16_kernel(const T* __restrict__ x, long long ld, int rows, int dim, int* __restrict__ flag) {
  bool bad = false;
  const long long total = (long long)rows * dim;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total && !bad; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / dim;
    const float v = to_f32<T>(x[r * ld + (i - r * dim)]);
    bad = __bfloat162float(__float2bfloat16_rn(v)) != v;                 // NaN compares unequal: counted as inexact
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicExch(flag, 1);
}

}  // namespace b2

namespace b2host {
using namespace b2;

int inexact_bf16(const void* x, int dtype, long long ld, int rows, int dim, int* flag, cudaStream_t s) {
  if (rows <= 0 || dim <= 0) return B2_EINVAL;
  const long long total = (long long)rows * dim;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  switch (dtype) {
    case 0: inexact_bf16_kernel<float><<<(int)blocks, 256, 0, s>>>((const float*)x, ld, rows, dim, flag); break;
    case 2: inexact_bf16_kernel<__half><<<(int)blocks, 256, 0, s>>>((const __half*)x, ld, rows, dim, flag); break;
    default: return B2_EINVAL;
  }
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int l2norm_fwd(const void* x, int dtype, long ldx, int rows, int dim, void* out, int ldo, int Kp, int split3_role,
               float* inv_norm, float* xhat_f32, int ldh, int normalize, cudaStream_t s) {
  if (rows <= 0 || dim <= 0 || Kp < dim || Kp % 64) return B2_EINVAL;
  const int blocks = (rows + 7) / 8;
  auto o = reinterpret_cast<__nv_bfloat16*>(out);
  switch (dtype) {
    case 0: l2norm_fwd_kernel<float><<<blocks, 256, 0, s>>>((const float*)x, ldx, rows, dim, o, ldo, Kp, split3_role, inv_norm, xhat_f32, ldh, normalize); break;
    case 1: l2norm_fwd_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)x, ldx, rows, dim, o, ldo, Kp, split3_role, inv_norm, xhat_f32, ldh, normalize); break;
    case 2: l2norm_fwd_kernel<__half><<<blocks, 256, 0, s>>>((const __half*)x, ldx, rows, dim, o, ldo, Kp, split3_role, inv_norm, xhat_f32, ldh, normalize); break;
    default: return B2_EINVAL;
  }
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

template <typename T>
static int l2norm_bwd_t(const float* dxh, int ldg, const T* x, long ldx, const float* inv_norm, const void* ox,
                        int odtype, long ldox, const float* oinv, const __nv_bfloat16* ohi, int ldohi, const float2* dc,
                        const float* usum, float gscale, float ucoef, const float* dev_omul, const float* dev_gmul,
                        int rows, int dim, float* dx, long lddx, cudaStream_t s) {
  const int blocks = (rows + 7) / 8;
  switch (ox ? odtype : 0) {
    case 0: l2norm_bwd_kernel<T, float><<<blocks, 256, 0, s>>>(dxh, ldg, x, ldx, inv_norm, (const float*)ox, ldox, oinv, ohi, ldohi, dc, usum, gscale, ucoef, dev_omul, dev_gmul, rows, dim, dx, lddx); break;
    case 1: l2norm_bwd_kernel<T, __nv_bfloat16><<<blocks, 256, 0, s>>>(dxh, ldg, x, ldx, inv_norm, (const __nv_bfloat16*)ox, ldox, oinv, ohi, ldohi, dc, usum, gscale, ucoef, dev_omul, dev_gmul, rows, dim, dx, lddx); break;
    case 2: l2norm_bwd_kernel<T, __half><<<blocks, 256, 0, s>>>(dxh, ldg, x, ldx, inv_norm, (const __half*)ox, ldox, oinv, ohi, ldohi, dc, usum, gscale, ucoef, dev_omul, dev_gmul, rows, dim, dx, lddx); break;
    default: return B2_EINVAL;
  }
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int l2norm_bwd(const float* dxh, int ldg, const void* x, int dtype, long ldx, const float* inv_norm, const void* ox,
               int odtype, long ldox, const float* oinv, const void* ohi, int ldohi, const float* dc,
               const float* usum, float gscale, float ucoef, const float* dev_omul, const float* dev_gmul, int rows,
               int dim, float* dx, long lddx, cudaStream_t s) {
  if (rows <= 0 || dim <= 0) return B2_EINVAL;
  if (dc && (!ox || !oinv || !ohi)) return B2_EINVAL;
  auto h = (const __nv_bfloat16*)ohi;
  auto d2 = (const float2*)dc;
  switch (dtype) {
    case 0: return l2norm_bwd_t<float>(dxh, ldg, (const float*)x, ldx, inv_norm, ox, odtype, ldox, oinv, h, ldohi, d2, usum, gscale, ucoef, dev_omul, dev_gmul, rows, dim, dx, lddx, s);
    case 1: return l2norm_bwd_t<__nv_bfloat16>(dxh, ldg, (const __nv_bfloat16*)x, ldx, inv_norm, ox, odtype, ldox, oinv, h, ldohi, d2, usum, gscale, ucoef, dev_omul, dev_gmul, rows, dim, dx, lddx, s);
    case 2: return l2norm_bwd_t<__half>(dxh, ldg, (const __half*)x, ldx, inv_norm, ox, odtype, ldox, oinv, h, ldohi, d2, usum, gscale, ucoef, dev_omul, dev_gmul, rows, dim, dx, lddx, s);
    default: return B2_EINVAL;
  }
}

int colsum_bf16(const void* xh, int ld, int rows, int dim, float* out, cudaStream_t s) {
  if (rows <= 0 || dim <= 0) return B2_EINVAL;
  dim3 grid((dim + 255) / 256, rows < 256 ? 1 : 64);
  colsum_bf16_kernel<<<grid, 256, 0, s>>>((const __nv_bfloat16*)xh, ld, rows, dim, out);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int gather_rows_bf16(const void* src, int lds, const long long* idx, int rows, int src_rows, int K, void* dst, int ldd,
                     cudaStream_t s) {
  if (rows <= 0 || K <= 0 || (K & 7) || (lds & 7) || (ldd & 7)) return B2_EINVAL;
  gather_rows_bf16_kernel<<<(rows + 7) / 8, 256, 0, s>>>((const uint4*)src, lds / 8, idx, rows, src_rows, K / 8,
                                                         (uint4*)dst, ldd / 8);
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
