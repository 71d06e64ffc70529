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

#include "l2norm_kernels.cuh"
#include "l2norm_multi_kernels.cuh"

namespace b2 {

// l2norm_bwd_kernel for the shapes of the contrastive step (fp32 everywhere, dim = NCH * 128, 16-byte aligned rows): a lane
// keeps its NCH float4 of the gradient and of x in registers, so every array is read exactly once with 16-byte loads (the
// general kernel walks each row twice with 4-byte loads: 61 us for 301 MB at 32k x 512).
template <int NCH>
__global__ void __launch_bounds__(256)
l2norm_bwd_vec_kernel(const float* __restrict__ dxh, int ldg, const float* __restrict__ x, long ldx,
                      const float* __restrict__ inv_norm, const float* __restrict__ ox, long ldox,
                      const float* __restrict__ oinv, const __nv_bfloat16* __restrict__ ohi, int ldohi,
                      const float2* __restrict__ dc, const float* __restrict__ usum, float gscale, float ucoef,
                      const float* __restrict__ dev_omul, const float* __restrict__ dev_gmul, int rows,
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
  const float inv = inv_norm[warp];
  float gv[NCH][4], xv[NCH][4];
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = k * 128 + lane * 4;
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(dxh + (size_t)warp * ldg + c));
    const float4 x4 = __ldg(reinterpret_cast<const float4*>(x + (size_t)warp * ldx + c));
    gv[k][0] = gscale * g4.x; gv[k][1] = gscale * g4.y; gv[k][2] = gscale * g4.z; gv[k][3] = gscale * g4.w;
    xv[k][0] = x4.x * inv; xv[k][1] = x4.y * inv; xv[k][2] = x4.z * inv; xv[k][3] = x4.w * inv;
    if (has_dc) {
      const float4 o4 = __ldg(reinterpret_cast<const float4*>(ox + (size_t)warp * ldox + c));
      const uint2 hraw = __ldg(reinterpret_cast<const uint2*>(ohi + (size_t)warp * ldohi + c));
      const float yh[4] = {o4.x * oi, o4.y * oi, o4.z * oi, o4.w * oi};
      const float yhi[4] = {__uint_as_float(hraw.x << 16), __uint_as_float(hraw.x & 0xffff0000u), __uint_as_float(hraw.y << 16),
                            __uint_as_float(hraw.y & 0xffff0000u)};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        gv[k][e] = fmaf(res, yh[e], gv[k][e]);
        gv[k][e] = fmaf(gb, yh[e] - yhi[e], gv[k][e]);
      }
    }
    if (usum) {
      const float4 u4 = __ldg(reinterpret_cast<const float4*>(usum + c));
      gv[k][0] = fmaf(uc, u4.x, gv[k][0]); gv[k][1] = fmaf(uc, u4.y, gv[k][1]);
      gv[k][2] = fmaf(uc, u4.z, gv[k][2]); gv[k][3] = fmaf(uc, u4.w, gv[k][3]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) dot = fmaf(gv[k][e], xv[k][e], dot);
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const int c = k * 128 + lane * 4;
    *reinterpret_cast<float4*>(dx + (size_t)warp * lddx + c) =
        make_float4(gmul * (gv[k][0] - dot * xv[k][0]) * inv, gmul * (gv[k][1] - dot * xv[k][1]) * inv,
                    gmul * (gv[k][2] - dot * xv[k][2]) * inv, gmul * (gv[k][3] - dot * xv[k][3]) * inv);
  }
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

int l2norm_fwd_multi(const void* x, int dtype, long ldx, int rows, int dim, void* const* outs_host, int n_out,
                     long long row_offset, int ldo, int Kp, float* inv_norm, int normalize, cudaStream_t s) {
  if (rows <= 0 || dim <= 0 || Kp < dim || Kp % 64 || Kp > 1024 || dim % 8 || n_out < 1 || n_out > L2N_MAX_DEST ||
      !outs_host || (ldo & 7) || row_offset < 0)
    return B2_EINVAL;
  const int esz = dtype == 0 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || ((ldx * esz) & 15)) return B2_EINVAL;
  L2nDests d;
  d.n = n_out;
  for (int i = 0; i < L2N_MAX_DEST; ++i) {
    d.ptr[i] = i < n_out ? reinterpret_cast<__nv_bfloat16*>(outs_host[i]) : nullptr;
    if (i < n_out && (!outs_host[i] || (reinterpret_cast<uintptr_t>(outs_host[i]) & 15))) return B2_EINVAL;
  }
  const int blocks = (rows + 7) / 8;
  switch (dtype) {
    case 0: l2norm_fwd_multi_kernel<float><<<blocks, 256, 0, s>>>((const float*)x, ldx, rows, dim, d, (long)row_offset, ldo, Kp, inv_norm, normalize); break;
    case 1: l2norm_fwd_multi_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)x, ldx, rows, dim, d, (long)row_offset, ldo, Kp, inv_norm, normalize); break;
    case 2: l2norm_fwd_multi_kernel<__half><<<blocks, 256, 0, s>>>((const __half*)x, ldx, rows, dim, d, (long)row_offset, ldo, Kp, inv_norm, normalize); break;
    default: return B2_EINVAL;
  }
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int l2norm_fwd_mc(const void* x, int dtype, long ldx, int rows, int dim, void* mc_out, long long row_offset, int ldo, int Kp,
                  float* inv_norm, int normalize, cudaStream_t s) {
  if (rows <= 0 || dim <= 0 || Kp < dim || Kp % 64 || Kp > 1024 || dim % 8 || !mc_out || (ldo & 7) || row_offset < 0 ||
      (reinterpret_cast<uintptr_t>(mc_out) & 15))
    return B2_EINVAL;
  const int esz = dtype == 0 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || ((ldx * esz) & 15)) return B2_EINVAL;
  L2nDests d{};
  d.n = 1;
  d.ptr[0] = reinterpret_cast<__nv_bfloat16*>(mc_out);
  const int blocks = (rows + 7) / 8;
  switch (dtype) {
    case 0: l2norm_fwd_multi_kernel<float, true><<<blocks, 256, 0, s>>>((const float*)x, ldx, rows, dim, d, (long)row_offset, ldo, Kp, inv_norm, normalize); break;
    case 1: l2norm_fwd_multi_kernel<__nv_bfloat16, true><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)x, ldx, rows, dim, d, (long)row_offset, ldo, Kp, inv_norm, normalize); break;
    case 2: l2norm_fwd_multi_kernel<__half, true><<<blocks, 256, 0, s>>>((const __half*)x, ldx, rows, dim, d, (long)row_offset, ldo, Kp, inv_norm, normalize); break;
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
  // fp32 rows of 128 k floats with 16-byte aligned pitches: the single-pass vector kernel
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (dtype == 0 && (!ox || odtype == 0) && dim % 128 == 0 && dim <= 1024 && ldg % 4 == 0 && ldx % 4 == 0 && lddx % 4 == 0 &&
      al16(dxh) && al16(x) && al16(dx) && (!ox || (ldox % 4 == 0 && al16(ox))) &&
      (!ohi || (ldohi % 4 == 0 && (reinterpret_cast<uintptr_t>(ohi) & 7) == 0)) && (!usum || al16(usum))) {
    const int blocks = (rows + 7) / 8;
#define L2B_VEC(N)                                                                                                        \
  case N:                                                                                                                 \
    l2norm_bwd_vec_kernel<N><<<blocks, 256, 0, s>>>(dxh, ldg, (const float*)x, ldx, inv_norm, (const float*)ox, ldox, oinv, h, \
                                                    ldohi, d2, usum, gscale, ucoef, dev_omul, dev_gmul, rows, dx, lddx);  \
    break;
    switch (dim / 128) {
      L2B_VEC(1) L2B_VEC(2) L2B_VEC(3) L2B_VEC(4) L2B_VEC(5) L2B_VEC(6) L2B_VEC(7) L2B_VEC(8)
    }
#undef L2B_VEC
    return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
  }
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

template <typename TA>
static int rowdot_raw_t(const TA* a, long long lda, const float* ainv, const void* b, int bdtype, long long ldb,
                        const float* binv, int rows, int dim, float* out, cudaStream_t s) {
  const int blocks = (rows + 7) / 8;
  switch (bdtype) {
    case 0: rowdot_raw_kernel<TA, float><<<blocks, 256, 0, s>>>(a, lda, ainv, (const float*)b, ldb, binv, rows, dim, out); break;
    case 1: rowdot_raw_kernel<TA, __nv_bfloat16><<<blocks, 256, 0, s>>>(a, lda, ainv, (const __nv_bfloat16*)b, ldb, binv, rows, dim, out); break;
    case 2: rowdot_raw_kernel<TA, __half><<<blocks, 256, 0, s>>>(a, lda, ainv, (const __half*)b, ldb, binv, rows, dim, out); break;
    default: return B2_EINVAL;
  }
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int rowdot_raw(const void* a, int adtype, long long lda, const float* ainv, const void* b, int bdtype, long long ldb,
               const float* binv, int rows, int dim, float* out, cudaStream_t s) {
  if (rows <= 0 || dim <= 0 || !a || !b || !ainv || !binv || !out) return B2_EINVAL;
  switch (adtype) {
    case 0: return rowdot_raw_t<float>((const float*)a, lda, ainv, b, bdtype, ldb, binv, rows, dim, out, s);
    case 1: return rowdot_raw_t<__nv_bfloat16>((const __nv_bfloat16*)a, lda, ainv, b, bdtype, ldb, binv, rows, dim, out, s);
    case 2: return rowdot_raw_t<__half>((const __half*)a, lda, ainv, b, bdtype, ldb, binv, rows, dim, out, s);
    default: return B2_EINVAL;
  }
}

int rowdot_bf16(const void* a, int lda, const void* b, int ldb, const long long* idx, int rows, int b_rows, int K,
                float* out, cudaStream_t s) {
  if (rows <= 0 || K <= 0 || (K & 1)) return B2_EINVAL;
  rowdot_bf16_kernel<<<(rows + 7) / 8, 256, 0, s>>>((const __nv_bfloat16*)a, lda, (const __nv_bfloat16*)b, ldb, idx,
                                                    rows, b_rows, K, out);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
