// Kernels of l2norm.cu (normalise forward / backward, operand helpers; see there).
// No inline PTX and no include: CUDA types / intrinsics and warp_sum come from the including translation unit
// (common.cuh) or from the host emulation (tests/emul/).
#pragma once

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

// out[r] = (a[r, :dim] . b[r, :dim]) * ainv[r] * binv[r]: the cosine of a pair from the RAW features in fp32 (the target
// logit of the softmax losses; the tensor-core value of the same pair carries the 2^-9 rounding of the bf16 operands, which
// — unlike the errors inside the log-sum-exps — is not averaged over a row: it was the whole 1e-5 loss error at N <= 16k).
template <typename TA, typename TB>
__global__ void __launch_bounds__(256)
rowdot_raw_kernel(const TA* __restrict__ a, long long lda, const float* __restrict__ ainv, const TB* __restrict__ b,
                  long long ldb, const float* __restrict__ binv, int rows, int dim, float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const TA* ar = a + (size_t)warp * lda;
  const TB* br = b + (size_t)warp * ldb;
  float s0 = 0.f, s1 = 0.f;
  int c = lane;
  for (; c + 32 < dim; c += 64) {
    s0 = fmaf(to_f32(ar[c]), to_f32(br[c]), s0);
    s1 = fmaf(to_f32(ar[c + 32]), to_f32(br[c + 32]), s1);
  }
  if (c < dim) s0 = fmaf(to_f32(ar[c]), to_f32(br[c]), s0);
  const float s = warp_sum(s0 + s1);
  if (lane == 0) out[warp] = s * ainv[warp] * binv[warp];
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
