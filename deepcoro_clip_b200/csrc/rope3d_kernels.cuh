// Kernels of rope3d.cu (see there). No inline PTX and no include: CUDA types / intrinsics come from the including
// translation unit or from the host emulation (tests/emul/).
#pragma once

namespace b2 {

template <typename T> struct RopeT;
template <> struct RopeT<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ float rnd(float v) { return v; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct RopeT<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ float rnd(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};
template <> struct RopeT<__half> {
  static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
  static __device__ __forceinline__ float rnd(float v) { return __half2float(__float2half_rn(v)); }
  static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
};

struct RopeTensor {
  const void* in;
  void* out;
  long long sb, sh, sn;      // input strides (elements)
};

// one thread = VEC consecutive channels (VEC even) of one (b, h, n) row; blockIdx.y selects q / k
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
rope3d_kernel(RopeTensor tq, RopeTensor tk, const T* __restrict__ sin_t, const T* __restrict__ cos_t, int B, int Hh,
              int N, int Dh, float sgn) {
  const RopeTensor t = blockIdx.y == 0 ? tq : tk;
  const int vpr = Dh / VEC;
  const long long total = (long long)B * Hh * N * vpr;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % vpr);
    long long row = idx / vpr;
    const int n = (int)(row % N);
    row /= N;
    const int h = (int)(row % Hh);
    const int b = (int)(row / Hh);
    const T* src = reinterpret_cast<const T*>(t.in) + b * t.sb + h * t.sh + n * t.sn + v * VEC;
    T* dst = reinterpret_cast<T*>(t.out) + ((((long long)b * Hh + h) * N + n) * Dh) + v * VEC;
    const T* sp = sin_t + (long long)n * Dh + v * VEC;
    const T* cp = cos_t + (long long)n * Dh + v * VEC;
    float x[VEC], s[VEC], c[VEC];
    if (VEC * sizeof(T) == 16 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
      // 16-byte vector path
      const uint4 xv = *reinterpret_cast<const uint4*>(src);
      const uint4 sv = *reinterpret_cast<const uint4*>(sp);
      const uint4 cv = *reinterpret_cast<const uint4*>(cp);
      const T* xe = reinterpret_cast<const T*>(&xv);
      const T* se = reinterpret_cast<const T*>(&sv);
      const T* ce = reinterpret_cast<const T*>(&cv);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        x[i] = RopeT<T>::ld(xe + i);
        s[i] = RopeT<T>::ld(se + i);
        c[i] = RopeT<T>::ld(ce + i);
      }
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        x[i] = RopeT<T>::ld(src + i);
        s[i] = RopeT<T>::ld(sp + i);
        c[i] = RopeT<T>::ld(cp + i);
      }
    }
    T outv[VEC];
#pragma unroll
    for (int i = 0; i < VEC; i += 2) {
      // rotate_half: r[2i] = -x[2i+1], r[2i+1] = x[2i]; every product and the sum rounded like the reference ops
      const float a0 = RopeT<T>::rnd(__fmul_rn(x[i], c[i]));
      const float a1 = RopeT<T>::rnd(__fmul_rn(x[i + 1], c[i + 1]));
      float b0, b1;
      if (sgn > 0.f) {   // forward: y = x*cos + r*sin
        b0 = RopeT<T>::rnd(__fmul_rn(-x[i + 1], s[i]));
        b1 = RopeT<T>::rnd(__fmul_rn(x[i], s[i + 1]));
      } else {           // backward: dx[2i] = dy[2i]c[2i] + dy[2i+1]s[2i+1] ; dx[2i+1] = dy[2i+1]c[2i+1] - dy[2i]s[2i]
        b0 = RopeT<T>::rnd(__fmul_rn(x[i + 1], s[i + 1]));
        b1 = RopeT<T>::rnd(__fmul_rn(-x[i], s[i]));
      }
      RopeT<T>::st(&outv[i], __fadd_rn(a0, b0));
      RopeT<T>::st(&outv[i + 1], __fadd_rn(a1, b1));
    }
    if (VEC * sizeof(T) == 16) {
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(outv);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) dst[i] = outv[i];
    }
  }
}

// Fast path (16-byte vectors): one thread = one 16-byte vector position (n, v) of the [N, Dh] plane, kept for RPB
// consecutive (b, h) rows, so the sin / cos vectors are loaded once per thread and ROPE_UNROLL independent 16-byte row
// loads are in flight per thread; a warp reads 512 contiguous bytes of one row per load. grid = (plane chunks,
// row groups, q|k). All index math is 32-bit.
constexpr int ROPE_UNROLL = 4;
template <typename T>
__global__ void __launch_bounds__(256)
rope3d_plane_kernel(RopeTensor tq, RopeTensor tk, const T* __restrict__ sin_t, const T* __restrict__ cos_t, int rows,
                    int Hh, int N, int Dh, int rpb, float sgn) {
  constexpr int VEC = 16 / sizeof(T);
  const RopeTensor t = blockIdx.z == 0 ? tq : tk;
  const int vpr = Dh / VEC;
  const int pv = blockIdx.x * blockDim.x + threadIdx.x;        // vector index inside the [N, Dh] plane
  if (pv >= N * vpr) return;
  const int n = pv / vpr, v = pv - n * vpr;
  float s[VEC], c[VEC];
  {
    const uint4 sv = *reinterpret_cast<const uint4*>(sin_t + (size_t)pv * VEC);
    const uint4 cv = *reinterpret_cast<const uint4*>(cos_t + (size_t)pv * VEC);
    const T* se = reinterpret_cast<const T*>(&sv);
    const T* ce = reinterpret_cast<const T*>(&cv);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      s[i] = RopeT<T>::ld(se + i);
      c[i] = RopeT<T>::ld(ce + i);
    }
  }
  const int r0 = blockIdx.y * rpb;
  const int r1 = min(rows, r0 + rpb);
  const long long in_off = (long long)n * t.sn + v * VEC;
  const size_t out_off = (size_t)n * Dh + v * VEC;
  for (int r = r0; r < r1; r += ROPE_UNROLL) {
    uint4 xv[ROPE_UNROLL];
#pragma unroll
    for (int u = 0; u < ROPE_UNROLL; ++u) {
      const int rr = min(r + u, r1 - 1);
      const int b = rr / Hh, h = rr - b * Hh;
      xv[u] = *reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(t.in) + b * t.sb + h * t.sh + in_off);
    }
#pragma unroll
    for (int u = 0; u < ROPE_UNROLL; ++u) {
      if (r + u < r1) {
        const T* xe = reinterpret_cast<const T*>(&xv[u]);
        T outv[VEC];
#pragma unroll
        for (int i = 0; i < VEC; i += 2) {
          const float x0 = RopeT<T>::ld(xe + i), x1 = RopeT<T>::ld(xe + i + 1);
          const float a0 = RopeT<T>::rnd(__fmul_rn(x0, c[i]));
          const float a1 = RopeT<T>::rnd(__fmul_rn(x1, c[i + 1]));
          float b0, b1;
          if (sgn > 0.f) {
            b0 = RopeT<T>::rnd(__fmul_rn(-x1, s[i]));
            b1 = RopeT<T>::rnd(__fmul_rn(x0, s[i + 1]));
          } else {
            b0 = RopeT<T>::rnd(__fmul_rn(x1, s[i + 1]));
            b1 = RopeT<T>::rnd(__fmul_rn(-x0, s[i]));
          }
          RopeT<T>::st(&outv[i], __fadd_rn(a0, b0));
          RopeT<T>::st(&outv[i + 1], __fadd_rn(a1, b1));
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<T*>(t.out) + (size_t)(r + u) * N * Dh + out_off) =
            *reinterpret_cast<const uint4*>(outv);
      }
    }
  }
}

}  // namespace b2
