// Multi-destination normalise kernel of l2norm.cu (device only: uses vector conversions the host emulation shim of
// tests/emul/ does not provide, so it lives apart from l2norm_kernels.cuh).
#pragma once

namespace b2 {

// Normalise + all-gather in ONE kernel (multi-GPU row-slab path): every bf16 operand row is stored straight into the
// [N, ld] operand buffer of EVERY rank (peer pointers of a symmetric-memory allocation: plain st.global over NVLink /
// NVSwitch for the remote ones), at row row_offset + r. The NCCL all-gather that used to follow the normalise (and its
// wait in front of the logits forward) disappears; a cross-rank barrier after the launch makes the rows visible.
// One warp per row, the row is read ONCE with 16-byte loads and kept in registers (dim <= 1024, dim % 8 == 0, 16-byte
// aligned rows), plain bf16 operands only; 16-byte stores (one NVLink packet per 8 elements).
constexpr int L2N_MAX_DEST = 8;
struct L2nDests {
  __nv_bfloat16* ptr[L2N_MAX_DEST];
  int n;
};

template <typename T>
__device__ __forceinline__ void l2n_load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void l2n_load8<float>(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void l2n_load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <>
__device__ __forceinline__ void l2n_load8<__half>(const __half* p, float (&v)[8]) {
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&a);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// kMc: dst.ptr[0] is the MULTICAST address of the symmetric operand buffer (NVSwitch replicates one multimem.st to every
// rank of the group, this rank included): each row leaves the GPU once instead of once per peer — 4 MB instead of 29 MB per
// operand and rank at 8 ranks, which was ~50 us of NVLink time per operand in front of the forward.
template <typename T, bool kMc = false>
__global__ void __launch_bounds__(256)
l2norm_fwd_multi_kernel(const T* __restrict__ x, long ldx, int rows, int dim, L2nDests dst, long row_offset, int ldo,
                        int Kp, float* __restrict__ inv_norm, int normalize) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const T* xr = x + (size_t)warp * ldx;
  float v[4][8];                                   // up to 1024 columns: 4 x (32 lanes x 8)
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < dim) {
      l2n_load8<T>(xr + c, v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss = fmaf(v[i][e], v[i][e], ss);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
    }
  }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  const float inv = normalize ? 1.f / fmaxf(nrm, 1e-12f) : 1.f;
  if (lane == 0 && inv_norm) inv_norm[warp] = normalize ? inv : nrm;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 8;
    if (c < Kp) {
      uint4 pk;
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[i][2 * e] * inv, v[i][2 * e + 1] * inv);
      const size_t off = (size_t)(row_offset + warp) * ldo + c;
      if (kMc) {
        asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst.ptr[0] + off), "f"(__uint_as_float(pk.x)),
                     "f"(__uint_as_float(pk.y)), "f"(__uint_as_float(pk.z)), "f"(__uint_as_float(pk.w)) : "memory");
      } else {
        for (int d = 0; d < dst.n; ++d) *reinterpret_cast<uint4*>(dst.ptr[d] + off) = pk;
      }
    }
  }
}

}  // namespace b2
