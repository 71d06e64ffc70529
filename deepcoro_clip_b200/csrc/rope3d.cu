// K7: 3D axial RoPE apply (forward and backward) for q and k in ONE launch.
//   y = x*cos + rotate_half(x)*sin   (models/rope_3d.py:13-17, 240-246); backward is the inverse rotation
//   dx = dy*cos - rotate_half(dy)*sin. The tables are the reference's own [N, Dh] sin/cos in
//   the tensor dtype (built once per (T,H,W) on the host side with the reference's formula and cached; 600 KB at
//   N=3137, L2 resident), so the ~30 elementwise kernels + autograd saves of the reference become one streaming
//   read + write of q and k: 4 * B*heads*N*Dh*sizeof(dtype) bytes.
//   Rounding follows the reference op by op (mul, mul, add each rounded to the tensor dtype), so bf16/fp16/fp32
//   results are bit-identical to the PyTorch module.
// Layout: x[b, h, n, :] at b*sb + h*sh + n*sn (+ contiguous Dh), any strides (MViT hands permuted views).
#include "common.cuh"
#include "host_api.h"

#include "rope3d_kernels.cuh"

namespace b2host {
using namespace b2;

template <typename T>
static int rope_launch(const RopeTensor& q, const RopeTensor& k, int ntens, const void* sin_t, const void* cos_t, int B,
                       int Hh, int N, int Dh, float sgn, cudaStream_t s) {
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = Dh % VEC == 0;
  auto aligned = [&](const RopeTensor& t) {
    return (reinterpret_cast<uintptr_t>(t.in) % 16 == 0) && (t.sb * sizeof(T)) % 16 == 0 && (t.sh * sizeof(T)) % 16 == 0 &&
           (t.sn * sizeof(T)) % 16 == 0;
  };
  const long long plane = (long long)N * (Dh / VEC);
  if (vec_ok && aligned(q) && (ntens == 1 || aligned(k)) && (long long)N * Dh < (1ll << 31) && plane < (1ll << 31)) {
    const int rows = B * Hh;
    const int pblocks = (int)((plane + 255) / 256);
    // enough row groups for >= 8 CTAs per SM in total, at least ROPE_UNROLL rows per CTA
    int groups = (8 * sm_count() + pblocks * ntens - 1) / (pblocks * ntens);
    if (groups < 1) groups = 1;
    int rpb = (rows + groups - 1) / groups;
    rpb = (rpb + ROPE_UNROLL - 1) / ROPE_UNROLL * ROPE_UNROLL;
    groups = (rows + rpb - 1) / rpb;
    if (groups <= 65535) {
      dim3 grid((unsigned)pblocks, (unsigned)groups, ntens);
      rope3d_plane_kernel<T><<<grid, 256, 0, s>>>(q, k, (const T*)sin_t, (const T*)cos_t, rows, Hh, N, Dh, rpb, sgn);
      return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
    }
  }
  const long long total = (long long)B * Hh * N * (vec_ok ? Dh / VEC : Dh / 2);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks, ntens);
  if (vec_ok)
    rope3d_kernel<T, VEC><<<grid, 256, 0, s>>>(q, k, (const T*)sin_t, (const T*)cos_t, B, Hh, N, Dh, sgn);
  else
    rope3d_kernel<T, 2><<<grid, 256, 0, s>>>(q, k, (const T*)sin_t, (const T*)cos_t, B, Hh, N, Dh, sgn);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int rope3d_apply(const void* q, long long qsb, long long qsh, long long qsn, void* q_out, const void* k, long long ksb,
                 long long ksh, long long ksn, void* k_out, const void* sin_t, const void* cos_t, int dtype, int B,
                 int Hh, int N, int Dh, int backward, cudaStream_t s) {
  if (!q || !q_out || !sin_t || !cos_t || B <= 0 || Hh <= 0 || N <= 0 || Dh <= 0 || (Dh & 1)) return B2_EINVAL;
  RopeTensor tq{q, q_out, qsb, qsh, qsn};
  RopeTensor tk{k, k_out, ksb, ksh, ksn};
  const int ntens = (k && k_out) ? 2 : 1;
  const float sgn = backward ? -1.f : 1.f;
  switch (dtype) {
    case 0: return rope_launch<float>(tq, tk, ntens, sin_t, cos_t, B, Hh, N, Dh, sgn, s);
    case 1: return rope_launch<__nv_bfloat16>(tq, tk, ntens, sin_t, cos_t, B, Hh, N, Dh, sgn, s);
    case 2: return rope_launch<__half>(tq, tk, ntens, sin_t, cos_t, B, Hh, N, Dh, sgn, s);
    default: return B2_EINVAL;
  }
}

}  // namespace b2host
