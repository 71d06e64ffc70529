// Kernels of attnpool.cu (CUDA-core attention pool, merge, backward; see there for the design). No inline PTX and no
// include: CUDA types / intrinsics come from the including translation unit (attnpool.cu) or from the host emulation
// (tests/emul/), which must also define B2_DYN_SMEM16(name) (the dynamic shared memory window).
#pragma once

namespace b2 {

constexpr int AP_TOK = 32;   // tokens per smem tile


template <typename T> __device__ __forceinline__ float ap_ld(const T* p);
template <> __device__ __forceinline__ float ap_ld<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ap_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float ap_ld<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ T ap_cv(float v);
template <> __device__ __forceinline__ float ap_cv<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 ap_cv<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half ap_cv<__half>(float v) { return __float2half_rn(v); }

// lane l owns, in every 16-byte-chunk row c, the VE elements d = (c*32 + l)*VE + i  (conflict-free LDS.128)
template <typename T, int NCH>
struct ApLane {
  static constexpr int VE = 16 / sizeof(T);
  static constexpr int EPL = NCH * VE;
  static __device__ __forceinline__ int d(int lane, int c, int i) { return (c * 32 + lane) * VE + i; }
};

struct PoolFwdParams {
  const void* x; long long sb, sn;     // [B, N, D], D contiguous
  const unsigned char* mask; long long mb;   // [B, N] (may be null)
  const float* qt;                     // [H, D]  (online-softmax mode)
  const float* w; long long wb, wh;    // [B, H, N] given weights (weighted-sum mode), else null
  float* part_m; float* part_l; float* part_acc;   // [B, S, H], [B, S, H], [B, S, H, D]
  int B, N, D, H, S;
  float drop_p; unsigned long long drop_seed; float* part_l2;   // attention dropout (training): sum of kept weights
};

template <typename T, int NCH>
__global__ void __launch_bounds__(512) pool_fwd_kernel(PoolFwdParams p) {
  using L = ApLane<T, NCH>;
  B2_DYN_SMEM16(smem);
  T* tile[2] = {reinterpret_cast<T*>(smem), reinterpret_cast<T*>(smem) + (size_t)AP_TOK * p.D};
  const int b = blockIdx.x, sp = blockIdx.y;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (int)((long long)p.N * sp / p.S), n1 = (int)((long long)p.N * (sp + 1) / p.S);
  const T* xb = reinterpret_cast<const T*>(p.x) + b * p.sb;
  const int chunks_per_row = p.D * (int)sizeof(T) / 16;

  auto load_tile = [&](int buf, int t0) {
    const int rows = min(AP_TOK, n1 - t0);
    for (int i = threadIdx.x; i < rows * chunks_per_row; i += blockDim.x) {
      const int r = i / chunks_per_row, c = i - r * chunks_per_row;
      __pipeline_memcpy_async(reinterpret_cast<unsigned char*>(tile[buf]) + ((size_t)r * p.D * sizeof(T) + c * 16),
                              reinterpret_cast<const unsigned char*>(xb + (t0 + r) * p.sn) + c * 16, 16);
    }
    __pipeline_commit();
  };

  float q[L::EPL], acc[L::EPL];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int i = 0; i < L::VE; ++i) {
      q[c * L::VE + i] = p.w ? 0.f : p.qt[(size_t)h * p.D + L::d(lane, c, i)];
      acc[c * L::VE + i] = 0.f;
    }
  float m = -INFINITY, l = 0.f, l2 = 0.f;
  const float keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  const unsigned char* mk = p.mask ? p.mask + b * p.mb : nullptr;
  const float* wrow = p.w ? p.w + b * p.wb + h * p.wh : nullptr;

  int buf = 0;
  if (n0 < n1) load_tile(0, n0);
  for (int t0 = n0; t0 < n1; t0 += AP_TOK) {
    if (t0 + AP_TOK < n1) {
      load_tile(buf ^ 1, t0 + AP_TOK);
      __pipeline_wait_prior(1);
    } else {
      __pipeline_wait_prior(0);
    }
    __syncthreads();
    const int rows = min(AP_TOK, n1 - t0);
    const T* tl = tile[buf];
    for (int r = 0; r < rows; ++r) {
      float xv[L::EPL];
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(tl + (size_t)r * p.D + (c * 32 + lane) * L::VE);
        const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
        for (int i = 0; i < L::VE; ++i) xv[c * L::VE + i] = ap_ld<T>(e + i);
      }
      float wgt;
      if (wrow) {
        wgt = wrow[t0 + r];
      } else {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < L::EPL; ++i) s = fmaf(xv[i], q[i], s);
        s = warp_sum(s);
        if (mk && mk[t0 + r]) s = -INFINITY;
        const float mn = fmaxf(m, s);
        if (mn == -INFINITY) continue;                   // every token so far masked
        const float sc = __expf(m - mn);                 // exp(-inf) = 0 on the first real token
        wgt = __expf(s - mn);
        l = l * sc + wgt;                                // softmax denominator: BEFORE dropout (nn.MultiheadAttention)
        if (p.drop_p > 0.f) wgt = attn_keep(p.drop_seed, b * p.H + h, t0 + r, p.drop_p) ? wgt * keep_scale : 0.f;
        l2 = l2 * sc + wgt;
        m = mn;
#pragma unroll
        for (int i = 0; i < L::EPL; ++i) acc[i] *= sc;
      }
#pragma unroll
      for (int i = 0; i < L::EPL; ++i) acc[i] = fmaf(wgt, xv[i], acc[i]);
    }
    __syncthreads();
    buf ^= 1;
  }
  const size_t slot = ((size_t)b * p.S + sp) * p.H + h;
  if (lane == 0 && p.part_m) {
    p.part_m[slot] = m;
    p.part_l[slot] = l;
    if (p.part_l2) p.part_l2[slot] = l2;
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int i = 0; i < L::VE; ++i) p.part_acc[slot * p.D + L::d(lane, c, i)] = acc[c * L::VE + i];
}

// merge the S splits of one (b, h): online-softmax mode -> xbar = sum_s e^{m_s - m} acc_s / sum_s e^{m_s - m} l_s;
// weighted-sum mode (part_m == null) -> plain sum, optionally accumulated atomically over b into out (dqt).
__global__ void __launch_bounds__(256)
pool_merge_kernel(const float* __restrict__ part_m, const float* __restrict__ part_l, const float* __restrict__ part_acc,
                  int B, int S, int H, int D, float* __restrict__ out, float* __restrict__ out_m,
                  float* __restrict__ out_l, int sum_over_b, const float* __restrict__ part_l2,
                  float* __restrict__ out_sa) {
  const int bh = blockIdx.x;               // b * H + h
  const int b = bh / H, h = bh - b * H;
  float m = -INFINITY;
  if (part_m)
    for (int s = 0; s < S; ++s) m = fmaxf(m, part_m[((size_t)b * S + s) * H + h]);
  float l = 0.f, l2 = 0.f;
  if (part_m)
    for (int s = 0; s < S; ++s) {
      const float ms = part_m[((size_t)b * S + s) * H + h];
      l += ms == -INFINITY ? 0.f : part_l[((size_t)b * S + s) * H + h] * __expf(ms - m);
      if (part_l2) l2 += ms == -INFINITY ? 0.f : part_l2[((size_t)b * S + s) * H + h] * __expf(ms - m);
    }
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) {
      const size_t slot = ((size_t)b * S + s) * H + h;
      float sc = 1.f;
      if (part_m) {
        const float ms = part_m[slot];
        sc = ms == -INFINITY ? 0.f : __expf(ms - m);
      }
      a = fmaf(sc, part_acc[slot * D + d], a);
    }
    if (part_m) out[(size_t)bh * D + d] = a / l;           // all-masked row -> 0/0 = NaN like nn.MultiheadAttention
    else if (sum_over_b) atomicAdd(out + (size_t)h * D + d, a);
    else out[(size_t)bh * D + d] = a;
  }
  if (threadIdx.x == 0 && part_m) {
    out_m[bh] = m;
    out_l[bh] = l;
    if (out_sa) out_sa[bh] = part_l2 ? l2 / l : 1.f;       // sum of the (dropped, rescaled) attention weights
  }
}

struct PoolBwdParams {
  const void* x; long long sb, sn;
  const unsigned char* mask; long long mb;
  const float* qt;        // [H, D]
  const float* dxbar;     // [B, H, D]
  const float* xbar;      // [B, H, D]
  const float* m; const float* l;   // [B, H]
  void* dx;               // [B, N, D] contiguous, dtype T
  float* ds;              // [B, H, N]
  int B, N, D, H;
  const float* sa; const float* dsa;   // [B, H] sum of dropped weights and its upstream gradient (dropout only)
  float drop_p; unsigned long long drop_seed;
  const float* dlse;                   // [B, H] upstream gradient of lse_h = m_h + log l_h (may be null)
};

// warp = token. smem: qt [H][D], dxbar_b [H][D] (fp32), c_h = dxbar_h . xbar_h
template <typename T, int NCH>
__global__ void __launch_bounds__(256) pool_bwd_dx_kernel(PoolBwdParams p) {
  using L = ApLane<T, NCH>;
  B2_DYN_SMEM16(smem);
  float* sq = reinterpret_cast<float*>(smem);
  float* sd = sq + (size_t)p.H * p.D;
  float* sc = sd + (size_t)p.H * p.D;          // [H] dxbar.xbar
  float* sm = sc + p.H;                        // [H] m
  float* sl = sm + p.H;                        // [H] 1/l
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < p.H * p.D; i += blockDim.x) {
    sq[i] = p.qt[i];
    sd[i] = p.dxbar[(size_t)b * p.H * p.D + i];
  }
  __syncthreads();
  for (int h = warp; h < p.H; h += nw) {
    float c = 0.f;
    for (int d = lane; d < p.D; d += 32) c = fmaf(sd[h * p.D + d], p.xbar[((size_t)b * p.H + h) * p.D + d], c);
    c = warp_sum(c);
    if (p.dsa) c = fmaf(p.dsa[b * p.H + h], p.sa[b * p.H + h], c);      // c = sum_n a_n kappa_n (T_n + dsa)
    if (p.dlse) c -= p.dlse[b * p.H + h];                                // d lse / d s_n = a_n
    if (lane == 0) {
      sc[h] = c;
      sm[h] = p.m[b * p.H + h];
      sl[h] = 1.f / p.l[b * p.H + h];
    }
  }
  __syncthreads();
  const T* xb = reinterpret_cast<const T*>(p.x) + b * p.sb;
  const unsigned char* mk = p.mask ? p.mask + b * p.mb : nullptr;
  T* dxb = reinterpret_cast<T*>(p.dx) + (size_t)b * p.N * p.D;
  for (int n = blockIdx.y * nw + warp; n < p.N; n += gridDim.y * nw) {
    float xv[L::EPL], dxv[L::EPL];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const uint4 v = *reinterpret_cast<const uint4*>(xb + n * p.sn + (c * 32 + lane) * L::VE);
      const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
      for (int i = 0; i < L::VE; ++i) {
        xv[c * L::VE + i] = ap_ld<T>(e + i);
        dxv[c * L::VE + i] = 0.f;
      }
    }
    const bool masked = mk && mk[n];
    for (int h = 0; h < p.H; ++h) {
      float s = 0.f, da = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const float* qh = sq + h * p.D + (c * 32 + lane) * L::VE;
        const float* dh = sd + h * p.D + (c * 32 + lane) * L::VE;
#pragma unroll
        for (int i = 0; i < L::VE; i += 4) {
          const float4 qv = *reinterpret_cast<const float4*>(qh + i);
          const float4 dv = *reinterpret_cast<const float4*>(dh + i);
          const int o = c * L::VE + i;
          s = fmaf(xv[o], qv.x, s); s = fmaf(xv[o + 1], qv.y, s); s = fmaf(xv[o + 2], qv.z, s); s = fmaf(xv[o + 3], qv.w, s);
          da = fmaf(xv[o], dv.x, da); da = fmaf(xv[o + 1], dv.y, da); da = fmaf(xv[o + 2], dv.z, da); da = fmaf(xv[o + 3], dv.w, da);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        da += __shfl_xor_sync(0xffffffffu, da, o);
      }
      const float a0 = masked ? 0.f : __expf(s - sm[h]) * sl[h];
      float kap = 1.f;
      if (p.drop_p > 0.f) kap = attn_keep(p.drop_seed, b * p.H + h, n, p.drop_p) ? 1.f / (1.f - p.drop_p) : 0.f;
      const float dsv = a0 * (kap * (da + (p.dsa ? p.dsa[b * p.H + h] : 0.f)) - sc[h]);
      const float a = a0 * kap;                            // weight that multiplied x_n in the forward
      if (lane == 0) p.ds[((size_t)b * p.H + h) * p.N + n] = dsv;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const float* qh = sq + h * p.D + (c * 32 + lane) * L::VE;
        const float* dh = sd + h * p.D + (c * 32 + lane) * L::VE;
#pragma unroll
        for (int i = 0; i < L::VE; i += 4) {
          const float4 qv = *reinterpret_cast<const float4*>(qh + i);
          const float4 dv = *reinterpret_cast<const float4*>(dh + i);
          const int o = c * L::VE + i;
          dxv[o] = fmaf(a, dv.x, fmaf(dsv, qv.x, dxv[o]));
          dxv[o + 1] = fmaf(a, dv.y, fmaf(dsv, qv.y, dxv[o + 1]));
          dxv[o + 2] = fmaf(a, dv.z, fmaf(dsv, qv.z, dxv[o + 2]));
          dxv[o + 3] = fmaf(a, dv.w, fmaf(dsv, qv.w, dxv[o + 3]));
        }
      }
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      T ov[L::VE];
#pragma unroll
      for (int i = 0; i < L::VE; ++i) ov[i] = ap_cv<T>(dxv[c * L::VE + i]);
      *reinterpret_cast<uint4*>(dxb + (size_t)n * p.D + (c * 32 + lane) * L::VE) = *reinterpret_cast<const uint4*>(ov);
    }
  }
}

}  // namespace b2
