// The [B, D]-vector work around the streaming attention-pool kernels (reference models/attention_pool.py:77-99: the
// in-projection of the learnable query, the value / output projections, LayerNorm and the optional output Linear of
// nn.MultiheadAttention + AttentionPool), forward and backward, in a handful of launches instead of ~45 framework
// launches on [B, D] operands (the module was host-bound, 1.2 ms eager for 0.24 ms of pool kernels).
//
//   pool_prep        params only: q0 = W_q query + b_q, qt_h = W_k,h^T q0_h / sqrt(Dh), and the 16-bit hi / lo operand image
//                    of qt that the tcgen05 forward kernel bulk-copies into shared memory.
//   pool_tail_fwd    clusters of 8 CTAs x 4 batch rows: merge of the token-split partials -> xbar, o = W_v,h xbar_h + b_v sa,
//                    y = W_o o + b_o, LayerNorm (two-pass statistics), optional proj. CTA r owns output columns
//                    [r D/8, (r+1) D/8); full rows are exchanged through distributed shared memory, every weight slice is
//                    read once per cluster.
//   pool_tail_bwd    the same clusters backwards: dproj, LayerNorm backward, do = W_o^T dy, dxbar_h = W_v,h^T do_h (+ dsa),
//                    plus the per-row operand image [qt_hi ; dxbar_hi ; qt_lo ; dxbar_lo] and c_h = dxbar_h . xbar_h that the
//                    tcgen05 backward kernel consumes.
//   pool_param_grads dW_o, db_o, dW_v, db_v, dgamma, dbeta (dW_p, db_p): rank-B updates, one CTA per 8 weight rows.
//   pool_qgrads      one cluster: dqt = sum of the per-(b, split) partials, dW_k, dq0, then dW_q, db_q, dquery.
// All fp32 on CUDA cores: O(B D^2) FLOP next to a 100 MB streaming pass.
#include "common.cuh"
#include "host_api.h"

namespace b2 {

constexpr int TL_THREADS = 256;
constexpr int TL_RB = 4;        // batch rows per cluster
constexpr int TL_CL = 8;        // CTAs per cluster = column slices
constexpr int TL_MAXW = 64;     // widest slice (D <= 512)

__device__ __forceinline__ float ld_dsmem(uint32_t local_addr, uint32_t rank) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(mapa_cluster(local_addr, rank)) : "memory");
  return v;
}
__device__ __forceinline__ uint16_t tl_bits(float v, int fp16) {
  return fp16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float tl_val(uint16_t b, int fp16) {
  return fp16 ? __half2float(__ushort_as_half(b)) : __bfloat162float(__ushort_as_bfloat16(b));
}
__device__ __forceinline__ uint32_t tl_sw128(int row, int col) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 3) ^ row) & 7) << 4) + (col & 7) * 2);
}
__device__ __forceinline__ float ld_any(const void* p, size_t i, int dtype) {
  if (dtype == 0) return reinterpret_cast<const float*>(p)[i];
  if (dtype == 1) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, size_t i, int dtype, float v) {
  if (dtype == 0) reinterpret_cast<float*>(p)[i] = v;
  else if (dtype == 1) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
}

// out[r * ostride + j] = bias[j0 + j] * bscale[r] + sum_i W[(j0 + j) * ldw + i] v[r * vstride + i]   (j < nj, i < K, K % 128 == 0)
// v in shared memory. A warp takes four output columns at a time (four independent weight-row requests in flight per step —
// these kernels are latency-bound on the weight reads — sharing the reads of v); the weight slice is read exactly once.
__device__ void slice_matvec(const float* __restrict__ W, long long ldw, int j0, int nj, int K, const float* v, int vstride,
                             float* out, int ostride, const float* __restrict__ bias, const float* bscale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = TL_THREADS / 32;
  for (int jb = warp * 4; jb < nj; jb += nw * 4) {
    float acc[4][TL_RB];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int r = 0; r < TL_RB; ++r) acc[q][r] = 0.f;
    const float* wr[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) wr[q] = W + (long long)(j0 + min(jb + q, nj - 1)) * ldw;
#pragma unroll 2
    for (int i = lane * 4; i < K; i += 128) {
      float4 w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = *reinterpret_cast<const float4*>(wr[q] + i);
#pragma unroll
      for (int r = 0; r < TL_RB; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(v + r * vstride + i);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          acc[q][r] = fmaf(w[q].x, x.x, fmaf(w[q].y, x.y, fmaf(w[q].z, x.z, fmaf(w[q].w, x.w, acc[q][r]))));
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int r = 0; r < TL_RB; ++r) {
        const float s = warp_sum(acc[q][r]);
        if (lane == 0 && jb + q < nj)
          out[r * ostride + jb + q] = s + (bias ? bias[j0 + jb + q] * (bscale ? bscale[r] : 1.f) : 0.f);
      }
  }
}

// out[r * ostride + i] = sum_{j in [ja, jb)} W[j * ldw + i0 + i] u[r * ustride + j]   (i < ni <= 64); u in shared memory.
// thread = (column i, one of 4 row groups), eight weight loads in flight per thread; `scratch` holds 4 * TL_RB * 64 floats.
// Ends with a __syncthreads.
__device__ void slice_matvec_t(const float* __restrict__ W, long long ldw, int i0, int ni, int ja, int jb, const float* u,
                               int ustride, float* out, int ostride, float* scratch) {
  const int i = threadIdx.x & 63, jp = threadIdx.x >> 6;
  float acc[TL_RB];
#pragma unroll
  for (int r = 0; r < TL_RB; ++r) acc[r] = 0.f;
  if (i < ni) {
    const float* wc = W + i0 + i;
    int j = ja + jp;
    for (; j + 28 < jb; j += 32) {
      float w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) w[e] = wc[(long long)(j + 4 * e) * ldw];
#pragma unroll
      for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int r = 0; r < TL_RB; ++r) acc[r] = fmaf(w[e], u[r * ustride + j + 4 * e], acc[r]);
    }
    for (; j < jb; j += 4) {
      const float w = wc[(long long)j * ldw];
#pragma unroll
      for (int r = 0; r < TL_RB; ++r) acc[r] = fmaf(w, u[r * ustride + j], acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < TL_RB; ++r) scratch[(jp * TL_RB + r) * 64 + i] = acc[r];
  __syncthreads();
  if (threadIdx.x < TL_RB * 64) {
    const int r = threadIdx.x >> 6;
    if (i < ni)
      out[r * ostride + i] = scratch[(0 * TL_RB + r) * 64 + i] + scratch[(1 * TL_RB + r) * 64 + i] +
                             scratch[(2 * TL_RB + r) * 64 + i] + scratch[(3 * TL_RB + r) * 64 + i];
  }
  __syncthreads();
}

// full[r][c * w8 + j] = slice_of_cta_c[r][j]: gathers the column slices of all 8 CTAs of the cluster (call after a cluster sync)
__device__ void gather_rows(float* full, int stride, int D, const float* own, int w8) {
  const uint32_t own_addr = smem_u32(own);
  for (int t = threadIdx.x; t < TL_RB * D; t += TL_THREADS) {
    const int r = t / D, c = t - r * D, rank = c / w8, j = c - rank * w8;
    full[r * stride + c] = ld_dsmem(own_addr + (uint32_t)(r * TL_MAXW + j) * 4, rank);
  }
  __syncthreads();
}
// sum over the cluster of one float per row (stat[r] of every CTA); call after a cluster sync
__device__ __forceinline__ float cluster_row_sum(const float* stat, int r) {
  float s = 0.f;
  const uint32_t a = smem_u32(stat + r);
#pragma unroll
  for (int c = 0; c < TL_CL; ++c) s += ld_dsmem(a, c);
  return s;
}

// ------------------------------------------------------------------------------------------------------------------
struct PrepParams {
  const float* query; const float* w_in; const float* b_in;   // [D], [3D, D], [3D]
  float* q0; float* qt; void* qt_img;                          // [D], [H, D], 16-bit [D/64][16][64] swizzled (or null)
  int D, H, fp16;
};
// grid (D/64, 8): CTA (kc, h) -> qt[h, 64 kc .. + 64) and image rows h (hi) / 8 + h (lo) of sub-tile kc; h >= H: zero rows
__global__ void __launch_bounds__(TL_THREADS) pool_prep_kernel(PrepParams p) {
  __shared__ float s_q[512];
  __shared__ float s_q0[512];
  __shared__ float s_part[4][64];
  const int kc = blockIdx.x, h = blockIdx.y, D = p.D, Dh = D / p.H;
  unsigned char* img = reinterpret_cast<unsigned char*>(p.qt_img);
  if (h >= p.H) {
    if (img && threadIdx.x < 64) {
      *reinterpret_cast<uint16_t*>(img + kc * 2048 + tl_sw128(h, threadIdx.x)) = 0;
      *reinterpret_cast<uint16_t*>(img + kc * 2048 + tl_sw128(8 + h, threadIdx.x)) = 0;
    }
    return;
  }
  for (int i = threadIdx.x; i < D; i += TL_THREADS) s_q[i] = p.query[i];
  __syncthreads();
  // q0_h[k] = W_q[h Dh + k, :] . query + b_q[h Dh + k]: one warp per row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int kb = warp * 4; kb < Dh; kb += TL_THREADS / 32 * 4) {       // four rows per warp and step: four requests in flight
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    const float* wr[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) wr[q] = p.w_in + (size_t)(h * Dh + min(kb + q, Dh - 1)) * D;
    for (int i = lane * 4; i < D; i += 128) {
      float4 w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = *reinterpret_cast<const float4*>(wr[q] + i);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        a[q] = fmaf(w[q].x, s_q[i], fmaf(w[q].y, s_q[i + 1], fmaf(w[q].z, s_q[i + 2], fmaf(w[q].w, s_q[i + 3], a[q]))));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float v = warp_sum(a[q]) + p.b_in[h * Dh + min(kb + q, Dh - 1)];
      if (lane == 0 && kb + q < Dh) {
        s_q0[kb + q] = v;
        if (kc == 0) p.q0[h * Dh + kb + q] = v;
      }
    }
  }
  __syncthreads();
  // qt[h, d] = sum_k W_k[h Dh + k, d] q0[k] / sqrt(Dh), d = 64 kc + (tid & 63), 4 groups of k
  const int dl = threadIdx.x & 63, kp = threadIdx.x >> 6;
  const float* wk = p.w_in + (size_t)D * D + (size_t)(h * Dh) * D + kc * 64 + dl;
  float a = 0.f;
#pragma unroll 8
  for (int k = kp; k < Dh; k += 4) a = fmaf(wk[(size_t)k * D], s_q0[k], a);
  s_part[kp][dl] = a;
  __syncthreads();
  if (threadIdx.x < 64) {
    const float v = (s_part[0][dl] + s_part[1][dl] + s_part[2][dl] + s_part[3][dl]) * rsqrtf((float)Dh);
    p.qt[(size_t)h * D + kc * 64 + dl] = v;
    if (img) {
      const uint16_t hi = tl_bits(v, p.fp16);
      *reinterpret_cast<uint16_t*>(img + kc * 2048 + tl_sw128(h, dl)) = hi;
      *reinterpret_cast<uint16_t*>(img + kc * 2048 + tl_sw128(8 + h, dl)) = tl_bits(v - tl_val(hi, p.fp16), p.fp16);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
struct TailFwdParams {
  const float* pm; const float* pl; const float* pl2; const float* pa;    // partials [B, S, H] x3, [B, S, H, D]
  const float* w_v; const float* b_v; const float* w_o; const float* b_o; // [D, D], [D] (w_v = in_proj_weight + 2 D D)
  const float* gamma; const float* beta; float eps;
  const float* w_p; const float* b_p; int Do;                              // optional output Linear [Do, D]
  float* xbar; float* m; float* l; float* sa;                              // [B, H, D], [B, H] x3
  float* o; float* yhat; float* rstd; float* yln;                          // [B, D], [B, D], [B], [B, D]
  void* out; int out_dtype;                                                // [B, Do or D]
  int B, S, H, D;
};

__global__ void __cluster_dims__(TL_CL, 1, 1) __launch_bounds__(TL_THREADS) pool_tail_fwd_kernel(TailFwdParams p) {
  extern __shared__ float tl_smem[];
  const int D = p.D, w8 = D / TL_CL, Dh = D / p.H;
  float* vec = tl_smem;                          // [RB][D]
  float* ownA = vec + TL_RB * D;                 // [RB][64] o slice
  float* ownB = ownA + TL_RB * TL_MAXW;          // y slice
  float* ownC = ownB + TL_RB * TL_MAXW;          // LayerNorm output slice
  float* stat1 = ownC + TL_RB * TL_MAXW;         // [RB]
  float* stat2 = stat1 + TL_RB;
  float* s_sa = stat2 + TL_RB;                   // [RB]
  const int rank = (int)cluster_ctarank(), b0 = blockIdx.y * TL_RB;
  const int j0 = rank * w8, h = j0 / Dh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- merge of the token splits for (row, head h) -> vec[r][:] = xbar[b, h, :] ----
  for (int r = 0; r < TL_RB; ++r) {
    const int b = b0 + r;
    if (b >= p.B) {
      for (int d = threadIdx.x; d < D; d += TL_THREADS) vec[r * D + d] = 0.f;
      if (threadIdx.x == 0) s_sa[r] = 0.f;
      continue;
    }
    float mx = -INFINITY;
    for (int s = 0; s < p.S; ++s) mx = fmaxf(mx, p.pm[((size_t)b * p.S + s) * p.H + h]);
    float lsum = 0.f, l2sum = 0.f;
    for (int s = 0; s < p.S; ++s) {
      const size_t slot = ((size_t)b * p.S + s) * p.H + h;
      const float ms = p.pm[slot];
      const float e = ms == -INFINITY ? 0.f : __expf(ms - mx);
      lsum = fmaf(p.pl[slot], e, lsum);
      if (p.pl2) l2sum = fmaf(p.pl2[slot], e, l2sum);
    }
    const float il = 1.f / lsum;
    for (int d = threadIdx.x; d < D; d += TL_THREADS) {
      float a = 0.f;
      for (int s = 0; s < p.S; ++s) {
        const size_t slot = ((size_t)b * p.S + s) * p.H + h;
        const float ms = p.pm[slot];
        a = fmaf(ms == -INFINITY ? 0.f : __expf(ms - mx), p.pa[slot * D + d], a);
      }
      const float xb = a * il;                   // all-masked row -> 0/0 = NaN like nn.MultiheadAttention
      vec[r * D + d] = xb;
      if (j0 % Dh == 0) p.xbar[((size_t)b * p.H + h) * D + d] = xb;
    }
    if (threadIdx.x == 0) {
      const float sav = p.pl2 ? l2sum * il : 1.f;
      s_sa[r] = sav;
      if (j0 % Dh == 0) { p.m[b * p.H + h] = mx; p.l[b * p.H + h] = lsum; p.sa[b * p.H + h] = sav; }
    }
  }
  __syncthreads();
  // ---- o slice = W_v[j0.., :] xbar_h + b_v sa ----
  slice_matvec(p.w_v, D, j0, w8, D, vec, D, ownA, TL_MAXW, p.b_v, s_sa);
  __syncthreads();
  for (int t = threadIdx.x; t < TL_RB * w8; t += TL_THREADS) {
    const int r = t / w8, j = t - r * w8;
    if (b0 + r < p.B) p.o[(size_t)(b0 + r) * D + j0 + j] = ownA[r * TL_MAXW + j];
  }
  cluster_sync_all();
  gather_rows(vec, D, D, ownA, w8);
  // ---- y slice = W_o[j0.., :] o + b_o ----
  slice_matvec(p.w_o, D, j0, w8, D, vec, D, ownB, TL_MAXW, p.b_o, nullptr);
  __syncthreads();
  // ---- LayerNorm over the full row: two passes over the cluster ----
  if (warp < TL_RB) {
    float s = 0.f;
    for (int j = lane; j < w8; j += 32) s += ownB[warp * TL_MAXW + j];
    s = warp_sum(s);
    if (lane == 0) stat1[warp] = s;
  }
  cluster_sync_all();
  float mean = 0.f;
  if (warp < TL_RB) {
    mean = cluster_row_sum(stat1, warp) / (float)D;
    float s = 0.f;
    for (int j = lane; j < w8; j += 32) { const float dv = ownB[warp * TL_MAXW + j] - mean; s = fmaf(dv, dv, s); }
    s = warp_sum(s);
    if (lane == 0) stat2[warp] = s;
  }
  cluster_sync_all();
  if (warp < TL_RB) {
    const int b = b0 + warp;
    const float rs = rsqrtf(cluster_row_sum(stat2, warp) / (float)D + p.eps);
    for (int j = lane; j < w8; j += 32) {
      const float yh = (ownB[warp * TL_MAXW + j] - mean) * rs;
      const float yl = fmaf(yh, p.gamma[j0 + j], p.beta[j0 + j]);
      ownC[warp * TL_MAXW + j] = yl;
      if (b < p.B) {
        p.yhat[(size_t)b * D + j0 + j] = yh;
        if (p.w_p) p.yln[(size_t)b * D + j0 + j] = yl;
        else st_any(p.out, (size_t)b * D + j0 + j, p.out_dtype, yl);
      }
    }
    if (lane == 0 && rank == 0 && b < p.B) p.rstd[b] = rs;
  }
  if (p.w_p) {
    // ---- out slice = W_p[jo0.., :] LN(y) + b_p ----
    cluster_sync_all();
    gather_rows(vec, D, D, ownC, w8);
    const int wo8 = p.Do / TL_CL, jo0 = rank * wo8;
    for (int c0 = 0; c0 < wo8; c0 += TL_MAXW) {
      const int nj = min(TL_MAXW, wo8 - c0);
      slice_matvec(p.w_p, D, jo0 + c0, nj, D, vec, D, ownA, TL_MAXW, p.b_p, nullptr);
      __syncthreads();
      for (int t = threadIdx.x; t < TL_RB * nj; t += TL_THREADS) {
        const int r = t / nj, j = t - r * nj;
        if (b0 + r < p.B) st_any(p.out, (size_t)(b0 + r) * p.Do + jo0 + c0 + j, p.out_dtype, ownA[r * TL_MAXW + j]);
      }
      __syncthreads();
    }
  }
  cluster_sync_all();      // no CTA leaves while a peer may still read its shared memory
}

// ------------------------------------------------------------------------------------------------------------------
struct TailBwdParams {
  const void* dout; int dout_dtype;                                        // [B, Do or D]
  const float* yhat; const float* rstd; const float* xbar; const float* sa;
  const float* w_v; const float* b_v; const float* w_o; const float* gamma; const float* w_p; int Do;
  const float* qt;                                                         // [H, D]
  float* dyln; float* dy; float* do_; float* dxbar; float* dsa; float* cdot;   // [B, D] x3, [B, H, D], [B, H] (or null), [B, H]
  void* w_img; int fp16;                                                   // [B][D/64][32][64] 16-bit swizzled (or null)
  int B, H, D;
};

__global__ void __cluster_dims__(TL_CL, 1, 1) __launch_bounds__(TL_THREADS) pool_tail_bwd_kernel(TailBwdParams p) {
  extern __shared__ float tl_smem[];
  const int D = p.D, w8 = D / TL_CL, Dh = D / p.H;
  float* vec = tl_smem;                          // [RB][max(D, Do)]
  const int vs = max(D, p.Do);
  float* ownA = vec + TL_RB * vs;                // dyln slice
  float* ownB = ownA + TL_RB * TL_MAXW;          // dy slice
  float* ownC = ownB + TL_RB * TL_MAXW;          // do slice
  float* scratch = ownC + TL_RB * TL_MAXW;       // [4][RB][64]
  float* stat1 = scratch + 4 * TL_RB * 64;       // [RB]
  float* stat2 = stat1 + TL_RB;
  float* cpart = stat2 + TL_RB;                  // [RB][8]
  const int rank = (int)cluster_ctarank(), b0 = blockIdx.y * TL_RB;
  const int i0 = rank * w8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- dyln slice: dout (no proj) or W_p^T dout ----
  if (p.w_p) {
    for (int t = threadIdx.x; t < TL_RB * p.Do; t += TL_THREADS) {
      const int r = t / p.Do, j = t - r * p.Do;
      vec[r * vs + j] = b0 + r < p.B ? ld_any(p.dout, (size_t)(b0 + r) * p.Do + j, p.dout_dtype) : 0.f;
    }
    __syncthreads();
    slice_matvec_t(p.w_p, D, i0, w8, 0, p.Do, vec, vs, ownA, TL_MAXW, scratch);
  } else {
    for (int t = threadIdx.x; t < TL_RB * w8; t += TL_THREADS) {
      const int r = t / w8, j = t - r * w8;
      ownA[r * TL_MAXW + j] = b0 + r < p.B ? ld_any(p.dout, (size_t)(b0 + r) * D + i0 + j, p.dout_dtype) : 0.f;
    }
    __syncthreads();
  }
  // ---- LayerNorm backward: dy = rstd (g - mean(g) - yhat mean(g yhat)), g = dyln gamma ----
  if (warp < TL_RB) {
    const int b = b0 + warp;
    float s1 = 0.f, s2 = 0.f;
    for (int j = lane; j < w8; j += 32) {
      const float dl = ownA[warp * TL_MAXW + j];
      const float g = dl * p.gamma[i0 + j];
      const float yh = b < p.B ? p.yhat[(size_t)b * D + i0 + j] : 0.f;
      s1 += g;
      s2 = fmaf(g, yh, s2);
      if (b < p.B) p.dyln[(size_t)b * D + i0 + j] = dl;
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { stat1[warp] = s1; stat2[warp] = s2; }
  }
  cluster_sync_all();
  if (warp < TL_RB) {
    const int b = b0 + warp;
    const float a1 = cluster_row_sum(stat1, warp) / (float)D, a2 = cluster_row_sum(stat2, warp) / (float)D;
    const float rs = b < p.B ? p.rstd[b] : 0.f;
    for (int j = lane; j < w8; j += 32) {
      const float g = ownA[warp * TL_MAXW + j] * p.gamma[i0 + j];
      const float yh = b < p.B ? p.yhat[(size_t)b * D + i0 + j] : 0.f;
      const float dyv = rs * (g - a1 - yh * a2);
      ownB[warp * TL_MAXW + j] = dyv;
      if (b < p.B) p.dy[(size_t)b * D + i0 + j] = dyv;
    }
  }
  cluster_sync_all();
  gather_rows(vec, vs, D, ownB, w8);                // vec[r][0..D) = dy rows (row stride vs)
  // ---- do slice = W_o^T dy ----
  slice_matvec_t(p.w_o, D, i0, w8, 0, D, vec, vs, ownC, TL_MAXW, scratch);
  for (int t = threadIdx.x; t < TL_RB * w8; t += TL_THREADS) {
    const int r = t / w8, j = t - r * w8;
    if (b0 + r < p.B) p.do_[(size_t)(b0 + r) * D + i0 + j] = ownC[r * TL_MAXW + j];
  }
  cluster_sync_all();
  gather_rows(vec, vs, D, ownC, w8);                // vec = do rows
  // ---- dsa[b, h] = sum_{j in head h} b_v[j] do[b, j] (only needed with attention dropout) ----
  if (p.dsa && rank == 0 && warp < TL_RB && b0 + warp < p.B)
    for (int h = 0; h < p.H; ++h) {
      float s = 0.f;
      for (int j = lane; j < Dh; j += 32) s = fmaf(p.b_v[h * Dh + j], vec[warp * vs + h * Dh + j], s);
      s = warp_sum(s);
      if (lane == 0) p.dsa[(b0 + warp) * p.H + h] = s;
    }
  // ---- dxbar[b, h, i0..] = sum_{j in head h} W_v[j, i] do[b, j]; operand image rows; c partials ----
  unsigned char* img = reinterpret_cast<unsigned char*>(p.w_img);
  const size_t img_row = (size_t)(D / 64) * 4096;
  for (int h = 0; h < p.H; ++h) {
    slice_matvec_t(p.w_v, D, i0, w8, h * Dh, (h + 1) * Dh, vec, vs, ownA, TL_MAXW, scratch);
    if (warp < TL_RB) {
      const int b = b0 + warp;
      float cs = 0.f;
      for (int j = lane; j < w8; j += 32) {
        const float dv = ownA[warp * TL_MAXW + j];
        if (b < p.B) {
          const int d = i0 + j;
          p.dxbar[((size_t)b * p.H + h) * D + d] = dv;
          cs = fmaf(dv, p.xbar[((size_t)b * p.H + h) * D + d], cs);
          if (img) {
            unsigned char* sub = img + b * img_row + (d >> 6) * 4096;
            const float qv = p.qt[(size_t)h * D + d];
            const uint16_t qh = tl_bits(qv, p.fp16), dh = tl_bits(dv, p.fp16);
            *reinterpret_cast<uint16_t*>(sub + tl_sw128(h, d & 63)) = qh;
            *reinterpret_cast<uint16_t*>(sub + tl_sw128(8 + h, d & 63)) = dh;
            *reinterpret_cast<uint16_t*>(sub + tl_sw128(16 + h, d & 63)) = tl_bits(qv - tl_val(qh, p.fp16), p.fp16);
            *reinterpret_cast<uint16_t*>(sub + tl_sw128(24 + h, d & 63)) = tl_bits(dv - tl_val(dh, p.fp16), p.fp16);
          }
        }
      }
      cs = warp_sum(cs);
      if (lane == 0) cpart[warp * 8 + h] = cs;
    }
    __syncthreads();
  }
  if (img && p.H < 8)
    for (int t = threadIdx.x; t < TL_RB * (8 - p.H) * w8; t += TL_THREADS) {
      const int r = t / ((8 - p.H) * w8), rem = t - r * (8 - p.H) * w8, h = p.H + rem / w8, d = i0 + rem % w8;
      if (b0 + r < p.B) {
        unsigned char* sub = img + (b0 + r) * img_row + (d >> 6) * 4096;
        for (int g = 0; g < 4; ++g) *reinterpret_cast<uint16_t*>(sub + tl_sw128(8 * g + h, d & 63)) = 0;
      }
    }
  cluster_sync_all();
  if (rank == 0 && threadIdx.x < TL_RB * 8) {
    const int r = threadIdx.x >> 3, h = threadIdx.x & 7;
    if (b0 + r < p.B && h < p.H) {
      float s = 0.f;
      const uint32_t a = smem_u32(cpart + r * 8 + h);
      for (int c = 0; c < TL_CL; ++c) s += ld_dsmem(a, c);
      p.cdot[(b0 + r) * p.H + h] = s;
    }
  }
  cluster_sync_all();
}

// ------------------------------------------------------------------------------------------------------------------
struct ParamGradParams {
  const float* dy; const float* o; const float* do_; const float* xbar; const float* sa; const float* dyln; const float* yhat;
  const void* dout; int dout_dtype; const float* yln; int Do;
  float* dw_o; float* db_o; float* dw_v; float* db_v; float* dgamma; float* dbeta; float* dw_p; float* db_p;
  int B, H, D, use_sa;
};
// grid (rows / 8, 3): y = 0 dW_o / db_o, y = 1 dW_v / db_v, y = 2 dgamma / dbeta (+ dW_p / db_p). 8 weight rows per CTA,
// threads over the D columns, loop over the batch.
__global__ void __launch_bounds__(TL_THREADS) pool_param_grads_kernel(ParamGradParams p) {
  __shared__ float coef[8][128];       // [row][b chunk]
  const int D = p.D, Dh = D / p.H, which = blockIdx.y, j0 = blockIdx.x * 8;
  if (which == 2 && !p.dw_p) {
    // dgamma / dbeta only: CTA x handles columns [x * 8 * ..): spread the D columns over the grid
    const int cols = (D + gridDim.x - 1) / gridDim.x;
    for (int i = blockIdx.x * cols + threadIdx.x; i < min(D, (int)(blockIdx.x + 1) * cols); i += TL_THREADS) {
      float g = 0.f, bb = 0.f;
      for (int b = 0; b < p.B; ++b) { const float dl = p.dyln[(size_t)b * D + i]; g = fmaf(dl, p.yhat[(size_t)b * D + i], g); bb += dl; }
      p.dgamma[i] = g; p.dbeta[i] = bb;
    }
    return;
  }
  const int rows = which == 2 ? p.Do : D;
  if (j0 >= rows) {
    return;
  }
  float acc[8][2];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r][0] = acc[r][1] = 0.f;
  float bsum = 0.f;                    // bias gradient of row j0 + (tid >> 5) accumulated by lane 0..: done below per chunk
  for (int bc = 0; bc < p.B; bc += 128) {
    const int nb = min(128, p.B - bc);
    __syncthreads();
    for (int t = threadIdx.x; t < 8 * nb; t += TL_THREADS) {
      const int r = t / nb, b = bc + t - r * nb, j = j0 + r;
      float c = 0.f;
      if (j < rows) c = which == 0 ? p.dy[(size_t)b * D + j] : which == 1 ? p.do_[(size_t)b * D + j]
                                                                        : ld_any(p.dout, (size_t)b * p.Do + j, p.dout_dtype);
      coef[r][b - bc] = c;
    }
    __syncthreads();
#pragma unroll 4
    for (int b = 0; b < nb; ++b) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = threadIdx.x + e * TL_THREADS;
        if (i < D) {
          if (which == 1) {
            // rows j0 .. j0 + 7 lie in one head when Dh % 8 == 0
            const float v = p.xbar[((size_t)(bc + b) * p.H + j0 / Dh) * D + i];
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r][e] = fmaf(coef[r][b], v, acc[r][e]);
          } else {
            const float v = which == 0 ? p.o[(size_t)(bc + b) * D + i] : p.yln[(size_t)(bc + b) * D + i];
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r][e] = fmaf(coef[r][b], v, acc[r][e]);
          }
        }
      }
    }
    if (threadIdx.x < 8) {
      const int r = threadIdx.x;
      for (int b = 0; b < nb; ++b)
        bsum = fmaf(coef[r][b], which == 1 ? (p.use_sa ? p.sa[(bc + b) * p.H + (j0 + r) / Dh] : 1.f) : 1.f, bsum);
    }
  }
  float* dw = which == 0 ? p.dw_o : which == 1 ? p.dw_v : p.dw_p;
  float* db = which == 0 ? p.db_o : which == 1 ? p.db_v : p.db_p;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int i = threadIdx.x + e * TL_THREADS;
    if (i < D)
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (j0 + r < rows) dw[(size_t)(j0 + r) * D + i] = acc[r][e];
  }
  if (threadIdx.x < 8 && j0 + threadIdx.x < rows) db[j0 + threadIdx.x] = bsum;
  if (which == 2 && blockIdx.x == 0) {
    for (int i = threadIdx.x; i < D; i += TL_THREADS) {
      float g = 0.f, bb = 0.f;
      for (int b = 0; b < p.B; ++b) { const float dl = p.dyln[(size_t)b * D + i]; g = fmaf(dl, p.yhat[(size_t)b * D + i], g); bb += dl; }
      p.dgamma[i] = g; p.dbeta[i] = bb;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
struct QGradParams {
  const float* part_dq; int nparts;      // [nparts, H, D] (nparts = B * splits)
  const float* q0; const float* query; const float* w_in;
  float* dqt;                            // [H, D]
  float* dw_in; float* db_in; float* dquery;   // [3D, D] (rows [0, 2D) written here), [3D] (entries [0, 2D)), [D]
  int H, D;
};
// one cluster of 8 CTAs: CTA r owns weight rows / columns [r D/8, (r+1) D/8)
__global__ void __cluster_dims__(TL_CL, 1, 1) __launch_bounds__(TL_THREADS) pool_qgrads_kernel(QGradParams p) {
  __shared__ __align__(16) float s_dqt[512];
  __shared__ __align__(16) float s_full[TL_RB * 512];
  __shared__ float s_own[TL_RB * TL_MAXW];
  __shared__ float s_scr[4 * TL_RB * 64];
  const int D = p.D, w8 = D / TL_CL, Dh = D / p.H;
  const int rank = (int)cluster_ctarank(), j0 = rank * w8, h = j0 / Dh;
  const float isq = rsqrtf((float)Dh);
  // dqt[h, :] = sum over the partials: thread = (4 channels, one of TL_THREADS / (D / 4) partial groups), 8 loads in flight
  {
    const int nv = D / 4, v4 = threadIdx.x % nv, grp = threadIdx.x / nv, ngrp = TL_THREADS / nv;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (grp < ngrp) {
      const float4* src = reinterpret_cast<const float4*>(p.part_dq + (size_t)h * D) + v4;
      const size_t stride = (size_t)p.H * D / 4;
      int sidx = grp;
      for (; sidx + 7 * ngrp < p.nparts; sidx += 8 * ngrp) {
        float4 t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = src[(size_t)(sidx + e * ngrp) * stride];
#pragma unroll
        for (int e = 0; e < 8; ++e) { a.x += t[e].x; a.y += t[e].y; a.z += t[e].z; a.w += t[e].w; }
      }
      for (; sidx < p.nparts; sidx += ngrp) {
        const float4 t = src[(size_t)sidx * stride];
        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      }
      reinterpret_cast<float4*>(s_full)[grp * nv + v4] = a;       // s_full is free until the dq0 stage
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += TL_THREADS) {
      float t = 0.f;
      for (int g2 = 0; g2 < ngrp; ++g2) t += s_full[g2 * D + d];
      s_dqt[d] = t;
      if (j0 % Dh == 0) p.dqt[(size_t)h * D + d] = t;
    }
  }
  __syncthreads();
  // dW_k[j, :] = q0[j] dqt[h, :] / sqrt(Dh);  dq0[j] = W_k[j, :] . dqt[h, :] / sqrt(Dh)   (j in the slice; b_k gets no gradient)
  for (int t = threadIdx.x; t < w8 * D; t += TL_THREADS) {
    const int j = t / D, d = t - j * D;
    p.dw_in[(size_t)(D + j0 + j) * D + d] = p.q0[j0 + j] * s_dqt[d] * isq;
  }
  for (int t = threadIdx.x; t < TL_RB * D; t += TL_THREADS) s_full[t] = t < D ? s_dqt[t] : 0.f;
  __syncthreads();
  slice_matvec(p.w_in + (size_t)D * D, D, j0, w8, D, s_full, D, s_own, TL_MAXW, nullptr, nullptr);
  __syncthreads();
  for (int j = threadIdx.x; j < w8; j += TL_THREADS) {
    s_own[j] *= isq;
    p.db_in[j0 + j] = s_own[j];          // db_q = dq0
    p.db_in[D + j0 + j] = 0.f;           // db_k: the key bias shifts every score of a head equally
  }
  cluster_sync_all();
  gather_rows(s_full, D, D, s_own, w8);     // s_full[0][0..D) = dq0 (rows 1.. unused)
  // dW_q[j, i] = dq0[j] query[i] (rows of the slice);  dquery[i] = sum_j W_q[j, i] dq0[j] (columns of the slice)
  for (int t = threadIdx.x; t < w8 * D; t += TL_THREADS) {
    const int j = t / D, i = t - j * D;
    p.dw_in[(size_t)(j0 + j) * D + i] = s_full[j0 + j] * p.query[i];
  }
  slice_matvec_t(p.w_in, D, j0, w8, 0, D, s_full, D, s_own, TL_MAXW, s_scr);
  for (int j = threadIdx.x; j < w8; j += TL_THREADS) p.dquery[j0 + j] = s_own[j];
  cluster_sync_all();
}

}  // namespace b2

namespace b2host {
using namespace b2;

bool pooltail_ok(int D, int H, int Do) {
  return D % 128 == 0 && D <= 512 && H >= 1 && H <= 8 && 8 % H == 0 && (Do == 0 || (Do % 8 == 0 && Do <= 4096));
}

int pool_prep(const float* query, const float* w_in, const float* b_in, int D, int H, float* q0, float* qt, void* qt_img,
              int fp16, cudaStream_t s) {
  if (!query || !w_in || !b_in || !q0 || !qt || !pooltail_ok(D, H, 0)) return B2_EINVAL;
  PrepParams p{query, w_in, b_in, q0, qt, qt_img, D, H, fp16};
  pool_prep_kernel<<<dim3(D / 64, 8), TL_THREADS, 0, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int pool_tail_fwd(const float* pm, const float* pl, const float* pl2, const float* pa, int B, int S, int H, int D,
                  const float* w_v, const float* b_v, const float* w_o, const float* b_o, const float* gamma,
                  const float* beta, float eps, const float* w_p, const float* b_p, int Do, float* xbar, float* m, float* l,
                  float* sa, float* o, float* yhat, float* rstd, float* yln, void* out, int out_dtype, cudaStream_t s) {
  if (!pm || !pl || !pa || !w_v || !b_v || !w_o || !b_o || !gamma || !beta || !xbar || !m || !l || !sa || !o || !yhat || !rstd ||
      !out || B <= 0 || S <= 0 || !pooltail_ok(D, H, w_p ? Do : 0) || (w_p && (!b_p || !yln)))
    return B2_EINVAL;
  TailFwdParams p{pm, pl, pl2, pa, w_v, b_v, w_o, b_o, gamma, beta, eps, w_p, b_p, w_p ? Do : 0, xbar, m, l, sa, o, yhat, rstd,
                  yln, out, out_dtype, B, S, H, D};
  const size_t smem = (size_t)(TL_RB * D + 3 * TL_RB * TL_MAXW + 3 * TL_RB) * sizeof(float);
  pool_tail_fwd_kernel<<<dim3(TL_CL, (B + TL_RB - 1) / TL_RB), TL_THREADS, smem, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int pool_tail_bwd(const void* dout, int dout_dtype, const float* yhat, const float* rstd, const float* xbar, const float* sa,
                  const float* w_v, const float* b_v, const float* w_o, const float* gamma, const float* w_p, int Do,
                  const float* qt, int B, int H, int D, float* dyln, float* dy, float* do_, float* dxbar, float* dsa,
                  float* cdot, void* w_img, int fp16, cudaStream_t s) {
  if (!dout || !yhat || !rstd || !xbar || !w_v || !b_v || !w_o || !gamma || !qt || !dyln || !dy || !do_ || !dxbar || !cdot ||
      B <= 0 || !pooltail_ok(D, H, w_p ? Do : 0))
    return B2_EINVAL;
  TailBwdParams p{dout, dout_dtype, yhat, rstd, xbar, sa, w_v, b_v, w_o, gamma, w_p, w_p ? Do : 0, qt, dyln, dy, do_, dxbar, dsa,
                  cdot, w_img, fp16, B, H, D};
  const int vs = D > p.Do ? D : p.Do;
  const size_t smem = (size_t)(TL_RB * vs + 3 * TL_RB * TL_MAXW + 4 * TL_RB * 64 + 2 * TL_RB + TL_RB * 8) * sizeof(float);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(pool_tail_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return B2_ECUDA;
  pool_tail_bwd_kernel<<<dim3(TL_CL, (B + TL_RB - 1) / TL_RB), TL_THREADS, smem, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int pool_param_grads(const float* dy, const float* o, const float* do_, const float* xbar, const float* sa, int use_sa,
                     const float* dyln, const float* yhat, const void* dout, int dout_dtype, const float* yln, int Do, int B,
                     int H, int D, float* dw_o, float* db_o, float* dw_v, float* db_v, float* dgamma, float* dbeta, float* dw_p,
                     float* db_p, cudaStream_t s) {
  if (!dy || !o || !do_ || !xbar || !dyln || !yhat || !dw_o || !db_o || !dw_v || !db_v || !dgamma || !dbeta || B <= 0 ||
      !pooltail_ok(D, H, dw_p ? Do : 0) || (D / H) % 8 != 0 || (use_sa && !sa) || (dw_p && (!dout || !yln || !db_p)))
    return B2_EINVAL;
  ParamGradParams p{dy, o, do_, xbar, sa, dyln, yhat, dout, dout_dtype, yln, dw_p ? Do : 0, dw_o, db_o, dw_v, db_v, dgamma, dbeta,
                    dw_p, db_p, B, H, D, use_sa};
  const int rows = dw_p && Do > D ? Do : D;
  pool_param_grads_kernel<<<dim3((rows + 7) / 8, 3), TL_THREADS, 0, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int pool_qgrads(const float* part_dq, int nparts, const float* q0, const float* query, const float* w_in, int H, int D,
                float* dqt, float* dw_in, float* db_in, float* dquery, cudaStream_t s) {
  if (!part_dq || nparts <= 0 || !q0 || !query || !w_in || !dqt || !dw_in || !db_in || !dquery || !pooltail_ok(D, H, 0))
    return B2_EINVAL;
  QGradParams p{part_dq, nparts, q0, query, w_in, dqt, dw_in, db_in, dquery, H, D};
  pool_qgrads_kernel<<<dim3(TL_CL), TL_THREADS, 0, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
