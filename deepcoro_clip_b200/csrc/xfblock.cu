// Pre-LN transformer block over the <= 16 views of a study (reference models/video_aggregator.py:7-54: LayerNorm ->
// nn.MultiheadAttention -> dropout -> residual, LayerNorm -> Linear(D, 4D) -> GELU -> dropout -> Linear(4D, D) -> dropout ->
// residual), forward and backward, fp32. In PyTorch this is ~75 launches forward + backward per block on
// [B, N <= 15, 512] operands — pure launch latency (2.5 ms of the 3.9 ms C3 step of round 1).
//
// One cluster of 8 CTAs per study. CTA r owns the column slice [r D/8, (r+1) D/8) of every D-wide row vector and
// [r F/8, (r+1) F/8) of the hidden layer (F = 4D): each weight slice is read once per cluster, full rows are exchanged through
// distributed shared memory (ld.shared::cluster), the N x N attention of a head is formed from the partial dot products of
// the CTAs that share the head. Everything of the study stays on chip between the first read of x and the last write.
//   xfblock_fwd   : out = block(x); saves the tensors the backward needs (normalised rows, q/k/v, probabilities, ...).
//   xfblock_bwd   : dx and the row-level gradients (d qkv, d attn-out, d hidden, d fc2-out, d LN outputs).
//   xfblock_wgrad : all parameter gradients as rank-(B N) updates dW[j, i] = sum_r a[r, j] b[r, i] (+ bias / LN sums).
// Dropout (training, p > 0) uses the counter-based keep mask of the attention pool (common.cuh attn_keep) at the four
// sites; forward and backward regenerate it.
#include "common.cuh"
#include "host_api.h"

namespace b2 {

constexpr int XB_UNROLL = 2;
// 8 warps. Measured with 16 warps (r02_xfblock_variants.log): one block fwd + bwd 253 us instead of 269 us at 8 studies x 4
// views, 598 instead of 722 us at 16 views, but 595 instead of 455 us at 32 studies (one CTA per SM: 256 CTAs take two waves).
constexpr int XB_THREADS = 256;
constexpr int XB_TG = XB_THREADS / 64;  // row groups of the transposed mat-vec
constexpr int WG_THREADS = 256;         // weight-gradient kernel
constexpr int XB_CL = 8;
constexpr int XB_W = 64;         // widest D slice (D <= 512)
constexpr int XB_FW = 256;       // widest hidden slice (F <= 2048)

__device__ __forceinline__ float xb_ld_dsmem(uint32_t local_addr, uint32_t rank) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(mapa_cluster(local_addr, rank)) : "memory");
  return v;
}
__device__ __forceinline__ float xb_keep(float p, unsigned long long seed, int site, int a, int b) {
  if (p <= 0.f) return 1.f;
  return attn_keep(seed + 0x51ED27ull * (unsigned)(site + 1), a, b, p) ? 1.f / (1.f - p) : 0.f;
}
__device__ __forceinline__ float xb_gelu(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float xb_gelu_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// out[r * ostride + j] = bias[j0 + j] + sum_i W[(j0 + j) * ldw + i] v[r * vstride + i]   (j < nj, K % 128 == 0).
// A warp takes FOUR output columns at a time: the four weight rows are loaded together (4 independent 512-byte requests per
// warp and step in flight instead of one — these kernels are latency-bound on the weight reads) and share the reads of v.
template <int RB>
__device__ void xb_matvec(const float* __restrict__ W, long long ldw, int j0, int nj, int K, const float* v, int vstride,
                          float* out, int ostride, const float* __restrict__ bias) {
  constexpr int JU = RB <= 8 ? 4 : 2;      // output columns per warp and step (register budget at RB = 16)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = XB_THREADS / 32;
  for (int jb = warp * JU; jb < nj; jb += nw * JU) {
    float acc[JU][RB];
#pragma unroll
    for (int q = 0; q < JU; ++q)
#pragma unroll
      for (int r = 0; r < RB; ++r) acc[q][r] = 0.f;
    const float* wr[JU];
#pragma unroll
    for (int q = 0; q < JU; ++q) wr[q] = W + (long long)(j0 + min(jb + q, nj - 1)) * ldw;
#pragma unroll XB_UNROLL
    for (int i = lane * 4; i < K; i += 128) {
      float4 w[JU];
#pragma unroll
      for (int q = 0; q < JU; ++q) w[q] = *reinterpret_cast<const float4*>(wr[q] + i);
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const float4 x = *reinterpret_cast<const float4*>(v + r * vstride + i);
#pragma unroll
        for (int q = 0; q < JU; ++q)
          acc[q][r] = fmaf(w[q].x, x.x, fmaf(w[q].y, x.y, fmaf(w[q].z, x.z, fmaf(w[q].w, x.w, acc[q][r]))));
      }
    }
#pragma unroll
    for (int q = 0; q < JU; ++q)
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const float s = warp_sum(acc[q][r]);
        if (lane == 0 && jb + q < nj) out[r * ostride + jb + q] = s + (bias ? bias[j0 + jb + q] : 0.f);
      }
  }
  __syncthreads();
}

// out[r * ostride + i] (+)= sum_{j < J} W[j * ldw + i0 + i] u[r * ustride + j]   (i < ni <= 64): thread = (column, 1 of XB_TG
// row groups), eight weight loads in flight per thread; `scratch` = XB_TG * RB * 64 floats. Ends with __syncthreads.
template <int RB>
__device__ void xb_matvec_t(const float* __restrict__ W, long long ldw, int i0, int ni, int J, const float* u, int ustride,
                            float* out, int ostride, float* scratch, bool accumulate) {
  const int i = threadIdx.x & 63, jp = threadIdx.x >> 6;
  float acc[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) acc[r] = 0.f;
  if (i < ni) {
    const float* wc = W + i0 + i;
    int j = jp;
    for (; j + 7 * XB_TG < J; j += 8 * XB_TG) {
      float w[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) w[e] = wc[(long long)(j + XB_TG * e) * ldw];
#pragma unroll
      for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int r = 0; r < RB; ++r) acc[r] = fmaf(w[e], u[r * ustride + j + XB_TG * e], acc[r]);
    }
    for (; j < J; j += XB_TG) {
      const float w = wc[(long long)j * ldw];
#pragma unroll
      for (int r = 0; r < RB; ++r) acc[r] = fmaf(w, u[r * ustride + j], acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < RB; ++r) scratch[(jp * RB + r) * 64 + i] = acc[r];
  __syncthreads();
  for (int t = threadIdx.x; t < RB * 64; t += XB_THREADS) {
    const int r = t >> 6, ii = t & 63;
    if (ii < ni) {
      float s = 0.f;
#pragma unroll
      for (int g = 0; g < XB_TG; ++g) s += scratch[(g * RB + r) * 64 + ii];
      out[r * ostride + ii] = accumulate ? out[r * ostride + ii] + s : s;
    }
  }
  __syncthreads();
}

// full[r * stride + c * w + j] = own_of_cta_c[r * ows + j]: column slices of all 8 CTAs (call after a cluster sync)
template <int RB>
__device__ void xb_gather(float* full, int stride, int ncols, const float* own, int ows, int w) {
  const uint32_t own_addr = smem_u32(own);
  for (int t = threadIdx.x; t < RB * ncols; t += XB_THREADS) {
    const int r = t / ncols, c = t - r * ncols, rank = c / w, j = c - rank * w;
    full[r * stride + c] = xb_ld_dsmem(own_addr + (uint32_t)(r * ows + j) * 4, rank);
  }
  __syncthreads();
}
__device__ __forceinline__ float xb_cluster_sum(const float* stat, int r) {
  float s = 0.f;
  const uint32_t a = smem_u32(stat + r);
#pragma unroll
  for (int c = 0; c < XB_CL; ++c) s += xb_ld_dsmem(a, c);
  return s;
}

struct XbParams {
  const float* x; float* out;                      // [B, N, D]
  const unsigned char* mask; long long mb;         // key padding [B, N] (non-zero = ignore) or null
  const float* ln1w; const float* ln1b; const float* w_in; const float* b_in; const float* w_o; const float* b_o;
  const float* ln2w; const float* ln2b; const float* w1; const float* b1; const float* w2; const float* b2;
  float eps1, eps2;
  // saved by the forward, read by the backward (rows R = B * N)
  float* xhat1; float* rstd1; float* h1;           // [R, D], [R], [R, D]
  float* qkv;                                      // [R, 3D]
  float* attn;                                     // [B, H, N, N] softmax probabilities (before dropout)
  float* o;                                        // [R, D] attention output before the out-projection
  float* x1;                                       // [R, D]
  float* xhat2; float* rstd2; float* h2;           // [R, D], [R], [R, D]
  float* z; float* u;                              // [R, F] pre-activation, [R, F] GELU (+ dropout) output
  // backward only
  const float* dout; float* dx;                    // [B, N, D]
  float* d_f2; float* d_z; float* d_ao; float* d_qkv; float* d_h2; float* d_h1;   // [R, D], [R, F], [R, D], [R, 3D], [R, D], [R, D]
  int B, N, D, H, F;
  float drop_p; unsigned long long seed;
};

// shared memory (floats): BIG [RB][F] (x rows | LN rows / gathered rows; later the gathered hidden rows) |
// qs, ks, vs, os, x1s, ts [RB][64] | zs [RB][256] | sp, aa [RB][RB] | stat1, stat2 [RB] | scratch [XB_TG][RB][64]
template <int RB>
struct XbSmem {
  static constexpr int big = 0;
  static constexpr int slices = RB * 2048;                  // F <= 2048
  static constexpr int zs = slices + 6 * RB * XB_W;
  static constexpr int sp = zs + RB * XB_FW;
  static constexpr int aa = sp + RB * RB;
  static constexpr int st = aa + RB * RB;
  static constexpr int scratch = st + 2 * RB;
  static constexpr int total = scratch + XB_TG * RB * 64;
};

template <int RB>
__global__ void __cluster_dims__(XB_CL, 1, 1) __launch_bounds__(XB_THREADS, 1) xfblock_fwd_kernel(XbParams p) {
  extern __shared__ __align__(16) float xb_smem[];
  using L = XbSmem<RB>;
  const int D = p.D, F = p.F, N = p.N, w8 = D / XB_CL, f8 = F / XB_CL, Dh = D / p.H;
  float* X = xb_smem + L::big;            // [RB][D]
  float* Hb = X + RB * D;                 // [RB][D]
  float* U = xb_smem + L::big;            // [RB][F] (aliases X, Hb once they are dead)
  float* qs = xb_smem + L::slices; float* ks = qs + RB * XB_W; float* vs = ks + RB * XB_W; float* os = vs + RB * XB_W;
  float* x1s = os + RB * XB_W; float* ts = x1s + RB * XB_W;
  float* zs = xb_smem + L::zs; float* sp = xb_smem + L::sp; float* aa = xb_smem + L::aa;
  float* stat1 = xb_smem + L::st; float* stat2 = stat1 + RB;
  const int rank = (int)cluster_ctarank(), b = blockIdx.y, c0 = rank * w8, hh = c0 / Dh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t row0 = (size_t)b * N;
  const unsigned char* mk = p.mask ? p.mask + (long long)b * p.mb : nullptr;

  // ---- x rows, LayerNorm 1 (every CTA, full rows) ----
  for (int t = threadIdx.x; t < RB * D; t += XB_THREADS) {
    const int r = t / D, d = t - r * D;
    X[t] = r < N ? p.x[(row0 + r) * D + d] : 0.f;
  }
  __syncthreads();
  for (int r = warp; r < RB; r += XB_THREADS / 32) {
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s += X[r * D + d];
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
    for (int d = lane; d < D; d += 32) { const float dv = X[r * D + d] - mean; q = fmaf(dv, dv, q); }
    const float rs = rsqrtf(warp_sum(q) / (float)D + p.eps1);
    for (int d = lane; d < D; d += 32) {
      const float xh = (X[r * D + d] - mean) * rs;
      const float h = fmaf(xh, p.ln1w[d], p.ln1b[d]);
      Hb[r * D + d] = h;
      if (rank == 0 && r < N) { p.xhat1[(row0 + r) * D + d] = xh; p.h1[(row0 + r) * D + d] = h; }
    }
    if (rank == 0 && lane == 0 && r < N) p.rstd1[row0 + r] = rs;
  }
  __syncthreads();
  // ---- q, k, v slices ----
  xb_matvec<RB>(p.w_in, D, c0, w8, D, Hb, D, qs, XB_W, p.b_in);
  xb_matvec<RB>(p.w_in + (size_t)D * D, D, c0, w8, D, Hb, D, ks, XB_W, p.b_in + D);
  xb_matvec<RB>(p.w_in + (size_t)2 * D * D, D, c0, w8, D, Hb, D, vs, XB_W, p.b_in + 2 * D);
  for (int t = threadIdx.x; t < RB * w8; t += XB_THREADS) {
    const int r = t / w8, j = t - r * w8;
    if (r < N) {
      float* dst = p.qkv + (row0 + r) * 3 * D + c0 + j;
      dst[0] = qs[r * XB_W + j]; dst[D] = ks[r * XB_W + j]; dst[2 * D] = vs[r * XB_W + j];
    }
  }
  // ---- partial scores of head hh over this slice ----
  for (int t = threadIdx.x; t < RB * RB; t += XB_THREADS) {
    const int i = t / RB, j = t - i * RB;
    float s = 0.f;
    for (int c = 0; c < w8; ++c) s = fmaf(qs[i * XB_W + c], ks[j * XB_W + c], s);
    sp[t] = s;
  }
  cluster_sync_all();
  {
    const int per = Dh / w8, first = hh * per;       // CTAs of this head: first .. first + per - 1
    const float isq = rsqrtf((float)Dh);
    for (int i = warp; i < RB; i += XB_THREADS / 32) {
      // lane = key j (RB <= 32)
      float s = -INFINITY;
      if (lane < N && !(mk && mk[lane])) {
        s = 0.f;
        const uint32_t a = smem_u32(sp + i * RB + lane);
        for (int c = 0; c < per; ++c) s += xb_ld_dsmem(a, first + c);
        s *= isq;
      }
      const float m = warp_max(s);
      const float e = __expf(s - m);                 // all keys masked: exp(-inf + inf) = NaN like nn.MultiheadAttention
      const float pr = e / warp_sum(lane < RB ? e : 0.f);
      if (lane < RB) {
        aa[i * RB + lane] = (lane < N ? pr : 0.f) * xb_keep(p.drop_p, p.seed, 0, (b * p.H + hh) * N + i, lane);
        if (c0 % Dh == 0 && i < N && lane < N) p.attn[(((size_t)b * p.H + hh) * N + i) * N + lane] = pr;
      }
    }
  }
  __syncthreads();
  // ---- attention output slice, saved; out-projection needs full rows ----
  for (int t = threadIdx.x; t < RB * w8; t += XB_THREADS) {
    const int i = t / w8, c = t - i * w8;
    float s = 0.f;
    for (int j = 0; j < N; ++j) s = fmaf(aa[i * RB + j], vs[j * XB_W + c], s);
    os[i * XB_W + c] = s;
    if (i < N) p.o[(row0 + i) * D + c0 + c] = s;
  }
  cluster_sync_all();
  xb_gather<RB>(Hb, D, D, os, XB_W, w8);
  xb_matvec<RB>(p.w_o, D, c0, w8, D, Hb, D, ts, XB_W, p.b_o);
  for (int t = threadIdx.x; t < RB * w8; t += XB_THREADS) {
    const int r = t / w8, j = t - r * w8;
    const float v = X[r * D + c0 + j] + ts[r * XB_W + j] * xb_keep(p.drop_p, p.seed, 1, (int)(row0 + r), c0 + j);
    x1s[r * XB_W + j] = v;
    if (r < N) p.x1[(row0 + r) * D + c0 + j] = v;
  }
  __syncthreads();
  // ---- LayerNorm 2 over full rows: two exchanges ----
  for (int r = warp; r < RB; r += XB_THREADS / 32) {
    float s = 0.f;
    for (int j = lane; j < w8; j += 32) s += x1s[r * XB_W + j];
    s = warp_sum(s);
    if (lane == 0) stat1[r] = s;
  }
  cluster_sync_all();
  float mean_r[2] = {0.f, 0.f};
  for (int r = warp, k = 0; r < RB; r += XB_THREADS / 32, ++k) {
    const float mean = xb_cluster_sum(stat1, r) / (float)D;
    mean_r[k] = mean;
    float q = 0.f;
    for (int j = lane; j < w8; j += 32) { const float dv = x1s[r * XB_W + j] - mean; q = fmaf(dv, dv, q); }
    q = warp_sum(q);
    if (lane == 0) stat2[r] = q;
  }
  cluster_sync_all();
  for (int r = warp, k = 0; r < RB; r += XB_THREADS / 32, ++k) {
    const float rs = rsqrtf(xb_cluster_sum(stat2, r) / (float)D + p.eps2);
    for (int j = lane; j < w8; j += 32) {
      const float xh = (x1s[r * XB_W + j] - mean_r[k]) * rs;
      const float h = fmaf(xh, p.ln2w[c0 + j], p.ln2b[c0 + j]);
      ts[r * XB_W + j] = h;
      if (r < N) { p.xhat2[(row0 + r) * D + c0 + j] = xh; p.h2[(row0 + r) * D + c0 + j] = h; }
    }
    if (rank == 0 && lane == 0 && r < N) p.rstd2[row0 + r] = rs;
  }
  cluster_sync_all();
  xb_gather<RB>(Hb, D, D, ts, XB_W, w8);
  // ---- hidden slice: z = W1 h2 + b1, u = dropout(gelu(z)) ----
  const int k0 = rank * f8;
  for (int cc = 0; cc < f8; cc += 64)
    xb_matvec<RB>(p.w1, D, k0 + cc, min(64, f8 - cc), D, Hb, D, zs + cc, XB_FW, p.b1);
  for (int t = threadIdx.x; t < RB * f8; t += XB_THREADS) {
    const int r = t / f8, k = t - r * f8;
    const float z = zs[r * XB_FW + k];
    const float u = xb_gelu(z) * xb_keep(p.drop_p, p.seed, 2, (int)(row0 + r), k0 + k);
    zs[r * XB_FW + k] = u;
    if (r < N) { p.z[(row0 + r) * F + k0 + k] = z; p.u[(row0 + r) * F + k0 + k] = u; }
  }
  cluster_sync_all();
  xb_gather<RB>(U, F, F, zs, XB_FW, f8);             // overwrites X / Hb: both dead
  // ---- out slice = x1 + dropout(W2 u + b2) ----
  xb_matvec<RB>(p.w2, F, c0, w8, F, U, F, ts, XB_W, p.b2);
  for (int t = threadIdx.x; t < RB * w8; t += XB_THREADS) {
    const int r = t / w8, j = t - r * w8;
    if (r < N)
      p.out[(row0 + r) * D + c0 + j] = x1s[r * XB_W + j] + ts[r * XB_W + j] * xb_keep(p.drop_p, p.seed, 3, (int)(row0 + r), c0 + j);
  }
  cluster_sync_all();
}

template <int RB>
__global__ void __cluster_dims__(XB_CL, 1, 1) __launch_bounds__(XB_THREADS, 1) xfblock_bwd_kernel(XbParams p) {
  extern __shared__ __align__(16) float xb_smem[];
  using L = XbSmem<RB>;
  const int D = p.D, F = p.F, N = p.N, w8 = D / XB_CL, f8 = F / XB_CL, Dh = D / p.H;
  float* BIG = xb_smem + L::big;          // gathered full rows, [RB][F] at most
  float* qs = xb_smem + L::slices; float* ks = qs + RB * XB_W; float* vs = ks + RB * XB_W; float* os = vs + RB * XB_W;
  float* x1s = os + RB * XB_W; float* ts = x1s + RB * XB_W;
  float* zs = xb_smem + L::zs; float* sp = xb_smem + L::sp; float* aa = xb_smem + L::aa;
  float* stat1 = xb_smem + L::st; float* stat2 = stat1 + RB;
  float* scratch = xb_smem + L::scratch;
  const int rank = (int)cluster_ctarank(), b = blockIdx.y, c0 = rank * w8, k0 = rank * f8, hh = c0 / Dh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t row0 = (size_t)b * N;

  // ---- d f2 = dout * mask3 (slice), gathered; d u = W2^T d f2 over this CTA's hidden slice; d z = d u * mask2 * gelu'(z) ----
  for (int t = threadIdx.x; t < RB * w8; t += XB_THREADS) {
    const int r = t / w8, j = t - r * w8;
    float g = 0.f;
    if (r < N) {
      g = p.dout[(row0 + r) * D + c0 + j];
      const float gf = g * xb_keep(p.drop_p, p.seed, 3, (int)(row0 + r), c0 + j);
      p.d_f2[(row0 + r) * D + c0 + j] = gf;
      ts[r * XB_W + j] = gf;
    } else {
      ts[r * XB_W + j] = 0.f;
    }
    x1s[r * XB_W + j] = g;                 // d x1 accumulates here: starts with the residual path of the second half
  }
  cluster_sync_all();
  xb_gather<RB>(BIG, D, D, ts, XB_W, w8);  // BIG[r][0..D) = d f2 rows
  for (int cc = 0; cc < f8; cc += 64)
    xb_matvec_t<RB>(p.w2, F, k0 + cc, min(64, f8 - cc), D, BIG, D, zs + cc, XB_FW, scratch, false);
  for (int t = threadIdx.x; t < RB * f8; t += XB_THREADS) {
    const int r = t / f8, k = t - r * f8;
    float dz = 0.f;
    if (r < N) {
      dz = zs[r * XB_FW + k] * xb_keep(p.drop_p, p.seed, 2, (int)(row0 + r), k0 + k) * xb_gelu_grad(p.z[(row0 + r) * F + k0 + k]);
      p.d_z[(row0 + r) * F + k0 + k] = dz;
    }
    zs[r * XB_FW + k] = dz;
  }
  cluster_sync_all();
  xb_gather<RB>(BIG, F, F, zs, XB_FW, f8);           // d z full rows
  // ---- d h2 slice = W1^T d z; LayerNorm 2 backward; d x1 += ----
  xb_matvec_t<RB>(p.w1, D, c0, w8, F, BIG, F, ts, XB_W, scratch, false);
  for (int r = warp; r < RB; r += XB_THREADS / 32) {
    float s1 = 0.f, s2 = 0.f;
    for (int j = lane; j < w8; j += 32) {
      const float dh = ts[r * XB_W + j];
      const float g = dh * p.ln2w[c0 + j];
      const float xh = r < N ? p.xhat2[(row0 + r) * D + c0 + j] : 0.f;
      s1 += g; s2 = fmaf(g, xh, s2);
      if (r < N) p.d_h2[(row0 + r) * D + c0 + j] = dh;
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { stat1[r] = s1; stat2[r] = s2; }
  }
  cluster_sync_all();
  for (int r = warp; r < RB; r += XB_THREADS / 32) {
    const float a1 = xb_cluster_sum(stat1, r) / (float)D, a2 = xb_cluster_sum(stat2, r) / (float)D;
    const float rs = r < N ? p.rstd2[row0 + r] : 0.f;
    for (int j = lane; j < w8; j += 32) {
      const float g = ts[r * XB_W + j] * p.ln2w[c0 + j];
      const float xh = r < N ? p.xhat2[(row0 + r) * D + c0 + j] : 0.f;
      x1s[r * XB_W + j] += rs * (g - a1 - xh * a2);
    }
  }
  __syncthreads();
  // ---- d ao = d x1 * mask1 (slice), gathered; d o slice = W_o^T d ao ----
  for (int t = threadIdx.x; t < RB * w8; t += XB_THREADS) {
    const int r = t / w8, j = t - r * w8;
    const float g = r < N ? x1s[r * XB_W + j] * xb_keep(p.drop_p, p.seed, 1, (int)(row0 + r), c0 + j) : 0.f;
    ts[r * XB_W + j] = g;
    if (r < N) p.d_ao[(row0 + r) * D + c0 + j] = g;
  }
  cluster_sync_all();                       // also: every CTA is done reading the peers' stat1 / stat2
  xb_gather<RB>(BIG, D, D, ts, XB_W, w8);
  xb_matvec_t<RB>(p.w_o, D, c0, w8, D, BIG, D, os, XB_W, scratch, false);       // os = d o slice (columns of head hh)
  // ---- attention backward of head hh: partial d a' over this slice, exchanged within the head ----
  for (int t = threadIdx.x; t < RB * w8; t += XB_THREADS) {
    const int r = t / w8, j = t - r * w8;
    const float* src = p.qkv + (row0 + (r < N ? r : 0)) * 3 * D + c0 + j;
    qs[r * XB_W + j] = r < N ? src[0] : 0.f;
    ks[r * XB_W + j] = r < N ? src[D] : 0.f;
    vs[r * XB_W + j] = r < N ? src[2 * D] : 0.f;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < RB * RB; t += XB_THREADS) {
    const int i = t / RB, j = t - i * RB;
    float s = 0.f;
    for (int c = 0; c < w8; ++c) s = fmaf(os[i * XB_W + c], vs[j * XB_W + c], s);
    sp[t] = s;
  }
  cluster_sync_all();
  {
    const int per = Dh / w8, first = hh * per;
    const float isq = rsqrtf((float)Dh);
    for (int i = warp; i < RB; i += XB_THREADS / 32) {
      float da = 0.f, pr = 0.f, kp = 0.f;
      if (lane < N && i < N) {
        const uint32_t a = smem_u32(sp + i * RB + lane);
        for (int c = 0; c < per; ++c) da += xb_ld_dsmem(a, first + c);
        pr = p.attn[(((size_t)b * p.H + hh) * N + i) * N + lane];
        kp = xb_keep(p.drop_p, p.seed, 0, (b * p.H + hh) * N + i, lane);
      }
      const float dpr = da * kp;                       // gradient w.r.t. the softmax probability
      const float dot = warp_sum(dpr * pr);
      if (lane < RB) {
        aa[i * RB + lane] = pr * kp;                    // a' (dropped probabilities): for d v
        zs[i * RB + lane] = pr * (dpr - dot) * isq;     // d s_ij with the 1 / sqrt(Dh) of the scores folded in
      }
    }
  }
  __syncthreads();
  // d v[j, c] = sum_i a'[i, j] d o[i, c];  d q[i, c] = sum_j ds[i, j] k[j, c];  d k[j, c] = sum_i ds[i, j] q[i, c]
  // (into scratch first: q / k / v slices are still being read)
  for (int t = threadIdx.x; t < RB * w8; t += XB_THREADS) {
    const int r = t / w8, c = t - r * w8;
    float dv = 0.f, dq = 0.f, dk = 0.f;
    if (r < N)
      for (int o = 0; o < N; ++o) {
        dv = fmaf(aa[o * RB + r], os[o * XB_W + c], dv);
        dq = fmaf(zs[r * RB + o], ks[o * XB_W + c], dq);
        dk = fmaf(zs[o * RB + r], qs[o * XB_W + c], dk);
      }
    scratch[t] = dq; scratch[RB * XB_W + t] = dk; scratch[2 * RB * XB_W + t] = dv;
    if (r < N) {
      float* dst = p.d_qkv + (row0 + r) * 3 * D + c0 + c;
      dst[0] = dq; dst[D] = dk; dst[2 * D] = dv;
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < RB * w8; t += XB_THREADS) {
    const int r = t / w8, c = t - r * w8;
    qs[r * XB_W + c] = scratch[t]; ks[r * XB_W + c] = scratch[RB * XB_W + t]; vs[r * XB_W + c] = scratch[2 * RB * XB_W + t];
  }
  cluster_sync_all();
  // ---- d h1 slice = W_q^T dq + W_k^T dk + W_v^T dv (full rows of each gathered in turn) ----
  xb_gather<RB>(BIG, D, D, qs, XB_W, w8);
  xb_matvec_t<RB>(p.w_in, D, c0, w8, D, BIG, D, ts, XB_W, scratch, false);
  xb_gather<RB>(BIG, D, D, ks, XB_W, w8);
  xb_matvec_t<RB>(p.w_in + (size_t)D * D, D, c0, w8, D, BIG, D, ts, XB_W, scratch, true);
  xb_gather<RB>(BIG, D, D, vs, XB_W, w8);
  xb_matvec_t<RB>(p.w_in + (size_t)2 * D * D, D, c0, w8, D, BIG, D, ts, XB_W, scratch, true);
  // ---- LayerNorm 1 backward + residual ----
  for (int r = warp; r < RB; r += XB_THREADS / 32) {
    float s1 = 0.f, s2 = 0.f;
    for (int j = lane; j < w8; j += 32) {
      const float dh = ts[r * XB_W + j];
      const float g = dh * p.ln1w[c0 + j];
      const float xh = r < N ? p.xhat1[(row0 + r) * D + c0 + j] : 0.f;
      s1 += g; s2 = fmaf(g, xh, s2);
      if (r < N) p.d_h1[(row0 + r) * D + c0 + j] = dh;
    }
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { stat1[r] = s1; stat2[r] = s2; }
  }
  cluster_sync_all();
  for (int r = warp; r < RB; r += XB_THREADS / 32) {
    const float a1 = xb_cluster_sum(stat1, r) / (float)D, a2 = xb_cluster_sum(stat2, r) / (float)D;
    const float rs = r < N ? p.rstd1[row0 + r] : 0.f;
    for (int j = lane; j < w8; j += 32) {
      const float g = ts[r * XB_W + j] * p.ln1w[c0 + j];
      const float xh = r < N ? p.xhat1[(row0 + r) * D + c0 + j] : 0.f;
      if (r < N) p.dx[(row0 + r) * D + c0 + j] = x1s[r * XB_W + j] + rs * (g - a1 - xh * a2);
    }
  }
  cluster_sync_all();
}

// dW[j, i] = sum_r a[r * lda + j] b[r * ldb + i] (j < J, i < I), db[j] = sum_r a[r, j]; optional LayerNorm sums
// dg[i] = sum_r a2[r, i] xh[r, i], dbeta[i] = sum_r a2[r, i] by the CTAs of blockIdx.y == 1. One CTA per 8 weight rows.
struct WgParams {
  const float* a; long long lda; const float* bm; long long ldb; float* dw; float* db; int J, I, R;
  const float* a2; const float* xh; float* dg; float* dbeta; int D2;
};
__global__ void __launch_bounds__(WG_THREADS) xfblock_wgrad_kernel(WgParams p) {
  __shared__ float coef[8][128];
  if (blockIdx.y == 1) {
    if (!p.dg || blockIdx.z != 0) return;
    const int i = blockIdx.x * WG_THREADS + threadIdx.x;
    if (i < p.D2) {
      float g = 0.f, bb = 0.f;
      for (int r = 0; r < p.R; ++r) { const float v = p.a2[(size_t)r * p.D2 + i]; g = fmaf(v, p.xh[(size_t)r * p.D2 + i], g); bb += v; }
      p.dg[i] = g; p.dbeta[i] = bb;
    }
    return;
  }
  const int j0 = blockIdx.x * 8;
  if (j0 >= p.J) return;
  constexpr int MAXE = 4;                 // 1024 columns per CTA (blockIdx.z selects the column block)
  const int ibase = blockIdx.z * MAXE * WG_THREADS;
  float acc[8][MAXE];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int e = 0; e < MAXE; ++e) acc[r][e] = 0.f;
  float bsum = 0.f;
  for (int rc = 0; rc < p.R; rc += 128) {
    const int nr = min(128, p.R - rc);
    __syncthreads();
    for (int t = threadIdx.x; t < 8 * nr; t += WG_THREADS) {
      const int jj = t / nr, r = rc + t - jj * nr;
      coef[jj][r - rc] = j0 + jj < p.J ? p.a[(size_t)r * p.lda + j0 + jj] : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int r = 0; r < nr; ++r) {
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        const int i = ibase + threadIdx.x + e * WG_THREADS;
        if (i < p.I) {
          const float v = p.bm[(size_t)(rc + r) * p.ldb + i];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) acc[jj][e] = fmaf(coef[jj][r], v, acc[jj][e]);
        }
      }
    }
    if (threadIdx.x < 8)
      for (int r = 0; r < nr; ++r) bsum += coef[threadIdx.x][r];
  }
#pragma unroll
  for (int e = 0; e < MAXE; ++e) {
    const int i = ibase + threadIdx.x + e * WG_THREADS;
    if (i < p.I)
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
        if (j0 + jj < p.J) p.dw[(size_t)(j0 + jj) * p.I + i] = acc[jj][e];
  }
  if (blockIdx.z == 0 && threadIdx.x < 8 && j0 + threadIdx.x < p.J && p.db) p.db[j0 + threadIdx.x] = bsum;
}


// acts0[b, n, :] = x[b, n, :] (+ pos[n, :]): the positional add in front of the first block
__global__ void agg_addpos_kernel(const float* __restrict__ x, long long sb, long long sn, const float* __restrict__ pos,
                                  float* __restrict__ out, int B, int N, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * N * D) return;
  const int d = (int)(i % D), n = (int)((i / D) % N);
  const long long b = i / ((long long)N * D);
  out[i] = x[b * sb + n * sn + d] + (pos ? pos[(long long)n * D + d] : 0.f);
}

// gpos[n, :] = sum_b dx[b, n, :] for n < N, 0 for the unused rows of the positional table
__global__ void agg_posgrad_kernel(const float* __restrict__ dx, float* __restrict__ gpos, int B, int N, int D, int rows) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * D) return;
  float t = 0.f;
  if (i < (long long)N * D)
    for (int b = 0; b < B; ++b) t += dx[(long long)b * N * D + i];
  gpos[i] = t;
}

}  // namespace b2

namespace b2host {
using namespace b2;

bool xfblock_ok(int N, int D, int H, int F) {
  return N >= 1 && N <= 16 && D % 128 == 0 && D <= 512 && F % 512 == 0 && F <= 2048 && H >= 1 && H <= 8 && 8 % H == 0 &&
         (D / H) % (D / 8) == 0;
}

template <int RB>
static int xb_launch(bool bwd, const XbParams& p, cudaStream_t s) {
  const size_t smem = (size_t)XbSmem<RB>::total * sizeof(float);
  auto k = bwd ? xfblock_bwd_kernel<RB> : xfblock_fwd_kernel<RB>;
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return B2_ECUDA;
  k<<<dim3(XB_CL, p.B), XB_THREADS, smem, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int xfblock_run(int backward, const XbParams& p, cudaStream_t s) {
  if (!xfblock_ok(p.N, p.D, p.H, p.F) || p.B <= 0) return B2_EINVAL;
  if (p.N <= 4) return xb_launch<4>(backward != 0, p, s);
  if (p.N <= 8) return xb_launch<8>(backward != 0, p, s);
  return xb_launch<16>(backward != 0, p, s);
}

// ptrs: host array of 35 device pointers, order = the fields of XbParams (see include/b200clip.h)
int xfblock(int backward, const void* const* q, int B, int N, int D, int H, int F, float eps1, float eps2, float drop_p,
            unsigned long long seed, long long mask_sb, cudaStream_t s) {
  if (!q) return B2_EINVAL;
  auto f = [&](int i) { return const_cast<float*>(reinterpret_cast<const float*>(q[i])); };
  for (int i : {0, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26})
    if (!q[i]) return B2_EINVAL;
  if (backward) {
    for (int i = 27; i < 35; ++i)
      if (!q[i]) return B2_EINVAL;
  } else if (!q[1]) {
    return B2_EINVAL;
  }
  XbParams p{f(0), f(1), reinterpret_cast<const unsigned char*>(q[2]), mask_sb, f(3), f(4), f(5), f(6), f(7), f(8), f(9), f(10),
             f(11), f(12), f(13), f(14), eps1, eps2, f(15), f(16), f(17), f(18), f(19), f(20), f(21), f(22), f(23), f(24), f(25),
             f(26), f(27), f(28), f(29), f(30), f(31), f(32), f(33), f(34), B, N, D, H, F, drop_p, seed};
  return xfblock_run(backward, p, s);
}

int xfblock_wgrad(const float* a, long long lda, const float* bm, long long ldb, float* dw, float* db, int J, int I, int R,
                  const float* a2, const float* xh, float* dg, float* dbeta, int D2, cudaStream_t s) {
  if (!a || !bm || !dw || J <= 0 || I <= 0 || I > 2048 || R <= 0) return B2_EINVAL;
  WgParams p{a, lda, bm, ldb, dw, db, J, I, R, a2, xh, dg, dbeta, D2};
  const int gx = max((J + 7) / 8, dg ? (D2 + WG_THREADS - 1) / WG_THREADS : 1);
  xfblock_wgrad_kernel<<<dim3(gx, dg ? 2 : 1, (I + 4 * WG_THREADS - 1) / (4 * WG_THREADS)), WG_THREADS, 0, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}


// ---- the whole aggregator step as one host call (EnhancedVideoAggregator.forward with depth >= 1, reference
// models/video_aggregator.py:128-158): positional add, the blocks, final LayerNorm + query pool; and its backward --------
static long long al64(long long n) { return (n + 63) / 64 * 64; }

// sizes[0] = floats saved per block by the forward, [1] = floats of backward work space, [2] = parameter-gradient floats
// per block (packed in the order ln1.w ln1.b in_proj.w in_proj.b out_proj.w out_proj.b ln2.w ln2.b fc1.w fc1.b fc2.w fc2.b)
void aggregator_sizes(int B, int N, int D, int H, int F, long long* sizes) {
  const long long R = (long long)B * N;
  sizes[0] = 6 * al64(R * D) + 2 * al64(R) + al64(3 * R * D) + al64((long long)B * H * N * N) + 2 * al64(R * F);
  sizes[1] = 4 * al64(R * D) + al64(R * F) + al64(3 * R * D);
  sizes[2] = 4ll * D * D + 2ll * F * D + 9ll * D + F;
}

int aggregator(int backward, const void* const* q, int depth, int B, int N, int D, int H, int F, const float* eps,
               float drop_p, const long long* seeds, long long mask_sb, long long x_sb, long long x_sn, int pos_rows,
               cudaStream_t s) {
  if (!q || !eps || !seeds || depth < 1 || B < 1 || !xfblock_ok(N, D, H, F)) return B2_EINVAL;
  auto f = [&](int i) { return const_cast<float*>(reinterpret_cast<const float*>(q[i])); };
  for (int i : {3, 4, 6, 7, 8})
    if (!q[i]) return B2_EINVAL;
  for (int i = 15; i < 15 + 12 * depth; ++i)
    if (!q[i]) return B2_EINVAL;
  if (backward ? (!q[9] || !q[10] || !q[11] || !q[12] || !q[14] || (q[1] && !q[13])) : (!q[0] || !q[5])) return B2_EINVAL;
  const long long R = (long long)B * N, act = R * D;
  long long sz[3];
  aggregator_sizes(B, N, D, H, F, sz);
  const unsigned char* mask = reinterpret_cast<const unsigned char*>(q[2]);
  auto block = [&](int i, XbParams& p) {
    const void* const* w = q + 15 + 12 * i;
    auto g = [&](int j) { return reinterpret_cast<const float*>(w[j]); };
    float* c = f(4) + (long long)i * sz[0];
    p = XbParams{};
    p.x = f(3) + (long long)i * act; p.out = f(3) + (long long)(i + 1) * act; p.mask = mask; p.mb = mask_sb;
    p.ln1w = g(0); p.ln1b = g(1); p.w_in = g(2); p.b_in = g(3); p.w_o = g(4); p.b_o = g(5);
    p.ln2w = g(6); p.ln2b = g(7); p.w1 = g(8); p.b1 = g(9); p.w2 = g(10); p.b2 = g(11);
    p.eps1 = eps[2 * i]; p.eps2 = eps[2 * i + 1];
    p.xhat1 = c; c += al64(R * D); p.rstd1 = c; c += al64(R); p.h1 = c; c += al64(R * D); p.qkv = c; c += al64(3 * R * D);
    p.attn = c; c += al64((long long)B * H * N * N); p.o = c; c += al64(R * D); p.x1 = c; c += al64(R * D);
    p.xhat2 = c; c += al64(R * D); p.rstd2 = c; c += al64(R); p.h2 = c; c += al64(R * D); p.z = c; c += al64(R * F); p.u = c;
    p.B = B; p.N = N; p.D = D; p.H = H; p.F = F; p.drop_p = drop_p; p.seed = (unsigned long long)seeds[i];
  };
  XbParams p;
  if (!backward) {
    agg_addpos_kernel<<<(unsigned)((act + 255) / 256), 256, 0, s>>>(f(0), x_sb, x_sn, f(1), f(3), B, N, D);
    for (int i = 0; i < depth; ++i) {
      block(i, p);
      if (int rc = xfblock_run(0, p, s)) return rc;
    }
    return querypool(0, f(3) + (long long)depth * act, (long long)N * D, D, nullptr, f(6), f(7), f(8), mask, mask_sb, B, N, D,
                     eps[2 * depth], f(5), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, s);
  }
  float* dact = f(10);
  float* tail = f(12);
  if (cudaMemsetAsync(tail, 0, 3ull * D * sizeof(float), s) != cudaSuccess) return B2_ECUDA;
  if (int rc = querypool(1, f(3) + (long long)depth * act, (long long)N * D, D, nullptr, f(6), f(7), f(8), mask, mask_sb, B, N, D,
                         eps[2 * depth], nullptr, f(9), dact + (depth & 1) * act, nullptr, tail, tail + D, tail + 2 * D, s))
    return rc;
  for (int i = depth - 1; i >= 0; --i) {
    block(i, p);
    float* c = f(11);
    p.dout = dact + ((i + 1) & 1) * act; p.dx = dact + (i & 1) * act;
    p.d_f2 = c; c += al64(R * D); p.d_z = c; c += al64(R * F); p.d_ao = c; c += al64(R * D); p.d_qkv = c; c += al64(3 * R * D);
    p.d_h2 = c; c += al64(R * D); p.d_h1 = c;
    if (int rc = xfblock_run(1, p, s)) return rc;
    float* g = f(14) + (long long)i * sz[2];
    float *ln1w = g, *ln1b = ln1w + D, *w_in = ln1b + D, *b_in = w_in + 3ll * D * D, *w_o = b_in + 3 * D, *b_o = w_o + (long long)D * D,
          *ln2w = b_o + D, *ln2b = ln2w + D, *w1 = ln2b + D, *b1 = w1 + (long long)F * D, *w2 = b1 + F, *b2 = w2 + (long long)D * F;
    int rc = xfblock_wgrad(p.d_f2, D, p.u, F, w2, b2, D, F, (int)R, nullptr, nullptr, nullptr, nullptr, 0, s);
    if (!rc) rc = xfblock_wgrad(p.d_z, F, p.h2, D, w1, b1, F, D, (int)R, p.d_h2, p.xhat2, ln2w, ln2b, D, s);
    if (!rc) rc = xfblock_wgrad(p.d_ao, D, p.o, D, w_o, b_o, D, D, (int)R, nullptr, nullptr, nullptr, nullptr, 0, s);
    if (!rc) rc = xfblock_wgrad(p.d_qkv, 3 * D, p.h1, D, w_in, b_in, 3 * D, D, (int)R, p.d_h1, p.xhat1, ln1w, ln1b, D, s);
    if (rc) return rc;
  }
  if (q[13])
    agg_posgrad_kernel<<<(unsigned)(((long long)pos_rows * D + 255) / 256), 256, 0, s>>>(dact, f(13), B, N, D, pos_rows);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
