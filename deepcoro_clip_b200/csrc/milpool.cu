// Gated-attention pooling of the multi-instance probing head (reference models/multi_instance_linear_probing.py:493-507 for
// [B, N, D] inputs and :509-536 for the two-level [B, N, L, D] case), forward and backward, fp32:
//     a_l = w . (tanh(V x_l + bV) * sigmoid(U x_l + bU)) + bw ;  A = softmax_l(a masked) ;  out = sum_l drop(A_l) x_l
// over S sequences of L instances (S = B, L = N views; or S = B N, L = patch tokens for the first level of the two-level
// case, then S = B, L = N on its result with the same weights).
//
// The gate logits are a [R x D] x [D x 2 Hd] product (R = S L rows) followed by an elementwise gate and a row reduction;
// its backward is two more products of the same size. fp32 parity with the reference rules out single-pass bf16 / tf32
// tensor-core products, so these are register-blocked fp32 FMA tiles (128 x 128 x 16, 8 x 8 per thread, double-buffered
// shared memory) with the gate applied in the epilogue of the first:
//   mil_gate_fwd   : tile of 128 rows x 64 hidden units (V and U columns of a unit in the same thread); writes t = tanh(.),
//                    g = sigmoid(.) [R, 2 Hd] for the backward and the partial logits per unit tile.
//   mil_pool_fwd   : per sequence (x P parts along L): logits -> masked softmax -> (dropout) -> weighted row sum.
//   mil_dA / mil_ds: dA_l = dout . x_l ; ds = A (dA - sum A dA).
//   mil_dpre       : dpre = ds w (g (1 - t^2) | t g (1 - g)) [R, 2 Hd] written once + partial bias / w gradient sums.
//   mil_dx         : dx = drop(A) dout + dpre [V; U].
//   mil_dw         : d[V; U] = dpre^T x, rows split over grid.z into partial products;
//   mil_reduce     : fixed-order sums of the partials (deterministic).
#include "tile_engine.cuh"
#include "host_api.h"

namespace b2 {

constexpr int MG_BM = 128, MG_BK = 16, MG_LD = 132, MG_THREADS = 256, MG_FCH = 64;

struct MilParams {
  const float* x;
  long long sx_seq, sx_tok, R;
  int S, L, D, Hd;
  const float *V, *bV, *U, *bU, *w, *bw;
  const uint8_t* mask;
  long long smask;
  float drop_p;
  unsigned long long seed;
  float *tg, *spart, *attn, *opart, *out;
  int P, nut;
  const float* dout;
  float *ds, *dx, *dpre, *wpart, *fpart, *dW, *dsmall;
  int Z, chunks;
};

__device__ __forceinline__ const float* mil_xrow(const MilParams& p, long long r) {
  return p.x + (r / p.L) * p.sx_seq + (r % p.L) * p.sx_tok;
}
__device__ __forceinline__ float4 mg_ld4(const float* q) { return __ldg(reinterpret_cast<const float4*>(q)); }
__device__ __forceinline__ int mg_row(int t, int i) { return i < 4 ? t * 4 + i : 64 + t * 4 + i - 4; }
__device__ __forceinline__ float mil_keepscale(const MilParams& p, long long r) {
  if (p.drop_p <= 0.f) return 1.f;
  return attn_keep(p.seed, (int)(r / p.L), (int)(r % p.L), p.drop_p) ? 1.f / (1.f - p.drop_p) : 0.f;
}

// acc[i][j] += sum_k A[k][row(i)] B[k][col(j)] over one 16-deep slab; rows / columns of a thread are two groups of four,
// 64 apart, so a quarter-warp reads consecutive float4 (no bank conflicts) and the others broadcast.
__device__ __forceinline__ void mg_fma_tile(const float* As, const float* Bs, float (&acc)[8][8], int ty, int tx) {
#pragma unroll
  for (int k = 0; k < MG_BK; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(As + k * MG_LD + ty * 4);
    const float4 a1 = *reinterpret_cast<const float4*>(As + k * MG_LD + 64 + ty * 4);
    const float4 b0 = *reinterpret_cast<const float4*>(Bs + k * MG_LD + tx * 4);
    const float4 b1 = *reinterpret_cast<const float4*>(Bs + k * MG_LD + 64 + tx * 4);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// fetch(kt): global -> registers for slab kt; commit(As, Bs): registers -> shared memory. One __syncthreads per slab.
template <class Fetch, class Commit>
__device__ __forceinline__ void mg_mainloop(int nk, Fetch fetch, Commit commit, float* sA, float* sB, float (&acc)[8][8]) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  if (nk <= 0) return;
  fetch(0);
  commit(sA, sB);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) fetch(kt + 1);
    mg_fma_tile(sA + cur * MG_BK * MG_LD, sB + cur * MG_BK * MG_LD, acc, ty, tx);
    if (kt + 1 < nk) commit(sA + (cur ^ 1) * MG_BK * MG_LD, sB + (cur ^ 1) * MG_BK * MG_LD);
    __syncthreads();
  }
}

__device__ __forceinline__ void mg_store_t(float* S, int k4, int col, const float4& v) {   // transposed: 4 slab rows, 1 col
  S[(k4 + 0) * MG_LD + col] = v.x;
  S[(k4 + 1) * MG_LD + col] = v.y;
  S[(k4 + 2) * MG_LD + col] = v.z;
  S[(k4 + 3) * MG_LD + col] = v.w;
}

// ---- gate logits: 128 rows x 64 hidden units per CTA -------------------------------------------------------------------
__global__ void __launch_bounds__(MG_THREADS) mil_gate_fwd_kernel(MilParams p) {
  __shared__ __align__(16) float sA[2 * MG_BK * MG_LD], sB[2 * MG_BK * MG_LD];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const long long r0 = (long long)blockIdx.x * MG_BM;
  const int u0 = blockIdx.y * 64;
  const int lrow = tid >> 2, kq = tid & 3;
  const float *ap[2], *bp[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const long long r = r0 + lrow + 64 * h;
    ap[h] = r < p.R ? mil_xrow(p, r) + kq * 4 : nullptr;
    const int u = u0 + lrow;
    bp[h] = u < p.Hd ? (h == 0 ? p.V : p.U) + (long long)u * p.D + kq * 4 : nullptr;
  }
  float4 ra[2], rb[2];
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  auto fetch = [&](int kt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      ra[h] = ap[h] ? mg_ld4(ap[h] + kt * MG_BK) : zero;
      rb[h] = bp[h] ? mg_ld4(bp[h] + kt * MG_BK) : zero;
    }
  };
  auto commit = [&](float* As, float* Bs) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mg_store_t(As, kq * 4, lrow + 64 * h, ra[h]);
      mg_store_t(Bs, kq * 4, lrow + 64 * h, rb[h]);
    }
  };
  mg_mainloop(p.D / MG_BK, fetch, commit, sA, sB, acc);

  const int ub = u0 + tx * 4;
  const bool uok = ub < p.Hd;
  float bv[4], bu[4], ww[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    bv[e] = uok ? p.bV[ub + e] : 0.f;
    bu[e] = uok ? p.bU[ub + e] : 0.f;
    ww[e] = uok ? p.w[ub + e] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long r = r0 + mg_row(ty, i);
    float t[4], g[4], part = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      t[e] = tanhf(acc[i][e] + bv[e]);
      g[e] = 1.f / (1.f + expf(-(acc[i][4 + e] + bu[e])));
      part = fmaf(ww[e], t[e] * g[e], part);
    }
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (r < p.R) {
      if (uok) {
        float* tgr = p.tg + r * (2ll * p.Hd);
        *reinterpret_cast<float4*>(tgr + ub) = make_float4(t[0], t[1], t[2], t[3]);
        *reinterpret_cast<float4*>(tgr + p.Hd + ub) = make_float4(g[0], g[1], g[2], g[3]);
      }
      if (tx == 0) p.spart[(long long)blockIdx.y * p.R + r] = part;
    }
  }
}

__device__ __forceinline__ float mil_block_reduce(float v, bool is_max, float* sred) {
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = sred[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = is_max ? fmaxf(r, sred[i]) : r + sred[i];
  return r;
}

// ---- masked softmax over the sequence + weighted sum of its rows (part blockIdx.x of P along L) -------------------------
__global__ void __launch_bounds__(256) mil_pool_fwd_kernel(MilParams p) {
  extern __shared__ float sc[];   // [L]
  __shared__ float4 red[256];
  __shared__ float sred[8];
  const int s = blockIdx.y, part = blockIdx.x, tid = threadIdx.x, L = p.L;
  const long long rb = (long long)s * L;
  const float bw = p.bw[0];
  float lmax = -INFINITY;
  for (int l = tid; l < L; l += 256) {
    float v = bw;
    for (int ut = 0; ut < p.nut; ++ut) v += p.spart[(long long)ut * p.R + rb + l];
    if (p.mask && p.mask[(long long)s * p.smask + l] == 0) v = -INFINITY;
    sc[l] = v;
    lmax = fmaxf(lmax, v);
  }
  const float m = mil_block_reduce(lmax, true, sred);
  float lsum = 0.f;
  for (int l = tid; l < L; l += 256) {
    const float e = expf(sc[l] - m);      // every instance masked: -inf - -inf = NaN, like the reference's softmax
    sc[l] = e;
    lsum += e;
  }
  const float inv = 1.f / mil_block_reduce(lsum, false, sred);
  const int lp = (L + p.P - 1) / p.P, l0 = part * lp, l1 = min(L, l0 + lp);
  for (int l = l0 + tid; l < l1; l += 256) {
    const float a = sc[l] * inv;
    p.attn[rb + l] = a;
    sc[l] = a * mil_keepscale(p, rb + l);
  }
  __syncthreads();
  const int nq = p.D >> 2;
  float* dst = p.P == 1 ? p.out + (long long)s * p.D : p.opart + ((long long)s * p.P + part) * p.D;
  for (int q0 = 0; q0 < nq; q0 += 256) {
    const int nqc = min(256, nq - q0), ng = 256 / nqc, g = tid / nqc, qi = tid - g * nqc;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g < ng) {
      const float* xb = p.x + (long long)s * p.sx_seq + (q0 + qi) * 4;
      int l = l0 + g;
      for (; l + 3 * ng < l1; l += 4 * ng) {
        float4 v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = mg_ld4(xb + (long long)(l + e * ng) * p.sx_tok);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = sc[l + e * ng];
          acc.x = fmaf(a, v[e].x, acc.x);
          acc.y = fmaf(a, v[e].y, acc.y);
          acc.z = fmaf(a, v[e].z, acc.z);
          acc.w = fmaf(a, v[e].w, acc.w);
        }
      }
      for (; l < l1; l += ng) {
        const float4 v = mg_ld4(xb + (long long)l * p.sx_tok);
        const float a = sc[l];
        acc.x = fmaf(a, v.x, acc.x);
        acc.y = fmaf(a, v.y, acc.y);
        acc.z = fmaf(a, v.z, acc.z);
        acc.w = fmaf(a, v.w, acc.w);
      }
    }
    red[tid] = acc;
    __syncthreads();
    if (tid < nqc) {
      float4 t = red[tid];
      for (int gg = 1; gg < ng; ++gg) {
        const float4 o = red[gg * nqc + tid];
        t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
      }
      *reinterpret_cast<float4*>(dst + (q0 + tid) * 4) = t;
    }
    __syncthreads();
  }
}

__global__ void mil_pool_combine_kernel(MilParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)p.S * p.D) return;
  const long long s = i / p.D, d = i % p.D;
  float t = 0.f;
  for (int q = 0; q < p.P; ++q) t += p.opart[(s * p.P + q) * p.D + d];
  p.out[i] = t;
}

// ---- backward of the weighted sum: dA_l = keepscale_l (dout . x_l) (one warp per row) -----------------------------------
__global__ void __launch_bounds__(256) mil_dA_kernel(MilParams p) {
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= p.R) return;
  const int lane = threadIdx.x & 31;
  const float* xr = mil_xrow(p, r);
  const float* dr = p.dout + (r / p.L) * p.D;
  float a = 0.f;
  for (int d = lane * 4; d < p.D; d += 128) {
    const float4 x = mg_ld4(xr + d), g = mg_ld4(dr + d);
    a = fmaf(x.x, g.x, fmaf(x.y, g.y, fmaf(x.z, g.z, fmaf(x.w, g.w, a))));
  }
  a = warp_sum(a);
  if (lane == 0) p.ds[r] = a * mil_keepscale(p, r);
}

// softmax backward per sequence, in place: ds = A (dA - sum_j A_j dA_j)
__global__ void __launch_bounds__(256) mil_ds_kernel(MilParams p) {
  __shared__ float sred[8];
  const long long rb = (long long)blockIdx.x * p.L;
  float a = 0.f;
  for (int l = threadIdx.x; l < p.L; l += 256) a = fmaf(p.attn[rb + l], p.ds[rb + l], a);
  const float dot = mil_block_reduce(a, false, sred);
  for (int l = threadIdx.x; l < p.L; l += 256) p.ds[rb + l] = p.attn[rb + l] * (p.ds[rb + l] - dot);
}

// ---- dx = drop(A) dout + dpre [V; U] : 128 rows x 128 columns per CTA, K = 2 Hd ------------------------------------------
__global__ void __launch_bounds__(MG_THREADS, 2) mil_dx_kernel(MilParams p) {
  __shared__ __align__(16) float sA[2 * MG_BK * MG_LD], sB[2 * MG_BK * MG_LD];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const long long r0 = (long long)blockIdx.x * MG_BM;
  const int n0 = blockIdx.y * 128, Hd = p.Hd;
  const int lrow = tid >> 2, cq = tid & 3, crow = tid >> 5, nq = tid & 31;
  const float* dpr[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const long long r = r0 + lrow + 64 * h;
    dpr[h] = r < p.R ? p.dpre + r * (2ll * Hd) + cq * 4 : nullptr;
  }
  const bool nok = n0 + nq * 4 < p.D;
  float4 ra[2], rb[2];
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  auto fetch = [&](int kt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      ra[h] = dpr[h] ? mg_ld4(dpr[h] + kt * MG_BK) : zero;
      const int c = kt * MG_BK + crow + 8 * h;
      const float* wr = (c < Hd ? p.V + (long long)c * p.D : p.U + (long long)(c - Hd) * p.D) + n0 + nq * 4;
      rb[h] = nok ? mg_ld4(wr) : zero;
    }
  };
  auto commit = [&](float* As, float* Bs) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mg_store_t(As, cq * 4, lrow + 64 * h, ra[h]);
      *reinterpret_cast<float4*>(Bs + (crow + 8 * h) * MG_LD + nq * 4) = rb[h];
    }
  };
  mg_mainloop(2 * Hd / MG_BK, fetch, commit, sA, sB, acc);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long r = r0 + mg_row(ty, i);
    if (r >= p.R) continue;
    const float ad = p.attn[r] * mil_keepscale(p, r);
    const float* dr = p.dout + (r / p.L) * p.D;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + 64 * h + tx * 4;
      if (n < p.D) {
        const float4 g = mg_ld4(dr + n);
        *reinterpret_cast<float4*>(p.dx + r * p.D + n) =
            make_float4(fmaf(ad, g.x, acc[i][4 * h + 0]), fmaf(ad, g.y, acc[i][4 * h + 1]), fmaf(ad, g.z, acc[i][4 * h + 2]),
                        fmaf(ad, g.w, acc[i][4 * h + 3]));
      }
    }
  }
}

// ---- d[V; U] partial products over the row range of blockIdx.z : 128 gate columns x 128 input columns --------------------
__global__ void __launch_bounds__(MG_THREADS, 2) mil_dw_kernel(MilParams p) {
  __shared__ __align__(16) float sA[2 * MG_BK * MG_LD], sB[2 * MG_BK * MG_LD];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int c0 = blockIdx.x * 128, n0 = blockIdx.y * 128, Hd = p.Hd;
  long long per = (p.R + p.Z - 1) / p.Z;
  per = (per + MG_BK - 1) / MG_BK * MG_BK;
  const long long rbeg = (long long)blockIdx.z * per, rend = min(p.R, rbeg + per);
  const int krow = tid >> 5, q = tid & 31;
  const bool cok = c0 + q * 4 < 2 * Hd, nok = n0 + q * 4 < p.D;
  float4 ra[2], rb[2];
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  // rows of slab kt are fetched in order, so (sequence, instance) of a thread's two rows advance by 16 per call instead of
  // a 64-bit division per load
  long long rs[2];
  int rl[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const long long r = rbeg + krow + 8 * h;
    rs[h] = r / p.L;
    rl[h] = (int)(r % p.L);
  }
  auto fetch = [&](int kt) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long r = rbeg + (long long)kt * MG_BK + krow + 8 * h;
      const bool rok = r < rend;
      ra[h] = rok && cok ? mg_ld4(p.dpre + r * (2ll * Hd) + c0 + q * 4) : zero;
      rb[h] = rok && nok ? mg_ld4(p.x + rs[h] * p.sx_seq + rl[h] * p.sx_tok + n0 + q * 4) : zero;
      rl[h] += MG_BK;
      while (rl[h] >= p.L) {
        rl[h] -= p.L;
        ++rs[h];
      }
    }
  };
  auto commit = [&](float* As, float* Bs) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *reinterpret_cast<float4*>(As + (krow + 8 * h) * MG_LD + q * 4) = ra[h];
      *reinterpret_cast<float4*>(Bs + (krow + 8 * h) * MG_LD + q * 4) = rb[h];
    }
  };
  const int nk = rend > rbeg ? (int)((rend - rbeg + MG_BK - 1) / MG_BK) : 0;
  mg_mainloop(nk, fetch, commit, sA, sB, acc);
  float* dst = p.wpart + (long long)blockIdx.z * (2ll * Hd) * p.D;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + mg_row(ty, i);
    if (c >= 2 * Hd) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + 64 * h + tx * 4;
      if (n < p.D)
        *reinterpret_cast<float4*>(dst + (long long)c * p.D + n) =
            make_float4(acc[i][4 * h + 0], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
    }
  }
}

// ---- dpre = ds w (g (1 - t^2) | t g (1 - g)) [R, 2 Hd], written once for the two products, and the partial sums of its
// columns over MG_FCH rows: fpart[chunk] = [dbV (Hd) | dbU (Hd) | dw (Hd) | dbw] (dw_u = sum_r ds_r t_u g_u, dbw = sum ds) ----
__global__ void __launch_bounds__(256) mil_dpre_kernel(MilParams p) {
  __shared__ float sd[MG_FCH];
  __shared__ float4 red[3][256];
  __shared__ float sred[8];
  const int tid = threadIdx.x, Hd = p.Hd;
  const long long rbeg = (long long)blockIdx.x * MG_FCH;
  const int nr = (int)min((long long)MG_FCH, p.R - rbeg);
  float mine = 0.f;
  if (tid < MG_FCH) {
    sd[tid] = tid < nr ? p.ds[rbeg + tid] : 0.f;
    mine = sd[tid];
  }
  const float tot = mil_block_reduce(mine, false, sred);     // also publishes sd
  float* dst = p.fpart + (long long)blockIdx.x * (3 * Hd + 4);     // row pitch 3 Hd + 4: float4 stores stay aligned
  if (tid == 0) dst[3 * Hd] = tot;
  const int nq = Hd >> 2;
  for (int q0 = 0; q0 < nq; q0 += 256) {
    const int nqc = min(256, nq - q0), ng = 256 / nqc, g = tid / nqc, qi = tid - g * nqc, u = (q0 + qi) * 4;
    float4 sv = make_float4(0.f, 0.f, 0.f, 0.f), su = sv, sw = sv;
    if (g < ng) {
      const float4 ww = mg_ld4(p.w + u);
      for (int i = g; i < nr; i += ng) {
        const long long off = (rbeg + i) * (2ll * Hd) + u;
        const float4 t = mg_ld4(p.tg + off), gg = mg_ld4(p.tg + off + Hd);
        const float d = sd[i];
        float4 pv, pu;
#define MIL_ONE(c)                                   \
  {                                                  \
    const float tgp = t.c * gg.c, b = d * ww.c;      \
    pv.c = b * gg.c * (1.f - t.c * t.c);             \
    pu.c = b * tgp * (1.f - gg.c);                   \
    sv.c += pv.c;                                    \
    su.c += pu.c;                                    \
    sw.c = fmaf(d, tgp, sw.c);                       \
  }
        MIL_ONE(x) MIL_ONE(y) MIL_ONE(z) MIL_ONE(w)
#undef MIL_ONE
        *reinterpret_cast<float4*>(p.dpre + off) = pv;
        *reinterpret_cast<float4*>(p.dpre + off + Hd) = pu;
      }
    }
    red[0][tid] = sv;
    red[1][tid] = su;
    red[2][tid] = sw;
    __syncthreads();
    if (tid < nqc) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float4 t = red[k][tid];
        for (int gg = 1; gg < ng; ++gg) {
          const float4 o = red[k][gg * nqc + tid];
          t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
        }
        *reinterpret_cast<float4*>(dst + k * Hd + (q0 + tid) * 4) = t;
      }
    }
    __syncthreads();
  }
}

// fixed-order sums of the partials: one thread per element of d[V; U] over the Z row ranges, one warp per bias / w element
// over the row chunks (lane-strided, then a shuffle tree)
__global__ void __launch_bounds__(256) mil_reduce_kernel(MilParams p) {
  const long long n1 = p.Z > 0 ? 2ll * p.Hd * p.D : 0, n2 = 3 * p.Hd + 1, nb1 = (n1 + 255) / 256;
  if (blockIdx.x < nb1) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i < n1) {
      float t = 0.f;
      for (int z = 0; z < p.Z; ++z) t += p.wpart[(long long)z * n1 + i];
      p.dW[i] = t;
    }
  } else {
    const long long j = (blockIdx.x - nb1) * 8 + (threadIdx.x >> 5);
    if (j < n2) {
      float t = 0.f;
      for (int c = threadIdx.x & 31; c < p.chunks; c += 32) t += p.fpart[(long long)c * (n2 + 3) + j];
      t = warp_sum(t);
      if ((threadIdx.x & 31) == 0) p.dsmall[j] = t;
    }
  }
}


// =====================================================================================================================
// Tensor-core variant (R >= 1024 rows, D in {256, 512, 768}; smaller problems keep fp32 products: they are launch-bound
// either way — 0.30 ms for [32, 4, 512] on both paths — and stay at fp32 accuracy, like precision="auto" of the losses): the three products on tcgen05 with split-precision operands.
// A value v is carried as hi = bf16(v), lo = bf16(v - hi) (16 mantissa bits); a product a b is formed as
// lo_a hi_b + hi_a lo_b + hi_a hi_b by ONE bf16 product over a three times longer K: A operand rows [lo | hi | hi], B operand
// rows [hi | lo | hi] (the bf16x3 scheme of the loss kernels, 2^-17 relative per product, fp32 accumulation in TMEM).
//   gate logits : tile engine (tile_engine.cuh) over A = x3 [R, 3 D], B = w3 [2 Hd, 3 D] with the gate columns interleaved
//                 (row 2u = V_u, 2u + 1 = U_u: a unit's pair sits in the same 32-column chunk of a thread); the epilogue
//                 applies tanh * sigmoid * w, writes t / g and the partial logit of its 128-column half.
//   dx          : same engine over A = dpre3 [R, 3 Hp], B = wt3 [D, 3 Hp] (Hp = 2 Hd rounded up to 64, zero padded); the
//                 epilogue adds drop(A_l) dout and writes dx.
//   d[V; U]     : gt_gemm.cu (dY += G^T X from [64 x 64] blocks of G, both operands MN-major) called three times:
//                 (dpre_hi, x_hi), (dpre_hi, x_lo), (dpre_lo, x_hi) — x_hi / x_lo are panels of x3.
// =====================================================================================================================
__device__ __forceinline__ void mil_split(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t mil_pack(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// x3[r, 0:D] = lo(x_r), x3[r, D:2D] = x3[r, 2D:3D] = hi(x_r)
__global__ void __launch_bounds__(256) mil_split_x_kernel(MilParams p, __nv_bfloat16* __restrict__ x3) {
  const int dq = p.D >> 2;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= p.R * dq) return;
  const long long r = i / dq;
  const int d = (int)(i - r * dq) * 4;
  const float4 v = mg_ld4(mil_xrow(p, r) + d);
  __nv_bfloat16 h[4], l[4];
  mil_split(v.x, h[0], l[0]); mil_split(v.y, h[1], l[1]); mil_split(v.z, h[2], l[2]); mil_split(v.w, h[3], l[3]);
  const uint2 hv = make_uint2(mil_pack(h[0], h[1]), mil_pack(h[2], h[3])), lv = make_uint2(mil_pack(l[0], l[1]), mil_pack(l[2], l[3]));
  __nv_bfloat16* row = x3 + r * (3ll * p.D) + d;
  *reinterpret_cast<uint2*>(row) = lv;
  *reinterpret_cast<uint2*>(row + p.D) = hv;
  *reinterpret_cast<uint2*>(row + 2 * p.D) = hv;
}

// w3 [2 Hd, 3 D] rows interleaved (2u = V_u, 2u + 1 = U_u): [hi | lo | hi];  wt3 [D, 3 Hp]: wt3[n, c] the same split of
// column n of row c, zero for c >= 2 Hd
__global__ void __launch_bounds__(256) mil_prep_w_kernel(MilParams p, int Hp, __nv_bfloat16* __restrict__ w3,
                                                         __nv_bfloat16* __restrict__ wt3) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)Hp * p.D) return;
  const int c = (int)(i / p.D), n = (int)(i - (long long)c * p.D);
  __nv_bfloat16 hi = __float2bfloat16_rn(0.f), lo = hi;
  if (c < 2 * p.Hd) {
    mil_split(((c & 1) ? p.U : p.V)[(long long)(c >> 1) * p.D + n], hi, lo);
    __nv_bfloat16* wr = w3 + (long long)c * (3ll * p.D) + n;
    wr[0] = hi;
    wr[p.D] = lo;
    wr[2 * p.D] = hi;
  }
  __nv_bfloat16* tr = wt3 + (long long)n * (3ll * Hp) + c;
  tr[0] = hi;
  tr[Hp] = lo;
  tr[2 * Hp] = hi;
}

struct MilGateEpiParams {
  const float *bV, *bU, *w;
  float *tg, *spart;
  int Hd;
  long long R;
};
struct MilGateEpi {
  using Params = MilGateEpiParams;
  struct State { float part; };
  __device__ static __forceinline__ void init(State& st, const Params&) { st.part = 0.f; }
  __device__ static __forceinline__ void begin_outer(State&, const Params&, int, const TeCtx&) {}
  __device__ static __forceinline__ void chunk(State& st, const Params& p, const TeCtx& ctx, int c, const uint32_t (&acc)[32]) {
    const int u0 = (ctx.col0 + c * 32) >> 1;        // 16 units per chunk: columns 2u (V) and 2u + 1 (U)
    if (!ctx.row_ok || u0 >= p.Hd) return;
    float* trow = p.tg + (long long)ctx.row * (2ll * p.Hd) + u0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (u0 + 4 * k >= p.Hd) break;                // Hd % 4 == 0: whole quads
      float t[4], g[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int u = u0 + 4 * k + e;
        t[e] = tanhf(__uint_as_float(acc[2 * (4 * k + e)]) + __ldg(p.bV + u));
        g[e] = 1.f / (1.f + expf(-(__uint_as_float(acc[2 * (4 * k + e) + 1]) + __ldg(p.bU + u))));
        st.part = fmaf(__ldg(p.w + u), t[e] * g[e], st.part);
      }
      *reinterpret_cast<float4*>(trow + 4 * k) = make_float4(t[0], t[1], t[2], t[3]);
      *reinterpret_cast<float4*>(trow + p.Hd + 4 * k) = make_float4(g[0], g[1], g[2], g[3]);
    }
  }
  __device__ static __forceinline__ void end_tile(State& st, const Params& p, const TeCtx& ctx) {
    if (ctx.row_ok) p.spart[(long long)(2 * ctx.n_block + ctx.wg) * p.R + ctx.row] = st.part;
    st.part = 0.f;
  }
  __device__ static __forceinline__ void end_outer(State&, const Params&, int, const TeCtx&) {}
};

struct MilDxEpiParams {
  const float *ad, *dout;      // drop(A_l) per row, upstream gradient [S, D]
  float* dx;                   // [R, D]
  int D, L;
};
struct MilDxEpi {
  using Params = MilDxEpiParams;
  struct State {};
  __device__ static __forceinline__ void init(State&, const Params&) {}
  __device__ static __forceinline__ void begin_outer(State&, const Params&, int, const TeCtx&) {}
  __device__ static __forceinline__ void chunk(State&, const Params& p, const TeCtx& ctx, int c, const uint32_t (&acc)[32]) {
    const int n0 = ctx.col0 + c * 32;
    if (!ctx.row_ok || n0 >= p.D) return;           // D % 32 == 0
    const float a = __ldg(p.ad + ctx.row);
    const float* dr = p.dout + (long long)(ctx.row / p.L) * p.D + n0;
    float* xr = p.dx + (long long)ctx.row * p.D + n0;
#pragma unroll
    for (int e = 0; e < 32; e += 4) {
      const float4 g = mg_ld4(dr + e);
      *reinterpret_cast<float4*>(xr + e) = make_float4(fmaf(a, g.x, __uint_as_float(acc[e])), fmaf(a, g.y, __uint_as_float(acc[e + 1])),
                                                      fmaf(a, g.z, __uint_as_float(acc[e + 2])), fmaf(a, g.w, __uint_as_float(acc[e + 3])));
    }
  }
  __device__ static __forceinline__ void end_tile(State&, const Params&, const TeCtx&) {}
  __device__ static __forceinline__ void end_outer(State&, const Params&, int, const TeCtx&) {}
};

// dpre in the three forms its consumers read, from t / g / ds (64 rows per CTA, a thread = 4 units = 8 interleaved columns):
//   dpre3 [R, 3 Hp] = [lo | hi | hi] (A operand of the dx product, zero in the padded columns),
//   ghi / glo: [64 x 64] blocks of gt_gemm.cu (8 KB each, 16-byte unit q of row r at position q ^ (r & 7)), zero rows past R,
//   ad[r] = drop(A_r); and the partial sums of the bias / w gradients as mil_dpre_kernel.
__global__ void __launch_bounds__(256) mil_dpre_tc_kernel(MilParams p, int Hp, int njb, __nv_bfloat16* __restrict__ dpre3,
                                                          __nv_bfloat16* __restrict__ ghi, __nv_bfloat16* __restrict__ glo,
                                                          float* __restrict__ ad) {
  __shared__ float sd[MG_FCH];
  __shared__ float4 red[3][256];
  __shared__ float sred[8];
  const int tid = threadIdx.x, Hd = p.Hd;
  const long long rbeg = (long long)blockIdx.x * MG_FCH;
  const int nr = (int)max(0ll, min((long long)MG_FCH, p.R - rbeg));
  float mine = 0.f;
  if (tid < MG_FCH) {
    sd[tid] = tid < nr ? p.ds[rbeg + tid] : 0.f;
    mine = sd[tid];
    if (tid < nr) ad[rbeg + tid] = p.attn[rbeg + tid] * mil_keepscale(p, rbeg + tid);
  }
  const float tot = mil_block_reduce(mine, false, sred);
  float* dst = p.fpart + (long long)blockIdx.x * (3 * Hd + 4);
  if (tid == 0) dst[3 * Hd] = tot;
  const int nq = Hp >> 3;                              // 8-column units per row (4 hidden units each)
  for (int q0 = 0; q0 < nq; q0 += 256) {
    const int nqc = min(256, nq - q0), ng = 256 / nqc, g = tid / nqc, qi = tid - g * nqc, u = (q0 + qi) * 4, c = 2 * u;
    float4 sv = make_float4(0.f, 0.f, 0.f, 0.f), su = sv, sw = sv;
    if (g < ng) {
      const bool uok = u < Hd;
      const float4 ww = uok ? mg_ld4(p.w + u) : sv;
      for (int i = g; i < MG_FCH; i += ng) {
        const long long r = rbeg + i;
        float pvv[4] = {0.f, 0.f, 0.f, 0.f}, puv[4] = {0.f, 0.f, 0.f, 0.f};
        if (i < nr && uok) {
          const long long off = r * (2ll * Hd) + u;
          const float4 t = mg_ld4(p.tg + off), gg = mg_ld4(p.tg + off + Hd);
          const float d = sd[i];
#define MIL_ONE(k, cmp)                              \
  {                                                  \
    const float tgp = t.cmp * gg.cmp, b = d * ww.cmp; \
    pvv[k] = b * gg.cmp * (1.f - t.cmp * t.cmp);     \
    puv[k] = b * tgp * (1.f - gg.cmp);               \
    sv.cmp += pvv[k];                                \
    su.cmp += puv[k];                                \
    sw.cmp = fmaf(d, tgp, sw.cmp);                   \
  }
          MIL_ONE(0, x) MIL_ONE(1, y) MIL_ONE(2, z) MIL_ONE(3, w)
#undef MIL_ONE
        }
        __nv_bfloat16 h[8], l[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          mil_split(pvv[k], h[2 * k], l[2 * k]);
          mil_split(puv[k], h[2 * k + 1], l[2 * k + 1]);
        }
        const uint4 hv = make_uint4(mil_pack(h[0], h[1]), mil_pack(h[2], h[3]), mil_pack(h[4], h[5]), mil_pack(h[6], h[7]));
        const uint4 lv = make_uint4(mil_pack(l[0], l[1]), mil_pack(l[2], l[3]), mil_pack(l[4], l[5]), mil_pack(l[6], l[7]));
        if (i < nr) {
          __nv_bfloat16* row = dpre3 + r * (3ll * Hp) + c;
          *reinterpret_cast<uint4*>(row) = lv;
          *reinterpret_cast<uint4*>(row + Hp) = hv;
          *reinterpret_cast<uint4*>(row + 2 * Hp) = hv;
        }
        // block (r / 64 = blockIdx.x, c / 64), row i, 16-byte unit (c % 64) / 8 swizzled with the row
        const long long boff = ((long long)blockIdx.x * njb + (c >> 6)) * 4096 + i * 64 + ((((c & 63) >> 3) ^ (i & 7)) << 3);
        *reinterpret_cast<uint4*>(ghi + boff) = hv;
        *reinterpret_cast<uint4*>(glo + boff) = lv;
      }
    }
    red[0][tid] = sv;
    red[1][tid] = su;
    red[2][tid] = sw;
    __syncthreads();
    if (tid < nqc && (q0 + tid) * 4 < Hd) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float4 t = red[k][tid];
        for (int gg = 1; gg < ng; ++gg) {
          const float4 o = red[k][gg * nqc + tid];
          t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
        }
        *reinterpret_cast<float4*>(dst + k * Hd + (q0 + tid) * 4) = t;
      }
    }
    __syncthreads();
  }
}

}  // namespace b2

namespace b2host {
using namespace b2;

bool milpool_ok(int L, int D, int Hd) { return L >= 1 && L <= 49152 && D >= 16 && D % 16 == 0 && Hd >= 8 && Hd % 8 == 0; }

// plan[0] = P (parts along L in the pooling pass), [1] = Z (row splits of the weight-gradient product),
// [2] = chunks (row chunks of the bias sums), [3] = unit tiles of the gate product
void milpool_plan(int S, int L, int D, int Hd, int* plan) {
  const long long R = (long long)S * L;
  int P = 1;
  if (L >= 256 && S < 296) P = (int)std::max(1ll, std::min<long long>(std::min(16, L / 128), (296 + S - 1) / S));
  const int tiles = ((2 * Hd + 127) / 128) * ((D + 127) / 128);
  // one full wave of the weight-gradient product: 2 CTAs per SM x 148 SMs = 296 tiles (a partial second wave costs a whole one)
  const int Z = (int)std::max(1ll, std::min<long long>((R + 255) / 256, std::max(1, 296 / tiles)));
  plan[0] = P;
  plan[1] = Z;
  plan[2] = (int)((R + MG_FCH - 1) / MG_FCH);
  plan[3] = (Hd + 63) / 64;
}

static int mil_fill(MilParams& p, const float* x, long long sx_seq, long long sx_tok, int S, int L, int D, int Hd,
                    const float* V, const float* U, const float* w, float drop_p, unsigned long long seed) {
  if (!milpool_ok(L, D, Hd) || S < 1 || !x || !V || !U || !w) return B2_EINVAL;
  if ((sx_seq | sx_tok) & 3 || ((uintptr_t)x | (uintptr_t)V | (uintptr_t)U | (uintptr_t)w) & 15) return B2_EINVAL;
  if (drop_p < 0.f || drop_p >= 1.f) return B2_EINVAL;
  p = MilParams{};
  p.x = x; p.sx_seq = sx_seq; p.sx_tok = sx_tok; p.S = S; p.L = L; p.D = D; p.Hd = Hd; p.R = (long long)S * L;
  p.V = V; p.U = U; p.w = w; p.drop_p = drop_p; p.seed = seed;
  int plan[4];
  milpool_plan(S, L, D, Hd, plan);
  p.P = plan[0]; p.Z = plan[1]; p.chunks = plan[2]; p.nut = plan[3];
  return B2_OK;
}

int milpool_fwd(const float* x, long long sx_seq, long long sx_tok, const uint8_t* mask, long long smask, const float* V,
                const float* bV, const float* U, const float* bU, const float* w, const float* bw, int S, int L, int D,
                int Hd, float drop_p, unsigned long long seed, float* tg, float* spart, float* attn, float* opart, float* out,
                cudaStream_t s) {
  MilParams p;
  if (int rc = mil_fill(p, x, sx_seq, sx_tok, S, L, D, Hd, V, U, w, drop_p, seed)) return rc;
  if (!bV || !bU || !bw || !tg || !spart || !attn || !out || (p.P > 1 && !opart)) return B2_EINVAL;
  p.bV = bV; p.bU = bU; p.bw = bw; p.mask = mask; p.smask = smask;
  p.tg = tg; p.spart = spart; p.attn = attn; p.opart = opart; p.out = out;
  mil_gate_fwd_kernel<<<dim3((unsigned)((p.R + MG_BM - 1) / MG_BM), p.nut), MG_THREADS, 0, s>>>(p);
  const size_t smem = (size_t)L * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(mil_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mil_pool_fwd_kernel<<<dim3(p.P, S), 256, smem, s>>>(p);
  if (p.P > 1) mil_pool_combine_kernel<<<(unsigned)(((long long)S * D + 255) / 256), 256, 0, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int milpool_bwd(const float* x, long long sx_seq, long long sx_tok, const float* V, const float* U, const float* w, int S,
                int L, int D, int Hd, float drop_p, unsigned long long seed, const float* tg, const float* attn,
                const float* dout, float* ds, float* dx, float* dpre, float* wpart, float* fpart, float* dW, float* dsmall,
                cudaStream_t s) {
  MilParams p;
  if (int rc = mil_fill(p, x, sx_seq, sx_tok, S, L, D, Hd, V, U, w, drop_p, seed)) return rc;
  if (!tg || !attn || !dout || !ds || !dx || !dpre || !wpart || !fpart || !dW || !dsmall) return B2_EINVAL;
  p.tg = const_cast<float*>(tg); p.attn = const_cast<float*>(attn); p.dout = dout;
  p.ds = ds; p.dx = dx; p.dpre = dpre; p.wpart = wpart; p.fpart = fpart; p.dW = dW; p.dsmall = dsmall;
  mil_dA_kernel<<<(unsigned)((p.R + 7) / 8), 256, 0, s>>>(p);
  mil_ds_kernel<<<S, 256, 0, s>>>(p);
  mil_dpre_kernel<<<p.chunks, 256, 0, s>>>(p);
  mil_dx_kernel<<<dim3((unsigned)((p.R + MG_BM - 1) / MG_BM), (D + 127) / 128), MG_THREADS, 0, s>>>(p);
  mil_dw_kernel<<<dim3((2 * Hd + 127) / 128, (D + 127) / 128, p.Z), MG_THREADS, 0, s>>>(p);
  const long long n1 = 2ll * Hd * D, n2 = 3 * Hd + 1;
  mil_reduce_kernel<<<(unsigned)((n1 + 255) / 256 + (n2 + 7) / 8), 256, 0, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}


// ---- tensor-core variant ---------------------------------------------------------------------------------------------
bool milpool_tc_ok(long long R, int L, int D, int Hd) {
  return milpool_ok(L, D, Hd) && (D == 256 || D == 512 || D == 768) && R >= 1024 && R < (1ll << 31) - 256 && Hd <= 4096 &&
         sm_count() >= 2;
}

// plan[0] = P, [1] = row chunks (64 rows each, rounded up to whole 128-row groups), [2] = partial-logit slots,
// [3] = Hp (gate columns padded to 64), [4] = bf16 elements of each blocked dpre buffer (gt_gemm.cu layout)
void milpool_tc_plan(int S, int L, int D, int Hd, long long* plan) {
  int base[4];
  milpool_plan(S, L, D, Hd, base);
  const long long R = (long long)S * L;
  plan[0] = base[0];
  plan[1] = 2 * ((R + 127) / 128);
  plan[2] = 2 * ((2 * Hd + TE_BN - 1) / TE_BN);
  plan[3] = (2 * Hd + 63) / 64 * 64;
  plan[4] = gstore_elems((int)R, 2 * Hd);
}

template <class Epi>
static int mil_launch_te(const void* A, const void* B, int Ma, int Nb, int Kp, const typename Epi::Params& ep, cudaStream_t s) {
  TeShape g{};
  g.Ma = Ma; g.Nb = Nb; g.Kp = Kp;
  g.m_tiles = (Ma + TE_BM - 1) / TE_BM;
  g.n_blocks = (Nb + TE_BN - 1) / TE_BN;
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmA, A, Ma, Kp, Kp, TE_BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmB, B, Nb, Kp, Kp, TE_BN))) return rc;
  auto kern = te_kernel<Epi, true>;
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device() & 63];
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TE_SMEM_BYTES) != cudaSuccess) return B2_ECUDA;
    attr_done = true;
  }
  const long long total = (long long)g.m_tiles * g.n_blocks;
  int grid = sm_count();
  if (total < grid) grid = (int)total;
  kern<<<grid, TE_THREADS, TE_SMEM_BYTES, s>>>(tmA, tmB, g, ep);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int milpool_tc_fwd(const float* x, long long sx_seq, long long sx_tok, const uint8_t* mask, long long smask, const float* V,
                   const float* bV, const float* U, const float* bU, const float* w, const float* bw, int S, int L, int D,
                   int Hd, float drop_p, unsigned long long seed, void* x3, void* w3, void* wt3, float* tg, float* spart,
                   float* attn, float* opart, float* out, cudaStream_t s) {
  MilParams p;
  if (int rc = mil_fill(p, x, sx_seq, sx_tok, S, L, D, Hd, V, U, w, drop_p, seed)) return rc;
  if (!milpool_tc_ok(p.R, L, D, Hd)) return B2_ENOSYS;
  long long plan[5];
  milpool_tc_plan(S, L, D, Hd, plan);
  if (!bV || !bU || !bw || !x3 || !w3 || !wt3 || !tg || !spart || !attn || !out || (p.P > 1 && !opart)) return B2_EINVAL;
  p.bV = bV; p.bU = bU; p.bw = bw; p.mask = mask; p.smask = smask;
  p.tg = tg; p.spart = spart; p.attn = attn; p.opart = opart; p.out = out;
  p.nut = (int)plan[2];
  const int Hp = (int)plan[3];
  mil_split_x_kernel<<<(unsigned)((p.R * (D / 4) + 255) / 256), 256, 0, s>>>(p, static_cast<__nv_bfloat16*>(x3));
  mil_prep_w_kernel<<<(unsigned)(((long long)Hp * D + 255) / 256), 256, 0, s>>>(p, Hp, static_cast<__nv_bfloat16*>(w3),
                                                                                static_cast<__nv_bfloat16*>(wt3));
  MilGateEpiParams ep{bV, bU, w, tg, spart, Hd, p.R};
  if (int rc = mil_launch_te<MilGateEpi>(x3, w3, (int)p.R, 2 * Hd, 3 * D, ep, s)) return rc;
  const size_t smem = (size_t)L * sizeof(float);
  if (smem > 48 * 1024) cudaFuncSetAttribute(mil_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mil_pool_fwd_kernel<<<dim3(p.P, S), 256, smem, s>>>(p);
  if (p.P > 1) mil_pool_combine_kernel<<<(unsigned)(((long long)S * D + 255) / 256), 256, 0, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

// one3: device floats with one3[2] == 1 (the scale slot gt_gemm reads). dW [2 Hd, D] comes back with INTERLEAVED rows
// (2u = dV_u, 2u + 1 = dU_u); dsmall as in milpool_bwd.
int milpool_tc_bwd(const float* x, long long sx_seq, long long sx_tok, const float* w, int S, int L, int D, int Hd,
                   float drop_p, unsigned long long seed, const void* x3, const void* wt3, const float* tg, const float* attn,
                   const float* dout, float* ds, float* dx, void* dpre3, void* ghi, void* glo, float* ad, float* fpart,
                   float* dW, float* dsmall, const float* one3, cudaStream_t s) {
  MilParams p;
  if (int rc = mil_fill(p, x, sx_seq, sx_tok, S, L, D, Hd, w, w, w, drop_p, seed)) return rc;      // V / U are not read here
  if (!milpool_tc_ok(p.R, L, D, Hd)) return B2_ENOSYS;
  if (!x3 || !wt3 || !tg || !attn || !dout || !ds || !dx || !dpre3 || !ghi || !glo || !ad || !fpart || !dW || !dsmall || !one3)
    return B2_EINVAL;
  long long plan[5];
  milpool_tc_plan(S, L, D, Hd, plan);
  const int Hp = (int)plan[3], njb = 4 * ((2 * Hd + 255) / 256);
  p.tg = const_cast<float*>(tg); p.attn = const_cast<float*>(attn); p.dout = dout;
  p.ds = ds; p.dx = dx; p.fpart = fpart; p.dW = dW; p.dsmall = dsmall;
  p.chunks = (int)plan[1];
  p.Z = 0;                                           // mil_reduce: the bias / w sums only
  mil_dA_kernel<<<(unsigned)((p.R + 7) / 8), 256, 0, s>>>(p);
  mil_ds_kernel<<<S, 256, 0, s>>>(p);
  mil_dpre_tc_kernel<<<p.chunks, 256, 0, s>>>(p, Hp, njb, static_cast<__nv_bfloat16*>(dpre3), static_cast<__nv_bfloat16*>(ghi),
                                               static_cast<__nv_bfloat16*>(glo), ad);
  MilDxEpiParams ep{ad, dout, dx, D, L};
  if (int rc = mil_launch_te<MilDxEpi>(dpre3, wt3, (int)p.R, D, 3 * Hp, ep, s)) return rc;
  if (cudaMemsetAsync(dW, 0, (size_t)2 * Hd * D * sizeof(float), s) != cudaSuccess) return B2_ECUDA;
  // hi*hi + hi*lo + lo*hi in ONE pass (one accumulator, one drain): glo follows ghi in memory, x_hi / x_lo are panels of x3
  const long long ge = plan[4];
  if (static_cast<__nv_bfloat16*>(glo) != static_cast<__nv_bfloat16*>(ghi) + ge) return B2_EINVAL;
  const int nib = (int)plan[1];
  const int gsrc[3] = {0, 0, nib}, xsrc[3] = {D, 0, D};
  if (int rc = gt_gemm_multi(ghi, 2 * ge, 3, gsrc, (int)p.R, 2 * Hd, x3, 3 * D, 3 * D, xsrc, D, D, one3, 1.f, dW, D, s)) return rc;
  const long long n2 = 3 * Hd + 1;
  mil_reduce_kernel<<<(unsigned)((n2 + 7) / 8), 256, 0, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
