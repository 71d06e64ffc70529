// Epilogue policies and selection kernels of the streaming retrieval (K5 / K6), see retrieval.cu. Plain CUDA C++ without
// inline PTX and without any include besides te_ctx.cuh, so that tests/emul/ can compile this very file for the host
// under a CUDA emulation shim and run the logic on CPU (the tile engine that feeds the policies is tcgen05 code and is
// covered by the -m gpu tests only).
#pragma once
#include <stdint.h>
#include "te_ctx.cuh"

namespace b2 {

struct RetrParams {
  const float* sgt;        // [Ma] ground-truth similarity per row (null: no rank counting)
  const long long* gt;     // [Ma] ground-truth GLOBAL column per row
  int col_offset;          // global index of B row 0 (text shard offset)
  int* counts;             // [Ma] += number of columns ranked before the ground truth
  float* part_score;       // [Ma][slots][k] partial top-k lists (null when kMaxK == 0)
  int* part_idx;
  int slots;
  int k;
};

template <int kMaxK>
struct RetrEpi {
  using Params = RetrParams;
  static constexpr int KL = kMaxK > 0 ? kMaxK : 1;
  struct State {
    float ls[KL];
    int li[KL];
    float thr;
    int cnt;
    float sg;
    int g;        // ground-truth column relative to this shard (may be out of range)
  };
  __device__ static __forceinline__ void init(State&, const Params&) {}
  __device__ static __forceinline__ void begin_outer(State& st, const Params& p, int, const TeCtx& ctx) {
#pragma unroll
    for (int i = 0; i < KL; ++i) {
      st.ls[i] = -INFINITY;
      st.li[i] = 0x7fffffff;
    }
    st.thr = -INFINITY;
    st.cnt = 0;
    st.sg = 0.f;
    st.g = -1;
    if (p.sgt && ctx.row_ok) {
      st.sg = p.sgt[ctx.row];
      const long long gg = p.gt[ctx.row] - p.col_offset;
      st.g = gg < -1 ? -1 : (gg > 0x3fffffff ? 0x3fffffff : (int)gg);
    } else if (p.sgt) {
      st.g = 0x3fffffff;
    }
  }
  __device__ static __forceinline__ void insert(State& st, float s, int idx, int k) {
    float cs = s;
    int ci = idx;
    bool ins = false;
#pragma unroll
    for (int q = 0; q < KL; ++q) {
      const bool b = ins || (cs > st.ls[q]);
      const float ts = b ? st.ls[q] : cs;
      const int ti = b ? st.li[q] : ci;
      st.ls[q] = b ? cs : st.ls[q];
      st.li[q] = b ? ci : st.li[q];
      cs = ts;
      ci = ti;
      ins = b;
    }
    float t = st.ls[0];
#pragma unroll
    for (int q = 1; q < KL; ++q) t = (q == k - 1) ? st.ls[q] : t;
    st.thr = (k == 1) ? st.ls[0] : t;
  }
  __device__ static __forceinline__ void chunk(State& st, const Params& p, const TeCtx& ctx, int c,
                                               const uint32_t (&acc)[32]) {
    const int cbase = ctx.col0 + c * 32;               // shard-local column of element 0
    const int nvalid = ctx.Nb - cbase;                 // elements e < nvalid are real columns
    if (p.sgt) {
      // ---- rank counting ----
      const int gl = st.g - cbase;                     // position of the ground truth inside this chunk
      int cnt = 0;
      if (nvalid >= 32 && (gl >= 32 || gl < 0)) {
        if (gl >= 32) {
#pragma unroll
          for (int e = 0; e < 32; ++e) cnt += (__uint_as_float(acc[e]) >= st.sg) ? 1 : 0;
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) cnt += (__uint_as_float(acc[e]) > st.sg) ? 1 : 0;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float s = __uint_as_float(acc[e]);
          const bool before = e < gl;
          const bool better = before ? (s >= st.sg) : (s > st.sg);
          cnt += (e < nvalid && e != gl && better) ? 1 : 0;
        }
      }
      st.cnt += cnt;
    }
    if (kMaxK > 0) {
      // ---- top-k ----
      // interior chunks (all 32 columns real — warp-uniform) skip the per-element bound test: the epilogue, not the tensor
      // pipe, bounds this sweep, so every instruction of the branch-free part counts
      float cmax = -INFINITY;
      if (nvalid >= 32) {
#pragma unroll
        for (int e = 0; e < 32; ++e) cmax = fmaxf(cmax, __uint_as_float(acc[e]));
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) cmax = fmaxf(cmax, (e < nvalid) ? __uint_as_float(acc[e]) : -INFINITY);
      }
      if (cmax > st.thr && ctx.row_ok) {
        uint32_t m = 0;
#pragma unroll
        for (int e = 0; e < 32; ++e) m |= (__uint_as_float(acc[e]) > st.thr) ? (1u << e) : 0u;
        if (nvalid < 32) m = nvalid <= 0 ? 0u : (m & ((1u << nvalid) - 1u));
        while (m) {
          const int e = __ffs(m) - 1;
          m &= m - 1;
          uint32_t v = acc[0];
#pragma unroll
          for (int q = 1; q < 32; ++q) v = (q == e) ? acc[q] : v;
          const float s = __uint_as_float(v);
          if (s > st.thr) insert(st, s, p.col_offset + cbase + e, p.k);
        }
      }
    }
  }
  __device__ static __forceinline__ void end_tile(State&, const Params&, const TeCtx&) {}
  __device__ static __forceinline__ void end_outer(State& st, const Params& p, int, const TeCtx& ctx) {
    if (!ctx.row_ok) return;
    if (p.sgt && st.cnt) atomicAdd(p.counts + ctx.row, st.cnt);
    if (kMaxK > 0) {
      const int slot = ctx.seg * 2 + ctx.wg;
      const size_t base = ((size_t)ctx.row * p.slots + slot) * p.k;
#pragma unroll
      for (int q = 0; q < KL; ++q)
        if (q < p.k) {
          p.part_score[base + q] = st.ls[q];
          p.part_idx[base + q] = st.li[q];
        }
    }
  }
};

// ---- two-sweep top-k (opt-in, B200CLIP_TOPK2=1): threshold first, then collect ----
// The register lists of RetrEpi make the WHOLE warp pay for every insertion of any of its 32 rows (~84 insertions per list,
// so nearly every 32-column chunk takes the slow path: 5x the counts-only sweep at k = 10). Two cheap sweeps instead:
//   sweep 1 (ColMaxEpi): per (row, slot = segment x column half) the maximum of each of the 32 column residue classes
//     (element e of every chunk) — one FMNMX per element, no branches. The subsets are disjoint, so the k-th largest of a
//     row's subset maxima (kth_largest_kernel) is a LOWER BOUND tau_i of its k-th best score.
//   sweep 2 (CollectEpi): every s_ij >= tau_i is appended to a per-row candidate buffer (one atomicAdd per chunk that
//     has a hit; ~12 hits per row at k = 10 with 64 subsets), which contains the exact top-k set including all ties at
//     the k-th score; topk_merge then orders it by (score desc, index asc). A row whose buffer overflows raises a flag
//     and the caller falls back to the register-list sweep (exactness is never traded).
struct ColMaxParams {
  float* part_max;   // [Ma][slots][32]
  int slots;
};
struct ColMaxEpi {
  using Params = ColMaxParams;
  struct State { float mx[32]; };
  __device__ static __forceinline__ void init(State&, const Params&) {}
  __device__ static __forceinline__ void begin_outer(State& st, const Params&, int, const TeCtx&) {
#pragma unroll
    for (int e = 0; e < 32; ++e) st.mx[e] = -INFINITY;
  }
  __device__ static __forceinline__ void chunk(State& st, const Params&, const TeCtx& ctx, int c,
                                               const uint32_t (&acc)[32]) {
    const int nvalid = ctx.Nb - (ctx.col0 + c * 32);
    if (nvalid >= 32) {
#pragma unroll
      for (int e = 0; e < 32; ++e) st.mx[e] = fmaxf(st.mx[e], __uint_as_float(acc[e]));
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) st.mx[e] = fmaxf(st.mx[e], e < nvalid ? __uint_as_float(acc[e]) : -INFINITY);
    }
  }
  __device__ static __forceinline__ void end_tile(State&, const Params&, const TeCtx&) {}
  __device__ static __forceinline__ void end_outer(State& st, const Params& p, int, const TeCtx& ctx) {
    if (!ctx.row_ok) return;
    float4* dst = reinterpret_cast<float4*>(p.part_max + ((size_t)ctx.row * p.slots + (ctx.seg * 2 + ctx.wg)) * 32);
#pragma unroll
    for (int e = 0; e < 8; ++e) dst[e] = make_float4(st.mx[4 * e], st.mx[4 * e + 1], st.mx[4 * e + 2], st.mx[4 * e + 3]);
  }
};

struct CollectParams {
  const float* thr;   // [Ma] per-row lower bound of the k-th best score
  int col_offset;     // global index of B row 0 (text shard offset)
  int* cnt;           // [Ma] candidates appended so far (zeroed by the caller)
  float* buf_s;       // [Ma][cap]
  int* buf_i;         // [Ma][cap], pre-filled with 0x7fffffff (= empty for topk_merge)
  int cap;
  int* overflow;      // set to 1 if any row had more than cap candidates
};
struct CollectEpi {
  using Params = CollectParams;
  struct State { float thr; };
  __device__ static __forceinline__ void init(State&, const Params&) {}
  __device__ static __forceinline__ void begin_outer(State& st, const Params& p, int, const TeCtx& ctx) {
    st.thr = ctx.row_ok ? p.thr[ctx.row] : INFINITY;
  }
  __device__ static __forceinline__ void chunk(State& st, const Params& p, const TeCtx& ctx, int c,
                                               const uint32_t (&acc)[32]) {
    const int cbase = ctx.col0 + c * 32;
    const int nvalid = ctx.Nb - cbase;
    // cheap pre-test (one FMNMX per element): most chunks hold no candidate of this row; padded columns read 0 and can
    // only make the pre-test pass spuriously — the exact mask below ignores them
    float cmax = -INFINITY;
#pragma unroll
    for (int e = 0; e < 32; ++e) cmax = fmaxf(cmax, __uint_as_float(acc[e]));
    if (!(cmax >= st.thr) || !ctx.row_ok) return;
    uint32_t m = 0;
#pragma unroll
    for (int e = 0; e < 32; ++e) m |= (__uint_as_float(acc[e]) >= st.thr) ? (1u << e) : 0u;
    if (nvalid < 32) m = nvalid <= 0 ? 0u : (m & ((1u << nvalid) - 1u));
    if (m != 0u) {
      const int n = __popc(m);
      int pos = atomicAdd(p.cnt + ctx.row, n);
      if (pos + n > p.cap) *p.overflow = 1;
      float* bs = p.buf_s + (size_t)ctx.row * p.cap;
      int* bi = p.buf_i + (size_t)ctx.row * p.cap;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        if ((m >> e) & 1u) {
          if (pos < p.cap) {
            bs[pos] = __uint_as_float(acc[e]);
            bi[pos] = p.col_offset + cbase + e;
          }
          ++pos;
        }
      }
    }
  }
  __device__ static __forceinline__ void end_tile(State&, const Params&, const TeCtx&) {}
  __device__ static __forceinline__ void end_outer(State&, const Params&, int, const TeCtx&) {}
};

// K6: out[row][0..k) = best k of the row's `slots * kin` candidates by (score desc, index asc). One warp per row.
__device__ __forceinline__ unsigned long long retr_key(float s, int idx) {
  uint32_t u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (uint32_t)(0x7fffffff - idx);
}
__global__ void __launch_bounds__(256)
topk_merge_kernel(const float* __restrict__ ps, const int* __restrict__ pi, int rows, int cand, int k,
                  float* __restrict__ out_s, long long* __restrict__ out_i) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* s = ps + (size_t)warp * cand;
  const int* id = pi + (size_t)warp * cand;
  // keys are unique per (score, idx): round r takes the largest key below the previous winner
  unsigned long long last = ~0ull;
  for (int r = 0; r < k; ++r) {
    unsigned long long best = 0ull;
    for (int c = lane; c < cand; c += 32) {
      const int ix = id[c];
      const unsigned long long key = ix != 0x7fffffff ? retr_key(s[c], ix) : 0ull;
      if (key < last && key > best) best = key;
    }
    unsigned long long wbest = best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, wbest, o);
      wbest = other > wbest ? other : wbest;
    }
    last = wbest != 0ull ? wbest : 0ull;
    if (lane == 0) {
      if (wbest == 0ull) {
        out_s[(size_t)warp * k + r] = -INFINITY;
        out_i[(size_t)warp * k + r] = -1;
      } else {
        uint32_t u = (uint32_t)(wbest >> 32);
        u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
        out_s[(size_t)warp * k + r] = __uint_as_float(u);
        out_i[(size_t)warp * k + r] = (long long)(0x7fffffff - (uint32_t)(wbest & 0xffffffffu));
      }
    }
  }
}

// thr[row] = k-th largest of the row's `cand` values counted with multiplicity (-inf when cand < k). One warp per row;
// round r takes the largest (value, position) key below the previous winner, as topk_merge does.
__global__ void __launch_bounds__(256)
kth_largest_kernel(const float* __restrict__ vals, int rows, int cand, int k, float* __restrict__ thr) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* v = vals + (size_t)warp * cand;
  unsigned long long last = ~0ull;
  for (int r = 0; r < k; ++r) {
    unsigned long long best = 0ull;
    for (int c = lane; c < cand; c += 32) {
      const unsigned long long key = retr_key(v[c], c);
      if (key < last && key > best) best = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    last = best;
    if (best == 0ull) break;            // fewer than k values
  }
  if (lane == 0) {
    float t = -INFINITY;
    if (last != 0ull && last != ~0ull) {
      uint32_t u = (uint32_t)(last >> 32);
      u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
      t = __uint_as_float(u);
    }
    thr[warp] = t;
  }
}

// hits[j] += #{rows: counts[row] < k_values[j]} (recall numerators, integer exact); mrr in double is done on the host
__global__ void __launch_bounds__(256)
recall_hits_kernel(const int* __restrict__ counts, int rows, const int* __restrict__ kvals, int nk,
                   unsigned long long* __restrict__ hits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (int j = 0; j < nk; ++j) {
    const bool hit = i < rows && counts[i] < kvals[j];
    const unsigned b = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(hits + j, (unsigned long long)__popc(b));
  }
}

// MRR numerator sum_i 1 / (counts[i] + 1) without a host pass over the rows: ranks are integers <= n_bins, so the sum is
// sum_r hist[r] / (r + 1) over the rank histogram — accumulated in fp64 in increasing-rank order by ONE CTA with a fixed
// reduction tree: deterministic and independent of the row order and of the text sharding (it differs from the
// reference's row-order double sum, retrieval_metrics_streaming.py:162-172, only by fp64 rounding, ~1e-16 relative).
__global__ void __launch_bounds__(256)
rank_hist_kernel(const int* __restrict__ counts, int rows, int n_bins, int* __restrict__ hist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) {
    int c = counts[i];
    c = c < 0 ? 0 : (c >= n_bins ? n_bins - 1 : c);
    atomicAdd(hist + c, 1);
  }
}
__global__ void __launch_bounds__(1024)
rank_hist_mrr_kernel(const int* __restrict__ hist, int n_bins, double* __restrict__ out) {
  double a = 0.0;
  for (int r = threadIdx.x; r < n_bins; r += blockDim.x) a += (double)hist[r] / (double)(r + 1);
  __shared__ double sh[1024];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}

}  // namespace b2
