// K2: fused logits forward for the softmax (CLIP / legacy gated) losses.
//   S = A B^T on tcgen05 tiles; epilogue computes P = 2^(f(S)*scale2 - shift2) once per element and
//   accumulates BOTH the row sums (over columns) and the column sums (over rows). The N x N matrix never
//   leaves TMEM / registers. Reference math: utils/loss/contrastive.py:146-162 (CLIPLoss),
//   utils/loss/losses.py:190-210 (gated SiglipLoss: f(s) = s*sigmoid(s)).
//
//   Column sums: each epilogue thread owns one A row (TMEM lane) and 128 columns; it keeps 128 fp32
//   column accumulators in registers across the whole sweep over A tiles of one 256-column B block and
//   reduces them across lanes only when the block changes (warp shuffles + one atomicAdd per column per
//   warp). Row sums: one atomicAdd per thread per tile.
#include "tile_engine2.cuh"
#include <stdlib.h>

namespace b2 {

struct LseParams {
  float scale2;    // log2(e) / tau
  float shift2;    // subtract after scaling (keeps 2^x inside fp32 range)
  float* rowsum;   // [Ma] += sum_j P_ij
  float* colsum;   // [Nb] += sum_i P_ij
  int gated;       // 1: f(s) = s * sigmoid(s) (legacy "siglip" gating), 0: f(s) = s
  const float* dyn;   // optional device block from dyn_prep (overrides scale2 / shift2): no host sync on tau
  float* diag;        // optional [Ma]: diag[i] = S[i, i + diag_off] exactly as the tensor core produced it, so the
  int diag_off;       //   target logit and the row/column LSE share the same rounding (they cancel in the loss)
};

template <bool kGated>
struct LseEpi {
  using Params = LseParams;
  struct State {
    float colacc[4][32];
    float rowacc;
    float scale2, shift2;
  };
  __device__ static __forceinline__ void init(State& st, const Params& p) {
    st.scale2 = p.dyn ? p.dyn[0] : p.scale2;
    st.shift2 = p.dyn ? p.dyn[1] : p.shift2;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int e = 0; e < 32; ++e) st.colacc[c][e] = 0.f;
    st.rowacc = 0.f;
  }
  __device__ static __forceinline__ void begin_outer(State&, const Params&, int, const TeCtx&) {}
  __device__ static __forceinline__ float prob(float s, const State& p) {
    if (kGated) {
      // s * sigmoid(s) = s / (1 + 2^(-s*log2e))
      const float e = ex2_approx(-1.4426950408889634f * s);
      s = __fdividef(s, 1.f + e);
    }
    return ex2_approx(fmaf(s, p.scale2, -p.shift2));
  }
  __device__ static __forceinline__ void chunk(State& st, const Params& p, const TeCtx& ctx, int c,
                                               const uint32_t (&acc)[32]) {
    if (p.diag) {
      const int d = ctx.row + p.diag_off - (ctx.col0 + c * 32);
      if (d >= 0 && d < 32 && ctx.row_ok) {
        uint32_t v = 0;
#pragma unroll
        for (int e = 0; e < 32; ++e) v = (e == d) ? acc[e] : v;
        p.diag[ctx.row] = __uint_as_float(v);
      }
    }
    if (ctx.full) {
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float v = prob(__uint_as_float(acc[e]), st);
        st.rowacc += v;
        st.colacc[c][e] += v;
      }
    } else {
      const int cbase = ctx.col0 + c * 32;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        float v = prob(__uint_as_float(acc[e]), st);
        v = (ctx.row_ok && (cbase + e) < ctx.Nb) ? v : 0.f;
        st.rowacc += v;
        st.colacc[c][e] += v;
      }
    }
  }
  __device__ static __forceinline__ void end_tile(State& st, const Params& p, const TeCtx& ctx) {
    if (ctx.row_ok) atomicAdd(p.rowsum + ctx.row, st.rowacc);
    st.rowacc = 0.f;
  }
  __device__ static __forceinline__ void end_outer(State& st, const Params& p, int outer, const TeCtx& ctx) {
    // cross-lane reduce of the 128 per-thread column accumulators; lane e keeps column c*32+e
    const int lane = threadIdx.x & 31;
    const int cb = outer * TE_BN + ctx.wg * 128;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float mine = 0.f;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float s = warp_sum(st.colacc[c][e]);
        if (lane == e) mine = s;
        st.colacc[c][e] = 0.f;
      }
      const int col = cb + c * 32 + lane;
      if (col < ctx.Nb) atomicAdd(p.colsum + col, mine);
    }
  }
};

// Stable mode of the softmax losses (dyn[11] != 0: tau below the fixed-shift window, down to the reference's clamp floor
// 1e-4, utils/loss/contrastive.py:153): per-row log-sum-exp with a RUNNING MAXIMUM. Tile order outer = A tile, so the
// epilogue thread owns its row across the whole sweep of B blocks and keeps (m, s) in two registers (online softmax in the
// log2 domain: s = sum_j 2^(L2_ij - m), L2 = f(S) * log2(e) / tau); one exponential per element, a rescale only when the
// maximum moves. Every (row, sweep segment, column half) writes its (m, s) pair; the LAST writer of a row (ticket counter)
// merges the pairs into lse2[row] = M + log2(sum_k s_k 2^(m_k - M)), so no second launch is needed. The column
// log-sum-exps of the symmetric loss are the row log-sum-exps of the role-swapped problem (A = text, B = video).
struct RowLseParams {
  const float* dyn;
  float2* part;      // [Ma][slots] (m, s) pairs
  int* ticket;       // [Ma], zero on entry
  float* lse2;       // [Ma] out
  float* diag;       // optional [Ma]: S[i, i + diag_off] with the tensor core's own rounding (the target logit)
  int diag_off;
  int slots;         // 2 * segs
  float* gap;        // optional [Ma] (needs diag): lse2[i] - L2_ii formed as (M - L2_ii) + log2(sum): no cancellation between
                     // two numbers of the size of the logits (up to 1.4e4 at tau = 1e-4), exactly log2(sum) when the target is
                     // the row maximum — the loss term of an almost separated batch keeps its relative accuracy, like the
                     // reference's log_softmax (x - max - log sum exp(x - max))
};

template <bool kGated>
struct RowLseEpi {
  using Params = RowLseParams;
  struct State {
    float m, s, scale2;
  };
  __device__ static __forceinline__ void init(State& st, const Params& p) {
    st.scale2 = p.dyn[0];
    st.m = -INFINITY;
    st.s = 0.f;
  }
  __device__ static __forceinline__ void begin_outer(State& st, const Params&, int, const TeCtx&) {
    st.m = -INFINITY;
    st.s = 0.f;
  }
  __device__ static __forceinline__ void chunk(State& st, const Params& p, const TeCtx& ctx, int c,
                                               const uint32_t (&acc)[32]) {
    if (p.diag) {
      const int d = ctx.row + p.diag_off - (ctx.col0 + c * 32);
      if (d >= 0 && d < 32 && ctx.row_ok) {
        uint32_t v = 0;
#pragma unroll
        for (int e = 0; e < 32; ++e) v = (e == d) ? acc[e] : v;
        p.diag[ctx.row] = __uint_as_float(v);
      }
    }
    const int cbase = ctx.col0 + c * 32;
    float l[32];
    float cm = -INFINITY;
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      float s = __uint_as_float(acc[e]);
      if (kGated) s = __fdividef(s, 1.f + ex2_approx(-1.4426950408889634f * s));
      float v = s * st.scale2;
      if (!ctx.full) v = (cbase + e) < ctx.Nb ? v : -INFINITY;
      l[e] = v;
      cm = fmaxf(cm, v);
    }
    if (cm > -INFINITY) {                     // at least one valid column in this chunk
      if (cm > st.m) {
        st.s *= ex2_approx(st.m - cm);        // m = -inf on the first chunk: 0 * 0
        st.m = cm;
      }
      float a = 0.f;
#pragma unroll
      for (int e = 0; e < 32; ++e) a += ex2_approx(l[e] - st.m);
      st.s += a;
    }
  }
  __device__ static __forceinline__ void end_tile(State&, const Params&, const TeCtx&) {}
  __device__ static __forceinline__ void end_outer(State& st, const Params& p, int, const TeCtx& ctx) {
    if (!ctx.row_ok) return;
    float2* pr = p.part + (size_t)ctx.row * p.slots;
    __stcg(pr + ctx.seg * 2 + ctx.wg, make_float2(st.m, st.s));
    __threadfence();
    if (atomicAdd(p.ticket + ctx.row, 1) == p.slots - 1) {
      __threadfence();
      float M = -INFINITY;
      for (int k = 0; k < p.slots; ++k) M = fmaxf(M, __ldcg(pr + k).x);
      float S = 0.f;
      for (int k = 0; k < p.slots; ++k) {
        const float2 v = __ldcg(pr + k);
        if (v.y > 0.f) S += v.y * ex2_approx(v.x - M);
      }
      const float lg = log2f(S);
      p.lse2[ctx.row] = M + lg;
      if (p.gap) {
        // the target logit exactly as chunk() formed it (same instructions on the same tensor-core value)
        float d = __ldcg(p.diag + ctx.row);
        if (kGated) d = __fdividef(d, 1.f + ex2_approx(-1.4426950408889634f * d));
        p.gap[ctx.row] = (M - d * st.scale2) + lg;
      }
      p.ticket[ctx.row] = 0;                  // re-armed for a replay of the same buffers (CUDA graph)
    }
  }
};

// out[i] = S[i, i] exactly as the tensor core produces it (used with TeShape::diag: only the diagonal tiles run).
struct DiagParams {
  float* out;
};
struct DiagEpi {
  using Params = DiagParams;
  struct State {};
  __device__ static __forceinline__ void init(State&, const Params&) {}
  __device__ static __forceinline__ void begin_outer(State&, const Params&, int, const TeCtx&) {}
  __device__ static __forceinline__ void chunk(State&, const Params& p, const TeCtx& ctx, int c,
                                               const uint32_t (&acc)[32]) {
    const int d = ctx.row - (ctx.col0 + c * 32);
    if (d >= 0 && d < 32 && ctx.row_ok) {
      uint32_t v = 0;
#pragma unroll
      for (int e = 0; e < 32; ++e) v = (e == d) ? acc[e] : v;
      p.out[ctx.row] = __uint_as_float(v);
    }
  }
  __device__ static __forceinline__ void end_tile(State&, const Params&, const TeCtx&) {}
  __device__ static __forceinline__ void end_outer(State&, const Params&, int, const TeCtx&) {}
};

// Debug / validation epilogue: dumps the raw fp32 S tile to global memory (only used by the self test).
struct DumpParams {
  float* out;
  int ld;
  int Ma;
};
struct DumpEpi {
  using Params = DumpParams;
  struct State {};
  __device__ static __forceinline__ void init(State&, const Params&) {}
  __device__ static __forceinline__ void begin_outer(State&, const Params&, int, const TeCtx&) {}
  __device__ static __forceinline__ void chunk(State&, const Params& p, const TeCtx& ctx, int c,
                                               const uint32_t (&acc)[32]) {
    if (!ctx.row_ok) return;
    const int cbase = ctx.col0 + c * 32;
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (cbase + e < ctx.Nb) p.out[(size_t)ctx.row * p.ld + cbase + e] = __uint_as_float(acc[e]);
  }
  __device__ static __forceinline__ void end_tile(State&, const Params&, const TeCtx&) {}
  __device__ static __forceinline__ void end_outer(State&, const Params&, int, const TeCtx&) {}
};

}  // namespace b2

namespace b2host {
using namespace b2;

int make_shape(TeShape& g, int Ma, int Nb, int Kp) {
  if (Ma <= 0 || Nb <= 0 || Kp <= 0 || (Kp % TE_BK) != 0) return B2_EINVAL;
  g.Ma = Ma;
  g.Nb = Nb;
  g.Kp = Kp;
  g.m_tiles = (Ma + TE_BM - 1) / TE_BM;
  g.n_blocks = (Nb + TE_BN - 1) / TE_BN;
  g.segs = 0;
  g.diag = 0;
  g.gate = nullptr;
  g.gate_on = 0;
  return B2_OK;
}

template <class Epi, bool kOuterIsB>
static int launch_te(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb,
                     const typename Epi::Params& ep, int max_ctas, cudaStream_t stream, int diag = 0,
                     const float* gate = nullptr, int gate_on = 0) {
  TeShape g;
  int rc = make_shape(g, Ma, Nb, Kp);
  if (rc) return rc;
  g.diag = diag;
  g.gate = gate;
  g.gate_on = gate_on;
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap_bf16_2d(&tmA, A, Ma, Kp, lda, TE_BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmB, B, Nb, Kp, ldb, TE_BN))) return rc;
  auto kern = te_kernel<Epi, kOuterIsB>;
  static bool attr_done_dev[64] = {};   // per template instantiation
  bool& attr_done = attr_done_dev[current_device() & 63];   // cudaFuncSetAttribute is per device
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TE_SMEM_BYTES) != cudaSuccess)
      return B2_ECUDA;
    attr_done = true;
  }
  long long total = diag ? (long long)g.m_tiles : (long long)g.m_tiles * g.n_blocks;
  int grid = sm_count();
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  if (total < grid) grid = (int)total;
  kern<<<grid, TE_THREADS, TE_SMEM_BYTES, stream>>>(tmA, tmB, g, ep);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

// CTA-pair engine (tile_engine2.cuh): Kp <= 1024 (first 512 columns of the outer operand resident).
// B200CLIP_TE_PAIR=0 keeps the single-CTA engine (A/B runs).
bool te_pair_enabled(int Kp) {
  static const bool on = [] { const char* e = getenv("B200CLIP_TE_PAIR"); return !(e && e[0] == '0'); }();
  return on && Kp <= 1024 && sm_count() >= 2;
}

template <class Epi, bool kOuterIsB>
static int launch_te2(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb,
                      const typename Epi::Params& ep, cudaStream_t stream, const float* gate = nullptr,
                      int gate_on = 0) {
  TeShape g;
  int rc = make_shape(g, Ma, Nb, Kp);
  if (rc) return rc;
  g.gate = gate;
  g.gate_on = gate_on;
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap_bf16_2d(&tmA, A, Ma, Kp, lda, TE_BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmB, B, Nb, Kp, ldb, 128))) return rc;      // each CTA loads 128 of the 256 block rows
  auto kern = te2_kernel<Epi, kOuterIsB>;
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device() & 63];   // cudaFuncSetAttribute is per device
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TE2_SMEM_BYTES) != cudaSuccess)
      return B2_ECUDA;
    attr_done = true;
  }
  const long long total = (long long)((g.m_tiles + 1) / 2) * g.n_blocks;
  long long clusters = sm_count() / 2;
  if (total < clusters) clusters = total;
  kern<<<(int)(2 * clusters), TE_THREADS, TE2_SMEM_BYTES, stream>>>(tmA, tmB, g, ep);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int logits_lse_fwd(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, float scale2,
                   float shift2, int gated, const float* dyn, int skip_if_stable, float* rowsum, float* colsum,
                   float* diag, int diag_off, cudaStream_t stream) {
  LseParams p{scale2, shift2, rowsum, colsum, gated, dyn, diag, diag_off};
  const float* gate = (dyn && skip_if_stable) ? dyn + 11 : nullptr;     // run only while dyn[11] == 0
  if (te_pair_enabled(Kp)) {
    if (gated) return launch_te2<LseEpi<true>, true>(A, B, Ma, Nb, Kp, lda, ldb, p, stream, gate, 0);
    return launch_te2<LseEpi<false>, true>(A, B, Ma, Nb, Kp, lda, ldb, p, stream, gate, 0);
  }
  if (gated) return launch_te<LseEpi<true>, true>(A, B, Ma, Nb, Kp, lda, ldb, p, 0, stream, 0, gate, 0);
  return launch_te<LseEpi<false>, true>(A, B, Ma, Nb, Kp, lda, ldb, p, 0, stream, 0, gate, 0);
}

// Sweep segments of the row-LSE pass: outer = A tile (pair), so few row tiles leave SMs idle unless the sweep over the B
// blocks is split (same rule as the retrieval sweep, never more segments than B blocks: every segment is non-empty).
int rowlse_slots(int Ma, int Nb, int Kp) {
  const int m_tiles = (Ma + TE_BM - 1) / TE_BM, n_blocks = (Nb + TE_BN - 1) / TE_BN;
  const int units = te_pair_enabled(Kp) ? sm_count() / 2 : sm_count();
  const int outer = te_pair_enabled(Kp) ? (m_tiles + 1) / 2 : m_tiles;
  int segs = 1;
  if (outer < 4 * units) {
    segs = (4 * units + outer - 1) / outer;
    if (segs > n_blocks) segs = n_blocks;
    if (segs > 64) segs = 64;
    if (segs < 1) segs = 1;
  }
  return 2 * segs;
}

template <bool kGated>
static int launch_rowlse(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, const RowLseParams& p,
                         int segs, const float* gate, cudaStream_t stream) {
  using Epi = RowLseEpi<kGated>;
  TeShape g;
  int rc = make_shape(g, Ma, Nb, Kp);
  if (rc) return rc;
  g.segs = segs;
  g.gate = gate;
  g.gate_on = 1;
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap_bf16_2d(&tmA, A, Ma, Kp, lda, TE_BM))) return rc;
  if (te_pair_enabled(Kp)) {
    if ((rc = make_tmap_bf16_2d(&tmB, B, Nb, Kp, ldb, 128))) return rc;
    auto kern2 = te2_kernel<Epi, false>;
    static bool attr2_done_dev[64] = {};
    bool& attr2_done = attr2_done_dev[current_device() & 63];
    if (!attr2_done) {
      if (cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, TE2_SMEM_BYTES) != cudaSuccess)
        return B2_ECUDA;
      attr2_done = true;
    }
    const long long items2 = (long long)((g.m_tiles + 1) / 2) * segs;
    long long clusters = sm_count() / 2;
    if (items2 < clusters) clusters = items2;
    kern2<<<(int)(2 * clusters), TE_THREADS, TE2_SMEM_BYTES, stream>>>(tmA, tmB, g, p);
    return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
  }
  if ((rc = make_tmap_bf16_2d(&tmB, B, Nb, Kp, ldb, TE_BN))) return rc;
  auto kern = te_kernel<Epi, false>;
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device() & 63];
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TE_SMEM_BYTES) != cudaSuccess)
      return B2_ECUDA;
    attr_done = true;
  }
  const long long items = (long long)g.m_tiles * segs;
  int grid = sm_count();
  if (items < grid) grid = (int)items;
  kern<<<grid, TE_THREADS, TE_SMEM_BYTES, stream>>>(tmA, tmB, g, p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int logits_rowlse(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, int gated, const float* dyn,
                  int only_if_stable, float* part, int slots, int* ticket, float* lse2, float* diag, int diag_off,
                  float* gap, cudaStream_t stream) {
  if (!dyn || !part || !ticket || !lse2 || slots != rowlse_slots(Ma, Nb, Kp) ||
      (reinterpret_cast<uintptr_t>(part) & 7) || (gap && !diag))
    return B2_EINVAL;
  RowLseParams p{dyn, reinterpret_cast<float2*>(part), ticket, lse2, diag, diag_off, slots, gap};
  const float* gate = only_if_stable ? dyn + 11 : nullptr;
  if (gated) return launch_rowlse<true>(A, B, Ma, Nb, Kp, lda, ldb, p, slots / 2, gate, stream);
  return launch_rowlse<false>(A, B, Ma, Nb, Kp, lda, ldb, p, slots / 2, gate, stream);
}

// out[i] = (A[i,:] . B[i,:]) with the tensor core's own rounding: bit-identical to the value any S = A B^T tile of the
// engine produces for that pair (same k-chunk order), which a CUDA-core dot product is not.
int rowdot_tc(const void* A, const void* B, int rows, int Kp, int lda, int ldb, float* out, cudaStream_t stream) {
  DiagParams p{out};
  return launch_te<DiagEpi, false>(A, B, rows, rows, Kp, lda, ldb, p, 0, stream, 1);
}

int logits_dump(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, float* out, int ldo,
                int max_ctas, cudaStream_t stream) {
  DumpParams p{out, ldo, Ma};
  return launch_te<DumpEpi, true>(A, B, Ma, Nb, Kp, lda, ldb, p, max_ctas, stream);
}

}  // namespace b2host
