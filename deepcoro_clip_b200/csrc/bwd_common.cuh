// Shared pieces of the logits-backward kernels (logits_bwd.cu: one CTA per item, cta_group::1;
// logits_bwd2.cu: CTA pairs, cta_group::2): parameters, the per-tile gradient epilogue and the accumulator drain.
#pragma once
#include "common.cuh"

namespace b2 {

constexpr int BW_BM = 128;     // X rows per item
constexpr int BW_BN = 128;     // Y rows per tile
constexpr int BW_BK = 64;
constexpr int BW_DP = 256;     // output columns per dp sweep
constexpr int BW_SLOTS = 6;
constexpr int BW_CHUNK = BW_BM * BW_BK * 2;           // 16 KB: [128 rows x 64 bf16]
constexpr int BW_XRES_CHUNKS = 8;                     // resident X panel: Kp <= 512
constexpr int BW_THREADS = 384;

template <bool kXRes>
struct BwSmem {
  static constexpr int kXBytes = kXRes ? BW_XRES_CHUNKS * BW_CHUNK : 0;
  static constexpr int kSlotBytes = kXRes ? BW_CHUNK : 2 * BW_CHUNK;
  static constexpr int kRingOff = kXBytes;
  static constexpr int kBarOff = kXBytes + BW_SLOTS * kSlotBytes;
  static constexpr int kColOff = kBarOff + 256;
  static constexpr int kBytes = kColOff + 2 * 128 * 4 + 1024 /*alignment slack*/;
};

enum { BW_CLIP = 0, BW_GATED = 1, BW_SIGLIP = 2, BW_SIGLIP_ENT = 3 };
// BW_SIGLIP_ENT = BW_SIGLIP plus the gradient of the entropy regulariser (utils/loss/contrastive.py:19-68):
//   G_ij += ent_coef * p_ij * (h_ij - m_i) * [|R| <= clamp],  p_ij = exp(L_ij - 30) / Z_i,
//   h_ij = -ln(p_ij + 1e-10) - p_ij / (p_ij + 1e-10),  m_i = sum_j p_ij h_ij.
// The per-video pairs {1/Z_i, m_i} arrive through `rowscale` ([Nx][2], X = video) or `colscale` ([Ny][2], Y = video).
template <int kMode>
struct BwIsSiglip { static constexpr bool value = kMode == BW_SIGLIP || kMode == BW_SIGLIP_ENT; };

struct BwParams {
  int Nx, Ny;          // valid rows of X and Y
  int Kp;              // K of the S product (multiple of 64; 3*Dp in bf16x3 mode)
  int Dp;              // padded width of the hi panel (multiple of 64): out-product columns come from Y[:, :Dp]
  int D;               // valid output columns
  int hi_off;          // column offset of the hi panel inside Y (0 plain bf16, 2*Dp in bf16x3 mode)
  float ydiag;         // CLIP: subtracted from G where (row + diag_off == column) BEFORE the bf16 rounding, so the
  int diag_off;        //   diagonal target (1-eps)/N cancels against P_ii(...) at full precision; 0 disables
  float* diag_corr;    // optional [Nx][2]: {g_ii - bf16(g_ii), bf16(g_ii)} for the fp32 fix-up in l2norm_bwd
  int x_tiles, y_tiles, dparts, nseg;
  float scale2, shift2;        // CLIP: P = 2^(f(S)*scale2 - shift2)
  float inv_tau, bias, wneg_c; // SigLIP: R = S*inv_tau + bias ; G = wneg_c * (sigmoid(clamp R) - yneg) * [|R|<=lclamp]
  float lclamp;                // logit clamp (30; 3e38 for the SigLIP2 BCE variants that do not clamp)   dyn[8]
  float yneg;                  // target of the non-positive pairs (label smoothing eps/2, default 0)      dyn[9]
  float ent_coef;              // BW_SIGLIP_ENT: -entropy_weight / B_global while the deficit is positive  dyn[10]
  const float* rowscale;       // [Nx]  c / rowsum_x  (CLIP)
  const float* colscale;       // [Ny]  c / colsum_y  (CLIP)
  float out_scale;             // 1 / tau
  float gnorm;                 // G is formed, rounded (bf16) and fed to the tensor core as G*gnorm = O(1); the
                               // accumulator and the scalar sums are multiplied back by 1/gnorm
  int hp;                      // 1: G is split into bf16 hi + lo (two TS-MMAs per K step): gradient rounding error
                               // 2^-17 instead of 2^-9; used together with the bf16x3 operands on small problems
  float* dX;                   // [Nx, ldd] fp32, accumulated with atomics
  int ldd;
  double* scal;                // [4] fp64 atomics: 0: sum G*f(S), 1: sum softplus(L), 2: sum G ; may be null
  const float* dyn;            // optional device block from dyn_prep: overrides scale2/shift2/inv_tau/bias/out_scale
  int gnjb;                    // column blocks per block row of the stored-G buffer (4 * ceil(Ny / 256))
  int gstore;                  // logits_bwd3.cu only: every G tile is also stored (bf16, scaled by gnorm) through the kernel's
                               // fourth tensor map, for the transposed product of gt_gemm.cu
  int stable;                  // CLIP / gated only, read from dyn[11] inside the kernel: rowscale / colscale hold log2-domain
                               // log-sum-exps (minus log2 c) and G = 2^(L2 - rowscale_i) + 2^(L2 - colscale_j), two
                               // exponentials that are each <= c, instead of P * (c / rowsum_i + c / colsum_j) with the
                               // fixed shift (tau below the window of dyn_prep, down to the reference's floor 1e-4)
};

// Row / column statistic in the form the epilogue consumes it. Fixed shift: scale * gnorm (0 outside the problem: no
// contribution). Stable: lse2 - log2(gnorm) (+huge outside the problem: 2^(L2 - huge) = 0).
__device__ __forceinline__ float bw_stat(const BwParams& p, const float* __restrict__ v, int i, bool ok) {
  if (p.stable) return ok ? __ldg(v + i) - log2f(p.gnorm) : 3.0e38f;
  return ok ? __ldg(v + i) * p.gnorm : 0.f;
}

// c * gnorm * (softmax over the row + softmax over the column) of one logit, f = f(S)
template <bool kStable>
__device__ __forceinline__ float bw_softmax_g(float f, float scale2, float nshift2, float rs, float cs) {
  if (kStable) {
    const float l2 = f * scale2;
    return ex2_approx(l2 - rs) + ex2_approx(l2 - cs);
  }
  return ex2_approx(fmaf(f, scale2, nshift2)) * (rs + cs);
}

__device__ __forceinline__ void lds128(uint32_t addr, float (&v)[4]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}


// Per-thread constants of one epilogue thread (one TMEM lane = one X row, 64 of the 128 tile columns).
struct BwThread {
  int row;          // global X row
  bool row_ok;
  int wg;           // column half of the S tile / accumulator
  float rs;         // rowscale[row] * gnorm (CLIP / gated)
  float ydn, wn, ign, nshift2;
  float ent_iz = 0.f, ent_m = 0.f;   // BW_SIGLIP_ENT with X = video: 1/Z_row, m_row
};

// Interior tile of bw_g_tile (no bounds / diagonal tests, single bf16 gradient operand): 64 S values of one thread ->
// packed bf16 G, written back in place.
template <int kMode, bool kStable>
__device__ __forceinline__ void bw_g_fast(const BwParams& p, const BwThread& th, const uint32_t (&acc)[2][32],
                                          uint32_t sbase, uint32_t cs_addr, bool want_scal, float& tacc, float& lacc,
                                          float& bacc) {
  constexpr bool kSig = BwIsSiglip<kMode>::value;
  const float rs = th.rs, wn = th.wn, nshift2 = th.nshift2;
  const float lc = p.lclamp, yneg = p.yneg;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t packed[16];
#pragma unroll
      for (int e = 0; e < 32; e += 4) {
        float cs4[4] = {0.f, 0.f, 0.f, 0.f};
        if (!kSig) lds128(cs_addr + (c * 32 + e) * 4, cs4);
        float g4[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float s = __uint_as_float(acc[c][e + h]);
          float g;
          if (kSig) {
            const float R = fmaf(s, p.inv_tau, p.bias);
            const float Lc = fminf(fmaxf(R, -lc), lc);
            const float ex = ex2_approx(-1.4426950408889634f * fabsf(Lc));
            const float den = 1.f + ex;
            const float r = __fdividef(1.f, den);
            const float sig = Lc >= 0.f ? r : ex * r;
            g = (fabsf(R) <= lc) ? wn * (sig - yneg) : 0.f;
            if (want_scal) {
              lacc += fmaf(-yneg, Lc, fmaxf(Lc, 0.f) + log1p_ex(ex));
              bacc += g;
              tacc = fmaf(g, s, tacc);
            }
          } else if (kMode == BW_GATED) {
            const float ex = ex2_approx(-1.4426950408889634f * s);
            const float sig = __fdividef(1.f, 1.f + ex);
            const float f = s * sig;
            const float fp = sig * (1.f + s * (1.f - sig));
            g = bw_softmax_g<kStable>(f, p.scale2, nshift2, rs, cs4[h]);
            if (want_scal) tacc = fmaf(g, f, tacc);
            g *= fp;
          } else {
            g = bw_softmax_g<kStable>(s, p.scale2, nshift2, rs, cs4[h]);
            if (want_scal) tacc = fmaf(g, s, tacc);
          }
          g4[h] = g;
        }
        packed[e >> 1] = pack_bf16x2(g4[0], g4[1]);
        packed[(e >> 1) + 1] = pack_bf16x2(g4[2], g4[3]);
      }
      tmem_st16(sbase + c * 16, packed);
    }
}

// G tile of one step: S (fp32, TMEM) -> elementwise gradient -> bf16x2 packed IN PLACE (tcgen05.st).
// sbase: TMEM address of this thread's 64 S columns; cs_addr / cs: staged colscale*gnorm of the tile (shared).
template <int kMode>
__device__ __forceinline__ void bw_g_tile(const BwParams& p, const BwThread& th, uint32_t sbase, uint32_t cs_addr,
                                          const float* cs, int xt, int j, int dp, bool want_scal, float& tacc,
                                          float& lacc, float& bacc) {
  constexpr bool kSig = BwIsSiglip<kMode>::value;
  constexpr bool kEnt = kMode == BW_SIGLIP_ENT;
  const int row = th.row, wg = th.wg;
  const bool row_ok = th.row_ok;
  const float rs = th.rs, ydn = th.ydn, wn = th.wn, ign = th.ign, nshift2 = th.nshift2;
  const float lc = p.lclamp, yneg = p.yneg;
  const bool full = (xt * BW_BM + BW_BM <= p.Nx) && (j * BW_BN + BW_BN <= p.Ny);
  // does the target diagonal cross this tile? (block-uniform)
  const int dlo = xt * BW_BM + p.diag_off - j * BW_BN;
  const bool has_diag = !kSig && p.ydiag != 0.f && dlo > -BW_BM && dlo < BW_BN;
  const int dcol = row + p.diag_off - j * BW_BN - wg * 64;   // diagonal column relative to this thread's half
  uint32_t acc[2][32];
  tmem_ld32(sbase, acc[0]);
  tmem_ld32(sbase + 32, acc[1]);
  tc_wait_ld();
  if (full && !has_diag && !p.hp && !kEnt) {
    // -------- fast path: interior tile, no bounds / diagonal tests --------
    if (!kSig && p.stable)
      bw_g_fast<kMode, true>(p, th, acc, sbase, cs_addr, want_scal, tacc, lacc, bacc);
    else
      bw_g_fast<kMode, false>(p, th, acc, sbase, cs_addr, want_scal, tacc, lacc, bacc);
  } else {
    // -------- general path: edge tiles, diagonal tiles, hi+lo gradient operand --------
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t packed[16], packed_lo[16];
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        float g2[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float s = __uint_as_float(acc[c][e + h]);
          const int cl = wg * 64 + c * 32 + e + h;     // column inside the tile
          float g, f = s;
          if (kSig) {
            const float R = fmaf(s, p.inv_tau, p.bias);
            const float Lc = fminf(fmaxf(R, -lc), lc);
            const float ex = ex2_approx(-1.4426950408889634f * fabsf(Lc));
            const float den = 1.f + ex;
            const float r = __fdividef(1.f, den);
            const float sig = Lc >= 0.f ? r : ex * r;
            const bool inr = fabsf(R) <= lc;
            g = inr ? wn * (sig - yneg) : 0.f;
            float sp = fmaf(-yneg, Lc, fmaxf(Lc, 0.f) + log1p_ex(ex));
            if (kEnt) {
              // entropy regulariser: p = exp(L - 30) / Z_video, video = this row (rowscale) or this column (colscale)
              const int colg = j * BW_BN + cl;
              float iz = th.ent_iz, mv = th.ent_m;
              if (p.colscale) {
                const bool cok = colg < p.Ny;
                iz = cok ? __ldg(p.colscale + 2 * colg) : 0.f;
                mv = cok ? __ldg(p.colscale + 2 * colg + 1) : 0.f;
              }
              const float pij = ex2_approx((Lc - 30.f) * 1.4426950408889634f) * iz;
              const float pe = pij + 1e-10f;
              const float h = -0.6931471805599453f * lg2_approx(pe) - __fdividef(pij, pe);
              if (inr) g = fmaf(p.ent_coef * p.gnorm, pij * (h - mv), g);
            }
            if (!full && !(row_ok && (j * BW_BN + cl) < p.Ny)) { g = 0.f; sp = 0.f; }
            lacc += sp;
            bacc += g;
            tacc = fmaf(g, s, tacc);
          } else {
            float fp = 1.f;
            if (kMode == BW_GATED) {
              const float ex = ex2_approx(-1.4426950408889634f * s);
              const float sig = __fdividef(1.f, 1.f + ex);
              f = s * sig;
              fp = sig * (1.f + s * (1.f - sig));
            }
            g = p.stable ? bw_softmax_g<true>(f, p.scale2, nshift2, rs, cs[cl])
                         : bw_softmax_g<false>(f, p.scale2, nshift2, rs, cs[cl]);
            if (has_diag && (c * 32 + e + h) == dcol) g -= ydn;
            if (!full && !(row_ok && (j * BW_BN + cl) < p.Ny)) g = 0.f;
            tacc = fmaf(g, f, tacc);
            if (kMode == BW_GATED) g *= fp;
            if (has_diag && (c * 32 + e + h) == dcol && dp == 0 && row_ok && p.diag_corr) {
              float gb = __bfloat162float(__float2bfloat16_rn(g));
              if (p.hp) gb += __bfloat162float(__float2bfloat16_rn(g - gb));
              p.diag_corr[2 * row] = (g - gb) * ign;
              p.diag_corr[2 * row + 1] = gb * ign;
            }
          }
          g2[h] = g;
        }
        packed[e >> 1] = pack_bf16x2(g2[0], g2[1]);
        if (p.hp) {
          const float r0 = g2[0] - __bfloat162float(__float2bfloat16_rn(g2[0]));
          const float r1 = g2[1] - __bfloat162float(__float2bfloat16_rn(g2[1]));
          packed_lo[e >> 1] = pack_bf16x2(r0, r1);
        }
      }
      tmem_st16(sbase + c * 16, packed);
      if (p.hp) tmem_st16(sbase + 32 + c * 16, packed_lo);
    }
    if (!want_scal) { tacc = 0.f; lacc = 0.f; bacc = 0.f; }
  }
  tc_wait_st();
}

// Accumulator of a finished dp sweep -> registers -> red.global.add into dX. abase: TMEM address of this thread's
// 128 accumulator columns.
__device__ __forceinline__ void bw_drain(const BwParams& p, const BwThread& th, uint32_t abase, int dp) {
  const int row = th.row, wg = th.wg;
  const bool row_ok = th.row_ok;
  const float ign = th.ign;
  float* drow = p.dX + (size_t)row * p.ldd + dp * BW_DP + wg * 128;
  const int cvalid = p.D - (dp * BW_DP + wg * 128);     // valid columns in this half
  const float osc = p.out_scale * ign;
  const bool vec_ok = (p.ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dX) & 15) == 0;
  uint32_t a[32];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c * 32 < cvalid) {     // warp-uniform
      tmem_ld32(abase + c * 32, a);
      tc_wait_ld();
      if (row_ok) {
        if (vec_ok && c * 32 + 32 <= cvalid) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            red_add_v4(drow + c * 32 + e, __uint_as_float(a[e]) * osc, __uint_as_float(a[e + 1]) * osc,
                       __uint_as_float(a[e + 2]) * osc, __uint_as_float(a[e + 3]) * osc);
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c * 32 + e < cvalid) atomicAdd(drow + c * 32 + e, __uint_as_float(a[e]) * osc);
        }
      }
    }
  }
}

}  // namespace b2
