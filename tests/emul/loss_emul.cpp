// Host library with the C-ABI entry points the CLIP loss path calls (include/b200clip.h), for running the package's own
// autograd function on CPU (tests/test_emulated_losses.py, single process and 2-rank gloo):
//   * the CUDA-core kernels are the SHIPPED device code (l2norm_kernels.cuh, scalars_kernels.cuh) under the emulation;
//   * the two tcgen05 tile kernels cannot be emulated; b200clip_logits_lse_fwd / b200clip_logits_bwd are MODELS that follow
//     the contracts written in include/b200clip.h operation by operation (bf16 operands, fp32 logits, G rounded to bf16
//     (hi + lo when hp) before the output product, diagonal target subtracted in fp32 before the rounding, diag_corr,
//     scal[0], gnorm / out_scale scaling). They exist to exercise the host code around the kernels, not to test the kernels.
#include "pool_mma_prims_emul.h"
#include "../../deepcoro_clip_b200/csrc/l2norm_kernels.cuh"
#include "../../deepcoro_clip_b200/csrc/scalars_kernels.cuh"
#include "../../deepcoro_clip_b200/csrc/siglip_kernels.cuh"
#include "../../deepcoro_clip_b200/csrc/retrieval_epi.cuh"
#include "../../deepcoro_clip_b200/csrc/alignment_diag.cuh"

using namespace b2;
using bf16 = __nv_bfloat16;

static inline float bf(const void* base, long long idx) { return __bfloat162float(static_cast<const bf16*>(base)[idx]); }

extern "C" int b200clip_l2norm_fwd(const void* x, int dtype, long long ldx, int rows, int dim, void* out, int ldo, int Kp,
                                   int split3_role, float* inv_norm, float* xhat_f32, int ldh, int normalize, void*) {
  if (rows <= 0 || dim <= 0 || Kp < dim || Kp % 64) return -22;
  const int blocks = (rows + 7) / 8;
  auto o = reinterpret_cast<bf16*>(out);
  switch (dtype) {
    case 0: emul::launch(blocks, 256, [&] { l2norm_fwd_kernel<float>((const float*)x, ldx, rows, dim, o, ldo, Kp, split3_role, inv_norm, xhat_f32, ldh, normalize); }); break;
    case 1: emul::launch(blocks, 256, [&] { l2norm_fwd_kernel<bf16>((const bf16*)x, ldx, rows, dim, o, ldo, Kp, split3_role, inv_norm, xhat_f32, ldh, normalize); }); break;
    case 2: emul::launch(blocks, 256, [&] { l2norm_fwd_kernel<__half>((const __half*)x, ldx, rows, dim, o, ldo, Kp, split3_role, inv_norm, xhat_f32, ldh, normalize); }); break;
    default: return -22;
  }
  return 0;
}

template <typename T>
static int bwd_t(const float* dxh, int ldg, const T* x, long ldx, const float* inv_norm, const void* ox, int odtype, long ldox,
                 const float* oinv, const bf16* ohi, int ldohi, const float2* dc, const float* usum, float gscale, float ucoef,
                 const float* omul, const float* gmul, int rows, int dim, float* dx, long lddx) {
  const int blocks = (rows + 7) / 8;
  switch (ox ? odtype : 0) {
    case 0: emul::launch(blocks, 256, [&] { l2norm_bwd_kernel<T, float>(dxh, ldg, x, ldx, inv_norm, (const float*)ox, ldox, oinv, ohi, ldohi, dc, usum, gscale, ucoef, omul, gmul, rows, dim, dx, lddx); }); break;
    case 1: emul::launch(blocks, 256, [&] { l2norm_bwd_kernel<T, bf16>(dxh, ldg, x, ldx, inv_norm, (const bf16*)ox, ldox, oinv, ohi, ldohi, dc, usum, gscale, ucoef, omul, gmul, rows, dim, dx, lddx); }); break;
    case 2: emul::launch(blocks, 256, [&] { l2norm_bwd_kernel<T, __half>(dxh, ldg, x, ldx, inv_norm, (const __half*)ox, ldox, oinv, ohi, ldohi, dc, usum, gscale, ucoef, omul, gmul, rows, dim, dx, lddx); }); break;
    default: return -22;
  }
  return 0;
}
extern "C" int b200clip_l2norm_bwd(const float* dxh, int ldg, const void* x, int dtype, long long ldx, const float* inv_norm,
                                   const void* ox, int odtype, long long ldox, const float* oinv, const void* ohi, int ldohi,
                                   const float* dc, const float* usum, float gscale, float ucoef, const float* omul,
                                   const float* gmul, int rows, int dim, float* dx, long long lddx, void*) {
  if (rows <= 0 || dim <= 0 || !dxh || !x || !inv_norm || !dx) return -22;
  if (dc && (!ox || !oinv || !ohi)) return -22;
  auto h = (const bf16*)ohi;
  auto d2 = (const float2*)dc;
  switch (dtype) {
    case 0: return bwd_t<float>(dxh, ldg, (const float*)x, ldx, inv_norm, ox, odtype, ldox, oinv, h, ldohi, d2, usum, gscale, ucoef, omul, gmul, rows, dim, dx, lddx);
    case 1: return bwd_t<bf16>(dxh, ldg, (const bf16*)x, ldx, inv_norm, ox, odtype, ldox, oinv, h, ldohi, d2, usum, gscale, ucoef, omul, gmul, rows, dim, dx, lddx);
    case 2: return bwd_t<__half>(dxh, ldg, (const __half*)x, ldx, inv_norm, ox, odtype, ldox, oinv, h, ldohi, d2, usum, gscale, ucoef, omul, gmul, rows, dim, dx, lddx);
    default: return -22;
  }
}

extern "C" int b200clip_colsum_bf16(const void* xh, int ld, int rows, int dim, float* out, void*) {
  if (rows <= 0 || dim <= 0 || !xh || !out) return -22;
  emul::launch(emul::Dim{(unsigned)((dim + 255) / 256), rows < 256 ? 1u : 64u, 1}, 256,
               [&] { colsum_bf16_kernel((const bf16*)xh, ld, rows, dim, out); });
  return 0;
}

template <typename TA>
static int rowdot_raw_t(const TA* a, long long lda, const float* ainv, const void* b, int bdtype, long long ldb,
                        const float* binv, int rows, int dim, float* out) {
  const int blocks = (rows + 7) / 8;
  switch (bdtype) {
    case 0: emul::launch(blocks, 256, [&] { rowdot_raw_kernel<TA, float>(a, lda, ainv, (const float*)b, ldb, binv, rows, dim, out); }); break;
    case 1: emul::launch(blocks, 256, [&] { rowdot_raw_kernel<TA, bf16>(a, lda, ainv, (const bf16*)b, ldb, binv, rows, dim, out); }); break;
    case 2: emul::launch(blocks, 256, [&] { rowdot_raw_kernel<TA, __half>(a, lda, ainv, (const __half*)b, ldb, binv, rows, dim, out); }); break;
    default: return -22;
  }
  return 0;
}
extern "C" int b200clip_rowdot_raw(const void* a, int adtype, long long lda, const float* ainv, const void* b, int bdtype,
                                   long long ldb, const float* binv, int rows, int dim, float* out, void*) {
  if (rows <= 0 || dim <= 0 || !a || !b || !ainv || !binv || !out) return -22;
  switch (adtype) {
    case 0: return rowdot_raw_t<float>((const float*)a, lda, ainv, b, bdtype, ldb, binv, rows, dim, out);
    case 1: return rowdot_raw_t<bf16>((const bf16*)a, lda, ainv, b, bdtype, ldb, binv, rows, dim, out);
    case 2: return rowdot_raw_t<__half>((const __half*)a, lda, ainv, b, bdtype, ldb, binv, rows, dim, out);
    default: return -22;
  }
}

extern "C" int b200clip_dyn_prep(const float* log_temp, const float* bias, float clamp_min, float bound, float* dyn, void*) {
  emul::launch(1, 32, [&] { dyn_prep_kernel(log_temp, bias, clamp_min, bound, dyn); });
  return 0;
}
extern "C" int b200clip_dyn_set_stable(float* dyn, int stable, void*) {
  emul::launch(1, 32, [&] { dyn_set_stable_kernel(dyn, stable); });
  return 0;
}
extern "C" int b200clip_clip_finalize(const float* sums, int n, int nvec, const float* dyn, float eps, int gated,
                                      const double* unif, float* rowscale, float* colscale, float* loss_out, double* acc_out,
                                      void*) {
  if (n <= 0 || !sums || !dyn || !rowscale || !colscale || !loss_out || (nvec != 3 && nvec != 7)) return -22;
  int blocks = (n + 1023) / 1024;
  if (blocks > FIN_MAX_BLOCKS) blocks = FIN_MAX_BLOCKS;
  FinPeers ps{};
  ps.ptr[0] = sums;
  ps.world = 0;
  ps.rows_per_rank = n;
  emul::launch(blocks, 1024, [&] { clip_finalize_kernel(ps, n, nvec, dyn, eps, gated, unif, rowscale, colscale, loss_out, acc_out); });
  return 0;
}
extern "C" int b200clip_clip_dlogtemp(const double* scal0, const float* dyn, const float* gmul, const double* unif, int n,
                                      float* out, void*) {
  if (!scal0 || !dyn || !gmul || !out || n <= 0) return -22;
  emul::launch(1, 32, [&] { clip_dlogtemp_kernel(scal0, dyn, gmul, unif, n, out); });
  return 0;
}

// ---------------- models of the two tcgen05 tile kernels (contracts: include/b200clip.h K2 / K3) ----------------
static inline float gate(float s) { return s / (1.f + expf(-s)); }

extern "C" int b200clip_logits_lse_fwd(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, float scale2,
                                       float shift2, int gated, const float* dyn, int skip_if_stable, float* rowsum,
                                       float* colsum, float* diag, int diag_off, void*) {
  if (dyn) { scale2 = dyn[0]; shift2 = dyn[1]; }
  if (dyn && skip_if_stable && dyn[11] != 0.f) return 0;          // launch gate: the stable sweeps run instead
  for (int i = 0; i < Ma; ++i)
    for (int j = 0; j < Nb; ++j) {
      float s = 0.f;
      for (int k = 0; k < Kp; ++k) s = fmaf(bf(A, (long long)i * lda + k), bf(B, (long long)j * ldb + k), s);
      if (diag && i + diag_off == j) diag[i] = s;
      const float p = exp2f(fmaf(gated ? gate(s) : s, scale2, -shift2));
      rowsum[i] += p;
      colsum[j] += p;
    }
  return 0;
}

// Model of the stable row log-sum-exp sweep (contract: include/b200clip.h K2s): lse2[i] = log2 sum_j 2^(f(S_ij) scale2),
// with a per-row maximum; the ticket array must arrive zeroed and is left zeroed.
extern "C" int b200clip_rowlse_slots(int, int, int) { return 2; }
extern "C" int b200clip_logits_rowlse(const void* A, const void* B, int Ma, int Nb, int Kp, int lda, int ldb, int gated,
                                      const float* dyn, int only_if_stable, float* part, int slots, int* ticket, float* lse2,
                                      float* diag, int diag_off, float* gap, void*) {
  if (!dyn || !part || !ticket || !lse2 || slots != 2 || (gap && !diag)) return -22;
  if (only_if_stable && dyn[11] == 0.f) return 0;
  const float scale2 = dyn[0];
  std::vector<float> l(Nb);
  for (int i = 0; i < Ma; ++i) {
    if (ticket[i] != 0) return -22;
    float m = -INFINITY;
    for (int j = 0; j < Nb; ++j) {
      float s = 0.f;
      for (int k = 0; k < Kp; ++k) s = fmaf(bf(A, (long long)i * lda + k), bf(B, (long long)j * ldb + k), s);
      if (diag && i + diag_off == j) diag[i] = s;
      l[j] = (gated ? gate(s) : s) * scale2;
      m = fmaxf(m, l[j]);
    }
    float sum = 0.f;
    for (int j = 0; j < Nb; ++j) sum += exp2f(l[j] - m);
    part[(size_t)i * 4] = m;
    part[(size_t)i * 4 + 1] = sum;
    lse2[i] = m + log2f(sum);
    if (gap) gap[i] = (m - (gated ? gate(diag[i]) : diag[i]) * scale2) + log2f(sum);
  }
  return 0;
}

extern "C" int b200clip_logits_bwd(int mode, const void* X, const void* Y, int Nx, int Ny, int Kp, int Dp, int D, int hi_off,
                                   int ldx, int ldy, float scale2, float shift2, float inv_tau, float bias, float wneg_c,
                                   const float* rowscale, const float* colscale, float out_scale, float gnorm, int hp,
                                   const float* dyn, float ydiag, int diag_off, float* diag_corr, float* dX, int ldd,
                                   double* scal, int, void*) {
  if (mode < 0 || mode > 3) return -38;
  (void)Dp;
  float lclamp = 30.f, yneg = 0.f;
  if (dyn) { scale2 = dyn[0]; shift2 = dyn[1]; inv_tau = dyn[2]; bias = dyn[5]; out_scale = dyn[2]; lclamp = dyn[8]; yneg = dyn[9]; }
  if (!(gnorm > 0.f)) gnorm = 1.f;
  const float ign = 1.f / gnorm, ydn = ydiag * gnorm;
  if (mode == 2 || mode == 3) {
    // G = wneg_c (sigmoid(clamp(R)) - yneg) [|R| <= lc], R = S / tau + bias; scal: [0] sum G S, [1] sum softplus, [2] sum G
    // mode 3 adds the entropy regulariser's gradient dyn[10] p_ij (h_ij - m_v) [|R| <= lc], p = exp(L - 30) / Z_v, with
    // {1 / Z_v, m_v} per video: rowscale [Nx][2] when X holds the videos, colscale [Ny][2] when Y does
    const float ent_coef = (mode == 3 && dyn) ? dyn[10] : 0.f;
    const float wn = wneg_c * gnorm;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int i = 0; i < Nx; ++i)
      for (int j = 0; j < Ny; ++j) {
        float sdot = 0.f;
        for (int k = 0; k < Kp; ++k) sdot = fmaf(bf(X, (long long)i * ldx + k), bf(Y, (long long)j * ldy + k), sdot);
        const float R = fmaf(sdot, inv_tau, bias);
        const float Lc = fminf(fmaxf(R, -lclamp), lclamp);
        const float ex = expf(-fabsf(Lc));
        const float sig = Lc >= 0.f ? 1.f / (1.f + ex) : ex / (1.f + ex);
        float g = fabsf(R) <= lclamp ? wn * (sig - yneg) : 0.f;
        if (mode == 3 && fabsf(R) <= lclamp) {
          const float iz = colscale ? colscale[2 * j] : rowscale[2 * i];
          const float mv = colscale ? colscale[2 * j + 1] : rowscale[2 * i + 1];
          const float pij = expf(Lc - 30.f) * iz;
          const float pe = pij + 1e-10f;
          g = fmaf(ent_coef * gnorm, pij * (-logf(pe) - pij / pe - mv), g);
        }
        t1 += (double)(fmaf(-yneg, Lc, fmaxf(Lc, 0.f) + log1pf(ex)));
        t2 += (double)g;
        t0 += (double)g * (double)sdot;
        float gb = __bfloat162float(__float2bfloat16_rn(g));
        if (hp) gb += __bfloat162float(__float2bfloat16_rn(g - gb));
        const float ge = gb * out_scale * ign;
        for (int d = 0; d < D; ++d) dX[(long long)i * ldd + d] += ge * bf(Y, (long long)j * ldy + hi_off + d);
      }
    if (scal) { scal[0] += t0 * (double)ign; scal[1] += t1; scal[2] += t2 * (double)ign; }
    return 0;
  }
  double tsum = 0.0;
  const bool stable = dyn && dyn[11] != 0.f;          // rowscale / colscale hold log2-domain log-sum-exps (minus log2 c)
  const float lg = log2f(gnorm);
  for (int i = 0; i < Nx; ++i) {
    const float rs = stable ? rowscale[i] - lg : rowscale[i] * gnorm;
    for (int j = 0; j < Ny; ++j) {
      float s = 0.f;
      for (int k = 0; k < Kp; ++k) s = fmaf(bf(X, (long long)i * ldx + k), bf(Y, (long long)j * ldy + k), s);
      float f = s, fp = 1.f;
      if (mode == 1) {
        const float sig = 1.f / (1.f + expf(-s));
        f = s * sig;
        fp = sig * (1.f + s * (1.f - sig));
      }
      float g = stable ? exp2f(f * scale2 - rs) + exp2f(f * scale2 - (colscale[j] - lg))
                       : exp2f(fmaf(f, scale2, -shift2)) * (rs + colscale[j] * gnorm);
      const bool on_diag = ydiag != 0.f && i + diag_off == j;
      if (on_diag) g -= ydn;
      tsum += (double)g * (double)f;
      if (mode == 1) g *= fp;
      float gb = __bfloat162float(__float2bfloat16_rn(g));
      if (hp) gb += __bfloat162float(__float2bfloat16_rn(g - gb));      // hi + lo: what the two TS-MMAs consume
      if (on_diag && diag_corr) {
        diag_corr[2 * i] = (g - gb) * ign;
        diag_corr[2 * i + 1] = gb * ign;
      }
      const float ge = gb * out_scale * ign;
      for (int d = 0; d < D; ++d) dX[(long long)i * ldd + d] += ge * bf(Y, (long long)j * ldy + hi_off + d);
    }
  }
  if (scal) scal[0] += tsum * (double)ign;
  return 0;
}

// ---------------- SigLIP pieces: shipped CUDA-core kernels + the model of the forward-only dense tile kernel ----------------
extern "C" int b200clip_dyn_set_siglip(float* dyn, float lclamp, float yneg, void*) {
  emul::launch(1, 32, [&] { dyn_set_siglip_kernel(dyn, lclamp, yneg); });
  return 0;
}
extern "C" int b200clip_siglip_dense_fwd(const void* V, const void* T, int B, int Tn, int Kp, int ldv, int ldt, const float* dyn,
                                         double* acc, void*) {
  const float inv_tau = dyn[2], bias = dyn[5], lc = dyn[8], yneg = dyn[9];
  double t = 0.0;
  for (int i = 0; i < B; ++i)
    for (int j = 0; j < Tn; ++j) {
      float s = 0.f;
      for (int k = 0; k < Kp; ++k) s = fmaf(bf(V, (long long)i * ldv + k), bf(T, (long long)j * ldt + k), s);
      const float L = fminf(fmaxf(fmaf(s, inv_tau, bias), -lc), lc);
      t += (double)(fmaf(-yneg, L, fmaxf(L, 0.f) + log1pf(expf(-fabsf(L)))));
    }
  acc[0] += t;
  return 0;
}
extern "C" int b200clip_siglip_combine(const double* acc, double wn_c, const float* tinv, int T, double* red, void*) {
  emul::launch(1, 256, [&] { siglip_combine_kernel(acc, wn_c, tinv, T, red); });
  return 0;
}
extern "C" int b200clip_siglip_loss_out(const double* red, const int* overflow, const float* ent, int world, float* loss_out,
                                        float* diag, void*) {
  emul::launch(1, 32, [&] { siglip_loss_out_kernel(red, overflow, ent, world, loss_out, diag); });
  return 0;
}
extern "C" int b200clip_siglip_scalar_grads(const double* red, const float* dyn, const float* gmul, float* dlt, float* dbias,
                                            void*) {
  emul::launch(1, 32, [&] { siglip_scalar_grads_kernel(red, dyn, gmul, dlt, dbias); });
  return 0;
}
extern "C" int b200clip_siglip_compact(const float* mask, long long ldm, const float* pw, long long ldw, int B, int T, int cap,
                                       int* col, float* y, float* w, int* cnt, float* ysum, int* overflow, void*) {
  if (B <= 0 || T <= 0 || cap <= 0) return -22;
  emul::launch((B + 7) / 8, 256, [&] { siglip_compact_kernel(mask, ldm, pw, ldw, B, T, cap, col, y, w, cnt, ysum, overflow); });
  return 0;
}
extern "C" int b200clip_siglip_pos(const void* V, int ldv, const void* T, int ldt, int K, int Dp, int D, int hi_off, int B, int Tn,
                                   int cap, const int* col, const float* y, const float* w, const int* cnt, const float* ysum,
                                   const float* dyn, float positive_weight, float negative_weight, float c, float gnorm, int hp,
                                   int use_pw, int auto_balance, float* dV, int lddv, float* dT, int lddt, double* acc,
                                   const void* Vraw, int v_dtype, long long ld_vraw, const float* vinv, const void* Traw,
                                   int t_dtype, long long ld_traw, const float* tinv, void*) {
  if (B <= 0 || K <= 0 || (K & 1)) return -22;
  if ((dV == nullptr) != (dT == nullptr)) return -22;
  PosParams p{(const bf16*)V, ldv, (const bf16*)T, ldt, K, Dp, D, hi_off, B, Tn, cap, col, y, w, cnt, ysum, dyn, positive_weight,
              negative_weight, c, gnorm > 0.f ? gnorm : 1.f, hp ? 1 : 0, use_pw, auto_balance, dV, lddv, dT, lddt, acc,
              Vraw, v_dtype, ld_vraw, vinv, Traw, t_dtype, ld_traw, tinv};
  emul::launch((B + 7) / 8, 256, [&] { siglip_pos_kernel(p); });
  return 0;
}

// ---------------- streaming retrieval: shipped epilogue policies / selection kernels + a modelled GEMM ----------------
// The tile engine is replaced by a driver that follows its epilogue contract (tile_engine.cuh: TileSeq item mode, TeCtx,
// begin_outer / chunk x 4 / end_tile / end_outer per thread, zero-filled padding) and computes each similarity from the bf16
// operands with an fp32 accumulator (exact for the exact-grid evaluation embeddings, as on the tensor core).
static const int kSmsR = 148;
static inline float sim_at(const void* V, int ldv, const void* T, int ldt, int K, int i, int j) {
  float s = 0.f;
  for (int k = 0; k < K; ++k) s = fmaf(bf(V, (long long)i * ldv + k), bf(T, (long long)j * ldt + k), s);
  return s;
}
template <class Epi>
static void run_epi(const std::vector<float>& S, int Ma, int Nb, int segs, const typename Epi::Params& p) {
  const int m_tiles = (Ma + 127) / 128, n_blocks = (Nb + 255) / 256;
  for (int m_tile = 0; m_tile < m_tiles; ++m_tile)
    for (int seg = 0; seg < segs; ++seg) {
      const int i0 = (int)((long long)n_blocks * seg / segs), i1 = (int)((long long)n_blocks * (seg + 1) / segs);
      if (i0 >= i1) continue;
      for (int wg = 0; wg < 2; ++wg)
        for (int q = 0; q < 4; ++q)
          for (int lane = 0; lane < 32; ++lane) {
            typename Epi::State st;
            Epi::init(st, p);
            TeCtx ctx;
            ctx.wg = wg; ctx.Nb = Nb; ctx.seg = seg; ctx.m_tile = m_tile;
            ctx.row = m_tile * 128 + q * 32 + lane;
            ctx.row_ok = ctx.row < Ma;
            for (int nb = i0; nb < i1; ++nb) {
              ctx.n_block = nb;
              ctx.col0 = nb * 256 + wg * 128;
              ctx.full = (m_tile * 128 + 128 <= Ma) && (nb * 256 + 256 <= Nb);
              if (nb == i0) Epi::begin_outer(st, p, m_tile, ctx);
              for (int c = 0; c < 4; ++c) {
                uint32_t acc[32];
                for (int e = 0; e < 32; ++e) {
                  const int col = ctx.col0 + c * 32 + e;
                  acc[e] = __float_as_uint((ctx.row_ok && col < Nb) ? S[(size_t)ctx.row * Nb + col] : 0.f);
                }
                Epi::chunk(st, p, ctx, c, acc);
              }
              Epi::end_tile(st, p, ctx);
            }
            Epi::end_outer(st, p, m_tile, ctx);
          }
    }
}
static std::vector<float> sim_matrix(const void* V, const void* T, int N, int M, int K, int ldv, int ldt) {
  std::vector<float> S((size_t)N * M);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < M; ++j) S[(size_t)i * M + j] = sim_at(V, ldv, T, ldt, K, i, j);
  return S;
}

extern "C" int b200clip_retrieval_segments(int Ma, int Nb) {
  const int m_tiles = (Ma + 127) / 128, n_blocks = (Nb + 255) / 256;
  int segs = 1;
  if (m_tiles < 4 * kSmsR) {
    segs = (4 * kSmsR + m_tiles - 1) / m_tiles;
    if (segs > n_blocks) segs = n_blocks;
    if (segs > 64) segs = 64;
    if (segs < 1) segs = 1;
  }
  return segs;
}
extern "C" int b200clip_retrieval_sweep(const void* V, const void* T, int Nv, int Mt, int Kp, int ldv, int ldt, const float* sgt,
                                        const long long* gt, int col_offset, int* counts, int k, int segs, float* part_score,
                                        int* part_idx, void*) {
  if (Nv <= 0 || Mt <= 0 || k < 0 || k > 64 || segs < 1) return -22;
  if (k > 0 && (!part_score || !part_idx)) return -22;
  if (sgt && (!gt || !counts)) return -22;
  const std::vector<float> S = sim_matrix(V, T, Nv, Mt, Kp, ldv, ldt);
  RetrParams p{sgt, gt, col_offset, counts, part_score, part_idx, 2 * segs, k};
  if (k == 0) run_epi<RetrEpi<0>>(S, Nv, Mt, segs, p);
  else if (k <= 16) run_epi<RetrEpi<16>>(S, Nv, Mt, segs, p);
  else run_epi<RetrEpi<64>>(S, Nv, Mt, segs, p);
  return 0;
}
extern "C" int b200clip_retrieval_colmax(const void* V, const void* T, int Nv, int Mt, int Kp, int ldv, int ldt, int segs,
                                         float* part_max, void*) {
  if (Nv <= 0 || Mt <= 0 || segs < 1 || !part_max) return -22;
  ColMaxParams p{part_max, 2 * segs};
  run_epi<ColMaxEpi>(sim_matrix(V, T, Nv, Mt, Kp, ldv, ldt), Nv, Mt, segs, p);
  return 0;
}
extern "C" int b200clip_retrieval_collect(const void* V, const void* T, int Nv, int Mt, int Kp, int ldv, int ldt, const float* thr,
                                          int col_offset, int segs, int* cnt, float* buf_s, int* buf_i, int cap, int* overflow,
                                          void*) {
  if (Nv <= 0 || Mt <= 0 || segs < 1 || cap < 1 || !thr || !cnt || !buf_s || !buf_i || !overflow) return -22;
  CollectParams p{thr, col_offset, cnt, buf_s, buf_i, cap, overflow};
  run_epi<CollectEpi>(sim_matrix(V, T, Nv, Mt, Kp, ldv, ldt), Nv, Mt, segs, p);
  return 0;
}
extern "C" int b200clip_kth_largest(const float* vals, int rows, int cand, int k, float* thr, void*) {
  if (rows <= 0 || cand <= 0 || k <= 0 || !vals || !thr) return -22;
  emul::launch((rows + 7) / 8, 256, [&] { kth_largest_kernel(vals, rows, cand, k, thr); });
  return 0;
}
extern "C" int b200clip_topk_merge(const float* ps, const int* pi, int rows, int cand, int k, float* out_s, long long* out_i,
                                   void*) {
  if (rows <= 0 || cand <= 0 || k <= 0) return -22;
  emul::launch((rows + 7) / 8, 256, [&] { topk_merge_kernel(ps, pi, rows, cand, k, out_s, out_i); });
  return 0;
}
extern "C" int b200clip_recall_hits(const int* counts, int rows, const int* kvals, int nk, unsigned long long* hits, void*) {
  if (rows <= 0 || nk <= 0) return -22;
  emul::launch((rows + 255) / 256, 256, [&] { recall_hits_kernel(counts, rows, kvals, nk, hits); });
  return 0;
}
extern "C" int b200clip_mrr_from_counts(const int* counts, int rows, int n_bins, int* hist, double* out, void*) {
  if (rows <= 0 || n_bins <= 0) return -22;
  emul::launch((rows + 255) / 256, 256, [&] { rank_hist_kernel(counts, rows, n_bins, hist); });
  emul::launch(1, 1024, [&] { rank_hist_mrr_kernel(hist, n_bins, out); });
  return 0;
}
// operand helpers of l2norm.cu (shipped kernels) and the tensor-core ground-truth dot (model: same fp32 dot as sim_at)
extern "C" int b200clip_inexact_bf16(const void* x, int dtype, long long ld, int rows, int dim, int* flag, void*) {
  if (rows <= 0 || dim <= 0) return -22;
  long long blocks = ((long long)rows * dim + 255) / 256;
  if (blocks > 32) blocks = 32;
  if (dtype == 0) emul::launch((unsigned)blocks, 256, [&] { inexact_bf16_kernel<float>((const float*)x, ld, rows, dim, flag); });
  else if (dtype == 2) emul::launch((unsigned)blocks, 256, [&] { inexact_bf16_kernel<__half>((const __half*)x, ld, rows, dim, flag); });
  else return -22;
  return 0;
}
extern "C" int b200clip_gather_rows_bf16(const void* src, int lds, const long long* idx, int rows, int src_rows, int K, void* dst,
                                         int ldd, void*) {
  if (rows <= 0 || K <= 0 || (K & 7) || (lds & 7) || (ldd & 7) || !src || !idx || !dst) return -22;
  emul::launch((rows + 7) / 8, 256, [&] {
    gather_rows_bf16_kernel((const uint4*)src, lds / 8, idx, rows, src_rows, K / 8, (uint4*)dst, ldd / 8);
  });
  return 0;
}
extern "C" int b200clip_rowdot_bf16(const void* a, int lda, const void* b, int ldb, const long long* idx, int rows, int b_rows,
                                    int K, float* out, void*) {
  if (rows <= 0 || K <= 0 || (K & 1) || !a || !b || !out) return -22;
  emul::launch((rows + 7) / 8, 256, [&] { rowdot_bf16_kernel((const bf16*)a, lda, (const bf16*)b, ldb, idx, rows, b_rows, K, out); });
  return 0;
}
extern "C" int b200clip_rowdot_tc(const void* a, int lda, const void* b, int ldb, int rows, int Kp, float* out, void*) {
  for (int i = 0; i < rows; ++i) out[i] = sim_at(a, lda, b, ldb, Kp, i, i);
  return 0;
}

extern "C" int b200clip_alignment_diag(const float* sums, int n, const float* dyn, int gated, float* out, void*) {
  if (!sums || !dyn || !out || n <= 0) return -22;
  emul::launch(1, 1024, [&] { alignment_diag_kernel(sums, n, dyn, gated, out); });
  return 0;
}

// ---------------- entropy regulariser (contrastive.py:19-68): models of the two row-statistics sweeps + shipped kernels ----------------
extern "C" int b200clip_siglip_entropy_rowsum(const void* V, const void* T, int B, int Tn, int Kp, int ldv, int ldt,
                                              const float* dyn, float* Z, void*) {
  const float inv_tau = dyn[2], bias = dyn[5];
  for (int i = 0; i < B; ++i)
    for (int j = 0; j < Tn; ++j) {
      float s = 0.f;
      for (int k = 0; k < Kp; ++k) s = fmaf(bf(V, (long long)i * ldv + k), bf(T, (long long)j * ldt + k), s);
      Z[i] += expf(fminf(fmaxf(fmaf(s, inv_tau, bias), -30.f), 30.f) - 30.f);
    }
  return 0;
}
extern "C" int b200clip_siglip_entropy_stats(const void* V, const void* T, int B, int Tn, int Kp, int ldv, int ldt,
                                             const float* dyn, const float* Z, float* H, float* Q, void*) {
  const float inv_tau = dyn[2], bias = dyn[5];
  for (int i = 0; i < B; ++i) {
    const float iz = 1.f / Z[i];
    for (int j = 0; j < Tn; ++j) {
      float s = 0.f;
      for (int k = 0; k < Kp; ++k) s = fmaf(bf(V, (long long)i * ldv + k), bf(T, (long long)j * ldt + k), s);
      const float pij = expf(fminf(fmaxf(fmaf(s, inv_tau, bias), -30.f), 30.f) - 30.f) * iz;
      const float pe = pij + 1e-10f;
      H[i] -= pij * logf(pe);
      Q[i] += pij * (pij / pe);
    }
  }
  return 0;
}
extern "C" int b200clip_siglip_entropy_rows(const float* Z, const float* H, const float* Q, int B, float* rowvec, double* stats,
                                            void*) {
  if (B <= 0) return -22;
  emul::launch(1, 1024, [&] { siglip_entropy_rows_kernel(Z, H, Q, B, rowvec, stats); });
  return 0;
}
extern "C" int b200clip_siglip_entropy_coef(const double* stats_all, int W, int Bg, int T, float weight, float thr, float* dyn,
                                            float* out, void*) {
  if (W <= 0 || Bg <= 0 || T <= 0) return -22;
  emul::launch(1, 32, [&] { siglip_entropy_coef_kernel(stats_all, W, Bg, T, weight, thr, dyn, out); });
  return 0;
}
