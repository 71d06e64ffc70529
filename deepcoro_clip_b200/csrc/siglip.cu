// SigLIP multi-positive sigmoid loss (utils/loss/contrastive.py:230-315), split into
//   dense part   : every pair treated as a NEGATIVE with weight negative_weight — logits_bwd mode 2 (training:
//                  loss + gradients in one recompute pass) or the forward-only SoftplusEpi below (no_grad);
//   sparse part  : the <= cap positives per video row (pos_mask > 0) — compacted once per call into per-row lists
//                  and applied as exact fp32 corrections (loss, dVhat, dThat, dbias, dlog_temp) by siglip_pos.
// The dense [B, T] fp32 pos_mask / pos_weights are therefore read exactly once (one streaming pass), never per tile.
#include "tile_engine.cuh"
#include "host_api.h"

namespace b2 {

// ---------------------------------------------------------------------------------------------------------------
// forward-only dense term: acc += sum_ij [softplus(L_ij) - yneg L_ij], L = clamp(S_ij/tau + bias, -lclamp, lclamp)
// ---------------------------------------------------------------------------------------------------------------
struct SpParams {
  const float* dyn;     // [2] = 1/tau, [5] = bias
  double* acc;
};
struct SoftplusEpi {
  using Params = SpParams;
  struct State {
    double total;
    float inv_tau, bias, lc, yneg;
  };
  __device__ static __forceinline__ void init(State& st, const Params& p) {
    st.total = 0.0;
    st.inv_tau = p.dyn[2];
    st.bias = p.dyn[5];
    st.lc = p.dyn[8];
    st.yneg = p.dyn[9];
  }
  __device__ static __forceinline__ void begin_outer(State&, const Params&, int, const TeCtx&) {}
  __device__ static __forceinline__ void chunk(State& st, const Params&, const TeCtx& ctx, int c,
                                               const uint32_t (&acc)[32]) {
    float part = 0.f;
    const int nvalid = ctx.row_ok ? ctx.Nb - (ctx.col0 + c * 32) : 0;
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const float R = fmaf(__uint_as_float(acc[e]), st.inv_tau, st.bias);
      const float L = fminf(fmaxf(R, -st.lc), st.lc);
      const float ex = ex2_approx(-1.4426950408889634f * fabsf(L));
      const float sp = fmaf(-st.yneg, L, fmaxf(L, 0.f) + log1p_ex(ex));
      part += (ctx.full || e < nvalid) ? sp : 0.f;
    }
    st.total += (double)part;
  }
  __device__ static __forceinline__ void end_tile(State&, const Params&, const TeCtx&) {}
  __device__ static __forceinline__ void end_outer(State& st, const Params& p, int, const TeCtx&) {
    double v = st.total;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(p.acc, v);
    st.total = 0.0;
  }
};


// ---------------------------------------------------------------------------------------------------------------
// entropy regulariser (utils/loss/contrastive.py:19-68), forward row statistics over the clamped logits L:
//   pass 1 (EntSumEpi) : Z_i = sum_j exp(L_ij - 30)                     (L <= 30, so every term is in [e^-60, 1])
//   pass 2 (EntStatEpi): p_ij = exp(L_ij - 30) / Z_i ; H_i = -sum_j p ln(p + 1e-10) ; Q_i = sum_j p^2 / (p + 1e-10)
// (m_i = sum_j p_ij h_ij = H_i - Q_i is what the backward needs.) One fp32 atomic per thread per tile and row.
// ---------------------------------------------------------------------------------------------------------------
struct EntParams {
  const float* dyn;
  const float* Z;    // pass 2: [Ma] row sums of pass 1
  float* out0;       // pass 1: Z ; pass 2: H
  float* out1;       // pass 2: Q
};
template <bool kStats>
struct EntEpi {
  using Params = EntParams;
  struct State {
    float inv_tau, bias, a0, a1, iz;
  };
  __device__ static __forceinline__ void init(State& st, const Params& p) {
    st.inv_tau = p.dyn[2];
    st.bias = p.dyn[5];
    st.a0 = st.a1 = 0.f;
    st.iz = 0.f;
  }
  __device__ static __forceinline__ void begin_outer(State&, const Params&, int, const TeCtx&) {}
  __device__ static __forceinline__ void chunk(State& st, const Params& p, const TeCtx& ctx, int c,
                                               const uint32_t (&acc)[32]) {
    if (kStats && c == 0) st.iz = ctx.row_ok ? 1.f / p.Z[ctx.row] : 0.f;
    const int nvalid = ctx.row_ok ? ctx.Nb - (ctx.col0 + c * 32) : 0;
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const float R = fmaf(__uint_as_float(acc[e]), st.inv_tau, st.bias);
      const float L = fminf(fmaxf(R, -30.f), 30.f);
      const float ex = ex2_approx((L - 30.f) * 1.4426950408889634f);
      const bool ok = ctx.full || e < nvalid;
      if (!kStats) {
        st.a0 += ok ? ex : 0.f;
      } else {
        const float pij = ex * st.iz;
        const float pe = pij + 1e-10f;
        st.a0 -= ok ? pij * 0.6931471805599453f * lg2_approx(pe) : 0.f;
        st.a1 += ok ? pij * __fdividef(pij, pe) : 0.f;
      }
    }
  }
  __device__ static __forceinline__ void end_tile(State& st, const Params& p, const TeCtx& ctx) {
    if (ctx.row_ok) {
      atomicAdd(p.out0 + ctx.row, st.a0);
      if (kStats) atomicAdd(p.out1 + ctx.row, st.a1);
    }
    st.a0 = st.a1 = 0.f;
  }
  __device__ static __forceinline__ void end_outer(State&, const Params&, int, const TeCtx&) {}
};

// ---------------------------------------------------------------------------------------------------------------
// scalar tails (one thread each; they replace ~20 single-element PyTorch launches per step):
//   siglip_combine : red = {loss, dbias, sum G*s} from the dense sums acc[0..2] and the positive corrections acc[4..6]
//   siglip_loss_out: loss_out = red[0] (+ NaN when a row overflowed its positive list) (+ entropy penalty ent[5]);
//                    diag (optional, 7 floats) = {ent[0..5], bce loss} for get_entropy_diagnostics()
//   siglip_scalar_grads: dlog_temp = -red[2] / tau * [tau not clamped] * grad_out ; dbias = red[1] * grad_out
// ---------------------------------------------------------------------------------------------------------------
}  // namespace b2

#include "siglip_kernels.cuh"

namespace b2 {

}  // namespace b2

namespace b2host {
using namespace b2;
int make_shape(TeShape& g, int Ma, int Nb, int Kp);   // logits_fwd.cu

int siglip_dense_fwd(const void* V, const void* T, int B, int Tn, int Kp, int ldv, int ldt, const float* dyn,
                     double* acc, cudaStream_t stream) {
  TeShape g;
  int rc = make_shape(g, B, Tn, Kp);
  if (rc) return rc;
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap_bf16_2d(&tmA, V, B, Kp, ldv, TE_BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmB, T, Tn, Kp, ldt, TE_BN))) return rc;
  auto kern = te_kernel<SoftplusEpi, true>;
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device() & 63];   // cudaFuncSetAttribute is per device
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TE_SMEM_BYTES) != cudaSuccess)
      return B2_ECUDA;
    attr_done = true;
  }
  const long long total = (long long)g.m_tiles * g.n_blocks;
  int grid = sm_count();
  if (total < grid) grid = (int)total;
  SpParams p{dyn, acc};
  kern<<<grid, TE_THREADS, TE_SMEM_BYTES, stream>>>(tmA, tmB, g, p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

template <bool kStats>
static int launch_ent(const void* V, const void* T, int B, int Tn, int Kp, int ldv, int ldt, const EntParams& p,
                      cudaStream_t stream) {
  TeShape g;
  int rc = make_shape(g, B, Tn, Kp);
  if (rc) return rc;
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap_bf16_2d(&tmA, V, B, Kp, ldv, TE_BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmB, T, Tn, Kp, ldt, TE_BN))) return rc;
  auto kern = te_kernel<EntEpi<kStats>, false>;
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device() & 63];   // cudaFuncSetAttribute is per device
  if (!attr_done) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TE_SMEM_BYTES) != cudaSuccess)
      return B2_ECUDA;
    attr_done = true;
  }
  const long long total = (long long)g.m_tiles * g.n_blocks;
  int grid = sm_count();
  if (total < grid) grid = (int)total;
  kern<<<grid, TE_THREADS, TE_SMEM_BYTES, stream>>>(tmA, tmB, g, p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int siglip_entropy_rowsum(const void* V, const void* T, int B, int Tn, int Kp, int ldv, int ldt, const float* dyn,
                          float* Z, cudaStream_t stream) {
  EntParams p{dyn, nullptr, Z, nullptr};
  return launch_ent<false>(V, T, B, Tn, Kp, ldv, ldt, p, stream);
}

int siglip_entropy_stats(const void* V, const void* T, int B, int Tn, int Kp, int ldv, int ldt, const float* dyn,
                         const float* Z, float* H, float* Q, cudaStream_t stream) {
  EntParams p{dyn, Z, H, Q};
  return launch_ent<true>(V, T, B, Tn, Kp, ldv, ldt, p, stream);
}

int siglip_entropy_rows(const float* Z, const float* H, const float* Q, int B, float* rowvec, double* stats,
                        cudaStream_t s) {
  if (B <= 0) return B2_EINVAL;
  siglip_entropy_rows_kernel<<<1, 1024, 0, s>>>(Z, H, Q, B, rowvec, stats);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int siglip_entropy_coef(const double* stats_all, int W, int Bg, int T, float weight, float thr, float* dyn, float* out,
                        cudaStream_t s) {
  if (W <= 0 || Bg <= 0 || T <= 0) return B2_EINVAL;
  siglip_entropy_coef_kernel<<<1, 32, 0, s>>>(stats_all, W, Bg, T, weight, thr, dyn, out);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int siglip_combine(const double* acc, double wn_c, const float* tinv, int T, double* red, cudaStream_t s) {
  siglip_combine_kernel<<<1, 256, 0, s>>>(acc, wn_c, tinv, T, red);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}
int siglip_loss_out(const double* red, const int* overflow, const float* ent, int world, float* loss_out, float* diag,
                    cudaStream_t s) {
  siglip_loss_out_kernel<<<1, 32, 0, s>>>(red, overflow, ent, world, loss_out, diag);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}
int siglip_scalar_grads(const double* red, const float* dyn, const float* gmul, float* dlt, float* dbias,
                        cudaStream_t s) {
  siglip_scalar_grads_kernel<<<1, 32, 0, s>>>(red, dyn, gmul, dlt, dbias);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
namespace b2 {
__global__ void siglip_compact_vec_kernel(const float* __restrict__ mask, long ldm, const float* __restrict__ pw, long ldw, int B,
                                          int T, int cap, int* __restrict__ col, float* __restrict__ y, float* __restrict__ w,
                                          int* __restrict__ cnt, float* __restrict__ ysum, int* __restrict__ overflow);
}
namespace b2host {
using namespace b2;

int siglip_compact(const float* mask, long ldm, const float* pw, long ldw, int B, int T, int cap, int* col, float* y,
                   float* w, int* cnt, float* ysum, int* overflow, cudaStream_t s) {
  if (B <= 0 || T <= 0 || cap <= 0) return B2_EINVAL;
  if (mask && T % 4 == 0 && ldm % 4 == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0)
    siglip_compact_vec_kernel<<<(B + 7) / 8, 256, 0, s>>>(mask, ldm, pw, ldw, B, T, cap, col, y, w, cnt, ysum, overflow);
  else
    siglip_compact_kernel<<<(B + 7) / 8, 256, 0, s>>>(mask, ldm, pw, ldw, B, T, cap, col, y, w, cnt, ysum, overflow);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host

namespace b2 {

// siglip_pos_kernel (siglip_kernels.cuh) for the common layout — fp32 raw features, D % 256 == 0, 16-byte aligned rows: the
// scalar kernel walks every positive with 2-byte / 4-byte loads and one atomic per element (186 us for the 32 K positives of
// config 2, the longest kernel of that step). Here a lane owns 8 contiguous channels per 256-channel chunk: the row's own
// operand / raw values stay in registers across its positives, the partner rows come in as 16-byte loads issued together,
// dV is accumulated in registers and written once, dT goes out as red.global.add.v4.f32. Same arithmetic, same order of
// the per-pair scalar work.
// siglip_compact_kernel (siglip_kernels.cuh) with 16-byte loads, four of them in flight per lane: the scalar kernel streams the
// dense fp32 mask at 2 TB/s (one 4-byte load per lane and ballot step). Same lists in the same (column) order.
__global__ void __launch_bounds__(256) siglip_compact_vec_kernel(const float* __restrict__ mask, long ldm,
                                                                  const float* __restrict__ pw, long ldw, int B, int T, int cap,
                                                                  int* __restrict__ col, float* __restrict__ y,
                                                                  float* __restrict__ w, int* __restrict__ cnt,
                                                                  float* __restrict__ ysum, int* __restrict__ overflow) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* mr = mask + (size_t)row * ldm;
  const float* wr = pw ? pw + (size_t)row * ldw : nullptr;
  const unsigned lt = (1u << lane) - 1;
  int n = 0;
  float ys = 0.f;
  auto emit = [&](const float4& v4, int cbase) {          // the 128 columns cbase .. cbase + 127, lane owns cbase + 4 lane + q
    const float v[4] = {fminf(fmaxf(v4.x, 0.f), 1.f), fminf(fmaxf(v4.y, 0.f), 1.f), fminf(fmaxf(v4.z, 0.f), 1.f),
                        fminf(fmaxf(v4.w, 0.f), 1.f)};
    ys += (v[0] + v[1]) + (v[2] + v[3]);
    const bool any = v[0] > 0.f || v[1] > 0.f || v[2] > 0.f || v[3] > 0.f;
    if (__ballot_sync(0xffffffffu, any) == 0u) return;
    int before = 0, total = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const unsigned b = __ballot_sync(0xffffffffu, v[q] > 0.f);
      before += __popc(b & lt);
      total += __popc(b);
    }
    int slot = n + before;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (v[q] > 0.f) {
        const int c = cbase + 4 * lane + q;
        if (slot < cap) {
          col[(size_t)row * cap + slot] = c;
          y[(size_t)row * cap + slot] = v[q];
          w[(size_t)row * cap + slot] = wr ? wr[c] : 1.f;
        }
        ++slot;
      }
    n += total;
  };
  int c0 = 0;
  for (; c0 + 512 <= T; c0 += 512) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(mr + c0 + u * 128 + lane * 4);
#pragma unroll
    for (int u = 0; u < 4; ++u) emit(v[u], c0 + u * 128);
  }
  for (; c0 < T; c0 += 128) {                             // tail: T % 4 == 0 is guaranteed by the caller, not T % 128
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 + lane * 4 < T) v = *reinterpret_cast<const float4*>(mr + c0 + lane * 4);
    emit(v, c0);
  }
  ys = warp_sum(ys);
  if (lane == 0) {
    cnt[row] = n < cap ? n : cap;
    ysum[row] = ys;
    if (n > cap) atomicExch(overflow, 1);
  }
}

template <int NC>
__global__ void __launch_bounds__(256) siglip_pos_vec_kernel(PosParams p) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  double a_loss = 0.0, a_bias = 0.0, a_t = 0.0;
  if (row < p.B) {
    const float inv_tau = p.dyn[2], bias = p.dyn[5], lc = p.dyn[8], yneg = p.dyn[9];
    const bool use_pw = (p.use_pw & 1) != 0, rule_mask = (p.use_pw & 2) != 0;
    const int n = p.cnt[row];
    constexpr int nc = NC;                             // 256-channel chunks
    float ratio = 1.f;
    if (p.auto_balance) {
      const float pc = fmaxf(p.ysum[row], 1.f);
      ratio = fmaxf(((float)p.Tn - pc) / pc, 1.f);
    }
    const __nv_bfloat16* vr = p.V + (size_t)row * p.ldv;
    const float* vraw = static_cast<const float*>(p.Vraw) + (long long)row * p.ld_vraw;
    const float vi = p.vinv[row];
    float vx[NC][8], vh[NC][8], dv[NC][8];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c < nc) {
        const int d0 = (lane + 32 * c) * 8;
        const float4 a = *reinterpret_cast<const float4*>(vraw + d0), b = *reinterpret_cast<const float4*>(vraw + d0 + 4);
        vx[c][0] = a.x * vi; vx[c][1] = a.y * vi; vx[c][2] = a.z * vi; vx[c][3] = a.w * vi;
        vx[c][4] = b.x * vi; vx[c][5] = b.y * vi; vx[c][6] = b.z * vi; vx[c][7] = b.w * vi;
        const uint4 h = *reinterpret_cast<const uint4*>(vr + p.hi_off + d0);
        const __nv_bfloat162* hp2 = reinterpret_cast<const __nv_bfloat162*>(&h);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(hp2[q]); vh[c][2 * q] = f.x; vh[c][2 * q + 1] = f.y; }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) dv[c][q] = 0.f;
    }
    for (int e = 0; e < n; ++e) {
      const int j = p.col[(size_t)row * p.cap + e];
      const float yraw = p.y[(size_t)row * p.cap + e];
      const float yv = fmaf(yraw, 1.f - 2.f * yneg, yneg);
      const float pwv = p.w[(size_t)row * p.cap + e];
      const __nv_bfloat16* tr = p.T + (size_t)j * p.ldt;
      const float* traw = static_cast<const float*>(p.Traw) + (long long)j * p.ld_traw;
      const float ti = p.tinv[j];
      // partner row: raw (fp32) and hi operand (bf16), all loads first
      float tx[NC][8], th[NC][8];
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c < nc) {
          const int d0 = (lane + 32 * c) * 8;
          const float4 a = *reinterpret_cast<const float4*>(traw + d0), b = *reinterpret_cast<const float4*>(traw + d0 + 4);
          const uint4 h = *reinterpret_cast<const uint4*>(tr + p.hi_off + d0);
          tx[c][0] = a.x * ti; tx[c][1] = a.y * ti; tx[c][2] = a.z * ti; tx[c][3] = a.w * ti;
          tx[c][4] = b.x * ti; tx[c][5] = b.y * ti; tx[c][6] = b.z * ti; tx[c][7] = b.w * ti;
          const __nv_bfloat162* hp2 = reinterpret_cast<const __nv_bfloat162*>(&h);
#pragma unroll
          for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(hp2[q]); th[c][2 * q] = f.x; th[c][2 * q + 1] = f.y; }
        }
      // what the dense pass saw: the dot product of the bf16 operands over all K columns (all panels of bf16x3 operands)
      float s = 0.f;
      for (int k = lane * 8; k < p.K; k += 256) {
        const uint4 a = *reinterpret_cast<const uint4*>(vr + k), b = *reinterpret_cast<const uint4*>(tr + k);
        const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 fa = __bfloat1622float2(a2[q]), fb = __bfloat1622float2(b2[q]);
          s = fmaf(fa.x, fb.x, s);
          s = fmaf(fa.y, fb.y, s);
        }
      }
      float sx = 0.f;
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c < nc)
#pragma unroll
          for (int q = 0; q < 8; ++q) sx = fmaf(vx[c][q], tx[c][q], sx);
      s = warp_sum(s);
      sx = warp_sum(sx);
      const float R = fmaf(s, inv_tau, bias);
      const float L = fminf(fmaxf(R, -lc), lc);
      const float ex = __expf(-fabsf(L));
      const float sp = fmaxf(L, 0.f) + log1pf(ex);
      const float sig = L >= 0.f ? 1.f / (1.f + ex) : ex / (1.f + ex);
      const float inr = fabsf(R) <= lc ? 1.f : 0.f;
      const float Rx = fmaf(sx, inv_tau, bias);
      const float Lx = fminf(fmaxf(Rx, -lc), lc);
      const float exx = __expf(-fabsf(Lx));
      const float spx = fmaxf(Lx, 0.f) + log1pf(exx);
      const float sigx = Lx >= 0.f ? 1.f / (1.f + exx) : exx / (1.f + exx);
      const float inrx = fabsf(Rx) <= lc ? 1.f : 0.f;
      float w = p.negative_weight;
      if (rule_mask ? yraw > 0.f : yv > 0.5f)
        w = p.auto_balance ? ratio : (use_pw ? pwv * p.positive_weight : p.positive_weight);
      const float g_full = w * (sigx - yv) * inrx * p.c;
      const float g_dense = p.negative_weight * p.c * (sig - yneg) * inr;
      if (lane == 0) {
        a_loss += (double)((w * (spx - Lx * yv) - p.negative_weight * (sp - L * yneg)) * p.c);
        a_bias += (double)(g_full - g_dense);
        a_t += (double)(g_full * sx - g_dense * s);
      }
      if (p.dV) {
        const float gs = g_dense * p.gnorm;
        float gr = __bfloat162float(__float2bfloat16_rn(gs));
        if (p.hp) gr += __bfloat162float(__float2bfloat16_rn(gs - gr));
        const float dgf = g_full * inv_tau, dgd = gr / p.gnorm * inv_tau;
        float* dt = p.dT + (size_t)j * p.lddt;
#pragma unroll
        for (int c = 0; c < NC; ++c)
          if (c < nc) {
            const int d0 = (lane + 32 * c) * 8;
            float o[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              dv[c][q] += dgf * tx[c][q] - dgd * th[c][q];
              o[q] = dgf * vx[c][q] - dgd * vh[c][q];
            }
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dt + d0), "f"(o[0]), "f"(o[1]), "f"(o[2]), "f"(o[3]) : "memory");
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dt + d0 + 4), "f"(o[4]), "f"(o[5]), "f"(o[6]), "f"(o[7]) : "memory");
          }
      }
    }
    if (p.dV && n > 0) {
      float* dvr = p.dV + (size_t)row * p.lddv;                // this warp owns row i
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c < nc) {
          const int d0 = (lane + 32 * c) * 8;
          float4 a = *reinterpret_cast<float4*>(dvr + d0), b = *reinterpret_cast<float4*>(dvr + d0 + 4);
          a.x += dv[c][0]; a.y += dv[c][1]; a.z += dv[c][2]; a.w += dv[c][3];
          b.x += dv[c][4]; b.y += dv[c][5]; b.z += dv[c][6]; b.w += dv[c][7];
          *reinterpret_cast<float4*>(dvr + d0) = a;
          *reinterpret_cast<float4*>(dvr + d0 + 4) = b;
        }
    }
  }
  __shared__ double sh[3][8];
  if (lane == 0) {
    sh[0][threadIdx.x >> 5] = a_loss;
    sh[1][threadIdx.x >> 5] = a_bias;
    sh[2][threadIdx.x >> 5] = a_t;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
    if (t != 0.0) atomicAdd(p.acc + threadIdx.x, t);
  }
}

}  // namespace b2

namespace b2host {
using namespace b2;

int siglip_pos(const void* V, int ldv, const void* T, int ldt, int K, int Dp, int D, int hi_off, int B, int Tn, int cap,
               const int* col, const float* y, const float* w, const int* cnt, const float* ysum, const float* dyn,
               float positive_weight, float negative_weight, float c, float gnorm, int hp, int use_pw, int auto_balance,
               float* dV, int lddv, float* dT, int lddt, double* acc, const void* Vraw, int v_dtype, long long ld_vraw,
               const float* vinv, const void* Traw, int t_dtype, long long ld_traw, const float* tinv, cudaStream_t s) {
  if (B <= 0 || K <= 0 || (K & 1)) return B2_EINVAL;
  if ((dV == nullptr) != (dT == nullptr)) return B2_EINVAL;
  if ((Vraw && (!vinv || v_dtype < 0 || v_dtype > 2)) || (Traw && (!tinv || t_dtype < 0 || t_dtype > 2))) return B2_EINVAL;
  PosParams p{(const __nv_bfloat16*)V, ldv, (const __nv_bfloat16*)T, ldt, K, Dp, D, hi_off, B, Tn, cap, col, y, w, cnt,
              ysum, dyn, positive_weight, negative_weight, c, gnorm > 0.f ? gnorm : 1.f, hp ? 1 : 0, use_pw, auto_balance, dV, lddv, dT,
              lddt, acc, Vraw, v_dtype, ld_vraw, vinv, Traw, t_dtype, ld_traw, tinv};
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool vec = Vraw && Traw && v_dtype == 0 && t_dtype == 0 && D % 256 == 0 && D <= 1024 && K % 256 == 0 && hi_off % 8 == 0 &&
                   ldv % 8 == 0 && ldt % 8 == 0 && ld_vraw % 4 == 0 && ld_traw % 4 == 0 && al16(V) && al16(T) && al16(Vraw) &&
                   al16(Traw) && (!dV || (al16(dV) && al16(dT) && lddv % 4 == 0 && lddt % 4 == 0));
  if (vec) {
    switch (D / 256) {
      case 1: siglip_pos_vec_kernel<1><<<(B + 7) / 8, 256, 0, s>>>(p); break;
      case 2: siglip_pos_vec_kernel<2><<<(B + 7) / 8, 256, 0, s>>>(p); break;
      case 3: siglip_pos_vec_kernel<3><<<(B + 7) / 8, 256, 0, s>>>(p); break;
      default: siglip_pos_vec_kernel<4><<<(B + 7) / 8, 256, 0, s>>>(p); break;
    }
  } else {
    siglip_pos_kernel<<<(B + 7) / 8, 256, 0, s>>>(p);
  }
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
