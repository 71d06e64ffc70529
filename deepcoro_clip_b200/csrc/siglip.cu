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
// forward-only dense term: acc += sum_ij softplus(clamp(S_ij/tau + bias, -30, 30))
// ---------------------------------------------------------------------------------------------------------------
struct SpParams {
  const float* dyn;     // [2] = 1/tau, [5] = bias
  double* acc;
};
struct SoftplusEpi {
  using Params = SpParams;
  struct State {
    double total;
    float inv_tau, bias;
  };
  __device__ static __forceinline__ void init(State& st, const Params& p) {
    st.total = 0.0;
    st.inv_tau = p.dyn[2];
    st.bias = p.dyn[5];
  }
  __device__ static __forceinline__ void begin_outer(State&, const Params&, int, const TeCtx&) {}
  __device__ static __forceinline__ void chunk(State& st, const Params&, const TeCtx& ctx, int c,
                                               const uint32_t (&acc)[32]) {
    float part = 0.f;
    const int nvalid = ctx.row_ok ? ctx.Nb - (ctx.col0 + c * 32) : 0;
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      const float R = fmaf(__uint_as_float(acc[e]), st.inv_tau, st.bias);
      const float L = fminf(fmaxf(R, -30.f), 30.f);
      const float ex = ex2_approx(-1.4426950408889634f * fabsf(L));
      const float sp = fmaxf(L, 0.f) + 0.6931471805599453f * lg2_approx(1.f + ex);
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
// positive-list compaction: one warp per video row scans pos_mask[row, :T] (and pos_weights) once.
//   lists: col[row][cap] int32, y[row][cap] (= clamp(mask,0,1)), pw[row][cap] (raw pos_weights or 1)
//   cnt[row] = number of entries, ysum[row] = sum_j y_ij (auto_balance), overflow flag if a row has > cap positives
// mask == nullptr: diagonal targets (row i -> column i for i < min(B, T)), contrastive.py:274-278.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
siglip_compact_kernel(const float* __restrict__ mask, long ldm, const float* __restrict__ pw, long ldw, int B, int T,
                      int cap, int* __restrict__ col, float* __restrict__ y, float* __restrict__ w,
                      int* __restrict__ cnt, float* __restrict__ ysum, int* __restrict__ overflow) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  if (!mask) {
    if (lane == 0) {
      const bool has = row < T;
      cnt[row] = has ? 1 : 0;
      ysum[row] = has ? 1.f : 0.f;
      if (has) {
        col[(size_t)row * cap] = row;
        y[(size_t)row * cap] = 1.f;
        w[(size_t)row * cap] = 1.f;
      }
    }
    return;
  }
  const float* mr = mask + (size_t)row * ldm;
  const float* wr = pw ? pw + (size_t)row * ldw : nullptr;
  int n = 0;
  float ys = 0.f;
  for (int c0 = 0; c0 < T; c0 += 32) {
    const int c = c0 + lane;
    float v = c < T ? mr[c] : 0.f;
    v = fminf(fmaxf(v, 0.f), 1.f);
    const bool pos = v > 0.f;
    const unsigned b = __ballot_sync(0xffffffffu, pos);
    if (pos) {
      const int slot = n + __popc(b & ((1u << lane) - 1));
      if (slot < cap) {
        col[(size_t)row * cap + slot] = c;
        y[(size_t)row * cap + slot] = v;
        w[(size_t)row * cap + slot] = wr ? wr[c] : 1.f;
      }
    }
    n += __popc(b);
    ys += v;
  }
  ys = warp_sum(ys);
  if (lane == 0) {
    cnt[row] = n < cap ? n : cap;
    ysum[row] = ys;
    if (n > cap) atomicExch(overflow, 1);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// sparse corrections. One warp per video row; for each list entry (row i, column j):
//   s = vhat_i . that_j (bf16 operands, all K panels), R = s/tau + b, L = clamp(R), sigma, softplus
//   weight rule (contrastive.py:283-298): y > 0.5 -> positive weight (severity * positive_weight | positive_weight |
//   auto-balance ratio), else negative_weight.
//   loss   += w (sp - L y) - wn sp                                   (dense part already added wn * sp)
//   dG      = [w (sigma - y) - rounded(wn sigma)] * inr * c             (dense MMA used the bf16-rounded gradient)
//   dVhat_i += dG/tau * that_j(hi) ; dThat_j += dG/tau * vhat_i(hi)  (atomic: several rows may share a text)
//   scalar sums use the unrounded dense value: dbias += [w(sigma-y) - wn sigma] inr c ; tsum += (same) * s
// acc (double): [0] loss correction (already * c), [1] dbias correction, [2] sum dG_fp32 * s correction
// ---------------------------------------------------------------------------------------------------------------
struct PosParams {
  const __nv_bfloat16* V; int ldv;
  const __nv_bfloat16* T; int ldt;
  int K, Dp, D, hi_off;
  int B, Tn, cap;
  const int* col; const float* y; const float* w; const int* cnt; const float* ysum;
  const float* dyn;
  float positive_weight, negative_weight, c, gnorm;
  int hp;
  int use_pw, auto_balance;
  float* dV; int lddv;     // may be null (loss only)
  float* dT; int lddt;
  double* acc;
};

__global__ void __launch_bounds__(256) siglip_pos_kernel(PosParams p) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  double a_loss = 0.0, a_bias = 0.0, a_t = 0.0;
  if (row < p.B) {
    const float inv_tau = p.dyn[2], bias = p.dyn[5];
    const int n = p.cnt[row];
    float ratio = 1.f;
    if (p.auto_balance) {
      const float pc = fmaxf(p.ysum[row], 1.f);
      ratio = fmaxf(((float)p.Tn - pc) / pc, 1.f);
    }
    const __nv_bfloat16* vr = p.V + (size_t)row * p.ldv;
    for (int e = 0; e < n; ++e) {
      const int j = p.col[(size_t)row * p.cap + e];
      const float yv = p.y[(size_t)row * p.cap + e];
      const float pwv = p.w[(size_t)row * p.cap + e];
      const __nv_bfloat16* tr = p.T + (size_t)j * p.ldt;
      float s = 0.f;
      for (int k = lane * 2; k < p.K; k += 64) {
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(vr + k));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(tr + k));
        s = fmaf(a.x, b.x, s);
        s = fmaf(a.y, b.y, s);
      }
      s = warp_sum(s);
      const float R = fmaf(s, inv_tau, bias);
      const float L = fminf(fmaxf(R, -30.f), 30.f);
      const float ex = __expf(-fabsf(L));
      const float sp = fmaxf(L, 0.f) + log1pf(ex);
      const float sig = L >= 0.f ? 1.f / (1.f + ex) : ex / (1.f + ex);
      const float inr = fabsf(R) <= 30.f ? 1.f : 0.f;
      float w = p.negative_weight;
      if (yv > 0.5f) w = p.auto_balance ? ratio : (p.use_pw ? pwv * p.positive_weight : p.positive_weight);
      const float g_full = w * (sig - yv) * inr * p.c;
      const float g_dense = p.negative_weight * p.c * sig * inr;
      if (lane == 0) {
        a_loss += (double)((w * (sp - L * yv) - p.negative_weight * sp) * p.c);
        a_bias += (double)(g_full - g_dense);
        a_t += (double)((g_full - g_dense) * s);
      }
      if (p.dV) {
        // the dense tile kernel fed bf16(g_dense * gnorm) (+ the bf16 residual when hp) to the tensor core
        const float gs = g_dense * p.gnorm;
        float gr = __bfloat162float(__float2bfloat16_rn(gs));
        if (p.hp) gr += __bfloat162float(__float2bfloat16_rn(gs - gr));
        const float dg = (g_full - gr / p.gnorm) * inv_tau;
        float* dv = p.dV + (size_t)row * p.lddv;
        float* dt = p.dT + (size_t)j * p.lddt;
        for (int d = lane; d < p.D; d += 32) {
          dv[d] += dg * __bfloat162float(tr[p.hi_off + d]);                  // this warp owns row i
          atomicAdd(dt + d, dg * __bfloat162float(vr[p.hi_off + d]));
        }
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
  static bool attr_done = false;
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

int siglip_compact(const float* mask, long ldm, const float* pw, long ldw, int B, int T, int cap, int* col, float* y,
                   float* w, int* cnt, float* ysum, int* overflow, cudaStream_t s) {
  if (B <= 0 || T <= 0 || cap <= 0) return B2_EINVAL;
  siglip_compact_kernel<<<(B + 7) / 8, 256, 0, s>>>(mask, ldm, pw, ldw, B, T, cap, col, y, w, cnt, ysum, overflow);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int siglip_pos(const void* V, int ldv, const void* T, int ldt, int K, int Dp, int D, int hi_off, int B, int Tn, int cap,
               const int* col, const float* y, const float* w, const int* cnt, const float* ysum, const float* dyn,
               float positive_weight, float negative_weight, float c, float gnorm, int hp, int use_pw, int auto_balance,
               float* dV, int lddv, float* dT, int lddt, double* acc, cudaStream_t s) {
  if (B <= 0 || K <= 0 || (K & 1)) return B2_EINVAL;
  if ((dV == nullptr) != (dT == nullptr)) return B2_EINVAL;
  PosParams p{(const __nv_bfloat16*)V, ldv, (const __nv_bfloat16*)T, ldt, K, Dp, D, hi_off, B, Tn, cap, col, y, w, cnt,
              ysum, dyn, positive_weight, negative_weight, c, gnorm > 0.f ? gnorm : 1.f, hp ? 1 : 0, use_pw, auto_balance, dV, lddv, dT,
              lddt, acc};
  siglip_pos_kernel<<<(B + 7) / 8, 256, 0, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
