// CUDA-core kernels of siglip.cu (scalar tails, positive-list compaction, sparse positive corrections; see there).
// No inline PTX and no include: CUDA types / intrinsics and warp_sum come from the including translation unit (common.cuh)
// or from the host emulation (tests/emul/).
#pragma once

namespace b2 {

// rowvec[i] = {1/Z_i, m_i = H_i - Q_i}; stats = {sum_i H_i, min_i H_i, max_i H_i} over the local rows. One CTA,
// fixed reduction order (deterministic).
__global__ void __launch_bounds__(1024)
siglip_entropy_rows_kernel(const float* __restrict__ Z, const float* __restrict__ H, const float* __restrict__ Q, int B,
                           float* __restrict__ rowvec, double* __restrict__ stats) {
  double sum = 0.0;
  float mn = 3.4e38f, mx = -3.4e38f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float h = H[i];
    rowvec[2 * i] = 1.f / Z[i];
    rowvec[2 * i + 1] = h - Q[i];
    sum += (double)h;
    mn = fminf(mn, h);
    mx = fmaxf(mx, h);
  }
  __shared__ double ssum[32];
  __shared__ float smn[32], smx[32];
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    ssum[threadIdx.x >> 5] = sum;
    smn[threadIdx.x >> 5] = mn;
    smx[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      sum += ssum[w];
      mn = fminf(mn, smn[w]);
      mx = fmaxf(mx, smx[w]);
    }
    stats[0] = sum;
    stats[1] = (double)mn;
    stats[2] = (double)mx;
  }
}

// stats_all [W][3] (one triple per rank) -> mean / min / max entropy over the B_global rows, the penalty and the gradient
// coefficient: out = {mean, min, max, mean / ln T, deficit = relu(thr - mean), weight * deficit};
// dyn[10] = deficit > 0 ? -weight / B_global : 0.
__global__ void siglip_entropy_coef_kernel(const double* __restrict__ stats_all, int W, int Bg, int T, float weight,
                                           float thr, float* __restrict__ dyn, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double sum = 0.0, mn = 3.4e38, mx = -3.4e38;
  for (int r = 0; r < W; ++r) {
    sum += stats_all[3 * r];
    mn = fmin(mn, stats_all[3 * r + 1]);
    mx = fmax(mx, stats_all[3 * r + 2]);
  }
  const float mean = (float)(sum / (double)Bg);
  const float deficit = fmaxf(thr - mean, 0.f);
  out[0] = mean;
  out[1] = (float)mn;
  out[2] = (float)mx;
  out[3] = mean / logf((float)T);
  out[4] = deficit;
  out[5] = weight * deficit;
  dyn[10] = deficit > 0.f ? -weight / (float)Bg : 0.f;
}

// red[0..2] = loss, dbias, sum G*s of this rank; red[3] / red[4] = a position-weighted fp64 checksum of the text operand's
// inverse norms and its square: after the all-reduce over W ranks, W * sum c^2 == (sum c)^2 exactly when every rank
// held the same texts (Cauchy-Schwarz) — siglip_loss_out turns the loss into NaN otherwise (text_replicated contract).
// One block of 256 threads, fixed reduction order (bit-identical on every rank for identical texts).
__global__ void __launch_bounds__(256)
siglip_combine_kernel(const double* __restrict__ acc, double wn_c, const float* __restrict__ tinv, int T,
                      double* __restrict__ red) {
  double c = 0.0;
  if (tinv)
    for (int j = threadIdx.x; j < T; j += 256) c += (double)tinv[j] * (1.0 + (double)(j & 1023) * (1.0 / 1024.0));
  __shared__ double sh[256];
  sh[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  red[0] = wn_c * acc[1] + acc[4];
  red[1] = acc[2] + acc[5];
  red[2] = acc[0] + acc[6];
  red[3] = sh[0];
  red[4] = sh[0] * sh[0];
}
__global__ void siglip_loss_out_kernel(const double* __restrict__ red, const int* __restrict__ overflow,
                                       const float* __restrict__ ent, int world, float* __restrict__ loss_out,
                                       float* __restrict__ diag) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double loss = red[0];
  if (overflow && overflow[0] > 0) loss = nan("");
  if (world > 1) {
    const double s2 = red[3] * red[3];
    if (fabs((double)world * red[4] - s2) > 1e-11 * s2) loss = nan("");      // the ranks' texts differ
  }
  if (diag) {
    for (int i = 0; i < 6; ++i) diag[i] = ent ? ent[i] : 0.f;
    diag[6] = (float)loss;
  }
  if (ent) loss += (double)ent[5];
  loss_out[0] = (float)loss;
}
__global__ void siglip_scalar_grads_kernel(const double* __restrict__ red, const float* __restrict__ dyn,
                                           const float* __restrict__ gmul, float* __restrict__ dlt,
                                           float* __restrict__ dbias) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double g = (double)gmul[0];
  if (dlt) dlt[0] = (float)(-(red[2] * (double)dyn[2]) * (double)dyn[7] * g);
  if (dbias) dbias[0] = (float)(red[1] * g);
}

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
  int use_pw, auto_balance;  // use_pw: bit 0 = per-pair pos_weights, bit 1 = weight rule "mask > 0" instead of "target > 0.5"
  float* dV; int lddv;     // may be null (loss only)
  float* dT; int lddt;
  double* acc;
  // Optional raw features (dtype code 0 fp32 / 1 bf16 / 2 fp16) + 1/max(||x||, eps): the gradient of a positive pair
  // is then formed with the fp32 normalised partner row x * inv instead of its bf16 operand. A video row has a handful of
  // positives carrying almost all of its gradient, so the 2^-9 rounding of the operand rows is not averaged away as it is
  // over the thousands of negatives: it was a flat 1.7e-3 relative gradient error at every size (north_star: 2e-3).
  const void* Vraw; int v_dtype; long long ld_vraw; const float* vinv;
  const void* Traw; int t_dtype; long long ld_traw; const float* tinv;
};

__device__ __forceinline__ float pos_ld(const void* base, int dtype, long long idx) {
  if (dtype == 0) return static_cast<const float*>(base)[idx];
  if (dtype == 1) return __bfloat162float(static_cast<const __nv_bfloat16*>(base)[idx]);
  return __half2float(static_cast<const __half*>(base)[idx]);
}

__global__ void __launch_bounds__(256) siglip_pos_kernel(PosParams p) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  double a_loss = 0.0, a_bias = 0.0, a_t = 0.0;
  if (row < p.B) {
    const float inv_tau = p.dyn[2], bias = p.dyn[5], lc = p.dyn[8], yneg = p.dyn[9];
    const bool use_pw = (p.use_pw & 1) != 0, rule_mask = (p.use_pw & 2) != 0;
    const int n = p.cnt[row];
    float ratio = 1.f;
    if (p.auto_balance) {
      const float pc = fmaxf(p.ysum[row], 1.f);
      ratio = fmaxf(((float)p.Tn - pc) / pc, 1.f);
    }
    const __nv_bfloat16* vr = p.V + (size_t)row * p.ldv;
    for (int e = 0; e < n; ++e) {
      const int j = p.col[(size_t)row * p.cap + e];
      const float yraw = p.y[(size_t)row * p.cap + e];
      const float yv = fmaf(yraw, 1.f - 2.f * yneg, yneg);          // label smoothing: y (1 - eps) + eps / 2
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
      // what the dense pass saw (bf16 operands): removed again below
      const float R = fmaf(s, inv_tau, bias);
      const float L = fminf(fmaxf(R, -lc), lc);
      const float ex = __expf(-fabsf(L));
      const float sp = fmaxf(L, 0.f) + log1pf(ex);
      const float sig = L >= 0.f ? 1.f / (1.f + ex) : ex / (1.f + ex);
      const float inr = fabsf(R) <= lc ? 1.f : 0.f;
      // the pair's true term: cosine from the raw features in fp32 when they are given
      float sx = s, Lx = L, spx = sp, sigx = sig, inrx = inr;
      if (p.Vraw && p.Traw) {
        float a = 0.f;
        for (int d = lane; d < p.D; d += 32)
          a = fmaf(pos_ld(p.Vraw, p.v_dtype, (long long)row * p.ld_vraw + d),
                   pos_ld(p.Traw, p.t_dtype, (long long)j * p.ld_traw + d), a);
        sx = warp_sum(a) * p.vinv[row] * p.tinv[j];
        const float Rx = fmaf(sx, inv_tau, bias);
        Lx = fminf(fmaxf(Rx, -lc), lc);
        const float exx = __expf(-fabsf(Lx));
        spx = fmaxf(Lx, 0.f) + log1pf(exx);
        sigx = Lx >= 0.f ? 1.f / (1.f + exx) : exx / (1.f + exx);
        inrx = fabsf(Rx) <= lc ? 1.f : 0.f;
      }
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
        // the dense tile kernel fed bf16(g_dense * gnorm) (+ the bf16 residual when hp) to the tensor core
        const float gs = g_dense * p.gnorm;
        float gr = __bfloat162float(__float2bfloat16_rn(gs));
        if (p.hp) gr += __bfloat162float(__float2bfloat16_rn(gs - gr));
        const float dgf = g_full * inv_tau, dgd = gr / p.gnorm * inv_tau;
        float* dv = p.dV + (size_t)row * p.lddv;
        float* dt = p.dT + (size_t)j * p.lddt;
        const float ti = p.Traw ? p.tinv[j] : 0.f, vi = p.Vraw ? p.vinv[row] : 0.f;
        for (int d = lane; d < p.D; d += 32) {
          const float th = __bfloat162float(tr[p.hi_off + d]), vh = __bfloat162float(vr[p.hi_off + d]);
          const float tx = p.Traw ? pos_ld(p.Traw, p.t_dtype, (long long)j * p.ld_traw + d) * ti : th;
          const float vx = p.Vraw ? pos_ld(p.Vraw, p.v_dtype, (long long)row * p.ld_vraw + d) * vi : vh;
          dv[d] += dgf * tx - dgd * th;                                        // this warp owns row i
          atomicAdd(dt + d, dgf * vx - dgd * vh);
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
