// The runner's inline multi-positive branch FROM FEATURES (reference runners/video_constrative_learning_runner.py:1256-1322,
// validation twin :1585-1641): L2-normalise both sides, gated logits L_ij = s sigmoid(s) / tau (+ margin on "abnormal" text
// columns), then either the weighted multi-positive softmax CE both ways (utils/loss/weighted_siglip.py:38-51) or the
// weighted BCE-sum / max(1, sum targets), plus the three alignment scalars the runner logs from the same logits (:1296-1311).
//
// The reference materialises five [B, M] matrices per step (similarity, gated, logits, two log-softmaxes) and their
// autograd twins. Here NOTHING of size B x M is written: the branch runs on the rank-LOCAL batch (it bypasses the gathering
// registry losses), B and M are a few dozen to a few hundred rows, so the problem is latency-bound and far below one
// 128 x 256 tensor-core tile per SM — one CTA per row recomputes its similarity row in fp32 on CUDA cores (exact fp32 logits:
// no bf16 operand rounding at all) once per statistic sweep and once per gradient sweep:
//   imp_stats<side>  CTA = one video row (side 0) or one text column (side 1): running max / sum-exp, sum pos, sum pos L,
//                    the BCE sums and the alignment partials of that row.
//   imp_finalize     one CTA: loss, sum targets, alignment_logprob / prob / cosine.
//   imp_grad<side>   CTA = one row again: dL_ij from the row AND column statistics, dS = dL g'(s) / tau, the row of
//                    dXhat accumulated in registers, L2-normalise backward applied in place, dlog_temp by fp64 atomics.
#include "common.cuh"
#include "host_api.h"

namespace b2 {

constexpr int IMP_THREADS = 256;
constexpr int IMP_MAXD = 1024;       // features per row (registers: D / 32 accumulators per lane)

struct ImpParams {
  const float* v; const float* t; long long ldv, ldt;      // [B, D], [M, D] raw features, fp32
  const float* targets; const float* pw; long long ldm;    // [B, M] positive mask, optional positive weights (same pitch)
  const float* abn; float margin;                          // [M] 0/1 abnormal flags or null
  const float* log_temp;
  const int* pw_nonzero;                                   // device flag: any(positive_weights != 0)
  int B, M, D, mode;                                       // mode 0: weighted softmax CE both ways, 1: weighted BCE sum
  float eps, neg_w;
  float* rstat; float* cstat;                              // [B, 8], [M, 8] statistics
  float* scal;                                             // [8]: loss, sum targets, logprob, prob, cosine, denom, n valid rows
  const float* gout;                                       // upstream gradient (device scalar) or null
  float* dv; float* dt; double* dlt_acc;                   // gradients
};

__device__ __forceinline__ float imp_sigmoid(float x) { return 1.f / (1.f + __expf(-x)); }

// positive weight used by the loss for pair (i, j): targets * positive_weights when weights were given and not all zero
__device__ __forceinline__ float imp_pos(const ImpParams& p, float tg, float pwv, bool use_pw) {
  return use_pw ? tg * pwv : tg;
}

// One CTA per row of `side` (0: video row i against all texts, 1: text row j against all videos).
// stat row: [0] lse, [1] R = sum pos (clamped at 0), [2] A = sum pos L, [3] inverse norm of the row,
//           side 0 only: [4] sum pwm logp-numerator pieces -> stored as sum pwm L, [5] sum pwm, [6] sum_{targets>0} S, [7] count
// BCE mode (side 0): [1] sum w bce, [2] sum targets.
template <int SIDE>
__global__ void __launch_bounds__(IMP_THREADS) imp_stats_kernel(ImpParams p) {
  __shared__ float xs[IMP_MAXD];
  __shared__ float red[8][8];
  const int row = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nother = SIDE == 0 ? p.M : p.B, D = p.D;
  const float* x = SIDE == 0 ? p.v + (long long)row * p.ldv : p.t + (long long)row * p.ldt;
  // normalise the row into shared memory
  float ss = 0.f;
  for (int d = threadIdx.x; d < D; d += IMP_THREADS) { const float a = x[d]; xs[d] = a; ss = fmaf(a, a, ss); }
  ss = warp_sum(ss);
  if (lane == 0) red[warp][0] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[w][0];
  const float inv = 1.f / fmaxf(sqrtf(tot), 1e-12f);
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += IMP_THREADS) xs[d] *= inv;
  __syncthreads();
  const float itau = __expf(-p.log_temp[0]);
  const bool use_pw = p.pw && p.pw_nonzero[0] != 0;
  float m = -INFINITY, se = 0.f, R = 0.f, A = 0.f, pl = 0.f, ps = 0.f, cs = 0.f, cn = 0.f;
  for (int o = warp; o < nother; o += 8) {
    const float* y = SIDE == 0 ? p.t + (long long)o * p.ldt : p.v + (long long)o * p.ldv;
    float dot = 0.f, yy = 0.f;
    for (int d = lane; d < D; d += 32) { const float b = y[d]; dot = fmaf(xs[d], b, dot); yy = fmaf(b, b, yy); }
    dot = warp_sum(dot); yy = warp_sum(yy);
    const float s = dot / fmaxf(sqrtf(yy), 1e-12f);
    const int vi = SIDE == 0 ? row : o, tj = SIDE == 0 ? o : row;
    const float L = s * imp_sigmoid(s) * itau + (p.abn ? p.abn[tj] * p.margin : 0.f);
    const float tg = p.targets[(long long)vi * p.ldm + tj];
    const float pwv = p.pw ? p.pw[(long long)vi * p.ldm + tj] : 0.f;
    if (p.mode == 0) {
      const float pos = fmaxf(imp_pos(p, tg, pwv, use_pw), 0.f);
      const float mn = fmaxf(m, L);
      se = se * __expf(m - mn) + __expf(L - mn);
      m = mn;
      R += pos;
      A = fmaf(pos, L, A);
    } else if (SIDE == 0) {
      const float w = p.pw ? (tg > 0.f ? pwv : p.neg_w) : p.neg_w;
      const float bce = fmaxf(L, 0.f) - L * tg + log1pf(__expf(-fabsf(L)));     // BCE-with-logits, stable form
      R = fmaf(w, bce, R);
      A += tg;
      const float mn = fmaxf(m, L);
      se = se * __expf(m - mn) + __expf(L - mn);
      m = mn;
    }
    if (SIDE == 0) {
      const float pwm = use_pw ? pwv * tg : tg;          // weights of the logged alignment log-probability (:1298-1302)
      pl = fmaf(pwm, L, pl);
      ps += pwm;
      if (tg != 0.f) { cs += s; cn += 1.f; }            // similarity[targets.bool()] (:1309)
    }
  }
  // every lane of a warp holds the same values: combine the 8 warps
  if (lane == 0) { red[warp][0] = m; red[warp][1] = se; red[warp][2] = R; red[warp][3] = A; red[warp][4] = pl; red[warp][5] = ps;
                   red[warp][6] = cs; red[warp][7] = cn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float M_ = -INFINITY;
    for (int w = 0; w < 8; ++w) M_ = fmaxf(M_, red[w][0]);
    float S_ = 0.f, R_ = 0.f, A_ = 0.f, pl_ = 0.f, ps_ = 0.f, cs_ = 0.f, cn_ = 0.f;
    for (int w = 0; w < 8; ++w) {
      S_ += red[w][0] == -INFINITY ? 0.f : red[w][1] * __expf(red[w][0] - M_);
      R_ += red[w][2]; A_ += red[w][3]; pl_ += red[w][4]; ps_ += red[w][5]; cs_ += red[w][6]; cn_ += red[w][7];
    }
    float* st = (SIDE == 0 ? p.rstat : p.cstat) + (size_t)row * 8;
    st[0] = M_ + __logf(S_);
    st[1] = R_; st[2] = A_; st[3] = inv; st[4] = pl_; st[5] = ps_; st[6] = cs_; st[7] = cn_;
  }
}

// one CTA: scalars. scal[0] loss, [1] sum targets, [2] alignment_logprob, [3] alignment_prob, [4] alignment_cosine,
// [5] BCE denominator max(1, sum targets), [6] rows with a positive (valid rows of the alignment mean)
__global__ void __launch_bounds__(IMP_THREADS) imp_finalize_kernel(ImpParams p) {
  __shared__ double sh[IMP_THREADS][6];
  double lr = 0., lc = 0., tg = 0., lp = 0., nv = 0., cs = 0., cn = 0., bsum = 0.;
  for (int i = threadIdx.x; i < p.B; i += IMP_THREADS) {
    const float* st = p.rstat + (size_t)i * 8;
    if (p.mode == 0) lr += ((double)st[0] * st[1] - st[2]) / fmax((double)st[1], (double)p.eps);
    else { bsum += st[1]; tg += st[2]; }
    if (st[5] > 0.f) { lp += ((double)st[4] - (double)st[0] * st[5]) / st[5]; nv += 1.; }
    cs += st[6]; cn += st[7];
  }
  if (p.mode == 0)
    for (int j = threadIdx.x; j < p.M; j += IMP_THREADS) {
      const float* st = p.cstat + (size_t)j * 8;
      lc += ((double)st[0] * st[1] - st[2]) / fmax((double)st[1], (double)p.eps);
    }
  double vals[8] = {lr, lc, tg, lp, nv, cs, cn, bsum};
  __shared__ double tot[8];
  for (int k = 0; k < 8; ++k) {
    sh[threadIdx.x][0] = vals[k];
    __syncthreads();
    for (int s = IMP_THREADS / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s) sh[threadIdx.x][0] += sh[threadIdx.x + s][0];
      __syncthreads();
    }
    if (threadIdx.x == 0) tot[k] = sh[0][0];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double denom = fmax(1.0, tot[2]);
    p.scal[0] = p.mode == 0 ? (float)(0.5 * (tot[0] / p.B + tot[1] / p.M)) : (float)(tot[7] / denom);
    p.scal[1] = (float)tot[2];
    const double alp = tot[4] > 0. ? tot[3] / tot[4] : NAN;          // no valid row: the reference leaves it None
    p.scal[2] = (float)alp;
    p.scal[3] = (float)exp(alp);
    p.scal[4] = tot[6] > 0. ? (float)(tot[5] / tot[6]) : NAN;
    p.scal[5] = (float)denom;
    p.scal[6] = (float)tot[4];
  }
}

// One CTA per row of `side`: gradient row of the normalised features, normalise-backward, dlog_temp (side 0 only).
template <int SIDE>
__global__ void __launch_bounds__(IMP_THREADS) imp_grad_kernel(ImpParams p) {
  __shared__ float xs[IMP_MAXD];
  __shared__ float acc_s[8][IMP_MAXD];
  __shared__ float red[8];
  const int row = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nother = SIDE == 0 ? p.M : p.B, D = p.D;
  const float* x = SIDE == 0 ? p.v + (long long)row * p.ldv : p.t + (long long)row * p.ldt;
  const float* own = (SIDE == 0 ? p.rstat : p.cstat) + (size_t)row * 8;
  const float* oth = SIDE == 0 ? p.cstat : p.rstat;
  const float inv = own[3];
  for (int d = threadIdx.x; d < D; d += IMP_THREADS) xs[d] = x[d] * inv;
  __syncthreads();
  const float itau = __expf(-p.log_temp[0]);
  const bool use_pw = p.pw && p.pw_nonzero[0] != 0;
  const float gmul = p.gout ? p.gout[0] : 1.f;
  const float denom = p.scal[5];
  float acc[IMP_MAXD / 32];
#pragma unroll
  for (int c = 0; c < IMP_MAXD / 32; ++c) acc[c] = 0.f;
  float dlt = 0.f;
  for (int o = warp; o < nother; o += 8) {
    const float* y = SIDE == 0 ? p.t + (long long)o * p.ldt : p.v + (long long)o * p.ldv;
    const float oinv = oth[(size_t)o * 8 + 3];
    float dot = 0.f;
    for (int d = lane; d < D; d += 32) dot = fmaf(xs[d], y[d], dot);
    const float s = warp_sum(dot) * oinv;
    const int vi = SIDE == 0 ? row : o, tj = SIDE == 0 ? o : row;
    const float sg = imp_sigmoid(s);
    const float g = s * sg;
    const float L = g * itau + (p.abn ? p.abn[tj] * p.margin : 0.f);
    const float tg = p.targets[(long long)vi * p.ldm + tj];
    const float pwv = p.pw ? p.pw[(long long)vi * p.ldm + tj] : 0.f;
    const float* rs = p.rstat + (size_t)vi * 8;
    const float* cst = p.cstat + (size_t)tj * 8;
    float dL;
    if (p.mode == 0) {
      const float pos = fmaxf(imp_pos(p, tg, pwv, use_pw), 0.f);
      const float dr = fmaxf(rs[1], p.eps), dc = fmaxf(cst[1], p.eps);
      dL = 0.5f / (float)p.B * (rs[1] * __expf(L - rs[0]) - pos) / dr + 0.5f / (float)p.M * (cst[1] * __expf(L - cst[0]) - pos) / dc;
    } else {
      const float w = p.pw ? (tg > 0.f ? pwv : p.neg_w) : p.neg_w;
      dL = w * (imp_sigmoid(L) - tg) / denom;
    }
    dL *= gmul;
    const float dS = dL * itau * sg * (1.f + s * (1.f - sg));        // d (s sigmoid(s)) / ds
    if (SIDE == 0) dlt = fmaf(dL, g * itau, dlt);                    // d L / d log_temp = -g / tau
    const float coef = dS * oinv;
#pragma unroll
    for (int c = 0; c < IMP_MAXD / 32; ++c) {
      const int d = lane + 32 * c;
      if (d < D) acc[c] = fmaf(coef, y[d], acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < IMP_MAXD / 32; ++c) {
    const int d = lane + 32 * c;
    if (d < D) acc_s[warp][d] = acc[c];
  }
  __syncthreads();
  // dxhat[d] = sum over warps; dx = inv (dxhat - xhat (xhat . dxhat))
  float part = 0.f;
  for (int d = threadIdx.x; d < D; d += IMP_THREADS) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += acc_s[w][d];
    acc_s[0][d] = a;
    part = fmaf(a, xs[d], part);
  }
  part = warp_sum(part);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  float proj = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) proj += red[w];
  float* out = SIDE == 0 ? p.dv + (size_t)row * D : p.dt + (size_t)row * D;
  for (int d = threadIdx.x; d < D; d += IMP_THREADS) out[d] = inv * (acc_s[0][d] - xs[d] * proj);
  if (SIDE == 0 && p.dlt_acc && lane == 0 && dlt != 0.f) atomicAdd(p.dlt_acc, -(double)dlt);
}

__global__ void imp_flag_kernel(const float* pw, long long ld, int B, int M, int* flag) {
  // flag = any(positive_weights != 0); *flag zeroed by the caller
  int any = 0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)B * M; t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / M, j = t - i * M;
    any |= pw[i * ld + j] != 0.f;
  }
  if (__syncthreads_or(any) && threadIdx.x == 0) atomicOr(flag, 1);
}

}  // namespace b2

namespace b2host {
using namespace b2;

int inline_mp_fwd(const float* v, long long ldv, const float* t, long long ldt, const float* targets, const float* pw,
                  long long ldm, const float* abn, float margin, const float* log_temp, int B, int M, int D, int mode,
                  float eps, float neg_w, float* rstat, float* cstat, float* scal, int* flag, cudaStream_t s) {
  if (!v || !t || !targets || !log_temp || !rstat || !cstat || !scal || !flag || B <= 0 || M <= 0 || D <= 0 || D > IMP_MAXD ||
      (mode != 0 && mode != 1))
    return B2_EINVAL;
  ImpParams p{v, t, ldv, ldt, targets, pw, ldm, abn, margin, log_temp, flag, B, M, D, mode, eps, neg_w, rstat, cstat, scal,
              nullptr, nullptr, nullptr, nullptr};
  if (cudaMemsetAsync(flag, 0, sizeof(int), s) != cudaSuccess) return B2_ECUDA;
  if (pw) imp_flag_kernel<<<32, 256, 0, s>>>(pw, ldm, B, M, flag);
  imp_stats_kernel<0><<<B, IMP_THREADS, 0, s>>>(p);
  imp_stats_kernel<1><<<M, IMP_THREADS, 0, s>>>(p);
  imp_finalize_kernel<<<1, IMP_THREADS, 0, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int inline_mp_bwd(const float* v, long long ldv, const float* t, long long ldt, const float* targets, const float* pw,
                  long long ldm, const float* abn, float margin, const float* log_temp, int B, int M, int D, int mode,
                  float eps, float neg_w, const float* rstat, const float* cstat, const float* scal, const int* flag,
                  const float* gout, float* dv, float* dt, double* dlt_acc, cudaStream_t s) {
  if (!v || !t || !targets || !log_temp || !rstat || !cstat || !scal || !flag || B <= 0 || M <= 0 || D <= 0 || D > IMP_MAXD)
    return B2_EINVAL;
  ImpParams p{v, t, ldv, ldt, targets, pw, ldm, abn, margin, log_temp, flag, B, M, D, mode, eps, neg_w,
              const_cast<float*>(rstat), const_cast<float*>(cstat), const_cast<float*>(scal), gout, dv, dt, dlt_acc};
  if (dlt_acc && cudaMemsetAsync(dlt_acc, 0, sizeof(double), s) != cudaSuccess) return B2_ECUDA;
  if (dv) imp_grad_kernel<0><<<B, IMP_THREADS, 0, s>>>(p);
  if (dt) imp_grad_kernel<1><<<M, IMP_THREADS, 0, s>>>(p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
