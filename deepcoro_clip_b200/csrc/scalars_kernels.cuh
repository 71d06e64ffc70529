// Kernels of scalars.cu (device-side scalar plumbing of the losses; see there).
// No inline PTX and no include: CUDA types / intrinsics and warp_sum come from the including translation unit
// (common.cuh) or from the host emulation (tests/emul/).
#pragma once

namespace b2 {

constexpr float kStableWindowBits = 224.f;

__global__ void dyn_prep_kernel(const float* __restrict__ log_temp, const float* __restrict__ bias, float clamp_min,
                                float bound, float* __restrict__ dyn) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float tau = expf(log_temp[0]);
  float clamped = 0.f;
  if (clamp_min > 0.f && tau < clamp_min) {   // torch.clamp(min=1e-4): gradient is zero when active
    tau = clamp_min;
    clamped = 1.f;
  }
  const float scale2 = 1.4426950408889634f / tau;
  // P = 2^(f(S)*scale2 - shift2) with f(S) <= bound. shift2 = scale2*bound - OFF keeps the largest term at
  // 2^OFF; OFF > 0 only when the dynamic range 2*bound/tau would otherwise underflow fp32 (tau < ~0.023).
  float off = 2.f * bound * scale2 - 120.f;
  off = fminf(fmaxf(off, 0.f), 100.f);
  const float shift2 = scale2 * bound - off;
  dyn[0] = scale2;
  dyn[1] = shift2;
  dyn[2] = 1.f / tau;
  dyn[3] = tau;
  dyn[4] = clamped;
  dyn[5] = bias ? bias[0] : 0.f;
  dyn[6] = 0.6931471805599453f * shift2;
  dyn[7] = 1.f - clamped;
  dyn[8] = 30.f;
  dyn[9] = 0.f;
  dyn[10] = 0.f;
  // Stable softmax mode (selected here, on the device, from tau alone so that every rank of a job takes the same path and
  // the host never reads tau): the fixed shift keeps every term of every row / column representable only while the whole
  // range 2*bound*scale2 of the scaled logits fits into the fp32 exponent range next to the offset above
  // (100 + 126 bits), i.e. tau >= ~0.0128 for unit vectors. Below that (down to the reference's clamp floor 1e-4,
  // utils/loss/contrastive.py:153, and without any floor for the legacy classes, utils/loss/losses.py:53, 146) the
  // forward runs per-row / per-column maxima (logits_rowlse) and the backward forms two exponentials that are each <= 1.
  dyn[11] = (2.f * bound * scale2 > kStableWindowBits) ? 1.f : 0.f;
}

__global__ void dyn_set_siglip_kernel(float* __restrict__ dyn, float lclamp, float yneg) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  dyn[8] = lclamp;
  dyn[9] = yneg;
}

// Overrides the mode dyn_prep chose from tau: stable = 1 forces the running-maximum sweeps / two-exponential gradient (valid
// for every tau), stable = 0 forces the fixed shift (valid only inside its window). For A/B runs and the parity tests.
__global__ void dyn_set_stable_kernel(float* __restrict__ dyn, int stable) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  dyn[11] = stable ? 1.f : 0.f;
}

// acc[slot] += sum_r ln( sums[r] ) + ln2*shift2 ;  scale_out[r] = c / sums[r]
__global__ void __launch_bounds__(256)
lse_finalize_kernel(const float* __restrict__ sums, int n, const float* __restrict__ dyn, float c,
                    float* __restrict__ scale_out, double* __restrict__ acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (i < n) {
    const float s = sums[i];
    v = (double)logf(s) + (double)dyn[6];
    if (scale_out) scale_out[i] = c / s;
  }
  // block reduce in double
  __shared__ double sh[8];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 8) {
    v = sh[threadIdx.x];
    for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffu, v, o);
    if (threadIdx.x == 0 && acc) atomicAdd(acc, v);
  }
}

// acc += sum_i f(v[i])
__global__ void __launch_bounds__(256)
vec_fsum_kernel(const float* __restrict__ v, int n, int gated, double* __restrict__ acc) {
  double t = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double s = (double)v[i];
    t += gated ? s / (1.0 + exp(-s)) : s;
  }
  __shared__ double sh[8];
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double u = 0.0;
    for (int w = 0; w < 8; ++w) u += sh[w];
    atomicAdd(acc, u);
  }
}

// acc += sum_r f(a[r,:K] . b[r,:K]),  f = identity or s*sigmoid(s);  optionally stores the raw dots
__global__ void __launch_bounds__(256)
diag_sum_kernel(const __nv_bfloat16* __restrict__ a, int lda, const __nv_bfloat16* __restrict__ b, int ldb, int rows,
                int K, int gated, float* __restrict__ dots, double* __restrict__ acc) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  if (warp < rows) {
    const __nv_bfloat162* ar = reinterpret_cast<const __nv_bfloat162*>(a + (size_t)warp * lda);
    const __nv_bfloat162* br = reinterpret_cast<const __nv_bfloat162*>(b + (size_t)warp * ldb);
    for (int c = lane; c < K / 2; c += 32) {
      const float2 av = __bfloat1622float2(ar[c]);
      const float2 bv = __bfloat1622float2(br[c]);
      s = fmaf(av.x, bv.x, s);
      s = fmaf(av.y, bv.y, s);
    }
  }
  s = warp_sum(s);
  __shared__ double sh[8];
  if (lane == 0) {
    if (warp < rows && dots) dots[warp] = s;
    double f = 0.0;
    if (warp < rows) f = gated ? (double)s / (1.0 + exp(-(double)s)) : (double)s;
    sh[threadIdx.x >> 5] = f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(acc, t);
  }
}


// One launch for the whole scalar tail of the softmax (CLIP / gated) loss forward (utils/loss/contrastive.py:155-164).
//   sums = nvec vectors of n floats (already all-reduced across ranks):
//     [0] colsum            [1] rowsum            [2] target dots S_ii (tensor-core rounding)              (nvec >= 3)
//     [3] lse2 of the rows  [4] lse2 of the columns (log2 domain, logits_rowlse)
//     [5] S_ii from the raw features in fp32 (rowdot_raw)   [6] S_ii as the column sweep's tensor core saw it  (nvec == 7)
//   In the stable mode (dyn[11] != 0, nvec == 7) slots [0] / [1] hold the column / row GAPS lse2 - L2_ii of logits_rowlse.
// Fixed-shift mode: rowscale[i] = c / rowsum[i], colscale[j] = c / colsum[j] (c = 0.5 / N, the backward's softmax
//   denominators); row term t_i = ln rowsum_i + ln2 shift2 - L_ii.
// Stable mode: rowscale[i] = lse2_row[i] - log2 c, colscale likewise (the backward forms c * softmax as
//   2^(L2 - rowscale) + 2^(L2 - colscale), every exponential <= c); t_i = ln2 * gap_i.
// Both: L_ii above is the TENSOR-CORE value (it cancels against the same term inside the log-sum-exp when the softmax is
//   peaked); the fp32 target logit enters as a first-order correction, t_i -= (1 - p_ii) * (L_ii^fp32 - L_ii^tc) with
//   p_ii = exp(-t_i): the 2^-9 operand rounding of the target pair is not averaged over a row like the errors inside the
//   log-sum-exp, it was the whole 1e-5 loss error of plain bf16 operands at N <= 16k.
//   loss = c * sum_i (t_row_i + t_col_i + 2 eps L_ii^fp32) - unif / N        (eps: label smoothing; fp64 inside)
// Up to 64 CTAs; every CTA reduces its slice in fp64 and parks its partial sums in a device scratch block, the CTA that
// draws the last ticket adds the partials in CTA order (deterministic) and writes the loss. (Was one CTA: 28 us at
// N = 32k — 3 % of an 8-GPU step.) The scratch block is shared by all launches of the process: one stream at a time.
constexpr int FIN_MAX_BLOCKS = 64;
__device__ double g_fin_partial[3 * FIN_MAX_BLOCKS];
__device__ unsigned int g_fin_ticket = 0;

// Peer view of the statistics (multi-GPU one-shot exchange, replaces the NCCL all-reduce of the nvec * n floats): every rank
// accumulates into its OWN symmetric-memory block; after a cross-rank barrier the finalize kernel of every rank reads the
// blocks of all W peers directly (ld.global over NVLink): vector 0 (partial column sums of every rank's row slab) is summed
// over the peers, the other vectors are read from the one rank that owns the row (rows_per_rank = n / W).
constexpr int FIN_MAX_PEERS = 8;
struct FinPeers {
  const float* ptr[FIN_MAX_PEERS];
  int world;            // 0: single block (ptr[0] holds fully reduced sums)
  int rows_per_rank;
};
__device__ __forceinline__ float fin_ld(const FinPeers& ps, int n, int vec, int i) {
  if (ps.world <= 1) return ps.ptr[0][(size_t)vec * n + i];
  if (vec == 0) {
    float s = 0.f;
    for (int r = 0; r < ps.world; ++r) s += ps.ptr[r][i];          // fixed order: bit-identical on every rank
    return s;
  }
  return ps.ptr[i / ps.rows_per_rank][(size_t)vec * n + i];
}

__global__ void __launch_bounds__(1024)
clip_finalize_kernel(FinPeers sums, int n, int nvec, const float* __restrict__ dyn, float eps, int gated,
                     const double* __restrict__ unif, float* __restrict__ rowscale, float* __restrict__ colscale,
                     float* __restrict__ loss_out, double* __restrict__ acc_out) {
  const float c = 0.5f / (float)n;
  const double shift = (double)dyn[6], inv_tau = (double)dyn[2];
  const bool ext = nvec >= 7;
  const bool stable = ext && dyn[11] != 0.f;
  const float l2c = log2f(c);
  double a_row = 0.0, a_col = 0.0, a_dot = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double d = (double)fin_ld(sums, n, 2, i);
    const double fd = gated ? d / (1.0 + exp(-d)) : d;
    double fx = fd, fdc = fd;                                // fp32 target logit / the column sweep's tensor-core value
    if (ext) {
      const double dx = (double)fin_ld(sums, n, 5, i);
      fx = gated ? dx / (1.0 + exp(-dx)) : dx;
    }
    double tr, tc;
    if (stable) {
      rowscale[i] = fin_ld(sums, n, 3, i) - l2c;
      colscale[i] = fin_ld(sums, n, 4, i) - l2c;
      tr = 0.6931471805599453 * (double)fin_ld(sums, n, 1, i);
      tc = 0.6931471805599453 * (double)fin_ld(sums, n, 0, i);
      const double dc = (double)fin_ld(sums, n, 6, i);
      fdc = gated ? dc / (1.0 + exp(-dc)) : dc;
    } else {
      const float cs = fin_ld(sums, n, 0, i), rs = fin_ld(sums, n, 1, i);
      colscale[i] = c / cs;
      rowscale[i] = c / rs;
      tr = (double)logf(rs) + shift - fd * inv_tau;
      tc = (double)logf(cs) + shift - fd * inv_tau;
    }
    tr -= (1.0 - exp(-tr)) * (fx - fd) * inv_tau;
    tc -= (1.0 - exp(-tc)) * (fx - fdc) * inv_tau;
    a_row += tr;
    a_col += tc;
    a_dot += fx;
  }
  __shared__ double sh[3][32];
  __shared__ bool last;
  for (int o = 16; o > 0; o >>= 1) {
    a_row += __shfl_xor_sync(0xffffffffu, a_row, o);
    a_col += __shfl_xor_sync(0xffffffffu, a_col, o);
    a_dot += __shfl_xor_sync(0xffffffffu, a_dot, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = a_row;
    sh[1][threadIdx.x >> 5] = a_col;
    sh[2][threadIdx.x >> 5] = a_dot;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0.0, cc = 0.0, d = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { r += sh[0][w]; cc += sh[1][w]; d += sh[2][w]; }
    g_fin_partial[3 * blockIdx.x] = r;
    g_fin_partial[3 * blockIdx.x + 1] = cc;
    g_fin_partial[3 * blockIdx.x + 2] = d;
    __threadfence();
    last = atomicAdd(&g_fin_ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double r = 0.0, cc = 0.0, d = 0.0;
    for (int b = 0; b < (int)gridDim.x; ++b) {
      r += g_fin_partial[3 * b];
      cc += g_fin_partial[3 * b + 1];
      d += g_fin_partial[3 * b + 2];
    }
    g_fin_ticket = 0;                                   // ready for the next launch (stream order)
    const double u = unif ? unif[0] : 0.0;
    const double loss = (0.5 / n) * (r + cc) + ((double)eps * d * inv_tau - u) / n;
    loss_out[0] = (float)loss;
    if (acc_out) { acc_out[0] = r; acc_out[1] = cc; acc_out[2] = d; }
  }
}

// d loss / d log_temp = (unif / N - scal0 / tau) * [tau not clamped] * grad_out   (Appendix A.1: -sum G L)
__global__ void clip_dlogtemp_kernel(const double* __restrict__ scal0, const float* __restrict__ dyn,
                                     const float* __restrict__ gmul, const double* __restrict__ unif, int n,
                                     float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double u = unif ? unif[0] : 0.0;
  out[0] = (float)((u / n - scal0[0] * (double)dyn[2]) * (double)dyn[7] * (double)gmul[0]);
}

}  // namespace b2
