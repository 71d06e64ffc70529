// Kernel of querypool.cu (see there). No inline PTX and no include: CUDA intrinsics and B2_DYN_SMEM_F32(name) (the dynamic
// shared memory window as float[]) come from the including translation unit or from the host emulation (tests/emul/).
#pragma once

namespace b2 {

struct QpParams {
  const float* x; long long sb, sn;          // [B, N, D]
  const float* pos;                          // [>=N, D] or null
  const float* lnw; const float* lnb; const float* q;   // [D]
  const unsigned char* mask; long long mb;   // [B, N] or null
  int B, N, D; float eps;
  float* out;                                // fwd: [B, D]
  const float* dout;                         // bwd: [B, D]
  float* dx; float* dpos; float* dlnw; float* dlnb; float* dq;   // bwd outputs (dpos/dlnw/dlnb/dq accumulated atomically)
};

// shared: ln [N][D], w [N], s [N], xh-stats mean/rstd [N]
template <bool kBwd>
__global__ void __launch_bounds__(256) querypool_kernel(QpParams p) {
  B2_DYN_SMEM_F32(sm);
  float* ln = sm;                             // [N][D]  (post-LN, masked rows zeroed)
  float* w = ln + (size_t)p.N * p.D;          // [N]
  float* sc = w + p.N;                        // [N]
  float* mean = sc + p.N;                     // [N]
  float* rstd = mean + p.N;                   // [N]
  float* red = rstd + p.N;                    // [N] scratch (dw / ds)
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float* xb = p.x + b * p.sb;
  const unsigned char* mk = p.mask ? p.mask + b * p.mb : nullptr;

  for (int n = warp; n < p.N; n += nw) {
    const float* xr = xb + n * p.sn;
    const float* pr = p.pos ? p.pos + (size_t)n * p.D : nullptr;
    float s1 = 0.f;
    for (int d = lane; d < p.D; d += 32) s1 += xr[d] + (pr ? pr[d] : 0.f);
    s1 = warp_sum(s1);
    const float mu = s1 / p.D;
    float s2 = 0.f;
    for (int d = lane; d < p.D; d += 32) {
      const float v = xr[d] + (pr ? pr[d] : 0.f) - mu;
      s2 = fmaf(v, v, s2);
    }
    s2 = warp_sum(s2);
    const float rs = rsqrtf(s2 / p.D + p.eps);
    const bool masked = mk && mk[n];
    float dot = 0.f;
    for (int d = lane; d < p.D; d += 32) {
      const float xh = (xr[d] + (pr ? pr[d] : 0.f) - mu) * rs;
      const float v = masked ? 0.f : fmaf(xh, p.lnw[d], p.lnb[d]);
      ln[(size_t)n * p.D + d] = v;
      dot = fmaf(v, p.q[d], dot);
    }
    dot = warp_sum(dot);
    if (lane == 0) {
      mean[n] = mu;
      rstd[n] = rs;
      sc[n] = masked ? -INFINITY : dot;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = -INFINITY;
    for (int n = 0; n < p.N; ++n) m = fmaxf(m, sc[n]);
    float l = 0.f;
    for (int n = 0; n < p.N; ++n) l += (sc[n] == -INFINITY) ? 0.f : __expf(sc[n] - m);
    for (int n = 0; n < p.N; ++n) w[n] = (m == -INFINITY || sc[n] == -INFINITY) ? 0.f : __expf(sc[n] - m) / l;
    // all views masked: softmax is NaN -> nan_to_num -> 0 -> fallback valid/sum(valid) = 0 (reference :143-152)
  }
  __syncthreads();
  if (!kBwd) {
    for (int d = threadIdx.x; d < p.D; d += blockDim.x) {
      float o = 0.f;
      for (int n = 0; n < p.N; ++n) o = fmaf(w[n], ln[(size_t)n * p.D + d], o);
      p.out[(size_t)b * p.D + d] = o;
    }
    return;
  }
  // ---------------- backward ----------------
  const float* go = p.dout + (size_t)b * p.D;
  for (int n = warp; n < p.N; n += nw) {
    float dw = 0.f;
    for (int d = lane; d < p.D; d += 32) dw = fmaf(go[d], ln[(size_t)n * p.D + d], dw);
    dw = warp_sum(dw);
    if (lane == 0) red[n] = dw;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int n = 0; n < p.N; ++n) t = fmaf(w[n], red[n], t);
    for (int n = 0; n < p.N; ++n) red[n] = w[n] * (red[n] - t);       // ds_n (0 for masked / all-masked)
  }
  __syncthreads();
  // dq += sum_n ds_n ln_n
  for (int d = threadIdx.x; d < p.D; d += blockDim.x) {
    float a = 0.f;
    for (int n = 0; n < p.N; ++n) a = fmaf(red[n], ln[(size_t)n * p.D + d], a);
    if (a != 0.f) atomicAdd(p.dq + d, a);
  }
  // per view: dln = w_n dout + ds_n q (0 on masked rows) -> LayerNorm backward -> dx ; parameter grads
  for (int n = warp; n < p.N; n += nw) {
    const bool masked = mk && mk[n];
    const float* xr = xb + n * p.sn;
    const float* pr = p.pos ? p.pos + (size_t)n * p.D : nullptr;
    const float mu = mean[n], rs = rstd[n], wn = w[n], dsn = red[n];
    float a1 = 0.f, a2 = 0.f;     // mean(dxh), mean(dxh * xh)
    for (int d = lane; d < p.D; d += 32) {
      const float xh = (xr[d] + (pr ? pr[d] : 0.f) - mu) * rs;
      const float dl = masked ? 0.f : fmaf(wn, go[d], dsn * p.q[d]);
      const float dxh = dl * p.lnw[d];
      a1 += dxh;
      a2 = fmaf(dxh, xh, a2);
      if (dl != 0.f) {
        atomicAdd(p.dlnw + d, dl * xh);
        atomicAdd(p.dlnb + d, dl);
      }
    }
    a1 = warp_sum(a1) / p.D;
    a2 = warp_sum(a2) / p.D;
    for (int d = lane; d < p.D; d += 32) {
      const float xh = (xr[d] + (pr ? pr[d] : 0.f) - mu) * rs;
      const float dl = masked ? 0.f : fmaf(wn, go[d], dsn * p.q[d]);
      const float g = (dl * p.lnw[d] - a1 - xh * a2) * rs;
      p.dx[((size_t)b * p.N + n) * p.D + d] = g;
      if (p.dpos && g != 0.f) atomicAdd(p.dpos + (size_t)n * p.D + d, g);
    }
  }
}

}  // namespace b2
