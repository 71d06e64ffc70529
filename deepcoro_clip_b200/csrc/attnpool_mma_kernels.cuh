// Kernels of attnpool_mma.cu (see there for the design). No inline PTX and no include: every hardware primitive comes from
// attnpool_mma_prims.cuh (device) or from tests/emul/pool_mma_prims_emul.h (host emulation), which must be included first.
#pragma once

namespace b2 {

constexpr int PM_THREADS = 256;    // 8 warps

// x tile in shared memory: D/64 TMA boxes of [PM_TT rows x 64 columns] (4 KB each, 128-byte rows, SWIZZLE_128B): the
// 16-byte chunk c (0..D/8) of row r lives in box c >> 3 at row r, chunk (c & 7) ^ (r & 7)  -> conflict-free ldmatrix.
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk, int /*row_bytes*/) {
  return base + (chunk >> 3) * (PM_TT * 128) + row * 128 + (((chunk & 7) ^ (row & 7)) << 4);
}

struct PmFwdParams {
  const void* x; long long sb, sn;
  const unsigned char* mask; long long mb;
  const float* qt;                     // [H, D] (softmax mode) or null
  const float* w; long long wb, wh;    // [B, H, N] given weights, or null
  float* part_m; float* part_l; float* part_acc;
  int B, N, D, H, S;
  float drop_p; unsigned long long drop_seed; float* part_l2;   // attention dropout: see attnpool.cu
  int stages;                          // cp.async pipeline depth (2..4 tiles of 32 tokens in flight)
};

// shared memory layout (bytes): x tiles stages * 32 * D * 2 | qt hi [8][D+8] | qt lo [8][D+8] | P hi [8][40] | P lo [8][40] |
//                               S_part [4][32][8] fp32 | scale [8] fp32
template <typename T, int NW>
__global__ void __launch_bounds__(NW * 32, NW == 8 ? 2 : 1) pool_fwd_mma_kernel(const __grid_constant__ CUtensorMap tmx, PmFwdParams p) {
  constexpr int KS = NW / 2;            // K splits of phase 1 (2 token groups x KS)
  constexpr int MAXC = 128 / NW;        // 16-channel chunks per warp at D = 1024
  B2_DYN_SMEM(smem_raw);
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int D = p.D, row_bytes = D * 2, QP = D + 8, PP = PM_TT + 8;
  const uint32_t s_base = smem_u32(smem);
  const int NS = p.stages, tile_bytes = PM_TT * row_bytes;
  uint16_t* q_hi = reinterpret_cast<uint16_t*>(smem + (size_t)NS * tile_bytes);
  uint16_t* q_lo = q_hi + 8 * QP;
  uint16_t* p_hi = q_lo + 8 * QP;
  uint16_t* p_lo = p_hi + 8 * PP;
  float* s_part = reinterpret_cast<float*>(p_lo + 8 * PP);
  float* s_scale = s_part + KS * PM_TT * 8;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_scale + 8);      // [NS] TMA tile landed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.x, sp = blockIdx.y;
  const int n0 = (int)((long long)p.N * sp / p.S), n1 = (int)((long long)p.N * (sp + 1) / p.S);
  const T* xb = reinterpret_cast<const T*>(p.x) + b * p.sb;
  const bool softmax_mode = p.w == nullptr;

  // prologue: barriers, then NS - 1 tiles in flight
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) mbar_init(&full_bar[i], 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmx);
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int i = 0; i < NS - 1; ++i) {
      const int tt = n0 + i * PM_TT;
      if (tt < n1) tma_x_tile(s_base + i * tile_bytes, &tmx, &full_bar[i], b * p.N + tt, D);
    }
  // qt -> hi / lo 16-bit, rows >= H zero
  if (softmax_mode)
    for (int i = threadIdx.x; i < 8 * D; i += NW * 32) {
      const int h = i / D, d = i - h * D;
      const float v = h < p.H ? p.qt[(size_t)h * D + d] : 0.f;
      const uint16_t hi = PmT<T>::bits(v);
      q_hi[h * QP + d] = hi;
      q_lo[h * QP + d] = PmT<T>::bits(v - PmT<T>::val(hi));
    }

  const int DW = D / NW;                // output channels per warp (multiple of 16)
  const int nchunk = DW / 16;
  float acc[MAXC][4];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
  float m_run = -INFINITY, l_run = 0.f, l2_run = 0.f;  // warp h < H owns head h's running max / sums (lane-uniform)
  const float keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  const unsigned char* mk = p.mask ? p.mask + b * p.mb : nullptr;
  const uint32_t qh_addr = smem_u32(q_hi), ql_addr = smem_u32(q_lo), ph_addr = smem_u32(p_hi), pl_addr = smem_u32(p_lo);

  int buf = 0;
  for (int t0 = n0; t0 < n1; t0 += PM_TT, buf = (buf + 1 == NS ? 0 : buf + 1)) {
    const int rows = min(PM_TT, n1 - t0);
    mbar_wait(&full_bar[buf], ((t0 - n0) / (PM_TT * NS)) & 1);
    __syncthreads();                                   // everyone is done with the previous tile and P
    if (threadIdx.x == 0) {
      const int tn = t0 + (NS - 1) * PM_TT, bn = (buf + NS - 1) % NS;      // refill the buffer the previous tile used
      if (tn < n1) tma_x_tile(s_base + bn * tile_bytes, &tmx, &full_bar[bn], b * p.N + tn, D);
    }
    const uint32_t tile = s_base + buf * tile_bytes;

    if (softmax_mode) {
      // ---- phase 1: S_part[kq][tok][h], warp = (token group tg, K quarter kq) ----
      const int tg = warp & 1, kq = warp >> 1;
      float c[4] = {0.f, 0.f, 0.f, 0.f}, cl[4] = {0.f, 0.f, 0.f, 0.f};
      const int kbeg = kq * (D / KS), kend = kbeg + D / KS;
      const int arow = tg * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      for (int k0 = kbeg; k0 < kend; k0 += 32) {          // D/KS is a multiple of 32: two k-steps, loads first
        uint32_t a0[4], a1[4];
        ldsm_x4(a0, tile_addr(tile, arow, (k0 >> 3) + (lane >> 4), row_bytes));
        ldsm_x4(a1, tile_addr(tile, arow, (k0 >> 3) + 2 + (lane >> 4), row_bytes));
        const uint32_t off = (uint32_t)(g * QP + k0 + 2 * t) * 2;
        const uint32_t l0 = lds32(ql_addr + off), l1 = lds32(ql_addr + off + 16), l2 = lds32(ql_addr + off + 32),
                       l3 = lds32(ql_addr + off + 48);
        const uint32_t h0 = lds32(qh_addr + off), h1 = lds32(qh_addr + off + 16), h2 = lds32(qh_addr + off + 32),
                       h3 = lds32(qh_addr + off + 48);
        PmT<T>::mma(cl, a0, l0, l1);
        PmT<T>::mma(c, a0, h0, h1);
        PmT<T>::mma(cl, a1, l2, l3);
        PmT<T>::mma(c, a1, h2, h3);
      }
      c[0] += cl[0]; c[1] += cl[1]; c[2] += cl[2]; c[3] += cl[3];
      float* sp_ = s_part + ((size_t)kq * PM_TT + tg * 16) * 8;
      *reinterpret_cast<float2*>(sp_ + g * 8 + 2 * t) = make_float2(c[0], c[1]);
      *reinterpret_cast<float2*>(sp_ + (g + 8) * 8 + 2 * t) = make_float2(c[2], c[3]);
      __syncthreads();
      // ---- online softmax: warp h handles head h, lane = token ----
      if (warp < 8) {
        const int h = warp;
        float s = -INFINITY;
        if (h < p.H && lane < rows && !(mk && mk[t0 + lane])) {
          s = 0.f;
#pragma unroll
          for (int q = 0; q < KS; ++q) s += s_part[(q * PM_TT + lane) * 8 + h];
        }
        const float m_new = fmaxf(m_run, warp_max(s));
        float sc = 1.f, pv = 0.f;
        if (m_new != -INFINITY) {
          sc = __expf(m_run - m_new);                  // exp(-inf) = 0 on the first unmasked tile
          pv = __expf(s - m_new);
        }
        l_run = l_run * sc + warp_sum(pv);             // softmax denominator: before dropout
        if (p.drop_p > 0.f) {
          pv = (h < p.H && attn_keep(p.drop_seed, b * p.H + h, t0 + lane, p.drop_p)) ? pv * keep_scale : 0.f;
          l2_run = l2_run * sc + warp_sum(pv);
        }
        m_run = m_new;
        const uint16_t hi = PmT<T>::bits(pv);
        p_hi[h * PP + lane] = hi;
        p_lo[h * PP + lane] = PmT<T>::bits(pv - PmT<T>::val(hi));
        if (lane == 0) s_scale[h] = sc;
      }
    } else {
      // ---- given weights: P[h][tok] = w[b, h, t0 + tok] ----
      if (warp < 8) {
        const int h = warp;
        float pv = 0.f;
        if (h < p.H && lane < rows) pv = p.w[b * p.wb + h * p.wh + t0 + lane];
        const uint16_t hi = PmT<T>::bits(pv);
        p_hi[h * PP + lane] = hi;
        p_lo[h * PP + lane] = PmT<T>::bits(pv - PmT<T>::val(hi));
      }
    }
    __syncthreads();

    // ---- phase 2: acc[d, h] += x^T P, warp owns channels [warp*DW, warp*DW + DW) ----
    if (softmax_mode) {
      const float s0 = s_scale[2 * t], s1 = s_scale[2 * t + 1];
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < nchunk) { acc[c][0] *= s0; acc[c][1] *= s1; acc[c][2] *= s0; acc[c][3] *= s1; }
    }
#pragma unroll
    for (int ks = 0; ks < PM_TT / 16; ++ks) {
      const uint32_t off = (uint32_t)(g * PP + ks * 16 + 2 * t) * 2;
      const uint32_t bh0 = lds32(ph_addr + off), bh1 = lds32(ph_addr + off + 16);
      const uint32_t bl0 = lds32(pl_addr + off), bl1 = lds32(pl_addr + off + 16);
      const int trow = ks * 16 + (lane & 7) + (lane >> 4) * 8;        // matrices 2,3: tokens +8
#pragma unroll
      for (int c4 = 0; c4 < MAXC; c4 += 4) {           // groups of 4 chunks: loads first, then 8 independent MMAs
        uint32_t a[4][4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c4 + c < nchunk)
            ldsm_x4_trans(a[c], tile_addr(tile, trow, ((warp * DW + (c4 + c) * 16) >> 3) + ((lane >> 3) & 1), row_bytes));   // matrices 1,3: d +8
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c4 + c < nchunk) PmT<T>::mma(acc[c4 + c], a[c], bl0, bl1);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c4 + c < nchunk) PmT<T>::mma(acc[c4 + c], a[c], bh0, bh1);
      }
    }
  }
  // ---- partial results of this (b, split): rows d = g / g+8 of each chunk, heads 2t / 2t+1 ----
  if (softmax_mode && warp < p.H && lane == 0 && p.part_m) {
    const size_t slot = ((size_t)b * p.S + sp) * p.H + warp;
    p.part_m[slot] = m_run;
    p.part_l[slot] = l_run;
    if (p.part_l2) p.part_l2[slot] = p.drop_p > 0.f ? l2_run : l_run;
  }
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    if (c < nchunk) {
      const int d = warp * DW + c * 16 + g;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int h = 2 * t + (e & 1), dd = d + (e >> 1) * 8;
        if (h < p.H) p.part_acc[(((size_t)b * p.S + sp) * p.H + h) * D + dd] = acc[c][e];
      }
    }
  }
}


struct PmBwdParams {
  const void* x; long long sb, sn;
  const unsigned char* mask; long long mb;
  const float* qt;        // [H, D]
  const float* dxbar;     // [B, H, D]
  const float* xbar;      // [B, H, D]
  const float* m; const float* l;   // [B, H]
  void* dx;               // [B, N, D] contiguous
  float* ds;              // [B, H, N]
  int B, N, D, H, S;      // S = token splits (gridDim.y)
  const float* sa; const float* dsa; float drop_p; unsigned long long drop_seed;
  int stages;
  const float* dlse;      // [B, H] upstream gradient of lse_h = m_h + log l_h (may be null)
  float* part_dq;         // kDq kernels: [B, S, H, D] per-(b, split) partials of dqt = sum_n ds_hn x_n (zeroed by the caller)
};

// shared memory: x / dx tiles 2 * 32 * D * 2 | qt hi, qt lo, dxbar hi, dxbar lo: 4 x [8][D+8] | Wt [D][24] 16-bit
//                (k-contiguous rows of [dxbar ; qt], 48-byte pitch) | C hi [32][16], C lo [32][16] |
//                S/T partials [4][32][16] fp32 | c, m, 1/l [3][8] fp32 | barriers | kDq: ds^T hi, lo [8][40] 16-bit
// kDq = true additionally accumulates dqt = sum_n ds_hn x_n from the SAME pass over x (the forward's phase 2 with P := ds,
// [D, 8] fp32 accumulator in registers, written as per-(b, split) partials): the separate "given weights" forward launch
// that re-reads all of x for the query gradient disappears (bwd reads x once instead of twice).
template <typename T, int NW, bool kDq = false>
__global__ void __launch_bounds__(NW * 32) pool_bwd_mma_kernel(const __grid_constant__ CUtensorMap tmx, PmBwdParams p) {
  constexpr int KS = NW / 2;
  constexpr int MAXC = 128 / NW;        // 16-channel chunks per warp at D = 1024 (kDq accumulator)
  constexpr int PP = PM_TT + 8;
  B2_DYN_SMEM(smem_raw);
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int D = p.D, row_bytes = D * 2, QP = D + 8;
  const uint32_t s_base = smem_u32(smem);
  const int NS = p.stages, tile_bytes = PM_TT * row_bytes;
  uint16_t* q_hi = reinterpret_cast<uint16_t*>(smem + (size_t)NS * tile_bytes);
  uint16_t* q_lo = q_hi + 8 * QP;
  uint16_t* d_hi = q_lo + 8 * QP;
  uint16_t* d_lo = d_hi + 8 * QP;
  uint16_t* wt = d_lo + 8 * QP;                          // [D][24]
  uint16_t* c_hi = wt + (size_t)D * 24;                  // [32][16]
  uint16_t* c_lo = c_hi + PM_TT * 16;
  float* st_part = reinterpret_cast<float*>(c_lo + PM_TT * 16);   // [4][32][16]
  float* s_c = st_part + KS * PM_TT * 16;
  float* s_m = s_c + 8;
  float* s_il = s_m + 8;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_il + 8);
  uint16_t* t_hi = reinterpret_cast<uint16_t*>(full_bar + 4);      // kDq only: ds^T [8][PP] hi / lo
  uint16_t* t_lo = t_hi + 8 * PP;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.x, sp = blockIdx.y;
  const int n0 = (int)((long long)p.N * sp / p.S), n1 = (int)((long long)p.N * (sp + 1) / p.S);
  const T* xb = reinterpret_cast<const T*>(p.x) + b * p.sb;
  T* dxb = reinterpret_cast<T*>(p.dx) + (size_t)b * p.N * D;
  if (n0 >= n1) return;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) mbar_init(&full_bar[i], 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmx);
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int i = 0; i < NS - 1; ++i) {
      const int tt = n0 + i * PM_TT;
      if (tt < n1) tma_x_tile(s_base + i * tile_bytes, &tmx, &full_bar[i], b * p.N + tt, D);
    }

  for (int i = threadIdx.x; i < 8 * D; i += NW * 32) {
    const int h = i / D, d = i - h * D;
    const float qv = h < p.H ? p.qt[(size_t)h * D + d] : 0.f;
    const float dv = h < p.H ? p.dxbar[((size_t)b * p.H + h) * D + d] : 0.f;
    const uint16_t qh = PmT<T>::bits(qv), dh = PmT<T>::bits(dv);
    q_hi[h * QP + d] = qh;
    q_lo[h * QP + d] = PmT<T>::bits(qv - PmT<T>::val(qh));
    d_hi[h * QP + d] = dh;
    d_lo[h * QP + d] = PmT<T>::bits(dv - PmT<T>::val(dh));
    wt[(size_t)d * 24 + h] = dh;           // k = h     : dxbar_h[d]  (multiplied by a_h)
    wt[(size_t)d * 24 + 8 + h] = qh;       // k = 8 + h : qt_h[d]     (multiplied by ds_h)
  }
  if (warp < 8) {
    const int h = warp;
    float c = 0.f;
    if (h < p.H)
      for (int d = lane; d < D; d += 32)
        c = fmaf(p.dxbar[((size_t)b * p.H + h) * D + d], p.xbar[((size_t)b * p.H + h) * D + d], c);
    c = warp_sum(c);
    if (p.dsa && h < p.H) c = fmaf(p.dsa[b * p.H + h], p.sa[b * p.H + h], c);
    if (p.dlse && h < p.H) c -= p.dlse[b * p.H + h];     // d lse / d s_n = a_n
    if (lane == 0) {
      s_c[h] = c;
      s_m[h] = h < p.H ? p.m[b * p.H + h] : 0.f;
      s_il[h] = h < p.H ? 1.f / p.l[b * p.H + h] : 0.f;
    }
  }
  const unsigned char* mk = p.mask ? p.mask + b * p.mb : nullptr;
  const uint32_t qh_addr = smem_u32(q_hi), ql_addr = smem_u32(q_lo), dh_addr = smem_u32(d_hi), dl_addr = smem_u32(d_lo);
  const uint32_t wt_addr = smem_u32(wt), ch_addr = smem_u32(c_hi), cl_addr = smem_u32(c_lo);
  const int DW = D / NW;
  const int nchunk = DW / 16;
  const uint32_t th_addr = smem_u32(t_hi), tl_addr = smem_u32(t_lo);
  float dq[kDq ? MAXC : 1][4];
#pragma unroll
  for (int c = 0; c < (kDq ? MAXC : 1); ++c) dq[c][0] = dq[c][1] = dq[c][2] = dq[c][3] = 0.f;

  int buf = 0;
  for (int t0 = n0; t0 < n1; t0 += PM_TT, buf = (buf + 1 == NS ? 0 : buf + 1)) {
    const int rows = min(PM_TT, n1 - t0);
    mbar_wait(&full_bar[buf], ((t0 - n0) / (PM_TT * NS)) & 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      const int tn = t0 + (NS - 1) * PM_TT, bn = (buf + NS - 1) % NS;
      if (tn < n1) {
        fence_proxy_async_smem();        // the previous tile's dx staging (generic writes) precedes the TMA overwrite
        tma_x_tile(s_base + bn * tile_bytes, &tmx, &full_bar[bn], b * p.N + tn, D);
      }
    }
    const uint32_t tile = s_base + buf * tile_bytes;
    // ---- phase 1: [S | T] partials, warp = (token group, K quarter) ----
    {
      const int tg = warp & 1, kq = warp >> 1;
      float cs[4] = {0.f, 0.f, 0.f, 0.f}, ct[4] = {0.f, 0.f, 0.f, 0.f}, csl[4] = {0.f, 0.f, 0.f, 0.f},
            ctl[4] = {0.f, 0.f, 0.f, 0.f};
      const int kbeg = kq * (D / KS), kend = kbeg + D / KS;
      const int arow = tg * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
      for (int k0 = kbeg; k0 < kend; k0 += 16) {
        uint32_t a[4];
        ldsm_x4(a, tile_addr(tile, arow, (k0 >> 3) + (lane >> 4), row_bytes));
        const uint32_t off = (uint32_t)(g * QP + k0 + 2 * t) * 2;
        const uint32_t ql0 = lds32(ql_addr + off), ql1 = lds32(ql_addr + off + 16), qh0 = lds32(qh_addr + off),
                       qh1 = lds32(qh_addr + off + 16);
        const uint32_t dl0 = lds32(dl_addr + off), dl1 = lds32(dl_addr + off + 16), dh0 = lds32(dh_addr + off),
                       dh1 = lds32(dh_addr + off + 16);
        PmT<T>::mma(csl, a, ql0, ql1);
        PmT<T>::mma(cs, a, qh0, qh1);
        PmT<T>::mma(ctl, a, dl0, dl1);
        PmT<T>::mma(ct, a, dh0, dh1);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { cs[e] += csl[e]; ct[e] += ctl[e]; }
      float* pp = st_part + ((size_t)kq * PM_TT + tg * 16) * 16;
      *reinterpret_cast<float2*>(pp + g * 16 + 2 * t) = make_float2(cs[0], cs[1]);
      *reinterpret_cast<float2*>(pp + (g + 8) * 16 + 2 * t) = make_float2(cs[2], cs[3]);
      *reinterpret_cast<float2*>(pp + g * 16 + 8 + 2 * t) = make_float2(ct[0], ct[1]);
      *reinterpret_cast<float2*>(pp + (g + 8) * 16 + 8 + 2 * t) = make_float2(ct[2], ct[3]);
    }
    __syncthreads();
    // ---- a, ds: warp = head, lane = token ----
    if (warp < 8) {
      const int h = warp;
      float sv = 0.f, tv = 0.f;
#pragma unroll
      for (int kq = 0; kq < KS; ++kq) {
        sv += st_part[((size_t)kq * PM_TT + lane) * 16 + h];
        tv += st_part[((size_t)kq * PM_TT + lane) * 16 + 8 + h];
      }
      const bool live = h < p.H && lane < rows && !(mk && mk[t0 + lane]);
      const float a0 = live ? __expf(sv - s_m[h]) * s_il[h] : 0.f;
      float kap = 1.f;
      if (p.drop_p > 0.f && h < p.H) kap = attn_keep(p.drop_seed, b * p.H + h, t0 + lane, p.drop_p) ? 1.f / (1.f - p.drop_p) : 0.f;
      const float dsv = a0 * (kap * (tv + ((p.dsa && h < p.H) ? p.dsa[b * p.H + h] : 0.f)) - s_c[h]);
      const float a = a0 * kap;
      if (h < p.H && lane < rows) p.ds[((size_t)b * p.H + h) * p.N + t0 + lane] = dsv;
      const uint16_t ah = PmT<T>::bits(a), dh2 = PmT<T>::bits(dsv);
      c_hi[lane * 16 + h] = ah;
      c_lo[lane * 16 + h] = PmT<T>::bits(a - PmT<T>::val(ah));
      c_hi[lane * 16 + 8 + h] = dh2;
      c_lo[lane * 16 + 8 + h] = PmT<T>::bits(dsv - PmT<T>::val(dh2));
      if (kDq) {                              // the same ds, token-contiguous per head: B operand of the dq product
        t_hi[h * PP + lane] = dh2;
        t_lo[h * PP + lane] = PmT<T>::bits(dsv - PmT<T>::val(dh2));
      }
    }
    __syncthreads();
    if (kDq) {
      // ---- phase 2b: dq[d, h] += sum_tok x[tok, d] ds[h, tok] (the forward's phase 2 with P := ds); the warp reads only
      //      its own channels [warp*DW, +DW) of the x tile, the ones it overwrites with dx right below ----
#pragma unroll
      for (int ks = 0; ks < PM_TT / 16; ++ks) {
        const uint32_t off = (uint32_t)(g * PP + ks * 16 + 2 * t) * 2;
        const uint32_t bh0 = lds32(th_addr + off), bh1 = lds32(th_addr + off + 16);
        const uint32_t bl0 = lds32(tl_addr + off), bl1 = lds32(tl_addr + off + 16);
        const int trow = ks * 16 + (lane & 7) + (lane >> 4) * 8;
#pragma unroll
        for (int c = 0; c < (kDq ? MAXC : 1); ++c) {
          if (c < nchunk) {
            uint32_t a[4];
            ldsm_x4_trans(a, tile_addr(tile, trow, ((warp * DW + c * 16) >> 3) + ((lane >> 3) & 1), row_bytes));
            PmT<T>::mma(dq[c], a, bl0, bl1);
            PmT<T>::mma(dq[c], a, bh0, bh1);
          }
        }
      }
      __syncwarp();                           // every lane's reads of the tile precede the dx staging stores of the warp
    }
    // ---- phase 2: dx[tok, d] = C . Wt^T, warp owns channels [warp*DW, +DW); staged into the (now free) x tile ----
#pragma unroll
    for (int tg = 0; tg < 2; ++tg) {
      uint32_t ah[4], al[4];
      const uint32_t o0 = (uint32_t)((tg * 16 + g) * 16 + 2 * t) * 2, o1 = o0 + 8 * 16 * 2;
      ah[0] = lds32(ch_addr + o0); ah[1] = lds32(ch_addr + o1); ah[2] = lds32(ch_addr + o0 + 16); ah[3] = lds32(ch_addr + o1 + 16);
      al[0] = lds32(cl_addr + o0); al[1] = lds32(cl_addr + o1); al[2] = lds32(cl_addr + o0 + 16); al[3] = lds32(cl_addr + o1 + 16);
      for (int d0 = warp * DW; d0 < warp * DW + DW; d0 += 8) {
        const uint32_t wo = wt_addr + (uint32_t)(d0 + g) * 48 + 4 * t;
        const uint32_t b0 = lds32(wo), b1 = lds32(wo + 16);
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        PmT<T>::mma(c, al, b0, b1);
        PmT<T>::mma(c, ah, b0, b1);
        // (tok = tg*16 + g [+8], d = d0 + 2t, 2t+1) -> 16-bit pairs into the swizzled tile
        const uint32_t v0 = (uint32_t)PmT<T>::bits(c[0]) | ((uint32_t)PmT<T>::bits(c[1]) << 16);
        const uint32_t v1 = (uint32_t)PmT<T>::bits(c[2]) | ((uint32_t)PmT<T>::bits(c[3]) << 16);
        const int r0 = tg * 16 + g;
        sts32(tile_addr(tile, r0, d0 >> 3, row_bytes) + 4 * t, v0);
        sts32(tile_addr(tile, r0 + 8, d0 >> 3, row_bytes) + 4 * t, v1);
      }
    }
    __syncthreads();
    {
      const int cpr = D / 8;
      for (int i = threadIdx.x; i < rows * cpr; i += NW * 32) {
        const int r = i / cpr, c = i - r * cpr;
        const uint4 v = lds128v(tile_addr(tile, r, c, row_bytes));
        *reinterpret_cast<uint4*>(dxb + (size_t)(t0 + r) * D + c * 8) = v;
      }
    }
  }
  if (kDq && p.part_dq) {
    // rows d = g / g + 8 of each chunk, heads 2t / 2t + 1 (the accumulator layout of the forward's phase 2)
#pragma unroll
    for (int c = 0; c < (kDq ? MAXC : 1); ++c) {
      if (c < nchunk) {
        const int d = warp * DW + c * 16 + g;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int h = 2 * t + (e & 1), dd = d + (e >> 1) * 8;
          if (h < p.H) p.part_dq[(((size_t)b * p.S + sp) * p.H + h) * D + dd] = dq[c][e];
        }
      }
    }
  }
}

}  // namespace b2
