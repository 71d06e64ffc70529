// K3: fused logits backward (tile recompute) for CLIP / gated-CLIP / SigLIP.
//
//   dX[i, :] += out_scale * sum_j G_ij * Yhat[j, :],   G_ij = epi(S_ij),  S = Xhat Yhat^T  (recomputed)
//
// Called twice per loss backward with the operand roles swapped: (X,Y) = (video,text) gives dVhat,
// (X,Y) = (text,video) gives dThat; the row/column scale vectors swap with them.
// The N x N matrices S, P and G exist only as 128 x 128 tiles in TMEM:
//
//   work item = (128-row X tile xt, segment seg of the Y tiles); inside an item the CTA walks the flattened step
//               sequence (dp, j): output column slice dp (256 wide) x Y tile j of the segment
//   smem      = [X panel of the item, resident: Kp/64 chunks of 128 x 64 bf16 (kXRes, Kp <= 512)] + 6-slot TMA ring of
//               Y chunks (16 KB each). kXRes = false (Kp > 512: D = 768, bf16x3 panels) streams X chunks through the
//               ring next to the Y chunks (32 KB slots) like the first version of this kernel did.
//   TMEM      = [0,256) fp32 accumulator of the current dp | [256,384) S/G buffer 0 | [384,512) S/G buffer 1
//   per step:  S_ij  <- tcgen05.mma SS (M128 N128 K=Kp)
//              G_ij  <- epilogue warps: tcgen05.ld S, elementwise gradient, pack bf16x2, tcgen05.st IN PLACE
//              acc   += tcgen05.mma TS: A = G_ij straight from TMEM, B = Yhat_j[:, slice] as an MN-major
//                       SWIZZLE_128B smem operand (the same [128 x 64] TMA boxes the S product uses)
//   The MMA warp issues S(step+1) before out(step), so the tensor pipe works while the epilogue of a step runs.
//   End of a dp sweep: accumulator -> registers -> red.global.add.v4.f32 into dX (caller zeroes dX).
//
// Measured machine limits that shaped this (tools/ubench, B200): a tcgen05.mma issued from a warp-uniform branch by
// one elected lane retires at its floor (SS N=128: 64 cycles, TS N=64: 32 cycles per K=16); issued under `lane == 0`
// the compiler wraps each UTCHMMA in a uniform-register waterfall loop costing ~112 cycles per instruction.
// TMA L2 -> smem streams at 61 B/cycle/SM with all 148 CTAs running, so the X panel must stay resident: streaming
// it again for every Y tile (320 KB per step) is L2-bound at ~5200 cycles per step versus 3072 tensor cycles.
//
// The diagonal (-Y_ij / N) part of the CLIP gradient and the SigLIP positives are NOT handled here: they are
// rank-sparse and are added analytically by l2norm_bwd / siglip_pos kernels, so this dense kernel never tests
// i == j except to subtract the target before the bf16 rounding.  Scalars (sum G*f(S) for dlog_temp, sum softplus
// for the SigLIP loss, sum G for dbias) are reduced per item during the dp == 0 sweep only.
#include "bwd_common.cuh"
#include "host_api.h"
#include <stdlib.h>

namespace b2 {

template <int kMode, bool kXRes>
__global__ void __launch_bounds__(BW_THREADS, 1)
bw_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, BwParams p) {
  using SM = BwSmem<kXRes>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = smem;                                  // resident X panel (kXRes)
  uint8_t* ring = smem + SM::kRingOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
  uint64_t* full_bar = bars;                       // [6]  TMA -> MMA
  uint64_t* empty_bar = bars + BW_SLOTS;           // [6]  MMA -> TMA
  uint64_t* sfull_bar = bars + 2 * BW_SLOTS;       // [2]  S tile ready      (MMA -> epilogue)
  uint64_t* gready_bar = sfull_bar + 2;            // [2]  G tile written    (epilogue -> MMA), 8 arrivals
  uint64_t* accfull_bar = gready_bar + 2;          // [1]  accumulator ready (MMA -> epilogue)
  uint64_t* accempty_bar = accfull_bar + 1;        // [1]  accumulator drained (epilogue -> MMA), 8 arrivals
  uint64_t* xfull_bar = accempty_bar + 1;          // [1]  X panel landed    (TMA -> MMA)
  uint64_t* xempty_bar = xfull_bar + 1;            // [1]  X panel free      (MMA -> TMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xempty_bar + 1);
  float* col_s = reinterpret_cast<float*>(smem + SM::kColOff);   // [2][128]

  // warp index through a shuffle: role branches are provably warp-uniform (see the header comment)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int kchunks = p.Kp / BW_BK;
  const int items = p.x_tiles * p.nseg;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < BW_SLOTS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sfull_bar[s], 1);
      mbar_init(&gready_bar[s], 8);
    }
    mbar_init(accfull_bar, 1);
    mbar_init(accempty_bar, 8);
    mbar_init(xfull_bar, 1);
    mbar_init(xempty_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_col = 0, s_col0 = 256;
  if (p.dyn) {   // temperature / bias live on the device: no host sync
    p.scale2 = p.dyn[0];
    p.shift2 = p.dyn[1];
    p.inv_tau = p.dyn[2];
    p.bias = p.dyn[5];
    p.out_scale = p.dyn[2];
    p.lclamp = p.dyn[8];
    p.yneg = p.dyn[9];
    p.ent_coef = p.dyn[10];
    p.stable = !BwIsSiglip<kMode>::value && p.dyn[11] != 0.f;
  }

  // item decode (identical in every role): item -> X tile, Y tile range [j0, j1)
  auto decode = [&](int item, int& xt, int& j0, int& j1) {
    const int seg = item % p.nseg;
    xt = item / p.nseg;
    j0 = (int)((long long)p.y_tiles * seg / p.nseg);
    j1 = (int)((long long)p.y_tiles * (seg + 1) / p.nseg);
  };
  auto n_dchunks = [&](int dp) {
    const int rem = (p.Dp - dp * BW_DP) / BW_BK;
    return rem < 4 ? rem : 4;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int slot = 0;
      uint32_t phase = 0, xphase = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int xt, j0, j1;
        decode(item, xt, j0, j1);
        const int nj = j1 - j0;
        if (nj <= 0) continue;
        const int T = nj * p.dparts;
        if (kXRes) {
          mbar_wait(xempty_bar, xphase ^ 1);
          xphase ^= 1;
          mbar_expect_tx(xfull_bar, kchunks * BW_CHUNK);
          for (int kc = 0; kc < kchunks; ++kc) tma_load_2d(xs + kc * BW_CHUNK, &tmX, xfull_bar, kc * BW_BK, xt * BW_BM);
        }
        auto load_s = [&](int t) {
          const int j = j0 + t % nj;
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(&empty_bar[slot], phase ^ 1);
            uint8_t* sl = ring + slot * SM::kSlotBytes;
            mbar_expect_tx(&full_bar[slot], SM::kSlotBytes);
            if (!kXRes) tma_load_2d(sl + BW_CHUNK, &tmX, &full_bar[slot], kc * BW_BK, xt * BW_BM);
            tma_load_2d(sl, &tmY, &full_bar[slot], kc * BW_BK, j * BW_BN);
            if (++slot == BW_SLOTS) { slot = 0; phase ^= 1; }
          }
        };
        auto load_out = [&](int t) {
          const int dp = t / nj, j = j0 + t % nj;
          const int nd = n_dchunks(dp);
          for (int dc = 0; dc < nd; ++dc) {
            mbar_wait(&empty_bar[slot], phase ^ 1);
            mbar_expect_tx(&full_bar[slot], BW_CHUNK);
            tma_load_2d(ring + slot * SM::kSlotBytes, &tmY, &full_bar[slot], p.hi_off + dp * BW_DP + dc * BW_BK,
                        j * BW_BN);
            if (++slot == BW_SLOTS) { slot = 0; phase ^= 1; }
          }
        };
        load_s(0);
        for (int t = 0; t < T; ++t) {
          if (t + 1 < T) load_s(t + 1);
          load_out(t);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BW_BM, BW_BN, 0, 0);    // A, B K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(BW_BM, BW_BK, 0, 1);    // A = bf16 G from TMEM, B MN-major, N = 64
      int slot = 0;
      uint32_t phase = 0, xphase = 0;
      uint32_t tile_ctr = 0;      // S/G buffer = tile_ctr & 1, phase = (tile_ctr >> 1) & 1
      uint32_t acc_ctr = 0;
      const uint32_t xs_addr = smem_u32(xs);
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int xt, j0, j1;
        decode(item, xt, j0, j1);
        const int nj = j1 - j0;
        if (nj <= 0) continue;
        const int T = nj * p.dparts;
        if (kXRes) {
          mbar_wait(xfull_bar, xphase);
          xphase ^= 1;
          tc_fence_after();
        }
        auto mma_s = [&](uint32_t tc) {
          const uint32_t d_tmem = tmem_base + s_col0 + (tc & 1) * BW_BN;
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(&full_bar[slot], phase);
            tc_fence_after();
            const uint32_t sl = smem_u32(ring + slot * SM::kSlotBytes);
            const uint64_t bdesc = make_smem_desc_sw128(sl, 1024);
            const uint64_t adesc = make_smem_desc_sw128(kXRes ? xs_addr + kc * BW_CHUNK : sl + BW_CHUNK, 1024);
#pragma unroll
            for (int k = 0; k < BW_BK / 16; ++k) mma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc_s, (kc | k) != 0);
            tc_commit(&empty_bar[slot]);
            if (++slot == BW_SLOTS) { slot = 0; phase ^= 1; }
          }
          tc_commit(&sfull_bar[tc & 1]);
        };
        auto mma_out = [&](uint32_t tc, int dp, bool first) {
          const uint32_t g_tmem = tmem_base + s_col0 + (tc & 1) * BW_BN;
          const int nd = n_dchunks(dp);
          for (int dc = 0; dc < nd; ++dc) {
            mbar_wait(&full_bar[slot], phase);
            tc_fence_after();
            const uint32_t sy = smem_u32(ring + slot * SM::kSlotBytes);
            const uint32_t d_tmem = tmem_base + acc_col + dc * BW_BK;
            const uint64_t bdesc0 = make_smem_desc_sw128(sy, 1024);
#pragma unroll
            for (int ks = 0; ks < BW_BN / 16; ++ks) {
              // B: rows = K (Y rows 16ks..16ks+15), 64 output columns contiguous per 128-byte row (+2048 B per ks)
              const uint64_t bdesc = bdesc0 + uint64_t(ks * (2048 >> 4));
              // A: G bf16x2-packed; each epilogue half keeps its 64 K-values in its own 32 columns (hi), lo next 32
              const uint32_t a_tmem = g_tmem + (ks >> 2) * 64 + (ks & 3) * 8;
              mma_ts(d_tmem, a_tmem, bdesc, idesc_o, !(first && ks == 0));
              if (p.hp) mma_ts(d_tmem, a_tmem + 32, bdesc, idesc_o, 1u);
            }
            tc_commit(&empty_bar[slot]);
            if (++slot == BW_SLOTS) { slot = 0; phase ^= 1; }
          }
        };
        mma_s(tile_ctr);
        if (kXRes && T == 1) tc_commit(xempty_bar);
        for (int t = 0; t < T; ++t) {
          const int dp = t / nj, jr = t - dp * nj;
          if (t + 1 < T) {
            mma_s(tile_ctr + 1);
            if (kXRes && t + 2 == T) tc_commit(xempty_bar);   // last S product of the item issued: X panel is free
          }
          if (jr == 0) {
            // accumulator must have been drained by the epilogue of the previous dp sweep
            mbar_wait(accempty_bar, (acc_ctr & 1) ^ 1);
            tc_fence_after();
          }
          mbar_wait(&gready_bar[tile_ctr & 1], (tile_ctr >> 1) & 1);
          tc_fence_after();
          mma_out(tile_ctr, dp, jr == 0);
          ++tile_ctr;
          if (jr == nj - 1) {
            tc_commit(accfull_bar);
            ++acc_ctr;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps 4..11 =====================
    const int q = warp & 3;               // TMEM lane quarter
    const int wg = (warp - 4) >> 2;       // column half of the 128-wide S tile / 256-wide accumulator
    const int etid = threadIdx.x - 128;   // 0..255
    uint32_t tile_ctr = 0, acc_ctr = 0;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const float ydn = p.ydiag * p.gnorm, wn = p.wneg_c * p.gnorm, ign = 1.f / p.gnorm;
    const float nshift2 = -p.shift2;
    const uint32_t cs_base = smem_u32(col_s) + wg * 64 * 4;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int xt, j0, j1;
      decode(item, xt, j0, j1);
      const int nj = j1 - j0;
      if (nj <= 0) continue;
      const int T = nj * p.dparts;
      const int row = xt * BW_BM + q * 32 + lane;
      const bool row_ok = row < p.Nx;
      float rs = 0.f, ent_iz = 0.f, ent_m = 0.f;
      if (!BwIsSiglip<kMode>::value) rs = bw_stat(p, p.rowscale, row, row_ok);
      if (kMode == BW_SIGLIP_ENT && p.rowscale && row_ok) {
        ent_iz = p.rowscale[2 * row];
        ent_m = p.rowscale[2 * row + 1];
      }
      double dtacc = 0.0, dlacc = 0.0, dbacc = 0.0;
      for (int t = 0; t < T; ++t, ++tile_ctr) {
        const int dp = t / nj, jr = t - dp * nj, j = j0 + jr;
        const bool want_scal = p.scal != nullptr && dp == 0;
        float tacc = 0.f, lacc = 0.f, bacc = 0.f;      // per-tile fp32 partials, accumulated in fp64 across tiles
        const int buf = tile_ctr & 1;
        if (!BwIsSiglip<kMode>::value) {
          if (etid < 128) {
            const int col = j * BW_BN + etid;
            col_s[buf * 128 + etid] = bw_stat(p, p.colscale, col, col < p.Ny);
          }
          named_bar_sync(1, 256);
        }
        mbar_wait(&sfull_bar[buf], (tile_ctr >> 1) & 1);
        tc_fence_after();
        const uint32_t sbase = tmem_base + lane_off + s_col0 + buf * BW_BN + wg * 64;
        const uint32_t cs_addr = cs_base + buf * 128 * 4;
        {
          BwThread th{row, row_ok, wg, rs, ydn, wn, ign, nshift2, ent_iz, ent_m};
          bw_g_tile<kMode>(p, th, sbase, cs_addr, col_s + buf * 128, xt, j, dp, want_scal, tacc, lacc, bacc);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&gready_bar[buf]);
        dtacc += (double)tacc;
        if (BwIsSiglip<kMode>::value) {
          dlacc += (double)lacc;
          dbacc += (double)bacc;
        }
        if (jr == nj - 1) {
          // ---- drain the accumulator of this dp sweep ----
          mbar_wait(accfull_bar, acc_ctr & 1);
          ++acc_ctr;
          tc_fence_after();
          {
            BwThread th{row, row_ok, wg, rs, ydn, wn, ign, nshift2, ent_iz, ent_m};
            bw_drain(p, th, tmem_base + lane_off + acc_col + wg * 128, dp);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(accempty_bar);
        }
      }
      if (p.scal) {
        for (int o = 16; o > 0; o >>= 1) {
          dtacc += __shfl_xor_sync(0xffffffffu, dtacc, o);
          if (BwIsSiglip<kMode>::value) {
            dlacc += __shfl_xor_sync(0xffffffffu, dlacc, o);
            dbacc += __shfl_xor_sync(0xffffffffu, dbacc, o);
          }
        }
        if (lane == 0) {
          atomicAdd(p.scal + 0, dtacc * (double)ign);
          if (BwIsSiglip<kMode>::value) {
            atomicAdd(p.scal + 1, dlacc);
            atomicAdd(p.scal + 2, dbacc * (double)ign);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace b2

namespace b2host {
using namespace b2;

template <int kMode, bool kXRes>
static int launch_bw(const CUtensorMap& tmX, const CUtensorMap& tmY, const BwParams& p, int grid, cudaStream_t stream) {
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device() & 63];   // cudaFuncSetAttribute is per device
  constexpr int smem = BwSmem<kXRes>::kBytes;
  if (!attr_done) {
    if (cudaFuncSetAttribute(bw_kernel<kMode, kXRes>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return B2_ECUDA;
    attr_done = true;
  }
  bw_kernel<kMode, kXRes><<<grid, BW_THREADS, smem, stream>>>(tmX, tmY, p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int logits_bwd(int mode, const void* X, const void* Y, int Nx, int Ny, int Kp, int Dp, int D, int hi_off, int ldx,
               int ldy,
               float scale2, float shift2, float inv_tau, float bias, float wneg_c, const float* rowscale,
               const float* colscale, float out_scale, float gnorm, int hp, const float* dyn, float ydiag,
               int diag_off, float* diag_corr,
               float* dX, int ldd, double* scal, int nseg_hint, cudaStream_t stream) {
  if (Nx <= 0 || Ny <= 0 || Kp <= 0 || Kp % 64 || Dp <= 0 || Dp % 64 || hi_off < 0 || hi_off % 64 ||
      hi_off + Dp > Kp || D > Dp || D <= 0)
    return B2_EINVAL;
  if (mode != BW_SIGLIP && mode != BW_SIGLIP_ENT && (!rowscale || !colscale)) return B2_EINVAL;
  if (mode == BW_SIGLIP_ENT && ((rowscale != nullptr) == (colscale != nullptr) || !dyn)) return B2_EINVAL;
  const bool ent = mode == BW_SIGLIP_ENT;   // entropy term: single-CTA kernel only (general epilogue path)
  // headline shapes (plain bf16 operands, D <= 512): CTA-pair kernel, half the shared-memory ingest per SM.
  // B200CLIP_BWD_PAIR=0 keeps the single-CTA kernel (A/B measurements).
  // 64-row CTA pairs with the whole output width in TMEM (logits_bwd3.cu): no S recompute. Default whenever the shape
  // qualifies (plain bf16 operands, Dp in {256, 512, 768}): measured 1.54 vs 1.98 ms (D = 512) and 2.28 vs 5.50 ms
  // (D = 768) per launch at 32k x 32k. B200CLIP_BWD3=0 keeps the 128-row kernels (A/B measurements).
  static const bool bw3_on = [] { const char* e = getenv("B200CLIP_BWD3"); return !(e && e[0] == '0'); }();
  if (!ent && bw3_on) {
    const int rc = logits_bwd_pair64(mode, X, Y, Nx, Ny, Kp, Dp, D, hi_off, ldx, ldy, scale2, shift2, inv_tau, bias,
                                     wneg_c, rowscale, colscale, out_scale, gnorm, hp, dyn, ydiag, diag_off, diag_corr,
                                     dX, ldd, scal, nseg_hint, stream);
    if (rc != B2_ENOSYS) return rc;
  }
  static const bool pair_ok = [] { const char* e = getenv("B200CLIP_BWD_PAIR"); return !(e && e[0] == '0'); }();
  if (!ent && pair_ok && Kp <= BW_XRES_CHUNKS * BW_BK && !hp && Dp % 128 == 0 && sm_count() >= 2)
    return logits_bwd_pair(mode, X, Y, Nx, Ny, Kp, Dp, D, hi_off, ldx, ldy, scale2, shift2, inv_tau, bias, wneg_c,
                           rowscale, colscale, out_scale, gnorm, hp, dyn, ydiag, diag_off, diag_corr, dX, ldd, scal,
                           nseg_hint, stream);
  BwParams p;
  p.Nx = Nx; p.Ny = Ny; p.Kp = Kp; p.Dp = Dp; p.D = D; p.hi_off = hi_off; p.ydiag = ydiag; p.diag_off = diag_off; p.diag_corr = diag_corr;
  p.x_tiles = (Nx + BW_BM - 1) / BW_BM;
  p.y_tiles = (Ny + BW_BN - 1) / BW_BN;
  p.dparts = (Dp + BW_DP - 1) / BW_DP;
  const int sms = sm_count();
  int nseg = nseg_hint;
  if (nseg <= 0) {
    // items = x_tiles * nseg dealt round-robin to the CTAs: pick the segment count (never shorter than 8 Y tiles)
    // that minimises the last-wave tail; ties go to fewer segments (fewer X panel loads and dX atomics)
    const int max_seg = p.y_tiles / 8 > 1 ? p.y_tiles / 8 : 1;
    double best = 1e30;
    nseg = 1;
    for (int s = 1; s <= max_seg && s <= 64; ++s) {
      const long long it = (long long)p.x_tiles * s;
      const long long waves = (it + sms - 1) / sms;
      const double cost = (double)waves * ((p.y_tiles + s - 1) / s + 1.5);   // +1.5 tiles of per-item overhead
      if (cost < best * 0.995) { best = cost; nseg = s; }
    }
  }
  if (nseg > p.y_tiles) nseg = p.y_tiles;
  p.nseg = nseg;
  p.scale2 = scale2; p.shift2 = shift2; p.inv_tau = inv_tau; p.bias = bias; p.wneg_c = wneg_c;
  p.rowscale = rowscale; p.colscale = colscale; p.out_scale = out_scale;
  p.gnorm = gnorm > 0.f ? gnorm : 1.f;
  p.hp = hp ? 1 : 0;
  p.lclamp = 30.f; p.yneg = 0.f; p.ent_coef = 0.f; p.stable = 0;
  p.dX = dX; p.ldd = ldd; p.scal = scal; p.dyn = dyn;
  CUtensorMap tmX, tmY;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmX, X, Nx, Kp, ldx, BW_BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmY, Y, Ny, Kp, ldy, BW_BN))) return rc;
  const int items = p.x_tiles * p.nseg;
  const int grid = items < sms ? items : sms;
  const bool xres = Kp <= BW_XRES_CHUNKS * BW_BK;
#define LAUNCH(M) (xres ? launch_bw<M, true>(tmX, tmY, p, grid, stream) : launch_bw<M, false>(tmX, tmY, p, grid, stream))
  if (mode == BW_CLIP) return LAUNCH(BW_CLIP);
  if (mode == BW_GATED) return LAUNCH(BW_GATED);
  if (mode == BW_SIGLIP) return LAUNCH(BW_SIGLIP);
  if (mode == BW_SIGLIP_ENT) return LAUNCH(BW_SIGLIP_ENT);
#undef LAUNCH
  return B2_EINVAL;
}

}  // namespace b2host
