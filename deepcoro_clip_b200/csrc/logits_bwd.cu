// K3: fused logits backward (tile recompute) for CLIP / gated-CLIP / SigLIP.
//
//   dX[i, :] += out_scale * sum_j G_ij * Yhat[j, :],   G_ij = epi(S_ij),  S = Xhat Yhat^T  (recomputed)
//
// Called twice per loss backward with the operand roles swapped: (X,Y) = (video,text) gives dVhat,
// (X,Y) = (text,video) gives dThat; the row/column scale vectors swap with them.
// The N x N matrices S, P and G exist only as 128 x 128 tiles in TMEM:
//
//   work item = (128-row X tile i, 256-column slice dp of the output, segment seg of the Y tiles)
//   TMEM      = [0,256) fp32 accumulator of the item | [256,384) S/G buffer 0 | [384,512) S/G buffer 1
//   per Y tile j:  S_ij  <- tcgen05.mma SS (M128 N128 K=Kp), operands streamed by TMA (6-slot ring)
//                  G_ij  <- epilogue warps: tcgen05.ld S, elementwise gradient, pack bf16x2, tcgen05.st IN PLACE
//                  acc   += tcgen05.mma TS: A = G_ij straight from TMEM, B = Yhat_j[:, slice] as an MN-major
//                           SWIZZLE_128B smem operand (the same [128 x 64] TMA boxes the S product uses)
//   The MMA warp issues S(j+1) before out(j), so the tensor pipe works while the epilogue of tile j runs.
//   Item end: accumulator -> registers -> red.global.add.f32 into dX (caller zeroes dX).
//
// The diagonal (-Y_ij / N) part of the CLIP gradient and the SigLIP positives are NOT handled here: they are
// rank-sparse and are added analytically by l2norm_bwd / siglip_pos kernels, so this dense kernel never tests
// i == j.  Scalars (sum G*f(S) for dlog_temp, sum softplus for the SigLIP loss, sum G for dbias) are reduced per
// item and atomically added to `scal` by the dp == 0 items only.
#include "common.cuh"
#include "host_api.h"

namespace b2 {

constexpr int BW_BM = 128;     // X rows per item
constexpr int BW_BN = 128;     // Y rows per tile
constexpr int BW_BK = 64;
constexpr int BW_DP = 256;     // output columns per item
constexpr int BW_SLOTS = 6;
constexpr int BW_CHUNK = BW_BM * BW_BK * 2;           // 16 KB: [128 rows x 64 bf16]
constexpr int BW_SLOT_BYTES = 2 * BW_CHUNK;           // X chunk + Y chunk
constexpr int BW_THREADS = 384;
constexpr int BW_SMEM_BYTES = BW_SLOTS * BW_SLOT_BYTES + 1024 + 256 + 2 * 128 * 4;

enum { BW_CLIP = 0, BW_GATED = 1, BW_SIGLIP = 2 };

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
  float inv_tau, bias, wneg_c; // SigLIP: R = S*inv_tau + bias ; G = wneg_c * sigmoid(clamp R) * [|R|<=30]
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
};

template <int kMode>
__global__ void __launch_bounds__(BW_THREADS, 1)
bw_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, BwParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BW_SLOTS * BW_SLOT_BYTES);
  uint64_t* full_bar = bars;                       // [6]  TMA -> MMA
  uint64_t* empty_bar = bars + BW_SLOTS;           // [6]  MMA -> TMA
  uint64_t* sfull_bar = bars + 2 * BW_SLOTS;       // [2]  S tile ready      (MMA -> epilogue)
  uint64_t* gready_bar = sfull_bar + 2;            // [2]  G tile written    (epilogue -> MMA), 8 arrivals
  uint64_t* accfull_bar = gready_bar + 2;          // [1]  accumulator ready (MMA -> epilogue)
  uint64_t* accempty_bar = accfull_bar + 1;        // [1]  accumulator drained (epilogue -> MMA), 8 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty_bar + 1);
  float* col_s = reinterpret_cast<float*>(smem + BW_SLOTS * BW_SLOT_BYTES + 256);   // [2][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kchunks = p.Kp / BW_BK;
  const int items = p.x_tiles * p.dparts * p.nseg;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
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
  }

  // item decode (identical in every role)
  auto decode = [&](int item, int& xt, int& dp, int& j0, int& j1) {
    const int seg = item % p.nseg;
    const int r = item / p.nseg;
    dp = r % p.dparts;
    xt = r / p.dparts;
    j0 = (int)((long long)p.y_tiles * seg / p.nseg);
    j1 = (int)((long long)p.y_tiles * (seg + 1) / p.nseg);
  };
  auto n_dchunks = [&](int dp) {
    const int rem = (p.Dp - dp * BW_DP) / BW_BK;
    return rem < 4 ? rem : 4;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      auto load_s = [&](int xt, int j) {
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&empty_bar[slot], phase ^ 1);
          uint8_t* sx = smem + slot * BW_SLOT_BYTES;
          mbar_expect_tx(&full_bar[slot], BW_SLOT_BYTES);
          tma_load_2d(sx, &tmX, &full_bar[slot], kc * BW_BK, xt * BW_BM);
          tma_load_2d(sx + BW_CHUNK, &tmY, &full_bar[slot], kc * BW_BK, j * BW_BN);
          if (++slot == BW_SLOTS) { slot = 0; phase ^= 1; }
        }
      };
      auto load_out = [&](int dp, int j) {
        const int nd = n_dchunks(dp);
        for (int dc = 0; dc < nd; ++dc) {
          mbar_wait(&empty_bar[slot], phase ^ 1);
          uint8_t* sy = smem + slot * BW_SLOT_BYTES;
          mbar_expect_tx(&full_bar[slot], BW_CHUNK);
          tma_load_2d(sy, &tmY, &full_bar[slot], p.hi_off + dp * BW_DP + dc * BW_BK, j * BW_BN);
          if (++slot == BW_SLOTS) { slot = 0; phase ^= 1; }
        }
      };
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int xt, dp, j0, j1;
        decode(item, xt, dp, j0, j1);
        if (j0 >= j1) continue;
        load_s(xt, j0);
        for (int j = j0; j < j1; ++j) {
          if (j + 1 < j1) load_s(xt, j + 1);
          load_out(dp, j);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BW_BM, BW_BN, 0, 0);    // A, B K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(BW_BM, BW_BK, 0, 1);    // A = bf16 G from TMEM, B MN-major, N = 64
      int slot = 0;
      uint32_t phase = 0;
      uint32_t tile_ctr = 0;      // S/G buffer = tile_ctr & 1, phase = (tile_ctr >> 1) & 1
      uint32_t item_ctr = 0;
      auto mma_s = [&](uint32_t tc) {
        const uint32_t d_tmem = tmem_base + s_col0 + (tc & 1) * BW_BN;
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&full_bar[slot], phase);
          tc_fence_after();
          const uint32_t sx = smem_u32(smem + slot * BW_SLOT_BYTES);
          const uint64_t adesc = make_smem_desc_sw128(sx, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(sx + BW_CHUNK, 1024);
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
          const uint32_t sy = smem_u32(smem + slot * BW_SLOT_BYTES);
          const uint32_t d_tmem = tmem_base + acc_col + dc * BW_BK;
#pragma unroll
          for (int ks = 0; ks < BW_BN / 16; ++ks) {
            // B: rows = K (Y rows 16ks..16ks+15), 64 output columns contiguous per 128-byte row
            const uint64_t bdesc = make_smem_desc_sw128(sy + ks * 2048, 1024);
            // A: G bf16x2-packed; each epilogue half keeps its 64 K-values in its own 32 columns (hi), lo next 32
            const uint32_t a_tmem = g_tmem + (ks >> 2) * 64 + (ks & 3) * 8;
            mma_ts(d_tmem, a_tmem, bdesc, idesc_o, !(first && ks == 0));
            if (p.hp) mma_ts(d_tmem, a_tmem + 32, bdesc, idesc_o, 1u);
          }
          tc_commit(&empty_bar[slot]);
          if (++slot == BW_SLOTS) { slot = 0; phase ^= 1; }
        }
      };
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        int xt, dp, j0, j1;
        decode(item, xt, dp, j0, j1);
        if (j0 >= j1) continue;
        mma_s(tile_ctr);
        for (int j = j0; j < j1; ++j) {
          if (j + 1 < j1) mma_s(tile_ctr + 1);
          if (j == j0) {
            // accumulator must have been drained by the epilogue of the previous item
            mbar_wait(accempty_bar, (item_ctr & 1) ^ 1);
            tc_fence_after();
          }
          mbar_wait(&gready_bar[tile_ctr & 1], (tile_ctr >> 1) & 1);
          tc_fence_after();
          mma_out(tile_ctr, dp, j == j0);
          ++tile_ctr;
        }
        tc_commit(accfull_bar);
        ++item_ctr;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps 4..11 =====================
    const int q = warp & 3;               // TMEM lane quarter
    const int wg = (warp - 4) >> 2;       // column half of the 128-wide S tile / 256-wide accumulator
    const int etid = threadIdx.x - 128;   // 0..255
    uint32_t tile_ctr = 0, item_ctr = 0;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      int xt, dp, j0, j1;
      decode(item, xt, dp, j0, j1);
      if (j0 >= j1) continue;
      const int row = xt * BW_BM + q * 32 + lane;
      const bool row_ok = row < p.Nx;
      float rs = 0.f;
      if (kMode != BW_SIGLIP) rs = row_ok ? p.rowscale[row] * p.gnorm : 0.f;
      const float ydn = p.ydiag * p.gnorm, wn = p.wneg_c * p.gnorm, ign = 1.f / p.gnorm;
      double dtacc = 0.0, dlacc = 0.0, dbacc = 0.0;
      for (int j = j0; j < j1; ++j, ++tile_ctr) {
        float tacc = 0.f, lacc = 0.f, bacc = 0.f;      // per-tile fp32 partials, accumulated in fp64 across tiles
        const int buf = tile_ctr & 1;
        float* cs = col_s + buf * 128;
        if (kMode != BW_SIGLIP) {
          if (etid < 128) {
            const int col = j * BW_BN + etid;
            cs[etid] = col < p.Ny ? p.colscale[col] * p.gnorm : 0.f;
          }
          named_bar_sync(1, 256);
        }
        const bool full = (xt * BW_BM + BW_BM <= p.Nx) && (j * BW_BN + BW_BN <= p.Ny);
        // does the target diagonal cross this tile? (block-uniform)
        const int dlo = xt * BW_BM + p.diag_off - j * BW_BN;
        const bool has_diag = kMode != BW_SIGLIP && p.ydiag != 0.f && dlo > -BW_BM && dlo < BW_BN;
        const int dcol = row + p.diag_off - j * BW_BN - wg * 64;   // diagonal column relative to this thread's half
        mbar_wait(&sfull_bar[buf], (tile_ctr >> 1) & 1);
        tc_fence_after();
        const uint32_t sbase = tmem_base + lane_off + s_col0 + buf * BW_BN + wg * 64;
        uint32_t acc[2][32];
        tmem_ld32(sbase, acc[0]);
        tmem_ld32(sbase + 32, acc[1]);
        tc_wait_ld();
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
              if (kMode == BW_SIGLIP) {
                const float R = fmaf(s, p.inv_tau, p.bias);
                const float Lc = fminf(fmaxf(R, -30.f), 30.f);
                const float ex = ex2_approx(-1.4426950408889634f * fabsf(Lc));
                const float den = 1.f + ex;
                const float r = __fdividef(1.f, den);
                const float sig = Lc >= 0.f ? r : ex * r;
                g = (fabsf(R) <= 30.f) ? wn * sig : 0.f;
                float sp = fmaxf(Lc, 0.f) + 0.6931471805599453f * lg2_approx(den);
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
                const float pr = ex2_approx(fmaf(f, p.scale2, -p.shift2));
                g = pr * (rs + cs[cl]);
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
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&gready_bar[buf]);
        dtacc += (double)tacc;
        if (kMode == BW_SIGLIP) {
          dlacc += (double)lacc;
          dbacc += (double)bacc;
        }
      }
      // ---- drain the accumulator of this item ----
      mbar_wait(accfull_bar, item_ctr & 1);
      tc_fence_after();
      {
        const uint32_t abase = tmem_base + lane_off + acc_col + wg * 128;
        float* drow = p.dX + (size_t)row * p.ldd + dp * BW_DP + wg * 128;
        const int cvalid = p.D - (dp * BW_DP + wg * 128);     // valid columns in this half
        uint32_t a[32];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c * 32 < cvalid) {     // warp-uniform
            tmem_ld32(abase + c * 32, a);
            tc_wait_ld();
            if (row_ok) {
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (c * 32 + e < cvalid) atomicAdd(drow + c * 32 + e, __uint_as_float(a[e]) * (p.out_scale * ign));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accempty_bar);
      ++item_ctr;
      if (p.scal && dp == 0) {
        for (int o = 16; o > 0; o >>= 1) {
          dtacc += __shfl_xor_sync(0xffffffffu, dtacc, o);
          if (kMode == BW_SIGLIP) {
            dlacc += __shfl_xor_sync(0xffffffffu, dlacc, o);
            dbacc += __shfl_xor_sync(0xffffffffu, dbacc, o);
          }
        }
        if (lane == 0) {
          atomicAdd(p.scal + 0, dtacc * (double)ign);
          if (kMode == BW_SIGLIP) {
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

int logits_bwd(int mode, const void* X, const void* Y, int Nx, int Ny, int Kp, int Dp, int D, int hi_off, int ldx,
               int ldy,
               float scale2, float shift2, float inv_tau, float bias, float wneg_c, const float* rowscale,
               const float* colscale, float out_scale, float gnorm, int hp, const float* dyn, float ydiag,
               int diag_off, float* diag_corr,
               float* dX, int ldd, double* scal, int nseg_hint, cudaStream_t stream) {
  if (Nx <= 0 || Ny <= 0 || Kp <= 0 || Kp % 64 || Dp <= 0 || Dp % 64 || hi_off < 0 || hi_off % 64 ||
      hi_off + Dp > Kp || D > Dp || D <= 0)
    return B2_EINVAL;
  if (mode != BW_SIGLIP && (!rowscale || !colscale)) return B2_EINVAL;
  BwParams p;
  p.Nx = Nx; p.Ny = Ny; p.Kp = Kp; p.Dp = Dp; p.D = D; p.hi_off = hi_off; p.ydiag = ydiag; p.diag_off = diag_off; p.diag_corr = diag_corr;
  p.x_tiles = (Nx + BW_BM - 1) / BW_BM;
  p.y_tiles = (Ny + BW_BN - 1) / BW_BN;
  p.dparts = (Dp + BW_DP - 1) / BW_DP;
  const int sms = sm_count();
  int nseg = nseg_hint;
  if (nseg <= 0) {
    // enough items for >= 6 waves (tail < ~15 %) but never segments shorter than 8 Y tiles
    const int base = p.x_tiles * p.dparts;
    nseg = (6 * sms + base - 1) / base;
    const int max_seg = p.y_tiles / 8 > 1 ? p.y_tiles / 8 : 1;
    if (nseg > max_seg) nseg = max_seg;
    if (nseg < 1) nseg = 1;
  }
  if (nseg > p.y_tiles) nseg = p.y_tiles;
  p.nseg = nseg;
  p.scale2 = scale2; p.shift2 = shift2; p.inv_tau = inv_tau; p.bias = bias; p.wneg_c = wneg_c;
  p.rowscale = rowscale; p.colscale = colscale; p.out_scale = out_scale;
  p.gnorm = gnorm > 0.f ? gnorm : 1.f;
  p.hp = hp ? 1 : 0;
  p.dX = dX; p.ldd = ldd; p.scal = scal; p.dyn = dyn;
  CUtensorMap tmX, tmY;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmX, X, Nx, Kp, ldx, BW_BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmY, Y, Ny, Kp, ldy, BW_BN))) return rc;
  const int items = p.x_tiles * p.dparts * p.nseg;
  const int grid = items < sms ? items : sms;
  static bool attr_done[3] = {false, false, false};
#define LAUNCH(M)                                                                                              \
  {                                                                                                            \
    if (!attr_done[M]) {                                                                                       \
      if (cudaFuncSetAttribute(bw_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM_BYTES) !=    \
          cudaSuccess)                                                                                         \
        return B2_ECUDA;                                                                                       \
      attr_done[M] = true;                                                                                     \
    }                                                                                                          \
    bw_kernel<M><<<grid, BW_THREADS, BW_SMEM_BYTES, stream>>>(tmX, tmY, p);                                    \
  }
  if (mode == BW_CLIP) LAUNCH(BW_CLIP)
  else if (mode == BW_GATED) LAUNCH(BW_GATED)
  else if (mode == BW_SIGLIP) LAUNCH(BW_SIGLIP)
  else return B2_EINVAL;
#undef LAUNCH
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
