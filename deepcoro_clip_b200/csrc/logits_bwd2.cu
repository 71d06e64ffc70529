// K3 on CTA pairs: the logits backward of logits_bwd.cu with tcgen05 cta_group::2 (Kp <= 512, the headline shapes).
//
// Why pairs: one SM ingests at most 61.5 B/cycle from L2 through TMA (tools/ubench: the same figure for 1 or 148
// CTAs, so it is a per-SM port limit that multicast cannot lift). The single-CTA kernel needs the whole Y tile for the
// S product (128 KB) plus a 64 KB slice for the output product per step = 62.5 B/cycle against 3072 tensor cycles:
// ingest-bound (measured 54-60 % tensor pipe). With M = 256 MMAs spanning two SMs each CTA supplies only HALF of
// every B operand from its own shared memory, so a CTA ingests 96 KB per step (31 B/cycle) and the tensor pipe
// becomes the limiter.
//
//   cluster (2 CTAs) = two consecutive 128-row X tiles (CTA rank r owns X tile 2*pair + r, resident in its smem)
//   item             = (X tile pair, segment of Y tiles); steps (dp, j) as in logits_bwd.cu
//   S step           : tcgen05.mma.cta_group::2 SS, M = 256, N = 128: A = each CTA's X panel, B = Y_j rows
//                      [64 r, 64 r + 64) per CTA ([64 x 64] K-major boxes, two k-chunks per 16 KB ring slot)
//   G                : every CTA's epilogue warps turn ITS 128 x 128 S block (own TMEM) into bf16 G in place
//   out step         : tcgen05.mma.cta_group::2 TS, M = 256, N = 128 twice: A = G (own TMEM), B = Y_j[:, d-slice] as
//                      MN-major [128 x 64] boxes; CTA r supplies output columns dp*256 + 128 g + 64 r + [0, 64)
//   TMEM per CTA     : [0,256) accumulator (own 128 rows x 256 d) | [256,384) S/G 0 | [384,512) S/G 1
//   barriers         : TMA of BOTH CTAs credits the leader's full barriers; the leader's MMA lane commits with a
//                      multicast arrive to both CTAs' empty / S-ready / accumulator-ready barriers; the peer's
//                      epilogue warps arrive remotely on the leader's G-ready / accumulator-drained barriers.
#include "bwd_common.cuh"
#include "host_api.h"

namespace b2 {

constexpr int BW2_SLOTS = 6;
constexpr int BW2_SLOT = 16384;
constexpr int BW2_XBYTES = BW_XRES_CHUNKS * BW_CHUNK;             // 128 KB resident X panel
constexpr int BW2_BAR_OFF = BW2_XBYTES + BW2_SLOTS * BW2_SLOT;
constexpr int BW2_COL_OFF = BW2_BAR_OFF + 256;
constexpr int BW2_SMEM = BW2_COL_OFF + 2 * 128 * 4 + 1024;

template <int kMode>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BW_THREADS, 1)
bw2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmYs,
           const __grid_constant__ CUtensorMap tmYo, BwParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = smem;
  uint8_t* ring = smem + BW2_XBYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BW2_BAR_OFF);
  uint64_t* full_bar = bars;                       // [6]  leader only: TMA (both CTAs) -> MMA
  uint64_t* empty_bar = bars + BW2_SLOTS;          // [6]  both: MMA (multicast commit) -> TMA
  uint64_t* sfull_bar = bars + 2 * BW2_SLOTS;      // [2]  both: S tile ready (multicast commit)
  uint64_t* gready_bar = sfull_bar + 2;            // [2]  leader only: G written, 16 arrivals (8 warps x 2 CTAs)
  uint64_t* accfull_bar = gready_bar + 2;          // [1]  both: accumulator ready (multicast commit)
  uint64_t* accempty_bar = accfull_bar + 1;        // [1]  leader only: accumulator drained, 16 arrivals
  uint64_t* xfull_bar = accempty_bar + 1;          // [1]  leader only: both X panels landed
  uint64_t* xempty_bar = xfull_bar + 1;            // [1]  both: X panels free (multicast commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xempty_bar + 1);
  float* col_s = reinterpret_cast<float*>(smem + BW2_COL_OFF);   // [2][128]

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int kchunks = p.Kp / BW_BK;
  const int kpairs = (kchunks + 1) / 2;
  const int x_pairs = (p.x_tiles + 1) / 2;
  const int items = x_pairs * p.nseg;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmYs);
    tma_prefetch_desc(&tmYo);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < BW2_SLOTS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sfull_bar[s], 1);
      mbar_init(&gready_bar[s], 16);
    }
    mbar_init(accfull_bar, 1);
    mbar_init(accempty_bar, 16);
    mbar_init(xfull_bar, 1);
    mbar_init(xempty_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_col = 0, s_col0 = 256;
  if (p.dyn) {
    p.scale2 = p.dyn[0];
    p.shift2 = p.dyn[1];
    p.inv_tau = p.dyn[2];
    p.bias = p.dyn[5];
    p.out_scale = p.dyn[2];
    p.lclamp = p.dyn[8];
    p.yneg = p.dyn[9];
    p.stable = kMode != BW_SIGLIP && p.dyn[11] != 0.f;
  }

  auto decode = [&](int item, int& xp, int& j0, int& j1) {
    const int seg = item % p.nseg;
    xp = item / p.nseg;
    j0 = (int)((long long)p.y_tiles * seg / p.nseg);
    j1 = (int)((long long)p.y_tiles * (seg + 1) / p.nseg);
  };
  // 64-column groups of the dp-th output slice (<= 4); group g is supplied by CTA (g & 1) to out-MMA (g >> 1)
  auto n_dgroups = [&](int dp) {
    const int rem = (p.Dp - dp * BW_DP) / BW_BK;
    return rem < 4 ? rem : 4;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs, each for its own halves) =====================
    if (elect_one()) {
      int slot = 0;
      uint32_t phase = 0, xphase = 0;
      for (int item = cluster_id; item < items; item += n_clusters) {
        int xp, j0, j1;
        decode(item, xp, j0, j1);
        const int nj = j1 - j0;
        if (nj <= 0) continue;
        const int T = nj * p.dparts;
        const int xt = 2 * xp + (int)rank;
        mbar_wait(xempty_bar, xphase ^ 1);
        xphase ^= 1;
        if (leader) mbar_expect_tx(xfull_bar, 2 * kchunks * BW_CHUNK);
        for (int kc = 0; kc < kchunks; ++kc) tma_load_2d_pair(xs + kc * BW_CHUNK, &tmX, xfull_bar, kc * BW_BK, xt * BW_BM);
        auto load_s = [&](int t) {
          const int j = j0 + t % nj;
          for (int kp = 0; kp < kpairs; ++kp) {
            const int nk = (2 * kp + 1 < kchunks) ? 2 : 1;
            mbar_wait(&empty_bar[slot], phase ^ 1);
            uint8_t* sl = ring + slot * BW2_SLOT;
            if (leader) mbar_expect_tx(&full_bar[slot], 2 * nk * 8192);
            for (int h = 0; h < nk; ++h)
              tma_load_2d_pair(sl + h * 8192, &tmYs, &full_bar[slot], (2 * kp + h) * BW_BK, j * BW_BN + 64 * (int)rank);
            if (++slot == BW2_SLOTS) { slot = 0; phase ^= 1; }
          }
        };
        auto load_out = [&](int t) {
          const int dp = t / nj, j = j0 + t % nj;
          const int ng = n_dgroups(dp);          // even (host checks Dp % 128 == 0)
          for (int g2 = 0; g2 < ng / 2; ++g2) {
            mbar_wait(&empty_bar[slot], phase ^ 1);
            if (leader) mbar_expect_tx(&full_bar[slot], 2 * BW_CHUNK);
            tma_load_2d_pair(ring + slot * BW2_SLOT, &tmYo, &full_bar[slot],
                             p.hi_off + dp * BW_DP + (2 * g2 + (int)rank) * BW_BK, j * BW_BN);
            if (++slot == BW2_SLOTS) { slot = 0; phase ^= 1; }
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
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(256, BW_BN, 0, 0);      // A, B K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(256, 128, 0, 1);        // A = G (TMEM), B MN-major, 64 columns / CTA
      int slot = 0;
      uint32_t phase = 0, xphase = 0;
      uint32_t tile_ctr = 0, acc_ctr = 0;
      const uint32_t xs_addr = smem_u32(xs);
      for (int item = cluster_id; item < items; item += n_clusters) {
        int xp, j0, j1;
        decode(item, xp, j0, j1);
        const int nj = j1 - j0;
        if (nj <= 0) continue;
        const int T = nj * p.dparts;
        mbar_wait(xfull_bar, xphase);
        xphase ^= 1;
        tc_fence_after();
        auto mma_s = [&](uint32_t tc) {
          const uint32_t d_tmem = tmem_base + s_col0 + (tc & 1) * BW_BN;
          for (int kp = 0; kp < kpairs; ++kp) {
            const int nk = (2 * kp + 1 < kchunks) ? 2 : 1;
            mbar_wait(&full_bar[slot], phase);
            tc_fence_after();
            const uint32_t sl = smem_u32(ring + slot * BW2_SLOT);
            for (int h = 0; h < nk; ++h) {
              const uint64_t bdesc = make_smem_desc_sw128(sl + h * 8192, 1024);
              const uint64_t adesc = make_smem_desc_sw128(xs_addr + (2 * kp + h) * BW_CHUNK, 1024);
#pragma unroll
              for (int k = 0; k < BW_BK / 16; ++k)
                mma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc_s, (kp | h | k) != 0);
            }
            tc_commit_pair(&empty_bar[slot], 3);
            if (++slot == BW2_SLOTS) { slot = 0; phase ^= 1; }
          }
          tc_commit_pair(&sfull_bar[tc & 1], 3);
        };
        auto mma_out = [&](uint32_t tc, int dp, bool first) {
          const uint32_t g_tmem = tmem_base + s_col0 + (tc & 1) * BW_BN;
          const int ng = n_dgroups(dp);
          for (int g2 = 0; g2 < ng / 2; ++g2) {
            mbar_wait(&full_bar[slot], phase);
            tc_fence_after();
            const uint32_t sy = smem_u32(ring + slot * BW2_SLOT);
            const uint32_t d_tmem = tmem_base + acc_col + g2 * 128;
            const uint64_t bdesc0 = make_smem_desc_sw128(sy, 1024);
            constexpr uint32_t idesc = idesc_o;
#pragma unroll
            for (int ks = 0; ks < BW_BN / 16; ++ks) {
              const uint64_t bdesc = bdesc0 + uint64_t(ks * (2048 >> 4));
              const uint32_t a_tmem = g_tmem + (ks >> 2) * 64 + (ks & 3) * 8;
              mma_ts_pair(d_tmem, a_tmem, bdesc, idesc, !(first && ks == 0));
            }
            tc_commit_pair(&empty_bar[slot], 3);
            if (++slot == BW2_SLOTS) { slot = 0; phase ^= 1; }
          }
        };
        mma_s(tile_ctr);
        if (T == 1) tc_commit_pair(xempty_bar, 3);
        for (int t = 0; t < T; ++t) {
          const int dp = t / nj, jr = t - dp * nj;
          if (t + 1 < T) {
            mma_s(tile_ctr + 1);
            if (t + 2 == T) tc_commit_pair(xempty_bar, 3);
          }
          if (jr == 0) {
            mbar_wait(accempty_bar, (acc_ctr & 1) ^ 1);
            tc_fence_after();
          }
          mbar_wait(&gready_bar[tile_ctr & 1], (tile_ctr >> 1) & 1);
          tc_fence_after();
          mma_out(tile_ctr, dp, jr == 0);
          ++tile_ctr;
          if (jr == nj - 1) {
            tc_commit_pair(accfull_bar, 3);
            ++acc_ctr;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps 4..11 (both CTAs, own TMEM) =====================
    const int q = warp & 3;
    const int wg = (warp - 4) >> 2;
    const int etid = threadIdx.x - 128;
    uint32_t tile_ctr = 0, acc_ctr = 0;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t cs_base = smem_u32(col_s) + wg * 64 * 4;
    // the leader's G-ready / accumulator-drained barriers, as shared::cluster addresses
    const uint32_t gready_remote0 = mapa_cluster(smem_u32(&gready_bar[0]), 0);
    const uint32_t gready_remote1 = mapa_cluster(smem_u32(&gready_bar[1]), 0);
    const uint32_t accempty_remote = mapa_cluster(smem_u32(accempty_bar), 0);
    BwThread th;
    th.wg = wg;
    th.ydn = p.ydiag * p.gnorm;
    th.wn = p.wneg_c * p.gnorm;
    th.ign = 1.f / p.gnorm;
    th.nshift2 = -p.shift2;
    for (int item = cluster_id; item < items; item += n_clusters) {
      int xp, j0, j1;
      decode(item, xp, j0, j1);
      const int nj = j1 - j0;
      if (nj <= 0) continue;
      const int T = nj * p.dparts;
      const int xt = 2 * xp + (int)rank;
      th.row = xt * BW_BM + q * 32 + lane;
      th.row_ok = th.row < p.Nx;
      th.rs = 0.f;
      if (kMode != BW_SIGLIP) th.rs = bw_stat(p, p.rowscale, th.row, th.row_ok);
      double dtacc = 0.0, dlacc = 0.0, dbacc = 0.0;
      for (int t = 0; t < T; ++t, ++tile_ctr) {
        const int dp = t / nj, jr = t - dp * nj, j = j0 + jr;
        const bool want_scal = p.scal != nullptr && dp == 0;
        float tacc = 0.f, lacc = 0.f, bacc = 0.f;
        const int buf = tile_ctr & 1;
        if (kMode != BW_SIGLIP) {
          if (etid < 128) {
            const int col = j * BW_BN + etid;
            col_s[buf * 128 + etid] = bw_stat(p, p.colscale, col, col < p.Ny);
          }
          named_bar_sync(1, 256);
        }
        mbar_wait(&sfull_bar[buf], (tile_ctr >> 1) & 1);
        tc_fence_after();
        const uint32_t sbase = tmem_base + lane_off + s_col0 + buf * BW_BN + wg * 64;
        bw_g_tile<kMode>(p, th, sbase, cs_base + buf * 128 * 4, col_s + buf * 128, xt, j, dp, want_scal, tacc, lacc,
                         bacc);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(buf ? gready_remote1 : gready_remote0);
        dtacc += (double)tacc;
        if (kMode == BW_SIGLIP) {
          dlacc += (double)lacc;
          dbacc += (double)bacc;
        }
        if (jr == nj - 1) {
          mbar_wait(accfull_bar, acc_ctr & 1);
          ++acc_ctr;
          tc_fence_after();
          bw_drain(p, th, tmem_base + lane_off + acc_col + wg * 128, dp);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(accempty_remote);
        }
      }
      if (p.scal) {
        for (int o = 16; o > 0; o >>= 1) {
          dtacc += __shfl_xor_sync(0xffffffffu, dtacc, o);
          if (kMode == BW_SIGLIP) {
            dlacc += __shfl_xor_sync(0xffffffffu, dlacc, o);
            dbacc += __shfl_xor_sync(0xffffffffu, dbacc, o);
          }
        }
        if (lane == 0) {
          atomicAdd(p.scal + 0, dtacc * (double)th.ign);
          if (kMode == BW_SIGLIP) {
            atomicAdd(p.scal + 1, dlacc);
            atomicAdd(p.scal + 2, dbacc * (double)th.ign);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // no CTA may exit (or free TMEM) while its peer can still signal it / read its smem
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace b2

namespace b2host {
using namespace b2;

template <int kMode>
static int launch_bw2(const CUtensorMap& tmX, const CUtensorMap& tmYs, const CUtensorMap& tmYo, const BwParams& p,
                      int grid, cudaStream_t stream) {
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device() & 63];   // cudaFuncSetAttribute is per device
  if (!attr_done) {
    if (cudaFuncSetAttribute(bw2_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW2_SMEM) != cudaSuccess)
      return B2_ECUDA;
    attr_done = true;
  }
  bw2_kernel<kMode><<<grid, BW_THREADS, BW2_SMEM, stream>>>(tmX, tmYs, tmYo, p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

// Same contract as logits_bwd() (host_api.h); requires Kp <= 512. Returns B2_EINVAL for shapes it does not take.
int logits_bwd_pair(int mode, const void* X, const void* Y, int Nx, int Ny, int Kp, int Dp, int D, int hi_off, int ldx,
                    int ldy, float scale2, float shift2, float inv_tau, float bias, float wneg_c,
                    const float* rowscale, const float* colscale, float out_scale, float gnorm, int hp,
                    const float* dyn, float ydiag, int diag_off, float* diag_corr, float* dX, int ldd, double* scal,
                    int nseg_hint, cudaStream_t stream) {
  if (Kp > BW_XRES_CHUNKS * BW_BK || hp || Dp % 128) return B2_EINVAL;
  BwParams p;
  p.Nx = Nx; p.Ny = Ny; p.Kp = Kp; p.Dp = Dp; p.D = D; p.hi_off = hi_off; p.ydiag = ydiag; p.diag_off = diag_off;
  p.diag_corr = diag_corr;
  p.x_tiles = (Nx + BW_BM - 1) / BW_BM;
  p.y_tiles = (Ny + BW_BN - 1) / BW_BN;
  p.dparts = (Dp + BW_DP - 1) / BW_DP;
  const int clusters = sm_count() / 2;
  const int x_pairs = (p.x_tiles + 1) / 2;
  int nseg = nseg_hint;
  if (nseg <= 0) {
    const int max_seg = p.y_tiles / 4 > 1 ? p.y_tiles / 4 : 1;
    double best = 1e30;
    nseg = 1;
    for (int s = 1; s <= max_seg && s <= 64; ++s) {
      const long long it = (long long)x_pairs * s;
      const long long waves = (it + clusters - 1) / clusters;
      const double cost = (double)waves * (p.dparts * ((p.y_tiles + s - 1) / s) + 3.0);
      if (cost < best * 0.995) { best = cost; nseg = s; }
    }
  }
  if (nseg > p.y_tiles) nseg = p.y_tiles;
  p.nseg = nseg;
  p.scale2 = scale2; p.shift2 = shift2; p.inv_tau = inv_tau; p.bias = bias; p.wneg_c = wneg_c;
  p.rowscale = rowscale; p.colscale = colscale; p.out_scale = out_scale;
  p.gnorm = gnorm > 0.f ? gnorm : 1.f;
  p.hp = 0;
  p.lclamp = 30.f; p.yneg = 0.f; p.ent_coef = 0.f; p.stable = 0;
  p.dX = dX; p.ldd = ldd; p.scal = scal; p.dyn = dyn;
  CUtensorMap tmX, tmYs, tmYo;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmX, X, Nx, Kp, ldx, BW_BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmYs, Y, Ny, Kp, ldy, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmYo, Y, Ny, Kp, ldy, BW_BN))) return rc;
  const int items = x_pairs * p.nseg;
  const int grid = 2 * (items < clusters ? items : clusters);
  if (mode == BW_CLIP) return launch_bw2<BW_CLIP>(tmX, tmYs, tmYo, p, grid, stream);
  if (mode == BW_GATED) return launch_bw2<BW_GATED>(tmX, tmYs, tmYo, p, grid, stream);
  if (mode == BW_SIGLIP) return launch_bw2<BW_SIGLIP>(tmX, tmYs, tmYo, p, grid, stream);
  return B2_EINVAL;
}

}  // namespace b2host
