// EXPERIMENTAL, opt-in (B200CLIP_BWD_QUAD=1) — correct but measured slower than logits_bwd2.cu (see the note at the
// dispatch in logits_bwd.cu and DESIGN.md 5.2).
// K3 on 4-CTA clusters: the CTA-pair backward of logits_bwd2.cu with the S / G tile SHARED between the two 256-column
// halves of D, so S is recomputed once per (X tile pair, Y tile) instead of once per D half (executed work per launch
// 4*B*N*D instead of 6*B*N*D).
//
//   cluster rank r = 2 q + p :  p = X tile inside the tile pair (the tcgen05 cta_group::2 peer, r ^ 1)
//                               q = D half  (output columns [256 q, 256 q + 256)); D partner = r ^ 2
//   pair q = CTAs {2q, 2q+1} issues M = 256 MMAs exactly like logits_bwd2.cu (leader = rank 2q).
//   Y tile t of the item belongs to pair (t & 1): that pair computes S_t (SS-MMA), its epilogue warps turn each CTA's
//   128 x 128 block into bf16 G_t, store it IN PLACE in TMEM (A operand of the pair's own TS output MMA) and PUSH the same
//   packed rows over DSMEM into the D partner's staging buffer (32 KB, SWIZZLE_128B K-major), which the other pair
//   consumes as the A operand of an SS output MMA. Every pair therefore runs S on every second step and the output
//   product on every step: 2048 instead of 3072 tensor cycles per step, 64 KB instead of 96 KB ingested per step.
//   DSMEM traffic: 32 KB per CTA every second step = 8 B/cycle.
//
//   shared memory per CTA: X panel 128 KB | G staging 32 KB | 4-slot TMA ring 64 KB | barriers | colscale staging
//   issue order of pair q (2 S tiles in flight so the epilogue latency is hidden):
//       S(q), S(q+2);  for t = 0..T-1:  out(t) [TS if t is ours, SS from the staging buffer otherwise];
//                                        if t is ours: S(t+4)
//   cross-proxy ordering of the pushed rows: st.shared::cluster (generic) -> fence.proxy.async -> mbarrier.arrive
//   .release.cluster on the consuming leader; the leader waits with .acquire.cluster, fences the proxy and issues.
#include "bwd_common.cuh"
#include "host_api.h"
#include <stdlib.h>

namespace b2 {

constexpr int BW4_SLOTS = 4;
constexpr int BW4_SLOT = 16384;
constexpr int BW4_XBYTES = BW_XRES_CHUNKS * BW_CHUNK;            // 128 KB
constexpr int BW4_STAGE_OFF = BW4_XBYTES;                         // 2 atoms [128 x 64] bf16 = 32 KB
constexpr int BW4_RING_OFF = BW4_STAGE_OFF + 2 * BW_CHUNK;
constexpr int BW4_BAR_OFF = BW4_RING_OFF + BW4_SLOTS * BW4_SLOT;
constexpr int BW4_COL_OFF = BW4_BAR_OFF + 256;
constexpr int BW4_SMEM = BW4_COL_OFF + 2 * 128 * 4 + 1024;

template <int kMode>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(BW_THREADS, 1)
bw4_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmYs,
           const __grid_constant__ CUtensorMap tmYo, BwParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* xs = smem;
  uint8_t* stage = smem + BW4_STAGE_OFF;
  uint8_t* ring = smem + BW4_RING_OFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BW4_BAR_OFF);
  uint64_t* full_bar = bars;                       // [4]  pair leader: TMA of both pair CTAs -> MMA
  uint64_t* empty_bar = bars + BW4_SLOTS;          // [4]  pair CTAs: MMA (multicast commit) -> TMA
  uint64_t* sfull_bar = bars + 2 * BW4_SLOTS;      // [2]  pair CTAs: own S tile ready
  uint64_t* gready_bar = sfull_bar + 2;            // [2]  pair leader: own G written (16 arrivals)
  uint64_t* accfull_bar = gready_bar + 2;          // [1]  pair CTAs
  uint64_t* accempty_bar = accfull_bar + 1;        // [1]  pair leader (16 arrivals)
  uint64_t* xfull_bar = accempty_bar + 1;          // [1]  pair leader
  uint64_t* xempty_bar = xfull_bar + 1;            // [1]  pair CTAs
  uint64_t* gsfull_bar = xempty_bar + 1;           // [1]  pair leader: partner pair's G landed in BOTH staging buffers (16)
  uint64_t* gsempty_bar = gsfull_bar + 1;          // [1]  every CTA: its D partner's staging buffer is free again
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gsempty_bar + 1);
  float* col_s = reinterpret_cast<float*>(smem + BW4_COL_OFF);   // [2][128]

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pp = rank & 1;                 // X tile inside the pair
  const int qq = rank >> 1;                // D half / pair index
  const bool leader = pp == 0;             // pair leader
  const uint16_t pair_mask = uint16_t(3u << (2 * qq));
  const uint16_t other_mask = uint16_t(3u << (2 * (1 - qq)));
  const int kchunks = p.Kp / BW_BK;
  const int kpairs = (kchunks + 1) / 2;
  const int x_pairs = (p.x_tiles + 1) / 2;
  const int items = x_pairs * p.nseg;
  const int cluster_id = blockIdx.x >> 2, n_clusters = gridDim.x >> 2;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmYs);
    tma_prefetch_desc(&tmYo);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < BW4_SLOTS; ++s) {
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
    mbar_init(gsfull_bar, 16);
    mbar_init(gsempty_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
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
  }

  auto decode = [&](int item, int& xp, int& j0, int& j1) {
    const int seg = item % p.nseg;
    xp = item / p.nseg;
    j0 = (int)((long long)p.y_tiles * seg / p.nseg);
    j1 = (int)((long long)p.y_tiles * (seg + 1) / p.nseg);
  };
  auto own = [&](int t) { return (t & 1) == qq; };
  // 64-column groups of this pair's output half (host guarantees Dp = 512: 4 groups, two per CTA)
  constexpr int NG2 = 2;

  if (warp == 0) {
    // ===================== TMA producer (every CTA for its own halves) =====================
    if (elect_one()) {
      int slot = 0;
      uint32_t phase = 0, xphase = 0;
      for (int item = cluster_id; item < items; item += n_clusters) {
        int xp, j0, j1;
        decode(item, xp, j0, j1);
        const int T = j1 - j0;
        if (T <= 0) continue;
        const int xt = 2 * xp + pp;
        mbar_wait(xempty_bar, xphase ^ 1);
        xphase ^= 1;
        if (leader) mbar_expect_tx(xfull_bar, 2 * kchunks * BW_CHUNK);
        for (int kc = 0; kc < kchunks; ++kc) tma_load_2d_pair(xs + kc * BW_CHUNK, &tmX, xfull_bar, kc * BW_BK, xt * BW_BM);
        auto load_s = [&](int t) {
          const int j = j0 + t;
          for (int kp = 0; kp < kpairs; ++kp) {
            const int nk = (2 * kp + 1 < kchunks) ? 2 : 1;
            mbar_wait(&empty_bar[slot], phase ^ 1);
            uint8_t* sl = ring + slot * BW4_SLOT;
            if (leader) mbar_expect_tx(&full_bar[slot], 2 * nk * 8192);
            for (int h = 0; h < nk; ++h)
              tma_load_2d_pair(sl + h * 8192, &tmYs, &full_bar[slot], (2 * kp + h) * BW_BK, j * BW_BN + 64 * pp);
            if (++slot == BW4_SLOTS) { slot = 0; phase ^= 1; }
          }
        };
        auto load_out = [&](int t) {
          const int j = j0 + t;
          for (int g2 = 0; g2 < NG2; ++g2) {
            mbar_wait(&empty_bar[slot], phase ^ 1);
            if (leader) mbar_expect_tx(&full_bar[slot], 2 * BW_CHUNK);
            tma_load_2d_pair(ring + slot * BW4_SLOT, &tmYo, &full_bar[slot],
                             p.hi_off + qq * BW_DP + (2 * g2 + pp) * BW_BK, j * BW_BN);
            if (++slot == BW4_SLOTS) { slot = 0; phase ^= 1; }
          }
        };
        if (qq < T) load_s(qq);
        if (qq + 2 < T) load_s(qq + 2);
        for (int t = 0; t < T; ++t) {
          load_out(t);
          if (own(t) && t + 4 < T) load_s(t + 4);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair leaders) =====================
    if (leader && elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(256, BW_BN, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(256, 128, 0, 1);        // A K-major (TMEM or staging), B MN-major
      int slot = 0;
      uint32_t phase = 0, xphase = 0;
      uint32_t own_issued = 0, own_used = 0;     // own S tiles issued / consumed (buffer = counter & 1)
      uint32_t recv_ctr = 0, acc_ctr = 0;
      const uint32_t xs_addr = smem_u32(xs), stage_addr = smem_u32(stage);
      for (int item = cluster_id; item < items; item += n_clusters) {
        int xp, j0, j1;
        decode(item, xp, j0, j1);
        const int T = j1 - j0;
        if (T <= 0) continue;
        mbar_wait(xfull_bar, xphase);
        xphase ^= 1;
        tc_fence_after();
        int own_left = (T - qq + 1) / 2;          // own tiles of this item still to issue
        auto mma_s = [&]() {
          const uint32_t tc = own_issued++;
          const uint32_t d_tmem = tmem_base + s_col0 + (tc & 1) * BW_BN;
          for (int kp = 0; kp < kpairs; ++kp) {
            const int nk = (2 * kp + 1 < kchunks) ? 2 : 1;
            mbar_wait(&full_bar[slot], phase);
            tc_fence_after();
            const uint32_t sl = smem_u32(ring + slot * BW4_SLOT);
            for (int h = 0; h < nk; ++h) {
              const uint64_t bdesc = make_smem_desc_sw128(sl + h * 8192, 1024);
              const uint64_t adesc = make_smem_desc_sw128(xs_addr + (2 * kp + h) * BW_CHUNK, 1024);
#pragma unroll
              for (int k = 0; k < BW_BK / 16; ++k)
                mma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc_s, (kp | h | k) != 0);
            }
            tc_commit_pair(&empty_bar[slot], pair_mask);
            if (++slot == BW4_SLOTS) { slot = 0; phase ^= 1; }
          }
          tc_commit_pair(&sfull_bar[tc & 1], pair_mask);
          if (--own_left == 0) tc_commit_pair(xempty_bar, pair_mask);      // last S product of the item issued
        };
        if (own_left == 0) tc_commit_pair(xempty_bar, pair_mask);          // (T == 1 and the tile is not ours)
        if (qq < T) mma_s();
        if (qq + 2 < T) mma_s();
        for (int t = 0; t < T; ++t) {
          if (t == 0) {
            mbar_wait(accempty_bar, (acc_ctr & 1) ^ 1);
            tc_fence_after();
          }
          const bool mine = own(t);
          uint32_t g_tmem = 0;
          if (mine) {
            const uint32_t tc = own_used++;
            mbar_wait(&gready_bar[tc & 1], (tc >> 1) & 1);
            tc_fence_after();
            g_tmem = tmem_base + s_col0 + (tc & 1) * BW_BN;
          } else {
            mbar_wait_cluster(gsfull_bar, recv_ctr & 1);
            ++recv_ctr;
            fence_proxy_async_all();
            tc_fence_after();
          }
          for (int g2 = 0; g2 < NG2; ++g2) {
            mbar_wait(&full_bar[slot], phase);
            tc_fence_after();
            const uint32_t sy = smem_u32(ring + slot * BW4_SLOT);
            const uint32_t d_tmem = tmem_base + acc_col + g2 * 128;
            const uint64_t bdesc0 = make_smem_desc_sw128(sy, 1024);
#pragma unroll
            for (int ks = 0; ks < BW_BN / 16; ++ks) {
              const uint64_t bdesc = bdesc0 + uint64_t(ks * (2048 >> 4));
              const uint32_t acc_on = !(t == 0 && ks == 0);
              if (mine) {
                mma_ts_pair(d_tmem, g_tmem + (ks >> 2) * 64 + (ks & 3) * 8, bdesc, idesc_o, acc_on);
              } else {
                const uint64_t adesc = make_smem_desc_sw128(stage_addr + (ks >> 2) * BW_CHUNK, 1024) + 2 * (ks & 3);
                mma_ss_pair(d_tmem, adesc, bdesc, idesc_o, acc_on);
              }
            }
            tc_commit_pair(&empty_bar[slot], pair_mask);
            if (++slot == BW4_SLOTS) { slot = 0; phase ^= 1; }
          }
          if (!mine) tc_commit_pair(gsempty_bar, other_mask);     // the senders may overwrite our staging buffers
          if (mine && t + 4 < T) mma_s();
          if (t == T - 1) {
            tc_commit_pair(accfull_bar, pair_mask);
            ++acc_ctr;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps 4..11 (every CTA, own TMEM, own tiles only) =====================
    const int q = warp & 3;
    const int wg = (warp - 4) >> 2;
    const int etid = threadIdx.x - 128;
    uint32_t own_ctr = 0, acc_ctr = 0, push_ctr = 0;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t cs_base = smem_u32(col_s) + wg * 64 * 4;
    const uint32_t lead_rank = 2 * qq, other_lead = 2 * (1 - qq);
    const uint32_t gready_remote0 = mapa_cluster(smem_u32(&gready_bar[0]), lead_rank);
    const uint32_t gready_remote1 = mapa_cluster(smem_u32(&gready_bar[1]), lead_rank);
    const uint32_t accempty_remote = mapa_cluster(smem_u32(accempty_bar), lead_rank);
    const uint32_t gsfull_remote = mapa_cluster(smem_u32(gsfull_bar), other_lead);
    // this thread's row inside the D partner's staging atom `wg`
    const int trow = q * 32 + lane;
    const uint32_t push_row = mapa_cluster(smem_u32(stage), rank ^ 2) + wg * BW_CHUNK + trow * 128;
    BwThread th;
    th.wg = wg;
    th.ydn = p.ydiag * p.gnorm;
    th.wn = p.wneg_c * p.gnorm;
    th.ign = 1.f / p.gnorm;
    th.nshift2 = -p.shift2;
    for (int item = cluster_id; item < items; item += n_clusters) {
      int xp, j0, j1;
      decode(item, xp, j0, j1);
      const int T = j1 - j0;
      if (T <= 0) continue;
      const int xt = 2 * xp + pp;
      th.row = xt * BW_BM + trow;
      th.row_ok = th.row < p.Nx;
      th.rs = 0.f;
      if (kMode != BW_SIGLIP) th.rs = th.row_ok ? p.rowscale[th.row] * p.gnorm : 0.f;
      double dtacc = 0.0, dlacc = 0.0, dbacc = 0.0;
      for (int t = qq; t < T; t += 2, ++own_ctr) {
        const int j = j0 + t;
        const bool want_scal = p.scal != nullptr;
        float tacc = 0.f, lacc = 0.f, bacc = 0.f;
        const int buf = own_ctr & 1;
        if (kMode != BW_SIGLIP) {
          if (etid < 128) {
            const int col = j * BW_BN + etid;
            col_s[buf * 128 + etid] = col < p.Ny ? p.colscale[col] * p.gnorm : 0.f;
          }
          named_bar_sync(1, 256);
        }
        mbar_wait(&sfull_bar[buf], (own_ctr >> 1) & 1);
        tc_fence_after();
        mbar_wait(gsempty_bar, (push_ctr & 1) ^ 1);        // partner consumed the G we pushed two steps ago
        ++push_ctr;
        const uint32_t sbase = tmem_base + lane_off + s_col0 + buf * BW_BN + wg * 64;
        // diag_corr / dp == 0 bookkeeping: the tile is processed exactly once (by its owner pair)
        bw_g_tile<kMode>(p, th, sbase, cs_base + buf * 128 * 4, col_s + buf * 128, xt, j, 0, want_scal, tacc, lacc, bacc,
                         push_row);
        fence_proxy_async_all();           // pushed rows (generic proxy) before the partner's MMA (async proxy)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(buf ? gready_remote1 : gready_remote0);
          mbar_arrive_cluster_release(gsfull_remote);
        }
        dtacc += (double)tacc;
        if (kMode == BW_SIGLIP) {
          dlacc += (double)lacc;
          dbacc += (double)bacc;
        }
      }
      // ---- drain this CTA's accumulator (own 128 rows x own 256 columns) ----
      mbar_wait(accfull_bar, acc_ctr & 1);
      ++acc_ctr;
      tc_fence_after();
      bw_drain(p, th, tmem_base + lane_off + acc_col + wg * 128, qq);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(accempty_remote);
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
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace b2

namespace b2host {
using namespace b2;

template <int kMode>
static int launch_bw4(const CUtensorMap& tmX, const CUtensorMap& tmYs, const CUtensorMap& tmYo, BwParams& p,
                      int x_pairs, int nseg_hint, cudaStream_t stream) {
  static int max_clusters = -1;
  if (max_clusters < 0) {
    if (cudaFuncSetAttribute(bw4_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW4_SMEM) != cudaSuccess)
      return B2_ECUDA;
    // how many 4-CTA clusters can be co-resident (GPC boundaries strand some SMs): launch exactly that many
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * (sm_count() / 4));
    cfg.blockDim = dim3(BW_THREADS);
    cfg.dynamicSmemBytes = BW4_SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, bw4_kernel<kMode>, &cfg) != cudaSuccess || n < 1) {
      cudaGetLastError();
      max_clusters = 0;
    } else {
      max_clusters = n;
    }
  }
  if (max_clusters < 8) return B2_ENOSYS;
  const int clusters_avail = max_clusters;
  int nseg = nseg_hint;
  if (nseg <= 0) {
    const int max_seg = p.y_tiles / 8 > 1 ? p.y_tiles / 8 : 1;
    double best = 1e30;
    nseg = 1;
    for (int s = 1; s <= max_seg && s <= 64; ++s) {
      const long long it = (long long)x_pairs * s;
      const long long waves = (it + clusters_avail - 1) / clusters_avail;
      const double cost = (double)waves * ((p.y_tiles + s - 1) / s + 4.0);
      if (cost < best * 0.995) { best = cost; nseg = s; }
    }
  }
  if (nseg > p.y_tiles) nseg = p.y_tiles;
  p.nseg = nseg;
  const int items = x_pairs * nseg;
  const int grid = 4 * (items < clusters_avail ? items : clusters_avail);
  bw4_kernel<kMode><<<grid, BW_THREADS, BW4_SMEM, stream>>>(tmX, tmYs, tmYo, p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

// Same contract as logits_bwd(); Kp <= 512 and Dp == 512 (two D halves), hp == 0. B2_ENOSYS if 4-CTA clusters
// cannot be scheduled usefully on this device (the caller falls back to the pair kernel).
int logits_bwd_quad(int mode, const void* X, const void* Y, int Nx, int Ny, int Kp, int Dp, int D, int hi_off, int ldx,
                    int ldy, float scale2, float shift2, float inv_tau, float bias, float wneg_c,
                    const float* rowscale, const float* colscale, float out_scale, float gnorm, int hp,
                    const float* dyn, float ydiag, int diag_off, float* diag_corr, float* dX, int ldd, double* scal,
                    int nseg_hint, cudaStream_t stream) {
  if (Kp > BW_XRES_CHUNKS * BW_BK || hp || Dp != 2 * BW_DP) return B2_EINVAL;
  BwParams p;
  p.Nx = Nx; p.Ny = Ny; p.Kp = Kp; p.Dp = Dp; p.D = D; p.hi_off = hi_off; p.ydiag = ydiag; p.diag_off = diag_off;
  p.diag_corr = diag_corr;
  p.x_tiles = (Nx + BW_BM - 1) / BW_BM;
  p.y_tiles = (Ny + BW_BN - 1) / BW_BN;
  p.dparts = 2;
  p.nseg = 1;
  p.scale2 = scale2; p.shift2 = shift2; p.inv_tau = inv_tau; p.bias = bias; p.wneg_c = wneg_c;
  p.rowscale = rowscale; p.colscale = colscale; p.out_scale = out_scale;
  p.gnorm = gnorm > 0.f ? gnorm : 1.f;
  p.hp = 0;
  p.lclamp = 30.f; p.yneg = 0.f; p.ent_coef = 0.f;
  p.dX = dX; p.ldd = ldd; p.scal = scal; p.dyn = dyn;
  CUtensorMap tmX, tmYs, tmYo;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmX, X, Nx, Kp, ldx, BW_BM))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmYs, Y, Ny, Kp, ldy, 64))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmYo, Y, Ny, Kp, ldy, BW_BN))) return rc;
  const int x_pairs = (p.x_tiles + 1) / 2;
  if (mode == BW_CLIP) return launch_bw4<BW_CLIP>(tmX, tmYs, tmYo, p, x_pairs, nseg_hint, stream);
  if (mode == BW_GATED) return launch_bw4<BW_GATED>(tmX, tmYs, tmYo, p, x_pairs, nseg_hint, stream);
  if (mode == BW_SIGLIP) return launch_bw4<BW_SIGLIP>(tmX, tmYs, tmYo, p, x_pairs, nseg_hint, stream);
  return B2_EINVAL;
}

}  // namespace b2host
