// K3 on CTA pairs with the WHOLE output width in TMEM: no S recompute per 256-column slice of D (Dp = 256/512/768).
//
// Why: the 128-row kernels (logits_bwd.cu, logits_bwd2.cu) keep a [128 x 256] fp32 accumulator next to the S/G buffers,
// so D = 512 recomputes every S tile twice and D = 768 three times, and at Kp = 768 the 128-row X panel (192 KB) no
// longer fits in shared memory, so X is re-streamed for every Y tile (measured 5.5 ms per launch at 32k x 32k x 768,
// 0.18 of the bf16 peak). Here a CTA owns only 64 X rows:
//
//   cluster (2 CTAs) = one 128-row X tile; CTA rank r keeps rows [64 r, 64 r + 64) resident (Kp/64 boxes of 8 KB)
//   tcgen05.mma.cta_group::2 with M = 128: every CTA holds a [64 x N] slice of the accumulator in the "2x2" TMEM
//     layout — row m in lane m for columns [0, N/2) and in lane 64 + m for columns [N/2, N) — i.e. N/2 TMEM columns
//     per N accumulator columns, so [64 x 768] fp32 takes 384 of the 512 columns and two [64 x 128] S buffers 128.
//   S step   : SS, M = 128, N = 128: A = the X panels, B = Y_j rows [64 r, 64 r + 64) per CTA, K = Kp
//   G        : epilogue warps read S (lanes 32q.., 32 columns per thread), form the gradient, round to bf16 and store it
//              into SHARED memory as the K-major SWIZZLE_128B A operand [64 rows x 128 j] (the 2x2 accumulator layout
//              splits the columns over the two lane halves, so G cannot be consumed from TMEM as in logits_bwd2.cu)
//   out step : SS, M = 128, N = 256, K = 128 (the j of the tile), once per 256 output columns: A = G (own shared
//              memory), B = Y_j[:, 256 n + 128 r + [0, 128)] as two MN-major [128 j x 64] boxes (LBO = 16 KB apart)
//   The accumulator is drained once per (X tile, Y segment): executed work 4 * B * N * D per launch instead of
//   (2 + 2 * Dp/256) * B * N * D.
//
// Barriers as in logits_bwd2.cu (TMA of both CTAs credits the leader's full barriers, multicast commits, remote arrives
// of the peer's epilogue warps); the G hand-off crosses proxies (generic st.shared -> tensor-core read of the SAME
// CTA's shared memory), so every writer executes fence.proxy.async.shared::cta before its warp arrives.
#include "bwd_common.cuh"
#include "host_api.h"
#include <stdlib.h>

namespace b2 {

constexpr int BW3_XROWS = 64;               // X rows per CTA
constexpr int BW3_XCHUNK = BW3_XROWS * BW_BK * 2;     // 8 KB: [64 rows x 64 bf16]
constexpr int BW3_SLOTS = 6;
constexpr int BW3_SLOT = 16384;

// kNJ = Y rows per step. 128: Kp <= 768 (X panel 96 KB, two [64 x 128] S buffers in 128 TMEM columns next to a 384-column
// accumulator). 256: Kp <= 512 only (X panel 64 KB, two [64 x 256] S buffers in 256 columns next to a 256-column
// accumulator): the S product re-reads its A operand (the X panel) half as often per logit, which matters because this
// kernel is bound by the 128 B/cycle/SM shared-memory port (tensor-core operand reads + TMA writes + G stores).
template <int kNJ>
struct Bw3Cfg {
  static constexpr int kMaxK = kNJ == 128 ? 12 : 8;                   // resident X chunks
  static constexpr int kSRows = kNJ / 2;                              // Y rows per CTA in the S product
  static constexpr int kSBox = kSRows * 128;                          // bytes of one [kSRows x 64] S box
  static constexpr int kCps = BW3_SLOT / kSBox;                       // k-chunks per ring slot (2 | 1)
  static constexpr int kGBuf = (kNJ / 64) * BW3_XCHUNK;               // [64 rows x kNJ] bf16
  static constexpr int kGOff = kMaxK * BW3_XCHUNK;
  static constexpr int kRingOff = kGOff + 2 * kGBuf;
  static constexpr int kBarOff = kRingOff + BW3_SLOTS * BW3_SLOT;
  static constexpr int kColOff = kBarOff + 256;
  static constexpr int kSmem = kColOff + 8 * 32 * 4 + 1024;
  static constexpr int kSCol = kNJ == 128 ? 384 : 256;                // TMEM: accumulator | S 0 | S 1
  static_assert(kSmem <= 232448, "shared memory budget");
};

// TMA store with an evict-first L2 policy: the 2 N^2 bytes of G stream through L2 once and must not displace the Y operand
// that every cluster keeps re-reading from it
__device__ __forceinline__ void bw3_tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int x, int y) {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(x), "r"(y), "l"(pol) : "memory");
}
__device__ __forceinline__ void bw3_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bw3_bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bw3_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 32 S values of one thread (row th.row, tile columns c0 .. c0+31) -> 32 gradient values, packed bf16x2.
// cs_addr: shared address of the 32 staged colscale*gnorm values; colg0: global column of element 0.
template <int kMode, bool kFast, bool kStable>
__device__ __forceinline__ void bw3_g32(const BwParams& p, const BwThread& th, const uint32_t (&acc)[32], uint32_t cs_addr,
                                        int colg0, int dcol, float& tacc, float& lacc, float& bacc,
                                        uint32_t (&packed)[16]) {
  constexpr bool kSig = BwIsSiglip<kMode>::value;
  const float lc = p.lclamp, yneg = p.yneg;
#pragma unroll
  for (int e = 0; e < 32; e += 4) {
    float cs4[4] = {0.f, 0.f, 0.f, 0.f};
    if (!kSig) lds128(cs_addr + e * 4, cs4);
    float g4[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float s = __uint_as_float(acc[e + h]);
      const bool ok = kFast || (th.row_ok && (colg0 + e + h) < p.Ny);
      float g;
      if (kSig) {
        const float R = fmaf(s, p.inv_tau, p.bias);
        const float Lc = fminf(fmaxf(R, -lc), lc);
        const float ex = ex2_approx(-1.4426950408889634f * fabsf(Lc));
        const float den = 1.f + ex;
        const float r = __fdividef(1.f, den);
        const float sig = Lc >= 0.f ? r : ex * r;
        g = (fabsf(R) <= lc) ? th.wn * (sig - yneg) : 0.f;
        float sp = fmaf(-yneg, Lc, fmaxf(Lc, 0.f) + log1p_ex(ex));
        if (!ok) { g = 0.f; sp = 0.f; }
        lacc += sp;
        bacc += g;
        tacc = fmaf(g, s, tacc);
      } else {
        float f = s, fp = 1.f;
        if (kMode == BW_GATED) {
          const float ex = ex2_approx(-1.4426950408889634f * s);
          const float sig = __fdividef(1.f, 1.f + ex);
          f = s * sig;
          fp = sig * (1.f + s * (1.f - sig));
        }
        g = bw_softmax_g<kStable>(f, p.scale2, th.nshift2, th.rs, cs4[h]);
        if (!kFast) {
          if (e + h == dcol) g -= th.ydn;
          if (!ok) g = 0.f;
        }
        tacc = fmaf(g, f, tacc);
        if (kMode == BW_GATED) g *= fp;
        if (!kFast && e + h == dcol && th.row_ok && p.diag_corr) {
          const float gb = __bfloat162float(__float2bfloat16_rn(g));
          p.diag_corr[2 * th.row] = (g - gb) * th.ign;
          p.diag_corr[2 * th.row + 1] = gb * th.ign;
        }
      }
      g4[h] = g;
    }
    packed[e >> 1] = pack_bf16x2(g4[0], g4[1]);
    packed[(e >> 1) + 1] = pack_bf16x2(g4[2], g4[3]);
  }
}

// Work of one cluster, identical in every warp role. nseg > 0: items = (X tile, segment of the Y tiles) dealt round-robin
// (explicit nseg_hint). nseg == 0 (default): the flattened (X tile, Y tile) sequence is cut into one CONTIGUOUS, equally
// long range per cluster — perfect balance for any shape (at 8 ranks a rank has only 32 X tiles for 74 clusters); a range
// that crosses X tiles is walked as up to a few (X tile, [j0, j1)) pieces, each with its own X panel load and drain.
struct Bw3Sched {
  long long r, r1;
  int item, items, stride, nseg, y_tiles;
  __device__ __forceinline__ void init(const BwParams& p, int cluster_id, int n_clusters) {
    nseg = p.nseg;
    y_tiles = p.y_tiles;
    if (nseg > 0) {
      items = p.x_tiles * nseg;
      item = cluster_id;
      stride = n_clusters;
    } else {
      const long long total = (long long)p.x_tiles * p.y_tiles;
      r = total * cluster_id / n_clusters;
      r1 = total * (cluster_id + 1) / n_clusters;
    }
  }
  __device__ __forceinline__ bool next(int& xt, int& j0, int& j1) {
    if (nseg > 0) {
      while (item < items) {
        const int seg = item % nseg;
        xt = item / nseg;
        j0 = (int)((long long)y_tiles * seg / nseg);
        j1 = (int)((long long)y_tiles * (seg + 1) / nseg);
        item += stride;
        if (j1 > j0) return true;
      }
      return false;
    }
    if (r >= r1) return false;
    xt = (int)(r / y_tiles);
    j0 = (int)(r - (long long)xt * y_tiles);
    const long long left = r1 - r;
    j1 = (left < (long long)(y_tiles - j0)) ? j0 + (int)left : y_tiles;
    r += j1 - j0;
    return true;
  }
};

template <int kMode, int kNJ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BW_THREADS, 1)
bw3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmYs,
           const __grid_constant__ CUtensorMap tmYo, const __grid_constant__ CUtensorMap tmG, BwParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using C = Bw3Cfg<kNJ>;
  uint8_t* xs = smem;
  uint8_t* gbuf = smem + C::kGOff;
  uint8_t* ring = smem + C::kRingOff;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* full_bar = bars;                       // [6]  leader only: TMA (both CTAs) -> MMA
  uint64_t* empty_bar = bars + BW3_SLOTS;          // [6]  both: MMA (multicast commit) -> TMA
  uint64_t* sfull_bar = bars + 2 * BW3_SLOTS;      // [2]  both: S tile ready (multicast commit)
  uint64_t* gready_bar = sfull_bar + 2;            // [2]  leader only: G in shared memory, 16 arrivals (8 warps x 2 CTAs)
  uint64_t* accfull_bar = gready_bar + 2;          // [1]  both: accumulator ready (multicast commit)
  uint64_t* accempty_bar = accfull_bar + 1;        // [1]  leader only: accumulator drained, 16 arrivals
  uint64_t* xfull_bar = accempty_bar + 1;          // [1]  leader only: both X panels landed
  uint64_t* xempty_bar = xfull_bar + 1;            // [1]  both: X panels free (multicast commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xempty_bar + 1);
  float* col_s = reinterpret_cast<float*>(smem + C::kColOff);   // [8 epilogue warps][32]: warp-private colscale stage

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int kchunks = p.Kp / BW_BK;                // even (host checks Kp % 256 == 0)
  const int nparts = p.Dp / 256;                   // N = 256 accumulator parts
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmYs);
    tma_prefetch_desc(&tmYo);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < BW3_SLOTS; ++s) {
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
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.dyn) {
    p.scale2 = p.dyn[0];
    p.shift2 = p.dyn[1];
    p.inv_tau = p.dyn[2];
    p.bias = p.dyn[5];
    p.out_scale = p.dyn[2];
    p.lclamp = p.dyn[8];
    p.yneg = p.dyn[9];
    p.stable = !BwIsSiglip<kMode>::value && p.dyn[11] != 0.f;
  }

  if (warp == 0) {
    // ===================== TMA producer (both CTAs, each for its own halves) =====================
    if (elect_one()) {
      int slot = 0;
      uint32_t phase = 0, xphase = 0;
      Bw3Sched sched;
      sched.init(p, cluster_id, n_clusters);
      int xt, j0, j1;
      while (sched.next(xt, j0, j1)) {
        const int nj = j1 - j0;
        mbar_wait(xempty_bar, xphase ^ 1);
        xphase ^= 1;
        if (leader) mbar_expect_tx(xfull_bar, 2 * kchunks * BW3_XCHUNK);
        for (int kc = 0; kc < kchunks; ++kc)
          tma_load_2d_pair(xs + kc * BW3_XCHUNK, &tmX, xfull_bar, kc * BW_BK, xt * BW_BM + BW3_XROWS * (int)rank);
        auto load_s = [&](int t) {
          const int j = j0 + t;
          for (int kp = 0; kp < kchunks / C::kCps; ++kp) {
            mbar_wait(&empty_bar[slot], phase ^ 1);
            uint8_t* sl = ring + slot * BW3_SLOT;
            if (leader) mbar_expect_tx(&full_bar[slot], 2 * BW3_SLOT);
            for (int h = 0; h < C::kCps; ++h)
              tma_load_2d_pair(sl + h * C::kSBox, &tmYs, &full_bar[slot], (C::kCps * kp + h) * BW_BK,
                               j * kNJ + C::kSRows * (int)rank);
            if (++slot == BW3_SLOTS) { slot = 0; phase ^= 1; }
          }
        };
        auto load_out = [&](int t) {
          const int j = j0 + t;
          for (int n = 0; n < nparts; ++n)
            for (int jh = 0; jh < kNJ / 128; ++jh)
              for (int h = 0; h < 2; ++h) {
                mbar_wait(&empty_bar[slot], phase ^ 1);
                if (leader) mbar_expect_tx(&full_bar[slot], 2 * BW3_SLOT);
                tma_load_2d_pair(ring + slot * BW3_SLOT, &tmYo, &full_bar[slot],
                                 p.hi_off + (4 * n + 2 * (int)rank + h) * BW_BK, j * kNJ + jh * 128);
                if (++slot == BW3_SLOTS) { slot = 0; phase ^= 1; }
              }
        };
        load_s(0);
        for (int t = 0; t < nj; ++t) {
          if (t + 1 < nj) load_s(t + 1);
          load_out(t);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one()) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, kNJ, 0, 0);        // A, B K-major; 64 rows of A, kNJ/2 of B per CTA
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 256, 0, 1);        // A = G (K-major), B MN-major, 128 columns / CTA
      int slot = 0;
      uint32_t phase = 0, xphase = 0;
      uint32_t tile_ctr = 0, acc_ctr = 0;
      const uint32_t xs_addr = smem_u32(xs), g_addr = smem_u32(gbuf);
      Bw3Sched sched;
      sched.init(p, cluster_id, n_clusters);
      int xt, j0, j1;
      while (sched.next(xt, j0, j1)) {
        const int nj = j1 - j0;
        mbar_wait(xfull_bar, xphase);
        xphase ^= 1;
        tc_fence_after();
        auto mma_s = [&](uint32_t tc) {
          const uint32_t d_tmem = tmem_base + C::kSCol + (tc & 1) * (kNJ / 2);
          for (int kp = 0; kp < kchunks / C::kCps; ++kp) {
            mbar_wait(&full_bar[slot], phase);
            tc_fence_after();
            const uint32_t sl = smem_u32(ring + slot * BW3_SLOT);
#pragma unroll
            for (int h = 0; h < C::kCps; ++h) {
              const uint64_t bdesc = make_smem_desc_sw128(sl + h * C::kSBox, 1024);
              const uint64_t adesc = make_smem_desc_sw128(xs_addr + (C::kCps * kp + h) * BW3_XCHUNK, 1024);
#pragma unroll
              for (int k = 0; k < BW_BK / 16; ++k)
                mma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc_s, (kp | h | k) != 0);
            }
            tc_commit_pair(&empty_bar[slot], 3);
            if (++slot == BW3_SLOTS) { slot = 0; phase ^= 1; }
          }
          tc_commit_pair(&sfull_bar[tc & 1], 3);
        };
        auto mma_out = [&](uint32_t tc, bool first) {
          const uint32_t ga = g_addr + (tc & 1) * C::kGBuf;
          for (int n = 0; n < nparts; ++n)
            for (int jh = 0; jh < kNJ / 128; ++jh) {
              // two consecutive ring slots = the two [128 j x 64] boxes of this CTA's 128 output columns (never wraps:
              // every step uses an even number of slots and BW3_SLOTS is even)
              mbar_wait(&full_bar[slot], phase);
              mbar_wait(&full_bar[slot + 1], phase);
              tc_fence_after();
              const uint32_t sy = smem_u32(ring + slot * BW3_SLOT);
              const uint32_t d_tmem = tmem_base + n * 128;
              const uint64_t bdesc0 = make_smem_desc_sw128(sy, BW3_SLOT);
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                const uint64_t bdesc = bdesc0 + uint64_t(ks * (2048 >> 4));
                const uint64_t adesc = make_smem_desc_sw128(ga + (2 * jh + (ks >> 2)) * BW3_XCHUNK, 1024) + 2 * (ks & 3);
                mma_ss_pair(d_tmem, adesc, bdesc, idesc_o, !(first && jh == 0 && ks == 0));
              }
              tc_commit_pair(&empty_bar[slot], 3);
              tc_commit_pair(&empty_bar[slot + 1], 3);
              slot += 2;
              if (slot == BW3_SLOTS) { slot = 0; phase ^= 1; }
            }
        };
        mma_s(tile_ctr);
        if (nj == 1) tc_commit_pair(xempty_bar, 3);
        for (int t = 0; t < nj; ++t) {
          if (t + 1 < nj) {
            mma_s(tile_ctr + 1);
            if (t + 2 == nj) tc_commit_pair(xempty_bar, 3);
          }
          if (t == 0) {
            mbar_wait(accempty_bar, (acc_ctr & 1) ^ 1);
            tc_fence_after();
          }
          mbar_wait(&gready_bar[tile_ctr & 1], (tile_ctr >> 1) & 1);
          tc_fence_after();
          mma_out(tile_ctr, t == 0);
          ++tile_ctr;
          if (t == nj - 1) {
            tc_commit_pair(accfull_bar, 3);
            ++acc_ctr;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps 4..11 (both CTAs, own TMEM / own shared memory) =====================
    const int q = warp & 3;               // TMEM lane quarter: rows 32 (q & 1) .., tile columns 64 (q >> 1) ..
    const int wg = (warp - 4) >> 2;       // half of the quarter's kNJ/2 TMEM columns
    uint32_t tile_ctr = 0, acc_ctr = 0;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const int rloc = (q & 1) * 32 + lane;                 // row inside this CTA's 64
    constexpr int NC = kNJ / 128;                         // 32-column chunks per thread and tile
    const int ctile = (q >> 1) * (kNJ / 2) + wg * (kNJ / 4);   // first tile column of this thread
    const uint32_t gready_remote0 = mapa_cluster(smem_u32(&gready_bar[0]), 0);
    const uint32_t gready_remote1 = mapa_cluster(smem_u32(&gready_bar[1]), 0);
    const uint32_t accempty_remote = mapa_cluster(smem_u32(accempty_bar), 0);
    // this thread's bytes of the K-major SWIZZLE_128B G row rloc: tile column cc lives in chunk cc >> 6, 16-byte unit
    // (cc & 63) >> 3 (XOR-swizzled with the row)
    const uint32_t grow_addr = smem_u32(gbuf) + rloc * 128;
    BwThread th;
    th.wg = wg;
    th.ydn = p.ydiag * p.gnorm;
    th.wn = p.wneg_c * p.gnorm;
    th.ign = 1.f / p.gnorm;
    th.nshift2 = -p.shift2;
    Bw3Sched sched;
    sched.init(p, cluster_id, n_clusters);
    int xt, j0, j1;
    while (sched.next(xt, j0, j1)) {
      const int nj = j1 - j0;
      th.row = xt * BW_BM + BW3_XROWS * (int)rank + rloc;
      th.row_ok = th.row < p.Nx;
      th.rs = 0.f;
      if (!BwIsSiglip<kMode>::value) th.rs = bw_stat(p, p.rowscale, th.row, th.row_ok);
      const bool rows_full = xt * BW_BM + BW_BM <= p.Nx;
      double dtacc = 0.0, dlacc = 0.0, dbacc = 0.0;
      // column scales: lane e of every warp fetches the scale of the warp's e-th column one tile ahead (registers),
      // stages it in a warp-private 32-float slot (__syncwarp only, no CTA barrier) and reads it back broadcast
      auto load_cs = [&](int jt, int c) -> float {
        const int col = jt * kNJ + ctile + 32 * c + lane;
        return bw_stat(p, p.colscale, col, col < p.Ny);
      };
      float cs_next[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) cs_next[c] = BwIsSiglip<kMode>::value ? 0.f : load_cs(j0, c);
      for (int t = 0; t < nj; ++t, ++tile_ctr) {
        const int j = j0 + t;
        float tacc = 0.f, lacc = 0.f, bacc = 0.f;
        const int buf = tile_ctr & 1;
        float cs_cur[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          cs_cur[c] = cs_next[c];
          if (!BwIsSiglip<kMode>::value && t + 1 < nj) cs_next[c] = load_cs(j + 1, c);
        }
        mbar_wait(&sfull_bar[buf], (tile_ctr >> 1) & 1);
        tc_fence_after();
        const bool full = rows_full && (j * kNJ + kNJ <= p.Ny);
        const int dlo = xt * BW_BM + p.diag_off - j * kNJ;        // does the diagonal cross this pair tile? (uniform)
        const bool has_diag = !BwIsSiglip<kMode>::value && p.ydiag != 0.f && dlo > -BW_BM && dlo < kNJ;
        const uint32_t cs_addr = smem_u32(col_s) + (warp - 4) * 32 * 4;
        const uint32_t ga = grow_addr + buf * C::kGBuf;
        if (p.gstore) {
          // the TMA store of tile t - 2 (same G buffer) has finished reading shared memory. kNJ = 256: a warp owns a whole
          // [32 x 64] box; kNJ = 128: the two warps of a lane quarter share one (32 columns each), warp wg = 0 stores it
          if (kNJ == 256 || wg == 0) {
            if (lane == 0) bw3_bulk_wait_read1();
            __syncwarp();
          }
          if (kNJ == 128) named_bar_sync(1 + q, 64);
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int cc = ctile + 32 * c;
          uint32_t acc[32];
          tmem_ld32(tmem_base + lane_off + C::kSCol + buf * (kNJ / 2) + wg * (kNJ / 4) + 32 * c, acc);
          if (!BwIsSiglip<kMode>::value) {
            __syncwarp();                                 // previous reads of the warp's slot are done
            col_s[(warp - 4) * 32 + lane] = cs_cur[c];
            __syncwarp();
          }
          tc_wait_ld();
          const int colg0 = j * kNJ + cc;
          int dcol = -1;                                  // diagonal target column relative to these 32 columns
          if (has_diag) {
            const int d = th.row + p.diag_off - colg0;
            dcol = (d >= 0 && d < 32) ? d : -1;
          }
          uint32_t packed[16];
          // (the stable-softmax variants exist for the CLIP / gated modes only; the branch is grid-uniform)
          if (!BwIsSiglip<kMode>::value && p.stable) {
            if (full && !has_diag)
              bw3_g32<kMode, true, !BwIsSiglip<kMode>::value>(p, th, acc, cs_addr, colg0, -1, tacc, lacc, bacc, packed);
            else
              bw3_g32<kMode, false, !BwIsSiglip<kMode>::value>(p, th, acc, cs_addr, colg0, dcol, tacc, lacc, bacc, packed);
          } else if (full && !has_diag) {
            bw3_g32<kMode, true, false>(p, th, acc, cs_addr, colg0, -1, tacc, lacc, bacc, packed);
          } else {
            bw3_g32<kMode, false, false>(p, th, acc, cs_addr, colg0, dcol, tacc, lacc, bacc, packed);
          }
          const uint32_t gc = ga + (cc >> 6) * BW3_XCHUNK;
          const int u0 = (cc & 63) >> 3;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            sts128(gc + (uint32_t((u0 + u) ^ (rloc & 7)) << 4), packed[4 * u], packed[4 * u + 1], packed[4 * u + 2],
                   packed[4 * u + 3]);
        }
        // generic-proxy stores -> visible to the tensor core's (async proxy) reads of THIS CTA's shared memory. The
        // .shared::cta form is enough (each SM's tensor core reads its own CTA's G half) and, unlike the unqualified
        // fence.proxy.async / mbarrier.arrive.release.cluster pair, does not compile to MEMBAR.ALL.GPU + ERRBAR
        // (measured: those two took 40 % of the epilogue's issue slots, profiles/r01d_*).
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (kNJ == 128 && p.gstore) named_bar_sync(1 + q, 64);      // both halves of the box are written and fenced
        if (lane == 0) mbar_arrive_cluster(buf ? gready_remote1 : gready_remote0);
        if (p.gstore && lane == 0 && (kNJ == 256 || wg == 0)) {
          // the warp's own [32 rows x 64 columns] half of chunk ctile / 64: 4 KB, contiguous in shared memory, already in the
          // SWIZZLE_128B operand layout -> copied byte for byte into block (row block 2 xt + rank, column block 4 j + chunk) of
          // the blocked G buffer (8 KB per [64 x 64] block, gt_gemm.cu loads the blocks as its MN-major A operand): 4 KB
          // sequential writes instead of 32 row segments 2 Ny bytes apart
          bw3_tma_store_2d(&tmG, smem_u32(gbuf) + buf * C::kGBuf + (ctile >> 6) * BW3_XCHUNK + (q & 1) * 32 * 128, 0,
                           ((2 * xt + (int)rank) * p.gnjb + (kNJ / 64) * j + (ctile >> 6)) * 64 + (q & 1) * 32);
          bw3_bulk_commit();
        }
        if (p.scal) {
          dtacc += (double)tacc;
          if (BwIsSiglip<kMode>::value) {
            dlacc += (double)lacc;
            dbacc += (double)bacc;
          }
        }
        if (t == nj - 1) {
          // ---- drain: accumulator part n, TMEM columns n*128 + wg*64 + [0, 64) of this lane quarter hold output
          //      columns 256 n + 128 (q >> 1) + 64 wg + [0, 64) of row rloc ----
          mbar_wait(accfull_bar, acc_ctr & 1);
          ++acc_ctr;
          tc_fence_after();
          const float osc = p.out_scale * th.ign;
          const bool vec_ok = (p.ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dX) & 15) == 0;
          for (int n = 0; n < nparts; ++n) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int d0 = 256 * n + 128 * (q >> 1) + 64 * wg + 32 * c;
              if (d0 < p.D) {                     // warp-uniform
                uint32_t a[32];
                tmem_ld32(tmem_base + lane_off + n * 128 + wg * 64 + c * 32, a);
                tc_wait_ld();
                if (th.row_ok) {
                  float* drow = p.dX + (size_t)th.row * p.ldd + d0;
                  if (vec_ok && d0 + 32 <= p.D) {
#pragma unroll
                    for (int e = 0; e < 32; e += 4)
                      red_add_v4(drow + e, __uint_as_float(a[e]) * osc, __uint_as_float(a[e + 1]) * osc,
                                 __uint_as_float(a[e + 2]) * osc, __uint_as_float(a[e + 3]) * osc);
                  } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                      if (d0 + e < p.D) atomicAdd(drow + e, __uint_as_float(a[e]) * osc);
                  }
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(accempty_remote);
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
          atomicAdd(p.scal + 0, dtacc * (double)th.ign);
          if (BwIsSiglip<kMode>::value) {
            atomicAdd(p.scal + 1, dlacc);
            atomicAdd(p.scal + 2, dbacc * (double)th.ign);
          }
        }
      }
    }
  }

  if (p.gstore && warp >= 4 && lane == 0) bw3_bulk_wait_all();
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

template <int kMode, int kNJ>
static int launch_bw3(const CUtensorMap& tmX, const CUtensorMap& tmYs, const CUtensorMap& tmYo, const CUtensorMap& tmG,
                      const BwParams& p, int grid, cudaStream_t stream) {
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device() & 63];   // cudaFuncSetAttribute is per device
  constexpr int smem = Bw3Cfg<kNJ>::kSmem;
  if (!attr_done) {
    if (cudaFuncSetAttribute(bw3_kernel<kMode, kNJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return B2_ECUDA;
    attr_done = true;
  }
  bw3_kernel<kMode, kNJ><<<grid, BW_THREADS, smem, stream>>>(tmX, tmYs, tmYo, tmG, p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

// Same contract as logits_bwd() (host_api.h) for plain bf16 operands with Kp == Dp in {256, 512, 768}; B2_ENOSYS for
// anything else (the caller falls back to the other kernels).
int logits_bwd_pair64(int mode, const void* X, const void* Y, int Nx, int Ny, int Kp, int Dp, int D, int hi_off, int ldx,
                      int ldy, float scale2, float shift2, float inv_tau, float bias, float wneg_c,
                      const float* rowscale, const float* colscale, float out_scale, float gnorm, int hp,
                      const float* dyn, float ydiag, int diag_off, float* diag_corr, float* dX, int ldd, double* scal,
                      int nseg_hint, cudaStream_t stream, void* gstore, long long g_elems) {
  if (hp || hi_off != 0 || Kp != Dp || Dp % 256 || Dp > 768 || sm_count() < 2) return B2_ENOSYS;
  // 256 Y rows per step whenever the X panel and the S buffers allow it (Kp <= 512); B200CLIP_BWD3_NJ=128 forces 128
  static const bool nj256_ok = [] { const char* e = getenv("B200CLIP_BWD3_NJ"); return !(e && e[0] == '1' && e[1] == '2'); }();
  const int NJ = (Kp <= 512 && nj256_ok) ? 256 : 128;
  if (mode != BW_CLIP && mode != BW_GATED && mode != BW_SIGLIP) return B2_ENOSYS;
  if (gstore && g_elems < gstore_elems(Nx, Ny)) return B2_ENOMEM;
  BwParams p;
  p.Nx = Nx; p.Ny = Ny; p.Kp = Kp; p.Dp = Dp; p.D = D; p.hi_off = 0; p.ydiag = ydiag; p.diag_off = diag_off;
  p.diag_corr = diag_corr;
  p.x_tiles = (Nx + BW_BM - 1) / BW_BM;              // 128-row X tiles, one per cluster item
  p.y_tiles = (Ny + NJ - 1) / NJ;
  p.dparts = 1;
  const int clusters = sm_count() / 2;
  // nseg_hint > 0: explicit (X tile, Y segment) items; otherwise balanced contiguous ranges (Bw3Sched)
  int nseg = nseg_hint > 0 ? nseg_hint : 0;
  if (nseg > p.y_tiles) nseg = p.y_tiles;
  p.nseg = nseg;
  p.scale2 = scale2; p.shift2 = shift2; p.inv_tau = inv_tau; p.bias = bias; p.wneg_c = wneg_c;
  p.rowscale = rowscale; p.colscale = colscale; p.out_scale = out_scale;
  p.gnorm = gnorm > 0.f ? gnorm : 1.f;
  p.hp = 0;
  p.lclamp = 30.f; p.yneg = 0.f; p.ent_coef = 0.f; p.stable = 0;
  p.dX = dX; p.ldd = ldd; p.scal = scal; p.dyn = dyn; p.gstore = gstore ? 1 : 0; p.gnjb = 4 * ((Ny + 255) / 256);
  CUtensorMap tmX, tmYs, tmYo, tmG;
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmX, X, Nx, Kp, ldx, BW3_XROWS))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmYs, Y, Ny, Kp, ldy, NJ / 2))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmYo, Y, Ny, Kp, ldy, 128))) return rc;
  if (gstore) {
    if ((rc = make_tmap_bf16_rows64(&tmG, gstore, (uint64_t)(gstore_elems(Nx, Ny) / 64), 32))) return rc;
  } else {
    tmG = tmX;      // never dereferenced
  }
  const long long items = p.nseg > 0 ? (long long)p.x_tiles * p.nseg : (long long)p.x_tiles * p.y_tiles;
  const int grid = 2 * (int)(items < clusters ? items : clusters);
#define LAUNCH3(M) (NJ == 256 ? launch_bw3<M, 256>(tmX, tmYs, tmYo, tmG, p, grid, stream) \
                              : launch_bw3<M, 128>(tmX, tmYs, tmYo, tmG, p, grid, stream))
  if (mode == BW_CLIP) return LAUNCH3(BW_CLIP);
  if (mode == BW_GATED) return LAUNCH3(BW_GATED);
  return LAUNCH3(BW_SIGLIP);
#undef LAUNCH3
}

// Both gradients of the softmax / gated / sigmoid contrastive step from ONE recompute of the logits: dX += G Y as above with every G tile
// stored (bf16, blocked layout of gt_gemm.cu) and dY += G^T X by gt_gemm.cu. Square single-GPU problems whose G fits the caller's buffer; the
// diagonal corrections of the Y side equal those of the X side (same G_ii). B2_ENOSYS when the shape does not qualify.
int logits_bwd_both(int mode, const void* X, const void* Y, int Nx, int Ny, int Kp, int Dp, int D, int ldx, int ldy,
                    float wneg_c, const float* rowscale, const float* colscale, float gnorm, const float* dyn, float ydiag,
                    int diag_off, float* diag_corr, float* dX, int ldd, float* dY, int lddy, double* scal, void* G,
                    long long g_elems, cudaStream_t stream) {
  if (!G || !dY || !dyn || (mode != BW_CLIP && mode != BW_GATED && mode != BW_SIGLIP)) return B2_ENOSYS;
  int rc = logits_bwd_pair64(mode, X, Y, Nx, Ny, Kp, Dp, D, 0, ldx, ldy, 0.f, 0.f, 0.f, 0.f, wneg_c, rowscale, colscale, 0.f,
                             gnorm, 0, dyn, ydiag, diag_off, diag_corr, dX, ldd, scal, 0, stream, G, g_elems);
  if (rc) return rc;
  return gt_gemm(G, g_elems, Nx, Ny, X, ldx, Dp, D, dyn, gnorm, dY, lddy, stream);
}

}  // namespace b2host
