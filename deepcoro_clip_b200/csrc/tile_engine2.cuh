// CTA-pair version of the S = A * B^T tile engine (tile_engine.cuh): tcgen05 cta_group::2 with ONE OPERAND RESIDENT in
// shared memory (all of it for Kp <= 512; for 512 < Kp <= 1024 its first 512 columns, the tail k-chunks of the resident
// operand are re-streamed per tile next to the other operand: 42 B/cycle per SM at Kp = 768, still under the ingest port).
//
// Why: a single SM ingests at most 61.5 B/cycle through TMA (tools/ubench). The single-CTA engine streams a 16 KB A
// chunk and a 32 KB B chunk per 64-wide k-chunk for 512 tensor cycles = 94 B/cycle: ingest-bound (~65 % tensor pipe).
// Here a pair of CTAs computes a 256 x 256 tile per step with M = 256, N = 256 MMAs: each CTA owns 128 A rows and
// supplies 128 of the 256 B rows. The operand of the OUTER loop stays resident (128 KB per CTA), the other one streams
// at 16 KB per k-chunk per CTA = 32 B/cycle.
//
//   kOuterIsB = true  (logits forward): outer = 256-row B block — each CTA keeps ITS 128 rows of the block resident,
//                     inner = pairs of 128-row A tiles, streamed (CTA r takes A tile 2*pair + r).
//   kOuterIsB = false (retrieval)     : outer = pair of A tiles — each CTA keeps ITS A tile resident (per-row epilogue
//                     state), inner = 256-row B blocks, streamed (CTA r loads rows [128 r, 128 r + 128) of the block).
//
// Every CTA's epilogue sees ordinary 128 x 256 tiles (its own TMEM lanes), so the epilogue policies of
// tile_engine.cuh are used unchanged. Barrier protocol as in logits_bwd2.cu: both CTAs' TMA credits the leader's full
// barriers, the leader's MMA lane multicasts its commits, the peer's epilogue arrives remotely on the leader's
// accumulator-empty barriers.
#pragma once
#include "tile_engine.cuh"

namespace b2 {

constexpr int TE2_SLOTS = 6;
constexpr int TE2_SLOT = 16384;
constexpr int TE2_RES_BYTES = 8 * 16384;                          // resident operand: 128 rows x Kp <= 512
constexpr int TE2_BAR_OFF = TE2_RES_BYTES + TE2_SLOTS * TE2_SLOT;
constexpr int TE2_SMEM_BYTES = TE2_BAR_OFF + 256 + 1024;

// Tile sequence of one CLUSTER: the single-CTA TileSeq over (outer, inner) with A tiles replaced by A-tile pairs.
struct TileSeq2 {
  TileSeq s;
  __device__ __forceinline__ void init(const TeShape& g, bool outer_is_b, int cluster, int n_clusters) {
    TeShape h = g;
    h.m_tiles = (g.m_tiles + 1) / 2;
    // TileSeq::init reads blockIdx.x / gridDim.x: re-derive its fields for (cluster, n_clusters)
    s.diag = false;
    s.diag_outer_is_b = outer_is_b;
    s.inner_n = outer_is_b ? h.m_tiles : h.n_blocks;
    const int outer_n = outer_is_b ? h.n_blocks : h.m_tiles;
    s.segs = h.segs;
    if (s.segs > 0) {
      s.items = outer_n * s.segs;
      s.item = cluster;
      s.stride = n_clusters;
      s.i = s.i1 = 0;
    } else {
      const long long total = (long long)h.m_tiles * h.n_blocks;
      s.t = (int)(total * cluster / n_clusters);
      s.t1 = (int)(total * (cluster + 1) / n_clusters);
    }
  }
  __device__ __forceinline__ bool next(int& o, int& in, int& sg) { return s.next(o, in, sg); }
};

template <class Epi, bool kOuterIsB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TE_THREADS, 1)
te2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TeShape g,
           typename Epi::Params ep) {
  if (te_gate_closed(g)) return;     // grid-uniform (both CTAs of every cluster): before any barrier / TMEM allocation
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* res = smem;                             // resident operand panel
  uint8_t* ring = smem + TE2_RES_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TE2_BAR_OFF);
  uint64_t* full_bar = bars;                       // [6] leader: TMA of both CTAs -> MMA
  uint64_t* empty_bar = bars + TE2_SLOTS;          // [6] both: MMA (multicast commit) -> TMA
  uint64_t* tfull_bar = bars + 2 * TE2_SLOTS;      // [2] both: accumulator ready (multicast commit)
  uint64_t* tempty_bar = tfull_bar + 2;            // [2] leader: accumulator drained, 16 arrivals
  uint64_t* rfull_bar = tempty_bar + 2;            // [1] leader: both resident panels landed
  uint64_t* rempty_bar = rfull_bar + 1;            // [1] both: resident panels free (multicast commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rempty_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int kchunks = g.Kp / TE_BK;
  const int rch = kchunks < 8 ? kchunks : 8;       // resident k-chunks of the outer operand (the rest is streamed)
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < TE2_SLOTS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 16);
    }
    mbar_init(rfull_bar, 1);
    mbar_init(rempty_bar, 1);
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

  // coordinates of the streamed / resident 128-row boxes of this CTA for the pair-tile (outer, inner)
  //   kOuterIsB: resident = B rows outer*256 + 128*rank ; streamed = A rows (2*inner + rank)*128
  //   else     : resident = A rows (2*outer + rank)*128 ; streamed = B rows inner*256 + 128*rank
  auto res_row = [&](int outer) { return kOuterIsB ? outer * TE_BN + 128 * (int)rank : (2 * outer + (int)rank) * TE_BM; };
  auto str_row = [&](int inner) { return kOuterIsB ? (2 * inner + (int)rank) * TE_BM : inner * TE_BN + 128 * (int)rank; };
  const CUtensorMap* tmRes = kOuterIsB ? &tmB : &tmA;
  const CUtensorMap* tmStr = kOuterIsB ? &tmA : &tmB;

  if (warp < 4) {
    setmaxnreg_dec<40>();
    if (warp == 0) {
      // ===================== TMA producer (both CTAs) =====================
      if (elect_one()) {
        int slot = 0;
        uint32_t phase = 0, rphase = 0;
        int cur_outer = -1, cur_seg = -1;
        TileSeq2 seq;
        seq.init(g, kOuterIsB, cluster_id, n_clusters);
        int outer, inner, sg;
        while (seq.next(outer, inner, sg)) {
          if (outer != cur_outer || sg != cur_seg) {
            cur_outer = outer;
            cur_seg = sg;
            mbar_wait(rempty_bar, rphase ^ 1);
            rphase ^= 1;
            if (leader) mbar_expect_tx(rfull_bar, 2 * rch * 16384);
            for (int kc = 0; kc < rch; ++kc)
              tma_load_2d_pair(res + kc * 16384, tmRes, rfull_bar, kc * TE_BK, res_row(outer));
          }
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(&empty_bar[slot], phase ^ 1);
            if (leader) mbar_expect_tx(&full_bar[slot], 2 * TE2_SLOT);
            tma_load_2d_pair(ring + slot * TE2_SLOT, tmStr, &full_bar[slot], kc * TE_BK, str_row(inner));
            if (++slot == TE2_SLOTS) { slot = 0; phase ^= 1; }
            if (kc >= rch) {             // tail chunk of the outer operand: not resident, streamed for every tile
              mbar_wait(&empty_bar[slot], phase ^ 1);
              if (leader) mbar_expect_tx(&full_bar[slot], 2 * TE2_SLOT);
              tma_load_2d_pair(ring + slot * TE2_SLOT, tmRes, &full_bar[slot], kc * TE_BK, res_row(outer));
              if (++slot == TE2_SLOTS) { slot = 0; phase ^= 1; }
            }
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (leader only) =====================
      if (leader && elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(256, TE_BN, 0, 0);
        int slot = 0;
        uint32_t phase = 0, rphase = 0;
        int lt = 0;
        int cur_outer = -1, cur_seg = -1;
        const uint32_t res_addr = smem_u32(res);
        TileSeq2 seq;
        seq.init(g, kOuterIsB, cluster_id, n_clusters);
        int outer, inner, sg;
        bool have = seq.next(outer, inner, sg);
        while (have) {
          if (outer != cur_outer || sg != cur_seg) {
            cur_outer = outer;
            cur_seg = sg;
            mbar_wait(rfull_bar, rphase);
            rphase ^= 1;
            tc_fence_after();
          }
          const int as = lt & 1;
          const uint32_t aphase = (lt >> 1) & 1;
          mbar_wait(&tempty_bar[as], aphase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * TE_BN;
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(&full_bar[slot], phase);
            tc_fence_after();
            const uint32_t st = smem_u32(ring + slot * TE2_SLOT);
            uint32_t rs = res_addr + kc * 16384;
            const int slot0 = slot;
            if (++slot == TE2_SLOTS) { slot = 0; phase ^= 1; }
            int slot1 = -1;
            if (kc >= rch) {
              mbar_wait(&full_bar[slot], phase);
              tc_fence_after();
              rs = smem_u32(ring + slot * TE2_SLOT);
              slot1 = slot;
              if (++slot == TE2_SLOTS) { slot = 0; phase ^= 1; }
            }
            const uint64_t adesc = make_smem_desc_sw128(kOuterIsB ? st : rs, 1024);
            const uint64_t bdesc = make_smem_desc_sw128(kOuterIsB ? rs : st, 1024);
#pragma unroll
            for (int k = 0; k < TE_BK / 16; ++k) mma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) != 0);
            tc_commit_pair(&empty_bar[slot0], 3);
            if (slot1 >= 0) tc_commit_pair(&empty_bar[slot1], 3);
          }
          tc_commit_pair(&tfull_bar[as], 3);
          ++lt;
          // next tile: if the resident panel changes, it becomes free once every MMA issued so far has completed
          int no, ni, ns;
          have = seq.next(no, ni, ns);
          if (!have || no != cur_outer || ns != cur_seg) tc_commit_pair(rempty_bar, 3);
          outer = no; inner = ni; sg = ns;
        }
      }
    }
  } else {
    // ===================== epilogue warps 4..11 (both CTAs, own TMEM lanes) =====================
    setmaxnreg_inc<232>();
    const int q = warp & 3;
    const int wg = (warp - 4) >> 2;
    const uint32_t tempty_remote0 = mapa_cluster(smem_u32(&tempty_bar[0]), 0);
    const uint32_t tempty_remote1 = mapa_cluster(smem_u32(&tempty_bar[1]), 0);
    typename Epi::State st;
    Epi::init(st, ep);
    TeCtx ctx, last;
    ctx.wg = wg;
    ctx.Nb = g.Nb;
    ctx.seg = 0;
    last = ctx;
    int cur_outer = -1, cur_seg = -1;
    int lt = 0;
    TileSeq2 seq;
    seq.init(g, kOuterIsB, cluster_id, n_clusters);
    int outer, inner, sg;
    for (; seq.next(outer, inner, sg); ++lt) {
      // this CTA's view: an ordinary (m_tile, n_block) tile of the single-CTA engine
      ctx.m_tile = 2 * (kOuterIsB ? inner : outer) + (int)rank;
      ctx.n_block = kOuterIsB ? outer : inner;
      ctx.row = ctx.m_tile * TE_BM + q * 32 + lane;
      ctx.col0 = ctx.n_block * TE_BN + wg * 128;
      ctx.row_ok = ctx.row < g.Ma;
      ctx.full = (ctx.m_tile * TE_BM + TE_BM <= g.Ma) && (ctx.n_block * TE_BN + TE_BN <= g.Nb);
      // the epilogue's "outer" index: n_block (kOuterIsB) or this CTA's own m_tile
      const int my_outer = kOuterIsB ? outer : ctx.m_tile;
      if (outer != cur_outer || sg != cur_seg) {
        if (cur_outer >= 0) Epi::end_outer(st, ep, kOuterIsB ? cur_outer : last.m_tile, last);
        cur_outer = outer;
        cur_seg = sg;
        ctx.seg = sg;
        Epi::begin_outer(st, ep, my_outer, ctx);
      }
      const int as = lt & 1;
      const uint32_t aphase = (lt >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + as * TE_BN + wg * 128;
      uint32_t acc[2][32];
      tmem_ld32(taddr, acc[0]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tc_wait_ld();
        if (c < 3) tmem_ld32(taddr + (c + 1) * 32, acc[(c + 1) & 1]);
        Epi::chunk(st, ep, ctx, c, acc[c & 1]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(as ? tempty_remote1 : tempty_remote0);
      Epi::end_tile(st, ep, ctx);
      last = ctx;
    }
    if (cur_outer >= 0) Epi::end_outer(st, ep, kOuterIsB ? cur_outer : last.m_tile, last);
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
