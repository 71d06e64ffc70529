// Persistent, warp-specialised S = A * B^T tile engine for sm_100a.
//
//   A [Ma, Kp] bf16 row-major (K contiguous), B [Nb, Kp] bf16 row-major, Kp a multiple of 64.
//   One CTA owns a contiguous range of 128 x 256 output tiles. Per tile the K loop streams 64-wide
//   k-chunks of A (16 KB) and B (32 KB) through a 4-stage TMA/mbarrier ring; one elected thread issues
//   tcgen05.mma (M=128, N=256, K=16) into one of two 256-column TMEM accumulator buffers; eight epilogue
//   warps drain the other buffer with tcgen05.ld and hand every 32-column chunk to an epilogue policy.
//   The S tile never leaves TMEM/registers.
//
// Tile order: kOuterIsB = true  -> outer index = 256-wide B block, inner = 128-row A tile (column state
//                                  such as per-column sums stays in registers across the inner sweep);
//             kOuterIsB = false -> outer = A tile, inner = B block (per-row state, e.g. top-k lists).
//
// Epilogue policy contract (all __device__):
//   struct Params;                               // trivially copyable kernel argument
//   struct State;                                // per-thread registers, lives across the whole CTA range
//   static void init(State&, const Params&);
//   static void begin_outer(State&, const Params&, int outer, const Ctx&);
//   static void chunk(State&, const Params&, const Ctx&, int c, const uint32_t (&acc)[32]);   // c = 0..3
//   static void end_tile(State&, const Params&, const Ctx&);
//   static void end_outer(State&, const Params&, int outer, const Ctx&);
// Ctx carries the tile coordinates of the calling thread (one TMEM lane = one A row, 128 columns per
// epilogue warpgroup).
#pragma once
#include "common.cuh"
#include "te_ctx.cuh"

namespace b2 {

constexpr int TE_BM = 128;
constexpr int TE_BN = 256;
constexpr int TE_BK = 64;
constexpr int TE_STAGES = 4;
constexpr int TE_A_BYTES = TE_BM * TE_BK * 2;   // 16 KB
constexpr int TE_B_BYTES = TE_BN * TE_BK * 2;   // 32 KB
constexpr int TE_STAGE_BYTES = TE_A_BYTES + TE_B_BYTES;
constexpr int TE_THREADS = 384;
constexpr int TE_SMEM_BYTES = TE_STAGES * TE_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct TeShape {
  int Ma;        // valid rows of A
  int Nb;        // valid rows of B
  int Kp;        // padded K (multiple of 64)
  int m_tiles;   // ceil(Ma / 128)
  int n_blocks;  // ceil(Nb / 256)
  int segs;      // 0: each CTA takes one contiguous range of the (outer, inner) tile order (long sweeps);
                 // s > 0: work items = (outer, segment of the inner sweep), items dealt round-robin to CTAs
  int diag;      // 1: only the tiles crossed by the main diagonal (tile t = A tile t x B block t/2); segs must be 0
  // Device-side launch gate (optional): the whole grid returns at once unless (*gate != 0) == (gate_on != 0). Lets the host
  // enqueue both variants of a pass whose choice depends on a device scalar (dyn[11]: stable softmax mode, decided from a
  // learnable temperature) without ever reading that scalar.
  const float* gate;
  int gate_on;
};
__device__ __forceinline__ bool te_gate_closed(const TeShape& g) {
  return g.gate != nullptr && ((*g.gate != 0.f) != (g.gate_on != 0));
}

// Identical tile sequence for the producer, MMA and epilogue roles of one CTA.
struct TileSeq {
  int inner_n, segs, items, item, stride;   // item mode
  int t, t1;                                // contiguous mode
  int i, i1, outer, seg;
  bool diag, diag_outer_is_b;
  __device__ __forceinline__ void init(const TeShape& g, bool outer_is_b) {
    diag = g.diag != 0;
    diag_outer_is_b = outer_is_b;
    inner_n = outer_is_b ? g.m_tiles : g.n_blocks;
    const int outer_n = outer_is_b ? g.n_blocks : g.m_tiles;
    segs = g.segs;
    if (segs > 0) {
      items = outer_n * segs;
      item = blockIdx.x;
      stride = gridDim.x;
      i = i1 = 0;
    } else {
      const long long total = diag ? (long long)g.m_tiles : (long long)g.m_tiles * g.n_blocks;
      t = (int)(total * blockIdx.x / gridDim.x);
      t1 = (int)(total * (blockIdx.x + 1) / gridDim.x);
    }
  }
  // returns false when the CTA is done; (outer, inner, seg) describe the next tile
  __device__ __forceinline__ bool next(int& o, int& in, int& sg) {
    if (segs > 0) {
      while (i >= i1) {
        if (item >= items) return false;
        outer = item / segs;
        seg = item - outer * segs;
        i = (int)((long long)inner_n * seg / segs);
        i1 = (int)((long long)inner_n * (seg + 1) / segs);
        item += stride;
      }
      o = outer; in = i++; sg = seg;
      return true;
    }
    if (t >= t1) return false;
    if (diag) {
      const int mt = t, nb = (t * TE_BM) / TE_BN;
      o = diag_outer_is_b ? nb : mt;
      in = diag_outer_is_b ? mt : nb;
      sg = 0;
      ++t;
      return true;
    }
    o = t / inner_n;
    in = t - o * inner_n;
    sg = 0;
    ++t;
    return true;
  }
};

template <class Epi, bool kOuterIsB>
__global__ void __launch_bounds__(TE_THREADS, 1)
te_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TeShape g,
          typename Epi::Params ep) {
  if (te_gate_closed(g)) return;     // grid-uniform: before any barrier / TMEM allocation
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TE_STAGES * TE_STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [TE_STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + TE_STAGES;       // [TE_STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * TE_STAGES;   // [2]          MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]          epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  // warp index through a shuffle: the compiler then knows the role branches are warp-uniform and keeps the MMA /
  // TMA operands in uniform registers (a `lane == 0` branch makes it wrap every UTCHMMA in a ~100-cycle waterfall loop)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  const int kchunks = g.Kp / TE_BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < TE_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8);   // one arrival per epilogue warp
    }
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

  if (warp < 4) {
    setmaxnreg_dec<40>();
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        int stage = 0;
        uint32_t phase = 0;
        TileSeq seq;
        seq.init(g, kOuterIsB);
        int outer, inner, sg;
        while (seq.next(outer, inner, sg)) {
          const int m_tile = kOuterIsB ? inner : outer;
          const int n_block = kOuterIsB ? outer : inner;
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = stage_base + stage * TE_STAGE_BYTES;
            uint8_t* sb = sa + TE_A_BYTES;
            mbar_expect_tx(&full_bar[stage], TE_STAGE_BYTES);
            tma_load_2d(sa, &tmA, &full_bar[stage], kc * TE_BK, m_tile * TE_BM);
            tma_load_2d(sb, &tmB, &full_bar[stage], kc * TE_BK, n_block * TE_BN);
            if (++stage == TE_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      if (elect_one()) {
        constexpr uint32_t idesc = make_idesc_bf16(TE_BM, TE_BN, 0, 0);
        int stage = 0;
        uint32_t phase = 0;
        int lt = 0;
        TileSeq seq;
        seq.init(g, kOuterIsB);
        int outer, inner, sg;
        for (; seq.next(outer, inner, sg); ++lt) {
          const int as = lt & 1;
          const uint32_t aphase = (lt >> 1) & 1;
          mbar_wait(&tempty_bar[as], aphase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * TE_BN;
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(stage_base + stage * TE_STAGE_BYTES);
            const uint64_t adesc = make_smem_desc_sw128(sa, 0);
            const uint64_t bdesc = make_smem_desc_sw128(sa + TE_A_BYTES, 0);
#pragma unroll
            for (int k = 0; k < TE_BK / 16; ++k) {
              // +32 bytes per K=16 step inside the 128-byte swizzle row: start-address field += 2
              mma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kc | k) != 0);
            }
            tc_commit(&empty_bar[stage]);
            if (++stage == TE_STAGES) { stage = 0; phase ^= 1; }
          }
          tc_commit(&tfull_bar[as]);
        }
      }
    }
  } else {
    // ===================== epilogue warps 4..11 =====================
    setmaxnreg_inc<232>();
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int wg = (warp - 4) >> 2;    // column half
    typename Epi::State st;
    Epi::init(st, ep);
    TeCtx ctx, last;
    ctx.wg = wg;
    ctx.Nb = g.Nb;
    ctx.seg = 0;
    last = ctx;
    int cur_outer = -1, cur_seg = -1;
    int lt = 0;
    TileSeq seq;
    seq.init(g, kOuterIsB);
    int outer, inner, sg;
    for (; seq.next(outer, inner, sg); ++lt) {
      ctx.m_tile = kOuterIsB ? inner : outer;
      ctx.n_block = kOuterIsB ? outer : inner;
      ctx.row = ctx.m_tile * TE_BM + q * 32 + lane;
      ctx.col0 = ctx.n_block * TE_BN + wg * 128;
      ctx.row_ok = ctx.row < g.Ma;
      ctx.full = (ctx.m_tile * TE_BM + TE_BM <= g.Ma) && (ctx.n_block * TE_BN + TE_BN <= g.Nb);
      if (outer != cur_outer || sg != cur_seg) {
        if (cur_outer >= 0) Epi::end_outer(st, ep, cur_outer, last);   // coordinates of the finished sweep
        cur_outer = outer;
        cur_seg = sg;
        ctx.seg = sg;
        Epi::begin_outer(st, ep, outer, ctx);
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
      // all TMEM reads of this buffer are complete (wait::ld above): release it to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      Epi::end_tile(st, ep, ctx);
      last = ctx;
    }
    if (cur_outer >= 0) Epi::end_outer(st, ep, cur_outer, last);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace b2
