// Host build of the library's retrieval epilogue policies / selection kernels and of the alignment scalar kernel under
// the emulation shim (cuda_emul.h). The tile engine itself (tcgen05 / TMA) is replaced by `run_epi`, which follows its
// contract exactly (tile_engine.cuh: TileSeq item mode, TeCtx, begin_outer / chunk x 4 / end_tile / end_outer per
// thread = (lane quarter q, lane, warpgroup wg); out-of-range rows and columns of a tile read 0 as TMA zero-fills them).
#include "cuda_emul.h"
#include "../../deepcoro_clip_b200/csrc/retrieval_epi.cuh"
#include "../../deepcoro_clip_b200/csrc/alignment_diag.cuh"

using namespace b2;

template <class Epi>
static void run_epi(const float* S, int Ma, int Nb, int segs, const typename Epi::Params& p) {
  const int m_tiles = (Ma + 127) / 128, n_blocks = (Nb + 255) / 256;
  for (int m_tile = 0; m_tile < m_tiles; ++m_tile)
    for (int seg = 0; seg < segs; ++seg) {
      const int i0 = (int)((long long)n_blocks * seg / segs), i1 = (int)((long long)n_blocks * (seg + 1) / segs);
      if (i0 >= i1) continue;                       // TileSeq::next skips empty segments
      for (int wg = 0; wg < 2; ++wg)
        for (int q = 0; q < 4; ++q)
          for (int lane = 0; lane < 32; ++lane) {
            typename Epi::State st;
            Epi::init(st, p);
            TeCtx ctx;
            ctx.wg = wg;
            ctx.Nb = Nb;
            ctx.seg = seg;
            ctx.m_tile = m_tile;
            ctx.row = m_tile * 128 + q * 32 + lane;
            ctx.row_ok = ctx.row < Ma;
            for (int nb = i0; nb < i1; ++nb) {
              ctx.n_block = nb;
              ctx.col0 = nb * 256 + wg * 128;
              ctx.full = (m_tile * 128 + 128 <= Ma) && (nb * 256 + 256 <= Nb);
              if (nb == i0) Epi::begin_outer(st, p, m_tile, ctx);
              for (int c = 0; c < 4; ++c) {
                uint32_t acc[32];
                for (int e = 0; e < 32; ++e) {
                  const int col = ctx.col0 + c * 32 + e;
                  const float v = (ctx.row_ok && col < Nb) ? S[(size_t)ctx.row * Nb + col] : 0.f;
                  acc[e] = __float_as_uint(v);
                }
                Epi::chunk(st, p, ctx, c, acc);
              }
              Epi::end_tile(st, p, ctx);
            }
            Epi::end_outer(st, p, m_tile, ctx);
          }
    }
}

extern "C" {

// The opt-in two-sweep top-k exactly as retrieval_metrics_streaming._topk_two_sweeps drives it. Returns the overflow flag.
int emul_topk_two_sweeps(const float* S, int N, int M, int k, int segs, int cap, int col_offset, float* part_max,
                         float* thr, int* cnt, float* buf_s, int* buf_i, float* out_s, long long* out_i) {
  ColMaxParams pc{part_max, 2 * segs};
  run_epi<ColMaxEpi>(S, N, M, segs, pc);
  emul::launch((N + 7) / 8, 256, [&] { kth_largest_kernel(part_max, N, 2 * segs * 32, k, thr); });
  int overflow = 0;
  CollectParams pk{thr, col_offset, cnt, buf_s, buf_i, cap, &overflow};
  run_epi<CollectEpi>(S, N, M, segs, pk);
  if (overflow) return 1;
  emul::launch((N + 7) / 8, 256, [&] { topk_merge_kernel(buf_s, buf_i, N, cap, k, out_s, out_i); });
  return 0;
}

// The validated register-list path (RetrEpi<16> partial lists + topk_merge) through the same driver, as a cross-check of
// the emulation itself against behaviour that has been measured on the GPU.
void emul_topk_register_lists(const float* S, int N, int M, int k, int segs, int col_offset, float* part_s, int* part_i,
                              float* out_s, long long* out_i) {
  RetrParams p{nullptr, nullptr, col_offset, nullptr, part_s, part_i, 2 * segs, k};
  run_epi<RetrEpi<16>>(S, N, M, segs, p);
  emul::launch((N + 7) / 8, 256, [&] { topk_merge_kernel(part_s, part_i, N, 2 * segs * k, k, out_s, out_i); });
}

// Rank counts of RetrEpi<0> (validated on the GPU) through the same driver.
void emul_rank_counts(const float* S, int N, int M, int segs, const float* sgt, const long long* gt, int col_offset,
                      int* counts) {
  RetrParams p{sgt, gt, col_offset, counts, nullptr, nullptr, 2 * segs, 0};
  run_epi<RetrEpi<0>>(S, N, M, segs, p);
}

void emul_alignment_diag(const float* sums, int n, const float* dyn, int gated, float* out) {
  emul::launch(1, 1024, [&] { alignment_diag_kernel(sums, n, dyn, gated, out); });
}

}  // extern "C"
