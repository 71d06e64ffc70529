// Coordinates handed to the epilogue policies of the tile engines (tile_engine.cuh / tile_engine2.cuh). Kept free of any
// other include so that policy headers can also be compiled for the host by the emulation tests (tests/emul/).
#pragma once

namespace b2 {

struct TeCtx {
  int m_tile;     // 128-row tile index of A
  int n_block;    // 256-row block index of B
  int row;        // global A row owned by this thread (m_tile*128 + lane quarter*32 + lane)
  int col0;       // first global B row (output column) of this thread's 128-column half
  int wg;         // epilogue warpgroup 0/1
  int seg;        // segment of the inner sweep (item schedule), 0 otherwise
  bool row_ok;    // row < Ma
  bool full;      // whole 128x256 tile inside [Ma, Nb]
  int Nb;
};

}  // namespace b2
