// Host build of the attention-pool MMA kernels (deepcoro_clip_b200/csrc/attnpool_mma_kernels.cuh, the file that ships in
// libb200clip.so) against the emulated primitives. Launch geometry and shared-memory sizing follow attnpool_mma.cu.
#include "pool_mma_prims_emul.h"
#include "../../deepcoro_clip_b200/csrc/attnpool_mma_kernels.cuh"

using namespace b2;

static size_t fwd_smem(int D, int stages) {
  return (size_t)stages * PM_TT * D * 2 + 2 * 8 * (D + 8) * 2 + 2 * 8 * (PM_TT + 8) * 2 + 8 * PM_TT * 8 * 4 + 8 * 4 + 64 + 2048;
}
static size_t bwd_smem(int D, int stages) {
  return (size_t)stages * PM_TT * D * 2 + 4 * 8 * (D + 8) * 2 + (size_t)D * 48 + 2 * PM_TT * 16 * 2 + 8 * PM_TT * 16 * 4 + 3 * 8 * 4 +
         64 + 2048;
}

extern "C" {

// x [B, N, D] bf16 (dtype 1) or fp16 (2), contiguous. softmax mode (w == null) or given weights w [B, H, N].
void emul_pool_fwd(const void* x, int dtype, const unsigned char* mask, const float* qt, const float* w, int B, int N, int D,
                   int H, int S, int NW, int stages, float* part_m, float* part_l, float* part_acc) {
  PmFwdParams p{x, (long long)N * D, (long long)D, mask, (long long)N, qt, w, (long long)H * N, (long long)N, part_m, part_l,
                part_acc, B, N, D, H, S, 0.f, 0ull, nullptr, stages};
  CUtensorMap tm{x, (uint64_t)B * N, (uint64_t)D, (long long)D};
  const emul::Dim grid{(unsigned)B, (unsigned)S, 1};
  const size_t smem = fwd_smem(D, stages);
  if (dtype == 1) {
    if (NW == 8) emul::launch(grid, 256, [&] { pool_fwd_mma_kernel<__nv_bfloat16, 8>(tm, p); }, smem);
    else emul::launch(grid, 512, [&] { pool_fwd_mma_kernel<__nv_bfloat16, 16>(tm, p); }, smem);
  } else {
    if (NW == 8) emul::launch(grid, 256, [&] { pool_fwd_mma_kernel<__half, 8>(tm, p); }, smem);
    else emul::launch(grid, 512, [&] { pool_fwd_mma_kernel<__half, 16>(tm, p); }, smem);
  }
}

// part_dq != null: the fused-dq instantiation (kDq = true), part_dq [B, S, H, D] zeroed by the caller
void emul_pool_bwd(const void* x, int dtype, const unsigned char* mask, const float* qt, const float* dxbar,
                   const float* xbar, const float* m, const float* l, int B, int N, int D, int H, int S, int NW, int stages,
                   void* dx, float* ds, const float* dlse, float* part_dq) {
  PmBwdParams p{x, (long long)N * D, (long long)D, mask, (long long)N, qt, dxbar, xbar, m, l, dx, ds, B, N, D, H, S,
                nullptr, nullptr, 0.f, 0ull, stages, dlse, part_dq};
  CUtensorMap tm{x, (uint64_t)B * N, (uint64_t)D, (long long)D};
  const emul::Dim grid{(unsigned)B, (unsigned)S, 1};
  const size_t smem = bwd_smem(D, stages) + 2 * 8 * (PM_TT + 8) * 2;
  if (part_dq) {
    if (dtype == 1) {
      if (NW == 8) emul::launch(grid, 256, [&] { pool_bwd_mma_kernel<__nv_bfloat16, 8, true>(tm, p); }, smem);
      else emul::launch(grid, 512, [&] { pool_bwd_mma_kernel<__nv_bfloat16, 16, true>(tm, p); }, smem);
    } else {
      if (NW == 8) emul::launch(grid, 256, [&] { pool_bwd_mma_kernel<__half, 8, true>(tm, p); }, smem);
      else emul::launch(grid, 512, [&] { pool_bwd_mma_kernel<__half, 16, true>(tm, p); }, smem);
    }
    return;
  }
  if (dtype == 1) {
    if (NW == 8) emul::launch(grid, 256, [&] { pool_bwd_mma_kernel<__nv_bfloat16, 8>(tm, p); }, smem);
    else emul::launch(grid, 512, [&] { pool_bwd_mma_kernel<__nv_bfloat16, 16>(tm, p); }, smem);
  } else {
    if (NW == 8) emul::launch(grid, 256, [&] { pool_bwd_mma_kernel<__half, 8>(tm, p); }, smem);
    else emul::launch(grid, 512, [&] { pool_bwd_mma_kernel<__half, 16>(tm, p); }, smem);
  }
}

}  // extern "C"
