// dY[Ny, D] += alpha * G^T X : the second gradient of the contrastive step from the STORED gradient tile matrix.
//
// logits_bwd3.cu forms every G tile (recompute S, softmax both ways, round to bf16) for dX += G Y. The text-side gradient
// dY = G^T X used to be a second launch of the same kernel with the operands swapped — a second S recompute, 4 B N D executed
// FLOP for 2 B N D of work. When the whole G fits in HBM (2 B N bytes: 2 GB at 32k x 32k, single-GPU case) the first launch
// stores its G tiles with TMA and this kernel does the plain product: executed work of the step 8 instead of 10 B N D
// (SURVEY 8d counts 6).
//
// G layout (bf16): [64 x 64] blocks of 8 KB, block (ib, jb) = rows 64 ib .., columns 64 jb .. at element offset
// (ib * nJB + jb) * 4096 with nJB = 4 * ceil(Ny / 256) and ib < 2 * ceil(Nx / 128); inside a block row r holds its 64 columns
// as eight 16-byte units, unit u at position u ^ (r & 7) — the SWIZZLE_128B operand image logits_bwd3.cu already builds in
// shared memory, so its store and this kernel's load are byte-for-byte copies of 4 / 16 KB (row-major G measured 0.33 ms
// slower in the storing kernel: 32 row segments 2 Ny bytes apart per box). Blocks past the edges hold zeros.
//
//   cluster (2 CTAs) = one 256-row tile of dY (256 columns j of G); CTA rank r owns rows [128 r, 128 r + 128)
//   tcgen05.mma.cta_group::2, M = 256, N = 256 per accumulator part (Dp / 256 parts: the whole output width in TMEM),
//   K = 64 rows i of G per stage:
//     A = G[i0 .. i0+63, j0 + 128 r .. +127] : MN-major (j contiguous), two consecutive [64 i x 64 j] blocks, LBO 8 KB
//     B = X[i0 .. i0+63, 256 n + 128 r .. +127] : MN-major (d contiguous), two boxes per part — each CTA feeds half of N
//   The flattened (tile, k-step) sequence is cut into one contiguous, equally long range per cluster (perfect balance:
//   128 tiles on 74 clusters would otherwise be 1.73 waves); every piece drains with red.global.add into the zeroed dY.
#include "bwd_common.cuh"
#include "host_api.h"

namespace b2 {

constexpr int GG_BK = 64;
constexpr int GG_BOX = 64 * 64 * 2;        // 8 KB: [64 i x 64 columns] bf16
constexpr int GG_STAGE = 6 * GG_BOX;       // A: 2 boxes | B: 2 parts x 2 boxes
constexpr int GG_STAGES = 4;
constexpr int GG_THREADS = 384;
constexpr int GG_SMEM = GG_STAGES * GG_STAGE + 256 + 1024;

struct GgParams {
  int Nx, Ny, Dp, D;
  float inv_gnorm;
  const float* dyn;      // dyn[2] = 1 / tau (device)
  float* dY;
  int ldd;
  int tiles, ksteps, njb;
  int d_off;             // first output column of this launch (Dp = its width: 256 or 512)
  // up to three (G, X) source pairs summed into the same accumulator (split-precision products: hi*hi + hi*lo + lo*hi):
  // source s reads the G block rows g_blk_off[s] + k and the X columns x_col_off[s] + d
  int nsrc, g_blk_off[3], x_col_off[3];
};

struct GgSched {
  long long r, r1;
  int ksteps;
  __device__ __forceinline__ void init(const GgParams& p, int cid, int ncl) {
    ksteps = p.ksteps * p.nsrc;
    const long long total = (long long)p.tiles * ksteps;
    r = total * cid / ncl;
    r1 = total * (cid + 1) / ncl;
  }
  __device__ __forceinline__ bool next(int& tile, int& k0, int& k1) {
    if (r >= r1) return false;
    tile = (int)(r / ksteps);
    k0 = (int)(r - (long long)tile * ksteps);
    const long long left = r1 - r;
    k1 = left < (long long)(ksteps - k0) ? k0 + (int)left : ksteps;
    r += k1 - k0;
    return true;
  }
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GG_THREADS, 1)
gt_gemm_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, GgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GG_STAGES * GG_STAGE);
  uint64_t* full_bar = bars;                       // [4] leader only: TMA of both CTAs -> MMA
  uint64_t* empty_bar = bars + GG_STAGES;          // [4] both: MMA (multicast commit) -> TMA
  uint64_t* accfull_bar = empty_bar + GG_STAGES;   // [1] both
  uint64_t* accempty_bar = accfull_bar + 1;        // [1] leader only, 16 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int nparts = p.Dp / 256;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < GG_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accfull_bar, 1);
    mbar_init(accempty_bar, 16);
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

  if (warp == 0) {
    if (elect_one()) {
      int slot = 0;
      uint32_t phase = 0;
      GgSched sched;
      sched.init(p, cluster_id, n_clusters);
      int tile, k0, k1;
      while (sched.next(tile, k0, k1)) {
        const int j0 = tile * 256 + 128 * (int)rank;
        for (int ks = k0; ks < k1; ++ks) {
          mbar_wait(&empty_bar[slot], phase ^ 1);
          uint8_t* st = smem + slot * GG_STAGE;
          if (leader) mbar_expect_tx(&full_bar[slot], 2 * (2 + 2 * nparts) * GG_BOX);
          const int src = ks / p.ksteps, kk = ks - src * p.ksteps, i0 = kk * GG_BK;
          tma_load_2d_pair(st, &tmG, &full_bar[slot], 0, ((p.g_blk_off[src] + kk) * p.njb + (j0 >> 6)) * 64);   // two blocks, 16 KB
          for (int n = 0; n < nparts; ++n) {
            const int d0 = p.x_col_off[src] + p.d_off + 256 * n + 128 * (int)rank;
            tma_load_2d_pair(st + (2 + 2 * n) * GG_BOX, &tmX, &full_bar[slot], d0, i0);
            tma_load_2d_pair(st + (3 + 2 * n) * GG_BOX, &tmX, &full_bar[slot], d0 + 64, i0);
          }
          if (++slot == GG_STAGES) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(256, 256, 1, 1);       // A and B MN-major; 128 rows of A, 128 columns of B per CTA
      int slot = 0;
      uint32_t phase = 0, acc_ctr = 0;
      GgSched sched;
      sched.init(p, cluster_id, n_clusters);
      int tile, k0, k1;
      while (sched.next(tile, k0, k1)) {
        mbar_wait(accempty_bar, (acc_ctr & 1) ^ 1);
        tc_fence_after();
        for (int ks = k0; ks < k1; ++ks) {
          mbar_wait(&full_bar[slot], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + slot * GG_STAGE);
          const uint64_t adesc0 = make_smem_desc_sw128(st, GG_BOX);
          for (int n = 0; n < nparts; ++n) {
            const uint64_t bdesc0 = make_smem_desc_sw128(st + (2 + 2 * n) * GG_BOX, GG_BOX);
#pragma unroll
            for (int k = 0; k < GG_BK / 16; ++k)
              mma_ss_pair(tmem_base + n * 256, adesc0 + uint64_t(k * (2048 >> 4)), bdesc0 + uint64_t(k * (2048 >> 4)), idesc,
                          !(ks == k0 && k == 0));
          }
          tc_commit_pair(&empty_bar[slot], 3);
          if (++slot == GG_STAGES) { slot = 0; phase ^= 1; }
        }
        tc_commit_pair(accfull_bar, 3);
        ++acc_ctr;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3, half = (warp - 4) >> 2;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t accempty_remote = mapa_cluster(smem_u32(accempty_bar), 0);
    const float alpha = p.dyn[2] * p.inv_gnorm;
    const bool vec_ok = (p.ldd & 3) == 0 && (reinterpret_cast<uintptr_t>(p.dY) & 15) == 0;
    uint32_t acc_ctr = 0;
    GgSched sched;
    sched.init(p, cluster_id, n_clusters);
    int tile, k0, k1;
    while (sched.next(tile, k0, k1)) {
      const int row = tile * 256 + 128 * (int)rank + 32 * q + lane;
      mbar_wait(accfull_bar, acc_ctr & 1);
      ++acc_ctr;
      tc_fence_after();
      for (int n = 0; n < nparts; ++n) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int d0 = p.d_off + 256 * n + 128 * half + 32 * c;
          if (d0 < p.D) {                        // warp-uniform
            uint32_t a[32];
            tmem_ld32(tmem_base + lane_off + n * 256 + half * 128 + c * 32, a);
            tc_wait_ld();
            if (row < p.Ny) {
              float* drow = p.dY + (size_t)row * p.ldd + d0;
              if (vec_ok && d0 + 32 <= p.D) {
#pragma unroll
                for (int e = 0; e < 32; e += 4)
                  red_add_v4(drow + e, __uint_as_float(a[e]) * alpha, __uint_as_float(a[e + 1]) * alpha,
                             __uint_as_float(a[e + 2]) * alpha, __uint_as_float(a[e + 3]) * alpha);
              } else {
#pragma unroll
                for (int e = 0; e < 32; ++e)
                  if (d0 + e < p.D) atomicAdd(drow + e, __uint_as_float(a[e]) * alpha);
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

long long gstore_elems(int Nx, int Ny) {
  return 2ll * ((Nx + 127) / 128) * 4ll * ((Ny + 255) / 256) * 4096ll;
}

// dY[Ny, D] (fp32, pitch ldd, pre-zeroed or holding a partial sum) += dyn[2] / gnorm * sum_s G_s^T X_s ; the G_s are block-row
// ranges of ONE blocked buffer (g_elems_total elements, 128-byte aligned; source s starts at block row g_blk_off[s], each
// holding 2 ceil(Nx / 128) block rows), the X_s column ranges [x_col_off[s], x_col_off[s] + Dp) of ONE bf16 matrix
// [Nx, x_cols] with pitch ldx; Dp in {256, 512, 768}. B2_ENOSYS for other widths.
int gt_gemm_multi(const void* G, long long g_elems_total, int nsrc, const int* g_blk_off, int Nx, int Ny, const void* X,
                  int ldx, int x_cols, const int* x_col_off, int Dp, int D, const float* dyn, float gnorm, float* dY, int ldd,
                  cudaStream_t stream) {
  if (Dp % 256 || Dp > 768 || sm_count() < 2) return B2_ENOSYS;
  if (!G || !X || !dyn || !dY || Nx < 1 || Ny < 1 || D < 1 || D > Dp || nsrc < 1 || nsrc > 3) return B2_EINVAL;
  const long long per = gstore_elems(Nx, Ny);
  GgParams p;
  p.Nx = Nx; p.Ny = Ny; p.D = D; p.inv_gnorm = 1.f / (gnorm > 0.f ? gnorm : 1.f); p.dyn = dyn; p.dY = dY; p.ldd = ldd;
  p.tiles = (Ny + 255) / 256;
  p.ksteps = 2 * ((Nx + 127) / 128);      // every row block the storing kernel wrote (zeros past Nx)
  p.njb = 4 * p.tiles;
  p.nsrc = nsrc;
  for (int s = 0; s < 3; ++s) {
    p.g_blk_off[s] = s < nsrc ? g_blk_off[s] : 0;
    p.x_col_off[s] = s < nsrc ? x_col_off[s] : 0;
    if (s < nsrc && ((long long)(p.g_blk_off[s] + p.ksteps) * p.njb * 4096 > g_elems_total || p.x_col_off[s] + Dp > x_cols))
      return B2_ENOMEM;
  }
  if (g_elems_total < per) return B2_ENOMEM;
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[current_device() & 63];
  if (!attr_done) {
    if (cudaFuncSetAttribute(gt_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GG_SMEM) != cudaSuccess) return B2_ECUDA;
    attr_done = true;
  }
  CUtensorMap tmG, tmX;
  int rc;
  if ((rc = make_tmap_bf16_rows64(&tmG, G, (uint64_t)(g_elems_total / 64), 128))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmX, X, Nx, x_cols, ldx, 64))) return rc;
  const long long total = (long long)p.tiles * p.ksteps * nsrc;
  const int clusters = sm_count() / 2;
  const int grid = 2 * (int)(total < clusters ? total : clusters);
  // the accumulator of a CTA is [128 x 512] fp32 (all of TMEM): wider outputs take a second sweep over G for the rest
  for (int off = 0; off < Dp; off += 512) {
    p.d_off = off;
    p.Dp = Dp - off < 512 ? Dp - off : 512;
    if (off >= D) break;
    gt_gemm_kernel<<<grid, GG_THREADS, GG_SMEM, stream>>>(tmG, tmX, p);
  }
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int gt_gemm(const void* G, long long g_elems, int Nx, int Ny, const void* X, int ldx, int Dp, int D, const float* dyn,
            float gnorm, float* dY, int ldd, cudaStream_t stream) {
  const int zero = 0;
  return gt_gemm_multi(G, g_elems, 1, &zero, Nx, Ny, X, ldx, Dp, &zero, Dp, D, dyn, gnorm, dY, ldd, stream);
}

}  // namespace b2host
