// Minimal host emulation of the CUDA execution model for the library's plain CUDA-core device code (no inline PTX, no
// tensor cores): every thread of a block is a std::thread, warps exchange through a 32-party barrier (shuffles, ballots,
// __syncwarp), __syncthreads is a block barrier, blocks run one after the other (so `__shared__` can be a static), atomics
// take a global lock. Test infrastructure only (tests/emul/): it lets the CPU suite execute the very device functions
// that ship in libb200clip.so — slow, but bit-faithful for integer / compare / select logic and IEEE for the arithmetic.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

namespace emul {
struct Dim { unsigned x = 1, y = 1, z = 1; };
struct Warp {
  std::barrier<> bar{32};
  uint64_t slot[32];
  uint32_t frag[32][8];      // mma / ldmatrix operand exchange (pool_mma_prims_emul.h)
};
struct Block {
  explicit Block(unsigned n, size_t smem_bytes = 0)
      : bar((std::ptrdiff_t)n), warps((n + 31) / 32), smem_store(smem_bytes + 2048) {
    dyn_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_store.data()) + 1023) & ~uintptr_t(1023));
  }
  std::barrier<> bar;
  std::vector<Warp> warps;
  std::vector<unsigned char> smem_store;
  unsigned char* dyn_smem;   // 1024-byte aligned dynamic shared memory of the running block
};
inline thread_local Warp* tl_warp = nullptr;
inline thread_local Block* tl_block = nullptr;
inline thread_local int tl_lane = 0;
inline std::mutex g_atomic_lock;
}  // namespace emul

inline thread_local emul::Dim threadIdx, blockIdx;
inline emul::Dim blockDim, gridDim;

template <class T>
inline T __shfl_xor_sync(unsigned, T v, int o) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  emul::tl_warp->slot[emul::tl_lane] = raw;
  emul::tl_warp->bar.arrive_and_wait();
  raw = emul::tl_warp->slot[(emul::tl_lane ^ o) & 31];
  emul::tl_warp->bar.arrive_and_wait();
  T r;
  std::memcpy(&r, &raw, sizeof(T));
  return r;
}
inline unsigned __ballot_sync(unsigned, bool pred) {
  emul::tl_warp->slot[emul::tl_lane] = pred ? 1 : 0;
  emul::tl_warp->bar.arrive_and_wait();
  unsigned b = 0;
  for (int l = 0; l < 32; ++l) b |= emul::tl_warp->slot[l] ? (1u << l) : 0u;
  emul::tl_warp->bar.arrive_and_wait();
  return b;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emul::tl_warp->bar.arrive_and_wait(); }
inline void __syncthreads() { emul::tl_block->bar.arrive_and_wait(); }
inline void __threadfence() {}
template <class T>
inline T __ldg(const T* p) { return *p; }

template <class T>
inline T atomicAdd(T* p, T v) {
  std::lock_guard<std::mutex> g(emul::g_atomic_lock);
  const T old = *p;
  *p = old + v;
  return old;
}
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __expf(float x) { return expf(x); }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float __logf(float x) { return logf(x); }
struct float2 { float x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

namespace emul {
// Runs kernel body `f` (a callable without arguments) for a 1-D grid of 1-D blocks. Blocks are sequential; a thread that
// returns leaves its barriers (so early-exiting warps do not block the rest of the block).
template <class F>
inline void launch(Dim grid, unsigned block, F f, size_t smem_bytes = 0) {
  gridDim = grid;
  blockDim = Dim{block, 1, 1};
  for (unsigned bz = 0; bz < grid.z; ++bz)
  for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned b = 0; b < grid.x; ++b) {
      Block bs(block, smem_bytes);
      // a partially filled last warp would need a smaller warp barrier: the library only launches multiples of 32
      std::vector<std::thread> th;
      th.reserve(block);
      for (unsigned t = 0; t < block; ++t)
        th.emplace_back([&, t, b, by, bz] {
          threadIdx = Dim{t, 0, 0};
          blockIdx = Dim{b, by, bz};
          tl_block = &bs;
          tl_warp = &bs.warps[t / 32];
          tl_lane = (int)(t % 32);
          f();
          tl_warp->bar.arrive_and_drop();
          bs.bar.arrive_and_drop();
        });
      for (auto& x : th) x.join();
    }
}
template <class F>
inline void launch(unsigned grid, unsigned block, F f) { launch(Dim{grid, 1, 1}, block, f); }
}  // namespace emul
