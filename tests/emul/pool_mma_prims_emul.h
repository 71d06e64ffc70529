// Host emulation of deepcoro_clip_b200/csrc/attnpool_mma_prims.cuh: the same names with the semantics the PTX ISA gives
// the instructions they wrap, so that attnpool_mma_kernels.cuh compiles unchanged for the CPU.
//   * 16-bit conversions: bf16 round-to-nearest-even by hand, fp16 through _Float16;
//   * mma.sync m16n8k16 (row.col, fp32 accumulate) and ldmatrix.x4(.trans): warp-collective — every lane deposits its
//     operands in the warp's exchange area, the warp meets at a barrier, every lane gathers what the fragment layout of the
//     PTX ISA assigns to it (A: a0..a3 = (g, 2t) (g+8, 2t) (g, 2t+8) (g+8, 2t+8); B: b0 = (k 2t, n g), b1 = (k 2t+8, n g);
//     C: c0,c1 = (g, 2t | 2t+1), c2,c3 = (g+8, ..); ldmatrix: lane l supplies row l%8 of matrix l/8, receives
//     (row g, elements 2t, 2t+1), transposed: (rows 2t, 2t+1, element g)), g = lane / 4, t = lane % 4;
//   * shared memory: the block's dynamic buffer, "shared-window addresses" are byte offsets into it;
//   * TMA tile load: synchronous copy of D/64 boxes [32 rows x 64 columns] with SWIZZLE_128B placement and zero fill
//     past the end of the tensor, completing the mbarrier (a phase counter) on the spot.
#pragma once
#include "cuda_emul.h"
#include <algorithm>
#include <atomic>

#define B2_DYN_SMEM(name) unsigned char* name = emul::tl_block->dyn_smem
#define __grid_constant__
using std::min;

#define B2_DYN_SMEM16(name) unsigned char* name = emul::tl_block->dyn_smem
#define B2_DYN_SMEM_F32(name) float* name = reinterpret_cast<float*>(emul::tl_block->dyn_smem)
struct uint4 { uint32_t x, y, z, w; };
struct __nv_bfloat16 { uint16_t x; };
struct __half { uint16_t x; };
// CUDA conversion intrinsics and the cp.async pipeline primitives used by the CUDA-core kernels (attnpool_kernels.cuh)
inline float __bfloat162float(__nv_bfloat16 v) { return __uint_as_float((uint32_t)v.x << 16); }
inline __nv_bfloat16 __float2bfloat16_rn(float v) {
  uint32_t u = __float_as_uint(v);
  if ((u & 0x7fffffffu) > 0x7f800000u) return __nv_bfloat16{(uint16_t)((u >> 16) | 0x40)};
  u += 0x7fffu + ((u >> 16) & 1u);
  return __nv_bfloat16{(uint16_t)(u >> 16)};
}
inline float __half2float(__half v) { _Float16 h; std::memcpy(&h, &v.x, 2); return (float)h; }
inline __half __float2half_rn(float v) { _Float16 h = (_Float16)v; __half r; std::memcpy(&r.x, &h, 2); return r; }
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
inline float2 __bfloat1622float2(__nv_bfloat162 v) { return float2{__bfloat162float(v.x), __bfloat162float(v.y)}; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0u; }
inline int atomicExch(int* p, int v) { std::lock_guard<std::mutex> g(emul::g_atomic_lock); const int o = *p; *p = v; return o; }
inline void __pipeline_memcpy_async(void* dst, const void* src, size_t n) { std::memcpy(dst, src, n); }
inline void __pipeline_commit() {}
inline void __pipeline_wait_prior(int) {}
struct CUtensorMap {          // what make_tmap_bf16_2d encodes: [rows, cols] 16-bit elements, row pitch in elements
  const void* base;
  uint64_t rows, cols;
  long long pitch;
};

namespace b2 {

constexpr int PM_TT = 32;

template <typename T> struct PmT;
namespace emu {
inline uint16_t f2bf(float v) {
  uint32_t u = __float_as_uint(v);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline float bf2f(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }
inline uint16_t f2h(float v) { _Float16 h = (_Float16)v; uint16_t b; std::memcpy(&b, &h, 2); return b; }
inline float h2f(uint16_t b) { _Float16 h; std::memcpy(&h, &b, 2); return (float)h; }
inline unsigned char* smem() { return emul::tl_block->dyn_smem; }

// D += A * B for one warp; cvt = 16-bit pattern -> float
template <class Cvt>
inline void mma_m16n8k16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, Cvt cvt) {
  emul::Warp& w = *emul::tl_warp;
  const int lane = emul::tl_lane, g = lane >> 2, t = lane & 3;
  for (int i = 0; i < 4; ++i) w.frag[lane][i] = a[i];
  w.frag[lane][4] = b0;
  w.frag[lane][5] = b1;
  w.bar.arrive_and_wait();
  auto A = [&](int r, int k) {
    const uint32_t v = w.frag[(r & 7) * 4 + ((k & 7) >> 1)][(r >> 3) + 2 * (k >> 3)];
    return cvt((uint16_t)(k & 1 ? v >> 16 : v & 0xffffu));
  };
  auto B = [&](int k, int n) {
    const uint32_t v = w.frag[n * 4 + ((k & 7) >> 1)][4 + (k >> 3)];
    return cvt((uint16_t)(k & 1 ? v >> 16 : v & 0xffffu));
  };
  float d[4] = {c[0], c[1], c[2], c[3]};
  for (int e = 0; e < 4; ++e) {
    const int r = g + (e >> 1) * 8, n = 2 * t + (e & 1);
    double s = 0.0;                                  // exact products, one rounding into the fp32 accumulator
    for (int k = 0; k < 16; ++k) s += (double)A(r, k) * (double)B(k, n);
    d[e] = (float)((double)d[e] + s);
  }
  w.bar.arrive_and_wait();
  for (int e = 0; e < 4; ++e) c[e] = d[e];
}
}  // namespace emu

template <> struct PmT<__nv_bfloat16> {
  static inline uint16_t bits(float v) { return emu::f2bf(v); }
  static inline float val(uint16_t b) { return emu::bf2f(b); }
  static inline void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    emu::mma_m16n8k16(c, a, b0, b1, emu::bf2f);
  }
};
template <> struct PmT<__half> {
  static inline uint16_t bits(float v) { return emu::f2h(v); }
  static inline float val(uint16_t b) { return emu::h2f(b); }
  static inline void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    emu::mma_m16n8k16(c, a, b0, b1, emu::h2f);
  }
};

inline uint32_t smem_u32(const void* p) { return (uint32_t)((const unsigned char*)p - emu::smem()); }
inline uint32_t lds32(uint32_t addr) { uint32_t v; std::memcpy(&v, emu::smem() + addr, 4); return v; }
inline void sts32(uint32_t addr, uint32_t v) { std::memcpy(emu::smem() + addr, &v, 4); }
inline uint4 lds128v(uint32_t addr) { uint4 v; std::memcpy(&v, emu::smem() + addr, 16); return v; }

inline void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  emul::Warp& w = *emul::tl_warp;
  const int lane = emul::tl_lane, g = lane >> 2, t = lane & 3;
  w.frag[lane][6] = addr;
  w.bar.arrive_and_wait();
  for (int i = 0; i < 4; ++i) r[i] = lds32(w.frag[i * 8 + g][6] + 4 * t);
  w.bar.arrive_and_wait();
}
inline void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  emul::Warp& w = *emul::tl_warp;
  const int lane = emul::tl_lane, g = lane >> 2, t = lane & 3;
  w.frag[lane][6] = addr;
  w.bar.arrive_and_wait();
  for (int i = 0; i < 4; ++i) {
    uint16_t lo, hi;
    std::memcpy(&lo, emu::smem() + w.frag[i * 8 + 2 * t][6] + 2 * g, 2);
    std::memcpy(&hi, emu::smem() + w.frag[i * 8 + 2 * t + 1][6] + 2 * g, 2);
    r[i] = (uint32_t)lo | ((uint32_t)hi << 16);
  }
  w.bar.arrive_and_wait();
}

// mbarrier as a completed-phase counter (one arrival per phase in these kernels: the TMA of one tile)
inline void mbar_init(uint64_t* bar, uint32_t) { reinterpret_cast<std::atomic<uint64_t>*>(bar)->store(0); }
inline void fence_mbar_init() {}
inline void fence_proxy_async_smem() {}
inline void tma_prefetch_desc(const CUtensorMap*) {}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  while ((reinterpret_cast<std::atomic<uint64_t>*>(bar)->load() & 1u) == parity) std::this_thread::yield();
}
inline void tma_x_tile(uint32_t tile, const CUtensorMap* tm, uint64_t* bar, int grow, int D) {
  const uint16_t* base = static_cast<const uint16_t*>(tm->base);
  for (int cb = 0; cb < D / 64; ++cb)
    for (int r = 0; r < PM_TT; ++r)
      for (int ch = 0; ch < 8; ++ch) {
        unsigned char* dst = emu::smem() + tile + cb * (PM_TT * 128) + r * 128 + ((ch ^ (r & 7)) << 4);
        const long long row = (long long)grow + r;
        if (row < (long long)tm->rows) std::memcpy(dst, base + row * tm->pitch + cb * 64 + ch * 8, 16);
        else std::memset(dst, 0, 16);
      }
  reinterpret_cast<std::atomic<uint64_t>*>(bar)->fetch_add(1);
}

inline float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
inline float warp_max(float v) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// common.cuh: counter-based attention-dropout mask (only reached with drop_p > 0)
inline bool attn_keep(unsigned long long seed, int row, int n, float p) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)(unsigned)row * 0x100000001B3ull + (unsigned)n + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f) >= p;
}

}  // namespace b2
