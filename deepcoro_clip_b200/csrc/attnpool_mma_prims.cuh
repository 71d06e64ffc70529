// Device primitives of the attention-pool MMA kernels (attnpool_mma_kernels.cuh): 16-bit conversions, mma.sync m16n8k16,
// ldmatrix, 32-bit shared-window loads / stores and the TMA tile load. Everything that is inline PTX or a CUDA intrinsic
// type lives here, so that tests/emul/ can substitute a host implementation of exactly this interface and compile the
// kernels file unchanged.
#pragma once
#include "common.cuh"

#define B2_DYN_SMEM(name) extern __shared__ unsigned char name[]

namespace b2 {

constexpr int PM_TT = 32;          // tokens per tile

template <typename T> struct PmT;
template <> struct PmT<__nv_bfloat16> {
  static __device__ __forceinline__ uint16_t bits(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
  static __device__ __forceinline__ float val(uint16_t b) { return __bfloat162float(__ushort_as_bfloat16(b)); }
  static __device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};
template <> struct PmT<__half> {
  static __device__ __forceinline__ uint16_t bits(float v) { return __half_as_ushort(__float2half_rn(v)); }
  static __device__ __forceinline__ float val(uint16_t b) { return __half2float(__ushort_as_half(b)); }
  static __device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds128v(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// One elected thread: D/64 box loads of the tile starting at global row `grow` (rows of the [B*N, D] matrix); rows past
// the end of the tensor are zero-filled by TMA, rows past the end of this CTA's token range are real neighbouring
// tokens and are neutralised by zero weights.
__device__ __forceinline__ void tma_x_tile(uint32_t tile, const CUtensorMap* tm, uint64_t* bar, int grow, int D) {
  mbar_expect_tx(bar, PM_TT * D * 2);
  for (int cb = 0; cb < D / 64; ++cb)
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(tile + cb * (PM_TT * 128)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(cb * 64), "r"(grow)
        : "memory");
}

}  // namespace b2
