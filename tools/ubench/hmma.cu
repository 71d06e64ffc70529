// mma.sync m16n8k16 bf16 throughput / latency on sm_100a (legacy tensor path)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int CH>
__global__ void k(float* out, int iters, unsigned long long* cyc) {
  float c[CH][4] = {};
  uint32_t a[4] = {0x3f803f80u + threadIdx.x, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u};
  unsigned long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CH; ++j) mma(c[j], a, 0x3f803f80u, 0x3f803f80u + i);
  }
  unsigned long long t1 = clock64();
  float s = 0; for (int j = 0; j < CH; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH> void run(int warps) {
  float* out; unsigned long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  k<CH><<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  k<CH><<<148, warps * 32>>>(out, iters, cyc); cudaDeviceSynchronize();
  unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (auto v : h) c += v; c /= 148;
  printf("warps/SM %2d  independent chains %d: %.2f cyc per MMA per warp, %.1f MMA/cyc/SM... = %.0f MAC/cyc/SM\n", warps, CH,
         c / (iters * CH), warps * iters * CH / c, 2048.0 * warps * iters * CH / c);
}
int main() { run<1>(1); run<4>(1); run<8>(1); run<1>(8); run<4>(8); run<8>(8); run<8>(16); return 0; }
