// Microbenchmarks that size the logits-backward redesign (DESIGN.md "measured machine limits"):
//   tma   : L2 -> shared streaming rate of [128 x 64] bf16 SWIZZLE_128B boxes, all CTAs sweeping the same 32 MB panel
//   mma   : tcgen05.mma issue rate for SS (A,B in smem) and TS (A in TMEM) at N = 64 / 128 / 256, operands resident
//   both  : the MMA loop with the TMA stream running in the same CTA
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I deepcoro_clip_b200/csrc tools/ubench/ubench.cu \
//        -L deepcoro_clip_b200 -lb200clip -Xlinker -rpath,'$ORIGIN/../../deepcoro_clip_b200' -o tools/ubench/ubench
#include "common.cuh"
#include <vector>
#include <cstdlib>
using namespace b2;

constexpr int CH = 128 * 64 * 2;   // 16 KB chunk
constexpr int SLOTS = 8;
constexpr int OPER_BYTES = 64 * 1024;   // resident operand area: A 16 KB (128 x 64) + B up to 32 KB (256 x 64)

struct UB {
  int mode;        // bit0: tma stream, bit1: mma loop
  int n;           // MMA N
  int ts;          // A from TMEM
  int iters;       // MMA k-chunks (each = 4 instructions of K=16)
  int tma_chunks;  // chunks per CTA
  int rows;        // rows of the panel
  int stagger;     // CTA-dependent start offset
  int nacc;        // accumulators rotated between consecutive MMA instructions (1 = dependent chain)
  int kmajor_b;    // 0: B K-major, 1: B MN-major
  int nslots;      // ring slots actually used by the tma test (<= SLOTS); 0 = SLOTS
  unsigned long long* cyc;   // [grid][2]
};

__global__ void __launch_bounds__(128, 1) ub_kernel(const __grid_constant__ CUtensorMap tm, UB p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* oper = smem;                              // 64 KB
  uint8_t* ring = smem + OPER_BYTES;                 // SLOTS * 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + SLOTS * CH);
  uint64_t* full = bars; uint64_t* empty = bars + SLOTS; uint64_t* done = bars + 2 * SLOTS;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  // pseudo-random bf16 operands in (-1, 1)
  for (int i = threadIdx.x; i < OPER_BYTES / 2; i += blockDim.x) {
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    float f = ((h >> 8) & 0xffff) / 32768.f - 1.f;
    reinterpret_cast<__nv_bfloat16*>(oper)[i] = __float2bfloat16(f);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm);
  }
  if (warp == 2) { tmem_alloc(tslot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tslot;
  const int row_tiles = p.rows / 128;
  const int NSL = p.nslots > 0 ? p.nslots : SLOTS;
  if (warp == 0 && lane == 0 && (p.mode & 1)) {
    int slot = 0; uint32_t ph = 0;
    int t = p.stagger ? (blockIdx.x * 7) % row_tiles : 0, kc = 0;
    for (int c = 0; c < p.tma_chunks; ++c) {
      mbar_wait(&empty[slot], ph ^ 1);
      mbar_expect_tx(&full[slot], CH);
      tma_load_2d(ring + slot * CH, &tm, &full[slot], kc * 64, t * 128);
      if (++kc == 8) { kc = 0; if (++t == row_tiles) t = 0; }
      if (++slot == NSL) { slot = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0 && (p.mode & 1)) {
    unsigned long long t0 = clock64();
    int slot = 0; uint32_t ph = 0;
    for (int c = 0; c < p.tma_chunks; ++c) {
      mbar_wait(&full[slot], ph);
      mbar_arrive(&empty[slot]);
      if (++slot == NSL) { slot = 0; ph ^= 1; }
    }
    p.cyc[blockIdx.x * 2 + 0] = clock64() - t0;
  } else if (warp == 3 && (p.mode & 2)) {
    const uint32_t idesc = make_idesc_bf16(128, p.n, 0, p.kmajor_b);
    const uint32_t sa = smem_u32(oper);
    const uint64_t adesc = make_smem_desc_sw128(sa, 1024);
    const uint64_t bdesc = make_smem_desc_sw128(sa + 16384, 1024);
    unsigned long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t d = tbase + (p.nacc > 1 ? ((it * 4 + k) & (p.nacc - 1)) * p.n : 0);
        const uint64_t bd = p.kmajor_b ? make_smem_desc_sw128(sa + 16384 + k * 2048, 8192) : bdesc + 2 * k;
        if (elect_one()) {
          if (p.ts) mma_ts(d, tbase + 448 + k * 8, bd, idesc, 1u);
          else mma_ss(d, adesc + 2 * k, bd, idesc, 1u);
        }
        __syncwarp();
      }
    }
    if (elect_one()) tc_commit(done);
    __syncwarp();
    mbar_wait(done, 0);
    if (lane == 0) p.cyc[blockIdx.x * 2 + 1] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}


template <int N, int TS>
__global__ void __launch_bounds__(128, 1) tight_kernel(unsigned long long* cyc, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + OPER_BYTES);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  for (int i = threadIdx.x; i < OPER_BYTES / 2; i += blockDim.x) {
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16(((h >> 8) & 0xffff) / 32768.f - 1.f);
  }
  if (threadIdx.x == 0) { mbar_init(done, 1); fence_mbar_init(); }
  if (warp == 2) { tmem_alloc(tslot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tslot;
  if (warp == 3) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t sa = smem_u32(smem);
    const uint64_t adesc = make_smem_desc_sw128(sa, 1024);
    const uint64_t bdesc = make_smem_desc_sw128(sa + 16384, 1024);
    if (elect_one()) {
      unsigned long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (TS) mma_ts(tbase, tbase + 448 + k * 8, bdesc + 2 * k, idesc, 1u);
          else mma_ss(tbase, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
        }
      }
      tc_commit(done);
      mbar_wait(done, 0);
      cyc[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tbase, 512); }
}

template <int N, int TS>
void run_tight(unsigned long long* cyc, int sms) {
  const int smem = OPER_BYTES + 2048;
  cudaFuncSetAttribute(tight_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 8192 * 64 / N;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  tight_kernel<N, TS><<<sms, 128, smem>>>(cyc, iters);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  tight_kernel<N, TS><<<sms, 128, smem>>>(cyc, iters);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<unsigned long long> h(sms); cudaMemcpy(h.data(), cyc, sms * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (auto v : h) c += v; c /= sms;
  printf("tight %s N=%-3d  %7.3f ms  %6.1f cyc per K16 instr  %7.1f TFLOP/s chip  err=%d\n", TS ? "TS" : "SS", N, ms,
         c / (iters * 4.0), 2.0 * 128 * N * 16 * 4.0 * iters * sms / ms / 1e9, (int)cudaGetLastError());
}

int main(int argc, char** argv) {
  const int rows = 32768, cols = 512;
  __nv_bfloat16* panel; cudaMalloc(&panel, (size_t)rows * cols * 2);
  cudaMemset(panel, 0x3c, (size_t)rows * cols * 2);
  CUtensorMap tm;
  if (b2host::make_tmap_bf16_2d(&tm, panel, rows, cols, cols, 128)) { printf("tmap failed\n"); return 1; }
  int sms = b2host::sm_count();
  unsigned long long* cyc; cudaMalloc(&cyc, sms * 16);
  const int smem = OPER_BYTES + SLOTS * CH + 1024 + 256;
  cudaFuncSetAttribute(ub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int grid = sms;
  auto run = [&](const char* name, UB p) {
    p.cyc = cyc; p.rows = rows;
    cudaMemset(cyc, 0, sms * 16);
    ub_kernel<<<grid, 128, smem>>>(tm, p);   // warm
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    ub_kernel<<<grid, 128, smem>>>(tm, p);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaError_t err = cudaGetLastError();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned long long> h(sms * 2); cudaMemcpy(h.data(), cyc, sms * 16, cudaMemcpyDeviceToHost);
    double c0 = 0, c1 = 0; for (int i = 0; i < grid; ++i) { c0 += h[2 * i]; c1 += h[2 * i + 1]; } c0 /= grid; c1 /= grid;
    printf("%-34s %8.3f ms  err=%d", name, ms, (int)err);
    if (p.mode & 1) printf("  tma: %7.1f GB/s chip, %6.2f B/cyc/SM (avg %0.f cyc)", (double)p.tma_chunks * CH * grid / ms / 1e6,
                           (double)p.tma_chunks * CH / c0, c0);
    if (p.mode & 2) printf("  mma: %6.1f cyc per K16 instr, %7.1f TFLOP/s chip", c1 / (p.iters * 4.0),
                           2.0 * 128 * p.n * 16 * 4.0 * p.iters * sms / ms / 1e9);
    printf("\n");
  };
  run_tight<64,0>(cyc,sms); run_tight<128,0>(cyc,sms); run_tight<256,0>(cyc,sms);
  run_tight<64,1>(cyc,sms); run_tight<128,1>(cyc,sms); run_tight<256,1>(cyc,sms);
  if (argc > 1 && argv[1][0]=='t') return 0;
  UB p{}; 
  if (argc > 1 && argv[1][0]=='l') {     // latency: bytes in flight vs achieved rate (Little's law), all 148 CTAs streaming
    for (int ns : {1, 2, 3, 4, 6, 8}) { char nm[64]; snprintf(nm, 64, "tma slots in flight=%d", ns); UB q{}; q.mode = 1; q.tma_chunks = 6000; q.stagger = 1; q.nslots = ns; run(nm, q); }
    grid = 8;
    for (int ns : {1, 2, 4}) { char nm[64]; snprintf(nm, 64, "tma slots=%d, 8 CTAs only", ns); UB q{}; q.mode = 1; q.tma_chunks = 6000; q.stagger = 1; q.nslots = ns; run(nm, q); }
    return 0;
  }
  if (argc > 1 && argv[1][0]=='g') {
    for (int g : {1, 8, 37, 74, 111, 148}) { grid = g; char nm[64]; snprintf(nm, 64, "tma grid=%d", g); UB q{}; q.mode = 1; q.tma_chunks = 20000; q.stagger = 1; run(nm, q); }
    return 0;
  }
  p.mode = 1; p.tma_chunks = 20000; p.stagger = 0; run("tma same-order", p);
  p.stagger = 1; run("tma staggered", p);
  for (int ts = 0; ts < 2; ++ts)
    for (int n : {64, 128, 256}) {
      char nm[64]; snprintf(nm, 64, "mma %s N=%d", ts ? "TS" : "SS", n);
      UB q{}; q.mode = 2; q.n = n; q.ts = ts; q.iters = 8192 * 64 / n; run(nm, q);
    }
  for (int ts = 0; ts < 2; ++ts)
    for (int n : {64, 128}) {
      char nm[64]; snprintf(nm, 64, "mma %s N=%d 2 accumulators", ts ? "TS" : "SS", n);
      UB q{}; q.mode = 2; q.n = n; q.ts = ts; q.nacc = 2; q.iters = 8192 * 64 / n; run(nm, q);
    }
  for (int ts = 0; ts < 2; ++ts)
    for (int n : {96, 160, 192, 224}) {
      char nm[64]; snprintf(nm, 64, "mma %s N=%d", ts ? "TS" : "SS", n);
      UB q{}; q.mode = 2; q.n = n; q.ts = ts; q.iters = 8192 * 64 / n; run(nm, q);
    }
  for (int n : {128, 256}) {
      char nm[64]; snprintf(nm, 64, "mma TS N=%d B MN-major", n);
      UB q{}; q.mode = 2; q.n = n; q.ts = 1; q.kmajor_b = 1; q.iters = 8192 * 64 / n; run(nm, q);
  }
  for (int ts = 0; ts < 2; ++ts)
    for (int n : {64, 128, 256}) {
      char nm[64]; snprintf(nm, 64, "mma %s N=%d + tma", ts ? "TS" : "SS", n);
      UB q{}; q.mode = 3; q.n = n; q.ts = ts; q.iters = 8192 * 64 / n; q.tma_chunks = 20000; q.stagger = 1; run(nm, q);
    }
  return 0;
}
