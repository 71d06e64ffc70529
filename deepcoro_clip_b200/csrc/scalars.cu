// Small device-side scalar plumbing so the loss never forces a host sync:
//   dyn_prep      : log_temp (device) -> tau, 1/tau, log2e/tau, exponent shift, clamp flag, bias
//   lse_finalize  : sums -> log-sum-exp total (double) + per-row gradient scales c / sum
//   diag_sum      : sum_i f(a_i . b_i) (double) for the target-logit term
// Layout of the `dyn` float[16] block (read by logits_fwd / logits_bwd when their dyn pointer is non-null):
//   [0] scale2 = log2(e)/tau   [1] shift2   [2] 1/tau   [3] tau   [4] 1 if tau was clamped (dlog_temp = 0)
//   [5] bias                    [6] ln(2)*shift2        [7] 1 if (learnable) temperature gradient is live
//   [8] SigLIP logit clamp (30) [9] SigLIP target of non-positive pairs (label smoothing eps/2, default 0)
//   [10] entropy-regulariser gradient coefficient (written by siglip_entropy_coef, default 0)
#include "common.cuh"
#include "host_api.h"
#include "alignment_diag.cuh"

#include "scalars_kernels.cuh"

namespace b2host {
using namespace b2;

int dyn_prep(const float* log_temp, const float* bias, float clamp_min, float bound, float* dyn, cudaStream_t s) {
  dyn_prep_kernel<<<1, 32, 0, s>>>(log_temp, bias, clamp_min, bound, dyn);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int dyn_set_siglip(float* dyn, float lclamp, float yneg, cudaStream_t s) {
  dyn_set_siglip_kernel<<<1, 32, 0, s>>>(dyn, lclamp, yneg);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int dyn_set_stable(float* dyn, int stable, cudaStream_t s) {
  dyn_set_stable_kernel<<<1, 32, 0, s>>>(dyn, stable);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int lse_finalize(const float* sums, int n, const float* dyn, float c, float* scale_out, double* acc, cudaStream_t s) {
  if (n <= 0) return B2_EINVAL;
  lse_finalize_kernel<<<(n + 255) / 256, 256, 0, s>>>(sums, n, dyn, c, scale_out, acc);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int vec_fsum(const float* v, int n, int gated, double* acc, cudaStream_t s) {
  if (n <= 0) return B2_EINVAL;
  int blocks = (n + 255) / 256;
  if (blocks > 64) blocks = 64;
  vec_fsum_kernel<<<blocks, 256, 0, s>>>(v, n, gated, acc);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int clip_finalize(const float* sums, int n, int nvec, const float* dyn, float eps, int gated, const double* unif,
                  float* rowscale, float* colscale, float* loss_out, double* acc_out, cudaStream_t s) {
  if (n <= 0 || !sums || !dyn || !rowscale || !colscale || !loss_out || (nvec != 3 && nvec != 7)) return B2_EINVAL;
  int blocks = (n + 1023) / 1024;
  if (blocks > FIN_MAX_BLOCKS) blocks = FIN_MAX_BLOCKS;
  FinPeers ps{};
  ps.ptr[0] = sums;
  ps.world = 0;
  ps.rows_per_rank = n;
  clip_finalize_kernel<<<blocks, 1024, 0, s>>>(ps, n, nvec, dyn, eps, gated, unif, rowscale, colscale, loss_out, acc_out);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int clip_finalize_peers(const float* const* peer_sums_host, int world, int n, int nvec, const float* dyn, float eps, int gated,
                        const double* unif, float* rowscale, float* colscale, float* loss_out, double* acc_out,
                        cudaStream_t s) {
  if (n <= 0 || !peer_sums_host || world < 1 || world > FIN_MAX_PEERS || n % world || !dyn || !rowscale || !colscale ||
      !loss_out || (nvec != 3 && nvec != 7))
    return B2_EINVAL;
  int blocks = (n + 1023) / 1024;
  if (blocks > FIN_MAX_BLOCKS) blocks = FIN_MAX_BLOCKS;
  FinPeers ps{};
  for (int r = 0; r < world; ++r) {
    if (!peer_sums_host[r]) return B2_EINVAL;
    ps.ptr[r] = peer_sums_host[r];
  }
  ps.world = world;
  ps.rows_per_rank = n / world;
  clip_finalize_kernel<<<blocks, 1024, 0, s>>>(ps, n, nvec, dyn, eps, gated, unif, rowscale, colscale, loss_out, acc_out);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host

namespace b2 {

// Cross-rank barrier over symmetric-memory flags (symm.py): flags[rank] = int32 [4 channels][8 sources] arrival epochs followed
// by 4 local epoch counters at +32. Every rank enqueues the same sequence of barriers per channel, so the device-side epoch
// counters agree without any host state (CUDA-graph safe). One CTA: thread p writes this rank's new epoch into peer p's flag
// (release, system scope) and spins until peer p's epoch has arrived here (acquire). Replaces the framework's barrier op:
// one ctypes call instead of a dispatcher round trip (~10 us instead of 40-75 us of host time per barrier).
struct SymmFlags { unsigned* ptr[8]; };
__global__ void symm_barrier_kernel(SymmFlags f, int world, int rank, int ch) {
  __shared__ unsigned e_sh;
  unsigned* mine = f.ptr[rank];
  if (threadIdx.x == 0) e_sh = atomicAdd(mine + 32 + ch, 1u) + 1u;
  __syncthreads();
  const unsigned e = e_sh;
  if ((int)threadIdx.x < world) {
    __threadfence_system();
    unsigned* dst = f.ptr[threadIdx.x] + ch * 8 + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(e) : "memory");
    const unsigned* src = mine + ch * 8 + threadIdx.x;
    unsigned v;
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (true) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
      if ((int)(v - e) >= 0) break;
      if ((++spins & 0xfff) == 0) {
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 5000000000ull) {
          printf("b200clip: symmetric-memory barrier timeout (rank %d waits for rank %d, channel %d, epoch %u, seen %u)\n", rank,
                 (int)threadIdx.x, ch, e, v);
          __trap();
        }
      }
    }
  }
}

// d loss / d log_temp with the partial sums of sum G L read from every rank's symmetric block (after a barrier)
struct ScalPeers { const double* ptr[8]; };
__global__ void clip_dlogtemp_peers_kernel(ScalPeers ps, int world, const float* __restrict__ dyn, const float* __restrict__ gmul,
                                           const double* __restrict__ unif, int n, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double sgl = 0.0;
  for (int r = 0; r < world; ++r) sgl += ps.ptr[r][0];
  const double u = unif ? unif[0] : 0.0;
  out[0] = (float)((u / n - sgl * (double)dyn[2]) * (double)dyn[7] * (double)gmul[0]);
}

// Sum-all-reduce of one fp32 buffer that exists once per rank in symmetric memory (<= 8 ranks of one NVLink domain): rank r
// reduces slice r (n / W elements, 16-byte vectors) over the W peers' copies and writes the result back into every copy.
// Per rank 2 (W - 1) / W of the buffer crosses NVLink in each direction. The caller brackets it with symm_barrier calls
// (all partial sums complete before, all slices broadcast after). No NCCL call, i.e. ~10 us of host time instead of the
// 150-350 us an eager dist.all_reduce costs when the per-rank problem is small.
struct SymmBufs { float* ptr[8]; };
__global__ void __launch_bounds__(256) symm_allreduce_f32_kernel(SymmBufs b, long long n4, int world, int rank) {
  // n4 = number of float4 elements; slice of this rank = [rank * per, min(n4, (rank + 1) * per))
  const long long per = (n4 + world - 1) / world;
  const long long lo = rank * per, hi = min(n4, lo + per);
  for (long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
    float4 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < world) v[r] = reinterpret_cast<const float4*>(b.ptr[r])[i];
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < 8; ++r)
      if (r < world) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < world) reinterpret_cast<float4*>(b.ptr[r])[i] = s;
  }
}
// out[i] = sum over ranks of peers[r][i] (fp64, n <= 32): scalar tails (summed in rank order on every rank: identical results)
struct SymmF64 { const double* ptr[8]; };
__global__ void symm_sum_f64_kernel(SymmF64 b, int n, int world, double* __restrict__ out) {
  const int i = threadIdx.x;
  if (i < n) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += b.ptr[r][i];
    out[i] = s;
  }
}

}  // namespace b2

namespace b2host {
using namespace b2;

int symm_allreduce_f32(void* const* bufs_host, long long n, int world, int rank, cudaStream_t s) {
  if (!bufs_host || world < 1 || world > 8 || rank < 0 || rank >= world || n <= 0 || (n & 3)) return B2_EINVAL;
  SymmBufs b{};
  for (int r = 0; r < world; ++r) {
    if (!bufs_host[r] || (reinterpret_cast<uintptr_t>(bufs_host[r]) & 15)) return B2_EINVAL;
    b.ptr[r] = reinterpret_cast<float*>(bufs_host[r]);
  }
  const long long n4 = n / 4, per = (n4 + world - 1) / world;
  long long blocks = (per + 255) / 256;
  const long long cap = 2LL * sm_count();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  symm_allreduce_f32_kernel<<<(unsigned)blocks, 256, 0, s>>>(b, n4, world, rank);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int symm_sum_f64(const void* const* peers_host, int n, int world, double* out, cudaStream_t s) {
  if (!peers_host || !out || n < 1 || n > 32 || world < 1 || world > 8) return B2_EINVAL;
  SymmF64 b{};
  for (int r = 0; r < world; ++r) {
    if (!peers_host[r]) return B2_EINVAL;
    b.ptr[r] = reinterpret_cast<const double*>(peers_host[r]);
  }
  symm_sum_f64_kernel<<<1, 32, 0, s>>>(b, n, world, out);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int symm_barrier(void* const* flags_host, int world, int rank, int channel, cudaStream_t s) {
  if (!flags_host || world < 1 || world > 8 || rank < 0 || rank >= world || channel < 0 || channel > 3) return B2_EINVAL;
  SymmFlags f{};
  for (int r = 0; r < world; ++r) {
    if (!flags_host[r]) return B2_EINVAL;
    f.ptr[r] = reinterpret_cast<unsigned*>(flags_host[r]);
  }
  symm_barrier_kernel<<<1, 32, 0, s>>>(f, world, rank, channel);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int clip_dlogtemp_peers(const void* const* scal_host, int world, const float* dyn, const float* gmul, const double* unif, int n,
                        float* out, cudaStream_t s) {
  if (!scal_host || world < 1 || world > 8 || !dyn || !gmul || !out || n <= 0) return B2_EINVAL;
  ScalPeers ps{};
  for (int r = 0; r < world; ++r) {
    if (!scal_host[r]) return B2_EINVAL;
    ps.ptr[r] = reinterpret_cast<const double*>(scal_host[r]);
  }
  clip_dlogtemp_peers_kernel<<<1, 32, 0, s>>>(ps, world, dyn, gmul, unif, n, out);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int clip_dlogtemp(const double* scal0, const float* dyn, const float* gmul, const double* unif, int n, float* out,
                  cudaStream_t s) {
  if (!scal0 || !dyn || !gmul || !out || n <= 0) return B2_EINVAL;
  clip_dlogtemp_kernel<<<1, 32, 0, s>>>(scal0, dyn, gmul, unif, n, out);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int alignment_diag(const float* sums, int n, const float* dyn, int gated, float* out, cudaStream_t s) {
  if (!sums || !dyn || !out || n <= 0) return B2_EINVAL;
  alignment_diag_kernel<<<1, 1024, 0, s>>>(sums, n, dyn, gated, out);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int diag_sum(const void* a, int lda, const void* b, int ldb, int rows, int K, int gated, float* dots, double* acc,
             cudaStream_t s) {
  if (rows <= 0 || K <= 0 || (K & 1)) return B2_EINVAL;
  diag_sum_kernel<<<(rows + 7) / 8, 256, 0, s>>>((const __nv_bfloat16*)a, lda, (const __nv_bfloat16*)b, ldb, rows, K,
                                                 gated, dots, acc);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
