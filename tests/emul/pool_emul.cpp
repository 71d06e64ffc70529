// Host build of the attention-pool MMA kernels (deepcoro_clip_b200/csrc/attnpool_mma_kernels.cuh, the file that ships in
// libb200clip.so) against the emulated primitives. Launch geometry and shared-memory sizing follow attnpool_mma.cu.
#include "pool_mma_prims_emul.h"
#include "../../deepcoro_clip_b200/csrc/attnpool_mma_kernels.cuh"
#include "../../deepcoro_clip_b200/csrc/attnpool_kernels.cuh"
#include "../../deepcoro_clip_b200/csrc/rope3d_kernels.cuh"
#include "../../deepcoro_clip_b200/csrc/querypool_kernels.cuh"
#include "../../deepcoro_clip_b200/csrc/multipos_kernels.cuh"
#include "../../deepcoro_clip_b200/csrc/dense_metrics_kernels.cuh"

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

// ---- the attention-pool entry points of include/b200clip.h with the library's own dispatch and launch geometry
//      (attnpool.cu / attnpool_mma.cu on a 148-SM device), so the package's Python host code can run on top of the
//      emulated kernels unchanged (tests/test_emulated_modules.py). `stream` is ignored. ----
static const int kSms = 148;
static bool mma_ok(const void* x, int dtype, long long sb, long long sn, int D, int H, int N) {
  return (dtype == 1 || dtype == 2) && H <= 8 && D % 128 == 0 && D <= 1024 && (reinterpret_cast<uintptr_t>(x) % 16) == 0 &&
         (sn % 8) == 0 && sb == (long long)N * sn;
}
static int stages_for(int D, size_t fixed, bool wide) {
  if (!wide && (size_t)2 * PM_TT * D * 2 + fixed <= 112 * 1024) return 2;
  int st = 4;
  while (st > 2 && (size_t)st * PM_TT * D * 2 + fixed > 220 * 1024) --st;
  return st;
}
extern "C" int b200clip_attnpool_splits(int B, int N) {
  int S = (2 * kSms) / B;
  const int maxS = (N + 4 * AP_TOK - 1) / (4 * AP_TOK);
  if (S > maxS) S = maxS;
  if (S < 1) S = 1;
  if (S > 64) S = 64;
  return S;
}
extern "C" int b200clip_attnpool_bwd_splits(int B, int N) {
  int S = kSms / B;
  const int maxS = (N + 2 * PM_TT - 1) / (2 * PM_TT);
  if (S > maxS) S = maxS;
  if (S < 1) S = 1;
  return S;
}

template <typename T, int NCH>
static void fwd_cc(const PoolFwdParams& p) {
  emul::launch(emul::Dim{(unsigned)p.B, (unsigned)p.S, 1}, 32 * p.H, [&] { pool_fwd_kernel<T, NCH>(p); },
               2 * (size_t)AP_TOK * p.D * sizeof(T));
}
template <typename T>
static int fwd_t(const PoolFwdParams& p) {
  const int VE = 16 / (int)sizeof(T);
  if (p.D % (32 * VE)) return -22;
  switch (p.D / (32 * VE)) {
    case 1: fwd_cc<T, 1>(p); break;
    case 2: fwd_cc<T, 2>(p); break;
    case 3: fwd_cc<T, 3>(p); break;
    case 4: fwd_cc<T, 4>(p); break;
    case 6: fwd_cc<T, 6>(p); break;
    case 8: fwd_cc<T, 8>(p); break;
    default: return -22;
  }
  return 0;
}
extern "C" int b200clip_attnpool_fwd(const void* x, int dtype, long long sb, long long sn, const unsigned char* mask, long long mb,
                          const float* qt, const float* w, long long wb, long long wh, int B, int N, int D, int H, int S,
                          float* pm, float* pl, float* pa, float drop_p, long long seed, float* pl2, void*) {
  if (mma_ok(x, dtype, sb, sn, D, H, N)) {
    const bool wide = D % 256 == 0 && (size_t)2 * PM_TT * D * 2 + fwd_smem(D, 0) > 112 * 1024;
    const int stages = stages_for(D, fwd_smem(D, 0), wide);
    PmFwdParams p{x, sb, sn, mask, mb, qt, w, wb, wh, pm, pl, pa, B, N, D, H, S, drop_p, (unsigned long long)seed, pl2, stages};
    CUtensorMap tm{x, (uint64_t)B * N, (uint64_t)D, sn};
    const emul::Dim grid{(unsigned)B, (unsigned)S, 1};
    const size_t smem = fwd_smem(D, stages);
    if (dtype == 1) {
      if (wide) emul::launch(grid, 512, [&] { pool_fwd_mma_kernel<__nv_bfloat16, 16>(tm, p); }, smem);
      else emul::launch(grid, 256, [&] { pool_fwd_mma_kernel<__nv_bfloat16, 8>(tm, p); }, smem);
    } else {
      if (wide) emul::launch(grid, 512, [&] { pool_fwd_mma_kernel<__half, 16>(tm, p); }, smem);
      else emul::launch(grid, 256, [&] { pool_fwd_mma_kernel<__half, 8>(tm, p); }, smem);
    }
    return 0;
  }
  PoolFwdParams p{x, sb, sn, mask, mb, qt, w, wb, wh, pm, pl, pa, B, N, D, H, S, drop_p, (unsigned long long)seed, pl2};
  return dtype == 0 ? fwd_t<float>(p) : dtype == 1 ? fwd_t<__nv_bfloat16>(p) : dtype == 2 ? fwd_t<__half>(p) : -22;
}

extern "C" int b200clip_attnpool_merge(const float* pm, const float* pl, const float* pa, int B, int S, int H, int D, float* out,
                            float* out_m, float* out_l, int sum_over_b, const float* pl2, float* out_sa, void*) {
  emul::launch(B * H, 256, [&] { pool_merge_kernel(pm, pl, pa, B, S, H, D, out, out_m, out_l, sum_over_b, pl2, out_sa); });
  return 0;
}

template <typename T, int NCH>
static void bwd_cc(const PoolBwdParams& p, int gy) {
  emul::launch(emul::Dim{(unsigned)p.B, (unsigned)gy, 1}, 256, [&] { pool_bwd_dx_kernel<T, NCH>(p); },
               (2 * (size_t)p.H * p.D + 3 * p.H) * sizeof(float));
}
template <typename T>
static int bwd_t(const PoolBwdParams& p) {
  const int VE = 16 / (int)sizeof(T);
  if (p.D % (32 * VE)) return -22;
  int gy = (2 * kSms + p.B - 1) / p.B;
  const int maxy = (p.N + 7) / 8;
  if (gy > maxy) gy = maxy;
  if (gy < 1) gy = 1;
  switch (p.D / (32 * VE)) {
    case 1: bwd_cc<T, 1>(p, gy); break;
    case 2: bwd_cc<T, 2>(p, gy); break;
    case 3: bwd_cc<T, 3>(p, gy); break;
    case 4: bwd_cc<T, 4>(p, gy); break;
    case 6: bwd_cc<T, 6>(p, gy); break;
    case 8: bwd_cc<T, 8>(p, gy); break;
    default: return -22;
  }
  return 0;
}
static int bwd_any(const void* x, int dtype, long long sb, long long sn, const unsigned char* mask, long long mb,
                   const float* qt, const float* dxbar, const float* xbar, const float* m, const float* l, int B, int N,
                   int D, int H, void* dx, float* ds, const float* sa, const float* dsa, float drop_p, long long seed,
                   const float* dlse, float* part_dq, bool need_mma) {
  if (mma_ok(x, dtype, sb, sn, D, H, N) && (reinterpret_cast<uintptr_t>(dx) % 16) == 0) {
    const int S = b200clip_attnpool_bwd_splits(B, N);
    const size_t dq_smem = part_dq ? 2 * 8 * (PM_TT + 8) * 2 : 0;
    const int stages = stages_for(D, bwd_smem(D, 0) + dq_smem, D % 256 == 0);
    PmBwdParams p{x, sb, sn, mask, mb, qt, dxbar, xbar, m, l, dx, ds, B, N, D, H, S, sa, dsa, drop_p, (unsigned long long)seed,
                  stages, dlse, part_dq};
    CUtensorMap tm{x, (uint64_t)B * N, (uint64_t)D, sn};
    const emul::Dim grid{(unsigned)B, (unsigned)S, 1};
    const size_t smem = bwd_smem(D, stages) + dq_smem;
    const bool wide = D % 256 == 0;
#define EM_BWD(TT_, NW_, DQ_) emul::launch(grid, NW_ * 32, [&] { pool_bwd_mma_kernel<TT_, NW_, DQ_>(tm, p); }, smem)
    if (part_dq) {
      if (dtype == 1) { if (wide) EM_BWD(__nv_bfloat16, 16, true); else EM_BWD(__nv_bfloat16, 8, true); }
      else { if (wide) EM_BWD(__half, 16, true); else EM_BWD(__half, 8, true); }
    } else {
      if (dtype == 1) { if (wide) EM_BWD(__nv_bfloat16, 16, false); else EM_BWD(__nv_bfloat16, 8, false); }
      else { if (wide) EM_BWD(__half, 16, false); else EM_BWD(__half, 8, false); }
    }
#undef EM_BWD
    return 0;
  }
  if (need_mma) return -38;
  PoolBwdParams p{x, sb, sn, mask, mb, qt, dxbar, xbar, m, l, dx, ds, B, N, D, H, sa, dsa, drop_p, (unsigned long long)seed, dlse};
  return dtype == 0 ? bwd_t<float>(p) : dtype == 1 ? bwd_t<__nv_bfloat16>(p) : dtype == 2 ? bwd_t<__half>(p) : -22;
}
extern "C" int b200clip_attnpool_bwd_dx(const void* x, int dtype, long long sb, long long sn, const unsigned char* mask, long long mb,
                             const float* qt, const float* dxbar, const float* xbar, const float* m, const float* l, int B,
                             int N, int D, int H, void* dx, float* ds, const float* sa, const float* dsa, float drop_p,
                             long long seed, const float* dlse, void*) {
  return bwd_any(x, dtype, sb, sn, mask, mb, qt, dxbar, xbar, m, l, B, N, D, H, dx, ds, sa, dsa, drop_p, seed, dlse, nullptr,
                 false);
}
extern "C" int b200clip_attnpool_bwd_dx_dq(const void* x, int dtype, long long sb, long long sn, const unsigned char* mask, long long mb,
                                const float* qt, const float* dxbar, const float* xbar, const float* m, const float* l,
                                int B, int N, int D, int H, void* dx, float* ds, const float* sa, const float* dsa,
                                float drop_p, long long seed, const float* dlse, float* part_dq, void*) {
  return bwd_any(x, dtype, sb, sn, mask, mb, qt, dxbar, xbar, m, l, B, N, D, H, dx, ds, sa, dsa, drop_p, seed, dlse, part_dq,
                 true);
}

// ---- b200clip_rope3d_apply with rope3d.cu's dispatch (plane fast path / generic kernel) ----
template <typename T>
static int rope_launch_emul(const RopeTensor& q, const RopeTensor& k, int ntens, const void* sin_t, const void* cos_t, int B,
                            int Hh, int N, int Dh, float sgn) {
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = Dh % VEC == 0;
  auto aligned = [&](const RopeTensor& t) {
    return (reinterpret_cast<uintptr_t>(t.in) % 16 == 0) && (t.sb * sizeof(T)) % 16 == 0 && (t.sh * sizeof(T)) % 16 == 0 &&
           (t.sn * sizeof(T)) % 16 == 0;
  };
  const long long plane = (long long)N * (Dh / VEC);
  if (vec_ok && aligned(q) && (ntens == 1 || aligned(k))) {
    const int rows = B * Hh;
    const int pblocks = (int)((plane + 255) / 256);
    int groups = (8 * kSms + pblocks * ntens - 1) / (pblocks * ntens);
    if (groups < 1) groups = 1;
    int rpb = (rows + groups - 1) / groups;
    rpb = (rpb + ROPE_UNROLL - 1) / ROPE_UNROLL * ROPE_UNROLL;
    groups = (rows + rpb - 1) / rpb;
    emul::launch(emul::Dim{(unsigned)pblocks, (unsigned)groups, (unsigned)ntens}, 256, [&] {
      rope3d_plane_kernel<T>(q, k, (const T*)sin_t, (const T*)cos_t, rows, Hh, N, Dh, rpb, sgn);
    });
    return 0;
  }
  const long long total = (long long)B * Hh * N * (vec_ok ? Dh / VEC : Dh / 2);
  long long blocks = (total + 255) / 256;
  if (blocks > 64) blocks = 64;                 // grid-stride kernel: fewer CTAs than the device would get, same result
  const emul::Dim grid{(unsigned)blocks, (unsigned)ntens, 1};
  if (vec_ok) emul::launch(grid, 256, [&] { rope3d_kernel<T, VEC>(q, k, (const T*)sin_t, (const T*)cos_t, B, Hh, N, Dh, sgn); });
  else emul::launch(grid, 256, [&] { rope3d_kernel<T, 2>(q, k, (const T*)sin_t, (const T*)cos_t, B, Hh, N, Dh, sgn); });
  return 0;
}
extern "C" int b200clip_rope3d_apply(const void* q, long long qsb, long long qsh, long long qsn, void* q_out, const void* k,
                                     long long ksb, long long ksh, long long ksn, void* k_out, const void* sin_t,
                                     const void* cos_t, int dtype, int B, int Hh, int N, int Dh, int backward, void*) {
  if (!q || !q_out || !sin_t || !cos_t || B <= 0 || Hh <= 0 || N <= 0 || Dh <= 0 || (Dh & 1)) return -22;
  RopeTensor tq{q, q_out, qsb, qsh, qsn};
  RopeTensor tk{k, k_out, ksb, ksh, ksn};
  const int ntens = (k && k_out) ? 2 : 1;
  const float sgn = backward ? -1.f : 1.f;
  switch (dtype) {
    case 0: return rope_launch_emul<float>(tq, tk, ntens, sin_t, cos_t, B, Hh, N, Dh, sgn);
    case 1: return rope_launch_emul<__nv_bfloat16>(tq, tk, ntens, sin_t, cos_t, B, Hh, N, Dh, sgn);
    case 2: return rope_launch_emul<__half>(tq, tk, ntens, sin_t, cos_t, B, Hh, N, Dh, sgn);
    default: return -22;
  }
}

// ---- b200clip_querypool (querypool.cu: one CTA per study, forward or backward) ----
extern "C" int b200clip_querypool(int backward, const float* x, long long sb, long long sn, const float* pos, const float* lnw,
                                  const float* lnb, const float* q, const unsigned char* mask, long long mb, int B, int N,
                                  int D, float eps, float* out, const float* dout, float* dx, float* dpos, float* dlnw,
                                  float* dlnb, float* dq, void*) {
  if (!x || !lnw || !lnb || !q || B <= 0 || N <= 0 || D <= 0) return -22;
  const size_t smem = ((size_t)N * D + 5 * (size_t)N) * sizeof(float);
  QpParams p{x, sb, sn, pos, lnw, lnb, q, mask, mb, B, N, D, eps, out, dout, dx, dpos, dlnw, dlnb, dq};
  if (backward) {
    if (!dout || !dx || !dlnw || !dlnb || !dq) return -22;
    emul::launch(emul::Dim{(unsigned)B, 1, 1}, 256, [&] { querypool_kernel<true>(p); }, smem);
  } else {
    if (!out) return -22;
    emul::launch(emul::Dim{(unsigned)B, 1, 1}, 256, [&] { querypool_kernel<false>(p); }, smem);
  }
  return 0;
}

// ---- b200clip_multipos_* (multipos.cu launch geometry on a 148-SM device) ----
static int mp_chunks(int N, int M) {
  const int col_blocks = (M + 127) / 128;
  int ch = (4 * kSms + col_blocks - 1) / col_blocks;
  if (ch > N / 32) ch = N / 32;
  if (ch < 1) ch = 1;
  if (ch > 256) ch = 256;
  return ch;
}
extern "C" int b200clip_multipos_workspace_bytes(int N, int M) { return mp_chunks(N, M) * M * (int)sizeof(MpAcc); }
extern "C" int b200clip_multipos_fwd(const float* L, long long ldl, const float* pw, const float* mk, long long ldw, int N, int M,
                                     int mode, float eps, int reduce_sum, float* rstat, float* cstat, float* coef,
                                     float* loss_out, void* workspace, void*) {
  if (N <= 0 || M <= 0 || mode < 0 || mode > 2 || (!pw && !mk)) return -22;
  const int ch = mp_chunks(N, M);
  float* imp_r = mode == 2 ? coef : nullptr;
  float* imp_c = mode == 2 ? coef + N : nullptr;
  emul::launch(N, 256, [&] { mp_row_stats_kernel(L, ldl, pw, mk, ldw, N, M, reinterpret_cast<float4*>(rstat), imp_r); });
  emul::launch(emul::Dim{(unsigned)((M + 127) / 128), (unsigned)ch, 1}, 128,
               [&] { mp_col_partial_kernel(L, ldl, pw, mk, ldw, N, M, ch, reinterpret_cast<MpAcc*>(workspace), mode == 2); });
  emul::launch((M + 127) / 128, 128, [&] {
    mp_col_merge_kernel(reinterpret_cast<const MpAcc*>(workspace), M, ch, reinterpret_cast<float4*>(cstat), imp_c);
  });
  emul::launch(1, 1024, [&] {
    mp_finalize_kernel(reinterpret_cast<const float4*>(rstat), reinterpret_cast<const float4*>(cstat), N, M, mode, eps,
                       reduce_sum, coef, loss_out);
  });
  return 0;
}
extern "C" int b200clip_multipos_bwd(const float* L, long long ldl, const float* pw, const float* mk, long long ldw, int N, int M,
                                     const float* rstat, const float* cstat, const float* coef, const float* gmul, float* dL,
                                     long long ldd, void*) {
  if (N <= 0 || M <= 0) return -22;
  emul::launch(N, 256, [&] {
    mp_backward_kernel(L, ldl, pw, mk, ldw, N, M, reinterpret_cast<const float4*>(rstat), reinterpret_cast<const float4*>(cstat),
                       coef, gmul, dL, ldd);
  });
  return 0;
}

// ---- b200clip_dense_gt_ranks / b200clip_dense_rank_metrics (dense_metrics.cu) ----
template <typename T>
static void ranks_emul(const void* sim, long long ld, int N, int M, const int* gt, int G, int sanitize, int* ranks) {
  const T* p = reinterpret_cast<const T*>(sim);
  if (G <= 1) emul::launch(N, 256, [&] { dense_gt_ranks_kernel<T, 1>(p, ld, N, M, gt, G, sanitize, ranks); });
  else if (G <= 4) emul::launch(N, 256, [&] { dense_gt_ranks_kernel<T, 4>(p, ld, N, M, gt, G, sanitize, ranks); });
  else if (G <= 8) emul::launch(N, 256, [&] { dense_gt_ranks_kernel<T, 8>(p, ld, N, M, gt, G, sanitize, ranks); });
  else emul::launch(N, 256, [&] { dense_gt_ranks_kernel<T, 16>(p, ld, N, M, gt, G, sanitize, ranks); });
}
extern "C" int b200clip_dense_gt_ranks(const void* sim, int dtype, long long ld, int N, int M, const int* gt, int G,
                                       int sanitize, int* ranks, void*) {
  if (N <= 0 || M <= 0 || G <= 0 || G > 16) return -22;
  if (dtype == 0) ranks_emul<float>(sim, ld, N, M, gt, G, sanitize, ranks);
  else if (dtype == 1) ranks_emul<__nv_bfloat16>(sim, ld, N, M, gt, G, sanitize, ranks);
  else if (dtype == 2) ranks_emul<__half>(sim, ld, N, M, gt, G, sanitize, ranks);
  else return -22;
  return 0;
}
extern "C" int b200clip_dense_rank_metrics(const int* ranks, const int* gsize, int N, int G, int M, const int* recall_k, int nrk,
                                           const int* ndcg_k, int nnk, int* best, double* rr, double* ap, unsigned char* hit,
                                           double* ndcg, void*) {
  if (N <= 0 || G <= 0 || G > 16 || M <= 0) return -22;
  emul::launch((N + 255) / 256, 256, [&] {
    dense_rank_metrics_kernel(ranks, gsize, N, G, M, recall_k, nrk, ndcg_k, nnk, best, rr, ap, hit, ndcg);
  });
  return 0;
}
