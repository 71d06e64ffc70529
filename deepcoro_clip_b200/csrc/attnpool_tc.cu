// K8 on the 5th-generation tensor cores: the single-query attention pool (reference
// models/attention_pool.py:77-99) for 16-bit x with D in {128, 256, 384, 512} and heads <= 8 — the C3 shape
// (x [32, 3136, 512] bf16). The mma.sync kernels of attnpool_mma.cu were issue-bound (about 870 warp instructions and
// three CTA barriers per 32-token tile, 0.26 of the HBM peak); here the two skinny contractions are tcgen05 MMAs issued by
// one thread, the x tile is staged ONCE by TMA and read by the tensor core both K-major (scores) and MN-major (weighted
// sum), and four epilogue warps do the per-token softmax arithmetic straight out of TMEM.
//
// forward, one CTA per (batch row, token split), 64-token tiles in a 3-deep TMA ring:
//   P1  S[tok, 16]   = x[tok, :] . [qt_hi ; qt_lo]^T          M = 64 (tokens), N = 16, K = D     (A = x tile, K-major)
//   epi p[tok, h]    = exp(S_h - m_ref_h)   (lazy reference maximum: rescale only when a score exceeds it by 8 nats)
//       P[16, tok]   = [p_hi ; p_lo] (bf16) written to shared memory as the K-major B operand of
//   P2  acc[d, 16]  += x[tok, d]^T . P^T                       M = 128 (channels), N = 16, K = 64 (A = x tile, MN-major)
//   end partial (m_ref, l, acc_hi + acc_lo) per (b, split) for the existing merge kernel.
// backward, same tiling, 2-deep ring:
//   P1  [S | T][tok, 32] = x . [qt_hi ; dxbar_hi ; qt_lo ; dxbar_lo]^T                       M = 64, N = 32, K = D
//   epi a = exp(S - m) / l,  ds = a (kappa (T + dsa) - c)  ->  C[32, tok] = [ds_hi ; a_hi ; ds_lo ; a_lo]
//   P2  dx^T[d, tok] = [qt_hi ; dxbar_hi]^T . (C_hi + C_lo)     M = 128, N = 64, K = 16 twice (A = the P1 operand, MN-major)
//   P3  dq^T[d, 16] += x^T . [ds_hi ; ds_lo]^T                  M = 128, N = 16, K = 64     (query gradient, same pass)
//   epi dx^T tile -> 16-bit, pairs of channels, 64-byte segments per token straight to global memory.
// x is read once forward and once backward (the ds [B, H, N] round trip and the second read of x for the query gradient
// of the mma.sync path are gone). hi + lo bf16 splits keep the contractions at fp32-level accuracy as before.
#include "common.cuh"
#include "host_api.h"

namespace b2 {

constexpr int PT_TT = 64;          // tokens per tile
constexpr int PT_THREADS = 192;    // warp 0 TMA producer, warp 1 MMA issuer (+ TMEM owner), warps 2..5 epilogue
constexpr int PT_BWD_THREADS = 320;   // backward: + warps 6..9 that drain dx^T while warps 2..5 work on the next tile
constexpr int PT_BOX = PT_TT * 128;   // one [64 tokens x 64 channels] 16-bit box

__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// TMEM -> registers in the MMA accumulator-fragment layout: 16 lanes x 64 columns; thread t receives, for each 8-column
// group j, r[4j], r[4j+1] = (lane t/4, columns 8j + 2(t%4), +1) and r[4j+2], r[4j+3] = (lane t/4 + 8, same columns)
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// four 8x8 16-bit matrices, stored TRANSPOSED: register i of thread t holds M_i[2(t%4)][t/4] (low) and M_i[2(t%4)+1][t/4]
// (high); row r of matrix i goes to the 16 bytes at the address supplied by thread 8i + r
__device__ __forceinline__ void stsm_x4_trans(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};"
               ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t smem_src, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// CTA-uniform OR over the `nthreads` threads of named barrier `id`
__device__ __forceinline__ bool named_bar_or(int id, int nthreads, bool pred) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 p, %3, 0;\n\t"
      "bar.red.or.pred q, %1, %2, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(r)
      : "r"(id), "r"(nthreads), "r"((uint32_t)pred)
      : "memory");
  return r != 0;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
// Hand-built operands (qt, dxbar, P, C) use x's own 16-bit format: tcgen05.mma kind::f16 takes A and B in the SAME format
// (an fp16 A against a bf16 B raised an illegal-instruction fault on sm_100a).
__device__ __forceinline__ uint16_t op_bits(float v, int fp16) {
  return fp16 ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float op_val(uint16_t b, int fp16) {
  return fp16 ? __half2float(__ushort_as_half(b)) : __bfloat162float(__ushort_as_bfloat16(b));
}
// instruction descriptor, both operands bf16 (fp16 = 0) or both fp16 (fp16 = 1), fp32 accumulate
__device__ __forceinline__ uint32_t pt_idesc(int M, int N, int a_mn_major, int b_mn_major, int fp16) {
  const uint32_t fmt = fp16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16) |
         (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// byte offset of element (row, col) of a [rows x 64] 16-bit sub-tile with 128-byte rows, SWIZZLE_128B (8-row groups of
// 1024 B): the layout TMA writes and both the K-major and the MN-major tcgen05 descriptors read
__device__ __forceinline__ uint32_t sw128_off(int row, int col) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((col >> 3) ^ row) & 7) << 4) + (col & 7) * 2);
}
// MN-major operand spanning two adjacent 64-element blocks `lbo` bytes apart
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo) { return make_smem_desc_sw128(addr, lbo); }

// One tile = D/64 TMA boxes of [64 tokens x 64 channels] (the tallest box the tile allows: 32 / 16 / 8-row boxes measured
// 1 / 6 / 23 % slower — the issue cost per box outweighs any DRAM-locality gain)
__device__ __forceinline__ void pt_load_tile(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int KC, int tok, int b) {
  for (int kc = 0; kc < KC; ++kc) tma_load_3d(dst + kc * PT_BOX, tm, bar, kc * 64, tok, b);
}

// plain bulk copy global -> shared (both 16-byte aligned, size % 16 == 0), completion on an mbarrier
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct PtFwdParams {
  const unsigned char* mask; long long mb;
  const float* qt;                     // [H, D]
  const void* qt_img;                  // or: the prepared operand image (pool_prep), D/64 x [16 x 64] 16-bit, swizzled
  float* part_m; float* part_l; float* part_l2; float* part_acc;
  int B, N, D, H, S, fp16;
  float drop_p; unsigned long long drop_seed;
};

// dynamic shared memory (1024-aligned): ring NS x (D/64) boxes of 8 KB | qt operand (D/64) x [16 x 64] (2 KB each) |
// P operand 2 x [16 x 64] (2 KB each) | red [2][4][8] fp32 | barriers
template <int NS>
__global__ void __launch_bounds__(PT_THREADS, 1) pool_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmx, PtFwdParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int D = p.D, KC = D / 64, MB = D / 128;
  const uint32_t tile_bytes = (uint32_t)KC * PT_BOX;
  const uint32_t ring = smem_u32(smem);
  const uint32_t qt_op = ring + NS * tile_bytes;
  const uint32_t p_op = qt_op + KC * 2048;
  float* red = reinterpret_cast<float*>(smem + (size_t)NS * tile_bytes + KC * 2048 + 2 * 2048);
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 64);
  uint64_t* full = bars;            // [NS] tile landed (TMA tx)
  uint64_t* empty = bars + NS;      // [NS] P2 of the tile complete (tcgen05.commit)
  uint64_t* s_full = bars + 2 * NS; // [2]  scores ready (commit)
  uint64_t* p_ready = s_full + 2;   // [2]  P written (128 epilogue threads)
  uint64_t* acc_done = p_ready + 2; // [1]
  uint64_t* op_full = acc_done + 1; // [1]  operand image landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(op_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x, sp = blockIdx.y;
  const int TPB = (p.N + PT_TT - 1) / PT_TT;
  const int tile0 = (int)((long long)TPB * sp / p.S), tile1 = (int)((long long)TPB * (sp + 1) / p.S);
  const int T = tile1 - tile0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], 128); }
    mbar_init(acc_done, 1);
    mbar_init(op_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmx);
    if (p.qt_img) {
      mbar_expect_tx(op_full, KC * 2048);
      bulk_load(qt_op, p.qt_img, KC * 2048, op_full);
    }
    // the first NS tiles start loading before anything else is set up
    for (int i = 0; i < NS && i < T; ++i) {
      mbar_expect_tx(&full[i], tile_bytes);
      pt_load_tile(ring + i * tile_bytes, &tmx, &full[i], KC, (tile0 + i) * PT_TT, b);
    }
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 128); tmem_relinquish(); }
  // qt -> [qt_hi (rows 0..7) ; qt_lo (rows 8..15)] K-major operand, heads >= H zero (unless the image was prepared)
  for (int i = threadIdx.x; i < (p.qt_img ? 0 : 8 * D); i += PT_THREADS) {
    const int h = i / D, d = i - h * D;
    const float v = h < p.H ? p.qt[(size_t)h * D + d] : 0.f;
    const uint16_t hi = op_bits(v, p.fp16);
    const uint32_t base = qt_op + (d >> 6) * 2048;
    sts16(base + sw128_off(h, d & 63), hi);
    sts16(base + sw128_off(8 + h, d & 63), op_bits(v - op_val(hi, p.fp16), p.fp16));
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t S_COL = 0, ACC_COL = 32;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      for (int i = NS; i < T; ++i) {
        const int slot = i % NS;
        mbar_wait(&empty[slot], ((i / NS) & 1) ^ 1);
        mbar_expect_tx(&full[slot], tile_bytes);
        pt_load_tile(ring + slot * tile_bytes, &tmx, &full[slot], KC, (tile0 + i) * PT_TT, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      const uint32_t idesc1 = pt_idesc(64, 16, 0, 0, p.fp16);
      const uint32_t idesc2 = pt_idesc(128, 16, 1, 0, p.fp16);
      auto p1 = [&](int i) {
        const int slot = i % NS;
        mbar_wait(&full[slot], (i / NS) & 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + S_COL + (i & 1) * 16;
        for (int kc = 0; kc < KC; ++kc) {
          const uint64_t adesc = make_smem_desc_sw128(ring + slot * tile_bytes + kc * PT_BOX, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(qt_op + kc * 2048, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc1, (kc | k) != 0);
        }
        tc_commit(&s_full[i & 1]);
      };
      if (p.qt_img) mbar_wait(op_full, 0);
      p1(0);
      for (int i = 0; i < T; ++i) {
        if (i + 1 < T) p1(i + 1);
        mbar_wait(&p_ready[i & 1], (i >> 1) & 1);
        tc_fence_after();
        const int slot = i % NS;
        const uint32_t tile = ring + slot * tile_bytes;
        const uint64_t bdesc = make_smem_desc_sw128(p_op + (i & 1) * 2048, 1024);
        for (int mb = 0; mb < MB; ++mb) {
          const uint64_t adesc = desc_mn(tile + 2 * mb * PT_BOX, PT_BOX);
#pragma unroll
          for (int ks = 0; ks < PT_TT / 16; ++ks)
            mma_ss(tmem_base + ACC_COL + mb * 16, adesc + uint64_t(ks * (2048 >> 4)), bdesc + 2 * ks, idesc2,
                   (i | ks) != 0);
        }
        tc_commit(&empty[slot]);
      }
      tc_commit(acc_done);
    }
  } else {
    // ===================== epilogue warps 2..5: TMEM lane quarter q = warp & 3 =====================
    const int q = warp & 3;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const int r = q * 16 + lane;                 // token row of the tile (M = 64: 16 rows per lane quarter, lanes 0..15)
    const unsigned char* mk = p.mask ? p.mask + (long long)b * p.mb : nullptr;
    const float keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    float m_ref[8], l[8], l2[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) { m_ref[h] = -INFINITY; l[h] = 0.f; l2[h] = 0.f; }
    for (int i = 0; i < T; ++i) {
      const int sb = i & 1;
      mbar_wait(&s_full[sb], (i >> 1) & 1);
      tc_fence_after();
      uint32_t c[16];
      tmem_ld16(tmem_base + S_COL + sb * 16 + lane_off, c);
      tc_wait_ld();
      const int n = (tile0 + i) * PT_TT + r;
      const bool live = lane < 16 && n < p.N && !(mk && mk[n]);
      float s[8];
      bool over = false;
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        s[h] = __uint_as_float(c[h]) + __uint_as_float(c[8 + h]);
        over |= live && h < p.H && s[h] > m_ref[h] + 8.f;
      }
      if (named_bar_or(1, 128, over)) {
        // a score left the window of the reference maximum: new reference = running maximum, rescale sums and accumulator
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const float v = warp_max((live && h < p.H) ? s[h] : -INFINITY);
          if (lane == 0) red[q * 8 + h] = v;
        }
        named_bar_sync(2, 128);
        float sc[8];
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const float mt = fmaxf(fmaxf(red[h], red[8 + h]), fmaxf(red[16 + h], red[24 + h]));
          const float mn = fmaxf(m_ref[h], mt);
          sc[h] = (mn == m_ref[h]) ? 1.f : (m_ref[h] == -INFINITY ? 0.f : __expf(m_ref[h] - mn));
          m_ref[h] = mn;
          l[h] *= sc[h];
          l2[h] *= sc[h];
        }
        if (i > 0) {
          mbar_wait(&empty[(i - 1) % NS], ((i - 1) / NS) & 1);     // P2 of every earlier tile is complete
          tc_fence_after();
          for (int mb = 0; mb < MB; ++mb) {
            uint32_t a[16];
            tmem_ld16(tmem_base + ACC_COL + mb * 16 + lane_off, a);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = __float_as_uint(__uint_as_float(a[j]) * sc[j & 7]);
            tmem_st16(tmem_base + ACC_COL + mb * 16 + lane_off, a);
          }
          tc_wait_st();
        }
      }
      if (lane < 16) {
        const uint32_t pb = p_op + sb * 2048;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          float pv = (live && h < p.H) ? __expf(s[h] - m_ref[h]) : 0.f;
          l[h] += pv;
          if (p.drop_p > 0.f) {
            pv = (pv != 0.f && attn_keep(p.drop_seed, b * p.H + h, n, p.drop_p)) ? pv * keep_scale : 0.f;
            l2[h] += pv;
          }
          const uint16_t hi = op_bits(pv, p.fp16);
          sts16(pb + sw128_off(h, r), hi);
          sts16(pb + sw128_off(8 + h, r), op_bits(pv - op_val(hi, p.fp16), p.fp16));
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_ready[sb]);
    }
    // ---- partial results of this (b, split) ----
    mbar_wait(acc_done, 0);
    tc_fence_after();
    const size_t slot = (size_t)b * p.S + sp;
    for (int mb = 0; mb < MB; ++mb) {
      uint32_t a[16];
      tmem_ld16(tmem_base + ACC_COL + mb * 16 + lane_off, a);
      tc_wait_ld();
      const int d = mb * 128 + q * 32 + lane;
#pragma unroll
      for (int h = 0; h < 8; ++h)
        if (h < p.H) p.part_acc[(slot * p.H + h) * D + d] = __uint_as_float(a[h]) + __uint_as_float(a[8 + h]);
    }
    named_bar_sync(2, 128);
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const float v = warp_sum(l[h]), v2 = warp_sum(l2[h]);
      if (lane == 0) { red[q * 8 + h] = v; red[32 + q * 8 + h] = v2; }
    }
    named_bar_sync(2, 128);
    const int et = threadIdx.x - 64;
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      if (et == h && h < p.H) {
        const float lt = red[h] + red[8 + h] + red[16 + h] + red[24 + h];
        const float lt2 = red[32 + h] + red[40 + h] + red[48 + h] + red[56 + h];
        p.part_m[slot * p.H + h] = m_ref[h];
        p.part_l[slot * p.H + h] = lt;
        if (p.part_l2) p.part_l2[slot * p.H + h] = p.drop_p > 0.f ? lt2 : lt;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

struct PtBwdParams {
  const unsigned char* mask; long long mb;
  const float* qt;        // [H, D]
  const float* dxbar;     // [B, H, D]
  const float* xbar;      // [B, H, D]
  const void* w_img;      // or: prepared operand images (pool_tail_bwd), [B][D/64][32 x 64] 16-bit, swizzled ...
  const float* cdot;      // ... with c_h = dxbar_h . xbar_h [B, H]
  const float* m; const float* l;   // [B, H]
  void* dx;               // [B, N, D] contiguous, 16-bit
  float* part_dq;         // [B, S, H, D] or null
  const float* sa; const float* dsa; const float* dlse;
  int B, N, D, H, S, fp16;
  float drop_p; unsigned long long drop_seed;
};

// dynamic shared memory: ring NS x (D/64) boxes | W operand (D/64) x [32 x 64] (4 KB each: rows qt_hi, dxbar_hi, qt_lo,
// dxbar_lo) | C operand 2 x [32 x 64] (4 KB each: rows ds_hi, a_hi, ds_lo, a_lo) | c, m, 1/l, dsa [4][8] fp32 | barriers
template <int NS>
__global__ void __launch_bounds__(PT_BWD_THREADS, 1) pool_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmx,
                                                                    const __grid_constant__ CUtensorMap tmdx, PtBwdParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int D = p.D, KC = D / 64, MB = D / 128;
  const uint32_t tile_bytes = (uint32_t)KC * PT_BOX;
  const uint32_t ring = smem_u32(smem);
  const uint32_t w_op = ring + NS * tile_bytes;
  const uint32_t c_op = w_op + KC * 4096;
  const uint32_t stage = c_op + 2 * 4096;          // 2 x [64 tokens x 128 channels] (two 8 KB boxes each): dx staging
  float* s_c = reinterpret_cast<float*>(smem + (size_t)NS * tile_bytes + KC * 4096 + 2 * 4096 + 2 * 2 * PT_BOX);
  float* s_m = s_c + 8;
  float* s_il = s_m + 8;
  float* s_dsa = s_il + 8;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_dsa + 8);
  uint64_t* full = bars;               // [NS]
  uint64_t* empty = bars + NS;         // [NS] P1 and P3 of the tile complete
  uint64_t* st_full = bars + 2 * NS;   // [2]
  uint64_t* c_ready = st_full + 2;     // [2]  (128 epilogue threads)
  uint64_t* dx_full = c_ready + 2;     // [1]  P2 (+ P3) of the tile complete
  uint64_t* dx_empty = dx_full + 1;    // [1]  dx^T drained out of TMEM (128 epilogue threads)
  uint64_t* op_full = dx_empty + 1;    // [1]  operand image landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(op_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x, sp = blockIdx.y;
  const int TPB = (p.N + PT_TT - 1) / PT_TT;
  const int tile0 = (int)((long long)TPB * sp / p.S), tile1 = (int)((long long)TPB * (sp + 1) / p.S);
  const int T = tile1 - tile0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&st_full[i], 1); mbar_init(&c_ready[i], 128); }
    mbar_init(dx_full, 1);
    mbar_init(dx_empty, 128);
    mbar_init(op_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmdx);
    if (p.w_img) {
      mbar_expect_tx(op_full, KC * 4096);
      bulk_load(w_op, reinterpret_cast<const unsigned char*>(p.w_img) + (size_t)b * KC * 4096, KC * 4096, op_full);
    }
    for (int i = 0; i < NS && i < T; ++i) {
      mbar_expect_tx(&full[i], tile_bytes);
      pt_load_tile(ring + i * tile_bytes, &tmx, &full[i], KC, (tile0 + i) * PT_TT, b);
    }
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < (p.w_img ? 0 : 8 * D); i += PT_BWD_THREADS) {
    const int h = i / D, d = i - h * D;
    const float qv = h < p.H ? p.qt[(size_t)h * D + d] : 0.f;
    const float dv = h < p.H ? p.dxbar[((size_t)b * p.H + h) * D + d] : 0.f;
    const uint16_t qh = op_bits(qv, p.fp16), dh = op_bits(dv, p.fp16);
    const uint32_t base = w_op + (d >> 6) * 4096;
    sts16(base + sw128_off(h, d & 63), qh);
    sts16(base + sw128_off(8 + h, d & 63), dh);
    sts16(base + sw128_off(16 + h, d & 63), op_bits(qv - op_val(qh, p.fp16), p.fp16));
    sts16(base + sw128_off(24 + h, d & 63), op_bits(dv - op_val(dh, p.fp16), p.fp16));
  }
  if (warp >= 2) {
    // c_h = dxbar_h . xbar_h (+ dsa_h sa_h) (- dlse_h): warp 2 + h takes head h
    for (int h = warp - 2; h < 8; h += 8) {
      float c = 0.f;
      if (p.cdot) {
        c = h < p.H ? p.cdot[b * p.H + h] : 0.f;
      } else {
        if (h < p.H)
          for (int d = lane; d < D; d += 32)
            c = fmaf(p.dxbar[((size_t)b * p.H + h) * D + d], p.xbar[((size_t)b * p.H + h) * D + d], c);
        c = warp_sum(c);
      }
      if (p.dsa && h < p.H) c = fmaf(p.dsa[b * p.H + h], p.sa[b * p.H + h], c);
      if (p.dlse && h < p.H) c -= p.dlse[b * p.H + h];
      if (lane == 0) {
        s_c[h] = c;
        s_m[h] = h < p.H ? p.m[b * p.H + h] : 0.f;
        s_il[h] = h < p.H ? 1.f / p.l[b * p.H + h] : 0.f;
        s_dsa[h] = (p.dsa && h < p.H) ? p.dsa[b * p.H + h] : 0.f;
      }
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t ST_COL = 0, DQ_COL = 64, DX_COL = 128;   // [S|T] 2 x 32 | dq^T MB x 16 | dx^T MB x 64

  if (warp == 0) {
    if (elect_one()) {
      for (int i = NS; i < T; ++i) {
        const int slot = i % NS;
        mbar_wait(&empty[slot], ((i / NS) & 1) ^ 1);
        mbar_expect_tx(&full[slot], tile_bytes);
        pt_load_tile(ring + slot * tile_bytes, &tmx, &full[slot], KC, (tile0 + i) * PT_TT, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc1 = pt_idesc(64, 32, 0, 0, p.fp16);       // [S|T] = x . W^T
      const uint32_t idesc2 = pt_idesc(128, 64, 1, 1, p.fp16);          // dx^T = W_hi^T . C     (both MN-major)
      const uint32_t idesc3 = pt_idesc(128, 16, 1, 0, p.fp16);      // dq^T = x^T . ds^T
      auto p1 = [&](int i) {
        const int slot = i % NS;
        mbar_wait(&full[slot], (i / NS) & 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ST_COL + (i & 1) * 32;
        for (int kc = 0; kc < KC; ++kc) {
          const uint64_t adesc = make_smem_desc_sw128(ring + slot * tile_bytes + kc * PT_BOX, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(w_op + kc * 4096, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc1, (kc | k) != 0);
        }
        tc_commit(&st_full[i & 1]);
      };
      if (p.w_img) mbar_wait(op_full, 0);
      p1(0);
      for (int i = 0; i < T; ++i) {
        if (i + 1 < T) p1(i + 1);
        mbar_wait(&c_ready[i & 1], (i >> 1) & 1);
        if (i > 0) mbar_wait(dx_empty, (i - 1) & 1);       // the previous tile's dx^T has left TMEM
        tc_fence_after();
        const int slot = i % NS;
        const uint32_t tile = ring + slot * tile_bytes;
        const uint32_t cb = c_op + (i & 1) * 4096;
        // P2: dx^T[d, tok] = sum_k W[k, d] C[k, tok], k = 0..15 against C rows 0..15 (hi) and again against rows 16..31 (lo)
        for (int mb = 0; mb < MB; ++mb) {
          const uint64_t adesc = desc_mn(w_op + 2 * mb * 4096, 4096);
          mma_ss(tmem_base + DX_COL + mb * 64, adesc, desc_mn(cb, 0), idesc2, 0u);
          mma_ss(tmem_base + DX_COL + mb * 64, adesc, desc_mn(cb + 2048, 0), idesc2, 1u);
        }
        tc_commit(dx_full);
        // P3: dq^T[d, 16] += sum_tok x[tok, d] [ds_hi ; ds_lo][., tok]   (C rows 0..7 and 16..23: 8-row groups 2048 B apart)
        if (p.part_dq) {
          uint64_t bdesc = make_smem_desc_sw128(cb, 1024);
          bdesc = (bdesc & ~(uint64_t(0x3FFF) << 32)) | (uint64_t((2048 >> 4) & 0x3FFF) << 32);
          for (int mb = 0; mb < MB; ++mb) {
            const uint64_t adesc = desc_mn(tile + 2 * mb * PT_BOX, PT_BOX);
#pragma unroll
            for (int ks = 0; ks < PT_TT / 16; ++ks)
              mma_ss(tmem_base + DQ_COL + mb * 16, adesc + uint64_t(ks * (2048 >> 4)), bdesc + 2 * ks, idesc3, (i | ks) != 0);
          }
        }
        tc_commit(&empty[slot]);
      }
    }
  } else if (warp < 6) {
    // ===================== warps 2..5: attention weights and score gradients of tile i -> C operand =====================
    const int q = warp & 3;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const int r = q * 16 + lane;
    const unsigned char* mk = p.mask ? p.mask + (long long)b * p.mb : nullptr;
    const float keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    for (int i = 0; i < T; ++i) {
      const int sb = i & 1;
      mbar_wait(&st_full[sb], (i >> 1) & 1);
      tc_fence_after();
      uint32_t c[32];
      tmem_ld32(tmem_base + ST_COL + sb * 32 + lane_off, c);
      tc_wait_ld();
      const int tok0 = (tile0 + i) * PT_TT;
      const int n = tok0 + r;
      const bool live = lane < 16 && n < p.N && !(mk && mk[n]);
      if (lane < 16) {
        const uint32_t cb = c_op + sb * 4096;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const float sv = __uint_as_float(c[h]) + __uint_as_float(c[16 + h]);
          const float tv = __uint_as_float(c[8 + h]) + __uint_as_float(c[24 + h]);
          const float a0 = (live && h < p.H) ? __expf(sv - s_m[h]) * s_il[h] : 0.f;
          float kap = 1.f;
          if (p.drop_p > 0.f && h < p.H) kap = attn_keep(p.drop_seed, b * p.H + h, n, p.drop_p) ? keep_scale : 0.f;
          const float dsv = a0 * (kap * (tv + s_dsa[h]) - s_c[h]);
          const float av = a0 * kap;
          const uint16_t dh = op_bits(dsv, p.fp16), ah = op_bits(av, p.fp16);
          sts16(cb + sw128_off(h, r), dh);
          sts16(cb + sw128_off(8 + h, r), ah);
          sts16(cb + sw128_off(16 + h, r), op_bits(dsv - op_val(dh, p.fp16), p.fp16));
          sts16(cb + sw128_off(24 + h, r), op_bits(av - op_val(ah, p.fp16), p.fp16));
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&c_ready[sb]);
    }
    if (p.part_dq) {
      const int last = T - 1;
      mbar_wait(&empty[last % NS], (last / NS) & 1);        // P3 of the last tile (and everything before) complete
      tc_fence_after();
      const size_t slot = (size_t)b * p.S + sp;
      for (int mb = 0; mb < MB; ++mb) {
        uint32_t a[16];
        tmem_ld16(tmem_base + DQ_COL + mb * 16 + lane_off, a);
        tc_wait_ld();
        const int d = mb * 128 + q * 32 + lane;
#pragma unroll
        for (int h = 0; h < 8; ++h)
          if (h < p.H) p.part_dq[(slot * p.H + h) * D + d] = __uint_as_float(a[h]) + __uint_as_float(a[8 + h]);
      }
      tc_fence_before();
    }
  } else {
    // ===================== warps 6..9: drain dx^T of tile i while warps 2..5 are on tile i + 1 =====================
    const int q = warp & 3;
    int blk = 0;
    for (int i = 0; i < T; ++i) {
      const int tok0 = (tile0 + i) * PT_TT;
      // ---- per 128-channel block: TMEM (fragment layout) -> 16-bit pairs -> stmatrix.trans into a
      //      [64 tokens x 128 channels] staging block (rows = tokens) -> TMA store (rows past N are clipped) ----
      mbar_wait(dx_full, i & 1);
      tc_fence_after();
      for (int mb = 0; mb < MB; ++mb, ++blk) {
        const uint32_t sbuf = stage + (blk & 1) * (2 * PT_BOX);
        if (threadIdx.x == 192) bulk_wait_read<1>();        // the store that last read this buffer (two blocks ago) is done
        named_bar_sync(2, 128);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t v[32];
          tmem_ld_16x256b_x8(tmem_base + DX_COL + mb * 64 + (uint32_t(q * 32 + g * 16) << 16), v);
          tc_wait_ld();
          // matrix mi = lane / 8 of each stmatrix: channels (mi & 1) * 8 .. + 7 of this 16-lane group, tokens 8 (2 jj + mi / 2) ..
          const int mi = lane >> 3, chunk = q * 4 + g * 2 + (mi & 1);          // 16-byte chunk (8 channels) of the 128
          const uint32_t cbase = sbuf + (chunk >> 3) * PT_BOX;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int tok = 8 * (2 * jj + (mi >> 1)) + (lane & 7);
            uint32_t r[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float lo = __uint_as_float(v[4 * (2 * jj + (e >> 1)) + 2 * (e & 1)]);
              const float hi = __uint_as_float(v[4 * (2 * jj + (e >> 1)) + 2 * (e & 1) + 1]);
              r[e] = p.fp16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
            }
            stsm_x4_trans(cbase + tok * 128 + ((((chunk & 7) ^ tok) & 7) << 4), r[0], r[1], r[2], r[3]);
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (threadIdx.x == 192) {
          tma_store_3d(&tmdx, sbuf, mb * 128, tok0, b);
          tma_store_3d(&tmdx, sbuf + PT_BOX, mb * 128 + 64, tok0, b);
          bulk_commit();
        }
      }
      tc_fence_before();
      mbar_arrive(dx_empty);
    }
    if (threadIdx.x == 192) bulk_wait_all();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace b2

namespace b2host {
using namespace b2;

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [B, N, D] 16-bit tensor (contiguous), box [1, 64 tokens, 64 channels], SWIZZLE_128B; tokens past N are zero-filled
static int make_tmap_x3d(CUtensorMap* out, const void* base, int B, int N, int D) {
  static EncodeTiledFn3 enc = nullptr;
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return B2_ENOSYS;
    enc = reinterpret_cast<EncodeTiledFn3>(fp);
  }
  cuuint64_t gdim[3] = {(cuuint64_t)D, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t gstride[2] = {(cuuint64_t)D * 2, (cuuint64_t)N * D * 2};
  cuuint32_t box[3] = {64, PT_TT, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? B2_OK : B2_EINVAL;
}

bool attnpool_tc_ok(const void* x, int dtype, long long sb, long long sn, int N, int D, int H) {
  return (dtype == 1 || dtype == 2) && H >= 1 && H <= 8 && D % 128 == 0 && D <= 512 && sn == D && sb == (long long)N * D &&
         (reinterpret_cast<uintptr_t>(x) % 16) == 0;
}

// Token splits per batch row: (B x S) CTAs, one per SM. Picks the S that maximises (SM occupancy of the waves) x (tile
// balance of the splits); B200CLIP_POOL_SPLITS overrides.
int attnpool_tc_splits(int B, int N) {
  if (B <= 0 || N <= 0) return 1;
  const int tpb = (N + PT_TT - 1) / PT_TT;
  if (const char* e = getenv("B200CLIP_POOL_SPLITS")) {
    const int s = atoi(e);
    if (s >= 1) return s > tpb ? tpb : s;
  }
  const int sms = sm_count();
  int best = 1;
  double best_eff = 0.0;
  for (int S = 1; S <= tpb && S <= 64; ++S) {
    const long long ctas = (long long)B * S;
    const long long waves = (ctas + sms - 1) / sms;
    const int per = (tpb + S - 1) / S;
    if (per < 3 && S > 1) break;
    const double eff = (double)ctas / (double)(waves * sms) * (double)tpb / (double)(S * per) / (1.0 + 0.15 * (waves - 1));
    if (eff > best_eff + 1e-9) { best_eff = eff; best = S; }
  }
  return best;
}

static size_t pt_fwd_smem(int D, int ns) { return (size_t)ns * (D / 64) * PT_BOX + (D / 64) * 2048 + 2 * 2048 + 64 * 4 + 16 * 8 + 16 + 1024; }
static size_t pt_bwd_smem(int D, int ns) {
  return (size_t)ns * (D / 64) * PT_BOX + (D / 64) * 4096 + 2 * 4096 + 2 * 2 * PT_BOX + 32 * 4 + 16 * 8 + 16 + 1024;
}

int attnpool_tc_fwd(const void* x, int dtype, const unsigned char* mask, long long mb, const float* qt, const void* qt_img,
                    int B, int N, int D, int H, int S, float* part_m, float* part_l, float* part_acc, float drop_p,
                    unsigned long long drop_seed, float* part_l2, cudaStream_t s) {
  if (!x || (!qt && !qt_img) || (reinterpret_cast<uintptr_t>(qt_img) & 15) || !part_m || !part_l || !part_acc || S < 1 || S > (N + PT_TT - 1) / PT_TT) return B2_EINVAL;
  CUtensorMap tmx;
  if (int rc = make_tmap_x3d(&tmx, x, B, N, D)) return rc;
  PtFwdParams p{mask, mb, qt, qt_img, part_m, part_l, part_l2, part_acc, B, N, D, H, S, dtype == 2 ? 1 : 0, drop_p, drop_seed};
  dim3 grid(B, S);
  const int ns = pt_fwd_smem(D, 4) <= 227 * 1024 ? 4 : 3;
  const size_t smem = pt_fwd_smem(D, ns);
  auto k = ns == 4 ? pool_fwd_tc_kernel<4> : pool_fwd_tc_kernel<3>;
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return B2_ECUDA;
  k<<<grid, PT_THREADS, smem, s>>>(tmx, p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

int attnpool_tc_bwd(const void* x, int dtype, const unsigned char* mask, long long mb, const float* qt, const float* dxbar,
                    const float* xbar, const void* w_img, const float* cdot, const float* m, const float* l, int B, int N,
                    int D, int H, int S, void* dx, const float* sa, const float* dsa, float drop_p,
                    unsigned long long drop_seed, const float* dlse, float* part_dq, cudaStream_t s) {
  if (!x || (!w_img && (!qt || !dxbar || !xbar)) || (w_img && !cdot) || (reinterpret_cast<uintptr_t>(w_img) & 15) || !m || !l || !dx || S < 1 || S > (N + PT_TT - 1) / PT_TT) return B2_EINVAL;
  CUtensorMap tmx;
  if (int rc = make_tmap_x3d(&tmx, x, B, N, D)) return rc;
  CUtensorMap tmdx;
  if (int rc = make_tmap_x3d(&tmdx, dx, B, N, D)) return rc;
  PtBwdParams p{mask, mb, qt, dxbar, xbar, w_img, cdot, m, l, dx, part_dq, sa, dsa, dlse, B, N, D, H, S, dtype == 2 ? 1 : 0, drop_p, drop_seed};
  dim3 grid(B, S);
  const int ns = pt_bwd_smem(D, 3) <= 227 * 1024 ? 3 : 2;
  const size_t smem = pt_bwd_smem(D, ns);
  auto k = ns == 3 ? pool_bwd_tc_kernel<3> : pool_bwd_tc_kernel<2>;
  if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return B2_ECUDA;
  k<<<grid, PT_BWD_THREADS, smem, s>>>(tmx, tmdx, p);
  return cudaGetLastError() == cudaSuccess ? B2_OK : B2_ECUDA;
}

}  // namespace b2host
