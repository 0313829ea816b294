// zg_attn.cu -- attention kernels of the batched paths.
//
// (1) attn_prefill_kernel: causal self-attention over whole prompts (what the reference computes one token at a time
//     through CausalSelfAttention.forward, ops.zig:129-173, with the implicit causality of a growing KV cache) as a
//     fused flash-style kernel on the tensor cores.  One CTA per (128-query tile, head, sequence):
//       warp 0      TMA: Q tile once, K/V tiles of the sequence double-buffered (128-byte swizzle)
//       warp 1      tcgen05.mma: S = Q K^T (128x128x64, K-major x K-major) into TMEM, then O_j = P V (128x64x128,
//                   K-major P from shared memory x MN-major V exactly as TMA delivered it)
//       warps 2..5  one query row per thread: tcgen05.ld S -> online softmax in registers (thread-local row max /
//                   sum, no shuffles) -> P as f16 into swizzled shared memory -> running O in registers
//     f16 operands, fp32 accumulation and softmax.  Two CTAs fit per SM so one CTA's softmax overlaps the other's MMAs.
//
// (2) attn_decode_batch_kernel: one new token per sequence against that sequence's fp32 KV cache (ops.zig:249-307,
//     query length 1, no mask) for B sequences at once; one CTA per (head, sequence), each K/V row read exactly once
//     with 128-bit loads straight from the time-major cache (no transposed copies), online softmax across warps.
#include <stdlib.h>

#include "zg_attn.cuh"

namespace zg {

namespace {

constexpr int QT = 128, KT = 128, HD = 64;
constexpr int TILE_BYTES = QT * HD * 2;  // 16 KB: 128 rows x 128 bytes
constexpr int ATT_THREADS = 192;
constexpr int ATT_SMEM = TILE_BYTES * (1 + 2 + 2) + 2 * TILE_BYTES + 128;  // Q, K x2, V x2, P (2 atoms), barriers
constexpr uint32_t ATT_TMEM_COLS = 256;                                   // S: 128 columns, PV: 64 columns

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for x <= 0 on the FMA/ALU pipes: round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-4 Taylor of 2^f
// (relative error < 5e-5, below fp16 resolution of P), exponent add by integer arithmetic.  The softmax needs
// 128 x 128 exponentials per key tile and MUFU.EX2 issues only 16 per clock per SM (ncu: XU pipe 89 % busy with
// MUFU alone), so every other element takes this path and the two pipes work in parallel.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 9.6181291e-3f, 5.5504109e-2f);
  p = fmaf(p, f, 2.4022651e-1f);
  p = fmaf(p, f, 6.9314718e-1f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

__device__ __forceinline__ float max3(float a, float b, float c) {  // three-input maximum, one instruction on sm_100
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Row maximum of one 128-key score tile (this thread's query row r lives in TMEM lane r).  DIAG: the tile crosses the
// causal boundary and keys c > r are masked; every other tile runs the mask-free instantiation (the per-element
// compare / select pairs were a third of the kernel's instructions when the mask was evaluated for all tiles).
// (Measured and rejected: 16-column tcgen05.ld kept one load ahead of the arithmetic -- 2 % slower end to end.)
template <bool DIAG>
__device__ __forceinline__ float tile_row_max(uint32_t ts_row, int r) {
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < KT; c += 32) {
    uint32_t sr[32];
    tc::tmem_ld32(ts_row + c, sr);
    tc::tmem_ld_wait();
    if (DIAG) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c + i <= r) mx = fmaxf(mx, __uint_as_float(sr[i]));
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 2) mx = max3(mx, __uint_as_float(sr[i]), __uint_as_float(sr[i + 1]));  // FMNMX3
    }
  }
  return mx;
}

// P = 2^(S scale2 - m_new) for one row of the tile, as f16 into the swizzled K-major P buffer; returns the row sum.
template <bool DIAG>
__device__ __forceinline__ float tile_row_exp(uint32_t ts_row, int r, float scale2, float m_new, uint32_t sP) {
  float sum = 0.0f;
#pragma unroll 1
  for (int c = 0; c < KT; c += 32) {
    uint32_t sr[32];
    tc::tmem_ld32(ts_row + c, sr);
    tc::tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      float p0 = fast_exp2(fmaf(__uint_as_float(sr[i]), scale2, -m_new));
      float p1 = exp2_poly(fmaf(__uint_as_float(sr[i + 1]), scale2, -m_new));
      if (DIAG && c + i > r) p0 = 0.0f;
      if (DIAG && c + i + 1 > r) p1 = 0.0f;
      sum += p0 + p1;
      __half2 t = __floats2half2_rn(p0, p1);
      pk[i >> 1] = *reinterpret_cast<uint32_t *>(&t);
    }
    // K-major SWIZZLE_128B: atom = 64 keys; row r at r * 128 bytes; 16-byte chunk index XOR (r % 8)
    const uint32_t atom = sP + (c >> 6) * TILE_BYTES + r * 128;
    const int ch0 = (c & 63) >> 3;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t a = atom + ((uint32_t)((ch0 + q) ^ (r & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                   "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3])
                   : "memory");
    }
  }
  return sum;
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_prefill_kernel(const __grid_constant__ CUtensorMap tm_qkv, __half *__restrict__ out, int T, int H, int E,
                    int n_bh, unsigned *err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = tc::smem_addr(smem_raw);
  const uint32_t sQ = base, sK = base + TILE_BYTES, sV = sK + 2 * TILE_BYTES, sP = sV + 2 * TILE_BYTES;
  const uint32_t bars = sP + 2 * TILE_BYTES;
  const uint32_t q_full = bars, kv_full = bars + 8, kv_empty = bars + 24, s_full = bars + 40, p_full = bars + 48,
                 pv_full = bars + 56, slot = bars + 64, abort_flag = bars + 68;
  const tc::Guard guard{err, abort_flag};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // heavy (late) query tiles first: tile qt attends to qt + 1 key tiles
  const int num_qt = (T + QT - 1) / QT;
  const int qt = num_qt - 1 - (int)(blockIdx.x / n_bh);
  const int bh = blockIdx.x % n_bh, b = bh / H, h = bh % H;
  const int q0 = qt * QT, n_kv = qt + 1;
  const int row_base = b * T;

  if (threadIdx.x == 0) {
    if (base & 1023u) atomicExch(err, 4u);  // swizzled tiles need a 1024-byte aligned base
    tc::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(kv_full + 8 * s, 1);
      tc::mbar_init(kv_empty + 8 * s, 1);
    }
    tc::mbar_init(s_full, 1);
    tc::mbar_init(p_full, 4);
    tc::mbar_init(pv_full, 1);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(abort_flag), "r"(0u));
    tc::fence_mbar_init();
    tc::prefetch_tmap(&tm_qkv);
  }
  if (warp == 1) tc::tmem_alloc<ATT_TMEM_COLS>(slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<uint32_t *>(smem_raw + (slot - base));
  const uint32_t tS = tmem, tO = tmem + 128;

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer ----------------
      tc::mbar_expect_tx(q_full, TILE_BYTES);
      tc::tma_load_2d(sQ, &tm_qkv, h * HD, row_base + q0, q_full);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        if (!tc::mbar_wait(kv_empty + 8 * s, ((j >> 1) & 1) ^ 1, guard)) break;
        tc::mbar_expect_tx(kv_full + 8 * s, 2 * TILE_BYTES);
        tc::tma_load_2d(sK + s * TILE_BYTES, &tm_qkv, E + h * HD, row_base + j * KT, kv_full + 8 * s);
        tc::tma_load_2d(sV + s * TILE_BYTES, &tm_qkv, 2 * E + h * HD, row_base + j * KT, kv_full + 8 * s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc_s = tc::umma_idesc(0, QT, KT, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = tc::umma_idesc(0, QT, HD, 0, 1);  // P (K-major) x V (MN-major: head dim contiguous)
      bool ok = tc::mbar_wait(q_full, 0, guard);
      for (int j = 0; j < n_kv && ok; ++j) {
        const int s = j & 1;
        if (!tc::mbar_wait(kv_full + 8 * s, (j >> 1) & 1, guard)) break;
        tc::fence_after_sync();
        const uint32_t k_addr = sK + s * TILE_BYTES, v_addr = sV + s * TILE_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::umma<false>(tS, tc::umma_desc_sw128(sQ + 32 * k, 16, 1024), tc::umma_desc_sw128(k_addr + 32 * k, 16, 1024),
                          idesc_s, (uint32_t)(k != 0));
        tc::umma_commit(s_full);
        if (!tc::mbar_wait(p_full, j & 1, guard)) break;
        tc::fence_after_sync();
#pragma unroll
        for (int i = 0; i < 8; ++i)  // 16 keys per MMA: P atom i/4, 32-byte step i%4; V rows 16 i .. 16 i + 15
          tc::umma<false>(tO, tc::umma_desc_sw128(sP + (i >> 2) * TILE_BYTES + 32 * (i & 3), 16, 1024),
                          tc::umma_desc_sw128(v_addr + 2048 * i, 1024, 1024), idesc_o, (uint32_t)(i != 0));
        tc::umma_commit(pv_full);
        tc::umma_commit(kv_empty + 8 * s);
      }
    }
  } else {  // ---------------- softmax / output warps: one query row per thread ----------------
    const int quad = warp & 3, r = quad * 32 + lane;  // TMEM lane == tile row
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const float scale2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    float m = -INFINITY, l = 0.0f, alpha_prev = 0.0f;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.0f;
    bool ok = true;
    for (int j = 0; j < n_kv; ++j) {
      if (!tc::mbar_wait(s_full, j & 1, guard)) { ok = false; break; }
      tc::fence_after_sync();
      const bool diag = (j == qt);  // only the last key tile crosses the causal boundary
      const float mx = diag ? tile_row_max<true>(tS + lane_base, r) : tile_row_max<false>(tS + lane_base, r);
      const float m_new = fmaxf(m, mx * scale2);
      const float alpha = fast_exp2(m - m_new);
      if (j > 0) {  // fold the previous tile's P V (it was computed against the previous running maximum)
        if (!tc::mbar_wait(pv_full, (j - 1) & 1, guard)) { ok = false; break; }
        tc::fence_after_sync();
#pragma unroll
        for (int c = 0; c < HD; c += 32) {
          uint32_t pr[32];
          tc::tmem_ld32(tO + lane_base + c, pr);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c + i] = fmaf(o[c + i], alpha_prev, __uint_as_float(pr[i]));
        }
      }
      alpha_prev = alpha;
      m = m_new;
      const float sum = diag ? tile_row_exp<true>(tS + lane_base, r, scale2, m_new, sP)
                             : tile_row_exp<false>(tS + lane_base, r, scale2, m_new, sP);
      l = fmaf(l, alpha, sum);
      tc::fence_proxy_async_smem();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(p_full);
    }
    if (ok && tc::mbar_wait(pv_full, (n_kv - 1) & 1, guard)) {
      tc::fence_after_sync();
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        uint32_t pr[32];
        tc::tmem_ld32(tO + lane_base + c, pr);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c + i] = fmaf(o[c + i], alpha_prev, __uint_as_float(pr[i]));
      }
      if (q0 + r < T) {
        const float inv = 1.0f / l;
        __half *dst = out + (size_t)(row_base + q0 + r) * E + h * HD;
#pragma unroll
        for (int d = 0; d < HD; d += 8) {
          uint4 pk;
          __half2 t0 = __floats2half2_rn(o[d] * inv, o[d + 1] * inv),
                         t1 = __floats2half2_rn(o[d + 2] * inv, o[d + 3] * inv),
                         t2 = __floats2half2_rn(o[d + 4] * inv, o[d + 5] * inv),
                         t3 = __floats2half2_rn(o[d + 6] * inv, o[d + 7] * inv);
          pk.x = *reinterpret_cast<uint32_t *>(&t0); pk.y = *reinterpret_cast<uint32_t *>(&t1);
          pk.z = *reinterpret_cast<uint32_t *>(&t2); pk.w = *reinterpret_cast<uint32_t *>(&t3);
          *reinterpret_cast<uint4 *>(dst + d) = pk;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (warp == 1) tc::tmem_dealloc<ATT_TMEM_COLS>(tmem);
}

// ---- version 2 of the prefill kernel (round 2): P and O never leave tensor memory, two threads per query row ---------------
//   * P = 2^(S scale - m) goes back to TMEM as packed f16 (tcgen05.st, columns [192, 256)) and the P V product reads its A
//     operand straight from TMEM (tcgen05.mma, A-from-TMEM form): no shared-memory round trip, no swizzle arithmetic, no
//     generic->async proxy fence, 32 KB less shared memory per CTA;
//   * O accumulates in TMEM across the key tiles (P V with accumulate) instead of being pulled into registers and
//     folded with 64 FMAs per thread and tile.  The running maximum is LAZY: a row keeps scaling by the maximum it
//     last committed to until a tile exceeds it by more than 2^8 (P then still fits f16 with room to spare); only then
//     is O rescaled in TMEM (ld, multiply, st), for the whole warp at once because tcgen05.ld / st are warp-collective.
//     l and O are always scaled by the same reference, so the final O / l is exact;
//   * EIGHT softmax warps: the two threads (warp w, warp w + 4, same lane) share a query row and split its 128 key
//     columns, 64 each (and the 64 output columns, 32 each); the row maximum of a tile is exchanged through shared
//     memory (double-buffered by tile parity, one named barrier per tile).  With 83 registers per thread two such CTAs
//     (20 warps) fit an SM: five warps per scheduler hide the tcgen05.ld and MUFU latencies that four could not.
//   tcgen05.commit semantics order everything: s_full(j) arrives only after every earlier MMA of the issuing thread --
//   including P V of tile j-1 -- has completed, so a softmax thread that has seen s_full(j) may rescale O and overwrite P.
constexpr int ATT2_THREADS = (2 + 8) * 32;
constexpr int ATT2_SMEM = TILE_BYTES * (1 + 2 + 2) + 128 + 2 * 2 * QT * 4;  // Q, K x2, V x2, barriers, row maxima [parity][half][row]
#ifndef ZG_ATTN_POLY_EVERY
#define ZG_ATTN_POLY_EVERY 0  // 2: every other element takes the FMA-pipe polynomial (round 1), 4 / 8: every fourth / eighth,
                              // 0: MUFU.EX2 only -- measured 112.5 / 104.7 / 101.3 / 99.3 us per layer at 355M, 16 x 1024
#endif
constexpr float LAZY_RESCALE = 8.0f;                                         // log2 domain

// 64 key columns [c0, c0 + 64) of this thread's row: running maximum
template <bool DIAG>
__device__ __forceinline__ float half_row_max(uint32_t ts_row, int r, int c0) {
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = c0; c < c0 + 64; c += 32) {
    uint32_t sr[32];
    tc::tmem_ld32(ts_row + c, sr);
    tc::tmem_ld_wait();
    if (DIAG) {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (c + i <= r) mx = fmaxf(mx, __uint_as_float(sr[i]));
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 2) mx = max3(mx, __uint_as_float(sr[i]), __uint_as_float(sr[i + 1]));  // FMNMX3
    }
  }
  return mx;
}
// the same 64 columns: P as packed f16 into TMEM columns tp_row + c/2 .., returns the partial row sum
template <bool DIAG>
__device__ __forceinline__ float half_row_exp(uint32_t ts_row, uint32_t tp_row, int r, int c0, float scale2, float m_ref,
                                              float &tile_max) {
  float sum = 0.0f, mx = -INFINITY;
#pragma unroll 1
  for (int c = c0; c < c0 + 64; c += 32) {
    uint32_t sr[32];
    tc::tmem_ld32(ts_row + c, sr);
    tc::tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      // one element in ZG_ATTN_POLY_EVERY takes the FMA-pipe polynomial, the others MUFU.EX2: with eight softmax warps
      // per CTA the kernel is bound by instruction issue (ncu: issue slots 54 % busy, XU 18 %), and the polynomial costs
      // ~10 instructions against 2
      const float x0 = fmaf(__uint_as_float(sr[i]), scale2, -m_ref);
      const float x1 = fmaf(__uint_as_float(sr[i + 1]), scale2, -m_ref);
      float p0 = fast_exp2(x0);
      float p1 = (ZG_ATTN_POLY_EVERY > 0 && ((i >> 1) % (ZG_ATTN_POLY_EVERY / 2 > 0 ? ZG_ATTN_POLY_EVERY / 2 : 1)) == 0) ? exp2_poly(x1) : fast_exp2(x1);
      if (DIAG) {
        if (c + i > r) p0 = 0.0f; else mx = fmaxf(mx, x0);
        if (c + i + 1 > r) p1 = 0.0f; else mx = fmaxf(mx, x1);
      } else {
        mx = max3(mx, x0, x1);  // how far this tile rises above the reference (log2 domain)
      }
      sum += p0 + p1;
      __half2 t = __floats2half2_rn(p0, p1);
      pk[i >> 1] = *reinterpret_cast<uint32_t *>(&t);
    }
    tc::tmem_st16(tp_row + (c >> 1), pk);  // keys c .. c+31 -> 16 packed columns
  }
  tc::tmem_st_wait();
  tile_max = mx;
  return sum;
}

__global__ void __launch_bounds__(ATT2_THREADS, 2)
attn_prefill_kernel_v2(const __grid_constant__ CUtensorMap tm_qkv, __half *__restrict__ out, int T, int H, int E,
                       int n_bh, unsigned *err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = tc::smem_addr(smem_raw);
  const uint32_t sQ = base, sK = base + TILE_BYTES, sV = sK + 2 * TILE_BYTES;
  const uint32_t bars = sV + 2 * TILE_BYTES;
  const uint32_t q_full = bars, kv_full = bars + 8, kv_empty = bars + 24, s_full = bars + 40, p_full = bars + 48,
                 pv_full = bars + 56, slot = bars + 64, abort_flag = bars + 68;
  float *rowx = reinterpret_cast<float *>(smem_raw + (bars + 128 - base));  // [2 parity][2 half][128 rows]
  const tc::Guard guard{err, abort_flag};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int num_qt = (T + QT - 1) / QT;
  const int qt = num_qt - 1 - (int)(blockIdx.x / n_bh);  // heavy (late) query tiles first
  const int bh = blockIdx.x % n_bh, b = bh / H, h = bh % H;
  const int q0 = qt * QT, n_kv = qt + 1;
  const int row_base = b * T;

  if (threadIdx.x == 0) {
    if (base & 1023u) atomicExch(err, 4u);
    tc::mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(kv_full + 8 * s, 1);
      tc::mbar_init(kv_empty + 8 * s, 1);
    }
    tc::mbar_init(s_full, 1);
    tc::mbar_init(p_full, 8);
    tc::mbar_init(pv_full, 1);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(abort_flag), "r"(0u));
    tc::fence_mbar_init();
    tc::prefetch_tmap(&tm_qkv);
  }
  if (warp == 1) tc::tmem_alloc<ATT_TMEM_COLS>(slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *reinterpret_cast<uint32_t *>(smem_raw + (slot - base));
  const uint32_t tS = tmem, tO = tmem + 128, tP = tmem + 192;

  // warps 0 and 1 run their loops with all 32 lanes and warp-uniform operands; one elected lane issues (zg_tc.cuh, *_u)
  if (warp == 0) {  // ---------------- TMA producer ----------------
    tc::mbar_expect_tx_u(q_full, TILE_BYTES);
    tc::tma_load_2d_u(sQ, &tm_qkv, h * HD, row_base + q0, q_full);
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      if (!tc::mbar_wait_u(kv_empty + 8 * s, ((j >> 1) & 1) ^ 1, guard)) break;
      tc::mbar_expect_tx_u(kv_full + 8 * s, 2 * TILE_BYTES);
      tc::tma_load_2d_u(sK + s * TILE_BYTES, &tm_qkv, E + h * HD, row_base + j * KT, kv_full + 8 * s);
      tc::tma_load_2d_u(sV + s * TILE_BYTES, &tm_qkv, 2 * E + h * HD, row_base + j * KT, kv_full + 8 * s);
    }
  } else if (warp == 1) {  // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc_s = tc::umma_idesc(0, QT, KT, 0, 0);  // Q (K-major) x K (K-major)
    constexpr uint32_t idesc_o = tc::umma_idesc(0, QT, HD, 0, 1);  // P (TMEM, K-major) x V (MN-major: head dim contiguous)
    // descriptors differ only in the 14-bit start-address field (bits 0..13, units of 16 bytes): built once, then an add
    // on the low word per MMA (shared memory ends below 256 KB, so the field never carries)
    const uint64_t dq0 = tc::umma_desc_sw128(sQ, 16, 1024), dk0 = tc::umma_desc_sw128(sK, 16, 1024),
                   dv0 = tc::umma_desc_sw128(sV, 1024, 1024);
    bool ok = tc::mbar_wait_u(q_full, 0, guard);
    for (int j = 0; j < n_kv && ok; ++j) {
      const int s = j & 1;
      if (!tc::mbar_wait_u(kv_full + 8 * s, (j >> 1) & 1, guard)) break;
      tc::fence_after_sync();
      const uint64_t dk = dk0 + (uint64_t)(s * (TILE_BYTES >> 4)), dv = dv0 + (uint64_t)(s * (TILE_BYTES >> 4));
#pragma unroll
      for (int k = 0; k < 4; ++k) tc::umma_u<false>(tS, dq0 + 2 * k, dk + 2 * k, idesc_s, (uint32_t)(k != 0));
      tc::umma_commit_u(s_full);
      if (!tc::mbar_wait_u(p_full, j & 1, guard)) break;
      tc::fence_after_sync();
#pragma unroll
      for (int i = 0; i < 8; ++i)  // 16 keys per MMA = 8 packed columns of P; V rows 16 i .. 16 i + 15
        tc::umma_ts_f16_u(tO, tP + 8 * i, dv + 128 * i, idesc_o, (uint32_t)((j | i) != 0));
      tc::umma_commit_u(kv_empty + 8 * s);
      if (j == n_kv - 1) tc::umma_commit_u(pv_full);
    }
  } else {  // ---------------- softmax warps: two threads per query row ----------------
    const int quad = warp & 3, r = quad * 32 + lane;  // TMEM lane == tile row (tcgen05.ld: warp w touches lanes 32 (w % 4) ..)
    const int hf = (warp - 2) >> 2;                   // which 64 key columns / 32 output columns of the row
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const float scale2 = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
    // S is read from TMEM ONCE per tile: tcgen05.ld moves 64 bytes per clock and SM, so the classic two passes (row
    // maximum, then exponentials) cost 2 x 64 KB = 2,048 cycles per 128 x 128 tile -- more than everything else together.
    // The exponentials are taken against the reference maximum the row already has and the tile's own maximum is
    // tracked on the side; the reference moves (and O is rescaled) BEFORE THE NEXT tile when the tile rose more than
    // 2^8 above it, and only a rise above 2^15 (P would leave the f16 range) makes the row pair redo the tile at once.
    // The very first tile has no reference yet and takes the two passes.
    float m_ref = -INFINITY, l = 0.0f, m_next = 0.0f;
    bool pending = false;  // the reference moves to m_next before the next tile
    bool ok = true;
    for (int j = 0; j < n_kv; ++j) {
      if (!tc::mbar_wait(s_full, j & 1, guard)) { ok = false; break; }
      tc::fence_after_sync();
      const bool diag = (j == qt);  // only the last key tile crosses the causal boundary
      float *xb = rowx + (j & 1) * 2 * QT;
      auto rescale_o = [&](float alpha) {  // warp-collective: this warp's 32 rows x this half's 32 columns of O
        uint32_t pr[32];
        tc::tmem_ld32(tO + lane_base + 32 * hf, pr);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) pr[i] = __float_as_uint(__uint_as_float(pr[i]) * alpha);
        tc::tmem_st32(tO + lane_base + 32 * hf, pr);
        tc::tmem_st_wait();
      };
      if (j == 0) {
        float mx = (diag ? half_row_max<true>(tS + lane_base, r, 64 * hf) : half_row_max<false>(tS + lane_base, r, 64 * hf)) * scale2;
        xb[hf * QT + r] = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");  // the two warps that share these 32 rows
        m_ref = fmaxf(mx, xb[(hf ^ 1) * QT + r]);
      } else if (__any_sync(0xffffffffu, pending)) {
        const float alpha = pending ? fast_exp2(m_ref - m_next) : 1.0f;
        rescale_o(alpha);
        l *= alpha;
        if (pending) m_ref = m_next;
        pending = false;
      }
      float tmax;
      float sum = diag ? half_row_exp<true>(tS + lane_base, tP + lane_base, r, 64 * hf, scale2, m_ref, tmax)
                       : half_row_exp<false>(tS + lane_base, tP + lane_base, r, 64 * hf, scale2, m_ref, tmax);
      if (j > 0) {
        xb[hf * QT + r] = tmax;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        const float rise = fmaxf(tmax, xb[(hf ^ 1) * QT + r]);  // of the whole row, relative to m_ref
        const bool redo = rise > 15.0f;
        if (__any_sync(0xffffffffu, redo)) {  // P would overflow f16: move the reference now and recompute (rare)
          const float alpha = redo ? fast_exp2(-rise) : 1.0f;
          rescale_o(alpha);
          l *= alpha;
          if (redo) m_ref += rise;
          sum = diag ? half_row_exp<true>(tS + lane_base, tP + lane_base, r, 64 * hf, scale2, m_ref, tmax)
                     : half_row_exp<false>(tS + lane_base, tP + lane_base, r, 64 * hf, scale2, m_ref, tmax);
        } else if (rise > LAZY_RESCALE) {
          pending = true;
          m_next = m_ref + rise;
        }
      }
      l += sum;
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(p_full);
    }
    // total row sum = the two halves' partial sums (same reference maximum)
    float *xb = rowx + ((n_kv & 1) * 2 * QT);
    xb[hf * QT + r] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
    l += xb[(hf ^ 1) * QT + r];
    if (ok && tc::mbar_wait(pv_full, 0, guard)) {
      tc::fence_after_sync();
      const float inv = 1.0f / l;
      __half *dst = out + (size_t)(row_base + q0 + r) * E + h * HD + 32 * hf;
      uint32_t pr[32];
      tc::tmem_ld32(tO + lane_base + 32 * hf, pr);
      tc::tmem_ld_wait();
      if (q0 + r < T) {
#pragma unroll
        for (int d = 0; d < 32; d += 8) {
          uint4 pk;
          __half2 t0 = __floats2half2_rn(__uint_as_float(pr[d]) * inv, __uint_as_float(pr[d + 1]) * inv),
                         t1 = __floats2half2_rn(__uint_as_float(pr[d + 2]) * inv, __uint_as_float(pr[d + 3]) * inv),
                         t2 = __floats2half2_rn(__uint_as_float(pr[d + 4]) * inv, __uint_as_float(pr[d + 5]) * inv),
                         t3 = __floats2half2_rn(__uint_as_float(pr[d + 6]) * inv, __uint_as_float(pr[d + 7]) * inv);
          pk.x = *reinterpret_cast<uint32_t *>(&t0); pk.y = *reinterpret_cast<uint32_t *>(&t1);
          pk.z = *reinterpret_cast<uint32_t *>(&t2); pk.w = *reinterpret_cast<uint32_t *>(&t3);
          *reinterpret_cast<uint4 *>(dst + d) = pk;
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (warp == 1) tc::tmem_dealloc<ATT_TMEM_COLS>(tmem);
}

// -------------------------------------------------------------------------------------------------------------------
// Batched single-query attention off the fp32 caches.  cache layout: [B][C][E] (time-major, heads interleaved -- the
// reference's per-block cache, main.zig:298-299, once per sequence).  q: [B, ldq] (row b holds q at columns h*64..).
// T = *pos_dev + pos_base + 1 rows are attended.  knew == null: the new token's K/V row is already in the cache (the
// general GEMM's epilogue appended it).  knew != null: row T-1 is taken from knew / vnew (row b, columns h*64.., pitch
// ldq: the c_attn output, complete only at this kernel boundary because the stream-K GEMM reduces partial sums) and this
// CTA appends it to the caches (ops.zig:151-152,156-157) -- read through plain loads, never through the cache, because
// the non-coherent streaming loads below must not see data written by this same kernel.
// -------------------------------------------------------------------------------------------------------------------
constexpr int DEC_WARPS = 8;

// four consecutive cache elements: the caches are fp32 (the reference's storage) or f16 (16-bit KV storage).  Loaded raw
// (rows in flight stay in their storage format: an f16 row costs half the registers) and widened at the point of use.
template <typename CT> struct Raw4;
template <> struct Raw4<float> { typedef float4 type; };
template <> struct Raw4<__half> { typedef uint2 type; };
__device__ __forceinline__ float4 ld_raw4(const float *p) { return ld_stream(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ uint2 ld_raw4(const __half *p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 widen(float4 v) { return v; }
__device__ __forceinline__ float4 widen(uint2 r) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2 *>(&r.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void narrow(float4 v, float4 *o) { *o = v; }
__device__ __forceinline__ void narrow(float4 v, uint2 *o) {
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  o->x = *reinterpret_cast<uint32_t *>(&a);
  o->y = *reinterpret_cast<uint32_t *>(&b);
}
__device__ __forceinline__ void st_out(float *p, float v) { *p = v; }
__device__ __forceinline__ void st_out(__half *p, float v) { *p = __float2half_rn(v); }

// CT: cache element type; OT: output element type (f16 when the c_proj GEMM that follows reads f16 operands)
template <typename CT, typename OT>
__global__ void __launch_bounds__(DEC_WARPS * 32)
attn_decode_batch_kernel(const float *q, int ldq, const CT *k_cache, const CT *v_cache, long long seq_stride, int E, OT *out, int ldo,
                         const int *pos_dev, int pos_base, const float *knew, const float *vnew, int trigger,
                         int rows_per_seq) {
  __shared__ float s_m[DEC_WARPS], s_l[DEC_WARPS];
  __shared__ float s_o[DEC_WARPS][HD];
  // rows_per_seq == 0: decode step, row blockIdx.y = sequence b at the common position.  rows_per_seq = T_prompt > 0:
  // fp32-class causal prefill -- row blockIdx.y = (b, t) attends to cache rows [0, t] of sequence b, which is what t + 1
  // calls of CausalSelfAttention.forward leave in _q (ops.zig:129-173; tests.zig:245-334 proves the equivalence).
  const int h = blockIdx.x, qrow = blockIdx.y;
  const int b = rows_per_seq ? qrow / rows_per_seq : qrow;
  if (trigger) pdl_trigger();  // the c_proj GEMM that follows fills its ring with weights while the caches stream
  pdl_wait();  // q / the new K,V row are the c_attn GEMM's output (no-op for an ordinary launch)
  const int T = rows_per_seq ? (qrow % rows_per_seq) + 1 : pos_base + (pos_dev ? *pos_dev : 0) + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;  // a half-warp covers one 256-byte head row with float4 loads
  const float4 qv = *reinterpret_cast<const float4 *>(q + (size_t)qrow * ldq + h * HD + 4 * l16);
  const CT *kb = k_cache + (size_t)b * seq_stride + h * HD + 4 * l16;
  const CT *vb = v_cache + (size_t)b * seq_stride + h * HD + 4 * l16;
  typedef typename Raw4<CT>::type RawT;
  RawT k_last, v_last;  // row T-1 in storage format, from the c_attn output (knew != null)
  narrow(make_float4(0.f, 0.f, 0.f, 0.f), &k_last);
  v_last = k_last;
  if (knew) {
    narrow(*reinterpret_cast<const float4 *>(knew + (size_t)b * ldq + h * HD + 4 * l16), &k_last);
    narrow(*reinterpret_cast<const float4 *>(vnew + (size_t)b * ldq + h * HD + 4 * l16), &v_last);
    if (threadIdx.x < 16) {  // cache append: 16 lanes x 4 elements = this head's 64 values of row T-1
      *reinterpret_cast<RawT *>(const_cast<CT *>(kb) + (size_t)(T - 1) * E) = k_last;
      *reinterpret_cast<RawT *>(const_cast<CT *>(vb) + (size_t)(T - 1) * E) = v_last;
    }
  }
  const float scale = 0.125f;
  float m = -INFINITY, l = 0.0f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // each half-warp walks rows t = 2 * (warp + DEC_WARPS * i) + half; U rows in flight per half-warp -- 4 KB per warp
  // whatever the storage width (an f16 row is half as long, so twice as many are kept in flight)
  constexpr int U = sizeof(CT) == 2 ? 8 : 4;
  constexpr int STEP = 2 * DEC_WARPS;
  for (int tb = 2 * warp; tb < T; tb += U * STEP) {  // warp-uniform trip count: the shuffles below need all 32 lanes
    const int t0 = tb + half;
    RawT kk[U], vv[U];
    float s[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u * STEP;
      kk[u] = k_last;
      vv[u] = v_last;
      if (t < T && !(knew && t == T - 1)) {
        kk[u] = ld_raw4(kb + (size_t)t * E);
        vv[u] = ld_raw4(vb + (size_t)t * E);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u * STEP;
      const float4 kf = widen(kk[u]);
      float d = (t < T) ? (qv.x * kf.x + qv.y * kf.y + qv.z * kf.z + qv.w * kf.w) : 0.0f;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      s[u] = (t < T) ? d * scale : -INFINITY;
    }
    float mx = s[0];
#pragma unroll
    for (int u = 1; u < U; ++u) mx = fmaxf(mx, s[u]);
    const float m_new = fmaxf(m, mx);
    if (m_new > -INFINITY) {
      const float a = __expf(m - m_new);
      acc.x *= a; acc.y *= a; acc.z *= a; acc.w *= a;
      l *= a;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (s[u] > -INFINITY) {
          const float p = __expf(s[u] - m_new);
          const float4 vf = widen(vv[u]);
          l += p;
          acc.x = fmaf(p, vf.x, acc.x); acc.y = fmaf(p, vf.y, acc.y);
          acc.z = fmaf(p, vf.z, acc.z); acc.w = fmaf(p, vf.w, acc.w);
        }
      }
      m = m_new;
    }
  }
  // merge the two half-warps, then the warps
  {
    const float m_o = __shfl_xor_sync(0xffffffffu, m, 16), l_o = __shfl_xor_sync(0xffffffffu, l, 16);
    float4 a_o;
    a_o.x = __shfl_xor_sync(0xffffffffu, acc.x, 16); a_o.y = __shfl_xor_sync(0xffffffffu, acc.y, 16);
    a_o.z = __shfl_xor_sync(0xffffffffu, acc.z, 16); a_o.w = __shfl_xor_sync(0xffffffffu, acc.w, 16);
    const float mm = fmaxf(m, m_o);
    const float f0 = (m > -INFINITY) ? __expf(m - mm) : 0.0f, f1 = (m_o > -INFINITY) ? __expf(m_o - mm) : 0.0f;
    acc.x = acc.x * f0 + a_o.x * f1; acc.y = acc.y * f0 + a_o.y * f1;
    acc.z = acc.z * f0 + a_o.z * f1; acc.w = acc.w * f0 + a_o.w * f1;
    l = l * f0 + l_o * f1;
    m = mm;
  }
  if (lane < 16) {
    if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
    *reinterpret_cast<float4 *>(&s_o[warp][4 * l16]) = acc;
  }
  __syncthreads();
  if (threadIdx.x < HD) {
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) mm = fmaxf(mm, s_m[w]);
    float num = 0.0f, den = 0.0f;
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) {
      const float f = (s_m[w] > -INFINITY) ? __expf(s_m[w] - mm) : 0.0f;
      num = fmaf(f, s_o[w][threadIdx.x], num);
      den = fmaf(f, s_l[w], den);
    }
    st_out(out + (size_t)qrow * ldo + h * HD + threadIdx.x, num / den);
  }
}

}  // namespace

bool attn_prefill_plan(AttnPrefillPlan *p, const void *qkv_f16, void *out_f16, int B, int T, int H, int E) {
  if (E != H * HD) {
    set_error(1, "prefill attention: head_dim must be 64", __FILE__, __LINE__);
    return false;
  }
  p->out = out_f16;
  p->B = B; p->T = T; p->H = H; p->E = E;
  return make_tmap_2d(&p->tm_qkv, qkv_f16, 1, (uint64_t)B * T, (uint64_t)3 * E, (uint64_t)3 * E * 2, QT, HD);
}

static bool attn_v1() {  // A/B switch: ZG_ATTN_V1=1 selects the round-1 kernel (P through shared memory, O in registers)
  static const bool v1 = getenv("ZG_ATTN_V1") != nullptr;
  return v1;
}

void attn_init_attrs() {
  static unsigned attr_gen = 0;  // per device: redone after every zg_init
  if (attr_gen != ctx().generation) {
    ZG_CUDA(cudaFuncSetAttribute(attn_prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    ZG_CUDA(cudaFuncSetAttribute(attn_prefill_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT2_SMEM));
    attr_gen = ctx().generation;
  }
}

void attn_prefill_launch(const AttnPrefillPlan &p) {
  attn_init_attrs();
  const int num_qt = (p.T + QT - 1) / QT, n_bh = p.B * p.H;
  if (attn_v1())
    attn_prefill_kernel<<<num_qt * n_bh, ATT_THREADS, ATT_SMEM, ctx().stream>>>(
        p.tm_qkv, reinterpret_cast<__half *>(p.out), p.T, p.H, p.E, n_bh, gemm_error_word());
  else
    attn_prefill_kernel_v2<<<num_qt * n_bh, ATT2_THREADS, ATT2_SMEM, ctx().stream>>>(
        p.tm_qkv, reinterpret_cast<__half *>(p.out), p.T, p.H, p.E, n_bh, gemm_error_word());
  ZG_LAUNCH_CHECK();
}

void attn_decode_batch_launch(const float *q, int ldq, const float *k_cache, const float *v_cache, long long seq_stride,
                              int B, int H, int E, float *out, int ldo, const int *pos_dev, int pos_base,
                              const float *knew, const float *vnew, int rows_per_seq) {
  // rows_per_seq > 0 (causal prefill): B * rows_per_seq query rows
  ZG_CUDA(launch_pdl(PDL_ATTN_DEP, attn_decode_batch_kernel<float, float>, dim3(H, rows_per_seq ? B * rows_per_seq : B),
                     dim3(DEC_WARPS * 32), 0, ctx().stream, q, ldq, k_cache, v_cache, seq_stride, E, out, ldo, pos_dev, pos_base, knew,
                     vnew, (knew && (pdl_mask() & PDL_ATTN_TRIGGER)) ? 1 : 0, rows_per_seq));
  ZG_LAUNCH_CHECK();
}

// the same step over f16 caches, f16 output (16-bit KV storage: half the bytes the step has to stream)
void attn_decode_batch_launch_f16(const float *q, int ldq, const void *k_cache, const void *v_cache, long long seq_stride, int B,
                                  int H, int E, void *out_f16, int ldo, const int *pos_dev, const float *knew, const float *vnew) {
  ZG_CUDA(launch_pdl(PDL_ATTN_DEP, attn_decode_batch_kernel<__half, __half>, dim3(H, B), dim3(DEC_WARPS * 32), 0, ctx().stream, q, ldq,
                     (const __half *)k_cache, (const __half *)v_cache, seq_stride, E, (__half *)out_f16, ldo, pos_dev, 0, knew, vnew,
                     (knew && (pdl_mask() & PDL_ATTN_TRIGGER)) ? 1 : 0, 0));
  ZG_LAUNCH_CHECK();
}

}  // namespace zg

using namespace zg;

extern "C" {

// Causal self-attention over B prompts of T tokens: qkv f16 [B*T, 3E] (the c_attn output) -> out f16 [B*T, E].
void zg_attention_prefill(const void *qkv_f16, void *out_f16, size_t B, size_t T, size_t n_heads, size_t n_embed) {
  if (!require_ready("zg_attention_prefill")) return;
  AttnPrefillPlan p;
  if (!attn_prefill_plan(&p, qkv_f16, out_f16, (int)B, (int)T, (int)n_heads, (int)n_embed)) return;
  attn_prefill_launch(p);
}

// One-token attention for B sequences off their fp32 caches ([B][context][E] each): q [B,E] -> out [B,E];
// rows [0, seq_len) of every cache are attended.
void zg_attention_decode_batch(const float *q, const float *k_cache, const float *v_cache, size_t B, size_t context,
                               size_t n_heads, size_t n_embed, size_t seq_len, float *out) {
  if (!require_ready("zg_attention_decode_batch") || seq_len == 0) return;
  attn_decode_batch_launch(q, (int)n_embed, k_cache, v_cache, (long long)(context * n_embed), (int)B, (int)n_heads,
                           (int)n_embed, out, (int)n_embed, nullptr, (int)seq_len - 1, nullptr, nullptr);
}

}  // extern "C"
