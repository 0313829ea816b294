// zg_gemm.cu -- Linear.forward (ops.zig:21-46) for M >= 16 as a hand-written Blackwell GEMM:
//   out[M,N] = epilogue(bias + A[M,K] . W[N,K]^T)
// Both operands are K-major, exactly the reference's `cblas_sgemm(RowMajor, NoTrans, Trans)` call (ops.zig:30-45),
// so the weights stay in the reference's [out, in] layout.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer   cp.async.bulk.tensor (UTMALDG) of 128-byte-swizzled A / W tiles into a shared-memory ring
//   warp 1      MMA issuer     one elected thread issues tcgen05.mma (UTC*MMA), accumulators in TMEM, double-buffered
//   warps 2..9  epilogue       tcgen05.ld (LDTM) -> bias / GELU / residual / K-V cache append -> global
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).  The epilogue of tile i
// overlaps the main loop of tile i+1.  kind::tf32 reads the fp32 tensors as they are; kind::f16 reads f16 copies.
#include <initializer_list>
#include <stdio.h>
#include <stdlib.h>

#include "zg_gemm.cuh"

namespace zg {

namespace {

constexpr int BM = 128;           // UMMA M (cta_group::1): TMEM lane i <-> tile row i
constexpr int ROW_BYTES = 128;    // one swizzle row = BLOCK_K elements
constexpr int A_STAGE = BM * ROW_BYTES;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (2 + EPI_WARPS) * 32;
constexpr int SMEM_BUDGET = 192 * 1024;
constexpr int STAGING = EPI_WARPS * 4096;  // one swizzled 32-row x 128-byte chunk per epilogue warp

enum { MODE_F16 = 0, MODE_TF32 = 1, MODE_TF32X3 = 2 };

template <int MODE, int BN>
struct Cfg {
  // 3xTF32: every stage holds the W tile twice in shared memory (as landed, and its lo part); the A tile's hi and lo
  // parts live in TMEM (A_COLS columns per stage behind the two accumulators), so the MMAs read only W from shared memory
  static constexpr bool SPLIT = MODE == MODE_TF32X3;
  static_assert(!SPLIT || BN <= 192, "3xTF32: the accumulator(s) + the A ring must fit 512 TMEM columns");
  // accumulator stages: two (the epilogue of a tile overlaps the next tile's main loop); the 192-column 3xTF32 tile has TMEM
  // for one only and is used where a CTA gets a single tile anyway (gemm_plan)
  static constexpr int ACC = (SPLIT && BN > 128) ? 1 : 2;
  static constexpr int B_STAGE = BN * ROW_BYTES;
  static constexpr int STAGE = A_STAGE + B_STAGE;  // what TMA writes per stage
  static constexpr int STAGE_ALL = SPLIT ? STAGE + B_STAGE : STAGE;
  static constexpr int A_COLS = 64;  // SPLIT: A_hi in columns 0..31, A_lo in 32..63 (one 32-bit column per K element)
  static constexpr int MAX_STAGES = SPLIT ? (512 - ACC * BN) / A_COLS : 10;
  static constexpr int STAGES = (SMEM_BUDGET / STAGE_ALL) > MAX_STAGES ? MAX_STAGES : (SMEM_BUDGET / STAGE_ALL);
  static constexpr int TMEM_COLS = SPLIT ? 512 : ((2 * BN) < 32 ? 32 : 2 * BN);  // two accumulator stages (+ the A ring)
  // epilogue warps e and e+4 split the columns; in the 3xTF32 mode warps 6..9 split operands instead
  static constexpr int HALVES = (SPLIT || BN < 64) ? 1 : 2;
  static constexpr int COLS_PER_HALF = BN / HALVES;
  static constexpr int SMEM = STAGES * STAGE_ALL + STAGING + 1024 /*alignment slack*/ + 384 /*barriers*/;
};

__device__ __forceinline__ float gelu_fast(float x) {
  float t;
  const float u = x * 0.7978845608f * (1.0f + 0.044715f * x * x);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return 0.5f * x * (1.0f + t);
}

// One 32-column chunk of one accumulator row: v[j] is out[row, col0 + j] before the epilogue.
struct EpiTma {
  const CUtensorMap *out, *k, *v;
  uint32_t stage;  // this warp's 4 KB staging chunk (1024-byte aligned)
  int row0;        // first row of this warp's 32 rows
};

// Stage one 32-row chunk (this thread's row = `lane`) and hand it to the TMA engine.  fp32: 128-byte rows, 16-byte
// chunk c of row r lands at c ^ (r % 8) (SWIZZLE_128B); fp16: 64-byte rows, chunk c at c ^ ((r / 2) % 4) (SWIZZLE_64B).
__device__ __forceinline__ void stage_and_store(const float (&v)[32], bool as_f16, uint32_t stage, int lane,
                                                const CUtensorMap *tm, int c0, int c1, bool reduce) {
  if (lane == 0) tc::tma_wait_read0();  // the previous store from this buffer has been read out
  __syncwarp();
  if (as_f16) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      __half2 t0 = __floats2half2_rn(v[8 * q], v[8 * q + 1]), t1 = __floats2half2_rn(v[8 * q + 2], v[8 * q + 3]),
              t2 = __floats2half2_rn(v[8 * q + 4], v[8 * q + 5]), t3 = __floats2half2_rn(v[8 * q + 6], v[8 * q + 7]);
      const uint32_t a = stage + lane * 64 + ((uint32_t)(q ^ ((lane >> 1) & 3)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(*reinterpret_cast<uint32_t *>(&t0)),
                   "r"(*reinterpret_cast<uint32_t *>(&t1)), "r"(*reinterpret_cast<uint32_t *>(&t2)),
                   "r"(*reinterpret_cast<uint32_t *>(&t3))
                   : "memory");
    }
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t a = stage + lane * 128 + ((uint32_t)(q ^ (lane & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(__float_as_uint(v[4 * q])),
                   "r"(__float_as_uint(v[4 * q + 1])), "r"(__float_as_uint(v[4 * q + 2])), "r"(__float_as_uint(v[4 * q + 3]))
                   : "memory");
    }
  }
  tc::fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    if (reduce) tc::tma_reduce_add_2d(tm, stage, c0, c1);
    else tc::tma_store_2d(tm, stage, c0, c1);
    tc::tma_commit_group();
  }
}

__device__ __forceinline__ void epilogue_chunk(const GemmArgs &g, float (&v)[32], int row, int col0, int pos_now,
                                               const EpiTma &t, int lane, bool add_bias) {
  const bool full = (col0 + 32 <= g.N);
  if (g.bias && add_bias) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4 *>(g.bias + col0 + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < g.N) v[j] += __ldg(g.bias + col0 + j);
    }
  }
  if (g.epi == TC_EPI_GELU) {
    if (g.gelu_fast) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_fast(v[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_ref(v[j]);
    }
  }
  if (g.tma_out) {  // TMA clips rows >= M and columns >= N itself
    stage_and_store(v, g.out_f16 != 0, t.stage, lane, t.out, col0, t.row0, g.tma_reduce != 0);
    if (g.k_cache && col0 >= g.E) {
      const int part = col0 / g.E;
      if (g.tma_kv) {  // 32 consecutive rows of one sequence -> 32 consecutive cache rows
        const int seq = t.row0 / g.rows_per_seq, tt = pos_now + t.row0 % g.rows_per_seq;
        stage_and_store(v, false, t.stage, lane, part == 1 ? t.k : t.v, col0 - part * g.E, seq * g.cache_rows + tt, false);
      } else if (row < g.M) {
        float *cache = (part == 1) ? g.k_cache : g.v_cache;
        const int seq = row / g.rows_per_seq, tt = pos_now + row % g.rows_per_seq;
        float *d = cache + (size_t)seq * g.cache_seq_stride + (size_t)tt * g.E + (col0 - part * g.E);
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(d + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    return;
  }
  if (row >= g.M) return;
  if (g.epi == TC_EPI_RESIDUAL) {
    const float *rr = g.resid + (size_t)row * g.ldr + col0;
    if (full && (g.ldr & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4 *>(rr + j);
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < g.N) v[j] += rr[j];
    }
  }
  if (g.out_f16) {
    __half *o = reinterpret_cast<__half *>(g.out) + (size_t)row * g.ldo + col0;
    if (full && (g.ldo & 7) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        uint4 pk;
        __half2 t0 = __floats2half2_rn(v[j], v[j + 1]), t1 = __floats2half2_rn(v[j + 2], v[j + 3]),
                       t2 = __floats2half2_rn(v[j + 4], v[j + 5]), t3 = __floats2half2_rn(v[j + 6], v[j + 7]);
        pk.x = *reinterpret_cast<uint32_t *>(&t0); pk.y = *reinterpret_cast<uint32_t *>(&t1);
        pk.z = *reinterpret_cast<uint32_t *>(&t2); pk.w = *reinterpret_cast<uint32_t *>(&t3);
        *reinterpret_cast<uint4 *>(o + j) = pk;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < g.N) o[j] = __float2half_rn(v[j]);
    }
  } else {
    float *o = reinterpret_cast<float *>(g.out) + (size_t)row * g.ldo + col0;
    if (full && (g.ldo & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (col0 + j < g.N) o[j] = v[j];
    }
  }
  if (g.k_cache && col0 >= g.E) {  // K / V rows of this token -> the caches (fp32), E % 32 == 0 so a chunk never straddles
    const int part = col0 / g.E;   // 1 = K, 2 = V
    float *cache = (part == 1) ? g.k_cache : g.v_cache;
    const int seq = row / g.rows_per_seq, t = pos_now + row % g.rows_per_seq;
    float *d = cache + (size_t)seq * g.cache_seq_stride + (size_t)t * g.E + (col0 - part * g.E);
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(d + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
}

// MODE_F16: fp16 operands, kind::f16.  MODE_TF32: fp32 operands read in place as kind::tf32 (the tensor core drops the
// low 13 mantissa bits).  MODE_TF32X3: error-compensated fp32 -- warps 6..9 split every landed tile into
// hi = x & 0xffffe000 (what kind::tf32 reads out of the raw word anyway) and lo = x - hi (exact in fp32), and the MMA
// warp accumulates A_lo B_hi + A_hi B_lo + A_hi B_hi, which restores fp32-class accuracy (~2^-20 relative per product)
// on the tensor cores.  Twelve MMAs per k-block re-read their operands from shared memory, which made the mode
// shared-memory-bound (ncu: tensor pipe 30 % busy); the A tile therefore goes to TMEM (tcgen05.st by the splitter
// warps, raw words + lo parts) and the MMAs take A from there: half the operand reads, no A_lo copy in shared memory.
template <int MODE, int BN>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
               const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_k,
               const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ GemmArgs g) {
  using C = Cfg<MODE, BN>;
  constexpr bool TF32 = MODE != MODE_F16;
  constexpr bool SPLIT = C::SPLIT;
  constexpr int BK = TF32 ? 32 : 64;  // elements per 128-byte swizzle row
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_addr(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t sA = base, sB = base + C::STAGES * A_STAGE;
  const uint32_t sBlo = base + C::STAGES * C::STAGE;  // SPLIT only
  const uint32_t staging = base + C::STAGES * C::STAGE_ALL;  // 1024-byte aligned: every stage size is a multiple of 1024
  const uint32_t bars = staging + STAGING;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * C::STAGES, split_bar = bars + 16 * C::STAGES;
  const uint32_t tfull_bar = bars + 24 * C::STAGES, tempty_bar = tfull_bar + 16;
  const uint32_t slot = tempty_bar + 16, abort_flag = slot + 4;
  uint32_t *slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (slot - raw));
  const tc::Guard guard{g.err, abort_flag};

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      tc::mbar_init(full_bar + 8 * s, 1);
      tc::mbar_init(empty_bar + 8 * s, 1);
      tc::mbar_init(split_bar + 8 * s, 4);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(tfull_bar + 8 * a, 1);
      tc::mbar_init(tempty_bar + 8 * a, 4 * C::HALVES);
    }
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(abort_flag), "r"(0u));
    tc::fence_mbar_init();
    tc::prefetch_tmap(&tm_a);
    tc::prefetch_tmap(&tm_b);
    if (g.tma_out) tc::prefetch_tmap(&tm_out);
  }
  if (warp == 1) tc::tmem_alloc<C::TMEM_COLS>(slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot_ptr;

  const int num_m = (g.M + BM - 1) / BM, num_n = (g.N + BN - 1) / BN, tiles = num_m * num_n;
  const int num_kb = (g.K + BK - 1) / BK;
  // work item = (output tile, K slice); slice ks covers k-blocks [ks num_kb / ksplit, (ks + 1) num_kb / ksplit)
  const int ksplit = g.ksplit, items = tiles * ksplit;

  // warps 0 and 1 run their loops with all 32 lanes and warp-uniform operands; one elected lane issues (zg_tc.cuh, *_u)
  if (warp == 0) {  // ---------------- TMA producer ----------------
    uint32_t stage = 0, phase = 0;
    bool ok = true;
    for (int item = blockIdx.x; item < items && ok; item += gridDim.x) {
      const int tile = item % tiles, ks = item / tiles;
      const int m_blk = g.n_fastest ? tile / num_n : tile % num_m, n_blk = g.n_fastest ? tile % num_n : tile / num_m;
      const int kb0 = ks * num_kb / ksplit, kb1 = (ks + 1) * num_kb / ksplit;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!tc::mbar_wait_u(empty_bar + 8 * stage, phase ^ 1, guard)) { ok = false; break; }
        tc::mbar_expect_tx_u(full_bar + 8 * stage, C::STAGE);
        tc::tma_load_2d_u(sA + stage * A_STAGE, &tm_a, kb * BK, m_blk * BM, full_bar + 8 * stage);
        tc::tma_load_2d_u(sB + stage * C::B_STAGE, &tm_b, kb * BK, n_blk * BN, full_bar + 8 * stage);
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {  // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = tc::umma_idesc(TF32 ? 2u : 0u, BM, BN, 0, 0);
    const uint32_t ready_bar = SPLIT ? split_bar : full_bar;
    // descriptors differ only in the 14-bit start-address field (units of 16 bytes): built once, then an add per MMA
    const uint64_t da0 = tc::umma_desc_sw128(sA, 16, 1024), db0 = tc::umma_desc_sw128(sB, 16, 1024);
    const uint64_t db0_lo = tc::umma_desc_sw128(SPLIT ? sBlo : sB, 16, 1024);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    bool ok = true;
    for (int item = blockIdx.x; item < items && ok; item += gridDim.x) {
      const int ks = item / tiles;
      const int kb0 = ks * num_kb / ksplit, kb1 = (ks + 1) * num_kb / ksplit;
      if (!tc::mbar_wait_u(tempty_bar + 8 * acc, acc_phase ^ 1, guard)) break;
      tc::fence_after_sync();
      const uint32_t d = tmem + acc * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!tc::mbar_wait_u(ready_bar + 8 * stage, phase, guard)) { ok = false; break; }
        tc::fence_after_sync();
        const uint64_t oa = (uint64_t)(stage * (A_STAGE >> 4)), ob = (uint64_t)(stage * (C::B_STAGE >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 x 32 bytes of K per swizzle row: UMMA_K = 8 (tf32) / 16 (f16)
          const uint64_t da = da0 + oa + 2 * k, db = db0 + ob + 2 * k;
          if constexpr (SPLIT) {
            const uint64_t db_lo = db0_lo + ob + 2 * k;
            const uint32_t ta = tmem + C::ACC * BN + stage * C::A_COLS + 8 * k;  // 8 K elements per MMA = 8 columns
            tc::umma_ts_u<true>(d, ta + 32, db, idesc, (uint32_t)(((kb - kb0) | k) != 0));  // A_lo W_hi
            tc::umma_ts_u<true>(d, ta, db_lo, idesc, 1u);                                   // A_hi W_lo
            tc::umma_ts_u<true>(d, ta, db, idesc, 1u);                                      // A_hi W_hi
          } else {
            tc::umma_u<TF32>(d, da, db, idesc, (uint32_t)(((kb - kb0) | k) != 0));
          }
        }
        tc::umma_commit_u(empty_bar + 8 * stage);  // frees the smem slot once these MMAs have read it
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      if (ok) tc::umma_commit_u(tfull_bar + 8 * acc);  // accumulator complete -> epilogue
      if (++acc == C::ACC) { acc = 0; acc_phase ^= 1; }
    }
  } else if (SPLIT && warp >= 6) {  // ---------------- operand splitters (3xTF32) ----------------
    const int t = threadIdx.x - 6 * 32;  // 0..127
    const int a_row = (warp & 3) * 32 + lane;  // tcgen05.st: warp w may touch TMEM lanes 32 (w % 4) ..
    const uint32_t a_lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t stage = 0, phase = 0;
    bool ok = true;
    for (int item = blockIdx.x; item < items && ok; item += gridDim.x) {
      const int ks = item / tiles;
      const int kb0 = ks * num_kb / ksplit, kb1 = (ks + 1) * num_kb / ksplit;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!tc::mbar_wait(full_bar + 8 * stage, phase, guard)) { ok = false; break; }
        {  // A: this thread's tile row (TMEM lane) -> 32 raw words + 32 lo parts.  The row's 16-byte chunk c sits at
           // c ^ (row % 8) (SWIZZLE_128B), so the eight rows of a quarter-warp hit eight different bank groups.
          const uint32_t arow = sA + stage * A_STAGE + (uint32_t)a_row * 128u;
          uint32_t x[32], lo[32];
#pragma unroll
          for (int c = 0; c < 8; ++c)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x[4 * c]), "=r"(x[4 * c + 1]), "=r"(x[4 * c + 2]), "=r"(x[4 * c + 3])
                         : "r"(arow + ((uint32_t)(c ^ (a_row & 7)) << 4)));
#pragma unroll
          for (int i = 0; i < 32; ++i) lo[i] = __float_as_uint(__uint_as_float(x[i]) - __uint_as_float(x[i] & 0xffffe000u));
          const uint32_t ta = tmem + C::ACC * BN + a_lane_base + stage * C::A_COLS;
          tc::tmem_st32(ta, x);
          tc::tmem_st32(ta + 32, lo);
        }
        {  // W: lo part next to the landed tile (elementwise, so the swizzled placement is irrelevant)
          const uint32_t hi = sB + stage * C::B_STAGE, lo = sBlo + stage * C::B_STAGE;
#pragma unroll 4
          for (int i = t; i < C::B_STAGE / 16; i += 128) {
            uint32_t x0, x1, x2, x3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(hi + 16 * i));
            const uint32_t l0 = __float_as_uint(__uint_as_float(x0) - __uint_as_float(x0 & 0xffffe000u)),
                           l1 = __float_as_uint(__uint_as_float(x1) - __uint_as_float(x1 & 0xffffe000u)),
                           l2 = __float_as_uint(__uint_as_float(x2) - __uint_as_float(x2 & 0xffffe000u)),
                           l3 = __float_as_uint(__uint_as_float(x3) - __uint_as_float(x3 & 0xffffe000u));
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lo + 16 * i), "r"(l0), "r"(l1), "r"(l2), "r"(l3) : "memory");
          }
        }
        tc::tmem_st_wait();
        tc::fence_before_sync();
        tc::fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(split_bar + 8 * stage);
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {  // ---------------- epilogue warps ----------------
    const int e = warp - 2, quad = warp & 3, half = e >> 2;  // tcgen05.ld: warp w may touch TMEM lanes 32 (w % 4) ..
    if (half < C::HALVES) {
      uint32_t acc = 0, acc_phase = 0;
      const int pos_now = g.pos_base + (g.pos_dev ? *g.pos_dev : 0);
      EpiTma et{&tm_out, &tm_k, &tm_v, staging + (uint32_t)e * 4096u, 0};
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int tile = item % tiles, ks = item / tiles;
        const int m_blk = g.n_fastest ? tile / num_n : tile % num_m, n_blk = g.n_fastest ? tile % num_n : tile / num_m;
        if (!tc::mbar_wait(tfull_bar + 8 * acc, acc_phase, guard)) break;
        tc::fence_after_sync();
        const int row = m_blk * BM + quad * 32 + lane;
        et.row0 = m_blk * BM + quad * 32;
        unsigned long long row_best = 0ull;  // fused argmax: this row's best over the tile's columns
#pragma unroll 1
        for (int c = 0; c < C::COLS_PER_HALF; c += 32) {
          const int col_in_tile = half * C::COLS_PER_HALF + c;
          const int col0 = n_blk * BN + col_in_tile;
          uint32_t r[32];
          tc::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + acc * BN + col_in_tile, r);
          tc::tmem_ld_wait();
          if (col0 < g.N && g.best) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              if (col < g.N) {
                const unsigned bits = __float_as_uint(__uint_as_float(r[j]) + (g.bias ? __ldg(g.bias + col) : 0.0f));
                const unsigned key = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);  // order-preserving float -> uint
                const unsigned long long cand = ((unsigned long long)key << 32) | (0xffffffffu - (unsigned)col);
                row_best = cand > row_best ? cand : row_best;
              }
            }
          } else if (col0 < g.N) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            epilogue_chunk(g, v, row, col0, pos_now, et, lane, ks == 0);
          }
        }
        if (g.best && row < g.M) atomicMax(g.best + 2 * row, row_best);
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tempty_bar + 8 * acc);
        if (++acc == C::ACC) { acc = 0; acc_phase ^= 1; }
      }
      if (g.tma_out && lane == 0) tc::tma_wait_all0();  // staged chunks fully written before the CTA retires
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (warp == 1) tc::tmem_dealloc<C::TMEM_COLS>(tmem);
}

// -------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) of the plain f16 / tf32 GEMM for wide problems: a cluster of two CTAs owns a
// 256 x 256 output tile.  Each CTA loads its 128 rows of A and its 128 of the tile's 256 W rows (32 KB per stage instead
// of 48: six stages), the leader's MMA warp issues one 256 x 256 x K-step MMA for the pair, and each CTA's epilogue
// drains its own 128 accumulator rows exactly like the single-CTA kernel.  Barriers: `full` lives in the leader (one
// arrive.expect_tx per producer, transaction bytes of both CTAs' TMA loads); `empty` and `tfull` are multicast commits
// (every CTA waits on its own copy); `tempty` lives in the leader and collects both CTAs' epilogue warps.
// -------------------------------------------------------------------------------------------------------------------
unsigned long long g_pair_launches = 0;  // launches of the CTA-pair kernel (tests assert the path they mean to cover)
constexpr int PAIR_BN = 256;
constexpr int PAIR_B_STAGE = (PAIR_BN / 2) * ROW_BYTES;  // this CTA's half of the W tile
constexpr int PAIR_STAGE = A_STAGE + PAIR_B_STAGE;
constexpr int PAIR_STAGES = SMEM_BUDGET / PAIR_STAGE;
constexpr int PAIR_SMEM = PAIR_STAGES * PAIR_STAGE + STAGING + 1024 + 384;

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ GemmArgs g) {
  static_assert(MODE != MODE_TF32X3, "the pair kernel has no operand splitters");
  constexpr bool TF32 = MODE != MODE_F16;
  constexpr int BN = PAIR_BN, BK = TF32 ? 32 : 64, HALVES = 2, COLS_PER_HALF = BN / HALVES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_addr(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + PAIR_STAGES * A_STAGE;
  const uint32_t staging = base + PAIR_STAGES * PAIR_STAGE;
  const uint32_t bars = staging + STAGING;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * PAIR_STAGES;
  const uint32_t tfull_bar = bars + 16 * PAIR_STAGES, tempty_bar = tfull_bar + 16;
  const uint32_t slot = tempty_bar + 16, abort_flag = slot + 4;
  uint32_t *slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (slot - raw));
  const tc::Guard guard{g.err, abort_flag};
  const uint32_t cr = tc::cluster_ctarank();  // 0 = leader
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PAIR_STAGES; ++s) {
      tc::mbar_init(full_bar + 8 * s, 2);   // leader's copy is the one in use: one arrive.expect_tx per CTA
      tc::mbar_init(empty_bar + 8 * s, 1);  // multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(tfull_bar + 8 * a, 1);                // multicast commit
      tc::mbar_init(tempty_bar + 8 * a, 2 * 4 * HALVES);  // leader's copy: the epilogue warps of both CTAs
    }
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(abort_flag), "r"(0u));
    tc::fence_mbar_init();
    tc::prefetch_tmap(&tm_a);
    tc::prefetch_tmap(&tm_b);
    if (g.tma_out) tc::prefetch_tmap(&tm_out);
  }
  if (warp == 1) tc::tmem_alloc2<512>(slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::cluster_sync();  // the peer's barriers are initialised before anything arrives on them
  tc::fence_after_sync();
  const uint32_t tmem = *slot_ptr;

  const int num_m = (g.M + 2 * BM - 1) / (2 * BM), num_n = (g.N + BN - 1) / BN, tiles = num_m * num_n;
  const int num_kb = (g.K + BK - 1) / BK;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const uint32_t lead_full = tc::mapa(full_bar, 0), lead_tempty = tc::mapa(tempty_bar, 0);

  if (warp == 0) {  // ---------------- TMA producer (both CTAs) ----------------
    uint32_t stage = 0, phase = 0;
    bool ok = true;
    for (int tile = pair; tile < tiles && ok; tile += npairs) {
      const int m_blk = g.n_fastest ? tile / num_n : tile % num_m, n_blk = g.n_fastest ? tile % num_n : tile / num_m;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (!tc::mbar_wait_u(empty_bar + 8 * stage, phase ^ 1, guard)) { ok = false; break; }
        tc::mbar_expect_tx_cluster_u(lead_full + 8 * stage, PAIR_STAGE);
        tc::tma_load_2d_pair_u(sA + stage * A_STAGE, &tm_a, kb * BK, m_blk * 2 * BM + (int)cr * BM, lead_full + 8 * stage);
        tc::tma_load_2d_pair_u(sB + stage * PAIR_B_STAGE, &tm_b, kb * BK, n_blk * BN + (int)cr * (BN / 2), lead_full + 8 * stage);
        if (++stage == PAIR_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {  // ---------------- MMA issuer (leader only) ----------------
    if (cr == 0) {
      constexpr uint32_t idesc = tc::umma_idesc(TF32 ? 2u : 0u, 2 * BM, BN, 0, 0);
      const uint64_t da0 = tc::umma_desc_sw128(sA, 16, 1024), db0 = tc::umma_desc_sw128(sB, 16, 1024);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      bool ok = true;
      for (int tile = pair; tile < tiles && ok; tile += npairs) {
        if (!tc::mbar_wait_u(tempty_bar + 8 * acc, acc_phase ^ 1, guard)) break;
        tc::fence_after_sync();
        const uint32_t d = tmem + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (!tc::mbar_wait_u(full_bar + 8 * stage, phase, guard)) { ok = false; break; }
          tc::fence_after_sync();
          const uint64_t oa = (uint64_t)(stage * (A_STAGE >> 4)), ob = (uint64_t)(stage * (PAIR_B_STAGE >> 4));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma2_u<TF32>(d, da0 + oa + 2 * k, db0 + ob + 2 * k, idesc, (uint32_t)((kb | k) != 0));
          tc::umma2_commit_both_u(empty_bar + 8 * stage);
          if (++stage == PAIR_STAGES) { stage = 0; phase ^= 1; }
        }
        if (ok) tc::umma2_commit_both_u(tfull_bar + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {  // ---------------- epilogue warps (both CTAs): this CTA's 128 accumulator rows ----------------
    const int e = warp - 2, quad = warp & 3, half = e >> 2;
    uint32_t acc = 0, acc_phase = 0;
    const int pos_now = g.pos_base + (g.pos_dev ? *g.pos_dev : 0);
    EpiTma et{&tm_out, &tm_k, &tm_v, staging + (uint32_t)e * 4096u, 0};
    for (int tile = pair; tile < tiles; tile += npairs) {
      const int m_blk = g.n_fastest ? tile / num_n : tile % num_m, n_blk = g.n_fastest ? tile % num_n : tile / num_m;
      if (!tc::mbar_wait(tfull_bar + 8 * acc, acc_phase, guard)) break;
      tc::fence_after_sync();
      const int row = m_blk * 2 * BM + (int)cr * BM + quad * 32 + lane;
      et.row0 = m_blk * 2 * BM + (int)cr * BM + quad * 32;
      unsigned long long row_best = 0ull;
#pragma unroll 1
      for (int c = 0; c < COLS_PER_HALF; c += 32) {
        const int col_in_tile = half * COLS_PER_HALF + c;
        const int col0 = n_blk * BN + col_in_tile;
        uint32_t r[32];
        tc::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + acc * BN + col_in_tile, r);
        tc::tmem_ld_wait();
        if (col0 < g.N && g.best) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            if (col < g.N) {
              const unsigned bits = __float_as_uint(__uint_as_float(r[j]) + (g.bias ? __ldg(g.bias + col) : 0.0f));
              const unsigned key = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
              const unsigned long long cand = ((unsigned long long)key << 32) | (0xffffffffu - (unsigned)col);
              row_best = cand > row_best ? cand : row_best;
            }
          }
        } else if (col0 < g.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          epilogue_chunk(g, v, row, col0, pos_now, et, lane, true);
        }
      }
      if (g.best && row < g.M) atomicMax(g.best + 2 * row, row_best);
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_cluster(lead_tempty + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (g.tma_out && lane == 0) tc::tma_wait_all0();
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::cluster_sync();  // nobody leaves (or frees TMEM) while the peer may still signal this CTA's barriers
  tc::fence_after_sync();
  if (warp == 1) tc::tmem_dealloc2<512>(tmem);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

template <int MODE, int BN>
void set_attr() {
  static unsigned attr_gen = 0;  // the attribute is per device: redo it after every zg_init
  if (attr_gen != ctx().generation) {
    ZG_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<MODE, BN>::SMEM));
    attr_gen = ctx().generation;
  }
}

template <int MODE, int BN>
void launch_one(const GemmPlan &p) {
  set_attr<MODE, BN>();
  gemm_tc_kernel<MODE, BN><<<p.grid, THREADS, Cfg<MODE, BN>::SMEM, ctx().stream>>>(p.tm_a, p.tm_b, p.tm_out, p.tm_k, p.tm_v, p.args);
  ZG_LAUNCH_CHECK();
}

template <int MODE>
void set_attr_pair() {
  static unsigned attr_gen = 0;
  if (attr_gen != ctx().generation) {
    ZG_CUDA(cudaFuncSetAttribute(gemm_pair_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM));
    attr_gen = ctx().generation;
  }
}
// how many CTA pairs the device runs at once (a persistent grid must not exceed it, or the excess pairs run as a second wave)
template <int MODE>
int max_pairs() {
  static unsigned gen = 0;
  static int n = 0;
  if (gen != ctx().generation) {
    set_attr_pair<MODE>();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (ctx().sm_count > 0 ? ctx().sm_count : 148));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = PAIR_SMEM;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&clusters, gemm_pair_kernel<MODE>, &cfg) != cudaSuccess) {
      cudaGetLastError();
      clusters = 0;
    }
    n = clusters;
    gen = ctx().generation;
    if (getenv("ZG_DEBUG_PAIR")) fprintf(stderr, "zg: max active CTA pairs (mode %d) = %d\n", MODE, n);
  }
  return n;
}

template <int MODE>
void launch_pair(const GemmPlan &p) {
  set_attr_pair<MODE>();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = PAIR_SMEM;
  cfg.stream = ctx().stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ZG_CUDA(cudaLaunchKernelEx(&cfg, gemm_pair_kernel<MODE>, p.tm_a, p.tm_b, p.tm_out, p.tm_k, p.tm_v, p.args));
  ZG_LAUNCH_CHECK();
  g_pair_launches++;
}

template <int MODE>
void launch_mode(const GemmPlan &p) {
  if constexpr (MODE != MODE_TF32X3) {
    if (p.pair) {
      launch_pair<MODE>(p);
      return;
    }
  }
  switch (p.bn) {
    case 256:
      if constexpr (MODE != MODE_TF32X3) {  // (gemm_plan never picks 256 columns for 3xTF32: no TMEM left for the A ring)
        launch_one<MODE, 256>(p);
        break;
      }
    case 192:
      if constexpr (MODE == MODE_TF32X3) {
        launch_one<MODE, 192>(p);
        break;
      }
    case 128: launch_one<MODE, 128>(p); break;
    case 64: launch_one<MODE, 64>(p); break;
    default: launch_one<MODE, 32>(p); break;
  }
}
template <int MODE>
void set_attr_mode() {
  if constexpr (MODE != MODE_TF32X3) {
    set_attr<MODE, 256>();
    set_attr_pair<MODE>();
  }
  if constexpr (MODE == MODE_TF32X3) set_attr<MODE, 192>();
  set_attr<MODE, 128>(); set_attr<MODE, 64>(); set_attr<MODE, 32>();
}

unsigned *g_err_word = nullptr;
bool g_disable_tma_out = false;  // test hook: force the direct-store epilogue
bool g_disable_split_k = false;  // test / A-B hook: never split K (ZG_NO_SPLIT_K=1 in the environment)

}  // namespace

void gemm_init_attrs() {  // outside any stream capture
  set_attr_mode<MODE_F16>();
  set_attr_mode<MODE_TF32>();
  set_attr_mode<MODE_TF32X3>();
  gemm_error_word();
}

static void gemm_forget_device_state() {  // zg_shutdown hook: the watchdog word lives on the device being left
  if (g_err_word) cudaFree(g_err_word);
  g_err_word = nullptr;
}

unsigned *gemm_error_word() {
  if (!g_err_word) {
    ZG_CUDA(cudaMalloc(&g_err_word, sizeof(unsigned)));
    ZG_CUDA(cudaMemset(g_err_word, 0, sizeof(unsigned)));
    note_alloc();
    register_shutdown_hook(gemm_forget_device_state);
  }
  return g_err_word;
}

bool make_tmap_2d(CUtensorMap *out, const void *base, int dtype, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                  uint32_t box_rows, uint32_t box_cols, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error(1, "cuTensorMapEncodeTiled is not available from this driver", __FILE__, __LINE__);
    return false;
  }
  if (((uintptr_t)base & 15) || (pitch_bytes & 15)) {
    set_error(1, "tensor map: base address and row pitch must be multiples of 16 bytes", __FILE__, __LINE__);
    return false;
  }
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstride[1] = {pitch_bytes};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estride[2] = {1, 1};
  const CUtensorMapDataType dt = dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  note_alloc();  // a tensor-map encode counts as start-up work: the hot path replays pre-encoded plans
  const CUresult r = fn(out, dt, 2, const_cast<void *>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error(1, "cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
    return false;
  }
  return true;
}

bool gemm_plan(GemmPlan *p, int mode, const void *A, size_t lda, const void *W, const GemmArgs &args, int bn) {
  const int tf32 = mode != MODE_F16;
  const int es = tf32 ? 4 : 2, bk = 128 / es;
  if (args.M <= 0 || args.N <= 0 || args.K <= 0) return false;
  if ((args.K * es) % 16 != 0) {
    set_error(1, "Linear (tensor-core path): in_features * sizeof(element) must be a multiple of 16", __FILE__, __LINE__);
    return false;
  }
  if (args.k_cache && (args.E % 32 != 0)) {
    set_error(1, "Linear (tensor-core path): n_embed must be a multiple of 32 for the fused cache append", __FILE__, __LINE__);
    return false;
  }
  static const bool env_no_split = getenv("ZG_NO_SPLIT_K") != nullptr;
  if (env_no_split) g_disable_split_k = true;
  const int sms = ctx().sm_count > 0 ? ctx().sm_count : 148;
  const bool explicit_bn = bn != 0;
  const int num_m = (args.M + BM - 1) / BM, num_kb = (args.K + bk - 1) / bk;
  // An in-place residual (x += Linear(h), main.zig:136-145) goes out as TMA reduce-adds, so its K range may be split
  // across work items: pick the widest tile whose (tiles x K slices) still occupy ~every SM with >= 6 k-blocks each.
  const bool can_split = !args.best && !g_disable_tma_out && !g_disable_split_k && bn == 0 && args.epi == TC_EPI_RESIDUAL &&
                         args.resid == args.out && args.ldr == args.ldo && !args.out_f16 &&
                         ((uintptr_t)args.out & 15) == 0 && ((size_t)args.ldo * 4) % 16 == 0;
  int ksplit = 1;
  if (bn == 0) {  // widest tile that still gives every SM a tile; skinny problems stream W with narrow tiles
    bn = 32;
    if (num_m == 1) {
      // One row tile (batched decode, M <= 128): HBM/L2-bound.  Every tile re-reads the whole A panel (128 rows) next to
      // its bn weight rows, and persistent CTAs work in waves, so cost ~ ceil(tiles / SMs) * (128 + bn): 150 tiles
      // of 32 columns (two waves) lose to 75 tiles of 64.
      long best = -1;
      for (int cand : {256, 128, 64, 32}) {
        if (cand > 128 && mode == MODE_TF32X3) continue;
        const int t = (args.N + cand - 1) / cand;
        const long cost = (long)((t + sms - 1) / sms) * (BM + cand);
        if (best < 0 || cost < best) { best = cost; bn = cand; }
      }
    } else {
      // 3xTF32 tiles are shared-memory-bound (operand split + three MMA reads per k-block), which favours wide tiles:
      // accept a tile count a little under the SM count (144 tiles of 128 beat 288 of 64)
      const int need = mode == MODE_TF32X3 ? (sms * 9) / 10 : sms;
      for (int cand : {256, 128, 64}) {
        if (cand > 128 && mode == MODE_TF32X3) continue;
        if (num_m * ((args.N + cand - 1) / cand) >= need) { bn = cand; break; }
      }
      if (mode == MODE_TF32X3) {
        // rounds x shared-memory bytes per k-block (A in and out once, W in, read for the split, lo written, read by three
        // MMAs): a 192-column tile (one accumulator stage) wins where it turns two rounds of 128 into one -- c_fc of the
        // 124M decode step at 1024 rows: 192 tiles of 128 on 148 SMs vs 128 tiles of 192
        const long t128 = (long)num_m * ((args.N + 127) / 128), t192 = (long)num_m * ((args.N + 191) / 192);
        const long c_now = ((num_m * ((args.N + bn - 1) / bn) + sms - 1) / sms) * (32 + 5 * bn / 8);
        if (t192 <= sms && t128 > sms && (32 + 5 * 192 / 8) < c_now && args.epi != TC_EPI_RESIDUAL) bn = 192;
      }
    }
    if (can_split && bn < 128) {
      for (int cand : {128, 64, 32}) {
        if (cand < bn) break;
        const int t = num_m * ((args.N + cand - 1) / cand);
        int ks = (sms + t - 1) / t;
        if (ks > 16) ks = 16;
        while (ks > 1 && num_kb / ks < 6) --ks;
        if (t * ks * 10 >= sms * 9) { bn = cand; ksplit = ks; break; }
      }
    }
  }
  if (mode == MODE_TF32X3 && bn > 192) bn = 128;  // an explicit request for 256 columns: the 3xTF32 kernel has no such tile
  // wide problems: CTA pairs on 256 x 256 tiles (gemm_pair_kernel) when that still gives every pair a tile
  static const bool env_no_pair = getenv("ZG_NO_PAIR") != nullptr;
  const int pair_tiles = ((args.M + 2 * BM - 1) / (2 * BM)) * ((args.N + PAIR_BN - 1) / PAIR_BN);
  p->pair = (!env_no_pair && !explicit_bn && mode != MODE_TF32X3 && bn == 256 && ksplit == 1 && pair_tiles >= sms / 2) ? 1 : 0;
  p->bn = bn;
  p->mode = mode;
  p->args = args;
  p->args.ksplit = ksplit;
  if (!p->args.err) p->args.err = gemm_error_word();
  const int tiles = ((args.M + BM - 1) / BM) * ((args.N + bn - 1) / bn) * ksplit;
  p->grid = tiles < sms ? tiles : sms;
  if (p->pair) {
    const int mp = mode == MODE_F16 ? max_pairs<MODE_F16>() : max_pairs<MODE_TF32>();
    if (mp <= 0) p->pair = 0;
    else p->grid = 2 * (pair_tiles < mp ? pair_tiles : mp);
  }
  if (!make_tmap_2d(&p->tm_a, A, tf32 ? 0 : 1, (uint64_t)args.M, (uint64_t)args.K, lda * es, BM, bk)) return false;
  if (!make_tmap_2d(&p->tm_b, W, tf32 ? 0 : 1, (uint64_t)args.N, (uint64_t)args.K, (uint64_t)args.K * es, p->pair ? bn / 2 : bn, bk)) return false;
  // Epilogue through TMA stores when the output is addressable by a tensor map; a residual that aliases the output
  // (x += ..., main.zig:136-145) becomes a TMA reduce-add so the kernel never reads it.
  GemmArgs &g = p->args;
  // Tile order: concurrently running CTAs should share the smaller operand through L2 and read the larger one from
  // DRAM exactly once.  A (activations) larger than W -> walk all N tiles of one M tile first.
  g.n_fastest = ((size_t)g.M >= (size_t)g.N) ? 1 : 0;
  const int oes = g.out_f16 ? 2 : 4;
  const bool resid_inplace = g.epi == TC_EPI_RESIDUAL && g.resid == g.out && g.ldr == g.ldo && !g.out_f16;
  g.tma_out = g.tma_reduce = g.tma_kv = 0;
  p->tm_out = p->tm_a; p->tm_k = p->tm_a; p->tm_v = p->tm_a;  // valid placeholders
  if (!g.best && !g_disable_tma_out && ((uintptr_t)g.out & 15) == 0 && ((size_t)g.ldo * oes) % 16 == 0 &&
      (g.epi != TC_EPI_RESIDUAL || resid_inplace)) {
    if (!make_tmap_2d(&p->tm_out, g.out, g.out_f16 ? 1 : 0, (uint64_t)g.M, (uint64_t)g.N, (uint64_t)g.ldo * oes, 32, 32,
                      g.out_f16 ? 64 : 128))
      return false;
    g.tma_out = 1;
    g.tma_reduce = resid_inplace ? 1 : 0;
    if (g.k_cache && !g.pos_dev && g.rows_per_seq % 32 == 0 && g.pos_base % 32 == 0 && g.cache_rows > 0 &&
        g.cache_seq_stride == (long long)g.cache_rows * g.E) {
      const uint64_t n_seq = (uint64_t)((g.M + g.rows_per_seq - 1) / g.rows_per_seq);
      if (!make_tmap_2d(&p->tm_k, g.k_cache, 0, n_seq * g.cache_rows, (uint64_t)g.E, (uint64_t)g.E * 4, 32, 32) ||
          !make_tmap_2d(&p->tm_v, g.v_cache, 0, n_seq * g.cache_rows, (uint64_t)g.E, (uint64_t)g.E * 4, 32, 32))
        return false;
      g.tma_kv = 1;
    }
  }
  if (g.ksplit > 1 && !g.tma_reduce) {
    set_error(1, "gemm_plan: split-K needs the TMA reduce-add epilogue", __FILE__, __LINE__);
    return false;
  }
  return true;
}

void gemm_launch(const GemmPlan &p) {
  if (p.mode == MODE_TF32X3) launch_mode<MODE_TF32X3>(p);
  else if (p.mode == MODE_TF32) launch_mode<MODE_TF32>(p);
  else launch_mode<MODE_F16>(p);
}

}  // namespace zg

// =================================================================================================
// C-ABI
// =================================================================================================
using namespace zg;

extern "C" {

// Linear.forward on the tensor cores.  precision: 0 = fp32 operands as kind::tf32 (no copies), 2 = the same with
// 3xTF32 error compensation (fp32-class accuracy), 1 = f16 operands
// (inputs_f16 / weight_f16 are device pointers to f16 copies made by zg_to_f16).  epi: 0 none, 1 GELU, 2 residual.
void zg_linear_forward_tc(const zg_linear *self, const void *inputs, size_t inputs_len, float *outputs, int precision,
                          const void *weight_lowp, int epi, const float *resid, int tile_n) {
  if (!require_ready("zg_linear_forward_tc")) return;
  GemmArgs a;
  a.M = (int)(inputs_len / self->in_features);
  a.N = (int)self->out_features;
  a.K = (int)self->in_features;
  a.bias = self->bias;
  a.out = outputs;
  a.ldo = a.N;
  a.epi = epi;
  a.resid = resid;
  a.ldr = a.N;
  GemmPlan p;
  const void *w = precision == 1 ? weight_lowp : (const void *)self->weight;
  const int mode = precision == 1 ? 0 : (precision == 2 ? 2 : 1);
  if (!gemm_plan(&p, mode, inputs, self->in_features, w, a, tile_n)) return;
  gemm_launch(p);
}

// fp32 -> f16 (round to nearest even) copy: start-up conversion of weights for the kind::f16 path.
__global__ void to_f16_kernel(const float *__restrict__ src, __half *__restrict__ dst, size_t n) {
  for (size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4; i < n; i += (size_t)blockDim.x * gridDim.x * 4) {
    if (i + 4 <= n) {
      const float4 v = *reinterpret_cast<const float4 *>(src + i);
      __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t *>(&a);
      pk.y = *reinterpret_cast<uint32_t *>(&b);
      *reinterpret_cast<uint2 *>(dst + i) = pk;
    } else {
      for (size_t j = i; j < n; ++j) dst[j] = __float2half_rn(src[j]);
    }
  }
}
void zg_to_f16(const float *src, void *dst_f16, size_t n) {
  if (!require_ready("zg_to_f16") || n == 0) return;
  const size_t want = (n / 4 + 255) / 256;
  to_f16_kernel<<<(unsigned)(want < 2368 ? (want ? want : 1) : 2368), 256, 0, ctx().stream>>>(
      src, reinterpret_cast<__half *>(dst_f16), n);
  ZG_LAUNCH_CHECK();
}

unsigned long long zg_tc_pair_launch_count(void) { return g_pair_launches; }
void zg_tc_set_direct_epilogue(int on) { g_disable_tma_out = on != 0; }  // test hook (per-op parity of both epilogues)

int zg_tc_error(void) {  // watchdog word of the tensor-core kernels; 0 when clean (synchronises)
  if (!require_ready("zg_tc_error")) return 1;
  unsigned v = 0;
  unsigned *w = gemm_error_word();
  ZG_CUDA(cudaStreamSynchronize(ctx().stream));
  ZG_CUDA(cudaMemcpy(&v, w, sizeof(v), cudaMemcpyDeviceToHost));
  return (int)v;
}

}  // extern "C"
