// zg_skinny.cu -- Linear.forward (ops.zig:21-46) for the batched DECODE step, M <= 128 rows (one new token of up to 128
// sequences): out[M,N] += bias + X[M,K] . W[N,K]^T with the operands of the tensor-core instruction SWAPPED.
//
// Why a second GEMM.  At M = 64 the step is bound by streaming W (fp32, read once) from HBM, not by flops, and the general
// kernel (zg_gemm.cu: 128 activation rows x BN weight rows per tile) leaves half the SMs idle (75 tiles of 64 columns for
// the 1.5B c_attn) and re-reads the activation panel per tile.  Here:
//   * the UMMA "A" operand (M = 128 rows, TMEM lanes) is a tile of 128 WEIGHT rows, the "B" operand (N = 64 / 128) is
//     the whole batch, so the accumulator D[128 weight rows, batch] uses every lane and W is what the pipeline streams;
//   * work is cut stream-K style: the (weight tile, k-block) grid is divided EVENLY over the CTAs (one per SM), a CTA
//     walks a contiguous range that spans at most a few weight tiles, and every segment's partial sums are added into
//     `out` with fp32 reductions (red.global.add.f32, lanes = consecutive output columns -> 128-byte lines).  `out`
//     therefore starts as zero -- or as the residual stream for `x += Linear(h)` (main.zig:136-145); the segment that
//     owns k-block 0 of a tile adds the bias.  Every SM streams the same number of weight bytes, whatever N is;
//   * an optional transform of the landed X tile in shared memory folds the GELU that follows c_fc (main.zig:80) into
//     the consumer (mlp c_proj): c_fc's split-K partial sums cannot be activated in its own epilogue;
//   * 3xTF32 (fp32-class accuracy, the mode the parity tests bless): the four transform/epilogue warps split every
//     landed tile into hi = x & 0xffffe000 and lo = x - hi; three MMAs per k-step (lo.hi + hi.lo + hi.hi).
// Roles: warp 0 TMA producer (UTMALDG, 128-byte swizzle), warp 1 tcgen05.mma issuer (accumulator in TMEM), warps 2-5
// tile transform during the main loop and tcgen05.ld epilogue at the end of each segment.
// Summation order across segments is not fixed (fp32 atomics), exactly like the split-K mode of zg_gemm.cu: the
// batched decode path is tolerance- and token-checked, not bit-reproducible.
#include "zg_gemm.cuh"
#include "zg_skinny.cuh"

namespace zg {

namespace {

constexpr int WM = 128;          // weight rows per tile = UMMA M = TMEM lanes
constexpr int ROW_BYTES = 128;   // one swizzle row = 32 fp32 (or 64 f16) of K
constexpr int W_STAGE = WM * ROW_BYTES;
constexpr int SK_THREADS = 6 * 32;
constexpr int SK_SMEM_BUDGET = 196 * 1024;

// GELU (ops.zig:225) as x / (1 + e^(-2u)), u = x 0.7978845608 (1 + 0.044715 x^2), on the SFU: the transform warps must
// keep up with the weight stream (libm tanhf is ~10x the instructions); relative error ~1e-6.
__device__ __forceinline__ float gelu_sfu(float x) {
  const float u = x * 0.7978845608f * (1.0f + 0.044715f * x * x);
  return __fdividef(x, 1.0f + __expf(-2.0f * u));
}

template <bool SPLIT, int MB>
struct SkCfg {
  static constexpr int X_STAGE = MB * ROW_BYTES;
  // 3xTF32: X twice in shared memory (as landed / GELU'd, and its lo part); the weight tile's raw words and lo parts go
  // to TMEM (W_COLS columns per stage behind the accumulator) and are the MMAs' A operand from there, so shared memory
  // carries each weight byte twice (TMA in, one read by the splitter) instead of five times
  static constexpr int STAGE = W_STAGE + X_STAGE * (SPLIT ? 2 : 1);
  static constexpr int W_COLS = 64;  // SPLIT: W_hi in columns 0..31, W_lo in 32..63 (one 32-bit column per K element)
  static constexpr int MAX_STAGES = SPLIT ? (512 - MB) / W_COLS : 10;
  static constexpr int STAGES = (SK_SMEM_BUDGET / STAGE) > MAX_STAGES ? MAX_STAGES : (SK_SMEM_BUDGET / STAGE);
  static constexpr int TMEM_COLS = SPLIT ? 512 : MB;
  static constexpr int STAGING = 4 * 4096;  // one 32 x 32 fp32 epilogue chunk per transform/epilogue warp
  static constexpr int SMEM = STAGES * STAGE + STAGING + 1024 /*alignment slack*/ + 512 /*barriers*/;
};

// SPLIT: 3xTF32.  MB: batch columns of the accumulator (64 or 128).  XF: the transform warps touch every landed tile
// (always when SPLIT; otherwise only when X needs the GELU).  F16: both operands are f16 copies (16-bit weight storage:
// half the bytes per weight), kind::f16 with fp32 accumulation; partial sums still leave as fp32.
template <bool SPLIT, int MB, bool XF, bool F16>
__global__ void __launch_bounds__(SK_THREADS, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x,
                   const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ SkinnyArgs g) {
  using C = SkCfg<SPLIT, MB>;
  static_assert(!SPLIT || XF, "the 3xTF32 split is done by the transform warps");
  static_assert(!F16 || (!SPLIT && !XF), "f16 operands are neither split nor transformed");
  constexpr int BK = F16 ? 64 : 32;  // elements per 128-byte swizzle row
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = tc::smem_addr(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sW = base, sX = base + C::STAGES * W_STAGE;
  const uint32_t sXlo = sX + C::STAGES * C::X_STAGE;  // SPLIT only
  const uint32_t staging = base + C::STAGES * C::STAGE;
  const uint32_t bars = staging + C::STAGING;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * C::STAGES, xf_bar = bars + 16 * C::STAGES;
  const uint32_t tfull_bar = bars + 24 * C::STAGES, tempty_bar = tfull_bar + 8;
  const uint32_t slot = tempty_bar + 8, abort_flag = slot + 4;
  uint32_t *slot_ptr = reinterpret_cast<uint32_t *>(smem_raw + (slot - raw));
  const tc::Guard guard{g.err, abort_flag};

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      tc::mbar_init(full_bar + 8 * s, 1);
      tc::mbar_init(empty_bar + 8 * s, 1);
      tc::mbar_init(xf_bar + 8 * s, 4);
    }
    tc::mbar_init(tfull_bar, 1);
    tc::mbar_init(tempty_bar, 4);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(abort_flag), "r"(0u));
    tc::fence_mbar_init();
    tc::prefetch_tmap(&tm_w);
    tc::prefetch_tmap(&tm_x);
    if (g.tma_out) tc::prefetch_tmap(&tm_out);
  }
  if (warp == 1) tc::tmem_alloc<C::TMEM_COLS>(slot);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot_ptr;
  if (g.pdl_trigger) pdl_trigger();  // the next kernel of the step may start its own prologue as SMs free up

  // stream-K: units = (weight tile, k-block) in tile-major order, divided evenly over the CTAs
  const int num_n = (g.N + WM - 1) / WM, num_kb = g.K / BK;
  const long long units = (long long)num_n * num_kb;
  int u0 = (int)(units * blockIdx.x / gridDim.x), u1 = (int)(units * (blockIdx.x + 1) / gridDim.x);
  if (g.best) {  // fused argmax: whole weight tiles per CTA, so that every logit is complete in one accumulator
    u0 = (int)((long long)num_n * blockIdx.x / gridDim.x) * num_kb;
    u1 = (int)((long long)num_n * (blockIdx.x + 1) / gridDim.x) * num_kb;
  }

  // warps 0 and 1 run their loops with all 32 lanes and warp-uniform operands; one elected lane issues (zg_tc.cuh, *_u)
  if (warp == 0) {  // ---------------- TMA producer ----------------
    const uint64_t pol_w = tc::policy_evict_first(), pol_x = tc::policy_evict_last();
    uint32_t stage = 0, phase = 0;
    // The weights are written by no kernel of the step: the first ring-full of W tiles is requested BEFORE waiting for
    // the predecessor grid (programmatic dependent launch), so the pipeline fill overlaps the predecessor's tail.
    const int npre = min(u1 - u0, C::STAGES);
    for (int i = 0; i < npre; ++i) {
      const int u = u0 + i, nt = u / num_kb, kb = u - nt * num_kb;
      tc::mbar_expect_tx_u(full_bar + 8 * i, W_STAGE + C::X_STAGE);
      tc::tma_load_2d_hint_u(sW + i * W_STAGE, &tm_w, kb * BK, nt * WM, full_bar + 8 * i, pol_w);
    }
    pdl_wait();  // X is the predecessor's output
    for (int u = u0; u < u1; ++u) {
      const int nt = u / num_kb, kb = u - nt * num_kb;
      if (u - u0 >= npre) {
        if (!tc::mbar_wait_u(empty_bar + 8 * stage, phase ^ 1, guard)) break;
        tc::mbar_expect_tx_u(full_bar + 8 * stage, W_STAGE + C::X_STAGE);
        tc::tma_load_2d_hint_u(sW + stage * W_STAGE, &tm_w, kb * BK, nt * WM, full_bar + 8 * stage, pol_w);
      }
      tc::tma_load_2d_hint_u(sX + stage * C::X_STAGE, &tm_x, kb * BK, 0, full_bar + 8 * stage, pol_x);
      if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {  // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = tc::umma_idesc(F16 ? 0u : 2u, WM, MB, 0, 0);
    const uint32_t ready_bar = XF ? xf_bar : full_bar;
    // descriptors differ only in the 14-bit start-address field (units of 16 bytes): built once, then an add per MMA
    const uint64_t da0 = tc::umma_desc_sw128(sW, 16, 1024), db0 = tc::umma_desc_sw128(sX, 16, 1024);
    const uint64_t db0_lo = tc::umma_desc_sw128(SPLIT ? sXlo : sX, 16, 1024);
    uint32_t stage = 0, phase = 0, tphase = 0;
    bool ok = true;
    int u = u0;
    while (u < u1 && ok) {
      const int nt = u / num_kb;
      const int seg_end = min(u1, (nt + 1) * num_kb);
      if (!tc::mbar_wait_u(tempty_bar, tphase ^ 1, guard)) break;  // the previous segment's accumulator has been read
      tc::fence_after_sync();
      for (int uu = u; uu < seg_end; ++uu) {
        if (!tc::mbar_wait_u(ready_bar + 8 * stage, phase, guard)) { ok = false; break; }
        tc::fence_after_sync();
        const uint64_t oa = (uint64_t)(stage * (W_STAGE >> 4)), ob = (uint64_t)(stage * (C::X_STAGE >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 4 x 32 bytes of K per swizzle row: UMMA_K = 8 (tf32)
          const uint64_t da = da0 + oa + 2 * k, db = db0 + ob + 2 * k;
          const uint32_t acc = (uint32_t)((uu != u) | (k != 0));
          if constexpr (SPLIT) {
            const uint64_t db_lo = db0_lo + ob + 2 * k;
            const uint32_t tw = tmem + MB + stage * C::W_COLS + 8 * k;  // 8 K elements per MMA = 8 columns
            tc::umma_ts_u<true>(tmem, tw + 32, db, idesc, acc);  // W_lo X_hi
            tc::umma_ts_u<true>(tmem, tw, db_lo, idesc, 1u);     // W_hi X_lo
            tc::umma_ts_u<true>(tmem, tw, db, idesc, 1u);        // W_hi X_hi
          } else {
            tc::umma_u<!F16>(tmem, da, db, idesc, acc);
          }
        }
        tc::umma_commit_u(empty_bar + 8 * stage);
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
      if (ok) tc::umma_commit_u(tfull_bar);
      tphase ^= 1;
      u = seg_end;
    }
  } else {  // ---------------- warps 2..5: tile transform, then the segment's epilogue ----------------
    const int t = threadIdx.x - 64;  // 0..127
    const int quad = warp & 3;       // tcgen05.ld: warp w touches TMEM lanes 32 (w % 4) ..
    uint32_t stage = 0, phase = 0, tphase = 0;
    bool ok = true;
    int u = u0;
    while (u < u1 && ok) {
      const int nt = u / num_kb, kb_first = u - nt * num_kb;
      const int seg_end = min(u1, (nt + 1) * num_kb);
      if constexpr (XF) {
        for (int uu = u; uu < seg_end; ++uu) {
          if (!tc::mbar_wait(full_bar + 8 * stage, phase, guard)) { ok = false; break; }
          const uint32_t xs = sX + stage * C::X_STAGE, ws = sW + stage * W_STAGE;
          // X tile: optional GELU (main.zig:80: the tile holds c_fc's pre-activation), then the hi / lo split.
          // Elementwise, so the swizzled placement is irrelevant; rows past M / N were zero-filled by TMA.
#pragma unroll 2
          for (int i = t; i < C::X_STAGE / 16; i += 128) {
            float4 x;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(xs + 16 * i));
            if (g.xform == SK_XFORM_GELU) {
              x.x = gelu_sfu(x.x); x.y = gelu_sfu(x.y); x.z = gelu_sfu(x.z); x.w = gelu_sfu(x.w);
            }
            if constexpr (SPLIT) {
              const float4 h = make_float4(__uint_as_float(__float_as_uint(x.x) & 0xffffe000u), __uint_as_float(__float_as_uint(x.y) & 0xffffe000u),
                                           __uint_as_float(__float_as_uint(x.z) & 0xffffe000u), __uint_as_float(__float_as_uint(x.w) & 0xffffe000u));
              if (g.xform == SK_XFORM_GELU)  // otherwise the landed word already is the hi operand (zg_tc.cuh)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(xs + 16 * i), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sXlo + stage * C::X_STAGE + 16 * i), "f"(x.x - h.x), "f"(x.y - h.y),
                           "f"(x.z - h.z), "f"(x.w - h.w) : "memory");
            } else {
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(xs + 16 * i), "f"(x.x), "f"(x.y), "f"(x.z), "f"(x.w) : "memory");
            }
          }
          if constexpr (SPLIT) {
            // W: this thread's tile row (TMEM lane) -> 32 raw words + 32 lo parts.  The row's 16-byte chunk c sits at
            // c ^ (row % 8) (SWIZZLE_128B), so the eight rows of a quarter-warp hit eight different bank groups.
            const int w_row = quad * 32 + lane;
            const uint32_t wrow = ws + (uint32_t)w_row * 128u;
            uint32_t x[32], lo[32];
#pragma unroll
            for (int c = 0; c < 8; ++c)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x[4 * c]), "=r"(x[4 * c + 1]), "=r"(x[4 * c + 2]), "=r"(x[4 * c + 3])
                           : "r"(wrow + ((uint32_t)(c ^ (w_row & 7)) << 4)));
#pragma unroll
            for (int i = 0; i < 32; ++i) lo[i] = __float_as_uint(__uint_as_float(x[i]) - __uint_as_float(x[i] & 0xffffe000u));
            const uint32_t tw = tmem + MB + ((uint32_t)(quad * 32) << 16) + stage * C::W_COLS;
            tc::tmem_st32(tw, x);
            tc::tmem_st32(tw + 32, lo);
            tc::tmem_st_wait();
            tc::fence_before_sync();
          }
          tc::fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(xf_bar + 8 * stage);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (!ok) break;
      }
      // ---- epilogue of this segment: out[m, n] += partial (+ bias when the segment starts the tile's K range) ----
      if (!tc::mbar_wait(tfull_bar, tphase, guard)) break;
      tc::fence_after_sync();
      const int n = nt * WM + quad * 32 + lane;
      const float bias = (kb_first == 0 && g.bias && n < g.N) ? __ldg(g.bias + n) : 0.0f;
      const uint32_t my_stage = staging + (uint32_t)(warp - 2) * 4096u;
#pragma unroll 1
      for (int c = 0; c < MB; c += 32) {
        if (c >= g.M) break;  // warp-uniform
        uint32_t r[32];
        tc::tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + c, r);
        tc::tmem_ld_wait();
        if (g.best) {
          // greedy sampling: per batch row m the warp's best (logit, column) pair, raised into best[2 m]
          const unsigned nkey = 0xffffffffu - (unsigned)n;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const unsigned bits = __float_as_uint(__uint_as_float(r[j]) + bias);
            unsigned key = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);  // order-preserving float -> uint
            if (n >= g.N) key = 0u;
            const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
            const unsigned nbest = __reduce_max_sync(0xffffffffu, key == kmax ? nkey : 0u);  // smallest column among the maxima
            if (lane == 0 && c + j < g.M) atomicMax(g.best + 2 * (c + j), ((unsigned long long)kmax << 32) | nbest);
          }
        } else if (g.tma_out) {
          // stage the chunk as [32 batch rows][32 output columns] (lane = column: every store instruction writes one
          // 128-byte row, conflict-free) and hand it to the TMA engine as one reduce-add; TMA clips rows >= M, columns >= N
          if (lane == 0) tc::tma_wait_read0();  // the previous chunk has been read out of this buffer
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(my_stage + (uint32_t)(j * 128 + lane * 4)), "f"(__uint_as_float(r[j]) + bias) : "memory");
          tc::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tc::tma_reduce_add_2d(&tm_out, my_stage, nt * WM + quad * 32, c);
            tc::tma_commit_group();
          }
        } else if (n < g.N) {
          float *dst = g.out + n;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int m = c + j;
            if (m < g.M) atomicAdd(dst + (size_t)m * g.ldo, __uint_as_float(r[j]) + bias);  // RED.E.ADD.F32, result unused
          }
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tempty_bar);
      tphase ^= 1;
      u = seg_end;
    }
    if (g.tma_out && lane == 0) tc::tma_wait_all0();  // staged chunks fully written before the CTA retires
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  if (warp == 1) tc::tmem_dealloc<C::TMEM_COLS>(tmem);
}

template <bool SPLIT, int MB, bool XF, bool F16 = false>
void launch_skinny(const SkinnyPlan &p) {
  static unsigned attr_gen = 0;  // per device: redone after every zg_init
  if (attr_gen != ctx().generation) {
    ZG_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<SPLIT, MB, XF, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<SPLIT, MB>::SMEM));
    attr_gen = ctx().generation;
  }
  ZG_CUDA(launch_pdl(PDL_GEMM_DEP, gemm_skinny_kernel<SPLIT, MB, XF, F16>, dim3(p.grid), dim3(SK_THREADS), (size_t)SkCfg<SPLIT, MB>::SMEM, ctx().stream,
                     p.tm_w, p.tm_x, p.tm_out, p.args));
  ZG_LAUNCH_CHECK();
}

}  // namespace

bool g_skinny_scalar_atomics = false;  // test hook: element-wise fp32 atomics instead of TMA reduce-adds

bool skinny_supported(int M, int N, int K) { return M >= 1 && M <= 128 && N >= 1 && K >= 64 && K % 64 == 0; }

bool skinny_plan(SkinnyPlan *p, int mode, const void *X, size_t ldx, const void *W, const SkinnyArgs &args) {
  const bool f16 = mode == 0;
  const int BK = f16 ? 64 : 32, es = f16 ? 2 : 4;
  if (args.M < 1 || args.M > 128 || args.N < 1 || args.K < BK || args.K % BK != 0 || (f16 && args.xform != SK_XFORM_NONE)) {
    set_error(1, "skinny_plan: needs 1 <= M <= 128 and in_features a multiple of 32 (f16 operands: 64, no transform)", __FILE__, __LINE__);
    return false;
  }
  p->args = args;
  p->args.pdl_trigger = (pdl_mask() & PDL_GEMM_TRIGGER) ? 1 : 0;
  p->mode = mode;
  p->mb = args.M <= 64 ? 64 : 128;
  if (!p->args.err) p->args.err = gemm_error_word();
  const int sms = ctx().sm_count > 0 ? ctx().sm_count : 148;
  const long long units = (long long)((args.N + WM - 1) / WM) * (args.K / BK);
  // every CTA should stream at least ~4 k-blocks, or the pipeline never fills
  long long grid = units / 4;
  if (grid < 1) grid = 1;
  p->grid = (int)(grid < sms ? grid : sms);
  if (!make_tmap_2d(&p->tm_w, W, f16 ? 1 : 0, (uint64_t)args.N, (uint64_t)args.K, (uint64_t)args.K * es, WM, BK)) return false;
  if (!make_tmap_2d(&p->tm_x, X, f16 ? 1 : 0, (uint64_t)args.M, (uint64_t)args.K, (uint64_t)ldx * es, (uint32_t)p->mb, BK)) return false;
  p->tm_out = p->tm_x;  // valid placeholder
  p->args.tma_out = 0;
  if (args.best) {
    const long long tiles = (args.N + WM - 1) / WM;
    p->grid = (int)(tiles < sms ? tiles : sms);
  } else if (!g_skinny_scalar_atomics && ((uintptr_t)args.out & 15) == 0 && ((size_t)args.ldo * 4) % 16 == 0) {
    if (!make_tmap_2d(&p->tm_out, args.out, 0, (uint64_t)args.M, (uint64_t)args.N, (uint64_t)args.ldo * 4, 32, 32, 0)) return false;
    p->args.tma_out = 1;
  }
  return true;
}

void skinny_launch(const SkinnyPlan &p) {
  const bool split = p.mode == 2, xf = split || p.args.xform != SK_XFORM_NONE;
  if (p.mode == 0) {
    if (p.mb == 64) launch_skinny<false, 64, false, true>(p);
    else launch_skinny<false, 128, false, true>(p);
    return;
  }
  if (p.mb == 64) {
    if (split) launch_skinny<true, 64, true>(p);
    else if (xf) launch_skinny<false, 64, true>(p);
    else launch_skinny<false, 64, false>(p);
  } else {
    if (split) launch_skinny<true, 128, true>(p);
    else if (xf) launch_skinny<false, 128, true>(p);
    else launch_skinny<false, 128, false>(p);
  }
}

// tok[m] = column of row m's first maximum, from the packed (orderable logit, ~column) word; also the history row
__global__ void finish_argmax_kernel(const unsigned long long *__restrict__ best, unsigned long long *__restrict__ tok,
                                     unsigned long long *__restrict__ hist, int B, const int *pos_dev) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= B) return;
  const unsigned long long t = (unsigned long long)(0xffffffffu - (unsigned)(best[2 * m] & 0xffffffffull));
  tok[m] = t;
  if (hist) hist[(size_t)(*pos_dev) * B + m] = t;
}
void skinny_finish_argmax(const unsigned long long *best, unsigned long long *tok, unsigned long long *hist, int B, const int *pos_dev) {
  finish_argmax_kernel<<<(B + 127) / 128, 128, 0, ctx().stream>>>(best, tok, hist, B, pos_dev);
  ZG_LAUNCH_CHECK();
}

void skinny_init_attrs() {  // outside any stream capture: cudaFuncSetAttribute is not capturable
  static unsigned gen = 0;
  if (gen == ctx().generation) return;
  gen = ctx().generation;
  ZG_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<true, 64, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<true, 64>::SMEM));
  ZG_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<false, 64, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<false, 64>::SMEM));
  ZG_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<false, 64, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<false, 64>::SMEM));
  ZG_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<true, 128, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<true, 128>::SMEM));
  ZG_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<false, 128, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<false, 128>::SMEM));
  ZG_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<false, 128, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<false, 128>::SMEM));
  ZG_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<false, 64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<false, 64>::SMEM));
  ZG_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<false, 128, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SkCfg<false, 128>::SMEM));
}

}  // namespace zg

using namespace zg;

extern "C" {

// Linear.forward for M <= 128 rows through the swapped-operand stream-K kernel: outputs[M,N] += bias + inputs . W^T.
// `outputs` must hold zeros (plain Linear) or the residual (x += Linear(h)) on entry.  precision: 0 = TF32, 2 = 3xTF32.
// xform: 0 none, 1 = inputs := gelu(inputs) on the fly (the GELU between c_fc and mlp c_proj, main.zig:80).
void zg_linear_forward_skinny(const zg_linear *self, const void *inputs, size_t inputs_len, float *outputs, int precision,
                              int xform, const void *weight_lowp) {
  if (!require_ready("zg_linear_forward_skinny")) return;
  g_skinny_scalar_atomics = (xform & 2) != 0;  // test hook (bit 1): the scalar-atomic epilogue
  xform &= 1;
  SkinnyArgs a;
  a.M = (int)(inputs_len / self->in_features);
  a.N = (int)self->out_features;
  a.K = (int)self->in_features;
  a.bias = self->bias;
  a.out = outputs;
  a.ldo = a.N;
  a.xform = xform;
  SkinnyPlan p;
  const bool planned = precision == 1 ? skinny_plan(&p, 0, inputs, self->in_features, weight_lowp, a)
                                      : skinny_plan(&p, precision == 2 ? 2 : 1, inputs, self->in_features, self->weight, a);
  g_skinny_scalar_atomics = false;
  if (planned) skinny_launch(p);
}

// Greedy sampling through the tied lm_head without materialising logits (main.zig:193 + argmax): tokens[m] = index of the
// first maximum of inputs[m,:] . W^T (+ bias).  `best_scratch` = 2 * M u64 of device scratch.  Asynchronous.
void zg_linear_argmax_skinny(const zg_linear *self, const float *inputs, size_t inputs_len, int precision,
                             unsigned long long *best_scratch, size_t *tokens_dev) {
  if (!require_ready("zg_linear_argmax_skinny")) return;
  SkinnyArgs a;
  a.M = (int)(inputs_len / self->in_features);
  a.N = (int)self->out_features;
  a.K = (int)self->in_features;
  a.bias = self->bias;
  a.best = best_scratch;
  SkinnyPlan p;
  if (!skinny_plan(&p, precision == 2 ? 2 : 1, inputs, self->in_features, self->weight, a)) return;
  ZG_CUDA(cudaMemsetAsync(best_scratch, 0, (size_t)a.M * 16, ctx().stream));
  skinny_launch(p);
  skinny_finish_argmax(best_scratch, (unsigned long long *)tokens_dev, nullptr, a.M, nullptr);
}

}  // extern "C"
