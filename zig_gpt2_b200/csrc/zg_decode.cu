// zg_decode.cu -- the fused batch-1 decode engine: GPT.forward / GPT.sample / generate
// (src/main.zig:178-207, 322-342) as ONE persistent cooperative kernel.
//
// Design (B200, 148 SMs, HBM-bound: 495 MB of fp32 weights per token at 124M):
//   * one CTA per SM, 7 consumer warps + 1 producer warp (two warps per scheduler: 255 registers per thread);
//   * the producer warp streams this CTA's share of every weight matrix, in execution order, through a
//     shared-memory ring with cp.async.bulk (UBLKCP) + mbarrier complete_tx, L2 evict-first.  The stream does
//     not depend on activations, so it runs ahead across layer phases and tokens.  All units of a batch (up to
//     28 rows) complete on ONE mbarrier, so a consumer waits once per batch;
//   * consumers: warp w owns rows w, w + 7, ... of a batch; the activation vector sits in the warp's registers;
//     a packed butterfly reduces the warp's (up to 4) row sums in 6 shuffles; lanes 0/8/16/24 apply the fused
//     epilogue.  No shared-memory partials and no CTA sync after the dot;
//   * phases of a layer (every phase boundary is a cross-SM exchange, the dominant cost at this size)
//        P1 LN1 + c_attn (+ K/V append)   ops.zig:143-158, main.zig:121-123
//        P2 attention over the time-major cache (flash-decoding splits when T is long)  ops.zig:160-171
//        P3 attn c_proj + residual         ops.zig:172, main.zig:136-139
//        P4 LN2 + c_fc + GELU + mlp c_proj main.zig:140, :79-81 -- the CTA that owns hidden units [j0, j1)
//           computes them AND multiplies them into its rows of c_proj^T, so the 4E-wide GELU vector never leaves
//           the SM; every CTA publishes an E-wide partial result
//        P5 reduce-scatter: CTA c sums its 5-6 elements over the 148 partial vectors in a fixed order
//           (deterministic; same-address atomics were measured at ~45 ns each, 7 us per layer), adds the bias and
//           the residual (main.zig:142-145) and publishes its slice of the new stream; no weights, no GEMV
//     then ln_f + tied lm_head + argmax (main.zig:189-194);
//   * FOUR phases per layer carry weights; the fifth is a pure exchange step;
//   * LayerNorm is folded into the matrix that follows it (start-up copies W' = W diag(g), c1 = W' 1,
//     c2 = W b + bias): y = rstd (W' x - mean c1) + c2 -- algebraically ops.zig:86-101 followed by ops.zig:21-46.
//     The statistics (single pass E[x], E[x^2], ops.zig:86-95) are computed by every warp from its register copy
//     while the dot products are already running; nothing waits for them until the epilogue;
//   * NO grid barrier and NO memory fence between phases.  Every activation word crosses SMs as one 64-bit
//     store {epoch : 32 | fp32 bits : 32}; a consumer gathers the vector it needs with 128-bit relaxed loads
//     and spins until every word carries the epoch of the phase that produces it (the NCCL "LL" idea: the
//     flag travels inside the datum, so there is nothing to fence);
//   * a CTA that is early does NOT start polling at once: before the first poll of the QKV / lm_head gather, of the
//     reduce phase and of the argmax gather it spins on the clock for a hold-off (ZG_PD_*): polling the L2 lines that
//     the slower CTAs are still storing to slowed those stores -- the largest single loss found in the exchange
//     (124M: 161 -> 147 us/token);
//   * the layer table lives in __constant__ memory, so no phase starts with a dependent global load;
//   * the token loop of generate() runs inside the kernel; each token is written to device memory and to a
//     pinned host ring, so the host only waits once per call.
#include <chrono>
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include "zg_common.cuh"
#include "zg_ptx.cuh"

namespace zg {
void launch_sample_rows(const float *logits, size_t pitch, int V, const void *sample_params_dev, const int *step_dev, int step,
                        unsigned long long *tok, unsigned long long *hist, int B, unsigned long long *host_ring);
void launch_softmax_temp(float *x, size_t n, float temp);
void launch_weighted_index(const float *p, size_t n, float u, unsigned long long *out);

typedef unsigned long long u64;

constexpr int NTHREADS = NCT + 32;  // + producer warp
constexpr int MAXSLOTS = 32;
constexpr int MAX_LAYERS = 64;
constexpr int ATT_ROWS = 16;  // rows per warp of an attention work item before the head is split across CTAs
constexpr int ATT_CHUNK = ATT_ROWS * NCW;  // KV rows per attention work item before splitting (one register round)
constexpr int ATT_SMAX = 8;          // at most this many splits per head; longer contexts loop over rounds inside a split
constexpr int PROF_MAX = 16384;
// sm.red layout: per-warp attention (m, l), per-warp argmax (value, index), the reduced token
constexpr int RED_M = 0, RED_L = 16, RED_B = 32, RED_I = 48, RED_TOK = 64, RED_FLOATS = 96;
static_assert(NCW <= 16, "sm.red and sm.part hold 16 per-warp entries");
constexpr int MAXNE = 16;        // elements of the stream a CTA owns in the reduce phase: ceil(E / SMs) <= 16

struct LayerDesc {
  const float *wq, *c1q, *c2q;    // LN1-folded c_attn: W diag(g) [3E,E], its row sums, W b + bias
  const float *w_proj, *b_proj;   // attention c_proj (reference layout)
  const float *wfc, *c1f, *c2f;   // LN2-folded c_fc [4E,E]
  const float *w2t, *b_proj2;     // mlp c_proj TRANSPOSED [4E,E]; its bias
  float *k_cache, *v_cache;
};
__constant__ LayerDesc c_layers[MAX_LAYERS];

struct DecodeParams {
  int E, H, hd, L, V, C;
  int nslot, slotf;  // ring geometry: slotf = 4E floats per slot
  const float *wte, *wpe, *lnf_g, *lnf_b;
  const float *wte_f, *c1h, *c2h;  // ln_f-folded tied lm_head
  // flagged exchange buffers: word = {epoch << 32 | fp32 bits}
  u64 *xres_f;  // [E]   residual stream after the attention half (P3 output)
  u64 *q_f;     // [E]   query of the current token
  u64 *kvn_f;   // [2E]  K row then V row of the current token (the cache gets the same values, unflagged)
  u64 *att_f;   // [E]   attention output
  u64 *amax_f;  // [2G]  per-CTA argmax partial: value word, index word
  u64 *part_f;  // [G][E] partial mlp c_proj outputs, one vector per CTA
  u64 *xnew_f;  // [E]   residual stream after the MLP half (P5 output)
  unsigned epoch_base;
  float *xres_out;  // [E] state.o: the reference leaves the pre-ln_f stream there (main.zig:116-118)
  float *xout;      // [E] state.x: ln_f output
  float *logits;    // [V] state.logits
  u64 *attp_f;           // [H][ATT_SMAX][hd+2] flagged flash-decoding partials: m, l, unnormalised o[hd]
  unsigned *err;         // sticky watchdog word
  const u64 *prompt;     // device, n_prompt entries; null => `single_token` is the forced token
  u64 single_token;
  int n_prompt;
  u64 *tokens;                 // device [C]: token forwarded/sampled at every step
  volatile u64 *tokens_host;   // pinned host ring [C]
  u64 *last_token;             // device: argmax of the last logits computed
  volatile u64 *result_host;   // pinned host {token, sequence number}: GPT.sample's answer without a copy + synchronise
  u64 result_seq;
  int first_step, n_steps;
  int force_logits;  // compute logits + argmax on every step (GPT.forward(compute_logits=true) on a prompt step)
  int store_logits;  // also write the logits vector to global memory
  int write_xout;    // write ln_f(x) to xout (state.x) and the stream to xres_out on the last step
  u64 *prof;
  int prof_cta;  // CTA whose thread 0 records the cycle timeline
};

// Cycle-level breakdown of a phase (zg_engine_read_profile): thread 0 of one chosen CTA keeps up to 12 %clock readings
// in registers and dumps them as (tag = 512 + 16 * phase kind + point, cycles) pairs when the phase ends.  Always
// compiled in: with profiling off every hook is one not-taken uniform branch.  (Round 1 measured builds without the
// hooks 5-8 % slower -- the never-taken dump block at the end of every phase changes ptxas's schedule; the A/B log is
// profiles/r01_ab_experiments.txt.  The kernel is sensitive to code layout at this level: every change is re-timed.)
struct Clk {
  unsigned t[12];
  u64 *buf;
  int i;
  __device__ __forceinline__ void at(int k) {
    if (buf) {
      if (k == 0) {
#pragma unroll
        for (int q = 1; q < 12; ++q) t[q] = 0u;
      }
      t[k] = (unsigned)clock64();
    }
  }
  __device__ __forceinline__ void dump(int kind) {
    if (buf && i + 12 < PROF_MAX) {
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        buf[2 * (i + k)] = (u64)(512 + 16 * kind + k);
        buf[2 * (i + k) + 1] = (u64)t[k];
      }
      i += 12;
      buf[2 * PROF_MAX] = (u64)i;
    }
  }
};


// Hold-off before the first poll of a phase.  A CTA that finishes a phase early starts to poll the flagged words of the
// next phase's input at once; 148 CTAs x 224 threads polling the very L2 lines that the slower CTAs are still storing
// to slows those stores down (measured: 161.6 -> 145.7 us/token at 124M with the hold-offs below, the largest single
// gain of the round; a back-off BETWEEN polls instead costs detection latency and loses).  The input of the QKV /
// lm_head phases (P5's output) and of the reduce phase (P4's partials) never arrives sooner than ~1,000 cycles after a
// CTA is ready for it, so these phases spin on the clock first; the attention consumers (P3) are not held off: the
// attention CTAs themselves are the late ones there.  Cycles at the SM clock; 0 disables.
#ifndef ZG_PD_P1
#define ZG_PD_P1 800
#endif
#ifndef ZG_PD_P3
#define ZG_PD_P3 0
#endif
#ifndef ZG_PD_P4
#define ZG_PD_P4 0
#endif
#ifndef ZG_PD_P5
#define ZG_PD_P5 800
#endif
#ifndef ZG_PD_AMAX  // before gathering the 148 argmax partials at the end of the lm_head phase
#define ZG_PD_AMAX 1500
#endif
__device__ __forceinline__ void predelay(int cycles) {
  if (cycles <= 0) return;
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
}

// Transcendentals of the attention softmax and of GELU.  Default: the SFU forms (ex2.approx / rcp.approx, relative error
// ~1e-6: two orders of magnitude inside the 1e-4 tolerance the per-op comparator allows, tests/test_gpu_model.py), because
// libm-accurate expf / tanhf / division are 20-40 dependent instructions each ON the critical path of a phase that runs
// once per layer (the final combine of the attention phase alone went from ~720 to ~300 cycles).  -DZG_EXACT_MATH
// restores the libm forms.  GELU: 0.5 x (1 + tanh u) == x / (1 + e^(-2u)) with u = x 0.7978845608 (1 + 0.044715 x^2), ops.zig:225.
#ifdef ZG_EXACT_MATH
__device__ __forceinline__ float fexp(float x) { return expf(x); }
__device__ __forceinline__ float fdiv(float a, float b) { return a / b; }
__device__ __forceinline__ float gelu_dec(float x) { return gelu_ref(x); }
#else
__device__ __forceinline__ float fexp(float x) { return __expf(x); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ float gelu_dec(float x) {
  const float u = x * 0.7978845608f * (1.0f + 0.044715f * x * x);
  return __fdividef(x, 1.0f + __expf(-2.0f * u));
}
#endif

// ---- flag-in-data exchange ------------------------------------------------------------------------
__device__ __forceinline__ void st_flag(u64 *p, float v, unsigned ep) {
  const u64 w = ((u64)ep << 32) | (u64)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
// The new K/V row goes to the cache with a gpu-scope relaxed store: other CTAs read it with ld.global.cg in LATER steps
// of the same launch (this step's readers take it from the flagged exchange), ordered only by the chain of flagged
// words in between -- a morally strong store keeps that well-defined in the PTX memory model.
__device__ __forceinline__ void st_cache(float *p, float v) {
  asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_pair(const u64 *p) {
  ulonglong2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ u64 ld_word(const u64 *p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
#ifdef ZG_NOWAIT  // timing experiment only: never wait for another CTA (results are garbage)
__device__ __forceinline__ bool pair_ok(const ulonglong2 &, unsigned) { return true; }
#else
__device__ __forceinline__ bool pair_ok(const ulonglong2 &v, unsigned ep) {
  return (unsigned)(v.x >> 32) == ep && (unsigned)(v.y >> 32) == ep;
}
#endif
__device__ __forceinline__ float lo_f(u64 w) { return __uint_as_float((unsigned)w); }

__device__ __forceinline__ bool wd_tripped(const Watchdog &wd) {
  uint32_t t;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(t) : "r"(wd.tripped_smem));
  return t != 0;
}
__device__ __forceinline__ void wd_trip(const Watchdog &wd, unsigned code) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(wd.tripped_smem), "r"(1u));
  atomicExch(wd.err_global, code);
}
// spin until both words of a pair carry `ep` (watchdogged: a protocol bug must not hang the GPU)
__device__ __noinline__ ulonglong2 spin_pair(const u64 *p, unsigned ep, Watchdog wd) {
  ulonglong2 v = ld_pair(p);
  if (wd_tripped(wd)) return v;
  const long long t0 = clock64();
  unsigned spins = 0;
  while (!pair_ok(v, ep)) {
    v = ld_pair(p);
    if ((++spins & 255u) == 0 && clock64() - t0 > WATCHDOG_CYCLES) {
      wd_trip(wd, 3u);
      break;
    }
  }
  return v;
}
__device__ __noinline__ u64 spin_word(const u64 *p, unsigned ep, Watchdog wd) {
  u64 v = ld_word(p);
  if (wd_tripped(wd)) return v;
  const long long t0 = clock64();
  unsigned spins = 0;
  while ((unsigned)(v >> 32) != ep) {
    v = ld_word(p);
    if ((++spins & 255u) == 0 && clock64() - t0 > WATCHDOG_CYCLES) {
      wd_trip(wd, 4u);
      break;
    }
  }
  return v;
}
// Gather n floats (n % 4 == 0) whose words must carry epoch `ep` into shared memory: all loads of a thread are issued
// before the first check, so the common case costs one L2 round trip; late quads are re-polled together.  ONE 256-bit load
// per thread and pass (LDG.E.ENL2.256.STRONG.GPU: four flagged words = one 32-byte sector): a thread's strong loads do not
// overlap well -- every additional flagged load per thread and pass was measured at +100..200 cycles on the gather's
// critical path (one polling warp with 12 128-bit pairs per lane: 207 us/token; seven warps with 2 pairs per lane: 145.3;
// one 256-bit quad per thread: 143.3) -- so the widest load wins: E = 768 needs 192 loads, one per thread.
struct Quad { u64 a, b, c, d; };
__device__ __forceinline__ Quad ld_quad(const u64 *p) {
  Quad v;
  asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p) : "memory");
  return v;
}
#ifdef ZG_NOWAIT
__device__ __forceinline__ bool quad_ok(const Quad &, unsigned) { return true; }
#else
__device__ __forceinline__ bool quad_ok(const Quad &v, unsigned ep) {
  return (unsigned)(v.a >> 32) == ep && (unsigned)(v.b >> 32) == ep && (unsigned)(v.c >> 32) == ep && (unsigned)(v.d >> 32) == ep;
}
#endif
template <int GB>
__device__ __forceinline__ void gather_flagged256(float *dst_smem, const u64 *src, int n, unsigned ep, Watchdog wd) {
  const int nquads = n >> 2;  // n % 4 == 0 (n_embed % 8 == 0 is checked at create)
#pragma unroll 1
  for (int base = 0; base < nquads; base += GB * NCT) {
    Quad v[GB];
    bool all_ok = true;
#pragma unroll
    for (int j = 0; j < GB; ++j) {
      const int idx = base + j * NCT + (int)threadIdx.x;
      if (idx < nquads) v[j] = ld_quad(src + 4 * idx);
    }
#pragma unroll
    for (int j = 0; j < GB; ++j) {
      const int idx = base + j * NCT + (int)threadIdx.x;
      if (idx < nquads) all_ok = all_ok && quad_ok(v[j], ep);
    }
    if (!all_ok && !wd_tripped(wd)) {
      const long long t0 = clock64();
      do {
        all_ok = true;
#pragma unroll
        for (int j = 0; j < GB; ++j) {
          const int idx = base + j * NCT + (int)threadIdx.x;
          if (idx < nquads && !quad_ok(v[j], ep)) v[j] = ld_quad(src + 4 * idx);
        }
#pragma unroll
        for (int j = 0; j < GB; ++j) {
          const int idx = base + j * NCT + (int)threadIdx.x;
          if (idx < nquads) all_ok = all_ok && quad_ok(v[j], ep);
        }
        if (!all_ok && clock64() - t0 > WATCHDOG_CYCLES) {
          wd_trip(wd, 3u);
          break;
        }
      } while (!all_ok);
    }
#pragma unroll
    for (int j = 0; j < GB; ++j) {
      const int idx = base + j * NCT + (int)threadIdx.x;
      if (idx < nquads)
        reinterpret_cast<float4 *>(dst_smem)[idx] = make_float4(lo_f(v[j].a), lo_f(v[j].b), lo_f(v[j].c), lo_f(v[j].d));
    }
  }
}
__device__ __forceinline__ void st_flag2(u64 *p, float v0, float v1, unsigned ep) {  // two adjacent flagged words
  const u64 w0 = ((u64)ep << 32) | (u64)__float_as_uint(v0), w1 = ((u64)ep << 32) | (u64)__float_as_uint(v1);
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}

// rows [r0, r1) of an N-row matrix owned by this CTA in a phase; `rot` rotates which CTAs get the remainder rows
__device__ __forceinline__ void row_range(int cta, int G, int rot, int N, int &r0, int &r1) {
  int c = cta + rot;
  if (c >= G) c -= G;
  r0 = (int)(((unsigned)c * (unsigned)N) / (unsigned)G);  // c < G <= ~150 SMs, N <= 4E or V: fits 32 bits
  r1 = (int)(((unsigned)(c + 1) * (unsigned)N) / (unsigned)G);
}
__device__ __forceinline__ int phase_rot(int layer, int ph, int G) { return ((layer * 5 + ph) * 29) % G; }

// Phases of a layer: 0 = LN1 + c_attn, 1 = attention (no weights), 2 = attn c_proj + residual,
// 3 = LN2 + c_fc + GELU + mlp c_proj (partial sums), 4 = reduce-scatter of the partial sums (no weights); the tied
// lm_head is phase index 5L of a step.
enum { M_REDUCE = -2, M_ATTN = -1, M_QKV = 0, M_RESID = 1, M_MLP = 2, M_LMHEAD = 3 };

struct Smem {
  float *ring;   // nslot * slotf
  float *vec;    // 2 x E: activation vector of the current GEMV phase, double-buffered: the previous phase's vector
                 // stays readable (residual operands), and with no CTA-wide sync at the end of a phase a fast warp
                 // may already be gathering the next vector while a slow one still reads the current one
  float *fbuf;   // 64: GELU outputs of the hidden units this CTA owns
  float *part;   // NCW * hd attention partial outputs
  float *red;    // RED_FLOATS
  uint32_t full0, empty0;  // shared addresses of mbarrier arrays
  Watchdog wd;             // sticky global error word + CTA-local tripped flag
};

// warp-wide float max in one instruction (CREDUX.MAX.F32, sm_100a)
__device__ __forceinline__ float warp_max_redux(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

// Packed butterfly: every lane holds N partial sums v[0..N); afterwards (N = 2^k <= 16) the lane whose bits
// 4, 3, ... (k bits, most significant first) spell index i holds the warp total of v[i] in v[0].  N + log2(32/N)
// shuffles instead of 5 N.
template <int N>
__device__ __forceinline__ float packed_reduce(float (&v)[N], int lane) {
  static_assert(N == 1 || N == 2 || N == 4 || N == 8 || N == 16, "power of two");
  int off = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float keep = hi ? v[i + n / 2] : v[i];
      const float send = hi ? v[i] : v[i + n / 2];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  float r = v[0];
  for (; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
  return r;
}

// Attention work item (head h, split s of S) over cache rows [t0,t1) -- ops.zig:249-307 without the
// whole-cache transposes: K/V of earlier tokens are read in place from the time-major cache (head stride hd = 64,
// time stride E); q and the current token's K/V row arrive through the flagged exchange (epoch `ep_in`).
// A warp owns rows t0 + warp + NCW u; a round covers AR rows per warp, all of them in registers BEFORE q is
// waited for (cache rows do not depend on this step), so a split that fits one round costs no exposed L2 latency
// after q lands.  Scores: per-lane partial dot over the lane's 2 dims, packed butterfly (AR + 5 - log2 AR shuffles
// for AR rows), one exp per lane, p broadcast by shuffle for the PV accumulation.
template <int AR>
__device__ __forceinline__ void attention_item(const DecodeParams &p, const Smem &sm, int l, int h, int s, int S, int T,
                                               unsigned ep_in, unsigned ep_out, Clk &ck) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int hd = 64;
  constexpr int LG = AR == 16 ? 4 : AR == 8 ? 3 : 2;  // log2 AR
  const int E = p.E, pos = T - 1;
  const int chunk = (T + S - 1) / S;
  const int t0 = s * chunk, t1 = min(T, t0 + chunk);
  const float scale = 0.125f;  // 1 / sqrt(64)
  const float *kh = c_layers[l].k_cache + h * hd;
  const float *vh = c_layers[l].v_cache + h * hd;
  float *po = sm.part;  // [NCW][hd] per-warp partial outputs

  const bool has_new = (pos >= t0 && pos < t1);
  float2 kv[AR], vv[AR];
  const int tfirst = t0 + warp;
#pragma unroll
  for (int u = 0; u < AR; ++u) {
    const int tt = tfirst + u * NCW;
    kv[u] = make_float2(0.0f, 0.0f);
    vv[u] = make_float2(0.0f, 0.0f);
    if (tt < t1 && tt != pos) {
      kv[u] = __ldcg(reinterpret_cast<const float2 *>(kh + (size_t)tt * E) + lane);
      vv[u] = __ldcg(reinterpret_cast<const float2 *>(vh + (size_t)tt * E) + lane);
    }
  }
  ck.at(1);
  ulonglong2 qw = ld_pair(p.q_f + h * hd + 2 * lane);
  ulonglong2 kw = qw, vw = qw;
  if (has_new) {
    kw = ld_pair(p.kvn_f + h * hd + 2 * lane);
    vw = ld_pair(p.kvn_f + E + h * hd + 2 * lane);
  }
  if (!pair_ok(qw, ep_in)) qw = spin_pair(p.q_f + h * hd + 2 * lane, ep_in, sm.wd);
  if (has_new) {
    if (!pair_ok(kw, ep_in)) kw = spin_pair(p.kvn_f + h * hd + 2 * lane, ep_in, sm.wd);
    if (!pair_ok(vw, ep_in)) vw = spin_pair(p.kvn_f + E + h * hd + 2 * lane, ep_in, sm.wd);
  }
  ck.at(2);
  const float2 qv = make_float2(lo_f(qw.x), lo_f(qw.y));
  const float2 knew = make_float2(lo_f(kw.x), lo_f(kw.y)), vnew = make_float2(lo_f(vw.x), lo_f(vw.y));
  // row index (within a round) whose score this lane holds after the packed butterfly: the top LG lane bits
  const int myu = lane >> (5 - LG);

  float mw = -INFINITY, lw = 0.0f;
  float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll 1
  for (int t = tfirst; t < t1; t += AR * NCW) {
    if (t != tfirst) {
#pragma unroll
      for (int u = 0; u < AR; ++u) {
        const int tt = t + u * NCW;
        if (tt < t1 && tt != pos) {
          kv[u] = __ldcg(reinterpret_cast<const float2 *>(kh + (size_t)tt * E) + lane);
          vv[u] = __ldcg(reinterpret_cast<const float2 *>(vh + (size_t)tt * E) + lane);
        }
      }
    }
    float a[AR];
#pragma unroll
    for (int u = 0; u < AR; ++u) {
      if (t + u * NCW == pos) {
        kv[u] = knew;
        vv[u] = vnew;
      }
      a[u] = fmaf(qv.x, kv[u].x, qv.y * kv[u].y);
    }
    const float sc = packed_reduce<AR>(a, lane) * scale;  // score of row t + myu * NCW
    const bool valid = (t + myu * NCW) < t1;
    const float mnew = fmaxf(mw, warp_max_redux(valid ? sc : -INFINITY));
    const float corr = (mw == -INFINITY) ? 0.0f : fexp(mw - mnew);
    const float pt = valid ? fexp(sc - mnew) : 0.0f;
    lw *= corr;
    acc.x *= corr;
    acc.y *= corr;
#pragma unroll
    for (int u = 0; u < AR; ++u) {
      const float pu = __shfl_sync(0xffffffffu, pt, u << (5 - LG));
      lw += pu;                          // every lane accumulates the whole denominator: no reduction afterwards
      acc.x = fmaf(pu, vv[u].x, acc.x);  // rows past t1 carry pu == 0 (their stale vv is finite)
      acc.y = fmaf(pu, vv[u].y, acc.y);
    }
    mw = mnew;
  }
  ck.at(3);
  po[warp * hd + 2 * lane] = acc.x;
  po[warp * hd + 2 * lane + 1] = acc.y;
  if (lane == 0) {
    sm.red[RED_M + warp] = mw;
    sm.red[RED_L + warp] = lw;
  }
  consumer_sync();
  ck.at(4);
  if (tid < hd) {
    float m = -INFINITY;
#pragma unroll
    for (int w = 0; w < NCW; ++w) m = fmaxf(m, sm.red[RED_M + w]);
    float lsum = 0.0f, o = 0.0f;
#pragma unroll
    for (int w = 0; w < NCW; ++w) {
      const float mwv = sm.red[RED_M + w];
      const float sc = (mwv == -INFINITY) ? 0.0f : fexp(mwv - m);
      lsum = fmaf(sm.red[RED_L + w], sc, lsum);
      o = fmaf(po[w * hd + tid], sc, o);
    }
    if (S == 1) {
      st_flag(p.att_f + h * hd + tid, fdiv(o, lsum), ep_out);
    } else {  // flash-decoding partial (m, l, unnormalised o): every consumer of the attention output combines the S
              // partials of a head itself while it gathers the vector (gather_att_partials), so no fence, no counter
      u64 *mine = p.attp_f + ((size_t)h * ATT_SMAX + s) * (hd + 2);
      st_flag(mine + 2 + tid, o, ep_out);
      if (tid == 0) st_flag2(mine, m, lsum, ep_out);
    }
  }
}

// Attention output vector from S (2..ATT_SMAX) flagged partials per head: out = sum_s o_s e^(m_s - M) / sum_s l_s e^(m_s - M).
// A thread handles pairs of adjacent elements; all 2 S loads of a pair are in flight before the first check.
__device__ __forceinline__ void gather_att_partials(float *dst_smem, const u64 *attp, int E, int S, unsigned ep, Watchdog wd) {
  constexpr int hd = 64;
  const int npairs = E >> 1;
#pragma unroll 1
  for (int idx = (int)threadIdx.x; idx < npairs; idx += NCT) {
    const int h = idx >> 5, d = (idx & 31) * 2;
    const u64 *base = attp + (size_t)h * ATT_SMAX * (hd + 2);
    ulonglong2 ml[ATT_SMAX], ov[ATT_SMAX];
    bool all_ok = true;
#pragma unroll
    for (int q = 0; q < ATT_SMAX; ++q) {
      if (q < S) {
        ml[q] = ld_pair(base + (size_t)q * (hd + 2));
        ov[q] = ld_pair(base + (size_t)q * (hd + 2) + 2 + d);
      }
    }
#pragma unroll
    for (int q = 0; q < ATT_SMAX; ++q)
      if (q < S) all_ok = all_ok && pair_ok(ml[q], ep) && pair_ok(ov[q], ep);
    if (!all_ok && !wd_tripped(wd)) {
      const long long t0 = clock64();
      do {
        all_ok = true;
#pragma unroll
        for (int q = 0; q < ATT_SMAX; ++q) {
          if (q < S) {
            if (!pair_ok(ml[q], ep)) ml[q] = ld_pair(base + (size_t)q * (hd + 2));
            if (!pair_ok(ov[q], ep)) ov[q] = ld_pair(base + (size_t)q * (hd + 2) + 2 + d);
            all_ok = all_ok && pair_ok(ml[q], ep) && pair_ok(ov[q], ep);
          }
        }
        if (!all_ok && clock64() - t0 > WATCHDOG_CYCLES) {
          wd_trip(wd, 6u);
          break;
        }
      } while (!all_ok);
    }
    float M = -INFINITY;
#pragma unroll
    for (int q = 0; q < ATT_SMAX; ++q)
      if (q < S) M = fmaxf(M, lo_f(ml[q].x));
    float lsum = 0.0f, o0 = 0.0f, o1 = 0.0f;
#pragma unroll
    for (int q = 0; q < ATT_SMAX; ++q) {
      if (q < S) {
        const float sc = fexp(lo_f(ml[q].x) - M);
        lsum = fmaf(lo_f(ml[q].y), sc, lsum);
        o0 = fmaf(lo_f(ov[q].x), sc, o0);
        o1 = fmaf(lo_f(ov[q].y), sc, o1);
      }
    }
    reinterpret_cast<float2 *>(dst_smem)[idx] = make_float2(fdiv(o0, lsum), fdiv(o1, lsum));
  }
}

// splits per head of the attention phase at context length T
__device__ __forceinline__ int att_splits(int T, int G, int H) {
  int S = (T + ATT_CHUNK - 1) / ATT_CHUNK;
  S = min(S, min(ATT_SMAX, G / H));
  return S;
}

__device__ __forceinline__ bool step_needs_logits(const DecodeParams &p, int step) {
  return p.force_logits || step >= p.n_prompt;
}

// ---- phase table ------------------------------------------------------------------------------------
// Everything about a GEMV phase that does not depend on the token is worked out once per launch and kept in
// shared memory, so that the top of a phase is a handful of LDS instead of integer divisions and branches
// (each phase executes its code exactly once: every instruction on its critical path is latency).
struct PhaseEnt {
  const float *W;     // rows of the (folded) matrix, K = E
  const float *W2;    // M_MLP: rows of c_proj^T for the same hidden units
  const float *bias;  // c2 (folded phases) or the plain bias
  const float *c1;    // row sums of the folded matrix; null when no LayerNorm precedes
  const u64 *src;     // flagged input vector
  int r0, nrows;      // rows this CTA owns
  int mode;
};

// LayerNorm statistics of a warp's register copy of the E-vector (lane holds float4 number lane + 32 j, zeros past
// the end); reference formula ops.zig:86-95: single pass E[x], E[x^2]; std = sqrt(var + eps).  Two steps so that the
// shuffle reduction can be scheduled after the dot products (nothing needs mean / rstd before the epilogue).
template <int NJ>
__device__ __forceinline__ void ln_partial(const float4 (&xs)[NJ], float &s, float &ss) {
  float s0 = 0.0f, s1 = 0.0f, q0 = 0.0f, q1 = 0.0f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    s0 += xs[j].x + xs[j].y;
    s1 += xs[j].z + xs[j].w;
    q0 = fmaf(xs[j].x, xs[j].x, q0); q1 = fmaf(xs[j].y, xs[j].y, q1);
    q0 = fmaf(xs[j].z, xs[j].z, q0); q1 = fmaf(xs[j].w, xs[j].w, q1);
  }
  s = s0 + s1;
  ss = q0 + q1;
}
__device__ __forceinline__ void ln_finish(float s, float ss, float inv_E, float &mean, float &rstd) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  mean = s * inv_E;
  rstd = rsqrtf(fmaf(ss, inv_E, -mean * mean) + 1e-5f);
}
// explicit LayerNorm output (only needed when GPT.forward has to leave ln_f(x) in state.x, main.zig:189)
template <int NJ>
__device__ __forceinline__ void ln_write(const float4 (&xs)[NJ], float mean, float rstd, const float *__restrict__ g,
                                         const float *__restrict__ b, float *out, float *raw_out, int E, int lane) {
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int i4 = lane + 32 * j;
    if (i4 < (E >> 2)) {
      const float4 gg = __ldg(reinterpret_cast<const float4 *>(g) + i4), bb = __ldg(reinterpret_cast<const float4 *>(b) + i4);
      float4 y;
      y.x = (xs[j].x - mean) * rstd * gg.x + bb.x;
      y.y = (xs[j].y - mean) * rstd * gg.y + bb.y;
      y.z = (xs[j].z - mean) * rstd * gg.z + bb.z;
      y.w = (xs[j].w - mean) * rstd * gg.w + bb.w;
      reinterpret_cast<float4 *>(out)[i4] = y;
      reinterpret_cast<float4 *>(raw_out)[i4] = xs[j];
    }
  }
}

// NJ = float4 per lane that cover one E-vector: E <= 128 NJ.
template <int NJ>
__global__ void __launch_bounds__(NTHREADS, 1) decode_persistent_kernel(const DecodeParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) u64 mbar_store[2 * MAXSLOTS];
  __shared__ unsigned wd_flag;
  const int G = gridDim.x, cta = blockIdx.x;
  const int E = p.E, Eq = p.E >> 2;  // Eq: float4 per E-vector
  const int nslot = p.nslot, slotf = p.slotf;
  const int L5 = 5 * p.L;
  Smem sm;
  sm.ring = reinterpret_cast<float *>(smem_raw);
  sm.vec = sm.ring + (size_t)nslot * slotf;
  sm.fbuf = sm.vec + 2 * E;
  sm.part = sm.fbuf + 64;
  sm.red = sm.part + NCW * p.hd;
  PhaseEnt *table = reinterpret_cast<PhaseEnt *>(sm.red + RED_FLOATS);
  sm.full0 = smem_u32(mbar_store);
  sm.empty0 = smem_u32(mbar_store + MAXSLOTS);
  sm.wd.err_global = p.err;
  sm.wd.tripped_smem = smem_u32(&wd_flag);
  const int bsz = min(NCW, nslot >> 1);  // ring units per batch: two batches always fit in the ring
  const float inv_E = 1.0f / (float)p.E;
  const int rb = 4 * bsz;                // rows per batch: at most 4 per warp

  if (threadIdx.x == 0) {
    wd_flag = 0u;
    for (int i = 0; i < nslot; ++i) {
      mbar_init(sm.full0 + 8u * i, 1);      // one arrive.expect_tx per batch, on the batch's first slot
      mbar_init(sm.empty0 + 8u * i, NCW);   // every consumer warp releases every slot of a batch
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int g = threadIdx.x; g <= L5; g += blockDim.x) {
    const int l = g / 5, ph = g - 5 * l;
    PhaseEnt e;
    e.W = e.W2 = e.bias = e.c1 = nullptr;
    e.src = nullptr;
    e.r0 = 0; e.nrows = 0; e.mode = M_ATTN;
    int N = 0, rot = 0;
    if (g == L5) {
      e.W = p.wte_f; e.bias = p.c2h; e.c1 = p.c1h; e.src = p.xnew_f; e.mode = M_LMHEAD; N = p.V;
    } else {
      const LayerDesc &ld = c_layers[l];
      rot = phase_rot(l, ph, G);
      if (ph == 0) {
        e.W = ld.wq; e.bias = ld.c2q; e.c1 = ld.c1q; e.src = p.xnew_f; e.mode = M_QKV; N = 3 * E;
      } else if (ph == 2) {
        e.W = ld.w_proj; e.bias = ld.b_proj; e.src = p.att_f; e.mode = M_RESID; N = E;
      } else if (ph == 3) {
        e.W = ld.wfc; e.W2 = ld.w2t; e.bias = ld.c2f; e.c1 = ld.c1f; e.src = p.xres_f; e.mode = M_MLP; N = 4 * E;
      } else if (ph == 4) {
        e.bias = ld.b_proj2; e.mode = M_REDUCE; N = E;
      }
    }
    if (e.mode != M_ATTN) {
      int r0, r1;
      row_range(cta, G, rot, N, r0, r1);
      e.r0 = r0; e.nrows = r1 - r0;
    }
    table[g] = e;
  }
  __syncthreads();

  const int last_step = p.first_step + p.n_steps - 1;

  if (threadIdx.x >= NCT) {
    // =============================== producer warp ===============================
    // Streams, in consumption order, the rows this CTA owns in every GEMV phase.  Unit = up to 4 rows of E floats
    // = one cp.async.bulk into one ring slot; the units of a batch all complete on the full-barrier of the batch's
    // first slot, so a consumer waits once per batch.  The MLP phase streams its c_fc rows, then its c_proj^T rows.
    const int lane = threadIdx.x - NCT;
    const uint64_t pol = policy_evict_first();
    int slot = 0;
    uint32_t parity = 0;
#pragma unroll 1
    for (int step = p.first_step; step <= last_step; ++step) {
      const int nph = L5 + (step_needs_logits(p, step) ? 1 : 0);
#pragma unroll 1
      for (int g = 0; g < nph; ++g) {
        const PhaseEnt &e = table[g];
        if (e.mode < 0) continue;
        if (e.mode == M_QKV && lane < 6) {
          // pull the layer's small vectors (folded constants + biases, 10E floats) into L2 ahead of the consumers
          const LayerDesc &ld = c_layers[g / 5];
          const float *arr = lane == 0 ? ld.c1q : lane == 1 ? ld.c2q : lane == 2 ? ld.b_proj : lane == 3 ? ld.c1f
                           : lane == 4 ? ld.c2f : ld.b_proj2;
          const int len = lane < 2 ? 3 * E : (lane == 3 || lane == 4) ? 4 * E : E;
          const int nlines = (len * 4 + 127) / 128;
          for (int i = cta; i < nlines; i += G) prefetch_l2(reinterpret_cast<const char *>(arr) + (size_t)i * 128);
        }
        if (lane == 0) {
#pragma unroll 1
          for (int pass = 0; pass < 2; ++pass) {
            const float *Wm = pass ? e.W2 : e.W;
            if (Wm == nullptr) break;
            const float *W = Wm + (size_t)e.r0 * E;
#pragma unroll 1
            for (int b0 = 0; b0 < e.nrows; b0 += rb) {
              const int nbr = min(rb, e.nrows - b0);
              const uint32_t fb = sm.full0 + 8u * slot;
              mbar_wait(sm.empty0 + 8u * slot, parity ^ 1u, sm.wd);  // first slot free => its full barrier is idle too
              mbar_expect_tx(fb, (uint32_t)nbr * (uint32_t)E * 4u);
#pragma unroll 1
              for (int r = 0; r < nbr; r += 4) {
                const int nr = min(4, nbr - r);
                if (r) mbar_wait(sm.empty0 + 8u * slot, parity ^ 1u, sm.wd);
                bulk_g2s(smem_u32(sm.ring + (size_t)slot * slotf), W + (size_t)(b0 + r) * E, (uint32_t)nr * (uint32_t)E * 4u,
                         fb, pol);
                if (++slot == nslot) { slot = 0; parity ^= 1u; }
              }
            }
          }
        }
        __syncwarp();
      }
    }
    return;
  }

  // ================================= consumer warps =================================
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Clk ck;
  ck.buf = (p.prof && cta == p.prof_cta && tid == 0) ? p.prof : nullptr;
  ck.i = 0;
#pragma unroll
  for (int k = 0; k < 12; ++k) ck.t[k] = 0u;
  u64 prev_token = 0;
  int bslot = 0;               // ring slot of the first unit of the next batch
  uint32_t fpar = 0;           // bit s: parity the next wait on full barrier s expects
  unsigned ep = p.epoch_base;  // epoch of the phase being executed; its inputs carry ep - 1
  int vsel = 0;                // which half of sm.vec the current GEMV phase reads
  // after packed_reduce<4> lane L holds the total of the warp's row number (L >> 3); lanes 0, 8, 16, 24 finish rows
  const int eidx = lane >> 3;
  const bool elane = (lane & 7) == 0;
  const int iloc = warp + NCW * eidx;  // row of a batch this lane finishes

#pragma unroll 1
  for (int step = p.first_step; step <= last_step; ++step) {
    const int pos = step, T = step + 1;  // seq_len = step + 1 (main.zig:333,337)
    u64 tok;
    if (step < p.n_prompt) tok = p.prompt ? p.prompt[step] : p.single_token;
    else if (step == p.first_step) tok = step > 0 ? __ldcg(p.tokens + step - 1) : 0ull;
    else tok = prev_token;
    if (tok >= (u64)p.V) tok = 0;  // never index the embedding out of bounds, whatever came in
    const bool want_logits = step_needs_logits(p, step);
    const int nph = L5 + (want_logits ? 1 : 0);
    u64 out_tok = tok;

#pragma unroll 1
    for (int g = 0; g < nph; ++g) {
      ++ep;
      const PhaseEnt ent = table[g];
      const bool is_head = (g == L5);
      const int mode = ent.mode;
      ck.at(0);

      if (mode == M_ATTN) {
        // ---------------- attention over the cache (ops.zig:160-171) ----------------
        const int S = att_splits(T, G, p.H);
        if (cta < p.H * S) {
          const int chunk = (T + S - 1) / S;
          if (chunk <= 4 * NCW) attention_item<4>(p, sm, g / 5, cta / S, cta % S, S, T, ep - 1, ep, ck);
          else if (chunk <= 8 * NCW) attention_item<8>(p, sm, g / 5, cta / S, cta % S, S, T, ep - 1, ep, ck);
          else attention_item<16>(p, sm, g / 5, cta / S, cta % S, S, T, ep - 1, ep, ck);
        }
        ck.at(11);
        ck.dump(1);
        continue;
      }
      if (mode == M_REDUCE) {
        // ---------------- residual 2, main.zig:142-145: x = x_mid + bias + sum over the G partial c_proj outputs,
        // for the elements [r0, r0 + ne) this CTA owns.  8 (or 16) adjacent threads read the ne adjacent words of
        // one source CTA (coalesced), a thread sums its element over its sources in a fixed order, then two
        // shuffles and a fixed-order sum over the warps: deterministic. ----------------
        const int ne = ent.nrows;
        const float *xmid = sm.vec + vsel * E;  // the MLP phase's input vector: this CTA's copy of the stream
        float bmine = 0.0f;
        if (tid < ne) bmine = __ldg(ent.bias + ent.r0 + tid);
        const int lg = ne <= 8 ? 3 : 4;              // log2(threads per source)
        const int k = tid & ((1 << lg) - 1);         // element this thread sums
        const int sgrp = tid >> lg, nsg = NCT >> lg;  // sources advance by nsg per pass
        // passes over the sources: 8 threads per source (E <= 1024: at most 8 elements per CTA, checked at create) need
        // ceil(G / 28) <= 6, 16 threads per source need ceil(G / 14) <= 12 (G <= 168); half the unrolled code at 124M / 355M
        constexpr int NP = NJ <= 8 ? 6 : 12;
        const u64 *col = p.part_f + ent.r0 + k;
        predelay(ZG_PD_P5);
        u64 w[NP];
        bool all_ok = true;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          const int src = sgrp + q * nsg;
          w[q] = (u64)(ep - 1) << 32;  // a source that does not exist contributes +0.0f
          if (src < G && k < ne) w[q] = ld_word(col + (size_t)src * E);
        }
#ifndef ZG_NOWAIT
#pragma unroll
        for (int q = 0; q < NP; ++q) all_ok = all_ok && (unsigned)(w[q] >> 32) == ep - 1;
        if (!all_ok && !wd_tripped(sm.wd)) {
          const long long t0 = clock64();
          do {
                all_ok = true;
#pragma unroll
            for (int q = 0; q < NP; ++q)
              if ((unsigned)(w[q] >> 32) != ep - 1) w[q] = ld_word(col + (size_t)(sgrp + q * nsg) * E);
#pragma unroll
            for (int q = 0; q < NP; ++q) all_ok = all_ok && (unsigned)(w[q] >> 32) == ep - 1;
            if (!all_ok && clock64() - t0 > WATCHDOG_CYCLES) {
              wd_trip(sm.wd, 4u);
              break;
            }
          } while (!all_ok);
        }
#endif
        ck.at(2);
        float tot = 0.0f;
#pragma unroll
        for (int q = 0; q < NP; ++q) tot += lo_f(w[q]);
        tot += __shfl_xor_sync(0xffffffffu, tot, 16);
        if (lg == 3) tot += __shfl_xor_sync(0xffffffffu, tot, 8);
        if (lane < (1 << lg)) sm.part[lane * 16 + warp] = tot;
        ck.at(3);
        consumer_sync();
        ck.at(4);
        if (tid < ne) {
          float v = 0.0f;
#pragma unroll
          for (int w2 = 0; w2 < NCW; ++w2) v += sm.part[tid * 16 + w2];
          st_flag(p.xnew_f + ent.r0 + tid, v + (xmid[ent.r0 + tid] + bmine), ep);
        }
        ck.at(11);
        ck.dump(4);
        continue;
      }

      vsel ^= 1;
      float *vec = sm.vec + vsel * E;              // this phase's input vector
      const float *vprev = sm.vec + (vsel ^ 1) * E;  // the previous GEMV phase's input vector
      const float4 *vec4 = reinterpret_cast<const float4 *>(vec);
      const bool has_ln = ent.c1 != nullptr;
      // ---------------- phase top: issue every load whose address is known before the activation arrives ----
      // lanes 0/8/16/24 of warp w finish rows w, w + NCW, w + 2 NCW, w + 3 NCW of a batch
      // the weights of the first batch are almost always in the ring already: test its barrier now, off the critical path
      bool wready = mbar_try(sm.full0 + 8u * bslot, (fpar >> bslot) & 1u);
      float bias_v = 0.0f, c1_v = 0.0f;
      if (elane && iloc < min(rb, ent.nrows)) {
        const int r = ent.r0 + iloc;
        if (ent.bias) bias_v = __ldg(ent.bias + r);
        if (has_ln) c1_v = __ldg(ent.c1 + r);
      }

      ck.at(1);
      if (mode == M_RESID) {  // attention CTAs are the late ones
        if (ZG_PD_P3 > 0 && cta >= p.H * att_splits(T, G, p.H)) predelay(ZG_PD_P3);
      }
      else if (mode == M_MLP) predelay(ZG_PD_P4);
      else if (g != 0) predelay(ZG_PD_P1);
      // ---------------- activation vector -> shared memory -> registers ----------------
      if (g == 0) {  // wte[token] + wpe[pos] (main.zig:179-183), recomputed by every CTA
        const float4 *te = reinterpret_cast<const float4 *>(p.wte + (size_t)tok * E);
        const float4 *pe = reinterpret_cast<const float4 *>(p.wpe + (size_t)pos * E);
#pragma unroll 1
        for (int i4 = tid; i4 < Eq; i4 += NCT) {
          const float4 a = __ldg(te + i4), b = __ldg(pe + i4);
          reinterpret_cast<float4 *>(vec)[i4] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
        }
      } else if (mode == M_RESID && att_splits(T, G, p.H) > 1) {
        gather_att_partials(vec, p.attp_f, E, att_splits(T, G, p.H), ep - 1, sm.wd);
      } else {
        gather_flagged256<(NJ <= 7 ? 1 : 2)>(vec, ent.src, E, ep - 1, sm.wd);
      }
      ck.at(2);
      consumer_sync();
      ck.at(3);
      float4 xs[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int i4 = lane + 32 * j;
        xs[j] = (i4 < Eq) ? vec4[i4] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float mean = 0.0f, rstd = 1.0f, ln_s = 0.0f, ln_ss = 0.0f;
      if (has_ln) {
        ln_partial<NJ>(xs, ln_s, ln_ss);
        if (is_head && p.write_xout && step == last_step && cta == 0 && warp == 0) {
          ln_finish(ln_s, ln_ss, inv_E, mean, rstd);
          ln_write<NJ>(xs, mean, rstd, p.lnf_g, p.lnf_b, p.xout, p.xres_out, E, lane);
        }
      }
      ck.at(4);

      // ---------------- GEMV: warp w takes rows w, w + NCW, ... of every batch ----------------
      float *kc = nullptr, *vc = nullptr;
      if (mode == M_QKV) {
        kc = c_layers[g / 5].k_cache + (size_t)pos * E;
        vc = c_layers[g / 5].v_cache + (size_t)pos * E;
      }
      float *logits = (is_head && p.store_logits && step == last_step) ? p.logits : nullptr;
      float best = -INFINITY;  // running argmax of the rows this lane finishes (lm_head only)
      unsigned best_i = 0xffffffffu;

#pragma unroll 1
      for (int b0 = 0; b0 < ent.nrows; b0 += rb) {
        const int nbr = min(rb, ent.nrows - b0);
        const int nun = (nbr + 3) >> 2;
        const bool own = elane && iloc < nbr;
        const int r = ent.r0 + b0 + iloc;
        if (b0 > 0 && own) {  // operands of later batches (the first batch's were fetched at the top of the phase)
          bias_v = ent.bias ? __ldg(ent.bias + r) : 0.0f;
          if (has_ln) c1_v = __ldg(ent.c1 + r);
        }
        if (!wready) mbar_wait(sm.full0 + 8u * bslot, (fpar >> bslot) & 1u, sm.wd);
        wready = false;
        fpar ^= 1u << bslot;
        if (b0 == 0) ck.at(5);
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = warp + NCW * u;
          if (i < nbr) {  // warp-uniform branch: a branch-free variant (every row slot loaded, invalid sums dropped)
                          // measured 20 us/token SLOWER at 124M
            int sl = bslot + (i >> 2);
            if (sl >= nslot) sl -= nslot;
            const float4 *w4 = reinterpret_cast<const float4 *>(sm.ring + (size_t)sl * slotf + (size_t)(i & 3) * E);
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
              const int i4 = min(lane + 32 * j, Eq - 1);  // clamped lanes meet a zero activation
              const float4 w = w4[i4];
              a0 = fmaf(w.x, xs[j].x, a0); a1 = fmaf(w.y, xs[j].y, a1);
              a2 = fmaf(w.z, xs[j].z, a2); a3 = fmaf(w.w, xs[j].w, a3);
            }
            acc[u] = (a0 + a1) + (a2 + a3);
          }
        }
        __syncwarp();
        if (lane < nun) {
          int sl = bslot + lane;
          if (sl >= nslot) sl -= nslot;
          mbar_arrive(sm.empty0 + 8u * (uint32_t)sl);
        }
        if (b0 == 0) ck.at(6);
        float v;  // lane L: total of the warp's row number L >> 3
        if (b0 == 0) {  // the LayerNorm statistics ride along with the first batch's reduction (same basic block)
          v = packed_reduce<4>(acc, lane);
          ln_finish(ln_s, ln_ss, inv_E, mean, rstd);
          ck.at(7);
        } else {
          v = packed_reduce<4>(acc, lane);
        }
        bslot += nun;
        if (bslot >= nslot) bslot -= nslot;
        // ---------------- epilogue: lanes 0/8/16/24 finish one row each ----------------
        if (own) {
          // folded LayerNorm: W (g (x - mean) rstd + b) + bias = rstd (W' x - mean c1) + c2
          v = has_ln ? fmaf(rstd, v - mean * c1_v, bias_v) : v + bias_v;
          if (mode == M_QKV) {  // q to the exchange, k/v to cache row `pos` (ops.zig:146-158) and to the exchange
            if (r < E) {
              st_flag(p.q_f + r, v, ep);
            } else if (r < 2 * E) {
              st_cache(kc + (r - E), v);
              st_flag(p.kvn_f + (r - E), v, ep);
            } else {
              st_cache(vc + (r - 2 * E), v);
              st_flag(p.kvn_f + E + (r - 2 * E), v, ep);
            }
          } else if (mode == M_RESID) {  // residual 1, main.zig:136-139: the stream is this CTA's own copy
            st_flag(p.xres_f + r, v + vprev[r], ep);
          } else if (mode == M_MLP) {  // main.zig:79-80
            sm.fbuf[b0 + iloc] = gelu_dec(v);
          } else {  // tied lm_head (main.zig:193) + running argmax; this lane sees increasing r, so strict >
            if (logits) logits[r] = v;
            if (v > best) { best = v; best_i = (unsigned)r; }
          }
        }
      }

      ck.at(8);
      if (mode == M_MLP) {
        // ---------------- mlp c_proj, main.zig:81: out += f_j * c_proj^T[j, :] over the hidden units j this CTA owns.
        // Thread t accumulates output float4 t (and t + 224 for wide models); no shuffles. ----------------
        // test the barrier of the first c_proj^T batch before the CTA sync
        bool w2ready = mbar_try(sm.full0 + 8u * bslot, (fpar >> bslot) & 1u);
        consumer_sync();  // fbuf complete
        ck.at(9);
        constexpr int NK = (NJ * 32 > NCT) ? 2 : 1;  // output float4 per thread
        float4 o4[NK];
#pragma unroll
        for (int k = 0; k < NK; ++k) o4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int b0 = 0; b0 < ent.nrows; b0 += rb) {
          const int nbr = min(rb, ent.nrows - b0);
          const int nun = (nbr + 3) >> 2;
          if (!w2ready) mbar_wait(sm.full0 + 8u * bslot, (fpar >> bslot) & 1u, sm.wd);
          w2ready = false;
          fpar ^= 1u << bslot;
          // One ring unit (4 rows of c_proj^T) per trip: one broadcast LDS.128 for the unit's four GELU factors, four
          // LDS.128 for this thread's column of the four rows, 16 FMAs.  Rows past the batch re-read row 0 of their unit
          // (always valid) with a zero factor.  Two trips are unrolled together so that the loads of unit q + 1 are in
          // flight under the FMAs of unit q; the slot index wraps with one compare per unit (uniform registers).
#pragma unroll
          for (int k = 0; k < NK; ++k) {
            const int i4 = min(tid + k * NCT, Eq - 1);
            const bool live = tid + k * NCT < Eq;
            int sl = bslot;
#pragma unroll 2
            for (int q = 0; q < nun; ++q) {
              const float4 *w4 = reinterpret_cast<const float4 *>(sm.ring + (size_t)sl * slotf) + i4;
              const int nv = live ? nbr - 4 * q : 0;  // valid rows of this unit (>= 1 for a live thread)
              const float4 f4 = *reinterpret_cast<const float4 *>(sm.fbuf + b0 + 4 * q);
              const float4 w0 = w4[0];
              const float4 w1 = w4[nv > 1 ? Eq : 0];
              const float4 w2 = w4[nv > 2 ? 2 * Eq : 0];
              const float4 w3 = w4[nv > 3 ? 3 * Eq : 0];
              const float f0 = nv > 0 ? f4.x : 0.0f, f1 = nv > 1 ? f4.y : 0.0f, f2 = nv > 2 ? f4.z : 0.0f, f3 = nv > 3 ? f4.w : 0.0f;
              o4[k].x = fmaf(f0, w0.x, o4[k].x); o4[k].y = fmaf(f0, w0.y, o4[k].y);
              o4[k].z = fmaf(f0, w0.z, o4[k].z); o4[k].w = fmaf(f0, w0.w, o4[k].w);
              o4[k].x = fmaf(f1, w1.x, o4[k].x); o4[k].y = fmaf(f1, w1.y, o4[k].y);
              o4[k].z = fmaf(f1, w1.z, o4[k].z); o4[k].w = fmaf(f1, w1.w, o4[k].w);
              o4[k].x = fmaf(f2, w2.x, o4[k].x); o4[k].y = fmaf(f2, w2.y, o4[k].y);
              o4[k].z = fmaf(f2, w2.z, o4[k].z); o4[k].w = fmaf(f2, w2.w, o4[k].w);
              o4[k].x = fmaf(f3, w3.x, o4[k].x); o4[k].y = fmaf(f3, w3.y, o4[k].y);
              o4[k].z = fmaf(f3, w3.z, o4[k].z); o4[k].w = fmaf(f3, w3.w, o4[k].w);
              if (++sl == nslot) sl = 0;
            }
          }
          __syncwarp();
          if (lane < nun) {
            int sl = bslot + lane;
            if (sl >= nslot) sl -= nslot;
            mbar_arrive(sm.empty0 + 8u * (uint32_t)sl);
          }
          bslot += nun;
          if (bslot >= nslot) bslot -= nslot;
        }
        ck.at(10);
        u64 *mine = p.part_f + (size_t)cta * E;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
          const int i4 = tid + k * NCT;
          if (i4 < Eq) {
            st_flag2(mine + 4 * i4, o4[k].x, o4[k].y, ep);
            st_flag2(mine + 4 * i4 + 2, o4[k].z, o4[k].w, ep);
          }
        }
      }
      ck.at(11);
      ck.dump(is_head ? 5 : mode == M_QKV ? 0 : mode == M_RESID ? 2 : 3);

      if (is_head) {
        // argmax (value desc, index asc): warp -> CTA -> one flagged partial per CTA, then every CTA reduces the
        // G partials itself: the next step's embedding needs the token everywhere.
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, best, o);
          const unsigned oi = __shfl_xor_sync(0xffffffffu, best_i, o);
          if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
        }
        if (lane == 0) {
          sm.red[RED_B + warp] = best;
          sm.red[RED_I + warp] = __uint_as_float(best_i);
        }
        consumer_sync();
        if (tid < 32) {
          best = (lane < NCW) ? sm.red[RED_B + lane] : -INFINITY;
          best_i = (lane < NCW) ? __float_as_uint(sm.red[RED_I + lane]) : 0xffffffffu;
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const unsigned oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
          }
          if (tid == 0) st_flag2(p.amax_f + 2 * cta, best, __uint_as_float(best_i), ep);
          predelay(ZG_PD_AMAX);
          float bv = -INFINITY;
          unsigned bi = 0xffffffffu;
          for (int i = tid; i < G; i += 32) {
            ulonglong2 w = ld_pair(p.amax_f + 2 * i);
            if (!pair_ok(w, ep)) w = spin_pair(p.amax_f + 2 * i, ep, sm.wd);
            const float ov = lo_f(w.x);
            const unsigned oi = (unsigned)w.y;
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const unsigned oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
          }
          if (tid == 0) sm.red[RED_TOK] = __uint_as_float(bi);
        }
        consumer_sync();
        const u64 amax = (u64)__float_as_uint(sm.red[RED_TOK]);
        if (cta == 0 && tid == 0) {
          *p.last_token = amax;
          if (p.result_host && step == last_step) {  // the host spins on the sequence number (zg_engine_sample_greedy)
            p.result_host[0] = amax;
            __threadfence_system();
            p.result_host[1] = p.result_seq;
          }
        }
        if (step >= p.n_prompt) out_tok = amax;  // generate(): main.zig:335-338
        consumer_sync();
      }
    }

    if (!want_logits && p.write_xout && step == last_step && cta == 0) {
      // GPT.forward(compute_logits = false) still leaves ln_f(x) in state.x (main.zig:189)
      vsel ^= 1;
      float *vec = sm.vec + vsel * E;
      gather_flagged256<(NJ <= 7 ? 1 : 2)>(vec, p.xnew_f, E, ep, sm.wd);
      consumer_sync();
      if (warp == 0) {
        float4 xs[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int i4 = lane + 32 * j;
          xs[j] = (i4 < Eq) ? reinterpret_cast<const float4 *>(vec)[i4] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float mean, rstd, ln_s, ln_ss;
        ln_partial<NJ>(xs, ln_s, ln_ss);
        ln_finish(ln_s, ln_ss, inv_E, mean, rstd);
        ln_write<NJ>(xs, mean, rstd, p.lnf_g, p.lnf_b, p.xout, p.xres_out, E, lane);
      }
    }
    if (cta == 0 && tid == 0) {
      p.tokens[step] = out_tok;
      if (p.tokens_host) p.tokens_host[step] = out_tok;
    }
    prev_token = out_tok;
  }
}

typedef void (*decode_kernel_t)(const DecodeParams);
// smallest instantiation whose register copy covers E (E <= 128 NJ)
static decode_kernel_t decode_kernel_for(int E) {
  if (E <= 256) return decode_persistent_kernel<2>;
  if (E <= 768) return decode_persistent_kernel<6>;
  if (E <= 1024) return decode_persistent_kernel<8>;
  if (E <= 1280) return decode_persistent_kernel<10>;
  return decode_persistent_kernel<13>;
}

// ---- start-up kernels: derived weight copies --------------------------------------------------------
// W'[r,k] = W[r,k] g[k];  c1[r] = sum_k W'[r,k];  c2[r] = sum_k W[r,k] b[k] + bias[r].  One warp per row.
__global__ void fold_ln_kernel(const float *__restrict__ W, const float *__restrict__ g, const float *__restrict__ b,
                               const float *__restrict__ bias, size_t N, int K, float *__restrict__ Wf,
                               float *__restrict__ c1, float *__restrict__ c2) {
  const size_t r = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= N) return;
  float s1 = 0.0f, s2 = 0.0f;
  for (int k = lane; k < K; k += 32) {
    const float w = W[r * K + k];
    const float wf = w * g[k];
    Wf[r * K + k] = wf;
    s1 += wf;
    s2 = fmaf(w, b[k], s2);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) {
    c1[r] = s1;
    c2[r] = s2 + (bias ? bias[r] : 0.0f);
  }
}
// out[k, n] = in[n, k] for in [N, K]
__global__ void transpose_weight_kernel(const float *__restrict__ in, int N, int K, float *__restrict__ out) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    tile[i][threadIdx.x] = (n < N && k < K) ? in[(size_t)n * K + k] : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    if (k < K && n < N) out[(size_t)k * N + n] = tile[threadIdx.x][i];
  }
}

}  // namespace zg

// =================================================================================================
// host side
// =================================================================================================
using namespace zg;

struct zg_engine {
  zg_config cfg;
  zg_state state;
  DecodeParams base;
  LayerDesc *layers_host;
  u64 *exchange_dev;  // all flagged buffers, one allocation
  float *derived_dev;  // folded / transposed weight copies, one allocation
  u64 *prompt_dev;
  u64 *tokens_dev;
  u64 *tokens_host;  // pinned, mapped
  u64 *tokens_host_devptr;
  u64 *last_token_dev;
  u64 result_seq;   // sequence number of the last zg_engine_sample_greedy answer (pinned slot tokens_host[C .. C+1])
  int sample_slot;  // zg_engine_forward is being called by zg_engine_sample_greedy
  u64 *prof_dev;
  unsigned *err_dev;
  void *samp_dev;  // {temp, seed, sequence} of the sampling generate loop (zg_engine_generate_sample)
  unsigned epoch_count;  // host mirror of the phase epoch (monotonic across launches)
  int grid;
  size_t smem_bytes;
  int n_prompt;
  int prof_enabled;
};

static zg_engine *g_table_owner = nullptr;  // whose layer table currently sits in __constant__ memory ...
static unsigned g_table_gen = 0;            // ... of the device selected by this zg_init generation

static size_t engine_smem_bytes(const zg_config &c, int nslot) {
  const size_t E = c.n_embed, hd = E / c.n_heads;
  const size_t floats = (size_t)nslot * 4 * E + 2 * E + 64 + NCW * hd + RED_FLOATS;
  return floats * sizeof(float) + (5 * c.n_layer + 1) * sizeof(PhaseEnt);
}

extern "C" {

zg_engine *zg_engine_create(const zg_gpt *gpt, const zg_state *state) {
  if (!require_ready("zg_engine_create")) return nullptr;
  Context &c = ctx();
  const zg_config &cfg = gpt->config;
  const size_t E = cfg.n_embed, V = cfg.vocab_size, L = cfg.n_layer;
  if (E % 8 != 0 || cfg.n_heads * 64 != E || E > 1664 || L > (size_t)MAX_LAYERS) {
    set_error(1, "zg_engine_create: needs head_dim 64, n_embed % 8 == 0, n_embed <= 1664, n_layer <= 64", __FILE__, __LINE__);
    return nullptr;
  }
  zg_engine *e = (zg_engine *)calloc(1, sizeof(zg_engine));
  if (!e) return nullptr;
  e->cfg = cfg;
  e->state = *state;
  e->grid = c.sm_count;
  if ((size_t)e->grid < cfg.n_heads || e->grid > 168 || (size_t)e->grid > E ||
      (4 * E + (size_t)e->grid - 1) / (size_t)e->grid > 64 || (E + (size_t)e->grid - 1) / (size_t)e->grid > MAXNE ||
      (E <= 1024 && (E + (size_t)e->grid - 1) / (size_t)e->grid > 8)) {
    set_error(1, "zg_engine_create: needs n_heads <= SMs <= min(168, n_embed), <= 64 hidden units and <= 16 stream elements per SM",
              __FILE__, __LINE__);
    free(e);
    return nullptr;
  }

  int max_smem = 0;
  ZG_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c.device));
  int nslot = MAXSLOTS;
  while (nslot > 2 && engine_smem_bytes(cfg, nslot) + 1024 > (size_t)max_smem) --nslot;
  if (engine_smem_bytes(cfg, nslot) + 1024 > (size_t)max_smem) {
    set_error(1, "zg_engine_create: model too wide for the shared-memory ring", __FILE__, __LINE__);
    free(e);
    return nullptr;
  }
  e->smem_bytes = engine_smem_bytes(cfg, nslot);
  const decode_kernel_t kern = decode_kernel_for((int)E);
  ZG_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes));
  int per_sm = 0;
  ZG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)kern, NTHREADS, e->smem_bytes));
  if (per_sm < 1) {
    set_error(1, "zg_engine_create: persistent kernel does not fit on an SM", __FILE__, __LINE__);
    free(e);
    return nullptr;
  }

  // derived weight copies (start-up only): LayerNorm-folded c_attn / c_fc / lm_head, transposed mlp c_proj
  const size_t per_layer = 3 * E * E + 2 * 3 * E + 4 * E * E + 2 * 4 * E + 4 * E * E;
  const size_t n_derived = L * per_layer + V * E + 2 * V;
  e->derived_dev = (float *)zg_alloc(n_derived * sizeof(float));
  if (!e->derived_dev) {
    free(e);
    return nullptr;
  }
  e->layers_host = (LayerDesc *)calloc(MAX_LAYERS, sizeof(LayerDesc));
  float *d = e->derived_dev;
  const int fold_threads = 256, rows_per_block = fold_threads / 32;
  for (size_t l = 0; l < L; ++l) {
    const zg_block &b = gpt->h[l];
    float *wq = d; d += 3 * E * E;
    float *c1q = d; d += 3 * E;
    float *c2q = d; d += 3 * E;
    float *wfc = d; d += 4 * E * E;
    float *c1f = d; d += 4 * E;
    float *c2f = d; d += 4 * E;
    float *w2t = d; d += 4 * E * E;
    fold_ln_kernel<<<(unsigned)((3 * E + rows_per_block - 1) / rows_per_block), fold_threads, 0, c.stream>>>(
        b.attn.c_attn.weight, b.ln_1.weight, b.ln_1.bias, b.attn.c_attn.bias, 3 * E, (int)E, wq, c1q, c2q);
    fold_ln_kernel<<<(unsigned)((4 * E + rows_per_block - 1) / rows_per_block), fold_threads, 0, c.stream>>>(
        b.mlp.c_fc.weight, b.ln_2.weight, b.ln_2.bias, b.mlp.c_fc.bias, 4 * E, (int)E, wfc, c1f, c2f);
    // mlp c_proj.weight is [E, 4E] (out, in); the fused MLP phase walks it by hidden unit: [4E, E]
    transpose_weight_kernel<<<dim3((unsigned)((4 * E + 31) / 32), (unsigned)((E + 31) / 32)), dim3(32, 8), 0, c.stream>>>(
        b.mlp.c_proj.weight, (int)E, (int)(4 * E), w2t);
    e->layers_host[l] = LayerDesc{wq, c1q, c2q, b.attn.c_proj.weight, b.attn.c_proj.bias, wfc, c1f, c2f, w2t,
                                  b.mlp.c_proj.bias, b.k_cache, b.v_cache};
  }
  float *wte_f = d; d += V * E;
  float *c1h = d; d += V;
  float *c2h = d; d += V;
  fold_ln_kernel<<<(unsigned)((V + rows_per_block - 1) / rows_per_block), fold_threads, 0, c.stream>>>(
      gpt->lm_head.weight, gpt->ln_f.weight, gpt->ln_f.bias, gpt->lm_head.bias, V, (int)E, wte_f, c1h, c2h);
  ZG_CUDA(cudaGetLastError());

  const size_t C = cfg.context_size, hd = 64;
  const size_t n_attp = cfg.n_heads * (size_t)ATT_SMAX * (hd + 2);
  const size_t n_amax = (2 * (size_t)e->grid + 3) & ~(size_t)3;  // keeps the vectors behind it 32-byte aligned (256-bit loads)
  const size_t n_exchange = E + E + 2 * E + E + n_amax + E + n_attp + (size_t)e->grid * E;
  e->exchange_dev = (u64 *)zg_alloc(n_exchange * 8);
  e->prompt_dev = (u64 *)zg_alloc(C * 8);
  e->tokens_dev = (u64 *)zg_alloc(C * 8);
  e->last_token_dev = (u64 *)zg_alloc(8);
  e->prof_dev = (u64 *)zg_alloc((2 * PROF_MAX + 4) * 8);
  e->err_dev = (unsigned *)zg_alloc(256);
  e->samp_dev = zg_alloc(64);
  ZG_CUDA(cudaHostAlloc(&e->tokens_host, (C + 2) * 8, cudaHostAllocMapped));  // + {token, sequence number} of GPT.sample
  note_alloc();
  ZG_CUDA(cudaHostGetDevicePointer((void **)&e->tokens_host_devptr, e->tokens_host, 0));
  if (zg_last_error()) {
    free(e);
    return nullptr;
  }
  memset(e->tokens_host, 0xff, (C + 2) * 8);
  zg_memset(e->exchange_dev, 0, n_exchange * 8);  // epoch 0 everywhere; the first phase of the first launch is epoch 1
  zg_memset(e->err_dev, 0, 256);
  zg_memset(e->tokens_dev, 0, C * 8);
  zg_memset(e->prof_dev, 0, (2 * PROF_MAX + 4) * 8);
  e->epoch_count = 0;

  DecodeParams &p = e->base;
  memset(&p, 0, sizeof(p));
  p.E = (int)E; p.H = (int)cfg.n_heads; p.hd = (int)hd; p.L = (int)L; p.V = (int)V; p.C = (int)C;
  p.nslot = nslot; p.slotf = 4 * (int)E;
  p.wte = gpt->wte.weight; p.wpe = gpt->wpe.weight; p.lnf_g = gpt->ln_f.weight; p.lnf_b = gpt->ln_f.bias;
  p.wte_f = wte_f; p.c1h = c1h; p.c2h = c2h;
  u64 *x = e->exchange_dev;
  p.xres_f = x; x += E;
  p.q_f = x; x += E;
  p.kvn_f = x; x += 2 * E;
  p.att_f = x; x += E;
  p.amax_f = x; x += n_amax;
  p.xnew_f = x; x += E;
  p.attp_f = x; x += n_attp;
  p.part_f = x;
  p.xres_out = state->o; p.xout = state->x; p.logits = state->logits;
  p.err = e->err_dev;
  p.tokens = e->tokens_dev; p.tokens_host = e->tokens_host_devptr; p.last_token = e->last_token_dev;
  zg_sync();
  return zg_last_error() ? (free(e), nullptr) : e;
}

void zg_engine_destroy(zg_engine *e) {
  if (!e) return;
  zg_sync();
  if (g_table_owner == e) g_table_owner = nullptr;
  zg_free(e->exchange_dev); zg_free(e->derived_dev); zg_free(e->prompt_dev); zg_free(e->tokens_dev); zg_free(e->last_token_dev);
  zg_free(e->prof_dev); zg_free(e->err_dev); zg_free(e->samp_dev);
  cudaFreeHost(e->tokens_host);
  free(e->layers_host);
  free(e);
}

}  // extern "C"

// number of phases a launch executes (the exchange epoch is monotonic across launches)
static unsigned phases_for(const zg_engine *e, const DecodeParams &p) {
  unsigned n = 0;
  for (int s = p.first_step; s < p.first_step + p.n_steps; ++s)
    n += 5u * (unsigned)e->cfg.n_layer + ((p.force_logits || s >= p.n_prompt) ? 1u : 0u);
  return n;
}

static void engine_launch(zg_engine *e, DecodeParams &p) {
  if (p.n_steps <= 0) return;
  if (p.first_step < 0 || p.first_step + p.n_steps > (int)e->cfg.context_size) {
    set_error(1, "decode engine: step range exceeds context_size", __FILE__, __LINE__);
    return;
  }
  if (g_table_owner != e || g_table_gen != ctx().generation) {  // stream-ordered, so a launch in flight keeps the table it was given
    ZG_CUDA(cudaMemcpyToSymbolAsync(c_layers, e->layers_host, sizeof(LayerDesc) * MAX_LAYERS, 0,
                                    cudaMemcpyHostToDevice, ctx().stream));
    g_table_owner = e;
    g_table_gen = ctx().generation;
  }
  p.epoch_base = e->epoch_count;
  p.prof = e->prof_enabled ? e->prof_dev : nullptr;
  p.prof_cta = e->prof_enabled - 1;
  e->epoch_count += phases_for(e, p);
  void *args[] = {(void *)&p};
  ZG_CUDA(cudaLaunchCooperativeKernel((const void *)decode_kernel_for((int)e->cfg.n_embed), dim3(e->grid), dim3(NTHREADS),
                                      args, e->smem_bytes, ctx().stream));
  ctx().launches++;
}

// after a synchronisation: did the in-kernel watchdog fire (a wait exceeded ~2 s)?
static int engine_check_watchdog(zg_engine *e) {
  unsigned w = 0;
  ZG_CUDA(cudaMemcpyAsync(&w, e->err_dev, sizeof(w), cudaMemcpyDeviceToHost, ctx().stream));
  ZG_CUDA(cudaStreamSynchronize(ctx().stream));
  if (w != 0) {
    set_error(1, w == 2 ? "decode engine watchdog: mbarrier wait timed out"
                        : "decode engine watchdog: flagged-exchange wait timed out", __FILE__, __LINE__);
    return 1;
  }
  return zg_last_error();
}

extern "C" {

void zg_engine_forward(zg_engine *e, size_t seq_len, size_t token, int compute_logits) {
  if (!require_ready("zg_engine_forward")) return;
  if (seq_len == 0 || seq_len > e->cfg.context_size || token >= e->cfg.vocab_size) {  // same contract as zg_gpt_forward
    set_error(1, "zg_engine_forward: need 1 <= seq_len <= context_size and token < vocab_size", __FILE__, __LINE__);
    return;
  }
  const size_t step = seq_len - 1;
  DecodeParams p = e->base;
  p.prompt = nullptr;  // the forced token rides in the kernel parameters
  p.single_token = token;
  p.n_prompt = (int)step + 1;
  p.first_step = (int)step;
  p.n_steps = 1;
  p.force_logits = compute_logits ? 1 : 0;
  p.store_logits = compute_logits ? 1 : 0;
  p.write_xout = 1;
  if (e->sample_slot) {
    p.result_host = e->tokens_host_devptr + e->cfg.context_size;
    p.result_seq = e->result_seq;
  }
  engine_launch(e, p);
}

// GPT.sample with temp -> 0 (main.zig:198-207): host token in, host token out.  The kernel writes {token, sequence number}
// into pinned host memory and the host spins on the sequence number: no copy, no stream synchronise on the per-token
// path (~10 us of a 150 us step).  If the answer does not show up the slow path synchronises and reads the watchdog.
size_t zg_engine_sample_greedy(zg_engine *e, size_t seq_len, size_t token) {
  if (!require_ready("zg_engine_sample_greedy")) return (size_t)-1;
  Context &c = ctx();
  const size_t C = e->cfg.context_size;
  volatile u64 *slot = e->tokens_host + C;
  e->result_seq = (e->result_seq + 1) & 0x7fffffffffffffffull;  // never the 0xff.. fill pattern
  e->sample_slot = 1;
  zg_engine_forward(e, seq_len, token, 1);
  e->sample_slot = 0;
  if (zg_last_error()) return (size_t)-1;
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned spins = 0; slot[1] != e->result_seq; ++spins) {
    __builtin_ia32_pause();
    if ((spins & 0xffffu) == 0xffffu && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(3)) break;
  }
  if (slot[1] != e->result_seq) {  // slow path: the launch failed or the watchdog fired
    ZG_CUDA(cudaStreamSynchronize(c.stream));
    if (engine_check_watchdog(e) || slot[1] != e->result_seq) return (size_t)-1;
  }
  return (size_t)slot[0];
}

size_t zg_engine_sample(zg_engine *e, size_t seq_len, float temp, size_t token, double u) {
  if (!require_ready("zg_engine_sample")) return (size_t)-1;
  Context &c = ctx();
  zg_engine_forward(e, seq_len, token, 1);
  if (zg_last_error()) return (size_t)-1;
  launch_softmax_temp(e->state.logits, e->cfg.vocab_size, temp);  // main.zig:200-203
  launch_weighted_index(e->state.logits, e->cfg.vocab_size, (float)u, c.token_slot);
  ZG_CUDA(cudaMemcpyAsync(c.token_slot_host, c.token_slot, 8, cudaMemcpyDeviceToHost, c.stream));
  ZG_CUDA(cudaStreamSynchronize(c.stream));
  return (size_t)c.token_slot_host[0];
}

int zg_engine_set_prompt(zg_engine *e, const size_t *inputs, size_t n_inputs) {
  if (!require_ready("zg_engine_set_prompt")) return 1;
  if (n_inputs > e->cfg.context_size) return 1;
  e->n_prompt = (int)n_inputs;
  if (n_inputs) ZG_CUDA(cudaMemcpyAsync(e->prompt_dev, inputs, n_inputs * 8, cudaMemcpyHostToDevice, ctx().stream));
  return zg_last_error();
}

void zg_engine_run_steps(zg_engine *e, size_t first_step, size_t n_steps) {
  if (!require_ready("zg_engine_run_steps")) return;
  DecodeParams p = e->base;
  p.prompt = e->prompt_dev;
  p.n_prompt = e->n_prompt;
  p.first_step = (int)first_step;
  p.n_steps = (int)n_steps;
  engine_launch(e, p);
}

int zg_engine_read_tokens(zg_engine *e, size_t first_step, size_t n_steps, size_t *out_tokens) {
  if (!require_ready("zg_engine_read_tokens")) return 1;
  if (zg_download(out_tokens, e->tokens_dev + first_step, n_steps * 8)) return zg_last_error();
  return engine_check_watchdog(e);
}

int zg_engine_generate_greedy(zg_engine *e, const size_t *inputs, size_t n_inputs, size_t n_total, size_t *out_tokens) {
  if (!require_ready("zg_engine_generate_greedy")) return 1;
  if (n_total > e->cfg.context_size || n_inputs > n_total) return 1;
  if (zg_engine_set_prompt(e, inputs, n_inputs)) return zg_last_error();
  zg_engine_run_steps(e, 0, n_total);
  // tokens were streamed into the pinned ring as they were produced; one wait for the whole call
  ZG_CUDA(cudaStreamSynchronize(ctx().stream));
  for (size_t i = 0; i < n_total; ++i) out_tokens[i] = (size_t)e->tokens_host[i];
  return engine_check_watchdog(e);
}

// generate() with temperature sampling (main.zig:322-342 with GPT.sample, :198-207), device resident: every sampling step is
// one persistent-kernel launch that leaves the logits in state.logits plus one sampling kernel (logits / temp, softmax,
// inverse-CDF draw with u = philox_uniform(seed, step, sequence)) that writes the token where the next launch reads it.
// No host round trip per token; the host waits once.  Greedy semantics otherwise: prompt tokens are forwarded without
// logits, the last prompt token is forwarded twice.
int zg_engine_generate_sample(zg_engine *e, const size_t *inputs, size_t n_inputs, size_t n_total, float temp,
                              unsigned long long seed, unsigned long long sequence, size_t *out_tokens) {
  if (!require_ready("zg_engine_generate_sample")) return 1;
  if (n_inputs == 0 || n_total > e->cfg.context_size || n_inputs > n_total || !(temp > 0.0f)) {
    set_error(1, "zg_engine_generate_sample: need 1 <= n_inputs <= n_total <= context_size and temp > 0", __FILE__, __LINE__);
    return 1;
  }
  if (zg_engine_set_prompt(e, inputs, n_inputs)) return zg_last_error();
  struct { float temp; unsigned long long seed, seq_base; } sp = {temp, seed, sequence};
  ZG_CUDA(cudaMemcpyAsync(e->samp_dev, &sp, sizeof(sp), cudaMemcpyHostToDevice, ctx().stream));
  zg_engine_run_steps(e, 0, n_inputs);
  for (size_t s = n_inputs; s < n_total; ++s) {
    DecodeParams p = e->base;
    p.prompt = e->prompt_dev;
    p.n_prompt = e->n_prompt;
    p.first_step = (int)s;
    p.n_steps = 1;
    p.store_logits = 1;
    engine_launch(e, p);
    launch_sample_rows(e->state.logits, e->cfg.vocab_size, (int)e->cfg.vocab_size, e->samp_dev, nullptr, (int)s,
                       e->tokens_dev + s, nullptr, 1, e->tokens_host_devptr + s);
  }
  ZG_CUDA(cudaStreamSynchronize(ctx().stream));
  for (size_t i = 0; i < n_total; ++i) out_tokens[i] = (size_t)e->tokens_host[i];
  return engine_check_watchdog(e);
}

size_t zg_engine_read_profile(zg_engine *e, unsigned long long *out, size_t max_entries) {
  if (!require_ready("zg_engine_read_profile")) return 0;
  if (out == nullptr) {  // toggle: NULL + n enables the timeline of CTA n - 1 (n != 0) or disables profiling (n == 0)
    e->prof_enabled = (int)max_entries;
    return 0;
  }
  // out receives (tag, SM cycles) pairs; returns the number of pairs
  u64 *tmp = (u64 *)malloc((2 * PROF_MAX + 4) * 8);
  zg_download(tmp, e->prof_dev, (2 * PROF_MAX + 4) * 8);
  size_t n = (size_t)tmp[2 * PROF_MAX];
  if (n > PROF_MAX) n = PROF_MAX;
  if (2 * n > max_entries) n = max_entries / 2;
  memcpy(out, tmp, 2 * n * 8);
  free(tmp);
  return n;
}

}  // extern "C"
