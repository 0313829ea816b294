// zg_decode.cu -- the fused batch-1 decode engine: GPT.forward / GPT.sample / generate
// (src/main.zig:178-207, 322-342) as ONE persistent cooperative kernel.
//
// Design (B200, 148 SMs, HBM-bound: 495 MB of fp32 weights per token at 124M):
//   * one CTA per SM, 8 consumer warps + 1 producer warp;
//   * the producer warp streams this CTA's share of every weight matrix, in execution order, through a
//     shared-memory ring with cp.async.bulk (UBLKCP) + mbarrier complete_tx, L2 evict-first.  The stream does
//     not depend on activations, so it runs ahead across layer phases and tokens;
//   * consumers take weight rows from the ring (one warp per ring unit, conflict-free 128-bit LDS), dot them
//     with the activation vector held in shared memory, reduce with warp shuffles, and apply the fused
//     epilogue (bias, GELU, residual add, KV-cache append, running argmax);
//   * five phases per layer
//        P1 LN1 + c_attn (+ K/V append)   ops.zig:143-158, main.zig:121-123
//        P2 attention over the time-major cache (flash-decoding splits when T is long)  ops.zig:160-171
//        P3 attn c_proj + residual         ops.zig:172, main.zig:136-139
//        P4 LN2 + c_fc + GELU              main.zig:140, :79-80
//        P5 mlp c_proj + residual          main.zig:81, :142-145
//     then ln_f + tied lm_head + argmax (main.zig:189-194);
//   * NO grid barrier and NO memory fence between phases.  Every activation word crosses SMs as one 64-bit
//     store {epoch : 32 | fp32 bits : 32}; a consumer gathers the vector it needs with 128-bit relaxed loads
//     and spins until every word carries the epoch of the phase that produces it (the NCCL "LL" idea: the
//     flag travels inside the datum, so there is nothing to fence).  A measured grid barrier costs ~2.6 us per
//     phase inside this kernel (0.9 us release fence + 0.7 us poll + 0.5 us acquire fence + skew); the
//     flagged gather costs one L2 round trip;
//   * the layer table lives in __constant__ memory, so no phase starts with a dependent global load;
//   * the token loop of generate() runs inside the kernel; each token is written to device memory and to a
//     pinned host ring, so the host only waits once per call.
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>

#include "zg_common.cuh"
#include "zg_ptx.cuh"

namespace zg {
void launch_softmax_temp(float *x, size_t n, float temp);
void launch_weighted_index(const float *p, size_t n, float u, unsigned long long *out);

typedef unsigned long long u64;

constexpr int NTHREADS = NCT + 32;  // + producer warp
constexpr int MAXSLOTS = 32;
constexpr int MAX_LAYERS = 64;
constexpr int ATT_CHUNK = 128;  // KV rows per attention work item before splitting
constexpr int PROF_MAX = 16384;
constexpr int LNR = 7;          // LayerNorm elements per thread: E <= 7 * 256
constexpr int GB = 6;           // flagged pairs a thread keeps in flight while gathering

struct LayerDesc {
  const float *ln1_g, *ln1_b, *w_attn, *b_attn, *w_proj, *b_proj, *ln2_g, *ln2_b, *w_fc, *b_fc, *w_proj2, *b_proj2;
  float *k_cache, *v_cache;
};
__constant__ LayerDesc c_layers[MAX_LAYERS];

struct DecodeParams {
  int E, H, hd, L, V, C;
  int nslot, slotf;  // ring geometry: slotf = 4E floats per slot
  const float *wte, *wpe, *lnf_g, *lnf_b;
  // flagged exchange buffers: word = {epoch << 32 | fp32 bits}
  u64 *xres_f;  // [E]   residual stream
  u64 *q_f;     // [E]   query of the current token
  u64 *kvn_f;   // [2E]  K row then V row of the current token (the cache gets the same values, unflagged)
  u64 *att_f;   // [E]   attention output
  u64 *f_f;     // [4E]  GELU(c_fc)
  u64 *amax_f;  // [2G]  per-CTA argmax partial: value word, index word
  unsigned epoch_base;
  float *xres_out;  // [E] state.o: the reference leaves the pre-ln_f stream there (main.zig:116-118)
  float *xout;      // [E] state.x: ln_f output
  float *logits;    // [V] state.logits
  float *att_part;       // [H][S][hd+2] flash-decoding partials
  unsigned *head_count;  // [H] arrival counters for the split combine
  unsigned *err;         // sticky watchdog word
  const u64 *prompt;     // device, n_prompt entries; null => `single_token` is the forced token
  u64 single_token;
  int n_prompt;
  u64 *tokens;                 // device [C]: token forwarded/sampled at every step
  volatile u64 *tokens_host;   // pinned host ring [C]
  u64 *last_token;             // device: argmax of the last logits computed
  int first_step, n_steps;
  int force_logits;  // compute logits + argmax on every step (GPT.forward(compute_logits=true) on a prompt step)
  int store_logits;  // also write the logits vector to global memory
  int write_xout;    // write ln_f(x) to xout (state.x) and the stream to xres_out on the last step
  u64 *prof;
  int dbg;
};

// optional timeline of CTA 0 / thread 0: (tag, %globaltimer) pairs
struct Prof {
  u64 *buf;
  int i;
  bool fine;
  __device__ __forceinline__ void fmark(int tag) {
    if (fine) mark(tag);
  }
  __device__ __forceinline__ void mark(int tag) {
    if (buf != nullptr) record(tag);
  }
  __device__ __noinline__ void record(int tag) {
    if (i < PROF_MAX) {
      buf[2 * i] = (u64)tag;
      buf[2 * i + 1] = globaltimer();
    }
    ++i;
  }
};

// ---- flag-in-data exchange ------------------------------------------------------------------------
__device__ __forceinline__ void st_flag(u64 *p, float v, unsigned ep) {
  const u64 w = ((u64)ep << 32) | (u64)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
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
__device__ __forceinline__ bool pair_ok(const ulonglong2 &v, unsigned ep) {
  return (unsigned)(v.x >> 32) == ep && (unsigned)(v.y >> 32) == ep;
}
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
// gather n floats (n even) whose words must carry epoch `ep` into shared memory; all loads of a thread are
// issued before the first check, so the common case costs one L2 round trip
__device__ __noinline__ void gather_flagged(float *dst_smem, const u64 *src, int n, unsigned ep, Watchdog wd) {
  const int npairs = n >> 1;
#pragma unroll 1
  for (int base = 0; base < npairs; base += GB * NCT) {
    ulonglong2 v[GB];
#pragma unroll
    for (int j = 0; j < GB; ++j) {
      const int idx = base + j * NCT + (int)threadIdx.x;
      if (idx < npairs) v[j] = ld_pair(src + 2 * idx);
    }
#pragma unroll
    for (int j = 0; j < GB; ++j) {
      const int idx = base + j * NCT + (int)threadIdx.x;
      if (idx < npairs) {
        if (!pair_ok(v[j], ep)) v[j] = spin_pair(src + 2 * idx, ep, wd);
        reinterpret_cast<float2 *>(dst_smem)[idx] = make_float2(lo_f(v[j].x), lo_f(v[j].y));
      }
    }
  }
}

// rows [r0, r1) of an N-row matrix owned by this CTA in a phase; `rot` rotates which CTAs get the remainder rows
__device__ __forceinline__ void row_range(int cta, int G, int rot, int N, int &r0, int &r1) {
  int c = cta + rot;
  if (c >= G) c -= G;
  r0 = (int)(((unsigned)c * (unsigned)N) / (unsigned)G);  // c < G <= ~150 SMs, N <= 4E or V: fits 32 bits
  r1 = (int)(((unsigned)(c + 1) * (unsigned)N) / (unsigned)G);
}
__device__ __forceinline__ int phase_rot(int layer, int ph, int G) { return ((layer * 5 + ph) * 29) % G; }

// GEMV phases.  ph numbering inside a layer: 0 = c_attn, 1 = attention (no weights), 2 = attn c_proj,
// 3 = c_fc, 4 = mlp c_proj; the tied lm_head is phase index 5L of a step.
enum { M_QKV = 0, M_RESID = 1, M_GELU = 2, M_LMHEAD = 3 };
struct PhaseDesc {
  const float *W, *bias, *ln_g, *ln_b;
  const u64 *src;
  int N, K, mode, rot;
};
__device__ __forceinline__ PhaseDesc phase_desc(const DecodeParams &p, int l, int ph, bool is_head, int G) {
  PhaseDesc d;
  const int E = p.E;
  if (is_head) {
    d = PhaseDesc{p.wte, nullptr, p.lnf_g, p.lnf_b, p.xres_f, p.V, E, M_LMHEAD, 0};
    return d;
  }
  const LayerDesc &ld = c_layers[l];
  d.rot = phase_rot(l, ph, G);
  d.K = E;
  d.ln_g = d.ln_b = nullptr;
  if (ph == 0) {
    d.W = ld.w_attn; d.bias = ld.b_attn; d.ln_g = ld.ln1_g; d.ln_b = ld.ln1_b; d.src = p.xres_f; d.N = 3 * E; d.mode = M_QKV;
  } else if (ph == 2) {
    d.W = ld.w_proj; d.bias = ld.b_proj; d.src = p.att_f; d.N = E; d.mode = M_RESID;
  } else if (ph == 3) {
    d.W = ld.w_fc; d.bias = ld.b_fc; d.ln_g = ld.ln2_g; d.ln_b = ld.ln2_b; d.src = p.xres_f; d.N = 4 * E; d.mode = M_GELU;
  } else {
    d.W = ld.w_proj2; d.bias = ld.b_proj2; d.src = p.f_f; d.N = E; d.K = 4 * E; d.mode = M_RESID;
  }
  return d;
}

struct Smem {
  float *ring;   // nslot * slotf
  float *vec;    // 2 x 4E: activation vector the GEMV phases read, double-buffered: with no CTA-wide sync at the
                 // end of a phase, a fast warp may already be gathering the next vector while a slow one still reads
                 // the current one (a gather into buffer b is two CTA syncs after the last read of buffer b)
  float *xv;     // E: staging for LayerNorm input
  float *part;   // NCW * hd attention partial outputs
  float *red;    // 64
  uint32_t full0, empty0;  // shared addresses of mbarrier arrays
  Watchdog wd;             // sticky global error word + CTA-local tripped flag
};

// LayerNorm of the E-vector in smem `src` into smem `dst`; reference formula ops.zig:86-101 (single pass E[x],
// E[x^2]; std = sqrt(var + eps)).  The affine parameters are fetched first so that their latency overlaps the
// two block reductions.  One copy of this code serves every call site (instruction-cache footprint matters:
// each phase executes its code exactly once).
__device__ __noinline__ void layer_norm_to_smem(const float *src_smem, float *dst, const float *__restrict__ g,
                                                const float *__restrict__ b, int E, float eps, float *red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float xr[LNR], lg[LNR], lb[LNR];
  float s = 0.0f, ss = 0.0f;
#pragma unroll
  for (int j = 0; j < LNR; ++j) {
    const int i = tid + j * NCT;
    lg[j] = (i < E) ? __ldg(g + i) : 0.0f;
    lb[j] = (i < E) ? __ldg(b + i) : 0.0f;
  }
#pragma unroll
  for (int j = 0; j < LNR; ++j) {
    const int i = tid + j * NCT;
    xr[j] = (i < E) ? src_smem[i] : 0.0f;
    s += xr[j];
    ss = fmaf(xr[j], xr[j], ss);
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  if (lane == 0) {
    red[warp] = s;
    red[NCW + warp] = ss;
  }
  consumer_sync();
  float ts = 0.0f, tss = 0.0f;
#pragma unroll
  for (int w = 0; w < NCW; ++w) {
    ts += red[w];
    tss += red[NCW + w];
  }
  const float n = (float)E;
  const float mean = ts / n;
  const float rstd = 1.0f / sqrtf(tss / n - mean * mean + eps);
#pragma unroll
  for (int j = 0; j < LNR; ++j) {
    const int i = tid + j * NCT;
    if (i < E) dst[i] = (xr[j] - mean) * rstd * lg[j] + lb[j];
  }
  consumer_sync();
}

// Attention work item (head h, split s of S) over cache rows [t0,t1) -- ops.zig:249-307 without the
// whole-cache transposes: K/V of earlier tokens are read in place from the time-major cache (head stride hd = 64,
// time stride E); q and the current token's K/V row arrive through the flagged exchange (epoch `ep_in`).
// One pass with an online softmax per warp, merged across warps in shared memory.
__device__ __noinline__ void attention_item(const DecodeParams &p, const Smem &sm, int l, int h, int s, int S, int T,
                                            unsigned ep_in, unsigned ep_out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int hd = 64;
  const int E = p.E, pos = T - 1;
  const int chunk = (T + S - 1) / S;
  const int t0 = s * chunk, t1 = min(T, t0 + chunk);
  const float scale = 0.125f;  // 1 / sqrt(64)
  const float *kh = c_layers[l].k_cache + h * hd;
  const float *vh = c_layers[l].v_cache + h * hd;
  float *po = sm.part;  // [NCW][hd] per-warp partial outputs

  // cache rows do not depend on this step: put the first batch in flight before waiting for q
  const bool has_new = (pos >= t0 && pos < t1);
  float2 kv[4], vv[4];
  const int tfirst = t0 + warp;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int tt = tfirst + u * NCW;
    kv[u] = make_float2(0.0f, 0.0f);
    vv[u] = make_float2(0.0f, 0.0f);
    if (tt < t1 && tt != pos) {
      kv[u] = __ldcg(reinterpret_cast<const float2 *>(kh + (size_t)tt * E) + lane);
      vv[u] = __ldcg(reinterpret_cast<const float2 *>(vh + (size_t)tt * E) + lane);
    }
  }
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
  const float2 qv = make_float2(lo_f(qw.x), lo_f(qw.y));
  const float2 knew = make_float2(lo_f(kw.x), lo_f(kw.y)), vnew = make_float2(lo_f(vw.x), lo_f(vw.y));

  float mw = -INFINITY, lw = 0.0f;
  float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll 1
  for (int t = tfirst; t < t1; t += 4 * NCW) {
    if (t != tfirst) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int tt = t + u * NCW;
        if (tt < t1 && tt != pos) {
          kv[u] = __ldcg(reinterpret_cast<const float2 *>(kh + (size_t)tt * E) + lane);
          vv[u] = __ldcg(reinterpret_cast<const float2 *>(vh + (size_t)tt * E) + lane);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (t + u * NCW == pos) {
        kv[u] = knew;
        vv[u] = vnew;
      }
    }
    float a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = fmaf(qv.x, kv[u].x, qv.y * kv[u].y);
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = warp_sum(a[u]) * scale;
    float mnew = mw;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (t + u * NCW < t1) mnew = fmaxf(mnew, a[u]);
    const float corr = (mw == -INFINITY) ? 0.0f : expf(mw - mnew);
    lw *= corr;
    acc.x *= corr;
    acc.y *= corr;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (t + u * NCW < t1) {
        const float pt = expf(a[u] - mnew);
        lw += pt;
        acc.x = fmaf(pt, vv[u].x, acc.x);
        acc.y = fmaf(pt, vv[u].y, acc.y);
      }
    }
    mw = mnew;
  }
  po[warp * hd + 2 * lane] = acc.x;
  po[warp * hd + 2 * lane + 1] = acc.y;
  if (lane == 0) {
    sm.red[16 + warp] = mw;
    sm.red[24 + warp] = lw;
  }
  consumer_sync();
  if (tid < hd) {
    float m = -INFINITY;
#pragma unroll
    for (int w = 0; w < NCW; ++w) m = fmaxf(m, sm.red[16 + w]);
    float lsum = 0.0f, o = 0.0f;
#pragma unroll
    for (int w = 0; w < NCW; ++w) {
      const float mwv = sm.red[16 + w];
      const float sc = (mwv == -INFINITY) ? 0.0f : expf(mwv - m);
      lsum = fmaf(sm.red[24 + w], sc, lsum);
      o = fmaf(po[w * hd + tid], sc, o);
    }
    if (S == 1) {
      st_flag(p.att_f + h * hd + tid, o / lsum, ep_out);
    } else {  // flash-decoding partial: (m, l, unnormalised o)
      float *mine = p.att_part + ((size_t)h * S + s) * (hd + 2);
      mine[2 + tid] = o;
      if (tid == 0) {
        mine[0] = m;
        mine[1] = lsum;
      }
    }
  }
  if (S == 1) return;
  // the last split of this head to arrive combines the partials (the only fence left: long contexts only)
  __threadfence();
  consumer_sync();
  if (tid == 0) {
    const unsigned old = atomicAdd(p.head_count + h, 1u);
    const bool last = (old == (unsigned)(S - 1));
    if (last) p.head_count[h] = 0u;
    sm.red[32] = last ? 1.0f : 0.0f;
    __threadfence();
  }
  consumer_sync();
  if (sm.red[32] != 0.0f && tid < hd) {
    const float *base = p.att_part + (size_t)h * S * (hd + 2);
    float M = -INFINITY;
    for (int i = 0; i < S; ++i) M = fmaxf(M, __ldcg(base + (size_t)i * (hd + 2)));
    float Lsum = 0.0f, a = 0.0f;
    for (int i = 0; i < S; ++i) {
      const float sc = expf(__ldcg(base + (size_t)i * (hd + 2)) - M);
      Lsum = fmaf(__ldcg(base + (size_t)i * (hd + 2) + 1), sc, Lsum);
      a = fmaf(__ldcg(base + (size_t)i * (hd + 2) + 2 + tid), sc, a);
    }
    st_flag(p.att_f + h * hd + tid, a / Lsum, ep_out);
  }
}

__device__ __forceinline__ bool step_needs_logits(const DecodeParams &p, int step) {
  return p.force_logits || step >= p.n_prompt;
}

// ---- phase table ------------------------------------------------------------------------------------
// Everything about a GEMV phase that does not depend on the token is worked out once per launch and kept in
// shared memory, so that the top of a phase is a handful of LDS instead of integer divisions and branches
// (each phase executes its code exactly once: every instruction on its critical path is latency).
struct PhaseEnt {
  const float *W, *bias, *ln_g, *ln_b;
  const u64 *src;
  int r0, nrows;  // rows of W this CTA owns in this phase
  int K, mode;
  int n_units, rps;
};

// up to 8 ring units at once: issue every try_wait before looking at any result, so their latencies overlap
__device__ __forceinline__ bool mbar_try8(const uint32_t (&bar)[8], uint32_t parity_bits) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p0, p1, p2, p3, p4, p5, p6, p7;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p0, [%1], %9;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p1, [%2], %10;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p2, [%3], %11;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p3, [%4], %12;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p4, [%5], %13;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p5, [%6], %14;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p6, [%7], %15;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p7, [%8], %16;\n\t"
      "and.pred p0, p0, p1;\n\tand.pred p2, p2, p3;\n\tand.pred p4, p4, p5;\n\tand.pred p6, p6, p7;\n\t"
      "and.pred p0, p0, p2;\n\tand.pred p4, p4, p6;\n\tand.pred p0, p0, p4;\n\t"
      "selp.u32 %0, 1, 0, p0;\n\t}"
      : "=r"(ok)
      : "r"(bar[0]), "r"(bar[1]), "r"(bar[2]), "r"(bar[3]), "r"(bar[4]), "r"(bar[5]), "r"(bar[6]), "r"(bar[7]),
        "r"(parity_bits & 1u), "r"((parity_bits >> 1) & 1u), "r"((parity_bits >> 2) & 1u), "r"((parity_bits >> 3) & 1u),
        "r"((parity_bits >> 4) & 1u), "r"((parity_bits >> 5) & 1u), "r"((parity_bits >> 6) & 1u), "r"((parity_bits >> 7) & 1u)
      : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(NTHREADS, 1) decode_persistent_kernel(const DecodeParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) u64 mbar_store[2 * MAXSLOTS];
  __shared__ unsigned wd_flag;
  const int G = gridDim.x, cta = blockIdx.x;
  const int E = p.E, E4 = 4 * p.E;
  const int nslot = p.nslot, slotf = p.slotf;
  const int L5 = 5 * p.L;
  Smem sm;
  sm.ring = reinterpret_cast<float *>(smem_raw);
  sm.vec = sm.ring + (size_t)nslot * slotf;
  sm.xv = sm.vec + 2 * E4;
  sm.part = sm.xv + E;
  sm.red = sm.part + NCW * p.hd;
  PhaseEnt *table = reinterpret_cast<PhaseEnt *>(sm.red + 64);
  sm.full0 = smem_u32(mbar_store);
  sm.empty0 = smem_u32(mbar_store + MAXSLOTS);
  sm.wd.err_global = p.err;
  sm.wd.tripped_smem = smem_u32(&wd_flag);

  if (threadIdx.x == 0) {
    wd_flag = 0u;
    for (int i = 0; i < nslot; ++i) {
      mbar_init(sm.full0 + 8u * i, 1);
      mbar_init(sm.empty0 + 8u * i, NCW);  // every consumer warp reads its slice of a unit, then releases it
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int g = threadIdx.x; g <= L5; g += blockDim.x) {
    const int l = g / 5, ph = g - 5 * l;
    const bool is_head = (g == L5);
    PhaseEnt e;
    e.W = nullptr; e.bias = nullptr; e.ln_g = nullptr; e.ln_b = nullptr; e.src = nullptr;
    e.r0 = 0; e.nrows = 0; e.K = E; e.mode = -1; e.n_units = 0; e.rps = 4;
    if (is_head || ph != 1) {
      const PhaseDesc d = phase_desc(p, l, ph, is_head, G);
      int r0, r1;
      row_range(cta, G, d.rot, d.N, r0, r1);
      e.W = d.W; e.bias = d.bias; e.ln_g = d.ln_g; e.ln_b = d.ln_b; e.src = d.src;
      e.r0 = r0; e.nrows = r1 - r0; e.K = d.K; e.mode = d.mode;
      e.rps = (d.K == E) ? 4 : 1;
      e.n_units = (e.nrows + e.rps - 1) / e.rps;
    }
    table[g] = e;
  }
  __syncthreads();

  const int last_step = p.first_step + p.n_steps - 1;
  const int bsz = nslot < 8 ? nslot : 8;  // ring units handled per GEMV batch (all on distinct slots)

  if (threadIdx.x >= NCT) {
    // =============================== producer warp ===============================
    // Streams, in consumption order, the rows this CTA owns in every GEMV phase.  Unit = up to slotf floats
    // (4 rows of K = E, or 1 row of K = 4E) = one cp.async.bulk into one ring slot.
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
        if (e.mode == M_QKV && lane < 8) {
          // pull the layer's small vectors (LayerNorm affine + biases, 13E floats) into L2 ahead of the consumers
          const LayerDesc &ld = c_layers[g / 5];
          const float *arr = lane == 0 ? ld.ln1_g : lane == 1 ? ld.ln1_b : lane == 2 ? ld.b_attn : lane == 3 ? ld.b_proj
                           : lane == 4 ? ld.ln2_g : lane == 5 ? ld.ln2_b : lane == 6 ? ld.b_fc : ld.b_proj2;
          const int len = lane == 2 ? 3 * E : lane == 6 ? E4 : E;
          const int nlines = (len * 4 + 127) / 128;
          for (int i = cta; i < nlines; i += G) prefetch_l2(reinterpret_cast<const char *>(arr) + (size_t)i * 128);
        }
        if (lane == 0) {
          const int rps = e.rps, K = e.K;
          const float *W = e.W + (size_t)e.r0 * K;
#pragma unroll 1
          for (int r = 0; r < e.nrows; r += rps) {
            const int nr = min(rps, e.nrows - r);
            const uint32_t bytes = (uint32_t)nr * (uint32_t)K * 4u;
            mbar_wait(sm.empty0 + 8u * slot, parity ^ 1u, sm.wd);
            const uint32_t fb = sm.full0 + 8u * slot;
            mbar_expect_tx(fb, bytes);
            bulk_g2s(smem_u32(sm.ring + (size_t)slot * slotf), W + (size_t)r * K, bytes, fb, pol);
            if (++slot == nslot) { slot = 0; parity ^= 1u; }
          }
        }
        __syncwarp();
      }
    }
    return;
  }

  // ================================= consumer warps =================================
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Prof pf{(p.prof && cta == 0 && tid == 0) ? p.prof : nullptr, 0, (p.dbg & 8) != 0};
  pf.mark(0);
  u64 prev_token = 0;
  int bslot = 0;               // ring slot / parity of the first unit of the current GEMV batch
  uint32_t bpar = 0;
  unsigned ep = p.epoch_base;  // epoch of the phase being executed; its inputs carry ep - 1
  int vsel = 0;                // which half of sm.vec the current GEMV phase reads
  // each warp reduces one eighth of every ring unit: floats [warp * slice, (warp + 1) * slice) of the slot
  const int slice4 = slotf >> 5;           // float4 per warp slice (slotf / 8 / 4)
  const bool small_slice = slice4 <= 96;   // E <= 768: the warp's x-slice fits 3 float4 per lane (registers)

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
    const float *te = p.wte + (size_t)tok * E, *pe = p.wpe + (size_t)pos * E;  // main.zig:179-180

#pragma unroll 1
    for (int g = 0; g < nph; ++g) {
      ++ep;
      const PhaseEnt ent = table[g];
      const bool is_head = (g == L5);
      const int tag = is_head ? 96 : 16 * (g % 5 + 1);

      if (ent.mode < 0) {
        // ---------------- attention over the cache (ops.zig:160-171) ----------------
        int S = (T + ATT_CHUNK - 1) / ATT_CHUNK;
        const int smax = G / p.H;
        if (S > smax) S = smax;
        if (cta < p.H * S) attention_item(p, sm, g / 5, cta / S, cta % S, S, T, ep - 1, ep);
        pf.mark(tag + 3);
        continue;
      }

      vsel ^= 1;
      float *vec = sm.vec + vsel * E4;
      const int rps = ent.rps, K = ent.K, mode = ent.mode;
      // ---------------- phase top: issue every load whose address is known before the activation arrives ----
      // thread t finishes row r0 + t of the first batch: bias and residual operand.  The residual is the
      // embedding itself in the first block (main.zig:181-183), otherwise the stream word written two (P5) or
      // three (P3) phases ago.
      const bool resid_from_emb = (g == 2);
      const unsigned ep_resid = ep - (K == E ? 3u : 2u);
      float bias_v = 0.0f, resid_v = 0.0f;
      u64 resid_w = 0;
      const bool own_row = tid < min(ent.nrows, bsz * rps);
      if (own_row) {
        const int r = ent.r0 + tid;
        if (ent.bias) bias_v = __ldg(ent.bias + r);
        if (mode == M_RESID) {
          if (resid_from_emb) resid_v = __ldg(te + r) + __ldg(pe + r);
          else resid_w = ld_word(p.xres_f + r);
        }
      }
      float lg[LNR], lb[LNR];
      if (ent.ln_g != nullptr) {
#pragma unroll
        for (int j = 0; j < LNR; ++j) {
          lg[j] = lb[j] = 0.0f;
          if (j * NCT < E) {  // uniform: skips the iterations a narrow model does not need
            const int i = tid + j * NCT;
            if (i < E) {
              lg[j] = __ldg(ent.ln_g + i);
              lb[j] = __ldg(ent.ln_b + i);
            }
          }
        }
      }
      pf.fmark(256 + 4);

      // ---------------- activation vector -> shared memory ----------------
      if (ent.ln_g != nullptr) {
        if (g == 0) {  // wte[token] + wpe[pos] (main.zig:179-183), recomputed by every CTA
#pragma unroll 1
          for (int i = tid; i < E; i += NCT) sm.xv[i] = __ldg(te + i) + __ldg(pe + i);
        } else {
          gather_flagged(sm.xv, ent.src, E, ep - 1, sm.wd);
        }
        pf.fmark(256 + 5);
        consumer_sync();
        pf.fmark(256 + 6);
        // LayerNorm, reference formula ops.zig:86-101 (single pass E[x], E[x^2]; std = sqrt(var + eps))
        float xr[LNR];
        float s = 0.0f, ss = 0.0f;
#pragma unroll
        for (int j = 0; j < LNR; ++j) {
          xr[j] = 0.0f;
          if (j * NCT < E) {
            const int i = tid + j * NCT;
            if (i < E) xr[j] = sm.xv[i];
            s += xr[j];
            ss = fmaf(xr[j], xr[j], ss);
          }
        }
        s = warp_sum(s);
        ss = warp_sum(ss);
        if (lane == 0) {
          sm.red[warp] = s;
          sm.red[NCW + warp] = ss;
        }
        consumer_sync();
        float ts = 0.0f, tss = 0.0f;
#pragma unroll
        for (int w = 0; w < NCW; ++w) {
          ts += sm.red[w];
          tss += sm.red[NCW + w];
        }
        const float nE = (float)E;
        const float mean = ts / nE;
        const float rstd = 1.0f / sqrtf(tss / nE - mean * mean + 1e-5f);
#pragma unroll
        for (int j = 0; j < LNR; ++j) {
          if (j * NCT < E) {
            const int i = tid + j * NCT;
            if (i < E) vec[i] = (xr[j] - mean) * rstd * lg[j] + lb[j];
          }
        }
        consumer_sync();
        if (is_head && p.write_xout && step == last_step && cta == 0) {
#pragma unroll 1
          for (int i = tid; i < E; i += NCT) {
            p.xout[i] = vec[i];
            p.xres_out[i] = sm.xv[i];
          }
        }
      } else {
        gather_flagged(vec, ent.src, K, ep - 1, sm.wd);
        pf.fmark(256 + 5);
        consumer_sync();
      }
      pf.mark(tag + 1);

      // ---------------- GEMV: every warp reduces its eighth of every ring unit ----------------
      // A unit is slotf floats: 4 rows of K = E (warp w covers half of row w / 2) or one row of K = 4E (warp w
      // covers an eighth of it).  Either way the warp's activation slice is the same for every unit of the
      // phase, so it is read once.  Per-unit partial sums stay in registers until the whole batch is consumed;
      // then 8 interleaved shuffle reductions, one CTA sync, and the first threads finish one row each.
      float *kc = nullptr, *vc = nullptr;
      if (mode == M_QKV) {
        kc = c_layers[g / 5].k_cache + (size_t)pos * E;
        vc = c_layers[g / 5].v_cache + (size_t)pos * E;
      }
      float *logits = (is_head && p.store_logits && step == last_step) ? p.logits : nullptr;
      const int xoff4 = (rps == 4) ? (warp & 1) * slice4 : warp * slice4;  // the warp's slice of the activation vector
      const int row_in_unit = (rps == 4) ? (warp >> 1) : 0;
      const float4 *vec4 = reinterpret_cast<const float4 *>(vec) + xoff4;
      float4 xs0 = make_float4(0.f, 0.f, 0.f, 0.f), xs1 = xs0, xs2 = xs0;
      if (small_slice) {
        if (lane < slice4) xs0 = vec4[lane];
        if (lane + 32 < slice4) xs1 = vec4[lane + 32];
        if (lane + 64 < slice4) xs2 = vec4[lane + 64];
      }
      float best = -INFINITY;  // running argmax of the rows this thread finishes (lm_head only)
      unsigned best_i = 0xffffffffu;
      const int spr = NCW / rps;  // warp slices per row

#pragma unroll 1
      for (int ub = 0; ub < ent.n_units; ub += bsz) {
        const int nb = min(bsz, ent.n_units - ub);
        uint32_t fbar[8];
        int slotj[8];
        uint32_t pbits = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          int sj = bslot + (j < nb ? j : 0);
          uint32_t pj = bpar;
          if (sj >= nslot) { sj -= nslot; pj ^= 1u; }
          slotj[j] = sj;
          fbar[j] = sm.full0 + 8u * sj;
          pbits |= pj << j;
        }
        if (!mbar_try8(fbar, pbits)) {
#pragma unroll 1
          for (int j = 0; j < nb; ++j) mbar_wait(fbar[j], (pbits >> j) & 1u, sm.wd);
        }
        pf.fmark(256 + 7);
        float acc[8];
        if (small_slice) {
          // branch-free: every load is unconditional (slots past nb alias slot 0, lanes past the slice are clamped and
          // meet a zero activation), so the compiler can put all 24 LDS.128 in flight before the first FMA
          const int i0 = min(lane, slice4 - 1), i1 = min(lane + 32, slice4 - 1), i2 = min(lane + 64, slice4 - 1);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 *w4 = reinterpret_cast<const float4 *>(sm.ring + (size_t)slotj[j] * slotf) + warp * slice4;
            const float4 wa = w4[i0], wb = w4[i1], wc = w4[i2];
            float a0 = wa.x * xs0.x, a1 = wa.y * xs0.y;
            a0 = fmaf(wa.z, xs0.z, a0); a1 = fmaf(wa.w, xs0.w, a1);
            a0 = fmaf(wb.x, xs1.x, a0); a1 = fmaf(wb.y, xs1.y, a1); a0 = fmaf(wb.z, xs1.z, a0); a1 = fmaf(wb.w, xs1.w, a1);
            a0 = fmaf(wc.x, xs2.x, a0); a1 = fmaf(wc.y, xs2.y, a1); a0 = fmaf(wc.z, xs2.z, a0); a1 = fmaf(wc.w, xs2.w, a1);
            const bool valid = (j < nb) && (row_in_unit < min(rps, ent.nrows - (ub + j) * rps));
            acc[j] = valid ? a0 + a1 : 0.0f;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            acc[j] = 0.0f;
            if (j < nb && row_in_unit < min(rps, ent.nrows - (ub + j) * rps)) {
              const float4 *w4 = reinterpret_cast<const float4 *>(sm.ring + (size_t)slotj[j] * slotf) + warp * slice4;
              float a0 = 0.0f, a1 = 0.0f;
#pragma unroll 1
              for (int i = lane; i < slice4; i += 32) {
                const float4 w = w4[i];
                const float4 x = vec4[i];
                a0 = fmaf(w.x, x.x, a0); a1 = fmaf(w.y, x.y, a1); a0 = fmaf(w.z, x.z, a0); a1 = fmaf(w.w, x.w, a1);
              }
              acc[j] = a0 + a1;
            }
          }
        }
        __syncwarp();
        if (lane < nb) mbar_arrive(sm.empty0 + 8u * (uint32_t)((bslot + lane >= nslot) ? bslot + lane - nslot : bslot + lane));
        pf.fmark(256 + 8);
        // packed reduction of the 8 per-unit sums: exchange halves (xor 16, 8, 4), then two plain steps;
        // afterwards lane L (L % 4 == 0) holds the warp total of unit ((L >> 4) & 1) * 4 + ((L >> 3) & 1) * 2 + ((L >> 2) & 1)
        {
          const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
          float b0 = h16 ? acc[4] : acc[0], b1 = h16 ? acc[5] : acc[1], b2 = h16 ? acc[6] : acc[2], b3 = h16 ? acc[7] : acc[3];
          const float s0 = h16 ? acc[0] : acc[4], s1 = h16 ? acc[1] : acc[5], s2 = h16 ? acc[2] : acc[6], s3 = h16 ? acc[3] : acc[7];
          b0 += __shfl_xor_sync(0xffffffffu, s0, 16);
          b1 += __shfl_xor_sync(0xffffffffu, s1, 16);
          b2 += __shfl_xor_sync(0xffffffffu, s2, 16);
          b3 += __shfl_xor_sync(0xffffffffu, s3, 16);
          float c0 = h8 ? b2 : b0, c1 = h8 ? b3 : b1;
          const float t0 = h8 ? b0 : b2, t1 = h8 ? b1 : b3;
          c0 += __shfl_xor_sync(0xffffffffu, t0, 8);
          c1 += __shfl_xor_sync(0xffffffffu, t1, 8);
          float d0 = h4 ? c1 : c0;
          const float u0 = h4 ? c0 : c1;
          d0 += __shfl_xor_sync(0xffffffffu, u0, 4);
          d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
          d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
          if ((lane & 3) == 0) sm.part[(lane >> 2) * NCW + warp] = d0;
        }
        bslot += nb;
        if (bslot >= nslot) { bslot -= nslot; bpar ^= 1u; }
        consumer_sync();
        pf.fmark(256 + 9);
        // ---------------- epilogue: thread t finishes row (ub * rps + t) of this phase ----------------
        const int rr = ub * rps + tid;
        if (tid < nb * rps && rr < ent.nrows) {
          const int j = (rps == 4) ? (tid >> 2) : tid, ri = tid - j * rps;
          const int r = ent.r0 + rr;
          float v = 0.0f;
          for (int sgm = 0; sgm < spr; ++sgm) v += sm.part[j * NCW + ri * spr + sgm];
          if (ub > 0) {  // operands of later batches were not prefetched at the top of the phase
            bias_v = ent.bias ? __ldg(ent.bias + r) : 0.0f;
            if (mode == M_RESID) {
              if (resid_from_emb) resid_v = __ldg(te + r) + __ldg(pe + r);
              else resid_w = ld_word(p.xres_f + r);
            }
          }
          v += bias_v;
          if (mode == M_QKV) {  // q to the exchange, k/v to cache row `pos` (ops.zig:146-158) and to the exchange
            if (r < E) {
              st_flag(p.q_f + r, v, ep);
            } else if (r < 2 * E) {
              kc[r - E] = v;
              st_flag(p.kvn_f + (r - E), v, ep);
            } else {
              vc[r - 2 * E] = v;
              st_flag(p.kvn_f + E + (r - 2 * E), v, ep);
            }
          } else if (mode == M_RESID) {  // main.zig:136-139,142-145
            if (!resid_from_emb) {
              if ((unsigned)(resid_w >> 32) != ep_resid) resid_w = spin_word(p.xres_f + r, ep_resid, sm.wd);
              resid_v = lo_f(resid_w);
            }
            st_flag(p.xres_f + r, v + resid_v, ep);
          } else if (mode == M_GELU) {  // main.zig:80
            st_flag(p.f_f + r, gelu_ref(v), ep);
          } else {  // tied lm_head (main.zig:193) + running argmax; this thread sees increasing r, so strict >
            if (logits) logits[r] = v;
            if (v > best) { best = v; best_i = (unsigned)r; }
          }
        }
        if (ub + bsz < ent.n_units) consumer_sync();  // sm.part is rewritten by the next batch
      }
      pf.mark(tag + 3);

      if (is_head) {
        // CTA-level argmax (value desc, index asc): the finishing threads all live in warp 0.  One flagged
        // partial per CTA, then every CTA reduces the G partials itself: the next step's embedding needs the
        // token everywhere.
        if (tid < 32) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const unsigned oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
          }
          if (tid == 0) {
            st_flag(p.amax_f + 2 * cta, best, ep);
            st_flag(p.amax_f + 2 * cta + 1, __uint_as_float(best_i), ep);
          }
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
          if (tid == 0) sm.red[40] = __uint_as_float(bi);
        }
        consumer_sync();
        const u64 amax = (u64)__float_as_uint(sm.red[40]);
        if (cta == 0 && tid == 0) *p.last_token = amax;
        if (step >= p.n_prompt) out_tok = amax;  // generate(): main.zig:335-338
        consumer_sync();
        pf.mark(tag + 4);
      }
    }

    if (!want_logits && p.write_xout && step == last_step && cta == 0) {
      // GPT.forward(compute_logits = false) still leaves ln_f(x) in state.x (main.zig:189)
      gather_flagged(sm.xv, p.xres_f, E, ep, sm.wd);
      consumer_sync();  // also: every warp is past the last GEMV's reads of sm.vec
      layer_norm_to_smem(sm.xv, sm.vec, p.lnf_g, p.lnf_b, E, 1e-5f, sm.red);
#pragma unroll 1
      for (int i = tid; i < E; i += NCT) {
        p.xout[i] = sm.vec[i];
        p.xres_out[i] = sm.xv[i];
      }
    }
    if (cta == 0 && tid == 0) {
      p.tokens[step] = out_tok;
      if (p.tokens_host) p.tokens_host[step] = out_tok;
    }
    prev_token = out_tok;
  }
  pf.mark(1);
  if (pf.buf) pf.buf[2 * PROF_MAX] = (u64)pf.i;
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
  u64 *prompt_dev;
  u64 *tokens_dev;
  u64 *tokens_host;  // pinned, mapped
  u64 *tokens_host_devptr;
  u64 *last_token_dev;
  u64 *prof_dev;
  unsigned *err_dev;
  unsigned epoch_count;  // host mirror of the phase epoch (monotonic across launches)
  int grid;
  size_t smem_bytes;
  int n_prompt;
  int prof_enabled;
};

static zg_engine *g_table_owner = nullptr;  // whose layer table currently sits in __constant__ memory

static size_t engine_smem_bytes(const zg_config &c, int nslot) {
  const size_t E = c.n_embed, hd = E / c.n_heads;
  const size_t floats = (size_t)nslot * 4 * E + 2 * 4 * E + E + NCW * hd + 64;
  return floats * sizeof(float) + (5 * c.n_layer + 1) * sizeof(PhaseEnt);
}

extern "C" {

zg_engine *zg_engine_create(const zg_gpt *gpt, const zg_state *state) {
  if (!require_ready("zg_engine_create")) return nullptr;
  Context &c = ctx();
  const zg_config &cfg = gpt->config;
  const size_t E = cfg.n_embed;
  if (E % 8 != 0 || cfg.n_heads * 64 != E || E > (size_t)LNR * NCT || cfg.n_layer > (size_t)MAX_LAYERS) {
    set_error(1, "zg_engine_create: needs head_dim 64, n_embed % 8 == 0, n_embed <= 1792, n_layer <= 64", __FILE__, __LINE__);
    return nullptr;
  }
  zg_engine *e = (zg_engine *)calloc(1, sizeof(zg_engine));
  if (!e) return nullptr;
  e->cfg = cfg;
  e->state = *state;
  e->grid = c.sm_count;
  if ((size_t)e->grid < cfg.n_heads) {
    set_error(1, "zg_engine_create: fewer SMs than attention heads", __FILE__, __LINE__);
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
  ZG_CUDA(cudaFuncSetAttribute(decode_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)e->smem_bytes));
  int per_sm = 0;
  ZG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_persistent_kernel, NTHREADS, e->smem_bytes));
  if (per_sm < 1) {
    set_error(1, "zg_engine_create: persistent kernel does not fit on an SM", __FILE__, __LINE__);
    free(e);
    return nullptr;
  }

  // layer table (start-up only); copied into __constant__ memory before a launch when another engine owned it
  e->layers_host = (LayerDesc *)calloc(MAX_LAYERS, sizeof(LayerDesc));
  for (size_t l = 0; l < cfg.n_layer; ++l) {
    const zg_block &b = gpt->h[l];
    e->layers_host[l] = LayerDesc{b.ln_1.weight, b.ln_1.bias, b.attn.c_attn.weight, b.attn.c_attn.bias,
                                  b.attn.c_proj.weight, b.attn.c_proj.bias, b.ln_2.weight, b.ln_2.bias,
                                  b.mlp.c_fc.weight, b.mlp.c_fc.bias, b.mlp.c_proj.weight, b.mlp.c_proj.bias,
                                  b.k_cache, b.v_cache};
  }

  const size_t C = cfg.context_size, hd = 64;
  const int smax = e->grid / (int)cfg.n_heads > 0 ? e->grid / (int)cfg.n_heads : 1;
  const size_t n_exchange = E + E + 2 * E + E + 4 * E + 2 * (size_t)e->grid;
  e->exchange_dev = (u64 *)zg_alloc(n_exchange * 8);
  e->prompt_dev = (u64 *)zg_alloc(C * 8);
  e->tokens_dev = (u64 *)zg_alloc(C * 8);
  e->last_token_dev = (u64 *)zg_alloc(8);
  e->prof_dev = (u64 *)zg_alloc((2 * PROF_MAX + 4) * 8);
  e->err_dev = (unsigned *)zg_alloc(256);
  float *att_part = (float *)zg_alloc(cfg.n_heads * (size_t)smax * (hd + 2) * sizeof(float));
  unsigned *head_count = (unsigned *)zg_alloc(cfg.n_heads * sizeof(unsigned));
  ZG_CUDA(cudaHostAlloc(&e->tokens_host, C * 8, cudaHostAllocMapped));
  ZG_CUDA(cudaHostGetDevicePointer((void **)&e->tokens_host_devptr, e->tokens_host, 0));
  if (zg_last_error()) {
    free(e);
    return nullptr;
  }
  memset(e->tokens_host, 0xff, C * 8);
  zg_memset(e->exchange_dev, 0, n_exchange * 8);  // epoch 0 everywhere; the first phase of the first launch is epoch 1
  zg_memset(e->err_dev, 0, 256);
  zg_memset(head_count, 0, cfg.n_heads * sizeof(unsigned));
  zg_memset(e->tokens_dev, 0, C * 8);
  zg_memset(e->prof_dev, 0, (2 * PROF_MAX + 4) * 8);
  e->epoch_count = 0;

  DecodeParams &p = e->base;
  memset(&p, 0, sizeof(p));
  p.E = (int)E; p.H = (int)cfg.n_heads; p.hd = (int)hd; p.L = (int)cfg.n_layer; p.V = (int)cfg.vocab_size; p.C = (int)C;
  p.nslot = nslot; p.slotf = 4 * (int)E;
  p.wte = gpt->wte.weight; p.wpe = gpt->wpe.weight; p.lnf_g = gpt->ln_f.weight; p.lnf_b = gpt->ln_f.bias;
  u64 *x = e->exchange_dev;
  p.xres_f = x; x += E;
  p.q_f = x; x += E;
  p.kvn_f = x; x += 2 * E;
  p.att_f = x; x += E;
  p.f_f = x; x += 4 * E;
  p.amax_f = x;
  p.xres_out = state->o; p.xout = state->x; p.logits = state->logits;
  p.att_part = att_part; p.head_count = head_count; p.err = e->err_dev;
  p.tokens = e->tokens_dev; p.tokens_host = e->tokens_host_devptr; p.last_token = e->last_token_dev;
  p.dbg = getenv("ZG_DEBUG") ? atoi(getenv("ZG_DEBUG")) : 0;
  zg_sync();
  return zg_last_error() ? (free(e), nullptr) : e;
}

void zg_engine_destroy(zg_engine *e) {
  if (!e) return;
  zg_sync();
  if (g_table_owner == e) g_table_owner = nullptr;
  zg_free(e->exchange_dev); zg_free(e->prompt_dev); zg_free(e->tokens_dev); zg_free(e->last_token_dev);
  zg_free(e->prof_dev); zg_free(e->err_dev); zg_free(e->base.att_part); zg_free(e->base.head_count);
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
  if (g_table_owner != e) {  // stream-ordered, so a launch in flight keeps the table it was given
    ZG_CUDA(cudaMemcpyToSymbolAsync(c_layers, e->layers_host, sizeof(LayerDesc) * MAX_LAYERS, 0,
                                    cudaMemcpyHostToDevice, ctx().stream));
    g_table_owner = e;
  }
  p.epoch_base = e->epoch_count;
  p.prof = e->prof_enabled ? e->prof_dev : nullptr;
  e->epoch_count += phases_for(e, p);
  void *args[] = {(void *)&p};
  ZG_CUDA(cudaLaunchCooperativeKernel((const void *)decode_persistent_kernel, dim3(e->grid), dim3(NTHREADS), args,
                                      e->smem_bytes, ctx().stream));
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
  engine_launch(e, p);
}

size_t zg_engine_sample_greedy(zg_engine *e, size_t seq_len, size_t token) {
  if (!require_ready("zg_engine_sample_greedy")) return (size_t)-1;
  Context &c = ctx();
  zg_engine_forward(e, seq_len, token, 1);
  ZG_CUDA(cudaMemcpyAsync(c.token_slot_host, e->last_token_dev, 8, cudaMemcpyDeviceToHost, c.stream));
  ZG_CUDA(cudaStreamSynchronize(c.stream));
  if (engine_check_watchdog(e)) return (size_t)-1;
  return (size_t)c.token_slot_host[0];
}

size_t zg_engine_sample(zg_engine *e, size_t seq_len, float temp, size_t token, double u) {
  if (!require_ready("zg_engine_sample")) return (size_t)-1;
  Context &c = ctx();
  zg_engine_forward(e, seq_len, token, 1);
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

size_t zg_engine_read_profile(zg_engine *e, unsigned long long *out, size_t max_entries) {
  if (!require_ready("zg_engine_read_profile")) return 0;
  if (out == nullptr) {  // toggle: calling with NULL enables (max_entries != 0) or disables profiling
    e->prof_enabled = max_entries ? 1 : 0;
    return 0;
  }
  // out receives (tag, ns) pairs; returns the number of pairs
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
