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
constexpr int ATT_CHUNK = 16 * NCW;  // KV rows per attention work item before splitting (one register round)
constexpr int PROF_MAX = 16384;
constexpr int GB = 7;           // flagged pairs a thread keeps in flight while gathering (7 x 224 threads >= 4 x 768 / 2)

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
// gather n floats (n even) whose words must carry epoch `ep` into shared memory; all loads of a thread are
// issued before the first check, so the common case costs one L2 round trip
__device__ __forceinline__ void gather_flagged(float *dst_smem, const u64 *src, int n, unsigned ep, Watchdog wd) {
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
  float *lnp;    // 2E: LayerNorm gain then shift of the current phase (cp.async at the top of the phase)
  float *part;   // NCW * hd attention partial outputs
  float *red;    // 64
  uint32_t full0, empty0;  // shared addresses of mbarrier arrays
  Watchdog wd;             // sticky global error word + CTA-local tripped flag
};

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
// A warp owns rows t0 + warp + 8u; a round covers 16 rows per warp (128 per CTA), all of them in registers
// BEFORE q is waited for (cache rows do not depend on this step), so a context of up to 128 rows per split costs
// no exposed L2 latency after q lands.  Scores: per-lane partial dot over the lane's 2 dims, packed butterfly
// (16 shuffles for 16 rows), one exp per lane, p broadcast by shuffle for the PV accumulation.
constexpr int AR = 16;  // rows per warp per round
__device__ __forceinline__ void attention_item(const DecodeParams &p, const Smem &sm, int l, int h, int s, int S, int T,
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
  // row index (within a round) whose score this lane holds after the packed butterfly
  const int myu = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);

  float mw = -INFINITY, lw = 0.0f;  // lw: this lane's share of the softmax denominator (even lanes only)
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
    const float mnew = fmaxf(mw, warp_max(valid ? sc : -INFINITY));
    const float corr = (mw == -INFINITY) ? 0.0f : expf(mw - mnew);
    const float pt = valid ? expf(sc - mnew) : 0.0f;
    lw = fmaf(lw, corr, (lane & 1) ? 0.0f : pt);
    acc.x *= corr;
    acc.y *= corr;
#pragma unroll
    for (int u = 0; u < AR; ++u) {
      const int src_lane = ((u >> 3) & 1) * 16 + ((u >> 2) & 1) * 8 + ((u >> 1) & 1) * 4 + (u & 1) * 2;
      const float pu = __shfl_sync(0xffffffffu, pt, src_lane);
      acc.x = fmaf(pu, vv[u].x, acc.x);  // rows past t1 carry pu == 0 (their stale vv is finite)
      acc.y = fmaf(pu, vv[u].y, acc.y);
    }
    mw = mnew;
  }
  lw = warp_sum(lw);
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
  int rb, rps;    // rows per batch (one mbarrier wait), rows per ring unit (4 when K = E, 1 when K = 4E)
};

// LayerNorm on a warp's register copy of the E-vector (lane holds float4 number lane + 32 j); reference formula
// ops.zig:86-101: single pass E[x], E[x^2]; std = sqrt(var + eps); divide.  Every warp computes the statistics
// redundantly from its own registers: no shared-memory round trip and no CTA sync.
template <int NJ>
__device__ __forceinline__ void ln_regs(float4 (&xs)[NJ], const float4 *__restrict__ g4, const float4 *__restrict__ b4,
                                        int E, int lane) {
  float s0 = 0.0f, s1 = 0.0f, q0 = 0.0f, q1 = 0.0f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {  // lanes past the vector hold zeros
    s0 += xs[j].x + xs[j].y;
    s1 += xs[j].z + xs[j].w;
    q0 = fmaf(xs[j].x, xs[j].x, q0); q1 = fmaf(xs[j].y, xs[j].y, q1);
    q0 = fmaf(xs[j].z, xs[j].z, q0); q1 = fmaf(xs[j].w, xs[j].w, q1);
  }
  float s = s0 + s1, ss = q0 + q1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  const float nE = (float)E;
  const float mean = s / nE;
  const float rstd = 1.0f / sqrtf(ss / nE - mean * mean + 1e-5f);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int i4 = lane + 32 * j;
    if (i4 < (E >> 2)) {  // padding lanes keep their zeros
      const float4 g = g4[i4], b = b4[i4];
      xs[j].x = (xs[j].x - mean) * rstd * g.x + b.x;
      xs[j].y = (xs[j].y - mean) * rstd * g.y + b.y;
      xs[j].z = (xs[j].z - mean) * rstd * g.z + b.z;
      xs[j].w = (xs[j].w - mean) * rstd * g.w + b.w;
    }
  }
}

// NJ = float4 per lane that cover one E-vector: E <= 128 NJ.
template <int NJ>
__global__ void __launch_bounds__(NTHREADS, 1) decode_persistent_kernel(const DecodeParams p) {
#ifndef ZG_XREG_MAXNJ
#define ZG_XREG_MAXNJ 6
#endif
  constexpr bool XREG4 = (NJ <= ZG_XREG_MAXNJ);  // the 4E-vector of the mlp c_proj phase also fits in registers
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) u64 mbar_store[2 * MAXSLOTS];
  __shared__ unsigned wd_flag;
  const int G = gridDim.x, cta = blockIdx.x;
  const int E = p.E, E4 = 4 * p.E, Eq = p.E >> 2;  // Eq: float4 per E-vector
  const int nslot = p.nslot, slotf = p.slotf;
  const int L5 = 5 * p.L;
  Smem sm;
  sm.ring = reinterpret_cast<float *>(smem_raw);
  sm.vec = sm.ring + (size_t)nslot * slotf;
  sm.lnp = sm.vec + 2 * E4;
  sm.part = sm.lnp + 2 * E;
  sm.red = sm.part + NCW * p.hd;
  PhaseEnt *table = reinterpret_cast<PhaseEnt *>(sm.red + 64);
  sm.full0 = smem_u32(mbar_store);
  sm.empty0 = smem_u32(mbar_store + MAXSLOTS);
  sm.wd.err_global = p.err;
  sm.wd.tripped_smem = smem_u32(&wd_flag);
  const int bsz = min(NCW, nslot >> 1);  // ring units per batch: two batches always fit in the ring, and a
                                         // batch of K = 4E rows gives every warp at most one row

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
    const bool is_head = (g == L5);
    PhaseEnt e;
    e.W = nullptr; e.bias = nullptr; e.ln_g = nullptr; e.ln_b = nullptr; e.src = nullptr;
    e.r0 = 0; e.nrows = 0; e.K = E; e.mode = -1; e.rb = 4; e.rps = 4;
    if (is_head || ph != 1) {
      const PhaseDesc d = phase_desc(p, l, ph, is_head, G);
      int r0, r1;
      row_range(cta, G, d.rot, d.N, r0, r1);
      e.W = d.W; e.bias = d.bias; e.ln_g = d.ln_g; e.ln_b = d.ln_b; e.src = d.src;
      e.r0 = r0; e.nrows = r1 - r0; e.K = d.K; e.mode = d.mode;
      e.rps = (d.K == E) ? 4 : 1;
      e.rb = bsz * e.rps;
    }
    table[g] = e;
  }
  __syncthreads();

  const int last_step = p.first_step + p.n_steps - 1;

  if (threadIdx.x >= NCT) {
    // =============================== producer warp ===============================
    // Streams, in consumption order, the rows this CTA owns in every GEMV phase.  Unit = up to slotf floats
    // (4 rows of K = E, or 1 row of K = 4E) = one cp.async.bulk into one ring slot; the units of a batch all
    // complete on the full-barrier of the batch's first slot, so a consumer waits once per batch.
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
          const int rps = e.rps, K = e.K, rb = e.rb;
          const float *W = e.W + (size_t)e.r0 * K;
#pragma unroll 1
          for (int b0 = 0; b0 < e.nrows; b0 += rb) {
            const int nbr = min(rb, e.nrows - b0);
            const uint32_t fb = sm.full0 + 8u * slot;
            mbar_wait(sm.empty0 + 8u * slot, parity ^ 1u, sm.wd);  // first slot free => its full barrier is idle too
            mbar_expect_tx(fb, (uint32_t)nbr * (uint32_t)K * 4u);
#pragma unroll 1
            for (int r = 0; r < nbr; r += rps) {
              const int nr = min(rps, nbr - r);
              if (r) mbar_wait(sm.empty0 + 8u * slot, parity ^ 1u, sm.wd);
              bulk_g2s(smem_u32(sm.ring + (size_t)slot * slotf), W + (size_t)(b0 + r) * K, (uint32_t)nr * (uint32_t)K * 4u,
                       fb, pol);
              if (++slot == nslot) { slot = 0; parity ^= 1u; }
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
  Prof pf{(p.prof && cta == 0 && tid == 0) ? p.prof : nullptr, 0, (p.dbg & 8) != 0};
  pf.mark(0);
  u64 prev_token = 0;
  int bslot = 0;               // ring slot of the first unit of the next batch
  uint32_t fpar = 0;           // bit s: parity the next wait on full barrier s expects
  unsigned ep = p.epoch_base;  // epoch of the phase being executed; its inputs carry ep - 1
  int vsel = 0;                // which half of sm.vec the current GEMV phase reads
  // after packed_reduce<4> lane L holds the total of the warp's row number (L >> 3); lanes 0, 8, 16, 24 finish rows
  const int eidx = lane >> 3;
  const bool elane = (lane & 7) == 0;

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
      const float4 *vec4 = reinterpret_cast<const float4 *>(vec);
      const int rps = ent.rps, K = ent.K, mode = ent.mode, rb = ent.rb;
      const bool k4 = (rps == 1);
      const bool has_ln = ent.ln_g != nullptr;
      // ---------------- phase top: issue every load whose address is known before the activation arrives ----
      // lanes 0/8/16/24 of warp w finish rows w, w + NCW, w + 2 NCW, w + 3 NCW of a batch: bias and residual operand.
      // The residual is the embedding itself in the first block (main.zig:181-183), otherwise the stream word
      // written two (P5) or three (P3) phases ago.
      const bool resid_from_emb = (g == 2);
      const unsigned ep_resid = ep - (k4 ? 2u : 3u);
      const int iloc = warp + NCW * eidx;  // row of a batch this lane finishes
      float bias_v = 0.0f, resid_v = 0.0f;
      u64 resid_w = 0;
      if (elane && iloc < min(rb, ent.nrows)) {
        const int r = ent.r0 + iloc;
        if (ent.bias) bias_v = __ldg(ent.bias + r);
        if (mode == M_RESID) {
          if (resid_from_emb) resid_v = __ldg(te + r) + __ldg(pe + r);
          else resid_w = ld_word(p.xres_f + r);
        }
      }
      if (has_ln) {  // LayerNorm affine parameters: global -> shared without passing through registers
#pragma unroll 1
        for (int i4 = tid; i4 < 2 * Eq; i4 += NCT) {
          const float *src = (i4 < Eq) ? ent.ln_g + 4 * i4 : ent.ln_b + 4 * (i4 - Eq);
          cp_async16(smem_u32(sm.lnp + 4 * i4), src);
        }
        cp_async_commit();
      }

      // ---------------- activation vector -> shared memory -> registers ----------------
      if (g == 0) {  // wte[token] + wpe[pos] (main.zig:179-183), recomputed by every CTA
#pragma unroll 1
        for (int i = tid; i < E; i += NCT) vec[i] = __ldg(te + i) + __ldg(pe + i);
      } else {
        gather_flagged(vec, ent.src, K, ep - 1, sm.wd);
      }
      if (has_ln) cp_async_wait_all();
      pf.fmark(256 + 5);
      consumer_sync();
      float4 xs[NJ];
      float4 xl[XREG4 ? 4 * NJ : 1];
      if (!k4) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int i4 = lane + 32 * j;
          xs[j] = (i4 < Eq) ? vec4[i4] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (has_ln) {
          ln_regs<NJ>(xs, reinterpret_cast<const float4 *>(sm.lnp), reinterpret_cast<const float4 *>(sm.lnp) + Eq, E, lane);
          if (is_head && p.write_xout && step == last_step && cta == 0 && warp == 0) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
              const int i4 = lane + 32 * j;
              if (i4 < Eq) {
                reinterpret_cast<float4 *>(p.xout)[i4] = xs[j];
                reinterpret_cast<float4 *>(p.xres_out)[i4] = vec4[i4];
              }
            }
          }
        }
      } else if (XREG4) {
#pragma unroll
        for (int j = 0; j < 4 * NJ; ++j) {
          const int i4 = lane + 32 * j;
          xl[XREG4 ? j : 0] = (i4 < E) ? vec4[i4] : make_float4(0.f, 0.f, 0.f, 0.f);  // E = float4 per 4E-vector
        }
      }
      pf.mark(tag + 1);

      // ---------------- GEMV: warp w takes rows w, w + 8, ... of every batch ----------------
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
        const int nun = k4 ? nbr : ((nbr + 3) >> 2);
        const bool own = elane && iloc < nbr;
        const int r = ent.r0 + b0 + iloc;
        if (b0 > 0 && own) {  // operands of later batches (the first batch's were fetched at the top of the phase)
          bias_v = ent.bias ? __ldg(ent.bias + r) : 0.0f;
          if (mode == M_RESID) {
            if (resid_from_emb) resid_v = __ldg(te + r) + __ldg(pe + r);
            else resid_w = ld_word(p.xres_f + r);
          }
        }
        mbar_wait(sm.full0 + 8u * bslot, (fpar >> bslot) & 1u, sm.wd);
        fpar ^= 1u << bslot;
        pf.fmark(256 + 7);
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (!k4) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = warp + NCW * u;
            if (i < nbr) {  // warp-uniform
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
        } else if (warp < nbr) {
          int sl = bslot + warp;
          if (sl >= nslot) sl -= nslot;
          const float4 *w4 = reinterpret_cast<const float4 *>(sm.ring + (size_t)sl * slotf);
          float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
          if (XREG4) {
#pragma unroll
            for (int j = 0; j < 4 * NJ; ++j) {
              const int i4 = min(lane + 32 * j, E - 1);
              const float4 w = w4[i4];
              const float4 x = xl[XREG4 ? j : 0];
              a0 = fmaf(w.x, x.x, a0); a1 = fmaf(w.y, x.y, a1); a2 = fmaf(w.z, x.z, a2); a3 = fmaf(w.w, x.w, a3);
            }
          } else {
#pragma unroll 4
            for (int i4 = lane; i4 < E; i4 += 32) {
              const float4 w = w4[i4];
              const float4 x = vec4[i4];
              a0 = fmaf(w.x, x.x, a0); a1 = fmaf(w.y, x.y, a1); a2 = fmaf(w.z, x.z, a2); a3 = fmaf(w.w, x.w, a3);
            }
          }
          acc[0] = (a0 + a1) + (a2 + a3);
        }
        __syncwarp();
        if (lane < nun) {
          int sl = bslot + lane;
          if (sl >= nslot) sl -= nslot;
          mbar_arrive(sm.empty0 + 8u * (uint32_t)sl);
        }
        pf.fmark(256 + 8);
        float v = packed_reduce<4>(acc, lane);  // lane L: total of the warp's row number L >> 3
        bslot += nun;
        if (bslot >= nslot) bslot -= nslot;
        // ---------------- epilogue: lanes 0/8/16/24 finish one row each ----------------
        if (own) {
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
#ifndef ZG_NOWAIT
              if ((unsigned)(resid_w >> 32) != ep_resid) resid_w = spin_word(p.xres_f + r, ep_resid, sm.wd);
#endif
              resid_v = lo_f(resid_w);
            }
            st_flag(p.xres_f + r, v + resid_v, ep);
          } else if (mode == M_GELU) {  // main.zig:80
            st_flag(p.f_f + r, gelu_ref(v), ep);
          } else {  // tied lm_head (main.zig:193) + running argmax; this lane sees increasing r, so strict >
            if (logits) logits[r] = v;
            if (v > best) { best = v; best_i = (unsigned)r; }
          }
        }
      }
      pf.mark(tag + 3);

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
          sm.red[44 + warp] = best;
          sm.red[52 + warp] = __uint_as_float(best_i);
        }
        consumer_sync();
        if (tid < 32) {
          best = (lane < NCW) ? sm.red[44 + lane] : -INFINITY;
          best_i = (lane < NCW) ? __float_as_uint(sm.red[52 + lane]) : 0xffffffffu;
#pragma unroll
          for (int o = 4; o > 0; o >>= 1) {
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
      vsel ^= 1;
      float *vec = sm.vec + vsel * E4;
      gather_flagged(vec, p.xres_f, E, ep, sm.wd);
      consumer_sync();
      if (warp == 0) {
        float4 xs[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int i4 = lane + 32 * j;
          xs[j] = (i4 < Eq) ? reinterpret_cast<const float4 *>(vec)[i4] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        ln_regs<NJ>(xs, reinterpret_cast<const float4 *>(p.lnf_g), reinterpret_cast<const float4 *>(p.lnf_b), E, lane);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int i4 = lane + 32 * j;
          if (i4 < Eq) {
            reinterpret_cast<float4 *>(p.xout)[i4] = xs[j];
            reinterpret_cast<float4 *>(p.xres_out)[i4] = reinterpret_cast<const float4 *>(vec)[i4];
          }
        }
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

typedef void (*decode_kernel_t)(const DecodeParams);
// smallest instantiation whose register copy covers E (E <= 128 NJ)
static decode_kernel_t decode_kernel_for(int E) {
  if (E <= 256) return decode_persistent_kernel<2>;
  if (E <= 768) return decode_persistent_kernel<6>;
  if (E <= 1024) return decode_persistent_kernel<8>;
  if (E <= 1280) return decode_persistent_kernel<10>;
  return decode_persistent_kernel<13>;
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
  const size_t floats = (size_t)nslot * 4 * E + 2 * 4 * E + 2 * E + NCW * hd + 64;
  return floats * sizeof(float) + (5 * c.n_layer + 1) * sizeof(PhaseEnt);
}

extern "C" {

zg_engine *zg_engine_create(const zg_gpt *gpt, const zg_state *state) {
  if (!require_ready("zg_engine_create")) return nullptr;
  Context &c = ctx();
  const zg_config &cfg = gpt->config;
  const size_t E = cfg.n_embed;
  if (E % 8 != 0 || cfg.n_heads * 64 != E || E > 1664 || cfg.n_layer > (size_t)MAX_LAYERS) {
    set_error(1, "zg_engine_create: needs head_dim 64, n_embed % 8 == 0, n_embed <= 1664, n_layer <= 64", __FILE__, __LINE__);
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
  const decode_kernel_t kern = decode_kernel_for((int)E);
  ZG_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes));
  int per_sm = 0;
  ZG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)kern, NTHREADS, e->smem_bytes));
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
