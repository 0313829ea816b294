// zg_batch.cu -- GPT.forward / generate (main.zig:178-207, 322-342) for B independent sequences at once.
//
// The reference forwards one token of one sequence per call; sequences never interact (no cross-sequence op exists
// in ops.zig / main.zig), so B sequences are B rows of every activation matrix and the Linear layers become GEMMs:
//   * decode step   rows = B (one new token per sequence, all at the same position): tcgen05 kind::tf32 GEMMs that read
//                   the reference's fp32 weights in place, fp32 KV caches, batched single-query attention; the
//                   whole step is one CUDA graph whose kernels read the position from a device word.
//   * prefill       rows = B*T (whole prompts): f16 operand copies of the weights (made once at create), kind::f16
//                   GEMMs with fused bias / GELU / residual / KV-cache-append epilogues, causal flash attention on the
//                   tensor cores; replaces the reference's token-at-a-time prompt loop (main.zig:331-334).
// Nothing here allocates after zg_batch_create.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "zg_attn.cuh"
#include "zg_gemm.cuh"
#include "zg_skinny.cuh"

namespace zg {
void launch_sample_rows(const float *logits, size_t pitch, int V, const void *sample_params_dev, const int *step_dev, int step,
                        unsigned long long *tok, unsigned long long *hist, int B, unsigned long long *host_ring);

namespace {

typedef unsigned long long u64;

// x[m,:] = wte[tok(m)] + wpe[pos(m)]  (main.zig:179-183).  decode: one row per sequence, pos = *pos_dev;
// prefill: row m = b*T + t, pos = t.
__global__ void embed_rows_kernel(const float *__restrict__ wte, const float *__restrict__ wpe,
                                  const u64 *__restrict__ tok, int T, const int *pos_dev, int E, int V,
                                  float *__restrict__ x) {
  const int m = blockIdx.x;
  size_t token = (size_t)tok[m];
  if (token >= (size_t)V) token = 0;  // host-supplied ids never index wte out of bounds (same clamp as the batch-1 engine)
  const int pos = pos_dev ? *pos_dev : (m % T);
  const float4 *a = reinterpret_cast<const float4 *>(wte + token * E), *p = reinterpret_cast<const float4 *>(wpe + (size_t)pos * E);
  float4 *o = reinterpret_cast<float4 *>(x + (size_t)m * E);
  for (int i = threadIdx.x; i < (E >> 2); i += blockDim.x) {
    const float4 u = __ldg(a + i), v = __ldg(p + i);
    o[i] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
  }
}

// LayerNorm.forward (ops.zig:82-104: single pass sums of x and x^2, eps inside the square root, division) over rows
// that may be strided in the input (row r starts at in + r * in_stride); output dense, fp32 or fp16.  One warp per
// row: the row is read once with 128-bit loads and stays in registers between the statistics and the normalisation,
// reductions are warp shuffles, stores are 128-bit (fp32) / 64-bit (fp16).  HBM-bound: E*4 bytes in, E*(4|2) out.
constexpr int LN_WARPS = 8;
// LN_MAXV float4 per lane (n_embed <= 128 LN_MAXV): 8 keeps the kernel at ~60 registers, i.e. 32 resident warps per SM
template <bool OUT_F16, int LN_MAXV>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_rows_kernel(const float *__restrict__ in, size_t in_stride, void *out,
                                                                const float *__restrict__ g, const float *__restrict__ b,
                                                                int E, float eps, int rows) {
  const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4 *src = reinterpret_cast<const float4 *>(in + (size_t)row * in_stride);
  const int nv = E >> 2;
  float4 v[LN_MAXV];
  float s = 0.0f, ss = 0.0f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = i * 32 + lane;
    if (c < nv) {
      v[i] = src[c];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      ss = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, fmaf(v[i].w, v[i].w, ss))));
    }
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  const float n = (float)E, mean = s / n;
  const float std_ = sqrtf(ss / n - mean * mean + eps);
  const float rinv = 1.0f / std_;
  const float4 *g4 = reinterpret_cast<const float4 *>(g), *b4 = reinterpret_cast<const float4 *>(b);
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = i * 32 + lane;
    if (c < nv) {
      const float4 gg = __ldg(g4 + c), bb = __ldg(b4 + c);
      float4 y;
      if (OUT_F16) {  // the f16 operand rounds at 2^-11: one reciprocal per row instead of four divisions per float4
        y.x = (v[i].x - mean) * rinv * gg.x + bb.x;
        y.y = (v[i].y - mean) * rinv * gg.y + bb.y;
        y.z = (v[i].z - mean) * rinv * gg.z + bb.z;
        y.w = (v[i].w - mean) * rinv * gg.w + bb.w;
      } else {  // fp32 path: the reference's division (ops.zig:101), bit-for-bit the per-op kernel's arithmetic
        y.x = (v[i].x - mean) / std_ * gg.x + bb.x;
        y.y = (v[i].y - mean) / std_ * gg.y + bb.y;
        y.z = (v[i].z - mean) / std_ * gg.z + bb.z;
        y.w = (v[i].w - mean) / std_ * gg.w + bb.w;
      }
      if (OUT_F16) {
        __half2 lo = __floats2half2_rn(y.x, y.y), hi = __floats2half2_rn(y.z, y.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t *>(&lo);
        pk.y = *reinterpret_cast<uint32_t *>(&hi);
        reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(out) + (size_t)row * E)[c] = pk;
      } else {
        reinterpret_cast<float4 *>(reinterpret_cast<float *>(out) + (size_t)row * E)[c] = y;
      }
    }
  }
}

template <bool OUT_F16>
void launch_ln_rows(const float *in, size_t in_stride, void *out, const float *g, const float *b, int E, int rows, cudaStream_t s) {
  const int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  if (E <= 1024) ln_rows_kernel<OUT_F16, 8><<<grid, LN_WARPS * 32, 0, s>>>(in, in_stride, out, g, b, E, 1e-5f, rows);
  else ln_rows_kernel<OUT_F16, 16><<<grid, LN_WARPS * 32, 0, s>>>(in, in_stride, out, g, b, E, 1e-5f, rows);
  ZG_LAUNCH_CHECK();
}

// LayerNorm.forward of row r of x into h (ops.zig:82-104: single pass sums of x and x^2, eps inside the square root,
// division) AND zero-fill of row r of the output that the next (stream-K) GEMM reduces its partial sums into.  One CTA of
// four warps per row -- the decode step has <= 128 rows, so the row is split over 128 threads instead of giving a single
// warp ~2,000 dependent instructions (the one-warp-per-row kernel took 11-12 us per call at E = 1600).
constexpr int LNZ_THREADS = 128;
template <int LN_MAXV, bool OUT_F16>  // float4 per thread: n_embed <= 4 * LNZ_THREADS * LN_MAXV
// (`in` is NOT __restrict__: see the PDL note in zg_common.cuh -- a const __restrict__ load may be hoisted above pdl_wait)
__global__ void __launch_bounds__(LNZ_THREADS) ln_zero_rows_kernel(const float *in, void *__restrict__ out,
                                                                   const float *__restrict__ g, const float *__restrict__ b,
                                                                   int E, float eps, float *__restrict__ zero, int zero_n,
                                                                   int trigger) {
  __shared__ float red[2][LNZ_THREADS / 32];
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (trigger) pdl_trigger();
  const int nv = E >> 2;
  const float4 *g4 = reinterpret_cast<const float4 *>(g), *b4 = reinterpret_cast<const float4 *>(b);
  pdl_wait();  // x is the previous GEMM's output; the buffer zeroed below may still be read by it
  const float4 *src = reinterpret_cast<const float4 *>(in + (size_t)row * E);
  float4 v[LN_MAXV];
  float s = 0.0f, ss = 0.0f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = i * LNZ_THREADS + tid;
    if (c < nv) {
      v[i] = src[c];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      ss = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, fmaf(v[i].z, v[i].z, fmaf(v[i].w, v[i].w, ss))));
    }
  }
  if (zero) {  // independent of the row's statistics: issued while the loads above are in flight
    float4 *z = reinterpret_cast<float4 *>(zero + (size_t)row * zero_n);
    for (int c = tid; c < (zero_n >> 2); c += LNZ_THREADS) z[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  if (lane == 0) { red[0][warp] = s; red[1][warp] = ss; }
  __syncthreads();
  s = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
  ss = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
  const float n = (float)E, mean = s / n;
  const float std_ = sqrtf(ss / n - mean * mean + eps);
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int c = i * LNZ_THREADS + tid;
    if (c < nv) {
      const float4 gg = __ldg(g4 + c), bb = __ldg(b4 + c);
      float4 y;
      y.x = (v[i].x - mean) / std_ * gg.x + bb.x;
      y.y = (v[i].y - mean) / std_ * gg.y + bb.y;
      y.z = (v[i].z - mean) / std_ * gg.z + bb.z;
      y.w = (v[i].w - mean) / std_ * gg.w + bb.w;
      if (OUT_F16) {
        __half2 lo = __floats2half2_rn(y.x, y.y), hi = __floats2half2_rn(y.z, y.w);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t *>(&lo);
        pk.y = *reinterpret_cast<uint32_t *>(&hi);
        reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(out) + (size_t)row * E)[c] = pk;
      } else {
        reinterpret_cast<float4 *>(reinterpret_cast<float *>(out) + (size_t)row * E)[c] = y;
      }
    }
  }
}
template <bool OUT_F16>
void launch_ln_zero_rows(const float *in, void *out, const float *g, const float *b, int E, int rows, float *zero, int zero_n,
                         cudaStream_t s) {
  const int trig = (pdl_mask() & PDL_LN_TRIGGER) ? 1 : 0;
  if (E <= 1024) ZG_CUDA(launch_pdl(PDL_LN_DEP, ln_zero_rows_kernel<2, OUT_F16>, dim3(rows), dim3(LNZ_THREADS), 0, s, in, out, g, b, E, 1e-5f, zero, zero_n, trig));
  else ZG_CUDA(launch_pdl(PDL_LN_DEP, ln_zero_rows_kernel<4, OUT_F16>, dim3(rows), dim3(LNZ_THREADS), 0, s, in, out, g, b, E, 1e-5f, zero, zero_n, trig));
  ZG_LAUNCH_CHECK();
}

// pre := gelu(pre) in place (main.zig:80), exact tanhf GELU (ops.zig:225): the stream-K decode step applies it here, once
// per element, instead of in mlp c_proj's operand load -- there every weight tile re-applies it to the X chunk it
// multiplies (12.5x redundant at 1.5B), and the single-pass TF32 kernel needs its transform warps only for that.
// cfg 4: 8.39 -> 8.18 ms (TF32), 8.94 -> 8.78 ms (3xTF32).
__global__ void __launch_bounds__(256) gelu_inplace_kernel(float *pre, size_t n4) {
  pdl_trigger();
  pdl_wait();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<float4 *>(pre)[i];
    v.x = gelu_ref(v.x); v.y = gelu_ref(v.y); v.z = gelu_ref(v.z); v.w = gelu_ref(v.w);
    reinterpret_cast<float4 *>(pre)[i] = v;
  }
}

// h16 = f16(gelu(pre)): the GELU between c_fc and mlp c_proj (main.zig:80) of the 16-bit decode step, where c_proj's f16
// operand cannot be produced by c_fc's epilogue (stream-K partial sums).  Exact tanhf GELU (ops.zig:225).
__global__ void __launch_bounds__(256) gelu_to_f16_kernel(const float *pre, __half *out, size_t n4) {  // (no __restrict__: PDL note)
  pdl_trigger();
  pdl_wait();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4 *>(pre)[i];
    __half2 a = __floats2half2_rn(gelu_ref(v.x), gelu_ref(v.y)), b = __floats2half2_rn(gelu_ref(v.z), gelu_ref(v.w));
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t *>(&a);
    pk.y = *reinterpret_cast<uint32_t *>(&b);
    reinterpret_cast<uint2 *>(out)[i] = pk;
  }
}

// greedy argmax per row (first maximum wins, like the oracle); writes the next token and the history row
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float *__restrict__ logits, size_t pitch, int V,
                                                          u64 *__restrict__ tok, u64 *__restrict__ hist, int B,
                                                          const int *pos_dev) {
  __shared__ float sv[8];
  __shared__ int si[8];
  const float *row = logits + (size_t)blockIdx.x * pitch;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float v = row[i];
    if (v > best) { best = v; bi = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sv[w] = best; si[w] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k)
      if (sv[k] > best || (sv[k] == best && si[k] < bi)) { best = sv[k]; bi = si[k]; }
    tok[blockIdx.x] = (u64)bi;
    if (hist) hist[(size_t)(*pos_dev) * B + blockIdx.x] = (u64)bi;
  }
}

// prompt step: tok[b] = prompts[b][s], hist[s][b] = tok[b]  with s = *pos_dev
__global__ void load_prompt_tokens_kernel(const u64 *__restrict__ prompts, int n_inputs, u64 *__restrict__ tok,
                                          u64 *__restrict__ hist, int B, const int *pos_dev) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int s = *pos_dev;
  const u64 t = prompts[(size_t)b * n_inputs + s];
  tok[b] = t;
  if (hist) hist[(size_t)s * B + b] = t;
}
// after a prefill: hist[s][b] = prompts[b][s] for every prompt position, tok[b] = last prompt token
__global__ void prompts_to_hist_kernel(const u64 *__restrict__ prompts, int n_inputs, u64 *__restrict__ tok,
                                       u64 *__restrict__ hist, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * n_inputs) return;
  const int b = i / n_inputs, s = i % n_inputs;
  const u64 t = prompts[i];
  hist[(size_t)s * B + b] = t;
  if (s == n_inputs - 1) tok[b] = t;
}
__global__ void set_pos_kernel(int *pos, int v, int add) { *pos = add ? *pos + v : v; }

struct LayerW {
  const float *ln1_g, *ln1_b, *attn_w, *attn_b, *proj_w, *proj_b, *ln2_g, *ln2_b, *fc_w, *fc_b, *proj2_w, *proj2_b;
  const __half *attn_w16, *proj_w16, *fc_w16, *proj2_w16;
};

struct LayerPlans {
  GemmPlan attn, proj, fc, proj2;
};
struct SkinnyLayerPlans {
  SkinnyPlan attn, proj, fc, proj2;
};

}  // namespace

}  // namespace zg

using namespace zg;

struct zg_batch {
  zg_config cfg;
  int B = 0, cap = 0, max_prompt = 0, Vp = 0;
  bool f16_prefill = false;   // prefill enabled (max_prompt > 0)
  bool exact_prefill = false;  // ... with 3xTF32 GEMMs on fp32 activations and fp32 causal attention instead of the f16 pipeline
  const float *wte = nullptr, *wpe = nullptr, *lnf_g = nullptr, *lnf_b = nullptr;
  const __half *wte16 = nullptr;
  std::vector<LayerW> layers;
  // caches: per layer [B][cap][E]
  float *k_cache = nullptr, *v_cache = nullptr;
  size_t layer_stride = 0, seq_stride = 0;
  // decode-step activations (rows = B)
  float *x = nullptr, *h = nullptr, *qkv = nullptr, *att = nullptr, *h4 = nullptr, *logits = nullptr;
  // prefill activations (rows = B * max_prompt)
  float *px = nullptr, *plast = nullptr;
  __half *ph = nullptr, *pqkv = nullptr, *patt = nullptr, *ph4 = nullptr, *plast16 = nullptr;
  float *ph32 = nullptr, *pqkv32 = nullptr, *patt32 = nullptr, *ph4_32 = nullptr;  // exact_prefill
  u64 *tok = nullptr, *hist = nullptr, *prompts = nullptr, *ptok = nullptr;
  u64 *hist_host = nullptr;  // pinned [hist_cap][B]: generate() reads the token history back through it
  size_t hist_cap = 0;
  int *pos = nullptr;
  // plans
  std::vector<LayerPlans> dec_plans, pre_plans;
  std::vector<SkinnyLayerPlans> sk_plans;  // decode step with <= 128 sequences: swapped-operand stream-K GEMMs
  SkinnyPlan sk_head;                      // tied lm_head with the argmax in its epilogue (no logits leave the kernel)
  unsigned long long *best = nullptr;      // [B][2] packed (orderable logit, ~column) words of the fused argmax
  void *samp = nullptr;                    // device {temp, seed, seq_base} of the sampling generate loop
  cudaGraphExec_t graph_sampling = nullptr;
  bool skinny = false;
  // 16-bit storage (flags bit 4): f16 copies of every weight, f16 KV caches, f16 operands between the kernels of the
  // stream-K decode step; fp32 residual stream and fp32 accumulation throughout
  bool store16 = false;
  __half *h16 = nullptr, *att16 = nullptr, *h4_16 = nullptr;
  void *k_cache16 = nullptr, *v_cache16 = nullptr;
  SkinnyPlan sk_head_logits;  // lm_head into a zeroed logits buffer (compute_logits = 1)
  GemmPlan dec_head, dec_head_best, pre_head;
  std::vector<AttnPrefillPlan> pre_attn;
  int pre_T = -1;
  cudaGraphExec_t graph_sample = nullptr, graph_prompt = nullptr;
  int graph_n_inputs = -1;
  int host_pos = 0;  // host mirror of *pos (every entry point that moves the position updates it)
  bool use_graph = true;
  int dec_mode = 2;  // decode-step GEMMs: 2 = 3xTF32 (fp32-class accuracy), 1 = single-pass TF32
  std::vector<void *> owned;
};

namespace {

template <typename T>
T *balloc(zg_batch *e, size_t n) {
  void *p = zg_alloc(n * sizeof(T));
  if (p) e->owned.push_back(p);
  return (T *)p;
}

const __half *f16_copy(zg_batch *e, const float *src, size_t n) {
  __half *d = balloc<__half>(e, n);
  if (d) zg_to_f16(src, d, n);
  return d;
}

GemmArgs base_args(int M, int N, int K, const float *bias, void *out, int ldo, int out_f16) {
  GemmArgs a;
  a.M = M; a.N = N; a.K = K; a.bias = bias; a.out = out; a.ldo = ldo; a.out_f16 = out_f16;
  return a;
}

bool build_decode_plans(zg_batch *e) {
  const int B = e->B, E = (int)e->cfg.n_embed, V = (int)e->cfg.vocab_size;
  e->dec_plans.resize(e->layers.size());
  for (size_t l = 0; l < e->layers.size(); ++l) {
    const LayerW &w = e->layers[l];
    LayerPlans &p = e->dec_plans[l];
    GemmArgs a = base_args(B, 3 * E, E, w.attn_b, e->qkv, 3 * E, 0);
    a.k_cache = e->k_cache + l * e->layer_stride;
    a.v_cache = e->v_cache + l * e->layer_stride;
    a.E = E; a.rows_per_seq = 1; a.cache_seq_stride = (long long)e->seq_stride; a.pos_dev = e->pos; a.cache_rows = e->cap;
    if (!gemm_plan(&p.attn, e->dec_mode, e->h, E, w.attn_w, a, 0)) return false;
    a = base_args(B, E, E, w.proj_b, e->x, E, 0);
    a.epi = TC_EPI_RESIDUAL; a.resid = e->x; a.ldr = E;
    if (!gemm_plan(&p.proj, e->dec_mode, e->att, E, w.proj_w, a, 0)) return false;
    a = base_args(B, 4 * E, E, w.fc_b, e->h4, 4 * E, 0);
    a.epi = TC_EPI_GELU;
    if (!gemm_plan(&p.fc, e->dec_mode, e->h, E, w.fc_w, a, 0)) return false;
    a = base_args(B, E, 4 * E, w.proj2_b, e->x, E, 0);
    a.epi = TC_EPI_RESIDUAL; a.resid = e->x; a.ldr = E;
    if (!gemm_plan(&p.proj2, e->dec_mode, e->h4, 4 * E, w.proj2_w, a, 0)) return false;
  }
  GemmArgs a = base_args(B, V, E, nullptr, e->logits, e->Vp, 0);
  if (!gemm_plan(&e->dec_head, e->dec_mode, e->h, E, e->wte, a, 0)) return false;
  a.best = e->best;  // greedy steps: the argmax rides in the lm_head's epilogue, no logits are written
  return gemm_plan(&e->dec_head_best, e->dec_mode, e->h, E, e->wte, a, 0);
}

// Decode step with <= 128 sequences: every layer GEMM through the stream-K kernel.  c_attn reduces into a zeroed qkv
// (its K/V columns are appended to the caches by the attention kernel), c_fc into a zeroed pre-activation buffer whose
// GELU is applied by mlp c_proj's operand load, both c_proj's into the residual stream in place.
bool build_skinny_plans16(zg_batch *e) {
  const int B = e->B, E = (int)e->cfg.n_embed;
  e->sk_plans.resize(e->layers.size());
  for (size_t l = 0; l < e->layers.size(); ++l) {
    const LayerW &w = e->layers[l];
    SkinnyLayerPlans &p = e->sk_plans[l];
    SkinnyArgs a;
    a.M = B; a.N = 3 * E; a.K = E; a.bias = w.attn_b; a.out = e->qkv; a.ldo = 3 * E;
    if (!skinny_plan(&p.attn, 0, e->h16, E, w.attn_w16, a)) return false;
    a = SkinnyArgs();
    a.M = B; a.N = E; a.K = E; a.bias = w.proj_b; a.out = e->x; a.ldo = E;
    if (!skinny_plan(&p.proj, 0, e->att16, E, w.proj_w16, a)) return false;
    a = SkinnyArgs();
    a.M = B; a.N = 4 * E; a.K = E; a.bias = w.fc_b; a.out = e->h4; a.ldo = 4 * E;
    if (!skinny_plan(&p.fc, 0, e->h16, E, w.fc_w16, a)) return false;
    a = SkinnyArgs();
    a.M = B; a.N = E; a.K = 4 * E; a.bias = w.proj2_b; a.out = e->x; a.ldo = E;
    if (!skinny_plan(&p.proj2, 0, e->h4_16, 4 * E, w.proj2_w16, a)) return false;
  }
  SkinnyArgs a;
  a.M = B; a.N = (int)e->cfg.vocab_size; a.K = E; a.best = e->best;
  if (!skinny_plan(&e->sk_head, 0, e->h16, E, e->wte16, a)) return false;
  a = SkinnyArgs();
  a.M = B; a.N = (int)e->cfg.vocab_size; a.K = E; a.out = e->logits; a.ldo = e->Vp;
  return skinny_plan(&e->sk_head_logits, 0, e->h16, E, e->wte16, a);
}

static bool gelu_separate(const zg_batch *e) {  // ZG_GELU_FUSED=1: A/B switch back to the operand-load GELU
  static const bool fused = getenv("ZG_GELU_FUSED") != nullptr;
  (void)e;
  return !fused;
}

bool build_skinny_plans(zg_batch *e) {
  if (e->store16) return build_skinny_plans16(e);
  const int B = e->B, E = (int)e->cfg.n_embed;
  e->sk_plans.resize(e->layers.size());
  for (size_t l = 0; l < e->layers.size(); ++l) {
    const LayerW &w = e->layers[l];
    SkinnyLayerPlans &p = e->sk_plans[l];
    SkinnyArgs a;
    a.M = B; a.N = 3 * E; a.K = E; a.bias = w.attn_b; a.out = e->qkv; a.ldo = 3 * E;
    if (!skinny_plan(&p.attn, e->dec_mode, e->h, E, w.attn_w, a)) return false;
    a = SkinnyArgs();
    a.M = B; a.N = E; a.K = E; a.bias = w.proj_b; a.out = e->x; a.ldo = E;
    if (!skinny_plan(&p.proj, e->dec_mode, e->att, E, w.proj_w, a)) return false;
    a = SkinnyArgs();
    a.M = B; a.N = 4 * E; a.K = E; a.bias = w.fc_b; a.out = e->h4; a.ldo = 4 * E;
    if (!skinny_plan(&p.fc, e->dec_mode, e->h, E, w.fc_w, a)) return false;
    a = SkinnyArgs();
    // the GELU is a separate in-place pass over h4 (gelu_inplace_kernel); SK_XFORM_GELU in the operand load is the A/B form
    a.M = B; a.N = E; a.K = 4 * E; a.bias = w.proj2_b; a.out = e->x; a.ldo = E; a.xform = gelu_separate(e) ? 0 : SK_XFORM_GELU;
    if (!skinny_plan(&p.proj2, e->dec_mode, e->h4, 4 * E, w.proj2_w, a)) return false;
  }
  SkinnyArgs a;
  a.M = B; a.N = (int)e->cfg.vocab_size; a.K = E; a.best = e->best;
  return skinny_plan(&e->sk_head, e->dec_mode, e->h, E, e->wte, a);
}

// fp32-class prefill: the same four GEMMs per layer in the 3xTF32 mode on fp32 activations; logits through the decode
// head (e->h -> e->logits).  Token-identical to the reference's token-at-a-time prompt loop wherever the decode step is.
bool build_exact_prefill_plans(zg_batch *e, int T) {
  const int B = e->B, E = (int)e->cfg.n_embed, M = B * T;
  e->pre_plans.resize(e->layers.size());
  for (size_t l = 0; l < e->layers.size(); ++l) {
    const LayerW &w = e->layers[l];
    LayerPlans &p = e->pre_plans[l];
    GemmArgs a = base_args(M, 3 * E, E, w.attn_b, e->pqkv32, 3 * E, 0);
    a.k_cache = e->k_cache + l * e->layer_stride;
    a.v_cache = e->v_cache + l * e->layer_stride;
    a.E = E; a.rows_per_seq = T; a.cache_seq_stride = (long long)e->seq_stride; a.pos_dev = nullptr; a.pos_base = 0; a.cache_rows = e->cap;
    if (!gemm_plan(&p.attn, 2, e->ph32, E, w.attn_w, a, 0)) return false;
    a = base_args(M, E, E, w.proj_b, e->px, E, 0);
    a.epi = TC_EPI_RESIDUAL; a.resid = e->px; a.ldr = E;
    if (!gemm_plan(&p.proj, 2, e->patt32, E, w.proj_w, a, 0)) return false;
    a = base_args(M, 4 * E, E, w.fc_b, e->ph4_32, 4 * E, 0);
    a.epi = TC_EPI_GELU;
    if (!gemm_plan(&p.fc, 2, e->ph32, E, w.fc_w, a, 0)) return false;
    a = base_args(M, E, 4 * E, w.proj2_b, e->px, E, 0);
    a.epi = TC_EPI_RESIDUAL; a.resid = e->px; a.ldr = E;
    if (!gemm_plan(&p.proj2, 2, e->ph4_32, 4 * E, w.proj2_w, a, 0)) return false;
  }
  e->pre_T = T;
  return true;
}

bool build_prefill_plans(zg_batch *e, int T) {
  if (e->pre_T == T) return true;
  if (e->exact_prefill) return build_exact_prefill_plans(e, T);
  const int B = e->B, E = (int)e->cfg.n_embed, V = (int)e->cfg.vocab_size, M = B * T, H = (int)e->cfg.n_heads;
  e->pre_plans.resize(e->layers.size());
  e->pre_attn.resize(e->layers.size());
  for (size_t l = 0; l < e->layers.size(); ++l) {
    const LayerW &w = e->layers[l];
    LayerPlans &p = e->pre_plans[l];
    GemmArgs a = base_args(M, 3 * E, E, w.attn_b, e->pqkv, 3 * E, 1);
    a.k_cache = e->k_cache + l * e->layer_stride;
    a.v_cache = e->v_cache + l * e->layer_stride;
    a.E = E; a.rows_per_seq = T; a.cache_seq_stride = (long long)e->seq_stride; a.pos_dev = nullptr; a.pos_base = 0; a.cache_rows = e->cap;
    if (!gemm_plan(&p.attn, 0, e->ph, E, w.attn_w16, a, 0)) return false;
    if (!attn_prefill_plan(&e->pre_attn[l], e->pqkv, e->patt, B, T, H, E)) return false;
    a = base_args(M, E, E, w.proj_b, e->px, E, 0);
    a.epi = TC_EPI_RESIDUAL; a.resid = e->px; a.ldr = E;
    if (!gemm_plan(&p.proj, 0, e->patt, E, w.proj_w16, a, 0)) return false;
    a = base_args(M, 4 * E, E, w.fc_b, e->ph4, 4 * E, 1);
    a.epi = TC_EPI_GELU; a.gelu_fast = 1;
    if (!gemm_plan(&p.fc, 0, e->ph, E, w.fc_w16, a, 0)) return false;
    a = base_args(M, E, 4 * E, w.proj2_b, e->px, E, 0);
    a.epi = TC_EPI_RESIDUAL; a.resid = e->px; a.ldr = E;
    if (!gemm_plan(&p.proj2, 0, e->ph4, 4 * E, w.proj2_w16, a, 0)) return false;
  }
  GemmArgs a = base_args(B, V, E, nullptr, e->logits, e->Vp, 0);
  if (!gemm_plan(&e->pre_head, 0, e->plast16, E, e->wte16, a, 0)) return false;
  e->pre_T = T;
  return true;
}

// One decode step for every sequence at position *pos: GPT.forward(seq_len = *pos + 1, tok[b]) (main.zig:178-195).
// head: 0 = no logits, 1 = logits in e->logits + argmax over them, 2 = argmax only, fused into the lm_head GEMM when the
// stream-K path is active (greedy generate / run_steps never need the logits themselves), 3 = logits + sampled token
// The greedy lm_head of a stream-K engine runs on the GENERAL kernel: batch rows = TMEM lanes, so the argmax is a per-thread
// running maximum and one atomicMax per row and tile, and the accumulator is double-buffered.  The swapped stream-K kernel
// (vocabulary rows = lanes: two warp reductions per batch column and tile, single accumulator) measured 94.8 vs 54.9 us at
// 124M / 128 sequences and 76.7 vs 65.9 us (TF32), 104.2 vs 97.7 us (3xTF32) at 1.5B / 64.  ZG_HEAD_SKINNY=1 selects it (A/B).
static bool head_general(int) {
  static const bool skinny_head = getenv("ZG_HEAD_SKINNY") != nullptr;
  return !skinny_head;
}

void enqueue_step(zg_batch *e, bool from_prompt, int n_inputs, int head) {
  cudaStream_t s = ctx().stream;
  const int B = e->B, E = (int)e->cfg.n_embed, H = (int)e->cfg.n_heads, V = (int)e->cfg.vocab_size;
  if (from_prompt) {
    load_prompt_tokens_kernel<<<(B + 127) / 128, 128, 0, s>>>(e->prompts, n_inputs, e->tok, e->hist, B, e->pos);
    ZG_LAUNCH_CHECK();
  }
  embed_rows_kernel<<<B, 128, 0, s>>>(e->wte, e->wpe, e->tok, 1, e->pos, E, V, e->x);
  ZG_LAUNCH_CHECK();
  for (size_t l = 0; e->store16 && l < e->layers.size(); ++l) {  // 16-bit storage: same step, f16 operands and caches
    const LayerW &w = e->layers[l];
    const SkinnyLayerPlans &p = e->sk_plans[l];
    launch_ln_zero_rows<true>(e->x, e->h16, w.ln1_g, w.ln1_b, E, B, e->qkv, 3 * E, s);
    skinny_launch(p.attn);
    attn_decode_batch_launch_f16(e->qkv, 3 * E, (const __half *)e->k_cache16 + l * e->layer_stride,
                                 (const __half *)e->v_cache16 + l * e->layer_stride, (long long)e->seq_stride, B, H, E, e->att16, E,
                                 e->pos, e->qkv + E, e->qkv + 2 * E);
    skinny_launch(p.proj);
    launch_ln_zero_rows<true>(e->x, e->h16, w.ln2_g, w.ln2_b, E, B, e->h4, 4 * E, s);
    skinny_launch(p.fc);
    ZG_CUDA(launch_pdl(PDL_LN_DEP, gelu_to_f16_kernel, dim3(2 * ctx().sm_count), dim3(256), 0, s, (const float *)e->h4, e->h4_16,
                       (size_t)B * E));  // B * 4E / 4 float4
    ZG_LAUNCH_CHECK();
    skinny_launch(p.proj2);
  }
  for (size_t l = 0; e->skinny && !e->store16 && l < e->layers.size(); ++l) {
    const LayerW &w = e->layers[l];
    const SkinnyLayerPlans &p = e->sk_plans[l];
    launch_ln_zero_rows<false>(e->x, e->h, w.ln1_g, w.ln1_b, E, B, e->qkv, 3 * E, s);  // main.zig:121-123; qkv := 0
    skinny_launch(p.attn);                                                       // c_attn (ops.zig:143)
    attn_decode_batch_launch(e->qkv, 3 * E, e->k_cache + l * e->layer_stride, e->v_cache + l * e->layer_stride,
                             (long long)e->seq_stride, B, H, E, e->att, E, e->pos, 0, e->qkv + E, e->qkv + 2 * E);
    skinny_launch(p.proj);                                                       // x += c_proj(att) (ops.zig:172, main.zig:136-139)
    launch_ln_zero_rows<false>(e->x, e->h, w.ln2_g, w.ln2_b, E, B, e->h4, 4 * E, s);    // main.zig:140; h4 := 0
    skinny_launch(p.fc);                                                         // c_fc pre-activation (main.zig:79)
    if (gelu_separate(e)) {
      ZG_CUDA(launch_pdl(PDL_LN_DEP, gelu_inplace_kernel, dim3(2 * ctx().sm_count), dim3(256), 0, s, e->h4, (size_t)B * E));  // B * 4E / 4
      ZG_LAUNCH_CHECK();
    }
    skinny_launch(p.proj2);                                                      // x += c_proj(gelu(.)) (main.zig:80-81,142-145)
  }
  for (size_t l = 0; !e->skinny && l < e->layers.size(); ++l) {
    const LayerW &w = e->layers[l];
    const LayerPlans &p = e->dec_plans[l];
    launch_ln_rows<false>(e->x, E, e->h, w.ln1_g, w.ln1_b, E, B, s);  // main.zig:121-123
    gemm_launch(p.attn);  // c_attn + K/V append at row *pos (ops.zig:143,151-152,156-157)
    attn_decode_batch_launch(e->qkv, 3 * E, e->k_cache + l * e->layer_stride, e->v_cache + l * e->layer_stride,
                             (long long)e->seq_stride, B, H, E, e->att, E, e->pos, 0);  // ops.zig:160-169
    gemm_launch(p.proj);  // c_proj + residual (ops.zig:172, main.zig:136-139)
    launch_ln_rows<false>(e->x, E, e->h, w.ln2_g, w.ln2_b, E, B, s);  // main.zig:140
    gemm_launch(p.fc);     // c_fc + GELU (main.zig:79-80)
    gemm_launch(p.proj2);  // c_proj + residual (main.zig:81,142-145)
  }
  if (head && e->store16) {
    if (head == 2) {
      launch_ln_zero_rows<true>(e->x, e->h16, e->lnf_g, e->lnf_b, E, B, reinterpret_cast<float *>(e->best), 4, s);
      skinny_launch(e->sk_head);
      skinny_finish_argmax(e->best, e->tok, e->hist, B, e->pos);
    } else {
      launch_ln_zero_rows<true>(e->x, e->h16, e->lnf_g, e->lnf_b, E, B, e->logits, e->Vp, s);  // logits := 0
      skinny_launch(e->sk_head_logits);
      if (head == 3) {
        launch_sample_rows(e->logits, (size_t)e->Vp, V, e->samp, e->pos, 0, e->tok, e->hist, B, nullptr);
      } else {
        argmax_rows_kernel<<<B, 256, 0, s>>>(e->logits, (size_t)e->Vp, V, e->tok, e->hist, B, e->pos);
        ZG_LAUNCH_CHECK();
      }
    }
  } else if (head == 2 && e->skinny && !head_general(B)) {
    launch_ln_zero_rows<false>(e->x, e->h, e->lnf_g, e->lnf_b, E, B, reinterpret_cast<float *>(e->best), 4, s);  // main.zig:189; best := 0
    skinny_launch(e->sk_head);                                         // tied lm_head + argmax (main.zig:192-194)
    skinny_finish_argmax(e->best, e->tok, e->hist, B, e->pos);
  } else if (head == 2) {
    launch_ln_zero_rows<false>(e->x, e->h, e->lnf_g, e->lnf_b, E, B, reinterpret_cast<float *>(e->best), 4, s);  // main.zig:189; best := 0
    gemm_launch(e->dec_head_best);                                     // tied lm_head + argmax (main.zig:192-194)
    skinny_finish_argmax(e->best, e->tok, e->hist, B, e->pos);
  } else if (head) {
    launch_ln_rows<false>(e->x, E, e->h, e->lnf_g, e->lnf_b, E, B, s);  // main.zig:189
    gemm_launch(e->dec_head);  // tied lm_head (main.zig:192-194)
    if (head == 3) {  // GPT.sample (main.zig:198-207): temperature softmax + inverse-CDF draw per sequence, on the device
      launch_sample_rows(e->logits, (size_t)e->Vp, V, e->samp, e->pos, 0, e->tok, e->hist, B, nullptr);
    } else {
      argmax_rows_kernel<<<B, 256, 0, s>>>(e->logits, (size_t)e->Vp, V, e->tok, e->hist, B, e->pos);
      ZG_LAUNCH_CHECK();
    }
  }
  set_pos_kernel<<<1, 1, 0, s>>>(e->pos, 1, 1);
  ZG_LAUNCH_CHECK();
}

void enqueue_prefill(zg_batch *e, int T, bool with_logits) {
  cudaStream_t s = ctx().stream;
  const int B = e->B, E = (int)e->cfg.n_embed, M = B * T;
  embed_rows_kernel<<<M, 128, 0, s>>>(e->wte, e->wpe, e->ptok, T, nullptr, E, (int)e->cfg.vocab_size, e->px);
  ZG_LAUNCH_CHECK();
  if (e->exact_prefill) {
    const int H = (int)e->cfg.n_heads;
    for (size_t l = 0; l < e->layers.size(); ++l) {
      const LayerW &w = e->layers[l];
      const LayerPlans &p = e->pre_plans[l];
      launch_ln_rows<false>(e->px, E, e->ph32, w.ln1_g, w.ln1_b, E, M, s);
      gemm_launch(p.attn);  // c_attn + K/V rows [0, T) of every sequence into the caches
      attn_decode_batch_launch(e->pqkv32, 3 * E, e->k_cache + l * e->layer_stride, e->v_cache + l * e->layer_stride,
                               (long long)e->seq_stride, B, H, E, e->patt32, E, nullptr, 0, nullptr, nullptr, T);
      gemm_launch(p.proj);
      launch_ln_rows<false>(e->px, E, e->ph32, w.ln2_g, w.ln2_b, E, M, s);
      gemm_launch(p.fc);
      gemm_launch(p.proj2);
    }
    if (with_logits) {  // last position of every prompt only (main.zig:192), through the decode head
      launch_ln_rows<false>(e->px + (size_t)(T - 1) * E, (size_t)T * E, e->h, e->lnf_g, e->lnf_b, E, B, s);
      gemm_launch(e->dec_head);
    }
    return;
  }
  for (size_t l = 0; l < e->layers.size(); ++l) {
    const LayerW &w = e->layers[l];
    const LayerPlans &p = e->pre_plans[l];
    launch_ln_rows<true>(e->px, E, e->ph, w.ln1_g, w.ln1_b, E, M, s);
    gemm_launch(p.attn);
    attn_prefill_launch(e->pre_attn[l]);
    gemm_launch(p.proj);
    launch_ln_rows<true>(e->px, E, e->ph, w.ln2_g, w.ln2_b, E, M, s);
    gemm_launch(p.fc);
    gemm_launch(p.proj2);
  }
  if (with_logits) {  // last position of every prompt only (main.zig:192)
    launch_ln_rows<true>(e->px + (size_t)(T - 1) * E, (size_t)T * E, e->plast16, e->lnf_g, e->lnf_b, E, B, s);
    gemm_launch(e->pre_head);
  }
}

bool capture(zg_batch *e, cudaGraphExec_t *exec, bool from_prompt, int n_inputs, int head) {
  cudaStream_t s = ctx().stream;
  cudaGraph_t g = nullptr;
  if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return false;
  ctx().capturing = true;  // recorded, not executed: zg_launch_count counts the replays
  enqueue_step(e, from_prompt, n_inputs, head);
  ctx().capturing = false;
  if (cudaStreamEndCapture(s, &g) != cudaSuccess || !g) {
    cudaGetLastError();
    return false;
  }
  const cudaError_t r = cudaGraphInstantiate(exec, g, 0);
  note_alloc();  // start-up work (first generate / run_steps of an engine), never repeated per step
  cudaGraphDestroy(g);
  return r == cudaSuccess;
}

}  // namespace

extern "C" {

zg_batch *zg_batch_create(const zg_gpt *gpt, size_t n_seqs, size_t cache_rows, size_t max_prompt, int flags) {
  if (!require_ready("zg_batch_create")) return nullptr;
  const zg_config &c = gpt->config;
  const size_t E = c.n_embed, V = c.vocab_size, L = c.n_layer;
  if (E > 2048 || E % 4 != 0) {
    set_error(1, "zg_batch_create: n_embed must be a multiple of 4 and <= 2048 (row kernels keep a row in registers)", __FILE__, __LINE__);
    return nullptr;
  }
  if (E != c.n_heads * 64 || n_seqs == 0 || cache_rows == 0 || cache_rows > c.context_size || max_prompt > cache_rows) {
    set_error(1, "zg_batch_create: head_dim must be 64, 0 < max_prompt <= cache_rows <= context_size", __FILE__, __LINE__);
    return nullptr;
  }
  zg_batch *e = new zg_batch();
  e->cfg = c;
  e->B = (int)n_seqs;
  e->cap = (int)cache_rows;
  e->max_prompt = (int)max_prompt;
  e->Vp = (int)((V + 3) & ~(size_t)3);
  e->f16_prefill = max_prompt > 0;
  e->exact_prefill = e->f16_prefill && (flags & 4);
  e->store16 = (flags & 16) != 0;
  if (e->store16 && (max_prompt > 0 || n_seqs > 128 || E % 64 != 0 || getenv("ZG_NO_SPLIT_K") != nullptr)) {
    set_error(1, "zg_batch_create: 16-bit storage needs n_seqs <= 128, n_embed % 64 == 0, max_prompt == 0 (prompts go token by token)",
              __FILE__, __LINE__);
    delete e;
    return nullptr;
  }
  e->use_graph = !(flags & 1);
  e->dec_mode = (flags & 2) ? 1 : 2;
  e->wte = gpt->wte.weight;
  e->wpe = gpt->wpe.weight;
  e->lnf_g = gpt->ln_f.weight;
  e->lnf_b = gpt->ln_f.bias;
  e->layers.resize(L);
  for (size_t l = 0; l < L; ++l) {
    const zg_block &b = gpt->h[l];
    LayerW &w = e->layers[l];
    w.ln1_g = b.ln_1.weight; w.ln1_b = b.ln_1.bias;
    w.attn_w = b.attn.c_attn.weight; w.attn_b = b.attn.c_attn.bias;
    w.proj_w = b.attn.c_proj.weight; w.proj_b = b.attn.c_proj.bias;
    w.ln2_g = b.ln_2.weight; w.ln2_b = b.ln_2.bias;
    w.fc_w = b.mlp.c_fc.weight; w.fc_b = b.mlp.c_fc.bias;
    w.proj2_w = b.mlp.c_proj.weight; w.proj2_b = b.mlp.c_proj.bias;
    w.attn_w16 = w.proj_w16 = w.fc_w16 = w.proj2_w16 = nullptr;
    if ((e->f16_prefill && !e->exact_prefill) || e->store16) {
      w.attn_w16 = f16_copy(e, w.attn_w, 3 * E * E);
      w.proj_w16 = f16_copy(e, w.proj_w, E * E);
      w.fc_w16 = f16_copy(e, w.fc_w, 4 * E * E);
      w.proj2_w16 = f16_copy(e, w.proj2_w, 4 * E * E);
    }
  }
  if ((e->f16_prefill && !e->exact_prefill) || e->store16) e->wte16 = f16_copy(e, e->wte, V * E);
  const size_t B = n_seqs;
  e->seq_stride = cache_rows * E;
  e->layer_stride = B * e->seq_stride;
  if (e->store16) {
    e->k_cache16 = balloc<__half>(e, L * e->layer_stride);
    e->v_cache16 = balloc<__half>(e, L * e->layer_stride);
    e->h16 = balloc<__half>(e, B * E);
    e->att16 = balloc<__half>(e, B * E);
    e->h4_16 = balloc<__half>(e, B * 4 * E);
    e->k_cache = e->v_cache = balloc<float>(e, 4);  // placeholders: the fp32 caches do not exist in this mode
  } else {
    e->k_cache = balloc<float>(e, L * e->layer_stride);
    e->v_cache = balloc<float>(e, L * e->layer_stride);
  }
  e->x = balloc<float>(e, B * E);
  e->h = balloc<float>(e, B * E);
  e->qkv = balloc<float>(e, B * 3 * E);
  e->att = balloc<float>(e, B * E);
  e->h4 = balloc<float>(e, B * 4 * E);
  e->logits = balloc<float>(e, B * (size_t)e->Vp);
  e->tok = balloc<u64>(e, B);
  e->hist_cap = (size_t)c.context_size;
  e->hist = balloc<u64>(e, e->hist_cap * B);
  e->prompts = balloc<u64>(e, B * (size_t)c.context_size);
  if (cudaHostAlloc(&e->hist_host, e->hist_cap * B * sizeof(u64), cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    e->hist_host = nullptr;
    set_error(1, "zg_batch_create: pinned token history", __FILE__, __LINE__);
  }
  note_alloc();
  e->pos = balloc<int>(e, 4);
  e->best = balloc<unsigned long long>(e, 2 * B);
  e->samp = balloc<unsigned long long>(e, 8);
  if (e->exact_prefill) {
    const size_t M = B * max_prompt;
    e->px = balloc<float>(e, M * E);
    e->ph32 = balloc<float>(e, M * E);
    e->pqkv32 = balloc<float>(e, M * 3 * E);
    e->patt32 = balloc<float>(e, M * E);
    e->ph4_32 = balloc<float>(e, M * 4 * E);
    e->ptok = balloc<u64>(e, M);
  } else if (e->f16_prefill) {
    const size_t M = B * max_prompt;
    e->px = balloc<float>(e, M * E);
    e->ph = balloc<__half>(e, M * E);
    e->pqkv = balloc<__half>(e, M * 3 * E);
    e->patt = balloc<__half>(e, M * E);
    e->ph4 = balloc<__half>(e, M * 4 * E);
    e->plast16 = balloc<__half>(e, B * E);
    e->ptok = balloc<u64>(e, M);
  }
  if (zg_last_error() || !e->k_cache || !e->v_cache || !e->hist) {
    zg_batch_destroy(e);
    return nullptr;
  }
  if (e->store16) {
    zg_memset(e->k_cache16, 0, L * e->layer_stride * sizeof(__half));
    zg_memset(e->v_cache16, 0, L * e->layer_stride * sizeof(__half));
  } else {
    zg_memset(e->k_cache, 0, L * e->layer_stride * sizeof(float));
    zg_memset(e->v_cache, 0, L * e->layer_stride * sizeof(float));
  }
  zg_memset(e->pos, 0, 4 * sizeof(int));
  zg_memset(e->tok, 0, B * sizeof(u64));
  gemm_init_attrs();
  attn_init_attrs();
  skinny_init_attrs();
  // ZG_NO_SPLIT_K=1 asks for run-to-run bit-reproducible steps: no reduction in arrival order anywhere, i.e. no split-K
  // in the general kernel and no stream-K kernel at all
  e->skinny = !(flags & 8) && getenv("ZG_NO_SPLIT_K") == nullptr && skinny_supported(e->B, (int)E, (int)E);
  if ((!e->store16 && !build_decode_plans(e)) || (e->skinny && !build_skinny_plans(e))) {
    zg_batch_destroy(e);
    return nullptr;
  }
  zg_sync();
  return e;
}

void zg_batch_destroy(zg_batch *e) {
  if (!e) return;
  if (ctx().ready) cudaStreamSynchronize(ctx().stream);
  if (e->graph_sample) cudaGraphExecDestroy(e->graph_sample);
  if (e->graph_prompt) cudaGraphExecDestroy(e->graph_prompt);
  if (e->graph_sampling) cudaGraphExecDestroy(e->graph_sampling);
  for (void *p : e->owned) zg_free(p);
  if (e->hist_host) cudaFreeHost(e->hist_host);
  delete e;
}

const float *zg_batch_logits(const zg_batch *e) { return e->logits; }
size_t zg_batch_logits_pitch(const zg_batch *e) { return (size_t)e->Vp; }

// GPT.forward(seq_len, tokens[b], compute_logits) for every sequence b (HOST tokens); logits -> zg_batch_logits().
void zg_batch_forward(zg_batch *e, size_t seq_len, const size_t *tokens, int compute_logits) {
  if (!require_ready("zg_batch_forward")) return;
  if (seq_len == 0 || seq_len > (size_t)e->cap) {
    set_error(1, "zg_batch_forward: seq_len outside the cache", __FILE__, __LINE__);
    return;
  }
  cudaStream_t s = ctx().stream;
  ZG_CUDA(cudaMemcpyAsync(e->tok, tokens, e->B * sizeof(u64), cudaMemcpyHostToDevice, s));
  set_pos_kernel<<<1, 1, 0, s>>>(e->pos, (int)seq_len - 1, 0);
  ZG_LAUNCH_CHECK();
  u64 *hist = e->hist;
  e->hist = nullptr;  // explicit forwards do not record history
  enqueue_step(e, false, 0, compute_logits);
  e->hist = hist;
  e->host_pos = (int)seq_len;
}

// The whole prompt of every sequence at once: tokens[b*T + t] (HOST).  Fills cache rows [0, T) of every block and,
// if asked, the logits of the last position -- what T calls of GPT.forward per sequence leave behind (main.zig:330-334).
int zg_batch_prefill(zg_batch *e, const size_t *tokens, size_t T, int compute_logits) {
  if (!require_ready("zg_batch_prefill")) return 1;
  if (!e->f16_prefill || T == 0 || T > (size_t)e->max_prompt) {
    set_error(1, "zg_batch_prefill: T must be in [1, max_prompt] given to zg_batch_create", __FILE__, __LINE__);
    return 1;
  }
  if (!build_prefill_plans(e, (int)T)) return 1;
  cudaStream_t s = ctx().stream;
  ZG_CUDA(cudaMemcpyAsync(e->ptok, tokens, e->B * T * sizeof(u64), cudaMemcpyHostToDevice, s));
  enqueue_prefill(e, (int)T, compute_logits != 0);
  set_pos_kernel<<<1, 1, 0, s>>>(e->pos, (int)T, 0);
  ZG_LAUNCH_CHECK();
  e->host_pos = (int)T;
  return zg_last_error();
}
// same, tokens already on the device (timing without the upload)
int zg_batch_prefill_resident(zg_batch *e, size_t T, int compute_logits) {
  if (!require_ready("zg_batch_prefill_resident")) return 1;
  if (!e->f16_prefill || T == 0 || T > (size_t)e->max_prompt || !build_prefill_plans(e, (int)T)) return 1;
  enqueue_prefill(e, (int)T, compute_logits != 0);
  return zg_last_error();
}

// generate() (main.zig:322-342), greedy, for B sequences with prompts[b*n_inputs + s] (HOST) of equal length.
// Steps s < n_inputs forward the prompt (one token at a time like the reference, or all at once when use_prefill);
// step n_inputs forwards the last prompt token again (the reference's duplicate, main.zig:329-338); every step's
// token goes to out_tokens[b*n_total + s] (HOST).  Asynchronous until the final copy.
static int batch_generate(zg_batch *e, const size_t *prompts, size_t n_inputs, size_t n_total, size_t *out_tokens,
                          int use_prefill, bool sample) {
  const int B = e->B;
  if (n_inputs == 0 || n_inputs > n_total || n_total > (size_t)e->cap || n_total > e->hist_cap) {
    set_error(1, "zg_batch_generate_greedy: need 1 <= n_inputs <= n_total <= cache rows", __FILE__, __LINE__);
    return 1;
  }
  if (use_prefill && (!e->f16_prefill || n_inputs > (size_t)e->max_prompt)) {
    set_error(1, "zg_batch_generate_greedy: prompt longer than max_prompt", __FILE__, __LINE__);
    return 1;
  }
  cudaStream_t s = ctx().stream;
  ZG_CUDA(cudaMemcpyAsync(e->prompts, prompts, (size_t)B * n_inputs * sizeof(u64), cudaMemcpyHostToDevice, s));
  size_t first = 0;
  if (use_prefill) {
    if (!build_prefill_plans(e, (int)n_inputs)) return 1;
    ZG_CUDA(cudaMemcpyAsync(e->ptok, e->prompts, (size_t)B * n_inputs * sizeof(u64), cudaMemcpyDeviceToDevice, s));
    enqueue_prefill(e, (int)n_inputs, false);
    prompts_to_hist_kernel<<<(unsigned)((B * n_inputs + 255) / 256), 256, 0, s>>>(e->prompts, (int)n_inputs, e->tok,
                                                                                   e->hist, B);
    ZG_LAUNCH_CHECK();
    set_pos_kernel<<<1, 1, 0, s>>>(e->pos, (int)n_inputs, 0);
    ZG_LAUNCH_CHECK();
    first = n_inputs;
  } else {
    set_pos_kernel<<<1, 1, 0, s>>>(e->pos, 0, 0);
    ZG_LAUNCH_CHECK();
  }
  if (e->use_graph) {
    if (!sample && !e->graph_sample && !capture(e, &e->graph_sample, false, 0, 2)) e->use_graph = false;
    if (sample && !e->graph_sampling && !capture(e, &e->graph_sampling, false, 0, 3)) e->use_graph = false;
    if (e->use_graph && first < n_inputs && (e->graph_n_inputs != (int)n_inputs || !e->graph_prompt)) {
      if (e->graph_prompt) cudaGraphExecDestroy(e->graph_prompt);
      e->graph_prompt = nullptr;
      if (capture(e, &e->graph_prompt, true, (int)n_inputs, 0)) e->graph_n_inputs = (int)n_inputs;
      else e->use_graph = false;
    }
  }
  for (size_t st = first; st < n_total; ++st) {
    const bool prompt_step = st < n_inputs;
    if (e->use_graph) {
      ZG_CUDA(cudaGraphLaunch(prompt_step ? e->graph_prompt : (sample ? e->graph_sampling : e->graph_sample), s));
      ctx().launches += prompt_step ? 2 + 7 * e->layers.size() + 1 : 1 + 7 * e->layers.size() + 4;
    } else {
      enqueue_step(e, prompt_step, (int)n_inputs, prompt_step ? 0 : (sample ? 3 : 2));
    }
  }
  e->host_pos = (int)n_total;
  // history [n_total][B] -> out [B][n_total]
  ZG_CUDA(cudaMemcpyAsync(e->hist_host, e->hist, (size_t)B * n_total * sizeof(u64), cudaMemcpyDeviceToHost, s));
  ZG_CUDA(cudaStreamSynchronize(s));
  for (size_t st = 0; st < n_total; ++st)
    for (int b = 0; b < B; ++b) out_tokens[(size_t)b * n_total + st] = (size_t)e->hist_host[st * B + b];
  if (zg_tc_error()) set_error(1, "tensor-core kernel watchdog tripped", __FILE__, __LINE__);
  return zg_last_error();
}

int zg_batch_generate_greedy(zg_batch *e, const size_t *prompts, size_t n_inputs, size_t n_total, size_t *out_tokens,
                             int use_prefill) {
  if (!require_ready("zg_batch_generate_greedy")) return 1;
  return batch_generate(e, prompts, n_inputs, n_total, out_tokens, use_prefill, false);
}

// generate() with temperature sampling for every sequence: sequence b of this engine draws its step-s uniform as
// philox_uniform(seed, s, seq_base + b), so a sequence's tokens do not depend on which GPU / batch slot it runs in.
int zg_batch_generate_sample(zg_batch *e, const size_t *prompts, size_t n_inputs, size_t n_total, float temp,
                             unsigned long long seed, unsigned long long seq_base, size_t *out_tokens, int use_prefill) {
  if (!require_ready("zg_batch_generate_sample")) return 1;
  if (!(temp > 0.0f)) {
    set_error(1, "zg_batch_generate_sample: temp must be > 0", __FILE__, __LINE__);
    return 1;
  }
  struct { float temp; unsigned long long seed, seq_base; } sp = {temp, seed, seq_base};
  ZG_CUDA(cudaMemcpyAsync(e->samp, &sp, sizeof(sp), cudaMemcpyHostToDevice, ctx().stream));
  return batch_generate(e, prompts, n_inputs, n_total, out_tokens, use_prefill, true);
}

// Device-resident stepping for timing: run n_steps sampling steps from the current position (no host copies).
void zg_batch_run_steps(zg_batch *e, size_t n_steps) {
  if (!require_ready("zg_batch_run_steps")) return;
  cudaStream_t s = ctx().stream;
  if (e->host_pos + n_steps > (size_t)e->cap) {
    set_error(1, "zg_batch_run_steps: would run past the KV cache", __FILE__, __LINE__);
    return;
  }
  e->host_pos += (int)n_steps;
  if (e->use_graph && !e->graph_sample && !capture(e, &e->graph_sample, false, 0, 2)) e->use_graph = false;
  for (size_t i = 0; i < n_steps; ++i) {
    if (e->use_graph) {
      ZG_CUDA(cudaGraphLaunch(e->graph_sample, s));
      ctx().launches += 1 + 7 * e->layers.size() + 4;
    } else {
      enqueue_step(e, false, 0, 2);
    }
  }
}
int zg_batch_fused_argmax(const zg_batch *e) { (void)e; return 1; }
int zg_batch_storage_bits(const zg_batch *e) { return e->store16 ? 16 : 32; }
// Set the common position (and so the attended length) directly: timing a step at T = 1024 needs no 1023 real steps.
void zg_batch_set_position(zg_batch *e, size_t pos) {
  if (!require_ready("zg_batch_set_position")) return;
  set_pos_kernel<<<1, 1, 0, ctx().stream>>>(e->pos, (int)pos, 0);
  ZG_LAUNCH_CHECK();
  e->host_pos = (int)pos;
}
// The token every sequence produced in its last sampling step (argmax of its logits): n_seqs ids, HOST.  Synchronises.
int zg_batch_read_tokens(zg_batch *e, size_t *out_tokens) {
  if (!require_ready("zg_batch_read_tokens")) return 1;
  cudaStream_t s = ctx().stream;
  ZG_CUDA(cudaMemcpyAsync(e->hist_host, e->tok, (size_t)e->B * sizeof(u64), cudaMemcpyDeviceToHost, s));
  ZG_CUDA(cudaStreamSynchronize(s));
  for (int b = 0; b < e->B; ++b) out_tokens[b] = (size_t)e->hist_host[b];
  return zg_last_error();
}
const float *zg_batch_k_cache(const zg_batch *e, size_t layer) {  // f16 data when zg_batch_storage_bits() == 16
  if (e->store16) return reinterpret_cast<const float *>((const __half *)e->k_cache16 + layer * e->layer_stride);
  return e->k_cache + layer * e->layer_stride;
}
const float *zg_batch_v_cache(const zg_batch *e, size_t layer) {
  if (e->store16) return reinterpret_cast<const float *>((const __half *)e->v_cache16 + layer * e->layer_stride);
  return e->v_cache + layer * e->layer_stride;
}

}  // extern "C"
