/*
 * zg_b200.h -- C-ABI of the B200-native (sm_100a) GPT-2 forward/decode path that replaces the
 * CPU arithmetic of EugenHotaj/zig_gpt2 behind its own operator surface.
 *
 * The reference has exactly two FFI imports: CBLAS (src/ops.zig:2, used at :30,:268,:289) and
 * regex.h (src/bpe.zig:2).  This header replaces the CBLAS import: the Zig host keeps the
 * reference's structs and method signatures (src/ops.zig:4-307, src/main.zig:5-342) and each
 * `forward` body becomes one call below.  Slices become (pointer, length) pairs whose pointers
 * are DEVICE pointers obtained from zg_alloc(); the host never dereferences them.  The struct
 * layouts below are the reference's struct fields in declaration order, so a Zig `extern struct`
 * with the same fields is ABI-compatible.
 *
 * Rules kept from the reference:
 *   - no allocation after start-up (README.md "No memory allocations at runtime"): only
 *     zg_init / zg_alloc / zg_engine_create / zg_batch_create allocate; every hot-path entry
 *     point takes caller-owned buffers and enqueues kernels on one stream;
 *   - hot-path calls return void (Zig `void`); failures are sticky and read with zg_last_error();
 *   - the caller is single-threaded per device.
 *
 * There is no CPU fallback: every entry point launches hand-written CUDA kernels and fails
 * (sticky error, non-zero status from init calls) when no sm_100 device is present.
 */
#ifndef ZG_B200_H
#define ZG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------------
 * Lifecycle / memory (start-up only).  Status: 0 = ok, otherwise a cudaError_t value.
 * ------------------------------------------------------------------------------------------- */
int zg_init(int device);            /* select device, create the stream and scratch; idempotent */
int zg_shutdown(void);
int zg_device_count(void);
int zg_sm_count(void);
void *zg_alloc(size_t bytes);       /* replaces allocator.alloc() for tensors (ops.zig:314, main.zig:48-60,298-299) */
int zg_free(void *dev_ptr);
int zg_memset(void *dev_ptr, int value, size_t bytes);
int zg_upload(void *dst_dev, const void *src_host, size_t bytes);   /* replaces fd.readAll into the slice (ops.zig:318) */
int zg_download(void *dst_host, const void *src_dev, size_t bytes); /* synchronises the stream first */
int zg_sync(void);
int zg_last_error(void);            /* sticky; 0 when clean */
const char *zg_last_error_string(void);
void zg_clear_error(void);
/* Use an externally owned CUDA stream (e.g. torch's current stream) for subsequent calls; NULL restores the library's. */
int zg_set_stream(void *cuda_stream);
/* Number of kernels this library has launched since zg_init (bench.py's gpu_launches). */
unsigned long long zg_launch_count(void);
/* Number of allocations (cudaMalloc, cudaHostAlloc, CUDA-graph instantiations) and tensor-map encodes since zg_init.
 * The reference's rule is "no memory allocations at runtime" (README.md): this counter must not move across
 * zg_engine_generate_greedy / zg_engine_run_steps / zg_engine_forward / zg_batch_forward / zg_batch_run_steps once the
 * engine exists (tests/test_gpu_model.py, tests/test_gpu_batch.py assert it). */
unsigned long long zg_alloc_count(void);
/* Philox4x32-10 counter-based generator behind the sampling paths: the uniform draw in [0,1) of sampling step `step` of
 * sequence `sequence` under `seed` -- a pure function, identical on host and device (csrc/zg_philox.cuh).  The
 * reference re-seeds its PRNG from the wall clock on every GPT.sample call (main.zig:204), which cannot be reproduced. */
double zg_philox_uniform(unsigned long long seed, unsigned long long step, unsigned long long sequence);
void zg_philox4x32_10(const unsigned counter[4], const unsigned key[2], unsigned out[4]); /* the raw block function (known-answer tests) */
/* CUDA-event stopwatch on the library's stream (bench.py times kernels on the stream they are launched on). */
int zg_timer_begin(void);
float zg_timer_end_ms(void); /* records the stop event, synchronises, returns elapsed milliseconds */

/* ---------------------------------------------------------------------------------------------
 * ops.zig
 * ------------------------------------------------------------------------------------------- */
typedef struct { /* ops.zig:4-19 */
  size_t in_features, out_features;
  const float *weight; /* device, [out_features, in_features] row-major */
  const float *bias;   /* device or NULL */
} zg_linear;
/* Linear.forward, ops.zig:21-46: outputs[M,N] = bias + inputs[M,K] weight[N,K]^T, M = inputs_len / in_features.
 * M < 16: fp32 SIMT path (warp-per-row GEMV, HBM-bound, fp32 accumulation).  M >= 16: tcgen05 tensor-core GEMM in
 * the error-compensated 3xTF32 mode (see zg_linear_forward_tc, precision 2). */
void zg_linear_forward(const zg_linear *self, const float *inputs, size_t inputs_len, float *outputs);

/* Linear.forward on the 5th-generation tensor cores (tcgen05.mma, TMEM accumulators, TMA-fed shared-memory ring),
 * the replacement of the reference's cblas_sgemm(RowMajor, NoTrans, Trans) call (ops.zig:30-45) when M >= 16.
 *   precision 0: fp32 operands read in place as kind::tf32 (`inputs` fp32, weight = self->weight);
 *   precision 2: as 0 with 3xTF32 error compensation (hi/lo operand split in shared memory, three MMAs per K step):
 *                fp32-class accuracy on the tensor cores, used by the HBM-bound batched decode step;
 *   precision 1: f16 operands (`inputs` and `weight_lowp` are f16 copies made with zg_to_f16), fp32 accumulation.
 *   epi: 0 = bias only, 1 = bias + GELU (main.zig:79-80 fused), 2 = bias + residual add of `resid` [M,N] (main.zig:136-145).
 *   tile_n: 0 = automatic, else 32/64/128/256 (N width of the CTA tile).
 * zg_linear_forward itself takes this path (precision 2) when M >= 16. */
void zg_linear_forward_tc(const zg_linear *self, const void *inputs, size_t inputs_len, float *outputs, int precision,
                          const void *weight_lowp, int epi, const float *resid, int tile_n);
/* Linear.forward for 1 <= M <= 128 rows (the batched decode step) on the tensor cores with the operands swapped: 128
 * weight rows are the UMMA M operand, the whole batch is N, the (weight tile, k-block) grid is split evenly over the SMs
 * (stream-K) and partial sums are reduced into `outputs` with fp32 atomics.  `outputs` must hold zeros (plain Linear) or
 * the residual (x += Linear(h), main.zig:136-145) on entry; in_features % 32 == 0.  precision 0 = TF32, 2 = 3xTF32,
 * 1 = f16 operands (`inputs` and `weight_lowp` are f16 copies made with zg_to_f16; in_features % 64 == 0; no xform).
 * xform bit 0 applies GELU (ops.zig:221-228) to `inputs` on the fly (mlp c_proj reading c_fc's pre-activation,
 * main.zig:80); bit 1 is a test hook (element-wise fp32 atomics instead of TMA reduce-adds in the epilogue). */
void zg_linear_forward_skinny(const zg_linear *self, const void *inputs, size_t inputs_len, float *outputs, int precision,
                              int xform, const void *weight_lowp);
/* Greedy sampling through a Linear without materialising its outputs (the tied lm_head + argmax, main.zig:193): for every
 * row m of `inputs`, tokens_dev[m] (DEVICE, 64-bit) = index of the first maximum of inputs[m,:] . W^T + bias.  The argmax
 * runs in the GEMM epilogue (whole weight tiles per CTA, packed atomicMax per row); `best_scratch` is 2 * M 64-bit words
 * of device scratch.  1 <= M <= 128, in_features % 32 == 0.  Asynchronous. */
void zg_linear_argmax_skinny(const zg_linear *self, const float *inputs, size_t inputs_len, int precision,
                             unsigned long long *best_scratch, size_t *tokens_dev);
void zg_to_f16(const float *src, void *dst_f16, size_t n); /* fp32 -> f16 round-to-nearest-even copy (start-up) */
void zg_tc_set_direct_epilogue(int on); /* test hook: 1 = per-row direct stores instead of the staged TMA-store epilogue */
unsigned long long zg_tc_pair_launch_count(void); /* launches of the CTA-pair (cta_group::2) GEMM so far: tests assert the path they cover */
int zg_tc_error(void); /* watchdog word of the tensor-core kernels (0 = clean); synchronises */

typedef struct { size_t emb_dim; const float *weight; } zg_embedding; /* ops.zig:49-57 */
/* Embedding.forward, ops.zig:59-67.  `idxs` is a HOST array of 64-bit indices (the reference passes
 * `&[1]usize{token}`, main.zig:179-180); `embeddings` is a device pointer. */
void zg_embedding_forward(const zg_embedding *self, const size_t *idxs, size_t n_idxs, float *embeddings);

typedef struct { size_t n_features; const float *weight, *bias; float eps; } zg_layer_norm; /* ops.zig:70-80 */
void zg_layer_norm_forward(const zg_layer_norm *self, float *inputs, size_t inputs_len); /* ops.zig:82-104, in place */

typedef struct { /* ops.zig:107-124 */
  size_t n_heads, n_embed, head_dim;
  zg_linear c_attn, c_proj;
} zg_attention;
/* CausalSelfAttention.forward, ops.zig:129-173: one new token, cache append at row seq_len-1.
 * The scratch arguments keep the reference's meaning; _k/_v/_attn are accepted and left untouched
 * (the fused kernel reads the time-major cache in place instead of transposing it every token). */
void zg_attention_forward(const zg_attention *self, size_t seq_len, const float *inputs, float *k_cache,
                          float *v_cache, float *outputs, float *_qkv, float *_q, float *_k, float *_v,
                          float *_attn);
void zg_split_qkv(const zg_attention *self, size_t seq_len, const float *inputs, size_t inputs_len,
                  size_t split_idx, float *outputs);                                      /* ops.zig:177-196 */
void zg_transpose(const size_t shape[3], const float *inputs, size_t inputs_len, float *outputs); /* ops.zig:199-216 */
void zg_gelu(float *inputs, size_t n);    /* ops.zig:221-228, in place */
void zg_softmax(float *inputs, size_t n); /* ops.zig:231-241, in place over the whole slice */
/* scaled_dot_product_attention, ops.zig:249-307: q[B,n,1,hd], k/v[B,n,T,hd] -> outputs[B,n,1,hd]. */
void zg_sdpa(const float *q, const float *k, size_t k_len, const float *v, size_t n_heads, size_t seq_len,
             size_t head_dim, float *outputs, float *_attn);

/* ---------------------------------------------------------------------------------------------
 * main.zig
 * ------------------------------------------------------------------------------------------- */
typedef struct { size_t vocab_size, context_size, n_layer, n_heads, n_embed; } zg_config; /* main.zig:5-23 */

typedef struct { /* main.zig:26-65; every float* is a device pointer, `decoded` is host memory */
  float *pos_emb, *x, *o, *logits;
  unsigned char *decoded;
  float *_h, *_4xh, *_qkv, *_q, *_k, *_v, *_attn;
} zg_state;
/* State.init, main.zig:46-64, with zg_alloc as the allocator.  _k/_v (2 x context x n_embed floats of
 * transposed-cache scratch in the reference) are allocated only if want_transpose_scratch != 0. */
int zg_state_init(zg_state *s, const zg_config *c, int want_transpose_scratch);
void zg_state_free(zg_state *s);

typedef struct { zg_linear c_fc, c_proj; } zg_mlp; /* main.zig:67-83 */
typedef struct { /* main.zig:85-117 */
  size_t n_embed;
  zg_layer_norm ln_1;
  zg_attention attn;
  zg_layer_norm ln_2;
  zg_mlp mlp;
  float *k_cache, *v_cache; /* device, [context_size, n_embed] time-major, heads interleaved (main.zig:298-299) */
} zg_block;
typedef struct { /* main.zig:149-176 */
  zg_config config;
  zg_embedding wte, wpe;
  const zg_block *h; /* host array of n_layer blocks */
  zg_layer_norm ln_f;
  zg_linear lm_head; /* weight tied to wte.weight, no bias (main.zig:312) */
} zg_gpt;

void zg_mlp_forward(const zg_mlp *self, const float *inputs, size_t inputs_len, const zg_state *state); /* main.zig:78-82 */
void zg_block_forward(const zg_block *self, size_t seq_len, const float *inputs, const zg_state *state); /* main.zig:119-146 */
/* GPT.forward, main.zig:178-195, op by op (one kernel per reference op). */
void zg_gpt_forward(const zg_gpt *self, size_t seq_len, size_t token, int compute_logits, const zg_state *state);
/* GPT.sample, main.zig:198-207: temperature softmax + inverse-CDF draw on the device; `u` in [0,1) replaces
 * the reference's wall-clock-seeded PRNG draw.  Returns the token id (synchronises). */
size_t zg_gpt_sample(const zg_gpt *self, size_t seq_len, float temp, size_t token, const zg_state *state, double u);
size_t zg_gpt_sample_greedy(const zg_gpt *self, size_t seq_len, size_t token, const zg_state *state);

/* Model assembly helpers (main.zig:210-314).  `w` = device pointers in canonical order:
 * wte, wpe, then per block {ln_1-g, ln_1-b, attn-c_attn-w, -b, attn-c_proj-w, -b, ln_2-g, ln_2-b,
 * mlp-c_fc-w, -b, mlp-c_proj-w, -b}, then ln_f-g, ln_f-b.  Allocates the block array (host) and the KV caches. */
size_t zg_weight_count(const zg_config *c);
size_t zg_weight_elems(const zg_config *c, size_t index);
int zg_gpt_init(zg_gpt *g, const zg_config *c, const float *const *w);
void zg_gpt_free(zg_gpt *g);
/* load_gpt, main.zig:304-314: reads `<raw_dir>/model-<name>` files (headerless LE fp32) straight to the device. */
int zg_load_gpt(zg_gpt *g, const zg_config *c, const char *raw_dir);

/* ---------------------------------------------------------------------------------------------
 * The fused decode engine: GPT.forward / GPT.sample / generate (main.zig:178-207,322-342) as ONE
 * persistent cooperative kernel per call -- every SM streams its share of the weights through a
 * shared-memory ring (cp.async.bulk + mbarrier) while grid-wide barriers separate the layer phases.
 * Batch 1, fp32 weights, fp32 accumulation.
 * ------------------------------------------------------------------------------------------- */
typedef struct zg_engine zg_engine;
zg_engine *zg_engine_create(const zg_gpt *gpt, const zg_state *state); /* start-up: descriptor tables, barrier words, pinned token ring */
void zg_engine_destroy(zg_engine *e);
void zg_engine_forward(zg_engine *e, size_t seq_len, size_t token, int compute_logits); /* == GPT.forward; logits land in state.logits */
size_t zg_engine_sample_greedy(zg_engine *e, size_t seq_len, size_t token);             /* forward + argmax; synchronises */
size_t zg_engine_sample(zg_engine *e, size_t seq_len, float temp, size_t token, double u);
/* generate(), main.zig:322-342, greedy: steps s in [0, n_total); prompt tokens are forwarded one at a time
 * without logits, the last prompt token is forwarded twice (as the reference does), every step's token is
 * written to out_tokens (HOST).  One kernel launch covers all steps; tokens are also streamed into a pinned
 * host ring as they are produced.  Returns 0 or an error code. */
int zg_engine_generate_greedy(zg_engine *e, const size_t *inputs, size_t n_inputs, size_t n_total, size_t *out_tokens);
/* generate() with GPT.sample (main.zig:198-207, 322-342) instead of greedy argmax, device resident: temperature softmax and
 * inverse-CDF draw run on the device with u = zg_philox_uniform(seed, step, sequence); one wait for the whole call. */
int zg_engine_generate_sample(zg_engine *e, const size_t *inputs, size_t n_inputs, size_t n_total, float temp,
                              unsigned long long seed, unsigned long long sequence, size_t *out_tokens);
/* Device-only variant for timing: runs steps [first_step, first_step + n_steps) of a generation whose prompt
 * (device-resident copy made by zg_engine_set_prompt) has n_inputs tokens.  Asynchronous. */
int zg_engine_set_prompt(zg_engine *e, const size_t *inputs, size_t n_inputs);
void zg_engine_run_steps(zg_engine *e, size_t first_step, size_t n_steps);
int zg_engine_read_tokens(zg_engine *e, size_t first_step, size_t n_steps, size_t *out_tokens);
/* Cycle timeline of the last launch, for profiling.  zg_engine_read_profile(e, NULL, n): n != 0 makes thread 0 of CTA
 * n - 1 record (tag = 512 + 16 * phase kind + point, SM clock cycles) pairs in later launches, n == 0 switches it off.
 * With a buffer: copies up to max_entries / 2 pairs into `out`, returns the number of pairs.  scripts/clock_profile.py */
size_t zg_engine_read_profile(zg_engine *e, unsigned long long *out, size_t max_entries);

/* ---------------------------------------------------------------------------------------------
 * Batched paths (BASELINE configs 3-5): B independent sequences as B rows of every Linear, so the reference's
 * M=1 sgemm calls become tensor-core GEMMs.  Sequences never interact (no cross-sequence op in ops.zig/main.zig).
 *   decode step: kind::tf32 GEMMs over the fp32 weights in place + batched single-query attention over per-sequence
 *                fp32 KV caches [cache_rows, n_embed] per block (main.zig:298-299 once per sequence), one CUDA graph per step;
 *   prefill:     whole prompts at once -- f16 copies of the weights (made here), kind::f16 GEMMs with fused
 *                bias/GELU/residual/KV-append epilogues and causal flash attention (tcgen05); replaces the
 *                token-at-a-time prompt loop of generate() (main.zig:331-334) and leaves the same caches behind.
 * ------------------------------------------------------------------------------------------- */
typedef struct zg_batch zg_batch;
/* start-up: caches for n_seqs sequences of up to cache_rows positions, activation sets, plans (tensor maps), f16
 * weight copies when max_prompt > 0 (prefill enabled for prompts up to max_prompt tokens).  flags bit 0: no CUDA graph;
 * bit 1: single-pass TF32 decode GEMMs instead of the error-compensated 3xTF32 default; bit 2: fp32-class prefill (3xTF32
 * GEMMs on fp32 activations + fp32 causal attention) instead of the f16 pipeline -- the prefilled generate() is then
 * token-identical to the reference's token-at-a-time prompt loop, at about a third of the f16 prefill's speed; bit 3: never the swapped-operand
 * stream-K GEMMs (the default decode step for n_seqs <= 128, zg_linear_forward_skinny), always the general kernel;
 * bit 4: 16-bit storage for the decode step: f16 copies of every weight (made here, once), f16 KV
 * caches, f16 operands, fp32 residual stream and accumulation -- half the bytes per step, tensor-core tolerance (<= 2e-2 on
 * logits).  Needs n_seqs <= 128 and max_prompt == 0; zg_batch_k_cache / v_cache then return f16 data. */
zg_batch *zg_batch_create(const zg_gpt *gpt, size_t n_seqs, size_t cache_rows, size_t max_prompt, int flags);
void zg_batch_destroy(zg_batch *e);
/* GPT.forward(seq_len, tokens[b], compute_logits) for every sequence b (tokens: HOST, n_seqs entries).  compute_logits:
 * 0 = none, 1 = logits (zg_batch_logits) and their argmax (zg_batch_read_tokens), 2 = argmax only -- fused into the
 * lm_head GEMM's epilogue (zg_batch_fused_argmax() == 1), so no logits are written (what generate / run_steps use). */
void zg_batch_forward(zg_batch *e, size_t seq_len, const size_t *tokens, int compute_logits);
const float *zg_batch_logits(const zg_batch *e);      /* device, [n_seqs, pitch] */
size_t zg_batch_logits_pitch(const zg_batch *e);      /* floats per row (vocab_size rounded up to 4) */
/* all prompt positions at once: tokens[b*T + t] (HOST); equals T calls of GPT.forward per sequence (cache rows [0,T),
 * logits of the last position when compute_logits). */
int zg_batch_prefill(zg_batch *e, const size_t *tokens, size_t T, int compute_logits);
int zg_batch_prefill_resident(zg_batch *e, size_t T, int compute_logits); /* tokens of the last prefill, no upload (timing) */
/* generate() (main.zig:322-342), greedy, per sequence: prompts[b*n_inputs + s] (HOST) -> out_tokens[b*n_total + s] (HOST). */
int zg_batch_generate_greedy(zg_batch *e, const size_t *prompts, size_t n_inputs, size_t n_total, size_t *out_tokens,
                             int use_prefill);
/* the same with GPT.sample: sequence b draws u = zg_philox_uniform(seed, step, seq_base + b) at every sampling step */
int zg_batch_generate_sample(zg_batch *e, const size_t *prompts, size_t n_inputs, size_t n_total, float temp,
                             unsigned long long seed, unsigned long long seq_base, size_t *out_tokens, int use_prefill);
void zg_batch_set_position(zg_batch *e, size_t pos); /* next step attends to cache rows [0, pos] (timing at a given context) */
void zg_batch_run_steps(zg_batch *e, size_t n_steps); /* n greedy steps from the current position, device resident, async */
int zg_batch_storage_bits(const zg_batch *e); /* 32 (the reference's fp32 weights and caches) or 16 */
int zg_batch_fused_argmax(const zg_batch *e); /* 1: greedy steps carry the argmax in the lm_head epilogue (both GEMM kernels) */
int zg_batch_read_tokens(zg_batch *e, size_t *out_tokens); /* argmax token of every sequence's last step (HOST, n_seqs ids); synchronises */
const float *zg_batch_k_cache(const zg_batch *e, size_t layer); /* device, [n_seqs, cache_rows, n_embed] */
const float *zg_batch_v_cache(const zg_batch *e, size_t layer);
/* the two attention kernels on their own (per-op parity tests) */
void zg_attention_prefill(const void *qkv_f16, void *out_f16, size_t B, size_t T, size_t n_heads, size_t n_embed);
void zg_attention_decode_batch(const float *q, const float *k_cache, const float *v_cache, size_t B, size_t context,
                               size_t n_heads, size_t n_embed, size_t seq_len, float *out);

#ifdef __cplusplus
}
#endif
#endif /* ZG_B200_H */
