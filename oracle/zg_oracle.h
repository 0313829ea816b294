/*
 * zg_oracle.h -- CPU ORACLE for the zig_gpt2 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference's algorithm (src/ops.zig, src/main.zig,
 * src/bpe.zig of EugenHotaj/zig_gpt2).  It exists so that the CUDA product path in
 * zig_gpt2_b200/ can be checked against the reference's arithmetic, and so that bench.py can
 * time the reference's CPU path on the GPU box's host cores (`cpu_baseline`, `--impl reference`).
 *
 * Nothing in zig_gpt2_b200/ (the product) may include, link, import or call this.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * The reference itself cannot be built in this image (no Zig toolchain; HEAD has two compile
 * errors, main.zig:136,142; ops.zig:2 imports the macOS-only Accelerate header), so there is
 * no oracle/_ref.  Parity pinning: the restatement is checked against the reference's own 8
 * unit tests (src/tests.zig:22-388, comparator src/tests.zig:4-20) on fixtures made by the
 * generate_test_data.py procedure (tests/golden/make_golden.py), and against the reference's
 * PyTorch model (generate_nano_gpt.py:24-152) executed from /root/reference at
 * fixture-generation time.
 *
 * Every function cites the reference file:line it follows.
 */
#ifndef ZG_ORACLE_H
#define ZG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- BLAS plumbing ------------------------------------------------------------------ */
/* The reference calls cblas_sgemm from Accelerate/OpenBLAS (ops.zig:30,268,289).  The oracle
 * calls through a function pointer that defaults to a scalar triple loop and can be pointed
 * at a real OpenBLAS (the scipy-bundled libscipy_openblas, symbol scipy_cblas_sgemm). */
int zo_blas_load(const char *so_path, const char *sgemm_symbol, const char *set_threads_symbol);
void zo_blas_set_threads(int n);
int zo_blas_is_openblas(void);

/* ---- ops.zig ------------------------------------------------------------------------ */
typedef struct { /* ops.zig:4-19 */
  size_t in_features, out_features;
  const float *weight; /* [out_features, in_features] row-major ("column major" of TF's [in,out]) */
  const float *bias;   /* NULL == Zig's `?[]const f32` null */
} zo_linear;
void zo_linear_forward(const zo_linear *self, const float *inputs, size_t inputs_len, float *outputs);

typedef struct { size_t emb_dim; const float *weight; } zo_embedding; /* ops.zig:49-57 */
void zo_embedding_forward(const zo_embedding *self, const size_t *idxs, size_t n_idxs, float *embeddings);

typedef struct { size_t n_features; const float *weight, *bias; float eps; } zo_layer_norm; /* ops.zig:70-80 */
void zo_layer_norm_forward(const zo_layer_norm *self, float *inputs, size_t inputs_len);

typedef struct { /* ops.zig:107-124 */
  size_t n_heads, n_embed, head_dim;
  zo_linear c_attn, c_proj;
} zo_attention;
void zo_attention_forward(const zo_attention *self, size_t seq_len, const float *inputs,
                          float *k_cache, float *v_cache, float *outputs, float *_qkv, float *_q,
                          float *_k, float *_v, float *_attn);
void zo_split_qkv(const zo_attention *self, size_t seq_len, const float *inputs, size_t inputs_len,
                  size_t split_idx, float *outputs);
void zo_transpose(const size_t shape[3], const float *inputs, size_t inputs_len, float *outputs);
void zo_gelu(float *inputs, size_t n);
void zo_softmax(float *inputs, size_t n);
void zo_sdpa(const float *q, const float *k, size_t k_len, const float *v, size_t n_heads,
             size_t seq_len, size_t head_dim, float *outputs, float *_attn);
/* ops.zig:309-320; returns number of elements actually read (short reads are accepted). */
long zo_load_tensor(const char *path, void *dst, size_t n_bytes);

/* ---- main.zig ----------------------------------------------------------------------- */
typedef struct { size_t vocab_size, context_size, n_layer, n_heads, n_embed; } zo_config; /* main.zig:5-23 */

typedef struct { /* main.zig:26-65 */
  float *pos_emb, *x, *o, *logits;
  unsigned char *decoded;
  float *_h, *_4xh, *_qkv, *_q, *_k, *_v, *_attn;
} zo_state;

typedef struct { zo_linear c_fc, c_proj; } zo_mlp; /* main.zig:67-83 */
typedef struct { /* main.zig:85-117 */
  size_t n_embed;
  zo_layer_norm ln_1;
  zo_attention attn;
  zo_layer_norm ln_2;
  zo_mlp mlp;
  float *k_cache, *v_cache;
} zo_block;
typedef struct { /* main.zig:149-176 */
  zo_config config;
  zo_embedding wte, wpe;
  zo_block *h;
  zo_layer_norm ln_f;
  zo_linear lm_head;
} zo_gpt;

int zo_state_init(zo_state *s, const zo_config *c);
void zo_state_free(zo_state *s);
void zo_mlp_forward(const zo_mlp *self, const float *inputs, size_t inputs_len, const zo_state *state);
void zo_block_forward(const zo_block *self, size_t seq_len, const float *inputs, const zo_state *state);
void zo_gpt_forward(const zo_gpt *self, size_t seq_len, size_t token, int compute_logits, const zo_state *state);
/* main.zig:198-207 with the RNG made explicit: `u` in [0,1) is the uniform draw that
 * weightedIndex would take from its (wall-clock seeded) PRNG. */
size_t zo_gpt_sample(const zo_gpt *self, size_t seq_len, float temp, size_t token, const zo_state *state, double u);
/* Extension (not in the reference): greedy argmax, first maximum wins. */
size_t zo_gpt_sample_greedy(const zo_gpt *self, size_t seq_len, size_t token, const zo_state *state);

/* Build a GPT over caller-owned weight memory; allocates the block array + KV caches
 * (main.zig:271-314).  `w` holds pointers in zo_weight_index() order. */
enum { ZO_W_PER_BLOCK = 12 };
size_t zo_weight_count(const zo_config *c); /* 2 + 12*n_layer + 2 */
int zo_gpt_init(zo_gpt *g, const zo_config *c, const float *const *w);
void zo_gpt_free(zo_gpt *g);
/* main.zig:210-314: load every tensor from `<dir>/model-*` files in the reference's raw format. */
int zo_load_gpt(zo_gpt *g, const zo_config *c, const char *raw_dir, float ***owned_out);

/* main.zig:322-342 generate(), returning the token stream instead of printing it.
 * out_tokens receives one token per loop iteration s in [0, n_total): the prompt tokens as
 * forwarded, then sampled tokens (greedy).  n_total <= context_size.  Reproduces the
 * duplicate-last-prompt-token behaviour.  If logits_dump != NULL, the logits of every sampled
 * step are appended ([n_total - n_prompt, vocab]). */
void zo_generate_greedy(const zo_gpt *gpt, const size_t *inputs, size_t n_inputs, size_t n_total,
                        const zo_state *state, size_t *out_tokens, float *logits_dump);

/* ---- bpe.zig ------------------------------------------------------------------------ */
typedef struct zo_encoder zo_encoder;
/* tokens[i] is the (unicode-mapped, UTF-8) string of vocabulary id ids[i]; uni[j] is the UTF-8
 * string that maps to byte uni_byte[j] (the byte_encoder.json of download_weights.py:69-90). */
zo_encoder *zo_encoder_init(const char *const *tokens, const size_t *token_lens, const size_t *ids,
                            size_t n_tokens, const char *const *uni, const size_t *uni_lens,
                            const unsigned char *uni_byte, size_t n_uni);
void zo_encoder_deinit(zo_encoder *e);
/* bpe.zig:59-97.  `inputs` must be NUL terminated at inputs[len] (regexec takes a bare pointer).
 * Words longer than the reference's 20-byte buffer make the reference overflow (UB); the oracle
 * returns (size_t)-1 for them instead. */
size_t zo_encoder_encode(const zo_encoder *e, const char *inputs, size_t len, size_t *outputs, size_t max_out);
/* bpe.zig:99-118 */
size_t zo_encoder_decode(const zo_encoder *e, const size_t *inputs, size_t n, unsigned char *outputs, size_t max_out);

#ifdef __cplusplus
}
#endif
#endif
