// zg_model.cu -- src/main.zig's State / MLP / Block / GPT composed op by op from the kernels of zg_ops.cu
// (one kernel per reference op, with the GELU and residual adds folded into the Linear epilogues), plus
// model assembly and the raw-file loader.  The fused persistent path lives in zg_decode.cu.
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "zg_common.cuh"

namespace zg {
enum { EPI_NONE = 0, EPI_GELU = 1, EPI_RESIDUAL = 2 };
void launch_linear(const float *x, const float *w, const float *bias, float *out, size_t M, size_t K, size_t N,
                   int epi, const float *resid);
void launch_embed_add(const float *wte, const float *wpe, size_t token, size_t pos, int E, float *x, float *pos_emb);
void launch_layernorm(const float *in, float *out, const float *g, const float *b, size_t rows, int E, float eps);
void launch_residual1(const float *o, const float *inputs, float *h, float *x, int E);
void launch_residual2(float *o, float *x, int E);
void launch_argmax(const float *logits, size_t n, unsigned long long *out);
void launch_softmax_temp(float *x, size_t n, float temp);
void launch_weighted_index(const float *p, size_t n, float u, unsigned long long *out);
}  // namespace zg

using namespace zg;

extern "C" {

// ---- State.init, main.zig:46-64 -------------------------------------------------------------------
int zg_state_init(zg_state *s, const zg_config *c, int want_transpose_scratch) {
  memset(s, 0, sizeof(*s));
  if (!require_ready("zg_state_init")) return 1;
  const size_t E = c->n_embed, C = c->context_size;
  s->pos_emb = (float *)zg_alloc(E * sizeof(float));
  s->x = (float *)zg_alloc(E * sizeof(float));
  s->o = (float *)zg_alloc(E * sizeof(float));
  s->logits = (float *)zg_alloc(c->vocab_size * sizeof(float));
  s->decoded = (unsigned char *)calloc(20, 1);  // main.zig:52, host bytes for the tokenizer
  s->_h = (float *)zg_alloc(E * sizeof(float));
  s->_4xh = (float *)zg_alloc(4 * E * sizeof(float));
  s->_qkv = (float *)zg_alloc(3 * E * sizeof(float));
  s->_q = (float *)zg_alloc(E * sizeof(float));
  s->_attn = (float *)zg_alloc(C * sizeof(float));
  if (want_transpose_scratch) {  // the reference's [n,T,hd] copies of the whole cache; unused by the CUDA path
    s->_k = (float *)zg_alloc(C * E * sizeof(float));
    s->_v = (float *)zg_alloc(C * E * sizeof(float));
  }
  return zg_last_error();
}

void zg_state_free(zg_state *s) {
  float *ptrs[] = {s->pos_emb, s->x, s->o, s->logits, s->_h, s->_4xh, s->_qkv, s->_q, s->_k, s->_v, s->_attn};
  for (float *p : ptrs)
    if (p) zg_free(p);
  free(s->decoded);
  memset(s, 0, sizeof(*s));
}

// ---- MLP.forward, main.zig:78-82 (GELU folded into c_fc's epilogue) -------------------------------
void zg_mlp_forward(const zg_mlp *self, const float *inputs, size_t inputs_len, const zg_state *state) {
  if (!require_ready("zg_mlp_forward")) return;
  const size_t M = inputs_len / self->c_fc.in_features;
  launch_linear(inputs, self->c_fc.weight, self->c_fc.bias, state->_4xh, M, self->c_fc.in_features,
                self->c_fc.out_features, EPI_GELU, nullptr);
  launch_linear(state->_4xh, self->c_proj.weight, self->c_proj.bias, state->o, M, self->c_proj.in_features,
                self->c_proj.out_features, EPI_NONE, nullptr);
}

// ---- Block.forward, main.zig:119-146 --------------------------------------------------------------
void zg_block_forward(const zg_block *self, size_t seq_len, const float *inputs, const zg_state *state) {
  if (!require_ready("zg_block_forward")) return;
  const int E = (int)self->n_embed;
  launch_layernorm(inputs, state->_h, self->ln_1.weight, self->ln_1.bias, 1, E, self->ln_1.eps);  // :121-123
  zg_attention_forward(&self->attn, seq_len, state->_h, self->k_cache, self->v_cache, state->o, state->_qkv,
                       state->_q, state->_k, state->_v, state->_attn);                            // :124-135
  launch_residual1(state->o, inputs, state->_h, state->x, E);                                      // :136-139
  launch_layernorm(state->_h, state->_h, self->ln_2.weight, self->ln_2.bias, 1, E, self->ln_2.eps); // :140
  zg_mlp_forward(&self->mlp, state->_h, (size_t)E, state);                                         // :141
  launch_residual2(state->o, state->x, E);                                                         // :142-145
}

// ---- GPT.forward, main.zig:178-195 ----------------------------------------------------------------
void zg_gpt_forward(const zg_gpt *self, size_t seq_len, size_t token, int compute_logits, const zg_state *state) {
  if (!require_ready("zg_gpt_forward")) return;
  // The reference indexes wte / wpe / the caches unchecked (a Zig safety panic in debug builds); here a bad position
  // or token would silently corrupt device memory, so both are rejected like the fused engine rejects them.
  if (seq_len == 0 || seq_len > self->config.context_size || token >= self->config.vocab_size) {
    set_error(1, "zg_gpt_forward: need 1 <= seq_len <= context_size and token < vocab_size", __FILE__, __LINE__);
    return;
  }
  const int E = (int)self->config.n_embed;
  launch_embed_add(self->wte.weight, self->wpe.weight, token, seq_len - 1, E, state->x, state->pos_emb);
  for (size_t i = 0; i < self->config.n_layer; ++i) zg_block_forward(&self->h[i], seq_len, state->x, state);
  launch_layernorm(state->x, state->x, self->ln_f.weight, self->ln_f.bias, 1, E, self->ln_f.eps);
  if (compute_logits)
    launch_linear(state->x, self->lm_head.weight, self->lm_head.bias, state->logits, 1, E,
                  self->lm_head.out_features, EPI_NONE, nullptr);
}

static size_t read_token_slot() {
  Context &c = ctx();
  ZG_CUDA(cudaMemcpyAsync(c.token_slot_host, c.token_slot, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                          c.stream));
  ZG_CUDA(cudaStreamSynchronize(c.stream));
  return (size_t)c.token_slot_host[0];
}

size_t zg_gpt_sample_greedy(const zg_gpt *self, size_t seq_len, size_t token, const zg_state *state) {
  if (!require_ready("zg_gpt_sample_greedy")) return (size_t)-1;
  zg_gpt_forward(self, seq_len, token, 1, state);
  if (zg_last_error()) return (size_t)-1;
  launch_argmax(state->logits, self->config.vocab_size, ctx().token_slot);
  return read_token_slot();
}

// ---- GPT.sample, main.zig:198-207 -----------------------------------------------------------------
size_t zg_gpt_sample(const zg_gpt *self, size_t seq_len, float temp, size_t token, const zg_state *state, double u) {
  if (!require_ready("zg_gpt_sample")) return (size_t)-1;
  zg_gpt_forward(self, seq_len, token, 1, state);
  if (zg_last_error()) return (size_t)-1;
  launch_softmax_temp(state->logits, self->config.vocab_size, temp);  // :200-203
  launch_weighted_index(state->logits, self->config.vocab_size, (float)u, ctx().token_slot);
  return read_token_slot();
}

// ---- model assembly, main.zig:271-314 -------------------------------------------------------------
size_t zg_weight_count(const zg_config *c) { return 2 + 12 * c->n_layer + 2; }

size_t zg_weight_elems(const zg_config *c, size_t index) {
  const size_t E = c->n_embed;
  if (index == 0) return c->vocab_size * E;
  if (index == 1) return c->context_size * E;
  const size_t nb = 12 * c->n_layer;
  if (index >= 2 + nb) return E;
  const size_t per[12] = {E, E, 3 * E * E, 3 * E, E * E, E, E, E, 4 * E * E, 4 * E, 4 * E * E, E};
  return per[(index - 2) % 12];
}

static zg_linear mk_linear(size_t in_f, size_t out_f, const float *w, const float *b) {
  zg_linear l = {in_f, out_f, w, b};
  return l;
}
static zg_layer_norm mk_ln(size_t n, const float *g, const float *b) {
  zg_layer_norm l = {n, g, b, 1e-5f};  // ops.zig:76
  return l;
}

int zg_gpt_init(zg_gpt *g, const zg_config *c, const float *const *w) {
  memset(g, 0, sizeof(*g));
  if (!require_ready("zg_gpt_init")) return 1;
  const size_t E = c->n_embed;
  g->config = *c;
  g->wte.emb_dim = E;
  g->wte.weight = w[0];
  g->wpe.emb_dim = E;
  g->wpe.weight = w[1];
  zg_block *h = (zg_block *)calloc(c->n_layer, sizeof(zg_block));
  if (!h) return 2;
  const size_t cache_bytes = c->context_size * E * sizeof(float);
  for (size_t l = 0; l < c->n_layer; ++l) {
    const float *const *b = w + 2 + 12 * l;
    zg_block *blk = &h[l];
    blk->n_embed = E;
    blk->ln_1 = mk_ln(E, b[0], b[1]);
    blk->attn.n_heads = c->n_heads;
    blk->attn.n_embed = E;
    blk->attn.head_dim = E / c->n_heads;  // ops.zig:120
    blk->attn.c_attn = mk_linear(E, 3 * E, b[2], b[3]);
    blk->attn.c_proj = mk_linear(E, E, b[4], b[5]);
    blk->ln_2 = mk_ln(E, b[6], b[7]);
    blk->mlp.c_fc = mk_linear(E, 4 * E, b[8], b[9]);
    blk->mlp.c_proj = mk_linear(4 * E, E, b[10], b[11]);
    blk->k_cache = (float *)zg_alloc(cache_bytes);  // main.zig:298-299
    blk->v_cache = (float *)zg_alloc(cache_bytes);
    if (!blk->k_cache || !blk->v_cache) return zg_last_error();
    zg_memset(blk->k_cache, 0, cache_bytes);
    zg_memset(blk->v_cache, 0, cache_bytes);
  }
  g->h = h;
  const float *const *tail = w + 2 + 12 * c->n_layer;
  g->ln_f = mk_ln(E, tail[0], tail[1]);
  g->lm_head = mk_linear(E, c->vocab_size, g->wte.weight, nullptr);  // main.zig:312, weight tying
  return zg_sync();
}

void zg_gpt_free(zg_gpt *g) {
  if (g->h) {
    for (size_t l = 0; l < g->config.n_layer; ++l) {
      zg_free(g->h[l].k_cache);
      zg_free(g->h[l].v_cache);
    }
    free((void *)g->h);
  }
  memset(g, 0, sizeof(*g));
}

// ---- load_gpt, main.zig:304-314 with load_tensor (ops.zig:309-320) ----
static const char *const kBlockKinds[12] = {"ln_1-g",        "ln_1-b",     "attn-c_attn-w", "attn-c_attn-b",
                                            "attn-c_proj-w", "attn-c_proj-b", "ln_2-g",     "ln_2-b",
                                            "mlp-c_fc-w",    "mlp-c_fc-b", "mlp-c_proj-w",  "mlp-c_proj-b"};

// Files are mapped (mmap) and streamed through TWO pinned staging buffers: while the H2D copy of chunk i runs on the
// stream (cudaMemcpyAsync from pinned memory, a real DMA), the host copies chunk i + 1 out of the page cache into the
// other buffer -- disk/page-cache reads, the staging memcpy and the PCIe transfer overlap, and a 6 GB checkpoint needs
// 2 x 64 MB of pinned memory instead of a buffer as large as the largest tensor.  A file shorter than its tensor is an
// error (the reference reads with readAll and ignores the count, ops.zig:318: a short file silently leaves garbage).
int zg_load_gpt(zg_gpt *g, const zg_config *c, const char *raw_dir) {
  if (!require_ready("zg_load_gpt")) return 1;
  const size_t nw = zg_weight_count(c);
  float **w = (float **)calloc(nw, sizeof(float *));
  if (!w) return 2;
  size_t chunk = (size_t)64 << 20;
  if (const char *kb = getenv("ZG_LOAD_CHUNK_KB")) {  // tests: exercise the multi-chunk path on small tensors
    const long v = atol(kb);
    if (v > 0) chunk = (size_t)v << 10;
  }
  char *staging[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaHostAlloc((void **)&staging[i], chunk, cudaHostAllocDefault);
    note_alloc();
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    set_error((int)e, "zg_load_gpt staging", __FILE__, __LINE__);
    for (int i = 0; i < 2; ++i) {
      if (staging[i]) cudaFreeHost(staging[i]);
      if (done[i]) cudaEventDestroy(done[i]);
    }
    free(w);
    return (int)e;
  }
  cudaStream_t s = ctx().stream;
  int rc = 0, buf = 0;
  bool used[2] = {false, false};
  for (size_t i = 0; i < nw && rc == 0; ++i) {
    char name[64], path[1024];
    if (i == 0) snprintf(name, sizeof(name), "wte");
    else if (i == 1) snprintf(name, sizeof(name), "wpe");
    else if (i >= 2 + 12 * c->n_layer) snprintf(name, sizeof(name), (i - 2 - 12 * c->n_layer) == 0 ? "ln_f-g" : "ln_f-b");
    else snprintf(name, sizeof(name), "h%zu-%s", (i - 2) / 12, kBlockKinds[(i - 2) % 12]);
    snprintf(path, sizeof(path), "%s/model-%s", raw_dir, name);
    const size_t bytes = zg_weight_elems(c, i) * sizeof(float);
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { rc = 3; set_error(1, "zg_load_gpt: cannot open tensor file", path, 0); break; }
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < bytes) {
      close(fd);
      rc = 4;
      set_error(1, "zg_load_gpt: short read (the reference accepts it silently, ops.zig:318; we do not)", path, 0);
      break;
    }
    const char *map = (const char *)mmap(nullptr, bytes, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) { rc = 5; set_error(1, "zg_load_gpt: mmap failed", path, 0); break; }
    madvise((void *)map, bytes, MADV_SEQUENTIAL);
    w[i] = (float *)zg_alloc(bytes);
    if (!w[i]) { rc = zg_last_error(); munmap((void *)map, bytes); break; }
    for (size_t off = 0; off < bytes && rc == 0; off += chunk) {
      const size_t n = bytes - off < chunk ? bytes - off : chunk;
      if (used[buf]) ZG_CUDA(cudaEventSynchronize(done[buf]));  // the copy that last used this buffer has drained
      memcpy(staging[buf], map + off, n);
      ZG_CUDA(cudaMemcpyAsync((char *)w[i] + off, staging[buf], n, cudaMemcpyHostToDevice, s));
      ZG_CUDA(cudaEventRecord(done[buf], s));
      used[buf] = true;
      buf ^= 1;
      rc = zg_last_error();
    }
    munmap((void *)map, bytes);
  }
  ZG_CUDA(cudaStreamSynchronize(s));
  for (int i = 0; i < 2; ++i) {
    cudaFreeHost(staging[i]);
    cudaEventDestroy(done[i]);
  }
  if (rc == 0) rc = zg_gpt_init(g, c, (const float *const *)w);
  free(w);  // the device tensors stay alive for the life of the process, like the reference's arena (main.zig:349-350)
  return rc;
}

}  // extern "C"
