/*
 * zg_model.c -- CPU ORACLE (test infrastructure only; see zg_oracle.h).
 * Plain-C restatement of /root/reference/src/main.zig (State, MLP, Block, GPT, loaders,
 * generate).  Each function cites the lines it follows.
 */
#include "zg_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- State.init, main.zig:46-64 ----------------------------------------------------- */
int zo_state_init(zo_state *s, const zo_config *c) {
  const size_t E = c->n_embed, C = c->context_size;
  memset(s, 0, sizeof(*s));
  s->pos_emb = calloc(E, sizeof(float));
  s->x = calloc(E, sizeof(float));
  s->o = calloc(E, sizeof(float));
  s->logits = calloc(c->vocab_size, sizeof(float));
  s->decoded = calloc(20, 1);
  s->_h = calloc(E, sizeof(float));
  s->_4xh = calloc(4 * E, sizeof(float));
  s->_qkv = calloc(3 * E, sizeof(float));
  s->_q = calloc(E, sizeof(float));
  s->_k = calloc(C * E, sizeof(float));
  s->_v = calloc(C * E, sizeof(float));
  s->_attn = calloc(C, sizeof(float));
  return (s->pos_emb && s->x && s->o && s->logits && s->decoded && s->_h && s->_4xh && s->_qkv &&
          s->_q && s->_k && s->_v && s->_attn)
             ? 0
             : -1;
}
void zo_state_free(zo_state *s) {
  free(s->pos_emb); free(s->x); free(s->o); free(s->logits); free(s->decoded); free(s->_h);
  free(s->_4xh); free(s->_qkv); free(s->_q); free(s->_k); free(s->_v); free(s->_attn);
  memset(s, 0, sizeof(*s));
}

/* ---- MLP.forward, main.zig:78-82: result in state.o --------------------------------- */
void zo_mlp_forward(const zo_mlp *self, const float *inputs, size_t inputs_len, const zo_state *state) {
  zo_linear_forward(&self->c_fc, inputs, inputs_len, state->_4xh);
  zo_gelu(state->_4xh, inputs_len / self->c_fc.in_features * self->c_fc.out_features);
  zo_linear_forward(&self->c_proj, state->_4xh,
                    inputs_len / self->c_fc.in_features * self->c_fc.out_features, state->o);
}

/* ---- Block.forward, main.zig:119-146 (with `state.o` read as `state.o.len`, the obvious
 * intent of :136 and :142, which do not compile as written) ---------------------------- */
void zo_block_forward(const zo_block *self, size_t seq_len, const float *inputs, const zo_state *state) {
  const size_t E = self->n_embed;
  memcpy(state->_h, inputs, E * sizeof(float)); /* :121 */
  zo_layer_norm_forward(&self->ln_1, state->_h, E); /* :123 */
  zo_attention_forward(&self->attn, seq_len, state->_h, self->k_cache, self->v_cache, state->o,
                       state->_qkv, state->_q, state->_k, state->_v, state->_attn); /* :124-135 */
  for (size_t i = 0; i < E; ++i) { /* :136-139 */
    state->_h[i] = state->o[i] + inputs[i];
    state->x[i] = state->_h[i];
  }
  zo_layer_norm_forward(&self->ln_2, state->_h, E); /* :140 */
  zo_mlp_forward(&self->mlp, state->_h, E, state);  /* :141 */
  for (size_t i = 0; i < E; ++i) {                  /* :142-145 */
    state->o[i] += state->x[i];
    state->x[i] = state->o[i];
  }
}

/* ---- GPT.forward, main.zig:178-195 -------------------------------------------------- */
void zo_gpt_forward(const zo_gpt *self, size_t seq_len, size_t token, int compute_logits, const zo_state *state) {
  const size_t pos = seq_len - 1;
  zo_embedding_forward(&self->wpe, &pos, 1, state->pos_emb); /* :179 */
  zo_embedding_forward(&self->wte, &token, 1, state->x);     /* :180 */
  for (size_t i = 0; i < self->config.n_embed; ++i) state->x[i] += state->pos_emb[i]; /* :181-183 */
  for (size_t i = 0; i < self->config.n_layer; ++i) zo_block_forward(&self->h[i], seq_len, state->x, state);
  zo_layer_norm_forward(&self->ln_f, state->x, self->config.n_embed); /* :189 */
  if (compute_logits) zo_linear_forward(&self->lm_head, state->x, self->config.n_embed, state->logits);
}

/* ---- GPT.sample, main.zig:198-207.  The reference seeds xoshiro256++ from wall-clock seconds
 * on every call and draws with std.rand.weightedIndex: point = uniform * sum(p); return the
 * first index whose running fp32 sum exceeds point.  `u` replaces the uniform draw. -------- */
size_t zo_gpt_sample(const zo_gpt *self, size_t seq_len, float temp, size_t token, const zo_state *state, double u) {
  const size_t V = self->config.vocab_size;
  zo_gpt_forward(self, seq_len, token, 1, state);
  for (size_t i = 0; i < V; ++i) state->logits[i] /= temp; /* :200-202 */
  zo_softmax(state->logits, V);                            /* :203 */
  float sum = 0.0f;
  for (size_t i = 0; i < V; ++i) sum += state->logits[i];
  const float point = (float)u * sum;
  float acc = 0.0f;
  for (size_t i = 0; i < V; ++i) {
    acc += state->logits[i];
    if (point < acc) return i;
  }
  return V - 1;
}

/* Extension: greedy decode (the north_star's parity mode).  First maximum wins. */
size_t zo_gpt_sample_greedy(const zo_gpt *self, size_t seq_len, size_t token, const zo_state *state) {
  const size_t V = self->config.vocab_size;
  zo_gpt_forward(self, seq_len, token, 1, state);
  size_t best = 0;
  for (size_t i = 1; i < V; ++i)
    if (state->logits[i] > state->logits[best]) best = i;
  return best;
}

/* ---- generate, main.zig:322-342 ----------------------------------------------------- */
void zo_generate_greedy(const zo_gpt *gpt, const size_t *inputs, size_t n_inputs, size_t n_total,
                        const zo_state *state, size_t *out_tokens, float *logits_dump) {
  size_t token = 0;
  for (size_t s = 0; s < n_total; ++s) {
    if (s < n_inputs) { /* :331-334 fill the KV cache one prompt token at a time, no logits */
      token = inputs[s];
      zo_gpt_forward(gpt, s + 1, token, 0, state);
    } else { /* :335-338; the first sampled step re-forwards the last prompt token at position s */
      token = zo_gpt_sample_greedy(gpt, s + 1, token, state);
      if (logits_dump) {
        memcpy(logits_dump, state->logits, gpt->config.vocab_size * sizeof(float));
        logits_dump += gpt->config.vocab_size;
      }
    }
    out_tokens[s] = token;
  }
}

/* ---- model assembly, main.zig:271-314 ----------------------------------------------- */
size_t zo_weight_count(const zo_config *c) { return 2 + ZO_W_PER_BLOCK * c->n_layer + 2; }

static zo_linear mk_linear(size_t in_f, size_t out_f, const float *w, const float *b) {
  zo_linear l = {in_f, out_f, w, b};
  return l;
}
static zo_layer_norm mk_ln(size_t n, const float *g, const float *b) {
  zo_layer_norm l = {n, g, b, 1e-5f}; /* ops.zig:76 */
  return l;
}

int zo_gpt_init(zo_gpt *g, const zo_config *c, const float *const *w) {
  const size_t E = c->n_embed;
  memset(g, 0, sizeof(*g));
  g->config = *c;
  g->wte.emb_dim = E; g->wte.weight = w[0];
  g->wpe.emb_dim = E; g->wpe.weight = w[1];
  g->h = calloc(c->n_layer, sizeof(zo_block));
  if (!g->h) return -1;
  for (size_t l = 0; l < c->n_layer; ++l) {
    const float *const *b = w + 2 + ZO_W_PER_BLOCK * l;
    zo_block *blk = &g->h[l];
    blk->n_embed = E;
    blk->ln_1 = mk_ln(E, b[0], b[1]);
    blk->attn.n_heads = c->n_heads;
    blk->attn.n_embed = E;
    blk->attn.head_dim = E / c->n_heads; /* ops.zig:120 */
    blk->attn.c_attn = mk_linear(E, 3 * E, b[2], b[3]);
    blk->attn.c_proj = mk_linear(E, E, b[4], b[5]);
    blk->ln_2 = mk_ln(E, b[6], b[7]);
    blk->mlp.c_fc = mk_linear(E, 4 * E, b[8], b[9]);
    blk->mlp.c_proj = mk_linear(4 * E, E, b[10], b[11]);
    blk->k_cache = calloc(c->context_size * E, sizeof(float)); /* main.zig:298-299 */
    blk->v_cache = calloc(c->context_size * E, sizeof(float));
    if (!blk->k_cache || !blk->v_cache) return -1;
  }
  const float *const *tail = w + 2 + ZO_W_PER_BLOCK * c->n_layer;
  g->ln_f = mk_ln(E, tail[0], tail[1]);
  g->lm_head = mk_linear(E, c->vocab_size, g->wte.weight, NULL); /* main.zig:312 weight tying */
  return 0;
}

void zo_gpt_free(zo_gpt *g) {
  if (g->h) {
    for (size_t l = 0; l < g->config.n_layer; ++l) {
      free(g->h[l].k_cache);
      free(g->h[l].v_cache);
    }
    free(g->h);
  }
  memset(g, 0, sizeof(*g));
}

/* File names: main.zig:216,224,240,248,260 and :272-294. */
static float *load_named(const char *dir, const char *name, size_t n) {
  char path[1024];
  snprintf(path, sizeof(path), "%s/model-%s", dir, name);
  float *p = calloc(n, sizeof(float));
  if (!p) return NULL;
  if (zo_load_tensor(path, p, n * sizeof(float)) < 0) { /* open failure is an error; short reads are not */
    free(p);
    return NULL;
  }
  return p;
}

int zo_load_gpt(zo_gpt *g, const zo_config *c, const char *raw_dir, float ***owned_out) {
  const size_t E = c->n_embed, nw = zo_weight_count(c);
  float **w = calloc(nw, sizeof(float *));
  if (!w) return -1;
  size_t k = 0;
  char name[128];
  w[k++] = load_named(raw_dir, "wte", c->vocab_size * E);
  w[k++] = load_named(raw_dir, "wpe", c->context_size * E);
  static const char *const kinds[ZO_W_PER_BLOCK] = {
      "ln_1-g", "ln_1-b", "attn-c_attn-w", "attn-c_attn-b", "attn-c_proj-w", "attn-c_proj-b",
      "ln_2-g", "ln_2-b", "mlp-c_fc-w",    "mlp-c_fc-b",    "mlp-c_proj-w",  "mlp-c_proj-b"};
  for (size_t l = 0; l < c->n_layer; ++l) {
    const size_t sizes[ZO_W_PER_BLOCK] = {E, E, 3 * E * E, 3 * E, E * E, E, E, E, 4 * E * E, 4 * E, 4 * E * E, E};
    for (int j = 0; j < ZO_W_PER_BLOCK; ++j) {
      snprintf(name, sizeof(name), "h%zu-%s", l, kinds[j]);
      w[k++] = load_named(raw_dir, name, sizes[j]);
    }
  }
  w[k++] = load_named(raw_dir, "ln_f-g", E);
  w[k++] = load_named(raw_dir, "ln_f-b", E);
  for (size_t i = 0; i < nw; ++i)
    if (!w[i]) return -2;
  *owned_out = w;
  return zo_gpt_init(g, c, (const float *const *)w);
}
