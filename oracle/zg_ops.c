/*
 * zg_ops.c -- CPU ORACLE (test infrastructure only; see zg_oracle.h).
 * Plain-C restatement of /root/reference/src/ops.zig.  Each function cites the lines it follows.
 */
#include "zg_oracle.h"

#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- cblas_sgemm stand-in ----------------------------------------------------------- */
enum { ZO_ROW_MAJOR = 101, ZO_NO_TRANS = 111, ZO_TRANS = 112 };

typedef void (*zo_sgemm_fn)(int order, int transa, int transb, int m, int n, int k, float alpha,
                            const float *a, int lda, const float *b, int ldb, float beta, float *c,
                            int ldc);
typedef void (*zo_set_threads_fn)(int);

/* Scalar sgemm for the two call shapes the reference uses (row-major, A not transposed,
 * B transposed or not).  Sequential fp32 accumulation over k. */
static void zo_sgemm_scalar(int order, int transa, int transb, int m, int n, int k, float alpha,
                            const float *a, int lda, const float *b, int ldb, float beta, float *c,
                            int ldc) {
  (void)order;
  (void)transa;
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j < n; ++j) {
      float acc = 0.0f;
      if (transb == ZO_TRANS) {
        const float *ar = a + (size_t)i * lda, *br = b + (size_t)j * ldb;
        for (int p = 0; p < k; ++p) acc += ar[p] * br[p];
      } else {
        const float *ar = a + (size_t)i * lda;
        for (int p = 0; p < k; ++p) acc += ar[p] * b[(size_t)p * ldb + j];
      }
      float *dst = c + (size_t)i * ldc + j;
      *dst = (beta == 0.0f) ? alpha * acc : alpha * acc + beta * *dst;
    }
  }
}

static zo_sgemm_fn g_sgemm = zo_sgemm_scalar;
static zo_set_threads_fn g_set_threads = NULL;
static void *g_blas_handle = NULL;

int zo_blas_load(const char *so_path, const char *sgemm_symbol, const char *set_threads_symbol) {
  if (so_path == NULL) { /* back to the scalar loop */
    g_sgemm = zo_sgemm_scalar;
    g_set_threads = NULL;
    return 0;
  }
  void *h = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return -1;
  void *f = dlsym(h, sgemm_symbol);
  if (!f) return -2;
  g_blas_handle = h;
  g_sgemm = (zo_sgemm_fn)f;
  g_set_threads = set_threads_symbol ? (zo_set_threads_fn)dlsym(h, set_threads_symbol) : NULL;
  return 0;
}
void zo_blas_set_threads(int n) {
  if (g_set_threads) g_set_threads(n);
}
int zo_blas_is_openblas(void) { return g_sgemm != zo_sgemm_scalar; }

/* ---- Linear.forward, ops.zig:21-46 -------------------------------------------------- */
void zo_linear_forward(const zo_linear *self, const float *inputs, size_t inputs_len, float *outputs) {
  const size_t batch_size = inputs_len / self->in_features; /* :22 */
  float beta = 0.0f;
  if (self->bias) { /* :24-29 bias rows are copied into the output, then sgemm accumulates */
    for (size_t b = 0; b < batch_size; ++b)
      memcpy(outputs + b * self->out_features, self->bias, self->out_features * sizeof(float));
    beta = 1.0f;
  }
  g_sgemm(ZO_ROW_MAJOR, ZO_NO_TRANS, ZO_TRANS, (int)batch_size, (int)self->out_features,
          (int)self->in_features, 1.0f, inputs, (int)self->in_features, self->weight,
          (int)self->in_features, beta, outputs, (int)self->out_features); /* :30-45 */
}

/* ---- Embedding.forward, ops.zig:59-67 ----------------------------------------------- */
void zo_embedding_forward(const zo_embedding *self, const size_t *idxs, size_t n_idxs, float *embeddings) {
  for (size_t i = 0; i < n_idxs; ++i)
    memcpy(embeddings + i * self->emb_dim, self->weight + self->emb_dim * idxs[i],
           self->emb_dim * sizeof(float));
}

/* ---- LayerNorm.forward, ops.zig:82-104 ---------------------------------------------- */
void zo_layer_norm_forward(const zo_layer_norm *self, float *inputs, size_t inputs_len) {
  const size_t nf = self->n_features;
  const size_t batch_size = inputs_len / nf;
  for (size_t b = 0; b < batch_size; ++b) {
    float mean = 0.0f, std_ = 0.0f; /* :86-92 single pass, sequential fp32 sums */
    for (size_t i = 0; i < nf; ++i) {
      const float x = inputs[b * nf + i];
      mean += x;
      std_ += x * x;
    }
    const float n = (float)nf;
    mean /= n;
    std_ = sqrtf((std_ / n) - (mean * mean) + self->eps); /* :95 */
    for (size_t i = 0; i < nf; ++i) {                      /* :98-102 */
      const size_t idx = b * nf + i;
      const float x = inputs[idx];
      inputs[idx] = (x - mean) / std_ * self->weight[i] + self->bias[i];
    }
  }
}

/* ---- CausalSelfAttention.split_qkv, ops.zig:177-196 --------------------------------- */
void zo_split_qkv(const zo_attention *self, size_t seq_len, const float *inputs, size_t inputs_len,
                  size_t split_idx, float *outputs) {
  const size_t ne = self->n_embed, ne3 = 3 * ne;
  const size_t batch_size = inputs_len / (seq_len * ne3);
  for (size_t b = 0; b < batch_size; ++b)
    for (size_t r = 0; r < seq_len; ++r) {
      const size_t out_offset = (b * seq_len * ne) + (r * ne);
      const size_t in_offset = (b * seq_len * ne3) + (r * ne3) + (split_idx * ne);
      memcpy(outputs + out_offset, inputs + in_offset, ne * sizeof(float));
    }
}

/* ---- CausalSelfAttention.transpose, ops.zig:199-216: (b,t,n,h) -> (b,n,t,h) --------- */
void zo_transpose(const size_t shape[3], const float *inputs, size_t inputs_len, float *outputs) {
  const size_t seq_len = shape[0], n_heads = shape[1], head_dim = shape[2];
  const size_t per = seq_len * n_heads * head_dim;
  const size_t batch_size = inputs_len / per;
  for (size_t b = 0; b < batch_size; ++b)
    for (size_t h = 0; h < n_heads; ++h)
      for (size_t s = 0; s < seq_len; ++s) {
        const size_t in_offset = (b * per) + (s * n_heads * head_dim) + (h * head_dim);
        const size_t out_offset = (b * per) + (h * seq_len * head_dim) + (s * head_dim);
        memcpy(outputs + out_offset, inputs + in_offset, head_dim * sizeof(float));
      }
}

/* ---- gelu, ops.zig:221-228 ---------------------------------------------------------- */
void zo_gelu(float *inputs, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    const float x = inputs[i];
    inputs[i] = 0.5f * x * (1.0f + tanhf(x * 0.7978845608f * (1.0f + 0.044715f * x * x)));
  }
}

/* ---- softmax over the whole slice, ops.zig:231-241 ---------------------------------- */
void zo_softmax(float *inputs, size_t n) {
  float max = inputs[0];
  for (size_t i = 1; i < n; ++i)
    if (inputs[i] > max) max = inputs[i];
  float sum = 0.0f;
  for (size_t i = 0; i < n; ++i) {
    inputs[i] = expf(inputs[i] - max);
    sum += inputs[i];
  }
  for (size_t i = 0; i < n; ++i) inputs[i] /= sum;
}

/* ---- scaled_dot_product_attention, ops.zig:249-307 (query length 1, no mask) -------- */
void zo_sdpa(const float *q, const float *k, size_t k_len, const float *v, size_t n_heads,
             size_t seq_len, size_t head_dim, float *outputs, float *_attn) {
  const size_t batch_size = k_len / (n_heads * seq_len * head_dim); /* :259 */
  for (size_t b = 0; b < batch_size; ++b)
    for (size_t h = 0; h < n_heads; ++h) {
      const size_t qo_offset = (b * n_heads * head_dim) + (h * head_dim);
      const size_t kv_offset = (b * n_heads * seq_len * head_dim) + (h * seq_len * head_dim);
      /* :268-283 attn[1,T] = (1/sqrt(hd)) q[1,hd] K_h[T,hd]^T */
      g_sgemm(ZO_ROW_MAJOR, ZO_NO_TRANS, ZO_TRANS, 1, (int)seq_len, (int)head_dim,
              1.0f / sqrtf((float)head_dim), q + qo_offset, (int)head_dim, k + kv_offset,
              (int)head_dim, 0.0f, _attn, (int)seq_len);
      zo_softmax(_attn, seq_len); /* :284; the caller slices _attn to [0,T) (main.zig:134) */
      /* :289-304 out[1,hd] = attn[1,T] V_h[T,hd] */
      g_sgemm(ZO_ROW_MAJOR, ZO_NO_TRANS, ZO_NO_TRANS, 1, (int)head_dim, (int)seq_len, 1.0f, _attn,
              (int)seq_len, v + kv_offset, (int)head_dim, 0.0f, outputs + qo_offset, (int)head_dim);
    }
}

/* ---- CausalSelfAttention.forward, ops.zig:129-173 ----------------------------------- */
void zo_attention_forward(const zo_attention *self, size_t seq_len, const float *inputs,
                          float *k_cache, float *v_cache, float *outputs, float *_qkv, float *_q,
                          float *_k, float *_v, float *_attn) {
  const size_t ne = self->n_embed;
  zo_linear_forward(&self->c_attn, inputs, ne, _qkv); /* :143 */

  /* Q (:146-147): split into `outputs` (used as scratch), "transpose" (identity for t=1). */
  zo_split_qkv(self, 1, _qkv, 3 * ne, 0, outputs);
  const size_t q_shape[3] = {1, self->n_heads, self->head_dim};
  zo_transpose(q_shape, outputs, ne, _q);

  const size_t t_shape[3] = {seq_len, self->n_heads, self->head_dim};
  /* K (:151-153): append row seq_len-1 to the cache, then transpose the WHOLE cache. */
  zo_split_qkv(self, 1, _qkv, 3 * ne, 1, outputs);
  memcpy(k_cache + (seq_len - 1) * ne, outputs, ne * sizeof(float));
  zo_transpose(t_shape, k_cache, seq_len * ne, _k);
  /* V (:156-158) */
  zo_split_qkv(self, 1, _qkv, 3 * ne, 2, outputs);
  memcpy(v_cache + (seq_len - 1) * ne, outputs, ne * sizeof(float));
  zo_transpose(t_shape, v_cache, seq_len * ne, _v);

  zo_sdpa(_q, _k, seq_len * ne, _v, self->n_heads, seq_len, self->head_dim, outputs, _attn); /* :160-169 */
  const size_t o_shape[3] = {self->n_heads, 1, self->head_dim};
  zo_transpose(o_shape, outputs, ne, _q);               /* :171 "hack" un-transpose */
  zo_linear_forward(&self->c_proj, _q, ne, outputs);    /* :172 */
}

/* ---- load_tensor, ops.zig:309-320 (headerless little-endian raw; short reads accepted) */
long zo_load_tensor(const char *path, void *dst, size_t n_bytes) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  size_t got = fread(dst, 1, n_bytes, f);
  fclose(f);
  return (long)got;
}
