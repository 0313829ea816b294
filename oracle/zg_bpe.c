/*
 * zg_bpe.c -- CPU ORACLE (test infrastructure only; see zg_oracle.h).
 * Plain-C restatement of /root/reference/src/bpe.zig: POSIX-ERE word split, byte->unicode
 * mapping, greedy longest-prefix vocabulary match (this is NOT merge-rank BPE; vocab.bpe is
 * never read by the reference), and the inverse decode.
 */
#include "zg_oracle.h"

#include <locale.h>
#include <regex.h>
#include <stdlib.h>
#include <string.h>

typedef struct { const char *key; size_t len; size_t val; int used; } zo_slot;
typedef struct { zo_slot *slots; size_t cap; } zo_map;

static uint64_t zo_hash(const char *s, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)s[i]; h *= 1099511628211ull; }
  return h;
}
static int zo_map_init(zo_map *m, size_t n) {
  m->cap = 16;
  while (m->cap < 2 * n + 2) m->cap <<= 1;
  m->slots = calloc(m->cap, sizeof(zo_slot));
  return m->slots ? 0 : -1;
}
static void zo_map_put(zo_map *m, const char *k, size_t len, size_t val) {
  size_t i = zo_hash(k, len) & (m->cap - 1);
  while (m->slots[i].used) {
    if (m->slots[i].len == len && memcmp(m->slots[i].key, k, len) == 0) break;
    i = (i + 1) & (m->cap - 1);
  }
  m->slots[i].key = k; m->slots[i].len = len; m->slots[i].val = val; m->slots[i].used = 1;
}
static const zo_slot *zo_map_get(const zo_map *m, const char *k, size_t len) {
  size_t i = zo_hash(k, len) & (m->cap - 1);
  while (m->slots[i].used) {
    if (m->slots[i].len == len && memcmp(m->slots[i].key, k, len) == 0) return &m->slots[i];
    i = (i + 1) & (m->cap - 1);
  }
  return NULL;
}

struct zo_encoder { /* bpe.zig:8-12 */
  zo_map token_to_idx;
  char **idx_to_token; size_t *idx_to_token_len; size_t n_idx;
  zo_map unicode_to_byte;
  char *byte_to_unicode[256]; size_t byte_to_unicode_len[256];
  char *arena; /* owns every key string */
  regex_t regex;
  locale_t c_locale; /* the Zig program never calls setlocale, so the reference's regex runs in the "C" locale */
};

/* bpe.zig:14-49 */
zo_encoder *zo_encoder_init(const char *const *tokens, const size_t *token_lens, const size_t *ids,
                            size_t n_tokens, const char *const *uni, const size_t *uni_lens,
                            const unsigned char *uni_byte, size_t n_uni) {
  zo_encoder *e = calloc(1, sizeof(*e));
  if (!e) return NULL;
  size_t total = 0, max_id = 0;
  for (size_t i = 0; i < n_tokens; ++i) { total += token_lens[i] + 1; if (ids[i] > max_id) max_id = ids[i]; }
  for (size_t i = 0; i < n_uni; ++i) total += uni_lens[i] + 1;
  e->arena = malloc(total ? total : 1);
  e->n_idx = max_id + 1;
  e->idx_to_token = calloc(e->n_idx, sizeof(char *));
  e->idx_to_token_len = calloc(e->n_idx, sizeof(size_t));
  if (!e->arena || !e->idx_to_token || !e->idx_to_token_len) return NULL;
  if (zo_map_init(&e->token_to_idx, n_tokens) || zo_map_init(&e->unicode_to_byte, n_uni)) return NULL;
  char *p = e->arena;
  for (size_t i = 0; i < n_tokens; ++i) { /* :20-24 */
    memcpy(p, tokens[i], token_lens[i]); p[token_lens[i]] = 0;
    zo_map_put(&e->token_to_idx, p, token_lens[i], ids[i]);
    e->idx_to_token[ids[i]] = p; e->idx_to_token_len[ids[i]] = token_lens[i];
    p += token_lens[i] + 1;
  }
  for (size_t i = 0; i < n_uni; ++i) { /* :25-29 */
    memcpy(p, uni[i], uni_lens[i]); p[uni_lens[i]] = 0;
    zo_map_put(&e->unicode_to_byte, p, uni_lens[i], uni_byte[i]);
    e->byte_to_unicode[uni_byte[i]] = p; e->byte_to_unicode_len[uni_byte[i]] = uni_lens[i];
    p += uni_lens[i] + 1;
  }
  /* :34-40 -- the five alternatives concatenated, REG_EXTENDED, C locale */
  static const char pattern[] =
      "'s|'t|'re|'ve|'m|'ll|'d"
      "|[[:space:]]?[[:alpha:]]+"
      "|[[:space:]]?[[:digit:]]+"
      "|[[:space:]]?[^[:space:][:alpha:][:digit:]]+"
      "|[[:space:]]+";
  e->c_locale = newlocale(LC_ALL_MASK, "C", (locale_t)0);
  locale_t prev = uselocale(e->c_locale);
  const int rc = regcomp(&e->regex, pattern, REG_EXTENDED);
  uselocale(prev);
  if (rc != 0) return NULL;
  return e;
}

void zo_encoder_deinit(zo_encoder *e) { /* :51-57 */
  if (!e) return;
  regfree(&e->regex);
  if (e->c_locale) freelocale(e->c_locale);
  free(e->token_to_idx.slots); free(e->unicode_to_byte.slots);
  free(e->idx_to_token); free(e->idx_to_token_len); free(e->arena); free(e);
}

/* bpe.zig:59-97 */
size_t zo_encoder_encode(const zo_encoder *e, const char *inputs, size_t len, size_t *outputs, size_t max_out) {
  regmatch_t matches[1];
  size_t token_idx = 0, offset = 0;
  locale_t prev = uselocale(e->c_locale);
  while (offset < len) {
    /* :65 -- the reference ignores the return code; a failed or empty match would make it spin forever on
     * stale offsets (only an embedded NUL can cause that in the C locale), so the oracle reports it instead */
    if (regexec(&e->regex, inputs + offset, 1, matches, 0) != 0 || matches[0].rm_eo <= 0) {
      uselocale(prev);
      return (size_t)-1;
    }
    const size_t match_so = offset + (size_t)matches[0].rm_so;
    const size_t match_eo = offset + (size_t)matches[0].rm_eo;

    char word[20]; /* :71 fixed 20-byte buffer */
    size_t word_eo = 0;
    for (size_t i = match_so; i < match_eo; ++i) { /* :73-78 bytes -> unicode */
      const unsigned char byte = (unsigned char)inputs[i];
      const char *u = e->byte_to_unicode[byte];
      if (!u) { uselocale(prev); return (size_t)-1; } /* `.?` unwrap panic in the reference */
      for (size_t j = 0; j < e->byte_to_unicode_len[byte]; ++j) {
        if (word_eo >= sizeof(word)) { uselocale(prev); return (size_t)-1; } /* reference overflows here (UB) */
        word[word_eo++] = u[j];
      }
    }
    size_t token_so = 0, token_eo = word_eo; /* :81-92 greedy longest prefix */
    while (token_so < token_eo) {
      const zo_slot *s = zo_map_get(&e->token_to_idx, word + token_so, token_eo - token_so);
      if (s) {
        if (token_idx >= max_out) { uselocale(prev); return (size_t)-1; }
        outputs[token_idx++] = s->val;
        token_so = token_eo;
        token_eo = word_eo;
      } else {
        token_eo -= 1; /* when this reaches token_so the rest of the word is silently dropped */
      }
    }
    offset = match_eo; /* :94 */
  }
  uselocale(prev);
  return token_idx;
}

/* bpe.zig:99-118 */
size_t zo_encoder_decode(const zo_encoder *e, const size_t *inputs, size_t n, unsigned char *outputs, size_t max_out) {
  size_t outputs_len = 0;
  for (size_t t = 0; t < n; ++t) {
    if (inputs[t] >= e->n_idx || !e->idx_to_token[inputs[t]]) return (size_t)-1; /* `.?` panic */
    const char *token = e->idx_to_token[inputs[t]];
    const size_t tl = e->idx_to_token_len[inputs[t]];
    size_t i = 0;
    while (i < tl) { /* :104-115: 1-byte key if present, else 2-byte key */
      const zo_slot *s = zo_map_get(&e->unicode_to_byte, token + i, 1);
      if (s) {
        i += 1;
      } else {
        if (i + 2 > tl) return (size_t)-1;
        s = zo_map_get(&e->unicode_to_byte, token + i, 2);
        if (!s) return (size_t)-1;
        i += 2;
      }
      if (outputs_len >= max_out) return (size_t)-1; /* reference writes past its 20-byte buffer */
      outputs[outputs_len++] = (unsigned char)s->val;
    }
  }
  return outputs_len;
}
