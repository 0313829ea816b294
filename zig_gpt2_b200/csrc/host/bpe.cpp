#include "bpe.hpp"

#include <cstdio>
#include <cstring>

namespace zgh {

bool read_file(const std::string &path, std::string *out) {
  FILE *f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  out->resize(n > 0 ? (size_t)n : 0);
  size_t got = n > 0 ? std::fread(&(*out)[0], 1, (size_t)n, f) : 0;
  std::fclose(f);
  return got == out->size();
}

static void append_utf8(std::string *s, uint32_t cp) {
  if (cp < 0x80) {
    s->push_back((char)cp);
  } else if (cp < 0x800) {
    s->push_back((char)(0xC0 | (cp >> 6)));
    s->push_back((char)(0x80 | (cp & 0x3F)));
  } else if (cp < 0x10000) {
    s->push_back((char)(0xE0 | (cp >> 12)));
    s->push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
    s->push_back((char)(0x80 | (cp & 0x3F)));
  } else {
    s->push_back((char)(0xF0 | (cp >> 18)));
    s->push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
    s->push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
    s->push_back((char)(0x80 | (cp & 0x3F)));
  }
}

static bool hex4(const std::string &t, size_t i, uint32_t *v) {
  if (i + 4 > t.size()) return false;
  uint32_t r = 0;
  for (size_t k = 0; k < 4; ++k) {
    const char c = t[i + k];
    r <<= 4;
    if (c >= '0' && c <= '9') r |= (uint32_t)(c - '0');
    else if (c >= 'a' && c <= 'f') r |= (uint32_t)(c - 'a' + 10);
    else if (c >= 'A' && c <= 'F') r |= (uint32_t)(c - 'A' + 10);
    else return false;
  }
  *v = r;
  return true;
}

bool parse_flat_json(const std::string &t, std::vector<std::pair<std::string, long>> *out) {
  size_t i = 0;
  auto ws = [&]() { while (i < t.size() && (t[i] == ' ' || t[i] == '\n' || t[i] == '\t' || t[i] == '\r')) ++i; };
  ws();
  if (i >= t.size() || t[i] != '{') return false;
  ++i;
  ws();
  if (i < t.size() && t[i] == '}') return true;
  while (i < t.size()) {
    ws();
    if (t[i] != '"') return false;
    ++i;
    std::string key;
    while (i < t.size() && t[i] != '"') {
      if (t[i] == '\\') {
        if (++i >= t.size()) return false;
        switch (t[i]) {
          case 'n': key.push_back('\n'); break;
          case 't': key.push_back('\t'); break;
          case 'r': key.push_back('\r'); break;
          case 'b': key.push_back('\b'); break;
          case 'f': key.push_back('\f'); break;
          case 'u': {
            uint32_t cp;
            if (!hex4(t, i + 1, &cp)) return false;
            i += 4;
            if (cp >= 0xD800 && cp <= 0xDBFF && i + 6 < t.size() && t[i + 1] == '\\' && t[i + 2] == 'u') {
              uint32_t lo;
              if (!hex4(t, i + 3, &lo)) return false;
              if (lo >= 0xDC00 && lo <= 0xDFFF) {
                cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                i += 6;
              }
            }
            append_utf8(&key, cp);
            break;
          }
          default: key.push_back(t[i]);  // \" \\ \/
        }
        ++i;
      } else {
        key.push_back(t[i++]);
      }
    }
    if (i >= t.size()) return false;
    ++i;  // closing quote
    ws();
    if (i >= t.size() || t[i] != ':') return false;
    ++i;
    ws();
    bool neg = false;
    if (i < t.size() && t[i] == '-') { neg = true; ++i; }
    if (i >= t.size() || t[i] < '0' || t[i] > '9') return false;
    long v = 0;
    while (i < t.size() && t[i] >= '0' && t[i] <= '9') v = v * 10 + (t[i++] - '0');
    out->emplace_back(std::move(key), neg ? -v : v);
    ws();
    if (i < t.size() && t[i] == ',') { ++i; continue; }
    if (i < t.size() && t[i] == '}') return true;
    return false;
  }
  return false;
}

Encoder::~Encoder() { deinit(); }

void Encoder::deinit() {  // bpe.zig:51-57
  if (compiled_) regfree(&regex_);
  compiled_ = false;
  if (c_locale_) freelocale(c_locale_);
  c_locale_ = (locale_t)0;
  token_to_idx_.clear();
  idx_to_token_.clear();
  unicode_to_byte_.clear();
}

bool Encoder::init(const std::vector<std::pair<std::string, long>> &token_to_idx,
                   const std::vector<std::pair<std::string, long>> &unicode_to_byte) {
  deinit();
  token_to_idx_.reserve(token_to_idx.size() * 2);
  idx_to_token_.reserve(token_to_idx.size() * 2);
  for (const auto &kv : token_to_idx) {  // bpe.zig:20-24
    token_to_idx_[kv.first] = (size_t)kv.second;
    idx_to_token_[(size_t)kv.second] = kv.first;
  }
  for (const auto &kv : unicode_to_byte) {  // bpe.zig:25-29
    if (kv.second < 0 || kv.second > 255) return false;
    unicode_to_byte_[kv.first] = (unsigned char)kv.second;
    byte_to_unicode_[kv.second] = kv.first;
    have_byte_[kv.second] = true;
  }
  // bpe.zig:34-40: five alternatives concatenated without spaces; REG_EXTENDED; no setlocale => C locale
  static const char kPattern[] =
      "'s|'t|'re|'ve|'m|'ll|'d"
      "|[[:space:]]?[[:alpha:]]+"
      "|[[:space:]]?[[:digit:]]+"
      "|[[:space:]]?[^[:space:][:alpha:][:digit:]]+"
      "|[[:space:]]+";
  c_locale_ = newlocale(LC_ALL_MASK, "C", (locale_t)0);
  locale_t prev = uselocale(c_locale_);
  compiled_ = regcomp(&regex_, kPattern, REG_EXTENDED) == 0;
  uselocale(prev);
  return compiled_;
}

bool Encoder::init_from_files(const std::string &encoder_json, const std::string &byte_encoder_json) {
  std::string a, b;
  if (!read_file(encoder_json, &a) || !read_file(byte_encoder_json, &b)) return false;
  std::vector<std::pair<std::string, long>> t2i, u2b;
  if (!parse_flat_json(a, &t2i) || !parse_flat_json(b, &u2b)) return false;
  return init(t2i, u2b);
}

size_t Encoder::encode(const std::string &inputs, std::vector<size_t> *outputs) const {  // bpe.zig:59-97
  const char *base = inputs.c_str();  // NUL-terminated
  const size_t len = inputs.size();
  regmatch_t matches[1];
  size_t offset = 0, emitted = 0;
  std::string word;
  struct LocaleGuard {
    locale_t prev;
    explicit LocaleGuard(locale_t l) : prev(uselocale(l)) {}
    ~LocaleGuard() { uselocale(prev); }
  } guard(c_locale_);
  while (offset < len) {
    if (regexec(&regex_, base + offset, 1, matches, 0) != 0 || matches[0].rm_eo <= 0) break;  // embedded NUL ends the text
    const size_t match_so = offset + (size_t)matches[0].rm_so, match_eo = offset + (size_t)matches[0].rm_eo;
    word.clear();
    for (size_t i = match_so; i < match_eo; ++i) {  // bpe.zig:73-78
      const unsigned char b = (unsigned char)base[i];
      if (!have_byte_[b]) return (size_t)-1;
      word += byte_to_unicode_[b];
    }
    size_t token_so = 0, token_eo = word.size();  // bpe.zig:81-92
    while (token_so < token_eo) {
      auto it = token_to_idx_.find(word.substr(token_so, token_eo - token_so));
      if (it != token_to_idx_.end()) {
        outputs->push_back(it->second);
        ++emitted;
        token_so = token_eo;
        token_eo = word.size();
      } else {
        token_eo -= 1;  // reaching token_so drops the rest of the word, as the reference does
      }
    }
    offset = match_eo;  // bpe.zig:94
  }
  return emitted;
}

size_t Encoder::decode(const size_t *inputs, size_t n, std::string *outputs) const {  // bpe.zig:99-118
  size_t produced = 0;
  for (size_t t = 0; t < n; ++t) {
    auto it = idx_to_token_.find(inputs[t]);
    if (it == idx_to_token_.end()) return (size_t)-1;
    const std::string &token = it->second;
    size_t i = 0;
    while (i < token.size()) {
      auto u = unicode_to_byte_.find(token.substr(i, 1));
      if (u != unicode_to_byte_.end()) {
        i += 1;
      } else {
        if (i + 2 > token.size()) return (size_t)-1;
        u = unicode_to_byte_.find(token.substr(i, 2));
        if (u == unicode_to_byte_.end()) return (size_t)-1;
        i += 2;
      }
      outputs->push_back((char)u->second);
      ++produced;
    }
  }
  return produced;
}

}  // namespace zgh

extern "C" {

void *zgh_encoder_create(const char *const *tokens, const size_t *token_lens, const size_t *ids, size_t n_tokens,
                         const char *const *uni, const size_t *uni_lens, const unsigned char *uni_byte, size_t n_uni) {
  std::vector<std::pair<std::string, long>> t2i, u2b;
  t2i.reserve(n_tokens);
  for (size_t i = 0; i < n_tokens; ++i) t2i.emplace_back(std::string(tokens[i], token_lens[i]), (long)ids[i]);
  for (size_t i = 0; i < n_uni; ++i) u2b.emplace_back(std::string(uni[i], uni_lens[i]), (long)uni_byte[i]);
  auto *e = new zgh::Encoder();
  if (!e->init(t2i, u2b)) {
    delete e;
    return nullptr;
  }
  return e;
}

void *zgh_encoder_create_from_files(const char *encoder_json, const char *byte_encoder_json) {
  auto *e = new zgh::Encoder();
  if (!e->init_from_files(encoder_json, byte_encoder_json)) {
    delete e;
    return nullptr;
  }
  return e;
}

void zgh_encoder_destroy(void *e) { delete static_cast<zgh::Encoder *>(e); }

size_t zgh_encoder_encode(const void *e, const char *inputs, size_t len, size_t *outputs, size_t max_out) {
  std::vector<size_t> out;
  const size_t n = static_cast<const zgh::Encoder *>(e)->encode(std::string(inputs, len), &out);
  if (n == (size_t)-1 || n > max_out) return (size_t)-1;
  std::memcpy(outputs, out.data(), n * sizeof(size_t));
  return n;
}

size_t zgh_encoder_decode(const void *e, const size_t *inputs, size_t n, unsigned char *outputs, size_t max_out) {
  std::string out;
  const size_t m = static_cast<const zgh::Encoder *>(e)->decode(inputs, n, &out);
  if (m == (size_t)-1 || m > max_out) return (size_t)-1;
  std::memcpy(outputs, out.data(), m);
  return m;
}

}  // extern "C"
