#include "bpe_gpt2.hpp"

#include <climits>
#include <cstring>

#include "bpe.hpp"
#include "unicode_tables.hpp"

namespace zgh {

static bool in_ranges(const CpRange *r, unsigned n, uint32_t cp) {
  unsigned lo = 0, hi = n;
  while (lo < hi) {
    const unsigned mid = (lo + hi) / 2;
    if (cp < r[mid].first) hi = mid;
    else if (cp > r[mid].last) lo = mid + 1;
    else return true;
  }
  return false;
}
bool is_letter(uint32_t cp) { return in_ranges(kLetterRanges, kLetterRangesCount, cp); }
bool is_number(uint32_t cp) { return in_ranges(kNumberRanges, kNumberRangesCount, cp); }
bool is_whitespace(uint32_t cp) {  // Unicode White_Space
  return (cp >= 0x09 && cp <= 0x0D) || cp == 0x20 || cp == 0x85 || cp == 0xA0 || cp == 0x1680 ||
         (cp >= 0x2000 && cp <= 0x200A) || cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F || cp == 0x3000;
}

// one code point starting at s[i] (malformed bytes decode as themselves, one at a time: category "other")
static uint32_t decode_cp(const std::string &s, size_t i, size_t *len) {
  const unsigned char c = (unsigned char)s[i];
  auto cont = [&](size_t k) { return i + k < s.size() && (((unsigned char)s[i + k]) & 0xC0) == 0x80; };
  if (c < 0x80) { *len = 1; return c; }
  if ((c & 0xE0) == 0xC0 && cont(1)) { *len = 2; return ((c & 0x1Fu) << 6) | ((unsigned char)s[i + 1] & 0x3Fu); }
  if ((c & 0xF0) == 0xE0 && cont(1) && cont(2)) {
    *len = 3;
    return ((c & 0x0Fu) << 12) | (((unsigned char)s[i + 1] & 0x3Fu) << 6) | ((unsigned char)s[i + 2] & 0x3Fu);
  }
  if ((c & 0xF8) == 0xF0 && cont(1) && cont(2) && cont(3)) {
    *len = 4;
    return ((c & 0x07u) << 18) | (((unsigned char)s[i + 1] & 0x3Fu) << 12) | (((unsigned char)s[i + 2] & 0x3Fu) << 6) |
           ((unsigned char)s[i + 3] & 0x3Fu);
  }
  *len = 1;
  return 0xFFFD;
}

enum Cls { C_LETTER, C_NUMBER, C_SPACE, C_OTHER };
static Cls classify(uint32_t cp) {
  if (is_whitespace(cp)) return C_SPACE;
  if (is_letter(cp)) return C_LETTER;
  if (is_number(cp)) return C_NUMBER;
  return C_OTHER;
}

// Ordered alternation, leftmost match at every position -- what a backtracking regex engine does with GPT-2's pattern.
void gpt2_pretokenize(const std::string &text, std::vector<std::pair<size_t, size_t>> *pieces) {
  pieces->clear();
  // code points with byte offsets and classes
  std::vector<size_t> off;
  std::vector<Cls> cls;
  std::vector<uint32_t> cps;
  for (size_t i = 0; i < text.size();) {
    size_t len;
    const uint32_t cp = decode_cp(text, i, &len);
    off.push_back(i);
    cps.push_back(cp);
    cls.push_back(classify(cp));
    i += len;
  }
  off.push_back(text.size());
  const size_t n = cps.size();
  static const char *const kContractions[] = {"'s", "'t", "'re", "'ve", "'m", "'ll", "'d"};
  size_t i = 0;
  while (i < n) {
    size_t end = 0;  // code-point index one past the match; 0 = no alternative matched yet
    // 's|'t|'re|'ve|'m|'ll|'d
    if (cps[i] == '\'') {
      for (const char *c : kContractions) {
        const size_t l = strlen(c);
        if (off[i] + l <= text.size() && text.compare(off[i], l, c) == 0) { end = i + l; break; }  // ASCII: 1 byte per cp
      }
    }
    if (!end) {
      // " ?\p{L}+", " ?\p{N}+", " ?[^\s\p{L}\p{N}]+": an optional U+0020, then a run of one class
      const size_t j = (cps[i] == ' ' && i + 1 < n) ? i + 1 : i;
      const Cls c = cls[j];
      if (c != C_SPACE) {
        size_t k = j;
        while (k < n && cls[k] == c) ++k;
        end = k;
      } else if (j != i) {
        // the optional space was taken but a space class follows: retry without it (cps[i] is a space itself -> falls through)
      }
    }
    if (!end) {
      // here cls[i] == C_SPACE.  "\s+(?!\S)" then "\s+"
      size_t k = i;
      while (k < n && cls[k] == C_SPACE) ++k;
      if (k == n) end = k;              // the run reaches the end of the text: nothing non-space follows
      else if (k - i >= 2) end = k - 1;  // leave the last whitespace character for the next piece
      else end = k;                      // a single whitespace character before a non-space: "\s+"
    }
    pieces->emplace_back(off[i], off[end]);
    i = end;
  }
}

bool Gpt2Tokenizer::init(const std::vector<std::pair<std::string, long>> &token_to_idx,
                         const std::vector<std::pair<std::string, std::string>> &merges,
                         const std::vector<std::pair<std::string, long>> &unicode_to_byte) {
  token_to_idx_.clear(); idx_to_token_.clear(); rank_.clear(); unicode_to_byte_.clear();
  for (auto &s : byte_to_unicode_) s.clear();
  for (const auto &kv : token_to_idx) {
    if (kv.second < 0) return false;
    token_to_idx_[kv.first] = (size_t)kv.second;
    idx_to_token_[(size_t)kv.second] = kv.first;
  }
  for (size_t r = 0; r < merges.size(); ++r) rank_.emplace(merges[r].first + '\x01' + merges[r].second, r);
  for (const auto &kv : unicode_to_byte) {
    if (kv.second < 0 || kv.second > 255) return false;
    unicode_to_byte_[kv.first] = (unsigned char)kv.second;
    byte_to_unicode_[kv.second] = kv.first;
  }
  for (const auto &s : byte_to_unicode_)
    if (s.empty()) return false;  // every byte needs its printable stand-in
  return true;
}

bool Gpt2Tokenizer::init_from_files(const std::string &encoder_json, const std::string &merges_txt, const std::string &byte_encoder_json) {
  std::string a, b, m;
  if (!read_file(encoder_json, &a) || !read_file(byte_encoder_json, &b) || !read_file(merges_txt, &m)) return false;
  std::vector<std::pair<std::string, long>> t2i, u2b;
  if (!parse_flat_json(a, &t2i) || !parse_flat_json(b, &u2b)) return false;
  std::vector<std::pair<std::string, std::string>> merges;
  size_t i = 0;
  while (i < m.size()) {
    size_t e = m.find('\n', i);
    if (e == std::string::npos) e = m.size();
    std::string line = m.substr(i, e - i);
    i = e + 1;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line.compare(0, 9, "#version:") == 0) continue;
    const size_t sp = line.find(' ');
    if (sp == std::string::npos) return false;
    merges.emplace_back(line.substr(0, sp), line.substr(sp + 1));
  }
  return init(t2i, merges, u2b);
}

// byte-pair encoding of one piece: start from its bytes (as their printable stand-ins) and repeatedly merge the adjacent
// pair of lowest rank, all its occurrences left to right, until no adjacent pair has a rank
void Gpt2Tokenizer::bpe(const std::string &piece, std::vector<std::string> *sym) const {
  sym->clear();
  for (unsigned char c : piece) sym->push_back(byte_to_unicode_[c]);
  while (sym->size() > 1) {
    size_t best = SIZE_MAX;
    std::string key;
    for (size_t k = 0; k + 1 < sym->size(); ++k) {
      key.assign((*sym)[k]).push_back('\x01');
      key += (*sym)[k + 1];
      const auto it = rank_.find(key);
      if (it != rank_.end() && it->second < best) best = it->second;
    }
    if (best == SIZE_MAX) break;
    std::vector<std::string> merged;
    for (size_t k = 0; k < sym->size();) {
      if (k + 1 < sym->size()) {
        key.assign((*sym)[k]).push_back('\x01');
        key += (*sym)[k + 1];
        const auto it = rank_.find(key);
        if (it != rank_.end() && it->second == best) {
          merged.push_back((*sym)[k] + (*sym)[k + 1]);
          k += 2;
          continue;
        }
      }
      merged.push_back((*sym)[k]);
      ++k;
    }
    sym->swap(merged);
  }
}

size_t Gpt2Tokenizer::encode(const std::string &text, std::vector<size_t> *out) const {
  out->clear();
  std::vector<std::pair<size_t, size_t>> pieces;
  gpt2_pretokenize(text, &pieces);
  std::vector<std::string> sym;
  for (const auto &p : pieces) {
    bpe(text.substr(p.first, p.second - p.first), &sym);
    for (const auto &s : sym) {
      const auto it = token_to_idx_.find(s);
      if (it == token_to_idx_.end()) return (size_t)-1;
      out->push_back(it->second);
    }
  }
  return out->size();
}

size_t Gpt2Tokenizer::decode(const size_t *ids, size_t n, std::string *out) const {
  out->clear();
  for (size_t i = 0; i < n; ++i) {
    const auto it = idx_to_token_.find(ids[i]);
    if (it == idx_to_token_.end()) return (size_t)-1;
    const std::string &tok = it->second;
    for (size_t k = 0; k < tok.size();) {  // every character of a token is a byte's stand-in (1 or 2 bytes of UTF-8)
      size_t len = ((unsigned char)tok[k] & 0x80) ? 2 : 1;
      if (k + len > tok.size()) len = 1;
      const auto b = unicode_to_byte_.find(tok.substr(k, len));
      if (b == unicode_to_byte_.end()) return (size_t)-1;
      out->push_back((char)b->second);
      k += len;
    }
  }
  return out->size();
}

}  // namespace zgh

extern "C" {

void *zgh_gpt2_create_from_files(const char *encoder_json, const char *merges_txt, const char *byte_encoder_json) {
  auto *t = new zgh::Gpt2Tokenizer();
  if (!t->init_from_files(encoder_json, merges_txt, byte_encoder_json)) {
    delete t;
    return nullptr;
  }
  return t;
}
void zgh_gpt2_destroy(void *t) { delete static_cast<zgh::Gpt2Tokenizer *>(t); }
size_t zgh_gpt2_encode(const void *t, const char *text, size_t len, size_t *out, size_t max_out) {
  std::vector<size_t> ids;
  const size_t n = static_cast<const zgh::Gpt2Tokenizer *>(t)->encode(std::string(text, len), &ids);
  if (n == (size_t)-1 || n > max_out) return (size_t)-1;
  for (size_t i = 0; i < n; ++i) out[i] = ids[i];
  return n;
}
size_t zgh_gpt2_decode(const void *t, const size_t *ids, size_t n, unsigned char *out, size_t max_out) {
  std::string s;
  const size_t m = static_cast<const zgh::Gpt2Tokenizer *>(t)->decode(ids, n, &s);
  if (m == (size_t)-1 || m > max_out) return (size_t)-1;
  memcpy(out, s.data(), m);
  return m;
}
size_t zgh_gpt2_pretokenize(const char *text, size_t len, size_t *ends, size_t max_out) {
  std::vector<std::pair<size_t, size_t>> pieces;
  zgh::gpt2_pretokenize(std::string(text, len), &pieces);
  if (pieces.size() > max_out) return (size_t)-1;
  for (size_t i = 0; i < pieces.size(); ++i) ends[i] = pieces[i].second;
  return pieces.size();
}

}  // extern "C"
