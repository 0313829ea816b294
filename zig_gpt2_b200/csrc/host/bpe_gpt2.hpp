// bpe_gpt2.hpp -- OPT-IN tokenizer fidelity mode: real byte-pair encoding with GPT-2's merge
// ranks (`vocab.bpe` / merges.txt) and GPT-2's pre-tokenizer pattern
//     's|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+
// The reference's src/bpe.zig is NOT this: its POSIX pattern has no lookahead and no literal-space prefix
// (bpe.zig:34-40) and its encoder is a greedy longest-prefix vocabulary match (bpe.zig:80-92), with 20-byte word and
// decode buffers (bpe.zig:71, main.zig:52).  `zgh::Encoder` (bpe.hpp) stays the default and stays bit-exact to the
// reference; this class is what produces the token ids GPT-2 was trained on, with no length limits.
// Pinned against transformers' GPT2Tokenizer on synthetic vocab / merges files (tests/test_tokenizer_gpt2.py).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace zgh {

// Unicode classes of the pattern (tables: unicode_tables.hpp, generated from Python's unicodedata)
bool is_letter(uint32_t cp);      // \p{L}
bool is_number(uint32_t cp);      // \p{N}
bool is_whitespace(uint32_t cp);  // \s (Unicode White_Space)

// The pre-tokenizer: byte offsets [begin, end) of consecutive pieces covering the whole text.
void gpt2_pretokenize(const std::string &text, std::vector<std::pair<size_t, size_t>> *pieces);

class Gpt2Tokenizer {
 public:
  // token_to_idx: encoder.json / vocab.json; merges: ordered (left, right) pairs of merges.txt (rank = position);
  // unicode_to_byte: byte_encoder.json (download_weights.py:69-90)
  bool init(const std::vector<std::pair<std::string, long>> &token_to_idx,
            const std::vector<std::pair<std::string, std::string>> &merges,
            const std::vector<std::pair<std::string, long>> &unicode_to_byte);
  bool init_from_files(const std::string &encoder_json, const std::string &merges_txt, const std::string &byte_encoder_json);
  size_t encode(const std::string &text, std::vector<size_t> *out) const;  // (size_t)-1: a piece has no vocabulary entry
  size_t decode(const size_t *ids, size_t n, std::string *out) const;      // (size_t)-1: unknown id

 private:
  void bpe(const std::string &piece_bytes, std::vector<std::string> *symbols) const;
  std::unordered_map<std::string, size_t> token_to_idx_;
  std::unordered_map<size_t, std::string> idx_to_token_;
  std::unordered_map<std::string, size_t> rank_;  // "left\x01right" -> merge rank
  std::unordered_map<std::string, unsigned char> unicode_to_byte_;
  std::string byte_to_unicode_[256];
};

}  // namespace zgh

extern "C" {
void *zgh_gpt2_create_from_files(const char *encoder_json, const char *merges_txt, const char *byte_encoder_json);
void zgh_gpt2_destroy(void *t);
size_t zgh_gpt2_encode(const void *t, const char *text, size_t len, size_t *out, size_t max_out);
size_t zgh_gpt2_decode(const void *t, const size_t *ids, size_t n, unsigned char *out, size_t max_out);
size_t zgh_gpt2_pretokenize(const char *text, size_t len, size_t *ends, size_t max_out);  // end offsets of the pieces
}
