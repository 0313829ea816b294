// bpe.hpp -- host-side tokenizer, the C++ mirror of the reference's src/bpe.zig (which stays CPU code in
// the Zig host too: it never touches the CUDA shim).  Same algorithm, same names:
//   Encoder.init   bpe.zig:14-49   reverse maps + POSIX ERE compile (REG_EXTENDED, C locale)
//   Encoder.encode bpe.zig:59-97   regex word split -> byte->unicode map -> greedy longest-prefix vocabulary match
//   Encoder.decode bpe.zig:99-118  id -> token string -> unicode->byte (1-byte key first, else 2-byte key)
// Documented divergence: the reference's fixed [20]u8 word buffer (bpe.zig:71) and 20-byte decode buffer
// (main.zig:52) overflow on longer words; here words and outputs are unbounded std::string/vector.
#pragma once
#include <locale.h>
#include <regex.h>

#include <cstddef>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace zgh {

// Parses a flat JSON object {"key": int, ...} (encoder.json / byte_encoder.json, main.zig:316-320),
// including \uXXXX escapes (surrogate pairs too) -> UTF-8 keys.  Returns false on malformed input.
bool parse_flat_json(const std::string &text, std::vector<std::pair<std::string, long>> *out);
bool read_file(const std::string &path, std::string *out);

class Encoder {
 public:
  Encoder() = default;
  ~Encoder();
  Encoder(const Encoder &) = delete;
  Encoder &operator=(const Encoder &) = delete;

  bool init(const std::vector<std::pair<std::string, long>> &token_to_idx,
            const std::vector<std::pair<std::string, long>> &unicode_to_byte);
  bool init_from_files(const std::string &encoder_json, const std::string &byte_encoder_json);
  void deinit();

  // `inputs` need not be NUL-terminated (a terminated copy is made; regexec takes a bare pointer, bpe.zig:65).
  size_t encode(const std::string &inputs, std::vector<size_t> *outputs) const;
  size_t decode(const size_t *inputs, size_t n, std::string *outputs) const;

 private:
  std::unordered_map<std::string, size_t> token_to_idx_;
  std::unordered_map<size_t, std::string> idx_to_token_;
  std::unordered_map<std::string, unsigned char> unicode_to_byte_;
  std::string byte_to_unicode_[256];
  bool have_byte_[256] = {false};
  regex_t regex_;
  bool compiled_ = false;
  locale_t c_locale_ = (locale_t)0;  // the Zig program never calls setlocale: its regex runs in the "C" locale
};

}  // namespace zgh

extern "C" {
// C-ABI for tests / bindings
void *zgh_encoder_create(const char *const *tokens, const size_t *token_lens, const size_t *ids, size_t n_tokens,
                         const char *const *uni, const size_t *uni_lens, const unsigned char *uni_byte, size_t n_uni);
void *zgh_encoder_create_from_files(const char *encoder_json, const char *byte_encoder_json);
void zgh_encoder_destroy(void *e);
size_t zgh_encoder_encode(const void *e, const char *inputs, size_t len, size_t *outputs, size_t max_out);
size_t zgh_encoder_decode(const void *e, const size_t *inputs, size_t n, unsigned char *outputs, size_t max_out);
}
