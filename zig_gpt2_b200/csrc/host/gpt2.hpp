// gpt2.hpp -- C++ host mirror of the reference's src/main.zig over the CUDA C-ABI (include/zg_b200.h).
// The reference's host is Zig; no Zig toolchain exists in the build image, so this file is the compiled,
// tested twin of zig/src/main.zig (which a maintainer with Zig 0.11 builds against the same shim).
// Same names and call structure: GPTConfig, State, load_gpt, GPT::forward / sample, generate.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "../../../include/zg_b200.h"
#include "bpe.hpp"

namespace zgh {

struct GPTConfig {  // main.zig:5-23
  size_t vocab_size, context_size, n_layer, n_heads, n_embed;
  zg_config c() const { return zg_config{vocab_size, context_size, n_layer, n_heads, n_embed}; }
};

inline bool config_for_size(const std::string &size, GPTConfig *out) {  // main.zig:346 hard-codes 124M
  if (size == "124M") *out = {50257, 1024, 12, 12, 768};
  else if (size == "355M") *out = {50257, 1024, 24, 16, 1024};
  else if (size == "774M") *out = {50257, 1024, 36, 20, 1280};
  else if (size == "1.5B") *out = {50257, 1024, 48, 25, 1600};
  else return false;
  return true;
}

class GPT {  // main.zig:149-208; owns the device model, the preallocated State and the fused engine
 public:
  ~GPT() {
    if (engine_) zg_engine_destroy(engine_);
    if (loaded_) {
      zg_state_free(&state_);
      zg_gpt_free(&gpt_);
    }
  }
  // load_gpt (main.zig:304-314) + State.init (main.zig:46-64): every allocation happens here
  bool load(const GPTConfig &config, const std::string &model_dir, int device) {
    config_ = config;
    if (zg_init(device) != 0) return false;
    const zg_config c = config.c();
    if (zg_load_gpt(&gpt_, &c, (model_dir + "/raw").c_str()) != 0) return false;
    if (zg_state_init(&state_, &c, /*want_transpose_scratch=*/0) != 0) return false;
    loaded_ = true;
    engine_ = zg_engine_create(&gpt_, &state_);
    return engine_ != nullptr;
  }
  void forward(size_t seq_len, size_t token, bool compute_logits) {  // main.zig:178-195
    zg_engine_forward(engine_, seq_len, token, compute_logits ? 1 : 0);
  }
  size_t sample(size_t seq_len, float temp, size_t token, double u) {  // main.zig:198-207, explicit uniform draw
    return zg_engine_sample(engine_, seq_len, temp, token, u);
  }
  size_t sample_greedy(size_t seq_len, size_t token) { return zg_engine_sample_greedy(engine_, seq_len, token); }
  // the whole sampling loop on the device, u(step) = zg_philox_uniform(seed, step, 0): reproducible with --seed
  int generate_sample(const std::vector<size_t> &inputs, size_t n_total, float temp, uint64_t seed, std::vector<size_t> *out) {
    out->assign(n_total, 0);
    return zg_engine_generate_sample(engine_, inputs.data(), inputs.size(), n_total, temp, seed, 0, out->data());
  }
  int generate_greedy(const std::vector<size_t> &inputs, size_t n_total, std::vector<size_t> *out) {
    out->assign(n_total, 0);
    return zg_engine_generate_greedy(engine_, inputs.data(), inputs.size(), n_total, out->data());
  }
  const GPTConfig &config() const { return config_; }

 private:
  GPTConfig config_{};
  zg_gpt gpt_{};
  zg_state state_{};
  zg_engine *engine_ = nullptr;
  bool loaded_ = false;
};

// generate (main.zig:322-342): prompt tokens one at a time without logits, then sampling up to n_total; the last
// prompt token is forwarded twice, as in the reference.  Every token (prompt included) is decoded and emitted.
// `decode` turns one token id into its bytes: Encoder::decode (bpe.zig:99-118) or the fidelity tokenizer's.
inline void generate(GPT &gpt, const std::function<void(size_t, std::string *)> &decode, float temp,
                     const std::vector<size_t> &inputs, size_t n_total, bool greedy, uint64_t seed,
                     const std::function<void(const std::string &)> &emit) {
  std::string piece;
  // greedy: one persistent-kernel launch for the whole loop; sampling: one launch + one sampling kernel per token,
  // still without a host round trip (GPT.sample's draw comes from the counter-based generator, main.zig:198-207)
  std::vector<size_t> toks;
  const int rc = greedy ? gpt.generate_greedy(inputs, n_total, &toks) : gpt.generate_sample(inputs, n_total, temp, seed, &toks);
  if (rc != 0) return;
  for (size_t t : toks) {
    piece.clear();
    decode(t, &piece);
    emit(piece);
  }
}
inline void generate(GPT &gpt, const Encoder &encoder, float temp, const std::vector<size_t> &inputs, size_t n_total,
                     bool greedy, uint64_t seed, const std::function<void(const std::string &)> &emit) {
  generate(gpt, [&](size_t t, std::string *out) { encoder.decode(&t, 1, out); }, temp, inputs, n_total, greedy, seed, emit);
}

}  // namespace zgh
