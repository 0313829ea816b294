// main.cpp -- `zig_gpt2 "<prompt>"` (main.zig:344-371) over the CUDA shim.
//   zig_gpt2 [--size 124M | --config V,C,L,H,E] [--model-dir models/124M] [--device 0] [--greedy] [--temp 0.8]
//            [--seed N] [--max-tokens N] [--tokenizer reference|gpt2 [--merges FILE]] "<prompt>"
// --tokenizer gpt2 is the opt-in fidelity mode (bpe_gpt2.hpp: real merges + GPT-2's pre-tokenizer pattern); the default is
// the bit-exact mirror of src/bpe.zig.
// Built by zig_gpt2_b200/build.py:build_cli (g++ main.cpp bpe.cpp + libzg_b200.so) -> zig_gpt2_b200/zig_gpt2.bin;
// tests/test_gpu_cli.py runs it end to end against the oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>

#include "bpe_gpt2.hpp"
#include "gpt2.hpp"

int main(int argc, char **argv) {
  std::string size = "124M", model_dir, prompt, custom, tokenizer = "reference", merges;
  int device = 0;
  bool greedy = false;
  float temp = 0.8f;  // main.zig:345
  uint64_t seed = (uint64_t)time(nullptr);
  size_t max_tokens = 0;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&](const char *what) -> const char * {
      if (i + 1 >= argc) { std::cerr << what << " needs a value\n"; exit(2); }
      return argv[++i];
    };
    if (a == "--size") size = next("--size");
    else if (a == "--config") custom = next("--config");  // vocab,context,layers,heads,embed (synthetic test models)
    else if (a == "--model-dir") model_dir = next("--model-dir");
    else if (a == "--device") device = atoi(next("--device"));
    else if (a == "--greedy") greedy = true;
    else if (a == "--temp") temp = (float)atof(next("--temp"));
    else if (a == "--seed") seed = strtoull(next("--seed"), nullptr, 10);
    else if (a == "--max-tokens") max_tokens = strtoull(next("--max-tokens"), nullptr, 10);
    else if (a == "--tokenizer") tokenizer = next("--tokenizer");
    else if (a == "--merges") merges = next("--merges");
    else prompt = a;
  }
  if (prompt.empty()) {  // the reference indexes args[1] unchecked (main.zig:361)
    std::cerr << "usage: zig_gpt2 [--size S] [--model-dir D] [--greedy] [--temp T] [--seed N] [--max-tokens N] \"<prompt>\"\n";
    return 2;
  }
  if (model_dir.empty()) model_dir = "models/" + size;
  zgh::GPTConfig config;
  if (!custom.empty()) {
    unsigned long v[5];
    if (sscanf(custom.c_str(), "%lu,%lu,%lu,%lu,%lu", &v[0], &v[1], &v[2], &v[3], &v[4]) != 5) {
      std::cerr << "--config wants vocab,context,layers,heads,embed\n";
      return 2;
    }
    config = {v[0], v[1], v[2], v[3], v[4]};
  } else if (!zgh::config_for_size(size, &config)) { std::cerr << "unknown size " << size << "\n"; return 2; }

  zgh::Encoder encoder;  // load_encoder, main.zig:316-320
  zgh::Gpt2Tokenizer gpt2_tok;
  const bool fidelity = tokenizer == "gpt2";
  if (fidelity) {
    if (merges.empty()) merges = model_dir + "/vocab.bpe";
    if (!gpt2_tok.init_from_files(model_dir + "/encoder.json", merges, model_dir + "/byte_encoder.json")) {
      std::cerr << "cannot load " << model_dir << "/encoder.json + byte_encoder.json + " << merges << "\n";
      return 1;
    }
  } else if (!encoder.init_from_files(model_dir + "/encoder.json", model_dir + "/byte_encoder.json")) {
    std::cerr << "cannot load " << model_dir << "/encoder.json + byte_encoder.json\n";
    return 1;
  }
  zgh::GPT gpt;
  if (!gpt.load(config, model_dir, device)) {
    std::cerr << "cannot load model: " << zg_last_error_string() << "\n";
    return 1;
  }
  std::vector<size_t> inputs;
  const size_t n_prompt = fidelity ? gpt2_tok.encode(prompt, &inputs) : encoder.encode(prompt, &inputs);
  if (n_prompt == (size_t)-1 || inputs.empty() || inputs.size() > config.context_size) {
    std::cerr << "prompt does not tokenize into 1.." << config.context_size << " tokens\n";
    return 1;
  }
  const size_t n_total = max_tokens ? std::min(config.context_size, inputs.size() + max_tokens) : config.context_size;
  zgh::generate(gpt,
                [&](size_t t, std::string *out) {
                  if (fidelity) gpt2_tok.decode(&t, 1, out);
                  else encoder.decode(&t, 1, out);
                },
                temp, inputs, n_total, greedy, seed,
                [](const std::string &piece) { std::cerr << piece << std::flush; });  // main.zig:340 prints to stderr
  std::cerr << "\n";
  if (zg_last_error()) { std::cerr << zg_last_error_string() << "\n"; return 1; }
  return 0;
}
