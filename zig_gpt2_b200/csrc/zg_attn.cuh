// zg_attn.cuh -- attention kernels of the batched paths (zg_attn.cu).
#pragma once
#include "zg_gemm.cuh"

namespace zg {

struct AttnPrefillPlan {
  CUtensorMap tm_qkv;  // f16 [B*T, 3E], box 128 rows x 64 columns, 128-byte swizzle
  void *out = nullptr; // f16 [B*T, E]
  int B = 0, T = 0, H = 0, E = 0;
};
bool attn_prefill_plan(AttnPrefillPlan *p, const void *qkv_f16, void *out_f16, int B, int T, int H, int E);
void attn_init_attrs();
void attn_prefill_launch(const AttnPrefillPlan &p);
void attn_decode_batch_launch(const float *q, int ldq, const float *k_cache, const float *v_cache, long long seq_stride,
                              int B, int H, int E, float *out, int ldo, const int *pos_dev, int pos_base,
                              const float *knew = nullptr, const float *vnew = nullptr, int rows_per_seq = 0);

void attn_decode_batch_launch_f16(const float *q, int ldq, const void *k_cache, const void *v_cache, long long seq_stride, int B,
                                  int H, int E, void *out_f16, int ldo, const int *pos_dev, const float *knew, const float *vnew);

}  // namespace zg
