// zg_gemm.cuh -- interface of the tcgen05 GEMM behind Linear.forward for M >= 16 (zg_gemm.cu).
#pragma once
#include "zg_common.cuh"
#include "zg_tc.cuh"

namespace zg {

enum { TC_EPI_NONE = 0, TC_EPI_GELU = 1, TC_EPI_RESIDUAL = 2 };

// Epilogue description of one Linear call: out[M,N] = epi(bias + A[M,K] . W[N,K]^T).
struct GemmArgs {
  int M = 0, N = 0, K = 0;
  const float *bias = nullptr;  // [N] or null (lm_head, main.zig:312)
  void *out = nullptr;          // [M, ldo] fp32 or f16
  int ldo = 0;
  int out_f16 = 0;
  int epi = TC_EPI_NONE;
  int gelu_fast = 0;            // tanh.approx instead of tanhf (f16 pipelines only)
  const float *resid = nullptr;  // [M, ldr] fp32, added after the bias (main.zig:136-139,142-145); may alias out
  int ldr = 0;
  // c_attn only: columns [E, 2E) / [2E, 3E) are additionally appended, as fp32, to the K / V caches
  // (ops.zig:151-152,156-157): cache row = pos_base + *pos_dev + (row % rows_per_seq) of sequence row / rows_per_seq.
  float *k_cache = nullptr, *v_cache = nullptr;
  int E = 0, rows_per_seq = 1;
  long long cache_seq_stride = 0;  // floats between consecutive sequences' caches
  const int *pos_dev = nullptr;
  int pos_base = 0;
  unsigned *err = nullptr;  // sticky device error word (watchdog)
  // set by gemm_plan: the epilogue stages 32x32 chunks in swizzled shared memory and writes them with TMA stores
  // (full 128-byte lines) instead of one 16-byte store per row; a residual that aliases `out` becomes a TMA reduce-add
  int tma_out = 0, tma_reduce = 0, tma_kv = 0;
  int cache_rows = 0;  // rows per sequence in the caches (tma_kv addressing)
  int n_fastest = 0;   // tile order (set by gemm_plan)
  // split-K (set by gemm_plan, in-place residual GEMMs only): the K range is cut into `ksplit` slices that are separate
  // work items; every slice reduce-adds its partial tile into `out` (which already holds the residual), slice 0 adds
  // the bias.  Lets a skinny GEMM (few output tiles, long K) occupy every SM with wide tiles.
  int ksplit = 1;
  // fused greedy head (main.zig:192-194 + the caller's argmax): no output leaves the kernel; every tile raises its best
  // (order-preserving logit bits << 32 | ~column) per row into best[2 * row] with one atomicMax.  The caller zeroes
  // `best` before the launch and decodes it with skinny_finish_argmax.  First maximum wins ties, as in the reference.
  unsigned long long *best = nullptr;
};

// A prepared launch: tensor maps are encoded once (start-up for the engines, per call for the op-level API).
struct GemmPlan {
  CUtensorMap tm_a, tm_b, tm_out, tm_k, tm_v;
  GemmArgs args;
  int bn = 256;
  int mode = 1;  // 0 = fp16 operands, 1 = fp32 operands as tf32, 2 = fp32 operands, 3xTF32 error-compensated
  int grid = 0;
  int pair = 0;  // 1: CTA pairs on 256 x 256 tiles (cta_group::2), grid = 2 x pairs launched as clusters of two
};

// A: [M, K] row-major with pitch lda elements; W: [N, K] row-major (the reference's Linear.weight layout, ops.zig:9).
// mode: see GemmPlan::mode.  bn = 0 picks the tile width.
bool gemm_plan(GemmPlan *p, int mode, const void *A, size_t lda, const void *W, const GemmArgs &args, int bn);
void gemm_launch(const GemmPlan &p);
void gemm_init_attrs();
unsigned *gemm_error_word();  // device word shared by every tensor-core kernel of this library

}  // namespace zg
