// zg_skinny.cuh -- interface of the swapped-operand stream-K GEMM of the batched decode step (zg_skinny.cu).
#pragma once
#include "zg_common.cuh"
#include "zg_tc.cuh"

namespace zg {

enum { SK_XFORM_NONE = 0, SK_XFORM_GELU = 1 };

// out[M, N] += bias + X[M, K] . W[N, K]^T.  `out` holds zeros or the residual on entry (partial sums are reduced into it).
struct SkinnyArgs {
  int M = 0, N = 0, K = 0;
  const float *bias = nullptr;  // [N] or null; added once per output element (by the segment that owns k-block 0)
  float *out = nullptr;         // [M, ldo] fp32
  int ldo = 0;
  int xform = SK_XFORM_NONE;    // transform of the landed X tile: GELU (main.zig:80) folded into mlp c_proj's operand load
  unsigned *err = nullptr;      // sticky device error word (watchdog)
  // set by skinny_plan: partial tiles leave through TMA reduce-adds (cp.reduce.async.bulk.tensor .add.f32, one instruction
  // per 32 x 32 chunk) instead of one fp32 atomic per element; needs a 16-byte aligned output with ldo % 4 == 0
  int tma_out = 0;
  int pdl_trigger = 0;  // set by skinny_plan: let the next kernel start launching early (PDL_GEMM_TRIGGER)
  // greedy sampling fused into the tied lm_head (main.zig:193 + argmax): no logits leave the kernel.  best[2 m] is raised
  // (atomicMax) to (orderable(logit) << 32 | ~column) for row m, so ties resolve to the first maximum like the reference
  // loop; work is split by whole weight tiles (no partial sums), `out` is not written.  best must be zero on entry.
  unsigned long long *best = nullptr;
};

struct SkinnyPlan {
  CUtensorMap tm_w, tm_x, tm_out;
  SkinnyArgs args;
  int mode = 1;  // 0 = f16 operands (16-bit weight storage), 1 = fp32 operands as tf32, 2 = 3xTF32 error-compensated
  int mb = 64;   // batch columns of the accumulator
  int grid = 0;
};

bool skinny_supported(int M, int N, int K);
// X: [M, K] row-major with pitch ldx elements; W: [N, K] row-major (Linear.weight, ops.zig:9); fp32, or f16 for mode 0
bool skinny_plan(SkinnyPlan *p, int mode, const void *X, size_t ldx, const void *W, const SkinnyArgs &args);
void skinny_launch(const SkinnyPlan &p);
void skinny_init_attrs();
// after a fused-argmax launch: unpack best[2 m] into tok[m] (and hist[*pos_dev][m] when hist != null)
void skinny_finish_argmax(const unsigned long long *best, unsigned long long *tok, unsigned long long *hist, int B, const int *pos_dev);

}  // namespace zg
