// zg_common.cuh -- shared host/device helpers for the sm_100a kernels behind include/zg_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/zg_b200.h"

namespace zg {

// ---- runtime context (one per process; the reference is single-threaded, main.zig:344-371) ----
struct Context {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  int last_error = 0;
  char last_error_msg[256] = {0};
  unsigned long long launches = 0;
  unsigned generation = 0;           // bumped by every successful zg_init: per-device lazily initialised state keys on it
  unsigned long long alloc_calls = 0;  // cudaMalloc / cudaHostAlloc / tensor-map encodes since zg_init (zg_alloc_count)
  bool capturing = false;              // a CUDA graph is being captured: launches are recorded, not executed
  // start-up scratch
  size_t *idx_staging = nullptr;  // device, for Embedding.forward with many indices
  size_t idx_staging_cap = 0;
  float *scratch = nullptr;  // device scratch for split-KV partials etc.
  size_t scratch_floats = 0;
  unsigned long long *token_slot = nullptr;  // device, 1 x u64 (argmax / sample result)
  unsigned long long *token_slot_host = nullptr;  // pinned
};
Context &ctx();
void set_error(int code, const char *what, const char *file, int line);
bool require_ready(const char *fn);
// Called by zg_shutdown before the context is torn down: frees / forgets lazily created per-device state (watchdog
// words, timer events, which engine owns the __constant__ layer table), so that zg_shutdown + zg_init -- on the same
// or on another device -- starts clean.
void register_shutdown_hook(void (*fn)());
inline void note_alloc() { ctx().alloc_calls++; }

#define ZG_CUDA(expr)                                                   \
  do {                                                                  \
    cudaError_t _e = (expr);                                            \
    if (_e != cudaSuccess) ::zg::set_error((int)_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define ZG_LAUNCH_CHECK()                                               \
  do {                                                                  \
    if (!::zg::ctx().capturing) ::zg::ctx().launches++;                 \
    cudaError_t _e = cudaGetLastError();                                \
    if (_e != cudaSuccess) ::zg::set_error((int)_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

// ---- device helpers ---------------------------------------------------------------------------
#ifdef __CUDACC__
// Programmatic dependent launch (PDL): a kernel launched with launch_pdl may begin -- prologue, barrier / TMEM set-up,
// loads of data that no earlier kernel writes (weights) -- while its predecessor in the stream is still draining;
// pdl_wait() returns once every prerequisite grid has completed and its writes are visible, and must precede the first
// access to anything an earlier kernel produces and the first global write.  pdl_trigger() lets the NEXT kernel start
// launching.  Both are no-ops in a kernel launched the ordinary way.
// PITFALL (found the hard way): data the predecessor writes must NOT be read through a `const T *__restrict__` parameter in
// a dependent kernel.  Such loads become LDG.CONSTANT (ld.global.nc) of memory the compiler may assume is never written
// during the kernel, so it is free to hoist them ABOVE the wait despite the "memory" clobber -- the LayerNorm kernel read
// x while the GEMM before it was still reduce-adding into it (SASS: the first LDG of x sat two instructions before
// ACQBULK).  Predecessor-produced operands are plain pointers; __restrict__ / __ldg only for data no kernel of the step writes.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

enum { PDL_GEMM_DEP = 1, PDL_LN_TRIGGER = 2, PDL_ATTN_TRIGGER = 4, PDL_GEMM_TRIGGER = 8, PDL_ATTN_DEP = 16, PDL_LN_DEP = 32 };
int pdl_mask();  // which kernels of the stream-K decode step take part (zg_runtime.cu; environment ZG_PDL for A/B runs)
// the launch carries the PDL attribute only when `bit` is set in pdl_mask()
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(int bit, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (pdl_mask() & bit) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// streaming 128-bit read-only load that does not allocate in L1 (weights are read once per token)
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// The reference's GELU, ops.zig:225: 0.5 x (1 + tanh(x * 0.7978845608 * (1 + 0.044715 x^2)))
__device__ __forceinline__ float gelu_ref(float x) {
  return 0.5f * x * (1.0f + tanhf(x * 0.7978845608f * (1.0f + 0.044715f * x * x)));
}
// block-wide sum / max over blockDim.x threads (multiple of 32, <= 1024); `red` is 32 floats of smem
__device__ __forceinline__ float block_sum(float v, float *red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.0f;
  return warp_sum(t);
}
__device__ __forceinline__ float block_max(float v, float *red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : -INFINITY;
  return warp_max(t);
}
#endif

}  // namespace zg
