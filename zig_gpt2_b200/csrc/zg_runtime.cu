// zg_runtime.cu -- lifecycle, memory and error plumbing of the C-ABI (include/zg_b200.h).
// Everything that allocates lives here or in the *_create / *_init entry points; the hot path never does.
#include <stdlib.h>
#include <string.h>

#include "zg_common.cuh"
#include "zg_philox.cuh"

namespace zg {

Context &ctx() {
  static Context c;
  return c;
}

void set_error(int code, const char *what, const char *file, int line) {
  Context &c = ctx();
  if (c.last_error != 0) return;  // sticky: keep the first failure
  c.last_error = code;
  const char *cs = (code > 1) ? cudaGetErrorString((cudaError_t)code) : "invalid argument / not initialised";
  snprintf(c.last_error_msg, sizeof(c.last_error_msg), "%s:%d: %s: %s", file, line, what, cs);
}

bool require_ready(const char *fn) {
  Context &c = ctx();
  if (!c.ready) {
    // No CPU fallback: without an initialised CUDA device every op fails loudly.
    set_error(1, fn, "zg_init() has not succeeded; there is no CPU fallback", 0);
    return false;
  }
  return true;
}

// ZG_PDL (A/B runs) is a bit mask over the kernels of the stream-K decode step:
//   1 the GEMMs are launched as programmatic dependents (prologue + weight prefetch before the predecessor has drained)
//   2 the LayerNorm row kernels trigger early        4 the attention kernel triggers early
//   8 the GEMMs trigger early                        16 / 32 the attention / row kernels are launched as dependents
// Default 47: a GEMM's ring is full of weights by the time the kernel that feeds it finishes, and the row kernels start
// under the tail of the GEMM before them.  Measured on B200 (1.5B, batch 64, context 1024, TF32), ms per step:
// 0 -> 9.09, 3 -> 8.92, 7 -> 8.81, 11 -> 8.75, 15 -> 8.65, 47 -> 8.63; with the attention kernel launched as a dependent
// (31, 63) -> 10.0: its 1,600 CTAs become resident and wait inside the c_attn GEMM.
int pdl_mask() {
  static const int m = getenv("ZG_PDL") ? atoi(getenv("ZG_PDL")) : 47;
  return m;
}

static void (*g_hooks[16])() = {nullptr};
static int g_nhooks = 0;
void register_shutdown_hook(void (*fn)()) {
  for (int i = 0; i < g_nhooks; ++i)
    if (g_hooks[i] == fn) return;
  if (g_nhooks < 16) g_hooks[g_nhooks++] = fn;
}

}  // namespace zg

using namespace zg;

static unsigned g_generation = 0;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;

extern "C" {

int zg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int zg_init(int device) {
  Context &c = ctx();
  if (c.ready && c.device == device) return 0;
  if (c.ready) zg_shutdown();
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    set_error((int)e, "cudaSetDevice", __FILE__, __LINE__);
    return (int)e;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    set_error((int)e, "cudaGetDeviceProperties", __FILE__, __LINE__);
    return (int)e;
  }
  if (prop.major != 10) {
    set_error(1, "zg_init: this library is built for sm_100a (Blackwell B200) only", __FILE__, __LINE__);
    return 1;
  }
  c.device = device;
  c.sm_count = prop.multiProcessorCount;
  e = cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error((int)e, "cudaStreamCreate", __FILE__, __LINE__);
    return (int)e;
  }
  c.stream = c.own_stream;
  c.idx_staging_cap = 1 << 16;
  c.scratch_floats = 1 << 20;
  if ((e = cudaMalloc(&c.idx_staging, c.idx_staging_cap * sizeof(size_t))) != cudaSuccess ||
      (e = cudaMalloc(&c.scratch, c.scratch_floats * sizeof(float))) != cudaSuccess ||
      (e = cudaMalloc(&c.token_slot, 8 * sizeof(unsigned long long))) != cudaSuccess ||
      (e = cudaHostAlloc(&c.token_slot_host, 8 * sizeof(unsigned long long), cudaHostAllocDefault)) != cudaSuccess) {
    set_error((int)e, "zg_init scratch", __FILE__, __LINE__);
    return (int)e;
  }
  c.launches = 0;
  c.alloc_calls = 0;
  c.generation = ++g_generation;
  c.ready = true;
  return 0;
}

int zg_shutdown(void) {
  Context &c = ctx();
  if (!c.ready) return 0;
  cudaStreamSynchronize(c.stream);
  for (int i = 0; i < g_nhooks; ++i) g_hooks[i]();  // per-device lazily created state of the other translation units
  if (g_ev0) {
    cudaEventDestroy(g_ev0);
    cudaEventDestroy(g_ev1);
    g_ev0 = g_ev1 = nullptr;
  }
  cudaFree(c.idx_staging);
  cudaFree(c.scratch);
  cudaFree(c.token_slot);
  cudaFreeHost(c.token_slot_host);
  cudaStreamDestroy(c.own_stream);
  c = Context();
  return 0;
}

int zg_sm_count(void) { return ctx().sm_count; }

void *zg_alloc(size_t bytes) {
  if (!require_ready("zg_alloc")) return nullptr;
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
  note_alloc();
  if (e != cudaSuccess) {
    set_error((int)e, "zg_alloc", __FILE__, __LINE__);
    return nullptr;
  }
  return p;
}

int zg_free(void *p) {
  cudaError_t e = cudaFree(p);
  return (int)e;
}

int zg_memset(void *p, int value, size_t bytes) {
  if (!require_ready("zg_memset")) return 1;
  cudaError_t e = cudaMemsetAsync(p, value, bytes, ctx().stream);
  if (e != cudaSuccess) set_error((int)e, "zg_memset", __FILE__, __LINE__);
  return (int)e;
}

int zg_upload(void *dst, const void *src, size_t bytes) {
  if (!require_ready("zg_upload")) return 1;
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx().stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx().stream);
  if (e != cudaSuccess) set_error((int)e, "zg_upload", __FILE__, __LINE__);
  return (int)e;
}

int zg_download(void *dst, const void *src, size_t bytes) {
  if (!require_ready("zg_download")) return 1;
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx().stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx().stream);
  if (e != cudaSuccess) set_error((int)e, "zg_download", __FILE__, __LINE__);
  return (int)e;
}

int zg_sync(void) {
  if (!require_ready("zg_sync")) return 1;
  cudaError_t e = cudaStreamSynchronize(ctx().stream);
  if (e != cudaSuccess) set_error((int)e, "zg_sync", __FILE__, __LINE__);
  return (int)e;
}

int zg_last_error(void) { return ctx().last_error; }
const char *zg_last_error_string(void) { return ctx().last_error ? ctx().last_error_msg : ""; }
void zg_clear_error(void) {
  ctx().last_error = 0;
  ctx().last_error_msg[0] = 0;
  cudaGetLastError();
}

int zg_set_stream(void *s) {
  if (!require_ready("zg_set_stream")) return 1;
  ctx().stream = s ? (cudaStream_t)s : ctx().own_stream;
  return 0;
}

// the uniform draw of sampling step `step` of sequence `sequence` (host copy of the device function: no device needed)
double zg_philox_uniform(unsigned long long seed, unsigned long long step, unsigned long long sequence) {
  return (double)philox_uniform(seed, step, sequence);
}
void zg_philox4x32_10(const unsigned counter[4], const unsigned key[2], unsigned out[4]) {
  const Philox4 r = philox4x32_10(counter[0], counter[1], counter[2], counter[3], key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = r.v[i];
}

unsigned long long zg_launch_count(void) { return ctx().launches; }
unsigned long long zg_alloc_count(void) { return ctx().alloc_calls; }

int zg_timer_begin(void) {
  if (!require_ready("zg_timer_begin")) return 1;
  if (!g_ev0) {
    ZG_CUDA(cudaEventCreate(&g_ev0));
    ZG_CUDA(cudaEventCreate(&g_ev1));
  }
  ZG_CUDA(cudaEventRecord(g_ev0, ctx().stream));
  return zg_last_error();
}
float zg_timer_end_ms(void) {
  if (!require_ready("zg_timer_end_ms") || !g_ev0) return -1.0f;
  float ms = -1.0f;
  ZG_CUDA(cudaEventRecord(g_ev1, ctx().stream));
  ZG_CUDA(cudaEventSynchronize(g_ev1));
  ZG_CUDA(cudaEventElapsedTime(&ms, g_ev0, g_ev1));
  return ms;
}

}  // extern "C"
