// zg_ptx.cuh -- inline-PTX wrappers (mbarrier, bulk async copy, acquire loads, timers) for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace zg {

constexpr int NCW = 7;             // consumer warps of the persistent decode kernel (+ 1 producer warp = 8 warps:
                                   // two per scheduler, so a thread may use up to 255 registers)
constexpr int NCT = NCW * 32;      // consumer threads

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with an explicit suspend-time hint: the thread sleeps in hardware until the phase completes or the
// hint expires, instead of spinning and competing for issue slots with the warps that share its scheduler
// (the arbiter favours the highest warp id, which is the producer warp).
__device__ __forceinline__ bool mbar_try_sleepy(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
// Watchdog: a protocol bug must not hang the GPU.  After ~2 s of waiting the kernel raises the sticky
// global error word and a CTA-local shared flag; every later wait in the CTA falls through at once.  The
// waits themselves never touch global memory (an earlier version polled the error word from the slow path
// and throttled the producer to ~1 unit/us).
constexpr long long WATCHDOG_CYCLES = 4000000000ll;
struct Watchdog {
  unsigned *err_global;   // sticky word the host checks after synchronising
  uint32_t tripped_smem;  // shared address of a CTA-local u32 flag
};
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity, Watchdog wd) {
  uint32_t tripped;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(tripped) : "r"(wd.tripped_smem));
  if (tripped) return;
  const long long t0 = clock64();
  while (!mbar_try_sleepy(bar, parity, 4000u)) {
    if (clock64() - t0 > WATCHDOG_CYCLES) {
      asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(wd.tripped_smem), "r"(1u));
      atomicExch(wd.err_global, 2u);
      return;
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, Watchdog wd) {
  if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity, wd);
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// bulk global -> shared copy (SASS: UBLKCP), completion signalled on an mbarrier as transaction bytes
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}
// 16-byte asynchronous global -> shared copy (SASS: LDGSTS), L2 only; completion via commit/wait groups
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acquire_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_release_gpu(unsigned *p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}


}  // namespace zg
