// zg_tc.cuh -- inline-PTX wrappers for the Blackwell tensor path: TMA tensor copies, tcgen05.mma with TMEM
// accumulators, tcgen05.ld, UMMA shared-memory / instruction descriptors.  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace zg {
namespace tc {

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes or the hint expires, so
// a waiting single-lane role (TMA producer, MMA issuer) or epilogue warp does not burn issue slots of the scheduler
// it shares with the warps doing arithmetic.
#ifndef ZG_TC_WAIT_HINT_NS
#define ZG_TC_WAIT_HINT_NS 2000
#endif
__device__ __forceinline__ bool mbar_try_sleepy(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"((uint32_t)ZG_TC_WAIT_HINT_NS)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must never hang the GPU.  After ~1 s the CTA-local abort flag (shared) and the
// sticky global error word are raised; every later wait in the CTA falls through immediately.
struct Guard {
  unsigned *err_global;
  uint32_t abort_smem;  // shared address of a u32
};
static __device__ __noinline__ bool mbar_wait_slow(uint32_t bar, uint32_t parity, Guard g) {
  uint32_t aborted;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(aborted) : "r"(g.abort_smem));
  if (aborted) return false;
  const long long t0 = clock64();
  while (!mbar_try_sleepy(bar, parity)) {
    if (clock64() - t0 > 2000000000ll) {
      asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(g.abort_smem), "r"(1u));
      if (g.err_global) atomicExch(g.err_global, 3u);
      return false;
    }
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(aborted) : "r"(g.abort_smem));
    if (aborted) return false;
  }
  return true;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, Guard g) {
  if (mbar_try(bar, parity)) return true;
  return mbar_wait_slow(bar, parity, g);
}
// all 32 lanes wait; the verdict is made warp-uniform so that the *_u issue forms below stay convergent
__device__ __forceinline__ bool mbar_wait_u(uint32_t bar, uint32_t parity, Guard g) {
  return __all_sync(0xffffffffu, mbar_wait(bar, parity, g));
}

// ---- TMA ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// 2-D tiled copy global -> shared (SASS: UTMALDG); c0 = innermost (contiguous) coordinate, c1 = row
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *m, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap *m, int c0, int c1, uint32_t bar,
                                                 uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, "
      "%3}], [%4], %5;"
      ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(bar), "l"(pol)
      : "memory");
}
// 2-D tiled copy shared -> global (SASS: UTMASTG), bulk-group completion; `reduce_add` makes it out += tile in L2
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(m), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap *m, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(m), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (UMMA / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------------
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {  // whole warp; COLS power of two >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp, the one that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 columns of 32-bit: thread i of the warp gets lane (base_lane + i), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM: thread i of the warp writes lane (base_lane + i), columns [col, col + N)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors ---------------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle (the layout TMA writes with CU_TENSOR_MAP_SWIZZLE_128B into a
// 1024-byte aligned tile whose rows are 128 bytes): bits [0,14) address>>4, [16,30) leading byte offset>>4,
// [32,46) stride byte offset>>4 (8 rows x 128 B = 1024), bits [46,48) = 1 (sm_100 descriptor version),
// bits [61,64) = 2 (SWIZZLE_128B).  The same encoding serves K-major tiles (rows = M/N index, 128 B of K per row)
// and MN-major tiles (rows = K index, 128 B of M/N per row); which one is meant is said by the instruction
// descriptor's a_major / b_major bits.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor for kind::f16 / kind::tf32 with fp32 accumulation.
// fmt: 0 = f16, 1 = bf16, 2 = tf32.  major: 0 = K-major, 1 = MN-major.
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t fmt, uint32_t M, uint32_t N, uint32_t a_major,
                                                  uint32_t b_major) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_major << 15) | (b_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}
// D[tmem] (+)= A[smem] . B[smem]; issued by ONE thread for the whole CTA
template <bool TF32>
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// D[tmem] (+)= A[TMEM] . B[smem], kind::f16: the A operand (M = 128 rows = TMEM lanes, K-major, two f16 per 32-bit column)
// is read from tensor memory -- the P matrix of flash attention never touches shared memory
__device__ __forceinline__ void umma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 3xTF32 operand split, x = hi + lo with hi = x & 0xffffe000: kind::tf32 reads the fp32 word and ignores its low 13
// mantissa bits, so the raw word already IS the hi operand and only lo = x - hi is ever written (measured: the GEMM
// errors are bit-identical with and without storing the masked word; tests/test_gpu_batch.py pins the 3xTF32 bound).

// ---- warp-uniform issue ------------------------------------------------------------------------------------------
// The single-thread instructions (tcgen05.mma, tcgen05.commit, TMA, arrive.expect_tx) take their operands from UNIFORM
// registers.  Issued under `if (lane == 0)` the compiler sees divergent control flow, cannot keep descriptors in uniform
// registers and wraps every such instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall: ~90 cycles per MMA
// (measured: 755 cycles to issue the eight P V MMAs of one attention tile -- the tensor pipe starved by its own issuer).
// The *_u forms below are called by ALL 32 lanes of the role's warp with warp-uniform arguments; the election and the
// predication happen inside the asm block, so the surrounding address arithmetic stays on the uniform datapath.
template <bool TF32>
__device__ __forceinline__ void umma_u(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (TF32) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// A operand from TMEM (lane = tile row, one 32-bit column per K element for tf32, two packed f16 per column)
template <bool TF32>
__device__ __forceinline__ void umma_ts_u(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (TF32) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void umma_ts_f16_u(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  umma_ts_u<false>(d_tmem, a_tmem, bdesc, idesc, accumulate);
}
__device__ __forceinline__ void umma_commit_u(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_u(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d_u(uint32_t dst, const CUtensorMap *m, int c0, int c1, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n\t}"
      ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint_u(uint32_t dst, const CUtensorMap *m, int c0, int c1, uint32_t bar, uint64_t pol) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;\n\t}"
      ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(bar), "l"(pol)
      : "memory");
}

// ---- CTA pair (cta_group::2) -------------------------------------------------------------------------------------
// Two CTAs of a cluster on the two SMs of a TPC run one 256-row MMA: each holds its 128 rows of A and HALF of the B tile,
// the leader (cluster rank 0) issues, both tensor cores see the whole B tile.  Per SM that halves the B bytes written
// into and read out of shared memory, which is what bounds the single-CTA f16 GEMM.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `saddr` (a shared::cta address of this CTA's layout) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster_u(uint32_t cluster_bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;\n\t}" ::"r"(cluster_bar), "r"(bytes) : "memory");
}
// tile into THIS CTA's shared memory, transaction bytes onto the leader's barrier (a shared::cluster address)
__device__ __forceinline__ void tma_load_2d_pair_u(uint32_t dst, const CUtensorMap *m, int c0, int c1, uint32_t cluster_bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n\t}"
      ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(cluster_bar)
      : "memory");
}
template <bool TF32>
__device__ __forceinline__ void umma2_u(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (TF32) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs once every MMA issued so far has completed
__device__ __forceinline__ void umma2_commit_both_u(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(bar) : "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem) {  // the same warp of both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

}  // namespace tc

// ---- host side: tensor maps ---------------------------------------------------------------------------
// 2-D row-major tensor [rows, cols] of `elem_bytes`-wide elements with row pitch `pitch_bytes`; box = [box_rows,
// box_cols] with box_cols * elem_bytes == 128 (one swizzle row).  dtype: 0 = fp32 (tf32 operand), 1 = fp16.
// swizzle_bytes: 128 (default) or 64 -- must equal box_cols * elem_bytes -- or 0 for an unswizzled box (plain row-major
// rows of box_cols elements in shared memory: staged epilogue tiles that no tensor-core instruction reads).
bool make_tmap_2d(CUtensorMap *out, const void *base, int dtype, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                  uint32_t box_rows, uint32_t box_cols, int swizzle_bytes = 128);

}  // namespace zg
