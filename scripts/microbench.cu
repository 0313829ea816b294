// microbench.cu -- isolates the two costs of the persistent decode kernel on a B200:
//   (1) grid-barrier latency for 148 co-resident CTAs, several implementations;
//   (2) weight-stream throughput of the cp.async.bulk + mbarrier ring, vs ring depth / unit size / consumer work;
//   (3) a plain LDG.128 read kernel as the achievable-HBM-read reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/microbench scripts/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../zig_gpt2_b200/csrc/zg_common.cuh"
#include "../zig_gpt2_b200/csrc/zg_ptx.cuh"

using namespace zg;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void red_release_add(unsigned *p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(288, 1) barrier_bench(unsigned *bar, int iters, unsigned long long *out) {
  if (threadIdx.x >= NCT) return;
  const int G = gridDim.x;
  unsigned target = 0;
  unsigned long long t0 = 0;
  for (int it = 0; it < iters + 10; ++it) {
    if (it == 10 && blockIdx.x == 0 && threadIdx.x == 0) t0 = globaltimer();
    target += G;
    consumer_sync();
    if (threadIdx.x == 0) {
      if (MODE == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while ((int)(ld_acquire(bar) - target) < 0) {}
      } else if (MODE == 1) {
        red_release_add(bar, 1u);
        while ((int)(ld_acquire(bar) - target) < 0) {}
      } else if (MODE == 2) {
        red_release_add(bar, 1u);
        while ((int)(ld_relaxed(bar) - target) < 0) {}
        asm volatile("fence.acquire.gpu;" ::: "memory");
      } else if (MODE == 3) {  // hierarchical: arrive on one of 4 counters (by cta%4), poll a flag written by last arriver
        __threadfence();
        const unsigned old = atomicAdd(bar + 32 * (1 + (blockIdx.x & 3)), 1u);
        (void)old;
        // each CTA polls all 4 sub-counters (spread contention over 4 L2 lines)
        const unsigned sub = target / 4;  // G divisible by 4 (148)
        for (int k = 0; k < 4; ++k)
          while ((int)(ld_acquire(bar + 32 * (1 + k)) - sub) < 0) {}
      }
    }
    consumer_sync();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = globaltimer() - t0;
}

// K sub-counters on separate 128-byte lines: each CTA posts one red.release to counter (cta % K); lanes 0..K-1 of
// warp 0 poll the K counters in parallel (one L2 round trip per poll round instead of K).
template <int K, int SLEEP>
__global__ void __launch_bounds__(288, 1) barrier_bench_k(unsigned *bar, int iters, unsigned long long *out) {
  if (threadIdx.x >= NCT) return;
  const int G = gridDim.x;
  const int lane = threadIdx.x & 31;
  const unsigned my_count = (lane < K) ? (unsigned)((G - lane + K - 1) / K) : 0u;
  unsigned epoch = 0;
  unsigned long long t0 = 0;
  for (int it = 0; it < iters + 10; ++it) {
    if (it == 10 && blockIdx.x == 0 && threadIdx.x == 0) t0 = globaltimer();
    ++epoch;
    consumer_sync();
    if (threadIdx.x < 32) {
      if (lane == 0) red_release_add(bar + 32 * (blockIdx.x % K), 1u);
      const unsigned want = epoch * my_count;
      bool done = (lane >= K);
      while (true) {
        if (!done) done = (int)(ld_acquire(bar + 32 * lane) - want) >= 0;
        if (__all_sync(0xffffffffu, done)) break;
        if (SLEEP) __nanosleep(SLEEP);
      }
    }
    consumer_sync();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = globaltimer() - t0;
}

// reduction of one E-vector partial per CTA into a shared accumulator, followed by a grid barrier
// MODE 0: none (barrier only)  1: red.add.u64 (fixed point)  2: red.add.v4.f32  3: red.add.f32  4: u64 over REP replicas
template <int MODE, int REP>
__global__ void __launch_bounds__(288, 1) reduce_bench(unsigned *bar, unsigned long long *acc64, float *acc32, int E, int iters,
                                                       unsigned long long *out) {
  if (threadIdx.x >= NCT) return;
  const int G = gridDim.x;
  unsigned target = 0;
  unsigned long long t0 = 0;
  for (int it = 0; it < iters + 10; ++it) {
    if (it == 10 && blockIdx.x == 0 && threadIdx.x == 0) t0 = globaltimer();
    if (MODE == 1 || MODE == 4) {
      unsigned long long *dst = acc64 + (size_t)(MODE == 4 ? (blockIdx.x % REP) : 0) * E;
      for (int i = threadIdx.x; i < E; i += NCT) {
        const long long v = (long long)((float)(i + it) * 0.001f * 4294967296.0f);
        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(dst + i), "l"(v) : "memory");
      }
    } else if (MODE == 2) {
      for (int i = threadIdx.x; i < E / 4; i += NCT) {
        const float v = (float)(i + it) * 0.001f;
        asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(acc32 + 4 * i), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
      }
    } else if (MODE == 3) {
      for (int i = threadIdx.x; i < E; i += NCT) atomicAdd(acc32 + i, (float)(i + it) * 0.001f);
    }
    target += G;
    consumer_sync();
    if (threadIdx.x == 0) {
      red_release_add(bar, 1u);
      while ((int)(ld_acquire(bar) - target) < 0) {}
    }
    consumer_sync();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = globaltimer() - t0;
}

// all CTAs read the same n-float vector from L2 after a barrier (the activation broadcast of every phase)
template <int NFLOATS>
__global__ void __launch_bounds__(288, 1) bcast_bench(unsigned *bar, const float *src, float *sink, int iters, unsigned long long *out) {
  __shared__ float4 buf[NFLOATS / 4];
  if (threadIdx.x >= NCT) return;
  const int G = gridDim.x;
  unsigned target = 0;
  unsigned long long t0 = 0;
  float acc = 0.f;
  for (int it = 0; it < iters + 10; ++it) {
    if (it == 10 && blockIdx.x == 0 && threadIdx.x == 0) t0 = globaltimer();
    for (int i = threadIdx.x; i < NFLOATS / 4; i += NCT) buf[i] = __ldcg(reinterpret_cast<const float4 *>(src) + i);
    consumer_sync();
    acc += buf[(threadIdx.x + it) % (NFLOATS / 4)].x;
    target += G;
    consumer_sync();
    if (threadIdx.x == 0) {
      red_release_add(bar, 1u);
      while ((int)(ld_acquire(bar) - target) < 0) {}
    }
    consumer_sync();
  }
  if (acc == 123.456f) sink[0] = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = globaltimer() - t0;
}

// ---- streaming ring -----------------------------------------------------------------------------
struct Pipe2 { int slot; uint32_t parity; __device__ void adv(int n) { if (++slot == n) { slot = 0; parity ^= 1u; } } };

template <int CONSUME>
__global__ void __launch_bounds__(288, 1) stream_bench(const float *buf, size_t floats_per_cta, int unit_floats, int nslot,
                                                       float *sink, unsigned *err_g, unsigned long long *out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mb[64];
  __shared__ unsigned tripped;
  if (threadIdx.x == 0) tripped = 0;
  Watchdog err{err_g, smem_u32(&tripped)};
  float *ring = reinterpret_cast<float *>(smem_raw);
  float *vec = ring + (size_t)nslot * unit_floats;
  const uint32_t full0 = smem_u32(mb), empty0 = smem_u32(mb + 32);
  if (threadIdx.x == 0) {
    for (int i = 0; i < nslot; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, NCW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < unit_floats; i += blockDim.x) vec[i] = 1.0f;
  __syncthreads();
  const float *src = buf + (size_t)blockIdx.x * floats_per_cta;
  const int n_units = (int)(floats_per_cta / unit_floats);
  Pipe2 pipe{0, 0};
  unsigned long long t0 = globaltimer();
  if (threadIdx.x >= NCT) {
    if (threadIdx.x == NCT) {
      const uint64_t pol = policy_evict_first();
      for (int u = 0; u < n_units; ++u) {
        mbar_wait(empty0 + 8 * pipe.slot, pipe.parity ^ 1u, err);
        mbar_expect_tx(full0 + 8 * pipe.slot, unit_floats * 4);
        bulk_g2s(smem_u32(ring + (size_t)pipe.slot * unit_floats), src + (size_t)u * unit_floats, unit_floats * 4,
                 full0 + 8 * pipe.slot, pol);
        pipe.adv(nslot);
      }
    }
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seg4 = unit_floats / NCW / 4;
  float acc = 0.f;
  for (int u = 0; u < n_units; ++u) {
    mbar_wait(full0 + 8 * pipe.slot, pipe.parity, err);
    if (CONSUME) {
      const float4 *w4 = reinterpret_cast<const float4 *>(ring + (size_t)pipe.slot * unit_floats) + warp * seg4;
      const float4 *v4 = reinterpret_cast<const float4 *>(vec) + warp * seg4;
      for (int i = lane; i < seg4; i += 32) {
        const float4 a = w4[i], b = v4[i];
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * pipe.slot);
    pipe.adv(nslot);
  }
  if (acc == 123.456f) sink[0] = acc;
  consumer_sync();
  if (threadIdx.x == 0) out[blockIdx.x] = globaltimer() - t0;
}

__global__ void __launch_bounds__(512) ldg_read(const float4 *buf, size_t n4, float *sink) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 a = ld_stream(buf + i), b = ld_stream(buf + i + stride), c = ld_stream(buf + i + 2 * stride), d = ld_stream(buf + i + 3 * stride);
    acc += a.x + b.y + c.z + d.w;
  }
  if (acc == 123.456f) sink[0] = acc;
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  printf("SMs %d\n", sms);
  unsigned *bar; unsigned long long *out; float *sink; unsigned *err;
  CK(cudaMalloc(&bar, 4096)); CK(cudaMalloc(&out, 8 * 256)); CK(cudaMalloc(&sink, 64)); CK(cudaMalloc(&err, 64));
  CK(cudaMemset(err, 0, 64));
  unsigned long long h[256];

  // (1) barriers
  const int iters = 2000;
  auto run_bar = [&](auto kern, const char *name) {
    CK(cudaMemset(bar, 0, 4096));
    void *args[] = {&bar, (void *)&iters, &out};
    CK(cudaLaunchCooperativeKernel((const void *)kern, dim3(sms), dim3(288), args, 0, 0));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
    printf("barrier %-34s %8.1f ns/barrier\n", name, (double)h[0] / iters);
  };
  run_bar(barrier_bench<0>, "fence+atomicAdd, ld.acquire spin");
  run_bar(barrier_bench<1>, "red.release, ld.acquire spin");
  run_bar(barrier_bench<2>, "red.release, ld.relaxed spin+fence");
  run_bar(barrier_bench<3>, "4 sub-counters");
  run_bar(barrier_bench_k<1, 0>, "k=1 lanes");

  {
    unsigned long long *acc64; float *acc32; float *src;
    CK(cudaMalloc(&acc64, 8 * 8192 * 16)); CK(cudaMalloc(&acc32, 4 * 8192)); CK(cudaMalloc(&src, 4 * 16384));
    CK(cudaMemset(acc64, 0, 8 * 8192 * 16)); CK(cudaMemset(acc32, 0, 4 * 8192)); CK(cudaMemset(src, 0, 4 * 16384));
    for (int E : {768, 1600}) {
      auto run_red = [&](auto kern, const char *name) {
        CK(cudaMemset(bar, 0, 4096));
        void *args[] = {&bar, &acc64, &acc32, &E, (void *)&iters, &out};
        CK(cudaLaunchCooperativeKernel((const void *)kern, dim3(sms), dim3(288), args, 0, 0));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
        printf("reduce E=%4d %-28s %8.1f ns/(reduce+barrier)\n", E, name, (double)h[0] / iters);
      };
      run_red(reduce_bench<0, 1>, "barrier only");
      run_red(reduce_bench<1, 1>, "red.add.u64");
      run_red(reduce_bench<2, 1>, "red.add.v4.f32");
      run_red(reduce_bench<3, 1>, "atomicAdd f32");
      run_red(reduce_bench<4, 4>, "red.add.u64 x4 replicas");
      run_red(reduce_bench<4, 16>, "red.add.u64 x16 replicas");
    }
    auto run_bc = [&](auto kern, const char *name) {
      CK(cudaMemset(bar, 0, 4096));
      void *args[] = {&bar, &src, &sink, (void *)&iters, &out};
      CK(cudaLaunchCooperativeKernel((const void *)kern, dim3(sms), dim3(288), args, 0, 0));
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
      printf("bcast %-20s %8.1f ns/(load+barrier)\n", name, (double)h[0] / iters);
    };
    run_bc(bcast_bench<768>, "768 floats");
    run_bc(bcast_bench<1536>, "1536 floats");
    run_bc(bcast_bench<3072>, "3072 floats");
  }

  // (2) stream
  const size_t total_floats = (size_t)160 * 1024 * 1024;  // 640 MB
  float *buf;
  CK(cudaMalloc(&buf, total_floats * 4));
  CK(cudaMemset(buf, 0, total_floats * 4));
  CK(cudaFuncSetAttribute(stream_bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CK(cudaFuncSetAttribute(stream_bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  const int unit_list[] = {3072};
  for (int consume = 0; consume < 2; ++consume)
    for (int unit_floats : unit_list)
      for (int nslot : {8, 16}) {
        size_t smem = ((size_t)nslot + 1) * unit_floats * 4;
        if (smem > 210 * 1024 || nslot > 32) continue;
        size_t per_cta = total_floats / sms / unit_floats * unit_floats;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        float best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
          void *args[] = {&buf, &per_cta, &unit_floats, &nslot, &sink, &err, &out};
          CK(cudaEventRecord(e0));
          if (consume) CK(cudaLaunchCooperativeKernel((const void *)stream_bench<1>, dim3(sms), dim3(288), args, smem, 0));
          else CK(cudaLaunchCooperativeKernel((const void *)stream_bench<0>, dim3(sms), dim3(288), args, smem, 0));
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
          if (ms < best) best = ms;
        }
        printf("stream consume=%d unit=%6d B nslot=%2d (%3zu KB): %7.1f GB/s\n", consume, unit_floats * 4, nslot, smem / 1024,
               (double)per_cta * sms * 4 / best / 1e6);
      }
  // (3) LDG read
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int blocks : {sms * 2, sms * 4, sms * 8}) {
      float best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        ldg_read<<<blocks, 512>>>((const float4 *)buf, total_floats / 4, sink);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
      }
      printf("ldg.128 read blocks=%d: %7.1f GB/s\n", blocks, (double)total_floats * 4 / best / 1e6);
    }
  }
  unsigned herr; CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
  printf("watchdog %u\n", herr);
  return 0;
}
