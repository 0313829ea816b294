// microbench2.cu -- cycle cost of the per-unit work of the decode kernel on one SM:
//   dot of a 12 KB ring unit against the activation vector (4-row and 1-row forms), with 1..8 warps active,
//   and the 4 interleaved warp-shuffle reductions.
#include <cstdio>
#include <cstdlib>
#include "../zig_gpt2_b200/csrc/zg_common.cuh"
using namespace zg;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

template <int RPS>
__global__ void __launch_bounds__(256, 1) dot_bench(int nw, int E, int iters, float *sink, long long *out) {
  extern __shared__ __align__(128) float sm[];
  const int slotf = 4 * E;
  float *ring = sm;             // 8 units
  float *vec = sm + 8 * slotf;  // 4E
  for (int i = threadIdx.x; i < 9 * slotf; i += blockDim.x) sm[i] = 1.0f / (1 + (i & 7));
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= nw) return;
  const int K = (RPS == 4) ? E : 4 * E;
  const int k4 = K >> 2;
  const float4 *vec4 = reinterpret_cast<const float4 *>(vec);
  const float4 *w4 = reinterpret_cast<const float4 *>(ring + (size_t)warp * slotf);
  float tot = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    if (RPS == 4) {
#pragma unroll 2
      for (int i = lane; i < k4; i += 32) {
        const float4 xv = vec4[i];
        const float4 w0 = w4[i], w1 = w4[k4 + i], w2 = w4[2 * k4 + i], w3 = w4[3 * k4 + i];
        a0 = fmaf(w0.x, xv.x, a0); a0 = fmaf(w0.y, xv.y, a0); a0 = fmaf(w0.z, xv.z, a0); a0 = fmaf(w0.w, xv.w, a0);
        a1 = fmaf(w1.x, xv.x, a1); a1 = fmaf(w1.y, xv.y, a1); a1 = fmaf(w1.z, xv.z, a1); a1 = fmaf(w1.w, xv.w, a1);
        a2 = fmaf(w2.x, xv.x, a2); a2 = fmaf(w2.y, xv.y, a2); a2 = fmaf(w2.z, xv.z, a2); a2 = fmaf(w2.w, xv.w, a2);
        a3 = fmaf(w3.x, xv.x, a3); a3 = fmaf(w3.y, xv.y, a3); a3 = fmaf(w3.z, xv.z, a3); a3 = fmaf(w3.w, xv.w, a3);
      }
    } else {
#pragma unroll 4
      for (int i = lane; i < k4; i += 32) {
        const float4 xv = vec4[i];
        const float4 w0 = w4[i];
        a0 = fmaf(w0.x, xv.x, a0); a1 = fmaf(w0.y, xv.y, a1); a2 = fmaf(w0.z, xv.z, a2); a3 = fmaf(w0.w, xv.w, a3);
      }
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
    tot += a0 + a1 + a2 + a3;
    asm volatile("" ::: "memory");
  }
  long long t1 = clock64();
  if (lane == 0) out[warp] = t1 - t0;
  if (tot == 123.0f) sink[0] = tot;
}

int main() {
  CK(cudaSetDevice(0));
  float *sink; long long *out; long long h[8];
  CK(cudaMalloc(&sink, 64)); CK(cudaMalloc(&out, 64));
  const int E = 768, iters = 2000;
  const size_t smem = 9 * 4 * E * sizeof(float);
  CK(cudaFuncSetAttribute(dot_bench<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(dot_bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rps : {4, 1})
    for (int nw : {1, 2, 4, 6, 8}) {
      if (rps == 4) dot_bench<4><<<1, 256, smem>>>(nw, E, iters, sink, out);
      else dot_bench<1><<<1, 256, smem>>>(nw, E, iters, sink, out);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost));
      printf("dot+4 shuffles rps=%d warps=%d: %.0f cycles per 12KB unit (warp 0)\n", rps, nw, (double)h[0] / iters);
    }
  return 0;
}
