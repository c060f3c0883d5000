// Developer microbenchmark (GPU box): MUFU.EX2 cadence per warp and per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_rate scripts/micro/mufu_rate.cu && /tmp/mufu_rate
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int ILP, int FMA_PER>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[ILP], acc[8];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = -1.0f - threadIdx.x * 1e-3f - i;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      v[i] = ex2(v[i]);
#pragma unroll
      for (int f = 0; f < FMA_PER; ++f) acc[(i * FMA_PER + f) & 7] = fmaf(acc[(i * FMA_PER + f) & 7], 1.0001f, 0.5f);
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int FMA_PER>
void run(int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<ILP, FMA_PER><<<148, warps * 32>>>(out, cyc, iters);
  k<ILP, FMA_PER><<<148, warps * 32>>>(out, cyc, iters);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("warps/block=%2d (per SMSP %.1f) ILP=%2d fma/mufu=%d: %.2f cycles per MUFU per warp, %.2f per SMSP-MUFU\n", warps, warps / 4.0, ILP, FMA_PER,
         (double)h / (iters * ILP), (double)h / (iters * ILP) / (warps / 4.0 < 1 ? 1 : warps / 4.0));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<16, 0>(4); run<16, 0>(8); run<16, 0>(16); run<16, 0>(32);
  run<16, 2>(4); run<16, 4>(4); run<16, 6>(4); run<16, 8>(4);
  run<16, 4>(8); run<16, 6>(8);
  return 0;
}
