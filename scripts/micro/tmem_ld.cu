// Developer microbenchmark (GPU box): tcgen05.ld latency / tcgen05.wait::ld cost for different shapes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/tmem_ld.bin scripts/micro/tmem_ld.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define LD16(addr, r) asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
  : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(addr) : "memory")
#define WAITLD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
// MODE 0: N back-to-back x16 loads then one wait, result consumed   (latency + throughput of N*2 KB)
// MODE 1: one x16 load, SPIN independent FMAs, then wait            (does the wait cost anything once data is there?)
template <int N, int SPIN>
__global__ void k(uint32_t* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot + ((uint32_t)(warp & 3) * 32 << 16);
  uint32_t r[N][16]; uint32_t acc = 0; float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = i + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int n = 0; n < N; ++n) LD16(base + ((n * 16 + (acc & 1)) & 255), r[n]);
#pragma unroll
    for (int s = 0; s < SPIN; ++s) f[s & 7] = fmaf(f[s & 7], 1.0001f, 0.5f);
    WAITLD();
#pragma unroll
    for (int n = 0; n < N; ++n) acc += r[n][0] & 1;   // next addresses depend on the loaded data
  }
  long long t1 = clock64();
  float fs = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) fs += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (uint32_t)fs;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}
template <int N, int SPIN>
void run(int warps) {
  uint32_t* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 500;
  k<N, SPIN><<<148, warps * 32>>>(out, cyc, iters);
  k<N, SPIN><<<148, warps * 32>>>(out, cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("warps=%d  %d x (x16 = 2 KB/warp) per wait, %3d independent FMAs before the wait: %.1f cycles per iteration (%s)\n", warps, N, SPIN,
         (double)h / iters, cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1, 0>(1); run<2, 0>(1); run<4, 0>(1); run<8, 0>(1);
  run<1, 0>(4); run<2, 0>(4); run<4, 0>(4); run<8, 0>(4);
  run<1, 64>(4); run<1, 128>(4); run<1, 256>(4); run<2, 256>(4); run<4, 256>(4);
  run<2, 0>(8); run<4, 0>(8);
  return 0;
}
