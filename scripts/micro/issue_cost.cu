// Developer microbenchmark (GPU box): issue cost (cycles per warp instruction, one warp per SM
// sub-partition, 8 independent chains) of the instructions the softmax / epilogue loops are made of.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/issue_cost.bin scripts/micro/issue_cost.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
enum Op { F2FP, FMNMX, FFMA2, FADD2, PRMT, IADD, FMUL, MUFU, F2FP_MUFU, RND_PRMT };
template <int OP>
__global__ void k(uint32_t* out, long long* cyc, int iters) {
  uint32_t a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = 0x3f800000u + threadIdx.x * 977 + i * 131; b[i] = 0x3f000000u + i * 77 + threadIdx.x; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == F2FP) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(a[i]) : "f"(__uint_as_float(a[i])), "f"(__uint_as_float(b[i])));
        if (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(*(float*)&a[i]) : "f"(__uint_as_float(b[i])));
        if (OP == FFMA2) { unsigned long long x = ((unsigned long long)a[i] << 32) | b[i], y; asm volatile("fma.rn.f32x2 %0, %1, %1, %1;" : "=l"(y) : "l"(x)); a[i] = (uint32_t)(y >> 32); b[i] = (uint32_t)y; }
        if (OP == FADD2) { unsigned long long x = ((unsigned long long)a[i] << 32) | b[i], y; asm volatile("add.rn.f32x2 %0, %1, %1;" : "=l"(y) : "l"(x)); a[i] = (uint32_t)(y >> 32); b[i] = (uint32_t)y; }
        if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(a[i]) : "r"(b[i]));
        if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
        if (OP == FMUL) asm volatile("mul.f32 %0, %0, %1;" : "+f"(*(float*)&a[i]) : "f"(__uint_as_float(b[i])));
        if (OP == MUFU) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(*(float*)&a[i]));
        if (OP == F2FP_MUFU) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(*(float*)&b[i])); asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(a[i]) : "f"(__uint_as_float(a[i])), "f"(__uint_as_float(b[i]))); }
        if (OP == RND_PRMT) { asm volatile("add.u32 %0, %0, 0x8000;" : "+r"(a[i])); asm volatile("add.u32 %0, %0, 0x8000;" : "+r"(b[i])); asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(a[i]) : "r"(b[i])); }
      }
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name, int warps) {
  uint32_t* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 1000;
  k<OP><<<148, warps * 32>>>(out, cyc, iters);
  k<OP><<<148, warps * 32>>>(out, cyc, iters);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps/SMSP=%d: %.2f cycles per loop body op per warp\n", name, warps / 4, (double)h / (iters * 32));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8}) {
    run<F2FP>("cvt.rn.bf16x2.f32 (F2FP)", w); run<FMNMX>("max.f32", w); run<FFMA2>("fma.f32x2", w); run<FADD2>("add.f32x2", w);
    run<PRMT>("prmt", w); run<IADD>("add.u32", w); run<FMUL>("mul.f32", w); run<MUFU>("ex2.approx", w);
    run<F2FP_MUFU>("ex2 + cvt pack", w); run<RND_PRMT>("2 x add 0x8000 + prmt", w);
  }
  return 0;
}
