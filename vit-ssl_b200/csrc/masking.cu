// vitssl_b200 — SimMIM mask generation in ONE launch, bit-exact with the reference.
//
// The reference draws the mask of sample b as `torch.randperm(N, device=cuda)[:n_m]`, B times in
// sequence from the device's default generator (ssl/simmim/masking.py:22-25), then scatters the
// indices into a bool mask (:27-33) and gathers `patches[bool_mask]` (:35). On CUDA one
// torch.randperm(n) (n small enough that `bits` <= 32) is the following integer algorithm
// (ATen/native/cuda/Randperm.cu + Randperm.cuh + DistributionTemplates.h, torch 2.11):
//   1. keys: thread i of a 256-thread block does curand_init(seed, /*subsequence*/ i, offset) and
//      one curand4(); key_i = int32((((x << 32) | y) % (2^32 - 1)) + INT_MIN). The generator
//      offset advances by 4.
//   2. stable LSD radix sort of (key, i) pairs on key bits [0, bits),
//      bits = ceil(log2(n - (6 n^2 + 1) / (12 ln 0.9))).
//   3. "islands" of equal masked keys are re-shuffled by their first element's thread with a
//      serial Fisher-Yates driven by curand_init(seed, /*subsequence*/ pos, offset + 4) and
//      successive curand() words. The generator offset advances by n rounded up to 4.
// This kernel replays exactly that with one CTA per sample (sample b starts at generator offset
// offset0 + b * (4 + roundup4(N))) and derives the mask tables in the same pass: the bool mask,
// the flat ids of the masked patches in ascending (b, n) order (== order of x[bool_mask]) and
// the inverse map. The caller advances the torch generator by B * (4 + roundup4(N)).
#include "common.cuh"
#include "vitssl_b200.h"

namespace vitssl {
namespace {

constexpr int MASK_THREADS = 256;
constexpr int MASK_MAX_N = 1024;

// curand Philox4_32_10 stream addressing: counter = (offset / 4 [64 bit], subsequence [64 bit])
__device__ __forceinline__ uint4 curand_philox_block(uint64_t seed, uint64_t subsequence,
                                                     uint64_t offset_div4) {
  return philox4x32(seed, /*c2,c3=*/subsequence, /*c0,c1=*/offset_div4);
}

__global__ void __launch_bounds__(MASK_THREADS)
simmim_mask_kernel(long long* __restrict__ perm_out, uint8_t* __restrict__ bool_mask,
                   int* __restrict__ rows, int* __restrict__ inv, int N, int n_keep, int bits,
                   unsigned long long seed, unsigned long long offset0,
                   unsigned long long per_sample) {
  __shared__ uint32_t key[MASK_MAX_N];
  __shared__ uint32_t skey[MASK_MAX_N];
  __shared__ int data[MASK_MAX_N];
  __shared__ uint8_t mark[MASK_MAX_N];
  __shared__ int warp_cnt[MASK_THREADS / 32];
  __shared__ int running;

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const unsigned long long offset = offset0 + static_cast<unsigned long long>(b) * per_sample;
  const uint32_t kmask = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);

  // 1. keys
  for (int i = tid; i < N; i += MASK_THREADS) {
    const uint4 r = curand_philox_block(seed, static_cast<uint64_t>(i), offset >> 2);
    const unsigned long long v = (static_cast<unsigned long long>(r.x) << 32) | r.y;
    const uint32_t k = static_cast<uint32_t>(v % 0xffffffffull);
    // + INT_MIN only flips bit 31, which the signed radix sort flips back: order is by k & kmask
    key[i] = k & kmask;
    mark[i] = 0;
  }
  __syncthreads();
  // 2. stable sort by rank counting (N <= 1024: N^2 comparisons per sample are negligible)
  for (int i = tid; i < N; i += MASK_THREADS) {
    const uint32_t ki = key[i];
    int rank = 0;
    for (int j = 0; j < N; ++j) {
      const uint32_t kj = key[j];
      rank += (kj < ki) || (kj == ki && j < i);
    }
    skey[rank] = ki;
    data[rank] = i;
  }
  __syncthreads();
  // 3. islands of duplicate keys: serial Fisher-Yates by the island leader
  for (int pos = tid; pos < N - 1; pos += MASK_THREADS) {
    if (skey[pos] != skey[pos + 1]) continue;
    if (pos != 0 && skey[pos] == skey[pos - 1]) continue;
    int island = 0;
    do { ++island; } while (pos + island < N && skey[pos + island] == skey[pos]);
    uint64_t ctr = (offset + 4) >> 2;
    uint4 blk = curand_philox_block(seed, static_cast<uint64_t>(pos), ctr);
    int word = 0;
    for (int i = island - 1; i > 0; --i) {
      const uint32_t w = word == 0 ? blk.x : word == 1 ? blk.y : word == 2 ? blk.z : blk.w;
      if (++word == 4) {
        word = 0;
        blk = curand_philox_block(seed, static_cast<uint64_t>(pos), ++ctr);
      }
      const int r = static_cast<int>(w % static_cast<uint32_t>(i + 1));
      if (i != r) {
        const int tmp = data[pos + i];
        data[pos + i] = data[pos + r];
        data[pos + r] = tmp;
      }
    }
  }
  __syncthreads();
  // 4. first n_keep entries of the permutation are the masked patches
  for (int j = tid; j < n_keep; j += MASK_THREADS) {
    const int n = data[j];
    if (perm_out) perm_out[static_cast<long long>(b) * n_keep + j] = n;
    mark[n] = 1;
  }
  if (tid == 0) running = 0;
  __syncthreads();
  // 5. ascending order of masked ids: block-wide exclusive scan of the marks
  for (int base = 0; base < N; base += MASK_THREADS) {
    const int n = base + tid;
    const int m = (n < N) ? mark[n] : 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, m);
    const int lane = tid & 31, warp = tid >> 5;
    if (lane == 0) warp_cnt[warp] = __popc(ballot);
    __syncthreads();
    int before = running;
    for (int w = 0; w < warp; ++w) before += warp_cnt[w];
    before += __popc(ballot & ((1u << lane) - 1u));
    if (n < N) {
      const long long flat = static_cast<long long>(b) * N + n;
      bool_mask[flat] = static_cast<uint8_t>(m);
      if (m) {
        rows[static_cast<long long>(b) * n_keep + before] = static_cast<int>(flat);
        inv[flat] = b * n_keep + before;
      } else {
        inv[flat] = -1;
      }
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < MASK_THREADS / 32; ++w) tot += warp_cnt[w];
      running += tot;
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace vitssl

using namespace vitssl;

extern "C" int vitssl_randperm_bits(int64_t n) {
  if (n <= 0) return 0;
  const double log_threshold_12 = log(0.9) * 12.0;
  const double nd = static_cast<double>(n);
  const int bits = static_cast<int>(ceil(log2(nd - (6.0 * nd * nd + 1.0) / log_threshold_12)));
  return bits < 64 ? bits : 64;
}

extern "C" int64_t vitssl_randperm_offset_per_call(int64_t n) {
  // key generation consumes 4 (one curand4 per thread while n <= 256 * grid), the duplicate-key
  // pass consumes n rounded up to a multiple of 4 (CUDAGeneratorImpl::philox_cuda_state)
  return n <= 0 ? 0 : 4 + ((n + 3) / 4) * 4;
}

extern "C" int vitssl_simmim_mask(int64_t* perm_out, uint8_t* bool_mask, int32_t* rows,
                                  int32_t* inv, int64_t B, int64_t N, int64_t n_keep,
                                  uint64_t philox_seed, uint64_t philox_offset,
                                  cudaStream_t stream) {
  VITSSL_REQUIRE(bool_mask && inv && (rows || n_keep == 0), VITSSL_ERR_ARG, "simmim_mask: null output");
  VITSSL_REQUIRE(B > 0 && N > 0 && n_keep >= 0 && n_keep <= N, VITSSL_ERR_SHAPE,
                 "simmim_mask: bad sizes B=%lld N=%lld n_keep=%lld", (long long)B, (long long)N,
                 (long long)n_keep);
  VITSSL_REQUIRE(N <= MASK_MAX_N, VITSSL_ERR_SHAPE, "simmim_mask: N=%lld exceeds %d", (long long)N,
                 MASK_MAX_N);
  VITSSL_REQUIRE(B * N < (1ll << 31), VITSSL_ERR_SHAPE, "simmim_mask: B*N exceeds int32");
  VITSSL_REQUIRE(philox_offset % 4 == 0, VITSSL_ERR_ARG, "simmim_mask: offset must be a multiple of 4");
  const int bits = vitssl_randperm_bits(N);
  VITSSL_REQUIRE(bits <= 32, VITSSL_ERR_SHAPE, "simmim_mask: N=%lld needs 64-bit keys", (long long)N);
  simmim_mask_kernel<<<static_cast<unsigned>(B), MASK_THREADS, 0, stream>>>(
      reinterpret_cast<long long*>(perm_out), bool_mask, rows, inv, static_cast<int>(N),
      static_cast<int>(n_keep), bits, philox_seed, philox_offset,
      static_cast<unsigned long long>(vitssl_randperm_offset_per_call(N)));
  return check_launch("simmim_mask");
}
