// vitssl_b200 — evaluation path: token mean-pooling of SimMIMViT.inference_forward
// (ssl/simmim/model.py:91-93) and the cosine k-nearest-neighbour classifier the reference's evaluator
// runs with scikit-learn on CPU copies of the features
// (evaluators/unsupervised_evaluator.py:38-66: KNeighborsClassifier(n_neighbors=num_classes,
// metric="cosine"), uniform vote). Everything is fp32 on CUDA cores on purpose: neighbour ranks must
// not move with bf16 rounding, and the work is tiny (8 000 x 5 000 x 384 for STL10).
#include "common.cuh"
#include "vitssl_b200.h"

namespace vitssl {
namespace {

// out[b, :] = mean over s of x[b, s, :]     (x fp32 [B, S, D] dense)
__global__ void __launch_bounds__(128) mean_tokens_kernel(const float* __restrict__ x, float* __restrict__ out, int S,
                                                          int D) {
  const int b = blockIdx.y;
  const int d = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (d >= D) return;
  const float* p = x + (static_cast<long long>(b) * S) * D + d;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < S; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(p + static_cast<long long>(s) * D);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const float inv = 1.0f / S;
  *reinterpret_cast<float4*>(out + static_cast<long long>(b) * D + d) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
}

// y[r, :] = x[r, :] / max(||x[r, :]||, 1e-12): one warp per row
__global__ void __launch_bounds__(128) normalize_rows_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                             long long rows, int D) {
  const int lane = threadIdx.x & 31;
  const long long r = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) { const float v = x[r * D + c]; ss = fmaf(v, v, ss); }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  for (int c = lane; c < D; c += 32) y[r * D + c] = x[r * D + c] * inv;
}

// sims[v, t] = <val[v, :], train[t, :]>  (rows already normalised): 32 x 32 output tile per CTA of
// 256 threads, 32-wide k-slices staged in shared memory, 4 outputs per thread
__global__ void __launch_bounds__(256) sim_matrix_kernel(const float* __restrict__ val, const float* __restrict__ train,
                                                         float* __restrict__ sims, int Nv, int Nt, int D) {
  __shared__ float sv[32][33], st[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty: 0..7
  const int v0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < D; k0 += 32) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = ty + 8 * j, k = k0 + tx;
      sv[row][tx] = (v0 + row < Nv && k < D) ? val[static_cast<long long>(v0 + row) * D + k] : 0.f;
      st[row][tx] = (t0 + row < Nt && k < D) ? train[static_cast<long long>(t0 + row) * D + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float b = st[tx][k];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(sv[ty + 8 * j][k], b, acc[j]);
    }
    __syncthreads();
  }
  if (t0 + tx < Nt) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (v0 + ty + 8 * j < Nv) sims[static_cast<long long>(v0 + ty + 8 * j) * Nt + t0 + tx] = acc[j];
  }
}

// One warp per validation row: k rounds of (max similarity, smallest index on ties) with removal,
// then a uniform vote over the k labels (most frequent; smallest label on ties — scipy's mode, which
// KNeighborsClassifier.predict uses). sims is consumed (selected entries are overwritten).
__global__ void __launch_bounds__(128) knn_vote_kernel(float* __restrict__ sims, const int* __restrict__ labels,
                                                       int* __restrict__ pred, int* __restrict__ nbr_out, int Nv, int Nt,
                                                       int k, int num_classes) {
  extern __shared__ int votes[];  // [4 warps][num_classes]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int v = blockIdx.x * 4 + w;
  int* my = votes + w * num_classes;
  for (int c = lane; c < num_classes; c += 32) my[c] = 0;
  __syncwarp();
  if (v >= Nv) return;
  float* row = sims + static_cast<long long>(v) * Nt;
  for (int r = 0; r < k; ++r) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int t = lane; t < Nt; t += 32) {
      const float s = row[t];
      if (s > best) { best = s; bi = t; }  // ascending t per lane: first (smallest) index kept on ties
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
      row[bi] = -INFINITY;
      const int lab = labels[bi];
      if (lab >= 0 && lab < num_classes) my[lab] += 1;
      if (nbr_out) nbr_out[static_cast<long long>(v) * k + r] = bi;
    }
    __syncwarp();
  }
  int bc = -1, bl = 0x7fffffff;
  for (int c = lane; c < num_classes; c += 32) {
    const int n = my[c];
    if (n > bc) { bc = n; bl = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int oc = __shfl_xor_sync(0xffffffffu, bc, o), ol = __shfl_xor_sync(0xffffffffu, bl, o);
    if (oc > bc || (oc == bc && ol < bl)) { bc = oc; bl = ol; }
  }
  if (lane == 0) pred[v] = bl;
}

}  // namespace
}  // namespace vitssl

using namespace vitssl;

extern "C" int vitssl_mean_tokens_f32(const float* x, float* out, int64_t B, int64_t S, int64_t D, cudaStream_t stream) {
  VITSSL_REQUIRE(x && out && B > 0 && S > 0 && D > 0, VITSSL_ERR_ARG, "mean_tokens_f32: bad args");
  VITSSL_REQUIRE(D % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                 VITSSL_ERR_SHAPE, "mean_tokens_f32: D %% 4 and 16-byte alignment required");
  dim3 grid((unsigned)((D / 4 + 127) / 128), (unsigned)B);
  mean_tokens_kernel<<<grid, 128, 0, stream>>>(x, out, (int)S, (int)D);
  return check_launch("mean_tokens_f32");
}

extern "C" int vitssl_knn_cosine(const float* val, const float* train, const int32_t* train_labels, float* val_n,
                                 float* train_n, float* sims, int32_t* pred, int32_t* neighbors, int64_t Nv,
                                 int64_t Nt, int64_t D, int64_t k, int64_t num_classes, cudaStream_t stream) {
  VITSSL_REQUIRE(val && train && train_labels && val_n && train_n && sims && pred, VITSSL_ERR_ARG, "knn_cosine: null pointer");
  VITSSL_REQUIRE(Nv > 0 && Nt > 0 && D > 0 && k > 0 && k <= Nt && num_classes > 0 && num_classes <= 8192, VITSSL_ERR_SHAPE,
                 "knn_cosine: bad sizes (Nv=%lld Nt=%lld D=%lld k=%lld classes=%lld)", (long long)Nv, (long long)Nt,
                 (long long)D, (long long)k, (long long)num_classes);
  normalize_rows_kernel<<<(unsigned)((Nv + 3) / 4), 128, 0, stream>>>(val, val_n, Nv, (int)D);
  int rc = check_launch("knn_cosine/normalize");
  if (rc) return rc;
  normalize_rows_kernel<<<(unsigned)((Nt + 3) / 4), 128, 0, stream>>>(train, train_n, Nt, (int)D);
  if ((rc = check_launch("knn_cosine/normalize"))) return rc;
  dim3 grid((unsigned)((Nt + 31) / 32), (unsigned)((Nv + 31) / 32));
  sim_matrix_kernel<<<grid, 256, 0, stream>>>(val_n, train_n, sims, (int)Nv, (int)Nt, (int)D);
  if ((rc = check_launch("knn_cosine/similarity"))) return rc;
  const size_t smem = 4 * (size_t)num_classes * sizeof(int);
  if (smem > 48 * 1024) cudaFuncSetAttribute(knn_vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  knn_vote_kernel<<<(unsigned)((Nv + 3) / 4), 128, smem, stream>>>(sims, train_labels, pred, neighbors, (int)Nv, (int)Nt,
                                                                   (int)k, (int)num_classes);
  return check_launch("knn_cosine/vote");
}
