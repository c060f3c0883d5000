// vitssl_b200 — bandwidth-bound helper kernels around the GEMMs: weight casts, EMA teacher
// update, bias-gradient column sums, patch extraction, token assembly (CLS / positional
// embedding / SimMIM mask-token substitution) and the masked-row gathers. All are single-pass,
// 16-byte vectorised and sized so that consecutive threads touch consecutive addresses.
#include "common.cuh"
#include "vitssl_b200.h"

namespace vitssl {
namespace {

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------
// multi-tensor apply: up to MT_MAX tensors per launch, 2048 elements per CTA
// ---------------------------------------------------------------------------------------
constexpr int MT_MAX = 48;
constexpr int MT_CHUNK = 2048;
struct MultiArgs {
  void* a[MT_MAX];
  void* b[MT_MAX];
  long long n[MT_MAX];
  int block_start[MT_MAX + 1];
  int count;
};

__device__ __forceinline__ int mt_find(const MultiArgs& m, int block) {
  int t = 0;
  while (t + 1 < m.count && m.block_start[t + 1] <= block) ++t;
  return t;
}

// a: fp32 source, b: bf16 destination
__global__ void __launch_bounds__(256) multi_cast_kernel(const __grid_constant__ MultiArgs m) {
  const int t = mt_find(m, blockIdx.x);
  const long long base = static_cast<long long>(blockIdx.x - m.block_start[t]) * MT_CHUNK;
  const float* src = reinterpret_cast<const float*>(m.a[t]);
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(m.b[t]);
  const long long n = m.n[t];
  const long long i = base + threadIdx.x * 8;
  if (i + 8 <= n && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
      ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    const float4 v0 = *reinterpret_cast<const float4*>(src + i);
    const float4 v1 = *reinterpret_cast<const float4*>(src + i + 4);
    uint4 o;
    o.x = pack_bf16(v0.x, v0.y); o.y = pack_bf16(v0.z, v0.w);
    o.z = pack_bf16(v1.x, v1.y); o.w = pack_bf16(v1.z, v1.w);
    *reinterpret_cast<uint4*>(dst + i) = o;
  } else {
    for (long long j = i; j < n && j < i + 8; ++j) dst[j] = __float2bfloat16_rn(src[j]);
  }
}

// a: teacher (in/out), b: student.  t <- m t + (1 - m) s   (ssl/dino/model.py:126-139)
__global__ void __launch_bounds__(256) multi_ema_kernel(const __grid_constant__ MultiArgs m,
                                                        const float mom) {
  const int t = mt_find(m, blockIdx.x);
  const long long base = static_cast<long long>(blockIdx.x - m.block_start[t]) * MT_CHUNK;
  float* te = reinterpret_cast<float*>(m.a[t]);
  const float* st = reinterpret_cast<const float*>(m.b[t]);
  const long long n = m.n[t];
  const float om = 1.0f - mom;
  const bool vec = ((reinterpret_cast<uintptr_t>(te) & 15) == 0) && ((reinterpret_cast<uintptr_t>(st) & 15) == 0);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const long long i = base + u * 1024 + threadIdx.x * 4;
    if (i + 4 <= n && vec) {
      float4 a = *reinterpret_cast<float4*>(te + i);
      const float4 s = *reinterpret_cast<const float4*>(st + i);
      // same operation order as param.mul_(m).add_((1 - m) * student)
      a.x = a.x * mom + om * s.x; a.y = a.y * mom + om * s.y;
      a.z = a.z * mom + om * s.z; a.w = a.w * mom + om * s.w;
      *reinterpret_cast<float4*>(te + i) = a;
    } else {
      for (long long j = i; j < n && j < i + 4; ++j) te[j] = te[j] * mom + om * st[j];
    }
  }
}

template <typename F>
int multi_launch(const void* const* a, const void* const* b, const int64_t* n, int count, F launch) {
  int done = 0;
  while (done < count) {
    MultiArgs m{};
    int blocks = 0, k = 0;
    while (done + k < count && k < MT_MAX) {
      m.a[k] = const_cast<void*>(a[done + k]);
      m.b[k] = const_cast<void*>(b[done + k]);
      m.n[k] = n[done + k];
      m.block_start[k] = blocks;
      blocks += (int)((n[done + k] + MT_CHUNK - 1) / MT_CHUNK);
      ++k;
    }
    m.block_start[k] = blocks;
    m.count = k;
    if (blocks > 0) {
      int rc = launch(m, blocks);
      if (rc) return rc;
    }
    done += k;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// column sums of a bf16 matrix (bias gradients): out[c] = sum_r x[r, c]
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, long long ld,
                                                     long long rows, int cols, float* __restrict__ out,
                                                     int rows_per_cta) {
  __shared__ float red[8][256 + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < cols) {
    for (long long r = r0 + warp; r < r1; r += 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(x + r * ld + c);
      acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
      acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int cc = blockIdx.x * 256 + threadIdx.x;
  if (cc < cols) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    atomicAdd(out + cc, s);
  }
}

// dense matrices (ld == cols) whose 16-byte column-group count g = cols / 8 divides the block size:
// thread t owns column group t % g for the whole kernel and walks the rows t / g, t / g + k, ...
// with four independent 16-byte loads in flight; one smem reduction + one atomic per column per CTA.
__global__ void __launch_bounds__(288) colsum_dense_kernel(const __nv_bfloat16* __restrict__ x, long long rows,
                                                           int g, float* __restrict__ out, int rows_per_cta) {
  extern __shared__ float red[];  // [blockDim.x][8]
  pdl_launch_dependents();
  pdl_wait();
  const int k = blockDim.x / g;   // rows per pass
  const int cg = threadIdx.x % g, rl = threadIdx.x / g;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  const uint4* base = reinterpret_cast<const uint4*>(x) + cg;
  f32x2 acc[4] = {pk2(0.f, 0.f), pk2(0.f, 0.f), pk2(0.f, 0.f), pk2(0.f, 0.f)};
  auto add = [&](const uint4& v) {
    acc[0] = fadd2(acc[0], pk2(bf16_lo(v.x), bf16_hi(v.x)));
    acc[1] = fadd2(acc[1], pk2(bf16_lo(v.y), bf16_hi(v.y)));
    acc[2] = fadd2(acc[2], pk2(bf16_lo(v.z), bf16_hi(v.z)));
    acc[3] = fadd2(acc[3], pk2(bf16_lo(v.w), bf16_hi(v.w)));
  };
  long long r = r0 + rl;
  for (; r + 3ll * k < r1; r += 4ll * k) {
    const uint4 v0 = __ldg(base + r * g), v1 = __ldg(base + (r + k) * g);
    const uint4 v2 = __ldg(base + (r + 2ll * k) * g), v3 = __ldg(base + (r + 3ll * k) * g);
    add(v0); add(v1); add(v2); add(v3);
  }
  for (; r < r1; r += k) add(__ldg(base + r * g));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float lo, hi;
    upk2(acc[i], lo, hi);
    red[threadIdx.x * 8 + 2 * i] = lo;
    red[threadIdx.x * 8 + 2 * i + 1] = hi;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < g * 8; c += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < k; ++j) s += red[(j * g + (c >> 3)) * 8 + (c & 7)];
    atomicAdd(out + c, s);
  }
}

__global__ void colsum_generic_kernel(const __nv_bfloat16* __restrict__ x, long long ld, long long rows,
                                      int cols, float* __restrict__ out, int rows_per_cta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  float s = 0.f;
  for (long long r = r0; r < r1; ++r) s += __bfloat162float(x[r * ld + c]);
  atomicAdd(out + c, s);
}

// ---------------------------------------------------------------------------------------
// im2col + cast: img fp32 [B,C,H,W] -> patches bf16 [B*gh*gw, C*p*p], feature order (c,ph,pw)
// (nn.Unfold, ssl/simmim/model.py:43; equals the Conv2d(k=s=p) im2col, patch_embedding.py:22,79)
// ---------------------------------------------------------------------------------------
// uint8 sources are raw image bytes; value = byte / 255 (torchvision ToTensor), §8(f)3
__device__ __forceinline__ float u8_to_unit(uint32_t b) { return __fdiv_rn(static_cast<float>(b), 255.0f); }

template <int VEC, typename SRC>
__global__ void __launch_bounds__(256) im2col_kernel(const SRC* __restrict__ img,
                                                     __nv_bfloat16* __restrict__ out, int B, int C, int H,
                                                     int W, int p, long long total_groups) {
  const long long g = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (g >= total_groups) return;
  const int gw = W / p, gh = H / p, P = C * p * p;
  const long long e = g * VEC;
  const int f = (int)(e % P);
  const long long patch = e / P;
  const int gx = (int)(patch % gw), gy = (int)((patch / gw) % gh), b = (int)(patch / (static_cast<long long>(gw) * gh));
  const int c = f / (p * p), ph = (f / p) % p, pw = f % p;
  const SRC* src = img + ((static_cast<long long>(b) * C + c) * H + gy * p + ph) * W + gx * p + pw;
  if constexpr (VEC == 8) {
    float v[8];
    if constexpr (sizeof(SRC) == 1) {
      const uint2 q = *reinterpret_cast<const uint2*>(src);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = u8_to_unit((q.x >> (8 * i)) & 0xffu);
        v[4 + i] = u8_to_unit((q.y >> (8 * i)) & 0xffu);
      }
    } else {
      const float4 v0 = *reinterpret_cast<const float4*>(src);
      const float4 v1 = *reinterpret_cast<const float4*>(src + 4);
      v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w; v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
    }
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
    o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + e) = o;
  } else {
    if constexpr (sizeof(SRC) == 1) out[e] = __float2bfloat16_rn(u8_to_unit(*src));
    else out[e] = __float2bfloat16_rn(*src);
  }
}

// gather of raw fp32 patches for the rows listed in idx (SimMIM targets, masking.py:35)
template <int VEC, typename SRC>
__global__ void __launch_bounds__(256) gather_patches_kernel(const SRC* __restrict__ img,
                                                             const int* __restrict__ idx,
                                                             float* __restrict__ out, int C, int H, int W,
                                                             int p, long long total_groups) {
  const long long g = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (g >= total_groups) return;
  const int gw = W / p, gh = H / p, P = C * p * p;
  const long long e = g * VEC;
  const int f = (int)(e % P);
  const long long patch = idx[e / P];
  const int gx = (int)(patch % gw), gy = (int)((patch / gw) % gh), b = (int)(patch / (static_cast<long long>(gw) * gh));
  const int c = f / (p * p), ph = (f / p) % p, pw = f % p;
  const SRC* src = img + ((static_cast<long long>(b) * C + c) * H + gy * p + ph) * W + gx * p + pw;
  if constexpr (VEC == 4) {
    if constexpr (sizeof(SRC) == 1) {
      const uint32_t q = *reinterpret_cast<const uint32_t*>(src);
      *reinterpret_cast<float4*>(out + e) = make_float4(u8_to_unit(q & 0xffu), u8_to_unit((q >> 8) & 0xffu),
                                                        u8_to_unit((q >> 16) & 0xffu), u8_to_unit(q >> 24));
    } else {
      *reinterpret_cast<float4*>(out + e) = *reinterpret_cast<const float4*>(src);
    }
  } else {
    if constexpr (sizeof(SRC) == 1) out[e] = u8_to_unit(*src);
    else out[e] = *src;
  }
}

// ---------------------------------------------------------------------------------------
// sparse row interpolation: dst[i,:] = sum_t w[i,t] * src[idx[i,t],:]  (bicubic resize of the
// positional-embedding grid, patch_embedding.py:26-48, as a 16-tap table built once on the host).
// Backward scatters: dsrc[idx[i,t],:] += w[i,t] * ddst[i,:]  (dsrc zeroed by the entry point).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) interp_rows_fwd_kernel(const float* __restrict__ src, const int* __restrict__ idx,
                                                              const float* __restrict__ w, float* __restrict__ dst,
                                                              int n_out, int D, int taps) {
  const int i = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += 128) {
    float acc = 0.f;
    for (int t = 0; t < taps; ++t) acc = fmaf(w[i * taps + t], src[static_cast<long long>(idx[i * taps + t]) * D + d], acc);
    dst[static_cast<long long>(i) * D + d] = acc;
  }
}

__global__ void __launch_bounds__(128) interp_rows_bwd_kernel(const float* __restrict__ ddst, const int* __restrict__ idx,
                                                              const float* __restrict__ w, float* __restrict__ dsrc,
                                                              int n_out, int D, int taps) {
  const int i = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += 128) {
    const float gval = ddst[static_cast<long long>(i) * D + d];
    for (int t = 0; t < taps; ++t) atomicAdd(dsrc + static_cast<long long>(idx[i * taps + t]) * D + d, w[i * taps + t] * gval);
  }
}

// ---------------------------------------------------------------------------------------
// token assembly. x[b,s,:] = (s==0 && cls ? cls : (mask[b,n] ? mask_token : proj[b,n,:])) + pos[s,:]
// patch_embedding.py:61-63,94-95 ; ssl/simmim/model.py:47-49
// ---------------------------------------------------------------------------------------
struct EmbedArgs {
  const __nv_bfloat16* proj;  // [B*N, D]
  const float* cls;           // [D] or null
  const float* pos;           // [S, D]
  const uint8_t* mask;        // [B*N] or null
  const float* mask_token;    // [D] or null
  float* x;                   // [B, S, D]
  int B, N, S, D;
};

__global__ void __launch_bounds__(256) embed_fwd_kernel(const EmbedArgs a, long long total_groups) {
  const long long g = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (g >= total_groups) return;
  const int d8 = a.D / 8;
  const int d = (int)(g % d8) * 8;
  const long long row = g / d8;  // b*S + s
  const int s = (int)(row % a.S);
  const long long b = row / a.S;
  float v[8];
  const int has_cls = a.cls != nullptr;
  if (has_cls && s == 0) {
    const float4 c0 = *reinterpret_cast<const float4*>(a.cls + d);
    const float4 c1 = *reinterpret_cast<const float4*>(a.cls + d + 4);
    v[0] = c0.x; v[1] = c0.y; v[2] = c0.z; v[3] = c0.w; v[4] = c1.x; v[5] = c1.y; v[6] = c1.z; v[7] = c1.w;
  } else {
    const long long pr = b * a.N + (s - has_cls);
    if (a.mask && a.mask[pr]) {
      const float4 c0 = *reinterpret_cast<const float4*>(a.mask_token + d);
      const float4 c1 = *reinterpret_cast<const float4*>(a.mask_token + d + 4);
      v[0] = c0.x; v[1] = c0.y; v[2] = c0.z; v[3] = c0.w; v[4] = c1.x; v[5] = c1.y; v[6] = c1.z; v[7] = c1.w;
    } else {
      const uint4 q = *reinterpret_cast<const uint4*>(a.proj + pr * a.D + d);
      v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
      v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
    }
  }
  const float4 p0 = *reinterpret_cast<const float4*>(a.pos + static_cast<long long>(s) * a.D + d);
  const float4 p1 = *reinterpret_cast<const float4*>(a.pos + static_cast<long long>(s) * a.D + d + 4);
  float* xo = a.x + row * a.D + d;
  *reinterpret_cast<float4*>(xo) = make_float4(v[0] + p0.x, v[1] + p0.y, v[2] + p0.z, v[3] + p0.w);
  *reinterpret_cast<float4*>(xo + 4) = make_float4(v[4] + p1.x, v[5] + p1.y, v[6] + p1.z, v[7] + p1.w);
}

struct EmbedBwdArgs {
  const float* dx; long long ld_b, ld_s;  // gradient of x: element strides for batch and token
  const uint8_t* mask;         // [B*N] or null
  __nv_bfloat16* dproj;        // [B*N, D]
  float* dpos;                 // [S, D]  (pre-zeroed; includes the CLS row)
  float* dmask_token;          // [D] or null (pre-zeroed)
  int B, N, S, D, has_cls, b_per_cta;
};

// one thread per (token s, 8 features); loops over a chunk of the batch
__global__ void __launch_bounds__(128) embed_bwd_kernel(const EmbedBwdArgs a) {
  const int d8 = a.D / 8;
  const int g = blockIdx.x * 128 + threadIdx.x;
  if (g >= a.S * d8) return;
  const int d = (g % d8) * 8, s = g / d8;
  const int b0 = blockIdx.y * a.b_per_cta, b1 = min(a.B, b0 + a.b_per_cta);
  float accp[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float accm[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bool is_cls = a.has_cls && s == 0;
#pragma unroll 4
  for (int b = b0; b < b1; ++b) {
    const float* src = a.dx + b * a.ld_b + s * a.ld_s + d;
    const float4 v0 = *reinterpret_cast<const float4*>(src);
    const float4 v1 = *reinterpret_cast<const float4*>(src + 4);
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) accp[i] += v[i];
    if (!is_cls) {
      const long long pr = static_cast<long long>(b) * a.N + (s - a.has_cls);
      const bool m = a.mask && a.mask[pr];
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (m) {
#pragma unroll
        for (int i = 0; i < 8; ++i) accm[i] += v[i];
      } else {
        o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
        o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
      }
      *reinterpret_cast<uint4*>(a.dproj + pr * a.D + d) = o;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(a.dpos + static_cast<long long>(s) * a.D + d + i, accp[i]);
  if (a.dmask_token && !is_cls) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (accm[i] != 0.f) atomicAdd(a.dmask_token + d + i, accm[i]);
  }
}

// Same, for D <= 1024 with a mask token: every thread of the grid would otherwise add its masked
// partial sums onto the same D addresses (thousands of colliding global atomics per address, which
// made this the slowest elementwise kernel of the SimMIM step). The CTA first folds its tokens
// together in shared memory and then issues one global atomic per feature.
__global__ void __launch_bounds__(128) embed_bwd_masktoken_kernel(const EmbedBwdArgs a) {
  __shared__ float sm_tok[1024];
  for (int i = threadIdx.x; i < a.D; i += 128) sm_tok[i] = 0.f;
  __syncthreads();
  const int d8 = a.D / 8;
  const int g = blockIdx.x * 128 + threadIdx.x;
  if (g < a.S * d8) {
    const int d = (g % d8) * 8, s = g / d8;
    const int b0 = blockIdx.y * a.b_per_cta, b1 = min(a.B, b0 + a.b_per_cta);
    float accp[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float accm[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool is_cls = a.has_cls && s == 0;
#pragma unroll 4
    for (int b = b0; b < b1; ++b) {
      const float* src = a.dx + b * a.ld_b + s * a.ld_s + d;
      const float4 v0 = *reinterpret_cast<const float4*>(src);
      const float4 v1 = *reinterpret_cast<const float4*>(src + 4);
      const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) accp[i] += v[i];
      if (!is_cls) {
        const long long pr = static_cast<long long>(b) * a.N + (s - a.has_cls);
        const bool m = a.mask[pr];
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (m) {
#pragma unroll
          for (int i = 0; i < 8; ++i) accm[i] += v[i];
        } else {
          o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
          o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
        }
        *reinterpret_cast<uint4*>(a.dproj + pr * a.D + d) = o;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(a.dpos + static_cast<long long>(s) * a.D + d + i, accp[i]);
    if (!is_cls) {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&sm_tok[d + i], accm[i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a.D; i += 128)
    if (sm_tok[i] != 0.f) atomicAdd(a.dmask_token + i, sm_tok[i]);
}

// out[i,:] = bf16(x[idx[i], :])   (ssl/simmim/model.py:56 boolean-mask gather, sync-free)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ x, long long ldx,
                                                          const int* __restrict__ idx,
                                                          __nv_bfloat16* __restrict__ out, int D,
                                                          long long total_groups) {
  const long long g = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (g >= total_groups) return;
  const int d8 = D / 8;
  const int d = (int)(g % d8) * 8;
  const long long i = g / d8;
  const float* src = x + static_cast<long long>(idx[i]) * ldx + d;
  const float4 v0 = *reinterpret_cast<const float4*>(src);
  const float4 v1 = *reinterpret_cast<const float4*>(src + 4);
  uint4 o;
  o.x = pack_bf16(v0.x, v0.y); o.y = pack_bf16(v0.z, v0.w);
  o.z = pack_bf16(v1.x, v1.y); o.w = pack_bf16(v1.z, v1.w);
  *reinterpret_cast<uint4*>(out + i * D + d) = o;
}

// dx[r,:] = inv[r] >= 0 ? float(dy[inv[r],:]) : 0   (backward of the gather; writes every row)
__global__ void __launch_bounds__(256) scatter_rows_kernel(const __nv_bfloat16* __restrict__ dy,
                                                           const int* __restrict__ inv,
                                                           float* __restrict__ dx, int D,
                                                           long long total_groups) {
  const long long g = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
  if (g >= total_groups) return;
  const int d8 = D / 8;
  const int d = (int)(g % d8) * 8;
  const long long r = g / d8;
  const int src = inv[r];
  float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
  if (src >= 0) {
    const uint4 q = *reinterpret_cast<const uint4*>(dy + static_cast<long long>(src) * D + d);
    o0 = make_float4(bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y));
    o1 = make_float4(bf16_lo(q.z), bf16_hi(q.z), bf16_lo(q.w), bf16_hi(q.w));
  }
  *reinterpret_cast<float4*>(dx + r * D + d) = o0;
  *reinterpret_cast<float4*>(dx + r * D + d + 4) = o1;
}

}  // namespace
}  // namespace vitssl

using namespace vitssl;

extern "C" int vitssl_multi_cast_bf16(const void* const* host_src, void* const* host_dst,
                                      const int64_t* host_numel, int count, cudaStream_t stream) {
  VITSSL_REQUIRE(count >= 0 && (count == 0 || (host_src && host_dst && host_numel)), VITSSL_ERR_ARG,
                 "multi_cast_bf16: bad args");
  return multi_launch(host_src, const_cast<const void* const*>(host_dst), host_numel, count,
                      [&](const MultiArgs& m, int blocks) {
                        multi_cast_kernel<<<blocks, 256, 0, stream>>>(m);
                        return check_launch("multi_cast_bf16");
                      });
}

extern "C" int vitssl_multi_ema(void* const* host_teacher, const void* const* host_student,
                                const int64_t* host_numel, int count, float momentum,
                                cudaStream_t stream) {
  VITSSL_REQUIRE(count >= 0 && (count == 0 || (host_teacher && host_student && host_numel)),
                 VITSSL_ERR_ARG, "multi_ema: bad args");
  return multi_launch(const_cast<const void* const*>(host_teacher), host_student, host_numel, count,
                      [&](const MultiArgs& m, int blocks) {
                        multi_ema_kernel<<<blocks, 256, 0, stream>>>(m, momentum);
                        return check_launch("multi_ema");
                      });
}

extern "C" int vitssl_colsum_bf16(const void* x, int64_t ld, int64_t rows, int64_t cols, float* out,
                                  cudaStream_t stream) {
  VITSSL_REQUIRE(x && out && rows >= 0 && cols > 0, VITSSL_ERR_ARG, "colsum_bf16: bad args");
  cudaMemsetAsync(out, 0, cols * sizeof(float), stream);
  return vitssl_colsum_bf16_acc(x, ld, rows, cols, out, stream);
}

extern "C" int vitssl_colsum_bf16_acc(const void* x, int64_t ld, int64_t rows, int64_t cols, float* out,
                                      cudaStream_t stream) {
  VITSSL_REQUIRE(x && out && rows >= 0 && cols > 0, VITSSL_ERR_ARG, "colsum_bf16: bad args");
  if (rows == 0) return 0;
  const bool fast = (cols % 8 == 0) && (ld % 8 == 0) && aligned16(x);
  if (fast && ld == cols && cols / 8 <= 288) {
    const int g = (int)(cols / 8);
    const int threads = (288 / g) * g >= 128 ? (288 / g) * g : g * ((128 + g - 1) / g);
    if (threads <= 288) {
      const int k = threads / g;
      long long ctas = static_cast<long long>(num_sms()) * 4;
      const long long max_ctas = (rows + 4ll * k - 1) / (4ll * k);
      if (ctas > max_ctas) ctas = max_ctas;
      const int rows_per_cta = (int)((((rows + ctas - 1) / ctas) + k - 1) / k * k);
      const unsigned grid = (unsigned)((rows + rows_per_cta - 1) / rows_per_cta);
      launch_pdl(colsum_dense_kernel, dim3(grid), dim3(threads), threads * 8 * sizeof(float), stream,
                 (const __nv_bfloat16*)x, (long long)rows, g, out, rows_per_cta);
      return check_launch("colsum_bf16");
    }
  }
  const int col_blocks = (int)((cols + 255) / 256);
  int row_splits = (num_sms() * 4 + col_blocks - 1) / col_blocks;
  if (row_splits > (rows + 31) / 32) row_splits = (int)((rows + 31) / 32);
  if (row_splits < 1) row_splits = 1;
  const int rows_per_cta = (int)((rows + row_splits - 1) / row_splits);
  dim3 grid(col_blocks, (unsigned)((rows + rows_per_cta - 1) / rows_per_cta));
  if (fast)
    colsum_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, ld, rows, (int)cols, out, rows_per_cta);
  else
    colsum_generic_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, ld, rows, (int)cols, out, rows_per_cta);
  return check_launch("colsum_bf16");
}

namespace {
template <typename SRC>
int im2col_launch(const SRC* img, void* out, int64_t B, int64_t C, int64_t H, int64_t W, int64_t p, cudaStream_t stream) {
  VITSSL_REQUIRE(img && out && B > 0 && C > 0 && p > 0, VITSSL_ERR_ARG, "im2col_bf16: bad args");
  VITSSL_REQUIRE(H % p == 0 && W % p == 0, VITSSL_ERR_SHAPE,
                 "im2col_bf16: image %lldx%lld not divisible by patch %lld", (long long)H, (long long)W, (long long)p);
  const long long total = B * C * H * W;
  const bool src_ok = sizeof(SRC) == 1 ? (W % 8 == 0 && (reinterpret_cast<uintptr_t>(img) & 7) == 0) : (W % 4 == 0 && aligned16(img));
  if (p % 8 == 0 && src_ok && aligned16(out)) {
    const long long groups = total / 8;
    im2col_kernel<8, SRC><<<(unsigned)((groups + 255) / 256), 256, 0, stream>>>(img, (__nv_bfloat16*)out, (int)B, (int)C, (int)H, (int)W, (int)p, groups);
  } else {
    im2col_kernel<1, SRC><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(img, (__nv_bfloat16*)out, (int)B, (int)C, (int)H, (int)W, (int)p, total);
  }
  return check_launch("im2col_bf16");
}

template <typename SRC>
int gather_patches_launch(const SRC* img, const int32_t* rows_idx, float* out, int64_t n_rows, int64_t C, int64_t H,
                          int64_t W, int64_t p, cudaStream_t stream) {
  VITSSL_REQUIRE(img && rows_idx && out && n_rows >= 0, VITSSL_ERR_ARG, "gather_patches_f32: bad args");
  VITSSL_REQUIRE(H % p == 0 && W % p == 0, VITSSL_ERR_SHAPE, "gather_patches_f32: image not divisible by patch");
  if (n_rows == 0) return 0;
  const long long total = n_rows * C * p * p;
  const bool src_ok = sizeof(SRC) == 1 ? (reinterpret_cast<uintptr_t>(img) & 3) == 0 : aligned16(img);
  if (p % 4 == 0 && W % 4 == 0 && src_ok && aligned16(out)) {
    const long long groups = total / 4;
    gather_patches_kernel<4, SRC><<<(unsigned)((groups + 255) / 256), 256, 0, stream>>>(img, rows_idx, out, (int)C, (int)H, (int)W, (int)p, groups);
  } else {
    gather_patches_kernel<1, SRC><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(img, rows_idx, out, (int)C, (int)H, (int)W, (int)p, total);
  }
  return check_launch("gather_patches_f32");
}
}  // namespace

extern "C" int vitssl_im2col_bf16(const float* img, void* out, int64_t B, int64_t C, int64_t H,
                                  int64_t W, int64_t p, cudaStream_t stream) {
  return im2col_launch(img, out, B, C, H, W, p, stream);
}

extern "C" int vitssl_im2col_u8_bf16(const uint8_t* img, void* out, int64_t B, int64_t C, int64_t H,
                                     int64_t W, int64_t p, cudaStream_t stream) {
  return im2col_launch(img, out, B, C, H, W, p, stream);
}

extern "C" int vitssl_gather_patches_f32(const float* img, const int32_t* rows_idx, float* out,
                                         int64_t n_rows, int64_t C, int64_t H, int64_t W, int64_t p,
                                         cudaStream_t stream) {
  return gather_patches_launch(img, rows_idx, out, n_rows, C, H, W, p, stream);
}

extern "C" int vitssl_gather_patches_u8_f32(const uint8_t* img, const int32_t* rows_idx, float* out,
                                            int64_t n_rows, int64_t C, int64_t H, int64_t W, int64_t p,
                                            cudaStream_t stream) {
  return gather_patches_launch(img, rows_idx, out, n_rows, C, H, W, p, stream);
}

extern "C" int vitssl_interp_rows_fwd(const float* src, const int32_t* idx, const float* w, float* dst,
                                      int64_t n_out, int64_t D, int64_t taps, cudaStream_t stream) {
  VITSSL_REQUIRE(src && idx && w && dst && n_out > 0 && D > 0 && taps > 0, VITSSL_ERR_ARG, "interp_rows_fwd: bad args");
  interp_rows_fwd_kernel<<<(unsigned)n_out, 128, 0, stream>>>(src, idx, w, dst, (int)n_out, (int)D, (int)taps);
  return check_launch("interp_rows_fwd");
}

extern "C" int vitssl_interp_rows_bwd(const float* ddst, const int32_t* idx, const float* w, float* dsrc,
                                      int64_t n_in, int64_t n_out, int64_t D, int64_t taps, cudaStream_t stream) {
  VITSSL_REQUIRE(ddst && idx && w && dsrc && n_in > 0 && n_out > 0 && D > 0 && taps > 0, VITSSL_ERR_ARG,
                 "interp_rows_bwd: bad args");
  cudaMemsetAsync(dsrc, 0, sizeof(float) * n_in * D, stream);
  interp_rows_bwd_kernel<<<(unsigned)n_out, 128, 0, stream>>>(ddst, idx, w, dsrc, (int)n_out, (int)D, (int)taps);
  return check_launch("interp_rows_bwd");
}

extern "C" int vitssl_embed_tokens_fwd(const void* proj, const float* cls, const float* pos,
                                       const uint8_t* mask, const float* mask_token, float* x,
                                       int64_t B, int64_t N, int64_t D, cudaStream_t stream) {
  VITSSL_REQUIRE(proj && pos && x && B > 0 && N > 0, VITSSL_ERR_ARG, "embed_tokens_fwd: bad args");
  VITSSL_REQUIRE(D % 8 == 0, VITSSL_ERR_SHAPE, "embed_tokens_fwd: embed_dim %lld must be a multiple of 8", (long long)D);
  VITSSL_REQUIRE((mask == nullptr) == (mask_token == nullptr), VITSSL_ERR_ARG, "embed_tokens_fwd: mask and mask_token go together");
  VITSSL_REQUIRE(aligned16(proj) && aligned16(pos) && aligned16(x) && aligned16(cls) && aligned16(mask_token),
                 VITSSL_ERR_ARG, "embed_tokens_fwd: pointers must be 16-byte aligned");
  EmbedArgs a{};
  a.proj = (const __nv_bfloat16*)proj; a.cls = cls; a.pos = pos; a.mask = mask; a.mask_token = mask_token;
  a.x = x; a.B = (int)B; a.N = (int)N; a.S = (int)(N + (cls ? 1 : 0)); a.D = (int)D;
  const long long groups = B * a.S * (D / 8);
  embed_fwd_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, stream>>>(a, groups);
  return check_launch("embed_tokens_fwd");
}

extern "C" int vitssl_embed_tokens_bwd(const float* dx, int64_t ld_b, int64_t ld_s, const uint8_t* mask,
                                       void* dproj, float* dpos, float* dmask_token, int64_t B,
                                       int64_t N, int64_t D, int has_cls, cudaStream_t stream) {
  VITSSL_REQUIRE(dx && dproj && dpos && B > 0 && N > 0, VITSSL_ERR_ARG, "embed_tokens_bwd: bad args");
  VITSSL_REQUIRE(D % 8 == 0 && ld_b % 4 == 0 && ld_s % 4 == 0, VITSSL_ERR_SHAPE, "embed_tokens_bwd: D %% 8 and 16-byte pitches required");
  EmbedBwdArgs a{};
  a.dx = dx; a.ld_b = ld_b; a.ld_s = ld_s; a.mask = mask; a.dproj = (__nv_bfloat16*)dproj; a.dpos = dpos;
  a.dmask_token = dmask_token; a.B = (int)B; a.N = (int)N; a.S = (int)(N + (has_cls ? 1 : 0)); a.D = (int)D;
  a.has_cls = has_cls ? 1 : 0;
  cudaMemsetAsync(dpos, 0, sizeof(float) * a.S * D, stream);
  if (dmask_token) cudaMemsetAsync(dmask_token, 0, sizeof(float) * D, stream);
  const int gx = (int)((a.S * (D / 8) + 127) / 128);
  int splits = (num_sms() * 8 + gx - 1) / gx;
  if (splits > B) splits = (int)B;
  if (splits < 1) splits = 1;
  a.b_per_cta = (int)((B + splits - 1) / splits);
  dim3 grid(gx, (unsigned)((B + a.b_per_cta - 1) / a.b_per_cta));
  if (dmask_token && mask && D <= 1024) embed_bwd_masktoken_kernel<<<grid, 128, 0, stream>>>(a);
  else embed_bwd_kernel<<<grid, 128, 0, stream>>>(a);
  return check_launch("embed_tokens_bwd");
}

extern "C" int vitssl_gather_rows_bf16(const float* x, int64_t ldx, const int32_t* idx, void* out,
                                       int64_t n_rows, int64_t D, cudaStream_t stream) {
  VITSSL_REQUIRE(x && idx && out && n_rows >= 0, VITSSL_ERR_ARG, "gather_rows_bf16: bad args");
  VITSSL_REQUIRE(D % 8 == 0 && ldx % 4 == 0 && aligned16(x) && aligned16(out), VITSSL_ERR_SHAPE, "gather_rows_bf16: D %% 8 and alignment required");
  if (n_rows == 0) return 0;
  const long long groups = n_rows * (D / 8);
  gather_rows_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, stream>>>(x, ldx, idx, (__nv_bfloat16*)out, (int)D, groups);
  return check_launch("gather_rows_bf16");
}

extern "C" int vitssl_scatter_rows_f32(const void* dy, const int32_t* inv_idx, float* dx, int64_t rows,
                                       int64_t D, cudaStream_t stream) {
  VITSSL_REQUIRE(dy && inv_idx && dx && rows >= 0, VITSSL_ERR_ARG, "scatter_rows_f32: bad args");
  VITSSL_REQUIRE(D % 8 == 0 && aligned16(dy) && aligned16(dx), VITSSL_ERR_SHAPE, "scatter_rows_f32: D %% 8 and alignment required");
  if (rows == 0) return 0;
  const long long groups = rows * (D / 8);
  scatter_rows_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, stream>>>((const __nv_bfloat16*)dy, inv_idx, dx, (int)D, groups);
  return check_launch("scatter_rows_f32");
}
