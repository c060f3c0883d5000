// vitssl_b200 — SimMIM and DINO objective kernels (all HBM-bound, single pass).
//   SimMIM: masked-patch L1 reconstruction loss (nn.L1Loss(mean), utils/train_utils.py:19-22,
//           configs/simmim/training.yaml:2-5) emitting the loss and sign(pred - target).
//   DINO:   L2 row normalisation (ssl/dino/head.py:21), weight-norm of the last layer (head.py:17),
//           center EMA (ssl/dino/model.py:91-99), the temperature-sharpened cross-entropy over all
//           (teacher view, student view) pairs in its factorised form (ssl/dino/loss.py:13-29;
//           SURVEY App. A-7) and its closed-form gradient.
#include "common.cuh"
#include "vitssl_b200.h"

namespace vitssl {
namespace {

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
  v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
}

__device__ __forceinline__ float block_sum(float v, float* red) {  // red: >= 32 floats
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float s = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (warp == 0) s = warp_sum(s);
  return s;  // valid in warp 0
}

// ---------------------------------------------------------------------------------------
// SimMIM masked L1: loss += sum|p - t| / n ; sign[i] = sign(p - t) (bf16)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l1_loss_kernel(const __nv_bfloat16* __restrict__ pred,
                                                      const float* __restrict__ target,
                                                      __nv_bfloat16* __restrict__ sign, long long n,
                                                      float inv_n, float* __restrict__ loss) {
  __shared__ float red[32];
  float acc = 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * 256 * 8;
  for (long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      float p[8];
      unpack8(*reinterpret_cast<const uint4*>(pred + i), p);
      const float4 t0 = *reinterpret_cast<const float4*>(target + i);
      const float4 t1 = *reinterpret_cast<const float4*>(target + i + 4);
      const float t[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
      float sg[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = p[j] - t[j];
        acc += fabsf(d);
        sg[j] = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      }
      if (sign) {
        uint4 o;
        o.x = pack_bf16(sg[0], sg[1]); o.y = pack_bf16(sg[2], sg[3]);
        o.z = pack_bf16(sg[4], sg[5]); o.w = pack_bf16(sg[6], sg[7]);
        *reinterpret_cast<uint4*>(sign + i) = o;
      }
    } else {
      for (long long j = i; j < n; ++j) {
        const float d = __bfloat162float(pred[j]) - target[j];
        acc += fabsf(d);
        if (sign) sign[j] = __float2bfloat16_rn(d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
      }
    }
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss, s * inv_n);
}

// d(pred) = sign * (grad_out / n): grad_out is a DEVICE scalar (the GradScaler-scaled upstream gradient)
__global__ void __launch_bounds__(256) l1_loss_bwd_kernel(const __nv_bfloat16* __restrict__ sign,
                                                          const float* __restrict__ grad_out, float inv_n,
                                                          __nv_bfloat16* __restrict__ dpred, long long n) {
  const float s = *grad_out * inv_n;
  const long long stride = static_cast<long long>(gridDim.x) * 256 * 8;
  for (long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(sign + i), v);
      uint4 o;
      o.x = pack_bf16(v[0] * s, v[1] * s); o.y = pack_bf16(v[2] * s, v[3] * s);
      o.z = pack_bf16(v[4] * s, v[5] * s); o.w = pack_bf16(v[6] * s, v[7] * s);
      *reinterpret_cast<uint4*>(dpred + i) = o;
    } else {
      for (long long j = i; j < n; ++j) dpred[j] = __float2bfloat16_rn(__bfloat162float(sign[j]) * s);
    }
  }
}

// ---------------------------------------------------------------------------------------
// row L2 normalisation (F.normalize(dim=1), eps 1e-12): y = x / max(||x||, eps)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) l2norm_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                         __nv_bfloat16* __restrict__ y,
                                                         float* __restrict__ inv_norm, long long rows, int D) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) { const float v = __bfloat162float(x[row * D + c]); ss += v * v; }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  if (lane == 0) inv_norm[row] = inv;
  for (int c = lane; c < D; c += 32) y[row * D + c] = __float2bfloat16_rn(__bfloat162float(x[row * D + c]) * inv);
}
// dx = inv * (dy - y * <y, dy>)   with y = x * inv recomputed in fp32
__global__ void __launch_bounds__(128) l2norm_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                         const float* __restrict__ inv_norm,
                                                         const __nv_bfloat16* __restrict__ dy,
                                                         __nv_bfloat16* __restrict__ dx, long long rows, int D) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float inv = inv_norm[row];
  float dot = 0.f;
  for (int c = lane; c < D; c += 32)
    dot += __bfloat162float(x[row * D + c]) * inv * __bfloat162float(dy[row * D + c]);
  dot = warp_sum(dot);
  for (int c = lane; c < D; c += 32) {
    const float yv = __bfloat162float(x[row * D + c]) * inv;
    dx[row * D + c] = __float2bfloat16_rn(inv * (__bfloat162float(dy[row * D + c]) - yv * dot));
  }
}

// ---------------------------------------------------------------------------------------
// weight norm: w[k,:] = g[k] * v[k,:] / ||v[k,:]||  -> bf16 (+ inv_norm saved)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) weight_norm_fwd_kernel(const float* __restrict__ v,
                                                              const float* __restrict__ g,
                                                              __nv_bfloat16* __restrict__ w,
                                                              float* __restrict__ inv_norm, long long rows, int D) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* vr = v + row * D;
  float ss = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 q = *reinterpret_cast<const float4*>(vr + c);
    ss += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  }
  ss = warp_sum(ss);
  const float inv = rsqrtf(ss);
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
  const float sc = g[row] * inv;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 q = *reinterpret_cast<const float4*>(vr + c);
    uint2 o;
    o.x = pack_bf16(q.x * sc, q.y * sc); o.y = pack_bf16(q.z * sc, q.w * sc);
    *reinterpret_cast<uint2*>(w + row * D + c) = o;
  }
}
// dg[k] = <dW[k], v[k]> * inv ; dv[k] = g*inv * (dW[k] - v[k] * <dW[k],v[k]> * inv^2)
__global__ void __launch_bounds__(128) weight_norm_bwd_kernel(const float* __restrict__ dw,
                                                              const float* __restrict__ v,
                                                              const float* __restrict__ g,
                                                              const float* __restrict__ inv_norm,
                                                              float* __restrict__ dg, float* __restrict__ dv,
                                                              long long rows, int D) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* vr = v + row * D;
  const float* dr = dw + row * D;
  float dot = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 a = *reinterpret_cast<const float4*>(vr + c);
    const float4 b = *reinterpret_cast<const float4*>(dr + c);
    dot += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  }
  dot = warp_sum(dot);
  const float inv = inv_norm[row];
  if (lane == 0) dg[row] = dot * inv;
  const float a1 = g[row] * inv, a2 = dot * inv * inv;
  for (int c = lane * 4; c < D; c += 128) {
    const float4 a = *reinterpret_cast<const float4*>(vr + c);
    const float4 b = *reinterpret_cast<const float4*>(dr + c);
    *reinterpret_cast<float4*>(dv + row * D + c) =
        make_float4(a1 * (b.x - a.x * a2), a1 * (b.y - a.y * a2), a1 * (b.z - a.z * a2), a1 * (b.w - a.w * a2));
  }
}

// center <- m * center + (1 - m) * colsum * inv_rows     (ssl/dino/model.py:96-99)
__global__ void center_ema_kernel(const float* __restrict__ center, const float* __restrict__ colsum,
                                  float* __restrict__ out, int K, float m, float inv_rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K) out[i] = m * center[i] + (1.0f - m) * (colsum[i] * inv_rows);
}

// ---------------------------------------------------------------------------------------
// DINO loss, factorised:  L = -(1/(G B K)) sum_b [ sum_g N_g / Z_g  -  G * sum_v lse_v ]
//   a_gk = (t_gk - c_k)/tau_t,  X_k = sum_v s'_vk,  s'_vk = s_vk / tau_s in fp32 (loss.py:23; the
//   reference under autocast rounds s/tau_s to bf16 first — we keep fp32, which is closer to its fp32 math)
//   N_g = sum_k exp(a_gk - m_g) X_k,  Z_g = sum_k exp(a_gk - m_g),  lse_v = logsumexp_k s'_vk
// One CTA per batch sample streams its G + V rows once with online (max, sum) rescaling.
// ---------------------------------------------------------------------------------------
constexpr int DL_MAX_G = 4;
constexpr int DL_MAX_V = 12;
constexpr int DL_THREADS = 512;

struct DinoLossArgs {
  const __nv_bfloat16* teacher;  // [G,B,K]
  const __nv_bfloat16* student;  // [V,B,K]
  const float* center;           // [K]
  float* loss;                   // [1], pre-zeroed
  float* t_stats;                // [G,B,2] (max, Z) of a_g
  float* s_lse;                  // [V,B]
  int G, V, B, K;
  float inv_tt, inv_ts;
};

__device__ __forceinline__ void online_merge(float& m, float& z, float m2, float z2) {
  const float mn = fmaxf(m, m2);
  z = z * __expf(m - mn) + z2 * __expf(m2 - mn);
  m = mn;
}

template <int G, int V>
__global__ void __launch_bounds__(DL_THREADS) dino_loss_fwd_kernel(const DinoLossArgs a) {
  __shared__ float sh[(3 * DL_MAX_G + 2 * DL_MAX_V) * (DL_THREADS / 32)];
  const int b = blockIdx.x;
  float tm[G], tz[G], tn[G], sm[V], sz[V];
#pragma unroll
  for (int g = 0; g < G; ++g) { tm[g] = -INFINITY; tz[g] = 0.f; tn[g] = 0.f; }
#pragma unroll
  for (int v = 0; v < V; ++v) { sm[v] = -INFINITY; sz[v] = 0.f; }

  for (int k = threadIdx.x * 8; k < a.K; k += DL_THREADS * 8) {
    float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float s[8];
      unpack8(*reinterpret_cast<const uint4*>(a.student + (static_cast<long long>(v) * a.B + b) * a.K + k), s);
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] = s[j] * a.inv_ts; mx = fmaxf(mx, s[j]); x[j] += s[j]; }
      const float mn = fmaxf(sm[v], mx);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += __expf(s[j] - mn);
      sz[v] = sz[v] * __expf(sm[v] - mn) + acc;
      sm[v] = mn;
    }
    const float4 c0 = *reinterpret_cast<const float4*>(a.center + k);
    const float4 c1 = *reinterpret_cast<const float4*>(a.center + k + 4);
    const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float t[8];
      unpack8(*reinterpret_cast<const uint4*>(a.teacher + (static_cast<long long>(g) * a.B + b) * a.K + k), t);
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) { t[j] = (t[j] - c[j]) * a.inv_tt; mx = fmaxf(mx, t[j]); }
      const float mn = fmaxf(tm[g], mx);
      const float resc = __expf(tm[g] - mn);
      float accz = 0.f, accn = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float e = __expf(t[j] - mn); accz += e; accn += e * x[j]; }
      tz[g] = tz[g] * resc + accz;
      tn[g] = tn[g] * resc + accn;
      tm[g] = mn;
    }
  }
  // warp-level merge, then cross-warp through shared memory
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = DL_THREADS / 32;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float m2 = __shfl_xor_sync(0xffffffffu, tm[g], o), z2 = __shfl_xor_sync(0xffffffffu, tz[g], o),
                  n2 = __shfl_xor_sync(0xffffffffu, tn[g], o);
      const float mn = fmaxf(tm[g], m2);
      const float e1 = (tm[g] == -INFINITY) ? 0.f : __expf(tm[g] - mn), e2 = (m2 == -INFINITY) ? 0.f : __expf(m2 - mn);
      tz[g] = tz[g] * e1 + z2 * e2; tn[g] = tn[g] * e1 + n2 * e2; tm[g] = mn;
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float m2 = __shfl_xor_sync(0xffffffffu, sm[v], o), z2 = __shfl_xor_sync(0xffffffffu, sz[v], o);
      const float mn = fmaxf(sm[v], m2);
      const float e1 = (sm[v] == -INFINITY) ? 0.f : __expf(sm[v] - mn), e2 = (m2 == -INFINITY) ? 0.f : __expf(m2 - mn);
      sz[v] = sz[v] * e1 + z2 * e2; sm[v] = mn;
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      sh[(3 * g + 0) * NW + warp] = tm[g]; sh[(3 * g + 1) * NW + warp] = tz[g]; sh[(3 * g + 2) * NW + warp] = tn[g];
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
      sh[(3 * G + 2 * v) * NW + warp] = sm[v]; sh[(3 * G + 2 * v + 1) * NW + warp] = sz[v];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float total = 0.f, lse_sum = 0.f;
    for (int g = 0; g < G; ++g) {
      float m = -INFINITY, z = 0.f, n = 0.f;
      for (int w = 0; w < NW; ++w) {
        const float m2 = sh[(3 * g) * NW + w], z2 = sh[(3 * g + 1) * NW + w], n2 = sh[(3 * g + 2) * NW + w];
        if (m2 == -INFINITY) continue;
        const float mn = fmaxf(m, m2);
        const float e1 = (m == -INFINITY) ? 0.f : __expf(m - mn), e2 = __expf(m2 - mn);
        z = z * e1 + z2 * e2; n = n * e1 + n2 * e2; m = mn;
      }
      a.t_stats[(static_cast<long long>(g) * a.B + b) * 2 + 0] = m;
      a.t_stats[(static_cast<long long>(g) * a.B + b) * 2 + 1] = z;
      total += n / z;
    }
    for (int v = 0; v < V; ++v) {
      float m = -INFINITY, z = 0.f;
      for (int w = 0; w < NW; ++w) {
        const float m2 = sh[(3 * G + 2 * v) * NW + w], z2 = sh[(3 * G + 2 * v + 1) * NW + w];
        if (m2 == -INFINITY) continue;
        const float mn = fmaxf(m, m2);
        const float e1 = (m == -INFINITY) ? 0.f : __expf(m - mn), e2 = __expf(m2 - mn);
        z = z * e1 + z2 * e2; m = mn;
      }
      const float lse = m + __logf(z);
      a.s_lse[static_cast<long long>(v) * a.B + b] = lse;
      lse_sum += lse;
    }
    const float lb = total - static_cast<float>(G) * lse_sum;
    atomicAdd(a.loss, -lb / (static_cast<float>(G) * a.B * a.K));
  }
}

// dS[v,b,k] = -(go / (tau_s G B K)) * (Pbar[b,k] - G * softmax(s'_v)[k])      
struct DinoLossBwdArgs {
  const __nv_bfloat16* teacher; const __nv_bfloat16* student; const float* center;
  const float* t_stats; const float* s_lse; const float* grad_out;  // device scalar
  __nv_bfloat16* dstudent;
  int G, V, B, K;
  float inv_tt, inv_ts;
};

template <int G, int V>
__global__ void __launch_bounds__(256) dino_loss_bwd_kernel(const DinoLossBwdArgs a) {
  const int b = blockIdx.y;
  const int k = (blockIdx.x * 256 + threadIdx.x) * 8;
  if (k >= a.K) return;
  const float coef = -(*a.grad_out) * a.inv_ts / (static_cast<float>(G) * a.B * a.K);
  const float4 c0 = *reinterpret_cast<const float4*>(a.center + k);
  const float4 c1 = *reinterpret_cast<const float4*>(a.center + k + 4);
  const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
  float pbar[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float t[8];
    unpack8(*reinterpret_cast<const uint4*>(a.teacher + (static_cast<long long>(g) * a.B + b) * a.K + k), t);
    const float m = a.t_stats[(static_cast<long long>(g) * a.B + b) * 2], iz = 1.0f / a.t_stats[(static_cast<long long>(g) * a.B + b) * 2 + 1];
#pragma unroll
    for (int j = 0; j < 8; ++j) pbar[j] += __expf((t[j] - c[j]) * a.inv_tt - m) * iz;
  }
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const long long off = (static_cast<long long>(v) * a.B + b) * a.K + k;
    float s[8];
    unpack8(*reinterpret_cast<const uint4*>(a.student + off), s);
    const float lse = a.s_lse[static_cast<long long>(v) * a.B + b];
    float d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      d[j] = coef * (pbar[j] - static_cast<float>(G) * __expf(s[j] * a.inv_ts - lse));
    uint4 o;
    o.x = pack_bf16(d[0], d[1]); o.y = pack_bf16(d[2], d[3]); o.z = pack_bf16(d[4], d[5]); o.w = pack_bf16(d[6], d[7]);
    *reinterpret_cast<uint4*>(a.dstudent + off) = o;
  }
}

template <int G>
int dino_fwd_dispatch_v(const DinoLossArgs& a, cudaStream_t st) {
  switch (a.V) {
    case 1: dino_loss_fwd_kernel<G, 1><<<a.B, DL_THREADS, 0, st>>>(a); break;
    case 2: dino_loss_fwd_kernel<G, 2><<<a.B, DL_THREADS, 0, st>>>(a); break;
    case 3: dino_loss_fwd_kernel<G, 3><<<a.B, DL_THREADS, 0, st>>>(a); break;
    case 4: dino_loss_fwd_kernel<G, 4><<<a.B, DL_THREADS, 0, st>>>(a); break;
    case 5: dino_loss_fwd_kernel<G, 5><<<a.B, DL_THREADS, 0, st>>>(a); break;
    case 6: dino_loss_fwd_kernel<G, 6><<<a.B, DL_THREADS, 0, st>>>(a); break;
    case 7: dino_loss_fwd_kernel<G, 7><<<a.B, DL_THREADS, 0, st>>>(a); break;
    case 8: dino_loss_fwd_kernel<G, 8><<<a.B, DL_THREADS, 0, st>>>(a); break;
    case 9: dino_loss_fwd_kernel<G, 9><<<a.B, DL_THREADS, 0, st>>>(a); break;
    case 10: dino_loss_fwd_kernel<G, 10><<<a.B, DL_THREADS, 0, st>>>(a); break;
    case 11: dino_loss_fwd_kernel<G, 11><<<a.B, DL_THREADS, 0, st>>>(a); break;
    case 12: dino_loss_fwd_kernel<G, 12><<<a.B, DL_THREADS, 0, st>>>(a); break;
    default: set_error("dino_loss: unsupported number of views V=%d (supported: 1..12)", a.V); return VITSSL_ERR_SHAPE;
  }
  return check_launch("dino_loss_fwd");
}
template <int G>
int dino_bwd_dispatch_v(const DinoLossBwdArgs& a, dim3 grid, cudaStream_t st) {
  switch (a.V) {
    case 1: dino_loss_bwd_kernel<G, 1><<<grid, 256, 0, st>>>(a); break;
    case 2: dino_loss_bwd_kernel<G, 2><<<grid, 256, 0, st>>>(a); break;
    case 3: dino_loss_bwd_kernel<G, 3><<<grid, 256, 0, st>>>(a); break;
    case 4: dino_loss_bwd_kernel<G, 4><<<grid, 256, 0, st>>>(a); break;
    case 5: dino_loss_bwd_kernel<G, 5><<<grid, 256, 0, st>>>(a); break;
    case 6: dino_loss_bwd_kernel<G, 6><<<grid, 256, 0, st>>>(a); break;
    case 7: dino_loss_bwd_kernel<G, 7><<<grid, 256, 0, st>>>(a); break;
    case 8: dino_loss_bwd_kernel<G, 8><<<grid, 256, 0, st>>>(a); break;
    case 9: dino_loss_bwd_kernel<G, 9><<<grid, 256, 0, st>>>(a); break;
    case 10: dino_loss_bwd_kernel<G, 10><<<grid, 256, 0, st>>>(a); break;
    case 11: dino_loss_bwd_kernel<G, 11><<<grid, 256, 0, st>>>(a); break;
    case 12: dino_loss_bwd_kernel<G, 12><<<grid, 256, 0, st>>>(a); break;
    default: set_error("dino_loss: unsupported number of views V=%d", a.V); return VITSSL_ERR_SHAPE;
  }
  return check_launch("dino_loss_bwd");
}

}  // namespace
}  // namespace vitssl

using namespace vitssl;

extern "C" int vitssl_l1_loss_fwd(const void* pred, const float* target, void* sign, float* loss,
                                  int64_t n, cudaStream_t stream) {
  VITSSL_REQUIRE(pred && target && loss && n > 0, VITSSL_ERR_ARG, "l1_loss_fwd: bad args");
  VITSSL_REQUIRE(aligned16(pred) && aligned16(target) && aligned16(sign), VITSSL_ERR_ARG, "l1_loss_fwd: 16-byte alignment required");
  cudaMemsetAsync(loss, 0, sizeof(float), stream);
  long long blocks = (n / 8 + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  l1_loss_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const __nv_bfloat16*)pred, target, (__nv_bfloat16*)sign, n, 1.0f / (float)n, loss);
  return check_launch("l1_loss_fwd");
}

extern "C" int vitssl_l1_loss_bwd(const void* sign, const float* grad_out, void* dpred, int64_t n,
                                  cudaStream_t stream) {
  VITSSL_REQUIRE(sign && grad_out && dpred && n > 0, VITSSL_ERR_ARG, "l1_loss_bwd: bad args");
  VITSSL_REQUIRE(aligned16(sign) && aligned16(dpred), VITSSL_ERR_ARG, "l1_loss_bwd: 16-byte alignment required");
  long long blocks = (n / 8 + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  l1_loss_bwd_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const __nv_bfloat16*)sign, grad_out, 1.0f / (float)n,
                                                           (__nv_bfloat16*)dpred, n);
  return check_launch("l1_loss_bwd");
}

extern "C" int vitssl_l2norm_fwd(const void* x, void* y, float* inv_norm, int64_t rows, int64_t D, cudaStream_t stream) {
  VITSSL_REQUIRE(x && y && inv_norm && rows >= 0 && D > 0, VITSSL_ERR_ARG, "l2norm_fwd: bad args");
  if (rows == 0) return 0;
  l2norm_fwd_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, inv_norm, rows, (int)D);
  return check_launch("l2norm_fwd");
}
extern "C" int vitssl_l2norm_bwd(const void* x, const float* inv_norm, const void* dy, void* dx, int64_t rows, int64_t D, cudaStream_t stream) {
  VITSSL_REQUIRE(x && inv_norm && dy && dx && rows >= 0 && D > 0, VITSSL_ERR_ARG, "l2norm_bwd: bad args");
  if (rows == 0) return 0;
  l2norm_bwd_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, stream>>>((const __nv_bfloat16*)x, inv_norm, (const __nv_bfloat16*)dy, (__nv_bfloat16*)dx, rows, (int)D);
  return check_launch("l2norm_bwd");
}

extern "C" int vitssl_weight_norm_fwd(const float* v, const float* g, void* w, float* inv_norm, int64_t rows, int64_t D, cudaStream_t stream) {
  VITSSL_REQUIRE(v && g && w && rows > 0 && D > 0, VITSSL_ERR_ARG, "weight_norm_fwd: bad args");
  VITSSL_REQUIRE(D % 4 == 0 && aligned16(v) && aligned16(w), VITSSL_ERR_SHAPE, "weight_norm_fwd: D %% 4 and alignment required");
  weight_norm_fwd_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, stream>>>(v, g, (__nv_bfloat16*)w, inv_norm, rows, (int)D);
  return check_launch("weight_norm_fwd");
}
extern "C" int vitssl_weight_norm_bwd(const float* dw, const float* v, const float* g, const float* inv_norm, float* dg, float* dv, int64_t rows, int64_t D, cudaStream_t stream) {
  VITSSL_REQUIRE(dw && v && g && inv_norm && dg && dv && rows > 0, VITSSL_ERR_ARG, "weight_norm_bwd: bad args");
  VITSSL_REQUIRE(D % 4 == 0 && aligned16(v) && aligned16(dw) && aligned16(dv), VITSSL_ERR_SHAPE, "weight_norm_bwd: D %% 4 and alignment required");
  weight_norm_bwd_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, stream>>>(dw, v, g, inv_norm, dg, dv, rows, (int)D);
  return check_launch("weight_norm_bwd");
}

extern "C" int vitssl_center_ema(const float* center, const float* colsum, float* out, int64_t K, float momentum, float inv_rows, cudaStream_t stream) {
  VITSSL_REQUIRE(center && colsum && out && K > 0, VITSSL_ERR_ARG, "center_ema: bad args");
  center_ema_kernel<<<(unsigned)((K + 255) / 256), 256, 0, stream>>>(center, colsum, out, (int)K, momentum, inv_rows);
  return check_launch("center_ema");
}

extern "C" int vitssl_dino_loss_fwd(const void* teacher, const void* student, const float* center, float* loss,
                                    float* t_stats, float* s_lse, int64_t G, int64_t V, int64_t B, int64_t K,
                                    float teacher_temp, float student_temp, cudaStream_t stream) {
  VITSSL_REQUIRE(teacher && student && center && loss && t_stats && s_lse, VITSSL_ERR_ARG, "dino_loss_fwd: null pointer");
  VITSSL_REQUIRE(G >= 1 && G <= DL_MAX_G && B > 0 && K > 0 && K % 8 == 0, VITSSL_ERR_SHAPE,
                 "dino_loss_fwd: unsupported G=%lld (1..4) or K=%lld (multiple of 8)", (long long)G, (long long)K);
  VITSSL_REQUIRE(teacher_temp > 0.f && student_temp > 0.f, VITSSL_ERR_ARG, "dino_loss_fwd: temperatures must be positive");
  DinoLossArgs a{};
  a.teacher = (const __nv_bfloat16*)teacher; a.student = (const __nv_bfloat16*)student; a.center = center;
  a.loss = loss; a.t_stats = t_stats; a.s_lse = s_lse; a.G = (int)G; a.V = (int)V; a.B = (int)B; a.K = (int)K;
  a.inv_tt = 1.0f / teacher_temp; a.inv_ts = 1.0f / student_temp;
  cudaMemsetAsync(loss, 0, sizeof(float), stream);
  switch (G) {
    case 1: return dino_fwd_dispatch_v<1>(a, stream);
    case 2: return dino_fwd_dispatch_v<2>(a, stream);
    case 3: return dino_fwd_dispatch_v<3>(a, stream);
    default: return dino_fwd_dispatch_v<4>(a, stream);
  }
}

extern "C" int vitssl_dino_loss_bwd(const void* teacher, const void* student, const float* center, const float* t_stats,
                                    const float* s_lse, const float* grad_out, void* dstudent, int64_t G, int64_t V,
                                    int64_t B, int64_t K, float teacher_temp, float student_temp, cudaStream_t stream) {
  VITSSL_REQUIRE(teacher && student && center && t_stats && s_lse && grad_out && dstudent, VITSSL_ERR_ARG, "dino_loss_bwd: null pointer");
  VITSSL_REQUIRE(G >= 1 && G <= DL_MAX_G && B > 0 && K > 0 && K % 8 == 0, VITSSL_ERR_SHAPE, "dino_loss_bwd: unsupported shape");
  DinoLossBwdArgs a{};
  a.teacher = (const __nv_bfloat16*)teacher; a.student = (const __nv_bfloat16*)student; a.center = center;
  a.t_stats = t_stats; a.s_lse = s_lse; a.grad_out = grad_out; a.dstudent = (__nv_bfloat16*)dstudent;
  a.G = (int)G; a.V = (int)V; a.B = (int)B; a.K = (int)K; a.inv_tt = 1.0f / teacher_temp; a.inv_ts = 1.0f / student_temp;
  dim3 grid((unsigned)((K / 8 + 255) / 256), (unsigned)B);
  switch (G) {
    case 1: return dino_bwd_dispatch_v<1>(a, grid, stream);
    case 2: return dino_bwd_dispatch_v<2>(a, grid, stream);
    case 3: return dino_bwd_dispatch_v<3>(a, grid, stream);
    default: return dino_bwd_dispatch_v<4>(a, grid, stream);
  }
}
