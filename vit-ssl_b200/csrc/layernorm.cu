// vitssl_b200 — fused residual-add (+dropout) + LayerNorm, forward and backward.
// Reference: encoder_block.py:40-52 (x = drop(branch) + residual; x = layer_norm(x)), LayerNorm
// eps 1e-5 with affine (encoder_block.py:26-27, mlp_head.py:9). The reference under autocast keeps
// the residual stream and the LayerNorm in fp32 and feeds bf16 to the next linear.
//
// HBM-bound, single pass: one warp owns a row, the row lives in registers (8 contiguous elements
// per lane per step -> 2x float4 loads of the stream, 1x 16-byte load of the bf16 branch).
//   fwd traffic / element: 4 (x) + 2 (branch) + 4 (x_out) + 2 (y)            = 12 B
//   bwd traffic / element: 2 (dy) + 4 (x) + 4 (dres) + 4 (dx) + 2 (dbranch) = 16 B
#include <stdlib.h>

#include "common.cuh"
#include "vitssl_b200.h"

namespace vitssl {
namespace {

#ifndef VITSSL_LN_WARPS
#define VITSSL_LN_WARPS 4  // rows (= warps) per CTA; 2 and 8 measured slower or equal (profiles/README.md)
#endif
constexpr int LN_WARPS = VITSSL_LN_WARPS;

struct LnFwdArgs {
  const float* x; long long ldx;          // input stream rows (pitch in elements)
  const __nv_bfloat16* branch;             // nullable, [rows, D] dense
  float* x_out;                            // nullable (required iff branch), dense
  const float* gamma; const float* beta;   // nullable together -> no LayerNorm, add only
  __nv_bfloat16* y;                        // LN output (bf16), dense
  float* mean; float* rstd;                // [rows]
  long long rows; int D; float eps;
  uint32_t drop_thresh16; float drop_scale; unsigned long long seed, offset;
  const unsigned long long* seed_slot;  // non-null: the seed is read from this device word (graph replay, encoder.cu)
};

struct LnBwdArgs {
  const __nv_bfloat16* dy;   // nullable: grad wrt LN output, dense [rows, D]
  const float* x; long long ldx;  // LN input rows
  const float* mean; const float* rstd; const float* gamma;
  const float* dres; long long ld_dres;  // nullable: grad arriving on the residual stream
  float* dx; long long ld_dx;            // out: total grad wrt the LN input / stream (fp32)
  __nv_bfloat16* dbranch;    // nullable out: grad wrt the branch = mask/(1-p) * dx, bf16 dense
  float* dgamma; float* dbeta;  // [D], pre-zeroed, accumulated with atomics
  long long rows; int D;
  uint32_t drop_thresh16; float drop_scale; unsigned long long seed, offset;
  const unsigned long long* seed_slot;  // non-null: the seed is read from this device word (graph replay, encoder.cu)
};

// dropout seed of a launch: its argument, or the device word a replayed CUDA graph points at
template <typename Args>
__device__ __forceinline__ unsigned long long eff_seed(const Args& a) {
  return a.seed_slot != nullptr ? __ldg(a.seed_slot) : a.seed;
}

// 8 consecutive elements of a row as four packed fp32 pairs (FFMA2 / FADD2 / FMUL2 operate on both)
struct Vec8 {
  f32x2 p[4];
};
__device__ __forceinline__ Vec8 load8_f32(const float* ptr) {
  const float4 a = *reinterpret_cast<const float4*>(ptr);
  const float4 b = *reinterpret_cast<const float4*>(ptr + 4);
  Vec8 v;
  v.p[0] = pk2(a.x, a.y); v.p[1] = pk2(a.z, a.w); v.p[2] = pk2(b.x, b.y); v.p[3] = pk2(b.z, b.w);
  return v;
}
__device__ __forceinline__ Vec8 ldg8_f32(const float* ptr) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(ptr));
  const float4 b = __ldg(reinterpret_cast<const float4*>(ptr + 4));
  Vec8 v;
  v.p[0] = pk2(a.x, a.y); v.p[1] = pk2(a.z, a.w); v.p[2] = pk2(b.x, b.y); v.p[3] = pk2(b.z, b.w);
  return v;
}
__device__ __forceinline__ void store8_f32(float* ptr, const Vec8& v) {
  float4 a, b;
  upk2(v.p[0], a.x, a.y); upk2(v.p[1], a.z, a.w); upk2(v.p[2], b.x, b.y); upk2(v.p[3], b.z, b.w);
  *reinterpret_cast<float4*>(ptr) = a;
  *reinterpret_cast<float4*>(ptr + 4) = b;
}
__device__ __forceinline__ Vec8 load8_bf16(const __nv_bfloat16* ptr) {
  const uint4 b = *reinterpret_cast<const uint4*>(ptr);
  Vec8 v;
  v.p[0] = pk2(bf16_lo(b.x), bf16_hi(b.x)); v.p[1] = pk2(bf16_lo(b.y), bf16_hi(b.y));
  v.p[2] = pk2(bf16_lo(b.z), bf16_hi(b.z)); v.p[3] = pk2(bf16_lo(b.w), bf16_hi(b.w));
  return v;
}
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* ptr, const Vec8& v) {
  uint4 pk;
  float a, b;
  upk2(v.p[0], a, b); pk.x = pack_bf16(a, b);
  upk2(v.p[1], a, b); pk.y = pack_bf16(a, b);
  upk2(v.p[2], a, b); pk.z = pack_bf16(a, b);
  upk2(v.p[3], a, b); pk.w = pack_bf16(a, b);
  *reinterpret_cast<uint4*>(ptr) = pk;
}
__device__ __forceinline__ float hsum8(const Vec8& v) {
  float a, b;
  upk2(fadd2(fadd2(v.p[0], v.p[1]), fadd2(v.p[2], v.p[3])), a, b);
  return a + b;
}
// keep-multipliers (0 or 1/(1-p)) of the 8 elements of dropout group `group8`
__device__ __forceinline__ Vec8 dropout_mult8(unsigned long long seed, unsigned long long offset,
                                              unsigned long long group8, uint32_t thresh16, float scale) {
  float m[8];
  dropout_scale8(seed, offset, group8, thresh16, scale, m);
  Vec8 v;
  v.p[0] = pk2(m[0], m[1]); v.p[1] = pk2(m[2], m[3]); v.p[2] = pk2(m[4], m[5]); v.p[3] = pk2(m[6], m[7]);
  return v;
}

template <int NG>  // NG = ceil(D / 256), D % 8 == 0
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const LnFwdArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const f32x2 zero = pk2(0.f, 0.f);
  // grid-stride over rows: a few resident CTAs per SM stream the whole tensor
  for (long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5); row < a.rows;
       row += static_cast<long long>(gridDim.x) * LN_WARPS) {
  const float* xr = a.x + row * a.ldx;
  Vec8 v[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const int c = (g * 32 + lane) * 8;
    if (c < a.D) {
      v[g] = load8_f32(xr + c);
      if (a.branch) {
        const Vec8 bf = load8_bf16(a.branch + row * a.D + c);
        if (a.drop_thresh16) {
          const Vec8 m = dropout_mult8(eff_seed(a), a.offset, static_cast<unsigned long long>(row * a.D + c) >> 3,
                                       a.drop_thresh16, a.drop_scale);
#pragma unroll
          for (int i = 0; i < 4; ++i) v[g].p[i] = ffma2(bf.p[i], m.p[i], v[g].p[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) v[g].p[i] = fadd2(v[g].p[i], bf.p[i]);
        }
        store8_f32(a.x_out + row * a.D + c, v[g]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[g].p[i] = zero;
    }
  }
  if (a.gamma == nullptr) continue;
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < NG; ++g) s += hsum8(v[g]);
  const float mean = warp_sum(s) / a.D;
  const f32x2 nmean = pk2(-mean, -mean);
  f32x2 sq2 = zero;
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const int c = (g * 32 + lane) * 8;
    if (c < a.D) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[g].p[i] = fadd2(v[g].p[i], nmean);  // centred values, reused by the normalisation
        sq2 = ffma2(v[g].p[i], v[g].p[i], sq2);
      }
    }
  }
  float sq;
  {
    float q0, q1;
    upk2(sq2, q0, q1);
    sq = q0 + q1;
  }
  const float rstd = rsqrtf(warp_sum(sq) / a.D + a.eps);
  if (lane == 0) { a.mean[row] = mean; a.rstd[row] = rstd; }
  const f32x2 rstd2 = pk2(rstd, rstd);
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const int c = (g * 32 + lane) * 8;
    if (c < a.D) {
      const Vec8 gm = ldg8_f32(a.gamma + c), bt = ldg8_f32(a.beta + c);
      Vec8 o;
#pragma unroll
      for (int i = 0; i < 4; ++i) o.p[i] = ffma2(fmul2(v[g].p[i], rstd2), gm.p[i], bt.p[i]);
      store8_bf16(a.y + row * a.D + c, o);
    }
  }
  }
}


// ------------------------------------------------------------------------------------------
// D % 128 == 0 (every ViT width here: 128 ... 1024): 4 consecutive elements per lane per group, so
// all 32 lanes carry the same load (with 8 per lane, D = 384 keeps half the warp idle in its
// second group), the per-lane state shrinks (12 instead of 16 values per array at D = 384) and
// more warps fit on an SM — these kernels are bound by bytes in flight, not by issue slots.
// ------------------------------------------------------------------------------------------
struct Vec4 {
  f32x2 p[2];
};
__device__ __forceinline__ Vec4 load4_f32(const float* ptr) {
  const float4 a = *reinterpret_cast<const float4*>(ptr);
  Vec4 v;
  v.p[0] = pk2(a.x, a.y); v.p[1] = pk2(a.z, a.w);
  return v;
}
__device__ __forceinline__ Vec4 ldg4_f32(const float* ptr) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(ptr));
  Vec4 v;
  v.p[0] = pk2(a.x, a.y); v.p[1] = pk2(a.z, a.w);
  return v;
}
__device__ __forceinline__ void store4_f32(float* ptr, const Vec4& v) {
  float4 a;
  upk2(v.p[0], a.x, a.y); upk2(v.p[1], a.z, a.w);
  *reinterpret_cast<float4*>(ptr) = a;
}
__device__ __forceinline__ Vec4 unpack4_bf16(const uint2 b) {
  Vec4 v;
  v.p[0] = pk2(bf16_lo(b.x), bf16_hi(b.x)); v.p[1] = pk2(bf16_lo(b.y), bf16_hi(b.y));
  return v;
}
__device__ __forceinline__ void store4_bf16(__nv_bfloat16* ptr, const Vec4& v) {
  uint2 pk;
  float a, b;
  upk2(v.p[0], a, b); pk.x = pack_bf16(a, b);
  upk2(v.p[1], a, b); pk.y = pack_bf16(a, b);
  *reinterpret_cast<uint2*>(ptr) = pk;
}
// keep-multipliers of elements e .. e+3 (e % 4 == 0): the half of dropout group e >> 3 they fall in
// (same generator call and the same per-element rule as dropout_scale8)
__device__ __forceinline__ Vec4 dropout_mult4(unsigned long long seed, unsigned long long offset,
                                              unsigned long long e, uint32_t thresh16, float scale) {
  const uint4 r = philox4x32<DROPOUT_PHILOX_ROUNDS>(seed, offset, e >> 3);
  const uint32_t lo = (e & 4) ? r.z : r.x, hi = (e & 4) ? r.w : r.y;
  const uint32_t th = thresh16 << 16;
  Vec4 v;
  v.p[0] = pk2(((lo << 16) >= th) ? scale : 0.0f, (lo >= th) ? scale : 0.0f);
  v.p[1] = pk2(((hi << 16) >= th) ? scale : 0.0f, (hi >= th) ? scale : 0.0f);
  return v;
}

template <int NG4>  // NG4 = D / 128
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd4_kernel(const LnFwdArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const f32x2 zero = pk2(0.f, 0.f);
  for (long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5); row < a.rows;
       row += static_cast<long long>(gridDim.x) * LN_WARPS) {
    const float* xr = a.x + row * a.ldx;
    Vec4 v[NG4];
    if (a.branch) {
      uint2 braw[NG4];
#pragma unroll
      for (int g = 0; g < NG4; ++g) {  // all loads of the row first
        const int c = (g * 32 + lane) * 4;
        v[g] = load4_f32(xr + c);
        braw[g] = *reinterpret_cast<const uint2*>(a.branch + row * a.D + c);
      }
#pragma unroll
      for (int g = 0; g < NG4; ++g) {
        const int c = (g * 32 + lane) * 4;
        const Vec4 bf = unpack4_bf16(braw[g]);
        if (a.drop_thresh16) {
          const Vec4 m = dropout_mult4(eff_seed(a), a.offset, static_cast<unsigned long long>(row * a.D + c),
                                       a.drop_thresh16, a.drop_scale);
          v[g].p[0] = ffma2(bf.p[0], m.p[0], v[g].p[0]);
          v[g].p[1] = ffma2(bf.p[1], m.p[1], v[g].p[1]);
        } else {
          v[g].p[0] = fadd2(v[g].p[0], bf.p[0]);
          v[g].p[1] = fadd2(v[g].p[1], bf.p[1]);
        }
        store4_f32(a.x_out + row * a.D + c, v[g]);
      }
    } else {
#pragma unroll
      for (int g = 0; g < NG4; ++g) v[g] = load4_f32(xr + (g * 32 + lane) * 4);
    }
    if (a.gamma == nullptr) continue;
    f32x2 s2 = zero;
#pragma unroll
    for (int g = 0; g < NG4; ++g) s2 = fadd2(s2, fadd2(v[g].p[0], v[g].p[1]));
    float s0, s1;
    upk2(s2, s0, s1);
    const float mean = warp_sum(s0 + s1) / a.D;
    const f32x2 nmean = pk2(-mean, -mean);
    f32x2 sq2 = zero;
#pragma unroll
    for (int g = 0; g < NG4; ++g) {
      v[g].p[0] = fadd2(v[g].p[0], nmean);
      v[g].p[1] = fadd2(v[g].p[1], nmean);
      sq2 = ffma2(v[g].p[0], v[g].p[0], sq2);
      sq2 = ffma2(v[g].p[1], v[g].p[1], sq2);
    }
    float q0, q1;
    upk2(sq2, q0, q1);
    const float rstd = rsqrtf(warp_sum(q0 + q1) / a.D + a.eps);
    if (lane == 0) { a.mean[row] = mean; a.rstd[row] = rstd; }
    const f32x2 rstd2 = pk2(rstd, rstd);
#pragma unroll
    for (int g = 0; g < NG4; ++g) {
      const int c = (g * 32 + lane) * 4;
      const Vec4 gm = ldg4_f32(a.gamma + c), bt = ldg4_f32(a.beta + c);
      Vec4 o;
      o.p[0] = ffma2(fmul2(v[g].p[0], rstd2), gm.p[0], bt.p[0]);
      o.p[1] = ffma2(fmul2(v[g].p[1], rstd2), gm.p[1], bt.p[1]);
      store4_bf16(a.y + row * a.D + c, o);
    }
  }
}

// Backward, same mapping. Per-lane state is the raw row (packed bf16 dy, x, dres): x_hat and dy*gamma
// are recomputed in the second phase instead of being kept, and the dgamma / dbeta partial sums live
// in a private shared-memory row per warp instead of 2 x 16 registers, so 8 CTAs (32 warps) fit on an
// SM where the 8-per-lane kernel has 4; the next row of the grid-stride loop is prefetched into L2
// while the current one is reduced.
constexpr int bwd4_ctas_per_sm(int ng4) { return ng4 <= 2 ? 8 : (ng4 <= 4 ? 6 : 4); }  // 64 / 85 / 128 registers

template <int NG4>
__global__ void __launch_bounds__(LN_WARPS * 32, bwd4_ctas_per_sm(NG4)) ln_bwd4_kernel(const LnBwdArgs a, const int prefetch) {
  __shared__ float4 acc_g[LN_WARPS][NG4 * 32];
  __shared__ float4 acc_b[LN_WARPS][NG4 * 32];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const f32x2 zero = pk2(0.f, 0.f);
  const bool stats = a.dy != nullptr && a.dgamma != nullptr;
  if (stats) {
#pragma unroll
    for (int g = 0; g < NG4; ++g) {
      acc_g[warp][g * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
      acc_b[warp][g * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const long long stride = static_cast<long long>(gridDim.x) * LN_WARPS;
  for (long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + warp; row < a.rows; row += stride) {
    if (prefetch && row + stride < a.rows) {
      // one 128-byte line per lane: NG4*4 lines of dres, NG4*4 of x, NG4*2 of dy
      const long long nr = row + stride;
      const int l = lane;
      if (a.dres && l < NG4 * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.dres + nr * a.ld_dres + l * 32));
      if (a.dy) {
        if (l < NG4 * 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.x + nr * a.ldx + l * 32));
        if (l < NG4 * 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.dy + nr * a.D + l * 64));
      }
    }
    Vec4 res[NG4], xs[NG4];
    uint2 draw[NG4];
    float mean = 0.f, rstd = 0.f;
    if (a.dy) { mean = a.mean[row]; rstd = a.rstd[row]; }
#pragma unroll
    for (int g = 0; g < NG4; ++g) {  // every load of the row is issued before the first use
      const int c = (g * 32 + lane) * 4;
      if (a.dres) res[g] = load4_f32(a.dres + row * a.ld_dres + c);
      if (a.dy) {
        draw[g] = *reinterpret_cast<const uint2*>(a.dy + row * a.D + c);
        xs[g] = load4_f32(a.x + row * a.ldx + c);
      }
    }
    const f32x2 rstd2 = pk2(rstd, rstd), nmr = pk2(-mean * rstd, -mean * rstd);
    f32x2 ka = zero, kb = zero, kc = zero;  // dx = ka * (dy*gamma) + kb + kc * x_hat
    if (a.dy) {
      f32x2 c1v = zero, c2v = zero;
#pragma unroll
      for (int g = 0; g < NG4; ++g) {
        const int c = (g * 32 + lane) * 4;
        const Vec4 d = unpack4_bf16(draw[g]);
        const Vec4 gm = ldg4_f32(a.gamma + c);
        f32x2 xh[2], dg[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          xh[i] = ffma2(xs[g].p[i], rstd2, nmr);  // (x - mean) * rstd
          xs[g].p[i] = xh[i];                      // keep x_hat in place of x
          const f32x2 dyg = fmul2(d.p[i], gm.p[i]);
          c1v = fadd2(c1v, dyg);
          c2v = ffma2(dyg, xh[i], c2v);
          dg[i] = fmul2(d.p[i], xh[i]);
        }
        if (stats) {
          float4 tg = acc_g[warp][g * 32 + lane], tb = acc_b[warp][g * 32 + lane];
          float e0, e1, e2, e3, f0, f1, f2, f3;
          upk2(dg[0], e0, e1); upk2(dg[1], e2, e3);
          upk2(d.p[0], f0, f1); upk2(d.p[1], f2, f3);
          tg.x += e0; tg.y += e1; tg.z += e2; tg.w += e3;
          tb.x += f0; tb.y += f1; tb.z += f2; tb.w += f3;
          acc_g[warp][g * 32 + lane] = tg;
          acc_b[warp][g * 32 + lane] = tb;
        }
      }
      float s1a, s1b, s2a, s2b;
      upk2(c1v, s1a, s1b);
      upk2(c2v, s2a, s2b);
      const float c1 = warp_sum(s1a + s1b) / a.D, c2 = warp_sum(s2a + s2b) / a.D;
      ka = rstd2; kb = pk2(-rstd * c1, -rstd * c1); kc = pk2(-rstd * c2, -rstd * c2);
    }
#pragma unroll
    for (int g = 0; g < NG4; ++g) {
      const int c = (g * 32 + lane) * 4;
      Vec4 o;
      if (a.dy) {
        const Vec4 d = unpack4_bf16(draw[g]);
        const Vec4 gm = ldg4_f32(a.gamma + c);
        o.p[0] = ffma2(xs[g].p[0], kc, ffma2(fmul2(d.p[0], gm.p[0]), ka, kb));
        o.p[1] = ffma2(xs[g].p[1], kc, ffma2(fmul2(d.p[1], gm.p[1]), ka, kb));
      } else {
        o.p[0] = zero; o.p[1] = zero;
      }
      if (a.dres) { o.p[0] = fadd2(o.p[0], res[g].p[0]); o.p[1] = fadd2(o.p[1], res[g].p[1]); }
      if (a.dx) store4_f32(a.dx + row * a.ld_dx + c, o);
      if (a.dbranch) {
        if (a.drop_thresh16) {
          const Vec4 m = dropout_mult4(eff_seed(a), a.offset, static_cast<unsigned long long>(row * a.D + c),
                                       a.drop_thresh16, a.drop_scale);
          o.p[0] = fmul2(o.p[0], m.p[0]);
          o.p[1] = fmul2(o.p[1], m.p[1]);
        }
        store4_bf16(a.dbranch + row * a.D + c, o);
      }
    }
  }
  if (!stats) return;
  // cross-warp reduction of the column partials, then one atomic per column per CTA
  __syncthreads();
  const float* pg = reinterpret_cast<const float*>(&acc_g[0][0]);
  const float* pb = reinterpret_cast<const float*>(&acc_b[0][0]);
  for (int c = threadIdx.x; c < a.D; c += LN_WARPS * 32) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) { sg += pg[w * NG4 * 128 + c]; sb += pb[w * NG4 * 128 + c]; }
    atomicAdd(a.dgamma + c, sg);
    atomicAdd(a.dbeta + c, sb);
  }
}

// generic shapes (D % 8 != 0 or D > 1024): one warp per row, three passes served by L1/L2
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_generic_kernel(const LnFwdArgs a) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const float* xr = a.x + row * a.ldx;
  if (a.branch) {
    for (int c = lane; c < a.D; c += 32) {
      float b = __bfloat162float(a.branch[row * a.D + c]);
      if (a.drop_thresh16) {
        const unsigned long long e = static_cast<unsigned long long>(row * a.D + c);
        const uint32_t keep = dropout_keep8(eff_seed(a), a.offset, e >> 3, a.drop_thresh16);
        b = ((keep >> (e & 7)) & 1u) ? b * a.drop_scale : 0.f;
      }
      a.x_out[row * a.D + c] = xr[c] + b;
    }
    __syncwarp();
    xr = a.x_out + row * a.D;
  }
  if (a.gamma == nullptr) return;
  float s = 0.f;
  for (int c = lane; c < a.D; c += 32) s += xr[c];
  const float mean = warp_sum(s) / a.D;
  float sq = 0.f;
  for (int c = lane; c < a.D; c += 32) { const float d = xr[c] - mean; sq += d * d; }
  const float rstd = rsqrtf(warp_sum(sq) / a.D + a.eps);
  if (lane == 0) { a.mean[row] = mean; a.rstd[row] = rstd; }
  for (int c = lane; c < a.D; c += 32)
    a.y[row * a.D + c] = __float2bfloat16_rn((xr[c] - mean) * rstd * a.gamma[c] + a.beta[c]);
}


template <int NG>
__global__ void __launch_bounds__(LN_WARPS * 32, NG <= 2 ? 4 : 2) ln_bwd_kernel(const LnBwdArgs a) {
  __shared__ float red[LN_WARPS][NG * 256 + 8];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const f32x2 zero = pk2(0.f, 0.f);
  Vec8 dg[NG], db[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int i = 0; i < 4; ++i) { dg[g].p[i] = zero; db[g].p[i] = zero; }

  for (long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + warp; row < a.rows;
       row += static_cast<long long>(gridDim.x) * LN_WARPS) {
    Vec8 dyv[NG], xh[NG], res[NG];
    f32x2 c1v = zero, c2v = zero;
    float mean = 0.f, rstd = 0.f;
    if (a.dy) { mean = a.mean[row]; rstd = a.rstd[row]; }
    const f32x2 rstd2 = pk2(rstd, rstd), nmr = pk2(-mean * rstd, -mean * rstd);
    // every load of the row is issued before the first reduction
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int c = (g * 32 + lane) * 8;
      if (c < a.D) {
        if (a.dres) res[g] = load8_f32(a.dres + row * a.ld_dres + c);
        if (a.dy) {
          const Vec8 d = load8_bf16(a.dy + row * a.D + c);
          const Vec8 xs = load8_f32(a.x + row * a.ldx + c);
          const Vec8 gm = ldg8_f32(a.gamma + c);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            xh[g].p[i] = ffma2(xs.p[i], rstd2, nmr);          // (x - mean) * rstd
            dg[g].p[i] = ffma2(d.p[i], xh[g].p[i], dg[g].p[i]);
            db[g].p[i] = fadd2(db[g].p[i], d.p[i]);
            dyv[g].p[i] = fmul2(d.p[i], gm.p[i]);
            c1v = fadd2(c1v, dyv[g].p[i]);
            c2v = ffma2(dyv[g].p[i], xh[g].p[i], c2v);
          }
        }
      }
    }
    f32x2 ka = zero, kb = zero, kc = zero;  // dx = ka * dyv + kb + kc * xh
    if (a.dy) {
      float s1a, s1b, s2a, s2b;
      upk2(c1v, s1a, s1b);
      upk2(c2v, s2a, s2b);
      const float c1 = warp_sum(s1a + s1b) / a.D, c2 = warp_sum(s2a + s2b) / a.D;
      ka = rstd2; kb = pk2(-rstd * c1, -rstd * c1); kc = pk2(-rstd * c2, -rstd * c2);
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int c = (g * 32 + lane) * 8;
      if (c < a.D) {
        Vec8 o;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          o.p[i] = a.dy ? ffma2(xh[g].p[i], kc, ffma2(dyv[g].p[i], ka, kb)) : zero;
          if (a.dres) o.p[i] = fadd2(o.p[i], res[g].p[i]);
        }
        if (a.dx) store8_f32(a.dx + row * a.ld_dx + c, o);
        if (a.dbranch) {
          if (a.drop_thresh16) {
            const Vec8 m = dropout_mult8(eff_seed(a), a.offset, static_cast<unsigned long long>(row * a.D + c) >> 3,
                                         a.drop_thresh16, a.drop_scale);
#pragma unroll
            for (int i = 0; i < 4; ++i) o.p[i] = fmul2(o.p[i], m.p[i]);
          }
          store8_bf16(a.dbranch + row * a.D + c, o);
        }
      }
    }
  }
  if (a.dy == nullptr || a.dgamma == nullptr) return;
  // cross-warp reduction of the per-lane column partials, then one atomic per column per CTA
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int g = 0; g < NG; ++g)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float lo, hi;
        upk2(pass == 0 ? dg[g].p[i] : db[g].p[i], lo, hi);
        red[warp][(g * 32 + lane) * 8 + 2 * i] = lo;
        red[warp][(g * 32 + lane) * 8 + 2 * i + 1] = hi;
      }
    __syncthreads();
    for (int c = threadIdx.x; c < a.D; c += LN_WARPS * 32) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < LN_WARPS; ++w) s += red[w][c];
      atomicAdd((pass == 0 ? a.dgamma : a.dbeta) + c, s);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_generic_kernel(const LnBwdArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + warp; row < a.rows;
       row += static_cast<long long>(gridDim.x) * LN_WARPS) {
    float c1 = 0.f, c2 = 0.f, mean = 0.f, rstd = 0.f;
    if (a.dy) {
      mean = a.mean[row]; rstd = a.rstd[row];
      for (int c = lane; c < a.D; c += 32) {
        const float d = __bfloat162float(a.dy[row * a.D + c]);
        const float xh = (a.x[row * a.ldx + c] - mean) * rstd;
        const float dyg = d * a.gamma[c];
        c1 += dyg; c2 += dyg * xh;
        if (a.dgamma) { atomicAdd(a.dgamma + c, d * xh); atomicAdd(a.dbeta + c, d); }
      }
      c1 = warp_sum(c1) / a.D; c2 = warp_sum(c2) / a.D;
    }
    for (int c = lane; c < a.D; c += 32) {
      float o = 0.f;
      if (a.dy) {
        const float d = __bfloat162float(a.dy[row * a.D + c]) * a.gamma[c];
        const float xh = (a.x[row * a.ldx + c] - mean) * rstd;
        o = rstd * (d - c1 - xh * c2);
      }
      if (a.dres) o += a.dres[row * a.ld_dres + c];
      if (a.dx) a.dx[row * a.ld_dx + c] = o;
      if (a.dbranch) {
        if (a.drop_thresh16) {
          const unsigned long long e = static_cast<unsigned long long>(row * a.D + c);
          const uint32_t keep = dropout_keep8(eff_seed(a), a.offset, e >> 3, a.drop_thresh16);
          o = ((keep >> (e & 7)) & 1u) ? o * a.drop_scale : 0.f;
        }
        a.dbranch[row * a.D + c] = __float2bfloat16_rn(o);
      }
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace vitssl

using namespace vitssl;

extern "C" int vitssl_add_layernorm_fwd(const float* x, int64_t ldx, const void* branch,
                                        float* x_out, const float* gamma, const float* beta,
                                        void* y, float* mean, float* rstd, int64_t rows, int64_t D,
                                        float eps, float dropout_p, uint64_t philox_seed,
                                        uint64_t philox_offset, cudaStream_t stream) {
  VITSSL_REQUIRE(x != nullptr && rows >= 0 && D > 0, VITSSL_ERR_ARG, "add_layernorm_fwd: bad args");
  VITSSL_REQUIRE((branch == nullptr) == (x_out == nullptr), VITSSL_ERR_ARG,
                 "add_layernorm_fwd: branch and x_out go together");
  VITSSL_REQUIRE((gamma == nullptr) == (beta == nullptr), VITSSL_ERR_ARG,
                 "add_layernorm_fwd: gamma and beta go together");
  if (gamma) VITSSL_REQUIRE(y && mean && rstd, VITSSL_ERR_ARG, "add_layernorm_fwd: LN outputs missing");
  VITSSL_REQUIRE(gamma || branch, VITSSL_ERR_ARG, "add_layernorm_fwd: nothing to do");
  VITSSL_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, VITSSL_ERR_ARG, "dropout_p out of range");
  if (rows == 0) return 0;
  LnFwdArgs a{};
  a.x = x; a.ldx = ldx; a.branch = reinterpret_cast<const __nv_bfloat16*>(branch);
  a.x_out = x_out; a.gamma = gamma; a.beta = beta; a.y = reinterpret_cast<__nv_bfloat16*>(y);
  a.mean = mean; a.rstd = rstd; a.rows = rows; a.D = (int)D; a.eps = eps;
  a.drop_thresh16 = static_cast<uint32_t>(dropout_p * 65536.0f);
  a.drop_scale = 1.0f / (1.0f - dropout_p); a.seed = philox_seed; a.offset = philox_offset; a.seed_slot = seed_slot_override();
  const long long want_f = (rows + LN_WARPS - 1) / LN_WARPS;
  // one warp per row, no grid cap: measured faster than a capped grid-stride launch (44 vs 53 us
  // for 50176 x 384) — the row loop only matters beyond 2^31 / 4 rows
  const long long cap_f = (1ll << 31) - 1;
  const unsigned grid = (unsigned)(want_f < cap_f ? want_f : cap_f);
  const bool fast = (D % 8 == 0) && D <= 1024 && (ldx % 4 == 0) && aligned16(x) &&
                    aligned16(branch) && aligned16(x_out) && aligned16(y) && aligned16(gamma) &&
                    aligned16(beta);
  static const int impl = getenv("VITSSL_LN_IMPL") ? atoi(getenv("VITSSL_LN_IMPL")) : 1;  // 0: 8-per-lane kernels only
  const int ng4 = (int)(D / 128);
  if (fast && impl != 0 && D % 128 == 0 && (ng4 <= 4 || ng4 == 6 || ng4 == 8)) {
    switch (ng4) {
      case 1: launch_pdl(ln_fwd4_kernel<1>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a); break;
      case 2: launch_pdl(ln_fwd4_kernel<2>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a); break;
      case 3: launch_pdl(ln_fwd4_kernel<3>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a); break;
      case 4: launch_pdl(ln_fwd4_kernel<4>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a); break;
      case 6: launch_pdl(ln_fwd4_kernel<6>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a); break;
      default: launch_pdl(ln_fwd4_kernel<8>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a); break;
    }
  } else if (fast) {
    const int ng = (int)((D + 255) / 256);
    switch (ng) {
      case 1: launch_pdl(ln_fwd_kernel<1>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a); break;
      case 2: launch_pdl(ln_fwd_kernel<2>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a); break;
      case 3: launch_pdl(ln_fwd_kernel<3>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a); break;
      default: launch_pdl(ln_fwd_kernel<4>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a); break;
    }
  } else {
    ln_fwd_generic_kernel<<<(unsigned)want_f, LN_WARPS * 32, 0, stream>>>(a);
  }
  return check_launch("add_layernorm_fwd");
}

extern "C" int vitssl_add_layernorm_bwd(const void* dy, const float* x, int64_t ldx,
                                        const float* mean, const float* rstd, const float* gamma,
                                        const float* dres, int64_t ld_dres, float* dx,
                                        int64_t ld_dx, void* dbranch, float* dgamma, float* dbeta,
                                        int64_t rows, int64_t D, float dropout_p,
                                        uint64_t philox_seed, uint64_t philox_offset,
                                        cudaStream_t stream) {
  if (dgamma && dbeta && D > 0) {
    cudaMemsetAsync(dgamma, 0, D * sizeof(float), stream);
    cudaMemsetAsync(dbeta, 0, D * sizeof(float), stream);
  }
  return vitssl_add_layernorm_bwd_acc(dy, x, ldx, mean, rstd, gamma, dres, ld_dres, dx, ld_dx, dbranch, dgamma,
                                      dbeta, rows, D, dropout_p, philox_seed, philox_offset, stream);
}

extern "C" int vitssl_add_layernorm_bwd_acc(const void* dy, const float* x, int64_t ldx,
                                            const float* mean, const float* rstd, const float* gamma,
                                            const float* dres, int64_t ld_dres, float* dx,
                                            int64_t ld_dx, void* dbranch, float* dgamma, float* dbeta,
                                            int64_t rows, int64_t D, float dropout_p,
                                            uint64_t philox_seed, uint64_t philox_offset,
                                            cudaStream_t stream) {
  VITSSL_REQUIRE(rows >= 0 && D > 0, VITSSL_ERR_ARG, "add_layernorm_bwd: bad args");
  if (dy) VITSSL_REQUIRE(x && mean && rstd && gamma, VITSSL_ERR_ARG, "add_layernorm_bwd: LN inputs missing");
  VITSSL_REQUIRE(dy || dres, VITSSL_ERR_ARG, "add_layernorm_bwd: no incoming gradient");
  VITSSL_REQUIRE(dx || dbranch, VITSSL_ERR_ARG, "add_layernorm_bwd: no output requested");
  VITSSL_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), VITSSL_ERR_ARG, "dgamma/dbeta go together");
  if (rows == 0) return 0;
  LnBwdArgs a{};
  a.dy = reinterpret_cast<const __nv_bfloat16*>(dy); a.x = x; a.ldx = ldx; a.mean = mean;
  a.rstd = rstd; a.gamma = gamma; a.dres = dres; a.ld_dres = ld_dres; a.dx = dx; a.ld_dx = ld_dx;
  a.dbranch = reinterpret_cast<__nv_bfloat16*>(dbranch); a.dgamma = dgamma; a.dbeta = dbeta;
  a.rows = rows; a.D = (int)D;
  a.drop_thresh16 = static_cast<uint32_t>(dropout_p * 65536.0f);
  a.drop_scale = 1.0f / (1.0f - dropout_p); a.seed = philox_seed; a.offset = philox_offset; a.seed_slot = seed_slot_override();
  long long want = (rows + LN_WARPS - 1) / LN_WARPS;
  static const int impl = getenv("VITSSL_LN_IMPL") ? atoi(getenv("VITSSL_LN_IMPL")) : 1;  // 0: 8-per-lane kernels only
  const int ng4 = (int)(D / 128);
  const bool use4 = impl != 0 && D % 128 == 0 && (ng4 <= 4 || ng4 == 6 || ng4 == 8);
  // one resident wave: 4 CTAs/SM for the 8-per-lane kernel (128 registers), 8 / 6 / 4 for the 4-per-lane one
  static const int cap_env = getenv("VITSSL_LN_BWD_CAP") ? atoi(getenv("VITSSL_LN_BWD_CAP")) : 0;
  static const int pf_env = getenv("VITSSL_LN_BWD_PREFETCH") ? atoi(getenv("VITSSL_LN_BWD_PREFETCH")) : 1;
  const int cap_mult = cap_env > 0 ? cap_env : (use4 ? bwd4_ctas_per_sm(ng4) : 4);
  const long long cap = static_cast<long long>(num_sms()) * cap_mult;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  const bool fast = (D % 8 == 0) && D <= 1024 && (ldx % 4 == 0) && (ld_dres % 4 == 0) &&
                    (ld_dx % 4 == 0) && aligned16(dy) && aligned16(x) && aligned16(dres) &&
                    aligned16(dx) && aligned16(dbranch) && aligned16(gamma);
  if (fast && use4) {
    switch (ng4) {
      case 1: launch_pdl(ln_bwd4_kernel<1>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a, pf_env); break;
      case 2: launch_pdl(ln_bwd4_kernel<2>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a, pf_env); break;
      case 3: launch_pdl(ln_bwd4_kernel<3>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a, pf_env); break;
      case 4: launch_pdl(ln_bwd4_kernel<4>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a, pf_env); break;
      case 6: launch_pdl(ln_bwd4_kernel<6>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a, pf_env); break;
      default: launch_pdl(ln_bwd4_kernel<8>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a, pf_env); break;
    }
  } else if (fast) {
    const int ng = (int)((D + 255) / 256);
    switch (ng) {
      case 1: launch_pdl(ln_bwd_kernel<1>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a); break;
      case 2: launch_pdl(ln_bwd_kernel<2>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a); break;
      case 3: launch_pdl(ln_bwd_kernel<3>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a); break;
      default: launch_pdl(ln_bwd_kernel<4>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, a); break;
    }
  } else {
    ln_bwd_generic_kernel<<<grid, LN_WARPS * 32, 0, stream>>>(a);
  }
  return check_launch("add_layernorm_bwd");
}
