// vitssl_b200 — fused residual-add (+dropout) + LayerNorm, forward and backward.
// Reference: encoder_block.py:40-52 (x = drop(branch) + residual; x = layer_norm(x)), LayerNorm
// eps 1e-5 with affine (encoder_block.py:26-27, mlp_head.py:9). The reference under autocast keeps
// the residual stream and the LayerNorm in fp32 and feeds bf16 to the next linear.
//
// HBM-bound, single pass: one warp owns a row, the row lives in registers (8 contiguous elements
// per lane per step -> 2x float4 loads of the stream, 1x 16-byte load of the bf16 branch).
//   fwd traffic / element: 4 (x) + 2 (branch) + 4 (x_out) + 2 (y)            = 12 B
//   bwd traffic / element: 2 (dy) + 4 (x) + 4 (dres) + 4 (dx) + 2 (dbranch) = 16 B
#include "common.cuh"
#include "vitssl_b200.h"

namespace vitssl {
namespace {

constexpr int LN_WARPS = 4;

struct LnFwdArgs {
  const float* x; long long ldx;          // input stream rows (pitch in elements)
  const __nv_bfloat16* branch;             // nullable, [rows, D] dense
  float* x_out;                            // nullable (required iff branch), dense
  const float* gamma; const float* beta;   // nullable together -> no LayerNorm, add only
  __nv_bfloat16* y;                        // LN output (bf16), dense
  float* mean; float* rstd;                // [rows]
  long long rows; int D; float eps;
  uint32_t drop_thresh16; float drop_scale; unsigned long long seed, offset;
};

template <int NG>  // NG = ceil(D / 256), D % 8 == 0
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(const LnFwdArgs a) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  const float* xr = a.x + row * a.ldx;
  float v[NG][8];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const int c = (g * 32 + lane) * 8;
    if (c < a.D) {
      const float4 p0 = *reinterpret_cast<const float4*>(xr + c);
      const float4 p1 = *reinterpret_cast<const float4*>(xr + c + 4);
      v[g][0] = p0.x; v[g][1] = p0.y; v[g][2] = p0.z; v[g][3] = p0.w;
      v[g][4] = p1.x; v[g][5] = p1.y; v[g][6] = p1.z; v[g][7] = p1.w;
      if (a.branch) {
        const uint4 b = *reinterpret_cast<const uint4*>(a.branch + row * a.D + c);
        float bf[8] = {bf16_lo(b.x), bf16_hi(b.x), bf16_lo(b.y), bf16_hi(b.y),
                       bf16_lo(b.z), bf16_hi(b.z), bf16_lo(b.w), bf16_hi(b.w)};
        if (a.drop_thresh16) {
          const uint32_t keep = dropout_keep8(a.seed, a.offset,
                                              static_cast<unsigned long long>(row * a.D + c) >> 3,
                                              a.drop_thresh16);
#pragma unroll
          for (int i = 0; i < 8; ++i) bf[i] = ((keep >> i) & 1u) ? bf[i] * a.drop_scale : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[g][i] += bf[i];
        float* xo = a.x_out + row * a.D + c;
        *reinterpret_cast<float4*>(xo) = make_float4(v[g][0], v[g][1], v[g][2], v[g][3]);
        *reinterpret_cast<float4*>(xo + 4) = make_float4(v[g][4], v[g][5], v[g][6], v[g][7]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[g][i] = 0.f;
    }
  }
  if (a.gamma == nullptr) return;
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[g][i];
  const float mean = warp_sum(s) / a.D;
  float sq = 0.f;
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const int c = (g * 32 + lane) * 8;
    if (c < a.D) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[g][i] - mean; sq += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / a.D + a.eps);
  if (lane == 0) { a.mean[row] = mean; a.rstd[row] = rstd; }
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const int c = (g * 32 + lane) * 8;
    if (c < a.D) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.gamma + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.beta + c));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.beta + c + 4));
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (v[g][i] - mean) * rstd * gm[i] + bt[i];
      uint4 pk;
      pk.x = pack_bf16(o[0], o[1]); pk.y = pack_bf16(o[2], o[3]);
      pk.z = pack_bf16(o[4], o[5]); pk.w = pack_bf16(o[6], o[7]);
      *reinterpret_cast<uint4*>(a.y + row * a.D + c) = pk;
    }
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
        const uint32_t keep = dropout_keep8(a.seed, a.offset, e >> 3, a.drop_thresh16);
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
};

template <int NG>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_bwd_kernel(const LnBwdArgs a) {
  __shared__ float red[LN_WARPS][NG * 256 + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float dg[NG][8], db[NG][8];
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int i = 0; i < 8; ++i) { dg[g][i] = 0.f; db[g][i] = 0.f; }

  for (long long row = static_cast<long long>(blockIdx.x) * LN_WARPS + warp; row < a.rows;
       row += static_cast<long long>(gridDim.x) * LN_WARPS) {
    float dyv[NG][8], xh[NG][8];
    float c1 = 0.f, c2 = 0.f;
    float mean = 0.f, rstd = 0.f;
    if (a.dy) { mean = a.mean[row]; rstd = a.rstd[row]; }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int c = (g * 32 + lane) * 8;
      if (a.dy && c < a.D) {
        const uint4 p = *reinterpret_cast<const uint4*>(a.dy + row * a.D + c);
        const float d[8] = {bf16_lo(p.x), bf16_hi(p.x), bf16_lo(p.y), bf16_hi(p.y),
                            bf16_lo(p.z), bf16_hi(p.z), bf16_lo(p.w), bf16_hi(p.w)};
        const float4 x0 = *reinterpret_cast<const float4*>(a.x + row * a.ldx + c);
        const float4 x1 = *reinterpret_cast<const float4*>(a.x + row * a.ldx + c + 4);
        const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma + c));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(a.gamma + c + 4));
        const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[g][i] = (xs[i] - mean) * rstd;
          dg[g][i] += d[i] * xh[g][i];
          db[g][i] += d[i];
          dyv[g][i] = d[i] * gm[i];
          c1 += dyv[g][i];
          c2 += dyv[g][i] * xh[g][i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { dyv[g][i] = 0.f; xh[g][i] = 0.f; }
      }
    }
    if (a.dy) {
      c1 = warp_sum(c1) / a.D;
      c2 = warp_sum(c2) / a.D;
    }
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      const int c = (g * 32 + lane) * 8;
      if (c < a.D) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = a.dy ? rstd * (dyv[g][i] - c1 - xh[g][i] * c2) : 0.f;
        if (a.dres) {
          const float4 r0 = *reinterpret_cast<const float4*>(a.dres + row * a.ld_dres + c);
          const float4 r1 = *reinterpret_cast<const float4*>(a.dres + row * a.ld_dres + c + 4);
          o[0] += r0.x; o[1] += r0.y; o[2] += r0.z; o[3] += r0.w;
          o[4] += r1.x; o[5] += r1.y; o[6] += r1.z; o[7] += r1.w;
        }
        if (a.dx) {
          float* dxp = a.dx + row * a.ld_dx + c;
          *reinterpret_cast<float4*>(dxp) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(dxp + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
        if (a.dbranch) {
          if (a.drop_thresh16) {
            const uint32_t keep = dropout_keep8(
                a.seed, a.offset, static_cast<unsigned long long>(row * a.D + c) >> 3,
                a.drop_thresh16);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = ((keep >> i) & 1u) ? o[i] * a.drop_scale : 0.f;
          }
          uint4 pk;
          pk.x = pack_bf16(o[0], o[1]); pk.y = pack_bf16(o[2], o[3]);
          pk.z = pack_bf16(o[4], o[5]); pk.w = pack_bf16(o[6], o[7]);
          *reinterpret_cast<uint4*>(a.dbranch + row * a.D + c) = pk;
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
      for (int i = 0; i < 8; ++i)
        red[warp][(g * 32 + lane) * 8 + i] = pass == 0 ? dg[g][i] : db[g][i];
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
          const uint32_t keep = dropout_keep8(a.seed, a.offset, e >> 3, a.drop_thresh16);
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
  a.drop_scale = 1.0f / (1.0f - dropout_p); a.seed = philox_seed; a.offset = philox_offset;
  const unsigned grid = (unsigned)((rows + LN_WARPS - 1) / LN_WARPS);
  const bool fast = (D % 8 == 0) && D <= 1024 && (ldx % 4 == 0) && aligned16(x) &&
                    aligned16(branch) && aligned16(x_out) && aligned16(y) && aligned16(gamma) &&
                    aligned16(beta);
  if (fast) {
    const int ng = (int)((D + 255) / 256);
    switch (ng) {
      case 1: ln_fwd_kernel<1><<<grid, LN_WARPS * 32, 0, stream>>>(a); break;
      case 2: ln_fwd_kernel<2><<<grid, LN_WARPS * 32, 0, stream>>>(a); break;
      case 3: ln_fwd_kernel<3><<<grid, LN_WARPS * 32, 0, stream>>>(a); break;
      default: ln_fwd_kernel<4><<<grid, LN_WARPS * 32, 0, stream>>>(a); break;
    }
  } else {
    ln_fwd_generic_kernel<<<grid, LN_WARPS * 32, 0, stream>>>(a);
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
  VITSSL_REQUIRE(rows >= 0 && D > 0, VITSSL_ERR_ARG, "add_layernorm_bwd: bad args");
  if (dy) VITSSL_REQUIRE(x && mean && rstd && gamma, VITSSL_ERR_ARG, "add_layernorm_bwd: LN inputs missing");
  VITSSL_REQUIRE(dy || dres, VITSSL_ERR_ARG, "add_layernorm_bwd: no incoming gradient");
  VITSSL_REQUIRE(dx || dbranch, VITSSL_ERR_ARG, "add_layernorm_bwd: no output requested");
  VITSSL_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), VITSSL_ERR_ARG, "dgamma/dbeta go together");
  if (dgamma) {
    cudaMemsetAsync(dgamma, 0, D * sizeof(float), stream);
    cudaMemsetAsync(dbeta, 0, D * sizeof(float), stream);
  }
  if (rows == 0) return 0;
  LnBwdArgs a{};
  a.dy = reinterpret_cast<const __nv_bfloat16*>(dy); a.x = x; a.ldx = ldx; a.mean = mean;
  a.rstd = rstd; a.gamma = gamma; a.dres = dres; a.ld_dres = ld_dres; a.dx = dx; a.ld_dx = ld_dx;
  a.dbranch = reinterpret_cast<__nv_bfloat16*>(dbranch); a.dgamma = dgamma; a.dbeta = dbeta;
  a.rows = rows; a.D = (int)D;
  a.drop_thresh16 = static_cast<uint32_t>(dropout_p * 65536.0f);
  a.drop_scale = 1.0f / (1.0f - dropout_p); a.seed = philox_seed; a.offset = philox_offset;
  long long want = (rows + LN_WARPS - 1) / LN_WARPS;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  const bool fast = (D % 8 == 0) && D <= 1024 && (ldx % 4 == 0) && (ld_dres % 4 == 0) &&
                    (ld_dx % 4 == 0) && aligned16(dy) && aligned16(x) && aligned16(dres) &&
                    aligned16(dx) && aligned16(dbranch) && aligned16(gamma);
  if (fast) {
    const int ng = (int)((D + 255) / 256);
    switch (ng) {
      case 1: ln_bwd_kernel<1><<<grid, LN_WARPS * 32, 0, stream>>>(a); break;
      case 2: ln_bwd_kernel<2><<<grid, LN_WARPS * 32, 0, stream>>>(a); break;
      case 3: ln_bwd_kernel<3><<<grid, LN_WARPS * 32, 0, stream>>>(a); break;
      default: ln_bwd_kernel<4><<<grid, LN_WARPS * 32, 0, stream>>>(a); break;
    }
  } else {
    ln_bwd_generic_kernel<<<grid, LN_WARPS * 32, 0, stream>>>(a);
  }
  return check_launch("add_layernorm_bwd");
}
