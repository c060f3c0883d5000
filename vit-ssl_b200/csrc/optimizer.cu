// vitssl_b200 — fused multi-tensor AdamW step.
//
// Reference: the trainers' `scaler.step(optimizer)` on `torch.optim.AdamW`
// (utils/train_utils.py:25-29, utils/trainers/simmim_trainer.py:69-71, base_trainer.py:44), followed
// on our side by the fp32 -> bf16 re-cast of the GEMM weight shadows. One pass over
// (param, grad, exp_avg, exp_avg_sq) does, per element,
//     g      = grad / grad_scale                      (GradScaler unscale, folded in)
//     skip everything when found_inf != 0             (GradScaler's skipped step)
//     p     -= lr * weight_decay * p                  (decoupled weight decay)
//     m      = m + (g - m) * (1 - beta1)              (torch's lerp form)
//     v      = beta2 * v + (1 - beta2) * g * g
//     p     -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
//     shadow = bf16(p)                                (optional: the GEMM operand copy)
// i.e. exactly torch's fused AdamW arithmetic (ATen/native/cuda/fused_adam_utils.cuh), 28 B of
// reads + 12 (+2) B of writes per parameter: HBM-bound. Step counters live on the device (one fp32
// per tensor, as torch keeps them) so that a step skipped by found_inf does not advance t and no
// host synchronisation is needed.
#include "common.cuh"
#include "vitssl_b200.h"

namespace vitssl {
namespace {

constexpr int AD_MAX = 36;      // tensors per launch (kernel parameter space)
constexpr int AD_CHUNK = 2048;  // elements per CTA

struct AdamArgs {
  float* p[AD_MAX];
  float* g[AD_MAX];
  float* m[AD_MAX];
  float* v[AD_MAX];
  __nv_bfloat16* sh[AD_MAX];  // nullable
  float* step[AD_MAX];        // device step counter of the tensor (already advanced for this step)
  long long n[AD_MAX];
  int block_start[AD_MAX + 1];
  int count;
};

struct AdamHyper {
  float lr, beta1, beta2, eps, weight_decay;
  const float* grad_scale;  // nullable device scalar
  const float* found_inf;   // nullable device scalar
};

// step[t] += 1 unless the GradScaler found an inf (runs before the update kernel of the same step)
__global__ void adamw_advance_steps_kernel(const __grid_constant__ AdamArgs a, const float* found_inf) {
  const int t = threadIdx.x;
  if (t < a.count && !(found_inf && *found_inf != 0.f)) *a.step[t] += 1.0f;
}

__global__ void __launch_bounds__(256) adamw_kernel(const __grid_constant__ AdamArgs a, const AdamHyper h) {
  if (h.found_inf && *h.found_inf != 0.f) return;
  int t = 0;
  while (t + 1 < a.count && a.block_start[t + 1] <= static_cast<int>(blockIdx.x)) ++t;
  const long long base = static_cast<long long>(blockIdx.x - a.block_start[t]) * AD_CHUNK;
  float* __restrict__ p = a.p[t];
  float* __restrict__ g = a.g[t];
  float* __restrict__ m = a.m[t];
  float* __restrict__ v = a.v[t];
  __nv_bfloat16* __restrict__ sh = a.sh[t];
  const long long n = a.n[t];
  const float step = *a.step[t];
  const float inv_scale = h.grad_scale ? 1.0f / *h.grad_scale : 1.0f;
  const float bc1 = 1.0f - powf(h.beta1, step);
  const float bc2_sqrt = sqrtf(1.0f - powf(h.beta2, step));
  const float step_size = h.lr / bc1;
  const float decay = h.lr * h.weight_decay;
  const float omb1 = 1.0f - h.beta1, omb2 = 1.0f - h.beta2;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= inv_scale;
    pp -= decay * pp;
    mm = mm + (gg - mm) * omb1;
    vv = h.beta2 * vv + omb2 * gg * gg;
    const float denom = sqrtf(vv) / bc2_sqrt + h.eps;
    pp -= step_size * mm / denom;
  };
  const bool vec = (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                      reinterpret_cast<uintptr_t>(v)) & 15) == 0) &&
                   (sh == nullptr || (reinterpret_cast<uintptr_t>(sh) & 7) == 0);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const long long i = base + u * 1024 + threadIdx.x * 4;
    if (i + 4 <= n && vec) {
      float4 pp = *reinterpret_cast<const float4*>(p + i);
      const float4 gg = *reinterpret_cast<const float4*>(g + i);
      float4 mm = *reinterpret_cast<const float4*>(m + i);
      float4 vv = *reinterpret_cast<const float4*>(v + i);
      upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y);
      upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
      *reinterpret_cast<float4*>(p + i) = pp;
      *reinterpret_cast<float4*>(m + i) = mm;
      *reinterpret_cast<float4*>(v + i) = vv;
      if (sh) *reinterpret_cast<uint2*>(sh + i) = make_uint2(pack_bf16(pp.x, pp.y), pack_bf16(pp.z, pp.w));
    } else {
      for (long long j = i; j < n && j < i + 4; ++j) {
        float pp = p[j], mm = m[j], vv = v[j];
        upd(pp, g[j], mm, vv);
        p[j] = pp; m[j] = mm; v[j] = vv;
        if (sh) sh[j] = __float2bfloat16_rn(pp);
      }
    }
  }
}

// teacher <- m * teacher + (1 - m) * student (ssl/dino/model.py:126-139) with the teacher's bf16
// GEMM-operand shadow written in the same pass (p = teacher, g = student, sh = shadow or null)
__global__ void __launch_bounds__(256) ema_shadow_kernel(const __grid_constant__ AdamArgs a, const float mom) {
  int t = 0;
  while (t + 1 < a.count && a.block_start[t + 1] <= static_cast<int>(blockIdx.x)) ++t;
  const long long base = static_cast<long long>(blockIdx.x - a.block_start[t]) * AD_CHUNK;
  float* __restrict__ te = a.p[t];
  const float* __restrict__ st = a.g[t];
  __nv_bfloat16* __restrict__ sh = a.sh[t];
  const long long n = a.n[t];
  const float om = 1.0f - mom;
  const bool vec = (((reinterpret_cast<uintptr_t>(te) | reinterpret_cast<uintptr_t>(st)) & 15) == 0) &&
                   (sh == nullptr || (reinterpret_cast<uintptr_t>(sh) & 7) == 0);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const long long i = base + u * 1024 + threadIdx.x * 4;
    if (i + 4 <= n && vec) {
      float4 x = *reinterpret_cast<const float4*>(te + i);
      const float4 s = *reinterpret_cast<const float4*>(st + i);
      // same operation order as param.mul_(m).add_((1 - m) * student)
      x.x = x.x * mom + om * s.x; x.y = x.y * mom + om * s.y;
      x.z = x.z * mom + om * s.z; x.w = x.w * mom + om * s.w;
      *reinterpret_cast<float4*>(te + i) = x;
      if (sh) *reinterpret_cast<uint2*>(sh + i) = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
    } else {
      for (long long j = i; j < n && j < i + 4; ++j) {
        const float x = te[j] * mom + om * st[j];
        te[j] = x;
        if (sh) sh[j] = __float2bfloat16_rn(x);
      }
    }
  }
}

}  // namespace
}  // namespace vitssl

using namespace vitssl;

extern "C" int vitssl_multi_ema_shadow(void* const* host_teacher, const void* const* host_student,
                                       void* const* host_shadow, const int64_t* host_numel, int count,
                                       float momentum, cudaStream_t stream) {
  VITSSL_REQUIRE(count >= 0 && (count == 0 || (host_teacher && host_student && host_numel)), VITSSL_ERR_ARG,
                 "multi_ema_shadow: bad args");
  int done = 0;
  while (done < count) {
    AdamArgs a{};
    int blocks = 0, k = 0;
    while (done + k < count && k < AD_MAX) {
      const int i = done + k;
      VITSSL_REQUIRE(host_teacher[i] && host_student[i] && host_numel[i] >= 0, VITSSL_ERR_ARG,
                     "multi_ema_shadow: null tensor %d", i);
      a.p[k] = static_cast<float*>(host_teacher[i]);
      a.g[k] = static_cast<float*>(const_cast<void*>(host_student[i]));
      a.sh[k] = host_shadow ? static_cast<__nv_bfloat16*>(host_shadow[i]) : nullptr;
      a.n[k] = host_numel[i];
      a.block_start[k] = blocks;
      blocks += static_cast<int>((host_numel[i] + AD_CHUNK - 1) / AD_CHUNK);
      ++k;
    }
    a.block_start[k] = blocks;
    a.count = k;
    if (blocks > 0) {
      ema_shadow_kernel<<<blocks, 256, 0, stream>>>(a, momentum);
      const int rc = check_launch("multi_ema_shadow");
      if (rc) return rc;
    }
    done += k;
  }
  return 0;
}

extern "C" int vitssl_adamw_step(void* const* host_param, const void* const* host_grad, void* const* host_exp_avg,
                                 void* const* host_exp_avg_sq, void* const* host_shadow, void* const* host_step,
                                 const int64_t* host_numel, int count, float lr, float beta1, float beta2,
                                 float eps, float weight_decay, const float* grad_scale, const float* found_inf,
                                 cudaStream_t stream) {
  VITSSL_REQUIRE(count >= 0 && (count == 0 || (host_param && host_grad && host_exp_avg && host_exp_avg_sq &&
                                               host_step && host_numel)),
                 VITSSL_ERR_ARG, "adamw_step: bad args");
  VITSSL_REQUIRE(lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f,
                 VITSSL_ERR_ARG, "adamw_step: invalid hyper-parameters");
  AdamHyper h{lr, beta1, beta2, eps, weight_decay, grad_scale, found_inf};
  int done = 0;
  while (done < count) {
    AdamArgs a{};
    int blocks = 0, k = 0;
    while (done + k < count && k < AD_MAX) {
      const int i = done + k;
      VITSSL_REQUIRE(host_param[i] && host_grad[i] && host_exp_avg[i] && host_exp_avg_sq[i] && host_step[i] &&
                         host_numel[i] >= 0,
                     VITSSL_ERR_ARG, "adamw_step: null tensor %d", i);
      a.p[k] = static_cast<float*>(host_param[i]);
      a.g[k] = static_cast<float*>(const_cast<void*>(host_grad[i]));
      a.m[k] = static_cast<float*>(host_exp_avg[i]);
      a.v[k] = static_cast<float*>(host_exp_avg_sq[i]);
      a.sh[k] = host_shadow ? static_cast<__nv_bfloat16*>(host_shadow[i]) : nullptr;
      a.step[k] = static_cast<float*>(host_step[i]);
      a.n[k] = host_numel[i];
      a.block_start[k] = blocks;
      blocks += static_cast<int>((host_numel[i] + AD_CHUNK - 1) / AD_CHUNK);
      ++k;
    }
    a.block_start[k] = blocks;
    a.count = k;
    adamw_advance_steps_kernel<<<1, 64, 0, stream>>>(a, found_inf);
    {
      const int rc = check_launch("adamw_advance_steps");
      if (rc) return rc;
    }
    if (blocks > 0) {
      adamw_kernel<<<blocks, 256, 0, stream>>>(a, h);
      const int rc = check_launch("adamw_step");
      if (rc) return rc;
    }
    done += k;
  }
  return 0;
}
