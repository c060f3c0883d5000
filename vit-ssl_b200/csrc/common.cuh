// vitssl_b200 — shared device/host helpers for the sm_100a kernels.
// Raw PTX wrappers for mbarrier / TMA / tcgen05 / TMEM, the Philox4x32-10
// generator used for dropout masks, and the thread-local error slot of the C-ABI.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace vitssl {

// ----------------------------------------------------------------------------------
// Error plumbing (include/vitssl_b200.h: vitssl_last_error)
// ----------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // returns 0 or negative code, records message

#define VITSSL_REQUIRE(cond, code, ...)      \
  do {                                       \
    if (!(cond)) {                           \
      ::vitssl::set_error(__VA_ARGS__);      \
      return (code);                         \
    }                                        \
  } while (0)

enum : int {
  VITSSL_OK = 0,
  VITSSL_ERR_ARG = -1,
  VITSSL_ERR_SHAPE = -2,
  VITSSL_ERR_CUDA = -3,
  VITSSL_ERR_DEVICE = -4,
  VITSSL_ERR_WORKSPACE = -5,
};

int num_sms();  // cached SM count of the current device

// Per-launch device timing of the C-sequenced paths (vitssl_profile_begin / _end, api.cu). While a
// profile is open, a ProfScope records a CUDA event on `stream` before and after the launches it
// brackets, tagged with a kernel family and its algorithmic work (FLOPs or bytes). Closed: no-op.
bool prof_on();
struct ProfScope {
  ProfScope(const char* kind, double work, cudaStream_t stream) : idx_(-1), stream_(stream) {
    if (prof_on()) begin(kind, work);
  }
  ~ProfScope() { if (idx_ >= 0) end(); }
  ProfScope(const ProfScope&) = delete;
  ProfScope& operator=(const ProfScope&) = delete;
 private:
  void begin(const char* kind, double work);
  void end();
  int idx_;
  cudaStream_t stream_;
};
// non-null while encoder.cu captures a stack call into a CUDA graph: dropout kernels launched by this
// thread then read their seed from that device word instead of their by-value argument
const unsigned long long* seed_slot_override();
// kernels a replayed graph launches are counted like individually launched ones (vitssl_launch_count)
long long launches_now();
void add_launches(long long n);
bool pdl_enabled();  // programmatic dependent launch for the hot kernels (VITSSL_PDL=0 disables)

// Launch with programmatic stream serialization: the grid may begin (prologue: barrier init,
// TMEM allocation, descriptor prefetch, launch latency) while the previous kernel on the stream
// is still draining; the kernel itself calls pdl_wait() before its first global-memory access,
// which blocks until that kernel has completed and its writes are visible. Only kernels that
// contain pdl_wait() may be launched this way.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ----------------------------------------------------------------------------------
// Small device helpers
// ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
// programmatic dependent launch (see launch_pdl): let the next kernel on the stream start its
// prologue now / wait for the previous kernel's completion and memory flush
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
// IEEE half pairs: the saved GELU backward factor travels as fp16 (11 significant bits in the same
// 2 bytes: its range is [-0.2, 1.2] / (1 - p), so fp16 loses nothing a bf16 copy would keep)
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_f16(uint32_t v) {
  return __half22float2(*reinterpret_cast<__half2*>(&v));
}
__device__ __forceinline__ float bf16_round(float x) {
  return __bfloat162float(__float2bfloat16_rn(x));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact (erf) GELU and its derivative — reference: feed_forward.py:26 (F.gelu default)
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// ----------------------------------------------------------------------------------
// Philox4x32 (counter-based; one 128-bit draw per (seed, offset, index)). 10 rounds is the
// curand-compatible generator (masking.cu replays torch's streams with it); the dropout masks
// use the 7-round variant (the shortest Crush-resistant one), they only have to be reproducible
// between forward and backward.
// ----------------------------------------------------------------------------------
template <int ROUNDS = 10>
__device__ __forceinline__ uint4 philox4x32(uint64_t seed, uint64_t offset, uint64_t index) {
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  uint32_t c0 = static_cast<uint32_t>(index), c1 = static_cast<uint32_t>(index >> 32);
  uint32_t c2 = static_cast<uint32_t>(offset), c3 = static_cast<uint32_t>(offset >> 32);
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    uint32_t hi0, lo0, hi1, lo1;  // one IMAD.WIDE.U32 per product
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}"
        : "=r"(lo0), "=r"(hi0) : "r"(c0), "r"(0xD2511F53u));
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}"
        : "=r"(lo1), "=r"(hi1) : "r"(c2), "r"(0xCD9E8D57u));
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
constexpr int DROPOUT_PHILOX_ROUNDS = 7;
// Same generator with the round keys precomputed on the host: as kernel parameters they become
// constant-bank operands of the round's 3-input XOR, so a round is 2 IMAD.WIDE + 2 LOP3 and no
// key arithmetic (the in-kernel schedule costs two more ALU instructions per round and draw).
// The KEY is the dropout site (`offset`, the same every step, so the round keys are launch constants
// that a replayed CUDA graph may keep) and the per-step `seed` rides in the two high counter words,
// which a kernel can take from its arguments or from a device word: a counter-based generator is a
// keyed bijection of the whole 128-bit counter, so (site, seed, index) -> draw is as good either way.
struct PhiloxKeys7 {
  uint32_t k[2 * DROPOUT_PHILOX_ROUNDS];  // (k0, k1) of round r at [2r], [2r+1]
  uint32_t c2, c3;                        // seed words
};
inline PhiloxKeys7 make_philox_keys7(uint64_t seed, uint64_t offset) {
  PhiloxKeys7 pk;
  const uint64_t key = offset * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;  // spread the small site index
  uint32_t k0 = static_cast<uint32_t>(key), k1 = static_cast<uint32_t>(key >> 32);
  for (int r = 0; r < DROPOUT_PHILOX_ROUNDS; ++r) {
    pk.k[2 * r] = k0; pk.k[2 * r + 1] = k1;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  pk.c2 = static_cast<uint32_t>(seed); pk.c3 = static_cast<uint32_t>(seed >> 32);
  return pk;
}
__device__ __forceinline__ uint4 philox4x32_7_keyed(const PhiloxKeys7& pk, uint32_t c2, uint32_t c3, uint64_t index) {
  uint32_t c0 = static_cast<uint32_t>(index), c1 = static_cast<uint32_t>(index >> 32);
#pragma unroll
  for (int r = 0; r < DROPOUT_PHILOX_ROUNDS; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}"
        : "=r"(lo0), "=r"(hi0) : "r"(c0), "r"(0xD2511F53u));
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}"
        : "=r"(lo1), "=r"(hi1) : "r"(c2), "r"(0xCD9E8D57u));
    const uint32_t n0 = hi1 ^ c1 ^ pk.k[2 * r], n2 = hi0 ^ c3 ^ pk.k[2 * r + 1];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  return make_uint4(c0, c1, c2, c3);
}
// Dropout of element e: the 16-bit lanes of one Philox draw cover 8 consecutive elements
// (draw index = e / 8, lane = e % 8 counted from the low half of .x); keep iff u16 >= thresh16
// (thresh16 = p * 65536). Both forms below implement exactly this rule.
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint64_t offset, uint64_t group8,
                                                  uint32_t thresh16) {
  const uint4 r = philox4x32<DROPOUT_PHILOX_ROUNDS>(seed, offset, group8);
  uint32_t m = 0;
  m |= ((r.x & 0xffffu) >= thresh16) << 0;
  m |= ((r.x >> 16) >= thresh16) << 1;
  m |= ((r.y & 0xffffu) >= thresh16) << 2;
  m |= ((r.y >> 16) >= thresh16) << 3;
  m |= ((r.z & 0xffffu) >= thresh16) << 4;
  m |= ((r.z >> 16) >= thresh16) << 5;
  m |= ((r.w & 0xffffu) >= thresh16) << 6;
  m |= ((r.w >> 16) >= thresh16) << 7;
  return m;
}
// same rule as per-element multipliers: m[i] = keep ? scale : 0 (the 16-bit compares are done
// in place on the 32-bit words: hi >= t <=> word >= t << 16; lo >= t <=> word << 16 >= t << 16)
__device__ __forceinline__ void dropout_scale8(uint64_t seed, uint64_t offset, uint64_t group8,
                                               uint32_t thresh16, float scale, float (&m)[8]) {
  const uint4 r = philox4x32<DROPOUT_PHILOX_ROUNDS>(seed, offset, group8);
  const uint32_t th = thresh16 << 16;
  m[0] = ((r.x << 16) >= th) ? scale : 0.0f;
  m[1] = (r.x >= th) ? scale : 0.0f;
  m[2] = ((r.y << 16) >= th) ? scale : 0.0f;
  m[3] = (r.y >= th) ? scale : 0.0f;
  m[4] = ((r.z << 16) >= th) ? scale : 0.0f;
  m[5] = (r.z >= th) ? scale : 0.0f;
  m[6] = ((r.w << 16) >= th) ? scale : 0.0f;
  m[7] = (r.w >= th) ? scale : 0.0f;
}

// Lane-mask form used by the GEMM epilogues (their dropout site is private to the GELU GEMM pair,
// so its rule may differ from the LayerNorm sites): element e keeps iff (u16 & 0x7fff) >= th15,
// th15 = p * 32768. For a Philox word holding two elements the test runs on both 16-bit lanes at
// once: with the lanes' top bits forced to 1 the subtraction cannot borrow across lanes and
// leaves each lane's top bit set iff that lane passes; PRMT's sign-replicate mode then widens the
// bit to 0xffff, which is ANDed onto the packed bf16x2 result. th2 = th15 * 0x10001.
__device__ __forceinline__ uint32_t dropout_lane_mask2(uint32_t r, uint32_t th2) {
  const uint32_t t = (r | 0x80008000u) - th2;
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, 0xBB99;" : "=r"(m) : "r"(t));  // bytes {1s,1s,3s,3s}: s = sign replicate
  return m;
}

// ----------------------------------------------------------------------------------
// packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2): two lanes per issue slot
// ----------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// standard normal cdf of two values: Phi(x) = 0.5 + t * Q(t^2), t = clamp(x, +-4.2), Q the
// degree-8 minimax polynomial (|error| <= 1.1e-5 over the whole real line, fp32 evaluation).
// No MUFU, 10 packed FMAs per pair. GELU(x) = x * Phi(x) (feed_forward.py:26, exact-erf GELU).
__device__ __forceinline__ f32x2 normal_cdf2(float x0, float x1) {
  const float t0 = fminf(fmaxf(x0, -4.2f), 4.2f), t1 = fminf(fmaxf(x1, -4.2f), 4.2f);
  const f32x2 t = pk2(t0, t1);
  const f32x2 s = fmul2(t, t);
  f32x2 q = ffma2(s, pk2(5.994787999e-11f, 5.994787999e-11f), pk2(-5.630807376e-09f, -5.630807376e-09f));
  q = ffma2(s, q, pk2(2.342876257e-07f, 2.342876257e-07f));
  q = ffma2(s, q, pk2(-5.759411124e-06f, -5.759411124e-06f));
  q = ffma2(s, q, pk2(9.456112457e-05f, 9.456112457e-05f));
  q = ffma2(s, q, pk2(-1.114074141e-03f, -1.114074141e-03f));
  q = ffma2(s, q, pk2(9.829915129e-03f, 9.829915129e-03f));
  q = ffma2(s, q, pk2(-6.636010855e-02f, -6.636010855e-02f));
  q = ffma2(s, q, pk2(3.989073634e-01f, 3.989073634e-01f));
  return ffma2(t, q, pk2(0.5f, 0.5f));
}

// Same function with one clamp per value instead of two: the polynomial argument is
// s = min(x^2, 4.2^2) and the final multiply-add uses the unclamped x with .sat, which pins the
// result to [0, 1]; past |x| = 4.2 the line 0.5 + x Q(4.2^2) leaves [0, 1] within 1.3e-5 of the
// true tail. (min/max run on the half-rate ALU pipe, which bounds the GELU epilogues.)
__device__ __forceinline__ f32x2 normal_cdf2_sat(float x0, float x1) {
  const f32x2 x = pk2(x0, x1);
  float s0, s1;
  upk2(fmul2(x, x), s0, s1);
  const f32x2 s = pk2(fminf(s0, 17.64f), fminf(s1, 17.64f));
  f32x2 q = ffma2(s, pk2(5.994787999e-11f, 5.994787999e-11f), pk2(-5.630807376e-09f, -5.630807376e-09f));
  q = ffma2(s, q, pk2(2.342876257e-07f, 2.342876257e-07f));
  q = ffma2(s, q, pk2(-5.759411124e-06f, -5.759411124e-06f));
  q = ffma2(s, q, pk2(9.456112457e-05f, 9.456112457e-05f));
  q = ffma2(s, q, pk2(-1.114074141e-03f, -1.114074141e-03f));
  q = ffma2(s, q, pk2(9.829915129e-03f, 9.829915129e-03f));
  q = ffma2(s, q, pk2(-6.636010855e-02f, -6.636010855e-02f));
  q = ffma2(s, q, pk2(3.989073634e-01f, 3.989073634e-01f));
  float q0, q1, r0, r1;
  upk2(q, q0, q1);
  asm("fma.rn.sat.f32 %0, %1, %2, 0f3F000000;" : "=f"(r0) : "f"(x0), "f"(q0));
  asm("fma.rn.sat.f32 %0, %1, %2, 0f3F000000;" : "=f"(r1) : "f"(x1), "f"(q1));
  return pk2(r0, r1);
}

// One lane of a converged warp (always the same one). Unlike `if (lane == 0)`, the compiler knows
// the branch is taken by a single elected lane of a warp in uniform control flow, so operands
// computed from warp-uniform values stay in uniform registers (tcgen05.mma / TMA operands).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a legitimate wait is microseconds; after ~2 s of spinning the kernel traps so a
// protocol bug surfaces as a launch failure instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("vitssl: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// Same, for the single-thread producer / MMA-issuer roles: the try_wait carries a suspend-time
// hint so a waiting warp parks in hardware instead of burning issue slots of the epilogue warps
// that share its scheduler (it still wakes as soon as the phase completes).
__device__ __forceinline__ bool mbar_try_wait_parked(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(200000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_parked(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_parked(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("vitssl: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_parked_addr(uint32_t cta_addr, uint32_t parity) {
  auto try_wait = [&]() {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(cta_addr), "r"(parity), "r"(200000u)
        : "memory");
    return ok != 0;
  };
  if (try_wait()) return;
  const long long t0 = clock64();
  while (!try_wait()) {
    if (clock64() - t0 > 4000000000LL) {
      printf("vitssl: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t cta_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(cta_addr) : "memory");
}

// ----------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — loads complete on an mbarrier
// ----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// TMA store: shared memory tile -> global (clipped at the tensor bounds), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
      : "memory");
}
// TMA reduction: global[tile] += shared tile (fp32 add performed at L2 on whole sectors), same
// bulk-group completion as a store. Split-K partial sums use it instead of per-element red.add.
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(m)),
      "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {  // smem of all but the N newest groups is reusable
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ----------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------
template <int kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same MMA with the two shared-memory descriptors given as (low word, shared high word): an issuer
// that keeps the per-tile low words (start address >> 4 | LBO field) in uniform registers only
// adds a constant per k-step. PAIR selects cta_group::2.
// (a_hi, b_hi) variant for operands with different layouts (swizzled A, unswizzled constant B)
__device__ __forceinline__ void umma_bf16_ss_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                  uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void umma_bf16_ss_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi,
                                                uint32_t idesc, uint32_t accumulate) {
  if constexpr (PAIR) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(hi)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(hi)
        : "memory");
  }
}
// arrive on `bar` once every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base_lane + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// D[tmem] (+)= A[tmem, K-major rows = lanes, bf16 packed 2 per column] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_lo(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM: thread i of the warp writes 16 consecutive 32-bit columns of lane (base + i)
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t tmem_ld_32x1(uint32_t taddr) {  // one column: thread i gets lane (base + i)
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// ----------------------------------------------------------------------------------
// CTA pairs (cluster of 2 on one TPC) for tcgen05 cta_group::2
// ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local smem address` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; completion bytes are credited to the mbarrier at
// `bar_cluster_addr` (the leader CTA's barrier), data lands in the issuing CTA's own smem
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result) {  // same warp id in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols)
               : "memory");
}
// 256 x N tile across the pair: A rows and B columns are split between the two CTAs' smem,
// each CTA's TMEM receives its 128 rows. Issued by the leader CTA only.
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                  uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this smem offset in BOTH CTAs once the pair's MMAs so far have retired
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
          "r"(smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// Shared-memory matrix descriptor for tcgen05.mma operands, 128-byte swizzle.
// Bit layout (PTX ISA "tcgen05 shared memory descriptor"): [0,14) start>>4, [16,30) LBO>>4,
// [32,46) SBO>>4, [46,48) version=1, [61,64) layout type (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                    uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// Shared-memory matrix descriptor without swizzle (8-row x 16-byte core matrices; LBO / SBO are the
// byte distances between core matrices along the two tile dimensions). Only used for the constant
// all-ones operand of the fused row-sum MMA, where every core matrix holds the same value, so the
// two strides merely have to stay inside the ones block.
__device__ __forceinline__ uint64_t umma_desc_noswizzle(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;
  return d;
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
// [4,6) D fmt (1=f32), [7,10) A fmt (1=bf16), [10,13) B fmt, [15] A major (1=MN), [16] B major,
// [17,23) N>>3, [24,29) M>>4.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) |
         (static_cast<uint32_t>(b_mn) << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ----------------------------------------------------------------------------------
// Host: TMA descriptor encoding through the driver entry point (no libcuda link dependency)
// ----------------------------------------------------------------------------------
// 2-D bf16 tensor [outer][inner] with row pitch ld_bytes; 128B-swizzled box.
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t ld_bytes, uint32_t box_inner, uint32_t box_outer);
// general 2-D map: elem_bytes 2 (bf16) or 4 (fp32); swizzle_bytes 0/32/64/128
int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t outer,
                 uint64_t ld_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes);
// 3-D bf16 tensor [d2][d1][inner]
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t d1, uint64_t d2,
                      uint64_t ld1_bytes, uint64_t ld2_bytes, uint32_t box_inner, uint32_t box_d1,
                      uint32_t box_d2);
int make_tmap_bf16_3d_sw(CUtensorMap* out, const void* base, uint64_t inner, uint64_t d1, uint64_t d2,
                         uint64_t ld1_bytes, uint64_t ld2_bytes, uint32_t box_inner, uint32_t box_d1,
                         uint32_t box_d2, int swizzle_bytes);

}  // namespace vitssl
