// vitssl_b200 — fused multi-head attention for sm_100a (reference: attention.py:5-27,86-103).
//   softmax(Q K^T / sqrt(d_k)) V, no mask, no attention dropout, d_k = 64, S <= 256 keys.
// Forward: one CTA per (batch, head, 128-query tile). TMA stages Q, K, V of the head in shared
// memory (zero-filled past S); S = Q K^T is accumulated in TMEM by tcgen05.mma; 128 threads (one
// per query row = TMEM lane) do the softmax in registers and write P (bf16) back to shared memory
// in the 128B-swizzled K-major layout; O = P V is a second tcgen05.mma whose accumulator reuses
// the S columns; the epilogue normalises by the row sum and stores bf16 context + fp32 LSE.
// With S <= 256 the whole key range is one block, so the online-softmax rescale degenerates to
// a single max/sum pass. Two CTAs are resident per SM (96 KB smem, 256 TMEM columns each) so the
// MMA of one overlaps the softmax of the other.
// Backward: one CTA per (batch, head); recomputes P from the saved LSE, keeps dQ/dK/dV
// accumulators in TMEM (512 columns) and uses both operand majors so that no transposes are
// materialised: dV = P^T dO, dK = dS^T Q (MN-major A and B), dQ = dS K (K-major A, MN-major B).
// Shapes outside (d_k = 64, S <= 256) and the return_attn=True path use the generic SIMT kernels
// at the bottom of this file.
#include "common.cuh"
#include "vitssl_b200.h"

namespace vitssl {
namespace {

constexpr float kLog2e = 1.4426950408889634f;

// per-phase clock64 timelines of CTA 0 (scripts/trace_attn_fwd.py, trace_attn_bwd.py); compiled out by default
#ifdef VITSSL_ATTN_TRACE
__device__ long long g_attn_trace[8192];
#define TRACE(slot) do { if (blockIdx.x == 0 && (slot) < 8192) g_attn_trace[(slot)] = clock64(); } while (0)
#define FTRACE(slot) do { if (blockIdx.x == 0 && (slot) < 4096) g_attn_trace[4096 + (slot)] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#define FTRACE(slot) do { } while (0)
#endif

// 16-byte shared-memory store with the state space spelled out
__device__ __forceinline__ void sts128_(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// One thread's 32 consecutive fp32 values (raw bits) of row `lane` -> a 32 x 32 bf16 staging block
// (64-byte rows in the TMA SWIZZLE_64B pattern) that a TMA store then writes as full lines; per-thread
// 16-byte global stores would touch 32 different lines per instruction. LO = false: bf16(x * mul);
// LO = true: the rounding residual bf16(x * mul - bf16(x * mul)).
template <bool LO>
__device__ __forceinline__ void stage_block32(uint32_t dst, int lane, const uint32_t (&v)[32], float mul) {
  const uint32_t base = dst + lane * 64;
  const int sw = (lane >> 1) & 3;
  const f32x2 m2 = pk2(mul, mul);
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float x0, x1;
      upk2(fmul2(pk2(__uint_as_float(v[8 * q4 + 2 * e]), __uint_as_float(v[8 * q4 + 2 * e + 1])), m2), x0, x1);
      uint32_t hi = pack_bf16(x0, x1);
      w[e] = LO ? pack_bf16(x0 - bf16_lo(hi), x1 - bf16_hi(hi)) : hi;
    }
    sts128_(base + ((q4 ^ sw) << 4), w[0], w[1], w[2], w[3]);
  }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
struct AttnFwdParams {
  __nv_bfloat16* out; long long ldo;  // context [B*Sq, ldo]; head h occupies cols [64h, 64h+64)
  __nv_bfloat16* out_lo;              // nullable, same layout: bf16(O - bf16(O)), see attn_delta_kernel
  float* lse;                         // [B, H, Sq] log-sum-exp of the scaled scores
  int B, H, Sq, Sk, kv_rows;          // kv_rows = round_up(Sk, 16) <= 256
  float scale;
  int staged;                         // 1: context leaves through shared-memory staging + TMA stores
  // Packed short sequences (self-attention, S <= 64: DINO local crops, cfg 1): G = 128 / S images share
  // one 128-row tile. The token rows of consecutive images are contiguous ([B*S] rows), so the tile is
  // simply rows [g*G*S, g*G*S + 128) of that matrix and the attention becomes block-diagonal with
  // block size S; rows past G*S belong to the next group and are masked / never stored.
  int G, GS;                          // images per tile (1 = one image per tile), G * S
};

constexpr int FWD_THREADS = 192;  // warps 0-3 softmax (TMEM lane quadrant = warp), 4 = TMA, 5 = MMA
constexpr int FWD_SMEM_Q = 0;               // 16 KB: 128 query rows x 128 B
constexpr int FWD_SMEM_K = 16384;           // 32 KB: up to 256 key rows
constexpr int FWD_SMEM_V = 16384 + 32768;   // 32 KB
constexpr int FWD_SMEM_STG = 16384 + 65536;  // 16 KB: two 32 x 32 bf16 staging blocks (2 KB each) per softmax warp
constexpr int FWD_SMEM_BAR = FWD_SMEM_STG + 16384;
constexpr int FWD_SMEM_BYTES = FWD_SMEM_BAR + 128 + 1024;
constexpr uint32_t FWD_COL_O = 128;  // O accumulator columns [128,192): inside S, past the packed P

// Persistent: each CTA (two per SM) loops over (batch, head, 128-query tile) items. Per item
//   TMA warp : Q,K as soon as the previous item's S MMA retired; V once its PV MMA retired
//   MMA warp : S = Q K^T (SS) -> [softmax] -> O = P V with P read from TMEM (TS form)
//   softmax  : row max, p = 2^((s - max) * scale * log2e) with packed FMAs + MUFU.EX2, P written
//              back over S in TMEM as packed bf16 (tcgen05.st), then O / rowsum -> global.
// Nothing of the S x S probability matrix ever reaches shared or global memory
// (attention.py:20-23 materialises it three times).
// STAGED: context leaves through staging + TMA stores (S > 64) or per-thread stores.
// PACKED: several short sequences per tile (see AttnFwdParams::G), always with per-thread stores.
template <bool STAGED, bool PACKED>
__global__ void __launch_bounds__(FWD_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o,
                const __grid_constant__ CUtensorMap tmap_olo, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem + FWD_SMEM_Q;
  uint8_t* sK = smem + FWD_SMEM_K;
  uint8_t* sV = smem + FWD_SMEM_V;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + FWD_SMEM_BAR);
  uint64_t* bar_qk = bar + 0;      // Q,K landed
  uint64_t* bar_v = bar + 1;       // V landed
  uint64_t* bar_s = bar + 2;       // S MMA retired: S readable, Q/K smem reusable
  uint64_t* bar_p = bar + 3;       // P written to TMEM by all 128 softmax threads
  uint64_t* bar_o = bar + 4;       // PV MMA retired: O readable, V smem reusable
  uint64_t* bar_oread = bar + 5;   // O drained: TMEM reusable for the next item's S
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 6);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool packed = PACKED;
  const int nq = (p.Sq + 127) / 128;                                        // 1 when packed
  const int items = (packed ? (p.B + p.G - 1) / p.G : p.B) * p.H * nq;      // packed: (group, head)

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
      tma_prefetch_desc(&tmap_o); tma_prefetch_desc(&tmap_olo);
      mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1);
      mbar_init(bar_p, 4); mbar_init(bar_o, 1); mbar_init(bar_oread, 4);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<256>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();  // the prologue overlapped the previous kernel's tail; global memory from here on
  const uint32_t kv_bytes = static_cast<uint32_t>(p.kv_rows) * 128u;

  if (warp == 4) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int qt = item % nq, h = (item / nq) % p.H, b = item / (nq * p.H);
        // packed: b is the group index, the maps are [B*S] rows x 1 batch
        const int qrow0 = packed ? b * p.GS : qt * 128, krow0 = packed ? b * p.GS : 0, bc = packed ? 0 : b;
        if (it > 0) mbar_wait(bar_s, (it - 1) & 1);
        mbar_expect_tx(bar_qk, 16384u + kv_bytes);
        tma_load_3d(sQ, &tmap_q, bar_qk, h * 64, qrow0, bc);
        tma_load_3d(sK, &tmap_k, bar_qk, h * 64, krow0, bc);
        if (it > 0) mbar_wait(bar_o, (it - 1) & 1);
        mbar_expect_tx(bar_v, kv_bytes);
        tma_load_3d(sV, &tmap_v, bar_v, h * 64, krow0, bc);
      }
    }
  } else if (warp == 5) {
    // ------------------------------ MMA issuer ------------------------------
    // the whole warp walks the items and the barrier waits (uniform control flow); one elected lane
    // issues, with descriptor low words kept in uniform registers (see gemm_sm100.cu)
    {
      const uint32_t idesc_s = umma_idesc_bf16(128, p.kv_rows, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, false, true);
      constexpr uint32_t HI = 0x40004040u;  // SBO 1024 | descriptor version 1 | SWIZZLE_128B
      const uint32_t q_lo = (smem_u32(sQ) >> 4) | ((16u >> 4) << 16), k_lo = (smem_u32(sK) >> 4) | ((16u >> 4) << 16);
      const uint32_t v_lo = (smem_u32(sV) >> 4) | ((8192u >> 4) << 16);  // MN-major B operand
      const int ksteps = p.kv_rows / 16;
      int it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        mbar_wait(bar_qk, it & 1);
        if (lane == 0) FTRACE(16 * it + 0);
        if (it > 0) mbar_wait(bar_oread, (it - 1) & 1);
        tc_fence_after();
        if (lane == 0) FTRACE(16 * it + 1);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss_lo<false>(tmem, q_lo + 2 * k, k_lo + 2 * k, HI, idesc_s, k > 0);
          umma_commit(bar_s);
        }
        __syncwarp();
        mbar_wait(bar_p, it & 1);
        if (lane == 0) FTRACE(16 * it + 2);
        mbar_wait(bar_v, it & 1);
        tc_fence_after();
        if (lane == 0) FTRACE(16 * it + 3);
        if (elect_one()) {
          for (int ks = 0; ks < ksteps; ++ks)  // P[128, 16 keys] = 8 packed columns per step
            umma_bf16_ts_lo(tmem + FWD_COL_O, tmem + ks * 8, v_lo + ks * 128, HI, idesc_o, ks > 0);
          umma_commit(bar_o);
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------ softmax + epilogue ------------------------------
    // thread t owns query row qt*128 + t == TMEM lane t
    const int t = threadIdx.x;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const float sl2 = p.scale * kLog2e;
    const f32x2 sl2v = pk2(sl2, sl2);
    const int nchunks = (p.kv_rows + 31) / 32;
    // this thread's valid key columns [clo, chi): all Sk keys, or (packed) the S keys of its own image
    int clo = 0, chi = p.Sk;
    if (packed) {
      const int gi = t / p.Sq;
      clo = t < p.GS ? gi * p.Sq : 0;
      chi = t < p.GS ? clo + p.Sq : 0;
    }
    const unsigned clen = static_cast<unsigned>(chi - clo);
    auto col_ok = [&](int col) { return packed ? static_cast<unsigned>(col - clo) < clen : col < chi; };
    auto chunk_full = [&](int c) { return packed ? (c * 32 >= clo && c * 32 + 32 <= chi) : c * 32 + 32 <= chi; };
    auto chunk_none = [&](int c) { return packed && (c * 32 + 32 <= clo || c * 32 >= chi); };
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int qt = item % nq, h = (item / nq) % p.H, b = item / (nq * p.H);
      if (t == 0) FTRACE(16 * it + 7);
      mbar_wait(bar_s, it & 1);
      tc_fence_after();
      if (t == 0) FTRACE(16 * it + 8);
      // pass 1: row maximum (TMEM loads run one chunk ahead; only the chunk straddling Sk is masked)
      float mx = -INFINITY;
      {
        uint32_t ra[32], rb[32];
        __syncwarp();
        tmem_ld_32x32(lane_addr, ra);
        auto pass1 = [&](uint32_t (&cur)[32], uint32_t (&nxt)[32], int c) {
          tmem_ld_wait();
          if (c + 1 < nchunks) tmem_ld_32x32(lane_addr + (c + 1) * 32, nxt);
          if (chunk_none(c)) {
            // (packed) none of this chunk's keys belong to this row's image
          } else if (chunk_full(c)) {
            float m0 = __uint_as_float(cur[0]), m1 = __uint_as_float(cur[1]);
#pragma unroll
            for (int i = 2; i < 32; i += 2) {
              m0 = fmaxf(m0, __uint_as_float(cur[i]));
              m1 = fmaxf(m1, __uint_as_float(cur[i + 1]));
            }
            mx = fmaxf(mx, fmaxf(m0, m1));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col_ok(c * 32 + i)) mx = fmaxf(mx, __uint_as_float(cur[i]));
          }
        };
        for (int c = 0; c < nchunks; c += 2) {
          pass1(ra, rb, c);
          if (c + 1 < nchunks) pass1(rb, ra, c + 1);
        }
      }
      if (t == 0) FTRACE(16 * it + 9);
      // pass 2: P over S in place (chunk c of 32 fp32 columns -> 16 packed bf16x2 columns at 16c)
      const float mneg = -mx * sl2;
      const f32x2 mnegv = pk2(mneg, mneg);
      f32x2 sum2 = pk2(0.f, 0.f);
      {
        uint32_t ra[32], rb[32];
        tmem_ld_32x32(lane_addr, ra);
        auto pass2 = [&](uint32_t (&cur)[32], uint32_t (&nxt)[32], int c) {
          tmem_ld_wait();
          if (c + 1 < nchunks) tmem_ld_32x32(lane_addr + (c + 1) * 32, nxt);
          uint32_t pk[16];
          const bool full = chunk_full(c);
          if (chunk_none(c)) {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = 0u;
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float a0, a1;
              upk2(ffma2(pk2(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])), sl2v, mnegv), a0, a1);
              float e0 = ex2_approx(a0), e1 = ex2_approx(a1);
              if (!full) {
                e0 = col_ok(c * 32 + i) ? e0 : 0.f;
                e1 = col_ok(c * 32 + i + 1) ? e1 : 0.f;
              }
              sum2 = fadd2(sum2, pk2(e0, e1));  // fp32 row sum: exact LSE for the backward recomputation
              pk[i >> 1] = pack_bf16(e0, e1);
            }
          }
          tmem_st_32x16(lane_addr + c * 16, pk);
        };
        for (int c = 0; c < nchunks; c += 2) {
          pass2(ra, rb, c);
          if (c + 1 < nchunks) pass2(rb, ra, c + 1);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);  // one arrival per warp
      if (t == 0) FTRACE(16 * it + 10);
      float sum;
      {
        float s0, s1;
        upk2(sum2, s0, s1);
        sum = s0 + s1;
      }

      mbar_wait(bar_o, it & 1);
      tc_fence_after();
      if (t == 0) FTRACE(16 * it + 11);
      // packed: row t of the tile is global token row b * GS + t (b = group index)
      const int qrow = packed ? t : qt * 128 + t;
      const long long grow = packed ? static_cast<long long>(b) * p.GS + t : 0;
      const bool row_ok = packed ? (t < p.GS && grow < static_cast<long long>(p.B) * p.Sq) : qrow < p.Sq;
      const float inv = 1.0f / sum;
      {
        uint32_t r0[32], r1[32];
        __syncwarp();
        tmem_ld_32x32(lane_addr + FWD_COL_O, r0);
        tmem_ld_32x32(lane_addr + FWD_COL_O + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_oread);  // O is in registers: the next item's S MMA may overwrite TMEM
        if (t == 0) FTRACE(16 * it + 12);
        if constexpr (STAGED) {
          // 32 rows x 64 columns per warp as two 32 x 32 blocks -> swizzled staging -> TMA stores (full
          // 64-byte row segments, clipped at Sq); the rounding residual reuses the blocks afterwards
          const uint32_t stg = smem_u32(smem + FWD_SMEM_STG + warp * 4096);
          const int row0 = qt * 128 + warp * 32;
          if (lane == 0) tma_store_wait_read<0>();  // the previous item's stores have drained the blocks
          __syncwarp();
          if (t == 0) FTRACE(16 * it + 4);
          stage_block32<false>(stg, lane, r0, inv);
          stage_block32<false>(stg + 2048, lane, r1, inv);
          if (t == 0) FTRACE(16 * it + 5);
          fence_proxy_async_smem();
          __syncwarp();
          if (t == 0) FTRACE(16 * it + 6);
          if (lane == 0) {
            tma_store_3d(&tmap_o, smem + FWD_SMEM_STG + warp * 4096, h * 64, row0, b);
            tma_store_3d(&tmap_o, smem + FWD_SMEM_STG + warp * 4096 + 2048, h * 64 + 32, row0, b);
            tma_store_commit();
          }
          if (t == 0) FTRACE(16 * it + 14);
          if (p.out_lo) {
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
            if (t == 0) FTRACE(16 * it + 15);
            stage_block32<true>(stg, lane, r0, inv);
            stage_block32<true>(stg + 2048, lane, r1, inv);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&tmap_olo, smem + FWD_SMEM_STG + warp * 4096, h * 64, row0, b);
              tma_store_3d(&tmap_olo, smem + FWD_SMEM_STG + warp * 4096 + 2048, h * 64 + 32, row0, b);
              tma_store_commit();
            }
          }
          if (qrow < p.Sq && p.lse)
            p.lse[(static_cast<long long>(b) * p.H + h) * p.Sq + qrow] = mx * p.scale + __logf(sum);
        } else if (row_ok) {
          const long long ooff = (packed ? grow : static_cast<long long>(b) * p.Sq + qrow) * p.ldo + h * 64;
          __nv_bfloat16* op = p.out + ooff;
          // 8 context values -> 16 bytes of bf16, and (training) 16 bytes of their bf16 rounding
          // residuals: hi + lo carries 16 significant bits of O, which the backward's
          // delta = rowsum(O * dO) needs (attn_delta_kernel)
          auto emit8 = [&](const uint32_t (&r)[32], int i, int col) {
            float x[8];
            uint32_t hi[4];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(r[i + e]) * inv;
#pragma unroll
            for (int e = 0; e < 4; ++e) hi[e] = pack_bf16(x[2 * e], x[2 * e + 1]);
            *reinterpret_cast<uint4*>(op + col) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (p.out_lo) {
              uint32_t lo[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) lo[e] = pack_bf16(x[2 * e] - bf16_lo(hi[e]), x[2 * e + 1] - bf16_hi(hi[e]));
              *reinterpret_cast<uint4*>(p.out_lo + ooff + col) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          };
#pragma unroll
          for (int i = 0; i < 32; i += 8) emit8(r0, i, i);
#pragma unroll
          for (int i = 0; i < 32; i += 8) emit8(r1, i, 32 + i);
          if (p.lse) {
            const long long bi = packed ? grow / p.Sq : b;                 // image, token inside it
            const long long ti = packed ? grow - bi * p.Sq : qrow;
            p.lse[(bi * p.H + h) * p.Sq + ti] = mx * p.scale + __logf(sum);
          }
        }
      }
      if (t == 0) FTRACE(16 * it + 13);
    }
  }
  if constexpr (STAGED) {
    if (warp < 4 && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
struct AttnBwdParams {
  const float* lse;                                                 // [B,H,Sq]
  const float* delta;                                               // [B,H,Sq] rowsum(O * dO), attn_delta_kernel
  __nv_bfloat16* dq; long long lddq;                                // [B*Sq, lddq]
  __nv_bfloat16* dk; long long lddk;                                // [B*Sk, lddk]
  __nv_bfloat16* dv; long long lddv;
  int B, H, Sq, Sk;
  float scale;
  int G, GS;  // packed short sequences: images per 128-row tile and G * S (see AttnFwdParams::G)
};

constexpr int BWD_MATH_WARPS = 16;  // 4 per TMEM lane quadrant, 32 of the 128 key columns each
constexpr int BWD_WARP_MMA = BWD_MATH_WARPS;      // single-thread tcgen05 issuer
constexpr int BWD_WARP_TMA = BWD_MATH_WARPS + 1;  // single-thread TMA producer
constexpr int BWD_THREADS = 32 * (BWD_MATH_WARPS + 2);
constexpr int BWD_TILE = 16384;  // one 128-row x 64-column bf16 tile (128-byte swizzled rows)
constexpr int BWD_SMEM_Q = 0 * BWD_TILE;    // 2 slots each: Q, dO, O (query tiles), K, V (key tiles)
constexpr int BWD_SMEM_DO = 2 * BWD_TILE;
constexpr int BWD_SMEM_O = 4 * BWD_TILE;    // (no O tiles any more: delta has its own pass) staging of the accumulator drains
constexpr int BWD_SMEM_K = 6 * BWD_TILE;
constexpr int BWD_SMEM_VV = 8 * BWD_TILE;
constexpr int BWD_SMEM_P = 10 * BWD_TILE;   // P and dS of the current iteration: 2 key blocks x 16 KB each
constexpr int BWD_SMEM_DS = 12 * BWD_TILE;
constexpr int BWD_SMEM_BAR = 14 * BWD_TILE;
constexpr int BWD_SMEM_BYTES = BWD_SMEM_BAR + 256 + 1024;

// 16-byte shared-memory accesses with the state space spelled out (the generic-address forms
// cost an address-space check per access)
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Persistent: one CTA per SM loops over (batch, head) items. Per item the (key tile j, query tile
// i) iterations run j-outer as before, but the three roles are decoupled so that nothing on the
// critical path waits for a round trip:
//   TMA warp  : refills a tile slot as soon as the last MMA that reads it has retired (per-slot
//               "free" barriers fed by tcgen05.commit), so the next item's Q/dO/O/K/V arrive while
//               the current item is still computing;
//   MMA warp  : issues S = Q K^T and dP = dO V^T of iteration g+1 as soon as the math warps hold
//               S/dP of iteration g in registers (bar_sdp_read), i.e. underneath their exp math,
//               then dV/dK/dQ of iteration g when P/dS are staged;
//   math warps: recompute P from the saved LSE, dS = P (dP - delta), stage both as bf16 MMA
//               operands, and drain dV/dK (per key tile) and dQ (per item) from TMEM.
// With a single 128x128 tile per item (S <= 128, the DINO local crops) the items alternate
// between the two tile slots, so loads are double-buffered there too.
// PACKED (S <= 64 self-attention): G images per tile. Their token rows are contiguous, so the tile of
// group g is rows [g*G*S, +128) of the [B*S] row matrix, one iteration per item, and P / dS are
// block-diagonal with block size S; rows and keys past G*S belong to the next group: masked, and
// never stored (the last partly valid 32-row block of an accumulator leaves through the *_t maps,
// whose box has G*S % 32 rows).
template <bool PACKED>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                const __grid_constant__ CUtensorMap tmap_dq,
                const __grid_constant__ CUtensorMap tmap_dk, const __grid_constant__ CUtensorMap tmap_dv,
                const __grid_constant__ CUtensorMap tmap_dq_t,
                const __grid_constant__ CUtensorMap tmap_dk_t, const __grid_constant__ CUtensorMap tmap_dv_t,
                const AttnBwdParams p) {
  constexpr bool packed = PACKED;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + BWD_SMEM_BAR);
  uint64_t* bar_ldq = bar + 0;       // [2] Q, dO, O of a query-tile slot landed
  uint64_t* bar_ldkv = bar + 2;      // [2] K, V of a key-tile slot landed
  uint64_t* bar_freeq = bar + 4;     // [2] last MMA reading the slot retired
  uint64_t* bar_freekv = bar + 6;    // [2]
  uint64_t* bar_sdp_full = bar + 8;  // S / dP of an iteration are in TMEM
  uint64_t* bar_sdp_read = bar + 9;  // ... and now in the math warps' registers
  uint64_t* bar_pds_ready = bar + 10;  // P / dS staged in shared memory
  uint64_t* bar_pds_free = bar + 11;   // dV / dK / dQ MMAs of an iteration retired
  uint64_t* bar_dkv_free = bar + 12;   // dV / dK accumulators drained
  uint64_t* bar_dq_free = bar + 13;    // dQ accumulators drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 14);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = (p.Sq + 127) >> 7, nk = (p.Sk + 127) >> 7;  // each 1 or 2
  const int nit = nq * nk;
  const bool alt = nit == 1;  // single-tile items alternate between the two slots
  const int items = (packed ? (p.B + p.G - 1) / p.G : p.B) * p.H;  // packed: (group, head)
  // iteration it -> (key tile j, query tile i), j outer
  auto it_j = [&](int it) { return nq == 2 ? it >> 1 : it; };
  auto it_i = [&](int it) { return nq == 2 ? it & 1 : 0; };

  if (warp == BWD_WARP_TMA) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k);
      tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_do);
      tma_prefetch_desc(&tmap_dq); tma_prefetch_desc(&tmap_dk); tma_prefetch_desc(&tmap_dv);
      if (packed) { tma_prefetch_desc(&tmap_dq_t); tma_prefetch_desc(&tmap_dk_t); tma_prefetch_desc(&tmap_dv_t); }
      for (int i = 0; i < 8; ++i) mbar_init(bar + i, 1);
      mbar_init(bar_sdp_full, 1); mbar_init(bar_sdp_read, BWD_MATH_WARPS);
      mbar_init(bar_pds_ready, BWD_MATH_WARPS); mbar_init(bar_pds_free, 1);
      mbar_init(bar_dkv_free, BWD_MATH_WARPS); mbar_init(bar_dq_free, BWD_MATH_WARPS);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // The CTA owns the SM's whole tensor memory (512 columns, one CTA per SM), so the allocation
  // starts at lane 0 / column 0. Treating that as a constant keeps every tcgen05.mma operand of the
  // issuing thread in uniform registers (a base loaded from shared memory costs a vector->uniform
  // hand-off sequence per MMA).
  if (*tmem_slot != 0u) {
    if (threadIdx.x == 0) printf("vitssl: attention_bwd expects the TMEM allocation at column 0\n");
    __trap();
  }
  pdl_wait();  // the prologue overlapped the previous kernel's tail; global memory from here on
  constexpr uint32_t tmem = 0;
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 320, COL_DQ = 384;

  if (warp == BWD_WARP_TMA) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int n = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
        const int h = item % p.H, b = item / p.H;
        const int u = alt ? n >> 1 : n;  // how often this item's slots have been used before
        auto load_kv = [&](int j) {
          const int sl = alt ? (n & 1) : j;
          if (u > 0) mbar_wait_parked(&bar_freekv[sl], (u - 1) & 1);
          mbar_expect_tx(&bar_ldkv[sl], 2 * BWD_TILE);
          tma_load_3d(smem + BWD_SMEM_K + sl * BWD_TILE, &tmap_k, &bar_ldkv[sl], h * 64, packed ? b * p.GS : j * 128, packed ? 0 : b);
          tma_load_3d(smem + BWD_SMEM_VV + sl * BWD_TILE, &tmap_v, &bar_ldkv[sl], h * 64, packed ? b * p.GS : j * 128, packed ? 0 : b);
        };
        auto load_q = [&](int i) {
          const int sl = alt ? (n & 1) : i;
          if (u > 0) mbar_wait_parked(&bar_freeq[sl], (u - 1) & 1);
          mbar_expect_tx(&bar_ldq[sl], 2 * BWD_TILE);
          tma_load_3d(smem + BWD_SMEM_Q + sl * BWD_TILE, &tmap_q, &bar_ldq[sl], h * 64, packed ? b * p.GS : i * 128, packed ? 0 : b);
          tma_load_3d(smem + BWD_SMEM_DO + sl * BWD_TILE, &tmap_do, &bar_ldq[sl], h * 64, packed ? b * p.GS : i * 128, packed ? 0 : b);
        };
        // in the order the slots are released by the previous item and first needed by this one
        load_kv(0);
        load_q(0);
        if (nq > 1) load_q(1);
        if (nk > 1) load_kv(1);
      }
    }
  } else if (warp == BWD_WARP_MMA) {
    // ------------------------------ MMA issuer ------------------------------
    // The whole warp walks the iterations and the barrier waits (uniform control flow) and one elected
    // lane issues, so every tcgen05.mma is a bare UTCHMMA with uniform-register operands (inside an
    // `if (lane == 0)` region each one is compiled into its own ELECT / BRA.U.ANY loop). Measured
    // neutral (159 us either way): the issuing thread is back-pressured by the tensor pipe, which is
    // itself paced by shared-memory operand reads — S / dP: 8 MMAs x 8 KB of operands = 64 KB, dV / dK /
    // dQ: 24 x 6 KB = 144 KB per iteration against 128 B/clk, i.e. ~700 + ~1 500 cycles in the trace.
    {
      constexpr uint32_t idesc_sdp = umma_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_dkv = umma_idesc_bf16(128, 64, true, true);
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 64, false, true);
      // Descriptor low words (start address >> 4 | leading-byte-offset field) of every operand
      // tile, computed once; advancing inside a tile adds (bytes >> 4). All forms share the high
      // word (SBO = 1024, version 1, 128-byte swizzle).
      constexpr uint32_t HI = 0x40004040u;
      constexpr uint32_t LBO_K = (16u >> 4) << 16;        // K-major operand
      constexpr uint32_t LBO_MN = (8192u >> 4) << 16;     // MN-major, 64-wide groups 8 KB apart
      constexpr uint32_t LBO_MN2 = (16384u >> 4) << 16;   // MN-major P / dS: key blocks 16 KB apart
      const uint32_t sb = smem_u32(smem) >> 4;
      auto lo = [&](int byte_off, uint32_t lbo) { return (sb + (static_cast<uint32_t>(byte_off) >> 4)) | lbo; };
      // S / dP of iteration `it` of the CTA's n-th item; waits for tiles this iteration uses first
      auto issue_sdp = [&](int n, int it) {
        const int j = it_j(it), i = it_i(it);
        const int u = alt ? n >> 1 : n;
        const int slq = alt ? (n & 1) : i, slk = alt ? (n & 1) : j;
        if (j == 0) mbar_wait(&bar_ldq[slq], u & 1);
        if (i == 0) mbar_wait(&bar_ldkv[slk], u & 1);
        tc_fence_after();
        const uint32_t q_k = lo(BWD_SMEM_Q + slq * BWD_TILE, LBO_K), do_k = lo(BWD_SMEM_DO + slq * BWD_TILE, LBO_K);
        const uint32_t k_k = lo(BWD_SMEM_K + slk * BWD_TILE, LBO_K), v_k = lo(BWD_SMEM_VV + slk * BWD_TILE, LBO_K);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss_lo<false>(COL_S, q_k + 2 * k, k_k + 2 * k, HI, idesc_sdp, k > 0);    // S = Q_i K_j^T
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss_lo<false>(COL_DP, do_k + 2 * k, v_k + 2 * k, HI, idesc_sdp, k > 0);  // dP = dO_i V_j^T
          umma_commit(bar_sdp_full);
        }
        __syncwarp();
      };
      // the next item's first S / dP may be issued under the current item's last iteration only
      // if its tiles are released by earlier iterations (true for the 2 x 2 tiling)
      const bool early_cross = nq == 2 && nk == 2;
      const uint32_t p_mn = lo(BWD_SMEM_P, LBO_MN2), ds_mn = lo(BWD_SMEM_DS, LBO_MN2), ds_k = lo(BWD_SMEM_DS, LBO_K);
      uint32_t g = 0, kt = 0;  // iterations / key tiles processed by this CTA so far
      int n = 0;
      if (blockIdx.x < items) issue_sdp(0, 0);
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
        const bool more_items = item + static_cast<int>(gridDim.x) < items;
        for (int it = 0; it < nit; ++it, ++g) {
          const int j = it_j(it), i = it_i(it);
          const bool in_item = it + 1 < nit;
          const bool have_next = in_item || more_items;
          const bool early = in_item || early_cross;
          if (have_next && early) {
            mbar_wait(bar_sdp_read, g & 1);
            if (lane == 0) TRACE(16 * g + 0);
            tc_fence_after();
            if (in_item) issue_sdp(n, it + 1); else issue_sdp(n + 1, 0);
          }
          if (lane == 0) TRACE(16 * g + 1);
          mbar_wait(bar_pds_ready, g & 1);
          if (lane == 0) TRACE(16 * g + 2);
          if (i == 0 && kt > 0) mbar_wait(bar_dkv_free, (kt - 1) & 1);  // previous key tile drained
          if (it == 0 && n > 0) mbar_wait(bar_dq_free, (n - 1) & 1);     // previous item's dQ drained
          tc_fence_after();
          const int slq = alt ? (n & 1) : i, slk = alt ? (n & 1) : j;
          const uint32_t q_mn = lo(BWD_SMEM_Q + slq * BWD_TILE, LBO_MN), do_mn = lo(BWD_SMEM_DO + slq * BWD_TILE, LBO_MN);
          const uint32_t k_mn = lo(BWD_SMEM_K + slk * BWD_TILE, LBO_MN);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)  // dV_j += P^T dO_i   (reduction over the 128 query rows)
              umma_bf16_ss_lo<false>(COL_DV, p_mn + 128 * ks, do_mn + 128 * ks, HI, idesc_dkv, i > 0 || ks > 0);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)  // dK_j += dS^T Q_i
              umma_bf16_ss_lo<false>(COL_DK, ds_mn + 128 * ks, q_mn + 128 * ks, HI, idesc_dkv, i > 0 || ks > 0);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)  // dQ_i += dS K_j     (reduction over the 128 keys)
              umma_bf16_ss_lo<false>(COL_DQ + i * 64, ds_k + (ks >> 2) * 1024 + (ks & 3) * 2, k_mn + 128 * ks, HI, idesc_dq,
                                     j > 0 || ks > 0);
            umma_commit(bar_pds_free);
            if (j == nk - 1) umma_commit(&bar_freeq[slq]);   // last reader of Q_i / dO_i
            if (i == nq - 1) umma_commit(&bar_freekv[slk]);  // last reader of K_j / V_j
          }
          __syncwarp();
          if (lane == 0) TRACE(16 * g + 3);
          if (have_next && !early) issue_sdp(n + 1, 0);    // (S / dP were read before P / dS were staged)
          if (i == nq - 1) ++kt;
        }
      }
    }
  } else {
    // 512 math threads: row r = TMEM lane, `cq` selects 32 of the 128 key columns of the tile
    const int r = threadIdx.x & 127, cq = threadIdx.x >> 7;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t sbar = sbase + BWD_SMEM_BAR;  // barrier k lives at sbar + 8k
    // this thread's 4 x 16-byte chunks (32 keys) inside key block (cq >> 1) of the P / dS tiles
    const uint32_t sP = sbase + BWD_SMEM_P + (cq >> 1) * 16384 + r * 128;
    uint32_t g = 0;  // iterations processed by this CTA so far
    int n = 0;       // items processed

    // TMEM accumulators -> global: dV_j / dK_j when a key tile is complete (warps 0-7 take dV,
    // warps 8-15 dK; 32 of the 64 columns each), dQ when the item is complete (thread group cq
    // takes columns (cq & 1) * 32 .. +32 of query tile cq >> 1). The drains of iteration g run
    // inside iteration g+1, after its first exp math and before its P / dS stores, so waiting for
    // the MMAs of iteration g costs nothing; each warp stages its 32 x 32 bf16 block (64-byte
    // swizzle) and hands it to a TMA store, which clips at the sequence end. (Per-thread 16-byte
    // global stores hit 32 different lines per instruction and took ~2500 cycles per drain in the LSU.)
    auto stage32 = [&](uint32_t dst, const uint32_t (&v)[32], float mul) {
      const uint32_t base = dst + lane * 64;
      const int sw = (lane >> 1) & 3;
      const f32x2 m2 = pk2(mul, mul);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float x0, x1;
          upk2(fmul2(pk2(__uint_as_float(v[8 * q4 + 2 * e]), __uint_as_float(v[8 * q4 + 2 * e + 1])), m2), x0, x1);
          w[e] = pack_bf16(x0, x1);
        }
        sts128(base + ((q4 ^ sw) << 4), w[0], w[1], w[2], w[3]);
      }
    };
    // Staging lives in the two tile slots that held the O tiles before delta moved to its own pass
    // (16 warps x 2 KB): the P / dS region is never touched, so a drain needs neither a wait for the
    // TMA engine to finish reading nor a CTA-wide barrier before P / dS of the current iteration are
    // stored (per-iteration trace of the former form: ~2 500-3 000 of ~6 000 cycles in the two
    // draining iterations of an item). A warp only waits for ITS previous store before it restages.
    auto drain = [&](int item, int j, bool item_done) {
      const int h = item % p.H, b = item / p.H;
      const int sel = cq >> 1, cc = cq & 1;
      uint8_t* stg = smem + BWD_SMEM_O + warp * 2048;
      const int row0 = (warp & 3) * 32, col0 = h * 64 + cc * 32;
      const bool dq_mine = item_done && sel < nq;
      uint32_t v[32];
      __syncwarp();
      tmem_ld_32x32(lane_addr + (sel == 0 ? COL_DV : COL_DK) + cc * 32, v);
      if (lane == 0) tma_store_wait_read<0>();  // this warp's block: its previous store has read it (long ago)
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_addr(sbar + 8 * 12);  // bar_dkv_free: accumulators are in registers
      stage32(smem_u32(stg), v, sel == 0 ? 1.0f : p.scale);
      fence_proxy_async_smem();
      __syncwarp();
      // packed: only the G*S rows of this group may be written (the tile's remaining rows are the next
      // group's): whole 32-row blocks through the 32-row maps, the partly valid one through the tail maps
      const int prow = b * p.GS + row0, pvalid = p.GS - row0;
      if (lane == 0) {
        if (!packed) tma_store_3d(sel == 0 ? &tmap_dv : &tmap_dk, stg, col0, j * 128 + row0, b);
        else if (pvalid >= 32) tma_store_3d(sel == 0 ? &tmap_dv : &tmap_dk, stg, col0, prow, 0);
        else if (pvalid > 0) tma_store_3d(sel == 0 ? &tmap_dv_t : &tmap_dk_t, stg, col0, prow, 0);
        tma_store_commit();
      }
      if (item_done) {
        __syncwarp();
        if (dq_mine) {
          tmem_ld_32x32(lane_addr + COL_DQ + sel * 64 + cc * 32, v);
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_addr(sbar + 8 * 13);  // bar_dq_free
        if (dq_mine) {
          if (lane == 0) tma_store_wait_read<0>();  // the dV / dK store above has read the block
          __syncwarp();
          stage32(smem_u32(stg), v, p.scale);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (!packed) tma_store_3d(&tmap_dq, stg, col0, sel * 128 + row0, b);
            else if (pvalid >= 32) tma_store_3d(&tmap_dq, stg, col0, prow, 0);
            else if (pvalid > 0) tma_store_3d(&tmap_dq_t, stg, col0, prow, 0);
            tma_store_commit();
          }
        }
      }
    };

    // -lse * log2(e) and delta = rowsum(O * dO) of this thread's two query rows ([B, H, Sq] = [item, Sq]
    // arrays; delta comes from attn_delta_kernel, which evaluates it from the hi + lo context so that
    // its error stays far below the (dP - delta) cancellation it enters); the next item's values are
    // fetched during the current item's last iteration
    // index of this thread's query row of tile i of `item` in the [B, H, Sq] arrays, or -1
    auto stat_index = [&](int item, int i) -> long long {
      if (packed) {
        const int h = item % p.H, g = item / p.H;
        const long long grow = static_cast<long long>(g) * p.GS + r;   // global token row
        if (i != 0 || r >= p.GS || grow >= static_cast<long long>(p.B) * p.Sq) return -1;
        const long long bi = grow / p.Sq;
        return (bi * p.H + h) * p.Sq + (grow - bi * p.Sq);
      }
      const int qrow = i * 128 + r;
      return (i < nq && qrow < p.Sq) ? static_cast<long long>(item) * p.Sq + qrow : -1;
    };
    auto load_nlse = [&](int item, int i) {
      const long long idx = stat_index(item, i);
      return idx >= 0 ? -p.lse[idx] * kLog2e : -INFINITY;
    };
    auto load_delta = [&](int item, int i) {
      const long long idx = stat_index(item, i);
      return idx >= 0 ? p.delta[idx] : 0.f;
    };
    // packed: this thread's keys are those of its own image, [klo, khi); the warp's 32 rows touch the
    // images between its first and last valid row, [wlo, whi) (warp-uniform: TMEM loads are collective)
    int klo = 0, khi = 0, wlo = 0, whi = 0;
    if (packed) {
      const int w0 = (warp & 3) * 32;
      if (r < p.GS) { klo = (r / p.Sq) * p.Sq; khi = klo + p.Sq; }
      if (w0 < p.GS) {
        const int wl = w0 + 31 < p.GS - 1 ? w0 + 31 : p.GS - 1;
        wlo = (w0 / p.Sq) * p.Sq;
        whi = (wl / p.Sq) * p.Sq + p.Sq;
      }
    }
    float nx0 = 0.f, nx1 = 0.f, nd0 = 0.f, nd1 = 0.f;
    if (blockIdx.x < items) {
      nx0 = load_nlse(blockIdx.x, 0); nx1 = load_nlse(blockIdx.x, 1);
      nd0 = load_delta(blockIdx.x, 0); nd1 = load_delta(blockIdx.x, 1);
    }

    for (int item = blockIdx.x; item < items; item += gridDim.x, ++n) {
      const float delta0 = nd0, delta1 = nd1;
      const float nlse0 = nx0, nlse1 = nx1;  // -inf on invalid rows -> p = 0
      for (int it = 0; it < nit; ++it, ++g) {
        const int j = it_j(it), i = it_i(it);
        if (it == nit - 1 && item + static_cast<int>(gridDim.x) < items) {
          nx0 = load_nlse(item + gridDim.x, 0);
          nx1 = load_nlse(item + gridDim.x, 1);
          nd0 = load_delta(item + gridDim.x, 0);
          nd1 = load_delta(item + gridDim.x, 1);
        }
        if (threadIdx.x == 0) TRACE(16 * g + 8);
        mbar_wait_parked_addr(sbar + 8 * 8, g & 1);  // bar_sdp_full
        if (threadIdx.x == 0) TRACE(16 * g + 9);
        tc_fence_after();
        const float sl2 = p.scale * kLog2e;
        const float dl = i ? delta1 : delta0, nl = i ? nlse1 : nlse0;
        const f32x2 sl2v = pk2(sl2, sl2), nlv = pk2(nl, nl), ndlv = pk2(-dl, -dl);
        const int keyb = j * 128 + cq * 32;  // this thread's first key
        // a warp whose 32 query rows all lie past Sq (S = 196: the last quadrant of the second query
        // tile; S = 37: three of four quadrants) has nothing to compute: its P / dS rows are zero. It
        // keeps every barrier hand-off but skips the TMEM reads and the exp math (key_lim = 0).
        const int key_lim = (i * 128 + (warp & 3) * 32 < p.Sq) ? p.Sk : 0;
        // 8 keys: P = 2^(S * scale * log2e - lse * log2e), dS = P * (dP - delta), one 16-byte store each
        auto chunk_math = [&](const uint32_t (&sv)[8], const uint32_t (&dp)[8], int c, uint32_t (&pp)[4], uint32_t (&dd)[4]) {
          const int key0 = keyb + c * 8;
          const bool some = packed ? (key0 < khi && key0 + 8 > klo) : key0 < key_lim;
          if (some) {
            const bool full = packed ? (key0 >= klo && key0 + 8 <= khi) : key0 + 8 <= key_lim;
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
              float a0, a1;
              upk2(ffma2(pk2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), sl2v, nlv), a0, a1);
              float p0 = ex2_approx(a0), p1 = ex2_approx(a1);
              if (!full) {
                if (packed) {
                  p0 = static_cast<unsigned>(key0 + e - klo) < static_cast<unsigned>(khi - klo) ? p0 : 0.f;
                  p1 = static_cast<unsigned>(key0 + e + 1 - klo) < static_cast<unsigned>(khi - klo) ? p1 : 0.f;
                } else {
                  p0 = (key0 + e < key_lim) ? p0 : 0.f;
                  p1 = (key0 + e + 1 < key_lim) ? p1 : 0.f;
                }
              }
              pp[e >> 1] = pack_bf16(p0, p1);
              float d0, d1;
              upk2(fmul2(pk2(p0, p1), fadd2(pk2(__uint_as_float(dp[e]), __uint_as_float(dp[e + 1])), ndlv)), d0, d1);
              dd[e >> 1] = pack_bf16(d0, d1);
            }
          } else {  // key chunk entirely past Sk
#pragma unroll
            for (int e = 0; e < 4; ++e) pp[e] = dd[e] = 0u;
          }
        };
        auto chunk_store = [&](int c, const uint32_t (&pp)[4], const uint32_t (&dd)[4]) {
          const int sw = (((cq & 1) * 4 + c) ^ (r & 7)) << 4;
          sts128(sP + sw, pp[0], pp[1], pp[2], pp[3]);
          sts128(sP + (BWD_SMEM_DS - BWD_SMEM_P) + sw, dd[0], dd[1], dd[2], dd[3]);
        };
        const uint32_t ts = lane_addr + COL_S + cq * 32, td = lane_addr + COL_DP + cq * 32;
        {
          // first half (16 keys) in ONE TMEM round trip; its exp math overlaps the previous iteration's
          // dV / dK / dQ MMAs. Its stores need those MMAs' P / dS operands consumed; they go out before
          // that iteration's finished accumulators are drained, so the drain's 32 registers never
          // coexist with pending results.
          uint32_t sv0[8], dp0[8], sv1[8], dp1[8], pp0[4], dd0[4], pp1[4], dd1[4];
          __syncwarp();
          if (packed ? (keyb < whi && keyb + 16 > wlo) : keyb < key_lim) {
            tmem_ld_32x8(ts, sv0);
            tmem_ld_32x8(td, dp0);
            if (packed || keyb + 8 < key_lim) {
              tmem_ld_32x8(ts + 8, sv1);
              tmem_ld_32x8(td + 8, dp1);
            }
            tmem_ld_wait();
          }
          chunk_math(sv0, dp0, 0, pp0, dd0);
          chunk_math(sv1, dp1, 1, pp1, dd1);
          if (g > 0) {
            if (threadIdx.x == 0) TRACE(16 * g + 10);
            mbar_wait_parked_addr(sbar + 8 * 11, (g - 1) & 1);  // bar_pds_free
            if (threadIdx.x == 0) TRACE(16 * g + 11);
            tc_fence_after();
          }
          chunk_store(0, pp0, dd0);
          chunk_store(1, pp1, dd1);
        }
        if (g > 0) {
          if (it == 0) drain(item - static_cast<int>(gridDim.x), nk - 1, true);
          else if (i == 0) drain(item, j - 1, false);
        }
        {
          // the last two sub-chunks are fetched together as well, so S / dP can be handed back to the
          // MMA warp (next iteration's S / dP) with half of this iteration's exp math still to do
          uint32_t sv2[8], dp2[8], sv3[8], dp3[8], pp[4], dd[4];
          __syncwarp();
          if (packed ? (keyb + 16 < whi && keyb + 32 > wlo) : keyb + 16 < key_lim) {
            tmem_ld_32x8(ts + 16, sv2);
            tmem_ld_32x8(td + 16, dp2);
            if (packed || keyb + 24 < key_lim) {
              tmem_ld_32x8(ts + 24, sv3);
              tmem_ld_32x8(td + 24, dp3);
            }
            tmem_ld_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_addr(sbar + 8 * 9);  // bar_sdp_read
          chunk_math(sv2, dp2, 2, pp, dd);
          chunk_store(2, pp, dd);
          chunk_math(sv3, dp3, 3, pp, dd);
          chunk_store(3, pp, dd);
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (threadIdx.x == 0) TRACE(16 * g + 12);
        if (lane == 0) mbar_arrive_addr(sbar + 8 * 10);  // bar_pds_ready: one arrival per warp
      }
    }
    if (g > 0) {  // the CTA's last item
      mbar_wait_parked_addr(sbar + 8 * 11, (g - 1) & 1);
      tc_fence_after();
      drain(blockIdx.x + (n - 1) * static_cast<int>(gridDim.x), nk - 1, true);
      if (lane == 0) tma_store_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == BWD_WARP_TMA) {
    tc_fence_after();
    tmem_dealloc<512>(0u);
  }
}

// ------------------------------------------------------------------------------------------
// delta[b, h, s] = sum_d O[b, s, h, d] * dO[b, s, h, d] — the row term of the softmax backward,
// dS = P * (dP - delta). The backward kernel used to take it from the bf16 context tiles; on real
// activations (tokens share a large common component, so dP - delta cancels to a small difference)
// the 2^-9 rounding of O then dominates the error of dQ: measured 4.8e-2 relative on a late ViT-Ti
// block against 5e-3 for the reference under autocast. The forward therefore also stores the bf16
// rounding residual of the context; O = hi + lo has 16 significant bits and the same case measures
// 8.5e-3 (4.3e-3 is the floor of this formulation). HBM-bound: 6 bytes per context element.
// One warp per token row, 8 elements per lane per step; a head is 8 consecutive lanes.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_delta_kernel(const __nv_bfloat16* __restrict__ o,
                                                         const __nv_bfloat16* __restrict__ o_lo,
                                                         const __nv_bfloat16* __restrict__ d_o, long long ldo,
                                                         float* __restrict__ delta, long long rows, int H, int S) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long b = row / S;
  const int sidx = static_cast<int>(row - b * S);
  const int groups = H * 8;
  for (int g0 = 0; g0 < groups; g0 += 32) {
    const int g = g0 + lane;
    float acc = 0.f;
    if (g < groups) {
      const long long off = row * ldo + g * 8;
      const uint4 a = *reinterpret_cast<const uint4*>(o + off);
      const uint4 d = *reinterpret_cast<const uint4*>(d_o + off);
      uint4 l = make_uint4(0u, 0u, 0u, 0u);
      if (o_lo) l = *reinterpret_cast<const uint4*>(o_lo + off);
      f32x2 acc2 = pk2(0.f, 0.f);
      acc2 = ffma2(fadd2(pk2(bf16_lo(a.x), bf16_hi(a.x)), pk2(bf16_lo(l.x), bf16_hi(l.x))), pk2(bf16_lo(d.x), bf16_hi(d.x)), acc2);
      acc2 = ffma2(fadd2(pk2(bf16_lo(a.y), bf16_hi(a.y)), pk2(bf16_lo(l.y), bf16_hi(l.y))), pk2(bf16_lo(d.y), bf16_hi(d.y)), acc2);
      acc2 = ffma2(fadd2(pk2(bf16_lo(a.z), bf16_hi(a.z)), pk2(bf16_lo(l.z), bf16_hi(l.z))), pk2(bf16_lo(d.z), bf16_hi(d.z)), acc2);
      acc2 = ffma2(fadd2(pk2(bf16_lo(a.w), bf16_hi(a.w)), pk2(bf16_lo(l.w), bf16_hi(l.w))), pk2(bf16_lo(d.w), bf16_hi(d.w)), acc2);
      float a0, a1;
      upk2(acc2, a0, a1);
      acc = a0 + a1;
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (g < groups && (lane & 7) == 0) delta[(b * H + (g >> 3)) * S + sidx] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// generic SIMT attention (any head dim <= 256, any lengths, explicit strides). One CTA per
// (query row, head, batch). Used for: return_attn=True (probabilities materialised, fp32),
// head dims other than 64, and S > 256. attention.py:20-27.
// ------------------------------------------------------------------------------------------
struct AttnGenericParams {
  const __nv_bfloat16 *q, *k, *v;
  long long q_sb, q_sh, q_ss, k_sb, k_sh, k_ss, v_sb, v_sh, v_ss;  // element strides (batch, head, row)
  __nv_bfloat16* out; long long o_sb, o_sh, o_ss;
  float* probs;  // nullable [B,H,Sq,Sk]
  float* lse;    // nullable [B,H,Sq]
  int B, H, Sq, Sk, d;
  float scale;
};

__global__ void __launch_bounds__(128) attn_generic_fwd_kernel(const AttnGenericParams p) {
  extern __shared__ float sm[];  // scores[Sk] | q[d] | red[8]
  float* scores = sm;
  float* qs = sm + p.Sk;
  float* red = qs + p.d;
  const int qi = blockIdx.x, h = blockIdx.y, b = blockIdx.z, t = threadIdx.x;
  const __nv_bfloat16* qp = p.q + b * p.q_sb + h * p.q_sh + qi * p.q_ss;
  for (int i = t; i < p.d; i += 128) qs[i] = __bfloat162float(qp[i]);
  __syncthreads();
  float mx = -INFINITY;
  for (int j = t; j < p.Sk; j += 128) {
    const __nv_bfloat16* kp = p.k + b * p.k_sb + h * p.k_sh + j * p.k_ss;
    float acc = 0.f;
    for (int i = 0; i < p.d; ++i) acc = fmaf(qs[i], __bfloat162float(kp[i]), acc);
    acc *= p.scale;
    scores[j] = acc;
    mx = fmaxf(mx, acc);
  }
  mx = warp_max(mx);
  if ((t & 31) == 0) red[t >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  float sum = 0.f;
  for (int j = t; j < p.Sk; j += 128) {
    const float e = __expf(scores[j] - mx);
    scores[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncthreads();
  if ((t & 31) == 0) red[4 + (t >> 5)] = sum;
  __syncthreads();
  sum = red[4] + red[5] + red[6] + red[7];
  const float inv = 1.0f / sum;
  if (p.probs) {
    float* pr = p.probs + ((static_cast<long long>(b) * p.H + h) * p.Sq + qi) * p.Sk;
    for (int j = t; j < p.Sk; j += 128) pr[j] = scores[j] * inv;
  }
  if (p.lse && t == 0) p.lse[(static_cast<long long>(b) * p.H + h) * p.Sq + qi] = mx + __logf(sum);
  __nv_bfloat16* op = p.out + b * p.o_sb + h * p.o_sh + qi * p.o_ss;
  for (int i = t; i < p.d; i += 128) {
    float acc = 0.f;
    const __nv_bfloat16* vp = p.v + b * p.v_sb + h * p.v_sh + i;
    // probabilities are rounded to bf16 before the PV product, as the reference does under autocast
    for (int j = 0; j < p.Sk; ++j) acc = fmaf(bf16_round(scores[j] * inv), __bfloat162float(vp[j * p.v_ss]), acc);
    op[i] = __float2bfloat16_rn(acc);
  }
}

struct AttnGenericBwdParams {
  AttnGenericParams f;          // q,k,v strides; out = forward output (for delta)
  const __nv_bfloat16* d_o;     // same strides as out
  const float* lse;
  __nv_bfloat16* dq;            // same strides as q
  float* dk_acc; float* dv_acc; // fp32 [B,Sk,H,d] dense accumulators (pre-zeroed)
};

__global__ void __launch_bounds__(128) attn_generic_bwd_kernel(const AttnGenericBwdParams a) {
  const AttnGenericParams& p = a.f;
  extern __shared__ float sm[];  // ds[Sk] | pj[Sk] | q[d] | do[d] | red[4]
  float* ds = sm;
  float* pj = sm + p.Sk;
  float* qs = pj + p.Sk;
  float* dos = qs + p.d;
  float* red = dos + p.d;
  const int qi = blockIdx.x, h = blockIdx.y, b = blockIdx.z, t = threadIdx.x;
  const __nv_bfloat16* qp = p.q + b * p.q_sb + h * p.q_sh + qi * p.q_ss;
  const __nv_bfloat16* dop = a.d_o + b * p.o_sb + h * p.o_sh + qi * p.o_ss;
  const __nv_bfloat16* op = p.out + b * p.o_sb + h * p.o_sh + qi * p.o_ss;
  float dl = 0.f;
  for (int i = t; i < p.d; i += 128) {
    qs[i] = __bfloat162float(qp[i]);
    dos[i] = __bfloat162float(dop[i]);
    dl += dos[i] * __bfloat162float(op[i]);
  }
  dl = warp_sum(dl);
  if ((t & 31) == 0) red[t >> 5] = dl;
  __syncthreads();
  dl = red[0] + red[1] + red[2] + red[3];
  const float lse = a.lse[(static_cast<long long>(b) * p.H + h) * p.Sq + qi];
  for (int j = t; j < p.Sk; j += 128) {
    const __nv_bfloat16* kp = p.k + b * p.k_sb + h * p.k_sh + j * p.k_ss;
    const __nv_bfloat16* vp = p.v + b * p.v_sb + h * p.v_sh + j * p.v_ss;
    float s = 0.f, dp = 0.f;
    for (int i = 0; i < p.d; ++i) {
      s = fmaf(qs[i], __bfloat162float(kp[i]), s);
      dp = fmaf(dos[i], __bfloat162float(vp[i]), dp);
    }
    const float pr = __expf(s * p.scale - lse);
    pj[j] = pr;
    ds[j] = pr * (dp - dl) * p.scale;
  }
  __syncthreads();
  __nv_bfloat16* dqp = a.dq + b * p.q_sb + h * p.q_sh + qi * p.q_ss;
  for (int i = t; i < p.d; i += 128) {
    float acc = 0.f;
    const __nv_bfloat16* kp = p.k + b * p.k_sb + h * p.k_sh + i;
    for (int j = 0; j < p.Sk; ++j) acc = fmaf(ds[j], __bfloat162float(kp[j * p.k_ss]), acc);
    dqp[i] = __float2bfloat16_rn(acc);
  }
  for (int idx = t; idx < p.Sk * p.d; idx += 128) {
    const int j = idx / p.d, i = idx % p.d;
    const long long o = ((static_cast<long long>(b) * p.Sk + j) * p.H + h) * p.d + i;
    atomicAdd(a.dk_acc + o, ds[j] * qs[i]);
    atomicAdd(a.dv_acc + o, pj[j] * dos[i]);
  }
}

int make_head_map(CUtensorMap* m, const void* base, int B, int S, int H, long long ld, int box_rows) {
  // tensor [B][S][H*64] with row pitch ld elements; 128B-swizzled box (64 cols, box_rows, 1)
  return make_tmap_bf16_3d(m, base, (uint64_t)H * 64, (uint64_t)S, (uint64_t)B, (uint64_t)ld * 2,
                           (uint64_t)S * ld * 2, 64, (uint32_t)box_rows, 1);
}

}  // namespace
}  // namespace vitssl

using namespace vitssl;

#ifdef VITSSL_ATTN_TRACE
extern "C" int vitssl_debug_attn_trace(long long* host_out, int n) {
  return cudaMemcpyFromSymbol(host_out, g_attn_trace, sizeof(long long) * n) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int vitssl_attention_supported(int64_t Sq, int64_t Sk, int64_t d_head) {
  return (d_head == 64 && Sk >= 1 && Sk <= 256 && Sq >= 1) ? 1 : 0;
}

extern "C" int vitssl_attention_fwd(const void* q, const void* k, const void* v, int64_t ldq,
                                    int64_t ldk, int64_t ldv, void* out, void* out_lo, int64_t ldo, float* lse,
                                    int64_t B, int64_t H, int64_t Sq, int64_t Sk, float scale,
                                    cudaStream_t stream) {
  VITSSL_REQUIRE(q && k && v && out, VITSSL_ERR_ARG, "attention_fwd: null pointer");
  VITSSL_REQUIRE(B > 0 && H > 0 && Sq > 0 && Sk > 0 && Sk <= 256, VITSSL_ERR_SHAPE,
                 "attention_fwd: unsupported shape B=%lld H=%lld Sq=%lld Sk=%lld (tcgen05 path needs Sk <= 256)",
                 (long long)B, (long long)H, (long long)Sq, (long long)Sk);
  VITSSL_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, VITSSL_ERR_SHAPE,
                 "attention_fwd: row pitches must be multiples of 8 elements");
  VITSSL_REQUIRE(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0) &&
                 ((uintptr_t)out % 16 == 0), VITSSL_ERR_ARG, "attention_fwd: pointers must be 16-byte aligned");
  AttnFwdParams p{};
  VITSSL_REQUIRE(out_lo == nullptr || (uintptr_t)out_lo % 16 == 0, VITSSL_ERR_ARG, "attention_fwd: out_lo must be 16-byte aligned");
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.out_lo = reinterpret_cast<__nv_bfloat16*>(out_lo); p.ldo = ldo; p.lse = lse;
  p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk;
  p.kv_rows = (int)((Sk + 15) / 16 * 16); p.scale = scale;
  // short self-attention sequences: G images per 128-row tile (VITSSL_ATTN_PACK=0 disables)
  static const int pack_env = getenv("VITSSL_ATTN_PACK") ? atoi(getenv("VITSSL_ATTN_PACK")) : 1;
  p.G = (pack_env != 0 && Sq == Sk && Sq <= 64 && B > 1) ? (int)(128 / Sq) : 1;
  if (p.G > B) p.G = (int)B;
  p.GS = p.G * (int)Sq;
  CUtensorMap mq, mk, mv;
  int rc;
  if (p.G > 1) {
    p.kv_rows = (p.GS + 15) / 16 * 16;
    // [B*S] token rows as ONE sequence: the tile of group g starts at row g*G*S
    if ((rc = make_head_map(&mq, q, 1, p.B * p.Sq, p.H, ldq, 128))) return rc;
    if ((rc = make_head_map(&mk, k, 1, p.B * p.Sk, p.H, ldk, p.kv_rows))) return rc;
    if ((rc = make_head_map(&mv, v, 1, p.B * p.Sk, p.H, ldv, p.kv_rows))) return rc;
  } else {
    if ((rc = make_head_map(&mq, q, p.B, p.Sq, p.H, ldq, 128))) return rc;
    if ((rc = make_head_map(&mk, k, p.B, p.Sk, p.H, ldk, p.kv_rows))) return rc;
    if ((rc = make_head_map(&mv, v, p.B, p.Sk, p.H, ldv, p.kv_rows))) return rc;
  }
  // the context (and its rounding residual) leave through 32 x 32 staging blocks + TMA stores when the
  // output rows are TMA-addressable (VITSSL_ATTN_FWD_STAGED=0 forces per-thread stores)
  static const int staged_env = getenv("VITSSL_ATTN_FWD_STAGED") ? atoi(getenv("VITSSL_ATTN_FWD_STAGED")) : 1;
  CUtensorMap mo, molo;
  memset(&mo, 0, sizeof(mo));
  memset(&molo, 0, sizeof(molo));
  // short sequences (S = 37: DINO local crops, cfg 1) keep per-thread stores: most of a 32-row block
  // would be clipped and the staging round trip costs more than it saves (58 vs 54 us at S = 37)
  p.staged = (p.G == 1 && (staged_env == 2 || (staged_env == 1 && Sq > 64))) ? 1 : 0;
  if (p.staged) {
    auto out_map = [&](CUtensorMap* m, void* base) {
      return make_tmap_bf16_3d_sw(m, base, (uint64_t)p.H * 64, (uint64_t)p.Sq, (uint64_t)p.B, (uint64_t)ldo * 2,
                                  (uint64_t)p.Sq * ldo * 2, 32, 32, 1, 64);
    };
    if ((rc = out_map(&mo, out))) return rc;
    if ((rc = out_map(&molo, out_lo ? out_lo : out))) return rc;
  }
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(attn_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BYTES);
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(attn_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BYTES);
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(attn_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BYTES);
    if (err != cudaSuccess) { set_error("attention_fwd: smem attribute: %s", cudaGetErrorString(err)); return VITSSL_ERR_CUDA; }
    configured = true;
  }
  const long long items = (p.G > 1 ? (B + p.G - 1) / p.G : B) * H * ((Sq + 127) / 128);
  const long long slots = 2ll * num_sms();  // persistent: two CTAs per SM
  const unsigned grid = (unsigned)(items < slots ? items : slots);
  cudaError_t lerr = p.G > 1
      ? launch_pdl(attn_fwd_kernel<false, true>, dim3(grid), dim3(FWD_THREADS), FWD_SMEM_BYTES, stream, mq, mk, mv, mo, molo, p)
      : p.staged
      ? launch_pdl(attn_fwd_kernel<true, false>, dim3(grid), dim3(FWD_THREADS), FWD_SMEM_BYTES, stream, mq, mk, mv, mo, molo, p)
      : launch_pdl(attn_fwd_kernel<false, false>, dim3(grid), dim3(FWD_THREADS), FWD_SMEM_BYTES, stream, mq, mk, mv, mo, molo, p);
  if (lerr != cudaSuccess) { set_error("attention_fwd: launch failed: %s", cudaGetErrorString(lerr)); return VITSSL_ERR_CUDA; }
  return check_launch("attention_fwd");
}

extern "C" int vitssl_attention_bwd(const void* q, const void* k, const void* v, int64_t ldq,
                                    int64_t ldk, int64_t ldv, const void* out, const void* out_lo, const void* d_out,
                                    int64_t ldo, const float* lse, float* delta, void* dq, int64_t lddq, void* dk,
                                    int64_t lddk, void* dv, int64_t lddv, int64_t B, int64_t H,
                                    int64_t Sq, int64_t Sk, float scale, cudaStream_t stream) {
  VITSSL_REQUIRE(q && k && v && out && d_out && lse && delta && dq && dk && dv, VITSSL_ERR_ARG, "attention_bwd: null pointer");
  VITSSL_REQUIRE(((uintptr_t)out % 16 == 0) && ((uintptr_t)d_out % 16 == 0) && ((uintptr_t)out_lo % 16 == 0), VITSSL_ERR_ARG,
                 "attention_bwd: out / out_lo / d_out must be 16-byte aligned");
  VITSSL_REQUIRE(B > 0 && H > 0 && Sq > 0 && Sk > 0 && Sk <= 256 && Sq <= 256, VITSSL_ERR_SHAPE,
                 "attention_bwd: unsupported shape Sq=%lld Sk=%lld (tcgen05 path needs both <= 256)",
                 (long long)Sq, (long long)Sk);
  VITSSL_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddq % 8 == 0 &&
                 lddk % 8 == 0 && lddv % 8 == 0, VITSSL_ERR_SHAPE, "attention_bwd: row pitches must be multiples of 8");
  AttnBwdParams p{};
  p.lse = lse; p.delta = delta;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq); p.lddq = lddq;
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.lddk = lddk;
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv); p.lddv = lddv;
  p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk; p.scale = scale;
  {
    // delta = rowsum(O * dO) per (batch, head, query) from the hi (+ lo) context: a small HBM-bound pass
    const long long rows = (long long)B * Sq;
    cudaError_t derr = launch_pdl(attn_delta_kernel, dim3((unsigned)((rows + 3) / 4)), dim3(128), 0, stream,
                                  reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(out_lo),
                                  reinterpret_cast<const __nv_bfloat16*>(d_out), (long long)ldo, delta, rows, (int)H, (int)Sq);
    if (derr != cudaSuccess) { set_error("attention_bwd: delta launch failed: %s", cudaGetErrorString(derr)); return VITSSL_ERR_CUDA; }
    const int rc0 = check_launch("attention_delta");
    if (rc0) return rc0;
  }
  // short self-attention sequences: G images per 128-row tile (VITSSL_ATTN_PACK=0 disables)
  static const int pack_env = getenv("VITSSL_ATTN_PACK") ? atoi(getenv("VITSSL_ATTN_PACK")) : 1;
  p.G = (pack_env != 0 && Sq == Sk && Sq <= 64 && B > 1) ? (int)(128 / Sq) : 1;
  if (p.G > B) p.G = (int)B;
  p.GS = p.G * (int)Sq;
  const bool packed = p.G > 1;
  // packed: the [B*S] token rows as one sequence (tile of group g = rows g*G*S ...)
  const int mapB = packed ? 1 : p.B, mapSq = packed ? p.B * p.Sq : p.Sq, mapSk = packed ? p.B * p.Sk : p.Sk;
  CUtensorMap mq, mk, mv, mdo;
  int rc;
  if ((rc = make_head_map(&mq, q, mapB, mapSq, p.H, ldq, 128))) return rc;
  if ((rc = make_head_map(&mk, k, mapB, mapSk, p.H, ldk, 128))) return rc;
  if ((rc = make_head_map(&mv, v, mapB, mapSk, p.H, ldv, 128))) return rc;
  if ((rc = make_head_map(&mdo, d_out, mapB, mapSq, p.H, ldo, 128))) return rc;
  // outputs leave through 32-column x 32-row staging blocks (64-byte swizzle), clipped at S; packed:
  // clipped at the end of the [B*S] rows, and the block that straddles G*S uses a map with fewer rows
  CUtensorMap mdq, mdk, mdv, mdq_t, mdk_t, mdv_t;
  auto out_map = [&](CUtensorMap* m, void* base, int S, long long ld, int box_rows) {
    return make_tmap_bf16_3d_sw(m, base, (uint64_t)p.H * 64, (uint64_t)S, (uint64_t)mapB, (uint64_t)ld * 2,
                                (uint64_t)S * ld * 2, 32, (uint32_t)box_rows, 1, 64);
  };
  if ((rc = out_map(&mdq, dq, mapSq, lddq, 32))) return rc;
  if ((rc = out_map(&mdk, dk, mapSk, lddk, 32))) return rc;
  if ((rc = out_map(&mdv, dv, mapSk, lddv, 32))) return rc;
  const int tail_rows = (packed && p.GS % 32) ? p.GS % 32 : 32;
  if ((rc = out_map(&mdq_t, dq, mapSq, lddq, tail_rows))) return rc;
  if ((rc = out_map(&mdk_t, dk, mapSk, lddk, tail_rows))) return rc;
  if ((rc = out_map(&mdv_t, dv, mapSk, lddv, tail_rows))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM_BYTES);
    if (err == cudaSuccess)
      err = cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM_BYTES);
    if (err != cudaSuccess) { set_error("attention_bwd: smem attribute: %s", cudaGetErrorString(err)); return VITSSL_ERR_CUDA; }
    configured = true;
  }
  // persistent: one CTA per SM walks the (batch, head) items — packed: (group of G images, head)
  const long long items = (packed ? (B + p.G - 1) / p.G : B) * H;
  const unsigned grid = (unsigned)(items < num_sms() ? items : num_sms());
  cudaError_t lerr = packed
      ? launch_pdl(attn_bwd_kernel<true>, dim3(grid), dim3(BWD_THREADS), BWD_SMEM_BYTES, stream, mq, mk, mv, mdo, mdq, mdk,
                   mdv, mdq_t, mdk_t, mdv_t, p)
      : launch_pdl(attn_bwd_kernel<false>, dim3(grid), dim3(BWD_THREADS), BWD_SMEM_BYTES, stream, mq, mk, mv, mdo, mdq, mdk,
                   mdv, mdq_t, mdk_t, mdv_t, p);
  if (lerr != cudaSuccess) { set_error("attention_bwd: launch failed: %s", cudaGetErrorString(lerr)); return VITSSL_ERR_CUDA; }
  return check_launch("attention_bwd");
}

extern "C" int vitssl_attention_generic_fwd(const void* q, const void* k, const void* v,
                                            const int64_t* host_strides, void* out, float* probs,
                                            float* lse, int64_t B, int64_t H, int64_t Sq, int64_t Sk,
                                            int64_t d, float scale, cudaStream_t stream) {
  VITSSL_REQUIRE(q && k && v && out && host_strides, VITSSL_ERR_ARG, "attention_generic_fwd: null pointer");
  VITSSL_REQUIRE(B > 0 && H > 0 && Sq > 0 && Sk > 0 && d > 0, VITSSL_ERR_SHAPE, "attention_generic_fwd: empty shape");
  const size_t smem = (size_t)(Sk + d + 8) * sizeof(float);
  VITSSL_REQUIRE(smem <= 200 * 1024, VITSSL_ERR_SHAPE, "attention_generic_fwd: Sk=%lld too long", (long long)Sk);
  AttnGenericParams p{};
  p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v;
  p.q_sb = host_strides[0]; p.q_sh = host_strides[1]; p.q_ss = host_strides[2];
  p.k_sb = host_strides[3]; p.k_sh = host_strides[4]; p.k_ss = host_strides[5];
  p.v_sb = host_strides[6]; p.v_sh = host_strides[7]; p.v_ss = host_strides[8];
  p.out = (__nv_bfloat16*)out; p.o_sb = host_strides[9]; p.o_sh = host_strides[10]; p.o_ss = host_strides[11];
  p.probs = probs; p.lse = lse; p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk; p.d = (int)d;
  p.scale = scale;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(attn_generic_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((unsigned)Sq, (unsigned)H, (unsigned)B);
  attn_generic_fwd_kernel<<<grid, 128, smem, stream>>>(p);
  return check_launch("attention_generic_fwd");
}

extern "C" int vitssl_attention_generic_bwd(const void* q, const void* k, const void* v,
                                            const int64_t* host_strides, const void* out,
                                            const void* d_out, const float* lse, void* dq,
                                            float* dk_acc, float* dv_acc, int64_t B, int64_t H,
                                            int64_t Sq, int64_t Sk, int64_t d, float scale,
                                            cudaStream_t stream) {
  VITSSL_REQUIRE(q && k && v && out && d_out && lse && dq && dk_acc && dv_acc && host_strides,
                 VITSSL_ERR_ARG, "attention_generic_bwd: null pointer");
  const size_t smem = (size_t)(2 * Sk + 2 * d + 4) * sizeof(float);
  VITSSL_REQUIRE(smem <= 200 * 1024, VITSSL_ERR_SHAPE, "attention_generic_bwd: Sk=%lld too long", (long long)Sk);
  AttnGenericBwdParams a{};
  AttnGenericParams& p = a.f;
  p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v;
  p.q_sb = host_strides[0]; p.q_sh = host_strides[1]; p.q_ss = host_strides[2];
  p.k_sb = host_strides[3]; p.k_sh = host_strides[4]; p.k_ss = host_strides[5];
  p.v_sb = host_strides[6]; p.v_sh = host_strides[7]; p.v_ss = host_strides[8];
  p.out = (__nv_bfloat16*)const_cast<void*>(out);
  p.o_sb = host_strides[9]; p.o_sh = host_strides[10]; p.o_ss = host_strides[11];
  p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk; p.d = (int)d; p.scale = scale;
  a.d_o = (const __nv_bfloat16*)d_out; a.lse = lse; a.dq = (__nv_bfloat16*)dq;
  a.dk_acc = dk_acc; a.dv_acc = dv_acc;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(attn_generic_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  dim3 grid((unsigned)Sq, (unsigned)H, (unsigned)B);
  attn_generic_bwd_kernel<<<grid, 128, smem, stream>>>(a);
  return check_launch("attention_generic_bwd");
}
