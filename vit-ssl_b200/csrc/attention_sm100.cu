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

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
struct AttnFwdParams {
  __nv_bfloat16* out; long long ldo;  // context [B*Sq, ldo]; head h occupies cols [64h, 64h+64)
  float* lse;                         // [B, H, Sq] log-sum-exp of the scaled scores
  int B, H, Sq, Sk, kv_rows;          // kv_rows = round_up(Sk, 16) <= 256
  float scale;
};

constexpr int FWD_THREADS = 192;  // warps 0-3 softmax (TMEM lane quadrant = warp), 4 = TMA, 5 = MMA
constexpr int FWD_SMEM_Q = 0;               // 16 KB: 128 query rows x 128 B
constexpr int FWD_SMEM_K = 16384;           // 32 KB: up to 256 key rows
constexpr int FWD_SMEM_V = 16384 + 32768;   // 32 KB
constexpr int FWD_SMEM_BAR = 16384 + 65536;
constexpr int FWD_SMEM_BYTES = FWD_SMEM_BAR + 128 + 1024;
constexpr uint32_t FWD_COL_O = 128;  // O accumulator columns [128,192): inside S, past the packed P

// Persistent: each CTA (two per SM) loops over (batch, head, 128-query tile) items. Per item
//   TMA warp : Q,K as soon as the previous item's S MMA retired; V once its PV MMA retired
//   MMA warp : S = Q K^T (SS) -> [softmax] -> O = P V with P read from TMEM (TS form)
//   softmax  : row max, p = 2^((s - max) * scale * log2e) with packed FMAs + MUFU.EX2, P written
//              back over S in TMEM as packed bf16 (tcgen05.st), then O / rowsum -> global.
// Nothing of the S x S probability matrix ever reaches shared or global memory
// (attention.py:20-23 materialises it three times).
__global__ void __launch_bounds__(FWD_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnFwdParams p) {
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = (p.Sq + 127) / 128;
  const int items = p.B * p.H * nq;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
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
  const uint32_t kv_bytes = static_cast<uint32_t>(p.kv_rows) * 128u;

  if (warp == 4) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int qt = item % nq, h = (item / nq) % p.H, b = item / (nq * p.H);
        if (it > 0) mbar_wait(bar_s, (it - 1) & 1);
        mbar_expect_tx(bar_qk, 16384u + kv_bytes);
        tma_load_3d(sQ, &tmap_q, bar_qk, h * 64, qt * 128, b);
        tma_load_3d(sK, &tmap_k, bar_qk, h * 64, 0, b);
        if (it > 0) mbar_wait(bar_o, (it - 1) & 1);
        mbar_expect_tx(bar_v, kv_bytes);
        tma_load_3d(sV, &tmap_v, bar_v, h * 64, 0, b);
      }
    }
  } else if (warp == 5) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(128, p.kv_rows, false, false);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, false, true);
      const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK), av = smem_u32(sV);
      const int ksteps = p.kv_rows / 16;
      int it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        mbar_wait(bar_qk, it & 1);
        if (it > 0) mbar_wait(bar_oread, (it - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tmem, umma_desc_sw128(aq + k * 32, 16, 1024),
                       umma_desc_sw128(ak + k * 32, 16, 1024), idesc_s, k > 0);
        umma_commit(bar_s);
        mbar_wait(bar_p, it & 1);
        mbar_wait(bar_v, it & 1);
        tc_fence_after();
        for (int ks = 0; ks < ksteps; ++ks)  // P[128, 16 keys] = 8 packed columns per step
          umma_bf16_ts(tmem + FWD_COL_O, tmem + ks * 8, umma_desc_sw128(av + ks * 2048, 8192, 1024),
                       idesc_o, ks > 0);
        umma_commit(bar_o);
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
    int it = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
      const int qt = item % nq, h = (item / nq) % p.H, b = item / (nq * p.H);
      mbar_wait(bar_s, it & 1);
      tc_fence_after();
      // pass 1: row maximum (TMEM loads run one chunk ahead; only the chunk straddling Sk is masked)
      float mx = -INFINITY;
      {
        uint32_t ra[32], rb[32];
        __syncwarp();
        tmem_ld_32x32(lane_addr, ra);
        auto pass1 = [&](uint32_t (&cur)[32], uint32_t (&nxt)[32], int c) {
          tmem_ld_wait();
          if (c + 1 < nchunks) tmem_ld_32x32(lane_addr + (c + 1) * 32, nxt);
          if (c * 32 + 32 <= p.Sk) {
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
              if (c * 32 + i < p.Sk) mx = fmaxf(mx, __uint_as_float(cur[i]));
          }
        };
        for (int c = 0; c < nchunks; c += 2) {
          pass1(ra, rb, c);
          if (c + 1 < nchunks) pass1(rb, ra, c + 1);
        }
      }
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
          const bool full = c * 32 + 32 <= p.Sk;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float a0, a1;
            upk2(ffma2(pk2(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])), sl2v, mnegv), a0, a1);
            float e0 = ex2_approx(a0), e1 = ex2_approx(a1);
            if (!full) {
              e0 = (c * 32 + i < p.Sk) ? e0 : 0.f;
              e1 = (c * 32 + i + 1 < p.Sk) ? e1 : 0.f;
            }
            sum2 = fadd2(sum2, pk2(e0, e1));  // fp32 row sum: exact LSE for the backward recomputation
            pk[i >> 1] = pack_bf16(e0, e1);
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
      float sum;
      {
        float s0, s1;
        upk2(sum2, s0, s1);
        sum = s0 + s1;
      }

      mbar_wait(bar_o, it & 1);
      tc_fence_after();
      const int qrow = qt * 128 + t;
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
        if (qrow < p.Sq) {
          __nv_bfloat16* op = p.out + (static_cast<long long>(b) * p.Sq + qrow) * p.ldo + h * 64;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(r0[i]) * inv, __uint_as_float(r0[i + 1]) * inv);
            o.y = pack_bf16(__uint_as_float(r0[i + 2]) * inv, __uint_as_float(r0[i + 3]) * inv);
            o.z = pack_bf16(__uint_as_float(r0[i + 4]) * inv, __uint_as_float(r0[i + 5]) * inv);
            o.w = pack_bf16(__uint_as_float(r0[i + 6]) * inv, __uint_as_float(r0[i + 7]) * inv);
            *reinterpret_cast<uint4*>(op + i) = o;
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(r1[i]) * inv, __uint_as_float(r1[i + 1]) * inv);
            o.y = pack_bf16(__uint_as_float(r1[i + 2]) * inv, __uint_as_float(r1[i + 3]) * inv);
            o.z = pack_bf16(__uint_as_float(r1[i + 4]) * inv, __uint_as_float(r1[i + 5]) * inv);
            o.w = pack_bf16(__uint_as_float(r1[i + 6]) * inv, __uint_as_float(r1[i + 7]) * inv);
            *reinterpret_cast<uint4*>(op + 32 + i) = o;
          }
          if (p.lse)
            p.lse[(static_cast<long long>(b) * p.H + h) * p.Sq + qrow] = mx * p.scale + __logf(sum);
        }
      }
    }
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
  const __nv_bfloat16* o; const __nv_bfloat16* d_o; long long ldo;  // [B*Sq, ldo]
  const float* lse;                                                 // [B,H,Sq]
  __nv_bfloat16* dq; long long lddq;                                // [B*Sq, lddq]
  __nv_bfloat16* dk; long long lddk;                                // [B*Sk, lddk]
  __nv_bfloat16* dv; long long lddv;
  int B, H, Sq, Sk;
  float scale;
};

constexpr int BWD_MATH_WARPS = 16;  // 4 per TMEM lane quadrant, 32 of the 128 key columns each
constexpr int BWD_THREADS = 32 * BWD_MATH_WARPS + 32;
constexpr int BWD_SMEM_Q = 0;          // 2 x 16 KB
constexpr int BWD_SMEM_DO = 32768;     // 2 x 16 KB
constexpr int BWD_SMEM_K = 65536;      // 2 x 16 KB
constexpr int BWD_SMEM_VV = 98304;     // 2 x 16 KB
constexpr int BWD_SMEM_P = 131072;     // 2 key blocks x 16 KB
constexpr int BWD_SMEM_DS = 163840;    // 2 key blocks x 16 KB
constexpr int BWD_SMEM_BAR = 196608;
constexpr int BWD_SMEM_BYTES = BWD_SMEM_BAR + 128 + 1024;

__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                const __grid_constant__ CUtensorMap tmap_o, const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + BWD_SMEM_BAR);
  uint64_t* bar_load = bar + 0;
  uint64_t* bar_sdp_full = bar + 1;
  uint64_t* bar_pds_ready = bar + 2;
  uint64_t* bar_pds_free = bar + 3;
  uint64_t* bar_dkv_free = bar + 4;
  uint64_t* bar_sdp_read = bar + 5;  // S / dP of the current iteration are in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int nq = (p.Sq + 127) / 128, nk = (p.Sk + 127) / 128;  // each <= 2
  const int nit = nq * nk;

  if (warp == BWD_MATH_WARPS) {
    if (lane == 0) {
      tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k);
      tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_do); tma_prefetch_desc(&tmap_o);
      mbar_init(bar_load, 1); mbar_init(bar_sdp_full, 1); mbar_init(bar_pds_ready, BWD_MATH_WARPS);
      mbar_init(bar_pds_free, 1); mbar_init(bar_dkv_free, BWD_MATH_WARPS);
      mbar_init(bar_sdp_read, BWD_MATH_WARPS);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 320, COL_DQ = 384;

  if (warp == BWD_MATH_WARPS) {
    if (lane == 0) {
      mbar_expect_tx(bar_load, static_cast<uint32_t>(3 * nq + 2 * nk) * 16384u);
      for (int i = 0; i < nq; ++i) {
        tma_load_3d(smem + BWD_SMEM_Q + i * 16384, &tmap_q, bar_load, h * 64, i * 128, b);
        tma_load_3d(smem + BWD_SMEM_DO + i * 16384, &tmap_do, bar_load, h * 64, i * 128, b);
        // O only feeds delta = rowsum(O * dO); it borrows the P region, which is idle until then
        tma_load_3d(smem + BWD_SMEM_P + i * 16384, &tmap_o, bar_load, h * 64, i * 128, b);
      }
      for (int j = 0; j < nk; ++j) {
        tma_load_3d(smem + BWD_SMEM_K + j * 16384, &tmap_k, bar_load, h * 64, j * 128, b);
        tma_load_3d(smem + BWD_SMEM_VV + j * 16384, &tmap_v, bar_load, h * 64, j * 128, b);
      }
      constexpr uint32_t idesc_sdp = umma_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_dkv = umma_idesc_bf16(128, 64, true, true);
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, 64, false, true);
      const uint32_t sq = smem_u32(smem + BWD_SMEM_Q), sdo = smem_u32(smem + BWD_SMEM_DO);
      const uint32_t sk = smem_u32(smem + BWD_SMEM_K), sv = smem_u32(smem + BWD_SMEM_VV);
      const uint32_t sp = smem_u32(smem + BWD_SMEM_P), sds = smem_u32(smem + BWD_SMEM_DS);

      auto issue_sdp = [&](int it) {
        const int j = it / nq, i = it % nq;
        const uint32_t qi = sq + i * 16384, doi = sdo + i * 16384;
        const uint32_t kj = sk + j * 16384, vj = sv + j * 16384;
#pragma unroll
        for (int k = 0; k < 4; ++k)  // S = Q_i K_j^T
          umma_bf16_ss(tmem + COL_S, umma_desc_sw128(qi + k * 32, 16, 1024),
                       umma_desc_sw128(kj + k * 32, 16, 1024), idesc_sdp, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // dP = dO_i V_j^T
          umma_bf16_ss(tmem + COL_DP, umma_desc_sw128(doi + k * 32, 16, 1024),
                       umma_desc_sw128(vj + k * 32, 16, 1024), idesc_sdp, k > 0);
        umma_commit(bar_sdp_full);
      };

      mbar_wait(bar_load, 0);
      tc_fence_after();
      issue_sdp(0);
      for (int it = 0; it < nit; ++it) {
        const int j = it / nq, i = it % nq;
        mbar_wait(bar_pds_ready, it & 1);
        tc_fence_after();
        if (it + 1 < nit) issue_sdp(it + 1);
        if (i == 0 && j > 0) {
          mbar_wait(bar_dkv_free, (j - 1) & 1);
          tc_fence_after();
        }
        const uint32_t qi = sq + i * 16384, doi = sdo + i * 16384, kj = sk + j * 16384;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)  // dV_j += P^T dO_i   (reduction over the 128 query rows)
          umma_bf16_ss(tmem + COL_DV, umma_desc_sw128(sp + ks * 2048, 16384, 1024),
                       umma_desc_sw128(doi + ks * 2048, 8192, 1024), idesc_dkv, (i > 0 || ks > 0));
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)  // dK_j += dS^T Q_i
          umma_bf16_ss(tmem + COL_DK, umma_desc_sw128(sds + ks * 2048, 16384, 1024),
                       umma_desc_sw128(qi + ks * 2048, 8192, 1024), idesc_dkv, (i > 0 || ks > 0));
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)  // dQ_i += dS K_j     (reduction over the 128 keys)
          umma_bf16_ss(tmem + COL_DQ + i * 64,
                       umma_desc_sw128(sds + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                       umma_desc_sw128(kj + ks * 2048, 8192, 1024), idesc_dq, (j > 0 || ks > 0));
        umma_commit(bar_pds_free);
      }
    }
  } else {
    // 512 math threads: row r = TMEM lane, `cq` selects 32 of the 128 key columns of the tile
    const int t = threadIdx.x;
    const int r = t & 127, cq = t >> 7;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const float sl2 = p.scale * kLog2e;
    float delta0 = 0.f, delta1 = 0.f, nlse0 = 0.f, nlse1 = 0.f;
    bool qvalid0 = false, qvalid1 = false;
    mbar_wait(bar_load, 0);
    // delta_i = sum_d O[i, d] * dO[i, d] from the TMA-staged (128B-swizzled) tiles in shared memory
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int qrow = i * 128 + r;
      const bool valid = i < nq && qrow < p.Sq;
      if (i == 0) qvalid0 = valid; else qvalid1 = valid;
      if (valid) {
        const uint8_t* so = smem + BWD_SMEM_P + i * 16384 + r * 128;
        const uint8_t* sd = smem + BWD_SMEM_DO + i * 16384 + r * 128;
        f32x2 acc2 = pk2(0.f, 0.f);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const int sw = (v ^ (r & 7)) << 4;
          const uint4 a = *reinterpret_cast<const uint4*>(so + sw);
          const uint4 d = *reinterpret_cast<const uint4*>(sd + sw);
          acc2 = ffma2(pk2(bf16_lo(a.x), bf16_hi(a.x)), pk2(bf16_lo(d.x), bf16_hi(d.x)), acc2);
          acc2 = ffma2(pk2(bf16_lo(a.y), bf16_hi(a.y)), pk2(bf16_lo(d.y), bf16_hi(d.y)), acc2);
          acc2 = ffma2(pk2(bf16_lo(a.z), bf16_hi(a.z)), pk2(bf16_lo(d.z), bf16_hi(d.z)), acc2);
          acc2 = ffma2(pk2(bf16_lo(a.w), bf16_hi(a.w)), pk2(bf16_lo(d.w), bf16_hi(d.w)), acc2);
        }
        float a0, a1;
        upk2(acc2, a0, a1);
        const float nl = -p.lse[(static_cast<long long>(b) * p.H + h) * p.Sq + qrow] * kLog2e;
        if (i == 0) { delta0 = a0 + a1; nlse0 = nl; } else { delta1 = a0 + a1; nlse1 = nl; }
      }
    }
    // every math thread has read O before anyone overwrites the region with P
    asm volatile("bar.sync 1, %0;" ::"n"(32 * BWD_MATH_WARPS) : "memory");
    // this thread's 4 x 16-byte chunks (32 keys) inside key block (cq >> 1) of the P / dS tiles
    uint8_t* sP = smem + BWD_SMEM_P + (cq >> 1) * 16384 + r * 128;
    uint8_t* sDS = smem + BWD_SMEM_DS + (cq >> 1) * 16384 + r * 128;

    for (int it = 0; it < nit; ++it) {
      const int j = it / nq, i = it % nq;
      mbar_wait(bar_sdp_full, it & 1);
      tc_fence_after();
      // invalid query rows recompute p = 2^(-inf) = 0 (their S rows are zero: TMA zero-fills Q)
      const bool rv = i ? qvalid1 : qvalid0;
      const float dl = i ? delta1 : delta0, nl = rv ? (i ? nlse1 : nlse0) : -INFINITY;
      const f32x2 sl2v = pk2(sl2, sl2), nlv = pk2(nl, nl), ndlv = pk2(-dl, -dl);
      uint32_t pp[2][8], dd[2][8];
#pragma unroll
      for (int c = 0; c < 2; ++c) {  // two sub-chunks of 16 keys
        const int key0 = j * 128 + cq * 32 + c * 16;
        if (key0 < p.Sk) {
          uint32_t sv[16], dp[16];
          __syncwarp();
          tmem_ld_32x16(lane_addr + COL_S + cq * 32 + c * 16, sv);
          tmem_ld_32x16(lane_addr + COL_DP + cq * 32 + c * 16, dp);
          tmem_ld_wait();
          const bool full = key0 + 16 <= p.Sk;
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            float a0, a1;
            upk2(ffma2(pk2(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])), sl2v, nlv), a0, a1);
            float p0 = ex2_approx(a0), p1 = ex2_approx(a1);
            if (!full) {
              p0 = (key0 + e < p.Sk) ? p0 : 0.f;
              p1 = (key0 + e + 1 < p.Sk) ? p1 : 0.f;
            }
            pp[c][e >> 1] = pack_bf16(p0, p1);
            float d0, d1;  // dS = P * (dP - delta)
            upk2(fmul2(pk2(p0, p1), fadd2(pk2(__uint_as_float(dp[e]), __uint_as_float(dp[e + 1])), ndlv)), d0, d1);
            dd[c][e >> 1] = pack_bf16(d0, d1);
          }
        } else {  // key chunk entirely past Sk
#pragma unroll
          for (int e = 0; e < 8; ++e) pp[c][e] = dd[c][e] = 0u;
        }
      }
      // everything above overlapped the previous iteration's dV / dK / dQ MMAs; only the stores
      // need their P / dS operands to have been consumed
      if (it > 0) mbar_wait(bar_pds_free, (it - 1) & 1);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int chunk = (cq & 1) * 4 + c * 2 + g;
          const int sw = (chunk ^ (r & 7)) << 4;
          *reinterpret_cast<uint4*>(sP + sw) = make_uint4(pp[c][4 * g], pp[c][4 * g + 1], pp[c][4 * g + 2], pp[c][4 * g + 3]);
          *reinterpret_cast<uint4*>(sDS + sw) = make_uint4(dd[c][4 * g], dd[c][4 * g + 1], dd[c][4 * g + 2], dd[c][4 * g + 3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_pds_ready);  // one arrival per warp

      if (i == nq - 1) {
        // key tile j finished: warps 0-7 drain dV_j, warps 8-15 dK_j; 32 of the 64 columns each
        mbar_wait(bar_pds_free, it & 1);
        tc_fence_after();
        const int krow = j * 128 + r;
        const int sel = cq >> 1, cc = cq & 1;
        const uint32_t col = (sel == 0 ? COL_DV : COL_DK) + cc * 32;
        const float mul = sel == 0 ? 1.0f : p.scale;
        __nv_bfloat16* base = sel == 0 ? p.dv : p.dk;
        const long long ld = sel == 0 ? p.lddv : p.lddk;
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(lane_addr + col, v);
        tmem_ld_wait();
        if (krow < p.Sk) {
          __nv_bfloat16* op = base + (static_cast<long long>(b) * p.Sk + krow) * ld + h * 64 + cc * 32;
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(v[e]) * mul, __uint_as_float(v[e + 1]) * mul);
            o.y = pack_bf16(__uint_as_float(v[e + 2]) * mul, __uint_as_float(v[e + 3]) * mul);
            o.z = pack_bf16(__uint_as_float(v[e + 4]) * mul, __uint_as_float(v[e + 5]) * mul);
            o.w = pack_bf16(__uint_as_float(v[e + 6]) * mul, __uint_as_float(v[e + 7]) * mul);
            *reinterpret_cast<uint4*>(op + e) = o;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dkv_free);
      }
    }
    // dQ: thread group cq drains columns (cq & 1) * 32 .. +32 of query tile cq >> 1 (all MMAs have
    // retired: the last bar_pds_free phase was waited on above)
    {
      const int tile = cq >> 1, cc = cq & 1;
      if (tile < nq) {
        const int qrow = tile * 128 + r;
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(lane_addr + COL_DQ + tile * 64 + cc * 32, v);
        tmem_ld_wait();
        if (qrow < p.Sq) {
          __nv_bfloat16* op = p.dq + (static_cast<long long>(b) * p.Sq + qrow) * p.lddq + h * 64 + cc * 32;
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(v[e]) * p.scale, __uint_as_float(v[e + 1]) * p.scale);
            o.y = pack_bf16(__uint_as_float(v[e + 2]) * p.scale, __uint_as_float(v[e + 3]) * p.scale);
            o.z = pack_bf16(__uint_as_float(v[e + 4]) * p.scale, __uint_as_float(v[e + 5]) * p.scale);
            o.w = pack_bf16(__uint_as_float(v[e + 6]) * p.scale, __uint_as_float(v[e + 7]) * p.scale);
            *reinterpret_cast<uint4*>(op + e) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == BWD_MATH_WARPS) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
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

extern "C" int vitssl_attention_supported(int64_t Sq, int64_t Sk, int64_t d_head) {
  return (d_head == 64 && Sk >= 1 && Sk <= 256 && Sq >= 1) ? 1 : 0;
}

extern "C" int vitssl_attention_fwd(const void* q, const void* k, const void* v, int64_t ldq,
                                    int64_t ldk, int64_t ldv, void* out, int64_t ldo, float* lse,
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
  p.out = reinterpret_cast<__nv_bfloat16*>(out); p.ldo = ldo; p.lse = lse;
  p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk;
  p.kv_rows = (int)((Sk + 15) / 16 * 16); p.scale = scale;
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_head_map(&mq, q, p.B, p.Sq, p.H, ldq, 128))) return rc;
  if ((rc = make_head_map(&mk, k, p.B, p.Sk, p.H, ldk, p.kv_rows))) return rc;
  if ((rc = make_head_map(&mv, v, p.B, p.Sk, p.H, ldv, p.kv_rows))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM_BYTES);
    if (err != cudaSuccess) { set_error("attention_fwd: smem attribute: %s", cudaGetErrorString(err)); return VITSSL_ERR_CUDA; }
    configured = true;
  }
  const long long items = (long long)B * H * ((Sq + 127) / 128);
  const long long slots = 2ll * num_sms();  // persistent: two CTAs per SM
  const unsigned grid = (unsigned)(items < slots ? items : slots);
  attn_fwd_kernel<<<grid, FWD_THREADS, FWD_SMEM_BYTES, stream>>>(mq, mk, mv, p);
  return check_launch("attention_fwd");
}

extern "C" int vitssl_attention_bwd(const void* q, const void* k, const void* v, int64_t ldq,
                                    int64_t ldk, int64_t ldv, const void* out, const void* d_out,
                                    int64_t ldo, const float* lse, void* dq, int64_t lddq, void* dk,
                                    int64_t lddk, void* dv, int64_t lddv, int64_t B, int64_t H,
                                    int64_t Sq, int64_t Sk, float scale, cudaStream_t stream) {
  VITSSL_REQUIRE(q && k && v && out && d_out && lse && dq && dk && dv, VITSSL_ERR_ARG, "attention_bwd: null pointer");
  VITSSL_REQUIRE(B > 0 && H > 0 && Sq > 0 && Sk > 0 && Sk <= 256 && Sq <= 256, VITSSL_ERR_SHAPE,
                 "attention_bwd: unsupported shape Sq=%lld Sk=%lld (tcgen05 path needs both <= 256)",
                 (long long)Sq, (long long)Sk);
  VITSSL_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddq % 8 == 0 &&
                 lddk % 8 == 0 && lddv % 8 == 0, VITSSL_ERR_SHAPE, "attention_bwd: row pitches must be multiples of 8");
  AttnBwdParams p{};
  p.o = reinterpret_cast<const __nv_bfloat16*>(out); p.d_o = reinterpret_cast<const __nv_bfloat16*>(d_out);
  p.ldo = ldo; p.lse = lse;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq); p.lddq = lddq;
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.lddk = lddk;
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv); p.lddv = lddv;
  p.B = (int)B; p.H = (int)H; p.Sq = (int)Sq; p.Sk = (int)Sk; p.scale = scale;
  CUtensorMap mq, mk, mv, mdo, mo;
  int rc;
  if ((rc = make_head_map(&mq, q, p.B, p.Sq, p.H, ldq, 128))) return rc;
  if ((rc = make_head_map(&mk, k, p.B, p.Sk, p.H, ldk, 128))) return rc;
  if ((rc = make_head_map(&mv, v, p.B, p.Sk, p.H, ldv, 128))) return rc;
  if ((rc = make_head_map(&mdo, d_out, p.B, p.Sq, p.H, ldo, 128))) return rc;
  if ((rc = make_head_map(&mo, out, p.B, p.Sq, p.H, ldo, 128))) return rc;
  static bool configured = false;
  if (!configured) {
    cudaError_t err = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM_BYTES);
    if (err != cudaSuccess) { set_error("attention_bwd: smem attribute: %s", cudaGetErrorString(err)); return VITSSL_ERR_CUDA; }
    configured = true;
  }
  dim3 grid((unsigned)H, (unsigned)B);
  attn_bwd_kernel<<<grid, BWD_THREADS, BWD_SMEM_BYTES, stream>>>(mq, mk, mv, mdo, mo, p);
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
