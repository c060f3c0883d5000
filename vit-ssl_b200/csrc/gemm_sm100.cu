// vitssl_b200 — bf16 GEMM for sm_100a: TMA -> shared-memory ring -> tcgen05.mma -> TMEM ->
// fused epilogue -> swizzled shared-memory staging -> TMA store. One persistent, warp-specialised
// kernel serves every linear layer on the hot path (reference call sites: attention.py:82-84,105;
// feed_forward.py:26-28; patch_embedding.py:22,79-84,113-116; ssl/simmim/model.py:45,57;
// ssl/dino/head.py:10-17):
//   forward   C[M,N]  = A[M,K] * W[N,K]^T      (both operands K-major)
//   dgrad     dX[M,K] = dY[M,N] * W[N,K]       (B operand MN-major)
//   wgrad     dW[N,K] = dY[M,N]^T * X[M,K]     (both operands MN-major, split-K + TMA reduce-add)
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = MMA issuer / TMEM owner,
// warps 2-17 = epilogue (four per TMEM lane quadrant; the tile's 32-column chunks are dealt
// round-robin, so every SM sub-partition has four warps to hide the epilogue's arithmetic
// latency). Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of
// tile i+1. Each epilogue warp drains its 32 rows one 32-column chunk at a time: tcgen05.ld ->
// bias / GELU / dGELU / dropout in registers -> 16-byte st.shared into a private 4 KB staging tile
// laid out in the TMA swizzle (bank-conflict free) -> one lane issues the bulk tensor store, so
// global writes are full 64/128-byte row segments instead of per-thread row fragments. Split-K
// partial sums leave the same way as TMA reduce-adds. The dGELU epilogue reads its saved
// pre-activations straight from global memory (64 contiguous bytes per thread). Tiles whose
// output cannot be a TMA target (row pitch not a multiple of 16 bytes) use the direct
// register->global epilogue.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "vitssl_b200.h"

namespace vitssl {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int EPI_WARPS = 16;  // four per TMEM lane quadrant: 32-column chunks are dealt round-robin
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int STG_BUF_BYTES = 4096;                   // one 32x32 fp32 chunk, or bf16 C + bf16 aux
constexpr int STG_BYTES = EPI_WARPS * STG_BUF_BYTES;  // one private staging tile per epilogue warp
constexpr int SMEM_LIMIT = 232448;                    // 227 KB per CTA

struct GemmShape {
  int M, N, K;
  int m_tiles, n_tiles, splits;
  int kblocks_total, kblocks_per_split;
};

struct GemmEpi {
  void* c;
  long long ldc;
  const float* bias;
  __nv_bfloat16* aux;
  long long ld_aux;
  float alpha;
  int mode;      // VITSSL_EPI_*
  int out_fp32;  // 1: fp32 output, 0: bf16
  int atomic;    // 1: red.add.f32 into C (split-K)
  int vec_ok;    // C rows are 16-byte aligned -> vector stores in the direct epilogue
  int tma_out;   // 1: epilogue goes through staging + TMA store (tmap_c / tmap_aux valid)
  uint32_t drop_th2;  // (p * 32768) * 0x10001, 0 = no dropout (common.cuh: dropout_lane_mask2)
  float drop_scale;
  PhiloxKeys7 keys;  // dropout generator state: site round keys (host-expanded) + the seed words
  const unsigned long long* seed_slot;  // non-null: the seed words come from this device word (graph replay)
  float* a_colsum;   // nullable: a_colsum[m] += alpha * sum_k op(A)[m, k] (bias gradient fused into a wgrad)
};
constexpr int ONES_OFFSET = 512;  // all-ones bf16 block (512 B) inside the 1 KB barrier page

// PAIR = true: two CTAs of a cluster share one 256 x BN tile (tcgen05 cta_group::2): each CTA
// stages its own 128 rows of A but only BN/2 columns of B, so the shared-memory fill and the
// tensor core's operand reads both drop by a third per FLOP — the single-CTA form is bound by
// shared-memory bandwidth (TMA writes + MMA operand reads) well below the tensor peak.
template <int BN, bool PAIR>
struct GemmCfg {
  static constexpr int kBRows = PAIR ? BN / 2 : BN;  // B columns staged by this CTA
  static constexpr int kBBytes = kBRows * BLOCK_K * 2;
  static constexpr int kStageBytes = A_TILE_BYTES + kBBytes;
  // the ring is deeper when the epilogue needs no staging tiles (split-K / direct epilogue)
  static constexpr int kAvail = SMEM_LIMIT - 1024 /*align*/ - 1024 /*barriers*/;
  static constexpr int kStagesStaged = (kAvail - STG_BYTES) / kStageBytes > 8 ? 8 : (kAvail - STG_BYTES) / kStageBytes;
  static constexpr int kStagesDirect = kAvail / kStageBytes > 8 ? 8 : kAvail / kStageBytes;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int smem_bytes(bool staged) {
    return 2048 + (staged ? kStagesStaged * kStageBytes + STG_BYTES : kStagesDirect * kStageBytes);
  }
};

// where this CTA sits in the persistent tile loop
struct TileCtx {
  int worker, nworkers;   // tile-loop start and stride (CTAs, or CTA pairs)
  int tile_rows;          // rows of one scheduled tile: 128, or 256 for a pair
  int row_off;            // this CTA's first row inside the tile (rank * 128)
  uint32_t tempty_addr;   // address of tempty_bar[0] in the CTA that issues the MMAs
  bool pair;              // tempty_addr is a shared::cluster address (the leader may be the peer)
};
// A cluster-scope release arrive costs a MEMBAR.ALL + ERRBAR per call (7 % of all stall samples of
// the GELU GEMM in ncu); a single CTA only needs the CTA-scope form.
__device__ __forceinline__ void tempty_arrive(const TileCtx& t, int acc) {
  if (t.pair) mbar_arrive_cluster(t.tempty_addr + acc * 8);
  else mbar_arrive_addr(t.tempty_addr + acc * 8);
}

// ---- epilogue math on one 32-column chunk of one accumulator row (all in registers) ----------
__device__ __forceinline__ void load_bias8(const float* bias, int nb, int g, int ncols, float (&b)[8]) {
  if (ncols == 32) {
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(bias + nb + 8 * g));
    const float4 t1 = __ldg(reinterpret_cast<const float4*>(bias + nb + 8 * g + 4));
    b[0] = t0.x; b[1] = t0.y; b[2] = t0.z; b[3] = t0.w; b[4] = t1.x; b[5] = t1.y; b[6] = t1.z; b[7] = t1.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (8 * g + i < ncols) ? __ldg(bias + nb + 8 * g + i) : 0.0f;
  }
}

// fp32 output (NONE / BIAS): v[] <- alpha * acc (+ bias)
template <int MODE>
__device__ __forceinline__ void epilogue_math_f32(uint32_t (&v)[32], const GemmEpi& e, int nb, int ncols) {
  const f32x2 al = pk2(e.alpha, e.alpha);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float b[8];
    if constexpr (MODE == VITSSL_EPI_BIAS) {
      load_bias8(e.bias, nb, g, ncols, b);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) b[i] = 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const int i = 8 * g + j;
      float x0, x1;
      upk2(ffma2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), al, pk2(b[j], b[j + 1])), x0, x1);
      v[i] = __float_as_uint(x0);
      v[i + 1] = __float_as_uint(x1);
    }
  }
}

// bf16 output: c[] <- packed bf16x2 of C. BIAS_GELU also produces u[] = packed bf16(alpha*acc+bias)
// (the aux output, feed_forward.py:26 pre-activation); DGELU consumes u[].
//   BIAS_GELU: C = dropout(u * Phi(u))          DGELU: C = alpha*acc * mask/(1-p) * (Phi(u) + u phi(u))
// Dropout: the 1/(1-p) factor rides on a multiply that is there anyway and the keep decision is a
// lane mask ANDed onto the packed bf16 pair (no per-element compare / select).
template <int MODE>
__device__ __forceinline__ void epilogue_math_bf16(const uint32_t (&v)[32], uint32_t (&c)[16], uint32_t (&u)[16],
                                                   const GemmEpi& e, const GemmShape& s, int row, int nb,
                                                   int ncols) {
  const f32x2 al = pk2(e.alpha, e.alpha);
  if constexpr (MODE == VITSSL_EPI_NONE || MODE == VITSSL_EPI_BIAS) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float b[8];
      if constexpr (MODE == VITSSL_EPI_BIAS) {
        load_bias8(e.bias, nb, g, ncols, b);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) b[i] = 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const int i = 8 * g + j;
        float x0, x1;
        upk2(ffma2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), al, pk2(b[j], b[j + 1])), x0, x1);
        c[i >> 1] = pack_bf16(x0, x1);
      }
    }
  } else if constexpr (MODE == VITSSL_EPI_MUL) {
    // C = alpha * acc * aux: backward of GELU(+dropout) with the factor mask/(1-p) * gelu'(u) saved by
    // the forward (BIAS_GELU_D) — no polynomial, no Philox, no exp in the backward epilogue
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float2 f = unpack_f16(u[i >> 1]);  // saved factor: fp16 pairs
      float d0, d1;
      upk2(fmul2(fmul2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), al), pk2(f.x, f.y)), d0, d1);
      c[i >> 1] = pack_bf16(d0, d1);
    }
  } else {
    const bool drop = e.drop_th2 != 0;
    const float sc = drop ? e.drop_scale : 1.0f;
    const f32x2 scv = MODE == VITSSL_EPI_DGELU ? pk2(sc * e.alpha, sc * e.alpha) : pk2(sc, sc);
    const unsigned long long g8 =
        (static_cast<unsigned long long>(row) * s.N + nb) >> 3;  // N % 8 == 0 enforced when dropping
    uint32_t sd2 = e.keys.c2, sd3 = e.keys.c3;
    if (drop && e.seed_slot != nullptr) {
      const unsigned long long sv = __ldg(e.seed_slot);
      sd2 = static_cast<uint32_t>(sv);
      sd3 = static_cast<uint32_t>(sv >> 32);
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint32_t keep[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
      if (drop) {
        const uint4 r = philox4x32_7_keyed(e.keys, sd2, sd3, g8 + g);
        keep[0] = dropout_lane_mask2(r.x, e.drop_th2);
        keep[1] = dropout_lane_mask2(r.y, e.drop_th2);
        keep[2] = dropout_lane_mask2(r.z, e.drop_th2);
        keep[3] = dropout_lane_mask2(r.w, e.drop_th2);
      }
      float b[8];
      if constexpr (MODE == VITSSL_EPI_BIAS_GELU || MODE == VITSSL_EPI_BIAS_GELU_D) load_bias8(e.bias, nb, g, ncols, b);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const int i = 8 * g + j;
        if constexpr (MODE == VITSSL_EPI_BIAS_GELU) {
          float x0, x1;
          upk2(ffma2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), al, pk2(b[j], b[j + 1])), x0, x1);
          const uint32_t up = pack_bf16(x0, x1);
          u[i >> 1] = up;
          const float u0 = bf16_lo(up), u1 = bf16_hi(up);
          float h0, h1;
          upk2(fmul2(fmul2(pk2(u0, u1), scv), normal_cdf2_sat(u0, u1)), h0, h1);
          c[i >> 1] = pack_bf16(h0, h1) & keep[j >> 1];
        } else if constexpr (MODE == VITSSL_EPI_BIAS_GELU_D) {
          // h = dropout(u Phi(u)) as above; the aux output is the whole backward factor
          // mask/(1-p) * (Phi(u) + u phi(u)) instead of u: the exp rides on the idle MUFU pipe here
          // and the backward GEMM's epilogue shrinks to one multiply (VITSSL_EPI_MUL)
          float x0, x1;
          upk2(ffma2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), al, pk2(b[j], b[j + 1])), x0, x1);
          const uint32_t up = pack_bf16(x0, x1);
          const float u0 = bf16_lo(up), u1 = bf16_hi(up);
          const f32x2 uu = pk2(u0, u1);
          const f32x2 cdf = normal_cdf2_sat(u0, u1);
          float a0, a1;
          upk2(fmul2(fmul2(uu, uu), pk2(-0.72134752044448170f, -0.72134752044448170f)), a0, a1);
          const f32x2 pdf = pk2(ex2_approx(a0), ex2_approx(a1));
          const f32x2 gp = ffma2(fmul2(uu, pk2(0.3989422804014327f, 0.3989422804014327f)), pdf, cdf);  // Phi + u phi
          float h0, h1, g0, g1;
          upk2(fmul2(fmul2(uu, scv), cdf), h0, h1);
          upk2(fmul2(gp, scv), g0, g1);
          c[i >> 1] = pack_bf16(h0, h1) & keep[j >> 1];
          u[i >> 1] = pack_f16(g0, g1) & keep[j >> 1];  // fp16: 11 significant bits for a factor in [-0.2, 1.2]/(1-p)
        } else {
          const uint32_t up = u[i >> 1];
          const float u0 = bf16_lo(up), u1 = bf16_hi(up);
          const f32x2 uu = pk2(u0, u1);
          // phi(u) = exp(-u^2/2) / sqrt(2 pi)
          float a0, a1;
          upk2(fmul2(fmul2(uu, uu), pk2(-0.72134752044448170f, -0.72134752044448170f)), a0, a1);
          const f32x2 pdf = pk2(ex2_approx(a0), ex2_approx(a1));
          const f32x2 upd = fmul2(fmul2(uu, pk2(0.3989422804014327f, 0.3989422804014327f)), pdf);
          const f32x2 gp = fadd2(upd, normal_cdf2_sat(u0, u1));  // Phi + u phi
          float d0, d1;
          upk2(fmul2(fmul2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), scv), gp), d0, d1);
          c[i >> 1] = pack_bf16(d0, d1) & keep[j >> 1];
        }
      }
    }
  }
}

// 32 bf16 (16 packed words) of row `lane` -> staging tile with rows of 64 bytes in the TMA
// SWIZZLE_64B pattern (16-byte chunk index ^= (row >> 1) & 3): a quarter-warp's 8 stores hit 8
// distinct bank groups.
__device__ __forceinline__ void stage_row_bf16(uint8_t* tile, int lane, const uint32_t (&c)[16]) {
  const uint32_t base = smem_u32(tile) + lane * 64;
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t a = base + ((j ^ sw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(c[4 * j]), "r"(c[4 * j + 1]),
                 "r"(c[4 * j + 2]), "r"(c[4 * j + 3])
                 : "memory");
  }
}
// 32 fp32 of row `lane` -> staging tile with rows of 128 bytes in the SWIZZLE_128B pattern
__device__ __forceinline__ void stage_row_f32(uint8_t* tile, int lane, const uint32_t (&v)[32]) {
  const uint32_t base = smem_u32(tile) + lane * 128;
  const int sw = lane & 7;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t a = base + ((j ^ sw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                 "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                 : "memory");
  }
}
// DGELU operand: 32 saved pre-activations (bf16) of one row straight from global memory; each
// thread reads one contiguous 64-byte segment (two full sectors), issued before the accumulator
// wait so the latency hides behind it. Rows past M / columns past N read as 0.
__device__ __forceinline__ void load_aux_row(const __nv_bfloat16* aux, long long ld, int row, int M, int nb,
                                             int ncols, uint32_t (&u)[16]) {
  if (row < M && ncols == 32) {
    const uint4* p = reinterpret_cast<const uint4*>(aux + static_cast<long long>(row) * ld + nb);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 t = __ldg(p + j);
      u[4 * j] = t.x; u[4 * j + 1] = t.y; u[4 * j + 2] = t.z; u[4 * j + 3] = t.w;
    }
  } else {
    const unsigned short* p = reinterpret_cast<const unsigned short*>(aux) + static_cast<long long>(row) * ld + nb;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint32_t lo = (row < M && 2 * i < ncols) ? __ldg(p + 2 * i) : 0u;
      const uint32_t hi = (row < M && 2 * i + 1 < ncols) ? __ldg(p + 2 * i + 1) : 0u;
      u[i] = lo | (hi << 16);
    }
  }
}

// ---- staged epilogue of one warp over its share of the persistent tile loop -------------------
// The warp owns TMEM lanes [32q, 32q+32) and the 32-column chunks cg, cg+4, ... of every tile.
// Four warps per scheduler give the chunk math (several hundred dependent ALU/FMA instructions
// for the GELU forms) the thread-level parallelism it needs; each warp stages one chunk at a
// time in a private 4 KB tile and hands it to a TMA store.
template <int BN, int MODE, bool OUT_F32>
__device__ __forceinline__ void epilogue_staged(const CUtensorMap* tmap_c, const CUtensorMap* tmap_aux,
                                                const GemmShape& s, const GemmEpi& e, uint8_t* stg,
                                                uint64_t* tfull_bar, const TileCtx& tc,
                                                uint32_t tmem_base, int q, int cg, int lane) {
  static_assert(!(OUT_F32 && MODE >= VITSSL_EPI_BIAS_GELU), "GELU / MUL epilogues write bf16");
  constexpr bool kAuxIn = MODE == VITSSL_EPI_DGELU || MODE == VITSSL_EPI_MUL;             // aux is read
  constexpr bool kAuxOut = MODE == VITSSL_EPI_BIAS_GELU || MODE == VITSSL_EPI_BIAS_GELU_D;  // aux is written
  constexpr int NC = BN / 32;
  const int num_work = s.m_tiles * s.n_tiles * s.splits;
  int acc = 0;
  uint32_t acc_phase = 0;
  for (int w = tc.worker; w < num_work; w += tc.nworkers) {
    const int n_blk = w % s.n_tiles;
    const int m_blk = (w / s.n_tiles) % s.m_tiles;
    const int row0 = m_blk * tc.tile_rows + tc.row_off + q * 32;
    const int nbase = n_blk * BN;
    const int nch = min(NC, (s.N - nbase + 31) >> 5);  // chunks of this tile that hold columns < N
    const uint32_t tcol = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;

    uint32_t up[16];
    if constexpr (kAuxIn) {  // first chunk's saved operand: in flight while the MMAs finish
      if (cg < nch) load_aux_row(e.aux, e.ld_aux, row0 + lane, s.M, nbase + cg * 32, min(32, s.N - nbase - cg * 32), up);
    }
    mbar_wait(&tfull_bar[acc], acc_phase);
    tc_fence_after();
    __syncwarp();
    if constexpr (OUT_F32 && MODE == VITSSL_EPI_NONE) {
      // fused row sums of op(A) (the tile's extra 16-column accumulator, all columns equal):
      // one warp per lane quadrant adds its 32 rows' partial sums
      if (e.a_colsum != nullptr && n_blk == 0 && cg == 0) {
        const uint32_t v = tmem_ld_32x1(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 2 * BN + acc * 16);
        tmem_ld_wait();
        if (row0 + lane < s.M) atomicAdd(e.a_colsum + row0 + lane, __uint_as_float(v) * e.alpha);
        __syncwarp();
      }
    }
    if (cg >= nch) {
      tc_fence_before();
      if (lane == 0) tempty_arrive(tc, acc);
    }
#pragma unroll 1
    for (int c = cg; c < nch; c += 4) {
      const int nb = nbase + c * 32;
      const int ncols = min(32, s.N - nb);
      uint32_t r[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the single-lane sections
      tmem_ld_32x32(tcol + c * 32, r);
      if constexpr (kAuxIn) {
        if (c != cg) load_aux_row(e.aux, e.ld_aux, row0 + lane, s.M, nb, ncols, up);
      }
      tmem_ld_wait();
      if (c + 4 >= nch) {  // this warp's columns are all in registers: hand the TMEM buffer back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) tempty_arrive(tc, acc);
      }
      if constexpr (OUT_F32) {
        epilogue_math_f32<MODE>(r, e, nb, ncols);
        if (lane == 0) tma_store_wait_read<0>();  // the previous chunk's store has drained the tile
        __syncwarp();
        stage_row_f32(stg, lane, r);
      } else {
        uint32_t cp[16];
        epilogue_math_bf16<MODE>(r, cp, up, e, s, row0 + lane, nb, ncols);
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        stage_row_bf16(stg, lane, cp);
        if constexpr (kAuxOut) stage_row_bf16(stg + 2048, lane, up);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if (OUT_F32 && e.atomic) tma_reduce_add_2d(tmap_c, stg, nb, row0);  // split-K partial sum
        else tma_store_2d(tmap_c, stg, nb, row0);
        if constexpr (kAuxOut) tma_store_2d(tmap_aux, stg + 2048, nb, row0);
        tma_store_commit();
      }
    }
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1;
  }
  if (lane == 0) tma_store_wait_all();
}

// ---- direct epilogue: registers -> global, per-thread row fragments (split-K red.add, or an
// output whose pitch TMA cannot address). NONE / BIAS only. -------------------------------------
template <int BN>
__device__ __forceinline__ void epilogue_direct(const GemmShape& s, const GemmEpi& e,
                                                uint64_t* tfull_bar, const TileCtx& tc,
                                                uint32_t tmem_base, int q, int cg, int lane) {
  constexpr int NC = BN / 32;
  const int num_work = s.m_tiles * s.n_tiles * s.splits;
  int acc = 0;
  uint32_t acc_phase = 0;
  for (int w = tc.worker; w < num_work; w += tc.nworkers) {
    const int n_blk = w % s.n_tiles;
    const int m_blk = (w / s.n_tiles) % s.m_tiles;
    const int m0 = m_blk * tc.tile_rows + tc.row_off, n0 = n_blk * BN;
    const int row = m0 + q * 32 + lane;
    mbar_wait(&tfull_bar[acc], acc_phase);
    tc_fence_after();
    if (cg >= NC) {
      __syncwarp();
      tc_fence_before();
      if (lane == 0) tempty_arrive(tc, acc);
    }
#pragma unroll 1
    for (int c = cg; c < NC; c += 4) {
      uint32_t r[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the predicated stores
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN + c * 32, r);
      tmem_ld_wait();
      if (c + 4 >= NC) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) tempty_arrive(tc, acc);
      }
      const int nb = n0 + c * 32;
      if (row >= s.M || nb >= s.N) continue;
      const int ncols = min(32, s.N - nb);
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) * e.alpha;
      if (e.mode == VITSSL_EPI_BIAS) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < ncols) v[i] += __ldg(e.bias + nb + i);
      }
      if (e.atomic) {
        float* cp = reinterpret_cast<float*>(e.c) + static_cast<long long>(row) * e.ldc + nb;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < ncols) atomicAdd(cp + i, v[i]);
      } else if (e.out_fp32) {
        float* cp = reinterpret_cast<float*>(e.c) + static_cast<long long>(row) * e.ldc + nb;
        if (ncols == 32 && e.vec_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(cp + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < ncols) cp[i] = v[i];
        }
      } else {
        __nv_bfloat16* cp =
            reinterpret_cast<__nv_bfloat16*>(e.c) + static_cast<long long>(row) * e.ldc + nb;
        if (ncols == 32 && e.vec_ok) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 pk;
            pk.x = pack_bf16(v[i], v[i + 1]); pk.y = pack_bf16(v[i + 2], v[i + 3]);
            pk.z = pack_bf16(v[i + 4], v[i + 5]); pk.w = pack_bf16(v[i + 6], v[i + 7]);
            *reinterpret_cast<uint4*>(cp + i) = pk;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < ncols) cp[i] = __float2bfloat16_rn(v[i]);
        }
      }
    }
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1;
  }
}

template <int BN, bool A_MN, bool B_MN, bool PAIR>
__global__ void __launch_bounds__(GEMM_THREADS, 1)  // 18 warps are allocated as 20: 96 registers each
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c,
                    const __grid_constant__ CUtensorMap tmap_aux, const GemmShape s,
                    const GemmEpi e) {
  using Cfg = GemmCfg<BN, PAIR>;
  constexpr int STAGE_BYTES = Cfg::kStageBytes;
  constexpr int B_ROWS = Cfg::kBRows;
  constexpr uint32_t IDESC = umma_idesc_bf16(PAIR ? 256 : BLOCK_M, BN, A_MN, B_MN);
  const int STAGES = e.tma_out ? Cfg::kStagesStaged : Cfg::kStagesDirect;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem_al = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  // [barriers 1 KB][ring: STAGES x STAGE_BYTES][staging tiles] — everything 1024-aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_al);
  uint8_t* smem = smem_al + 1024;
  uint8_t* staging = smem + STAGES * STAGE_BYTES;
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tfull_bar = empty_bar + 8;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;  // rank 0 of a pair issues the MMAs
  TileCtx tc;
  tc.worker = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  tc.nworkers = PAIR ? (gridDim.x >> 1) : gridDim.x;
  tc.tile_rows = PAIR ? 256 : BLOCK_M;
  tc.row_off = PAIR ? static_cast<int>(rank) * BLOCK_M : 0;
  tc.tempty_addr = PAIR ? mapa_u32(smem_u32(tempty_bar), 0) : smem_u32(tempty_bar);
  tc.pair = PAIR;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (e.tma_out) {
      tma_prefetch_desc(&tmap_c);
      if (e.mode >= VITSSL_EPI_BIAS_GELU) tma_prefetch_desc(&tmap_aux);
    }
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PAIR ? 2 * EPI_WARPS : EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
    else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  if (warp == 2 && e.a_colsum != nullptr) {  // 512 bytes of bf16 1.0 for the row-sum MMA's B operand
    reinterpret_cast<uint4*>(smem_al + ONES_OFFSET)[lane] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // the peer's barriers exist before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // everything above overlapped the previous kernel's tail; global memory from here on

  const int num_work = s.m_tiles * s.n_tiles * s.splits;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full0 = PAIR ? mapa_u32(smem_u32(full_bar), 0) : 0u;  // the leader's full barriers
      for (int w = tc.worker; w < num_work; w += tc.nworkers) {
        const int n_blk = w % s.n_tiles;
        const int m_blk = (w / s.n_tiles) % s.m_tiles;
        const int sp = w / (s.n_tiles * s.m_tiles);
        const int kb0 = sp * s.kblocks_per_split;
        const int kb1 = min(kb0 + s.kblocks_per_split, s.kblocks_total);
        const int m0 = m_blk * tc.tile_rows + tc.row_off;
        const int n0 = n_blk * BN + (PAIR ? static_cast<int>(rank) * B_ROWS : 0);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_parked(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + stage * STAGE_BYTES;
          uint8_t* b_dst = a_dst + A_TILE_BYTES;
          const int k0 = kb * BLOCK_K;
          if constexpr (PAIR) {
            // both CTAs' bytes land on the leader's barrier; only the leader arms it
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
            const uint32_t fb = full0 + stage * 8;
            if constexpr (!A_MN) {
              tma_load_2d_pair(a_dst, &tmap_a, fb, k0, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j)
                tma_load_2d_pair(a_dst + j * 8192, &tmap_a, fb, m0 + 64 * j, k0);
            }
            if constexpr (!B_MN) {
              tma_load_2d_pair(b_dst, &tmap_b, fb, k0, n0);
            } else {
#pragma unroll
              for (int j = 0; j < B_ROWS / 64; ++j)
                tma_load_2d_pair(b_dst + j * 8192, &tmap_b, fb, n0 + 64 * j, k0);
            }
          } else {
            mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
            if constexpr (!A_MN) {
              tma_load_2d(a_dst, &tmap_a, &full_bar[stage], k0, m0);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j)
                tma_load_2d(a_dst + j * 8192, &tmap_a, &full_bar[stage], m0 + 64 * j, k0);
            }
            if constexpr (!B_MN) {
              tma_load_2d(b_dst, &tmap_b, &full_bar[stage], k0, n0);
            } else {
#pragma unroll
              for (int j = 0; j < B_ROWS / 64; ++j)
                tma_load_2d(b_dst + j * 8192, &tmap_b, &full_bar[stage], n0 + 64 * j, k0);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    // The whole warp walks the tile / k-block schedule and the barrier waits (uniform control
    // flow); one elected lane issues the MMAs and commits. Descriptor low words are derived from
    // the stage base with one add per k-step and every form shares the high word, so an MMA costs a
    // few uniform-datapath instructions. (Issued from `if (lane == 0)` the same code needed a
    // vector->uniform hand-off loop per operand: ~18 instructions per MMA against a 96-cycle MMA at
    // BN = 192 — the issuer, not the tensor pipe, paced the small-tile GEMMs.)
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      constexpr uint32_t HI = 0x40004040u;  // SBO 1024 | descriptor version 1 | SWIZZLE_128B
      constexpr uint32_t LBO_K = (16u >> 4) << 16, LBO_MN = (8192u >> 4) << 16;
      constexpr uint32_t A_STEP = A_MN ? (2048u >> 4) : (32u >> 4), B_STEP = B_MN ? (2048u >> 4) : (32u >> 4);
      const uint32_t ring_lo = smem_u32(smem) >> 4;
      // fused row sums: D1[128 x 16] += A_tile * ones, issued for the first column tile of every row tile
      constexpr uint32_t IDESC_ONES = umma_idesc_bf16(BLOCK_M, 16, A_MN, false);
      constexpr uint32_t HI_ONES = (256u >> 4) | (1u << 14);  // SBO 256, version 1, no swizzle
      const uint32_t ones_lo = (smem_u32(smem_al + ONES_OFFSET) >> 4) | ((128u >> 4) << 16);
      for (int w = tc.worker; w < num_work; w += tc.nworkers) {
        const int sp = w / (s.n_tiles * s.m_tiles);
        const int kb0 = sp * s.kblocks_per_split;
        const int kb1 = min(kb0 + s.kblocks_per_split, s.kblocks_total);
        const bool row_sums = !PAIR && e.a_colsum != nullptr && (w % s.n_tiles) == 0;
        mbar_wait_parked(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_parked(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = (ring_lo + stage * (STAGE_BYTES >> 4)) | (A_MN ? LBO_MN : LBO_K);
          const uint32_t b_lo = (ring_lo + stage * (STAGE_BYTES >> 4) + (A_TILE_BYTES >> 4)) | (B_MN ? LBO_MN : LBO_K);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              // K-major: 16 bf16 = 32 bytes along the swizzled row; 8-row groups 1024 B apart.
              // MN-major: 16 k-rows = 2048 bytes; 64-wide MN groups 8192 B apart (LBO),
              //           8-row k groups 1024 B apart (SBO).
              umma_bf16_ss_lo<PAIR>(d_tmem, a_lo + k * A_STEP, b_lo + k * B_STEP, HI, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            if (row_sums) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / 16; ++k)
                umma_bf16_ss_lohi(tmem_base + 2 * BN + acc * 16, a_lo + k * A_STEP, HI, ones_lo, HI_ONES, IDESC_ONES,
                                  (kb > kb0 || k > 0) ? 1u : 0u);
            }
            // frees the smem slot (in both CTAs of a pair) once these MMAs retire
            if constexpr (PAIR) umma_commit_pair(&empty_bar[stage]);
            else umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (elect_one()) {
          if constexpr (PAIR) umma_commit_pair(&tfull_bar[acc]);
          else umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------ epilogue ------------------------------
    const int q = warp & 3;          // TMEM lane quadrant this warp may access
    const int cg = (warp - 2) >> 2;  // its 32-column chunks: cg, cg + 4, ...
    uint8_t* stg = staging + (warp - 2) * STG_BUF_BYTES;
    if (!e.tma_out) {
      epilogue_direct<BN>(s, e, tfull_bar, tc, tmem_base, q, cg, lane);
    } else if (e.mode == VITSSL_EPI_BIAS_GELU) {
      epilogue_staged<BN, VITSSL_EPI_BIAS_GELU, false>(&tmap_c, &tmap_aux, s, e, stg, tfull_bar, tc, tmem_base, q,
                                                       cg, lane);
    } else if (e.mode == VITSSL_EPI_DGELU) {
      epilogue_staged<BN, VITSSL_EPI_DGELU, false>(&tmap_c, &tmap_aux, s, e, stg, tfull_bar, tc, tmem_base, q, cg,
                                                   lane);
    } else if (e.mode == VITSSL_EPI_BIAS_GELU_D) {
      epilogue_staged<BN, VITSSL_EPI_BIAS_GELU_D, false>(&tmap_c, &tmap_aux, s, e, stg, tfull_bar, tc, tmem_base, q,
                                                         cg, lane);
    } else if (e.mode == VITSSL_EPI_MUL) {
      epilogue_staged<BN, VITSSL_EPI_MUL, false>(&tmap_c, &tmap_aux, s, e, stg, tfull_bar, tc, tmem_base, q, cg,
                                                 lane);
    } else if (e.mode == VITSSL_EPI_BIAS) {
      if (e.out_fp32)
        epilogue_staged<BN, VITSSL_EPI_BIAS, true>(&tmap_c, &tmap_aux, s, e, stg, tfull_bar, tc, tmem_base, q, cg,
                                                   lane);
      else
        epilogue_staged<BN, VITSSL_EPI_BIAS, false>(&tmap_c, &tmap_aux, s, e, stg, tfull_bar, tc, tmem_base, q, cg,
                                                    lane);
    } else {
      if (e.out_fp32)
        epilogue_staged<BN, VITSSL_EPI_NONE, true>(&tmap_c, &tmap_aux, s, e, stg, tfull_bar, tc, tmem_base, q, cg,
                                                   lane);
      else
        epilogue_staged<BN, VITSSL_EPI_NONE, false>(&tmap_c, &tmap_aux, s, e, stg, tfull_bar, tc, tmem_base, q, cg,
                                                    lane);
    }
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // no CTA leaves while its peer may still signal it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ----------------------------------------------------------------------------------
// Generic SIMT GEMM for shapes TMA cannot address (row pitch not a multiple of 16 bytes, e.g.
// the 10-class MLP head, mlp_head.py:10). 32x32 tiles, fp32 accumulate. Tiny problems only.
// ----------------------------------------------------------------------------------
__global__ void gemm_simt_kernel(const __nv_bfloat16* __restrict__ A,
                                 const __nv_bfloat16* __restrict__ B, int M, int N, int K,
                                 long long lda, long long ldb, int a_mn, int b_mn,
                                 const GemmEpi e) {
  __shared__ float As[32][33];
  __shared__ float Bs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < K; k0 += 32) {
    for (int i = ty; i < 32; i += 8) {
      // As[i][tx] = A(m0+i, k0+tx); Bs[i][tx] = B(k0+tx, n0+i)
      float av = 0.f, bv = 0.f;
      if (a_mn) {  // A stored [K][M]: read with M fastest
        const int m = m0 + tx, k = k0 + i;
        if (m < M && k < K) av = __bfloat162float(A[static_cast<long long>(k) * lda + m]);
        As[tx][i] = av;
      } else {
        const int m = m0 + i, k = k0 + tx;
        if (m < M && k < K) av = __bfloat162float(A[static_cast<long long>(m) * lda + k]);
        As[i][tx] = av;
      }
      if (b_mn) {  // B stored [K][N]
        const int n = n0 + tx, k = k0 + i;
        if (n < N && k < K) bv = __bfloat162float(B[static_cast<long long>(k) * ldb + n]);
        Bs[tx][i] = bv;
      } else {  // B stored [N][K]
        const int n = n0 + i, k = k0 + tx;
        if (n < N && k < K) bv = __bfloat162float(B[static_cast<long long>(n) * ldb + k]);
        Bs[i][tx] = bv;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float b = Bs[tx][k];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = fmaf(As[ty + 8 * j][k], b, acc[j]);
    }
    __syncthreads();
  }
  const int n = n0 + tx;
  if (n >= N) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int m = m0 + ty + 8 * j;
    if (m >= M) continue;
    float v = acc[j] * e.alpha;
    if (e.mode == VITSSL_EPI_BIAS || e.mode == VITSSL_EPI_BIAS_GELU || e.mode == VITSSL_EPI_BIAS_GELU_D) v += e.bias[n];
    if (e.mode == VITSSL_EPI_BIAS_GELU) {
      v = bf16_round(v);
      e.aux[static_cast<long long>(m) * e.ld_aux + n] = __float2bfloat16_rn(v);
      v = gelu_erf(v);
    } else if (e.mode == VITSSL_EPI_BIAS_GELU_D) {  // (no dropout on this path); the factor is stored as fp16
      v = bf16_round(v);
      reinterpret_cast<__half*>(e.aux)[static_cast<long long>(m) * e.ld_aux + n] = __float2half_rn(gelu_erf_grad(v));
      v = gelu_erf(v);
    } else if (e.mode == VITSSL_EPI_DGELU) {
      v *= gelu_erf_grad(__bfloat162float(e.aux[static_cast<long long>(m) * e.ld_aux + n]));
    } else if (e.mode == VITSSL_EPI_MUL) {
      v *= __half2float(reinterpret_cast<const __half*>(e.aux)[static_cast<long long>(m) * e.ld_aux + n]);
    }
    if (e.out_fp32)
      reinterpret_cast<float*>(e.c)[static_cast<long long>(m) * e.ldc + n] = v;
    else
      reinterpret_cast<__nv_bfloat16*>(e.c)[static_cast<long long>(m) * e.ldc + n] =
          __float2bfloat16_rn(v);
  }
}

template <int BN, bool A_MN, bool B_MN, bool PAIR>
int launch_tcgen05(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                   const CUtensorMap& tx, const GemmShape& s, const GemmEpi& e, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, PAIR>;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, PAIR>;
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t err =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg::smem_bytes(false) > Cfg::smem_bytes(true) ? Cfg::smem_bytes(false) : Cfg::smem_bytes(true));
    if (err != cudaSuccess) {
      set_error("gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(err));
      return VITSSL_ERR_CUDA;
    }
    configured = true;
  }
  const int num_work = s.m_tiles * s.n_tiles * s.splits;
  if constexpr (PAIR) {
    const int pairs = num_sms() / 2;
    const int nclusters = num_work < pairs ? num_work : pairs;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * nclusters);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::smem_bytes(e.tma_out != 0);
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    cudaError_t err = cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, tx, s, e);
    if (err != cudaSuccess) {
      set_error("gemm: cluster launch failed: %s", cudaGetErrorString(err));
      return VITSSL_ERR_CUDA;
    }
    return check_launch("gemm_tcgen05_pair");
  } else {
    const int grid = num_work < num_sms() ? num_work : num_sms();
    cudaError_t err = launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::smem_bytes(e.tma_out != 0), stream, ta, tb,
                                 tc, tx, s, e);
    if (err != cudaSuccess) {
      set_error("gemm: launch failed: %s", cudaGetErrorString(err));
      return VITSSL_ERR_CUDA;
    }
    return check_launch("gemm_tcgen05");
  }
}

template <bool A_MN, bool B_MN>
int dispatch_bn(int bn, bool pair, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                const CUtensorMap& tx, const GemmShape& s, const GemmEpi& e, cudaStream_t stream) {
  if (pair) {
    switch (bn) {
      case 128: return launch_tcgen05<128, A_MN, B_MN, true>(ta, tb, tc, tx, s, e, stream);
      case 192: return launch_tcgen05<192, A_MN, B_MN, true>(ta, tb, tc, tx, s, e, stream);
      default: return launch_tcgen05<256, A_MN, B_MN, true>(ta, tb, tc, tx, s, e, stream);
    }
  }
  switch (bn) {
    case 64: return launch_tcgen05<64, A_MN, B_MN, false>(ta, tb, tc, tx, s, e, stream);
    case 128: return launch_tcgen05<128, A_MN, B_MN, false>(ta, tb, tc, tx, s, e, stream);
    case 192: return launch_tcgen05<192, A_MN, B_MN, false>(ta, tb, tc, tx, s, e, stream);
    default: return launch_tcgen05<256, A_MN, B_MN, false>(ta, tb, tc, tx, s, e, stream);
  }
}

int pick_block_n(int N) {
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  if (N <= 64) return 64;
  if (N <= 128) return 128;
  if (N <= 192) return 192;
  return 256;
}

}  // namespace

}  // namespace vitssl

using namespace vitssl;

static int gemm_impl(const void* A, const void* B, void* C, int64_t M, int64_t N,
                     int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_mn,
                     int b_mn, int epilogue, const float* bias, void* aux,
                     int64_t ld_aux, float alpha, int out_fp32, int split_k,
                     float dropout_p, uint64_t philox_seed, uint64_t philox_offset, float* a_colsum,
                     cudaStream_t stream) {
  VITSSL_REQUIRE(A && B && C, VITSSL_ERR_ARG, "gemm: null operand");
  VITSSL_REQUIRE(M > 0 && N > 0 && K > 0, VITSSL_ERR_SHAPE, "gemm: empty problem %lld x %lld x %lld",
                 (long long)M, (long long)N, (long long)K);
  VITSSL_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), VITSSL_ERR_SHAPE,
                 "gemm: dimension exceeds int32");
  VITSSL_REQUIRE(epilogue >= VITSSL_EPI_NONE && epilogue <= VITSSL_EPI_MUL, VITSSL_ERR_ARG,
                 "gemm: bad epilogue %d", epilogue);
  if (epilogue == VITSSL_EPI_BIAS || epilogue == VITSSL_EPI_BIAS_GELU || epilogue == VITSSL_EPI_BIAS_GELU_D)
    VITSSL_REQUIRE(bias != nullptr, VITSSL_ERR_ARG, "gemm: epilogue needs bias");
  if (epilogue >= VITSSL_EPI_BIAS_GELU)
    VITSSL_REQUIRE(aux != nullptr && ld_aux >= N, VITSSL_ERR_ARG, "gemm: epilogue needs aux");
  VITSSL_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, VITSSL_ERR_ARG, "gemm: dropout_p out of range");
  if (dropout_p > 0.f)
    VITSSL_REQUIRE(N % 8 == 0, VITSSL_ERR_SHAPE, "gemm: dropout epilogue needs N %% 8 == 0");

  GemmEpi e{};
  e.c = C; e.ldc = ldc; e.bias = bias; e.aux = reinterpret_cast<__nv_bfloat16*>(aux);
  e.ld_aux = ld_aux; e.alpha = alpha; e.mode = epilogue; e.out_fp32 = out_fp32;
  e.drop_th2 = static_cast<uint32_t>(dropout_p * 32768.0f) * 0x10001u;
  e.drop_scale = 1.0f / (1.0f - dropout_p);
  e.keys = make_philox_keys7(philox_seed, philox_offset);
  e.seed_slot = dropout_p > 0.f ? seed_slot_override() : nullptr;
  e.a_colsum = a_colsum;

  // TMA needs 16-byte aligned bases and row pitches
  const bool tma_ok = (reinterpret_cast<uintptr_t>(A) % 16 == 0) &&
                      (reinterpret_cast<uintptr_t>(B) % 16 == 0) && (lda % 8 == 0) &&
                      (ldb % 8 == 0);
  const int elt = out_fp32 ? 4 : 2;
  const bool gelu_mode = epilogue >= VITSSL_EPI_BIAS_GELU;  // every epilogue with an aux operand (bf16 in/out)
  bool out_ok = (reinterpret_cast<uintptr_t>(C) % 16 == 0) && ((ldc * elt) % 16 == 0);
  if (gelu_mode)
    out_ok = out_ok && !out_fp32 && (reinterpret_cast<uintptr_t>(aux) % 16 == 0) && (ld_aux % 8 == 0);
  if (!tma_ok || (gelu_mode && !out_ok)) {
    VITSSL_REQUIRE(dropout_p == 0.f, VITSSL_ERR_SHAPE, "gemm: dropout unsupported on unaligned path");
    VITSSL_REQUIRE(a_colsum == nullptr, VITSSL_ERR_SHAPE, "gemm_rowsum: unaligned operands");
    e.atomic = 0; e.vec_ok = 0; e.tma_out = 0;
    dim3 grid((unsigned)((N + 31) / 32), (unsigned)((M + 31) / 32));
    gemm_simt_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(A),
                                               reinterpret_cast<const __nv_bfloat16*>(B), (int)M,
                                               (int)N, (int)K, lda, ldb, a_mn, b_mn, e);
    return check_launch("gemm_simt");
  }

  int bn = pick_block_n((int)N);
  {
    // experiment hook: VITSSL_GEMM_BN=<n> forces the tile width when it divides N (64/128/192/256)
    static const int bn_env = getenv("VITSSL_GEMM_BN") ? atoi(getenv("VITSSL_GEMM_BN")) : 0;
    if (bn_env > 0 && N % bn_env == 0 && (bn_env == 64 || bn_env == 128 || bn_env == 192 || bn_env == 256)) bn = bn_env;
  }
  // CTA pairs (cta_group::2) for the compute-bound GEMMs: 256-row tiles, B columns split across
  // the pair. Measured on B200: +12..19 % where both N and K are large (8192^3: 1162 -> 1388 TF/s,
  // ViT-B shapes 1240-1360 TF/s = cuBLAS parity); no gain on ViT-S shapes (N or K = 384), which are
  // bound by HBM and the epilogue, so those stay single-CTA. VITSSL_GEMM_PAIR=0/1/2 = off/auto/force.
  // An MN-major B operand is staged in 64-column boxes, so its per-CTA half must be a multiple of 64.
  static const int pair_env = getenv("VITSSL_GEMM_PAIR") ? atoi(getenv("VITSSL_GEMM_PAIR")) : 1;
  // Round 2, re-measured at the ViT-S shapes with pairs forced: the long-K plain / bias GEMMs gain
  // (FFN2 forward 50176x384x1536: 66.1 -> 59.8 us, FFN1 dgrad 64.5 -> 61.5, QKV dgrad 51.4 -> 49.0),
  // the arithmetic-heavy epilogues (GELU forms, MUL), K = 384 and the weight gradients do not.
  const bool long_k_plain = K >= 1024 && N >= 256 && !a_mn && (epilogue == VITSSL_EPI_NONE || epilogue == VITSSL_EPI_BIAS);
  bool pair = pair_env != 0 && M >= 1024 && (M % 256 == 0 || M >= 4096) && N >= 128 &&
              (pair_env == 2 || (N >= 512 && K >= 512) || long_k_plain);
  if (pair) {
    if (bn == 64) pair = false;
    if (b_mn && bn == 192) bn = 128;
  }
  if (a_colsum != nullptr) {
    // fused row sums need 32 spare TMEM columns (BN <= 192), the staged fp32 epilogue and a single CTA per tile
    if (bn == 256) bn = (N % 192 == 0) ? 192 : 128;
    VITSSL_REQUIRE(!pair && out_ok && out_fp32 && epilogue == VITSSL_EPI_NONE && 2 * bn + 32 <= 512, VITSSL_ERR_SHAPE,
                   "gemm_rowsum: configuration not supported (CTA-pair shape, bf16 output or unaligned C)");
  }
  const int tile_m = pair ? 2 * BLOCK_M : BLOCK_M;
  GemmShape s{};
  s.M = (int)M; s.N = (int)N; s.K = (int)K;
  s.m_tiles = (int)((M + tile_m - 1) / tile_m);
  s.n_tiles = (int)((N + bn - 1) / bn);
  s.kblocks_total = (int)((K + BLOCK_K - 1) / BLOCK_K);
  int splits = 1;
  if (split_k != 0 && out_fp32 && epilogue == VITSSL_EPI_NONE) {
    if (split_k > 0) {
      splits = split_k;
    } else {  // auto: one wave — tiles * splits <= SM count — and >= 4 k-blocks per split
      const int tiles = s.m_tiles * s.n_tiles;
      splits = (pair ? num_sms() / 2 : num_sms()) / tiles;
      const int max_splits = s.kblocks_total / 4 > 0 ? s.kblocks_total / 4 : 1;
      if (splits > max_splits) splits = max_splits;
    }
    if (splits < 1) splits = 1;
    if (splits > s.kblocks_total) splits = s.kblocks_total;
  }
  s.kblocks_per_split = (s.kblocks_total + splits - 1) / splits;
  s.splits = (s.kblocks_total + s.kblocks_per_split - 1) / s.kblocks_per_split;
  // split_k == -2 means "accumulate into C": with a single split the partial sum is still ADDED (TMA
  // reduce-add), so several calls can contribute to one gradient buffer (two passes over the same
  // encoder blocks accumulate their weight gradients into one zeroed buffer)
  e.atomic = (s.splits > 1 || (split_k == -2 && out_fp32 && epilogue == VITSSL_EPI_NONE)) ? 1 : 0;
  e.vec_ok = out_ok ? 1 : 0;
  e.tma_out = out_ok ? 1 : 0;  // split-K partials go through the same staging tiles as a TMA reduce-add
  if (e.atomic && split_k != -2) {  // -2: the caller hands over a zeroed C (one memset for many GEMMs)
    cudaError_t err = cudaMemset2DAsync(C, ldc * 4, 0, N * 4, M, stream);
    if (err != cudaSuccess) {
      set_error("gemm: memset failed: %s", cudaGetErrorString(err));
      return VITSSL_ERR_CUDA;
    }
  }

  CUtensorMap ta, tb, tc, tx;
  memset(&tc, 0, sizeof(tc));
  memset(&tx, 0, sizeof(tx));
  int rc;
  if (!a_mn) rc = make_tmap_bf16_2d(&ta, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, 64, BLOCK_M);
  else       rc = make_tmap_bf16_2d(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda * 2, 64, 64);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap_bf16_2d(&tb, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb * 2, 64, pair ? bn / 2 : bn);
  else       rc = make_tmap_bf16_2d(&tb, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb * 2, 64, 64);
  if (rc) return rc;
  if (e.tma_out) {
    // epilogue staging tiles: 32 rows x 32 columns (bf16: 64-byte rows, SWIZZLE_64B; fp32: 128B)
    rc = make_tmap_2d(&tc, C, elt, (uint64_t)N, (uint64_t)M, (uint64_t)ldc * elt, 32, 32,
                      out_fp32 ? 128 : 64);
    if (rc) return rc;
    if (gelu_mode) {
      rc = make_tmap_2d(&tx, aux, 2, (uint64_t)N, (uint64_t)M, (uint64_t)ld_aux * 2, 32, 32, 64);
      if (rc) return rc;
    }
  }

  if (!a_mn && !b_mn) return dispatch_bn<false, false>(bn, pair, ta, tb, tc, tx, s, e, stream);
  if (!a_mn && b_mn) return dispatch_bn<false, true>(bn, pair, ta, tb, tc, tx, s, e, stream);
  if (a_mn && b_mn) return dispatch_bn<true, true>(bn, pair, ta, tb, tc, tx, s, e, stream);
  return dispatch_bn<true, false>(bn, pair, ta, tb, tc, tx, s, e, stream);
}

extern "C" int vitssl_gemm_bf16(const void* A, const void* B, void* C, int64_t M, int64_t N,
                                int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_mn,
                                int b_mn, int epilogue, const float* bias, void* aux,
                                int64_t ld_aux, float alpha, int out_fp32, int split_k,
                                float dropout_p, uint64_t philox_seed, uint64_t philox_offset,
                                cudaStream_t stream) {
  return gemm_impl(A, B, C, M, N, K, lda, ldb, ldc, a_mn, b_mn, epilogue, bias, aux, ld_aux, alpha, out_fp32, split_k,
                   dropout_p, philox_seed, philox_offset, nullptr, stream);
}

extern "C" int vitssl_gemm_bf16_rowsum(const void* A, const void* B, void* C, float* a_rowsum, int64_t M, int64_t N,
                                       int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_mn, int b_mn,
                                       float alpha, int split_k, cudaStream_t stream) {
  VITSSL_REQUIRE(a_rowsum != nullptr, VITSSL_ERR_ARG, "gemm_rowsum: null a_rowsum");
  return gemm_impl(A, B, C, M, N, K, lda, ldb, ldc, a_mn, b_mn, VITSSL_EPI_NONE, nullptr, nullptr, 0, alpha, 1, split_k,
                   0.f, 0, 0, a_rowsum, stream);
}
