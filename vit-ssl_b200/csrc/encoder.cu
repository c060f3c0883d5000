// vitssl_b200 — host-side sequencing of a whole pre-LN encoder stack in ONE C-ABI call.
//
// Reference: the L-iteration hot loop `for block in encoder_blocks: x, attn = block(x)`
// (vit.py:35-37, ssl/simmim/model.py:52-55, ssl/dino/model.py:36-38) over EncoderBlock.forward
// (encoder_block.py:40-52). Issuing the ~11 (forward) / ~16 (backward) kernels of every block from
// Python costs ~10 us of host time per launch, which makes the step host-bound for the DINO
// multi-crop passes and for small models; here the caller hands over every buffer once and the
// launches are issued back to back from C++. No allocation, no synchronisation, same kernels and
// the same dropout stream layout as the per-op path (functional.py::_EncoderStackFn):
//   site 3l   : after the attention out-projection (applied inside the next add+LayerNorm)
//   site 3l+1 : after GELU (GEMM epilogue)        site 3l+2 : after the FFN (next add+LayerNorm)
#include "common.cuh"
#include "vitssl_b200.h"

using namespace vitssl;

#define VITSSL_TRY(call)          \
  do {                            \
    const int rc_ = (call);       \
    if (rc_ != 0) return rc_;     \
  } while (0)

// Per-launch timing while a profile is open (vitssl_profile_begin): CUDA events on `stream` around
// the launch, tagged with the kernel family and its algorithmic work, so that bench.py's roofline
// numbers are taken on THIS sequencing path — the one the timed step runs.
namespace {
struct Kind { char s[40]; };
inline Kind gemm_kind(int64_t M, int64_t N, int64_t K, int a_mn, int b_mn, int epi) {
  Kind k;
  snprintf(k.s, sizeof(k.s), "gemm|%lldx%lldx%lld|a_mn=%d b_mn=%d epi=%d", (long long)M, (long long)N, (long long)K,
           a_mn, b_mn, epi);
  return k;
}
inline double gemm_flops(int64_t M, int64_t N, int64_t K) { return 2.0 * (double)M * (double)N * (double)K; }
inline double attn_flops(int64_t B, int64_t H, int64_t S) { return 4.0 * (double)B * (double)H * (double)S * (double)S * 64.0; }
}  // namespace
#define VITSSL_TIMED(kind, work, call)        \
  do {                                        \
    ProfScope ps_((kind), (work), stream);    \
    VITSSL_TRY(call);                         \
  } while (0)

static int stack_fwd_body(const vitssl_encoder_fwd_args* a, cudaStream_t stream) {
  VITSSL_REQUIRE(a != nullptr && a->L >= 1, VITSSL_ERR_ARG, "encoder_stack_fwd: bad args");
  const int64_t B = a->B, S = a->S, D = a->D, H = a->H, F = a->F, L = a->L;
  const int64_t M = B * S;
  VITSSL_REQUIRE(D == H * 64 && vitssl_attention_supported(S, S, 64), VITSSL_ERR_SHAPE,
                 "encoder_stack_fwd: needs d_head = 64 and S <= 256 (got D=%lld H=%lld S=%lld)",
                 (long long)D, (long long)H, (long long)S);
  const float p = a->dropout_p;
  const float scale = 0.125f;  // 1 / sqrt(64)
  const float* stream_in = a->x_in;  // fp32 residual stream entering the block
  const void* branch = nullptr;      // bf16 branch output waiting to be added
  for (int64_t l = 0; l < L; ++l) {
    // x = x + drop(branch); xn1 = LN1(x)                                   encoder_block.py:40-41
    float* xs = (l == 0) ? nullptr : a->xs[l];
    VITSSL_TIMED("add_layernorm", (double)M * D * (branch ? 12 : 6),
                 vitssl_add_layernorm_fwd(stream_in, D, branch, xs, a->g1[l], a->be1[l], a->xn1[l], a->mean1[l],
                                          a->rstd1[l], M, D, a->eps, branch ? p : 0.f, a->seed,
                                          (uint64_t)(l > 0 ? 3 * l - 1 : 0), stream));
    const float* x_l = (l == 0) ? a->x_in : xs;
    // fused QKV projection                                                   attention.py:82-84
    VITSSL_TIMED(gemm_kind(M, 3 * D, D, 0, 0, 0).s, gemm_flops(M, 3 * D, D),
                 vitssl_gemm_bf16(a->xn1[l], a->wqkv[l], a->qkv[l], M, 3 * D, D, D, D, 3 * D, 0, 0, VITSSL_EPI_NONE,
                                  nullptr, nullptr, 0, 1.0f, 0, 0, 0.f, 0, 0, stream));
    const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(a->qkv[l]);
    VITSSL_TIMED("attn_fwd", attn_flops(B, H, S),
                 vitssl_attention_fwd(qkv, qkv + D, qkv + 2 * D, 3 * D, 3 * D, 3 * D, a->ctx[l],
                                      a->ctx_lo ? a->ctx_lo[l] : nullptr, D, a->lse[l], B, H, S, S, scale,
                                      stream));                               // attention.py:20-23
    VITSSL_TIMED(gemm_kind(M, D, D, 0, 0, 0).s, gemm_flops(M, D, D),
                 vitssl_gemm_bf16(a->ctx[l], a->wo[l], a->y1, M, D, D, D, D, D, 0, 0, VITSSL_EPI_NONE, nullptr,
                                  nullptr, 0, 1.0f, 0, 0, 0.f, 0, 0, stream));  // attention.py:105
    // x = x + drop(attn); xn2 = LN2(x)                                     encoder_block.py:45-48
    VITSSL_TIMED("add_layernorm", (double)M * D * 12,
                 vitssl_add_layernorm_fwd(x_l, D, a->y1, a->xmid[l], a->g2[l], a->be2[l], a->xn2[l], a->mean2[l],
                                          a->rstd2[l], M, D, a->eps, p, a->seed, (uint64_t)(3 * l), stream));
    // FFN: u = xn2 W1^T + b1, h = drop(gelu(u)), saved: h and mask/(1-p) gelu'(u); y2 = h W2^T + b2   feed_forward.py:26-28
    VITSSL_TIMED(gemm_kind(M, F, D, 0, 0, VITSSL_EPI_BIAS_GELU_D).s, gemm_flops(M, F, D),
                 vitssl_gemm_bf16(a->xn2[l], a->w1[l], a->h[l], M, F, D, D, D, F, 0, 0, VITSSL_EPI_BIAS_GELU_D, a->b1[l],
                                  a->u[l], F, 1.0f, 0, 0, p, a->seed, (uint64_t)(3 * l + 1), stream));
    void* y2 = a->y2[l & 1];
    VITSSL_TIMED(gemm_kind(M, D, F, 0, 0, 1).s, gemm_flops(M, D, F),
                 vitssl_gemm_bf16(a->h[l], a->w2[l], y2, M, D, F, F, F, D, 0, 0, VITSSL_EPI_BIAS, a->b2[l], nullptr, 0,
                                  1.0f, 0, 0, 0.f, 0, 0, stream));
    stream_in = a->xmid[l];
    branch = y2;
  }
  // final residual add: out = x + drop(ffn)                                 encoder_block.py:51-52
  VITSSL_TIMED("add_layernorm", (double)M * D * 10,
               vitssl_add_layernorm_fwd(stream_in, D, branch, a->out, nullptr, nullptr, nullptr, nullptr, nullptr, M, D,
                                        a->eps, p, a->seed, (uint64_t)(3 * L - 1), stream));
  return 0;
}

static int stack_bwd_body(const vitssl_encoder_bwd_args* a, cudaStream_t stream) {
  VITSSL_REQUIRE(a != nullptr && a->fwd != nullptr && a->fwd->L >= 1 && a->delta != nullptr, VITSSL_ERR_ARG,
                 "encoder_stack_bwd: bad args");
  const vitssl_encoder_fwd_args* f = a->fwd;
  const int64_t B = f->B, S = f->S, D = f->D, H = f->H, F = f->F, L = f->L;
  const int64_t M = B * S;
  const float p = f->dropout_p;
  const float scale = 0.125f;
  const int64_t l_hi = a->l_end > 0 ? a->l_end : L, l_lo = a->l_begin;
  VITSSL_REQUIRE(0 <= l_lo && l_lo < l_hi && l_hi <= L, VITSSL_ERR_ARG, "encoder_stack_bwd: bad layer range [%lld, %lld)",
                 (long long)l_lo, (long long)l_hi);
  const float* gs;  // gradient on the residual stream
  if (l_hi == L) {
    // gradient of the final add: the stream gradient passes through, the branch gets mask/(1-p) * g
    VITSSL_TIMED("add_layernorm", (double)M * D * 6,
                 vitssl_add_layernorm_bwd_acc(nullptr, nullptr, 0, nullptr, nullptr, nullptr, a->gout, D, nullptr, 0,
                                              a->dbranch, nullptr, nullptr, M, D, p, f->seed, (uint64_t)(3 * L - 1), stream));
    gs = a->gout;
  } else {
    gs = a->gs[1];  // left there (with dbranch) by the previous call
  }
  for (int64_t l = l_hi - 1; l >= l_lo; --l) {
    const void* dy2 = a->dbranch;
    // FFN backward
    VITSSL_TIMED(gemm_kind(M, F, D, 0, 1, VITSSL_EPI_MUL).s, gemm_flops(M, F, D),
                 vitssl_gemm_bf16(dy2, f->w2[l], a->du, M, F, D, D, F, F, 0, 1, VITSSL_EPI_MUL, nullptr, f->u[l], F,
                                  1.0f, 0, 0, 0.f, 0, 0, stream));
    // dW2 = dy2^T h with db2 = colsum(dy2) from the same kernel (ones-operand MMA); shapes the fused
    // form does not cover (CTA pairs) fall back to GEMM + column-sum pass
    {
      ProfScope ps_(gemm_kind(D, F, M, 1, 1, 0).s, gemm_flops(D, F, M), stream);
      if (vitssl_gemm_bf16_rowsum(dy2, f->h[l], a->dw2[l], a->db2[l], D, F, M, D, F, F, 1, 1, 1.0f, -2, stream) != 0) {
        VITSSL_TRY(vitssl_gemm_bf16(dy2, f->h[l], a->dw2[l], D, F, M, D, F, F, 1, 1, VITSSL_EPI_NONE, nullptr, nullptr, 0,
                                    1.0f, 1, -2, 0.f, 0, 0, stream));
        VITSSL_TRY(vitssl_colsum_bf16_acc(dy2, D, M, D, a->db2[l], stream));
      }
    }
    VITSSL_TIMED(gemm_kind(M, D, F, 0, 1, 0).s, gemm_flops(M, D, F),
                 vitssl_gemm_bf16(a->du, f->w1[l], a->dxn, M, D, F, F, D, D, 0, 1, VITSSL_EPI_NONE, nullptr, nullptr, 0,
                                  1.0f, 0, 0, 0.f, 0, 0, stream));
    {
      ProfScope ps_(gemm_kind(F, D, M, 1, 1, 0).s, gemm_flops(F, D, M), stream);
      if (vitssl_gemm_bf16_rowsum(a->du, f->xn2[l], a->dw1[l], a->db1[l], F, D, M, F, D, D, 1, 1, 1.0f, -2, stream) != 0) {
        VITSSL_TRY(vitssl_gemm_bf16(a->du, f->xn2[l], a->dw1[l], F, D, M, F, D, D, 1, 1, VITSSL_EPI_NONE, nullptr, nullptr,
                                    0, 1.0f, 1, -2, 0.f, 0, 0, stream));
        VITSSL_TRY(vitssl_colsum_bf16_acc(a->du, F, M, F, a->db1[l], stream));
      }
    }
    // LN2 + residual: gs <- d(xmid), dy1 = dropout-masked gradient of the attention branch
    float* gs_mid = a->gs[0];
    VITSSL_TIMED("add_layernorm", (double)M * D * 16,
                 vitssl_add_layernorm_bwd_acc(a->dxn, f->xmid[l], D, f->mean2[l], f->rstd2[l], f->g2[l], gs, D, gs_mid, D,
                                              a->dbranch, a->dg2[l], a->dbe2[l], M, D, p, f->seed, (uint64_t)(3 * l),
                                              stream));
    const void* dy1 = a->dbranch;
    VITSSL_TIMED(gemm_kind(M, D, D, 0, 1, 0).s, gemm_flops(M, D, D),
                 vitssl_gemm_bf16(dy1, f->wo[l], a->dctx, M, D, D, D, D, D, 0, 1, VITSSL_EPI_NONE, nullptr, nullptr, 0,
                                  1.0f, 0, 0, 0.f, 0, 0, stream));
    VITSSL_TIMED(gemm_kind(D, D, M, 1, 1, 0).s, gemm_flops(D, D, M),
                 vitssl_gemm_bf16(dy1, f->ctx[l], a->dwo[l], D, D, M, D, D, D, 1, 1, VITSSL_EPI_NONE, nullptr, nullptr, 0,
                                  1.0f, 1, -2, 0.f, 0, 0, stream));
    const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(f->qkv[l]);
    __nv_bfloat16* dqkv = reinterpret_cast<__nv_bfloat16*>(a->dqkv);
    VITSSL_TIMED("attn_bwd", 2.5 * attn_flops(B, H, S),
                 vitssl_attention_bwd(qkv, qkv + D, qkv + 2 * D, 3 * D, 3 * D, 3 * D, f->ctx[l],
                                      f->ctx_lo ? f->ctx_lo[l] : nullptr, a->dctx, D, f->lse[l], a->delta, dqkv, 3 * D,
                                      dqkv + D, 3 * D, dqkv + 2 * D, 3 * D, B, H, S, S, scale, stream));
    VITSSL_TIMED(gemm_kind(M, D, 3 * D, 0, 1, 0).s, gemm_flops(M, D, 3 * D),
                 vitssl_gemm_bf16(dqkv, f->wqkv[l], a->dxn, M, D, 3 * D, 3 * D, D, D, 0, 1, VITSSL_EPI_NONE, nullptr,
                                  nullptr, 0, 1.0f, 0, 0, 0.f, 0, 0, stream));
    VITSSL_TIMED(gemm_kind(3 * D, D, M, 1, 1, 0).s, gemm_flops(3 * D, D, M),
                 vitssl_gemm_bf16(dqkv, f->xn1[l], a->dwqkv[l], 3 * D, D, M, 3 * D, D, D, 1, 1, VITSSL_EPI_NONE, nullptr,
                                  nullptr, 0, 1.0f, 1, -2, 0.f, 0, 0, stream));
    // LN1 + residual: gs <- d(block input); for l > 0 also the masked gradient of block l-1's FFN
    const float* x_l = (l == 0) ? f->x_in : f->xs[l];
    float* gs_in = (l == 0) ? a->dx : a->gs[1];
    VITSSL_TIMED("add_layernorm", (double)M * D * (l > 0 ? 16 : 14),
                 vitssl_add_layernorm_bwd_acc(a->dxn, x_l, D, f->mean1[l], f->rstd1[l], f->g1[l], gs_mid, D, gs_in, D,
                                              l > 0 ? a->dbranch : nullptr, a->dg1[l], a->dbe1[l], M, D, l > 0 ? p : 0.f,
                                              f->seed, (uint64_t)(l > 0 ? 3 * l - 1 : 0), stream));
    gs = gs_in;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// CUDA-graph replay of a stack call. The ~100 (forward) / ~150 (backward) launches of a stack are
// the same every step when the caller's buffers sit at the same addresses (the steady state of a
// caching allocator in a training loop); between dependent launches in a stream the GPU idles 2-3.5 us
// (profiles/r2_gaps.md: 0.53 ms of a 14.7 ms SimMIM step, 1.3 ms of a DINO step), a graph replay of
// the same kernels does not (scripts/micro/graph_gap.py: -2.0 us per launch). A call is keyed on
// every scalar and pointer it hands over; the second call with a known key is captured
// (cudaStreamBeginCapture around the very same sequencing code) and instantiated, later ones are
// one cudaGraphLaunch. The only per-step value, the dropout seed, travels through a device word
// that the dropout kernels read (layernorm.cu / gemm_sm100.cu: seed_slot_override) and that a
// one-thread kernel sets ahead of the replay, in stream order. A key that misses takes the direct
// path, so results never depend on the cache. VITSSL_GRAPH=0 disables; profiling bypasses it.
// ------------------------------------------------------------------------------------------
#include <atomic>
#include <mutex>
#include <vector>

namespace {
__device__ unsigned long long g_seed_word;
__global__ void set_seed_kernel(unsigned long long s) { g_seed_word = s; }

struct GraphEntry {
  std::vector<unsigned char> key;
  cudaGraphExec_t exec = nullptr;
  int seen = 0;
  long long launches = 0;  // kernels and memsets the captured sequence holds
  bool bad = false;  // capture or instantiation failed once: stay on the direct path
  unsigned long long stamp = 0;
};
std::mutex g_graph_mu;
std::vector<GraphEntry> g_graphs;
unsigned long long g_graph_stamp = 0;
long long g_graph_captured = 0, g_graph_replayed = 0;
constexpr size_t GRAPH_CACHE_ENTRIES = 24;
thread_local const unsigned long long* tl_seed_slot = nullptr;

std::atomic<bool> g_graph_user_on{true};  // vitssl_graph_enable()
bool graphs_enabled() {
  static const bool on = !(getenv("VITSSL_GRAPH") && atoi(getenv("VITSSL_GRAPH")) == 0);
  return on && g_graph_user_on.load(std::memory_order_relaxed);
}
const unsigned long long* seed_word_address() {
  static const unsigned long long* addr = [] {
    void* p = nullptr;
    return cudaGetSymbolAddress(&p, g_seed_word) == cudaSuccess ? static_cast<const unsigned long long*>(p) : nullptr;
  }();
  return addr;
}
template <typename T>
void key_put(std::vector<unsigned char>& k, const T& v) {
  const unsigned char* b = reinterpret_cast<const unsigned char*>(&v);
  k.insert(k.end(), b, b + sizeof(T));
}
template <typename P>
void key_put_array(std::vector<unsigned char>& k, P const* arr, int64_t L) {
  if (arr == nullptr) { key_put(k, (void*)nullptr); return; }
  for (int64_t l = 0; l < L; ++l) key_put(k, arr[l]);
}
void key_fwd(std::vector<unsigned char>& k, const vitssl_encoder_fwd_args* a) {
  key_put(k, a->B); key_put(k, a->S); key_put(k, a->D); key_put(k, a->H); key_put(k, a->F); key_put(k, a->L);
  key_put(k, a->dropout_p); key_put(k, a->eps);
  key_put(k, a->x_in); key_put(k, a->out); key_put(k, a->y1); key_put(k, a->y2[0]); key_put(k, a->y2[1]);
  const int64_t L = a->L;
  key_put_array(k, a->wqkv, L); key_put_array(k, a->wo, L); key_put_array(k, a->w1, L); key_put_array(k, a->w2, L);
  key_put_array(k, a->b1, L); key_put_array(k, a->b2, L); key_put_array(k, a->g1, L); key_put_array(k, a->be1, L);
  key_put_array(k, a->g2, L); key_put_array(k, a->be2, L); key_put_array(k, a->xs, L); key_put_array(k, a->mean1, L);
  key_put_array(k, a->rstd1, L); key_put_array(k, a->xn1, L); key_put_array(k, a->qkv, L); key_put_array(k, a->ctx, L);
  key_put_array(k, a->lse, L); key_put_array(k, a->ctx_lo, L); key_put_array(k, a->xmid, L); key_put_array(k, a->mean2, L);
  key_put_array(k, a->rstd2, L); key_put_array(k, a->xn2, L); key_put_array(k, a->u, L); key_put_array(k, a->h, L);
}
void key_bwd(std::vector<unsigned char>& k, const vitssl_encoder_bwd_args* a) {
  key_fwd(k, a->fwd);
  key_put(k, a->gout); key_put(k, a->dx); key_put(k, a->dbranch); key_put(k, a->du); key_put(k, a->dxn); key_put(k, a->dctx);
  key_put(k, a->dqkv); key_put(k, a->delta); key_put(k, a->gs[0]); key_put(k, a->gs[1]);
  const int64_t L = a->fwd->L;
  key_put_array(k, a->dwqkv, L); key_put_array(k, a->dwo, L); key_put_array(k, a->dw1, L); key_put_array(k, a->db1, L);
  key_put_array(k, a->dw2, L); key_put_array(k, a->db2, L); key_put_array(k, a->dg1, L); key_put_array(k, a->dbe1, L);
  key_put_array(k, a->dg2, L); key_put_array(k, a->dbe2, L);
  key_put(k, a->l_begin); key_put(k, a->l_end);
}

// run `body` (which sequences the launches on `stream`) directly, or as a captured / replayed graph
template <typename Body>
int graph_dispatch(std::vector<unsigned char>& key, unsigned long long seed, bool uses_seed, cudaStream_t stream, Body&& body) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  const unsigned long long* slot = seed_word_address();
  if (!graphs_enabled() || prof_on() || slot == nullptr || cudaStreamIsCapturing(stream, &cs) != cudaSuccess ||
      cs != cudaStreamCaptureStatusNone)
    return body(stream);
  // the sequence is captured on a stream of our own (the caller's may be the legacy default stream,
  // which cannot be captured) and the instantiated graph is launched on the caller's stream
  static cudaStream_t cap_stream = [] {
    cudaStream_t st = nullptr;
    return cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess ? st : nullptr;
  }();
  if (cap_stream == nullptr) return body(stream);
  std::unique_lock<std::mutex> lock(g_graph_mu);
  GraphEntry* e = nullptr;
  for (auto& g : g_graphs)
    if (g.key == key) { e = &g; break; }
  if (e == nullptr) {
    if (g_graphs.size() >= GRAPH_CACHE_ENTRIES) {  // evict the least recently used entry
      size_t victim = 0;
      for (size_t i = 1; i < g_graphs.size(); ++i)
        if (g_graphs[i].stamp < g_graphs[victim].stamp) victim = i;
      if (g_graphs[victim].exec) cudaGraphExecDestroy(g_graphs[victim].exec);
      g_graphs.erase(g_graphs.begin() + victim);
    }
    g_graphs.emplace_back();
    e = &g_graphs.back();
    e->key = key;
  }
  e->stamp = ++g_graph_stamp;
  e->seen += 1;
  if (e->bad || e->seen < 2) {  // first sighting (one-off shapes never pay for a capture) or known failure
    lock.unlock();
    return body(stream);
  }
  if (e->exec == nullptr && g_graph_captured >= 48 && g_graph_replayed < 2 * g_graph_captured) {
    // the caller's buffer addresses do not repeat (captures are not being reused): stop paying for them
    g_graph_user_on.store(false);
    lock.unlock();
    return body(stream);
  }
  if (e->exec == nullptr) {
    // capture the same sequencing code; the dropout kernels read the seed word instead of their argument
    if (cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      e->bad = true;
      cudaGetLastError();
      lock.unlock();
      return body(stream);
    }
    tl_seed_slot = slot;
    const long long n0 = launches_now();
    const int rc = body(cap_stream);
    e->launches = launches_now() - n0;
    add_launches(-e->launches);  // counted below, once per replay
    tl_seed_slot = nullptr;
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(cap_stream, &graph);
    if (rc != 0 || ce != cudaSuccess || graph == nullptr ||
        cudaGraphInstantiate(&e->exec, graph, 0) != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      e->exec = nullptr;
      e->bad = true;
      cudaGetLastError();
      lock.unlock();
      return rc != 0 ? rc : body(stream);  // nothing was launched by the capture: run it for real
    }
    cudaGraphDestroy(graph);
    g_graph_captured += 1;
  }
  cudaGraphExec_t exec = e->exec;
  g_graph_replayed += 1;
  add_launches(e->launches);
  if (uses_seed) set_seed_kernel<<<1, 1, 0, stream>>>(seed);
  const cudaError_t le = cudaGraphLaunch(exec, stream);
  lock.unlock();
  if (le != cudaSuccess) {
    set_error("encoder_stack: graph launch failed: %s", cudaGetErrorString(le));
    return VITSSL_ERR_CUDA;
  }
  return check_launch("encoder_stack_graph");
}
}  // namespace

namespace vitssl {
const unsigned long long* seed_slot_override() { return tl_seed_slot; }
}  // namespace vitssl

extern "C" int vitssl_encoder_stack_fwd(const vitssl_encoder_fwd_args* a, cudaStream_t stream) {
  if (a == nullptr || a->L < 1) return stack_fwd_body(a, stream);
  std::vector<unsigned char> key;
  key.reserve(4096);
  key_put(key, 'F');
  key_fwd(key, a);
  return graph_dispatch(key, a->seed, a->dropout_p > 0.f, stream, [&](cudaStream_t st) { return stack_fwd_body(a, st); });
}

extern "C" int vitssl_encoder_stack_bwd(const vitssl_encoder_bwd_args* a, cudaStream_t stream) {
  if (a == nullptr || a->fwd == nullptr || a->fwd->L < 1) return stack_bwd_body(a, stream);
  std::vector<unsigned char> key;
  key.reserve(4096);
  key_put(key, 'B');
  key_bwd(key, a);
  return graph_dispatch(key, a->fwd->seed, a->fwd->dropout_p > 0.f, stream,
                        [&](cudaStream_t st) { return stack_bwd_body(a, st); });
}

extern "C" int vitssl_graph_enable(int on) {
  g_graph_user_on.store(on != 0);
  return 0;
}

extern "C" int vitssl_graph_stats(int64_t* captured, int64_t* replayed) {
  std::lock_guard<std::mutex> lock(g_graph_mu);
  if (captured) *captured = g_graph_captured;
  if (replayed) *replayed = g_graph_replayed;
  return 0;
}
