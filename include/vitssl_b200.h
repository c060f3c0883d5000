/* vitssl_b200 — C ABI of the B200-native ViT-SSL training hot path.
 *
 * The reference (kristi700/ViT-SSL) has no FFI layer: its hot path is the Python nn.Module API of
 * package `vit_core`. This library is what our drop-in `vit_core` package binds with ctypes; each
 * entry point cites the reference code it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless named host_*; the caller owns all buffers; the
 *    library never allocates, frees or synchronises; every call is enqueued on `stream`;
 *  - returns 0 on success, a negative VITSSL_ERR_* code otherwise; vitssl_last_error() gives the
 *    thread-local message;
 *  - "bf16" buffers are raw 16-bit bfloat16; activations on the residual stream are fp32
 *    (the reference under autocast keeps an fp32 stream: encoder_block.py:46,52);
 *  - dropout masks are never stored: they are regenerated from (philox_seed, philox_offset).
 */
#ifndef VITSSL_B200_H_
#define VITSSL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* vitssl_stream_t; /* == cudaStream_t */

#define VITSSL_ERR_ARG (-1)
#define VITSSL_ERR_SHAPE (-2)
#define VITSSL_ERR_CUDA (-3)
#define VITSSL_ERR_DEVICE (-4)
#define VITSSL_ERR_WORKSPACE (-5)

/* ---- runtime ------------------------------------------------------------------------- */
int vitssl_version(void);              /* major*10000 + minor*100 + patch */
const char* vitssl_last_error(void);   /* thread-local, never NULL */
int vitssl_device_check(void);         /* 0 iff the current device is sm_100 (B200) */
int vitssl_num_sms(void);
/* number of kernels this library has launched (all threads of the process) since the last reset
 * (bench.py reports it as gpu_launches) */
int64_t vitssl_launch_count(int reset);

/* Per-launch device timing of the C-sequenced encoder stack (bench.py's roofline numbers): while a
 * profile is open every kernel family the stack launches is bracketed by CUDA events on its stream.
 * vitssl_profile_end() closes it and returns the number of records; vitssl_profile_read() is the ONE
 * entry point that waits for the device (cudaEventSynchronize on the record). `kind` is
 * "gemm|MxNxK|a_mn=. b_mn=. epi=." | "attn_fwd" | "attn_bwd" | "add_layernorm"; `work` is the
 * algorithmic FLOPs (tensor-bound families) or bytes (add_layernorm). */
int vitssl_profile_begin(void);
int64_t vitssl_profile_end(void);
int vitssl_profile_read(int64_t index, char* kind_out, int64_t kind_cap, double* work, float* ms);

/* ---- GEMM: every nn.Linear / Conv2d-patchify on the path -------------------------------
 * C[M,N] = alpha * op(A) * op(B), bf16 inputs, fp32 accumulation (tcgen05.mma, TMEM).
 *   a_mn = 0: A stored [M][K] (pitch lda)      a_mn = 1: A stored [K][M]
 *   b_mn = 0: B stored [N][K] (nn.Linear weight) b_mn = 1: B stored [K][N]
 * forward  y = x W^T   : a_mn=0,b_mn=0 (attention.py:82-84,105; feed_forward.py:26,28)
 * dgrad    dx = dy W   : a_mn=0,b_mn=1        wgrad dW = dy^T x : a_mn=1,b_mn=1
 * epilogue: NONE | BIAS (+bias[N]) | BIAS_GELU (aux <- bf16(acc+bias), C <- dropout(gelu(aux));
 *           feed_forward.py:26-27) | DGELU (C <- acc * mask/(1-p) * gelu'(aux)).
 * out_fp32: C is fp32, else bf16. split_k: 0 = off, -1 = auto, n = n splits (fp32 NONE only;
 * C is zeroed on `stream` and the partial sums are accumulated with TMA reduce-add), -2 = auto
 * and ACCUMULATE into C, which the caller initialised (one memset can then serve many GEMMs, and
 * several calls — e.g. two passes over the same layer — can add into one gradient buffer). */
#define VITSSL_EPI_NONE 0
#define VITSSL_EPI_BIAS 1
#define VITSSL_EPI_BIAS_GELU 2
#define VITSSL_EPI_DGELU 3
/* The pair the encoder stack uses (feed_forward.py:26-27 and its backward): the forward GEMM saves the
 * whole backward factor instead of the pre-activation, so the backward epilogue is one multiply —
 *   BIAS_GELU_D: u = bf16(acc+bias); C <- dropout(gelu(u)) (bf16); aux <- mask/(1-p) * gelu'(u)
 *   MUL:         C <- alpha * acc * aux (bf16)
 * For this pair aux holds IEEE fp16 values (2 bytes, same pitch rules): the factor lies in
 * [-0.2, 1.2]/(1-p), where fp16's 11 significant bits make the saved copy as good as recomputing. */
#define VITSSL_EPI_BIAS_GELU_D 4
#define VITSSL_EPI_MUL 5
int vitssl_gemm_bf16(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K,
                     int64_t lda, int64_t ldb, int64_t ldc, int a_mn, int b_mn, int epilogue,
                     const float* bias, void* aux, int64_t ld_aux, float alpha, int out_fp32,
                     int split_k, float dropout_p, uint64_t philox_seed, uint64_t philox_offset,
                     vitssl_stream_t stream);

/* fp32 C = alpha * op(A) op(B) as above (epilogue NONE) and, fused into the same kernel,
 * a_rowsum[m] += alpha * sum_k op(A)[m, k]: with A = dY^T (a_mn = 1) this is the bias gradient of
 * the Linear whose weight gradient the GEMM computes (feed_forward.py:26,28), obtained as one extra
 * 16-column tcgen05.mma per k-step against a constant all-ones operand instead of a separate pass
 * over dY. a_rowsum (fp32 [M]) is ACCUMULATED into: the caller zeroes it. Returns VITSSL_ERR_SHAPE
 * for configurations it does not cover (CTA-pair shapes, unaligned operands); callers then fall
 * back to vitssl_gemm_bf16 + vitssl_colsum_bf16. split_k as for vitssl_gemm_bf16. */
int vitssl_gemm_bf16_rowsum(const void* A, const void* B, void* C, float* a_rowsum, int64_t M,
                            int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_mn,
                            int b_mn, float alpha, int split_k, vitssl_stream_t stream);

/* ---- fused residual-add (+dropout) + LayerNorm ------------------------------------------
 * encoder_block.py:40-52:  x_out = x + dropout(branch);  y = LayerNorm(x_out) (eps, affine).
 *   x      fp32 rows of pitch ldx (elements);  branch bf16 dense [rows,D] or NULL (then x_out
 *   must be NULL and the LayerNorm reads x);  gamma/beta NULL -> add only (y/mean/rstd unused).
 *   y bf16 dense; mean/rstd fp32 [rows] are saved for backward. */
int vitssl_add_layernorm_fwd(const float* x, int64_t ldx, const void* branch, float* x_out,
                             const float* gamma, const float* beta, void* y, float* mean,
                             float* rstd, int64_t rows, int64_t D, float eps, float dropout_p,
                             uint64_t philox_seed, uint64_t philox_offset, vitssl_stream_t stream);
/* backward of the above. dy (bf16, NULL for the add-only form) is the gradient of y; dres (fp32,
 * nullable) the gradient arriving on the residual stream. Outputs: dx (fp32, total gradient of
 * x_out, which is also the gradient of x) and dbranch (bf16, nullable) = dropout-mask * dx.
 * dgamma/dbeta [D] are zeroed on `stream` and accumulated. */
int vitssl_add_layernorm_bwd(const void* dy, const float* x, int64_t ldx, const float* mean,
                             const float* rstd, const float* gamma, const float* dres,
                             int64_t ld_dres, float* dx, int64_t ld_dx, void* dbranch,
                             float* dgamma, float* dbeta, int64_t rows, int64_t D, float dropout_p,
                             uint64_t philox_seed, uint64_t philox_offset, vitssl_stream_t stream);
/* same, but dgamma/dbeta are accumulated into (the caller zeroed them, e.g. with one memset for
 * every layer's gradients) */
int vitssl_add_layernorm_bwd_acc(const void* dy, const float* x, int64_t ldx, const float* mean,
                                 const float* rstd, const float* gamma, const float* dres,
                                 int64_t ld_dres, float* dx, int64_t ld_dx, void* dbranch,
                                 float* dgamma, float* dbeta, int64_t rows, int64_t D, float dropout_p,
                                 uint64_t philox_seed, uint64_t philox_offset, vitssl_stream_t stream);

/* ---- multi-head attention (attention.py:5-27, 86-103) -----------------------------------
 * q/k/v/out are bf16 [B, S, H, 64] views: token rows of pitch ld* elements, head h at column
 * 64*h (so the fused QKV GEMM output is addressed in place). tcgen05 path: d_head = 64,
 * Sk <= 256 (and Sq <= 256 for backward). lse fp32 [B,H,Sq] is saved by forward for backward.
 * out_lo (nullable; training) receives bf16(O - bf16(O)) in the layout of `out`: the backward's
 * delta = rowsum(O * dO) is evaluated from out + out_lo (16 significant bits) because its rounding
 * error enters dS = P (dP - delta) coherently along a row and, on real activations, dominated the
 * error of dQ when taken from the bf16 context alone. `delta` is a caller-provided fp32 [B,H,Sq]
 * workspace that vitssl_attention_bwd fills itself (one extra HBM-bound launch).
 * Self-attention with Sq == Sk <= 64 (DINO local crops, cfg 1) runs with floor(128 / S) images per
 * 128-row tile and block-diagonal masking; this needs the images' token rows to be contiguous
 * (row b*S + s at pitch ld*, which is what every caller passes) and changes no result. */
int vitssl_attention_supported(int64_t Sq, int64_t Sk, int64_t d_head);
int vitssl_attention_fwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk,
                         int64_t ldv, void* out, void* out_lo, int64_t ldo, float* lse, int64_t B,
                         int64_t H, int64_t Sq, int64_t Sk, float scale, vitssl_stream_t stream);
int vitssl_attention_bwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk,
                         int64_t ldv, const void* out, const void* out_lo, const void* d_out,
                         int64_t ldo, const float* lse, float* delta, void* dq, int64_t lddq, void* dk,
                         int64_t lddk, void* dv, int64_t lddv, int64_t B, int64_t H, int64_t Sq,
                         int64_t Sk, float scale, vitssl_stream_t stream);
/* generic SIMT attention: any head dim / length, explicit element strides
 * host_strides[12] = {q_b,q_h,q_s, k_b,k_h,k_s, v_b,v_h,v_s, o_b,o_h,o_s}; probs (fp32
 * [B,H,Sq,Sk]) and lse are optional outputs (return_attn=True path, attention.py:24-25).
 * backward accumulates dK/dV into pre-zeroed fp32 [B,Sk,H,d] buffers. */
int vitssl_attention_generic_fwd(const void* q, const void* k, const void* v,
                                 const int64_t* host_strides, void* out, float* probs, float* lse,
                                 int64_t B, int64_t H, int64_t Sq, int64_t Sk, int64_t d,
                                 float scale, vitssl_stream_t stream);
int vitssl_attention_generic_bwd(const void* q, const void* k, const void* v,
                                 const int64_t* host_strides, const void* out, const void* d_out,
                                 const float* lse, void* dq, float* dk_acc, float* dv_acc,
                                 int64_t B, int64_t H, int64_t Sq, int64_t Sk, int64_t d,
                                 float scale, vitssl_stream_t stream);

/* ---- whole encoder stack in one call (host-side launch sequencing) ---------------------------
 * The hot loop `for block in encoder_blocks` (vit.py:35-37, ssl/simmim/model.py:52-55,
 * ssl/dino/model.py:36-38) over EncoderBlock.forward (encoder_block.py:40-52): every kernel of all
 * L pre-LN blocks is enqueued from C++, forward or backward, with buffers supplied by the caller.
 * Arrays named `const T* const*` are HOST arrays of L DEVICE pointers. d_head must be 64, S <= 256.
 * Layer l reads the stream from x_in (l = 0) or xs[l]; xs[0] is unused. Saved-for-backward
 * buffers: xs, mean1/rstd1, xn1, qkv ([M,3D]), ctx, ctx_lo, lse, xmid, mean2/rstd2, xn2, u, h. In
 * inference every layer may point at the same buffers. */
typedef struct {
  int64_t B, S, D, H, F, L;
  float dropout_p, eps;
  uint64_t seed;
  const float* x_in;              /* fp32 [B*S, D] */
  float* out;                     /* fp32 [B*S, D] */
  void* y1;                       /* bf16 [B*S, D] scratch */
  void* y2[2];                    /* bf16 [B*S, D] scratch (ping-pong) */
  const void* const* wqkv;        /* bf16 [3D, D]: rows = (w_query; w_key; w_value) */
  const void* const* wo;          /* bf16 [D, D] */
  const void* const* w1;          /* bf16 [F, D] */
  const void* const* w2;          /* bf16 [D, F] */
  const float* const* b1; const float* const* b2;
  const float* const* g1; const float* const* be1; const float* const* g2; const float* const* be2;
  float* const* xs; float* const* mean1; float* const* rstd1; void* const* xn1;
  void* const* qkv; void* const* ctx; float* const* lse;
  void* const* ctx_lo;            /* bf16 [B*S, D] per layer: rounding residual of ctx (NULL array or NULL entries in inference) */
  float* const* xmid; float* const* mean2; float* const* rstd2; void* const* xn2;
  void* const* u; void* const* h;  /* u: saved GELU backward factor mask/(1-p) * gelu'(pre-activation) */
} vitssl_encoder_fwd_args;
int vitssl_encoder_stack_fwd(const vitssl_encoder_fwd_args* args, vitssl_stream_t stream);

typedef struct {
  const vitssl_encoder_fwd_args* fwd;  /* the forward call's arguments (saved buffers, weights) */
  const float* gout;                   /* fp32 [B*S, D] gradient of `out` */
  float* dx;                           /* fp32 [B*S, D] gradient of x_in */
  void* dbranch; void* du; void* dxn; void* dctx; void* dqkv;  /* bf16 scratch: [M,D] [M,F] [M,D] [M,D] [M,3D] */
  float* delta;                        /* fp32 scratch [B*H*S]: attention backward row terms */
  float* gs[2];                        /* fp32 [B*S, D] scratch (stream gradient ping-pong) */
  /* parameter gradients: ACCUMULATED into — every buffer below must be zero on entry */
  float* const* dwqkv; float* const* dwo; float* const* dw1; float* const* db1;
  float* const* dw2; float* const* db2;
  float* const* dg1; float* const* dbe1; float* const* dg2; float* const* dbe2;
  /* layer range of this call: blocks l_end-1 ... l_begin (l_end = 0 means L). A caller may split the
   * backward into several calls, highest layers first, with the same scratch buffers (they carry
   * the state between calls) — e.g. to start reducing a chunk's gradients across ranks while the
   * next chunk is computed. */
  int64_t l_begin, l_end;
} vitssl_encoder_bwd_args;
int vitssl_encoder_stack_bwd(const vitssl_encoder_bwd_args* args, vitssl_stream_t stream);
/* Both stack calls replay a CUDA graph of their own launch sequence when they are called again with
 * identical arguments (every scalar and pointer; the dropout seed is exempt: it travels through a
 * device word) — the steady state of a training loop. A call whose arguments were not seen before
 * takes the direct path, so results never depend on the cache; VITSSL_GRAPH=0 disables it. The seed
 * word is set in stream order, so concurrent stack calls on DIFFERENT streams with dropout must switch
 * the replay off (vitssl_graph_enable). Counters
 * for tests and bench.py: graphs captured / calls served by a replay since the library was loaded. */
int vitssl_graph_stats(int64_t* captured, int64_t* replayed);
/* Switch the replay off (0) / on (1) at run time. The data-parallel module switches it off: buffers
 * that a collective touches on another stream are released at timing-dependent moments, so the
 * caller's addresses do not repeat from step to step there and every capture would be wasted. */
int vitssl_graph_enable(int on);

/* ---- weight casts, EMA, bias gradients (host_* arguments are HOST arrays of device pointers) */
/* fp32 -> bf16 for `count` tensors in one or a few launches (parameters stay fp32 nn.Parameters;
 * the GEMMs read bf16 shadows, like autocast's per-call weight casts). */
int vitssl_multi_cast_bf16(const void* const* host_src, void* const* host_dst,
                           const int64_t* host_numel, int count, vitssl_stream_t stream);
/* teacher <- m * teacher + (1 - m) * student over parameter lists (ssl/dino/model.py:126-139). */
int vitssl_multi_ema(void* const* host_teacher, const void* const* host_student,
                     const int64_t* host_numel, int count, float momentum, vitssl_stream_t stream);
/* same update with the teacher's bf16 GEMM-operand shadows written in the same pass
 * (host_shadow[i] NULL = parameter without a shadow), so the next teacher forward needs no cast. */
int vitssl_multi_ema_shadow(void* const* host_teacher, const void* const* host_student,
                            void* const* host_shadow, const int64_t* host_numel, int count,
                            float momentum, vitssl_stream_t stream);
/* out[c] = sum_r x[r,c] (bf16 in, fp32 out; out is zeroed on `stream`): bias gradients. */
int vitssl_colsum_bf16(const void* x, int64_t ld, int64_t rows, int64_t cols, float* out,
                       vitssl_stream_t stream);
/* out[c] += sum_r x[r,c] (out zeroed by the caller) */
int vitssl_colsum_bf16_acc(const void* x, int64_t ld, int64_t rows, int64_t cols, float* out,
                           vitssl_stream_t stream);

/* ---- patches and tokens ------------------------------------------------------------------ */
/* img fp32 [B,C,H,W] -> bf16 [B*(H/p)*(W/p), C*p*p], feature order (c,ph,pw), patches row-major
 * (nn.Unfold: ssl/simmim/model.py:43; Conv2d(k=s=p) im2col: patch_embedding.py:22,79-84). */
int vitssl_im2col_bf16(const float* img, void* out, int64_t B, int64_t C, int64_t H, int64_t W,
                       int64_t p, vitssl_stream_t stream);
/* raw fp32 patches of the listed flat patch ids (b*N + n): SimMIM targets (masking.py:35). */
int vitssl_gather_patches_f32(const float* img, const int32_t* rows_idx, float* out, int64_t n_rows,
                              int64_t C, int64_t H, int64_t W, int64_t p, vitssl_stream_t stream);
/* Input side (data/datasets.py:102-123 feed `ToTensor` output): the same two kernels
 * reading raw uint8 [B,C,H,W] image bytes, value = byte / 255 exactly as torchvision's ToTensor
 * computes it, so the host sends 1 byte per pixel instead of 4 and no fp32 image is ever stored. */
int vitssl_im2col_u8_bf16(const uint8_t* img, void* out, int64_t B, int64_t C, int64_t H, int64_t W,
                          int64_t p, vitssl_stream_t stream);
int vitssl_gather_patches_u8_f32(const uint8_t* img, const int32_t* rows_idx, float* out, int64_t n_rows,
                                 int64_t C, int64_t H, int64_t W, int64_t p, vitssl_stream_t stream);
/* Bicubic resize of the positional-embedding grid (patch_embedding.py:26-48) as a sparse row
 * interpolation: dst[i,:] = sum_t w[i,t] * src[idx[i,t],:], idx/w = [n_out, taps] tables (16 taps for
 * bicubic; built once per grid pair on the host with torch's upsample_bicubic2d index arithmetic).
 * Backward: dsrc (fp32 [n_in, D], zeroed on `stream`) += w[i,t] * ddst[i,:]. */
int vitssl_interp_rows_fwd(const float* src, const int32_t* idx, const float* w, float* dst,
                           int64_t n_out, int64_t D, int64_t taps, vitssl_stream_t stream);
int vitssl_interp_rows_bwd(const float* ddst, const int32_t* idx, const float* w, float* dsrc,
                           int64_t n_in, int64_t n_out, int64_t D, int64_t taps, vitssl_stream_t stream);
/* x[b,s,:] = (CLS | mask_token | proj[b,n,:]) + pos[s,:]   (patch_embedding.py:61-63,94-95;
 * ssl/simmim/model.py:47-49). cls NULL -> no CLS row (S = N); mask (uint8 [B*N]) NULL -> none. */
int vitssl_embed_tokens_fwd(const void* proj, const float* cls, const float* pos,
                            const uint8_t* mask, const float* mask_token, float* x, int64_t B,
                            int64_t N, int64_t D, vitssl_stream_t stream);
/* backward: dproj (bf16, zero on masked rows), dpos [S,D] (row 0 is also d(cls) when has_cls),
 * dmask_token [D] (nullable). dx is addressed with element strides (ld_b, ld_s). */
int vitssl_embed_tokens_bwd(const float* dx, int64_t ld_b, int64_t ld_s, const uint8_t* mask,
                            void* dproj, float* dpos, float* dmask_token, int64_t B, int64_t N,
                            int64_t D, int has_cls, vitssl_stream_t stream);
/* out[i,:] = bf16(x[idx[i],:]) — sync-free form of x[bool_mask] (ssl/simmim/model.py:56). */
int vitssl_gather_rows_bf16(const float* x, int64_t ldx, const int32_t* idx, void* out,
                            int64_t n_rows, int64_t D, vitssl_stream_t stream);
/* its backward: dx[r,:] = inv_idx[r] >= 0 ? dy[inv_idx[r],:] : 0 for every row r. */
int vitssl_scatter_rows_f32(const void* dy, const int32_t* inv_idx, float* dx, int64_t rows,
                            int64_t D, vitssl_stream_t stream);

/* ---- SimMIM masking (ssl/simmim/masking.py:6-37) ------------------------------------------ */
/* Bit-exact replay of B sequential `torch.randperm(N, device=cuda)[:n_keep]` draws from the
 * Philox state (seed, offset) of torch's CUDA generator, fused with the mask bookkeeping:
 *   perm_out  int64 [B, n_keep] (nullable)  the drawn indices, in draw order (masking.py:22-25)
 *   bool_mask uint8 [B, N]                   1 where masked (masking.py:27-33)
 *   rows      int32 [B*n_keep]               flat ids b*N+n of masked patches, ascending — the
 *                                            row order of `patches[bool_mask]` (masking.py:35)
 *   inv       int32 [B*N]                    position of a patch in `rows`, or -1
 * N <= 1024. The caller must advance the generator offset by
 * B * vitssl_randperm_offset_per_call(N) so later draws continue the reference's stream. */
int vitssl_simmim_mask(int64_t* perm_out, uint8_t* bool_mask, int32_t* rows, int32_t* inv,
                       int64_t B, int64_t N, int64_t n_keep, uint64_t philox_seed,
                       uint64_t philox_offset, vitssl_stream_t stream);
/* key width torch.randperm uses for n elements, and the generator offset one call consumes */
int vitssl_randperm_bits(int64_t n);
int64_t vitssl_randperm_offset_per_call(int64_t n);

/* ---- SimMIM objective --------------------------------------------------------------------- */
/* loss[0] = mean |pred - target| (nn.L1Loss(mean): utils/train_utils.py:19-22); sign (bf16,
 * nullable) receives sign(pred - target) so that d(pred) = sign * grad / n needs no second read
 * of the targets. pred bf16, target fp32. */
int vitssl_l1_loss_fwd(const void* pred, const float* target, void* sign, float* loss, int64_t n,
                       vitssl_stream_t stream);

/* d(pred) = sign * grad_out / n (bf16); grad_out is a DEVICE scalar (GradScaler-scaled upstream
 * gradient), so backward needs no host read and no re-read of pred/target. */
int vitssl_l1_loss_bwd(const void* sign, const float* grad_out, void* dpred, int64_t n,
                       vitssl_stream_t stream);

/* ---- optimizer step ------------------------------------------------------------
 * Fused multi-tensor AdamW replacing `scaler.step(optimizer)` on torch.optim.AdamW
 * (utils/train_utils.py:25-29, utils/trainers/simmim_trainer.py:69-71, base_trainer.py:44): one pass
 * over fp32 (param, grad, exp_avg, exp_avg_sq) with the GradScaler unscale (grad / *grad_scale) and
 * skip (*found_inf != 0) folded in, torch's fused-AdamW arithmetic, and the bf16 GEMM-operand shadow
 * written in the same pass (host_shadow[i] may be NULL). host_* are HOST arrays of `count` DEVICE
 * pointers; host_step[i] is the tensor's fp32 device step counter (advanced here unless found_inf). */
int vitssl_adamw_step(void* const* host_param, const void* const* host_grad, void* const* host_exp_avg,
                      void* const* host_exp_avg_sq, void* const* host_shadow, void* const* host_step,
                      const int64_t* host_numel, int count, float lr, float beta1, float beta2, float eps,
                      float weight_decay, const float* grad_scale, const float* found_inf,
                      vitssl_stream_t stream);

/* ---- DINO objective ------------------------------------------------------------------------ */
/* F.normalize(dim=1) (ssl/dino/head.py:21), bf16 rows. */
int vitssl_l2norm_fwd(const void* x, void* y, float* inv_norm, int64_t rows, int64_t D,
                      vitssl_stream_t stream);
int vitssl_l2norm_bwd(const void* x, const float* inv_norm, const void* dy, void* dx, int64_t rows,
                      int64_t D, vitssl_stream_t stream);
/* weight_norm(Linear) (head.py:17): w = g * v / ||v||_row -> bf16; backward maps dW to (dg, dv). */
int vitssl_weight_norm_fwd(const float* v, const float* g, void* w, float* inv_norm, int64_t rows,
                           int64_t D, vitssl_stream_t stream);
int vitssl_weight_norm_bwd(const float* dw, const float* v, const float* g, const float* inv_norm,
                           float* dg, float* dv, int64_t rows, int64_t D, vitssl_stream_t stream);
/* out = m * center + (1 - m) * colsum * inv_rows (ssl/dino/model.py:96-99); colsum comes from
 * vitssl_colsum_bf16 over the teacher logits (all-reduced across ranks first under DP). */
int vitssl_center_ema(const float* center, const float* colsum, float* out, int64_t K, float momentum,
                      float inv_rows, vitssl_stream_t stream);
/* DINOLoss (ssl/dino/loss.py:13-29) for teacher bf16 [G,B,K], student bf16 [V,B,K], center [K]:
 * loss[0] = -(1/(G B K)) sum_b sum_k (sum_g softmax((T_g-c)/tt))(sum_v log_softmax(S_v/ts)).
 * t_stats [G,B,2] and s_lse [V,B] are saved for backward. G <= 4, V <= 12 (callers chunk more views: the loss is additive over view groups), K % 8 == 0. */
int vitssl_dino_loss_fwd(const void* teacher, const void* student, const float* center, float* loss,
                         float* t_stats, float* s_lse, int64_t G, int64_t V, int64_t B, int64_t K,
                         float teacher_temp, float student_temp, vitssl_stream_t stream);
/* dstudent[v,b,k] = -(grad_out/(ts G B K)) (Pbar[b,k] - G softmax(S_v/ts)[k]); grad_out is a
 * DEVICE scalar (the GradScaler-scaled upstream gradient) so no host sync is needed. */
int vitssl_dino_loss_bwd(const void* teacher, const void* student, const float* center,
                         const float* t_stats, const float* s_lse, const float* grad_out,
                         void* dstudent, int64_t G, int64_t V, int64_t B, int64_t K,
                         float teacher_temp, float student_temp, vitssl_stream_t stream);

/* ---- evaluation path ---------------------------------------------------------- */
/* out[b,:] = mean_s x[b,s,:] — the token mean-pool of SimMIMViT.inference_forward
 * (ssl/simmim/model.py:91-93). fp32 [B,S,D] dense, D % 4 == 0. */
int vitssl_mean_tokens_f32(const float* x, float* out, int64_t B, int64_t S, int64_t D,
                           vitssl_stream_t stream);
/* Cosine k-nearest-neighbour classification of fp32 feature rows — what the reference's evaluator
 * does with scikit-learn on the CPU (evaluators/unsupervised_evaluator.py:38-66:
 * KNeighborsClassifier(n_neighbors=num_classes, metric="cosine").fit(train).predict(val)): rows are
 * L2-normalised, the k most similar training rows are taken (smaller index first on exact ties) and
 * vote uniformly (smallest label on ties). Caller-provided workspaces: val_n [Nv,D], train_n [Nt,D],
 * sims [Nv,Nt] fp32 (consumed). pred int32 [Nv]; neighbors int32 [Nv,k] nullable. */
int vitssl_knn_cosine(const float* val, const float* train, const int32_t* train_labels, float* val_n,
                      float* train_n, float* sims, int32_t* pred, int32_t* neighbors, int64_t Nv,
                      int64_t Nt, int64_t D, int64_t k, int64_t num_classes, vitssl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VITSSL_B200_H_ */
