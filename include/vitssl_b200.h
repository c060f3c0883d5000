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
/* number of kernels this library has launched on the calling thread since the last reset
 * (bench.py reports it as gpu_launches) */
int64_t vitssl_launch_count(int reset);

/* ---- GEMM: every nn.Linear / Conv2d-patchify on the path -------------------------------
 * C[M,N] = alpha * op(A) * op(B), bf16 inputs, fp32 accumulation (tcgen05.mma, TMEM).
 *   a_mn = 0: A stored [M][K] (pitch lda)      a_mn = 1: A stored [K][M]
 *   b_mn = 0: B stored [N][K] (nn.Linear weight) b_mn = 1: B stored [K][N]
 * forward  y = x W^T   : a_mn=0,b_mn=0 (attention.py:82-84,105; feed_forward.py:26,28)
 * dgrad    dx = dy W   : a_mn=0,b_mn=1        wgrad dW = dy^T x : a_mn=1,b_mn=1
 * epilogue: NONE | BIAS (+bias[N]) | BIAS_GELU (aux <- bf16(acc+bias), C <- dropout(gelu(aux));
 *           feed_forward.py:26-27) | DGELU (C <- acc * mask/(1-p) * gelu'(aux)).
 * out_fp32: C is fp32, else bf16. split_k: 0 = off, -1 = auto, n = n splits (fp32 NONE only;
 * C is zeroed on `stream` and accumulated with red.add). */
#define VITSSL_EPI_NONE 0
#define VITSSL_EPI_BIAS 1
#define VITSSL_EPI_BIAS_GELU 2
#define VITSSL_EPI_DGELU 3
int vitssl_gemm_bf16(const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K,
                     int64_t lda, int64_t ldb, int64_t ldc, int a_mn, int b_mn, int epilogue,
                     const float* bias, void* aux, int64_t ld_aux, float alpha, int out_fp32,
                     int split_k, float dropout_p, uint64_t philox_seed, uint64_t philox_offset,
                     vitssl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VITSSL_B200_H_ */
