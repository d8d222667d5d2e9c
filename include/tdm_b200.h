/* libtdm_b200 -- C ABI of the B200-native TDM distillation hot path.
 *
 * The reference (nauka117/TDM) has no FFI layer: its seam is the Python duck-typed surface between
 * its solver/loss glue (src/models.py, src/predictor.py, src/utils.py, src/main.py:441-540) and the
 * third-party diffusers PixArt transformer it calls (main.py:168-176, models.py:38-39,
 * predictor.py:33-39).  This header is the boundary a maintainer binds instead (ctypes stub in
 * INTEGRATION.md): every entry point takes plain device pointers, sizes and a CUDA stream
 * (`void* stream` == cudaStream_t), never allocates tensor memory, enqueues asynchronously on the
 * caller's stream, returns 0 on success or a negative TDM_ERR_* code, and leaves a message in
 * tdm_last_error().  No C++ exception crosses the ABI.  There is no CPU fallback.
 *
 * Conventions: bf16 = 2-byte bfloat16 storage; "ld*" are leading dimensions in ELEMENTS;
 * latents are NCHW; token matrices are row-major [rows, features].
 */
#ifndef TDM_B200_H
#define TDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDM_OK 0
#define TDM_ERR_INVALID (-1) /* shape / alignment / argument violation */
#define TDM_ERR_CUDA (-2)    /* a CUDA runtime or driver call failed */
#define TDM_ERR_STATE (-3)   /* object used in the wrong state */

const char* tdm_last_error(void); /* thread-local, never NULL */
int tdm_version(void);            /* ABI version, bumped on any signature change */
int tdm_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * Dense contraction on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA-fed).
 *   C[M,N] = alpha * ( A[M,K] . B[N,K]^T  +  A2[M,K2] . B2[N,K2]^T )   then the fused epilogue.
 * Replaces the nn.Linear calls inside diffusers' BasicTransformerBlock that the reference triggers
 * at models.py:38-39 / predictor.py:33-39 (SURVEY.md K2,K4,K5,K6,K8,K10,K11) and their autograd
 * backward (K21).  The second operand pair is the LoRA side product folded into the same K loop.
 *
 * Operand storage: a_mn_major=0 -> A stored [M,K] (K contiguous), lda = row stride;
 *                  a_mn_major=1 -> A stored [K,M] (M contiguous), lda = row stride.   Same for B.
 *   forward  y  = x W^T      : A=x (0), B=W[N,K] (0)
 *   dgrad    dx = dy W       : A=dy (0), B=W[N_out,K_in] read as [K=N_out rows, N=K_in] (1)
 *   wgrad    dW = dy^T x     : A=dy [tokens, N_out] (1), B=x [tokens, K_in] (1)
 * Epilogue, per element (m,n), in this order:
 *   v = alpha*acc + bias[n]; pre_out[m,n] = bf16(v); act; v *= gate[m / rows_per_gate, n];
 *   v += res[m,n]; out_f32[m,n] = v; out_bf16[m,n] = bf16(v)
 * Every pointer is optional except A, B and at least one output.  N, K, K2 and all ld* must be
 * multiples of 8 (16-byte TMA strides / vector stores). */
#define TDM_ACT_NONE 0
#define TDM_ACT_GELU_TANH 1     /* v = gelu_tanh(v) */
#define TDM_ACT_GELU_TANH_BWD 2 /* v = v * gelu_tanh'(act_in[m,n]) */

typedef struct TdmGemm {
  int32_t M, N, K;
  int32_t a_mn_major, b_mn_major;
  const void* A; /* bf16 */
  int64_t lda;
  const void* B; /* bf16 */
  int64_t ldb;
  /* optional second pair (same majors as the first) */
  int32_t K2;
  int32_t seg_n;  /* if >0: output columns are split into segments of seg_n; segment s reads */
  int32_t seg_k2; /*        A2 columns [s*seg_k2, s*seg_k2+K2)  (fused q/k/v LoRA)             */
  const void* A2;
  int64_t lda2;
  const void* B2;
  int64_t ldb2;
  /* epilogue */
  float alpha;
  int32_t act;
  const float* bias;  /* [N] */
  const void* act_in; /* bf16 [M, ld_act] */
  int64_t ld_act;
  void* pre_out; /* bf16 [M, ld_pre] */
  int64_t ld_pre;
  const float* gate; /* fp32 [ceil(M/rows_per_gate), gate_stride] */
  int32_t rows_per_gate;
  int64_t gate_stride;
  const float* res; /* fp32 [M, ld_res] (may alias out_f32) */
  int64_t ld_res;
  float* out_f32;
  int64_t ld_out_f32;
  void* out_bf16;
  int64_t ld_out_bf16;
  int32_t tile_n; /* 0 = choose; else 64/128/256 */
} TdmGemm;

int tdm_gemm_bf16(const TdmGemm* g, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TDM_B200_H */
