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

/* element types of the solver / latent tensors (the reference hard-codes fp16 at its cast sites,
 * models.py:58, predictor.py:85, main.py:138-139,443; north_star asks for bf16) */
#define TDM_F32 0
#define TDM_F16 1
#define TDM_BF16 2

#define TDM_OK 0
#define TDM_ERR_INVALID (-1) /* shape / alignment / argument violation */
#define TDM_ERR_CUDA (-2)    /* a CUDA runtime or driver call failed */
#define TDM_ERR_STATE (-3)   /* object used in the wrong state */

const char* tdm_last_error(void); /* thread-local, never NULL */
int tdm_version(void);            /* ABI version, bumped on any signature change */
int tdm_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* number of CUDA kernels this library has launched in this process (bench.py's gpu_launches); reset != 0 zeroes it */
uint64_t tdm_launch_count(int reset);

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
 *   v = alpha*acc + bias[n]; pre_out[m,n] = bf16(v); pre_out_f32[m,n] = v; act;
 *   v *= gate[m / rows_per_gate, n]; v += res[m (% res_row_mod), n]; out_f32[m,n] = v; out_bf16[m,n] = bf16(v)
 * Every pointer is optional except A, B and at least one output.  N, K, K2 and all ld* must be
 * multiples of 8 (16-byte TMA strides / vector stores). */
#define TDM_ACT_NONE 0
#define TDM_ACT_GELU_TANH 1     /* v = gelu_tanh(v) */
#define TDM_ACT_GELU_TANH_BWD 2 /* v = v * gelu_tanh'(act_in[m,n]) */
#define TDM_ACT_SILU 3          /* v = v * sigmoid(v) */
#define TDM_ACT_SILU_BWD 4      /* v = v * silu'(act_in[m,n]) */

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
  float* pre_out_f32; /* fp32 [M, ld_pre_f32]: same value as pre_out, unrounded */
  int64_t ld_pre_f32;
  const float* gate; /* fp32 [ceil(M/rows_per_gate), gate_stride] */
  int32_t rows_per_gate;
  int64_t gate_stride;
  const float* res; /* fp32 [M, ld_res] (may alias out_f32) */
  int64_t ld_res;
  int32_t res_row_mod; /* if >0 the residual row is m % res_row_mod (pos_embed broadcast over the batch) */
  float* out_f32;
  int64_t ld_out_f32;
  void* out_bf16;
  int64_t ld_out_bf16;
  int32_t tile_n;  /* 0 = choose; else 64/128/192/256 */
  int32_t split_k; /* 0 = choose; >1 forces split-K (accumulate-only epilogues: res == out_f32, nothing else) */
  int32_t cta_pair; /* 0 = choose; 1 = one CTA per 128-row tile; 2 = CTA pair (cta_group::2) per 256-row tile */
  /* Stacked problems that share one launch (single-CTA kernel, no second operand pair; seg_k needs K % 64 == 0):
   *  seg_k  > 0 (with seg_n): output columns of segment s = n / seg_n contract over A columns / B rows
   *            [s*seg_k, s*seg_k + K) and read B columns n - s*seg_n  (LoRA dt = dy_s B_s for stacked q/k/v)
   *  seg_m  > 0: output rows of segment s = m / seg_m read B columns n + s*seg_nb          (LoRA dB_s = dy_s^T t_s) */
  int32_t seg_k, seg_m, seg_nb;
  int32_t epilogue; /* 0 = choose; 1 = generic epilogue (direct loads / stores); 2 = streaming TMA epilogue with the
                       run-time tested code; 3 = streaming, feature-specialised instance when one exists (same results,
                       bit for bit: tests/test_gemm_gpu.py) */
} TdmGemm;

int tdm_gemm_bf16(const TdmGemm* g, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention on tcgen05/TMEM: O = softmax(Q K^T * scale + bias[b, key]) V  per (batch, head).
 * Replaces xformers.memory_efficient_attention, which the reference enables unconditionally
 * (main.py:177-178, 372), for diffusers' attn1 (self) and attn2 (cross, T5 context with the additive
 * mask bias (1-mask)*-10000 broadcast over heads and queries).  head_dim must be a multiple of 8 in
 * [16,128] (PixArt: 72).  Q/K/V/O/dO/dQ/dK/dV are bf16 [B, S, H, d] VIEWS with d contiguous, given as a
 * base pointer + element strides, so the fused QKV GEMM output [tokens, 3*H*d] is consumed in place.
 * Backward needs lse from the forward, the forward output O, and two fp32 scratch buffers. */
typedef struct TdmAttn {
  int32_t B, H, Sq, Sk, d;
  float scale;
  const void* Q;
  int64_t q_stride_b, q_stride_s, q_stride_h;
  const void* K;
  int64_t k_stride_b, k_stride_s, k_stride_h;
  const void* V;
  int64_t v_stride_b, v_stride_s, v_stride_h;
  const float* bias; /* fp32 [B, Sk] additive key bias, or NULL */
  void* O;           /* forward: output; backward: the saved forward output */
  int64_t o_stride_b, o_stride_s, o_stride_h;
  float* lse; /* fp32 [B, H, Sq] natural-log row sum-exp: written by fwd (may be NULL there), read by bwd */
  /* backward only */
  const void* dO; /* same strides as O */
  float* delta;    /* fp32 [B, H, Sq] scratch */
  float* dq_accum; /* fp32 [B*Sq, H*d] scratch (zero-filled by the call) */
  void* dQ;        /* must be a token-matrix view: dq_stride_h == d, dq_stride_b == Sq*dq_stride_s */
  int64_t dq_stride_b, dq_stride_s, dq_stride_h;
  void* dK;
  int64_t dk_stride_b, dk_stride_s, dk_stride_h;
  void* dV;
  int64_t dv_stride_b, dv_stride_s, dv_stride_h;
} TdmAttn;

int tdm_attn_fwd(const TdmAttn* a, void* stream);
int tdm_attn_bwd(const TdmAttn* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Few-step solver / re-noising / loss: the reference-owned elementwise math, fused.
 * Given identical inputs these are bit-identical to the reference's eager op sequence. */

/* One solver step: eps-half select (models.py:40) + CFG combine (models.py:45, predictor.py:42) +
 * predicted_origin (utils.py:47-59) + noise_scheduler.add_noise(x0, eps, t_next).to(lowp)
 * (models.py:58).  model_out is [B,Cout,HW] fp32 and eps is its first C channels; x_t is [B,C,HW].
 * prediction_type: 0 "epsilon", 1 "v_prediction"; anything else is the reference's ValueError.
 * Negative timesteps index from the end of the table like torch does (t=-1 after the last step). */
int tdm_solver_step(const float* model_out, const float* model_out_uncond, double cfg, const void* x_t, int x_dtype,
                    const int64_t* t, const int64_t* t_next, const void* alphas, const void* sigmas, int table_dtype,
                    int table_len, const float* alphas_cumprod, int prediction_type, float* eps_out, float* x0_out,
                    void* x_next, int next_dtype, int B, int C, int Cout, int HW, void* stream);

/* Predictor.add_noise (predictor.py:76-85; mixed=0) / Predictor.obtain_mixed_noise (predictor.py:87-97;
 * mixed=1), optionally fused with the per-sample trajectory select of main.py:479
 * (`noisy_imgs_list[ind_t]` made a gather): sample b reads states[index[b]].
 * Every op is rounded in the dtype torch's promotion rules give it.  t2 < t1 yields NaN as in the
 * reference (D3). */
int tdm_renoise(const void* const* states, const int* state_dtypes, int n_states, const int64_t* index,
                const void* noise, int noise_dtype, const int64_t* t1, const int64_t* t2, const void* alphas,
                const void* sigmas, int table_dtype, int table_len, void* out, int out_dtype, int mixed, int B,
                long per_sample, void* stream);

/* TDM loss (main.py:514-529) and its closed-form gradient d loss / d model_latents (times grad_scale).
 * scratch_w: [B] doubles, loss_out: [1] double (both overwritten). grad_out may be NULL. */
int tdm_loss_fwd_bwd(const float* model_latents, const float* sd_latents, const float* sd_latents_uncond, double cfg,
                     int use_huber, double huber_c, double* scratch_w, double* loss_out, float* grad_out,
                     float grad_scale, int B, long per_sample, void* stream);

/* ------------------------------------------------------------------------------------------------
 * HBM-bound block kernels (diffusers BasicTransformerBlock with norm_type="ada_norm_single";
 * SURVEY.md Appendix A, K1/K9/K12/K13/K15/K21). */

/* h = LayerNorm(x; eps, no affine) * (1 + scale[b]) + shift[b];  b = row / rows_per_sample;
 * shift/scale point into a [B, ..] fp32 tensor with mod_stride elements between samples. */
int tdm_ln_modulate_fwd(const float* x, const float* shift, const float* scale, int64_t mod_stride,
                        int rows_per_sample, void* out_bf16, float* mean, float* rstd, int M, int D, float eps,
                        void* stream);
/* dx_out = dx_in + LN-modulate backward of dh; optional bf16 copy; optional accumulation of
 * d shift / d scale into [B, dmod_stride]-strided fp32 buffers. */
int tdm_ln_modulate_bwd(const void* dh_bf16, const float* x, const float* mean, const float* rstd, const float* scale,
                        int64_t mod_stride, int rows_per_sample, const float* dx_in, float* dx_out, void* dx_out_bf16,
                        float* dshift, float* dscale, int64_t dmod_stride, int M, int D, void* stream);
/* backward of x_out = x + gate[b] * y:  dy = bf16(gate[b] * dx);  dgate[b] += sum_rows dx * y (optional) */
int tdm_gate_bwd(const float* dx, const void* y_bf16, const float* gate, int64_t gate_stride, int rows_per_sample,
                 void* dy_bf16, float* dgate, int64_t dgate_stride, int M, int D, void* stream);
/* out[n] += sum_m a[m,n]  (bias gradients) */
int tdm_colsum_bf16(const void* a_bf16, int64_t ld, int M, int N, float* out, void* stream);
/* mod[b,j,:] = table[j,:] + (per_j ? tvec[b, j*D:(j+1)*D] : tvec[b,:]) and its backward */
int tdm_modulation(const float* table, const float* tvec, int per_j, float* out, int B, int J, int D, void* stream);
int tdm_modulation_bwd(const float* dmod, float* dtable, float* dtvec, int per_j, int accumulate_tvec, int B, int J,
                       int D, void* stream);
/* Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0) -> bf16 [B, dim] */
int tdm_timestep_sinusoid(const int64_t* t, void* out_bf16, int B, int dim, void* stream);
/* PatchEmbed im2col: latent [B,C,H,W] (dtype enum) -> bf16 [B*(H/p)*(W/p), C*p*p] */
int tdm_patchify(const void* latent, int dtype, void* out_bf16, int B, int C, int H, int W, int p, void* stream);
/* proj_out rows -> NCHW ("nhwpqc->nchpwq") and the matching gradient scatter (first Cgrad channels) */
int tdm_unpatchify(const float* y, int64_t ldy, float* out, int B, int Cout, int H, int W, int p, void* stream);
int tdm_unpatchify_bwd(const float* dout, void* dy_bf16, int B, int Cgrad, int Cout, int H, int W, int p, void* stream);
int tdm_cast_f32_bf16(const float* in, void* out_bf16, int64_t n, void* stream);
/* DDPMScheduler.add_noise (third-party diffusers; reference call sites models.py:58, predictor.py:58,70):
 * alphas_cumprod is cast to x0's dtype, a = ac[t]**0.5, s = (1-ac[t])**0.5, out = a*x0 + s*noise with torch's
 * type promotion; out_dtype must be that promoted type. */
int tdm_scheduler_add_noise(const void* x0, int x0_dtype, const void* noise, int noise_dtype, const int64_t* t,
                            const float* alphas_cumprod, int table_len, void* out, int out_dtype, int B,
                            long per_sample, void* stream);
/* backward of tdm_solver_step w.r.t. the model outputs (epsilon prediction):
 *   d out_c[:, :C] = cfg' * (g_eps - sigma/alpha * g_x0),  d out_u[:, :C] = (1 - cfg) * (same), other channels 0,
 *   with cfg' = cfg when an unconditional output took part, else 1.  g_eps / g_x0 may be NULL. */
int tdm_solver_step_bwd(const float* g_eps, const float* g_x0, double cfg, int has_uncond, const int64_t* t,
                        const void* alphas, const void* sigmas, int table_dtype, int table_len, float* d_out_c,
                        float* d_out_u, int B, int C, int Cout, int HW, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training-step glue (src/main.py:473-479, 535-540).  Device-side scalars only: nothing here syncs. */

/* out[b] = states[index[b]][b]: `noisy_imgs_list[ind_t]` (main.py:479) as the per-sample gather it was
 * meant to be; negative indices count from the end like Python's. */
int tdm_select_states(const void* const* states, const int* state_dtypes, int n_states, const int64_t* index, void* out,
                      int out_dtype, int B, long per_sample, void* stream);
/* Fake-score denoising loss (repair of SURVEY.md D5; the reference never trains unet_fake, main.py:470-505):
 * loss = mean((pred[:, :C] - target)^2) in fp64, grad_out[B,C,HW] = 2 (pred - target) / numel * grad_scale. */
int tdm_mse_fwd_bwd(const float* pred, const float* target, double* loss_out, float* grad_out, float grad_scale, int B,
                    int C, int Cout, int HW, void* stream);
/* accelerator.clip_grad_norm_ (main.py:537) on flat gradient arenas: accumulate sum of squares, then
 * coef = min(1, max_norm / (norm + 1e-6)) * inv_scale with norm = sqrt(sumsq) * inv_scale. */
int tdm_sumsq_f32(const float* x, int64_t n, double* accum, int zero_first, void* stream);
int tdm_clip_coef(const double* sumsq, float max_norm, float inv_scale, float* coef_out, float* norm_out, void* stream);
/* optimizer.step() (main.py:206-224, 538): torch.optim.AdamW semantics on a flat fp32 arena; the gradient is
 * multiplied by *grad_coef (device scalar, may be NULL); param_bf16 (optional) receives the re-cast weights.
 * Hyper-parameters are doubles: 1 - beta, 1 - lr * wd, lr / (1 - beta1^t) are formed in double like torch does. */
int tdm_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16, int64_t n,
                   double lr, double beta1, double beta2, double eps, double weight_decay, int step,
                   const float* grad_coef, void* stream);

/* EMAModel.step (diffusers; reference main.py:188-189, 553-554) on a flat fp32 arena:
 * shadow -= (1 - decay) * (shadow - param). */
int tdm_ema_update(float* shadow, const float* param, int64_t n, float decay, void* stream);

/* ------------------------------------------------------------------------------------------------
 * One PixArt DiT block (diffusers BasicTransformerBlock, norm_type="ada_norm_single"; SURVEY.md
 * Appendix A) forward and backward as ONE call each: the library sequences its own kernels
 * (LN-modulate, tcgen05 GEMMs with fused bias/GELU/gate/residual/LoRA epilogues, tcgen05 attention)
 * on the caller's stream.  All buffers are caller-owned.
 *   x       fp32 [B*N, D] residual stream (in) -> x_out fp32 [B*N, D]
 *   tmod    fp32 [B, 6*D]  output of adaln_single.linear for this batch's timesteps
 *   ctx     bf16 [B*L, D]  projected caption tokens; ctx_bias fp32 [B, L] = (1-mask)*-10000
 * Weights are bf16 [out, in] (q/k/v and k2/v2 stacked on the output dimension), biases fp32.
 * LoRA (peft semantics y = Wx + b + (alpha/r) B(Ax)): lora_a_* bf16 [r(,x3/x2), in], lora_b_* bf16 [out, r],
 * lora_r must be 64 when present; a NULL lora_a_* pointer disables LoRA for that linear.
 * `save` != 0 keeps what backward needs in `ws` (tdm_dit_block_ws_bytes(cfg, 1) bytes, 256-byte
 * aligned); with save == 0 `ws` is scratch of tdm_dit_block_ws_bytes(cfg, 0) bytes. */
typedef struct TdmDitShape {
  int32_t B, N, D, H, L, Dff;
  int32_t lora_r;     /* 0 or 64 */
  float lora_scale;   /* alpha / r */
  float eps;
  int32_t train_base; /* backward also produces base weight/bias/table gradients (full fine-tune) */
} TdmDitShape;

typedef struct TdmBlockWeights {
  const float* table; /* scale_shift_table [6, D] fp32 */
  const void* w_qkv; const float* b_qkv; /* [3D, D], [3D] */
  const void* w_o;   const float* b_o;   /* attn1.to_out.0 */
  const void* w_q2;  const float* b_q2;  /* attn2.to_q */
  const void* w_kv2; const float* b_kv2; /* attn2.to_k/to_v stacked [2D, D] */
  const void* w_o2;  const float* b_o2;
  const void* w_ff1; const float* b_ff1; /* ff.net.0.proj [Dff, D] */
  const void* w_ff2; const float* b_ff2; /* ff.net.2 [D, Dff] */
  const void* lora_a_qkv; const void* lora_b_qkv; /* [3r, D], [3D, r] */
  const void* lora_a_o;   const void* lora_b_o;
  const void* lora_a_q2;  const void* lora_b_q2;
  const void* lora_a_kv2; const void* lora_b_kv2; /* [2r, D], [2D, r] */
  const void* lora_a_o2;  const void* lora_b_o2;
  const void* lora_a_ff1; const void* lora_b_ff1;
  const void* lora_a_ff2; const void* lora_b_ff2;
} TdmBlockWeights;

/* fp32 gradient buffers mirroring TdmBlockWeights; every pointer optional; gradients are ACCUMULATED (+=). */
typedef struct TdmBlockGrads {
  float* table;
  float* w_qkv; float* b_qkv;
  float* w_o;   float* b_o;
  float* w_q2;  float* b_q2;
  float* w_kv2; float* b_kv2;
  float* w_o2;  float* b_o2;
  float* w_ff1; float* b_ff1;
  float* w_ff2; float* b_ff2;
  float* lora_a_qkv; float* lora_b_qkv;
  float* lora_a_o;   float* lora_b_o;
  float* lora_a_q2;  float* lora_b_q2;
  float* lora_a_kv2; float* lora_b_kv2;
  float* lora_a_o2;  float* lora_b_o2;
  float* lora_a_ff1; float* lora_b_ff1;
  float* lora_a_ff2; float* lora_b_ff2;
  float* d_tmod; /* [B, 6D] accumulated gradient of tmod (needed when train_base) */
  float* d_ctx;  /* [B*L, D] fp32 accumulated gradient of ctx (needed when train_base) */
} TdmBlockGrads;

int64_t tdm_dit_block_ws_bytes(const TdmDitShape* s, int save);
int64_t tdm_dit_block_bwd_scratch_bytes(const TdmDitShape* s);
int tdm_dit_block_fwd(const TdmDitShape* s, const TdmBlockWeights* w, const float* x, const float* tmod,
                      const void* ctx_bf16, const float* ctx_bias, float* x_out, void* ws, int save, void* stream);
/* dx_out fp32 [B*N, D] = gradient w.r.t. the block input, given dx fp32 [B*N, D] w.r.t. its output.
 * `x` is the same block input that was given to the forward call; `ws` its saved workspace. */
int tdm_dit_block_bwd(const TdmDitShape* s, const TdmBlockWeights* w, const TdmBlockGrads* g, const float* x,
                      const void* ctx_bf16, const float* ctx_bias, const float* dx, float* dx_out, const void* ws,
                      void* scratch, void* stream);

/* LoRA merge for the no-grad forwards (peft semantics y = W x + (alpha / r) B A x): for every entry of a
 * DEVICE-resident problem table, w_eff = bf16(w32 + scale * lora_b lora_a) with r = 64, one launch for the whole model.
 * w32 / w_eff: [n_out, k_in] row-major (fp32 master / bf16 copy), lora_a: bf16 [64, k_in], lora_b: bf16 [n_out, 64];
 * n_out and k_in multiples of 128; tile_begin = running count of 128 x 128 tiles before this problem. */
typedef struct TdmMergeProblem {
  const float* w32;
  void* w_eff;
  const void* lora_a;
  const void* lora_b;
  int32_t n_out, k_in;
  int32_t tile_begin, reserved;
} TdmMergeProblem;
int tdm_lora_merge(const TdmMergeProblem* problems_dev, int n_problems, int total_tiles, float scale, void* stream);

/* -------------------------------------------------------------------------------------------------
 * TAESD decoder: diffusers.AutoencoderTiny.decode (reference main.py:166, 338; training.py:113-117), SURVEY.md 8f rank 2.
 *   x = tanh(x / 3) * 3 ; conv3x3(4 -> 64) + ReLU ; [3 blocks ; nearest up x2 ; conv3x3(64 -> 64, no bias)] x 3 ; 1 block ;
 *   conv3x3(64 -> 3) ; x * 2 - 1,      block(x) = ReLU(conv(ReLU(conv(ReLU(conv(x))))) + x)
 * The 64 -> 64 convolutions run as tcgen05 implicit GEMMs over NHWC bf16 activations (TMA zero fill == padding).
 * conv_in_w [64,4,3,3], conv_out_w [3,64,3,3] and all biases: fp32 in PyTorch layout.  block_w[i] / up_w[i]: one
 * 64 -> 64 convolution packed as bf16 [9 taps][64 out][64 in] (tap = ky * 3 + kx); block_w holds the 10 blocks'
 * three convolutions in order (block_b their biases), up_w the convolution that follows each upsample. */
typedef struct TdmTaesdWeights {
  const float* conv_in_w;  const float* conv_in_b;
  const void* block_w[30]; const float* block_b[30];
  const void* up_w[3];
  const float* conv_out_w; const float* conv_out_b;
} TdmTaesdWeights;
/* bytes of workspace for B latents of H x W (three NHWC bf16 activation buffers at the output resolution) */
int64_t tdm_taesd_ws_bytes(int B, int H, int W);
/* latents [B,4,H,W] NCHW (dtype TDM_F32 / TDM_F16 / TDM_BF16) -> out fp32 [B,3,8H,8W] NCHW.  to_image != 0 applies the
 * reference's post-processing clamp(-1, 1) * 0.5 + 0.5 (training.py:113-117) in the last kernel. */
int tdm_taesd_decode(const TdmTaesdWeights* w, const void* latents, int latent_dtype, int B, int H, int W, float* out,
                     int to_image, void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TDM_B200_H */
