// Training-step glue kernels around the DiT passes: per-sample trajectory select (main.py:479 made a
// gather), the fake-score denoising loss with its closed-form gradient (the repair of
// SURVEY.md D5), and the step that follows the hot path -- global-norm clip (main.py:537) + AdamW
// (main.py:206-224, 538) over the flat parameter arenas, fused with the bf16 re-cast of the weights
// the tensor-core kernels read.  All HBM-bound: 128-bit vectorised, grid sized in SM multiples,
// warp-shuffle reductions, device-side scalars (no host sync anywhere in the step).
#include <cuda_fp16.h>
#include "common.cuh"

namespace tdm {

static inline int ts_grid(long work_items, int threads) {
  long g = (work_items + threads - 1) / threads;
  const long cap = (long)num_sms() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

__device__ __forceinline__ float ld_any(const void* p, long i, int dt) {
  if (dt == TDM_F32) return reinterpret_cast<const float*>(p)[i];
  if (dt == TDM_F16) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, long i, float v, int dt) {
  if (dt == TDM_F32) reinterpret_cast<float*>(p)[i] = v;
  else if (dt == TDM_F16) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

struct SelectArgs {
  const void* states[8];
  int state_dt[8];
  int n_states;
  const int64_t* index;
  void* out;
  int out_dt;
  long per, B;
};

__global__ void select_states_kernel(SelectArgs a) {
  const long total = a.B * a.per;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int b = int(i / a.per);
    int si = (int)a.index[b];
    if (si < 0) si += a.n_states;  // python indexing
    st_any(a.out, i, ld_any(a.states[si], i, a.state_dt[si]), a.out_dt);
  }
}

__device__ __forceinline__ double block_sum_d(double v) {
  __shared__ double sm[32];
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sm[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sm[threadIdx.x] : 0.0;
  if (w == 0) v = warp_sum(v);
  __syncthreads();
  return v;
}

// loss = mean((pred[:, :C] - target)^2) accumulated in fp64; grad = 2 (pred - target) / numel * grad_scale
__global__ void mse_fwd_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target, double* loss,
                                   float* __restrict__ grad, float grad_scale, int B, int C, int Cout, int HW) {
  const long per = (long)C * HW, per_out = (long)Cout * HW;
  const long total = (long)B * per;
  const double inv_n = 1.0 / (double)total;
  double acc = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long b = i / per;
    const float d = pred[b * per_out + (i - b * per)] - target[i];
    acc += (double)d * (double)d;
    if (grad) grad[i] = (float)(2.0 * (double)d * inv_n) * grad_scale;
  }
  acc = block_sum_d(acc);
  if (threadIdx.x == 0) atomicAdd(loss, acc * inv_n);
}

// sum of squares of a flat fp32 buffer, accumulated into *out (fp64)
__global__ void sumsq_kernel(const float* __restrict__ x, long n, double* out) {
  const long n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float acc = 0.f;
  double dacc = 0.0;
  int cnt = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 v = x4[i];
    acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    if (++cnt == 64) {  // bound the fp32 partial before folding it into the fp64 accumulator
      dacc += (double)acc;
      acc = 0.f;
      cnt = 0;
    }
  }
  dacc += (double)acc;
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long i = n4 << 2; i < n; ++i) dacc += (double)x[i] * (double)x[i];
  dacc = block_sum_d(dacc);
  if (threadIdx.x == 0) atomicAdd(out, dacc);
}

// coef = min(1, max_norm / (sqrt(sumsq) * inv_scale + 1e-6)) * inv_scale  (torch.nn.utils.clip_grad_norm_ with the
// gradients still carrying a 1/inv_scale factor, e.g. a sum over ranks); norm_out = sqrt(sumsq) * inv_scale
__global__ void clip_coef_kernel(const double* sumsq, float max_norm, float inv_scale, float* coef, float* norm_out) {
  const float norm = (float)sqrt(*sumsq) * inv_scale;
  float c = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.f;
  c = c < 1.f ? c : 1.f;
  *coef = c * inv_scale;
  if (norm_out) *norm_out = norm;
}

// AdamW (decoupled weight decay, bias-corrected; torch.optim.AdamW semantics) on a flat arena.
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, bf16* __restrict__ p16, long n, float b2, float omb1,
                             float omb2, float eps, float decay, float step, float bc2_sqrt,
                             const float* __restrict__ coef_ptr) {
  const float coef = coef_ptr ? *coef_ptr : 1.f;
  const long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  uint2* o4 = reinterpret_cast<uint2*>(p16);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    float* P = &pp.x;
    float* Mv = &mm.x;
    float* V = &vv.x;
    const float* G = &gg.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = G[k] * coef;
      P[k] *= decay;
      Mv[k] = Mv[k] + omb1 * (gk - Mv[k]);      // exp_avg.lerp_(grad, 1 - beta1)
      V[k] = b2 * V[k] + omb2 * (gk * gk);      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
      const float denom = sqrtf(V[k]) / bc2_sqrt + eps;
      P[k] -= step * (Mv[k] / denom);           // param.addcdiv_(exp_avg, denom, value=-step_size)
    }
    p4[i] = pp;
    m4[i] = mm;
    v4[i] = vv;
    if (p16) {
      uint2 w;
      w.x = pack_bf16(pp.x, pp.y);
      w.y = pack_bf16(pp.z, pp.w);
      o4[i] = w;
    }
  }
}

// EMA of the student parameters (diffusers EMAModel.step, reference main.py:553-554): shadow -= (1 - decay) * (shadow - p)
__global__ void ema_kernel(float* __restrict__ shadow, const float* __restrict__ p, long n, float one_minus_decay) {
  const long n4 = n >> 2;
  float4* s4 = reinterpret_cast<float4*>(shadow);
  const float4* p4 = reinterpret_cast<const float4*>(p);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    float4 a = s4[i];
    const float4 b = p4[i];
    a.x -= one_minus_decay * (a.x - b.x);
    a.y -= one_minus_decay * (a.y - b.y);
    a.z -= one_minus_decay * (a.z - b.z);
    a.w -= one_minus_decay * (a.w - b.w);
    s4[i] = a;
  }
}

}  // namespace tdm

using namespace tdm;

static bool dt_ok2(int d) { return d == TDM_F32 || d == TDM_F16 || d == TDM_BF16; }

extern "C" int tdm_select_states(const void* const* states, const int* state_dtypes, int n_states,
                                 const int64_t* index, void* out, int out_dtype, int B, long per_sample, void* stream) {
  TDM_REQUIRE(states && state_dtypes && index && out, "tdm_select_states: null input");
  TDM_REQUIRE(n_states >= 1 && n_states <= 8, "tdm_select_states: n_states must be in [1,8]");
  TDM_REQUIRE(dt_ok2(out_dtype) && B > 0 && per_sample > 0, "tdm_select_states: bad dtype or shape");
  SelectArgs a;
  for (int i = 0; i < 8; ++i) {
    a.states[i] = i < n_states ? states[i] : nullptr;
    a.state_dt[i] = i < n_states ? state_dtypes[i] : 0;
    TDM_REQUIRE(i >= n_states || (states[i] && dt_ok2(state_dtypes[i])), "tdm_select_states: bad state %d", i);
  }
  a.n_states = n_states; a.index = index; a.out = out; a.out_dt = out_dtype; a.per = per_sample; a.B = B;
  ::tdm::count_launch(); select_states_kernel<<<ts_grid((long)B * per_sample, 256), 256, 0, (cudaStream_t)stream>>>(a);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_mse_fwd_bwd(const float* pred, const float* target, double* loss_out, float* grad_out,
                               float grad_scale, int B, int C, int Cout, int HW, void* stream) {
  TDM_REQUIRE(pred && target && loss_out, "tdm_mse_fwd_bwd: null input");
  TDM_REQUIRE(B > 0 && C > 0 && Cout >= C && HW > 0, "tdm_mse_fwd_bwd: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  TDM_CHECK_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(double), s));
  ::tdm::count_launch(); mse_fwd_bwd_kernel<<<ts_grid((long)B * C * HW, 256), 256, 0, s>>>(pred, target, loss_out, grad_out, grad_scale, B, C,
                                                                  Cout, HW);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_sumsq_f32(const float* x, int64_t n, double* accum, int zero_first, void* stream) {
  TDM_REQUIRE(x && accum && n > 0, "tdm_sumsq_f32: null input");
  TDM_REQUIRE(((uintptr_t)x & 15) == 0, "tdm_sumsq_f32: buffer must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (zero_first) TDM_CHECK_CUDA(cudaMemsetAsync(accum, 0, sizeof(double), s));
  ::tdm::count_launch(); sumsq_kernel<<<ts_grid(n / 4 + 1, 256), 256, 0, s>>>(x, (long)n, accum);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_clip_coef(const double* sumsq, float max_norm, float inv_scale, float* coef_out, float* norm_out,
                             void* stream) {
  TDM_REQUIRE(sumsq && coef_out, "tdm_clip_coef: null input");
  ::tdm::count_launch(); clip_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sumsq, max_norm, inv_scale, coef_out, norm_out);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, void* param_bf16,
                              int64_t n, double lr, double beta1, double beta2, double eps, double weight_decay, int step,
                              const float* grad_coef, void* stream) {
  TDM_REQUIRE(param && grad && exp_avg && exp_avg_sq, "tdm_adamw_step: null input");
  TDM_REQUIRE(n > 0 && n % 4 == 0, "tdm_adamw_step: n=%lld must be a positive multiple of 4", (long long)n);
  TDM_REQUIRE(step >= 1, "tdm_adamw_step: step counts from 1");
  TDM_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0 &&
                  ((uintptr_t)param_bf16 & 7) == 0,
              "tdm_adamw_step: buffers must be 16-byte aligned");
  // Hyper-parameters arrive as doubles and every derived scalar is formed in double before it is rounded to fp32, the way
  // torch.optim.AdamW forms them in Python (1 - beta2 = 0.001, not 1.f - 0.999f = 0.00100005).
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2_sqrt = sqrt(1.0 - pow(beta2, (double)step));
  ::tdm::count_launch();
  adamw_kernel<<<ts_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(
      param, grad, exp_avg, exp_avg_sq, (bf16*)param_bf16, (long)n, (float)beta2, (float)(1.0 - beta1),
      (float)(1.0 - beta2), (float)eps, (float)(1.0 - lr * weight_decay), (float)(lr / bc1), (float)bc2_sqrt, grad_coef);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_ema_update(float* shadow, const float* param, int64_t n, float decay, void* stream) {
  TDM_REQUIRE(shadow && param && n > 0 && n % 4 == 0, "tdm_ema_update: bad input");
  TDM_REQUIRE((((uintptr_t)shadow | (uintptr_t)param) & 15) == 0, "tdm_ema_update: buffers must be 16-byte aligned");
  ::tdm::count_launch(); ema_kernel<<<ts_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(shadow, param, (long)n, 1.f - decay);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}
