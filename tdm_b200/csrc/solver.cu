// Few-step solver update, t1->t2 re-noising, trajectory select and the TDM loss + closed-form
// gradient: the reference-owned elementwise math (src/utils.py:47-73, src/models.py:46-58,
// src/predictor.py:42-58,76-97, src/main.py:473-479,514-529) as fused, 128-bit vectorised kernels.
//
// Rounding discipline: every reference op is one IEEE-rounded step (torch eager launches one kernel
// per op), so the fused kernels use __f*_rn intrinsics (no FMA contraction) and re-round through the
// low-precision type wherever the reference op would have produced fp16/bf16.  Given identical
// inputs the results are bit-identical to the reference's eager path.
#include <cuda_fp16.h>
#include "common.cuh"

namespace tdm {

template <int DT> struct LowP;
template <> struct LowP<TDM_F32> {
  typedef float T;
  static __device__ __forceinline__ float rd(const void* p, long i) { return reinterpret_cast<const float*>(p)[i]; }
  static __device__ __forceinline__ float rn(float v) { return v; }
  static __device__ __forceinline__ void wr(void* p, long i, float v) { reinterpret_cast<float*>(p)[i] = v; }
};
template <> struct LowP<TDM_F16> {
  typedef __half T;
  static __device__ __forceinline__ float rd(const void* p, long i) { return __half2float(reinterpret_cast<const __half*>(p)[i]); }
  static __device__ __forceinline__ float rn(float v) { return __half2float(__float2half_rn(v)); }
  static __device__ __forceinline__ void wr(void* p, long i, float v) { reinterpret_cast<__half*>(p)[i] = __float2half_rn(v); }
};
template <> struct LowP<TDM_BF16> {
  typedef __nv_bfloat16 T;
  static __device__ __forceinline__ float rd(const void* p, long i) { return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]); }
  static __device__ __forceinline__ float rn(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
  static __device__ __forceinline__ void wr(void* p, long i, float v) { reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v); }
};

__device__ __forceinline__ float rd_dt(const void* p, long i, int dt) {
  if (dt == TDM_F32) return LowP<TDM_F32>::rd(p, i);
  if (dt == TDM_F16) return LowP<TDM_F16>::rd(p, i);
  return LowP<TDM_BF16>::rd(p, i);
}
__device__ __forceinline__ float rn_dt(float v, int dt) {
  if (dt == TDM_F32) return v;
  if (dt == TDM_F16) return LowP<TDM_F16>::rn(v);
  return LowP<TDM_BF16>::rn(v);
}
__device__ __forceinline__ void wr_dt(void* p, long i, float v, int dt) {
  if (dt == TDM_F32) LowP<TDM_F32>::wr(p, i, v);
  else if (dt == TDM_F16) LowP<TDM_F16>::wr(p, i, v);
  else LowP<TDM_BF16>::wr(p, i, v);
}
__device__ __forceinline__ long wrap_index(long t, int len) { return t < 0 ? t + len : t; }  // python indexing

// 128-bit (fp32) / 64-bit (fp16, bf16) vector access to 4 consecutive elements; i % 4 == 0 and 16-byte aligned bases.
__device__ __forceinline__ void ld4_dt(const void* p, long i, int dt, float* o) {
  if (dt == TDM_F32) {
    const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  } else {
    const uint2 v = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p) + i);
    if (dt == TDM_F16) {
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
      o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    } else {
      const float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y);
      o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    }
  }
}
__device__ __forceinline__ void st4_dt(void* p, long i, int dt, const float* v) {
  if (dt == TDM_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = make_float4(v[0], v[1], v[2], v[3]);
  } else if (dt == TDM_F16) {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    uint2 w;
    w.x = *reinterpret_cast<uint32_t*>(&a);
    w.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p) + i) = w;
  } else {
    uint2 w;
    w.x = pack_bf16(v[0], v[1]);
    w.y = pack_bf16(v[2], v[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p) + i) = w;
  }
}

// ------------------------------------------------------------------------------------------------
// solver step: eps-half select + CFG combine + predicted_origin + scheduler.add_noise + low-p cast
// ------------------------------------------------------------------------------------------------
struct SolverArgs {
  const float* out_c;   // [B, Cout, HW] conditional model output
  const float* out_u;   // optional unconditional output
  float cfg;
  const void* x_t;      // [B, C, HW]
  int x_dt;
  const int64_t* t;
  const int64_t* t_next;  // optional
  const void* alphas;   // sqrt(abar) table in table_dt
  const void* sigmas;
  int table_dt, table_len;
  const float* acp;     // alphas_cumprod fp32
  int pred_type;        // 0 eps, 1 v
  float* eps_out;       // optional [B,C,HW]
  float* x0_out;        // optional [B,C,HW]
  void* x_next;         // optional [B,C,HW] in next_dt
  int next_dt;
  int B, C, Cout, HW;
};

// One element of the solver step (identical op sequence for the scalar and the 4-wide kernel).
__device__ __forceinline__ void solver_elem(const SolverArgs& a, float e, float eu, float x, float al, float sg, float sa,
                                            float ss, float& e_out, float& x0, float& xn) {
  if (a.out_u) e = __fadd_rn(eu, __fmul_rn(a.cfg, __fsub_rn(e, eu)));  // predictor.py:42 / models.py:45
  // promotion: fp32 model output makes every op fp32; the table entries are exact in fp32
  if (a.pred_type == 0) x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(sg, e)), al);       // utils.py:51
  else  // utils.py:55: alphas * sample is a low-precision product when both operands are low precision
    x0 = __fsub_rn(rn_dt(__fmul_rn(al, x), a.table_dt == a.x_dt ? a.x_dt : TDM_F32), __fmul_rn(sg, e));
  e_out = e;
  xn = __fadd_rn(__fmul_rn(sa, x0), __fmul_rn(ss, e));  // scheduler.add_noise, cast by the caller (models.py:58)
}

template <int VEC>
__global__ void solver_step_kernel(SolverArgs a) {
  const long per = (long)a.C * a.HW;
  const long total = (long)a.B * per / VEC;
  for (long iv = (long)blockIdx.x * blockDim.x + threadIdx.x; iv < total; iv += (long)gridDim.x * blockDim.x) {
    const long i = iv * VEC;
    const int b = int(i / per);
    const long r = i - (long)b * per;  // c*HW + hw; the eps half is the first C channels of Cout
    const long src = (long)b * a.Cout * a.HW + r;
    const long t = wrap_index(a.t[b], a.table_len);
    const float sg = rd_dt(a.sigmas, t, a.table_dt);
    const float al = rd_dt(a.alphas, t, a.table_dt);
    float sa = 0.f, ss = 0.f;
    if (a.x_next) {
      const float ac = a.acp[wrap_index(a.t_next[b], a.table_len)];
      sa = __fsqrt_rn(ac);                  // alphas_cumprod[t] ** 0.5  (torch: pow 0.5 == sqrt)
      ss = __fsqrt_rn(__fsub_rn(1.f, ac));  // (1 - alphas_cumprod[t]) ** 0.5
    }
    float e[VEC], eu[VEC], x[VEC], eo[VEC], x0[VEC], xn[VEC];
    if (VEC == 4) {
      ld4_dt(a.out_c, src, TDM_F32, e);
      if (a.out_u) ld4_dt(a.out_u, src, TDM_F32, eu);
      ld4_dt(a.x_t, i, a.x_dt, x);
    } else {
      e[0] = a.out_c[src];
      if (a.out_u) eu[0] = a.out_u[src];
      x[0] = rd_dt(a.x_t, i, a.x_dt);
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) solver_elem(a, e[k], a.out_u ? eu[k] : 0.f, x[k], al, sg, sa, ss, eo[k], x0[k], xn[k]);
    if (VEC == 4) {
      if (a.eps_out) st4_dt(a.eps_out, i, TDM_F32, eo);
      if (a.x0_out) st4_dt(a.x0_out, i, TDM_F32, x0);
      if (a.x_next) st4_dt(a.x_next, i, a.next_dt, xn);
    } else {
      if (a.eps_out) a.eps_out[i] = eo[0];
      if (a.x0_out) a.x0_out[i] = x0[0];
      if (a.x_next) wr_dt(a.x_next, i, xn[0], a.next_dt);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Predictor.add_noise / obtain_mixed_noise (predictor.py:76-97) with optional per-sample trajectory
// select (main.py:479 made a per-sample gather).
// ------------------------------------------------------------------------------------------------
struct RenoiseArgs {
  const void* states[8];  // candidate states [B,C,HW]; when n_states==1 it is the sample itself
  int state_dt[8];
  int n_states;
  const int64_t* index;   // [B] which state each sample takes (ignored when n_states==1)
  const void* noise;      // [B,C,HW]
  int noise_dt;
  const int64_t* t1;
  const int64_t* t2;
  const void* alphas;
  const void* sigmas;
  int table_dt, table_len;
  void* out;              // samples at level t2, out_dt
  int out_dt;
  int mixed;              // 1: obtain_mixed_noise (states[0] is the model noise)
  long per, B;
};

__device__ __forceinline__ int promote(int a, int b) {  // torch type promotion among {f32,f16,bf16}
  if (a == b) return a;
  return TDM_F32;  // f16 x bf16 also promotes to f32
}

template <int VEC>
__global__ void renoise_kernel(RenoiseArgs a) {
  const long total = a.B * a.per / VEC;
  for (long iv = (long)blockIdx.x * blockDim.x + threadIdx.x; iv < total; iv += (long)gridDim.x * blockDim.x) {
    const long i = iv * VEC;
    const int b = int(i / a.per);
    const int td = a.table_dt;
    const long t1 = wrap_index(a.t1[b], a.table_len), t2 = wrap_index(a.t2[b], a.table_len);
    const float s1 = rd_dt(a.sigmas, t1, td), a1 = rd_dt(a.alphas, t1, td);
    const float s2 = rd_dt(a.sigmas, t2, td), a2 = rd_dt(a.alphas, t2, td);
    // beta = (s2**2 - (a2/a1*s1)**2) ** 0.5, every op rounded in the table dtype (predictor.py:82-83)
    const float q = rn_dt(__fmul_rn(rn_dt(__fdiv_rn(a2, a1), td), s1), td);
    const float beta2 = rn_dt(__fsub_rn(rn_dt(__fmul_rn(s2, s2), td), rn_dt(__fmul_rn(q, q), td)), td);
    const float beta = rn_dt(__fsqrt_rn(beta2), td);  // NaN when t2 < t1 (reference defect D3 kept)
    int si = 0;
    if (a.n_states > 1) {
      si = (int)a.index[b];
      if (si < 0) si += a.n_states;
    }
    const int sd = a.state_dt[si];
    const int d1 = promote(sd, td);
    const int dn = promote(td, a.noise_dt);
    const int d2 = promote(d1, dn);
    float x[VEC], nz[VEC], r[VEC];
    if (VEC == 4) {
      ld4_dt(a.states[si], i, sd, x);
      ld4_dt(a.noise, i, a.noise_dt, nz);
    } else {
      x[0] = rd_dt(a.states[si], i, sd);
      nz[0] = rd_dt(a.noise, i, a.noise_dt);
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const float y = rn_dt(__fmul_rn(rn_dt(__fdiv_rn(x[k], a1), d1), a2), d1);  // samples / alphas * alphas_new
      const float bn = rn_dt(__fmul_rn(beta, nz[k]), dn);
      r[k] = rn_dt(__fadd_rn(y, bn), d2);
      if (a.mixed) r[k] = rn_dt(__fdiv_rn(r[k], s2), promote(d2, td));  // predictor.py:96
    }
    if (VEC == 4) st4_dt(a.out, i, a.out_dt, r);
    else wr_dt(a.out, i, r[0], a.out_dt);
  }
}

// ------------------------------------------------------------------------------------------------
// TDM loss (main.py:514-529) and its closed-form gradient.
//   phase 1: w_b = mean_{chw} |double(m) - double(sd_cfg)|            (fp64, per sample)
//   phase 2: loss = mean( double(rho(m - coop)) / w_b ),  grad = rho'(m - coop) / (w_b * numel)
// ------------------------------------------------------------------------------------------------
struct LossArgs {
  const float* m;     // model_latents
  const float* sd;    // teacher cond x0
  const float* sdu;   // teacher uncond x0
  float cfg_m1;       // cfg - 1 (as fp32 scalar)
  int use_cfg;        // cfg > 1
  int use_huber;
  float c, c2;        // huber_c and huber_c**2 rounded to fp32
  double* wsum;       // [B] zero-initialised scratch
  double* loss;       // [1] zero-initialised
  float* grad;        // optional [B, per]
  float grad_scale;   // upstream scale on the gradient (loss scaling / accumulation)
  long per, B;
  int vec4;           // per % 4 == 0 and 16-byte aligned inputs: 128-bit loads
};

__device__ __forceinline__ double block_sum_double(double v) {
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

__global__ void loss_weight_kernel(LossArgs a) {  // grid (chunks, B)
  const int b = blockIdx.y;
  const float* m = a.m + (long)b * a.per;
  const float* sd = a.sd + (long)b * a.per;
  const float* su = a.sdu + (long)b * a.per;
  double acc = 0.0;
  if (a.vec4) {  // 128-bit loads; the per-element op sequence is unchanged
    for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < a.per; i += (long)gridDim.x * blockDim.x * 4) {
      const float4 mv = *reinterpret_cast<const float4*>(m + i), sv = *reinterpret_cast<const float4*>(sd + i),
                   uv = *reinterpret_cast<const float4*>(su + i);
      const float mm[4] = {mv.x, mv.y, mv.z, mv.w}, ss[4] = {sv.x, sv.y, sv.z, sv.w}, uu[4] = {uv.x, uv.y, uv.z, uv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float sc = __fadd_rn(ss[k], __fmul_rn(a.cfg_m1, __fsub_rn(ss[k], uu[k])));
        acc += fabs((double)mm[k] - (double)sc);
      }
    }
  } else
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < a.per; i += (long)gridDim.x * blockDim.x) {
    // sd_latents_cfg = sd + (cfg-1) * (sd - sd_u), fp32 op by op (main.py:518)
    const float sc = __fadd_rn(sd[i], __fmul_rn(a.cfg_m1, __fsub_rn(sd[i], su[i])));
    acc += fabs((double)m[i] - (double)sc);
  }
  acc = block_sum_double(acc);
  if (threadIdx.x == 0) atomicAdd(&a.wsum[b], acc);
}

__global__ void loss_value_kernel(LossArgs a) {  // grid (chunks, B)
  const int b = blockIdx.y;
  const float* m = a.m + (long)b * a.per;
  const float* sd = a.sd + (long)b * a.per;
  const float* su = a.sdu + (long)b * a.per;
  const double w = a.wsum[b] / (double)a.per;
  const double n = (double)a.per * (double)a.B;
  double acc = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < a.per; i += (long)gridDim.x * blockDim.x) {
    float coop = sd[i];
    if (a.use_cfg) coop = __fadd_rn(sd[i], __fmul_rn(a.cfg_m1, __fsub_rn(sd[i], su[i])));
    const float d = __fsub_rn(m[i], coop);
    const float d2 = __fmul_rn(d, d);
    double g;
    if (a.use_huber) {
      const float rt = __fsqrt_rn(__fadd_rn(d2, a.c2));
      acc += (double)__fsub_rn(rt, a.c) / w;
      g = (double)d / (double)rt / (w * n);
    } else {
      acc += (double)d2 / w;
      g = 2.0 * (double)d / (w * n);
    }
    if (a.grad) a.grad[(long)b * a.per + i] = (float)(g * (double)a.grad_scale);
  }
  acc = block_sum_double(acc);
  if (threadIdx.x == 0) atomicAdd(a.loss, acc / n);
}

// DDPMScheduler.add_noise with torch's dtype semantics (alphas_cumprod cast to x0's dtype first).
__global__ void sched_add_noise_kernel(const void* x0, int xd, const void* noise, int nd, const int64_t* t,
                                       const float* acp, int table_len, void* out, int od, long per, long B) {
  const long total = B * per;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int b = int(i / per);
    const float ac = rn_dt(acp[wrap_index(t[b], table_len)], xd);
    const float sa = rn_dt(__fsqrt_rn(ac), xd);
    const float ss = rn_dt(__fsqrt_rn(rn_dt(__fsub_rn(1.f, ac), xd)), xd);
    const float t1 = rn_dt(__fmul_rn(sa, rd_dt(x0, i, xd)), xd);
    const int dn = promote(xd, nd);
    const float t2 = rn_dt(__fmul_rn(ss, rd_dt(noise, i, nd)), dn);
    wr_dt(out, i, rn_dt(__fadd_rn(t1, t2), promote(xd, dn)), od);
  }
}

// d(model outputs) from d(eps), d(x0) of one epsilon-prediction solver step
__global__ void solver_step_bwd_kernel(const float* g_eps, const float* g_x0, float cfg, int has_u, const int64_t* t,
                                       const void* alphas, const void* sigmas, int td, int table_len, float* d_c,
                                       float* d_u, int B, int C, int Cout, int HW) {
  const long per_out = (long)Cout * HW, per_in = (long)C * HW;
  const long total = (long)B * per_out;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int b = int(i / per_out);
    const long r = i - (long)b * per_out;
    float g = 0.f;
    if (r < per_in) {
      const long j = (long)b * per_in + r;
      if (g_eps) g = g_eps[j];
      if (g_x0) {
        const long tt = wrap_index(t[b], table_len);
        g -= rd_dt(sigmas, tt, td) / rd_dt(alphas, tt, td) * g_x0[j];
      }
    }
    if (d_c) d_c[i] = has_u ? cfg * g : g;
    if (d_u) d_u[i] = (1.f - cfg) * g;
  }
}

static inline int ew_grid(long total, int threads) {
  long g = (total + threads - 1) / threads;
  const long cap = (long)num_sms() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace tdm

using namespace tdm;

static bool dt_ok(int d) { return d == TDM_F32 || d == TDM_F16 || d == TDM_BF16; }

extern "C" int tdm_solver_step(const float* model_out, const float* model_out_uncond, double cfg, const void* x_t,
                               int x_dtype, const int64_t* t, const int64_t* t_next, const void* alphas,
                               const void* sigmas, int table_dtype, int table_len, const float* alphas_cumprod,
                               int prediction_type, float* eps_out, float* x0_out, void* x_next, int next_dtype,
                               int B, int C, int Cout, int HW, void* stream) {
  TDM_REQUIRE(model_out && x_t && t && alphas && sigmas, "tdm_solver_step: null input");
  TDM_REQUIRE(dt_ok(x_dtype) && dt_ok(table_dtype) && dt_ok(next_dtype), "tdm_solver_step: bad dtype enum");
  TDM_REQUIRE(prediction_type == 0 || prediction_type == 1,
              "Prediction type %d currently not supported.", prediction_type);  // utils.py:56-57
  TDM_REQUIRE(B > 0 && C > 0 && Cout >= C && HW > 0, "tdm_solver_step: bad shape");
  TDM_REQUIRE(!x_next || (t_next && alphas_cumprod), "tdm_solver_step: x_next needs t_next and alphas_cumprod");
  SolverArgs a;
  a.out_c = model_out; a.out_u = model_out_uncond; a.cfg = (float)cfg; a.x_t = x_t; a.x_dt = x_dtype;
  a.t = t; a.t_next = t_next; a.alphas = alphas; a.sigmas = sigmas; a.table_dt = table_dtype;
  a.table_len = table_len; a.acp = alphas_cumprod; a.pred_type = prediction_type;
  a.eps_out = eps_out; a.x0_out = x0_out; a.x_next = x_next; a.next_dt = next_dtype;
  a.B = B; a.C = C; a.Cout = Cout; a.HW = HW;
  const long total = (long)B * C * HW;
  const bool vec = (HW % 4 == 0) && ((((uintptr_t)model_out | (uintptr_t)model_out_uncond | (uintptr_t)x_t | (uintptr_t)eps_out |
                                       (uintptr_t)x0_out | (uintptr_t)x_next) & 15) == 0);
  ::tdm::count_launch();
  if (vec) solver_step_kernel<4><<<ew_grid(total / 4, 256), 256, 0, (cudaStream_t)stream>>>(a);
  else solver_step_kernel<1><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(a);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_renoise(const void* const* states, const int* state_dtypes, int n_states, const int64_t* index,
                           const void* noise, int noise_dtype, const int64_t* t1, const int64_t* t2,
                           const void* alphas, const void* sigmas, int table_dtype, int table_len, void* out,
                           int out_dtype, int mixed, int B, long per_sample, void* stream) {
  TDM_REQUIRE(states && state_dtypes && noise && t1 && t2 && alphas && sigmas && out, "tdm_renoise: null input");
  TDM_REQUIRE(n_states >= 1 && n_states <= 8, "tdm_renoise: n_states must be in [1,8]");
  TDM_REQUIRE(n_states == 1 || index, "tdm_renoise: index required when selecting among states");
  TDM_REQUIRE(dt_ok(noise_dtype) && dt_ok(table_dtype) && dt_ok(out_dtype), "tdm_renoise: bad dtype enum");
  RenoiseArgs a;
  for (int i = 0; i < 8; ++i) {
    a.states[i] = i < n_states ? states[i] : nullptr;
    a.state_dt[i] = i < n_states ? state_dtypes[i] : 0;
    TDM_REQUIRE(i >= n_states || (states[i] && dt_ok(state_dtypes[i])), "tdm_renoise: bad state %d", i);
  }
  a.n_states = n_states; a.index = index; a.noise = noise; a.noise_dt = noise_dtype; a.t1 = t1; a.t2 = t2;
  a.alphas = alphas; a.sigmas = sigmas; a.table_dt = table_dtype; a.table_len = table_len; a.out = out;
  a.out_dt = out_dtype; a.mixed = mixed; a.per = per_sample; a.B = B;
  bool vec = (per_sample % 4 == 0) && ((((uintptr_t)noise | (uintptr_t)out) & 15) == 0);
  for (int i = 0; i < n_states; ++i) vec = vec && (((uintptr_t)states[i] & 15) == 0);
  ::tdm::count_launch();
  if (vec) renoise_kernel<4><<<ew_grid((long)B * per_sample / 4, 256), 256, 0, (cudaStream_t)stream>>>(a);
  else renoise_kernel<1><<<ew_grid((long)B * per_sample, 256), 256, 0, (cudaStream_t)stream>>>(a);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_loss_fwd_bwd(const float* model_latents, const float* sd_latents, const float* sd_latents_uncond,
                                double cfg, int use_huber, double huber_c, double* scratch_w, double* loss_out,
                                float* grad_out, float grad_scale, int B, long per_sample, void* stream) {
  TDM_REQUIRE(model_latents && sd_latents && sd_latents_uncond && scratch_w && loss_out, "tdm_loss_fwd_bwd: null input");
  TDM_REQUIRE(B > 0 && per_sample > 0, "tdm_loss_fwd_bwd: bad shape");
  cudaStream_t s = (cudaStream_t)stream;
  TDM_CHECK_CUDA(cudaMemsetAsync(scratch_w, 0, sizeof(double) * B, s));
  TDM_CHECK_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(double), s));
  LossArgs a;
  a.m = model_latents; a.sd = sd_latents; a.sdu = sd_latents_uncond;
  a.cfg_m1 = (float)(cfg - 1.0); a.use_cfg = cfg > 1.0; a.use_huber = use_huber;  // python scalars -> fp32 scalars
  a.c = (float)huber_c; a.c2 = (float)(huber_c * huber_c);
  a.wsum = scratch_w; a.loss = loss_out; a.grad = grad_out; a.grad_scale = grad_scale;
  a.per = per_sample; a.B = B;
  a.vec4 = (per_sample % 4 == 0) && ((((uintptr_t)model_latents | (uintptr_t)sd_latents | (uintptr_t)sd_latents_uncond) & 15) == 0);
  int chunks = (int)((per_sample + 2047) / 2048);
  if (chunks > 64) chunks = 64;
  dim3 grid(chunks, B);
  ::tdm::count_launch(); loss_weight_kernel<<<grid, 256, 0, s>>>(a);
  ::tdm::count_launch(); loss_value_kernel<<<grid, 256, 0, s>>>(a);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_scheduler_add_noise(const void* x0, int x0_dtype, const void* noise, int noise_dtype,
                                       const int64_t* t, const float* alphas_cumprod, int table_len, void* out,
                                       int out_dtype, int B, long per_sample, void* stream) {
  TDM_REQUIRE(x0 && noise && t && alphas_cumprod && out, "tdm_scheduler_add_noise: null input");
  TDM_REQUIRE(dt_ok(x0_dtype) && dt_ok(noise_dtype) && dt_ok(out_dtype), "tdm_scheduler_add_noise: bad dtype enum");
  const int want = x0_dtype == noise_dtype ? x0_dtype : TDM_F32;
  TDM_REQUIRE(out_dtype == want, "tdm_scheduler_add_noise: out dtype must be the promoted type (%d)", want);
  ::tdm::count_launch(); sched_add_noise_kernel<<<ew_grid((long)B * per_sample, 256), 256, 0, (cudaStream_t)stream>>>(
      x0, x0_dtype, noise, noise_dtype, t, alphas_cumprod, table_len, out, out_dtype, per_sample, B);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_solver_step_bwd(const float* g_eps, const float* g_x0, double cfg, int has_uncond, const int64_t* t,
                                   const void* alphas, const void* sigmas, int table_dtype, int table_len,
                                   float* d_out_c, float* d_out_u, int B, int C, int Cout, int HW, void* stream) {
  TDM_REQUIRE((g_eps || g_x0) && d_out_c && t && alphas && sigmas, "tdm_solver_step_bwd: null input");
  TDM_REQUIRE(!has_uncond || d_out_u, "tdm_solver_step_bwd: d_out_u required with an unconditional branch");
  TDM_REQUIRE(dt_ok(table_dtype), "tdm_solver_step_bwd: bad dtype enum");
  ::tdm::count_launch(); solver_step_bwd_kernel<<<ew_grid((long)B * Cout * HW, 256), 256, 0, (cudaStream_t)stream>>>(
      g_eps, g_x0, (float)cfg, has_uncond, t, alphas, sigmas, table_dtype, table_len, d_out_c,
      has_uncond ? d_out_u : nullptr, B, C, Cout, HW);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}
