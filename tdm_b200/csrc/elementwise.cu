// HBM-bound kernels of the DiT block: adaLN-single LayerNorm+modulate (fwd/bwd), gate backward,
// per-sample column reductions (d shift / d scale / d gate / d bias), modulation-table add,
// timestep sinusoid, patchify / unpatchify.  All are coalesced, 128-bit vectorised and reduce with
// warp shuffles; none of them reshapes work into a GEMM.
// Reference semantics: SURVEY.md Appendix A (diffusers BasicTransformerBlock, ada_norm_single).
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"

namespace tdm {

constexpr int LN_MAXV = 16;  // float4 per lane: D <= 2048

// ---------------------------------------------------------------------------------------------
// h = LN(x) * (1 + scale[b]) + shift[b]   (no affine, eps inside sqrt), x fp32 -> h bf16.
// One warp per row, row held in registers, two-pass mean / variance in fp32.
// ---------------------------------------------------------------------------------------------
template <int NV>  // float4 per lane actually used (D == NV * 128), or 0: runtime count up to LN_MAXV
__global__ void __launch_bounds__(256) ln_modulate_fwd_kernel(const float* __restrict__ x, const float* __restrict__ shift,
                                                            const float* __restrict__ scale, long mod_stride,
                                                            int rows_per_sample, bf16* __restrict__ out,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            int M, int D, float eps) {
  // The eight rows of a CTA almost always belong to one sample: its shift / (1 + scale) rows are staged in shared
  // memory once per CTA, in flight together with the row loads, instead of 18 dependent 16-byte loads per lane after
  // the two reductions (they sat on the critical path of a latency-bound kernel and cost 8x the L1 traffic).
  extern __shared__ float4 ln_mod[];  // [2][D / 4]
  pdl_sync();
  const int warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * (blockDim.x >> 5);
  const int row = row0 + warp;
  const int lane = threadIdx.x & 31;
  constexpr int LN_MAXV = NV > 0 ? NV : tdm::LN_MAXV;  // shadows the global bound: registers sized to the real row
  const int nv = NV > 0 ? NV : (D >> 7);               // float4 per lane
  const int last_row = (row0 + (int)(blockDim.x >> 5) - 1 < M ? row0 + (int)(blockDim.x >> 5) - 1 : M - 1);
  const bool one_sample = (row0 / rows_per_sample) == (last_row / rows_per_sample);  // CTA-uniform
  float4 v[LN_MAXV];
  float s = 0.f;
  if (row < M) {
    const float4* xr = reinterpret_cast<const float4*>(x + (long)row * D);
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (i < nv) {
        v[i] = xr[i * 32 + lane];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
  }
  // modulation rows: loaded into registers now (in flight with the row loads), parked in shared memory only after the
  // two reductions -- the CTA barrier then costs nothing, the loads have long landed
  const int nq = D >> 2;
  float4 m_sh[2], m_sc[2];
  if (one_sample) {
    const long mb0 = (long)(row0 / rows_per_sample) * mod_stride;
    const float4* sh = reinterpret_cast<const float4*>(shift + mb0);
    const float4* sc = reinterpret_cast<const float4*>(scale + mb0);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < nq) {
        m_sh[k] = sh[i];
        m_sc[k] = sc[i];
      }
    }
  }
  float mean = 0.f, rstd = 0.f;
  if (row < M) {
    mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (i < nv) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
    rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
  }
  if (one_sample) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < nq) {
        ln_mod[i] = m_sh[k];
        ln_mod[nq + i] = make_float4(1.f + m_sc[k].x, 1.f + m_sc[k].y, 1.f + m_sc[k].z, 1.f + m_sc[k].w);
      }
    }
    __syncthreads();
  }
  if (row >= M) return;
  uint2* o = reinterpret_cast<uint2*>(out + (long)row * D);
  if (one_sample) {
    const int nq = D >> 2;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (i < nv) {
        const float4 a = ln_mod[i * 32 + lane], b = ln_mod[nq + i * 32 + lane];
        uint2 w;
        w.x = pack_bf16((v[i].x - mean) * rstd * b.x + a.x, (v[i].y - mean) * rstd * b.y + a.y);
        w.y = pack_bf16((v[i].z - mean) * rstd * b.z + a.z, (v[i].w - mean) * rstd * b.w + a.w);
        o[i * 32 + lane] = w;
      }
  } else {
    const long mb = (long)(row / rows_per_sample) * mod_stride;
    const float4* sh = reinterpret_cast<const float4*>(shift + mb);
    const float4* sc = reinterpret_cast<const float4*>(scale + mb);
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (i < nv) {
        const float4 a = sh[i * 32 + lane], b = sc[i * 32 + lane];
        uint2 w;
        w.x = pack_bf16((v[i].x - mean) * rstd * (1.f + b.x) + a.x, (v[i].y - mean) * rstd * (1.f + b.y) + a.y);
        w.y = pack_bf16((v[i].z - mean) * rstd * (1.f + b.z) + a.z, (v[i].w - mean) * rstd * (1.f + b.w) + a.w);
        o[i * 32 + lane] = w;
      }
  }
}

// Streaming form of the kernel above for the shapes of the model (rows_per_sample % 32 == 0): one CTA per 32 consecutive
// rows -- ONE wave of resident CTAs for 8192 rows, no tail -- each warp walks its four rows with the NEXT row's loads in
// flight while the current one is reduced and written, the sample's shift / (1 + scale) rows staged once per CTA.  (The
// one-row-per-warp grid above runs 1024 CTAs in 2.3 waves of 444: 13.5-15 us for 56.6 MB = 0.58 of the measured copy
// bandwidth; the ramp and the last, 30 %-full wave are what this form removes.)
template <int NV>
__global__ void __launch_bounds__(256, 2) ln_modulate_fwd_stream_kernel(const float* __restrict__ x, const float* __restrict__ shift,
                                                                      const float* __restrict__ scale, long mod_stride,
                                                                      int rows_per_sample, bf16* __restrict__ out,
                                                                      float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                      int M, float eps) {
  constexpr int D = NV * 128;
  constexpr int RPW = 4;  // rows per warp
  __shared__ float4 s_mod[2 * NV * 32];  // shift, then 1 + scale
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_cta = blockIdx.x * 32;
  const int row0 = row_cta + warp * RPW;
  float4 v[NV], nx[NV];
  auto load_row = [&](float4* dst, int row) {
    const float4* xr = reinterpret_cast<const float4*>(x + (long)row * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) dst[i] = xr[i * 32 + lane];
  };
  if (row0 < M) load_row(v, row0);
  {  // the CTA's 32 rows belong to one sample
    const long mb = (long)(row_cta / rows_per_sample) * mod_stride;
    const float4* sh = reinterpret_cast<const float4*>(shift + mb);
    const float4* sc = reinterpret_cast<const float4*>(scale + mb);
    for (int i = threadIdx.x; i < NV * 32; i += 256) {
      const float4 a = sh[i], b = sc[i];
      s_mod[i] = a;
      s_mod[NV * 32 + i] = make_float4(1.f + b.x, 1.f + b.y, 1.f + b.z, 1.f + b.w);
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int r = 0; r < RPW; ++r) {
    const int row = row0 + r;
    if (row >= M) break;
    if (r + 1 < RPW && row + 1 < M) load_row(nx, row + 1);  // in flight under this row's reductions and stores
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
    uint2* o = reinterpret_cast<uint2*>(out + (long)row * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 a = s_mod[i * 32 + lane], b = s_mod[NV * 32 + i * 32 + lane];
      uint2 w;
      w.x = pack_bf16((v[i].x - mean) * rstd * b.x + a.x, (v[i].y - mean) * rstd * b.y + a.y);
      w.y = pack_bf16((v[i].z - mean) * rstd * b.z + a.z, (v[i].w - mean) * rstd * b.w + a.w);
      o[i * 32 + lane] = w;
    }
    if (r + 1 < RPW) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = nx[i];
    }
  }
}

// dx_out = dx_in + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dh * (1 + scale[b])
template <int NV>
__global__ void __launch_bounds__(256) ln_modulate_bwd_kernel(const bf16* __restrict__ dh, const float* __restrict__ x,
                                                            const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                                                            const float* __restrict__ scale, long mod_stride, int rows_per_sample,
                                                            const float* __restrict__ dx_in, float* __restrict__ dx_out,
                                                            bf16* __restrict__ dx_out_bf16, int M, int D) {
  pdl_sync();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  constexpr int LN_MAXV = NV > 0 ? NV : tdm::LN_MAXV;
  const int nv = NV > 0 ? NV : (D >> 7);
  const float mean = mean_in[row], rstd = rstd_in[row];
  const float4* xr = reinterpret_cast<const float4*>(x + (long)row * D);
  const uint2* dr = reinterpret_cast<const uint2*>(dh + (long)row * D);
  const float4* sc = reinterpret_cast<const float4*>(scale + (long)(row / rows_per_sample) * mod_stride);
  float4 g[LN_MAXV], xh[LN_MAXV];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
    if (i < nv) {
      const float4 xv = xr[i * 32 + lane];
      const uint2 dv = dr[i * 32 + lane];
      const float4 b = sc[i * 32 + lane];
      const float2 d0 = unpack_bf16(dv.x), d1 = unpack_bf16(dv.y);
      xh[i] = make_float4((xv.x - mean) * rstd, (xv.y - mean) * rstd, (xv.z - mean) * rstd, (xv.w - mean) * rstd);
      g[i] = make_float4(d0.x * (1.f + b.x), d0.y * (1.f + b.y), d1.x * (1.f + b.z), d1.y * (1.f + b.w));
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
    }
  const float c1 = warp_sum(s1) / (float)D, c2 = warp_sum(s2) / (float)D;
  const float4* di = dx_in ? reinterpret_cast<const float4*>(dx_in + (long)row * D) : nullptr;
  float4* o = reinterpret_cast<float4*>(dx_out + (long)row * D);
  uint2* ob = dx_out_bf16 ? reinterpret_cast<uint2*>(dx_out_bf16 + (long)row * D) : nullptr;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i)
    if (i < nv) {
      float4 r = make_float4(rstd * (g[i].x - c1 - xh[i].x * c2), rstd * (g[i].y - c1 - xh[i].y * c2),
                             rstd * (g[i].z - c1 - xh[i].z * c2), rstd * (g[i].w - c1 - xh[i].w * c2));
      if (di) {
        const float4 p = di[i * 32 + lane];
        r.x += p.x; r.y += p.y; r.z += p.z; r.w += p.w;
      }
      o[i * 32 + lane] = r;
      if (ob) {
        uint2 w;
        w.x = pack_bf16(r.x, r.y);
        w.y = pack_bf16(r.z, r.w);
        ob[i * 32 + lane] = w;
      }
    }
}

// ---------------------------------------------------------------------------------------------
// Blocked per-sample column reductions.  blockDim (32, 8): thread.x owns 4 consecutive columns
// (block = 128 columns), thread.y strides over the rows of one row-chunk; partials are combined
// in shared memory and added to out[sample, col] with one atomic per column per block.
//   MODE 0: out0 += sum dh            out1 += sum dh * xhat          (d shift, d scale)
//   MODE 1: out0 += sum dx * y        and dy = bf16(gate * dx)        (gate backward)
//   MODE 2: out0 += sum a (bf16)                                      (bias gradient)
//   MODE 3: dy = bf16(gate * dx) only                                  (gate backward, frozen gate)
// ---------------------------------------------------------------------------------------------
struct ColRedArgs {
  const bf16* a_bf16;   // dh (0), y (1), a (2)
  const float* x;       // x (0), dx (1,3)
  const float* mean;
  const float* rstd;
  const float* gate;    // (1,3) [B, gate_stride]
  long gate_stride;
  bf16* dy;             // (1,3)
  float* out0;
  float* out1;
  long out_stride;      // elements between samples in out0/out1
  int rows_per_sample, rows_per_block, M, D;
  long ld;              // row stride of the inputs
};

template <int MODE>
__global__ void __launch_bounds__(256) colreduce_kernel(ColRedArgs a) {
  __shared__ float4 sm0[8][32];
  __shared__ float4 sm1[8][32];
  pdl_sync();
  const int col = blockIdx.x * 128 + threadIdx.x * 4;
  const int row0 = blockIdx.y * a.rows_per_block;
  const int b = row0 / a.rows_per_sample;
  const bool col_ok = col < a.D;
  float4 acc0 = make_float4(0, 0, 0, 0), acc1 = make_float4(0, 0, 0, 0);
  float4 gt = make_float4(0, 0, 0, 0);
  if ((MODE == 1 || MODE == 3) && col_ok) gt = *reinterpret_cast<const float4*>(a.gate + (long)b * a.gate_stride + col);
  if (col_ok) {
    const int rend = min(row0 + a.rows_per_block, a.M);
#pragma unroll 4  // independent rows: four rows of loads in flight per thread (the kernel is latency-, not bandwidth-bound)
    for (int r = row0 + threadIdx.y; r < rend; r += 8) {
      const long off = (long)r * a.ld + col;
      if (MODE == 0) {
        const uint2 dv = *reinterpret_cast<const uint2*>(a.a_bf16 + off);
        const float4 xv = *reinterpret_cast<const float4*>(a.x + off);
        const float mu = a.mean[r], rs = a.rstd[r];
        const float2 d0 = unpack_bf16(dv.x), d1 = unpack_bf16(dv.y);
        acc0.x += d0.x; acc0.y += d0.y; acc0.z += d1.x; acc0.w += d1.y;
        acc1.x += d0.x * (xv.x - mu) * rs; acc1.y += d0.y * (xv.y - mu) * rs;
        acc1.z += d1.x * (xv.z - mu) * rs; acc1.w += d1.y * (xv.w - mu) * rs;
      } else if (MODE == 1 || MODE == 3) {
        const float4 dx = *reinterpret_cast<const float4*>(a.x + off);
        if (MODE == 1) {
          const uint2 yv = *reinterpret_cast<const uint2*>(a.a_bf16 + off);
          const float2 y0 = unpack_bf16(yv.x), y1 = unpack_bf16(yv.y);
          acc0.x += dx.x * y0.x; acc0.y += dx.y * y0.y; acc0.z += dx.z * y1.x; acc0.w += dx.w * y1.y;
        }
        uint2 w;
        w.x = pack_bf16(gt.x * dx.x, gt.y * dx.y);
        w.y = pack_bf16(gt.z * dx.z, gt.w * dx.w);
        *reinterpret_cast<uint2*>(a.dy + off) = w;
      } else {
        const uint2 av = *reinterpret_cast<const uint2*>(a.a_bf16 + off);
        const float2 a0 = unpack_bf16(av.x), a1 = unpack_bf16(av.y);
        acc0.x += a0.x; acc0.y += a0.y; acc0.z += a1.x; acc0.w += a1.y;
      }
    }
  }
  if (MODE == 3) return;
  sm0[threadIdx.y][threadIdx.x] = acc0;
  if (MODE == 0) sm1[threadIdx.y][threadIdx.x] = acc1;
  __syncthreads();
  if (threadIdx.y == 0 && col_ok) {
    float4 t0 = sm0[0][threadIdx.x], t1 = sm1[0][threadIdx.x];
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      const float4 u = sm0[j][threadIdx.x];
      t0.x += u.x; t0.y += u.y; t0.z += u.z; t0.w += u.w;
      if (MODE == 0) {
        const float4 w = sm1[j][threadIdx.x];
        t1.x += w.x; t1.y += w.y; t1.z += w.z; t1.w += w.w;
      }
    }
    float* o0 = a.out0 + (long)b * a.out_stride + col;
    atomicAdd(o0 + 0, t0.x); atomicAdd(o0 + 1, t0.y); atomicAdd(o0 + 2, t0.z); atomicAdd(o0 + 3, t0.w);
    if (MODE == 0) {
      float* o1 = a.out1 + (long)b * a.out_stride + col;
      atomicAdd(o1 + 0, t1.x); atomicAdd(o1 + 1, t1.y); atomicAdd(o1 + 2, t1.z); atomicAdd(o1 + 3, t1.w);
    }
  }
}

static int pick_rows_per_block(int rows_per_sample) {
  int r = 128;
  while (r > 1 && rows_per_sample % r != 0) r >>= 1;
  return r;
}

template <int MODE>
static int launch_colreduce(ColRedArgs a, cudaStream_t s) {
  a.rows_per_block = pick_rows_per_block(a.rows_per_sample);
  dim3 grid((a.D + 127) / 128, (a.M + a.rows_per_block - 1) / a.rows_per_block);
  ::tdm::count_launch();
  TDM_CHECK_CUDA(launch_pdl(colreduce_kernel<MODE>, grid, dim3(32, 8), (size_t)0, s, a));
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

// ---------------------------------------------------------------------------------------------
// small conditioning kernels
// ---------------------------------------------------------------------------------------------
// mod[b, j, d] = table[j, d] + (per_j ? tvec[b, j*D + d] : tvec[b, d])
__global__ void modulation_kernel(const float* table, const float* tvec, int per_j, float* out, int B, int J, int D) {
  pdl_sync();
  const long total = (long)B * J * D;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int d = int(i % D);
    const int j = int((i / D) % J);
    const int b = int(i / ((long)D * J));
    out[i] = table[(long)j * D + d] + (per_j ? tvec[(long)b * J * D + (long)j * D + d] : tvec[(long)b * D + d]);
  }
}
// dtable[j,d] += sum_b dmod[b,j,d];  dtvec (+)= dmod (per_j) or sum_j dmod (broadcast)
__global__ void modulation_bwd_kernel(const float* dmod, float* dtable, float* dtvec, int per_j, int accumulate_tvec,
                                      int B, int J, int D) {
  pdl_sync();
  const long jd = (long)J * D;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < jd; i += (long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) {
      const float g = dmod[(long)b * jd + i];
      s += g;
      if (dtvec && per_j) {
        float* p = dtvec + (long)b * jd + i;
        *p = accumulate_tvec ? *p + g : g;
      }
    }
    if (dtable) dtable[i] += s;
  }
  if (dtvec && !per_j) {
    const long bd = (long)B * D;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < bd; i += (long)gridDim.x * blockDim.x) {
      const int d = int(i % D);
      const int b = int(i / D);
      float s = 0.f;
      for (int j = 0; j < J; ++j) s += dmod[(long)b * jd + (long)j * D + d];
      dtvec[i] = accumulate_tvec ? dtvec[i] + s : s;
    }
  }
}

// Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0): out[b] = cat[cos(t f), sin(t f)]
__global__ void timestep_sinusoid_kernel(const int64_t* t, bf16* out, int B, int dim) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i % half;
  const float freq = expf(-logf(10000.f) * (float)k / (float)half);
  const float arg = (float)t[b] * freq;
  out[(long)b * dim + k] = __float2bfloat16_rn(cosf(arg));
  out[(long)b * dim + half + k] = __float2bfloat16_rn(sinf(arg));
}

// latent [B,C,H,W] (dt) -> patches bf16 [B*gh*gw, C*p*p], column = c*p*p + py*p + px (Conv2d weight order)
__global__ void patchify_kernel(const void* latent, int dt, bf16* out, int B, int C, int H, int W, int p) {
  const int gh = H / p, gw = W / p, kk = C * p * p;
  const long total = (long)B * gh * gw * kk;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int col = int(i % kk);
    const long tok = i / kk;
    const int px = col % p, py = (col / p) % p, c = col / (p * p);
    const int gx = int(tok % gw), gy = int((tok / gw) % gh), b = int(tok / ((long)gw * gh));
    const long src = (((long)b * C + c) * H + (gy * p + py)) * W + (gx * p + px);
    float v;
    if (dt == TDM_F32) v = reinterpret_cast<const float*>(latent)[src];
    else if (dt == TDM_F16) v = __half2float(reinterpret_cast<const __half*>(latent)[src]);
    else v = __bfloat162float(reinterpret_cast<const bf16*>(latent)[src]);
    out[i] = __float2bfloat16_rn(v);
  }
}
// y fp32 [B*gh*gw, ldy] (column = (py*p+px)*Cout + c) -> out [B, Cout, H, W] fp32 ("nhwpqc->nchpwq")
__global__ void unpatchify_kernel(const float* y, long ldy, float* out, int B, int Cout, int H, int W, int p) {
  const long total = (long)B * Cout * H * W;
  const int gw = W / p, gh = H / p;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int xw = int(i % W), yh = int((i / W) % H), c = int((i / ((long)W * H)) % Cout), b = int(i / ((long)W * H * Cout));
    const int gx = xw / p, px = xw % p, gy = yh / p, py = yh % p;
    const long tok = ((long)b * gh + gy) * gw + gx;
    out[i] = y[tok * ldy + (py * p + px) * Cout + c];
  }
}
// d out [B, Cgrad, H, W] fp32 (gradient of the first Cgrad channels; the rest are zero) -> dy bf16 [M, p*p*Cout]
__global__ void unpatchify_bwd_kernel(const float* dout, bf16* dy, int B, int Cgrad, int Cout, int H, int W, int p) {
  const int gw = W / p, gh = H / p, nc = p * p * Cout;
  const long total = (long)B * gh * gw * nc;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int col = int(i % nc);
    const long tok = i / nc;
    const int c = col % Cout, px = (col / Cout) % p, py = col / (Cout * p);
    const int gx = int(tok % gw), gy = int((tok / gw) % gh), b = int(tok / ((long)gw * gh));
    float v = 0.f;
    if (c < Cgrad) v = dout[(((long)b * Cgrad + c) * H + (gy * p + py)) * W + (gx * p + px)];
    dy[i] = __float2bfloat16_rn(v);
  }
}

__global__ void cast_f32_bf16_kernel(const float* in, bf16* out, long n) {
  for (long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long)gridDim.x * blockDim.x * 4) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(in + i);
      uint2 w;
      w.x = pack_bf16(v.x, v.y);
      w.y = pack_bf16(v.z, v.w);
      *reinterpret_cast<uint2*>(out + i) = w;
    } else {
      for (long j = i; j < n; ++j) out[j] = __float2bfloat16_rn(in[j]);
    }
  }
}

static inline int ew_grid(long total, int threads) {
  long g = (total + threads - 1) / threads;
  const long cap = (long)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace tdm

using namespace tdm;

extern "C" int tdm_ln_modulate_fwd(const float* x, const float* shift, const float* scale, int64_t mod_stride,
                                   int rows_per_sample, void* out_bf16, float* mean, float* rstd, int M, int D,
                                   float eps, void* stream) {
  TDM_REQUIRE(x && shift && scale && out_bf16, "tdm_ln_modulate_fwd: null input");
  TDM_REQUIRE(D % 128 == 0 && D <= 128 * LN_MAXV, "tdm_ln_modulate_fwd: D=%d must be a multiple of 128 and <= %d", D, 128 * LN_MAXV);
  TDM_REQUIRE(M > 0 && rows_per_sample > 0 && mod_stride % 4 == 0, "tdm_ln_modulate_fwd: bad shape");
  ::tdm::count_launch();
  // model shapes: the streaming form (TDM_B200_LN_STREAM=0 keeps the one-row-per-warp grid for A/B timing)
  const bool no_stream = getenv("TDM_B200_LN_STREAM") && atoi(getenv("TDM_B200_LN_STREAM")) == 0;
  if (!no_stream && rows_per_sample % 32 == 0 && (D == 1152 || D == 256)) {
    const int grid = (M + 31) / 32;
    if (D == 1152)
      TDM_CHECK_CUDA(launch_pdl(ln_modulate_fwd_stream_kernel<9>, dim3(grid), dim3(256), (size_t)0, (cudaStream_t)stream, x, shift, scale,
                                (long)mod_stride, rows_per_sample, (bf16*)out_bf16, mean, rstd, M, eps));
    else
      TDM_CHECK_CUDA(launch_pdl(ln_modulate_fwd_stream_kernel<2>, dim3(grid), dim3(256), (size_t)0, (cudaStream_t)stream, x, shift, scale,
                                (long)mod_stride, rows_per_sample, (bf16*)out_bf16, mean, rstd, M, eps));
    TDM_CHECK_CUDA(cudaGetLastError());
    return TDM_OK;
  }
#define TDM_LN_FWD(NVV) TDM_CHECK_CUDA(launch_pdl(ln_modulate_fwd_kernel<NVV>, dim3((M + 7) / 8), dim3(256), (size_t)(2 * D * 4), (cudaStream_t)stream, \
      x, shift, scale, (long)mod_stride, rows_per_sample, (bf16*)out_bf16, mean, rstd, M, D, eps))
  switch (D >> 7) {
    case 2: TDM_LN_FWD(2); break;
    case 9: TDM_LN_FWD(9); break;
    default: TDM_LN_FWD(0); break;
  }
#undef TDM_LN_FWD
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_ln_modulate_bwd(const void* dh_bf16, const float* x, const float* mean, const float* rstd,
                                   const float* scale, int64_t mod_stride, int rows_per_sample, const float* dx_in,
                                   float* dx_out, void* dx_out_bf16, float* dshift, float* dscale, int64_t dmod_stride,
                                   int M, int D, void* stream) {
  TDM_REQUIRE(dh_bf16 && x && mean && rstd && scale && dx_out, "tdm_ln_modulate_bwd: null input");
  TDM_REQUIRE(D % 128 == 0 && D <= 128 * LN_MAXV, "tdm_ln_modulate_bwd: D=%d must be a multiple of 128 and <= %d", D, 128 * LN_MAXV);
  TDM_REQUIRE((dshift == nullptr) == (dscale == nullptr), "tdm_ln_modulate_bwd: dshift and dscale come together");
  cudaStream_t s = (cudaStream_t)stream;
  ::tdm::count_launch();
#define TDM_LN_BWD(NVV) TDM_CHECK_CUDA(launch_pdl(ln_modulate_bwd_kernel<NVV>, dim3((M + 7) / 8), dim3(256), (size_t)0, s, \
      (const bf16*)dh_bf16, x, mean, rstd, scale, (long)mod_stride, rows_per_sample, dx_in, dx_out, (bf16*)dx_out_bf16, M, D))
  switch (D >> 7) {
    case 2: TDM_LN_BWD(2); break;
    case 9: TDM_LN_BWD(9); break;
    default: TDM_LN_BWD(0); break;
  }
#undef TDM_LN_BWD
  TDM_CHECK_CUDA(cudaGetLastError());
  if (dshift) {
    ColRedArgs a = {};
    a.a_bf16 = (const bf16*)dh_bf16; a.x = x; a.mean = mean; a.rstd = rstd;
    a.out0 = dshift; a.out1 = dscale; a.out_stride = dmod_stride;
    a.rows_per_sample = rows_per_sample; a.M = M; a.D = D; a.ld = D;
    return launch_colreduce<0>(a, s);
  }
  return TDM_OK;
}

extern "C" int tdm_gate_bwd(const float* dx, const void* y_bf16, const float* gate, int64_t gate_stride,
                            int rows_per_sample, void* dy_bf16, float* dgate, int64_t dgate_stride, int M, int D,
                            void* stream) {
  TDM_REQUIRE(dx && gate && dy_bf16, "tdm_gate_bwd: null input");
  TDM_REQUIRE(D % 4 == 0 && gate_stride % 4 == 0, "tdm_gate_bwd: D and gate_stride must be multiples of 4");
  TDM_REQUIRE(!dgate || y_bf16, "tdm_gate_bwd: dgate needs the saved pre-gate output y");
  ColRedArgs a = {};
  a.a_bf16 = (const bf16*)y_bf16; a.x = dx; a.gate = gate; a.gate_stride = gate_stride; a.dy = (bf16*)dy_bf16;
  a.out0 = dgate; a.out_stride = dgate_stride; a.rows_per_sample = rows_per_sample; a.M = M; a.D = D; a.ld = D;
  return dgate ? launch_colreduce<1>(a, (cudaStream_t)stream) : launch_colreduce<3>(a, (cudaStream_t)stream);
}

extern "C" int tdm_colsum_bf16(const void* a_bf16, int64_t ld, int M, int N, float* out, void* stream) {
  TDM_REQUIRE(a_bf16 && out, "tdm_colsum_bf16: null input");
  TDM_REQUIRE(N % 4 == 0 && ld % 4 == 0, "tdm_colsum_bf16: N and ld must be multiples of 4");
  ColRedArgs a = {};
  a.a_bf16 = (const bf16*)a_bf16; a.out0 = out; a.out_stride = 0; a.rows_per_sample = 128; a.M = M; a.D = N; a.ld = ld;
  return launch_colreduce<2>(a, (cudaStream_t)stream);
}

extern "C" int tdm_modulation(const float* table, const float* tvec, int per_j, float* out, int B, int J, int D,
                              void* stream) {
  TDM_REQUIRE(table && tvec && out, "tdm_modulation: null input");
  ::tdm::count_launch();
  TDM_CHECK_CUDA(launch_pdl(modulation_kernel, dim3(ew_grid((long)B * J * D, 256)), dim3(256), (size_t)0, (cudaStream_t)stream, table, tvec, per_j, out, B, J, D));
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_modulation_bwd(const float* dmod, float* dtable, float* dtvec, int per_j, int accumulate_tvec,
                                  int B, int J, int D, void* stream) {
  TDM_REQUIRE(dmod, "tdm_modulation_bwd: null input");
  ::tdm::count_launch();
  TDM_CHECK_CUDA(launch_pdl(modulation_bwd_kernel, dim3(ew_grid((long)J * D, 256)), dim3(256), (size_t)0, (cudaStream_t)stream, dmod, dtable,
                            dtvec, per_j, accumulate_tvec, B, J, D));
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_timestep_sinusoid(const int64_t* t, void* out_bf16, int B, int dim, void* stream) {
  TDM_REQUIRE(t && out_bf16 && dim % 2 == 0, "tdm_timestep_sinusoid: bad input");
  const int n = B * dim / 2;
  ::tdm::count_launch(); timestep_sinusoid_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t, (bf16*)out_bf16, B, dim);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_patchify(const void* latent, int dtype, void* out_bf16, int B, int C, int H, int W, int p, void* stream) {
  TDM_REQUIRE(latent && out_bf16 && H % p == 0 && W % p == 0, "tdm_patchify: bad input");
  ::tdm::count_launch(); patchify_kernel<<<ew_grid((long)B * C * H * W, 256), 256, 0, (cudaStream_t)stream>>>(latent, dtype, (bf16*)out_bf16, B, C, H, W, p);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_unpatchify(const float* y, int64_t ldy, float* out, int B, int Cout, int H, int W, int p, void* stream) {
  TDM_REQUIRE(y && out && H % p == 0 && W % p == 0, "tdm_unpatchify: bad input");
  ::tdm::count_launch(); unpatchify_kernel<<<ew_grid((long)B * Cout * H * W, 256), 256, 0, (cudaStream_t)stream>>>(y, ldy, out, B, Cout, H, W, p);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_unpatchify_bwd(const float* dout, void* dy_bf16, int B, int Cgrad, int Cout, int H, int W, int p,
                                  void* stream) {
  TDM_REQUIRE(dout && dy_bf16 && Cgrad <= Cout, "tdm_unpatchify_bwd: bad input");
  ::tdm::count_launch(); unpatchify_bwd_kernel<<<ew_grid((long)B * Cout * H * W, 256), 256, 0, (cudaStream_t)stream>>>(dout, (bf16*)dy_bf16, B, Cgrad, Cout, H, W, p);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

extern "C" int tdm_cast_f32_bf16(const float* in, void* out_bf16, int64_t n, void* stream) {
  TDM_REQUIRE(in && out_bf16, "tdm_cast_f32_bf16: null input");
  TDM_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out_bf16 & 7) == 0, "tdm_cast_f32_bf16: misaligned");
  ::tdm::count_launch(); cast_f32_bf16_kernel<<<ew_grid((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(in, (bf16*)out_bf16, n);
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}
