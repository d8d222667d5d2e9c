// TAESD decoder (diffusers AutoencoderTiny.decode; reference main.py:166, 338, training.py:113-117): latents -> RGB.
// SURVEY.md 8f rank 2, "the step right after sampling".
//
//   x = tanh(x / 3) * 3 ; conv3x3(4 -> 64) + ReLU ; [3 blocks, up x2, conv] x 3 ; 1 block ; conv3x3(64 -> 3) ; x * 2 - 1
//   block(x) = ReLU(conv(ReLU(conv(ReLU(conv(x))))) + x), every conv 3x3, 64 -> 64, padding 1
//
// 141 GFLOP per 512x512 image, 98 % of it in the 64 -> 64 convolutions: those run as implicit GEMMs on tcgen05.
// Activations are NHWC bf16, so one pixel is one 128-byte row and a tile of 128 output pixels is exactly one
// SWIZZLE_128B operand tile.  For every tap (dy, dx) the TMA producer loads the SHIFTED input window through a 4-D
// tensor map (C, W, H, B): out-of-bounds rows / columns are zero-filled by the TMA unit, which is the padding of the
// convolution -- no halo copies, no im2col buffer.  D[128 px, 64] += A_tap[128 px, 64 ci] * W_tap[64 co, 64 ci]^T, nine
// taps x four K = 16 steps per tile, fp32 accumulator in tensor memory (double-buffered, 2 x 64 columns); the nine
// weight taps (72 KB) stay resident in shared memory for the whole persistent CTA.  Epilogue (thread == pixel):
// + bias, + residual (the block's skip), ReLU, bf16, one 128-byte row store.
// The 4 -> 64 input convolution (K = 36), the 64 -> 3 output convolution (N = 3) and the nearest-neighbour upsample are
// HBM-bound and stay on the CUDA cores.
#include <cuda_fp16.h>
#include "common.cuh"

namespace tdm {

constexpr int TC_STAGES = 6;              // A-operand ring: 6 x 16 KB
constexpr int TC_TILE_BYTES = 128 * 128;  // 128 pixels x 64 bf16
constexpr int TC_W_BYTES = 9 * 64 * 128;  // nine [64 co x 64 ci] taps
constexpr int TC_THREADS = 192;           // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

struct ConvParams {
  int B, H, W;
  int bw, bh;          // tile = bh rows x bw columns of pixels, bw * bh == 128
  int tiles_x, tiles_y;
  const float* bias;   // [64] or null
  const bf16* res;     // NHWC residual or null
  bf16* out;           // NHWC
  int relu;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
taesd_conv64_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                          // 72 KB
  uint8_t* sA = sW + TC_W_BYTES;               // ring
  float* sBias = reinterpret_cast<float*>(sA + TC_STAGES * TC_TILE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 64);
  uint64_t* w_full = bars;
  uint64_t* a_full = bars + 1;                 // [TC_STAGES]
  uint64_t* a_empty = a_full + TC_STAGES;      // [TC_STAGES]
  uint64_t* acc_full = a_empty + TC_STAGES;    // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2], 4 elected arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int n_tiles = p.B * tiles_per_img;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < TC_STAGES; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    fence_barrier_init();
  }
  if (threadIdx.x < 64) sBias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;  // weights: written before this launch chain
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  pdl_sync();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmX);
      tma_prefetch_desc(&tmW);
      mbar_arrive_expect_tx(w_full, TC_W_BYTES);
      for (int t = 0; t < 9; ++t) tma_load_2d(sW + t * 64 * 128, &tmW, w_full, 0, t * 64);
      int st = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
        const int y0 = (r / p.tiles_x) * p.bh, x0 = (r % p.tiles_x) * p.bw;
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&a_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&a_full[st], TC_TILE_BYTES);
          // shifted window; rows / columns outside the image are zero-filled == the convolution's padding
          tma_load_4d(sA + st * TC_TILE_BYTES, &tmX, &a_full[st], 0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, b);
          if (++st == TC_STAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    const uint32_t aA = smem_u32(sA), aW = smem_u32(sW);
    mbar_wait(w_full, 0);
    int st = 0, it = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);  // the epilogue has pulled this accumulator into registers
      for (int tap = 0; tap < 9; ++tap) {
        mbar_wait(&a_full[st], ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_acc + buf * 64, make_sdesc_sw128(aA + st * TC_TILE_BYTES + k * 32, 16, 1024),
                      make_sdesc_sw128(aW + tap * 64 * 128 + k * 32, 16, 1024), idesc, (tap | k) != 0);
          umma_commit(&a_empty[st]);
          if (tap == 8) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
        if (++st == TC_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: thread == output pixel == TMEM lane =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
      const int y = (r / p.tiles_x) * p.bh + row / p.bw, x = (r % p.tiles_x) * p.bw + row % p.bw;
      mbar_wait(&acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      uint32_t acc[64];
      tmem_ld32(tmem_acc + buf * 64 + lane_off, acc);
      tmem_ld32(tmem_acc + buf * 64 + 32 + lane_off, acc + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      if (y < p.H && x < p.W) {
        const long pix = ((long)b * p.H + y) * p.W + x;
        const uint4* rrow = p.res ? reinterpret_cast<const uint4*>(p.res + pix * 64) : nullptr;
        uint4* orow = reinterpret_cast<uint4*>(p.out + pix * 64);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[c8 * 8 + i]) + sBias[c8 * 8 + i];
          if (rrow) {
            const uint4 rr = rrow[c8];
            const float2 r0 = unpack_bf16(rr.x), r1 = unpack_bf16(rr.y), r2 = unpack_bf16(rr.z), r3 = unpack_bf16(rr.w);
            v[0] += r0.x; v[1] += r0.y; v[2] += r1.x; v[3] += r1.y; v[4] += r2.x; v[5] += r2.y; v[6] += r3.x; v[7] += r3.y;
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          uint4 w;
          w.x = pack_bf16(v[0], v[1]); w.y = pack_bf16(v[2], v[3]); w.z = pack_bf16(v[4], v[5]); w.w = pack_bf16(v[6], v[7]);
          orow[c8] = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, 128);
  }
}

// x = tanh(x / 3) * 3 ; conv3x3(4 -> 64) + bias ; ReLU.  NCHW latents (fp32 / fp16 / bf16) -> NHWC bf16.
// One thread = one pixel x 8 output channels (36 inputs, 288 weights from shared memory).
template <typename TIn>
__global__ void taesd_conv_in_kernel(const TIn* __restrict__ lat, const float* __restrict__ w, const float* __restrict__ bias,
                                     bf16* __restrict__ out, int B, int H, int W) {
  __shared__ float sw[64 * 36];
  __shared__ float sb[64];
  for (int i = threadIdx.x; i < 64 * 36; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  pdl_sync();
  const long total = (long)B * H * W * 8;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx & 7);
    const long pix = idx >> 3;
    const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((long)W * H));
    float in[36];
#pragma unroll
    for (int ci = 0; ci < 4; ++ci)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
        float v = 0.f;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          v = (float)lat[(((long)b * 4 + ci) * H + yy) * W + xx];
          v = tanhf(v * (1.f / 3.f)) * 3.f;
        }
        in[ci * 9 + t] = v;
      }
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const float* wr = sw + (c8 * 8 + o) * 36;
      float a = sb[c8 * 8 + o];
#pragma unroll
      for (int i = 0; i < 36; ++i) a = fmaf(in[i], wr[i], a);
      acc[o] = fmaxf(a, 0.f);
    }
    uint4 v4;
    v4.x = pack_bf16(acc[0], acc[1]); v4.y = pack_bf16(acc[2], acc[3]); v4.z = pack_bf16(acc[4], acc[5]); v4.w = pack_bf16(acc[6], acc[7]);
    reinterpret_cast<uint4*>(out + pix * 64)[c8] = v4;
  }
}

// nearest-neighbour x2 (nn.Upsample(scale_factor=2)): NHWC bf16, one thread = one 16-byte chunk of an output pixel
__global__ void taesd_upsample2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W) {
  pdl_sync();
  const long total = (long)B * 2 * H * 2 * W * 8;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx & 7);
    const long pix = idx >> 3;
    const int x = (int)(pix % (2 * W)), y = (int)((pix / (2 * W)) % (2 * H)), b = (int)(pix / ((long)4 * W * H));
    out[idx] = in[((((long)b * H + (y >> 1)) * W + (x >> 1)) << 3) + c8];
  }
}

// conv3x3(64 -> 3) + bias ; x * 2 - 1 (DecoderTiny's output scaling) ; optional clamp(-1, 1) * 0.5 + 0.5 (training.py:113-117).
// NHWC bf16 -> NCHW fp32.  One thread = one output pixel.
__global__ void taesd_conv_out_kernel(const bf16* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                      float* __restrict__ out, int B, int H, int W, int to_image) {
  __shared__ float sw[9 * 64 * 3];  // [tap][ci][co]
  for (int i = threadIdx.x; i < 9 * 64 * 3; i += blockDim.x) {
    const int co = i % 3, ci = (i / 3) % 64, t = i / 192;
    sw[i] = w[(co * 64 + ci) * 9 + t];
  }
  __syncthreads();
  pdl_sync();
  const long total = (long)B * H * W;
  for (long pix = (long)blockIdx.x * blockDim.x + threadIdx.x; pix < total; pix += (long)gridDim.x * blockDim.x) {
    const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((long)W * H));
    float a0 = bias[0], a1 = bias[1], a2 = bias[2];
    for (int t = 0; t < 9; ++t) {
      const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      const uint4* row = reinterpret_cast<const uint4*>(in + (((long)b * H + yy) * W + xx) * 64);
      const float* wt = sw + t * 192;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        const uint4 v = row[c8];
        const float2 f[4] = {unpack_bf16(v.x), unpack_bf16(v.y), unpack_bf16(v.z), unpack_bf16(v.w)};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float* w0 = wt + (c8 * 8 + 2 * i) * 3;
          a0 = fmaf(f[i].x, w0[0], a0); a1 = fmaf(f[i].x, w0[1], a1); a2 = fmaf(f[i].x, w0[2], a2);
          a0 = fmaf(f[i].y, w0[3], a0); a1 = fmaf(f[i].y, w0[4], a1); a2 = fmaf(f[i].y, w0[5], a2);
        }
      }
    }
    float r[3] = {a0 * 2.f - 1.f, a1 * 2.f - 1.f, a2 * 2.f - 1.f};
    if (to_image) {
#pragma unroll
      for (int c = 0; c < 3; ++c) r[c] = fminf(fmaxf(r[c], -1.f), 1.f) * 0.5f + 0.5f;
    }
    const long hw = (long)H * W;
    float* o = out + (long)b * 3 * hw + (long)y * W + x;
    o[0] = r[0]; o[hw] = r[1]; o[2 * hw] = r[2];
  }
}

static int ew_grid(long n, int threads) {
  long b = (n + threads - 1) / threads;
  const long cap = (long)num_sms() * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

static int conv64(const bf16* x, const void* w_packed, const float* bias, const bf16* res, bf16* out, int B, int H, int W,
                  int relu, cudaStream_t st) {
  ConvParams p;
  p.B = B; p.H = H; p.W = W;
  p.bw = W < 128 ? W : 128;
  p.bh = 128 / p.bw;
  p.tiles_x = (W + p.bw - 1) / p.bw;
  p.tiles_y = (H + p.bh - 1) / p.bh;
  p.bias = bias; p.res = res; p.out = out; p.relu = relu;
  CUtensorMap tmX, tmW;
  {
    uint64_t dims[4] = {64, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {128, (uint64_t)W * 128, (uint64_t)H * W * 128};
    uint32_t box[4] = {64, (uint32_t)p.bw, (uint32_t)p.bh, 1};
    int rc = make_tma_bf16(&tmX, x, 4, dims, strides, box, TMA_SWZ_128);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {64, 9 * 64};
    uint64_t strides[1] = {128};
    uint32_t box[2] = {64, 64};
    int rc = make_tma_bf16(&tmW, w_packed, 2, dims, strides, box, TMA_SWZ_128);
    if (rc) return rc;
  }
  const int smem = 1024 + TC_W_BYTES + TC_STAGES * TC_TILE_BYTES + 256 + 256;
  static bool attr = false;
  if (!attr) {
    TDM_CHECK_CUDA(cudaFuncSetAttribute(taesd_conv64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const int n_tiles = B * p.tiles_x * p.tiles_y;
  const int grid = n_tiles < num_sms() ? n_tiles : num_sms();
  count_launch();
  TDM_CHECK_CUDA(launch_pdl(taesd_conv64_kernel, dim3(grid), dim3(TC_THREADS), (size_t)smem, st, tmX, tmW, p));
  return TDM_OK;
}

}  // namespace tdm

using namespace tdm;

extern "C" int64_t tdm_taesd_ws_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return 3 * ((int64_t)B * 8 * H * 8 * W * 64 * 2 + 1024);
}

extern "C" int tdm_taesd_decode(const TdmTaesdWeights* w, const void* latents, int latent_dtype, int B, int H, int W,
                                float* out, int to_image, void* workspace, int64_t workspace_bytes, void* stream) {
  TDM_REQUIRE(w && latents && out && workspace, "tdm_taesd_decode: null argument");
  TDM_REQUIRE(B > 0 && H > 0 && W > 0, "tdm_taesd_decode: bad shape");
  TDM_REQUIRE((W & (W - 1)) == 0 && W >= 8, "tdm_taesd_decode: latent width %d must be a power of two >= 8", W);
  TDM_REQUIRE(workspace_bytes >= tdm_taesd_ws_bytes(B, H, W), "tdm_taesd_decode: workspace too small");
  TDM_REQUIRE(((uintptr_t)workspace & 1023) == 0, "tdm_taesd_decode: workspace must be 1024-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t per = ((int64_t)B * 8 * H * 8 * W * 64 * 2 + 1023) & ~int64_t(1023);
  bf16* buf[3] = {reinterpret_cast<bf16*>(workspace), reinterpret_cast<bf16*>((uint8_t*)workspace + per),
                  reinterpret_cast<bf16*>((uint8_t*)workspace + 2 * per)};
  // conv_in
  {
    const long n = (long)B * H * W * 8;
    count_launch();
    if (latent_dtype == TDM_F32)
      TDM_CHECK_CUDA(launch_pdl(taesd_conv_in_kernel<float>, dim3(ew_grid(n, 256)), dim3(256), 0, st, (const float*)latents, w->conv_in_w, w->conv_in_b, buf[0], B, H, W));
    else if (latent_dtype == TDM_BF16)
      TDM_CHECK_CUDA(launch_pdl(taesd_conv_in_kernel<bf16>, dim3(ew_grid(n, 256)), dim3(256), 0, st, (const bf16*)latents, w->conv_in_w, w->conv_in_b, buf[0], B, H, W));
    else if (latent_dtype == TDM_F16)
      TDM_CHECK_CUDA(launch_pdl(taesd_conv_in_kernel<__half>, dim3(ew_grid(n, 256)), dim3(256), 0, st, (const __half*)latents, w->conv_in_w, w->conv_in_b, buf[0], B, H, W));
    else
      TDM_REQUIRE(false, "tdm_taesd_decode: unsupported latent dtype %d", latent_dtype);
  }
  int cur = 0, h = H, wd = W, blk = 0;
  const int n_blocks[4] = {3, 3, 3, 1};
  for (int s = 0; s < 4; ++s) {
    for (int k = 0; k < n_blocks[s]; ++k, ++blk) {
      const int a = (cur + 1) % 3, b = (cur + 2) % 3;
      int rc;
      if ((rc = conv64(buf[cur], w->block_w[blk * 3 + 0], w->block_b[blk * 3 + 0], nullptr, buf[a], B, h, wd, 1, st))) return rc;
      if ((rc = conv64(buf[a], w->block_w[blk * 3 + 1], w->block_b[blk * 3 + 1], nullptr, buf[b], B, h, wd, 1, st))) return rc;
      if ((rc = conv64(buf[b], w->block_w[blk * 3 + 2], w->block_b[blk * 3 + 2], buf[cur], buf[a], B, h, wd, 1, st))) return rc;
      cur = a;
    }
    if (s < 3) {
      const int a = (cur + 1) % 3, b = (cur + 2) % 3;
      const long n = (long)B * 2 * h * 2 * wd * 8;
      count_launch();
      TDM_CHECK_CUDA(launch_pdl(taesd_upsample2_kernel, dim3(ew_grid(n, 256)), dim3(256), 0, st, (const uint4*)buf[cur], (uint4*)buf[a], B, h, wd));
      h *= 2; wd *= 2;
      int rc;
      if ((rc = conv64(buf[a], w->up_w[s], nullptr, nullptr, buf[b], B, h, wd, 0, st))) return rc;
      cur = b;
    }
  }
  count_launch();
  TDM_CHECK_CUDA(launch_pdl(taesd_conv_out_kernel, dim3(ew_grid((long)B * h * wd, 128)), dim3(128), 0, st, (const bf16*)buf[cur], w->conv_out_w,
                            w->conv_out_b, out, B, h, wd, to_image));
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}
