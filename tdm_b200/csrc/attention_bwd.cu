// Flash attention backward on tcgen05 / TMEM (head_dim <= 128), self- and cross-attention.
// One CTA owns a tile of 128 keys of one (batch, head) and loops over the query tiles:
//   S^T  = K Q^T          (recomputed)        dP^T = V dO^T
//   P^T  = exp(S^T*scale + bias - lse)        dS^T = P^T o (dP^T - delta)
//   dV  += P^T dO         dK += dS^T Q        dQ  = dS K   (fp32 atomics into a [B,Sq,H,d] scratch)
// thread == key row == TMEM lane for the softmax recompute; thread == query row for the dQ drain.
// Operand reuse without copies: the Q / dO tiles are read K-major for S^T / dP^T and MN-major for
// dK / dV; dS^T in shared memory is read K-major for dK and MN-major (as dS) for dQ.
// TMEM: S^T [0,128) (reused for dQ), dP^T [128,256), dV [256,384), dK [384,512).
#include <stdlib.h>
#include "common.cuh"

namespace tdm {

constexpr int CHUNK = 128 * 128;  // bytes of one [128 rows x 64 bf16] swizzled chunk

int make_head_map(CUtensorMap* map, const void* ptr, int B, int S, int H, int d, long stride_b, long stride_s,
                  long stride_h);
int make_head_map64(CUtensorMap* map, const void* ptr, int B, int S, int H, int d, long stride_b, long stride_s,
                    long stride_h);  // attention_fwd2.cu: box 32 x 128, 64-byte swizzle
constexpr int CH64 = 128 * 64;  // one [128 rows x 32 bf16] chunk, 64-byte swizzle

struct AttnBwdParams {
  int B, H, Sq, Sk, d, nd, dk16;
  int nc;  // ceil(d / 32): chunks of the 64-byte-swizzled tiles (attn_bwd_kernel)
  int dq_stage_bytes;  // attn_bwd_kernel: bytes of the dQ staging area (column passes, or the bf16 tile of the direct path)
  float scale, scale_log2;
  const float* bias;   // [B, Sk] or null
  const float* lse;    // [B, H, Sq]
  const float* delta;  // [B, H, Sq]
  float* dq_accum;     // fp32 [B, Sq, H, d] contiguous
  bf16* dK;
  long dk_stride_b, dk_stride_s, dk_stride_h;
  bf16* dV;
  long dv_stride_b, dv_stride_s, dv_stride_h;
  int dq_direct;     // one key tile per (batch, head) (cross-attention, Sk <= 128): every dQ tile is produced exactly once --
                     // it is stored as bf16 straight into dQ (no zero fill, no fp32 accumulation, no cast kernel)
  long long* trace;  // TDM_B200_ATTN_TRACE: clock64 stamps of CTA (0,0,0), [query tile][16 events]
};

// delta[b,h,q] = sum_d dO[b,q,h,d] * O[b,q,h,d]; the same kernel zero-fills the fp32 dQ accumulator (layout
// [B*Sq, H*d]), which saves a separate memset node between the kernels.
// Row kernel (heads contiguous: stride_h == d, the only layout the DiT produces): one warp per (b, q) token row of
// H*d values.  Lanes read consecutive 16-byte units of O and dO (fully coalesced), the per-unit partial dots meet in
// shared memory and one lane per head adds the d/8 partials of its head; the row of the dQ accumulator is zeroed
// with consecutive float4 stores.  (The first version ran one THREAD per (b, q, h): 144-byte strides between lanes,
// 27 us for 76 MB.)
__global__ void __launch_bounds__(256) attn_delta_rows_kernel(const bf16* __restrict__ O, const bf16* __restrict__ dO, long sb, long ss,
                                                             float* __restrict__ delta, float* __restrict__ dq_accum, int B, int H,
                                                             int Sq, int d) {
  extern __shared__ float part[];  // [8 warps][H * d / 8]
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * 8 + warp;  // b * Sq + q
  if (row >= (long)B * Sq) return;
  const int b = int(row / Sq), q = int(row - (long)b * Sq);
  const int units = H * d / 8, upb = d / 8;
  const uint4* o = reinterpret_cast<const uint4*>(O + (long)b * sb + (long)q * ss);
  const uint4* g = reinterpret_cast<const uint4*>(dO + (long)b * sb + (long)q * ss);
  float* pw = part + warp * units;
  for (int u = lane; u < units; u += 32) {
    const uint4 a = o[u], c = g[u];
    const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
    const float2 c0 = unpack_bf16(c.x), c1 = unpack_bf16(c.y), c2 = unpack_bf16(c.z), c3 = unpack_bf16(c.w);
    pw[u] = a0.x * c0.x + a0.y * c0.y + a1.x * c1.x + a1.y * c1.y + a2.x * c2.x + a2.y * c2.y + a3.x * c3.x + a3.y * c3.y;
  }
  if (dq_accum) {  // (null: the main kernel writes dQ directly, nothing to zero)
    float4* z = reinterpret_cast<float4*>(dq_accum + row * (long)(H * d));
    for (int i = lane; i < H * d / 4; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncwarp();
  for (int h = lane; h < H; h += 32) {
    float acc = 0.f;
    for (int i = 0; i < upb; ++i) acc += pw[h * upb + i];
    delta[((long)b * H + h) * Sq + q] = acc;
  }
}

// General strides: one thread per (b, q, h).
__global__ void attn_delta_kernel(const bf16* __restrict__ O, const bf16* __restrict__ dO, long sb, long ss, long sh,
                                  float* __restrict__ delta, float* __restrict__ dq_accum, int B, int H, int Sq, int d) {
  pdl_sync();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)B * Sq * H;
  if (idx >= total) return;
  const int h = int(idx % H);
  const int q = int((idx / H) % Sq);
  const int b = int(idx / ((long)H * Sq));
  const long off = (long)b * sb + (long)q * ss + (long)h * sh;
  float acc = 0.f;
  for (int c = 0; c < d; c += 8) {
    const uint4 o = *reinterpret_cast<const uint4*>(O + off + c);
    const uint4 g = *reinterpret_cast<const uint4*>(dO + off + c);
    const float2 o0 = unpack_bf16(o.x), o1 = unpack_bf16(o.y), o2 = unpack_bf16(o.z), o3 = unpack_bf16(o.w);
    const float2 g0 = unpack_bf16(g.x), g1 = unpack_bf16(g.y), g2 = unpack_bf16(g.z), g3 = unpack_bf16(g.w);
    acc += o0.x * g0.x + o0.y * g0.y + o1.x * g1.x + o1.y * g1.y + o2.x * g2.x + o2.y * g2.y + o3.x * g3.x + o3.y * g3.y;
  }
  delta[((long)b * H + h) * Sq + q] = acc;
  if (dq_accum) {
    float4* z = reinterpret_cast<float4*>(dq_accum + (((long)b * Sq + q) * H + h) * d);  // d % 8 == 0: 16-byte aligned
    for (int c = 0; c < d / 4; ++c) z[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(192, 1)
attn_bwd_wide_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tile_bytes = p.nd * CHUNK;
  uint8_t* sK = smem;
  uint8_t* sV = sK + tile_bytes;
  uint8_t* sQ = sV + tile_bytes;
  uint8_t* sdO = sQ + tile_bytes;
  uint8_t* sPt = sdO + tile_bytes;   // [128 keys x 128 queries] bf16, 2 chunks
  uint8_t* sdSt = sPt + 2 * CHUNK;
  float* sLse = reinterpret_cast<float*>(sdSt + 2 * CHUNK);  // 128 (log2 units)
  float* sDelta = sLse + 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 128);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;
  uint64_t* s_full = bars + 2;
  uint64_t* pds_full = bars + 3;   // 128 arrivals
  uint64_t* mma_done = bars + 4;
  uint64_t* dq_drained = bars + 5; // 128 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int n_q = (p.Sq + 127) / 128;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    mbar_init(qdo_full, 1);
    mbar_init(s_full, 1);
    mbar_init(pds_full, 128);
    mbar_init(mma_done, 1);
    mbar_init(dq_drained, 128);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_St = tmem_base, tm_dPt = tmem_base + 128, tm_dV = tmem_base + 256, tm_dK = tmem_base + 384;

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * tile_bytes);
      for (int c = 0; c < p.nd; ++c) {
        tma_load_4d(sK + c * CHUNK, &tmK, kv_full, c * 64, k0, h, b);
        tma_load_4d(sV + c * CHUNK, &tmV, kv_full, c * 64, k0, h, b);
      }
      for (int i = 0; i < n_q; ++i) {
        if (i > 0) mbar_wait(mma_done, (i - 1) & 1);  // Q/dO tiles of the previous iteration fully consumed
        mbar_arrive_expect_tx(qdo_full, 2 * tile_bytes);
        for (int c = 0; c < p.nd; ++c) {
          tma_load_4d(sQ + c * CHUNK, &tmQ, qdo_full, c * 64, i * 128, h, b);
          tma_load_4d(sdO + c * CHUNK, &tmdO, qdo_full, c * 64, i * 128, h, b);
        }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const int nd64 = p.nd * 64;
      const uint32_t id_kk = make_idesc_bf16(128, 128, 0, 0);    // S^T, dP^T : A K-major, B K-major
      const uint32_t id_kmn = make_idesc_bf16(128, nd64, 0, 1);  // dV, dK    : A K-major, B MN-major
      const uint32_t id_mnmn = make_idesc_bf16(128, nd64, 1, 1); // dQ        : A MN-major, B MN-major
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), adO = smem_u32(sdO);
      const uint32_t aPt = smem_u32(sPt), adSt = smem_u32(sdSt);
      mbar_wait(kv_full, 0);
      for (int i = 0; i < n_q; ++i) {
        mbar_wait(qdo_full, i & 1);
        if (i > 0) mbar_wait(dq_drained, (i - 1) & 1);  // S^T columns are reused by dQ
        tc_fence_after();
        for (int k = 0; k < p.dk16; ++k) {
          const uint32_t off = (k >> 2) * CHUNK + (k & 3) * 32;
          umma_bf16(tm_St, make_sdesc_sw128(aK + off, 16, 1024), make_sdesc_sw128(aQ + off, 16, 1024), id_kk, k != 0);
        }
        for (int k = 0; k < p.dk16; ++k) {
          const uint32_t off = (k >> 2) * CHUNK + (k & 3) * 32;
          umma_bf16(tm_dPt, make_sdesc_sw128(aV + off, 16, 1024), make_sdesc_sw128(adO + off, 16, 1024), id_kk, k != 0);
        }
        umma_commit(s_full);
        mbar_wait(pds_full, i & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // contraction over the 128 queries of this tile
          const uint32_t offA = (k >> 2) * CHUNK + (k & 3) * 32;
          umma_bf16(tm_dV, make_sdesc_sw128(aPt + offA, 16, 1024), make_sdesc_sw128(adO + k * 2048, CHUNK, 1024),
                    id_kmn, (i | k) != 0);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t offA = (k >> 2) * CHUNK + (k & 3) * 32;
          umma_bf16(tm_dK, make_sdesc_sw128(adSt + offA, 16, 1024), make_sdesc_sw128(aQ + k * 2048, CHUNK, 1024),
                    id_kmn, (i | k) != 0);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // dQ = dS K: contraction over the 128 keys; A = dS^T read MN-major
          umma_bf16(tm_St, make_sdesc_sw128(adSt + k * 2048, CHUNK, 1024), make_sdesc_sw128(aK + k * 2048, CHUNK, 1024),
                    id_mnmn, k != 0);
        }
        umma_commit(mma_done);
      }
    }
  } else {
    // ===================== compute warps (128 threads) =====================
    const int tid = threadIdx.x;
    const uint32_t lane_off = uint32_t(warp * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    const int kv = k0 + tid;
    float bias_l2 = 0.f;
    if (kv >= p.Sk) bias_l2 = -INFINITY;
    else if (p.bias) bias_l2 = p.bias[(long)b * p.Sk + kv] * LOG2E;
    const long bh = (long)b * p.H + h;
    for (int i = 0; i < n_q; ++i) {
      const int q = i * 128 + tid;
      sLse[tid] = q < p.Sq ? p.lse[bh * p.Sq + q] * LOG2E : INFINITY;
      sDelta[tid] = q < p.Sq ? p.delta[bh * p.Sq + q] : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      uint8_t* prow = sPt + tid * 128;
      uint8_t* drow = sdSt + tid * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t rs[32], rp[32];
        tmem_ld32(tm_St + lane_off + c * 32, rs);
        tmem_ld32(tm_dPt + lane_off + c * 32, rp);
        tmem_ld_wait();
        float pv[32], ds[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float pj = exp2f(fmaf(__uint_as_float(rs[j]), p.scale_log2, bias_l2) - sLse[c * 32 + j]);
          pv[j] = pj;
          ds[j] = pj * (__uint_as_float(rp[j]) - sDelta[c * 32 + j]);
        }
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {
          const int U = c * 4 + u4;
          const int off = (U >> 3) * CHUNK + (((U & 7) ^ (tid & 7)) << 4);
          uint4 w;
          w.x = pack_bf16(pv[u4 * 8 + 0], pv[u4 * 8 + 1]); w.y = pack_bf16(pv[u4 * 8 + 2], pv[u4 * 8 + 3]);
          w.z = pack_bf16(pv[u4 * 8 + 4], pv[u4 * 8 + 5]); w.w = pack_bf16(pv[u4 * 8 + 6], pv[u4 * 8 + 7]);
          *reinterpret_cast<uint4*>(prow + off) = w;
          w.x = pack_bf16(ds[u4 * 8 + 0], ds[u4 * 8 + 1]); w.y = pack_bf16(ds[u4 * 8 + 2], ds[u4 * 8 + 3]);
          w.z = pack_bf16(ds[u4 * 8 + 4], ds[u4 * 8 + 5]); w.w = pack_bf16(ds[u4 * 8 + 6], ds[u4 * 8 + 7]);
          *reinterpret_cast<uint4*>(drow + off) = w;
        }
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(pds_full);
      // drain dQ for this query tile (thread == query row)
      mbar_wait(mma_done, i & 1);
      tc_fence_after();
      float* dq_row = p.dq_accum + (((long)b * p.Sq + q) * p.H + h) * p.d;
      for (int c = 0; c < p.d; c += 8) {
        uint32_t r[8];
        tmem_ld8(tm_St + lane_off + c, r);
        tmem_ld_wait();
        if (q < p.Sq) {
#pragma unroll
          for (int j = 0; j < 8; ++j) atomicAdd(dq_row + c + j, __uint_as_float(r[j]) * p.scale);
        }
      }
      tc_fence_before();
      mbar_arrive(dq_drained);
    }
    // epilogue: dV, dK for this key tile (thread == key row)
    bf16* dv_row = p.dV + (long)b * p.dv_stride_b + (long)kv * p.dv_stride_s + (long)h * p.dv_stride_h;
    bf16* dk_row = p.dK + (long)b * p.dk_stride_b + (long)kv * p.dk_stride_s + (long)h * p.dk_stride_h;
    for (int c = 0; c < p.d; c += 8) {
      uint32_t rv[8], rk[8];
      tmem_ld8(tm_dV + lane_off + c, rv);
      tmem_ld8(tm_dK + lane_off + c, rk);
      tmem_ld_wait();
      if (kv < p.Sk) {
        uint4 w;
        w.x = pack_bf16(__uint_as_float(rv[0]), __uint_as_float(rv[1]));
        w.y = pack_bf16(__uint_as_float(rv[2]), __uint_as_float(rv[3]));
        w.z = pack_bf16(__uint_as_float(rv[4]), __uint_as_float(rv[5]));
        w.w = pack_bf16(__uint_as_float(rv[6]), __uint_as_float(rv[7]));
        *reinterpret_cast<uint4*>(dv_row + c) = w;
        w.x = pack_bf16(__uint_as_float(rk[0]) * p.scale, __uint_as_float(rk[1]) * p.scale);
        w.y = pack_bf16(__uint_as_float(rk[2]) * p.scale, __uint_as_float(rk[3]) * p.scale);
        w.z = pack_bf16(__uint_as_float(rk[4]) * p.scale, __uint_as_float(rk[5]) * p.scale);
        w.w = pack_bf16(__uint_as_float(rk[6]) * p.scale, __uint_as_float(rk[7]) * p.scale);
        *reinterpret_cast<uint4*>(dk_row + c) = w;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Pipelined backward for head_dim <= 80 (PixArt: 72).  Same math as above, but
//   * dQ has its own TMEM columns and is drained by a SECOND warpgroup, so the drain overlaps the next query tile: each
//     drain warp stages its 32 query rows (fp32) in shared memory and ONE TMA reduce-add per warp and tile adds them into
//     the accumulator.  (Thread-per-row red.global.add.v4.f32 -- 32 rows = 32 transactions per instruction, 2300 per tile
//     -- clogged the SM's load/store queue, which every mbarrier wait and tcgen05.ld of the CTA shares; found in the
//     forward kernel, attention_fwd2.cu.)  dK / dV leave through shared memory and TMA stores for the same reason;
//   * Q / K / V / dO tiles use rows of 64 bytes (32 head-dim columns, 64-byte swizzle): 24 KB per tile for head_dim 72
//     instead of 32 KB, which is what makes room for the dQ staging buffer;
//   * Q / dO tiles are double-buffered in shared memory and the TMA producer runs one tile ahead;
//   * P^T (bf16) is written back into the S^T columns and feeds dV += P^T dO straight from tensor memory;
//   * dP^T of the next tile is issued as soon as the compute warps have pulled the current S^T / dP^T into
//     registers, S^T of the next tile right after dV (tcgen05.mma executes in issue order).
// TMEM: S^T/P^T [0,128)  dP^T [128,256)  dV [256,336)  dK [336,416)  dQ [416,496).
// warps 0-7: P / dS (thread == key row; two halves of the query columns)   warps 8-11: dQ drain (thread == query row)
// warp 12: TMA   warp 13: MMA
// ------------------------------------------------------------------------------------------------
constexpr int BWD_THREADS = 448;  // 8 compute warps (two halves of the query columns) + 4 drain warps + TMA + MMA

#define BW_TRACE(i_, e_)                                                                       \
  do {                                                                                         \
    if constexpr (TRACE) {                                                                     \
      if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (i_) < 32) p.trace[(i_) * 16 + (e_)] = clock64(); \
    }                                                                                          \
  } while (0)
template <bool TRACE>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                const __grid_constant__ CUtensorMap tmdQ, const __grid_constant__ CUtensorMap tmdK,
                const __grid_constant__ CUtensorMap tmdV, const __grid_constant__ CUtensorMap tmdQ16,
                const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // swizzled tiles need 1024-byte alignment
  const int tile_bytes = p.nc * CH64;
  uint8_t* sK = smem;
  uint8_t* sV = sK + tile_bytes;
  uint8_t* sQ = sV + tile_bytes;          // 2 stages: Q(i) is held until dK(i); the reload (~2000 clk) is needed at S^T(i+2)
  uint8_t* sdO = sQ + 2 * tile_bytes;     // 2 stages with their own barriers: dO(i) is released right behind dV(i), so the
                                          // reload is back for dP^T(i+2), which is issued a phase earlier than S^T(i+2)
  uint8_t* sdSt = sdO + 2 * tile_bytes;   // [128 keys x 128 queries] bf16, 2 chunks of 64 columns, 128-byte swizzle
  float* sdQ = reinterpret_cast<float*>(sdSt + 2 * CHUNK);   // [128 queries][d] fp32 staging of the dQ tile (bf16 on the direct path)
  float* sLse = sdQ + p.dq_stage_bytes / 4;                  // 2 x 128 (log2 units)
  float* sDelta = sLse + 256;                                 // 2 x 128
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDelta + 256);
  uint64_t* kv_full = bars;        // 1
  uint64_t* q_full = bars + 1;     // [2] (+1 spare)
  uint64_t* q_empty = bars + 4;    // [2] (+1 spare)
  uint64_t* do_full = bars + 7;    // [2]
  uint64_t* do_empty = bars + 9;   // [2]
  uint64_t* st_full = bars + 11;   // S^T of tile i complete
  uint64_t* dp_free = bars + 12;   // 8: compute warps hold dP^T of tile i in registers (its region takes S^T of tile i+1)
  uint64_t* ds_full = bars + 13;   // 8: dS^T (smem) of tile i written
  uint64_t* dq_full = bars + 14;   // all MMAs of tile i complete (dQ ready, dS^T smem free)
  uint64_t* dq_free = bars + 15;   // 4: dQ of tile i drained
  // dP^T of tile i complete: ONE BARRIER PER TILE PARITY.  dP^T(i+1) is issued as soon as P^T(i) is written, i.e. possibly
  // before a compute warp has started to wait for dP^T(i): on a single barrier the awaited parity would flip back and the
  // wait never return (seen as a hang once in ~100 000 launches; same rule as attention_fwd.cu's p_full).
  uint64_t* dpt_full = bars + 16;  // [2]
  uint64_t* p_full = bars + 18;    // 8: P^T of tile i written to tensor memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int n_q = (p.Sq + 127) / 128;
  const int dpad = p.dk16 * 16;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&do_full[i], 1);
      mbar_init(&do_empty[i], 1);
    }
    mbar_init(st_full, 1);
    mbar_init(&dpt_full[0], 1);
    mbar_init(&dpt_full[1], 1);
    mbar_init(dp_free, 8);   // one elected arrival per compute warp (per-thread arrivals serialise on the barrier word)
    mbar_init(p_full, 8);
    mbar_init(ds_full, 8);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 4);
    fence_barrier_init();
  }
  if (warp == 13) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Two 128-column regions R0 / R1 swap roles every tile: S^T(i) / P^T(i) live in R[i & 1], dP^T(i) in R[(i + 1) & 1].
  const uint32_t tm_R = tmem_base, tm_dV = tmem_base + 256, tm_dK = tmem_base + 336, tm_dQ = tmem_base + 416;
  pdl_sync();

  if (warp == 12) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmdO);
      mbar_arrive_expect_tx(kv_full, 2 * tile_bytes);
      for (int c = 0; c < p.nc; ++c) {
        tma_load_4d(sK + c * CH64, &tmK, kv_full, c * 32, k0, h, b);
        tma_load_4d(sV + c * CH64, &tmV, kv_full, c * 32, k0, h, b);
      }
      for (int i = 0; i < n_q; ++i) {
        const int sd = i & 1, sq = i & 1;
        const uint32_t qph = (i >> 1) & 1;
        mbar_wait(&do_empty[sd], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&do_full[sd], tile_bytes);
        for (int c = 0; c < p.nc; ++c) tma_load_4d(sdO + sd * tile_bytes + c * CH64, &tmdO, &do_full[sd], c * 32, i * 128, h, b);
        mbar_wait(&q_empty[sq], qph ^ 1);
        mbar_arrive_expect_tx(&q_full[sq], tile_bytes);
        for (int c = 0; c < p.nc; ++c) tma_load_4d(sQ + sq * tile_bytes + c * CH64, &tmQ, &q_full[sq], c * 32, i * 128, h, b);
      }
    }
  } else if (warp == 13) {
    // ===================== MMA issuer =====================
    // Warp-convergent: every lane polls and computes the (uniform) descriptors, tcgen05 ops under elect.sync only.
    const uint32_t id_kk = make_idesc_bf16(128, 128, 0, 0);    // S^T, dP^T : A K-major, B K-major
    const uint32_t id_kmn = make_idesc_bf16(128, dpad, 0, 1);  // dV, dK    : A K-major (TMEM / smem), B MN-major
    const uint32_t id_mnmn = make_idesc_bf16(128, dpad, 1, 1); // dQ        : A MN-major, B MN-major
    // Q / K / V / dO tiles: 64-byte swizzle.  K-major: 8-row groups 512 bytes apart, two 16-column k-steps per 32-column
    // chunk; MN-major: chunks CH64 apart, one k-step = 16 rows of 64 bytes.
    const uint64_t dK_k = make_sdesc_sw64(smem_u32(sK), 16, 512);       // K as K-major operand (S^T)
    const uint64_t dV_k = make_sdesc_sw64(smem_u32(sV), 16, 512);       // V as K-major operand (dP^T)
    const uint64_t dK_mn = make_sdesc_sw64(smem_u32(sK), CH64, 512);    // K as MN-major operand (dQ)
    const uint64_t dQ_k0 = make_sdesc_sw64(smem_u32(sQ), 16, 512);
    const uint64_t dQ_mn0 = make_sdesc_sw64(smem_u32(sQ), CH64, 512);
    const uint64_t ddO_k0 = make_sdesc_sw64(smem_u32(sdO), 16, 512);
    const uint64_t ddO_mn0 = make_sdesc_sw64(smem_u32(sdO), CH64, 512);
    const uint64_t ddS_k = make_sdesc_sw128(smem_u32(sdSt), 16, 1024);
    const uint64_t ddS_mn = make_sdesc_sw128(smem_u32(sdSt), CHUNK, 1024);
    const uint32_t stage_step = uint32_t(tile_bytes) >> 4;
    const int dk16 = p.dk16;
    auto issue_st = [&](int i, int qs) {  // caller holds the election; qs = Q stage of tile i
      const uint64_t qd = dQ_k0 + uint64_t(qs * stage_step);
      const uint32_t dst = tm_R + (i & 1) * 128;
#pragma unroll
      for (int k = 0; k < 5; ++k)
        if (k < dk16) {
          const uint32_t off = ((k >> 1) * CH64 + (k & 1) * 32) >> 4;
          umma_bf16(dst, dK_k + off, qd + off, id_kk, k != 0);
        }
    };
    auto issue_dpt = [&](int i) {
      const uint64_t od = ddO_k0 + uint64_t((i & 1) * stage_step);
      const uint32_t dst = tm_R + ((i + 1) & 1) * 128;
#pragma unroll
      for (int k = 0; k < 5; ++k)
        if (k < dk16) {
          const uint32_t off = ((k >> 1) * CH64 + (k & 1) * 32) >> 4;
          umma_bf16(dst, dV_k + off, od + off, id_kk, k != 0);
        }
    };
    // Per tile the compute warps run two phases -- A: P^T from S^T (exponentials), B: dS^T from dP^T -- and this warp
    // answers three events in the order they occur:
    //   P^T(i) written        -> dV(i), then dP^T(i+1) into the region that held S^T(i) / P^T(i) (in order behind dV(i))
    //   dP^T(i) in registers  -> S^T(i+1) into the region that held dP^T(i)
    //   dS^T(i) written       -> dK(i), dQ(i)
    // so S^T(i+1) is complete before the compute warps finish tile i and dP^T(i+1) long before they need it: the tensor
    // work of a tile runs under the compute phases instead of between them (the first form produced S^T(i+1) and
    // dP^T(i+1) only after P^T / dS^T of tile i: compute 2560 clk + tensor / hand-off 1350 clk per tile, serial).
    mbar_wait(kv_full, 0);
    mbar_wait(&q_full[0], 0);
    mbar_wait(&do_full[0], 0);
    tc_fence_after();
    if (elect_one()) {
      issue_st(0, 0);
      umma_commit(st_full);
      issue_dpt(0);
      umma_commit(&dpt_full[0]);
    }
    __syncwarp();
    for (int i = 0; i < n_q; ++i) {
      const int sd = i & 1, qs = i & 1;
      const bool more = i + 1 < n_q;
      const int qs1 = (i + 1) & 1;
      const uint32_t qph1 = ((i + 1) >> 1) & 1;
      if (lane == 0) BW_TRACE(i, 0);
      mbar_wait(p_full, i & 1);
      if (more) mbar_wait(&do_full[(i + 1) & 1], ((i + 1) >> 1) & 1);
      tc_fence_after();
      if (lane == 0) BW_TRACE(i, 1);
      if (elect_one()) {
        const uint64_t od_mn = ddO_mn0 + uint64_t(sd * stage_step);
        const uint32_t pt = tm_R + (i & 1) * 128;
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dV += P^T dO : contraction over the 128 queries, A from tensor memory
          umma_bf16_ts(tm_dV, pt + (k >> 2) * 64 + (k & 3) * 8, od_mn + uint64_t(k * (1024 >> 4)), id_kmn, (i | k) != 0);
        umma_commit(&do_empty[sd]);  // dV(i) is the last reader of dO(i)
        if (more) {
          issue_dpt(i + 1);  // overwrites P^T(i): ordered after dV(i) by the in-order tensor pipe
          umma_commit(&dpt_full[(i + 1) & 1]);
        }
      }
      __syncwarp();
      if (more) {
        mbar_wait(dp_free, i & 1);
        mbar_wait(&q_full[qs1], qph1);
        tc_fence_after();
        if (elect_one()) {
          issue_st(i + 1, qs1);
          umma_commit(st_full);
        }
        __syncwarp();
      }
      if (lane == 0) BW_TRACE(i, 2);
      mbar_wait(ds_full, i & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t qd_mn = dQ_mn0 + uint64_t(qs * stage_step);
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // dK += dS^T Q
          const uint32_t offA = ((k >> 2) * CHUNK + (k & 3) * 32) >> 4;
          umma_bf16(tm_dK, ddS_k + offA, qd_mn + uint64_t(k * (1024 >> 4)), id_kmn, (i | k) != 0);
        }
        umma_commit(&q_empty[qs]);  // dK(i) is the last reader of Q(i)
      }
      __syncwarp();
      if (i > 0) {
        mbar_wait(dq_free, (i - 1) & 1);
        tc_fence_after();
      }
      if (lane == 0) BW_TRACE(i, 3);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)  // dQ = dS K : contraction over the 128 keys; A = dS^T read MN-major
          umma_bf16(tm_dQ, ddS_mn + uint64_t(k * (2048 >> 4)), dK_mn + uint64_t(k * (1024 >> 4)), id_mnmn, k != 0);
        umma_commit(dq_full);
      }
      __syncwarp();
      if (lane == 0) BW_TRACE(i, 4);
    }
  } else if (warp < 8) {
    // ===================== P / dS compute (2 x 128 threads, thread == key row) =====================
    // warps 0-3 own query columns [0,64) of every tile, warps 4-7 columns [64,128): two warps per SM sub-partition
    // hide each other's TMEM / MUFU / shared-memory latencies.
    const int tid = threadIdx.x & 127;
    const int wg = threadIdx.x >> 7;
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    const int kv = k0 + tid;
    float bias_l2 = 0.f;
    if (kv >= p.Sk) bias_l2 = -INFINITY;
    else if (p.bias) bias_l2 = p.bias[(long)b * p.Sk + kv] * LOG2E;
    const long bh = (long)b * p.H + h;
    uint8_t* drow = sdSt + tid * 128;
    // lse (warpgroup 0) / delta (warpgroup 1) of the tile's 128 queries: loaded one tile AHEAD into a register -- the
    // global-memory round trip (~1000 clk) sat between two tiles on the loop's critical path (clock64 trace)
    const float* stat_g = (wg == 0 ? p.lse : p.delta) + bh * p.Sq;
    const float stat_oob = wg == 0 ? INFINITY : 0.f;
    float stat_next = tid < p.Sq ? stat_g[tid] : stat_oob;
    for (int i = 0; i < n_q; ++i) {
      float* lse_s = sLse + (i & 1) * 128;
      float* del_s = sDelta + (i & 1) * 128;
      if (wg == 0) lse_s[tid] = stat_next * LOG2E;
      else del_s[tid] = stat_next;
      {
        const int qn = (i + 1) * 128 + tid;
        stat_next = (i + 1 < n_q && qn < p.Sq) ? stat_g[qn] : stat_oob;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 0) BW_TRACE(i, 5);
      const uint32_t tm_S = tm_R + (i & 1) * 128 + lane_off, tm_dP = tm_R + ((i + 1) & 1) * 128 + lane_off;
      // ---- phase A: P^T = exp2(S^T * scale + bias - lse) -> packed bf16 into this warpgroup's OWN half of the S^T columns
      //      ([0,32) / [64,96): never passes the columns it still has to read, never touches the other warpgroup's) ----
      mbar_wait(st_full, i & 1);
      tc_fence_after();
      if (threadIdx.x == 0) BW_TRACE(i, 6);
      float pf[64];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = wg * 2 + cc;
        uint32_t rs[32];
        tmem_ld32(tm_S + c * 32, rs);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float4 l4 = *reinterpret_cast<const float4*>(lse_s + c * 32 + 2 * j);
          const float p0 = ex2_fast(fmaf(__uint_as_float(rs[2 * j]), p.scale_log2, bias_l2) - l4.x);
          const float p1 = ex2_fast(fmaf(__uint_as_float(rs[2 * j + 1]), p.scale_log2, bias_l2) - l4.y);
          const float p2 = ex2_fast(fmaf(__uint_as_float(rs[2 * j + 2]), p.scale_log2, bias_l2) - l4.z);
          const float p3 = ex2_fast(fmaf(__uint_as_float(rs[2 * j + 3]), p.scale_log2, bias_l2) - l4.w);
          pf[cc * 32 + 2 * j] = p0; pf[cc * 32 + 2 * j + 1] = p1; pf[cc * 32 + 2 * j + 2] = p2; pf[cc * 32 + 2 * j + 3] = p3;
          pk[j] = pack_bf16(p0, p1);
          pk[j + 1] = pack_bf16(p2, p3);
        }
        tmem_st16(tm_S + wg * 64 + cc * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (threadIdx.x == 0) BW_TRACE(i, 7);
      // ---- phase B: dS^T = P^T o (dP^T - delta) -> shared memory (K-major for dK, MN-major for dQ) ----
      mbar_wait(&dpt_full[i & 1], (i >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = wg * 2 + cc;
        uint32_t rp[32];
        tmem_ld32(tm_dP + c * 32, rp);
        tmem_ld_wait();
        if (cc == 1) {  // dP^T(i) is in registers: its region may take S^T(i+1)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dp_free);
        }
        uint32_t dk[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float4 d4 = *reinterpret_cast<const float4*>(del_s + c * 32 + 2 * j);
          dk[j] = pack_bf16(pf[cc * 32 + 2 * j] * (__uint_as_float(rp[2 * j]) - d4.x),
                            pf[cc * 32 + 2 * j + 1] * (__uint_as_float(rp[2 * j + 1]) - d4.y));
          dk[j + 1] = pack_bf16(pf[cc * 32 + 2 * j + 2] * (__uint_as_float(rp[2 * j + 2]) - d4.z),
                                pf[cc * 32 + 2 * j + 3] * (__uint_as_float(rp[2 * j + 3]) - d4.w));
        }
        if (threadIdx.x == 0 && cc == 0) BW_TRACE(i, 8);
        if (cc == 0 && i > 0) mbar_wait(dq_full, (i - 1) & 1);  // dK/dQ of the previous tile no longer read dS^T
#pragma unroll
        for (int u4 = 0; u4 < 4; ++u4) {
          const int U = c * 4 + u4;
          const int off = (U >> 3) * CHUNK + (((U & 7) ^ (tid & 7)) << 4);
          *reinterpret_cast<uint4*>(drow + off) = make_uint4(dk[u4 * 4], dk[u4 * 4 + 1], dk[u4 * 4 + 2], dk[u4 * 4 + 3]);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
      if (threadIdx.x == 0) BW_TRACE(i, 9);
    }
    // epilogue: warps 0-3 write dV, warps 4-7 write dK of this key tile (thread == key row)
    mbar_wait(dq_full, (n_q - 1) & 1);
    tc_fence_after();
    // dV (warps 0-3) / dK (warps 4-7): each warp stages its 32 key rows as bf16 in the Q buffers (free: dq_full of the
    // last tile means every MMA has retired) and stores them with one TMA store; rows >= Sk are clipped by the map.
    const uint32_t tm_src = wg == 0 ? tm_dV : tm_dK;
    const float osc = wg == 0 ? 1.f : p.scale;
    uint8_t* stage = sQ + wg * tile_bytes + (warp & 3) * (32 * p.d * 2);
    uint8_t* srow = stage + lane * (p.d * 2);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {  // d <= 80: 48 + 32 columns, one tcgen05.ld round each
      uint32_t r[3][16];
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (3 * hh + i < p.dk16) tmem_ld16(tm_src + lane_off + 16 * (3 * hh + i), r[i]);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        if (8 * (6 * hh + i) < p.d) {
#define BW_E(k_) (__uint_as_float(r[i >> 1][(i & 1) * 8 + (k_)]) * osc)
          uint4 w;
          w.x = pack_bf16(BW_E(0), BW_E(1));
          w.y = pack_bf16(BW_E(2), BW_E(3));
          w.z = pack_bf16(BW_E(4), BW_E(5));
          w.w = pack_bf16(BW_E(6), BW_E(7));
#undef BW_E
          *reinterpret_cast<uint4*>(srow + 16 * (6 * hh + i)) = w;
        }
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_4d(wg == 0 ? &tmdV : &tmdK, stage, 0, k0 + (warp & 3) * 32, h, b);
      tma_store_commit();
      tma_store_wait_all();
    }
  } else {
    // ===================== dQ drain (128 threads, thread == query row) =====================
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    float* stage = sdQ + (warp & 3) * 32 * p.d;  // this warp's 32 query rows, dense [32][d] fp32
    float* srow = stage + lane * p.d;
    if (lane == 0) tma_prefetch_desc(&tmdQ);
    for (int i = 0; i < n_q; ++i) {
      mbar_wait_relaxed(dq_full, i & 1);
      tc_fence_after();
      if (threadIdx.x == 256) BW_TRACE(i, 10);
      uint32_t r[5][16];  // d <= 80: one tcgen05.ld round
#pragma unroll
      for (int c = 0; c < 5; ++c)
        if (c < p.dk16) tmem_ld16(tm_dQ + lane_off + 16 * c, r[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);  // dQ is in registers: the next tile's dQ may be issued
      if (p.dq_direct) {  // warp-uniform: bf16 rows, one plain TMA store into the caller's dQ view
        if (lane == 0 && i > 0) tma_store_wait_read();
        __syncwarp();
        uint8_t* hrow = reinterpret_cast<uint8_t*>(sdQ) + (warp & 3) * (32 * p.d * 2) + lane * (p.d * 2);
#pragma unroll
        for (int c = 0; c < 10; ++c) {
          if (8 * c < p.d) {
#define BW_Q(k_) (__uint_as_float(r[c >> 1][(c & 1) * 8 + (k_)]) * p.scale)
            uint4 w;
            w.x = pack_bf16(BW_Q(0), BW_Q(1));
            w.y = pack_bf16(BW_Q(2), BW_Q(3));
            w.z = pack_bf16(BW_Q(4), BW_Q(5));
            w.w = pack_bf16(BW_Q(6), BW_Q(7));
#undef BW_Q
            *reinterpret_cast<uint4*>(hrow + 16 * c) = w;
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {  // rows >= Sq are clipped by the map
          tma_store_4d(&tmdQ16, reinterpret_cast<uint8_t*>(sdQ) + (warp & 3) * (32 * p.d * 2), 0, i * 128 + (warp & 3) * 32, h, b);
          tma_store_commit();
        }
      } else {
        if (lane == 0 && i > 0) tma_store_wait_read();  // the previous tile's reduce has read the staging rows
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 20; ++c) {
          if (4 * c < p.d)
            *reinterpret_cast<float4*>(srow + 4 * c) =
                make_float4(__uint_as_float(r[c >> 2][(c & 3) * 4]) * p.scale, __uint_as_float(r[c >> 2][(c & 3) * 4 + 1]) * p.scale,
                            __uint_as_float(r[c >> 2][(c & 3) * 4 + 2]) * p.scale, __uint_as_float(r[c >> 2][(c & 3) * 4 + 3]) * p.scale);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {  // rows >= Sq are clipped by the map
          tma_reduce_add_3d(&tmdQ, stage, h * p.d, i * 128 + (warp & 3) * 32, b);
          tma_store_commit();
        }
      }
      if (threadIdx.x == 256) BW_TRACE(i, 11);
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// fp32 [rows, cols] (ld_in) -> bf16 [rows, cols] (ld_out): dQ scratch -> its slot in the d(qkv) matrix
__global__ void cast2d_f32_bf16_kernel(const float* in, long ld_in, bf16* out, long ld_out, int rows, int cols) {
  pdl_sync();
  const int c4 = cols / 4;
  const long total = (long)rows * c4;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = int(i / c4), c = int(i % c4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(in + (long)r * ld_in + c);
    uint2 w;
    w.x = pack_bf16(v.x, v.y);
    w.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(out + (long)r * ld_out + c) = w;
  }
}

}  // namespace tdm

using namespace tdm;

extern "C" int tdm_attn_bwd(const TdmAttn* a, void* stream) {
  TDM_REQUIRE(a && a->Q && a->K && a->V && a->O && a->dO && a->lse && a->delta && a->dq_accum && a->dQ && a->dK && a->dV,
              "tdm_attn_bwd: null input");
  TDM_REQUIRE(a->d % 8 == 0 && a->d >= 16 && a->d <= 128, "tdm_attn_bwd: head_dim=%d must be a multiple of 8 in [16,128]", a->d);
  TDM_REQUIRE(a->dq_stride_h == a->d && a->dq_stride_b == (int64_t)a->Sq * a->dq_stride_s && a->dq_stride_s % 4 == 0,
              "tdm_attn_bwd: dQ must be a [B*Sq, >=H*d] token matrix view");
  cudaStream_t s = (cudaStream_t)stream;
  CUtensorMap tmQ, tmK, tmV, tmdO;
  int rc;
  if ((rc = make_head_map(&tmQ, a->Q, a->B, a->Sq, a->H, a->d, a->q_stride_b, a->q_stride_s, a->q_stride_h))) return rc;
  if ((rc = make_head_map(&tmK, a->K, a->B, a->Sk, a->H, a->d, a->k_stride_b, a->k_stride_s, a->k_stride_h))) return rc;
  if ((rc = make_head_map(&tmV, a->V, a->B, a->Sk, a->H, a->d, a->v_stride_b, a->v_stride_s, a->v_stride_h))) return rc;
  if ((rc = make_head_map(&tmdO, a->dO, a->B, a->Sq, a->H, a->d, a->o_stride_b, a->o_stride_s, a->o_stride_h))) return rc;
  const long rows = (long)a->B * a->Sq;
  const int HD = a->H * a->d;
  // one key tile per (batch, head) on the pipelined kernel: dQ is written directly (see AttnBwdParams::dq_direct)
  const bool narrow = ((a->d + 15) / 16) * 16 <= 80 && (((long)a->H * a->d) % 4 == 0) && (a->d % 4 == 0);
  const bool dq_direct = narrow && a->Sk <= 128;
  float* dq_zero = dq_direct ? nullptr : a->dq_accum;
  if (a->o_stride_h == a->d && (size_t)(8 * a->H * a->d / 8) * 4 <= 48 * 1024) {
    ::tdm::count_launch();
    TDM_CHECK_CUDA(launch_pdl(attn_delta_rows_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), (size_t)(8 * (a->H * a->d / 8) * 4), s,
                              (const bf16*)a->O, (const bf16*)a->dO, (long)a->o_stride_b, (long)a->o_stride_s, a->delta, dq_zero,
                              a->B, a->H, a->Sq, a->d));
  } else {
    const long total = rows * a->H;
    ::tdm::count_launch();
    TDM_CHECK_CUDA(launch_pdl(attn_delta_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)0, s, (const bf16*)a->O,
                              (const bf16*)a->dO, (long)a->o_stride_b, (long)a->o_stride_s, (long)a->o_stride_h, a->delta,
                              dq_zero, a->B, a->H, a->Sq, a->d));
  }
  AttnBwdParams p;
  p.B = a->B; p.H = a->H; p.Sq = a->Sq; p.Sk = a->Sk; p.d = a->d;
  p.nd = (a->d + 63) / 64; p.dk16 = (a->d + 15) / 16;
  p.scale = a->scale; p.scale_log2 = a->scale * 1.4426950408889634f;
  p.bias = a->bias; p.lse = a->lse; p.delta = a->delta; p.dq_accum = a->dq_accum;
  p.dK = (bf16*)a->dK; p.dk_stride_b = a->dk_stride_b; p.dk_stride_s = a->dk_stride_s; p.dk_stride_h = a->dk_stride_h;
  p.dV = (bf16*)a->dV; p.dv_stride_b = a->dv_stride_b; p.dv_stride_s = a->dv_stride_s; p.dv_stride_h = a->dv_stride_h;
  dim3 grid((a->Sk + 127) / 128, a->H, a->B);
  p.nc = (a->d + 31) / 32;
  p.dq_direct = dq_direct ? 1 : 0;
  if (narrow) {
    if ((rc = make_head_map64(&tmQ, a->Q, a->B, a->Sq, a->H, a->d, a->q_stride_b, a->q_stride_s, a->q_stride_h))) return rc;
    if ((rc = make_head_map64(&tmK, a->K, a->B, a->Sk, a->H, a->d, a->k_stride_b, a->k_stride_s, a->k_stride_h))) return rc;
    if ((rc = make_head_map64(&tmV, a->V, a->B, a->Sk, a->H, a->d, a->v_stride_b, a->v_stride_s, a->v_stride_h))) return rc;
    if ((rc = make_head_map64(&tmdO, a->dO, a->B, a->Sq, a->H, a->d, a->o_stride_b, a->o_stride_s, a->o_stride_h))) return rc;
    CUtensorMap tmdQ, tmdK, tmdV, tmdQ16;
    {  // fp32 accumulator [B, Sq, H*d]: one 32-row x d box per drain warp; bf16 dK / dV views [B, Sk, H, d]
      uint64_t dims[3] = {(uint64_t)HD, (uint64_t)a->Sq, (uint64_t)a->B};
      uint64_t strides[2] = {(uint64_t)HD * 4, (uint64_t)a->Sq * HD * 4};
      uint32_t box[3] = {(uint32_t)a->d, 32, 1};
      if ((rc = make_tma_typed(&tmdQ, a->dq_accum, 1, 3, dims, strides, box, TMA_SWZ_NONE))) return rc;
      uint64_t dims4[4] = {(uint64_t)a->d, (uint64_t)a->Sk, (uint64_t)a->H, (uint64_t)a->B};
      uint32_t box4[4] = {(uint32_t)a->d, 32, 1, 1};
      uint64_t sk[3] = {(uint64_t)a->dk_stride_s * 2, (uint64_t)a->dk_stride_h * 2, (uint64_t)a->dk_stride_b * 2};
      if ((rc = make_tma_bf16(&tmdK, a->dK, 4, dims4, sk, box4, TMA_SWZ_NONE))) return rc;
      uint64_t sv[3] = {(uint64_t)a->dv_stride_s * 2, (uint64_t)a->dv_stride_h * 2, (uint64_t)a->dv_stride_b * 2};
      if ((rc = make_tma_bf16(&tmdV, a->dV, 4, dims4, sv, box4, TMA_SWZ_NONE))) return rc;
      uint64_t dimsq[4] = {(uint64_t)a->d, (uint64_t)a->Sq, (uint64_t)a->H, (uint64_t)a->B};
      uint64_t sq[3] = {(uint64_t)a->dq_stride_s * 2, (uint64_t)a->dq_stride_h * 2, (uint64_t)a->dq_stride_b * 2};
      if ((rc = make_tma_bf16(&tmdQ16, a->dQ, 4, dimsq, sq, box4, TMA_SWZ_NONE))) return rc;
    }
    p.dq_stage_bytes = 128 * a->d * 4;
    const int smem = 6 * p.nc * CH64 + 2 * CHUNK + p.dq_stage_bytes + 2048 + 256;
    static int smem_set = 0;
    if (smem > smem_set) {
      TDM_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      TDM_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      smem_set = smem;
    }
    p.trace = nullptr;
    const char* trace_path = getenv("TDM_B200_ATTN_TRACE");  // debugging aid: per-tile clock64 stamps of CTA (0,0,0)
    static long long* trace_buf = nullptr;
    if (trace_path && *trace_path) {
      if (!trace_buf) TDM_CHECK_CUDA(cudaMalloc(&trace_buf, 32 * 16 * sizeof(long long)));
      TDM_CHECK_CUDA(cudaMemsetAsync(trace_buf, 0, 32 * 16 * sizeof(long long), s));
      p.trace = trace_buf;
    }
    ::tdm::count_launch();
    if (p.trace) TDM_CHECK_CUDA(launch_pdl(attn_bwd_kernel<true>, grid, dim3(BWD_THREADS), (size_t)smem, s, tmQ, tmK, tmV, tmdO, tmdQ, tmdK, tmdV, tmdQ16, p));
    else TDM_CHECK_CUDA(launch_pdl(attn_bwd_kernel<false>, grid, dim3(BWD_THREADS), (size_t)smem, s, tmQ, tmK, tmV, tmdO, tmdQ, tmdK, tmdV, tmdQ16, p));
    if (p.trace) {
      static long long host[32 * 16];
      TDM_CHECK_CUDA(cudaStreamSynchronize(s));
      TDM_CHECK_CUDA(cudaMemcpy(host, p.trace, sizeof(host), cudaMemcpyDeviceToHost));
      if (FILE* f = fopen(trace_path, "a")) {
        fprintf(f, "# bwd Sq=%d Sk=%d d=%d events: MMA 0 top 1 p_full seen 2 S^T(i+1) issued 3 dq_free seen 4 dQ issued | compute 5 top 6 st_full seen 7 P^T arrive 8 first dS chunk ready 9 ds arrive | drain 10 dq_full seen 11 reduce issued\n",
                a->Sq, a->Sk, a->d);
        const long long t0 = host[0];
        for (int g = 0; g < 32; ++g) {
          if (!host[g * 16 + 1]) break;
          fprintf(f, "%d", g);
          for (int e = 0; e < 12; ++e) fprintf(f, " %lld", host[g * 16 + e] ? host[g * 16 + e] - t0 : -1);
          fprintf(f, "\n");
        }
        fclose(f);
      }
    }
  } else {
    const int smem = 4 * p.nd * CHUNK + 4 * CHUNK + 1024 + 256 + 1024;
    static int smem_set_w = 0;
    if (smem > smem_set_w) {
      TDM_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      smem_set_w = smem;
    }
    ::tdm::count_launch(); attn_bwd_wide_kernel<<<grid, 192, smem, s>>>(tmQ, tmK, tmV, tmdO, p);
  }
  TDM_CHECK_CUDA(cudaGetLastError());
  if (dq_direct) return TDM_OK;
  // dQ scratch [B*Sq, H*d] fp32 -> bf16 dQ with the caller's strides (requires the [tokens, H*d] layout)
  const int nblk = num_sms() * 8;
  ::tdm::count_launch();
  TDM_CHECK_CUDA(launch_pdl(cast2d_f32_bf16_kernel, dim3(nblk), dim3(256), (size_t)0, s, (const float*)a->dq_accum, (long)HD, (bf16*)a->dQ, (long)a->dq_stride_s, (int)rows, (int)HD));
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}
