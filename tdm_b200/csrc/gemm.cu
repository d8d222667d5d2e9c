// tcgen05 / TMEM / TMA GEMM for sm_100a with a fused epilogue (see include/tdm_b200.h: TdmGemm).
//
// One persistent CTA per SM (gemm_bf16_kernel) or one CTA pair per SM pair (gemm2_bf16_kernel, cta_group::2: the
// default for every large problem), 384 threads, warp-specialised:
//   warp 0    : TMA producer  (one lane) -- fills a ring of {A 128x64, B BNx64 (or its half)} bf16 stages (SW128);
//               the ring depth is chosen per launch from the shared memory the epilogue leaves free
//   warp 1    : MMA issuer    (warp-convergent, tcgen05 ops under elect.sync) -- tcgen05.mma into one of two TMEM
//               accumulators
//   warp 2    : TMEM allocator
//   warps 4-11: epilogue      -- tcgen05.ld (thread == accumulator row) -> fused math -> TMA stores
// Three pipelines: smem full/empty (TMA<->MMA), TMEM full/empty (MMA<->epilogue), static tile loop.
// Two epilogues: the streaming one (FAST kernels: every epilogue tensor moves by TMA, feature-specialised instances)
// and the generic one (split-K reductions and the rare features); both produce bit-identical results.
// Tails in M, N and K are handled by TMA out-of-bounds zero fill / clipping.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"

namespace tdm {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;

struct GemmParams {
  int M, N, K, K2;
  int a_mn, b_mn;
  int seg_n, seg_k2;
  int seg_k, seg_m, seg_nb;  // stacked problems in one launch (single-CTA kernel): see TdmGemm
  int num_m_tiles, num_n_tiles;
  // CTA-pair kernel, BN = 256, last column tile at most BN / 2 wide (N = 1152, 3456: most GEMMs of a DiT block):
  //   narrow_last: that tile is computed with N = BN / 2 MMAs (half the tensor time and half the B traffic; the first form
  //                issued full-width MMAs on zero-filled B rows, so a 128-column tail cost as much as a 256-column tile);
  //   bal:         the narrow tiles are dealt to the pairs that would otherwise carry one tile more than the others.  Units
  //                go round-robin to the pairs; positions [bal_a0, bal_a1) u [bal_b0, bal_b1) u [bal_c0, bal_c1) of the unit
  //                sequence hold the narrow tiles: N = 1152 on 74 pairs = 128 full + 32 half tiles -> every pair <= 2.0
  //                tile times instead of 3 (gemm_unit_tile, gemm_plan_balance).
  int narrow_last, bal, bal_a0, bal_a1, bal_b0, bal_b1, bal_c0, bal_c1;
  int splits;        // split-K factor (>1 only in accumulate-only mode: partial sums are red.add'ed into out_f32)
  int stages;        // depth of the TMA<->MMA shared-memory ring (chosen on the host: what the epilogue leaves free)
  int epi_bytes;     // bytes of epilogue staging that follow the ring
  // streaming epilogue (FAST kernels): per warp a ring of `epi_nbuf` chunk sets of `epi_set_bytes`; a set holds the
  // fp32 tile (residual in -> out_f32, at 0), the first bf16 tile (act_in in -> out_bf16, at epi_off_h0) and the
  // second bf16 tile (pre_out, at epi_off_h1); `epi_in_bytes` = bytes the TMA loads per chunk (0: nothing to load)
  int epi_nbuf, epi_set_bytes, epi_off_h0, epi_off_h1, epi_in_bytes;
  int epi_feat_slot;  // which specialised instance of the streaming epilogue runs (-1: the run-time tested one)
  int gate_uniform;  // rows_per_gate % BM == 0: one gate row per tile
  float alpha;
  int act;
  const float* bias;
  const bf16* act_in;
  long ld_act;
  bf16* pre_out;
  long ld_pre;
  float* pre_out_f32;
  long ld_pre_f32;
  int res_row_mod;
  const float* gate;
  int rows_per_gate;
  long gate_stride;
  const float* res;
  long ld_res;
  float* out_f32;
  long ld_out_f32;
  bf16* out_bf16;
  long ld_out_bf16;
};

constexpr int EPI_WARPS = 8;   // two warps per TMEM lane quarter, each takes every other 32-column chunk of the tile
constexpr int GEMM_THREADS = 128 + 32 * EPI_WARPS;

// Epilogue shared memory: per epilogue warp one 32-row x 32-column staging tile for each output precision (fp32 with
// the 128-byte swizzle, bf16 with the 64-byte swizzle) that a TMA store drains -- every HBM write is a full line and
// the LSU only sees conflict-free 16-byte shared-memory stores -- plus the tile's bias / gate rows, staged once per
// tile by all epilogue warps BEFORE they wait for the accumulator (their latency hides behind the mainloop).
constexpr int EPI_STAGE_F32 = 32 * 32 * 4;                        // 4 KB
constexpr int EPI_STAGE_BF16 = 32 * 32 * 2;                       // 2 KB
constexpr int EPI_STAGE_WARP = EPI_STAGE_F32 + EPI_STAGE_BF16;    // 6 KB per warp
constexpr int EPI_VEC_FLOATS = 256;                               // >= widest tile
constexpr int EPI_GENERIC_BYTES = EPI_WARPS * EPI_STAGE_WARP;   // 48 KB: staging of the generic epilogue
constexpr int EPI_VEC_BYTES = 8192;  // generic: double-buffered bias / gate rows (4 KB); FAST: 1 KB of private rows per warp
constexpr int MAX_STAGES = 8;
constexpr int EPI_MAX_NBUF = 3;
constexpr int BAR_BYTES = 512;  // full[8] empty[8] tfull[2] tempty[2] ld[8 warps x 3] tmem slot
constexpr int SMEM_LIMIT = 232448;  // 227 KB dynamic shared memory per CTA on sm_100
// shared memory: [ring: stages x STAGE_BYTES][epilogue staging: epi_bytes][bias/gate rows][barriers]
__host__ __device__ constexpr int gemm_smem_bytes(int stages, int stage_bytes, int epi_bytes) {
  return stages * stage_bytes + epi_bytes + EPI_VEC_BYTES + BAR_BYTES + 1024 /*align*/;
}

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ACC_STRIDE = (BN <= 64) ? 64 : (BN <= 128 ? 128 : 256);  // TMEM columns between the two accumulators
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
};

__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.f + tanh_fast(k0 * (x + k1 * x * x * x)));
}
__device__ __forceinline__ float gelu_tanh_grad_fast(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float x2 = x * x;
  const float t = tanh_fast(k0 * (x + k1 * x * x2));
  return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * k0 * (1.f + 3.f * k1 * x2);
}

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory"); }

// All epilogue warps: stage bias[n0 .. n0+BN) and (per-tile uniform) gate[n0 .. n0+BN) into buffer `as`.
template <int BN>
__device__ __forceinline__ void gemm_epilogue_stage_vectors(const GemmParams& p, float* svec, int as, int m0, int n0, int et) {
  float* sb = svec + as * 2 * EPI_VEC_FLOATS;
  float* sg = sb + EPI_VEC_FLOATS;
  const float* gate_tile = (p.gate && p.gate_uniform) ? p.gate + (long)(m0 / p.rows_per_gate) * p.gate_stride : nullptr;
  for (int n = et; n < BN; n += 32 * EPI_WARPS) {
    const bool ok = n0 + n < p.N;
    sb[n] = (p.bias && ok) ? p.bias[n0 + n] : 0.f;
    sg[n] = (gate_tile && ok) ? gate_tile[n0 + n] : 1.f;
  }
  epi_barrier();
}

// Unit (position in the round-robin sequence) -> tile coordinates.  Identity order (n fastest) unless p.bal.
__device__ __forceinline__ void gemm_unit_tile(const GemmParams& p, int tile, int& mt, int& nt) {
  const int nn = p.num_n_tiles;
  if (!p.bal) {
    mt = tile / nn;
    nt = tile - mt * nn;
    return;
  }
  auto below = [](int u, int lo, int hi) { return u <= lo ? 0 : (u < hi ? u - lo : hi - lo); };  // positions of [lo, hi) before u
  const int narrow_before = below(tile, p.bal_a0, p.bal_a1) + below(tile, p.bal_b0, p.bal_b1) + below(tile, p.bal_c0, p.bal_c1);
  const bool is_narrow = (tile >= p.bal_a0 && tile < p.bal_a1) || (tile >= p.bal_b0 && tile < p.bal_b1) ||
                         (tile >= p.bal_c0 && tile < p.bal_c1);
  if (is_narrow) {
    mt = narrow_before;
    nt = nn - 1;
  } else {
    const int f = tile - narrow_before;
    mt = f / (nn - 1);
    nt = f - mt * (nn - 1);
  }
}

// Fused epilogue of one 128-row accumulator quarter: thread == accumulator row (tcgen05.ld 32x32b), so the 32 values
// of a chunk sit in registers (32-way ILP).  `q` = TMEM lane quarter of the calling warp, `ch` = even / odd chunks.
template <int BN>
__device__ __forceinline__ void gemm_epilogue_rows(const GemmParams& p, const CUtensorMap* tmC16, const CUtensorMap* tmC32,
                                                   uint32_t taddr, int m0, int n0, int q, int ch, int lane, uint8_t* stg,
                                                   const float* svec, int as) {
  const int m = m0 + q * 32 + lane;
  const bool row_ok = m < p.M;
  const float* sb = svec + as * 2 * EPI_VEC_FLOATS;
  const float* sg = sb + EPI_VEC_FLOATS;
  const float* gate_row = nullptr;  // only when the gate changes inside a tile
  if (p.gate && !p.gate_uniform && row_ok) gate_row = p.gate + (long)(m / p.rows_per_gate) * p.gate_stride;
  const long rr = p.res_row_mod > 0 ? (m % p.res_row_mod) : m;
  uint8_t* stg_f = stg + lane * 128;                      // fp32 row (128 B, SW128: unit ^= row & 7)
  uint8_t* stg_h = stg + EPI_STAGE_F32 + lane * 64;       // bf16 row (64 B, SW64: unit ^= (row >> 1) & 3)
  const int xf = lane & 7, xh = (lane >> 1) & 3;
#pragma unroll 1
  for (int c = ch * 32; c < BN; c += 64) {
    if (n0 + c >= p.N) break;  // warp-uniform
    uint32_t acc[32];
    tmem_ld32(taddr + c, acc);
    tmem_ld_wait();
    if (p.splits > 1) {  // accumulate-only split-K: partial sums meet in HBM
      // The 32 x 32 fp32 chunk is staged like an fp32 output and ADDED by the TMA engine (cp.reduce.async.bulk.tensor):
      // one bulk operation per chunk.  The first form issued red.global.add.v4.f32 per thread -- 32 rows = 32 memory
      // transactions per instruction, 256 per chunk -- which clogs the SM's load/store queue that the mbarrier waits and
      // tcgen05.ld of every warp share (found in the attention kernels, attention_fwd2.cu).  Rows >= M and columns >= N
      // are clipped by the tensor map.
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
#pragma unroll
      for (int g = 0; g < 8; ++g)
        *reinterpret_cast<float4*>(stg_f + ((g ^ xf) << 4)) =
            make_float4(p.alpha * __uint_as_float(acc[4 * g]), p.alpha * __uint_as_float(acc[4 * g + 1]),
                        p.alpha * __uint_as_float(acc[4 * g + 2]), p.alpha * __uint_as_float(acc[4 * g + 3]));
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_2d(tmC32, stg, n0 + c, m0 + q * 32);
        tma_store_commit();
      }
      continue;
    }
    // the TMA engine may still be reading this warp's staging tiles for the previous chunk
    if (lane == 0) tma_store_wait_read();
    __syncwarp();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int n = n0 + c + g * 8;
      const bool ok = row_ok && n < p.N;  // guards every direct global access; staged stores are clipped by the TMA
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(p.alpha, __uint_as_float(acc[g * 8 + i]), sb[c + g * 8 + i]);  // as the streaming epilogue
      if (p.pre_out && ok) {
        uint4 o;
        o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
        o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
        *reinterpret_cast<uint4*>(p.pre_out + (long)m * p.ld_pre + n) = o;
      }
      if (p.pre_out_f32 && ok) {
        float* o = p.pre_out_f32 + (long)m * p.ld_pre_f32 + n;
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
      if (p.act == TDM_ACT_GELU_TANH) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = gelu_tanh_fast(v[i]);
      } else if (p.act == TDM_ACT_GELU_TANH_BWD) {
        uint4 u = make_uint4(0, 0, 0, 0);
        if (ok) u = *reinterpret_cast<const uint4*>(p.act_in + (long)m * p.ld_act + n);
        const float2 u0 = unpack_bf16(u.x), u1 = unpack_bf16(u.y), u2 = unpack_bf16(u.z), u3 = unpack_bf16(u.w);
        v[0] *= gelu_tanh_grad_fast(u0.x); v[1] *= gelu_tanh_grad_fast(u0.y);
        v[2] *= gelu_tanh_grad_fast(u1.x); v[3] *= gelu_tanh_grad_fast(u1.y);
        v[4] *= gelu_tanh_grad_fast(u2.x); v[5] *= gelu_tanh_grad_fast(u2.y);
        v[6] *= gelu_tanh_grad_fast(u3.x); v[7] *= gelu_tanh_grad_fast(u3.y);
      } else if (p.act == TDM_ACT_SILU) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = v[i] / (1.f + __expf(-v[i]));
      } else if (p.act == TDM_ACT_SILU_BWD) {
        uint4 u = make_uint4(0, 0, 0, 0);
        if (ok) u = *reinterpret_cast<const uint4*>(p.act_in + (long)m * p.ld_act + n);
        const float2 u0 = unpack_bf16(u.x), u1 = unpack_bf16(u.y), u2 = unpack_bf16(u.z), u3 = unpack_bf16(u.w);
        const float uu[8] = {u0.x, u0.y, u1.x, u1.y, u2.x, u2.y, u3.x, u3.y};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float sgm = 1.f / (1.f + __expf(-uu[i]));
          v[i] *= sgm * (1.f + uu[i] * (1.f - sgm));
        }
      }
      if (p.gate) {
        float4 g0, g1;
        if (gate_row) {
          g0 = ok ? *reinterpret_cast<const float4*>(gate_row + n) : make_float4(1, 1, 1, 1);
          g1 = ok ? *reinterpret_cast<const float4*>(gate_row + n + 4) : make_float4(1, 1, 1, 1);
        } else {
          g0 = *reinterpret_cast<const float4*>(sg + c + g * 8);
          g1 = *reinterpret_cast<const float4*>(sg + c + g * 8 + 4);
        }
        v[0] = __fmul_rn(v[0], g0.x); v[1] = __fmul_rn(v[1], g0.y); v[2] = __fmul_rn(v[2], g0.z); v[3] = __fmul_rn(v[3], g0.w);
        v[4] = __fmul_rn(v[4], g1.x); v[5] = __fmul_rn(v[5], g1.y); v[6] = __fmul_rn(v[6], g1.z); v[7] = __fmul_rn(v[7], g1.w);
      }
      if (p.res && ok) {
        const float4 r0 = *reinterpret_cast<const float4*>(p.res + rr * p.ld_res + n);
        const float4 r1 = *reinterpret_cast<const float4*>(p.res + rr * p.ld_res + n + 4);
        v[0] = __fadd_rn(v[0], r0.x); v[1] = __fadd_rn(v[1], r0.y); v[2] = __fadd_rn(v[2], r0.z); v[3] = __fadd_rn(v[3], r0.w);
        v[4] = __fadd_rn(v[4], r1.x); v[5] = __fadd_rn(v[5], r1.y); v[6] = __fadd_rn(v[6], r1.z); v[7] = __fadd_rn(v[7], r1.w);
      }
      if (p.out_f32) {
        *reinterpret_cast<float4*>(stg_f + (((2 * g) ^ xf) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(stg_f + (((2 * g + 1) ^ xf) << 4)) = make_float4(v[4], v[5], v[6], v[7]);
      }
      if (p.out_bf16) {
        uint4 o;
        o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
        o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
        *reinterpret_cast<uint4*>(stg_h + ((g ^ xh) << 4)) = o;
      }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (p.out_f32) tma_store_2d(tmC32, stg, n0 + c, m0 + q * 32);
      if (p.out_bf16) tma_store_2d(tmC16, stg + EPI_STAGE_F32, n0 + c, m0 + q * 32);
      tma_store_commit();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Streaming epilogue (FAST kernels).  Everything that touches global memory is a TMA transfer of a 32 x 32 chunk:
// the residual and the backward-activation input are LOADED into the chunk set one chunk ahead (the first chunk of a
// tile before the accumulator is even waited for), every output is written into the set in place and drained by a
// TMA store; sets form a ring of 2-3 per warp so a store is never waited for right after it was issued.  The next
// chunk's tcgen05.ld is in flight while the current one is processed, and the accumulator is handed back to the MMA
// warp as soon as this warp's last chunk sits in registers.  All option tests are warp-uniform and made once per
// chunk, shared memory is addressed through 32-bit shared-space instructions: ~110 SASS instructions per chunk for a
// bias + bf16 epilogue (the generic epilogue below: ~620, and it was the bottleneck of every K <= 1152 GEMM).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tma_store_wait_read_n(int n) {  // at most n of this thread's store groups still reading
  if (n <= 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  else if (n == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
}

#ifdef TDM_DBG_TIMING
__device__ long long tdm_dbg_times[8192];
__device__ uint32_t tdm_dbg_cj[32 * 8];
#define DBG_T(tag)                                                                  \
  do {                                                                              \
    if (blockIdx.x == 3 && lane == 0 && q == 1 && ch == 0 && cj < 500)              \
      tdm_dbg_times[cj * 8 + (tag)] = clock64();                                    \
  } while (0)
#else
#define DBG_T(tag)
#endif

__device__ __forceinline__ float ld_global_f32(const float* ptr) {
  float v;
  asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(ptr) : "memory");
  return v;
}

struct EpiMaps {
  const CUtensorMap *c16, *c32, *res, *act, *pre;
};

// Feature sets the streaming epilogue is compiled for.  The hot GEMMs of a DiT block hit one of the exact sets and run
// branch-free chunk code; any other combination takes FEAT_DYN, the same code with the tests made at run time.
enum : int {
  EF_OUT16 = 1, EF_OUT32 = 2, EF_RES = 4, EF_GATE = 8, EF_PRE = 16, EF_GELU = 32, EF_GELU_BWD = 64, FEAT_DYN = 1 << 20
};
__host__ __device__ constexpr int epi_feat_slot(int feat) {  // index of the specialised instance, -1: dynamic
  switch (feat) {
    case EF_OUT16: return 0;                                  // QKV, q2, kv2, plain dgrads, LoRA t / dt
    case EF_OUT16 | EF_GELU: return 1;                        // FF1
    case EF_OUT16 | EF_GELU | EF_PRE: return 2;               // FF1, pre-activation kept for the backward
    case EF_OUT16 | EF_GELU_BWD: return 3;                    // FF2 dgrad
    case EF_OUT32 | EF_RES: return 4;                         // o2, q2 dgrad, un-split wgrad accumulation
    case EF_OUT32 | EF_RES | EF_GATE: return 5;               // FF2
    case EF_OUT32 | EF_RES | EF_GATE | EF_OUT16: return 6;    // O-proj
    case EF_OUT16 | EF_RES: return 7;                         // LoRA merge: bf16(W32 + s B A)
    default: return -1;
  }
}

// One tile, one epilogue warp.  `ring` / `ldbar` are this warp's chunk sets and load barriers, `cj` its running
// chunk count (ring position and barrier phase), `vec` its private bias / gate rows (chunk k of the tile at
// vec + k * 256: 32 bias floats, 32 gate floats).  Returns after the last store of the tile has been ISSUED.
template <int BN, bool PAIR, int FEAT>
__device__ __forceinline__ void gemm_epilogue_fast(const GemmParams& p, const EpiMaps& tm, uint32_t taddr, int m0, int n0, int q,
                                                   int ch, int lane, uint32_t ring, uint64_t* ldbar, uint32_t& buf, uint32_t& ph,
                                                   uint32_t vec, uint64_t* tfull, uint32_t aphase, uint64_t* tempty) {
  constexpr bool DYN = FEAT == FEAT_DYN;
  const bool f_out16 = DYN ? p.out_bf16 != nullptr : bool(FEAT & EF_OUT16);
  const bool f_out32 = DYN ? p.out_f32 != nullptr : bool(FEAT & EF_OUT32);
  const bool f_res = DYN ? p.res != nullptr : bool(FEAT & EF_RES);
  const bool f_gate = DYN ? p.gate != nullptr : bool(FEAT & EF_GATE);
  const bool f_pre = DYN ? p.pre_out != nullptr : bool(FEAT & EF_PRE);
  const bool f_gelu = DYN ? p.act == TDM_ACT_GELU_TANH : bool(FEAT & EF_GELU);
  const bool f_gelu_bwd = DYN ? p.act == TDM_ACT_GELU_TANH_BWD : bool(FEAT & EF_GELU_BWD);
  const bool f_silu = DYN ? p.act == TDM_ACT_SILU : false;
  const bool f_silu_bwd = DYN ? p.act == TDM_ACT_SILU_BWD : false;
  const bool f_actin = f_gelu_bwd || f_silu_bwd;
  const bool has_in = f_res || f_actin;
  const int row0 = m0 + q * 32;
  const int nlim = (p.N - n0) < BN ? (p.N - n0) : BN;  // columns of this tile that exist
  const uint32_t nbuf = (uint32_t)p.epi_nbuf;
#ifdef TDM_DBG_TIMING
  uint32_t& cj = tdm_dbg_cj[(blockIdx.x % 32) * EPI_WARPS + q + 4 * ch];
#endif
  const uint32_t xf = (lane & 7) << 4, xh = ((lane >> 1) & 3) << 4;  // SW128 / SW64 XOR of this thread's row
  int c = ch * 32;
  auto hand_back = [&]() {  // the accumulator is in registers (or not needed): the MMA warp may overwrite it
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if constexpr (PAIR) mbar_arrive_cluster(smem_u32(tempty) & PEER_BIT_MASK);  // the leader's barrier
      else mbar_arrive(tempty);
    }
  };
  auto issue_loads = [&](uint32_t buf, int col) {  // lane 0: residual / activation-input chunk -> set `buf`
    const uint32_t set = ring + buf * p.epi_set_bytes;
    uint64_t* bar = &ldbar[buf];
    mbar_arrive_expect_tx(bar, (uint32_t)p.epi_in_bytes);
    const uint32_t bar32 = smem_u32(bar);
    if (f_res)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(set), "l"(reinterpret_cast<uint64_t>(tm.res)), "r"(bar32), "r"(n0 + col), "r"(row0) : "memory");
    if (f_actin)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(set + p.epi_off_h0), "l"(reinterpret_cast<uint64_t>(tm.act)), "r"(bar32), "r"(n0 + col), "r"(row0) : "memory");
  };
  if (c >= nlim) {  // this warp has no chunk in a narrow tail tile
    mbar_wait_relaxed(tfull, aphase);
    hand_back();
    return;
  }
  DBG_T(1);
  if (has_in && lane == 0) {  // the first chunk's inputs travel while the mainloop finishes the accumulator
    tma_store_wait_read_n(nbuf - 1);
    issue_loads(buf, c);
  }
  mbar_wait_relaxed(tfull, aphase);
  tc_fence_after();
  DBG_T(2);
  uint32_t acc[32];
  tmem_ld32(taddr + c, acc);
  uint32_t sb = vec;  // bias row of this chunk (zeros without a bias), the gate row follows (ones without a gate)
#pragma unroll 1
  for (; c < nlim; c += 64, sb += 256) {
    const uint32_t set = ring + buf * p.epi_set_bytes;
    const uint32_t buf_next = buf + 1 == nbuf ? 0 : buf + 1;
    const uint32_t rowF = set + lane * 128, rowH0 = set + p.epi_off_h0 + lane * 64, rowH1 = set + p.epi_off_h1 + lane * 64;
    tmem_ld_wait();
    DBG_T(3);
    float v[32];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const uint4 b = lds128(sb + g * 16);
      v[g * 4 + 0] = fmaf(p.alpha, __uint_as_float(acc[g * 4 + 0]), __uint_as_float(b.x));
      v[g * 4 + 1] = fmaf(p.alpha, __uint_as_float(acc[g * 4 + 1]), __uint_as_float(b.y));
      v[g * 4 + 2] = fmaf(p.alpha, __uint_as_float(acc[g * 4 + 2]), __uint_as_float(b.z));
      v[g * 4 + 3] = fmaf(p.alpha, __uint_as_float(acc[g * 4 + 3]), __uint_as_float(b.w));
    }
    const int cn = c + 64;
    const bool more = cn < nlim;  // warp-uniform
    if (more) tmem_ld32(taddr + cn, acc);  // in flight while this chunk is processed
    else hand_back();
    if (lane == 0) {
      if (has_in) {
        if (more) {
          tma_store_wait_read_n(nbuf - 2);  // the set of chunk cj + 1 was last stored from by chunk cj + 1 - nbuf
          issue_loads(buf_next, cn);
        }
      } else {
        tma_store_wait_read_n(nbuf - 1);  // this set was last stored from by chunk cj - nbuf
      }
    }
    __syncwarp();
    DBG_T(4);
    if (has_in) mbar_wait(&ldbar[buf], ph);
    if (f_pre) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sts128(rowH1 + ((j << 4) ^ xh), pack_bf16(v[j * 8 + 0], v[j * 8 + 1]), pack_bf16(v[j * 8 + 2], v[j * 8 + 3]),
               pack_bf16(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16(v[j * 8 + 6], v[j * 8 + 7]));
    }
    if (f_gelu) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu_tanh_fast(v[i]);
    } else if (f_actin) {
      float u[32];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 w = lds128(rowH0 + ((j << 4) ^ xh));
        const float2 u0 = unpack_bf16(w.x), u1 = unpack_bf16(w.y), u2 = unpack_bf16(w.z), u3 = unpack_bf16(w.w);
        u[j * 8 + 0] = u0.x; u[j * 8 + 1] = u0.y; u[j * 8 + 2] = u1.x; u[j * 8 + 3] = u1.y;
        u[j * 8 + 4] = u2.x; u[j * 8 + 5] = u2.y; u[j * 8 + 6] = u3.x; u[j * 8 + 7] = u3.y;
      }
      if (f_gelu_bwd) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= gelu_tanh_grad_fast(u[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float sgm = 1.f / (1.f + __expf(-u[i]));
          v[i] *= sgm * (1.f + u[i] * (1.f - sgm));
        }
      }
    } else if (f_silu) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = v[i] / (1.f + __expf(-v[i]));
    }
    if (f_gate) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint4 w = lds128(sb + 128 + g * 16);
        v[g * 4 + 0] = __fmul_rn(v[g * 4 + 0], __uint_as_float(w.x)); v[g * 4 + 1] = __fmul_rn(v[g * 4 + 1], __uint_as_float(w.y));
        v[g * 4 + 2] = __fmul_rn(v[g * 4 + 2], __uint_as_float(w.z)); v[g * 4 + 3] = __fmul_rn(v[g * 4 + 3], __uint_as_float(w.w));
      }
    }
    if (f_res) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint4 w = lds128(rowF + ((g << 4) ^ xf));
        v[g * 4 + 0] = __fadd_rn(v[g * 4 + 0], __uint_as_float(w.x)); v[g * 4 + 1] = __fadd_rn(v[g * 4 + 1], __uint_as_float(w.y));
        v[g * 4 + 2] = __fadd_rn(v[g * 4 + 2], __uint_as_float(w.z)); v[g * 4 + 3] = __fadd_rn(v[g * 4 + 3], __uint_as_float(w.w));
      }
    }
    if (f_out32) {
#pragma unroll
      for (int g = 0; g < 8; ++g)
        sts128(rowF + ((g << 4) ^ xf), __float_as_uint(v[g * 4 + 0]), __float_as_uint(v[g * 4 + 1]), __float_as_uint(v[g * 4 + 2]),
               __float_as_uint(v[g * 4 + 3]));
    }
    if (f_out16) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sts128(rowH0 + ((j << 4) ^ xh), pack_bf16(v[j * 8 + 0], v[j * 8 + 1]), pack_bf16(v[j * 8 + 2], v[j * 8 + 3]),
               pack_bf16(v[j * 8 + 4], v[j * 8 + 5]), pack_bf16(v[j * 8 + 6], v[j * 8 + 7]));
    }
    DBG_T(5);
    fence_async_smem();
    __syncwarp();
    DBG_T(6);
#ifndef TDM_DBG_NOSTORE
    if (lane == 0) {
      const int col = n0 + c;
      if (f_out32)
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(tm.c32)), "r"(set), "r"(col), "r"(row0) : "memory");
      if (f_out16)
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(tm.c16)), "r"(set + p.epi_off_h0), "r"(col), "r"(row0) : "memory");
      if (f_pre)
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(tm.pre)), "r"(set + p.epi_off_h1), "r"(col), "r"(row0) : "memory");
      tma_store_commit();
    }
#endif
    DBG_T(7);
    buf = buf_next;
    if (buf == 0) ph ^= 1;
#ifdef TDM_DBG_TIMING
    ++cj;
#endif
  }
}

// The epilogue warps' tile loop of a FAST kernel (shared by the single-CTA and the CTA-pair kernel).  Bias / gate
// rows are private to each warp: the values of the NEXT tile are fetched into registers while the current tile is
// processed and parked in the warp's own shared-memory rows at the start of that tile -- no CTA-wide barrier.
template <int BN, bool PAIR>
__device__ __forceinline__ void gemm_epilogue_warp_loop(const GemmParams& p, const EpiMaps& maps, uint32_t tmem_base, int acc_stride,
                                                        int ew, int lane, uint32_t ring, uint64_t* ldbar, uint32_t vec,
                                                        uint64_t* tfull_bar, uint64_t* tempty_bar, int unit0, int unit_step,
                                                        int num_units, int rank) {
  constexpr int NCH = (BN + 63) / 64;  // chunks per warp and tile
  const int q = ew & 3;    // == warp % 4: the TMEM lane quarter this warp may read
  const int ch = ew >> 2;  // even / odd 32-column chunks of the tile
  const int feat_slot = p.epi_feat_slot;
  float bias_r[NCH], gate_r[NCH];
  // FAST kernels never split K: unit == tile.  One gemm_unit_tile per tile (two integer divisions every ~8 us).
  int mt, nt;
  gemm_unit_tile(p, unit0 < num_units ? unit0 : 0, mt, nt);
  auto fetch_vec = [&](int mt_, int nt_) {
    const int m0_ = mt_ * (PAIR ? 2 * BM : BM) + rank * BM, n0_ = nt_ * BN;
    const float* gate_tile = nullptr;
    if (p.gate) gate_tile = p.gate + (long)(m0_ / p.rows_per_gate) * p.gate_stride;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const int n = n0_ + ch * 32 + k * 64 + lane;
      const bool ok = n < p.N && ch * 32 + k * 64 < BN;
      // plain (coherent, ordered) loads: the gate rows are written by the kernel just before this one, and an
      // invariant (__ldg) load may be hoisted above griddepcontrol.wait by the compiler
      bias_r[k] = (p.bias && ok) ? ld_global_f32(p.bias + n) : 0.f;
      gate_r[k] = (gate_tile && ok) ? ld_global_f32(gate_tile + n) : 1.f;
    }
  };
  uint32_t buf = 0, ph = 0;
  int it = 0;
  if (unit0 < num_units) fetch_vec(mt, nt);
  for (int unit = unit0; unit < num_units; unit += unit_step, ++it) {
    const int as = it & 1;
    const uint32_t aphase = (it >> 1) & 1;
    const int m0 = mt * (PAIR ? 2 * BM : BM) + rank * BM, n0 = nt * BN;
    __syncwarp();  // the previous tile's reads of the rows are done
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(vec + k * 256 + lane * 4), "f"(bias_r[k]) : "memory");
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(vec + k * 256 + 128 + lane * 4), "f"(gate_r[k]) : "memory");
    }
    __syncwarp();
    if (unit + unit_step < num_units) {
      gemm_unit_tile(p, unit + unit_step, mt, nt);
      fetch_vec(mt, nt);  // the next tile's rows travel during this tile
    }
    const uint32_t taddr = tmem_base + as * acc_stride + (uint32_t(q * 32) << 16);
#ifndef TDM_DBG_NOEPI
#define TDM_EPI_CASE(F) \
  gemm_epilogue_fast<BN, PAIR, F>(p, maps, taddr, m0, n0, q, ch, lane, ring, ldbar, buf, ph, vec, &tfull_bar[as], aphase, &tempty_bar[as])
    switch (feat_slot) {
      case 0: TDM_EPI_CASE(EF_OUT16); break;
      case 1: TDM_EPI_CASE(EF_OUT16 | EF_GELU); break;
      case 2: TDM_EPI_CASE(EF_OUT16 | EF_GELU | EF_PRE); break;
      case 3: TDM_EPI_CASE(EF_OUT16 | EF_GELU_BWD); break;
      case 4: TDM_EPI_CASE(EF_OUT32 | EF_RES); break;
      case 5: TDM_EPI_CASE(EF_OUT32 | EF_RES | EF_GATE); break;
      case 6: TDM_EPI_CASE(EF_OUT32 | EF_RES | EF_GATE | EF_OUT16); break;
      case 7: TDM_EPI_CASE(EF_OUT16 | EF_RES); break;
      default: TDM_EPI_CASE(FEAT_DYN); break;
    }
#undef TDM_EPI_CASE
#else
    mbar_wait(&tfull_bar[as], aphase);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if constexpr (PAIR) mbar_arrive_cluster(smem_u32(&tempty_bar[as]) & PEER_BIT_MASK);
      else mbar_arrive(&tempty_bar[as]);
    }
#endif
  }
}

template <int BN, bool FAST>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                 const __grid_constant__ CUtensorMap tmC16, const __grid_constant__ CUtensorMap tmC32,
                 const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmAct,
                 const __grid_constant__ CUtensorMap tmPre, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  const int STAGES = p.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_stage = smem + STAGES * Cfg::STAGE_BYTES;  // 1024-byte aligned: stages are multiples of 8 KB
  float* epi_vec = reinterpret_cast<float*>(epi_stage + p.epi_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + p.epi_bytes + EPI_VEC_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* ld_bar = tempty_bar + 2;  // [EPI_WARPS][EPI_MAX_NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ld_bar + EPI_WARPS * EPI_MAX_NBUF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nk1 = (p.K + BK - 1) / BK;
  const int nk2 = (p.K2 + BK - 1) / BK;
  const int nk = nk1 + nk2;
  const int num_units = p.num_m_tiles * p.num_n_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (nk2) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
  }
  if (FAST && warp == 4 && lane == 0) {
    if (p.out_bf16) tma_prefetch_desc(&tmC16);
    if (p.out_f32) tma_prefetch_desc(&tmC32);
    if (p.res) tma_prefetch_desc(&tmRes);
    if (p.act_in) tma_prefetch_desc(&tmAct);
    if (p.pre_out) tma_prefetch_desc(&tmPre);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EPI_WARPS);
    }
    if constexpr (FAST)
      for (int s = 0; s < EPI_WARPS * EPI_MAX_NBUF; ++s) mbar_init(&ld_bar[s], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();  // everything above is private to this CTA; global memory is touched only from here on

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int tile = unit / p.splits, sp = unit - tile * p.splits;
        const int m0 = (tile / p.num_n_tiles) * BM;
        const int n0 = (tile % p.num_n_tiles) * BN;
        const int seg = p.seg_n > 0 ? n0 / p.seg_n : 0;
        // stacked problems: a column segment shifts the contraction window of both operands and restarts B's columns;
        // a row segment shifts B's columns
        const int k_off = p.seg_k > 0 ? seg * p.seg_k : 0;
        const int bn0 = n0 - (p.seg_k > 0 ? seg * p.seg_n : 0) + (p.seg_m > 0 ? (m0 / p.seg_m) * p.seg_nb : 0);
        const int kb0 = (int)((long)sp * nk / p.splits), kb1 = (int)((long)(sp + 1) * nk / p.splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + Cfg::A_BYTES;
#ifdef TDM_DBG_NOTMA
          mbar_arrive(&full_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          continue;
#endif
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const CUtensorMap* mA = &tmA;
          const CUtensorMap* mB = &tmB;
          int ka = kb * BK + k_off, kbb = kb * BK + k_off;
          if (kb >= nk1) {
            mA = &tmA2;
            mB = &tmB2;
            kbb = (kb - nk1) * BK;
            ka = seg * p.seg_k2 + kbb;
          }
          if (!p.a_mn) {
            tma_load_2d(sA, mA, &full_bar[stage], ka, m0);
          } else {
            tma_load_2d(sA, mA, &full_bar[stage], m0, ka);
            tma_load_2d(sA + 8192, mA, &full_bar[stage], m0 + 64, ka);
          }
          if (!p.b_mn) {
            tma_load_2d(sB, mB, &full_bar[stage], kbb, bn0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + j * 8192, mB, &full_bar[stage], bn0 + 64 * j, kbb);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Warp-convergent: every lane polls the barriers and computes the (warp-uniform) descriptors, only the tcgen05
    // instructions sit under elect.sync.  The compiler then keeps the operands in uniform registers; issuing from
    // inside an `if (lane == 0)` region made it emit a register-broadcast loop per MMA (~30 SASS instructions,
    // ~180 clocks) that capped the tensor pipe at ~54 % on 96-clock MMAs.
    {
      const uint32_t idesc = make_idesc_bf16(BM, BN, p.a_mn, p.b_mn);
      const uint32_t a_kstep = (p.a_mn ? 2048u : 32u) >> 4;  // descriptor address units (16 B) per UMMA_K = 16
      const uint32_t b_kstep = (p.b_mn ? 2048u : 32u) >> 4;
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u;
      const uint32_t b_lbo = p.b_mn ? 8192u : 16u;
      const uint64_t a_desc0 = make_sdesc_sw128(smem_u32(smem), a_lbo, 1024);
      const uint64_t b_desc0 = make_sdesc_sw128(smem_u32(smem) + Cfg::A_BYTES, b_lbo, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const int tile = unit / p.splits, sp = unit - tile * p.splits;
        const int kb0 = (int)((long)sp * nk / p.splits), kb1 = (int)((long)(sp + 1) * nk / p.splits);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * Cfg::ACC_STRIDE;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = a_desc0 + uint64_t(stage * (Cfg::STAGE_BYTES >> 4));
            const uint64_t db = b_desc0 + uint64_t(stage * (Cfg::STAGE_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
#ifndef TDM_DBG_NOMMA
              umma_bf16(tmem_d, da + uint64_t(k * a_kstep), db + uint64_t(k * b_kstep), idesc, (kb > kb0) || (k != 0));
#endif
            }
            umma_commit(&empty_bar[stage]);  // frees the smem stage when these MMAs retire
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tfull_bar[as]);  // accumulator complete
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // thread == accumulator row (tcgen05.ld 32x32b): all 32 values of a chunk sit in registers, so the fused math
    // has 32-way ILP and every global access is a 16-byte vector; each thread touches whole 32-byte sectors.
    const int ew = warp - 4;
    if constexpr (FAST) {
      const EpiMaps maps = {&tmC16, &tmC32, &tmRes, &tmAct, &tmPre};
      gemm_epilogue_warp_loop<BN, false>(p, maps, tmem_base, Cfg::ACC_STRIDE, ew, lane,
                                         smem_u32(epi_stage) + ew * p.epi_nbuf * p.epi_set_bytes, ld_bar + ew * EPI_MAX_NBUF,
                                         smem_u32(epi_vec) + ew * 1024, tfull_bar, tempty_bar, blockIdx.x, gridDim.x, num_units, 0);
    } else {
      const int q = ew & 3;    // == warp % 4: the TMEM lane quarter this warp may read
      const int ch = ew >> 2;  // even / odd 32-column chunks of the tile
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const int tile = unit / p.splits;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        const int m0 = (tile / p.num_n_tiles) * BM;
        const int n0 = (tile % p.num_n_tiles) * BN;
        gemm_epilogue_stage_vectors<BN>(p, epi_vec, as, m0, n0, ew * 32 + lane);
        const uint32_t taddr = tmem_base + as * Cfg::ACC_STRIDE + (uint32_t(q * 32) << 16);
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
#ifndef TDM_DBG_NOEPI
        gemm_epilogue_rows<BN>(p, &tmC16, &tmC32, taddr, m0, n0, q, ch, lane, epi_stage + ew * EPI_STAGE_WARP, epi_vec, as);
#endif
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);
      }
    }
    if (lane == 0) tma_store_wait_all();  // the staging tiles must outlive the bulk stores that read them
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): one 256 x BN output tile per pair of SMs.  Each CTA stages its own 128 rows of A
// and HALF of the B tile, the leader CTA's single thread issues 256 x BN x 16 MMAs that read both CTAs' shared
// memory and write both CTAs' tensor memory.  Per CTA and 64-deep k-block that is (128 + BN/2) * 128 bytes of
// L2 -> SMEM traffic instead of (128 + BN) * 128: the mainloop of the single-CTA kernel is bound by exactly that
// stream (B300_MICROARCH.md: chip-wide LTS cap ~6300 B/clk), so the pair runs ~1.4x closer to the tensor peak.
// Barriers: `full` lives in the leader (both CTAs' TMA loads credit it), `empty` / `tfull` are signalled in both
// CTAs by multicast tcgen05.commit, `tempty` lives in the leader (the peer's epilogue warps arrive remotely).
// ------------------------------------------------------------------------------------------------
template <int BN>
struct Gemm2Cfg {
  static constexpr int A_BYTES = BM * BK * 2;        // this CTA's 128 rows
  static constexpr int B_BYTES = (BN / 2) * BK * 2;  // this CTA's half of the N extent
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ACC_STRIDE = (BN <= 128) ? 128 : 256;
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
};

template <int BN, bool FAST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                  const __grid_constant__ CUtensorMap tmC16, const __grid_constant__ CUtensorMap tmC32,
                  const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmAct,
                  const __grid_constant__ CUtensorMap tmPre, const GemmParams p) {
  using Cfg = Gemm2Cfg<BN>;
  const int STAGES = p.stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  static_assert(Cfg::STAGE_BYTES % 1024 == 0, "epilogue staging must stay 1024-byte aligned");
  uint8_t* epi_stage = smem + STAGES * Cfg::STAGE_BYTES;
  float* epi_vec = reinterpret_cast<float*>(epi_stage + p.epi_bytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + p.epi_bytes + EPI_VEC_BYTES);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* ld_bar = tempty_bar + 2;  // [EPI_WARPS][EPI_MAX_NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ld_bar + EPI_WARPS * EPI_MAX_NBUF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nk1 = (p.K + BK - 1) / BK;
  const int nk2 = (p.K2 + BK - 1) / BK;
  const int nk = nk1 + nk2;
  const int num_units = p.num_m_tiles * p.num_n_tiles * p.splits;  // num_m_tiles counts 256-row pair tiles here

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (nk2) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
  }
  if (FAST && warp == 4 && lane == 0) {
    if (p.out_bf16) tma_prefetch_desc(&tmC16);
    if (p.out_f32) tma_prefetch_desc(&tmC32);
    if (p.res) tma_prefetch_desc(&tmRes);
    if (p.act_in) tma_prefetch_desc(&tmAct);
    if (p.pre_out) tma_prefetch_desc(&tmPre);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * EPI_WARPS);
    }
    if constexpr (FAST)
      for (int s = 0; s < EPI_WARPS * EPI_MAX_NBUF; ++s) mbar_init(&ld_bar[s], 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers and tensor memory exist before anyone signals or writes across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = pair; unit < num_units; unit += npairs) {
        const int tile = unit / p.splits, sp = unit - tile * p.splits;
        int mt, nt;
        gemm_unit_tile(p, tile, mt, nt);
        const int m0 = mt * (2 * BM) + (int)rank * BM;
        const int n0 = nt * BN;
        // this CTA's half of the B tile; a narrow last tile is an N = BN / 2 MMA whose halves are BN / 4 rows each (the box
        // still brings BN / 2 rows: the MMA reads the first BN / 4 of them)
        const bool narrow = p.narrow_last && nt == p.num_n_tiles - 1;
        const int nh = n0 + (int)rank * (narrow ? BN / 4 : BN / 2);
        const int seg = p.seg_n > 0 ? n0 / p.seg_n : 0;
        const int kb0 = (int)((long)sp * nk / p.splits), kb1 = (int)((long)(sp + 1) * nk / p.splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + Cfg::A_BYTES;
          const uint32_t fb = smem_u32(&full_bar[stage]) & PEER_BIT_MASK;  // the leader's barrier, from either CTA
#ifdef TDM_DBG_NOTMA
          if (leader) mbar_arrive(&full_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          continue;
#endif
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          const CUtensorMap* mA = &tmA;
          const CUtensorMap* mB = &tmB;
          int ka = kb * BK, kbb = kb * BK;
          if (kb >= nk1) {
            mA = &tmA2;
            mB = &tmB2;
            kbb = (kb - nk1) * BK;
            ka = seg * p.seg_k2 + kbb;
          }
          if (!p.a_mn) {
            tma_load_2d_2cta(sA, mA, fb, ka, m0);
          } else {
            tma_load_2d_2cta(sA, mA, fb, m0, ka);
            tma_load_2d_2cta(sA + 8192, mA, fb, m0 + 64, ka);
          }
          if (!p.b_mn) {
            tma_load_2d_2cta(sB, mB, fb, kbb, nh);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j) tma_load_2d_2cta(sB + j * 8192, mB, fb, nh + 64 * j, kbb);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; warp-convergent, tcgen05 ops under elect.sync) =====================
    if (leader) {
      const uint32_t idesc = make_idesc_bf16(2 * BM, BN, p.a_mn, p.b_mn);
      const uint32_t a_kstep = (p.a_mn ? 2048u : 32u) >> 4;  // descriptor address units (16 B) per UMMA_K = 16
      const uint32_t b_kstep = (p.b_mn ? 2048u : 32u) >> 4;
      const uint32_t a_lbo = p.a_mn ? 8192u : 16u;
      const uint32_t b_lbo = p.b_mn ? 8192u : 16u;
      const uint64_t a_desc0 = make_sdesc_sw128(smem_u32(smem), a_lbo, 1024);
      const uint64_t b_desc0 = make_sdesc_sw128(smem_u32(smem) + Cfg::A_BYTES, b_lbo, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint32_t idesc_narrow = make_idesc_bf16(2 * BM, BN / 2, p.a_mn, p.b_mn);
      for (int unit = pair; unit < num_units; unit += npairs, ++it) {
        const int tile = unit / p.splits, sp = unit - tile * p.splits;
        const int kb0 = (int)((long)sp * nk / p.splits), kb1 = (int)((long)(sp + 1) * nk / p.splits);
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        uint32_t idesc_u = idesc;
        if (p.narrow_last) {
          int mt, nt;
          gemm_unit_tile(p, tile, mt, nt);
          if (nt == p.num_n_tiles - 1) idesc_u = idesc_narrow;
        }
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * Cfg::ACC_STRIDE;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t da = a_desc0 + uint64_t(stage * (Cfg::STAGE_BYTES >> 4));
            const uint64_t db = b_desc0 + uint64_t(stage * (Cfg::STAGE_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
#ifndef TDM_DBG_NOMMA
              umma_bf16_2cta(tmem_d, da + uint64_t(k * a_kstep), db + uint64_t(k * b_kstep), idesc_u, (kb > kb0) || (k != 0));
#endif
            }
            umma_commit_2cta(&empty_bar[stage]);  // frees the stage in BOTH CTAs when these MMAs retire
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit_2cta(&tfull_bar[as]);  // accumulator complete, both CTAs' epilogues may read
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs: each owns its 128 accumulator rows) =====================
    const int ew = warp - 4;
    if constexpr (FAST) {
      const EpiMaps maps = {&tmC16, &tmC32, &tmRes, &tmAct, &tmPre};
      gemm_epilogue_warp_loop<BN, true>(p, maps, tmem_base, Cfg::ACC_STRIDE, ew, lane,
                                        smem_u32(epi_stage) + ew * p.epi_nbuf * p.epi_set_bytes, ld_bar + ew * EPI_MAX_NBUF,
                                        smem_u32(epi_vec) + ew * 1024, tfull_bar, tempty_bar, pair, npairs, num_units, (int)rank);
    } else {
      const int q = ew & 3;
      const int ch = ew >> 2;
      int it = 0;
      for (int unit = pair; unit < num_units; unit += npairs, ++it) {
        const int tile = unit / p.splits;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        int mt, nt;
        gemm_unit_tile(p, tile, mt, nt);
        const int m0 = mt * (2 * BM) + (int)rank * BM;
        const int n0 = nt * BN;
        gemm_epilogue_stage_vectors<BN>(p, epi_vec, as, m0, n0, ew * 32 + lane);
        const uint32_t taddr = tmem_base + as * Cfg::ACC_STRIDE + (uint32_t(q * 32) << 16);
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
#ifndef TDM_DBG_NOEPI
        gemm_epilogue_rows<BN>(p, &tmC16, &tmC32, taddr, m0, n0, q, ch, lane, epi_stage + ew * EPI_STAGE_WARP, epi_vec, as);
#endif
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(smem_u32(&tempty_bar[as]) & PEER_BIT_MASK);  // the leader's barrier
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static int make_operand_map(CUtensorMap* map, const void* ptr, int mn_major, long rows_mn, long k, long ld,
                            int box_mn) {
  // K-major: stored [rows_mn, k], inner = k.  MN-major: stored [k, rows_mn], inner = rows_mn.
  uint64_t dims[2], strides[1];
  uint32_t box[2];
  if (!mn_major) {
    dims[0] = (uint64_t)k;
    dims[1] = (uint64_t)rows_mn;
    box[0] = BK;
    box[1] = (uint32_t)box_mn;
  } else {
    dims[0] = (uint64_t)rows_mn;
    dims[1] = (uint64_t)k;
    box[0] = 64;
    box[1] = BK;
  }
  strides[0] = (uint64_t)ld * 2;
  return make_tma_bf16(map, ptr, 2, dims, strides, box, TMA_SWZ_128);
}

// Split-K applies only to pure accumulation (dW += dy^T x style: res aliases out_f32, nothing else in the epilogue):
// partial sums are red.add'ed into out_f32, so the existing contents play the role of `res`.
static bool gemm_acc_only(const TdmGemm* g) {
  return g->res && g->res == g->out_f32 && g->ld_res == g->ld_out_f32 && !g->bias && !g->gate && g->act == TDM_ACT_NONE &&
         !g->out_bf16 && !g->pre_out && !g->pre_out_f32 && g->res_row_mod == 0 && g->K2 <= 0;
}
// Split factor when the output tiles alone cannot fill the machine and K is long: the s in [1, 6] (>= 8 k-blocks per
// unit) with the fewest waves per split, ceil(tiles * s / slots) / s; ties go to fewer partial sums to reduce in HBM.
static int gemm_plan_splits(const TdmGemm* g, int tiles0, int slots) {
  const int nk = (g->K + BK - 1) / BK;
  if (g->split_k > 1) return g->split_k < nk ? g->split_k : nk;
  if (g->split_k == 1 || !gemm_acc_only(g) || nk < 16) return 1;
  if (tiles0 >= slots) {  // enough tiles: halve K only to repair a badly filled last wave (2.2 waves -> 3 half waves)
    const int w1 = (tiles0 + slots - 1) / slots;
    const double w2 = double((2 * tiles0 + slots - 1) / slots) / 2.0;
    return (nk >= 32 && w2 < 0.87 * w1) ? 2 : 1;
  }
  int best = 1;
  double best_c = 1e30;
  for (int sp = 1; sp <= 6 && sp <= nk / 8; ++sp) {
    const double c = double((tiles0 * sp + slots - 1) / slots) / sp;
    if (c < best_c - 1e-9) {
      best_c = c;
      best = sp;
    }
  }
  return best;
}

// Positions of the narrow (half-cost) tiles in the round-robin unit sequence of `slots` pairs: the last round's positions
// (pairs 0 .. L-1 carry one unit more than the others), then the same pairs' positions one round earlier (two halves ~ one
// full tile), the rest on the highest pairs of that round.  Left off (identity order) when the pattern does not fit.
static void gemm_plan_balance(GemmParams* p, int slots) {
  const int T = p->num_m_tiles * p->num_n_tiles, H = p->num_m_tiles;
  const int P = T < slots ? T : slots;
  const int R = (T + P - 1) / P;
  if (R < 2) return;
  const int L = T - (R - 1) * P;  // units of the last round
  if (L == P) return;             // full last round: nothing to balance
  const int c = H < L ? H : L;
  const int h1 = H - c;
  const int a = h1 < L ? h1 : L;
  const int h2 = h1 - a;
  if (h2 > P - a) return;
  p->bal = 1;
  p->bal_a0 = (R - 2) * P; p->bal_a1 = p->bal_a0 + a;
  p->bal_b1 = (R - 1) * P; p->bal_b0 = p->bal_b1 - h2;
  p->bal_c0 = (R - 1) * P; p->bal_c1 = p->bal_c0 + c;
}

template <int BN, bool PAIR>
static int launch_gemm(const TdmGemm* g, cudaStream_t stream) {
  constexpr int STAGE_BYTES = PAIR ? Gemm2Cfg<BN>::STAGE_BYTES : GemmCfg<BN>::STAGE_BYTES;
  constexpr int TILE_M = PAIR ? 2 * BM : BM;   // rows of one work unit
  constexpr int BOX_N = PAIR ? BN / 2 : BN;    // B rows one CTA stages per k-block
  CUtensorMap tmA, tmB, tmA2, tmB2;
  int rc;
  // stacked problems: the maps span every segment (A / B contraction extent nseg * seg_k, B columns per segment)
  const long nseg_col = g->seg_k > 0 ? (g->N + g->seg_n - 1) / g->seg_n : 1;
  const long nseg_row = g->seg_m > 0 ? (g->M + g->seg_m - 1) / g->seg_m : 1;
  const long k_map = g->seg_k > 0 ? nseg_col * g->seg_k : g->K;
  const long n_map = g->seg_k > 0 ? g->seg_n : (g->seg_m > 0 ? g->N + (nseg_row - 1) * g->seg_nb : g->N);
  if ((rc = make_operand_map(&tmA, g->A, g->a_mn_major, g->M, k_map, g->lda, BM))) return rc;
  if ((rc = make_operand_map(&tmB, g->B, g->b_mn_major, n_map, k_map, g->ldb, BOX_N))) return rc;
  if (g->K2 > 0) {
    const long a2_cols = g->seg_n > 0 ? (long)((g->N + g->seg_n - 1) / g->seg_n) * g->seg_k2 : g->K2;
    if ((rc = make_operand_map(&tmA2, g->A2, g->a_mn_major, g->M, a2_cols, g->lda2, BM))) return rc;
    if ((rc = make_operand_map(&tmB2, g->B2, g->b_mn_major, g->N, g->K2, g->ldb2, BOX_N))) return rc;
  } else {
    tmA2 = tmA;
    tmB2 = tmB;
  }
  // output tensor maps for the TMA-store epilogue: [M, N] row-major, 32 x 32 boxes
  CUtensorMap tmC16 = tmA, tmC32 = tmA;
  {
    const uint64_t dims[2] = {(uint64_t)g->N, (uint64_t)g->M};
    const uint32_t box[2] = {32, 32};
    if (g->out_bf16) {
      const uint64_t st[1] = {(uint64_t)g->ld_out_bf16 * 2};
      if ((rc = make_tma_typed(&tmC16, g->out_bf16, 0, 2, dims, st, box, TMA_SWZ_64))) return rc;
    }
    if (g->out_f32) {
      const uint64_t st[1] = {(uint64_t)g->ld_out_f32 * 4};
      if ((rc = make_tma_typed(&tmC32, g->out_f32, 1, 2, dims, st, box, TMA_SWZ_128))) return rc;
    }
  }
  GemmParams p;
  p.M = g->M; p.N = g->N; p.K = g->K; p.K2 = g->K2 > 0 ? g->K2 : 0;
  p.a_mn = g->a_mn_major; p.b_mn = g->b_mn_major;
  p.seg_n = g->seg_n; p.seg_k2 = g->seg_k2;
  p.seg_k = g->seg_k; p.seg_m = g->seg_m; p.seg_nb = g->seg_nb;
  p.num_m_tiles = (g->M + TILE_M - 1) / TILE_M;
  p.num_n_tiles = (g->N + BN - 1) / BN;
  p.alpha = g->alpha; p.act = g->act; p.bias = g->bias;
  p.act_in = (const bf16*)g->act_in; p.ld_act = g->ld_act;
  p.pre_out = (bf16*)g->pre_out; p.ld_pre = g->ld_pre;
  p.pre_out_f32 = g->pre_out_f32; p.ld_pre_f32 = g->ld_pre_f32; p.res_row_mod = g->res_row_mod;
  p.gate = g->gate; p.rows_per_gate = g->rows_per_gate > 0 ? g->rows_per_gate : 1; p.gate_stride = g->gate_stride;
  p.res = g->res; p.ld_res = g->ld_res;
  p.out_f32 = g->out_f32; p.ld_out_f32 = g->ld_out_f32;
  p.out_bf16 = (bf16*)g->out_bf16; p.ld_out_bf16 = g->ld_out_bf16;
  if (g->split_k > 1 && !gemm_acc_only(g))
    return set_error(TDM_ERR_INVALID, "tdm_gemm_bf16: split_k needs an accumulate-only epilogue (res == out_f32)");
  p.splits = gemm_plan_splits(g, p.num_m_tiles * p.num_n_tiles, PAIR ? num_sms() / 2 : num_sms());
  p.narrow_last = p.bal = 0;
  p.bal_a0 = p.bal_a1 = p.bal_b0 = p.bal_b1 = p.bal_c0 = p.bal_c1 = 0;
  if constexpr (PAIR && BN == 256) {
    const bool no_bal = getenv("TDM_B200_GEMM_BALANCE") && atoi(getenv("TDM_B200_GEMM_BALANCE")) == 0;  // A/B timing
    const int tail = g->N - (p.num_n_tiles - 1) * BN;  // width of the last column tile
    if (!no_bal && p.splits == 1 && p.num_n_tiles >= 2 && tail <= BN / 2 && g->seg_n == 0 && g->seg_k == 0 && g->seg_m == 0) {
      p.narrow_last = 1;
      gemm_plan_balance(&p, num_sms() / 2);
    }
  }

  // Epilogue flavour and shared-memory plan.  The streaming (FAST) epilogue moves every epilogue tensor by TMA; the
  // generic one keeps the rare features: split-K reductions, the row-modulo residual (pos_embed), the unrounded
  // pre-activation copy and gates that change inside a 128-row tile.
  p.gate_uniform = (p.rows_per_gate % BM) == 0;
  if (!(g->act == TDM_ACT_GELU_TANH_BWD || g->act == TDM_ACT_SILU_BWD)) p.act_in = nullptr;  // unused otherwise
  static const bool force_generic = getenv("TDM_B200_GEMM_GENERIC") && atoi(getenv("TDM_B200_GEMM_GENERIC")) != 0;
  const bool fast_ok = p.splits == 1 && g->res_row_mod == 0 && !g->pre_out_f32 && (!g->gate || p.gate_uniform) &&
                    (((uintptr_t)g->res | (uintptr_t)p.act_in | (uintptr_t)g->pre_out) & 15) == 0;
  if (g->epilogue >= 2 && !fast_ok)
    return set_error(TDM_ERR_INVALID, "tdm_gemm_bf16: the streaming epilogue does not cover this call (split-K, res_row_mod, pre_out_f32, "
                                      "a gate that changes inside a tile, or unaligned epilogue tensors)");
  const bool fast = fast_ok && g->epilogue != 1 && !(force_generic && g->epilogue == 0);
  constexpr int AVAIL = SMEM_LIMIT - 1024 - BAR_BYTES - EPI_VEC_BYTES;
  CUtensorMap tmRes = tmA, tmAct = tmA, tmPre = tmA;
  p.epi_nbuf = 1; p.epi_set_bytes = 0; p.epi_off_h0 = p.epi_off_h1 = 0; p.epi_in_bytes = 0; p.epi_feat_slot = -1;
  if (fast) {
    const bool need_f = g->res || g->out_f32, need_h0 = p.act_in || g->out_bf16, need_h1 = g->pre_out != nullptr;
    p.epi_off_h0 = need_f ? EPI_STAGE_F32 : 0;
    p.epi_off_h1 = p.epi_off_h0 + (need_h0 ? EPI_STAGE_BF16 : 0);
    p.epi_set_bytes = (p.epi_off_h1 + (need_h1 ? EPI_STAGE_BF16 : 0) + 1023) & ~1023;
    p.epi_in_bytes = (g->res ? EPI_STAGE_F32 : 0) + (p.act_in ? EPI_STAGE_BF16 : 0);
    // three sets per warp when something is loaded (load one chunk ahead without waiting for the youngest store), as
    // long as the operand ring keeps >= 4 stages; two otherwise
    p.epi_nbuf = 2;
    if (p.epi_in_bytes > 0 && (AVAIL - EPI_WARPS * 3 * p.epi_set_bytes) / STAGE_BYTES >= 4) p.epi_nbuf = 3;
    p.epi_bytes = EPI_WARPS * p.epi_nbuf * p.epi_set_bytes;
    const bool plain_act = g->act == TDM_ACT_NONE || g->act == TDM_ACT_GELU_TANH || g->act == TDM_ACT_GELU_TANH_BWD;
    const int feat = (g->out_bf16 ? EF_OUT16 : 0) | (g->out_f32 ? EF_OUT32 : 0) | (g->res ? EF_RES : 0) | (g->gate ? EF_GATE : 0) |
                     (g->pre_out ? EF_PRE : 0) | (g->act == TDM_ACT_GELU_TANH ? EF_GELU : 0) |
                     (g->act == TDM_ACT_GELU_TANH_BWD ? EF_GELU_BWD : 0);
    static const bool force_dyn = getenv("TDM_B200_GEMM_DYN") && atoi(getenv("TDM_B200_GEMM_DYN")) != 0;
    p.epi_feat_slot = (plain_act && g->epilogue != 2 && !(force_dyn && g->epilogue == 0)) ? epi_feat_slot(feat) : -1;
    const uint64_t dims[2] = {(uint64_t)g->N, (uint64_t)g->M};
    const uint32_t box[2] = {32, 32};
    if (g->res) {
      const uint64_t st[1] = {(uint64_t)g->ld_res * 4};
      if ((rc = make_tma_typed(&tmRes, g->res, 1, 2, dims, st, box, TMA_SWZ_128))) return rc;
    }
    if (p.act_in) {
      const uint64_t st[1] = {(uint64_t)g->ld_act * 2};
      if ((rc = make_tma_typed(&tmAct, p.act_in, 0, 2, dims, st, box, TMA_SWZ_64))) return rc;
    }
    if (g->pre_out) {
      const uint64_t st[1] = {(uint64_t)g->ld_pre * 2};
      if ((rc = make_tma_typed(&tmPre, g->pre_out, 0, 2, dims, st, box, TMA_SWZ_64))) return rc;
    }
  } else {
    p.epi_bytes = EPI_GENERIC_BYTES;
  }
  p.stages = (AVAIL - p.epi_bytes) / STAGE_BYTES;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  if (p.stages < 2) return set_error(TDM_ERR_INVALID, "tdm_gemm_bf16: no shared memory left for the operand ring (tile_n=%d)", BN);
  const int smem = gemm_smem_bytes(p.stages, STAGE_BYTES, p.epi_bytes);

  const int units = p.num_m_tiles * p.num_n_tiles * p.splits;
  const int sms = num_sms();
  static bool attr_done[2] = {false, false};  // per template instantiation: [generic, fast]
  auto launch = [&](auto kernel, int grid) -> int {
    bool& attr_set = attr_done[fast ? 1 : 0];
    if (!attr_set) {
      TDM_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
      attr_set = true;
    }
    ::tdm::count_launch();
    static const bool pdl_gemm = !(getenv("TDM_B200_PDL_GEMM") && atoi(getenv("TDM_B200_PDL_GEMM")) == 0);
    TDM_CHECK_CUDA(launch_pdl_if(pdl_gemm, kernel, dim3(grid), dim3(GEMM_THREADS), smem, stream, tmA, tmB, tmA2, tmB2, tmC16, tmC32,
                                 tmRes, tmAct, tmPre, p));
    return TDM_OK;
  };
  if constexpr (PAIR) {
    const int pairs = units < sms / 2 ? units : sms / 2;
    if (fast) { if ((rc = launch(gemm2_bf16_kernel<BN, true>, 2 * pairs))) return rc; }
    else { if ((rc = launch(gemm2_bf16_kernel<BN, false>, 2 * pairs))) return rc; }
  } else {
    const int grid = units < sms ? units : sms;
    if (fast) { if ((rc = launch(gemm_bf16_kernel<BN, true>, grid))) return rc; }
    else { if ((rc = launch(gemm_bf16_kernel<BN, false>, grid))) return rc; }
  }
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}

}  // namespace tdm

using namespace tdm;

#ifdef TDM_DBG_TIMING
extern "C" int tdm_dbg_read(long long* out, int max_pairs) {
  int n = max_pairs < 4096 ? max_pairs : 4096;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, tdm_dbg_times, sizeof(long long) * n);
  static uint32_t zeros[32 * 8];
  cudaMemcpyToSymbol(tdm_dbg_cj, zeros, sizeof(zeros));
  static long long zt[8192];
  cudaMemcpyToSymbol(tdm_dbg_times, zt, sizeof(zt));
  return n;
}
#endif

extern "C" int tdm_gemm_bf16(const TdmGemm* g, void* stream) {
  TDM_REQUIRE(g != nullptr, "tdm_gemm_bf16: null descriptor");
  TDM_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "tdm_gemm_bf16: M,N,K must be positive (got %d,%d,%d)", g->M,
              g->N, g->K);
  TDM_REQUIRE(g->A && g->B, "tdm_gemm_bf16: A and B are required");
  TDM_REQUIRE(g->out_f32 || g->out_bf16 || g->pre_out || g->pre_out_f32, "tdm_gemm_bf16: no output given");
  TDM_REQUIRE(g->N % 8 == 0, "tdm_gemm_bf16: N=%d must be a multiple of 8", g->N);
  TDM_REQUIRE(g->lda % 8 == 0 && g->ldb % 8 == 0, "tdm_gemm_bf16: lda/ldb must be multiples of 8");
  TDM_REQUIRE(((uintptr_t)g->A & 15) == 0 && ((uintptr_t)g->B & 15) == 0, "tdm_gemm_bf16: A/B must be 16-byte aligned");
  TDM_REQUIRE((((uintptr_t)g->out_f32 | (uintptr_t)g->out_bf16) & 15) == 0, "tdm_gemm_bf16: outputs must be 16-byte aligned");
  TDM_REQUIRE(g->act >= TDM_ACT_NONE && g->act <= TDM_ACT_SILU_BWD, "tdm_gemm_bf16: unknown act %d", g->act);
  TDM_REQUIRE((g->act != TDM_ACT_GELU_TANH_BWD && g->act != TDM_ACT_SILU_BWD) || g->act_in,
              "tdm_gemm_bf16: backward activations need act_in");
  if (g->K2 > 0) {
    TDM_REQUIRE(g->A2 && g->B2, "tdm_gemm_bf16: K2>0 needs A2 and B2");
    TDM_REQUIRE(g->lda2 % 8 == 0 && g->ldb2 % 8 == 0, "tdm_gemm_bf16: lda2/ldb2 must be multiples of 8");
    if (g->seg_n > 0)
      TDM_REQUIRE(g->seg_k2 % 64 == 0 && g->K2 <= g->seg_k2, "tdm_gemm_bf16: segmented pair needs seg_k2 %% 64 == 0 and K2 <= seg_k2");
  }
  TDM_REQUIRE((!g->out_f32 || g->ld_out_f32 % 4 == 0) && (!g->out_bf16 || g->ld_out_bf16 % 8 == 0) &&
                  (!g->res || g->ld_res % 4 == 0) && (!g->pre_out || g->ld_pre % 8 == 0) &&
                  (!g->pre_out_f32 || g->ld_pre_f32 % 4 == 0) &&
                  (!g->act_in || g->ld_act % 8 == 0) && (!g->gate || g->gate_stride % 4 == 0),
              "tdm_gemm_bf16: epilogue leading dimensions must keep 16-byte alignment");
  // Tile / mode choice, calibrated on B200 measurements under the sustained power cap (scripts/bench_gemm_one.py,
  // profiles/): the mainloop is bound by the L2 -> shared-memory operand stream, so a tile costs its operand rows per
  // k-block -- (BN + 128) for a CTA pair (each CTA stages 128 rows of A and BN / 2 rows of B for 128 x BN outputs... per
  // SM: (BN + 256) / 2 rows per 128 x BN), 1.5 x (BN + 64) for a single CTA -- times the number of waves over the SM
  // pairs / SMs.  8192 x {1152, 3456, 4608} x 1152: pair-192 26.1 us, pair-256 62.7 / 73.7 us against 27.0 / 69.4 /
  // 83.9 us single-CTA (cuBLAS in the same loop: 25.7 / 57.4 / 73.7 us).
  const int sms = num_sms();
  int bn = g->tile_n;
  int pair = g->cta_pair;  // 0 choose, 1 single CTA, 2 CTA pair
  static const int env_pair = getenv("TDM_B200_GEMM_PAIR") ? atoi(getenv("TDM_B200_GEMM_PAIR")) : 0;  // debugging aid
  if (pair == 0 && (env_pair == 1 || g->seg_k > 0 || g->seg_m > 0)) pair = 1;
  if (g->seg_k > 0 || g->seg_m > 0) {
    TDM_REQUIRE(g->K2 <= 0, "tdm_gemm_bf16: stacked problems take no second operand pair");
    TDM_REQUIRE(g->seg_k <= 0 || (g->seg_n > 0 && g->seg_k >= g->K && g->K % 64 == 0),
                "tdm_gemm_bf16: seg_k needs seg_n > 0, seg_k >= K and K %% 64 == 0");
    TDM_REQUIRE(g->seg_m <= 0 || (g->seg_m % 128 == 0 && g->seg_nb > 0), "tdm_gemm_bf16: seg_m must be a multiple of 128 with seg_nb > 0");
    TDM_REQUIRE(g->cta_pair != 2, "tdm_gemm_bf16: stacked problems run on the single-CTA kernel");
  }
  TDM_REQUIRE(g->epilogue >= 0 && g->epilogue <= 3, "tdm_gemm_bf16: epilogue must be 0 (choose), 1 (generic), 2 or 3 (streaming)");
  TDM_REQUIRE(pair >= 0 && pair <= 2, "tdm_gemm_bf16: cta_pair must be 0 (choose), 1 (single) or 2 (pair)");
  auto cost = [&](int b, bool pr) {
    const int mt = pr ? (g->M + 255) / 256 : (g->M + 127) / 128;
    const int nt = (g->N + b - 1) / b;
    const int slots = pr ? sms / 2 : sms;
    const int sp = gemm_plan_splits(g, mt * nt, slots);
    // a ragged last column tile fetches (and so costs) only its valid part: count it as a fraction of a tile
    const double units = double(mt) * (double(g->N) / b) * sp;
    const double waves = ceil(units / slots - 0.02);
    // three epilogue tensors per element (residual in, fp32 + bf16 out: the O projection): the tile is bound by its
    // epilogue bytes, which scale with the tile width alone, so three rounds of 192-wide tiles beat three rounds of 256
    // (measured 34.0 vs 37.1 us at 8192 x 1152 x 1152; scripts/bench_gemm_tiles.py)
    // (the same holds for FF2 -- gate + residual, K = 4608: 84.1 vs 86.5 us -- but not for o2, K = 1152: 29.7 vs 28.7)
    const double epi = (b == 256 && g->res && g->out_f32 && (g->out_bf16 || (g->gate && g->K >= 4096))) ? 1.3 : 1.0;
    return epi * (pr ? waves * double(b + 128) : waves * double(b + 64) * 1.5) / sp;
  };
  auto allowed = [&](int b, bool pr) {
    if (g->seg_n > 0 && g->seg_n % b != 0) return false;
    if (pr && (b < 128 || g->M <= 128)) return false;
    if (pr && g->b_mn_major && b % 128 != 0) return false;  // each CTA stages whole 64-column groups of its half
    return true;
  };
  if (bn == 0 || pair == 0) {
    const int cands[4] = {256, 192, 128, 64};
    double best = 1e30;
    int best_bn = 0, best_pair = 1;
    for (int pm = 2; pm >= 1; --pm) {
      if (pair != 0 && pair != pm) continue;
      for (int i = 0; i < 4; ++i) {
        const int b = cands[i];
        if (bn != 0 && b != bn) continue;
        if (!allowed(b, pm == 2)) continue;
        const double c = cost(b, pm == 2);
        if (c < best * 0.97) {  // prefer the earlier (wider, paired) candidate unless a later one is clearly better
          best = c;
          best_bn = b;
          best_pair = pm;
        }
      }
    }
    TDM_REQUIRE(best_bn != 0, "tdm_gemm_bf16: no tile configuration fits tile_n=%d cta_pair=%d seg_n=%d", g->tile_n, g->cta_pair, g->seg_n);
    bn = best_bn;
    pair = best_pair;
  }
  TDM_REQUIRE(allowed(bn, pair == 2), "tdm_gemm_bf16: tile_n=%d cta_pair=%d not valid for this problem (seg_n=%d, M=%d, b_mn=%d)", bn, pair,
              g->seg_n, g->M, g->b_mn_major);
  cudaStream_t s = (cudaStream_t)stream;
  if (pair == 2) {
    switch (bn) {
      case 128: return launch_gemm<128, true>(g, s);
      case 192: return launch_gemm<192, true>(g, s);
      case 256: return launch_gemm<256, true>(g, s);
      default: return set_error(TDM_ERR_INVALID, "tdm_gemm_bf16: CTA-pair tiles are 128, 192 or 256 wide (got %d)", bn);
    }
  }
  switch (bn) {
    case 64: return launch_gemm<64, false>(g, s);
    case 128: return launch_gemm<128, false>(g, s);
    case 192: return launch_gemm<192, false>(g, s);
    case 256: return launch_gemm<256, false>(g, s);
    default: return set_error(TDM_ERR_INVALID, "tdm_gemm_bf16: tile_n must be 0, 64, 128, 192 or 256 (got %d)", bn);
  }
}
