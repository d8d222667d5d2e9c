// Flash attention forward on tcgen05 / TMEM for head_dim <= 128 (PixArt: 72), self- and cross-attention.
// Replaces xformers.memory_efficient_attention (reference main.py:177-178, 372) inside the diffusers
// Attention processor:  O = softmax(Q K^T * scale + bias) V,  bias = (1 - mask) * -10000 broadcast over
// heads and queries.
//
// One CTA per (128 queries, head, batch); 192 threads:
//   warps 0-3 : softmax + O correction + epilogue (thread == query row == TMEM lane)
//   warp 4    : TMA producer (Q once, K/V tiles of 128 keys double-buffered)
//   warp 5    : MMA issuer  (S = Q K^T into a ping-pong TMEM buffer, O += P V)
// head_dim 72 is not a multiple of 64: every tile is loaded as ceil(d/64) 128-byte-swizzled column
// chunks straight from the [tokens, heads*d] activation matrix with a 4-D tensor map (d, S, H, B);
// TMA zero-fills columns >= d and rows >= S, so no padded copy of Q/K/V ever exists in HBM.
#include <stdlib.h>
#include "common.cuh"

namespace tdm {

constexpr int ATT_BM = 128;  // queries per CTA
constexpr int ATT_BN = 128;  // keys per tile
constexpr int CHUNK_BYTES = 128 * 128;  // one [128 rows x 64 bf16] swizzled chunk

struct AttnFwdParams {
  int B, H, Sq, Sk, d;
  int nd;         // ceil(d / 64) column chunks
  int kv_stages;  // K / V ring depth (3 when it fits in shared memory, else 2)
  int dk16;       // ceil(d / 16) MMA k-steps for Q K^T
  float scale_log2;  // softmax scale * log2(e)
  const float* bias; // [B, Sk] additive, or null
  bf16* O;        // [B, Sq, H, d] with strides
  long o_stride_b, o_stride_s, o_stride_h;
  float* lse;     // [B, H, Sq] natural-log sum-exp (for backward), or null
  long long* trace;  // DBG & 64: clock64 stamps of CTA 0, [tile][16 events] (TDM_B200_ATTN_TRACE=<file>)
};
constexpr int TRACE_TILES = 96;
#define ATT_TRACE(g_, e_)                                                                        \
  do {                                                                                           \
    if constexpr (DBG & 64) {                                                                    \
      if (blockIdx.x == 0 && (g_) < TRACE_TILES) p.trace[(g_) * 16 + (e_)] = clock64();           \
    }                                                                                            \
  } while (0)

__device__ __forceinline__ void tmem_st8w(uint32_t taddr, const float* v) {
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(v[i]);
  tmem_st8(taddr, r);
}

// Persistent kernel: one CTA per SM walks over (query tile, head, batch) work items; the K/V ring, the S ping-pong
// and every barrier phase run continuously across items, Q is double-buffered and the next item's Q / first K,V
// tiles are in flight while the current item finishes, so the per-item fixed cost (barrier set-up, TMEM allocation,
// descriptor fetch, first-load latency -- ~5 us of a 17 us CTA in the one-tile-per-CTA version) is paid once per SM.
// TMEM map: S ping-pong at columns [0,128) and [128,256); O at [256, 256+pad16(d)).  P(g) (bf16, two keys per
// 32-bit column) overwrites columns [0,64) of the S buffer it was computed from and is the A operand of the
// P V contraction directly from tensor memory: P never touches shared memory.
// Two softmax warpgroups share every 128-query tile: warps 0-3 own keys [0,64) of the tile, warps 4-7 keys
// [64,128) (thread == query row in both); they exchange row maxima through shared memory.
// Two TMA producer warps: Q + K (warp 8) and V (warp 10) run independently.  With one in-order producer thread the
// K(g+1) load queued behind the wait for V(g)'s stage (free only when P V(g-2) retired), which left ONE tile of
// prefetch distance: the loop period was the TMA round trip (~2800 clk) plus the P V time -- removing every softmax
// instruction did not change the kernel time (profiles/r02_attn_ablation.jsonl).
constexpr int FWD_THREADS = 352;  // 8 softmax warps + Q/K TMA warp + MMA warp + V TMA warp

// DBG != 0 instantiations are timing ablations (scripts/attn_ablate.py; results are WRONG by construction):
//   1 no MUFU (ex2 replaced by a move)   2 no half-row maxima exchange   4 no P store to TMEM   8 no S load from TMEM
//   16 no running-max loop   64 clock64 trace   128 no P V MMAs   256 no Q K^T MMAs   512 no K / V TMA loads
// (a variant with every fourth exponential as a degree-3 polynomial on the FMA pipe measured no faster: 90.4 vs 88.8 us)
// What they showed (profiles/r02_attn_ablation*.jsonl, r02_attn_trace_*.txt): with EVERY MMA, TMA load and softmax
// instruction removed (927) the kernel still takes 83 of 88 us -- the loop is bound by the latency of its own
// mbarrier / tcgen05.commit hand-offs (300-900 clk per hop, two hops per tile on the critical cycle), not by the
// tensor pipe (640 clk per tile, 20 % active), MUFU (1024-1300 clk per tile) or TMA.
template <int DBG>
__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tile_bytes = p.nd * CHUNK_BYTES;
  const int KS = p.kv_stages;
  uint8_t* sQ = smem;                      // 2 stages
  uint8_t* sK = sQ + 2 * tile_bytes;       // KS stages
  uint8_t* sV = sK + KS * tile_bytes;      // KS stages
  float* sBias = reinterpret_cast<float*>(sV + KS * tile_bytes);  // 2 x 128 floats
  float* sMax = sBias + 256;                                      // [2 tiles][2 halves][128 rows]
  float* sL = sMax + 512;                                         // [2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sL + 256);
  uint64_t* q_full = bars;        // [2]
  uint64_t* q_empty = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;    // [2]
  uint64_t* p_full = bars + 6;    // [2] by tile parity (bars + 6, bars + 7): see the MMA warp's wait
  uint64_t* pv_done = bars + 21;  // [2] by tile parity (see the softmax warps' waits)
  uint64_t* o_free = bars + 8;
  uint64_t* k_full = bars + 9;    // [3]
  uint64_t* v_full = bars + 12;   // [3]
  uint64_t* k_empty = bars + 15;  // [3]  K stage free as soon as Q K^T of the tile retired (a whole tile earlier than V)
  uint64_t* v_empty = bars + 18;  // [3]  V stage free: P V of the tile retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = (p.Sk + ATT_BN - 1) / ATT_BN;
  const int n_qt = (p.Sq + ATT_BM - 1) / ATT_BM;
  const int n_items = n_qt * p.H * p.B;
  const int my_items = ((int)blockIdx.x < n_items) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total_tiles = my_items * n_kv;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    mbar_init(&p_full[0], 8);   // one elected arrival per softmax warp (256 per-thread arrivals serialise on the barrier word)
    mbar_init(&p_full[1], 8);
    mbar_init(&pv_done[0], 1);
    mbar_init(&pv_done[1], 1);
    mbar_init(o_free, 8);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;        // 2 x 128 columns
  const uint32_t tmem_O = tmem_base + 256;  // <= 128 columns
  pdl_sync();

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      int g = 0, st = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_items; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int q0 = (item % n_qt) * ATT_BM, h = (item / n_qt) % p.H, b = item / (n_qt * p.H);
        const int qs = it & 1;
        mbar_wait(&q_empty[qs], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qs], tile_bytes);
        for (int c = 0; c < p.nd; ++c) tma_load_4d(sQ + qs * tile_bytes + c * CHUNK_BYTES, &tmQ, &q_full[qs], c * 64, q0, h, b);
        for (int j = 0; j < n_kv; ++j, ++g) {
          mbar_wait(&k_empty[st], ph ^ 1);
          ATT_TRACE(g, 6);
          if constexpr (DBG & 512) {
            mbar_arrive(&k_full[st]);
          } else {
          mbar_arrive_expect_tx(&k_full[st], tile_bytes);
          for (int c = 0; c < p.nd; ++c)
            tma_load_4d(sK + st * tile_bytes + c * CHUNK_BYTES, &tmK, &k_full[st], c * 64, j * ATT_BN, h, b);
          }
          if (++st == KS) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 10) {
    // ===================== TMA producer: V =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmV);
      int g = 0, st = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_items; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int h = (item / n_qt) % p.H, b = item / (n_qt * p.H);
        for (int j = 0; j < n_kv; ++j, ++g) {
          mbar_wait(&v_empty[st], ph ^ 1);
          ATT_TRACE(g, 7);
          if constexpr (DBG & 512) {
            mbar_arrive(&v_full[st]);
          } else {
          mbar_arrive_expect_tx(&v_full[st], tile_bytes);
          for (int c = 0; c < p.nd; ++c)
            tma_load_4d(sV + st * tile_bytes + c * CHUNK_BYTES, &tmV, &v_full[st], c * 64, j * ATT_BN, h, b);
          }
          if (++st == KS) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    // Warp-convergent (every lane polls and computes the warp-uniform descriptors), tcgen05 ops under elect.sync:
    // operands stay in uniform registers.  The tile stream is flat across items: S(g+1) = Q K^T of the NEXT tile --
    // possibly the next item's first -- is issued before P(g) V(g), so the tensor pipe never drains between items.
    const uint32_t idesc_qk = make_idesc_bf16(128, ATT_BN, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(128, p.dk16 * 16, 0, 1);  // A = P from TMEM; B = V MN-major (d contiguous)
    const uint64_t q_desc0 = make_sdesc_sw128(smem_u32(sQ), 16, 1024);
    const uint64_t k_desc0 = make_sdesc_sw128(smem_u32(sK), 16, 1024);
    const uint64_t v_desc0 = make_sdesc_sw128(smem_u32(sV), CHUNK_BYTES, 1024);
    const uint32_t stage_step = uint32_t(tile_bytes) >> 4;  // descriptor address field counts 16-byte units
    const int dk16 = p.dk16;
    // (one polling lane + __syncwarp instead of 32 polling lanes shortened the waits in the clock64 trace but measured
    // slower end to end: 100 vs 91 us in every warp, 92.6 vs 86.5 us in this warp alone)
    auto wait1 = [&](uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); };
    // running tile coordinates of the two issue streams (no integer divisions on the issue path)
    struct Cur { int it, j, st; uint32_t ph; };
    auto advance = [&](Cur& c) {
      if (++c.j == n_kv) { c.j = 0; ++c.it; }
      if (++c.st == KS) { c.st = 0; c.ph ^= 1; }
    };
    Cur cq{0, 0, 0, 0}, cp{0, 0, 0, 0};
    auto issue_qk = [&](int g) {
      if (lane == 0) ATT_TRACE(g, 0);
      if (cq.j == 0) wait1(&q_full[cq.it & 1], (cq.it >> 1) & 1);
      wait1(&k_full[cq.st], cq.ph);
      if (lane == 0) ATT_TRACE(g, 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t qd = q_desc0 + uint64_t((cq.it & 1) * stage_step);
        const uint64_t kd = k_desc0 + uint64_t(cq.st * stage_step);
        const uint32_t sd = tmem_S + (g & 1) * 128;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (k < dk16 && !(DBG & 256)) {  // 16 head-dim columns per step: four steps per 64-column chunk
            const uint32_t off = ((k >> 2) * CHUNK_BYTES + (k & 3) * 32) >> 4;
            umma_bf16(sd, qd + off, kd + off, idesc_qk, k != 0);
          }
        }
        umma_commit(&s_full[g & 1]);
        umma_commit(&k_empty[cq.st]);
        if (cq.j == n_kv - 1) umma_commit(&q_empty[cq.it & 1]);  // last read of this item's Q tile
      }
      __syncwarp();
      advance(cq);
      if (lane == 0) ATT_TRACE(g, 8);
    };
    if (total_tiles > 0) issue_qk(0);
    for (int g = 0; g < total_tiles; ++g) {
      // S((g+1)&1) still holds P(g-1): tcgen05.mma executes in issue order, so this S write follows P V (g-1)
      if (g + 1 < total_tiles) issue_qk(g + 1);
      // S(g+1) is issued BEFORE this wait, so the softmax warps can finish tile g and tile g+1 (two arrivals) before this
      // warp has observed the first: on one barrier the awaited parity would flip back and the wait never return
      // (seen as a rare hang, adjacent CTAs at once).  One barrier per tile parity: tile g+2 needs S(g+2), which is
      // only issued after this wait.
      wait1(&p_full[g & 1], (g >> 1) & 1);
      if (lane == 0) ATT_TRACE(g, 2);
      wait1(&v_full[cp.st], cp.ph);
      if (lane == 0) ATT_TRACE(g, 3);
      if (cp.j == 0 && cp.it > 0) wait1(o_free, (cp.it - 1) & 1);  // the previous item's epilogue has read O
      tc_fence_after();
      if (elect_one()) {
        const uint64_t vd = v_desc0 + uint64_t(cp.st * stage_step);
        const uint32_t p_addr = tmem_S + (g & 1) * 128;
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; ++k)  // 16 keys = 8 packed TMEM columns = 2048 smem bytes per step
          if (!(DBG & 128))
            umma_bf16_ts(tmem_O, p_addr + k * 8, vd + uint64_t(k * (2048 >> 4)), idesc_pv, (cp.j | k) != 0);
        umma_commit(&v_empty[cp.st]);
        umma_commit(&pv_done[g & 1]);
      }
      __syncwarp();
      advance(cp);
      if (lane == 0) ATT_TRACE(g, 9);
    }
  } else {
    // ===================== softmax / correction / epilogue (2 x 128 threads, thread == query row) =====================
    const int row = threadIdx.x & 127;  // query row within the tile == TMEM lane
    const int wg = threadIdx.x >> 7;    // 0: keys [0,64) of every tile, 1: keys [64,128)
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const float LOG2E = 1.4426950408889634f;
    // O columns this half rescales / writes out (8-column groups split between the two warpgroups)
    const int og = (p.d + 7) / 8;
    const int oc0 = wg == 0 ? 0 : ((og + 1) / 2) * 8, oc1 = wg == 0 ? ((og + 1) / 2) * 8 : p.d;
    int g = 0;
    for (int it = 0; it < my_items; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int q0 = (item % n_qt) * ATT_BM, h = (item / n_qt) % p.H, b = item / (n_qt * p.H);
      float m_used = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_kv; ++j, ++g) {
        const int buf = g & 1;
        const bool need_bias = p.bias != nullptr || (j + 1) * ATT_BN > p.Sk;  // CTA-uniform
        if (need_bias) {  // additive key bias (pre-multiplied by log2 e) and the key-padding mask of a ragged last tile
          if (wg == 0) {
            const int kv = j * ATT_BN + row;
            float bv = 0.f;
            if (kv >= p.Sk) bv = -INFINITY;
            else if (p.bias) bv = p.bias[(long)b * p.Sk + kv] * LOG2E;
            sBias[buf * 128 + row] = bv;
          }
          asm volatile("bar.sync 2, 256;" ::: "memory");
        }
        mbar_wait(&s_full[buf], (g >> 1) & 1);
        if (threadIdx.x == 0) ATT_TRACE(g, 4);
        tc_fence_after();
        const uint32_t s_addr = tmem_S + buf * 128 + lane_off;
        uint32_t sr[64];
        if constexpr (DBG & 8) {
#pragma unroll
          for (int i = 0; i < 64; ++i) sr[i] = __float_as_uint((float)(i + row) * 0.001f);
        } else {
          tmem_ld32(s_addr + wg * 64, sr);
          tmem_ld32(s_addr + wg * 64 + 32, sr + 32);
          tmem_ld_wait();
        }
        if (threadIdx.x == 0) ATT_TRACE(g, 10);
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
        if (need_bias) {
          const float* bias_s = sBias + buf * 128 + wg * 64;
#pragma unroll
          for (int i = 0; i < 64; i += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(bias_s + i);
            const float v0 = fmaf(__uint_as_float(sr[i]), p.scale_log2, bb.x), v1 = fmaf(__uint_as_float(sr[i + 1]), p.scale_log2, bb.y);
            const float v2 = fmaf(__uint_as_float(sr[i + 2]), p.scale_log2, bb.z), v3 = fmaf(__uint_as_float(sr[i + 3]), p.scale_log2, bb.w);
            sr[i] = __float_as_uint(v0); sr[i + 1] = __float_as_uint(v1); sr[i + 2] = __float_as_uint(v2); sr[i + 3] = __float_as_uint(v3);
            mx0 = fmaxf(mx0, v0); mx1 = fmaxf(mx1, v1); mx2 = fmaxf(mx2, v2); mx3 = fmaxf(mx3, v3);
          }
        } else if constexpr (DBG & 16) {
          mx0 = __uint_as_float(sr[0]);
        } else {
#pragma unroll
          for (int i = 0; i < 64; i += 4) {
            mx0 = fmaxf(mx0, __uint_as_float(sr[i])); mx1 = fmaxf(mx1, __uint_as_float(sr[i + 1]));
            mx2 = fmaxf(mx2, __uint_as_float(sr[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(sr[i + 3]));
          }
        }
        float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        if (!need_bias) mx *= p.scale_log2;  // scale > 0
        // exchange the half-row maxima (double-buffered by tile parity: one barrier per tile is enough)
        if constexpr (!(DBG & 2)) {
          float* mx_s = sMax + buf * 256;
          mx_s[wg * 128 + row] = mx;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          mx = fmaxf(mx, mx_s[(wg ^ 1) * 128 + row]);
        }
        // Lazy rescale: O and l keep the stale maximum until it is exceeded by more than 2^8 (exact after the final
        // division; P <= 256 stays well inside bf16 / fp32 range).  Warp-uniform (tcgen05.ld/st are collective) and
        // identical in both warpgroups, which see the same rows and the same exchanged maxima.
        if (threadIdx.x == 0) ATT_TRACE(g, 11);
        const float m_new = fmaxf(m_used, mx);
        if (j == 0) {
          m_used = m_new;
        } else if (__any_sync(0xffffffffu, m_new - m_used > 8.f)) {
          // O must not be rescaled under an in-flight P V.  The softmax warps wait for P V only here and at the end of an
          // item, not every tile, so a single barrier would alias: S(g) ready only implies P V(g-2) retired, and with
          // P V(g-2).. two phases behind a parity wait passes at once (seen as stale O when the V tiles arrive late).
          // One barrier per tile parity: its awaited phase is either current or complete, never two behind.
          mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);
          tc_fence_after();
          const float alpha = ex2_fast(m_used - m_new);
          l_run *= alpha;
          m_used = m_new;
          for (int c = oc0; c < oc1; c += 8) {
            uint32_t r[8];
            tmem_ld8(tmem_O + lane_off + c, r);
            tmem_ld_wait();
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]) * alpha;
            tmem_st8w(tmem_O + lane_off + c, v);
          }
          tmem_st_wait();
        }
        // P is packed in place: sr[i] <- bf16x2(p[2i], p[2i+1]) (the write index never passes the read index)
        float l0 = 0.f, l1 = 0.f;
        if (need_bias) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float p0 = ex2_fast(__uint_as_float(sr[2 * i]) - m_used);
            const float p1 = ex2_fast(__uint_as_float(sr[2 * i + 1]) - m_used);
            l0 += p0;
            l1 += p1;
            sr[i] = pack_bf16(p0, p1);
          }
        } else {
          const float nm = -m_used;
          auto ex = [](float x) { if constexpr (DBG & 1) return x; else return ex2_fast(x); };
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float p0 = ex(fmaf(__uint_as_float(sr[2 * i]), p.scale_log2, nm));
            const float p1 = ex(fmaf(__uint_as_float(sr[2 * i + 1]), p.scale_log2, nm));
            l0 += p0;
            l1 += p1;
            sr[i] = pack_bf16(p0, p1);
          }
        }
        l_run += l0 + l1;
        if (threadIdx.x == 0) ATT_TRACE(g, 12);
        if constexpr (!(DBG & 4)) {
          tmem_st32(s_addr + wg * 32, sr);
          tmem_st_wait();
          if (threadIdx.x == 0) ATT_TRACE(g, 13);
        } else {
          if (l_run == 123.456f) tmem_st32(s_addr + wg * 32, sr);  // keeps sr live
        }
        tc_fence_before();
        if (threadIdx.x == 0) ATT_TRACE(g, 5);
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[buf]);
      }
      // combine the two halves' row sums (sL is only touched here: the barrier pair orders it against the last item)
      sL[wg * 128 + row] = l_run;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      l_run += sL[(wg ^ 1) * 128 + row];
      mbar_wait(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);  // the item's last P V (the tensor pipe retires in order)
      tc_fence_after();
      const int q = q0 + row;
      const float inv_l = 1.f / l_run;
      if (q < p.Sq && wg == 0) {
        if (p.lse) p.lse[((long)b * p.H + h) * p.Sq + q] = (m_used + log2f(l_run)) * 0.6931471805599453f;
      }
      bf16* orow = p.O + (long)b * p.o_stride_b + (long)q * p.o_stride_s + (long)h * p.o_stride_h;
      uint4 w[8];  // at most 8 eight-column groups per half (d <= 128); fully unrolled so it stays in registers
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = oc0 + 8 * i;
        if (c < oc1) {  // warpgroup-uniform
          uint32_t r[8];
          tmem_ld8(tmem_O + lane_off + c, r);
          tmem_ld_wait();
          w[i].x = pack_bf16(__uint_as_float(r[0]) * inv_l, __uint_as_float(r[1]) * inv_l);
          w[i].y = pack_bf16(__uint_as_float(r[2]) * inv_l, __uint_as_float(r[3]) * inv_l);
          w[i].z = pack_bf16(__uint_as_float(r[4]) * inv_l, __uint_as_float(r[5]) * inv_l);
          w[i].w = pack_bf16(__uint_as_float(r[6]) * inv_l, __uint_as_float(r[7]) * inv_l);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);  // O is in registers: the next item's first P V may overwrite it
      if (q < p.Sq) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = oc0 + 8 * i;
          if (c < oc1) *reinterpret_cast<uint4*>(orow + c) = w[i];
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // sL reads done before the next item's write
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// 4-D tensor map over a [B, S, H, d] view (d contiguous) of a token matrix; box = 64 x 128 x 1 x 1.
int make_head_map(CUtensorMap* map, const void* ptr, int B, int S, int H, int d, long stride_b, long stride_s,
                  long stride_h) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)S, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)stride_s * 2, (uint64_t)stride_h * 2, (uint64_t)stride_b * 2};
  uint32_t box[4] = {64, 128, 1, 1};
  return make_tma_bf16(map, ptr, 4, dims, strides, box, TMA_SWZ_128);
}

int launch_attn_fwd2(const TdmAttn* a, void* stream);  // attention_fwd2.cu: two query tiles per CTA, head_dim <= 96

}  // namespace tdm

using namespace tdm;

extern "C" int tdm_attn_fwd(const TdmAttn* a, void* stream) {
  TDM_REQUIRE(a && a->Q && a->K && a->V && a->O, "tdm_attn_fwd: null input");
  TDM_REQUIRE(a->d % 8 == 0 && a->d >= 16 && a->d <= 128, "tdm_attn_fwd: head_dim=%d must be a multiple of 8 in [16,128]", a->d);
  TDM_REQUIRE(a->B > 0 && a->H > 0 && a->Sq > 0 && a->Sk > 0, "tdm_attn_fwd: bad shape");
  TDM_REQUIRE(a->q_stride_s % 8 == 0 && a->q_stride_h % 8 == 0 && a->q_stride_b % 8 == 0 &&
                  a->k_stride_s % 8 == 0 && a->k_stride_h % 8 == 0 && a->k_stride_b % 8 == 0 &&
                  a->v_stride_s % 8 == 0 && a->v_stride_h % 8 == 0 && a->v_stride_b % 8 == 0 &&
                  a->o_stride_s % 8 == 0 && a->o_stride_h % 8 == 0 && a->o_stride_b % 8 == 0,
              "tdm_attn_fwd: strides must be multiples of 8 elements (16 bytes)");
  // head_dim <= 96: the two-stream kernel (TDM_B200_ATTN_V1=1 keeps this file's one-tile kernel for A/B timing)
  const bool force_v1 = getenv("TDM_B200_ATTN_V1") && atoi(getenv("TDM_B200_ATTN_V1")) != 0;
  if (a->d <= 96 && !force_v1 && !getenv("TDM_B200_ATTN_DBG")) return launch_attn_fwd2(a, stream);
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_head_map(&tmQ, a->Q, a->B, a->Sq, a->H, a->d, a->q_stride_b, a->q_stride_s, a->q_stride_h))) return rc;
  if ((rc = make_head_map(&tmK, a->K, a->B, a->Sk, a->H, a->d, a->k_stride_b, a->k_stride_s, a->k_stride_h))) return rc;
  if ((rc = make_head_map(&tmV, a->V, a->B, a->Sk, a->H, a->d, a->v_stride_b, a->v_stride_s, a->v_stride_h))) return rc;
  AttnFwdParams p;
  p.B = a->B; p.H = a->H; p.Sq = a->Sq; p.Sk = a->Sk; p.d = a->d;
  p.nd = (a->d + 63) / 64;
  p.dk16 = (a->d + 15) / 16;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.bias = a->bias;
  p.O = (bf16*)a->O; p.o_stride_b = a->o_stride_b; p.o_stride_s = a->o_stride_s; p.o_stride_h = a->o_stride_h;
  p.lse = a->lse;
  const int extra = 1024 + 6144 + 512 + 512;  // alignment slack + bias / max / sum hand-over arrays + barriers
  p.kv_stages = (8 * p.nd * CHUNK_BYTES + extra <= 232448) ? 3 : 2;  // head_dim <= 64: 16 KB tiles, three stages fit
  const int smem = (2 + 2 * p.kv_stages) * p.nd * CHUNK_BYTES + extra;
  const int dbg = getenv("TDM_B200_ATTN_DBG") ? atoi(getenv("TDM_B200_ATTN_DBG")) : 0;
  auto kern = attn_fwd_kernel<0>;
  switch (dbg) {
    case 1: kern = attn_fwd_kernel<1>; break;
    case 2: kern = attn_fwd_kernel<2>; break;
    case 4: kern = attn_fwd_kernel<4>; break;
    case 8: kern = attn_fwd_kernel<8>; break;
    case 16: kern = attn_fwd_kernel<16>; break;
    case 31: kern = attn_fwd_kernel<31>; break;
    case 128: kern = attn_fwd_kernel<128>; break;
    case 256: kern = attn_fwd_kernel<256>; break;
    case 415: kern = attn_fwd_kernel<415>; break;
    case 512: kern = attn_fwd_kernel<512>; break;
    case 927: kern = attn_fwd_kernel<927>; break;
    default: break;
  }
  int threads = FWD_THREADS;
  p.trace = nullptr;
  const char* trace_path = getenv("TDM_B200_ATTN_TRACE");  // debugging aid: per-tile clock64 stamps of CTA 0
  static long long* trace_buf = nullptr;
  if (trace_path && *trace_path) {
    if (!trace_buf) TDM_CHECK_CUDA(cudaMalloc(&trace_buf, TRACE_TILES * 16 * sizeof(long long)));
    TDM_CHECK_CUDA(cudaMemsetAsync(trace_buf, 0, TRACE_TILES * 16 * sizeof(long long), (cudaStream_t)stream));
    p.trace = trace_buf;
    kern = attn_fwd_kernel<64>;
  }
  {
    static const void* seen[16];
    static int seen_smem[16];
    int slot = 0;
    while (slot < 15 && seen[slot] && seen[slot] != (const void*)kern) ++slot;
    if (seen[slot] != (const void*)kern || smem > seen_smem[slot]) {
      TDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      seen[slot] = (const void*)kern;
      seen_smem[slot] = smem;
    }
  }
  const long items = (long)((a->Sq + ATT_BM - 1) / ATT_BM) * a->H * a->B;
  const int grid = (int)(items < num_sms() ? items : num_sms());
  ::tdm::count_launch();
  static const bool pdl_attn = !(getenv("TDM_B200_PDL_ATTN") && atoi(getenv("TDM_B200_PDL_ATTN")) == 0);
  TDM_CHECK_CUDA(launch_pdl_if(pdl_attn, kern, dim3(grid), dim3(threads), (size_t)smem, (cudaStream_t)stream, tmQ, tmK, tmV, p));
  TDM_CHECK_CUDA(cudaGetLastError());
  if (p.trace) {
    static long long host[TRACE_TILES * 16];
    TDM_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    TDM_CHECK_CUDA(cudaMemcpy(host, p.trace, sizeof(host), cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(trace_path, "a")) {
      fprintf(f, "# Sq=%d Sk=%d d=%d events: 0 qk_wait 1 qk_issue 2 p_full 3 pv_issue 4 s_full_seen 5 p_arrive 6 k_load 7 v_load\n", a->Sq, a->Sk, a->d);
      for (int g = 0; g < TRACE_TILES; ++g) {
        if (!host[g * 16 + 1]) break;
        fprintf(f, "%d", g);
        for (int e = 0; e < 14; ++e) fprintf(f, " %lld", host[g * 16 + e] ? host[g * 16 + e] - host[1] : -1);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  return TDM_OK;
}
