// Flash attention forward, two query tiles per CTA ("ping-pong"), head_dim <= 96 (PixArt: 72).  Same contract as
// attention_fwd.cu (xformers.memory_efficient_attention of the reference, main.py:177-178, 372); tdm_attn_fwd picks
// this kernel whenever the head fits.
//
// Why a second kernel.  attn_fwd_kernel keeps ONE 128-query tile per CTA in flight: its loop is a ring of hand-offs
// (S ready -> softmax -> P ready -> P V -> S ready ...) whose latencies add up to ~3000 clk per 128 x 128 tile although
// the tensor pipe needs 640 and the exponentials ~1300 (profiles/r02_attn_*).  Here one CTA owns
// 256 query rows = two independent streams s = 0, 1 that share every K / V tile.  Each stream has its own softmax
// warpgroup (thread == query row, all 128 keys of a tile: no exchange of partial maxima, no CTA barrier), its own S / P
// buffer and its own O accumulator in tensor memory; the single MMA warp alternates
//     P V_0(g), Q K^T_0(g+1) | P V_1(g), Q K^T_1(g+1) | ...
// so that the tensor work of one stream runs under the softmax of the other, and each K / V tile is fetched once per 256
// query rows.  tcgen05.mma retires in issue order, hence (i) Q K^T_s(g+1) may overwrite the S buffer that still holds
// P_s(g) as soon as it is ISSUED behind P V_s(g), and (ii) "S_s(g+1) ready" implies "P V_s(g) retired": the softmax
// warps never wait for a P V barrier inside an item (O is quiescent whenever they hold S), only once per item before
// the epilogue.
// Shared memory: rows of 64 bytes (32 head-dim columns) with the 64-byte swizzle: head_dim 72 = three chunks = 24 KB per
// 128-row tile instead of 32 KB with 128-byte rows, which is what lets 2 x 2 query tiles (double-buffered across items)
// and a two-stage K / V ring fit (192 KB).  TMA zero-fills columns >= d and rows >= S.
// TMEM map: S_0 [0,128) S_1 [128,256) O_0 [256,256+16*ceil(d/16)) O_1 [384, ...); P_s (bf16 pairs) overwrites columns
// [0,64) of S_s and is the A operand of P V straight from tensor memory.
//
// Measured alternatives that were NOT kept (B200, self-attention 8 x 16 x 1024 x 72; this form: 74-77 us, one-tile kernel 90):
//   * whole 128-key row resident in registers (setmaxnreg 232 / 40): ptxas spills in every role, 117 us;
//   * half a row resident, the other half re-read under the first half's exponentials (two tcgen05.ld rounds instead of
//     four): 86-90 us -- register pressure costs more than the round trips save;
//   * single pass with the stale maximum as the reference and a two-pass fallback (saves the maximum pass): 78.3 us;
//   * one exponential in four as a degree-3 polynomial on the FMA pipe: 82.5 us (ncu: XU pipe 44 % busy, issue slots 36 %:
//     neither is the limit -- 19 % of all warp samples sit in the softmax warps' wait for S, the rest is spread over
//     long-scoreboard / fixed-latency stalls of ~700 instructions per warp and tile);
//   * two warpgroups per stream, thread == half a row, maxima exchanged through shared memory (16 softmax warps, 96
//     registers, or 104 / 64 with setmaxnreg): 122-128 us -- 3500 clk for a thread's 64 exponentials in the trace;
//   * stream 1 started 1000-2600 clk behind stream 0 at every item (so that one stream's exponentials run under the
//     other's hand-offs instead of both at once): 75.9-76.5 us against 76.6 -- the phases of the two streams are not what
//     limits the cycle;
//   * the item's epilogue moved between the two passes of the next item's first tile (hides the wait for the last P V)
//     with all of O in one tcgen05.ld round: 86.8 us -- the inlined epilogue's 96 registers spill the hot loop;
//   * both passes as four 32-column chunks with the next chunk's tcgen05.ld in flight under the current one's math (no
//     spills, 64 live registers as here, but four tcgen05.wait::ld per pass instead of two): 81.6 us;
//   * with setmaxnreg 200 / 104 (below): the whole row resident from the maximum pass to the exponentials (one tcgen05.ld
//     round per tile, no spills): 73.5 us, the same as two rounds -- with S in registers the exponential phase still takes
//     ~1950 clk per tile (128 MUFU.EX2 per thread, two softmax warps per sub-partition: the XU pipe's 2 x 1024 clk).
//   (ex2.approx.ftz.bf16x2, which would halve the XU work, is not a packed instruction on sm_100a: ptxas emits two
//   MUFU.EX2.BF16 and a PRMT.)
// What did pay: the O tile leaving through shared memory + ONE TMA store (thread-per-row 16-byte global stores clogged the
// load/store queue that mbarrier waits and tcgen05.ld share: 83.5 -> 75.5 us), a dedicated store warp (Q buffer released
// ~600 clk after the epilogue instead of a tile later: cross-attention 27 -> 25 us), register headroom for the softmax
// warpgroups (setmaxnreg 200 for them, 104 for the load / issue / store warps, placed at the top of each role's branch --
// in a common prologue ptxas spilled every role: 76.3 -> 74.0 us, 4096 tokens 266 -> 254 us).
#include <stdlib.h>
#include "common.cuh"

namespace tdm {

constexpr int F2_THREADS = 384;    // 2 x 4 softmax warps (200 registers each via setmaxnreg) + Q/K TMA warp + MMA warp + V TMA warp + O store warp (104)
constexpr int F2_CH = 128 * 64;    // one [128 rows x 32 bf16] chunk, 64-byte swizzle
constexpr int F2_MAX_KS = 4;

struct AttnFwd2Params {
  int B, H, Sq, Sk, d;
  int nc;         // ceil(d / 32) column chunks
  int dk16;       // ceil(d / 16) MMA k-steps of Q K^T; P V produces 16 * dk16 columns
  int kv_stages;
  float scale_log2;
  const float* bias;
  bf16* O;
  long o_stride_b, o_stride_s, o_stride_h;
  float* lse;
  long long* trace;  // DBG & 64: clock64 stamps of CTA 0
};
constexpr int F2_TRACE_TILES = 64;
#define F2_TRACE(g_, e_)                                                                         \
  do {                                                                                           \
    if constexpr (DBG & 64) {                                                                    \
      if (blockIdx.x == 0 && (g_) < F2_TRACE_TILES) p.trace[(g_) * 16 + (e_)] = clock64();        \
    }                                                                                            \
  } while (0)

template <int DBG>
__global__ void __launch_bounds__(F2_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnFwd2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tile_bytes = p.nc * F2_CH;
  const int KS = p.kv_stages;
  uint8_t* sQ = smem;                      // [stream][item parity]
  uint8_t* sK = sQ + 4 * tile_bytes;       // KS stages
  uint8_t* sV = sK + KS * tile_bytes;      // KS stages
  float* sBias = reinterpret_cast<float*>(sV + KS * tile_bytes);  // [stream][tile parity][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 512);
  uint64_t* q_full = bars;         // [stream * 2 + item parity]
  uint64_t* q_empty = bars + 4;    // [stream * 2 + item parity]
  uint64_t* s_full = bars + 8;     // [stream]
  uint64_t* p_full = bars + 10;    // [stream]
  uint64_t* o_full = bars + 12;    // [stream]  the item's last P V retired
  uint64_t* k_full = bars + 14;    // [KS]
  uint64_t* k_empty = bars + 18;
  uint64_t* v_full = bars + 22;
  uint64_t* v_empty = bars + 26;
  // [stream * 2 + item parity]  the item's O tile is in shared memory (the stream's Q buffer).  One barrier per item parity:
  // a stream needs nothing from the store warp to finish item it+1 (its Q arrived long ago), so on a single barrier a store
  // warp that is more than one item late would see the awaited parity flip back and wait forever; item it+2 cannot finish
  // before the store warp has released the Q buffer of item it.
  uint64_t* o_staged = bars + 30;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 34);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_kv = (p.Sk + 127) / 128;
  const int n_qp = (p.Sq + 255) / 256;
  const int n_items = n_qp * p.H * p.B;
  const int my_items = ((int)blockIdx.x < n_items) ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int total_tiles = my_items * n_kv;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 2);  // the stream's last Q K^T of the item retired + its O tile (staged in the Q buffer) stored
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);  // one elected arrival per softmax warp of the stream
      mbar_init(&o_full[i], 1);
      mbar_init(&o_staged[2 * i], 4);
      mbar_init(&o_staged[2 * i + 1], 4);
    }
    for (int i = 0; i < F2_MAX_KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;        // stream s: + 128 s
  const uint32_t tmem_O = tmem_base + 256;  // stream s: + 128 s
  pdl_sync();
  if (warp >= 8) {
  setmaxnreg_dec<104>();
  if (warp == 8) {
    // ===================== TMA producer: Q (both streams) + K =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      int st = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_items; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int q0 = (item % n_qp) * 256, h = (item / n_qp) % p.H, b = item / (n_qp * p.H);
        for (int s = 0; s < 2; ++s) {
          const int qb = s * 2 + (it & 1);
          mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
          if (s == 0) F2_TRACE(it * n_kv, 15);
          mbar_arrive_expect_tx(&q_full[qb], tile_bytes);
          for (int c = 0; c < p.nc; ++c) tma_load_4d(sQ + qb * tile_bytes + c * F2_CH, &tmQ, &q_full[qb], c * 32, q0 + s * 128, h, b);
        }
        for (int j = 0; j < n_kv; ++j) {
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[st], tile_bytes);
          for (int c = 0; c < p.nc; ++c) tma_load_4d(sK + st * tile_bytes + c * F2_CH, &tmK, &k_full[st], c * 32, j * 128, h, b);
          if (++st == KS) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 10) {
    // ===================== TMA producer: V =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmV);
      int st = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_items; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int h = (item / n_qp) % p.H, b = item / (n_qp * p.H);
        for (int j = 0; j < n_kv; ++j) {
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_arrive_expect_tx(&v_full[st], tile_bytes);
          for (int c = 0; c < p.nc; ++c) tma_load_4d(sV + st * tile_bytes + c * F2_CH, &tmV, &v_full[st], c * 32, j * 128, h, b);
          if (++st == KS) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 11) {
    // ===================== O store: one TMA store per (stream, item), then the Q buffer it was staged in is released =====
    if (lane == 0) {
      tma_prefetch_desc(&tmO);
      for (int it = 0; it < my_items; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int q0 = (item % n_qp) * 256, h = (item / n_qp) % p.H, b = item / (n_qp * p.H);
        for (int s = 0; s < 2; ++s) {
          const int qb = s * 2 + (it & 1);
          mbar_wait_relaxed(&o_staged[s * 2 + (it & 1)], (it >> 1) & 1);
          if constexpr (!(DBG & 1)) tma_store_4d(&tmO, sQ + qb * tile_bytes, 0, q0 + s * 128, h, b);  // rows >= Sq are clipped
          tma_store_commit();
          tma_store_wait_read();
          mbar_arrive(&q_empty[qb]);
        }
      }
      tma_store_wait_all();  // shared memory must outlive the last store's reads
    }
  } else if (warp == 9) {
    // ===================== MMA issuer (warp-convergent, tcgen05 under elect.sync) =====================
    const uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(128, p.dk16 * 16, 0, 1);  // A = P from TMEM; B = V MN-major (d contiguous)
    const uint64_t q_desc0 = make_sdesc_sw64(smem_u32(sQ), 16, 512);    // K-major: 8-row groups 512 bytes apart
    const uint64_t k_desc0 = make_sdesc_sw64(smem_u32(sK), 16, 512);
    const uint64_t v_desc0 = make_sdesc_sw64(smem_u32(sV), F2_CH, 512); // MN-major: 32-column chunks F2_CH apart
    const uint32_t tile_step = uint32_t(tile_bytes) >> 4;
    const int dk16 = p.dk16;
    // Running coordinates of the tile whose Q K^T (q_*) / P V (p_*) is issued next -- both streams, then advanced: plain
    // scalars (a struct captured by reference lived in local memory: LDL / STL on the issue path).
    int q_it = 0, q_j = 0, q_st = 0, gq = 0;
    uint32_t q_ph = 0;
    int p_j = 0, p_st = 0;
    uint32_t p_ph = 0;
#define F2_ISSUE_QK(S_)                                                                                      \
    do {                                                                                                     \
      if ((S_) == 0) mbar_wait(&k_full[q_st], q_ph);                                                         \
      const int qb_ = (S_) * 2 + (q_it & 1);                                                                 \
      if (q_j == 0) mbar_wait(&q_full[qb_], (q_it >> 1) & 1);                                                \
      tc_fence_after();                                                                                      \
      if (lane == 0) F2_TRACE(gq, (S_));                                                                     \
      if (elect_one()) {                                                                                     \
        const uint64_t qd = q_desc0 + uint64_t(qb_ * tile_step);                                             \
        const uint64_t kd = k_desc0 + uint64_t(q_st * tile_step);                                            \
        const uint32_t sd = tmem_S + (S_) * 128;                                                             \
        _Pragma("unroll") for (int k = 0; k < 6; ++k) {                                                      \
          if (k < dk16) { /* 16 head-dim columns per step: two steps per 32-column chunk */                  \
            const uint32_t off = ((k >> 1) * F2_CH + (k & 1) * 32) >> 4;                                     \
            umma_bf16(sd, qd + off, kd + off, idesc_qk, k != 0);                                             \
          }                                                                                                  \
        }                                                                                                    \
        umma_commit(&s_full[(S_)]);                                                                          \
        if ((S_) == 1) umma_commit(&k_empty[q_st]);                                                          \
        if (q_j == n_kv - 1) umma_commit(&q_empty[qb_]); /* last read of this stream's Q tile */             \
      }                                                                                                      \
      __syncwarp();                                                                                          \
      if ((S_) == 1) {                                                                                       \
        if (++q_j == n_kv) { q_j = 0; ++q_it; }                                                              \
        if (++q_st == KS) { q_st = 0; q_ph ^= 1; }                                                           \
        ++gq;                                                                                                \
      }                                                                                                      \
    } while (0)
    if (total_tiles > 0) {
      F2_ISSUE_QK(0);
      F2_ISSUE_QK(1);
    }
    for (int g = 0; g < total_tiles; ++g) {
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        mbar_wait(&p_full[s], g & 1);
        if (s == 0) mbar_wait(&v_full[p_st], p_ph);
        tc_fence_after();
        if (lane == 0) F2_TRACE(g, 2 + s);
        if (elect_one()) {
          const uint64_t vd = v_desc0 + uint64_t(p_st * tile_step);
          const uint32_t p_addr = tmem_S + s * 128;
#pragma unroll
          for (int k = 0; k < 8; ++k)  // 16 keys = 8 packed TMEM columns = 16 smem rows of 64 bytes per step
            umma_bf16_ts(tmem_O + s * 128, p_addr + k * 8, vd + uint64_t(k * (1024 >> 4)), idesc_pv, (p_j | k) != 0);
          if (s == 1) umma_commit(&v_empty[p_st]);
          if (p_j == n_kv - 1) umma_commit(&o_full[s]);
        }
        __syncwarp();
        if (g + 1 < total_tiles) F2_ISSUE_QK(s);
      }
      if (++p_j == n_kv) p_j = 0;
      if (++p_st == KS) { p_st = 0; p_ph ^= 1; }
    }
#undef F2_ISSUE_QK
  }
  } else {
    setmaxnreg_inc<200>();
    // ===================== softmax / correction / epilogue: warpgroup s owns stream s, thread == query row ==========
    const int s = warp >> 2;
    const int row = threadIdx.x & 127;
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const uint32_t s_addr = tmem_S + s * 128 + lane_off;
    const uint32_t o_addr = tmem_O + s * 128 + lane_off;
    float* bias_s = sBias + s * 256;
    const float LOG2E = 1.4426950408889634f;
    const float sc = p.scale_log2;
    int g = 0;
    // bias value of key (tile * 128 + row) of batch b_: loaded one tile AHEAD into a register (the global-memory round
    // trip sat between the epilogue of an item and its successor's first tile: ~1000 clk of a ~9800 clk cross-attention item)
    auto bias_of = [&](int b_, int j_) {
      const int kv = j_ * 128 + row;
      float bv = 0.f;
      if (kv >= p.Sk) bv = -INFINITY;
      else if (p.bias) bv = p.bias[(long)b_ * p.Sk + kv] * LOG2E;
      return bv;
    };
    float bias_next = 0.f;
    if (my_items > 0) bias_next = bias_of((int)blockIdx.x / (n_qp * p.H), 0);
    for (int it = 0; it < my_items; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int q0 = (item % n_qp) * 256 + s * 128, h = (item / n_qp) % p.H, b = item / (n_qp * p.H);
      float m_used = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_kv; ++j, ++g) {
        const bool need_bias = p.bias != nullptr || (j + 1) * 128 > p.Sk;  // CTA-uniform
        const uint32_t bs_addr = smem_u32(bias_s + (g & 1) * 128);
        if (need_bias) {  // additive key bias (times log2 e) and the key-padding mask of a ragged last tile
          bias_s[(g & 1) * 128 + row] = bias_next;
          if (s == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
        }
        {  // the next tile's value (this item's next key tile, or the first one of the next item)
          const bool last_j = j + 1 == n_kv;
          const int nb = last_j ? ((int)blockIdx.x + (it + 1) * (int)gridDim.x) / (n_qp * p.H) : b;
          if (!last_j || it + 1 < my_items) bias_next = bias_of(nb, last_j ? 0 : j + 1);
        }
        mbar_wait(&s_full[s], g & 1);
        tc_fence_after();
        if (row == 0) F2_TRACE(g, 4 + s);
        uint32_t a[32], c[32];
        // ---- pass 1: row maximum.  The row is read twice from tensor memory in 64-column halves instead of living in
        //      128 registers: a whole row resident (setmaxnreg 232 / 40) made ptxas spill in every role (117 us), half a
        //      row resident with the other half re-read under the first half's exponentials measured 86-90 us against
        //      74 us for this form -- the register pressure costs more than the extra tcgen05.ld round trips. ----
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
        if (!need_bias) {
          // the maximum pass holds nothing but the row: all 128 columns in ONE round trip (128 live registers: the softmax
          // warpgroups have 200 through setmaxnreg).  The same idea in pass 2 -- second half in flight under the first
          // half's exponentials -- measured 83 us against 73.5: loads in flight while the warp computes were slower in
          // every form tried.
          uint32_t a2[32], c2[32];
          tmem_ld32(s_addr, a);
          tmem_ld32(s_addr + 32, c);
          tmem_ld32(s_addr + 64, a2);
          tmem_ld32(s_addr + 96, c2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            mx0 = fmaxf(mx0, fmaxf(fmaxf(__uint_as_float(a[i]), __uint_as_float(c[i])), fmaxf(__uint_as_float(a2[i]), __uint_as_float(c2[i]))));
            mx1 = fmaxf(mx1, fmaxf(fmaxf(__uint_as_float(a[i + 1]), __uint_as_float(c[i + 1])), fmaxf(__uint_as_float(a2[i + 1]), __uint_as_float(c2[i + 1]))));
            mx2 = fmaxf(mx2, fmaxf(fmaxf(__uint_as_float(a[i + 2]), __uint_as_float(c[i + 2])), fmaxf(__uint_as_float(a2[i + 2]), __uint_as_float(c2[i + 2]))));
            mx3 = fmaxf(mx3, fmaxf(fmaxf(__uint_as_float(a[i + 3]), __uint_as_float(c[i + 3])), fmaxf(__uint_as_float(a2[i + 3]), __uint_as_float(c2[i + 3]))));
          }
        } else {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            tmem_ld32(s_addr + half * 64, a);
            tmem_ld32(s_addr + half * 64 + 32, c);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float4 ba, bc;  // volatile: ptxas otherwise hoists every bias load above the first use
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(ba.x), "=f"(ba.y), "=f"(ba.z), "=f"(ba.w) : "r"(bs_addr + 256 * half + 4 * i));
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(bc.x), "=f"(bc.y), "=f"(bc.z), "=f"(bc.w) : "r"(bs_addr + 256 * half + 128 + 4 * i));
              mx0 = fmaxf(mx0, fmaxf(fmaf(__uint_as_float(a[i]), sc, ba.x), fmaf(__uint_as_float(c[i]), sc, bc.x)));
              mx1 = fmaxf(mx1, fmaxf(fmaf(__uint_as_float(a[i + 1]), sc, ba.y), fmaf(__uint_as_float(c[i + 1]), sc, bc.y)));
              mx2 = fmaxf(mx2, fmaxf(fmaf(__uint_as_float(a[i + 2]), sc, ba.z), fmaf(__uint_as_float(c[i + 2]), sc, bc.z)));
              mx3 = fmaxf(mx3, fmaxf(fmaf(__uint_as_float(a[i + 3]), sc, ba.w), fmaf(__uint_as_float(c[i + 3]), sc, bc.w)));
            }
          }
        }
        float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        if (!need_bias) mx *= sc;  // scale > 0
        // Lazy rescale: O and l keep the stale maximum until it is exceeded by more than 2^8 (exact after the final
        // division; P <= 256 is well inside bf16 / fp32 range).  Warp-uniform: tcgen05.ld / st are warp-collective.
        // No barrier: S_s(g) ready implies P V_s(g-1) retired, and P V_s(g) is only issued after this warp's arrival.
        const float m_new = fmaxf(m_used, mx);
        if (j == 0) {
          m_used = m_new;
        } else if (__any_sync(0xffffffffu, m_new - m_used > 8.f)) {
          const float alpha = ex2_fast(m_used - m_new);
          l_run *= alpha;
          m_used = m_new;
#pragma unroll 1
          for (int cc = 0; cc < p.dk16; ++cc) {
            uint32_t r[16];
            tmem_ld16(o_addr + cc * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st16(o_addr + cc * 16, r);
          }
        }
        if (row == 0) F2_TRACE(g, 6 + s);
        // ---- pass 2: P = exp2(S * scale + bias - m), bf16 pairs written over the S columns already consumed ----
        const float nm = -m_used;
        float l0 = 0.f, l1 = 0.f;
        // DBG & 8: one exponential in four on the FMA pipe (degree-3 polynomial, relative error 1e-4 << bf16)
        auto exq = [](float x, int i) { if constexpr (DBG & 8) return (i & 1) ? ex2_poly(x) : ex2_fast(x); else return ex2_fast(x); };
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tmem_ld32(s_addr + half * 64, a);
          tmem_ld32(s_addr + half * 64 + 32, c);
          tmem_ld_wait();
          if (need_bias) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float2 ba, bc;
              asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(ba.x), "=f"(ba.y) : "r"(bs_addr + 256 * half + 8 * i));
              asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(bc.x), "=f"(bc.y) : "r"(bs_addr + 256 * half + 128 + 8 * i));
              const float p0 = ex2_fast(fmaf(__uint_as_float(a[2 * i]), sc, ba.x) + nm);
              const float p1 = ex2_fast(fmaf(__uint_as_float(a[2 * i + 1]), sc, ba.y) + nm);
              const float p2 = ex2_fast(fmaf(__uint_as_float(c[2 * i]), sc, bc.x) + nm);
              const float p3 = ex2_fast(fmaf(__uint_as_float(c[2 * i + 1]), sc, bc.y) + nm);
              l0 += p0 + p2;
              l1 += p1 + p3;
              a[i] = pack_bf16(p0, p1);
              c[i] = pack_bf16(p2, p3);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float p0 = ex2_fast(fmaf(__uint_as_float(a[2 * i]), sc, nm));
              const float p1 = exq(fmaf(__uint_as_float(a[2 * i + 1]), sc, nm), i);
              const float p2 = ex2_fast(fmaf(__uint_as_float(c[2 * i]), sc, nm));
              const float p3 = exq(fmaf(__uint_as_float(c[2 * i + 1]), sc, nm), i);
              l0 += p0 + p2;
              l1 += p1 + p3;
              a[i] = pack_bf16(p0, p1);
              c[i] = pack_bf16(p2, p3);
            }
          }
          // keys [64 half, 64 half + 32) -> P columns [32 half, 32 half + 16), the next 32 keys behind them: always inside
          // S columns this thread has already read
          tmem_st16(s_addr + half * 32, a);
          tmem_st16(s_addr + half * 32 + 16, c);
        }
        l_run += l0 + l1;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
        if (row == 0) F2_TRACE(g, 8 + s);
      }
      // ---- epilogue of the item: O / l -> bf16, log-sum-exp ----
      if (threadIdx.x == 0) F2_TRACE(g - 1, 10);
      mbar_wait(&o_full[s], it & 1);
      tc_fence_after();
      if (threadIdx.x == 0) F2_TRACE(g - 1, 11);
      const int q = q0 + row;
      const float inv_l = 1.f / l_run;
      // O_s -> bf16 rows in the stream's Q buffer of THIS item (free: o_full implies every Q K^T of the item retired; the
      // producer reloads it only after q_empty's second arrival, by the store warp) -> one TMA store per 128 x d tile.  Thread-per-row
      // 16-byte global stores (32 rows = 32 transactions per instruction) clogged the SM's load/store queue -- which the
      // mbarrier waits and tcgen05.ld of every warp share -- for ~5000 clk per item: 83.5 -> 68.9 us with the stores
      // compiled out.  The next item's first P V_s, which overwrites O_s, is issued only after this warp's next p_full
      // arrival, behind another tcgen05.fence::before_thread_sync in program order.
      if (q < p.Sq && p.lse) p.lse[((long)b * p.H + h) * p.Sq + q] = (m_used + log2f(l_run)) * 0.6931471805599453f;
      uint8_t* srow = sQ + (s * 2 + (it & 1)) * tile_bytes + row * (p.d * 2);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {  // d <= 96: two rounds of three 16-column loads
        uint32_t r[3][16];
#pragma unroll
        for (int i = 0; i < 3; ++i)
          if (3 * hh + i < p.dk16) tmem_ld16(o_addr + 16 * (3 * hh + i), r[i]);
        tmem_ld_wait();
        if (threadIdx.x == 0) F2_TRACE(g - 1, 12);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          if (8 * (6 * hh + i) < p.d) {
#define F2_E(k_) __uint_as_float(r[i >> 1][(i & 1) * 8 + (k_)])
            uint4 w;
            w.x = pack_bf16(F2_E(0) * inv_l, F2_E(1) * inv_l);
            w.y = pack_bf16(F2_E(2) * inv_l, F2_E(3) * inv_l);
            w.z = pack_bf16(F2_E(4) * inv_l, F2_E(5) * inv_l);
            w.w = pack_bf16(F2_E(6) * inv_l, F2_E(7) * inv_l);
#undef F2_E
            *reinterpret_cast<uint4*>(srow + 16 * (6 * hh + i)) = w;
          }
        }
      }
      if (threadIdx.x == 0) F2_TRACE(g - 1, 13);
      fence_async_smem();  // generic-proxy writes -> visible to the TMA store the store warp issues
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_staged[s * 2 + (it & 1)]);
      if (threadIdx.x == 0) F2_TRACE(g - 1, 14);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// 4-D tensor map over a [B, S, H, d] view (d contiguous); box = 32 x 128 x 1 x 1, 64-byte swizzle.
int make_head_map64(CUtensorMap* map, const void* ptr, int B, int S, int H, int d, long stride_b, long stride_s,
                           long stride_h) {
  uint64_t dims[4] = {(uint64_t)d, (uint64_t)S, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)stride_s * 2, (uint64_t)stride_h * 2, (uint64_t)stride_b * 2};
  uint32_t box[4] = {32, 128, 1, 1};
  return make_tma_bf16(map, ptr, 4, dims, strides, box, TMA_SWZ_64);
}

// Called by tdm_attn_fwd (attention_fwd.cu) after argument validation, for head_dim <= 96.
int launch_attn_fwd2(const TdmAttn* a, void* stream) {
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_head_map64(&tmQ, a->Q, a->B, a->Sq, a->H, a->d, a->q_stride_b, a->q_stride_s, a->q_stride_h))) return rc;
  if ((rc = make_head_map64(&tmK, a->K, a->B, a->Sk, a->H, a->d, a->k_stride_b, a->k_stride_s, a->k_stride_h))) return rc;
  if ((rc = make_head_map64(&tmV, a->V, a->B, a->Sk, a->H, a->d, a->v_stride_b, a->v_stride_s, a->v_stride_h))) return rc;
  CUtensorMap tmO;
  {
    uint64_t dims[4] = {(uint64_t)a->d, (uint64_t)a->Sq, (uint64_t)a->H, (uint64_t)a->B};
    uint64_t strides[3] = {(uint64_t)a->o_stride_s * 2, (uint64_t)a->o_stride_h * 2, (uint64_t)a->o_stride_b * 2};
    uint32_t box[4] = {(uint32_t)a->d, 128, 1, 1};
    if ((rc = make_tma_bf16(&tmO, a->O, 4, dims, strides, box, TMA_SWZ_NONE))) return rc;
  }
  AttnFwd2Params p;
  p.B = a->B; p.H = a->H; p.Sq = a->Sq; p.Sk = a->Sk; p.d = a->d;
  p.nc = (a->d + 31) / 32;
  p.dk16 = (a->d + 15) / 16;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.bias = a->bias;
  p.O = (bf16*)a->O; p.o_stride_b = a->o_stride_b; p.o_stride_s = a->o_stride_s; p.o_stride_h = a->o_stride_h;
  p.lse = a->lse;
  p.trace = nullptr;
  const int tile_bytes = p.nc * F2_CH;
  const int extra = 1024 + 2048 + 512;  // alignment slack + bias rows + barriers
  int ks = (232448 - extra - 4 * tile_bytes) / (2 * tile_bytes);
  p.kv_stages = ks > F2_MAX_KS ? F2_MAX_KS : ks;
  TDM_REQUIRE(p.kv_stages >= 2, "tdm_attn_fwd: head_dim=%d does not fit the two-stream kernel", a->d);
  const int smem = (4 + 2 * p.kv_stages) * tile_bytes + extra;
  auto kern = attn_fwd2_kernel<0>;
  switch (getenv("TDM_B200_ATTN2_DBG") ? atoi(getenv("TDM_B200_ATTN2_DBG")) : 0) {  // experiments
    case 1: kern = attn_fwd2_kernel<1>; break;  // no O stores (wrong results)
    case 8: kern = attn_fwd2_kernel<8>; break;  // one exponential in four as a polynomial on the FMA pipe
    default: break;
  }
  const char* trace_path = getenv("TDM_B200_ATTN_TRACE");  // debugging aid: per-tile clock64 stamps of CTA 0
  static long long* trace_buf = nullptr;
  if (trace_path && *trace_path) {
    if (!trace_buf) TDM_CHECK_CUDA(cudaMalloc(&trace_buf, F2_TRACE_TILES * 16 * sizeof(long long)));
    TDM_CHECK_CUDA(cudaMemsetAsync(trace_buf, 0, F2_TRACE_TILES * 16 * sizeof(long long), (cudaStream_t)stream));
    p.trace = trace_buf;
    kern = attn_fwd2_kernel<64>;
  }
  {
    static const void* seen[4];
    static int seen_smem[4];
    int slot = 0;
    while (slot < 3 && seen[slot] && seen[slot] != (const void*)kern) ++slot;
    if (seen[slot] != (const void*)kern || smem > seen_smem[slot]) {
      TDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      seen[slot] = (const void*)kern;
      seen_smem[slot] = smem;
    }
  }
  const long items = (long)((a->Sq + 255) / 256) * a->H * a->B;
  const int grid = (int)(items < num_sms() ? items : num_sms());
  count_launch();
  static const bool pdl_attn = !(getenv("TDM_B200_PDL_ATTN") && atoi(getenv("TDM_B200_PDL_ATTN")) == 0);
  TDM_CHECK_CUDA(launch_pdl_if(pdl_attn, kern, dim3(grid), dim3(F2_THREADS), (size_t)smem, (cudaStream_t)stream, tmQ, tmK, tmV, tmO, p));
  TDM_CHECK_CUDA(cudaGetLastError());
  if (p.trace) {
    static long long host[F2_TRACE_TILES * 16];
    TDM_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    TDM_CHECK_CUDA(cudaMemcpy(host, p.trace, sizeof(host), cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(trace_path, "a")) {
      fprintf(f, "# fwd2 Sq=%d Sk=%d d=%d events: 0/1 qk_issue s0/s1  2/3 pv_issue s0/s1  4/5 s_full seen  6/7 max done  8/9 p_arrive  10-14 epilogue s0  15 Q load issued (item)\n",
              a->Sq, a->Sk, a->d);
      long long t0 = host[0];
      for (int g = 0; g < F2_TRACE_TILES; ++g) {
        if (!host[g * 16 + 4]) break;
        fprintf(f, "%d", g);
        for (int e = 0; e < 16; ++e) fprintf(f, " %lld", host[g * 16 + e] ? host[g * 16 + e] - t0 : -1);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  return TDM_OK;
}

}  // namespace tdm
