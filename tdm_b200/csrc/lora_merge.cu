// LoRA merge for the no-grad forwards: W_eff = bf16(W32 + (alpha / r) * B A) for EVERY adapted linear of a model in ONE
// launch (r = 64).  The four trajectory forwards of an iteration (reference main.py:464-466) read the merged weights, so
// they cost what a teacher forward costs; the merge runs once per optimiser step.  Round 1 issued one small tcgen05 GEMM
// per adapted linear -- 280 launches of ~12 us, 3.5 ms per iteration, each using 9-36 of 148 SMs; the work itself is
// HBM-bound (read the fp32 master, write the bf16 copy: 3.6 GB, 0.55 ms at the measured copy bandwidth).
//
// One CTA per 128 x 128 tile of some W (device-side problem table, the arenas never move): LoRA-B rows -> SWIZZLE_128B
// K-major A operand, LoRA-A columns -> SWIZZLE_128B MN-major B operand (written by the CTA's own threads: 16 KB each,
// no tensor maps to build per problem), four tcgen05.mma (M 128, N 128, K 16) into 128 TMEM columns, thread == row epilogue
// handed over through shared memory so that global memory sees whole rows per warp.  Three CTAs share an SM (65 KB shared
// memory, 128 TMEM columns each), which is what overlaps one tile's loads with another's stores.
#include "common.cuh"

namespace tdm {

constexpr int MG_THREADS = 128;

constexpr int MG_MAX_PROBLEMS = 1024;  // tile_begin of every problem staged in shared memory (the search is on the tile path)

__device__ __forceinline__ int find_problem(const int* tile_begin, int n, int tile) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tile_begin[mid] <= tile) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(MG_THREADS, 3)
lora_merge_kernel(const TdmMergeProblem* __restrict__ probs, int n_probs, int total_tiles, float scale) {
  extern __shared__ uint8_t mg_smem_raw[];
  uint8_t* sA = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mg_smem_raw) + 1023) & ~uintptr_t(1023));  // LoRA-B tile: 128 rows x 64 bf16, K-major
  uint8_t* sB = sA + 128 * 128;                         // LoRA-A tile: two 64-column chunks of [64 k][64 n], MN-major
  float* sT = reinterpret_cast<float*>(sB + 2 * 64 * 128);  // accumulator chunk, transposed hand-over (row stride 65: no bank conflicts)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int s_begin[MG_MAX_PROBLEMS];
  const int warp = threadIdx.x >> 5, tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_slot;
  const uint32_t lane_off = uint32_t(warp * 32) << 16;
  const uint32_t idesc = make_idesc_bf16(128, 128, 0, 1);
  for (int i = tid; i < n_probs; i += MG_THREADS) s_begin[i] = probs[i].tile_begin;  // built once on the host: not PDL-ordered data
  __syncthreads();
  pdl_sync();  // the adapters come from the optimiser kernel that precedes this launch
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const TdmMergeProblem pr = probs[find_problem(s_begin, n_probs, tile)];
    const int t = tile - pr.tile_begin;
    const int tiles_n = pr.k_in >> 7;
    const int r0 = (t / tiles_n) << 7, c0 = (t % tiles_n) << 7;
    // ---- operands -> shared memory (generic proxy), swizzled the way the tensor core reads them ----
    {
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(pr.lora_b) + (long)r0 * 64);
#pragma unroll
      for (int i = 0; i < 8; ++i) {  // 128 rows x 8 sixteen-byte chunks
        const int idx = i * MG_THREADS + tid, r = idx >> 3, c = idx & 7;
        *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = src[idx];
      }
      const bf16* la = reinterpret_cast<const bf16*>(pr.lora_a);
#pragma unroll
      for (int i = 0; i < 8; ++i) {  // 64 k-rows x 16 sixteen-byte chunks (128 columns)
        const int idx = i * MG_THREADS + tid, k = idx >> 4, c = idx & 15;
        const uint4 v = *reinterpret_cast<const uint4*>(la + (long)k * pr.k_in + c0 + c * 8);
        *reinterpret_cast<uint4*>(sB + (c >> 3) * 8192 + k * 128 + (((c & 7) ^ (k & 7)) << 4)) = v;
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_acc, make_sdesc_sw128(aA + k * 32, 16, 1024), make_sdesc_sw128(aB + k * 2048, 8192, 1024), idesc, k != 0);
      umma_commit(&bar);
    }
    mbar_wait_relaxed(&bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: TMEM hands every thread one ROW of the accumulator, global memory wants a warp on one row: transpose
    //      64-column chunks through shared memory (each warp owns its 32 rows: no CTA barrier), then lane l streams columns
    //      2l, 2l+1 of row after row -- 256-byte fp32 reads, 128-byte bf16 writes per warp instruction.  (Thread == row
    //      all the way to global memory measured 3.96 ms for the whole model: 16-byte loads / 8-byte stores scattered over
    //      32 rows per instruction; the whole merge moves 3.6 GB.) ----
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t acc[64];
      tmem_ld32(tmem_acc + lane_off + cc * 64, acc);
      tmem_ld32(tmem_acc + lane_off + cc * 64 + 32, acc + 32);
      tmem_ld_wait();
      float* trow = sT + tid * 65;
#pragma unroll
      for (int j = 0; j < 64; ++j) trow[j] = __uint_as_float(acc[j]);
      __syncwarp();
      const int lane = tid & 31;
      // all 32 row loads of this warp in flight before the first use: 8 KB per warp, ~100 KB per SM (with four in flight the
      // kernel was latency-bound at 1.1 TB/s)
      float2 wv[32];
      const float* wbase = pr.w32 + (long)(r0 + warp * 32) * pr.k_in + c0 + cc * 64 + 2 * lane;
#pragma unroll
      for (int r = 0; r < 32; ++r) wv[r] = *reinterpret_cast<const float2*>(wbase + (long)r * pr.k_in);
      bf16* obase = reinterpret_cast<bf16*>(pr.w_eff) + (long)(r0 + warp * 32) * pr.k_in + c0 + cc * 64 + 2 * lane;
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const int rr = warp * 32 + r;
        const float a0 = sT[rr * 65 + 2 * lane], a1 = sT[rr * 65 + 2 * lane + 1];
        *reinterpret_cast<uint32_t*>(obase + (long)r * pr.k_in) = pack_bf16(fmaf(scale, a0, wv[r].x), fmaf(scale, a1, wv[r].y));
      }
      __syncwarp();
    }
    tc_fence_before();
    __syncthreads();  // accumulator and operand tiles are free for the next tile
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, 128);
  }
}

}  // namespace tdm

using namespace tdm;

extern "C" int tdm_lora_merge(const TdmMergeProblem* problems_dev, int n_problems, int total_tiles, float scale, void* stream) {
  TDM_REQUIRE(problems_dev && n_problems > 0 && total_tiles > 0, "tdm_lora_merge: empty problem table");
  TDM_REQUIRE(n_problems <= MG_MAX_PROBLEMS, "tdm_lora_merge: at most %d problems per launch (got %d)", MG_MAX_PROBLEMS, n_problems);
  const int cap = num_sms() * 3;
  const int grid = total_tiles < cap ? total_tiles : cap;
  const int smem = 1024 + 128 * 128 + 2 * 64 * 128 + 128 * 65 * 4;
  static bool attr = false;
  if (!attr) {
    TDM_CHECK_CUDA(cudaFuncSetAttribute(lora_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  count_launch();
  TDM_CHECK_CUDA(launch_pdl(lora_merge_kernel, dim3(grid), dim3(MG_THREADS), (size_t)smem, (cudaStream_t)stream, problems_dev, n_problems, total_tiles, scale));
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}
