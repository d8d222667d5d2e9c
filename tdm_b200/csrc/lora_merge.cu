// LoRA merge for the no-grad forwards: W_eff = bf16(W32 + (alpha / r) * B A) for EVERY adapted linear of a model in ONE
// launch (r = 64).  The four trajectory forwards of an iteration (reference main.py:464-466) read the merged weights, so
// they cost what a teacher forward costs; the merge runs once per optimiser step.  Round 1 issued one small tcgen05 GEMM
// per adapted linear -- 280 launches of ~12 us, 3.5 ms per iteration, each using 9-36 of 148 SMs; the work itself is
// HBM-bound (read the fp32 master, write the bf16 copy: 3.6 GB, 0.55 ms at the measured copy bandwidth).
//
// Persistent kernel, one CTA per SM, 128 x 128 tiles of some W from a device-side problem table (the arenas never move):
//   warp 4     : streams the fp32 master tile into a 2-stage shared-memory ring, one 512-byte bulk copy per row (no
//                tensor maps to build per problem); 128 KB per SM in flight is what keeps HBM busy while the other
//                warps are inside their per-tile latency chain
//   warps 0-3  : LoRA-B rows -> SWIZZLE_128B K-major A operand, LoRA-A columns -> SWIZZLE_128B MN-major B operand
//                (written by the threads themselves; the NEXT tile's operand rows are already in registers), four
//                tcgen05.mma (M 128, N 128, K 16) into 128 TMEM columns, then thread == row: accumulator row from TMEM,
//                master row from the ring (padded row pitch: conflict-free), bf16 row into a staging buffer, one 256-byte
//                bulk store per row.
// No thread touches global memory with a scattered 16-byte access: the first version (thread == row loads / stores)
// took 3.96 ms, a shared-memory transpose to warp-per-row accesses 3.30 ms (1.1 TB/s, latency-bound).
#include "common.cuh"

namespace tdm {

constexpr int MG_THREADS = 160;        // 4 compute warps + 1 loader warp
constexpr int MG_MAX_PROBLEMS = 1024;  // tile_begin of every problem staged in shared memory (the search is on the tile path)
constexpr int MG_WPITCH = 528;         // bytes between fp32 master rows in the ring (512 + 16: thread-per-row float4 reads hit 32 banks)
constexpr int MG_OPITCH = 272;         // bytes between bf16 output rows in the staging buffer (256 + 16)
constexpr int MG_STAGE = 128 * MG_WPITCH;
constexpr int MG_SMEM = 1024 + 2 * MG_STAGE + 128 * 128 + 2 * 64 * 128 + 128 * MG_OPITCH;

__device__ __forceinline__ int find_problem(const int* tile_begin, int n, int tile) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tile_begin[mid] <= tile) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}
__device__ __forceinline__ void bulk_load_row(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store_row(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}

__global__ void __launch_bounds__(MG_THREADS, 1)
lora_merge_kernel(const TdmMergeProblem* __restrict__ probs, int n_probs, int total_tiles, float scale) {
  extern __shared__ uint8_t mg_smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mg_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = base;                    // LoRA-B tile: 128 rows x 64 bf16, K-major, 128-byte swizzle
  uint8_t* sB = sA + 128 * 128;          // LoRA-A tile: two 64-column chunks of [64 k][64 n], MN-major
  uint8_t* sW = sB + 2 * 64 * 128;       // ring: 2 x 128 fp32 master rows of 512 bytes, pitch MG_WPITCH
  uint8_t* sO = sW + 2 * MG_STAGE;       // 128 bf16 output rows of 256 bytes, pitch MG_OPITCH
  __shared__ uint64_t mma_bar, w_full[2], w_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_begin[MG_MAX_PROBLEMS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&mma_bar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 4);  // one elected arrival per compute warp
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  for (int i = tid; i < n_probs; i += MG_THREADS) s_begin[i] = probs[i].tile_begin;  // built once on the host: not PDL-ordered data
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_slot;
  pdl_sync();  // the adapters and the master weights come from the optimiser kernel that precedes this launch

  if (warp == 4) {
    // ===================== loader: fp32 master tiles -> ring =====================
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const TdmMergeProblem pr = probs[find_problem(s_begin, n_probs, tile)];
      const int t = tile - pr.tile_begin;
      const int tiles_n = pr.k_in >> 7;
      const int r0 = (t / tiles_n) << 7, c0 = (t % tiles_n) << 7;
      const int st = it & 1;
      mbar_wait(&w_empty[st], ((it >> 1) & 1) ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(&w_full[st], 128 * 512);
      __syncwarp();
      const float* src = pr.w32 + (long)r0 * pr.k_in + c0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = k * 32 + lane;
        bulk_load_row(sW + st * MG_STAGE + r * MG_WPITCH, src + (long)r * pr.k_in, 512, &w_full[st]);
      }
    }
  } else {
    // ===================== operands, MMA, epilogue (thread == row of the tile) =====================
    const uint32_t lane_off = uint32_t(warp * 32) << 16;
    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 1);
    // operand rows of a tile: B rows r0 .. r0+127 (128 bytes each: 8 chunks per thread, chunk index idx = i * 128 + tid),
    // A rows 0 .. 63, columns c0 .. c0+127 (256 bytes each)
    uint4 ob[8], oa[8];
    auto load_operands = [&](int tile) {
      const TdmMergeProblem pr = probs[find_problem(s_begin, n_probs, tile)];
      const int t = tile - pr.tile_begin;
      const int tiles_n = pr.k_in >> 7;
      const int r0 = (t / tiles_n) << 7, c0 = (t % tiles_n) << 7;
      const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(pr.lora_b) + (long)r0 * 64);
      const bf16* la = reinterpret_cast<const bf16*>(pr.lora_a);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = i * 128 + tid;
        ob[i] = src[idx];
        oa[i] = *reinterpret_cast<const uint4*>(la + (long)(idx >> 4) * pr.k_in + c0 + (idx & 15) * 8);
      }
    };
    int it = 0;
    uint32_t mma_phase = 0;
    if ((int)blockIdx.x < total_tiles) load_operands(blockIdx.x);
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const TdmMergeProblem pr = probs[find_problem(s_begin, n_probs, tile)];
      const int t = tile - pr.tile_begin;
      const int tiles_n = pr.k_in >> 7;
      const int r0 = (t / tiles_n) << 7, c0 = (t % tiles_n) << 7;
      const int st = it & 1;
      // ---- operands -> shared memory (generic proxy), swizzled the way the tensor core reads them.  The previous
      //      tile's MMAs have retired (mma_bar was waited for), so the buffers are free. ----
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = i * 128 + tid;
        const int r = idx >> 3, c = idx & 7;        // B: 128 rows x 8 sixteen-byte chunks
        *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = ob[i];
        const int k = idx >> 4, cc = idx & 15;      // A: 64 k-rows x 16 sixteen-byte chunks (128 columns)
        *reinterpret_cast<uint4*>(sB + (cc >> 3) * 8192 + k * 128 + (((cc & 7) ^ (k & 7)) << 4)) = oa[i];
      }
      fence_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid == 0) {
        tc_fence_after();
        const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_acc, make_sdesc_sw128(aA + k * 32, 16, 1024), make_sdesc_sw128(aB + k * 2048, 8192, 1024), idesc, k != 0);
        umma_commit(&mma_bar);
      }
      if (tile + (int)gridDim.x < total_tiles) load_operands(tile + gridDim.x);  // in flight under the MMAs and the epilogue
      mbar_wait_relaxed(&mma_bar, mma_phase);
      mma_phase ^= 1;
      tc_fence_after();
      mbar_wait_relaxed(&w_full[st], (it >> 1) & 1);
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // this thread's output row of the previous tile has left
      const uint8_t* wrow = sW + st * MG_STAGE + tid * MG_WPITCH;
      uint8_t* orow = sO + tid * MG_OPITCH;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t acc[32];
        tmem_ld32(tmem_acc + lane_off + cc * 32, acc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w0 = *reinterpret_cast<const float4*>(wrow + cc * 128 + j * 32);
          const float4 w1 = *reinterpret_cast<const float4*>(wrow + cc * 128 + j * 32 + 16);
          uint4 o;
          o.x = pack_bf16(fmaf(scale, __uint_as_float(acc[8 * j]), w0.x), fmaf(scale, __uint_as_float(acc[8 * j + 1]), w0.y));
          o.y = pack_bf16(fmaf(scale, __uint_as_float(acc[8 * j + 2]), w0.z), fmaf(scale, __uint_as_float(acc[8 * j + 3]), w0.w));
          o.z = pack_bf16(fmaf(scale, __uint_as_float(acc[8 * j + 4]), w1.x), fmaf(scale, __uint_as_float(acc[8 * j + 5]), w1.y));
          o.w = pack_bf16(fmaf(scale, __uint_as_float(acc[8 * j + 6]), w1.z), fmaf(scale, __uint_as_float(acc[8 * j + 7]), w1.w));
          *reinterpret_cast<uint4*>(orow + cc * 64 + j * 16) = o;
        }
      }
      tc_fence_before();
      fence_async_smem();  // this thread's staged row -> visible to its own bulk store
      bulk_store_row(reinterpret_cast<bf16*>(pr.w_eff) + (long)(r0 + tid) * pr.k_in + c0, orow, 256);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&w_empty[st]);  // the ring stage has been read by this warp's 32 rows
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
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
  const int cap = num_sms();
  const int grid = total_tiles < cap ? total_tiles : cap;
  static bool attr = false;
  if (!attr) {
    TDM_CHECK_CUDA(cudaFuncSetAttribute(lora_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM));
    attr = true;
  }
  count_launch();
  TDM_CHECK_CUDA(launch_pdl(lora_merge_kernel, dim3(grid), dim3(MG_THREADS), (size_t)MG_SMEM, (cudaStream_t)stream, problems_dev, n_problems, total_tiles, scale));
  TDM_CHECK_CUDA(cudaGetLastError());
  return TDM_OK;
}
