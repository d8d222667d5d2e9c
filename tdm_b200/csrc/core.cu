// Library plumbing: error channel, device info, TMA tensor-map encoding through the driver entry
// point (no link-time dependency on libcuda).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace tdm {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
unsigned long long launches(int reset) {
  unsigned long long v = __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
  if (reset) __atomic_store_n(&g_launches, 0ull, __ATOMIC_RELAXED);
  return v;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TDM_B200_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tma_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, TmaSwizzle swz) {
  return make_tma_typed(map, base, 0, rank, dims, strides_bytes, box, swz);
}

int make_tma_typed(CUtensorMap* map, const void* base, int is_f32, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box, TmaSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(TDM_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gdims[5];
  cuuint64_t gstr[4];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapSwizzle s = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swz == TMA_SWZ_32) s = CU_TENSOR_MAP_SWIZZLE_32B;
  if (swz == TMA_SWZ_64) s = CU_TENSOR_MAP_SWIZZLE_64B;
  if (swz == TMA_SWZ_128) s = CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(map, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr,
                  gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, s, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(TDM_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (CUresult %d): rank %d base %p dims [%llu,%llu,%llu,%llu] "
                     "stride0 %llu box [%u,%u]",
                     (int)r, rank, base, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                     (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0);
  }
  return TDM_OK;
}

}  // namespace tdm

extern "C" const char* tdm_last_error(void) { return tdm::g_err; }
extern "C" int tdm_version(void) { return 2; }
extern "C" uint64_t tdm_launch_count(int reset) { return tdm::launches(reset); }
extern "C" int tdm_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  TDM_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  TDM_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return TDM_OK;
}
