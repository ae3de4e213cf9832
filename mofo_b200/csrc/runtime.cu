// Host-side runtime of libmofo_sm100.so: error reporting, device query, TMA tensor-map encoding.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "../../include/mofo_b200.h"
#include "common.cuh"

namespace mofo {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return MOFO_ERR_CUDA;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

bool pdl_enabled() {
  // measured on B200 (bench.py, CUDA-graph replay): 15.04 ms/step with PDL vs 14.84 ms without -> opt-in only
  static const bool on = [] { const char* e = getenv("MOFO_B200_PDL"); return e && e[0] == '1'; }();
  return on;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return MOFO_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15) || box_rows == 0 || box_rows > 256 || box_cols != 64) {
    set_error("tensor map: base must be 16-B aligned, ld*2 a multiple of 16, box_rows in 1..256 (base=%p ld=%llu box=%u)",
              base, static_cast<unsigned long long>(ld), box_rows);
    return MOFO_ERR_INVALID;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu ld=%llu box_rows=%u)", static_cast<int>(r),
              static_cast<unsigned long long>(rows), static_cast<unsigned long long>(cols),
              static_cast<unsigned long long>(ld), box_rows);
    return MOFO_ERR_CUDA;
  }
  return MOFO_OK;
}

}  // namespace mofo

extern "C" {

int mofo_version(void) { return MOFO_B200_VERSION; }

const char* mofo_last_error(void) { return mofo::g_err; }

int mofo_sm_count(void) { return mofo::sm_count(); }

}  // extern "C"
