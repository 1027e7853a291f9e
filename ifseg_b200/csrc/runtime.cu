// Host-side runtime glue: thread-local error string, launch counter, TMA descriptor encoding
// through the driver entry point (the library does not link libcuda).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>

#include <cudaTypedefs.h>

#include "common.cuh"

namespace sgf {

static thread_local char g_err[1024] = "";
static std::atomic<int64_t> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
    set_last_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  return fn;
}

int encode_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz,
                const uint32_t* elem_strides) {
  auto fn = get_encode_fn();
  if (!fn) return SGF_ERR_CUDA;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = fn(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (CUresult %d): rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u] "
                   "stride0=%llu base=%p",
                   static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                   (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                   rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0,
                   (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), base);
    return SGF_ERR_CUDA;
  }
  return SGF_OK;
}

}  // namespace sgf

extern "C" const char* sgf_last_error(void) { return sgf::g_err; }
extern "C" int sgf_abi_version(void) { return 3; }
extern "C" int64_t sgf_launch_count(void) { return sgf::g_launches.load(); }
extern "C" void sgf_reset_launch_count(void) { sgf::g_launches.store(0); }
