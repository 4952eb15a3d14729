// C-ABI plumbing: error strings, launch counter, device queries, tensor-map encoding.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "host_util.h"
#include "../../include/omnidata_b200.h"

namespace odb {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int fail(int status, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return status;
}
int fail_cuda(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
  return ODB_ERR_CUDA;
}
int check_launch(const char* where) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail_cuda(e, where);
  }
  return ODB_OK;
}
static unsigned long long* g_trace = nullptr;
unsigned long long* debug_trace() { return g_trace; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

bool pdl_enabled() { return true; }

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
  return dev < 0 ? 0 : (dev >= kMaxDevices ? kMaxDevices - 1 : dev);
}

int num_sms() {
  static int sms[kMaxDevices] = {};
  const int dev = current_device();
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) {
      cudaGetLastError();
      v = 148;
    }
    sms[dev] = v;
  }
  return sms[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    if (e != cudaSuccess || sym == nullptr || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

int encode_tiled(CUtensorMap* map, CUtensorMapDataType dtype, int rank, void* base,
                 const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                 const cuuint32_t* elem_strides, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr)
    return fail(ODB_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver / no device)");
  CUresult r = fn(map, dtype, (cuuint32_t)rank, base, dims, strides_bytes, box, elem_strides,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err),
             "cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u]",
             (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
             (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
             box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return ODB_ERR_CUDA;
  }
  return ODB_OK;
}

}  // namespace odb

extern "C" int odb_abi_version(void) { return ODB_ABI_VERSION; }
extern "C" const char* odb_last_error(void) { return odb::g_err; }
extern "C" int odb_debug_conv_trace(void* device_buffer) {
  odb::g_trace = static_cast<unsigned long long*>(device_buffer);
  return odb::kTraceSlots;
}
extern "C" int64_t odb_launch_count(void) { return odb::g_launches.load(); }

extern "C" int odb_fill_zero(void* ptr, int64_t bytes, void* stream) {
  if (ptr == nullptr || bytes < 0) return odb::fail(ODB_ERR_INVALID, "fill_zero: bad argument");
  cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return odb::fail_cuda(e, "fill_zero");
  return ODB_OK;
}
