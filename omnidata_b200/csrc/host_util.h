// Host-side helpers shared by the C-ABI entry points: error reporting, launch accounting,
// device properties and the TMA tensor-map encoder (resolved from the driver at run time so that
// the library links against libcudart only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace odb {

int fail(int status, const char* msg);
int fail_cuda(cudaError_t e, const char* where);
int check_launch(const char* where);
void count_launch();
int num_sms();            // of the current device (cached per device)
constexpr int kMaxDevices = 64;
int current_device();     // cudaGetDevice, clamped to [0, kMaxDevices)

int encode_tiled(CUtensorMap* map, CUtensorMapDataType dtype, int rank, void* base,
                 const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                 const cuuint32_t* elem_strides, CUtensorMapSwizzle swizzle);

bool pdl_enabled();

// fp32 correctness mode of odb_conv_gemm (fp32_path.cu)


// diagnostics (odb_debug_conv_trace): device buffer of kTraceSlots uint64 per CTA, or nullptr
constexpr int kTraceSlots = 128;
unsigned long long* debug_trace();

// Launch with programmatic dependent launch allowed (the kernel must call grid_dep_wait()).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace odb
