// Host-side helpers shared by the C-ABI entry points: error reporting, launch accounting,
// device properties and the TMA tensor-map encoder (resolved from the driver at run time so that
// the library links against libcudart only).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace odb {

int fail(int status, const char* msg);
int fail_cuda(cudaError_t e, const char* where);
int check_launch(const char* where);
void count_launch();
int num_sms();

int encode_tiled(CUtensorMap* map, CUtensorMapDataType dtype, int rank, void* base,
                 const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
                 const cuuint32_t* elem_strides, CUtensorMapSwizzle swizzle);

}  // namespace odb
