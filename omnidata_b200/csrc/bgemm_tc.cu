// tcgen05 batched GEMM with selectable operand majors: weight gradients and the attention backward (bf16).
#include "common.cuh"
#include "host_util.h"
#include "../../include/omnidata_b200.h"

namespace odb {

long long conv_wgrad_tc_workspace_bytes(const odb_wgrad_desc* d) { (void)d; return 0; }
int conv_wgrad_tc(const odb_wgrad_desc* d, cudaStream_t stream) {
  (void)d; (void)stream;
  return fail(ODB_ERR_UNSUPPORTED, "conv_wgrad: bf16 tensor-core path not built yet");
}
long long attention_bwd_tc_workspace_bytes(int b, int tokens, int heads) { (void)b; (void)tokens; (void)heads; return 0; }
int attention_bwd_tc(const void*, const void*, const void*, const float*, void*, void*, long long, int, int, int, float,
                     cudaStream_t) {
  return fail(ODB_ERR_UNSUPPORTED, "attention_bwd: bf16 tensor-core path not built yet");
}

}  // namespace odb
