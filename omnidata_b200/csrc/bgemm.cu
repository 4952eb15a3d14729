// Dispatch of the backward contractions (odb_conv_wgrad, odb_attention_bwd) by storage type.
#include "common.cuh"
#include "host_util.h"
#include "../../include/omnidata_b200.h"

namespace odb {
int conv_wgrad_f32(const odb_wgrad_desc* d, cudaStream_t stream);                                   // fp32_path.cu
int attention_bwd_f32(const float* qkv, const float* o, const float* d_o, float* dqkv, void* workspace,
                      long long workspace_bytes, int b, int tokens, int heads, float scale, cudaStream_t stream);
int conv_wgrad_tc(const odb_wgrad_desc* d, cudaStream_t stream);                                    // bgemm_tc.cu
long long conv_wgrad_tc_workspace_bytes(const odb_wgrad_desc* d);
int attention_bwd_tc(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, void* workspace,
                     long long workspace_bytes, int b, int tokens, int heads, float scale, cudaStream_t stream);
long long attention_bwd_tc_workspace_bytes(int b, int tokens, int heads);
}  // namespace odb

using namespace odb;

static int wgrad_desc_ok(const odb_wgrad_desc* d) {
  if (d == nullptr) return fail(ODB_ERR_INVALID, "conv_wgrad: null descriptor");
  if (d->num_views < 1 || d->num_views > ODB_MAX_VIEWS || d->num_taps < 1 || d->num_taps > ODB_MAX_TAPS)
    return fail(ODB_ERR_INVALID, "conv_wgrad: bad view/tap count");
  for (int t = 0; t < d->num_taps; ++t)
    if (d->tap_view[t] < 0 || d->tap_view[t] >= d->num_views) return fail(ODB_ERR_INVALID, "conv_wgrad: tap refers to a missing view");
  if (d->out == nullptr || d->dy.ptr == nullptr || d->n < 1 || d->dy.c != d->n) return fail(ODB_ERR_INVALID, "conv_wgrad: bad dy / out");
  return ODB_OK;
}

extern "C" int64_t odb_conv_wgrad_workspace_bytes(const odb_wgrad_desc* d) {
  if (wgrad_desc_ok(d)) return -1;
  const long long row = (long long)d->num_taps * d->views[0].c;
  if (d->dtype == ODB_DTYPE_F32) {
    // up to 256 pixel-range splits (the kernel uses fewer when the workspace is smaller)
    const long long tiles = (long long)((d->n + 63) / 64) * ((d->views[0].c + 63) / 64) * d->num_taps;
    long long splits = (4LL * num_sms() + tiles - 1) / tiles;
    if (splits > 256) splits = 256;
    if (splits < 1) splits = 1;
    return splits * d->n * row * 4;
  }
  return conv_wgrad_tc_workspace_bytes(d);
}

extern "C" int odb_conv_wgrad(const odb_wgrad_desc* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = wgrad_desc_ok(d);
  if (rc) return rc;
  if (d->dtype == ODB_DTYPE_F32) return conv_wgrad_f32(d, stream);
  if (d->dtype == ODB_DTYPE_BF16) return conv_wgrad_tc(d, stream);
  return fail(ODB_ERR_INVALID, "conv_wgrad: dtype must be ODB_DTYPE_BF16 or ODB_DTYPE_F32");
}

extern "C" int64_t odb_attention_bwd_workspace_bytes(int32_t b, int32_t tokens, int32_t heads, int32_t dtype) {
  if (b < 1 || tokens < 1 || heads < 1) return -1;
  if (dtype == ODB_DTYPE_F32) return 2LL * b * heads * tokens * tokens * 4;
  return attention_bwd_tc_workspace_bytes(b, tokens, heads);
}

extern "C" int odb_attention_bwd(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv,
                                 void* workspace, int64_t workspace_bytes, int32_t b, int32_t tokens, int32_t heads,
                                 float scale, int32_t dtype, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!qkv || !o || !d_o || !dqkv || b < 1 || tokens < 1 || heads < 1) return fail(ODB_ERR_INVALID, "attention_bwd: bad argument");
  if (dtype == ODB_DTYPE_F32)
    return attention_bwd_f32(static_cast<const float*>(qkv), static_cast<const float*>(o), static_cast<const float*>(d_o),
                             static_cast<float*>(dqkv), workspace, workspace_bytes, b, tokens, heads, scale, stream);
  if (dtype == ODB_DTYPE_BF16)
    return attention_bwd_tc(qkv, o, d_o, lse, dqkv, workspace, workspace_bytes, b, tokens, heads, scale, stream);
  return fail(ODB_ERR_INVALID, "attention_bwd: dtype must be ODB_DTYPE_BF16 or ODB_DTYPE_F32");
}
