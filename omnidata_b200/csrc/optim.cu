// Optimizer step of the depth train step: Depth.configure_optimizers =
// torch.optim.Adam(self.parameters(), lr) (train_depth.py:381-383; default betas (0.9, 0.999), eps 1e-8,
// no weight decay, no amsgrad) behind the Trainer's gradient_clip_val = 10 (train_depth.py:425: PL clips
// the global L2 norm with torch.nn.utils.clip_grad_norm_).
//
// All parameters / gradients / moments live in ONE flat fp32 buffer each (the model is 123 M parameters:
// 493 MB per buffer), so the whole step is two launches: a deterministic sum of squares (fixed-order
// fp64 partials, last-block finalize -> the clip coefficient on the device, no host sync) and one fused
// update pass.  HBM-bound: 16 B read + 12 B written per parameter.
#include "common.cuh"
#include "host_util.h"
#include "../../include/omnidata_b200.h"

namespace odb {

constexpr int kOptThreads = 256;
constexpr int kMaxNormBlocks = 2048;      // fixes the workspace size (no device query needed)

ODB_DEVINL double opt_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// partial[blockIdx.x] = sum of g^2 over the block's contiguous slab; the last block to finish adds the
// partials in index order and writes out2 = (total_norm, clip_coef = min(1, max_norm / (total_norm + 1e-6)))
__global__ void __launch_bounds__(kOptThreads) grad_norm_kernel(const float* __restrict__ g, long long n,
                                                                long long per_block, double* __restrict__ partial,
                                                                unsigned int* __restrict__ ticket, float max_norm,
                                                                float* __restrict__ out2) {
  __shared__ double scratch[kOptThreads / 32];
  __shared__ bool last;
  const long long lo = (long long)blockIdx.x * per_block;
  long long hi = lo + per_block;
  if (hi > n) hi = n;
  double acc = 0.0;
  // per_block is a multiple of 4 * blockDim: 16-byte loads, each thread a fixed set of elements
  for (long long i = lo + 4LL * threadIdx.x; i < hi; i += 4LL * kOptThreads) {
    if (i + 3 < hi) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(g + i));
      acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    } else {
      for (long long j = i; j < hi; ++j) acc += (double)g[j] * g[j];
    }
  }
  acc = opt_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kOptThreads / 32; ++w) s += scratch[w];
    partial[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double s = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) s += reinterpret_cast<volatile double*>(partial)[b];
    const float norm = (float)sqrt(s);
    const float coef = max_norm / (norm + 1e-6f);      // clip_grad_norm_: max_norm / (total_norm + 1e-6), clamped to 1
    out2[0] = norm;
    out2[1] = coef < 1.0f ? coef : 1.0f;
    *ticket = 0;                                       // re-armed for the next step
  }
}

// torch.optim.Adam single-tensor update (adam.py _single_tensor_adam, maximize / amsgrad / weight_decay off):
//   g' = g * clip;  m = lerp(m, g', 1 - beta1);  v = v * beta2 + (1 - beta2) g'^2
//   p -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
__global__ void __launch_bounds__(kOptThreads) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                                float* __restrict__ m, float* __restrict__ v,
                                                                long long n, const float* __restrict__ clip2,
                                                                float w1, float beta2, float w2, float step_size,
                                                                float bc2_sqrt, float eps,
                                                                const float* __restrict__ step_scalars) {
  const float clip = clip2 != nullptr ? clip2[1] : 1.0f;
  if (step_scalars != nullptr) {       // captured in a CUDA graph: the step-dependent scalars come from device memory
    step_size = step_scalars[0];
    bc2_sqrt = step_scalars[1];
  }
  const long long stride = 4LL * gridDim.x * blockDim.x;
  for (long long i = 4LL * (blockIdx.x * (long long)blockDim.x + threadIdx.x); i < n; i += stride) {
    if (i + 3 < n) {
      float4 pv = *reinterpret_cast<float4*>(p + i);
      const float4 gv = __ldg(reinterpret_cast<const float4*>(g + i));
      float4 mv = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      float* pp = &pv.x; const float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gg = __fmul_rn(gp[k], clip);
        mp[k] = __fadd_rn(mp[k], __fmul_rn(w1, __fsub_rn(gg, mp[k])));
        vp[k] = __fadd_rn(__fmul_rn(vp[k], beta2), __fmul_rn(__fmul_rn(w2, gg), gg));
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vp[k]), bc2_sqrt), eps);
        pp[k] = __fsub_rn(pp[k], __fmul_rn(step_size, __fdiv_rn(mp[k], denom)));
      }
      *reinterpret_cast<float4*>(p + i) = pv;
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (long long j = i; j < n; ++j) {
        const float gg = __fmul_rn(g[j], clip);
        const float mm = __fadd_rn(m[j], __fmul_rn(w1, __fsub_rn(gg, m[j])));
        const float vv = __fadd_rn(__fmul_rn(v[j], beta2), __fmul_rn(__fmul_rn(w2, gg), gg));
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv), bc2_sqrt), eps);
        m[j] = mm; v[j] = vv;
        p[j] = __fsub_rn(p[j], __fmul_rn(step_size, __fdiv_rn(mm, denom)));
      }
    }
  }
}

}  // namespace odb

using namespace odb;

extern "C" int64_t odb_grad_norm_workspace_bytes(void) { return (int64_t)kMaxNormBlocks * 8 + 256; }

extern "C" int odb_clip_grad_norm(const float* grads, int64_t n, float max_norm, void* workspace, float* out2,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!grads || !workspace || !out2 || n < 1 || (reinterpret_cast<uintptr_t>(grads) & 15u) ||
      (reinterpret_cast<uintptr_t>(workspace) & 7u))
    return fail(ODB_ERR_INVALID, "clip_grad_norm: bad argument (16-byte aligned gradients, workspace required)");
  int max_blocks = num_sms() * 8;
  if (max_blocks > kMaxNormBlocks) max_blocks = kMaxNormBlocks;
  const long long quantum = 4LL * kOptThreads;
  long long per_block = (n + max_blocks - 1) / max_blocks;
  per_block = (per_block + quantum - 1) / quantum * quantum;
  const int blocks = (int)((n + per_block - 1) / per_block);
  // layout of the workspace: [ticket (zero-initialised by the caller once; the kernel re-arms it)] [partials]
  unsigned int* ticket = static_cast<unsigned int*>(workspace);
  double* partial = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
  grad_norm_kernel<<<blocks, kOptThreads, 0, stream>>>(grads, n, per_block, partial, ticket, max_norm, out2);
  count_launch();
  return check_launch("clip_grad_norm");
}

extern "C" int odb_adam_step_scalars(float lr, float beta1, float beta2, int64_t step, float* out2) {
  if (out2 == nullptr || step < 1) return fail(ODB_ERR_INVALID, "adam_step_scalars: bad argument");
  // scalar preparation as torch does it on the host (python floats = doubles), then one rounding to fp32
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  out2[0] = (float)((double)lr / bc1);
  out2[1] = (float)sqrt(bc2);
  return ODB_OK;
}

extern "C" int odb_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                             const float* clip2, float lr, float beta1, float beta2, float eps, int64_t step,
                             const float* step_scalars, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (step_scalars != nullptr && step < 1) step = 1;     // unused: the scalars are read from device memory
  if (!params || !grads || !exp_avg || !exp_avg_sq || n < 1 || step < 1 ||
      ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
        reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15u))
    return fail(ODB_ERR_INVALID, "adam_step: bad argument (16-byte aligned flat fp32 buffers, step >= 1)");
  // scalar preparation as torch does it on the host (python floats = doubles), then one rounding to fp32
  float host2[2];
  odb_adam_step_scalars(lr, beta1, beta2, step, host2);
  const float step_size = host2[0], bc2_sqrt = host2[1];
  const float w1 = (float)(1.0 - (double)beta1), w2 = (float)(1.0 - (double)beta2);
  long long blocks = (n / 4 + kOptThreads - 1) / kOptThreads;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_step_kernel<<<(unsigned)blocks, kOptThreads, 0, stream>>>(params, grads, exp_avg, exp_avg_sq, n, clip2, w1, beta2,
                                                                 w2, step_size, bc2_sqrt, eps, step_scalars);
  count_launch();
  return check_launch("adam_step");
}
