// 3-D refocus augmentation on the device (omnidata_tools/torch/data/refocus_augmentation.py):
// a thin-lens depth-of-field effect driven by a (predicted) depth map.
//   compute_quantiles            :82-87   n_quantiles + 1 depth quantiles per image (torch.quantile, linear), ends -/+ eps
//   circle of confusion          :76-78   blur radius per quantile  r = aperture |q - focus| / q
//   separable_gaussian / stack   :30-58, :105-121  one Gaussian-blurred copy of the image per quantile (replicate
//                                          padding, cutoff = int(3 r) made odd, r < 0.1 -> copy)
//   membership + composite       :90-103, :124-141  each pixel blends the two stack entries bracketing its depth with
//                                          weights 1 - dist^2
// Kernels: exact quantiles by radix select (no sort), two separable passes per stack level (thread = output pixel,
// weights tabulated in shared memory), one fused membership + composite pass.  HBM-bound except for wide blurs.
#include "common.cuh"
#include "host_util.h"
#include "select.cuh"
#include "../../include/omnidata_b200.h"

namespace odb {

constexpr int kMaxCutoff = 4097;

// grid (n_q + 1, b): quantile q_i = i / n_q of depth[b] (torch.quantile 'linear': rank = q (n-1), lerp of the two
// neighbouring order statistics), then q_0 -= eps, q_last += eps
__global__ void __launch_bounds__(kLossThreads) refocus_quantiles_kernel(const float* __restrict__ depth, int hw,
                                                                         int n_q, float eps, float* __restrict__ qv) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t bc[2];
  const int i = blockIdx.x, b = blockIdx.y;
  const float* d = depth + (long long)b * hw;
  auto all = [&](long long) { return true; };
  const float q = (float)i / (float)n_q;
  const float pos = q * (float)(hw - 1);
  const long long lo = (long long)floorf(pos);
  const long long hi = min(lo + 1, (long long)hw - 1);
  const float frac = pos - (float)lo;
  const float vlo = key_to_float(block_radix_select(d, hw, (unsigned long long)lo, all, hist, bc));
  const float vhi = key_to_float(block_radix_select(d, hw, (unsigned long long)hi, all, hist, bc));
  if (threadIdx.x == 0) {
    float v = vlo + frac * (vhi - vlo);              // torch lerp (weight < 0.5 and >= 0.5 forms agree to 1 ulp)
    if (i == 0) v -= eps;
    if (i == n_q) v += eps;
    qv[b * (n_q + 1) + i] = v;
  }
}

ODB_DEVINL int refocus_cutoff(float r) {
  int c = (int)(r * 3.0f);                           // get_blur_stack_single_image: int(r * cutoff_multiplier), made odd
  if ((c & 1) == 0) c += 1;
  return c;
}

// One separable pass.  grid (x blocks, rows, b * levels * 3 planes); DIR 0: along x, 1: along y.
// out = conv(replicate-padded in, gaussian(cutoff, std = r)) / sum(gaussian)       (:41-56)
template <int DIR>
__global__ void __launch_bounds__(256) refocus_blur_kernel(const float* __restrict__ in, const float* __restrict__ radii,
                                                           int levels, int h, int w, int in_per_level,
                                                           float* __restrict__ out) {
  extern __shared__ float wts[];
  const int plane = blockIdx.z;                       // (b * levels + level) * 3 + c
  const int c = plane % 3, bl = plane / 3;
  const int level = bl % levels, b = bl / levels;
  const float r = radii[b * levels + level];
  const float* src = in + ((long long)(in_per_level ? bl : b) * 3 + c) * h * w;
  float* dst = out + (long long)plane * h * w;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (r < 0.1f) {                                     // separable_gaussian: `if r < 1e-1: return img`
    if (x < w) dst[(long long)y * w + x] = src[(long long)y * w + x];
    return;
  }
  const int cutoff = min(refocus_cutoff(r), kMaxCutoff);
  const int half = cutoff / 2;
  const float sig2 = 2.0f * r * r;
  float part = 0.f;
  for (int k = threadIdx.x; k < cutoff; k += blockDim.x) {
    const float n = (float)k - (float)(cutoff - 1) * 0.5f;
    const float wk = expf(-(n * n) / sig2);
    wts[k] = wk;
    part += wk;
  }
  // block sum of the weights (fixed order)
  __shared__ float wsum[8];
  for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = part;
  __syncthreads();
  float filsum = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) filsum += wsum[i];
  if (x >= w) return;
  float acc = 0.f;
  if (DIR == 0) {
    const float* row = src + (long long)y * w;
    for (int k = 0; k < cutoff; ++k) {
      int xx = x + k - half;
      xx = xx < 0 ? 0 : (xx > w - 1 ? w - 1 : xx);
      acc = fmaf(row[xx], wts[k], acc);
    }
  } else {
    for (int k = 0; k < cutoff; ++k) {
      int yy = y + k - half;
      yy = yy < 0 ? 0 : (yy > h - 1 ? h - 1 : yy);
      acc = fmaf(src[(long long)yy * w + x], wts[k], acc);
    }
  }
  dst[(long long)y * w + x] = acc / filsum;
}

// membership (:90-103) + composite (:124-141): out[b][c][y][x]
__global__ void __launch_bounds__(256) refocus_composite_kernel(const float* __restrict__ stack,
                                                                const float* __restrict__ depth,
                                                                const float* __restrict__ qv, int levels, int hw,
                                                                float* __restrict__ out, int32_t* __restrict__ segments) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw) return;
  const float d = depth[(long long)b * hw + i];
  const float* q = qv + b * levels;
  int idx = 0;                                        // torch.searchsorted (left): first index with q[idx] >= d
  while (idx < levels && q[idx] < d) ++idx;
  idx = idx < 1 ? 1 : (idx > levels - 1 ? levels - 1 : idx);
  const float ql = q[idx - 1], qr = q[idx];
  const float dist = qr - ql;
  const float dl = (d - ql) / dist, dr = (qr - d) / dist;
  const float sl = 1.f - dl * dl, sr = 1.f - dr * dr;
  const float tot = sl + sr;
  const float wl = sl / tot, wr = sr / tot;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float a = stack[(((long long)b * levels + idx - 1) * 3 + c) * hw + i];
    const float e = stack[(((long long)b * levels + idx) * 3 + c) * hw + i];
    out[((long long)b * 3 + c) * hw + i] = wl * a + wr * e;
  }
  if (segments != nullptr) segments[(long long)b * hw + i] = idx - 1;
}

}  // namespace odb

using namespace odb;

extern "C" int odb_refocus_quantiles(const float* depth, int32_t b, int32_t h, int32_t w, int32_t n_quantiles,
                                     float eps, float* quantile_vals, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!depth || !quantile_vals || b < 1 || h < 1 || w < 1 || n_quantiles < 1 || n_quantiles > 255)
    return fail(ODB_ERR_INVALID, "refocus_quantiles: bad argument");
  refocus_quantiles_kernel<<<dim3(n_quantiles + 1, b), kLossThreads, 0, stream>>>(depth, h * w, n_quantiles, eps,
                                                                                 quantile_vals);
  count_launch();
  return check_launch("refocus_quantiles");
}

extern "C" int odb_refocus_compose(const float* rgb, const float* depth, const float* quantile_vals,
                                   const float* blur_radii, int32_t b, int32_t h, int32_t w, int32_t levels,
                                   float* stack_tmp, float* stack, float* out, int32_t* segments, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!rgb || !depth || !quantile_vals || !blur_radii || !stack_tmp || !stack || !out || b < 1 || h < 1 || w < 1 ||
      levels < 2 || levels > 256 || h > 65535 || (long long)b * levels * 3 > 65535)
    return fail(ODB_ERR_INVALID, "refocus_compose: bad argument");
  const dim3 grid((w + 255) / 256, h, b * levels * 3);
  const size_t smem = kMaxCutoff * sizeof(float);
  refocus_blur_kernel<0><<<grid, 256, smem, stream>>>(rgb, blur_radii, levels, h, w, 0, stack_tmp);
  count_launch();
  refocus_blur_kernel<1><<<grid, 256, smem, stream>>>(stack_tmp, blur_radii, levels, h, w, 1, stack);
  count_launch();
  refocus_composite_kernel<<<dim3((h * w + 255) / 256, b), 256, 0, stream>>>(stack, depth, quantile_vals, levels, h * w,
                                                                            out, segments);
  count_launch();
  return check_launch("refocus_compose");
}
