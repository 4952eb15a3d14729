// Forward kernels of the depth-training losses of omnidata_tools/torch:
//   * make_valid_mask            train_depth.py:215-242
//   * MidasLoss.forward          losses/midas_loss.py:137-157  (SSIMAE :33-56,104-111; scale/shift :10-30;
//                                 4-scale gradient matching :59-100,114-134; masked_l1 losses/masked_losses.py:4-7)
//   * VNL_Loss.forward           losses/virtual_normal_loss.py:29-194
// All inputs are fp32 (the reference trains in fp32).  Reductions are deterministic: per-thread fp64
// partials combined in a fixed shuffle/shared-memory order, medians / order statistics by an exact
// 4-pass radix select with integer histograms.  MidasLoss and VNL_Loss also have their backward passes
// here (gradient with respect to the prediction: the first step of the train step's backward; the network's
// own backward is not built).
#include "common.cuh"
#include "host_util.h"
#include "select.cuh"
#include "../../include/omnidata_b200.h"

namespace odb {

// ------------------------------------------------------------------------------------------ make_valid_mask
// valid = nearest_upsample(max_pool2d(1 - m, k)) == 0        (train_depth.py:234-238)
__global__ void __launch_bounds__(256) make_valid_mask_kernel(const float* __restrict__ m,
                                                              uint8_t* __restrict__ valid, int b, int h,
                                                              int w, int k) {
  const int hp = h / k, wp = w / k;
  const float sy = (float)hp / (float)h, sx = (float)wp / (float)w;   // F.interpolate 'nearest' scale
  const long long total = (long long)b * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int bi = (int)(i / ((long long)w * h));
    const int py = min((int)floorf(y * sy), hp - 1), px = min((int)floorf(x * sx), wp - 1);
    float mx = -INFINITY;
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx)
        mx = fmaxf(mx, 1.0f - m[((long long)bi * h + py * k + dy) * w + px * k + dx]);
    valid[i] = (mx == 0.0f) ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------ MiDaS loss
struct MidasWs {
  float* med;      // [2][b]  nanmedian of (pred, target) over valid pixels, 0 if none
  float* scale;    // [2][b]  mean absolute deviation sum|x - med| / (n_mask + 1)
  double* sums;    // [b][8]  ssi_abs, n_mask, a00, a01, a11, b0, b1
  double* grad;    // [b][4][2] (gradient loss, mask count) per scale
};

// grid (2, b): tensor 0 = prediction, 1 = target.  masked_shift_and_scale (midas_loss.py:33-56)
__global__ void __launch_bounds__(kLossThreads) midas_median_kernel(const float* __restrict__ pred,
                                                                    const float* __restrict__ target,
                                                                    const uint8_t* __restrict__ mask,
                                                                    MidasWs ws, int b_n, int hw) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t bc[2];
  __shared__ double scratch[32];
  __shared__ unsigned long long s_n;
  __shared__ float s_med;
  const int sel = blockIdx.x, b = blockIdx.y;
  const float* v = (sel == 0 ? pred : target) + (long long)b * hw;
  const uint8_t* mk = mask + (long long)b * hw;
  auto take = [&](long long i) { return mk[i] != 0 && !isnan(v[i]); };
  // count of non-NaN valid values (integer, deterministic)
  double cnt = 0.0, nmask = 0.0;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    if (mk[i]) { nmask += 1.0; if (!isnan(v[i])) cnt += 1.0; }
  }
  const double tot = block_sum_d(cnt, scratch);
  const double totm = block_sum_d(nmask, scratch);
  if (threadIdx.x == 0) s_n = (unsigned long long)tot;
  __syncthreads();
  const unsigned long long n = s_n;
  float med = 0.f;   // t[isnan(t)] = 0
  if (n > 0) {
    const uint32_t key = block_radix_select(v, hw, (n - 1) / 2, take, hist, bc);   // lower median
    med = key_to_float(key);
  }
  // mean absolute deviation over the mask:  sum(|x - t| over mask) / (n_mask + 1)
  double dev = 0.0;
  for (int i = threadIdx.x; i < hw; i += blockDim.x)
    if (mk[i]) dev += (double)fabsf(v[i] - med);
  const double devs = block_sum_d(dev, scratch);
  if (threadIdx.x == 0) {
    s_med = med;
    ws.med[sel * b_n + b] = med;
    // n_mask is a block_sum_d result valid in thread 0 only: recompute from the shared copy below
  }
  __syncthreads();
  if (threadIdx.x == 0) ws.scale[sel * b_n + b] = (float)devs / ((float)totm + 1.0f);
}

// grid (b): SSIMAE numerator + the five sums of compute_scale_and_shift (midas_loss.py:10-30, :147-151)
__global__ void __launch_bounds__(kLossThreads) midas_sums_kernel(const float* __restrict__ pred,
                                                                  const float* __restrict__ target,
                                                                  const uint8_t* __restrict__ mask,
                                                                  MidasWs ws, int b_n, int hw) {
  __shared__ double scratch[32];
  const int b = blockIdx.x;
  const float* p = pred + (long long)b * hw;
  const float* g = target + (long long)b * hw;
  const uint8_t* mk = mask + (long long)b * hw;
  const float tp = ws.med[b], tg = ws.med[b_n + b];
  const float sp = ws.scale[b] + 1e-6f, sg = ws.scale[b_n + b] + 1e-6f;
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    if (!mk[i]) continue;
    const float pv = p[i], gv = g[i];
    const float pa = (pv - tp) / sp, ga = (gv - tg) / sg;
    acc[0] += (double)fabsf(pa - ga);
    acc[1] += 1.0;
    const float pi = 1.0f / (pv + 1e-6f), ti = 1.0f / (gv + 1e-6f);
    acc[2] += (double)(pi * pi);
    acc[3] += (double)pi;
    acc[4] += 1.0;
    acc[5] += (double)(pi * ti);
    acc[6] += (double)ti;
  }
  for (int k = 0; k < 7; ++k) {
    const double r = block_sum_d(acc[k], scratch);
    if (threadIdx.x == 0) ws.sums[b * 8 + k] = r;
  }
}

// grid (scales, b): gradient_loss at stride 2^s on prediction_ssi = scale * 1/(p+eps) + shift
// (midas_loss.py:83-100, 128-132, 151-153)
__global__ void __launch_bounds__(kLossThreads) midas_grad_kernel(const float* __restrict__ pred,
                                                                  const float* __restrict__ target,
                                                                  const uint8_t* __restrict__ mask,
                                                                  MidasWs ws, int h, int w) {
  __shared__ double scratch[32];
  const int s = blockIdx.x, b = blockIdx.y;
  const int step = 1 << s;
  const int hs = (h + step - 1) / step, wsz = (w + step - 1) / step;
  const long long img = (long long)b * h * w;
  // closed-form 2x2 solve in fp32 exactly as the reference does it
  const float a00 = (float)ws.sums[b * 8 + 2], a01 = (float)ws.sums[b * 8 + 3], a11 = (float)ws.sums[b * 8 + 4];
  const float b0 = (float)ws.sums[b * 8 + 5], b1 = (float)ws.sums[b * 8 + 6];
  const float det = a00 * a11 - a01 * a01;
  float x0 = 0.f, x1 = 0.f;
  if (det != 0.f) {
    x0 = (a11 * b0 - a01 * b1) / (det + 1e-6f);
    x1 = (-a01 * b0 + a00 * b1) / (det + 1e-6f);
  }
  auto dval = [&](int y, int x, float& m) {
    const long long i = img + (long long)(y * step) * w + x * step;
    m = mask[i] ? 1.f : 0.f;
    const float pssi = x0 * (1.0f / (pred[i] + 1e-6f)) + x1;
    return m * (pssi - 1.0f / (target[i] + 1e-6f));
  };
  double loss = 0.0, cnt = 0.0;
  const int total = hs * wsz;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int y = i / wsz, x = i - y * wsz;
    float m0;
    const float d0 = dval(y, x, m0);
    cnt += (double)m0;
    if (x + 1 < wsz) {
      float m1;
      const float d1 = dval(y, x + 1, m1);
      loss += (double)(m1 * m0 * fabsf(d1 - d0));
    }
    if (y + 1 < hs) {
      float m1;
      const float d1 = dval(y + 1, x, m1);
      loss += (double)(m1 * m0 * fabsf(d1 - d0));
    }
  }
  const double l = block_sum_d(loss, scratch);
  const double c = block_sum_d(cnt, scratch);
  if (threadIdx.x == 0) { ws.grad[(b * 4 + s) * 2] = l; ws.grad[(b * 4 + s) * 2 + 1] = c; }
}

__global__ void midas_finalize_kernel(MidasWs ws, int b_n, int scales, float alpha, float* __restrict__ out3) {
  if (threadIdx.x != 0) return;
  double num = 0.0, den = 0.0;
  for (int b = 0; b < b_n; ++b) { num += ws.sums[b * 8 + 0]; den += ws.sums[b * 8 + 1]; }
  const float ssi = (float)(num / den);               // masked_l1_loss: sum / mask.sum()
  float reg = 0.f;
  for (int s = 0; s < scales; ++s) {
    double acc = 0.0;                                  // reduction_image_based: mean_b(loss_b / M_b)
    for (int b = 0; b < b_n; ++b) {
      const double l = ws.grad[(b * 4 + s) * 2], m = ws.grad[(b * 4 + s) * 2 + 1];
      acc += (m != 0.0) ? l / m : l;
    }
    reg += (float)(acc / b_n);
  }
  out3[1] = ssi;
  out3[2] = reg;
  out3[0] = ssi + alpha * reg;
}

// ------------------------------------------------------------------------------------------ MiDaS loss, backward
// d(w_ssi * ssi + w_reg * reg) / d(prediction), what autograd produces for MidasLoss.forward (midas_loss.py:137-157).
// Needs the forward workspace of the SAME inputs (medians, deviations, the five sums, per-scale mask counts).
// With V the valid pixels of an image, n = |V|, N = sum of n over the batch, t / s the median / deviation of the
// prediction, pa = (p - t)/(s + eps), e = sign(pa - ga), sigma = sign(p - t)  (both zero outside V):
//   ssi:  dL/dp_j = e_j / (N (s+eps)) + dL/ds * sigma_j / (n+1) + [j = median element] dL/dt,
//         dL/ds = -sum(e (p - t)) / (N (s+eps)^2),  dL/dt = -sum(e) / (N (s+eps)) - dL/ds * sum(sigma) / (n+1)
//         (the nanmedian passes its gradient to one element: here the lowest index holding the median value)
//   reg:  P = x0 q + x1, q = 1/(p+eps);  G = d reg / d P (signs of the masked forward differences at the four scales,
//         weighted 1/(B M_s)),  gx0 = sum(G q), gx1 = sum(G);  x0, x1 depend on p through a00, a01, b0:
//         d reg/dq_j = G_j x0 + [j in V] (2 C00 q_j + C01 + Cb0 u_j),  dq/dp = -q^2.
struct MidasBwdWs {
  double* sc;      // [b][16]: 0 inv = 1/(N (s+eps)), 1 dL/ds / (n+1), 2 dL/dt, 3 median index, 4 gx0, 5 gx1, 6 x0, 7 x1
};

// grid (b): the ssi scalars of one image
__global__ void __launch_bounds__(kLossThreads) midas_bwd_ssi_kernel(const float* __restrict__ pred,
                                                                     const float* __restrict__ target,
                                                                     const uint8_t* __restrict__ mask, MidasWs ws,
                                                                     MidasBwdWs bw, int b_n, int hw) {
  __shared__ double scratch[32];
  __shared__ int s_idx[32];
  const int b = blockIdx.x;
  const float* p = pred + (long long)b * hw;
  const float* g = target + (long long)b * hw;
  const uint8_t* mk = mask + (long long)b * hw;
  const float tp = ws.med[b], tg = ws.med[b_n + b];
  const float sp = ws.scale[b] + 1e-6f, sg = ws.scale[b_n + b] + 1e-6f;
  double A = 0.0, Bs = 0.0, S = 0.0, n = 0.0;
  int first = 0x7fffffff;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    if (!mk[i]) continue;
    const float pv = p[i];
    const float pa = (pv - tp) / sp, ga = (g[i] - tg) / sg;
    const float d = pa - ga;
    const double e = d > 0.f ? 1.0 : (d < 0.f ? -1.0 : 0.0);
    const float c = pv - tp;
    A += e;
    Bs += e * (double)c;
    S += c > 0.f ? 1.0 : (c < 0.f ? -1.0 : 0.0);
    n += 1.0;
    if (pv == tp && i < first) first = i;
  }
  const double rA = block_sum_d(A, scratch);
  const double rB = block_sum_d(Bs, scratch);
  const double rS = block_sum_d(S, scratch);
  const double rn = block_sum_d(n, scratch);
  // lowest index holding the median value
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_down_sync(0xffffffffu, first, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_idx[threadIdx.x >> 5] = first;
  __syncthreads();
  if (threadIdx.x == 0) {
    int k = 0x7fffffff;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) k = min(k, s_idx[w]);
    double N = 0.0;
    for (int bb = 0; bb < b_n; ++bb) N += ws.sums[bb * 8 + 1];
    const double se = (double)sp;
    const double inv = 1.0 / (N * se);
    const double dLds = -rB / (N * se * se);
    const double dLdt = -rA * inv - dLds * rS / (rn + 1.0);
    double* sc = bw.sc + (long long)b * 16;
    sc[0] = inv;
    sc[1] = dLds / (rn + 1.0);
    sc[2] = rn > 0.0 ? dLdt : 0.0;
    sc[3] = (rn > 0.0 && k != 0x7fffffff) ? (double)k : -1.0;
  }
}

// grid (b): G = d reg / d prediction_ssi for every pixel (all scales), gx0 = sum G q, gx1 = sum G
__global__ void __launch_bounds__(kLossThreads) midas_bwd_reg_kernel(const float* __restrict__ pred,
                                                                     const float* __restrict__ target,
                                                                     const uint8_t* __restrict__ mask, MidasWs ws,
                                                                     MidasBwdWs bw, int b_n, int h, int w, int scales,
                                                                     float* __restrict__ gbuf) {
  __shared__ double scratch[32];
  const int b = blockIdx.x;
  const long long img = (long long)b * h * w;
  const float a00 = (float)ws.sums[b * 8 + 2], a01 = (float)ws.sums[b * 8 + 3], a11 = (float)ws.sums[b * 8 + 4];
  const float b0 = (float)ws.sums[b * 8 + 5], b1 = (float)ws.sums[b * 8 + 6];
  const float det = a00 * a11 - a01 * a01;
  float x0 = 0.f, x1 = 0.f;
  if (det != 0.f) {
    x0 = (a11 * b0 - a01 * b1) / (det + 1e-6f);
    x1 = (-a01 * b0 + a00 * b1) / (det + 1e-6f);
  }
  float wsc[4];
  for (int s = 0; s < 4; ++s) {
    const double M = s < scales ? ws.grad[(b * 4 + s) * 2 + 1] : 0.0;
    wsc[s] = s < scales ? (float)(1.0 / ((double)b_n * (M != 0.0 ? M : 1.0))) : 0.f;
  }
  auto dval = [&](int y, int x, bool& m) {
    const long long i = img + (long long)y * w + x;
    m = mask[i] != 0;
    const float pssi = x0 * (1.0f / (pred[i] + 1e-6f)) + x1;
    return m ? (pssi - 1.0f / (target[i] + 1e-6f)) : 0.f;
  };
  auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
  double gx0 = 0.0, gx1 = 0.0;
  const int total = h * w;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int y = i / w, x = i - y * w;
    float G = 0.f;
    bool m0;
    const float d0 = dval(y, x, m0);
    if (m0) {
      for (int s = 0; s < scales; ++s) {
        const int st = 1 << s;
        if ((y & (st - 1)) || (x & (st - 1))) break;       // not on the grids of this and coarser scales
        float acc = 0.f;
        bool m1;
        if (x - st >= 0) { const float d1 = dval(y, x - st, m1); if (m1) acc += sgn(d0 - d1); }
        if (x + st < w) { const float d1 = dval(y, x + st, m1); if (m1) acc -= sgn(d1 - d0); }
        if (y - st >= 0) { const float d1 = dval(y - st, x, m1); if (m1) acc += sgn(d0 - d1); }
        if (y + st < h) { const float d1 = dval(y + st, x, m1); if (m1) acc -= sgn(d1 - d0); }
        G += wsc[s] * acc;
      }
      const float q = 1.0f / (pred[img + i] + 1e-6f);
      gx0 += (double)G * (double)q;
      gx1 += (double)G;
    }
    gbuf[img + i] = G;
  }
  const double r0 = block_sum_d(gx0, scratch);
  const double r1 = block_sum_d(gx1, scratch);
  if (threadIdx.x == 0) {
    double* sc = bw.sc + (long long)b * 16;
    sc[4] = r0; sc[5] = r1; sc[6] = (double)x0; sc[7] = (double)x1;
  }
}

// grid (blocks per image, b): the gradient itself
__global__ void __launch_bounds__(256) midas_bwd_final_kernel(const float* __restrict__ pred,
                                                              const float* __restrict__ target,
                                                              const uint8_t* __restrict__ mask, MidasWs ws,
                                                              MidasBwdWs bw, int b_n, int hw, float w_ssi, float w_reg,
                                                              const float* __restrict__ gbuf, float* __restrict__ grad) {
  const int b = blockIdx.y;
  const double* sc = bw.sc + (long long)b * 16;
  const float inv = (float)sc[0], dlds_n1 = (float)sc[1], dldt = (float)sc[2];
  const int kmed = (int)sc[3];
  const double gx0 = sc[4], gx1 = sc[5], x0 = sc[6], x1 = sc[7];
  const double a00 = ws.sums[b * 8 + 2], a01 = ws.sums[b * 8 + 3], a11 = ws.sums[b * 8 + 4];
  const double b0 = ws.sums[b * 8 + 5], b1 = ws.sums[b * 8 + 6];
  double c00 = 0.0, c01 = 0.0, cb0 = 0.0;
  {
    const float detf = (float)a00 * (float)a11 - (float)a01 * (float)a01;     // the forward's fp32 test for det != 0
    if (detf != 0.f) {
      const double D = a00 * a11 - a01 * a01 + 1e-6;
      c00 = gx0 * (-x0 * a11 / D) + gx1 * (b1 / D - x1 * a11 / D);
      c01 = gx0 * ((-b1 + 2.0 * a01 * x0) / D) + gx1 * ((-b0 + 2.0 * a01 * x1) / D);
      cb0 = gx0 * (a11 / D) + gx1 * (-a01 / D);
    }
  }
  const float fc00 = (float)(2.0 * c00), fc01 = (float)c01, fcb0 = (float)cb0, fx0 = (float)x0;
  const float tp = ws.med[b], tg = ws.med[b_n + b];
  const float sp = ws.scale[b] + 1e-6f, sg = ws.scale[b_n + b] + 1e-6f;
  const long long img = (long long)b * hw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    float gr = 0.f;
    if (mask[img + i]) {
      const float pv = pred[img + i], gv = target[img + i];
      const float d = (pv - tp) / sp - (gv - tg) / sg;
      const float e = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      const float c = pv - tp;
      const float sigma = c > 0.f ? 1.f : (c < 0.f ? -1.f : 0.f);
      float gs = e * inv + dlds_n1 * sigma;
      if (i == kmed) gs += dldt;
      const float q = 1.0f / (pv + 1e-6f), u = 1.0f / (gv + 1e-6f);
      const float dq = gbuf[img + i] * fx0 + (fc00 * q + fc01 + fcb0 * u);
      gr = w_ssi * gs - w_reg * q * q * dq;
    }
    grad[img + i] = gr;
  }
}

// ------------------------------------------------------------------------------------------ normal-training losses
// masked_l1_loss + masked_cosine_angular_loss (losses/masked_losses.py:4-7,14-23) as train_normal.py:247-258
// uses them: preds = clamp(model(rgb), 0, 1); mask_valid [b,1,h,w] repeated over the 3 channels;
// loss = cos + 10 * l1.  grid (b): per-image fp64 partials (sum |p - t| over valid pixels x 3 channels,
// sum of -cos over valid pixels, valid pixel count), fixed order.
__global__ void __launch_bounds__(kLossThreads) normal_loss_partial_kernel(const float* __restrict__ pred,
                                                                           const float* __restrict__ target,
                                                                           const uint8_t* __restrict__ mask,
                                                                           int hw, int clamp_pred,
                                                                           double* __restrict__ partial) {
  __shared__ double scratch[32];
  const int b = blockIdx.x;
  const float* p = pred + (long long)b * 3 * hw;
  const float* g = target + (long long)b * 3 * hw;
  const uint8_t* mk = mask + (long long)b * hw;
  double l1 = 0.0, cs = 0.0, cnt = 0.0;
  for (int i = threadIdx.x; i < hw; i += blockDim.x) {
    if (!mk[i]) continue;
    float pv[3], gv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      pv[c] = p[c * hw + i];
      if (clamp_pred) pv[c] = fminf(fmaxf(pv[c], 0.f), 1.f);          // train_normal.py:249
      gv[c] = g[c * hw + i];
      l1 += (double)fabsf(pv[c] - gv[c]);
    }
    float pn = 0.f, gn = 0.f, dot = 0.f;
    float pc[3], gc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      pc[c] = fminf(fmaxf(2.f * pv[c] - 1.f, -1.f), 1.f);
      gc[c] = fminf(fmaxf(2.f * gv[c] - 1.f, -1.f), 1.f);
      pn = fmaf(pc[c], pc[c], pn);
      gn = fmaf(gc[c], gc[c], gn);
    }
    // F.normalize(p=2, dim=1): x / max(||x||, 1e-12)
    const float pin = __fdiv_rn(1.f, fmaxf(__fsqrt_rn(pn), 1e-12f));
    const float gin = __fdiv_rn(1.f, fmaxf(__fsqrt_rn(gn), 1e-12f));
#pragma unroll
    for (int c = 0; c < 3; ++c) dot = fmaf(pc[c] * pin, gc[c] * gin, dot);
    cs -= (double)dot;
    cnt += 1.0;
  }
  const double r0 = block_sum_d(l1, scratch);
  const double r1 = block_sum_d(cs, scratch);
  const double r2 = block_sum_d(cnt, scratch);
  if (threadIdx.x == 0) { partial[b * 3 + 0] = r0; partial[b * 3 + 1] = r1; partial[b * 3 + 2] = r2; }
}

__global__ void normal_loss_finalize_kernel(const double* __restrict__ partial, int b_n, float* __restrict__ out3) {
  if (threadIdx.x != 0) return;
  double l1 = 0.0, cs = 0.0, cnt = 0.0;
  for (int b = 0; b < b_n; ++b) { l1 += partial[b * 3]; cs += partial[b * 3 + 1]; cnt += partial[b * 3 + 2]; }
  const float l1_loss = (float)(l1 / (3.0 * cnt));     // element sum / mask_valid.sum() (mask repeated x3)
  const float cos_loss = (float)(cs / cnt);            // mean over valid pixels (NaN when there are none, as the reference)
  out3[1] = l1_loss;
  out3[2] = cos_loss;
  out3[0] = cos_loss + 10.0f * l1_loss;
}

// backward of the pair: grad[b][c][h][w] = d(w_l1 * l1 + w_cos * cos) / d(prediction) (through the optional clamp)
__global__ void __launch_bounds__(256) normal_loss_bwd_kernel(const float* __restrict__ pred,
                                                              const float* __restrict__ target,
                                                              const uint8_t* __restrict__ mask,
                                                              const double* __restrict__ partial, int b_n, int hw,
                                                              int clamp_pred, float w_l1, float w_cos,
                                                              float* __restrict__ grad) {
  const int b = blockIdx.y;
  double cnt = 0.0;
  for (int i = 0; i < b_n; ++i) cnt += partial[i * 3 + 2];
  const float k_l1 = (float)((double)w_l1 / (3.0 * cnt));
  const float k_cos = (float)((double)w_cos / cnt);
  const float* p = pred + (long long)b * 3 * hw;
  const float* g = target + (long long)b * 3 * hw;
  float* o = grad + (long long)b * 3 * hw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
    float out[3] = {0.f, 0.f, 0.f};
    if (mask[(long long)b * hw + i]) {
      float pv[3], gv[3], pass[3], pc[3], gc[3], pass2[3];
      float pn = 0.f, gn = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float raw = p[c * hw + i];
        pass[c] = (!clamp_pred || (raw >= 0.f && raw <= 1.f)) ? 1.f : 0.f;      // torch.clamp backward: inclusive
        pv[c] = clamp_pred ? fminf(fmaxf(raw, 0.f), 1.f) : raw;
        gv[c] = g[c * hw + i];
        const float t = 2.f * pv[c] - 1.f;
        pass2[c] = (t >= -1.f && t <= 1.f) ? 1.f : 0.f;
        pc[c] = fminf(fmaxf(t, -1.f), 1.f);
        gc[c] = fminf(fmaxf(2.f * gv[c] - 1.f, -1.f), 1.f);
        pn = fmaf(pc[c], pc[c], pn);
        gn = fmaf(gc[c], gc[c], gn);
      }
      const float pl = __fsqrt_rn(pn), gl = __fsqrt_rn(gn);
      const float pin = __fdiv_rn(1.f, fmaxf(pl, 1e-12f)), gin = __fdiv_rn(1.f, fmaxf(gl, 1e-12f));
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) dot = fmaf(pc[c] * pin, gc[c] * gin, dot);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float d = pv[c] - gv[c];
        const float s = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        // d(-<pn, gn>)/d pc = -(gn_c - <pn, gn> pn_c) / |pc|   (F.normalize backward; below eps the norm is clamped)
        const float dcos = pl > 1e-12f ? -(gc[c] * gin - dot * pc[c] * pin) * pin : -(gc[c] * gin) * 1e12f;
        out[c] = pass[c] * (k_l1 * s + k_cos * 2.f * pass2[c] * dcos);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c * hw + i] = out[c];
  }
}

// ------------------------------------------------------------------------------------------ virtual normal loss
struct Vec3 { float x, y, z; };
ODB_DEVINL Vec3 sub3(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
ODB_DEVINL float dot3(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
ODB_DEVINL Vec3 cross3(Vec3 a, Vec3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
ODB_DEVINL float comp(Vec3 v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : v.z); }

// thread = (image, point group).  transfer_xyz :44-50, filter_mask :95-128, select_points_groups
// :130-149, per-group loss :169-188.  loss[i] = NaN for filtered-out groups.
__global__ void __launch_bounds__(256) vnl_groups_kernel(const float* __restrict__ first,
                                                         const float* __restrict__ second,
                                                         const int* __restrict__ p1, const int* __restrict__ p2,
                                                         const int* __restrict__ p3, float* __restrict__ loss,
                                                         int b_n, int n_pts, int h, int w, float fx, float fy,
                                                         float delta_z) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)b_n * n_pts) return;
  const int b = (int)(gid / n_pts), n = (int)(gid % n_pts);
  const int idx[3] = {p1[n], p2[n], p3[n]};
  const float u0 = (float)(w / 2), v0 = (float)(h / 2);
  Vec3 G[3], D[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int y = idx[j] / w, x = idx[j] - y * w;
    const float gd = first[(long long)b * h * w + idx[j]];
    const float dd = second[(long long)b * h * w + idx[j]];
    G[j] = {((float)x - u0) * fabsf(gd) / fx, ((float)y - v0) * fabsf(gd) / fy, gd};
    D[j] = {((float)x - u0) * fabsf(dd) / fx, ((float)y - v0) * fabsf(dd) / fy, dd};
  }
  // ---- filter_mask on the first argument's points
  const Vec3 d12 = sub3(G[1], G[0]), d13 = sub3(G[2], G[0]), d23 = sub3(G[2], G[1]);
  const Vec3 dv[3] = {d12, d13, d23};
  float nrm[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) nrm[i] = sqrtf(dot3(dv[i], dv[i]));
  int ncos = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float e = dot3(dv[i], dv[j]) / (nrm[i] * nrm[j] + 1e-8f);
      ncos += (e > 0.867f || e < -0.867f) ? 1 : 0;
    }
  const bool mask_cos = ncos > 3;
  const bool mask_pad = (G[0].z > delta_z) && (G[1].z > delta_z) && (G[2].z > delta_z);
  bool near_[3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
    near_[c] = fabsf(comp(d12, c)) < 0.005f || fabsf(comp(d13, c)) < 0.005f || fabsf(comp(d23, c)) < 0.005f;
  const bool ignore = (near_[0] && near_[1] && near_[2]) || mask_cos;
  if (!(mask_pad && !ignore)) { loss[gid] = __int_as_float(0x7fc00000); return; }
  // ---- reference quirk (:144): the boolean mask z_j == 0 of POINT j indexes the COORDINATE axis, so
  // coordinate j of all three points of the second argument becomes 1e-4
  const bool z0[3] = {D[0].z == 0.f, D[1].z == 0.f, D[2].z == 0.f};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (z0[0]) D[j].x = 0.0001f;
    if (z0[1]) D[j].y = 0.0001f;
    if (z0[2]) D[j].z = 0.0001f;
  }
  Vec3 gn = cross3(sub3(G[1], G[0]), sub3(G[2], G[0]));
  Vec3 dn = cross3(sub3(D[1], D[0]), sub3(D[2], D[0]));
  float gl = sqrtf(dot3(gn, gn)), dl = sqrtf(dot3(dn, dn));
  if (gl == 0.f) gl += 0.01f;
  if (dl == 0.f) dl += 0.01f;
  loss[gid] = fabsf(gn.x / gl - dn.x / dl) + fabsf(gn.y / gl - dn.y / dl) + fabsf(gn.z / gl - dn.z / dl);
}

// single block: mean of the losses after dropping the lowest 25 % (:189-193)
__global__ void __launch_bounds__(kLossThreads) vnl_reduce_kernel(const float* __restrict__ loss, long long n,
                                                                  int select, float* __restrict__ out) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t bc[2];
  __shared__ double scratch[32];
  __shared__ double s_vals[2];
  auto take = [&](long long i) { return !isnan(loss[i]); };
  double cnt = 0.0, tot = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x)
    if (take(i)) { cnt += 1.0; tot += (double)loss[i]; }
  const double c = block_sum_d(cnt, scratch);
  const double t = block_sum_d(tot, scratch);
  if (threadIdx.x == 0) { s_vals[0] = c; s_vals[1] = t; }
  __syncthreads();
  const unsigned long long count = (unsigned long long)s_vals[0];
  const double total = s_vals[1];
  const unsigned long long k = select ? (unsigned long long)((double)count * 0.25) : 0ull;
  if (count == 0) { if (threadIdx.x == 0) out[0] = __int_as_float(0x7fc00000); return; }
  double dropped = 0.0;
  if (k > 0) {
    const uint32_t key = block_radix_select(loss, n, k - 1, take, hist, bc);   // largest dropped value
    const float thr = key_to_float(key);
    double below = 0.0, nbelow = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x)
      if (take(i) && loss[i] < thr) { below += (double)loss[i]; nbelow += 1.0; }
    const double sb = block_sum_d(below, scratch);
    const double nb = block_sum_d(nbelow, scratch);
    if (threadIdx.x == 0) dropped = sb + ((double)k - nb) * (double)thr;
  }
  if (threadIdx.x == 0) out[0] = (float)((total - dropped) / (double)(count - k));
}

// ------------------------------------------------------------------------------------------ virtual normal loss, backward
// d(mean of the kept group losses) / d(first): what autograd gives for VNL_Loss.forward(first, second)
// (virtual_normal_loss.py:151-194; train_depth.py:272 passes the prediction as `first`).  The boolean masks carry no
// gradient; a kept group k contributes through its normal n = (P1 - P0) x (P2 - P0) of the FIRST argument's points:
//   w = dL/dn = (r / |n| - (r . n) n / |n|^3) / K',   r = sign(n/|n| - m/|m|),   dL/dP1 = b x w, dL/dP2 = w x a,
//   dL/dP0 = -(dL/dP1 + dL/dP2),   P = (u |d|, v |d|, d)  ->  dL/dd += Gx u sign(d)/fx + Gy v sign(d)/fy + Gz.
// Pixels are sampled with replacement, so contributions are scattered with 64-bit FIXED-POINT atomics (2^-40 units):
// integer addition is associative, the result is bit-reproducible.
constexpr double kVnlFixScale = 1099511627776.0;   // 2^40

// single block: which groups survive the "drop the lowest 25 %" selection.  sel: [0] threshold, [1] cut index (equal
// values below it are dropped: the sort order among ties is by index), [2] number of kept groups
__global__ void __launch_bounds__(kLossThreads) vnl_bwd_select_kernel(const float* __restrict__ loss, long long n,
                                                                      int select, double* __restrict__ sel) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t bc[2];
  __shared__ double scratch[32];
  __shared__ double s_cnt;
  auto take = [&](long long i) { return !isnan(loss[i]); };
  double cnt = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x)
    if (take(i)) cnt += 1.0;
  const double c = block_sum_d(cnt, scratch);
  if (threadIdx.x == 0) s_cnt = c;
  __syncthreads();
  const unsigned long long count = (unsigned long long)s_cnt;
  const unsigned long long k = select ? (unsigned long long)((double)count * 0.25) : 0ull;
  if (count == 0 || k == 0) {
    if (threadIdx.x == 0) { sel[0] = -1e300; sel[1] = 0.0; sel[2] = (double)count; }
    return;
  }
  const uint32_t key = block_radix_select(loss, n, k - 1, take, hist, bc);   // largest dropped value
  const float thr = key_to_float(key);
  double nbelow = 0.0, neq = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x)
    if (take(i)) { if (loss[i] < thr) nbelow += 1.0; else if (loss[i] == thr) neq += 1.0; }
  const double nb = block_sum_d(nbelow, scratch);
  const double ne = block_sum_d(neq, scratch);
  if (threadIdx.x == 0) {
    const long long tie_drop = (long long)k - (long long)nb;     // >= 1
    long long cut = n;                                           // all equal values dropped
    if ((long long)ne > tie_drop) {                              // ambiguity: drop the first tie_drop of them
      long long seen = 0;
      for (long long i = 0; i < n; ++i)
        if (take(i) && loss[i] == thr && ++seen == tie_drop) { cut = i + 1; break; }
    }
    sel[0] = (double)thr; sel[1] = (double)cut; sel[2] = (double)(count - k);
  }
}

__global__ void __launch_bounds__(256) vnl_bwd_groups_kernel(const float* __restrict__ first,
                                                             const float* __restrict__ second,
                                                             const int* __restrict__ p1, const int* __restrict__ p2,
                                                             const int* __restrict__ p3, const float* __restrict__ loss,
                                                             const double* __restrict__ sel, int b_n, int n_pts, int h,
                                                             int w, float fx, float fy,
                                                             unsigned long long* __restrict__ acc) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)b_n * n_pts) return;
  const float l = loss[gid];
  if (isnan(l)) return;                                          // filtered out in the forward pass
  const float thr = (float)sel[0];
  if (sel[0] > -1e299 && (l < thr || (l == thr && gid < (long long)sel[1]))) return;   // dropped by the selection
  const float inv_k = (float)(1.0 / sel[2]);
  const int b = (int)(gid / n_pts), n = (int)(gid % n_pts);
  const int idx[3] = {p1[n], p2[n], p3[n]};
  const float u0 = (float)(w / 2), v0 = (float)(h / 2);
  Vec3 G[3], D[3];
  float uu[3], vv[3], sg[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int y = idx[j] / w, x = idx[j] - y * w;
    const float gd = first[(long long)b * h * w + idx[j]];
    const float dd = second[(long long)b * h * w + idx[j]];
    uu[j] = ((float)x - u0) / fx; vv[j] = ((float)y - v0) / fy;
    sg[j] = gd > 0.f ? 1.f : (gd < 0.f ? -1.f : 0.f);
    G[j] = {((float)x - u0) * fabsf(gd) / fx, ((float)y - v0) * fabsf(gd) / fy, gd};
    D[j] = {((float)x - u0) * fabsf(dd) / fx, ((float)y - v0) * fabsf(dd) / fy, dd};
  }
  const bool z0[3] = {D[0].z == 0.f, D[1].z == 0.f, D[2].z == 0.f};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (z0[0]) D[j].x = 0.0001f;
    if (z0[1]) D[j].y = 0.0001f;
    if (z0[2]) D[j].z = 0.0001f;
  }
  const Vec3 a = sub3(G[1], G[0]), bb = sub3(G[2], G[0]);
  const Vec3 gn = cross3(a, bb);
  const Vec3 dn = cross3(sub3(D[1], D[0]), sub3(D[2], D[0]));
  float gl = sqrtf(dot3(gn, gn)), dl = sqrtf(dot3(dn, dn));
  const bool gzero = gl == 0.f;
  if (gzero) gl += 0.01f;
  if (dl == 0.f) dl += 0.01f;
  auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
  const Vec3 r = {sgn(gn.x / gl - dn.x / dl), sgn(gn.y / gl - dn.y / dl), sgn(gn.z / gl - dn.z / dl)};
  const float rg = gzero ? 0.f : dot3(r, gn) / (gl * gl * gl);
  const Vec3 wv = {(r.x / gl - rg * gn.x) * inv_k, (r.y / gl - rg * gn.y) * inv_k, (r.z / gl - rg * gn.z) * inv_k};
  const Vec3 dA = cross3(bb, wv), dB = cross3(wv, a);
  const Vec3 dP[3] = {{-(dA.x + dB.x), -(dA.y + dB.y), -(dA.z + dB.z)}, dA, dB};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double g = (double)dP[j].x * (double)(uu[j] * sg[j]) + (double)dP[j].y * (double)(vv[j] * sg[j]) + (double)dP[j].z;
    const long long q = __double2ll_rn(g * kVnlFixScale);
    atomicAdd(acc + (long long)b * h * w + idx[j], (unsigned long long)q);
  }
}

__global__ void __launch_bounds__(256) vnl_bwd_finish_kernel(const unsigned long long* __restrict__ acc, long long n,
                                                             float upstream, float* __restrict__ grad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  grad[i] = (float)((double)(long long)acc[i] * (1.0 / kVnlFixScale)) * upstream;
}

}  // namespace odb

using namespace odb;

extern "C" int odb_make_valid_mask(const float* mask_float, uint8_t* mask_valid, int32_t b, int32_t h,
                                   int32_t w, int32_t pool, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!mask_float || !mask_valid || b < 1 || h < pool || w < pool || pool < 1)
    return fail(ODB_ERR_INVALID, "make_valid_mask: bad argument");
  const long long total = (long long)b * h * w;
  int blocks = (int)((total + 255) / 256);
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  make_valid_mask_kernel<<<blocks, 256, 0, stream>>>(mask_float, mask_valid, b, h, w, pool);
  count_launch();
  return check_launch("make_valid_mask");
}

extern "C" int64_t odb_midas_loss_workspace_bytes(int32_t b) {
  if (b < 1) return -1;
  return 256 + (int64_t)b * (4 * 4 + 8 * 8 + 8 * 8) + 256;
}

extern "C" int odb_midas_loss_fwd(const float* prediction, const float* target, const uint8_t* mask,
                                  int32_t b, int32_t h, int32_t w, float alpha, int32_t scales, float* out3,
                                  void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!prediction || !target || !mask || !out3 || !workspace || b < 1 || h < 1 || w < 1 || scales < 1 ||
      scales > 4 || (long long)h * w > 0x7fffffffLL)
    return fail(ODB_ERR_INVALID, "midas_loss_fwd: bad argument");
  if (workspace_bytes < odb_midas_loss_workspace_bytes(b) || (reinterpret_cast<uintptr_t>(workspace) & 255u))
    return fail(ODB_ERR_INVALID, "midas_loss_fwd: workspace too small or not 256-byte aligned");
  if (!(alpha > 0.f)) return fail(ODB_ERR_INVALID, "midas_loss_fwd: alpha must be > 0 (the reference leaves `total` undefined otherwise)");
  char* base = static_cast<char*>(workspace);
  MidasWs ws;
  ws.sums = reinterpret_cast<double*>(base);
  ws.grad = ws.sums + (size_t)b * 8;
  ws.med = reinterpret_cast<float*>(ws.grad + (size_t)b * 8);
  ws.scale = ws.med + 2 * (size_t)b;
  const int hw = h * w;
  midas_median_kernel<<<dim3(2, b), kLossThreads, 0, stream>>>(prediction, target, mask, ws, b, hw);
  count_launch();
  midas_sums_kernel<<<b, kLossThreads, 0, stream>>>(prediction, target, mask, ws, b, hw);
  count_launch();
  midas_grad_kernel<<<dim3(scales, b), kLossThreads, 0, stream>>>(prediction, target, mask, ws, h, w);
  count_launch();
  midas_finalize_kernel<<<1, 32, 0, stream>>>(ws, b, scales, alpha, out3);
  count_launch();
  return check_launch("midas_loss_fwd");
}

extern "C" int64_t odb_midas_loss_bwd_workspace_bytes(int32_t b) { return (int64_t)b * 16 * 8; }

extern "C" int odb_midas_loss_bwd(const float* prediction, const float* target, const uint8_t* mask, int32_t b,
                                  int32_t h, int32_t w, int32_t scales, float w_ssi, float w_reg,
                                  const void* fwd_workspace, void* bwd_workspace, float* gbuf, float* grad,
                                  void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!prediction || !target || !mask || !fwd_workspace || !bwd_workspace || !gbuf || !grad || b < 1 || h < 1 ||
      w < 1 || scales < 1 || scales > 4 || (long long)h * w > 0x7fffffffLL ||
      (reinterpret_cast<uintptr_t>(bwd_workspace) & 7u))
    return fail(ODB_ERR_INVALID, "midas_loss_bwd: bad argument");
  char* base = const_cast<char*>(static_cast<const char*>(fwd_workspace));
  MidasWs ws;
  ws.sums = reinterpret_cast<double*>(base);
  ws.grad = ws.sums + (size_t)b * 8;
  ws.med = reinterpret_cast<float*>(ws.grad + (size_t)b * 8);
  ws.scale = ws.med + 2 * (size_t)b;
  MidasBwdWs bw;
  bw.sc = static_cast<double*>(bwd_workspace);
  const int hw = h * w;
  midas_bwd_ssi_kernel<<<b, kLossThreads, 0, stream>>>(prediction, target, mask, ws, bw, b, hw);
  count_launch();
  midas_bwd_reg_kernel<<<b, kLossThreads, 0, stream>>>(prediction, target, mask, ws, bw, b, h, w, scales, gbuf);
  count_launch();
  int gx = (hw + 255) / 256;
  if (gx > 64) gx = 64;
  midas_bwd_final_kernel<<<dim3(gx, b), 256, 0, stream>>>(prediction, target, mask, ws, bw, b, hw, w_ssi, w_reg, gbuf,
                                                         grad);
  count_launch();
  return check_launch("midas_loss_bwd");
}

extern "C" int odb_normal_loss_bwd(const float* prediction, const float* target, const uint8_t* mask_valid, int32_t b,
                                   int32_t h, int32_t w, int32_t clamp_prediction, float w_l1, float w_cos,
                                   const double* fwd_workspace, float* grad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!prediction || !target || !mask_valid || !fwd_workspace || !grad || b < 1 || h < 1 || w < 1 ||
      (long long)h * w > 0x7fffffffLL / 3 || b > 65535)
    return fail(ODB_ERR_INVALID, "normal_loss_bwd: bad argument");
  int gx = (h * w + 255) / 256;
  if (gx > 64) gx = 64;
  normal_loss_bwd_kernel<<<dim3(gx, b), 256, 0, stream>>>(prediction, target, mask_valid, fwd_workspace, b, h * w,
                                                         clamp_prediction, w_l1, w_cos, grad);
  count_launch();
  return check_launch("normal_loss_bwd");
}

extern "C" int odb_vnl_loss_bwd(const float* first, const float* second, const int32_t* p1, const int32_t* p2,
                                const int32_t* p3, int32_t n_points, int32_t b, int32_t h, int32_t w, float fx, float fy,
                                int32_t select, const float* group_loss, float upstream, void* acc, double* sel4,
                                float* grad, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!first || !second || !p1 || !p2 || !p3 || !group_loss || !acc || !sel4 || !grad || n_points < 1 || b < 1 ||
      h < 1 || w < 1 || fx == 0.f || fy == 0.f || (reinterpret_cast<uintptr_t>(acc) & 7u))
    return fail(ODB_ERR_INVALID, "vnl_loss_bwd: bad argument");
  const long long total = (long long)b * n_points;
  const long long px = (long long)b * h * w;
  cudaError_t e = cudaMemsetAsync(acc, 0, (size_t)px * 8, stream);
  if (e != cudaSuccess) return fail_cuda(e, "vnl_loss_bwd: memset");
  vnl_bwd_select_kernel<<<1, kLossThreads, 0, stream>>>(group_loss, total, select, sel4);
  count_launch();
  vnl_bwd_groups_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      first, second, p1, p2, p3, group_loss, sel4, b, n_points, h, w, fx, fy, static_cast<unsigned long long*>(acc));
  count_launch();
  vnl_bwd_finish_kernel<<<(unsigned)((px + 255) / 256), 256, 0, stream>>>(static_cast<const unsigned long long*>(acc), px,
                                                                        upstream, grad);
  count_launch();
  return check_launch("vnl_loss_bwd");
}

extern "C" int odb_normal_loss_fwd(const float* prediction, const float* target, const uint8_t* mask_valid,
                                   int32_t b, int32_t h, int32_t w, int32_t clamp_prediction, float* out3,
                                   double* workspace, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!prediction || !target || !mask_valid || !out3 || !workspace || b < 1 || h < 1 || w < 1 ||
      (long long)h * w > 0x7fffffffLL / 3)
    return fail(ODB_ERR_INVALID, "normal_loss_fwd: bad argument");
  normal_loss_partial_kernel<<<b, kLossThreads, 0, stream>>>(prediction, target, mask_valid, h * w,
                                                            clamp_prediction, workspace);
  count_launch();
  normal_loss_finalize_kernel<<<1, 32, 0, stream>>>(workspace, b, out3);
  count_launch();
  return check_launch("normal_loss_fwd");
}

extern "C" int odb_vnl_loss_fwd(const float* first, const float* second, const int32_t* p1, const int32_t* p2,
                                const int32_t* p3, int32_t n_points, int32_t b, int32_t h, int32_t w, float fx,
                                float fy, float delta_z, int32_t select, float* out1, float* group_loss,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!first || !second || !p1 || !p2 || !p3 || !out1 || !group_loss || n_points < 1 || b < 1 || h < 1 ||
      w < 1 || fx == 0.f || fy == 0.f)
    return fail(ODB_ERR_INVALID, "vnl_loss_fwd: bad argument");
  const long long total = (long long)b * n_points;
  vnl_groups_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(first, second, p1, p2, p3, group_loss, b,
                                                                         n_points, h, w, fx, fy, delta_z);
  count_launch();
  vnl_reduce_kernel<<<1, kLossThreads, 0, stream>>>(group_loss, total, select, out1);
  count_launch();
  return check_launch("vnl_loss_fwd");
}
