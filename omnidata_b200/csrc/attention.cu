// Fused multi-head attention for the ViT-B/16 blocks of DPT-Hybrid (577 tokens, 12 heads, d = 64):
//   out = softmax(q k^T * scale) v       (timm Attention.forward; block loop at M/vit.py:150-151)
// One CTA per (image, head).  K and V of the whole sequence (<= 640 keys) stay resident in shared
// memory (160 KiB, XOR-swizzled 128-byte rows); each warp owns 16 query rows at a time and runs a
// flash-style online softmax over 64-key chunks, so S and P never leave registers.
// Round-1 tensor path: warp-level mma.sync m16n8k16 bf16 (fp32 accumulate).  The tcgen05/TMEM
// version of this kernel is listed in DESIGN.md as the next step for this op.
#include "common.cuh"
#include "host_util.h"
#include "../../include/omnidata_b200.h"

namespace odb {

constexpr int kHeadDim = 64;
constexpr int kMaxKeys = 640;
constexpr int kAttnThreads = 256;

ODB_DEVINL void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
ODB_DEVINL void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                  uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
ODB_DEVINL void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
      "{%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
ODB_DEVINL float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
ODB_DEVINL void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int tokens, int heads,
                 float scale_log2e) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sK = smem_u32(smem);
  const uint32_t sV = sK + kMaxKeys * 128;
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = heads * kHeadDim;         // 768
  const long long row_stride = 3LL * C;   // 2304
  const bf16* base = qkv + (long long)b * tokens * row_stride + h * kHeadDim;
  const int nchunks = (tokens + 63) / 64;
  const int keys_pad = nchunks * 64;

  // ---- stage K and V (zero rows beyond the sequence)
  for (int i = threadIdx.x; i < keys_pad * 8; i += kAttnThreads) {
    const int key = i >> 3, ch = i & 7;
    const uint32_t off = key * 128 + ((ch ^ (key & 7)) << 4);
    if (key < tokens) {
      const bf16* src = base + (long long)key * row_stride + ch * 8;
      cp_async16(sK + off, src + C);
      cp_async16(sV + off, src + 2 * C);
    } else {
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sK + off), "r"(0) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sV + off), "r"(0) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int qr = lane >> 2;        // row within the 8-row half
  const int qc = (lane & 3) * 2;   // column pair within an 8-wide tile
  const int lmat = lane >> 3, lrow = lane & 7;
  const int groups = (tokens + 15) / 16;

  for (int g = warp; g < groups; g += kAttnThreads / 32) {
    const int r0 = g * 16 + qr, r1 = r0 + 8;
    const int r0c = min(r0, tokens - 1), r1c = min(r1, tokens - 1);
    // ---- Q fragments straight from global (each row is read exactly once)
    uint32_t qa[4][4];
    {
      const bf16* q0 = base + (long long)r0c * row_stride;
      const bf16* q1 = base + (long long)r1c * row_stride;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        qa[kk][0] = __ldg(reinterpret_cast<const uint32_t*>(q0 + kk * 16 + qc));
        qa[kk][1] = __ldg(reinterpret_cast<const uint32_t*>(q1 + kk * 16 + qc));
        qa[kk][2] = __ldg(reinterpret_cast<const uint32_t*>(q0 + kk * 16 + 8 + qc));
        qa[kk][3] = __ldg(reinterpret_cast<const uint32_t*>(q1 + kk * 16 + 8 + qc));
      }
    }
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int ck = 0; ck < nchunks; ++ck) {
      const int key0 = ck * 64;
      float s[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
      // ---- S = Q K^T for 64 keys
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = key0 + j * 8 + lrow;
        const uint32_t rowaddr = sK + key * 128;
        uint32_t kb[8];
        ldmatrix_x4(rowaddr + (((lmat) ^ (key & 7)) << 4), kb[0], kb[1], kb[2], kb[3]);
        ldmatrix_x4(rowaddr + (((4 + lmat) ^ (key & 7)) << 4), kb[4], kb[5], kb[6], kb[7]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) mma_bf16_16816(s[j], qa[kk], kb[2 * kk], kb[2 * kk + 1]);
      }
      // ---- scale, mask, online softmax
      float cmax0 = -INFINITY, cmax1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kcol = key0 + j * 8 + qc;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = s[j][e] * scale_log2e;
          if (kcol + (e & 1) >= tokens) v = -INFINITY;
          s[j][e] = v;
        }
        cmax0 = fmaxf(cmax0, fmaxf(s[j][0], s[j][1]));
        cmax1 = fmaxf(cmax1, fmaxf(s[j][2], s[j][3]));
      }
      cmax0 = fmaxf(cmax0, __shfl_xor_sync(0xffffffffu, cmax0, 1));
      cmax0 = fmaxf(cmax0, __shfl_xor_sync(0xffffffffu, cmax0, 2));
      cmax1 = fmaxf(cmax1, __shfl_xor_sync(0xffffffffu, cmax1, 1));
      cmax1 = fmaxf(cmax1, __shfl_xor_sync(0xffffffffu, cmax1, 2));
      const float mn0 = fmaxf(m0, cmax0), mn1 = fmaxf(m1, cmax1);
      const float alpha0 = fast_exp2(m0 - mn0), alpha1 = fast_exp2(m1 - mn1);
      m0 = mn0; m1 = mn1;
      float ps0 = 0.f, ps1 = 0.f;
      uint32_t pa[4][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float p0 = fast_exp2(s[j][0] - mn0), p1 = fast_exp2(s[j][1] - mn0);
        const float p2 = fast_exp2(s[j][2] - mn1), p3 = fast_exp2(s[j][3] - mn1);
        ps0 += p0 + p1;
        ps1 += p2 + p3;
        pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p0, p1);
        pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p2, p3);
      }
      l0 = l0 * alpha0 + ps0;
      l1 = l1 * alpha1 + ps1;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j][0] *= alpha0; o[j][1] *= alpha0; o[j][2] *= alpha1; o[j][3] *= alpha1;
      }
      // ---- O += P V
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int key = key0 + t * 16 + (lmat & 1) * 8 + lrow;
        const uint32_t rowaddr = sV + key * 128;
#pragma unroll
        for (int jp = 0; jp < 4; ++jp) {
          uint32_t v0, v1, v2, v3;
          ldmatrix_x4_trans(rowaddr + (((2 * jp + (lmat >> 1)) ^ (key & 7)) << 4), v0, v1, v2, v3);
          mma_bf16_16816(o[2 * jp], pa[t], v0, v1);
          mma_bf16_16816(o[2 * jp + 1], pa[t], v2, v3);
        }
      }
    }
    // ---- finalize
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    bf16* ob = out + (long long)b * tokens * C + h * kHeadDim;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (r0 < tokens)
        *reinterpret_cast<uint32_t*>(ob + (long long)r0 * C + j * 8 + qc) =
            pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
      if (r1 < tokens)
        *reinterpret_cast<uint32_t*>(ob + (long long)r1 * C + j * 8 + qc) =
            pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
    }
  }
}

}  // namespace odb

using namespace odb;

extern "C" int odb_attention_mma(const void* qkv, void* out, int32_t b, int32_t tokens, int32_t heads,
                             float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!qkv || !out || b < 1 || heads < 1 || tokens < 1)
    return fail(ODB_ERR_INVALID, "attention: bad argument");
  if (tokens > kMaxKeys) return fail(ODB_ERR_UNSUPPORTED, "attention: at most 640 tokens");
  const int smem = 2 * kMaxKeys * 128;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail_cuda(e, "attention: cudaFuncSetAttribute");
    configured = true;
  }
  dim3 grid(heads, b);
  attention_kernel<<<grid, kAttnThreads, smem, stream>>>(static_cast<const bf16*>(qkv),
                                                         static_cast<bf16*>(out), tokens, heads,
                                                         scale * 1.4426950408889634f);
  count_launch();
  return check_launch("attention");
}
