// tcgen05 / TMEM fused attention, two query tiles in flight ("ping-pong").
//
// Same arithmetic as attention_tc.cu (exact two-pass softmax, P rounded to bf16 for PV, fp32 row sum
// of the unrounded P) but organised to hide the MMA <-> softmax hand-off latency that bounded the
// first kernel (ncu: tensor pipe 24 % active, stalls on the per-block mbarriers): the CTA runs TWO
// independent groups, each with its own producer warp, MMA-issuing warp, four softmax warps, Q slot,
// P slot and TMEM columns, working on alternate 128-query tiles of the same (image, head) unit against
// the shared resident K / V.  While one group waits for a tensor-core result the other one computes.
//   warps 0/2   TMA producers (group 0 also loads K and V of the unit)
//   warps 1/3   tcgen05.mma issuers:  S = Q K_j^T (M=128, N=64, K=64),  O += P_j V_j (M=128, N=64, K=64)
//   warps 4-7 / 8-11  softmax + epilogue of group 0 / 1: one thread per query row (whole 64-key block)
// TMEM per group: S double-buffered (2 x 64 columns) + O (64 columns) = 192 columns.
#include "common.cuh"
#include "host_util.h"
#include "../../include/omnidata_b200.h"

namespace odb {

constexpr int kT2Rows = 128;               // queries per tile
constexpr int kT2Keys = 64;                // keys per S / P block
constexpr int kT2MaxKeyRows = 640;         // resident K / V rows
constexpr int kT2TileBytes = kT2Rows * 128;
constexpr int kT2Threads = 384;

constexpr int kT2OffK = 0;
constexpr int kT2OffV = kT2MaxKeyRows * 128;
constexpr int kT2OffQ = 2 * kT2MaxKeyRows * 128;          // 2 slots (one per group)
constexpr int kT2OffP = kT2OffQ + 2 * kT2TileBytes;       // 2 slots (one per group), 128 x 64 bf16 each
constexpr int kT2OffBar = kT2OffP + 2 * kT2TileBytes;
constexpr int kT2SmemBytes = kT2OffBar + 256 + 1024;
static_assert(kT2SmemBytes <= 232448, "attention smem plan exceeds 227 KiB");

struct AttnT2Params {
  CUtensorMap kv_map;   // dims {64 d, 3*heads, tokens, batch}; box {64, 1, 128, 1}
  bf16* out;
  int tokens, heads, batch;
  float scale_log2e;
};

ODB_DEVINL float ex2_t2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kT2Threads, 1) attention_tc2_kernel(const __grid_constant__ AttnT2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar0 = sbase + kT2OffBar;
  // shared by both groups
  const uint32_t kv_full = bar0, kv_empty = bar0 + 8;
  // per group g: q_full, q_empty, s_full[2], s_empty[2], p_full, p_empty, o_full, o_empty  (10 barriers)
  auto gbar = [&](int g, int i) { return bar0 + 16u + 8u * (g * 10 + i); };
  const uint32_t tmem_slot = bar0 + 16u + 8u * 20;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + kT2OffBar + 16 + 8 * 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.tokens + kT2Rows - 1) / kT2Rows;
  const int n_kb = (p.tokens + kT2Keys - 1) / kT2Keys;     // 64-key blocks
  const int n_kv_tiles = (p.tokens + 127) / 128;           // 128-row TMA boxes of K / V
  const int units = p.batch * p.heads;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    mbar_init(kv_empty, 2);                                 // both groups' MMA warps release the unit
    for (int g = 0; g < 2; ++g) {
      mbar_init(gbar(g, 0), 1); mbar_init(gbar(g, 1), 1);   // q_full, q_empty
      mbar_init(gbar(g, 2), 1); mbar_init(gbar(g, 3), 1);   // s_full[2]
      mbar_init(gbar(g, 4), 4); mbar_init(gbar(g, 5), 4);   // s_empty[2] (4 softmax warps)
      mbar_init(gbar(g, 6), 128); mbar_init(gbar(g, 7), 1); // p_full, p_empty
      mbar_init(gbar(g, 8), 1); mbar_init(gbar(g, 9), 4);   // o_full, o_empty
    }
    mbar_fence_init();
    tma_prefetch_desc(&p.kv_map);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  grid_dep_wait();
  grid_dep_launch();

  // role decode: group of this warp
  const int grp = (warp < 4) ? (warp >> 1) : ((warp - 4) >> 2);
  const uint32_t q_full = gbar(grp, 0), q_empty = gbar(grp, 1);
  auto s_full = [&](int i) { return gbar(grp, 2 + i); };
  auto s_empty = [&](int i) { return gbar(grp, 4 + i); };
  const uint32_t p_full = gbar(grp, 6), p_empty = gbar(grp, 7), o_full = gbar(grp, 8), o_empty = gbar(grp, 9);
  const uint32_t q_smem = sbase + kT2OffQ + grp * kT2TileBytes;
  const uint32_t p_smem = sbase + kT2OffP + grp * kT2TileBytes;
  const uint32_t t_grp = tmem_base + grp * 192;             // S0 +0, S1 +64, O +128
  // tiles of unit number u_iter owned by this group: those with (tile + u_iter) odd/even alternating,
  // so that over two consecutive units both groups process the same number of tiles
  auto first_tile = [&](uint32_t u_iter) { return (grp + static_cast<int>(u_iter)) & 1; };

  if (warp == 0 || warp == 2) {
    // ------------------------------------------------------------------ TMA producers
    if (lane == 0) {
      uint32_t u_iter = 0, t_iter = 0;
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++u_iter) {
        const int b = unit / p.heads, h = unit % p.heads;
        if (grp == 0) {
          mbar_wait(kv_empty, (u_iter & 1u) ^ 1u);
          mbar_expect_tx(kv_full, 2u * n_kv_tiles * kT2TileBytes);
          for (int j = 0; j < n_kv_tiles; ++j) {
            tma_load_4d(sbase + kT2OffK + j * kT2TileBytes, &p.kv_map, kv_full, 0, p.heads + h, j * 128, b);
            tma_load_4d(sbase + kT2OffV + j * kT2TileBytes, &p.kv_map, kv_full, 0, 2 * p.heads + h, j * 128, b);
          }
        }
        for (int qt = first_tile(u_iter); qt < n_tiles; qt += 2, ++t_iter) {
          mbar_wait(q_empty, (t_iter & 1u) ^ 1u);
          mbar_expect_tx(q_full, kT2TileBytes);
          tma_load_4d(q_smem, &p.kv_map, q_full, 0, h, qt * kT2Rows, b);
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ------------------------------------------------------------------ MMA issuers
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(kT2Rows, kT2Keys);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(kT2Rows, 64) | (1u << 16);   // B (= V) is MN-major
      uint32_t u_iter = 0, t_iter = 0, sb_iter = 0, p_iter = 0;
      const uint64_t qdesc = umma_desc_sw128(q_smem);
      auto issue_s = [&](int j) {
        const uint32_t sbuf = sb_iter & 1u;
        mbar_wait(s_empty(sbuf), ((sb_iter >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint64_t kdesc = umma_desc_sw128(sbase + kT2OffK + j * (kT2Keys * 128));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(t_grp + sbuf * kT2Keys, qdesc + 2u * k, kdesc + 2u * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full(sbuf));
        ++sb_iter;
      };
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++u_iter) {
        mbar_wait(kv_full, u_iter & 1u);
        tc_fence_after();
        for (int qt = first_tile(u_iter); qt < n_tiles; qt += 2, ++t_iter) {
          mbar_wait(q_full, t_iter & 1u);
          tc_fence_after();
          int next_s = 0;
          auto pump = [&](int upto) {
            const int lim = upto < 2 * n_kb ? upto : 2 * n_kb;
            for (; next_s < lim; ++next_s) issue_s(next_s < n_kb ? next_s : next_s - n_kb);
          };
          pump(n_kb + 1);
          for (int j = 0; j < n_kb; ++j) {
            pump(n_kb + j + 2);
            mbar_wait(p_full, p_iter & 1u);
            tc_fence_after();
            if (j == 0) {
              mbar_wait(o_empty, (t_iter & 1u) ^ 1u);
              tc_fence_after();
            }
            const uint64_t pdesc = umma_desc_sw128(p_smem);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t vdesc = umma_desc_sw128(sbase + kT2OffV + (j * kT2Keys + kk * 16) * 128);
              umma_bf16_ss(t_grp + 128, pdesc + 2u * kk, vdesc, idesc_pv, (j | kk) != 0 ? 1u : 0u);
            }
            umma_commit(p_empty);
            ++p_iter;
          }
          umma_commit(o_full);
          umma_commit(q_empty);
        }
        umma_commit(kv_empty);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue (one row per thread)
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = t_grp + (static_cast<uint32_t>(quad * 32) << 16);
    const float c = p.scale_log2e;
    uint32_t u_iter = 0, t_iter = 0, sb_iter = 0, p_iter = 0;
    for (int unit = blockIdx.x; unit < units; unit += gridDim.x, ++u_iter) {
      const int b = unit / p.heads, h = unit % p.heads;
      for (int qt = first_tile(u_iter); qt < n_tiles; qt += 2, ++t_iter) {
        // ---- pass 1: row maximum
        float mx = -INFINITY;
        for (int j = 0; j < n_kb; ++j, ++sb_iter) {
          const uint32_t sbuf = sb_iter & 1u;
          mbar_wait(s_full(sbuf), (sb_iter >> 1) & 1u);
          tc_fence_after();
          uint32_t r[64];
          tmem_ld_32x32(t_lane + sbuf * kT2Keys, r);
          tmem_ld_32x32(t_lane + sbuf * kT2Keys + 32, r + 32);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty(sbuf));
          const int key0 = j * kT2Keys;
          if (key0 + kT2Keys <= p.tokens) {
#pragma unroll
            for (int i = 0; i < 64; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 64; ++i)
              if (key0 + i < p.tokens) mx = fmaxf(mx, __uint_as_float(r[i]));
          }
        }
        const float mc = mx * c;
        // ---- pass 2
        float l = 0.f;
        for (int j = 0; j < n_kb; ++j, ++sb_iter, ++p_iter) {
          const uint32_t sbuf = sb_iter & 1u;
          mbar_wait(s_full(sbuf), (sb_iter >> 1) & 1u);
          tc_fence_after();
          uint32_t r[64];
          tmem_ld_32x32(t_lane + sbuf * kT2Keys, r);
          tmem_ld_32x32(t_lane + sbuf * kT2Keys + 32, r + 32);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_empty(sbuf));
          const int key0 = j * kT2Keys;
          uint32_t packed[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float p0 = ex2_t2(fmaf(__uint_as_float(r[2 * i]), c, -mc));
            float p1 = ex2_t2(fmaf(__uint_as_float(r[2 * i + 1]), c, -mc));
            if (key0 + 2 * i >= p.tokens) p0 = 0.f;
            if (key0 + 2 * i + 1 >= p.tokens) p1 = 0.f;
            l += p0 + p1;
            packed[i] = pack_bf16x2(p0, p1);
          }
          mbar_wait(p_empty, (p_iter & 1u) ^ 1u);
          const uint32_t prow = p_smem + static_cast<uint32_t>(row) * 128u;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const uint32_t addr = prow + (static_cast<uint32_t>(jj ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(packed[4 * jj]),
                         "r"(packed[4 * jj + 1]), "r"(packed[4 * jj + 2]), "r"(packed[4 * jj + 3])
                         : "memory");
          }
          fence_proxy_async_smem();
          mbar_arrive(p_full);
        }
        const float inv = 1.0f / l;
        // ---- epilogue
        mbar_wait(o_full, t_iter & 1u);
        tc_fence_after();
        uint32_t o[64];
        tmem_ld_32x32(t_lane + 128, o);
        tmem_ld_32x32(t_lane + 160, o + 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);
        const int qrow = qt * kT2Rows + row;
        if (qrow < p.tokens) {
          bf16* dst = p.out + ((long long)b * p.tokens + qrow) * (p.heads * 64) + h * 64;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o[8 * jj + 0]) * inv, __uint_as_float(o[8 * jj + 1]) * inv);
            v.y = pack_bf16x2(__uint_as_float(o[8 * jj + 2]) * inv, __uint_as_float(o[8 * jj + 3]) * inv);
            v.z = pack_bf16x2(__uint_as_float(o[8 * jj + 4]) * inv, __uint_as_float(o[8 * jj + 5]) * inv);
            v.w = pack_bf16x2(__uint_as_float(o[8 * jj + 6]) * inv, __uint_as_float(o[8 * jj + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + jj * 8) = v;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace odb

using namespace odb;

extern "C" int odb_attention_pp(const void* qkv, void* out, int32_t b, int32_t tokens, int32_t heads,
                                float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!qkv || !out || b < 1 || heads < 1 || tokens < 1)
    return fail(ODB_ERR_INVALID, "attention: bad argument");
  if (tokens > kT2MaxKeyRows) return fail(ODB_ERR_UNSUPPORTED, "attention: at most 640 tokens");
  if (reinterpret_cast<uintptr_t>(qkv) & 15u) return fail(ODB_ERR_INVALID, "attention: qkv must be 16-byte aligned");
  AttnT2Params p;
  memset(&p, 0, sizeof(p));
  {
    cuuint64_t dims[4] = {64, (cuuint64_t)(3 * heads), (cuuint64_t)tokens, (cuuint64_t)b};
    cuuint64_t strides[3] = {128, (cuuint64_t)(3 * heads) * 128, (cuuint64_t)tokens * (3 * heads) * 128};
    cuuint32_t box[4] = {64, 1, 128, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    int rc = encode_tiled(&p.kv_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(qkv), dims,
                          strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  p.out = static_cast<bf16*>(out);
  p.tokens = tokens; p.heads = heads; p.batch = b;
  p.scale_log2e = scale * 1.4426950408889634f;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kT2SmemBytes);
    if (e != cudaSuccess) return fail_cuda(e, "attention: cudaFuncSetAttribute");
    configured = true;
  }
  const int units = b * heads;
  const int grid = units < num_sms() ? units : num_sms();
  cudaError_t le = launch_pdl(attention_tc2_kernel, dim3(grid), dim3(kT2Threads), kT2SmemBytes, stream, p);
  count_launch();
  if (le != cudaSuccess) return fail_cuda(le, "attention_pp: launch");
  return check_launch("attention_pp");
}
