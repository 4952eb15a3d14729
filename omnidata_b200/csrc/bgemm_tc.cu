// tcgen05 batched GEMM with selectable operand majors — the backward contractions of the train step (bf16):
//   * weight gradients  dW[n][t*C + c] = sum_pixels dY[pixel][n] * X_t[pixel + offset_t][c]
//     Both operands are MN-major: the 128-byte-swizzled TMA box {64 channels, 64 pixels} that the forward kernel
//     consumes as a K-major A tile (M = pixels, K = channels) IS the canonical MN-major UMMA operand of the
//     transposed product (K = pixels, MN = channels) — no transposes, no im2col, tap shifts and zero padding by the
//     TMA unit exactly as in the forward.  The pixel range is split across CTAs (fp32 partials, ordered reduction).
//   * attention backward per (image, head): P = exp2(QK^T c - lse) and dS = P o (dO V^T - D) * scale as K-major
//     GEMMs with fused epilogues (P / dS materialised once, bf16), then dQ = dS K (B MN-major), dK = dS^T Q and
//     dV = P^T dO (A and B MN-major).
// One persistent warp-specialised kernel: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (TMEM -> registers
// -> global, one accumulator row per thread), 4-stage operand ring, double-buffered TMEM accumulators.
// UMMA shared-memory descriptors (cute/atom/mma_traits_sm100.hpp, canonical layouts):
//   K-major  SW128: rows of 64 K-elements (128 B), 8-row groups 1024 B apart (SBO);
//   MN-major SW128: ((8,8,m),(8,k)) : ((1,8,LBO),(64,SBO)) in elements — 64 contiguous MN elements per 128-byte row,
//                   rows = K index, 8-row K groups SBO = 1024 B apart, 64-element MN atoms LBO apart (= one 64-row
//                   sub-box here: 8192 B).
#include "common.cuh"
#include "host_util.h"
#include "../../include/omnidata_b200.h"

namespace odb {

constexpr int kBgStages = 4;                // M = 128 tiles: 4 stages of 48 KiB; M = 256 tiles: 3 stages of 64 KiB
constexpr int kBgAStage = 128 * 128;        // 16 KiB: 128 x 64 bf16 either way
constexpr int kBgBStage = 256 * 128;        // 32 KiB: up to 256 x 64 bf16
constexpr int kBgSubBytes = 64 * 128;       // MN-major sub-box: 64 K-rows x 64 MN elements
constexpr int kBgMaxThreads = 320;          // warp 0 TMA, warp 1 MMA, 4 or 8 epilogue warps
constexpr int kBgRing = kBgStages * (kBgAStage + kBgBStage);      // 192 KiB operand ring
constexpr int kBgStagePitch = 144;                               // 128-byte row + 16: conflict-free row / piece access
constexpr int kBgWarpStage = 32 * kBgStagePitch;                  // one 32-row x 128-byte staging tile per epilogue warp
static int bg_smem_bytes(int stages, int mt2, int epi_warps) {
  return stages * (mt2 * kBgAStage + kBgBStage) + 256 + epi_warps * kBgWarpStage + 1024;
}
constexpr int kBgSmemMax = 227 * 1024;

enum : int { BG_EPI_F32 = 0, BG_EPI_BF16 = 1, BG_EPI_P = 2, BG_EPI_DS = 3 };

struct BgOperand {
  CUtensorMap map[ODB_MAX_VIEWS];
  int off[4], mul_mn[4], mul_k1[4], mul_k2[4], mul_k3[4], mul_z1[4], mul_z2[4];
  int mn_major;          // 0: K-major (one box of `rows` x 64 K), 1: MN-major (nsub boxes of 64 K-rows x 64 MN)
  int nsub;              // boxes per stage
  int sub_dim;           // tensor dim advanced by 64 per sub-box
  unsigned sub_bytes;    // smem bytes per box
};

struct BgParams {
  BgOperand a, b;
  int nk1, nk2, nk3;
  int MT, NT, Z1, Z2;
  int split_mode, ksteps_per_split;
  int stages;            // operand ring depth (3 or 4)
  int epi_warps;         // 4: one epilogue warp per TMEM lane quadrant; 8: two per quadrant, each owning half of the columns
                         //    (short-K GEMMs with heavy epilogues: the attention P / dS products)
  int mt2;               // 1: 128-row M tiles, double-buffered accumulator; 2: 256-row M tiles (two 128-row UMMAs sharing the
                         //    B tile: half the B traffic per flop), one accumulator set — for long K loops (weight gradients)
  int bn;
  int tap_mode;
  int8_t tap_view[ODB_MAX_TAPS], tap_dx[ODB_MAX_TAPS], tap_dy[ODB_MAX_TAPS];
  int epi;
  void* out;
  long long o_row, o_z1, o_z2;
  int m_valid, n_valid, n_store;
  const float* rowvec;           // EPI_P: lse, EPI_DS: D;  index z2 * rv_z2 + z1 * rv_z1 + row
  long long rv_z1, rv_z2;
  const bf16* pmat;              // EPI_DS: P, addressed like out
  float c1;
};

ODB_DEVINL uint64_t bg_desc(uint32_t addr, uint32_t lbo_bytes) {
  return static_cast<uint64_t>((addr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         (64ull << 32) | (1ull << 46) | (2ull << 61);
}

ODB_DEVINL void bg_issue_loads(const BgOperand& o, const CUtensorMap* map, uint32_t smem_dst, uint32_t bar, int mn, int k1,
                               int k2, int k3, int z1, int z2, int dx, int dy) {
  int c[4];
#pragma unroll
  for (int d = 0; d < 4; ++d)
    c[d] = o.off[d] + mn * o.mul_mn[d] + k1 * o.mul_k1[d] + k2 * o.mul_k2[d] + k3 * o.mul_k3[d] + z1 * o.mul_z1[d] +
           z2 * o.mul_z2[d];
  c[1] += dx;
  c[2] += dy;
  for (int s = 0; s < o.nsub; ++s) {
    int cc[4] = {c[0], c[1], c[2], c[3]};
    cc[o.sub_dim] += 64 * s;
    tma_load_4d(smem_dst + s * o.sub_bytes, map, bar, cc[0], cc[1], cc[2], cc[3]);
  }
}

__global__ void __launch_bounds__(kBgMaxThreads, 1) bgemm_kernel(const __grid_constant__ BgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stages = p.stages;
  const uint32_t a_stage = static_cast<uint32_t>(p.mt2) * kBgAStage;
  const uint32_t stage_stride = a_stage + kBgBStage;          // stage s: A at s * stride, B right behind it
  const uint32_t a_base = sbase, b_base = sbase + a_stage;
  const uint32_t bar0 = sbase + static_cast<uint32_t>(stages) * stage_stride;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kBgStages + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * kBgStages + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * kBgStages + 2 + a); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * kBgStages + 4);
  const uint32_t stage_base = bar0 + 256u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kBgStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), p.epi_warps); }
    mbar_fence_init();
    tma_prefetch_desc(&p.a.map[0]);
    tma_prefetch_desc(&p.b.map[0]);
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int total_units = p.MT * p.NT * p.Z1 * p.Z2;
  const int ksteps_all = p.nk1 * p.nk2 * p.nk3;
  auto decode = [&](int unit, int& mt, int& nt, int& z1, int& z2) {
    mt = unit % p.MT; unit /= p.MT;
    nt = unit % p.NT; unit /= p.NT;
    z1 = unit % p.Z1;
    z2 = unit / p.Z1;
  };
  auto krange = [&](int z2, int& kb, int& ke) {
    if (p.split_mode) { kb = z2 * p.ksteps_per_split; ke = min(ksteps_all, kb + p.ksteps_per_split); }
    else { kb = 0; ke = ksteps_all; }
  };
  const uint32_t a_bytes = static_cast<uint32_t>(p.a.nsub) * p.a.sub_bytes;
  const uint32_t b_bytes = static_cast<uint32_t>(p.b.nsub) * p.b.sub_bytes;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
        int mt, nt, z1, z2, kb, ke;
        decode(unit, mt, nt, z1, z2);
        krange(z2, kb, ke);
        const int zz2 = p.split_mode ? 0 : z2;
        const CUtensorMap* bmap = &p.b.map[p.tap_mode ? p.tap_view[z1] : 0];
        const int dx = p.tap_mode ? p.tap_dx[z1] : 0, dy = p.tap_mode ? p.tap_dy[z1] : 0;
        const int zb1 = p.tap_mode ? 0 : z1;
        for (int ks = kb; ks < ke; ++ks) {
          const int k1 = ks % p.nk1, k2 = (ks / p.nk1) % p.nk2, k3 = ks / (p.nk1 * p.nk2);
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), a_bytes + b_bytes);
          bg_issue_loads(p.a, &p.a.map[0], a_base + stage * stage_stride, full_bar(stage), mt * 128 * p.mt2, k1, k2, k3, zb1, zz2, 0, 0);
          bg_issue_loads(p.b, bmap, b_base + stage * stage_stride, full_bar(stage), nt * p.bn, k1, k2, k3, zb1, zz2, dx, dy);
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.bn) | (p.a.mn_major ? (1u << 15) : 0u) | (p.b.mn_major ? (1u << 16) : 0u);
      const uint32_t a_kstep = p.a.mn_major ? 128u : 2u;     // descriptor units (16 B) per UMMA_K = 16
      const uint32_t b_kstep = p.b.mn_major ? 128u : 2u;
      int stage = 0; uint32_t phase = 0, iter = 0;
      for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++iter) {
        int mt, nt, z1, z2, kb, ke;
        decode(unit, mt, nt, z1, z2);
        krange(z2, kb, ke);
        const uint32_t acc = p.mt2 == 2 ? 0u : (iter & 1u), acc_phase = p.mt2 == 2 ? (iter & 1u) : ((iter >> 1) & 1u);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256u;
        for (int ks = kb; ks < ke; ++ks) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t adesc = bg_desc(a_base + stage * stage_stride, p.a.mn_major ? kBgSubBytes : 16u);
          const uint64_t bdesc = bg_desc(b_base + stage * stage_stride, p.b.mn_major ? kBgSubBytes : 16u);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t accum = (ks > kb || k > 0) ? 1u : 0u;
            umma_bf16_ss(d_tmem, adesc + a_kstep * k, bdesc + b_kstep * k, idesc, accum);
            // second 128 rows of a 256-row tile: the A sub-tile one 16 KiB block further, accumulator columns [256, 512)
            if (p.mt2 == 2) umma_bf16_ss(d_tmem + 256u, adesc + (kBgAStage >> 4) + a_kstep * k, bdesc + b_kstep * k, idesc, accum);
          }
          umma_commit(empty_bar(stage));
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else {
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    uint32_t iter = 0;
    for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x, ++iter) {
      int mt, nt, z1, z2;
      decode(unit, mt, nt, z1, z2);
      const uint32_t acc = p.mt2 == 2 ? 0u : (iter & 1u), acc_phase = p.mt2 == 2 ? (iter & 1u) : ((iter >> 1) & 1u);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      // Every global access of the epilogue is staged through a per-warp shared-memory tile of 32 rows x 128 bytes:
      // a thread owns one accumulator ROW (TMEM lane), but a warp-wide store of "my row, 16 bytes" touches 32
      // different 128-byte lines.  Staged, lanes 8k..8k+7 move one row's 128 contiguous bytes: 4 full lines per
      // instruction.  (Measured on the attention backward: the direct version ran the P / dS GEMMs store-bound.)
      const uint32_t wst = stage_base + static_cast<uint32_t>(warp - 2) * kBgWarpStage;
      const int ncols = p.epi_warps == 8 ? p.bn / 2 : p.bn;          // columns this warp owns
      const int cbeg = p.epi_warps == 8 ? ((warp - 2) >> 2) * ncols : 0;
      const int piece = lane & 7, rsub = lane >> 3;
      const bool f32out = p.epi == BG_EPI_F32;
      const int cw = f32out ? 32 : 64;                               // columns per 128-byte row segment
      const int esz = f32out ? 4 : 2;
      for (int half = 0; half < p.mt2; ++half) {
      const int row0 = (mt * p.mt2 + half) * 128 + quad * 32;        // first row of this warp
      const int row = row0 + lane;
      const bool row_ok = row < p.m_valid;
      const long long zbase = z2 * p.o_z2 + z1 * p.o_z1;
      float rv = 0.f;
      if ((p.epi == BG_EPI_P || p.epi == BG_EPI_DS) && row_ok) rv = p.rowvec[z2 * p.rv_z2 + z1 * p.rv_z1 + row];
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (p.mt2 == 2 ? half * 256u : acc * 256u);
      for (int c0 = cbeg; c0 < cbeg + ncols; c0 += cw) {
        const int col0 = nt * p.bn + c0;
        uint32_t r[64];
        tmem_ld_32x32(t_row + c0, r);
        if (!f32out) tmem_ld_32x32(t_row + c0 + 32, r + 32);
        tmem_ld_wait();
        if (col0 >= p.n_store) continue;                              // warp-uniform
        uint32_t q[32];                                               // this row's 128 output bytes
        if (f32out) {
#pragma unroll
          for (int j = 0; j < 32; ++j) q[j] = r[j];
        } else {
          float v[64];
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = __uint_as_float(r[j]);
          if (p.epi == BG_EPI_P) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = (col0 + j < p.n_valid) ? ex2_approx(fmaf(v[j], p.c1, -rv)) : 0.f;
          } else if (p.epi == BG_EPI_DS) {
            // P tile: coalesced global -> staging tile, then every thread reads its own row
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = rsub + 4 * i, grow = row0 + rr;
              uint4 u = make_uint4(0u, 0u, 0u, 0u);
              if (grow < p.m_valid && col0 + piece * 8 < p.n_store)
                u = *reinterpret_cast<const uint4*>(p.pmat + zbase + static_cast<long long>(grow) * p.o_row + col0 + piece * 8);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wst + rr * kBgStagePitch + piece * 16), "r"(u.x),
                           "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint32_t a0, a1, a2, a3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                           : "r"(wst + lane * kBgStagePitch + j * 16) : "memory");
              const float2 p0 = unpack_bf16x2(a0), p1 = unpack_bf16x2(a1), p2 = unpack_bf16x2(a2), p3 = unpack_bf16x2(a3);
              const float pv[8] = {p0.x, p0.y, p1.x, p1.y, p2.x, p2.y, p3.x, p3.y};
#pragma unroll
              for (int e = 0; e < 8; ++e) v[8 * j + e] = pv[e] * (v[8 * j + e] - rv) * p.c1;
            }
            __syncwarp();
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) q[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wst + lane * kBgStagePitch + j * 16), "r"(q[4 * j]),
                       "r"(q[4 * j + 1]), "r"(q[4 * j + 2]), "r"(q[4 * j + 3]) : "memory");
        __syncwarp();
        const int ecol = col0 + piece * (16 / esz);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = rsub + 4 * i, grow = row0 + rr;
          if (grow < p.m_valid && ecol < p.n_store) {
            uint4 u;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                         : "r"(wst + rr * kBgStagePitch + piece * 16) : "memory");
            char* dst = static_cast<char*>(p.out) + (zbase + static_cast<long long>(grow) * p.o_row + ecol) * esz;
            *reinterpret_cast<uint4*>(dst) = u;
          }
        }
        __syncwarp();
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

__global__ void __launch_bounds__(256) bg_sum_splits_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                            int splits, long long n, int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double t = 0.0;
    for (int s = 0; s < splits; ++s) t += (double)partial[(long long)s * n + i];
    if (accumulate) t += (double)out[i];
    out[i] = (float)t;
  }
}
// many splits (small gradients split over the whole chip): 8 split lanes per element shorten the dependent chain;
// grid ceil(n / 32)
__global__ void __launch_bounds__(256) bg_sum_splits_lanes_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                                  int splits, long long n, int accumulate) {
  const long long i = (long long)blockIdx.x * 32 + (threadIdx.x & 31);
  const float* src = partial + i;
  double t = ordered_sum8(splits, i < n, [&](int s) { return __ldg(src + (long long)s * n); });
  if (threadIdx.x < 32 && i < n) {
    if (accumulate) t += (double)out[i];
    out[i] = (float)t;
  }
}

// D[b][h][row] = sum_d dO[b][row][h*64+d] * O[b][row][h*64+d]
__global__ void __launch_bounds__(256) attn_rowdot_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                                                          float* __restrict__ out, int batch, int tokens, int heads) {
  const long long total = (long long)batch * tokens * heads;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int h = (int)(i % heads);
    const long long br = i / heads;          // b * tokens + row
    const int row = (int)(br % tokens), b = (int)(br / tokens);
    const bf16* po = o + br * heads * 64 + h * 64;
    const bf16* pd = d_o + br * heads * 64 + h * 64;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 u = *reinterpret_cast<const uint4*>(po + 8 * j), w = *reinterpret_cast<const uint4*>(pd + 8 * j);
      const float2 a0 = unpack_bf16x2(u.x), a1 = unpack_bf16x2(u.y), a2 = unpack_bf16x2(u.z), a3 = unpack_bf16x2(u.w);
      const float2 b0 = unpack_bf16x2(w.x), b1 = unpack_bf16x2(w.y), b2 = unpack_bf16x2(w.z), b3 = unpack_bf16x2(w.w);
      acc += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
    }
    out[((long long)b * heads + h) * tokens + row] = acc;
  }
}

// ------------------------------------------------------------------------------------------ host
static int bg_encode(CUtensorMap* map, const void* ptr, const long long dims_[4], const long long strides_elems[3],
                     const int box_[4]) {
  cuuint64_t dims[4] = {(cuuint64_t)dims_[0], (cuuint64_t)dims_[1], (cuuint64_t)dims_[2], (cuuint64_t)dims_[3]};
  cuuint64_t strides[3] = {(cuuint64_t)strides_elems[0] * 2, (cuuint64_t)strides_elems[1] * 2, (cuuint64_t)strides_elems[2] * 2};
  for (int i = 0; i < 3; ++i)
    if (strides[i] % 16 != 0 || strides[i] == 0) return fail(ODB_ERR_INVALID, "bgemm: tensor strides must be positive multiples of 8 elements");
  cuuint32_t box[4] = {(cuuint32_t)box_[0], (cuuint32_t)box_[1], (cuuint32_t)box_[2], (cuuint32_t)box_[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                      CU_TENSOR_MAP_SWIZZLE_128B);
}

static int bg_launch(BgParams& p, cudaStream_t stream) {
  static bool configured[kMaxDevices] = {};
  const int dev = current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(bgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBgSmemMax);
    if (e != cudaSuccess) return fail_cuda(e, "bgemm: cudaFuncSetAttribute");
    configured[dev] = true;
  }
  if (p.epi_warps != 8) p.epi_warps = 4;
  if (p.mt2 == 2) { p.stages = 3; p.epi_warps = 4; }
  else if (p.stages != 3) p.stages = kBgStages;
  if (p.epi_warps == 8) p.stages = 3;
  const int smem = bg_smem_bytes(p.stages, p.mt2, p.epi_warps);
  if (smem > kBgSmemMax) return fail(ODB_ERR_INVALID, "bgemm: shared-memory plan too large");
  const long long units = (long long)p.MT * p.NT * p.Z1 * p.Z2;
  if (units < 1 || units > 0x7fffffffLL) return fail(ODB_ERR_INVALID, "bgemm: bad unit count");
  const int grid = units < num_sms() ? (int)units : num_sms();
  bgemm_kernel<<<grid, 64 + 32 * p.epi_warps, smem, stream>>>(p);
  count_launch();
  return check_launch("bgemm");
}

static void zero4(int* v) { v[0] = v[1] = v[2] = v[3] = 0; }
static void bg_operand_init(BgOperand& o) {
  zero4(o.off); zero4(o.mul_mn); zero4(o.mul_k1); zero4(o.mul_k2); zero4(o.mul_k3); zero4(o.mul_z1); zero4(o.mul_z2);
  o.mn_major = 0; o.nsub = 1; o.sub_dim = 0; o.sub_bytes = 0;
}

static void view_strides(const odb_view& v, long long* sx, long long* sy, long long* sb) {
  *sx = v.sx; *sy = v.sy; *sb = v.sb;
  if (v.w == 1 && *sx == 0) *sx = v.c;
  if (v.h == 1 && *sy == 0) *sy = (long long)v.w * *sx;
  if (v.b == 1 && *sb == 0) *sb = (long long)v.h * *sy;
}

struct WgradPlan { int tw, th, nk1, nk2, nk3, bn, mt2, MT, NT, splits, kps; long long row_len; };

static int wgrad_plan(const odb_wgrad_desc* d, long long workspace_bytes, WgradPlan* pl) {
  const int C = d->views[0].c, n = d->n;
  if (C % 8 || n % 8) return fail(ODB_ERR_INVALID, "conv_wgrad: C and n must be multiples of 8");
  const int ow = d->dy.w, oh = d->dy.h, ob = d->dy.b;
  if (ow >= 64) { pl->tw = 64; pl->th = 1; }
  else if (ow >= 32) { pl->tw = 32; pl->th = 2; }
  else if (ow >= 16) { pl->tw = 16; pl->th = 4; }
  else { pl->tw = 8; pl->th = 8; }
  pl->nk1 = (ow + pl->tw - 1) / pl->tw;
  pl->nk2 = (oh + pl->th - 1) / pl->th;
  pl->nk3 = ob;
  const int cpad = (C + 63) / 64 * 64;
  pl->bn = cpad < 256 ? cpad : 256;
  pl->mt2 = n > 128 ? 2 : 1;
  pl->MT = (n + 128 * pl->mt2 - 1) / (128 * pl->mt2);
  pl->NT = (C + pl->bn - 1) / pl->bn;
  pl->row_len = (long long)d->num_taps * C;
  const long long ksteps = (long long)pl->nk1 * pl->nk2 * pl->nk3;
  const long long tiles = (long long)pl->MT * pl->NT * d->num_taps;
  // one wave of CTAs: every extra split costs a full fp32 partial of the gradient (write + re-read in the reduction)
  long long splits = tiles >= num_sms() ? 1 : num_sms() / tiles;
  if (splits > ksteps / 8) splits = ksteps / 8;
  if (splits > 256) splits = 256;
  if (splits < 1) splits = 1;
  if (workspace_bytes >= 0) {
    const long long fit = workspace_bytes / ((long long)n * pl->row_len * 4);
    if (splits > fit) splits = fit;
    if (splits < 1) splits = 1;
  }
  pl->kps = (int)((ksteps + splits - 1) / splits);
  pl->splits = (int)((ksteps + pl->kps - 1) / pl->kps);
  return ODB_OK;
}

long long conv_wgrad_tc_workspace_bytes(const odb_wgrad_desc* d) {
  WgradPlan pl;
  if (wgrad_plan(d, -1, &pl)) return -1;
  return (long long)pl.splits * d->n * pl.row_len * 4;
}

int conv_wgrad_tc(const odb_wgrad_desc* d, cudaStream_t stream) {
  WgradPlan pl;
  int rc = wgrad_plan(d, d->workspace_bytes, &pl);
  if (rc) return rc;
  const int C = d->views[0].c, n = d->n;
  const bool direct = pl.splits == 1 && !d->accumulate;
  if (!direct && (d->workspace == nullptr || d->workspace_bytes < (long long)pl.splits * n * pl.row_len * 4))
    return fail(ODB_ERR_INVALID, "conv_wgrad: workspace too small");
  BgParams p;
  memset(&p, 0, sizeof(p));
  bg_operand_init(p.a);
  bg_operand_init(p.b);
  // A = dy, MN-major: M = output channel (dim 0), K = pixels (dims 1..3 in tw x th tiles)
  {
    long long sx, sy, sb;
    view_strides(d->dy, &sx, &sy, &sb);
    const long long dims[4] = {n, d->dy.w, d->dy.h, d->dy.b};
    const long long str[3] = {sx, sy, sb};
    const int box[4] = {64, pl.tw, pl.th, 1};
    if ((reinterpret_cast<uintptr_t>(d->dy.ptr) & 15u) != 0) return fail(ODB_ERR_INVALID, "conv_wgrad: dy must be 16-byte aligned");
    rc = bg_encode(&p.a.map[0], d->dy.ptr, dims, str, box);
    if (rc) return rc;
    for (int v = 1; v < ODB_MAX_VIEWS; ++v) p.a.map[v] = p.a.map[0];
    p.a.mn_major = 1; p.a.nsub = 2 * pl.mt2; p.a.sub_dim = 0; p.a.sub_bytes = kBgSubBytes;
    p.a.mul_mn[0] = 1; p.a.mul_k1[1] = pl.tw; p.a.mul_k2[2] = pl.th; p.a.mul_k3[3] = 1;
  }
  // B = input views, MN-major: N = input channel, same pixel traversal shifted by the tap
  for (int v = 0; v < ODB_MAX_VIEWS; ++v) {
    const odb_view& src = d->views[v < d->num_views ? v : 0];
    if (src.c != C) return fail(ODB_ERR_INVALID, "conv_wgrad: views disagree on channels");
    if ((reinterpret_cast<uintptr_t>(src.ptr) & 15u) != 0) return fail(ODB_ERR_INVALID, "conv_wgrad: view must be 16-byte aligned");
    long long sx, sy, sb;
    view_strides(src, &sx, &sy, &sb);
    const long long dims[4] = {C, src.w, src.h, src.b};
    const long long str[3] = {sx, sy, sb};
    const int box[4] = {64, pl.tw, pl.th, 1};
    rc = bg_encode(&p.b.map[v], src.ptr, dims, str, box);
    if (rc) return rc;
  }
  p.b.mn_major = 1; p.b.nsub = pl.bn / 64; p.b.sub_dim = 0; p.b.sub_bytes = kBgSubBytes;
  p.b.mul_mn[0] = 1; p.b.mul_k1[1] = pl.tw; p.b.mul_k2[2] = pl.th; p.b.mul_k3[3] = 1;
  p.nk1 = pl.nk1; p.nk2 = pl.nk2; p.nk3 = pl.nk3;
  p.MT = pl.MT; p.NT = pl.NT; p.Z1 = d->num_taps; p.Z2 = pl.splits;
  p.split_mode = 1; p.ksteps_per_split = pl.kps;
  p.mt2 = pl.mt2;
  p.bn = pl.bn;
  p.tap_mode = 1;
  for (int t = 0; t < d->num_taps; ++t) { p.tap_view[t] = d->tap_view[t]; p.tap_dx[t] = d->tap_dx[t]; p.tap_dy[t] = d->tap_dy[t]; }
  p.epi = BG_EPI_F32;
  p.out = direct ? static_cast<void*>(d->out) : d->workspace;
  p.o_row = pl.row_len; p.o_z1 = C; p.o_z2 = (long long)n * pl.row_len;
  p.m_valid = n; p.n_valid = C; p.n_store = C;
  rc = bg_launch(p, stream);
  if (rc || direct) return rc;
  const long long total = (long long)n * pl.row_len;
  if (pl.splits >= 16 && total <= 131072) {     // small gradient, many splits: the chain length is what costs
    bg_sum_splits_lanes_kernel<<<(unsigned)((total + 31) / 32), 256, 0, stream>>>(static_cast<const float*>(d->workspace), d->out,
                                                                                  pl.splits, total, d->accumulate);
  } else {
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
    bg_sum_splits_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float*>(d->workspace), d->out, pl.splits, total,
                                                               d->accumulate);
  }
  count_launch();
  return check_launch("conv_wgrad: split reduction");
}

// ------------------------------------------------------------------------------------------ attention backward
constexpr int kAttPad = 640;
long long attention_bwd_tc_workspace_bytes(int b, int tokens, int heads) {
  // P, dS: bf16 [b*heads][tokens][640]; D: fp32 [b*heads][tokens]
  return 2LL * b * heads * tokens * kAttPad * 2 + (long long)b * heads * tokens * 4 + 512;
}

int attention_bwd_tc(const void* qkv, const void* o, const void* d_o, const float* lse, void* dqkv, void* workspace,
                     long long workspace_bytes, int b, int tokens, int heads, float scale, cudaStream_t stream) {
  if (lse == nullptr) return fail(ODB_ERR_INVALID, "attention_bwd: the bf16 path needs the forward's lse");
  if (tokens > kAttPad) return fail(ODB_ERR_UNSUPPORTED, "attention_bwd: at most 640 tokens");
  if (workspace == nullptr || workspace_bytes < attention_bwd_tc_workspace_bytes(b, tokens, heads))
    return fail(ODB_ERR_INVALID, "attention_bwd: workspace too small");
  const int H = heads, T = tokens;
  const long long BH = (long long)b * H;
  bf16* P = static_cast<bf16*>(workspace);
  bf16* dS = P + BH * T * kAttPad;
  float* D = reinterpret_cast<float*>(dS + BH * T * kAttPad);
  {
    const long long total = BH * T;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
    attn_rowdot_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const bf16*>(o), static_cast<const bf16*>(d_o), D, b, T, H);
    count_launch();
  }
  // tensor maps
  const long long qkv_dims[4] = {64, 3LL * H, T, b};
  const long long qkv_str[3] = {64, 3LL * H * 64, (long long)T * 3 * H * 64};
  const long long do_dims[4] = {64, H, T, b};
  const long long do_str[3] = {64, (long long)H * 64, (long long)T * H * 64};
  const long long ps_dims[4] = {kAttPad, T, BH, 1};
  const long long ps_str[3] = {kAttPad, (long long)T * kAttPad, BH * T * kAttPad};
  const int box_rows128[4] = {64, 1, 128, 1}, box_rows64[4] = {64, 1, 64, 1};
  const int box_ps_k[4] = {64, 128, 1, 1}, box_ps_mn[4] = {64, 64, 1, 1};
  CUtensorMap qkv128, qkv64, do128, do64, p_mn, ds_k, ds_mn;
  int rc;
  if ((rc = bg_encode(&qkv128, qkv, qkv_dims, qkv_str, box_rows128))) return rc;
  if ((rc = bg_encode(&qkv64, qkv, qkv_dims, qkv_str, box_rows64))) return rc;
  if ((rc = bg_encode(&do128, d_o, do_dims, do_str, box_rows128))) return rc;
  if ((rc = bg_encode(&do64, d_o, do_dims, do_str, box_rows64))) return rc;
  if ((rc = bg_encode(&p_mn, P, ps_dims, ps_str, box_ps_mn))) return rc;
  if ((rc = bg_encode(&ds_k, dS, ps_dims, ps_str, box_ps_k))) return rc;
  if ((rc = bg_encode(&ds_mn, dS, ps_dims, ps_str, box_ps_mn))) return rc;
  auto set_maps = [](BgOperand& o, const CUtensorMap& m) { for (int v = 0; v < ODB_MAX_VIEWS; ++v) o.map[v] = m; };
  const int MTq = (T + 127) / 128;

  BgParams p;
  // ---- 1. P = exp2(Q K^T c - lse)   and   2. dS = P o (dO V^T - D) * scale      (K-major x K-major, K = 64)
  for (int pass = 0; pass < 2; ++pass) {
    memset(&p, 0, sizeof(p));
    bg_operand_init(p.a); bg_operand_init(p.b);
    set_maps(p.a, pass == 0 ? qkv128 : do128);
    p.a.sub_bytes = kBgAStage; p.a.mul_mn[2] = 1; p.a.mul_z1[1] = 1; p.a.mul_z2[3] = 1;
    set_maps(p.b, qkv128);
    p.b.sub_bytes = 128 * 128; p.b.mul_mn[2] = 1; p.b.mul_z1[1] = 1; p.b.mul_z2[3] = 1;
    p.b.off[1] = pass == 0 ? H : 2 * H;                         // K slot / V slot
    p.nk1 = p.nk2 = p.nk3 = 1;
    p.MT = MTq; p.NT = kAttPad / 128; p.Z1 = H; p.Z2 = b;
    p.mt2 = 1;
    p.epi_warps = 8;
    p.bn = 128;
    p.epi = pass == 0 ? BG_EPI_P : BG_EPI_DS;
    p.out = pass == 0 ? static_cast<void*>(P) : static_cast<void*>(dS);
    p.o_row = kAttPad; p.o_z1 = (long long)T * kAttPad; p.o_z2 = (long long)H * T * kAttPad;
    p.m_valid = T; p.n_valid = T; p.n_store = kAttPad;
    p.rowvec = pass == 0 ? lse : D; p.rv_z1 = T; p.rv_z2 = (long long)H * T;
    p.pmat = P;
    p.c1 = pass == 0 ? scale * 1.4426950408889634f : scale;
    if ((rc = bg_launch(p, stream))) return rc;
  }
  // ---- 3. dQ = dS K    (A K-major over the keys, B = K MN-major)
  bf16* dq = static_cast<bf16*>(dqkv);
  {
    memset(&p, 0, sizeof(p));
    bg_operand_init(p.a); bg_operand_init(p.b);
    set_maps(p.a, ds_k);
    p.a.sub_bytes = kBgAStage; p.a.mul_mn[1] = 1; p.a.mul_k1[0] = 64; p.a.mul_z1[2] = 1; p.a.mul_z2[2] = H;
    set_maps(p.b, qkv64);
    p.b.mn_major = 1; p.b.nsub = 1; p.b.sub_bytes = kBgSubBytes;
    p.b.off[1] = H; p.b.mul_k1[2] = 64; p.b.mul_z1[1] = 1; p.b.mul_z2[3] = 1;
    p.nk1 = kAttPad / 64; p.nk2 = p.nk3 = 1;
    p.MT = MTq; p.NT = 1; p.Z1 = H; p.Z2 = b;
    p.mt2 = 1;
    p.bn = 64;
    p.epi = BG_EPI_BF16;
    p.out = dq;
    p.o_row = 3LL * H * 64; p.o_z1 = 64; p.o_z2 = (long long)T * 3 * H * 64;
    p.m_valid = T; p.n_valid = 64; p.n_store = 64;
    if ((rc = bg_launch(p, stream))) return rc;
  }
  // ---- 4. dK = dS^T Q   and   5. dV = P^T dO     (A MN-major over the keys, B MN-major; K = queries)
  for (int pass = 0; pass < 2; ++pass) {
    memset(&p, 0, sizeof(p));
    bg_operand_init(p.a); bg_operand_init(p.b);
    set_maps(p.a, pass == 0 ? ds_mn : p_mn);
    p.a.mn_major = 1; p.a.nsub = 2; p.a.sub_dim = 0; p.a.sub_bytes = kBgSubBytes;
    p.a.mul_mn[0] = 1; p.a.mul_k1[1] = 64; p.a.mul_z1[2] = 1; p.a.mul_z2[2] = H;
    set_maps(p.b, pass == 0 ? qkv64 : do64);
    p.b.mn_major = 1; p.b.nsub = 1; p.b.sub_bytes = kBgSubBytes;
    p.b.mul_k1[2] = 64; p.b.mul_z1[1] = 1; p.b.mul_z2[3] = 1;           // Q slot 0 / dO
    p.nk1 = (T + 63) / 64; p.nk2 = p.nk3 = 1;
    p.MT = MTq; p.NT = 1; p.Z1 = H; p.Z2 = b;
    p.mt2 = 1;
    p.bn = 64;
    p.epi = BG_EPI_BF16;
    p.out = dq + (pass == 0 ? 1 : 2) * H * 64;
    p.o_row = 3LL * H * 64; p.o_z1 = 64; p.o_z2 = (long long)T * 3 * H * 64;
    p.m_valid = T; p.n_valid = 64; p.n_store = 64;
    if ((rc = bg_launch(p, stream))) return rc;
  }
  return ODB_OK;
}

}  // namespace odb
