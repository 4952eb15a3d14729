// Shared device helpers for the sm_100a kernels of omnidata_b200.
// Thin inline-PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (UMMA + TMEM).
// Nothing here is reference-derived; the reference (EPFL-VILAB/omnidata) has no native code on this path.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ODB_DEVINL __device__ __forceinline__

namespace odb {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- shared-memory addressing
ODB_DEVINL uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
ODB_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
ODB_DEVINL void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
ODB_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
ODB_DEVINL void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
ODB_DEVINL bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin with a watchdog: a protocol bug must abort the kernel (sticky error on the host)
// instead of hanging the GPU box.
ODB_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("odb: mbarrier watchdog block %d thread %d bar %u parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------- programmatic dependent launch
// Every forward-path kernel is launched with programmaticStreamSerialization: it may start (and run
// its prologue: barrier init, TMEM allocation, descriptor prefetch) while the previous kernel of the
// stream drains.  grid_dep_wait() must precede the first access to global memory.
ODB_DEVINL void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
ODB_DEVINL void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- fences / named barriers
ODB_DEVINL void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
ODB_DEVINL void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- TMA
ODB_DEVINL void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
ODB_DEVINL void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0,
                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
ODB_DEVINL void tma_load_4d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0,
                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
ODB_DEVINL void tma_store_4d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2,
                             int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(map)),
      "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
ODB_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
ODB_DEVINL void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
ODB_DEVINL void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------- tcgen05 / TMEM
ODB_DEVINL void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
ODB_DEVINL void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
ODB_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
ODB_DEVINL void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
ODB_DEVINL void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, single CTA.
ODB_DEVINL void umma_bf16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the two shared-memory descriptors given as their low words (start address >> 4 | LBO) plus the common
// high word: small tiles (N = 32: 16 tensor-pipe cycles per instruction) are bound by how fast ONE thread can issue,
// so the per-instruction descriptor arithmetic is kept to 32-bit adds.
constexpr uint32_t kUmmaDescHiSw128 = 64u | (1u << 14) | (2u << 29);
ODB_DEVINL uint32_t umma_desc_lo_sw128(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFFu) | (1u << 16); }
ODB_DEVINL void umma_bf16_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kUmmaDescHiSw128)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
// (implicitly performs tcgen05.fence::before_thread_sync).
ODB_DEVINL void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
ODB_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i).
ODB_DEVINL void tmem_ld_32x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// waits for this thread's outstanding tcgen05.ld AND ties the destination registers to the wait, so
// that no use of r can be scheduled above it (needed once loads are issued ahead of their use)
ODB_DEVINL void tmem_ld_wait_regs(uint32_t* r) {
  asm volatile(
      "tcgen05.wait::ld.sync.aligned;\n"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
        "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]),
        "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]),
        "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
        "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
      :
      : "memory");
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2)
ODB_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
ODB_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank`
ODB_DEVINL uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
ODB_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
// TMA loads issued by either CTA of a pair; the transaction bytes are credited to the mbarrier at
// `bar_cluster` (the leader CTA's barrier, a shared::cluster address).
ODB_DEVINL void tma_load_2d_cg2(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_cluster,
                                int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
ODB_DEVINL void tma_load_4d_cg2(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar_cluster,
                                int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
ODB_DEVINL void tmem_alloc_cg2(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(ncols)
               : "memory");
}
ODB_DEVINL void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
ODB_DEVINL void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of each CTA] * B[N halves in the two CTAs]; issued by the leader.
ODB_DEVINL void umma_bf16_ss_cg2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit -> arrive on the barrier at the same smem offset in every CTA of `cta_mask`
ODB_DEVINL void umma_commit_cg2(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(bar),
      "h"(cta_mask)
      : "memory");
}

// Shared-memory matrix descriptor for a K-major bf16 operand tile stored as dense 128-byte rows
// (64 bf16 of K per row) with the 128B TMA swizzle: 8-row groups are 1024 B apart (SBO), LBO = 1
// (ignored for swizzled K-major), descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.
ODB_DEVINL uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) |
         (1ull << 46) | (2ull << 61);
}
// Instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7, 10), both K-major, N>>3 at bit 17,
// M>>4 at bit 24.
ODB_DEVINL constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// ---------------------------------------------------------------- small numerics helpers
// erf via Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7, far below one bf16 ulp of the GELU output):
// 1 MUFU.RCP + 1 MUFU.EX2 + 9 FMA-class ops instead of the ~30-instruction libdevice erff.
ODB_DEVINL float erf_as(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = __expf(-ax * ax);
  return copysignf(fmaf(-p * t, e, 1.0f), x);
}
ODB_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752f)); }

// ---- packed fp32 pairs (sm_100: FFMA2 — one issue slot for two fused multiply-adds)
ODB_DEVINL uint64_t f32x2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
ODB_DEVINL void f32x2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
ODB_DEVINL uint64_t f32x2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
ODB_DEVINL float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Exact-erf GELU for the GEMM epilogue, two elements at a time:
//   gelu(x) = x Phi(x) = max(x, 0) - |x| Phi(-|x|),   Phi(-a) = 2^P(a),  a = min(|x|, 8)
// P = degree-10 weighted-minimax fit of log2 Phi(-a) on [0, 8] (fit: |gelu error| < 4e-8 absolute AND
// < 6e-6 relative — the tail x -> -inf keeps RELATIVE accuracy because the error sits in the
// exponent; measured in fp32 against float64 x Phi(x): 2.4e-7 = half an ulp of the result).
// Cost per element: 5 FFMA2 + 1 MUFU.EX2 + 3 ALU — half the MUFU work and ~60 % of the issue slots
// of the Abramowitz-Stegun form above, which is what bounds the fc1 epilogue (B200: 16 MUFU/clk/SM).
ODB_DEVINL void gelu_erf_x2(float& x0, float& x1) {
  const float a0 = fminf(fabsf(x0), 8.0f), a1 = fminf(fabsf(x1), 8.0f);
  const uint64_t a = f32x2_pack(a0, a1);
#define ODB_C2(c) f32x2_pack(c, c)
  uint64_t p = f32x2_fma(ODB_C2(-4.45074694e-09f), a, ODB_C2(1.77394618e-07f));
  p = f32x2_fma(p, a, ODB_C2(-2.91884744e-06f));
  p = f32x2_fma(p, a, ODB_C2(2.41618334e-05f));
  p = f32x2_fma(p, a, ODB_C2(-7.56097581e-05f));
  p = f32x2_fma(p, a, ODB_C2(-4.90890656e-04f));
  p = f32x2_fma(p, a, ODB_C2(7.66879548e-03f));
  p = f32x2_fma(p, a, ODB_C2(-5.30659795e-02f));
  p = f32x2_fma(p, a, ODB_C2(-4.58926226e-01f));
  p = f32x2_fma(p, a, ODB_C2(-1.15116936e+00f));
  p = f32x2_fma(p, a, ODB_C2(-9.99995267e-01f));
#undef ODB_C2
  float p0, p1;
  f32x2_unpack(p, p0, p1);
  x0 = fmaf(-a0, ex2_approx(p0), fmaxf(x0, 0.0f));
  x1 = fmaf(-a1, ex2_approx(p1), fmaxf(x1, 0.0f));
}

ODB_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
ODB_DEVINL float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

ODB_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
ODB_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- ordered reduction of per-block partials (256-thread blocks)
// Block = 32 adjacent columns x 8 part lanes: lane l sums parts l, l+8, ... in fp64 (independent loads, four in flight),
// then the lane-0 threads add the 8 lane sums in lane order.  The order is fixed for a given `parts`, so results are
// bit-reproducible, and the dependent chain is parts / 8 long instead of parts.  The value is returned to the threads
// of part lane 0 (threadIdx.x < 32).
template <typename F>
ODB_DEVINL double ordered_sum8(int parts, bool active, F&& part) {
  __shared__ double sh_os8[8][33];
  const int col = threadIdx.x & 31, pl = threadIdx.x >> 5;
  double t = 0.0;
  if (active) {
#pragma unroll 4
    for (int p = pl; p < parts; p += 8) t += (double)part(p);
  }
  sh_os8[pl][col] = t;
  __syncthreads();
  double s = 0.0;
  if (pl == 0) {
#pragma unroll
    for (int l = 0; l < 8; ++l) s += sh_os8[l][col];
  }
  return s;
}


}  // namespace odb
