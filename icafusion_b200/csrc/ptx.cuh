// Thin inline-PTX wrappers for the sm_100a features the kernels use:
// mbarrier, cp.async (LDGSTS), TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld), proxy fences.
// Bit layouts follow the PTX ISA (sm_100a) -- the same encodings CUTLASS' cute/arch/mma_sm100_desc.hpp documents.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace icaf {

#ifndef ICAF_SPIN_LIMIT
#define ICAF_SPIN_LIMIT (1u << 22)   // mbarrier polls before declaring a deadlock and trapping
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (CUDA error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > ICAF_SPIN_LIMIT) {
      printf("icaf: mbarrier deadlock block(%d,%d,%d) thread %d bar 0x%x parity %u\n", blockIdx.x, blockIdx.y,
             blockIdx.z, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// launch_dependents: the next kernel in the stream may start its prologue now; wait: block until every kernel this
// launch programmatically depends on has completed and flushed its writes.  Both are no-ops without the launch attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------ proxy fences
// generic-proxy smem writes (st.shared / completed cp.async) -> visible to the async proxy (UMMA, TMA)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------ cp.async (LDGSTS), 16B with zero-fill
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  uint32_t n = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
// Arrive on `bar` (counts as this thread's arrival, .noinc) once every cp.async this thread has issued so far has landed.
// Non-blocking: the producer keeps issuing; this is how CUTLASS' sm100 cp.async mainloops hand stages to tcgen05.mma.
__device__ __forceinline__ void cp_async_arrive_on(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------ TMA (2D tiled load, mbarrier completion)
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ------------------------------------------------------------------ tcgen05: TMEM allocation
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst) {   // one full warp
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM columns: power of two in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {    // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------ tcgen05: descriptors
// Instruction descriptor, kind::f16, fp16 A/B (K-major both), fp32 accumulate.
//  [4,6) D format 1=f32 | [7,10) A fmt 0=f16 | [10,13) B fmt 0=f16 | 15 A major 0=K | 16 B major 0=K
//  [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
// Shared-memory matrix descriptor for a K-major tile stored as rows of 128 B (64 halfs) with the 128-byte swizzle:
// 8-row groups are 1024 B apart (SBO), LBO unused for swizzled K-major. bits [46,48) = 1 (sm_100 version),
// [61,64) = 2 (SWIZZLE_128B). The tile base must be 1024-B aligned; stepping K by 16 halfs adds 32 B to the start.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4);        // start address  [0,14)
  d |= uint64_t(1) << 16;                       // LBO (ignored)  [16,30)
  d |= uint64_t(1024 >> 4) << 32;               // SBO            [32,46)
  d |= uint64_t(1) << 46;                       // version
  d |= uint64_t(2) << 61;                       // SWIZZLE_128B
  return d;
}

// K-major tile whose rows are `row_bytes` (32 / 64 / 128) wide with the matching 32B / 64B / 128B swizzle: 8-row groups are
// 8*row_bytes apart.  layout type field: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t row_bytes) {
  const uint64_t lt = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4);
  d |= uint64_t(1) << 16;
  d |= uint64_t((8 * row_bytes) >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= lt << 61;
  return d;
}

// MN-major operand tile (the M / N index is the contiguous one): rows of `row_bytes` (32 / 64 / 128) hold 16 / 32 / 64
// consecutive M|N elements of ONE k, consecutive k are consecutive rows, 8-row groups are 8*row_bytes apart (SBO); the next
// block of 64 (32, 16) M|N elements starts `lbo_bytes` further (LBO).  This is the layout a TMA box (cols = M|N, rows = k)
// with the matching swizzle produces -- CUTLASS' Layout_MN_SW{32,64,128}_Atom.  One MMA (K = 16) spans two 8-row groups:
// advance the start address by 16 * row_bytes per K step.  Needs the matching "major" bit in the instruction descriptor.
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t saddr, uint32_t row_bytes, uint32_t lbo_bytes) {
  const uint64_t lt = row_bytes == 128 ? 2 : (row_bytes == 64 ? 4 : 6);
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((8 * row_bytes) >> 4) << 32;
  d |= uint64_t(1) << 46;
  d |= lt << 61;
  return d;
}
// kind::f16 instruction descriptor with MN-major A and / or B (bit 15 / bit 16)
__host__ __device__ constexpr uint32_t umma_idesc_f16_major(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            bool accumulate) {
  uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// Arrive on an mbarrier when all tcgen05 ops issued so far by this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread i gets lane base+i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 16-column flavour (the persistent kernel double-buffers these against the epilogue math).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------ thread-block clusters / distributed shared memory
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {   // address of the same offset in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ------------------------------------------------------------------ 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256)
// One full 32-byte L2 sector per lane and instruction; the pointer must be 32-byte aligned.
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&o)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
               "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
}
__device__ __forceinline__ void ld_global_nc_v8(const void* p, uint32_t (&o)[8]) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]),
               "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]) : "l"(p));
}

// ------------------------------------------------------------------ counter-based dropout mask of the attention probabilities
// keep(dir, batch*head, query, key): the training forward (attn.cu) and the backward (attn_bwd.cu) regenerate the same mask.
__device__ __forceinline__ uint32_t attn_hash(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}
__device__ __forceinline__ bool attn_keep(uint32_t seed, int dir, int bh, int q, int k, float p) {
  const uint32_t h = attn_hash(uint32_t(q) * 65536u + uint32_t(k & 0xffff), uint32_t(bh) * 2u + uint32_t(dir) + (uint32_t(k) >> 16) * 0x10001u, seed);
  return float(h >> 8) * (1.f / 16777216.f) >= p;
}

// ------------------------------------------------------------------ small math helpers
// x * sigmoid(x) with MUFU.EX2 + MUFU.RCP (rel. error ~1e-6, far below the fp16 output rounding)
__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
__device__ __forceinline__ float gelu_erf_f(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

}  // namespace icaf
