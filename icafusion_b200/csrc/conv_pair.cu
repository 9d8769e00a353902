// CTA-pair (tcgen05 cta_group::2) variant of the persistent implicit-GEMM conv kernel for the wide layers (N >= 256).
//
// Why: the wide 3x3 / 1x1 layers are bound by the bytes each SM has to pull out of L2 per K block, not by the tensor
// pipe (tools/conv_probe.py: a 128x256 tile streams 16 KB of A + 32 KB of B per 512 MMA-clocks, the chip delivers
// ~40 B/clk/SM).  Two SMs of a TPC therefore work as one: a cluster of two CTAs owns a 256 x 256 output tile, each CTA
// loads its own 128 activation rows but only HALF of the filter tile (128 of the 256 output channels), and the leader
// CTA issues tcgen05.mma.cta_group::2 (M = 256, N = 256): each SM's tensor core reads its A rows and both halves of B
// (its own and the peer's shared memory).  Bytes per SM and K block drop from 48 KB to 32 KB for the same MMA work.
//
// Protocol (per CTA the same barrier layout; "leader" = cluster rank 0):
//   full[s]   leader only: 1 arrival (its producer's expect_tx of BOTH CTAs' bytes); both CTAs' TMA loads complete_tx on it
//             (cp.async.bulk.tensor ... .cta_group::2 with the leader's barrier address).
//   empty[s]  both: tcgen05.commit.cta_group::2 ... multicast::cluster (mask 0b11) from the leader's MMA thread.
//   tfull[b]  both: same multicast commit after the last K block of a tile.
//   tempty[b] leader only: one arrival per epilogue warp of the buffer, both CTAs (the peer's arrive remotely).
//   TMEM holds four accumulator buffers for BN <= 128 (two for BN = 256), so the cluster-scope hand-shakes of up to four
//   tiles overlap.
// Everything else (tile loop, four epilogue groups, bias staging, 256-bit stores) is conv_persist.cu's.
//
// Halo mode (P.halo, 3x3 / stride 1 / pad 1 layers, tile = 16 rows x 8 pixels): the nine taps of a 64-channel block read
// the same (16+2) x (8+2) pixel patch.  Instead of nine shifted TMA boxes, three x-shifted copies of the 18-row patch are
// loaded (box 64 ch x 8 px x 18 rows at x0-1+kx); a copy is 18 swizzle atoms of 8 pixels x 128 B, so the tap (ky, kx) is
// the plain K-major operand that starts ky atoms (ky * 1024 B) into copy kx -- a 1024-byte aligned descriptor, nothing
// exotic.  K loop: channel block -> kx; a ring stage holds one copy plus the three filter half-tiles of its ky taps, so
// one barrier round trip and one commit feed 12 MMAs.  Activation bytes per SM drop 2.7x (9 x 16 KB -> 3 x 18 KB per channel block).
// Halo mode 2: when the whole K extent is one channel block and N fits one tile (the stem, the 64 -> 64 channel P2
// bottlenecks) the nine filter half-tiles are loaded once per problem and stay resident; only activation copies stream.
// (ncu on the stem: 1.89 GB of TMA traffic for 84 MB of input -- 39 % of it the same 36 KB of filter, reloaded per tile.)
// 16- / 32-channel maps (the image stem over the space-to-depth frame) take the same path: the TMA box still asks for 64
// channels and the unit zero-fills the ones the tensor does not have, so the copies keep 128-byte rows and the 128B
// swizzle; only the K steps that hold real channels are issued, and the filter box of tap t starts at K offset t * Cin.
// (A first version staged 32- / 64-byte rows with the matching narrow swizzle: every MMA then cost ~200 clk instead of
// ~40 -- tools/stem_probe.py: 16 channels 277 us, 32 channels 412 us, 64 channels 240 us for the same output.)
// 576 threads per CTA: warps 0-15 epilogue, warp 16 MMA issuer (leader) + TMEM allocator (both), warp 17 TMA producer.
#include <cstring>

#include "conv_common.cuh"

namespace icaf {

constexpr int kQEpiWarps = 16;
constexpr int kQThreads = (kQEpiWarps + 2) * 32;
constexpr int kQMaxStages = 10;
constexpr int kQABytes = BM * BK * 2;         // this CTA's 128 activation rows
constexpr int kQHaloABytes = 18 * 8 * 128;    // halo mode: (16 + 2) rows x 8 pixels x 64 channels of one x-shifted copy

template <int BN>                             // BN = tile width = UMMA N (256, 128 or 64)
struct QSmem {
  static constexpr int kBBytes = (BN / 2) * BK * 2;   // this CTA's half of the filter tile
  static constexpr int kStageBytes = kQABytes + kBBytes;
  static constexpr int kBufs = BN <= 128 ? 4 : 2;     // accumulator buffers in TMEM
  static constexpr int kHalves = 4 / kBufs;           // epilogue groups per buffer (four groups in total)
  static constexpr int kCW = BN / kHalves;            // columns per epilogue group
  static constexpr int kBarBytes = 512;               // barrier block (see the index map in the kernel)
  static constexpr int kTail = kBarBytes + 4 * 2 * kCW * 4 + 1024;
  static constexpr int kStagesFit = (227 * 1024 - kTail) / kStageBytes;
  static constexpr int kStages = kStagesFit > kQMaxStages ? kQMaxStages : kStagesFit;
  static constexpr int kSmem = kStages * kStageBytes + kTail;
  static constexpr int kTmemCols = kBufs * BN;
  // halo mode (3x3 / stride 1): one ring of stages = [x-shifted activation copy | filter half-tiles of its three ky taps]
  static constexpr int kHaloStageBytes = kQHaloABytes + 3 * kBBytes;
  static constexpr int kHaloFit = (227 * 1024 - kTail) / kHaloStageBytes;
  static constexpr int kHaloStages = kHaloFit > kQMaxStages ? kQMaxStages : kHaloFit;
  static constexpr int kHaloRing = kHaloStages * kHaloStageBytes;
  static constexpr int kHaloSmem = kHaloRing + kTail;
  // halo mode 2 (one channel block, one N tile): the nine filter half-tiles stay resident, the ring holds copies only
  static constexpr int kResBytes = 9 * kBBytes;
  static constexpr int kResFit = (227 * 1024 - kTail - kResBytes) / kQHaloABytes;
  static constexpr int kResStages = kResFit > kQMaxStages ? kQMaxStages : (kResFit < 1 ? 1 : kResFit);
  static constexpr int kResRing = kResBytes + kResStages * kQHaloABytes;
  static constexpr int kResSmem = kResRing + kTail;
};

// ---- cta_group::2 PTX (same encodings CUTLASS' SM100_TMA_2SM_LOAD / umma_arrive_multicast_2x1SM use)
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst) {    // one warp in EACH CTA of the pair, same smem offset
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma2_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// Arrive on the barrier at this offset in both CTAs once every tcgen05 op issued so far by this thread has completed.
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const void* tmap, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const void* tmap, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// Remote arrive on the leader's barrier.  Deliberately the plain form (what CUTLASS' umma_arrive_2x1SM_sm0 emits): the
// explicit .release.cluster variant compiles to MEMBAR.ALL.GPU + ERRBAR in front of the arrive, i.e. every epilogue warp
// waited for all of its outstanding global stores once per tile (ncu: 13 % of the stem kernel's stall samples).  The
// accumulator hand-back needs no memory ordering beyond tcgen05.wait::ld + tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

struct PairTile { int z, mtile, n0, tb, oy0, ox0, m0; };

// pair-tile t -> this CTA's output tile.  Sequence: n fastest, then pair of M tiles, then problem.
__device__ __forceinline__ PairTile pair_tile(const ConvParams& P, int t, int rank, int m_pairs, int n_tiles, int BN) {
  PairTile c;
  const int per_z = m_pairs * n_tiles;
  c.z = t / per_z;
  t -= c.z * per_z;
  const int mp = t / n_tiles;
  c.n0 = (t - mp * n_tiles) * BN;
  c.mtile = 2 * mp + rank;
  c.m0 = c.mtile * BM; c.tb = 0; c.oy0 = 0; c.ox0 = 0;
  if (P.a_mode == A_TMA4D) {
    const int per_img = P.tiles_x * P.tiles_y;
    c.tb = c.mtile / per_img;                 // an odd tile count leaves the peer a tile past the last image: all zero fill
    const int r = c.mtile - c.tb * per_img;
    c.oy0 = (r / P.tiles_x) * P.th;
    c.ox0 = (r % P.tiles_x) * P.tw;
    c.m0 = 0;
  }
  return c;
}

// Same as conv_persist.cu's epi_tile, except that the accumulator buffer is handed back to the LEADER's MMA thread.
template <int CW, int ACT, int RES, int XM = 0>
__device__ __forceinline__ void epi_tile_pair(uint32_t trow, uint32_t tempty_leader, const float* sb, float rbias, float alpha,
                                              float beta, const __half* rrow, __half* yrow, int al_row, bool mvalid, int nrem, EpiRow& ex) {
  uint32_t acc0[16], acc1[16];
  auto chunk = [&](const uint32_t (&acc)[16], int cb) {
    const int nc = nrem - cb;
    if (mvalid && nc > 0)
      epi_chunk16<ACT, RES, XM>(acc, sb + cb, rbias, alpha, beta, rrow ? rrow + cb : nullptr, yrow + cb, nc >= 16 ? al_row : 0, nc, true, ex, cb);
  };
  tmem_ld16(trow, acc0);
#pragma unroll 1
  for (int cb = 0; cb < CW; cb += 32) {
    tmem_ld_wait();
    tmem_ld16(trow + cb + 16, acc1);
    chunk(acc0, cb);
    tmem_ld_wait();
    if (cb + 32 < CW) {
      tmem_ld16(trow + cb + 32, acc0);
    } else {
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive_cluster(tempty_leader);
    }
    chunk(acc1, cb + 16);
  }
}

template <int BN, bool XM>
__global__ void __launch_bounds__(kQThreads, 1)
conv_gemm_pair_kernel(const ConvParams P, const __grid_constant__ ConvMaps maps, int total_pairs, int m_tiles, int m_pairs, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  using L = QSmem<BN>;
  constexpr int kQStages = L::kStages;
  constexpr int kQStageBytes = L::kStageBytes;
  constexpr int kQBBytes = L::kBBytes;
  constexpr int kQCW = L::kCW;
  constexpr int kBufs = L::kBufs, kHalves = L::kHalves;
  constexpr int kQBN = BN;
  const int halo = P.halo;                            // 1: x-shifted copies instead of tap boxes (3x3 / stride 1); 2: + resident filter
  const int ncb = (P.Cin + BK - 1) / BK;              // halo: 64-channel blocks (1 for the 16- / 32-channel stem maps)
  const int kst = (P.Cin < BK ? P.Cin : BK) / 16;     // halo: 16-element K steps per tap that hold real channels
  const uint32_t bar_off = halo == 2 ? uint32_t(L::kResRing) : (halo ? uint32_t(L::kHaloRing) : uint32_t(kQStages) * kQStageBytes);
  const uint32_t bar_base = smem_base + bar_off;
  // barrier block (8-byte slots): 0-9 stage full, 10-19 stage empty, 20 / 30 resident filter full / released (halo mode 2),
  // 40-43 accumulator full, 44-47 accumulator empty, 48 TMEM base address
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (10 + s); };
  const uint32_t bres_full = bar_base + 8u * 20, bres_empty = bar_base + 8u * 30;
  auto tfull_bar = [&](int b) { return bar_base + 8u * (40 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (44 + b); };
  const uint32_t tmem_slot = bar_base + 8u * 48;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int tid = threadIdx.x;
  const int rank = int(cluster_ctarank());            // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int nkb = P.k_pad / BK;
  const int a_mode = P.a_mode;

  if (tid == 0) {
    for (int s = 0; s < kQMaxStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(bres_full, 1);
    mbar_init(bres_empty, 1);
    for (int b = 0; b < kBufs; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 8 * kHalves);      // one arrival per epilogue warp of the buffer, both CTAs
    }
    fence_mbar_init();
  }
  if (warp == kQEpiWarps) tmem_alloc2<L::kTmemCols>(tmem_slot);
  if (warp == kQEpiWarps + 1 && lane_id() == 0) {
    tma_prefetch_desc(&maps.w[0]);
    tma_prefetch_desc(&maps.w[1]);
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.a[1]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_arrive();                                   // the peer's barriers exist before anything is signalled across
  cluster_wait();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 8 * 48);

  if (warp < kQEpiWarps) {
    // ------------------------------------------------------------------ epilogue groups (both CTAs, own 128 rows)
    const int eg = warp >> 2;
    const int buf = eg / kHalves;
    const int half = eg % kHalves;
    const int gt = tid & 127;
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const int ry = a_mode == A_TMA4D ? gt / P.tw : 0;
    const int rx = a_mode == A_TMA4D ? gt - ry * P.tw : 0;
    const int mode = (P.epi & ICAF_EPI_SCALED_RES) ? 2 : ((P.epi & ICAF_EPI_ADD_RES) ? 1 : 0);
    // XM instantiation (DMFF linears): 9 / 10 = LayerNorm folded into this GEMM (no activation / GELU); 11 = scaled
    // residual + statistics of the output rows
    const int mode_act = XM ? (P.ln_parts > 0 ? (P.act == ICAF_ACT_GELU ? 10 : 9) : 11) : P.act * 3 + mode;
    const bool row_bias = (P.epi & ICAF_EPI_BIAS_ROW) != 0;
    float* sbias = reinterpret_cast<float*>(smem_gen + bar_off + L::kBarBytes) + eg * 2 * kQCW;
    const uint32_t tempty_leader = map_to_cta(tempty_bar(buf), 0);
    auto bias_of = [&](int t) -> float {
      if (t >= total_pairs || gt >= kQCW) return 0.f;
      const PairTile q = pair_tile(P, t, rank, m_pairs, n_tiles, BN);
      const float* pb = q.z ? P.p[1].bias : P.p[0].bias;
      const int n = q.n0 + half * kQCW + gt;
      return (pb && !row_bias && n < P.N) ? __ldg(pb + n) : 0.f;
    };
    const int step = kBufs * n_clusters;
    int t = cluster_id + buf * n_clusters;
    float bnext = bias_of(t);
    for (int it = 0; t < total_pairs; t += step, ++it) {
      const PairTile c = pair_tile(P, t, rank, m_pairs, n_tiles, BN);
      const ConvProblem pr = pick_problem(P, c.z);
      float* sb = sbias + (it & 1) * kQCW;
      if (gt < kQCW) sb[gt] = bnext;
      named_bar_sync(1 + eg, 128);
      bnext = bias_of(t + step);
      int m;
      bool mvalid;
      if (a_mode == A_TMA4D) {
        m = (c.tb * P.Ho + c.oy0 + ry) * P.Wo + c.ox0 + rx;
        mvalid = c.mtile < m_tiles && ry < P.th && c.oy0 + ry < P.Ho && c.ox0 + rx < P.Wo;   // halo tiles may hang over in x too
      } else {
        m = c.m0 + gt;
        mvalid = m < P.M;
      }
      float alpha = 0.f, beta = 1.f;
      if (mode == 2) { alpha = __ldg(pr.alpha); beta = __ldg(pr.beta); }
      const float rbias = (row_bias && pr.bias && mvalid) ? __ldg(pr.bias + m) : 0.f;
      const int nb0 = c.n0 + half * kQCW;
      __half* yrow = pr.y + size_t(mvalid ? m : 0) * pr.y_ld + nb0;
      const __half* rrow = (mode != 0 && pr.res) ? pr.res + size_t(mvalid ? m : 0) * pr.res_ld + nb0 : nullptr;
      if (rrow && mvalid) {
        for (int cb = 0; cb < kQCW && nb0 + cb < P.N; cb += 64) prefetch_l2(rrow + cb);
      }
      const uintptr_t ua = reinterpret_cast<uintptr_t>(yrow) | (rrow ? reinterpret_cast<uintptr_t>(rrow) : 0);
      const int al_row = (ua & 31) == 0 ? 2 : ((ua & 15) == 0 ? 1 : 0);
      mbar_wait(tfull_bar(buf), it & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + uint32_t(buf * kQBN + half * kQCW) + lane_off;
      const int nrem = P.N - nb0;
      EpiRow ex;
      ex.sum = ex.sumsq = 0.f; ex.ln_a = 1.f; ex.ln_mu = 0.f; ex.ln_s = nullptr;
      if (XM) {
        ex.ln_s = pr.ln_s ? pr.ln_s + nb0 : nullptr;
        if (P.ln_parts > 0) epi_row_ln(ex, P, pr, m, mvalid);
        switch (mode_act) {
          case 9: epi_tile_pair<kQCW, 0, 0, 1>(trow, tempty_leader, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, ex); break;
          case 10: epi_tile_pair<kQCW, 2, 0, 1>(trow, tempty_leader, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, ex); break;
          default: epi_tile_pair<kQCW, 0, 2, 2>(trow, tempty_leader, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, ex); break;
        }
        if (mode_act == 11 && mvalid && nrem > 0) epi_row_emit(ex, P, pr, m, nb0, min(nb0 + kQCW, P.N));
      } else {
        switch (mode_act) {
          case 0: epi_tile_pair<kQCW, 0, 0>(trow, tempty_leader, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, ex); break;
          case 1: epi_tile_pair<kQCW, 0, 1>(trow, tempty_leader, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, ex); break;
          case 2: epi_tile_pair<kQCW, 0, 2>(trow, tempty_leader, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, ex); break;
          case 3: epi_tile_pair<kQCW, 1, 0>(trow, tempty_leader, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, ex); break;
          case 4: epi_tile_pair<kQCW, 1, 1>(trow, tempty_leader, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, ex); break;
          case 5: epi_tile_pair<kQCW, 1, 2>(trow, tempty_leader, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, ex); break;
          case 6: epi_tile_pair<kQCW, 2, 0>(trow, tempty_leader, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, ex); break;
          case 7: epi_tile_pair<kQCW, 2, 1>(trow, tempty_leader, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, ex); break;
          default: epi_tile_pair<kQCW, 2, 2>(trow, tempty_leader, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, ex); break;
        }
      }
    }
  } else if (warp == kQEpiWarps) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * BM, kQBN);
      int s = 0;
      uint32_t ph = 0;
      int zprev = -1, loads = 0;                             // halo mode 2: problem whose filter is resident, loads so far
      int i = 0;
      for (int t = cluster_id; t < total_pairs; t += n_clusters, ++i) {
        const int buf = i % kBufs;
        mbar_wait(tempty_bar(buf), ((i / kBufs) & 1) ^ 1);      // both CTAs' epilogue groups have drained this buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + uint32_t(buf * kQBN);
        // kst is 4 (Cin % 64 == 0), 2 or 1 (32- / 16-channel maps): fully unrolled issue sequences, no loop overhead in the
        // one thread that feeds the tensor pipe.  BKY = distance between the filter tiles of ky and ky + 1 (16-byte units).
#define ICAF_ISSUE_TAPS(KST, BKY)                                                                                            \
  _Pragma("unroll") for (int ky = 0; ky < 3; ++ky) {                                                                          \
    _Pragma("unroll") for (int k = 0; k < (KST); ++k)                                                                         \
        umma2_f16_ss(tmem_d, ad0 + uint64_t(ky * (1024 >> 4) + 2 * k), bd0 + uint64_t(ky * (BKY) + 2 * k), idesc,             \
                     !first || ky != 0 || k != 0);                                                                            \
  }
        if (halo == 2) {
          // resident filter: all nine half-tiles of this problem's filter sit at the bottom of shared memory (tap t at
          // t * kQBBytes, t = ky * 3 + kx); a ring stage is just one activation copy
          const int per_z = m_pairs * n_tiles;
          const int z = t / per_z;
          const int zn = t + n_clusters < total_pairs ? (t + n_clusters) / per_z : z;
          if (z != zprev) { mbar_wait(bres_full, uint32_t(loads & 1)); ++loads; zprev = z; }
          bool first = true;
          for (int kx = 0; kx < 3; ++kx) {
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t ad0 = umma_desc_sw128(smem_base + L::kResBytes + s * kQHaloABytes);
              const uint64_t bd0 = umma_desc_sw128(smem_base + kx * kQBBytes);
              if (kst == 4) { ICAF_ISSUE_TAPS(4, 3 * (kQBBytes >> 4)) } else if (kst == 2) { ICAF_ISSUE_TAPS(2, 3 * (kQBBytes >> 4)) }
              else { ICAF_ISSUE_TAPS(1, 3 * (kQBBytes >> 4)) }
              umma2_commit_both(empty_bar(s));
              if (kx == 2) {
                umma2_commit_both(tfull_bar(buf));
                if (zn != z) umma2_commit_both(bres_empty);   // last tile of this problem: the filter may be replaced
              }
            }
            __syncwarp();
            first = false;
            if (++s == L::kResStages) { s = 0; ph ^= 1; }
          }
          continue;
        }
        if (halo) {
          // one stage = one copy + the filter half-tiles of its three ky taps: a single barrier round trip and a single
          // commit per 12 MMAs (the issuing thread, not the tensor pipe, sets the pace of these loops)
          bool first = true;
          for (int cbk = 0; cbk < ncb; ++cbk) {
            for (int kx = 0; kx < 3; ++kx) {
              mbar_wait(full_bar(s), ph);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t sa = smem_base + s * L::kHaloStageBytes;
                const uint64_t ad0 = umma_desc_sw128(sa);
                const uint64_t bd0 = umma_desc_sw128(sa + kQHaloABytes);
                if (kst == 4) { ICAF_ISSUE_TAPS(4, kQBBytes >> 4) } else if (kst == 2) { ICAF_ISSUE_TAPS(2, kQBBytes >> 4) }
                else { ICAF_ISSUE_TAPS(1, kQBBytes >> 4) }
                umma2_commit_both(empty_bar(s));
                if (kx == 2 && cbk == ncb - 1) umma2_commit_both(tfull_bar(buf));
              }
              __syncwarp();
              first = false;
              if (++s == L::kHaloStages) { s = 0; ph ^= 1; }
            }
          }
          continue;
        }
#undef ICAF_ISSUE_TAPS
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(s), ph);                        // both CTAs' operands of this stage have landed
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_base + s * kQStageBytes;
            const uint64_t ad = umma_desc_sw128(sa);
            const uint64_t bd = umma_desc_sw128(sa + kQABytes);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma2_f16_ss(tmem_d, ad + uint64_t(2 * k), bd + uint64_t(2 * k), idesc, (kb | k) != 0);
            umma2_commit_both(empty_bar(s));
            if (kb == nkb - 1) umma2_commit_both(tfull_bar(buf));
          }
          __syncwarp();
          if (++s == kQStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ TMA producer (one thread per CTA)
    if (elect_one()) {
      const uint32_t a_bytes = a_mode == A_TMA2D ? uint32_t(kQABytes) : uint32_t(P.tw * P.th) * 128u;
      const uint32_t stage_tx = 2u * (uint32_t(kQBBytes) + a_bytes);   // both CTAs' loads complete on the leader's barrier
      int s = 0;
      uint32_t ph = 0;
      int zprev = -1, loads = 0;
      for (int t = cluster_id; t < total_pairs; t += n_clusters) {
        const PairTile c = pair_tile(P, t, rank, m_pairs, n_tiles, BN);
        const CUtensorMap* mw = c.z ? &maps.w[1] : &maps.w[0];
        const CUtensorMap* ma = c.z ? &maps.a[1] : &maps.a[0];
        if (halo == 2) {
          if (c.z != zprev) {                                // first tile, or the sequence moved on to the other stream's filter
            if (loads > 0) mbar_wait(bres_empty, uint32_t((loads - 1) & 1));
            if (rank == 0) mbar_arrive_expect_tx(bres_full, 2u * uint32_t(L::kResBytes));
            const uint32_t lb = map_to_cta(bres_full, 0);
            for (int tap = 0; tap < 9; ++tap)
              tma2_load_2d(smem_base + tap * kQBBytes, mw, lb, tap * P.Cin, c.n0 + rank * (kQBN / 2));
            ++loads;
            zprev = c.z;
          }
          for (int kx = 0; kx < 3; ++kx) {
            mbar_wait(empty_bar(s), ph ^ 1);
            if (rank == 0) mbar_arrive_expect_tx(full_bar(s), 2u * kQHaloABytes);
            tma2_load_4d(smem_base + L::kResBytes + s * kQHaloABytes, ma, map_to_cta(full_bar(s), 0), 0, c.ox0 - 1 + kx, c.oy0 - 1, c.tb);
            if (++s == L::kResStages) { s = 0; ph ^= 1; }
          }
          continue;
        }
        if (halo) {
          for (int cbk = 0; cbk < ncb; ++cbk) {
            for (int kx = 0; kx < 3; ++kx) {
              mbar_wait(empty_bar(s), ph ^ 1);
              const uint32_t sa = smem_base + s * L::kHaloStageBytes;
              const uint32_t lfull = map_to_cta(full_bar(s), 0);
              if (rank == 0) mbar_arrive_expect_tx(full_bar(s), 2u * uint32_t(L::kHaloStageBytes));
              tma2_load_4d(sa, ma, lfull, cbk * BK, c.ox0 - 1 + kx, c.oy0 - 1, c.tb);
              for (int ky = 0; ky < 3; ++ky)
                tma2_load_2d(sa + kQHaloABytes + ky * kQBBytes, mw, lfull, (ky * 3 + kx) * P.Cin + cbk * BK, c.n0 + rank * (kQBN / 2));
              if (++s == L::kHaloStages) { s = 0; ph ^= 1; }
            }
          }
          continue;
        }
        int ky = 0, kx = 0, ch = 0;                        // filter tap / channel block of the running K block (no divisions)
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t sa = smem_base + s * kQStageBytes;
          const uint32_t lfull = map_to_cta(full_bar(s), 0);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(s), stage_tx);
          tma2_load_2d(sa + kQABytes, mw, lfull, kb * BK, c.n0 + rank * (kQBN / 2));
          if (a_mode == A_TMA2D) {
            tma2_load_2d(sa, ma, lfull, kb * BK, c.m0);
          } else {
            tma2_load_4d(sa, ma, lfull, ch, c.ox0 * P.stride - P.pad + kx, c.oy0 * P.stride - P.pad + ky, c.tb);
            ch += BK;
            if (ch >= P.Cin) { ch = 0; if (++kx == P.kw) { kx = 0; ++ky; } }
          }
          if (++s == kQStages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  cluster_arrive();                                   // neither CTA may free TMEM / exit while the pair still works
  cluster_wait();
  if (warp == kQEpiWarps) {
    tc_fence_after();
    tmem_dealloc2<L::kTmemCols>(tmem_base);
  }
}

template <int BN>
int plan_pair(ConvParams& P, const icaf_conv_geom* g, int n_io, ConvPlan& pl) {
  using L = QSmem<BN>;
  if (!(P.a_mode == A_TMA2D || P.a_mode == A_TMA4D)) return set_error(ICAF_ERR_BAD_ARG, "conv2d(pair): both operands must arrive by TMA");
  const int m_tiles = P.a_mode == A_TMA4D ? P.B * P.tiles_x * P.tiles_y : (P.M + BM - 1) / BM;
  const int m_pairs = (m_tiles + 1) / 2;
  const int n_tiles = (P.N + BN - 1) / BN;
  const int total = m_pairs * n_tiles * n_io;
  // invariants the kernel's paths rely on (the dispatcher in conv_gemm.cu establishes them; fail loudly if it ever does not)
  if (P.a_mode == A_TMA4D && !(P.tw >= 1 && P.th >= 1 && P.tw * P.th <= BM && P.tiles_x * P.tw >= P.Wo && P.tiles_y * P.th >= P.Ho))
    return set_error(ICAF_ERR_BAD_ARG, "conv2d(pair): 4-D tiles must cover the map with at most 128 pixels each");
  if (P.halo && !(P.a_mode == A_TMA4D && g->kh == 3 && g->kw == 3 && g->stride == 1 && g->pad == 1 && P.tw == 8 && P.th == 16 &&
                  (P.Cin % 64 == 0 || P.Cin == 16 || P.Cin == 32)))
    return set_error(ICAF_ERR_BAD_ARG, "conv2d(pair): halo copies need a 3x3 / stride 1 / pad 1 layer on 16 x 8 tiles");
  if (!P.halo && P.a_mode == A_TMA4D && P.cblk != 64)
    return set_error(ICAF_ERR_BAD_ARG, "conv2d(pair): tap boxes need 64-channel blocks");
  if (P.halo == 2 && !(n_tiles == 1 && P.Cin <= 64))
    return set_error(ICAF_ERR_BAD_ARG, "conv2d(pair): the resident-filter mode needs one channel block and one N tile");
  P.stages = L::kStages;
  P.splits = 1;
  const int max_clusters = pl.sms / 2;
  if (max_clusters < 1) return set_error(ICAF_ERR_BAD_ARG, "conv2d(pair): needs at least two SMs");
  const int waves = (total + max_clusters - 1) / max_clusters;
  const int clusters = (total + waves - 1) / waves;
  pl.kernel = ICAF_KERNEL_PAIR; pl.bn = BN;
  pl.grid_x = unsigned(2 * clusters); pl.grid_y = pl.grid_z = 1; pl.cluster = 2;
  pl.smem = P.halo == 2 ? L::kResSmem : (P.halo ? L::kHaloSmem : L::kSmem);
  if (pl.smem > 227 * 1024) return set_error(ICAF_ERR_BAD_ARG, "conv2d(pair): shared-memory plan exceeds 227 KB");
  pl.total = total; pl.m_tiles = m_tiles; pl.m_pairs = m_pairs; pl.n_tiles = n_tiles;
  return ICAF_OK;
}

template <int BN>
int launch_pair(const ConvParams& P, const ConvPlan& pl, const __half* const (&w)[2], const icaf_conv_geom* g, int n_io, cudaStream_t st) {
  const bool xm = (P.epi & (ICAF_EPI_LN_FOLD | ICAF_EPI_EMIT_STATS)) != 0;
  static bool configured[2][kMaxDevices] = {{false}, {false}};
  if (int rc = xm ? configure_smem(conv_gemm_pair_kernel<BN, true>, 227 * 1024, configured[1], "conv2d: cudaFuncSetAttribute (pair)")
                  : configure_smem(conv_gemm_pair_kernel<BN, false>, 227 * 1024, configured[0], "conv2d: cudaFuncSetAttribute (pair)"))
    return rc;
  ConvMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int i = 0; i < n_io; ++i) {
    int rc = encode_tmap_2d(&maps.w[i], w[i], (uint64_t)P.k_pad, (uint64_t)g->w_rows, (uint64_t)P.k_pad * 2, BK, BN / 2);
    if (rc) return rc;
    const ConvProblem& pr = P.p[i];
    if (P.a_mode == A_TMA2D)
      rc = encode_tmap_2d(&maps.a[i], pr.x, (uint64_t)P.Cin, (uint64_t)P.M, (uint64_t)pr.x_ld * 2, BK, BM);
    else if (P.halo)   // one x-shifted copy: 8 px x (16+2) rows x 64 channels (a 16- / 32-channel map is zero-filled up to 64)
      rc = encode_tmap_nhwc(&maps.a[i], pr.x, P.Cin, P.Wi, P.Hi, P.B, pr.x_ld, BK, 8, 18, 1, 1);
    else
      rc = encode_tmap_nhwc(&maps.a[i], pr.x, P.Cin, P.Wi, P.Hi, P.B, pr.x_ld, BK, P.tw * P.stride, P.th * P.stride, P.stride, P.stride);
    if (rc) return rc;
  }
  if (n_io == 1) { maps.w[1] = maps.w[0]; maps.a[1] = maps.a[0]; }
  if (xm) launch_kc(conv_gemm_pair_kernel<BN, true>, dim3(pl.grid_x), dim3(kQThreads), (size_t)pl.smem, st, 2u, P, maps, pl.total, pl.m_tiles, pl.m_pairs, pl.n_tiles);
  else launch_kc(conv_gemm_pair_kernel<BN, false>, dim3(pl.grid_x), dim3(kQThreads), (size_t)pl.smem, st, 2u, P, maps, pl.total, pl.m_tiles, pl.m_pairs, pl.n_tiles);
  return check_launch("conv2d_fwd(pair)");
}

#define ICAF_INST(BN)                                                                                              \
  template int plan_pair<BN>(ConvParams&, const icaf_conv_geom*, int, ConvPlan&);                                  \
  template int launch_pair<BN>(const ConvParams&, const ConvPlan&, const __half* const (&)[2], const icaf_conv_geom*, int, cudaStream_t);
ICAF_INST(64)
ICAF_INST(128)
ICAF_INST(256)
#undef ICAF_INST

}  // namespace icaf
