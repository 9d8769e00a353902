// Persistent variant of the implicit-GEMM conv kernel for layers with many output tiles (>= ~2 tiles per SM).
//
// One CTA per SM loops over output tiles (static round-robin).  The three pipelines of conv_gemm.cu are kept but
// decoupled across tiles so that every unit stays busy:
//   * the TMA producer runs ahead through the tile sequence, bounded only by the smem ring;
//   * the MMA warp accumulates tile i into TMEM buffer (i mod kBufs; four buffers for BN <= 128, two for BN = 256) while
//   * the epilogue group(s) of that buffer -- four warps each; BN = 256 splits a tile's columns over two groups -- drain the
//     previous tile of the buffer (bias, SiLU/GELU, residual, fp16 store).  Four groups in total, so up to four tiles
//     are between MMA and store at once and the buffer hand-shake latency stays off the critical path:
//     the short-K layers are bound by the latency of that chain, not by a throughput limit (tools/conv_probe.py).
// Per-tile bookkeeping is kept off that chain: tile coordinates advance as a mixed-radix counter (no divisions), the
// bias is read straight from global memory (L1-resident across tiles), TMEM loads are double-buffered against the math,
// rows are written with 256-bit stores (full L2 sectors).
// 576 threads: warps 0-15 epilogue groups 0-3 (group = 2*buffer + column half), warp 16 MMA issuer + TMEM allocator,
// warp 17 TMA producer.  Both operands arrive by TMA; A_GATHER layers stay on conv_gemm_tc_kernel.
// TMEM: kBufs x BN fp32 columns.  Shared memory: the whole SM (ring of up to 10 stages).
#include <cstring>

#include "conv_common.cuh"

namespace icaf {

constexpr int kPEpiWarps = 16;
constexpr int kPThreads = (kPEpiWarps + 2) * 32;
constexpr int kPMaxStages = 10;

template <int BN>
struct PSmem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBufs = BN <= 128 ? 4 : 2;     // accumulator buffers in TMEM (tiles in flight between MMA and epilogue)
  static constexpr int kHalves = 4 / kBufs;           // epilogue groups per buffer: four groups in total
  static constexpr int kCW = BN / kHalves;            // columns per group
  static constexpr int kTmemCols = kBufs * BN;
  // [ring] [barriers 256 B] [bias: 4 groups x 2 tiles x kCW fp32] ; + 1024 B alignment slack
  static constexpr int kTailBytes = 256 + 4 * 2 * kCW * 4 + 1024;
  static int total(int stages) { return stages * kStageBytes + kTailBytes; }
};

// Tile sequence t, t + step, t + 2*step, ... as a mixed-radix counter (n fastest: CTAs working side by side share the
// activation tile in L2).  Digits: n tile | x | y | image | problem.  For 2-D A tiles x is the M tile and y, image are unused.
struct TileIter {
  int d[5], st[5], r[4];
  int t, step;
  __device__ __forceinline__ void init(const ConvParams& P, int t0, int step_, int n_tiles, int m_tiles) {
    const bool four = P.a_mode == A_TMA4D;
    r[0] = n_tiles; r[1] = four ? P.tiles_x : m_tiles; r[2] = four ? P.tiles_y : 1; r[3] = four ? P.B : 1;
    t = t0; step = step_;
    int a = t0, b = step_;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      d[i] = a % r[i]; a /= r[i];
      st[i] = b % r[i]; b /= r[i];
    }
    d[4] = a; st[4] = b;
  }
  __device__ __forceinline__ void next() {
    t += step;
    int carry = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int v = d[i] + st[i] + carry;
      carry = v >= r[i];
      d[i] = carry ? v - r[i] : v;
    }
    d[4] += st[4] + carry;
  }
};

struct TileCoord { int z, m0, n0, tb, oy0, ox0; };

__device__ __forceinline__ TileCoord tile_coord(const ConvParams& P, const TileIter& it, int BN) {
  TileCoord c;
  c.z = it.d[4]; c.n0 = it.d[0] * BN;
  if (P.a_mode == A_TMA4D) { c.m0 = 0; c.ox0 = it.d[1] * P.tw; c.oy0 = it.d[2] * P.th; c.tb = it.d[3]; }
  else { c.m0 = it.d[1] * BM; c.ox0 = 0; c.oy0 = 0; c.tb = 0; }
  return c;
}

// One epilogue group's share of a tile: CW accumulator columns of this thread's row in 16-column chunks.  The TMEM load
// of chunk i+1 is in flight during the math of chunk i (two register buffers, loop unrolled by two only: the hot loop
// must fit the instruction cache); the accumulator buffer goes back to the MMA warp as soon as the last chunk sits in
// registers.
template <int CW, int ACT, int RES, int XM = 0>
__device__ __forceinline__ void epi_tile(uint32_t trow, uint32_t tempty, const float* sb, float rbias, float alpha,
                                         float beta, const __half* rrow, __half* yrow, int al_row, bool mvalid, int nrem,
                                         bool do_store, EpiRow& ex) {
  static_assert(CW % 32 == 0, "two chunks per iteration");
  uint32_t acc0[16], acc1[16];
  auto chunk = [&](const uint32_t (&acc)[16], int cb) {
    const int nc = nrem - cb;
    if (mvalid && nc > 0)
      epi_chunk16<ACT, RES, XM>(acc, sb + cb, rbias, alpha, beta, rrow ? rrow + cb : nullptr, yrow + cb, nc >= 16 ? al_row : 0, nc,
                                do_store, ex, cb);
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
      if (lane_id() == 0) mbar_arrive(tempty);
    }
    chunk(acc1, cb + 16);
  }
}

template <int BN, bool XM>
__global__ void __launch_bounds__(kPThreads, 1)
conv_gemm_persist_kernel(const ConvParams P, const __grid_constant__ ConvMaps maps, int total_tiles, int m_tiles, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  using L = PSmem<BN>;
  constexpr int kBufs = L::kBufs, kHalves = L::kHalves, kCW = L::kCW;
  const int kStages = P.stages;
  const uint32_t bar_off = uint32_t(kStages) * L::kStageBytes;
  const uint32_t bar_base = smem_base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kPMaxStages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * kPMaxStages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * kPMaxStages + 4 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kPMaxStages + 8);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int tid = threadIdx.x;
  const int nkb = P.k_pad / BK;
  const int a_mode = P.a_mode;
  const uint32_t cblk = uint32_t(P.cblk);                 // channels per A box (64, or Cin for the small-Cin layers)
  const uint32_t sub_bytes = uint32_t(BM) * cblk * 2u;    // one per-tap sub-tile of the stage

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < kBufs; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 4 * kHalves);     // one elected arrival per epilogue warp of the buffer
    }
    fence_mbar_init();
  }
  if (warp == kPEpiWarps) tmem_alloc<L::kTmemCols>(tmem_slot);
  if (warp == kPEpiWarps + 1 && lane_id() == 0) {
    tma_prefetch_desc(&maps.w[0]);
    tma_prefetch_desc(&maps.w[1]);
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.a[1]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 8 * (2 * kPMaxStages + 8));

  if (warp < kPEpiWarps) {
    // ------------------------------------------------------------------ epilogue groups
    const int eg = warp >> 2;                      // group
    const int buf = eg / kHalves;                  // accumulator buffer
    const int half = eg % kHalves;                 // column half of the tile (BN = 256 only)
    {
      const int gt = tid & 127;                    // thread within the group == TMEM lane == tile row
      const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
      const int ry = a_mode == A_TMA4D ? gt / P.tw : 0;
      const int rx = a_mode == A_TMA4D ? gt - ry * P.tw : 0;
      const int mode = (P.epi & ICAF_EPI_SCALED_RES) ? 2 : ((P.epi & ICAF_EPI_ADD_RES) ? 1 : 0);
      // XM instantiation (DMFF linears): 9 / 10 = LayerNorm folded into this GEMM (no activation / GELU); 11 = scaled
      // residual + statistics of the output rows
      const int mode_act = XM ? (P.ln_parts > 0 ? (P.act == ICAF_ACT_GELU ? 10 : 9) : 11) : (ICAF_DBG(P, 2) ? mode : P.act * 3 + mode);
      const bool dst = !ICAF_DBG(P, 1);
      const bool row_bias = (P.epi & ICAF_EPI_BIAS_ROW) != 0;
      float* sbias = reinterpret_cast<float*>(smem_gen + bar_off + 256) + eg * 2 * kCW;   // [tile parity][kCW]
      // this thread's bias column of a tile (fetched one tile ahead, so its latency hides behind the current tile)
      auto bias_of = [&](const TileIter& q) -> float {
        const float* pb = q.d[4] ? P.p[1].bias : P.p[0].bias;
        const int n = q.d[0] * BN + half * kCW + gt;
        return (gt < kCW && pb && !row_bias && q.t < total_tiles && n < P.N) ? __ldg(pb + n) : 0.f;
      };
      TileIter ti;
      ti.init(P, blockIdx.x + buf * gridDim.x, kBufs * gridDim.x, n_tiles, m_tiles);
      float bnext = bias_of(ti);
      for (int it = 0; ti.t < total_tiles; ++it) {
        const TileCoord c = tile_coord(P, ti, BN);
        const ConvProblem pr = pick_problem(P, c.z);
        float* sb = sbias + (it & 1) * kCW;
        if (gt < kCW) sb[gt] = bnext;
        named_bar_sync(1 + eg, 128);               // slice (it & 1) was last read two tiles ago: no second barrier needed
        ti.next();
        bnext = bias_of(ti);
        int m;
        bool mvalid;
        if (a_mode == A_TMA4D) {
          m = (c.tb * P.Ho + c.oy0 + ry) * P.Wo + c.ox0 + rx;
          mvalid = ry < P.th && c.oy0 + ry < P.Ho;
        } else {
          m = c.m0 + gt;
          mvalid = m < P.M;
        }
        float alpha = 0.f, beta = 1.f;
        if (mode == 2) { alpha = __ldg(pr.alpha); beta = __ldg(pr.beta); }
        const float rbias = (row_bias && pr.bias && mvalid) ? __ldg(pr.bias + m) : 0.f;
        const int nb0 = c.n0 + half * kCW;
        __half* yrow = pr.y + size_t(mvalid ? m : 0) * pr.y_ld + nb0;
        const __half* rrow = (mode != 0 && pr.res) ? pr.res + size_t(mvalid ? m : 0) * pr.res_ld + nb0 : nullptr;
        if (rrow && mvalid) {
          for (int cb = 0; cb < kCW && nb0 + cb < P.N; cb += 64) prefetch_l2(rrow + cb);
        }
        // alignment class of this row's 16-column chunks (chunks are 32 bytes apart, so one test covers them all)
        const uintptr_t ua = reinterpret_cast<uintptr_t>(yrow) | (rrow ? reinterpret_cast<uintptr_t>(rrow) : 0);
        const int al_row = (ua & 31) == 0 ? 2 : ((ua & 15) == 0 ? 1 : 0);
        mbar_wait(tfull_bar(buf), it & 1);
        tc_fence_after();
        const uint32_t trow = tmem_base + uint32_t(buf * BN + half * kCW) + lane_off;
        const uint32_t te = tempty_bar(buf);
        const int nrem = P.N - nb0;
        EpiRow ex;
        ex.sum = ex.sumsq = 0.f; ex.ln_a = 1.f; ex.ln_mu = 0.f; ex.ln_s = nullptr;
        if (XM) {
          ex.ln_s = pr.ln_s ? pr.ln_s + nb0 : nullptr;
          if (P.ln_parts > 0) epi_row_ln(ex, P, pr, m, mvalid);
          switch (mode_act) {
            case 9: epi_tile<kCW, 0, 0, 1>(trow, te, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, dst, ex); break;
            case 10: epi_tile<kCW, 2, 0, 1>(trow, te, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, dst, ex); break;
            default: epi_tile<kCW, 0, 2, 2>(trow, te, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, dst, ex); break;
          }
          if (mode_act == 11 && mvalid && nrem > 0) epi_row_emit(ex, P, pr, m, nb0, min(nb0 + kCW, P.N));
        } else {
          switch (mode_act) {
            case 0: epi_tile<kCW, 0, 0>(trow, te, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, dst, ex); break;
            case 1: epi_tile<kCW, 0, 1>(trow, te, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, dst, ex); break;
            case 2: epi_tile<kCW, 0, 2>(trow, te, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, dst, ex); break;
            case 3: epi_tile<kCW, 1, 0>(trow, te, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, dst, ex); break;
            case 4: epi_tile<kCW, 1, 1>(trow, te, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, dst, ex); break;
            case 5: epi_tile<kCW, 1, 2>(trow, te, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, dst, ex); break;
            case 6: epi_tile<kCW, 2, 0>(trow, te, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, dst, ex); break;
            case 7: epi_tile<kCW, 2, 1>(trow, te, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, dst, ex); break;
            default: epi_tile<kCW, 2, 2>(trow, te, sb, rbias, alpha, beta, rrow, yrow, al_row, mvalid, nrem, dst, ex); break;
          }
        }
      }
    }
  } else if (warp == kPEpiWarps) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
    int s = 0;
    uint32_t ph = 0;
    int i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int buf = i % kBufs;
      mbar_wait(tempty_bar(buf), ((i / kBufs) & 1) ^ 1);      // epilogue groups have drained this buffer (first use: free)
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + uint32_t(buf * BN);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_base + s * L::kStageBytes;
          const uint64_t bd = umma_desc_sw128(sa + L::kABytes);
          if (ICAF_DBG(P, 32)) {
          } else if (cblk == 64) {
            const uint64_t ad = umma_desc_sw128(sa);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16_ss(tmem_d, ad + uint64_t(2 * k), bd + uint64_t(2 * k), idesc, (kb | k) != 0);
          } else {
            // small-Cin staging: the K block is 64/cblk per-tap sub-tiles [128 rows][cblk] with 32B / 64B swizzle
            const int nks = min(BK / 16, (P.K - kb * BK) / 16);
            for (int k = 0; k < nks; ++k) {
              const uint32_t e = uint32_t(k * 16);
              const uint64_t ad = umma_desc_kmajor(sa + (e / cblk) * sub_bytes + (e % cblk) * 2u, cblk * 2u);
              umma_f16_ss(tmem_d, ad, bd + uint64_t(2 * k), idesc, (kb | k) != 0);
            }
          }
          umma_commit(empty_bar(s));
          if (kb == nkb - 1) umma_commit(tfull_bar(buf));
        }
        __syncwarp();
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == kPEpiWarps + 1) {
    // ------------------------------------------------------------------ TMA producer (one thread)
    if (elect_one()) {
      const uint32_t a_bytes = a_mode == A_TMA2D ? L::kABytes : (a_mode == A_TMA4D ? uint32_t(P.tw * P.th) * 128u : 0u);
      int s = 0;
      uint32_t ph = 0;
      TileIter ti;
      ti.init(P, blockIdx.x, gridDim.x, n_tiles, m_tiles);
      for (; ti.t < total_tiles; ti.next()) {
        const TileCoord c = tile_coord(P, ti, BN);
        const CUtensorMap* mw = c.z ? &maps.w[1] : &maps.w[0];
        const CUtensorMap* ma = c.z ? &maps.a[1] : &maps.a[0];
        int ky = 0, kx = 0, ch = 0;                        // tap / channel block of the running K block (64-channel path)
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t sa = smem_base + s * L::kStageBytes;
          if (a_mode == A_TMA4D && cblk < 64) {
            // one box per filter tap inside this K block; the tail block of K = 9*Cin holds fewer taps
            const int nsub = min(int(BK / cblk), (P.K - kb * BK) / int(cblk));
            mbar_arrive_expect_tx(full_bar(s), L::kBBytes + uint32_t(nsub) * uint32_t(P.tw * P.th) * cblk * 2u);
            tma_load_2d(sa + L::kABytes, mw, full_bar(s), kb * BK, c.n0);
            for (int j = 0; j < nsub; ++j) {
              const int k0 = kb * BK + j * int(cblk);
              const int tap = k0 / P.Cin;
              const int ch = k0 - tap * P.Cin;
              const int ky = tap / P.kw, kx = tap - ky * P.kw;
              tma_load_4d(sa + uint32_t(j) * sub_bytes, ma, full_bar(s), ch, c.ox0 * P.stride - P.pad + kx,
                          c.oy0 * P.stride - P.pad + ky, c.tb);
            }
            if (++s == kStages) { s = 0; ph ^= 1; }
            continue;
          }
          const bool ldA = !ICAF_DBG(P, 8), ldB = !ICAF_DBG(P, 16);
          mbar_arrive_expect_tx(full_bar(s), (ldB ? L::kBBytes : 0u) + (ldA ? a_bytes : 0u));
          if (ldB) tma_load_2d(sa + L::kABytes, mw, full_bar(s), kb * BK, c.n0);
          if (!ldA) {
          } else if (a_mode == A_TMA2D) {
            tma_load_2d(sa, ma, full_bar(s), kb * BK, c.m0);
          } else if (a_mode == A_TMA4D) {
            tma_load_4d(sa, ma, full_bar(s), ch, c.ox0 * P.stride - P.pad + kx, c.oy0 * P.stride - P.pad + ky, c.tb);
          }
          if (a_mode == A_TMA4D) {                         // advance even when the probe build skipped the load
            ch += BK;
            if (ch >= P.Cin) { ch = 0; if (++kx == P.kw) { kx = 0; ++ky; } }
          }
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kPEpiWarps) {
    tc_fence_after();
    tmem_dealloc<L::kTmemCols>(tmem_base);
  }
}

template <int BN>
int plan_persist(ConvParams& P, int n_io, ConvPlan& pl) {
  using L = PSmem<BN>;
  constexpr int kSmemCap = 227 * 1024;
  if (P.a_mode == A_GATHER) return set_error(ICAF_ERR_BAD_ARG, "conv2d(persistent): both operands must arrive by TMA");
  const int m_tiles = P.a_mode == A_TMA4D ? P.B * P.tiles_x * P.tiles_y : (P.M + BM - 1) / BM;
  const int n_tiles = (P.N + BN - 1) / BN;
  const int total = m_tiles * n_tiles * n_io;
  if (P.a_mode == A_TMA4D && !(P.tw >= 1 && P.th >= 1 && P.tw * P.th <= BM && P.tiles_x * P.tw >= P.Wo && P.tiles_y * P.th >= P.Ho &&
                               (P.cblk == 64 || P.cblk == 32 || P.cblk == 16) && P.Cin % P.cblk == 0))
    return set_error(ICAF_ERR_BAD_ARG, "conv2d(persistent): 4-D tiles must cover the map with at most 128 pixels each");
  int stages = (kSmemCap - L::kTailBytes) / L::kStageBytes;
  if (stages > kPMaxStages) stages = kPMaxStages;
  if (stages < 2) return set_error(ICAF_ERR_BAD_ARG, "conv2d(persistent): ring shallower than two stages");
  P.stages = stages;
  P.splits = 1;
  // balanced static schedule: every CTA gets ceil(total/waves) or one fewer tiles
  const int waves = (total + pl.sms - 1) / pl.sms;
  pl.kernel = ICAF_KERNEL_PERSIST; pl.bn = BN;
  pl.grid_x = unsigned((total + waves - 1) / waves); pl.grid_y = pl.grid_z = 1; pl.cluster = 1;
  pl.smem = L::total(stages);
  pl.total = total; pl.m_tiles = m_tiles; pl.m_pairs = 0; pl.n_tiles = n_tiles;
  return ICAF_OK;
}

template <int BN>
int launch_persist(const ConvParams& P, const ConvPlan& pl, const __half* const (&w)[2], const icaf_conv_geom* g, int n_io, cudaStream_t st) {
  const bool xm = (P.epi & (ICAF_EPI_LN_FOLD | ICAF_EPI_EMIT_STATS)) != 0;
  static bool configured[2][kMaxDevices] = {{false}, {false}};
  if (int rc = xm ? configure_smem(conv_gemm_persist_kernel<BN, true>, 227 * 1024, configured[1], "conv2d: cudaFuncSetAttribute (persistent)")
                  : configure_smem(conv_gemm_persist_kernel<BN, false>, 227 * 1024, configured[0], "conv2d: cudaFuncSetAttribute (persistent)"))
    return rc;
  ConvMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int i = 0; i < n_io; ++i) {
    int rc = encode_tmap_2d(&maps.w[i], w[i], (uint64_t)P.k_pad, (uint64_t)g->w_rows, (uint64_t)P.k_pad * 2, BK, BN);
    if (rc) return rc;
    const ConvProblem& pr = P.p[i];
    if (P.a_mode == A_TMA2D)
      rc = encode_tmap_2d(&maps.a[i], pr.x, (uint64_t)P.Cin, (uint64_t)P.M, (uint64_t)pr.x_ld * 2, BK, BM);
    else if (P.a_mode == A_TMA4D)
      rc = encode_tmap_nhwc(&maps.a[i], pr.x, P.Cin, P.Wi, P.Hi, P.B, pr.x_ld, (uint32_t)P.cblk, P.tw * P.stride, P.th * P.stride,
                            P.stride, P.stride);
    if (rc) return rc;
  }
  if (n_io == 1) { maps.w[1] = maps.w[0]; maps.a[1] = maps.a[0]; }
  if (xm) launch_k(conv_gemm_persist_kernel<BN, true>, dim3(pl.grid_x), dim3(kPThreads), (size_t)pl.smem, st, P, maps, pl.total, pl.m_tiles, pl.n_tiles);
  else launch_k(conv_gemm_persist_kernel<BN, false>, dim3(pl.grid_x), dim3(kPThreads), (size_t)pl.smem, st, P, maps, pl.total, pl.m_tiles, pl.n_tiles);
  return check_launch("conv2d_fwd(persistent)");
}

#define ICAF_INST(BN)                                                                                              \
  template int plan_persist<BN>(ConvParams&, int, ConvPlan&);                                                      \
  template int launch_persist<BN>(const ConvParams&, const ConvPlan&, const __half* const (&)[2], const icaf_conv_geom*, int, cudaStream_t);
ICAF_INST(32)
ICAF_INST(64)
ICAF_INST(128)
ICAF_INST(256)
#undef ICAF_INST

}  // namespace icaf
