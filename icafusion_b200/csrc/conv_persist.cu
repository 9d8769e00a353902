// Persistent variant of the implicit-GEMM conv kernel for layers with many output tiles (>= ~2 tiles per SM).
//
// One CTA per SM loops over output tiles (static round-robin).  The three pipelines of conv_gemm.cu are kept but
// decoupled across tiles so that every unit stays busy:
//   * the TMA / gather producers run ahead through the tile sequence, bounded only by the smem ring;
//   * the MMA warp accumulates tile i into TMEM buffer (i & 1) while
//   * epilogue group (i & 1) -- four warps -- drains the previous tile of that buffer (bias, SiLU/GELU, residual,
//     fp16 store).  Two epilogue groups alternate, so two tiles can be in their epilogue while a third is in the tensor
//     core: for the 1x1 / small-K layers, whose main loop is shorter than their epilogue, that is what lets the kernel
//     run at the HBM roofline instead of at the latency of one CTA's serial phases.
// 448 threads: warps 0-3 epilogue group 0, warps 4-7 epilogue group 1, warp 8 MMA issuer + TMEM allocator,
// warp 9 TMA producer, warps 10-13 cp.async gather producers (A_GATHER layers only).
// TMEM: 2 x BN fp32 columns.  Shared memory: the whole SM (ring of up to 10 stages).
#include <cstring>

#include "conv_common.cuh"

namespace icaf {

constexpr int kPThreads = 448;
constexpr int kPMaxStages = 10;

template <int BN>
struct PSmem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  // [ring] [barriers 256 B] [bias 2 x BN fp32] ; + 1024 B alignment slack
  static constexpr int kTailBytes = 256 + 2 * BN * 4 + 1024;
  static int total(int stages) { return stages * kStageBytes + kTailBytes; }
};

struct TileCoord { int z, m0, n0, tb, oy0, ox0; };

__device__ __forceinline__ TileCoord tile_coord(const ConvParams& P, int t, int n_tiles, int m_tiles, int BN) {
  TileCoord c;
  const int per_z = m_tiles * n_tiles;
  c.z = t / per_z;
  t -= c.z * per_z;
  const int mt = t / n_tiles;            // n fastest: CTAs working side by side share the activation tile in L2
  c.n0 = (t - mt * n_tiles) * BN;
  c.m0 = mt * BM; c.tb = 0; c.oy0 = 0; c.ox0 = 0;
  if (P.a_mode == A_TMA4D) {
    const int per_img = P.tiles_x * P.tiles_y;
    c.tb = mt / per_img;
    const int r = mt - c.tb * per_img;
    c.oy0 = (r / P.tiles_x) * P.th;
    c.ox0 = (r % P.tiles_x) * P.tw;
    c.m0 = 0;
  }
  return c;
}

template <int BN>
__global__ void __launch_bounds__(kPThreads, 1)
conv_gemm_persist_kernel(const ConvParams P, const __grid_constant__ ConvMaps maps, int total_tiles, int m_tiles, int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  using L = PSmem<BN>;
  const int kStages = P.stages;
  const uint32_t bar_off = uint32_t(kStages) * L::kStageBytes;
  const uint32_t bar_base = smem_base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kPMaxStages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * kPMaxStages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * kPMaxStages + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kPMaxStages + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int tid = threadIdx.x;
  const int nkb = P.k_pad / BK;
  const int a_mode = P.a_mode;
  const uint32_t cblk = uint32_t(P.cblk);                 // channels per A box (64, or Cin for the small-Cin layers)
  const uint32_t sub_bytes = uint32_t(BM) * cblk * 2u;    // one per-tap sub-tile of the stage

  if (tid == 0) {
    const uint32_t nfull = a_mode == A_GATHER ? 129u : 1u;
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), nfull);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 128);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<L::kTmemCols>(tmem_slot);
  if (warp == 9 && lane_id() == 0) {
    tma_prefetch_desc(&maps.w[0]);
    tma_prefetch_desc(&maps.w[1]);
    if (a_mode != A_GATHER) { tma_prefetch_desc(&maps.a[0]); tma_prefetch_desc(&maps.a[1]); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 8 * (2 * kPMaxStages + 4));

  if (warp < 8) {
    // ------------------------------------------------------------------ epilogue groups
    const int g = warp >> 2;                       // group = accumulator buffer
    const int gt = tid & 127;                      // thread within the group == TMEM lane == tile row
    float* sbias = reinterpret_cast<float*>(smem_gen + bar_off + 256) + g * BN;
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    int it = 0;                                    // tiles this group has drained
    for (int i = g, t = blockIdx.x + g * gridDim.x; t < total_tiles; i += 2, t += 2 * gridDim.x, ++it) {
      const TileCoord c = tile_coord(P, t, n_tiles, m_tiles, BN);
      const ConvProblem pr = pick_problem(P, c.z);
      int m;
      bool mvalid;
      if (a_mode == A_TMA4D) {
        const int ry = gt / P.tw, rx = gt - ry * P.tw;
        m = (c.tb * P.Ho + c.oy0 + ry) * P.Wo + c.ox0 + rx;
        mvalid = ry < P.th && c.oy0 + ry < P.Ho;
      } else {
        m = c.m0 + gt;
        mvalid = m < P.M;
      }
      float alpha = 0.f, beta = 1.f;
      if (P.epi & ICAF_EPI_SCALED_RES) { alpha = __ldg(pr.alpha); beta = __ldg(pr.beta); }
      const float rbias = ((P.epi & ICAF_EPI_BIAS_ROW) && pr.bias && mvalid) ? __ldg(pr.bias + m) : 0.f;
      __half* yrow = pr.y + size_t(mvalid ? m : 0) * pr.y_ld;
      const __half* rrow = pr.res ? pr.res + size_t(mvalid ? m : 0) * pr.res_ld : nullptr;
      const int mode = (P.epi & ICAF_EPI_SCALED_RES) ? 2 : (rrow ? 1 : 0);
      if (rrow && mvalid) {
        for (int cb = 0; cb < BN && c.n0 + cb < P.N; cb += 64) prefetch_l2(rrow + c.n0 + cb);
      }
      named_bar_sync(1 + g, 128);                  // previous tile's readers are done with this group's bias slice
      for (int col = gt; col < BN; col += 128)
        sbias[col] = (pr.bias && !(P.epi & ICAF_EPI_BIAS_ROW) && c.n0 + col < P.N) ? __ldg(pr.bias + c.n0 + col) : 0.f;
      named_bar_sync(1 + g, 128);
      mbar_wait(tfull_bar(g), it & 1);
      tc_fence_after();
      const uint32_t trow = tmem_base + uint32_t(g * BN) + lane_off;
      const int mode_act = P.act * 3 + mode;
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 32) {
        uint32_t acc[32];
        __syncwarp();
        tmem_ld32(trow + cb, acc);
        tmem_ld_wait();
        const int nb = c.n0 + cb;
        if (mvalid && nb < P.N) {
          const int ncols = min(32, P.N - nb);
          const bool vec = ncols == 32 && ((reinterpret_cast<uintptr_t>(yrow + nb) & 15) == 0) &&
                           (!rrow || (reinterpret_cast<uintptr_t>(rrow + nb) & 15) == 0);
          const float* sb = sbias + cb;
          const __half* rp = rrow ? rrow + nb : nullptr;
          __half* yp = yrow + nb;
          switch (mode_act) {
            case 0: epi_chunk<0, 0>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols); break;
            case 1: epi_chunk<0, 1>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols); break;
            case 2: epi_chunk<0, 2>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols); break;
            case 3: epi_chunk<1, 0>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols); break;
            case 4: epi_chunk<1, 1>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols); break;
            case 5: epi_chunk<1, 2>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols); break;
            case 6: epi_chunk<2, 0>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols); break;
            case 7: epi_chunk<2, 1>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols); break;
            default: epi_chunk<2, 2>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols); break;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(g));                  // 128 arrivals: the MMA warp may overwrite this accumulator buffer
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
    int s = 0;
    uint32_t ph = 0;
    int i = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
      const int buf = i & 1;
      mbar_wait(tempty_bar(buf), ((i >> 1) & 1) ^ 1);      // epilogue group has drained this buffer (first use: free)
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + uint32_t(buf * BN);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_base + s * L::kStageBytes;
          const uint64_t bd = umma_desc_sw128(sa + L::kABytes);
          if (cblk == 64) {
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
  } else if (warp == 9) {
    // ------------------------------------------------------------------ TMA producer (one thread)
    if (elect_one()) {
      const uint32_t a_bytes = a_mode == A_TMA2D ? L::kABytes : (a_mode == A_TMA4D ? uint32_t(P.tw * P.th) * 128u : 0u);
      const uint32_t bytes = L::kBBytes + a_bytes;
      int s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileCoord c = tile_coord(P, t, n_tiles, m_tiles, BN);
        const CUtensorMap* mw = c.z ? &maps.w[1] : &maps.w[0];
        const CUtensorMap* ma = c.z ? &maps.a[1] : &maps.a[0];
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
          mbar_arrive_expect_tx(full_bar(s), bytes);
          tma_load_2d(sa + L::kABytes, mw, full_bar(s), kb * BK, c.n0);
          if (a_mode == A_TMA2D) {
            tma_load_2d(sa, ma, full_bar(s), kb * BK, c.m0);
          } else if (a_mode == A_TMA4D) {
            const int k0 = kb * BK;
            const int tap = k0 / P.Cin;
            const int ch = k0 - tap * P.Cin;
            const int ky = tap / P.kw, kx = tap - ky * P.kw;
            tma_load_4d(sa, ma, full_bar(s), ch, c.ox0 * P.stride - P.pad + kx, c.oy0 * P.stride - P.pad + ky, c.tb);
          }
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (a_mode == A_GATHER) {
    // ------------------------------------------------------------------ cp.async gather producers (warps 10-13)
    const int pt = tid - 320;         // 0..127
    const int c8 = pt & 7;            // 16-byte chunk within the 128-byte K row
    const int r0 = pt >> 3;           // rows r0 + 16*i
    const uint32_t sw = uint32_t(c8 ^ (r0 & 7)) << 4;
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileCoord c = tile_coord(P, t, n_tiles, m_tiles, BN);
      const ConvProblem pr = pick_problem(P, c.z);
      uint32_t base[8];
      int iy0[8], ix0[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int m = c.m0 + r0 + 16 * i;
        bool mv = m < P.M;
        int mm = mv ? m : 0;
        int ox = mm % P.Wo;
        int q = mm / P.Wo;
        int oy = q % P.Ho;
        int b = q / P.Ho;
        base[i] = uint32_t(b) * uint32_t(P.Hi * P.Wi);
        iy0[i] = mv ? oy * P.stride - P.pad : -100000;
        ix0[i] = ox * P.stride - P.pad;
      }
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t sa = smem_base + s * L::kStageBytes;
        const int k0 = kb * BK + c8 * 8;
        const bool kvalid = k0 < P.K;
        const int tap = k0 / P.Cin;
        const int ch = k0 - tap * P.Cin;
        const int ky = tap / P.kw;
        const int kx = tap - ky * P.kw;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int iy = iy0[i] + ky, ix = ix0[i] + kx;
          bool ok = kvalid && (unsigned)iy < (unsigned)P.Hi && (unsigned)ix < (unsigned)P.Wi;
          size_t off = ok ? (size_t(base[i] + uint32_t(iy * P.Wi + ix)) * size_t(pr.x_ld) + ch) : 0;
          cp_async16(sa + uint32_t(r0 + 16 * i) * 128u + sw, pr.x + off, ok);
        }
        cp_async_arrive_on(full_bar(s));     // asynchronous arrival: the whole ring can be in flight, nobody blocks
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<L::kTmemCols>(tmem_base);
  }
}

template <int BN>
int launch_persist(ConvParams& P, const __half* const (&w)[2], const icaf_conv_geom* g, int n_io, cudaStream_t st) {
  using L = PSmem<BN>;
  constexpr int kSmemCap = 227 * 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_persist_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemCap);
    if (e != cudaSuccess) return set_cuda_error(e, "conv2d: cudaFuncSetAttribute (persistent)");
    configured = true;
  }
  const int m_tiles = P.a_mode == A_TMA4D ? P.B * P.tiles_x * P.tiles_y : (P.M + BM - 1) / BM;
  const int n_tiles = (P.N + BN - 1) / BN;
  const int total = m_tiles * n_tiles * n_io;
  int stages = (kSmemCap - L::kTailBytes) / L::kStageBytes;
  if (stages > kPMaxStages) stages = kPMaxStages;
  P.stages = stages;
  P.splits = 1;
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
  const int sms = sm_count_cached();
  // balanced static schedule: every CTA gets ceil(total/waves) or one fewer tiles
  const int waves = (total + sms - 1) / sms;
  const int grid = (total + waves - 1) / waves;
  launch_k(conv_gemm_persist_kernel<BN>, dim3(grid), dim3(kPThreads), (size_t)L::total(stages), st, P, maps, total, m_tiles, n_tiles);
  return check_launch("conv2d_fwd(persistent)");
}

template int launch_persist<32>(ConvParams&, const __half* const (&)[2], const icaf_conv_geom*, int, cudaStream_t);
template int launch_persist<64>(ConvParams&, const __half* const (&)[2], const icaf_conv_geom*, int, cudaStream_t);
template int launch_persist<128>(ConvParams&, const __half* const (&)[2], const icaf_conv_geom*, int, cudaStream_t);
template int launch_persist<256>(ConvParams&, const __half* const (&)[2], const icaf_conv_geom*, int, cudaStream_t);

}  // namespace icaf
