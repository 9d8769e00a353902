// Implicit-GEMM Conv(+folded BN)+bias+activation(+residual) / Linear on the 5th-gen tensor cores (tcgen05).
//
//   C[M=B*Ho*Wo, N=Cout] = A[M, K=kh*kw*Cin] * W[N, K]^T           fp16 operands, fp32 accumulate in TMEM
//
// A is never materialised (no im2col).  Three ways to stage the A tile (128 rows x 64 K) into the 128B-swizzled
// K-major shared-memory layout the UMMA descriptors expect, picked per layer on the host:
//   A_TMA2D  : 1x1 conv / linear -- A is a plain [M][Cin] matrix: one 2-D TMA box per stage.
//   A_TMA4D  : kxk conv whose 128-row tile is a TH x TW patch of one image and Cin % 64 == 0: one 4-D TMA box
//              (64 ch, TW*s, TH*s, 1) of the NHWC input per stage at coordinates shifted by the filter tap;
//              out-of-bounds (= padding) is zero-filled by the TMA unit, the stride is the box traversal stride.
//              (Cin = 16 / 32: one box per filter tap, 32- / 64-byte swizzle -- persistent and pair kernels only.)
//   A_GATHER : anything else (Cin = 4 / 8 / odd multiples of 8, Cin < 64 on small grids): producer warps gather 16-byte
//              channel runs with cp.async (zero-fill), 4 rows x 128 B per warp instruction.
// The filter tile (BN rows x 64 K) always arrives by 2-D TMA.
// Three kernels share this contract; icaf_conv2d_fwd picks per layer:
//   conv_gemm_tc_kernel      (this file)     one tile per CTA, split-K over clusters     -- grids below ~2 tiles per SM
//   conv_gemm_persist_kernel (conv_persist.cu) one CTA per SM looping over tiles         -- many tiles, short K
//   conv_gemm_pair_kernel    (conv_pair.cu)  CTA pairs (cta_group::2), halo copies for 3x3 -- wide / deep-K layers
// conv_gemm_tc_kernel: CTA = one 128 x BN output tile, 192 threads:
//   warps 0-3 : A_GATHER producers, then the epilogue (thread t owns TMEM lane t = output row t)
//   warp  4   : TMEM allocator + single-thread tcgen05.mma issuer
//   warp  5   : TMA producer (one elected thread)
// Pipelines: smem ring full[]/empty[] (producers <-> MMA), accum_full (MMA -> epilogue).  Two CTAs share an SM
// (BN <= 128) so one CTA's epilogue overlaps the other's main loop.
#include <cstdlib>
#include <cstring>

#include "conv_common.cuh"

namespace icaf {

constexpr int kMaxStages = 10;
template <int BN>
struct SmemLayout {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  // [ring: stages x (A|B)] [barriers 256 B] [bias BN fp32] ; + 1024 B alignment slack
  static constexpr int kTailBytes = 256 + BN * 4 + 1024;
  static int total(int stages) { return stages * kStageBytes + kTailBytes; }
};

// XM: the LayerNorm-fold / row-statistics epilogues (DMFF linears) live in their own instantiation so that the epilogue
// of every other layer keeps its round-1 size (the hot loops are instruction-cache sensitive).
template <int BN, bool XM>
__global__ void __launch_bounds__(kThreads, (BN <= 128 ? 2 : 1))
conv_gemm_tc_kernel(const ConvParams P, const __grid_constant__ ConvMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  using L = SmemLayout<BN>;
  const int kStages = P.stages;
  const uint32_t bar_off = uint32_t(kStages) * L::kStageBytes;
  const uint32_t bar_base = smem_base + bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * kMaxStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 1);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int tid = threadIdx.x;
  // split-K: the `splits` CTAs of a cluster (consecutive blockIdx.x) share one output tile and take a K range each
  const int splits = P.splits;
  const int mtile = splits > 1 ? blockIdx.x / splits : blockIdx.x;
  const int ntile = blockIdx.y, bz = blockIdx.z;
  const ConvProblem pr = pick_problem(P, bz);           // by value: a dynamic param index would spill to local
  const int n0 = ntile * BN;
  const int a_mode = P.a_mode;
  const uint32_t crank = splits > 1 ? cluster_ctarank() : 0u;
  const int nkb_all = P.k_pad / BK;
  const int kb_begin = (nkb_all * int(crank)) / splits;
  const int nkb = (nkb_all * (int(crank) + 1)) / splits - kb_begin;
  // tile origin: linear rows (gather / 2-D) or a th x tw patch of image tb (4-D)
  int m0 = mtile * BM, tb = 0, oy0 = 0, ox0 = 0;
  if (a_mode == A_TMA4D) {
    const int per_img = P.tiles_x * P.tiles_y;
    m0 = 0;
    tb = mtile / per_img;
    const int t = mtile - tb * per_img;
    oy0 = (t / P.tiles_x) * P.th;
    ox0 = (t % P.tiles_x) * P.tw;
  }

  if (tid == 0) {
    const uint32_t nfull = a_mode == A_GATHER ? 129u : 1u;    // 128 gather threads + the TMA thread
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), nfull);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<L::kTmemCols>(tmem_slot);
  if (warp == 5 && lane_id() == 0) {
    tma_prefetch_desc(bz ? &maps.w[1] : &maps.w[0]);
    if (a_mode != A_GATHER) tma_prefetch_desc(bz ? &maps.a[1] : &maps.a[0]);
  }
  float* sbias = reinterpret_cast<float*>(smem_gen + bar_off + 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();   // prologue (barriers, TMEM, descriptor prefetch, bias = parameters only) overlapped the previous kernel
  const uint32_t tmem_d = *reinterpret_cast<volatile uint32_t*>(smem_gen + bar_off + 8 * (2 * kMaxStages + 1));

  if (warp < 4) {
    // Epilogue operands that do not depend on the main loop are fetched now so their DRAM latency hides behind it:
    // bias slice and (alpha, beta) into registers, this thread's residual row into L2.
    float bias_r[(BN + 127) / 128];
#pragma unroll
    for (int i = 0; i < (BN + 127) / 128; ++i) {
      const int col = tid + 128 * i;
      bias_r[i] = (col < BN && pr.bias && !(P.epi & ICAF_EPI_BIAS_ROW) && n0 + col < P.N) ? __ldg(pr.bias + n0 + col) : 0.f;
    }
    float alpha = 0.f, beta = 1.f;
    if (P.epi & ICAF_EPI_SCALED_RES) { alpha = __ldg(pr.alpha); beta = __ldg(pr.beta); }
    if (a_mode == A_GATHER) {
      // ---------------------------------------------------------------- cp.async gather producers
      const int c = tid & 7;          // 16-byte chunk within the 128-byte K row
      const int r0 = tid >> 3;        // rows r0 + 16*i
      const uint32_t sw = uint32_t(c ^ (r0 & 7)) << 4;
      uint32_t base[8];
      int iy0[8], ix0[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int m = m0 + r0 + 16 * i;
        bool mv = m < P.M;
        int mm = mv ? m : 0;
        int ox = mm % P.Wo;
        int t = mm / P.Wo;
        int oy = t % P.Ho;
        int b = t / P.Ho;
        base[i] = uint32_t(b) * uint32_t(P.Hi * P.Wi);
        iy0[i] = mv ? oy * P.stride - P.pad : -100000;   // invalid rows fall out of bounds -> zero fill
        ix0[i] = ox * P.stride - P.pad;
      }
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t sa = smem_base + s * L::kStageBytes;
        const int k0 = (kb_begin + kb) * BK + c * 8;
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
        cp_async_arrive_on(full_bar(s));     // asynchronous arrival when this thread's chunks have landed: no wait_group
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }

    // ------------------------------------------------------------------ epilogue
    const int row = tid;                   // TMEM lane == tile row
    int m;
    bool mvalid;
    if (a_mode == A_TMA4D) {
      const int ry = row / P.tw, rx = row - ry * P.tw;
      m = (tb * P.Ho + oy0 + ry) * P.Wo + ox0 + rx;
      mvalid = ry < P.th && oy0 + ry < P.Ho;      // tw divides Wo; the last tile row of an image may hang over
    } else {
      m = m0 + row;
      mvalid = m < P.M;
    }
    const uint32_t trow = tmem_d + (uint32_t(warp * 32) << 16);
    const float rbias = ((P.epi & ICAF_EPI_BIAS_ROW) && pr.bias && mvalid) ? __ldg(pr.bias + m) : 0.f;
    __half* yrow = pr.y + size_t(mvalid ? m : 0) * pr.y_ld;
    const __half* rrow = pr.res ? pr.res + size_t(mvalid ? m : 0) * pr.res_ld : nullptr;
    const int mode = (P.epi & ICAF_EPI_SCALED_RES) ? 2 : (rrow ? 1 : 0);
    if (rrow && mvalid) {
      for (int cb = 0; cb < BN && n0 + cb < P.N; cb += 64) prefetch_l2(rrow + n0 + cb);   // 128-byte lines of the residual row
    }
#pragma unroll
    for (int i = 0; i < (BN + 127) / 128; ++i)
      if (tid + 128 * i < BN) sbias[tid + 128 * i] = bias_r[i];
    named_bar_sync(1, 128);                // bias tile visible to the four epilogue warps
    EpiRow ex;
    ex.sum = ex.sumsq = 0.f; ex.ln_a = 1.f; ex.ln_mu = 0.f; ex.ln_s = nullptr;
    if (XM) {
      ex.ln_s = pr.ln_s ? pr.ln_s + n0 : nullptr;
      if (P.ln_parts > 0) epi_row_ln(ex, P, pr, m, mvalid);   // row statistics: fetched while the main loop still runs
    }
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    // XM: 9 / 10 = LayerNorm folded into this GEMM (no activation / GELU); 11 = scaled residual + statistics of the output rows
    const int mode_act = XM ? (P.ln_parts > 0 ? (P.act == ICAF_ACT_GELU ? 10 : 9) : 11) : P.act * 3 + mode;
    if (splits > 1) {
      // ---- split-K reduction through distributed shared memory ----
      // Every CTA's ring is idle once its accumulator is complete.  Barrier A: all accumulators done (so the leader's
      // ring may be overwritten); CTAs 1..S-1 then push their fp32 partial tile into the leader's ring; barrier B:
      // the leader adds them to its own accumulator and runs the real epilogue.
      constexpr int kPitch = BN + 4;                 // floats per staged row (+4: spreads the rows over the banks)
      cluster_arrive();
      cluster_wait();
      if (crank != 0) {
        const uint32_t dst0 = map_to_cta(smem_base, 0) + uint32_t(((crank - 1) * BM + row) * kPitch) * 4u;
#pragma unroll 1
        for (int cb = 0; cb < BN; cb += 32) {
          uint32_t acc[32];
          __syncwarp();
          tmem_ld32(trow + cb, acc);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            st_cluster_v4(dst0 + uint32_t(cb + 4 * q) * 4u, acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        }
      }
      cluster_arrive();
      cluster_wait();
    }
    if (crank == 0) {
      const float* part = reinterpret_cast<const float*>(smem_gen) + size_t(row) * (BN + 4);
#pragma unroll 1
      for (int cb = 0; cb < BN; cb += 32) {
        uint32_t acc[32];
        __syncwarp();
        tmem_ld32(trow + cb, acc);      // .sync.aligned: executed by the whole (converged) warp
        tmem_ld_wait();
        for (int r = 1; r < splits; ++r) {
          const float4* pp = reinterpret_cast<const float4*>(part + size_t(r - 1) * BM * (BN + 4) + cb);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 v = pp[q];
            acc[4 * q] = __float_as_uint(__uint_as_float(acc[4 * q]) + v.x);
            acc[4 * q + 1] = __float_as_uint(__uint_as_float(acc[4 * q + 1]) + v.y);
            acc[4 * q + 2] = __float_as_uint(__uint_as_float(acc[4 * q + 2]) + v.z);
            acc[4 * q + 3] = __float_as_uint(__uint_as_float(acc[4 * q + 3]) + v.w);
          }
        }
        const int nb = n0 + cb;
        if (mvalid && nb < P.N) {
          const int ncols = min(32, P.N - nb);
          const bool vec = ncols == 32 && ((reinterpret_cast<uintptr_t>(yrow + nb) & 15) == 0) &&
                           (!rrow || (reinterpret_cast<uintptr_t>(rrow + nb) & 15) == 0);
          const float* sb = sbias + cb;
          const __half* rp = rrow ? rrow + nb : nullptr;
          __half* yp = yrow + nb;
          // act / residual mode are warp-uniform: dispatch once per chunk to straight-line specialisations
          if (XM) {
            switch (mode_act) {
              case 9: epi_chunk<0, 0, 1>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols, ex, cb); break;
              case 10: epi_chunk<2, 0, 1>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols, ex, cb); break;
              default: epi_chunk<0, 2, 2>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols, ex, cb); break;
            }
          } else {
            switch (mode_act) {
              case 0: epi_chunk<0, 0>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols, ex, cb); break;
              case 1: epi_chunk<0, 1>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols, ex, cb); break;
              case 2: epi_chunk<0, 2>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols, ex, cb); break;
              case 3: epi_chunk<1, 0>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols, ex, cb); break;
              case 4: epi_chunk<1, 1>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols, ex, cb); break;
              case 5: epi_chunk<1, 2>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols, ex, cb); break;
              case 6: epi_chunk<2, 0>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols, ex, cb); break;
              case 7: epi_chunk<2, 1>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols, ex, cb); break;
              default: epi_chunk<2, 2>(acc, sb, rbias, alpha, beta, rp, yp, vec, ncols, ex, cb); break;
            }
          }
        }
      }
      if (XM && mode_act == 11 && mvalid && n0 < P.N) epi_row_emit(ex, P, pr, m, n0, min(n0 + BN, P.N));
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
    int s = 0;
    uint32_t ph = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_base + s * L::kStageBytes;
        const uint64_t ad = umma_desc_sw128(sa);
        const uint64_t bd = umma_desc_sw128(sa + L::kABytes);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_f16_ss(tmem_d, ad + uint64_t(2 * k), bd + uint64_t(2 * k), idesc, (kb | k) != 0);
        umma_commit(empty_bar(s));
        if (kb == nkb - 1) umma_commit(accum_bar);
      }
      __syncwarp();
      if (++s == kStages) { s = 0; ph ^= 1; }
    }
    if (splits > 1) { cluster_arrive(); cluster_wait(); cluster_arrive(); cluster_wait(); }
  } else {
    // ------------------------------------------------------------------ TMA producer (warp 5, one thread)
    if (elect_one()) {
      const CUtensorMap* mw = bz ? &maps.w[1] : &maps.w[0];
      const CUtensorMap* ma = bz ? &maps.a[1] : &maps.a[0];
      const uint32_t a_bytes = a_mode == A_TMA2D ? L::kABytes : (a_mode == A_TMA4D ? uint32_t(P.tw * P.th) * 128u : 0u);
      const uint32_t bytes = L::kBBytes + a_bytes;
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t sa = smem_base + s * L::kStageBytes;
        mbar_arrive_expect_tx(full_bar(s), bytes);
        tma_load_2d(sa + L::kABytes, mw, full_bar(s), (kb_begin + kb) * BK, n0);
        if (a_mode == A_TMA2D) {
          tma_load_2d(sa, ma, full_bar(s), (kb_begin + kb) * BK, m0);
        } else if (a_mode == A_TMA4D) {
          const int k0 = (kb_begin + kb) * BK;
          const int tap = k0 / P.Cin;
          const int ch = k0 - tap * P.Cin;
          const int ky = tap / P.kw, kx = tap - ky * P.kw;
          tma_load_4d(sa, ma, full_bar(s), ch, ox0 * P.stride - P.pad + kx, oy0 * P.stride - P.pad + ky, tb);
        }
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
    if (splits > 1) { cluster_arrive(); cluster_wait(); cluster_arrive(); cluster_wait(); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<L::kTmemCols>(tmem_d);
  }
}

// ---------------------------------------------------------------------------------------------------
// CUDA-core reference with the identical contract (tests only).
struct SimtParams { ConvParams P; const __half* w[2]; };
__global__ void conv_gemm_simt_kernel(const SimtParams S) {
  pdl_launch_dependents();
  pdl_wait();
  const ConvParams& P = S.P;
  const ConvProblem pr = pick_problem(P, blockIdx.z);
  const __half* w = blockIdx.z ? S.w[1] : S.w[0];
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)P.M * P.N) return;
  int n = int(idx % P.N);
  int m = int(idx / P.N);
  int ox = m % P.Wo, t = m / P.Wo, oy = t % P.Ho, b = t / P.Ho;
  float acc = 0.f;
  for (int ky = 0; ky < P.kh; ++ky)
    for (int kx = 0; kx < P.kw; ++kx) {
      int iy = oy * P.stride - P.pad + ky, ix = ox * P.stride - P.pad + kx;
      if ((unsigned)iy >= (unsigned)P.Hi || (unsigned)ix >= (unsigned)P.Wi) continue;
      const __half* xp = pr.x + (size_t(b) * P.Hi * P.Wi + size_t(iy) * P.Wi + ix) * pr.x_ld;
      const __half* wp = w + size_t(n) * P.k_pad + (ky * P.kw + kx) * P.Cin;
      for (int c = 0; c < P.Cin; ++c) acc += __half2float(xp[c]) * __half2float(wp[c]);
    }
  if (pr.bias) acc += (P.epi & ICAF_EPI_BIAS_ROW) ? pr.bias[m] : pr.bias[n];
  if (P.act == ICAF_ACT_SILU) acc = acc / (1.0f + expf(-acc));
  else if (P.act == ICAF_ACT_GELU) acc = gelu_erf_f(acc);
  if (pr.res) {
    float rf = __half2float(pr.res[size_t(m) * pr.res_ld + n]);
    acc = (P.epi & ICAF_EPI_SCALED_RES) ? (*pr.alpha) * rf + (*pr.beta) * acc : acc + rf;
  }
  pr.y[size_t(m) * pr.y_ld + n] = __float2half_rn(acc);
}

#ifdef ICAF_PROBE
static int g_dbg = 0, g_dbg_bn = 0;       // probe builds only: see icaf_debug_set below
#else
constexpr int g_dbg = 0, g_dbg_bn = 0;
#endif
// Geometry half of the argument check (host only): everything the planner needs, no pointers.
static int fill_geom(const icaf_conv_geom* g, int n_io, ConvParams& P) {
  if (!g || n_io < 1 || n_io > 2) return set_error(ICAF_ERR_BAD_ARG, "conv2d: need 1 or 2 problems");
  if (!(g->Cin == 4 || g->Cin % 8 == 0)) return set_error(ICAF_ERR_UNSUPPORTED, "conv2d: Cin must be 4 or a multiple of 8");
  if (g->Cin == 4 && (g->Wi % 2 || g->stride % 2 || g->pad % 2 || g->kw % 2))
    return set_error(ICAF_ERR_UNSUPPORTED, "conv2d: packed-image (Cin=4) path needs even Wi, stride, pad, kw");
  if (g->k_pad % 64 || g->k_pad < g->kh * g->kw * g->Cin) return set_error(ICAF_ERR_BAD_ARG, "conv2d: bad k_pad");
  if (g->w_rows < g->Cout) return set_error(ICAF_ERR_BAD_ARG, "conv2d: w_rows < Cout");
  long long M = (long long)g->B * g->Ho * g->Wo;
  if (M <= 0 || M > 0x7fffffffLL || (long long)g->B * g->Hi * g->Wi > 0x7fffffffLL)
    return set_error(ICAF_ERR_BAD_ARG, "conv2d: size out of range");
  if ((g->epi & (ICAF_EPI_ADD_RES | ICAF_EPI_SCALED_RES)) == (ICAF_EPI_ADD_RES | ICAF_EPI_SCALED_RES))
    return set_error(ICAF_ERR_BAD_ARG, "conv2d: ADD_RES and SCALED_RES are exclusive");
  if ((g->epi & ICAF_EPI_LN_FOLD) && ((g->epi & ~ICAF_EPI_LN_FOLD) != 0 || !(g->act == ICAF_ACT_NONE || g->act == ICAF_ACT_GELU) ||
                                      g->kh != 1 || g->kw != 1 || g->stride != 1 || g->pad != 0))
    return set_error(ICAF_ERR_UNSUPPORTED, "conv2d: LN_FOLD is a linear-layer (1x1) epilogue without residual, activation none / GELU");
  if ((g->epi & ICAF_EPI_EMIT_STATS) && ((g->epi & ~ICAF_EPI_EMIT_STATS) != ICAF_EPI_SCALED_RES || g->act != ICAF_ACT_NONE))
    return set_error(ICAF_ERR_UNSUPPORTED, "conv2d: EMIT_STATS rides on the SCALED_RES epilogue without activation");
  P.M = int(M); P.N = g->Cout; P.K = g->kh * g->kw * g->Cin; P.k_pad = g->k_pad;
  P.B = g->B; P.Hi = g->Hi; P.Wi = g->Wi; P.Cin = g->Cin; P.Ho = g->Ho; P.Wo = g->Wo;
  P.kh = g->kh; P.kw = g->kw; P.stride = g->stride; P.pad = g->pad; P.act = g->act; P.epi = g->epi;
  P.a_mode = A_GATHER; P.tw = P.th = P.tiles_x = P.tiles_y = 0; P.stages = 2; P.splits = 1; P.cblk = 64; P.halo = 0; P.dbg = g_dbg;
  P.ln_parts = 0; P.ln_eps = 0.f; P.ln_inv_k = 0.f;
  P.dense16 = 1;                          // plan-only calls have no pointers: assume a dense input frame
  memset(P.p, 0, sizeof(P.p));
  return ICAF_OK;
}

static int fill_params(const icaf_conv_geom* g, const icaf_conv_io* io, int n_io, ConvParams& P, const __half* (&w)[2]) {
  if (!io) return set_error(ICAF_ERR_BAD_ARG, "conv2d: null io");
  if (int rc = fill_geom(g, n_io, P)) return rc;
  for (int i = 0; i < 2; ++i) {
    const icaf_conv_io& s = io[i < n_io ? i : 0];
    bool need_res = g->epi & (ICAF_EPI_ADD_RES | ICAF_EPI_SCALED_RES);
    if (!s.x || !s.w || !s.y || (need_res && !s.res) || ((g->epi & ICAF_EPI_SCALED_RES) && (!s.alpha || !s.beta)))
      return set_error(ICAF_ERR_BAD_ARG, "conv2d: null pointer");
    if ((reinterpret_cast<uintptr_t>(s.x) & 15) || (reinterpret_cast<uintptr_t>(s.w) & 15) || (s.x_ld % 8 && g->Cin != 4) ||
        (g->Cin == 4 && s.x_ld != 4))
      return set_error(ICAF_ERR_BAD_ARG, "conv2d: x / w must be 16-byte aligned with x_ld a multiple of 8");
    if ((g->epi & ICAF_EPI_LN_FOLD) && (!s.ln_stats || !s.ln_colsum || s.ln_parts < 1 || s.ln_parts > 64))
      return set_error(ICAF_ERR_BAD_ARG, "conv2d: LN_FOLD needs ln_stats, ln_colsum and 1..64 partials per row");
    if ((g->epi & ICAF_EPI_EMIT_STATS) && !s.stats_out) return set_error(ICAF_ERR_BAD_ARG, "conv2d: EMIT_STATS needs stats_out");
    if (i > 0 && (g->epi & ICAF_EPI_LN_FOLD) && (s.ln_parts != io[0].ln_parts || s.ln_eps != io[0].ln_eps))
      return set_error(ICAF_ERR_BAD_ARG, "conv2d: grouped problems must share ln_parts / ln_eps");
    P.p[i] = ConvProblem{(const __half*)s.x, s.bias, need_res ? (const __half*)s.res : nullptr, (__half*)s.y, s.alpha, s.beta,
                         s.x_ld, s.res_ld, s.y_ld,
                         (g->epi & ICAF_EPI_LN_FOLD) ? (const float2*)s.ln_stats : nullptr,
                         (g->epi & ICAF_EPI_LN_FOLD) ? s.ln_colsum : nullptr,
                         (g->epi & ICAF_EPI_EMIT_STATS) ? (float2*)s.stats_out : nullptr};
    if (i == 0 && (g->epi & ICAF_EPI_LN_FOLD)) { P.ln_parts = s.ln_parts; P.ln_eps = s.ln_eps; P.ln_inv_k = 1.0f / float(P.K); }
    if (i < n_io && s.x_ld != 16) P.dense16 = 0;
    w[i] = (const __half*)s.w;
  }
  return ICAF_OK;
}

// Pick how the A tile is staged (see the header comment) and, for A_TMA4D, the tile shape.
static void plan_a_mode(const icaf_conv_geom* g, ConvParams& P) {
  if (g->kh == 1 && g->kw == 1 && g->stride == 1 && g->pad == 0 && g->Cin % 8 == 0) {
    P.a_mode = A_TMA2D;
    return;
  }
  if ((g->Cin % 64 == 0 || g->Cin == 32 || g->Cin == 16) && g->stride <= 2) {
    // tile = th x tw output pixels of one image, tw | Wo, tw*th <= 128: maximise the fraction of useful MMA rows
    int best_tw = 0, best_th = 0;
    double best_u = 0.0;
    for (int tw = 1; tw <= 128 && tw <= g->Wo; ++tw) {
      if (g->Wo % tw || tw * g->stride > 256) continue;
      int th = 128 / tw;
      if (th > g->Ho) th = g->Ho;
      if (th * g->stride > 256) th = 256 / g->stride;
      int ty = (g->Ho + th - 1) / th;
      double u = double(g->Wo) * g->Ho / (double(g->Wo / tw) * ty * 128.0);
      if (u > best_u + 1e-9 || (u > best_u - 1e-9 && tw > best_tw)) { best_u = u; best_tw = tw; best_th = th; }
    }
    if (best_u >= 0.6) {
      P.a_mode = A_TMA4D; P.tw = best_tw; P.th = best_th; P.tiles_x = g->Wo / best_tw; P.tiles_y = (g->Ho + best_th - 1) / best_th;
      P.cblk = g->Cin < 64 ? g->Cin : 64;
    }
  }
}

template <int BN>
static int plan_tc(ConvParams& P, int n_io, ConvPlan& pl) {
  using L = SmemLayout<BN>;
  constexpr int kSmemCap = 227 * 1024;
  const int mt = P.a_mode == A_TMA4D ? P.B * P.tiles_x * P.tiles_y : (P.M + BM - 1) / BM;
  if (P.a_mode == A_TMA4D && !(P.tw >= 1 && P.th >= 1 && P.tw * P.th <= BM && P.tiles_x * P.tw >= P.Wo && P.tiles_y * P.th >= P.Ho &&
                               P.cblk == 64 && P.Cin % 64 == 0))
    return set_error(ICAF_ERR_BAD_ARG, "conv2d(tc): 4-D tiles must cover the map with at most 128 pixels each, 64-channel blocks");
  unsigned gx = unsigned(mt), gy = unsigned((P.N + BN - 1) / BN), gz = unsigned(n_io);
  // Ring depth: a grid that fits in one wave gets the whole SM (deep ring: the K loop is latency-bound at small M);
  // otherwise two CTAs share an SM so that one CTA's epilogue overlaps the other's main loop.
  const long long ctas = (long long)gx * gy * gz;
  const int budget = (ctas <= pl.sms || BN > 128) ? kSmemCap : (kSmemCap / 2 - 1024);
  int stages = (budget - L::kTailBytes) / L::kStageBytes;
  const int nkb = P.k_pad / BK;
  if (stages > nkb) stages = nkb;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  // Split-K: a grid that leaves most SMs idle on a deep K loop is spread over clusters of `splits` CTAs per tile.
  int splits = 1;
  if (ctas * 2 <= pl.sms && nkb >= 8) {
    splits = int(pl.sms / ctas);
    if (splits > 8) splits = 8;                    // portable cluster size
    if (splits > nkb / 4) splits = nkb / 4;        // >= 4 K blocks per CTA
    const int per_split = BM * (BN + 4) * 4;       // staged fp32 partial tile in the leader's ring
    while (splits > 1 && (splits - 1) * per_split > stages * L::kStageBytes) --splits;
    if (splits < 1) splits = 1;
  }
  if (splits > 1) {
    const int per = (nkb + splits - 1) / splits;
    if (stages > per) stages = per < 2 ? 2 : per;
    while ((splits - 1) * (BM * (BN + 4) * 4) > stages * L::kStageBytes) ++stages;   // keep room for the partial tiles
    gx *= splits;
  }
  if (stages > kMaxStages || L::total(stages) > kSmemCap)
    return set_error(ICAF_ERR_BAD_ARG, "conv2d(tc): shared-memory plan exceeds the ring / 227 KB");
  P.stages = stages;
  P.splits = splits;
  pl.kernel = ICAF_KERNEL_TC; pl.bn = BN;
  pl.grid_x = gx; pl.grid_y = gy; pl.grid_z = gz; pl.cluster = unsigned(splits);
  pl.smem = L::total(stages);
  pl.total = int(ctas); pl.m_tiles = mt; pl.m_pairs = 0; pl.n_tiles = int(gy);
  return ICAF_OK;
}

template <int BN, bool XM>
static int launch_tc_x(const ConvParams& P, const ConvPlan& pl, const __half* const (&w)[2], const icaf_conv_geom* g, int n_io, cudaStream_t st) {
  static bool configured[kMaxDevices] = {false};
  if (int rc = configure_smem(conv_gemm_tc_kernel<BN, XM>, 227 * 1024, configured, "conv2d: cudaFuncSetAttribute")) return rc;
  ConvMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int i = 0; i < n_io; ++i) {
    int rc = encode_tmap_2d(&maps.w[i], w[i], (uint64_t)P.k_pad, (uint64_t)g->w_rows, (uint64_t)P.k_pad * 2, BK, BN);
    if (rc) return rc;
    const ConvProblem& pr = P.p[i];
    if (P.a_mode == A_TMA2D)
      rc = encode_tmap_2d(&maps.a[i], pr.x, (uint64_t)P.Cin, (uint64_t)P.M, (uint64_t)pr.x_ld * 2, BK, BM);
    else if (P.a_mode == A_TMA4D)
      rc = encode_tmap_nhwc(&maps.a[i], pr.x, P.Cin, P.Wi, P.Hi, P.B, pr.x_ld, BK, P.tw * P.stride, P.th * P.stride, P.stride,
                            P.stride);
    if (rc) return rc;
  }
  if (n_io == 1) { maps.w[1] = maps.w[0]; maps.a[1] = maps.a[0]; }
  launch_kc(conv_gemm_tc_kernel<BN, XM>, dim3(pl.grid_x, pl.grid_y, pl.grid_z), dim3(kThreads), (size_t)pl.smem, st, pl.cluster, P, maps);
  return check_launch("conv2d_fwd");
}
template <int BN>
static int launch_tc(const ConvParams& P, const ConvPlan& pl, const __half* const (&w)[2], const icaf_conv_geom* g, int n_io, cudaStream_t st) {
  return (P.epi & (ICAF_EPI_LN_FOLD | ICAF_EPI_EMIT_STATS)) ? launch_tc_x<BN, true>(P, pl, w, g, n_io, st) : launch_tc_x<BN, false>(P, pl, w, g, n_io, st);
}

// ---------------------------------------------------------------------------------------------------
// The dispatcher, host only: staging mode, tile shapes, kernel family and tile width for one layer geometry.  No CUDA call.
// pair_mode: -1 = the ICAF_PAIR environment switch (default 1), 0 = never CTA pairs, 1 = heuristic, 2 = pairs wherever they can run.
static int env_pair_mode() {
  static const int v = []() { const char* e = getenv("ICAF_PAIR"); return !e ? 1 : (e[0] == '0' ? 0 : (e[0] == 'a' ? 2 : 1)); }();
  return v;
}

static int plan_conv(const icaf_conv_geom* g, int n_io, int sms, int pair_mode, ConvParams& P, ConvPlan& pl) {
  memset(&pl, 0, sizeof(pl));
  pl.sms = sms;
  if (sms < 1) return set_error(ICAF_ERR_BAD_ARG, "conv2d: SM count must be positive");
  plan_a_mode(g, P);
  // Tile width: the widest BN that still yields at least ~one CTA per SM (two waves for the 1-CTA/SM BN=256); small
  // problems take BN=32 so that more SMs share the K loop.
  const long long mt = P.a_mode == A_TMA4D ? (long long)P.B * P.tiles_x * P.tiles_y : (P.M + BM - 1) / BM;
  auto ctas = [&](int bn) { return mt * ((P.N + bn - 1) / bn) * n_io; };
  // CTA pairs (conv_pair.cu): two SMs share one 256 x BN tile and each loads only half of the filter tile.  Measured
  // (yolov5l batch 16, profiles/): always a win at BN = 256 once a wave of clusters is full (or half full with a deep K
  // loop); at BN = 128 / 64 only for the deep-K (3x3) layers -- the short-K 1x1 layers are HBM / epilogue bound and lose
  // to the pair's extra synchronisation.  ICAF_PAIR=0 disables it, ICAF_PAIR=all forces it wherever it can run (tests).
  const int pair_env = pair_mode < 0 ? env_pair_mode() : pair_mode;
  const bool pair_ok = pair_env && (P.a_mode == A_TMA2D || (P.a_mode == A_TMA4D && P.cblk == 64));
  const int nkb_all = P.k_pad / BK;
  auto pair_wanted = [&](int b) {
    if (!pair_ok || b < 64) return false;
    if (pair_env == 2) return true;
    const long long pairs = ctas(b) / 2;
    if (b == 256) return pairs >= sms / 2 || (pairs >= sms / 4 && nkb_all >= 32);
    return pairs >= sms / 2 && nkb_all >= 8;
  };
  // 3x3 / stride 1 layers the pair kernel can run with halo copies (conv_pair.cu): once those have removed most of the
  // activation traffic, BN = 128 pairs beat BN = 256 (probe: 50 vs 52-54 us on M20480 N256 K2304, 54 vs 63 us on the P5 layer:
  // finer wave balance, four accumulator buffers)
  static const bool halo_on = []() { const char* e = getenv("ICAF_HALO"); return !(e && e[0] == '0'); }();
  static const bool bres_on = []() { const char* e = getenv("ICAF_HALO"); return !(e && e[0] == '1'); }();   // ICAF_HALO=1: copies, no resident filter
  const bool halo64 = pair_ok && halo_on && P.a_mode == A_TMA4D && P.cblk == 64 && g->kh == 3 && g->kw == 3 && g->stride == 1 &&
                      g->pad == 1 && g->Cin % 64 == 0 &&
                      double(g->Wo) * g->Ho >= 0.6 * (double((g->Wo + 7) / 8) * ((g->Ho + 15) / 16) * 128.0);
  int bn = 32;
  if (halo64 && P.N >= 128 && pair_wanted(128)) bn = 128;
  else if (P.N >= 256 && (ctas(256) >= 2 * sms || pair_wanted(256))) bn = 256;
  else if (P.N > 64 && ctas(128) >= sms) bn = 128;
  else if (P.N > 32 && ctas(64) >= sms) bn = 64;
  if (g_dbg_bn) bn = g_dbg_bn;
  // Many tiles per SM: the persistent kernel overlaps main loop and epilogue across tiles (conv_persist.cu).
  static const bool persist_on = []() { const char* e = getenv("ICAF_PERSISTENT"); return !(e && e[0] == '0'); }();
  const bool persistent = persist_on && ctas(bn) >= 2 * sms && P.a_mode != A_GATHER;   // both operands by TMA
  if (P.a_mode == A_TMA4D && P.cblk < 64 && !persistent) {      // small-Cin TMA staging exists in the persistent kernel only
    P.a_mode = A_GATHER; P.tw = P.th = P.tiles_x = P.tiles_y = 0; P.cblk = 64;
  }
  // (A wave-tail scheme -- peel total % SMs tiles off into a split-K cluster launch -- was measured and dropped: these
  // layers are bound by chip-wide L2->SM bandwidth, so a partly filled last wave just streams the same bytes through
  // fewer, faster CTAs; the second launch only added its fixed cost: 65 -> 87 us on the 320-tile P4 3x3 layer.)
  // The image stem (3x3 / s1 over the 16-channel space-to-depth frame) has its own kernel once a wave of tiles exists
  // (conv_stem.cu: x-merged rows, four accumulators per tile, resident filter).  ICAF_STEM=0 keeps the generic kernels.
  static const bool stem_on = []() { const char* e = getenv("ICAF_STEM"); return !(e && e[0] == '0'); }();
  if (stem_on && P.dense16 && stem_eligible(g) &&
      (long long)g->B * ((g->Wo / 4 + 7) / 8) * ((g->Ho + 15) / 16) * n_io >= sms / 2)
    return plan_stem(P, g, n_io, pl);
  if (pair_env && halo_on && P.a_mode == A_TMA4D && P.cblk < 64 && g->kh == 3 && g->kw == 3 && g->stride == 1 && g->pad == 1) {
    // 16- / 32-channel 3x3 layers (the image stem over the space-to-depth frame): CTA pairs + halo copies, 64-wide tiles
    const int tx = (g->Wo + 7) / 8, ty = (g->Ho + 15) / 16;
    const long long pairs = (long long)g->B * tx * ty * ((P.N + 63) / 64) * n_io / 2;
    if (double(g->Wo) * g->Ho >= 0.6 * (double(tx) * ty * 128.0) && (pair_env == 2 || pairs >= sms / 2)) {
      P.halo = (bres_on && P.N <= 64) ? 2 : 1;           // 2: the filter (one 64-wide N tile, one channel block) stays resident
      P.tw = 8;
      P.th = 16;
      P.tiles_x = tx;
      P.tiles_y = ty;
      return plan_pair<64>(P, g, n_io, pl);
    }
  }
  if (pair_wanted(bn) && P.a_mode != A_GATHER) {
    // 3x3 / stride 1 layers on 16 x 8 pixel tiles: every activation row is fetched three times instead of nine (conv_pair.cu)
    if (halo64) {   // tiles may hang over the right / bottom edge (P5: 16 x 20)
      // (resident filter, halo mode 2, measured slower here: 233 vs 209 us at Cin = 64, N = 64)
      P.halo = 1;
      P.tw = 8;
      P.th = 16;
      P.tiles_x = (g->Wo + 7) / 8;
      P.tiles_y = (g->Ho + 15) / 16;
    }
    switch (bn) {
      case 256: return plan_pair<256>(P, g, n_io, pl);
      case 128: return plan_pair<128>(P, g, n_io, pl);
      default: return plan_pair<64>(P, g, n_io, pl);
    }
  }
  if (persistent) {
    switch (bn) {
      case 256: return plan_persist<256>(P, n_io, pl);
      case 128: return plan_persist<128>(P, n_io, pl);
      case 64: return plan_persist<64>(P, n_io, pl);
      default: return plan_persist<32>(P, n_io, pl);
    }
  }
  switch (bn) {
    case 256: return plan_tc<256>(P, n_io, pl);
    case 128: return plan_tc<128>(P, n_io, pl);
    case 64: return plan_tc<64>(P, n_io, pl);
    default: return plan_tc<32>(P, n_io, pl);
  }
}

}  // namespace icaf

using namespace icaf;

#ifdef ICAF_PROBE
// Probe builds only (not part of the ABI, absent from the shipped library): kernel-stage switches + forced tile width.
extern "C" void icaf_debug_set(int dbg, int bn) { g_dbg = dbg; g_dbg_bn = bn; }
#endif

extern "C" int icaf_conv2d_plan(const icaf_conv_geom* g, int n_io, int sm_count, int pair_mode, icaf_conv_plan* out) {
  if (!out) return set_error(ICAF_ERR_BAD_ARG, "conv2d_plan: null output");
  ConvParams P;
  int rc = fill_geom(g, n_io, P);
  if (rc) return rc;
  ConvPlan pl;
  rc = plan_conv(g, n_io, sm_count, pair_mode, P, pl);
  if (rc) return rc;
  out->kernel = pl.kernel; out->bn = pl.bn; out->a_mode = P.a_mode;
  out->tile_w = P.tw; out->tile_h = P.th; out->tiles_x = P.tiles_x; out->tiles_y = P.tiles_y;
  out->cblk = P.cblk; out->halo = P.halo; out->stages = P.stages; out->splits = P.splits;
  out->grid_x = int(pl.grid_x); out->grid_y = int(pl.grid_y); out->grid_z = int(pl.grid_z); out->cluster = int(pl.cluster);
  out->smem_bytes = pl.smem; out->work_items = pl.total;
  return ICAF_OK;
}

extern "C" int icaf_conv2d_fwd(const icaf_conv_geom* g, const icaf_conv_io* io, int n_io, void* stream) {
  ConvParams P;
  const __half* w[2];
  int rc = fill_params(g, io, n_io, P, w);
  if (rc) return rc;
  ConvPlan pl;
  rc = plan_conv(g, n_io, sm_count_cached(), -1, P, pl);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int key = pl.kernel * 1000 + pl.bn;
  switch (key) {
    case ICAF_KERNEL_PAIR * 1000 + 256: return launch_pair<256>(P, pl, w, g, n_io, st);
    case ICAF_KERNEL_PAIR * 1000 + 128: return launch_pair<128>(P, pl, w, g, n_io, st);
    case ICAF_KERNEL_PAIR * 1000 + 64: return launch_pair<64>(P, pl, w, g, n_io, st);
    case ICAF_KERNEL_PERSIST * 1000 + 256: return launch_persist<256>(P, pl, w, g, n_io, st);
    case ICAF_KERNEL_PERSIST * 1000 + 128: return launch_persist<128>(P, pl, w, g, n_io, st);
    case ICAF_KERNEL_PERSIST * 1000 + 64: return launch_persist<64>(P, pl, w, g, n_io, st);
    case ICAF_KERNEL_PERSIST * 1000 + 32: return launch_persist<32>(P, pl, w, g, n_io, st);
    case ICAF_KERNEL_TC * 1000 + 256: return launch_tc<256>(P, pl, w, g, n_io, st);
    case ICAF_KERNEL_TC * 1000 + 128: return launch_tc<128>(P, pl, w, g, n_io, st);
    case ICAF_KERNEL_TC * 1000 + 64: return launch_tc<64>(P, pl, w, g, n_io, st);
    case ICAF_KERNEL_TC * 1000 + 32: return launch_tc<32>(P, pl, w, g, n_io, st);
    case ICAF_KERNEL_STEM * 1000 + 64:
    case ICAF_KERNEL_STEM * 1000 + 32: return launch_stem(P, pl, w, g, n_io, st);
    default: return set_error(ICAF_ERR_BAD_ARG, "conv2d: the planner produced an unknown kernel / tile width");
  }
}

extern "C" int icaf_conv2d_fwd_simt(const icaf_conv_geom* g, const icaf_conv_io* io, int n_io, void* stream) {
  SimtParams S;
  int rc = fill_params(g, io, n_io, S.P, S.w);
  if (rc) return rc;
  long long total = (long long)S.P.M * S.P.N;
  dim3 grid((unsigned)((total + 255) / 256), 1, n_io);
  launch_k(conv_gemm_simt_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, S);
  return check_launch("conv2d_fwd_simt");
}
