// The image stem as its own kernel: 3x3 / stride 1 / pad 1 convolution over the 16-channel space-to-depth frame
// (= the reference's Conv(3, c, 6, 2, 2), models/common.py:48-60, YAML rows 0 and 10; see ops.pack_stem_weight).
//
// What the generic kernels cannot do for this layer: its K is 9 taps x 16 channels, so a TMA box of "64 channels" is three
// quarters zero fill (round 1: 1.9 GB of box traffic for 84 MB of input) and a 128-pixel tile is retired after nine tiny
// MMAs (2.2 us of hand-shakes per tile).  Here the frame is read through an x-MERGED view: four neighbouring pixels x 16
// channels = one dense 128-byte row ("super-pixel").  A tile is 16 rows x 8 super-pixels = 512 output pixels:
//   * three x-shifted copies of the 18-row patch (TMA box 64 x 8 x 18 of the (64, W/4, H, B) view) per tile -- every byte real;
//   * output pixel 4X + r of super-pixel X needs input pixels 4X + r + dx, dx in {-1, 0, 1}: sub-pixel j = r + dx of super-pixel
//     X (j = -1 -> sub-pixel 3 of X-1 = copy 0, j = 4 -> sub-pixel 0 of X+1 = copy 2).  Sub-pixel j of a row is the 32-byte chunk
//     j of the 128-byte row -- exactly a K step of the SW128 K-major operand.  So for each r there is ONE accumulator
//     (128 super-pixels x N channels) fed by nine M=128, K=16 MMAs whose A descriptors differ only in copy, row offset ky and
//     chunk: 36 MMAs per 512 pixels, no wasted FLOP, four accumulators = 4 N TMEM columns, double buffered;
//   * the filter (9 taps x N x 16, both streams) stays resident in shared memory;
//   * an epilogue thread owns one super-pixel: its four output pixels are 4 N contiguous halfs in NHWC memory.
// 576 threads: four epilogue groups of four warps (group = accumulator buffer x half of the four sub-pixel accumulators; TMEM
// loads double-buffered against the math), warp 16 TMEM + MMA issue, warp 17 TMA.
#include <cstring>

#include "conv_common.cuh"

namespace icaf {

constexpr int kSThreads = 576;
constexpr int kSEpiWarps = 16;
constexpr int kSCopyBytes = 18 * 8 * 128;          // one x-shifted copy: 18 rows x 8 super-pixels x 128 B
constexpr int kSStageBytes = 3 * kSCopyBytes;      // 55296
constexpr int kSStages = 3;
constexpr int kSFilterBytes = 3 * 64 * 128;        // per stream: K blocks 0-63 / 64-127 / 128-191 of [64 rows][64 K], SW128
constexpr int kSBarOff = kSStages * kSStageBytes + 2 * kSFilterBytes;
constexpr int kSSmem = kSBarOff + 256 + 2 * 64 * 4 + 1024;

struct StemParams {
  ConvProblem p[2];
  int B, H, W, N, tiles_x, tiles_y, total, n_io, act;
};
struct StemMaps { CUtensorMap w[2]; CUtensorMap a[2]; };

template <int BN>
__global__ void __launch_bounds__(kSThreads, 1) conv_stem_kernel(const StemParams P, const __grid_constant__ StemMaps maps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t filt = smem_base + kSStages * kSStageBytes;
  const uint32_t bar_base = smem_base + kSBarOff;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (8 + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (10 + b); };
  const uint32_t filt_bar = bar_base + 8u * 12;
  const uint32_t tmem_slot = bar_base + 8u * 13;
  float* sbias = reinterpret_cast<float*>(smem_gen + kSBarOff + 256);          // [2][64]

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < kSStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 8); }     // two groups x four warps drain a buffer
    mbar_init(filt_bar, 1);
    fence_mbar_init();
  }
  if (warp == kSEpiWarps) tmem_alloc<8 * BN>(tmem_slot);                       // 2 buffers x 4 accumulators x BN columns
  if (warp == kSEpiWarps + 1 && lane_id() == 0) {
    tma_prefetch_desc(&maps.w[0]); tma_prefetch_desc(&maps.a[0]);
    if (P.n_io > 1) { tma_prefetch_desc(&maps.w[1]); tma_prefetch_desc(&maps.a[1]); }
  }
  if (tid < 128) {                                                            // bias of both streams (parameters: before the PDL wait)
    const int z = tid >> 6, n = tid & 63;
    const float* pb = z ? P.p[1].bias : P.p[0].bias;
    sbias[tid] = (z < P.n_io && pb && n < P.N) ? __ldg(pb + n) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + kSBarOff + 8 * 13);
  const int per_img = P.tiles_x * P.tiles_y;
  const int per_prob = P.B * per_img;

  if (warp < kSEpiWarps) {
    // ------------------------------------------------------------------ epilogue groups
    const int eg = warp >> 2, gt = tid & 127;
    const int buf = eg >> 1, r_lo = (eg & 1) * 2;                              // this group's buffer and its two sub-pixel accumulators
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const int yl = gt >> 3, xl = gt & 7;                                      // super-pixel of this thread inside the tile
    constexpr int kChunks = 2 * (BN / 16);                                    // 16-column chunks of this group's two accumulators
    int it = 0;
    for (int t = blockIdx.x + buf * gridDim.x; t < P.total; t += 2 * gridDim.x, ++it) {
      const int z = t / per_prob;
      int r0 = t - z * per_prob;
      const int b = r0 / per_img;
      r0 -= b * per_img;
      const int ty = r0 / P.tiles_x, tx = r0 - ty * P.tiles_x;
      const ConvProblem pr = pick_problem_stem(P.p, z);
      const int y = ty * 16 + yl, x0 = (tx * 8 + xl) * 4 + r_lo;
      const bool valid = y < P.H && x0 < P.W;
      __half* yrow = pr.y + (size_t(size_t(b) * P.H + (valid ? y : 0)) * P.W + (valid ? x0 : 0)) * pr.y_ld;
      const int al = (reinterpret_cast<uintptr_t>(yrow) & 31) == 0 && (pr.y_ld * 2) % 32 == 0 ? 2 : ((reinterpret_cast<uintptr_t>(yrow) & 15) == 0 && (pr.y_ld * 2) % 16 == 0 ? 1 : 0);
      const float* sb = sbias + z * 64;
      mbar_wait(tfull_bar(buf), it & 1);
      tc_fence_after();
      EpiRow ex;
      ex.sum = ex.sumsq = 0.f; ex.ln_a = 1.f; ex.ln_mu = 0.f; ex.ln_s = nullptr;
      const uint32_t trow = tmem_base + uint32_t(buf * 4 * BN + r_lo * BN) + lane_off;   // 2 BN consecutive columns: accumulators r_lo, r_lo + 1
      auto chunk = [&](const uint32_t (&acc)[16], int ci) {
        const int r = ci / (BN / 16), cb = (ci % (BN / 16)) * 16;
        const int nc = P.N - cb;
        if (valid && x0 + r < P.W && nc > 0) {
          __half* yp = yrow + size_t(r) * pr.y_ld + cb;
          if (P.act == ICAF_ACT_SILU) epi_chunk16<1, 0>(acc, sb + cb, 0.f, 0.f, 1.f, nullptr, yp, nc >= 16 ? al : 0, nc, true, ex, cb);
          else epi_chunk16<0, 0>(acc, sb + cb, 0.f, 0.f, 1.f, nullptr, yp, nc >= 16 ? al : 0, nc, true, ex, cb);
        }
      };
      uint32_t acc0[16], acc1[16];
      __syncwarp();
      tmem_ld16(trow, acc0);
#pragma unroll 1
      for (int ci = 0; ci < kChunks; ci += 2) {                                // chunk ci+1 in flight during the math of chunk ci
        tmem_ld_wait();
        tmem_ld16(trow + (ci + 1) * 16, acc1);
        chunk(acc0, ci);
        tmem_ld_wait();
        if (ci + 2 < kChunks) {
          tmem_ld16(trow + (ci + 2) * 16, acc0);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane_id() == 0) mbar_arrive(tempty_bar(buf));                    // accumulators are in registers: hand the buffer back
        }
        chunk(acc1, ci + 1);
      }
    }
  } else if (warp == kSEpiWarps) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
    mbar_wait(filt_bar, 0);
    int s = 0, i = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < P.total; t += gridDim.x, ++i) {
      const int buf = i & 1;
      const int z = t / per_prob;
      mbar_wait(tempty_bar(buf), ((i >> 1) & 1) ^ 1);                          // the buffer's epilogue group has drained it
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_base + s * kSStageBytes;
        const uint32_t fb = filt + z * kSFilterBytes;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const uint32_t tmem_d = tmem_base + uint32_t(buf * 4 * BN + r * BN);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int j = r + kx - 1;                                        // sub-pixel of the input pixel
              const int copy = j < 0 ? 0 : (j > 3 ? 2 : 1);
              const int chunk = (j + 4) & 3;
              const int tap = ky * 3 + kx;
              const uint64_t ad = umma_desc_sw128(sa + copy * kSCopyBytes + ky * 1024) + uint64_t(2 * chunk);
              const uint64_t bd = umma_desc_sw128(fb + (tap >> 2) * (64 * 128)) + uint64_t(2 * (tap & 3));
              umma_f16_ss(tmem_d, ad, bd, idesc, (ky | kx) != 0);
            }
          }
        }
        umma_commit(empty_bar(s));
        umma_commit(tfull_bar(buf));
      }
      __syncwarp();
      if (++s == kSStages) { s = 0; ph ^= 1; }
    }
  } else if (lane_id() == 0) {
    // ------------------------------------------------------------------ TMA producer (one thread)
    mbar_arrive_expect_tx(filt_bar, uint32_t(P.n_io) * kSFilterBytes);
    for (int z = 0; z < P.n_io; ++z)
      for (int kb = 0; kb < 3; ++kb)
        tma_load_2d(filt + z * kSFilterBytes + kb * (64 * 128), z ? &maps.w[1] : &maps.w[0], filt_bar, kb * 64, 0);
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < P.total; t += gridDim.x) {
      const int z = t / per_prob;
      int r0 = t - z * per_prob;
      const int b = r0 / per_img;
      r0 -= b * per_img;
      const int ty = r0 / P.tiles_x, tx = r0 - ty * P.tiles_x;
      mbar_wait(empty_bar(s), ph ^ 1);
      const uint32_t sa = smem_base + s * kSStageBytes;
      mbar_arrive_expect_tx(full_bar(s), kSStageBytes);
      const CUtensorMap* ma = z ? &maps.a[1] : &maps.a[0];
#pragma unroll
      for (int c = 0; c < 3; ++c) tma_load_4d(sa + c * kSCopyBytes, ma, full_bar(s), 0, tx * 8 - 1 + c, ty * 16 - 1, b);
      if (++s == kSStages) { s = 0; ph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kSEpiWarps) {
    tc_fence_after();
    tmem_dealloc<8 * BN>(tmem_base);
  }
}

// Host side ---------------------------------------------------------------------------------------------------
bool stem_eligible(const icaf_conv_geom* g) {
  return g->kh == 3 && g->kw == 3 && g->stride == 1 && g->pad == 1 && g->Cin == 16 && g->Cout <= 64 && g->Cout % 16 == 0 &&
         g->w_rows >= g->Cout && g->k_pad == 192 && g->epi == 0 && (g->act == ICAF_ACT_SILU || g->act == ICAF_ACT_NONE) &&
         g->Wo % 4 == 0 && g->Hi == g->Ho && g->Wi == g->Wo;
}

int plan_stem(ConvParams& P, const icaf_conv_geom* g, int n_io, ConvPlan& pl) {
  if (!stem_eligible(g)) return set_error(ICAF_ERR_BAD_ARG, "conv2d(stem): geometry is not the space-to-depth stem");
  P.a_mode = A_TMA4D; P.cblk = 64; P.tw = 32; P.th = 16; P.halo = 0;           // 16 rows x 8 super-pixels (32 pixels) per tile
  P.tiles_x = (g->Wo / 4 + 7) / 8; P.tiles_y = (g->Ho + 15) / 16;
  P.stages = kSStages; P.splits = 1;
  const int total = g->B * P.tiles_x * P.tiles_y * n_io;
  const int waves = (total + pl.sms - 1) / pl.sms;
  pl.kernel = ICAF_KERNEL_STEM; pl.bn = g->Cout <= 32 ? 32 : 64;
  pl.grid_x = unsigned((total + waves - 1) / waves); pl.grid_y = pl.grid_z = 1; pl.cluster = 1;
  pl.smem = kSSmem;
  pl.total = total; pl.m_tiles = total / n_io; pl.m_pairs = 0; pl.n_tiles = 1;
  return ICAF_OK;
}

int launch_stem(const ConvParams& P, const ConvPlan& pl, const __half* const (&w)[2], const icaf_conv_geom* g, int n_io, cudaStream_t st) {
  StemParams S;
  memset(&S, 0, sizeof(S));
  S.p[0] = P.p[0]; S.p[1] = P.p[1];
  S.B = g->B; S.H = g->Ho; S.W = g->Wo; S.N = g->Cout; S.tiles_x = P.tiles_x; S.tiles_y = P.tiles_y; S.total = pl.total; S.n_io = n_io;
  S.act = g->act;
  StemMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int i = 0; i < n_io; ++i) {
    if (P.p[i].x_ld != 16) return set_error(ICAF_ERR_BAD_ARG, "conv2d(stem): the space-to-depth frame must be dense (pixel pitch 16)");
    int rc = encode_tmap_2d(&maps.w[i], w[i], (uint64_t)g->k_pad, (uint64_t)g->w_rows, (uint64_t)g->k_pad * 2, 64, 64);
    if (rc) return rc;
    rc = encode_tmap_nhwc(&maps.a[i], P.p[i].x, 64, g->Wi / 4, g->Hi, g->B, 64, 64, 8, 18, 1, 1);     // x-merged view: 4 pixels = 1 row
    if (rc) return rc;
  }
  if (n_io == 1) { maps.w[1] = maps.w[0]; maps.a[1] = maps.a[0]; }
  static bool configured[2][kMaxDevices] = {{false}, {false}};
  if (pl.bn == 32) {
    if (int rc = configure_smem(conv_stem_kernel<32>, kSSmem, configured[0], "conv2d: cudaFuncSetAttribute (stem)")) return rc;
    launch_k(conv_stem_kernel<32>, dim3(pl.grid_x), dim3(kSThreads), (size_t)kSSmem, st, S, maps);
  } else {
    if (int rc = configure_smem(conv_stem_kernel<64>, kSSmem, configured[1], "conv2d: cudaFuncSetAttribute (stem)")) return rc;
    launch_k(conv_stem_kernel<64>, dim3(pl.grid_x), dim3(kSThreads), (size_t)kSSmem, st, S, maps);
  }
  return check_launch("conv2d_fwd(stem)");
}

}  // namespace icaf
