// Shared pieces of the implicit-GEMM conv kernels (one-tile-per-CTA kernel in conv_gemm.cu, persistent kernel in
// conv_persist.cu): problem descriptors, TMA descriptor bundle, epilogue chunk.
#pragma once
#include "icaf_internal.cuh"

namespace icaf {

constexpr int BM = 128;
constexpr int BK = 64;              // 64 halfs = one 128-byte swizzle atom row
constexpr int kLag = 2;             // cp.async groups kept in flight per gather thread
constexpr int kThreads = 192;

enum AMode { A_GATHER = 0, A_TMA2D = 1, A_TMA4D = 2 };

struct ConvProblem {
  const __half* x; const float* bias; const __half* res; __half* y;
  const float* alpha; const float* beta;
  long long x_ld, res_ld, y_ld;
  const float2* ln_stats;   // LN fold: (sum, sum of squares) partials of every INPUT row, [M][ln_parts]
  const float* ln_s;        // LN fold: column sums of the gamma-folded filter, fp32 [Cout]
  float2* stats_out;        // EMIT_STATS: (sum, sum of squares) partials of every OUTPUT row, [M][ceil(N/32)]
};
// ICAF_DBG(P, bit): compile-time false unless the library is built with -DICAF_PROBE (ICAF_PROBE=1 python -m icafusion_b200.build)
#ifdef ICAF_PROBE
#define ICAF_DBG(P, bit) (((P).dbg & (bit)) != 0)
#else
#define ICAF_DBG(P, bit) false
#endif

struct ConvParams {
  ConvProblem p[2];
  int M, N, K, k_pad;
  int B, Hi, Wi, Cin, Ho, Wo, kh, kw, stride, pad;
  int act, epi;
  int a_mode, tw, th, tiles_x, tiles_y;   // A_TMA4D: tile = th x tw output pixels (tw*th <= 128), tiles per image
  int stages;                             // smem ring depth (runtime: deep rings for small grids, 2 CTAs/SM otherwise)
  int splits;                             // split-K factor = cluster size along x (1 = no cluster); partial sums meet in DSMEM
  int cblk;                               // A_TMA4D: channels per TMA box = min(Cin, 64); < 64 only in the persistent kernel
  int halo;                               // conv_pair.cu: 3x3/s1 layers stage three x-shifted (th+2)-row copies per channel block (0 / 1)
  int dense16;                            // host only: every problem's input has pixel pitch 16 (stem kernel eligibility)
  int ln_parts;                           // LN fold: partials per input row (0 = no fold)
  float ln_eps, ln_inv_k;                 // LN fold: epsilon, 1 / (normalised features = K)
  int dbg;                                // probe builds (-DICAF_PROBE, tools/conv_probe.py): 1 no stores, 2 no activation,
                                          // 8 no A loads, 16 no B loads, 32 no MMA; always 0 in the shipped library
};
struct ConvMaps {          // TMA descriptors, passed by value as a __grid_constant__ kernel parameter
  CUtensorMap w[2];
  CUtensorMap a[2];
};

__device__ __forceinline__ ConvProblem pick_problem(const ConvParams& P, unsigned z) {
  ConvProblem r;
  r.x = z ? P.p[1].x : P.p[0].x;          r.bias = z ? P.p[1].bias : P.p[0].bias;
  r.res = z ? P.p[1].res : P.p[0].res;    r.y = z ? P.p[1].y : P.p[0].y;
  r.alpha = z ? P.p[1].alpha : P.p[0].alpha; r.beta = z ? P.p[1].beta : P.p[0].beta;
  r.x_ld = z ? P.p[1].x_ld : P.p[0].x_ld; r.res_ld = z ? P.p[1].res_ld : P.p[0].res_ld;
  r.y_ld = z ? P.p[1].y_ld : P.p[0].y_ld;
  r.ln_stats = z ? P.p[1].ln_stats : P.p[0].ln_stats; r.ln_s = z ? P.p[1].ln_s : P.p[0].ln_s;
  r.stats_out = z ? P.p[1].stats_out : P.p[0].stats_out;
  return r;
}

// Per-row extras of an epilogue thread.  XM = 1 (LayerNorm folded into this GEMM, common.py:660,665,749-750): the filter
// carries gamma, the bias carries beta . W, and the row is normalised after the fact:
//     LN(x) . W^T + b  =  rstd * (x . W'^T  -  mean * s)  +  b'          W' = W diag(gamma), s_n = sum_k W'[n][k]
// XM = 2 (EMIT_STATS): sum and sum of squares of the fp16-rounded outputs of this row, for the LN fold of the NEXT GEMM.
struct EpiRow {
  float ln_a, ln_mu;     // rstd, mean of this thread's input row
  const float* ln_s;     // s, offset to the first column of the current pass
  float sum, sumsq;
};
__device__ __forceinline__ void epi_row_ln(EpiRow& ex, const ConvParams& P, const ConvProblem& pr, int m, bool mvalid) {
  float su = 0.f, sq = 0.f;
  if (mvalid) {
    const float2* sp = pr.ln_stats + size_t(m) * P.ln_parts;
    for (int i = 0; i < P.ln_parts; ++i) { const float2 v = __ldg(sp + i); su += v.x; sq += v.y; }
  }
  const float mu = su * P.ln_inv_k;
  ex.ln_mu = mu;
  ex.ln_a = rsqrtf(fmaxf(sq * P.ln_inv_k - mu * mu, 0.f) + P.ln_eps);
}
// slots [n_begin/32, n_end/32) of this row's partials: the first one carries the sums, the others zero
__device__ __forceinline__ void epi_row_emit(const EpiRow& ex, const ConvParams& P, const ConvProblem& pr, int m, int n_begin, int n_end) {
  const int slots = (P.N + 31) >> 5;
  float2* sp = pr.stats_out + size_t(m) * slots;
  const int s0 = n_begin >> 5, s1 = min((n_end + 31) >> 5, slots);
  for (int i = s0; i < s1; ++i) sp[i] = i == s0 ? make_float2(ex.sum, ex.sumsq) : make_float2(0.f, 0.f);
}

// One 32-column chunk of one output row: + bias, activation, residual, fp16 store.
// ACT: 0 none, 1 SiLU, 2 GELU(erf).  RES: 0 none, 1 y = act(v) + res, 2 y = alpha*res + beta*v.
template <int ACT, int RES, int XM = 0>
__device__ __forceinline__ void epi_chunk(const uint32_t (&acc)[32], const float* __restrict__ sb, float rbias,
                                          float alpha, float beta, const __half* __restrict__ rp,
                                          __half* __restrict__ yp, bool vec, int ncols, EpiRow& ex, int cb, bool do_store = true) {
  auto f = [&](int j) {
    float a = __uint_as_float(acc[j]);
    if (XM == 1) a = ex.ln_a * (a - ex.ln_mu * __ldg(ex.ln_s + cb + j));
    float t = a + sb[j] + rbias;
    if (ACT == ICAF_ACT_SILU) t = silu_f(t);
    if (ACT == ICAF_ACT_GELU) t = gelu_erf_f(t);
    return t;
  };
  if (vec) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = f(q * 8 + e);
      if (RES != 0) {
        uint4 rr = *reinterpret_cast<const uint4*>(rp + q * 8);
        const __half2* rh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 rf = __half22float2(rh[e]);
          if (RES == 2) {
            v[2 * e] = alpha * rf.x + beta * v[2 * e];
            v[2 * e + 1] = alpha * rf.y + beta * v[2 * e + 1];
          } else {
            v[2 * e] += rf.x;
            v[2 * e + 1] += rf.y;
          }
        }
      }
      uint4 o;
      o.x = pack_half2(v[0], v[1]); o.y = pack_half2(v[2], v[3]);
      o.z = pack_half2(v[4], v[5]); o.w = pack_half2(v[6], v[7]);
      if (XM == 2) {
        const __half2* oh = reinterpret_cast<const __half2*>(&o);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 w = __half22float2(oh[e]);
          ex.sum += w.x + w.y;
          ex.sumsq += w.x * w.x + w.y * w.y;
        }
      }
      if (do_store) *reinterpret_cast<uint4*>(yp + q * 8) = o;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < ncols) {
        float t = f(j);
        if (RES != 0) {
          float rf = __half2float(rp[j]);
          t = RES == 2 ? alpha * rf + beta * t : t + rf;
        }
        const __half h = __float2half_rn(t);
        if (XM == 2) { const float w = __half2float(h); ex.sum += w; ex.sumsq += w * w; }
        yp[j] = h;
      }
    }
  }
}

// Persistent-kernel flavour of epi_chunk, 16 columns of one output row.  Rows whose output (and residual) are 32-byte
// aligned are written with one 256-bit store: a full L2 sector per lane and instruction.  `sb`: 16 bias floats (smem).
// `al`: 2 = 32-byte aligned, 1 = 16-byte aligned, 0 = element-wise loads / stores of the first `ncols` columns (ragged N,
// odd pitches); the math is shared by the three so the hot loop stays small (instruction cache, tools/conv_probe.py).
template <int ACT, int RES, int XM = 0>
__device__ __forceinline__ void epi_chunk16(const uint32_t (&acc)[16], const float* __restrict__ sb, float rbias,
                                            float alpha, float beta, const __half* __restrict__ rp,
                                            __half* __restrict__ yp, int al, int ncols, bool do_store, EpiRow& ex, int cb) {
  auto act = [&](float t) {
    if (ACT == ICAF_ACT_SILU) t = silu_f(t);
    if (ACT == ICAF_ACT_GELU) t = gelu_erf_f(t);
    return t;
  };
  float v[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b4 = *reinterpret_cast<const float4*>(sb + 4 * q);
    float a0 = __uint_as_float(acc[4 * q + 0]), a1 = __uint_as_float(acc[4 * q + 1]);
    float a2 = __uint_as_float(acc[4 * q + 2]), a3 = __uint_as_float(acc[4 * q + 3]);
    if (XM == 1) {
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(ex.ln_s + cb + 4 * q));
      a0 = ex.ln_a * (a0 - ex.ln_mu * s4.x); a1 = ex.ln_a * (a1 - ex.ln_mu * s4.y);
      a2 = ex.ln_a * (a2 - ex.ln_mu * s4.z); a3 = ex.ln_a * (a3 - ex.ln_mu * s4.w);
    }
    v[4 * q + 0] = act(a0 + b4.x + rbias);
    v[4 * q + 1] = act(a1 + b4.y + rbias);
    v[4 * q + 2] = act(a2 + b4.z + rbias);
    v[4 * q + 3] = act(a3 + b4.w + rbias);
  }
  if (RES != 0) {
    uint32_t rr[8];
    if (al == 2) {
      ld_global_nc_v8(rp, rr);
    } else if (al == 1) {
      uint4 r0 = __ldg(reinterpret_cast<const uint4*>(rp));
      uint4 r1 = __ldg(reinterpret_cast<const uint4*>(rp + 8));
      rr[0] = r0.x; rr[1] = r0.y; rr[2] = r0.z; rr[3] = r0.w; rr[4] = r1.x; rr[5] = r1.y; rr[6] = r1.z; rr[7] = r1.w;
    } else {
      const unsigned short* rs = reinterpret_cast<const unsigned short*>(rp);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t lo = 2 * e < ncols ? rs[2 * e] : 0u;
        const uint32_t hi = 2 * e + 1 < ncols ? rs[2 * e + 1] : 0u;
        rr[e] = lo | (hi << 16);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float2 rf = __half22float2(*reinterpret_cast<const __half2*>(&rr[e]));
      if (RES == 2) {
        v[2 * e] = alpha * rf.x + beta * v[2 * e];
        v[2 * e + 1] = alpha * rf.y + beta * v[2 * e + 1];
      } else {
        v[2 * e] += rf.x;
        v[2 * e + 1] += rf.y;
      }
    }
  }
  uint32_t o[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = pack_half2(v[2 * e], v[2 * e + 1]);
  if (XM == 2) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (2 * e < ncols) {                         // ragged tails: only the columns that exist
        const float2 w = __half22float2(*reinterpret_cast<const __half2*>(&o[e]));
        ex.sum += w.x; ex.sumsq += w.x * w.x;
        if (2 * e + 1 < ncols) { ex.sum += w.y; ex.sumsq += w.y * w.y; }
      }
    }
  }
  if (do_store) {
    if (al == 2) {
      st_global_v8(yp, o);
    } else if (al == 1) {
      *reinterpret_cast<uint4*>(yp) = make_uint4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<uint4*>(yp + 8) = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
      unsigned short* ys = reinterpret_cast<unsigned short*>(yp);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (2 * e < ncols) ys[2 * e] = (unsigned short)(o[e] & 0xffffu);
        if (2 * e + 1 < ncols) ys[2 * e + 1] = (unsigned short)(o[e] >> 16);
      }
    }
  }
}

__device__ __forceinline__ ConvProblem pick_problem_stem(const ConvProblem (&p)[2], int z) {
  ConvProblem r = p[0];
  if (z) r = p[1];
  return r;
}

// Host-side launch plan: everything the dispatcher decides before it touches CUDA.  icaf_conv2d_plan (host only, no
// device needed) exposes it so that a CPU test can walk every layer geometry through the dispatcher's invariants.
struct ConvPlan {
  int kernel;                     // ICAF_KERNEL_TC / _PERSIST / _PAIR
  int bn;                         // output-channel tile width
  unsigned grid_x, grid_y, grid_z, cluster;
  int smem;                       // dynamic shared memory per CTA (bytes)
  int total, m_tiles, m_pairs, n_tiles;
  int sms;                        // SM count the plan was made for
};

// Each kernel family: plan_* fills the launch shape (and P.stages / P.splits) and checks the kernel's invariants without
// any CUDA call; launch_* encodes the TMA descriptors and launches exactly that plan.
// persistent kernel (conv_persist.cu), BN = 32, 64, 128, 256
template <int BN>
int plan_persist(ConvParams& P, int n_io, ConvPlan& pl);
template <int BN>
int launch_persist(const ConvParams& P, const ConvPlan& pl, const __half* const (&w)[2], const icaf_conv_geom* g, int n_io, cudaStream_t st);

// image-stem kernel (conv_stem.cu): 3x3 / s1 over the 16-channel space-to-depth frame, x-merged rows
bool stem_eligible(const icaf_conv_geom* g);
int plan_stem(ConvParams& P, const icaf_conv_geom* g, int n_io, ConvPlan& pl);
int launch_stem(const ConvParams& P, const ConvPlan& pl, const __half* const (&w)[2], const icaf_conv_geom* g, int n_io, cudaStream_t st);

// CTA-pair kernel (conv_pair.cu): 256 x BN tiles over two SMs, tcgen05.mma.cta_group::2 (BN = 64, 128, 256)
template <int BN>
int plan_pair(ConvParams& P, const icaf_conv_geom* g, int n_io, ConvPlan& pl);
template <int BN>
int launch_pair(const ConvParams& P, const ConvPlan& pl, const __half* const (&w)[2], const icaf_conv_geom* g, int n_io, cudaStream_t st);

}  // namespace icaf
