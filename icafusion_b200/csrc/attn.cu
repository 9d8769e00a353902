// Bidirectional DMFF cross-attention core (models/common.py:670-684) as a flash-style tcgen05 kernel:
//   out_vis = softmax(q_ir k_vis^T / sqrt(d)) v_vis        out_ir = softmax(q_vis k_ir^T / sqrt(d)) v_ir
// The N x N score matrix never leaves the SM: S = Q K^T is produced by the tensor core into TMEM, the 128
// softmax threads (one per query row == TMEM lane) run the online softmax straight out of TMEM, write P (fp16)
// into 128B-swizzled shared memory, and a second UMMA computes P V into TMEM; running (max, sum, acc) live in
// registers.  Q / K / V^T tiles are staged by TMA (cp.async.bulk.tensor) from a producer warp, K / V^T double-buffered.
//
// Layouts (see include/icaf_b200.h): either qkv (B, Npad, 3C) = [q | k | v] rows as ONE fused projection emits them (V is
// then consumed as an MN-major UMMA operand straight from its token-major tile), or qk (B, Npad, 2C) + vt (C, B*Npad) = V^T
// (K-major); out (B, Npad, C).
// CTA = (128-query tile, batch*head, direction).
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "icaf_internal.cuh"

namespace icaf {

constexpr int kQT = 128;    // queries per CTA (UMMA M)

struct AttnParams {
  const __half* qk[2];   // [0]=vis, [1]=ir
  const __half* vt[2];   // NULL in the fused-qkv form
  __half* out[2];
  int B, N, n_pad, C, heads;
  int ld;                // row pitch of qk: 2C, or 3C in the fused-qkv form ([q | k | v])
  float p_drop;          // training forward only (TRAIN instantiation): dropout on the probabilities (common.py:677,680)
  uint32_t seed;
  const uint32_t* seed_off;   // optional device-side offset (icaf_set_seed_offset)
  float scale_log2;      // log2(e) / sqrt(d)
};

// ---------------------------------------------------------------------------------------------------
// The kernel: Q / K / V^T tiles arrive by cp.async.bulk.tensor issued from a dedicated producer warp,
// so the 128 softmax threads do nothing but softmax; two CTAs share an SM for head dims <= 64 (one CTA's softmax overlaps
// the other's MMAs and loads -- with a single CTA every SM sub-partition holds exactly one softmax warp and every TMEM /
// MUFU latency is exposed).  192 threads: warps 0-3 softmax (thread = query row = TMEM lane), warp 4 TMEM + MMA issue,
// warp 5 TMA producer.
//   Q / K tiles: 2-D boxes (min(D,64) columns x 128 token rows) of the (B*Npad, 2C) projection matrix -> K-major rows of
//                32 / 64 / 128 bytes with the matching swizzle (D = 128: two 64-column blocks)
//   V^T tiles  : two boxes (64 keys x D feature rows) of the (C, B*Npad) matrix -> K-major SW128 (keys are the K dim of PV)
//   V tiles    : (VF, fused [q|k|v] rows) boxes like K's at column 2C + head*D -> rows = keys, i.e. an MN-major B operand
// Rows / keys past the tensor are zero-filled by the TMA unit; keys in [N, ...) are masked in the softmax.
struct AttnMaps {
  CUtensorMap qk[2];   // [0] = vis, [1] = ir : (B*Npad rows, 2C | 3C cols), box (min(D,64), 128) -- Q tiles
  CUtensorMap kv[2];   // same matrices, box (min(D,64), KV) -- K tiles (and V tiles in the fused form)
  CUtensorMap vt[2];   // split form: (C rows, B*Npad cols), box (64, D)
};

// KV = keys per tile (UMMA N of S, K of PV).  Head dims 16 / 32 run 64-key tiles: S + O then fit 128 TMEM columns and the
// small tiles let 4 / 3 CTAs share an SM -- these head dims are exp-bound, and with one or two CTAs per SM every TMEM load,
// MUFU and barrier latency of the single softmax warp per sub-partition is exposed (probe: profiles/r02_attn_probe_*).
template <int D>
struct AttnCfg {
  static constexpr int kKV = D <= 32 ? 64 : 128;
  static constexpr int kCtas = D == 16 ? 4 : (D == 32 ? 3 : (D == 64 ? 2 : 1));
};

template <int D, int KV>
struct AttnSmemT {
  static constexpr int kKB = (D + 63) / 64;                 // 64-wide column blocks of the head dim
  static constexpr int kRowB = D >= 64 ? 128 : D * 2;       // bytes per staged Q / K row (= swizzle span)
  static constexpr int kQBytes = kKB * kQT * kRowB;
  static constexpr int kKBytes = kKB * KV * kRowB;          // per buffer
  static constexpr int kVBytes = (KV / 64) * D * 128;       // per buffer: 64-key blocks of D rows (= KV rows of D halfs)
  static constexpr int kPBytes = (KV / 64) * kQT * 128;
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kQOff + kQBytes;
  static constexpr int kVOff = kKOff + 2 * kKBytes;
  static constexpr int kPOff = kVOff + 2 * kVBytes;
  static constexpr int kBarOff = kPOff + kPBytes;
  static constexpr int kCtas = AttnCfg<D>::kCtas;           // CTAs per SM
  // two CTAs of the D = 64 variant fill the SM to the byte: no alignment slack (the kernel checks its base is 1024-aligned)
  static constexpr bool kSlack = (kBarOff + 128 + 1024) * kCtas + kCtas * 1024 <= 228 * 1024;
  static constexpr int kTotal = kBarOff + 128 + (kSlack ? 1024 : 0);
  static constexpr int kTmemCols = KV + D <= 128 ? 128 : 256;   // S (KV) + O (D)
  static_assert(kQBytes % 1024 == 0 && kKBytes % 1024 == 0 && kVBytes % 1024 == 0, "tiles must keep 1024-byte alignment");
};

__device__ __forceinline__ float fast_exp2_t(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// TRAIN: dropout on the attention probabilities (the row sum still runs over the un-dropped values, like
// `att = softmax(..); att = attn_drop(att)`); its own instantiation, so the inference kernels keep their size.
template <int D, bool VF, bool TRAIN = false>
__global__ void __launch_bounds__(192, AttnCfg<D>::kCtas) cross_attn_tma_kernel(const AttnParams P, const __grid_constant__ AttnMaps M) {
  constexpr int kKV = AttnCfg<D>::kKV;
  using L = AttnSmemT<D, kKV>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t sbase = L::kSlack ? ((smem_u32(smem_raw) + 1023u) & ~1023u) : smem_u32(smem_raw);
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar = sbase + L::kBarOff;
  const uint32_t q_full = bar, s_full = bar + 8, p_full = bar + 16, o_full = bar + 24;
  auto kv_full = [&](int i) { return bar + 32u + 8u * i; };
  auto kv_empty = [&](int i) { return bar + 48u + 8u * i; };
  const uint32_t tmem_slot = bar + 64;

  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5;
  const int dir = blockIdx.z;                 // 0: out_vis (q_ir, k_vis, v_vis)   1: out_ir (q_vis, k_ir, v_ir)
  const int b = blockIdx.y / P.heads, head = blockIdx.y % P.heads;
  const int q0 = blockIdx.x * kQT;
  const int C = P.C, N = P.N, n_pad = P.n_pad;
  const int nkv = (N + kKV - 1) / kKV;

  if (tid == 0) {
    if (!L::kSlack && (sbase & 1023u)) {
      printf("icaf: cross_attn_tma_kernel<%d>: dynamic shared memory base 0x%x is not 1024-byte aligned\n", D, sbase);
      __trap();
    }
    mbar_init(q_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1);
    mbar_init(kv_full(0), 1); mbar_init(kv_full(1), 1);
    mbar_init(kv_empty(0), 1); mbar_init(kv_empty(1), 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<L::kTmemCols>(tmem_slot);
  if (warp == 5 && lane_id() == 0) {
    tma_prefetch_desc(dir == 0 ? &M.qk[1] : &M.qk[0]);
    tma_prefetch_desc(dir == 0 ? &M.kv[0] : &M.kv[1]);
    if (!VF) tma_prefetch_desc(dir == 0 ? &M.vt[0] : &M.vt[1]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();           // everything above overlapped the previous kernel's tail; q/k/v are its outputs
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + L::kBarOff + 64);
  const uint32_t tmem_S = tmem, tmem_O = tmem + kKV;

  if (warp < 4) {
    // ------------------------------------------------------------------ online softmax (TMEM lanes = query rows)
    const int row = tid;
    const int qn = q0 + row;
    const uint32_t lane_off = uint32_t(warp * 32) << 16;
    const float sl2 = P.scale_log2;
    float m_run = -INFINITY, l_run = 0.f;
    float acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = 0.f;
    uint8_t* prow0 = sgen + L::kPOff + row * 128;
    const int rsw = row & 7;

    for (int j = 0; j < nkv; ++j) {
      const int kv0 = j * kKV;
      const bool full = kv0 + kKV <= N;       // every key of the tile is valid: no masking
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float mx = -INFINITY;
#pragma unroll 1
      for (int cb = 0; cb < kKV; cb += 32) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_S + lane_off + cb, r);
        tmem_ld_wait();
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (kv0 + cb + i < N) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
      const float m_new = fmaxf(m_run, mx);               // finite: every tile has >= 1 valid key
      const float corr = fast_exp2_t((m_run - m_new) * sl2);
      const float moff = m_new * sl2;
      float rs = 0.f;
#pragma unroll 1
      for (int cb = 0; cb < kKV; cb += 32) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_S + lane_off + cb, r);
        tmem_ld_wait();
        uint32_t pk[16];
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = fast_exp2_t(fmaf(__uint_as_float(r[i]), sl2, -moff));
            const float p1 = fast_exp2_t(fmaf(__uint_as_float(r[i + 1]), sl2, -moff));
            __half2 h = __floats2half2_rn(p0, p1);
            const float2 hf = __half22float2(h);            // sum what the PV MMA will actually see
            rs += hf.x + hf.y;
            if (TRAIN) {
              const float ks = 1.f / (1.f - P.p_drop);
              const bool k0 = attn_keep(P.seed + (P.seed_off ? __ldg(P.seed_off) : 0u), dir, blockIdx.y, qn, kv0 + cb + i, P.p_drop), k1 = attn_keep(P.seed + (P.seed_off ? __ldg(P.seed_off) : 0u), dir, blockIdx.y, qn, kv0 + cb + i + 1, P.p_drop);
              h = __floats2half2_rn(k0 ? hf.x * ks : 0.f, k1 ? hf.y * ks : 0.f);
            }
            pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&h);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = (kv0 + cb + i < N) ? fast_exp2_t(fmaf(__uint_as_float(r[i]), sl2, -moff)) : 0.f;
            const float p1 = (kv0 + cb + i + 1 < N) ? fast_exp2_t(fmaf(__uint_as_float(r[i + 1]), sl2, -moff)) : 0.f;
            __half2 h = __floats2half2_rn(p0, p1);
            const float2 hf = __half22float2(h);
            rs += hf.x + hf.y;
            if (TRAIN) {
              const float ks = 1.f / (1.f - P.p_drop);
              const bool k0 = attn_keep(P.seed + (P.seed_off ? __ldg(P.seed_off) : 0u), dir, blockIdx.y, qn, kv0 + cb + i, P.p_drop), k1 = attn_keep(P.seed + (P.seed_off ? __ldg(P.seed_off) : 0u), dir, blockIdx.y, qn, kv0 + cb + i + 1, P.p_drop);
              h = __floats2half2_rn(k0 ? hf.x * ks : 0.f, k1 ? hf.y * ks : 0.f);
            }
            pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&h);
          }
        }
        // P row -> K-major SW128 smem (block = cb/64, 16-byte chunks (cb%64)/8 ..)
        uint8_t* prow = prow0 + (cb >> 6) * (kQT * 128);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = ((cb & 63) >> 3) + q;
          *reinterpret_cast<uint4*>(prow + ((c ^ rsw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      l_run = l_run * corr + rs;
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      // ---- O_tile = P V ; fold into the running accumulator ----
      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int cb = 0; cb < D; cb += 32) {
        if (D - cb >= 32) {
          uint32_t r[32];
          __syncwarp();
          tmem_ld32(tmem_O + lane_off + cb, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[cb + i] = fmaf(acc[cb + i], corr, __uint_as_float(r[i]));
        } else {
          uint32_t r[16];
          __syncwarp();
          tmem_ld16(tmem_O + lane_off + cb, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[cb + i] = fmaf(acc[cb + i], corr, __uint_as_float(r[i]));
        }
      }
      tc_fence_before();
    }
    // ---- normalise and store (heads merged: column head*D) ; zero the pad rows ----
    if (qn < n_pad) {
      const float inv = qn < N ? 1.f / l_run : 0.f;
      __half* o = (dir == 0 ? P.out[0] : P.out[1]) + (size_t(b) * n_pad + qn) * C + head * D;
#pragma unroll
      for (int i = 0; i < D; i += 8) {
        uint4 v;
        v.x = pack_half2(acc[i] * inv, acc[i + 1] * inv);
        v.y = pack_half2(acc[i + 2] * inv, acc[i + 3] * inv);
        v.z = pack_half2(acc[i + 4] * inv, acc[i + 5] * inv);
        v.w = pack_half2(acc[i + 6] * inv, acc[i + 7] * inv);
        *reinterpret_cast<uint4*>(o + i) = v;
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_f16(kQT, kKV);
    constexpr uint32_t idesc_o = umma_idesc_f16_major(kQT, D, false, VF);      // VF: V tile is MN-major (rows = keys)
    mbar_wait(q_full, 0);
    for (int j = 0; j < nkv; ++j) {
      const int buf = j & 1;
      mbar_wait(kv_full(buf), (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint64_t ad = umma_desc_kmajor(sbase + L::kQOff + (k >> 2) * (kQT * 128) + (k & 3) * 32, L::kRowB);
          const uint64_t bd = umma_desc_kmajor(sbase + L::kKOff + buf * L::kKBytes + (k >> 2) * (kKV * 128) + (k & 3) * 32, L::kRowB);
          umma_f16_ss(tmem_S, ad, bd, idesc_s, k != 0);
        }
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kKV / 16; ++k) {
          const uint64_t ad = umma_desc_sw128(sbase + L::kPOff + (k >> 2) * (kQT * 128)) + uint64_t(2 * (k & 3));
          const uint64_t bd = VF ? umma_desc_mnmajor(sbase + L::kVOff + buf * L::kVBytes + k * 16 * L::kRowB, L::kRowB, kKV * 128)
                                 : umma_desc_sw128(sbase + L::kVOff + buf * L::kVBytes + (k >> 2) * (D * 128)) + uint64_t(2 * (k & 3));
          umma_f16_ss(tmem_O, ad, bd, idesc_o, k != 0);
        }
        umma_commit(o_full);
        umma_commit(kv_empty(buf));
      }
      __syncwarp();
    }
  } else if (lane_id() == 0) {
    // ------------------------------------------------------------------ TMA producer (one thread)
    const CUtensorMap* mq = dir == 0 ? &M.qk[1] : &M.qk[0];
    const CUtensorMap* mk = dir == 0 ? &M.kv[0] : &M.kv[1];
    const CUtensorMap* mv = dir == 0 ? &M.vt[0] : &M.vt[1];
    (void)mv;
    const int row_b = b * n_pad;
    mbar_arrive_expect_tx(q_full, L::kQBytes);
#pragma unroll
    for (int kb = 0; kb < L::kKB; ++kb)
      tma_load_2d(sbase + L::kQOff + kb * (kQT * 128), mq, q_full, head * D + kb * 64, row_b + q0);
    for (int j = 0; j < nkv; ++j) {
      const int buf = j & 1;
      if (j >= 2) mbar_wait(kv_empty(buf), ((j >> 1) & 1) ^ 1);     // PV(j-2) has drained this buffer
      mbar_arrive_expect_tx(kv_full(buf), L::kKBytes + L::kVBytes);
#pragma unroll
      for (int kb = 0; kb < L::kKB; ++kb)
        tma_load_2d(sbase + L::kKOff + buf * L::kKBytes + kb * (kKV * 128), mk, kv_full(buf), C + head * D + kb * 64, row_b + j * kKV);
      if (VF) {
#pragma unroll
        for (int kb = 0; kb < L::kKB; ++kb)
          tma_load_2d(sbase + L::kVOff + buf * L::kVBytes + kb * (kKV * 128), mk, kv_full(buf), 2 * C + head * D + kb * 64, row_b + j * kKV);
      } else {
#pragma unroll
        for (int kb = 0; kb < kKV / 64; ++kb)
          tma_load_2d(sbase + L::kVOff + buf * L::kVBytes + kb * (D * 128), mv, kv_full(buf), row_b + j * kKV + kb * 64, head * D);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<L::kTmemCols>(tmem);
  }
}

// ---------------------------------------------------------------------------------------------------
// CUDA-core reference, one thread per (query, head, batch, direction). Tests only.
__global__ void cross_attn_simt_kernel(const AttnParams P) {
  pdl_launch_dependents();
  pdl_wait();
  const int d = P.C / P.heads;
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = 2LL * P.B * P.heads * P.n_pad;
  if (idx >= total) return;
  int qn = int(idx % P.n_pad);
  long long t = idx / P.n_pad;
  int head = int(t % P.heads); t /= P.heads;
  int b = int(t % P.B);
  int dir = int(t / P.B);
  const __half* qsrc = dir == 0 ? P.qk[1] : P.qk[0];
  const __half* ksrc = dir == 0 ? P.qk[0] : P.qk[1];
  const __half* vsrc = dir == 0 ? P.vt[0] : P.vt[1];
  __half* o = (dir == 0 ? P.out[0] : P.out[1]) + (size_t(b) * P.n_pad + qn) * P.C + head * d;
  if (qn >= P.N) {
    for (int i = 0; i < d; ++i) o[i] = __float2half(0.f);
    return;
  }
  const __half* q = qsrc + (size_t(b) * P.n_pad + qn) * P.ld + head * d;
  float mx = -INFINITY;
  for (int k = 0; k < P.N; ++k) {
    const __half* kp = ksrc + (size_t(b) * P.n_pad + k) * P.ld + P.C + head * d;
    float s = 0.f;
    for (int i = 0; i < d; ++i) s += __half2float(q[i]) * __half2float(kp[i]);
    mx = fmaxf(mx, s);
  }
  float l = 0.f;
  float acc[128];
  for (int i = 0; i < d; ++i) acc[i] = 0.f;
  for (int k = 0; k < P.N; ++k) {
    const __half* kp = ksrc + (size_t(b) * P.n_pad + k) * P.ld + P.C + head * d;
    float s = 0.f;
    for (int i = 0; i < d; ++i) s += __half2float(q[i]) * __half2float(kp[i]);
    float p = exp2f((s - mx) * P.scale_log2);
    l += p;
    for (int i = 0; i < d; ++i)
      acc[i] += p * __half2float(vsrc ? vsrc[size_t(head * d + i) * (size_t(P.B) * P.n_pad) + size_t(b) * P.n_pad + k]
                                      : kp[P.C + i]);                     // fused form: v sits one C further in the key's row
  }
  for (int i = 0; i < d; ++i) o[i] = __float2half_rn(acc[i] / l);
}

static int fill_attn(const void* qk_vis, const void* qk_ir, const void* vt_vis, const void* vt_ir, void* out_vis,
                     void* out_ir, int B, int N, int n_pad, int C, int heads, AttnParams& P) {
  // vt_* both NULL: the fused form, qk_* are (B, Npad, 3C) [q | k | v] matrices
  if (!qk_vis || !qk_ir || (!vt_vis != !vt_ir) || !out_vis || !out_ir) return set_error(ICAF_ERR_BAD_ARG, "cross_attention: null pointer");
  if (B < 1 || N < 1 || n_pad < N || n_pad % 8 || heads < 1 || C % heads) return set_error(ICAF_ERR_BAD_ARG, "cross_attention: bad shape");
  int d = C / heads;
  if (d != 16 && d != 32 && d != 64 && d != 128) return set_error(ICAF_ERR_UNSUPPORTED, "cross_attention: head dim must be 16/32/64/128");
  P.qk[0] = (const __half*)qk_vis; P.qk[1] = (const __half*)qk_ir;
  P.vt[0] = (const __half*)vt_vis; P.vt[1] = (const __half*)vt_ir;
  P.out[0] = (__half*)out_vis; P.out[1] = (__half*)out_ir;
  P.B = B; P.N = N; P.n_pad = n_pad; P.C = C; P.heads = heads;
  P.ld = vt_vis ? 2 * C : 3 * C;
  P.p_drop = 0.f; P.seed = 0u; P.seed_off = nullptr;
  P.scale_log2 = 1.4426950408889634f / sqrtf(float(d));   // 1/sqrt(d_k), common.py:670
  return ICAF_OK;
}

template <int D, bool VF, bool TRAIN = false>
static int launch_attn_tma(const AttnParams& P, cudaStream_t st) {
  constexpr int kKV = AttnCfg<D>::kKV;
  using L = AttnSmemT<D, kKV>;
  static bool configured[kMaxDevices] = {false};
  if (int rc = configure_smem(cross_attn_tma_kernel<D, VF, TRAIN>, L::kTotal, configured, "cross_attention: cudaFuncSetAttribute")) return rc;
  AttnMaps maps;
  memset(&maps, 0, sizeof(maps));
  const uint64_t rows = uint64_t(P.B) * P.n_pad;
  for (int i = 0; i < 2; ++i) {
    int rc = encode_tmap_2d(&maps.qk[i], P.qk[i], uint64_t(P.ld), rows, uint64_t(P.ld) * 2, D < 64 ? D : 64, kQT);
    if (rc) return rc;
    rc = encode_tmap_2d(&maps.kv[i], P.qk[i], uint64_t(P.ld), rows, uint64_t(P.ld) * 2, D < 64 ? D : 64, kKV);
    if (rc) return rc;
    if (!VF) {
      rc = encode_tmap_2d(&maps.vt[i], P.vt[i], rows, uint64_t(P.C), rows * 2, 64, D);
      if (rc) return rc;
    }
  }
  dim3 grid((P.n_pad + kQT - 1) / kQT, P.B * P.heads, 2);
  launch_k(cross_attn_tma_kernel<D, VF, TRAIN>, dim3(grid), dim3(192), L::kTotal, st, P, maps);
  return check_launch("cross_attention");
}

template <bool VF>
static int dispatch_attn(const AttnParams& P, cudaStream_t st) {
  switch (P.C / P.heads) {
    case 16: return launch_attn_tma<16, VF>(P, st);
    case 32: return launch_attn_tma<32, VF>(P, st);
    case 64: return launch_attn_tma<64, VF>(P, st);
    default: return launch_attn_tma<128, VF>(P, st);
  }
}

}  // namespace icaf

using namespace icaf;

extern "C" int icaf_cross_attention(const void* qk_vis, const void* qk_ir, const void* vt_vis, const void* vt_ir,
                                    void* out_vis, void* out_ir, int B, int N, int n_pad, int C, int heads, void* stream) {
  AttnParams P;
  int rc = fill_attn(qk_vis, qk_ir, vt_vis, vt_ir, out_vis, out_ir, B, N, n_pad, C, heads, P);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if ((uint64_t(B) * n_pad * 2) % 16 || (reinterpret_cast<uintptr_t>(vt_vis) & 15) || (reinterpret_cast<uintptr_t>(qk_vis) & 15) ||
      (reinterpret_cast<uintptr_t>(vt_ir) & 15) || (reinterpret_cast<uintptr_t>(qk_ir) & 15))
    return set_error(ICAF_ERR_BAD_ARG, "cross_attention: TMA needs 16-byte aligned tensors and row pitches");
  return vt_vis ? dispatch_attn<false>(P, st) : dispatch_attn<true>(P, st);
}

extern "C" int icaf_cross_attention_train(const void* qkv_vis, const void* qkv_ir, void* out_vis, void* out_ir, int B, int N, int n_pad, int C, int heads,
                                          float p_drop, uint32_t seed, void* stream) {
  AttnParams P;
  int rc = fill_attn(qkv_vis, qkv_ir, nullptr, nullptr, out_vis, out_ir, B, N, n_pad, C, heads, P);
  if (rc) return rc;
  if (!(p_drop >= 0.f && p_drop < 1.f)) return set_error(ICAF_ERR_BAD_ARG, "cross_attention_train: dropout probability must be in [0, 1)");
  if ((uint64_t(B) * n_pad * 2) % 16 || (reinterpret_cast<uintptr_t>(qkv_vis) & 15) || (reinterpret_cast<uintptr_t>(qkv_ir) & 15))
    return set_error(ICAF_ERR_BAD_ARG, "cross_attention_train: TMA needs 16-byte aligned tensors and row pitches");
  P.p_drop = p_drop; P.seed = seed; P.seed_off = seed_offset_ptr();
  cudaStream_t st = (cudaStream_t)stream;
  if (p_drop == 0.f) return dispatch_attn<true>(P, st);
  switch (C / heads) {
    case 16: return launch_attn_tma<16, true, true>(P, st);
    case 32: return launch_attn_tma<32, true, true>(P, st);
    case 64: return launch_attn_tma<64, true, true>(P, st);
    default: return launch_attn_tma<128, true, true>(P, st);
  }
}

extern "C" int icaf_cross_attention_simt(const void* qk_vis, const void* qk_ir, const void* vt_vis, const void* vt_ir,
                                         void* out_vis, void* out_ir, int B, int N, int n_pad, int C, int heads,
                                         void* stream) {
  AttnParams P;
  int rc = fill_attn(qk_vis, qk_ir, vt_vis, vt_ir, out_vis, out_ir, B, N, n_pad, C, heads, P);
  if (rc) return rc;
  long long total = 2LL * B * heads * n_pad;
  launch_k(cross_attn_simt_kernel, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, P);
  return check_launch("cross_attention_simt");
}
