// Bidirectional DMFF cross-attention core (models/common.py:670-684) as a flash-style tcgen05 kernel:
//   out_vis = softmax(q_ir k_vis^T / sqrt(d)) v_vis        out_ir = softmax(q_vis k_ir^T / sqrt(d)) v_ir
// The N x N score matrix never leaves the SM: S = Q K^T is produced by the tensor core into TMEM, the 128
// softmax threads (one per query row == TMEM lane) run the online softmax straight out of TMEM, write P (fp16)
// into 128B-swizzled shared memory, and a second UMMA computes P V into TMEM; running (max, sum, acc) live in
// registers.  K / V^T tiles are double-buffered with cp.async so the next tile lands while the current one
// is being reduced.
//
// Layouts (see include/icaf_b200.h): qk (B, Npad, 2C) = [q | k] rows; vt (C, B*Npad) = V^T; out (B, Npad, C).
// CTA = (128-query tile, batch*head, direction); 160 threads: warps 0-3 gather + softmax, warp 4 MMA issuer.
#include <cmath>
#include <cstdlib>

#include "icaf_internal.cuh"

namespace icaf {

constexpr int kQT = 128;    // queries per CTA (UMMA M)
constexpr int kKV = 128;    // keys per tile   (UMMA N of S, K of PV)

struct AttnParams {
  const __half* qk[2];   // [0]=vis, [1]=ir
  const __half* vt[2];
  __half* out[2];
  int B, N, n_pad, C, heads;
  float scale_log2;      // log2(e) / sqrt(d)
};

template <int D>
struct AttnSmem {
  static constexpr int kKB = (D + 63) / 64;             // 64-wide K blocks of the head dim
  static constexpr int kQBytes = kKB * kQT * 128;
  static constexpr int kKBytes = kKB * kKV * 128;       // per buffer
  static constexpr int kVBytes = 2 * D * 128;           // V^T: two 64-key blocks of D rows, per buffer
  static constexpr int kPBytes = 2 * kQT * 128;
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kQOff + kQBytes;
  static constexpr int kVOff = kKOff + 2 * kKBytes;
  static constexpr int kPOff = kVOff + 2 * ((kVBytes + 1023) / 1024 * 1024);
  static constexpr int kBarOff = kPOff + kPBytes;
  static constexpr int kTotal = kBarOff + 128 + 1024;
  static constexpr int kVStride = (kVBytes + 1023) / 1024 * 1024;
};

template <int D>
__global__ void __launch_bounds__(160, 1) cross_attn_tc_kernel(const AttnParams P) {
  using L = AttnSmem<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
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
  const __half* qsrc = dir == 0 ? P.qk[1] : P.qk[0];
  const __half* ksrc = dir == 0 ? P.qk[0] : P.qk[1];
  const __half* vsrc = dir == 0 ? P.vt[0] : P.vt[1];
  __half* outp = dir == 0 ? P.out[0] : P.out[1];
  const int C = P.C, N = P.N, n_pad = P.n_pad;
  const int nkv = (N + kKV - 1) / kKV;

  if (tid == 0) {
    mbar_init(q_full, 128); mbar_init(s_full, 1); mbar_init(p_full, 128); mbar_init(o_full, 1);
    mbar_init(kv_full(0), 128); mbar_init(kv_full(1), 128);
    mbar_init(kv_empty(0), 1); mbar_init(kv_empty(1), 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();           // everything above overlapped the previous kernel's tail; q/k/v are its outputs
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + L::kBarOff + 64);
  const uint32_t tmem_S = tmem, tmem_O = tmem + kKV;

  if (warp < 4) {
    constexpr int CPR = D / 8;                // 16-byte chunks per Q/K row
    // gather helpers ---------------------------------------------------------------------------
    auto load_rows = [&](uint32_t dst_base, const __half* src, int row0, int col0) {
      // 128 token rows x D halfs from the (B,Npad,2C) projection buffer into K-major SW128 block(s)
      for (int id = tid; id < 128 * CPR; id += 128) {
        int row = id / CPR, cc = id % CPR;
        int n = row0 + row;
        bool ok = n < N;
        const __half* g = src + (size_t(b) * n_pad + (ok ? n : 0)) * (2 * C) + col0 + cc * 8;
        uint32_t dst = dst_base + uint32_t(cc >> 3) * (128u * 128u) + uint32_t(row) * 128u + (uint32_t((cc & 7) ^ (row & 7)) << 4);
        cp_async16(dst, g, ok);
      }
    };
    auto load_vt = [&](uint32_t dst_base, int kv0) {
      // D feature rows x 128 keys from V^T (C, B*Npad): two 64-key K blocks
      for (int id = tid; id < D * 16; id += 128) {
        int row = id >> 4, cc = id & 15;
        int key = kv0 + cc * 8;
        bool ok = key < n_pad;
        const __half* g = vsrc + size_t(head * D + row) * (size_t(P.B) * n_pad) + size_t(b) * n_pad + (ok ? key : 0);
        uint32_t dst = dst_base + uint32_t(cc >> 3) * uint32_t(D * 128) + uint32_t(row) * 128u + (uint32_t((cc & 7) ^ (row & 7)) << 4);
        cp_async16(dst, g, ok);
      }
    };

    load_rows(sbase + L::kQOff, qsrc, q0, head * D);                 // Q tile (q part: cols [0,C))
    cp_async_arrive_on(q_full);                                      // asynchronous arrivals: nobody blocks on the loads
    load_rows(sbase + L::kKOff, ksrc, 0, C + head * D);              // K tile 0 (k part: cols [C,2C))
    load_vt(sbase + L::kVOff, 0);
    cp_async_arrive_on(kv_full(0));

    const int row = tid;
    const int qn = q0 + row;
    const uint32_t lane_off = uint32_t(warp * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    float acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = 0.f;

    for (int j = 0; j < nkv; ++j) {
      const int kv0 = j * kKV;
      if (j + 1 < nkv) {                       // prefetch tile j+1 into the other buffer
        const int nb = (j + 1) & 1;
        mbar_wait(kv_empty(nb), (((j + 1) >> 1) & 1) ^ 1);
        load_rows(sbase + L::kKOff + nb * L::kKBytes, ksrc, kv0 + kKV, C + head * D);
        load_vt(sbase + L::kVOff + nb * L::kVStride, kv0 + kKV);
        cp_async_arrive_on(kv_full(nb));
      }
      // ---- online softmax on S (TMEM lanes = query rows) ----
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float mx = -INFINITY;
#pragma unroll 1
      for (int cb = 0; cb < kKV; cb += 32) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_S + lane_off + cb, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (kv0 + cb + i < N) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
      const float m_new = fmaxf(m_run, mx);               // finite: every tile has >= 1 valid key
      const float corr = exp2f((m_run - m_new) * P.scale_log2);
      const float moff = m_new * P.scale_log2;
      float rs = 0.f;
#pragma unroll 1
      for (int cb = 0; cb < kKV; cb += 32) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_S + lane_off + cb, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = (kv0 + cb + i < N) ? exp2f(__uint_as_float(r[i]) * P.scale_log2 - moff) : 0.f;
          float p1 = (kv0 + cb + i + 1 < N) ? exp2f(__uint_as_float(r[i + 1]) * P.scale_log2 - moff) : 0.f;
          __half2 h = __floats2half2_rn(p0, p1);
          float2 hf = __half22float2(h);                  // sum what the PV MMA will actually see
          rs += hf.x + hf.y;
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
        // P row -> K-major SW128 smem (block = cb/64, 16-byte chunks (cb%64)/8 ..)
        uint8_t* prow = sgen + L::kPOff + (cb >> 6) * (kQT * 128) + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int c = ((cb & 63) >> 3) + q;
          *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
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
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_O + lane_off + cb, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (cb + i < D) acc[cb + i] = acc[cb + i] * corr + __uint_as_float(r[i]);
      }
      tc_fence_before();
    }
    // ---- normalise and store (heads merged: column head*D) ; zero the pad rows ----
    if (qn < n_pad) {
      const float inv = qn < N ? 1.f / l_run : 0.f;
      __half* o = outp + (size_t(b) * n_pad + qn) * C + head * D;
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
  } else {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_f16(kQT, kKV);
    constexpr uint32_t idesc_o = umma_idesc_f16(kQT, D);
    mbar_wait(q_full, 0);
    for (int j = 0; j < nkv; ++j) {
      const int buf = j & 1;
      mbar_wait(kv_full(buf), (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          uint64_t ad = umma_desc_sw128(sbase + L::kQOff + (k >> 2) * (kQT * 128)) + uint64_t(2 * (k & 3));
          uint64_t bd = umma_desc_sw128(sbase + L::kKOff + buf * L::kKBytes + (k >> 2) * (kKV * 128)) + uint64_t(2 * (k & 3));
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
          uint64_t ad = umma_desc_sw128(sbase + L::kPOff + (k >> 2) * (kQT * 128)) + uint64_t(2 * (k & 3));
          uint64_t bd = umma_desc_sw128(sbase + L::kVOff + buf * L::kVStride + (k >> 2) * (D * 128)) + uint64_t(2 * (k & 3));
          umma_f16_ss(tmem_O, ad, bd, idesc_o, k != 0);
        }
        umma_commit(o_full);
        umma_commit(kv_empty(buf));
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
}

// ---------------------------------------------------------------------------------------------------
// Software-pipelined variant for head dims <= 64 (everything in yolov5s, P3/P4 of yolov5l, the whole DMFF sweep):
// S, P and O are double-buffered (TMEM: S0|S1|O0|O1 = 512 columns; smem: two P tiles), so
//   * the tensor core computes S(j+1) = Q K(j+1)^T while the softmax threads are still reducing S(j), and
//   * the threads fold O(j-1) = P(j-1) V(j-1) into their registers only after they have handed P(j) to the tensor
//     core, i.e. the PV MMA latency is hidden behind the next tile's softmax.
// exp2 goes through MUFU.EX2 directly (ex2.approx.ftz).
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int D>
struct AttnSmemP {
  static constexpr int kQBytes = kQT * 128;
  static constexpr int kKBytes = kKV * 128;             // per buffer (one 64-wide K block, D <= 64)
  static constexpr int kVBytes = 2 * D * 128;
  static constexpr int kVStride = (kVBytes + 1023) / 1024 * 1024;
  static constexpr int kPBytes = 2 * kQT * 128;         // per buffer
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kQOff + kQBytes;
  static constexpr int kVOff = kKOff + 2 * kKBytes;
  static constexpr int kPOff = kVOff + 2 * kVStride;
  static constexpr int kBarOff = kPOff + 2 * kPBytes;
  static constexpr int kTotal = kBarOff + 128 + 1024;
};

template <int D>
__global__ void __launch_bounds__(160, 1) cross_attn_pipe_kernel(const AttnParams P) {
  static_assert(D <= 64, "pipelined variant: one 64-wide K block per Q/K row");
  using L = AttnSmemP<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const uint32_t bar = sbase + L::kBarOff;
  const uint32_t q_full = bar;
  auto s_full = [&](int i) { return bar + 8u + 8u * i; };
  auto p_full = [&](int i) { return bar + 24u + 8u * i; };
  auto o_full = [&](int i) { return bar + 40u + 8u * i; };
  auto kv_full = [&](int i) { return bar + 56u + 8u * i; };
  auto kv_empty = [&](int i) { return bar + 72u + 8u * i; };
  const uint32_t tmem_slot = bar + 88;

  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5;
  const int dir = blockIdx.z;
  const int b = blockIdx.y / P.heads, head = blockIdx.y % P.heads;
  const int q0 = blockIdx.x * kQT;
  const __half* qsrc = dir == 0 ? P.qk[1] : P.qk[0];
  const __half* ksrc = dir == 0 ? P.qk[0] : P.qk[1];
  const __half* vsrc = dir == 0 ? P.vt[0] : P.vt[1];
  __half* outp = dir == 0 ? P.out[0] : P.out[1];
  const int C = P.C, N = P.N, n_pad = P.n_pad;
  const int nkv = (N + kKV - 1) / kKV;

  if (tid == 0) {
    mbar_init(q_full, 128);
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full(i), 1); mbar_init(p_full(i), 128); mbar_init(o_full(i), 1);
      mbar_init(kv_full(i), 128); mbar_init(kv_empty(i), 1);
    }
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sgen + L::kBarOff + 88);
  auto tmem_S = [&](int i) { return tmem + uint32_t(i) * kKV; };
  auto tmem_O = [&](int i) { return tmem + 256u + uint32_t(i) * 128u; };

  if (warp < 4) {
    constexpr int CPR = D / 8;
    auto load_rows = [&](uint32_t dst_base, const __half* src, int row0, int col0) {
      for (int id = tid; id < 128 * CPR; id += 128) {
        int row = id / CPR, cc = id % CPR;
        int n = row0 + row;
        bool ok = n < N;
        const __half* g = src + (size_t(b) * n_pad + (ok ? n : 0)) * (2 * C) + col0 + cc * 8;
        cp_async16(dst_base + uint32_t(row) * 128u + (uint32_t(cc ^ (row & 7)) << 4), g, ok);
      }
    };
    auto load_vt = [&](uint32_t dst_base, int kv0) {
      for (int id = tid; id < D * 16; id += 128) {
        int row = id >> 4, cc = id & 15;
        int key = kv0 + cc * 8;
        bool ok = key < n_pad;
        const __half* g = vsrc + size_t(head * D + row) * (size_t(P.B) * n_pad) + size_t(b) * n_pad + (ok ? key : 0);
        cp_async16(dst_base + uint32_t(cc >> 3) * uint32_t(D * 128) + uint32_t(row) * 128u + (uint32_t((cc & 7) ^ (row & 7)) << 4), g, ok);
      }
    };
    load_rows(sbase + L::kQOff, qsrc, q0, head * D);
    cp_async_arrive_on(q_full);                // asynchronous arrivals: nobody blocks on the loads
    load_rows(sbase + L::kKOff, ksrc, 0, C + head * D);
    load_vt(sbase + L::kVOff, 0);
    cp_async_arrive_on(kv_full(0));

    const int row = tid;
    const int qn = q0 + row;
    const uint32_t lane_off = uint32_t(warp * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f, corr_prev = 0.f;
    float acc[D];
#pragma unroll
    for (int i = 0; i < D; ++i) acc[i] = 0.f;

    auto fold_o = [&](int t, float corr) {           // acc = acc*corr + O(t)
      mbar_wait(o_full(t & 1), (t >> 1) & 1);
      tc_fence_after();
      uint32_t r[32];
#pragma unroll
      for (int cb = 0; cb < D; cb += 32) {
        __syncwarp();
        tmem_ld32(tmem_O(t & 1) + lane_off + cb, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (cb + i < D) acc[cb + i] = acc[cb + i] * corr + __uint_as_float(r[i]);
      }
      tc_fence_before();
    };

    for (int j = 0; j < nkv; ++j) {
      const int kv0 = j * kKV, sb = j & 1;
      if (j + 1 < nkv) {                       // prefetch tile j+1 (its buffer is free once PV(j-1) has completed)
        const int nb = (j + 1) & 1;
        mbar_wait(kv_empty(nb), (((j + 1) >> 1) & 1) ^ 1);
        load_rows(sbase + L::kKOff + nb * L::kKBytes, ksrc, kv0 + kKV, C + head * D);
        load_vt(sbase + L::kVOff + nb * L::kVStride, kv0 + kKV);
        cp_async_arrive_on(kv_full(nb));       // lands during this tile's softmax; the MMA warp issues S(j+1) right then
      }
      mbar_wait(s_full(sb), (j >> 1) & 1);
      tc_fence_after();
      float mx = -INFINITY;
#pragma unroll 1
      for (int cb = 0; cb < kKV; cb += 32) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_S(sb) + lane_off + cb, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (kv0 + cb + i < N) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
      const float m_new = fmaxf(m_run, mx);
      const float corr = fast_exp2((m_run - m_new) * P.scale_log2);
      const float moff = m_new * P.scale_log2;
      float rs = 0.f;
      uint8_t* pbuf = sgen + L::kPOff + sb * L::kPBytes;
#pragma unroll 1
      for (int cb = 0; cb < kKV; cb += 32) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_S(sb) + lane_off + cb, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = (kv0 + cb + i < N) ? fast_exp2(fmaf(__uint_as_float(r[i]), P.scale_log2, -moff)) : 0.f;
          float p1 = (kv0 + cb + i + 1 < N) ? fast_exp2(fmaf(__uint_as_float(r[i + 1]), P.scale_log2, -moff)) : 0.f;
          __half2 h = __floats2half2_rn(p0, p1);
          float2 hf = __half22float2(h);
          rs += hf.x + hf.y;
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
        uint8_t* prow = pbuf + (cb >> 6) * (kQT * 128) + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          int c = ((cb & 63) >> 3) + q;
          *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      l_run = l_run * corr + rs;
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full(sb));                 // P(j) handed over; S buffer sb is free for S(j+2)
      if (j >= 1) fold_o(j - 1, corr_prev);    // PV(j-1) ran while this tile's softmax was computed
      corr_prev = corr;
    }
    fold_o(nkv - 1, corr_prev);
    if (qn < n_pad) {
      const float inv = qn < N ? 1.f / l_run : 0.f;
      __half* o = outp + (size_t(b) * n_pad + qn) * C + head * D;
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
  } else {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = umma_idesc_f16(kQT, kKV);
    constexpr uint32_t idesc_o = umma_idesc_f16(kQT, D);
    auto issue_s = [&](int t) {                // S(t) = Q K(t)^T into S buffer t&1
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          uint64_t ad = umma_desc_sw128(sbase + L::kQOff) + uint64_t(2 * k);
          uint64_t bd = umma_desc_sw128(sbase + L::kKOff + (t & 1) * L::kKBytes) + uint64_t(2 * k);
          umma_f16_ss(tmem_S(t & 1), ad, bd, idesc_s, k != 0);
        }
        umma_commit(s_full(t & 1));
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(kv_full(0), 0);
    tc_fence_after();
    issue_s(0);
    for (int j = 0; j < nkv; ++j) {
      const int buf = j & 1;
      if (j + 1 < nkv) {                       // next tile's scores while the softmax threads work on S(j)
        mbar_wait(kv_full((j + 1) & 1), ((j + 1) >> 1) & 1);
        tc_fence_after();
        issue_s(j + 1);                        // S buffer (j+1)&1 was released by p_full(j-1), waited for last iteration
      }
      mbar_wait(p_full(buf), (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kKV / 16; ++k) {
          uint64_t ad = umma_desc_sw128(sbase + L::kPOff + buf * L::kPBytes + (k >> 2) * (kQT * 128)) + uint64_t(2 * (k & 3));
          uint64_t bd = umma_desc_sw128(sbase + L::kVOff + buf * L::kVStride + (k >> 2) * (D * 128)) + uint64_t(2 * (k & 3));
          umma_f16_ss(tmem_O(buf), ad, bd, idesc_o, k != 0);
        }
        umma_commit(o_full(buf));
        umma_commit(kv_empty(buf));
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
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
  const __half* q = qsrc + (size_t(b) * P.n_pad + qn) * (2 * P.C) + head * d;
  float mx = -INFINITY;
  for (int k = 0; k < P.N; ++k) {
    const __half* kp = ksrc + (size_t(b) * P.n_pad + k) * (2 * P.C) + P.C + head * d;
    float s = 0.f;
    for (int i = 0; i < d; ++i) s += __half2float(q[i]) * __half2float(kp[i]);
    mx = fmaxf(mx, s);
  }
  float l = 0.f;
  float acc[128];
  for (int i = 0; i < d; ++i) acc[i] = 0.f;
  for (int k = 0; k < P.N; ++k) {
    const __half* kp = ksrc + (size_t(b) * P.n_pad + k) * (2 * P.C) + P.C + head * d;
    float s = 0.f;
    for (int i = 0; i < d; ++i) s += __half2float(q[i]) * __half2float(kp[i]);
    float p = exp2f((s - mx) * P.scale_log2);
    l += p;
    for (int i = 0; i < d; ++i)
      acc[i] += p * __half2float(vsrc[size_t(head * d + i) * (size_t(P.B) * P.n_pad) + size_t(b) * P.n_pad + k]);
  }
  for (int i = 0; i < d; ++i) o[i] = __float2half_rn(acc[i] / l);
}

static int fill_attn(const void* qk_vis, const void* qk_ir, const void* vt_vis, const void* vt_ir, void* out_vis,
                     void* out_ir, int B, int N, int n_pad, int C, int heads, AttnParams& P) {
  if (!qk_vis || !qk_ir || !vt_vis || !vt_ir || !out_vis || !out_ir) return set_error(ICAF_ERR_BAD_ARG, "cross_attention: null pointer");
  if (B < 1 || N < 1 || n_pad < N || n_pad % 8 || heads < 1 || C % heads) return set_error(ICAF_ERR_BAD_ARG, "cross_attention: bad shape");
  int d = C / heads;
  if (d != 16 && d != 32 && d != 64 && d != 128) return set_error(ICAF_ERR_UNSUPPORTED, "cross_attention: head dim must be 16/32/64/128");
  P.qk[0] = (const __half*)qk_vis; P.qk[1] = (const __half*)qk_ir;
  P.vt[0] = (const __half*)vt_vis; P.vt[1] = (const __half*)vt_ir;
  P.out[0] = (__half*)out_vis; P.out[1] = (__half*)out_ir;
  P.B = B; P.N = N; P.n_pad = n_pad; P.C = C; P.heads = heads;
  P.scale_log2 = 1.4426950408889634f / sqrtf(float(d));   // 1/sqrt(d_k), common.py:670
  return ICAF_OK;
}

template <int D>
static int launch_attn_pipe(const AttnParams& P, cudaStream_t st) {
  using L = AttnSmemP<D>;
  static bool configured[kMaxDevices] = {false};
  if (int rc = configure_smem(cross_attn_pipe_kernel<D>, L::kTotal, configured, "cross_attention: cudaFuncSetAttribute")) return rc;
  dim3 grid((P.n_pad + kQT - 1) / kQT, P.B * P.heads, 2);
  launch_k(cross_attn_pipe_kernel<D>, dim3(grid), dim3(160), L::kTotal, st, P);
  return check_launch("cross_attention");
}

template <int D>
static int launch_attn(const AttnParams& P, cudaStream_t st) {
  using L = AttnSmem<D>;
  static bool configured[kMaxDevices] = {false};
  if (int rc = configure_smem(cross_attn_tc_kernel<D>, L::kTotal, configured, "cross_attention: cudaFuncSetAttribute")) return rc;
  dim3 grid((P.n_pad + kQT - 1) / kQT, P.B * P.heads, 2);
  launch_k(cross_attn_tc_kernel<D>, dim3(grid), dim3(160), L::kTotal, st, P);
  return check_launch("cross_attention");
}

}  // namespace icaf

using namespace icaf;

extern "C" int icaf_cross_attention(const void* qk_vis, const void* qk_ir, const void* vt_vis, const void* vt_ir,
                                    void* out_vis, void* out_ir, int B, int N, int n_pad, int C, int heads, void* stream) {
  AttnParams P;
  int rc = fill_attn(qk_vis, qk_ir, vt_vis, vt_ir, out_vis, out_ir, B, N, n_pad, C, heads, P);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  static const bool pipe_on = []() { const char* e = getenv("ICAF_ATTN_PIPE"); return !(e && e[0] == '0'); }();
  switch (C / heads) {
    // the pipelined kernel pays off where the two MMAs are a visible share of a tile (d = 64: -28 % at 5120 tokens);
    // at d = 16 / 32 the tile is exp-bound (4*d flop per MUFU.EX2) and the serial kernel is 4-8 % faster
    case 16: return launch_attn<16>(P, st);
    case 32: return launch_attn<32>(P, st);
    case 64: return pipe_on ? launch_attn_pipe<64>(P, st) : launch_attn<64>(P, st);
    default: return launch_attn<128>(P, st);
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
