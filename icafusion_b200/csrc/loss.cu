// Detection loss, forward (utils/loss.py:325-463 ComputeLoss.__call__ + build_targets): target assignment, CIoU box loss,
// objectness BCE against IoU-valued targets, class BCE -- three launches, no host round trip, deterministic (no floating
// point atomics: objectness targets meet in an integer atomicMax, every sum is reduced in a fixed order).
//   1. loss_candidates_kernel: one thread per (level, target, anchor, offset) candidate of build_targets (:405-463):
//      anchor-ratio match, the four half-cell neighbour offsets, grid cell, CIoU of the decoded prediction against the
//      target box (general.py:410-447), its (1 - iou) and class-BCE terms, objectness target into tobj by atomicMax
//      (the reference sorts by IoU before its scatter so that the largest IoU wins a contested cell, :374-377).
//   2. loss_obj_kernel: BCEWithLogits(p[..., 4], tobj) with pos_weight, per-block partial sums over fixed chunks.
//   3. loss_finalize_kernel: means, level balance, gains -> (loss * batch, lbox, lobj, lcls, 0).
#include <cmath>
#include <cstring>

#include "icaf_internal.cuh"

namespace icaf {

constexpr int kLossMaxLevels = 5;
constexpr int kLossObjBlocks = 256;     // partial sums per level

struct LossParams {
  const void* p[kLossMaxLevels];
  int ny[kLossMaxLevels], nx[kLossMaxLevels];
  long long cell_off[kLossMaxLevels];   // offset of the level's tobj slab (floats)
  float anchors[kLossMaxLevels * 8 * 2];
  float balance[kLossMaxLevels];
  int p_fp32, nl, B, na, no, nt;
  float box, obj, cls, cls_pw, obj_pw, anchor_t, gr, cp, cn;
  const float* targets;                 // (nt, 6): image, class, x, y, w, h (normalised)
  float* tobj;                          // all levels, (B, na, ny, nx) each
  float* cand_box; float* cand_cls; int* cand_valid;     // [nl][nt][na][5]
  float* obj_part;                      // [nl][kLossObjBlocks]
  float* out;                           // 5 floats
};

__device__ __forceinline__ float loss_ld(const void* p, int fp32, long long i) {
  return fp32 ? reinterpret_cast<const float*>(p)[i] : __half2float(reinterpret_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ float sigmoid_f(float v) { return 1.f / (1.f + expf(-v)); }
// BCEWithLogitsLoss element with pos_weight (torch semantics): (1 - y) x + (1 + (pw - 1) y) (log(1 + exp(-|x|)) + max(-x, 0))
__device__ __forceinline__ float bce_logits(float x, float y, float pw) {
  const float lw = 1.f + (pw - 1.f) * y;
  return (1.f - y) * x + lw * (log1pf(expf(-fabsf(x))) + fmaxf(-x, 0.f));
}

__global__ void loss_candidates_kernel(const LossParams P) {
  pdl_launch_dependents();
  pdl_wait();
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long per_level = (long long)P.nt * P.na * 5;
  if (idx >= per_level * P.nl) return;
  const int lvl = int(idx / per_level);
  long long r = idx - lvl * per_level;
  const int t = int(r / (P.na * 5));
  r -= (long long)t * P.na * 5;
  const int a = int(r / 5), k = int(r - a * 5);
  P.cand_valid[idx] = 0;
  const float* tg = P.targets + (long long)t * 6;
  const int nx = P.nx[lvl], ny = P.ny[lvl];
  const float gx = tg[2] * nx, gy = tg[3] * ny, gw = tg[4] * nx, gh = tg[5] * ny;     // :425-426 targets * gain
  const float aw = P.anchors[(lvl * P.na + a) * 2], ah = P.anchors[(lvl * P.na + a) * 2 + 1];
  const float rw = gw / aw, rh = gh / ah;                                           // :429-430 anchor-multiple match
  if (!(fmaxf(fmaxf(rw, 1.f / rw), fmaxf(rh, 1.f / rh)) < P.anchor_t)) return;
  float ox = 0.f, oy = 0.f;                                                         // :434-441 neighbour cells
  const float g = 0.5f;
  if (k == 1) { if (!(fmodf(gx, 1.f) < g && gx > 1.f)) return; ox = g; }
  else if (k == 2) { if (!(fmodf(gy, 1.f) < g && gy > 1.f)) return; oy = g; }
  else if (k == 3) { const float ix = nx - gx; if (!(fmodf(ix, 1.f) < g && ix > 1.f)) return; ox = -g; }
  else if (k == 4) { const float iy = ny - gy; if (!(fmodf(iy, 1.f) < g && iy > 1.f)) return; oy = -g; }
  const int b = int(tg[0]), c = int(tg[1]);
  int gi = int(gx - ox), gj = int(gy - oy);                                         // .long() truncates
  gi = min(max(gi, 0), nx - 1); gj = min(max(gj, 0), ny - 1);                       // :455 clamp_ (in place: tbox sees it too)
  if (b < 0 || b >= P.B) return;
  const float tx = gx - gi, ty = gy - gj;                                           // :456 target box in cell units
  // prediction at (b, a, gj, gi)                                                     :355-360
  const long long cell = (((long long)b * P.na + a) * ny + gj) * nx + gi;
  const long long pb = cell * P.no;
  const void* pl = P.p[lvl];
  const float sx = sigmoid_f(loss_ld(pl, P.p_fp32, pb)), sy = sigmoid_f(loss_ld(pl, P.p_fp32, pb + 1));
  const float sw = sigmoid_f(loss_ld(pl, P.p_fp32, pb + 2)), sh = sigmoid_f(loss_ld(pl, P.p_fp32, pb + 3));
  const float px = sx * 2.f - 0.5f, py = sy * 2.f - 0.5f;
  const float pw = (sw * 2.f) * (sw * 2.f) * aw, ph = (sh * 2.f) * (sh * 2.f) * ah;
  // CIoU, general.py:418-447 (xywh form, eps = 1e-7)
  const float eps = 1e-7f;
  const float b1x1 = px - pw / 2, b1x2 = px + pw / 2, b1y1 = py - ph / 2, b1y2 = py + ph / 2;
  const float b2x1 = tx - gw / 2, b2x2 = tx + gw / 2, b2y1 = ty - gh / 2, b2y2 = ty + gh / 2;
  const float inter = fmaxf(fminf(b1x2, b2x2) - fmaxf(b1x1, b2x1), 0.f) * fmaxf(fminf(b1y2, b2y2) - fmaxf(b1y1, b2y1), 0.f);
  const float w1 = b1x2 - b1x1, h1 = b1y2 - b1y1 + eps, w2 = b2x2 - b2x1, h2 = b2y2 - b2y1 + eps;
  const float uni = w1 * h1 + w2 * h2 - inter + eps;
  const float iou = inter / uni;
  const float cw = fmaxf(b1x2, b2x2) - fminf(b1x1, b2x1), ch = fmaxf(b1y2, b2y2) - fminf(b1y1, b2y1);
  const float c2 = cw * cw + ch * ch + eps;
  const float dx = b2x1 + b2x2 - b1x1 - b1x2, dy = b2y1 + b2y2 - b1y1 - b1y2;
  const float rho2 = (dx * dx + dy * dy) / 4.f;
  const float da = atanf(w2 / h2) - atanf(w1 / h1);
  const float v = (4.f / (3.14159265358979323846f * 3.14159265358979323846f)) * da * da;
  const float alpha = v / (v - iou + (1.f + eps));
  const float ciou = iou - (rho2 / c2 + v * alpha);
  P.cand_box[idx] = 1.f - ciou;                                                     // :361
  float lc = 0.f;
  if (P.no - 5 > 1) {                                                               // :380-383
    for (int j = 0; j < P.no - 5; ++j)
      lc += bce_logits(loss_ld(pl, P.p_fp32, pb + 5 + j), (j == c) ? P.cp : P.cn, P.cls_pw);
  }
  P.cand_cls[idx] = lc;
  P.cand_valid[idx] = 1;
  const float score = (1.f - P.gr) + P.gr * fmaxf(ciou, 0.f);                       // :364-377 (largest IoU wins a cell)
  atomicMax(reinterpret_cast<int*>(P.tobj + P.cell_off[lvl] + cell), __float_as_int(score));
}

__global__ void __launch_bounds__(256) loss_obj_kernel(const LossParams P) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[256];
  const int lvl = blockIdx.y;
  const long long cells = (long long)P.B * P.na * P.ny[lvl] * P.nx[lvl];
  const long long per = (cells + kLossObjBlocks - 1) / kLossObjBlocks;
  const long long c0 = blockIdx.x * per, c1 = min(c0 + per, cells);
  const float* tobj = P.tobj + P.cell_off[lvl];
  float s = 0.f;
  for (long long c = c0 + threadIdx.x; c < c1; c += 256)
    s += bce_logits(loss_ld(P.p[lvl], P.p_fp32, c * P.no + 4), tobj[c], P.obj_pw);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) P.obj_part[lvl * kLossObjBlocks + blockIdx.x] = red[0];
}

__global__ void __launch_bounds__(256) loss_finalize_kernel(const LossParams P) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float rb[256], rc[256];
  __shared__ int rn[256];
  float lbox = 0.f, lobj = 0.f, lcls = 0.f;
  const long long per_level = (long long)P.nt * P.na * 5;
  for (int lvl = 0; lvl < P.nl; ++lvl) {
    float sb = 0.f, sc = 0.f;
    int n = 0;
    for (long long i = threadIdx.x; i < per_level; i += 256) {
      const long long j = lvl * per_level + i;
      if (P.cand_valid[j]) { sb += P.cand_box[j]; sc += P.cand_cls[j]; ++n; }
    }
    rb[threadIdx.x] = sb; rc[threadIdx.x] = sc; rn[threadIdx.x] = n;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) { rb[threadIdx.x] += rb[threadIdx.x + o]; rc[threadIdx.x] += rc[threadIdx.x + o]; rn[threadIdx.x] += rn[threadIdx.x + o]; }
      __syncthreads();
    }
    float so = 0.f;
    if (threadIdx.x == 0) {
      for (int i = 0; i < kLossObjBlocks; ++i) so += P.obj_part[lvl * kLossObjBlocks + i];
      const long long cells = (long long)P.B * P.na * P.ny[lvl] * P.nx[lvl];
      if (rn[0] > 0) {
        lbox += rb[0] / float(rn[0]);                                              // (1 - iou).mean()
        if (P.no - 5 > 1) lcls += rc[0] / (float(rn[0]) * float(P.no - 5));        // BCEcls mean over n x nc
      }
      lobj += (so / float(cells)) * P.balance[lvl];                                // :389-390
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    lbox *= P.box; lobj *= P.obj; lcls *= P.cls;                                    // :396-398
    P.out[0] = (lbox + lobj + lcls) * float(P.B);                                   // loss * bs
    P.out[1] = lbox; P.out[2] = lobj; P.out[3] = lcls; P.out[4] = 0.f;              // lrk (ranking loss) is disabled upstream (:386)
  }
}

static size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace icaf

using namespace icaf;

extern "C" size_t icaf_loss_workspace_bytes(int B, int na, int nt, const int* ny, const int* nx, int nl) {
  if (B < 1 || na < 1 || nt < 0 || nl < 1 || nl > kLossMaxLevels || !ny || !nx) return 0;
  size_t cells = 0;
  for (int i = 0; i < nl; ++i) cells += (size_t)B * na * ny[i] * nx[i];
  const size_t cand = (size_t)nl * (nt > 0 ? nt : 1) * na * 5;
  return align256(cells * 4) + 3 * align256(cand * 4) + align256((size_t)nl * kLossObjBlocks * 4);
}

extern "C" int icaf_compute_loss_fwd(const void* const* p, int p_fp32, const int* ny, const int* nx, int nl, int B, int na, int no,
                                     const float* targets, int nt, const float* anchors_host, const icaf_loss_hyp* hyp, float* out,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  if (!p || !ny || !nx || !anchors_host || !hyp || !out || !workspace) return set_error(ICAF_ERR_BAD_ARG, "compute_loss: null pointer");
  if (nl < 1 || nl > kLossMaxLevels || B < 1 || na < 1 || na > 8 || no < 6 || nt < 0 || (nt > 0 && !targets))
    return set_error(ICAF_ERR_BAD_ARG, "compute_loss: bad shape (nl <= 5, na <= 8)");
  if (hyp->fl_gamma > 0.f) return set_error(ICAF_ERR_UNSUPPORTED, "compute_loss: focal loss (fl_gamma > 0) is not built");
  if (workspace_bytes < icaf_loss_workspace_bytes(B, na, nt, ny, nx, nl) || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return set_error(ICAF_ERR_BAD_ARG, "compute_loss: workspace too small (icaf_loss_workspace_bytes) or not 256-byte aligned");
  LossParams P;
  memset(&P, 0, sizeof(P));
  size_t cells = 0;
  for (int i = 0; i < nl; ++i) {
    if (!p[i] || ny[i] < 1 || nx[i] < 1) return set_error(ICAF_ERR_BAD_ARG, "compute_loss: bad level");
    P.p[i] = p[i]; P.ny[i] = ny[i]; P.nx[i] = nx[i]; P.cell_off[i] = (long long)cells;
    cells += (size_t)B * na * ny[i] * nx[i];
    P.balance[i] = hyp->balance[i];
  }
  for (int i = 0; i < nl * na * 2; ++i) P.anchors[i] = anchors_host[i];
  P.p_fp32 = p_fp32; P.nl = nl; P.B = B; P.na = na; P.no = no; P.nt = nt;
  P.box = hyp->box; P.obj = hyp->obj; P.cls = hyp->cls; P.cls_pw = hyp->cls_pw; P.obj_pw = hyp->obj_pw;
  P.anchor_t = hyp->anchor_t; P.gr = hyp->gr; P.cp = hyp->cp; P.cn = hyp->cn;
  P.targets = targets; P.out = out;
  const size_t cand = (size_t)nl * (nt > 0 ? nt : 1) * na * 5;
  char* w = (char*)workspace;
  P.tobj = (float*)w; w += align256(cells * 4);
  P.cand_box = (float*)w; w += align256(cand * 4);
  P.cand_cls = (float*)w; w += align256(cand * 4);
  P.cand_valid = (int*)w; w += align256(cand * 4);
  P.obj_part = (float*)w;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(P.tobj, 0, cells * 4, st);                        // tobj = zeros_like(pi[..., 0])   :351
  if (e == cudaSuccess) e = cudaMemsetAsync(P.cand_valid, 0, cand * 4, st);
  if (e != cudaSuccess) return set_cuda_error(e, "compute_loss: cudaMemsetAsync");
  if (nt > 0) {
    const long long total = (long long)nl * nt * na * 5;
    launch_k(loss_candidates_kernel, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, st, P);
    if (int rc = check_launch("compute_loss(candidates)")) return rc;
  }
  launch_k(loss_obj_kernel, dim3(kLossObjBlocks, nl), dim3(256), 0, st, P);
  if (int rc = check_launch("compute_loss(objectness)")) return rc;
  launch_k(loss_finalize_kernel, dim3(1), dim3(256), 0, st, P);
  return check_launch("compute_loss(finalize)");
}
