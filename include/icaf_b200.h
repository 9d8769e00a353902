/*
 * icaf_b200 -- C ABI of the B200 (sm_100a) kernels for the ICAFusion hot path:
 * the two-stream CSPDarknet Conv+BN+SiLU backbone and the DMFF cross-attention fusion block.
 *
 * The reference (chanchanchan97/ICAFusion) is pure Python/PyTorch and has no FFI of its own; these
 * entry points are what a binding for its operator library (models/common.py) calls instead of the
 * PyTorch ops cited at each function.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - Plain C types only.  Every pointer is a DEVICE pointer unless stated otherwise.
 *   - Activations are fp16 NHWC "views": a base pointer plus a pixel pitch `ld` in elements, so a
 *     kernel can read or write a channel slice of a wider (concatenated) buffer in place.
 *     (A torch tensor of logical shape (B,C,H,W) in channels_last memory format is exactly this.)
 *   - Token tensors are fp16 (B, Npad, C) with Npad = N rounded up to 8.
 *   - The caller owns all memory (inputs, outputs, workspaces); the library never allocates device
 *     memory, never synchronises, and enqueues all work on the `stream` argument (a cudaStream_t),
 *     so it composes with the PyTorch caching allocator, autograd hooks and CUDA-graph capture.
 *   - Return value: ICAF_OK or an error code; icaf_last_error() gives a thread-local message.
 */
#ifndef ICAF_B200_H
#define ICAF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICAF_OK 0
#define ICAF_ERR_BAD_ARG 1
#define ICAF_ERR_UNSUPPORTED 2
#define ICAF_ERR_CUDA 3

#define ICAF_ACT_NONE 0
#define ICAF_ACT_SILU 1 /* nn.SiLU, models/common.py:54 */
#define ICAF_ACT_GELU 2 /* nn.GELU (erf), models/common.py:706 */

#define ICAF_EPI_BIAS_ROW 1   /* bias indexed by output row (swap-AB linears) instead of channel   */
#define ICAF_EPI_ADD_RES 2    /* y = act(acc+bias) + res          (Bottleneck shortcut, common.py:194) */
#define ICAF_EPI_SCALED_RES 4 /* y = alpha*res + beta*(acc+bias)  (LearnableCoefficient pairs, common.py:747-750) */
/* LayerNorm folded into the linear layer that consumes it (common.py:660-668, 749-750 + 704-706): the caller packs
 * W' = W diag(gamma), bias' = bias + W beta, ln_colsum[n] = sum_k W'[n][k]; the kernel runs the GEMM on the RAW rows and
 * normalises in the epilogue, y = act( rstd_m * (acc - mean_m * ln_colsum[n]) + bias'[n] ), with (mean, rstd) of input row m
 * from `ln_stats`: ln_parts (sum, sum of squares) fp32 pairs per row, as an EMIT_STATS producer or icaf_row_stats wrote
 * them.  1x1 geometry only, no residual, activation none or GELU. */
#define ICAF_EPI_LN_FOLD 8
/* With SCALED_RES: also write, per output row, (sum, sum of squares) partials of the fp16-rounded outputs to `stats_out`:
 * float2 [M][ceil(Cout/32)] (partials of one row sum to the row's totals) -- the ln_stats of the next LN_FOLD layer. */
#define ICAF_EPI_EMIT_STATS 16

int icaf_version(void);
const char* icaf_last_error(void);
/* Kernels this library has enqueued so far in this process (monotonic; one C-ABI compute call may launch more than one
 * kernel, e.g. a convolution whose last wave is peeled off as a split-K tail).  bench.py reports its per-step delta. */
long long icaf_kernel_launches(void);
/* Number of SMs of the current device (grid sizing of callers' workspaces); <0 on error. */
int icaf_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution / linear layer on the tcgen05 tensor cores.
 *   y[b,oy,ox,n] = epi( sum_{ky,kx,c} x[b, oy*s-p+ky, ox*s-p+kx, c] * w[n][(ky*kw+kx)*Cin + c] + bias[n] )
 * Replaces: Conv.forward / Conv.fuseforward (models/common.py:56-60: nn.Conv2d -> BatchNorm2d -> SiLU,
 * BN folded as utils/torch_utils.py:182-202), nn.Linear calls of CrossAttention / MLP
 * (common.py:660-668,683-685,704-715; a linear is the 1x1 case with Hi=Wi=1 rows as pixels) and
 * Detect's nn.Conv2d (models/yolo_test.py:49).
 * `n_io` problems of identical geometry (e.g. the RGB and the IR stream) run in one launch.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int B, Hi, Wi, Cin; /* input  NHWC; Cin multiple of 8, or exactly 4 (packed 3-channel image, even Wi) */
  int Ho, Wo, Cout;   /* output NHWC */
  int kh, kw, stride, pad;
  int k_pad;  /* row pitch of the packed filter matrix: multiple of 64, >= kh*kw*Cin (zero padded) */
  int w_rows; /* rows present in the packed filter matrix (>= Cout, zero padded) */
  int act;    /* ICAF_ACT_* */
  int epi;    /* ICAF_EPI_* flags */
} icaf_conv_geom;

typedef struct {
  const void* x;     int64_t x_ld;   /* fp16 input view                                   */
  const void* w;                     /* fp16 [w_rows][k_pad], K order (ky,kx,c)           */
  const float* bias;                 /* fp32 [Cout] (or [rows] with BIAS_ROW); may be NULL */
  const void* res;   int64_t res_ld; /* fp16 residual view (ADD_RES / SCALED_RES) or NULL  */
  void* y;           int64_t y_ld;   /* fp16 output view                                   */
  const float* alpha;                /* device scalars for SCALED_RES                      */
  const float* beta;
  const float* ln_stats;             /* LN_FOLD: fp32 pairs [M][ln_parts] (sum, sum of squares) of the input rows */
  const float* ln_colsum;            /* LN_FOLD: fp32 [w_rows]                              */
  int ln_parts;                      /* LN_FOLD: partials per row, 1..64                     */
  float ln_eps;                      /* LN_FOLD: LayerNorm epsilon                           */
  float* stats_out;                  /* EMIT_STATS: fp32 pairs [M][ceil(Cout/32)]            */
} icaf_conv_io;

int icaf_conv2d_fwd(const icaf_conv_geom* g, const icaf_conv_io* io, int n_io, void* stream);

/* The dispatcher's decision for one layer geometry, computed on the HOST only (no device, no stream, no pointers):
 * which kernel family runs it, with which tile shapes, grid and shared memory.  icaf_conv2d_fwd launches exactly the
 * plan this returns for (geometry, n_io, SM count of the current device, ICAF_PAIR switch); the same invariant checks
 * run in both, so a GPU-less test can walk every layer of a model through the dispatcher (tests/test_abi_cpu.py).
 * pair_mode: -1 = environment default (ICAF_PAIR), 0 = never CTA pairs, 1 = heuristic, 2 = pairs wherever possible. */
#define ICAF_KERNEL_TC 0      /* one 128 x BN tile per CTA, split-K clusters   (conv_gemm.cu)    */
#define ICAF_KERNEL_PERSIST 1 /* one CTA per SM looping over tiles              (conv_persist.cu) */
#define ICAF_KERNEL_PAIR 2    /* CTA pairs, tcgen05 cta_group::2, halo copies   (conv_pair.cu)    */
#define ICAF_KERNEL_STEM 3    /* image stem over the space-to-depth frame, x-merged rows (conv_stem.cu); the plan
                                 assumes a dense frame (pixel pitch 16), which icaf_conv2d_fwd checks on the pointers */
typedef struct {
  int kernel;                            /* ICAF_KERNEL_*                                                     */
  int bn;                                /* output-channel tile width (32/64/128/256)                         */
  int a_mode;                            /* activation staging: 0 cp.async gather, 1 2-D TMA, 2 4-D TMA       */
  int tile_w, tile_h, tiles_x, tiles_y;  /* 4-D TMA: output-pixel tile and tiles per image                    */
  int cblk;                              /* channels per TMA box                                              */
  int halo;                              /* pair kernel: 0 tap boxes, 1 x-shifted halo copies, 2 + resident filter */
  int stages, splits;                    /* smem ring depth; split-K factor (= cluster size of the tc kernel) */
  int grid_x, grid_y, grid_z, cluster;   /* launch shape                                                      */
  int smem_bytes;                        /* dynamic shared memory per CTA                                     */
  int work_items;                        /* tiles (tc/persist) or tile pairs (pair) the grid iterates over    */
} icaf_conv_plan;
int icaf_conv2d_plan(const icaf_conv_geom* g, int n_io, int sm_count, int pair_mode, icaf_conv_plan* out);

/* Test-only CUDA-core reference of the same contract (slow, obviously-correct); used by tests to
 * localise tensor-core bugs on the device.  Not called by the product path. */
int icaf_conv2d_fwd_simt(const icaf_conv_geom* g, const icaf_conv_io* io, int n_io, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Input staging: (B,3,H,W) planar image -> fp16 NHWC with C padded to 4 (r,g,b,0).
 * Replaces the `.half()` / `/255` staging of detect_twostream.py:70-80 / train.py:295-297.
 * src_dtype: 0 = fp16, 1 = fp32, 2 = uint8 (scaled by `scale`, e.g. 1/255).
 * ------------------------------------------------------------------------------------------- */
int icaf_pack_image(const void* src, int src_dtype, float scale, int B, int H, int W, void* dst, void* stream);

/* Letterbox + BGR->RGB + HWC->CHW for a batch of decoded frames, on the device.  Replaces utils/datasets.py:1404-1427
 * letterbox (cv2.resize INTER_LINEAR + cv2.copyMakeBorder) and the `img[:, :, ::-1].transpose(2, 0, 1)` of datasets.py:238.
 * src: uint8 (B, H0, W0, 3) BGR; dst: uint8 (B, 3, H, W) RGB planar; the resized (new_h, new_w) frame sits at (top, left),
 * the rest is pad_value (114).  xtab / ytab: device int32 [new_w][4] = {x0, x1, a0, a1} and [new_h][4] = {y0, y1, b0, b1},
 * cv2's fixed-point bilinear taps (x 2048) -- built on the host (icafusion_b200/datasets.py:resize_taps); may be NULL when
 * (new_h, new_w) == (H0, W0).  Bit-exact against the cv2 pipeline. */
int icaf_letterbox(const void* src, int B, int H0, int W0, void* dst, int H, int W, int top, int left, int new_h, int new_w,
                   const int* xtab, const int* ytab, int pad_value, void* stream);

/* Same staging, space-to-depth layout: dst is (B, H/2, W/2, 16) fp16 with channel (dy*2+dx)*4 + c (c = r,g,b,0).
 * A 6x6 / stride 2 / pad 2 stem convolution over the image (yolov5 "P1/2" row of the model YAML) is then exactly a
 * 3x3 / stride 1 / pad 1 convolution over this tensor (ky = 2*ty+dy, kx = 2*tx+dx), which runs on the TMA path. H, W even. */
int icaf_pack_image_s2d(const void* src, int src_dtype, float scale, int B, int H, int W, void* dst, void* stream);

/* SPPF's three chained MaxPool2d(5,1,2) (models/common.py:259-266): y1,y2,y3 written as channel
 * slices; x is (B,H,W,C) view. */
int icaf_sppf_pool(const void* x, int64_t x_ld, void* y1, void* y2, void* y3, int64_t y_ld, int B, int H, int W,
                   int C, void* stream);

/* nn.Upsample(None, 2, 'nearest') (yolov5l_Transfusion_kaist.yaml:48,53) into a channel slice. */
int icaf_upsample2x(const void* x, int64_t x_ld, void* y, int64_t y_ld, int B, int H, int W, int C, void* stream);

/* Pull a read-only region (packed filters) into L2 ahead of its consumers: after an L2 flush every layer would
 * otherwise pay a DRAM round trip for its first filter tile.  Touches no data; purely a cache hint. */
int icaf_prefetch_l2(const void* ptr, size_t bytes, void* stream);

/* Copy a channel slice (Concat, models/common.py:313-321, when producer-side slice writes are not possible). */
int icaf_copy_channels(const void* x, int64_t x_ld, void* y, int64_t y_ld, int64_t pixels, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DMFF block pieces (models/common.py:762-891).
 * ------------------------------------------------------------------------------------------- */
/* AdaptivePool2d avg+max (common.py:868-891) + LearnableWeights (:579-587) + flatten/permute + pos_emb (:817-823):
 *   tok[b, n, c] = w[0]*avgpool + w[1]*maxpool + pos[n, c];  rows n in [N, Npad) are zeroed.
 * Two modalities per launch (x0/x1 ...). `mix` = 4 device floats {w1_vis, w2_vis, w1_ir, w2_ir}.
 * stats_* (both or neither; C % 32 == 0): fp32 pairs [B*Npad][C/32] (sum, sum of squares) of every token row per 32
 * channels -- the ln_stats (ln_parts = C/32) of the LN_FOLD projection that consumes the tokens. */
int icaf_dmff_pool_tokens(const void* x_vis, const void* x_ir, int64_t x_ld, const void* pos_vis, const void* pos_ir,
                          const float* mix, void* tok_vis, void* tok_ir, float* stats_vis, float* stats_ir, int B, int H,
                          int W, int C, int nh, int nw, int n_pad, void* stream);

/* nn.LayerNorm over the last dim (eps 1e-5) of `rows` x C fp16 tokens; two independent problems per launch
 * (x1 may be NULL).  gamma/beta are fp32 [C].  Replaces common.py:660,665,749-750. */
int icaf_layernorm(const void* x0, const void* x1, const float* g0, const float* b0, const float* g1,
                   const float* b1, void* y0, void* y1, int64_t rows, int C, float eps, void* stream);

/* (sum, sum of squares) of every row of a (rows, C) fp16 matrix -> fp32 pairs [rows][1]: the `ln_stats` (ln_parts = 1) of
 * an LN_FOLD layer whose input no EMIT_STATS epilogue produced (first loop of a DMFF block, stand-alone calls).
 * Two problems per launch (x1 may be NULL). */
int icaf_row_stats(const void* x0, const void* x1, float* stats0, float* stats1, int64_t rows, int C, void* stream);

/* Bidirectional cross-attention core (common.py:670-684), flash style: no N x N score matrix in HBM.
 *   out_vis = softmax(q_ir k_vis^T / sqrt(d)) v_vis ;  out_ir = softmax(q_vis k_ir^T / sqrt(d)) v_ir
 * Two input forms:
 *   fused  (vt_vis == vt_ir == NULL): qk_* are fp16 (B, Npad, 3C) rows [q | k | v] exactly as ONE fused projection GEMM
 *          emits them; the V tiles feed the tensor core as MN-major operands, no transpose anywhere;
 *   split  : qk_* fp16 (B, Npad, 2C) rows [q | k], vt_* fp16 (C, B*Npad) value projection stored transposed.
 * out_*: fp16 (B, Npad, C) heads merged (the layout out_proj consumes).  Every tile is staged by TMA. */
int icaf_cross_attention(const void* qk_vis, const void* qk_ir, const void* vt_vis, const void* vt_ir, void* out_vis,
                         void* out_ir, int B, int N, int n_pad, int C, int heads, void* stream);

/* Test-only CUDA-core reference of icaf_cross_attention (same contract). */
int icaf_cross_attention_simt(const void* qk_vis, const void* qk_ir, const void* vt_vis, const void* vt_ir,
                              void* out_vis, void* out_ir, int B, int N, int n_pad, int C, int heads, void* stream);

/* Token map -> feature map tail (common.py:827-840): reshape (B,nh,nw,C), F.interpolate to (H,W)
 * (mode 0 = bilinear align_corners=False [eval], 1 = nearest [train]), add the stream's own features,
 * write both modalities into one (B,H,W,2C) buffer = the Concat that feeds conv1x1_out. */
int icaf_dmff_upsample_cat(const void* tok_vis, const void* tok_ir, int n_pad, const void* x_vis, const void* x_ir,
                           int64_t x_ld, void* y, int64_t y_ld, int B, int H, int W, int C, int nh, int nw, int mode,
                           void* stream);

/* Detect head decode (models/yolo_test.py:49-65) for one level.  `p` is the 1x1-conv output
 * (B,ny,nx,p_ld) fp16 with na*no valid channels.  Writes
 *   x_out  (B,na,ny,nx,no) fp16  raw,   z (B, total_rows, no) rows [row_off, row_off+na*ny*nx) decoded,
 *   logits (B, total_rows, no-5) same rows.  anchors: host array na*2 floats (pixels). */
int icaf_detect_decode(const void* p, int64_t p_ld, void* x_out, void* z, void* logits, int B, int ny, int nx, int na,
                       int no, int total_rows, int row_off, float stride, const float* anchors_host, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batched non-maximum suppression on the decoded predictions, fully on the device.
 * Replaces utils/general.py:518-607 non_max_suppression (best-class branch: conf = obj * max cls, both > conf_thres,
 * class-offset boxes unless `agnostic`, torchvision.ops.nms greedy suppression at iou_thres, max_nms = 30000 candidates,
 * first max_det kept).  z: fp16 (B, R, no) as icaf_detect_decode writes it; arithmetic in fp32 like the reference on
 * z.float().  class_mask: bit k set = keep class k (0 = all classes; at most 64 classes with a mask).
 * det: fp32 (B, max_det, 6) rows [x1, y1, x2, y2, conf, cls] in confidence order; count: int32 (B) rows valid per image.
 * workspace: icaf_nms_workspace_bytes(B, R) bytes of device memory, 8-byte aligned (caller owned).
 * ------------------------------------------------------------------------------------------- */
size_t icaf_nms_workspace_bytes(int B, int R);
int icaf_nms(const void* z, int B, int R, int no, float conf_thres, float iou_thres, int agnostic, uint64_t class_mask,
             int max_det, float* det, int* count, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Detection loss, forward only (the validation loss test.py:132-133 accumulates; the training backward is not built):
 * utils/loss.py:325-463 ComputeLoss.__call__ + build_targets -- anchor-ratio matching with the four half-cell neighbour
 * offsets, CIoU box loss, objectness BCE against IoU-valued targets (largest IoU wins a contested cell), class BCE.
 * p: nl device pointers to the Detect training outputs (B, na, ny[i], nx[i], no), fp16 (p_fp32 = 0) or fp32 (1); arithmetic
 * in fp32.  targets: device fp32 (nt, 6) rows [image, class, x, y, w, h] normalised to [0, 1].  anchors_host: nl*na*2 floats
 * in grid units (Detect.anchors).  out: 5 device floats [loss * batch, lbox, lobj, lcls, 0].  fl_gamma must be 0.
 * Deterministic: no floating-point atomics.  workspace: icaf_loss_workspace_bytes(...) bytes, 256-byte aligned.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  float box, obj, cls;   /* loss gains hyp['box'], hyp['obj'], hyp['cls'] (already scaled as train.py:226-228 does) */
  float cls_pw, obj_pw;  /* BCE positive weights                                                                   */
  float anchor_t;        /* anchor-multiple threshold                                                              */
  float fl_gamma;        /* focal-loss gamma; only 0 is supported                                                  */
  float gr;              /* model.gr: objectness target = (1 - gr) + gr * iou                                      */
  float cp, cn;          /* smoothed positive / negative class targets (smooth_BCE, loss.py:15-17)                 */
  float balance[5];      /* per-level objectness weights: {4, 1, 0.4} for three levels (loss.py:346)               */
} icaf_loss_hyp;
/* p_ld: 0 when p[i] is the reference's (B, na, ny, nx, no) contiguous tensor; otherwise the pixel pitch (elements) of the
 * head's own (B, ny, nx, na*no) NHWC map (the training path hands the 1x1 head convolution's output over without a permute).
 * no_bwd: 0 for a forward-only workspace, `no` when icaf_compute_loss_bwd will follow. */
size_t icaf_loss_workspace_bytes(int B, int na, int nt, const int* ny, const int* nx, int nl, int no_bwd);
int icaf_compute_loss_fwd(const void* const* p, int p_fp32, int p_ld, const int* ny, const int* nx, int nl, int B, int na, int no,
                          const float* targets, int nt, const float* anchors_host, const icaf_loss_hyp* hyp, float* out,
                          void* workspace, size_t workspace_bytes, void* stream);
/* Backward of the loss (train.py:344 starts here): dp[i] = grad_out[0] * d out[0] / d p[i], in the dtype and memory layout of
 * p[i] (every element of the (cells x no) slab is written; pad channels of an NHWC map with p_ld > na*no are left alone).
 * Same arguments as the forward call that filled `workspace` (sized with no_bwd = no); grad_out: device fp32 scalar (the
 * GradScaler factor arrives here).  The CIoU derivative is the forward expression evaluated on forward-mode duals; the
 * objectness target and CIoU's alpha are constants, as in the reference (loss.py:372, general.py:444).  Candidate gradients
 * meeting in one cell are summed with fp32 atomics (sum order not fixed), then rounded once. */
int icaf_compute_loss_bwd(const void* const* p, int p_fp32, int p_ld, const int* ny, const int* nx, int nl, int B, int na, int no,
                          const float* targets, int nt, const float* anchors_host, const icaf_loss_hyp* hyp, const float* grad_out,
                          void* const* dp, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training-step building blocks (train.py:344 backward of the hot path's nn.Conv2d / nn.Linear layers).  Operator level
 * only: the whole-model backward / optimiser / DDP loop is not built.
 * ------------------------------------------------------------------------------------------- */
/* Weight gradient dW[n][c][ky][kx] (fp32, PyTorch layout) = (accumulate ? dW : 0) + scale * sum_pixels dy[.., n] * x[.. shifted .., c]
 * of the convolution / linear layer described by `g` (Cin 16, 32 or a multiple of 64; Cout % 8 == 0; stride 1 or 2).
 * x, dy: fp16 NHWC views (pixel pitch x_ld / dy_ld).  tcgen05 GEMM over pixels with both operands MN-major, split over
 * the pixel range, splits summed in a fixed order (deterministic).  scale: 1 / loss scale.  workspace: caller owned,
 * icaf_conv2d_wgrad_workspace_bytes(g) bytes, 16-byte aligned. */
size_t icaf_conv2d_wgrad_workspace_bytes(const icaf_conv_geom* g);
int icaf_conv2d_wgrad(const icaf_conv_geom* g, const void* x, int64_t x_ld, const void* dy, int64_t dy_ld, float* dw, float scale,
                      int accumulate, void* workspace, size_t workspace_bytes, void* stream);
/* y (B, H2, W2, C) = x (B, H, W, C) with a zero between every two pixels (y[b, 2i, 2j] = x[b, i, j]): the data gradient of a
 * stride-2 convolution is the stride-1 icaf_conv2d_fwd of this tensor with the flipped, transposed filter. */
int icaf_zero_stuff2(const void* x, void* y, int B, int H, int W, int C, int H2, int W2, void* stream);
/* out[c] (fp32) = (accumulate ? out[c] : 0) + scale * sum_rows x[r][c] of a dense fp16 (rows, C) matrix: bias gradients.
 * Deterministic two-stage sum; workspace: 64 * C floats. */
int icaf_colsum(const void* x, int64_t rows, int C, float* out, float scale, int accumulate, float* workspace, size_t workspace_bytes,
                void* stream);

/* Reduction scratch of the kernels below: icaf_train_workspace_bytes(C) bytes (fp32, 16-byte aligned), plus what each states. */
size_t icaf_train_workspace_bytes(int C);
/* BatchNorm2d with BATCH statistics + activation (Conv.forward in training mode, models/common.py:56-57; act: 0 none, 1 SiLU):
 * x, y: dense fp16 (rows, C) = NHWC maps; statistics in fp32 over the rows; run_mean / run_var (may be NULL) get the momentum
 * update with the unbiased variance like nn.BatchNorm2d; save_mean / save_invstd (fp32 [C]) feed the backward. */
int icaf_bn_act_fwd(const void* x, const float* gamma, const float* beta, float* run_mean, float* run_var, void* y, float* save_mean,
                    float* save_invstd, int64_t rows, int C, float eps, float momentum, int act, float* workspace, size_t workspace_bytes,
                    void* stream);
/* Its backward: dx (fp16) and dgamma / dbeta (fp32, (accumulate ? += : =) grad_scale * value; may be NULL).
 * workspace: icaf_train_workspace_bytes(C) (it includes the per-channel coefficient block of the apply pass). */
int icaf_bn_act_bwd(const void* x, const void* dy, const float* gamma, const float* beta, const float* save_mean, const float* save_invstd,
                    void* dx, float* dgamma, float* dbeta, int64_t rows, int C, int act, float grad_scale, int accumulate, float* workspace,
                    size_t workspace_bytes, void* stream);
/* Element-wise over n fp16 values (n % 8 == 0): mode 0 y = GELU_erf(x) (common.py:706); 1 y = dy * GELU'(x);
 * 2 y = dropout(x, p) with a counter-based mask keyed by (element index, seed) -- calling it on dy with the same seed is the backward. */
int icaf_eltwise(int mode, const void* x, const void* dy, void* y, int64_t n, float p, uint32_t seed, void* stream);
/* nn.LayerNorm backward over dense fp16 (rows, C), C <= 2048: dx, dgamma, dbeta as above.
 * workspace: icaf_train_workspace_bytes(C) + 2 rows floats. */
int icaf_layernorm_bwd(const void* x, const void* dy, const float* gamma, void* dx, float* dgamma, float* dbeta, int64_t rows, int C, float eps,
                       float grad_scale, int accumulate, float* workspace, size_t workspace_bytes, void* stream);
/* out[0] = (accumulate ? out[0] : 0) + scale * <x, y> over dense fp16 (rows, C): gradients of LearnableCoefficient / LearnableWeights. */
int icaf_dot(const void* x, const void* y, int64_t rows, int C, float* out, float scale, int accumulate, float* workspace, size_t workspace_bytes,
             void* stream);
/* Backward of nn.Upsample(None, 2, 'nearest'): dx (B, H, W, C) = sums of the 2 x 2 blocks of dy (B, 2H, 2W, C). */
int icaf_upsample2x_bwd(const void* dy, void* dx, int B, int H, int W, int C, void* stream);
/* Backward of one MaxPool2d(5, 1, 2) of SPPF's chain (common.py:259-266): x is that pool's input, dy the gradient of its output
 * (dense fp16 NHWC, C % 8 == 0).  workspace: B*H*W*C bytes (one arg-max code per window and channel), 8-byte aligned. */
int icaf_maxpool5_bwd(const void* x, const void* dy, void* dx, int B, int H, int W, int C, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of icaf_dmff_pool_tokens w.r.t. the two feature maps (dense fp16 (B,H,W,C) gradients): every pixel gathers, from
 * each pooling window that contains it, dtok * (w_avg / window + w_max * [pixel is the window's first maximum]).  The
 * gradients of the mixing weights and positional embeddings are plain reductions of dtok (icaf_dot / icaf_colsum). */
int icaf_dmff_pool_tokens_bwd(const void* x_vis, const void* x_ir, int64_t x_ld, const void* dtok_vis, const void* dtok_ir, const float* mix,
                              void* dx_vis, void* dx_ir, int B, int H, int W, int C, int nh, int nw, int n_pad, void* workspace,
                              size_t workspace_bytes, void* stream);   /* workspace: 2*B*nh*nw*C bytes (arg-max codes), 8-byte aligned */
/* Backward of icaf_dmff_upsample_cat (mode 1, nearest: the training-mode tail, common.py:828-829) w.r.t. the token streams:
 * dtok[b][n] = sum of dcat over the pixels token n was copied to (pad rows get 0).  dcat: (B,H,W,2C) with pixel pitch d_ld;
 * the gradients of the two residual inputs are its channel halves. */
int icaf_dmff_upsample_cat_bwd(const void* dcat, int64_t d_ld, void* dtok_vis, void* dtok_ir, int B, int H, int W, int C, int nh, int nw,
                               int n_pad, int mode, void* stream);
/* fp32 master filter (Cout,Cin,kh,kw) -> fp16 bank [rows][k_pad] of icaf_conv2d_fwd, K order (ky,kx,channel), channel count
 * padded to chan_pad, padding written as zero.  transpose_flip = 0: the forward filter (rows >= Cout, channels = Cin);
 * 1: the data-gradient filter W'[c][n][ky][kx] = W[n][c][kh-1-ky][kw-1-kx] (rows >= Cin, channels = Cout). */
int icaf_pack_weight(const float* w, int Cout, int Cin, int kh, int kw, int chan_pad, int rows, int k_pad, int transpose_flip, void* out,
                     void* stream);
/* Both banks of one filter in one launch: out_fwd [rows_f][kpad_f] (channels = Cin) and out_dgrad [rows_d][kpad_d] (channels = Cout
 * padded to chan_pad_d). */
int icaf_pack_weight_pair(const float* w, int Cout, int Cin, int kh, int kw, int rows_f, int kpad_f, void* out_fwd, int chan_pad_d, int rows_d,
                          int kpad_d, void* out_dgrad, void* stream);

/* Optional device-side counter (one uint32) added to every dropout seed by the kernels at run time (NULL switches it off).  A
 * captured CUDA graph of the training step keeps its host-side seeds; bumping this counter on the device between replays gives
 * every replay fresh dropout masks, and the forward / backward kernels of one step still regenerate identical masks. */
int icaf_set_seed_offset(const void* device_u32);

/* Training-mode forward of the fused form: like icaf_cross_attention(qkv_vis, qkv_ir, NULL, NULL, ...) plus dropout with
 * probability p_drop on the attention probabilities (common.py:677,680; counter-based mask keyed by `seed`, reproduced by the
 * backward below). */
int icaf_cross_attention_train(const void* qkv_vis, const void* qkv_ir, void* out_vis, void* out_ir, int B, int N, int n_pad, int C, int heads,
                               float p_drop, uint32_t seed, void* stream);

/* Backward of icaf_cross_attention in its fused form (qkv_* fp16 (B, Npad, 3C) rows [q | k | v]; out_* the forward outputs,
 * dout_* their gradients, fp16 (B, Npad, C)): dqkv_* (same layout as qkv_*; pad rows zeroed).  Probabilities are recomputed;
 * p_drop / seed reproduce the dropout mask of a training forward (0 = none).  CUDA-core kernels for the pooled-token regime.
 * workspace: icaf_cross_attention_bwd_workspace_bytes(B, n_pad, heads). */
size_t icaf_cross_attention_bwd_workspace_bytes(int B, int n_pad, int heads);
int icaf_cross_attention_bwd(const void* qkv_vis, const void* qkv_ir, const void* out_vis, const void* out_ir, const void* dout_vis,
                             const void* dout_ir, void* dqkv_vis, void* dqkv_ir, int B, int N, int n_pad, int C, int heads, float p_drop,
                             uint32_t seed, void* workspace, size_t workspace_bytes, void* stream);

/* out = a[0] * x (+ b[0] * y when y != NULL) over n fp16 elements (n % 8 == 0, 16-byte aligned); a, b device fp32 scalars.
 * LearnableCoefficient.forward / LearnableWeights.forward called stand-alone (models/common.py:569-587). */
int icaf_axpby(const void* x, const void* y, const float* a, const float* b, void* out, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ICAF_B200_H */
