"""CPU oracle for the ICAFusion hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain fp32 restatement (torch CPU functional ops on a flat ``state_dict``) of the
reference's two-stream CSPDarknet + DMFF forward.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this file; nothing under ``icafusion_b200/`` does, and the product path raises
when its CUDA library is missing rather than falling back to this.

Pinning: the reference has no tests / golden vectors of its own, so
this oracle is pinned against *outputs of the reference itself*, executed in the build
container by ``oracle/gen_golden.py`` (which imports /root/reference through
``oracle/ref_shim.py``) and committed as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays them.

Every function cites the reference lines it restates (paths relative to /root/reference).
The restatement is deliberately *functional* (no nn.Module, no parameters objects): the
reference's algorithm is "call these torch ops in this order with these tensors".
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

BN_EPS_MODEL = 1e-3      # utils/torch_utils.py:151  (initialize_weights sets eps on every BN of a built Model)
BN_MOMENTUM_MODEL = 0.03  # utils/torch_utils.py:152
_BN_TRAIN = [False]       # set by model_forward(training=True): Conv.forward with the module in train() (common.py:56-57)
BN_EPS_STANDALONE = 1e-5  # nn.BatchNorm2d default, stand-alone TransformerFusionBlock
LN_EPS = 1e-5            # nn.LayerNorm default, models/common.py:631-632,723-724


# ----------------------------------------------------------------------------------------
# Conv / Bottleneck / C3 / SPPF            models/common.py:36-60, 184-194, 216-227, 252-267
# ----------------------------------------------------------------------------------------
def autopad(k: int, p=None) -> int:
    """models/common.py:36-40"""
    return k // 2 if p is None else p


def conv_bn_silu(x, sd: SD, pre: str, k: int, s: int, p=None, act=True, bn_eps=BN_EPS_MODEL):
    """`Conv.forward` (common.py:56-57) when `pre.bn.*` exists, else `Conv.fuseforward`
    (common.py:59-60) on the BN-folded conv produced by fuse_conv_and_bn (torch_utils.py:182-202)."""
    w = sd[pre + ".conv.weight"]
    pad = autopad(k, p)
    if pre + ".bn.weight" in sd:
        y = F.conv2d(x, w, None, stride=s, padding=pad)
        if _BN_TRAIN[0]:      # module in train(): batch statistics, running stats updated in place (momentum 0.03, torch_utils.py:152)
            y = F.batch_norm(y, sd[pre + ".bn.running_mean"], sd[pre + ".bn.running_var"],
                             sd[pre + ".bn.weight"], sd[pre + ".bn.bias"], True, BN_MOMENTUM_MODEL, bn_eps)
        else:
            y = F.batch_norm(y, sd[pre + ".bn.running_mean"], sd[pre + ".bn.running_var"],
                             sd[pre + ".bn.weight"], sd[pre + ".bn.bias"], False, 0.0, bn_eps)
    else:
        y = F.conv2d(x, w, sd[pre + ".conv.bias"], stride=s, padding=pad)
    return F.silu(y) if act else y


def bottleneck(x, sd: SD, pre: str, shortcut: bool, bn_eps=BN_EPS_MODEL):
    """common.py:184-194  (c1 == c2 always inside C3, e=1.0)"""
    y = conv_bn_silu(conv_bn_silu(x, sd, pre + ".cv1", 1, 1, bn_eps=bn_eps), sd, pre + ".cv2", 3, 1, bn_eps=bn_eps)
    return x + y if shortcut else y


def c3(x, sd: SD, pre: str, n: int, shortcut: bool, bn_eps=BN_EPS_MODEL):
    """common.py:216-227"""
    a = conv_bn_silu(x, sd, pre + ".cv1", 1, 1, bn_eps=bn_eps)
    for j in range(n):
        a = bottleneck(a, sd, f"{pre}.m.{j}", shortcut, bn_eps)
    b = conv_bn_silu(x, sd, pre + ".cv2", 1, 1, bn_eps=bn_eps)
    return conv_bn_silu(torch.cat((a, b), 1), sd, pre + ".cv3", 1, 1, bn_eps=bn_eps)


def sppf(x, sd: SD, pre: str, k: int = 5, bn_eps=BN_EPS_MODEL):
    """common.py:252-267"""
    x = conv_bn_silu(x, sd, pre + ".cv1", 1, 1, bn_eps=bn_eps)
    y1 = F.max_pool2d(x, k, 1, k // 2)
    y2 = F.max_pool2d(y1, k, 1, k // 2)
    y3 = F.max_pool2d(y2, k, 1, k // 2)
    return conv_bn_silu(torch.cat([x, y1, y2, y3], 1), sd, pre + ".cv2", 1, 1, bn_eps=bn_eps)


# ----------------------------------------------------------------------------------------
# DMFF block                                                      models/common.py:569-891
# ----------------------------------------------------------------------------------------
def adaptive_pool(x, out_h: int, out_w: int, kind: str):
    """AdaptivePool2d.forward, common.py:868-891 -- NOT nn.AdaptiveAvgPool2d semantics."""
    H, W = x.shape[-2:]
    if H > out_h or W > out_w:
        sh, sw = H // out_h, W // out_w
        kh, kw = H - (out_h - 1) * sh, W - (out_w - 1) * sw
        if kind == "avg":
            return F.avg_pool2d(x, (kh, kw), (sh, sw), 0)
        return F.max_pool2d(x, (kh, kw), (sh, sw), 0)
    return x


def dmff_tokens(fea, sd: SD, pre: str, which: str, va: int, ha: int):
    """common.py:817-823: LearnableWeights(avg, max) (:579-587) -> (B,N,C) + pos_emb."""
    w1, w2 = sd[f"{pre}.{which}_coefficient.w1"], sd[f"{pre}.{which}_coefficient.w2"]
    t = adaptive_pool(fea, va, ha, "avg") * w1 + adaptive_pool(fea, va, ha, "max") * w2
    B, C, nh, nw = t.shape
    return t.reshape(B, C, nh * nw).permute(0, 2, 1) + sd[f"{pre}.pos_emb_{which}"], nh, nw


def _ln(x, sd: SD, pre: str):
    return F.layer_norm(x, (x.shape[-1],), sd[pre + ".weight"], sd[pre + ".bias"], LN_EPS)


def _lin(x, sd: SD, pre: str):
    return F.linear(x, sd[pre + ".weight"], sd[pre + ".bias"])


def cross_attention(r, i, sd: SD, pre: str, h: int = 8):
    """CrossAttention.forward, common.py:641-687 (eval: dropouts are identity).
    NOTE the crossing: RGB output = softmax(q_ir k_vis^T) v_vis (:670,682)."""
    B, N, C = r.shape
    d = C // h

    def heads(t):  # (B,N,C) -> (B,h,N,d)
        return t.reshape(B, N, h, d).permute(0, 2, 1, 3)

    rn = _ln(r, sd, pre + ".LN1")                                              # :660
    q_v, k_v, v_v = (heads(_lin(rn, sd, f"{pre}.{n}_proj_vis")) for n in ("que", "key", "val"))
    inn = _ln(i, sd, pre + ".LN2")                                             # :665
    q_i, k_i, v_i = (heads(_lin(inn, sd, f"{pre}.{n}_proj_ir")) for n in ("que", "key", "val"))
    scale = 1.0 / math.sqrt(d)                                                 # :670-671
    att_vis = torch.softmax(q_i @ k_v.transpose(-1, -2) * scale, -1)
    att_ir = torch.softmax(q_v @ k_i.transpose(-1, -2) * scale, -1)
    o_vis = (att_vis @ v_v).permute(0, 2, 1, 3).reshape(B, N, C)               # :682
    o_ir = (att_ir @ v_i).permute(0, 2, 1, 3).reshape(B, N, C)                 # :684
    return _lin(o_vis, sd, pre + ".out_proj_vis"), _lin(o_ir, sd, pre + ".out_proj_ir")


def _mlp(x, sd: SD, pre: str):
    """nn.Sequential(Linear, GELU(erf), Linear, Dropout), common.py:704-715"""
    return _lin(F.gelu(_lin(x, sd, pre + ".0")), sd, pre + ".2")


def cross_transformer_block(r, i, sd: SD, pre: str, loops: int = 1, h: int = 8):
    """CrossTransformerBlock.forward, common.py:737-759.  Block-level LN2 feeds BOTH MLPs;
    ln_input / ln_output / mlp / LN1 are dead parameters."""
    c = [sd[f"{pre}.coefficient{j}.bias"] for j in range(1, 9)]
    for _ in range(loops):
        o_r, o_i = cross_attention(r, i, sd, pre + ".crossatt", h)
        ra = r * c[0] + o_r * c[1]
        ia = i * c[2] + o_i * c[3]
        r = ra * c[4] + _mlp(_ln(ra, sd, pre + ".LN2"), sd, pre + ".mlp_vis") * c[5]
        i = ia * c[6] + _mlp(_ln(ia, sd, pre + ".LN2"), sd, pre + ".mlp_ir") * c[7]
    return r, i


def dmff_block(rgb, ir, sd: SD, pre: str, va: int, ha: int, loops: int = 1, h: int = 8,
               training: bool = False, bn_eps=BN_EPS_MODEL):
    """TransformerFusionBlock.forward, common.py:809-865 (eval: bilinear; train: nearest)."""
    B, C, H, W = rgb.shape
    r, nh, nw = dmff_tokens(rgb, sd, pre, "vis", va, ha)
    i, _, _ = dmff_tokens(ir, sd, pre, "ir", va, ha)
    r, i = cross_transformer_block(r, i, sd, pre + ".crosstransformer.0", loops, h)
    mode = "nearest" if training else "bilinear"

    def up(t):
        t = t.reshape(B, nh, nw, C).permute(0, 3, 1, 2)
        return F.interpolate(t, size=(H, W), mode=mode)

    cat = torch.cat([up(r) + rgb, up(i) + ir], 1)
    return conv_bn_silu(cat, sd, pre + ".conv1x1_out", 1, 1, 0, bn_eps=bn_eps)


# ----------------------------------------------------------------------------------------
# Detect head + model walk                                  models/yolo_test.py:26-70,136-163
# ----------------------------------------------------------------------------------------
def detect(xs: Sequence[torch.Tensor], sd: SD, pre: str, nc: int, anchors, stride, training=False):
    """Detect.forward, yolo_test.py:43-65. eval returns (cat(z), cat(logits), [x_i])."""
    na = len(anchors[0]) // 2
    no = nc + 5
    a = torch.tensor(anchors, dtype=torch.float32).view(len(anchors), -1, 2)
    z, lg, outs = [], [], []
    for li, x in enumerate(xs):
        x = F.conv2d(x, sd[f"{pre}.m.{li}.weight"], sd[f"{pre}.m.{li}.bias"])
        bs, _, ny, nx = x.shape
        x = x.view(bs, na, no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
        outs.append(x)
        if not training:
            yv, xv = torch.meshgrid(torch.arange(ny), torch.arange(nx), indexing="ij")
            grid = torch.stack((xv, yv), 2).view(1, 1, ny, nx, 2).float()
            y = x.sigmoid()
            xy = (y[..., 0:2] * 2.0 - 0.5 + grid) * stride[li]
            wh = (y[..., 2:4] * 2) ** 2 * a[li].view(1, na, 1, 1, 2)
            y = torch.cat((xy, wh, y[..., 4:]), -1)
            z.append(y.view(bs, -1, no))
            lg.append(x[..., 5:].reshape(bs, -1, no - 5))
    return outs if training else (torch.cat(z, 1), torch.cat(lg, 1), outs)


def make_divisible(x, divisor):
    """utils/general.py:234-236"""
    return math.ceil(x / divisor) * divisor


def parse_layers(cfg: dict) -> List[dict]:
    """The subset of parse_model (yolo_test.py:216-302) the Transfusion YAMLs exercise:
    Conv / C3 / SPPF / nn.Upsample / Concat / TransformerFusionBlock / Detect."""
    gd, gw = cfg["depth_multiple"], cfg["width_multiple"]
    nc, anchors = cfg["nc"], cfg["anchors"]
    no = (len(anchors[0]) // 2) * (nc + 5)
    ch: List[int] = []
    layers = []
    for i, (f, n, m, args) in enumerate(cfg["backbone"] + cfg["head"]):
        args = list(args)
        n = max(round(n * gd), 1) if n > 1 else n
        c_in_prev = ch[-1] if ch else 3
        if m in ("Conv", "C3", "SPPF"):
            if m == "Conv" and args[0] == 64:          # yolo_test.py:242-246: both stems take the image
                c1 = 3
            else:
                c1 = ch[f] if ch else 3
            c2 = args[0]
            if c2 != no:
                c2 = make_divisible(c2 * gw, 8)
            spec = dict(type=m, c1=c1, c2=c2, args=args[1:], n=n)
        elif m == "nn.Upsample":
            c2 = c_in_prev
            spec = dict(type="Upsample", scale=args[1], mode=args[2])
        elif m == "Concat":
            c2 = sum(ch[x] for x in f)
            spec = dict(type="Concat")
        elif m == "TransformerFusionBlock":
            c2 = ch[f[0]]
            spec = dict(type="DMFF", c=c2, va=args[1], ha=args[2])
        elif m == "Detect":
            c2 = None
            spec = dict(type="Detect", nc=nc, anchors=anchors, ch=[ch[x] for x in f])
        else:
            raise ValueError(f"oracle does not restate module {m}")
        spec.update(i=i, f=f)
        layers.append(spec)
        ch.append(c2)
    return layers


def model_forward(sd: SD, cfg: dict, rgb, ir, training=False, dmff_loops: int = 1, taps: list = None):
    """Model.forward_once, yolo_test.py:136-163: sequential walk; `f == -4` feeds the IR image."""
    layers = parse_layers(cfg)
    save = set()
    for L in layers:
        fs = [L["f"]] if isinstance(L["f"], int) else L["f"]
        save.update(x % L["i"] for x in fs if x != -1)
    y: List = []
    x = rgb
    _BN_TRAIN[0] = bool(training)
    try:
        return _walk(layers, save, sd, rgb, ir, training, dmff_loops, taps)
    finally:
        _BN_TRAIN[0] = False


def _walk(layers, save, sd: SD, rgb, ir, training: bool, dmff_loops: int, taps: list = None):
    """Training mode restates the dropout-free graph (every nn.Dropout with p = 0): the reference's masks come from torch's
    generator and cannot be reproduced by another implementation."""
    y: List = []
    x = rgb
    for L in layers:
        f, i, t = L["f"], L["i"], L["type"]
        pre = f"model.{i}"
        if f == -4:
            x = ir
        elif f != -1:
            x = y[f] if isinstance(f, int) else [x if j == -1 else y[j] for j in f]
        if t == "Conv":
            a = L["args"]
            k = a[0] if len(a) > 0 else 1
            s = a[1] if len(a) > 1 else 1
            p = a[2] if len(a) > 2 else None
            x = conv_bn_silu(x, sd, pre, k, s, p)
        elif t == "C3":
            shortcut = L["args"][0] if L["args"] else True
            x = c3(x, sd, pre, L["n"], shortcut)
        elif t == "SPPF":
            x = sppf(x, sd, pre, L["args"][0] if L["args"] else 5)
        elif t == "Upsample":
            x = F.interpolate(x, scale_factor=float(L["scale"]), mode=L["mode"])
        elif t == "Concat":
            x = torch.cat(x, 1)
        elif t == "DMFF":
            x = dmff_block(x[0], x[1], sd, pre, L["va"], L["ha"], dmff_loops, training=training)
        elif t == "Detect":
            x = detect(list(x), sd, pre, L["nc"], L["anchors"], [8.0, 16.0, 32.0], training)
        y.append(x if i in save else None)
        if taps is not None:
            taps.append(x)
    return x


def fold_bn(sd: SD, bn_eps=BN_EPS_MODEL) -> SD:
    """fuse_conv_and_bn (utils/torch_utils.py:182-202) applied to every Conv in a state_dict,
    i.e. what Model.fuse() (yolo_test.py:182-190) leaves behind."""
    out = {}
    for k, v in sd.items():
        if ".bn." in k:
            continue
        if k.endswith(".conv.weight") and k[:-len("conv.weight")] + "bn.weight" in sd:
            pre = k[:-len(".conv.weight")]
            g, b = sd[pre + ".bn.weight"], sd[pre + ".bn.bias"]
            mu, var = sd[pre + ".bn.running_mean"], sd[pre + ".bn.running_var"]
            scale = g / torch.sqrt(var + bn_eps)
            out[k] = v * scale.view(-1, 1, 1, 1)
            out[pre + ".conv.bias"] = b - mu * scale
        else:
            out[k] = v
    return out


# ----------------------------------------------------------------------------------------
# detection loss (utils/loss.py), forward
# ----------------------------------------------------------------------------------------
def ciou_xywh(box1, box2, eps=1e-7):
    """bbox_iou(box1, box2, x1y1x2y2=False, CIoU=True), utils/general.py:410-447.  box1: (4, n), box2: (n, 4), xywh."""
    b2 = box2.T
    x1a, x1b = box1[0] - box1[2] / 2, box1[0] + box1[2] / 2
    y1a, y1b = box1[1] - box1[3] / 2, box1[1] + box1[3] / 2
    x2a, x2b = b2[0] - b2[2] / 2, b2[0] + b2[2] / 2
    y2a, y2b = b2[1] - b2[3] / 2, b2[1] + b2[3] / 2
    inter = (torch.min(x1b, x2b) - torch.max(x1a, x2a)).clamp(0) * (torch.min(y1b, y2b) - torch.max(y1a, y2a)).clamp(0)
    w1, h1 = x1b - x1a, y1b - y1a + eps
    w2, h2 = x2b - x2a, y2b - y2a + eps
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    cw = torch.max(x1b, x2b) - torch.min(x1a, x2a)
    ch = torch.max(y1b, y2b) - torch.min(y1a, y2a)
    c2 = cw ** 2 + ch ** 2 + eps
    rho2 = ((x2a + x2b - x1a - x1b) ** 2 + (y2a + y2b - y1a - y1b) ** 2) / 4
    v = (4 / math.pi ** 2) * torch.pow(torch.atan(w2 / h2) - torch.atan(w1 / h1), 2)
    alpha = (v / (v - iou + (1 + eps))).detach()          # torch.no_grad() in the reference (general.py:444-445)
    return iou - (rho2 / c2 + v * alpha)


def compute_loss(p, targets, anchors, hyp: dict, gr: float = 1.0):
    """ComputeLoss.__call__ + build_targets, utils/loss.py:356-463 (fl_gamma = 0, autobalance off, ranking loss disabled as
    upstream).  p: list of (B, na, ny, nx, no) fp32; targets (nt, 6) [image, class, x, y, w, h]; anchors (nl, na, 2) in grid
    units.  Candidates are enumerated target-major per (level, anchor, offset) like the device kernel; the reference's
    order only matters for its IoU-sorted scatter, restated here as a max per cell."""
    nl, na = anchors.shape[0], anchors.shape[1]
    nc = p[0].shape[-1] - 5
    balance = {3: [4.0, 1.0, 0.4]}.get(nl, [4.0, 1.0, 0.25, 0.06, 0.02])
    eps_s = hyp.get("label_smoothing", 0.0)
    cp, cn = 1.0 - 0.5 * eps_s, 0.5 * eps_s
    bce = lambda x, y, pw: F.binary_cross_entropy_with_logits(x, y, pos_weight=torch.tensor([pw]))   # noqa: E731
    lbox = lobj = lcls = torch.zeros(())
    off = torch.tensor([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1]]).float() * 0.5
    nt = targets.shape[0]
    for i, pi in enumerate(p):
        B, _, ny, nx, no = pi.shape
        tobj = torch.zeros_like(pi[..., 0])
        rows = []
        if nt:
            gain = torch.tensor([1, 1, nx, ny, nx, ny]).float()
            t = targets * gain
            for a in range(na):
                r = t[:, 4:6] / anchors[i, a][None]
                keep = torch.max(r, 1.0 / r).max(1)[0] < hyp["anchor_t"]                          # :429-430
                ta = t[keep]
                gxy = ta[:, 2:4]
                gxi = torch.tensor([nx, ny]).float() - gxy
                jk = (gxy % 1.0 < 0.5) & (gxy > 1.0)
                lm = (gxi % 1.0 < 0.5) & (gxi > 1.0)
                sel = torch.stack((torch.ones_like(jk[:, 0]), jk[:, 0], jk[:, 1], lm[:, 0], lm[:, 1]))  # (5, n)   :434-440
                for k in range(5):
                    tk = ta[sel[k]]
                    if tk.shape[0]:
                        rows.append(torch.cat((tk, torch.full((tk.shape[0], 1), float(a)), off[k].expand(tk.shape[0], 2)), 1))
        if rows:
            t = torch.cat(rows, 0)
            b, c, a = t[:, 0].long(), t[:, 1].long(), t[:, 6].long()
            gxy, gwh, offs = t[:, 2:4], t[:, 4:6], t[:, 7:9]
            gij = (gxy - offs).long()
            gi, gj = gij[:, 0].clamp(0, nx - 1), gij[:, 1].clamp(0, ny - 1)                         # :455 (clamp_ in place)
            tbox = torch.cat((gxy - torch.stack((gi, gj), 1), gwh), 1)
            ps = pi[b, a, gj, gi]
            pxy = ps[:, :2].sigmoid() * 2.0 - 0.5
            pwh = (ps[:, 2:4].sigmoid() * 2) ** 2 * anchors[i][a]
            iou = ciou_xywh(torch.cat((pxy, pwh), 1).T, tbox)
            lbox = lbox + (1.0 - iou).mean()
            score = (1.0 - gr) + gr * iou.detach().clamp(0)
            order = torch.argsort(score)                                                           # :374-377: largest IoU written last
            tobj[b[order], a[order], gj[order], gi[order]] = score[order]
            if nc > 1:
                tc = torch.full_like(ps[:, 5:], cn)
                tc[torch.arange(t.shape[0]), c] = cp
                lcls = lcls + bce(ps[:, 5:], tc, hyp["cls_pw"])
        lobj = lobj + bce(pi[..., 4], tobj, hyp["obj_pw"]) * balance[i]
    lbox, lobj, lcls = lbox * hyp["box"], lobj * hyp["obj"], lcls * hyp["cls"]
    bs = p[0].shape[0]
    return (lbox + lobj + lcls) * bs, torch.stack((lbox, lobj, lcls, torch.zeros(())))


# ----------------------------------------------------------------------------------------
# post-processing (utils/general.py)
# ----------------------------------------------------------------------------------------
def greedy_nms(boxes, scores, iou_thres: float):
    """torchvision.ops.nms restated (the reference calls it at utils/general.py:592; torchvision is a third-party
    dependency, unpinned in requirements.txt:10, 0.26 in this image -- csrc/ops/cpu/nms_kernel.cpp): stable sort by
    descending score, keep a box unless a kept higher-scored box overlaps it with IoU > iou_thres.  fp32 arithmetic in
    the kernel's order.  Returns kept indices in score order."""
    import numpy as np
    b = boxes.detach().cpu().numpy().astype(np.float32)
    sc = scores.detach().cpu().numpy().astype(np.float32)
    order = np.argsort(-sc, kind="stable")
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = ((x2 - x1) * (y2 - y1)).astype(np.float32)
    suppressed = np.zeros(len(sc), dtype=bool)
    keep = []
    for _i in range(len(order)):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(int(i))
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest]); yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest]); yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), (xx2 - xx1).astype(np.float32))
        h = np.maximum(np.float32(0), (yy2 - yy1).astype(np.float32))
        inter = (w * h).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / ((areas[i] + areas[rest]).astype(np.float32) - inter).astype(np.float32)
        suppressed[rest[ovr > np.float32(iou_thres)]] = True
    return torch.as_tensor(keep, dtype=torch.long)


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, max_det=300):
    """utils/general.py:518-607, best-class branch (multi_label is off whenever nc == 1, :533): per image
    x = x[obj > conf]; cls *= obj; box = xywh2xyxy; conf, j = cls.max(1); keep conf > conf_thres; optional class filter;
    top max_nms = 30000 by confidence; boxes + cls * 4096 unless agnostic; NMS; first max_det = 300."""
    max_wh, max_nms = 4096, 30000
    out = []
    pred = prediction.float()
    for x in pred:
        x = x[x[:, 4] > conf_thres].clone()
        if not x.shape[0]:
            out.append(torch.zeros((0, 6)))
            continue
        x[:, 5:] *= x[:, 4:5]
        box = x[:, :4].clone()
        box[:, 0] = x[:, 0] - x[:, 2] / 2
        box[:, 1] = x[:, 1] - x[:, 3] / 2
        box[:, 2] = x[:, 0] + x[:, 2] / 2
        box[:, 3] = x[:, 1] + x[:, 3] / 2
        conf, j = x[:, 5:].max(1, keepdim=True)
        x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
        if classes is not None:
            x = x[(x[:, 5:6] == torch.tensor(classes, dtype=x.dtype)).any(1)]
        n = x.shape[0]
        if not n:
            out.append(torch.zeros((0, 6)))
            continue
        if n > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]
        c = x[:, 5:6] * (0 if agnostic else max_wh)
        i = greedy_nms(x[:, :4] + c, x[:, 4], iou_thres)[:max_det]
        out.append(x[i])
    return out


# ----------------------------------------------------------------------------------------
# Work accounting -- used by bench.py to turn time into GFLOP/s
# ----------------------------------------------------------------------------------------
def dmff_flops(B, C, H, W, N, loops=1):
    """F_dmff = B*[L*(48 N C^2 + 8 N^2 C) + 4 H W C^2]"""
    return B * (loops * (48 * N * C * C + 8 * N * N * C) + 4 * H * W * C * C)


def model_conv_flops(cfg: dict, H: int, W: int) -> float:
    """Algorithmic FLOPs (2*MAC) of every conv/linear/attention matmul for one RGB+IR pair."""
    layers = parse_layers(cfg)
    shapes: List = []
    total = 0.0

    def conv(ci, co, k, ho, wo):
        return 2.0 * ho * wo * co * ci * k * k

    for L in layers:
        f, t = L["f"], L["type"]
        if f == -4 or not shapes:
            cin, h, w = 3, H, W
        elif isinstance(f, int):
            cin, h, w = shapes[f] if f != -1 else shapes[-1]
        else:
            srcs = [shapes[-1] if j == -1 else shapes[j] for j in f]
            cin, h, w = sum(s[0] for s in srcs), srcs[0][1], srcs[0][2]
        if t == "Conv":
            a = L["args"]
            k = a[0] if a else 1
            s = a[1] if len(a) > 1 else 1
            p = a[2] if len(a) > 2 else k // 2
            ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
            total += conv(L["c1"], L["c2"], k, ho, wo)
            shapes.append((L["c2"], ho, wo))
        elif t == "C3":
            c1, c2, n = L["c1"], L["c2"], L["n"]
            c_ = c2 // 2
            total += 2 * conv(c1, c_, 1, h, w) + conv(2 * c_, c2, 1, h, w)
            total += n * (conv(c_, c_, 1, h, w) + conv(c_, c_, 3, h, w))
            shapes.append((c2, h, w))
        elif t == "SPPF":
            c1, c2 = L["c1"], L["c2"]
            total += conv(c1, c1 // 2, 1, h, w) + conv(2 * c1, c2, 1, h, w)
            shapes.append((c2, h, w))
        elif t == "Upsample":
            shapes.append((cin, h * L["scale"], w * L["scale"]))
        elif t == "Concat":
            shapes.append((cin, h, w))
        elif t == "DMFF":
            c = L["c"]
            nh, nw = min(L["va"], h), min(L["ha"], w)
            total += dmff_flops(1, c, h, w, nh * nw)
            shapes.append((c, h, w))
        elif t == "Detect":
            na = len(L["anchors"][0]) // 2
            for j in f:
                cj, hj, wj = shapes[j]
                total += conv(cj, na * (L["nc"] + 5), 1, hj, wj)
            shapes.append(None)
    return total


# ----------------------------------------------------------------------------------------
# one training step (train.py:334-344): forward in train mode, loss, backward
# ----------------------------------------------------------------------------------------
def train_step(sd: SD, cfg: dict, rgb, ir, targets, hyp: dict, gr: float = 1.0, dmff_loops: int = 1, autocast_device=None,
               loss_scale: float = 1.0):
    """-> (loss (1,), loss_items (4,), {name: gradient}, [raw Detect maps], updated state).  `sd` is left untouched: parameters
    are cloned into autograd leaves, BatchNorm buffers are cloned and updated like nn.BatchNorm2d does in train().
    autocast_device: run the same graph on that CUDA device under fp16 autocast with a static loss scale -- the reference's own
    training regime (train.py:334-344); the GPU tests use it to measure how far that regime sits from fp32."""
    if autocast_device is not None:
        sd = {k: v.to(autocast_device) for k, v in sd.items()}
        rgb, ir, targets = rgb.to(autocast_device), ir.to(autocast_device), targets.to(autocast_device)
        with torch.autocast("cuda", dtype=torch.float16):
            return _train_step(sd, cfg, rgb, ir, targets, hyp, gr, dmff_loops, loss_scale)
    return _train_step(sd, cfg, rgb, ir, targets, hyp, gr, dmff_loops, loss_scale)


def _train_step(sd: SD, cfg: dict, rgb, ir, targets, hyp: dict, gr: float, dmff_loops: int, loss_scale: float):
    state = {k: v.clone() for k, v in sd.items()}
    params = {}
    for k, v in state.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var", "anchors", "anchor_grid")):
            params[k] = v.requires_grad_(True)
    pred = model_forward(state, cfg, rgb, ir, training=True, dmff_loops=dmff_loops)
    det = [L for L in parse_layers(cfg) if L["type"] == "Detect"][0]
    anchors = torch.tensor(det["anchors"], dtype=torch.float32).view(len(det["anchors"]), -1, 2) / torch.tensor([8.0, 16.0, 32.0]).view(-1, 1, 1)
    # (the loss itself always runs in fp32 on the CPU; under autocast the fp16 maps are cast up, their gradient cast back down)
    loss, items = compute_loss([p.float().cpu() for p in pred], targets.cpu(), anchors, hyp, gr)
    (loss.sum() * loss_scale).backward()
    grads = {k: p.grad / loss_scale for k, p in params.items() if p.grad is not None}
    return loss.detach(), items.detach(), grads, [p.detach() for p in pred], {k: v.detach() for k, v in state.items()}
