"""Synthetic weights / inputs for the oracle-side tests -- TEST INFRASTRUCTURE.

The generators themselves live in icafusion_b200/synth.py (the benchmark uses them too); this module re-exports them and
adds the state_dict layout of the reference Model derived from the oracle's own layer parser."""
from collections import OrderedDict

from icafusion_b200.synth import (dmff_param_shapes, synth_features, synth_images, synth_state_dict,  # noqa: F401
                                  synth_tensor)


def model_param_shapes(cfg: dict) -> "OrderedDict[str, tuple]":
    """state_dict layout (names -> shapes) of models/yolo_test.py Model(cfg) for the module
    subset the Transfusion configs use; checked against the real reference by gen_golden."""
    from oracle.icaf_oracle import parse_layers
    s: "OrderedDict[str, tuple]" = OrderedDict()

    def conv(pre, c1, c2, k):
        s[pre + ".conv.weight"] = (c2, c1, k, k)
        for b, shp in (("weight", (c2,)), ("bias", (c2,)), ("running_mean", (c2,)),
                       ("running_var", (c2,)), ("num_batches_tracked", ())):
            s[f"{pre}.bn.{b}"] = shp

    for L in parse_layers(cfg):
        pre, t = f"model.{L['i']}", L["type"]
        if t == "Conv":
            conv(pre, L["c1"], L["c2"], L["args"][0] if L["args"] else 1)
        elif t == "C3":
            c1, c2, n = L["c1"], L["c2"], L["n"]
            c_ = c2 // 2
            conv(pre + ".cv1", c1, c_, 1)
            conv(pre + ".cv2", c1, c_, 1)
            conv(pre + ".cv3", 2 * c_, c2, 1)
            for j in range(n):
                conv(f"{pre}.m.{j}.cv1", c_, c_, 1)
                conv(f"{pre}.m.{j}.cv2", c_, c_, 3)
        elif t == "SPPF":
            c1, c2 = L["c1"], L["c2"]
            conv(pre + ".cv1", c1, c1 // 2, 1)
            conv(pre + ".cv2", 2 * c1, c2, 1)
        elif t == "DMFF":
            s.update(dmff_param_shapes(L["c"], L["va"] * L["ha"], pre))
        elif t == "Detect":
            na = len(L["anchors"][0]) // 2
            for j, c in enumerate(L["ch"]):
                s[f"{pre}.m.{j}.weight"] = (na * (L["nc"] + 5), c, 1, 1)
                s[f"{pre}.m.{j}.bias"] = (na * (L["nc"] + 5),)
    return s
