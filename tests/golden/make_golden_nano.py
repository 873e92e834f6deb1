"""Golden vectors for phi = 'nano' (depthwise-separable DWConv blocks, models/base/baseConv.py:22-30) from the REAL
reference: models/ffa/yolox_ffa.py YoloBody(nc, 'nano') (GLSDet P0, import shim D1), models/base/yolox.py
YoloBody(nc, 'nano') (stock three-level YOLOX), models/new/yolox10.py (GLSDet P1, import shim D2) and
models/block/non_local/yolo_patch_nonlocal_plus.py (GLSDet P2), all loaded strictly with the seeded synthetic weights and run from an
IMAGE.  Stored: the image, the backbone's feature maps, the neck outputs, the raw per-level logits, the decoded
predictions and the NMS rows.  tests/golden/nano_cases.npz pins oracle/ref_path.py's DWConv branch and, through it, the
CUDA path (csrc/dwconv.cu + the tcgen05 1x1 convs).

Weights: glsdet_b200.synthetic "calibrated" flavour, recorded by
    python tools/calibrate_synthetic.py --phi nano --nc 10 --seed 11 --size 256 --bn-beta 1.0
    python tools/calibrate_synthetic.py --phi nano --nc 3 --seed 12 --variant stock --size 256 --bn-beta 1.0
    python tools/calibrate_synthetic.py --phi nano --nc 3 --seed 13 --variant p1 --size 256 --bn-beta 1.0
    python tools/calibrate_synthetic.py --phi nano --nc 3 --seed 14 --variant p2 --size 256 --bn-beta 1.0

Run in the build container only (the GPU box has no /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_nano.py
"""
import contextlib
import io
import json
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/yolox-drone")

CASES = {"p0": dict(variant="ffa", nc=10, seed=11, batch=2, in_h=96, in_w=128, conf=0.01, nms_thr=0.65),
         "stock": dict(variant="stock", nc=3, seed=12, batch=1, in_h=64, in_w=96, conf=0.01, nms_thr=0.65),
         "p1": dict(variant="p1", nc=3, seed=13, batch=1, in_h=128, in_w=192, conf=0.01, nms_thr=0.65),
         "p2": dict(variant="p2", nc=3, seed=14, batch=1, in_h=128, in_w=192, conf=0.01, nms_thr=0.65)}


def main():
    import types

    import models.base.yolox as yb
    import models.block.non_local.yolo_patch_nonlocal_plus as yp2
    import models.ffa.ffa as ffa_mod
    import models.ffa.yolox_ffa as yf
    import models.new.darknet as dk
    from models.core import utils_bbox as ub

    yf.FTT = ffa_mod.FFA  # D1
    sys.modules.setdefault("models.decouple", types.ModuleType("models.decouple"))   # D2: yolox10.py imports models.decouple.darknet
    sys.modules["models.decouple.darknet"] = dk
    import models.new.yolox10 as y10
    from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict

    out, meta = {}, {}
    for name, c in CASES.items():
        mod = {"ffa": yf, "stock": yb, "p1": y10, "p2": yp2}[c["variant"]]
        sd = synthetic_state_dict(c["nc"], "nano", seed=c["seed"], flavour="calibrated", variant=c["variant"])
        with contextlib.redirect_stdout(io.StringIO()):
            net = mod.YoloBody(c["nc"], "nano").eval()
        net.load_state_dict(sd, strict=True)
        x = synthetic_images(c["batch"], c["in_h"], c["in_w"], seed=c["seed"] + 50)
        hw = [c["in_h"], c["in_w"]]
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            feats = net.backbone.backbone(x)
            neck = net.backbone(x)
            logits = net(x)
            pred = ub.decode_outputs([o.clone() for o in logits], hw).contiguous()
            res = ub.non_max_suppression(pred.clone(), c["nc"], hw, np.array(hw), False, conf_thres=c["conf"],
                                         nms_thres=c["nms_thr"])
        out[f"{name}_image"] = x.numpy()
        for k, v in feats.items():
            out[f"{name}_{k}"] = v.numpy()
        for i, t in enumerate(neck):
            out[f"{name}_neck{i}"] = t.numpy()
        for i, t in enumerate(logits):
            out[f"{name}_logits{i}"] = t.numpy()
        out[f"{name}_pred"] = pred.numpy()
        for i, r in enumerate(res):
            out[f"{name}_nms{i}"] = r if r is not None else np.zeros((0, 7), np.float32)
        meta[name] = dict(c, n_keys=len(sd), keys=list(net.state_dict().keys()),
                          kept=[int(len(out[f"{name}_nms{i}"])) for i in range(c["batch"])],
                          logit_rms=[float(t.pow(2).mean().sqrt()) for t in logits])
        print(name, meta[name]["kept"], meta[name]["logit_rms"])
    np.savez_compressed(HERE / "nano_cases.npz", **out)
    (HERE / "nano_meta.json").write_text(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
