"""Golden vectors for the CSPDarknet backbone from the REAL reference.

Run in the build container only (the GPU box has no /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_backbone.py

The real models/ffa/yolox_ffa.py YoloBody(10, 's') (import shim D1 only, stdout swallowed: D3) is loaded strictly with
the seeded calibrated weights and run from an IMAGE: the backbone's four feature maps and the raw per-level logits
of the whole model are stored in tests/golden/backbone_s.npz together with the image.
"""
import contextlib
import io
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/yolox-drone")


def main():
    import models.ffa.ffa as ffa_mod
    import models.ffa.yolox_ffa as yf

    yf.FTT = ffa_mod.FFA  # D1
    from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict

    nc, phi, seed, in_h, in_w = 10, "s", 0, 96, 128
    sd = synthetic_state_dict(nc, phi, seed=seed, flavour="calibrated")
    with contextlib.redirect_stdout(io.StringIO()):
        net = yf.YoloBody(nc, phi).eval()
    net.load_state_dict(sd, strict=True)
    x = synthetic_images(2, in_h, in_w, seed=7)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        feats = net.backbone.backbone(x)
        logits = net(x)
    out = {"image": x.numpy()}
    out.update({k: v.numpy() for k, v in feats.items()})
    out.update({f"logits{i}": t.numpy() for i, t in enumerate(logits)})
    np.savez_compressed(HERE / "backbone_s.npz", **out)
    print({k: (v.shape, float(np.abs(v).mean())) for k, v in out.items()})


if __name__ == "__main__":
    main()
