"""Golden vectors of the decode variants (models/core/utils_bbox.py:36-251) recorded from the REAL reference functions.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_decode.py      (build container only: needs /root/reference)

yolo.py:75-82 picks one of decode_outputs / decode_outputs_no_sigmoid / decode_outputs_no_sigmoid_all /
decode_outputs_cls_sigmoid by `decode_mode`; decode_outputs_xyxy is used by the loss-side tools.  Inputs: seeded raw
head outputs of three ragged levels for a 96 x 160 input (batch 2, nc = 4)."""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/yolox-drone")


def main():
    from models.core import utils_bbox as ub

    g = torch.Generator().manual_seed(31)
    in_h, in_w, nc = 96, 160, 4
    levels = [torch.randn(2, 5 + nc, in_h // s, in_w // s, generator=g) * 1.5 for s in (8, 16, 32)]
    out = {f"level{i}": l.numpy() for i, l in enumerate(levels)}
    out["input_shape"] = np.array([in_h, in_w])
    for name, fn in (("default", ub.decode_outputs), ("no_sigmoid", ub.decode_outputs_no_sigmoid),
                     ("no_sigmoid_all", ub.decode_outputs_no_sigmoid_all), ("cls_sigmoid", ub.decode_outputs_cls_sigmoid),
                     ("xyxy", ub.decode_outputs_xyxy)):
        with torch.no_grad():
            out[name] = fn([l.clone() for l in levels], [in_h, in_w]).contiguous().numpy()
    np.savez_compressed(HERE / "decode_variants.npz", **out)
    print("decode_variants.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
