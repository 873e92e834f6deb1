"""Golden vectors of the facade's resize_image recorded from the REAL reference function.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_resize.py      (build container only: needs /root/reference)

models/core/utils.py:21-34 resize_image (PIL Image.resize BICUBIC, optional letterbox) on the synthetic test image of
tests/_helpers.py::ufp_synth_image.  The outputs are megabytes, so each case stores its SHA-256 and a corner crop."""
import hashlib
import sys
from pathlib import Path

import numpy as np
from PIL import Image

HERE = Path(__file__).resolve().parent
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/yolox-drone")
sys.path.insert(0, str(HERE.parent))
from _helpers import ufp_synth_image  # noqa: E402

CASES = (  # name, seed, (ih, iw), (w, h), letterbox
    ("stretch", 1, (480, 640), (512, 384), False),
    ("up", 2, (96, 128), (320, 256), False),
    ("letterbox_wide", 3, (300, 500), (416, 416), True),
    ("letterbox_tall", 4, (765, 333), (640, 512), True),
    ("same_width", 5, (200, 320), (320, 256), False),
    ("visdrone", 6, (765, 1360), (1024, 1024), False),
)


def main():
    from models.core.utils import resize_image

    out = {"names": np.array([c[0] for c in CASES])}
    for name, seed, (ih, iw), size, lb in CASES:
        img = ufp_synth_image(seed, ih, iw)
        res = np.array(resize_image(Image.fromarray(img), size, lb))
        assert res.dtype == np.uint8 and res.shape == (size[1], size[0], 3)
        out[f"{name}_meta"] = np.array([seed, ih, iw, size[0], size[1], int(lb)])
        out[f"{name}_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(res).tobytes()).digest(), dtype=np.uint8)
        out[f"{name}_crop"] = res[:64, :96].copy()
        out[f"{name}_center"] = res[size[1] // 2 - 16:size[1] // 2 + 16, size[0] // 2 - 24:size[0] // 2 + 24].copy()
    np.savez_compressed(HERE / "resize_cases.npz", **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
