"""Drop-in for yolox-drone/models/core/utils.py (cvtColor :9-14, resize_image :21-34, get_classes :40-44,
preprocess_input :47-51) with the image work on the device.

resize_image returns a uint8 [h, w, 3] CUDA tensor holding exactly the bytes of the PIL image the reference builds
(Image.resize(..., Image.BICUBIC), optional letterbox paste onto (128, 128, 128)): Pillow's 8-bit resampler is restated in
csrc/resize.cu (fixed-point tables on the host, two passes on the device).  preprocess_input of a device tensor is what
YoloBody.detect_uint8 fuses into the Focus kernel; the host version below is the reference's arithmetic for callers that
still preprocess on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Sequence, Tuple

import numpy as np
import torch

from . import _native as N


def cvtColor(image):
    """utils.py:9-14: anything that is not an H x W x 3 image is converted to RGB (PIL, host)."""
    if len(np.shape(image)) == 3 and np.shape(image)[2] == 3:
        return image
    return image.convert("RGB")


def get_classes(classes_path):
    """utils.py:40-44."""
    with open(classes_path, encoding="utf-8") as f:
        class_names = [c.strip() for c in f.readlines()]
    return class_names, len(class_names)


def preprocess_input(image):
    """utils.py:47-51 on a host float32 array (in place, like the reference)."""
    image /= 255.0
    image -= np.array([0.485, 0.456, 0.406])
    image /= np.array([0.229, 0.224, 0.225])
    return image


_tables: Dict[Tuple[int, int, str], Tuple[torch.Tensor, torch.Tensor, int]] = {}


def _table(in_size: int, out_size: int, device) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """Pillow's bicubic window bounds and 22-bit weights for in_size -> out_size samples, cached on the device."""
    key = (in_size, out_size, str(device))
    hit = _tables.get(key)
    if hit is None:
        lib = N.load()
        ks = lib.glsdet_pil_bicubic_ksize(in_size, out_size)
        bounds = np.zeros((out_size, 2), dtype=np.int32)
        kk = np.zeros((out_size, ks), dtype=np.int32)
        N.check(lib.glsdet_pil_bicubic_table(in_size, out_size, bounds.ctypes.data_as(C.c_void_p), kk.ctypes.data_as(C.c_void_p)),
                "glsdet_pil_bicubic_table")
        if len(_tables) > 64:
            _tables.clear()
        hit = _tables[key] = (torch.from_numpy(bounds).to(device), torch.from_numpy(kk).to(device), ks)
    return hit


def letterbox_geometry(iw: int, ih: int, w: int, h: int, letterbox_image: bool):
    """(nw, nh, off_x, off_y) of utils.py:24-33."""
    if not letterbox_image:
        return w, h, 0, 0
    scale = min(w / iw, h / ih)
    nw, nh = int(iw * scale), int(ih * scale)
    return nw, nh, (w - nw) // 2, (h - nh) // 2


def resize_image(image, size: Sequence[int], letterbox_image: bool, device=None, out: torch.Tensor = None, stream=None) -> torch.Tensor:
    """utils.py:21-34.  `image`: PIL RGB image, uint8 H x W x 3 array, or uint8 [H, W, 3] CUDA tensor; size = (w, h).
    Returns the resized (letterboxed) image as a uint8 [h, w, 3] CUDA tensor - the bytes np.array(new_image) holds in
    the reference."""
    lib = N.load()
    if isinstance(image, torch.Tensor):
        src = image
    else:
        arr = np.asarray(image)
        if arr.ndim != 3 or arr.shape[2] != 3 or arr.dtype != np.uint8:
            raise ValueError("resize_image expects an RGB uint8 image (use cvtColor first)")
        src = torch.from_numpy(np.array(arr, copy=True))
    dev = torch.device(device) if device is not None else (src.device if src.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    if dev.type != "cuda":
        raise N.NativeError("resize_image runs on a CUDA device (glsdet_b200 has no CPU path)")
    src = src.to(dev, non_blocking=True).contiguous()
    ih, iw = int(src.shape[0]), int(src.shape[1])
    w, h = int(size[0]), int(size[1])
    nw, nh, ox, oy = letterbox_geometry(iw, ih, w, h, letterbox_image)
    if nw <= 0 or nh <= 0:
        raise ValueError("the letterboxed image is empty")
    canvas = out if out is not None else torch.empty((h, w, 3), dtype=torch.uint8, device=dev)
    assert canvas.is_cuda and canvas.dtype == torch.uint8 and tuple(canvas.shape) == (h, w, 3) and canvas.is_contiguous()
    bh = kh = bv = kv = None
    ksh = ksv = 0
    if nw != iw:
        bh, kh, ksh = _table(iw, nw, dev)
    if nh != ih:
        bv, kv, ksv = _table(ih, nh, dev)
    tmp = torch.empty((ih, nw, 3), dtype=torch.uint8, device=dev) if (nw != iw and nh != ih) else None
    N.check(lib.glsdet_resize_bicubic_u8(src.data_ptr(), ih, iw, canvas.data_ptr(), h, w, nh, nw, oy, ox, 128, N.ptr(tmp),
                                         N.ptr(bh), N.ptr(kh), ksh, N.ptr(bv), N.ptr(kv), ksv, N.stream_ptr(stream)),
            "glsdet_resize_bicubic_u8")
    return canvas
