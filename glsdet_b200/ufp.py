"""UFP stage of UFPMP-Det on the native library: what runs between the coarse detector and
the MP-Det pass, and after it.  Drop-in names for the reference's functions (/root/reference/yolox-ufp):

  * UnifiedForegroundPacking(bbox_list, scale, input_shape, output_shape)  - mmdet/core/ufp/unified_foreground_packing.py:185-197
    (scale_boxes :6-32, ForegroundRegionGeneration :68-103, Packing :140-181, spp.py phsppog :69-168): strictly sequential
    algorithms over a few hundred boxes -> C++ on the host inside libglsdet_b200.so (glsdet_ufp_pack);
  * display_merge_result(results, img, img_name, w, h)                     - ufpmp_det_eval.py:182-193: crop, integer-factor
    bilinear resize (cv2.resize INTER_LINEAR, bit-exact) and paste of every chip -> one CUDA kernel (glsdet_ufp_mosaic);
  * merge_second_stage(rec, second_results, nms_thresh)                    - ufpmp_det_eval.py:270-306: map the second-stage
    detections back through the chips (intersection over the smaller area > 0.9) and merge them per class with py_cpu_nms
    (:149-179, legacy "+1" areas) -> one CUDA kernel, one CTA per class (glsdet_ufp_merge);
  * coco_rows(merged, image_id)                                            - :307-322, the JSON rows (host, trivial).

There is no CPU path for the device functions: a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _native as N

MERGE_CAP = 4096   # mapped detections per class (shared-memory sort of the merge kernel)


def UnifiedForegroundPacking(bbox_list, scale: float, input_shape: Sequence[int], output_shape=(1333, 800)):
    """bbox_list: [n, 4] float32 (x1, y1, x2, y2) coarse detections in image pixels; input_shape = [width, height] of the
    image.  Returns (rows, new_width, new_height), rows = [x0, y0, w, h, new_x, new_y, factor] per foreground region, like
    the reference.  `output_shape` is unused there as well (:140)."""
    lib = N.load()
    if isinstance(bbox_list, torch.Tensor):
        bbox_list = bbox_list.detach().cpu().numpy()
    boxes = np.ascontiguousarray(np.asarray(bbox_list, dtype=np.float32).reshape(-1, 4))
    n = boxes.shape[0]
    rows = np.zeros((max(n, 1), 7), dtype=np.float64)
    n_rows, new_w, new_h = C.c_int32(0), C.c_double(0.0), C.c_double(0.0)
    N.check(lib.glsdet_ufp_pack(boxes.ctypes.data_as(C.c_void_p), n, float(scale), int(input_shape[0]), int(input_shape[1]),
                                rows.ctypes.data_as(C.c_void_p), C.byref(n_rows), C.byref(new_w), C.byref(new_h)),
            "glsdet_ufp_pack")
    return [list(r) for r in rows[:n_rows.value]], new_w.value, new_h.value


def chip_table(rows) -> np.ndarray:
    """[math.floor(_) for _ in result] (ufpmp_det_eval.py:188,271) -> int32 [n, 7] (x1, y1, w, h, new_x, new_y, factor)."""
    return np.array([[math.floor(v) for v in r] for r in rows], dtype=np.int32).reshape(-1, 7)


def display_merge_result(results, img: torch.Tensor, img_name=None, w: float = 0.0, h: float = 0.0, stream=None) -> torch.Tensor:
    """The mosaic of the foreground chips: uint8 [ceil(h), ceil(w), 3] on the device.  `img`: uint8 HWC CUDA tensor (the
    array cv2.imread returns, uploaded).  The reference builds a float64 canvas holding the same integer values."""
    lib = N.load()
    if not (isinstance(img, torch.Tensor) and img.is_cuda and img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 3):
        raise N.NativeError("display_merge_result needs a uint8 [H, W, 3] CUDA tensor (glsdet_b200 has no CPU path)")
    img = img.contiguous()
    chips = chip_table(results)
    H, W = int(img.shape[0]), int(img.shape[1])
    for x1, y1, cw, ch, nx, ny, sf in chips:
        if cw == 0 or ch == 0:
            continue
        if sf not in (1, 2, 4):
            raise ValueError(f"chip scale factor {sf}: the reference only produces 1, 2 or 4")
        if x1 < 0 or y1 < 0 or x1 + cw > W or y1 + ch > H:
            raise ValueError("a chip leaves the image (scale_boxes clips regions to the image, unified_foreground_packing.py:28-31)")
    can_w, can_h = math.ceil(w), math.ceil(h)
    canvas = torch.empty((can_h, can_w, 3), dtype=torch.uint8, device=img.device)
    dchips = torch.from_numpy(chips).to(img.device) if len(chips) else None
    N.check(lib.glsdet_ufp_mosaic(img.data_ptr(), H, W, dchips.data_ptr() if dchips is not None else None, len(chips),
                                  canvas.data_ptr(), can_h, can_w, N.stream_ptr(stream)), "glsdet_ufp_mosaic")
    return canvas


def merge_second_stage(rec, second_results: Sequence[torch.Tensor], nms_thresh: float = 0.6, cap: int = MERGE_CAP,
                       stream=None) -> List[torch.Tensor]:
    """second_results: per class a [K_c, 5] float32 CUDA tensor (x1, y1, x2, y2, score) in mosaic pixels (what the MP-Det
    pass returns with rescale=True).  Returns per class the merged detections [K'_c, 5] in image pixels, score order."""
    lib = N.load()
    nc = len(second_results)
    if nc == 0:
        return []
    dev = second_results[0].device
    if dev.type != "cuda":
        raise N.NativeError("merge_second_stage needs CUDA tensors (glsdet_b200 has no CPU path)")
    counts = [int(t.shape[0]) for t in second_results]
    dets = torch.cat([t.reshape(-1, 5).float() for t in second_results]).contiguous() if sum(counts) else torch.zeros((1, 5), device=dev)
    off = torch.tensor(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32), device=dev)
    chips = chip_table(rec)
    dchips = torch.from_numpy(chips).to(dev) if len(chips) else None
    mapped = torch.empty((nc, cap, 5), dtype=torch.float32, device=dev)
    out = torch.empty((nc, cap, 5), dtype=torch.float32, device=dev)
    cnt = torch.zeros((2, nc), dtype=torch.int32, device=dev)
    N.check(lib.glsdet_ufp_merge(dets.data_ptr(), off.data_ptr(), nc, dchips.data_ptr() if dchips is not None else None,
                                 len(chips), float(nms_thresh), mapped.data_ptr(), cap, out.data_ptr(), cnt[0].data_ptr(),
                                 cnt[1].data_ptr(), N.stream_ptr(stream)), "glsdet_ufp_merge")
    host = cnt.cpu()   # the result is variable-length: one small device->host read, like the reference's .cpu() of its results
    if int(host[1].max()) > cap:
        raise N.NativeError(f"a class maps {int(host[1].max())} detections back, more than cap={cap}")
    return [out[c, :int(host[0, c])] for c in range(nc)]


def coco_rows(merged: Sequence[torch.Tensor], image_id: int) -> List[dict]:
    """ufpmp_det_eval.py:307-322: int() truncation of the corners, [x, y, w, h] boxes, category = class index."""
    rows = []
    for c, dets in enumerate(merged):
        for x1, y1, x2, y2, score in dets.detach().cpu().numpy():
            x1, y1, x2, y2 = int(x1), int(y1), int(x2), int(y2)
            rows.append({"image_id": image_id, "category_id": c, "score": float(score), "bbox": [x1, y1, x2 - x1, y2 - y1]})
    return rows
