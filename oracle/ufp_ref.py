"""CPU restatement of the UFP stage of UFPMP-Det - TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(glsdet_b200/ufp.py, csrc/ufp.cu) never does.  Reference files (read-only, /root/reference/yolox-ufp):

  * mmdet/core/ufp/unified_foreground_packing.py:6-32   scale_boxes
  * mmdet/core/ufp/unified_foreground_packing.py:35-47  get_merge_bbox_aera
  * mmdet/core/ufp/unified_foreground_packing.py:68-103 ForegroundRegionGeneration
  * mmdet/core/ufp/unified_foreground_packing.py:140-181 Packing
  * mmdet/core/ufp/unified_foreground_packing.py:185-197 UnifiedForegroundPacking
  * mmdet/core/ufp/spp.py:69-112 phsppog, :115-168 recursive_packing
  * ufpmp_det_eval.py:36-50 compute_iof, :149-179 py_cpu_nms, :182-193 display_merge_result,
    :270-296 map-back of the second-stage detections, :299-306 per-class merge NMS, :307-322 COCO rows

Pinned: tests/golden/ufp_cases.npz is produced by tests/golden/make_golden_ufp.py from the REAL functions above (the
two ufp modules are imported as they are; compute_iof / py_cpu_nms / display_merge_result are executed from the
reference file's own source through `ast`, because the script imports mmcv / pycocotools at module level; the
map-back loop lives inside main() and is restated here, line by line, on top of the real compute_iof).
Arithmetic types follow NumPy 2 promotion rules (numpy 2.3 is what this image has; the reference pins no version):
float32 boxes stay float32 when combined with Python ints, float32 / int64 arrays give float64.  cv2.resize is
OpenCV 4.13's INTER_LINEAR for uint8 (restated in resize_linear_u8, verified bit-exact against cv2 here).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np


# ------------------------------------------------------------------------------------------------ packing (host side)
def scale_boxes(bboxes: np.ndarray, scale: float, image_shape=(1333, 1333)) -> np.ndarray:
    """unified_foreground_packing.py:6-32."""
    assert bboxes.shape[1] == 4
    w_half = (bboxes[:, 2] - bboxes[:, 0]) * 0.5
    h_half = (bboxes[:, 3] - bboxes[:, 1]) * 0.5
    x_c = (bboxes[:, 2] + bboxes[:, 0]) * 0.5
    y_c = (bboxes[:, 3] + bboxes[:, 1]) * 0.5
    w_half = w_half * scale
    h_half = h_half * scale
    w, h = image_shape
    out = np.zeros_like(bboxes)
    out[:, 0] = np.clip(x_c - w_half, 0, w - 1)
    out[:, 2] = np.clip(x_c + w_half, 0, w - 1)
    out[:, 1] = np.clip(y_c - h_half, 0, h - 1)
    out[:, 3] = np.clip(y_c + h_half, 0, h - 1)
    return out


def foreground_region_generation(bbox_list: np.ndarray, scaled: np.ndarray):
    """unified_foreground_packing.py:68-103.  Greedy, order-dependent merge: region idx absorbs every later-tested box
    whose union rectangle with it is smaller than the sum of the two areas; the running rectangle A grows at once."""
    n = bbox_list.shape[0]
    f32 = np.float32
    areas = ((bbox_list[:, 2] - bbox_list[:, 0] + 1) * (bbox_list[:, 3] - bbox_list[:, 1] + 1)).astype(f32)
    avg = areas.copy()
    cnt = np.ones(n, dtype=np.int64)
    used = [True] * n
    scaled = scaled.astype(f32).copy()
    for i in range(n):
        if not used[i]:
            continue
        a = [f32(v) for v in scaled[i]]
        for j in range(n):
            if not used[j] or i == j:
                continue
            b = scaled[j]
            a1 = f32(a[2] - a[0]) * f32(a[3] - a[1])
            a2 = f32(b[2] - b[0]) * f32(b[3] - b[1])
            x0, y0, x1, y1 = min(a[0], b[0]), min(a[1], b[1]), max(a[2], b[2]), max(a[3], b[3])
            merge = f32(x1 - x0) * f32(y1 - y0)
            if merge < f32(a1 + a2):
                a = [x0, y0, x1, y1]
                used[j] = False
                avg[i] = f32(avg[i] + avg[j])
                cnt[i] += cnt[j]
        scaled[i] = a
    mean = avg / cnt                                # float32 / int64 -> float64
    factor = np.where(mean < 32 * 32, 4, np.where(mean < 96 * 96, 2, 1)).astype(np.int64)
    used = np.array(used, dtype=bool)
    return scaled[used], factor[used]


def _recursive_packing(x, y, w, h, remaining, indices, result):
    """spp.py:115-168 with D = 0 (no rotation): j only takes the value 0."""
    priority = 6
    best = None
    for idx in indices:
        rw, rh = remaining[idx]
        if priority > 1 and rw == w and rh == h:
            priority, best = 1, idx
            break
        elif priority > 2 and rw == w and rh < h:
            priority, best = 2, idx
        elif priority > 3 and rw < w and rh == h:
            priority, best = 3, idx
        elif priority > 4 and rw < w and rh < h:
            priority, best = 4, idx
        elif priority > 5:
            priority, best = 5, idx
    if priority < 5:
        omega, d = remaining[best]
        result[best] = (x, y, omega, d)
        indices.remove(best)
        if priority == 2:
            _recursive_packing(x, y + d, w, h - d, remaining, indices, result)
        elif priority == 3:
            _recursive_packing(x + omega, y, w - omega, h, remaining, indices, result)
        elif priority == 4:
            min_w = min_h = float("inf")   # sys.maxsize in the reference: only compared, never stored
            for idx in indices:
                min_w = min(min_w, remaining[idx][0])
                min_h = min(min_h, remaining[idx][1])
            min_w = min(min_h, min_w)
            min_h = min_w
            if w - omega < min_w:
                _recursive_packing(x, y + d, w, h - d, remaining, indices, result)
            elif h - d < min_h:
                _recursive_packing(x + omega, y, w - omega, h, remaining, indices, result)
            elif omega < min_w:
                _recursive_packing(x + omega, y, w - omega, d, remaining, indices, result)
                _recursive_packing(x, y + d, w, h - d, remaining, indices, result)
            else:
                _recursive_packing(x, y + d, omega, h - d, remaining, indices, result)
                _recursive_packing(x + omega, y, w - omega, h, remaining, indices, result)


def phsppog(width: float, rectangles: Sequence[Sequence[float]], sorting: str = "width"):
    """spp.py:69-112: PH strip-packing heuristic, no rotation, guillotine cuts."""
    wh = 0 if sorting == "width" else 1
    result = [None] * len(rectangles)
    remaining = [list(r) for r in rectangles]
    order = sorted(range(len(remaining)), key=lambda i: -remaining[i][wh])
    H = 0
    while order:
        idx = order.pop(0)
        r = remaining[idx]
        result[idx] = (0, H, r[0], r[1])
        x, y, w, h, H = r[0], H, width - r[0], r[1], H + r[1]
        _recursive_packing(x, y, w, h, remaining, order, result)
    return H, result


def packing(regions: np.ndarray, factor: np.ndarray):
    """unified_foreground_packing.py:140-181: binary search of the strip width (the LAST probe's layout is kept), then
    every rectangle is matched back to the first unused region of the same scaled size."""
    boxes = []
    for i in range(len(factor)):
        w = regions[i][2] - regions[i][0]
        h = regions[i][3] - regions[i][1]
        boxes.append([float(w) * int(factor[i]), float(h) * int(factor[i])])
    lo, hi = 300, 2666
    rects = []
    while lo <= hi:
        mid = (lo + hi) / 2
        height, rects = phsppog(mid, boxes, sorting="height")
        if height > mid:
            lo = mid + 1
        else:
            hi = mid - 1
    flag = [True] * regions.shape[0]
    result = []
    new_w = new_h = 0
    for (x, y, w, h) in rects:
        new_w = max(new_w, x + w)
        new_h = max(new_h, y + h)
        for i in range(regions.shape[0]):
            if not flag[i]:
                continue
            f = int(factor[i])
            _w = regions[i, 2] - regions[i, 0]
            _h = regions[i, 3] - regions[i, 1]
            if float(_w) * f == w and float(_h) * f == h:
                flag[i] = False
                result.append([float(regions[i, 0]), float(regions[i, 1]), float(_w), float(_h), float(x), float(y), float(f)])
    return result, new_w, new_h


def unified_foreground_packing(bbox_list: np.ndarray, scale: float, input_shape, output_shape=(1333, 800)):
    """unified_foreground_packing.py:185-197 -> (rows [x0, y0, w, h, new_x, new_y, factor], new_width, new_height)."""
    scaled = scale_boxes(bbox_list, scale, input_shape)
    regions, factor = foreground_region_generation(bbox_list, scaled)
    return packing(regions, factor)


# ------------------------------------------------------------------------------------------------ mosaic
def resize_linear_u8(img: np.ndarray, sf: int) -> np.ndarray:
    """cv2.resize(img, (w * sf, h * sf)) for uint8, INTER_LINEAR, integer factor (OpenCV 4.13 imgproc/resize.cpp,
    third party, not under /root/reference): 11-bit fixed-point coefficients; columns clamp the coefficient at the
    borders, rows clamp the source row and keep the coefficient; dst = ((b0 (S0 >> 4) >> 16) + (b1 (S1 >> 4) >> 16) + 2) >> 2."""
    h, w, _ = img.shape
    src = img.astype(np.int64)

    def taps(n_src, n_dst, clamp_coef):
        i0 = np.zeros(n_dst, np.int64)
        i1 = np.zeros(n_dst, np.int64)
        a = np.zeros((n_dst, 2), np.int64)
        for d in range(n_dst):
            f = (d + 0.5) * (n_src / n_dst) - 0.5
            s = int(math.floor(f))
            f -= s
            if clamp_coef:
                if s < 0:
                    f, s = 0.0, 0
                if s >= n_src - 1:
                    f, s = 0.0, n_src - 1
            i0[d] = min(max(s, 0), n_src - 1)
            i1[d] = min(max(s + 1, 0), n_src - 1)
            a[d] = (int(np.rint((1.0 - f) * 2048)), int(np.rint(f * 2048)))
        return i0, i1, a

    x0, x1, xa = taps(w, w * sf, True)
    y0, y1, ya = taps(h, h * sf, False)
    rows = src[:, x0, :] * xa[:, 0][None, :, None] + src[:, x1, :] * xa[:, 1][None, :, None]
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    out = (((b0 * (rows[y0] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def chip_ints(rows) -> np.ndarray:
    """[math.floor(_) for _ in result] of ufpmp_det_eval.py:188,271 -> int64 [n, 7]."""
    return np.array([[math.floor(v) for v in r] for r in rows], dtype=np.int64).reshape(-1, 7)


def display_merge_result(rows, img: np.ndarray, w: float, h: float) -> np.ndarray:
    """ufpmp_det_eval.py:182-193: the mosaic fed to the second detector (float64 canvas of uint8 values there; uint8
    here).  `img` is what cv2.imread returned (BGR uint8 HWC)."""
    W, H = math.ceil(w), math.ceil(h)
    canvas = np.zeros((H, W, 3), dtype=np.uint8)
    for x1, y1, cw, ch, nx, ny, sf in chip_ints(rows):
        if cw == 0 or ch == 0:
            continue
        crop = img[y1:y1 + ch, x1:x1 + cw, :]
        assert crop.shape[:2] == (ch, cw), "scale_boxes keeps every region inside the image"
        canvas[ny:ny + ch * sf, nx:nx + cw * sf, :] = resize_linear_u8(crop, int(sf))
    return canvas


# ------------------------------------------------------------------------------------------------ map-back + merge
def compute_iof(pos1, pos2):
    """ufpmp_det_eval.py:36-50 (pos1: float32 detection, pos2: integer chip rectangle)."""
    left1, top1, right1, down1 = pos1
    left2, top2, right2, down2 = pos2
    area1 = (right1 - left1) * (down1 - top1)
    area2 = (right2 - left2) * (down2 - top2)
    left, right = max(left1, left2), min(right1, right2)
    top, bottom = max(top1, top2), min(down1, down2)
    if left >= right or top >= bottom:
        return 0
    inter = (right - left) * (bottom - top)
    return inter / min(area1, area2)


def map_back(rows, second_results: List[np.ndarray]) -> List[np.ndarray]:
    """ufpmp_det_eval.py:270-296: every second-stage detection [x1, y1, x2, y2, score] (float32, mosaic pixels, one
    array per class) that lies > 0.9 (intersection over the smaller area) inside a chip goes back to image coordinates;
    per class, chip-major order.  float32 arithmetic (NumPy 2: Python ints are weak)."""
    out = [[] for _ in second_results]
    for ox, oy, cw, ch, nx, ny, sf in (tuple(int(v) for v in r) for r in chip_ints(rows)):
        chip = [nx, ny, nx + cw * sf, ny + ch * sf]
        for c, dets in enumerate(second_results):
            for x1, y1, x2, y2, score in dets:
                if compute_iof([x1, y1, x2, y2], chip) > 0.9:
                    nw = (x2 - x1) / sf
                    nh = (y2 - y1) / sf
                    bx = (x1 - nx) / sf + ox
                    by = (y1 - ny) / sf + oy
                    out[c].append([bx, by, bx + nw, by + nh, score])
    return [np.array(o, dtype=np.float32).reshape(-1, 5) for o in out]


def py_cpu_nms(dets: np.ndarray, thresh: float) -> List[int]:
    """ufpmp_det_eval.py:149-179: areas with the legacy +1, keep while ovr <= thresh.  The reference orders by
    scores.argsort()[::-1] (unstable for ties); this restatement breaks ties by the lower index."""
    dets = np.asarray(dets)
    x1, y1, x2, y2, scores = (dets[:, k] for k in range(5))
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = np.lexsort((np.arange(len(scores)), -scores))
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        xx1 = np.maximum(x1[i], x1[order[1:]])
        yy1 = np.maximum(y1[i], y1[order[1:]])
        xx2 = np.minimum(x2[i], x2[order[1:]])
        yy2 = np.minimum(y2[i], y2[order[1:]])
        w = np.maximum(0.0, xx2 - xx1 + 1)
        h = np.maximum(0.0, yy2 - yy1 + 1)
        inter = w * h
        ovr = inter / (areas[i] + areas[order[1:]] - inter)
        order = order[np.where(ovr <= thresh)[0] + 1]
    return keep


def merge_results(rows, second_results: List[np.ndarray], thresh: float = 0.6) -> List[np.ndarray]:
    """ufpmp_det_eval.py:270-306: map-back, then the per-class merge NMS; [K_c, 5] float32 per class, score order."""
    out = []
    for dets in map_back(rows, second_results):
        out.append(dets[py_cpu_nms(dets, thresh)] if len(dets) else dets)
    return out


def coco_rows(merged: List[np.ndarray], image_id: int) -> List[dict]:
    """ufpmp_det_eval.py:307-322: int() truncation of the corners, [x, y, w, h] boxes, category = class index."""
    rows = []
    for c, dets in enumerate(merged):
        for x1, y1, x2, y2, score in dets:
            x1, y1, x2, y2 = int(x1), int(y1), int(x2), int(y2)
            rows.append({"image_id": image_id, "category_id": c, "score": float(score), "bbox": [x1, y1, x2 - x1, y2 - y1]})
    return rows
