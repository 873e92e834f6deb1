"""Drop-in for yolox-drone/models/core/utils_bbox.py (decode_outputs :254-306, non_max_suppression :375-484,
yolo_correct_boxes :8-33) running on the native sm_100a kernels.  Same names, arguments and return types;
differences a caller can observe:

  * `prediction` is not overwritten in place (the reference converts it to corner form in place, :386);
  * decode_outputs returns a contiguous [B, A, 5+nc] tensor (the reference returns a permuted view);
  * non_max_suppression has one extra keyword, `strategy`, naming which torchvision batched_nms branch to
    reproduce bit-exactly (default: torchvision's own dispatch for CUDA tensors).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _native as N

STRATEGIES = {"trick": N.NMS_COORD_TRICK, "per_class": N.NMS_PER_CLASS, "auto_cuda": N.NMS_AUTO_CUDA,
              "auto_cpu": N.NMS_AUTO_CPU, "mmcv": N.NMS_MMCV}


def decode_outputs(outputs: Sequence[torch.Tensor], input_shape: Sequence[int]) -> torch.Tensor:
    """utils_bbox.py:254-306: list of raw [B, 5+nc, h, w] maps -> [B, sum(h*w), 5+nc] with normalised
    (cx, cy, w, h), sigmoid(obj), sigmoid(cls)."""
    return _decode(outputs, input_shape, N.DECODE_SIGMOID_OBJ | N.DECODE_SIGMOID_CLS | N.DECODE_NORMALISE)


def decode_outputs_no_sigmoid(outputs: Sequence[torch.Tensor], input_shape: Sequence[int]) -> torch.Tensor:
    """utils_bbox.py:149-200 (decode_mode 'obj_sigmoid', yolo.py:77-78): sigmoid on the objectness only, classes raw."""
    return _decode(outputs, input_shape, N.DECODE_SIGMOID_OBJ | N.DECODE_NORMALISE)


def decode_outputs_no_sigmoid_all(outputs: Sequence[torch.Tensor], input_shape: Sequence[int]) -> torch.Tensor:
    """utils_bbox.py:202-251 (decode_mode 'no_sigmoid', yolo.py:79-80): objectness and classes stay raw."""
    return _decode(outputs, input_shape, N.DECODE_NORMALISE)


def decode_outputs_cls_sigmoid(outputs: Sequence[torch.Tensor], input_shape: Sequence[int]) -> torch.Tensor:
    """utils_bbox.py:95-147 (decode_mode 'cls_sigmoid', yolo.py:81-82): sigmoid on the classes only, objectness raw."""
    return _decode(outputs, input_shape, N.DECODE_SIGMOID_CLS | N.DECODE_NORMALISE)


def decode_outputs_xyxy(outputs: Sequence[torch.Tensor], input_shape: Sequence[int]) -> torch.Tensor:
    """utils_bbox.py:36-93: corner boxes (x1, y1, x2, y2) in input pixels, raw objectness / class logits."""
    return _decode(outputs, input_shape, N.DECODE_XYXY)


def _decode(outputs: Sequence[torch.Tensor], input_shape: Sequence[int], mode: int) -> torch.Tensor:
    lib = N.load()
    outs = [o.float().contiguous() for o in outputs]
    if not outs or not outs[0].is_cuda:
        raise N.NativeError("decode_outputs needs CUDA tensors (glsdet_b200 has no CPU path)")
    b, nch = outs[0].shape[:2]
    n = len(outs)
    ptrs = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    hs = (C.c_int32 * n)(*[o.shape[2] for o in outs])
    ws = (C.c_int32 * n)(*[o.shape[3] for o in outs])
    a = sum(o.shape[2] * o.shape[3] for o in outs)
    pred = torch.empty((b, a, nch), dtype=torch.float32, device=outs[0].device)
    N.check(lib.glsdet_decode_outputs_mode(ptrs, hs, ws, n, b, nch - 5, int(input_shape[0]), int(input_shape[1]), mode,
                                           pred.data_ptr(), N.stream_ptr()), "glsdet_decode_outputs_mode")
    return pred


def yolo_correct_boxes(box_xy, box_wh, input_shape, image_shape, letterbox_image):
    """utils_bbox.py:8-33: undo the letterbox and scale to the original image; returns (y1, x1, y2, x2).
    Host-side numpy exactly like the reference (it runs after the device->host copy there too, :481-483)."""
    yx = np.asarray(box_xy)[..., ::-1]
    hw = np.asarray(box_wh)[..., ::-1]
    inp = np.array(input_shape)
    img = np.array(image_shape)
    if letterbox_image:
        new_shape = np.round(img * np.min(inp / img))
        offset = (inp - new_shape) / 2.0 / inp
        scale = inp / new_shape
        yx = (yx - offset) * scale
        hw = (hw * scale).astype(hw.dtype)  # in-place `*=` in the reference keeps the input dtype (:25)
    mins = yx - hw / 2.0
    maxes = yx + hw / 2.0
    out = np.concatenate([mins[..., 0:1], mins[..., 1:2], maxes[..., 0:1], maxes[..., 1:2]], axis=-1)
    return out * np.concatenate([img, img], axis=-1)


class DeviceNMS:
    """Score filter + class-aware NMS for a fixed (batch, anchors, classes); no host synchronisation.
    det: [B, max_det, 7] rows (x1, y1, x2, y2, obj_conf, class_conf, class_pred) sorted by score,
    count: [B] int32, keep_index: [B, max_det] int32 anchor indices."""

    def __init__(self, batch: int, anchors: int, num_classes: int, max_det: Optional[int] = None, device=None):
        self._lib = N.load()
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.batch, self.anchors, self.nc = batch, anchors, num_classes
        self.max_det = anchors if max_det is None else int(max_det)
        nbytes = self._lib.glsdet_nms_workspace_bytes(batch, anchors, num_classes)
        if nbytes <= 0:
            raise N.NativeError("glsdet_nms_workspace_bytes rejected the problem size")
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.device = dev
        self.det = self.count = self.keep_index = None   # results of the most recent launch
        self.handle = C.c_void_p()
        N.check(self._lib.glsdet_nms_create(batch, anchors, num_classes, self.max_det, self.workspace.data_ptr(),
                                            nbytes, C.byref(self.handle)), "glsdet_nms_create")

    def launch(self, pred: torch.Tensor, conf_thres: float, nms_thres: float, strategy: str = "auto_cuda",
               stream=None, cls_logits: bool = False, out=None, box_div: Optional[torch.Tensor] = None):
        """`pred`: [B, A, 5+nc] fp32, either contiguous rows or the permuted view of a [B, 5+nc, A] tensor (what the
        reference's decode_outputs returns and what the fused path writes); anything else is made contiguous.
        Every launch returns FRESH (det, count) tensors (caching-allocator blocks, no synchronisation), so a caller may keep
        step i's detections across step i+1; rows beyond count[b] are unspecified.  `out=(det, count, keep_index)` writes
        into caller-provided tensors instead (e.g. double buffers of a pipelined loop).  `box_div`: optional [B, 4] fp32
        divisors of the corner boxes (mmdet's rescale, yolox_head.py:283-285)."""
        assert pred.dtype == torch.float32 and pred.is_cuda
        nch = 5 + self.nc
        assert tuple(pred.shape) == (self.batch, self.anchors, nch), (pred.shape, self.batch, self.anchors)
        if pred.is_contiguous():
            layout = N.PRED_ROWS
        elif tuple(pred.stride()) == (nch * self.anchors, 1, self.anchors):
            layout = N.PRED_PLANES
        else:
            pred, layout = pred.contiguous(), N.PRED_ROWS
        if cls_logits:   # class columns are raw logits (FFAPathPlan detect mode): the filter applies the sigmoid
            layout |= N.PRED_CLS_LOGITS
        if out is None:
            self.det = torch.empty((self.batch, self.max_det, 7), dtype=torch.float32, device=self.device)
            self.count = torch.empty((self.batch,), dtype=torch.int32, device=self.device)
            self.keep_index = torch.empty((self.batch, self.max_det), dtype=torch.int32, device=self.device)
        else:
            self.det, self.count, self.keep_index = out
            assert tuple(self.det.shape) == (self.batch, self.max_det, 7) and self.det.dtype == torch.float32 and self.det.is_contiguous()
            assert tuple(self.count.shape) == (self.batch,) and self.count.dtype == torch.int32
            assert tuple(self.keep_index.shape) == (self.batch, self.max_det) and self.keep_index.dtype == torch.int32
        if box_div is not None:
            assert box_div.dtype == torch.float32 and tuple(box_div.shape) == (self.batch, 4) and box_div.is_contiguous()
        N.check(self._lib.glsdet_nms_launch_layout(self.handle, pred.data_ptr(), layout, N.ptr(box_div), float(conf_thres),
                                                   float(nms_thres), STRATEGIES[strategy], self.det.data_ptr(),
                                                   self.count.data_ptr(), self.keep_index.data_ptr(),
                                                   N.stream_ptr(stream)), "glsdet_nms_launch_layout")
        return self.det, self.count

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self._lib.glsdet_nms_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


_nms_cache = {}


def _device_nms(batch, anchors, nc, device) -> DeviceNMS:
    key = (batch, anchors, nc, str(device))
    if key not in _nms_cache:
        if len(_nms_cache) > 8:
            _nms_cache.clear()
        _nms_cache[key] = DeviceNMS(batch, anchors, nc, device=device)
    return _nms_cache[key]


def non_max_suppression(prediction: torch.Tensor, num_classes: int, input_shape, image_shape, letterbox_image,
                        conf_thres: float = 0.5, nms_thres: float = 0.4, strategy: str = "auto_cuda",
                        ) -> List[Optional[np.ndarray]]:
    """utils_bbox.py:375-484.  Returns, per image, a float32 ndarray [K, 7] with rows
    (y1, x1, y2, x2 in original-image pixels, obj_conf, class_conf, class_pred) sorted by score, or None for
    an input without anchors (:405-406)."""
    b, a = prediction.shape[:2]
    if a == 0:
        return [None for _ in range(b)]
    if not prediction.is_cuda:
        raise N.NativeError("non_max_suppression needs a CUDA tensor (glsdet_b200 has no CPU path)")
    pred = prediction.float()
    if pred.shape[-1] != 5 + num_classes:
        pred = pred[..., :5 + num_classes].contiguous()
    op = _device_nms(b, a, num_classes, pred.device)
    det, count = op.launch(pred, conf_thres, nms_thres, strategy)
    counts = count.cpu().numpy()                 # the one device->host sync of the reference (:481)
    kmax = int(counts.max()) if b else 0
    rows = det[:, :kmax].cpu().numpy()
    output: List[Optional[np.ndarray]] = []
    for i in range(b):
        out = rows[i, :counts[i]].copy()
        box_xy, box_wh = (out[:, 0:2] + out[:, 2:4]) / 2, out[:, 2:4] - out[:, 0:2]
        out[:, :4] = yolo_correct_boxes(box_xy, box_wh, input_shape, image_shape, letterbox_image)
        output.append(out)
    return output


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float,
                strategy: str = "auto_cuda") -> torch.Tensor:
    """Same contract as torchvision.ops.boxes.batched_nms (the call at utils_bbox.py:414-419): int64 indices of
    the kept boxes, sorted by decreasing score."""
    lib = N.load()
    k = boxes.shape[0]
    if k == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    if not boxes.is_cuda:
        raise N.NativeError("batched_nms needs CUDA tensors (glsdet_b200 has no CPU path)")
    bx = boxes.float().contiguous()
    sc = scores.float().contiguous()
    lb = idxs.float().contiguous()
    # any idxs (negative, batch * nc + cls, > 255): dense ids 0..U-1 pick the class segments, the original values keep
    # driving the coordinate-trick offsets exactly like torchvision's idxs * (max + 1)
    uniq, ids = torch.unique(idxs, return_inverse=True)
    if uniq.numel() > 256:
        raise NotImplementedError(f"batched_nms: {uniq.numel()} distinct idxs (at most 256 class segments are supported)")
    ids = ids.to(torch.int32).contiguous()
    uf = uniq.float()
    lmax = float(uf.abs().max()) if bool((uf == uf.round()).all()) else -1.0
    nbytes, ws = _batched_nms_workspace(k, boxes.device)
    keep = torch.empty((k,), dtype=torch.int32, device=boxes.device)
    cnt = torch.zeros((1,), dtype=torch.int32, device=boxes.device)
    N.check(lib.glsdet_batched_nms_ids(bx.data_ptr(), sc.data_ptr(), lb.data_ptr(), ids.data_ptr(), lmax, k, float(iou_threshold),
                                       STRATEGIES[strategy], ws.data_ptr(), nbytes, keep.data_ptr(), cnt.data_ptr(),
                                       N.stream_ptr()), "glsdet_batched_nms_ids")
    return keep[:int(cnt.item())].long()


_bnms_ws = {}


def _batched_nms_workspace(k: int, device):
    """Workspace of batched_nms, cached per device and grown geometrically (it was allocated on every call)."""
    lib = N.load()
    key = str(device)
    cap, ws = _bnms_ws.get(key, (0, None))
    if k > cap:
        cap = max(k, 2 * cap, 1024)
        ws = torch.empty(lib.glsdet_batched_nms_workspace_bytes(cap), dtype=torch.uint8, device=device)
        _bnms_ws[key] = (cap, ws)
    return ws.numel(), ws


def detection_lines(result: Optional[np.ndarray], class_names: Sequence[str], keep_classes: Optional[Sequence[str]] = None) -> List[str]:
    """The detection-results wire format of yolox-drone/yolo.py:288-303 (get_map_txt): one line
    "<class> <score[:6]> <left> <top> <right> <bottom>" per row of a non_max_suppression result (rows are
    (top, left, bottom, right, obj_conf, class_conf, class_pred)); score = str(obj_conf * class_conf)[:6] of the numpy
    float32 product, coordinates truncated with int().  `keep_classes` mirrors the `class_names` filter of :298-299."""
    if result is None:
        return []
    lines = []
    top_label = np.array(result[:, 6], dtype="int32")
    top_conf = result[:, 4] * result[:, 5]
    top_boxes = result[:, :4]
    for i, c in enumerate(top_label):
        predicted_class = class_names[int(c)]
        if keep_classes is not None and predicted_class not in keep_classes:
            continue
        top, left, bottom, right = top_boxes[i]
        score = str(top_conf[i])
        lines.append("%s %s %s %s %s %s" % (predicted_class, score[:6], str(int(left)), str(int(top)), str(int(right)),
                                            str(int(bottom))))
    return lines
