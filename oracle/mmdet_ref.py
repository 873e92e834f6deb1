"""ORACLE (test infrastructure only) - CPU fp32 restatement of the mmdet 2.19.1 YOLOX neck/head inference path.

mmdet's modules cannot be imported here or on the GPU box (mmcv-full is not installed and
mmdet/models/detectors/__init__.py is missing from the checkout), so this file restates them from the
vendored sources under /root/reference/yolox-ufp/mmdet (cited per function) plus the published semantics of the
mmcv 1.x pieces they call (mmcv-full>=1.3.17,<=1.5.0, requirements/mminstall.txt:1):
  * mmcv.cnn.ConvModule      = conv (bias only without norm) -> norm ('bn') -> activation; Swish = x * sigmoid(x)
  * mmcv.ops.nms.batched_nms = boxes + label * (max + 1); one nms when K < split_thr (10000) else one per class on the
                               shifted boxes; result sorted by score; returns (cat(boxes, scores)[keep], keep)
  * mmcv.ops.nms.nms         = greedy NMS, suppress iff IoU > thr, offset 0 (same arithmetic as torchvision's)
PARITY UNPINNED for the mmcv pieces (no reference test pins them, the library is absent).  The conv/decode math is
pinned indirectly: with weights renamed by the key map of SURVEY.md section 8c this path must reproduce the outputs of
yolox-drone/models/base/yolox.py, whose real outputs are committed as tests/golden/stock_s_calibrated.npz.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import nms_oracle
from .ref_path import base_conv

StateDict = Dict[str, torch.Tensor]


def drone_to_mmdet_keys(sd: StateDict) -> (StateDict, StateDict):
    """yolox-drone base/yolox.py state_dict -> (neck, head) state_dicts in mmdet naming."""
    neck_map = (("backbone.lateral_conv0.", "reduce_layers.0."), ("backbone.reduce_conv1.", "reduce_layers.1."),
                ("backbone.C3_p4.", "top_down_blocks.0."), ("backbone.C3_p3.", "top_down_blocks.1."),
                ("backbone.bu_conv2.", "downsamples.0."), ("backbone.bu_conv1.", "downsamples.1."),
                ("backbone.C3_n3.", "bottom_up_blocks.0."), ("backbone.C3_n4.", "bottom_up_blocks.1."),
                ("head.stems.", "out_convs."))
    csp = ((".conv1.", ".main_conv."), (".conv2.", ".short_conv."), (".conv3.", ".final_conv."))
    head_map = (("head.cls_convs.", "multi_level_cls_convs."), ("head.reg_convs.", "multi_level_reg_convs."),
                ("head.cls_preds.", "multi_level_conv_cls."), ("head.reg_preds.", "multi_level_conv_reg."),
                ("head.obj_preds.", "multi_level_conv_obj."))
    neck, head = {}, {}
    for k, v in sd.items():
        if k.startswith("backbone.backbone."):
            continue
        for a, b in head_map:
            if k.startswith(a):
                head[b + k[len(a):]] = v
                break
        else:
            for a, b in neck_map:
                if k.startswith(a):
                    k2 = b + k[len(a):]
                    if "blocks." in b and ".m." in k2:
                        k2 = k2.replace(".m.", ".blocks.")          # bottleneck convs keep conv1 / conv2
                    elif "blocks." in b:
                        for x, y in csp:
                            k2 = k2.replace(x, y)
                    neck[k2] = v
                    break
    return neck, head


def _csp(sd: StateDict, p: str, x: torch.Tensor) -> torch.Tensor:
    """mmdet/models/utils/csp_layer.py:137-150 with DarknetBottleneck :63-72 (add_identity=False in the neck)."""
    x_short = base_conv(sd, p + ".short_conv", x)
    x_main = base_conv(sd, p + ".main_conv", x)
    j = 0
    while f"{p}.blocks.{j}.conv1.conv.weight" in sd:
        x_main = base_conv(sd, f"{p}.blocks.{j}.conv2", base_conv(sd, f"{p}.blocks.{j}.conv1", x_main))
        j += 1
    return base_conv(sd, p + ".final_conv", torch.cat((x_main, x_short), dim=1))


def yolox_pafpn(sd: StateDict, inputs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """mmdet/models/necks/yolox_pafpn.py:117-156."""
    n = len(inputs)
    with torch.no_grad():
        inner = [inputs[-1]]
        for idx in range(n - 1, 0, -1):
            hi = base_conv(sd, f"reduce_layers.{n - 1 - idx}", inner[0])
            inner[0] = hi
            up = F.interpolate(hi, scale_factor=2, mode="nearest")
            inner.insert(0, _csp(sd, f"top_down_blocks.{n - 1 - idx}", torch.cat([up, inputs[idx - 1]], 1)))
        outs = [inner[0]]
        for idx in range(n - 1):
            down = base_conv(sd, f"downsamples.{idx}", outs[-1], stride=2)
            outs.append(_csp(sd, f"bottom_up_blocks.{idx}", torch.cat([down, inner[idx + 1]], 1)))
        return [base_conv(sd, f"out_convs.{i}", o) for i, o in enumerate(outs)]


def yolox_head_forward(sd: StateDict, feats: Sequence[torch.Tensor]):
    """mmdet/models/dense_heads/yolox_head.py:184-213."""
    cls_scores, bbox_preds, objs = [], [], []
    with torch.no_grad():
        for l, x in enumerate(feats):
            cf = base_conv(sd, f"multi_level_cls_convs.{l}.1", base_conv(sd, f"multi_level_cls_convs.{l}.0", x))
            rf = base_conv(sd, f"multi_level_reg_convs.{l}.1", base_conv(sd, f"multi_level_reg_convs.{l}.0", x))
            cls_scores.append(F.conv2d(cf, sd[f"multi_level_conv_cls.{l}.weight"], sd[f"multi_level_conv_cls.{l}.bias"]))
            bbox_preds.append(F.conv2d(rf, sd[f"multi_level_conv_reg.{l}.weight"], sd[f"multi_level_conv_reg.{l}.bias"]))
            objs.append(F.conv2d(rf, sd[f"multi_level_conv_obj.{l}.weight"], sd[f"multi_level_conv_obj.{l}.bias"]))
    return cls_scores, bbox_preds, objs


def mmcv_batched_nms(boxes: np.ndarray, scores: np.ndarray, idxs: np.ndarray, iou_threshold: float,
                     split_thr: int = 10000):
    """mmcv.ops.nms.batched_nms (class_agnostic=False).  Returns (dets [n,5], keep)."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float32).reshape(-1, 4)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    idxs = np.asarray(idxs)
    if boxes.shape[0] == 0:
        return np.zeros((0, 5), np.float32), np.zeros((0,), np.int64)
    max_coordinate = boxes.max()
    offsets = (idxs.astype(np.float32) * np.float32(max_coordinate + np.float32(1))).astype(np.float32)
    shifted = (boxes + offsets[:, None]).astype(np.float32)
    if shifted.shape[0] < split_thr:
        keep = nms_oracle.nms(shifted, scores, iou_threshold)
    else:
        mask = np.zeros(len(scores), dtype=bool)
        for c in np.unique(idxs):
            cur = np.nonzero(idxs == c)[0]
            mask[cur[nms_oracle.nms(shifted[cur], scores[cur], iou_threshold)]] = True
        keep = np.nonzero(mask)[0]
        keep = keep[np.argsort(-scores[keep], kind="stable")]
    return np.concatenate([boxes[keep], scores[keep, None]], axis=1), keep


def get_bboxes(cls_scores, bbox_preds, objectnesses, strides, score_thr: float, iou_threshold: float,
               scale_factors=None):
    """mmdet/models/dense_heads/yolox_head.py:215-322 (priors: core/anchor/point_generator.py:148-175, offset 0).
    Returns per image (dets [n,5] float32, labels [n] int64)."""
    num_imgs = cls_scores[0].shape[0]
    nc = cls_scores[0].shape[1]
    priors = []
    for (h, w), s in zip([c.shape[2:] for c in cls_scores], strides):
        sx = (torch.arange(0, w) * s).float()
        sy = (torch.arange(0, h) * s).float()
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        priors.append(torch.stack([xx.reshape(-1), yy.reshape(-1), torch.full((h * w,), float(s)),
                                   torch.full((h * w,), float(s))], dim=-1))
    priors = torch.cat(priors)
    cls = torch.cat([c.permute(0, 2, 3, 1).reshape(num_imgs, -1, nc) for c in cls_scores], dim=1).sigmoid()
    box = torch.cat([b.permute(0, 2, 3, 1).reshape(num_imgs, -1, 4) for b in bbox_preds], dim=1)
    obj = torch.cat([o.permute(0, 2, 3, 1).reshape(num_imgs, -1) for o in objectnesses], dim=1).sigmoid()
    xys = (box[..., :2] * priors[:, 2:]) + priors[:, :2]                      # :299
    whs = box[..., 2:].exp() * priors[:, 2:]                                  # :300
    bboxes = torch.stack([xys[..., 0] - whs[..., 0] / 2, xys[..., 1] - whs[..., 1] / 2,
                          xys[..., 0] + whs[..., 0] / 2, xys[..., 1] + whs[..., 1] / 2], -1)
    if scale_factors is not None:
        bboxes = bboxes / torch.tensor(scale_factors, dtype=bboxes.dtype).unsqueeze(1)   # :283-285
    out = []
    for i in range(num_imgs):
        max_scores, labels = torch.max(cls[i], 1)                             # :311
        valid = obj[i] * max_scores >= score_thr                              # :312
        b = bboxes[i][valid].numpy()
        s = (max_scores[valid] * obj[i][valid]).numpy()
        l = labels[valid].numpy()
        if l.size == 0:
            out.append((b.reshape(0, 4), l))
            continue
        dets, keep = mmcv_batched_nms(b, s, l, iou_threshold)
        out.append((dets, l[keep]))
    return out


def expected_keys(in_channels, out_channels, num_csp_blocks, num_classes, feat_channels, levels=3):
    """State-dict key order the mmdet modules produce (derived from the constructors' registration order:
    yolox_pafpn.py:52-115, csp_layer.py:107-135, yolox_head.py:130-170); BN contributes weight, bias, running_mean,
    running_var, num_batches_tracked."""
    def cm(p):
        return [p + ".conv.weight"] + [f"{p}.bn.{n}" for n in ("weight", "bias", "running_mean", "running_var",
                                                              "num_batches_tracked")]

    def csp(p):
        k = cm(p + ".main_conv") + cm(p + ".short_conv") + cm(p + ".final_conv")
        for j in range(num_csp_blocks):
            k += cm(f"{p}.blocks.{j}.conv1") + cm(f"{p}.blocks.{j}.conv2")
        return k

    neck = []
    for i in range(2):
        neck += cm(f"reduce_layers.{i}")
    for i in range(2):
        neck += csp(f"top_down_blocks.{i}")
    for i in range(2):
        neck += cm(f"downsamples.{i}")
    for i in range(2):
        neck += csp(f"bottom_up_blocks.{i}")
    for i in range(3):
        neck += cm(f"out_convs.{i}")
    head = []
    for name in ("multi_level_cls_convs", "multi_level_reg_convs"):
        for l in range(levels):
            head += cm(f"{name}.{l}.0") + cm(f"{name}.{l}.1")
    for name in ("multi_level_conv_cls", "multi_level_conv_reg", "multi_level_conv_obj"):
        for l in range(levels):
            head += [f"{name}.{l}.weight", f"{name}.{l}.bias"]
    return neck, head
