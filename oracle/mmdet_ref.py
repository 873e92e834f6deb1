"""ORACLE (test infrastructure only) - CPU fp32 restatement of the mmdet 2.19.1 YOLOX neck/head inference path.

mmdet's modules cannot be imported here or on the GPU box (mmcv-full is not installed and
mmdet/models/detectors/__init__.py is missing from the checkout), so this file restates them from the
vendored sources under /root/reference/yolox-ufp/mmdet (cited per function) plus the published semantics of the
mmcv 1.x pieces they call (mmcv-full>=1.3.17,<=1.5.0, requirements/mminstall.txt:1):
  * mmcv.cnn.ConvModule      = conv (bias only without norm) -> norm ('bn') -> activation; Swish = x * sigmoid(x)
  * mmcv.ops.nms.batched_nms = boxes + label * (max + 1); one nms when K < split_thr (10000) else one per class on the
                               shifted boxes; result sorted by score; returns (cat(boxes, scores)[keep], keep)
  * mmcv.ops.nms.nms         = greedy NMS, suppress iff IoU > thr, offset 0 (same arithmetic as torchvision's)
PARITY UNPINNED for the mmcv pieces only (no reference test pins them, the library is absent).  What IS pinned:
  * MP-Det (fpn_forward, mp_forward_proxy, mp_head_forward, gfl_decode_level, gfl_get_bboxes_single): against
    tests/golden/mpdet_cases.npz, recorded by tests/golden/make_golden_mpdet.py, which EXECUTES the reference's own source
    of FPN.forward, MPHead.forward_single / forward_proxy, Integral, GFLHead._get_bboxes_single / anchor_center,
    BaseDenseHead._bbox_post_process, filter_scores_and_topk, distance2bbox and DistancePointBBoxCoder.decode (compiled
    from the files under /root/reference/yolox-ufp through `ast`, bound to attribute-only stand-in objects);
  * YOLOX neck / head conv and decode math: with weights renamed by the key map of SURVEY.md section 8c this path must
    reproduce the outputs of yolox-drone/models/base/yolox.py, whose real outputs are tests/golden/stock_s_calibrated.npz.
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
        # phi = 'nano' (DWConv, baseConv.py:22-30) <-> mmcv DepthwiseSeparableConvModule (use_depthwise=True)
        k = k.replace(".dconv.", ".depthwise_conv.").replace(".pconv.", ".pointwise_conv.")
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


# ------------------------------------------------------------------------------------------ MP-Det (BASELINE configs[2])
# FPN + MPHead inference path of yolox-ufp (necks/fpn.py, dense_heads/mp_head.py, dense_heads/gfl_head.py).  PARITY
# UNPINNED like the rest of this file: mmcv is absent and the MP-Det config file is missing from the checkout (SURVEY.md
# D6), so the configuration is the reconstruction of SURVEY.md section 8d row 3: FPN(in_channels=[256,512,1024,2048],
# out_channels=256, start_level=1, add_extra_convs='on_output', num_outs=5), MPHead(num_classes=10, in_channels=256,
# stacked_convs=4, feat_channels=256, reg_max=16, GN32, strides 8..128, proxies_list, gamma=10), AnchorGenerator with
# center_offset 0 (cell (x, y) -> point (x*s, y*s)), test_cfg(score_thr=0.05, nms_pre=1000, nms iou 0.6, max_per_img=500).
from . import ref_path as _rp

MP_PROXIES = (2, 3, 2, 5, 4, 8, 8, 4, 3, 3)   # mp_head.py:31
MP_STRIDES = (8, 16, 32, 64, 128)


def fpn_forward(sd: StateDict, inputs: Sequence[torch.Tensor], p: str = "", start_level: int = 1, num_outs: int = 5):
    """necks/fpn.py:152-203 with add_extra_convs='on_output', no norm, no activation (ConvModule = conv + bias)."""
    q = _rp._q
    n_lat = len(inputs) - start_level
    lat = [q(F.conv2d(q(inputs[i + start_level]), q(sd[f"{p}lateral_convs.{i}.conv.weight"]),
                      sd[f"{p}lateral_convs.{i}.conv.bias"])) for i in range(n_lat)]
    for i in range(n_lat - 1, 0, -1):                                                   # :166-175
        lat[i - 1] = q(lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest"))
    outs = [q(F.conv2d(lat[i], q(sd[f"{p}fpn_convs.{i}.conv.weight"]), sd[f"{p}fpn_convs.{i}.conv.bias"], padding=1))
            for i in range(n_lat)]                                                      # :179-181
    for i in range(n_lat, num_outs):                                                    # :196-202 ('on_output')
        outs.append(q(F.conv2d(outs[-1], q(sd[f"{p}fpn_convs.{i}.conv.weight"]), sd[f"{p}fpn_convs.{i}.conv.bias"],
                               stride=2, padding=1)))
    return outs


def _conv_gn_relu(sd, p, x):
    """mmcv ConvModule(conv without bias -> GroupNorm(32) -> ReLU), mp_head.py:46-62."""
    q = _rp._q
    y = q(F.conv2d(x, q(sd[p + ".conv.weight"]), None, padding=1))
    return q(torch.relu(F.group_norm(y, 32, sd[p + ".gn.weight"], sd[p + ".gn.bias"], eps=1e-5)))


def mp_forward_proxy(feat: torch.Tensor, proxies: torch.Tensor, proxies_list=MP_PROXIES, gamma: float = 10.0):
    """mp_head.py:105-121: cosine similarity to the class proxies, softmax(gamma * sim)-weighted sum per class, * gamma."""
    centers = F.normalize(proxies, p=2, dim=1)
    feat = F.normalize(feat, p=2, dim=1)
    sim = feat.matmul(centers.t())
    out, pre = [], 0
    for n in proxies_list:
        sub = sim[:, pre:pre + n]
        prob = F.softmax(sub * gamma, dim=1)
        out.append(torch.sum(prob * sub, dim=1)[:, None])
        pre += n
    return torch.cat(out, dim=1) * gamma


def mp_head_forward(sd: StateDict, feats: Sequence[torch.Tensor], p: str = "", stacked_convs: int = 4,
                    proxies_list=MP_PROXIES, gamma: float = 10.0):
    """MPHead.forward_single per level (mp_head.py:123-154): shared towers, gfl_reg * Scale, proxy classification.
    Returns (cls_scores [B, nc, H, W], bbox_preds [B, 4*(reg_max+1), H, W]) lists."""
    q = _rp._q
    cls_scores, bbox_preds = [], []
    for l, x in enumerate(feats):
        cf = rf = q(x)
        for i in range(stacked_convs):
            cf = _conv_gn_relu(sd, f"{p}cls_convs.{i}", cf)
            rf = _conv_gn_relu(sd, f"{p}reg_convs.{i}", rf)
        bbox = F.conv2d(rf, q(sd[p + "gfl_reg.weight"]), sd[p + "gfl_reg.bias"], padding=1) * sd[f"{p}scales.{l}.scale"]
        f = F.conv2d(cf, q(sd[p + "gfl_cls_conv.weight"]), sd[p + "gfl_cls_conv.bias"], padding=1)
        b, c, h, w = f.shape
        s = mp_forward_proxy(f.permute(0, 2, 3, 1).reshape(-1, c), sd[p + "proxies"], proxies_list, gamma)
        cls_scores.append(s.reshape(b, h, w, -1).permute(0, 3, 1, 2).contiguous())
        bbox_preds.append(bbox.float())
    return cls_scores, bbox_preds


def gfl_decode_level(bbox_pred: torch.Tensor, stride: int, img_shape, reg_max: int = 16) -> torch.Tensor:
    """Integral (gfl_head.py:35-49) * stride and DistancePointBBoxCoder.decode on the cell points (x*s, y*s)
    (gfl_head.py:437-438,456-457; core/bbox/transforms.py:136-165): [4*(reg_max+1), H, W] -> [H*W, 4] xyxy clamped."""
    _, h, w = bbox_pred.shape
    x = bbox_pred.permute(1, 2, 0).reshape(-1, reg_max + 1)
    d = F.linear(F.softmax(x, dim=1), torch.linspace(0, reg_max, reg_max + 1)[None]).reshape(-1, 4) * stride
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    px, py = (xs.reshape(-1) * stride).float(), (ys.reshape(-1) * stride).float()
    boxes = torch.stack([px - d[:, 0], py - d[:, 1], px + d[:, 2], py + d[:, 3]], -1)
    boxes[:, 0::2] = boxes[:, 0::2].clamp(min=0, max=img_shape[1])
    boxes[:, 1::2] = boxes[:, 1::2].clamp(min=0, max=img_shape[0])
    return boxes


def gfl_get_bboxes_single(cls_scores, bbox_preds, img_shape, score_thr=0.05, nms_pre=1000, iou_thr=0.6, max_per_img=500,
                          strides=MP_STRIDES, scale_factor=None):
    """_get_bboxes_single (gfl_head.py:426-471) + filter_scores_and_topk (core/utils/misc.py:143-165) +
    _bbox_post_process (base_dense_head.py:276-301) for ONE image: per-level maps [nc, H, W] / [68, H, W].
    Ties in the per-level sort are broken by (anchor, class) order (torch.sort is not stable; documented choice)."""
    mb, ms, ml = [], [], []
    for cs, bp, s in zip(cls_scores, bbox_preds, strides):
        nc = cs.shape[0]
        scores = cs.permute(1, 2, 0).reshape(-1, nc).sigmoid()
        boxes = gfl_decode_level(bp, s, img_shape)
        flat = scores.reshape(-1)
        idx = torch.nonzero(flat > score_thr).reshape(-1)
        order = torch.sort(flat[idx], descending=True, stable=True)[1][:nms_pre]
        idx = idx[order]
        ms.append(flat[idx]); ml.append(idx % nc); mb.append(boxes[idx // nc])
    boxes, scores, labels = torch.cat(mb), torch.cat(ms), torch.cat(ml)
    if scale_factor is not None:      # rescale=True: mlvl_bboxes /= mlvl_bboxes.new_tensor(scale_factor) (base_dense_head.py:282-283)
        boxes = boxes / boxes.new_tensor(scale_factor)
    if boxes.numel() == 0:
        return torch.zeros((0, 5)), labels
    _, keep = mmcv_batched_nms(boxes.numpy(), scores.numpy(), labels.float().numpy(), iou_thr)
    keep = torch.from_numpy(np.asarray(keep, dtype=np.int64))[:max_per_img]
    return torch.cat([boxes[keep], scores[keep][:, None]], 1), labels[keep]


def mpdet_state_dict_shapes(num_classes: int = 10, proxies_list=MP_PROXIES, reg_max: int = 16, feat: int = 256,
                            in_channels=(256, 512, 1024, 2048), num_words: int = 200):
    """Keys / shapes of FPN ('neck.') and MPHead ('bbox_head.') parameters and buffers (mp_head.py:42-98)."""
    sh = {}
    for i, c in enumerate(in_channels[1:]):
        sh[f"neck.lateral_convs.{i}.conv.weight"], sh[f"neck.lateral_convs.{i}.conv.bias"] = (feat, c, 1, 1), (feat,)
    for i in range(5):
        sh[f"neck.fpn_convs.{i}.conv.weight"], sh[f"neck.fpn_convs.{i}.conv.bias"] = (feat, feat, 3, 3), (feat,)
    for br in ("cls_convs", "reg_convs"):
        for i in range(4):
            sh[f"bbox_head.{br}.{i}.conv.weight"] = (feat, feat, 3, 3)
            sh[f"bbox_head.{br}.{i}.gn.weight"], sh[f"bbox_head.{br}.{i}.gn.bias"] = (feat,), (feat,)
    sh["bbox_head.gfl_cls_conv.weight"], sh["bbox_head.gfl_cls_conv.bias"] = (feat, feat, 3, 3), (feat,)
    sh["bbox_head.gfl_reg.weight"], sh["bbox_head.gfl_reg.bias"] = (4 * (reg_max + 1), feat, 3, 3), (4 * (reg_max + 1),)
    for l in range(5):
        sh[f"bbox_head.scales.{l}.scale"] = ()
    sh["bbox_head.proxies"] = (sum(proxies_list), feat)
    sh["bbox_head._embedding"] = (num_classes + 1, num_words, feat)
    sh["bbox_head._pos_embedding_ptr"] = (num_classes + 1,)
    sh["bbox_head._proxies_prob"] = (sum(proxies_list),)
    sh["bbox_head.integral.project"] = (reg_max + 1,)
    return sh


def mpdet_synthetic_state_dict(seed: int = 0, num_classes: int = 10):
    """Seeded random weights with O(1) activations (variance-preserving convs, GN affine near identity, logit noise)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in mpdet_state_dict_shapes(num_classes).items():
        if k.endswith("conv.weight") or k.endswith("gfl_cls_conv.weight") or k.endswith("gfl_reg.weight"):
            fan = shp[1] * shp[2] * shp[3]
            gain = 1.0 if k.startswith("neck.") else (2.0 if "_convs." in k else 1.0)
            sd[k] = torch.randn(shp, generator=g) * (gain / fan) ** 0.5
        elif k.endswith("gn.weight"):
            sd[k] = 1.0 + 0.05 * torch.randn(shp, generator=g)
        elif k.endswith(".bias"):
            sd[k] = 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(".scale"):
            sd[k] = torch.tensor(1.0 + 0.1 * float(torch.randn((), generator=g)))
        elif k.endswith("proxies"):
            sd[k] = torch.randn(shp, generator=g)
        elif k.endswith("_embedding"):
            sd[k] = torch.randn(shp, generator=g)
        elif k.endswith("_pos_embedding_ptr"):
            sd[k] = torch.zeros(shp, dtype=torch.long)
        elif k.endswith("_proxies_prob"):
            sd[k] = torch.cat([torch.full((n,), 1.0 / n) for n in MP_PROXIES])
        elif k.endswith("integral.project"):
            sd[k] = torch.linspace(0, 16, 17)
        else:
            raise KeyError(k)
    return sd
