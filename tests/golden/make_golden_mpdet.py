"""Golden vectors of the MP-Det neck / head recorded by EXECUTING THE REFERENCE'S OWN
SOURCE for every function that lives under /root/reference/yolox-ufp:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_mpdet.py      (build container only)

`import mmdet` needs mmcv-full (absent), so the methods are compiled from the reference files through `ast` and bound to
small stand-in objects that only carry attributes (weights, config values):

  necks/fpn.py                 FPN.forward                      (:152-203, the @auto_fp16 decorator is a no-op here)
  dense_heads/mp_head.py       MPHead.forward_proxy, forward_single               (:105-154)
  dense_heads/gfl_head.py      class Integral, GFLHead._get_bboxes_single, anchor_center   (:16-49, :124-136, :391-471)
  dense_heads/base_dense_head.py  BaseDenseHead._bbox_post_process                (:226-301)
  core/utils/misc.py           filter_scores_and_topk           (:119-165)
  core/bbox/transforms.py      distance2bbox                    (:136-190)
  core/bbox/coder/distance_point_bbox_coder.py  DistancePointBBoxCoder.decode     (as a method of a stand-in)

What is NOT under /root/reference and therefore restated (mmcv 1.x semantics): ConvModule (conv -> GroupNorm(32) -> ReLU
for the towers, conv + bias for the FPN), Scale (x * scalar), AnchorGenerator priors of GFL (square anchors of 8 * stride
centred on (x * stride, y * stride)), and mmcv.ops.batched_nms (oracle.mmdet_ref.mmcv_batched_nms on top of the pinned NMS).
Weights: oracle.mmdet_ref.mpdet_synthetic_state_dict(0) (seeded); inputs: torch.randn with a seeded generator - both are
rebuilt by the tests, only the outputs are stored."""
import ast
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference/yolox-ufp/mmdet")
sys.dont_write_bytecode = True
sys.path.insert(0, str(ROOT))
from oracle import mmdet_ref as M  # noqa: E402


def extract(path, names, ns):
    """Compile the named top-level functions / classes, or `Class.method` methods, of a reference file into `ns`."""
    tree = ast.parse((REF / path).read_text())
    body = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            body.append(node)
        elif isinstance(node, ast.ClassDef):
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and f"{node.name}.{sub.name}" in names:
                    sub.decorator_list = []          # @auto_fp16 / @force_fp32: identity in fp32
                    body.append(sub)
    mod = ast.Module(body=body, type_ignores=[])
    exec(compile(mod, str(REF / path), "exec"), ns)
    return ns


class Cfg(dict):
    __getattr__ = dict.__getitem__


def main():
    ns = {"torch": torch, "nn": nn, "F": F, "np": np}
    extract("core/utils/misc.py", {"filter_scores_and_topk"}, ns)
    extract("core/bbox/transforms.py", {"distance2bbox"}, ns)
    extract("models/necks/fpn.py", {"FPN.forward"}, ns)
    fpn_forward = ns.pop("forward")
    extract("models/dense_heads/mp_head.py", {"MPHead.forward_proxy", "MPHead.forward_single"}, ns)
    forward_proxy, forward_single = ns.pop("forward_proxy"), ns.pop("forward_single")
    extract("models/dense_heads/gfl_head.py", {"Integral", "GFLHead._get_bboxes_single", "GFLHead.anchor_center"}, ns)
    get_bboxes_single, anchor_center = ns.pop("_get_bboxes_single"), ns.pop("anchor_center")
    extract("models/dense_heads/base_dense_head.py", {"BaseDenseHead._bbox_post_process"}, ns)
    bbox_post_process = ns.pop("_bbox_post_process")
    extract("core/bbox/coder/distance_point_bbox_coder.py", {"DistancePointBBoxCoder.decode"}, ns)
    coder_decode = ns.pop("decode")

    def batched_nms(boxes, scores, idxs, nms_cfg):     # mmcv.ops.batched_nms restated (third party, absent)
        dets, keep = M.mmcv_batched_nms(boxes.numpy(), scores.numpy(), idxs.numpy().astype(np.float32), float(nms_cfg["iou_threshold"]))
        return torch.from_numpy(dets), torch.from_numpy(np.asarray(keep, dtype=np.int64))
    ns["batched_nms"] = batched_nms

    sd = M.mpdet_synthetic_state_dict(0)
    nsd = {k[5:]: v for k, v in sd.items() if k.startswith("neck.")}
    hsd = {k[10:]: v for k, v in sd.items() if k.startswith("bbox_head.")}

    def conv(w, b, stride=1):
        k = w.shape[-1]
        return lambda x: F.conv2d(x, w, b, stride=stride, padding=(k - 1) // 2)

    def conv_gn_relu(p):    # mmcv ConvModule(norm_cfg=GN32): conv without bias -> GroupNorm -> ReLU
        return lambda x: F.relu(F.group_norm(F.conv2d(x, hsd[p + ".conv.weight"], None, padding=1), 32, hsd[p + ".gn.weight"], hsd[p + ".gn.bias"], 1e-5))

    out = {}
    for case, (H, W, seed) in {"small": (128, 192, 1), "odd": (200, 336, 2)}.items():
        g = torch.Generator().manual_seed(seed)
        ins = [torch.randn(1, c, -(-H // s), -(-W // s), generator=g) for c, s in zip((256, 512, 1024, 2048), (4, 8, 16, 32))]
        # ---- FPN.forward on a stand-in carrying the attributes it reads
        neck = types.SimpleNamespace(
            in_channels=[256, 512, 1024, 2048], start_level=1, backbone_end_level=4, num_outs=5, add_extra_convs="on_output",
            relu_before_extra_convs=False, upsample_cfg=dict(mode="nearest"),
            lateral_convs=[conv(nsd[f"lateral_convs.{i}.conv.weight"], nsd[f"lateral_convs.{i}.conv.bias"]) for i in range(3)],
            fpn_convs=[conv(nsd[f"fpn_convs.{i}.conv.weight"], nsd[f"fpn_convs.{i}.conv.bias"], 2 if i >= 3 else 1) for i in range(5)])
        with torch.no_grad():
            feats = fpn_forward(neck, ins)
        # ---- MPHead.forward_single per level
        head = types.SimpleNamespace(
            cls_convs=[conv_gn_relu(f"cls_convs.{i}") for i in range(4)], reg_convs=[conv_gn_relu(f"reg_convs.{i}") for i in range(4)],
            gfl_reg=conv(hsd["gfl_reg.weight"], hsd["gfl_reg.bias"]), gfl_cls_conv=conv(hsd["gfl_cls_conv.weight"], hsd["gfl_cls_conv.bias"]),
            training=False, feat_channels=256, proxies=hsd["proxies"], num_classes=10, proxies_list=list(M.MP_PROXIES), gamma=10)
        head.forward_proxy = types.MethodType(forward_proxy, head)
        cls_scores, bbox_preds = [], []
        with torch.no_grad():
            for l, x in enumerate(feats):
                s = hsd[f"scales.{l}.scale"]
                c, b = forward_single(head, x, lambda t, s=s: t * s)      # mmcv Scale
                cls_scores.append(c)
                bbox_preds.append(b)
        # ---- GFLHead._get_bboxes_single + BaseDenseHead._bbox_post_process
        img_shape = (H, W - 11, 3)
        strides = [(s, s) for s in M.MP_STRIDES]
        priors = []
        for (s, _), c in zip(strides, cls_scores):
            h, w = c.shape[2:]
            ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32) * s, torch.arange(w, dtype=torch.float32) * s, indexing="ij")
            cx, cy, half = xs.reshape(-1), ys.reshape(-1), 4.0 * s      # AnchorGenerator: scale 8, ratio 1, centre offset 0
            priors.append(torch.stack([cx - half, cy - half, cx + half, cy + half], 1))
        gfl = types.SimpleNamespace(test_cfg=None, prior_generator=types.SimpleNamespace(strides=strides), cls_out_channels=10,
                                    integral=ns["Integral"](16))
        gfl.bbox_coder = types.SimpleNamespace(clip_border=True)
        gfl.bbox_coder.decode = types.MethodType(coder_decode, gfl.bbox_coder)
        gfl.anchor_center = types.MethodType(anchor_center, gfl)
        gfl._bbox_post_process = types.MethodType(bbox_post_process, gfl)
        cfg = Cfg(nms_pre=1000, score_thr=0.05, nms=dict(type="nms", iou_threshold=0.6), max_per_img=500)
        with torch.no_grad():
            dets, labels = get_bboxes_single(gfl, [c[0] for c in cls_scores], [b[0] for b in bbox_preds], None, priors,
                                             dict(img_shape=img_shape, scale_factor=1.0), cfg, rescale=False, with_nms=True)
        out[f"{case}_meta"] = np.array([H, W, seed])
        fs, bs = (4, 1) if case == "small" else (16, 4)     # channel sub-sampling keeps the fixture small
        out[f"{case}_stride"] = np.array([fs, bs])
        for l in range(5):
            out[f"{case}_fpn{l}"] = feats[l][:, ::fs].numpy()
            out[f"{case}_cls{l}"] = cls_scores[l].numpy()
            out[f"{case}_box{l}"] = bbox_preds[l][:, ::bs].numpy()
        out[f"{case}_dets"] = dets.numpy()
        out[f"{case}_labels"] = labels.numpy()
        print(case, [tuple(f.shape) for f in feats], "dets", tuple(dets.shape))
    np.savez_compressed(HERE / "mpdet_cases.npz", **out)
    print((HERE / "mpdet_cases.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
