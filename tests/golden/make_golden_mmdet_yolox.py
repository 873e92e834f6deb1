"""Golden vectors of the mmdet YOLOX face recorded by EXECUTING THE REFERENCE'S OWN SOURCE:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_mmdet_yolox.py      (build container only)

`import mmdet` needs mmcv-full (absent), so the methods below are compiled from the files under
/root/reference/yolox-ufp/mmdet through `ast` and bound to attribute-only stand-in objects:

  models/necks/yolox_pafpn.py       YOLOXPAFPN.forward                                   (:117-156)
  models/utils/csp_layer.py         DarknetBottleneck.forward, CSPLayer.forward          (:63-72, :142-150)
  models/dense_heads/yolox_head.py  YOLOXHead.forward_single, get_bboxes, _bbox_decode, _bboxes_nms   (:184-195, :215-322)
  core/anchor/point_generator.py    class MlvlPointGenerator                             (grid_priors with_stride)

Restated because they live in mmcv (absent): ConvModule = conv (no bias) -> BatchNorm(eps 1e-3, eval) -> Swish, nn.Upsample
(nearest, x2), mmcv.ops.batched_nms (oracle.mmdet_ref.mmcv_batched_nms).  Weights: the seeded stock-YOLOX state dict of
tests/golden/meta.json renamed by oracle.mmdet_ref.drone_to_mmdet_keys; inputs: seeded torch.randn (rebuilt by the tests)."""
import ast
import json
import sys
import types
from functools import partial
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.modules.utils import _pair

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference/yolox-ufp/mmdet")
sys.dont_write_bytecode = True
sys.path.insert(0, str(ROOT))
from oracle import mmdet_ref as M, ref_path  # noqa: E402


def extract(path, names, ns):
    tree = ast.parse((REF / path).read_text())
    body = []
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in names:
            node.decorator_list = []
            node.bases = []
            body.append(node)
        elif isinstance(node, ast.ClassDef):
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and f"{node.name}.{sub.name}" in names:
                    sub.decorator_list = []
                    sub.name = f"{node.name}_{sub.name}"
                    body.append(sub)
    exec(compile(ast.Module(body=body, type_ignores=[]), str(REF / path), "exec"), ns)
    return ns


class Cfg(dict):
    __getattr__ = dict.__getitem__


def multi_apply(func, *args, **kwargs):      # mmdet/core/utils/misc.py:8-28 (three lines, restated)
    pfunc = partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


def main():
    meta = json.loads((HERE / "meta.json").read_text())["stock"]
    sd = ref_path.synthetic_state_dict(meta["nc"], meta["phi"], seed=meta["seed"], flavour="calibrated", variant="stock")
    nsd, hsd = M.drone_to_mmdet_keys(sd)
    ns = {"torch": torch, "nn": nn, "F": F, "np": np, "_pair": _pair, "multi_apply": multi_apply}
    extract("models/necks/yolox_pafpn.py", {"YOLOXPAFPN.forward"}, ns)
    extract("models/utils/csp_layer.py", {"DarknetBottleneck.forward", "CSPLayer.forward"}, ns)
    extract("models/dense_heads/yolox_head.py", {"YOLOXHead.forward_single", "YOLOXHead.forward", "YOLOXHead.get_bboxes",
                                                 "YOLOXHead._bbox_decode", "YOLOXHead._bboxes_nms"}, ns)
    extract("core/anchor/point_generator.py", {"MlvlPointGenerator"}, ns)

    def batched_nms(boxes, scores, idxs, nms_cfg):
        dets, keep = M.mmcv_batched_nms(boxes.numpy(), scores.numpy(), idxs.numpy().astype(np.float32), float(nms_cfg["iou_threshold"]))
        return torch.from_numpy(dets), torch.from_numpy(np.asarray(keep, dtype=np.int64))
    ns["batched_nms"] = batched_nms

    def conv_module(d, p, stride=1):     # mmcv ConvModule(norm=BN eps 1e-3 momentum 0.03, act=Swish), eval mode
        w = d[p + ".conv.weight"]
        k = w.shape[-1]
        def f(x):
            y = F.conv2d(x, w, None, stride=stride, padding=(k - 1) // 2)
            y = F.batch_norm(y, d[p + ".bn.running_mean"], d[p + ".bn.running_var"], d[p + ".bn.weight"], d[p + ".bn.bias"], False, 0.0, 1e-3)
            return y * torch.sigmoid(y)
        return f

    def bottleneck(d, p):
        o = types.SimpleNamespace(conv1=conv_module(d, p + ".conv1"), conv2=conv_module(d, p + ".conv2"), add_identity=False)
        return lambda x: ns["DarknetBottleneck_forward"](o, x)

    def csp(d, p):
        blocks, j = [], 0
        while f"{p}.blocks.{j}.conv1.conv.weight" in d:
            blocks.append(bottleneck(d, f"{p}.blocks.{j}"))
            j += 1
        def seq(x):
            for b in blocks:
                x = b(x)
            return x
        o = types.SimpleNamespace(short_conv=conv_module(d, p + ".short_conv"), main_conv=conv_module(d, p + ".main_conv"),
                                  final_conv=conv_module(d, p + ".final_conv"), blocks=seq)
        return lambda x: ns["CSPLayer_forward"](o, x)

    in_ch = [nsd["reduce_layers.1.conv.weight"].shape[0], nsd["reduce_layers.0.conv.weight"].shape[0], nsd["reduce_layers.0.conv.weight"].shape[1]]
    neck = types.SimpleNamespace(
        in_channels=in_ch, upsample=lambda x: F.interpolate(x, scale_factor=2, mode="nearest"),
        reduce_layers=[conv_module(nsd, f"reduce_layers.{i}") for i in range(2)],
        top_down_blocks=[csp(nsd, f"top_down_blocks.{i}") for i in range(2)],
        downsamples=[conv_module(nsd, f"downsamples.{i}", 2) for i in range(2)],
        bottom_up_blocks=[csp(nsd, f"bottom_up_blocks.{i}") for i in range(2)],
        out_convs=[conv_module(nsd, f"out_convs.{i}") for i in range(3)])

    def tower(p):
        a, b = conv_module(hsd, p + ".0"), conv_module(hsd, p + ".1")
        return lambda x: b(a(x))

    def plain(p):
        return lambda x: F.conv2d(x, hsd[p + ".weight"], hsd[p + ".bias"])

    strides = [8, 16, 32]
    head = types.SimpleNamespace(
        multi_level_cls_convs=[tower(f"multi_level_cls_convs.{l}") for l in range(3)],
        multi_level_reg_convs=[tower(f"multi_level_reg_convs.{l}") for l in range(3)],
        multi_level_conv_cls=[plain(f"multi_level_conv_cls.{l}") for l in range(3)],
        multi_level_conv_reg=[plain(f"multi_level_conv_reg.{l}") for l in range(3)],
        multi_level_conv_obj=[plain(f"multi_level_conv_obj.{l}") for l in range(3)],
        cls_out_channels=meta["nc"], test_cfg=None, prior_generator=ns["MlvlPointGenerator"](strides, offset=0))
    for name in ("forward_single", "forward", "get_bboxes", "_bbox_decode", "_bboxes_nms"):
        setattr(head, name, types.MethodType(ns[f"YOLOXHead_{name}"], head))

    out = {}
    H, W = 96, 160
    g = torch.Generator().manual_seed(77)
    feats = [torch.randn(2, c, H // s, W // s, generator=g) for c, s in zip(in_ch, strides)]
    with torch.no_grad():
        outs = ns["YOLOXPAFPN_forward"](neck, tuple(feats))
        cls_scores, bbox_preds, objs = head.forward(outs)
        cfg = Cfg(score_thr=0.01, nms=dict(type="nms", iou_threshold=0.65))
        metas = [dict(scale_factor=np.array([1.0, 1.0, 1.0, 1.0], dtype=np.float32)), dict(scale_factor=np.array([1.25, 1.5, 1.25, 1.5], dtype=np.float32))]
        plain_res = head.get_bboxes(cls_scores, bbox_preds, objs, img_metas=metas, cfg=cfg, rescale=False)
        scaled_res = head.get_bboxes(cls_scores, bbox_preds, objs, img_metas=metas, cfg=cfg, rescale=True)
    out["meta"] = np.array([H, W, 77])
    for l in range(3):
        out[f"neck{l}"] = outs[l].numpy()
        out[f"cls{l}"] = cls_scores[l].numpy()
        out[f"box{l}"] = bbox_preds[l].numpy()
        out[f"obj{l}"] = objs[l].numpy()
    for tag, res in (("plain", plain_res), ("scaled", scaled_res)):
        for i, (d, l) in enumerate(res):
            out[f"{tag}_dets{i}"] = d.numpy()
            out[f"{tag}_labels{i}"] = l.numpy()
    np.savez_compressed(HERE / "mmdet_yolox_cases.npz", **out)
    print({k: v.shape for k, v in out.items()}, (HERE / "mmdet_yolox_cases.npz").stat().st_size)


if __name__ == "__main__":
    main()
