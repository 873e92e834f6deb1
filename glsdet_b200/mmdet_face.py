"""mmdet registry face of the same path: `@NECKS.register_module() YOLOXPAFPN`, `@HEADS.register_module() YOLOXHead`
(yolox-ufp/mmdet/models/necks/yolox_pafpn.py:13-156, dense_heads/yolox_head.py:20-322, models/builder.py:7-45).

Same constructor arguments, forward signatures, tensor layouts and state_dict keys as the mmdet 2.19.1 modules, so
`NECKS.build(dict(type='YOLOXPAFPN', in_channels=[128, 256, 512], out_channels=128, num_csp_blocks=1))` and
`HEADS.build(dict(type='YOLOXHead', num_classes=80, in_channels=128, feat_channels=128, test_cfg=...))`
(configs/yolox/yolox_s_8x8_300e_coco.py:12-22) give drop-ins whose math runs in libglsdet_b200.so.  mmcv / mmdet are
not required (and not installed here): a minimal Registry with mmcv's `register_module` / `build` contract is
provided; when mmdet IS importable the classes are additionally registered into its registries with force=True by
`register_into_mmdet()`.

Key map onto the yolox-drone naming the plan builder uses:
  reduce_layers.0/1 -> lateral_conv0 / reduce_conv1, top_down_blocks.0/1 -> C3_p4 / C3_p3,
  downsamples.0/1 -> bu_conv2 / bu_conv1, bottom_up_blocks.0/1 -> C3_n3 / C3_n4, out_convs.i -> head.stems.i,
  CSP main_conv/short_conv/final_conv/blocks.j -> conv1/conv2/conv3/m.j,
  multi_level_{cls,reg}_convs.l -> {cls,reg}_convs.l, multi_level_conv_{cls,reg,obj}.l -> {cls,reg,obj}_preds.l.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _native as N
from .utils_bbox import STRATEGIES, DeviceNMS
from .yolox_ffa import BaseConv, _PlanOwner


class Registry:
    """The part of mmcv.utils.Registry that model configs rely on: register_module() and build(cfg)."""

    def __init__(self, name: str):
        self.name = name
        self._modules: Dict[str, type] = {}

    def register_module(self, name: Optional[str] = None, force: bool = False, module: Optional[type] = None):
        def _register(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._modules[key] = cls
            return cls

        return _register(module) if module is not None else _register

    def get(self, key: str):
        return self._modules.get(key)

    def build(self, cfg: dict, default_args: Optional[dict] = None):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise KeyError('cfg must be a dict with the key "type"')
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        typ = args.pop("type")
        cls = typ if isinstance(typ, type) else self._modules.get(typ)
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        return cls(**args)


MODELS = Registry("models")
NECKS = HEADS = MODELS   # mmdet/models/builder.py:9-15: every registry is the same MODELS object


class _Cfg(dict):
    """dict with attribute access (what mmcv.Config gives test_cfg)."""

    __getattr__ = dict.get


def _cfg(c):
    if c is None:
        return None
    return _Cfg({k: _cfg(v) if isinstance(v, dict) else v for k, v in dict(c).items()})


def _act_name(act_cfg) -> str:
    typ = (act_cfg or {}).get("type", "Swish")
    if typ in ("Swish", "SiLU"):
        return "silu"
    if typ == "ReLU":
        return "relu"
    if typ == "LeakyReLU":
        return "lrelu"
    raise NotImplementedError(f"activation {typ} is not supported by the native path")


def _check_norm(norm_cfg):
    if not norm_cfg or norm_cfg.get("type") != "BN":
        raise NotImplementedError("only BatchNorm (folded at load time) is supported")
    if abs(norm_cfg.get("eps", 1e-5) - 1e-3) > 1e-12:
        raise NotImplementedError("the native plan folds BatchNorm with eps=1e-3 (YOLOX default)")


class _DWSepConv(nn.Module):
    """Parameter holder of mmcv's DepthwiseSeparableConvModule as the YOLOX modules build it with `use_depthwise=True`
    (necks/yolox_pafpn.py:55, utils/csp_layer.py:44, dense_heads/yolox_head.py:146-147): depthwise_conv = ConvModule(C, C, k,
    stride, groups=C) and pointwise_conv = ConvModule(C, Cout, 1), each conv -> BN -> act.  mmcv is absent, so this layout
    (attribute names depthwise_conv / pointwise_conv, norm and activation on both halves: mmcv 1.x defaults dw_norm_cfg =
    pw_norm_cfg = norm_cfg, dw_act_cfg = pw_act_cfg = act_cfg) is restated; the math is the yolox-drone DWConv
    (models/base/baseConv.py:22-30), which IS pinned, and runs on glsdet_dwconv + the tcgen05 1x1 conv."""

    def __init__(self, in_channels, out_channels, ksize, stride, act):
        super().__init__()
        self.depthwise_conv = BaseConv(in_channels, in_channels, ksize, stride, groups=in_channels, act=act)
        self.pointwise_conv = BaseConv(in_channels, out_channels, 1, 1, act=act)


def _conv3(depthwise, cin, cout, stride, act):
    return _DWSepConv(cin, cout, 3, stride, act) if depthwise else BaseConv(cin, cout, 3, stride, act=act)


_DW_MAP = ((".depthwise_conv.", ".dconv."), (".pointwise_conv.", ".pconv."))


class _CSPLayer(nn.Module):
    """mmdet/models/utils/csp_layer.py:75-150 parameter layout (main_conv, short_conv, final_conv, blocks.j.conv1/2)."""

    def __init__(self, in_channels, out_channels, num_blocks, act, depthwise=False):
        super().__init__()
        mid = int(out_channels * 0.5)
        self.main_conv = BaseConv(in_channels, mid, 1, 1, act=act)
        self.short_conv = BaseConv(in_channels, mid, 1, 1, act=act)
        self.final_conv = BaseConv(2 * mid, out_channels, 1, 1, act=act)

        class _Block(nn.Module):
            def __init__(self):
                super().__init__()
                self.conv1 = BaseConv(mid, mid, 1, 1, act=act)
                self.conv2 = _conv3(depthwise, mid, mid, 1, act)

        self.blocks = nn.Sequential(*[_Block() for _ in range(num_blocks)])


_NECK_MAP = (("reduce_layers.0.", "lateral_conv0."), ("reduce_layers.1.", "reduce_conv1."),
             ("top_down_blocks.0.", "C3_p4."), ("top_down_blocks.1.", "C3_p3."),
             ("downsamples.0.", "bu_conv2."), ("downsamples.1.", "bu_conv1."),
             ("bottom_up_blocks.0.", "C3_n3."), ("bottom_up_blocks.1.", "C3_n4."))
_CSP_MAP = ((".main_conv.", ".conv1."), (".short_conv.", ".conv2."), (".final_conv.", ".conv3."), (".blocks.", ".m."))


@NECKS.register_module()
class YOLOXPAFPN(_PlanOwner):
    """forward(inputs: tuple of 3 NCHW maps) -> tuple of 3 NCHW maps with `out_channels` channels."""

    _neck_prefix = "backbone."
    _head_prefix = "head."
    _parts = ("neck", "stems")
    _variant = "stock"

    def __init__(self, in_channels, out_channels, num_csp_blocks=3, use_depthwise=False,
                 upsample_cfg=dict(scale_factor=2, mode="nearest"), conv_cfg=None,
                 norm_cfg=dict(type="BN", momentum=0.03, eps=0.001), act_cfg=dict(type="Swish"), init_cfg=None):
        super().__init__()
        if conv_cfg is not None:
            raise NotImplementedError("custom conv layers are not supported by the native path")
        if len(in_channels) != 3 or upsample_cfg.get("scale_factor", 2) != 2 or upsample_cfg.get("mode") != "nearest":
            raise NotImplementedError("the native plan implements the 3-level nearest-x2 YOLOX PAFPN")
        if not (in_channels[1] == 2 * in_channels[0] and in_channels[2] == 4 * in_channels[0]):
            raise NotImplementedError("in_channels must be (c, 2c, 4c)")
        _check_norm(norm_cfg)
        act = _act_name(act_cfg)
        self._act = act
        self.in_channels, self.out_channels = list(in_channels), out_channels
        self.reduce_layers, self.top_down_blocks = nn.ModuleList(), nn.ModuleList()
        for idx in range(2, 0, -1):
            self.reduce_layers.append(BaseConv(in_channels[idx], in_channels[idx - 1], 1, 1, act=act))
            self.top_down_blocks.append(_CSPLayer(in_channels[idx - 1] * 2, in_channels[idx - 1], num_csp_blocks, act, use_depthwise))
        self.downsamples, self.bottom_up_blocks = nn.ModuleList(), nn.ModuleList()
        for idx in range(2):
            self.downsamples.append(_conv3(use_depthwise, in_channels[idx], in_channels[idx], 2, act))
            self.bottom_up_blocks.append(_CSPLayer(in_channels[idx] * 2, in_channels[idx + 1], num_csp_blocks, act, use_depthwise))
        self.out_convs = nn.ModuleList([BaseConv(c, out_channels, 1, 1, act=act) for c in in_channels])
        nn.Module.train(self, False)

    def init_weights(self):
        pass

    def _num_classes(self):
        return 1

    def _plan_state_dict(self):
        out = {}
        for k, v in self._source_state_dict().items():
            if k.startswith("out_convs."):
                out["head.stems." + k[len("out_convs."):]] = v
                continue
            for a, b in _NECK_MAP:
                if k.startswith(a):
                    k = b + k[len(a):]
                    break
            for a, b in _CSP_MAP + _DW_MAP:
                k = k.replace(a, b)
            out["backbone." + k] = v
        return out

    def _plan(self, batch, input_hw, device):
        plan = super()._plan(batch, input_hw, device)
        return plan

    @torch.no_grad()
    def forward(self, inputs: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, ...]:
        assert len(inputs) == len(self.in_channels)
        f0 = inputs[0]
        plan = self._plan(f0.shape[0], (f0.shape[2] * 8, f0.shape[3] * 8), f0.device)
        plan.load_features([t.float() for t in inputs])
        plan.run_neck()
        plan.run_stems()
        return tuple(plan.stem_outputs_nchw())


@HEADS.register_module()
class YOLOXHead(_PlanOwner):
    """forward(feats) -> (cls_scores, bbox_preds, objectnesses); get_bboxes(...) -> [(dets [n,5], labels [n])]."""

    _head_prefix = "head."
    _parts = ("towers",)
    _variant = "stock"
    _decode = "mmdet"

    def __init__(self, num_classes, in_channels, feat_channels=256, stacked_convs=2, strides=[8, 16, 32],
                 use_depthwise=False, dcn_on_last_conv=False, conv_bias="auto", conv_cfg=None,
                 norm_cfg=dict(type="BN", momentum=0.03, eps=0.001), act_cfg=dict(type="Swish"), loss_cls=None,
                 loss_bbox=None, loss_obj=None, loss_l1=None, train_cfg=None, test_cfg=None, init_cfg=None):
        super().__init__()
        if dcn_on_last_conv or conv_cfg is not None:
            raise NotImplementedError("DCN / custom conv layers are not supported by the native path")
        if stacked_convs != 2 or list(strides) != [8, 16, 32] or in_channels != feat_channels:
            raise NotImplementedError("the native plan implements stacked_convs=2, strides [8,16,32], "
                                      "in_channels == feat_channels")
        if conv_bias not in ("auto", False):
            raise NotImplementedError("tower convs carry BatchNorm, hence no conv bias")
        _check_norm(norm_cfg)
        act = _act_name(act_cfg)
        self.num_classes = self.cls_out_channels = num_classes
        self.in_channels, self.feat_channels, self.strides = in_channels, feat_channels, list(strides)
        self.test_cfg, self.train_cfg = _cfg(test_cfg), train_cfg
        fc = feat_channels

        def tower():
            return nn.Sequential(_conv3(use_depthwise, in_channels, fc, 1, act), _conv3(use_depthwise, fc, fc, 1, act))

        self.multi_level_cls_convs = nn.ModuleList([tower() for _ in strides])
        self.multi_level_reg_convs = nn.ModuleList([tower() for _ in strides])
        self.multi_level_conv_cls = nn.ModuleList([nn.Conv2d(fc, num_classes, 1) for _ in strides])
        self.multi_level_conv_reg = nn.ModuleList([nn.Conv2d(fc, 4, 1) for _ in strides])
        self.multi_level_conv_obj = nn.ModuleList([nn.Conv2d(fc, 1, 1) for _ in strides])
        self._nms: Dict[tuple, DeviceNMS] = {}
        nn.Module.train(self, False)

    def init_weights(self):
        pass

    def _num_classes(self):
        return self.num_classes

    def _plan_state_dict(self):
        ren = (("multi_level_cls_convs.", "cls_convs."), ("multi_level_reg_convs.", "reg_convs."),
               ("multi_level_conv_cls.", "cls_preds."), ("multi_level_conv_reg.", "reg_preds."),
               ("multi_level_conv_obj.", "obj_preds."))
        out = {}
        for k, v in self._source_state_dict().items():
            for a, b in ren:
                if k.startswith(a):
                    k = b + k[len(a):]
                    break
            for a, b in _DW_MAP:
                k = k.replace(a, b)
            out["head." + k] = v
        return out

    @torch.no_grad()
    def forward(self, feats: Sequence[torch.Tensor]):
        """Returns three lists (one entry per level) of [B, nc, h, w], [B, 4, h, w], [B, 1, h, w] raw maps
        (views of one [B, 5+nc, h, w] buffer per level)."""
        f0 = feats[0]
        plan = self._plan(f0.shape[0], (f0.shape[2] * 8, f0.shape[3] * 8), f0.device)
        plan.load_tower_inputs([t.float() for t in feats])
        plan.run_towers(decoded=False)
        maps = [t.clone() for t in plan.logits]
        return [m[:, 5:] for m in maps], [m[:, :4] for m in maps], [m[:, 4:5] for m in maps]

    @torch.no_grad()
    def get_bboxes(self, cls_scores, bbox_preds, objectnesses, img_metas=None, cfg=None, rescale=False,
                   with_nms=True):
        """yolox_head.py:215-322.  Any raw maps with the reference layout are accepted (not only forward()'s)."""
        lib = N.load()
        cfg = self.test_cfg if cfg is None else _cfg(cfg)
        if cfg is None:
            raise ValueError("get_bboxes needs test_cfg (score_thr, nms=dict(type='nms', iou_threshold=...))")
        nms_cfg = dict(cfg["nms"])
        if nms_cfg.get("type", "nms") != "nms" or nms_cfg.get("class_agnostic", False):
            raise NotImplementedError("only nms=dict(type='nms', iou_threshold=...) is supported")
        n = len(cls_scores)
        assert n == len(bbox_preds) == len(objectnesses)
        b = cls_scores[0].shape[0]
        dev = cls_scores[0].device
        if not cls_scores[0].is_cuda:
            raise N.NativeError("get_bboxes needs CUDA tensors (glsdet_b200 has no CPU path)")

        def prep(ts):  # channel slices of a contiguous [B, C, h, w] buffer are passed without a copy
            ok = all(t.dtype == torch.float32 and t.stride(3) == 1 and t.stride(2) == t.shape[3] and
                     t.stride(1) == t.shape[2] * t.shape[3] for t in ts)
            ts = list(ts) if ok else [t.float().contiguous() for t in ts]
            return ts, (C.c_void_p * n)(*[t.data_ptr() for t in ts]), (C.c_int64 * n)(*[t.stride(0) for t in ts])

        cls_t, cls_p, cls_bs = prep(cls_scores)
        box_t, box_p, box_bs = prep(bbox_preds)
        obj_t, obj_p, obj_bs = prep(objectnesses)
        hs = (C.c_int32 * n)(*[t.shape[2] for t in cls_t])
        ws = (C.c_int32 * n)(*[t.shape[3] for t in cls_t])
        st = (C.c_int32 * n)(*self.strides[:n])
        a = sum(t.shape[2] * t.shape[3] for t in cls_t)
        nc = self.num_classes
        pred = torch.empty((b, a, 5 + nc), dtype=torch.float32, device=dev)
        N.check(lib.glsdet_decode_mmdet(cls_p, box_p, obj_p, cls_bs, box_bs, obj_bs, hs, ws, st, n, b, nc,
                                        pred.data_ptr(), N.stream_ptr()), "glsdet_decode_mmdet")
        key = (b, a, nc, str(dev))
        if key not in self._nms:
            self._nms = {key: DeviceNMS(b, a, nc, device=dev)}
        op = self._nms[key]
        div = None
        if rescale:
            sf = [list(m["scale_factor"]) for m in img_metas]
            div = torch.tensor(sf, dtype=torch.float32, device=dev).contiguous()
        det, count = op.launch(pred, float(cfg["score_thr"]), float(nms_cfg["iou_threshold"]), "mmcv", box_div=div)
        # the one host synchronisation of get_bboxes: the per-image (dets [n_i, 5], labels [n_i]) tensors it returns are
        # sized by the counts (mmdet's own _bboxes_nms indexes with boolean masks, which synchronises per image)
        counts = count.cpu().tolist()
        results = []
        for i in range(b):
            rows = det[i, :counts[i]]
            dets = torch.cat([rows[:, :4], (rows[:, 4] * rows[:, 5]).unsqueeze(1)], dim=1)
            results.append((dets, rows[:, 6].long()))
        return results


def register_into_mmdet() -> bool:
    """If a real mmdet is importable, put the drop-ins into its registries (force=True) and return True."""
    try:
        from mmdet.models.builder import HEADS as _H, NECKS as _N  # type: ignore
    except Exception:
        return False
    _N.register_module(name="YOLOXPAFPN", force=True, module=YOLOXPAFPN)
    _H.register_module(name="YOLOXHead", force=True, module=YOLOXHead)
    return True
