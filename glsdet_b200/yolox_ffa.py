"""Drop-in for yolox-drone/models/ffa/yolox_ffa.py: `YoloBody(num_classes, phi)`, `YOLOPAFPN`, `YOLOXHead`, `FFA`.

Same constructor signatures, forward signatures, tensor layouts (NCHW fp32 in and out) and state_dict keys as
the reference, so `yolox-drone/yolo.py` can load it by module path
(`importlib.import_module(config_path).YoloBody(num_classes, phi)`, yolo.py:100-102) and
`load_state_dict(torch.load(path))` strictly (yolo.py:105).

The modules below only HOLD parameters with the reference's names.  The math of the neck, the FFA block, the
head, the decode and the NMS runs in the native plan (glsdet_b200/engine.py -> libglsdet_b200.so), the CSPDarknet
backbone in glsdet_b200/backbone.py; there is no PyTorch fallback for any of it.
"""
from __future__ import annotations

import threading
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .backbone import BackbonePlan, backbone_supported
from .engine import FFAPathPlan
from .utils_bbox import DeviceNMS

_DEPTH = {"nano": 0.33, "tiny": 0.33, "s": 0.33, "m": 0.67, "l": 1.00, "x": 1.33}
_WIDTH = {"nano": 0.25, "tiny": 0.375, "s": 0.50, "m": 0.75, "l": 1.00, "x": 1.25}


class BaseConv(nn.Module):
    """Parameter holder for conv(bias=False) + BatchNorm2d(eps=1e-3, momentum=0.03) + activation
    (reference: models/base/baseConv.py:6-16).  No layer has a PyTorch forward: neck, head and backbone all run in
    native plans owned by their parent modules."""

    def __init__(self, in_channels, out_channels, ksize, stride, groups=1, bias=False, act="silu"):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, ksize, stride, (ksize - 1) // 2, groups=groups, bias=bias)
        self.bn = nn.BatchNorm2d(out_channels, eps=0.001, momentum=0.03)
        self.act_name = act

    def forward(self, x):
        raise RuntimeError("this layer belongs to the native GLSDet path and is executed by libglsdet_b200.so "
                           "through its parent module (CSPDarknet / YOLOPAFPN / YOLOXHead / YoloBody); it has no PyTorch forward")


class DWConv(nn.Module):
    """Parameter holder for the depthwise-separable conv of phi = 'nano' (reference: models/base/baseConv.py:22-30):
    dconv = BaseConv(C, C, k, stride, groups=C), pconv = BaseConv(C, Cout, 1, 1).  Executed as glsdet_dwconv (csrc/dwconv.cu)
    + the tcgen05 1x1 conv by the parent module's native plan."""

    def __init__(self, in_channels, out_channels, ksize, stride=1, act="silu"):
        super().__init__()
        self.dconv = BaseConv(in_channels, in_channels, ksize, stride, groups=in_channels, act=act)
        self.pconv = BaseConv(in_channels, out_channels, 1, 1, groups=1, act=act)

    def forward(self, x):
        return self.dconv(x)   # raises: no PyTorch forward


class Bottleneck(nn.Module):
    def __init__(self, cin, cout, shortcut=True, expansion=0.5, depthwise=False, act="silu"):
        super().__init__()
        hidden = int(cout * expansion)
        Conv = DWConv if depthwise else BaseConv
        self.conv1 = BaseConv(cin, hidden, 1, 1, act=act)
        self.conv2 = Conv(hidden, cout, 3, 1, act=act)
        self.use_add = shortcut and cin == cout


class CSPLayer(nn.Module):
    """models/ffa/darknet.py:66-112 layout: conv1, conv2, conv3, m.<j>.conv{1,2}."""

    def __init__(self, in_channels, out_channels, n=1, shortcut=True, expansion=0.5, depthwise=False, act="silu"):
        super().__init__()
        hidden = int(out_channels * expansion)
        self.conv1 = BaseConv(in_channels, hidden, 1, 1, act=act)
        self.conv2 = BaseConv(in_channels, hidden, 1, 1, act=act)
        self.conv3 = BaseConv(2 * hidden, out_channels, 1, 1, act=act)
        self.m = nn.Sequential(*[Bottleneck(hidden, hidden, shortcut, 1.0, depthwise, act=act) for _ in range(n)])


class Focus(nn.Module):
    def __init__(self, in_channels, out_channels, ksize=1, stride=1, act="silu"):
        super().__init__()
        self.conv = BaseConv(in_channels * 4, out_channels, ksize, stride, act=act)


class SPPBottleneck(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_sizes=(5, 9, 13), activation="silu"):
        super().__init__()
        hidden = in_channels // 2
        self.conv1 = BaseConv(in_channels, hidden, 1, 1, act=activation)
        self.m = nn.ModuleList([nn.MaxPool2d(ks, 1, ks // 2) for ks in kernel_sizes])
        self.conv2 = BaseConv(hidden * (len(kernel_sizes) + 1), out_channels, 1, 1, act=activation)


class SE(nn.Module):
    """models/ffa/ffa.py:5-20 parameter layout (fc.0.weight, fc.2.weight)."""

    def __init__(self, channel, reduction=16):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction, bias=False), nn.ReLU(),
                                nn.Linear(channel // reduction, channel, bias=False), nn.Sigmoid())


class FFA(nn.Module):
    """models/ffa/ffa.py:22-85 parameter layout; evaluated inside the head's native plan."""

    def __init__(self, num_channels):
        super().__init__()
        c = num_channels
        self.scale = BaseConv(c * 2, c * 4, 1, 1, act="relu")
        self.create_content_extractor = nn.Sequential(BaseConv(c * 4, c * 4, 1, 1, act="relu"),
                                                      BaseConv(c * 4, c * 4, 1, 1, act="relu"))
        self.create_text_extractor = nn.Sequential(BaseConv(c * 2, c * 2, 1, 1, act="relu"))
        self.conv3 = BaseConv(c * 2, c, 1, 1, act="relu")
        self.se1 = SE(c * 4)


FTT = FFA  # the reference constructs `FTT(...)` (yolox_ffa.py:31) while only FFA exists


class _PlanOwner(nn.Module):
    """Caches native plans per (batch, input size, device) and drops them when the weights change."""

    _neck_prefix = ""
    _head_prefix = ""
    _parts: Tuple[str, ...] = ()
    _variant = "ffa"      # "ffa": models/ffa/yolox_ffa.py; "stock": models/base/yolox.py and the mmdet pair
    _decode = "drone"     # decoded-row flavour of the fused path
    precision = "bf16"    # "bf16": tcgen05 path; "fp32": accuracy mode (SIMT fp32 kernels, 1e-3 parity bar)

    def set_precision(self, precision: str):
        """Select the arithmetic of the native plan for this module and its native children."""
        assert precision in ("bf16", "fp32")
        for m in self.modules():
            if isinstance(m, _PlanOwner):
                m.precision = precision
                m.invalidate_plans()
        return self

    def __init__(self):
        super().__init__()
        self._plans: Dict[tuple, FFAPathPlan] = {}
        self._plan_lock = threading.Lock()
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_plans())

    def _replicate_for_data_parallel(self):
        """nn.DataParallel (yolox-drone/yolo.py:110) re-creates shallow replicas of the module on every forward and strips
        their parameters into plain attributes.  The native plans are built from the ORIGINAL module's state_dict (copied to
        the replica's device once) and live in the original's plan cache, which the replicas share (the cache key holds the
        device); `_plan_lock` serialises the replica threads while a plan is being built."""
        replica = super()._replicate_for_data_parallel()
        replica._dp_source = getattr(self, "_dp_source", self)
        return replica

    def _source_state_dict(self):
        return getattr(self, "_dp_source", self).state_dict()

    def invalidate_plans(self):
        self._plans.clear()

    def _apply(self, fn, *args, **kwargs):
        self.invalidate_plans()
        return super()._apply(fn, *args, **kwargs)

    def train(self, mode: bool = True):
        if mode:
            raise RuntimeError("glsdet_b200 modules are inference-only (BatchNorm is folded into the convs)")
        return super().train(False)

    def _num_classes(self) -> int:
        raise NotImplementedError

    def _plan_state_dict(self):
        """state_dict in yolox-drone naming (subclasses with other naming translate here)."""
        return self._source_state_dict()

    def _plan(self, batch: int, input_hw: Sequence[int], device) -> FFAPathPlan:
        key = (batch, int(input_hw[0]), int(input_hw[1]), str(torch.device(device)), self.precision)
        with self._plan_lock:
            plan = self._plans.get(key)
            if plan is None:
                if len(self._plans) >= 8:
                    self._plans.clear()
                with torch.cuda.device(torch.device(device)):
                    plan = FFAPathPlan(self._plan_state_dict(), batch, input_hw, self._num_classes(), device=device,
                                       neck_prefix=self._neck_prefix, head_prefix=self._head_prefix, parts=self._parts,
                                       variant=self._variant, decode=self._decode, precision=self.precision)
                self._plans[key] = plan
        return plan

    def _fused_plan(self, x: torch.Tensor) -> Optional[FFAPathPlan]:
        """Plan with the native CSPDarknet chained in front (image -> backbone -> neck -> head in NHWC bf16), or None
        when the backbone cannot run natively for this input (CPU tensor, fp32 mode, phi='tiny') or the topology needs
        the NCHW fp32 features (P1 pre-loads)."""
        return self._fused_plan_for(x.shape[0], x.shape[2], x.shape[3], x.device)

    def _fused_plan_for(self, batch: int, height: int, width: int, device) -> Optional[FFAPathPlan]:
        bb = self.backbone.backbone
        if not (isinstance(bb, CSPDarknet) and bb.native_ok_for(height, width, device)):
            return None
        plan = self._plan(batch, (height, width), device)
        if plan.fp32 or plan.pre_loads:
            return None
        with self._plan_lock:
            if plan.backbone is None:
                with torch.cuda.device(torch.device(device)):
                    plan.attach_backbone(bb._source_state_dict(), "", bb.act_name)
        return plan


class CSPDarknet(_PlanOwner):
    """Backbone (models/ffa/darknet.py:115-195): forward(image batch) -> {"dark2": .., .. "dark5": ..} NCHW fp32.

    The forward runs as the native BackbonePlan (glsdet_b200/backbone.py: Focus kernel, tcgen05 convs - SIMT fp32 convs in
    the fp32 accuracy mode -, SPP pooling kernel).  The PyTorch layers below only hold the parameters: there is no PyTorch
    or CPU fallback, a CPU tensor raises."""

    def __init__(self, dep_mul, wid_mul, out_features=("dark2", "dark3", "dark4", "dark5"), depthwise=False,
                 act="silu"):
        super().__init__()
        self.out_features = out_features
        self.act_name = act
        c = int(wid_mul * 64)
        d = max(round(dep_mul * 3), 1)
        self.base_channels = c
        self.stem = Focus(3, c, ksize=3, act=act)
        Conv = DWConv if depthwise else BaseConv   # darknet.py:120

        def stage(cin, cout, n, shortcut=True, spp=False):
            layers = [Conv(cin, cout, 3, 2, act=act)]
            if spp:
                layers.append(SPPBottleneck(cout, cout, activation=act))
            layers.append(CSPLayer(cout, cout, n=n, shortcut=shortcut, depthwise=depthwise, act=act))
            return nn.Sequential(*layers)

        self.dark2 = stage(c, c * 2, d)
        self.dark3 = stage(c * 2, c * 4, d * 3)
        self.dark4 = stage(c * 4, c * 8, d * 3)
        self.dark5 = stage(c * 8, c * 16, d, shortcut=False, spp=True)
        self._bb_plans: Dict[tuple, BackbonePlan] = {}
        super().train(False)

    def invalidate_plans(self):
        super().invalidate_plans()
        if hasattr(self, "_bb_plans"):
            self._bb_plans.clear()

    def native_ok(self, x: torch.Tensor) -> bool:
        return self.native_ok_for(x.shape[2], x.shape[3], x.device)

    def native_ok_for(self, height: int, width: int, device) -> bool:
        """True when the backbone can be chained in front of a bf16 neck plan for this input."""
        return bool(torch.device(device).type == "cuda" and self.precision == "bf16" and
                    backbone_supported(self.base_channels) and height % 32 == 0 and width % 32 == 0)

    def native_plan(self, batch: int, input_hw: Sequence[int], device, outs=None, key_extra=None) -> BackbonePlan:
        key = (batch, int(input_hw[0]), int(input_hw[1]), str(torch.device(device)), key_extra, self.precision)
        with self._plan_lock:
            plan = self._bb_plans.get(key)
            if plan is None:
                if len(self._bb_plans) >= 8:
                    self._bb_plans.clear()
                with torch.cuda.device(torch.device(device)):
                    plan = BackbonePlan(self._source_state_dict(), batch, input_hw, device=device, act=self.act_name, prefix="",
                                        outs=outs, precision=self.precision)
                self._bb_plans[key] = plan
        return plan

    @torch.no_grad()
    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("CSPDarknet belongs to the native GLSDet path and is executed by libglsdet_b200.so on a CUDA "
                               "device; there is no CPU / PyTorch forward")
        plan = self.native_plan(x.shape[0], x.shape[2:], x.device)
        plan.run(x.float().contiguous())
        return plan.features_nchw(self.out_features)


class YOLOXHead(_PlanOwner):
    """models/ffa/yolox_ffa.py:12-118.  forward(inputs) takes the tuple YOLOPAFPN.forward returns
    (feat0, P3_out, P4_out, P5_out; NCHW fp32) and returns the list of raw [B, 5+nc, h, w] maps."""

    _parts = ("head",)

    def __init__(self, num_classes, width=1.0, in_channels=[256, 512, 1024, 256], act="silu", depthwise=False):
        super().__init__()
        self.num_classes = num_classes
        hc = int(256 * width)
        Conv = DWConv if depthwise else BaseConv   # yolox_ffa.py:15
        self.cls_convs, self.reg_convs = nn.ModuleList(), nn.ModuleList()
        self.cls_preds, self.reg_preds, self.obj_preds = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.stems = nn.ModuleList()
        self.csp = CSPLayer(int(0.5 * in_channels[0] * width), int(in_channels[0] * width), round(3 * 0.75), False,
                            depthwise=depthwise, act=act)
        self.ftt = FTT(int(width * in_channels[0]))
        for i, cin in enumerate(in_channels):
            if i != 3:
                self.stems.append(BaseConv(int(cin * width), hc, 1, 1, act=act))
            self.cls_convs.append(nn.Sequential(Conv(hc, hc, 3, 1, act=act), Conv(hc, hc, 3, 1, act=act)))
            self.reg_convs.append(nn.Sequential(Conv(hc, hc, 3, 1, act=act), Conv(hc, hc, 3, 1, act=act)))
            self.cls_preds.append(nn.Conv2d(hc, num_classes, 1, 1, 0))
            self.reg_preds.append(nn.Conv2d(hc, 4, 1, 1, 0))
            self.obj_preds.append(nn.Conv2d(hc, 1, 1, 1, 0))
        super().train(False)

    def _num_classes(self):
        return self.num_classes

    @torch.no_grad()
    def forward(self, inputs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        p3 = inputs[1]
        plan = self._plan(p3.shape[0], (p3.shape[2] * 8, p3.shape[3] * 8), p3.device)
        plan.load_head_inputs([t.float() for t in inputs])
        plan.run_head(decoded=False)
        return [t.clone() for t in plan.logits]


class YOLOPAFPN(_PlanOwner):
    """models/ffa/yolox_ffa.py:121-261.  forward(image batch) -> (feat0, P3_out, P4_out, P5_out), NCHW fp32."""

    _parts = ("neck",)

    def __init__(self, depth=1.0, width=1.0, in_features=("dark2", "dark3", "dark4", "dark5"),
                 in_channels=[256, 512, 1024], depthwise=False, act="silu"):
        super().__init__()
        self.backbone = CSPDarknet(depth, width, depthwise=depthwise, act=act)
        self.in_features = in_features
        c0, c1, c2 = (int(c * width) for c in in_channels)
        n = round(3 * depth)
        Conv = DWConv if depthwise else BaseConv   # yolox_ffa.py:125
        self.lateral_conv0 = BaseConv(c2, c1, 1, 1, act=act)
        self.C3_p4 = CSPLayer(2 * c1, c1, n, False, depthwise=depthwise, act=act)
        self.reduce_conv1 = BaseConv(c1, c0, 1, 1, act=act)
        self.C3_p3 = CSPLayer(2 * c0, c0, n, False, depthwise=depthwise, act=act)
        self.bu_conv2 = Conv(c0, c0, 3, 2, act=act)
        self.C3_n3 = CSPLayer(2 * c0, c1, n, False, depthwise=depthwise, act=act)
        self.bu_conv1 = Conv(c1, c1, 3, 2, act=act)
        self.C3_n4 = CSPLayer(2 * c1, c2, n, False, depthwise=depthwise, act=act)
        super().train(False)

    def _num_classes(self):
        return 1  # unused by the neck-only plan

    @torch.no_grad()
    def features(self, x: torch.Tensor) -> List[torch.Tensor]:
        out = self.backbone(x)
        return [out[f] for f in self.in_features]

    @torch.no_grad()
    def forward_features(self, feats: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, ...]:
        """Neck only, from (dark2, dark3, dark4, dark5) NCHW fp32."""
        f0 = feats[0]
        plan = self._plan(f0.shape[0], (f0.shape[2] * 4, f0.shape[3] * 4), f0.device)
        plan.load_features([t.float() for t in feats])
        plan.run_neck()
        outs = plan.neck_outputs_nchw()
        return (feats[0], outs[1], outs[2], outs[3])

    def forward(self, input: torch.Tensor):
        return self.forward_features(self.features(input))


class YoloBody(_PlanOwner):
    """models/ffa/yolox_ffa.py:264-284.  forward(x) returns the raw per-level maps exactly like the reference;
    `detect` is the fused neck -> head -> decode -> filter -> NMS path."""

    _neck_prefix = "backbone."
    _head_prefix = "head."
    _parts = ("neck", "head")

    def __init__(self, num_classes, phi):
        super().__init__()
        depth, width = _DEPTH[phi], _WIDTH[phi]
        depthwise = phi == "nano"
        self.num_classes = num_classes
        self.backbone = YOLOPAFPN(depth, width, depthwise=depthwise)
        self.head = YOLOXHead(num_classes, width, depthwise=depthwise)
        self._nms: Dict[tuple, DeviceNMS] = {}
        super().train(False)

    def _num_classes(self):
        return self.num_classes

    def plan_for(self, feats: Sequence[torch.Tensor]) -> FFAPathPlan:
        f0 = feats[0]
        return self._plan(f0.shape[0], (f0.shape[2] * 4, f0.shape[3] * 4), f0.device)

    @torch.no_grad()
    def forward_features(self, feats: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        """Neck + head from backbone features (dark2..dark5, NCHW fp32): list of raw [B, 5+nc, h, w]."""
        return [t.clone() for t in self.plan_for(feats).forward_logits(feats)]

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        plan = self._fused_plan(x)
        if plan is None:
            return self.forward_features(self.backbone.features(x))
        return [t.clone() for t in plan.forward_image(x.float().contiguous(), False)]

    @torch.no_grad()
    def decode_features(self, feats: Sequence[torch.Tensor]) -> torch.Tensor:
        """decode_outputs(forward(...)) fused: [B, A, 5+nc] (cx, cy, w, h normalised, obj, cls)."""
        return self.plan_for(feats).forward_decoded(feats)

    def nms_for(self, plan: FFAPathPlan, max_det: Optional[int] = None) -> DeviceNMS:
        key = (plan.B, plan.num_anchors, self.num_classes, max_det, str(plan.device))
        nms = self._nms.get(key)
        if nms is None:
            nms = DeviceNMS(plan.B, plan.num_anchors, self.num_classes, max_det=max_det, device=plan.device)
            self._nms = {key: nms}
        return nms

    @torch.no_grad()
    def detect_features(self, feats: Sequence[torch.Tensor], conf_thres: float = 0.5, nms_thres: float = 0.4,
                        strategy: str = "auto_cuda", max_det: Optional[int] = None, graph: bool = False):
        """Whole hot path on the device.  Returns (det [B, max_det, 7], count [B]) device tensors; rows are
        (x1, y1, x2, y2 normalised network coordinates, obj_conf, class_conf, class_pred), sorted by score.
        `graph=True` replays a CUDA graph of neck -> head -> filter -> NMS captured once per plan (engine.GraphedPath); the
        returned tensors are then the graph's fixed outputs, overwritten by the next call."""
        plan = self.plan_for(feats)
        if graph:
            key = (max_det, float(conf_thres), float(nms_thres), strategy)
            graphs = plan.__dict__.setdefault("_graphs", {})     # the captured graphs live and die with their plan
            gp = graphs.get(key)
            if gp is None:
                from .engine import GraphedPath
                graphs.clear()
                gp = graphs[key] = GraphedPath(plan, self.nms_for(plan, max_det), conf_thres, nms_thres, strategy)
            plan.load_features(feats)
            return gp.replay()
        pred = plan.forward_detect(feats)
        return self.nms_for(plan, max_det).launch(pred, conf_thres, nms_thres, strategy, cls_logits=plan.det_cls_logits)

    @torch.no_grad()
    def detect_uint8(self, images_u8: torch.Tensor, conf_thres: float = 0.5, nms_thres: float = 0.4,
                     strategy: str = "auto_cuda", max_det: Optional[int] = None, **norm):
        """uint8 HWC image batch [B, H, W, 3] on the device (already letterboxed to the network size, yolo.py:130) ->
        detections.  preprocess_input (models/core/utils.py:47-51) and the transpose of yolo.py:134 run inside the Focus
        kernel, bit-identical to the host preprocessing; `mean=` / `std=` override the ImageNet constants."""
        b, h, w, _ = images_u8.shape
        plan = self._fused_plan_for(b, h, w, images_u8.device)
        if plan is None:
            raise NotImplementedError("detect_uint8 needs the chained native backbone (bf16 plan without pre-loads)")
        pred = plan.forward_image_uint8(images_u8.contiguous(), "det", **norm)
        return self.nms_for(plan, max_det).launch(pred, conf_thres, nms_thres, strategy, cls_logits=plan.det_cls_logits)

    @torch.no_grad()
    def detect(self, x: torch.Tensor, conf_thres: float = 0.5, nms_thres: float = 0.4, strategy: str = "auto_cuda",
               max_det: Optional[int] = None):
        """Image batch -> detections: backbone -> neck -> head -> decode -> filter -> NMS, all native."""
        plan = self._fused_plan(x)
        if plan is None:
            return self.detect_features(self.backbone.features(x), conf_thres, nms_thres, strategy, max_det)
        pred = plan.forward_image(x.float().contiguous(), "det")
        return self.nms_for(plan, max_det).launch(pred, conf_thres, nms_thres, strategy, cls_logits=plan.det_cls_logits)
