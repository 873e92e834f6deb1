"""Drop-in for yolox-drone/models/base/yolox.py: the stock three-level YOLOX (`YoloBody`, `YOLOPAFPN`, `YOLOXHead`)
with the reference's constructor signatures, NCHW fp32 tensors and state_dict keys, executed by the native plan
(variant "stock" of glsdet_b200/engine.py).  The same math is what the mmdet pair YOLOXPAFPN + YOLOXHead computes;
see glsdet_b200/mmdet_face.py for that naming.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .engine import FFAPathPlan
from .utils_bbox import DeviceNMS
from .yolox_ffa import _DEPTH, _WIDTH, BaseConv, CSPDarknet, CSPLayer, DWConv, _PlanOwner


class YOLOXHead(_PlanOwner):
    """models/base/yolox.py YOLOXHead: forward(inputs=(P3_out, P4_out, P5_out)) -> list of raw [B, 5+nc, h, w]."""

    _parts = ("head",)
    _variant = "stock"

    def __init__(self, num_classes, width=1.0, in_channels=[256, 512, 1024], act="silu", depthwise=False):
        super().__init__()
        self.num_classes = num_classes
        hc = int(256 * width)
        Conv = DWConv if depthwise else BaseConv   # models/base/yolox.py:14
        self.cls_convs, self.reg_convs = nn.ModuleList(), nn.ModuleList()
        self.cls_preds, self.reg_preds, self.obj_preds = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.stems = nn.ModuleList()
        for cin in in_channels:
            self.stems.append(BaseConv(int(cin * width), hc, 1, 1, act=act))
            self.cls_convs.append(nn.Sequential(Conv(hc, hc, 3, 1, act=act), Conv(hc, hc, 3, 1, act=act)))
            self.cls_preds.append(nn.Conv2d(hc, num_classes, 1, 1, 0))
            self.reg_convs.append(nn.Sequential(Conv(hc, hc, 3, 1, act=act), Conv(hc, hc, 3, 1, act=act)))
            self.reg_preds.append(nn.Conv2d(hc, 4, 1, 1, 0))
            self.obj_preds.append(nn.Conv2d(hc, 1, 1, 1, 0))
        nn.Module.train(self, False)

    def _num_classes(self):
        return self.num_classes

    @torch.no_grad()
    def forward(self, inputs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        p3 = inputs[0]
        plan = self._plan(p3.shape[0], (p3.shape[2] * 8, p3.shape[3] * 8), p3.device)
        plan.load_head_inputs([t.float() for t in inputs])
        plan.run_head(decoded=False)
        return [t.clone() for t in plan.logits]


class YOLOPAFPN(_PlanOwner):
    """models/base/yolox.py YOLOPAFPN: forward(images) -> (P3_out, P4_out, P5_out)."""

    _parts = ("neck",)
    _variant = "stock"

    def __init__(self, depth=1.0, width=1.0, in_features=("dark3", "dark4", "dark5"), in_channels=[256, 512, 1024],
                 depthwise=False, act="silu"):
        super().__init__()
        self.backbone = CSPDarknet(depth, width, depthwise=depthwise, act=act)
        self.in_features = in_features
        c0, c1, c2 = (int(c * width) for c in in_channels)
        n = round(3 * depth)
        Conv = DWConv if depthwise else BaseConv   # models/base/yolox.py:99
        self.lateral_conv0 = BaseConv(c2, c1, 1, 1, act=act)
        self.C3_p4 = CSPLayer(2 * c1, c1, n, False, depthwise=depthwise, act=act)
        self.reduce_conv1 = BaseConv(c1, c0, 1, 1, act=act)
        self.C3_p3 = CSPLayer(2 * c0, c0, n, False, depthwise=depthwise, act=act)
        self.bu_conv2 = Conv(c0, c0, 3, 2, act=act)
        self.C3_n3 = CSPLayer(2 * c0, c1, n, False, depthwise=depthwise, act=act)
        self.bu_conv1 = Conv(c1, c1, 3, 2, act=act)
        self.C3_n4 = CSPLayer(2 * c1, c2, n, False, depthwise=depthwise, act=act)
        nn.Module.train(self, False)

    def _num_classes(self):
        return 1

    @torch.no_grad()
    def features(self, x: torch.Tensor) -> List[torch.Tensor]:
        out = self.backbone(x)
        return [out[f] for f in self.in_features]

    @torch.no_grad()
    def forward_features(self, feats: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, ...]:
        f0 = feats[0]
        plan = self._plan(f0.shape[0], (f0.shape[2] * 8, f0.shape[3] * 8), f0.device)
        plan.load_features([t.float() for t in feats])
        plan.run_neck()
        return tuple(plan.neck_outputs_nchw())

    def forward(self, input: torch.Tensor):
        return self.forward_features(self.features(input))


class YoloBody(_PlanOwner):
    """models/base/yolox.py YoloBody(num_classes, phi)."""

    _neck_prefix = "backbone."
    _head_prefix = "head."
    _parts = ("neck", "head")
    _variant = "stock"

    def __init__(self, num_classes, phi):
        super().__init__()
        depth, width = _DEPTH[phi], _WIDTH[phi]
        self.num_classes = num_classes
        self.backbone = YOLOPAFPN(depth, width, depthwise=phi == "nano")
        self.head = YOLOXHead(num_classes, width, depthwise=phi == "nano")
        self._nms = {}
        nn.Module.train(self, False)

    def _num_classes(self):
        return self.num_classes

    def plan_for(self, feats: Sequence[torch.Tensor]) -> FFAPathPlan:
        f0 = feats[0]
        return self._plan(f0.shape[0], (f0.shape[2] * 8, f0.shape[3] * 8), f0.device)

    @torch.no_grad()
    def forward_features(self, feats: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        return [t.clone() for t in self.plan_for(feats).forward_logits(feats)]

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        plan = self._fused_plan(x)
        if plan is None:
            return self.forward_features(self.backbone.features(x))
        return [t.clone() for t in plan.forward_image(x.float().contiguous(), False)]

    @torch.no_grad()
    def decode_features(self, feats: Sequence[torch.Tensor]) -> torch.Tensor:
        return self.plan_for(feats).forward_decoded(feats)

    @torch.no_grad()
    def detect_features(self, feats: Sequence[torch.Tensor], conf_thres: float = 0.5, nms_thres: float = 0.4,
                        strategy: str = "auto_cuda", max_det: Optional[int] = None):
        plan = self.plan_for(feats)
        pred = plan.forward_detect(feats)
        key = (plan.B, plan.num_anchors, self.num_classes, max_det, str(plan.device))
        if key not in self._nms:
            self._nms = {key: DeviceNMS(plan.B, plan.num_anchors, self.num_classes, max_det=max_det, device=plan.device)}
        return self._nms[key].launch(pred, conf_thres, nms_thres, strategy, cls_logits=plan.det_cls_logits)
