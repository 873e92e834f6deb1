"""Drop-in for yolox-drone/models/new/yolox10.py (GLSDet "P1"): `YoloBody(num_classes, phi)`, `YOLOPAFPN`, `YOLOXHead`,
plus `Non_local_Block` / `Patch_Conv_NonLocal_new` of models/new/Non_local_family.py.

Same constructors, forward signatures (NCHW fp32 in and out), level order (strides 8, 16, 32) and state_dict keys as
the reference (654 keys for phi='s').  The modules only hold parameters; the math runs in the
native plan (engine.FFAPathPlan, variant "p1"): the patch non-local attention is evaluated in its reassociated form
(engine.FFAPathPlan._build_nonlocal) by the tcgen05 conv kernel with per-image weight matrices.  The CSPDarknet
backbone runs as the native BackbonePlan (glsdet_b200/backbone.py).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn as nn

from .yolox_ffa import _DEPTH, _WIDTH, BaseConv, CSPDarknet, CSPLayer, DWConv, _PlanOwner
from .yolox_ffa import YoloBody as _FFAYoloBody


class Non_local_Block(nn.Module):
    """models/new/Non_local_family.py:6-48 parameter layout: g, theta, phi, conv_out - 1x1 convs with bias."""

    def __init__(self, in_channels, inter_channels=None):
        super().__init__()
        self.in_channels = in_channels
        self.inter_channels = inter_channels if inter_channels is not None else max(in_channels // 2, 1)
        self.g = nn.Conv2d(in_channels, self.inter_channels, 1, 1)
        self.theta = nn.Conv2d(in_channels, self.inter_channels, 1, 1)
        self.phi = nn.Conv2d(in_channels, self.inter_channels, 1, 1)
        self.conv_out = nn.Conv2d(self.inter_channels, in_channels, 1, 1)

    def forward(self, x):
        raise RuntimeError("Non_local_Block is executed by libglsdet_b200.so through its parent YOLOPAFPN / YoloBody")


class Patch_Conv_NonLocal_new(nn.Module):
    """models/new/Non_local_family.py:204-250 parameter layout (channel_cat='non_linear': 3x3 BaseConv)."""

    def __init__(self, in_channel=256, out_channel=512, channel_scale=0.5, patch_scale=2, act="silu",
                 channel_cat="non_linear"):
        super().__init__()
        if channel_cat != "non_linear":
            raise NotImplementedError("only channel_cat='non_linear' (the default, used by yolox10.py) is supported")
        mid = int(channel_scale * in_channel)
        self.feat_patchconv_lt_nonlocal = Non_local_Block(in_channel, mid)
        self.feat_patchconv_lb_nonlocal = Non_local_Block(in_channel, mid)
        self.feat_patchconv_rt_nonlocal = Non_local_Block(in_channel, mid)
        self.feat_patchconv_rb_nonlocal = Non_local_Block(in_channel, mid)
        self.channel_conv = BaseConv(mid, out_channel, 3, 1, act=act)

    def forward(self, x):
        raise RuntimeError("Patch_Conv_NonLocal_new is executed by libglsdet_b200.so through its parent module")


class YOLOXHead(_PlanOwner):
    """models/new/yolox10.py:8-158.  forward(inputs) takes (feat0, P3_out, P4_out, P5_out) and returns three raw
    [B, 5+nc, h, w] maps (strides 8, 16, 32), channel order cat([reg, obj, cls]) (:156)."""

    _parts = ("head",)
    _variant = "p1"

    def __init__(self, num_classes, width=1.0, in_channels=[256, 512, 1024], act="silu", depthwise=False):
        super().__init__()
        self.num_classes = num_classes
        hc = int(256 * width)
        Conv = DWConv if depthwise else BaseConv   # yolox10.py:11
        self.cls_convs, self.reg_convs = nn.ModuleList(), nn.ModuleList()
        self.cls_preds, self.reg_preds, self.obj_preds = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.stems = nn.ModuleList()
        self.csp_feat0 = CSPLayer(int(0.5 * in_channels[0] * width), int(in_channels[0] * width), round(3 * 0.75),
                                  False, depthwise=depthwise, act=act)
        self.up_convs = nn.ModuleList()
        for i, cin in enumerate(in_channels):
            self.stems.append(BaseConv(int(cin * width), hc, 1, 1, act=act))
            self.up_convs.append(nn.Sequential(Conv(hc, hc, 3, 1, act=act), Conv(hc, hc, 3, 2, act=act)))
            m = 2 if i == 2 else 3
            self.cls_convs.append(nn.Sequential(Conv(m * hc, m * hc, 3, 1, act=act), Conv(m * hc, hc, 3, 1, act=act)))
            self.cls_preds.append(nn.Conv2d(hc, num_classes, 1, 1, 0))
            self.reg_convs.append(nn.Sequential(Conv(hc, hc, 3, 1, act=act), Conv(hc, hc, 3, 1, act=act)))
            self.reg_preds.append(nn.Conv2d(hc, 4, 1, 1, 0))
            self.obj_preds.append(nn.Conv2d(hc, 1, 1, 1, 0))
        nn.Module.train(self, False)

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
    """models/new/yolox10.py:161-335: feat_k + Patch_conv_feat_k(feat_k) on dark3..dark5, then the PAFPN.
    forward(image batch) -> (feat0, P3_out, P4_out, P5_out), NCHW fp32."""

    _parts = ("neck",)
    _variant = "p1"

    def __init__(self, depth=1.0, width=1.0, in_features=("dark2", "dark3", "dark4", "dark5"),
                 in_channels=[256, 512, 1024], depthwise=False, act="silu"):
        super().__init__()
        self.backbone = CSPDarknet(depth, width, depthwise=depthwise, act=act)
        self.in_features = in_features
        c0, c1, c2 = (int(c * width) for c in in_channels)
        n = round(3 * depth)
        Conv = DWConv if depthwise else BaseConv   # yolox10.py:165
        self.lateral_conv0 = BaseConv(c2, c1, 1, 1, act=act)
        self.C3_p4 = CSPLayer(2 * c1, c1, n, False, depthwise=depthwise, act=act)
        self.reduce_conv1 = BaseConv(c1, c0, 1, 1, act=act)
        self.C3_p3 = CSPLayer(2 * c0, c0, n, False, depthwise=depthwise, act=act)
        self.bu_conv2 = Conv(c0, c0, 3, 2, act=act)
        self.C3_n3 = CSPLayer(2 * c0, c1, n, False, depthwise=depthwise, act=act)
        self.bu_conv1 = Conv(c1, c1, 3, 2, act=act)
        self.C3_n4 = CSPLayer(2 * c1, c2, n, False, depthwise=depthwise, act=act)
        self.Patch_conv_feat1 = Patch_Conv_NonLocal_new(c0, c0, channel_scale=1, patch_scale=2)
        self.Patch_conv_feat2 = Patch_Conv_NonLocal_new(c1, c1, channel_scale=1, patch_scale=2)
        self.Patch_conv_feat3 = Patch_Conv_NonLocal_new(c2, c2, channel_scale=1, patch_scale=2)
        nn.Module.train(self, False)

    def _num_classes(self):
        return 1  # unused by the neck-only plan

    @torch.no_grad()
    def features(self, x: torch.Tensor) -> List[torch.Tensor]:
        out = self.backbone(x)
        return [out[f] for f in self.in_features]

    @torch.no_grad()
    def forward_features(self, feats: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, ...]:
        f0 = feats[0]
        plan = self._plan(f0.shape[0], (f0.shape[2] * 4, f0.shape[3] * 4), f0.device)
        plan.load_features([t.float() for t in feats])
        plan.run_neck()
        outs = plan.neck_outputs_nchw()
        return (feats[0], outs[1], outs[2], outs[3])

    def forward(self, input: torch.Tensor):
        return self.forward_features(self.features(input))


class YoloBody(_FFAYoloBody):
    """models/new/yolox10.py:338-345; `detect` / `detect_features` run neck -> head -> decode -> filter -> NMS fused."""

    _variant = "p1"

    def __init__(self, num_classes, phi):
        _PlanOwner.__init__(self)
        depth, width = _DEPTH[phi], _WIDTH[phi]
        depthwise = phi == "nano"
        self.num_classes = num_classes
        self.backbone = YOLOPAFPN(depth, width, depthwise=depthwise)
        self.head = YOLOXHead(num_classes, width, depthwise=depthwise)
        self._nms = {}
        nn.Module.train(self, False)
