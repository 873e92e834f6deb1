"""Drop-in for yolox-drone/models/block/non_local/yolo_patch_nonlocal_plus.py (GLSDet "P2"): `YoloBody(num_classes,
phi)`, `YOLOPAFPN`, `YOLOXHead`, plus the parameter holders of models/block/non_local/Identity_Conv.py
(`Patch_Conv`, `Patch_Conv_NonLocal`, `Non_local_Block`, `Identity_Conv_{three,five,seven}`).

Same constructors, forward signatures (NCHW fp32), level order (strides 8, 16, 32) and state_dict keys (600 for
phi='s').  The math runs in the native plan (engine.FFAPathPlan, variant "p2")."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn as nn

from .yolox10 import Non_local_Block
from .yolox_ffa import _DEPTH, _WIDTH, BaseConv, CSPDarknet, CSPLayer, DWConv, _PlanOwner
from .yolox_ffa import YoloBody as _FFAYoloBody

_NATIVE = "executed by libglsdet_b200.so through its parent YOLOPAFPN / YoloBody"


class _IdentityConv(nn.Module):
    """Identity_Conv.py:27-84 parameter layout: `conv` = k x k nn.Conv2d with bias, padding k // 2, initialised to the
    identity map (centre tap 1) exactly like the reference."""

    def __init__(self, channels, k):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, k, 1, k // 2)
        with torch.no_grad():
            self.conv.weight.zero_()
            self.conv.bias.zero_()
            idx = torch.arange(channels)
            self.conv.weight[idx, idx, k // 2, k // 2] = 1.0

    def forward(self, x):
        raise RuntimeError("Identity_Conv is " + _NATIVE)


class Identity_Conv_three(_IdentityConv):
    def __init__(self, in_channels=3, out_channels=3, kernel_size=3, stride=1, padding=1, groups=1):
        super().__init__(in_channels, 3)


class Identity_Conv_five(_IdentityConv):
    def __init__(self, in_channels=3, out_channels=3, kernel_size=5, stride=1, padding=2, groups=1):
        super().__init__(in_channels, 5)


class Identity_Conv_seven(_IdentityConv):
    def __init__(self, in_channels=3, out_channels=3, kernel_size=7, stride=1, padding=3, groups=1):
        super().__init__(in_channels, 7)


class Patch_Conv(nn.Module):
    """Identity_Conv.py:269-318 parameter layout (channel_cat='linear': plain 1x1 conv with bias)."""

    _nonlocal = False

    def __init__(self, in_channel=256, out_channel=512, channel_scale=0.5, patch_scale=2, stride=2, act="silu",
                 channel_cat="linear"):
        super().__init__()
        if channel_cat != "linear":
            raise NotImplementedError("only channel_cat='linear' (the default, used by yolo_patch_nonlocal_plus.py)")
        mid = int(channel_scale * in_channel)
        self.stride = stride
        for pos in ("lt", "lb", "rt", "rb"):
            setattr(self, f"feat_patchconv_{pos}", BaseConv(in_channel, mid, 3, stride, act=act))
        if self._nonlocal:
            for pos in ("lt", "lb", "rt", "rb"):
                setattr(self, f"feat_patchconv_{pos}_nonlocal", Non_local_Block(mid, mid))
        for pos in ("r", "l", "t", "b"):
            setattr(self, f"feat_patchconv_{pos}", BaseConv(mid, mid, 3, 1, act=act))
        self.channel_conv = nn.Conv2d(2 * mid, out_channel, 1, 1)

    def forward(self, x):
        raise RuntimeError(type(self).__name__ + " is " + _NATIVE)


class Patch_Conv_NonLocal(Patch_Conv):
    """Identity_Conv.py:321-384."""

    _nonlocal = True


class YOLOXHead(_PlanOwner):
    """yolo_patch_nonlocal_plus.py:6-147: the stock decoupled head; forward(inputs = (P3_out, P4_out, P5_out))."""

    _parts = ("head",)
    _variant = "p2"

    def __init__(self, num_classes, width=1.0, in_channels=[256, 512, 1024], act="silu", depthwise=False):
        super().__init__()
        self.num_classes = num_classes
        hc = int(256 * width)
        Conv = DWConv if depthwise else BaseConv   # yolo_patch_nonlocal_plus.py:15
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
    """yolo_patch_nonlocal_plus.py:55-247.  forward(image batch) -> (P3_out, P4_out, P5_out), NCHW fp32."""

    _parts = ("neck",)
    _variant = "p2"

    def __init__(self, depth=1.0, width=1.0, in_features=("dark3", "dark4", "dark5"), in_channels=[256, 512, 1024],
                 depthwise=False, act="silu"):
        super().__init__()
        self.backbone = CSPDarknet(depth, width, out_features=("dark3", "dark4", "dark5"), depthwise=depthwise, act=act)
        self.in_features = in_features
        c0, c1, c2 = (int(c * width) for c in in_channels)
        n = round(3 * depth)
        Conv = DWConv if depthwise else BaseConv   # yolo_patch_nonlocal_plus.py:97
        self.lateral_conv0 = BaseConv(c2, c1, 1, 1, act=act)
        self.C3_p4 = CSPLayer(3 * c1, c1, n, False, depthwise=depthwise, act=act)
        self.reduce_conv1 = BaseConv(c1, c0, 1, 1, act=act)
        self.C3_p3 = CSPLayer(2 * c0, c0, n, False, depthwise=depthwise, act=act)
        self.P3_Identity = Identity_Conv_seven(c0, c0)
        self.bu_conv2 = Conv(c0, c0, 3, 2, act=act)
        self.C3_n3 = CSPLayer(3 * c0, c1, n, False, depthwise=depthwise, act=act)
        self.P4_Identity = Identity_Conv_five(c1, c1)
        self.bu_conv1 = Conv(c1, c1, 3, 2, act=act)
        self.C3_n4 = CSPLayer(2 * c1, c2, n, False, depthwise=depthwise, act=act)
        self.Patch_conv_feat1 = Patch_Conv_NonLocal(in_channel=c0, out_channel=c1, patch_scale=2)
        self.Patch_conv_feat2 = Patch_Conv(in_channel=c1, out_channel=c0, patch_scale=4, stride=1)
        self.P5_Identity = Identity_Conv_three(c2, c2)
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


class YoloBody(_FFAYoloBody):
    """yolo_patch_nonlocal_plus.py:249-264; `detect` / `detect_features` fuse neck -> head -> decode -> filter -> NMS."""

    _variant = "p2"

    def __init__(self, num_classes, phi):
        _PlanOwner.__init__(self)
        depth, width = _DEPTH[phi], _WIDTH[phi]
        depthwise = phi == "nano"
        self.num_classes = num_classes
        self.backbone = YOLOPAFPN(depth, width, depthwise=depthwise)
        self.head = YOLOXHead(num_classes, width, depthwise=depthwise)
        self._nms = {}
        nn.Module.train(self, False)

    def plan_for(self, feats):
        f0 = feats[0]   # dark3: stride 8
        return self._plan(f0.shape[0], (f0.shape[2] * 8, f0.shape[3] * 8), f0.device)
