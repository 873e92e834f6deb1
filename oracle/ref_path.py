"""ORACLE (test infrastructure only) - CPU fp32 restatement of the reference's GLSDet hot path.

This file restates, in plain functional PyTorch on the CPU, what WUTCM-Lab/GLSDet computes on the path
neck -> FFA -> decoupled head -> decode -> score filter -> class-aware NMS, driven by nothing but a
state_dict with the reference's keys.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product (glsdet_b200/) never does.

Pinning: tests/golden/make_golden.py imports the REAL reference from /root/reference (with the two
import-time shims of SURVEY.md section 0.2) in the build container, runs it on seeded inputs and commits the
input/output vectors under tests/golden/; tests/test_oracle_cpu.py checks this restatement against those
vectors.  The reference's own test-suite has no golden vectors for this path.

Every function cites the reference lines it follows (paths relative to /root/reference/yolox-drone/).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import nms_oracle

StateDict = Dict[str, torch.Tensor]
BN_EPS = 1e-3  # models/base/baseConv.py:12


def _act(x: torch.Tensor, act: str) -> torch.Tensor:
    # models/base/activation.py:4-17
    if act == "silu":
        return x * torch.sigmoid(x)
    if act == "relu":
        return torch.relu(x)
    if act == "lrelu":
        return F.leaky_relu(x, 0.1)
    raise AttributeError(f"Unsupported act type: {act}")


_EMULATE_BF16 = False  # see neck_head_bf16(): storage-precision emulation of the native 16-bit path
_EMU_IN_H = None       # network input height while emulating: the stride of a map is _EMU_IN_H / its height
_EMU_STORAGE = "mixed"  # "mixed": bf16 for the head's tensors at stride <= 4, fp16 for everything else (the policy of
                        # glsdet_b200/_native.py::storage_dtype, restated here because the oracle imports nothing from
                        # the product); "bf16" / "f16": one type everywhere
_EMU_ROLE = "head"      # "head" | "backbone": which part of the graph is being emulated


def _emu_dtype(t: torch.Tensor, role: str):
    if _EMU_STORAGE == "bf16":
        return torch.bfloat16
    if _EMU_STORAGE == "f16":
        return torch.float16
    head_fine = role == "head" and (_EMU_IN_H is None or t.dim() != 4 or _EMU_IN_H / t.shape[2] <= 4)
    return torch.bfloat16 if head_fine else torch.float16


def _q(t: torch.Tensor, like: torch.Tensor = None, role: str = None) -> torch.Tensor:
    """Round to the 16-bit storage type of the tensor and back when the storage emulation is on (identity otherwise).
    A tensor that already carries its storage type (attribute `_st`, set here) is returned as is.  `like`: the
    activation a weight tensor multiplies - weights are stored in the type of their conv's input."""
    if not _EMULATE_BF16:
        return t
    if like is None:
        if getattr(t, "_st", None) is not None:
            return t
        dt = _emu_dtype(t, role or _EMU_ROLE)
        out = t.to(dt).float()
        out._st = dt
        return out
    dt = getattr(like, "_st", None) or _emu_dtype(like, role or _EMU_ROLE)
    return t.to(dt).float()


def _qin(feats):
    """Plan inputs (backbone features converted from NCHW fp32): stored in the input type of the policy."""
    return [None if f is None else _q(f, role="input") for f in feats]


class _emulate:
    """with _emulate(on, in_h, storage): ... - scoped storage emulation."""

    def __init__(self, on: bool, in_h=None, storage: str = "mixed", role: str = "head"):
        self.new = (bool(on), in_h, storage, role)

    def __enter__(self):
        global _EMULATE_BF16, _EMU_IN_H, _EMU_STORAGE, _EMU_ROLE
        self.old = (_EMULATE_BF16, _EMU_IN_H, _EMU_STORAGE, _EMU_ROLE)
        _EMULATE_BF16, _EMU_IN_H, _EMU_STORAGE, _EMU_ROLE = self.new

    def __exit__(self, *exc):
        global _EMULATE_BF16, _EMU_IN_H, _EMU_STORAGE, _EMU_ROLE
        _EMULATE_BF16, _EMU_IN_H, _EMU_STORAGE, _EMU_ROLE = self.old


def base_conv(sd: StateDict, p: str, x: torch.Tensor, stride: int = 1, act: str = "silu") -> torch.Tensor:
    """act(bn(conv(x))): models/base/baseConv.py:6-16 (pad=(k-1)//2, conv bias=False, BN eps 1e-3, eval mode)."""
    # DWConv.forward (baseConv.py:22-30; phi = 'nano'): pconv(dconv(x)); the same block under mmcv's names
    # (DepthwiseSeparableConvModule: depthwise_conv / pointwise_conv) for the mmdet face with use_depthwise=True
    for dn, pn in ((".dconv", ".pconv"), (".depthwise_conv", ".pointwise_conv")):
        if (p + dn + ".conv.weight") in sd:
            return base_conv(sd, p + pn, base_conv(sd, p + dn, x, stride, act), 1, act)
    w = sd[p + ".conv.weight"]
    k = w.shape[-1]
    groups = x.shape[1] // w.shape[1]   # 1, or the channel count for the depthwise half of a DWConv (baseConv.py:25)
    if _EMULATE_BF16:
        # same arithmetic as the reference but with the storage precision of the bf16 path: BN folded into the
        # weights, weights and layer outputs rounded to bf16, accumulation and bias in fp32
        scale = sd[p + ".bn.weight"].double() / torch.sqrt(sd[p + ".bn.running_var"].double() + BN_EPS)
        wf = (w.double() * scale.view(-1, 1, 1, 1)).float()
        bf = (sd[p + ".bn.bias"].double() - sd[p + ".bn.running_mean"].double() * scale).float()
        if groups > 1:   # the depthwise kernel keeps its folded weights in fp32 (csrc/dwconv.cu)
            return _q(_act(F.conv2d(_q(x), wf, bf, stride=stride, padding=(k - 1) // 2, groups=groups), act))
        return _q(_act(F.conv2d(_q(x), _q(wf, x), bf, stride=stride, padding=(k - 1) // 2), act))
    y = F.conv2d(x, w, None, stride=stride, padding=(k - 1) // 2, groups=groups)
    y = F.batch_norm(y, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"], sd[p + ".bn.weight"],
                     sd[p + ".bn.bias"], training=False, eps=BN_EPS)
    return _act(y, act)


def _count_blocks(sd: StateDict, p: str) -> int:
    n = 0
    while f"{p}.m.{n}.conv1.conv.weight" in sd:
        n += 1
    return n


def csp_layer(sd: StateDict, p: str, x: torch.Tensor, shortcut: bool = False, act: str = "silu") -> torch.Tensor:
    """models/ffa/darknet.py:66-112 (CSPLayer) with Bottleneck :43-63 (expansion 1.0 inside CSP, :87)."""
    x1 = base_conv(sd, p + ".conv1", x, act=act)
    x2 = base_conv(sd, p + ".conv2", x, act=act)
    for j in range(_count_blocks(sd, p)):
        y = base_conv(sd, f"{p}.m.{j}.conv2", base_conv(sd, f"{p}.m.{j}.conv1", x1, act=act), act=act)
        x1 = y + x1 if shortcut else y  # use_add = shortcut and cin == cout (:57); cin == cout inside CSP
    return base_conv(sd, p + ".conv3", torch.cat((x1, x2), dim=1), act=act)


def _up2(x: torch.Tensor) -> torch.Tensor:
    return F.interpolate(x, scale_factor=2, mode="nearest")


def pafpn_neck(sd: StateDict, feats: Sequence[torch.Tensor], p: str = "backbone") -> List[torch.Tensor]:
    """models/ffa/yolox_ffa.py:196-261 after the backbone call (:197-198): feats = (dark2, dark3, dark4, dark5)."""
    feat0, feat1, feat2, feat3 = feats
    P5 = base_conv(sd, f"{p}.lateral_conv0", feat3)                          # :203
    x = csp_layer(sd, f"{p}.C3_p4", torch.cat([_up2(P5), feat2], 1))         # :207-215
    P4 = base_conv(sd, f"{p}.reduce_conv1", x)                               # :220
    P3_out = csp_layer(sd, f"{p}.C3_p3", torch.cat([_up2(P4), feat1], 1))    # :224-232
    d = base_conv(sd, f"{p}.bu_conv2", P3_out, stride=2)                     # :237
    P4_out = csp_layer(sd, f"{p}.C3_n3", torch.cat([d, P4], 1))              # :241-245
    d = base_conv(sd, f"{p}.bu_conv1", P4_out, stride=2)                     # :250
    P5_out = csp_layer(sd, f"{p}.C3_n4", torch.cat([d, P5], 1))              # :254-258
    return [feat0, P3_out, P4_out, P5_out]                                   # :261


def se_block(sd: StateDict, p: str, x: torch.Tensor) -> torch.Tensor:
    """models/ffa/ffa.py:5-20: x * sigmoid(W2 relu(W1 avgpool(x))), Linear layers without bias."""
    b, c = x.shape[:2]
    y = x.mean(dim=(2, 3))
    y = torch.sigmoid(F.linear(torch.relu(F.linear(y, sd[p + ".fc.0.weight"])), sd[p + ".fc.2.weight"]))
    return x * y.view(b, c, 1, 1)


def ffa(sd: StateDict, p: str, bottom: torch.Tensor, top: torch.Tensor) -> torch.Tensor:
    """models/ffa/ffa.py:74-85 (all five BaseConvs use ReLU, :27-66)."""
    t = base_conv(sd, p + ".scale", top, act="relu")
    t = base_conv(sd, p + ".create_content_extractor.0", t, act="relu")
    t = base_conv(sd, p + ".create_content_extractor.1", t, act="relu")
    t = _q(t + se_block(sd, p + ".se1", t))
    t = F.pixel_shuffle(t, 2)
    b = torch.cat((bottom, t), 1)
    b = base_conv(sd, p + ".create_text_extractor.0", b, act="relu")
    b = base_conv(sd, p + ".conv3", b, act="relu")
    return _q(t + b)


def yolox_head(sd: StateDict, inputs: Sequence[torch.Tensor], p: str = "head") -> List[torch.Tensor]:
    """models/ffa/yolox_ffa.py:58-118: FFA on (P3_out, P4_out), csp on dark2 (+ upsampled FFA output), stems on
    the other three; level 0 uses tower index 3, level k>0 uses k-1 (:84-87); output cat([reg, obj, cls], 1)."""
    zz = ffa(sd, f"{p}.ftt", inputs[1], inputs[2])                           # :66
    proc = []
    for k, x in enumerate(inputs):
        if k == 0:
            proc.append(csp_layer(sd, f"{p}.csp", x))                        # :70
        else:
            proc.append(base_conv(sd, f"{p}.stems.{k - 1}", x))              # :72
    proc[0] = _q(proc[0] + _up2(zz))                                         # :73
    outs = []
    for k, x in enumerate(proc):
        i = 3 if k == 0 else k - 1
        cf = base_conv(sd, f"{p}.cls_convs.{i}.1", base_conv(sd, f"{p}.cls_convs.{i}.0", x))
        cls_out = F.conv2d(cf, _q(sd[f"{p}.cls_preds.{i}.weight"], cf), sd[f"{p}.cls_preds.{i}.bias"])
        rf = base_conv(sd, f"{p}.reg_convs.{i}.1", base_conv(sd, f"{p}.reg_convs.{i}.0", x))
        reg_out = F.conv2d(rf, _q(sd[f"{p}.reg_preds.{i}.weight"], rf), sd[f"{p}.reg_preds.{i}.bias"])
        obj_out = F.conv2d(rf, _q(sd[f"{p}.obj_preds.{i}.weight"], rf), sd[f"{p}.obj_preds.{i}.bias"])
        outs.append(torch.cat([reg_out, obj_out, cls_out], 1))               # :116
    return outs


def stock_head(sd: StateDict, inputs: Sequence[torch.Tensor], p: str = "head") -> List[torch.Tensor]:
    """models/base/yolox.py YOLOXHead.forward: stem -> two 3x3 towers -> 1x1 preds per level, cat([reg, obj, cls])."""
    outs = []
    for k, x in enumerate(inputs):
        x = base_conv(sd, f"{p}.stems.{k}", x)
        cf = base_conv(sd, f"{p}.cls_convs.{k}.1", base_conv(sd, f"{p}.cls_convs.{k}.0", x))
        cls_out = F.conv2d(cf, _q(sd[f"{p}.cls_preds.{k}.weight"], cf), sd[f"{p}.cls_preds.{k}.bias"])
        rf = base_conv(sd, f"{p}.reg_convs.{k}.1", base_conv(sd, f"{p}.reg_convs.{k}.0", x))
        reg_out = F.conv2d(rf, _q(sd[f"{p}.reg_preds.{k}.weight"], rf), sd[f"{p}.reg_preds.{k}.bias"])
        obj_out = F.conv2d(rf, _q(sd[f"{p}.obj_preds.{k}.weight"], rf), sd[f"{p}.obj_preds.{k}.bias"])
        outs.append(torch.cat([reg_out, obj_out, cls_out], 1))
    return outs


def stock_neck_head(sd: StateDict, feats: Sequence[torch.Tensor], bf16: bool = False,
                    storage: str = "mixed") -> List[torch.Tensor]:
    """models/base/yolox.py YoloBody.forward minus CSPDarknet: feats = (dark3, dark4, dark5).  The neck is the same
    PAFPN as yolox_ffa.py (the FFA variant only adds dark2 as a pass-through fourth input)."""
    with _emulate(bf16, 8 * feats[0].shape[2], storage), torch.no_grad():
        f = _qin(feats)
        neck = pafpn_neck(sd, [None] + list(f))
        return stock_head(sd, neck[1:])


def neck_head(sd: StateDict, feats: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """YoloBody.forward minus the CSPDarknet call (yolox_ffa.py:275-284)."""
    with torch.no_grad():
        return yolox_head(sd, pafpn_neck(sd, feats))


def neck_bf16(sd: StateDict, feats: Sequence[torch.Tensor], storage: str = "mixed") -> List[torch.Tensor]:
    """pafpn_neck with the storage precision of the 16-bit path (see neck_head_bf16)."""
    with _emulate(True, 4 * feats[0].shape[2], storage), torch.no_grad():
        return pafpn_neck(sd, _qin(feats))


def neck_head_bf16(sd: StateDict, feats: Sequence[torch.Tensor], storage: str = "mixed") -> List[torch.Tensor]:
    """The same graph evaluated with the STORAGE precision of the native 16-bit path (inputs, folded weights and
    every layer output rounded to the storage type of their level - bf16 at stride 4, fp16 at the coarser levels for
    storage="mixed", the product's default; "bf16" / "f16" = one type everywhere; fp32 accumulation).  Separates two
    error sources in the tests: the CUDA path must match this closely (kernel correctness), while its distance to
    neck_head() is the quantisation error that BASELINE.json bounds by 2e-2."""
    with _emulate(True, 4 * feats[0].shape[2], storage), torch.no_grad():
        return yolox_head(sd, pafpn_neck(sd, _qin(feats)))


# ------------------------------------------------------------------------------------------ P1: models/new/yolox10.py
def non_local_block(sd: StateDict, p: str, x: torch.Tensor) -> torch.Tensor:
    """models/new/Non_local_family.py:32-48 (dot_product mode, :27-30): g / theta / phi / conv_out are 1x1 convs WITH
    bias and no BN; pairwise = theta^T phi / T with T = H*W (the divisor is pairwise.shape[-1], :29); no softmax."""
    n, _, h, w = x.shape
    ci = sd[p + ".g.weight"].shape[0]
    g_x = F.conv2d(_q(x), _q(sd[p + ".g.weight"], x), sd[p + ".g.bias"]).view(n, ci, -1).permute(0, 2, 1)          # [N, T, Ci]
    theta_x = F.conv2d(_q(x), _q(sd[p + ".theta.weight"], x), sd[p + ".theta.bias"]).view(n, ci, -1).permute(0, 2, 1)
    phi_x = F.conv2d(_q(x), _q(sd[p + ".phi.weight"], x), sd[p + ".phi.bias"]).view(n, ci, -1)                     # [N, Ci, T]
    pairwise = torch.matmul(theta_x, phi_x)
    pairwise = pairwise / pairwise.shape[-1]                                                                    # :29
    y = torch.matmul(pairwise, g_x).permute(0, 2, 1).reshape(n, ci, h, w)                                       # :44-45
    return x + F.conv2d(y, _q(sd[p + ".conv_out.weight"], x), sd[p + ".conv_out.bias"])                            # :46


def patch_conv_nonlocal_new(sd: StateDict, p: str, x: torch.Tensor) -> torch.Tensor:
    """models/new/Non_local_family.py:229-250: 2x2 patch split at int(H/2), int(W/2) (:230-233), one non-local block
    per patch, re-tile (:240-247), then channel_conv = 3x3 BaseConv + SiLU (channel_cat='non_linear', :224-227)."""
    h2, w2 = int(x.shape[2] / 2), int(x.shape[3] / 2)
    lt = non_local_block(sd, p + ".feat_patchconv_lt_nonlocal", x[:, :, :h2, :w2])
    lb = non_local_block(sd, p + ".feat_patchconv_lb_nonlocal", x[:, :, h2:, :w2])
    rt = non_local_block(sd, p + ".feat_patchconv_rt_nonlocal", x[:, :, :h2, w2:])
    rb = non_local_block(sd, p + ".feat_patchconv_rb_nonlocal", x[:, :, h2:, w2:])
    y = torch.cat((torch.cat((lt, rt), dim=3), torch.cat((lb, rb), dim=3)), dim=2)
    return base_conv(sd, p + ".channel_conv", _q(y))


def p1_neck(sd: StateDict, feats: Sequence[torch.Tensor], p: str = "backbone") -> List[torch.Tensor]:
    """models/new/yolox10.py:259-335 after the backbone call: feat_k + Patch_conv_feat_k(feat_k) for dark3..dark5
    (:262-266), then the same PAFPN as pafpn_neck; returns (feat0, P3_out, P4_out, P5_out) (:335)."""
    feat0, feat1, feat2, feat3 = feats
    feat1 = _q(feat1 + patch_conv_nonlocal_new(sd, f"{p}.Patch_conv_feat1", feat1))
    feat2 = _q(feat2 + patch_conv_nonlocal_new(sd, f"{p}.Patch_conv_feat2", feat2))
    feat3 = _q(feat3 + patch_conv_nonlocal_new(sd, f"{p}.Patch_conv_feat3", feat3))
    return pafpn_neck(sd, [feat0, feat1, feat2, feat3], p)


def p1_head(sd: StateDict, inputs: Sequence[torch.Tensor], p: str = "head") -> List[torch.Tensor]:
    """models/new/yolox10.py:70-158: csp_feat0 on dark2, stems on (P3_out, P4_out, P5_out); the cls branch of level k
    sees cat([x_k, up_convs[k](level below), upsample(x_{k+1})]) (:93-99; the top level has no upsampled input); the
    reg branch sees x_k alone (:139); output cat([reg, obj, cls], 1) (:156)."""
    f0 = csp_layer(sd, f"{p}.csp_feat0", inputs[0])                                                    # :83
    xs = [base_conv(sd, f"{p}.stems.{k}", inputs[k + 1]) for k in range(3)]                            # :86
    outs = []
    for k, x in enumerate(xs):
        lower = f0 if k == 0 else xs[k - 1]
        down = base_conv(sd, f"{p}.up_convs.{k}.1", base_conv(sd, f"{p}.up_convs.{k}.0", lower), stride=2)
        parts = [x, down] + ([_up2(xs[k + 1])] if k < 2 else [])
        cf = torch.cat(parts, 1)
        cf = base_conv(sd, f"{p}.cls_convs.{k}.1", base_conv(sd, f"{p}.cls_convs.{k}.0", cf))
        cls_out = F.conv2d(cf, _q(sd[f"{p}.cls_preds.{k}.weight"], cf), sd[f"{p}.cls_preds.{k}.bias"])
        rf = base_conv(sd, f"{p}.reg_convs.{k}.1", base_conv(sd, f"{p}.reg_convs.{k}.0", x))
        reg_out = F.conv2d(rf, _q(sd[f"{p}.reg_preds.{k}.weight"], rf), sd[f"{p}.reg_preds.{k}.bias"])
        obj_out = F.conv2d(rf, _q(sd[f"{p}.obj_preds.{k}.weight"], rf), sd[f"{p}.obj_preds.{k}.bias"])
        outs.append(torch.cat([reg_out, obj_out, cls_out], 1))
    return outs


def p1_neck_head(sd: StateDict, feats: Sequence[torch.Tensor], bf16: bool = False,
                 storage: str = "mixed") -> List[torch.Tensor]:
    """models/new/yolox10.py YoloBody.forward (:339-345) minus the CSPDarknet call; `bf16` = storage-precision
    emulation as in neck_head_bf16."""
    with _emulate(bf16, 4 * feats[0].shape[2], storage), torch.no_grad():
        return p1_head(sd, p1_neck(sd, _qin(feats)))


# ------------------------------------------------------------------------------------------ P2: yolo_patch_nonlocal_plus.py
def _patch_seams(sd: StateDict, p: str, lt, lb, rt, rb) -> torch.Tensor:
    """Seam convs + re-tiling shared by Patch_Conv and Patch_Conv_NonLocal
    (models/block/non_local/Identity_Conv.py:303-316 / 369-382): 3x3 BaseConvs on the left / right halves
    (cat along H) and the top / bottom halves (cat along W), lr = cat(l, r; W), tb = cat(t, b; H), channel cat,
    channel_conv = plain 1x1 nn.Conv2d with bias (channel_cat='linear', :287-288)."""
    l = base_conv(sd, p + ".feat_patchconv_l", torch.cat((lt, lb), dim=2))
    r = base_conv(sd, p + ".feat_patchconv_r", torch.cat((rt, rb), dim=2))
    t = base_conv(sd, p + ".feat_patchconv_t", torch.cat((lt, rt), dim=3))
    b = base_conv(sd, p + ".feat_patchconv_b", torch.cat((lb, rb), dim=3))
    y = torch.cat((torch.cat((l, r), dim=3), torch.cat((t, b), dim=2)), dim=1)
    return _q(F.conv2d(y, _q(sd[p + ".channel_conv.weight"], y), sd[p + ".channel_conv.bias"]))


def patch_conv(sd: StateDict, p: str, x: torch.Tensor, stride: int, nonlocal_: bool) -> torch.Tensor:
    """Patch_Conv.forward (Identity_Conv.py:292-318) / Patch_Conv_NonLocal.forward (:353-384): 2x2 split at
    int(H/2), int(W/2), one 3x3 BaseConv (stride `stride`) per patch, [one dot-product Non_local_Block per patch],
    seam convs, re-tile, 1x1."""
    h2, w2 = int(x.shape[2] / 2), int(x.shape[3] / 2)
    parts = {"lt": x[:, :, :h2, :w2], "lb": x[:, :, h2:, :w2], "rt": x[:, :, :h2, w2:], "rb": x[:, :, h2:, w2:]}
    out = {}
    for k, v in parts.items():
        y = base_conv(sd, f"{p}.feat_patchconv_{k}", v, stride=stride)
        if nonlocal_:
            y = _q(non_local_block(sd, f"{p}.feat_patchconv_{k}_nonlocal", y))
        out[k] = y
    return _patch_seams(sd, p, out["lt"], out["lb"], out["rt"], out["rb"])


def identity_conv(sd: StateDict, p: str, x: torch.Tensor) -> torch.Tensor:
    """Identity_Conv_{three,five,seven}.forward (Identity_Conv.py:27-84): a plain k x k nn.Conv2d with bias and
    padding k // 2 (the identity initialisation only matters for training from scratch)."""
    w = sd[p + ".conv.weight"]
    return _q(F.conv2d(_q(x), _q(w, x), sd[p + ".conv.bias"], padding=w.shape[-1] // 2))


def p2_neck(sd: StateDict, feats: Sequence[torch.Tensor], p: str = "backbone") -> List[torch.Tensor]:
    """models/block/non_local/yolo_patch_nonlocal_plus.py:180-247 after the backbone call: feats = (dark3, dark4, dark5)."""
    feat1, feat2, feat3 = feats
    feat1_patch = patch_conv(sd, f"{p}.Patch_conv_feat1", feat1, stride=2, nonlocal_=True)     # :184
    feat2_patch = patch_conv(sd, f"{p}.Patch_conv_feat2", feat2, stride=1, nonlocal_=False)    # :185
    P5 = base_conv(sd, f"{p}.lateral_conv0", feat3)
    x = csp_layer(sd, f"{p}.C3_p4", torch.cat([_up2(P5), feat2, feat1_patch], 1))             # :197-201
    P4 = base_conv(sd, f"{p}.reduce_conv1", x)
    P3_out = csp_layer(sd, f"{p}.C3_p3", torch.cat([_up2(P4), feat1], 1))
    P3_out = identity_conv(sd, f"{p}.P3_Identity", P3_out)                                      # :219
    d = base_conv(sd, f"{p}.bu_conv2", P3_out, stride=2)
    P4_out = csp_layer(sd, f"{p}.C3_n3", torch.cat([d, P4, feat2_patch], 1))                   # :228-232
    P4_out = identity_conv(sd, f"{p}.P4_Identity", P4_out)                                      # :233
    d = base_conv(sd, f"{p}.bu_conv1", P4_out, stride=2)
    P5_out = csp_layer(sd, f"{p}.C3_n4", torch.cat([d, P5], 1))
    P5_out = identity_conv(sd, f"{p}.P5_Identity", P5_out)                                      # :246
    return [P3_out, P4_out, P5_out]


def p2_neck_head(sd: StateDict, feats: Sequence[torch.Tensor], bf16: bool = False,
                 storage: str = "mixed") -> List[torch.Tensor]:
    """yolo_patch_nonlocal_plus.py YoloBody.forward minus the CSPDarknet call; the head (:6-147) is the stock
    three-level decoupled head."""
    with _emulate(bf16, 8 * feats[0].shape[2], storage), torch.no_grad():
        return stock_head(sd, p2_neck(sd, _qin(feats)))


# ------------------------------------------------------------------------------------------ backbone (upstream)
def csp_darknet(sd: StateDict, x: torch.Tensor, p: str = "backbone.backbone") -> List[torch.Tensor]:
    """models/ffa/darknet.py:10-37,115-195 (Focus :15-21, SPPBottleneck :33-37, CSPDarknet.forward :172-195).
    SURVEY.md section 8f row 1; pinned to tests/golden/backbone_s.npz (recorded from the real reference)."""
    with torch.no_grad():
        x = torch.cat((x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]), dim=1)
        x = base_conv(sd, f"{p}.stem.conv", x)
        outs = []
        for name in ("dark2", "dark3", "dark4", "dark5"):
            x = base_conv(sd, f"{p}.{name}.0", x, stride=2)
            if name == "dark5":
                x = base_conv(sd, f"{p}.{name}.1.conv1", x)
                x = torch.cat([x] + [F.max_pool2d(x, ks, 1, ks // 2) for ks in (5, 9, 13)], dim=1)
                x = base_conv(sd, f"{p}.{name}.1.conv2", x)
                x = csp_layer(sd, f"{p}.{name}.2", x, shortcut=False)
            else:
                x = csp_layer(sd, f"{p}.{name}.1", x, shortcut=True)
            outs.append(x)
        return outs


def csp_darknet_bf16(sd: StateDict, x: torch.Tensor, p: str = "backbone.backbone",
                     storage: str = "mixed") -> List[torch.Tensor]:
    """csp_darknet with the storage precision of the 16-bit path (see neck_head_bf16)."""
    with _emulate(True, x.shape[2], storage, role="backbone"):
        return csp_darknet(sd, x, p)   # the Focus kernel rounds the space-to-depth image (base_conv quantises its input)


# ------------------------------------------------------------------------------------------ decode
def decode_outputs(outputs: Sequence[torch.Tensor], input_shape: Sequence[int]) -> torch.Tensor:
    """models/core/utils_bbox.py:254-306.  Returns a contiguous [B, A, 5+nc] tensor (the reference returns the
    same values as a permuted view)."""
    hw = [o.shape[-2:] for o in outputs]
    out = torch.cat([o.flatten(start_dim=2) for o in outputs], dim=2).permute(0, 2, 1).contiguous()
    out[:, :, 4:] = torch.sigmoid(out[:, :, 4:])
    grids, strides = [], []
    for h, w in hw:
        gy, gx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        grids.append(torch.stack((gx, gy), 2).view(1, -1, 2))
        strides.append(torch.full((1, h * w, 1), input_shape[0] / h))       # :285 (H-stride for both axes)
    grids = torch.cat(grids, dim=1).to(out.dtype)
    strides = torch.cat(strides, dim=1).to(out.dtype)
    out[..., :2] = (out[..., :2] + grids) * strides
    out[..., 2:4] = torch.exp(out[..., 2:4]) * strides
    out[..., [0, 2]] = out[..., [0, 2]] / input_shape[1]
    out[..., [1, 3]] = out[..., [1, 3]] / input_shape[0]
    return out


# ------------------------------------------------------------------------------------------ post-processing
def decode_outputs_variant(outputs: Sequence[torch.Tensor], input_shape: Sequence[int], variant: str) -> torch.Tensor:
    """The decode variants of models/core/utils_bbox.py:36-251 that yolo.py:75-82 dispatches on decode_mode:
    'no_sigmoid' (:149-200, sigmoid on obj only), 'no_sigmoid_all' (:202-251, no sigmoid), 'cls_sigmoid' (:95-147,
    sigmoid on the classes only), 'xyxy' (:36-93: no sigmoid, no normalisation, corner boxes in input pixels)."""
    hw = [tuple(x.shape[-2:]) for x in outputs]
    out = torch.cat([x.flatten(start_dim=2) for x in outputs], dim=2).permute(0, 2, 1).contiguous().clone()
    if variant == "no_sigmoid":
        out[:, :, 4] = torch.sigmoid(out[:, :, 4])
    elif variant == "cls_sigmoid":
        out[:, :, 5:] = torch.sigmoid(out[:, :, 5:])
    elif variant not in ("no_sigmoid_all", "xyxy"):
        raise ValueError(variant)
    grids, strides = [], []
    for h, w in hw:
        gy, gx = torch.meshgrid([torch.arange(h), torch.arange(w)], indexing="ij")
        grid = torch.stack((gx, gy), 2).view(1, -1, 2)
        grids.append(grid)
        strides.append(torch.full((1, grid.shape[1], 1), input_shape[0] / h))
    grids = torch.cat(grids, dim=1).type(out.type())
    strides = torch.cat(strides, dim=1).type(out.type())
    out[..., :2] = (out[..., :2] + grids) * strides
    out[..., 2:4] = torch.exp(out[..., 2:4]) * strides
    if variant == "xyxy":
        c = out.clone()
        out[..., 0] = c[..., 0] - c[..., 2] / 2
        out[..., 1] = c[..., 1] - c[..., 3] / 2
        out[..., 2] = c[..., 0] + c[..., 2] / 2
        out[..., 3] = c[..., 1] + c[..., 3] / 2
        return out
    out[..., [0, 2]] = out[..., [0, 2]] / input_shape[1]
    out[..., [1, 3]] = out[..., [1, 3]] / input_shape[0]
    return out


def yolo_correct_boxes(box_xy, box_wh, input_shape, image_shape, letterbox_image):
    """models/core/utils_bbox.py:8-33 (numpy, float64 intermediates exactly as numpy promotes them)."""
    box_yx = box_xy[..., ::-1]
    box_hw = box_wh[..., ::-1]
    input_shape = np.array(input_shape)
    image_shape = np.array(image_shape)
    if letterbox_image:
        new_shape = np.round(image_shape * np.min(input_shape / image_shape))
        offset = (input_shape - new_shape) / 2.0 / input_shape
        scale = input_shape / new_shape
        box_yx = (box_yx - offset) * scale
        box_hw = (box_hw * scale).astype(box_hw.dtype)  # the reference multiplies in place (:25): stays float32
    box_mins = box_yx - (box_hw / 2.0)
    box_maxes = box_yx + (box_hw / 2.0)
    boxes = np.concatenate([box_mins[..., 0:1], box_mins[..., 1:2], box_maxes[..., 0:1], box_maxes[..., 1:2]], axis=-1)
    boxes = boxes * np.concatenate([image_shape, image_shape], axis=-1)
    return boxes


def filter_candidates(prediction: torch.Tensor, num_classes: int, conf_thres: float):
    """utils_bbox.py:381-412 for the whole batch: cxcywh -> xyxy, class max, score threshold, compaction.
    Returns per image (det [K,7] float32 numpy, anchor_index [K] int64)."""
    pred = prediction.clone()
    corner = pred.new_empty(pred.shape)
    corner[:, :, 0] = pred[:, :, 0] - pred[:, :, 2] / 2
    corner[:, :, 1] = pred[:, :, 1] - pred[:, :, 3] / 2
    corner[:, :, 2] = pred[:, :, 0] + pred[:, :, 2] / 2
    corner[:, :, 3] = pred[:, :, 1] + pred[:, :, 3] / 2
    pred[:, :, :4] = corner[:, :, :4]
    res = []
    for image_pred in pred:
        class_conf, class_pred = torch.max(image_pred[:, 5:5 + num_classes], 1, keepdim=True)
        mask = (image_pred[:, 4] * class_conf[:, 0] >= conf_thres)
        det = torch.cat((image_pred[:, :5], class_conf, class_pred.float()), 1)[mask]
        res.append((det.numpy().astype(np.float32), torch.nonzero(mask)[:, 0].numpy()))
    return res


def non_max_suppression(prediction: torch.Tensor, num_classes: int, input_shape, image_shape, letterbox_image,
                        conf_thres: float = 0.5, nms_thres: float = 0.4, strategy: str = "auto_cpu",
                        correct_boxes: bool = True, return_index: bool = False):
    """utils_bbox.py:375-484.  `strategy` names which torchvision batched_nms branch is restated (see
    nms_oracle.batched_nms).  With correct_boxes=False rows stay (x1,y1,x2,y2,obj,cls_conf,cls) in network
    coordinates (what the device produces before the host-side yolo_correct_boxes)."""
    output = []
    index = []
    for det, anchor_idx in filter_candidates(prediction, num_classes, conf_thres):
        keep = nms_oracle.batched_nms(det[:, :4], det[:, 4] * det[:, 5], det[:, 6], nms_thres, strategy)
        out = det[keep]
        index.append(anchor_idx[keep])
        if correct_boxes:
            box_xy, box_wh = (out[:, 0:2] + out[:, 2:4]) / 2, out[:, 2:4] - out[:, 0:2]
            out[:, :4] = yolo_correct_boxes(box_xy, box_wh, input_shape, image_shape, letterbox_image)
        output.append(out)
    return (output, index) if return_index else output


# ------------------------------------------------------------------------------------------ synthetic weights
# The seeded weight generator is plain data generation (no algorithm of the path); it lives in the package so that
# bench.py's product arm can use it without touching oracle/.
from glsdet_b200.synthetic import p0_state_dict_shapes, synthetic_state_dict  # noqa: E402,F401
