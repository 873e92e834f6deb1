"""Execution plan of the CSPDarknet backbone (yolox-drone/models/ffa/darknet.py:10-37,115-195; the identical file is
models/new/darknet.py and models/base/darknet.py) - SURVEY.md section 8(f) row 1, the caller side of the hot path.

Image batch [B, 3, H, W] fp32 NCHW -> (dark2, dark3, dark4, dark5) as NHWC bf16 buffers, which are the input buffers of
the neck plan when the two are chained (no NCHW fp32 round trip between backbone and neck):

  * Focus (darknet.py:15-21): one kernel does the space-to-depth and the layout / type change of the image
    (glsdet_focus_nchw_f32_to_nhwc_bf16, 16-channel pixels); the 3x3 stem conv reads it with zero weights on the 4 pad channels;
  * every BaseConv = the tcgen05 implicit-GEMM conv with folded BatchNorm and SiLU in the epilogue;
  * CSPLayer: conv1 | conv2 as one GEMM into the concat buffer, Bottleneck shortcut (darknet.py:59-63) as the bf16
    post-residual of the 3x3 conv written in place, conv3 on the concat buffer;
  * SPPBottleneck (darknet.py:24-37): conv1 writes window 0 of the 4C concat buffer, one kernel writes the 5/9/13 max
    pools into windows 1..3 (cascaded separable 5-wide maxima in shared memory), conv2 reads the buffer.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch

from . import _native as N
from .ops import (ConvOp, ConvOpF32, DepthwiseOp, FocusOp, FoldedView, SppPoolOp, View, fold_bn, fold_kx_pair_weight, fold_kx_weight,
                  pair_bias, pair_conv3_weight, pair_pointwise_weight,
                  nhwc_to_nchw, pair_stride2_weight)

BN_EPS = 1e-3
IMAGENET_MEAN = (0.485, 0.456, 0.406)   # models/core/utils.py:49-50
IMAGENET_STD = (0.229, 0.224, 0.225)
FEATURES = ("dark2", "dark3", "dark4", "dark5")


def backbone_supported(base_channels: int) -> bool:
    """The native plan needs channel counts that are multiples of 8 (TMA strides of 16 bytes): every phi of the reference
    (base 16 for the depthwise 'nano', 24 for 'tiny', 32 for 's', ...)."""
    return base_channels % 8 == 0


class BackbonePlan:
    def __init__(self, state_dict: Dict[str, torch.Tensor], batch: int, input_hw: Sequence[int], device=None,
                 act: str = "silu", prefix: str = "backbone.backbone.", outs: Optional[Dict[str, torch.Tensor]] = None,
                 precision: str = "bf16", storage: Optional[str] = None):
        """`outs` optionally maps feature names ("dark2".."dark5") to existing NHWC tensors [B, H/s, W/s, C] (the input
        buffers of a neck plan); missing ones are allocated here.  `precision` "fp32" = the accuracy mode (every tensor
        fp32, SIMT fp32 convs of csrc/fp32_path.cu, fp32 Focus / pooling kernels)."""
        assert precision in ("bf16", "fp32")
        self.fp32 = precision == "fp32"
        self.storage = storage   # 16-bit storage policy (see _native.storage_dtype); None = GLSDET_STORAGE / "mixed"
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        self.B = batch
        self.in_h, self.in_w = int(input_hw[0]), int(input_hw[1])
        if self.in_h % 32 or self.in_w % 32:
            raise ValueError("input size must be a multiple of 32 (yolox-drone/yolo.py:34-36)")
        self.act = N.ACT_BY_NAME[act]
        self.sd = {k[len(prefix):]: v.detach().cpu() for k, v in state_dict.items()   # host side: see engine.FFAPathPlan
                   if k.startswith(prefix) and not k.endswith("num_batches_tracked")}
        if "stem.conv.conv.weight" not in self.sd:
            raise KeyError(f"no CSPDarknet weights under prefix {prefix!r}")
        self.base = self.sd["stem.conv.conv.weight"].shape[0]
        if not backbone_supported(self.base):
            raise NotImplementedError(f"CSPDarknet base width {self.base} is not a multiple of 8")
        self.ops: List = []
        self.pair_csp_blocks: List[str] = []   # CSPLayers built in pixel-pair form (_csp_pairs)
        self.flops = 0.0
        self._bufs: Dict[str, torch.Tensor] = {}
        self.outs: Dict[str, torch.Tensor] = dict(outs or {})
        self._build()

    # ------------------------------------------------------------------ helpers
    def _dt(self, stride: int):
        return torch.float32 if self.fp32 else N.storage_dtype(stride, self.storage, role="backbone")

    def _buf(self, name: str, stride: int, channels: int) -> torch.Tensor:
        t = torch.empty((self.B, self.in_h // stride, self.in_w // stride, channels), dtype=self._dt(stride), device=self.device)
        self._bufs[name] = t
        return t

    def _folded(self, p: str):
        sd = self.sd
        return fold_bn(sd[p + ".conv.weight"], sd[p + ".bn.weight"], sd[p + ".bn.bias"], sd[p + ".bn.running_mean"],
                       sd[p + ".bn.running_var"], BN_EPS)

    def _conv(self, w, b, srcs, out, stride=1, **kw):
        cls = ConvOpF32 if self.fp32 else ConvOp
        op = cls(srcs, w, b, ksize=w.shape[-2], stride=stride, act=self.act, out=out, **kw)
        self.ops.append(op)
        self.flops += op.flops
        return op

    def _is_dw(self, p: str) -> bool:
        """`p` names a DWConv (models/base/baseConv.py:22-30; phi = 'nano', darknet.py:48,120)."""
        return (p + ".dconv.conv.weight") in self.sd

    def _base_conv(self, p, srcs, out, stride=1, **kw):
        if self._is_dw(p):   # DWConv.forward: pconv(dconv(x)); the depthwise half is csrc/dwconv.cu
            assert len(srcs) == 1
            src = srcs[0]
            w, b = self._folded(p + ".dconv")
            bb, h, w_ = src.bhw
            t = torch.empty((bb, h // stride, w_ // stride, src.c), dtype=src.t.dtype, device=self.device)
            self._bufs[p + ".dconv"] = t
            op = DepthwiseOp(src, w, b, stride=stride, act=self.act, out=View(t))
            self.ops.append(op)
            self.flops += op.flops
            return self._base_conv(p + ".pconv", [View(t)], out, 1, **kw)
        w, b = self._folded(p)
        return self._conv(w, b, srcs, out, stride, **kw)

    def _csp(self, p: str, stride: int, src: View, out: View, shortcut: bool):
        """CSPLayer.forward (darknet.py:91-112) with Bottleneck.forward (:59-63)."""
        w1, b1 = self._folded(p + ".conv1")
        w2, b2 = self._folded(p + ".conv2")
        hid = w1.shape[0]
        if self._pair_csp_ok(p, hid, src, out, shortcut):
            return self._csp_pairs(p, stride, src, out, (w1, b1, w2, b2))
        X = self._buf(p + ".X", stride, 2 * hid)
        self._conv(torch.cat([w1, w2], 0), torch.cat([b1, b2], 0), [src], View(X))
        bb = self._buf(p + ".b", stride, hid)
        j = 0
        while f"{p}.m.{j}.conv1.conv.weight" in self.sd:
            x1 = View(X, 0, hid)
            self._base_conv(f"{p}.m.{j}.conv1", [x1], View(bb))
            if shortcut:   # use_add: in_channels == out_channels always holds inside a CSPLayer (expansion 1.0)
                self._base_conv(f"{p}.m.{j}.conv2", [View(bb)], x1, post_res=x1, post_shift=0)
            else:
                self._base_conv(f"{p}.m.{j}.conv2", [View(bb)], x1)
            j += 1
        self._base_conv(p + ".conv3", [View(X)], out)

    def _pair_csp_ok(self, p: str, hid: int, src: View, out: View, shortcut: bool) -> bool:
        """The pixel-pair form of a whole CSPLayer: 32-channel halves only (dark2 of phi = 's'), whole buffers on both sides."""
        if self.fp32 or not shortcut or hid != 32 or os.environ.get("GLSDET_NO_PAIR_CSP"):
            return False
        if self._is_dw(p + ".m.0.conv2") or src.t.shape[2] % 2 or src.coff or out.coff:
            return False
        return src.c == src.t.shape[3] == 2 * hid and out.c == out.t.shape[3] == 2 * hid

    def _csp_pairs(self, p: str, stride: int, src: View, out: View, w12):
        """CSPLayer with 32-channel halves in PIXEL-PAIR form: every tensor is viewed as [B, H, W/2, 2C] (pair X = pixels 2X,
        2X+1) and every conv becomes a conv over pairs with block-structured weights (ops.pair_*): the 32-channel layers (one
        half-empty 64-channel K chunk per tap, N = 32 direct-store epilogue: 58 + 129 us at 16 x 256^2) turn into 64 -> 64
        layers on half as many GEMM rows with full K chunks and the TMA-store epilogue.  The concat buffer holds
        (x1(p0) | x1(p1) | x2(p0) | x2(p1)) per pair so that the bottleneck's x1 stays one contiguous window; conv3's weight
        columns are permuted to match and its pair output is the natural [B, H, W, 2 * hid] layout again."""
        w1, b1, w2, b2 = w12
        hid = w1.shape[0]
        bsz, h, w_, _ = src.t.shape
        self.pair_csp_blocks.append(p)
        flops0 = self.flops

        def pairs(t):
            return t.view(t.shape[0], t.shape[1], t.shape[2] // 2, 2 * t.shape[3])

        Xp = pairs(self._buf(p + ".X", stride, 2 * hid))
        bbp = pairs(self._buf(p + ".b", stride, hid))
        w12n = torch.cat([w1, w2], 0)
        self._conv(pair_pointwise_weight(w12n, 1, 2), pair_bias(torch.cat([b1, b2], 0), 2), [View(pairs(src.t))], View(Xp))
        ref = 2.0 * bsz * h * w_ * (2 * hid) * (2 * hid)
        x1 = View(Xp, 0, 2 * hid)
        j = 0
        while f"{p}.m.{j}.conv1.conv.weight" in self.sd:
            wa, ba = self._folded(f"{p}.m.{j}.conv1")
            wb, bb_ = self._folded(f"{p}.m.{j}.conv2")
            self._conv(pair_pointwise_weight(wa), pair_bias(ba), [x1], View(bbp))
            self._conv(pair_conv3_weight(wb), pair_bias(bb_), [View(bbp)], x1, post_res=x1, post_shift=0)
            ref += 2.0 * bsz * h * w_ * hid * hid * 10
            j += 1
        w3, b3 = self._folded(p + ".conv3")
        self._conv(pair_pointwise_weight(w3, 2, 1), pair_bias(b3), [View(Xp)], View(pairs(out.t)))
        ref += 2.0 * bsz * h * w_ * (2 * hid) * w3.shape[0]
        self.flops = flops0 + ref   # the reference's FLOPs: the structural zeros of the pair weights are not work

    def _out(self, name: str, stride: int, channels: int) -> torch.Tensor:
        t = self.outs.get(name)
        if t is None:
            t = self._buf(name, stride, channels)
            self.outs[name] = t
        assert tuple(t.shape) == (self.B, self.in_h // stride, self.in_w // stride, channels) and t.dtype == self._dt(stride), \
            (name, tuple(t.shape), t.dtype)
        return t

    # ------------------------------------------------------------------ graph (CSPDarknet.forward, darknet.py:172-195)
    def _build(self):
        c = self.base
        h2, w2 = self.in_h // 2, self.in_w // 2
        w, b = self._folded("stem.conv")                       # [c, 12, 3, 3]
        stem = self._buf("stem", 2, c)
        if self.fp32:                                          # accuracy mode: plain 3x3 conv over the 12 Focus channels
            s2d = self._buf("focus", 2, 12)
            self.focus = FocusOp(s2d)
            self._conv(w, b, [View(s2d)], View(stem))
        elif os.environ.get("GLSDET_STEM_UNFOLDED"):           # diagnostic: plain 3x3 conv over 16-channel pixels (K = 9 * 64)
            s2d = self._buf("focus", 2, 16)
            self.focus = FocusOp(s2d)
            self._conv(torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 4)), b, [View(s2d)], View(stem))
        else:
            # kx taps folded into the channel view (K = 3 * 64): zero-bordered rows of 16-channel pixels, the conv reads
            # 64 consecutive elements = (pixel x-1 | x | x+1 | ignored) per ky tap
            flat = torch.zeros(self.B * h2 * (w2 + 2) * 16 + 64, dtype=self._dt(2), device=self.device)
            self._bufs["focus"] = flat
            fv = FoldedView(flat, self.B, h2, w2, 16)
            self.focus = FocusOp(fv)
            if w2 % 2 == 0 and (2 * c) % 64 == 0 and not os.environ.get("GLSDET_STEM_SINGLE"):
                # pair form: one GEMM row = output pixels (2X, 2X+1); the window at step X covers input pixels 2X-1 .. 2X+2
                # with no padding left in K, N = 2c >= 64 (TMA-store epilogue) and the output is the [.., W/2, 2c] view
                pair = FoldedView(flat, self.B, h2, w2 // 2, 32, row_pitch=(w2 + 2) * 16)
                wp, bp = fold_kx_pair_weight(w, b, 16)
                self._conv(wp, bp, [pair], View(stem.view(self.B, h2, w2 // 2, 2 * c)), ksize_w=1)
            else:
                self._conv(fold_kx_weight(w, 16), b, [fv], View(stem), ksize_w=1)
        self.ops[-1].flops = 2.0 * self.B * h2 * w2 * c * 12 * 9   # the reference's FLOPs (12 channels, 9 taps)
        self.flops = self.ops[-1].flops
        x = stem
        for name, stride, cout in (("dark2", 4, 2 * c), ("dark3", 8, 4 * c), ("dark4", 16, 8 * c)):
            t = self._buf(name + ".0", stride, cout)
            cin = x.shape[3]
            maxc = int(os.environ.get("GLSDET_PAIR_STRIDE2_MAXC", "32"))
            if (not self.fp32 and cin % 32 == 0 and cin <= maxc and x.shape[2] % 2 == 0 and not self._is_dw(f"{name}.0")
                    and not os.environ.get("GLSDET_NO_PAIR_STRIDE2")):
                # 32 input channels: the pixel-pair form (K = 6 x 64 instead of 9 x 64 half-empty chunks, dense TMA boxes)
                w_, b_ = self._folded(f"{name}.0")
                xp = x.view(x.shape[0], x.shape[1], x.shape[2] // 2, 2 * cin)
                self._conv(pair_stride2_weight(w_), b_, [View(xp)], View(t), stride=2, ksize_w=2)
                self.ops[-1].flops = 2.0 * self.B * t.shape[1] * t.shape[2] * cout * cin * 9
                self.flops += self.ops[-1].flops - 2.0 * self.B * t.shape[1] * t.shape[2] * cout * 2 * cin * 6
            else:
                self._base_conv(f"{name}.0", [View(x)], View(t), stride=2)
            o = self._out(name, stride, cout)
            self._csp(f"{name}.1", stride, View(t), View(o), shortcut=True)
            x = o
        t = self._buf("dark5.0", 32, 16 * c)
        self._base_conv("dark5.0", [View(x)], View(t), stride=2)
        hid = 8 * c
        cat = self._buf("spp_cat", 32, 4 * hid)
        self._base_conv("dark5.1.conv1", [View(t)], View(cat, 0, hid))
        self.ops.append(SppPoolOp(cat, hid))
        t2 = self._buf("spp_out", 32, 16 * c)
        self._base_conv("dark5.1.conv2", [View(cat)], View(t2))
        self._csp("dark5.2", 32, View(t2), View(self._out("dark5", 32, 16 * c)), shortcut=False)

    # ------------------------------------------------------------------ execution
    def run(self, image: torch.Tensor, stream=None) -> None:
        self.focus.launch(image, stream)
        for op in self.ops:
            op.launch(stream)

    def run_uint8(self, image_u8: torch.Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD, stream=None) -> None:
        """Same from a uint8 HWC batch [B, H, W, 3] (letterboxed to the network size): preprocess_input
        (models/core/utils.py:47-51) and the HWC -> CHW transpose (yolo.py:134) are fused into the Focus kernel."""
        if self.fp32:
            raise NotImplementedError("the uint8 entry point exists for bf16 plans")
        self.focus.launch_u8(image_u8, mean, std, stream)
        for op in self.ops:
            op.launch(stream)

    def num_launches(self) -> int:
        return 1 + len(self.ops)

    def features_nchw(self, names: Sequence[str] = FEATURES, stream=None) -> Dict[str, torch.Tensor]:
        """The dict CSPDarknet.forward returns (NCHW fp32)."""
        res = {}
        for n in names:
            v = View(self.outs[n]) if n in self.outs else View(self._bufs[n])
            b, h, w = v.bhw
            t = torch.empty((b, v.c, h, w), dtype=torch.float32, device=self.device)
            nhwc_to_nchw(v, t, stream)
            res[n] = t
        return res
