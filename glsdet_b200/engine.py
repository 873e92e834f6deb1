"""Execution plans of the YOLOX-drone hot path: PAFPN neck -> (FFA) -> decoupled head -> decode.

Three topologies share one builder:
  * variant "p1"    - models/new/yolox10.py (GLSDet P1): patch non-local attention on dark3..dark5 in front of the
                      PAFPN (Non_local_family.py:204-250, evaluated in reassociated form, see _build_nonlocal),
                      three levels, CSP block on dark2, cross-level cls branch;
  * variant "ffa"   - models/ffa/yolox_ffa.py (GLSDet P0): four levels (strides 4..32), FFA fusion on (P3_out,
                      P4_out), CSP block on dark2, level 0 uses tower index 3;
  * variant "p2"    - models/block/non_local/yolo_patch_nonlocal_plus.py (GLSDet P2): Patch_Conv_NonLocal on dark3 and
                      Patch_Conv on dark4 feeding C3_p4 / C3_n3 as third inputs, 7x7 / 5x5 / 3x3 identity convs behind
                      the neck outputs, stock three-level head;
  * variant "stock" - models/base/yolox.py and, with renamed keys, the mmdet pair YOLOXPAFPN + YOLOXHead
                      (yolox-ufp/mmdet/models/necks/yolox_pafpn.py, dense_heads/yolox_head.py): three levels.

A plan is built once per (weights, batch, input size): BatchNorm is folded into the conv weights, weights are
re-laid out for the tcgen05 kernel, every intermediate lives in a preallocated NHWC bf16 buffer, and the forward
pass becomes a flat list of native launches (no PyTorch math on the path):

  * torch.cat is never materialised: producers write straight into channel windows of the concat buffer;
  * upsample + cat in front of a CSP block uses  conv1x1(cat(up(a), b)) = up(conv1x1_a(a)) + conv1x1_b(b):
    the low-resolution half is computed at a quarter of the FLOPs and added (fp32) in the epilogue;
  * CSP conv1/conv2 share their input -> one GEMM with stacked output channels; the same for the first conv of
    the cls/reg towers; reg_preds/obj_preds are one N=5 GEMM;
  * PixelShuffle is a pure re-addressing because the producing conv's output channels are permuted at pack time;
  * the 1x1 prediction convs are fused into the epilogue of the second tower conv (its activated tile never leaves
    the SM); in `decoded` mode that epilogue also applies sigmoid/exp/grid/stride and writes [B, A, 5+nc] rows
    (yolox-drone: normalised centre boxes, utils_bbox.py:254-306; mmdet: pixel centre boxes, yolox_head.py:298-301).

Op lists:  neck_ops -> stem_ops (everything that produces the per-level head inputs p_k) -> tower_ops (first tower
conv) -> pred_raw_ops | pred_dec_ops (second tower conv + fused prediction conv).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch

from . import _native as N
from .ops import (BGemmF32Op, ConvOp, ConvOpF32, DepthwiseOp, GatherBiasOp, NhwcTransposeOp, PatchTransposeOp, RectCopyOp, ScaleShuffleOp, SeGateOp, Upsample2xOp, View, fold_bn,
                  nchw_to_nhwc, nhwc_to_nchw)

BN_EPS = 1e-3


class FFAPathPlan:
    """Neck + head for fixed weights / batch / input size.  State-dict keys follow yolox-drone ("backbone.*" for the
    neck, "head.*" for the head) after `neck_prefix` / `head_prefix` have been stripped."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], batch: int, input_hw: Sequence[int], num_classes: int,
                 device=None, act: str = "silu", neck_prefix: str = "backbone.", head_prefix: str = "head.",
                 parts: Sequence[str] = ("neck", "head"), variant: str = "ffa", decode: str = "drone",
                 precision: str = "bf16", storage: Optional[str] = None):
        """`parts` selects which op lists are built - "neck", "stems" (everything producing the per-level head
        inputs), "towers" (tower + prediction convs), "head" = stems + towers - because a stand-alone neck or head
        module only owns its own weights; `decode` picks the decoded-row flavour: "drone" (normalised) or "mmdet"
        (input pixels); `precision` "bf16" is the tensor-core path, "fp32" the accuracy mode of BASELINE.json configs[0]
        (every tensor fp32, SIMT fp32 convs, csrc/fp32_path.cu; parity bar 1e-3 relative)."""
        assert precision in ("bf16", "fp32")
        self.fp32 = precision == "fp32"
        # 16-bit storage policy of the tensor-core path: "mixed" (default) = bf16 at stride 4, fp16 at the coarser
        # levels (glsdet_b200/_native.py::storage_dtype); "bf16" / "f16" force one type
        self.storage = storage
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.device = dev
        parts = set(parts)
        if "head" in parts:
            parts |= {"stems", "towers"}
        self.parts = parts
        self.variant = variant
        self.decode = decode
        assert variant in ("ffa", "stock", "p1", "p2") and decode in ("drone", "mmdet")
        sd = {}
        for k, v in state_dict.items():
            if k.endswith("num_batches_tracked"):
                continue
            if "neck" in self.parts and k.startswith(neck_prefix) and not k.startswith(neck_prefix + "backbone."):
                sd["backbone." + k[len(neck_prefix):]] = v.detach().cpu()     # CSPDarknet entries are skipped
            elif ({"stems", "towers"} & self.parts) and k.startswith(head_prefix):
                sd["head." + k[len(head_prefix):]] = v.detach().cpu()
        # the weights stay on the HOST while the plan is built (BatchNorm folding in fp64, concatenations, permutations,
        # packing: CPU tensor ops); every operator uploads its packed weights once (ops.ConvOp)
        self.sd = sd
        self.B = batch
        self.in_h, self.in_w = int(input_hw[0]), int(input_hw[1])
        if self.in_h % 32 or self.in_w % 32:
            raise ValueError("input size must be a multiple of 32 (yolox-drone/yolo.py:34-36)")
        self.nc = num_classes
        self.act = N.ACT_BY_NAME[act]
        if "neck" in self.parts:
            self.c0 = sd["backbone.reduce_conv1.conv.weight"].shape[0]   # int(256 * width)
            self.c1 = sd["backbone.lateral_conv0.conv.weight"].shape[0]  # int(512 * width)
            self.c2 = sd["backbone.lateral_conv0.conv.weight"].shape[1]  # int(1024 * width)
        elif "stems" in self.parts:
            self.c0, self.c1, self.c2 = (sd[f"head.stems.{i}.conv.weight"].shape[1] for i in range(3))
        else:   # towers only: the neck-side buffers are never touched, any consistent sizes do
            self.c0 = sd["head.cls_preds.0.weight"].shape[1]
            self.c1, self.c2 = 2 * self.c0, 4 * self.c0
        if "stems" in self.parts:
            self.hc = sd["head.stems.0.conv.weight"].shape[0]            # head width int(256 * width)
        elif "towers" in self.parts:
            self.hc = sd["head.cls_preds.0.weight"].shape[1]             # (the tower convs may be DWConvs: no .conv.weight)
        else:
            self.hc = self.c0
        if variant == "ffa" and "stems" in self.parts:
            self.cd2 = sd["head.csp.conv1.conv.weight"].shape[1]
        elif variant == "p1" and "stems" in self.parts:
            self.cd2 = sd["head.csp_feat0.conv1.conv.weight"].shape[1]
        else:
            self.cd2 = self.c0 // 2
        for c in (self.c0, self.c1, self.c2, self.cd2, self.hc):
            if c % 16:
                raise NotImplementedError(f"channel count {c} is not a multiple of 16")
        self.strides = (4, 8, 16, 32) if variant == "ffa" else (8, 16, 32)   # levels of the head outputs
        self.stride_hw = {s: (self.in_h // s, self.in_w // s) for s in (4, 8, 16, 32)}
        self.level_hw = [self.stride_hw[s] for s in self.strides]
        self.num_anchors = sum(h * w for h, w in self.level_hw)
        self._bufs: Dict[str, torch.Tensor] = {}
        self.neck_ops: List = []
        self.stem_ops: List = []
        self.tower_ops: List = []
        self.pred_raw_ops: List = []
        self.pred_dec_ops: List = []
        self.pred_det_ops: List = []   # decoded boxes / objectness + RAW class logits (the score filter applies the sigmoid)
        self.logits: List[torch.Tensor] = []
        self.pred: Optional[torch.Tensor] = None
        self.flops = 0.0
        self.backbone = None      # BackbonePlan writing straight into self.inputs (attach_backbone)
        self._build()

    # ------------------------------------------------------------------ helpers
    def _dt(self, stride: int, role: str = "head"):
        """Storage type of a tensor at `stride`: fp32 in the accuracy mode, else the 16-bit policy."""
        return torch.float32 if self.fp32 else N.storage_dtype(stride, self.storage, role)

    def _buf(self, name: str, stride: int, channels: int, dtype=None) -> torch.Tensor:
        h, w = self.stride_hw[stride]
        role = "input" if name in ("dark2", "dark3", "dark4", "dark5") else "head"
        t = torch.empty((self.B, h, w, channels), dtype=self._dt(stride, role) if dtype is None else dtype, device=self.device)
        self._bufs[name] = t
        return t

    def _folded(self, p: str):
        sd = self.sd
        return fold_bn(sd[p + ".conv.weight"], sd[p + ".bn.weight"], sd[p + ".bn.bias"], sd[p + ".bn.running_mean"],
                       sd[p + ".bn.running_var"], BN_EPS)

    def _conv(self, ops: List, w, b, srcs, out, k, stride=1, act=None, **kw) -> ConvOp:
        if self.fp32:
            return self._conv_f32(ops, w, b, srcs, out, k, stride, self.act if act is None else act, **kw)
        op = ConvOp(srcs, w, b, ksize=k, stride=stride, act=self.act if act is None else act, out=out, **kw)
        ops.append(op)
        if w is not None:   # the batched products of the non-local block are not part of the reference's conv FLOPs
            self.flops += op.flops
        else:
            self.attn_flops = getattr(self, "attn_flops", 0.0) + op.flops
        return op

    def _conv_f32(self, ops, w, b, srcs, out, k, stride, act, pred_weight=None, pred_bias=None, pred_act=N.ACT_NONE,
                  dec=(0.0, 0.0, 0.0), **kw):
        """fp32 accuracy mode: same graph on ConvOpF32; the fused prediction conv becomes a separate 1x1 conv."""
        if pred_weight is None:
            op = ConvOpF32(srcs, w, b, ksize=k, stride=stride, act=act, out=out, dec=dec, **kw)
            ops.append(op)
            self.flops += op.flops
            return op
        bb, hh, ww = srcs[0].bhw
        tmp = torch.empty((bb, hh // stride, ww // stride, w.shape[0]), dtype=torch.float32, device=self.device)
        op1 = ConvOpF32(srcs, w, b, ksize=k, stride=stride, act=act, out=View(tmp))
        op2 = ConvOpF32([View(tmp)], pred_weight.reshape(pred_weight.shape[0], -1, 1, 1), pred_bias, ksize=1, act=pred_act,
                        out=out, dec=dec, **kw)
        ops.extend([op1, op2])
        self.flops += op1.flops + op2.flops
        return op2

    def _is_dw(self, p: str) -> bool:
        """`p` names a DWConv (models/base/baseConv.py:22-30; phi = 'nano'): keys p.dconv.* and p.pconv.*."""
        return (p + ".dconv.conv.weight") in self.sd

    def _dw(self, ops, p: str, src: View, stride: int = 1, act=None) -> View:
        """dconv half of DWConv `p` (depthwise k x k + BN + act, csrc/dwconv.cu) into a temporary of the source's storage
        type; returns the view the pconv half reads."""
        w, b = self._folded(p + ".dconv")
        bb, h, w_ = src.bhw
        k = w.shape[-1]
        pad = (k - 1) // 2
        ho, wo = (h + 2 * pad - k) // stride + 1, (w_ + 2 * pad - k) // stride + 1
        t = torch.empty((bb, ho, wo, src.c), dtype=src.t.dtype, device=self.device)
        self._bufs[p + ".dconv"] = t
        op = DepthwiseOp(src, w, b, stride=stride, act=self.act if act is None else act, out=View(t))
        ops.append(op)
        self.flops += op.flops
        return View(t)

    def _base_conv(self, ops, p, srcs, out, stride=1, act=None, **kw):
        if self._is_dw(p):   # DWConv.forward: pconv(dconv(x)), both with BN + act
            assert len(srcs) == 1
            return self._base_conv(ops, p + ".pconv", [self._dw(ops, p, srcs[0], stride, act)], out, 1, act, **kw)
        w, b = self._folded(p)
        return self._conv(ops, w, b, srcs, out, w.shape[-1], stride, act, **kw)

    def _csp(self, ops, p: str, stride: int, src: Optional[View], out, *, up_src: Optional[View] = None,
             x_name: str, scratch=None, **out_kw):
        """CSPLayer (models/ffa/darknet.py:91-112), shortcut=False.  With `up_src` the block input is
        cat([upsample(up_src), src], 1) (yolox_ffa.py:207-211,224-228)."""
        sd = self.sd
        w1, b1 = self._folded(p + ".conv1")
        w2, b2 = self._folded(p + ".conv2")
        w12, b12 = torch.cat([w1, w2], 0), torch.cat([b1, b2], 0)
        hid = w1.shape[0]
        srcs = list(src) if isinstance(src, (list, tuple)) else [src]
        X = scratch[0] if scratch is not None else self._buf(x_name, stride, 2 * hid)
        if up_src is not None:
            ca = up_src.c
            T = self._buf(x_name + "_T", stride * 2, 2 * hid, torch.float32)
            self._conv(ops, w12[:, :ca].contiguous(), None, [up_src], View(T), 1, act=N.ACT_NONE)
            self._conv(ops, w12[:, ca:].contiguous(), b12, srcs, View(X), 1, pre_res=View(T), pre_shift=1)
        else:
            self._conv(ops, w12, b12, srcs, View(X), 1)
        bb = scratch[1] if scratch is not None else self._buf(x_name + "_b", stride, hid)
        j = 0
        while f"{p}.m.{j}.conv1.conv.weight" in sd:
            self._base_conv(ops, f"{p}.m.{j}.conv1", [View(X, 0, hid)], View(bb))
            self._base_conv(ops, f"{p}.m.{j}.conv2", [View(bb)], View(X, 0, hid))
            j += 1
        self._base_conv(ops, p + ".conv3", [View(X)], out, **out_kw)

    # ------------------------------------------------------------------ graph
    def _build_p2(self):
        """GLSDet P2 (yolo_patch_nonlocal_plus.py): patch-conv neck of _build_neck_p2 + the stock three-level head."""
        c0, c1, c2, hc = self.c0, self.c1, self.c2, self.hc
        d3, d4, d5 = self._buf("dark3", 8, c0), self._buf("dark4", 16, c1), self._buf("dark5", 32, c2)
        self.pre_loads = []
        self.inputs = (d3, d4, d5)
        if "neck" in self.parts:
            self.neck_out = self._build_neck_p2(self.neck_ops, d3, d4, d5)
        else:
            self.neck_out = (View(self._buf("P3_out", 8, c0)), View(self._buf("P4_out", 16, c1)),
                             View(self._buf("P5_out", 32, c2)))
        self.p = [self._buf(f"p{k}", s, hc) for k, s in enumerate(self.strides)]
        if "stems" in self.parts:
            for k, src in enumerate(self.neck_out):
                self._base_conv(self.stem_ops, f"head.stems.{k}", [src], View(self.p[k]))
        if "towers" in self.parts:
            self._build_towers()

    def _build(self):
        if self.variant == "p2":
            return self._build_p2()
        c0, c1, c2, hc = self.c0, self.c1, self.c2, self.hc
        ffa = self.variant == "ffa"
        p1 = self.variant == "p1"
        d3 = self._buf("dark3", 8, c0)
        d4 = self._buf("dark4", 16, c1)
        d5 = self._buf("dark5", 32, c2)
        self.pre_loads: List = []   # (input index, op): extra per-input conversions run by load_features
        raw = (d3, d4, d5)
        if p1 and "neck" in self.parts:
            # feat_k + Patch_conv_feat_k(feat_k) replaces dark3..dark5 in front of the PAFPN (yolox10.py:262-266)
            d3, d4, d5 = (self._buf(f"feat{i + 1}", s_, c_) for i, (s_, c_) in enumerate(((8, c0), (16, c1), (32, c2))))
        cat5 = self._buf("cat5", 32, 2 * c1)              # [bu_conv1(P4_out) | P5]
        cat4 = self._buf("cat4", 16, 2 * c0)              # [bu_conv2(P3_out) | P4]
        catf = self._buf("catf", 8, 2 * c0 if ffa else c0)  # [P3_out | FFA top after PixelShuffle]
        p5up = self._buf("c3p4_out", 16, c1)
        p4out = self._buf("P4_out", 16, c1)
        p5out = self._buf("P5_out", 32, c2)
        P5 = View(cat5, c1, c1)
        P4 = View(cat4, c0, c0)
        P3o = View(catf, 0, c0)
        if ffa or p1:
            d2 = self._buf("dark2", 4, self.cd2)
            self.inputs = (d2,) + raw
            self.neck_out = (View(d2), P3o, View(p4out), View(p5out))
        else:
            self.inputs = (d3, d4, d5)
            self.neck_out = (P3o, View(p4out), View(p5out))
        # per-level head inputs (what mmdet calls the neck outputs after out_convs)
        if p1:   # head inputs x_k live in the first hc channels of the cls-branch concat buffers
            self.p = []
        else:
            self.p = [self._buf(f"p{k}", s, hc) for k, s in enumerate(self.strides)]
        if "neck" in self.parts:
            if p1:
                for i, (x, f, s_) in enumerate(zip(raw, (d3, d4, d5), (8, 16, 32))):
                    self._build_nonlocal(self.neck_ops, f"backbone.Patch_conv_feat{i + 1}", i + 1, x, f, s_)
            self._build_neck(self.neck_ops, d3, d4, d5, cat4, cat5, p5up, p4out, p5out, P5, P4, P3o)
        if p1:
            if "stems" in self.parts or "towers" in self.parts:
                self._build_p1_head(self._bufs["dark2"], P3o, p4out, p5out)
            return
        if "stems" in self.parts:
            if ffa:
                self._build_ffa_stems(self.stem_ops, self._bufs["dark2"], catf, p4out, p5out, P3o)
            else:
                for k, src in enumerate((P3o, View(p4out), View(p5out))):
                    self._base_conv(self.stem_ops, f"head.stems.{k}", [src], View(self.p[k]))
        if "towers" in self.parts:
            self._build_towers()

    def _build_neck(self, nk, d3, d4, d5, cat4, cat5, p5up, p4out, p5out, P5, P4, P3o):
        c0, c1 = self.c0, self.c1
        # yolox_ffa.py:203-258 == base/yolox.py YOLOPAFPN.forward == mmdet yolox_pafpn.py:117-150
        self._base_conv(nk, "backbone.lateral_conv0", [View(d5)], P5)
        self._csp(nk, "backbone.C3_p4", 16, View(d4), View(p5up), up_src=P5, x_name="c3p4")
        self._base_conv(nk, "backbone.reduce_conv1", [View(p5up)], P4)
        self._csp(nk, "backbone.C3_p3", 8, View(d3), P3o, up_src=P4, x_name="c3p3")
        self._base_conv(nk, "backbone.bu_conv2", [P3o], View(cat4, 0, c0), stride=2)
        self._csp(nk, "backbone.C3_n3", 16, View(cat4), View(p4out), x_name="c3n3")
        self._base_conv(nk, "backbone.bu_conv1", [View(p4out)], View(cat5, 0, c1), stride=2)
        self._csp(nk, "backbone.C3_n4", 32, View(cat5), View(p5out), x_name="c3n4")

    def _build_ffa_stems(self, hd, d2, catf, p4out, p5out, P3o):
        c0, hc = self.c0, self.hc
        p = self.p
        # ---- FFA: ffa.py:74-85 (all ReLU)
        f = "head.ftt"
        relu = N.ACT_RELU
        s_a = self._buf("ffa_a", 16, 4 * hc)
        s_b = self._buf("ffa_b", 16, 4 * hc)
        self._base_conv(hd, f + ".scale", [View(p4out)], View(s_a), act=relu)
        self._base_conv(hd, f + ".create_content_extractor.0", [View(s_a)], View(s_b), act=relu)
        # PixelShuffle(2): out[c, 2y+i, 2x+j] = in[4c+2i+j, y, x]  ->  store channels in (i, j, c) order
        perm = torch.arange(4 * hc).view(hc, 4).t().reshape(-1)  # new n' -> old 4c + (2i+j)
        w, b = self._folded(f + ".create_content_extractor.1")
        self._conv(hd, w[perm].contiguous(), b[perm].contiguous(), [View(s_b)], View(s_a), 1, act=relu)
        se = SeGateOp(View(s_a), self.sd[f + ".se1.fc.0.weight"][:, perm], self.sd[f + ".se1.fc.2.weight"][perm])
        hd.append(se)
        hd.append(ScaleShuffleOp(View(s_a), se.gate, View(catf, c0, hc)))
        tx = self._buf("ffa_text", 8, 2 * hc)
        zz = self._buf("ffa_out", 8, hc)
        self._base_conv(hd, f + ".create_text_extractor.0", [View(catf)], View(tx), act=relu)
        self._base_conv(hd, f + ".conv3", [View(tx)], View(zz), act=relu, post_res=View(catf, c0, hc), post_shift=0)
        # ---- head inputs: yolox_ffa.py:66-73
        split = int(os.environ.get("GLSDET_HCSP_SPLIT", "1"))
        self.hcsp_split = split if (split > 1 and self.B % split == 0 and not self.fp32) else 1
        if self.hcsp_split > 1:
            # the six launches of the block at stride 4 run per group of B / split images on ONE set of scratch buffers
            # sized for a group, so that the intermediates (64 + 32 MB for four 1024^2 images) stay in the 126 MB L2
            # instead of round-tripping HBM between the launches
            nb = self.B // split
            hid = self.sd["head.csp.conv1.conv.weight"].shape[0]
            X = self._buf("hcsp", 4, 2 * hid)[:nb]
            bb = self._buf("hcsp_b", 4, hid)[:nb]
            for g in range(split):
                sl = slice(g * nb, (g + 1) * nb)
                self._csp(hd, "head.csp", 4, View(d2[sl]), View(p[0][sl]), x_name="hcsp", scratch=(X, bb),
                          post_res=View(zz[sl]), post_shift=1)
        else:
            self._csp(hd, "head.csp", 4, View(d2), View(p[0]), x_name="hcsp", post_res=View(zz), post_shift=1)
        self._base_conv(hd, "head.stems.0", [P3o], View(p[1]))
        self._base_conv(hd, "head.stems.1", [View(p4out)], View(p[2]))
        self._base_conv(hd, "head.stems.2", [View(p5out)], View(p[3]))

    def _build_nonlocal(self, ops, p: str, in_idx: int, x: torch.Tensor, out: torch.Tensor, stride: int):
        """out = x + channel_conv(retile(non_local(patch)))  -  Patch_Conv_NonLocal_new.forward
        (models/new/Non_local_family.py:229-250) + the residual of yolox10.py:262-266, with each Non_local_Block (:32-48,
        dot-product mode, no softmax) reassociated:  y = (theta^T phi / T) g = theta (phi^T g / T), hence
            block(X) = X + X W_eff^T + 1 b_eff^T,   W_eff = (Wo G / T) S (Phi^T Wtheta),   b_eff = (Wo G / T) S Phi^T btheta + bo
        with S = [X | 1]^T [X | 1] the Gram matrix of the patch (pixels x channels, plus a ones column that carries the
        conv biases), G = [Wg | bg], Phi = [Wphi | bphi].  O(T C^2) instead of O(T^2 C), no [T, T] intermediate.
        Everything is the tcgen05 conv kernel with per-image weight matrices:
          S = Xt Xt^T (Xt = per-patch transposed copy of the input, ones row appended), Z^T = A2'^T S, W = A1 Z, then
          a 1x1 conv over the patch with weights W[:, :C], bias column W[:, C] + bo, and the input as residual.
        Patch images are ordered b' = (b*2 + py)*2 + px; position (py, px) = lt (0,0), rt (0,1), lb (1,0), rb (1,1)."""
        sd, dev = self.sd, self.device
        B, H, W_, C = x.shape
        nl = self._buf(p + ".nl", stride, C)
        if self.fp32:
            # accuracy mode: every patch position is cut out as a dense fp32 batch and runs the block un-folded
            # (theta / phi / g convs, phi^T g / T, theta M, conv_out + residual: _nonlocal_f32)
            h2, w2 = H // 2, W_ // 2
            for pos, (y0, x0, ph, pw) in (("lt", (0, 0, h2, w2)), ("rt", (0, w2, h2, W_ - w2)),
                                          ("lb", (h2, 0, H - h2, w2)), ("rb", (h2, w2, H - h2, W_ - w2))):
                xp = torch.empty((B, ph, pw, C), dtype=x.dtype, device=dev)
                ops.append(RectCopyOp(View(x), View(xp), B, [(0, y0, x0, 0, 0, 0, ph, pw)]))
                nlp = self._nonlocal_f32(ops, f"{p}.feat_patchconv_{pos}_nonlocal", View(xp))
                ops.append(RectCopyOp(View(nlp), View(nl), B, [(0, 0, 0, 0, y0, x0, ph, pw)]))
                self._keepalive = getattr(self, "_keepalive", []) + [xp, nlp]
        elif H % 2 or W_ % 2:
            # unequal 2x2 split (int(H/2), int(W/2): Non_local_family.py:230-233, e.g. 17 x 32 at 544 x 1024): every
            # patch position is cut out as its own dense batch and runs its own chain of GEMMs
            h2, w2 = H // 2, W_ // 2
            for pos, (y0, x0, ph, pw) in (("lt", (0, 0, h2, w2)), ("rt", (0, w2, h2, W_ - w2)),
                                          ("lb", (h2, 0, H - h2, w2)), ("rb", (h2, w2, H - h2, W_ - w2))):
                xp = torch.empty((B, ph, pw, C), dtype=x.dtype, device=dev)
                ops.append(RectCopyOp(View(x), View(xp), B, [(0, y0, x0, 0, 0, 0, ph, pw)]))
                xt, Ca = self._nonlocal_operand(f"{p}.{pos}", B, C, ph * pw, x.dtype)
                ops.append(NhwcTransposeOp(View(xp), xt, scale=(ph * pw) ** -0.5))
                Wm, bias = self._nonlocal_gemms(ops, p, xt, C, ph * pw, B, dict(src_shared=1), positions=(pos,))
                nlp = torch.empty_like(xp)
                self._conv(ops, None, None, [View(xp)], View(nlp), 1, act=N.ACT_NONE, weight_raw=Wm.view(B, C, Ca),
                           n_out=C, pre_res=View(bias), pre_shift=30, post_res=View(xp), post_shift=0)
                ops.append(RectCopyOp(View(nlp), View(nl), B, [(0, 0, 0, 0, y0, x0, ph, pw)]))
                self._keepalive = getattr(self, "_keepalive", []) + [xp, nlp]
        else:
            T = (H // 2) * (W_ // 2)
            Bp = 4 * B
            xt, Ca = self._nonlocal_operand(p, Bp, C, T, x.dtype)
            self.pre_loads.append((in_idx, PatchTransposeOp(xt, C, H, W_, scale=T ** -0.5)))
            Wm, bias = self._nonlocal_gemms(ops, p, xt, C, T, Bp, dict(src_shared=4))
            self._conv(ops, None, None, [View(x)], View(nl), 1, act=N.ACT_NONE, weight_raw=Wm.view(Bp, C, Ca), n_out=C,
                       patch_mode=True, pre_res=View(bias), pre_shift=30, post_res=View(x), post_shift=0)
        self._base_conv(ops, p + ".channel_conv", [View(nl)], View(out), post_res=View(x), post_shift=0)

    def _nonlocal_f32(self, ops, q: str, x: View) -> torch.Tensor:
        """Non_local_Block.forward (models/new/Non_local_family.py:32-48; Identity_Conv.py:205-246, dot-product mode) on a
        dense fp32 batch of patches, fp32 accuracy mode:  theta | phi | g = 1x1 convs with bias (one stacked launch),
        M = phi^T g / T per image (the reference's  P = theta^T phi / T, y = P g  is  theta (phi^T g / T): no softmax sits
        between the two products), y = theta M, out = x + conv_out(y)."""
        sd, dev = self.sd, self.device
        wt, wp, wg, wo = (sd[f"{q}.{n}.weight"].float() for n in ("theta", "phi", "g", "conv_out"))
        bt, bp, bg, bo = (sd[f"{q}.{n}.bias"].float() for n in ("theta", "phi", "g", "conv_out"))
        ci = wt.shape[0]
        B, h, w = x.bhw
        tpg = torch.empty((B, h, w, 3 * ci), dtype=torch.float32, device=dev)
        self._conv(ops, torch.cat([wt, wp, wg], 0), torch.cat([bt, bp, bg], 0), [x], View(tpg), 1, act=N.ACT_NONE)
        m = torch.empty((B, ci, ci), dtype=torch.float32, device=dev)
        y = torch.empty((B, h, w, ci), dtype=torch.float32, device=dev)
        for op in (BGemmF32Op("gram", View(tpg, ci, ci), View(tpg, 2 * ci, ci), m, alpha=1.0 / (h * w)),
                   BGemmF32Op("apply", View(tpg, 0, ci), m, View(y))):
            ops.append(op)
            self.attn_flops = getattr(self, "attn_flops", 0.0) + op.flops
        out = torch.empty((B, h, w, wo.shape[0]), dtype=torch.float32, device=dev)
        self._conv(ops, wo, bo, [View(y)], View(out), 1, act=N.ACT_NONE, post_res=x, post_shift=0)
        self._keepalive = getattr(self, "_keepalive", []) + [tpg, m, y]
        return out

    def _nonlocal_operand(self, p: str, Bp: int, C: int, T: int, dtype):
        """Per-patch transposed operand Xt[b'][c][t] / sqrt(T) with the ones row (channel sums / conv biases) preset.
        The 1 / sqrt(T) on both Gram operands makes S = Xt Xt^T the MEAN outer product (the 1 / T of
        Non_local_family.py:29), whose entries are O(1) - the plain sum over T = 4096 pixels would leave the fp16 range."""
        Tp = (T + 63) // 64 * 64
        Ca = (C // 64 + 1) * 64      # C channels + the ones row, padded to whole 64-wide K chunks (C = 32 at nano width)
        xt = torch.zeros((Bp, Ca, Tp), dtype=dtype, device=self.device)
        xt[:, C, :T] = T ** -0.5
        self._bufs[p + ".xt"] = xt
        return xt, Ca

    def _nonlocal_gemms(self, ops, p: str, xt: torch.Tensor, C: int, T: int, Bp: int, shared_kw: dict,
                        positions=("lt", "rt", "lb", "rb")):
        """Gram matrix and the two C x C products of the reassociated non-local block for Bp patch images; returns
        (W [Bp, 1, C, Ca] 16-bit: columns 0..C-1 = W_eff, column C = b_eff - b_o;  bias [Bp, 1, 1, C] fp32 = b_eff).
        `shared_kw` tells the conv operator how a patch image selects its position's static matrices
        (position order lt, rt, lb, rb = py * 2 + px)."""
        sd, dev = self.sd, self.device
        Ca, Tp = xt.shape[1], xt.shape[2]
        npos = len(positions)
        a1 = torch.zeros((npos, C, Ca), dtype=torch.float64)
        a2t = torch.zeros((npos, Ca, Ca), dtype=torch.float64)
        bo = torch.zeros((npos, C), dtype=torch.float32)
        for i, pos in enumerate(positions):
            q = f"{p}.feat_patchconv_{pos}_nonlocal."
            wg, wt, wp, wo = (sd[q + n + ".weight"].double().flatten(1) for n in ("g", "theta", "phi", "conv_out"))
            bg, bt, bp = (sd[q + n + ".bias"].double() for n in ("g", "theta", "phi"))
            G = torch.cat([wg, bg[:, None]], 1)          # [Ci, C+1]
            Phi = torch.cat([wp, bp[:, None]], 1)        # [Ci, C+1]
            a1[i, :, :C + 1] = wo @ G                    # [C, C+1]; the 1 / T sits in S (see _nonlocal_operand)
            a2 = Phi.t() @ torch.cat([wt, bt[:, None]], 1)   # [C+1, C+1]: columns 0..C-1 -> W_eff, column C -> b_eff - bo
            a2t[i, :C + 1, :C + 1] = a2.t()
            bo[i] = sd[q + "conv_out.bias"].float()
        dt = xt.dtype
        a1 = a1.to(dt).view(npos, 1, C, Ca).contiguous().to(dev)
        a2t = a2t.to(dt).view(npos, 1, Ca, Ca).contiguous().to(dev)
        div = shared_kw.get("src_shared_div", 0)
        img_pos = (torch.arange(Bp) // div) if div else (torch.arange(Bp) % npos)
        bo_full = bo[img_pos].contiguous().to(dev)       # [Bp, C]: b_o of every patch image
        S = torch.empty((Bp, 1, Ca, Ca), dtype=dt, device=dev)
        Zt = torch.empty((Bp, 1, Ca, Ca), dtype=dt, device=dev)
        Wm = torch.empty((Bp, 1, C, Ca), dtype=dt, device=dev)
        bias = torch.empty((Bp, 1, 1, C), dtype=torch.float32, device=dev)
        none = N.ACT_NONE
        # S[c][c'] = sum_t Xt[c][t] Xt[c'][t]
        self._conv(ops, None, None, [View(xt.view(Bp, 1, Ca, Tp))], View(S), 1, act=none, weight_raw=xt, n_out=Ca)
        # Zt[k][i] = sum_j A2'[j][k] S[i][j]
        self._conv(ops, None, None, [View(a2t)], View(Zt), 1, act=none, weight_raw=S.view(Bp, Ca, Ca), n_out=Ca,
                   batch=Bp, **shared_kw)
        # W[n][k] = sum_i A1[n][i] Z[i][k]
        self._conv(ops, None, None, [View(a1)], View(Wm), 1, act=none, weight_raw=Zt.view(Bp, Ca, Ca), n_out=Ca,
                   batch=Bp, **shared_kw)
        ops.append(GatherBiasOp(Wm.view(Bp, C, Ca), bo_full, bias, C))
        self._keepalive = getattr(self, "_keepalive", []) + [a1, a2t, bo_full, S, Zt, Wm, bias]
        return Wm, bias

    def _build_patch_conv(self, ops, p: str, x: torch.Tensor, stride: int, nonlocal_: bool, out: View):
        """Patch_Conv / Patch_Conv_NonLocal (models/block/non_local/Identity_Conv.py:292-318, 353-384): 2x2 split, one
        3x3 BaseConv per patch position (own weights, zero padding at the patch borders), [the dot-product non-local
        block per patch, reassociated as in _build_nonlocal], seam convs on the left / right / top / bottom halves,
        re-tile, 1x1 channel conv with bias.  Patches live as dense images ordered by position
        (b' = position * B + b, position = py * 2 + px), so every conv is a plain dense launch; the split, the halves
        and the re-tiling are rectangle copies."""
        sd, dev = self.sd, self.device
        B, H, W_, Cin = x.shape
        if H % (2 * stride) or W_ % (2 * stride):
            raise NotImplementedError("patch convs need equal 2x2 patches (unequal splits are not supported)")
        mid = sd[f"{p}.feat_patchconv_lt.conv.weight"].shape[0]
        h2, w2 = H // 2, W_ // 2
        hq, wq = h2 // stride, w2 // stride
        pos_yx = ((0, 0), (0, 1), (1, 0), (1, 1))    # lt, rt, lb, rb
        names = ("lt", "rt", "lb", "rb")
        dt = x.dtype
        xp = torch.empty((4 * B, h2, w2, Cin), dtype=dt, device=dev)
        ops.append(RectCopyOp(View(x), View(xp), B, [(0, py * h2, px * w2, i * B, 0, 0, h2, w2)
                                                     for i, (py, px) in enumerate(pos_yx)]))
        y = torch.empty((4 * B, hq, wq, mid), dtype=dt, device=dev)
        for i, nm in enumerate(names):
            self._base_conv(ops, f"{p}.feat_patchconv_{nm}", [View(xp[i * B:(i + 1) * B])], View(y[i * B:(i + 1) * B]),
                            stride=stride)
        if nonlocal_ and self.fp32:
            nl = torch.empty_like(y)
            for i, nm in enumerate(names):
                nlp = self._nonlocal_f32(ops, f"{p}.feat_patchconv_{nm}_nonlocal", View(y[i * B:(i + 1) * B]))
                ops.append(RectCopyOp(View(nlp), View(nl[i * B:(i + 1) * B]), B, [(0, 0, 0, 0, 0, 0, hq, wq)]))
                self._keepalive = getattr(self, "_keepalive", []) + [nlp]
        elif nonlocal_:
            T = hq * wq
            xt, Ca = self._nonlocal_operand(p, 4 * B, mid, T, dt)
            ops.append(NhwcTransposeOp(View(y), xt, scale=T ** -0.5))
            Wm, bias = self._nonlocal_gemms(ops, p, xt, mid, T, 4 * B, dict(src_shared=4, src_shared_div=B))
            nl = torch.empty_like(y)
            self._conv(ops, None, None, [View(y)], View(nl), 1, act=N.ACT_NONE, weight_raw=Wm.view(4 * B, mid, Ca),
                       n_out=mid, pre_res=View(bias), pre_shift=30, post_res=View(y), post_shift=0)
        else:
            nl = y
        lr_in = torch.empty((2 * B, 2 * hq, wq, mid), dtype=dt, device=dev)   # L images, then R images
        tb_in = torch.empty((2 * B, hq, 2 * wq, mid), dtype=dt, device=dev)   # T images, then B images
        ops.append(RectCopyOp(View(nl), View(lr_in), B, [(i * B, 0, 0, px * B, py * hq, 0, hq, wq)
                                                         for i, (py, px) in enumerate(pos_yx)]))
        ops.append(RectCopyOp(View(nl), View(tb_in), B, [(i * B, 0, 0, py * B, 0, px * wq, hq, wq)
                                                         for i, (py, px) in enumerate(pos_yx)]))
        lr_o, tb_o = torch.empty_like(lr_in), torch.empty_like(tb_in)
        self._base_conv(ops, f"{p}.feat_patchconv_l", [View(lr_in[:B])], View(lr_o[:B]))
        self._base_conv(ops, f"{p}.feat_patchconv_r", [View(lr_in[B:])], View(lr_o[B:]))
        self._base_conv(ops, f"{p}.feat_patchconv_t", [View(tb_in[:B])], View(tb_o[:B]))
        self._base_conv(ops, f"{p}.feat_patchconv_b", [View(tb_in[B:])], View(tb_o[B:]))
        cc = torch.empty((B, 2 * hq, 2 * wq, 2 * mid), dtype=dt, device=dev)   # [cat(l, r; W) | cat(t, b; H)]
        ops.append(RectCopyOp(View(lr_o), View(cc, 0, mid), B, [(0, 0, 0, 0, 0, 0, 2 * hq, wq), (B, 0, 0, 0, 0, wq, 2 * hq, wq)]))
        ops.append(RectCopyOp(View(tb_o), View(cc, mid, mid), B, [(0, 0, 0, 0, 0, 0, hq, 2 * wq), (B, 0, 0, 0, hq, 0, hq, 2 * wq)]))
        self._conv(ops, sd[p + ".channel_conv.weight"].float(), sd[p + ".channel_conv.bias"].float(), [View(cc)], out, 1,
                   act=N.ACT_NONE)
        self._keepalive = getattr(self, "_keepalive", []) + [xp, y, nl, lr_in, tb_in, lr_o, tb_o, cc]

    def _build_neck_p2(self, nk, d3, d4, d5):
        """models/block/non_local/yolo_patch_nonlocal_plus.py:180-247: PAFPN with two extra inputs (the patch modules on
        dark3 / dark4, concatenated into C3_p4 / C3_n3) and k x k identity convs (7 / 5 / 3) behind the three outputs."""
        c0, c1, c2 = self.c0, self.c1, self.c2
        sd = self.sd
        f1p = self._buf("feat1_patch", 16, c1)
        cat4 = self._buf("cat4", 16, 3 * c0)               # [bu_conv2(P3_out) | P4 | feat2_patch]
        cat5 = self._buf("cat5", 32, 2 * c1)               # [bu_conv1(P4_out) | P5]
        P5, P4 = View(cat5, c1, c1), View(cat4, c0, c0)
        self._build_patch_conv(nk, "backbone.Patch_conv_feat1", d3, 2, True, View(f1p))
        self._build_patch_conv(nk, "backbone.Patch_conv_feat2", d4, 1, False, View(cat4, 2 * c0, c0))
        self._base_conv(nk, "backbone.lateral_conv0", [View(d5)], P5)
        p5up = self._buf("c3p4_out", 16, c1)
        self._csp(nk, "backbone.C3_p4", 16, [View(d4), View(f1p)], View(p5up), up_src=P5, x_name="c3p4")
        self._base_conv(nk, "backbone.reduce_conv1", [View(p5up)], P4)
        p3raw, p4raw, p5raw = self._buf("p3raw", 8, c0), self._buf("p4raw", 16, c1), self._buf("p5raw", 32, c2)
        p3o, p4out, p5out = self._buf("P3_out", 8, c0), self._buf("P4_out", 16, c1), self._buf("P5_out", 32, c2)
        self._csp(nk, "backbone.C3_p3", 8, View(d3), View(p3raw), up_src=P4, x_name="c3p3")
        ident = lambda q, src, dst: self._conv(nk, sd[q + ".conv.weight"].float(), sd[q + ".conv.bias"].float(), [View(src)],
                                               View(dst), sd[q + ".conv.weight"].shape[-1], act=N.ACT_NONE)
        ident("backbone.P3_Identity", p3raw, p3o)
        self._base_conv(nk, "backbone.bu_conv2", [View(p3o)], View(cat4, 0, c0), stride=2)
        self._csp(nk, "backbone.C3_n3", 16, View(cat4), View(p4raw), x_name="c3n3")
        ident("backbone.P4_Identity", p4raw, p4out)
        self._base_conv(nk, "backbone.bu_conv1", [View(p4out)], View(cat5, 0, c1), stride=2)
        self._csp(nk, "backbone.C3_n4", 32, View(cat5), View(p5raw), x_name="c3n4")
        ident("backbone.P5_Identity", p5raw, p5out)
        return View(p3o), View(p4out), View(p5out)

    def _build_p1_head(self, d2, P3o, p4out, p5out):
        """models/new/yolox10.py:70-158.  Level k (strides 8/16/32): x_k = stems[k](P_k); cls branch input
        cat([x_k, up_convs[k](level below), upsample(x_{k+1})]) (the top level has no upsampled part); reg branch
        input x_k.  The concat buffers are written in place by the producers; the prediction convs are fused into the
        second tower convs as in the other variants."""
        hc, nc, sd = self.hc, self.nc, self.sd
        nch = 5 + nc
        st, tw = self.stem_ops, self.tower_ops
        f0 = self._buf("p1_f0", 4, hc)
        cat = [self._buf(f"p1_cat{k}", s_, (3 if k < 2 else 2) * hc) for k, s_ in enumerate((8, 16, 32))]
        xs = [View(c, 0, hc) for c in cat]
        self.p = cat
        if "stems" in self.parts:
            self._csp(st, "head.csp_feat0", 4, View(d2), View(f0), x_name="hcsp")
            for k, src in enumerate((P3o, View(p4out), View(p5out))):
                self._base_conv(st, f"head.stems.{k}", [src], xs[k])
            for k, s_ in enumerate((8, 16, 32)):
                lower = View(f0) if k == 0 else xs[k - 1]
                tmp = self._buf(f"p1_up{k}", s_ // 2, hc)
                self._base_conv(st, f"head.up_convs.{k}.0", [lower], View(tmp))
                self._base_conv(st, f"head.up_convs.{k}.1", [View(tmp)], View(cat[k], hc, hc), stride=2)
                if k < 2:
                    st.append(Upsample2xOp(xs[k + 1], View(cat[k], 2 * hc, hc)))
        if "towers" not in self.parts:
            return
        self.logits = [torch.empty((self.B, nch, h, w), dtype=torch.float32, device=self.device) for h, w in self.level_hw]
        self._alloc_pred(nch)
        box_act = N.ACT_YOLOX_BOX if self.decode == "drone" else N.ACT_MMDET_BOX
        a_off = 0
        for k, (h, w) in enumerate(self.level_hw):
            s_ = self.strides[k]
            m = 3 if k < 2 else 2
            tc = self._buf(f"p1_tc{k}", s_, m * hc)
            tr = self._buf(f"p1_tr{k}", s_, hc)
            self._base_conv(tw, f"head.cls_convs.{k}.0", [View(cat[k])], View(tc))
            self._base_conv(tw, f"head.reg_convs.{k}.0", [xs[k]], View(tr))
            cls_in, reg_in, k2 = View(tc), View(tr), 3
            if self._is_dw(f"head.cls_convs.{k}.1"):   # phi = 'nano': depthwise halves here, predictions fused into the 1x1 halves
                cls_in = self._dw(tw, f"head.cls_convs.{k}.1", cls_in)
                reg_in = self._dw(tw, f"head.reg_convs.{k}.1", reg_in)
                wc1, bc1 = self._folded(f"head.cls_convs.{k}.1.pconv")
                wr1, br1 = self._folded(f"head.reg_convs.{k}.1.pconv")
                k2 = 1
            else:
                wc1, bc1 = self._folded(f"head.cls_convs.{k}.1")
                wr1, br1 = self._folded(f"head.reg_convs.{k}.1")
            w_ro = torch.cat([sd[f"head.reg_preds.{k}.weight"], sd[f"head.obj_preds.{k}.weight"]], 0).float()
            b_ro = torch.cat([sd[f"head.reg_preds.{k}.bias"], sd[f"head.obj_preds.{k}.bias"]], 0).float()
            w_cl, b_cl = sd[f"head.cls_preds.{k}.weight"].float(), sd[f"head.cls_preds.{k}.bias"].float()
            kw = dict(out_mode=N.OUT_NCHW_F32, out_ld=nch, out_batch_stride=nch * h * w)
            self._conv(self.pred_raw_ops, wr1, br1, [reg_in], self.logits[k], k2, out_coff=0, pred_weight=w_ro,
                       pred_bias=b_ro, pred_act=N.ACT_NONE, **kw)
            self._conv(self.pred_raw_ops, wc1, bc1, [cls_in], self.logits[k], k2, out_coff=5, pred_weight=w_cl,
                       pred_bias=b_cl, pred_act=N.ACT_NONE, **kw)
            stride = float(self.in_h / h)
            kw, c_reg, c_cls = self._pred_out(nch, a_off)
            op = self._conv(self.pred_dec_ops, wr1, br1, [reg_in], self.pred_store, k2, out_coff=c_reg, pred_weight=w_ro,
                            pred_bias=b_ro, pred_act=box_act, dec=(stride, float(self.in_w), float(self.in_h)), **kw)
            self._conv(self.pred_dec_ops, wc1, bc1, [cls_in], self.pred_store, k2, out_coff=c_cls, pred_weight=w_cl,
                       pred_bias=b_cl, pred_act=N.ACT_SIGMOID, **kw)
            if not self.fp32:
                self.pred_det_ops.append(op)
                self._conv(self.pred_det_ops, wc1, bc1, [cls_in], self.pred_store, k2, out_coff=c_cls, pred_weight=w_cl,
                           pred_bias=b_cl, pred_act=N.ACT_NONE, **kw)
            a_off += h * w
        self.flops -= sum(op.flops for op in self.pred_dec_ops) + sum(op.flops for op in self.pred_det_ops[1::2])
        if self.fp32:
            self.pred_det_ops = self.pred_dec_ops
        self.flops += sum(2.0 * self.B * h * w * hc * (5 + nc) for h, w in self.level_hw)

    def _alloc_pred(self, nch: int):
        """Decoded predictions.  bf16 plans store them as planes [B, 5+nc, A] - the physical layout of the tensor the
        reference's decode_outputs returns (utils_bbox.py:266,306: cat along dim 2, then permute(0, 2, 1) as a view) - so
        that the fused prediction convs store coalesced runs per channel and the score filter reads coalesced;
        `self.pred` is the [B, A, 5+nc] view with strides ((5+nc) A, 1, A), exactly like the reference's.  The fp32
        accuracy mode keeps contiguous rows."""
        if self.fp32:
            self.pred_store = torch.empty((self.B, self.num_anchors, nch), dtype=torch.float32, device=self.device)
            self.pred = self.pred_store
        else:
            self.pred_store = torch.empty((self.B, nch, self.num_anchors), dtype=torch.float32, device=self.device)
            self.pred = self.pred_store.permute(0, 2, 1)

    def _pred_out(self, nch: int, a_off: int):
        """Output addressing of a level's fused prediction convs: (kwargs, out_coff of reg|obj, out_coff of cls)."""
        if self.fp32:
            return (dict(out_mode=N.OUT_NHWC_F32, out_ld=nch, out_batch_stride=self.num_anchors * nch),
                    a_off * nch, a_off * nch + 5)
        return (dict(out_mode=N.OUT_NCHW_F32, out_ld=nch, out_batch_stride=self.num_anchors * nch,
                     out_plane_stride=self.num_anchors, out_elem_offset=a_off), 0, 5)

    def _build_towers(self):
        """Towers + predictions: yolox_ffa.py:76-117 (level 0 uses tower index 3) / base/yolox.py / mmdet
        yolox_head.py:184-195."""
        hc, nc, sd = self.hc, self.nc, self.sd
        nch = 5 + nc
        self.logits = [torch.empty((self.B, nch, h, w), dtype=torch.float32, device=self.device)
                       for h, w in self.level_hw]
        self._alloc_pred(nch)
        box_act = N.ACT_YOLOX_BOX if self.decode == "drone" else N.ACT_MMDET_BOX
        a_off = 0
        for k, (h, w) in enumerate(self.level_hw):
            i = k if self.variant in ("stock", "p2") else (3 if k == 0 else k - 1)
            F = self._buf(f"tower{k}", self.strides[k], 2 * hc)
            cls_in, reg_in = [View(F, 0, hc)], [View(F, hc, hc)]
            k2 = 3   # kernel size of the conv that carries the fused prediction conv
            if self._is_dw(f"head.cls_convs.{i}.0"):
                # phi = 'nano': every tower conv is a DWConv.  The first pair cannot share one GEMM any more (each tower
                # has its own depthwise filter on p_k); the depthwise halves of the second pair run in tower_ops and the
                # prediction convs are fused into their 1x1 pconv halves.
                self._base_conv(self.tower_ops, f"head.cls_convs.{i}.0", [View(self.p[k])], cls_in[0])
                self._base_conv(self.tower_ops, f"head.reg_convs.{i}.0", [View(self.p[k])], reg_in[0])
                cls_in = [self._dw(self.tower_ops, f"head.cls_convs.{i}.1", cls_in[0])]
                reg_in = [self._dw(self.tower_ops, f"head.reg_convs.{i}.1", reg_in[0])]
                wc1, bc1 = self._folded(f"head.cls_convs.{i}.1.pconv")
                wr1, br1 = self._folded(f"head.reg_convs.{i}.1.pconv")
                k2 = 1
            else:
                wc0, bc0 = self._folded(f"head.cls_convs.{i}.0")
                wr0, br0 = self._folded(f"head.reg_convs.{i}.0")
                self._conv(self.tower_ops, torch.cat([wc0, wr0], 0), torch.cat([bc0, br0], 0), [View(self.p[k])], View(F), 3)
                wc1, bc1 = self._folded(f"head.cls_convs.{i}.1")
                wr1, br1 = self._folded(f"head.reg_convs.{i}.1")
            w_ro = torch.cat([sd[f"head.reg_preds.{i}.weight"], sd[f"head.obj_preds.{i}.weight"]], 0).float()
            b_ro = torch.cat([sd[f"head.reg_preds.{i}.bias"], sd[f"head.obj_preds.{i}.bias"]], 0).float()
            w_cl, b_cl = sd[f"head.cls_preds.{i}.weight"].float(), sd[f"head.cls_preds.{i}.bias"].float()
            # The second tower conv and the 1x1 prediction conv run as one kernel (the activated tile stays on the
            # SM).  Raw variant: reference layout cat([reg, obj, cls], 1) (yolox_ffa.py:116).
            kw = dict(out_mode=N.OUT_NCHW_F32, out_ld=nch, out_batch_stride=nch * h * w)
            self._conv(self.pred_raw_ops, wr1, br1, reg_in, self.logits[k], k2, out_coff=0, pred_weight=w_ro,
                       pred_bias=b_ro, pred_act=N.ACT_NONE, **kw)
            self._conv(self.pred_raw_ops, wc1, bc1, cls_in, self.logits[k], k2, out_coff=5, pred_weight=w_cl,
                       pred_bias=b_cl, pred_act=N.ACT_NONE, **kw)
            # Decoded variant: the decode runs in the same epilogue, rows of [B, A, 5+nc].  Stride as the reference
            # computes it: input_shape[0] / h (utils_bbox.py:285); equal to the integer mmdet stride.
            stride = float(self.in_h / h)
            kw, c_reg, c_cls = self._pred_out(nch, a_off)
            op = self._conv(self.pred_dec_ops, wr1, br1, reg_in, self.pred_store, k2, out_coff=c_reg, pred_weight=w_ro,
                            pred_bias=b_ro, pred_act=box_act, dec=(stride, float(self.in_w), float(self.in_h)), **kw)
            self._conv(self.pred_dec_ops, wc1, bc1, cls_in, self.pred_store, k2, out_coff=c_cls, pred_weight=w_cl,
                       pred_bias=b_cl, pred_act=N.ACT_SIGMOID, **kw)
            if not self.fp32:
                self.pred_det_ops.append(op)
                self._conv(self.pred_det_ops, wc1, bc1, cls_in, self.pred_store, k2, out_coff=c_cls, pred_weight=w_cl,
                           pred_bias=b_cl, pred_act=N.ACT_NONE, **kw)
            a_off += h * w
        # the second tower convs exist twice (raw + decoded variants); count them once, plus the prediction convs
        self.flops -= sum(op.flops for op in self.pred_dec_ops) + sum(op.flops for op in self.pred_det_ops[1::2])
        if self.fp32:
            self.pred_det_ops = self.pred_dec_ops
        self.flops += sum(2.0 * self.B * h * w * hc * (5 + nc) for h, w in self.level_hw)

    def attach_backbone(self, state_dict: Dict[str, torch.Tensor], prefix: str, act: str = "silu"):
        """Chain the CSPDarknet plan (glsdet_b200/backbone.py) in front of the neck: its dark2..dark5 outputs ARE this
        plan's NHWC bf16 input buffers, so an image batch goes backbone -> neck -> head without leaving the native
        layout.  Not available for plans with per-input pre-loads (P1's Gram operands are
        built from the NCHW fp32 features) or in the fp32 accuracy mode."""
        from .backbone import FEATURES, BackbonePlan

        if self.fp32 or self.pre_loads:
            raise NotImplementedError("attach_backbone: bf16 plans without pre-loads only")
        names = FEATURES[-len(self.inputs):]
        self.backbone = BackbonePlan(state_dict, self.B, (self.in_h, self.in_w), device=self.device, act=act, prefix=prefix,
                                     outs=dict(zip(names, self.inputs)))
        return self.backbone

    def forward_image(self, image: torch.Tensor, decoded, stream=None):
        """Image batch [B, 3, H, W] fp32 -> raw logits (decoded=False) or decoded rows [B, A, 5+nc] (decoded=True / "det")."""
        self.backbone.run(image, stream)
        self.run_neck(stream)
        self.run_head(decoded, stream)
        return self.pred if decoded else self.logits

    def forward_image_uint8(self, image_u8: torch.Tensor, decoded, stream=None, **norm):
        """uint8 HWC batch [B, H, W, 3] -> like forward_image (normalisation fused into the Focus kernel)."""
        self.backbone.run_uint8(image_u8, stream=stream, **norm)
        self.run_neck(stream)
        self.run_head(decoded, stream)
        return self.pred if decoded else self.logits

    # ------------------------------------------------------------------ execution
    def load_features(self, feats: Sequence[torch.Tensor], stream=None) -> None:
        """Backbone features, NCHW fp32 -> internal NHWC bf16 buffers: (dark2, dark3, dark4, dark5) for the FFA
        variant, (dark3, dark4, dark5) for the stock one."""
        assert len(feats) == len(self.inputs), (len(feats), len(self.inputs))
        for src, dst in zip(feats, self.inputs):
            nchw_to_nhwc(src.contiguous(), View(dst), stream)
        for idx, op in self.pre_loads:
            op.launch(feats[idx].contiguous(), stream)

    def load_head_inputs(self, inputs: Sequence[torch.Tensor], stream=None) -> None:
        """The tuple YOLOPAFPN.forward returns (NCHW fp32) -> internal buffers."""
        assert len(inputs) == len(self.neck_out)
        for src, dst in zip(inputs, self.neck_out):
            nchw_to_nhwc(src.contiguous(), dst, stream)

    def load_tower_inputs(self, feats: Sequence[torch.Tensor], stream=None) -> None:
        """Per-level head inputs p_k (what mmdet's neck returns after out_convs), NCHW fp32 -> internal buffers."""
        assert len(feats) == len(self.p)
        for src, dst in zip(feats, self.p):
            nchw_to_nhwc(src.contiguous(), View(dst), stream)

    @staticmethod
    def _run(ops, stream):
        for op in ops:
            op.launch(stream)

    def run_neck(self, stream=None) -> None:
        self._run(self.neck_ops, stream)

    def run_stems(self, stream=None) -> None:
        self._run(self.stem_ops, stream)

    def run_towers(self, decoded, stream=None) -> None:
        """`decoded`: False = raw logits per level (YoloBody.forward), True = decoded predictions in self.pred
        (decode_outputs), "det" = decoded boxes / objectness with RAW class logits in self.pred for the fused detect path
        (DeviceNMS.launch(..., cls_logits=plan.det_cls_logits) applies the class sigmoid in the score filter)."""
        self._run(self.tower_ops, stream)
        self._run(self.pred_det_ops if decoded == "det" else self.pred_dec_ops if decoded else self.pred_raw_ops, stream)

    @property
    def det_cls_logits(self) -> bool:
        return not self.fp32

    def run_head(self, decoded, stream=None) -> None:
        self.run_stems(stream)
        self.run_towers(decoded, stream)

    def _to_nchw(self, views, stream=None) -> List[torch.Tensor]:
        outs = []
        for v in views:
            b, h, w = v.bhw
            t = torch.empty((b, v.c, h, w), dtype=torch.float32, device=self.device)
            nhwc_to_nchw(v, t, stream)
            outs.append(t)
        return outs

    def neck_outputs_nchw(self, stream=None) -> List[torch.Tensor]:
        return self._to_nchw(self.neck_out, stream)

    def stem_outputs_nchw(self, stream=None) -> List[torch.Tensor]:
        return self._to_nchw([View(t) for t in self.p], stream)

    def forward_logits(self, feats: Sequence[torch.Tensor], stream=None) -> List[torch.Tensor]:
        """Raw [B, 5+nc, h, w] logits per level, as YoloBody.forward returns them (yolox_ffa.py:116)."""
        self.load_features(feats, stream)
        self.run_neck(stream)
        self.run_head(False, stream)
        return self.logits

    def forward_decoded(self, feats: Sequence[torch.Tensor], stream=None) -> torch.Tensor:
        """decode_outputs(YoloBody.forward(x)) as one fused pass: [B, A, 5+nc]."""
        self.load_features(feats, stream)
        self.run_neck(stream)
        self.run_head(True, stream)
        return self.pred

    def forward_detect(self, feats: Sequence[torch.Tensor], stream=None) -> torch.Tensor:
        """forward_decoded for the fused detect path: class columns stay raw logits (see run_towers)."""
        self.load_features(feats, stream)
        self.run_neck(stream)
        self.run_head("det", stream)
        return self.pred

    def num_launches(self, decoded: bool) -> int:
        se = sum(1 for op in self.stem_ops if isinstance(op, SeGateOp))  # two kernels per SE gate
        return (len(self.inputs) + len(self.pre_loads) + len(self.neck_ops) + len(self.stem_ops) + se +
                len(self.tower_ops) + len(self.pred_dec_ops if decoded else self.pred_raw_ops))

    def buffer(self, name: str) -> torch.Tensor:
        return self._bufs[name]


class GraphedPath:
    """CUDA-graph capture of the device-resident part of a detect step for one plan (fixed weights, batch, input size):
    neck -> head (fused decode) -> score filter -> NMS, ~65 launches replayed as ONE graph launch (SURVEY.md section 7
    step 9).  The programmatic-dependent-launch attribute of every kernel becomes a programmatic edge of the graph.
    Inputs are the plan's own NHWC buffers (filled by plan.load_features or the chained backbone before `replay`), outputs
    the fixed (det, count, keep_index) tensors of this object - a replay overwrites them, so a caller that keeps results
    across steps copies them (or alternates two GraphedPath objects)."""

    def __init__(self, plan: "FFAPathPlan", nms, conf_thres: float, nms_thres: float, strategy: str = "auto_cuda"):
        self.plan, self.nms = plan, nms
        dev = plan.device
        self.det = torch.empty((nms.batch, nms.max_det, 7), dtype=torch.float32, device=dev)
        self.count = torch.empty((nms.batch,), dtype=torch.int32, device=dev)
        self.keep_index = torch.empty((nms.batch, nms.max_det), dtype=torch.int32, device=dev)
        self.args = (float(conf_thres), float(nms_thres), strategy)
        self.graph = None
        self._capture()

    def _body(self):
        self.plan.run_neck()
        self.plan.run_head("det")
        self.nms.launch(self.plan.pred, self.args[0], self.args[1], self.args[2], cls_logits=self.plan.det_cls_logits,
                        out=(self.det, self.count, self.keep_index))

    def _capture(self):
        side = torch.cuda.Stream(device=self.plan.device)
        side.wait_stream(torch.cuda.current_stream(self.plan.device))
        with torch.cuda.stream(side):
            self._body()            # warm-up on the capture stream: lazy one-time initialisation must not be captured
            self._body()
        torch.cuda.current_stream(self.plan.device).wait_stream(side)
        torch.cuda.synchronize(self.plan.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            self._body()
        self.graph = g

    def replay(self):
        self.graph.replay()
        return self.det, self.count
