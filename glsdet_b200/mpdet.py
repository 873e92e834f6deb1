"""Drop-ins for the MP-Det neck / head of yolox-ufp (BASELINE configs[2]):
`FPN` (mmdet/models/necks/fpn.py:10-203) and `MPHead` (mmdet/models/dense_heads/mp_head.py:21-154 on top of
GFLHead, dense_heads/gfl_head.py:52-471), with mmdet's constructor arguments, forward signatures (tuples of NCHW fp32
maps) and state_dict keys (`lateral_convs.i.conv`, `fpn_convs.i.conv`; `cls_convs.i.{conv,gn}`, `reg_convs.i.{conv,gn}`,
`gfl_cls_conv`, `gfl_reg`, `scales.l.scale`, `proxies`, `_embedding`, `_pos_embedding_ptr`, `_proxies_prob`,
`integral.project`).  Inference only; the math runs in libglsdet_b200.so:

  * FPN: lateral 1x1 convs with the top-down `+= upsample(coarser)` folded into the epilogue (bf16 post-residual read
    at (y >> 1, x >> 1)), 3x3 output convs, stride-2 extra convs 'on_output' (odd maps are copied into a zero-padded
    even buffer first: the extra row / column is exactly the conv's zero padding);
  * MPHead: towers shared by the five levels = 3x3 conv (tcgen05) + `glsdet_group_norm_relu`; `gfl_cls_conv` -> fp32
    features -> `glsdet_proxy_scores`; `gfl_reg` with the level's `Scale` folded into weights and bias ->
    `glsdet_gfl_decode` (integral + distance2bbox);
  * get_bboxes: `glsdet_gfl_select` per level (score filter + top nms_pre) -> mmcv-style batched NMS
    (`utils_bbox.batched_nms(..., "mmcv")`) -> max_per_img.

mmcv is absent in this environment, so these classes register in the minimal registries of mmdet_face.py (and in a real
mmdet through `register_into_mmdet`); the restated oracle (oracle/mmdet_ref.py) is PARITY UNPINNED for this path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _native as N
from .ops import ConvOp, RectCopyOp, View, nchw_to_nhwc, nhwc_to_nchw
from .utils_bbox import batched_nms


class _ConvHolder(nn.Module):
    """mmcv ConvModule parameter layout: `conv` (+ `gn`)."""

    def __init__(self, cin, cout, k, stride=1, gn=False):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, bias=not gn)
        if gn:
            self.gn = nn.GroupNorm(32, cout)

    def forward(self, x):
        raise RuntimeError("this layer is executed by libglsdet_b200.so through its parent FPN / MPHead")


class _Scale(nn.Module):
    def __init__(self, v=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(float(v)))


class _Integral(nn.Module):
    def __init__(self, reg_max):
        super().__init__()
        self.register_buffer("project", torch.linspace(0, reg_max, reg_max + 1))


class _Native(nn.Module):
    def __init__(self):
        super().__init__()
        self._plans: Dict[tuple, object] = {}
        self.register_load_state_dict_post_hook(lambda m, inc: m._plans.clear())

    def _apply(self, fn, *a, **k):
        self._plans.clear()
        return super()._apply(fn, *a, **k)

    def train(self, mode: bool = True):
        if mode:
            raise RuntimeError("glsdet_b200 modules are inference-only")
        return super().train(False)


class FPN(_Native):
    """necks/fpn.py: FPN(in_channels, out_channels, num_outs, start_level=0, add_extra_convs='on_output').
    forward(inputs: tuple of len(in_channels) NCHW fp32 maps) -> tuple of num_outs NCHW fp32 maps."""

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, add_extra_convs=False,
                 relu_before_extra_convs=False, no_norm_on_lateral=False, conv_cfg=None, norm_cfg=None, act_cfg=None,
                 upsample_cfg=dict(mode="nearest"), init_cfg=None):
        super().__init__()
        if end_level != -1 or norm_cfg is not None or act_cfg is not None or relu_before_extra_convs:
            raise NotImplementedError("FPN drop-in: end_level=-1, no norm / activation, no relu_before_extra_convs")
        if add_extra_convs not in ("on_output",) and num_outs > len(in_channels) - start_level:
            raise NotImplementedError("FPN drop-in: extra levels need add_extra_convs='on_output'")
        self.in_channels, self.out_channels, self.num_outs, self.start_level = list(in_channels), out_channels, num_outs, start_level
        n_lat = len(in_channels) - start_level
        self.lateral_convs = nn.ModuleList([_ConvHolder(in_channels[i + start_level], out_channels, 1) for i in range(n_lat)])
        self.fpn_convs = nn.ModuleList([_ConvHolder(out_channels, out_channels, 3) for _ in range(n_lat)] +
                                       [_ConvHolder(out_channels, out_channels, 3, 2) for _ in range(num_outs - n_lat)])
        nn.Module.train(self, False)

    def _plan(self, inputs):
        key = tuple(tuple(t.shape) for t in inputs) + (str(inputs[0].device),)
        if key not in self._plans:
            self._plans = {key: _FPNPlan(self, inputs)}
        return self._plans[key]

    @torch.no_grad()
    def forward(self, inputs: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, ...]:
        assert len(inputs) == len(self.in_channels)
        return self._plan(inputs).run(inputs)


class _FPNPlan:
    def __init__(self, m: FPN, inputs):
        dev = inputs[0].device
        sl, C = m.start_level, m.out_channels
        used = inputs[sl:]
        B = used[0].shape[0]
        self.inputs_idx = list(range(sl, len(inputs)))
        self.src = [torch.empty((B, t.shape[2], t.shape[3], t.shape[1]), dtype=torch.bfloat16, device=dev) for t in used]
        n_lat = len(used)
        self.lat = [torch.empty((B, t.shape[2], t.shape[3], C), dtype=torch.bfloat16, device=dev) for t in used]
        self.ops: List = []
        none = N.ACT_NONE
        for i in range(n_lat - 1, -1, -1):   # coarse to fine: lat[i] = conv(x_i) + upsample(lat[i+1])  (fpn.py:156-175)
            cm = m.lateral_convs[i].conv
            kw = {}
            if i < n_lat - 1:
                h, w = self.lat[i].shape[1:3]
                if (self.lat[i + 1].shape[1] * 2, self.lat[i + 1].shape[2] * 2) != (h, w):
                    raise NotImplementedError("FPN drop-in: every level must be exactly half of the level below it")
                kw = dict(post_res=View(self.lat[i + 1]), post_shift=1)
            self.ops.append(ConvOp([View(self.src[i])], cm.weight.detach().float(), cm.bias.detach().float(), ksize=1,
                                   act=none, out=View(self.lat[i]), **kw))
        self.outs = []
        for i in range(n_lat):
            cm = m.fpn_convs[i].conv
            o = torch.empty_like(self.lat[i])
            self.ops.append(ConvOp([View(self.lat[i])], cm.weight.detach().float(), cm.bias.detach().float(), ksize=3,
                                   act=none, out=View(o)))
            self.outs.append(o)
        self._pads = []
        for i in range(n_lat, m.num_outs):   # 'on_output': 3x3 stride-2 convs on the previous output (fpn.py:196-202)
            prev = self.outs[-1]
            h, w = prev.shape[1:3]
            if h % 2 or w % 2:   # the missing row / column is the conv's own zero padding
                pad = torch.zeros((B, h + h % 2, w + w % 2, C), dtype=torch.bfloat16, device=dev)
                self.ops.append(RectCopyOp(View(prev), View(pad), B, [(0, 0, 0, 0, 0, 0, h, w)]))
                self._pads.append(pad)
                prev = pad
            cm = m.fpn_convs[i].conv
            o = torch.empty((B, (h + 1) // 2, (w + 1) // 2, C), dtype=torch.bfloat16, device=dev)
            self.ops.append(ConvOp([View(prev)], cm.weight.detach().float(), cm.bias.detach().float(), ksize=3, stride=2,
                                   act=none, out=View(o)))
            self.outs.append(o)

    def run(self, inputs):
        for idx, dst in zip(self.inputs_idx, self.src):
            nchw_to_nhwc(inputs[idx].float().contiguous(), View(dst))
        for op in self.ops:
            op.launch()
        res = []
        for o in self.outs:
            t = torch.empty((o.shape[0], o.shape[3], o.shape[1], o.shape[2]), dtype=torch.float32, device=o.device)
            nhwc_to_nchw(View(o), t)
            res.append(t)
        return tuple(res)


class MPHead(_Native):
    """dense_heads/mp_head.py MPHead(num_classes, in_channels, stacked_convs=4, feat_channels=256, reg_max=16,
    proxies_list=[...], gamma=10, strides via anchor_generator, test_cfg).  forward(feats) -> (cls_scores, bbox_preds)
    lists of NCHW fp32 maps like MPHead.forward in eval mode; get_bboxes(...) -> [(dets [n, 5], labels [n])]."""

    def __init__(self, num_classes, in_channels, feat_channels=256, stacked_convs=4, reg_max=16, num_words=200, beta=0,
                 gamma=10, proxies_list=(2, 3, 2, 5, 4, 8, 8, 4, 3, 3), strides=(8, 16, 32, 64, 128), norm_cfg=None,
                 anchor_generator=None, test_cfg=None, train_cfg=None, **kwargs):
        super().__init__()
        assert num_classes == len(proxies_list)
        if anchor_generator is not None:
            strides = tuple(anchor_generator.get("strides", strides))
        self.num_classes, self.in_channels, self.feat_channels = num_classes, in_channels, feat_channels
        self.stacked_convs, self.reg_max, self.gamma = stacked_convs, reg_max, float(gamma)
        self.proxies_list, self.strides = tuple(proxies_list), tuple(strides)
        self.test_cfg = dict(test_cfg) if test_cfg is not None else None
        self.cls_convs = nn.ModuleList([_ConvHolder(in_channels if i == 0 else feat_channels, feat_channels, 3, gn=True)
                                        for i in range(stacked_convs)])
        self.reg_convs = nn.ModuleList([_ConvHolder(in_channels if i == 0 else feat_channels, feat_channels, 3, gn=True)
                                        for i in range(stacked_convs)])
        self.gfl_cls_conv = nn.Conv2d(feat_channels, feat_channels, 3, padding=1)
        self.gfl_reg = nn.Conv2d(feat_channels, 4 * (reg_max + 1), 3, padding=1)
        self.scales = nn.ModuleList([_Scale(1.0) for _ in strides])
        self.register_buffer("_embedding", torch.randn(num_classes + 1, num_words, feat_channels))
        self.register_buffer("_pos_embedding_ptr", torch.zeros(num_classes + 1, dtype=torch.long))
        self.proxies = nn.Parameter(torch.randn(sum(proxies_list), feat_channels))
        self.register_buffer("_proxies_prob", torch.cat([torch.full((n,), 1.0 / n) for n in proxies_list]))
        self.integral = _Integral(reg_max)
        nn.Module.train(self, False)

    def _plan(self, feats):
        key = tuple(tuple(t.shape) for t in feats) + (str(feats[0].device),)
        if key not in self._plans:
            self._plans = {key: _MPHeadPlan(self, feats)}
        return self._plans[key]

    @torch.no_grad()
    def forward(self, feats: Sequence[torch.Tensor]):
        plan = self._plan(feats)
        plan.run(feats, img_shape=None)
        return plan.cls_maps(), plan.bbox_maps()

    @torch.no_grad()
    def detect(self, feats: Sequence[torch.Tensor], img_metas, cfg=None, rescale=False):
        """forward + get_bboxes without materialising the NCHW maps: [(dets [n, 5], labels [n])] per image."""
        plan = self._plan(feats)
        return plan.get_bboxes(feats, img_metas, self.test_cfg if cfg is None else dict(cfg), rescale=rescale)

    @torch.no_grad()
    def get_bboxes(self, cls_scores, bbox_preds, score_factors=None, img_metas=None, cfg=None, rescale=False, with_nms=True, **kwargs):
        """base_dense_head.py:62-150 get_bboxes(cls_scores, bbox_preds, img_metas=..., cfg=..., rescale=..., with_nms=True)
        on maps returned by forward(): [(dets [n, 5], labels [n])] per image."""
        if not with_nms or score_factors is not None:
            raise NotImplementedError("MPHead.get_bboxes: with_nms=True and no score factors (GFL has none)")
        shapes = tuple((cls_scores[0].shape[0], self.in_channels) + tuple(c.shape[2:]) for c in cls_scores) + (str(cls_scores[0].device),)
        plan = self._plans.get(shapes)
        if plan is None:
            raise RuntimeError("MPHead.get_bboxes: call forward() on the same feature shapes first (the plan owns the buffers)")
        return plan.get_bboxes(None, img_metas, self.test_cfg if cfg is None else dict(cfg), rescale=rescale, maps=(cls_scores, bbox_preds))

    @torch.no_grad()
    def simple_test(self, feats, img_metas, rescale=False):
        """dense_test_mixins.py simple_test_bboxes: forward + get_bboxes; the fused path skips the NCHW maps."""
        return self.detect(feats, img_metas, rescale=rescale)


def _scale4(sf):
    """img_meta['scale_factor'] -> (w, h, w, h) divisors (mmdet stores a 4-vector or a scalar)."""
    try:
        v = [float(x) for x in sf]
    except TypeError:
        v = [float(sf)] * 4
    return v if len(v) == 4 else [v[0], v[1 % len(v)], v[0], v[1 % len(v)]]


class _MPHeadPlan:
    def __init__(self, m: MPHead, feats):
        self.m = m
        dev = feats[0].device
        lib = N.load()
        self.lib = lib
        B, Cin = feats[0].shape[:2]
        fc, nc, bins = m.feat_channels, m.num_classes, m.reg_max + 1
        self.B, self.dev = B, dev
        self.hw = [(t.shape[2], t.shape[3]) for t in feats]
        self.A = sum(h * w for h, w in self.hw)
        self.row0 = [sum(h * w for h, w in self.hw[:l]) for l in range(len(feats))]
        self.rows = torch.empty((B, self.A, nc), dtype=torch.float32, device=dev)     # raw class scores
        self.boxes = torch.empty((B, self.A, 4), dtype=torch.float32, device=dev)     # decoded xyxy
        self.gn_scratch = torch.zeros(int(lib.glsdet_group_norm_scratch_floats(B, fc)), dtype=torch.float32, device=dev)
        centers = torch.nn.functional.normalize(m.proxies.detach().float(), p=2, dim=1).contiguous()
        starts = [0]
        for n in m.proxies_list:
            starts.append(starts[-1] + n)
        self.centers, self.cls_start = centers, torch.tensor(starts, dtype=torch.int32, device=dev)
        self.n_prox_pad = (centers.shape[0] + 15) // 16 * 16
        self.proxy_w = torch.zeros((self.n_prox_pad, fc, 1, 1), dtype=torch.float32, device=dev)
        self.proxy_w[:centers.shape[0], :, 0, 0] = centers
        self.levels = []
        reg_ld = (4 * bins + 15) // 16 * 16
        self.reg_ld = reg_ld
        for l, t in enumerate(feats):
            h, w = self.hw[l]
            x = torch.empty((B, h, w, Cin), dtype=torch.bfloat16, device=dev)
            ta = torch.empty((B, h, w, fc), dtype=torch.bfloat16, device=dev)
            tb = torch.empty_like(ta)
            ra = torch.empty_like(ta)
            rb = torch.empty_like(ta)
            feat16 = torch.empty((B, h, w, fc), dtype=torch.bfloat16, device=dev)     # class features (bf16: operand of the proxy GEMM)
            sims = torch.empty((B, h, w, self.n_prox_pad), dtype=torch.float32, device=dev)
            reg32 = torch.zeros((B, h, w, reg_ld), dtype=torch.float32, device=dev)
            ops = []

            def tower(convs, first, a, b_):
                src, bufs = first, (a, b_)
                for i, cm in enumerate(convs):
                    dst = bufs[i % 2]
                    ops.append(("conv", ConvOp([View(src)], cm.conv.weight.detach().float(), None, ksize=3, act=N.ACT_NONE,
                                               out=View(dst))))
                    ops.append(("gn", (dst, cm.gn.weight.detach().float().contiguous(), cm.gn.bias.detach().float().contiguous(),
                                       float(cm.gn.eps))))
                    src = dst
                return src

            cf = tower(m.cls_convs, x, ta, tb)
            rf = tower(m.reg_convs, x, ra, rb)
            ops.append(("conv", ConvOp([View(cf)], m.gfl_cls_conv.weight.detach().float(), m.gfl_cls_conv.bias.detach().float(),
                                       ksize=3, act=N.ACT_NONE, out=View(feat16))))
            # forward_proxy (mp_head.py:105-121): similarities to the normalised proxies = a 1x1 conv on the tensor core
            ops.append(("conv", ConvOp([View(feat16)], self.proxy_w, None, ksize=1, act=N.ACT_NONE, out=View(sims))))
            sc = float(m.scales[l].scale.detach())
            ops.append(("conv", ConvOp([View(rf)], m.gfl_reg.weight.detach().float() * sc, m.gfl_reg.bias.detach().float() * sc,
                                       ksize=3, act=N.ACT_NONE, out=View(reg32, 0, 4 * bins))))
            self.levels.append(dict(x=x, feat16=feat16, sims=sims, reg32=reg32, ops=ops, keep=(ta, tb, ra, rb)))
        # candidate buffers of get_bboxes
        self.cap = 0
        self.keys = None

    # The five levels are independent until the candidates are selected.  The 100 x 168 level fills the device; the three
    # coarse levels are chains of ~30 launches of <= 96 CTAs that are bound by launch and pipeline-fill latency, so they run
    # on side streams underneath the fine levels (GLSDET_MPDET_STREAMS=0 puts everything on the caller's stream).
    _LEVEL_STREAM = (0, 1, 2, 2, 2)

    def _streams(self):
        if self.__dict__.get("_side") is None:
            import os
            on = os.environ.get("GLSDET_MPDET_STREAMS", "1") != "0"
            self._side = [torch.cuda.Stream(self.dev) for _ in range(2)] if on else []
            self._fork = torch.cuda.Event()
            self._join = [torch.cuda.Event() for _ in self._side]
            n = int(self.lib.glsdet_group_norm_scratch_floats(self.B, self.m.feat_channels))
            self._gn_scratch = [self.gn_scratch] + [torch.zeros(n, dtype=torch.float32, device=self.dev) for _ in self._side]
        return self._side

    def run(self, feats, img_shape):
        side = self._streams()
        main = torch.cuda.current_stream()
        if not side:
            for l in range(len(self.levels)):
                self._run_level(l, feats, img_shape, self.gn_scratch)
            return
        self._fork.record(main)
        order = sorted(range(len(self.levels)), key=lambda l: -self._LEVEL_STREAM[min(l, 4)])
        used = set()
        for l in order:               # coarse chains are issued first: they start while the host still issues the fine ones
            si = self._LEVEL_STREAM[min(l, 4)]
            if si == 0:
                self._run_level(l, feats, img_shape, self._gn_scratch[0])
                continue
            st = side[si - 1]
            if si not in used:
                st.wait_event(self._fork)
                used.add(si)
            with torch.cuda.stream(st):
                self._run_level(l, feats, img_shape, self._gn_scratch[si])
        for si in sorted(used):
            self._join[si - 1].record(side[si - 1])
            main.wait_event(self._join[si - 1])

    def _run_level(self, l, feats, img_shape, gn_scratch):
        m, lib, st = self.m, self.lib, N.stream_ptr(None)
        nc, bins = m.num_classes, m.reg_max + 1
        lv = self.levels[l]
        nchw_to_nhwc(feats[l].float().contiguous(), View(lv["x"]))
        h, w = self.hw[l]
        for kind, op in lv["ops"]:
            if kind == "conv":
                op.launch()
            else:
                buf, g, b_, eps = op
                N.check(lib.glsdet_group_norm_relu(buf.data_ptr(), self.B, h * w, buf.shape[3], buf.shape[3], 32,
                                                   g.data_ptr(), b_.data_ptr(), eps, gn_scratch.data_ptr(), st),
                        "glsdet_group_norm_relu")
        N.check(lib.glsdet_proxy_aggregate(lv["feat16"].data_ptr(), lv["sims"].data_ptr(), self.n_prox_pad,
                                           self.cls_start.data_ptr(), nc, m.feat_channels, self.B, h * w, m.gamma,
                                           self.rows.data_ptr(), nc, self.A * nc, self.row0[l], st), "glsdet_proxy_aggregate")
        if img_shape is not None:
            N.check(lib.glsdet_gfl_decode(lv["reg32"].data_ptr(), self.reg_ld, bins, self.B, h, w, float(m.strides[l]),
                                          float(img_shape[1]), float(img_shape[0]), self.boxes.data_ptr(), self.A * 4,
                                          self.row0[l], st), "glsdet_gfl_decode")

    def load_maps(self, cls_scores, bbox_preds, img_shape):
        """Given per-level NCHW maps (what MPHead.forward returned) -> the plan's score rows and decoded boxes."""
        m, lib, st = self.m, self.lib, N.stream_ptr(None)
        nc, bins = m.num_classes, m.reg_max + 1
        for l, (lv, (h, w)) in enumerate(zip(self.levels, self.hw)):
            self.rows[:, self.row0[l]:self.row0[l] + h * w] = cls_scores[l].float().permute(0, 2, 3, 1).reshape(self.B, h * w, nc)
            lv["reg32"][..., :4 * bins] = bbox_preds[l].float().permute(0, 2, 3, 1)
            N.check(lib.glsdet_gfl_decode(lv["reg32"].data_ptr(), self.reg_ld, bins, self.B, h, w, float(m.strides[l]),
                                          float(img_shape[1]), float(img_shape[0]), self.boxes.data_ptr(), self.A * 4,
                                          self.row0[l], st), "glsdet_gfl_decode")

    def cls_maps(self):
        nc = self.m.num_classes
        return [self.rows[:, r0:r0 + h * w].reshape(self.B, h, w, nc).permute(0, 3, 1, 2).contiguous()
                for r0, (h, w) in zip(self.row0, self.hw)]

    def bbox_maps(self):
        n = 4 * (self.m.reg_max + 1)
        return [lv["reg32"][..., :n].permute(0, 3, 1, 2).contiguous() for lv in self.levels]

    def get_bboxes(self, feats, img_metas, cfg, rescale=False, maps=None):
        """_get_bboxes_single + _bbox_post_process (gfl_head.py:426-471, base_dense_head.py:276-301); every image of the
        batch must share img_shape (the clamp bounds are launch parameters).  `maps=(cls_scores, bbox_preds)` post-processes
        given NCHW maps (the reference's get_bboxes signature) instead of running the head on `feats`."""
        m, lib, st = self.m, self.lib, N.stream_ptr(None)
        img_shape = tuple(img_metas[0]["img_shape"][:2])
        assert all(tuple(im["img_shape"][:2]) == img_shape for im in img_metas)
        score_thr, nms_pre = float(cfg.get("score_thr", 0.05)), int(cfg.get("nms_pre", 1000))
        iou = float(dict(cfg["nms"]).get("iou_threshold", 0.6))
        max_per_img = int(cfg.get("max_per_img", 100))
        nms_max = dict(cfg["nms"]).get("max_num", -1)       # mmcv nms_cfg 'max_num': keep at most that many
        if nms_max is not None and int(nms_max) > 0:
            max_per_img = min(max_per_img, int(nms_max))
        if maps is None:
            self.run(feats, img_shape)
        else:
            self.load_maps(maps[0], maps[1], img_shape)
        nc = m.num_classes
        cap = nms_pre * len(self.levels)
        big = max(h * w for h, w in self.hw) * nc
        kstride = 1
        while kstride < big:
            kstride <<= 1
        if self.keys is None or self.cap != cap:
            self.cap = cap
            self.keys = torch.empty((self.B, kstride), dtype=torch.int64, device=self.dev)
            self.ccount = torch.zeros((self.B,), dtype=torch.int32, device=self.dev)
            self.cboxes = torch.empty((self.B, cap, 4), dtype=torch.float32, device=self.dev)
            self.cscores = torch.empty((self.B, cap), dtype=torch.float32, device=self.dev)
            self.clabels = torch.empty((self.B, cap), dtype=torch.float32, device=self.dev)
            self.sel_scratch = torch.zeros(int(lib.glsdet_gfl_select_scratch_ints(self.B)), dtype=torch.int32, device=self.dev)
        self.ccount.zero_()
        for l, (h, w) in enumerate(self.hw):
            N.check(lib.glsdet_gfl_select(self.rows.data_ptr(), nc, self.A * nc, self.boxes.data_ptr(), self.A * 4, self.row0[l],
                                          h * w, nc, score_thr, nms_pre, self.B, self.keys.data_ptr(), kstride,
                                          self.ccount.data_ptr(), self.cboxes.data_ptr(), self.cscores.data_ptr(),
                                          self.clabels.data_ptr(), cap, self.sel_scratch.data_ptr(), st), "glsdet_gfl_select")
        # Post-processing without per-image host round trips: every image runs the mmcv-style batched NMS on its FULL
        # candidate buffer (cap rows); rows beyond the image's candidate count are padding - zero boxes (IoU 0 with
        # everything, they never suppress and never raise the coordinate-trick offset) with the lowest score, so they sort
        # behind every real candidate.  One device->host read of the B counts at the end sizes the returned tensors
        # (mmdet returns variable-length (dets, labels) per image, so that read is part of the interface).
        pad = torch.arange(cap, device=self.dev)[None, :] >= self.ccount[:, None]
        self.cboxes.masked_fill_(pad[:, :, None], 0.0)
        self.cscores.masked_fill_(pad, -3.0e38)
        self.clabels.masked_fill_(pad, 0.0)
        if rescale:   # base_dense_head.py:282-283: mlvl_bboxes /= scale_factor, before the NMS
            sf = torch.tensor([list(_scale4(im["scale_factor"])) for im in img_metas], dtype=torch.float32, device=self.dev)
            self.cboxes.div_(sf[:, None, :])
        from .utils_bbox import STRATEGIES
        if getattr(self, "keep", None) is None or self.keep.shape[1] != cap:
            self.keep = torch.empty((self.B, cap), dtype=torch.int32, device=self.dev)
            self.kcount = torch.zeros((self.B,), dtype=torch.int32, device=self.dev)
            nbytes = int(lib.glsdet_batched_nms_batch_workspace_bytes(self.B, cap, nc))
            self.nms_ws = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
        ids = self.clabels.to(torch.int32)
        # all images in ONE pipeline launch sequence (each image stays its own NMS problem: per-image coordinate-trick maximum)
        N.check(lib.glsdet_batched_nms_ids_batch(self.cboxes.data_ptr(), self.cscores.data_ptr(), self.clabels.data_ptr(),
                                                 ids.data_ptr(), float(nc - 1), nc, self.B, cap, iou, STRATEGIES["mmcv"],
                                                 self.nms_ws.data_ptr(), self.nms_ws.numel(), self.keep.data_ptr(),
                                                 self.kcount.data_ptr(), st), "glsdet_batched_nms_ids_batch")
        keep = self.keep.long()
        real = (keep < self.ccount[:, None].long()) & (torch.arange(cap, device=self.dev)[None, :] < self.kcount[:, None])
        n_real = real.sum(1).clamp(max=max_per_img)                     # kept real candidates lead the score order
        top = keep[:, :max_per_img].clamp(0, cap - 1)
        dets = torch.cat([torch.gather(self.cboxes, 1, top[:, :, None].expand(-1, -1, 4)), torch.gather(self.cscores, 1, top)[:, :, None]], 2)
        labels = torch.gather(self.clabels, 1, top).long()
        counts = n_real.cpu().tolist()
        return [(dets[b, :counts[b]], labels[b, :counts[b]]) for b in range(self.B)]


from .mmdet_face import HEADS, NECKS  # noqa: E402  (same minimal registries; mmdet's own when it is installed)

NECKS.register_module(module=FPN, force=True)
HEADS.register_module(module=MPHead, force=True)
