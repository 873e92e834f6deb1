"""Host logic of the execution plans (glsdet_b200/engine.py, glsdet_b200/backbone.py) checked WITHOUT a GPU.

A plan is a flat list of native launches plus the buffers they read and write; which conv reads which channel window,
which weights are stacked / permuted / folded, where a residual enters - all of that is decided on the host.  Here every
native operator class is replaced (inside this test only) by a small torch model of the operator's documented contract
(include/glsdet_b200.h), the plans are built on the CPU with exactly the product code, and the result is compared with
the oracle.  This pins the GRAPH (wiring, weight transforms, BatchNorm folding, DWConv splitting for phi = 'nano'); the
kernels themselves are pinned by the `-m gpu` tests.  Nothing here is reachable from the product: the operator models
live in this file.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_path

from glsdet_b200 import _native as N


def _act(y, act):
    if act == N.ACT_SILU:
        return y * torch.sigmoid(y)
    if act == N.ACT_RELU:
        return torch.relu(y)
    if act == N.ACT_LRELU:
        return F.leaky_relu(y, 0.1)
    assert act == N.ACT_NONE, act
    return y


def _win(v):
    """NHWC channel window -> NCHW fp32."""
    return v.t[..., v.coff:v.coff + v.c].float().permute(0, 3, 1, 2)


def _store(v, y):
    v.t[..., v.coff:v.coff + y.shape[1]] = y.permute(0, 2, 3, 1).to(v.t.dtype)


def _up(t, shift):
    return t if shift == 0 else F.interpolate(t, scale_factor=2 ** shift, mode="nearest")


class ModelConvOp:
    """glsdet_conv_desc: conv over cat(srcs) (+bias) (+pre_res) -> act (+post_res) [-> fused 1x1 prediction conv] -> store."""

    def __init__(self, srcs, weight, bias, *, ksize, stride=1, act=N.ACT_NONE, out, out_mode=None, out_ld=0, out_coff=0,
                 out_batch_stride=0, pre_res=None, pre_shift=0, post_res=None, post_shift=0, dec=(0.0, 0.0, 0.0),
                 pred_weight=None, pred_bias=None, pred_act=N.ACT_NONE, out_plane_stride=0, out_elem_offset=0, **other):
        assert not other, f"operator model does not cover {sorted(other)}"
        self.__dict__.update(srcs=srcs, w=weight.float(), b=None if bias is None else bias.float(), k=ksize, stride=stride,
                             act=act, out=out, out_mode=out_mode, out_coff=out_coff, pre=pre_res, pre_shift=pre_shift,
                             post=post_res, post_shift=post_shift, pw=pred_weight, pb=pred_bias, pred_act=pred_act,
                             plane=out_plane_stride)
        b, h, w = srcs[0].bhw
        self.flops = 2.0 * b * (h // stride) * (w // stride) * weight.shape[0] * weight.shape[1] * ksize * ksize

    def launch(self, stream=None):
        x = torch.cat([_win(s) for s in self.srcs], 1)
        y = F.conv2d(x, self.w, self.b, stride=self.stride, padding=(self.k - 1) // 2)
        if self.pre is not None:
            y = y + _up(_win(self.pre)[:, :y.shape[1]], self.pre_shift)
        y = _act(y, self.act)
        if self.post is not None:
            y = y + _up(_win(self.post)[:, :y.shape[1]], self.post_shift)
        if self.pw is not None:
            y = F.conv2d(y, self.pw.float().reshape(self.pw.shape[0], -1, 1, 1), self.pb.float())
            assert self.pred_act == N.ACT_NONE, "only the raw-logit variant of the fused prediction conv is modelled"
        if hasattr(self.out, "bhw"):
            _store(self.out, y)
        else:   # raw NCHW fp32 logits [B, 5+nc, h, w]
            assert self.out_mode == N.OUT_NCHW_F32 and self.plane == 0
            self.out[:, self.out_coff:self.out_coff + y.shape[1]] = y


class ModelDepthwiseOp:
    def __init__(self, src, weight, bias, *, stride=1, act=N.ACT_SILU, out):
        self.__dict__.update(src=src, w=weight.float(), b=bias.float(), stride=stride, act=act, out=out)
        self.flops = 0.0

    def launch(self, stream=None):
        k = self.w.shape[-1]
        y = F.conv2d(_win(self.src), self.w, self.b, stride=self.stride, padding=(k - 1) // 2, groups=self.w.shape[0])
        _store(self.out, _act(y, self.act))


class ModelFocusOp:
    def __init__(self, dst):
        self.dst = dst

    def launch(self, image, stream=None):
        x = torch.cat((image[..., ::2, ::2], image[..., 1::2, ::2], image[..., ::2, 1::2], image[..., 1::2, 1::2]), 1)
        self.dst.zero_()
        self.dst[..., :12] = x.permute(0, 2, 3, 1).to(self.dst.dtype)


class ModelSppPoolOp:
    def __init__(self, cat, channels):
        self.cat, self.c = cat, channels

    def launch(self, stream=None):
        c = self.c
        x = self.cat[..., :c].float().permute(0, 3, 1, 2)
        for i, ks in enumerate((5, 9, 13)):
            self.cat[..., (i + 1) * c:(i + 2) * c] = F.max_pool2d(x, ks, 1, ks // 2).permute(0, 2, 3, 1).to(self.cat.dtype)


class ModelSeGateOp:
    def __init__(self, x, w1, w2):
        self.x, self.w1, self.w2 = x, w1.float(), w2.float()
        self.gate = torch.empty((x.bhw[0], x.c), dtype=torch.float32)

    def launch(self, stream=None):
        m = _win(self.x).mean(dim=(2, 3))
        self.gate.copy_(1.0 + torch.sigmoid(torch.relu(m @ self.w1.t()) @ self.w2.t()))


class ModelScaleShuffleOp:
    def __init__(self, x, gate, dst):
        self.x, self.gate, self.dst = x, gate, dst

    def launch(self, stream=None):
        b, h, w = self.x.bhw
        co = self.x.c // 4
        t = self.x.t.float() * self.gate.view(b, 1, 1, -1)             # channels in (i, j, c) order
        t = t.view(b, h, w, 2, 2, co).permute(0, 1, 3, 2, 4, 5).reshape(b, 2 * h, 2 * w, co)
        self.dst.t[..., self.dst.coff:self.dst.coff + co] = t.to(self.dst.t.dtype)


class ModelRectCopyOp:
    """glsdet_rect_copy: rects = (src image0, sy, sx, dst image0, dy, dx, h, w), each applied to `batch` consecutive images."""

    def __init__(self, src, dst, batch, rects):
        self.src, self.dst, self.batch, self.rects = src, dst, batch, list(rects)

    def launch(self, stream=None):
        s, d = self.src, self.dst
        for sb, sy, sx, db, dy, dx, h, w in self.rects:
            for i in range(self.batch):
                d.t[db + i, dy:dy + h, dx:dx + w, d.coff:d.coff + d.c] = s.t[sb + i, sy:sy + h, sx:sx + w, s.coff:s.coff + s.c]


class ModelUpsample2xOp:
    def __init__(self, src, dst):
        self.src, self.dst = src, dst

    def launch(self, stream=None):
        _store(self.dst, F.interpolate(_win(self.src), scale_factor=2, mode="nearest"))


class ModelBGemmF32Op:
    """glsdet_bgemm_f32 in its two uses: "gram" out[b] = alpha * a[b]^T b[b] over the pixels, "apply" out[b] = a[b] m[b]."""

    def __init__(self, mode, a, b, out, alpha=1.0):
        self.mode, self.a, self.b, self.out, self.alpha = mode, a, b, out, alpha
        self.flops = 0.0

    def launch(self, stream=None):
        a = self.a.t[..., self.a.coff:self.a.coff + self.a.c].flatten(1, 2)          # [B, T, Ci]
        if self.mode == "gram":
            bm = self.b.t[..., self.b.coff:self.b.coff + self.b.c].flatten(1, 2)
            self.out.copy_(self.alpha * torch.einsum("bti,btj->bij", a, bm))
        else:
            y = torch.einsum("bti,bij->btj", a, self.b)
            o = self.out
            o.t[..., o.coff:o.coff + o.c] = y.view(o.t.shape[0], o.t.shape[1], o.t.shape[2], -1)


def _nchw_to_nhwc(src, dst, stream=None):
    _store(dst, src)


def _nhwc_to_nchw(src, dst, stream=None):
    dst.copy_(_win(src))


@pytest.fixture
def op_models(monkeypatch):
    from glsdet_b200 import backbone, engine

    monkeypatch.setenv("GLSDET_STEM_UNFOLDED", "1")      # plain 3x3 stem conv over the 16-channel Focus pixels
    monkeypatch.setenv("GLSDET_NO_PAIR_STRIDE2", "1")    # plain stride-2 convs (the pair views are kernel-side layouts)
    for mod in (engine, backbone):
        for name, model in (("ConvOp", ModelConvOp), ("ConvOpF32", ModelConvOp), ("DepthwiseOp", ModelDepthwiseOp),
                            ("FocusOp", ModelFocusOp), ("SppPoolOp", ModelSppPoolOp), ("SeGateOp", ModelSeGateOp),
                            ("ScaleShuffleOp", ModelScaleShuffleOp), ("RectCopyOp", ModelRectCopyOp),
                            ("Upsample2xOp", ModelUpsample2xOp), ("BGemmF32Op", ModelBGemmF32Op), ("nchw_to_nhwc", _nchw_to_nhwc),
                            ("nhwc_to_nchw", _nhwc_to_nchw)):
            if hasattr(mod, name):
                monkeypatch.setattr(mod, name, model)
    return engine, backbone


def _rel(a, b):
    return ((a.float() - b).norm() / b.norm()).item()


@pytest.mark.parametrize("variant,phi,precision", [("ffa", "nano", "fp32"), ("ffa", "nano", "bf16"), ("stock", "nano", "fp32"),
                                                   ("ffa", "s", "fp32"), ("stock", "tiny", "bf16")])
def test_plan_graph_matches_oracle(variant, phi, precision, op_models):
    _plan_graph_case(variant, phi, precision, op_models)


@pytest.mark.parametrize("phi", ["s", "nano"])
def test_head_csp_group_split_wiring(phi, op_models, monkeypatch):
    """GLSDET_HCSP_SPLIT: the stride-4 CSP block of the FFA head built once per group of images on shared scratch buffers
    (batch slices of the input, the output and the residual) gives the same logits."""
    monkeypatch.setenv("GLSDET_HCSP_SPLIT", "2")
    _plan_graph_case("ffa", phi, "bf16", op_models, expect_split=2)


def test_backbone_pair_form_csp_block(op_models, monkeypatch):
    """The dark2 CSPLayer of phi = 's' (32-channel halves) is built in pixel-pair form (backbone.py::_csp_pairs: pair views,
    block-structured weights of ops.pair_*); with the operators modelled in torch it gives the feature maps of the plain
    form (GLSDET_NO_PAIR_CSP=1) up to the 16-bit rounding of the stored tensors, and both match the oracle."""
    _, backbone = op_models
    from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict

    b, h, w = 2, 64, 96
    sd = synthetic_state_dict(3, "s", seed=5, flavour="kaiming", variant="ffa")
    x = synthetic_images(b, h, w, seed=6)
    with torch.no_grad():
        feats = ref_path.csp_darknet(sd, x)
    outs = []
    for plain in (False, True):
        if plain:
            monkeypatch.setenv("GLSDET_NO_PAIR_CSP", "1")
        bb = backbone.BackbonePlan(sd, b, (h, w), device="cpu", prefix="backbone.backbone.", precision="bf16")
        assert bb.pair_csp_blocks == ([] if plain else ["dark2.1"])
        bb.run(x)
        outs.append([bb.outs[n].float().permute(0, 3, 1, 2) for n in backbone.FEATURES])
        for t, r in zip(outs[-1], feats):
            assert _rel(t, r) <= 2e-2
    assert outs[0][0].shape == outs[1][0].shape
    for a, c in zip(outs[0], outs[1]):
        assert _rel(a, c) <= 1e-2
    # the weight transforms alone, in fp64: pair-form convs == the plain convs on the pixel view
    g = torch.Generator().manual_seed(3)
    from glsdet_b200.ops import pair_bias, pair_conv3_weight, pair_pointwise_weight
    xx = torch.randn(2, 8, 6, 10, generator=g, dtype=torch.float64)            # NCHW, W even
    w3 = torch.randn(12, 8, 3, 3, generator=g, dtype=torch.float64)
    b3 = torch.randn(12, generator=g, dtype=torch.float64)
    ref = F.conv2d(xx, w3, b3, padding=1)

    def to_pairs(t):   # NCHW -> pair view NCHW': channels (pixel, c)
        n, c, hh, ww = t.shape
        return t.permute(0, 2, 3, 1).reshape(n, hh, ww // 2, 2 * c).permute(0, 3, 1, 2)

    got = F.conv2d(to_pairs(xx), pair_conv3_weight(w3), pair_bias(b3), padding=1)
    assert torch.allclose(got, to_pairs(ref), atol=1e-12)
    w1 = torch.randn(12, 8, 1, 1, generator=g, dtype=torch.float64)
    got = F.conv2d(to_pairs(xx), pair_pointwise_weight(w1), pair_bias(b3))
    assert torch.allclose(got, to_pairs(F.conv2d(xx, w1, b3)), atol=1e-12)


def _plan_graph_case(variant, phi, precision, op_models, expect_split=1):
    """image -> BackbonePlan -> FFAPathPlan (neck, head, raw logits) with the operators modelled in torch, against
    oracle.ref_path on the same seeded weights: 1e-4 in the fp32 mode, 2e-2 (BASELINE.json's 16-bit bound) with the
    16-bit storage policy applied to the buffers."""
    engine, backbone = op_models
    from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict

    nc, b, h, w = 3, 2, 64, 96
    sd = synthetic_state_dict(nc, phi, seed=5, flavour="kaiming", variant=variant)
    x = synthetic_images(b, h, w, seed=6)
    with torch.no_grad():
        feats = ref_path.csp_darknet(sd, x)
        if variant == "ffa":
            ref = ref_path.neck_head(sd, feats)
        else:
            ref = ref_path.stock_head(sd, ref_path.pafpn_neck(sd, [None] + list(feats[1:]))[1:])
    plan = engine.FFAPathPlan(sd, b, (h, w), nc, device="cpu", variant=variant, precision=precision)
    names = backbone.FEATURES[-len(plan.inputs):]
    bb = backbone.BackbonePlan(sd, b, (h, w), device="cpu", prefix="backbone.backbone.", precision=precision,
                               outs=dict(zip(names, plan.inputs)))
    n_dw = sum(isinstance(op, ModelDepthwiseOp) for op in bb.ops + plan.neck_ops + plan.stem_ops + plan.tower_ops)
    assert (n_dw > 0) == (phi == "nano")
    if expect_split > 1:   # the block's scratch buffers hold one group of images
        assert plan.hcsp_split == expect_split
    bb.run(x)
    tol = 1e-4 if precision == "fp32" else 2e-2
    for name, t, r in zip(names, plan.inputs, feats[-len(names):]):
        assert _rel(t.float().permute(0, 3, 1, 2), r) <= tol, name
    plan.run_neck()
    plan.run_head(False)
    for i, (t, r) in enumerate(zip(plan.logits, ref)):
        assert t.shape == r.shape
        assert _rel(t, r) <= tol, (i, _rel(t, r))


def test_mmdet_depthwise_plans_match_golden(op_models):
    """mmdet face with use_depthwise=True: the neck (+ out_convs) plan and the towers-only plan built from the drop-in
    modules' translated state dicts, against the golden logits of the real stock nano model."""
    import json
    from pathlib import Path

    engine, _ = op_models
    from glsdet_b200.mmdet_face import YOLOXHead, YOLOXPAFPN
    from oracle import mmdet_ref

    gold = Path(__file__).resolve().parent / "golden"
    z = np.load(gold / "nano_cases.npz")
    m = json.loads((gold / "nano_meta.json").read_text())["stock"]
    sd = ref_path.synthetic_state_dict(m["nc"], "nano", seed=m["seed"], flavour="calibrated", variant="stock")
    neck_sd, head_sd = mmdet_ref.drone_to_mmdet_keys(sd)
    neck = YOLOXPAFPN(in_channels=[64, 128, 256], out_channels=64, num_csp_blocks=1, use_depthwise=True)
    head = YOLOXHead(num_classes=m["nc"], in_channels=64, feat_channels=64, use_depthwise=True)
    neck.load_state_dict(neck_sd, strict=True)
    head.load_state_dict(head_sd, strict=True)
    b, hw = m["batch"], (m["in_h"], m["in_w"])
    feats = [torch.from_numpy(z[f"stock_dark{i}"]) for i in (3, 4, 5)]
    pn = engine.FFAPathPlan(neck._plan_state_dict(), b, hw, 1, device="cpu", parts=neck._parts, variant="stock",
                            precision="fp32")
    pn.load_features(feats)
    pn.run_neck()
    pn.run_stems()
    p_k = pn.stem_outputs_nchw()
    ph = engine.FFAPathPlan(head._plan_state_dict(), b, hw, m["nc"], device="cpu", parts=head._parts, variant="stock",
                            decode="mmdet", precision="fp32")
    ph.load_tower_inputs(p_k)
    ph.run_towers(False)
    for i, t in enumerate(ph.logits):
        assert _rel(t, torch.from_numpy(z[f"stock_logits{i}"])) <= 1e-4, i


@pytest.mark.parametrize("variant,hw,phi", [("p1", (128, 192), "s"), ("p1", (96, 160), "s"), ("p2", (128, 192), "s"),
                                            ("p1", (128, 192), "nano"), ("p2", (128, 192), "nano")])
def test_fp32_plans_of_p1_p2_match_oracle(variant, hw, phi, op_models):
    """fp32 accuracy mode of the GLSDet P1 / P2 topologies (patch non-local attention un-folded into theta / phi / g convs and
    two batched fp32 GEMMs, engine.py::_nonlocal_f32): features -> logits against oracle.ref_path at 1e-4.  The second P1
    size has odd level sizes (unequal 2x2 patch splits, Non_local_family.py:230-233)."""
    engine, _ = op_models
    from glsdet_b200.synthetic import synthetic_state_dict

    nc, b = 3, 2
    h, w = hw
    sd = synthetic_state_dict(nc, phi, seed=7, flavour="kaiming", variant=variant)
    g = torch.Generator().manual_seed(8)
    wm = 2 if phi == "s" else 1     # width 0.5 / 0.25 (nano: every k > 1 conv of neck and head is a DWConv)
    chans = tuple(c * wm for c in ((32, 64, 128, 256) if variant == "p1" else (64, 128, 256)))
    strides = (4, 8, 16, 32) if variant == "p1" else (8, 16, 32)
    feats = [torch.randn(b, c, h // s_, w // s_, generator=g) for c, s_ in zip(chans, strides)]
    ref = ref_path.p1_neck_head(sd, feats) if variant == "p1" else ref_path.p2_neck_head(sd, feats)
    plan = engine.FFAPathPlan(sd, b, hw, nc, device="cpu", variant=variant, precision="fp32")
    assert not plan.pre_loads
    out = plan.forward_logits(feats)
    for i, (t, r) in enumerate(zip(out, ref)):
        assert t.shape == r.shape
        assert _rel(t, r) <= 1e-4, (variant, i, _rel(t, r))
