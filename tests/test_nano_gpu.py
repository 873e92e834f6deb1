"""GPU parity of phi = 'nano' (depthwise-separable DWConv blocks, yolox-drone/models/base/baseConv.py:22-30) through
the C ABI: the depthwise kernel (csrc/dwconv.cu) against F.conv2d(groups = C), the 1x1 conv + fused prediction conv the
nano towers end in, and the whole models (GLSDet P0 and the stock YOLOX) against tests/golden/nano_cases.npz - outputs of
the REAL reference modules recorded by tests/golden/make_golden_nano.py.  The seeded weights are calibrated with
trained-like BatchNorm shifts (tools/calibrate_synthetic.py --bn-beta 1.0, as for the YOLOX-l case of test_path_gpu.py): a
random net with beta = 0 sits in the chaotic regime where every layer multiplies a relative perturbation by 1.1, and the
DWConv variant has twice the layers (measured on the oracle's storage emulation: 3.6 % from the image with beta = 0 even in
fp16 storage, 0.16 % with beta = 1).  Bars: relative l2 <= 2e-2 in 16-bit storage,
<= 1e-3 in the fp32 accuracy mode (BASELINE.json), NMS rows bit-exact when fed the golden predictions.
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _helpers import assert_close_rel
from oracle import ref_path

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("generic", [False, True, "tiled"], ids=["k3", "generic", "k3_tiled"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32], ids=["bf16", "f16", "f32"])
def test_depthwise_kernel_matches_torch(dtype, generic, native_lib, cuda_device, monkeypatch):
    """glsdet_dwconv vs F.conv2d(groups = C) on the same rounded inputs: stride 1 / 2, k = 3 / 5, odd sizes (zero padding
    at every border, ragged x strips), one-pixel images, channel windows on both sides, every activation; the 3x3
    specialisations (row-streaming, and row-tiled with GLSDET_DW_STREAM=0) and the generic kernel (GLSDET_DW_GENERIC=1)."""
    from glsdet_b200 import _native as N
    from glsdet_b200.ops import DepthwiseOp, View

    if generic == "tiled":
        monkeypatch.setenv("GLSDET_DW_STREAM", "0")   # the row-tiled 3x3 kernel instead of the row-streaming one
    elif generic:
        monkeypatch.setenv("GLSDET_DW_GENERIC", "1")
    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    cases = [(2, 16, 24, 40, 3, 1, "silu"), (1, 32, 33, 17, 3, 2, "silu"), (3, 64, 7, 9, 5, 1, "relu"),
             (1, 8, 1, 1, 3, 1, "lrelu"), (2, 128, 16, 16, 3, 2, "none"), (1, 256, 5, 64, 3, 1, "silu"),
             (2, 24, 9, 7, 3, 1, "silu"), (1, 16, 6, 11, 3, 2, "relu"), (1, 8, 3, 2, 3, 2, "silu"), (1, 40, 2, 1, 3, 1, "none"),
             # tall maps: row bands of 16 / 8 output rows with a ragged last band (the row-streaming kernel's slot rotation)
             (4, 64, 600, 256, 3, 1, "silu"), (4, 64, 1201, 255, 3, 2, "silu")]
    for b, c, h, w, k, s, act in cases:
        x = torch.randn(b, c, h, w, generator=g).to(dev).to(dtype)
        wt = (torch.randn(c, 1, k, k, generator=g) / k).to(dev)
        bias = torch.randn(c, generator=g).to(dev)
        # channel windows: the source sits at offset 8 of a wider buffer, the output at offset 16
        src = torch.full((b, h, w, c + 24), 3.0, dtype=dtype, device=dev)
        src[..., 8:8 + c] = x.permute(0, 2, 3, 1)
        pad = (k - 1) // 2
        ho, wo = (h + 2 * pad - k) // s + 1, (w + 2 * pad - k) // s + 1
        out = torch.full((b, ho, wo, c + 16), 7.0, dtype=dtype, device=dev)
        DepthwiseOp(View(src, 8, c), wt, bias, stride=s, act=N.ACT_BY_NAME[act], out=View(out, 16, c)).launch()
        torch.cuda.synchronize()
        y = F.conv2d(x.double(), wt.double(), bias.double(), stride=s, padding=pad, groups=c)
        ref = {"silu": lambda t: t * torch.sigmoid(t), "relu": torch.relu, "none": lambda t: t,
               "lrelu": lambda t: F.leaky_relu(t, 0.1)}[act](y)
        got = out[..., 16:].permute(0, 3, 1, 2).double()
        assert got.shape == ref.shape, (got.shape, ref.shape)
        eps = {torch.bfloat16: 2.0 ** -8, torch.float16: 2.0 ** -11, torch.float32: 4e-6}[dtype]
        assert (got - ref).abs().max().item() <= eps * max(1.0, ref.abs().max().item()) * 1.01, (b, c, h, w, k, s, act)
        assert (out[..., :16] == 7.0).all()   # nothing outside the window is written


def test_depthwise_rejects_bad_arguments(native_lib, cuda_device):
    from glsdet_b200 import _native as N

    x = torch.zeros(1, 4, 4, 12, dtype=torch.bfloat16, device=cuda_device)
    w = torch.zeros(9, 12, device=cuda_device)
    b = torch.zeros(12, device=cuda_device)
    rc = native_lib.glsdet_dwconv(x.data_ptr(), 12, 0, x.data_ptr(), 12, 0, 1, 4, 4, 12, 3, 1, w.data_ptr(), b.data_ptr(),
                                  N.ACT_SILU, N.DT_BF16, N.stream_ptr())
    assert rc != 0 and b"multiples of 8" in native_lib.glsdet_last_error()
    rc = native_lib.glsdet_dwconv(x.data_ptr(), 16, 0, x.data_ptr(), 16, 0, 1, 4, 4, 16, 4, 1, w.data_ptr(), b.data_ptr(),
                                  N.ACT_SILU, N.DT_BF16, N.stream_ptr())
    assert rc != 0 and b"ksize" in native_lib.glsdet_last_error()


@pytest.mark.parametrize("n_tower,n_pred,pred_act", [(64, 5, "box"), (64, 10, "none_nchw"), (128, 3, "sigmoid")])
def test_pointwise_conv_with_fused_prediction(n_tower, n_pred, pred_act, native_lib, cuda_device):
    """1x1 conv + SiLU + fused 1x1 prediction conv (+ decode): what the second tower DWConv of phi = 'nano' ends in (its
    pconv half carries the prediction conv the 3x3 tower conv carries in the other models)."""
    from glsdet_b200 import _native as N
    from glsdet_b200.ops import ConvOp, View

    dev = cuda_device
    g = torch.Generator().manual_seed(n_tower + n_pred)
    B, H, W, Cin = 2, 24, 40, n_tower
    r16 = lambda t: t.to(torch.bfloat16).float()
    x = r16(torch.randn(B, Cin, H, W, generator=g)).to(dev)
    w = r16(torch.randn(n_tower, Cin, 1, 1, generator=g) / Cin ** 0.5).to(dev)
    bias = torch.randn(n_tower, generator=g).to(dev)
    wp = (torch.randn(n_pred, n_tower, 1, 1, generator=g) / n_tower ** 0.5).to(dev)
    bp = torch.randn(n_pred, generator=g).to(dev)
    t = F.conv2d(x, w, bias)
    t = t * torch.sigmoid(t)
    y = F.conv2d(r16(t), r16(wp), bp)   # tensor-core prediction path: activated tile and weights are 16-bit operands
    rows = y.permute(0, 2, 3, 1).reshape(B, H * W, n_pred)
    xin = View(x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))
    nch = n_pred + 3
    if pred_act == "none_nchw":
        out = torch.full((B, nch, H, W), float("nan"), device=dev)
        ConvOp([xin], w, bias, ksize=1, act=N.ACT_SILU, out=out, out_mode=N.OUT_NCHW_F32, out_ld=nch, out_coff=2,
               out_batch_stride=nch * H * W, pred_weight=wp, pred_bias=bp, pred_act=N.ACT_NONE).launch()
        torch.cuda.synchronize()
        assert torch.allclose(out[:, 2:2 + n_pred], y, rtol=6e-3, atol=6e-3), (out[:, 2:2 + n_pred] - y).abs().max()
        assert torch.isnan(out[:, :2]).all() and torch.isnan(out[:, 2 + n_pred:]).all()
        return
    out = torch.full((B, H * W, nch), float("nan"), device=dev)
    act = {"sigmoid": N.ACT_SIGMOID, "box": N.ACT_YOLOX_BOX}[pred_act]
    stride, in_h, in_w = 8.0, H * 8.0, W * 8.0
    ConvOp([xin], w, bias, ksize=1, act=N.ACT_SILU, out=out, out_mode=N.OUT_NHWC_F32, out_ld=nch, out_coff=1,
           out_batch_stride=H * W * nch, pred_weight=wp, pred_bias=bp, pred_act=act, dec=(stride, in_w, in_h)).launch()
    torch.cuda.synchronize()
    got = out[:, :, 1:1 + n_pred]
    if pred_act == "sigmoid":
        ref = torch.sigmoid(rows)
    else:
        gy, gx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
        gx, gy = gx.reshape(1, -1).float(), gy.reshape(1, -1).float()
        ref = torch.stack([(rows[..., 0] + gx) * stride / in_w, (rows[..., 1] + gy) * stride / in_h,
                           torch.exp(rows[..., 2]) * stride / in_w, torch.exp(rows[..., 3]) * stride / in_h,
                           torch.sigmoid(rows[..., 4])], dim=-1)
    assert torch.allclose(got, ref, rtol=6e-3, atol=6e-3), (got - ref).abs().max()
    assert torch.isnan(out[:, :, 0]).all() and torch.isnan(out[:, :, 1 + n_pred:]).all()


def _case(case):
    z = np.load(GOLD / "nano_cases.npz")
    m = json.loads((GOLD / "nano_meta.json").read_text())[case]
    sd = ref_path.synthetic_state_dict(m["nc"], "nano", seed=m["seed"], flavour="calibrated", variant=m["variant"])
    if case == "p0":
        from glsdet_b200.yolox_ffa import YoloBody
    else:
        from glsdet_b200.yolox_base import YoloBody
    net = YoloBody(m["nc"], "nano")
    net.load_state_dict(sd, strict=True)
    return z, m, sd, net


@pytest.mark.parametrize("case", ["p0", "stock"])
def test_nano_model_matches_reference_golden(case, native_lib, cuda_device):
    """YoloBody(nc, 'nano'): image -> backbone -> neck -> head on the native path (depthwise kernel + tcgen05 pointwise convs)
    against the real reference's feature maps, neck outputs and logits; then decode + NMS rows."""
    from glsdet_b200.utils_bbox import decode_outputs, non_max_suppression

    z, m, sd, net = _case(case)
    net = net.to(cuda_device).eval()
    x = torch.from_numpy(z[f"{case}_image"]).to(cuda_device)
    names = ("dark2", "dark3", "dark4", "dark5") if case == "p0" else ("dark3", "dark4", "dark5")
    feats = net.backbone.backbone(x)
    for n in names:
        assert_close_rel(feats[n], torch.from_numpy(z[f"{case}_{n}"]), what=f"{case} {n}")
    # neck + head from the REFERENCE's features (the metric's segment), then from the image (chained backbone)
    ref_feats = [torch.from_numpy(z[f"{case}_{n}"]).to(cuda_device) for n in names]
    neck = net.backbone.forward_features(ref_feats)
    for i, t in enumerate(neck):
        assert_close_rel(t, torch.from_numpy(z[f"{case}_neck{i}"]), what=f"{case} neck{i}")
    for tag, logits in (("features", net.forward_features(ref_feats)), ("image", net(x))):
        for i, t in enumerate(logits):
            ref = torch.from_numpy(z[f"{case}_logits{i}"])
            assert t.shape == ref.shape
            assert rel_l2(t, ref) <= 2e-2, (case, tag, i, rel_l2(t, ref))
    # post-processing on the golden logits: decode within 1e-5, NMS rows bit-exact when fed the golden predictions
    hw = [m["in_h"], m["in_w"]]
    gl = [torch.from_numpy(z[f"{case}_logits{i}"]).to(cuda_device) for i in range(len(logits))]
    pred = decode_outputs(gl, hw)
    assert torch.allclose(pred.cpu(), torch.from_numpy(z[f"{case}_pred"]), rtol=1e-5, atol=1e-5)
    got = non_max_suppression(torch.from_numpy(z[f"{case}_pred"]).to(cuda_device), m["nc"], hw, np.array(hw), False,
                              m["conf"], m["nms_thr"])
    for b in range(m["batch"]):
        assert np.array_equal(got[b], z[f"{case}_nms{b}"]), (case, b)
    # fused detect path runs and returns finite rows
    det, cnt = net.detect_features(ref_feats, conf_thres=m["conf"], nms_thres=m["nms_thr"])
    torch.cuda.synchronize()
    assert int(cnt.min()) > 0 and torch.isfinite(det[0, :int(cnt[0])]).all()


def test_nano_fp32_accuracy_mode(native_lib, cuda_device):
    """fp32 accuracy mode of the nano P0 model (depthwise kernel on fp32 tensors, SIMT fp32 pointwise convs): 1e-3 relative
    (BASELINE.json) against the real reference's logits, from the image."""
    z, m, sd, net = _case("p0")
    net = net.to(cuda_device).eval()
    net.set_precision("fp32")
    x = torch.from_numpy(z["p0_image"]).to(cuda_device)
    for i, t in enumerate(net(x)):
        ref = torch.from_numpy(z[f"p0_logits{i}"])
        assert rel_l2(t, ref) <= 1e-3, (i, rel_l2(t, ref))


def test_mmdet_face_use_depthwise(native_lib, cuda_device):
    """mmdet YOLOXPAFPN / YOLOXHead with use_depthwise=True (mmcv DepthwiseSeparableConvModule; YOLOX-nano config shapes)
    on the native path: the stock nano weights renamed to mmdet's keys give the logits of the REAL models/base/yolox.py
    nano model (golden), and get_bboxes runs on them."""
    from glsdet_b200.mmdet_face import YOLOXHead, YOLOXPAFPN
    from oracle import mmdet_ref

    z = np.load(GOLD / "nano_cases.npz")
    m = json.loads((GOLD / "nano_meta.json").read_text())["stock"]
    sd = ref_path.synthetic_state_dict(m["nc"], "nano", seed=m["seed"], flavour="calibrated", variant="stock")
    neck_sd, head_sd = mmdet_ref.drone_to_mmdet_keys(sd)
    neck = YOLOXPAFPN(in_channels=[64, 128, 256], out_channels=64, num_csp_blocks=1, use_depthwise=True)
    head = YOLOXHead(num_classes=m["nc"], in_channels=64, feat_channels=64, use_depthwise=True,
                     test_cfg=dict(score_thr=0.01, nms=dict(type="nms", iou_threshold=0.65)))
    neck.load_state_dict(neck_sd, strict=True)
    head.load_state_dict(head_sd, strict=True)
    neck, head = neck.to(cuda_device).eval(), head.to(cuda_device).eval()
    feats = tuple(torch.from_numpy(z[f"stock_dark{i}"]).to(cuda_device) for i in (3, 4, 5))
    cls, box, obj = head(neck(feats))
    for i in range(3):
        got = torch.cat([box[i], obj[i], cls[i]], 1)
        ref = torch.from_numpy(z[f"stock_logits{i}"])
        assert rel_l2(got, ref) <= 2e-2, (i, rel_l2(got, ref))
    res = head.get_bboxes(cls, box, obj, img_metas=[dict(scale_factor=[1.0, 1.0, 1.0, 1.0])] * m["batch"])
    assert len(res) == m["batch"] and res[0][0].shape[1] == 5 and torch.isfinite(res[0][0]).all()


@pytest.mark.parametrize("case", ["p1", "p2"])
def test_nano_p1_p2_models_match_reference_golden(case, native_lib, cuda_device):
    """phi = 'nano' of the GLSDet P1 (models/new/yolox10.py) and P2 (yolo_patch_nonlocal_plus.py) models - DWConv in the
    neck's bu_convs / Bottlenecks, the cross-level head's up / tower convs, plus the patch non-local attention at nano width
    - against the REAL reference's logits: 16-bit path within 2e-2 (from the reference's features and from the image), fp32
    accuracy mode within 1e-3."""
    import importlib

    z = np.load(GOLD / "nano_cases.npz")
    m = json.loads((GOLD / "nano_meta.json").read_text())[case]
    sd = ref_path.synthetic_state_dict(m["nc"], "nano", seed=m["seed"], flavour="calibrated", variant=m["variant"])
    mod = importlib.import_module("glsdet_b200.yolox10" if case == "p1" else "glsdet_b200.yolo_patch_nonlocal_plus")
    net = mod.YoloBody(m["nc"], "nano")
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda_device).eval()
    x = torch.from_numpy(z[f"{case}_image"]).to(cuda_device)
    names = ("dark2", "dark3", "dark4", "dark5") if case == "p1" else ("dark3", "dark4", "dark5")
    ref_feats = [torch.from_numpy(z[f"{case}_{n}"]).to(cuda_device) for n in names]
    refs = [torch.from_numpy(z[f"{case}_logits{i}"]) for i in range(3)]
    for tag, logits in (("features", net.forward_features(ref_feats)), ("image", net(x))):
        for i, t in enumerate(logits):
            assert t.shape == refs[i].shape
            assert rel_l2(t, refs[i]) <= 2e-2, (case, tag, i, rel_l2(t, refs[i]))
    det, cnt = net.detect_features(ref_feats, conf_thres=m["conf"], nms_thres=m["nms_thr"])
    torch.cuda.synchronize()
    assert int(cnt.min()) > 0 and torch.isfinite(det[0, :int(cnt[0])]).all()
    net.set_precision("fp32")
    for i, t in enumerate(net.forward_features(ref_feats)):
        assert rel_l2(t, refs[i]) <= 1e-3, (case, "fp32", i, rel_l2(t, refs[i]))
