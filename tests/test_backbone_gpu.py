"""GPU parity of the native CSPDarknet backbone (glsdet_b200/backbone.py) through the C ABI.

Tolerances follow tests/test_path_gpu.py: relative l2 error <= 2e-2 (the plain BASELINE.json bound) against the fp32
golden vectors of the real reference / the fp32 oracle, and <= 2e-2 against the storage-precision emulation of the same
graph (oracle.ref_path.csp_darknet_bf16; kernel error only); the two byte-moving kernels are bit-exact.  The backbone
stores fp16 (glsdet_b200/_native.py::storage_dtype): measured 0.1 - 0.8 %, bf16 storage alone would cost 2 - 6 %.
"""
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


def test_focus_kernel_bit_exact(native_lib, cuda_device):
    from glsdet_b200.ops import FocusOp

    g = torch.Generator().manual_seed(0)
    for b, h, w in ((2, 8, 12), (1, 64, 96), (3, 34, 50)):
        x = torch.randn(b, 3, h, w, generator=g).to(cuda_device)
        dst = torch.full((b, h // 2, w // 2, 16), 7.0, dtype=torch.bfloat16, device=cuda_device)
        FocusOp(dst).launch(x)
        ref = torch.cat((x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]), 1)   # darknet.py:16-20
        assert torch.equal(dst[..., :12].permute(0, 3, 1, 2), ref.to(torch.bfloat16))
        assert dst[..., 12:].abs().max().item() == 0


def test_focus_rejects_odd_sizes(native_lib, cuda_device):
    from glsdet_b200 import _native as N

    x = torch.zeros(1, 3, 7, 8, device=cuda_device)
    dst = torch.zeros(1, 3, 4, 16, dtype=torch.bfloat16, device=cuda_device)
    rc = native_lib.glsdet_focus_nchw_f32_to_nhwc_bf16(x.data_ptr(), dst.data_ptr(), 1, 7, 8, 0, N.stream_ptr())
    assert rc != 0 and b"even" in native_lib.glsdet_last_error()


def test_spp_pool_kernel_bit_exact(native_lib, cuda_device):
    from glsdet_b200.ops import SppPoolOp

    g = torch.Generator().manual_seed(1)
    for b, c, h, w in ((2, 16, 4, 5), (1, 64, 32, 32), (2, 8, 17, 32), (1, 24, 1, 3), (1, 8, 40, 50)):
        x = torch.randn(b, c, h, w, generator=g).to(cuda_device).to(torch.bfloat16)
        cat = torch.zeros(b, h, w, 4 * c, dtype=torch.bfloat16, device=cuda_device)
        cat[..., :c] = x.permute(0, 2, 3, 1)
        SppPoolOp(cat, c).launch()
        assert torch.equal(cat[..., :c].permute(0, 3, 1, 2), x)
        for i, ks in enumerate((5, 9, 13)):       # darknet.py:28
            ref = F.max_pool2d(x.float(), ks, 1, ks // 2).to(torch.bfloat16)
            assert torch.equal(cat[..., (i + 1) * c:(i + 2) * c].permute(0, 3, 1, 2), ref), (b, c, h, w, ks)


def test_kx_folded_conv_matches_torch(native_lib, cuda_device):
    """3x3 conv over 16-channel pixels with the kx taps folded into the channel view (glsdet_conv_desc.ksize_w = 1; the
    Focus stem) against fp64 torch on the same bf16-rounded operands; sizes with ragged tiles and one-row images."""
    from glsdet_b200 import _native as N
    from glsdet_b200.ops import ConvOp, FoldedView, View, fold_kx_weight

    g = torch.Generator().manual_seed(2)
    for b, h, w, n in ((2, 24, 40, 32), (1, 64, 64, 64), (3, 5, 7, 32), (1, 1, 130, 48)):
        x = torch.randn(b, 12, h, w, generator=g).to(cuda_device).to(torch.bfloat16)
        wt = (torch.randn(n, 12, 3, 3, generator=g) / 10).to(cuda_device)
        bias = torch.randn(n, generator=g).to(cuda_device)
        flat = torch.zeros(b * h * (w + 2) * 16 + 64, dtype=torch.bfloat16, device=cuda_device)
        fv = FoldedView(flat, b, h, w, 16)
        fv.interior()[..., :12] = x.permute(0, 2, 3, 1)
        out = torch.zeros(b, h, w, n, dtype=torch.bfloat16, device=cuda_device)
        op = ConvOp([fv], fold_kx_weight(wt, 16), bias, ksize=3, ksize_w=1, act=N.ACT_SILU, out=View(out))
        op.launch()
        ref = F.conv2d(x.double(), wt.to(torch.bfloat16).double(), bias.double(), padding=1)
        ref = ref * torch.sigmoid(ref)
        assert_close_rel(out.permute(0, 3, 1, 2), ref.float(), tol=1e-2, what=f"folded {b}x{h}x{w}->{n}")


def _net(cls, sd, dev):
    net = cls(10, "s")
    net.load_state_dict(sd, strict=True)
    return net.to(dev).eval()


def test_backbone_against_reference_golden(native_lib, cuda_device):
    """CSPDarknet.forward (native plan) against the real reference's feature maps."""
    from glsdet_b200.synthetic import synthetic_state_dict
    from glsdet_b200.yolox_ffa import YoloBody

    z = np.load(GOLD / "backbone_s.npz")
    sd = synthetic_state_dict(10, "s", seed=0, flavour="calibrated")
    net = _net(YoloBody, sd, cuda_device)
    n0 = native_lib.glsdet_launch_count()
    feats = net.backbone.backbone(torch.from_numpy(z["image"]).to(cuda_device))
    assert native_lib.glsdet_launch_count() - n0 >= 30        # the native plan ran, not the PyTorch layers
    assert list(feats) == ["dark2", "dark3", "dark4", "dark5"]
    emu = dict(zip(("dark2", "dark3", "dark4", "dark5"), ref_path.csp_darknet_bf16(sd, torch.from_numpy(z["image"]))))
    for name, f in feats.items():
        assert f.dtype == torch.float32 and tuple(f.shape) == z[name].shape
        ref = torch.from_numpy(z[name])
        tol = 2e-2
        assert_close_rel(f, ref, tol=tol, what=name, max_factor=6.0)
        assert rel_l2(f, emu[name]) <= 2e-2, (name, rel_l2(f, emu[name]))


def test_image_to_logits_against_reference_golden(native_lib, cuda_device):
    """YoloBody.forward(image): backbone -> neck -> FFA -> head chained in NHWC bf16 (no NCHW fp32 round trip)."""
    from glsdet_b200.synthetic import synthetic_state_dict
    from glsdet_b200.yolox_ffa import YoloBody

    z = np.load(GOLD / "backbone_s.npz")
    sd = synthetic_state_dict(10, "s", seed=0, flavour="calibrated")
    net = _net(YoloBody, sd, cuda_device)
    x = torch.from_numpy(z["image"])
    logits = net(x.to(cuda_device))
    plan = net._fused_plan(x.to(cuda_device))
    assert plan is not None and plan.backbone is not None
    # storage-precision emulation of the same graph (weights and every activation rounded to their storage type, fp32 math)
    feats_q = ref_path.csp_darknet_bf16(sd, x)
    emu = ref_path.neck_head_bf16(sd, feats_q)
    for i, t in enumerate(logits):
        ref = torch.from_numpy(z[f"logits{i}"])
        err = rel_l2(t, ref)
        assert err <= 2e-2, (i, err)
        assert rel_l2(t, emu[i]) <= 2e-2, (i, rel_l2(t, emu[i]))


def test_backbone_1024_against_oracle_and_batch_invariance(native_lib, cuda_device):
    """BASELINE configs[1] size: features of two 1024x1024 images against the fp32 oracle; running the same image in a
    batch of 3 gives bit-identical features."""
    from glsdet_b200.backbone import BackbonePlan
    from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict

    sd = synthetic_state_dict(10, "s", seed=0, flavour="calibrated")
    x = synthetic_images(2, 1024, 1024, seed=5)
    ref = ref_path.csp_darknet(sd, x)
    plan = BackbonePlan(sd, 2, (1024, 1024), device=cuda_device)
    plan.run(x.to(cuda_device))
    got = plan.features_nchw()
    emu = ref_path.csp_darknet_bf16(sd, x)
    for name, r, e in zip(("dark2", "dark3", "dark4", "dark5"), ref, emu):
        assert_close_rel(got[name], r, tol=2e-2, what=name, max_factor=8.0)
        assert rel_l2(got[name], e) <= 2e-2, (name, rel_l2(got[name], e))
    plan3 = BackbonePlan(sd, 3, (1024, 1024), device=cuda_device)
    x3 = torch.cat([x[1:], x, ])[:3].contiguous()     # images (1, 0, 1)
    plan3.run(x3.to(cuda_device))
    got3 = plan3.features_nchw()
    for name in got:
        assert torch.equal(got3[name][1], got[name][0]) and torch.equal(got3[name][0], got[name][1]), name


def test_detect_from_image_matches_detect_from_features(native_lib, cuda_device):
    """detect(image) = backbone chained into the fused path; its rows equal detect_features on the plan's own
    feature maps converted to the reference layout (same kernels, same inputs -> bit-exact)."""
    from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict
    from glsdet_b200.yolox_ffa import YoloBody

    sd = synthetic_state_dict(10, "s", seed=0, flavour="calibrated")
    net = _net(YoloBody, sd, cuda_device)
    x = synthetic_images(2, 256, 320, seed=9).to(cuda_device)
    det, cnt = net.detect(x, conf_thres=0.01, nms_thres=0.65)
    det, cnt = det.clone(), cnt.clone()
    assert int(cnt.min()) > 0
    plan = net._fused_plan(x)
    feats = plan.backbone.features_nchw()
    # bf16 values survive the fp32 round trip exactly, so the neck sees identical inputs
    det2, cnt2 = net.detect_features([feats[k] for k in ("dark2", "dark3", "dark4", "dark5")], conf_thres=0.01, nms_thres=0.65)
    assert torch.equal(cnt, cnt2)
    for b in range(2):
        assert torch.equal(det[b, :int(cnt[b])], det2[b, :int(cnt[b])])


def test_stock_model_from_image(native_lib, cuda_device):
    """models/base/yolox.py topology (three levels): forward(image) through the chained native backbone against the
    oracle restatement."""
    from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict
    from glsdet_b200.yolox_base import YoloBody

    sd = synthetic_state_dict(10, "s", seed=21, flavour="calibrated", variant="stock")
    net = _net(YoloBody, sd, cuda_device)
    x = synthetic_images(1, 160, 192, seed=3)
    logits = net(x.to(cuda_device))
    feats = ref_path.csp_darknet(sd, x)[1:]
    ref = ref_path.stock_neck_head(sd, feats)
    emu = ref_path.stock_neck_head(sd, ref_path.csp_darknet_bf16(sd, x)[1:], bf16=True)
    for i, t in enumerate(logits):
        assert rel_l2(t, ref[i]) <= 2e-2, (i, rel_l2(t, ref[i]), rel_l2(emu[i], ref[i]))


def test_backbone_fp32_accuracy_mode(native_lib, cuda_device):
    """set_precision('fp32'): Focus / convs / pools in fp32 (SIMT) - the 1e-3 bar of BASELINE configs[0] extended to the
    backbone: golden feature maps and image -> logits of the real reference."""
    from glsdet_b200.synthetic import synthetic_state_dict
    from glsdet_b200.yolox_ffa import YoloBody

    z = np.load(GOLD / "backbone_s.npz")
    sd = synthetic_state_dict(10, "s", seed=0, flavour="calibrated")
    net = _net(YoloBody, sd, cuda_device).set_precision("fp32")
    x = torch.from_numpy(z["image"]).to(cuda_device)
    n0 = native_lib.glsdet_launch_count()
    feats = net.backbone.backbone(x)
    assert native_lib.glsdet_launch_count() - n0 >= 30
    for name, f in feats.items():
        assert_close_rel(f, torch.from_numpy(z[name]), tol=1e-3, what="fp32 " + name)
    logits = net(x)
    for i, t in enumerate(logits):
        assert_close_rel(t, torch.from_numpy(z[f"logits{i}"]), tol=1e-3, what=f"fp32 logits{i}")


def test_backbone_tiny_width(native_lib, cuda_device):
    """phi='tiny' (base width 24: channel counts that are multiples of 8 but not of 16) against the oracle."""
    from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict
    from glsdet_b200.yolox_ffa import YoloBody

    sd = synthetic_state_dict(3, "tiny", seed=3, flavour="calibrated")
    net = YoloBody(3, "tiny")
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda_device).eval()
    x = synthetic_images(2, 128, 160, seed=2)
    feats = net.backbone.backbone(x.to(cuda_device))
    ref = ref_path.csp_darknet(sd, x)
    emu = ref_path.csp_darknet_bf16(sd, x)
    for (name, f), r, e in zip(feats.items(), ref, emu):
        assert_close_rel(f, r, tol=2e-2, what="tiny " + name, max_factor=8.0)


def test_backbone_rejects_cpu_tensor(native_lib):
    from glsdet_b200.yolox_ffa import CSPDarknet

    with pytest.raises(RuntimeError, match="no CPU"):
        CSPDarknet(0.33, 0.5)(torch.zeros(1, 3, 64, 64))


def test_uint8_entry_bit_identical_to_host_preprocessing(native_lib, cuda_device):
    """detect_uint8(uint8 HWC batch): the normalisation of models/core/utils.py:47-51 and the transpose of yolo.py:134 run
    inside the Focus kernel; the Focus output and the detections equal those of the float path fed with the host-side
    preprocessing (numpy, exactly the reference's statements)."""
    from glsdet_b200.ops import FocusOp, FoldedView
    from glsdet_b200.synthetic import synthetic_state_dict
    from glsdet_b200.yolox_ffa import YoloBody

    rng = np.random.default_rng(5)
    b, h, w = 2, 128, 160
    img = rng.integers(0, 256, (b, h, w, 3), dtype=np.uint8)
    x = np.array(img, dtype="float32")                      # yolo.py:134
    x /= 255.0                                              # utils.py:48-50
    x -= np.array([0.485, 0.456, 0.406])
    x /= np.array([0.229, 0.224, 0.225])
    x = torch.from_numpy(np.ascontiguousarray(np.transpose(x, (0, 3, 1, 2)))).to(cuda_device)
    fa = FoldedView(torch.zeros(b * (h // 2) * (w // 2 + 2) * 16 + 64, dtype=torch.bfloat16, device=cuda_device), b, h // 2, w // 2, 16)
    fb = FoldedView(torch.zeros_like(fa.t), b, h // 2, w // 2, 16)
    FocusOp(fa).launch(x)
    FocusOp(fb).launch_u8(torch.from_numpy(img).to(cuda_device), (0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    assert torch.equal(fa.t, fb.t)
    sd = synthetic_state_dict(10, "s", seed=0, flavour="calibrated")
    net = _net(YoloBody, sd, cuda_device)
    det_a, cnt_a = net.detect(x, conf_thres=0.01, nms_thres=0.65)
    det_a, cnt_a = det_a.clone(), cnt_a.clone()
    det_b, cnt_b = net.detect_uint8(torch.from_numpy(img).to(cuda_device), conf_thres=0.01, nms_thres=0.65)
    assert torch.equal(cnt_a, cnt_b)
    for i in range(b):
        assert torch.equal(det_a[i, :int(cnt_a[i])], det_b[i, :int(cnt_b[i])])   # rows beyond count are unspecified


def test_large_model_from_image(native_lib, cuda_device):
    """phi='l' (BASELINE configs[3] widths: base 64, up to 1024 channels = several N blocks per conv, SPP over 512 hidden
    channels, three bottlenecks per stage) from an image with a non-square size, against the oracle."""
    from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict
    from glsdet_b200.yolox_ffa import YoloBody

    sd = synthetic_state_dict(3, "l", seed=6, flavour="calibrated")
    net = YoloBody(3, "l")
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda_device).eval()
    x = synthetic_images(1, 160, 288, seed=6)
    feats = net.backbone.backbone(x.to(cuda_device))
    ref = ref_path.csp_darknet(sd, x)
    emu = ref_path.csp_darknet_bf16(sd, x)
    for (name, f), r, e in zip(feats.items(), ref, emu):
        assert_close_rel(f, r, tol=2e-2, what="l " + name, max_factor=8.0)
    logits = net(x.to(cuda_device))
    ref_l = ref_path.neck_head(sd, ref)
    emu_l = ref_path.neck_head_bf16(sd, emu)
    for i, t in enumerate(logits):
        assert rel_l2(t, ref_l[i]) <= 2e-2, (i, rel_l2(t, ref_l[i]), rel_l2(emu_l[i], ref_l[i]))


def test_p2_model_chains_the_backbone(native_lib, cuda_device):
    """GLSDet P2 has no per-input pre-loads, so forward(image) chains the native backbone into its plan; the result is
    bit-identical to going through the NCHW fp32 feature maps (bf16 values survive that round trip exactly)."""
    from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict
    from glsdet_b200.yolo_patch_nonlocal_plus import YoloBody

    sd = synthetic_state_dict(10, "s", seed=0, flavour="calibrated", variant="p2")
    net = _net(YoloBody, sd, cuda_device)
    x = synthetic_images(2, 256, 320, seed=8).to(cuda_device)
    a = [t.clone() for t in net(x)]
    plan = net._fused_plan(x)
    assert plan is not None and plan.backbone is not None
    b = net.forward_features(net.backbone.features(x))
    for u, v in zip(a, b):
        assert torch.equal(u, v)


def test_kx_folded_pair_conv_matches_torch(native_lib, cuda_device):
    """Pair form of the folded stem conv (two output pixels per GEMM row, fold_kx_pair_weight) against fp64 torch."""
    from glsdet_b200 import _native as N
    from glsdet_b200.ops import ConvOp, FoldedView, View, fold_kx_pair_weight

    g = torch.Generator().manual_seed(4)
    for b, h, w, n in ((2, 24, 40, 32), (1, 64, 64, 32), (3, 5, 6, 64), (1, 1, 130, 32)):
        x = torch.randn(b, 12, h, w, generator=g).to(cuda_device).to(torch.bfloat16)
        wt = (torch.randn(n, 12, 3, 3, generator=g) / 10).to(cuda_device)
        bias = torch.randn(n, generator=g).to(cuda_device)
        flat = torch.zeros(b * h * (w + 2) * 16 + 64, dtype=torch.bfloat16, device=cuda_device)
        FoldedView(flat, b, h, w, 16).interior()[..., :12] = x.permute(0, 2, 3, 1)
        pair = FoldedView(flat, b, h, w // 2, 32, row_pitch=(w + 2) * 16)
        out = torch.zeros(b, h, w, n, dtype=torch.bfloat16, device=cuda_device)
        wp, bp = fold_kx_pair_weight(wt, bias, 16)
        ConvOp([pair], wp, bp, ksize=3, ksize_w=1, act=N.ACT_SILU, out=View(out.view(b, h, w // 2, 2 * n))).launch()
        ref = F.conv2d(x.double(), wt.to(torch.bfloat16).double(), bias.double(), padding=1)
        ref = ref * torch.sigmoid(ref)
        assert_close_rel(out.permute(0, 3, 1, 2), ref.float(), tol=1e-2, what=f"pair {b}x{h}x{w}->{n}")


def test_pair_stride2_conv_matches_torch(native_lib, cuda_device):
    """3x3 stride-2 conv over 32-channel pixels in the pixel-pair form (glsdet_conv_desc.ksize_w = 2, pair_stride2_weight)
    against fp64 torch; ragged tiles and several sizes."""
    from glsdet_b200 import _native as N
    from glsdet_b200.ops import ConvOp, View, pair_stride2_weight

    g = torch.Generator().manual_seed(6)
    for b, h, w, n in ((2, 32, 48, 64), (1, 64, 64, 64), (3, 6, 20, 128), (1, 2, 260, 64)):
        x = torch.randn(b, 32, h, w, generator=g).to(cuda_device).to(torch.bfloat16)
        wt = (torch.randn(n, 32, 3, 3, generator=g) / 12).to(cuda_device)
        bias = torch.randn(n, generator=g).to(cuda_device)
        xin = x.permute(0, 2, 3, 1).contiguous()                       # [B, H, W, 32]
        out = torch.zeros(b, h // 2, w // 2, n, dtype=torch.bfloat16, device=cuda_device)
        op = ConvOp([View(xin.view(b, h, w // 2, 64))], pair_stride2_weight(wt), bias, ksize=3, stride=2, ksize_w=2,
                    act=N.ACT_SILU, out=View(out))
        op.launch()
        ref = F.conv2d(x.double(), wt.to(torch.bfloat16).double(), bias.double(), stride=2, padding=1)
        ref = ref * torch.sigmoid(ref)
        assert_close_rel(out.permute(0, 3, 1, 2), ref.float(), tol=1e-2, what=f"pair stride 2 {b}x{h}x{w}->{n}")
