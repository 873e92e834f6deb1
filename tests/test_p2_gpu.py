"""GPU parity of the GLSDet P2 slice (yolo_patch_nonlocal_plus.py: Patch_Conv_NonLocal / Patch_Conv inputs to the PAFPN,
7x7 / 5x5 / 3x3 identity convs, stock head) through the C ABI.  Tolerances as in test_path_gpu.py."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _helpers import TOL, assert_close_rel
from oracle import ref_path

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
META = json.loads((GOLD / "meta.json").read_text())["p2"]


def _bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("k,cin,n_out,hw", [(5, 64, 64, (24, 40)), (7, 128, 128, (40, 24)), (5, 256, 256, (16, 16)),
                                            (7, 128, 128, (128, 128))])
def test_large_kernel_convs(k, cin, n_out, hw, native_lib, cuda_device):
    """5x5 / 7x7 stride-1 convs (identity convs of P2) on the tcgen05 kernel: k ky-taps per A stage."""
    from glsdet_b200 import _native as N
    from glsdet_b200.ops import ConvOp, View

    dev = cuda_device
    g = torch.Generator().manual_seed(k * 100 + cin)
    H, W = hw
    x = _bf(torch.randn(2, cin, H, W, generator=g)).to(dev)
    w = _bf(torch.randn(n_out, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev)
    bias = torch.randn(n_out, generator=g).to(dev)
    ref = F.conv2d(x.double().cpu(), w.double().cpu(), bias.double().cpu(), padding=k // 2).float()
    out = torch.full((2, H, W, n_out), float("nan"), device=dev, dtype=torch.bfloat16)
    ConvOp([View(x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))], w, bias, ksize=k, act=N.ACT_NONE, out=View(out)).launch()
    torch.cuda.synchronize()
    got = out.float().permute(0, 3, 1, 2).cpu()
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()


def test_rect_copy_and_nhwc_transpose(native_lib, cuda_device):
    from glsdet_b200.ops import NhwcTransposeOp, RectCopyOp, View

    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    B, H, W, C = 2, 8, 12, 32
    x = torch.randn(B, H, W, C + 16, generator=g).to(dev).to(torch.bfloat16)
    h2, w2 = H // 2, W // 2
    xp = torch.zeros(4 * B, h2, w2, C, device=dev, dtype=torch.bfloat16)
    pos = ((0, 0), (0, 1), (1, 0), (1, 1))
    RectCopyOp(View(x, 8, C), View(xp), B, [(0, py * h2, px * w2, i * B, 0, 0, h2, w2) for i, (py, px) in enumerate(pos)]).launch()
    torch.cuda.synchronize()
    for i, (py, px) in enumerate(pos):
        assert torch.equal(xp[i * B:(i + 1) * B], x[:, py * h2:(py + 1) * h2, px * w2:(px + 1) * w2, 8:8 + C])
    # back into a channel window of a wider buffer
    back = torch.zeros(B, H, W, C + 8, device=dev, dtype=torch.bfloat16)
    RectCopyOp(View(xp), View(back, 8, C), B, [(i * B, 0, 0, 0, py * h2, px * w2, h2, w2) for i, (py, px) in enumerate(pos)]).launch()
    torch.cuda.synchronize()
    assert torch.equal(back[..., 8:], x[..., 8:8 + C]) and back[..., :8].abs().sum() == 0
    T = h2 * w2
    xt = torch.zeros(4 * B, C + 64, 64, device=dev, dtype=torch.bfloat16)
    NhwcTransposeOp(View(xp), xt).launch()
    torch.cuda.synchronize()
    assert torch.equal(xt[:, :C, :T], xp.reshape(4 * B, T, C).transpose(1, 2))
    assert xt[:, C:].abs().sum() == 0 and xt[:, :, T:].abs().sum() == 0


def _net(sd, dev):
    from glsdet_b200.yolo_patch_nonlocal_plus import YoloBody

    net = YoloBody(10, "s")
    net.load_state_dict(sd, strict=True)
    return net.to(dev).eval()


def test_p2_model_matches_reference_golden(native_lib, cuda_device):
    from glsdet_b200.utils_bbox import decode_outputs

    z = np.load(GOLD / f"{META['name']}.npz")
    sd = ref_path.synthetic_state_dict(META["nc"], META["phi"], seed=META["seed"], flavour="calibrated", variant="p2")
    net = _net(sd, cuda_device)
    feats = [torch.from_numpy(z[f"feat{i}"]).to(cuda_device) for i in range(3)]
    cpu_feats = [f.cpu() for f in feats]
    neck = net.backbone.forward_features(feats)
    plan = net.backbone._plan(1, (META["in_h"], META["in_w"]), cuda_device)
    assert_close_rel(plan.buffer("feat1_patch").float().permute(0, 3, 1, 2), torch.from_numpy(z["feat1_patch"]), TOL,
                     "Patch_Conv_NonLocal(dark3)")
    ref_path._EMULATE_BF16 = True
    try:
        with torch.no_grad():
            neck_emu = ref_path.p2_neck(sd, [ref_path._q(f) for f in cpu_feats])
    finally:
        ref_path._EMULATE_BF16 = False
    for i in range(3):
        ref_i = torch.from_numpy(z[f"neck{i}"])
        assert_close_rel(neck[i], ref_i, TOL, f"p2 neck{i}")
    logits = net.forward_features(feats)
    emu = ref_path.p2_neck_head(sd, cpu_feats, bf16=True)
    for i in range(3):
        ref_i = torch.from_numpy(z[f"logits{i}"])
        assert_close_rel(logits[i], ref_i, TOL, f"p2 logits{i}", frac=2e-2)
        assert_close_rel(logits[i], emu[i], 1.5e-2, f"p2 logits{i} vs bf16-storage emulation",
                         max_factor=4.0, frac=5e-2)
    pred_fused = net.decode_features(feats)
    pred_sep = decode_outputs(logits, [META["in_h"], META["in_w"]])
    assert torch.allclose(pred_fused, pred_sep, rtol=1e-5, atol=1e-6)


def test_p2_model_vs_oracle_1024(native_lib, cuda_device):
    """P2-s at 1024 x 1024, batch 2: one image against the oracle, batch invariance, detections."""
    from glsdet_b200.synthetic import synthetic_images

    sd = ref_path.synthetic_state_dict(10, "s", seed=0, flavour="calibrated", variant="p2")
    net = _net(sd, cuda_device)
    feats = ref_path.csp_darknet(sd, synthetic_images(2, 1024, 1024, seed=12))[1:]
    ref = ref_path.p2_neck_head(sd, [f[:1] for f in feats])
    emu = ref_path.p2_neck_head(sd, [f[:1] for f in feats], bf16=True)
    dfeats = [f.to(cuda_device) for f in feats]
    out = net.forward_features(dfeats)
    for i in range(3):
        assert_close_rel(out[i][:1], ref[i], TOL, f"p2 1024 logits{i}", frac=5e-2)
        assert_close_rel(out[i][:1], emu[i], 1.5e-2, f"p2 1024 logits{i} vs bf16-storage emulation", frac=8e-2)
    pred2 = net.decode_features(dfeats).clone()
    one = net.decode_features([f[1:2].contiguous() for f in dfeats])
    assert torch.equal(one[0], pred2[1]), "results must not depend on the batch an image is in"
    det, cnt = net.detect_features(dfeats, conf_thres=0.02, nms_thres=0.65)
    torch.cuda.synchronize()
    assert (cnt.cpu() >= 0).all() and det.shape[0] == 2


@pytest.mark.parametrize("size", [(96, 160), (544, 1024)])
def test_p2_odd_patch_sizes(size, native_lib, cuda_device):
    """Every legal input (a multiple of 32, yolo.py:34-36) splits into EQUAL 2 x 2 patches at dark3 / dark4
    (dark3 = 4m rows -> halves 2m -> stride-2 conv -> m; dark4 = 2m rows -> halves m), but for odd m the patches, the
    left/right/top/bottom halves and the stride-2 outputs have odd sizes: 96 x 160 (m = 3, 5) and config 4's 544 x 1024
    (m = 17: 17 x 32 patches).  Checked against the oracle (which follows Identity_Conv.py:292-384 slice by slice)."""
    from glsdet_b200.synthetic import synthetic_images

    in_h, in_w = size
    sd = ref_path.synthetic_state_dict(10, "s", seed=0, flavour="calibrated", variant="p2")
    net = _net(sd, cuda_device)
    feats = ref_path.csp_darknet(sd, synthetic_images(1, in_h, in_w, seed=31))[1:]
    assert feats[1].shape[2] // 2 % 2 == 1                     # odd patch height at dark4
    ref = ref_path.p2_neck_head(sd, feats)
    out = net.forward_features([f.to(cuda_device) for f in feats])
    for i in range(3):
        assert_close_rel(out[i], ref[i], TOL, f"p2 {in_h}x{in_w} logits{i}", frac=5e-2)
