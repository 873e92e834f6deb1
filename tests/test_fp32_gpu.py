"""fp32 accuracy mode (precision="fp32": SIMT fp32 kernels, csrc/fp32_path.cu) against the fp32 reference.

BASELINE.json: "feature maps and logits within 1e-3 relative in fp32"; configs[0] = GLSDet YOLOX-s neck+head forward,
one synthetic 640x640 image, random-init fp32.  Bounds here: relative l2 error <= 1e-3 AND every element within
1e-3 * max|ref| (measured: about 1e-6, fp32 summation order only).
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_path

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
META = json.loads((GOLD / "meta.json").read_text())
TOL32 = 1e-3


def _close(got, ref, what, tol=TOL32):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert torch.isfinite(got).all(), what
    rel = ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()
    mx = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
    assert rel <= tol and mx <= tol, f"{what}: rel l2 {rel:.3g}, max {mx:.3g} (bound {tol})"
    return rel


CASES = [
    # name, B, H, W, cins, N, k, stride, act
    ("1x1", 2, 9, 13, [48], 40, 1, 1, "silu"),
    ("3x3", 1, 17, 12, [24], 70, 3, 1, "silu"),
    ("3x3s2", 2, 16, 20, [32], 33, 3, 2, "relu"),
    ("cat3x3", 1, 10, 10, [16, 24], 20, 3, 1, "none"),
    ("5x5", 1, 12, 9, [8], 8, 5, 1, "lrelu"),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_f32_matches_torch(case, native_lib, cuda_device):
    from glsdet_b200 import _native as N
    from glsdet_b200.ops import ConvOpF32, View

    name, B, H, W, cins, n_out, k, stride, act = case
    dev = cuda_device
    g = torch.Generator().manual_seed(len(name) * 7 + k)
    xs = [torch.randn(B, c, H, W, generator=g).to(dev) for c in cins]
    w = (torch.randn(n_out, sum(cins), k, k, generator=g) / (sum(cins) * k * k) ** 0.5).to(dev)
    bias = torch.randn(n_out, generator=g).to(dev)
    # float64 CPU reference (cuDNN may pick Winograd / FFT algorithms whose fp32 error is 1e-4 .. 1e-3)
    ref = F.conv2d(torch.cat(xs, 1).double().cpu(), w.double().cpu(), bias.double().cpu(), stride=stride, padding=(k - 1) // 2)
    ref = {"silu": lambda t: t * torch.sigmoid(t), "relu": torch.relu, "none": lambda t: t,
           "lrelu": lambda t: F.leaky_relu(t, 0.1)}[act](ref)
    Ho, Wo = ref.shape[2:]
    post = torch.randn(B, Ho, Wo, n_out, generator=g).to(dev)
    pre = torch.randn(B, max(Ho // 2, 1), max(Wo // 2, 1), n_out, generator=g).to(dev) if Ho % 2 == 0 and Wo % 2 == 0 else None
    srcs = [View(x.permute(0, 2, 3, 1).contiguous()) for x in xs]
    out = torch.full((B, Ho, Wo, n_out + 8), float("nan"), device=dev)
    ConvOpF32(srcs, w, bias, ksize=k, stride=stride, act=N.ACT_BY_NAME[act], out=View(out, 4, n_out)).launch()
    torch.cuda.synchronize()
    _close(out[..., 4:4 + n_out].permute(0, 3, 1, 2), ref, name, 1e-5)
    assert torch.isnan(out[..., :4]).all() and torch.isnan(out[..., 4 + n_out:]).all()
    # residuals (+ upsampled fp32 partial sums before the activation), NCHW output
    if pre is not None:
        raw = F.conv2d(torch.cat(xs, 1).double().cpu(), w.double().cpu(), bias.double().cpu(), stride=stride, padding=(k - 1) // 2)
        raw = raw + F.interpolate(pre.permute(0, 3, 1, 2).double().cpu(), scale_factor=2, mode="nearest")
        ref2 = torch.relu(raw) + post.permute(0, 3, 1, 2).double().cpu()
        out2 = torch.full((B, n_out + 3, Ho, Wo), float("nan"), device=dev)
        ConvOpF32(srcs, w, bias, ksize=k, stride=stride, act=N.ACT_RELU, out=out2, out_mode=N.OUT_NCHW_F32,
                  out_ld=n_out + 3, out_coff=2, out_batch_stride=(n_out + 3) * Ho * Wo, pre_res=View(pre), pre_shift=1,
                  post_res=View(post), post_shift=0).launch()
        torch.cuda.synchronize()
        _close(out2[:, 2:2 + n_out], ref2, name + " residuals", 1e-5)


def _net(meta_or_phi, nc, sd, dev):
    from glsdet_b200.yolox_ffa import YoloBody

    net = YoloBody(nc, meta_or_phi)
    net.load_state_dict(sd, strict=True)
    return net.to(dev).eval().set_precision("fp32")


@pytest.mark.parametrize("meta", META["models"], ids=[m["name"] for m in META["models"]])
def test_fp32_model_matches_reference_golden(meta, native_lib, cuda_device):
    z = np.load(GOLD / f"{meta['name']}.npz")
    sd = ref_path.synthetic_state_dict(meta["nc"], meta["phi"], seed=meta["seed"], flavour=meta["flavour"])
    net = _net(meta["phi"], meta["nc"], sd, cuda_device)
    feats = [torch.from_numpy(z[f"feat{i}"]).to(cuda_device) for i in range(4)]
    neck = net.backbone.forward_features(feats)
    for i in range(1, 4):
        _close(neck[i], torch.from_numpy(z[f"neck{i}"]), f"fp32 neck{i}")
    logits = net.forward_features(feats)
    for i in range(4):
        _close(logits[i], torch.from_numpy(z[f"logits{i}"]), f"fp32 logits{i}")
    pred = net.decode_features(feats)
    _close(pred, torch.from_numpy(z["pred"]), "fp32 decoded rows")


def test_config1_yolox_s_640_fp32(native_lib, cuda_device):
    """BASELINE.json configs[0]: P0 YoloBody(10, 's'), reference random init (weights_init normal 0.02), one synthetic
    640 x 640 image: neck + head logits, decode and detections against the fp32 oracle."""
    from glsdet_b200.utils_bbox import non_max_suppression

    nc = 10
    sd = ref_path.synthetic_state_dict(nc, "s", seed=0, flavour="reference")
    net = _net("s", nc, sd, cuda_device)
    x = torch.randn(1, 3, 640, 640, generator=torch.Generator().manual_seed(0))
    feats = ref_path.csp_darknet(sd, x)
    ref = ref_path.neck_head(sd, feats)
    dfeats = [f.to(cuda_device) for f in feats]
    out = net.forward_features(dfeats)
    assert [tuple(o.shape) for o in out] == [(1, 15, 160, 160), (1, 15, 80, 80), (1, 15, 40, 40), (1, 15, 20, 20)]
    worst = max(_close(out[i], ref[i], f"config 1 logits{i}") for i in range(4))
    assert worst <= 1e-4, f"fp32 path should sit at summation-order noise, got {worst:.3g}"
    pred = net.decode_features(dfeats)
    ref_pred = ref_path.decode_outputs(ref, [640, 640])
    _close(pred, ref_pred, "config 1 decoded rows")
    # the calibrated flavour (O(1) activations at every depth) on the same image
    sd2 = ref_path.synthetic_state_dict(nc, "s", seed=0, flavour="calibrated")
    net2 = _net("s", nc, sd2, cuda_device)
    feats2 = ref_path.csp_darknet(sd2, x)
    ref2 = ref_path.neck_head(sd2, feats2)
    out2 = net2.forward_features([f.to(cuda_device) for f in feats2])
    for i in range(4):
        _close(out2[i], ref2[i], f"config 1 (calibrated weights) logits{i}")
    # detections from identical predictions are bit-exact (same post-processing kernels as the bf16 path)
    res = non_max_suppression(ref_path.decode_outputs(ref2, [640, 640]).to(cuda_device), nc, [640, 640], np.array([640, 640]),
                              False, 0.01, 0.65, "auto_cpu")
    ref_res = ref_path.non_max_suppression(ref_path.decode_outputs(ref2, [640, 640]), nc, [640, 640], np.array([640, 640]),
                                           False, 0.01, 0.65, strategy="auto_cpu")
    np.testing.assert_array_equal(res[0], ref_res[0])


def test_batched_gemm_f32_kernel(native_lib, cuda_device):
    """glsdet_bgemm_f32 (the two products of the un-folded non-local block) on channel windows of NHWC fp32 tensors, against
    float64 torch: the Gram form (A stored [K][M], k = pixel) and the apply form; ragged sizes."""
    from glsdet_b200.ops import BGemmF32Op, View

    dev = cuda_device
    g = torch.Generator().manual_seed(11)
    for b, h, w, ci in ((2, 8, 12, 64), (3, 5, 7, 48), (1, 32, 32, 128), (2, 1, 3, 16)):
        t = torch.randn(b, h, w, 3 * ci + 8, generator=g).to(dev)
        a, bb = View(t, 8, ci), View(t, 8 + ci, ci)
        m = torch.full((b, ci, ci), float("nan"), device=dev)
        BGemmF32Op("gram", a, bb, m, alpha=1.0 / (h * w)).launch()
        A = t[..., 8:8 + ci].flatten(1, 2).double()
        Bm = t[..., 8 + ci:8 + 2 * ci].flatten(1, 2).double()
        ref = torch.einsum("bti,btj->bij", A, Bm) / (h * w)
        assert (m.double() - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item()), (b, h, w, ci)
        y = torch.full((b, h, w, ci + 8), 5.0, device=dev)
        BGemmF32Op("apply", a, m, View(y, 8, ci)).launch()
        torch.cuda.synchronize()
        ref_y = torch.einsum("bti,bij->btj", A, m.double()).view(b, h, w, ci)
        assert (y[..., 8:].double() - ref_y).abs().max().item() <= 1e-5 * max(1.0, ref_y.abs().max().item())
        assert (y[..., :8] == 5.0).all()


@pytest.mark.parametrize("variant", ["p1", "p2"])
def test_fp32_mode_of_p1_p2_matches_reference_golden(variant, native_lib, cuda_device):
    """fp32 accuracy mode of the GLSDet P1 (models/new/yolox10.py: patch non-local attention + cross-level head) and P2
    (models/block/non_local/yolo_patch_nonlocal_plus.py) models: neck outputs and logits within 1e-3 relative (BASELINE.json)
    of the REAL reference's outputs (tests/golden/p{1,2}_s_calibrated.npz); measured at summation-order noise."""
    import importlib

    meta = META[variant]
    z = np.load(GOLD / f"{meta['name']}.npz")
    sd = ref_path.synthetic_state_dict(meta["nc"], meta["phi"], seed=meta["seed"], flavour="calibrated", variant=variant)
    mod = importlib.import_module("glsdet_b200.yolox10" if variant == "p1" else "glsdet_b200.yolo_patch_nonlocal_plus")
    net = mod.YoloBody(meta["nc"], meta["phi"])
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda_device).eval().set_precision("fp32")
    nf = 4 if variant == "p1" else 3
    feats = [torch.from_numpy(z[f"feat{i}"]).to(cuda_device) for i in range(nf)]
    logits = net.forward_features(feats)
    worst = 0.0
    for i, t in enumerate(logits):
        worst = max(worst, _close(t, torch.from_numpy(z[f"logits{i}"]), f"fp32 {variant} logits{i}"))
    assert worst <= 1e-4, f"fp32 {variant} should sit at summation-order noise, got {worst:.3g}"
