"""GPU parity of the whole hot path (through the C ABI) against the oracle and the reference's golden vectors.

Tolerances (BASELINE.json north_star: "feature maps and logits within ... 2e-2 relative in bf16", "NMS
keep-indices bit-exact when fed identical scores"):
  * against the fp32 oracle (quantisation + kernel error): ||diff||_2 <= 2e-2 * ||ref||_2 for every neck map and every
    logit level - the plain bound, no allowance relative to any "inherent" error - with 99 % of the elements within
    2e-2 * max|ref| and every element within 8e-2 * max|ref|.  Measured (profiles/r2_parity_report.txt): 0.03 - 0.25 %
    with the default storage policy (bf16 for the stride-4 head level, fp16 for everything else);
  * against the storage-precision emulation of the oracle (oracle.ref_path.neck_head_bf16: fp32 math, weights and
    activations rounded to the storage type of their level; kernel error only): ||diff||_2 <= 1.5e-2 * ||ref||_2;
  * decode within 1e-5; NMS rows bit-exact when fed identical predictions.
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _helpers import TOL, _clustered, assert_close_rel
from oracle import nms_oracle, ref_path

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
META = json.loads((GOLD / "meta.json").read_text())


# ---------------------------------------------------------------------------------------------- small kernels
def test_layout_converters_roundtrip(native_lib, cuda_device):
    from glsdet_b200.ops import View, nchw_to_nhwc, nhwc_to_nchw

    g = torch.Generator().manual_seed(0)
    for (b, c, h, w) in [(2, 64, 16, 24), (1, 48, 7, 9), (3, 200, 5, 33)]:
        x = torch.randn(b, c, h, w, generator=g).to(cuda_device)
        buf = torch.zeros(b, h, w, c + 16, device=cuda_device, dtype=torch.bfloat16)
        nchw_to_nhwc(x, View(buf, 8, c))
        ref = x.to(torch.bfloat16)
        assert torch.equal(buf[..., 8:8 + c].permute(0, 3, 1, 2), ref)
        assert buf[..., :8].abs().max() == 0 and buf[..., 8 + c:].abs().max() == 0
        back = torch.empty(b, c, h, w, device=cuda_device)
        nhwc_to_nchw(View(buf, 8, c), back)
        assert torch.equal(back, ref.float())


def test_se_gate_and_pixel_shuffle(native_lib, cuda_device):
    from glsdet_b200.ops import ScaleShuffleOp, SeGateOp, View

    g = torch.Generator().manual_seed(1)
    B, H, W, C = 2, 12, 20, 64     # 4C = 256 channels before the shuffle
    x = torch.randn(B, 4 * C, H, W, generator=g).abs().to(cuda_device)
    w1 = (torch.randn(16, 4 * C, generator=g) / 16).to(cuda_device)
    w2 = (torch.randn(4 * C, 16, generator=g) / 4).to(cuda_device)
    xb = x.to(torch.bfloat16).float()
    y = torch.sigmoid(F.linear(torch.relu(F.linear(xb.mean((2, 3)), w1)), w2))
    ref = F.pixel_shuffle(xb + xb * y.view(B, 4 * C, 1, 1), 2)          # ffa.py:77-78
    perm = torch.arange(4 * C, device=cuda_device).view(C, 4).t().reshape(-1)
    xin = xb[:, perm].permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)  # channels in (i, j, c) order
    se = SeGateOp(View(xin), w1[:, perm], w2[perm])
    se.launch()
    dst = torch.zeros(B, 2 * H, 2 * W, 2 * C, device=cuda_device, dtype=torch.bfloat16)
    ScaleShuffleOp(View(xin), se.gate, View(dst, C, C)).launch()
    torch.cuda.synchronize()
    assert torch.allclose(se.gate, (1 + y)[:, perm], rtol=1e-5, atol=1e-6)
    got = dst[..., C:].float().permute(0, 3, 1, 2)
    assert_close_rel(got, ref, 1e-2, "scale+shuffle")
    assert dst[..., :C].abs().max() == 0


def test_decode_outputs_matches_reference_golden(native_lib, cuda_device):
    from glsdet_b200.utils_bbox import decode_outputs

    m = META["postproc"]
    z = np.load(GOLD / "postproc_reference.npz")
    logits = [torch.from_numpy(z[f"logits{i}"]).to(cuda_device) for i in range(4)]
    keep = [t.clone() for t in logits]
    pred = decode_outputs(logits, [m["in_h"], m["in_w"]]).cpu()
    ref = torch.from_numpy(z["pred"]).clone()
    pred[2, :, 4] = 0.0
    assert torch.allclose(pred, ref, rtol=1e-5, atol=1e-6)
    for a, b in zip(logits, keep):
        assert torch.equal(a, b), "decode_outputs must not modify its inputs (utils_bbox.py:266 works on a copy)"


# ---------------------------------------------------------------------------------------------- NMS
@pytest.mark.parametrize("case", META["nms"]["cases"], ids=[c["name"] for c in META["nms"]["cases"]])
def test_batched_nms_bit_exact_vs_torchvision_golden(case, native_lib, cuda_device):
    from glsdet_b200.utils_bbox import batched_nms

    z = np.load(GOLD / "nms_torchvision.npz")
    n = case["name"]
    b, s, l = (torch.from_numpy(z[f"{n}_{k}"]).to(cuda_device) for k in ("boxes", "scores", "labels"))
    for thr in (0.45, 0.65):
        got = batched_nms(b, s, l, thr, "trick").cpu().numpy()
        np.testing.assert_array_equal(got, z[f"{n}_keep_trick_{thr}"])
        got = batched_nms(b, s, l, thr, "per_class").cpu().numpy()
        ref = z[f"{n}_keep_vanilla_{thr}"]
        if case["ties"]:   # the reference's final sort is unstable for tied scores; ours breaks ties by index
            np.testing.assert_array_equal(got, nms_oracle.batched_nms(z[f"{n}_boxes"], z[f"{n}_scores"],
                                                                      z[f"{n}_labels"], thr, "per_class"))
            assert sorted(got.tolist()) == sorted(ref.tolist())
        else:
            np.testing.assert_array_equal(got, ref)
        got = batched_nms(b, s, l, thr, "auto_cpu").cpu().numpy()
        if not case["ties"]:
            np.testing.assert_array_equal(got, z[f"{n}_keep_auto_{thr}"])


@pytest.mark.parametrize("k,nc", [(2, 1), (63, 3), (64, 3), (65, 10), (4095, 10), (4096, 10), (4097, 10), (9000, 10),
                                  (20000, 4), (40000, 80)])
def test_batched_nms_random_vs_oracle(k, nc, native_lib, cuda_device):
    from glsdet_b200.utils_bbox import batched_nms

    rng = np.random.default_rng(k * 131 + nc)
    boxes, scores, labels = _clustered(rng, k, nc)
    tb, ts, tl = (torch.from_numpy(a).to(cuda_device) for a in (boxes, scores, labels))
    for strat in ("trick", "per_class"):
        got = batched_nms(tb, ts, tl, 0.6, strat).cpu().numpy()
        ref = nms_oracle.batched_nms(boxes, scores, labels, 0.6, strat)
        np.testing.assert_array_equal(got, ref)
        assert 0 < len(got) < k or k < 64


def test_batched_nms_edge_cases(native_lib, cuda_device):
    from glsdet_b200.utils_bbox import batched_nms

    dev = cuda_device
    e = batched_nms(torch.zeros(0, 4, device=dev), torch.zeros(0, device=dev), torch.zeros(0, device=dev), 0.5)
    assert e.shape == (0,) and e.dtype == torch.int64
    # identical boxes, tied scores: the first index wins (stable sort), everything else is suppressed
    b = torch.tensor([[0.1, 0.1, 0.5, 0.5]] * 5, device=dev)
    s = torch.full((5,), 0.7, device=dev)
    l = torch.zeros(5, device=dev)
    assert batched_nms(b, s, l, 0.5, "trick").tolist() == [0]
    # different classes never suppress each other
    l2 = torch.arange(5, device=dev, dtype=torch.float32)
    assert batched_nms(b, s, l2, 0.5, "trick").tolist() == [0, 1, 2, 3, 4]
    assert batched_nms(b, s, l2, 0.5, "per_class").tolist() == [0, 1, 2, 3, 4]
    # zero-area boxes: IoU is 0/0 = NaN and never suppresses (torchvision semantics)
    z = torch.tensor([[0.3, 0.3, 0.3, 0.3]] * 3, device=dev)
    assert batched_nms(z, torch.tensor([0.9, 0.8, 0.7], device=dev), torch.zeros(3, device=dev), 0.5).tolist() == [0, 1, 2]
    # boxes far outside the image (min coordinate < -0.5) force the literal class-agnostic coordinate trick
    rng = np.random.default_rng(3)
    boxes, scores, labels = _clustered(rng, 700, 6)
    boxes -= 2.0
    got = batched_nms(*(torch.from_numpy(a).to(dev) for a in (boxes, scores, labels)), 0.5, "trick").cpu().numpy()
    np.testing.assert_array_equal(got, nms_oracle.batched_nms(boxes, scores, labels, 0.5, "trick"))


def test_batched_nms_negative_scores_and_arbitrary_idxs(native_lib, cuda_device):
    """torchvision's contract: scores may be raw logits (negative) and idxs any values (batch * nc + cls, negative,
    non-integral); checked against the oracle and against torchvision itself on the CPU."""
    import torchvision

    from glsdet_b200.utils_bbox import batched_nms

    rng = np.random.default_rng(77)
    boxes, scores, labels = _clustered(rng, 3000, 12)
    logit = np.log(scores / (1 - scores)).astype(np.float32)          # negative for half of the boxes
    for lab in (labels * 37 + 300, -labels - 1, labels * 0.25, labels * 1e6):
        lab = lab.astype(np.float32)
        tb, ts, tl = (torch.from_numpy(a).to(cuda_device) for a in (boxes, logit, lab))
        for strat in ("trick", "per_class"):
            got = batched_nms(tb, ts, tl, 0.55, strat).cpu().numpy()
            np.testing.assert_array_equal(got, nms_oracle.batched_nms(boxes, logit, lab, 0.55, strat))
        # the real torchvision op (CPU dispatch: 4 * 3000 > 4000 elements -> per-class branch)
        tv = torchvision.ops.boxes.batched_nms(torch.from_numpy(boxes), torch.from_numpy(logit), torch.from_numpy(lab), 0.55)
        np.testing.assert_array_equal(batched_nms(tb, ts, tl, 0.55, "auto_cpu").cpu().numpy(), tv.numpy())
    with pytest.raises(NotImplementedError):
        batched_nms(tb[:600], ts[:600], torch.arange(600, device=cuda_device, dtype=torch.float32), 0.5)


def test_filter_argmax_of_near_tied_saturated_logits(native_lib, cuda_device):
    """Fused detect path: class logits around 12-15 saturate to (almost) the same fp32 probability; torch.max returns the
    FIRST index among equal probabilities, so the filter must evaluate every candidate class, not only the largest logit."""
    from glsdet_b200.utils_bbox import DeviceNMS

    B, A, nc = 1, 4096, 10
    g = torch.Generator().manual_seed(5)
    logits = torch.full((B, A, nc), -6.0)
    m = 9.0 + 8.0 * torch.rand(A, generator=g)                  # largest logit 9 .. 17
    hi = torch.randint(1, nc, (A,), generator=g)                  # its class (never class 0)
    lo = (hi - 1 - torch.randint(0, 3, (A,), generator=g)).clamp(min=0)   # a lower-index class slightly below it
    gap = 0.2 * torch.rand(A, generator=g) ** 3
    idx = torch.arange(A)
    logits[0, idx, hi] = m
    logits[0, idx, lo] = m - gap
    pred = torch.zeros((B, A, 5 + nc))
    cx = ((idx % 64).float() + 0.5) / 64
    cy = ((idx // 64).float() + 0.5) / 64
    pred[0, :, 0], pred[0, :, 1], pred[0, :, 2], pred[0, :, 3], pred[0, :, 4] = cx, cy, 0.004, 0.004, 0.9
    pred[0, :, 5:] = logits[0]
    dpred = pred.to(cuda_device)
    dprob = dpred.clone()
    dprob[0, :, 5:] = 1.0 / (1.0 + torch.exp(-dpred[0, :, 5:]))    # the prediction epilogue's formula, fp32 on the device
    conf, label = dprob[0, :, 5:].max(1)                           # torch.max: first index among equal probabilities
    assert (label.cpu() != hi).float().mean() > 0.05, "the case must contain ties that resolve to the lower index"
    op = DeviceNMS(B, A, nc)
    det, cnt = op.launch(dpred, 0.01, 0.65, "per_class", cls_logits=True)
    torch.cuda.synchronize()
    assert int(cnt[0]) == A
    anchor = op.keep_index[0, :A].long()
    assert torch.equal(det[0, :A, 6], label[anchor].float())
    assert torch.equal(det[0, :A, 5], conf[anchor])
    ref_det = det.clone()
    det2, cnt2 = op.launch(dprob, 0.01, 0.65, "per_class")         # same probabilities, decoded: identical rows
    torch.cuda.synchronize()
    assert int(cnt2[0]) == A and torch.equal(det2, ref_det)


def test_nms_fallback_kernel_without_mask_budget(native_lib, cuda_device, monkeypatch):
    """With no bitmask budget every segment runs on the blocked-greedy fallback kernel: same keep sets."""
    from glsdet_b200.utils_bbox import batched_nms

    monkeypatch.setenv("GLSDET_NMS_MASK_WORDS", "0")
    for k, nc in ((65, 3), (3000, 2), (9000, 10)):
        rng = np.random.default_rng(k)
        boxes, scores, labels = _clustered(rng, k, nc)
        tb, ts, tl = (torch.from_numpy(a).to(cuda_device) for a in (boxes, scores, labels))
        for strat in ("trick", "per_class"):
            got = batched_nms(tb, ts, tl, 0.6, strat).cpu().numpy()
            np.testing.assert_array_equal(got, nms_oracle.batched_nms(boxes, scores, labels, 0.6, strat))
    monkeypatch.setenv("GLSDET_NMS_MASK_WORDS", "3000")   # mixed: small segments masked, large ones fall back
    rng = np.random.default_rng(1)
    boxes, scores, labels = _clustered(rng, 6000, 10)
    labels[:3000] = 0
    tb, ts, tl = (torch.from_numpy(a).to(cuda_device) for a in (boxes, scores, labels))
    got = batched_nms(tb, ts, tl, 0.6, "per_class").cpu().numpy()
    np.testing.assert_array_equal(got, nms_oracle.batched_nms(boxes, scores, labels, 0.6, "per_class"))


def test_non_max_suppression_matches_reference_golden(native_lib, cuda_device):
    """Reference decode + NMS outputs recorded on the CPU (torchvision auto dispatch for CPU tensors)."""
    from glsdet_b200.utils_bbox import non_max_suppression

    m = META["postproc"]
    z = np.load(GOLD / "postproc_reference.npz")
    pred = torch.from_numpy(z["pred"]).to(cuda_device)
    before = pred.clone()
    res = non_max_suppression(pred, m["nc"], [m["in_h"], m["in_w"]], np.array(m["image_shape"]), m["letterbox"],
                              conf_thres=m["conf"], nms_thres=m["nms_thr"], strategy="auto_cpu")
    assert torch.equal(pred, before)
    assert [len(r) for r in res] == m["kept"] and res[2].shape == (0, 7) and res[0].dtype == np.float32
    for b in range(m["batch"]):
        np.testing.assert_array_equal(res[b], z[f"nms{b}"])
    assert non_max_suppression(torch.zeros(2, 0, 15, device=cuda_device), 10, [32, 32], np.array([32, 32]), False) == [None, None]


def test_device_nms_full_anchor_count_properties(native_lib, cuda_device):
    """BASELINE config 2 anchor count (87 040 per image): size-independent properties + oracle on one image."""
    from glsdet_b200.utils_bbox import DeviceNMS

    B, A, nc = 4, 87040, 10
    g = torch.Generator().manual_seed(5)
    pred = torch.rand(B, A, 5 + nc, generator=g)
    pred[..., 0:2] = torch.rand(B, A, 2, generator=g)
    pred[..., 2:4] = torch.rand(B, A, 2, generator=g) * 0.08 + 0.01
    pred[..., 4] = torch.rand(B, A, generator=g) ** 4
    pred[3, :, 4] = 0.0
    dpred = pred.to(cuda_device)
    op = DeviceNMS(B, A, nc)
    det, cnt = op.launch(dpred, 0.3, 0.5, "auto_cuda")
    torch.cuda.synchronize()
    cnt = cnt.cpu().numpy()
    assert cnt[3] == 0 and (cnt[:3] > 0).all()
    det_h = det.cpu().numpy()
    ref = ref_path.non_max_suppression(pred[:1], nc, [1024, 1024], None, False, 0.3, 0.5, strategy="auto_cuda",
                                       correct_boxes=False)
    np.testing.assert_array_equal(det_h[0, :cnt[0]], ref[0])
    for b in range(3):
        rows = det_h[b, :cnt[b]]
        sc = rows[:, 4] * rows[:, 5]
        assert (np.diff(sc) <= 0).all(), "rows must be sorted by score"
        assert (sc >= 0.3).all()
        # idempotence: NMS of the survivors keeps all of them
        again = nms_oracle.batched_nms(rows[:, :4], sc, rows[:, 6], 0.5, "auto_cuda")
        assert len(again) == len(rows)


def test_device_nms_topk_mode_matches_full_nms(native_lib, cuda_device, monkeypatch):
    """max_det <= 1536 switches to the bitmask-free blocked-greedy kernel with early exit per (image, class) segment:
    its rows must be exactly the first max_det rows of the unlimited NMS (mask path), for every strategy."""
    from glsdet_b200.utils_bbox import DeviceNMS

    B, A, nc = 3, 20000, 4
    rng = np.random.default_rng(11)
    pred = np.zeros((B, A, 5 + nc), np.float32)
    for b in range(B):
        boxes, scores, labels = _clustered(rng, A, nc, spread=0.01)
        labels[: A // 2] = 0                      # one dominant class: its segment is cut short by the early exit
        cxcy = (boxes[:, :2] + boxes[:, 2:]) / 2
        pred[b, :, 0:2], pred[b, :, 2:4] = cxcy, boxes[:, 2:] - boxes[:, :2]
        pred[b, :, 4] = scores
        pred[b, np.arange(A), 5 + labels.astype(int)] = 1.0
    pred[2, :, 4] *= 0.2                          # an image with few candidates (no early exit)
    dpred = torch.from_numpy(pred).to(cuda_device)
    full = DeviceNMS(B, A, nc)
    for strategy in ("auto_cuda", "per_class", "trick"):
        det_f, cnt_f = (t.clone() for t in full.launch(dpred, 0.15, 0.5, strategy))
        for max_det in (100, 1000, 1536):
            lim = DeviceNMS(B, A, nc, max_det=max_det)
            det_l, cnt_l = lim.launch(dpred, 0.15, 0.5, strategy)
            torch.cuda.synchronize()
            for b in range(B):
                k = min(int(cnt_f[b]), max_det)
                assert int(cnt_l[b]) == k, (strategy, max_det, b, int(cnt_l[b]), k)
                assert torch.equal(det_l[b, :k], det_f[b, :k])
                assert torch.equal(lim.keep_index[b, :k], full.keep_index[b, :k])
    assert int(cnt_f[0]) > 1536, "the case must really truncate"
    monkeypatch.setenv("GLSDET_NMS_NO_TOPK", "1")    # same limit on the bitmask path
    lim = DeviceNMS(B, A, nc, max_det=1000)
    det_l, cnt_l = lim.launch(dpred, 0.15, 0.5, "auto_cuda")
    det_f, cnt_f = full.launch(dpred, 0.15, 0.5, "auto_cuda")
    torch.cuda.synchronize()
    for b in range(B):
        k = min(int(cnt_f[b]), 1000)
        assert int(cnt_l[b]) == k and torch.equal(det_l[b, :k], det_f[b, :k])


# ---------------------------------------------------------------------------------------------- whole model
@pytest.mark.parametrize("meta", META["models"], ids=[m["name"] for m in META["models"]])
def test_model_matches_reference_golden(meta, native_lib, cuda_device):
    from glsdet_b200.utils_bbox import decode_outputs, non_max_suppression
    from glsdet_b200.yolox_ffa import YoloBody

    z = np.load(GOLD / f"{meta['name']}.npz")
    sd = ref_path.synthetic_state_dict(meta["nc"], meta["phi"], seed=meta["seed"], flavour=meta["flavour"])
    net = YoloBody(meta["nc"], meta["phi"])
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda_device).eval()
    feats = [torch.from_numpy(z[f"feat{i}"]).to(cuda_device) for i in range(4)]

    neck = net.backbone.forward_features(feats)
    for i in range(1, 4):
        ref_i = torch.from_numpy(z[f"neck{i}"])
        assert_close_rel(neck[i], ref_i, TOL, f"neck{i}")
    logits = net.forward_features(feats)
    emu = ref_path.neck_head_bf16(sd, [f.cpu() for f in feats])
    for i in range(4):
        ref_i = torch.from_numpy(z[f"logits{i}"])
        assert_close_rel(logits[i], ref_i, TOL, f"logits{i}", frac=2e-2)
        assert_close_rel(logits[i], emu[i], 1.5e-2, f"logits{i} vs bf16-storage emulation", max_factor=4.0, frac=5e-2)
    # stand-alone head module fed with the reference's own neck outputs
    hl = net.head([torch.from_numpy(z[f"neck{i}"]).to(cuda_device) for i in range(4)])
    for i in range(4):
        assert_close_rel(hl[i], torch.from_numpy(z[f"logits{i}"]), TOL, f"head-only logits{i}")  # shorter chain
    # fused decode == decode_outputs(raw logits) on our own logits
    pred_fused = net.decode_features(feats)
    pred_sep = decode_outputs(logits, [meta["in_h"], meta["in_w"]])
    assert torch.allclose(pred_fused, pred_sep, rtol=1e-5, atol=1e-6)
    # post-processing fed with the reference's own predictions: bit-exact rows
    res = non_max_suppression(torch.from_numpy(z["pred"]).to(cuda_device), meta["nc"], [meta["in_h"], meta["in_w"]],
                              np.array(meta["image_shape"]), meta["letterbox"], meta["conf"], meta["nms_thr"], "auto_cpu")
    for b in range(meta["batch"]):
        np.testing.assert_array_equal(res[b], z[f"nms{b}"])


def test_model_vs_oracle_1024_and_batch_invariance(native_lib, cuda_device):
    """BASELINE config 2 shape (P0-s, 1024x1024): one image against the oracle, then batch-of-4 == 4 x batch-of-1."""
    from glsdet_b200.yolox_ffa import YoloBody

    nc = 10
    sd = ref_path.synthetic_state_dict(nc, "s", seed=11, flavour="calibrated")
    net = YoloBody(nc, "s")
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda_device).eval()
    from glsdet_b200.synthetic import synthetic_images

    feats = ref_path.csp_darknet(sd, synthetic_images(4, 1024, 1024, seed=12))
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref = ref_path.neck_head(sd, [f[:1] for f in feats])
    emu = ref_path.neck_head_bf16(sd, [f[:1] for f in feats])
    dfeats = [f.to(cuda_device) for f in feats]
    out4 = net.forward_features(dfeats)
    for i in range(4):
        assert_close_rel(out4[i][:1], ref[i], TOL, f"1024 logits{i}")
        assert_close_rel(out4[i][:1], emu[i], 1.5e-2, f"1024 logits{i} vs bf16-storage emulation", frac=5e-2)
    pred4 = net.decode_features(dfeats).clone()
    ref_pred = ref_path.decode_outputs(ref, [1024, 1024])
    assert pred4.shape == (4, 87040, 15)
    dxy = (pred4[:1, :, :2].cpu() - ref_pred[..., :2]).abs()
    dpr = (pred4[:1, :, 4:].cpu() - ref_pred[..., 4:]).abs()
    assert dxy.max() <= 0.01, "box centres (normalised) far from the oracle"
    assert dpr.mean() <= 2e-3 and torch.quantile(dpr.flatten()[::7], 0.999) <= 0.1, "probabilities far from the oracle"
    for b in range(4):
        one = net.decode_features([f[b:b + 1].contiguous() for f in dfeats])
        assert torch.equal(one[0], pred4[b]), "results must not depend on the batch an image is in"
    det, cnt = net.detect_features(dfeats, conf_thres=0.02, nms_thres=0.65)
    torch.cuda.synchronize()
    assert (cnt.cpu() >= 0).all() and det.shape[0] == 4


def test_config4_yolox_l_544x1024(native_lib, cuda_device):
    """BASELINE config 4: P0 YOLOX-l, UAVDT-shaped 544x1024 input, nc=3, detections capped at 1000 per
    image.  Odd level sizes (17x32 at stride 32) exercise ragged tiles; one image is checked against the oracle."""
    from glsdet_b200.synthetic import synthetic_images
    from glsdet_b200.yolox_ffa import YoloBody

    nc, in_h, in_w = 3, 544, 1024
    sd = ref_path.synthetic_state_dict(nc, "l", seed=6, flavour="calibrated")   # trained-like conditioning, see calibrate_synthetic.py --bn-beta
    net = YoloBody(nc, "l")
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda_device).eval()
    feats = ref_path.csp_darknet(sd, synthetic_images(2, in_h, in_w, seed=5))
    assert [tuple(f.shape[1:]) for f in feats] == [(128, 136, 256), (256, 68, 128), (512, 34, 64), (1024, 17, 32)]
    ref = ref_path.neck_head(sd, [f[:1] for f in feats])
    emu = ref_path.neck_head_bf16(sd, [f[:1] for f in feats])
    dfeats = [f.to(cuda_device) for f in feats]
    logits = net.forward_features(dfeats)
    for i in range(4):
        assert_close_rel(logits[i][:1], ref[i], TOL, f"l-544 logits{i}", frac=2e-2)
    det, cnt = net.detect_features(dfeats, conf_thres=0.01, nms_thres=0.65, max_det=1000)
    torch.cuda.synchronize()
    assert det.shape == (2, 1000, 7) and (cnt.cpu() <= 1000).all() and (cnt.cpu() > 0).all()
    # top-1000 by score of the uncapped result (mmdet max_per_img semantics, base_dense_head.py:297)
    full, full_cnt = net.detect_features(dfeats, conf_thres=0.01, nms_thres=0.65, max_det=None)
    torch.cuda.synchronize()
    for b in range(2):
        n = int(cnt[b])
        assert n == min(1000, int(full_cnt[b]))
        assert torch.equal(det[b, :n], full[b, :n])


def test_detect_mode_raw_class_logits_equals_decoded_mode(native_lib, cuda_device):
    """The fused detect path leaves the class logits raw and lets the score filter apply the sigmoid
    (GLSDET_PRED_CLS_LOGITS); its rows must be bit-identical to filter + NMS on the decoded probabilities."""
    from glsdet_b200.synthetic import synthetic_state_dict
    from glsdet_b200.yolox_ffa import YoloBody

    sd = synthetic_state_dict(10, "s", seed=0, flavour="calibrated")
    net = YoloBody(10, "s")
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda_device).eval()
    g = torch.Generator().manual_seed(31)
    feats = [torch.randn(2, c, 256 // s, 320 // s, generator=g).to(cuda_device) for c, s in ((64, 4), (128, 8), (256, 16), (512, 32))]
    plan = net.plan_for(feats)
    assert plan.det_cls_logits and len(plan.pred_det_ops) == len(plan.pred_dec_ops)
    prob = plan.forward_decoded(feats).clone()
    assert tuple(prob.stride()) == (15 * plan.num_anchors, 1, plan.num_anchors)      # the reference's permuted view
    nms = net.nms_for(plan, None)
    det_a, cnt_a = nms.launch(prob, 0.01, 0.65)
    det_a, cnt_a = det_a.clone(), cnt_a.clone()
    raw = plan.forward_detect(feats)
    assert torch.equal(raw[..., :5], prob[..., :5])
    assert (raw[..., 5:] < 0).any() and float(prob[..., 5:].min()) >= 0.0            # logits vs probabilities
    det_b, cnt_b = net.detect_features(feats, conf_thres=0.01, nms_thres=0.65)
    assert int(cnt_a.min()) > 0 and torch.equal(cnt_a, cnt_b)
    for b in range(2):
        assert torch.equal(det_a[b, :int(cnt_a[b])], det_b[b, :int(cnt_b[b])])
    # rows layout (contiguous) gives the same result as the planes view
    det_c, cnt_c = nms.launch(prob.contiguous(), 0.01, 0.65)
    assert torch.equal(cnt_a, cnt_c) and torch.equal(det_a[0, :int(cnt_a[0])], det_c[0, :int(cnt_c[0])])


def test_filter_sigmoid_window_exact_on_ties_and_saturation(native_lib, cuda_device):
    """GLSDET_PRED_CLS_LOGITS evaluates the sigmoid only for the classes that can attain the maximal probability; the
    chosen (class_conf, class_pred) must equal the first maximum over the probabilities of ALL classes, also for equal
    and nearly equal logits and in the saturated ranges."""
    from glsdet_b200.utils_bbox import DeviceNMS

    g = torch.Generator().manual_seed(77)
    A, nc = 4096, 10
    logits = torch.randn(1, A, nc, generator=g) * 2.0 - 2.0
    logits[0, :200] = torch.randn(200, 1, generator=g).expand(200, nc)                    # all classes equal
    base = torch.randn(200, generator=g)
    logits[0, 200:400, 3] = base
    logits[0, 200:400, 7] = torch.nextafter(base, torch.full_like(base, 10.0))           # one ulp above an earlier class
    logits[0, 400:600] = 16.0 + 20.0 * torch.rand(200, nc, generator=g)                   # saturated at 1.0
    logits[0, 600:800] = -85.0 - 30.0 * torch.rand(200, nc, generator=g)                  # saturated at 0
    logits[0, 800:1000, 2] = 15.5
    logits[0, 800:1000, 5] = 16.5
    boxes = torch.rand(1, A, 2, generator=g)
    pred = torch.cat([boxes, 0.01 + 0.02 * torch.rand(1, A, 2, generator=g), torch.full((1, A, 1), 0.9), logits], 2)
    raw = pred.to(cuda_device).contiguous()
    prob = raw.clone()
    prob[..., 5:] = 1.0 / (1.0 + torch.exp(-raw[..., 5:]))       # same formula as the kernels (expf, IEEE division)
    nms = DeviceNMS(1, A, nc, device=cuda_device)
    det_a, cnt_a = nms.launch(prob, 0.0, 0.99)
    det_a, cnt_a = det_a.clone(), cnt_a.clone()
    det_b, cnt_b = nms.launch(raw, 0.0, 0.99, cls_logits=True)
    assert int(cnt_a[0]) > 3000 and torch.equal(cnt_a, cnt_b)
    n = int(cnt_a[0])
    assert torch.equal(det_a[0, :n], det_b[0, :n])


def test_batched_nms_ids_batch_equals_per_image_calls(native_lib, cuda_device):
    """glsdet_batched_nms_ids_batch (B independent problems in one launch sequence) == B calls of batched_nms, for the
    mmcv and the torchvision strategies; images of different coordinate ranges (per-image coordinate-trick maxima)."""
    import ctypes as C

    from glsdet_b200 import _native as N
    from glsdet_b200.utils_bbox import STRATEGIES, batched_nms

    rng = np.random.default_rng(17)
    B, k, nc = 3, 1500, 7
    boxes, scores, labels = [], [], []
    for b in range(B):
        bx, sc, lb = _clustered(rng, k, nc)
        boxes.append(bx * (100.0 * (b + 1)))          # different extents per image
        scores.append(sc)
        labels.append(lb)
    bx = torch.from_numpy(np.stack(boxes)).to(cuda_device).contiguous()
    sc = torch.from_numpy(np.stack(scores)).to(cuda_device).contiguous()
    lb = torch.from_numpy(np.stack(labels)).to(cuda_device).contiguous()
    ids = lb.to(torch.int32)
    for strategy in ("mmcv", "auto_cuda", "per_class"):
        nbytes = int(native_lib.glsdet_batched_nms_batch_workspace_bytes(B, k, nc))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=cuda_device)
        keep = torch.empty((B, k), dtype=torch.int32, device=cuda_device)
        cnt = torch.zeros((B,), dtype=torch.int32, device=cuda_device)
        N.check(native_lib.glsdet_batched_nms_ids_batch(bx.data_ptr(), sc.data_ptr(), lb.data_ptr(), ids.data_ptr(), float(nc - 1), nc, B, k,
                                                        0.6, STRATEGIES[strategy], ws.data_ptr(), nbytes, keep.data_ptr(), cnt.data_ptr(),
                                                        N.stream_ptr()), "glsdet_batched_nms_ids_batch")
        for b in range(B):
            want = batched_nms(bx[b], sc[b], lb[b], 0.6, strategy)
            assert torch.equal(keep[b, :int(cnt[b])].long(), want), (strategy, b)
