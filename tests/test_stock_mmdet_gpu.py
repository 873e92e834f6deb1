"""GPU parity of the stock three-level YOLOX drop-in (models/base/yolox.py) and of the mmdet registry face
(YOLOXPAFPN + YOLOXHead incl. get_bboxes), through the C ABI, against goldens of the real reference and the oracle."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from _helpers import TOL, _clustered, assert_close_rel
from oracle import mmdet_ref, ref_path

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
META = json.loads((GOLD / "meta.json").read_text())["stock"]


def _stock_sd():
    return ref_path.synthetic_state_dict(META["nc"], META["phi"], seed=META["seed"], flavour="calibrated", variant="stock")


def test_stock_yolox_matches_reference_golden(native_lib, cuda_device):
    from glsdet_b200.utils_bbox import decode_outputs, non_max_suppression
    from glsdet_b200.yolox_base import YoloBody

    z = np.load(GOLD / "stock_s_calibrated.npz")
    sd = _stock_sd()
    net = YoloBody(META["nc"], META["phi"])
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda_device).eval()
    feats = [torch.from_numpy(z[f"feat{i}"]).to(cuda_device) for i in range(3)]
    emu = ref_path.stock_neck_head(sd, [f.cpu() for f in feats], bf16=True)
    neck = net.backbone.forward_features(feats)
    for i in range(3):
        assert_close_rel(neck[i], torch.from_numpy(z[f"neck{i}"]), 2.5e-2, f"stock neck{i}")
    logits = net.forward_features(feats)
    for i in range(3):
        ref_i = torch.from_numpy(z[f"logits{i}"])
        assert_close_rel(logits[i], ref_i, TOL, f"stock logits{i}", frac=2e-2)
        assert_close_rel(logits[i], emu[i], 1.5e-2, f"stock logits{i} vs bf16 emulation", frac=5e-2)
    hl = net.head([torch.from_numpy(z[f"neck{i}"]).to(cuda_device) for i in range(3)])
    for i in range(3):
        assert_close_rel(hl[i], torch.from_numpy(z[f"logits{i}"]), TOL, f"stock head-only logits{i}")
    pred_fused = net.decode_features(feats)
    assert torch.allclose(pred_fused, decode_outputs(logits, [META["in_h"], META["in_w"]]), rtol=1e-5, atol=1e-6)
    res = non_max_suppression(torch.from_numpy(z["pred"]).to(cuda_device), META["nc"], [META["in_h"], META["in_w"]],
                              np.array([META["in_h"], META["in_w"]]), False, META["conf"], META["nms_thr"], "auto_cpu")
    for b in range(META["batch"]):
        np.testing.assert_array_equal(res[b], z[f"nms{b}"])


def _match_dets(got, want, atol=2e-3):
    """Same detections up to ulp-level score/box differences (the GPU exp/sigmoid are not bit-identical to the CPU's)."""
    gd, gl = got[0].cpu().numpy(), got[1].cpu().numpy()
    wd, wl = want
    assert abs(len(gd) - len(wd)) <= max(1, len(wd) // 100), (len(gd), len(wd))
    used = np.zeros(len(gd), bool)
    missing = 0
    for d, l in zip(wd, wl):
        cand = np.nonzero((gl == l) & ~used)[0]
        if len(cand) == 0:
            missing += 1
            continue
        err = np.abs(gd[cand] - d).max(1)
        j = err.argmin()
        if err[j] <= atol * max(1.0, np.abs(d[:4]).max()):
            used[cand[j]] = True
        else:
            missing += 1
    assert missing <= max(1, len(wd) // 100), f"{missing} of {len(wd)} oracle detections have no counterpart"


def test_mmdet_face_matches_oracle(native_lib, cuda_device):
    from glsdet_b200.mmdet_face import HEADS, NECKS

    z = np.load(GOLD / "stock_s_calibrated.npz")
    nsd, hsd = mmdet_ref.drone_to_mmdet_keys(_stock_sd())
    test_cfg = dict(score_thr=META["conf"], nms=dict(type="nms", iou_threshold=META["nms_thr"]))
    neck = NECKS.build(dict(type="YOLOXPAFPN", in_channels=[128, 256, 512], out_channels=128, num_csp_blocks=1))
    head = HEADS.build(dict(type="YOLOXHead", num_classes=META["nc"], in_channels=128, feat_channels=128, test_cfg=test_cfg))
    neck.load_state_dict(nsd, strict=True)
    head.load_state_dict(hsd, strict=True)
    neck, head = neck.to(cuda_device).eval(), head.to(cuda_device).eval()
    feats = [torch.from_numpy(z[f"feat{i}"]) for i in range(3)]
    ref_p = mmdet_ref.yolox_pafpn(nsd, feats)
    got_p = neck(tuple(f.to(cuda_device) for f in feats))
    assert isinstance(got_p, tuple) and len(got_p) == 3
    for i in range(3):
        assert_close_rel(got_p[i], ref_p[i], 2.5e-2, f"mmdet neck out{i}")
    ref_cls, ref_box, ref_obj = mmdet_ref.yolox_head_forward(hsd, ref_p)
    cls, box, obj = head([p.to(cuda_device) for p in ref_p])
    for i in range(3):
        assert cls[i].shape == ref_cls[i].shape and box[i].shape == ref_box[i].shape and obj[i].shape == ref_obj[i].shape
        assert_close_rel(torch.cat([box[i], obj[i], cls[i]], 1), torch.cat([ref_box[i], ref_obj[i], ref_cls[i]], 1),
                         TOL, f"mmdet head level {i}")
    # get_bboxes on the oracle's raw maps (identical inputs on both sides), without and with rescale
    dev = lambda ts: [t.to(cuda_device) for t in ts]
    want = mmdet_ref.get_bboxes(ref_cls, ref_box, ref_obj, [8, 16, 32], META["conf"], META["nms_thr"])
    got = head.get_bboxes(dev(ref_cls), dev(ref_box), dev(ref_obj), img_metas=[{} for _ in range(META["batch"])])
    for b in range(META["batch"]):
        assert got[b][0].shape[1] == 5 and got[b][1].dtype == torch.int64
        _match_dets(got[b], want[b])
    sf = [[1.5, 1.25, 1.5, 1.25], [0.5, 0.75, 0.5, 0.75]]
    metas = [dict(scale_factor=np.array(s, np.float32)) for s in sf]
    want = mmdet_ref.get_bboxes(ref_cls, ref_box, ref_obj, [8, 16, 32], META["conf"], META["nms_thr"], scale_factors=sf)
    got = head.get_bboxes(dev(ref_cls), dev(ref_box), dev(ref_obj), img_metas=metas, rescale=True)
    for b in range(META["batch"]):
        _match_dets(got[b], want[b])
    # also accepts the views its own forward returns (channel slices, no copy)
    own = head.get_bboxes(cls, box, obj, img_metas=[{} for _ in range(META["batch"])])
    assert len(own) == META["batch"] and own[0][0].shape[1] == 5


def test_mmcv_strategy_large_k(native_lib, cuda_device):
    """mmcv batched_nms switches to per-class NMS on the shifted boxes from 10 000 boxes on."""
    from glsdet_b200.utils_bbox import batched_nms

    for k in (9000, 12000):
        rng = np.random.default_rng(k)
        boxes, scores, labels = _clustered(rng, k, 6)
        boxes *= 640.0
        boxes[::7] -= 40.0      # some boxes stick out of the image: classes are no longer separable by the shift
        got = batched_nms(*(torch.from_numpy(a).to(cuda_device) for a in (boxes, scores, labels)), 0.6, "mmcv").cpu().numpy()
        _, keep = mmdet_ref.mmcv_batched_nms(boxes, scores, labels.astype(np.int64), 0.6)
        np.testing.assert_array_equal(got, keep)


def test_mmdet_face_vs_reference_source_golden(native_lib, cuda_device):
    """The registry classes against outputs recorded by executing the reference's own YOLOXPAFPN / CSPLayer / YOLOXHead
    sources (tests/golden/make_golden_mmdet_yolox.py): neck and head maps within the 16-bit tolerance, get_bboxes fed the
    golden maps equal to the golden detections (labels and order identical), with and without rescale."""
    from glsdet_b200.mmdet_face import HEADS, NECKS

    z = np.load(GOLD / "mmdet_yolox_cases.npz")
    H, W, seed = (int(v) for v in z["meta"])
    nsd, hsd = mmdet_ref.drone_to_mmdet_keys(_stock_sd())
    test_cfg = dict(score_thr=0.01, nms=dict(type="nms", iou_threshold=0.65))
    neck = NECKS.build(dict(type="YOLOXPAFPN", in_channels=[128, 256, 512], out_channels=128, num_csp_blocks=1))
    head = HEADS.build(dict(type="YOLOXHead", num_classes=META["nc"], in_channels=128, feat_channels=128, test_cfg=test_cfg))
    neck.load_state_dict(nsd, strict=True)
    head.load_state_dict(hsd, strict=True)
    neck, head = neck.to(cuda_device).eval(), head.to(cuda_device).eval()
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(2, c, H // s, W // s, generator=g) for c, s in zip((128, 256, 512), (8, 16, 32))]
    got_p = neck(tuple(f.to(cuda_device) for f in feats))
    for l in range(3):
        assert_close_rel(got_p[l], torch.from_numpy(z[f"neck{l}"]), 2.5e-2, f"neck out{l} vs golden")
    cls, box, obj = head([torch.from_numpy(z[f"neck{l}"]).to(cuda_device) for l in range(3)])
    for l in range(3):
        assert_close_rel(torch.cat([box[l], obj[l], cls[l]], 1),
                         torch.from_numpy(np.concatenate([z[f"box{l}"], z[f"obj{l}"], z[f"cls{l}"]], 1)), TOL, f"head level {l} vs golden")
    maps = [[torch.from_numpy(z[f"{n}{l}"]).to(cuda_device) for l in range(3)] for n in ("cls", "box", "obj")]
    sf = [[1.0, 1.0, 1.0, 1.0], [1.25, 1.5, 1.25, 1.5]]
    metas = [dict(scale_factor=np.array(s, np.float32)) for s in sf]
    for tag, rescale in (("plain", False), ("scaled", True)):
        got = head.get_bboxes(*maps, img_metas=metas, rescale=rescale)
        for i in range(2):
            _match_dets(got[i], (z[f"{tag}_dets{i}"], z[f"{tag}_labels{i}"]))
