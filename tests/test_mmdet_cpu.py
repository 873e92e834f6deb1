"""CPU tests of the mmdet registry face: oracle restatement vs the pinned stock-YOLOX golden, key contract, registry."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import mmdet_ref, ref_path

GOLD = Path(__file__).resolve().parent / "golden"
META = json.loads((GOLD / "meta.json").read_text())["stock"]


def _stock_sd():
    return ref_path.synthetic_state_dict(META["nc"], META["phi"], seed=META["seed"], flavour="calibrated", variant="stock")


def test_stock_oracle_matches_reference_golden():
    z = np.load(GOLD / "stock_s_calibrated.npz")
    sd = _stock_sd()
    feats = [torch.from_numpy(z[f"feat{i}"]) for i in range(3)]
    out = ref_path.stock_neck_head(sd, feats)
    for i in range(3):
        np.testing.assert_allclose(out[i].numpy(), z[f"logits{i}"], rtol=1e-4, atol=2e-5)
    pred = ref_path.decode_outputs([torch.from_numpy(z[f"logits{i}"]) for i in range(3)], [META["in_h"], META["in_w"]])
    np.testing.assert_allclose(pred.numpy(), z["pred"], rtol=1e-6, atol=1e-7)


def test_mmdet_oracle_reproduces_stock_reference_outputs():
    """SURVEY 8c cross-check: yolox-drone base/yolox.py weights renamed onto the restated mmdet modules give the
    outputs of the real reference (golden)."""
    z = np.load(GOLD / "stock_s_calibrated.npz")
    neck_sd, head_sd = mmdet_ref.drone_to_mmdet_keys(_stock_sd())
    feats = [torch.from_numpy(z[f"feat{i}"]) for i in range(3)]
    p = mmdet_ref.yolox_pafpn(neck_sd, feats)
    cls, box, obj = mmdet_ref.yolox_head_forward(head_sd, p)
    for i in range(3):
        got = torch.cat([box[i], obj[i], cls[i]], 1).numpy()
        np.testing.assert_allclose(got, z[f"logits{i}"], rtol=1e-4, atol=2e-5)


def test_mmdet_get_bboxes_oracle_consistent_with_drone_postproc():
    """The mmdet decode (pixels) equals the yolox-drone decode (normalised) times the input size, and its NMS keeps
    the same anchors when both use the coordinate trick."""
    z = np.load(GOLD / "stock_s_calibrated.npz")
    logits = [torch.from_numpy(z[f"logits{i}"]) for i in range(3)]
    cls, box, obj = [l[:, 5:] for l in logits], [l[:, :4] for l in logits], [l[:, 4:5] for l in logits]
    res = mmdet_ref.get_bboxes(cls, box, obj, [8, 16, 32], META["conf"], META["nms_thr"])
    pred = torch.from_numpy(z["pred"])
    for b in range(META["batch"]):
        dets, labels = res[b]
        assert dets.shape[1] == 5 and len(dets) == len(labels) > 0
        assert (np.diff(dets[:, 4]) <= 0).all()
        drone = ref_path.non_max_suppression(pred[b:b + 1], META["nc"], [META["in_h"], META["in_w"]], None, False,
                                             META["conf"], META["nms_thr"], strategy="trick", correct_boxes=False)[0]
        # same number of survivors up to boxes whose IoU sits within rounding of the threshold
        assert abs(len(drone) - len(dets)) <= max(2, len(dets) // 50)
        k = min(len(drone), len(dets), 5)
        scale = np.array([META["in_w"], META["in_h"], META["in_w"], META["in_h"]], np.float32)
        np.testing.assert_allclose(dets[:k, :4], drone[:k, :4] * scale, rtol=1e-4, atol=1e-3)


def test_registry_and_state_dict_keys():
    from glsdet_b200.mmdet_face import HEADS, NECKS, YOLOXHead, YOLOXPAFPN

    neck = NECKS.build(dict(type="YOLOXPAFPN", in_channels=[128, 256, 512], out_channels=128, num_csp_blocks=1))
    head = HEADS.build(dict(type="YOLOXHead", num_classes=80, in_channels=128, feat_channels=128,
                            test_cfg=dict(score_thr=0.01, nms=dict(type="nms", iou_threshold=0.65))))
    assert isinstance(neck, YOLOXPAFPN) and isinstance(head, YOLOXHead)
    nk, hk = mmdet_ref.expected_keys([128, 256, 512], 128, 1, 80, 128)
    assert list(neck.state_dict().keys()) == nk
    assert list(head.state_dict().keys()) == hk
    assert head.test_cfg.score_thr == 0.01 and head.test_cfg.nms.iou_threshold == 0.65
    with pytest.raises(KeyError):
        NECKS.build(dict(type="NoSuchNeck"))
    # weights of the stock yolox-drone model load strictly after renaming (shapes agree)
    nsd, hsd = mmdet_ref.drone_to_mmdet_keys(_stock_sd())
    neck10 = YOLOXPAFPN([128, 256, 512], 128, num_csp_blocks=1)
    head10 = YOLOXHead(10, 128, feat_channels=128)
    neck10.load_state_dict(nsd, strict=True)
    head10.load_state_dict(hsd, strict=True)
    # and the module's own key translation inverts the oracle's
    back = {**neck10._plan_state_dict(), **head10._plan_state_dict()}
    want = {k: v for k, v in _stock_sd().items() if not k.startswith("backbone.backbone.")}
    assert set(back) == set(want)
    for k in want:
        assert torch.equal(back[k], want[k]), k


def test_mmdet_oracle_pinned_to_reference_source_golden():
    """oracle/mmdet_ref.py's YOLOX neck / head / get_bboxes against outputs recorded by executing the reference's own
    YOLOXPAFPN.forward, CSPLayer.forward, DarknetBottleneck.forward, YOLOXHead.forward / get_bboxes / _bbox_decode /
    _bboxes_nms and MlvlPointGenerator sources (tests/golden/make_golden_mmdet_yolox.py)."""
    import torch

    z = np.load(GOLD / "mmdet_yolox_cases.npz")
    H, W, seed = (int(v) for v in z["meta"])
    nsd, hsd = mmdet_ref.drone_to_mmdet_keys(_stock_sd())
    in_ch = [nsd["reduce_layers.1.conv.weight"].shape[0], nsd["reduce_layers.0.conv.weight"].shape[0], nsd["reduce_layers.0.conv.weight"].shape[1]]
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(2, c, H // s, W // s, generator=g) for c, s in zip(in_ch, (8, 16, 32))]
    outs = mmdet_ref.yolox_pafpn(nsd, feats)
    cls, box, obj = mmdet_ref.yolox_head_forward(hsd, outs)
    for l in range(3):
        assert np.allclose(outs[l].numpy(), z[f"neck{l}"], rtol=1e-5, atol=1e-5)
        assert np.allclose(cls[l].numpy(), z[f"cls{l}"], rtol=1e-4, atol=1e-5)
        assert np.allclose(box[l].numpy(), z[f"box{l}"], rtol=1e-4, atol=1e-5)
        assert np.allclose(obj[l].numpy(), z[f"obj{l}"], rtol=1e-4, atol=1e-5)
    gold_maps = [[torch.from_numpy(z[f"{n}{l}"]) for l in range(3)] for n in ("cls", "box", "obj")]
    sf = [[1.0, 1.0, 1.0, 1.0], [1.25, 1.5, 1.25, 1.5]]
    for tag, scale in (("plain", None), ("scaled", sf)):
        res = mmdet_ref.get_bboxes(*gold_maps, (8, 16, 32), 0.01, 0.65, scale_factors=scale)
        for i, (d, l) in enumerate(res):
            assert np.array_equal(np.asarray(l), z[f"{tag}_labels{i}"]), (tag, i)
            assert np.allclose(np.asarray(d), z[f"{tag}_dets{i}"], rtol=1e-6, atol=1e-5), (tag, i)


def test_mmdet_use_depthwise_matches_the_pinned_nano_model():
    """use_depthwise=True (mmcv DepthwiseSeparableConvModule in necks/yolox_pafpn.py:55, utils/csp_layer.py:44,
    dense_heads/yolox_head.py:146-147) is the yolox-drone DWConv under other key names: the nano weights of the stock
    YOLOX renamed onto the mmdet modules (oracle restatement AND the drop-in modules' own key translation) reproduce the
    golden logits recorded from the REAL models/base/yolox.py YoloBody(nc, 'nano') (tests/golden/make_golden_nano.py)."""
    from glsdet_b200.mmdet_face import YOLOXHead, YOLOXPAFPN

    z = np.load(GOLD / "nano_cases.npz")
    m = json.loads((GOLD / "nano_meta.json").read_text())["stock"]
    sd = ref_path.synthetic_state_dict(m["nc"], "nano", seed=m["seed"], flavour="calibrated", variant="stock")
    neck_sd, head_sd = mmdet_ref.drone_to_mmdet_keys(sd)
    assert any(".depthwise_conv." in k for k in neck_sd) and any(".pointwise_conv." in k for k in head_sd)
    feats = [torch.from_numpy(z[f"stock_dark{i}"]) for i in (3, 4, 5)]
    with torch.no_grad():
        p = mmdet_ref.yolox_pafpn(neck_sd, feats)
        cls, box, obj = mmdet_ref.yolox_head_forward(head_sd, p)
    for i in range(3):
        got = torch.cat([box[i], obj[i], cls[i]], 1).numpy()
        np.testing.assert_allclose(got, z[f"stock_logits{i}"], rtol=1e-4, atol=5e-5)
    # the drop-in modules: mmdet YOLOX-nano config (configs/yolox/yolox_nano_8x8_300e_coco.py shapes), strict load, and the
    # key translation back to the plan's naming is the inverse of the renaming above
    neck = YOLOXPAFPN(in_channels=[64, 128, 256], out_channels=64, num_csp_blocks=1, use_depthwise=True)
    head = YOLOXHead(num_classes=m["nc"], in_channels=64, feat_channels=64, use_depthwise=True)
    neck.load_state_dict(neck_sd, strict=True)
    head.load_state_dict(head_sd, strict=True)
    back = dict(neck._plan_state_dict())
    back.update(head._plan_state_dict())
    want = {k for k in sd if not k.startswith("backbone.backbone.")}
    assert set(back) == want
    assert all(torch.equal(back[k], sd[k]) for k in want)
