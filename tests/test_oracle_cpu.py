"""CPU tests: the oracle restatement against the golden vectors produced by the real reference
(tests/golden/make_golden.py), and the NMS oracle against torchvision's recorded outputs."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import nms_oracle, ref_path

GOLD = Path(__file__).resolve().parent / "golden"
META = json.loads((GOLD / "meta.json").read_text())


def _checksum(sd):
    return sum(float(v.double().abs().sum()) for k, v in sorted(sd.items()) if v.dtype.is_floating_point)


@pytest.mark.parametrize("meta", META["models"], ids=[m["name"] for m in META["models"]])
def test_oracle_model_matches_reference_golden(meta):
    z = np.load(GOLD / f"{meta['name']}.npz")
    sd = ref_path.synthetic_state_dict(meta["nc"], meta["phi"], seed=meta["seed"], flavour=meta["flavour"])
    assert len(sd) == meta["n_keys"]
    assert _checksum(sd) == pytest.approx(meta["weight_checksum"], rel=1e-9), "seeded weights differ from the golden run"
    feats = [torch.from_numpy(z[f"feat{i}"]) for i in range(4)]
    with torch.no_grad():
        neck = ref_path.pafpn_neck(sd, feats)
        logits = ref_path.yolox_head(sd, neck)
    for i in range(4):
        np.testing.assert_allclose(neck[i].numpy(), z[f"neck{i}"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(logits[i].numpy(), z[f"logits{i}"], rtol=1e-4, atol=2e-5)
    pred = ref_path.decode_outputs([torch.from_numpy(z[f"logits{i}"]) for i in range(4)], [meta["in_h"], meta["in_w"]])
    np.testing.assert_allclose(pred.numpy(), z["pred"], rtol=1e-6, atol=1e-7)
    res = ref_path.non_max_suppression(torch.from_numpy(z["pred"]), meta["nc"], [meta["in_h"], meta["in_w"]],
                                       np.array(meta["image_shape"]), meta["letterbox"], meta["conf"], meta["nms_thr"],
                                       strategy="auto_cpu")
    for b in range(meta["batch"]):
        assert res[b].shape == z[f"nms{b}"].shape
        np.testing.assert_array_equal(res[b], z[f"nms{b}"])


def test_oracle_postproc_matches_reference_golden():
    m = META["postproc"]
    z = np.load(GOLD / "postproc_reference.npz")
    pred = ref_path.decode_outputs([torch.from_numpy(z[f"logits{i}"]) for i in range(4)], [m["in_h"], m["in_w"]])
    pred[2, :, 4] = 0.0
    np.testing.assert_allclose(pred.numpy(), z["pred"], rtol=1e-6, atol=1e-7)
    res = ref_path.non_max_suppression(torch.from_numpy(z["pred"]), m["nc"], [m["in_h"], m["in_w"]],
                                       np.array(m["image_shape"]), m["letterbox"], m["conf"], m["nms_thr"], "auto_cpu")
    assert [len(r) for r in res] == m["kept"]
    assert res[2].shape == (0, 7)
    for b in range(m["batch"]):
        np.testing.assert_array_equal(res[b], z[f"nms{b}"])


@pytest.mark.parametrize("case", META["nms"]["cases"], ids=[c["name"] for c in META["nms"]["cases"]])
def test_nms_oracle_matches_torchvision_golden(case):
    z = np.load(GOLD / "nms_torchvision.npz")
    n = case["name"]
    b, s, l = z[f"{n}_boxes"], z[f"{n}_scores"], z[f"{n}_labels"]
    for thr in (0.45, 0.65):
        got_t = nms_oracle.batched_nms(b, s, l, thr, "trick")
        np.testing.assert_array_equal(got_t, z[f"{n}_keep_trick_{thr}"])
        got_v = nms_oracle.batched_nms(b, s, l, thr, "per_class")
        ref_v = z[f"{n}_keep_vanilla_{thr}"]
        if case["ties"]:  # torch.sort(descending) in the vanilla branch is not stable: tied scores may permute
            assert sorted(got_v.tolist()) == sorted(ref_v.tolist())
            np.testing.assert_array_equal(s[got_v], s[ref_v])
        else:
            np.testing.assert_array_equal(got_v, ref_v)
        got_a = nms_oracle.batched_nms(b, s, l, thr, "auto_cpu")
        if not (case["ties"] and 4 * case["k"] > 4000):
            np.testing.assert_array_equal(got_a, z[f"{n}_keep_auto_{thr}"])


def test_nms_oracle_c_build_matches_scalar_python():
    rng = np.random.default_rng(5)
    c = rng.uniform(0, 1, (200, 2))
    wh = rng.uniform(0.02, 0.3, (200, 2))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    scores = rng.uniform(0, 1, 200).astype(np.float32)
    for thr in (0.3, 0.5, 0.7):
        np.testing.assert_array_equal(nms_oracle.nms(boxes, scores, thr), nms_oracle.nms_python(boxes, scores, thr))


def test_nms_oracle_live_torchvision():
    """Same check against the installed torchvision binary, when importable (it is on both boxes)."""
    tv = pytest.importorskip("torchvision")
    rng = np.random.default_rng(11)
    for k in (0, 1, 37, 999, 1001, 2500):
        c = rng.uniform(0, 1, (k, 2))
        wh = np.exp(rng.normal(-3, 0.5, (k, 2)))
        boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
        scores = ((rng.permutation(k) + 1) / (k + 1.0)).astype(np.float32)
        labels = rng.integers(0, 10, k).astype(np.float32)
        ref = tv.ops.batched_nms(torch.from_numpy(boxes), torch.from_numpy(scores), torch.from_numpy(labels), 0.5)
        got = nms_oracle.batched_nms(boxes, scores, labels, 0.5, "auto_cpu")
        np.testing.assert_array_equal(got, ref.numpy())


def test_oracle_backbone_runs():
    sd = ref_path.synthetic_state_dict(3, "tiny", seed=0)
    feats = ref_path.csp_darknet(sd, torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(0)))
    assert [tuple(f.shape) for f in feats] == [(1, 48, 16, 16), (1, 96, 8, 8), (1, 192, 4, 4), (1, 384, 2, 2)]


def test_oracle_backbone_pinned_to_reference_golden():
    """oracle.ref_path.csp_darknet (and the whole image -> logits chain) against tests/golden/backbone_s.npz, which
    tests/golden/make_golden_backbone.py recorded from the REAL models/ffa/yolox_ffa.py YoloBody run from an image."""
    from pathlib import Path

    z = np.load(Path(__file__).resolve().parent / "golden" / "backbone_s.npz")
    sd = ref_path.synthetic_state_dict(10, "s", seed=0, flavour="calibrated")
    x = torch.from_numpy(z["image"])
    feats = ref_path.csp_darknet(sd, x)
    for name, f in zip(("dark2", "dark3", "dark4", "dark5"), feats):
        ref = torch.from_numpy(z[name])
        assert f.shape == ref.shape
        assert (f - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item()), name
    logits = ref_path.neck_head(sd, feats)
    for i, t in enumerate(logits):
        ref = torch.from_numpy(z[f"logits{i}"])
        assert (t - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item()), i


def test_oracle_decode_variants_pinned_to_reference_golden():
    """decode_outputs_{no_sigmoid, no_sigmoid_all, cls_sigmoid, xyxy} (utils_bbox.py:36-251): the restatement against the
    outputs of the real functions (tests/golden/make_golden_decode.py)."""
    z = np.load(GOLD / "decode_variants.npz")
    levels = [torch.from_numpy(z[f"level{i}"]) for i in range(3)]
    shape = [int(v) for v in z["input_shape"]]
    np.testing.assert_allclose(ref_path.decode_outputs(levels, shape).numpy(), z["default"], rtol=1e-6, atol=1e-6)
    for name in ("no_sigmoid", "no_sigmoid_all", "cls_sigmoid", "xyxy"):
        got = ref_path.decode_outputs_variant(levels, shape, name).numpy()
        np.testing.assert_allclose(got, z[name], rtol=1e-6, atol=1e-6, err_msg=name)


@pytest.mark.parametrize("case", ["p0", "stock", "p1", "p2"])
def test_oracle_nano_dwconv_pinned_to_reference_golden(case):
    """phi = 'nano' (DWConv = depthwise k x k + pointwise 1x1, models/base/baseConv.py:22-30): the oracle's backbone, neck,
    head, decode and NMS against tests/golden/nano_cases.npz, which tests/golden/make_golden_nano.py recorded from the REAL
    models/ffa/yolox_ffa.py / models/base/yolox.py YoloBody(nc, 'nano') run from an image; the state_dict keys (and their
    order) of the drop-in modules against the real modules' keys."""
    import json

    z = np.load(GOLD / "nano_cases.npz")
    m = json.loads((GOLD / "nano_meta.json").read_text())[case]
    sd = ref_path.synthetic_state_dict(m["nc"], "nano", seed=m["seed"], flavour="calibrated", variant=m["variant"])
    assert sorted(sd.keys()) == sorted(m["keys"])
    x = torch.from_numpy(z[f"{case}_image"])
    feats = ref_path.csp_darknet(sd, x)
    tol = lambda ref: 1e-4 * max(1.0, ref.abs().max().item())
    for name, f in zip(("dark2", "dark3", "dark4", "dark5"), feats):
        if f"{case}_{name}" in z.files:
            ref = torch.from_numpy(z[f"{case}_{name}"])
            assert f.shape == ref.shape and (f - ref).abs().max().item() <= tol(ref), name
    with torch.no_grad():
        if case == "p0":
            neck = ref_path.pafpn_neck(sd, feats)
            logits = ref_path.yolox_head(sd, neck)
        elif case == "p1":
            neck = ref_path.p1_neck(sd, feats)
            logits = ref_path.p1_head(sd, neck)
        elif case == "p2":
            neck = ref_path.p2_neck(sd, feats[1:])
            logits = ref_path.stock_head(sd, neck)
        else:
            neck = ref_path.pafpn_neck(sd, [None] + list(feats[1:]))[1:]
            logits = ref_path.stock_head(sd, neck)
    for i, t in enumerate(neck):
        ref = torch.from_numpy(z[f"{case}_neck{i}"])
        assert (t - ref).abs().max().item() <= tol(ref), ("neck", i)
    for i, t in enumerate(logits):
        ref = torch.from_numpy(z[f"{case}_logits{i}"])
        assert (t - ref).abs().max().item() <= tol(ref), ("logits", i)
    hw = [m["in_h"], m["in_w"]]
    pred = torch.from_numpy(z[f"{case}_pred"])
    res = ref_path.non_max_suppression(pred.clone(), m["nc"], hw, np.array(hw), False, m["conf"], m["nms_thr"])
    for b in range(m["batch"]):
        assert np.array_equal(res[b], z[f"{case}_nms{b}"]), b
    # the drop-in modules expose the reference's keys in the reference's order (strict load works, DataParallel-safe)
    import importlib

    YoloBody = importlib.import_module({"p0": "glsdet_b200.yolox_ffa", "stock": "glsdet_b200.yolox_base", "p1": "glsdet_b200.yolox10",
                                        "p2": "glsdet_b200.yolo_patch_nonlocal_plus"}[case]).YoloBody
    net = YoloBody(m["nc"], "nano")
    assert list(net.state_dict().keys()) == m["keys"]
    net.load_state_dict(sd, strict=True)
