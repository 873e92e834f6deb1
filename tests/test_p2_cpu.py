"""CPU checks of the GLSDet P2 slice (models/block/non_local/yolo_patch_nonlocal_plus.py): the oracle against the
reference's golden vectors and the state_dict contract of the drop-in module."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ref_path

GOLD = Path(__file__).resolve().parent / "golden"
META = json.loads((GOLD / "meta.json").read_text())["p2"]


def _sd():
    return ref_path.synthetic_state_dict(META["nc"], META["phi"], seed=META["seed"], flavour="calibrated", variant="p2")


def test_p2_oracle_matches_reference_golden():
    z = np.load(GOLD / f"{META['name']}.npz")
    sd = _sd()
    assert len(sd) == META["n_keys"] == 600
    feats = [torch.from_numpy(z[f"feat{i}"]) for i in range(3)]
    with torch.no_grad():
        f1p = ref_path.patch_conv(sd, "backbone.Patch_conv_feat1", feats[0], stride=2, nonlocal_=True)
        f2p = ref_path.patch_conv(sd, "backbone.Patch_conv_feat2", feats[1], stride=1, nonlocal_=False)
        neck = ref_path.p2_neck(sd, feats)
        logits = ref_path.stock_head(sd, neck)
    np.testing.assert_allclose(f1p.numpy(), z["feat1_patch"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(f2p.numpy(), z["feat2_patch"], rtol=1e-4, atol=1e-5)
    for i in range(3):
        np.testing.assert_allclose(neck[i].numpy(), z[f"neck{i}"], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(logits[i].numpy(), z[f"logits{i}"], rtol=1e-4, atol=5e-5)
    pred = ref_path.decode_outputs([torch.from_numpy(z[f"logits{i}"]) for i in range(3)], [META["in_h"], META["in_w"]])
    np.testing.assert_allclose(pred.numpy(), z["pred"], rtol=1e-6, atol=1e-7)
    res = ref_path.non_max_suppression(torch.from_numpy(z["pred"]), META["nc"], [META["in_h"], META["in_w"]],
                                       np.array([META["in_h"], META["in_w"]]), False, META["conf"], META["nms_thr"],
                                       strategy="auto_cpu")
    for b in range(META["batch"]):
        np.testing.assert_array_equal(res[b], z[f"nms{b}"])


def test_p2_state_dict_keys_match_reference():
    from glsdet_b200.yolo_patch_nonlocal_plus import YoloBody

    ref = json.loads((GOLD / "state_dict_keys_p2_s.json").read_text())
    net = YoloBody(10, "s")
    mine = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert list(mine.keys()) == list(ref.keys())
    assert mine == ref
    # identity initialisation of the k x k convs (Identity_Conv.py:27-84)
    w = net.backbone.P3_Identity.conv.weight
    assert w.shape[-1] == 7 and float(w.sum()) == w.shape[0] and float(w[5, 5, 3, 3]) == 1.0
    net.load_state_dict(_sd(), strict=True)
    with pytest.raises(RuntimeError, match="libglsdet_b200"):
        net.backbone.Patch_conv_feat2(torch.zeros(1, 256, 4, 4))
