"""CPU checks of the GLSDet P1 slice (models/new/yolox10.py): oracle vs the reference's golden vectors, the
state_dict contract of the drop-in module, and the reassociated form of the non-local block that the CUDA path uses."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ref_path

GOLD = Path(__file__).resolve().parent / "golden"
META = json.loads((GOLD / "meta.json").read_text())["p1"]


def _sd():
    return ref_path.synthetic_state_dict(META["nc"], META["phi"], seed=META["seed"], flavour="calibrated", variant="p1")


def test_p1_oracle_matches_reference_golden():
    z = np.load(GOLD / f"{META['name']}.npz")
    sd = _sd()
    assert len(sd) == META["n_keys"] == 654
    feats = [torch.from_numpy(z[f"feat{i}"]) for i in range(4)]
    with torch.no_grad():
        h2, w2 = feats[2].shape[2] // 2, feats[2].shape[3] // 2
        nl = ref_path.non_local_block(sd, "backbone.Patch_conv_feat2.feat_patchconv_lt_nonlocal", feats[2][:, :, :h2, :w2])
        pc = ref_path.patch_conv_nonlocal_new(sd, "backbone.Patch_conv_feat2", feats[2])
        neck = ref_path.p1_neck(sd, feats)
        logits = ref_path.p1_head(sd, neck)
    np.testing.assert_allclose(nl.numpy(), z["nonlocal_lt_feat2"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(pc.numpy(), z["patchconv_feat2"], rtol=1e-4, atol=1e-5)
    for i in range(4):
        np.testing.assert_allclose(neck[i].numpy(), z[f"neck{i}"], rtol=1e-4, atol=2e-5)
    assert len(logits) == 3
    for i in range(3):
        np.testing.assert_allclose(logits[i].numpy(), z[f"logits{i}"], rtol=1e-4, atol=5e-5)
    pred = ref_path.decode_outputs([torch.from_numpy(z[f"logits{i}"]) for i in range(3)], [META["in_h"], META["in_w"]])
    np.testing.assert_allclose(pred.numpy(), z["pred"], rtol=1e-6, atol=1e-7)
    res = ref_path.non_max_suppression(torch.from_numpy(z["pred"]), META["nc"], [META["in_h"], META["in_w"]],
                                       np.array([META["in_h"], META["in_w"]]), False, META["conf"], META["nms_thr"],
                                       strategy="auto_cpu")
    for b in range(META["batch"]):
        np.testing.assert_array_equal(res[b], z[f"nms{b}"])


def test_p1_state_dict_keys_match_reference():
    from glsdet_b200.yolox10 import YoloBody

    ref = json.loads((GOLD / "state_dict_keys_p1_s.json").read_text())
    net = YoloBody(10, "s")
    mine = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert list(mine.keys()) == list(ref.keys())
    assert mine == ref
    net.load_state_dict(_sd(), strict=True)
    with pytest.raises(RuntimeError, match="libglsdet_b200"):
        net.backbone.Patch_conv_feat1(torch.zeros(1, 128, 4, 4))
    with pytest.raises(RuntimeError, match="inference-only"):
        net.train()


def test_non_local_block_reassociation_identity():
    """Non_local_family.py:32-48 without softmax is linear in the pairwise matrix, so
    x + conv_out((theta^T phi / T) g) == x + W_eff x + b_eff with W_eff = (Wo [Wg|bg] / T) S ([Wphi|bphi]^T [Wtheta|btheta])
    and S the Gram matrix of [x | 1] - the form engine.FFAPathPlan._build_nonlocal evaluates on the GPU."""
    sd = _sd()
    g = torch.Generator().manual_seed(4)
    for C, H, W, p in ((128, 6, 10, "backbone.Patch_conv_feat1.feat_patchconv_rb_nonlocal"),
                       (512, 3, 5, "backbone.Patch_conv_feat3.feat_patchconv_lb_nonlocal")):
        x = torch.randn(2, C, H, W, generator=g)
        ref = ref_path.non_local_block(sd, p, x).double()
        wg, wt, wp, wo = (sd[f"{p}.{n}.weight"].double().flatten(1) for n in ("g", "theta", "phi", "conv_out"))
        bg, bt, bp, bo = (sd[f"{p}.{n}.bias"].double() for n in ("g", "theta", "phi", "conv_out"))
        T = H * W
        a1 = wo @ torch.cat([wg, bg[:, None]], 1) / T
        a2 = torch.cat([wp, bp[:, None]], 1).t() @ torch.cat([wt, bt[:, None]], 1)
        X = x.double().flatten(2).permute(0, 2, 1)
        Xa = torch.cat([X, torch.ones(2, T, 1, dtype=torch.float64)], 2)
        Wa = a1 @ (Xa.transpose(1, 2) @ Xa) @ a2
        out = X + X @ Wa[:, :, :C].transpose(1, 2) + (Wa[:, :, C] + bo)[:, None, :]
        out = out.permute(0, 2, 1).reshape(2, C, H, W)
        assert (out - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
