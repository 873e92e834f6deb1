"""CPU checks of the MP-Det drop-ins (no GPU): parameter layout against mmdet's names and the restated oracle's
internal consistency (PARITY UNPINNED: mmcv is absent, see oracle/mmdet_ref.py)."""
import torch

from oracle import mmdet_ref as M


def test_mpdet_state_dict_layout():
    import glsdet_b200.mpdet  # noqa: F401
    from glsdet_b200.mmdet_face import HEADS, NECKS

    neck = NECKS.build(dict(type="FPN", in_channels=[256, 512, 1024, 2048], out_channels=256, start_level=1,
                            add_extra_convs="on_output", num_outs=5))
    head = HEADS.build(dict(type="MPHead", num_classes=10, in_channels=256, stacked_convs=4, feat_channels=256))
    want = M.mpdet_state_dict_shapes()
    got = {"neck." + k: tuple(v.shape) for k, v in neck.state_dict().items()}
    got.update({"bbox_head." + k: tuple(v.shape) for k, v in head.state_dict().items()})
    assert got == {k: tuple(v) for k, v in want.items()}
    # mmdet names (necks/fpn.py:107-144, dense_heads/mp_head.py:42-98)
    for k in ("neck.lateral_convs.0.conv.weight", "neck.fpn_convs.4.conv.bias", "bbox_head.cls_convs.3.gn.weight",
              "bbox_head.reg_convs.0.conv.weight", "bbox_head.gfl_cls_conv.bias", "bbox_head.gfl_reg.weight",
              "bbox_head.scales.4.scale", "bbox_head.proxies", "bbox_head._embedding", "bbox_head._proxies_prob"):
        assert k in got, k
    assert "bbox_head.cls_convs.0.conv.bias" not in got      # ConvModule with a norm layer has no conv bias
    import pytest
    with pytest.raises(RuntimeError, match="inference-only"):
        head.train()


def test_mpdet_oracle_shapes_and_proxy_scores():
    sd = M.mpdet_synthetic_state_dict(0)
    g = torch.Generator().manual_seed(1)
    ins = [torch.randn(1, c, 128 // s, 192 // s, generator=g) for c, s in zip((256, 512, 1024, 2048), (4, 8, 16, 32))]
    nsd = {k[5:]: v for k, v in sd.items() if k.startswith("neck.")}
    hsd = {k[10:]: v for k, v in sd.items() if k.startswith("bbox_head.")}
    with torch.no_grad():
        outs = M.fpn_forward(nsd, ins)
        cls, box = M.mp_head_forward(hsd, outs)
    assert [tuple(o.shape[2:]) for o in outs] == [(16, 24), (8, 12), (4, 6), (2, 3), (1, 2)]
    assert all(c.shape[1] == 10 for c in cls) and all(b.shape[1] == 68 for b in box)
    # gamma-weighted cosine similarity lies in [-gamma, gamma]
    assert all(float(c.abs().max()) <= 10.0 + 1e-4 for c in cls)
    dets, labels = M.gfl_get_bboxes_single([c[0] for c in cls], [b[0] for b in box], (128, 190))
    assert dets.shape[1] == 5 and len(dets) == len(labels) <= 500
    assert (dets[:-1, 4] >= dets[1:, 4]).all() and float(dets[:, 2].max()) <= 190.0 and float(dets[:, 3].max()) <= 128.0


def _golden_case(case):
    import numpy as np
    from pathlib import Path

    z = np.load(Path(__file__).parent / "golden" / "mpdet_cases.npz")
    H, W, seed = (int(v) for v in z[f"{case}_meta"])
    g = torch.Generator().manual_seed(seed)
    ins = [torch.randn(1, c, -(-H // s), -(-W // s), generator=g) for c, s in zip((256, 512, 1024, 2048), (4, 8, 16, 32))]
    return z, H, W, ins


def test_mpdet_oracle_pinned_to_reference_source_golden():
    """oracle/mmdet_ref.py against outputs recorded by executing the reference's own FPN.forward / MPHead.forward_single /
    forward_proxy / Integral / _get_bboxes_single / filter_scores_and_topk / distance2bbox / _bbox_post_process sources
    (tests/golden/make_golden_mpdet.py).  Only the mmcv pieces (ConvModule, Scale, anchors, batched_nms) are restated."""
    import numpy as np

    sd = M.mpdet_synthetic_state_dict(0)
    nsd = {k[5:]: v for k, v in sd.items() if k.startswith("neck.")}
    hsd = {k[10:]: v for k, v in sd.items() if k.startswith("bbox_head.")}
    for case in ("small", "odd"):
        z, H, W, ins = _golden_case(case)
        fs, bs = (int(v) for v in z[f"{case}_stride"])
        with torch.no_grad():
            outs = M.fpn_forward(nsd, ins)
            cls, box = M.mp_head_forward(hsd, outs)
            dets, labels = M.gfl_get_bboxes_single([c[0] for c in cls], [b[0] for b in box], (H, W - 11))
        for l in range(5):
            assert np.allclose(outs[l][:, ::fs].numpy(), z[f"{case}_fpn{l}"], rtol=1e-5, atol=1e-5), (case, "fpn", l)
            assert np.allclose(cls[l].numpy(), z[f"{case}_cls{l}"], rtol=1e-4, atol=1e-4), (case, "cls", l)
            assert np.allclose(box[l][:, ::bs].numpy(), z[f"{case}_box{l}"], rtol=1e-4, atol=1e-4), (case, "box", l)
        assert dets.shape == z[f"{case}_dets"].shape, (case, dets.shape)
        assert np.array_equal(labels.numpy(), z[f"{case}_labels"])
        assert np.allclose(dets.numpy(), z[f"{case}_dets"], rtol=1e-5, atol=1e-4)
