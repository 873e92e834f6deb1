"""GPU parity of the MP-Det neck / head (BASELINE configs[2]) through the C ABI against the
restated oracle (oracle/mmdet_ref.py) and against goldens recorded by executing the reference's own sources
(tests/golden/make_golden_mpdet.py; mmcv is absent, so only its pieces are restated).
Tolerances: bf16 path, 2e-2 relative l2 (feature maps, class scores, box distributions); kernels without bf16 storage
(GroupNorm, proxy scores, integral decode, selection) against fp32 torch at 1e-4 / bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from _helpers import TOL, assert_close_rel
from oracle import mmdet_ref as M
from oracle import ref_path

pytestmark = pytest.mark.gpu


def test_group_norm_relu_kernel(native_lib, cuda_device):
    from glsdet_b200 import _native as N

    dev = cuda_device
    g = torch.Generator().manual_seed(0)
    for (B, H, W, C) in ((2, 13, 21, 256), (1, 100, 168, 256), (3, 7, 11, 128)):
        x = (torch.randn(B, C, H, W, generator=g) * 2 + 0.5).to(torch.bfloat16)
        gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
        ref = torch.relu(F.group_norm(x.float(), 32, gamma, beta, eps=1e-5))
        buf = x.permute(0, 2, 3, 1).contiguous().to(dev)
        scratch = torch.zeros(int(native_lib.glsdet_group_norm_scratch_floats(B, C)), device=dev)
        dg, db = gamma.to(dev), beta.to(dev)
        N.check(native_lib.glsdet_group_norm_relu(buf.data_ptr(), B, H * W, C, C, 32, dg.data_ptr(), db.data_ptr(), 1e-5,
                                                  scratch.data_ptr(), None), "gn")
        torch.cuda.synchronize()
        got = buf.float().permute(0, 3, 1, 2).cpu()
        assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()


def test_proxy_scores_and_decode_kernels(native_lib, cuda_device):
    from glsdet_b200 import _native as N

    dev = cuda_device
    g = torch.Generator().manual_seed(1)
    B, H, W, C, nc = 2, 9, 14, 256, 10
    feat = torch.randn(B, H, W, C, generator=g)
    prox = torch.randn(sum(M.MP_PROXIES), C, generator=g)
    ref = M.mp_forward_proxy(feat.reshape(-1, C), prox).reshape(B, H * W, nc)
    centers = F.normalize(prox, p=2, dim=1).contiguous().to(dev)
    starts = torch.tensor(np.concatenate([[0], np.cumsum(M.MP_PROXIES)]), dtype=torch.int32, device=dev)
    rows = torch.full((B, H * W + 5, nc), float("nan"), device=dev)
    dfeat = feat.to(dev)
    N.check(native_lib.glsdet_proxy_scores(dfeat.data_ptr(), centers.data_ptr(), starts.data_ptr(), nc, centers.shape[0],
                                           C, B, H * W, 10.0, rows.data_ptr(), nc, (H * W + 5) * nc, 3, None), "proxy")
    torch.cuda.synchronize()
    assert torch.allclose(rows[:, 3:3 + H * W].cpu(), ref, rtol=1e-4, atol=1e-4)
    assert torch.isnan(rows[:, :3]).all() and torch.isnan(rows[:, 3 + H * W:]).all()
    # integral decode
    reg = torch.randn(B, H, W, 80, generator=g) * 2
    boxes = torch.full((B, H * W + 2, 4), float("nan"), device=dev)
    dreg = reg.to(dev)
    N.check(native_lib.glsdet_gfl_decode(dreg.data_ptr(), 80, 17, B, H, W, 16.0, 200.0, 130.0, boxes.data_ptr(),
                                         (H * W + 2) * 4, 1, None), "decode")
    torch.cuda.synchronize()
    for b in range(B):
        want = M.gfl_decode_level(reg[b, :, :, :68].permute(2, 0, 1), 16, (130, 200))
        assert torch.allclose(boxes[b, 1:1 + H * W].cpu(), want, rtol=1e-5, atol=1e-4)


def _models(dev, seed=0):
    from glsdet_b200.mmdet_face import HEADS, NECKS
    import glsdet_b200.mpdet  # noqa: F401  (registers FPN / MPHead)

    sd = M.mpdet_synthetic_state_dict(seed)
    neck = NECKS.build(dict(type="FPN", in_channels=[256, 512, 1024, 2048], out_channels=256, start_level=1,
                            add_extra_convs="on_output", num_outs=5))
    head = HEADS.build(dict(type="MPHead", num_classes=10, in_channels=256, stacked_convs=4, feat_channels=256,
                            anchor_generator=dict(type="AnchorGenerator", ratios=[1.0], octave_base_scale=8,
                                                  scales_per_octave=1, strides=[8, 16, 32, 64, 128]),
                            test_cfg=dict(nms_pre=1000, min_bbox_size=0, score_thr=0.05,
                                          nms=dict(type="nms", iou_threshold=0.6), max_per_img=500)))
    nsd = {k[5:]: v for k, v in sd.items() if k.startswith("neck.")}
    hsd = {k[10:]: v for k, v in sd.items() if k.startswith("bbox_head.")}
    assert list(neck.state_dict().keys()) == list(nsd.keys())
    assert set(head.state_dict().keys()) == set(hsd.keys())
    neck.load_state_dict(nsd, strict=True)
    head.load_state_dict(hsd, strict=True)
    return neck.to(dev).eval(), head.to(dev).eval(), nsd, hsd


@pytest.mark.parametrize("size", [(192, 256), (800, 1344)])
def test_mpdet_fpn_and_head_vs_oracle(size, native_lib, cuda_device):
    """FPN + MPHead forward: config 3 shapes at 800 x 1344 (C3 100x168 .. P7 7x11, odd extra levels) and a small case."""
    H, W = size
    neck, head, nsd, hsd = _models(cuda_device)
    g = torch.Generator().manual_seed(7)
    ins = [torch.randn(1, c, H // s, W // s, generator=g) for c, s in zip((256, 512, 1024, 2048), (4, 8, 16, 32))]
    with torch.no_grad():
        ref_fpn = M.fpn_forward(nsd, ins)
        ref_cls, ref_box = M.mp_head_forward(hsd, ref_fpn)
    outs = neck([t.to(cuda_device) for t in ins])
    assert [tuple(o.shape) for o in outs] == [tuple(o.shape) for o in ref_fpn]
    for i, (a, b) in enumerate(zip(outs, ref_fpn)):
        assert_close_rel(a, b, TOL, f"FPN out {i}")
    # the head is checked on the ORACLE's FPN outputs (isolates it from the neck's bf16 error)
    cls, box = head([t.to(cuda_device) for t in ref_fpn])
    ref_path._EMULATE_BF16 = True
    try:
        with torch.no_grad():
            emu_cls, emu_box = M.mp_head_forward(hsd, ref_fpn)
    finally:
        ref_path._EMULATE_BF16 = False
    for l in range(5):
        inh_c = ((emu_cls[l] - ref_cls[l]).norm() / ref_cls[l].norm()).item()
        inh_b = ((emu_box[l] - ref_box[l]).norm() / ref_box[l].norm()).item()
        assert_close_rel(cls[l], ref_cls[l], max(TOL, 1.3 * inh_c), f"MPHead cls_score level {l}", frac=5e-2)
        assert_close_rel(box[l], ref_box[l], max(TOL, 1.3 * inh_b), f"MPHead bbox_pred level {l}", frac=5e-2)


def test_mpdet_get_bboxes_vs_oracle(native_lib, cuda_device):
    """Selection + decode + NMS fed with identical maps: the native post-processing on the maps the native head produced
    must equal the oracle's post-processing of those same maps (boxes to 1e-3 px, scores 1e-6, same labels and order)."""
    H, W = 320, 448
    neck, head, nsd, hsd = _models(cuda_device, seed=3)
    g = torch.Generator().manual_seed(9)
    ins = [torch.randn(2, c, H // s, W // s, generator=g) for c, s in zip((256, 512, 1024, 2048), (4, 8, 16, 32))]
    feats = neck([t.to(cuda_device) for t in ins])
    metas = [dict(img_shape=(H, W - 11, 3), scale_factor=1.0)] * 2
    res = head.detect(feats, metas)
    cls, box = head(feats)
    for b in range(2):
        dets, labels = M.gfl_get_bboxes_single([c[b].cpu() for c in cls], [x[b].cpu() for x in box], (H, W - 11))
        got_d, got_l = res[b][0].cpu(), res[b][1].cpu()
        assert got_d.shape == dets.shape and len(dets) > 20
        assert torch.equal(got_l, labels)
        assert torch.allclose(got_d[:, 4], dets[:, 4], rtol=0, atol=2e-6)
        assert torch.allclose(got_d[:, :4], dets[:, :4], rtol=0, atol=2e-3)


def test_mpdet_vs_reference_source_golden(native_lib, cuda_device):
    """FPN / MPHead maps against outputs recorded by executing the reference's own sources (tests/golden/make_golden_mpdet.py,
    case 'small'): the neck from the inputs, the head from the golden-consistent oracle FPN maps."""
    import numpy as np
    from pathlib import Path

    z = np.load(Path(__file__).parent / "golden" / "mpdet_cases.npz")
    H, W, seed = (int(v) for v in z["small_meta"])
    fs, bs = (int(v) for v in z["small_stride"])
    g = torch.Generator().manual_seed(seed)
    ins = [torch.randn(1, c, H // s, W // s, generator=g) for c, s in zip((256, 512, 1024, 2048), (4, 8, 16, 32))]
    neck, head, nsd, hsd = _models(cuda_device)
    outs = neck([t.to(cuda_device) for t in ins])
    for l in range(5):
        assert_close_rel(outs[l][:, ::fs], torch.from_numpy(z[f"small_fpn{l}"]), TOL, f"FPN out {l} vs golden")
    with torch.no_grad():
        ref_fpn = M.fpn_forward(nsd, ins)
    cls, box = head([t.to(cuda_device) for t in ref_fpn])
    for l in range(5):
        assert_close_rel(cls[l], torch.from_numpy(z[f"small_cls{l}"]), 3e-2, f"cls {l} vs golden", frac=5e-2)
        assert_close_rel(box[l][:, ::bs], torch.from_numpy(z[f"small_box{l}"]), 3e-2, f"box {l} vs golden", frac=5e-2)


def test_mpdet_get_bboxes_api_rescale_and_max_num(native_lib, cuda_device):
    """get_bboxes(cls_scores, bbox_preds, img_metas=..., rescale=...) on forward()'s maps == the fused detect(); rescale
    divides the boxes by scale_factor before the NMS (base_dense_head.py:282-283); nms max_num caps the result."""
    H, W = 192, 256
    neck, head, nsd, hsd = _models(cuda_device, seed=3)
    g = torch.Generator().manual_seed(5)
    ins = [torch.randn(2, c, H // s, W // s, generator=g) for c, s in zip((256, 512, 1024, 2048), (4, 8, 16, 32))]
    feats = neck([t.to(cuda_device) for t in ins])
    metas = [dict(img_shape=(H, W, 3), scale_factor=[2.0, 2.0, 2.0, 2.0])] * 2
    fused = head.simple_test(feats, metas, rescale=False)
    cls, box = head(feats)
    api = head.get_bboxes(cls, box, img_metas=metas, rescale=False)
    for (d0, l0), (d1, l1) in zip(fused, api):
        assert torch.equal(l0, l1) and torch.equal(d0, d1) and len(d0) > 5
    scaled = head.get_bboxes(cls, box, img_metas=metas, rescale=True)
    for b in range(2):
        dets, labels = M.gfl_get_bboxes_single([c[b].cpu() for c in cls], [x[b].cpu() for x in box], (H, W), scale_factor=2.0)
        assert scaled[b][0].shape == dets.shape and torch.equal(scaled[b][1].cpu(), labels)
        assert torch.allclose(scaled[b][0].cpu(), dets, rtol=0, atol=2e-3)
    capped = head.detect(feats, metas, cfg=dict(nms_pre=1000, score_thr=0.05, nms=dict(type="nms", iou_threshold=0.6, max_num=7), max_per_img=500))
    assert all(len(d) == min(7, len(f[0])) and torch.equal(d, f[0][:7]) for (d, _), f in zip(capped, fused))


@pytest.mark.parametrize("dist", ["wide", "narrow", "ties", "few"])
def test_gfl_select_kernel_distributions(dist, native_lib, cuda_device):
    """glsdet_gfl_select against `scores > thr` + stable top-k by (score desc, flattened index asc) (gfl_head.py:440-452):
    wide logits (one histogram pass), a narrow band (sub-bin refinement inside the crossing bin), massive exact ties
    (global-memory fallback), and fewer candidates than nms_pre.  Bit-exact."""
    from glsdet_b200 import _native as N
    dev = cuda_device
    B, A, nc, topk, thr = 3, 9000, 10, 1000, 0.05
    g = torch.Generator().manual_seed(11)
    if dist == "wide":
        logits = torch.randn(B, A, nc, generator=g) * 3.0
    elif dist == "narrow":
        logits = 0.4 + torch.randn(B, A, nc, generator=g) * 1e-3
    elif dist == "ties":
        logits = torch.full((B, A, nc), 0.25)
        logits[:, ::70, 3] = 0.5
    else:
        logits = torch.full((B, A, nc), -8.0)
        logits[:, :50, 2] = torch.randn(B, 50, generator=g)
    boxes = torch.rand(B, A, 4, generator=g)
    boxes[:, :, 0] = torch.arange(A, dtype=torch.float32)
    rows, bx = logits.to(dev).contiguous(), boxes.to(dev).contiguous()
    ks = 1
    while ks < A * nc:
        ks <<= 1
    keys = torch.empty((B, ks), dtype=torch.int64, device=dev)
    cap = topk
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    cb = torch.zeros((B, cap, 4), device=dev)
    cs = torch.zeros((B, cap), device=dev)
    cl = torch.zeros((B, cap), device=dev)
    scratch = torch.zeros(int(native_lib.glsdet_gfl_select_scratch_ints(B)), dtype=torch.int32, device=dev)
    for rep in range(2):   # the second call checks that the scratch was left clean
        cnt.zero_()
        N.check(native_lib.glsdet_gfl_select(rows.data_ptr(), nc, A * nc, bx.data_ptr(), A * 4, 0, A, nc, thr, topk, B, keys.data_ptr(),
                                             ks, cnt.data_ptr(), cb.data_ptr(), cs.data_ptr(), cl.data_ptr(), cap, scratch.data_ptr(),
                                             N.stream_ptr(None)), "glsdet_gfl_select")
        torch.cuda.synchronize()
        assert int(scratch.abs().sum()) == 0
        sc = torch.sigmoid(rows.float()).reshape(B, -1).cpu()
        for b in range(B):
            s = sc[b]
            idx = torch.nonzero(s > thr).flatten()
            order = torch.sort(s[idx], descending=True, stable=True).indices[:topk]
            sel = idx[order]
            n = int(cnt[b])
            assert n == sel.numel(), (dist, b, n, sel.numel())
            got_sc = cs[b, :n].cpu()
            flat = cb[b, :n, 0].cpu().long() * nc + cl[b, :n].cpu().long()      # boxes[..., 0] carries the anchor index
            assert flat.unique().numel() == n
            assert torch.equal(cb[b, :n].cpu(), boxes[b][flat // nc])
            # scores are the device's expf sigmoid: equal to torch's within an ulp, so order / membership are checked on
            # the returned scores and against torch's with that slack
            assert torch.allclose(got_sc, s[flat], rtol=0, atol=2e-7)
            d = got_sc[1:] - got_sc[:-1]
            assert bool(((d < 0) | ((d == 0) & (flat[1:] > flat[:-1]))).all()), "not sorted by (score desc, index asc)"
            if n:
                rest = torch.ones_like(s, dtype=torch.bool)
                rest[flat] = False
                rest &= s > thr
                if rest.any():
                    assert float(s[rest].max()) <= float(got_sc.min()) + 2e-7
            if dist in ("ties", "few"):
                assert torch.equal(flat, sel)
