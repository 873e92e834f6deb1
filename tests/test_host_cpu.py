"""CPU tests of the host side: C-ABI library surface, module/state_dict contract, weight packing, BN folding."""
import json
import re
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


def test_library_exports_every_declared_symbol(native_lib):
    from glsdet_b200 import _native

    header = (ROOT / "include" / "glsdet_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(glsdet_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found in the header"
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    for name in declared:
        assert hasattr(native_lib, name), f"{name} is declared in include/glsdet_b200.h but not exported"
    assert native_lib.glsdet_abi_version() == 2


def test_no_compute_without_gpu_but_errors_are_reported(native_lib):
    from glsdet_b200 import _native as N

    assert native_lib.glsdet_nms_workspace_bytes(2, 1000, 10) > 0
    assert native_lib.glsdet_nms_workspace_bytes(0, 1000, 10) < 0
    rc = native_lib.glsdet_conv_launch(None, None)
    assert rc != 0 and b"null op" in native_lib.glsdet_last_error()
    d = N.ConvDesc()
    d.ksize, d.stride, d.batch, d.height, d.width = 4, 1, 1, 8, 8   # even kernel sizes are rejected
    assert native_lib.glsdet_conv_weight_shape(d, None, None, None) != 0
    assert b"ksize" in native_lib.glsdet_last_error()


def test_state_dict_keys_match_reference():
    from glsdet_b200.yolox_ffa import YoloBody

    ref = json.loads((GOLD / "state_dict_keys_p0_s.json").read_text())
    net = YoloBody(10, "s")
    mine = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert list(mine.keys()) == list(ref.keys())
    assert mine == ref
    assert not net.training and not net.backbone.backbone.stem.conv.bn.training


def test_strict_load_of_reference_style_checkpoint_and_no_torch_forward():
    from glsdet_b200.yolox_ffa import YoloBody
    from oracle import ref_path

    sd = ref_path.synthetic_state_dict(3, "tiny", seed=5)
    net = YoloBody(3, "tiny")
    net.load_state_dict(sd, strict=True)
    with pytest.raises(RuntimeError, match="native GLSDet path"):
        net.head.stems[0](torch.zeros(1, 96, 4, 4))
    with pytest.raises(RuntimeError, match="inference-only"):
        net.train()
    # the backbone is native too: no PyTorch / CPU forward anywhere
    with pytest.raises(RuntimeError, match="native GLSDet path"):
        net.backbone.backbone.dark2[0](torch.zeros(1, 24, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU"):
        net.backbone.features(torch.zeros(1, 3, 64, 96))


def test_fold_bn_and_pack_order():
    from glsdet_b200.ops import fold_bn, pack_conv_weight

    g = torch.Generator().manual_seed(0)
    w = torch.randn(24, 40, 3, 3, generator=g)
    gamma, beta = torch.randn(24, generator=g), torch.randn(24, generator=g)
    mean, var = torch.randn(24, generator=g), torch.rand(24, generator=g) + 0.5
    x = torch.randn(2, 40, 9, 9, generator=g)
    ref = F.batch_norm(F.conv2d(x, w, padding=1), mean, var, gamma, beta, False, 0.0, 1e-3)
    wf, bf = fold_bn(w, gamma, beta, mean, var, 1e-3)
    assert torch.allclose(F.conv2d(x, wf, bf, padding=1), ref, rtol=1e-4, atol=1e-4)

    # K order: (source, tap, channel) with 64-channel zero padding per (source, tap)
    packed = pack_conv_weight(w, [24, 16], 32, 9 * 128).float()
    assert packed.shape == (32, 1152) and packed[24:].abs().max() == 0
    wb = w.to(torch.bfloat16).float()
    for tap in (0, 4, 8):
        ky, kx = divmod(tap, 3)
        seg0 = packed[:24, tap * 64: tap * 64 + 64]
        assert torch.equal(seg0[:, :24], wb[:, :24, ky, kx]) and seg0[:, 24:].abs().max() == 0
        seg1 = packed[:24, 9 * 64 + tap * 64: 9 * 64 + tap * 64 + 64]
        assert torch.equal(seg1[:, :16], wb[:, 24:, ky, kx]) and seg1[:, 16:].abs().max() == 0


def test_yolo_correct_boxes_matches_oracle():
    from glsdet_b200.utils_bbox import yolo_correct_boxes
    from oracle import ref_path

    rng = np.random.default_rng(0)
    xy = rng.uniform(0, 1, (50, 2)).astype(np.float32)
    wh = rng.uniform(0, 0.4, (50, 2)).astype(np.float32)
    for lb in (True, False):
        a = yolo_correct_boxes(xy, wh, [640, 1024], np.array([540, 1024]), lb)
        b = ref_path.yolo_correct_boxes(xy, wh, [640, 1024], np.array([540, 1024]), lb)
        np.testing.assert_array_equal(a, b)


def test_cpu_tensors_are_rejected_loudly(native_lib):
    from glsdet_b200 import _native as N
    from glsdet_b200.utils_bbox import batched_nms, decode_outputs, non_max_suppression

    with pytest.raises(N.NativeError):
        decode_outputs([torch.zeros(1, 15, 4, 4)], [32, 32])
    with pytest.raises(N.NativeError):
        non_max_suppression(torch.zeros(1, 16, 15), 10, [32, 32], np.array([32, 32]), False)
    with pytest.raises(N.NativeError):
        batched_nms(torch.zeros(3, 4), torch.zeros(3), torch.zeros(3), 0.5)


def test_detection_lines_wire_format():
    """yolo.py:288-303: "<class> <score[:6]> <left> <top> <right> <bottom>" per detection row."""
    import io

    from glsdet_b200.utils_bbox import detection_lines

    names = ["pedestrian", "car", "van"]
    res = np.array([[10.7, 20.2, 110.9, 220.5, 0.9, 0.8123456, 1.0],
                    [-3.2, 5.0, 40.0, 50.0, 0.5, 0.1, 0.0],
                    [1.0, 2.0, 3.0, 4.0, 1.0, 1e-5, 2.0]], dtype=np.float32)
    # the reference's own statements (yolo.py:288-303), executed verbatim on the same rows
    f = io.StringIO()
    top_label = np.array(res[:, 6], dtype='int32')
    top_conf = res[:, 4] * res[:, 5]
    top_boxes = res[:, :4]
    for i, c in list(enumerate(top_label)):
        predicted_class = names[int(c)]
        box = top_boxes[i]
        score = str(top_conf[i])
        top, left, bottom, right = box
        f.write("%s %s %s %s %s %s\n" % (predicted_class, score[:6], str(int(left)), str(int(top)), str(int(right)), str(int(bottom))))
    assert detection_lines(res, names) == f.getvalue().splitlines()
    assert detection_lines(res, names)[0] == "car 0.7311 20 10 220 110"
    assert detection_lines(res, names, keep_classes=["car"]) == ["car 0.7311 20 10 220 110"]
    assert detection_lines(None, names) == []


def _folded_windows(x_nhwc16: torch.Tensor, step: int) -> torch.Tensor:
    """CPU model of the overlapping TMA view of a zero-bordered 16-channel row buffer: 64 consecutive elements starting
    at padded pixel step * X (rows padded with one zero pixel on each side, slack behind the last row)."""
    b, h, w, c = x_nhwc16.shape
    padded = F.pad(x_nhwc16, (0, 0, 1, 1))                      # [B, H, W + 2, 16]
    flat = torch.cat([padded.reshape(b, h, -1), torch.zeros(b, h, 64)], 2)
    n = w // step
    idx = (torch.arange(n) * step * c)[:, None] + torch.arange(64)[None, :]
    return flat[:, :, idx]                                      # [B, H, n, 64]


def test_kx_folded_weight_transforms_reproduce_the_conv():
    """fold_kx_weight / fold_kx_pair_weight (Focus stem: kx taps folded into an overlapping channel view) and
    pair_stride2_weight (stride-2 conv over pixel pairs) against F.conv2d, with the views modelled on the CPU."""
    from glsdet_b200.ops import fold_kx_pair_weight, fold_kx_weight, pair_stride2_weight

    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 12, 6, 8, generator=g)
    w = torch.randn(5, 12, 3, 3, generator=g)
    bias = torch.randn(5, generator=g)
    ref = F.conv2d(x, w, bias, padding=1)
    x16 = F.pad(x.permute(0, 2, 3, 1), (0, 4))                  # [B, H, W, 16]
    # single form: K = (ky, 64-element window at pixel x - 1)
    win = F.pad(_folded_windows(x16, 1), (0, 0, 0, 0, 1, 1))    # rows padded for the ky taps: [B, H + 2, W, 64]
    wf = fold_kx_weight(w, 16)                                  # [N, 64, 3, 1]
    out = sum(torch.einsum("bhwk,nk->bnhw", win[:, ky:ky + 6], wf[:, :, ky, 0]) for ky in range(3)) + bias.view(1, -1, 1, 1)
    assert torch.allclose(out, ref, atol=1e-4)
    # pair form: one row of the GEMM = output pixels (2X, 2X + 1)
    winp = F.pad(_folded_windows(x16, 2), (0, 0, 0, 0, 1, 1))   # [B, H + 2, W / 2, 64]
    wp, bp = fold_kx_pair_weight(w, bias, 16)                   # [2N, 64, 3, 1], [2N]
    outp = sum(torch.einsum("bhwk,nk->bnhw", winp[:, ky:ky + 6], wp[:, :, ky, 0]) for ky in range(3)) + bp.view(1, -1, 1, 1)
    outp = outp.view(2, 2, 5, 6, 4).permute(0, 2, 3, 4, 1).reshape(2, 5, 6, 8)     # (half, n) -> pixels 2X + half
    assert torch.allclose(outp, ref, atol=1e-4)
    # stride-2 conv over pixel pairs: input [B, H, W/2, 2C], taps (ky, pair x - 1 | pair x), rows strided
    x2 = torch.randn(2, 32, 8, 12, generator=g)
    w2 = torch.randn(7, 32, 3, 3, generator=g)
    ref2 = F.conv2d(x2, w2, None, stride=2, padding=1)
    pairs = x2.permute(0, 2, 3, 1).reshape(2, 8, 6, 64)
    pp = F.pad(pairs, (0, 0, 1, 0, 1, 1))                       # one zero pair on the left, one zero row above / below
    ws = pair_stride2_weight(w2)                                # [N, 64, 3, 2]
    out2 = torch.zeros_like(ref2)
    for ky in range(3):
        rows = pp[:, ky:ky + 8:2]                               # input row 2 * oy + ky - 1
        for kx in range(2):
            out2 += torch.einsum("bhwk,nk->bnhw", rows[:, :, kx:kx + 6], ws[:, :, ky, kx])
    assert torch.allclose(out2, ref2, atol=1e-4)


def test_pil_bicubic_tables_reproduce_pillow(native_lib):
    """glsdet_pil_bicubic_table is a HOST function of the C ABI: its windows / 22-bit weights, applied with Pillow's two
    passes in numpy, reproduce Image.resize(..., BICUBIC) byte for byte (the device kernels apply the same tables)."""
    import ctypes as C

    Image = pytest.importorskip("PIL.Image")
    from glsdet_b200 import _native as N

    def table(i, o):
        ks = native_lib.glsdet_pil_bicubic_ksize(i, o)
        b = np.zeros((o, 2), np.int32)
        k = np.zeros((o, ks), np.int32)
        N.check(native_lib.glsdet_pil_bicubic_table(i, o, b.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p)))
        return b, k

    def resample(img, ow, oh):
        ih, iw, _ = img.shape
        cur = img.astype(np.int64)
        if ow != iw:
            b, k = table(iw, ow)
            out = np.zeros((ih, ow, 3), np.int64)
            for x in range(ow):
                lo, n = b[x]
                out[:, x, :] = (1 << 21) + (cur[:, lo:lo + n, :] * k[x, :n][None, :, None]).sum(1)
            cur = np.clip(out >> 22, 0, 255)
        if oh != ih:
            b, k = table(ih, oh)
            out = np.zeros((oh, cur.shape[1], 3), np.int64)
            for y in range(oh):
                lo, n = b[y]
                out[y] = (1 << 21) + (cur[lo:lo + n] * k[y, :n][:, None, None]).sum(0)
            cur = np.clip(out >> 22, 0, 255)
        return cur.astype(np.uint8)

    rng = np.random.default_rng(0)
    for ih, iw, oh, ow in ((153, 272, 256, 256), (100, 37, 64, 96), (50, 50, 50, 96), (33, 200, 128, 200), (17, 19, 160, 128)):
        img = rng.integers(0, 256, (ih, iw, 3), dtype=np.uint8)
        want = np.array(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(resample(img, ow, oh), want), (ih, iw, oh, ow)


def test_letterbox_geometry_follows_reference():
    from glsdet_b200.utils import letterbox_geometry

    assert letterbox_geometry(500, 300, 416, 416, True) == (416, 249, 0, 83)      # models/core/utils.py:24-33
    assert letterbox_geometry(333, 765, 640, 512, True) == (222, 511, 209, 0)      # int(765 * (512 / 765)) truncates to 511
    assert letterbox_geometry(500, 300, 416, 416, False) == (416, 416, 0, 0)
