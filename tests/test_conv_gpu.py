"""GPU parity of the tcgen05 implicit-GEMM conv against a plain fp32 PyTorch conv of the same op.

Inputs and weights are rounded to bf16 first so that both sides see identical operands; the remaining
difference is accumulation order (fp32) and the bf16 rounding of the stored output.
Tolerances: bf16 outputs 1e-2 relative to the tensor's max magnitude (bf16 has 8 mantissa bits -> 4e-3
per element), fp32 outputs 1e-4.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf16r(t):
    return t.to(torch.bfloat16).float()


def _nhwc(t):  # NCHW fp32 -> NHWC bf16 contiguous
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _act(x, name):
    if name == "silu":
        return x * torch.sigmoid(x)
    if name == "relu":
        return torch.relu(x)
    return x


CASES = [
    # name, B, H, W, cins, N, k, stride, act
    ("1x1_c64_n128", 2, 32, 32, [64], 128, 1, 1, "silu"),
    ("1x1_c512_n256", 1, 32, 32, [512], 256, 1, 1, "silu"),
    ("3x3_c128_n128_ragged", 2, 40, 40, [128], 128, 3, 1, "silu"),
    ("3x3_c64_n64", 1, 64, 64, [64], 64, 3, 1, "silu"),
    ("3x3s2_c128_n128", 2, 64, 64, [128], 128, 3, 2, "silu"),
    ("3x3s2_c256_n256_ragged", 1, 40, 40, [256], 256, 3, 2, "silu"),
    ("cat_1x1_128+128_n256", 2, 32, 32, [128, 128], 256, 1, 1, "relu"),
    ("1x1_c512_n512", 1, 32, 32, [512], 512, 1, 1, "relu"),
    ("1x1_c96_n192", 1, 24, 24, [96], 192, 1, 1, "none"),
    ("3x3_c128_n256", 1, 20, 20, [128], 256, 3, 1, "silu"),
    ("3x3_c128_n128_big", 2, 256, 256, [128], 128, 3, 1, "silu"),
    ("1x1_c64_n64_ragged", 2, 40, 36, [64], 64, 1, 1, "silu"),
    ("1x1_c128_n192_ragged", 1, 20, 44, [128], 192, 1, 1, "relu"),
    ("1x1_c64_n128_big", 2, 256, 256, [64], 128, 1, 1, "silu"),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_matches_torch(case, native_lib, cuda_device):
    from glsdet_b200 import _native as N
    from glsdet_b200.ops import ConvOp, View

    name, B, H, W, cins, n_out, k, stride, act = case
    g = torch.Generator(device="cpu").manual_seed(hash(name) % (2 ** 31))
    dev = cuda_device
    xs = [_bf16r(torch.randn(B, c, H, W, generator=g)).to(dev) for c in cins]
    cin = sum(cins)
    w = _bf16r(torch.randn(n_out, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev)
    bias = torch.randn(n_out, generator=g).to(dev)
    ref = F.conv2d(torch.cat(xs, 1), w, bias, stride=stride, padding=(k - 1) // 2)
    ref = _act(ref, act)

    srcs = [View(_nhwc(x)) for x in xs]
    Ho, Wo = H // stride, W // stride
    out = torch.full((B, Ho, Wo, n_out), float("nan"), device=dev, dtype=torch.bfloat16)
    op = ConvOp(srcs, w, bias, ksize=k, stride=stride, act=N.ACT_BY_NAME[act], out=View(out))
    op.launch()
    torch.cuda.synchronize()
    got = out.float().permute(0, 3, 1, 2)
    assert torch.isfinite(got).all(), "unwritten or non-finite outputs"
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 1e-2 * scale, f"{name}: max abs err {err} vs scale {scale}"


def test_conv_residuals_and_fp32_out(native_lib, cuda_device):
    """pre-activation fp32 low-res residual (upsampled), post-activation bf16 residual, fp32 NHWC output."""
    from glsdet_b200 import _native as N
    from glsdet_b200.ops import ConvOp, View

    dev = cuda_device
    g = torch.Generator().manual_seed(7)
    B, H, W, Cin, Nout = 2, 32, 32, 128, 128
    x = _bf16r(torch.randn(B, Cin, H, W, generator=g)).to(dev)
    w = _bf16r(torch.randn(Nout, Cin, 1, 1, generator=g) / Cin ** 0.5).to(dev)
    bias = torch.randn(Nout, generator=g).to(dev)
    pre = torch.randn(B, H // 2, W // 2, Nout, generator=g).to(dev)          # NHWC fp32, half resolution
    post = _bf16r(torch.randn(B, H, W, Nout, generator=g)).to(dev)            # NHWC bf16 values
    pre_up = F.interpolate(pre.permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    ref = F.conv2d(x, w, bias) + pre_up
    ref = ref * torch.sigmoid(ref) + post.permute(0, 3, 1, 2)

    out = torch.full((B, H, W, Nout), float("nan"), device=dev, dtype=torch.float32)
    op = ConvOp([View(_nhwc(x))], w, bias, ksize=1, act=N.ACT_SILU, out=View(out),
                pre_res=View(pre.contiguous()), pre_shift=1, post_res=View(post.to(torch.bfloat16).contiguous()),
                post_shift=0)
    op.launch()
    torch.cuda.synchronize()
    got = out.permute(0, 3, 1, 2)
    assert torch.isfinite(got).all()
    assert (got - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()


def test_conv_channel_windows(native_lib, cuda_device):
    """Read channels [64,128) of a 192-wide buffer, write channels [32,96) of a 128-wide buffer."""
    from glsdet_b200 import _native as N
    from glsdet_b200.ops import ConvOp, View

    dev = cuda_device
    g = torch.Generator().manual_seed(11)
    B, H, W = 1, 16, 48
    buf = _bf16r(torch.randn(B, H, W, 192, generator=g)).to(dev).to(torch.bfloat16)
    w = _bf16r(torch.randn(64, 64, 3, 3, generator=g) / 24.0).to(dev)
    out = torch.zeros((B, H, W, 128), device=dev, dtype=torch.bfloat16)
    op = ConvOp([View(buf, 64, 64)], w, None, ksize=3, act=N.ACT_RELU, out=View(out, 32, 64))
    op.launch()
    torch.cuda.synchronize()
    x = buf[..., 64:128].float().permute(0, 3, 1, 2)
    ref = torch.relu(F.conv2d(x, w, None, padding=1))
    got = out[..., 32:96].float().permute(0, 3, 1, 2)
    assert (got - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    assert out[..., :32].abs().max().item() == 0 and out[..., 96:].abs().max().item() == 0


def test_pred_conv_nchw_and_decode(native_lib, cuda_device):
    """Prediction convs: N=10 -> NCHW fp32 logits; N=5 with the fused YOLOX box decode into [B, A, 15] rows."""
    from glsdet_b200 import _native as N
    from glsdet_b200.ops import ConvOp, View

    dev = cuda_device
    g = torch.Generator().manual_seed(3)
    B, H, W, Cin, nc = 2, 24, 40, 128, 10
    x = _bf16r(torch.randn(B, Cin, H, W, generator=g)).to(dev)
    wc = _bf16r(torch.randn(nc, Cin, 1, 1, generator=g) / Cin ** 0.5).to(dev)
    bc = torch.randn(nc, generator=g).to(dev)
    wr = _bf16r(torch.randn(5, Cin, 1, 1, generator=g) / Cin ** 0.5).to(dev)
    br = torch.randn(5, generator=g).to(dev)
    xin = View(_nhwc(x))

    logits = torch.full((B, 5 + nc, H, W), float("nan"), device=dev)
    op = ConvOp([xin], wc, bc, ksize=1, act=N.ACT_NONE, out=logits, out_mode=N.OUT_NCHW_F32, out_ld=5 + nc,
                out_coff=5, out_batch_stride=(5 + nc) * H * W)
    op.launch()
    torch.cuda.synchronize()
    ref_c = F.conv2d(x, wc, bc)
    assert (logits[:, 5:] - ref_c).abs().max().item() <= 1e-4 * ref_c.abs().max().item()
    assert torch.isnan(logits[:, :5]).all()

    A = H * W + 7  # rows of this level sit at offset 3 inside a longer anchor list
    pred = torch.full((B, A, 5 + nc), float("nan"), device=dev)
    stride, in_h, in_w = 8.0, H * 8.0, W * 8.0
    base = pred[:, 3:, :]
    op2 = ConvOp([xin], wr, br, ksize=1, act=N.ACT_YOLOX_BOX, out=pred, out_mode=N.OUT_NHWC_F32, out_ld=5 + nc,
                 out_coff=3 * (5 + nc), out_batch_stride=A * (5 + nc), dec=(stride, in_w, in_h))
    op2.launch()
    op3 = ConvOp([xin], wc, bc, ksize=1, act=N.ACT_SIGMOID, out=pred, out_mode=N.OUT_NHWC_F32, out_ld=5 + nc,
                 out_coff=3 * (5 + nc) + 5, out_batch_stride=A * (5 + nc))
    op3.launch()
    torch.cuda.synchronize()
    r = F.conv2d(x, wr, br)  # [B,5,H,W]
    gy, gx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    cx = (r[:, 0] + gx) * stride / in_w
    cy = (r[:, 1] + gy) * stride / in_h
    bw = torch.exp(r[:, 2]) * stride / in_w
    bh = torch.exp(r[:, 3]) * stride / in_h
    ob = torch.sigmoid(r[:, 4])
    ref_rows = torch.stack([cx, cy, bw, bh, ob], dim=-1).reshape(B, H * W, 5)
    got_rows = base[:, :H * W, :5]
    assert torch.allclose(got_rows, ref_rows, rtol=2e-4, atol=2e-5)
    ref_cls = torch.sigmoid(ref_c).permute(0, 2, 3, 1).reshape(B, H * W, nc)
    assert torch.allclose(base[:, :H * W, 5:], ref_cls, rtol=2e-4, atol=2e-5)
    assert torch.isnan(pred[:, :3]).all() and torch.isnan(pred[:, 3 + H * W:]).all()


def test_generic_epilogue_fallback(native_lib, cuda_device, monkeypatch):
    """The runtime-dispatched generic epilogue (used for combinations without a specialisation) stays correct."""
    monkeypatch.setenv("GLSDET_CONV_GENERIC_EPILOGUE", "1")
    test_conv_residuals_and_fp32_out(native_lib, cuda_device)
    test_pred_conv_nchw_and_decode(native_lib, cuda_device)
    test_conv_matches_torch(CASES[2], native_lib, cuda_device)


def test_bf16_residual_epilogues(native_lib, cuda_device):
    """bf16 outputs with the pre-activation (fp32, half resolution) or post-activation (bf16, half resolution)
    residual: the specialised epilogues used by the CSP blocks in front of an upsample+concat and by head.csp."""
    from glsdet_b200 import _native as N
    from glsdet_b200.ops import ConvOp, View

    dev = cuda_device
    g = torch.Generator().manual_seed(21)
    B, H, W, Cin, Nout = 2, 40, 24, 64, 128
    x = _bf16r(torch.randn(B, Cin, H, W, generator=g)).to(dev)
    w = _bf16r(torch.randn(Nout, Cin, 1, 1, generator=g) / Cin ** 0.5).to(dev)
    bias = torch.randn(Nout, generator=g).to(dev)
    pre = torch.randn(B, H // 2, W // 2, Nout, generator=g).to(dev)
    post = _bf16r(torch.randn(B, H // 2, W // 2, Nout, generator=g)).to(dev)
    up = lambda t: F.interpolate(t.permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    conv = F.conv2d(x, w, bias)
    out = torch.zeros((B, H, W, Nout), device=dev, dtype=torch.bfloat16)
    ConvOp([View(_nhwc(x))], w, bias, ksize=1, act=N.ACT_SILU, out=View(out), pre_res=View(pre.contiguous()),
           pre_shift=1).launch()
    torch.cuda.synchronize()
    ref = conv + up(pre)
    ref = ref * torch.sigmoid(ref)
    assert (out.float().permute(0, 3, 1, 2) - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    out.zero_()
    ConvOp([View(_nhwc(x))], w, bias, ksize=1, act=N.ACT_RELU, out=View(out),
           post_res=View(post.to(torch.bfloat16).contiguous()), post_shift=1).launch()
    torch.cuda.synchronize()
    ref = torch.relu(conv) + up(post)
    assert (out.float().permute(0, 3, 1, 2) - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()


@pytest.mark.parametrize("n_tower,n_pred,pred_act", [(128, 10, "sigmoid"), (128, 5, "box"), (128, 10, "none_nchw"),
                                                     (256, 3, "sigmoid"), (64, 5, "none_nchw"), (96, 16, "none_rows")])
def test_fused_tower_pred_conv(n_tower, n_pred, pred_act, native_lib, cuda_device):
    """3x3 tower conv + SiLU + 1x1 prediction conv (+ decode) in one kernel vs the two-step PyTorch evaluation."""
    from glsdet_b200 import _native as N
    from glsdet_b200.ops import ConvOp, View

    dev = cuda_device
    g = torch.Generator().manual_seed(n_tower * 31 + n_pred)
    B, H, W, Cin = 2, 24, 40, 128
    x = _bf16r(torch.randn(B, Cin, H, W, generator=g)).to(dev)
    w = _bf16r(torch.randn(n_tower, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(dev)
    bias = torch.randn(n_tower, generator=g).to(dev)
    wp = (torch.randn(n_pred, n_tower, 1, 1, generator=g) / n_tower ** 0.5).to(dev)
    bp = torch.randn(n_pred, generator=g).to(dev)
    t = F.conv2d(x, w, bias, padding=1)
    t = t * torch.sigmoid(t)
    import os
    if n_tower % 64 == 0 and not os.environ.get("GLSDET_CONV_PRED_FMA"):
        # tensor-core prediction path: the activated tile and the prediction weights are bf16 MMA operands
        y = F.conv2d(_bf16r(t), _bf16r(wp), bp)
    else:
        y = F.conv2d(t, wp, bp)                                 # [B, n_pred, H, W]
    rows = y.permute(0, 2, 3, 1).reshape(B, H * W, n_pred)
    nch = n_pred + 3
    if pred_act == "none_nchw":
        out = torch.full((B, nch, H, W), float("nan"), device=dev)
        ConvOp([View(_nhwc(x))], w, bias, ksize=3, act=N.ACT_SILU, out=out, out_mode=N.OUT_NCHW_F32, out_ld=nch,
               out_coff=2, out_batch_stride=nch * H * W, pred_weight=wp, pred_bias=bp, pred_act=N.ACT_NONE).launch()
        torch.cuda.synchronize()
        # 6e-3: one bf16 rounding flip of a large activation (|t| 4 * 2^-8) times a prediction weight
        assert torch.allclose(out[:, 2:2 + n_pred], y, rtol=6e-3, atol=6e-3), (out[:, 2:2 + n_pred] - y).abs().max()
        assert torch.isnan(out[:, :2]).all() and torch.isnan(out[:, 2 + n_pred:]).all()
        return
    out = torch.full((B, H * W, nch), float("nan"), device=dev)
    act = {"sigmoid": N.ACT_SIGMOID, "box": N.ACT_YOLOX_BOX, "none_rows": N.ACT_NONE}[pred_act]
    stride, in_h, in_w = 8.0, H * 8.0, W * 8.0
    ConvOp([View(_nhwc(x))], w, bias, ksize=3, act=N.ACT_SILU, out=out, out_mode=N.OUT_NHWC_F32, out_ld=nch,
           out_coff=1, out_batch_stride=H * W * nch, pred_weight=wp, pred_bias=bp, pred_act=act,
           dec=(stride, in_w, in_h)).launch()
    torch.cuda.synchronize()
    got = out[:, :, 1:1 + n_pred]
    if pred_act == "sigmoid":
        ref = torch.sigmoid(rows)
    elif pred_act == "none_rows":
        ref = rows
    else:
        gy, gx = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
        gx, gy = gx.reshape(1, -1).float(), gy.reshape(1, -1).float()
        ref = torch.stack([(rows[..., 0] + gx) * stride / in_w, (rows[..., 1] + gy) * stride / in_h,
                           torch.exp(rows[..., 2]) * stride / in_w, torch.exp(rows[..., 3]) * stride / in_h,
                           torch.sigmoid(rows[..., 4])], dim=-1)
    assert torch.allclose(got, ref, rtol=6e-3, atol=6e-3), (got - ref).abs().max()
    assert torch.isnan(out[:, :, 0]).all() and torch.isnan(out[:, :, 1 + n_pred:]).all()


def test_two_cta_mode(native_lib, cuda_device, monkeypatch):
    """cta_group::2 variant (CTA pairs, M = 256 per tcgen05.mma, operands split over two SMs) forced for every stride-1
    conv including the 1x1 ones (it is the default for 3x3 only)."""
    monkeypatch.setenv("GLSDET_CONV_2CTA", "1")
    monkeypatch.setenv("GLSDET_CONV_PRED_FMA", "1")   # the 2-CTA kernel keeps the FMA prediction path
    for case in (CASES[2], CASES[3], CASES[6], CASES[9], CASES[10]):
        test_conv_matches_torch(case, native_lib, cuda_device)
    test_fused_tower_pred_conv(128, 10, "sigmoid", native_lib, cuda_device)
    test_bf16_residual_epilogues(native_lib, cuda_device)


def test_fused_pred_fma_path(native_lib, cuda_device, monkeypatch):
    """Per-thread FMA prediction path on the 1-CTA kernel (tower widths that are not a multiple of 64 use it; forced)."""
    monkeypatch.setenv("GLSDET_CONV_2CTA", "0")
    monkeypatch.setenv("GLSDET_CONV_PRED_FMA", "1")
    test_fused_tower_pred_conv(128, 10, "sigmoid", native_lib, cuda_device)
    test_fused_tower_pred_conv(128, 5, "box", native_lib, cuda_device)
    test_fused_tower_pred_conv(256, 3, "sigmoid", native_lib, cuda_device)


@pytest.mark.parametrize("env", [{"GLSDET_CONV_2CTA": "0"},
                                 {"GLSDET_CONV_2CTA": "0", "GLSDET_CONV_MT": "2"},
                                 {"GLSDET_CONV_2CTA": "0", "GLSDET_CONV_MT": "2", "GLSDET_CONV_NO_BRES": "1"},
                                 {"GLSDET_CONV_2CTA": "0", "GLSDET_CONV_MT": "1", "GLSDET_CONV_NO_BRES": "1", "GLSDET_CONV_NO_PDL": "1"},
                                 {"GLSDET_CONV_2CTA": "0", "GLSDET_CONV_MT": "2", "GLSDET_CONV_NO_BGROUP": "1"},
                                 {"GLSDET_CONV_NO_TMA_STORE": "1"}, {"GLSDET_CONV_2CTA": "0", "GLSDET_CONV_EPI8": "1"}],
                         ids=["one_cta", "mt2", "mt2_nobres", "mt1_nobres_nopdl", "mt2_nobgroup", "direct_stores", "epi8"])
def test_conv_schedule_variants(env, native_lib, cuda_device, monkeypatch):
    """Work-item shapes and weight staging are scheduling choices: two M tiles per weight stage (mt=2), resident
    weights, grouped ky taps, programmatic dependent launch - every combination must give the same results."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for case in CASES:
        test_conv_matches_torch(case, native_lib, cuda_device)
    test_conv_residuals_and_fp32_out(native_lib, cuda_device)
    test_bf16_residual_epilogues(native_lib, cuda_device)
    for args in [(128, 10, "sigmoid"), (128, 5, "box"), (256, 3, "sigmoid"), (64, 5, "none_nchw"), (96, 16, "none_rows")]:
        test_fused_tower_pred_conv(*args, native_lib, cuda_device)
