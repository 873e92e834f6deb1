"""Python handles for the native operators (thin: build a descriptor, keep tensors alive, launch)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _native as N

CHUNK_K = 64


def fold_bn(conv_weight: torch.Tensor, bn_weight, bn_bias, bn_mean, bn_var, eps: float):
    """Fold eval-mode BatchNorm into the conv (reference: BaseConv.forward, models/base/baseConv.py:15-16,
    BatchNorm2d(eps=1e-3) at :12):  y = (conv(x) - mean) / sqrt(var + eps) * gamma + beta."""
    scale = bn_weight.double() / torch.sqrt(bn_var.double() + eps)
    w = conv_weight.double() * scale.view(-1, 1, 1, 1)
    b = bn_bias.double() - bn_mean.double() * scale
    return w.float(), b.float()


def pack_conv_weight(weight: torch.Tensor, splits: Sequence[int], n_pad: int, k_pad: int,
                     dtype=torch.bfloat16) -> torch.Tensor:
    """[N, Cin, kh, kw] fp32 -> 16-bit [n_pad, k_pad] (the storage type of the conv's sources) in the K order the
    kernel streams: (source, tap = ky*kw + kx, channel) with every (source, tap) segment zero-padded to 64 channels."""
    n, cin, kh, kw = weight.shape
    assert sum(splits) == cin, (splits, cin)
    segs = []
    a = 0
    for c in splits:
        seg = weight[:, a:a + c].permute(0, 2, 3, 1)  # [N, kh, kw, c]
        cpad = (c + CHUNK_K - 1) // CHUNK_K * CHUNK_K
        if cpad != c:
            seg = torch.nn.functional.pad(seg, (0, cpad - c))
        segs.append(seg.reshape(n, kh * kw * cpad))
        a += c
    packed = torch.cat(segs, dim=1)
    assert packed.shape[1] == k_pad, (packed.shape, k_pad)
    out = torch.zeros((n_pad, k_pad), dtype=dtype, device=weight.device)
    if dtype == torch.float16:   # saturate like the kernels' fp16 stores do
        packed = packed.clamp(-65504.0, 65504.0)
    out[:n] = packed.to(dtype)
    return out.contiguous()


class View:
    """A channel window of an NHWC buffer: tensor [B, H, W, ld] (bf16 / fp16 or fp32), channels [coff, coff + c)."""

    __slots__ = ("t", "coff", "c")

    def __init__(self, t: torch.Tensor, coff: int = 0, c: Optional[int] = None):
        assert t.dim() == 4 and t.is_contiguous()
        self.t, self.coff = t, coff
        self.c = t.shape[3] - coff if c is None else c
        assert 0 <= coff and coff + self.c <= t.shape[3]

    @property
    def ld(self) -> int:
        return self.t.shape[3]

    @property
    def ptr(self) -> int:
        return self.t.data_ptr() + self.coff * self.t.element_size()

    @property
    def bhw(self):
        return tuple(self.t.shape[:3])


def as16(v: "View") -> "View":
    """An fp32 channel window seen as a 16-bit one with twice the channels (bit copies only: rectangle copies, nearest
    upsampling - the kernels move 16-byte vectors and never interpret the elements); 16-bit windows pass through."""
    if v.t.dtype != torch.float32:
        return v
    return View(v.t.view(torch.float16), 2 * v.coff, 2 * v.c)


class FoldedView:
    """Source view of a few-channel 3x3 conv with the kx taps folded into the channel dimension (glsdet_conv_desc.ksize_w
    = 1): `t` is a zero-bordered NHWC bf16 buffer [B, H, W + 2, c_pix] (+ slack behind the last row); the conv sees, at
    output pixel x, the 64 consecutive elements starting at padded pixel x = (pixel x-1 | pixel x | pixel x+1 | ...)."""

    __slots__ = ("t", "c", "c_pix", "_bhw", "_row_pitch")

    def __init__(self, flat: torch.Tensor, batch: int, height: int, width: int, c_pix: int, row_pitch: int = 0):
        """`row_pitch` (elements) overrides (width + 2) * c_pix: the PAIR view of a 16-channel buffer - c_pix = 32 (two
        pixels per step), width = W / 2, row_pitch = (W + 2) * 16 - whose 64-element window at step X covers pixels
        2X-1 .. 2X+2, i.e. the inputs of the two output pixels 2X and 2X+1 (see fold_kx_pair_weight)."""
        assert flat.dtype in (torch.bfloat16, torch.float16) and flat.dim() == 1 and c_pix % 8 == 0 and c_pix <= 64
        self._row_pitch = row_pitch if row_pitch else (width + 2) * c_pix
        assert flat.numel() >= batch * height * self._row_pitch + 64, "needs 64 elements of slack behind the last row"
        self.t, self.c, self.c_pix, self._bhw = flat, 64, c_pix, (batch, height, width)

    @property
    def ld(self) -> int:
        return self.c_pix

    @property
    def ptr(self) -> int:
        return self.t.data_ptr()

    @property
    def bhw(self):
        return self._bhw

    @property
    def row_pitch(self) -> int:
        return self._row_pitch

    @property
    def img_pitch(self) -> int:
        return self._bhw[1] * self.row_pitch

    def interior(self) -> torch.Tensor:
        """[B, H, W, c_pix] view of the real pixels (tests)."""
        b, h, w = self._bhw
        return self.t[:b * h * (w + 2) * self.c_pix].view(b, h, w + 2, self.c_pix)[:, :, 1:w + 1]


def fold_kx_weight(weight: torch.Tensor, c_pix: int) -> torch.Tensor:
    """[N, C, 3, 3] -> [N, 64, 3, 1] for a FoldedView source: channel kx * c_pix + c of tap ky = weight[:, c, ky, kx]."""
    n, c, kh, kw = weight.shape
    assert kw == 3 and c <= c_pix and 3 * c_pix <= 64
    out = torch.zeros((n, 64, kh, 1), dtype=weight.dtype, device=weight.device)
    for kx in range(3):
        out[:, kx * c_pix:kx * c_pix + c, :, 0] = weight[:, :, :, kx]
    return out


def fold_kx_pair_weight(weight: torch.Tensor, bias: torch.Tensor, c_pix: int = 16):
    """[N, C, 3, 3] (C <= c_pix = 16) -> ([2N, 64, 3, 1], [2N]) for the PAIR view: one GEMM row = two horizontally adjacent
    output pixels.  The 64-element window holds pixel slots s = 0..3 = input pixels 2X-1 .. 2X+2; output half h (pixel
    2X + h) uses slot s with tap kx = s - h.  No zero padding of K is left (3 x 64 against 9 x 64 for the plain form) and
    N = 2N output channels are the two pixels' channels side by side - a contiguous [.., W/2, 2N] view of the output."""
    n, c, kh, kw = weight.shape
    assert kw == 3 and c <= c_pix and 4 * c_pix == 64
    out = torch.zeros((2 * n, 64, kh, 1), dtype=weight.dtype, device=weight.device)
    for h in range(2):
        for kx in range(3):
            s_ = kx + h
            out[h * n:(h + 1) * n, s_ * c_pix:s_ * c_pix + c, :, 0] = weight[:, :, :, kx]
    return out, torch.cat([bias, bias], 0)


def pair_stride2_weight(weight: torch.Tensor) -> torch.Tensor:
    """[N, C, 3, 3] of a stride-2 conv -> [N, 2C, 3, 2] for the pixel-pair form (glsdet_conv_desc.ksize_w = 2): the input is
    viewed as [B, H, W/2, 2C] (pair X = pixels 2X, 2X+1); output column x reads pixels 2x-1, 2x, 2x+1 = the right pixel of
    pair x-1 (tap 0: kx = 0) and both pixels of pair x (tap 1: kx = 1, 2)."""
    n, c, kh, kw = weight.shape
    assert kh == 3 and kw == 3
    out = torch.zeros((n, 2 * c, 3, 2), dtype=weight.dtype, device=weight.device)
    out[:, c:, :, 0] = weight[:, :, :, 0]
    out[:, :c, :, 1] = weight[:, :, :, 1]
    out[:, c:, :, 1] = weight[:, :, :, 2]
    return out


def pair_pointwise_weight(weight: torch.Tensor, in_groups: int = 1, out_groups: int = 1) -> torch.Tensor:
    """[N, C, 1, 1] of a 1x1 conv -> [2N, 2C, 1, 1] for the pixel-pair form of a stride-1 layer: the tensors are viewed as
    [B, H, W/2, 2C] (pair X = pixels 2X, 2X+1) and both pixels of a pair go through the same weights.  The channels of a
    pair are ordered (group, pixel, channel within the group): with one group that is (pixel 0 | pixel 1), with two groups
    of C/2 channels (x1(p0) | x1(p1) | x2(p0) | x2(p1)) - the layout that keeps each half of a CSPLayer's concat buffer
    one contiguous window of the pair view."""
    n, c = weight.shape[:2]
    assert weight.shape[2:] == (1, 1) and c % in_groups == 0 and n % out_groups == 0
    w = weight[:, :, 0, 0]
    gi, go = c // in_groups, n // out_groups
    out = torch.zeros((2 * n, 2 * c), dtype=weight.dtype, device=weight.device)
    for a in range(out_groups):
        for b in range(in_groups):
            blk = w[a * go:(a + 1) * go, b * gi:(b + 1) * gi]
            for px in range(2):
                out[(2 * a + px) * go:(2 * a + px + 1) * go, (2 * b + px) * gi:(2 * b + px + 1) * gi] = blk
    return out[:, :, None, None].contiguous()


def pair_bias(bias: torch.Tensor, groups: int = 1) -> torch.Tensor:
    """Bias of a pair-form layer in the (group, pixel, channel) order of pair_pointwise_weight."""
    g = bias.shape[0] // groups
    return torch.cat([bias[a * g:(a + 1) * g] for a in range(groups) for _ in range(2)])


def pair_conv3_weight(weight: torch.Tensor) -> torch.Tensor:
    """[N, C, 3, 3] of a stride-1 3x3 conv -> [2N, 2C, 3, 3] for the pixel-pair form: output pixel 2X + po reads input pixels
    2X + po - 1 .. 2X + po + 1; input pixel 2(X + kx' - 1) + pi is tap kx = 2(kx' - 1) + pi - po + 1 of it (when 0 <= kx <= 2).
    The zero padding of the pair view (pairs -1 and W/2) is the zero padding of the pixels -1 and W."""
    n, c, kh, kw = weight.shape
    assert kh == 3 and kw == 3
    out = torch.zeros((2 * n, 2 * c, 3, 3), dtype=weight.dtype, device=weight.device)
    for po in range(2):
        for pi in range(2):
            for kxp in range(3):
                kx = 2 * (kxp - 1) + pi - po + 1
                if 0 <= kx <= 2:
                    out[po * n:(po + 1) * n, pi * c:(pi + 1) * c, :, kxp] = weight[:, :, :, kx]
    return out


class ConvOp:
    """conv(+cat)(+bias)(+residual) -> act (+residual) -> store; see glsdet_conv_desc in include/glsdet_b200.h."""

    def __init__(self, srcs: Sequence[View], weight: torch.Tensor, bias: Optional[torch.Tensor], *, ksize: int,
                 stride: int = 1, act: int = N.ACT_NONE, out, out_mode: int = N.OUT_NHWC_BF16, out_ld: int = 0,
                 out_coff: int = 0, out_batch_stride: int = 0, pre_res: Optional[View] = None, pre_shift: int = 0,
                 post_res: Optional[View] = None, post_shift: int = 0, dec=(0.0, 0.0, 0.0),
                 pred_weight: Optional[torch.Tensor] = None, pred_bias: Optional[torch.Tensor] = None,
                 pred_act: int = N.ACT_NONE, weight_raw: Optional[torch.Tensor] = None, n_out: Optional[int] = None,
                 src_shared: int = 0, src_shared_div: int = 0, patch_mode: bool = False, batch: Optional[int] = None,
                 ksize_w: int = 0, out_plane_stride: int = 0, out_elem_offset: int = 0):
        """`weight_raw`: a bf16 device matrix [N rows, pitch] (shared) or [batch, N rows, pitch] (one per image) used
        as is (the batched products of the non-local block); `src_shared` = k > 0: srcs[0] holds k static matrices and image b reads matrix b mod k;
        `patch_mode`: srcs[0] / post_res / out are [B, H, W, .] tensors processed as their 4*B 2x2 patches."""
        lib = N.load()
        assert 1 <= len(srcs) <= 2
        b, h, w = srcs[0].bhw
        sdt = srcs[0].t.dtype
        for s in srcs:
            assert s.bhw == (b, h, w) and s.t.dtype == sdt and sdt in (torch.bfloat16, torch.float16)
        if patch_mode:
            assert h % 2 == 0 and w % 2 == 0 and ksize == 1
            b, h, w = 4 * b, h // 2, w // 2
        if src_shared:
            assert b == src_shared and batch is not None
            b = batch
        n_out = weight.shape[0] if weight_raw is None else n_out
        d = N.ConvDesc()
        d.src0, d.src0_c, d.src0_ld = srcs[0].ptr, srcs[0].c, srcs[0].ld
        if len(srcs) == 2:
            d.src1, d.src1_c, d.src1_ld = srcs[1].ptr, srcs[1].c, srcs[1].ld
        d.batch, d.height, d.width = b, h, w
        d.ksize, d.stride = ksize, stride
        d.ksize_w = ksize_w
        if isinstance(srcs[0], FoldedView):
            assert ksize_w == 1 and len(srcs) == 1 and stride == 1
            d.src0_row_pitch, d.src0_img_pitch = srcs[0].row_pitch, srcs[0].img_pitch
        d.out_channels = n_out
        d.act = act
        n_pad, k_pad, block_n = C.c_int32(), C.c_int32(), C.c_int32()
        N.check(lib.glsdet_conv_weight_shape(C.byref(d), C.byref(n_pad), C.byref(k_pad), C.byref(block_n)),
                "glsdet_conv_weight_shape")
        self.block_n = block_n.value
        dev = srcs[0].t.device
        if weight_raw is None:
            # weights are folded / packed wherever they live (the plans keep them on the HOST: a plan build is then a few
            # hundred small CPU ops and one upload per layer instead of ~1000 framework kernels) and uploaded here
            self.packed = pack_conv_weight(weight, [s.c for s in srcs], n_pad.value, k_pad.value, sdt).to(dev)
        else:
            assert weight_raw.dtype == sdt and weight_raw.is_contiguous() and ksize == 1
            assert weight_raw.shape[-1] >= k_pad.value and weight_raw.shape[-2] >= n_out, (weight_raw.shape, k_pad.value)
            self.packed = weight_raw
            d.weight_ld = weight_raw.shape[-1]
            if weight_raw.dim() == 3:
                assert weight_raw.shape[0] == b, (weight_raw.shape, b)
                d.weight_batch_stride = weight_raw.shape[1] * weight_raw.shape[2]
        d.src_shared = int(src_shared)
        d.src_shared_div = int(src_shared_div)
        d.patch_mode = 1 if patch_mode else 0
        d.src_dtype = N.dt_code(sdt)
        self.bias = None if bias is None else bias.detach().float().contiguous().to(dev)
        d.weight = self.packed.data_ptr()
        d.bias = 0 if self.bias is None else self.bias.data_ptr()
        ho, wo = h // stride, w // stride
        if stride == 2 and ksize_w == 2:   # stride-2 conv over pixel pairs: the width is already halved by the view
            wo = w
        if pre_res is not None:
            assert pre_res.t.dtype == torch.float32
            assert pre_res.bhw == (b, max(ho >> pre_shift, 1), max(wo >> pre_shift, 1)), (pre_res.bhw, b, ho, wo, pre_shift)
            d.pre_res, d.pre_shift, d.pre_ld = pre_res.ptr, pre_shift, pre_res.ld
        if post_res is not None:
            d.post_dtype = N.dt_code(post_res.t.dtype)
            if patch_mode:
                assert post_res.bhw == (b // 4, 2 * ho, 2 * wo) and post_shift == 0
            else:
                assert post_res.bhw == (b, ho >> post_shift, wo >> post_shift)
            d.post_res, d.post_shift, d.post_ld = post_res.ptr, post_shift, post_res.ld
        if isinstance(out, View):
            if patch_mode:
                assert out.bhw == (b // 4, 2 * ho, 2 * wo) and out.c >= n_out
            else:
                assert out.bhw == (b, ho, wo) and (out.c >= n_out or pred_weight is not None)
            d.out, d.out_ld, d.out_coff = out.t.data_ptr(), out.ld, out.coff
            d.out_batch_stride = ho * wo * out.ld
            if out.t.dtype == torch.float32:
                d.out_mode = N.OUT_NHWC_F32
            else:
                d.out_mode, d.out_dtype = N.OUT_NHWC_BF16, N.dt_code(out.t.dtype)
            self._out_t = out.t
        else:  # raw tensor with explicit addressing (NCHW fp32 logits / [B, A, C] decoded rows)
            d.out, d.out_ld, d.out_coff = out.data_ptr() + out_elem_offset * out.element_size(), out_ld, out_coff
            d.out_batch_stride = out_batch_stride
            d.out_mode = out_mode
            d.out_plane_stride = out_plane_stride
            self._out_t = out
        d.dec_stride, d.dec_in_w, d.dec_in_h = dec
        self.pred_weight = self.pred_bias = None
        if pred_weight is not None:   # fused prediction conv on the activated tile (never stored)
            self.pred_weight = pred_weight.detach().float().reshape(pred_weight.shape[0], -1).contiguous().to(dev)
            assert self.pred_weight.shape[1] == n_out and self.pred_weight.shape[0] <= 16
            self.pred_bias = pred_bias.detach().float().contiguous().to(dev)
            d.pred_weight, d.pred_bias = self.pred_weight.data_ptr(), self.pred_bias.data_ptr()
            d.pred_channels, d.pred_act = self.pred_weight.shape[0], pred_act
        self._keep = (srcs, pre_res, post_res, weight_raw)
        self.desc = d
        self.handle = C.c_void_p()
        N.check(lib.glsdet_conv_create(C.byref(d), C.byref(self.handle)), "glsdet_conv_create")
        self._lib = lib
        self.flops = 2.0 * b * ho * wo * n_out * sum(s.c for s in srcs) * ksize * (ksize_w or ksize)

    def launch(self, stream=None):
        N.check(self._lib.glsdet_conv_launch(self.handle, N.stream_ptr(stream)), "glsdet_conv_launch")

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self._lib.glsdet_conv_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------- other operators
def nchw_to_nhwc(src: torch.Tensor, dst: View, stream=None) -> None:
    """NCHW fp32 (reference layout) -> channel window of an NHWC bf16 buffer."""
    assert src.dtype == torch.float32 and src.is_contiguous() and src.is_cuda
    b, c, h, w = src.shape
    lib = N.load()
    if dst.t.dtype == torch.float32:   # fp32 accuracy mode
        assert dst.bhw == (b, h, w) and dst.c == c
        N.check(lib.glsdet_nchw_nhwc_f32(src.data_ptr(), dst.t.data_ptr(), b, c, h, w, dst.ld, dst.coff, 1,
                                         N.stream_ptr(stream)), "glsdet_nchw_nhwc_f32")
        return
    assert dst.bhw == (b, h, w) and dst.c == c
    N.check(lib.glsdet_nchw_f32_to_nhwc_16(src.data_ptr(), dst.t.data_ptr(), b, c, h, w, dst.ld, dst.coff,
                                           N.dt_code(dst.t.dtype), N.stream_ptr(stream)), "glsdet_nchw_f32_to_nhwc_16")


def nhwc_to_nchw(src: View, dst: torch.Tensor, stream=None) -> None:
    assert dst.dtype == torch.float32 and dst.is_contiguous()
    b, h, w = src.bhw
    assert tuple(dst.shape) == (b, src.c, h, w)
    lib = N.load()
    if src.t.dtype == torch.float32:
        N.check(lib.glsdet_nchw_nhwc_f32(src.t.data_ptr(), dst.data_ptr(), b, src.c, h, w, src.ld, src.coff, 0,
                                         N.stream_ptr(stream)), "glsdet_nchw_nhwc_f32")
        return
    N.check(lib.glsdet_nhwc_16_to_nchw_f32(src.t.data_ptr(), dst.data_ptr(), b, src.c, h, w, src.ld, src.coff,
                                           N.dt_code(src.t.dtype), N.stream_ptr(stream)), "glsdet_nhwc_16_to_nchw_f32")


class SeGateOp:
    """gate[b, c] = 1 + sigmoid(W2 relu(W1 mean_hw(x)))  (models/ffa/ffa.py:16-20 and :77)."""

    def __init__(self, x: View, w1: torch.Tensor, w2: torch.Tensor):
        assert x.coff == 0 and x.t.dtype in (torch.bfloat16, torch.float16, torch.float32)
        b, h, w = x.bhw
        self.x, self.b, self.hw, self.c = x, b, h * w, x.c
        self.w1 = w1.detach().float().contiguous().to(x.t.device)
        self.w2 = w2.detach().float().contiguous().to(x.t.device)
        self.hidden = self.w1.shape[0]
        assert tuple(self.w1.shape) == (self.hidden, self.c) and tuple(self.w2.shape) == (self.c, self.hidden)
        self.scratch = torch.empty((b, N.SE_SLABS, self.c), dtype=torch.float32, device=x.t.device)
        self.gate = torch.empty((b, self.c), dtype=torch.float32, device=x.t.device)
        self._lib = N.load()

    def launch(self, stream=None):
        if self.x.t.dtype == torch.float32:
            st = N.stream_ptr(stream)
            N.check(self._lib.glsdet_se_partial_f32(self.x.ptr, self.b, self.hw, self.c, self.x.ld,
                                                    self.scratch.data_ptr(), st), "glsdet_se_partial_f32")
            N.check(self._lib.glsdet_se_fc(self.scratch.data_ptr(), self.w1.data_ptr(), self.w2.data_ptr(), self.hidden,
                                           self.gate.data_ptr(), self.b, self.hw, self.c, st), "glsdet_se_fc")
            return
        N.check(self._lib.glsdet_se_gate_16(self.x.ptr, self.b, self.hw, self.c, self.x.ld, self.w1.data_ptr(),
                                            self.w2.data_ptr(), self.hidden, self.scratch.data_ptr(),
                                            self.gate.data_ptr(), N.dt_code(self.x.t.dtype), N.stream_ptr(stream)),
                "glsdet_se_gate_16")


class ScaleShuffleOp:
    """dst = PixelShuffle(2)(x * gate) with x's channels in (i, j, c) order (ffa.py:77-78)."""

    def __init__(self, x: View, gate: torch.Tensor, dst: View):
        b, h, w = x.bhw
        assert x.coff == 0 and x.c == x.ld and x.c % 4 == 0
        self.cout = x.c // 4
        assert dst.bhw == (b, 2 * h, 2 * w) and dst.c == self.cout
        self.x, self.gate, self.dst, self.b, self.h, self.w = x, gate, dst, b, h, w
        self._lib = N.load()

    def launch(self, stream=None):
        if self.x.t.dtype == torch.float32:
            N.check(self._lib.glsdet_scale_pixel_shuffle_f32(self.x.ptr, self.gate.data_ptr(), self.dst.t.data_ptr(),
                                                             self.b, self.h, self.w, self.cout, self.dst.ld,
                                                             self.dst.coff, N.stream_ptr(stream)),
                    "glsdet_scale_pixel_shuffle_f32")
            return
        N.check(self._lib.glsdet_scale_pixel_shuffle_16(self.x.ptr, self.gate.data_ptr(), self.dst.t.data_ptr(), self.b,
                                                        self.h, self.w, self.cout, self.dst.ld, self.dst.coff,
                                                        N.dt_code(self.x.t.dtype), N.dt_code(self.dst.t.dtype),
                                                        N.stream_ptr(stream)), "glsdet_scale_pixel_shuffle_16")


class _Call:
    """Deferred plain function call so that converters can sit in an op list next to the op objects."""

    def __init__(self, fn, *args):
        self.fn, self.args = fn, args

    def launch(self, stream=None):
        self.fn(*self.args, stream=stream)


class PatchTransposeOp:
    """NCHW fp32 map -> per-patch transposed bf16 matrices [4B, rows, t_ld] (Gram operand of the non-local block)."""

    def __init__(self, dst: torch.Tensor, channels: int, height: int, width: int, scale: float = 1.0):
        assert dst.dtype in (torch.bfloat16, torch.float16) and dst.dim() == 3 and dst.is_contiguous()
        self.dst, self.c, self.h, self.w, self.scale = dst, channels, height, width, float(scale)
        self._lib = N.load()

    def launch(self, src: torch.Tensor, stream=None):
        assert src.dtype == torch.float32 and src.is_contiguous() and src.is_cuda
        b = src.shape[0]
        assert tuple(src.shape) == (b, self.c, self.h, self.w) and self.dst.shape[0] == 4 * b
        N.check(self._lib.glsdet_patch_transpose_16(src.data_ptr(), self.dst.data_ptr(), b, self.c, self.h, self.w,
                                                    self.dst.shape[1], self.dst.shape[2], N.dt_code(self.dst.dtype),
                                                    self.scale, N.stream_ptr(stream)), "glsdet_patch_transpose_16")


class GatherBiasOp:
    """bias[b][n] = base[b mod groups][n] + w[b][n][col]."""

    def __init__(self, w: torch.Tensor, base: torch.Tensor, bias: torch.Tensor, col: int):
        assert w.dtype in (torch.bfloat16, torch.float16) and w.dim() == 3 and w.is_contiguous()
        assert base.dtype == torch.float32 and bias.dtype == torch.float32 and base.is_contiguous()
        self.w, self.base, self.bias, self.col = w, base, bias, col
        self.groups = base.shape[0]
        assert base.shape[1] == w.shape[1] and bias.numel() == w.shape[0] * w.shape[1]
        self._lib = N.load()

    def launch(self, stream=None):
        w = self.w
        N.check(self._lib.glsdet_gather_bias_16(w.data_ptr(), self.base.data_ptr(), self.bias.data_ptr(), w.shape[0],
                                                w.shape[1], w.shape[2], self.col, w.shape[1] * w.shape[2], self.groups,
                                                N.dt_code(w.dtype), N.stream_ptr(stream)), "glsdet_gather_bias_16")


class Upsample2xOp:
    """dst window = nearest 2x upsampling of the src window (NHWC bf16)."""

    def __init__(self, src: View, dst: View):
        src, dst = as16(src), as16(dst)   # fp32 accuracy mode: copied as 16-bit pairs
        b, h, w = src.bhw
        assert dst.bhw == (b, 2 * h, 2 * w) and dst.c == src.c
        assert src.t.dtype == dst.t.dtype and src.t.dtype in (torch.bfloat16, torch.float16)   # a plain 16-bit copy
        self.src, self.dst = src, dst
        self._lib = N.load()

    def launch(self, stream=None):
        s, d = self.src, self.dst
        b, h, w = s.bhw
        N.check(self._lib.glsdet_upsample2x(s.t.data_ptr(), d.t.data_ptr(), b, h, w, s.c, s.ld, s.coff, d.ld, d.coff,
                                            N.stream_ptr(stream)), "glsdet_upsample2x")


class DepthwiseOp:
    """dconv half of the reference's DWConv (models/base/baseConv.py:22-30): depthwise k x k conv (groups = channels) with
    the folded BatchNorm and the activation, NHWC channel window -> NHWC channel window of the same storage type (16-bit,
    or fp32 in the accuracy mode).  `weight` [C, 1, k, k] and `bias` [C] are the BN-folded fp32 values."""

    def __init__(self, src: View, weight: torch.Tensor, bias: torch.Tensor, *, stride: int = 1, act: int = N.ACT_SILU,
                 out: View):
        c, one, k, k2 = weight.shape
        assert one == 1 and k == k2 and c == src.c == out.c, (tuple(weight.shape), src.c, out.c)
        b, h, w = src.bhw
        pad = (k - 1) // 2
        ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
        assert out.bhw == (b, ho, wo) and out.t.dtype == src.t.dtype, (out.bhw, (b, ho, wo), out.t.dtype, src.t.dtype)
        dt = src.t.dtype
        self.dt = N.DT_F32 if dt == torch.float32 else N.dt_code(dt)
        dev = src.t.device
        self.w = weight.detach().float().reshape(c, k * k).t().contiguous().to(dev)   # [tap][C]
        self.b = bias.detach().float().contiguous().to(dev)
        self.src, self.out, self.k, self.stride, self.act = src, out, k, stride, act
        self.flops = 2.0 * b * ho * wo * c * k * k
        self._lib = N.load()

    def launch(self, stream=None):
        s, o = self.src, self.out
        b, h, w = s.bhw
        N.check(self._lib.glsdet_dwconv(s.t.data_ptr(), s.ld, s.coff, o.t.data_ptr(), o.ld, o.coff, b, h, w, s.c, self.k,
                                        self.stride, self.w.data_ptr(), self.b.data_ptr(), self.act, self.dt,
                                        N.stream_ptr(stream)), "glsdet_dwconv")


class ConvOpF32:
    """fp32 accuracy-mode twin of ConvOp (glsdet_conv_f32: SIMT fp32 implicit GEMM, every tensor fp32).  Same
    arguments as ConvOp minus the tensor-core-only ones (fused prediction conv, batched / patch modes)."""

    def __init__(self, srcs: Sequence[View], weight: torch.Tensor, bias: Optional[torch.Tensor], *, ksize: int,
                 stride: int = 1, act: int = N.ACT_NONE, out, out_mode: int = N.OUT_NHWC_F32, out_ld: int = 0,
                 out_coff: int = 0, out_batch_stride: int = 0, pre_res: Optional[View] = None, pre_shift: int = 0,
                 post_res: Optional[View] = None, post_shift: int = 0, dec=(0.0, 0.0, 0.0)):
        self._lib = N.load()
        assert 1 <= len(srcs) <= 2
        b, h, w = srcs[0].bhw
        for s in srcs:
            assert s.bhw == (b, h, w) and s.t.dtype == torch.float32
        n_out, cin = weight.shape[0], weight.shape[1]
        assert cin == sum(s.c for s in srcs)
        # K order (source, tap, channel)
        segs, a = [], 0
        for s in srcs:
            segs.append(weight[:, a:a + s.c].permute(0, 2, 3, 1).reshape(n_out, -1))
            a += s.c
        dev = srcs[0].t.device
        self.packed = torch.cat(segs, 1).float().contiguous().to(dev)
        self.bias = None if bias is None else bias.detach().float().contiguous().to(dev)
        d = N.ConvF32Desc()
        d.src0, d.src0_c, d.src0_ld = srcs[0].ptr, srcs[0].c, srcs[0].ld
        if len(srcs) == 2:
            d.src1, d.src1_c, d.src1_ld = srcs[1].ptr, srcs[1].c, srcs[1].ld
        d.batch, d.height, d.width, d.ksize, d.stride = b, h, w, ksize, stride
        d.weight, d.out_channels = self.packed.data_ptr(), n_out
        d.bias = 0 if self.bias is None else self.bias.data_ptr()
        d.act = act
        ho, wo = h // stride, w // stride
        if pre_res is not None:
            assert pre_res.t.dtype == torch.float32 and pre_res.bhw == (b, max(ho >> pre_shift, 1), max(wo >> pre_shift, 1))
            d.pre_res, d.pre_shift, d.pre_ld = pre_res.ptr, pre_shift, pre_res.ld
        if post_res is not None:
            assert post_res.t.dtype == torch.float32 and post_res.bhw == (b, ho >> post_shift, wo >> post_shift)
            d.post_res, d.post_shift, d.post_ld = post_res.ptr, post_shift, post_res.ld
        if isinstance(out, View):
            assert out.bhw == (b, ho, wo) and out.c >= n_out and out.t.dtype == torch.float32
            d.out, d.out_ld, d.out_coff = out.t.data_ptr(), out.ld, out.coff
            d.out_batch_stride = ho * wo * out.ld
            d.out_mode = N.OUT_NHWC_F32
            self._out_t = out.t
        else:
            assert out.dtype == torch.float32
            d.out, d.out_ld, d.out_coff = out.data_ptr(), out_ld, out_coff
            d.out_batch_stride, d.out_mode = out_batch_stride, out_mode
            self._out_t = out
        d.dec_stride, d.dec_in_w, d.dec_in_h = dec
        self.desc = d
        self._keep = (srcs, pre_res, post_res)
        self.flops = 2.0 * b * ho * wo * n_out * cin * ksize * ksize

    def launch(self, stream=None):
        N.check(self._lib.glsdet_conv_f32(C.byref(self.desc), N.stream_ptr(stream)), "glsdet_conv_f32")


class BGemmF32Op:
    """Batched fp32 GEMM on NHWC fp32 channel windows (glsdet_bgemm_f32), the two products of the dot-product non-local
    block in the fp32 accuracy mode.  mode "gram":  out[b][i][j] = alpha * sum_t a[b, t, i] * b_[b, t, j]  (a, b_: windows of
    [B, h, w, .] tensors, t = pixel; out: [B, Ci, Cj] fp32);  mode "apply":  out[b, t, j] = sum_i a[b, t, i] * m[b][i][j]
    (a: window, m: [B, Ci, Cj] fp32, out: window)."""

    def __init__(self, mode: str, a: View, b, out, alpha: float = 1.0):
        assert mode in ("gram", "apply") and a.t.dtype == torch.float32
        bb, h, w = a.bhw
        t = h * w
        self._keep = (a, b, out)
        self._lib = N.load()
        es = 4
        if mode == "gram":
            assert isinstance(b, View) and b.bhw == a.bhw and b.t.dtype == torch.float32
            assert out.dtype == torch.float32 and tuple(out.shape) == (bb, a.c, b.c) and out.is_contiguous()
            self.args = (a.ptr, 1, a.ld, t * a.ld, b.ptr, b.ld, t * b.ld, out.data_ptr(), b.c, a.c * b.c, a.c, b.c, t,
                         float(alpha), bb)
            self.flops = 2.0 * bb * t * a.c * b.c
        else:
            assert b.dtype == torch.float32 and b.dim() == 3 and b.shape[0] == bb and b.shape[1] == a.c and b.is_contiguous()
            assert isinstance(out, View) and out.bhw == a.bhw and out.c == b.shape[2] and out.t.dtype == torch.float32
            self.args = (a.ptr, 0, a.ld, t * a.ld, b.data_ptr(), b.shape[2], b.shape[1] * b.shape[2], out.ptr, out.ld,
                         t * out.ld, t, b.shape[2], a.c, float(alpha), bb)
            self.flops = 2.0 * bb * t * a.c * b.shape[2]
        del es

    def launch(self, stream=None):
        N.check(self._lib.glsdet_bgemm_f32(*self.args, N.stream_ptr(stream)), "glsdet_bgemm_f32")


class RectCopyOp:
    """Up to 8 rectangle copies between two NHWC bf16 tensors in one launch (glsdet_rect_copy):
    rects = [(src_image0, sy, sx, dst_image0, dy, dx, h, w), ...], each applied to `batch` consecutive images."""

    def __init__(self, src: View, dst: View, batch: int, rects):
        src, dst = as16(src), as16(dst)   # fp32 accuracy mode: copied as 16-bit pairs
        assert src.t.dtype == dst.t.dtype and src.t.dtype in (torch.bfloat16, torch.float16) and src.c == dst.c
        assert 1 <= len(rects) <= 8
        self.src, self.dst, self.batch = src, dst, batch
        self.rects = (N.Rect * len(rects))(*[N.Rect(*r) for r in rects])
        for r in rects:
            assert r[0] + batch <= src.t.shape[0] and r[3] + batch <= dst.t.shape[0], (r, src.t.shape, dst.t.shape)
        self.n = len(rects)
        self._lib = N.load()

    def launch(self, stream=None):
        s, d = self.src, self.dst
        N.check(self._lib.glsdet_rect_copy(s.t.data_ptr(), s.t.shape[1], s.t.shape[2], s.ld, s.coff, d.t.data_ptr(),
                                           d.t.shape[1], d.t.shape[2], d.ld, d.coff, self.batch, s.c, self.rects, self.n,
                                           N.stream_ptr(stream)), "glsdet_rect_copy")


class NhwcTransposeOp:
    """NHWC bf16 [B', h, w, C] window -> per-image transposed matrices dst[b'][c][t] (Gram operand)."""

    def __init__(self, src: View, dst: torch.Tensor, scale: float = 1.0):
        self.scale = float(scale)
        assert src.t.dtype == dst.dtype and dst.dtype in (torch.bfloat16, torch.float16) and dst.dim() == 3 and dst.is_contiguous()
        b, h, w = src.bhw
        assert dst.shape[0] == b and dst.shape[1] >= src.c and dst.shape[2] >= h * w
        self.src, self.dst, self.b, self.t = src, dst, b, h * w
        self._lib = N.load()

    def launch(self, stream=None):
        s = self.src
        N.check(self._lib.glsdet_nhwc_transpose_16(s.t.data_ptr(), self.dst.data_ptr(), self.b, self.t, s.c, s.ld, s.coff,
                                                   self.dst.shape[1], self.dst.shape[2], N.dt_code(self.dst.dtype),
                                                   self.scale, N.stream_ptr(stream)), "glsdet_nhwc_transpose_16")


class FocusOp:
    """Focus space-to-depth of the image batch (models/ffa/darknet.py:15-21): NCHW fp32 [B, 3, H, W] -> NHWC bf16
    [B, H/2, W/2, 16] (12 channels in the reference's cat order + 4 zero channels)."""

    def __init__(self, dst):
        """`dst`: a contiguous [B, H/2, W/2, 16] tensor, or a FoldedView (zero-bordered rows, written at pixel x + 1)."""
        self.f32 = False
        if isinstance(dst, FoldedView):
            assert dst.c_pix == 16
            self.shape, self.border, self.dst = dst.bhw, 1, dst.t
        elif dst.dtype == torch.float32:   # fp32 accuracy mode: [B, H/2, W/2, 12]
            assert dst.dim() == 4 and dst.shape[3] == 12 and dst.is_contiguous()
            self.shape, self.border, self.dst, self.f32 = tuple(dst.shape[:3]), 0, dst, True
        else:
            assert dst.dtype in (torch.bfloat16, torch.float16) and dst.dim() == 4 and dst.shape[3] == 16 and dst.is_contiguous()
            self.shape, self.border, self.dst = tuple(dst.shape[:3]), 0, dst
        self._lib = N.load()

    def launch(self, image: torch.Tensor, stream=None):
        b, h2, w2 = self.shape
        assert image.dtype == torch.float32 and image.is_contiguous() and image.is_cuda
        assert tuple(image.shape) == (b, 3, 2 * h2, 2 * w2), (tuple(image.shape), self.shape)
        if self.f32:
            N.check(self._lib.glsdet_focus_nchw_f32_to_nhwc_f32(image.data_ptr(), self.dst.data_ptr(), b, 2 * h2, 2 * w2,
                                                                N.stream_ptr(stream)), "glsdet_focus_nchw_f32_to_nhwc_f32")
            return
        N.check(self._lib.glsdet_focus_nchw_f32_to_nhwc_16(image.data_ptr(), self.dst.data_ptr(), b, 2 * h2, 2 * w2,
                                                           self.border, N.dt_code(self.dst.dtype), N.stream_ptr(stream)),
                "glsdet_focus_nchw_f32_to_nhwc_16")


    def launch_u8(self, image_u8: torch.Tensor, mean, std, stream=None):
        """uint8 HWC image batch [B, H, W, 3]; normalisation of models/core/utils.py:47-51 fused (bf16 plans only)."""
        b, h2, w2 = self.shape
        assert not self.f32 and image_u8.dtype == torch.uint8 and image_u8.is_contiguous() and image_u8.is_cuda
        assert tuple(image_u8.shape) == (b, 2 * h2, 2 * w2, 3), (tuple(image_u8.shape), self.shape)
        m = (C.c_double * 3)(*[float(v) for v in mean])
        s = (C.c_double * 3)(*[float(v) for v in std])
        N.check(self._lib.glsdet_focus_u8_to_nhwc_16(image_u8.data_ptr(), self.dst.data_ptr(), b, 2 * h2, 2 * w2,
                                                     self.border, m, s, N.dt_code(self.dst.dtype), N.stream_ptr(stream)),
                "glsdet_focus_u8_to_nhwc_16")


class SppPoolOp:
    """MaxPool2d(5 / 9 / 13, 1, k // 2) of channel window 0 of an NHWC bf16 concat buffer [B, h, w, 4C] into windows
    1..3 (SPPBottleneck.forward, models/ffa/darknet.py:33-36)."""

    def __init__(self, cat: torch.Tensor, channels: int):
        assert cat.dtype in (torch.bfloat16, torch.float16, torch.float32) and cat.dim() == 4 and cat.is_contiguous() and cat.shape[3] == 4 * channels
        self.cat, self.c = cat, channels
        self._lib = N.load()

    def launch(self, stream=None):
        b, h, w, ld = self.cat.shape
        c = self.c
        if self.cat.dtype == torch.float32:
            N.check(self._lib.glsdet_spp_maxpool_f32(self.cat.data_ptr(), b, h, w, c, ld, 0, c, 2 * c, 3 * c,
                                                     N.stream_ptr(stream)), "glsdet_spp_maxpool_f32")
            return
        N.check(self._lib.glsdet_spp_maxpool_16(self.cat.data_ptr(), b, h, w, c, ld, 0, c, 2 * c, 3 * c,
                                                N.dt_code(self.cat.dtype), N.stream_ptr(stream)), "glsdet_spp_maxpool_16")
