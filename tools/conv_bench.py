"""Micro-benchmark of single conv ops (CUDA-event timing, L2 flushed between timed launches)."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from glsdet_b200 import _native as N  # noqa: E402
from glsdet_b200.ops import ConvOp, View  # noqa: E402

CASES = {
    # name: (B, H, W, cins, N, k, stride)
    "head3x3_s4_n128": (16, 256, 256, [128], 128, 3, 1),
    "head3x3_s4_n256": (16, 256, 256, [128], 256, 3, 1),
    "head3x3_s8_n256": (16, 128, 128, [128], 256, 3, 1),
    "csp1x1_c64_n128": (16, 256, 256, [64], 128, 1, 1),
    "csp3x3_c64_n64": (16, 256, 256, [64], 64, 3, 1),
    "ffa1x1_c512_n512": (16, 64, 64, [512], 512, 1, 1),
    "pred_n10": (16, 256, 256, [128], 10, 1, 1),
    "csp1x1_c64_n64": (16, 256, 256, [64], 64, 1, 1),
    "csp1x1_c128_n128": (16, 256, 256, [128], 128, 1, 1),
    "neck1x1_c256_n128_s8": (16, 128, 128, [256], 128, 1, 1),
    "neck3x3_c128_n128_s16": (16, 64, 64, [128], 128, 3, 1),
    # second tower conv + fused prediction conv (10 classes, sigmoid rows)
    "tower_pred_s4": (16, 256, 256, [128], 128, 3, 1, 10),
    "tower_pred_s8": (16, 128, 128, [128], 128, 3, 1, 10),
    # residual variants (8th field: "post1" = bf16 residual at half resolution added after the activation, "post0" = same
    # resolution, "pre1" = fp32 pre-activation partial sums at half resolution)
    "csp1x1_c128_n128_post1": (16, 256, 256, [128], 128, 1, 1, "post1"),
    "ffa1x1_c256_n128_post0": (16, 128, 128, [256], 128, 1, 1, "post0"),
    "ffa1x1_c256_n128": (16, 128, 128, [256], 128, 1, 1),
    "c3p3_c128_n128_pre1": (16, 128, 128, [128], 128, 1, 1, "pre1"),
    "c3p3_c128_n128": (16, 128, 128, [128], 128, 1, 1),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default=",".join(CASES))
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--act", default="silu")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name in args.cases.split(","):
        B, H, W, cins, n, k, s = CASES[name][:7]
        extra = CASES[name][7] if len(CASES[name]) > 7 else 0
        n_pred = extra if isinstance(extra, int) else 0
        res_kw = {}
        if isinstance(extra, str):
            sh = int(extra[-1])
            if extra.startswith("post"):
                res = torch.randn(B, (H // s) >> sh, (W // s) >> sh, n, device=dev).to(torch.bfloat16)
                res_kw = dict(post_res=View(res), post_shift=sh)
            else:
                res = torch.randn(B, (H // s) >> sh, (W // s) >> sh, n, device=dev)
                res_kw = dict(pre_res=View(res), pre_shift=sh)
        srcs = [View(torch.randn(B, H, W, c, device=dev).to(torch.bfloat16)) for c in cins]
        w = torch.randn(n, sum(cins), k, k, device=dev) * 0.05
        bias = torch.randn(n, device=dev)
        out = torch.empty(B, H // s, W // s, n, device=dev, dtype=torch.bfloat16)
        if n_pred:
            rows = torch.empty(B, (H // s) * (W // s), n_pred + 5, device=dev)
            wp = torch.randn(n_pred, n, 1, 1, device=dev) * 0.05
            op = ConvOp(srcs, w, bias, ksize=k, stride=s, act=N.ACT_BY_NAME[args.act], out=rows,
                        out_mode=N.OUT_NHWC_F32, out_ld=n_pred + 5, out_coff=5,
                        out_batch_stride=rows.shape[1] * (n_pred + 5), pred_weight=wp,
                        pred_bias=torch.zeros(n_pred, device=dev), pred_act=N.ACT_SIGMOID)
        else:
            op = ConvOp(srcs, w, bias, ksize=k, stride=s, act=N.ACT_BY_NAME[args.act], out=View(out), **res_kw)
        for _ in range(3):
            op.launch()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            op.launch()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        in_bytes = sum(B * H * W * c * 2 for c in cins)
        out_bytes = out.numel() * 2
        print(f"{name:22s} {t*1e3:9.1f} us  {op.flops/t/1e9:8.1f} TFLOP/s  "
              f"{(in_bytes+out_bytes)/t/1e6:8.1f} GB/s (algorithmic)  block_n={op.block_n}", flush=True)


if __name__ == "__main__":
    main()
