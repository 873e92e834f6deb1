"""Achieved HBM bandwidth of the depthwise kernel (csrc/dwconv.cu) on the phi = 'nano' layer shapes of a 16 x 1024^2 batch,
L2 flushed between launches, CUDA events; algorithmic bytes = input read once + output written once (16-bit storage),
against the measured copy bandwidth of MEASURED_PEAKS.json.  Then one nano P0 step (image -> detections) for reference."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from glsdet_b200 import _native as N  # noqa: E402
from glsdet_b200.ops import DepthwiseOp, View  # noqa: E402


def main():
    dev = torch.device("cuda")
    peak = 6548.5
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    b = 16
    # (channels, input h = w, k, stride): backbone stage entries, Bottleneck 3x3 convs, neck bu_convs, head towers of nano @1024^2
    shapes = [(16, 512, 3, 2), (32, 256, 3, 2), (16, 256, 3, 1), (64, 128, 3, 2), (32, 128, 3, 1), (64, 256, 3, 1),
              (64, 128, 3, 1), (128, 64, 3, 2), (64, 64, 3, 1), (128, 32, 3, 1)]
    for c, hw, k, s in shapes:
        x = torch.randn(b, hw, hw, c, device=dev).to(torch.float16)
        out = torch.empty(b, hw // s, hw // s, c, dtype=torch.float16, device=dev)
        wt = torch.randn(c, 1, k, k) / k
        op = DepthwiseOp(View(x), wt, torch.zeros(c), stride=s, act=N.ACT_SILU, out=View(out))
        ts = []
        for _ in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); op.launch(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = sorted(ts)[len(ts) // 2]
        nbytes = (x.numel() + out.numel()) * 2
        gbs = nbytes / us * 1e-3
        print(f"dwconv {k}x{k}/{s} C={c:4d} @{hw:4d}^2 x{b}: {us:7.1f} us  {nbytes / 1e6:7.1f} MB  {gbs:7.0f} GB/s  "
              f"{gbs / peak:5.2f} of the measured HBM peak ({peak:.0f} GB/s)")
    # one nano step, image -> detections
    from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict
    from glsdet_b200.yolox_ffa import YoloBody

    net = YoloBody(10, "nano")
    net.load_state_dict(synthetic_state_dict(10, "nano", seed=11, flavour="calibrated"), strict=True)
    net = net.to(dev).eval()
    x = synthetic_images(16, 1024, 1024, seed=3).to(dev)
    for _ in range(3):
        det, cnt = net.detect(x, conf_thres=0.01, nms_thres=0.65)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        det, cnt = net.detect(x, conf_thres=0.01, nms_thres=0.65)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"nano P0 image -> detections, 16 x 1024^2 device-resident: {ms:.2f} ms per batch = {16 / ms * 1e3:.0f} img/s; "
          f"kept per image {cnt.tolist()}")


if __name__ == "__main__":
    main()
