"""Throughput of the fp32 accuracy mode (BASELINE configs[0] semantics: every tensor fp32, SIMT fp32 kernels,
csrc/fp32_path.cu) on the headline workload, beside the 16-bit tensor-core path: P0 YOLOX-s, 16 x 1024^2, device-resident
features -> detections (neck + head + decode + NMS), CUDA events."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict  # noqa: E402
from glsdet_b200.yolox_ffa import YoloBody  # noqa: E402


def main():
    dev = torch.device("cuda")
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    net = YoloBody(10, "s")
    net.load_state_dict(synthetic_state_dict(10, "s", seed=0, flavour="calibrated"), strict=True)
    net = net.to(dev).eval()
    x = synthetic_images(b, 1024, 1024, seed=1000).to(dev)
    feats = [f.contiguous() for f in net.backbone.features(x)]
    for prec in ("bf16", "fp32"):
        net.set_precision(prec)
        for _ in range(2):
            det, cnt = net.detect_features(feats, conf_thres=0.01, nms_thres=0.65)
        torch.cuda.synchronize()
        n = 10 if prec == "bf16" else 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            det, cnt = net.detect_features(feats, conf_thres=0.01, nms_thres=0.65)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        gf = 140.2 * b
        print(f"{prec}: {ms:8.2f} ms per {b} x 1024^2 = {b / ms * 1e3:7.1f} img/s, {gf / ms:7.1f} TFLOP/s algorithmic; "
              f"kept {int(cnt.sum())} rows")


if __name__ == "__main__":
    main()
