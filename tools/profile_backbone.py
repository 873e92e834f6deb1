"""One pass of the native CSPDarknet plan (YOLOX-s, 16 x 1024^2) for ncu: python tools/profile_backbone.py [passes]."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from glsdet_b200.backbone import BackbonePlan  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    sd = bench.make_weights()
    plan = BackbonePlan(sd, 16, (bench.IN_H, bench.IN_W), device=dev)
    x = torch.randn(16, 3, bench.IN_H, bench.IN_W, device=dev)
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        plan.run(x)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
