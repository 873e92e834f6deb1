"""Warm CUDA-event time of the Focus kernels (fp32 NCHW and uint8 HWC entry) at 16 x 1024^2."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from glsdet_b200.ops import FocusOp, FoldedView  # noqa: E402


def main():
    dev = torch.device("cuda")
    b, h, w = 16, 1024, 1024
    fa = FoldedView(torch.zeros(b * (h // 2) * (w // 2 + 2) * 16 + 64, dtype=torch.bfloat16, device=dev), b, h // 2, w // 2, 16)
    x = torch.randn(b, 3, h, w, device=dev)
    u = torch.randint(0, 256, (b, h, w, 3), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    op = FocusOp(fa)
    for name, fn in (("fp32 NCHW", lambda: op.launch(x)), ("uint8 HWC", lambda: op.launch_u8(u, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)))):
        ts = []
        for _ in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"focus {name}: {sorted(ts)[len(ts) // 2]:.1f} us (L2 flushed)")


if __name__ == "__main__":
    main()
