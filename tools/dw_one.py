"""One depthwise launch shape for an ncu capture: 3x3/1, C = 64, 16 x 256^2, fp16 storage (the largest nano layer)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from glsdet_b200 import _native as N  # noqa: E402
from glsdet_b200.ops import DepthwiseOp, View  # noqa: E402

dev = torch.device("cuda")
c, hw, b = 64, 256, 16
x = torch.randn(b, hw, hw, c, device=dev).to(torch.float16)
out = torch.empty_like(x)
op = DepthwiseOp(View(x), torch.randn(c, 1, 3, 3) / 3, torch.zeros(c), stride=1, act=N.ACT_SILU, out=View(out))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()
    op.launch()
torch.cuda.synchronize()
