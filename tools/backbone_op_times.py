"""Warm CUDA-event time of every op of the native CSPDarknet plan (YOLOX-s, 16 x 1024^2), plus the whole backbone."""
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from glsdet_b200.backbone import BackbonePlan  # noqa: E402
from glsdet_b200.ops import ConvOp  # noqa: E402
from glsdet_b200.synthetic import synthetic_images  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    sd = bench.make_weights()
    B = 16
    plan = BackbonePlan(sd, B, (bench.IN_H, bench.IN_W), device=dev)
    x = torch.cat([synthetic_images(4, bench.IN_H, bench.IN_W, seed=i) for i in range(B // 4)]).to(dev)
    ops = [("focus", None)] + [("op", op) for op in plan.ops]
    reps = 12
    times = [[] for _ in ops]
    whole = []
    for r in range(reps + 3):
        evs = []
        for _, op in ops:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.focus.launch(x) if op is None else op.launch()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run(x)
        e1.record()
        torch.cuda.synchronize()
        if r >= 3:
            whole.append(e0.elapsed_time(e1) * 1e3)
            for i, (a, b) in enumerate(evs):
                times[i].append(a.elapsed_time(b) * 1e3)
    tot = 0.0
    for (g, op), ts in zip(ops, times):
        t = statistics.median(ts)
        tot += t
        if isinstance(op, ConvOp):
            d = op.desc
            tag = f"{d.ksize}x{d.ksize}/{d.stride} {d.src0_c + d.src1_c:4d}->{d.out_channels:3d} @{d.height}x{d.width}"
            print(f"{t:8.1f} us  {op.flops / t / 1e6:7.1f} TF/s  {tag}")
        else:
            print(f"{t:8.1f} us  {'':12s}  {'FocusOp' if op is None else type(op).__name__}")
    print(f"sum {tot:.1f} us over {len(ops)} ops; whole backbone (no events between ops) {statistics.median(whole):.1f} us; "
          f"{plan.flops / B / 1e9:.1f} GFLOP per image")


if __name__ == "__main__":
    main()
