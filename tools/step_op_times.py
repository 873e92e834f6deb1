"""Warm, in-pipeline time of every op of a bench step (CUDA events around each launch, median over repeated steps).
Event pairs serialise the ops (no programmatic-dependent-launch overlap), so the sum is a little above the real step."""
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from glsdet_b200.ops import ConvOp  # noqa: E402


def main():
    if len(sys.argv) > 1 and sys.argv[1] in ("p0", "p1", "p2"):
        bench.VARIANT = sys.argv[1]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    sd = bench.make_weights()
    net = bench.body_class()(bench.NUM_CLASSES, bench.PHI)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    B = 16
    feats = bench.make_features(net, B, 1000, dev)
    plan = net.plan_for(feats)
    groups = (("neck", plan.neck_ops), ("stems", plan.stem_ops), ("tower", plan.tower_ops), ("pred", plan.pred_det_ops))
    ops = [(g, op) for g, lst in groups for op in lst]
    reps = 12
    times = [[] for _ in ops]
    for r in range(reps + 3):
        plan.load_features(feats)
        evs = []
        for _, op in ops:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            op.launch()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        if r >= 3:
            for i, (e0, e1) in enumerate(evs):
                times[i].append(e0.elapsed_time(e1) * 1e3)
    tot = 0.0
    for (g, op), ts in zip(ops, times):
        t = statistics.median(ts)
        tot += t
        if isinstance(op, ConvOp):
            d = op.desc
            tag = f"{d.ksize}x{d.ksize}/{d.stride} {d.src0_c + d.src1_c:4d}->{d.out_channels:3d}{'+p' + str(d.pred_channels) if d.pred_channels else '':4s} @{d.height}x{d.width}"
            print(f"{g:5s} {t:8.1f} us  {op.flops / t / 1e6:7.1f} TF/s  {tag}")
        else:
            print(f"{g:5s} {t:8.1f} us  {'':12s}  {type(op).__name__}")
    print(f"sum {tot:.1f} us over {len(ops)} ops")


if __name__ == "__main__":
    main()
