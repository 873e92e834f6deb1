"""Where the time of every conv launch of a bench step goes (diagnostic, GLSDET_CONV_TRACE=1): %globaltimer stamps of
CTA 0 of each launch - launch gap after the previous kernel's exit, prologue, wait for the predecessor, set-up, first
operand, main loop, epilogue, tail."""
import ctypes as C
import os
import sys
from pathlib import Path

os.environ["GLSDET_CONV_TRACE"] = "1"
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from glsdet_b200.ops import ConvOp  # noqa: E402


def main():
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

    def step():
        plan.load_features(feats)
        plan.run_neck()
        plan.run_head("det")

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    buf = (C.c_uint64 * 16)()
    for _, op in ops:
        if isinstance(op, ConvOp):
            op._lib.glsdet_conv_read_trace(op.handle, buf)
    step()
    torch.cuda.synchronize()
    prev_exit = None
    print(f"{'layer':42s} {'gap':>6s} {'prolog':>6s} {'depwait':>7s} {'setup':>6s} {'1stTMA':>6s} {'1stA':>6s} {'mma':>7s} {'epi':>6s} {'drain':>6s} {'exit':>6s} {'total':>7s}  (us, CTA 0)")
    tot_fixed = 0.0
    for g, op in ops:
        if not isinstance(op, ConvOp):
            prev_exit = None
            print(f"{g:6s} {type(op).__name__}")
            continue
        op._lib.glsdet_conv_read_trace(op.handle, buf)
        t = [int(v) for v in buf]
        d = op.desc
        tag = f"{g:6s} {d.ksize}x{d.ksize}/{d.stride} {d.src0_c + d.src1_c:4d}->{d.out_channels:3d} @{d.height}x{d.width}"
        us = lambda a, b: (t[b] - t[a]) / 1e3 if t[a] and t[b] else float("nan")
        gap = (t[0] - prev_exit) / 1e3 if prev_exit else float("nan")
        end = t[10] or t[9]
        print(f"{tag:42s} {gap:6.1f} {us(0, 1):6.1f} {0.0:7.1f} {us(1, 2):6.1f} {us(2, 3):6.1f} {us(3, 4):6.1f} {us(4, 5):7.1f} "
              f"{us(5, 7):6.1f} {us(7, 8):6.1f} {(end - t[8]) / 1e3 if t[8] else float('nan'):6.1f} {(end - t[0]) / 1e3:7.1f}")
        prev_exit = end


if __name__ == "__main__":
    main()
