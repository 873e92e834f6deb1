"""One profiled step of the bench workload, bracketed by cudaProfilerStart/Stop (run under
`ncu --profile-from-start off ...`).  Also writes the op list of the step (shapes, FLOPs) so that the per-launch
ncu table can be joined with layer shapes: gpurun_out/step_ops.json."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from glsdet_b200.ops import ConvOp  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    if len(sys.argv) > 1 and sys.argv[1] in ("p0", "p1", "p2"):
        bench.VARIANT = sys.argv[1]
    sd = bench.make_weights()
    YoloBody = bench.body_class()

    net = YoloBody(bench.NUM_CLASSES, bench.PHI)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    B = 16
    feats = bench.make_features(net, B, 1000, dev)
    plan = net.plan_for(feats)
    import os
    md = int(os.environ.get('GLSDET_PROFILE_MAXDET', '0'))
    nms = net.nms_for(plan, md if md > 0 else None)

    def step():
        plan.load_features(feats)
        plan.run_neck()
        plan.run_head("det")
        nms.launch(plan.pred, bench.CONF_THRES, bench.NMS_THRES, "auto_cuda", cls_logits=plan.det_cls_logits)

    ops = []
    for group, lst in (("neck", plan.neck_ops), ("stems", plan.stem_ops), ("tower", plan.tower_ops), ("pred", plan.pred_det_ops)):
        for op in lst:
            if isinstance(op, ConvOp):
                d = op.desc
                ops.append(dict(group=group, k=d.ksize, s=d.stride, cin=d.src0_c + d.src1_c, n=d.out_channels,
                                h=d.height, w=d.width, gflop=op.flops / 1e9, pred=int(d.pred_channels),
                                batched=int(d.weight_batch_stride != 0), patch=int(d.patch_mode)))
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / f"step_ops_{bench.VARIANT}.json").write_text(json.dumps(ops))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
