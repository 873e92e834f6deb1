"""Does the post-processing of step i hide behind the conv segment of step i+1 (second stream)?  Serial vs overlapped
device-resident step time of the bench workload."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


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
    nms = net.nms_for(plan, None)
    main_s = torch.cuda.current_stream()
    post_s = torch.cuda.Stream(device=dev)
    outs = []
    for _ in range(2):
        det, cnt = nms.launch(plan.pred, bench.CONF_THRES, bench.NMS_THRES, "auto_cuda", cls_logits=plan.det_cls_logits)
        outs.append((det, cnt, nms.keep_index))
    torch.cuda.synchronize()

    def serial(n):
        for _ in range(n):
            plan.load_features(feats)
            plan.run_neck()
            plan.run_head("det")
            nms.launch(plan.pred, bench.CONF_THRES, bench.NMS_THRES, "auto_cuda", cls_logits=plan.det_cls_logits, out=outs[0])

    conv_done = [torch.cuda.Event() for _ in range(2)]
    nms_done = [torch.cuda.Event() for _ in range(2)]

    def overlapped(n):
        for i in range(n):
            plan.load_features(feats)
            plan.run_neck()
            plan.run_stems()
            plan._run(plan.tower_ops, None)
            if i > 0:
                main_s.wait_event(nms_done[(i - 1) & 1])     # the previous step's filter has consumed plan.pred
            plan._run(plan.pred_det_ops, None)
            conv_done[i & 1].record(main_s)
            post_s.wait_event(conv_done[i & 1])
            nms.launch(plan.pred, bench.CONF_THRES, bench.NMS_THRES, "auto_cuda", cls_logits=plan.det_cls_logits,
                       out=outs[i & 1], stream=post_s)
            nms_done[i & 1].record(post_s)
        main_s.wait_event(nms_done[(n - 1) & 1])

    for name, fn in (("serial", serial), ("overlapped", overlapped), ("serial", serial), ("overlapped", overlapped)):
        fn(5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        e0.record()
        fn(n)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{name:10s} {ms:.3f} ms/step  {B / ms * 1e3:.0f} img/s  kept {int(outs[0][1].sum())} {int(outs[1][1].sum())}")


if __name__ == "__main__":
    main()
