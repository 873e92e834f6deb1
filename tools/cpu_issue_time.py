"""Host time needed to ISSUE one device-resident step (71 ctypes launches) against the GPU time of the step: is the
launch thread the limiter?"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    sd = bench.make_weights()
    net = bench.body_class()(bench.NUM_CLASSES, bench.PHI)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    feats = bench.make_features(net, 16, 1000, dev)
    plan = net.plan_for(feats)
    nms = net.nms_for(plan, None)

    def step():
        plan.load_features(feats)
        plan.run_neck()
        plan.run_head("det")
        nms.launch(plan.pred, bench.CONF_THRES, bench.NMS_THRES, "auto_cuda", cls_logits=plan.det_cls_logits)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    for n in (1, 3):    # few steps: the launch queue never fills, so perf_counter measures pure issue time
        t0 = time.perf_counter()
        for _ in range(n):
            step()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"{n} step(s): issue {1e3 * (t1 - t0) / n:.3f} ms per step on the host, {1e3 * (t2 - t0) / n:.3f} ms until the GPU is done")


if __name__ == "__main__":
    main()
