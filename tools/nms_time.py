"""Warm time of the post-processing pipeline alone on the bench workload's predictions (CUDA events)."""
import sys
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
    import os
    md = int(os.environ.get("GLSDET_PROFILE_MAXDET", "0"))   # 0: every survivor (the bench's default), else the top-K variant
    nms = net.nms_for(plan, md if md > 0 else None)
    prob = plan.forward_decoded(feats).clone()       # probabilities, planes view
    raw = plan.forward_detect(feats).clone()         # raw class logits, planes view
    for name, pred, kw in (("planes, probabilities", prob, {}), ("planes, raw class logits", raw, dict(cls_logits=True)),
                           ("rows, probabilities", prob.contiguous(), {})):
        for _ in range(5):
            nms.launch(pred, bench.CONF_THRES, bench.NMS_THRES, "auto_cuda", **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            nms.launch(pred, bench.CONF_THRES, bench.NMS_THRES, "auto_cuda", **kw)
        e1.record()
        torch.cuda.synchronize()
        print(f"post-processing ({name}): {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per batch of 16")
    # the SE gate of the FFA stage alone (se_partial + se_fc), warm
    from glsdet_b200.ops import SeGateOp
    se = [op for op in getattr(plan, "stem_ops", []) if isinstance(op, SeGateOp)]
    for op in se[:1]:
        for _ in range(5):
            op.launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            op.launch()
        e1.record()
        torch.cuda.synchronize()
        print(f"SE gate (se_partial + se_fc, warm, back to back): {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")


if __name__ == "__main__":
    main()
