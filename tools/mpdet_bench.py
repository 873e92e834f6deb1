"""Timing of the MP-Det neck / head (BASELINE configs[2] shapes: 800 x 1344 mosaic -> C2..C5 of a ResNet-50) on one GPU:
FPN + MPHead + selection + NMS, device-resident, CUDA events; plus the oracle port on the host CPU for one image."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import glsdet_b200.mpdet  # noqa: E402,F401
from glsdet_b200.mmdet_face import HEADS, NECKS  # noqa: E402
from oracle import mmdet_ref as M  # noqa: E402


def main(batch=8, H=800, W=1344):
    dev = torch.device("cuda:0")
    sd = M.mpdet_synthetic_state_dict(0)
    neck = NECKS.build(dict(type="FPN", in_channels=[256, 512, 1024, 2048], out_channels=256, start_level=1,
                            add_extra_convs="on_output", num_outs=5))
    head = HEADS.build(dict(type="MPHead", num_classes=10, in_channels=256, stacked_convs=4, feat_channels=256,
                            test_cfg=dict(nms_pre=1000, score_thr=0.05, nms=dict(type="nms", iou_threshold=0.6), max_per_img=500)))
    nsd = {k[5:]: v for k, v in sd.items() if k.startswith("neck.")}
    hsd = {k[10:]: v for k, v in sd.items() if k.startswith("bbox_head.")}
    neck.load_state_dict(nsd, strict=True)
    head.load_state_dict(hsd, strict=True)
    neck, head = neck.to(dev).eval(), head.to(dev).eval()
    g = torch.Generator().manual_seed(0)
    ins = [torch.randn(batch, c, H // s, W // s, generator=g).to(dev) for c, s in zip((256, 512, 1024, 2048), (4, 8, 16, 32))]
    metas = [dict(img_shape=(H, W - 11, 3), scale_factor=1.0)] * batch

    def step():
        feats = neck(ins)
        return head.detect(feats, metas)

    for _ in range(3):
        res = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        res = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"MP-Det FPN + MPHead + get_bboxes, batch {batch} at {H}x{W}: {ms:.2f} ms / batch = {batch / ms * 1e3:.1f} img/s "
          f"(detections per image: {[len(r[0]) for r in res][:4]} ...)")
    cpu = [t[:1].cpu() for t in ins]
    torch.set_num_threads(torch.get_num_threads())
    t0 = time.perf_counter()
    with torch.no_grad():
        outs = M.fpn_forward(nsd, cpu)
        cs, bp = M.mp_head_forward(hsd, outs)
        M.gfl_get_bboxes_single([c[0] for c in cs], [b[0] for b in bp], (H, W - 11))
    print(f"oracle port on the host CPU ({torch.get_num_threads()} threads), one image: {time.perf_counter() - t0:.2f} s")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
