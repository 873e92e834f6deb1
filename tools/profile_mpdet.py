"""One profiled step of the MP-Det workload (bench.py --config cfg3), bracketed by cudaProfilerStart/Stop."""
import sys
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
    neck.load_state_dict({k[5:]: v for k, v in sd.items() if k.startswith("neck.")}, strict=True)
    head.load_state_dict({k[10:]: v for k, v in sd.items() if k.startswith("bbox_head.")}, strict=True)
    neck, head = neck.to(dev).eval(), head.to(dev).eval()
    g = torch.Generator().manual_seed(0)
    ins = [torch.randn(batch, c, H // s, W // s, generator=g).to(dev) for c, s in zip((256, 512, 1024, 2048), (4, 8, 16, 32))]
    metas = [dict(img_shape=(H, W - 11, 3), scale_factor=1.0)] * batch
    for _ in range(3):
        head.detect(neck(ins), metas)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    head.detect(neck(ins), metas)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
