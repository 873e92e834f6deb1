"""Measured relative l2 error of the native 16-bit path against the fp32 oracle, per neck map and logit level, for the
configurations BASELINE.json / VERDICT name (run on a GPU box; output committed under profiles/).
Columns: error of the CUDA path vs the fp32 oracle | error of the storage-precision emulation vs the fp32 oracle (what
16-bit storage alone costs) | CUDA path vs the emulation (kernel error)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict  # noqa: E402
from oracle import ref_path  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm()).item()


def case(variant, phi, nc, seed, h, w, dev, from_image=False):
    if variant == "p0":
        from glsdet_b200.yolox_ffa import YoloBody
        vname = "ffa"
    elif variant == "p1":
        from glsdet_b200.yolox10 import YoloBody
        vname = "p1"
    else:
        from glsdet_b200.yolo_patch_nonlocal_plus import YoloBody
        vname = "p2"
    sd = synthetic_state_dict(nc, phi, seed=seed, flavour="calibrated", variant=vname)
    net = YoloBody(nc, phi)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    torch.manual_seed(seed + 100)
    if from_image:
        x = synthetic_images(1, h, w, seed=seed + 5)
        feats = ref_path.csp_darknet(sd, x)
        emu_feats = ref_path.csp_darknet_bf16(sd, x)
        if variant == "p2":
            feats, emu_feats = feats[1:], emu_feats[1:]
        got = net(x.to(dev))
    else:
        x = synthetic_images(1, h, w, seed=seed + 5)
        feats = ref_path.csp_darknet(sd, x)       # realistic feature statistics, exact fp32 features as the input
        if variant == "p2":
            feats = feats[1:]
        emu_feats = feats
        got = net.forward_features([f.to(dev) for f in feats])
    fn = {"p0": (ref_path.neck_head, lambda f: ref_path.neck_head_bf16(sd, f)),
          "p1": (lambda s_, f: ref_path.p1_neck_head(s_, f), lambda f: ref_path.p1_neck_head(sd, f, bf16=True)),
          "p2": (lambda s_, f: ref_path.p2_neck_head(s_, f), lambda f: ref_path.p2_neck_head(sd, f, bf16=True))}[variant]
    ref = fn[0](sd, feats)
    emu = fn[1](emu_feats)
    tag = f"{variant}-{phi} {h}x{w} nc={nc} {'image->logits' if from_image else 'features->logits'}"
    for i, (g, r, e) in enumerate(zip(got, ref, emu)):
        print(f"{tag:48s} logits{i}: cuda-vs-fp32 {rel(g, r):.4f}   storage-emulation-vs-fp32 {rel(e, r):.4f}   "
              f"cuda-vs-emulation {rel(g, e):.4f}", flush=True)
    if variant == "p0" and not from_image:
        neck = net.backbone.forward_features([f.to(dev) for f in feats])
        nref = ref_path.pafpn_neck(sd, feats)
        for i in range(1, 4):
            print(f"{tag:48s} neck{i}:   cuda-vs-fp32 {rel(neck[i], nref[i]):.4f}", flush=True)


def main():
    dev = torch.device("cuda", 0)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    case("p0", "s", 10, 0, 1024, 1024, dev)
    case("p0", "s", 10, 0, 1024, 1024, dev, from_image=True)
    case("p0", "l", 3, 6, 544, 1024, dev)   # seed 6 = tests/test_path_gpu.py::test_config4_yolox_l_544x1024
    case("p1", "s", 10, 0, 1024, 1024, dev)
    case("p1", "l", 3, 6, 544, 1024, dev)
    case("p2", "s", 10, 0, 1024, 1024, dev)


if __name__ == "__main__":
    main()
