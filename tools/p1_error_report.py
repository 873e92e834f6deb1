"""Developer tool: per-stage error of the P1 path at 1024^2 against the fp32 oracle and its bf16-storage emulation."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from glsdet_b200.synthetic import synthetic_images  # noqa: E402
from glsdet_b200.yolox10 import YoloBody  # noqa: E402
from oracle import ref_path  # noqa: E402


def rl2(a, b):
    return ((a.float().cpu() - b).norm() / b.norm()).item()


def main(size=1024):
    dev = torch.device("cuda:0")
    sd = ref_path.synthetic_state_dict(10, "s", seed=0, flavour="calibrated", variant="p1")
    net = YoloBody(10, "s")
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    feats = ref_path.csp_darknet(sd, synthetic_images(1, size, size, seed=12))
    d = [f.to(dev) for f in feats]
    plan = net.plan_for(d)
    plan.load_features(d)
    plan.run_neck()
    torch.cuda.synchronize()
    with torch.no_grad():
        for i in (1, 2, 3):
            p = f"backbone.Patch_conv_feat{i}"
            ref = feats[i] + ref_path.patch_conv_nonlocal_new(sd, p, feats[i])
            ref_path._EMULATE_BF16 = True
            emu = ref_path._q(ref_path._q(feats[i]) + ref_path.patch_conv_nonlocal_new(sd, p, ref_path._q(feats[i])))
            ref_path._EMULATE_BF16 = False
            got = plan.buffer(f"feat{i}").float().permute(0, 3, 1, 2)
            print(f"feat{i}: gpu vs ref {rl2(got, ref):.4f}  emu vs ref {rl2(emu, ref):.4f}  gpu vs emu {rl2(got, emu):.4f}  "
                  f"|patchconv|/|feat| {(ref - feats[i]).norm() / feats[i].norm():.3f}")
        neck = ref_path.p1_neck(sd, feats)
        ref_path._EMULATE_BF16 = True
        neck_e = ref_path.p1_neck(sd, [ref_path._q(f) for f in feats])
        ref_path._EMULATE_BF16 = False
        outs = plan.neck_outputs_nchw()
        for i in (1, 2, 3):
            print(f"neck{i}: gpu vs ref {rl2(outs[i], neck[i]):.4f}  emu vs ref {rl2(neck_e[i], neck[i]):.4f}  gpu vs emu {rl2(outs[i], neck_e[i]):.4f}")
        lg = net.forward_features(d)
        ref = ref_path.p1_neck_head(sd, feats)
        emu = ref_path.p1_neck_head(sd, feats, bf16=True)
        for i in range(3):
            print(f"logits{i}: gpu vs ref {rl2(lg[i], ref[i]):.4f}  emu vs ref {rl2(emu[i], ref[i]):.4f}  gpu vs emu {rl2(lg[i], emu[i]):.4f}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1024)
