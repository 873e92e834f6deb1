"""Developer tool: make the seeded synthetic weights behave like a trained network.

A random network without normalisation drifts (activations grow ~1.3x per layer), which gives absurd logits and
a degenerate NMS load.  Real checkpoints have BatchNorm statistics that match their activations; this tool gives
the synthetic ones the same property: it runs the CPU oracle once on seeded noise images, sets every BatchNorm's
running_mean / running_var to the statistics of the conv output feeding it (exactly what BN training would have
recorded), and rescales the prediction convs so that obj / cls logits have a standard deviation of 2.0 / 0.5 around the
YOLOX prior and the box regressors 0.15, and the objectness biases are shifted by one common offset so that a
target fraction of the anchors (default 4 %) passes conf 0.01.  Only those tensors are stored, as
glsdet_b200/data/calib_p0_<phi>_nc<nc>_seed<seed>.npz; glsdet_b200.synthetic.synthetic_state_dict(flavour=
"calibrated") overlays them on the seeded base weights.  (Uses oracle/, so it is a tool, not product code.)

    python tools/calibrate_synthetic.py --phi s --nc 10 --seed 0
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict  # noqa: E402
from oracle import ref_path  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--phi", default="s")
    ap.add_argument("--nc", type=int, default=10)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--variant", default="ffa", choices=["ffa", "stock", "p1", "p2"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--obj-std", type=float, default=2.0)
    ap.add_argument("--target-pass", type=float, default=0.04, help="fraction of anchors with obj*cls >= 0.01")
    ap.add_argument("--cls-std", type=float, default=0.5)
    ap.add_argument("--bn-beta", type=float, default=None,
                    help="mean of the BatchNorm shifts beta (std 0.1).  Default: keep the base weights' N(0, 0.1).  With beta "
                         "= 0 the pre-activations are N(0, 1), where SiLU + re-normalisation multiplies any relative "
                         "perturbation by 1.10 per layer (chaotic regime of a random net: sqrt(E[silu'(z)^2] Var z / Var "
                         "silu(z))): through the ~45 layers of YOLOX-l even fp16 rounding noise grows to 2-3 %.  A trained "
                         "network does not sit there; beta = 1 gives a factor of 1.03 per layer.")
    ap.add_argument("--out", default=None, help="output .npz (default glsdet_b200/data/calib_<variant>_<phi>_nc<nc>_seed<seed>.npz)")
    args = ap.parse_args()
    torch.set_num_threads(8)
    sd = synthetic_state_dict(args.nc, args.phi, seed=args.seed, flavour="kaiming", variant=args.variant)
    changed = {}
    if args.bn_beta is not None:
        g = torch.Generator().manual_seed(args.seed + 977)
        for k in list(sd):
            if k.endswith(".bn.bias"):
                sd[k] = args.bn_beta + 0.1 * torch.randn(sd[k].shape, generator=g)
                changed[k] = sd[k]
    orig = ref_path.base_conv

    def calibrating_base_conv(sd_, p, x, stride=1, act="silu"):
        if (p + ".dconv.conv.weight") in sd_:   # DWConv (phi = 'nano'): the oracle recurses into dconv / pconv
            return orig(sd_, p, x, stride, act)
        w = sd_[p + ".conv.weight"]
        y = F.conv2d(x, w, None, stride=stride, padding=(w.shape[-1] - 1) // 2, groups=x.shape[1] // w.shape[1])
        sd_[p + ".bn.running_mean"] = y.mean(dim=(0, 2, 3))
        sd_[p + ".bn.running_var"] = y.var(dim=(0, 2, 3), unbiased=False).clamp_min(1e-4)
        changed[p + ".bn.running_mean"] = sd_[p + ".bn.running_mean"]
        changed[p + ".bn.running_var"] = sd_[p + ".bn.running_var"]
        return orig(sd_, p, x, stride, act)

    ref_path.base_conv = calibrating_base_conv
    x = synthetic_images(2, args.size, args.size, seed=args.seed + 4242)
    with torch.no_grad():
        feats = ref_path.csp_darknet(sd, x)
        if args.variant == "p2":
            neck = [None] + ref_path.p2_neck(sd, feats[1:])
        else:
            neck = ref_path.p1_neck(sd, feats) if args.variant == "p1" else ref_path.pafpn_neck(sd, feats)
        # towers: rescale the prediction convs on the calibrated tower outputs
        towers = None
        if args.variant == "p1":   # models/new/yolox10.py:83-139
            f0 = ref_path.csp_layer(sd, "head.csp_feat0", neck[0])
            xs = [ref_path.base_conv(sd, f"head.stems.{k}", neck[k + 1]) for k in range(3)]
            towers = []
            for k, xk in enumerate(xs):
                lower = f0 if k == 0 else xs[k - 1]
                down = ref_path.base_conv(sd, f"head.up_convs.{k}.1", ref_path.base_conv(sd, f"head.up_convs.{k}.0", lower), stride=2)
                parts = [xk, down] + ([F.interpolate(xs[k + 1], scale_factor=2, mode="nearest")] if k < 2 else [])
                cf = ref_path.base_conv(sd, f"head.cls_convs.{k}.1", ref_path.base_conv(sd, f"head.cls_convs.{k}.0", torch.cat(parts, 1)))
                rf = ref_path.base_conv(sd, f"head.reg_convs.{k}.1", ref_path.base_conv(sd, f"head.reg_convs.{k}.0", xk))
                towers.append((k, cf, rf))
            proc = []
        elif args.variant in ("stock", "p2"):
            proc = [ref_path.base_conv(sd, f"head.stems.{k}", neck[k + 1]) for k in range(3)]
        else:
            zz = ref_path.ffa(sd, "head.ftt", neck[1], neck[2])
            proc = [ref_path.csp_layer(sd, "head.csp", neck[0]) + F.interpolate(zz, scale_factor=2, mode="nearest")]
            proc += [ref_path.base_conv(sd, f"head.stems.{k}", neck[k + 1]) for k in range(3)]
        if towers is None:
            towers = []
            for k, xk in enumerate(proc):
                i = k if args.variant in ("stock", "p2") else (3 if k == 0 else k - 1)
                cf = ref_path.base_conv(sd, f"head.cls_convs.{i}.1", ref_path.base_conv(sd, f"head.cls_convs.{i}.0", xk))
                rf = ref_path.base_conv(sd, f"head.reg_convs.{i}.1", ref_path.base_conv(sd, f"head.reg_convs.{i}.0", xk))
                towers.append((i, cf, rf))
        for i, cf, rf in towers:
            for name, feat, target in (("cls_preds", cf, args.cls_std), ("obj_preds", rf, args.obj_std), ("reg_preds", rf, 0.15)):
                key = f"head.{name}.{i}.weight"
                y = F.conv2d(feat, sd[key])
                std = y.std(dim=(0, 2, 3)).clamp_min(1e-6)
                sd[key] = sd[key] * (target / std).view(-1, 1, 1, 1)
                changed[key] = sd[key]
    ref_path.base_conv = orig
    # shift the objectness biases so that the wanted fraction of anchors passes conf 0.01 (bisection on one offset)
    head_fn = {"stock": lambda n: ref_path.stock_head(sd, n[1:]), "p2": lambda n: ref_path.stock_head(sd, n[1:]), "p1": lambda n: ref_path.p1_head(sd, n),
               "ffa": lambda n: ref_path.yolox_head(sd, n)}[args.variant]
    with torch.no_grad():
        lg = head_fn(neck)
    obj = torch.cat([l[:, 4].flatten(1) for l in lg], 1)
    cls = torch.cat([torch.sigmoid(l[:, 5:]).max(1)[0].flatten(1) for l in lg], 1)
    lo, hi = -20.0, 20.0
    for _ in range(50):
        mid = 0.5 * (lo + hi)
        frac = float(((torch.sigmoid(obj + mid) * cls) >= 0.01).float().mean())
        lo, hi = (mid, hi) if frac < args.target_pass else (lo, mid)
    for i in range(len(lg)):
        key = f"head.obj_preds.{i}.bias"
        sd[key] = sd[key] + 0.5 * (lo + hi)
        changed[key] = sd[key]
    tag = "p0" if args.variant == "ffa" else args.variant
    out = Path(args.out) if args.out else ROOT / "glsdet_b200" / "data" / f"calib_{tag}_{args.phi}_nc{args.nc}_seed{args.seed}.npz"
    np.savez_compressed(out, **{k: v.numpy().astype(np.float32) for k, v in changed.items()})
    # report
    with torch.no_grad():
        lg = head_fn(neck)
        pred = ref_path.decode_outputs(lg, [args.size, args.size])
    sc = pred[:, :, 4] * pred[:, :, 5:].max(2)[0]
    print(out.name, f"{out.stat().st_size/1e3:.0f} kB", "anchors", pred.shape[1], "cand@0.01", (sc >= 0.01).sum(1).tolist(),
          "neck std", [round(float(t.std()), 2) for t in neck if t is not None], "logit std", [round(float(l.std()), 2) for l in lg])
    res = ref_path.non_max_suppression(pred, args.nc, [args.size, args.size], None, False, 0.01, 0.65, correct_boxes=False)
    print("kept", [len(r) for r in res])


if __name__ == "__main__":
    main()
