"""Seeded random-init weights and inputs of the P0 architecture (models/ffa/yolox_ffa.py YoloBody).

Data generation only - shared by the tests, the golden-vector script and bench.py (there is no network access for
checkpoints or datasets, so every run uses synthetic weights of the reference's exact shapes and keys).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

StateDict = Dict[str, torch.Tensor]

def p0_state_dict_shapes(num_classes: int, phi: str, variant: str = "ffa") -> Dict[str, tuple]:
    """Shapes of every state_dict entry of models/ffa/yolox_ffa.py YoloBody(num_classes, phi)."""
    depth = {"nano": 0.33, "tiny": 0.33, "s": 0.33, "m": 0.67, "l": 1.0, "x": 1.33}[phi]
    width = {"nano": 0.25, "tiny": 0.375, "s": 0.50, "m": 0.75, "l": 1.0, "x": 1.25}[phi]
    depthwise = phi == "nano"   # yolox_ffa.py:270: every k > 1 conv outside Focus becomes a DWConv (baseConv.py:22-30)
    shapes: Dict[str, tuple] = {}

    def bc(p, cin, cout, k):
        shapes[p + ".conv.weight"] = (cout, cin, k, k)
        for n in ("weight", "bias", "running_mean", "running_var"):
            shapes[f"{p}.bn.{n}"] = (cout,)
        shapes[p + ".bn.num_batches_tracked"] = ()

    def cv(p, cin, cout, k):   # `Conv = DWConv if depthwise else BaseConv`
        if depthwise:
            bc(p + ".dconv", 1, cin, k)      # [cin, 1, k, k], groups = cin
            bc(p + ".pconv", cin, cout, 1)
        else:
            bc(p, cin, cout, k)

    def csp(p, cin, cout, n):
        hid = int(cout * 0.5)
        bc(p + ".conv1", cin, hid, 1)
        bc(p + ".conv2", cin, hid, 1)
        bc(p + ".conv3", 2 * hid, cout, 1)
        for j in range(n):
            bc(f"{p}.m.{j}.conv1", hid, hid, 1)
            cv(f"{p}.m.{j}.conv2", hid, hid, 3)

    base = int(width * 64)
    bdep = max(round(depth * 3), 1)
    bb = "backbone.backbone"
    bc(f"{bb}.stem.conv", 12, base, 3)
    for name, cin, cout, n in (("dark2", base, base * 2, bdep), ("dark3", base * 2, base * 4, bdep * 3),
                               ("dark4", base * 4, base * 8, bdep * 3)):
        cv(f"{bb}.{name}.0", cin, cout, 3)
        csp(f"{bb}.{name}.1", cout, cout, n)
    cv(f"{bb}.dark5.0", base * 8, base * 16, 3)
    bc(f"{bb}.dark5.1.conv1", base * 16, base * 8, 1)
    bc(f"{bb}.dark5.1.conv2", base * 32, base * 16, 1)
    csp(f"{bb}.dark5.2", base * 16, base * 16, bdep)

    c0, c1, c2 = int(256 * width), int(512 * width), int(1024 * width)
    n = round(3 * depth)
    p2 = variant == "p2"   # models/block/non_local/yolo_patch_nonlocal_plus.py: C3_p4 / C3_n3 take a third input
    bc("backbone.lateral_conv0", c2, c1, 1)
    csp("backbone.C3_p4", (3 if p2 else 2) * c1, c1, n)
    bc("backbone.reduce_conv1", c1, c0, 1)
    csp("backbone.C3_p3", 2 * c0, c0, n)
    if p2:
        shapes["backbone.P3_Identity.conv.weight"], shapes["backbone.P3_Identity.conv.bias"] = (c0, c0, 7, 7), (c0,)
    cv("backbone.bu_conv2", c0, c0, 3)
    csp("backbone.C3_n3", (3 if p2 else 2) * c0, c1, n)
    if p2:
        shapes["backbone.P4_Identity.conv.weight"], shapes["backbone.P4_Identity.conv.bias"] = (c1, c1, 5, 5), (c1,)
    cv("backbone.bu_conv1", c1, c1, 3)
    csp("backbone.C3_n4", 2 * c1, c2, n)

    if p2:
        for name, cin, cout, nl in (("Patch_conv_feat1", c0, c1, True), ("Patch_conv_feat2", c1, c0, False)):
            q, mid = f"backbone.{name}", cin // 2
            for pos in ("lt", "lb", "rt", "rb"):
                bc(f"{q}.feat_patchconv_{pos}", cin, mid, 3)
            if nl:
                for pos in ("lt", "lb", "rt", "rb"):
                    for nm in ("g", "theta", "phi", "conv_out"):
                        shapes[f"{q}.feat_patchconv_{pos}_nonlocal.{nm}.weight"] = (mid, mid, 1, 1)
                        shapes[f"{q}.feat_patchconv_{pos}_nonlocal.{nm}.bias"] = (mid,)
            for pos in ("r", "l", "t", "b"):
                bc(f"{q}.feat_patchconv_{pos}", mid, mid, 3)
            shapes[f"{q}.channel_conv.weight"], shapes[f"{q}.channel_conv.bias"] = (cout, 2 * mid, 1, 1), (cout,)
        shapes["backbone.P5_Identity.conv.weight"], shapes["backbone.P5_Identity.conv.bias"] = (c2, c2, 3, 3), (c2,)
        variant = "stock"   # the head is the stock three-level one
    hc = int(256 * width)
    if variant == "p1":   # models/new/yolox10.py: patch non-local blocks in the neck, cross-level cls branch (App. C)
        for i, c in ((1, c0), (2, c1), (3, c2)):
            q = f"backbone.Patch_conv_feat{i}"
            for pos in ("lt", "lb", "rt", "rb"):
                for name in ("g", "theta", "phi", "conv_out"):
                    shapes[f"{q}.feat_patchconv_{pos}_nonlocal.{name}.weight"] = (c, c, 1, 1)
                    shapes[f"{q}.feat_patchconv_{pos}_nonlocal.{name}.bias"] = (c,)
            bc(q + ".channel_conv", c, c, 3)
        csp("head.csp_feat0", int(0.5 * 256 * width), hc, round(3 * 0.75))
        for i, cin in enumerate((c0, c1, c2)):
            bc(f"head.stems.{i}", cin, hc, 1)
            cv(f"head.up_convs.{i}.0", hc, hc, 3)
            cv(f"head.up_convs.{i}.1", hc, hc, 3)
            m = 2 if i == 2 else 3
            cv(f"head.cls_convs.{i}.0", m * hc, m * hc, 3)
            cv(f"head.cls_convs.{i}.1", m * hc, hc, 3)
            cv(f"head.reg_convs.{i}.0", hc, hc, 3)
            cv(f"head.reg_convs.{i}.1", hc, hc, 3)
            for name, co in (("cls_preds", num_classes), ("reg_preds", 4), ("obj_preds", 1)):
                shapes[f"head.{name}.{i}.weight"] = (co, hc, 1, 1)
                shapes[f"head.{name}.{i}.bias"] = (co,)
        return shapes
    if variant == "stock":   # models/base/yolox.py: three levels, stems on (P3_out, P4_out, P5_out), no FFA
        for i, cin in enumerate((c0, c1, c2)):
            bc(f"head.stems.{i}", cin, hc, 1)
            for br in ("cls_convs", "reg_convs"):
                cv(f"head.{br}.{i}.0", hc, hc, 3)
                cv(f"head.{br}.{i}.1", hc, hc, 3)
            for name, co in (("cls_preds", num_classes), ("reg_preds", 4), ("obj_preds", 1)):
                shapes[f"head.{name}.{i}.weight"] = (co, hc, 1, 1)
                shapes[f"head.{name}.{i}.bias"] = (co,)
        return shapes
    csp("head.csp", int(0.5 * 256 * width), hc, round(3 * 0.75))
    f = "head.ftt"
    bc(f + ".scale", 2 * hc, 4 * hc, 1)
    bc(f + ".create_content_extractor.0", 4 * hc, 4 * hc, 1)
    bc(f + ".create_content_extractor.1", 4 * hc, 4 * hc, 1)
    bc(f + ".create_text_extractor.0", 2 * hc, 2 * hc, 1)
    bc(f + ".conv3", 2 * hc, hc, 1)
    shapes[f + ".se1.fc.0.weight"] = (4 * hc // 16, 4 * hc)
    shapes[f + ".se1.fc.2.weight"] = (4 * hc, 4 * hc // 16)
    for i, cin in enumerate((c0, c1, c2)):
        bc(f"head.stems.{i}", cin, hc, 1)
    for i in range(4):
        for br in ("cls_convs", "reg_convs"):
            cv(f"head.{br}.{i}.0", hc, hc, 3)
            cv(f"head.{br}.{i}.1", hc, hc, 3)
        for name, co in (("cls_preds", num_classes), ("reg_preds", 4), ("obj_preds", 1)):
            shapes[f"head.{name}.{i}.weight"] = (co, hc, 1, 1)
            shapes[f"head.{name}.{i}.bias"] = (co,)
    return shapes


def synthetic_state_dict(num_classes: int, phi: str, seed: int = 0, flavour: str = "kaiming",
                         variant: str = "ffa") -> StateDict:
    """Random-init weights for the P0 architecture.

    flavour "reference": what train.py does - default PyTorch init then weights_init(normal, 0.02)
        (models/ffa/yolox_losses.py:402-421): conv weights N(0, 0.02), BN gamma N(1, 0.02), beta 0, stats (0, 1).
        Activations collapse towards zero through the depth (gain < 1 per layer) and every anchor scores ~0.25, so
        it exercises the kernels poorly; kept as a parity case because it is the reference's own initialisation.
    flavour "calibrated": "kaiming" plus BatchNorm running statistics that match the activations (as in any trained
        checkpoint) and prediction convs scaled to logit noise 1.2 / box noise 0.3; recorded once by
        tools/calibrate_synthetic.py into glsdet_b200/data/.  Feature maps are O(1) at every depth, a few per cent
        of the anchors pass conf 0.01 and neighbouring candidates suppress each other.  Used by bench.py and the
        full-size parity tests.
    flavour "kaiming": variance-preserving init for SiLU/ReLU layers (E[silu(z)^2] = 0.356, E[relu(z)^2] = 0.5 for
        z ~ N(0,1)), mildly randomised BN parameters and statistics, prediction biases at the YOLOX prior
        (-log((1-p)/p), p = 0.01), logit noise of about 1.2 and boxes about five cells wide, so that feature maps
        stay O(1) at every depth, 1-3 % of the anchors pass conf 0.01 and neighbouring candidates really suppress
        each other.  Used for parity at full depth and for the benchmark.
    """
    if flavour == "calibrated":
        # seeded base weights + BatchNorm statistics / prediction-conv scales recorded by tools/calibrate_synthetic.py
        import numpy as np
        from pathlib import Path

        tag = "p0" if variant == "ffa" else variant
        f = Path(__file__).resolve().parent / "data" / f"calib_{tag}_{phi}_nc{num_classes}_seed{seed}.npz"
        if not f.exists():
            raise FileNotFoundError(f"{f} missing: run tools/calibrate_synthetic.py --phi {phi} --nc {num_classes} "
                                    f"--seed {seed} --variant {variant}")
        sd = synthetic_state_dict(num_classes, phi, seed, "kaiming", variant)
        with np.load(f) as z:
            for k in z.files:
                assert tuple(z[k].shape) == tuple(sd[k].shape), k
                sd[k] = torch.from_numpy(z[k].copy())
        return sd
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    relu_layers = ("head.ftt.",)
    for k, shp in p0_state_dict_shapes(num_classes, phi, variant).items():
        part = k.rsplit(".", 1)[1]
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith(".conv.weight"):
            fan_in = shp[1] * shp[2] * shp[3]
            if flavour == "reference":
                std = 0.02
            else:
                second_moment = 0.5 if k.startswith(relu_layers) else 0.356
                std = math.sqrt(1.0 / (fan_in * second_moment))
            sd[k] = torch.randn(shp, generator=g) * std
        elif ".bn." in k:
            if flavour == "reference":
                sd[k] = {"weight": 1.0 + 0.02 * torch.randn(shp, generator=g), "bias": torch.zeros(shp),
                         "running_mean": torch.zeros(shp), "running_var": torch.ones(shp)}[part]
            else:
                sd[k] = {"weight": 1.0 + 0.05 * torch.randn(shp, generator=g),
                         "bias": 0.1 * torch.randn(shp, generator=g),
                         "running_mean": 0.1 * torch.randn(shp, generator=g),
                         "running_var": 0.9 + 0.2 * torch.rand(shp, generator=g)}[part]
        elif ".fc." in k:
            sd[k] = torch.randn(shp, generator=g) * math.sqrt(1.0 / shp[1])
        elif ".channel_conv." in k and ".conv." not in k and ".bn." not in k or "_Identity.conv." in k:
            # plain nn.Conv2d with bias (Identity_Conv.py:287-288 and :27-84); identity convs start near identity
            if part == "weight":
                fan_in = shp[1] * shp[2] * shp[3]
                wgt = torch.randn(shp, generator=g) * (0.02 if flavour == "reference" else 0.5 * math.sqrt(1.0 / fan_in))
                if "_Identity" in k:
                    idx = torch.arange(shp[0])
                    wgt[idx, idx, shp[2] // 2, shp[3] // 2] += 1.0
                sd[k] = wgt
            else:
                sd[k] = torch.randn(shp, generator=g) * (0.02 if flavour == "reference" else 0.1)
        elif "_nonlocal." in k:   # plain nn.Conv2d 1x1 with bias (Non_local_family.py:15-18)
            if part == "weight":
                sd[k] = torch.randn(shp, generator=g) * (0.02 if flavour == "reference" else math.sqrt(1.0 / shp[1]))
            else:
                sd[k] = torch.randn(shp, generator=g) * (0.02 if flavour == "reference" else 0.1)
        elif "_preds." in k and part == "weight":
            if flavour == "reference":
                sd[k] = torch.randn(shp, generator=g) * 0.02
            else:
                std = 1.2 / math.sqrt(shp[1] * 0.356)
                if "reg_preds" in k:
                    std *= 0.25
                sd[k] = torch.randn(shp, generator=g) * std
        elif "_preds." in k and part == "bias":
            if flavour == "reference":
                sd[k] = (torch.rand(shp, generator=g) * 2 - 1) * 0.05
            elif "reg_preds" in k:      # boxes about ten cells wide: neighbouring anchors overlap above 0.65
                sd[k] = torch.tensor([0.0, 0.0, 2.3, 2.3]) + 0.1 * torch.randn(shp, generator=g)
            elif "cls_preds" in k:      # a few dominant classes, as in VisDrone (car / pedestrian)
                sd[k] = torch.full(shp, -math.log((1 - 0.01) / 0.01)) + 1.0 * torch.randn(shp, generator=g)
            else:
                sd[k] = torch.full(shp, -math.log((1 - 0.01) / 0.01)) + 0.2 * torch.randn(shp, generator=g)
        else:
            raise KeyError(k)
    return sd


def synthetic_images(batch: int, height: int, width: int, seed: int = 0) -> torch.Tensor:
    """Seeded stand-in for ImageNet-normalised aerial images: unit-variance noise with a natural-image-like
    spectrum (white noise plus bilinearly upsampled noise at 1/4, 1/16 and 1/64 resolution), so that feature maps
    and score maps have spatial structure (clusters of candidates) instead of being white."""
    import torch.nn.functional as F

    g = torch.Generator().manual_seed(seed)
    x = 0.35 * torch.randn(batch, 3, height, width, generator=g)
    for s, a in ((4, 0.6), (16, 0.8), (64, 1.0)):
        low = torch.randn(batch, 3, max(2, height // s), max(2, width // s), generator=g)
        x = x + a * F.interpolate(low, size=(height, width), mode="bilinear", align_corners=False)
    return x / x.std()
