"""Generate the golden vectors under tests/golden/ by running the REAL reference (read-only at /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference's own test-suite pins nothing on this path, so these vectors - outputs of the
reference modules themselves on seeded inputs - are what pins oracle/ref_path.py and, through it, the CUDA path.
Import-time shims only: D1 `yolox_ffa.FTT = ffa.FFA`; stdout is swallowed because the
reference prints inside forward (D3).  Weights come from oracle.ref_path.synthetic_state_dict (seeded), are
loaded with strict=True (which also pins the state_dict key/shape table) and are NOT stored: a checksum is.
"""
import contextlib
import io
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.dont_write_bytecode = True
REF = "/root/reference/yolox-drone"


def import_reference():
    sys.path.insert(0, REF)
    import models.ffa.ffa as ffa_mod
    import models.ffa.yolox_ffa as yf

    yf.FTT = ffa_mod.FFA  # D1
    from models.core import utils_bbox as ub

    return yf, ub


def weight_checksum(sd):
    tot = 0.0
    for k in sorted(sd):
        if sd[k].dtype.is_floating_point:
            tot += float(sd[k].double().abs().sum())
    return tot


def seeded_feats(seed, batch, in_h, in_w, chans):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(batch, c, in_h // s, in_w // s, generator=g) for c, s in zip(chans, (4, 8, 16, 32))]


def run_model_case(yf, ub, name, phi, nc, flavour, seed, batch, in_h, in_w, conf, nms_thr):
    from oracle import ref_path

    sd = ref_path.synthetic_state_dict(nc, phi, seed=seed, flavour=flavour)
    with contextlib.redirect_stdout(io.StringIO()):
        net = yf.YoloBody(nc, phi).eval()
    net.load_state_dict(sd, strict=True)
    key_shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    width = {"tiny": 0.375, "s": 0.5, "m": 0.75, "l": 1.0}[phi]
    chans = [int(c * width) for c in (128, 256, 512, 1024)]
    feats = seeded_feats(seed + 100, batch, in_h, in_w, chans)

    class Stub(torch.nn.Module):
        def forward(self, x):
            return dict(zip(("dark2", "dark3", "dark4", "dark5"), feats))

    net.backbone.backbone = Stub()
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        neck_out = net.backbone(torch.zeros(batch, 3, in_h, in_w))
        logits = net.head(neck_out)
        pred = ub.decode_outputs([o.clone() for o in logits], [in_h, in_w]).contiguous()
        results = ub.non_max_suppression(pred.clone(), nc, [in_h, in_w], np.array([in_h * 2, in_w * 3]), True,
                                         conf_thres=conf, nms_thres=nms_thr)
    out = {f"feat{i}": f.numpy() for i, f in enumerate(feats)}
    out.update({f"neck{i}": t.numpy() for i, t in enumerate(neck_out)})
    out.update({f"logits{i}": t.numpy() for i, t in enumerate(logits)})
    out["pred"] = pred.numpy()
    for i, r in enumerate(results):
        out[f"nms{i}"] = r if r is not None else np.zeros((0, 7), np.float32)
    meta = dict(name=name, phi=phi, nc=nc, flavour=flavour, seed=seed, batch=batch, in_h=in_h, in_w=in_w, conf=conf,
                nms_thr=nms_thr, image_shape=[in_h * 2, in_w * 3], letterbox=True, weight_checksum=weight_checksum(sd),
                n_keys=len(sd), candidates=[int((pred[b, :, 4] * pred[b, :, 5:].max(1)[0] >= conf).sum()) for b in range(batch)],
                kept=[int(len(out[f"nms{i}"])) for i in range(batch)])
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(name, {k: meta[k] for k in ("candidates", "kept", "weight_checksum")})
    return meta, key_shapes


def run_stock_case(ub, name, phi, nc, seed, batch, in_h, in_w, conf, nms_thr):
    """Stock 3-level YOLOX (models/base/yolox.py, imports as shipped): the math the mmdet YOLOXPAFPN / YOLOXHead
    pair also computes (key map in SURVEY.md section 8c)."""
    import models.base.yolox as yb
    from oracle import ref_path

    sd = ref_path.synthetic_state_dict(nc, phi, seed=seed, flavour="calibrated", variant="stock")
    net = yb.YoloBody(nc, phi).eval()
    net.load_state_dict(sd, strict=True)
    key_shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    width = {"tiny": 0.375, "s": 0.5, "m": 0.75, "l": 1.0}[phi]
    chans = [int(c * width) for c in (256, 512, 1024)]
    g = torch.Generator().manual_seed(seed + 100)
    feats = [torch.randn(batch, c, in_h // s, in_w // s, generator=g) for c, s in zip(chans, (8, 16, 32))]

    class Stub(torch.nn.Module):
        def forward(self, x):
            return dict(zip(("dark3", "dark4", "dark5"), feats))

    net.backbone.backbone = Stub()
    with torch.no_grad():
        neck_out = net.backbone(torch.zeros(batch, 3, in_h, in_w))
        logits = net.head(neck_out)
        pred = ub.decode_outputs([o.clone() for o in logits], [in_h, in_w]).contiguous()
        results = ub.non_max_suppression(pred.clone(), nc, [in_h, in_w], np.array([in_h, in_w]), False,
                                         conf_thres=conf, nms_thres=nms_thr)
    out = {f"feat{i}": f.numpy() for i, f in enumerate(feats)}
    out.update({f"neck{i}": t.numpy() for i, t in enumerate(neck_out)})
    out.update({f"logits{i}": t.numpy() for i, t in enumerate(logits)})
    out["pred"] = pred.numpy()
    for i, r in enumerate(results):
        out[f"nms{i}"] = r if r is not None else np.zeros((0, 7), np.float32)
    meta = dict(name=name, phi=phi, nc=nc, seed=seed, batch=batch, in_h=in_h, in_w=in_w, conf=conf, nms_thr=nms_thr,
                weight_checksum=weight_checksum(sd), n_keys=len(sd), kept=[int(len(out[f"nms{i}"])) for i in range(batch)])
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(name, meta["kept"], meta["weight_checksum"])
    return meta, key_shapes


def run_p1_case(ub, name, phi, nc, seed, batch, in_h, in_w, conf, nms_thr):
    """GLSDet P1 (models/new/yolox10.py: patch non-local neck + cross-level head).  Import shim D2 (SURVEY.md section
    0.2): yolox10.py imports `models.decouple.darknet`, which does not exist; `models.new.darknet` is the identical
    file (same bytes as models/ffa/darknet.py) and is aliased to that name."""
    import types

    import models.new.darknet as dk

    sys.modules.setdefault("models.decouple", types.ModuleType("models.decouple"))
    sys.modules["models.decouple.darknet"] = dk
    import models.new.yolox10 as y10
    from oracle import ref_path

    sd = ref_path.synthetic_state_dict(nc, phi, seed=seed, flavour="calibrated", variant="p1")
    with contextlib.redirect_stdout(io.StringIO()):
        net = y10.YoloBody(nc, phi).eval()
    net.load_state_dict(sd, strict=True)
    key_shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    width = {"tiny": 0.375, "s": 0.5, "m": 0.75, "l": 1.0}[phi]
    chans = [int(c * width) for c in (128, 256, 512, 1024)]
    feats = seeded_feats(seed + 100, batch, in_h, in_w, chans)

    class Stub(torch.nn.Module):
        def forward(self, x):
            return dict(zip(("dark2", "dark3", "dark4", "dark5"), feats))

    net.backbone.backbone = Stub()
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        neck_out = net.backbone(torch.zeros(batch, 3, in_h, in_w))
        logits = net.head(neck_out)
        pred = ub.decode_outputs([o.clone() for o in logits], [in_h, in_w]).contiguous()
        results = ub.non_max_suppression(pred.clone(), nc, [in_h, in_w], np.array([in_h, in_w]), False,
                                         conf_thres=conf, nms_thres=nms_thr)
        # one stand-alone non-local block and patch module, for the unit-level pins
        pc = net.backbone.Patch_conv_feat2
        nl_lt = pc.feat_patchconv_lt_nonlocal(feats[2][:, :, :feats[2].shape[2] // 2, :feats[2].shape[3] // 2])
        pc_out = pc(feats[2])
    out = {f"feat{i}": f.numpy() for i, f in enumerate(feats)}
    out.update({f"neck{i}": t.numpy() for i, t in enumerate(neck_out)})
    out.update({f"logits{i}": t.numpy() for i, t in enumerate(logits)})
    out["pred"] = pred.numpy()
    out["nonlocal_lt_feat2"] = nl_lt.numpy()
    out["patchconv_feat2"] = pc_out.numpy()
    for i, r in enumerate(results):
        out[f"nms{i}"] = r if r is not None else np.zeros((0, 7), np.float32)
    meta = dict(name=name, phi=phi, nc=nc, seed=seed, batch=batch, in_h=in_h, in_w=in_w, conf=conf, nms_thr=nms_thr,
                weight_checksum=weight_checksum(sd), n_keys=len(sd), kept=[int(len(out[f"nms{i}"])) for i in range(batch)])
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(name, meta["kept"], meta["weight_checksum"])
    return meta, key_shapes


def run_p2_case(ub, name, phi, nc, seed, batch, in_h, in_w, conf, nms_thr):
    """GLSDet P2 (models/block/non_local/yolo_patch_nonlocal_plus.py, imports as shipped)."""
    import models.block.non_local.yolo_patch_nonlocal_plus as yp
    from oracle import ref_path

    sd = ref_path.synthetic_state_dict(nc, phi, seed=seed, flavour="calibrated", variant="p2")
    with contextlib.redirect_stdout(io.StringIO()):
        net = yp.YoloBody(nc, phi).eval()
    net.load_state_dict(sd, strict=True)
    key_shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    width = {"tiny": 0.375, "s": 0.5, "m": 0.75, "l": 1.0}[phi]
    chans = [int(c * width) for c in (256, 512, 1024)]
    g = torch.Generator().manual_seed(seed + 100)
    feats = [torch.randn(batch, c, in_h // s, in_w // s, generator=g) for c, s in zip(chans, (8, 16, 32))]

    class Stub(torch.nn.Module):
        def forward(self, x):
            return dict(zip(("dark3", "dark4", "dark5"), feats))

    net.backbone.backbone = Stub()
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        neck_out = net.backbone(torch.zeros(batch, 3, in_h, in_w))
        logits = net.head(neck_out)
        pred = ub.decode_outputs([o.clone() for o in logits], [in_h, in_w]).contiguous()
        results = ub.non_max_suppression(pred.clone(), nc, [in_h, in_w], np.array([in_h, in_w]), False,
                                         conf_thres=conf, nms_thres=nms_thr)
        f1p = net.backbone.Patch_conv_feat1(feats[0])
        f2p = net.backbone.Patch_conv_feat2(feats[1])
    out = {f"feat{i}": f.numpy() for i, f in enumerate(feats)}
    out.update({f"neck{i}": t.numpy() for i, t in enumerate(neck_out)})
    out.update({f"logits{i}": t.numpy() for i, t in enumerate(logits)})
    out["pred"] = pred.numpy()
    out["feat1_patch"], out["feat2_patch"] = f1p.numpy(), f2p.numpy()
    for i, r in enumerate(results):
        out[f"nms{i}"] = r if r is not None else np.zeros((0, 7), np.float32)
    meta = dict(name=name, phi=phi, nc=nc, seed=seed, batch=batch, in_h=in_h, in_w=in_w, conf=conf, nms_thr=nms_thr,
                weight_checksum=weight_checksum(sd), n_keys=len(sd), kept=[int(len(out[f"nms{i}"])) for i in range(batch)])
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(name, meta["kept"], meta["weight_checksum"])
    return meta, key_shapes


def clustered_boxes(rng, k, nc, normalised=True, ties=False):
    g = max(4, k // 12)
    cen = rng.uniform(0.05, 0.95, (g, 2))
    which = rng.integers(0, g, k)
    c = cen[which] + rng.normal(0, 0.004, (k, 2))
    wh = np.exp(rng.normal(-3.2, 0.4, (g, 2)))[which] * np.exp(rng.normal(0, 0.15, (k, 2)))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1)
    if not normalised:
        boxes = boxes * 1024.0
    scores = (rng.permutation(k) + 1) / (k + 1.0)
    if ties:
        scores = np.round(scores, 2)
    labels = rng.integers(0, nc, k)
    return boxes.astype(np.float32), scores.astype(np.float32), labels.astype(np.float32)


def nms_cases():
    import torchvision
    from torchvision.ops import boxes as tvb

    rng = np.random.default_rng(2024)
    out, meta = {}, []
    specs = [("small_trick", 300, 10, True, False), ("mid_trick", 950, 10, True, False),
             ("mid_vanilla", 1800, 10, True, False), ("pixels", 700, 3, False, False),
             ("one_class", 500, 1, True, False), ("ties_trick", 400, 10, True, True), ("single", 1, 10, True, False)]
    for name, k, nc, norm, ties in specs:
        b, s, l = clustered_boxes(rng, k, nc, norm, ties)
        tb, ts, tl = torch.from_numpy(b), torch.from_numpy(s), torch.from_numpy(l)
        for thr in (0.45, 0.65):
            kt = tvb._batched_nms_coordinate_trick(tb, ts, tl, thr).numpy()
            kv = tvb._batched_nms_vanilla(tb, ts, tl, thr).numpy()
            ka = tvb.batched_nms(tb, ts, tl, thr).numpy()
            out[f"{name}_keep_trick_{thr}"] = kt
            out[f"{name}_keep_vanilla_{thr}"] = kv
            out[f"{name}_keep_auto_{thr}"] = ka
        out[f"{name}_boxes"], out[f"{name}_scores"], out[f"{name}_labels"] = b, s, l
        meta.append(dict(name=name, k=k, nc=nc, ties=ties))
    np.savez_compressed(HERE / "nms_torchvision.npz", **out)
    return dict(torchvision=torchvision.__version__, cases=meta)


def postproc_case(ub):
    """Reference decode_outputs + non_max_suppression on synthetic logits with real overlaps."""
    rng = np.random.default_rng(7)
    nc, in_h, in_w, batch = 10, 128, 160, 3
    g = torch.Generator().manual_seed(99)
    logits = []
    for s in (4, 8, 16, 32):
        h, w = in_h // s, in_w // s
        t = torch.randn(batch, 5 + nc, h, w, generator=g)
        t[:, 2:4] = t[:, 2:4] * 0.5 + 1.0          # boxes a few cells wide -> neighbours overlap
        t[:, 4] = t[:, 4] * 2.0 - 1.0
        logits.append(t)
    with torch.no_grad():
        pred = ub.decode_outputs([o.clone() for o in logits], [in_h, in_w]).contiguous()
        res = ub.non_max_suppression(pred.clone(), nc, [in_h, in_w], np.array([300, 500]), False, conf_thres=0.3,
                                     nms_thres=0.5)
        pred[2, :, 4] = 0.0  # an image without candidates -> empty [0,7]
        res2 = ub.non_max_suppression(pred.clone(), nc, [in_h, in_w], np.array([300, 500]), False, conf_thres=0.3,
                                      nms_thres=0.5)
    out = {f"logits{i}": t.numpy() for i, t in enumerate(logits)}
    out["pred"] = pred.numpy()  # with image 2 zeroed
    for i in range(batch):
        out[f"nms{i}"] = res2[i] if res2[i] is not None else np.zeros((0, 7), np.float32)
    np.savez_compressed(HERE / "postproc_reference.npz", **out)
    meta = dict(nc=nc, in_h=in_h, in_w=in_w, batch=batch, conf=0.3, nms_thr=0.5, image_shape=[300, 500],
                letterbox=False, kept=[int(len(out[f"nms{i}"])) for i in range(batch)])
    print("postproc", meta["kept"])
    return meta


def main():
    os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
    torch.set_num_threads(8)
    yf, ub = import_reference()
    if "--only-p1" in sys.argv:   # add / refresh the P1 vectors without touching the other files
        metas = json.loads((HERE / "meta.json").read_text())
        mp, keys_p1 = run_p1_case(ub, "p1_s_calibrated", "s", 10, 0, 1, 128, 192, 0.01, 0.65)
        metas["p1"] = mp
        (HERE / "state_dict_keys_p1_s.json").write_text(json.dumps(keys_p1, indent=0))
        (HERE / "meta.json").write_text(json.dumps(metas, indent=1))
        return
    if "--only-p2" in sys.argv:
        metas = json.loads((HERE / "meta.json").read_text())
        mp, keys_p2 = run_p2_case(ub, "p2_s_calibrated", "s", 10, 0, 1, 128, 192, 0.01, 0.65)
        metas["p2"] = mp
        (HERE / "state_dict_keys_p2_s.json").write_text(json.dumps(keys_p2, indent=0))
        (HERE / "meta.json").write_text(json.dumps(metas, indent=1))
        return
    metas = {}
    m1, keys = run_model_case(yf, ub, "p0_s_calibrated", "s", 10, "calibrated", 1, 2, 64, 96, 0.01, 0.65)
    m2, _ = run_model_case(yf, ub, "p0_s_refinit", "s", 10, "reference", 2, 1, 64, 64, 0.01, 0.65)
    m3, keys_t = run_model_case(yf, ub, "p0_tiny_calibrated", "tiny", 3, "calibrated", 3, 1, 96, 64, 0.01, 0.65)
    metas["models"] = [m1, m2, m3]
    ms, keys_stock = run_stock_case(ub, "stock_s_calibrated", "s", 10, 21, 2, 96, 64, 0.01, 0.65)
    metas["stock"] = ms
    (HERE / "state_dict_keys_stock_s.json").write_text(json.dumps(keys_stock, indent=0))
    mp, keys_p1 = run_p1_case(ub, "p1_s_calibrated", "s", 10, 0, 1, 128, 192, 0.01, 0.65)
    metas["p1"] = mp
    (HERE / "state_dict_keys_p1_s.json").write_text(json.dumps(keys_p1, indent=0))
    mp2, keys_p2 = run_p2_case(ub, "p2_s_calibrated", "s", 10, 0, 1, 128, 192, 0.01, 0.65)
    metas["p2"] = mp2
    (HERE / "state_dict_keys_p2_s.json").write_text(json.dumps(keys_p2, indent=0))
    metas["nms"] = nms_cases()
    metas["postproc"] = postproc_case(ub)
    metas["torch"] = torch.__version__
    (HERE / "meta.json").write_text(json.dumps(metas, indent=1))
    (HERE / "state_dict_keys_p0_s.json").write_text(json.dumps(keys, indent=0))
    print("wrote", sorted(p.name for p in HERE.glob("*.npz")))


if __name__ == "__main__":
    main()
