#!/usr/bin/env python
"""Benchmark of the GLSDet hot path (neck + FFA + decoupled head + decode + score filter + class-aware NMS).

    python bench.py --gpus 1 --steps 20 --warmup 5                      # product arm (hand-written sm_100a kernels)
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1      # the reference path on the host CPU
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU, images sharded

Workload = BASELINE.json configs[1]: GLSDet YOLOX-s (models/ffa/yolox_ffa.py YoloBody(10, 's')), a batch of 16
synthetic 1024x1024 VisDrone-shaped images PER GPU (weak scaling), bf16 activations / fp32 accumulation.
A step = one pass of the hot path over that batch: layout conversion of the four backbone feature maps (NCHW fp32,
what the reference's CSPDarknet hands to the neck) -> neck -> FFA -> head -> fused decode -> score filter
(conf 0.01, yolo.py:44) -> class-aware NMS (iou 0.65, yolo.py:48) -> [K,7] detections; for N > 1 ranks the
per-step NCCL gather of the detections is part of the step.  Inputs are synthetic: random-init weights of the
reference's exact shapes (glsdet_b200/synthetic.py) and feature maps produced once by the CSPDarknet backbone from
seeded noise images; the backbone is upstream of the metric and is not part of `value` or of the roofline segment.
`e2e` goes through the public module API with host buffers: YoloBody.detect(pinned host image batch) -> detections on
the host, i.e. it additionally runs the native CSPDarknet backbone in front of the path;
`e2e.from_features` is the same with host feature maps (the metric's exact segment, PCIe-bound on 503 MB per step).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

NUM_CLASSES = 10
PHI = "s"
IN_H = IN_W = 1024
CONF_THRES = 0.01     # yolox-drone/yolo.py:44
NMS_THRES = 0.65      # yolox-drone/yolo.py:48
ALGO_GFLOP_PER_IMAGE = 140.2   # neck 13.7 + FFA 8.6 + head 117.9 (2*MAC, reference graph)
WEIGHT_SEED = 0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU per step")
    ap.add_argument("--max-det", type=int, default=0,
                    help="0 (default) = every NMS survivor, like the reference's non_max_suppression (utils_bbox.py:414-420); "
                         "N > 0 = keep the N best per image (BASELINE configs[3] semantics; a cheaper top-K NMS)")
    ap.add_argument("--gather-rows", type=int, default=16384,
                    help="rows per image of the fixed-layout buffers of the uncapped run (host copy-out and NCCL gather); "
                         "the run aborts if an image keeps more")
    ap.add_argument("--cpu-images", type=int, default=16, help="images per pass of the bounded CPU-baseline sample")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline sample: repeat passes for about this long")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", default="p0", choices=["p0", "p1", "p2"],
                    help="p0 = BASELINE configs[1] (models/ffa/yolox_ffa.py, the default and the judged metric); "
                         "p1 = models/new/yolox10.py (patch non-local attention neck, SURVEY.md section 8d row 2')")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def conv_traffic(batch):
    """DRAM bytes (read + write) of the conv segment of one step, from the committed ncu capture (same batch only)."""
    f = ROOT / "profiles" / "r1_conv_traffic.json"
    if f.exists():
        d = json.loads(f.read_text())
        if d.get("batch") == batch:
            return d["total_bytes_per_step"]
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.005)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ synthetic data
VARIANT = "p0"


def make_weights():
    from glsdet_b200.synthetic import synthetic_state_dict

    return synthetic_state_dict(NUM_CLASSES, PHI, seed=WEIGHT_SEED, flavour="calibrated",
                                variant={"p0": "ffa", "p1": "p1", "p2": "p2"}[VARIANT])


def body_class():
    if VARIANT == "p2":
        from glsdet_b200.yolo_patch_nonlocal_plus import YoloBody
    elif VARIANT == "p1":
        from glsdet_b200.yolox10 import YoloBody
    else:
        from glsdet_b200.yolox_ffa import YoloBody
    return YoloBody


def ref_neck_head(sd, feats):
    from oracle import ref_path

    if VARIANT == "p2":
        return ref_path.p2_neck_head(sd, feats)
    return ref_path.neck_head(sd, feats) if VARIANT == "p0" else ref_path.p1_neck_head(sd, feats)


def make_images(batch, seed):
    """Seeded synthetic images (unit variance, natural-image-like spectrum), [batch, 3, 1024, 1024] fp32 on the host."""
    import torch

    from glsdet_b200.synthetic import synthetic_images

    return torch.cat([synthetic_images(min(4, batch - i), IN_H, IN_W, seed=seed * 64 + i) for i in range(0, batch, 4)])


def make_features(net, batch, seed, device, images=None):
    """Backbone features (NCHW fp32, what the reference's CSPDarknet hands to the neck) of the synthetic images."""
    import torch

    images = make_images(batch, seed) if images is None else images
    feats = None
    with torch.no_grad():
        for i in range(0, batch, 4):
            f = [t.float().contiguous() for t in net.backbone.features(images[i:i + 4].to(device))]
            feats = f if feats is None else [torch.cat([a, b]) for a, b in zip(feats, f)]
    return feats


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_run(sd, feats_cpu, n_images, threads):
    """The reference's own CPU implementation of the path, restated in oracle/ref_path.py (the reference is pure
    PyTorch; its modules cannot be imported on the GPU box because /root/reference is absent there).  One image at
    a time, like yolo.py's per-image loop.  Returns (seconds, candidates, kept)."""
    import torch
    from oracle import ref_path

    torch.set_num_threads(threads)
    cand, kept = [], []
    t0 = time.perf_counter()
    for i in range(n_images):
        f = [t[i:i + 1] for t in feats_cpu]
        logits = ref_neck_head(sd, f)
        pred = ref_path.decode_outputs(logits, [IN_H, IN_W])
        res = ref_path.non_max_suppression(pred, NUM_CLASSES, [IN_H, IN_W], None, False, CONF_THRES, NMS_THRES,
                                           strategy="auto_cpu", correct_boxes=False)
        cand.append(int((pred[0, :, 4] * pred[0, :, 5:].max(1)[0] >= CONF_THRES).sum()))
        kept.append(len(res[0]))
    return time.perf_counter() - t0, cand, kept


def run_reference_arm(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    from oracle import ref_path

    sd = make_weights()
    n = max(1, min(args.cpu_images, 2))
    feats = ref_path.csp_darknet(sd, make_images(n, 1000))   # backbone features on the host (not timed)
    if VARIANT == "p2":
        feats = feats[1:]
    times = []
    cand = kept = None
    for s in range(args.warmup + args.steps):
        dt, cand, kept = cpu_reference_run(sd, feats, n, threads)
        if s >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    line = {"impl": "reference", "metric": "images/sec at 1024^2 (neck+head+NMS)", "value": value, "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"GLSDet YOLOX-s neck+head+decode+NMS, {n} synthetic 1024x1024 images per step on the host CPU",
                       "num_classes": NUM_CLASSES, "conf_thres": CONF_THRES, "nms_thres": NMS_THRES,
                       "candidates_per_image": cand, "kept_per_image": kept},
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                             "sample": f"{n} images per step, batch 1 each, fp32, torch CPU ({threads} threads)"},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ product arm
def run_native_arm(args):
    import torch
    import torch.distributed as dist

    from glsdet_b200 import _native as N
    from glsdet_b200.dist import gather_detections

    YoloBody = body_class()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (product arm) needs a CUDA device: there is no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL may print its version banner on stdout while the communicator is created; the contract is ONE JSON
        # line on stdout, so route fd 1 to stderr for the duration of the (eager) communicator set-up
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    lib = N.load()

    sd = make_weights()
    net = YoloBody(NUM_CLASSES, PHI)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    B = args.batch
    host_images = make_images(B, 1000 + rank).pin_memory()
    feats = make_features(net, B, 1000 + rank, dev, images=host_images)
    host_feats = [t.cpu().pin_memory() for t in feats]
    plan = net.plan_for(feats)
    max_det = args.max_det if args.max_det > 0 else None      # None: every survivor, the reference's semantics
    rows_cap = max_det if max_det is not None else min(args.gather_rows, plan.num_anchors)   # fixed-layout copy-out / gather rows
    nms = net.nms_for(plan, max_det)

    def device_step():
        det, cnt = net.detect_features(feats, conf_thres=CONF_THRES, nms_thres=NMS_THRES, strategy="auto_cuda",
                                       max_det=max_det)
        if world > 1:
            gather_detections(det, cnt, max_rows=rows_cap)
        return det, cnt

    stream = torch.cuda.current_stream()
    clock = ClockSampler(local_rank)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing (value) with the conv segment bracketed for the roofline
    for _ in range(args.warmup):
        device_step()
    sync_all()
    launches0 = lib.glsdet_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    seg = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(args.steps)]
    clock.start()
    ev[0].record(stream)
    for s in range(args.steps):
        plan.load_features(feats)
        seg[s][0].record(stream)
        plan.run_neck()
        plan.run_head("det")
        seg[s][1].record(stream)
        det, cnt = nms.launch(plan.pred, CONF_THRES, NMS_THRES, "auto_cuda", cls_logits=plan.det_cls_logits)
        seg[s][2].record(stream)
        if world > 1:
            gather_detections(det, cnt, max_rows=rows_cap)
    ev[1].record(stream)
    sync_all()
    launches = lib.glsdet_launch_count() - launches0
    elapsed_ms = ev[0].elapsed_time(ev[1])
    conv_ms = sum(a.elapsed_time(b) for a, b, _ in seg) / args.steps
    post_ms = sum(b.elapsed_time(c) for _, b, c in seg) / args.steps   # score filter + sort + NMS + row gather
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * B * args.steps / (elapsed_ms * 1e-3)

    # ---- end-to-end through the public module API with HOST buffers, every step: pinned host input -> device ->
    # path -> detections (counts + rows) back on the host.  Two device input sets: the upload of step i+1 (copy
    # stream) overlaps the compute of step i; the host waits for step i's detections before it moves on.
    #   "images":   YoloBody.detect(image batch) - what a user of the reference calls (yolo.py:116-160 feeds images);
    #               the CSPDarknet backbone (+27.9 GFLOP/image, NOT part of the metric's work and not
    #               part of the reference arm's) runs natively in front of the path, 201 MB of H2D per step;
    #   "features": YoloBody.detect_features(dark2..dark5 fp32 NCHW) - exactly the metric's segment, 503 MB of H2D per
    #               step (PCIe-bound).
    host_cnt = [torch.empty((B,), dtype=torch.int32).pin_memory() for _ in range(2)]
    host_det = [torch.empty((B, rows_cap, 7), dtype=torch.float32).pin_memory() for _ in range(2)]
    d2h_rows = [0]
    copy_stream = torch.cuda.Stream(device=dev)

    def e2e_measure(host_inputs, step_fn):
        dev_in = [[torch.empty_like(t, device=dev) for t in host_inputs] for _ in range(2)]
        uploaded = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]

        def upload(slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])       # the previous user of this input set has finished
                for d, h in zip(dev_in[slot], host_inputs):
                    d.copy_(h, non_blocking=True)
                uploaded[slot].record(copy_stream)

        def run(n_steps):
            for ev_ in consumed:
                ev_.record(stream)
            upload(0)
            for i in range(n_steps):
                slot = i & 1
                if i + 1 < n_steps:
                    upload(slot ^ 1)
                stream.wait_event(uploaded[slot])
                det, cnt = step_fn(dev_in[slot])
                consumed[slot].record(stream)
                if world > 1:
                    gather_detections(det, cnt, max_rows=rows_cap)
                # the reference's copy-out (utils_bbox.py:481 output[i].cpu()): counts first, then the rows that exist
                host_cnt[slot].copy_(cnt, non_blocking=True)
                done[slot].record(stream)
                done[slot].synchronize()
                kmax = int(host_cnt[slot].max())
                if kmax > rows_cap:
                    raise SystemExit(f"an image kept {kmax} boxes > --gather-rows {rows_cap}")
                host_det[slot][:, :kmax].copy_(det[:, :kmax], non_blocking=True)
                done[slot].record(stream)
                done[slot].synchronize()                     # the caller reads this step's detections now
                d2h_rows[0] = kmax

        run(max(2, args.warmup // 2))
        sync_all()
        t0 = time.perf_counter()
        run(args.steps)
        sync_all()
        ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        del dev_in
        return ms

    kw = dict(conf_thres=CONF_THRES, nms_thres=NMS_THRES, strategy="auto_cuda", max_det=max_det)
    if VARIANT in ("p0", "p2"):   # the backbone chains into the plan (no NCHW fp32 round trip)
        image_step = lambda inp: net.detect(inp[0], **kw)
    else:   # P1: native backbone -> NCHW fp32 features (its Gram operands are built from them) -> path
        image_step = lambda inp: net.detect_features(net.backbone.features(inp[0]), **kw)
    feat_ms = e2e_measure(host_feats, lambda inp: net.detect_features(inp, **kw))
    e2e_ms = e2e_measure([host_images], image_step)
    clocks = clock.stop()   # sampled over the device-resident timed region and the two end-to-end regions above
    clocks["sampled_over"] = "device-resident timed region + e2e regions (features, images), 5 ms interval"
    kept = [int(v) for v in host_cnt[(args.steps - 1) & 1]]
    u8_ms = None
    if VARIANT == "p0":   # uint8 HWC frames (what a decoder delivers): normalisation fused into the Focus kernel
        mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
        host_u8 = ((host_images * std + mean) * 255.0).round_().clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory()
        u8_ms = e2e_measure([host_u8], lambda inp: net.detect_uint8(inp[0], **kw))
        kept_u8 = [int(v) for v in host_cnt[(args.steps - 1) & 1]]
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    feat_value = world * B * args.steps / (feat_ms * 1e-3)
    h2d_bytes = host_images.numel() * 4
    h2d_feat_bytes = sum(t.numel() * 4 for t in host_feats)

    # ---- device-resident image -> detections (backbone included), for the record
    dev_images = host_images.to(dev)
    for _ in range(3):
        image_step([dev_images])
    sync_all()
    evi = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    evi[0].record(stream)
    for _ in range(args.steps):
        det, cnt = image_step([dev_images])
        if world > 1:
            gather_detections(det, cnt, max_rows=rows_cap)
    evi[1].record(stream)
    sync_all()
    img_ms = evi[0].elapsed_time(evi[1]) / args.steps
    if world > 1:
        t = torch.tensor([img_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        img_ms = float(t.item())
    d2h_bytes = host_cnt[0].numel() * 4 + B * d2h_rows[0] * 7 * 4   # counts + the [B, max kept, 7] rows of the last step
    plan.run_head(True)   # decoded probabilities (the timed steps leave raw class logits in plan.pred)
    cand = [int(v) for v in ((plan.pred[:, :, 4] * plan.pred[:, :, 5:].max(2)[0]) >= CONF_THRES).sum(1).cpu()]

    if rank == 0:
        pk = peaks()
        algo_gflop = ALGO_GFLOP_PER_IMAGE if VARIANT == "p0" else (plan.flops + getattr(plan, "attn_flops", 0.0)) / B / 1e9
        achieved_tf = algo_gflop * B / conv_ms  # GFLOP / ms = TFLOP/s
        # denominator: the burst figure unless the timed region is long enough (>= 2 s) for the sustained one to apply
        long_region = elapsed_ms >= 2000.0
        tf_peak = pk["tf_sustained"] if long_region else pk["tf_burst"]
        line = {"metric": "images/sec at 1024^2 (neck+head+NMS)", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": ("GLSDet YOLOX-s (yolox_ffa YoloBody(10,'s')) neck+FFA+head+decode+filter+NMS, "
                                        f"batch {B} of synthetic 1024x1024 images per GPU (BASELINE configs[1])") if VARIANT == "p0" else
                                       ("GLSDet P1 YOLOX-s (yolox10 YoloBody(10,'s')) non-local neck+head+decode+filter+NMS, "
                                        f"batch {B} of synthetic 1024x1024 images per GPU (SURVEY 8d row 2')"),
                           "images_per_gpu": B, "num_classes": NUM_CLASSES, "conf_thres": CONF_THRES,
                           "nms_thres": NMS_THRES, "nms_strategy": "torchvision auto dispatch for CUDA tensors",
                           "l2": "inputs (503 MB fp32 feature maps per batch) exceed the 126 MB L2; no explicit flush",
                           "candidates_per_image": cand, "kept_per_image": kept,
                           "max_det": max_det if max_det is not None else "none (every NMS survivor, utils_bbox.py:414-420)",
                           "postprocess_ms": post_ms,
                           "parallelism": f"dp{world} (images sharded, NCCL gather of detections)" if world > 1 else "single GPU"},
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms / args.steps,
                        "how": "wall clock; YoloBody.detect(pinned host image batch fp32) -> host detections; the native "
                               "CSPDarknet backbone (+27.9 GFLOP/image, outside the metric and outside the reference arm) "
                               "runs in front of the path; upload of step i+1 overlaps compute of step i (2 input sets, copy stream)",
                        "from_features": {"value": feat_value, "unit": "images/s", "h2d_bytes_per_step": h2d_feat_bytes,
                                          "ms_per_step": feat_ms / args.steps,
                                          "how": "same, but YoloBody.detect_features(pinned host dark2..dark5 fp32 NCHW): "
                                                 "exactly the metric's segment, PCIe-bound on 503 MB per step"}},
                "e2e_uint8": None if u8_ms is None else {
                    "value": world * B * args.steps / (u8_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": B * IN_H * IN_W * 3,
                    "ms_per_step": u8_ms / args.steps,
                    "how": "YoloBody.detect_uint8(pinned host uint8 HWC frames): preprocess_input (utils.py:47-51) fused into "
                           "the Focus kernel; the same synthetic images clipped and quantised to 8 bits (a different "
                           "NMS load: see kept_per_image)", "kept_per_image": kept_u8},
                "with_backbone": {"value": world * B / (img_ms * 1e-3), "unit": "images/s", "ms_per_step": img_ms,
                                  "what": "device-resident image batch -> backbone -> neck -> head -> NMS (CUDA events)"},
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": tf_peak, "unit": "TFLOP/s",
                             "frac": achieved_tf / tf_peak, "traffic": conv_traffic(B),
                             "kernel": "conv_gemm_kernel (all conv launches of a step, neck+FFA+head segment)",
                             "algorithmic_gflop_per_image": algo_gflop, "segment_ms": conv_ms,
                             "peak_source": pk["source"] + (", sustained figure (timed region >= 2 s)" if long_region else
                                                            f", burst figure (timed region {elapsed_ms * 1e-3:.2f} s < 2 s)")},
                "launches_per_step": int(launches) // args.steps}
        if not args.no_cpu_baseline and world == 1:   # N = 1 only: at N > 1 the other ranks would spin in a barrier meanwhile
            threads = os.cpu_count() or 1
            n = min(args.cpu_images, B)
            feats_cpu = [t[:n].clone() for t in host_feats]
            cpu_reference_run(sd, feats_cpu, 1, threads)  # warm-up
            total_dt, images, ccand, ckept = 0.0, 0, None, None
            while total_dt < args.cpu_seconds and images < 64 * n:
                dt, ccand, ckept = cpu_reference_run(sd, feats_cpu, n, threads)
                total_dt += dt
                images += n
            line["cpu_baseline"] = {"value": images / total_dt, "unit": "images/s", "cores": threads, "kind": "port",
                                    "sample": f"{images} images ({n} of the batch's images, repeated), one at a time "
                                              f"(yolo.py per-image loop), fp32, torch CPU {threads} threads, {total_dt:.1f} s",
                                    "candidates_per_image": ccand, "kept_per_image": ckept}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    global VARIANT
    args = parse_args()
    VARIANT = args.variant
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_native_arm(args)


if __name__ == "__main__":
    sys.exit(main())
