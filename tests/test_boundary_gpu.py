"""The drop-in boundary exactly as the reference uses it: the YOLO facade's loading sequence
(yolox-drone/yolo.py:99-111: module-path import -> YoloBody(num_classes, phi) -> strict load_state_dict(torch.load(path)) ->
.eval() -> nn.DataParallel(net).cuda()), its decode_mode dispatch (yolo.py:75-82) and the post-processing call (:139-143)."""
import importlib
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.nn as nn

from _helpers import TOL, assert_close_rel
from oracle import ref_path

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"

DECODE_FUNCS = {"default": "decode_outputs", "obj_sigmoid": "decode_outputs_no_sigmoid",
                "no_sigmoid": "decode_outputs_no_sigmoid_all", "cls_sigmoid": "decode_outputs_cls_sigmoid"}   # yolo.py:75-82


def test_decode_variants_match_reference_golden(native_lib, cuda_device):
    """utils_bbox.py:36-251 against the outputs of the real functions (tests/golden/make_golden_decode.py)."""
    from glsdet_b200 import utils_bbox

    z = np.load(GOLD / "decode_variants.npz")
    levels = [torch.from_numpy(z[f"level{i}"]).to(cuda_device) for i in range(3)]
    shape = [int(v) for v in z["input_shape"]]
    for name, fn in (("default", "decode_outputs"), ("no_sigmoid", "decode_outputs_no_sigmoid"),
                     ("no_sigmoid_all", "decode_outputs_no_sigmoid_all"), ("cls_sigmoid", "decode_outputs_cls_sigmoid"),
                     ("xyxy", "decode_outputs_xyxy")):
        got = getattr(utils_bbox, fn)([l.clone() for l in levels], shape)
        assert got.shape == z[name].shape
        # expf / division on the device against torch on the CPU: a few ulp
        np.testing.assert_allclose(got.cpu().numpy(), z[name], rtol=2e-6, atol=1e-6, err_msg=name)
    for l, orig in zip(levels, (z[f"level{i}"] for i in range(3))):
        assert np.array_equal(l.cpu().numpy(), orig), "decode must not modify the head outputs"


@pytest.mark.parametrize("module_path", ["glsdet_b200/yolox_ffa.py", "glsdet_b200/yolox_base.py"])
def test_yolo_generate_sequence_and_decode_modes(module_path, native_lib, cuda_device, tmp_path):
    from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict

    num_classes, phi, input_shape = 10, "s", [128, 160]
    stock = "yolox_base" in module_path
    sd = synthetic_state_dict(num_classes, phi, seed=21 if stock else 3, flavour="calibrated", variant="stock" if stock else "ffa")
    model_path = tmp_path / "weights.pth"
    torch.save(sd, model_path)

    # ---- yolo.py:99-111, literally
    config_path = module_path[:-3].replace('/', '.')
    x = importlib.import_module(config_path)
    net = x.YoloBody(num_classes, phi)
    device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
    net.load_state_dict(torch.load(model_path, map_location=device))
    net = net.eval()
    net = nn.DataParallel(net)
    net = net.cuda()

    # ---- yolo.py:136-143 for every decode_mode of :75-82 (batch 2 so that two visible devices get one image each)
    utils_bbox = importlib.import_module("glsdet_b200.utils_bbox")
    images_cpu = synthetic_images(2, input_shape[0], input_shape[1], seed=9)
    image_shape = np.array([256, 480])
    feats = ref_path.csp_darknet(sd, images_cpu)
    ref_logits = ref_path.stock_neck_head(sd, feats[1:]) if stock else ref_path.neck_head(sd, feats)
    with torch.no_grad():
        images = images_cpu.cuda()
        outputs = net(images)
    assert isinstance(outputs, (list, tuple)) and len(outputs) == len(ref_logits)
    for o, r in zip(outputs, ref_logits):
        assert o.dtype == torch.float32 and o.is_cuda
        assert_close_rel(o, r, TOL, f"{module_path} raw logits through nn.DataParallel", frac=2e-2)
    for mode, fn_name in DECODE_FUNCS.items():
        decode_func = getattr(utils_bbox, fn_name)
        pred = decode_func(outputs, input_shape)
        want = (ref_path.decode_outputs([o.cpu() for o in outputs], input_shape) if mode == "default" else
                ref_path.decode_outputs_variant([o.cpu() for o in outputs], input_shape,
                                                {"obj_sigmoid": "no_sigmoid", "no_sigmoid": "no_sigmoid_all", "cls_sigmoid": "cls_sigmoid"}[mode]))
        np.testing.assert_allclose(pred.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-6, err_msg=mode)
        results = utils_bbox.non_max_suppression(pred, num_classes, input_shape, image_shape, False, conf_thres=0.01,
                                                 nms_thres=0.65)
        ref_res = ref_path.non_max_suppression(pred.cpu().clone(), num_classes, input_shape, image_shape, False, 0.01, 0.65,
                                               "auto_cuda")
        assert len(results) == 2
        for got, ref in zip(results, ref_res):
            np.testing.assert_array_equal(got, ref)
    # a second forward re-creates the DataParallel replicas: the plans must be reused, and the result identical
    from glsdet_b200 import _native

    n0 = _native.load().glsdet_launch_count()
    with torch.no_grad():
        outputs2 = net(images)
    assert all(torch.equal(a, b) for a, b in zip(outputs, outputs2))
    assert _native.load().glsdet_launch_count() > n0


def test_device_nms_results_do_not_alias_across_launches(native_lib, cuda_device):
    """A caller may keep step i's detections across step i+1 (launch() used to return the same buffers)."""
    from glsdet_b200.utils_bbox import DeviceNMS

    B, A, nc = 2, 512, 3
    g = torch.Generator().manual_seed(1)
    pred = torch.rand((B, A, 5 + nc), generator=g)
    pred[..., 2:4] *= 0.05
    op = DeviceNMS(B, A, nc)
    det1, cnt1 = op.launch(pred.to(cuda_device), 0.2, 0.5)
    torch.cuda.synchronize()
    keep1, c1 = det1.clone(), cnt1.clone()
    pred2 = pred.clone()
    pred2[..., 4] *= 0.5
    det2, cnt2 = op.launch(pred2.to(cuda_device), 0.2, 0.5)
    torch.cuda.synchronize()
    assert det1.data_ptr() != det2.data_ptr() and cnt1.data_ptr() != cnt2.data_ptr()
    assert torch.equal(cnt1, c1) and not torch.equal(cnt1, cnt2)
    for b in range(B):
        assert torch.equal(det1[b, :int(c1[b])], keep1[b, :int(c1[b])])
    out = (torch.empty_like(det1), torch.empty_like(cnt1), torch.empty((B, A), dtype=torch.int32, device=cuda_device))
    det3, cnt3 = op.launch(pred.to(cuda_device), 0.2, 0.5, out=out)
    torch.cuda.synchronize()
    assert det3.data_ptr() == out[0].data_ptr() and torch.equal(cnt3, c1)


def test_cuda_graph_replay_equals_eager(native_lib, cuda_device):
    """neck -> head -> filter -> NMS captured once per plan as a CUDA graph (with the
    programmatic-dependent-launch edges) must return exactly the eager results, also for new inputs and after re-capture."""
    from glsdet_b200.synthetic import synthetic_images, synthetic_state_dict
    from glsdet_b200.yolox_ffa import YoloBody

    sd = synthetic_state_dict(10, "s", seed=3, flavour="calibrated")
    net = YoloBody(10, "s")
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda_device).eval()
    lib_count = native_lib.glsdet_launch_count
    for seed in (1, 2):
        x = synthetic_images(2, 256, 320, seed=seed).to(cuda_device)
        feats = [f.float().contiguous() for f in net.backbone.features(x)]
        det, cnt = net.detect_features(feats, conf_thres=0.02, nms_thres=0.65)
        det, cnt = det.clone(), cnt.clone()
        n0 = lib_count()
        gdet, gcnt = net.detect_features(feats, conf_thres=0.02, nms_thres=0.65, graph=True)
        torch.cuda.synchronize()
        launched = lib_count() - n0
        assert torch.equal(gcnt, cnt) and int(cnt.sum()) > 0
        for b in range(2):
            assert torch.equal(gdet[b, :int(cnt[b])], det[b, :int(cnt[b])])
        if seed == 2:   # the second call replays: only the four layout converters are launched from the host
            assert launched == 4, launched


def test_gather_payload_kernel_equals_framework_packing(native_lib, cuda_device):
    """glsdet_pack_detections (one launch) against the torch-op packing the CPU / gloo path uses, padded and packed
    layouts, including an image without detections, counts above the row cap and a payload that overflows total_rows."""
    from glsdet_b200.dist import DetectionGather

    g = torch.Generator().manual_seed(3)
    det = torch.rand(5, 40, 7, generator=g)
    cnt = torch.tensor([7, 0, 40, 55, 13], dtype=torch.int32)
    for total in (None, 64, 30):
        cpu = DetectionGather(5, 16, "cpu", total_rows=total)
        gpu = DetectionGather(5, 16, cuda_device, total_rows=total)
        a = cpu.result(cpu.submit(det, cnt))
        b = gpu.result(gpu.submit(det.to(cuda_device), cnt.to(cuda_device)))
        assert torch.equal(a[1], b[1].cpu())
        if total is None:
            for i in range(5):
                n = int(a[1][i])
                assert torch.equal(a[0][i, :n], b[0][i, :n].cpu())
        else:
            ra, rb = cpu.unpack(*a), gpu.unpack(*b)
            assert [len(r) for r in ra] == [len(r) for r in rb]
            assert all(torch.equal(x, y.cpu()) for x, y in zip(ra, rb))
