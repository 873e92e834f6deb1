"""YOLO facade (yolox-drone/yolo.py, models/core/utils.py) on the GPU: resize_image bit-exact
against goldens recorded from the REAL reference function (and against live Pillow), and the facade's methods against the
same pipeline assembled by hand from the already-tested pieces."""
import hashlib
from pathlib import Path

import numpy as np
import pytest
import torch

from _helpers import ufp_synth_image

pytestmark = pytest.mark.gpu
GOLD = np.load(Path(__file__).parent / "golden" / "resize_cases.npz")


@pytest.mark.parametrize("name", [str(n) for n in GOLD["names"]])
def test_resize_image_bit_exact_vs_reference_golden(name, native_lib, cuda_device):
    from glsdet_b200.utils import resize_image

    seed, ih, iw, w, h, lb = (int(v) for v in GOLD[f"{name}_meta"])
    img = ufp_synth_image(seed, ih, iw)
    got = resize_image(img, (w, h), bool(lb), device=cuda_device).cpu().numpy()
    assert got.shape == (h, w, 3)
    assert np.array_equal(got[:64, :96], GOLD[f"{name}_crop"])
    assert np.array_equal(got[h // 2 - 16:h // 2 + 16, w // 2 - 24:w // 2 + 24], GOLD[f"{name}_center"])
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).digest() == GOLD[f"{name}_sha256"].tobytes()


def test_resize_image_vs_live_pillow(native_lib, cuda_device):
    Image = pytest.importorskip("PIL.Image")
    from glsdet_b200.utils import resize_image

    rng = np.random.default_rng(5)
    for ih, iw, w, h, lb in ((37, 53, 64, 96, False), (1080, 1920, 1024, 544, True), (50, 50, 50, 96, False), (64, 64, 64, 64, True),
                             (17, 301, 160, 128, True)):
        img = rng.integers(0, 256, (ih, iw, 3), dtype=np.uint8)
        pil = Image.fromarray(img)
        if lb:
            scale = min(w / iw, h / ih)
            nw, nh = int(iw * scale), int(ih * scale)
            want = Image.new("RGB", (w, h), (128, 128, 128))
            want.paste(pil.resize((nw, nh), Image.BICUBIC), ((w - nw) // 2, (h - nh) // 2))
        else:
            want = pil.resize((w, h), Image.BICUBIC)
        got = resize_image(pil, (w, h), lb, device=cuda_device).cpu().numpy()
        assert np.array_equal(got, np.array(want)), (ih, iw, w, h, lb)


def _facade(cuda_device, letterbox, decode_mode="default", shape=(256, 320)):
    from glsdet_b200.synthetic import synthetic_state_dict
    from glsdet_b200.yolo import YOLO

    names = [f"c{i}" for i in range(10)]
    sd = synthetic_state_dict(10, "s", seed=3, flavour="calibrated")
    return YOLO(state_dict=sd, class_names=names, input_shape=list(shape), phi="s", config_path="models/ffa/yolox_ffa.py",
                letterbox_image=letterbox, confidence=0.02, nms_iou=0.65, decode_mode=decode_mode), names


@pytest.mark.parametrize("letterbox", [False, True])
def test_facade_detect_equals_hand_assembled_pipeline(letterbox, native_lib, cuda_device, tmp_path):
    Image = pytest.importorskip("PIL.Image")
    from glsdet_b200.utils import resize_image
    from glsdet_b200.utils_bbox import detection_lines, yolo_correct_boxes

    yolo, names = _facade(cuda_device, letterbox)
    img = Image.fromarray(ufp_synth_image(12, 300, 420))
    rows = yolo.detect(img)
    assert rows is not None and rows.ndim == 2 and rows.shape[1] == 7 and len(rows) > 0
    # the same by hand: resize (tested above) -> fused detect on the uint8 frame -> yolo_correct_boxes
    u8 = resize_image(img, (320, 256), letterbox, device=cuda_device).unsqueeze(0)
    det, cnt = yolo.net.module.detect_uint8(u8, conf_thres=0.02, nms_thres=0.65)
    n = int(cnt[0])
    raw = det[0, :n].cpu().numpy()
    want = raw.copy()
    want[:, :4] = yolo_correct_boxes((raw[:, 0:2] + raw[:, 2:4]) / 2, raw[:, 2:4] - raw[:, 0:2], [256, 320], np.array([300, 420]), letterbox)
    assert np.array_equal(rows, want)
    assert (np.diff(rows[:, 4] * rows[:, 5]) <= 0).all()          # score order
    # get_map_txt: the wire format of yolo.py:302-303, restricted to the requested classes
    yolo.get_map_txt("img0", img, names[:2], str(tmp_path))
    text = (tmp_path / "detection-results" / "img0.txt").read_text().splitlines()
    assert text == detection_lines(rows, names, names[:2])
    assert all(len(l.split()) == 6 for l in text)
    # detect_image draws on a copy-compatible PIL image of the same size; get_FPS returns seconds per image
    out = yolo.detect_image(img.copy())
    assert out.size == img.size
    assert 0 < yolo.get_FPS(img, 2) < 5
    # batch entry point = per-image results
    both = yolo.detect_batch([img, img])
    assert np.array_equal(both[0], rows) and np.array_equal(both[1], rows)


def test_facade_decode_modes_follow_reference_call_sequence(native_lib, cuda_device):
    """yolo.py:75-82: decode_mode picks the decode function; the facade then runs net(images) -> decode_func ->
    non_max_suppression like the reference."""
    Image = pytest.importorskip("PIL.Image")
    from glsdet_b200 import utils_bbox as ub
    from glsdet_b200.utils import preprocess_input, resize_image

    img = Image.fromarray(ufp_synth_image(13, 200, 260))
    for mode, fn in (("obj_sigmoid", ub.decode_outputs_no_sigmoid), ("cls_sigmoid", ub.decode_outputs_cls_sigmoid)):
        yolo, _ = _facade(cuda_device, False, decode_mode=mode, shape=(128, 160))
        assert yolo.decode_func is fn
        rows = yolo.detect(img)
        u8 = resize_image(img, (160, 128), False, device=cuda_device)
        x = torch.from_numpy(np.expand_dims(np.transpose(preprocess_input(u8.cpu().numpy().astype("float32")), (2, 0, 1)), 0)).to(cuda_device)
        want = ub.non_max_suppression(fn(yolo.net(x), [128, 160]), 10, [128, 160], np.array([200, 260]), False, conf_thres=0.02, nms_thres=0.65)
        assert np.array_equal(rows, want[0])
    with pytest.raises(ValueError):
        _facade(cuda_device, False, decode_mode="nonsense")
