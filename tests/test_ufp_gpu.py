"""UFP stage on the GPU: mosaic assembly and map-back + merge NMS through the C ABI, bit-exact
against goldens recorded from the REAL reference functions and against the oracle on seeded cases."""
import hashlib
from pathlib import Path

import numpy as np
import pytest
import torch

from _helpers import ufp_synth_image
from oracle import ufp_ref

pytestmark = pytest.mark.gpu
GOLD = np.load(Path(__file__).parent / "golden" / "ufp_cases.npz")
SEEDS = {"a": 0, "b": 1, "c": 2, "d": 3}


@pytest.mark.parametrize("case", ["a", "b"])
def test_mosaic_bit_exact_vs_reference_golden(case, native_lib, cuda_device):
    from glsdet_b200.ufp import display_merge_result

    w, h = (int(v) for v in GOLD[f"{case}_shape"])
    img = ufp_synth_image(SEEDS[case] + 100, h, w)
    rows = [list(r) for r in GOLD[f"{case}_rows"]]
    new_w, new_h = GOLD[f"{case}_extent"]
    got = display_merge_result(rows, torch.from_numpy(img).to(cuda_device), "synthetic", new_w, new_h).cpu().numpy()
    assert list(got.shape) == list(GOLD[f"{case}_mosaic_shape"])
    assert np.array_equal(got[:96, :128], GOLD[f"{case}_mosaic_crop"])
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).digest() == GOLD[f"{case}_mosaic_sha256"].tobytes()


def test_mosaic_vs_oracle_random_chips(native_lib, cuda_device):
    """Chips of every factor touching the image borders, one-pixel chips, an empty chip list."""
    from glsdet_b200.ufp import display_merge_result

    rng = np.random.default_rng(11)
    img = ufp_synth_image(5, 97, 131)
    rows, y = [], 0
    for cw, ch, sf in ((1, 1, 4), (1, 7, 2), (9, 1, 4), (131, 3, 1), (20, 97, 2), (33, 21, 4), (64, 64, 1), (5, 5, 2)):
        x1 = int(rng.integers(0, 131 - cw + 1))
        y1 = int(rng.integers(0, 97 - ch + 1))
        rows.append([x1 + 0.25, y1 + 0.5, cw + 0.75, ch + 0.5, 3.0, float(y), float(sf)])
        y += ch * sf + 1
    W = max(3 + r[2] * r[6] for r in rows) + 0.5
    got = display_merge_result(rows, torch.from_numpy(img).to(cuda_device), None, W, float(y)).cpu().numpy()
    want = ufp_ref.display_merge_result(rows, img, W, float(y))
    assert np.array_equal(got, want)
    empty = display_merge_result([], torch.from_numpy(img).to(cuda_device), None, 40.0, 30.0)
    assert empty.shape == (30, 40, 3) and int(empty.sum()) == 0
    with pytest.raises(Exception):
        display_merge_result(rows, torch.from_numpy(img), None, W, float(y))       # CPU tensor: no fallback path


@pytest.mark.parametrize("case", ["a", "b"])
def test_merge_bit_exact_vs_reference_golden(case, native_lib, cuda_device):
    from glsdet_b200.ufp import coco_rows, merge_second_stage

    rows = [list(r) for r in GOLD[f"{case}_rows"]]
    nc = int(GOLD[f"{case}_nc"])
    second = [torch.from_numpy(GOLD[f"{case}_second{i}"]).to(cuda_device) for i in range(nc)]
    merged = merge_second_stage(rows, second, 0.6)
    for i in range(nc):
        assert np.array_equal(merged[i].cpu().numpy(), GOLD[f"{case}_merged{i}"]), f"class {i}"
    want = ufp_ref.coco_rows([GOLD[f"{case}_merged{i}"] for i in range(nc)], 7)
    assert coco_rows(merged, 7) == want


def test_merge_vs_oracle_clustered_and_empty(native_lib, cuda_device):
    from glsdet_b200.ufp import merge_second_stage

    rng = np.random.default_rng(21)
    rows = [[10.0, 20.0, 50.0, 40.0, 0.0, 0.0, 4.0], [200.5, 100.25, 80.0, 60.0, 200.0, 0.0, 2.0], [5.0, 300.0, 150.0, 90.0, 0.0, 160.0, 1.0]]
    second = []
    for c in range(4):
        n = (0, 1, 300, 900)[c]
        chip = rng.integers(0, 3, n)
        nx = np.array([0, 200, 0])[chip]
        ny = np.array([0, 0, 160])[chip]
        ww = np.array([200, 160, 150])[chip]
        hh = np.array([160, 120, 90])[chip]
        cx = nx + rng.uniform(0.1, 0.9, n) * ww
        cy = ny + rng.uniform(0.1, 0.9, n) * hh
        bw, bh = rng.uniform(4, 40, n), rng.uniform(4, 40, n)
        sc = (rng.permutation(n) + 1) / (n + 1.0)
        second.append(np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2, sc], 1).astype(np.float32).reshape(-1, 5))
    got = merge_second_stage(rows, [torch.from_numpy(s).to(cuda_device) for s in second], 0.6)
    want = ufp_ref.merge_results(rows, second, 0.6)
    for c in range(4):
        assert np.array_equal(got[c].cpu().numpy(), want[c]), f"class {c}"
    # no chips at all: nothing maps back
    none = merge_second_stage([], [torch.from_numpy(s).to(cuda_device) for s in second], 0.6)
    assert all(t.shape == (0, 5) for t in none)
