"""UFP stage without a GPU: the oracle restatement and the host-side packing of the native
library against goldens recorded from the REAL reference functions (tests/golden/make_golden_ufp.py)."""
import hashlib
from pathlib import Path

import numpy as np
import pytest

from _helpers import ufp_synth_image
from oracle import ufp_ref

GOLD = np.load(Path(__file__).parent / "golden" / "ufp_cases.npz")
CASES = [str(c) for c in GOLD["cases"]]
SEEDS = {"a": 0, "b": 1, "c": 2, "d": 3}


@pytest.mark.parametrize("case", CASES)
def test_oracle_packing_matches_reference_golden(case):
    w, h = (int(v) for v in GOLD[f"{case}_shape"])
    rows, new_w, new_h = ufp_ref.unified_foreground_packing(GOLD[f"{case}_boxes"].copy(), 1.5, [w, h])
    assert np.array_equal(np.array(rows, dtype=np.float64).reshape(-1, 7), GOLD[f"{case}_rows"])
    assert [float(new_w), float(new_h)] == list(GOLD[f"{case}_extent"])


@pytest.mark.parametrize("case", CASES)
def test_native_host_packing_bit_exact(case, native_lib):
    """glsdet_ufp_pack is a HOST function of the C ABI (no device work): bit-exact rows, extent and order."""
    from glsdet_b200.ufp import UnifiedForegroundPacking

    w, h = (int(v) for v in GOLD[f"{case}_shape"])
    rows, new_w, new_h = UnifiedForegroundPacking(GOLD[f"{case}_boxes"].copy(), 1.5, input_shape=[w, h])
    assert np.array_equal(np.array(rows, dtype=np.float64).reshape(-1, 7), GOLD[f"{case}_rows"])
    assert [new_w, new_h] == list(GOLD[f"{case}_extent"])


def test_native_host_packing_edge_cases(native_lib):
    from glsdet_b200.ufp import UnifiedForegroundPacking

    rows, new_w, new_h = UnifiedForegroundPacking(np.zeros((0, 4), np.float32), 1.5, input_shape=[640, 480])
    assert rows == [] and new_w == 0 and new_h == 0
    ref = ufp_ref.unified_foreground_packing(np.zeros((0, 4), np.float32), 1.5, [640, 480])
    assert ref[0] == [] and ref[1] == 0 and ref[2] == 0
    # identical boxes (equality branches of the packer), a box larger than the strip, degenerate boxes
    rng = np.random.default_rng(7)
    b = np.array([[10, 10, 40, 40]] * 5 + [[100, 50, 130, 80]] * 4 + [[0, 0, 1300, 700], [5, 5, 5, 5], [300, 300, 310, 300]], np.float32)
    b = np.concatenate([b, rng.uniform(0, 600, (40, 4)).astype(np.float32)])
    b[:, 2:] = np.maximum(b[:, 2:], b[:, :2])
    got = UnifiedForegroundPacking(b.copy(), 1.5, input_shape=[1360, 765])
    want = ufp_ref.unified_foreground_packing(b.copy(), 1.5, [1360, 765])
    assert np.array_equal(np.array(got[0]).reshape(-1, 7), np.array(want[0], dtype=np.float64).reshape(-1, 7))
    assert (got[1], got[2]) == (float(want[1]), float(want[2]))


@pytest.mark.parametrize("case", ["a", "b"])
def test_oracle_mosaic_and_merge_match_reference_golden(case):
    w, h = (int(v) for v in GOLD[f"{case}_shape"])
    img = ufp_synth_image(SEEDS[case] + 100, h, w)
    rows = [list(r) for r in GOLD[f"{case}_rows"]]
    mosaic = ufp_ref.display_merge_result(rows, img, *GOLD[f"{case}_extent"])
    assert list(mosaic.shape) == list(GOLD[f"{case}_mosaic_shape"])
    assert np.array_equal(mosaic[:96, :128], GOLD[f"{case}_mosaic_crop"])
    assert hashlib.sha256(np.ascontiguousarray(mosaic).tobytes()).digest() == GOLD[f"{case}_mosaic_sha256"].tobytes()
    nc = int(GOLD[f"{case}_nc"])
    second = [GOLD[f"{case}_second{i}"] for i in range(nc)]
    mapped = ufp_ref.map_back(rows, second)
    merged = ufp_ref.merge_results(rows, second, 0.6)
    for i in range(nc):
        assert np.array_equal(mapped[i], GOLD[f"{case}_mapped{i}"])
        assert np.array_equal(merged[i], GOLD[f"{case}_merged{i}"])


def test_resize_restatement_matches_cv2_when_present():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for _ in range(40):
        h, w = int(rng.integers(1, 50)), int(rng.integers(1, 50))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for sf in (1, 2, 4):
            assert np.array_equal(ufp_ref.resize_linear_u8(img, sf), cv2.resize(img, (w * sf, h * sf)))
