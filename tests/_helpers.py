"""Helpers shared by the GPU parity tests."""
import numpy as np
import torch

TOL = 2e-2


def assert_close_rel(got: torch.Tensor, ref: torch.Tensor, tol=TOL, what="", max_factor=4.0, frac=1e-2):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite values"
    d = (got - ref).abs()
    scale = ref.abs().max().item()
    rel_l2 = (got - ref).norm().item() / max(ref.norm().item(), 1e-30)
    assert rel_l2 <= tol, f"{what}: relative l2 error {rel_l2:.4g} > {tol}"
    frac_bad = (d > tol * scale).float().mean().item()
    assert frac_bad <= frac, f"{what}: {frac_bad:.2%} of the elements differ by more than {tol} * max|ref|"
    assert d.max().item() <= max_factor * tol * scale, f"{what}: max abs diff {d.max().item():.4g} vs scale {scale:.4g}"



def _clustered(rng, k, nc, spread=0.004, lo=0.05):
    g = max(4, k // 12)
    cen = rng.uniform(lo, 0.95, (g, 2))
    which = rng.integers(0, g, k)
    c = cen[which] + rng.normal(0, spread, (k, 2))
    wh = np.exp(rng.normal(-3.2, 0.4, (g, 2)))[which] * np.exp(rng.normal(0, 0.15, (k, 2)))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)
    scores = ((rng.permutation(k) + 1) / (k + 1.0)).astype(np.float32)
    labels = rng.integers(0, nc, k).astype(np.float32)
    return boxes, scores, labels




def ufp_synth_image(seed: int, h: int, w: int) -> np.ndarray:
    """Deterministic BGR uint8 test image of the UFP goldens (integer arithmetic only, so tests/golden/make_golden_ufp.py
    and the tests rebuild the same bytes on any machine): smooth structure plus seeded noise."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.int64)
    rng = np.random.default_rng(seed)
    chans = [((xx * (3 + c) + yy * (5 - c) + ((xx * xx + 3 * yy * yy) >> (5 + c)) + 40 * c) & 255) for c in range(3)]
    img = np.stack(chans, -1) + rng.integers(-24, 25, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)
