"""Golden vectors of the UFP stage recorded from the REAL reference code.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_ufp.py      (build container only: needs /root/reference, cv2)

  * yolox-ufp/mmdet/core/ufp/{unified_foreground_packing,spp}.py are imported as they are (through a stand-in package,
    because `import mmdet` needs mmcv): UnifiedForegroundPacking -> rows, new_w, new_h.
  * ufpmp_det_eval.py imports mmcv / pycocotools at module level, so compute_iof, py_cpu_nms and display_merge_result are
    compiled from the reference file's own source through `ast` (the function bodies run unmodified; cv2.imread is handed
    the synthetic image).  The map-back loop (:270-296) and the merge (:297-306) live inside main(): they are replayed
    here statement by statement on top of the real compute_iof / py_cpu_nms.
Inputs: clustered VisDrone-like coarse detections (float32), a synthetic BGR uint8 image (tests/_helpers.py::ufp_synth_image), synthetic second-stage
detections placed inside / across the chips of the mosaic (distinct scores)."""
import ast
import hashlib
import importlib.util
import math
import sys
import types
from pathlib import Path

import cv2
import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from _helpers import ufp_synth_image  # noqa: E402  (the tests rebuild the image from the same function)

REF = Path("/root/reference/yolox-ufp")
sys.dont_write_bytecode = True


def load_ufp():
    pkg = types.ModuleType("refufp")
    pkg.__path__ = [str(REF / "mmdet/core/ufp")]
    sys.modules["refufp"] = pkg
    for name in ("spp", "unified_foreground_packing"):
        spec = importlib.util.spec_from_file_location(f"refufp.{name}", REF / f"mmdet/core/ufp/{name}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"refufp.{name}"] = mod
        spec.loader.exec_module(mod)
    return sys.modules["refufp.unified_foreground_packing"]


def load_eval_functions(image):
    src = (REF / "ufpmp_det_eval.py").read_text()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("compute_iof", "py_cpu_nms", "display_merge_result")]
    mod = ast.Module(body=keep, type_ignores=[])
    fake_cv2 = types.SimpleNamespace(imread=lambda path: image, resize=cv2.resize)
    ns = {"np": np, "math": math, "cv2": fake_cv2}
    exec(compile(mod, str(REF / "ufpmp_det_eval.py"), "exec"), ns)
    return ns


def coarse_boxes(rng, n, w, h, clusters):
    cx = rng.uniform(0.1 * w, 0.9 * w, clusters)
    cy = rng.uniform(0.1 * h, 0.9 * h, clusters)
    k = rng.integers(0, clusters, n)
    x = cx[k] + rng.normal(0, 0.04 * w, n)
    y = cy[k] + rng.normal(0, 0.04 * h, n)
    bw = np.exp(rng.normal(2.8, 0.6, n))
    bh = np.exp(rng.normal(2.8, 0.6, n))
    b = np.stack([x - bw / 2, y - bh / 2, x + bw / 2, y + bh / 2], 1)
    b[:, 0::2] = np.clip(b[:, 0::2], 0, w - 1)
    b[:, 1::2] = np.clip(b[:, 1::2], 0, h - 1)
    return b.astype(np.float32)



def second_stage(rng, rec, nc, per_class):
    """Detections in mosaic pixels: most inside a chip, some straddling chip borders, some covering a whole chip."""
    chips = [[math.floor(v) for v in r] for r in rec]
    out = []
    score_pool = rng.permutation(nc * per_class).astype(np.float32)
    for c in range(nc):
        rows = []
        for j in range(per_class):
            _, _, cw, ch, nx, ny, sf = chips[rng.integers(0, len(chips))]
            W, H = cw * sf, ch * sf
            mode = rng.integers(0, 10)
            if mode < 7:      # inside, clustered so that the NMS has work
                bw, bh = rng.uniform(4, max(5, W / 2)), rng.uniform(4, max(5, H / 2))
                x = nx + rng.uniform(0, max(1, W - bw)) + rng.normal(0, 1.0)
                y = ny + rng.uniform(0, max(1, H - bh)) + rng.normal(0, 1.0)
            elif mode < 9:    # straddling the right / bottom border
                bw, bh = rng.uniform(6, 30), rng.uniform(6, 30)
                x, y = nx + W - bw * rng.uniform(0.5, 1.0), ny + H - bh * rng.uniform(0.5, 1.0)
            else:             # covering the whole chip (the chip is the smaller rectangle)
                x, y, bw, bh = nx - 2.5, ny - 1.5, W + 5.0, H + 4.0
            rows.append([x, y, x + bw, y + bh, 0.05 + 0.9 * score_pool[c * per_class + j] / (nc * per_class)])
            if mode < 3 and j + 1 < per_class:   # a near-duplicate for the NMS
                rows.append([x + 1.0, y + 0.5, x + bw + 0.5, y + bh + 1.0, 0.05 + 0.9 * (score_pool[c * per_class + j] + 0.5) / (nc * per_class)])
        out.append(np.array(rows[:per_class], dtype=np.float32))
    return out


def main():
    ufp = load_ufp()
    out = {}
    cases = (("a", 0, 320, 240, 60, 6), ("b", 1, 640, 480, 160, 12), ("c", 2, 1360, 765, 420, 25), ("d", 3, 200, 150, 1, 1))
    names = []
    for name, seed, w, h, n, clusters in cases:
        rng = np.random.default_rng(seed)
        boxes = coarse_boxes(rng, n, w, h, clusters)
        rec, new_w, new_h = ufp.UnifiedForegroundPacking(boxes.copy(), 1.5, input_shape=[w, h])
        out[f"{name}_boxes"] = boxes
        out[f"{name}_shape"] = np.array([w, h])
        out[f"{name}_rows"] = np.array([[float(v) for v in r] for r in rec], dtype=np.float64).reshape(-1, 7)
        out[f"{name}_extent"] = np.array([float(new_w), float(new_h)])
        names.append(name)
        if name in ("a", "b"):
            image = ufp_synth_image(seed + 100, h, w)
            fn = load_eval_functions(image)
            mosaic = fn["display_merge_result"](rec, "synthetic", "synthetic", new_w, new_h)
            assert mosaic.dtype == np.float64 and np.array_equal(mosaic, np.rint(mosaic)) and mosaic.max() <= 255
            m8 = mosaic.astype(np.uint8)
            # the mosaic is megabytes: its SHA-256 pins every byte, a corner crop gives a readable diff
            out[f"{name}_mosaic_shape"] = np.array(m8.shape)
            out[f"{name}_mosaic_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(m8).tobytes()).digest(), dtype=np.uint8)
            out[f"{name}_mosaic_crop"] = m8[:96, :128].copy()
            nc, per_class = (3, 40) if name == "a" else (10, 120)
            second = second_stage(rng, rec, nc, per_class)
            # ufpmp_det_eval.py:270-306 replayed on the real compute_iof / py_cpu_nms
            new_second = [[] for _ in range(nc)]
            for chips in rec:
                o_x1, o_y1, cw, ch, n_x, n_y, scale_factor = [math.floor(_) for _ in chips]
                chip_bbox = [n_x, n_y, n_x + cw * scale_factor, n_y + ch * scale_factor]
                for idx, _results in enumerate(second):
                    for _result in _results:
                        x1, y1, x2, y2, score = _result
                        t_bbox = [x1, y1, x2, y2]
                        if fn["compute_iof"](t_bbox, chip_bbox) > 0.9:
                            nw_ = (x2 - x1) / scale_factor
                            nh_ = (y2 - y1) / scale_factor
                            nx_ = (x1 - n_x) / scale_factor + o_x1
                            ny_ = (y1 - n_y) / scale_factor + o_y1
                            new_second[idx].append([nx_, ny_, nx_ + nw_, ny_ + nh_, score])
            for idx in range(nc):
                mapped = np.array(new_second[idx])
                out[f"{name}_second{idx}"] = second[idx]
                if mapped.shape[0] == 0:
                    out[f"{name}_mapped{idx}"] = np.zeros((0, 5), np.float32)
                    out[f"{name}_merged{idx}"] = np.zeros((0, 5), np.float32)
                    continue
                assert mapped.dtype == np.float32, mapped.dtype
                keep = fn["py_cpu_nms"](mapped, 0.6)
                out[f"{name}_mapped{idx}"] = mapped
                out[f"{name}_merged{idx}"] = mapped[keep]
            out[f"{name}_nc"] = np.array(nc)
    out["cases"] = np.array(names)
    np.savez_compressed(HERE / "ufp_cases.npz", **out)
    print({k: (v.shape, v.dtype) for k, v in out.items() if "second" not in k and "mapped" not in k and "merged" not in k})
    for name in ("a", "b"):
        print(name, "mapped", [len(out[f"{name}_mapped{i}"]) for i in range(int(out[f"{name}_nc"]))],
              "merged", [len(out[f"{name}_merged{i}"]) for i in range(int(out[f"{name}_nc"]))])


if __name__ == "__main__":
    main()
