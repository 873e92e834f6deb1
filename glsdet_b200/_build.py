"""In-tree build of libglsdet_b200.so (hand-written sm_100a CUDA behind a C ABI).

nvcc cross-compiles without a GPU, so this runs on the CPU build box; the resulting .so is git-ignored
but travels to the GPU box with the working tree.  Only sm_100a is targeted - there is no other backend.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libglsdet_b200.so"
OBJ = PKG / "lib" / "obj"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-ffp-contract=off",   # host float32 arithmetic of ufp.cu follows NumPy (no fused multiply-add)
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libglsdet_b200.so cannot be built (there is no fallback path)")


def _digest(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((PKG.parent / "include").glob("*.h"))):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ and link the shared library. Incremental (content-hash stamps)."""
    nvcc = _nvcc()
    OBJ.mkdir(parents=True, exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    if not sources:
        raise RuntimeError(f"no CUDA sources under {CSRC}")

    def compile_one(src: Path):
        obj = OBJ / (src.stem + ".o")
        stamp = OBJ / (src.stem + ".sha")
        dig = _digest(src)
        if not force and obj.exists() and stamp.exists() and stamp.read_text() == dig:
            return obj, False
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
        if verbose:
            sys.stderr.write(res.stderr)
        stamp.write_text(dig)
        return obj, True

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        results = list(ex.map(compile_one, sources))
    objs = [str(o) for o, _ in results]
    if force or not LIB.exists() or any(changed for _, changed in results):
        cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
