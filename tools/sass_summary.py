"""Per-kernel counts of the Blackwell-specific SASS instructions in libglsdet_b200.so (cuobjdump -sass): tcgen05 MMAs
(UTCHMMA / .2CTA), TMEM loads (LDTM), TMA loads / stores (UTMALDG / UTMASTG), tensor-core barriers (UTCBAR), and - as a
check that nothing fell back to the legacy path - HMMA / wgmma.  Usage: python tools/sass_summary.py [out.txt]"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "glsdet_b200" / "lib" / "libglsdet_b200.so"
PATTERNS = ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "HMMA", "MUFU.TANH", "SYNCS")


def main(out=None):
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?PT?\d*\s+)?([A-Z][A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["instructions"] += 1
        for p in PATTERNS:
            if p == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    cur[p] += 1
            elif op.startswith(p):
                cur[p] += 1
    demangled = {}
    try:
        names = subprocess.run(["cu++filt"] + list(per), capture_output=True, text=True).stdout.splitlines()
        demangled = dict(zip(per, names))
    except Exception:
        pass
    lines = [f"# {LIB.relative_to(ROOT)}: SASS summary (sm_100a), one row per kernel; counts are static instruction counts",
             f"{'kernel':70s} {'instr':>7s} " + " ".join(f"{p:>13s}" for p in PATTERNS)]
    tot = collections.Counter()
    for k, c in per.items():
        name = demangled.get(k, k)
        name = name.replace("glsdet::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
        cut = name.find(">(")
        name = name[:cut + 1] if cut >= 0 else name.split("(")[0]
        name = name.replace("(bool)", "").replace("(int)", "")
        lines.append(f"{name[:70]:70s} {c['instructions']:7d} " + " ".join(f"{c[p]:13d}" for p in PATTERNS))
        tot.update(c)
    lines.append(f"{'TOTAL (' + str(len(per)) + ' kernels)':70s} {tot['instructions']:7d} " + " ".join(f"{tot[p]:13d}" for p in PATTERNS))
    s = "\n".join(lines)
    print(s)
    if out:
        Path(out).write_text(s + "\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
