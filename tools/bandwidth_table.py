"""Achieved DRAM bandwidth of the non-GEMM kernels from an ncu per-launch metrics CSV (duration, dram bytes):
python tools/bandwidth_table.py <csv> [out.txt] [peak GB/s]."""
import collections
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main(path, out=None, peak=None):
    if peak is None:
        p = ROOT / "MEASURED_PEAKS.json"
        peak = json.loads(p.read_text())["hbm_gbs"] if p.exists() else 6650.0
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        if "conv_gemm" in r["Kernel Name"]:
            continue
        d = per.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0].replace("void ", "").replace("glsdet::", "")})
        v = float(r["Metric Value"].replace(",", ""))
        unit, name = r["Metric Unit"], r["Metric Name"]
        if name == "gpu__time_duration.sum":
            v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "usecond": v, "nsecond": v / 1e3, "msecond": v * 1e3}.get(unit, v)
        if name.startswith("dram__bytes"):
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1) / 1e6
        d[name] = v
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0)
        a[3] += d.get("dram__bytes_write.sum", 0.0)
    text = [f"# non-GEMM kernels of {path}: DRAM traffic / duration against the measured HBM peak ({peak:.0f} GB/s); ncu launches are",
            "# cold-cache and serialised.  Kernels that move little data (sort steps, plans, FC) are latency-bound by design.",
            f"{'kernel':44s} {'launches':>8s} {'us':>9s} {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>8s} {'of peak':>8s}"]
    for n, (c, us, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = (rd + wr) / us * 1e3 if us else 0.0   # MB / us = TB/s -> GB/s
        text.append(f"{n[:44]:44s} {c:8d} {us:9.1f} {rd:9.1f} {wr:9.1f} {gbs:8.0f} {gbs / peak:8.2f}")
    s = "\n".join(text)
    print(s)
    if out:
        open(out, "w").write(s + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, float(sys.argv[3]) if len(sys.argv) > 3 else None)
