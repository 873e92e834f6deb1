"""Per-kernel totals of one profiled step (multi-metric ncu CSV written by tools/profile_step.py under ncu)."""
import collections
import csv
import sys


def main(path, out=None):
    with open(path) as f:
        rows = [r for r in csv.DictReader(l for l in f if not l.startswith("==")) if r["Metric Name"] == "gpu__time_duration.sum"]
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows:
        n = r["Kernel Name"].split("(")[0].replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "nsecond": v / 1e3, "usecond": v, "msecond": v * 1e3}[r["Metric Unit"]]
        a = agg.setdefault(n, [0.0, 0])
        a[0] += v
        a[1] += 1
        tot += v
    lines = [f"# one step ({path}); ncu per-launch times are cold-cache and serialised: shares, not absolutes",
             f"{'kernel':56s} {'launches':>8s} {'total us':>10s} {'share':>7s}"]
    for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        lines.append(f"{n[:56]:56s} {c:8d} {v:10.1f} {100 * v / tot:6.1f}%")
    lines.append(f"{'TOTAL':56s} {sum(c for _, c in agg.values()):8d} {tot:10.1f}")
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
