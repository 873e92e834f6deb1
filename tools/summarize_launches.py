"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one bench step, per kernel."""
import collections
import csv
import sys


def main(path, step_index=1, out=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    # 4 consecutive converts (dark2..dark5 -> NHWC bf16) open every device-resident step
    conv = [i for i, n in enumerate(names) if "nchw_to_nhwc" in n]
    starts = [i for k, i in enumerate(conv) if conv[k:k + 4] == [i, i + 1, i + 2, i + 3] and (k == 0 or conv[k - 1] != i - 1)]
    start = starts[step_index]
    end = starts[step_index + 1] if step_index + 1 < len(starts) else len(rows)
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows[start:end]:
        n = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}[r["Metric Unit"]]
        a = agg.setdefault(n, [0.0, 0])
        a[0] += v
        a[1] += 1
        tot += v
    lines = [f"# one step (launches {start}..{end - 1} of {path}); ncu per-launch times are cold-cache and serialised",
             f"{'kernel':56s} {'launches':>8s} {'total us':>10s} {'share':>7s}"]
    for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        lines.append(f"{n[:56]:56s} {c:8d} {v:10.1f} {100 * v / tot:6.1f}%")
    lines.append(f"{'TOTAL':56s} {sum(c for _, c in agg.values()):8d} {tot:10.1f}")
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1, sys.argv[3] if len(sys.argv) > 3 else None)
