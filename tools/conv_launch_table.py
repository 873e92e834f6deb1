"""Join an ncu per-launch metrics CSV of conv_gemm_kernel with the plan's op list (layer names, FLOPs)."""
import collections
import csv
import sys

sys.path.insert(0, ".")


def plan_ops():
    """Op order of a step (neck, head, decoded preds) with analytic FLOPs - needs no GPU: mirrors engine.py."""
    return None


def main(path, out=None, ops_path=None):
    import json
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if "conv_gemm" in r["Kernel Name"]]
    ops = json.load(open(ops_path)) if ops_path else None
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r["ID"], {})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        name = r["Metric Name"]
        if name == "gpu__time_duration.sum":
            v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "usecond": v, "nsecond": v / 1e3, "msecond": v * 1e3}.get(unit, v)
        if name.startswith("dram__bytes") or name.startswith("lts__t_bytes"):
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1) / 1e6
        d[name] = v
    text = [f"{'#':>3s} {'us':>8s} {'grid':>5s} {'tensor%':>8s} {'dram rd MB':>10s} {'dram wr MB':>10s} {'L2 MB':>9s}  layer (group k/s cin->n @h x w, GFLOP, TFLOP/s)"]
    tot = 0
    for i, (k, d) in enumerate(per.items()):
        tot += d["gpu__time_duration.sum"]
        text.append(f"{i:3d} {d['gpu__time_duration.sum']:8.1f} {int(d['launch__grid_size']):5d} "
                    f"{d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']:8.1f} "
                    f"{d['dram__bytes_read.sum']:10.1f} {d['dram__bytes_write.sum']:10.1f} {d['lts__t_bytes.sum']:9.1f}"
                    + (f"  {ops[i]['group']:5s} {ops[i]['k']}x{ops[i]['k']}/{ops[i]['s']} {ops[i]['cin']:4d}->{ops[i]['n']:3d}"
                       f"{'+p' + str(ops[i]['pred']) if ops[i]['pred'] else ('*W/img' if ops[i].get('batched') else ''):6s} @{ops[i]['h']}x{ops[i]['w']} {ops[i]['gflop']:7.1f} GF "
                       f"{ops[i]["gflop"] / d["gpu__time_duration.sum"] * 1e3:7.1f}" if ops and i < len(ops) else ""))
    text.append(f"total {tot:.1f} us over {len(per)} launches")
    s = "\n".join(text)
    print(s)
    if out:
        open(out, "w").write(s + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else None)
