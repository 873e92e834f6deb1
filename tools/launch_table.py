"""Per-launch table of an `ncu --metrics ... --csv` log (all kernels, launch order) joined with the plan's conv op list,
plus a per-kernel summary.  Usage: python tools/launch_table.py launches.csv [step_ops.json] [out.txt]"""
import collections
import csv
import json
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = per.setdefault(int(r["ID"]), {"name": r["Kernel Name"].split("(")[0].replace("glsdet::", ""), "grid": r["Grid Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit, name = r["Metric Unit"], r["Metric Name"]
        if name == "gpu__time_duration.sum":
            v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "usecond": v, "nsecond": v / 1e3, "msecond": v * 1e3}.get(unit, v)
        if name.startswith("dram__bytes") or name.startswith("lts__t_bytes"):
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1) / 1e6
        d[name] = v
    return per


def main(path, ops_path=None, out=None):
    per = load(path)
    ops = json.load(open(ops_path)) if ops_path else []
    text = [f"{'#':>3s} {'us':>8s} {'tensor%':>8s} {'rd MB':>8s} {'wr MB':>8s}  kernel / layer"]
    tot = 0.0
    ci = 0
    agg = collections.OrderedDict()
    rd = wr = 0.0
    for i, d in per.items():
        t = d["gpu__time_duration.sum"]
        tot += t
        a = agg.setdefault(d["name"], [0.0, 0]); a[0] += t; a[1] += 1
        extra = ""
        if "conv_gemm" in d["name"] and ci < len(ops):
            o = ops[ci]; ci += 1
            extra = (f"  {o['group']:5s} {o['k']}x{o['k']}/{o['s']} {o['cin']:4d}->{o['n']:3d}{'+p' + str(o['pred']) if o['pred'] else ''} "
                     f"@{o['h']}x{o['w']} {o['gflop']:6.1f} GF {o['gflop'] / t * 1e3:7.1f} TF/s")
        r_, w_ = d.get('dram__bytes_read.sum', 0), d.get('dram__bytes_write.sum', 0)
        rd += r_; wr += w_
        text.append(f"{i:3d} {t:8.1f} {d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):8.1f} "
                    f"{r_:8.1f} {w_:8.1f}  {d['name'][:40]:40s} {d['grid']}{extra}")
    text.append(f"total {tot:.1f} us over {len(per)} launches; dram read {rd:.0f} MB, written {wr:.0f} MB")
    text.append("")
    for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        text.append(f"{n[:56]:56s} {c:4d} {v:10.1f} us {100 * v / tot:6.1f}%")
    s = "\n".join(text)
    print(s)
    if out:
        open(out, "w").write(s + "\n")
    return per


def conv_traffic_json(per, out_json, batch=16):
    """DRAM bytes (read + write) of the conv_gemm launches of the profiled step -> profiles/r2_conv_traffic.json."""
    rd = sum(d.get("dram__bytes_read.sum", 0) for d in per.values() if "conv_gemm" in d["name"]) * 1e6
    wr = sum(d.get("dram__bytes_write.sum", 0) for d in per.values() if "conv_gemm" in d["name"]) * 1e6
    n = sum(1 for d in per.values() if "conv_gemm" in d["name"])
    json.dump({"batch": batch, "conv_launches": n, "dram_read_bytes": rd, "dram_write_bytes": wr, "total_bytes_per_step": rd + wr,
               "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over tools/profile_step.py p0 (one step, cold caches per launch)"},
              open(out_json, "w"), indent=1)


if __name__ == "__main__":
    per_ = main(*sys.argv[1:4])
    if len(sys.argv) > 4:
        conv_traffic_json(per_, sys.argv[4])
