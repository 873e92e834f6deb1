"""Top SASS instructions of an .ncu-rep by stall samples / executed count (reads `ncu --page source --csv`)."""
import csv
import subprocess
import sys


def main(rep, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]
    ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    data = []
    for r in rows[2:]:
        if len(r) < len(h):
            continue
        st = sorted(((int(r[i] or 0), n) for i, n in stall_cols), reverse=True)[:2]
        data.append((int(r[isamp] or 0), int(r[iex] or 0), r[isrc].strip(), st))
    tot = sum(d[0] for d in data)
    totex = sum(d[1] for d in data)
    print(f"total samples {tot}, total warp instructions {totex}")
    print("--- by samples")
    for i, d in sorted(enumerate(data), key=lambda x: -x[1][0])[:top]:
        print(f"{i:5d} {d[0]:7d} {100*d[0]/tot:5.1f}%  ex={d[1]:9d}  {d[2][:70]:70s} {d[3]}")
    print("--- by executed")
    for i, d in sorted(enumerate(data), key=lambda x: -x[1][1])[:top // 2]:
        print(f"{i:5d} {d[0]:7d}  ex={d[1]:9d} {100*d[1]/totex:5.1f}%  {d[2][:70]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
