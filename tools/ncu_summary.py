"""Key numbers of one `ncu --set full` capture (.ncu-rep) as text: duration, tensor pipe, issue slots, pipes, DRAM / L2
traffic, registers, stall reasons per issued instruction.  Usage: python tools/ncu_summary.py capture.ncu-rep [out.txt] [title]"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM pipe %"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue slots busy %"),
    ("smsp__issue_active.avg.per_cycle_active", "IPC per scheduler"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU / conversions) pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"),
    ("inst_executed", "warp instructions"),
]


def main(rep, out=None, title=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# {title or rep}: ncu --set full --clock-control none (one launch; numbers under the profiler are not bench values)"]
    for vals in rows[2:]:
        get = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        lines.append(f"kernel: {get['Kernel Name'][0]}   grid {get.get('Grid Size', ('', ''))[0]}  block {get.get('Block Size', ('', ''))[0]}")
        for k, label in KEYS:
            if k in get and get[k][0] != "":
                lines.append(f"  {label:34s} {get[k][0]:>16s} {get[k][1]}")
        lines.append("  stall reasons, warps per issued instruction:")
        st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v.replace(",", "")))
              for h, v in zip(hdr, vals) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v]
        for n, v in sorted(st, key=lambda kv: -kv[1])[:8]:
            lines.append(f"    {n:28s} {v:6.2f}")
    s = "\n".join(lines)
    print(s)
    if out:
        open(out, "w").write(s + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else None)
