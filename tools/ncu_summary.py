#!/usr/bin/env python
"""ncu_summary.py REPORT.ncu-rep [OUT.txt] -- the handful of numbers DESIGN.md / bench.py quote from an
`ncu --set full` capture: duration, DRAM bytes (-> roofline.traffic), pipe utilisation, occupancy, stall reasons.
Reads the report with `ncu -i ... --page raw --csv` (no GPU needed)."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"), ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__maximum_warps_per_active_cycle_pct", "theoretical occupancy %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    lines = []
    for vals in rows[2:]:
        m = {h: (v, u) for h, v, u in zip(hdr, vals, units)}
        seen = set()
        for k, label in KEYS:
            if k in m and label not in seen:
                seen.add(label)
                lines.append(f"{label:32s} {m[k][0]} {m[k][1]}")
        try:
            rd = float(m["dram__bytes_read.sum"][0]); wr = float(m["dram__bytes_write.sum"][0])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
            tot = rd * scale[m["dram__bytes_read.sum"][1]] + wr * scale[m["dram__bytes_write.sum"][1]]
            dur = float(m["gpu__time_duration.sum"][0]) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}[m["gpu__time_duration.sum"][1]]
            lines.append(f"{'dram traffic (read+write)':32s} {tot:.0f} byte  = {tot / dur / 1e9:.0f} GB/s under ncu")
        except Exception:
            pass
        stalls = sorted(((float(v[0]), h) for h, v in m.items()
                         if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and v[0]), reverse=True)
        lines.append("stall reasons (warps stalled per issue-active cycle):")
        for val, h in stalls[:6]:
            lines.append(f"    {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} {val:.2f}")
        lines.append("")
    txt = "\n".join(lines)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(f"# {rep} (ncu --set full --clock-control none), summarised by tools/ncu_summary.py\n" + txt)
    print(txt)


if __name__ == "__main__":
    main()
