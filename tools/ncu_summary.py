#!/usr/bin/env python
"""Key metrics per captured launch from an `ncu -i X.ncu-rep --page raw --csv` dump:  python tools/ncu_summary.py raw.csv [out.md]"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU wavefronts %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("launch__registers_per_thread", "registers"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall long_scoreboard (samples)"),
    ("smsp__pcsamp_warps_issue_stalled_barrier", "stall barrier (samples)"),
    ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "stall math_pipe_throttle (samples)"),
    ("smsp__pcsamp_warps_issue_stalled_mio_throttle", "stall mio_throttle (samples)"),
    ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "stall short_scoreboard (samples)"),
    ("smsp__pcsamp_warps_issue_stalled_wait", "stall wait (samples)"),
    ("smsp__pcsamp_warps_issue_stalled_sleeping", "stall sleeping (samples)"),
    ("smsp__pcsamp_warps_issue_stalled_selected", "selected (samples)"),
]


def main(path, out=None):
    rows = list(csv.reader(open(path, newline="")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    lines = [f"# ncu --set full summary ({path})", ""]
    for r in data:
        lines.append(f"## `{r[ki][:110]}`")
        lines.append("")
        lines.append("| metric | value |")
        lines.append("|---|---|")
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                lines.append(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
        lines.append("")
    text = "\n".join(lines)
    if out:
        open(out, "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main(*sys.argv[1:3])
