#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total
device time and share.  (Per-launch times under ncu are cold-cache and serialised: compare SHARES.)

    python tools/summarize_launches.py gpurun_out/launches.csv [out.md]
"""
import csv
import re
import sys
from collections import defaultdict


def main(path, out=None):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui].strip(), 1e-3)
        rows.append((r[ki], v * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for k, us in rows:
        k = re.sub(r"\(.*", "", k)
        k = re.sub(r"^void ", "", k)[:100]
        agg[k][0] += 1
        agg[k][1] += us
    tot = sum(v[1] for v in agg.values())
    lines = [f"# ncu launch list summary ({path}): {len(rows)} launches, {tot / 1e3:.2f} ms device time", "",
             "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {us / 1e3:.3f} | {100 * us / tot:.1f}% |")
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
