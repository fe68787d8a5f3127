#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list (one row per kernel launch) into a per-kernel
table: launches, total device time, share of the total.  The absolute times are cold-cache and serialised (ncu replays
each launch in isolation), so only the SHARES are meaningful (B200_PROFILING.md).

    python scripts/summarise_launches.py gpurun_out/launches.csv [--out profiles/x.md] [--mine accumulate,extract,...]
"""
from __future__ import annotations

import argparse
import csv
import re
import sys

MINE = ("mss::", "accumulate_kernel", "accumulate_cells_kernel", "accumulate_rows_kernel", "extract_", "finalize_", "vote_", "dice_kernel", "halo_add", "importance", "gaussian_profile",
        "resample_", "intensity_kernel")


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--out", default=None)
    ap.add_argument("--top", type=int, default=25)
    args = ap.parse_args()
    rows = []
    with open(args.csv, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        if len(r) != len(hdr):
            continue
        rows.append(r)
    if hdr is None:
        sys.exit("no ncu CSV header found")
    i_name, i_val, i_unit, i_metric = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    agg, total, n = {}, 0.0, 0
    for r in rows:
        if "gpu__time_duration" not in r[i_metric]:
            continue
        us = float(r[i_val].replace(",", "")) * scale.get(r[i_unit], 1.0)
        name = re.sub(r"\(.*", "", r[i_name])
        name = re.sub(r"<.*", "", name)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
        n += 1
    out = [f"# launch list summary of {args.csv}", "",
           f"{n} kernel launches, {total / 1e3:.2f} ms of summed device time under ncu (cold-cache, serialised: compare shares).", "",
           "| kernel | launches | total ms | share % | ours |", "|---|---|---|---|---|"]
    mine_total = 0.0
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        ours = any(m in name for m in MINE)
        if ours:
            mine_total += us
    shown = 0
    for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        ours = any(m in name for m in MINE)
        if shown < args.top or ours:
            out.append(f"| {name[:90]} | {cnt} | {us / 1e3:.3f} | {100 * us / total:.3f} | {'yes' if ours else ''} |")
            shown += 1
    out += ["", f"libmss_b200.so kernels: {mine_total / 1e3:.3f} ms = {100 * mine_total / total:.3f} % of the summed device time; "
            f"the rest is the backbone (PyTorch / cuDNN / cuBLAS kernels of the caller's module)."]
    text = "\n".join(out) + "\n"
    if args.out:
        open(args.out, "w").write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
