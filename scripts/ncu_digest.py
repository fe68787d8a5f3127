#!/usr/bin/env python
"""Digest of an Nsight Compute report (run where ncu is installed, no GPU needed):

    python scripts/ncu_digest.py gpurun_out/x.ncu-rep [--top 25] [--out profiles/x.md]

Prints, per profiled launch, the raw-page metrics the roofline argument rests on (duration, DRAM bytes, DRAM / SM
utilisation, occupancy, registers) and the hottest SASS lines of the source page (stall samples + executed counts).
"""
from __future__ import annotations

import argparse
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_lsu.sum",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lts__t_bytes.sum",
    "l1tex__t_bytes.sum",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps",
    "launch__waves_per_multiprocessor",
]


def ncu_csv(rep: str, page: str, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=20)
    ap.add_argument("--out", default=None)
    ap.add_argument("--traffic-key", default=None,
                    help="write launch 0's DRAM bytes (read + write) under this key into profiles/ncu_traffic.json (read by bench.py)")
    args = ap.parse_args()
    buf = io.StringIO()

    def p(*a):
        print(*a, file=buf)

    rows = ncu_csv(args.rep, "raw")
    hdr, units = rows[0], rows[1]
    p(f"# ncu digest of {args.rep}")
    for li, r in enumerate(rows[2:]):
        name = r[hdr.index("Kernel Name")]
        p(f"\n## launch {li}: {name}\n")
        p("| metric | value | unit |\n|---|---|---|")
        for k in RAW_KEYS:
            if k in hdr:
                i = hdr.index(k)
                p(f"| {k} | {r[i]} | {units[i]} |")
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")])
            wr = float(r[hdr.index("dram__bytes_write.sum")])
            ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = rd * scale[ur] + wr * scale[uw]
            dur = float(r[hdr.index("gpu__time_duration.sum")])
            du = units[hdr.index("gpu__time_duration.sum")]
            dur_s = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[du]
            p(f"| DRAM traffic (read+write) | {tot / 1e9:.4f} | GB |")
            if args.traffic_key and li == 0:
                import json
                import os
                path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
                d = json.load(open(path)) if os.path.exists(path) else {}
                d[args.traffic_key] = int(round(tot))
                d.setdefault("_source", {})[args.traffic_key] = f"{args.out or args.rep} ({name}: dram read {rd} {ur} + write {wr} {uw})"
                json.dump(d, open(path, "w"), indent=1)
            p(f"| DRAM traffic / duration (under ncu: cold, serialised) | {tot / dur_s / 1e9:.1f} | GB/s |")
        except Exception as e:  # noqa: BLE001
            p(f"| traffic | n/a ({e}) | |")

    src = ncu_csv(args.rep, "source", ["--print-source", "sass"])
    kernels, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    for li, k in enumerate(kernels):
        h = k["hdr"]
        try:
            i_src, i_smp, i_exe = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        except (ValueError, AttributeError):
            continue
        tot_smp = sum(int(r[i_smp] or 0) for r in k["rows"]) or 1
        tot_exe = sum(int(r[i_exe] or 0) for r in k["rows"]) or 1
        p(f"\n## launch {li} source page: {k['name']}\n")
        p(f"{len(k['rows'])} SASS instructions, {tot_exe} warp-instructions executed, {tot_smp} stall samples\n")
        mix = {}
        for r in k["rows"]:
            op = r[i_src].split()[0] if r[i_src].split() else "?"
            if op.startswith("@"):
                op = r[i_src].split()[1]
            op = op.split(".")[0]
            mix[op] = mix.get(op, 0) + int(r[i_exe] or 0)
        top_mix = sorted(mix.items(), key=lambda kv: -kv[1])[:14]
        p("executed mix: " + ", ".join(f"{o} {100.0 * c / tot_exe:.1f}%" for o, c in top_mix))
        p("\n| samples % | executed | SASS |\n|---|---|---|")
        for r in sorted(k["rows"], key=lambda r: -int(r[i_smp] or 0))[: args.top]:
            p(f"| {100.0 * int(r[i_smp] or 0) / tot_smp:.1f} | {r[i_exe]} | `{r[i_src].strip()}` |")
    text = buf.getvalue()
    if args.out:
        with open(args.out, "w") as f:
            f.write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
