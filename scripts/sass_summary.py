#!/usr/bin/env python
"""profiles/sass_summary.md: per kernel of libmss_b200.so, counts of the SASS mnemonics that prove what the sources claim
(TMA bulk tensor copies, asynchronous global->shared copies, mbarrier traffic, vector width).  Runs where cuobjdump is
installed; no GPU needed.

    python scripts/sass_summary.py > profiles/sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "medicalsemseg_b200", "lib", "libmss_b200.so")
WATCH = ["UTMALDG", "UTMASTG", "UBLKCP", "UTMACCTL", "SYNCS", "LDGSTS", "LDGDEPBAR", "DEPBAR", "LDG.E.128", "STG.E.128", "LDS.128",
         "ATOM", "RED", "BAR.SYNC", "FMUL", "FADD", "FFMA", "POPC", "LOP3", "MUFU"]


def main() -> None:
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op.startswith(w) or (w in ("LDG.E.128", "STG.E.128", "LDS.128") and w.split(".")[0] in op and ".128" in op
                                        and op.startswith(w.split(".")[0])):
                    kernels[cur][w] += 1
    arch = re.findall(r"arch = (sm_\w+)", out)
    print("# SASS summary of medicalsemseg_b200/lib/libmss_b200.so\n")
    print(f"`cuobjdump -sass`: {len(kernels)} kernels, architectures {sorted(set(arch))}.  Counts are static instruction counts.\n")
    print("| kernel | SASS instr | " + " | ".join(WATCH) + " |")
    print("|---|---|" + "---|" * len(WATCH))
    for name, c in kernels.items():
        short = re.sub(r"\(.*\)$", "", name)
        short = short.replace("mss::", "")
        if len(short) > 90:
            short = short[:87] + "..."
        print(f"| `{short}` | {c['_total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in WATCH) + " |")


if __name__ == "__main__":
    sys.exit(main())
