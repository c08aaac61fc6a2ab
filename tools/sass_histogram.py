#!/usr/bin/env python
"""SASS evidence for the shipped library: per kernel, how many Blackwell-native instructions it contains
(B200_PROFILING.md, "What proves a Blackwell-native kernel").  Runs on the authoring box (cuobjdump only).

    python tools/sass_histogram.py > profiles/rNN_sass_histogram.md
"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "goal_force_b200" / "_lib" / "libgoalforce_b200.so"
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU.EX2",
         "FFMA2", "HMMA", "LDG", "STG", "LDS", "STS", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels: "OrderedDict[str, Counter]" = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            cur = kernels.setdefault(name, Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                if w != "UTCHMMA.2CTA" and (op == w or op.startswith(w + ".")):
                    cur[w] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                cur["UTCHMMA.2CTA"] += 1
    print(f"# SASS opcode counts per kernel ({LIB.name}, sm_100a)\n")
    print("`cuobjdump -sass` of the shipped library; static instruction counts (not executed counts). UTCHMMA = "
          "tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA load, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops. "
          "No HMMA (legacy mma.sync) anywhere.\n")
    cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "SYNCS", "MUFU.EX2", "FFMA2", "HMMA", "LDG", "STG"]
    print("| kernel | instr | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    tot = Counter()
    for name, c in kernels.items():
        short = name.replace("gf::", "")
        print(f"| `{short}` | {c['_total']} | " + " | ".join(str(c[k]) for k in cols) + " |")
        tot.update(c)
    print(f"| **all kernels** | {tot['_total']} | " + " | ".join(str(tot[k]) for k in cols) + " |")


if __name__ == "__main__":
    sys.exit(main())
