#!/usr/bin/env python
"""Per-kernel SASS opcode evidence (B200_PROFILING.md "What proves a Blackwell-native kernel"): counts of the tcgen05 / TMEM / TMA
mnemonics (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = cp.async.bulk.tensor, UTCBAR = tcgen05.commit,
SYNCS = mbarrier) and of the legacy HMMA (mma.sync) in every kernel of the built library, plus registers / spills from ptxas.
No GPU needed.  Usage: python tools/sass_opcodes.py > profiles/sass_opcodes_r02.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "instructany2pix_b200", "libia2p_sm100a.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "MUFU", "LDG", "STG", "LDS", "STS", "R2UR"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
counts, cur = collections.OrderedDict(), None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = counts.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur is not None:
        op = m.group(1)
        for o in OPS:
            if op == o or op.startswith(o + "."):
                key = o
                if o == "UTCHMMA" and ".2CTA" in op:
                    key = "UTCHMMA.2CTA"
                if o in ("LDG", "STG") and ".128" in op:
                    key = o + ".128"
                cur[key] += 1
        cur["total"] += 1
cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "MUFU", "LDG.128", "STG.128", "R2UR", "total"]
print(f"# SASS opcode counts per kernel of {os.path.basename(LIB)} (cuobjdump -sass; sm_100a)")
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---:|" * len(cols))
for k, c in sorted(counts.items(), key=lambda kv: (-(kv[1]["UTCHMMA"] + kv[1]["UTCHMMA.2CTA"]), kv[0])):
    print(f"| `{k}` | " + " | ".join(str(c.get(x, 0)) for x in cols) + " |")
