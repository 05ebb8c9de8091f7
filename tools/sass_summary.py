#!/usr/bin/env python
"""cuobjdump -sass of the built library -> profiles/sass_summary_rNN.md: per kernel the instruction count and the counts of
the opcodes that say what the kernel is made of (FP64 / FP32 FMA, shared / global memory, TMA bulk copies, mbarriers,
cp.async, tensor-core and TMEM instructions).  Usage: python tools/sass_summary.py <lib.so> <out.md>"""
import collections
import re
import subprocess
import sys

WATCH = ["DFMA", "DADD", "DMUL", "FFMA", "F2F", "LDS", "STS", "LDG", "STG", "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "BAR", "RED", "ATOM",
         "SHFL", "UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "HMMA", "DMMA"]


def main(lib, out):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    archs = set()
    for line in txt.splitlines():
        m = re.match(r"\s*arch = (sm_\w+)", line)
        if m:
            archs.add(m.group(1))
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur["_total"] += 1
            op = m.group(1)
            cur[op] += 1
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    with open(out, "w") as f:
        f.write(f"`cuobjdump -sass {lib}`: {len(kernels)} kernels, arch {sorted(archs)}, {tot['_total']} SASS instructions.\n\n")
        f.write("Library totals: " + ", ".join(f"{k} {tot[k]}" for k in WATCH) + "\n\n")
        f.write("No tensor-core / TMEM instruction (UTC*MMA, LDTM, STTM, HMMA, DMMA) is expected: DESIGN.md section 3d.\n\n")
        cols = ["DFMA", "DADD", "DMUL", "FFMA", "F2F", "LDS", "STS", "LDG", "STG", "LDGSTS", "UBLKCP", "SYNCS", "BAR", "SHFL"]
        f.write("| kernel | SASS | " + " | ".join(cols) + " |\n|---|---:|" + "---:|" * len(cols) + "\n")
        for n, c in sorted(kernels.items(), key=lambda kv: -kv[1]["_total"]):
            f.write(f"| `{n}` | {c['_total']} | " + " | ".join(str(c[k]) for k in cols) + " |\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
