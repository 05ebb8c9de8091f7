#!/usr/bin/env python
"""Warp-state samples of one kernel of an `ncu --set full --import-source on` report -> a small json for profiles/:
share of every stall reason, share of the samples by SASS opcode, shared-memory wavefront excess, and the hottest
CUDA-C source lines (samples per file:line).  Usage: python tools/ncu_stalls.py <ncu-rep> <out.json> [kernel-regex]"""
import collections
import csv
import io
import json
import re
import subprocess
import sys


def page(rep, what, extra=()):
    cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", what, *extra]
    return subprocess.run(cmd, capture_output=True, text=True).stdout


def main(rep, out, kre=None):
    txt = page(rep, "sass")
    lines = txt.splitlines()
    kernel = lines[0].split(",", 1)[1].strip().strip('",') if lines and lines[0].startswith('"Kernel Name"') else "?"
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    col = {n: i for i, n in enumerate(hdr)}
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    stalls, ops = collections.Counter(), collections.Counter()
    total = 0
    wf = wf_ex = 0.0
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        n = int(r[col["# Samples"]] or 0)
        total += n
        for s in stall_cols:
            stalls[s] += int(r[col[s]] or 0)
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", r[col["Source"]])
        if m:
            op = m.group(1)
            op = ".".join(op.split(".")[:3]) if op.startswith("F2F") else ".".join(op.split(".")[:2])
            ops[op] += n
        wf += float(r[col["L1 Wavefronts Shared"]] or 0)
        wf_ex += float(r[col["L1 Wavefronts Shared Excessive"]] or 0)
    res = {"report": rep, "kernel": kernel, "samples": total,
           "stall_reasons_share": {k: round(v / max(sum(stalls.values()), 1), 4) for k, v in stalls.most_common(10)},
           "samples_by_opcode_share": {k: round(v / max(total, 1), 4) for k, v in ops.most_common(14)},
           "shared_wavefronts": wf, "shared_wavefronts_excessive": wf_ex, "excessive_share": round(wf_ex / wf, 4) if wf else None}
    # CUDA-C view (cuda,sass): rows with a line number carry the samples of all SASS instructions of that source line
    try:
        hot = []
        cur_file, hdr2 = None, None
        for ln in page(rep, "cuda,sass").splitlines():
            if ln.startswith('"File Path"'):
                cur_file = next(csv.reader([ln]))[1]
                continue
            if ln.startswith('"Function Name"'):
                continue
            r = next(csv.reader([ln]))
            if r and r[0] == "Line No":
                hdr2 = r
                continue
            if hdr2 and len(r) == len(hdr2) and r[0].isdigit():
                n = r[hdr2.index("# Samples")]
                n = int(n) if n.isdigit() else 0
                if n:
                    hot.append((n, cur_file.split("/")[-1], r[0], r[1].strip()[:120]))
        hot.sort(reverse=True)
        res["hottest_source_lines"] = [{"share": round(n / max(total, 1), 4), "where": f"{f}:{l}", "source": s2} for n, f, l, s2 in hot[:30]]
    except Exception as e:                      # the CUDA view is a convenience; never lose the SASS summary over it
        res["hottest_source_lines_error"] = repr(e)
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1)[:3000])


if __name__ == "__main__":
    main(*sys.argv[1:])
