#!/usr/bin/env python
"""Turns the ncu outputs brought back in gpurun_out/ into the small, committed files under profiles/:
  * a launch list (gpu__time_duration per launch) -> per-kernel count / mean / share table (CSV in, markdown + json out)
  * an `ncu --set full` report (.ncu-rep) -> per-kernel DRAM bytes, L2 hit rate, registers, shared memory (json)
Usage: python tools/ncu_summary.py launches <csv> <out_prefix> | full <ncu-rep> <out_json> [label]"""
import collections
import csv
import json
import subprocess
import sys


def launches(path, out_prefix):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    acc = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        if r[ui] in ("ns", "nsecond"):
            v /= 1e3
        acc.setdefault(name, []).append(v)
    total = sum(sum(v) for v in acc.values())
    table = sorted(((n, len(v), sum(v) / len(v), sum(v), sum(v) / total) for n, v in acc.items()), key=lambda t: -t[3])
    with open(out_prefix + ".md", "w") as f:
        f.write(f"ncu launch list `{path}`: {sum(t[1] for t in table)} launches, {total / 1e3:.2f} ms of kernel time "
                "(gpu__time_duration.sum, --clock-control none; cold-cache and serialised: compare shares, not absolutes)\n\n")
        f.write("| kernel | launches | mean us | total us | share |\n|---|---:|---:|---:|---:|\n")
        for n, c, m, tot, sh in table:
            f.write(f"| `{n}` | {c} | {m:.2f} | {tot:.0f} | {sh:.3f} |\n")
    json.dump({"source": path, "kernels": {n: {"launches": c, "mean_us": m, "share": sh} for n, c, m, tot, sh in table}},
              open(out_prefix + ".json", "w"), indent=1)
    print(open(out_prefix + ".md").read())


def full(rep, out_json, label=""):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    want = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
            "lts__t_sector_hit_rate.pct": "l2_hit_pct", "launch__registers_per_thread": "registers",
            "launch__grid_size": "grid", "launch__shared_mem_per_block_dynamic": "smem_dynamic_kb",
            "launch__shared_mem_per_block_static": "smem_static_kb", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
            "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct"}
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
    acc = collections.OrderedDict()
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
        d = acc.setdefault(name, collections.defaultdict(list))
        for k, short in want.items():
            if k in col:
                v = float(r[col[k]].replace(",", ""))
                if short.startswith("dram_r") or short.startswith("dram_w"):
                    v *= scale.get(units[col[k]], 1.0)
                d[short].append(v)
    out = {"source": rep, "label": label, "kernels": {}}
    for n, d in acc.items():
        k = {s: sum(v) / len(v) for s, v in d.items()}
        k["launches_profiled"] = len(d["duration_us"])
        k["dram_bytes_per_launch"] = k.pop("dram_read", 0.0) + k.pop("dram_write", 0.0)
        out["kernels"][n] = k
    json.dump(out, open(out_json, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
