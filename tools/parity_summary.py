#!/usr/bin/env python
"""gpurun_out/parity_log.jsonl (written by tests/helpers.py::record during `pytest -m gpu`) -> profiles/parity_rNN.md.
The last record of every (test, op, mode, ...) key wins.  Usage: python tools/parity_summary.py <jsonl> <out.md> [title]"""
import collections
import json
import sys


def fmt(v):
    if isinstance(v, float):
        return f"{v:.3g}"
    if isinstance(v, list):
        return "[" + ", ".join(fmt(x) for x in v) + "]"
    return str(v)


def main():
    src, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else "Parity numbers measured on B200"
    rows = collections.OrderedDict()
    for line in open(src):
        line = line.strip()
        if not line:
            continue
        r = json.loads(line)
        if r.get("device") != "cuda":
            continue
        key = "/".join(str(r[k]) for k in ("test", "cfg", "op", "mode", "iters", "sigma", "dct") if k in r)
        rows[key] = {k: v for k, v in r.items() if k not in ("test", "device", "op", "mode", "cfg", "errs") and not (k in ("niter", "ref_niter") and isinstance(v, list) and len(v) > 20)}
    with open(out, "w") as f:
        f.write(f"# {title}\n\nFrom `{src}` (tests/helpers.py::record): relative L2 errors of the CUDA path (through the C ABI, 3x256x256) "
                "against oracle/fh_oracle.py (float64 / complex128) and the golden vectors produced by the real reference.  "
                "`mode` = CG implementation (`auto`: float64 resident kernel fh_k_cg_res64, `streaming`: float32 five-launch iteration, `resident_f32`: the float32 resident kernel of round 1).\n\n"
                "| test | values |\n|---|---|\n")
        for k, v in rows.items():
            f.write(f"| {k} | " + ", ".join(f"{a}={fmt(b)}" for a, b in v.items()) + " |\n")
    print(f"{len(rows)} rows -> {out}")


if __name__ == "__main__":
    main()
