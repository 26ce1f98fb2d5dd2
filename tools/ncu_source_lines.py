#!/usr/bin/env python
"""Hot source lines of one profiled kernel straight from the report (`ncu --set full --import-source on`):
executed warp instructions, active lanes and stall samples per CUDA source line, aggregated by ncu itself
(works whatever the tree has been rebuilt to since, unlike tools/ncu_hotspots.py).

  python tools/ncu_source_lines.py gpurun_out/r02d_c5_extend.ncu-rep [--top 40] [--by samples]
"""
import argparse
import csv
import io
import os
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--by", default="inst", choices=["inst", "samples"])
    a = ap.parse_args()
    if a.report.endswith(".csv"):        # the source page exported on the GPU box (tools/gpu_g.sh)
        out = open(a.report).read()
    else:
        out = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                             capture_output=True, text=True).stdout
    rows, fname, hdr = [], "?", None
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Name":
            fname = os.path.basename(r[1])
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        if r[2] != "-":        # a SASS row; the source rows carry the aggregates
            continue
        d = dict(zip(hdr, r))
        d["Source"] = r[1]
        try:
            inst = float(d.get("Instructions Executed") or 0)
            thr = float(d.get("Thread Instructions Executed") or 0)
            smp = float(d.get("# Samples") or 0)
        except ValueError:
            continue
        if inst or smp:
            rows.append((inst, thr, smp, fname, d["Line No"], d["Source"].strip(), d))
    ti, ts = sum(r[0] for r in rows), sum(r[2] for r in rows)
    print(f"# executed warp instructions {ti:.0f}; average active lanes {sum(r[1] for r in rows) / max(ti, 1):.1f}; stall samples {ts:.0f}")
    print("# inst%  lanes  samples%  top stall      file:line  source")
    rows.sort(key=lambda r: -(r[0] if a.by == "inst" else r[2]))
    for inst, thr, smp, f, ln, src, d in rows[:a.top]:
        stalls = {k[6:]: float(v or 0) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k}
        top = max(stalls, key=stalls.get) if stalls and max(stalls.values()) > 0 else "-"
        print(f"{100 * inst / max(ti, 1):6.2f}  {thr / max(inst, 1):5.1f}  {100 * smp / max(ts, 1):7.2f}  {top:<14} {f}:{ln}  {src[:90]}")


if __name__ == "__main__":
    main()
