#!/usr/bin/env python
"""Executed warp instructions per SOURCE LINE of one profiled kernel: joins the per-instruction
counts of an ncu report (`ncu --set full --import-source on`, source page) with the line table of
the same build (`nvdisasm -g` on the cubin inside csrc/build/rtb_wavefront.o).  The report and the
object must come from the same build.

  python tools/ncu_hotspots.py gpurun_out/r01c_c1_fused.ncu-rep 'k_fusedILb1ELb0ELi2E' [--top 45]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "ray_tracing-rendering_b200", "csrc", "build", "rtb_wavefront.o")


def line_table(mangled_part):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call(["cuobjdump", "-xelf", "all", OBJ], cwd=tmp, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(dis) if l.startswith(".text.") and mangled_part in l][0]
    end = [i for i, l in enumerate(dis) if l.startswith(".text.") and i > start][0]
    cur, table = ("?", 0), {}
    for l in dis[start:end]:
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m:
            table[int(m.group(1), 16)] = cur
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel", help="part of the mangled kernel name, e.g. k_fusedILb1ELb0ELi2E")
    ap.add_argument("--top", type=int, default=45)
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, ie, it = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    ex = {int(r[ia], 16): (int(r[ie]), int(r[it])) for r in rows[2:] if len(r) > it}
    base = min(ex)
    table = line_table(a.kernel)
    assert len(table) == len(ex), f"report has {len(ex)} instructions, this build {len(table)}: not the same build"
    per, per_t = collections.Counter(), collections.Counter()
    for addr, (e, t) in ex.items():
        per[table[addr - base]] += e
        per_t[table[addr - base]] += t
    tot = sum(per.values())
    print(f"# {rows[0][1]}")
    print(f"# executed warp instructions: {tot}; average active lanes {sum(per_t.values()) / tot:.1f}")
    print("# share  lanes  file:line  source")
    files, acc = {}, 0
    for (f, l), e in per.most_common(a.top):
        if f not in files:
            try:
                files[f] = open(f).read().split("\n")
            except OSError:
                files[f] = []
        text = files[f][l - 1].strip()[:100] if 0 < l <= len(files[f]) else ""
        acc += e
        print(f"{e / tot * 100:5.2f}%  {per_t[(f, l)] / e:5.1f}  {os.path.basename(f)}:{l}  {text}")
    print(f"# these lines: {acc / tot * 100:.1f} % of all executed instructions")


if __name__ == "__main__":
    main()
