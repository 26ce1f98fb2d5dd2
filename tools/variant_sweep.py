#!/usr/bin/env python
"""Times every library under ray_tracing-rendering_b200/variants/ (tools/build_variants.sh) and the
default build on the given configurations: one subprocess per (library, configuration), so each
run loads exactly one build through RTB200_LIBRARY.

  python tools/variant_sweep.py C1 C3 [--reps 3]
"""
import argparse
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--spp", type=int, default=0)
    a = ap.parse_args()
    libs = [("default", "")] + [(os.path.basename(p)[len("librtb200_"):-3], p)
                                for p in sorted(glob.glob(os.path.join(ROOT, "ray_tracing-rendering_b200", "variants", "librtb200_*.so")))]
    for cfg in a.configs:
        for name, path in libs:
            env = dict(os.environ)
            if path:
                env["RTB200_LIBRARY"] = path
            cmd = [sys.executable, os.path.join(ROOT, "tools", "run_config.py"), cfg, "--reps", str(a.reps)]
            if a.spp:
                cmd += ["--spp", str(a.spp)]
            try:
                out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
                lines = [l for l in out.stdout.splitlines() if " ms " in l]
                ms = sorted(float(l.split(": ")[1].split(" ms")[0]) for l in lines)
                print(f"{cfg} {name:>12}: " + (" ".join(f"{m:.2f}" for m in ms) if ms else "FAILED " + out.stderr[-300:]), flush=True)
            except subprocess.TimeoutExpired:
                print(f"{cfg} {name:>12}: TIMEOUT", flush=True)


if __name__ == "__main__":
    main()
