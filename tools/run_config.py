#!/usr/bin/env python
"""Renders one BASELINE configuration (C1..C5) at a chosen spp and prints the
stage timing; the command ncu wraps when a profile of a configuration is wanted.

  python tools/run_config.py C5 --spp 4 [--pool N] [--wavefront] [--time] [--reps 2]
"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "ray_tracing-rendering_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--wavefront", action="store_true")
    ap.add_argument("--fused", action="store_true")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--count", action="store_true")
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--warm", type=int, default=1)
    ap.add_argument("--binary", action="store_true", help="round 1's traversal kernels (RTB_OPT_BINARY_TRAVERSAL = 1)")
    ap.add_argument("--wide", action="store_true", help="the warp-scheduled 4-wide kernels whatever the scene (RTB_OPT_BINARY_TRAVERSAL = 2)")
    ap.add_argument("--nogroup", action="store_true", help="every rect of a box object its own tree item (RTB_OPT_GROUP_BOXES = 0)")
    a = ap.parse_args()
    pkg = importlib.import_module(PKG)
    configs = importlib.import_module(PKG + ".configs")
    binding = importlib.import_module(PKG + ".binding")
    cfg = configs.get(a.config)
    spp = a.spp or cfg.spp
    ctx = pkg.Context(0)
    if a.binary or a.wide:
        ctx.set_option(binding.OPT_BINARY_TRAVERSAL, 1 if a.binary else 2)
    if a.nogroup:
        ctx.set_option(binding.OPT_GROUP_BOXES, 0)
    ctx.upload_scene(cfg.blob())
    flags = (binding.RENDER_FORCE_WAVEFRONT if a.wavefront else 0) | (binding.RENDER_FORCE_FUSED if a.fused else 0) | (binding.RENDER_TIME_EXTEND if a.time else 0) \
        | (binding.RENDER_COUNT_VISITS if a.count else 0)
    for _ in range(a.warm):
        ctx.render(ctx.params(cfg.width, cfg.height, min(spp, 2), cfg.integrator, cfg.depth, pool_paths=a.pool))
    for r in range(a.reps):
        _, st = ctx.render(ctx.params(cfg.width, cfg.height, spp, cfg.integrator, cfg.depth, seed=3 + r,
                                      pool_paths=a.pool, flags=flags))
        rays = st["rays_closest"] + st["rays_shadow"]
        line = (f"{cfg.name} {cfg.width}x{cfg.height} spp {spp} int {cfg.integrator}: {st['device_ms']:.2f} ms "
                f"{st['paths'] / st['device_ms'] / 1e3:.1f} Mpaths/s {rays / st['device_ms'] / 1e3:.1f} Mrays/s "
                f"iters {st['iterations']} launches {st['kernel_launches']} schedule {st['schedule']} "
                f"rays {st['rays_closest'] / 1e6:.1f}M+{st['rays_shadow'] / 1e6:.1f}M")
        if a.time:
            line += f" stage_ms {[round(x, 2) for x in st['stage_ms']]} extend_ms {st['extend_ms']:.2f}"
        if a.count:
            line += (f" nodes/ray {st['nodes_visited'] / max(rays, 1):.2f} prims/ray {st['prim_tests'] / max(rays, 1):.2f} "
                     f"max nodes/ray {st['max_nodes_per_ray']} chunk-sync lane utilisation bound "
                     f"{st['extend_nodes'] / max(32 * st['extend_chunk_max_nodes'], 1):.3f}")
        print(line, flush=True)


if __name__ == "__main__":
    main()
