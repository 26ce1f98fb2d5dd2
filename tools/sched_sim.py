#!/usr/bin/env python
"""Lane utilisation of the production traversal scheduler (csrc/rtb_trace.cuh) on emulated warps,
without a GPU: builds a sphere field of the C5 generator, makes a wavefront-like ray stream
(camera rays in sample-major order, then the diffuse bounces off their hit points in the same
order) and runs it through tests/hostcheck precision 38, sweeping the vote thresholds.

  python tools/sched_sim.py [--half 60] [--rays 40000]

Prints, per setting: node steps and the mean lanes taking part, the same for primitive steps, and
the instruction-weighted utilisation for a node step of ~100 and a primitive step of ~70 instructions.
"""
import argparse
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
PKG = "ray_tracing-rendering_b200"


def load_hostcheck():
    so = os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so")
    L = C.CDLL(so)
    L.hc_scene_create.restype = C.c_void_p
    L.hc_scene_create.argtypes = [C.c_char_p, C.c_uint64, C.c_int]
    L.hc_scene_destroy.argtypes = [C.c_void_p]
    L.hc_trace_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    L.hc_sched_stats.argtypes = [C.c_void_p]
    L.hc_sched_tuning.argtypes = [C.c_int, C.c_int]
    L.hc_sched_leaf_min.argtypes = [C.c_int]
    L.hc_sched_sim_warps.argtypes = [C.c_int]
    L.hc_sched_postpone.argtypes = [C.c_int]
    L.hc_sched_sim_warps(1)  # full-size windows, as in a long launch
    return L


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--half", type=int, default=60)
    ap.add_argument("--rays", type=int, default=40000)
    ap.add_argument("--scene", default="field")
    a = ap.parse_args()
    abi = importlib.import_module(PKG + ".abi")
    scenes = importlib.import_module(PKG + ".scenes")
    L = load_hostcheck()
    if a.scene == "field":
        blob = scenes.sphere_field(half_extent=a.half, width=256, height=144, spp=1)
        lookfrom = np.array([13.0, 2.0, 3.0]) * (a.half / 12.5)
    else:
        blob = scenes.final_scene(1)
        lookfrom = np.array([478.0, 278.0, -600.0])
    h = L.hc_scene_create(blob, len(blob), 4)
    rng = np.random.default_rng(1)
    # camera rays: a 2-D block of neighbouring pixels, row-major (sample-major order of the renderer)
    n = a.rays
    side = int(np.sqrt(n))
    n = side * side
    if a.scene == "field":
        target = np.zeros(3)
        span = 0.12
    else:
        target = np.array([278.0, 278.0, 0.0])
        span = 0.35
    w = (lookfrom - target) / np.linalg.norm(lookfrom - target)
    u = np.cross([0, 1, 0], w)
    u /= np.linalg.norm(u)
    v = np.cross(w, u)
    ii, jj = np.meshgrid(np.arange(side), np.arange(side))
    su = (ii.ravel() / side - 0.5) * span
    sv = (jj.ravel() / side - 0.5) * span
    d = -w[None, :] + su[:, None] * u[None, :] + sv[:, None] * v[None, :]
    rays = np.zeros(n, abi.RAY)
    rays["o"] = lookfrom
    rays["d"] = d
    rays["t_min"], rays["t_max"], rays["origin_prim"] = 0.001, np.inf, -1

    def run(r, prec):
        out = np.zeros(len(r), abi.HIT)
        st = np.zeros(2, np.uint64)
        L.hc_trace_batch(h, ptr(r), len(r), prec, ptr(out), ptr(st))
        return out, st

    def report(name, r):
        for node_min, switch_min, leaf_min in [(16, 8, 0), (20, 8, 0), (16, 8, 33), (16, 8, 16), (16, 8, 12), (16, 8, 10), (16, 8, 8), (20, 8, 12), (12, 8, 12), (16, 4, 12)]:
            L.hc_sched_tuning(node_min, switch_min)
            L.hc_sched_leaf_min(leaf_min)
            L.hc_sched_postpone(1 if leaf_min else 0)  # leaf_min 0: leaves are not parked (round 2's first scheduler)
            s = np.zeros(6, np.uint64)
            L.hc_sched_stats(ptr(s))
            _, st = run(r, 38)
            L.hc_sched_stats(ptr(s))
            s = s.astype(float)
            cn, cl = 100.0, 70.0
            useful = s[1] * cn + s[3] * cl
            issued = 32 * (s[0] * cn + s[2] * cl)
            print(f"{name:>10} node_min {node_min:2d} switch_min {switch_min:2d} leaf_min {leaf_min:2d}: node steps {int(s[0]):8d} x {s[1] / max(s[0], 1):5.1f} lanes,"
                  f" prim steps {int(s[2]):8d} x {s[3] / max(s[2], 1):5.1f} lanes, refills {int(s[4]):6d} x {s[5] / max(s[4], 1):4.1f};"
                  f" utilisation {useful / issued:.3f}; nodes/ray {st[0] / len(r):.1f} prims/ray {st[1] / len(r):.2f}", flush=True)

    report("camera", rays)
    hit, _ = run(rays, 37)
    ok = hit["prim"] >= 0
    p = rays["o"][ok] + hit["t"][ok, None] * rays["d"][ok]
    nb = len(p)
    dd = rng.normal(size=(nb, 3))
    dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    dd[:, 1] = np.abs(dd[:, 1])  # off the ground / upper hemisphere: what a diffuse bounce looks like here
    b = np.zeros(nb, abi.RAY)
    b["o"] = p + 1e-3 * dd
    b["d"] = dd
    b["t_min"], b["t_max"], b["origin_prim"] = 0.001, np.inf, -1
    report("bounce", b)
    L.hc_scene_destroy(h)


if __name__ == "__main__":
    main()
