import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import numpy as np
pkg = importlib.import_module("ray_tracing-rendering_b200")
scenes = importlib.import_module("ray_tracing-rendering_b200.scenes")
binding = importlib.import_module("ray_tracing-rendering_b200.binding")
import conftest
ctx = pkg.Context(0)
which = sys.argv[1]
if which == "s9":
    blob = conftest.load_golden(9).blob
    t = time.time(); ctx.upload_scene(blob); print("upload s", time.time() - t, ctx.scene_stats())
    for spp in (20, 100):
        for _ in range(2):
            acc, st = ctx.render(ctx.params(800, 800, spp, 1, seed=3))
        print(f"C2 scene9 800x800x{spp}: {st['device_ms']:.1f} ms {st['paths']/st['device_ms']/1e3:.1f} Mpaths/s {(st['rays_closest']+st['rays_shadow'])/st['device_ms']/1e3:.1f} Mrays/s iters {st['iterations']}")
    acc, st = ctx.render(ctx.params(800, 800, 8, 1, seed=3, flags=binding.RENDER_COUNT_VISITS | binding.RENDER_TIME_EXTEND))
    print("nodes/ray", st['nodes_visited']/st['rays_closest'], "tests/ray", st['prim_tests']/st['rays_closest'], "extend share", st['extend_ms']/st['device_ms'])
if which == "c5":
    half = int(sys.argv[2]) if len(sys.argv) > 2 else 500
    t = time.time(); blob = scenes.sphere_field(half, 3840, 2160, 1024); print("build blob s", time.time() - t, len(blob) / 1e6, "MB")
    t = time.time(); ctx.upload_scene(blob); print("upload+BVH s", time.time() - t, ctx.scene_stats())
    for spp in (1, 4):
        for _ in range(2):
            acc, st = ctx.render(ctx.params(3840, 2160, spp, 4, seed=3))
        print(f"C5 {4*half*half} spheres 3840x2160x{spp} int4: {st['device_ms']:.1f} ms {st['paths']/st['device_ms']/1e3:.1f} Mpaths/s {(st['rays_closest']+st['rays_shadow'])/st['device_ms']/1e3:.1f} Mrays/s iters {st['iterations']} mean {acc[...,:3].mean(axis=(0,1))/spp}")
    acc, st = ctx.render(ctx.params(3840, 2160, 1, 4, seed=3, flags=binding.RENDER_COUNT_VISITS | binding.RENDER_TIME_EXTEND))
    print("nodes/ray", st['nodes_visited']/(st['rays_closest']+st['rays_shadow']), "tests/ray", st['prim_tests']/(st['rays_closest']+st['rays_shadow']), "extend share", st['extend_ms']/st['device_ms'])
