import importlib, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import numpy as np
pkg = importlib.import_module("ray_tracing-rendering_b200")
scenes = importlib.import_module("ray_tracing-rendering_b200.scenes")
import conftest
ctx = pkg.Context(0)
ctx.upload_scene(scenes.sphere_field(500, 3840, 2160, 1024))
for pool in (1 << 20, 1 << 21, 1 << 22, 1 << 23):
    for _ in range(2):
        acc, st = ctx.render(ctx.params(3840, 2160, 8, 4, seed=3, pool_paths=pool))
    print(f"C5 pool {pool}: {st['device_ms']:.1f} ms {st['paths']/st['device_ms']/1e3:.1f} Mpaths/s iters {st['iterations']}")
ctx.upload_scene(conftest.load_golden(9).blob)
for pool in (1 << 19, 1 << 20, 1 << 21, 1 << 22):
    for _ in range(2):
        acc, st = ctx.render(ctx.params(800, 800, 100, 1, seed=3, pool_paths=pool))
    print(f"C2 pool {pool}: {st['device_ms']:.1f} ms {st['paths']/st['device_ms']/1e3:.1f} Mpaths/s iters {st['iterations']}")
