import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ray_tracing-rendering_b200")
scenes = importlib.import_module("ray_tracing-rendering_b200.scenes")
ctx = pkg.Context(0)
ctx.upload_scene(scenes.sphere_field(500, 3840, 2160, 1024))
acc, st = ctx.render(ctx.params(3840, 2160, 2, 4, seed=3))
print(f"C5 2spp: {st['device_ms']:.1f} ms {st['paths']/st['device_ms']/1e3:.1f} Mpaths/s iters {st['iterations']} launches {st['kernel_launches']}")
