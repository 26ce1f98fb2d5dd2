import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("ray_tracing-rendering_b200")
scenes = importlib.import_module("ray_tracing-rendering_b200.scenes")
ctx = pkg.Context(0)
sid = int(sys.argv[1]) if len(sys.argv) > 1 else 7
integ = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx.upload_scene(scenes.select_scene(sid))
T = pkg.abi.parse_blob(scenes.select_scene(sid))["globals"][0]
w, h, spp = int(T["image_width"]), int(T["image_height"]), int(T["samples_per_pixel"])
for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 3):
    acc, st = ctx.render(ctx.params(w, h, spp, integ, seed=3))
    print(f"scene{sid} int{integ} {w}x{h}x{spp}: {st['device_ms']:.1f} ms {st['paths']/st['device_ms']/1e3:.1f} Mpaths/s {(st['rays_closest']+st['rays_shadow'])/st['device_ms']/1e3:.1f} Mrays/s sched {st['schedule']}")
