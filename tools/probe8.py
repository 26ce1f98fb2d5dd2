import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
pkg = importlib.import_module("ray_tracing-rendering_b200")
scenes = importlib.import_module("ray_tracing-rendering_b200.scenes")
binding = importlib.import_module("ray_tracing-rendering_b200.binding")
import conftest
ctx = pkg.Context(0)
for name, blob, w, h, spp, integ in (("C5", scenes.sphere_field(500, 3840, 2160, 1024), 3840, 2160, 16, 4), ("C2", conftest.load_golden(9).blob, 800, 800, 100, 1)):
    ctx.upload_scene(blob)
    ctx.render(ctx.params(w, h, 2, integ))
    acc, a = ctx.render(ctx.params(w, h, spp, integ, seed=3))
    acc, st = ctx.render(ctx.params(w, h, spp, integ, seed=3, flags=binding.RENDER_TIME_EXTEND))
    print(f"{name}: untimed {a['device_ms']:.1f} ms {a['paths']/a['device_ms']/1e3:.0f} Mpaths/s | timed {st['device_ms']:.1f} ms iters {st['iterations']} stage_ms extend/shade/miss/connect {[round(x,1) for x in st['stage_ms']]} sum {sum(st['stage_ms']):.1f} rays {st['rays_closest']/1e6:.0f}M+{st['rays_shadow']/1e6:.0f}M")
