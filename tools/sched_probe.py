"""Fused vs wavefront schedule on the small (shared-memory sized) configurations, per integrator:
the measurements behind the schedule rule in rtb_wavefront.cu."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ray_tracing-rendering_b200"); cf = importlib.import_module("ray_tracing-rendering_b200.configs"); b = importlib.import_module("ray_tracing-rendering_b200.binding")
ctx = pkg.Context(0)
for name in ("C1", "C3", "C4", "C4env"):
    c = cf.get(name); ctx.upload_scene(c.blob())
    for integ in (1, 2, 3, 4):
        out = []
        for fl in (b.RENDER_FORCE_FUSED, b.RENDER_FORCE_WAVEFRONT, 0):
            best = 1e9
            for rep in range(3):
                _, st = ctx.render(ctx.params(c.width, c.height, min(c.spp, 64), integ, flags=fl, seed=rep + 1))
                best = min(best, st["device_ms"])
            out.append((st["schedule"], round(best, 2)))
        print(name, "integrator", integ, "fused/wavefront/default (schedule, ms)", out, flush=True)
