"""Times every catalogue fixture the fused schedule takes (scenes of <= 64 primitive records) at 400x400, 64 spp with
its own integrator: which k_fused instantiation a scene gets depends on its materials and integrator."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import conftest
pkg = importlib.import_module("ray_tracing-rendering_b200")
ctx = pkg.Context(0)
tot = {}
for sid in conftest.ALL_SCENES:
    g = conftest.load_golden(sid)
    ctx.upload_scene(g.blob)
    for integ in (1, 4):
        ctx.render(ctx.params(400, 400, 2, integ, 50))
        best = 1e9
        for r in range(3):
            _, st = ctx.render(ctx.params(400, 400, 64, integ, 50, seed=r + 1))
            best = min(best, st["device_ms"])
        if st["schedule"] == 1:
            print(f"scene {sid:2d} integrator {integ}: {best:7.2f} ms", flush=True)
            tot[integ] = tot.get(integ, 0.0) + best
print("total ms per integrator", {k: round(v, 1) for k, v in tot.items()})
