"""Small renders through every kernel family, for compute-sanitizer (memcheck / racecheck)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conftest
pkg = importlib.import_module("ray_tracing-rendering_b200"); b = importlib.import_module("ray_tracing-rendering_b200.binding")
scenes = importlib.import_module("ray_tracing-rendering_b200.scenes")
ctx = pkg.Context(0)
jobs = [(conftest.load_golden(9).blob, 48, 48, 4, 1, 0, 0), (conftest.load_golden(1).blob, 40, 30, 4, 4, 0, 1000),
        (conftest.load_golden(23).blob, 40, 30, 4, 4, 0, 0), (conftest.load_golden(23).blob, 40, 30, 4, 3, b.RENDER_FORCE_FUSED, 0),
        (conftest.load_golden(7).blob, 40, 40, 4, 1, 0, 0), (conftest.load_golden(21).blob, 40, 40, 4, 4, b.RENDER_FORCE_WAVEFRONT, 0),
        (scenes.hdr_demo(64, scenes.synthetic_hdr(64, 32, 1)), 64, 36, 4, 4, 0, 0),
        (scenes.sphere_field(20, 64, 36, 4), 64, 36, 4, 4, b.RENDER_COUNT_VISITS, 4096)]
for blob, w, h, spp, integ, flags, pool in jobs:
    ctx.upload_scene(blob)
    acc, st = ctx.render(ctx.params(w, h, spp, integ, seed=2, flags=flags, pool_paths=pool))
    assert st["paths"] == w * h * spp and np.isfinite(acc).all()
    rgb = ctx.resolve_rgb8(w, h, spp)
    print("ok", w, h, spp, integ, flags, st["schedule"], st["rays_closest"], st["rays_shadow"], flush=True)
g = conftest.load_golden(9)
ctx.upload_scene(g.blob)
ctx.trace(g["rays"][:512], 64); ctx.trace(g["rays"][:512], 32); ctx.trace(g["rays"][:512], 34); ctx.trace(g["rays"][:512], 36)
ctx.set_option(b.OPT_BINARY_TRAVERSAL, 2)          # the 4-wide kernels on the scene with media and an instance
ctx.upload_scene(g.blob)
acc, st = ctx.render(ctx.params(40, 40, 4, 1, seed=5))
assert st["traversal"] == 2 and np.isfinite(acc).all()
ctx.set_option(b.OPT_BINARY_TRAVERSAL, 0)
# round 2's additions: the reference-order walk with seeded draws, gated spheres (exact and in-tree), image textures,
# the device-built environment tables, the lazy fp64 tables
z = np.load(os.path.join(ROOT, "tests", "golden", "media09.npz"))
ctx.upload_scene(z["blob"].tobytes()); ctx.trace(z["rays"][:256], 65)
for sid, integ in ((34, 4), (35, 4), (36, 4), (22, 4)):
    gg = conftest.load_golden(sid)
    ctx.set_option(b.OPT_LAZY_F64_PRIMS, 0 if sid == 22 else 100000)
    ctx.upload_scene(gg.blob)
    acc, st = ctx.render(ctx.params(48, 27, 4, integ, seed=3))
    assert np.isfinite(acc).all()
    ctx.trace(gg["rays"][:256], 64); ctx.trace(gg["rays"][:256], 34)
    if sid == 36:
        assert len(ctx.env_tables()) > 0
    print("ok scene", sid, st["schedule"], st["traversal"], flush=True)
print("done")
