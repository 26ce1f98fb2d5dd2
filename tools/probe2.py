import importlib, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
np.set_printoptions(linewidth=200, precision=5, suppress=True)
from oracle import refbind
binding = importlib.import_module("ray_tracing-rendering_b200.binding")
abi = refbind.abi
ctx = binding.Context(0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "s9"):
    s = refbind.RefScene(9); ctx.upload_scene(s.blob())
    W = H = 200; spp = 64
    for md in (1, 2, 3, 4, 6, 10, 50):
        S, S2, cnt = s.render_linear(1, W, H, spp, max_depth=md)
        acc, st = ctx.render(ctx.params(W, H, spp, 1, max_depth=md, seed=5))
        print(f"s9 depth{md}: ref mean {(S/spp).mean(axis=(0,1))} rays/path {cnt.sum()/(W*H*spp):.4f} | gpu {acc[...,:3].mean(axis=(0,1))/spp} rays/path {st['rays_closest']/st['paths']:.4f}")
if which in ("all", "s23"):
    for integ in (2, 3, 4):
        s = refbind.RefScene(23); ctx.upload_scene(s.blob())
        W, H, spp = 400, 225, 128
        S, S2, cnt = s.render_linear(integ, W, H, spp)
        acc, st = ctx.render(ctx.params(W, H, spp, integ, seed=5))
        print(f"s23 int{integ}: ref {(S/spp).mean(axis=(0,1))} rays/path {cnt/(W*H*spp)} | gpu {acc[...,:3].mean(axis=(0,1))/spp} rays/path {st['rays_closest']/st['paths']:.4f} {st['rays_shadow']/st['paths']:.4f} nan {np.isnan(acc).sum()}")
    for sid, integ in ((24, 4), (24, 3), (11, 2), (12, 4), (5, 1), (8, 1), (22, 3)):
        s = refbind.RefScene(sid); ctx.upload_scene(s.blob())
        g = abi.parse_blob(s.blob())['globals'][0]
        W, H, spp = int(g['image_width']) // 4, int(g['image_height']) // 4, 128
        S, S2, cnt = s.render_linear(integ, W, H, spp)
        acc, st = ctx.render(ctx.params(W, H, spp, integ, seed=5))
        print(f"s{sid} int{integ}: ref {(S/spp).mean(axis=(0,1))} rays/path {cnt/(W*H*spp)} | gpu {acc[...,:3].mean(axis=(0,1))/spp} rays/path {st['rays_closest']/st['paths']:.4f} {st['rays_shadow']/st['paths']:.4f} nan {np.isnan(acc).sum()}")
if which in ("all", "c1"):
    s = refbind.RefScene(7); ctx.upload_scene(s.blob())
    p = ctx.params(600, 600, 400, 1, seed=3, pool_paths=1 << 20)
    for _ in range(3):
        acc, st = ctx.render(p)
        print(f"C1: {st['device_ms']:.1f} ms {st['paths']/st['device_ms']/1e3:.1f} Mpaths/s iters {st['iterations']}")
