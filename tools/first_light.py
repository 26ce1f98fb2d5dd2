"""Development probe: first GPU run of the library against the reference oracle."""
import importlib, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
np.set_printoptions(linewidth=200, precision=6, suppress=True)
from oracle import refbind
pkg = importlib.import_module("ray_tracing-rendering_b200")
binding = importlib.import_module("ray_tracing-rendering_b200.binding")
abi = refbind.abi

ctx = binding.Context(0)
print(binding.load().rtb_version())
for sid, integ in [(7, 1), (21, 3), (21, 4), (23, 4), (23, 3), (1, 1), (9, 1)]:
    s = refbind.RefScene(sid)
    blob = s.blob()
    T = abi.parse_blob(blob)
    ctx.upload_scene(blob)
    print("scene", sid, ctx.scene_stats())
    rays, hits, c = s.record_rays(integ, 20000, 300000)
    is_med = T['prims']['type'] == 5
    def med(p): return np.where(p >= 0, is_med[np.maximum(p, 0)], False)
    for prec in (64, 32):
        out, vis = ctx.trace(rays, prec, want_visits=True)
        ok = ~med(hits['prim']) & ~med(out['prim'])
        closest = ~np.isfinite(rays['t_max'])
        pm = (out['prim'] == hits['prim'])
        tm = (out['t'] == hits['t'])
        print(f"  prec{prec}: n={len(rays)} prim_match={pm[ok].mean():.6f} (closest {pm[ok&closest].mean():.6f}) t_bitexact={tm[ok].mean():.6f} nodes/ray={vis[0]/len(rays):.2f} tests/ray={vis[1]/len(rays):.2f}")
    g = T['globals'][0]
    W, H = int(g['image_width']) // 2, int(g['image_height']) // 2
    spp = 64
    t0 = time.time(); S, S2, cnt = s.render_linear(integ, W, H, spp); tref = time.time() - t0
    ref_mean = S / spp
    p = ctx.params(W, H, spp, integ, seed=7)
    acc, st = ctx.render(p)
    gpu_mean = acc[..., :3] / spp
    print(f"  render {W}x{H}x{spp} int{integ}: ref mean {ref_mean.mean(axis=(0,1))} gpu mean {gpu_mean.mean(axis=(0,1))}")
    print(f"     ref {tref:.2f}s rays/path {cnt.sum()/(W*H*spp):.3f} | gpu {st['device_ms']:.2f} ms rays/path {(st['rays_closest']+st['rays_shadow'])/st['paths']:.3f} iters {st['iterations']} launches {st['kernel_launches']}")
    var = np.maximum(S2 / spp - ref_mean**2, 0) / spp
    z = (gpu_mean - ref_mean) / np.sqrt(2 * var + 1e-12)
    print(f"     |z|<3 frac {np.mean(np.abs(z) < 3):.4f}  rmse {np.sqrt(np.mean((gpu_mean-ref_mean)**2)):.5f}")

# full config C1
s = refbind.RefScene(7); ctx.upload_scene(s.blob())
for pool in (1 << 19, 1 << 20, 1 << 21, 1 << 22):
    p = ctx.params(600, 600, 400, 1, seed=3, pool_paths=pool)
    acc, st = ctx.render(p)
    acc, st = ctx.render(p)
    paths = st['paths']
    print(f"C1 pool={pool}: {st['device_ms']:.1f} ms  {paths/st['device_ms']/1e3:.1f} Mpaths/s  {(st['rays_closest']+st['rays_shadow'])/st['device_ms']/1e3:.1f} Mrays/s iters {st['iterations']} launches {st['kernel_launches']} mean {acc[...,:3].mean(axis=(0,1))/400}")
