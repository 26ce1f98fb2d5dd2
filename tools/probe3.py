import importlib, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
np.set_printoptions(linewidth=200, precision=5, suppress=True)
from oracle import refbind
binding = importlib.import_module("ray_tracing-rendering_b200.binding")
abi = refbind.abi
ctx = binding.Context(0)
s = refbind.RefScene(9); blob = s.blob(); T = abi.parse_blob(blob); ctx.upload_scene(blob)
ptype = T['prims']['type']; pmat = T['prims']['material']; mtype = T['materials']['type']
rays, hits, c = s.record_rays(1, 60000, 400000)
for prec in (64, 32):
    out = ctx.trace(rays, prec)
    for name, h in (("ref", hits), (f"gpu{prec}", out)):
        miss = h['prim'] < 0
        pt = np.where(h['prim'] >= 0, ptype[np.maximum(h['prim'], 0)], -1)
        mt = np.where(h['prim'] >= 0, mtype[pmat[np.maximum(h['prim'], 0)]], -1)
        print(name, "miss", miss.mean(), "prim types", [(k, round((pt == k).mean(), 4)) for k in range(6)], "mat types", [(k, round((mt == k).mean(), 4)) for k in range(6)])
    # by origin type
    has_o = rays['origin_prim'] >= 0
    print("  rays with origin", has_o.mean())
    for k in range(6):
        sel = has_o & (ptype[np.maximum(rays['origin_prim'], 0)] == k)
        if sel.sum():
            print(f"   origin primtype {k}: n={sel.sum()} ref miss {(hits['prim'][sel]<0).mean():.4f} gpu miss {(out['prim'][sel]<0).mean():.4f}  ref med {(np.where(hits['prim'][sel]>=0, ptype[np.maximum(hits['prim'][sel],0)],-1)==5).mean():.4f} gpu med {(np.where(out['prim'][sel]>=0, ptype[np.maximum(out['prim'][sel],0)],-1)==5).mean():.4f}")
