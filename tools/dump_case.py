import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import numpy as np
import conftest
binding = importlib.import_module("ray_tracing-rendering_b200.binding")
sid, integ = int(sys.argv[1]), int(sys.argv[2])
g = conftest.load_golden(sid)
ctx = binding.Context(0); ctx.upload_scene(g.blob)
ref_spp = int(g[f"img_{integ}_spp"][0]); h, w, _ = g[f"img_{integ}_sum"].shape
means = [ctx.render(ctx.params(w, h, max(ref_spp // 2, 64), integ, seed=100 + i))[0][..., :3] / max(ref_spp // 2, 64) for i in range(8)]
np.save(f"gpurun_out/dump_{sid}_{integ}.npy", np.stack(means))
