import sys, time, importlib
sys.path.insert(0, '/root/repo')
pkg = importlib.import_module("ray_tracing-rendering_b200"); cfgs = importlib.import_module("ray_tracing-rendering_b200.configs")
ctx = pkg.Context(0)
for name in ("C4env", "C5"):
    blob = cfgs.get(name).blob()
    for i in range(3):
        t = time.time(); ctx.upload_scene(blob); print(name, "upload wall ms", round(1e3 * (time.time() - t), 1), flush=True)
