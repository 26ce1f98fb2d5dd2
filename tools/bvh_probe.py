"""Extend-stage time against the BVH builder knobs (largest leaf, SAH traversal cost, node order)."""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ray_tracing-rendering_b200"); cf = importlib.import_module("ray_tracing-rendering_b200.configs"); b = importlib.import_module("ray_tracing-rendering_b200.binding")
ctx = pkg.Context(0)
knobs = [(4, 100, 0), (4, 100, 1), (2, 100, 1), (4, 200, 1), (8, 200, 0)] if len(sys.argv) < 2 else [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for name, spp in (("C2", 50), ("C5", 8)):
    c = cf.get(name); blob = c.blob()
    for max_leaf, cost, dfs in knobs:
        ctx.set_option(3, max_leaf); ctx.set_option(4, cost); ctx.set_option(5, dfs)
        ctx.upload_scene(blob)
        ctx.render(ctx.params(c.width, c.height, 2, c.integrator))
        _, st = ctx.render(ctx.params(c.width, c.height, spp, c.integrator, seed=3, flags=b.RENDER_TIME_EXTEND))
        _, sc = ctx.render(ctx.params(c.width, c.height, 2, c.integrator, seed=3, flags=b.RENDER_COUNT_VISITS))
        rays = sc["rays_closest"] + sc["rays_shadow"]
        print(f"{name} max_leaf {max_leaf:2d} cost {cost:4d}% dfs {dfs}: nodes {ctx.scene_stats()['n_nodes']:8d} total {st['device_ms']:7.2f} ms stages {[round(x, 2) for x in st['stage_ms']]} "
              f"nodes/ray {sc['nodes_visited'] / rays:.2f} prims/ray {sc['prim_tests'] / rays:.2f}", flush=True)
