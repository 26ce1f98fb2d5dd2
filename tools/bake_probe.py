"""Upper bound of 'sphere-only instances baked into world space' on scene09: the same scene with the 1,000 small
spheres written as world-space spheres (no rotate_y + translate chain) against the stock scene."""
import importlib, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("ray_tracing-rendering_b200"); S = importlib.import_module("ray_tracing-rendering_b200.scenes")
b = importlib.import_module("ray_tracing-rendering_b200.binding")

def baked(seed=1):
    rng = S.XorShift32(seed)
    B = S.SceneBuilder()
    ground = B.lambertian((0.48, 0.83, 0.53))
    for i in range(20):
        for j in range(20):
            w = 100.0
            x0, z0 = -1000.0 + i * w, -1000.0 + j * w
            y1 = rng.uniform(1, 101)
            B.box((x0, 0.0, z0), (x0 + w, y1, z0 + w), ground)
    B.xz_rect(123, 423, 147, 412, 554, B.diffuse_light((7, 7, 7)))
    B.moving_sphere((400, 400, 200), (430, 400, 200), 0, 1, 50, B.lambertian((0.7, 0.3, 0.1)))
    B.sphere((260, 150, 45), 50, B.dielectric(1.5))
    B.sphere((0, 150, 145), 50, B.metal((0.8, 0.8, 0.9), 1.0))
    glass = B.dielectric(1.5)
    B.sphere((360, 150, 145), 70, glass)
    inner = B.sphere((360, 150, 145), 70, glass)
    B.prims[inner]["flags"] = 1
    B.constant_medium(inner, 0.2, (0.2, 0.4, 0.9))
    fog = B.sphere((0, 0, 0), 5000, B.dielectric(1.5))
    B.prims[fog]["flags"] = 1
    B.constant_medium(fog, 0.0001, (1, 1, 1))
    B.sphere((400, 200, 400), 100, B.lambertian_tex(B.image_missing()))
    B.sphere((220, 280, 300), 80, B.lambertian_tex(B.noise(0.1, rng)))
    white = B.lambertian((.73, .73, .73))
    c, s = math.cos(math.radians(15)), math.sin(math.radians(15))
    for _ in range(1000):
        x, y, z = rng.uniform(0, 165), rng.uniform(0, 165), rng.uniform(0, 165)
        B.sphere((c * x + s * z - 100, y + 270, -s * x + c * z + 395), 10, white)
    return B.finish(9, 800, 1.0, 500, (0, 0, 0), (478, 278, -600), (278, 278, 0), 40.0)

ctx = pkg.Context(0)
for name, blob in (("stock", S.final_scene(1)), ("baked", baked(1))):
    ctx.upload_scene(blob)
    ctx.render(ctx.params(800, 800, 2, 1, 50))
    for r in range(2):
        _, st = ctx.render(ctx.params(800, 800, 100, 1, 50, seed=3 + r, flags=b.RENDER_TIME_EXTEND))
        print(name, f"{st['device_ms']:.2f} ms stage_ms {[round(x, 2) for x in st['stage_ms']]} traversal {st['traversal']} rays {st['rays_closest'] / 1e6:.1f}M", flush=True)
    ctx.set_option(b.OPT_BINARY_TRAVERSAL, 2)
    ctx.upload_scene(blob)
    _, st = ctx.render(ctx.params(800, 800, 100, 1, 50, seed=3, flags=b.RENDER_TIME_EXTEND))
    print(name, "4-wide", f"{st['device_ms']:.2f} ms stage_ms {[round(x, 2) for x in st['stage_ms']]} traversal {st['traversal']}", flush=True)
    ctx.set_option(b.OPT_BINARY_TRAVERSAL, 0)
