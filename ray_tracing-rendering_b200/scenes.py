"""Scene builders for the BASELINE.json configurations, written against the flat-scene
tables directly (numpy).  They restate WHAT the reference's builders construct
(src/scene/scenes.cpp, cited per function) in the blob format of include/rtb200_scene.h;
tests/test_scenes.py checks them primitive by primitive against the blobs obtained by
walking the reference's own graphs.  The C++ mirror of the builder API lives in host/.
"""
import math

import numpy as np

from . import abi

T_LAMBERTIAN, T_METAL, T_DIELECTRIC, T_LIGHT, T_PBR, T_ISOTROPIC = range(6)
P_SPHERE, P_MSPHERE, P_XY, P_XZ, P_YZ, P_MEDIUM = range(6)
XF_TRANSLATE, XF_ROTATE_Y, XF_FLIP = range(3)


class SceneBuilder:
    """Accumulates the tables of one scene; finish() returns the blob."""

    def __init__(self):
        self.prims, self.chains, self.ops, self.mats, self.texs, self.lights = [], [], [], [], [], []
        self.images, self.perlins = [], []
        self.env_texels = np.zeros(0, np.float32)
        self.image_bytes = np.zeros(0, np.uint8)

    # ---- textures / materials (src/materials) ------------------------------------------------
    def solid(self, c):
        t = np.zeros(1, abi.TEXTURE)[0]
        t["type"] = 0
        t["even"] = t["odd"] = t["image"] = t["perlin"] = -1
        t["color"] = c
        self.texs.append(t)
        return len(self.texs) - 1

    def _mat(self, mtype, tex=(-1, -1, -1, -1), color=(0, 0, 0), fuzz=0.0, ir=0.0):
        m = np.zeros(1, abi.MATERIAL)[0]
        m["type"] = mtype
        m["tex"] = tex
        m["color"] = color
        m["fuzz"] = fuzz
        m["ir"] = ir
        self.mats.append(m)
        return len(self.mats) - 1

    def lambertian(self, c):                       # material.h:74-75
        return self._mat(T_LAMBERTIAN, (self.solid(c), -1, -1, -1))

    def metal(self, c, fuzz):                      # material.h:120-121 (fuzz clamped to 1)
        return self._mat(T_METAL, color=c, fuzz=fuzz if fuzz < 1 else 1.0)

    def dielectric(self, ir):                      # material.h:149-150
        return self._mat(T_DIELECTRIC, ir=ir)

    def diffuse_light(self, c):                    # material.h:208-211
        return self._mat(T_LIGHT, (self.solid(c), -1, -1, -1))

    def pbr(self, albedo, roughness, metallic):    # material.h:240-243 with solid_color textures
        return self._mat(T_PBR, (self.solid(albedo), self.solid((roughness,) * 3),
                                 self.solid((metallic,) * 3), -1))

    # ---- wrappers (src/geometry/hittable.h) --------------------------------------------------
    def chain(self, ops):
        """ops: list of ('translate', (x,y,z)) | ('rotate_y', degrees) | ('flip',), outermost first."""
        first = len(self.ops)
        for op in ops:
            r = np.zeros(1, abi.XFORM_OP)[0]
            if op[0] == "translate":
                r["kind"], (r["a"], r["b"], r["c"]) = XF_TRANSLATE, op[1]
            elif op[0] == "rotate_y":                # hittable.h:96-99
                rad = op[1] * 3.1415926535897932385 / 180.0
                r["kind"], r["a"], r["b"] = XF_ROTATE_Y, math.sin(rad), math.cos(rad)
            else:
                r["kind"] = XF_FLIP
            self.ops.append(r)
        c = np.zeros(1, abi.CHAIN)[0]
        c["first"], c["count"] = first, len(ops)
        self.chains.append(c)
        return len(self.chains) - 1

    # ---- primitives (src/geometry) -----------------------------------------------------------
    def _prim(self, ptype, mat, d, chain=-1, flags=0):
        p = np.zeros(1, abi.PRIM)[0]
        p["type"], p["material"], p["chain"], p["flags"] = ptype, mat, chain, flags
        p["d"][:len(d)] = d
        self.prims.append(p)
        return len(self.prims) - 1

    def sphere(self, c, r, mat, chain=-1):
        return self._prim(P_SPHERE, mat, [c[0], c[1], c[2], r], chain)

    def xy_rect(self, x0, x1, y0, y1, k, mat, chain=-1):
        return self._prim(P_XY, mat, [x0, x1, y0, y1, k], chain)

    def xz_rect(self, x0, x1, z0, z1, k, mat, chain=-1):
        return self._prim(P_XZ, mat, [x0, x1, z0, z1, k], chain)

    def yz_rect(self, y0, y1, z0, z1, k, mat, chain=-1):
        return self._prim(P_YZ, mat, [y0, y1, z0, z1, k], chain)

    def moving_sphere(self, c0, c1, t0, t1, r, mat, chain=-1):     # moving_sphere.h:10-17
        return self._prim(P_MSPHERE, mat, [c0[0], c0[1], c0[2], c1[0], c1[1], c1[2], t0, t1, r], chain)

    def constant_medium(self, boundary_prim, density, c):          # constant_medium.h:28-53
        """The boundary primitive must already exist; pass flags=BOUNDARY_ONLY when it is not
        itself a member of the world."""
        m = self._mat(T_ISOTROPIC, (self.solid(c), -1, -1, -1))
        i = self._prim(P_MEDIUM, m, [-1.0 / density])
        self.prims[i]["aux0"], self.prims[i]["aux1"] = boundary_prim, 1
        return i

    def image_missing(self):                                       # texture.h:96-110 (file not found)
        im = np.zeros(1, abi.IMAGE)[0]
        self.images.append(im)
        t = np.zeros(1, abi.TEXTURE)[0]
        t["type"] = 2
        t["even"] = t["odd"] = t["perlin"] = -1
        t["image"] = len(self.images) - 1
        self.texs.append(t)
        return len(self.texs) - 1

    def noise(self, scale, rng):                                   # texture.h:148-163, perlin.h:10-20,62-78
        pl = np.zeros(1, abi.PERLIN)[0]
        for i in range(256):
            v = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-1, 1)])
            pl["ranvec"][i] = (1.0 / math.sqrt(float(v @ v))) * v
        for name in ("perm_x", "perm_y", "perm_z"):
            perm = list(range(256))
            for i in range(255, 0, -1):
                tgt = rng.randint(0, i)
                perm[i], perm[tgt] = perm[tgt], perm[i]
            pl[name] = perm
        self.perlins.append(pl)
        t = np.zeros(1, abi.TEXTURE)[0]
        t["type"] = 3
        t["even"] = t["odd"] = t["image"] = -1
        t["perlin"], t["scale"] = len(self.perlins) - 1, scale
        self.texs.append(t)
        return len(self.texs) - 1

    def lambertian_tex(self, tex):
        return self._mat(T_LAMBERTIAN, (tex, -1, -1, -1))

    def box(self, p0, p1, mat, chain=-1):            # box.h:31-47: six rects, this order
        self.xy_rect(p0[0], p1[0], p0[1], p1[1], p1[2], mat, chain)
        self.xy_rect(p0[0], p1[0], p0[1], p1[1], p0[2], mat, chain)
        self.xz_rect(p0[0], p1[0], p0[2], p1[2], p1[1], mat, chain)
        self.xz_rect(p0[0], p1[0], p0[2], p1[2], p0[1], mat, chain)
        self.yz_rect(p0[1], p1[1], p0[2], p1[2], p1[0], mat, chain)
        self.yz_rect(p0[1], p1[1], p0[2], p1[2], p0[0], mat, chain)

    # ---- lights (src/lighting) ---------------------------------------------------------------
    def quad_light(self, Q, u, v, c):
        l = np.zeros(1, abi.LIGHT)[0]
        l["type"], l["Q"], l["u"], l["v"], l["intensity"] = 0, Q, u, v, c
        self.lights.append(l)

    def env_light(self, texels=None):
        """texels: (H, W, 3) float32 or None (= file missing, environmental_light.h:127-132)."""
        l = np.zeros(1, abi.LIGHT)[0]
        l["type"] = 4
        if texels is not None:
            h, w, _ = texels.shape
            l["env_width"], l["env_height"], l["env_is_probe"] = w, h, int(w == h)
            l["env_offset"] = self.env_texels.size
            self.env_texels = np.concatenate([self.env_texels, np.asarray(texels, np.float32).ravel()])
        self.lights.append(l)

    # ---- SceneConfig (scenes.h:11-24) + camera (main.cpp:63-70) -------------------------------
    def finish(self, scene_id, width, aspect, spp, background, lookfrom, lookat, vfov, aperture=0.0,
               focus_dist=10.0, vup=(0, 1, 0)):
        g = np.zeros(1, abi.GLOBALS)
        g[0]["background"] = background
        g[0]["image_width"] = width
        g[0]["image_height"] = int(width / aspect)
        g[0]["samples_per_pixel"] = spp
        g[0]["scene_id"] = scene_id
        c = np.zeros(1, abi.CAMERA)
        c[0]["lookfrom"], c[0]["lookat"], c[0]["vup"] = lookfrom, lookat, vup
        c[0]["vfov"], c[0]["aspect_ratio"], c[0]["aperture"], c[0]["focus_dist"] = vfov, aspect, aperture, focus_dist
        c[0]["time0"], c[0]["time1"] = 0.0, 1.0          # RenderConfig::kShutterOpen/Close, main.cpp:45-46

        def arr(lst, dt):
            return np.array(lst, dtype=dt) if lst else np.zeros(0, dt)
        return abi.build_blob({
            "globals": g, "camera": c, "prims": arr(self.prims, abi.PRIM), "chains": arr(self.chains, abi.CHAIN),
            "xform_ops": arr(self.ops, abi.XFORM_OP), "materials": arr(self.mats, abi.MATERIAL),
            "textures": arr(self.texs, abi.TEXTURE), "images": arr(self.images, abi.IMAGE),
            "image_bytes": self.image_bytes, "perlins": arr(self.perlins, abi.PERLIN),
            "lights": arr(self.lights, abi.LIGHT), "env_texels": self.env_texels})


def cornell_box(nee: bool = False) -> bytes:
    """scene 7 (cornell_box, scenes.cpp:159-187 + case 7, :1572-1582) or, with nee=True,
    scene 21 (cornell_box_nee, scenes.cpp:779-809 + case 21, :1729-1744)."""
    b = SceneBuilder()
    red = b.lambertian((.65, .05, .05))
    white = b.lambertian((.73, .73, .73))
    green = b.lambertian((.12, .45, .15))
    light = b.diffuse_light((15, 15, 15))
    b.yz_rect(0, 555, 0, 555, 555, green)
    b.yz_rect(0, 555, 0, 555, 0, red)
    b.xz_rect(213, 343, 227, 332, 554, light, b.chain([("flip",)]) if nee else -1)
    b.xz_rect(0, 555, 0, 555, 0, white)
    b.xz_rect(0, 555, 0, 555, 555, white)
    b.xy_rect(0, 555, 0, 555, 555, white)
    b.box((0, 0, 0), (165, 330, 165), white, b.chain([("translate", (265, 0, 295)), ("rotate_y", 15)]))
    b.box((0, 0, 0), (165, 165, 165), white, b.chain([("translate", (130, 0, 65)), ("rotate_y", -18)]))
    if nee:
        b.quad_light((213, 554, 227), (130, 0, 0), (0, 0, 105), (15, 15, 15))
    return b.finish(21 if nee else 7, 600, 1.0, 400, (0, 0, 0), (278, 278, -800), (278, 278, 0), 40.0)


def mis_comparison() -> bytes:
    """scene 23 (mis_comparison_scene, scenes.cpp:580-626 + case 23, :1762-1781)."""
    b = SceneBuilder()
    b.sphere((0, -1000, 0), 1000, b.lambertian((0.5, 0.5, 0.5)))
    b.sphere((-2.5, 1, 0), 1.0, b.pbr((0.9, 0.6, 0.2), 0.001, 1.0))
    b.sphere((0, 1, 0), 1.0, b.pbr((0.8, 0.8, 0.8), 0.4, 1.0))
    b.sphere((2.5, 1, 0), 1.0, b.dielectric(1.5))
    b.xz_rect(-10, 10, -10, 10, 10, b.diffuse_light((5, 5, 5)), b.chain([("flip",)]))
    b.yz_rect(3.75, 4.25, 1.75, 2.25, 6, b.diffuse_light((50, 50, 50)), b.chain([("flip",)]))
    b.quad_light((-10, 10, -10), (20, 0, 0), (0, 0, 20), (5, 5, 5))
    b.quad_light((6, 4, 2), (0, 0.5, 0), (0, 0, 0.5), (50, 50, 50))      # faces +X: reference quirk 4
    return b.finish(23, 800, 16.0 / 9.0, 64, (0, 0, 0), (0, 3, 8), (0, 1, 0), 35.0)


class XorShift32:
    """The generator of the reference (rtweekend.h:24-34) with an explicit seed, used only to
    make the synthetic C5 scene reproducible."""

    def __init__(self, seed=1):
        self.s = np.uint32(seed)

    def next(self):
        s = int(self.s)
        s ^= (s << 13) & 0xFFFFFFFF
        s ^= s >> 17
        s ^= (s << 5) & 0xFFFFFFFF
        self.s = np.uint32(s)
        return s * 2.3283064365386963e-10

    def uniform(self, lo, hi):                       # random_double(min, max), rtweekend.h:37-40
        return lo + (hi - lo) * self.next()

    def randint(self, lo, hi):                       # random_int, rtweekend.h:48-50
        return int(self.uniform(lo, hi + 1))

    def block(self, n):
        out = np.empty(n)
        s = int(self.s)
        for i in range(n):
            s ^= (s << 13) & 0xFFFFFFFF
            s ^= s >> 17
            s ^= (s << 5) & 0xFFFFFFFF
            out[i] = s * 2.3283064365386963e-10
        self.s = np.uint32(s)
        return out


def sphere_field(half_extent: int = 500, width: int = 3840, height: int = 2160, spp: int = 1024, seed: int = 1) -> bytes:
    """C5: the static generalisation of random_scene (scenes.cpp:15-59): a ground
    sphere plus one r=0.2 sphere per unit cell of [-half_extent, half_extent)^2 (1,000,000 spheres
    for half_extent 500), materials by xi (<0.8 lambertian, <0.95 metal, else glass), every sphere
    its own material record as in the reference, one 200x200 quad light at y=100 with its
    flip_face rect, sky background.  Vectorised construction; numbers come from numpy's PCG64
    seeded with `seed` (the scene is synthetic — there is no reference instance to reproduce)."""
    n_side = 2 * half_extent
    n = n_side * n_side
    rng = np.random.default_rng(seed)
    a, bb = np.meshgrid(np.arange(-half_extent, half_extent), np.arange(-half_extent, half_extent), indexing="ij")
    choose = rng.random(n)
    cx = a.ravel() + 0.9 * rng.random(n)
    cz = bb.ravel() + 0.9 * rng.random(n)
    is_lam, is_met = choose < 0.8, (choose >= 0.8) & (choose < 0.95)
    n_tex = int(is_lam.sum()) + 2
    texs = np.zeros(n_tex, abi.TEXTURE)
    texs["type"] = 0
    for f in ("even", "odd", "image", "perlin"):
        texs[f] = -1
    texs["color"][0] = (0.5, 0.5, 0.5)      # ground
    texs["color"][1] = (15, 15, 15)         # emitter
    texs["color"][2:] = rng.random((n_tex - 2, 3)) * rng.random((n_tex - 2, 3))
    mats = np.zeros(n + 2, abi.MATERIAL)
    mats["tex"] = -1
    mats["type"][0], mats["tex"][0, 0] = T_LAMBERTIAN, 0
    mats["type"][1], mats["tex"][1, 0] = T_LIGHT, 1
    m = mats[2:]
    m["type"] = np.where(is_lam, T_LAMBERTIAN, np.where(is_met, T_METAL, T_DIELECTRIC))
    tex0 = np.full(n, -1, np.int32)
    tex0[is_lam] = 2 + np.arange(int(is_lam.sum()))
    m["tex"][:, 0] = tex0
    m["color"][is_met] = 0.5 + 0.5 * rng.random((int(is_met.sum()), 3))
    m["fuzz"][is_met] = 0.5 * rng.random(int(is_met.sum()))
    m["ir"][~is_lam & ~is_met] = 1.5
    prims = np.zeros(n + 2, abi.PRIM)
    prims["chain"] = -1
    prims["type"][0], prims["material"][0], prims["d"][0, :4] = P_SPHERE, 0, (0, -1000, 0, 1000)
    prims["type"][1], prims["material"][1], prims["chain"][1] = P_XZ, 1, 0
    prims["d"][1, :5] = (-100, 100, -100, 100, 100)
    p = prims[2:]
    p["type"], p["material"] = P_SPHERE, 2 + np.arange(n)
    p["d"][:, 0], p["d"][:, 1], p["d"][:, 2], p["d"][:, 3] = cx, 0.2, cz, 0.2
    ops = np.zeros(1, abi.XFORM_OP)
    ops["kind"] = XF_FLIP
    chains = np.zeros(1, abi.CHAIN)
    chains["count"] = 1
    lights = np.zeros(1, abi.LIGHT)
    lights[0]["Q"], lights[0]["u"], lights[0]["v"], lights[0]["intensity"] = (-100, 100, -100), (200, 0, 0), (0, 0, 200), (15, 15, 15)
    g = np.zeros(1, abi.GLOBALS)
    g[0]["background"] = (0.7, 0.8, 1.0)
    g[0]["image_width"], g[0]["image_height"], g[0]["samples_per_pixel"], g[0]["scene_id"] = width, height, spp, -5
    c = np.zeros(1, abi.CAMERA)
    c[0]["lookfrom"], c[0]["lookat"], c[0]["vup"] = (13 * 40, 2 * 40, 3 * 40), (0, 0, 0), (0, 1, 0)
    c[0]["vfov"], c[0]["aspect_ratio"], c[0]["focus_dist"], c[0]["time1"] = 20.0, width / height, 10.0, 1.0
    return abi.build_blob({"globals": g, "camera": c, "prims": prims, "chains": chains, "xform_ops": ops,
                           "materials": mats, "textures": texs, "lights": lights})


def final_scene(seed: int = 1) -> bytes:
    """scene 9 (final_scene, scenes.cpp:221-290 + case 9, :1595-1604): 400 boxes of random height,
    the rect light, a moving sphere, glass / metal spheres, two constant media (one inside a glass
    boundary that is also a world member, one filling a 5000-radius boundary), the earth sphere
    (earthmap.jpg is absent from the reference repo => cyan, texture.h:96-110), a Perlin sphere
    and 1000 small spheres under rotate_y(15)+translate.  The reference draws from an RNG seeded
    by the thread id; here the same xorshift32 starts from `seed`."""
    rng = XorShift32(seed)
    b = SceneBuilder()
    ground = b.lambertian((0.48, 0.83, 0.53))
    for i in range(20):
        for j in range(20):
            w = 100.0
            x0, z0 = -1000.0 + i * w, -1000.0 + j * w
            y1 = rng.uniform(1, 101)
            b.box((x0, 0.0, z0), (x0 + w, y1, z0 + w), ground)
    b.xz_rect(123, 423, 147, 412, 554, b.diffuse_light((7, 7, 7)))
    b.moving_sphere((400, 400, 200), (430, 400, 200), 0, 1, 50, b.lambertian((0.7, 0.3, 0.1)))
    b.sphere((260, 150, 45), 50, b.dielectric(1.5))
    b.sphere((0, 150, 145), 50, b.metal((0.8, 0.8, 0.9), 1.0))
    glass = b.dielectric(1.5)
    boundary = b.sphere((360, 150, 145), 70, glass)               # world member AND medium boundary
    inner = b.sphere((360, 150, 145), 70, glass)
    b.prims[inner]["flags"] = 1
    b.constant_medium(inner, 0.2, (0.2, 0.4, 0.9))
    fog = b.sphere((0, 0, 0), 5000, b.dielectric(1.5))
    b.prims[fog]["flags"] = 1
    b.constant_medium(fog, 0.0001, (1, 1, 1))
    b.sphere((400, 200, 400), 100, b.lambertian_tex(b.image_missing()))
    b.sphere((220, 280, 300), 80, b.lambertian_tex(b.noise(0.1, rng)))
    white = b.lambertian((.73, .73, .73))
    ch = b.chain([("translate", (-100, 270, 395)), ("rotate_y", 15)])
    for _ in range(1000):
        b.sphere((rng.uniform(0, 165), rng.uniform(0, 165), rng.uniform(0, 165)), 10, white, ch)
    return b.finish(9, 800, 1.0, 500, (0, 0, 0), (478, 278, -600), (278, 278, 0), 40.0)


def synthetic_hdr(width: int = 2048, height: int = 1024, seed: int = 1) -> np.ndarray:
    """The synthetic equirectangular environment of C4-env: vertical sky gradient
    0.2 -> 1.0, one 32x32-texel sun at radiance 5e3, +-5 % xorshift32 noise.  (H, W, 3) float32."""
    v = np.linspace(1.0, 0.2, height, dtype=np.float64)[:, None, None]     # row 0 = zenith
    img = np.broadcast_to(v, (height, width, 3)).copy()
    s = np.uint32(seed)
    n = height * width
    # vectorised xorshift32 stream: one independent lane per row keeps this O(W) python steps
    lanes = (np.arange(height, dtype=np.uint64) * 2654435761 + int(s)) & 0xFFFFFFFF
    lanes = np.where(lanes == 0, 1, lanes).astype(np.uint64)
    noise = np.empty((height, width))
    for x in range(width):
        lanes ^= (lanes << np.uint64(13)) & np.uint64(0xFFFFFFFF)
        lanes ^= lanes >> np.uint64(17)
        lanes ^= (lanes << np.uint64(5)) & np.uint64(0xFFFFFFFF)
        noise[:, x] = lanes * 2.3283064365386963e-10
    img *= (0.95 + 0.1 * noise)[:, :, None]
    y0, x0 = height // 4, (3 * width) // 8
    img[y0:y0 + 32, x0:x0 + 32, :] = 5.0e3
    del n
    return img.astype(np.float32)


def hdr_demo(width: int = 1920, env: np.ndarray = None, spp: int = 200) -> bytes:
    """scene 24 (hdr_demo_scene, scenes.cpp:658-683 + case 24, :1783-1794) at `width` x 16:9 with
    an environment light; env=None reproduces the reference's missing-file behaviour (white)."""
    b = SceneBuilder()
    b.sphere((-4, 1, 0), 1.0, b.metal((0.9, 0.9, 0.9), 0.0))
    b.sphere((0, 1, 0), 1.0, b.pbr((1.0, 0.71, 0.29), 0.2, 1.0))
    b.sphere((4, 1, 0), 1.0, b.dielectric(1.5))
    b.env_light(env)
    return b.finish(24, width, 16.0 / 9.0, spp, (0, 0, 0), (0, 3, 10), (0, 1, 0), 30.0)


BUILDERS = {7: lambda: cornell_box(False), 21: lambda: cornell_box(True), 23: mis_comparison, 9: final_scene,
            24: lambda: hdr_demo(800, None)}


def select_scene(scene_id: int) -> bytes:
    """select_scene(int) of the reference (scenes.cpp:1523) for the BASELINE.json scene ids."""
    if scene_id not in BUILDERS:
        raise KeyError(f"scene {scene_id}: only the BASELINE.json configurations {sorted(BUILDERS)} are built in")
    return BUILDERS[scene_id]()
