"""Generates the golden fixtures in this directory from the UNMODIFIED reference
(oracle/_ref/libref_oracle.so, built by `make -C oracle ref` from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference's RNG is seeded from the thread id (rtweekend.h:26-27) and is not
reproducible, so the fixtures are generated ONCE and committed; they pin the
oracle restatement (oracle/port) and the CUDA library to answers computed by the
reference itself:

  sceneNN.npz
    blob            the flattened scene (include/rtb200_scene.h), produced by walking
                    the reference's own select_scene(NN) graph
    rays, hits      hit() queries issued by Integrator::Li (camera, bounce, shadow
                    rays) with the reference's answers
    camera          camera.h:44-50 derived members
    bsdf_q_<m>, bsdf_v_<m>   material::eval/pdf/emitted on random (normal, wo, wi)
    light_q_<l>, light_v_<l> Light::sample / pdf / Le
    tex_q_<t>, tex_v_<t>     texture::value
    img_<integrator>_{sum,sumsq,spp}  per-pixel sum / sum of squares of linear Li
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402

abi = refbind.abi
HERE = os.path.dirname(os.path.abspath(__file__))


def write_hdr(path, rgb):
    """Minimal flat (non-RLE) Radiance RGBE writer; rgb: (H, W, 3) float."""
    rgb = np.asarray(rgb, np.float64)
    h, w, _ = rgb.shape
    m = rgb.max(axis=2)
    e = np.where(m > 1e-32, np.floor(np.log2(np.maximum(m, 1e-38))) + 1, 0)
    scale = np.where(m > 1e-32, 256.0 / np.exp2(e), 0)
    out = np.zeros((h, w, 4), np.uint8)
    out[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    out[..., 3] = np.where(m > 1e-32, e + 128, 0).astype(np.uint8)
    # a scanline starting with bytes (2, 2, <128) would be parsed as RLE
    bad = (out[:, 0, 0] == 2) & (out[:, 0, 1] == 2) & (out[:, 0, 2] < 128)
    out[bad, 0, 0] = 3
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(f"-Y {h} +X {w}\n".encode())
        f.write(out.tobytes())


def synthetic_sky(w, h, seed, sun=(100.0, 90.0, 80.0)):
    """Vertical gradient + a small bright 'sun' + +-5 % noise (the shape of SURVEY §8d's C4 env
    input; the fixture's sun is 100x dimmer than the bench input so that 1k samples converge)."""
    rng = np.random.default_rng(seed)
    v = (np.arange(h) + 0.5) / h
    img = np.zeros((h, w, 3))
    img[:] = (1.0 - 0.8 * v)[:, None, None] * np.array([0.6, 0.75, 1.0])
    img *= 1.0 + 0.05 * (2 * rng.random((h, w, 1)) - 1)
    sy, sx = h // 4, (2 * w) // 3
    img[sy:sy + max(1, h // 16), sx:sx + max(1, w // 32)] = np.array(sun)
    return img


def unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def bsdf_queries(n, rng):
    q = np.zeros(n, abi.BSDF_QUERY)
    nrm = unit(rng.normal(size=(n, 3)))
    q["normal"] = nrm
    q["p"] = rng.uniform(-5, 5, size=(n, 3))
    q["u"] = rng.random(n)
    q["v"] = rng.random(n)
    q["front_face"] = rng.integers(0, 2, n)
    # wo in the normal's hemisphere (as after set_face_normal), wi mostly so; a quarter below
    wo = unit(rng.normal(size=(n, 3)))
    wo = np.where((np.sum(wo * nrm, axis=1) < 0)[:, None], -wo, wo)
    wi = unit(rng.normal(size=(n, 3)))
    flip = (np.sum(wi * nrm, axis=1) < 0) & (rng.random(n) < 0.75)
    wi = np.where(flip[:, None], -wi, wi)
    # a third of the queries near the specular peak (wi ~ reflect(-wo, n)), where GGX is stiff
    k = n // 3
    refl = 2 * np.sum(wo[:k] * nrm[:k], axis=1, keepdims=True) * nrm[:k] - wo[:k]
    wi[:k] = unit(refl + 0.02 * rng.normal(size=(k, 3)))
    q["wo"] = wo
    q["wi"] = wi
    return q


def light_queries(n, rng, scale):
    q = np.zeros(n, abi.LIGHT_QUERY)
    q["p"] = rng.uniform(-scale, scale, size=(n, 3))
    q["d"] = rng.normal(size=(n, 3)) * rng.uniform(0.2, 3.0, size=(n, 1))
    q["u"] = rng.random((n, 2))
    return q


only = set()

# (scene id, integrator, image size, spp, light query scale): integrator 4 (MIS) or 3 (NEE) where the scene
# has lights, 2 (PBR path) for the PBR galleries, 1 / 0 otherwise
CATALOGUE = [(2, 1, (64, 36), 128, 10.0), (4, 1, (64, 36), 128, 10.0), (5, 1, (64, 36), 256, 10.0), (6, 0, (64, 36), 256, 10.0),
             (10, 1, (64, 36), 128, 10.0), (11, 2, (64, 36), 256, 10.0), (12, 2, (48, 48), 256, 10.0),
             (13, 2, (64, 36), 256, 10.0), (14, 2, (64, 36), 256, 10.0), (16, 4, (64, 36), 256, 10.0),
             (20, 3, (64, 36), 256, 10.0), (22, 4, (48, 48), 128, 500.0), (25, 4, (64, 36), 512, 10.0),
             (27, 4, (64, 36), 512, 10.0), (28, 3, (64, 36), 512, 10.0), (30, 4, (64, 36), 256, 10.0),
             (31, 4, (48, 48), 256, 500.0), (32, 4, (64, 36), 256, 10.0), (33, 4, (64, 36), 256, 10.0),
             (34, 4, (64, 36), 256, 10.0), (35, 4, (64, 36), 256, 10.0), (36, 4, (64, 36), 256, 10.0),
             (37, 4, (48, 48), 256, 10.0), (38, 3, (64, 36), 256, 10.0), (39, 4, (64, 36), 256, 10.0),
             (40, 4, (64, 36), 256, 10.0), (41, 3, (64, 36), 256, 10.0), (42, 1, (64, 36), 256, 10.0)]


def write_assets(tmp):
    """Synthetic stand-ins for the assets the reference's scenes load by fixed relative names and that
    are absent from its repository (SURVEY 8c "Missing assets"): with them in the cwd image_texture /
    EnvironmentLight hold real texel data (stb decodes them inside the reference; the blob carries the
    decoded texels), without them the reference's cyan / white fallbacks apply."""
    from PIL import Image
    rng = np.random.default_rng(7)

    def pattern(w, h, base, amp, freq):
        y, x = np.mgrid[0:h, 0:w]
        img = np.zeros((h, w, 3))
        for c in range(3):
            img[..., c] = base[c] + amp[c] * np.sin(freq[0] * x / w * 2 * np.pi + c) * np.cos(freq[1] * y / h * 2 * np.pi - c)
        img += 0.04 * (rng.random((h, w, 3)) - 0.5)
        return (np.clip(img, 0, 1) * 255).astype(np.uint8)

    Image.fromarray(pattern(256, 128, (0.35, 0.5, 0.55), (0.3, 0.35, 0.4), (3, 2))).save(os.path.join(tmp, "earthmap.jpg"), quality=92)
    for d, stem, base in (("oak", "oak_veneer_01", (0.55, 0.38, 0.2)), ("brick", "red_brick", (0.6, 0.25, 0.2)),
                          ("rust", "rusty_metal_04", (0.45, 0.3, 0.22))):
        os.makedirs(os.path.join(tmp, "tex", d), exist_ok=True)
        Image.fromarray(pattern(64, 64, base, (0.25, 0.2, 0.15), (4, 3))).save(os.path.join(tmp, "tex", d, stem + "_diff_1k.png"))
        r = pattern(48, 40, (0.5, 0.5, 0.5), (0.4, 0.4, 0.4), (2, 5))
        r[..., 1] = r[..., 2] = r[..., 0]
        Image.fromarray(r).save(os.path.join(tmp, "tex", d, stem + "_rough_1k.png"))
        m = pattern(32, 32, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5), (3, 3))
        m[..., 1] = m[..., 2] = m[..., 0] = np.where(m[..., 0] > 127, 230, 20)
        Image.fromarray(m).save(os.path.join(tmp, "tex", d, stem + "_metal_1k.png"))
        # tangent-space normal map: z up, x / y tilted by a smooth bump field (texture.h:19-23 decodes 2c - 1)
        y, x = np.mgrid[0:64, 0:64]
        nx = 0.35 * np.sin(x / 64 * 6 * np.pi)
        ny = 0.35 * np.cos(y / 64 * 4 * np.pi)
        nz = np.sqrt(np.maximum(1 - nx * nx - ny * ny, 0.05))
        n = np.stack([nx, ny, nz], axis=-1)
        Image.fromarray(((n * 0.5 + 0.5) * 255).astype(np.uint8)).save(os.path.join(tmp, "tex", d, stem + "_nor_dx_1k.png"))
    write_hdr(os.path.join(tmp, "brown_photostudio_02_4k.hdr"), synthetic_sky(64, 32, 3, sun=(60.0, 55.0, 50.0)))
    write_hdr(os.path.join(tmp, "cedar_bridge_sunset_2_4k.hdr"), synthetic_sky(48, 24, 4, sun=(80.0, 50.0, 30.0)))
    write_hdr(os.path.join(tmp, "stpeters_probe.hdr"), synthetic_sky(32, 32, 5))
    write_hdr(os.path.join(tmp, "uffizi_probe.hdr"), synthetic_sky(24, 24, 6))


def make_scene(sid, integrators, n_paths, n_rays, img_wh, img_spp, rng, light_scale=10.0, n_bsdf=400, n_light=600,
               n_tex=300, max_mats=16):
    if only and sid not in only:
        return
    s = refbind.RefScene(sid)
    blob = s.blob()
    T = abi.parse_blob(blob)
    out = {"blob": np.frombuffer(blob, np.uint8), "camera": s.camera_derived()}
    rays_all, hits_all = [], []
    for integ in integrators:
        r, h, _ = s.record_rays(integ, n_paths, n_rays // len(integrators))
        rays_all.append(r)
        hits_all.append(h)
    out["rays"] = np.concatenate(rays_all)
    out["hits"] = np.concatenate(hits_all)
    nm = len(T["materials"])
    for m in (range(nm) if nm <= max_mats else rng.choice(nm, 6, replace=False)):
        q = bsdf_queries(n_bsdf, rng)
        out[f"bsdf_q_{m}"] = q
        out[f"bsdf_v_{m}"] = s.bsdf_eval(int(m), q)
    for l in range(len(T["lights"])):
        q = light_queries(n_light, rng, light_scale)
        out[f"light_q_{l}"] = q
        out[f"light_v_{l}"] = s.light_eval(l, q)
        out[f"light_flags_{l}"] = np.array([s.light_flags(l)])
    nt = len(T["textures"])
    for t in (range(nt) if nt <= max_mats else rng.choice(nt, 4, replace=False)):
        uvp = np.concatenate([rng.random((n_tex, 2)), rng.uniform(-300, 600, size=(n_tex, 3))], axis=1)
        out[f"tex_q_{t}"] = uvp
        out[f"tex_v_{t}"] = s.texture_value(int(t), uvp)
    w, h = img_wh
    for integ in integrators:
        S, S2, cnt = s.render_linear(integ, w, h, img_spp)
        out[f"img_{integ}_sum"] = S.astype(np.float32)
        out[f"img_{integ}_sumsq"] = S2.astype(np.float32)
        out[f"img_{integ}_spp"] = np.array([img_spp])
        out[f"img_{integ}_rays"] = cnt
    path = os.path.join(HERE, f"scene{sid:02d}.npz")
    np.savez_compressed(path, **out)
    print(f"scene {sid}: {os.path.getsize(path) / 1024:.0f} KiB, {len(out['rays'])} rays,"
          f" {nm} materials, {len(T['lights'])} lights")


def make_media_seeded(rng):
    """media08.npz / media09.npz: fresh instances of the two media scenes with the reference's generator
    state recorded per query (rtb_ray.reserved), for the deterministic constant_medium parity tests."""
    for sid, integ in ((8, 1), (9, 1)):
        s = refbind.RefScene(sid)
        r, h, _ = s.record_rays(integ, 1500, 4000)
        path = os.path.join(HERE, f"media{sid:02d}.npz")
        np.savez_compressed(path, blob=np.frombuffer(s.blob(), np.uint8), rays=r, hits=h)
        print(f"media fixture {sid}: {os.path.getsize(path) / 1024:.0f} KiB, {len(r)} rays")


def make_fullres():
    """Whole-image means of linear Li at the FULL resolution of the BASELINE.json configs
    (the image mean depends on the resolution through u=(i+xi)/(W-1), renderer.h:73-74).
    mean_<scene>_<integrator> = [r, g, b, standard error r, g, b]."""
    out = {}
    for sid, integ, spp in ((7, 1, 48), (7, 0, 24), (21, 3, 32), (21, 4, 32), (23, 4, 64), (23, 3, 64), (9, 1, 12)):
        s = refbind.RefScene(sid)
        g = abi.parse_blob(s.blob())["globals"][0]
        w, h = int(g["image_width"]), int(g["image_height"])
        S, S2, cnt = s.render_linear(integ, w, h, spp)
        mean = S / spp
        var = np.maximum(S2 / spp - mean ** 2, 0) / spp
        se = np.sqrt(var.sum(axis=(0, 1))) / (w * h)
        out[f"mean_{sid}_{integ}"] = np.concatenate([mean.mean(axis=(0, 1)), se])
        out[f"rays_{sid}_{integ}"] = cnt / (w * h * spp)
        print(f"fullres scene {sid} int {integ}: {w}x{h}x{spp} mean {mean.mean(axis=(0, 1))} se {se} rays/path {cnt / (w * h * spp)}")
    np.savez_compressed(os.path.join(HERE, "fullres_means.npz"), **out)


def main():
    only.update(int(a) for a in sys.argv[1:])      # scene ids; 0 = the full-resolution means; -1 = the seeded media fixtures
    if not refbind.available():
        raise SystemExit("oracle/_ref/libref_oracle.so missing: run `make -C oracle ref` first")
    rng = np.random.default_rng(20261018)
    with tempfile.TemporaryDirectory() as tmp:
        # the env-light scenes load their .hdr by bare file name from the cwd
        write_hdr(os.path.join(tmp, "sky.hdr"), synthetic_sky(64, 32, 1))        # scene 19, equirect
        write_hdr(os.path.join(tmp, "rnl_probe.hdr"), synthetic_sky(32, 32, 2))  # scene 26, light probe
        os.chdir(tmp)
        make_scene(7, [0, 1], 800, 2400, (64, 64), 512, rng)                  # C1 Cornell box
        make_scene(21, [3, 4], 500, 2400, (64, 64), 512, rng, 500.0)          # C3 Cornell + NEE
        make_scene(23, [2, 3, 4], 600, 2400, (80, 45), 512, rng)              # C4 MIS comparison
        make_scene(9, [1], 1000, 2400, (64, 64), 256, rng)                    # C2 final scene
        make_scene(1, [1], 1000, 2000, (64, 64), 256, rng)                    # moving spheres, template of C5
        make_scene(19, [3, 4], 600, 1600, (80, 45), 1024, rng)                 # equirect env map (synthetic)
        make_scene(26, [4], 800, 1200, (80, 45), 1024, rng)                   # light-probe env map (synthetic)
        make_scene(24, [4], 800, 1200, (80, 45), 256, rng)                    # env map missing -> white fallback
        make_scene(15, [3], 500, 1200, (80, 45), 256, rng)                    # point light
        make_scene(17, [4], 500, 1200, (80, 45), 256, rng)                    # directional light
        make_scene(18, [3], 500, 1200, (80, 45), 256, rng)                    # spot light
        make_scene(8, [1], 600, 1600, (64, 64), 256, rng)                     # cornell_smoke (media in instances)
        if not only or 0 in only:
            make_fullres()
        if not only or -1 in only:
            make_media_seeded(rng)
        # the rest of the reference's catalogue (scenes.cpp:1523-2096) as smaller fixtures: hits, a few BSDF /
        # light / texture grids and one image each; synthetic assets in the cwd (write_assets), so image
        # textures, roughness / metallic / normal maps and the four remaining .hdr names carry texel data
        write_assets(tmp)
        lite = dict(n_bsdf=120, n_light=200, n_tex=120, max_mats=8)
        for sid, integ, wh, spp, scale in CATALOGUE:
            make_scene(sid, [integ], 400, 1200, wh, spp, rng, scale, **lite)
        os.chdir(ROOT)


if __name__ == "__main__":
    main()
