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


def make_scene(sid, integrators, n_paths, n_rays, img_wh, img_spp, rng, light_scale=10.0):
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
    for m in (range(nm) if nm <= 16 else rng.choice(nm, 6, replace=False)):
        q = bsdf_queries(400, rng)
        out[f"bsdf_q_{m}"] = q
        out[f"bsdf_v_{m}"] = s.bsdf_eval(int(m), q)
    for l in range(len(T["lights"])):
        q = light_queries(600, rng, light_scale)
        out[f"light_q_{l}"] = q
        out[f"light_v_{l}"] = s.light_eval(l, q)
        out[f"light_flags_{l}"] = np.array([s.light_flags(l)])
    nt = len(T["textures"])
    for t in (range(nt) if nt <= 16 else rng.choice(nt, 4, replace=False)):
        uvp = np.concatenate([rng.random((300, 2)), rng.uniform(-300, 600, size=(300, 3))], axis=1)
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
    only.update(int(a) for a in sys.argv[1:])
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
        os.chdir(ROOT)


if __name__ == "__main__":
    main()
