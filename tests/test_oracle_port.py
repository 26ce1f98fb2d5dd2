"""Pins the CPU restatement (oracle/port) to the golden vectors the unmodified reference
produced: hits bit for bit, BSDF / light / texture values to 1e-12, camera frames bit for
bit, images statistically (same gates as the GPU)."""
import numpy as np
import pytest

import parity
from conftest import ALL_SCENES, CATALOGUE_CASES, GOLDEN_SCENES


@pytest.fixture(scope="module")
def port():
    from oracle import portbind
    assert portbind.available(), "oracle/liboracle_port.so missing: run __graft_entry__.build()"
    cache = {}

    def get(golden, sid):
        if sid not in cache:
            cache[sid] = portbind.PortScene(golden(sid).blob)
        return cache[sid]
    return get


@pytest.mark.parametrize("sid", ALL_SCENES)
def test_hits_bit_exact(port, golden, abi, sid):
    g = golden(sid)
    T = abi.parse_blob(g.blob)
    got = port(golden, sid).trace(g["rays"])
    mask = parity.deterministic_mask(T, g["hits"], got)
    # the port numbers primitives exactly as the blob does, and tests them in blob order, which is
    # the reference's own traversal order (the walker emitted leaves in that order): ids included
    assert parity.trace_mismatches(g["hits"], got, mask, fields=("t", "p", "normal", "front_face", "material")) == 0
    same_id = (got["prim"] == g["hits"]["prim"])[mask]
    assert same_id.mean() >= 0.9999
    ok = mask & (g["hits"]["prim"] >= 0)
    sph = ok & (T["prims"]["type"][np.maximum(g["hits"]["prim"], 0)] != 1)   # moving spheres leave u,v unset
    assert np.array_equal(got["u"][sph], g["hits"]["u"][sph]) and np.array_equal(got["v"][sph], g["hits"]["v"][sph])


@pytest.mark.parametrize("sid", ALL_SCENES)
def test_camera_bit_exact(port, golden, sid):
    assert np.array_equal(port(golden, sid).camera_derived(), golden(sid)["camera"])


@pytest.mark.parametrize("sid", ALL_SCENES)
def test_bsdf_light_texture_values(port, golden, sid):
    g = golden(sid)
    s = port(golden, sid)
    for m in g.keys("bsdf_q_"):
        got, ref = s.bsdf_eval(m, g[f"bsdf_q_{m}"]), g[f"bsdf_v_{m}"]
        for f in ("f", "pdf", "emitted_old", "emitted_new"):
            assert parity.values_close(got[f], ref[f], 1e-12, 1e-300).all(), (sid, m, f)
    for l in g.keys("light_q_"):
        got, ref = s.light_eval(l, g[f"light_q_{l}"]), g[f"light_v_{l}"]
        fields = ("pdf", "dist", "is_delta", "pdf_dir", "Le") if sid == 24 else \
            ("Li", "wi", "pdf", "dist", "is_delta", "pdf_dir", "Le")
        for f in fields:
            assert parity.values_close(got[f], ref[f], 1e-12, 1e-300).all(), (sid, l, f)
    for t in g.keys("tex_q_"):
        assert parity.values_close(s.texture_value(t, g[f"tex_q_{t}"]), g[f"tex_v_{t}"], 1e-12, 1e-300).all(), (sid, t)


@pytest.mark.parametrize("sid,integrator", [(7, 0), (7, 1), (21, 3), (21, 4), (23, 2), (23, 3), (23, 4), (9, 1), (19, 4),
                                            (24, 4), (15, 3), (17, 4), (18, 3), (8, 1)] + CATALOGUE_CASES)
def test_images_match_reference_statistics(port, golden, sid, integrator):
    g = golden(sid)
    ref_sum, ref_sumsq = g[f"img_{integrator}_sum"], g[f"img_{integrator}_sumsq"]
    ref_spp = int(g[f"img_{integrator}_spp"][0])
    h, w, _ = ref_sum.shape
    k, spp = 4, max(ref_spp // 8, 32)
    means = []
    rays = 0
    for i in range(k):
        s, _, cnt, _ = port(golden, sid).render_linear(integrator, w, h, spp, seed=11 + i)
        means.append(s / spp)
        rays += int(cnt[0]) + int(cnt[1])
    rep = parity.image_report(ref_sum, ref_sumsq, ref_spp, np.stack(means))
    cal = parity.selfcal(sid, integrator)
    assert parity.image_gates(rep, cal is not None) == [], rep
    if cal is not None and (sid, integrator) in CATALOGUE_CASES:
        # the calibrated gates the GPU renderer is held to (tests/test_gpu_render.py), rehearsed on the restatement
        equal = [port(golden, sid).render_linear(integrator, w, h, ref_spp, seed=300 + i)[0] / ref_spp for i in range(4)]
        hi = port(golden, sid).render_linear(integrator, w, h, 8 * ref_spp, seed=77)[0] / (8 * ref_spp)
        bad = parity.strict_image_gates(ref_sum, ref_sumsq, ref_spp, np.stack(equal), hi.mean(axis=(0, 1)), cal)
        assert bad == [], (bad, cal)
    # closest + shadow: the reference-side recorder classifies a shadow ray towards an infinite
    # light (t_max = inf) as a closest-hit query, so only the total is comparable
    ref_rpp = float(g[f"img_{integrator}_rays"].sum()) / (w * h * ref_spp)
    assert abs(rays / (k * w * h * spp) - ref_rpp) <= 0.02 * ref_rpp


MEDIA_FIXTURES = ["media08", "media09", "scene22"]


def load_media(name):
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    return z["blob"].tobytes(), z["rays"], z["hits"]


@pytest.mark.parametrize("name", MEDIA_FIXTURES)
def test_media_hits_bit_exact_with_the_reference_random_stream(abi, name):
    """constant_medium::hit draws its free-flight distance from the reference's generator
    (constant_medium.h:85).  The fixtures carry that generator's state per query (rtb_ray.reserved, read
    off by the harness), so the restatement — walking the leaves in the reference's order and drawing
    where the reference draws — must give the reference's answer on EVERY query, media included."""
    from oracle import portbind
    blob, rays, ref = load_media(name)
    assert (rays["reserved"] != 0).all()
    T = abi.parse_blob(blob)
    got = portbind.PortScene(blob).trace(rays)
    every = np.ones(len(rays), bool)
    assert parity.is_medium(T, ref["prim"]).sum() > 50          # the media are actually hit
    assert parity.trace_mismatches(ref, got, every, fields=("prim", "t", "p", "normal", "front_face", "material")) == 0


def test_reference_output_path_is_the_documented_conversion(tmp_path):
    """The formula the host-layer PNG test and the device resolve are held to IS the reference's output
    path: random sums through the reference's write_color_to_buffer + save_to_png (oracle/_ref), decoded."""
    from PIL import Image
    from oracle import refbind
    if not refbind.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    sums = rng.random((37, 53, 3)) * 40.0                 # 16 samples: means up to 2.5 (clamped)
    sums[0, 0] = 0
    sums[1, 1] = 16.0                                     # exactly 1.0
    buf = refbind.output_path(sums, 16, str(tmp_path / "o.png"))
    img = np.asarray(Image.open(str(tmp_path / "o.png")).convert("RGB"))
    want_buf = np.clip(np.sqrt((1.0 / 16) * sums), 0, 1)
    assert np.array_equal(buf, want_buf)
    assert np.array_equal(img, (want_buf[::-1] * 255).astype(np.uint8))
