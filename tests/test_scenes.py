"""The built-in scene builders (ray_tracing-rendering_b200/scenes.py) must construct exactly
what the reference's builders construct: compared primitive by primitive (with materials,
textures and wrapper chains resolved) against the blobs obtained by walking the reference's
own select_scene() graphs (tests/golden)."""
import importlib

import numpy as np
import pytest

from conftest import PKG


def canonical(T):
    def tex(i):
        if i < 0:
            return None
        t = T["textures"][i]
        return (int(t["type"]), tuple(t["color"]), float(t["scale"]), tex(int(t["even"])) if t["type"] == 1 else None,
                tex(int(t["odd"])) if t["type"] == 1 else None)

    def mat(i):
        m = T["materials"][i]
        return (int(m["type"]), tuple(tex(int(k)) for k in m["tex"]), tuple(m["color"]), float(m["fuzz"]), float(m["ir"]))

    def chain(i):
        if i < 0:
            return ()
        c = T["chains"][i]
        return tuple((int(o["kind"]), float(o["a"]), float(o["b"]), float(o["c"]))
                     for o in T["xform_ops"][c["first"]:c["first"] + c["count"]])
    prims = sorted((int(p["type"]), tuple(p["d"]), chain(int(p["chain"])), mat(int(p["material"])), int(p["flags"]) & 1)
                   for p in T["prims"])
    lights = [(int(l["type"]), tuple(l["Q"]), tuple(l["u"]), tuple(l["v"]), tuple(l["intensity"]), float(l["cos_cutoff"]),
               int(l["env_width"])) for l in T["lights"]]
    return prims, lights, T["globals"].tobytes(), T["camera"].tobytes()


@pytest.mark.parametrize("sid", [7, 21, 23])
def test_builtin_scene_equals_reference_scene(abi, golden, sid):
    scenes = importlib.import_module(PKG + ".scenes")
    mine = canonical(abi.parse_blob(scenes.select_scene(sid)))
    ref = canonical(abi.parse_blob(golden(sid).blob))
    assert mine[0] == ref[0]          # primitives, bit for bit (incl. sin/cos of rotate_y)
    assert mine[1] == ref[1]          # lights
    assert mine[2] == ref[2] and mine[3] == ref[3]   # SceneConfig + camera arguments


def test_unknown_scene_is_an_error():
    scenes = importlib.import_module(PKG + ".scenes")
    with pytest.raises(KeyError):
        scenes.select_scene(3)


def test_sphere_field_shape(abi):
    scenes = importlib.import_module(PKG + ".scenes")
    T = abi.parse_blob(scenes.sphere_field(half_extent=8, width=64, height=36, spp=4))
    assert len(T["prims"]) == 16 * 16 + 2 and len(T["materials"]) == 16 * 16 + 2   # one material per sphere
    assert len(T["lights"]) == 1 and (T["prims"]["type"] == 3).sum() == 1
    c = T["prims"]["d"][2:, :3]
    assert np.all(c[:, 1] == 0.2) and c[:, 0].min() >= -8 and c[:, 0].max() < 8
    kinds = T["materials"]["type"][2:]
    assert set(np.unique(kinds)) <= {0, 1, 2}
    # deterministic for a given seed
    assert scenes.sphere_field(8, 64, 36, 4) == scenes.sphere_field(8, 64, 36, 4)


def test_final_scene_has_the_reference_census(abi, golden):
    """scene 9 is random in the reference (thread-id seeded RNG), so the built-in builder cannot
    reproduce one instance; it must construct the same KIND of scene: the same primitive /
    material / texture tables up to the random numbers, the same wrapper chain and camera."""
    scenes = importlib.import_module(PKG + ".scenes")
    mine, ref = abi.parse_blob(scenes.final_scene(1)), abi.parse_blob(golden(9).blob)
    for sec in ("prims", "materials", "textures", "chains", "xform_ops", "images", "perlins", "lights"):
        assert len(mine[sec]) == len(ref[sec]), sec
    assert sorted(mine["prims"]["type"].tolist()) == sorted(ref["prims"]["type"].tolist())
    assert sorted(mine["materials"]["type"].tolist()) == sorted(ref["materials"]["type"].tolist())
    assert mine["xform_ops"].tobytes() == ref["xform_ops"].tobytes()
    assert mine["camera"].tobytes() == ref["camera"].tobytes() and mine["globals"].tobytes() == ref["globals"].tobytes()
    # the deterministic members are identical: light, moving sphere, the six free spheres, both media
    def fixed(T):
        p = T["prims"]
        keep = p[(p["chain"] < 0) & (p["type"] != 2) & (p["type"] != 4) & ~((p["type"] == 3) & (p["d"][:, 4] != 554))]
        return sorted((int(q["type"]), tuple(q["d"]), int(q["flags"]) & 1) for q in keep)
    assert fixed(mine) == fixed(ref)
    # boxes: 20 x 20 grid of 100-wide boxes with heights in [1, 101)
    tops = mine["prims"][(mine["prims"]["type"] == 3) & (mine["prims"]["d"][:, 4] != 554) & (mine["prims"]["d"][:, 4] != 0)]
    assert len(tops) == 400 and tops["d"][:, 4].min() >= 1 and tops["d"][:, 4].max() < 101
    # 1000 spheres of radius 10 in [0,165)^3 under the instance chain
    sp = mine["prims"][mine["prims"]["chain"] == 0]
    assert len(sp) == 1000 and np.all(sp["d"][:, 3] == 10) and sp["d"][:, :3].min() >= 0 and sp["d"][:, :3].max() < 165
    # Perlin tables: unit gradients, three permutations of 0..255
    pl = mine["perlins"][0]
    assert np.allclose(np.linalg.norm(pl["ranvec"], axis=1), 1.0)
    for k in ("perm_x", "perm_y", "perm_z"):
        assert sorted(pl[k].tolist()) == list(range(256))
    assert scenes.final_scene(1) == scenes.final_scene(1) and scenes.final_scene(1) != scenes.final_scene(2)


def test_hdr_demo_and_synthetic_environment(abi, golden):
    scenes = importlib.import_module(PKG + ".scenes")
    # without an image: exactly the reference's scene 24 (missing .hdr => white environment)
    mine, ref = canonical(abi.parse_blob(scenes.hdr_demo(800, None))), canonical(abi.parse_blob(golden(24).blob))
    assert mine == ref
    env = scenes.synthetic_hdr(256, 128, 1)
    assert env.shape == (128, 256, 3) and env.dtype == np.float32
    assert env.max() == 5.0e3 and (env == 5.0e3).sum() == 32 * 32 * 3          # the sun
    sky = np.where(env == 5.0e3, np.nan, env)
    assert np.nanmean(sky[:16]) > np.nanmean(sky[-16:]) and 0.15 < np.nanmin(sky) and np.nanmax(sky) < 1.1
    assert np.array_equal(env, scenes.synthetic_hdr(256, 128, 1))
    T = abi.parse_blob(scenes.hdr_demo(1920, env))
    assert T["lights"][0]["env_width"] == 256 and T["lights"][0]["env_height"] == 128 and len(T["env_texels"]) == env.size


def test_configs_name_the_baseline_workloads(abi):
    configs = importlib.import_module(PKG + ".configs")
    assert sorted(configs.CONFIGS) == ["C1", "C2", "C3", "C4", "C4env", "C5"]
    c1 = configs.get("C1")
    assert (c1.width, c1.height, c1.spp, c1.integrator, c1.depth, c1.scene_id) == (600, 600, 400, 1, 50, 7)
    for name in ("C1", "C2", "C3", "C4"):
        c = configs.get(name)
        g = abi.parse_blob(c.blob())["globals"][0]
        assert (g["image_width"], g["image_height"], g["samples_per_pixel"], g["scene_id"]) == (c.width, c.height, c.spp, c.scene_id)
    with pytest.raises(KeyError):
        configs.get("C9")
