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
