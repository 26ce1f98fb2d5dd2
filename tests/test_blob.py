"""The flat scene blob: framing, cross-reference contents of the fixtures produced by
walking the reference's own scene graphs, and the python writer/reader round trip."""
import numpy as np
import pytest

from conftest import ALL_SCENES, GOLDEN_SCENES


@pytest.mark.parametrize("sid", ALL_SCENES)
def test_blob_parses_and_is_consistent(abi, golden, sid):
    T = abi.parse_blob(golden(sid).blob)
    g = T["globals"][0]
    assert g["scene_id"] == sid
    assert g["image_height"] == int(g["image_width"] / T["camera"][0]["aspect_ratio"])  # main.cpp:69
    prims = T["prims"]
    assert len(prims) > 0
    assert prims["material"].min() >= 0 and prims["material"].max() < len(T["materials"])
    assert prims["chain"].min() >= -1 and prims["chain"].max() < len(T["chains"])
    ch = T["chains"]
    if len(ch):
        assert (ch["first"] + ch["count"]).max() <= len(T["xform_ops"])
    med = prims[prims["type"] == 5]
    for m in med:
        b = prims[m["aux0"]:m["aux0"] + m["aux1"]]
        assert len(b) == m["aux1"] and (b["flags"] & 1).all()  # boundary-only prims


def test_cornell_box_contents(abi, golden):
    # scenes.cpp:159-187: 6 free rects + 2 boxes x 6 rects, two rotate+translate chains
    T = abi.parse_blob(golden(7).blob)
    assert len(T["prims"]) == 18 and len(T["chains"]) == 2 and len(T["lights"]) == 0
    assert sorted(T["materials"]["type"].tolist()) == [0, 0, 0, 3]
    ops = T["xform_ops"]
    assert ops["kind"].tolist() == [0, 1, 0, 1]  # translate(rotate_y(box)) listed outermost first
    assert np.allclose([ops[0]["a"], ops[0]["b"], ops[0]["c"]], [265, 0, 295])
    assert np.isclose(ops[1]["a"], np.sin(np.radians(15))) and np.isclose(ops[3]["a"], np.sin(np.radians(-18)))
    g = T["globals"][0]
    assert (g["image_width"], g["image_height"], g["samples_per_pixel"]) == (600, 600, 400)


def test_scene21_and_23_lights(abi, golden):
    T = abi.parse_blob(golden(21).blob)
    assert len(T["lights"]) == 1 and T["lights"][0]["type"] == 0
    assert np.allclose(T["lights"][0]["Q"], [213, 554, 227])
    # the emitter is wrapped in flip_face (scenes.cpp:790-791)
    assert (T["xform_ops"]["kind"] == 2).sum() == 1
    T = abi.parse_blob(golden(23).blob)
    assert len(T["lights"]) == 2
    assert sorted(T["materials"]["type"].tolist()) == [0, 2, 3, 3, 4, 4]


def test_final_scene_census(abi, golden):
    # SURVEY §8a census of scene 9: 2,401 rects, 1 moving sphere, 2 media
    T = abi.parse_blob(golden(9).blob)
    t = T["prims"]["type"]
    assert ((t >= 2) & (t <= 4)).sum() == 2401
    assert (t == 1).sum() == 1 and (t == 5).sum() == 2
    assert (t == 0).sum() == 1007
    assert len(T["perlins"]) == 1 and len(T["images"]) == 1 and T["images"][0]["width"] == 0  # earthmap.jpg missing


def test_env_fixture_has_texels(abi, golden):
    T = abi.parse_blob(golden(19).blob)
    l = T["lights"][0]
    assert l["type"] == 4 and (l["env_width"], l["env_height"], l["env_is_probe"]) == (64, 32, 0)
    assert len(T["env_texels"]) == 64 * 32 * 3 and T["env_texels"].max() > 50
    T = abi.parse_blob(golden(26).blob)
    assert T["lights"][0]["env_is_probe"] == 1
    T = abi.parse_blob(golden(24).blob)
    assert T["lights"][0]["env_width"] == 0  # missing .hdr -> white fallback


@pytest.mark.parametrize("sid", [7, 9, 19])
def test_python_blob_roundtrip(abi, golden, sid):
    blob = golden(sid).blob
    T = abi.parse_blob(blob)
    again = abi.build_blob(T)
    T2 = abi.parse_blob(again)
    for k in T:
        assert T[k].tobytes() == T2[k].tobytes(), k


def test_bad_blob_rejected(abi):
    with pytest.raises(ValueError):
        abi.parse_blob(b"\0" * 64)
