"""Error behaviour and edge cases of the C-ABI on a live device."""
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_loaded_library_is_the_in_tree_build(binding, gpu_ctx):
    import os
    assert os.path.samefile(binding.LIB_PATH, os.path.join(os.path.dirname(binding.__file__), "librtb200.so"))
    with open("/proc/self/maps") as f:
        assert any("librtb200.so" in ln for ln in f)


def test_render_before_upload_fails(binding):
    ctx = binding.Context(0)
    with pytest.raises(binding.RtbError) as e:
        ctx.render(ctx.params(16, 16, 1, 1))
    assert e.value.status == -5
    with pytest.raises(binding.RtbError):
        ctx.trace(np.zeros(1, binding.abi.RAY))
    ctx.close()


def test_bad_blobs_are_rejected(binding, golden):
    ctx = binding.Context(0)
    blob = bytearray(golden(7).blob)
    for bad in (b"", b"\0" * 100, bytes(blob[:200])):
        with pytest.raises(binding.RtbError) as e:
            ctx.upload_scene(bad if bad else b"\0")
        assert e.value.status == -4
    T = binding.abi.parse_blob(bytes(blob))
    prims = T["prims"].copy()
    prims["material"][3] = 99
    T2 = dict(T)
    T2["prims"] = prims
    with pytest.raises(binding.RtbError) as e:
        ctx.upload_scene(binding.abi.build_blob(T2))
    assert e.value.status == -4 and "material" in str(e.value)
    ctx.upload_scene(bytes(blob))  # the context survives failed uploads
    acc, st = ctx.render(ctx.params(32, 32, 2, 1))
    assert st["paths"] == 32 * 32 * 2
    ctx.close()


def test_invalid_params(gpu_ctx, golden, binding):
    gpu_ctx.upload_scene(golden(7).blob)
    for kw in (dict(width=1), dict(integrator=7), dict(spp=-1), dict(sample_offset=2, sample_stride=2), dict(max_depth=65536)):
        args = dict(width=32, height=32, spp=1, integrator=1)
        args.update(kw)
        with pytest.raises(binding.RtbError) as e:
            gpu_ctx.render(gpu_ctx.params(**args))
        assert e.value.status == -1


def test_empty_and_ragged_batches(gpu_ctx, golden, binding):
    g = golden(23)
    gpu_ctx.upload_scene(g.blob)
    assert len(gpu_ctx.trace(np.zeros(0, binding.abi.RAY), 64)) == 0
    assert len(gpu_ctx.bsdf_eval(0, np.zeros(0, binding.abi.BSDF_QUERY), 32)) == 0
    for n in (1, 31, 33, 1000):  # not multiples of the warp / block size
        got = gpu_ctx.trace(g["rays"][:n], 64)
        assert np.array_equal(got["prim"], g["hits"]["prim"][:n])
    with pytest.raises(binding.RtbError):
        gpu_ctx.bsdf_eval(99, g["bsdf_q_0"], 64)
    with pytest.raises(binding.RtbError):
        gpu_ctx.trace(g["rays"][:4], 16)
    # degenerate rays: zero direction and NaN origins must not hang or crash
    r = g["rays"][:4].copy()
    r["d"][0] = 0
    r["o"][1] = np.nan
    r["t_max"][2] = -1
    out = gpu_ctx.trace(r, 32)
    assert out["prim"][2] == -1


def test_cancel_from_another_thread(gpu_ctx, golden, binding):
    gpu_ctx.upload_scene(golden(7).blob)
    p = gpu_ctx.params(600, 600, 60000, 0, seed=1)   # 21.6 G samples: seconds of work even at > 10 Gpaths/s
    res = {}

    def run():
        try:
            gpu_ctx.render(p)
            res["status"] = 0
        except binding.RtbError as e:
            res["status"] = e.status
    t = threading.Thread(target=run)
    t.start()
    time.sleep(0.1)
    gpu_ctx.cancel()
    t.join(60)
    assert res.get("status") == -6
    acc, st = gpu_ctx.render(gpu_ctx.params(64, 64, 4, 1))   # usable afterwards
    assert st["paths"] == 64 * 64 * 4


def test_fp64_tables_built_on_first_use(binding, golden):
    """RTB_OPT_LAZY_F64_PRIMS: a scene uploaded without its fp64 validation tables (what the 1 M-sphere
    scene gets by default) renders, and its first precision-64 call builds them: same bits as the
    tables built at upload."""
    import parity
    g = golden(9)
    ctx = binding.Context(0)
    try:
        ctx.set_option(binding.OPT_LAZY_F64_PRIMS, 0)
        ctx.upload_scene(g.blob)
        before = ctx.scene_stats()["device_bytes"]
        acc, st = ctx.render(ctx.params(32, 32, 4, 1, seed=3))
        assert np.isfinite(acc).all() and st["paths"] == 32 * 32 * 4
        got = ctx.trace(g["rays"], 64)
        assert ctx.scene_stats()["device_bytes"] > before
        ctx.set_option(binding.OPT_LAZY_F64_PRIMS, 1 << 40)
        ctx.upload_scene(g.blob)
        eager = ctx.trace(g["rays"], 64)
        assert np.array_equal(got["t"], eager["t"]) and np.array_equal(got["prim"], eager["prim"])
        tex = np.array([[0.3, 0.6, 1.0, 2.0, 3.0]])
        ctx.set_option(binding.OPT_LAZY_F64_PRIMS, 0)
        ctx.upload_scene(g.blob)
        lazy_tex = ctx.texture_eval(1, tex, 64)
        ctx.set_option(binding.OPT_LAZY_F64_PRIMS, 1 << 40)
        ctx.upload_scene(g.blob)
        assert np.array_equal(lazy_tex, ctx.texture_eval(1, tex, 64))
    finally:
        ctx.close()
