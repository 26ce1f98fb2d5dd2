"""GPU parity layer 3 (images) and size-independent properties of the wavefront renderer."""
import numpy as np
import pytest

import parity
from conftest import CATALOGUE_CASES

pytestmark = pytest.mark.gpu

# (scene, integrator): the BASELINE.json configs at fixture resolution + the feature tail
IMAGE_CASES = [(7, 0), (7, 1), (21, 3), (21, 4), (23, 2), (23, 3), (23, 4), (9, 1), (1, 1), (19, 3), (19, 4),
               (26, 4), (24, 4), (15, 3), (17, 4), (18, 3), (8, 1)] + CATALOGUE_CASES


@pytest.mark.parametrize("sid,integrator", IMAGE_CASES)
def test_image_statistics_match_reference(gpu_ctx, golden, sid, integrator):
    g = golden(sid)
    gpu_ctx.upload_scene(g.blob)
    ref_sum, ref_sumsq = g[f"img_{integrator}_sum"], g[f"img_{integrator}_sumsq"]
    ref_spp = int(g[f"img_{integrator}_spp"][0])
    h, w, _ = ref_sum.shape
    k, spp = 8, max(ref_spp // 2, 64)
    means = []
    for i in range(k):
        acc, st = gpu_ctx.render(gpu_ctx.params(w, h, spp, integrator, seed=100 + i))
        assert st["paths"] == w * h * spp
        assert np.isfinite(acc).all()
        means.append(acc[..., :3] / spp)
    rep = parity.image_report(ref_sum, ref_sumsq, ref_spp, np.stack(means))
    cal = parity.selfcal(sid, integrator)
    assert parity.image_gates(rep, cal is not None) == [], rep
    if cal is not None:
        # the contract's gates, against the reference-vs-itself calibration of this very case
        # (tests/golden/selfcal.npz): 4 renders at the reference's sample count, one at 32x
        equal = [gpu_ctx.render(gpu_ctx.params(w, h, ref_spp, integrator, seed=300 + i))[0][..., :3] / ref_spp for i in range(4)]
        hi = gpu_ctx.render(gpu_ctx.params(w, h, 32 * ref_spp, integrator, seed=77))[0][..., :3] / (32 * ref_spp)
        bad = parity.strict_image_gates(ref_sum, ref_sumsq, ref_spp, np.stack(equal), hi.mean(axis=(0, 1)), cal)
        assert bad == [], (bad, cal)


@pytest.mark.parametrize("mode", [1, 2])
def test_both_traversal_kernels_render_scene09(gpu_ctx, golden, binding, mode):
    """BVH scenes are traced by the warp-scheduled 4-wide kernels (csrc/rtb_trace.cuh) or by round 1's
    binary while-while kernels; the default picks by scene (scene09, with media and an instance in
    its tree, gets the binary ones).  Both, forced, must pass the image gates and the path-length
    check on the scene that uses every primitive type."""
    g = golden(9)
    gpu_ctx.set_option(binding.OPT_BINARY_TRAVERSAL, mode)
    try:
        gpu_ctx.upload_scene(g.blob)
        ref_sum, ref_sumsq, ref_spp = g["img_1_sum"], g["img_1_sumsq"], int(g["img_1_spp"][0])
        h, w, _ = ref_sum.shape
        spp = max(ref_spp // 2, 64)
        means, rays, paths = [], 0, 0
        for i in range(8):
            acc, st = gpu_ctx.render(gpu_ctx.params(w, h, spp, 1, seed=200 + i))
            assert st["paths"] == w * h * spp and np.isfinite(acc).all()
            means.append(acc[..., :3] / spp)
            rays += st["rays_closest"]
            paths += st["paths"]
        rep = parity.image_report(ref_sum, ref_sumsq, ref_spp, np.stack(means))
        assert parity.image_gates(rep) == [], rep
        ref_rpp = float(g["img_1_rays"][0]) / (w * h * ref_spp)
        assert abs(rays / paths - ref_rpp) <= 0.01 * ref_rpp
    finally:
        gpu_ctx.set_option(binding.OPT_BINARY_TRAVERSAL, 0)


def test_rays_per_path_match_reference(gpu_ctx, golden):
    """Path lengths are a sensitive whole-pipeline statistic (3.25 closest-hit
    rays per path on scene07/int1)."""
    for sid, integ in ((7, 1), (7, 0), (9, 1), (1, 1), (8, 1)):
        g = golden(sid)
        gpu_ctx.upload_scene(g.blob)
        ref_sum = g[f"img_{integ}_sum"]
        h, w, _ = ref_sum.shape
        spp = 256
        _, st = gpu_ctx.render(gpu_ctx.params(w, h, spp, integ, seed=3))
        ref_rpp = float(g[f"img_{integ}_rays"][0]) / (w * h * int(g[f"img_{integ}_spp"][0]))
        assert abs(st["rays_closest"] / st["paths"] - ref_rpp) <= 0.01 * ref_rpp, (sid, integ)


def test_sample_split_sums_to_the_whole(gpu_ctx, golden):
    """Multi-GPU contract: renders with sample_stride N / offsets 0..N-1 add up to the
    stride-1 image (same per-sample RNG streams); only the float summation order differs."""
    g = golden(21)
    gpu_ctx.upload_scene(g.blob)
    w = h = 128
    spp = 24
    whole, st = gpu_ctx.render(gpu_ctx.params(w, h, spp, 4, seed=9))
    for n in (2, 3, 8):
        parts = np.zeros_like(whole)
        rays = 0
        for r in range(n):
            acc, s = gpu_ctx.render(gpu_ctx.params(w, h, spp, 4, seed=9, sample_offset=r, sample_stride=n))
            parts += acc
            rays += s["rays_closest"] + s["rays_shadow"]
        assert rays == st["rays_closest"] + st["rays_shadow"]   # identical paths, exactly
        assert np.allclose(parts[..., :3], whole[..., :3], rtol=2e-4, atol=1e-3)


def test_result_is_independent_of_pool_size_and_repeatable(gpu_ctx, golden):
    g = golden(7)
    gpu_ctx.upload_scene(g.blob)
    w = h = 200
    base, st0 = gpu_ctx.render(gpu_ctx.params(w, h, 32, 1, seed=4, pool_paths=1 << 20))
    for pool in (1 << 12, 1 << 16, 1 << 20):
        acc, st = gpu_ctx.render(gpu_ctx.params(w, h, 32, 1, seed=4, pool_paths=pool))
        assert st["rays_closest"] == st0["rays_closest"] and st["paths"] == st0["paths"]
        assert np.allclose(acc[..., :3], base[..., :3], rtol=2e-4, atol=1e-3)
    other, _ = gpu_ctx.render(gpu_ctx.params(w, h, 32, 1, seed=5))
    assert not np.allclose(other[..., :3], base[..., :3], rtol=1e-3, atol=1e-3)


def test_full_size_c1_properties(gpu_ctx, golden):
    """BASELINE config C1 at full size (600x600, 400 spp, depth 50, integrator 1): too large for
    the CPU oracle in a test, so checked through properties: Russian roulette is unbiased
    (integrators 0 and 1 agree), every sample is accounted for, the image is finite, and the
    whole-image mean equals the reference's mean at the same resolution
    (tests/golden/fullres_means.npz: 600x600, 48 spp; the mean depends on the resolution
    through u = (i+xi)/(W-1), renderer.h:73-74)."""
    import os
    from conftest import GOLDEN
    full = np.load(os.path.join(GOLDEN, "fullres_means.npz"))
    g = golden(7)
    gpu_ctx.upload_scene(g.blob)
    acc1, st1 = gpu_ctx.render(gpu_ctx.params(600, 600, 400, 1, seed=1))
    assert st1["paths"] == 600 * 600 * 400 and np.isfinite(acc1).all()
    acc0, st0 = gpu_ctx.render(gpu_ctx.params(600, 600, 100, 0, seed=2))
    m1 = acc1[..., :3].mean(axis=(0, 1)) / 400
    m0 = acc0[..., :3].mean(axis=(0, 1)) / 100
    assert np.allclose(m0, m1, rtol=0.01)
    ref = full["mean_7_1"][:3]
    assert np.allclose(m1, ref, rtol=0.01), (m1, ref)
    assert np.allclose(m0, full["mean_7_0"][:3], rtol=0.01)
    assert 3.2 < st1["rays_closest"] / st1["paths"] < 3.3    # 3.25-3.26
    assert 6.4 < st0["rays_closest"] / st0["paths"] < 6.7    # 6.55-6.58


@pytest.mark.parametrize("sid,integrator", [(21, 3), (21, 4), (23, 3), (23, 4)])
def test_full_size_config_means(gpu_ctx, golden, sid, integrator):
    """BASELINE configs C3 / C4 at their full resolution and spp: whole-image mean and rays per
    path against the reference's at the same resolution."""
    import os
    from conftest import GOLDEN
    full = np.load(os.path.join(GOLDEN, "fullres_means.npz"))
    g = golden(sid)
    gpu_ctx.upload_scene(g.blob)
    w, h, spp = (600, 600, 400) if sid == 21 else (800, 450, 64)
    acc, st = gpu_ctx.render(gpu_ctx.params(w, h, spp, integrator, seed=1))
    m = acc[..., :3].mean(axis=(0, 1)) / spp
    ref = full[f"mean_{sid}_{integrator}"]
    assert np.allclose(m, ref[:3], rtol=0.01), (m, ref)
    assert abs(st["rays_closest"] / st["paths"] - full[f"rays_{sid}_{integrator}"][0]) < 0.02


def test_depth_limit_and_zero_work(gpu_ctx, golden):
    g = golden(7)
    gpu_ctx.upload_scene(g.blob)
    acc, st = gpu_ctx.render(gpu_ctx.params(64, 64, 16, 1, max_depth=1))
    assert st["rays_closest"] == st["paths"] == 64 * 64 * 16       # exactly one ray per path
    acc, st = gpu_ctx.render(gpu_ctx.params(64, 64, 0, 1))
    assert st["paths"] == 0 and not acc.any()
    acc, st = gpu_ctx.render(gpu_ctx.params(64, 64, 4, 1, max_depth=0))
    assert st["paths"] == 64 * 64 * 4 and not acc[..., :3].any()


def test_resolve_rgb8_is_the_reference_output_path(gpu_ctx, golden, tmp_path):
    """the device's resolve against the reference's OWN output path, not a formula — the
    sums of a GPU render go through the reference's Renderer::write_color_to_buffer (renderer.h:126-140) and
    RenderBuffer::save_to_png (render_buffer.h:35-55, stb's PNG encoder) inside oracle/_ref; the decoded
    file must equal rtb_resolve_rgb8's bytes exactly (sqrt in double, clamp, truncation, y flip), on an
    image with clamped highlights, black pixels and every byte value in between."""
    from PIL import Image
    from oracle import refbind
    if not refbind.available():
        pytest.fail("oracle/_ref/libref_oracle.so did not travel to this box")
    g = golden(23)
    gpu_ctx.upload_scene(g.blob)
    for (w, h, spp, integ) in ((160, 90, 32, 4), (97, 31, 5, 3)):
        acc, _ = gpu_ctx.render(gpu_ctx.params(w, h, spp, integ, seed=2))
        got = gpu_ctx.resolve_rgb8(w, h, spp)
        png = str(tmp_path / f"ref_{w}.png")
        buf = refbind.output_path(acc[..., :3], spp, png)
        want = np.asarray(Image.open(png).convert("RGB"))
        assert want.shape == got.shape == (h, w, 3)
        assert np.array_equal(got, want), f"{(got != want).sum()} bytes differ"
        assert (want == 255).any() and len(np.unique(want)) > 200
        # and the RenderBuffer itself (what the reference's window polls, main.cpp:119-126)
        assert np.array_equal(buf, np.clip(np.sqrt((1.0 / spp) * acc[..., :3].astype(np.float64)), 0, 1))


@pytest.mark.parametrize("sid,integrator", [(7, 1), (21, 4), (23, 4), (23, 3), (8, 1)])
def test_fused_and_wavefront_schedules_agree(gpu_ctx, golden, binding, sid, integrator):
    """Small scenes can run the fused persistent kernel; the wavefront schedule is the
    same stage functions behind HBM queues.  Both consume the per-sample RNG stream in the
    same order, so without media they trace exactly the same paths."""
    g = golden(sid)
    gpu_ctx.upload_scene(g.blob)
    w, h, spp = 96, 96, 64
    a, sa = gpu_ctx.render(gpu_ctx.params(w, h, spp, integrator, seed=21, flags=binding.RENDER_FORCE_FUSED))
    b, sb = gpu_ctx.render(gpu_ctx.params(w, h, spp, integrator, seed=21, flags=binding.RENDER_FORCE_WAVEFRONT))
    assert sa["schedule"] == 1 and sb["schedule"] == 0
    assert sa["paths"] == sb["paths"] == w * h * spp
    if sid == 23:  # no boxes, no media: identical paths, ray for ray
        assert sa["rays_closest"] == sb["rays_closest"] and sa["rays_shadow"] == sb["rays_shadow"]
        assert np.allclose(a[..., :3], b[..., :3], rtol=2e-4, atol=1e-3)
    elif sid != 8:
        # The fused kernel tests a `box` as one slab test, the wavefront as six rects: same t for
        # the same face, but a ray within rounding distance of a box edge may pick the other face,
        # after which that one path differs.  Almost every pixel is still identical.
        assert abs(sa["rays_closest"] - sb["rays_closest"]) <= 2e-4 * sb["rays_closest"]
        same = np.isclose(a[..., :3], b[..., :3], rtol=2e-4, atol=1e-3).all(axis=-1)
        assert same.mean() > 0.998
        assert np.allclose(a[..., :3].mean(axis=(0, 1)), b[..., :3].mean(axis=(0, 1)), rtol=2e-3)
    else:  # cornell_smoke: shadow/extension rays through media draw different random numbers
        assert abs(sa["rays_closest"] - sb["rays_closest"]) < 0.01 * sb["rays_closest"]
        assert np.allclose(a[..., :3].mean(axis=(0, 1)), b[..., :3].mean(axis=(0, 1)), rtol=0.02)


def test_large_scenes_use_the_wavefront_schedule(gpu_ctx, golden):
    gpu_ctx.upload_scene(golden(9).blob)
    _, st = gpu_ctx.render(gpu_ctx.params(64, 64, 8, 1))
    assert st["schedule"] == 0 and st["iterations"] > 4


@pytest.mark.gpu
def test_wavefront_result_does_not_depend_on_the_pool_size(gpu_ctx, golden, binding):
    """A sample's random stream is keyed by (pixel, sample), so the set of paths — and, without
    media, every ray — is the same however many paths are resident: tiny pools exercise the
    refill of empty queue entries and the final drain of the warps' private sample ranges."""
    gpu_ctx.upload_scene(golden(1).blob)          # random spheres: BVH scene, no media
    w, h, spp = 80, 45, 24
    ref = None
    for pool in (0, 1000, 4096, 65536):
        acc, st = gpu_ctx.render(gpu_ctx.params(w, h, spp, 1, seed=5, pool_paths=pool,
                                                flags=binding.RENDER_FORCE_WAVEFRONT))
        assert st["schedule"] == 0 and st["paths"] == w * h * spp
        assert np.isfinite(acc).all()
        if ref is None:
            ref = (acc.copy(), st)
        else:
            assert st["rays_closest"] == ref[1]["rays_closest"]
            assert np.allclose(acc[..., :3], ref[0][..., :3], rtol=2e-4, atol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,spp", [(7, 5, 3), (2, 2, 1), (33, 2, 2), (64, 64, 1)])
def test_wavefront_small_and_ragged_jobs(gpu_ctx, golden, binding, w, h, spp):
    """Fewer samples than one warp / than the pool, sample counts that are no multiple of 32."""
    gpu_ctx.upload_scene(golden(1).blob)
    for integrator in (1, 4):
        acc, st = gpu_ctx.render(gpu_ctx.params(w, h, spp, integrator, seed=3, flags=binding.RENDER_FORCE_WAVEFRONT))
        assert st["paths"] == w * h * spp and st["rays_closest"] >= w * h * spp
        assert np.isfinite(acc).all() and (acc[..., :3] >= 0).all()
    # max_depth 0: nothing is traced
    acc, st = gpu_ctx.render(gpu_ctx.params(w, h, spp, 1, max_depth=0, flags=binding.RENDER_FORCE_WAVEFRONT))
    assert st["rays_closest"] == 0 and not acc[..., :3].any()


@pytest.mark.gpu
def test_no_degenerate_rays_survive(gpu_ctx, golden, binding):
    """fp32 guard: a zero-length scatter direction (possible once per ~2^24 isotropic draws) used
    to turn into a NaN ray that passed every slab test and walked the whole tree for up to 50
    bounces.  With media in the scene, no ray may visit a large part of the tree and the image
    must be finite."""
    gpu_ctx.upload_scene(golden(9).blob)
    n_nodes = gpu_ctx.scene_stats()["n_nodes"]
    worst = 0
    for seed in range(1, 5):
        acc, st = gpu_ctx.render(gpu_ctx.params(400, 400, 24, 1, seed=seed, flags=binding.RENDER_COUNT_VISITS))
        assert np.isfinite(acc).all()
        worst = max(worst, st["max_nodes_per_ray"])
    assert 0 < worst < n_nodes // 4, (worst, n_nodes)


@pytest.mark.gpu
def test_row_split_tiles_the_image(gpu_ctx, golden):
    """Image split of the multi-GPU driver: the calls with row_stride N / offsets 0..N-1 write
    disjoint rows, and because a sample's stream is keyed by the IMAGE pixel their sum is the
    full image exactly (one contribution order per pixel: no float reordering at all for the
    closest-hit only integrators is not guaranteed by atomics, so a tolerance is kept)."""
    gpu_ctx.upload_scene(golden(21).blob)
    w, h, spp = 96, 50, 16
    whole, st = gpu_ctx.render(gpu_ctx.params(w, h, spp, 4, seed=4))
    for n in (2, 3, 7):
        parts = np.zeros_like(whole)
        rays = 0
        for r in range(n):
            acc, s = gpu_ctx.render(gpu_ctx.params(w, h, spp, 4, seed=4, row_offset=r, row_stride=n))
            others = np.ones(h, bool)
            others[r::n] = False
            assert not acc[others].any()                      # rows of other ranks stay zero
            assert s["paths"] == w * len(range(r, h, n)) * spp
            parts += acc
            rays += s["rays_closest"] + s["rays_shadow"]
        assert rays == st["rays_closest"] + st["rays_shadow"]
        assert np.allclose(parts[..., :3], whole[..., :3], rtol=2e-4, atol=1e-3)
    # rows x samples together (2 x 2 ranks)
    parts = np.zeros_like(whole)
    for r in range(2):
        for q in range(2):
            parts += gpu_ctx.render(gpu_ctx.params(w, h, spp, 4, seed=4, row_offset=r, row_stride=2,
                                                   sample_offset=q, sample_stride=2))[0]
    assert np.allclose(parts[..., :3], whole[..., :3], rtol=2e-4, atol=1e-3)
