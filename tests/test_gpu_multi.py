"""Multi-GPU rendering inside the library (csrc/rtb_multi.cu): rtb_group_* (one process, several
GPUs, NCCL bound at run time) against the one-GPU render of the same job.  The two-GPU cases need a
box with at least two devices (`gpurun --gpus 2`); on a one-GPU box they are skipped and the
group-of-one cases still run."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


def test_group_of_one_is_the_single_context_render(binding, golden):
    blob = golden(21).blob
    with binding.Context(0) as ctx, binding.Group([0]) as grp:
        ctx.upload_scene(blob)
        grp.upload_scene(blob)
        p = ctx.params(96, 64, 32, 4, seed=5)
        a, sa = ctx.render(p)
        rgb_ref = ctx.resolve_rgb8(96, 64, 32)
        rgb = np.zeros((64, 96, 3), np.uint8)
        b, sb = grp.render(p, rgb8=rgb)
        assert sa["paths"] == sb["paths"] == 96 * 64 * 32
        assert np.allclose(a, b, rtol=1e-4, atol=1e-4)
        assert (np.abs(rgb.astype(int) - rgb_ref.astype(int)) <= 1).all()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("n", [2])
def test_group_render_sums_to_the_one_gpu_image(binding, golden, n):
    """Every GPU renders its slice of the samples (the stream of a sample depends only on pixel,
    sample index and seed), ncclReduce adds the staged means: the group's image is the one-GPU image
    up to float summation order, whatever the split."""
    blob = golden(21).blob
    with binding.Context(0) as ctx, binding.Group(list(range(n))) as grp:
        assert grp.size() == n
        ctx.upload_scene(blob)
        grp.upload_scene(blob)
        for spp, integ in ((32, 4), (1, 1)):     # spp < GPUs: the rows are split instead
            p = ctx.params(96, 64, spp, integ, seed=9)
            a, sa = ctx.render(p)
            rgb = np.zeros((64, 96, 3), np.uint8)
            b, sb = grp.render(p, rgb8=rgb)
            assert sb["paths"] == sa["paths"] == 96 * 64 * spp
            assert sb["rays_closest"] == sa["rays_closest"]
            assert np.allclose(a[..., :3], b[..., :3], rtol=2e-3, atol=2e-3 * spp)
            ref8 = np.floor(np.clip(np.sqrt(np.flipud(a[..., :3]) / spp), 0, 1) * 255).astype(int)
            assert (np.abs(rgb.astype(int) - ref8) <= 1).mean() > 0.999


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_cpp_renderer_on_two_gpus_without_python(tmp_path):
    """Renderer::render of the C++ host layer with set_devices({0, 1}): compiled and run as a plain
    C++ program (no Python, no torch in the process)."""
    pkg = os.path.join(ROOT, "ray_tracing-rendering_b200")
    exe = str(tmp_path / "host_group_test")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(pkg, "host"),
                           os.path.join(ROOT, "tests", "host_group_test.cpp"), "-o", exe, "-L" + pkg, "-lrtb200",
                           "-Wl,-rpath," + pkg])
    out = subprocess.run([exe, "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    vals = dict(l.split(" ", 1) for l in out.stdout.splitlines() if " " in l)
    assert vals["devices"].strip() == "2"
    one = np.array(vals["mean1"].split(), float)
    two = np.array(vals["mean2"].split(), float)
    assert np.allclose(one, two, rtol=0.02)
    assert int(vals["paths2"]) == 128 * 128 * 64
