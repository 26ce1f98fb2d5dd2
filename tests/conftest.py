import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "ray_tracing-rendering_b200"
GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_SCENES = [7, 21, 23, 9, 1, 19, 26, 24, 15, 17, 18, 8]
# the rest of the reference's catalogue (scenes.cpp:1523-2096) as smaller fixtures, generated with synthetic
# image textures / normal maps / .hdr files in place (tests/golden/make_golden.py CATALOGUE): (scene, integrator)
CATALOGUE_CASES = [(2, 1), (4, 1), (5, 1), (6, 0), (10, 1), (11, 2), (12, 2), (13, 2), (14, 2), (16, 4), (20, 3), (22, 4),
                   (25, 4), (27, 4), (28, 3), (30, 4), (31, 4), (32, 4), (33, 4), (34, 4), (35, 4), (36, 4), (37, 4),
                   (38, 3), (39, 4), (40, 4), (41, 3), (42, 1)]
CATALOGUE_SCENES = [c[0] for c in CATALOGUE_CASES]
ALL_SCENES = GOLDEN_SCENES + CATALOGUE_SCENES


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def abi():
    return importlib.import_module(PKG + ".abi")


@pytest.fixture(scope="session")
def binding():
    return importlib.import_module(PKG + ".binding")


class Golden:
    def __init__(self, sid):
        self.sid = sid
        self.z = np.load(os.path.join(GOLDEN, f"scene{sid:02d}.npz"))
        self.blob = self.z["blob"].tobytes()

    def __getitem__(self, k):
        return self.z[k]

    def keys(self, prefix):
        return sorted(int(k[len(prefix):]) for k in self.z.files if k.startswith(prefix))


_golden_cache = {}


def load_golden(sid):
    if sid not in _golden_cache:
        _golden_cache[sid] = Golden(sid)
    return _golden_cache[sid]


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(scope="session")
def gpu_ctx(binding):
    """A context on cuda:0.  The product has no CPU path: this fixture fails (it does
    not skip) when the library is missing or no device is present."""
    ctx = binding.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def hostcheck():
    """tests/hostcheck/libhostcheck.so: the device headers instantiated with g++ (a
    development aid for the GPU-less container, never part of the product)."""
    import ctypes as C
    src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp")
    so = os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so")
    csrc = os.path.join(ROOT, PKG, "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".hpp"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                               "-I" + os.path.join(ROOT, "include"), "-I" + csrc, src, "-o", so])
    L = C.CDLL(so)
    L.hc_scene_create.restype = C.c_void_p
    L.hc_scene_create.argtypes = [C.c_char_p, C.c_uint64, C.c_int]
    L.hc_scene_destroy.argtypes = [C.c_void_p]
    L.hc_scene_info.argtypes = [C.c_void_p, C.c_void_p]
    L.hc_scene_tree_hash.restype = C.c_uint64
    L.hc_scene_tree_hash.argtypes = [C.c_void_p]
    L.hc_trace_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    L.hc_bsdf_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    L.hc_bsdf_sample.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_void_p]
    L.hc_light_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_void_p]
    L.hc_camera_derived.argtypes = [C.c_void_p, C.c_void_p]
    L.hc_wide_info.argtypes = [C.c_void_p, C.c_void_p]
    L.hc_sched_stats.argtypes = [C.c_void_p]
    L.hc_qnode_check.argtypes = [C.c_void_p, C.c_void_p]
    return L
