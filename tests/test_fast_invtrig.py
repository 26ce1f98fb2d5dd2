"""The fp32 acos_ / atan2_ of csrc/rtb_math.cuh (polynomial forms used by sphere u,v and the environment-map
look-ups, texture.h:19-23 / sphere.h:68-81 / environmental_light.h:250-356 call the C library on doubles) against
the exact functions of the same fp32 arguments."""
import ctypes as C

import numpy as np


def test_fast_acos_and_atan2_are_within_1e6_rad(hostcheck):
    rng = np.random.default_rng(3)
    n = 200_000
    x = np.concatenate([rng.uniform(-1, 1, n - 8), [-1, 1, 0, -0.0, 0.99999994, -0.99999994, 1e-20, -1e-20]]).astype(np.float32)
    y = np.concatenate([rng.uniform(-1, 1, n - 8), [0, 0, 0, 1, 1e-20, -1e-20, 1, -1]]).astype(np.float32)
    # a second batch with wildly different magnitudes (atan2 only sees the ratio)
    y[: n // 4] *= np.float32(1e-6)
    a = np.zeros(n, np.float32)
    t = np.zeros(n, np.float32)
    hostcheck.hc_invtrig.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    hostcheck.hc_invtrig(x.ctypes.data, y.ctypes.data, n, a.ctypes.data, t.ctypes.data)
    x64, y64 = x.astype(np.float64), y.astype(np.float64)
    ref_a = np.arccos(x64)
    assert np.all(np.abs(a - ref_a) <= 5e-7), np.abs(a - ref_a).max()      # (acosf itself: 3.3e-7 on these arguments)
    ref_t = np.arctan2(y64, x64)
    assert np.all(np.abs(t - ref_t) <= 8e-7), np.abs(t - ref_t).max()
    assert t[n - 8 + 2] == 0.0                                  # atan2(0, 0)
    assert abs(a[n - 8] - np.pi) < 1e-6 and a[n - 7] == 0.0     # acos(-1), acos(1)
