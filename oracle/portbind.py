"""ctypes binding of oracle/liboracle_port.so — the CPU restatement (oracle/port).  TEST
INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs when the compiled reference (oracle/_ref) is not available."""
import ctypes as C
import importlib
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_port.so")
abi = importlib.import_module("ray_tracing-rendering_b200.abi")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, u64, i32, u32 = C.c_void_p, C.c_uint64, C.c_int, C.c_uint32
        L.port_scene_create.restype = vp
        L.port_scene_create.argtypes = [vp, u64]
        L.port_scene_destroy.argtypes = [vp]
        L.port_camera_derived.argtypes = [vp, vp]
        L.port_trace_batch.argtypes = [vp, vp, u64, u32, vp]
        L.port_bsdf_eval.argtypes = [vp, i32, vp, u64, vp]
        L.port_light_eval.argtypes = [vp, i32, vp, u64, u32, vp]
        L.port_texture_value.argtypes = [vp, i32, vp, u64, vp]
        L.port_render_linear.restype = C.c_double
        L.port_render_linear.argtypes = [vp, i32, i32, i32, i32, i32, i32, u32, vp, vp, vp]
        L.port_hardware_threads.restype = i32
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PortScene:
    def __init__(self, blob: bytes):
        self._buf = C.create_string_buffer(blob, len(blob))
        self.h = lib().port_scene_create(C.cast(self._buf, C.c_void_p), len(blob))
        if not self.h:
            raise RuntimeError("port_scene_create failed")

    def close(self):
        if self.h:
            lib().port_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def camera_derived(self):
        out = np.zeros(24)
        lib().port_camera_derived(self.h, _ptr(out))
        return out

    def trace(self, rays, seed=1):
        rays = np.ascontiguousarray(rays, dtype=abi.RAY)
        hits = np.zeros(rays.size, abi.HIT)
        lib().port_trace_batch(self.h, _ptr(rays), rays.size, seed, _ptr(hits))
        return hits

    def bsdf_eval(self, material, queries):
        q = np.ascontiguousarray(queries, dtype=abi.BSDF_QUERY)
        out = np.zeros(q.size, abi.BSDF_VALUE)
        lib().port_bsdf_eval(self.h, material, _ptr(q), q.size, _ptr(out))
        return out

    def light_eval(self, light, queries, seed=1):
        q = np.ascontiguousarray(queries, dtype=abi.LIGHT_QUERY)
        out = np.zeros(q.size, abi.LIGHT_VALUE)
        lib().port_light_eval(self.h, light, _ptr(q), q.size, seed, _ptr(out))
        return out

    def texture_value(self, texture, uvp):
        uvp = np.ascontiguousarray(uvp, dtype=np.float64).reshape(-1, 5)
        out = np.zeros((uvp.shape[0], 3))
        lib().port_texture_value(self.h, texture, _ptr(uvp), uvp.shape[0], _ptr(out))
        return out

    def render_linear(self, integrator, width, height, spp, max_depth=50, threads=0, seed=1, want_images=True):
        s = np.zeros((height, width, 3)) if want_images else None
        s2 = np.zeros((height, width, 3)) if want_images else None
        counters = np.zeros(2, np.uint64)
        secs = lib().port_render_linear(self.h, integrator, width, height, spp, max_depth, threads, seed,
                                        _ptr(s) if want_images else None, _ptr(s2) if want_images else None,
                                        _ptr(counters))
        return s, s2, counters, secs


def hardware_threads():
    return lib().port_hardware_threads()
