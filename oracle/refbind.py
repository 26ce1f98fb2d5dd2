"""ctypes binding of oracle/_ref/libref_oracle.so (the unmodified reference,
compiled headless by oracle/Makefile).  TEST INFRASTRUCTURE ONLY: imported by
tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl
reference legs — never by the product package."""
import ctypes as C
import importlib
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_oracle.so")
abi = importlib.import_module("ray_tracing-rendering_b200.abi")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_scene_create.restype = C.c_void_p
        L.ref_scene_create.argtypes = [C.c_int]
        L.ref_scene_destroy.argtypes = [C.c_void_p]
        L.ref_scene_blob.restype = C.c_void_p
        L.ref_scene_blob.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.ref_camera_derived.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_trace_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_record_rays.restype = C.c_uint64
        L.ref_record_rays.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
        L.ref_bsdf_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_bsdf_sample.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_texture_value.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_light_eval.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_light_flags.restype = C.c_int
        L.ref_light_flags.argtypes = [C.c_void_p, C.c_int]
        L.ref_render_linear.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_render_timed.restype = C.c_double
        L.ref_render_timed.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p,
                                       C.c_uint64]
        L.ref_hardware_threads.restype = C.c_int
        L.ref_output_path.restype = C.c_int
        L.ref_output_path.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class RefScene:
    """select_scene(scene_id) of the reference, walked and tagged."""

    def __init__(self, scene_id: int):
        self.h = lib().ref_scene_create(int(scene_id))
        if not self.h:
            raise RuntimeError(f"ref_scene_create({scene_id}) failed")
        self.scene_id = scene_id

    def close(self):
        if self.h:
            lib().ref_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def blob(self) -> bytes:
        n = C.c_uint64()
        p = lib().ref_scene_blob(self.h, C.byref(n))
        return C.string_at(p, n.value)

    def camera_derived(self):
        out = np.zeros(24)
        lib().ref_camera_derived(self.h, _ptr(out))
        return out

    def trace(self, rays):
        rays = np.ascontiguousarray(rays, dtype=abi.RAY)
        hits = np.zeros(rays.size, abi.HIT)
        lib().ref_trace_batch(self.h, _ptr(rays), rays.size, _ptr(hits))
        return hits

    def record_rays(self, integrator, n_paths, max_rays):
        rays = np.zeros(max_rays, abi.RAY)
        hits = np.zeros(max_rays, abi.HIT)
        counters = np.zeros(2, np.uint64)
        n = lib().ref_record_rays(self.h, integrator, n_paths, max_rays, _ptr(rays), _ptr(hits),
                                  _ptr(counters))
        return rays[:n].copy(), hits[:n].copy(), counters

    def bsdf_eval(self, material, queries):
        q = np.ascontiguousarray(queries, dtype=abi.BSDF_QUERY)
        out = np.zeros(q.size, abi.BSDF_VALUE)
        lib().ref_bsdf_eval(self.h, material, _ptr(q), q.size, _ptr(out))
        return out

    def bsdf_sample(self, material, queries):
        q = np.ascontiguousarray(queries, dtype=abi.BSDF_QUERY)
        out = np.zeros(q.size, abi.BSDF_SAMPLE)
        lib().ref_bsdf_sample(self.h, material, _ptr(q), q.size, _ptr(out))
        return out

    def texture_value(self, texture, uvp):
        uvp = np.ascontiguousarray(uvp, dtype=np.float64).reshape(-1, 5)
        out = np.zeros((uvp.shape[0], 3))
        lib().ref_texture_value(self.h, texture, _ptr(uvp), uvp.shape[0], _ptr(out))
        return out

    def light_eval(self, light, queries):
        q = np.ascontiguousarray(queries, dtype=abi.LIGHT_QUERY)
        out = np.zeros(q.size, abi.LIGHT_VALUE)
        lib().ref_light_eval(self.h, light, _ptr(q), q.size, _ptr(out))
        return out

    def light_flags(self, light):
        return lib().ref_light_flags(self.h, light)

    def render_linear(self, integrator, width, height, spp, max_depth=50, threads=0):
        s = np.zeros((height, width, 3))
        s2 = np.zeros((height, width, 3))
        counters = np.zeros(2, np.uint64)
        lib().ref_render_linear(self.h, integrator, width, height, spp, max_depth, threads,
                                _ptr(s), _ptr(s2), _ptr(counters))
        return s, s2, counters


def render_timed(scene_id, integrator, width=0, spp=0, max_depth=50, want_image=False):
    """Unmodified Renderer::render on a fresh scene; returns (seconds, w, h, image|None)."""
    w, h = C.c_int(), C.c_int()
    img = None
    cap = 0
    if want_image:
        img = np.zeros(4096 * 4096 * 3)
        cap = img.size
    # Renderer::render prints "Rendering finished in ..." on the C++ stdout (renderer.h:100);
    # keep it off this process's stdout (bench.py prints exactly one JSON line there)
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        s = lib().ref_render_timed(scene_id, integrator, width, spp, max_depth, C.byref(w), C.byref(h),
                                   _ptr(img) if img is not None else None, cap)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)
    if img is not None:
        img = img[:w.value * h.value * 3].reshape(h.value, w.value, 3).copy()
    return s, w.value, h.value, img


def output_path(sums, samples, png_path):
    """The reference's own output path on caller-supplied linear sums (H x W x 3, row 0 = bottom of the
    image): Renderer::write_color_to_buffer (renderer.h:126-140) per pixel, then RenderBuffer::save_to_png
    (render_buffer.h:35-55) to png_path.  Returns the RenderBuffer's contents (H x W x 3, same layout)."""
    sums = np.ascontiguousarray(sums, dtype=np.float64)
    h, w, _ = sums.shape
    buf = np.zeros((h, w, 3))
    if not lib().ref_output_path(w, h, _ptr(sums), int(samples), os.fsencode(png_path), _ptr(buf)):
        raise RuntimeError("save_to_png failed")
    return buf


def hardware_threads():
    return lib().ref_hardware_threads()
