"""ctypes binding of librtb200.so (include/rtb200.h).

Plumbing only: every call goes straight through the C-ABI; there is no Python or
CPU implementation of anything behind it.  Loading fails loudly when the library
has not been built, and creating a context fails loudly without a CUDA device.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# RTB200_LIBRARY selects another build of the same library (kernel-tuning experiments)
LIB_PATH = os.environ.get("RTB200_LIBRARY") or os.path.join(_HERE, "librtb200.so")

RTB_OK = 0
STATUS_NAMES = {0: "RTB_OK", -1: "RTB_ERR_INVALID_ARGUMENT", -2: "RTB_ERR_NO_DEVICE",
                -3: "RTB_ERR_CUDA", -4: "RTB_ERR_BAD_SCENE", -5: "RTB_ERR_NO_SCENE",
                -6: "RTB_ERR_CANCELLED", -7: "RTB_ERR_OUT_OF_MEMORY"}
RENDER_COUNT_VISITS = 1
RENDER_TIME_EXTEND = 2
RENDER_FORCE_WAVEFRONT = 4
RENDER_FORCE_FUSED = 8
OPT_FLAT_TRAVERSAL, OPT_FUSED_SCHEDULE = 1, 2
OPT_BVH_MAX_LEAF, OPT_BVH_TRAVERSAL_COST_PCT, OPT_BVH_LAYOUT_DFS, OPT_BINARY_TRAVERSAL, OPT_LAZY_F64_PRIMS = 3, 4, 5, 6, 7
OPT_GROUP_BOXES = 8

# Every symbol include/rtb200.h declares (tests check the library exports them all).
EXPORTS = ["rtb_version", "rtb_context_create", "rtb_context_destroy", "rtb_last_error",
           "rtb_set_option", "rtb_scene_upload", "rtb_scene_get_stats", "rtb_camera_derived", "rtb_scene_env_tables", "rtb_render",
           "rtb_render_device", "rtb_cancel", "rtb_resolve_rgb8", "rtb_trace_batch",
           "rtb_bsdf_eval_batch", "rtb_bsdf_sample_batch", "rtb_light_eval_batch",
           "rtb_texture_eval_batch",
           "rtb_group_create", "rtb_group_destroy", "rtb_group_last_error", "rtb_group_size", "rtb_group_context",
           "rtb_group_scene_upload", "rtb_group_render", "rtb_group_cancel",
           "rtb_comm_unique_id", "rtb_comm_init", "rtb_comm_release", "rtb_render_reduce", "rtb_accum_copy_device"]
COMM_ID_BYTES = 128


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32),
                ("max_depth", C.c_int32), ("rr_start_depth", C.c_int32), ("integrator", C.c_int32),
                ("sample_offset", C.c_int32), ("sample_stride", C.c_int32), ("seed", C.c_uint64),
                ("pool_paths", C.c_int32), ("flags", C.c_int32), ("row_offset", C.c_int32),
                ("row_stride", C.c_int32)]


class RenderStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("prim_tests", C.c_uint64), ("iterations", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("device_ms", C.c_double), ("extend_ms", C.c_double),
                ("extend_launches", C.c_uint64), ("schedule", C.c_int32), ("traversal", C.c_int32), ("stage_ms", C.c_double * 4),
                ("max_nodes_per_ray", C.c_uint64), ("extend_nodes", C.c_uint64),
                ("extend_chunk_max_nodes", C.c_uint64)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["stage_ms"] = list(self.stage_ms)
        return d


class SceneStats(C.Structure):
    _fields_ = [("n_prims", C.c_int32), ("n_nodes", C.c_int32), ("n_instances", C.c_int32),
                ("n_materials", C.c_int32), ("n_lights", C.c_int32), ("has_media", C.c_int32),
                ("device_bytes", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class RtbError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


_lib = None


def load():
    """dlopen librtb200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                          "(there is no fallback implementation)")
    L = C.CDLL(LIB_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    L.rtb_version.restype = C.c_char_p
    L.rtb_context_create.argtypes = [i32, C.POINTER(vp)]
    L.rtb_context_destroy.argtypes = [vp]
    L.rtb_context_destroy.restype = None
    L.rtb_last_error.argtypes = [vp]
    L.rtb_last_error.restype = C.c_char_p
    L.rtb_set_option.argtypes = [vp, i32, C.c_int64]
    L.rtb_scene_env_tables.argtypes = [vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.rtb_scene_upload.argtypes = [vp, vp, u64]
    L.rtb_scene_get_stats.argtypes = [vp, C.POINTER(SceneStats)]
    L.rtb_camera_derived.argtypes = [vp, vp]
    L.rtb_render.argtypes = [vp, C.POINTER(RenderParams), vp, C.POINTER(RenderStats)]
    L.rtb_render_device.argtypes = [vp, C.POINTER(RenderParams), vp, vp, C.POINTER(RenderStats)]
    L.rtb_cancel.argtypes = [vp]
    L.rtb_resolve_rgb8.argtypes = [vp, C.c_int32, vp]
    L.rtb_trace_batch.argtypes = [vp, vp, u64, i32, vp, vp]
    L.rtb_bsdf_eval_batch.argtypes = [vp, i32, vp, u64, i32, vp]
    L.rtb_bsdf_sample_batch.argtypes = [vp, i32, vp, u64, i32, u64, vp]
    L.rtb_light_eval_batch.argtypes = [vp, i32, vp, u64, i32, u64, vp]
    L.rtb_texture_eval_batch.argtypes = [vp, i32, vp, u64, i32, vp]
    L.rtb_group_create.argtypes = [vp, i32, C.POINTER(vp)]
    L.rtb_group_destroy.argtypes = [vp]
    L.rtb_group_destroy.restype = None
    L.rtb_group_last_error.argtypes = [vp]
    L.rtb_group_last_error.restype = C.c_char_p
    L.rtb_group_size.argtypes = [vp]
    L.rtb_group_context.argtypes = [vp, i32]
    L.rtb_group_context.restype = vp
    L.rtb_group_scene_upload.argtypes = [vp, vp, u64]
    L.rtb_group_render.argtypes = [vp, C.POINTER(RenderParams), vp, vp, C.POINTER(RenderStats)]
    L.rtb_group_cancel.argtypes = [vp]
    L.rtb_comm_unique_id.argtypes = [vp]
    L.rtb_comm_init.argtypes = [vp, i32, i32, vp]
    L.rtb_comm_release.argtypes = [vp]
    L.rtb_comm_release.restype = None
    L.rtb_render_reduce.argtypes = [vp, C.POINTER(RenderParams), vp, vp, vp, C.POINTER(RenderStats)]
    L.rtb_accum_copy_device.argtypes = [vp, vp, vp]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One rtb_context (one GPU)."""

    def __init__(self, device: int = 0):
        self._lib = load()
        h = C.c_void_p()
        rc = self._lib.rtb_context_create(int(device), C.byref(h))
        if rc != RTB_OK:
            raise RtbError(rc, self._lib.rtb_last_error(None).decode())
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rtb_context_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != RTB_OK:
            raise RtbError(rc, self._lib.rtb_last_error(self._h).decode())

    def set_option(self, option: int, value: int):
        self._check(self._lib.rtb_set_option(self._h, option, value))

    # ---- scene
    def upload_scene(self, blob: bytes):
        # the library copies what it needs: hand it the bytes object's own buffer (no 200 MB staging copy)
        blob = bytes(blob) if not isinstance(blob, bytes) else blob
        self._check(self._lib.rtb_scene_upload(self._h, C.cast(C.c_char_p(blob), C.c_void_p), len(blob)))

    def scene_stats(self) -> dict:
        s = SceneStats()
        self._check(self._lib.rtb_scene_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def env_tables(self):
        n = C.c_uint64()
        self._check(self._lib.rtb_scene_env_tables(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value)
        if n.value:
            self._check(self._lib.rtb_scene_env_tables(self._h, _ptr(out), n.value, None))
        return out

    def camera_derived(self):
        out = np.zeros(24)
        self._check(self._lib.rtb_camera_derived(self._h, _ptr(out)))
        return out

    # ---- render
    @staticmethod
    def params(width, height, spp, integrator, max_depth=50, rr_start_depth=3, seed=1,
               sample_offset=0, sample_stride=1, pool_paths=0, flags=0, row_offset=0, row_stride=1) -> RenderParams:
        return RenderParams(width, height, spp, max_depth, rr_start_depth, integrator, sample_offset,
                            sample_stride, seed, pool_paths, flags, row_offset, row_stride)

    def render(self, params: RenderParams, out=None):
        """Host-buffer render: returns (accum[h, w, 4] float32 linear SUMS, stats dict)."""
        if out is None:
            out = np.empty((params.height, params.width, 4), np.float32)
        st = RenderStats()
        self._check(self._lib.rtb_render(self._h, C.byref(params), _ptr(out), C.byref(st)))
        return out, st.as_dict()

    def render_device(self, params: RenderParams, device_ptr: int, stream: int = 0):
        """Render into a caller-owned device buffer (e.g. a torch tensor's data_ptr())."""
        st = RenderStats()
        self._check(self._lib.rtb_render_device(self._h, C.byref(params), C.c_void_p(device_ptr),
                                                C.c_void_p(stream), C.byref(st)))
        return st.as_dict()

    def cancel(self):
        self._check(self._lib.rtb_cancel(self._h))

    # ---- multi-GPU, one process per GPU: the library's own NCCL communicator
    @staticmethod
    def comm_unique_id() -> bytes:
        """On rank 0; hand the bytes to every rank (e.g. torch.distributed.broadcast_object_list)."""
        buf = C.create_string_buffer(COMM_ID_BYTES)
        rc = load().rtb_comm_unique_id(C.cast(buf, C.c_void_p))
        if rc != RTB_OK:
            raise RtbError(rc, "rtb_comm_unique_id failed (libnccl.so.2 missing?)")
        return buf.raw

    def comm_init(self, n_ranks: int, rank: int, unique_id: bytes = None):
        buf = C.create_string_buffer(unique_id, COMM_ID_BYTES) if unique_id is not None else None
        self._check(self._lib.rtb_comm_init(self._h, n_ranks, rank, C.cast(buf, C.c_void_p) if buf is not None else None))

    def render_reduce(self, params: RenderParams, out=None, rgb8=None, stream: int = 0):
        """Collective: this rank's slice of the WHOLE job + ncclReduce onto rank 0 (inside the library).
        out / rgb8: optional host arrays (rank 0).  Returns the stats dict."""
        st = RenderStats()
        self._check(self._lib.rtb_render_reduce(self._h, C.byref(params), _ptr(out) if out is not None else None,
                                                _ptr(rgb8) if rgb8 is not None else None, C.c_void_p(stream), C.byref(st)))
        return st.as_dict()

    def accum_copy_device(self, device_ptr: int, stream: int = 0):
        self._check(self._lib.rtb_accum_copy_device(self._h, C.c_void_p(device_ptr), C.c_void_p(stream)))

    def resolve_rgb8(self, width, height, spp):
        out = np.empty((height, width, 3), np.uint8)
        self._check(self._lib.rtb_resolve_rgb8(self._h, spp, _ptr(out)))
        return out

    # ---- parity-layer batches
    def trace(self, rays, precision=64, want_visits=False):
        rays = np.ascontiguousarray(rays, dtype=abi.RAY)
        hits = np.zeros(rays.size, abi.HIT)
        visits = np.zeros(2, np.uint64)
        self._check(self._lib.rtb_trace_batch(self._h, _ptr(rays), rays.size, precision, _ptr(hits),
                                              _ptr(visits) if want_visits else None))
        return (hits, visits) if want_visits else hits

    def bsdf_eval(self, material, queries, precision=64):
        q = np.ascontiguousarray(queries, dtype=abi.BSDF_QUERY)
        out = np.zeros(q.size, abi.BSDF_VALUE)
        self._check(self._lib.rtb_bsdf_eval_batch(self._h, material, _ptr(q), q.size, precision, _ptr(out)))
        return out

    def bsdf_sample(self, material, queries, precision=64, seed=1):
        q = np.ascontiguousarray(queries, dtype=abi.BSDF_QUERY)
        out = np.zeros(q.size, abi.BSDF_SAMPLE)
        self._check(self._lib.rtb_bsdf_sample_batch(self._h, material, _ptr(q), q.size, precision, seed,
                                                    _ptr(out)))
        return out

    def light_eval(self, light, queries, precision=64, seed=1):
        q = np.ascontiguousarray(queries, dtype=abi.LIGHT_QUERY)
        out = np.zeros(q.size, abi.LIGHT_VALUE)
        self._check(self._lib.rtb_light_eval_batch(self._h, light, _ptr(q), q.size, precision, seed,
                                                   _ptr(out)))
        return out

    def texture_eval(self, texture, uvp, precision=64):
        uvp = np.ascontiguousarray(uvp, dtype=np.float64).reshape(-1, 5)
        out = np.zeros((uvp.shape[0], 3))
        self._check(self._lib.rtb_texture_eval_batch(self._h, texture, _ptr(uvp), uvp.shape[0], precision,
                                                     _ptr(out)))
        return out


class Group:
    """One rtb_group: several GPUs driven from this process (one host thread, stream and NCCL
    communicator per device inside the library)."""

    def __init__(self, devices):
        self._lib = load()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = self._lib.rtb_group_create(C.cast(devs, C.c_void_p), len(devices), C.byref(h))
        if rc != RTB_OK:
            raise RtbError(rc, self._lib.rtb_last_error(None).decode())
        self._h = h
        self.devices = list(devices)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rtb_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != RTB_OK:
            raise RtbError(rc, self._lib.rtb_group_last_error(self._h).decode())

    def size(self) -> int:
        return self._lib.rtb_group_size(self._h)

    def set_option(self, option: int, value: int):
        for i in range(self.size()):
            rc = self._lib.rtb_set_option(self._lib.rtb_group_context(self._h, i), option, value)
            if rc != RTB_OK:
                raise RtbError(rc, "rtb_set_option")

    def upload_scene(self, blob: bytes):
        buf = C.create_string_buffer(blob, len(blob))
        self._check(self._lib.rtb_group_scene_upload(self._h, C.cast(buf, C.c_void_p), len(blob)))

    def render(self, params: RenderParams, out=None, rgb8=None, want_accum=True):
        """The whole job on all GPUs.  Returns (accum[h, w, 4] float32 linear SUMS or None, stats)."""
        if out is None and want_accum:
            out = np.empty((params.height, params.width, 4), np.float32)
        st = RenderStats()
        self._check(self._lib.rtb_group_render(self._h, C.byref(params), _ptr(out) if out is not None else None,
                                               _ptr(rgb8) if rgb8 is not None else None, C.byref(st)))
        return out, st.as_dict()

    def cancel(self):
        self._check(self._lib.rtb_group_cancel(self._h))
