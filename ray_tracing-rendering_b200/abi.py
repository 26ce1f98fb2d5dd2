"""numpy views of the POD records that cross the C-ABI (include/rtb200_types.h,
include/rtb200_scene.h).  Layouts are asserted against the C structs by
tests/test_abi_layout.py."""
import numpy as np

RAY = np.dtype([("o", "<f8", 3), ("d", "<f8", 3), ("time", "<f8"), ("t_min", "<f8"),
                ("t_max", "<f8"), ("origin_prim", "<i4"), ("reserved", "<i4")])
HIT = np.dtype([("t", "<f8"), ("p", "<f8", 3), ("normal", "<f8", 3), ("u", "<f8"), ("v", "<f8"),
                ("prim", "<i4"), ("front_face", "<i4"), ("material", "<i4"), ("reserved", "<i4")])
BSDF_QUERY = np.dtype([("p", "<f8", 3), ("normal", "<f8", 3), ("u", "<f8"), ("v", "<f8"),
                       ("wo", "<f8", 3), ("wi", "<f8", 3), ("front_face", "<i4"),
                       ("reserved", "<i4")])
BSDF_VALUE = np.dtype([("f", "<f8", 3), ("pdf", "<f8"), ("emitted_old", "<f8", 3),
                       ("emitted_new", "<f8", 3)])
BSDF_SAMPLE = np.dtype([("wi", "<f8", 3), ("f", "<f8", 3), ("pdf", "<f8"), ("ok", "<i4"),
                        ("is_specular", "<i4"), ("scatter_dir", "<f8", 3),
                        ("scatter_atten", "<f8", 3), ("scatter_ok", "<i4"), ("reserved", "<i4")])
LIGHT_QUERY = np.dtype([("p", "<f8", 3), ("d", "<f8", 3), ("u", "<f8", 2)])
LIGHT_VALUE = np.dtype([("Li", "<f8", 3), ("wi", "<f8", 3), ("pdf", "<f8"), ("dist", "<f8"),
                        ("is_delta", "<i4"), ("reserved", "<i4"), ("pdf_dir", "<f8"),
                        ("Le", "<f8", 3)])

assert RAY.itemsize == 80 and HIT.itemsize == 88 and BSDF_QUERY.itemsize == 120
assert BSDF_VALUE.itemsize == 80 and BSDF_SAMPLE.itemsize == 120
assert LIGHT_QUERY.itemsize == 64 and LIGHT_VALUE.itemsize == 104

# ---- scene blob sections ------------------------------------------------------------------
SEC_GLOBALS, SEC_CAMERA, SEC_PRIMS, SEC_CHAINS, SEC_XFORM_OPS, SEC_MATERIALS = 1, 2, 3, 4, 5, 6
SEC_TEXTURES, SEC_IMAGES, SEC_IMAGE_BYTES, SEC_PERLIN, SEC_LIGHTS, SEC_ENV_TEXELS = 7, 8, 9, 10, 11, 12
SEC_GATES = 13
PRIM_FLAG_BOUNDARY_ONLY, PRIM_FLAG_DUP_LEAF, PRIM_FLAG_GATED = 1, 2, 4

GLOBALS = np.dtype([("background", "<f8", 3), ("image_width", "<i4"), ("image_height", "<i4"),
                    ("samples_per_pixel", "<i4"), ("scene_id", "<i4")])
CAMERA = np.dtype([("lookfrom", "<f8", 3), ("lookat", "<f8", 3), ("vup", "<f8", 3), ("vfov", "<f8"),
                   ("aspect_ratio", "<f8"), ("aperture", "<f8"), ("focus_dist", "<f8"),
                   ("time0", "<f8"), ("time1", "<f8")])
PRIM = np.dtype([("type", "<i4"), ("material", "<i4"), ("chain", "<i4"), ("flags", "<i4"),
                 ("aux0", "<i4"), ("aux1", "<i4"), ("d", "<f8", 9)])
CHAIN = np.dtype([("first", "<i4"), ("count", "<i4")])
XFORM_OP = np.dtype([("kind", "<i4"), ("reserved", "<i4"), ("a", "<f8"), ("b", "<f8"), ("c", "<f8")])
MATERIAL = np.dtype([("type", "<i4"), ("tex", "<i4", 4), ("reserved", "<i4"), ("color", "<f8", 3),
                     ("fuzz", "<f8"), ("ir", "<f8")])
TEXTURE = np.dtype([("type", "<i4"), ("even", "<i4"), ("odd", "<i4"), ("image", "<i4"),
                    ("perlin", "<i4"), ("reserved", "<i4"), ("color", "<f8", 3), ("scale", "<f8")])
IMAGE = np.dtype([("width", "<i4"), ("height", "<i4"), ("offset", "<u8")])
PERLIN = np.dtype([("ranvec", "<f8", (256, 3)), ("perm_x", "<i4", 256), ("perm_y", "<i4", 256),
                   ("perm_z", "<i4", 256)])
LIGHT = np.dtype([("type", "<i4"), ("env_width", "<i4"), ("env_height", "<i4"),
                  ("env_is_probe", "<i4"), ("env_offset", "<u8"), ("Q", "<f8", 3), ("u", "<f8", 3),
                  ("v", "<f8", 3), ("intensity", "<f8", 3), ("cos_cutoff", "<f8")])
GATE = np.dtype([("prim", "<i4"), ("reserved", "<i4"), ("lo", "<f8", 3), ("hi", "<f8", 3)])

SECTION_DTYPES = {
    SEC_GLOBALS: GLOBALS, SEC_CAMERA: CAMERA, SEC_PRIMS: PRIM, SEC_CHAINS: CHAIN,
    SEC_XFORM_OPS: XFORM_OP, SEC_MATERIALS: MATERIAL, SEC_TEXTURES: TEXTURE, SEC_IMAGES: IMAGE,
    SEC_IMAGE_BYTES: np.dtype("u1"), SEC_PERLIN: PERLIN, SEC_LIGHTS: LIGHT,
    SEC_ENV_TEXELS: np.dtype("<f4"), SEC_GATES: GATE,
}
SECTION_NAMES = {
    SEC_GLOBALS: "globals", SEC_CAMERA: "camera", SEC_PRIMS: "prims", SEC_CHAINS: "chains",
    SEC_XFORM_OPS: "xform_ops", SEC_MATERIALS: "materials", SEC_TEXTURES: "textures",
    SEC_IMAGES: "images", SEC_IMAGE_BYTES: "image_bytes", SEC_PERLIN: "perlins",
    SEC_LIGHTS: "lights", SEC_ENV_TEXELS: "env_texels", SEC_GATES: "gates",
}
SCENE_MAGIC = 0x31424C46
SCENE_VERSION = 2

_HDR = np.dtype([("magic", "<u4"), ("version", "<u4"), ("total_bytes", "<u8"),
                 ("n_sections", "<u4"), ("reserved", "<u4")])
_SEC = np.dtype([("id", "<u4"), ("stride", "<u4"), ("count", "<u8"), ("offset", "<u8")])


def parse_blob(blob: bytes) -> dict:
    """Scene blob -> {section name: numpy record array} (read-only views)."""
    buf = np.frombuffer(blob, dtype=np.uint8)
    hdr = buf[:_HDR.itemsize].view(_HDR)[0]
    if hdr["magic"] != SCENE_MAGIC or hdr["version"] != SCENE_VERSION:
        raise ValueError("not an rtb200 scene blob")
    n = int(hdr["n_sections"])
    secs = buf[_HDR.itemsize:_HDR.itemsize + n * _SEC.itemsize].view(_SEC)
    out = {}
    for s in secs:
        dt = SECTION_DTYPES[int(s["id"])]
        if dt.itemsize != int(s["stride"]):
            raise ValueError(f"section {int(s['id'])}: stride {int(s['stride'])} != {dt.itemsize}")
        off, cnt = int(s["offset"]), int(s["count"])
        out[SECTION_NAMES[int(s["id"])]] = buf[off:off + cnt * dt.itemsize].view(dt)
    for sid, name in SECTION_NAMES.items():     # optional sections a writer left out
        out.setdefault(name, np.zeros(0, SECTION_DTYPES[sid]))
    return out


def build_blob(tables: dict) -> bytes:
    """Inverse of parse_blob (used by tests to author synthetic scenes)."""
    ids = sorted(SECTION_DTYPES)
    off = _HDR.itemsize + len(ids) * _SEC.itemsize
    off = (off + 7) & ~7
    secs = np.zeros(len(ids), _SEC)
    payloads = []
    for k, sid in enumerate(ids):
        dt = SECTION_DTYPES[sid]
        arr = np.ascontiguousarray(tables.get(SECTION_NAMES[sid], np.zeros(0, dt)), dtype=dt).reshape(-1)
        secs[k] = (sid, dt.itemsize, arr.size, off)
        payloads.append((off, arr.tobytes()))
        off = (off + arr.size * dt.itemsize + 7) & ~7
    blob = bytearray(off)
    hdr = np.zeros(1, _HDR)
    hdr[0] = (SCENE_MAGIC, SCENE_VERSION, off, len(ids), 0)
    blob[:_HDR.itemsize] = hdr.tobytes()
    blob[_HDR.itemsize:_HDR.itemsize + secs.nbytes] = secs.tobytes()
    for o, b in payloads:
        blob[o:o + len(b)] = b
    return bytes(blob)
