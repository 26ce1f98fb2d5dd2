"""The numpy dtypes of ray_tracing-rendering_b200/abi.py must describe exactly the C structs
of include/rtb200_types.h and include/rtb200_scene.h (sizes and field offsets)."""
import os
import subprocess
import tempfile

from conftest import ROOT

STRUCTS = {
    "rtb_ray": "RAY", "rtb_hit": "HIT", "rtb_bsdf_query": "BSDF_QUERY", "rtb_bsdf_value": "BSDF_VALUE",
    "rtb_bsdf_sample": "BSDF_SAMPLE", "rtb_light_query": "LIGHT_QUERY", "rtb_light_value": "LIGHT_VALUE",
    "rtb_globals": "GLOBALS", "rtb_camera": "CAMERA", "rtb_prim": "PRIM", "rtb_chain": "CHAIN",
    "rtb_xform_op": "XFORM_OP", "rtb_material": "MATERIAL", "rtb_texture": "TEXTURE", "rtb_image": "IMAGE",
    "rtb_perlin": "PERLIN", "rtb_light": "LIGHT",
}


def test_struct_layouts_match_c(abi):
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "rtb200.h"', 'int main(void){']
    for c_name, py_name in STRUCTS.items():
        dt = getattr(abi, py_name)
        lines.append(f'printf("{c_name} size %zu\\n", sizeof({c_name}));')
        for f in dt.names:
            lines.append(f'printf("{c_name} {f} %zu\\n", offsetof({c_name}, {f}));')
    lines += ['return 0;}']
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "layout.c")
        exe = os.path.join(tmp, "layout")
        with open(src, "w") as fh:
            fh.write("\n".join(lines))
        # plain C compiler: the public headers must be valid C
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                               src, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    got = {}
    for ln in out.splitlines():
        s, f, v = ln.split()
        got[(s, f)] = int(v)
    for c_name, py_name in STRUCTS.items():
        dt = getattr(abi, py_name)
        assert got[(c_name, "size")] == dt.itemsize, c_name
        for f in dt.names:
            assert got[(c_name, f)] == dt.fields[f][1], (c_name, f)


def test_render_param_structs_match_c(binding):
    import ctypes as C
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "rtb200.h"
int main(void){
 printf("%zu %zu %zu\n", sizeof(rtb_render_params), sizeof(rtb_render_stats), sizeof(rtb_scene_stats));
 printf("%zu %zu %zu\n", offsetof(rtb_render_params, seed), offsetof(rtb_render_stats, device_ms), offsetof(rtb_scene_stats, device_bytes));
 return 0;}
'''
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "p.c")
        with open(src, "w") as fh:
            fh.write(prog)
        subprocess.check_call(["gcc", "-std=c99", "-I" + os.path.join(ROOT, "include"), src, "-o", src + ".x"])
        a, b = subprocess.check_output([src + ".x"], text=True).splitlines()
    sizes = [int(x) for x in a.split()]
    offs = [int(x) for x in b.split()]
    assert sizes == [C.sizeof(binding.RenderParams), C.sizeof(binding.RenderStats), C.sizeof(binding.SceneStats)]
    assert offs == [binding.RenderParams.seed.offset, binding.RenderStats.device_ms.offset,
                    binding.SceneStats.device_bytes.offset]
