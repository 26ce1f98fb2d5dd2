"""The BASELINE.json configurations as data: which scene blob, image size,
samples per pixel, depth and integrator id.  bench.py, tools/run_config.py and the GPU tests
take their workloads from here so that every number in DESIGN.md names one of these."""
from dataclasses import dataclass
from typing import Callable

from . import scenes


@dataclass(frozen=True)
class Config:
    name: str
    workload: str                 # the text BASELINE.json uses for it
    blob: Callable[[], bytes]     # scene builder (the product's own; no reference code)
    scene_id: int                 # select_scene() id of the reference, or -1 (synthetic)
    width: int
    height: int
    spp: int
    integrator: int
    depth: int = 50


_env_cache = {}


def _env_blob():
    if "e" not in _env_cache:
        _env_cache["e"] = scenes.hdr_demo(1920, scenes.synthetic_hdr(2048, 1024, 1), 200)
    return _env_cache["e"]


CONFIGS = {
    "C1": Config("C1", "scene07 Cornell box 600x600 spp=400 kMaxDepth=50 integrator 1 (Russian roulette)",
                 lambda: scenes.cornell_box(False), 7, 600, 600, 400, 1),
    "C2": Config("C2", "scene09 RTiOW final scene 800x800 spp=500 integrator 1",
                 lambda: scenes.final_scene(1), 9, 800, 800, 500, 1),
    "C3": Config("C3", "scene21 Cornell box with NEE 600x600 spp=400 depth 50 integrator 3",
                 lambda: scenes.cornell_box(True), 21, 600, 600, 400, 3),
    "C4": Config("C4", "scene23 area lights + glossy Cook-Torrance spheres 800x450 spp=64 integrator 4 (MIS)",
                 scenes.mis_comparison, 23, 800, 450, 64, 4),
    "C4env": Config("C4env", "hdr_demo_scene + synthetic 2048x1024 equirect HDR env light 1920x1080 spp=200 integrator 4",
                    _env_blob, 24, 1920, 1080, 200, 4),
    "C5": Config("C5", "synthetic 1M-sphere field 3840x2160 spp=1024 integrator 4 (MIS)",
                 lambda: scenes.sphere_field(500, 3840, 2160, 1024), -1, 3840, 2160, 1024, 4),
}


def get(name: str) -> Config:
    if name not in CONFIGS:
        raise KeyError(f"unknown configuration {name!r}; known: {sorted(CONFIGS)}")
    return CONFIGS[name]
