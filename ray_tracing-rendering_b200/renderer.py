"""Python mirror of the reference's Renderer (src/renderer/renderer.h:17-118) on top of
the C-ABI: same method names and argument meaning, `render` takes the flattened scene
blob instead of the shared_ptr graph.  The C++ mirror lives in host/."""
import numpy as np

from .binding import Context

# src/main.cpp:81-100
INTEGRATOR_NAMES = {0: "PathIntegrator", 1: "RRPathInterator", 2: "PBRPathIntegrator",
                    3: "DirectLightIntegrator", 4: "MISPathIntegrator"}


class Renderer:
    def __init__(self, device: int = 0, ctx: Context = None):
        self.ctx = ctx if ctx is not None else Context(device)
        self._integrator = 4      # main.cpp:52
        self._spp = 10            # Renderer::Settings, renderer.h:19-21
        self._max_depth = 50      # main.cpp:102
        self._rr_start_depth = 3  # e.g. rr_path_integrator.h:63
        self._rendering = False
        self.last_stats = None

    # renderer.h:26 — the reference passes an Integrator object; its id selects the kernels
    def set_integrator(self, integrator_id: int):
        if integrator_id not in INTEGRATOR_NAMES:
            integrator_id = 4  # main.cpp:97-99 default
        self._integrator = integrator_id

    def set_samples(self, samples: int):  # renderer.h:104
        self._spp = int(samples)

    def set_max_depth(self, depth: int):  # renderer.h:107-111
        self._max_depth = int(depth)

    def set_rr_start_depth(self, depth: int):
        self._rr_start_depth = int(depth)

    def cancel(self):  # renderer.h:113
        self.ctx.cancel()

    def is_rendering(self) -> bool:  # renderer.h:116
        return self._rendering

    def render(self, scene_blob: bytes, width: int, height: int, seed: int = 1, upload: bool = True):
        """renderer.h:30-102.  Returns the RenderBuffer contents: (height, width, 3) float64,
        sqrt(mean) clamped to [0,1], row 0 = bottom (renderer.h:126-140)."""
        if upload:
            self.ctx.upload_scene(scene_blob)
        p = self.ctx.params(width, height, self._spp, self._integrator, self._max_depth,
                            self._rr_start_depth, seed)
        self._rendering = True
        try:
            acc, self.last_stats = self.ctx.render(p)
        finally:
            self._rendering = False
        self.last_linear = acc[..., :3].astype(np.float64) / max(self._spp, 1)
        return np.clip(np.sqrt(self.last_linear), 0.0, 1.0)


def to_png_bytes(buffer):
    """RenderBuffer::save_to_png pixel conversion (render_buffer.h:35-55): y flip and
    (unsigned char)(x * 255) truncation.  Returns (height, width, 3) uint8, top row first."""
    return (np.asarray(buffer)[::-1] * 255.0).astype(np.uint8)
