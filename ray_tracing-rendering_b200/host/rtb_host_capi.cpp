// rtb_host_capi.cpp — small C entry points over the C++ host layer (librtb200_host.so), used by
// the tests and by programs that want the host layer without a C++ compiler:
//   rtbh_builtin_scene_blob   builds scene 7 / 21 / 23 with the host classes and flattens it;
//   rtbh_render_png           the headless equivalent of src/main.cpp:49-151.
// When compiled with -DRTBH_WITH_REFERENCE_SCENES the reference's own scenes.cpp (compiled
// against host/compat, unchanged) provides select_scene() for all 40 scene ids.
#include "rtb_scenes.hpp"

#ifdef RTBH_WITH_REFERENCE_SCENES
#include "scenes.h" // the reference's: declares SceneConfig and select_scene
#endif

#include <cstdlib>
#include <cstring>

namespace {
thread_local std::string g_err;

template <class Cfg> std::vector<uint8_t> blob_of(const Cfg &c, int id) {
    camera cam(c.lookfrom, c.lookat, c.vup, c.vfov, c.aspect_ratio, c.aperture, c.focus_dist, 0.0, 1.0); // main.cpp:63-66
    const int w = c.image_width, h = static_cast<int>(w / c.aspect_ratio);                                // main.cpp:68-69
    return rtb::flatten(*c.world, cam, c.background, c.lights, w, h, c.samples_per_pixel, id);
}
} // namespace

extern "C" {

__attribute__((visibility("default"))) const char *rtbh_last_error() { return g_err.c_str(); }

// Returns a malloc()ed blob (caller frees with rtbh_free) or NULL.
__attribute__((visibility("default"))) uint8_t *rtbh_builtin_scene_blob(int scene_id, uint64_t *nbytes) {
    try {
        const auto b = blob_of(rtb::builtin_scene(scene_id), scene_id);
        uint8_t *out = static_cast<uint8_t *>(std::malloc(b.size()));
        std::memcpy(out, b.data(), b.size());
        *nbytes = b.size();
        return out;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

#ifdef RTBH_WITH_REFERENCE_SCENES
__attribute__((visibility("default"))) uint8_t *rtbh_reference_scene_blob(int scene_id, uint32_t seed, uint64_t *nbytes) {
    try {
        rtb_host_seed(seed);
        const SceneConfig c = select_scene(scene_id);
        if (!c.world)
            throw std::runtime_error("select_scene returned no world");
        const auto b = blob_of(c, scene_id);
        uint8_t *out = static_cast<uint8_t *>(std::malloc(b.size()));
        std::memcpy(out, b.data(), b.size());
        *nbytes = b.size();
        return out;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
#endif

__attribute__((visibility("default"))) void rtbh_free(void *p) { std::free(p); }

// src/main.cpp:49-151 without the window: scene id + integrator id -> PNG file.
// width/spp <= 0 keep the scene's own values.  Returns 0 on success.
__attribute__((visibility("default"))) int rtbh_render_png(int scene_id, int integrator_id, int width, int spp, const char *path,
                                                           double *seconds) {
    try {
        const rtb::SceneSetup c = rtb::builtin_scene(scene_id);
        auto cam = make_shared<camera>(c.lookfrom, c.lookat, c.vup, c.vfov, c.aspect_ratio, c.aperture, c.focus_dist, 0.0, 1.0);
        const int w = width > 0 ? width : c.image_width, h = static_cast<int>(w / c.aspect_ratio);
        RenderBuffer buffer(w, h);
        Renderer renderer;
        renderer.set_samples(spp > 0 ? spp : c.samples_per_pixel);
        shared_ptr<Integrator> integ;
        switch (integrator_id) { // main.cpp:81-100
        case 0: integ = make_shared<PathIntegrator>(); break;
        case 1: integ = make_shared<RRPathInterator>(); break;
        case 2: integ = make_shared<PBRPathIntegrator>(); break;
        case 3: integ = make_shared<DirectLightIntegrator>(); break;
        default: integ = make_shared<MISPathIntegrator>(); break;
        }
        renderer.set_integrator(integ);
        renderer.set_max_depth(50); // main.cpp:102
        renderer.render(c.world, cam, c.background, buffer, c.lights);
        if (seconds)
            *seconds = renderer.last_stats().device_ms * 1e-3;
        if (path && !buffer.save_to_png(path))
            throw std::runtime_error(std::string("could not write ") + path);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}
}
