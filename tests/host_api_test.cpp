// Exercises the reference-shaped C++ API of host/rtb_host.hpp on a live GPU: hit(), eval(),
// pdf(), emitted(), sample(), Light::sample(), texture::value() and Renderer::render.
// Prints "key value..." lines that tests/test_host_layer.py checks.
#include "rtb_scenes.hpp"

#include <cmath>
#include <cstdio>

int main() {
    try {
        rtb::SceneSetup c = rtb::builtin_scene(7);
        // a ray through the upper right of the opening (clear of both boxes) hits the back wall (z = 555) of the Cornell box
        hit_record rec;
        ray r(point3(500, 500, -800), vec3(0, 0, 1), 0.5);
        const bool ok = c.world->hit(r, 0.001, infinity, rec);
        std::printf("hit %d %.17g %.17g %.17g %.17g %d\n", ok, rec.t, rec.p.z(), rec.normal.z(), rec.u, int(rec.front_face));
        // straight up from the floor centre: the light at y = 554
        hit_record up;
        const bool ok2 = c.world->hit(ray(point3(278, 1, 279.5), vec3(0, 1, 0)), 0.001, infinity, up);
        const color e = up.mat_ptr->emitted(up.u, up.v, up.p);
        std::printf("light %d %.17g %.17g\n", ok2, up.t, e.x());
        const bool miss = c.world->hit(ray(point3(278, 278, -800), vec3(0, 0, -1)), 0.001, infinity, rec);
        std::printf("miss %d\n", miss);
        // lambertian eval = albedo / pi, pdf = cos / pi
        lambertian white(color(.73, .73, .73));
        hit_record s;
        s.p = point3(0, 0, 0);
        s.normal = vec3(0, 1, 0);
        s.front_face = true;
        const vec3 wo = unit_vector(vec3(0.3, 1, 0.2)), wi = unit_vector(vec3(-0.4, 0.8, 0.1));
        const color f = white.eval(s, wo, wi);
        std::printf("lambert %.17g %.17g\n", f.x(), white.pdf(s, wo, wi));
        BSDFSample bs;
        const bool sok = white.sample(s, wo, bs);
        std::printf("sample %d %.17g %.17g %d\n", sok, bs.wi.length(), bs.pdf - dot(bs.wi, s.normal) / pi, int(bs.is_specular));
        diffuse_light lamp(color(15, 15, 15));
        std::printf("lamp_sample %d\n", int(lamp.sample(s, wo, bs)));
        QuadLight q(point3(213, 554, 227), vec3(130, 0, 0), vec3(0, 0, 105), color(15, 15, 15));
        const LightSample ls = q.sample(point3(278, 0, 279.5), vec2(0.5, 0.5));
        std::printf("quad %.17g %.17g %.17g %d\n", ls.dist, ls.pdf, ls.Li.x(), int(ls.is_delta));
        checker_texture chk(color(0.2, 0.3, 0.1), color(0.9, 0.9, 0.9));
        const color tv = chk.value(0, 0, point3(0.05, 0.05, 0.05));
        std::printf("checker %.17g\n", tv.x());
        // Renderer::render (renderer.h:30) with integrator 1 at a small size
        auto cam = make_shared<camera>(c.lookfrom, c.lookat, c.vup, c.vfov, c.aspect_ratio, c.aperture, c.focus_dist, 0.0, 1.0);
        RenderBuffer buf(64, 64);
        Renderer renderer;
        renderer.set_samples(256);
        renderer.set_integrator(make_shared<RRPathInterator>());
        renderer.set_max_depth(50);
        renderer.render(c.world, cam, c.background, buf, c.lights);
        double sum[3] = {0, 0, 0};
        for (const auto &row : buf.get_data())
            for (const auto &px : row)
                for (int k = 0; k < 3; ++k)
                    sum[k] += px[k] * px[k]; // undo the sqrt: mean of clamped linear values
        std::printf("render %.6f %.6f %.6f %d\n", sum[0] / 4096, sum[1] / 4096, sum[2] / 4096, int(renderer.is_rendering()));
        std::printf("png %d\n", int(buf.save_to_png("/tmp/rtb_host_api_test.png")));
        // progressive preview: four sample passes add up to the one-call image (same sample set)
        RenderBuffer buf4(64, 64);
        renderer.set_preview_passes(4);
        renderer.render(c.world, cam, c.background, buf4, c.lights);
        double worst = 0;
        for (int j = 0; j < 64; ++j)
            for (int i = 0; i < 64; ++i)
                for (int k = 0; k < 3; ++k)
                    worst = std::max(worst, std::fabs(buf.get_data()[j][i][k] - buf4.get_data()[j][i][k]));
        std::printf("passes %d %.3g %llu\n", renderer.passes_done(), worst, (unsigned long long)renderer.last_stats().paths);
        return 0;
    } catch (const std::exception &e) {
        std::printf("exception %s\n", e.what());
        return 2;
    }
}
