// Renderer::render of host/rtb_host.hpp on several GPUs of one box, with no Python in the process:
// the library drives one host thread, stream and NCCL communicator per device (rtb_group_*).
// Usage: host_group_test <n_devices>.  Prints "key value..." lines for tests/test_gpu_multi.py.
#include "rtb_scenes.hpp"

#include <cstdio>
#include <cstdlib>

static void image_mean(const RenderBuffer &buf, double mean[3]) {
    mean[0] = mean[1] = mean[2] = 0;
    double n = 0;
    for (const auto &row : buf.get_data())
        for (const auto &px : row) {
            for (int k = 0; k < 3; ++k)
                mean[k] += px[k] * px[k]; // undo the sqrt
            n += 1;
        }
    for (int k = 0; k < 3; ++k)
        mean[k] /= n;
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 2;
    try {
        rtb::SceneSetup c = rtb::builtin_scene(21);
        auto cam = make_shared<camera>(c.lookfrom, c.lookat, c.vup, c.vfov, c.aspect_ratio, c.aperture, c.focus_dist, 0.0, 1.0);
        double m1[3], m2[3];
        {
            RenderBuffer buf(128, 128);
            Renderer r;
            r.set_samples(64);
            r.set_integrator(make_shared<MISPathIntegrator>());
            r.set_max_depth(50);
            r.set_preview_passes(1);
            r.render(c.world, cam, c.background, buf, c.lights);
            image_mean(buf, m1);
        }
        RenderBuffer buf(128, 128);
        Renderer r;
        std::vector<int> devs;
        for (int i = 0; i < n; ++i)
            devs.push_back(i);
        r.set_devices(devs);
        r.set_samples(64);
        r.set_integrator(make_shared<MISPathIntegrator>());
        r.set_max_depth(50);
        r.set_preview_passes(1);
        r.render(c.world, cam, c.background, buf, c.lights);
        image_mean(buf, m2);
        std::printf("devices %d\n", r.device_count());
        std::printf("mean1 %.6f %.6f %.6f\n", m1[0], m1[1], m1[2]);
        std::printf("mean2 %.6f %.6f %.6f\n", m2[0], m2[1], m2[2]);
        std::printf("paths2 %llu\n", (unsigned long long)r.last_stats().paths);
        return 0;
    } catch (const std::exception &e) {
        std::printf("exception %s\n", e.what());
        return 2;
    }
}
