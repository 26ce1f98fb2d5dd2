// rtb_scenes.hpp — the BASELINE.json scenes built with the host-layer classes, for programs
// that do not link the reference's scenes.cpp.  What each scene CONTAINS is data taken from
// the reference (cited); the code is this repo's.
#ifndef RTB_SCENES_HPP
#define RTB_SCENES_HPP

#include "rtb_host.hpp"

namespace rtb {

struct SceneSetup { // SceneConfig of the reference (scenes.h:11-24)
    shared_ptr<hittable> world;
    std::vector<shared_ptr<Light>> lights;
    color background{0, 0, 0};
    point3 lookfrom{13, 2, 3}, lookat{0, 0, 0};
    vec3 vup{0, 1, 0};
    double vfov = 40.0, aperture = 0.0, focus_dist = 10.0, aspect_ratio = 16.0 / 9.0;
    int image_width = 1280, samples_per_pixel = 100;
};

// Cornell box: scenes.cpp:159-187 (scene 7) and :779-809 (scene 21, emitter behind flip_face
// + a QuadLight over it, scenes.cpp:1729-1744).
inline SceneSetup cornell(bool nee) {
    SceneSetup c;
    hittable_list objs;
    const auto white = make_shared<lambertian>(color(.73, .73, .73));
    const auto lamp = make_shared<diffuse_light>(color(15, 15, 15));
    objs.add(make_shared<yz_rect>(0, 555, 0, 555, 555, make_shared<lambertian>(color(.12, .45, .15))));
    objs.add(make_shared<yz_rect>(0, 555, 0, 555, 0, make_shared<lambertian>(color(.65, .05, .05))));
    shared_ptr<hittable> emitter = make_shared<xz_rect>(213, 343, 227, 332, 554, lamp);
    objs.add(nee ? shared_ptr<hittable>(make_shared<flip_face>(emitter)) : emitter);
    objs.add(make_shared<xz_rect>(0, 555, 0, 555, 0, white));
    objs.add(make_shared<xz_rect>(0, 555, 0, 555, 555, white));
    objs.add(make_shared<xy_rect>(0, 555, 0, 555, 555, white));
    struct Block { double h, angle; vec3 at; };
    for (const Block &b : {Block{330, 15, vec3(265, 0, 295)}, Block{165, -18, vec3(130, 0, 65)}})
        objs.add(make_shared<translate>(
            make_shared<rotate_y>(make_shared<box>(point3(0, 0, 0), point3(165, b.h, 165), white), b.angle), b.at));
    c.world = make_shared<bvh_node>(objs, 0, 1);
    if (nee)
        c.lights.push_back(make_shared<QuadLight>(point3(213, 554, 227), vec3(130, 0, 0), vec3(0, 0, 105), color(15, 15, 15)));
    c.aspect_ratio = 1.0;
    c.image_width = 600;
    c.samples_per_pixel = 400;
    c.lookfrom = point3(278, 278, -800);
    c.lookat = point3(278, 278, 0);
    return c;
}

// scenes.cpp:580-626 + case 23 (:1762-1781): three spheres under a large and a small area light.
inline SceneSetup mis_comparison() {
    SceneSetup c;
    hittable_list w;
    auto flat = [](double r, double g, double b) { return make_shared<solid_color>(r, g, b); };
    w.add(make_shared<sphere>(point3(0, -1000, 0), 1000, make_shared<lambertian>(color(0.5, 0.5, 0.5))));
    w.add(make_shared<sphere>(point3(-2.5, 1, 0), 1.0,
                              make_shared<PBRMaterial>(flat(0.9, 0.6, 0.2), flat(0.001, 0.001, 0.001), flat(1, 1, 1))));
    w.add(make_shared<sphere>(point3(0, 1, 0), 1.0,
                              make_shared<PBRMaterial>(flat(0.8, 0.8, 0.8), flat(0.4, 0.4, 0.4), flat(1, 1, 1))));
    w.add(make_shared<sphere>(point3(2.5, 1, 0), 1.0, make_shared<dielectric>(1.5)));
    w.add(make_shared<flip_face>(make_shared<xz_rect>(-10, 10, -10, 10, 10, make_shared<diffuse_light>(color(5, 5, 5)))));
    w.add(make_shared<flip_face>(make_shared<yz_rect>(3.75, 4.25, 1.75, 2.25, 6, make_shared<diffuse_light>(color(50, 50, 50)))));
    c.world = make_shared<bvh_node>(w, 0, 1);
    c.lights.push_back(make_shared<QuadLight>(point3(-10, 10, -10), vec3(20, 0, 0), vec3(0, 0, 20), color(5, 5, 5)));
    c.lights.push_back(make_shared<QuadLight>(point3(6, 4, 2), vec3(0, 0.5, 0), vec3(0, 0, 0.5), color(50, 50, 50)));
    c.image_width = 800;
    c.samples_per_pixel = 64;
    c.lookfrom = point3(0, 3, 8);
    c.lookat = point3(0, 1, 0);
    c.vfov = 35.0;
    return c;
}

inline SceneSetup builtin_scene(int id) {
    switch (id) {
    case 7: return cornell(false);
    case 21: return cornell(true);
    case 23: return mis_comparison();
    default: throw std::runtime_error("builtin_scene: only the BASELINE.json scenes 7, 21 and 23 are built in");
    }
}

} // namespace rtb

#endif // RTB_SCENES_HPP
