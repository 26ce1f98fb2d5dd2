// rtb_host.hpp — C++ host layer: the reference's class API as a scene DESCRIPTION that
// flattens itself into the blob of include/rtb200_scene.h and renders through the C-ABI
// of include/rtb200.h.
//
// Same class names, constructors and public members as the reference
// (JiGuang283/Ray_Tracing-Rendering, src/{core,geometry,materials,lighting,renderer,scene}),
// so scene-builder code written against the reference compiles against these headers
// unchanged (tests/test_host_layer.py compiles the reference's own scenes.cpp against them).
// What differs is where the work happens: there is no CPU intersection, shading or
// integration code here.  hittable::hit, material::eval/pdf/sample/scatter/emitted,
// Light::sample/pdf/Le and texture::value are implemented as batch-of-one calls into the
// GPU library (slow, but the answers are the kernels' answers); Renderer::render is one
// rtb_scene_upload + rtb_render.  Without a CUDA device every one of them throws.
//
// Header-only, C++14, depends on include/rtb200.h + rtb200_blob.hpp and -lrtb200.
#ifndef RTB_HOST_HPP
#define RTB_HOST_HPP

#include "rtb200.h"
#include "rtb200_blob.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

using std::make_shared;
using std::make_unique;
using std::shared_ptr;
using std::sqrt;
using std::unique_ptr;

// ---- src/core/rtweekend.h -----------------------------------------------------------------
constexpr double infinity = std::numeric_limits<double>::infinity();
constexpr double pi = 3.1415926535897932385;
inline constexpr double degrees_to_radians(double degrees) { return degrees * pi / 180.0; }

// The reference seeds its xorshift32 from the thread id (rtweekend.h:26-27), which makes
// every random scene unrepeatable; here the same generator starts from a fixed seed that
// rtb_host_seed() can change.
inline uint32_t &rtb_host_rng_state() {
    static thread_local uint32_t s = 0x9E3779B9u;
    return s;
}
inline void rtb_host_seed(uint32_t seed) { rtb_host_rng_state() = seed ? seed : 1u; }
// Marsaglia's 13/17/5 xorshift, the generator rtweekend.h:24-34 uses, so that a scene built here from the
// reference's seed has the reference's content; the draw is state / 2^32
inline uint32_t rtb_host_xorshift32(uint32_t x) {
    x ^= x << 13;
    x ^= x >> 17;
    return x ^ (x << 5);
}
inline double random_double() {
    uint32_t &st = rtb_host_rng_state();
    st = rtb_host_xorshift32(st);
    return double(st) * (1.0 / 4294967296.0);
}
inline double random_double(double min, double max) noexcept { return min + (max - min) * random_double(); }
inline double clamp(double x, double min, double max) noexcept { return x < min ? min : (x > max ? max : x); }
inline int random_int(int min, int max) { return static_cast<int>(random_double(min, max + 1)); }

// ---- src/core/vec3.h ------------------------------------------------------------------------
// (same public interface as the reference's vec3 — scenes.cpp and user code compile against it — written
// here over two element-wise helpers)
class vec3 {
  public:
    double e[3];
    vec3() : e{0, 0, 0} {}
    vec3(double e0, double e1, double e2) : e{e0, e1, e2} {}
    template <class F> static vec3 map(const vec3 &a, F f) { return vec3(f(a.e[0]), f(a.e[1]), f(a.e[2])); }
    template <class F> static vec3 zip(const vec3 &a, const vec3 &b, F f) {
        return vec3(f(a.e[0], b.e[0]), f(a.e[1], b.e[1]), f(a.e[2], b.e[2]));
    }
    double x() const noexcept { return e[0]; }
    double y() const noexcept { return e[1]; }
    double z() const noexcept { return e[2]; }
    double operator[](int i) const { return e[i]; }
    double &operator[](int i) { return e[i]; }
    vec3 operator-() const { return map(*this, [](double a) { return -a; }); }
    vec3 &operator+=(const vec3 &v) { return *this = zip(*this, v, [](double a, double b) { return a + b; }); }
    vec3 &operator*=(const vec3 &v) { return *this = zip(*this, v, [](double a, double b) { return a * b; }); }
    vec3 &operator*=(const double t) { return *this = map(*this, [t](double a) { return a * t; }); }
    vec3 &operator/=(const double t) { return *this *= 1 / t; }
    double length_squared() const noexcept { return e[0] * e[0] + e[1] * e[1] + e[2] * e[2]; }
    double length() const { return sqrt(length_squared()); }
    bool near_zero() const noexcept {
        const double tiny = 1e-8;
        return fabs(e[0]) < tiny && fabs(e[1]) < tiny && fabs(e[2]) < tiny;
    }
    static vec3 random() {
        const double a = random_double(), b = random_double(), c = random_double(); // (a fixed order: the reference leaves it to the compiler)
        return vec3(a, b, c);
    }
    static vec3 random(double min, double max) {
        const double a = random_double(min, max), b = random_double(min, max), c = random_double(min, max);
        return vec3(a, b, c);
    }
};
class vec2 {
  public:
    double e[2];
    vec2() : e{0, 0} {}
    vec2(double e0, double e1) : e{e0, e1} {}
    double x() const { return e[0]; }
    double y() const { return e[1]; }
    double operator[](int i) const { return e[i]; }
    double &operator[](int i) { return e[i]; }
};
using point3 = vec3;
using color = vec3;
inline std::ostream &operator<<(std::ostream &out, const vec3 &v) { return out << v.e[0] << ' ' << v.e[1] << ' ' << v.e[2]; }
inline vec3 operator+(const vec3 &u, const vec3 &v) { return vec3::zip(u, v, [](double a, double b) { return a + b; }); }
inline vec3 operator-(const vec3 &u, const vec3 &v) { return vec3::zip(u, v, [](double a, double b) { return a - b; }); }
inline vec3 operator*(const vec3 &u, const vec3 &v) { return vec3::zip(u, v, [](double a, double b) { return a * b; }); }
inline vec3 operator*(double t, const vec3 &v) { return vec3::map(v, [t](double a) { return t * a; }); }
inline vec3 operator*(const vec3 &v, double t) { return t * v; }
inline vec3 operator/(vec3 v, double t) { return (1 / t) * v; }
inline double dot(const vec3 &u, const vec3 &v) { return u.e[0] * v.e[0] + u.e[1] * v.e[1] + u.e[2] * v.e[2]; }
inline vec3 cross(const vec3 &u, const vec3 &v) {
    const double *a = u.e, *b = v.e;
    return vec3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
inline vec3 unit_vector(const vec3 &v) { return v / v.length(); }

// ---- src/core/ray.h -------------------------------------------------------------------------
class ray {
  public:
    ray() = default;
    ray(const point3 &origin, const vec3 &direction, double time = 0.0) noexcept : orig(origin), dir(direction), tm(time) {}
    point3 origin() const noexcept { return orig; }
    vec3 direction() const noexcept { return dir; }
    double time() const noexcept { return tm; }
    point3 at(double t) const noexcept { return orig + t * dir; }

  private:
    point3 orig;
    vec3 dir;
    double tm = 0.0;
};

class material;
class texture;
class hittable;
class Light;

// ---- src/geometry/aabb.h (bounds only; the slab test lives on the GPU) ---------------------------
class aabb {
  public:
    aabb() : minimum(0, 0, 0), maximum(0, 0, 0) {}
    aabb(const point3 &a, const point3 &b) : minimum(a), maximum(b) {}
    point3 min() const { return minimum; }
    point3 max() const { return maximum; }
    point3 minimum, maximum;
};
inline aabb surrounding_box(aabb a, aabb b) {
    return aabb(point3(fmin(a.min().x(), b.min().x()), fmin(a.min().y(), b.min().y()), fmin(a.min().z(), b.min().z())),
                point3(fmax(a.max().x(), b.max().x()), fmax(a.max().y(), b.max().y()), fmax(a.max().z(), b.max().z())));
}

// src/geometry/hittable.h:10-23
struct hit_record {
    point3 p;
    vec3 normal;
    material *mat_ptr = nullptr;
    double t = 0, u = 0, v = 0;
    bool front_face = false;
    inline void set_face_normal(const ray &r, const vec3 &outward_normal) {
        front_face = dot(r.direction(), outward_normal) < 0;
        normal = front_face ? outward_normal : -outward_normal;
    }
};

namespace rtb {

inline void check(int rc, rtb_context *ctx, const char *what) {
    if (rc != RTB_OK)
        throw std::runtime_error(std::string(what) + ": " + rtb_last_error(ctx));
}

// One process-wide GPU context for the batch-of-one API calls (hit / eval / sample ...).
inline rtb_context *api_context() {
    static rtb_context *ctx = nullptr;
    if (!ctx)
        check(rtb_context_create(0, &ctx), nullptr, "rtb_context_create");
    return ctx;
}

// Walks a description graph and accumulates the flat tables.
struct Flattener {
    SceneTables T;
    std::map<const texture *, int> tex_ids;
    std::map<const material *, int> mat_ids;
    std::vector<const material *> materials; // index -> object (to fill hit_record::mat_ptr)
    std::vector<const hittable *> wrappers;
    std::vector<rtb_xform_op> ops;
    std::map<std::vector<const hittable *>, int> chain_ids;
    bool boundary = false;

    int chain_id() {
        if (wrappers.empty())
            return -1;
        auto it = chain_ids.find(wrappers);
        if (it != chain_ids.end())
            return it->second;
        rtb_chain c;
        c.first = int(T.xform_ops.size());
        c.count = int(ops.size());
        T.xform_ops.insert(T.xform_ops.end(), ops.begin(), ops.end());
        T.chains.push_back(c);
        return chain_ids[wrappers] = int(T.chains.size()) - 1;
    }
    void push(const hittable *w, int kind, double a = 0, double b = 0, double c = 0) {
        rtb_xform_op op{};
        op.kind = kind;
        op.a = a;
        op.b = b;
        op.c = c;
        wrappers.push_back(w);
        ops.push_back(op);
    }
    void pop() {
        wrappers.pop_back();
        ops.pop_back();
    }
    int texture_id(const texture *t);
    int material_id(const material *m);
    int add_prim(int type, const material *m, int flags, std::initializer_list<double> d) {
        rtb_prim p{};
        p.type = type;
        p.material = material_id(m);
        p.chain = chain_id();
        p.flags = flags | (boundary ? int(RTB_PRIM_BOUNDARY_ONLY) : 0);
        int k = 0;
        for (double v : d)
            p.d[k++] = v;
        T.prims.push_back(p);
        return int(T.prims.size()) - 1;
    }
};

} // namespace rtb

// ---- src/materials/texture.h, perlin.h ---------------------------------------------------------
class texture {
  public:
    virtual ~texture() = default;
    virtual void describe(rtb::Flattener &F, rtb_texture &out) const = 0;
    // texture.h:13 — evaluated by the GPU library
    color value(double u, double v, const point3 &p) const;
    double value_scalar(double u, double v, const point3 &p) const { return value(u, v, p).x(); }
    vec3 value_normal(double u, double v, const point3 &p) const {
        color c = value(u, v, p);
        return unit_vector(c * 2.0 - color(1, 1, 1));
    }
    double value_roughness(double u, double v, const point3 &p) const { return value_scalar(u, v, p); }
    double value_metallic(double u, double v, const point3 &p) const { return value_scalar(u, v, p); }
};

class solid_color : public texture {
  public:
    solid_color() {}
    solid_color(color c) : color_value(c) {}
    solid_color(double r, double g, double b) : color_value(r, g, b) {}
    void describe(rtb::Flattener &, rtb_texture &o) const override {
        o.type = RTB_TEX_SOLID;
        for (int k = 0; k < 3; ++k)
            o.color[k] = color_value[k];
    }
    color color_value;
};

class checker_texture : public texture {
  public:
    checker_texture() {}
    checker_texture(shared_ptr<texture> _even, shared_ptr<texture> _odd) : odd(_odd), even(_even) {}
    checker_texture(color c1, color c2) : odd(make_shared<solid_color>(c2)), even(make_shared<solid_color>(c1)) {}
    void describe(rtb::Flattener &F, rtb_texture &o) const override {
        o.type = RTB_TEX_CHECKER;
        o.even = F.texture_id(even.get());
        o.odd = F.texture_id(odd.get());
    }
    shared_ptr<texture> odd, even;
};

// texture.h:82-146.  The reference decodes JPEG/PNG through the vendored stb_image, which is
// out of scope here; this loader reads binary PPM (P6) and otherwise behaves like the
// reference on a missing file: the texture renders cyan (texture.h:116-118).
class image_texture : public texture {
  public:
    image_texture() {}
    image_texture(const char *filename) {
        std::ifstream f(filename, std::ios::binary);
        std::string magic;
        int w = 0, h = 0, maxv = 0;
        if (f && (f >> magic >> w >> h >> maxv) && magic == "P6" && maxv == 255 && w > 0 && h > 0) {
            f.get();
            data.resize(size_t(w) * h * 3);
            f.read(reinterpret_cast<char *>(data.data()), std::streamsize(data.size()));
            if (f) {
                width = w;
                height = h;
            }
        }
        if (!width) {
            data.clear();
            std::cerr << "ERROR: Could not load texture image file '" << filename << "'.\n";
        }
    }
    image_texture(int w, int h, const unsigned char *rgb) : data(rgb, rgb + size_t(w) * h * 3), width(w), height(h) {}
    void describe(rtb::Flattener &F, rtb_texture &o) const override {
        o.type = RTB_TEX_IMAGE;
        rtb_image im{};
        im.width = width;
        im.height = height;
        im.offset = F.T.image_bytes.size();
        F.T.image_bytes.insert(F.T.image_bytes.end(), data.begin(), data.end());
        o.image = int(F.T.images.size());
        F.T.images.push_back(im);
    }
    std::vector<unsigned char> data;
    int width = 0, height = 0;
};

class perlin { // perlin.h:10-20, 77-94: the tables; evaluation is on the GPU
  public:
    perlin() {
        for (int i = 0; i < 256; ++i) {
            vec3 v = unit_vector(vec3::random(-1, 1));
            for (int k = 0; k < 3; ++k)
                tab.ranvec[i][k] = v[k];
        }
        permute(tab.perm_x);
        permute(tab.perm_y);
        permute(tab.perm_z);
    }
    rtb_perlin tab;

  private:
    static void permute(int32_t *p) {
        for (int i = 0; i < 256; i++)
            p[i] = i;
        for (int i = 255; i > 0; i--) {
            int target = random_int(0, i);
            std::swap(p[i], p[target]);
        }
    }
};

class noise_texture : public texture {
  public:
    noise_texture() {}
    noise_texture(double sc) : scale(sc) {}
    void describe(rtb::Flattener &F, rtb_texture &o) const override {
        o.type = RTB_TEX_NOISE;
        o.scale = scale;
        o.perlin = int(F.T.perlins.size());
        F.T.perlins.push_back(noise.tab);
    }
    perlin noise;
    double scale = 1.0;
};

// ---- src/materials/material.h -------------------------------------------------------------------
struct BSDFSample { // material.h:13-20
    vec3 wi;
    color f;
    double pdf = 0;
    bool is_specular = false;
    bool is_transmission = false;
};

class material {
  public:
    virtual ~material() = default;
    virtual void describe(rtb::Flattener &F, rtb_material &out) const = 0;
    // material.h:27-69 — every one of these is answered by the GPU library
    color emitted(double u, double v, const point3 &p) const;
    color emitted(const hit_record &rec, const vec3 &wo) const;
    virtual bool is_specular() const { return false; } // never overridden in the reference either
    bool sample(const hit_record &rec, const vec3 &wo, BSDFSample &sampled) const;
    color eval(const hit_record &rec, const vec3 &wo, const vec3 &wi) const;
    double pdf(const hit_record &rec, const vec3 &wo, const vec3 &wi) const;
    bool scatter(const ray &r_in, const hit_record &rec, color &attenuation, ray &scattered) const;
};

class lambertian : public material {
  public:
    lambertian(const color &a) : albedo(make_shared<solid_color>(a)) {}
    lambertian(shared_ptr<texture> a) : albedo(a) {}
    void describe(rtb::Flattener &F, rtb_material &o) const override {
        o.type = RTB_MAT_LAMBERTIAN;
        o.tex[0] = F.texture_id(albedo.get());
    }
    shared_ptr<texture> albedo;
};
class metal : public material {
  public:
    metal(const color &a, double f) : albedo(a), fuzz(f < 1 ? f : 1) {}
    void describe(rtb::Flattener &, rtb_material &o) const override {
        o.type = RTB_MAT_METAL;
        for (int k = 0; k < 3; ++k)
            o.color[k] = albedo[k];
        o.fuzz = fuzz;
    }
    color albedo;
    double fuzz;
};
class dielectric : public material {
  public:
    dielectric(double index_of_refraction) : ir(index_of_refraction) {}
    void describe(rtb::Flattener &, rtb_material &o) const override {
        o.type = RTB_MAT_DIELECTRIC;
        o.ir = ir;
    }
    double ir;
};
class diffuse_light : public material {
  public:
    diffuse_light(shared_ptr<texture> a) : emit(a) {}
    diffuse_light(color c) : emit(make_shared<solid_color>(c)) {}
    void describe(rtb::Flattener &F, rtb_material &o) const override {
        o.type = RTB_MAT_DIFFUSE_LIGHT;
        o.tex[0] = F.texture_id(emit.get());
    }
    shared_ptr<texture> emit;
};
class PBRMaterial : public material {
  public:
    PBRMaterial(shared_ptr<texture> a, shared_ptr<texture> r, shared_ptr<texture> m, shared_ptr<texture> n = nullptr)
        : albedo(a), roughness(r), metallic(m), normal_map(n) {}
    void describe(rtb::Flattener &F, rtb_material &o) const override {
        o.type = RTB_MAT_PBR;
        o.tex[0] = F.texture_id(albedo.get());
        o.tex[1] = F.texture_id(roughness.get());
        o.tex[2] = F.texture_id(metallic.get());
        o.tex[3] = F.texture_id(normal_map.get());
    }
    shared_ptr<texture> albedo, roughness, metallic, normal_map;
};
class isotropic : public material { // constant_medium.h:12-26
  public:
    isotropic(color c) : albedo(make_shared<solid_color>(c)) {}
    isotropic(shared_ptr<texture> a) : albedo(a) {}
    void describe(rtb::Flattener &F, rtb_material &o) const override {
        o.type = RTB_MAT_ISOTROPIC;
        o.tex[0] = F.texture_id(albedo.get());
    }
    shared_ptr<texture> albedo;
};

inline int rtb::Flattener::texture_id(const texture *t) {
    if (!t)
        return -1;
    auto it = tex_ids.find(t);
    if (it != tex_ids.end())
        return it->second;
    rtb_texture r{};
    r.even = r.odd = r.image = r.perlin = -1;
    t->describe(*this, r); // children first (checker), so indices are final before the push
    T.textures.push_back(r);
    return tex_ids[t] = int(T.textures.size()) - 1;
}
inline int rtb::Flattener::material_id(const material *m) {
    auto it = mat_ids.find(m);
    if (it != mat_ids.end())
        return it->second;
    rtb_material r{};
    for (int k = 0; k < 4; ++k)
        r.tex[k] = -1;
    m->describe(*this, r);
    T.materials.push_back(r);
    materials.push_back(m);
    return mat_ids[m] = int(T.materials.size()) - 1;
}

// ---- src/geometry -------------------------------------------------------------------------------
class hittable {
  public:
    virtual ~hittable() = default;
    // adds this object's leaves to the flat tables (wrapper chain and flags come from F)
    virtual void flatten(rtb::Flattener &F, int flags) const = 0;
    virtual bool bounding_box(double time0, double time1, aabb &output_box) const = 0;
    // hittable.h:28-29 — answered by the GPU library's fp64 validation kernels
    bool hit(const ray &r, double t_min, double t_max, hit_record &rec) const;

  private:
    struct Probe;
    mutable std::shared_ptr<Probe> probe_;
};

class sphere : public hittable {
  public:
    sphere(point3 cen, double r, shared_ptr<material> m) : center(cen), radius(r), mat_ptr(std::move(m)) {}
    void flatten(rtb::Flattener &F, int flags) const override {
        F.add_prim(RTB_PRIM_SPHERE, mat_ptr.get(), flags, {center[0], center[1], center[2], radius});
    }
    bool bounding_box(double, double, aabb &o) const override {
        o = aabb(center - vec3(radius, radius, radius), center + vec3(radius, radius, radius));
        return true;
    }
    point3 center;
    double radius;
    shared_ptr<material> mat_ptr;
};
class moving_sphere : public hittable {
  public:
    moving_sphere() {}
    moving_sphere(point3 cen0, point3 cen1, double _time0, double _time1, double r, shared_ptr<material> m)
        : center0(cen0), center1(cen1), time0(_time0), time1(_time1), radius(r), mat_ptr(m) {}
    point3 center(double time) const { return center0 + ((time - time0) / (time1 - time0)) * (center1 - center0); }
    void flatten(rtb::Flattener &F, int flags) const override {
        F.add_prim(RTB_PRIM_MOVING_SPHERE, mat_ptr.get(), flags,
                   {center0[0], center0[1], center0[2], center1[0], center1[1], center1[2], time0, time1, radius});
    }
    bool bounding_box(double t0, double t1, aabb &o) const override {
        const vec3 r(radius, radius, radius);
        o = surrounding_box(aabb(center(t0) - r, center(t0) + r), aabb(center(t1) - r, center(t1) + r));
        return true;
    }
    point3 center0, center1;
    double time0 = 0, time1 = 1, radius = 0;
    shared_ptr<material> mat_ptr;
};
class xy_rect : public hittable {
  public:
    xy_rect() {}
    xy_rect(double _x0, double _x1, double _y0, double _y1, double _k, shared_ptr<material> mat)
        : mp(mat), x0(_x0), x1(_x1), y0(_y0), y1(_y1), k(_k) {}
    void flatten(rtb::Flattener &F, int flags) const override { F.add_prim(RTB_PRIM_XY_RECT, mp.get(), flags, {x0, x1, y0, y1, k}); }
    bool bounding_box(double, double, aabb &o) const override {
        o = aabb(point3(x0, y0, k - 0.0001), point3(x1, y1, k + 0.0001));
        return true;
    }
    shared_ptr<material> mp;
    double x0 = 0, x1 = 0, y0 = 0, y1 = 0, k = 0;
};
class xz_rect : public hittable {
  public:
    xz_rect() {}
    xz_rect(double _x0, double _x1, double _z0, double _z1, double _k, shared_ptr<material> mat)
        : mp(mat), x0(_x0), x1(_x1), z0(_z0), z1(_z1), k(_k) {}
    void flatten(rtb::Flattener &F, int flags) const override { F.add_prim(RTB_PRIM_XZ_RECT, mp.get(), flags, {x0, x1, z0, z1, k}); }
    bool bounding_box(double, double, aabb &o) const override {
        o = aabb(point3(x0, k - 0.0001, z0), point3(x1, k + 0.0001, z1));
        return true;
    }
    shared_ptr<material> mp;
    double x0 = 0, x1 = 0, z0 = 0, z1 = 0, k = 0;
};
class yz_rect : public hittable {
  public:
    yz_rect() {}
    yz_rect(double _y0, double _y1, double _z0, double _z1, double _k, shared_ptr<material> mat)
        : mp(mat), y0(_y0), y1(_y1), z0(_z0), z1(_z1), k(_k) {}
    void flatten(rtb::Flattener &F, int flags) const override { F.add_prim(RTB_PRIM_YZ_RECT, mp.get(), flags, {y0, y1, z0, z1, k}); }
    bool bounding_box(double, double, aabb &o) const override {
        o = aabb(point3(k - 0.0001, y0, z0), point3(k + 0.0001, y1, z1));
        return true;
    }
    shared_ptr<material> mp;
    double y0 = 0, y1 = 0, z0 = 0, z1 = 0, k = 0;
};

class hittable_list : public hittable {
  public:
    hittable_list() {}
    hittable_list(shared_ptr<hittable> object) { add(object); }
    void clear() { objects.clear(); }
    void add(shared_ptr<hittable> object) { objects.push_back(object); }
    void flatten(rtb::Flattener &F, int flags) const override {
        for (const auto &o : objects)
            o->flatten(F, flags);
    }
    bool bounding_box(double t0, double t1, aabb &out) const override {
        if (objects.empty())
            return false;
        aabb tmp;
        bool first = true;
        for (const auto &o : objects) {
            if (!o->bounding_box(t0, t1, tmp))
                return false;
            out = first ? tmp : surrounding_box(out, tmp);
            first = false;
        }
        return true;
    }
    std::vector<shared_ptr<hittable>> objects;
};

class box : public hittable { // box.h:31-47
  public:
    box() {}
    box(const point3 &p0, const point3 &p1, shared_ptr<material> ptr) : box_min(p0), box_max(p1) {
        sides.add(make_shared<xy_rect>(p0.x(), p1.x(), p0.y(), p1.y(), p1.z(), ptr));
        sides.add(make_shared<xy_rect>(p0.x(), p1.x(), p0.y(), p1.y(), p0.z(), ptr));
        sides.add(make_shared<xz_rect>(p0.x(), p1.x(), p0.z(), p1.z(), p1.y(), ptr));
        sides.add(make_shared<xz_rect>(p0.x(), p1.x(), p0.z(), p1.z(), p0.y(), ptr));
        sides.add(make_shared<yz_rect>(p0.y(), p1.y(), p0.z(), p1.z(), p1.x(), ptr));
        sides.add(make_shared<yz_rect>(p0.y(), p1.y(), p0.z(), p1.z(), p0.x(), ptr));
    }
    void flatten(rtb::Flattener &F, int flags) const override { sides.flatten(F, flags); }
    bool bounding_box(double, double, aabb &o) const override {
        o = aabb(box_min, box_max);
        return true;
    }
    point3 box_min, box_max;
    hittable_list sides;
};

// bvh.h: the reference builds its random-axis median-split tree here; the device library
// builds its own SAH tree from the flat tables, so this class only keeps the members.
// A node over ONE object tests it twice in the reference (bvh.h:68-69): flagged DUP_LEAF.
class bvh_node : public hittable {
  public:
    bvh_node(const hittable_list &list, double time0, double time1) : objects(list.objects), t0(time0), t1(time1) {}
    bvh_node(const std::vector<shared_ptr<hittable>> &src, size_t start, size_t end, double time0, double time1)
        : objects(src.begin() + start, src.begin() + end), t0(time0), t1(time1) {}
    void flatten(rtb::Flattener &F, int flags) const override {
        const int f = flags | (objects.size() == 1 ? int(RTB_PRIM_DUP_LEAF) : 0);
        for (const auto &o : objects)
            o->flatten(F, f);
    }
    bool bounding_box(double time0, double time1, aabb &out) const override {
        hittable_list l;
        l.objects = objects;
        return l.bounding_box(time0, time1, out);
    }
    std::vector<shared_ptr<hittable>> objects;
    double t0, t1;
};

class translate : public hittable {
  public:
    translate(shared_ptr<hittable> p, const vec3 &displacement) : ptr(p), offset(displacement) {}
    void flatten(rtb::Flattener &F, int flags) const override {
        F.push(this, RTB_XF_TRANSLATE, offset[0], offset[1], offset[2]);
        ptr->flatten(F, flags);
        F.pop();
    }
    bool bounding_box(double t0, double t1, aabb &o) const override {
        if (!ptr->bounding_box(t0, t1, o))
            return false;
        o = aabb(o.min() + offset, o.max() + offset);
        return true;
    }
    shared_ptr<hittable> ptr;
    vec3 offset;
};
class rotate_y : public hittable {
  public:
    rotate_y(shared_ptr<hittable> p, double angle) : ptr(p) {
        auto radians = degrees_to_radians(angle);
        sin_theta = sin(radians);
        cos_theta = cos(radians);
        hasbox = ptr->bounding_box(0, 1, bbox);
        // bounds of the rotated box = bounds of its eight rotated corners (y is untouched by the rotation)
        const point3 lo = bbox.min(), hi = bbox.max();
        double x0 = infinity, x1 = -infinity, z0 = infinity, z1 = -infinity;
        for (int corner = 0; corner < 4; ++corner) {
            const double x = (corner & 1) ? hi.x() : lo.x(), z = (corner & 2) ? hi.z() : lo.z();
            const double rx = cos_theta * x + sin_theta * z, rz = -sin_theta * x + cos_theta * z;
            x0 = fmin(x0, rx);
            x1 = fmax(x1, rx);
            z0 = fmin(z0, rz);
            z1 = fmax(z1, rz);
        }
        bbox = aabb(point3(x0, lo.y(), z0), point3(x1, hi.y(), z1));
    }
    void flatten(rtb::Flattener &F, int flags) const override {
        F.push(this, RTB_XF_ROTATE_Y, sin_theta, cos_theta, 0);
        ptr->flatten(F, flags);
        F.pop();
    }
    bool bounding_box(double, double, aabb &o) const override {
        o = bbox;
        return hasbox;
    }
    shared_ptr<hittable> ptr;
    double sin_theta, cos_theta;
    bool hasbox;
    aabb bbox;
};
class flip_face : public hittable {
  public:
    flip_face(shared_ptr<hittable> p) : ptr(p) {}
    void flatten(rtb::Flattener &F, int flags) const override {
        F.push(this, RTB_XF_FLIP_FACE);
        ptr->flatten(F, flags);
        F.pop();
    }
    bool bounding_box(double t0, double t1, aabb &o) const override { return ptr->bounding_box(t0, t1, o); }
    shared_ptr<hittable> ptr;
};

class constant_medium : public hittable {
  public:
    constant_medium(shared_ptr<hittable> b, double d, shared_ptr<texture> a)
        : boundary(b), phase_function(make_shared<isotropic>(a)), neg_inv_density(-1 / d) {}
    constant_medium(shared_ptr<hittable> b, double d, color c)
        : boundary(b), phase_function(make_shared<isotropic>(c)), neg_inv_density(-1 / d) {}
    void flatten(rtb::Flattener &F, int flags) const override {
        if (F.boundary)
            throw std::runtime_error("constant_medium inside a constant_medium boundary is not supported");
        const int first = int(F.T.prims.size());
        F.boundary = true;
        boundary->flatten(F, 0);
        F.boundary = false;
        const int count = int(F.T.prims.size()) - first;
        const int id = F.add_prim(RTB_PRIM_MEDIUM, phase_function.get(), flags, {neg_inv_density});
        F.T.prims[id].aux0 = first;
        F.T.prims[id].aux1 = count;
    }
    bool bounding_box(double t0, double t1, aabb &o) const override { return boundary->bounding_box(t0, t1, o); }
    shared_ptr<hittable> boundary;
    shared_ptr<material> phase_function;
    double neg_inv_density;
};

// ---- src/lighting ---------------------------------------------------------------------------------
struct LightSample { // light.h:7-13
    color Li;
    vec3 wi;
    double pdf = 0;
    double dist = 0;
    bool is_delta = false;
};

class Light {
  public:
    virtual ~Light() = default;
    virtual void describe(rtb::SceneTables &T, rtb_light &out) const = 0;
    virtual bool is_delta() const { return false; }
    virtual bool is_infinite() const { return false; }
    virtual color power() const { return color(0, 0, 0); }
    // light.h:21-41 — answered by the GPU library
    LightSample sample(const point3 &p, const vec2 &u) const;
    double pdf(const point3 &origin, const vec3 &direction) const;
    color Le(const ray &r) const;
};

class QuadLight : public Light {
  public:
    QuadLight(const point3 &_Q, const vec3 &_u, const vec3 &_v, const color &_c) : Q(_Q), u(_u), v(_v), intensity(_c) {}
    void describe(rtb::SceneTables &, rtb_light &o) const override {
        o.type = RTB_LIGHT_QUAD;
        for (int k = 0; k < 3; ++k) {
            o.Q[k] = Q[k];
            o.u[k] = u[k];
            o.v[k] = v[k];
            o.intensity[k] = intensity[k];
        }
    }
    point3 Q;
    vec3 u, v;
    color intensity;
};
class PointLight : public Light {
  public:
    PointLight(const point3 &pos, const color &intensity) : m_position(pos), m_intensity(intensity) {}
    void describe(rtb::SceneTables &, rtb_light &o) const override {
        o.type = RTB_LIGHT_POINT;
        for (int k = 0; k < 3; ++k) {
            o.Q[k] = m_position[k];
            o.intensity[k] = m_intensity[k];
        }
    }
    bool is_delta() const override { return true; }
    color power() const override { return 4.0 * pi * m_intensity; }
    point3 m_position;
    color m_intensity;
};
class SpotLight : public Light {
  public:
    SpotLight(point3 pos, vec3 dir, double cutoff, color intensity_)
        : position(pos), direction(unit_vector(dir)), intensity(intensity_), cos_cutoff(cos(cutoff * (pi / 180.0))) {}
    void describe(rtb::SceneTables &, rtb_light &o) const override {
        o.type = RTB_LIGHT_SPOT;
        for (int k = 0; k < 3; ++k) {
            o.Q[k] = position[k];
            o.u[k] = direction[k];
            o.intensity[k] = intensity[k];
        }
        o.cos_cutoff = cos_cutoff;
    }
    bool is_delta() const override { return true; }
    point3 position;
    vec3 direction;
    color intensity;
    double cos_cutoff;
};
class DirectionalLight : public Light {
  public:
    DirectionalLight(const vec3 &dir, const color &c) : direction(unit_vector(dir)), L(c) {}
    void describe(rtb::SceneTables &, rtb_light &o) const override {
        o.type = RTB_LIGHT_DIRECTIONAL;
        for (int k = 0; k < 3; ++k) {
            o.u[k] = direction[k];
            o.intensity[k] = L[k];
        }
    }
    bool is_delta() const override { return true; }
    vec3 direction;
    color L;
};

// environmental_light.h:114-144.  Reads Radiance .hdr (RGBE, flat or new-style RLE); a file
// that cannot be read gives the reference's fallback: constant white, uniform sampling.
class EnvironmentLight : public Light {
  public:
    EnvironmentLight(const char *map_filename) {
        if (!load_hdr(map_filename)) {
            std::cerr << "ERROR: Could not load HDR environment map: " << map_filename << std::endl;
            width = height = 0;
            hdr_data.clear();
        }
    }
    EnvironmentLight(int w, int h, const float *rgb) : hdr_data(rgb, rgb + size_t(w) * h * 3), width(w), height(h) {}
    void describe(rtb::SceneTables &T, rtb_light &o) const override {
        o.type = RTB_LIGHT_ENV;
        o.env_width = width;
        o.env_height = height;
        o.env_is_probe = (width > 0 && width == height) ? 1 : 0; // environmental_light.h:138-140
        o.env_offset = T.env_texels.size();
        T.env_texels.insert(T.env_texels.end(), hdr_data.begin(), hdr_data.end());
    }
    bool is_infinite() const override { return true; }
    std::vector<float> hdr_data;
    int width = 0, height = 0;

  private:
    bool load_hdr(const char *path) {
        std::ifstream f(path, std::ios::binary);
        if (!f)
            return false;
        std::string line;
        if (!std::getline(f, line) || (line.rfind("#?RADIANCE", 0) != 0 && line.rfind("#?RGBE", 0) != 0))
            return false;
        while (std::getline(f, line) && !line.empty()) {
        }
        if (!std::getline(f, line))
            return false;
        int w = 0, h = 0;
        if (std::sscanf(line.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0)
            return false;
        std::vector<unsigned char> scan(size_t(w) * 4);
        hdr_data.assign(size_t(w) * h * 3, 0.f);
        for (int y = 0; y < h; ++y) {
            unsigned char hd[4];
            if (!f.read(reinterpret_cast<char *>(hd), 4))
                return false;
            if (w >= 8 && w < 32768 && hd[0] == 2 && hd[1] == 2 && !(hd[2] & 0x80)) { // new-style RLE
                for (int c = 0; c < 4; ++c)
                    for (int x = 0; x < w;) {
                        int n = f.get();
                        if (n < 0)
                            return false;
                        if (n > 128) {
                            int v = f.get();
                            for (n -= 128; n-- > 0 && x < w; ++x)
                                scan[size_t(x) * 4 + c] = static_cast<unsigned char>(v);
                        } else {
                            for (; n-- > 0 && x < w; ++x)
                                scan[size_t(x) * 4 + c] = static_cast<unsigned char>(f.get());
                        }
                    }
            } else { // flat
                std::memcpy(scan.data(), hd, 4);
                if (!f.read(reinterpret_cast<char *>(scan.data()) + 4, std::streamsize(size_t(w) * 4 - 4)))
                    return false;
            }
            for (int x = 0; x < w; ++x) {
                const unsigned char *p = &scan[size_t(x) * 4];
                const float s = p[3] ? std::ldexp(1.0f, int(p[3]) - (128 + 8)) : 0.f;
                for (int c = 0; c < 3; ++c)
                    hdr_data[(size_t(y) * w + x) * 3 + c] = p[c] * s;
            }
        }
        width = w;
        height = h;
        return true;
    }
};

// ---- src/renderer/camera.h ----------------------------------------------------------------------
class camera {
  public:
    camera(point3 lookfrom, point3 lookat, point3 vup, double vfov, double aspect_ratio, double aperture,
           double focus_dist, double _time0 = 0.0, double _time1 = 0.0) {
        for (int k = 0; k < 3; ++k) {
            args.lookfrom[k] = lookfrom[k];
            args.lookat[k] = lookat[k];
            args.vup[k] = vup[k];
        }
        args.vfov = vfov;
        args.aspect_ratio = aspect_ratio;
        args.aperture = aperture;
        args.focus_dist = focus_dist;
        args.time0 = _time0;
        args.time1 = _time1;
    }
    rtb_camera args; // the constructor arguments; the derived frame is built by the device library
};

// ---- src/scene/scenes.h ---------------------------------------------------------------------------
#ifndef RTB_HOST_NO_SCENECONFIG
struct RtbSceneConfig {
    shared_ptr<hittable> world;
    std::vector<shared_ptr<Light>> lights;
    color background{0, 0, 0};
    point3 lookfrom{13, 2, 3};
    point3 lookat{0, 0, 0};
    vec3 vup{0, 1, 0};
    double vfov = 40.0;
    double aperture = 0.0;
    double focus_dist = 10.0;
    double aspect_ratio = 16.0 / 9.0;
    int image_width = 1280;
    int samples_per_pixel = 100;
};
#endif

namespace rtb {

// world + camera + background + lights -> scene blob (what Renderer::render receives,
// renderer.h:30-32).  `materials_out` (optional) receives index -> material*.
inline std::vector<uint8_t> flatten(const hittable &world, const camera &cam, const color &background,
                                    const std::vector<shared_ptr<Light>> &lights, int width, int height, int spp,
                                    int scene_id = -1, std::vector<const material *> *materials_out = nullptr) {
    Flattener F;
    world.flatten(F, 0);
    for (const auto &l : lights) {
        rtb_light r{};
        l->describe(F.T, r);
        F.T.lights.push_back(r);
    }
    for (int k = 0; k < 3; ++k)
        F.T.globals.background[k] = background[k];
    F.T.globals.image_width = width;
    F.T.globals.image_height = height;
    F.T.globals.samples_per_pixel = spp;
    F.T.globals.scene_id = scene_id;
    F.T.camera = cam.args;
    if (materials_out)
        *materials_out = F.materials;
    return F.T.serialise();
}

inline camera default_probe_camera() { return camera(point3(0, 0, 1), point3(0, 0, 0), vec3(0, 1, 0), 40, 1, 0, 1); }

} // namespace rtb

// ---- GPU-backed single-query API ---------------------------------------------------------------
struct hittable::Probe {
    std::vector<const material *> materials;
};

inline bool hittable::hit(const ray &r, double t_min, double t_max, hit_record &rec) const {
    rtb_context *ctx = rtb::api_context();
    static thread_local const hittable *uploaded = nullptr;
    if (!probe_ || uploaded != this) { // (re)upload this object as the world of the API context
        auto pr = std::make_shared<Probe>();
        const auto blob = rtb::flatten(*this, rtb::default_probe_camera(), color(0, 0, 0), {}, 2, 2, 1, -1, &pr->materials);
        rtb::check(rtb_scene_upload(ctx, blob.data(), blob.size()), ctx, "rtb_scene_upload");
        probe_ = pr;
        uploaded = this;
    }
    rtb_ray q{};
    for (int k = 0; k < 3; ++k) {
        q.o[k] = r.origin()[k];
        q.d[k] = r.direction()[k];
    }
    q.time = r.time();
    q.t_min = t_min;
    q.t_max = t_max;
    q.origin_prim = -1;
    rtb_hit h{};
    rtb::check(rtb_trace_batch(ctx, &q, 1, 64, &h, nullptr), ctx, "rtb_trace_batch");
    if (h.prim < 0)
        return false;
    rec.t = h.t;
    rec.p = point3(h.p[0], h.p[1], h.p[2]);
    rec.normal = vec3(h.normal[0], h.normal[1], h.normal[2]);
    rec.u = h.u;
    rec.v = h.v;
    rec.front_face = h.front_face != 0;
    rec.mat_ptr = const_cast<material *>(probe_->materials[size_t(h.material)]);
    return true;
}

namespace rtb {
// Uploads a one-sphere scene carrying `m` (material 0) / `l` (light 0) / `t` (last texture).
inline void upload_probe_scene(const shared_ptr<material> &m, const std::vector<shared_ptr<Light>> &lights) {
    rtb_context *ctx = api_context();
    sphere s(point3(0, 0, 0), 1.0, m);
    const auto blob = flatten(s, default_probe_camera(), color(0, 0, 0), lights, 2, 2, 1);
    check(rtb_scene_upload(ctx, blob.data(), blob.size()), ctx, "rtb_scene_upload");
}
struct NoDelete {
    template <class T> void operator()(T *) const {}
};
inline rtb_bsdf_query make_query(const hit_record &rec, const vec3 &wo, const vec3 &wi) {
    rtb_bsdf_query q{};
    for (int k = 0; k < 3; ++k) {
        q.p[k] = rec.p[k];
        q.normal[k] = rec.normal[k];
        q.wo[k] = wo[k];
        q.wi[k] = wi[k];
    }
    q.u = rec.u;
    q.v = rec.v;
    q.front_face = rec.front_face ? 1 : 0;
    return q;
}
inline rtb_bsdf_value bsdf_value(const material *m, const hit_record &rec, const vec3 &wo, const vec3 &wi) {
    upload_probe_scene(shared_ptr<material>(const_cast<material *>(m), NoDelete()), {});
    rtb_context *ctx = api_context();
    const rtb_bsdf_query q = make_query(rec, wo, wi);
    rtb_bsdf_value v{};
    check(rtb_bsdf_eval_batch(ctx, 0, &q, 1, 64, &v), ctx, "rtb_bsdf_eval_batch");
    return v;
}
inline rtb_bsdf_sample bsdf_sample(const material *m, const hit_record &rec, const vec3 &wo) {
    upload_probe_scene(shared_ptr<material>(const_cast<material *>(m), NoDelete()), {});
    rtb_context *ctx = api_context();
    const rtb_bsdf_query q = make_query(rec, wo, wo);
    rtb_bsdf_sample s{};
    const uint64_t seed = uint64_t(random_double() * 4294967296.0) + 1;
    check(rtb_bsdf_sample_batch(ctx, 0, &q, 1, 64, seed, &s), ctx, "rtb_bsdf_sample_batch");
    return s;
}
} // namespace rtb

inline color material::eval(const hit_record &rec, const vec3 &wo, const vec3 &wi) const {
    const rtb_bsdf_value v = rtb::bsdf_value(this, rec, wo, wi);
    return color(v.f[0], v.f[1], v.f[2]);
}
inline double material::pdf(const hit_record &rec, const vec3 &wo, const vec3 &wi) const {
    return rtb::bsdf_value(this, rec, wo, wi).pdf;
}
inline color material::emitted(double u, double v, const point3 &p) const {
    hit_record rec;
    rec.u = u;
    rec.v = v;
    rec.p = p;
    rec.normal = vec3(0, 1, 0);
    const rtb_bsdf_value r = rtb::bsdf_value(this, rec, vec3(0, 1, 0), vec3(0, 1, 0));
    return color(r.emitted_old[0], r.emitted_old[1], r.emitted_old[2]);
}
inline color material::emitted(const hit_record &rec, const vec3 &wo) const {
    const rtb_bsdf_value r = rtb::bsdf_value(this, rec, wo, wo);
    return color(r.emitted_new[0], r.emitted_new[1], r.emitted_new[2]);
}
inline bool material::sample(const hit_record &rec, const vec3 &wo, BSDFSample &sampled) const {
    const rtb_bsdf_sample s = rtb::bsdf_sample(this, rec, wo);
    if (!s.ok)
        return false;
    sampled.wi = vec3(s.wi[0], s.wi[1], s.wi[2]);
    sampled.f = color(s.f[0], s.f[1], s.f[2]);
    sampled.pdf = s.pdf;
    sampled.is_specular = s.is_specular != 0;
    return true;
}
inline bool material::scatter(const ray &r_in, const hit_record &rec, color &attenuation, ray &scattered) const {
    const rtb_bsdf_sample s = rtb::bsdf_sample(this, rec, -unit_vector(r_in.direction()));
    if (!s.scatter_ok)
        return false;
    attenuation = color(s.scatter_atten[0], s.scatter_atten[1], s.scatter_atten[2]);
    scattered = ray(rec.p, vec3(s.scatter_dir[0], s.scatter_dir[1], s.scatter_dir[2]), r_in.time());
    return true;
}
inline color texture::value(double u, double v, const point3 &p) const {
    auto self = shared_ptr<texture>(const_cast<texture *>(this), rtb::NoDelete());
    rtb::upload_probe_scene(make_shared<lambertian>(self), {});
    rtb_context *ctx = rtb::api_context();
    const double q[5] = {u, v, p[0], p[1], p[2]};
    double rgb[3] = {0, 0, 0};
    // replaying the flattener on this texture alone yields its index in the uploaded scene
    rtb::Flattener F;
    const int tid = F.texture_id(this);
    rtb::check(rtb_texture_eval_batch(ctx, tid, q, 1, 64, rgb), ctx, "rtb_texture_eval_batch");
    return color(rgb[0], rgb[1], rgb[2]);
}
namespace rtb {
inline rtb_light_value light_value(const Light *l, const point3 &p, const vec3 &d, const vec2 &u) {
    upload_probe_scene(make_shared<lambertian>(color(0.5, 0.5, 0.5)), {shared_ptr<Light>(const_cast<Light *>(l), NoDelete())});
    rtb_context *ctx = api_context();
    rtb_light_query q{};
    for (int k = 0; k < 3; ++k) {
        q.p[k] = p[k];
        q.d[k] = d[k];
    }
    q.u[0] = u.x();
    q.u[1] = u.y();
    rtb_light_value v{};
    check(rtb_light_eval_batch(ctx, 0, &q, 1, 64, uint64_t(random_double() * 4294967296.0) + 1, &v), ctx,
          "rtb_light_eval_batch");
    return v;
}
} // namespace rtb
inline LightSample Light::sample(const point3 &p, const vec2 &u) const {
    const rtb_light_value v = rtb::light_value(this, p, vec3(0, 1, 0), u);
    LightSample s;
    s.Li = color(v.Li[0], v.Li[1], v.Li[2]);
    s.wi = vec3(v.wi[0], v.wi[1], v.wi[2]);
    s.pdf = v.pdf;
    s.dist = v.dist;
    s.is_delta = v.is_delta != 0;
    return s;
}
inline double Light::pdf(const point3 &origin, const vec3 &direction) const {
    return rtb::light_value(this, origin, direction, vec2(0.5, 0.5)).pdf_dir;
}
inline color Light::Le(const ray &r) const {
    const rtb_light_value v = rtb::light_value(this, r.origin(), r.direction(), vec2(0.5, 0.5));
    return color(v.Le[0], v.Le[1], v.Le[2]);
}

// ---- src/renderer ---------------------------------------------------------------------------------
// integrator.h:9-21 and the five implementations: on the GPU an integrator is its id
// (src/main.cpp:81-100) plus max depth and Russian-roulette start depth.
class Integrator {
  public:
    virtual ~Integrator() = default;
    virtual int id() const = 0;
    virtual void set_max_depth(int depth) { m_max_depth = depth; }
    void set_rr_start_depth(int depth) { m_rr_start_depth = depth; }
    int max_depth() const { return m_max_depth; }
    int rr_start_depth() const { return m_rr_start_depth; }

  protected:
    int m_max_depth = 50;
    int m_rr_start_depth = 3;
};
class PathIntegrator : public Integrator { public: int id() const override { return RTB_INTEGRATOR_PATH; } };
class RRPathInterator : public Integrator { public: int id() const override { return RTB_INTEGRATOR_RR; } };
class PBRPathIntegrator : public Integrator { public: int id() const override { return RTB_INTEGRATOR_PBR; } };
class DirectLightIntegrator : public Integrator { public: int id() const override { return RTB_INTEGRATOR_DIRECT; } };
class MISPathIntegrator : public Integrator { public: int id() const override { return RTB_INTEGRATOR_MIS; } };

// render_buffer.h:11-84
class RenderBuffer {
  public:
    RenderBuffer(int width, int height) : m_width(width), m_height(height) { m_pixels.resize(height, std::vector<color>(width)); }
    void set_pixel(int x, int y, const color &c) {
        if (x >= 0 && x < m_width && y >= 0 && y < m_height)
            m_pixels[y][x] = c;
    }
    const std::vector<std::vector<color>> &get_data() const { return m_pixels; }
    int width() const { return m_width; }
    int height() const { return m_height; }
    // render_buffer.h:35-55: y flip, (unsigned char)(x * 255)
    std::vector<unsigned char> to_rgb8() const {
        std::vector<unsigned char> img(size_t(m_width) * m_height * 3);
        for (int j = 0; j < m_height; ++j)
            for (int i = 0; i < m_width; ++i)
                for (int k = 0; k < 3; ++k)
                    img[(size_t(j) * m_width + i) * 3 + k] = static_cast<unsigned char>(m_pixels[m_height - 1 - j][i][k] * 255);
        return img;
    }
    // Uncompressed (stored-deflate) PNG; the reference goes through the vendored stb_image_write.
    bool save_to_png(const std::string &filename) const {
        const auto img = to_rgb8();
        std::vector<unsigned char> raw;
        raw.reserve(img.size() + m_height);
        for (int j = 0; j < m_height; ++j) {
            raw.push_back(0);
            raw.insert(raw.end(), img.begin() + size_t(j) * m_width * 3, img.begin() + size_t(j + 1) * m_width * 3);
        }
        std::vector<unsigned char> z = {0x78, 0x01};
        for (size_t pos = 0; pos < raw.size();) {
            const size_t n = std::min<size_t>(65535, raw.size() - pos);
            z.push_back(pos + n == raw.size() ? 1 : 0);
            z.push_back(n & 255);
            z.push_back(n >> 8);
            z.push_back(~n & 255);
            z.push_back((~n >> 8) & 255);
            z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
            pos += n;
        }
        uint32_t a = 1, b = 0;
        for (unsigned char c : raw) {
            a = (a + c) % 65521;
            b = (b + a) % 65521;
        }
        const uint32_t adler = (b << 16) | a;
        for (int s = 24; s >= 0; s -= 8)
            z.push_back((adler >> s) & 255);
        std::ofstream f(filename, std::ios::binary);
        if (!f)
            return false;
        const unsigned char sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
        f.write(reinterpret_cast<const char *>(sig), 8);
        auto chunk = [&](const char *tag, const std::vector<unsigned char> &d) {
            std::vector<unsigned char> buf(tag, tag + 4);
            buf.insert(buf.end(), d.begin(), d.end());
            uint32_t crc = 0xffffffffu;
            for (unsigned char c : buf) {
                crc ^= c;
                for (int k = 0; k < 8; ++k)
                    crc = (crc >> 1) ^ (0xedb88320u & (0u - (crc & 1)));
            }
            crc ^= 0xffffffffu;
            const uint32_t len = uint32_t(d.size());
            const unsigned char l[4] = {(unsigned char)(len >> 24), (unsigned char)(len >> 16), (unsigned char)(len >> 8), (unsigned char)len};
            const unsigned char c4[4] = {(unsigned char)(crc >> 24), (unsigned char)(crc >> 16), (unsigned char)(crc >> 8), (unsigned char)crc};
            f.write(reinterpret_cast<const char *>(l), 4);
            f.write(reinterpret_cast<const char *>(buf.data()), std::streamsize(buf.size()));
            f.write(reinterpret_cast<const char *>(c4), 4);
        };
        std::vector<unsigned char> ihdr(13, 0);
        for (int s = 0; s < 4; ++s) {
            ihdr[s] = (m_width >> (24 - 8 * s)) & 255;
            ihdr[4 + s] = (m_height >> (24 - 8 * s)) & 255;
        }
        ihdr[8] = 8;
        ihdr[9] = 2;
        chunk("IHDR", ihdr);
        chunk("IDAT", z);
        chunk("IEND", {});
        return bool(f);
    }

  private:
    int m_width, m_height;
    std::vector<std::vector<color>> m_pixels;
};

// renderer.h:17-140
class Renderer {
  public:
    struct Settings {
        int samples_per_pixel = 10;
    };
    Renderer() : m_is_rendering(false) {}
    ~Renderer() {
        if (m_group)
            rtb_group_destroy(m_group);
        else if (m_ctx)
            rtb_context_destroy(m_ctx);
    }
    Renderer(const Renderer &) = delete;
    Renderer &operator=(const Renderer &) = delete;

    void set_device(int device) { m_devices.assign(1, device); }
    // Several GPUs of this box: the samples of every pixel (rows too when there are fewer samples
    // than GPUs) are split over them and reduced over NCCL inside the library (rtb_group_*) — the
    // GPU counterpart of the reference's tile queue over CPU threads (renderer.h:40-94).
    void set_devices(const std::vector<int> &devices) {
        if (!devices.empty())
            m_devices = devices;
    }
    int device_count() const { return int(m_devices.size()); }
    void set_integrator(std::shared_ptr<Integrator> integrator) { m_integrator = integrator; }
    void set_samples(int samples) { m_settings.samples_per_pixel = samples; }
    void set_max_depth(int depth) {
        if (m_integrator)
            m_integrator->set_max_depth(depth);
    }
    void set_seed(uint64_t seed) { m_seed = seed; }
    // Progressive preview: how many sample passes the render is cut into (0 = automatic, about
    // 2^28 samples per pass; 1 = one call).  target_buffer is refreshed after every pass.
    void set_preview_passes(int passes) { m_preview_passes = passes; }
    int passes_done() const { return m_passes_done; }
    void cancel() { // renderer.h:113 — safe from another thread
        m_is_rendering = false;
        if (m_group)
            rtb_group_cancel(m_group);
        else if (m_ctx)
            rtb_cancel(m_ctx);
    }
    bool is_rendering() const { return m_is_rendering; }
    const rtb_render_stats &last_stats() const { return m_stats; }

    // renderer.h:30-102.  Throws std::runtime_error when there is no GPU (there is no CPU path).
    void render(shared_ptr<hittable> world, shared_ptr<camera> cam, const color &background,
                RenderBuffer &target_buffer, const std::vector<shared_ptr<Light>> &lights = {}) {
        m_is_rendering = true;
        const auto start = std::chrono::high_resolution_clock::now();
        const bool multi = m_devices.size() > 1;
        if (multi && !m_group) {
            if (rtb_group_create(m_devices.data(), int(m_devices.size()), &m_group) != RTB_OK)
                throw std::runtime_error(std::string("rtb_group_create: ") + rtb_last_error(nullptr));
            m_ctx = rtb_group_context(m_group, 0);
        }
        if (!m_ctx)
            rtb::check(rtb_context_create(m_devices[0], &m_ctx), nullptr, "rtb_context_create");
        const int w = target_buffer.width(), h = target_buffer.height();
        const auto blob = rtb::flatten(*world, *cam, background, lights, w, h, m_settings.samples_per_pixel);
        if (multi) {
            if (rtb_group_scene_upload(m_group, blob.data(), blob.size()) != RTB_OK)
                throw std::runtime_error(std::string("rtb_group_scene_upload: ") + rtb_group_last_error(m_group));
        } else {
            rtb::check(rtb_scene_upload(m_ctx, blob.data(), blob.size()), m_ctx, "rtb_scene_upload");
        }
        if (!m_integrator) { // renderer.h:76-79: without an integrator nothing is accumulated
            m_is_rendering = false;
            return;
        }
        rtb_render_params p{};
        p.width = w;
        p.height = h;
        p.spp = m_settings.samples_per_pixel;
        p.max_depth = m_integrator->max_depth();
        p.rr_start_depth = m_integrator->rr_start_depth();
        p.integrator = m_integrator->id();
        p.seed = m_seed;
        // The reference's worker threads fill target_buffer tile by tile while the UI thread polls
        // it every 33 ms (main.cpp:119-126).  Here the image fills in sample PASSES: pass k renders
        // the samples s = k (mod passes) of every pixel (the same split the multi-GPU driver uses;
        // a sample's random stream depends on (pixel, s, seed) only, so the passes add up to exactly
        // the one-call sample set) and the buffer shows sqrt(mean so far) after each.
        int passes = m_preview_passes;
        if (passes <= 0) { // automatic: about 2^28 samples (tens of ms on a B200) per pass
            const unsigned long long total = (unsigned long long)w * h * (unsigned long long)std::max(p.spp, 1);
            passes = int(std::min<unsigned long long>((total + (1ull << 28) - 1) >> 28, 64));
        }
        passes = std::max(1, std::min(passes, std::max(p.spp, 1)));
        p.sample_stride = passes;
        std::vector<float> acc(size_t(w) * h * 4), sum;
        rtb_render_stats total_stats{};
        int samples_done = 0;
        m_passes_done = 0;
        for (int pass = 0; pass < passes && m_is_rendering; ++pass) {
            p.sample_offset = pass;
            int rc;
            if (multi) { // the group plans its own split: a pass is a job of its own (spp / passes samples, its own seed)
                rtb_render_params q = p;
                q.sample_offset = 0;
                q.sample_stride = 1;
                q.spp = (p.spp - pass + passes - 1) / passes;
                q.seed = m_seed + 0x9e3779b97f4a7c15ull * uint64_t(pass);
                rc = rtb_group_render(m_group, &q, acc.data(), nullptr, &m_stats);
            } else {
                rc = rtb_render(m_ctx, &p, acc.data(), &m_stats);
            }
            if (rc == RTB_ERR_CANCELLED)
                break;
            if (rc != RTB_OK) {
                m_is_rendering = false;
                if (multi)
                    throw std::runtime_error(std::string("rtb_group_render: ") + rtb_group_last_error(m_group));
                rtb::check(rc, m_ctx, "rtb_render");
            }
            samples_done += (p.spp - pass + passes - 1) / passes;
            total_stats.paths += m_stats.paths;
            total_stats.rays_closest += m_stats.rays_closest;
            total_stats.rays_shadow += m_stats.rays_shadow;
            total_stats.iterations += m_stats.iterations;
            total_stats.kernel_launches += m_stats.kernel_launches;
            total_stats.device_ms += m_stats.device_ms;
            total_stats.schedule = m_stats.schedule;
            const float *show = acc.data();
            if (passes > 1) {
                if (sum.empty())
                    sum.assign(acc.size(), 0.f);
                for (size_t i = 0; i < acc.size(); ++i)
                    sum[i] += acc[i];
                show = sum.data();
            }
            const double scale = 1.0 / std::max(samples_done, 1);
            for (int j = 0; j < h; ++j) // renderer.h:126-140
                for (int i = 0; i < w; ++i) {
                    const float *a = &show[(size_t(j) * w + i) * 4];
                    target_buffer.set_pixel(i, j, color(clamp(sqrt(scale * a[0]), 0.0, 1.0), clamp(sqrt(scale * a[1]), 0.0, 1.0),
                                                        clamp(sqrt(scale * a[2]), 0.0, 1.0)));
                }
            m_passes_done = pass + 1;
        }
        m_stats = total_stats;
        const std::chrono::duration<double> elapsed = std::chrono::high_resolution_clock::now() - start;
        m_is_rendering = false;
        std::cout << "Rendering finished in " << elapsed.count() << " seconds." << std::endl; // renderer.h:100
    }

  private:
    Settings m_settings;
    std::atomic<bool> m_is_rendering;
    std::shared_ptr<Integrator> m_integrator;
    rtb_context *m_ctx = nullptr; // the single device's context, or the group's first
    rtb_group *m_group = nullptr; // several devices
    rtb_render_stats m_stats{};
    std::vector<int> m_devices = std::vector<int>(1, 0);
    uint64_t m_seed = 1;
    int m_preview_passes = 0;
    std::atomic<int> m_passes_done{0};
};

#endif // RTB_HOST_HPP
