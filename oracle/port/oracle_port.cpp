// oracle/port/oracle_port.cpp — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's path-integration algorithm on the flat scene tables of
// include/rtb200_scene.h: fp64, the reference's operation order, its rejection samplers and
// its xorshift32 generator (with an explicit seed).  Deliberately independent of the product:
// it shares no code with ray_tracing-rendering_b200/csrc (no BVH — every ray is tested
// against every primitive, closest hit wins, which is what any correct BVH must return).
// Every function cites the reference lines it restates (paths relative to /root/reference/src).
//
// PINNED: validated against the golden vectors the unmodified reference produced
// (tests/golden/*.npz, via tests/test_oracle_port.py): hits bit-exact, BSDF / light /
// texture values to the last ulp, images statistically.  It is the checker of record on
// machines where oracle/_ref (the compiled reference itself) is not available.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
#include "rtb200_blob.hpp"
#include "rtb200_types.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

namespace {

constexpr double kInf = std::numeric_limits<double>::infinity();
constexpr double kPi = 3.1415926535897932385; // core/rtweekend.h:18

struct V {
    double x, y, z;
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V operator+(V a, V b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V operator-(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V operator-(V a) { return {-a.x, -a.y, -a.z}; }
inline V operator*(V a, V b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V operator*(double t, V a) { return {t * a.x, t * a.y, t * a.z}; }
inline V operator*(V a, double t) { return t * a; }
inline V operator/(V a, double t) { return (1 / t) * a; }                             // core/vec3.h:208-210
inline double dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }             // vec3.h:212-214
inline V cross(V u, V v) { return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; }
inline double len2(V a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
inline double len(V a) { return std::sqrt(len2(a)); }
inline V unit(V a) { return a / len(a); }                                            // vec3.h:222-224
inline V reflect(V v, V n) { return v - 2 * dot(v, n) * n; }                          // vec3.h:239-241
inline V refract(V uv, V n, double eta) {                                            // vec3.h:243-248
    const double cos_theta = std::fmin(dot(-uv, n), 1.0);
    const V perp = eta * (uv + cos_theta * n);
    const V par = -std::sqrt(std::fabs(1.0 - len2(perp))) * n;
    return perp + par;
}
inline bool near_zero(V v) { return std::fabs(v.x) < 1e-8 && std::fabs(v.y) < 1e-8 && std::fabs(v.z) < 1e-8; }
inline double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); } // rtweekend.h:40-46
inline V mk(const double *p) { return {p[0], p[1], p[2]}; }

// core/rtweekend.h:24-50 — the reference's generator with an explicit seed
struct Rng {
    uint32_t s;
    explicit Rng(uint32_t seed) : s(seed ? seed : 1u) {}
    double next() {
        s ^= s << 13;
        s ^= s >> 17;
        s ^= s << 5;
        return s * 2.3283064365386963e-10;
    }
    double range(double a, double b) { return a + (b - a) * next(); }
    int irange(int a, int b) { return static_cast<int>(range(a, b + 1)); }
    V in_unit_sphere() { // vec3.h:226-233 (rejection)
        while (true) {
            const V p{range(-1, 1), range(-1, 1), range(-1, 1)};
            if (len2(p) >= 1)
                continue;
            return p;
        }
    }
    V unit_vector() { return unit(in_unit_sphere()); } // vec3.h:235-237
    V in_unit_disk() {                                 // vec3.h:250-257
        while (true) {
            const V p{range(-1, 1), range(-1, 1), 0};
            if (len2(p) >= 1)
                continue;
            return p;
        }
    }
    V cosine_direction() { // vec3.h:261-269
        const double r1 = next(), r2 = next();
        const double z = std::sqrt(1 - r2), phi = 2 * kPi * r1;
        return {std::cos(phi) * std::sqrt(r2), std::sin(phi) * std::sqrt(r2), z};
    }
};

struct Ray {
    V o, d;
    double tm;
};
struct Rec { // geometry/hittable.h:10-23
    V p, n;
    double t, u, v;
    bool ff;
    int mat, prim;
};
inline void set_face_normal(Rec &r, V dir, V outward) { // hittable.h:19-22
    r.ff = dot(dir, outward) < 0;
    r.n = r.ff ? outward : -outward;
}

struct Camera { // renderer/camera.h:9-30
    V origin, llc, horizontal, vertical, u, v, w;
    double lens_radius, time0, time1;
};

struct Scene {
    std::vector<uint8_t> blob;
    rtb::SceneView *S = nullptr;
    Camera cam;
    std::vector<std::vector<double>> env_tables; // per light (ENV only)
    std::vector<V> quad_normal;
    std::vector<double> quad_area;
    ~Scene() { delete S; }
};

Camera make_camera(const rtb_camera &c) {
    Camera o;
    const double theta = c.vfov * kPi / 180.0, h = std::tan(theta / 2);
    const double vh = 2.0 * h, vw = c.aspect_ratio * vh;
    const V from = mk(c.lookfrom), at = mk(c.lookat), up = mk(c.vup);
    o.w = unit(from - at);
    o.u = unit(cross(up, o.w));
    o.v = cross(o.w, o.u);
    o.origin = from;
    o.horizontal = c.focus_dist * vw * o.u;
    o.vertical = c.focus_dist * vh * o.v;
    o.llc = o.origin - o.horizontal / 2 - o.vertical / 2 - c.focus_dist * o.w;
    o.lens_radius = c.aperture / 2;
    o.time0 = c.time0;
    o.time1 = c.time1;
    return o;
}
Ray camera_ray(const Camera &c, double s, double t, Rng &g) { // camera.h:32-40
    const V rd = c.lens_radius * g.in_unit_disk();
    const V off = c.u * rd.x + c.v * rd.y;
    const V dir = c.llc + s * c.horizontal + t * c.vertical - c.origin - off;
    const double tm = g.range(c.time0, c.time1);
    return {c.origin + off, dir, tm};
}

// ---- geometry -----------------------------------------------------------------------------------
void apply_op(const rtb_xform_op &op, V &o, V &d) {
    if (op.kind == RTB_XF_TRANSLATE) { // hittable.h:53
        o = V{o.x - op.a, o.y - op.b, o.z - op.c};
    } else if (op.kind == RTB_XF_ROTATE_Y) { // hittable.h:129-138 (a = sin, b = cos)
        const V oo = o, dd = d;
        o.x = op.b * oo.x - op.a * oo.z;
        o.z = op.a * oo.x + op.b * oo.z;
        d.x = op.b * dd.x - op.a * dd.z;
        d.z = op.a * dd.x + op.b * dd.z;
    }
}

// leaf hit() in object space: sphere.h:33-60, moving_sphere.h:36-62, aarect.h:79-135
bool hit_leaf(const rtb_prim &p, V o, V d, double tm, double t_min, double t_max, Rec &rec) {
    if (p.type == RTB_PRIM_SPHERE || p.type == RTB_PRIM_MOVING_SPHERE) {
        V c;
        double radius;
        if (p.type == RTB_PRIM_SPHERE) {
            c = mk(p.d);
            radius = p.d[3];
        } else { // moving_sphere.h:32-34
            const V c0 = mk(p.d), c1 = mk(p.d + 3);
            c = c0 + ((tm - p.d[6]) / (p.d[7] - p.d[6])) * (c1 - c0);
            radius = p.d[8];
        }
        const V oc = o - c;
        const double a = len2(d), half_b = dot(oc, d), cc = len2(oc) - radius * radius;
        const double disc = half_b * half_b - a * cc;
        if (disc < 0)
            return false;
        const double sq = std::sqrt(disc);
        double root = (-half_b - sq) / a;
        if (root < t_min || root > t_max) {
            root = (-half_b + sq) / a;
            if (root < t_min || root > t_max)
                return false;
        }
        rec.t = root;
        rec.p = o + root * d;
        const V outward = (rec.p - c) / radius;
        set_face_normal(rec, d, outward);
        rec.u = rec.v = 0;
        if (p.type == RTB_PRIM_SPHERE) { // sphere.h:25-31
            const double theta = std::acos(-outward.y), phi = std::atan2(-outward.z, outward.x) + kPi;
            rec.u = phi / (2 * kPi);
            rec.v = theta / kPi;
        }
        return true;
    }
    const int AX = p.type == RTB_PRIM_XY_RECT ? 2 : (p.type == RTB_PRIM_XZ_RECT ? 1 : 0);
    const int A = p.type == RTB_PRIM_YZ_RECT ? 1 : 0, B = p.type == RTB_PRIM_XY_RECT ? 1 : 2;
    const double t = (p.d[4] - o[AX]) / d[AX];
    if (t < t_min || t > t_max)
        return false;
    const double a = o[A] + t * d[A], b = o[B] + t * d[B];
    if (a < p.d[0] || a > p.d[1] || b < p.d[2] || b > p.d[3])
        return false;
    rec.u = (a - p.d[0]) / (p.d[1] - p.d[0]);
    rec.v = (b - p.d[2]) / (p.d[3] - p.d[2]);
    rec.t = t;
    V outward{0, 0, 0};
    (AX == 0 ? outward.x : (AX == 1 ? outward.y : outward.z)) = 1;
    set_face_normal(rec, d, outward);
    rec.p = o + t * d;
    return true;
}

// aabb::hit (aabb.h:31-48) with the ray's 1/d and signs (ray.h:11-16)
bool box_hit(const double lo[3], const double hi[3], V o, V d, double t_min, double t_max) {
    for (int a = 0; a < 3; ++a) {
        const double inv = 1.0 / d[a];
        double t0 = (lo[a] - o[a]) * inv, t1 = (hi[a] - o[a]) * inv;
        if (inv < 0)
            std::swap(t0, t1);
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max <= t_min)
            return false;
    }
    return true;
}

// A leaf reached through its wrapper chain: translate::hit / rotate_y::hit / flip_face::hit
// (hittable.h:51-62, 127-156, 163-170), outermost wrapper first on the way in, last on the way out.
bool hit_wrapped(const Scene &sc, const rtb_prim &p, const Ray &r, double t_min, double t_max, Rec &rec) {
    V o = r.o, d = r.d;
    V dirs[17];
    int n = 0, first = 0;
    dirs[0] = d;
    if (p.chain >= 0) {
        const rtb_chain c = sc.S->chains()[p.chain];
        first = c.first;
        n = c.count;
        for (int i = 0; i < n; ++i) {
            apply_op(sc.S->xform_ops()[first + i], o, d);
            dirs[i + 1] = d;
        }
    }
    if (p.flags & RTB_PRIM_GATED) { // rtb_gate: the box test of the bvh_node that holds this sphere (bvh.h:42)
        const int64_t self = &p - sc.S->prims();
        for (uint64_t k = 0; k < sc.S->n_gates(); ++k)
            if (sc.S->gates()[k].prim == self && !box_hit(sc.S->gates()[k].lo, sc.S->gates()[k].hi, o, d, t_min, t_max))
                return false;
    }
    if (!hit_leaf(p, o, d, r.tm, t_min, t_max, rec))
        return false;
    for (int i = n - 1; i >= 0; --i) {
        const rtb_xform_op &op = sc.S->xform_ops()[first + i];
        if (op.kind == RTB_XF_TRANSLATE) { // hittable.h:58-59
            rec.p = V{rec.p.x + op.a, rec.p.y + op.b, rec.p.z + op.c};
            set_face_normal(rec, dirs[i + 1], rec.n);
        } else if (op.kind == RTB_XF_ROTATE_Y) { // hittable.h:142-153
            const V p0 = rec.p, n0 = rec.n;
            rec.p.x = op.b * p0.x + op.a * p0.z;
            rec.p.z = -op.a * p0.x + op.b * p0.z;
            const V nn{op.b * n0.x + op.a * n0.z, n0.y, -op.a * n0.x + op.b * n0.z};
            set_face_normal(rec, dirs[i + 1], nn);
        } else { // hittable.h:168
            rec.ff = !rec.ff;
        }
    }
    return true;
}

bool hit_boundary(const Scene &sc, const rtb_prim &m, const Ray &r, double t_min, double t_max, Rec &rec) {
    bool any = false; // hittable_list::hit, hittable_list.h:33-47
    for (int i = m.aux0; i < m.aux0 + m.aux1; ++i) {
        Rec tmp;
        if (hit_wrapped(sc, sc.S->prims()[i], r, t_min, t_max, tmp)) {
            any = true;
            t_max = tmp.t;
            rec = tmp;
        }
    }
    return any;
}

// constant_medium::hit, constant_medium.h:55-104
bool hit_medium(const Scene &sc, const rtb_prim &m, const Ray &r, double t_min, double t_max, Rng &g, Rec &rec) {
    Rec r1, r2;
    if (!hit_boundary(sc, m, r, -kInf, kInf, r1))
        return false;
    if (!hit_boundary(sc, m, r, r1.t + 0.0001, kInf, r2))
        return false;
    if (r1.t < t_min)
        r1.t = t_min;
    if (r2.t > t_max)
        r2.t = t_max;
    if (r1.t >= r2.t)
        return false;
    if (r1.t < 0)
        r1.t = 0;
    const double ray_length = len(r.d);
    const double inside = (r2.t - r1.t) * ray_length;
    const double hit_distance = m.d[0] * std::log(g.next());
    if (hit_distance > inside)
        return false;
    rec.t = r1.t + hit_distance / ray_length;
    rec.p = r.o + rec.t * r.d;
    rec.n = V{1, 0, 0};
    rec.ff = true;
    rec.u = rec.v = 0;
    return true;
}

// scene.hit(): closest over every world primitive (what bvh_node::hit, bvh.h:40-50, computes)
bool scene_hit(const Scene &sc, const Ray &r, double t_min, double t_max, Rng &g, Rec &rec) {
    bool any = false;
    const rtb_prim *P = sc.S->prims();
    const int n = int(sc.S->n_prims());
    for (int i = 0; i < n; ++i) {
        const rtb_prim &p = P[i];
        if (p.flags & RTB_PRIM_BOUNDARY_ONLY)
            continue;
        Rec tmp;
        bool h;
        if (p.type == RTB_PRIM_MEDIUM) {
            h = hit_medium(sc, p, r, t_min, t_max, g, tmp);
            if (p.flags & RTB_PRIM_DUP_LEAF) { // bvh.h:46-47 on a one-object node: tested twice
                Rec t2;
                if (hit_medium(sc, p, r, t_min, h ? tmp.t : t_max, g, t2)) {
                    h = true;
                    tmp = t2;
                }
            }
            // wrappers above a medium re-run set_face_normal on its arbitrary normal; not restated
            // (no reference scene wraps a medium)
        } else {
            h = hit_wrapped(sc, p, r, t_min, t_max, tmp);
        }
        if (h) {
            any = true;
            t_max = tmp.t;
            rec = tmp;
            rec.prim = i;
            rec.mat = p.material;
        }
    }
    return any;
}

// ---- textures (materials/texture.h, perlin.h) -----------------------------------------------------
double perlin_noise(const rtb_perlin &P, V p) { // perlin.h:22-42, 96-111
    const double u = p.x - std::floor(p.x), v = p.y - std::floor(p.y), w = p.z - std::floor(p.z);
    const int i = int(std::floor(p.x)), j = int(std::floor(p.y)), k = int(std::floor(p.z));
    const double uu = u * u * (3 - 2 * u), vv = v * v * (3 - 2 * v), ww = w * w * (3 - 2 * w);
    double accum = 0.0;
    for (int di = 0; di < 2; di++)
        for (int dj = 0; dj < 2; dj++)
            for (int dk = 0; dk < 2; dk++) {
                const int idx = P.perm_x[(i + di) & 255] ^ P.perm_y[(j + dj) & 255] ^ P.perm_z[(k + dk) & 255];
                const V c = mk(P.ranvec[idx]);
                const V wv{u - di, v - dj, w - dk};
                accum += (di * uu + (1 - di) * (1 - uu)) * (dj * vv + (1 - dj) * (1 - vv)) * (dk * ww + (1 - dk) * (1 - ww)) *
                         dot(c, wv);
            }
    return accum;
}
double perlin_turb(const rtb_perlin &P, V p) { // perlin.h:44-56
    double accum = 0.0, weight = 1.0;
    for (int i = 0; i < 7; i++) {
        accum += weight * perlin_noise(P, p);
        weight *= 0.5;
        p = 2.0 * p; // temp_p *= 2
    }
    return std::fabs(accum);
}
V tex_value(const Scene &sc, int id, double u, double v, V p) {
    const rtb_texture &t = sc.S->textures()[id];
    switch (t.type) {
    case RTB_TEX_SOLID: return mk(t.color); // texture.h:47-49
    case RTB_TEX_CHECKER: {                 // texture.h:70-77
        const double s = std::sin(10 * p.x) * std::sin(10 * p.y) * std::sin(10 * p.z);
        return tex_value(sc, s < 0 ? t.odd : t.even, u, v, p);
    }
    case RTB_TEX_IMAGE: { // texture.h:115-139
        const rtb_image &im = sc.S->images()[t.image];
        if (im.width == 0)
            return V{0, 1, 1};
        u = clampd(u, 0.0, 1.0);
        v = 1.0 - clampd(v, 0.0, 1.0);
        int i = int(u * im.width), j = int(v * im.height);
        if (i >= im.width)
            i = im.width - 1;
        if (j >= im.height)
            j = im.height - 1;
        const uint8_t *px = sc.S->image_bytes() + im.offset + (size_t(j) * im.width + i) * 3;
        const double s = 1.0 / 255.0;
        return V{s * px[0], s * px[1], s * px[2]};
    }
    default: { // texture.h:155-158: color(1,1,1) * 0.5 * (1 + sin(scale*z + 10*turb(p)))
        const double n = 0.5 * (1 + std::sin(t.scale * p.z + 10 * perlin_turb(sc.S->perlins()[t.perlin], p)));
        return V{n, n, n};
    }
    }
}

// ---- materials (materials/material.h) --------------------------------------------------------------
struct BS {
    V wi, f;
    double pdf;
    bool spec;
};
double schlick(double cosine, double ref_idx) { // material.h:199-203
    double r0 = (1 - ref_idx) / (1 + ref_idx);
    r0 = r0 * r0;
    return r0 + (1 - r0) * std::pow((1 - cosine), 5);
}
double ggx_D(V N, V H, double rough) { // material.h:397-408
    const double a = rough * rough, a2 = a * a, NdotH = std::max(dot(N, H), 0.0), NdotH2 = NdotH * NdotH;
    double denom = (NdotH2 * (a2 - 1.0) + 1.0);
    denom = kPi * denom * denom;
    return a2 / denom;
}
double ggx_G1(double NdotV, double rough) { // material.h:410-418
    const double k = (rough * rough) / 2.0;
    return NdotV / (NdotV * (1.0 - k) + k);
}
V pbr_normal(const Scene &sc, const rtb_material &m, const Rec &rec) { // material.h:247-262
    V N = rec.n;
    if (m.tex[3] >= 0) {
        V a0 = std::fabs(N.y) > 0.999 ? V{1, 0, 0} : unit(cross(N, V{0, 1, 0}));
        const V a1 = cross(N, a0);
        const V c = tex_value(sc, m.tex[3], rec.u, rec.v, rec.p);
        const V ln = unit(c * 2.0 - V{1, 1, 1}); // texture.h:19-22
        N = unit(ln.x * a0 + ln.y * a1 + ln.z * N);
    }
    return N;
}
double mat_pdf(const Scene &sc, const rtb_material &m, const Rec &rec, V wo, V wi) {
    if (m.type == RTB_MAT_LAMBERTIAN) { // material.h:92-96
        const double c = dot(rec.n, unit(wi));
        return c < 0 ? 0 : c / kPi;
    }
    if (m.type == RTB_MAT_PBR) { // material.h:310-344
        const V N = pbr_normal(sc, m, rec);
        if (dot(N, wi) <= 0)
            return 0;
        double rough = tex_value(sc, m.tex[1], rec.u, rec.v, rec.p).x;
        rough = clampd(rough, 0.01, 1.0);
        const double pdf_diff = dot(N, wi) / kPi;
        const V H = unit(wo + wi);
        const double D = ggx_D(N, H, rough), NdotH = std::max(dot(N, H), 0.0), HdotV = std::max(dot(H, wo), 0.0);
        const double pdf_spec = (D * NdotH) / (4.0 * HdotV + 0.0001);
        return 0.5 * pdf_diff + 0.5 * pdf_spec;
    }
    return 0.0; // material.h:53-56
}
V mat_eval(const Scene &sc, const rtb_material &m, const Rec &rec, V wo, V wi) {
    if (m.type == RTB_MAT_LAMBERTIAN) // material.h:98-101
        return tex_value(sc, m.tex[0], rec.u, rec.v, rec.p) / kPi;
    if (m.type == RTB_MAT_PBR) { // material.h:346-395
        const V N = pbr_normal(sc, m, rec);
        const double NdotL = dot(N, wi), NdotV = dot(N, wo);
        if (NdotL <= 0 || NdotV <= 0)
            return V{0, 0, 0};
        double rough = tex_value(sc, m.tex[1], rec.u, rec.v, rec.p).x;
        const double metal = tex_value(sc, m.tex[2], rec.u, rec.v, rec.p).x;
        const V base = tex_value(sc, m.tex[0], rec.u, rec.v, rec.p);
        rough = clampd(rough, 0.01, 1.0);
        const V H = unit(wo + wi);
        const V one{1, 1, 1}, mv{metal, metal, metal};
        const V F0 = (one - mv) * V{0.04, 0.04, 0.04} + mv * base;
        const V F = F0 + (one - F0) * std::pow(1.0 - std::max(dot(H, wo), 0.0), 5.0);
        const double D = ggx_D(N, H, rough);
        const double G = ggx_G1(std::max(dot(N, wi), 0.0), rough) * ggx_G1(std::max(dot(N, wo), 0.0), rough);
        const V numerator = D * G * F;
        const double denominator = 4.0 * NdotV * NdotL + 0.0001;
        const V specular = numerator / denominator;
        V kD = one - F;
        kD = kD * (1.0 - metal);
        const V diffuse = kD * base / kPi;
        return diffuse + specular;
    }
    return V{0, 0, 0}; // material.h:47-50
}
V emitted_old(const Scene &sc, const rtb_material &m, const Rec &rec) { // material.h:27-29, 220-222
    return m.type == RTB_MAT_DIFFUSE_LIGHT ? tex_value(sc, m.tex[0], rec.u, rec.v, rec.p) : V{0, 0, 0};
}
V emitted_new(const Scene &sc, const rtb_material &m, const Rec &rec) { // material.h:32-34, 224-229
    return (m.type == RTB_MAT_DIFFUSE_LIGHT && rec.ff) ? tex_value(sc, m.tex[0], rec.u, rec.v, rec.p) : V{0, 0, 0};
}
bool mat_sample(const Scene &sc, const rtb_material &m, const Rec &rec, V wo, Rng &g, BS &bs) {
    switch (m.type) {
    case RTB_MAT_LAMBERTIAN: { // material.h:79-90
        V dir = rec.n + g.unit_vector();
        if (near_zero(dir))
            dir = rec.n;
        bs.wi = unit(dir);
        bs.pdf = dot(rec.n, bs.wi) / kPi;
        bs.f = tex_value(sc, m.tex[0], rec.u, rec.v, rec.p) / kPi;
        bs.spec = false;
        return true;
    }
    case RTB_MAT_METAL: { // material.h:123-131
        const V reflected = reflect(unit(-wo), rec.n);
        bs.wi = unit(reflected + m.fuzz * g.in_unit_sphere());
        bs.f = mk(m.color);
        bs.pdf = 1.0;
        bs.spec = true;
        return dot(bs.wi, rec.n) > 0;
    }
    case RTB_MAT_DIELECTRIC: { // material.h:152-174
        bs.f = V{1, 1, 1};
        bs.spec = true;
        bs.pdf = 1.0;
        const double ratio = rec.ff ? (1.0 / m.ir) : m.ir;
        const V ud = -wo;
        const double cos_theta = std::fmin(dot(-ud, rec.n), 1.0), sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
        const bool cannot = ratio * sin_theta > 1.0;
        if (cannot || schlick(cos_theta, ratio) > g.next())
            bs.wi = reflect(ud, rec.n);
        else
            bs.wi = refract(ud, rec.n, ratio);
        return true;
    }
    case RTB_MAT_PBR: { // material.h:245-308
        const V N = pbr_normal(sc, m, rec);
        double rough = tex_value(sc, m.tex[1], rec.u, rec.v, rec.p).x;
        rough = clampd(rough, 0.01, 1.0);
        V w = unit(N); // onb::build_from_w, onb.h:29-34
        const V a = (std::fabs(w.x) > 0.9) ? V{0, 1, 0} : V{1, 0, 0};
        const V vv = unit(cross(w, a)), uu = cross(w, vv);
        if (g.next() < 0.5) {
            const double r1 = g.next(), r2 = g.next(), al = rough * rough, phi = 2.0 * kPi * r1;
            const double ct = std::sqrt((1.0 - r2) / (1.0 + (al * al - 1.0) * r2)), st = std::sqrt(1.0 - ct * ct);
            const V Hl{st * std::cos(phi), st * std::sin(phi), ct};
            const V H = Hl.x * uu + Hl.y * vv + Hl.z * w;
            const V L = reflect(-wo, H);
            if (dot(N, L) <= 0)
                return false;
            bs.wi = L;
        } else {
            const V c = g.cosine_direction();
            V L = c.x * uu + c.y * vv + c.z * w;
            if (dot(N, L) <= 0)
                L = N;
            bs.wi = unit(L);
        }
        bs.spec = false;
        bs.pdf = mat_pdf(sc, m, rec, wo, bs.wi);
        bs.f = mat_eval(sc, m, rec, wo, bs.wi);
        return !(bs.pdf < 1e-6);
    }
    default: return false; // diffuse_light material.h:213-216; isotropic inherits the base (material.h:41-44)
    }
}
bool mat_scatter(const Scene &sc, const rtb_material &m, const Rec &rec, const Ray &rin, Rng &g, V &atten, Ray &out) {
    switch (m.type) {
    case RTB_MAT_LAMBERTIAN: { // material.h:103-112
        V dir = rec.n + g.unit_vector();
        if (near_zero(dir))
            dir = rec.n;
        out = {rec.p, dir, rin.tm};
        atten = tex_value(sc, m.tex[0], rec.u, rec.v, rec.p);
        return true;
    }
    case RTB_MAT_METAL: { // material.h:133-140
        const V reflected = reflect(unit(rin.d), rec.n);
        out = {rec.p, reflected + m.fuzz * g.in_unit_sphere(), rin.tm};
        atten = mk(m.color);
        return dot(out.d, rec.n) > 0;
    }
    case RTB_MAT_DIELECTRIC: { // material.h:176-193
        atten = V{1, 1, 1};
        const double ratio = rec.ff ? (1.0 / m.ir) : m.ir;
        const V ud = unit(rin.d);
        const double cos_theta = std::fmin(dot(-ud, rec.n), 1.0), sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
        V dir;
        if (ratio * sin_theta > 1.0 || schlick(cos_theta, ratio) > g.next())
            dir = reflect(ud, rec.n);
        else
            dir = refract(ud, rec.n, ratio);
        out = {rec.p, dir, rin.tm};
        return true;
    }
    case RTB_MAT_ISOTROPIC: // constant_medium.h:19-24
        out = {rec.p, g.in_unit_sphere(), rin.tm};
        atten = tex_value(sc, m.tex[0], rec.u, rec.v, rec.p);
        return true;
    default: return false; // material.h:231-234; PBRMaterial has no scatter()
    }
}

// ---- lights (lighting/*.h) -------------------------------------------------------------------------
struct LS {
    V Li, wi;
    double pdf, dist;
    bool delta;
};
struct Env { // views into Scene::env_tables[l] — environmental_light.h:15-27, 62-80
    const double *func, *cdf, *rint, *mcdf;
    double mint;
    int W, H;
};
Env env_of(const Scene &sc, int l) {
    const rtb_light &L = sc.S->lights()[l];
    const int W = L.env_width, H = L.env_height;
    const double *b = sc.env_tables[l].data();
    Env e;
    e.W = W;
    e.H = H;
    e.func = b;
    e.cdf = b + size_t(W) * H;
    e.rint = e.cdf + size_t(W + 1) * H;
    e.mcdf = e.rint + H;
    e.mint = e.mcdf[H + 1];
    return e;
}
double dist1d_sample(const double *func, const double *cdf, double fint, int n, double u, double &pdf, int &off) {
    // environmental_light.h:30-44
    const double *it = std::lower_bound(cdf, cdf + n + 1, u);
    off = std::max(0, int(it - cdf) - 1);
    off = std::min(off, n - 1);
    double du = u - cdf[off];
    if (cdf[off + 1] - cdf[off] > 0)
        du /= (cdf[off + 1] - cdf[off]);
    pdf = (fint > 0) ? func[off] / fint : 0;
    return (off + du) / n;
}
V env_pixel(const float *tex, int W, int H, int i, int j) { // environmental_light.h:297-311
    if (i < 0)
        i += W;
    if (i >= W)
        i -= W;
    if (j < 0)
        j = 0;
    if (j >= H)
        j = H - 1;
    const float *q = tex + 3 * (size_t(j) * W + i);
    return V{q[0], q[1], q[2]};
}
void env_uv(const rtb_light &L, V ud, double &u, double &v, double &theta) {
    if (L.env_is_probe) { // environmental_light.h:258-266, 325-334
        const double d = std::sqrt(ud.x * ud.x + ud.y * ud.y);
        const double rc = (d > 0) ? (1.0 / kPi) * std::acos(ud.z) / d : 0.0;
        u = (ud.x * rc + 1.0) * 0.5;
        v = (ud.y * rc + 1.0) * 0.5;
        v = 1.0 - v;
        theta = std::acos(ud.z);
    } else { // environmental_light.h:268-274
        theta = std::acos(ud.y);
        const double phi = std::atan2(-ud.z, ud.x) + kPi;
        u = phi / (2 * kPi);
        v = theta / kPi;
    }
}
V env_Le(const Scene &sc, int l, V dir) { // environmental_light.h:250-295
    const rtb_light &L = sc.S->lights()[l];
    if (L.env_width == 0)
        return V{1, 1, 1};
    double u, v, theta;
    env_uv(L, unit(dir), u, v, theta);
    const int W = L.env_width, H = L.env_height;
    const double ui = u * W - 0.5, vi = v * H - 0.5;
    const int i0 = int(std::floor(ui)), j0 = int(std::floor(vi));
    const double du = ui - i0, dv = vi - j0;
    const float *tex = sc.S->env_texels() + L.env_offset;
    const V c00 = env_pixel(tex, W, H, i0, j0), c10 = env_pixel(tex, W, H, i0 + 1, j0);
    const V c01 = env_pixel(tex, W, H, i0, j0 + 1), c11 = env_pixel(tex, W, H, i0 + 1, j0 + 1);
    const V c0 = c00 * (1 - du) + c10 * du, c1 = c01 * (1 - du) + c11 * du;
    return c0 * (1 - dv) + c1 * dv;
}
V light_Le(const Scene &sc, int l, V dir) { return sc.S->lights()[l].type == RTB_LIGHT_ENV ? env_Le(sc, l, dir) : V{0, 0, 0}; }
double light_pdf(const Scene &sc, int l, V origin, V direction) {
    const rtb_light &L = sc.S->lights()[l];
    if (L.type == RTB_LIGHT_QUAD) { // quad_light.h:51-82
        const V normal = sc.quad_normal[l], Q = mk(L.Q), uu = mk(L.u), vv = mk(L.v);
        const double denom = dot(direction, normal);
        if (denom >= -1e-6)
            return 0;
        const double t = dot(Q - origin, normal) / denom;
        if (t < 0.001 || t > kInf)
            return 0;
        const V ph = origin + t * direction - Q;
        const double alpha = dot(ph, uu) / len2(uu), beta = dot(ph, vv) / len2(vv);
        if (alpha < 0 || alpha > 1 || beta < 0 || beta > 1)
            return 0;
        const double dist_sq = t * t * len2(direction), cos_theta = -denom / len(direction);
        return dist_sq / (sc.quad_area[l] * cos_theta);
    }
    if (L.type == RTB_LIGHT_ENV) { // environmental_light.h:314-356
        if (L.env_width == 0)
            return 1.0 / (4.0 * kPi);
        double u, v, theta;
        env_uv(L, unit(direction), u, v, theta);
        const double sin_theta = std::sin(theta);
        if (sin_theta < 1e-6)
            return 0;
        const Env e = env_of(sc, l);
        const int ui = int(clampd(int(u * e.W), 0, e.W - 1)), vi = int(clampd(int(v * e.H), 0, e.H - 1));
        const double ci = e.rint[vi];
        const double pc = ci > 0 ? e.func[size_t(vi) * e.W + ui] / (ci * e.W) : 0; // Distribution1D::pdf, :46-48
        const double pm = e.mint > 0 ? ci / (e.mint * e.H) : 0;
        return pc * pm * e.W * e.H / (2.0 * kPi * kPi * sin_theta);
    }
    return 0.0; // light.h:24-26
}
LS light_sample(const Scene &sc, int l, V p, double u0, double u1, Rng &g) {
    const rtb_light &L = sc.S->lights()[l];
    LS s{{0, 0, 0}, {0, 0, 0}, 0, 0, false};
    const V Q = mk(L.Q), I = mk(L.intensity);
    switch (L.type) {
    case RTB_LIGHT_QUAD: { // quad_light.h:18-49
        const V lp = Q + u0 * mk(L.u) + u1 * mk(L.v), d = lp - p;
        const double d2 = len2(d);
        s.dist = std::sqrt(d2);
        s.wi = d / s.dist;
        const double ct = dot(-s.wi, sc.quad_normal[l]);
        if (ct <= 0)
            return s;
        s.Li = I;
        s.pdf = d2 / (sc.quad_area[l] * ct);
        return s;
    }
    case RTB_LIGHT_POINT: { // point_light.h:13-26
        const V d = Q - p;
        const double d2 = len2(d);
        s.dist = std::sqrt(d2);
        s.wi = d / s.dist;
        s.Li = I / d2;
        s.pdf = 1.0;
        s.delta = true;
        return s;
    }
    case RTB_LIGHT_SPOT: { // spot_light.h:15-34
        const V d = Q - p;
        const double d2 = len2(d);
        s.dist = std::sqrt(d2);
        s.wi = d / s.dist;
        s.delta = true;
        s.pdf = 1.0;
        if (!(dot(-s.wi, mk(L.u)) < L.cos_cutoff))
            s.Li = I / d2;
        return s;
    }
    case RTB_LIGHT_DIRECTIONAL: // directional_light.h:14-22
        s.wi = -mk(L.u);
        s.dist = kInf;
        s.Li = I;
        s.delta = true;
        s.pdf = 1.0;
        return s;
    default: { // environmental_light.h:182-248
        s.dist = kInf;
        if (L.env_width == 0) {
            s.wi = g.unit_vector();
            s.pdf = 1.0 / (4.0 * kPi);
            s.Li = V{1, 1, 1};
            return s;
        }
        const Env e = env_of(sc, l);
        double pdfs[2];
        int vi, ui;
        const double v = dist1d_sample(e.rint, e.mcdf, e.mint, e.H, u1, pdfs[1], vi);
        const double u = dist1d_sample(e.func + size_t(vi) * e.W, e.cdf + size_t(vi) * (e.W + 1), e.rint[vi], e.W, u0, pdfs[0], ui);
        const double map_pdf = pdfs[0] * pdfs[1];
        if (map_pdf == 0)
            return s;
        double theta;
        if (L.env_is_probe) {
            const double uc = u * 2.0 - 1.0, vc = (1.0 - v) * 2.0 - 1.0, r = std::sqrt(uc * uc + vc * vc);
            if (r > 1.0)
                return s;
            theta = kPi * r;
            const double phi = std::atan2(vc, uc), st = std::sin(theta);
            s.wi = V{st * std::cos(phi), st * std::sin(phi), std::cos(theta)};
        } else {
            const double phi = u * 2 * kPi - kPi;
            theta = v * kPi;
            const double st = std::sin(theta), ct = std::cos(theta);
            s.wi = V{st * std::cos(phi), ct, -st * std::sin(phi)};
        }
        const double sin_theta = std::sin(theta);
        if (sin_theta < 1e-6)
            return s;
        s.pdf = map_pdf * e.W * e.H / (2.0 * kPi * kPi * sin_theta);
        s.Li = env_Le(sc, l, s.wi);
        return s;
    }
    }
}

// ---- integrators (renderer/*_integrator.h) -----------------------------------------------------------
struct Counters {
    uint64_t closest = 0, shadow = 0;
};
double power_heuristic(double a, double b) { // mis_path_integrator.h:165-170
    const double a2 = a * a, b2 = b * b, d = a2 + b2;
    return d > 0 ? a2 / d : 0.0;
}
V clamp_radiance(V L, double mx = 100.0) { // mis_path_integrator.h:154-162
    if (L.x > mx || L.y > mx || L.z > mx) {
        const double m = std::max({L.x, L.y, L.z});
        if (m > mx)
            return L * (mx / m);
    }
    return L;
}
double all_lights_pdf(const Scene &sc, const Ray &r) { // mis_path_integrator.h:173-188
    const int n = int(sc.S->n_lights());
    double total = 0.0;
    const double sel = 1.0 / n;
    for (int l = 0; l < n; ++l)
        total += light_pdf(sc, l, r.o, r.d) * sel;
    return total;
}

V Li(const Scene &sc, int integrator, Ray ray, int max_depth, int rr_start, Rng &g, Counters &cnt) {
    const V bg = mk(sc.S->globals().background);
    const int nl = int(sc.S->n_lights());
    V T{1, 1, 1}, L{0, 0, 0};
    bool spec = false;
    double prev_pdf = 0.0;
    for (int depth = 0; depth < max_depth; ++depth) {
        Rec rec;
        cnt.closest++;
        if (!scene_hit(sc, ray, 0.001, kInf, g, rec)) {
            if (integrator <= 2) { // path_integrator.h:28-30, rr:30-33, pbr:30-33
                L = L + T * bg;
                break;
            }
            V env{0, 0, 0};
            bool found = false;
            for (int l = 0; l < nl; ++l)
                if (sc.S->lights()[l].type == RTB_LIGHT_ENV) {
                    env = env + light_Le(sc, l, ray.d);
                    found = true;
                }
            if (!found)
                L = L + T * bg;
            else if (integrator == 3 || depth == 0 || spec) // direct:35-48, mis:51-52
                L = L + T * env;
            else // mis:53-63
                L = L + T * env * power_heuristic(prev_pdf, all_lights_pdf(sc, ray));
            break;
        }
        const rtb_material &m = sc.S->materials()[rec.mat];
        if (integrator <= 1) { // path_integrator.h:32-44, rr_path_integrator.h:35-57
            L = L + T * emitted_old(sc, m, rec);
            V atten;
            Ray scattered;
            if (!mat_scatter(sc, m, rec, ray, g, atten, scattered))
                break;
            T = T * atten;
            if (integrator == 1 && depth >= rr_start) {
                const double ps = clampd(std::max({T.x, T.y, T.z}), 0.005, 0.95);
                if (g.next() > ps)
                    break;
                T = T / ps;
            }
            ray = scattered;
            continue;
        }
        const V wo = -unit(ray.d);
        const V e = emitted_new(sc, m, rec);
        if (integrator == 2) { // pbr_path_integrator.h:38-39
            L = L + T * e;
        } else if (integrator == 3) { // direct_light_integrator.h:52-55
            if (depth == 0 || spec)
                L = L + T * e;
        } else if (len2(e) > 0) { // mis_path_integrator.h:72-94
            V Le;
            if (depth == 0 || spec)
                Le = T * e;
            else if (nl > 0)
                Le = T * e * power_heuristic(prev_pdf, all_lights_pdf(sc, ray));
            else
                Le = T * e;
            L = L + (depth == 0 ? Le : clamp_radiance(Le));
        }
        spec = false; // material::is_specular(), material.h:37 (never overridden)
        if (integrator >= 3 && nl > 0) {
            // sample_lights_direct (direct:98-142) / sample_lights_mis (mis:192-234)
            V Ld{0, 0, 0};
            const int li = g.irange(0, nl - 1);
            const double sel = 1.0 / nl;
            const double u0 = g.next(), u1 = g.next();
            const LS ls = light_sample(sc, li, rec.p, u0, u1, g);
            if (ls.pdf > 0 && len2(ls.Li) > 0) {
                const Ray sr{rec.p, ls.wi, 0};
                Rec srec;
                cnt.shadow++;
                if (!scene_hit(sc, sr, 0.001, ls.dist - 0.001, g, srec)) {
                    const V f = mat_eval(sc, m, rec, wo, ls.wi);
                    const double ct = std::fabs(dot(ls.wi, rec.n));
                    if (integrator == 3) {
                        Ld = ls.delta ? f * ls.Li * ct / sel : f * ls.Li * ct / (ls.pdf * sel);
                    } else if (ls.delta) {
                        Ld = f * ls.Li * ct / sel;
                    } else {
                        const double bp = mat_pdf(sc, m, rec, wo, ls.wi), lp = ls.pdf * sel;
                        Ld = f * ls.Li * ct * power_heuristic(lp, bp) / lp;
                    }
                }
            }
            if (integrator == 3) { // direct:133-139
                if (Ld.x > 100.0)
                    Ld = Ld * (100.0 / Ld.x);
                if (Ld.y > 100.0)
                    Ld = Ld * (100.0 / Ld.y);
                if (Ld.z > 100.0)
                    Ld = Ld * (100.0 / Ld.z);
                L = L + T * Ld;
            } else {
                L = L + clamp_radiance(T * Ld);
            }
        }
        BS bs;
        if (!mat_sample(sc, m, rec, wo, g, bs)) {
            if (integrator != 4) // pbr:43-45, direct:65-67
                break;
            V atten; // mis:106-117
            Ray scattered;
            if (!mat_scatter(sc, m, rec, ray, g, atten, scattered))
                break;
            T = T * atten;
            ray = scattered;
            spec = false;
            prev_pdf = 0.0;
        } else {
            if (bs.pdf < 1e-8 && !bs.spec)
                break;
            spec = bs.spec;
            prev_pdf = bs.spec ? 0.0 : bs.pdf;
            const double ct = std::fabs(dot(bs.wi, rec.n));
            T = bs.spec ? T * bs.f : T * (bs.f * ct / bs.pdf);
            ray = Ray{rec.p, bs.wi, ray.tm};
        }
        if (depth >= rr_start) { // pbr:59-69, direct:84-92, mis:137-146
            const double ps = clampd(std::max({T.x, T.y, T.z}), 0.05, 0.95);
            if (g.next() > ps)
                break;
            T = T / ps;
        }
    }
    return L;
}

void fill_hit(bool ok, const Rec &rec, rtb_hit &h) {
    std::memset(&h, 0, sizeof(h));
    h.prim = -1;
    h.material = -1;
    if (!ok)
        return;
    h.t = rec.t;
    h.p[0] = rec.p.x; h.p[1] = rec.p.y; h.p[2] = rec.p.z;
    h.normal[0] = rec.n.x; h.normal[1] = rec.n.y; h.normal[2] = rec.n.z;
    h.u = rec.u;
    h.v = rec.v;
    h.prim = rec.prim;
    h.front_face = rec.ff;
    h.material = rec.mat;
}
Rec rec_of(const rtb_bsdf_query &q) {
    Rec r;
    r.p = mk(q.p);
    r.n = mk(q.normal);
    r.u = q.u;
    r.v = q.v;
    r.t = 1;
    r.ff = q.front_face != 0;
    r.mat = r.prim = 0;
    return r;
}

} // namespace

extern "C" {

void *port_scene_create(const void *blob, uint64_t nbytes) {
    try {
        auto *sc = new Scene();
        sc->blob.assign(static_cast<const uint8_t *>(blob), static_cast<const uint8_t *>(blob) + nbytes);
        sc->S = new rtb::SceneView(sc->blob.data(), sc->blob.size());
        sc->S->validate();
        sc->cam = make_camera(sc->S->camera());
        const int nl = int(sc->S->n_lights());
        sc->env_tables.resize(nl);
        sc->quad_normal.resize(nl, V{0, 0, 0});
        sc->quad_area.resize(nl, 0.0);
        for (int l = 0; l < nl; ++l) {
            const rtb_light &L = sc->S->lights()[l];
            if (L.type == RTB_LIGHT_QUAD) { // quad_light.h:9-16
                const V n = cross(mk(L.u), mk(L.v));
                sc->quad_area[l] = len(n);
                sc->quad_normal[l] = unit(n);
            }
            if (L.type == RTB_LIGHT_ENV && L.env_width > 0) { // environmental_light.h:146-180, 15-27
                const int W = L.env_width, H = L.env_height;
                std::vector<double> &t = sc->env_tables[l];
                t.assign(size_t(W) * H + size_t(W + 1) * H + H + (H + 1) + 1, 0.0);
                double *func = t.data(), *cdf = func + size_t(W) * H, *rint = cdf + size_t(W + 1) * H, *mcdf = rint + H;
                const float *tex = sc->S->env_texels() + L.env_offset;
                for (int v = 0; v < H; ++v) {
                    const double st = std::sin(kPi * (v + 0.5) / H);
                    for (int u = 0; u < W; ++u) {
                        const size_t i = size_t(v) * W + u;
                        const double r = tex[3 * i], gg = tex[3 * i + 1], b = tex[3 * i + 2];
                        func[i] = (0.2126 * r + 0.7152 * gg + 0.0722 * b) * st;
                    }
                    double *c = cdf + size_t(v) * (W + 1);
                    c[0] = 0;
                    for (int i = 1; i <= W; ++i)
                        c[i] = c[i - 1] + func[size_t(v) * W + i - 1];
                    rint[v] = c[W];
                    if (rint[v] > 0)
                        for (int i = 0; i <= W; ++i)
                            c[i] /= rint[v];
                }
                mcdf[0] = 0;
                for (int i = 1; i <= H; ++i)
                    mcdf[i] = mcdf[i - 1] + rint[i - 1];
                mcdf[H + 1] = mcdf[H];
                if (mcdf[H + 1] > 0)
                    for (int i = 0; i <= H; ++i)
                        mcdf[i] /= mcdf[H + 1];
            }
        }
        return sc;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "port_scene_create: %s\n", e.what());
        return nullptr;
    }
}
void port_scene_destroy(void *h) { delete static_cast<Scene *>(h); }

void port_camera_derived(void *h, double out[24]) {
    const Camera &c = static_cast<Scene *>(h)->cam;
    const V *vs[7] = {&c.origin, &c.llc, &c.horizontal, &c.vertical, &c.u, &c.v, &c.w};
    for (int i = 0; i < 7; ++i) {
        out[3 * i] = vs[i]->x;
        out[3 * i + 1] = vs[i]->y;
        out[3 * i + 2] = vs[i]->z;
    }
    out[21] = c.lens_radius;
    out[22] = c.time0;
    out[23] = c.time1;
}

void port_trace_batch(void *h, const rtb_ray *rays, uint64_t n, uint32_t seed, rtb_hit *hits) {
    const Scene &sc = *static_cast<Scene *>(h);
    Rng g(seed);
    for (uint64_t i = 0; i < n; ++i) {
        const Ray r{mk(rays[i].o), mk(rays[i].d), rays[i].time};
        Rec rec;
        if (rays[i].reserved) // the state the reference's generator had when it answered this query (ref_harness recorder)
            g = Rng(uint32_t(rays[i].reserved));
        const bool ok = scene_hit(sc, r, rays[i].t_min, rays[i].t_max, g, rec);
        fill_hit(ok, rec, hits[i]);
    }
}

void port_bsdf_eval(void *h, int material, const rtb_bsdf_query *q, uint64_t n, rtb_bsdf_value *out) {
    const Scene &sc = *static_cast<Scene *>(h);
    const rtb_material &m = sc.S->materials()[material];
    for (uint64_t i = 0; i < n; ++i) {
        const Rec rec = rec_of(q[i]);
        const V wo = mk(q[i].wo), wi = mk(q[i].wi);
        const V f = mat_eval(sc, m, rec, wo, wi), e0 = emitted_old(sc, m, rec), e1 = emitted_new(sc, m, rec);
        out[i].f[0] = f.x; out[i].f[1] = f.y; out[i].f[2] = f.z;
        out[i].pdf = mat_pdf(sc, m, rec, wo, wi);
        out[i].emitted_old[0] = e0.x; out[i].emitted_old[1] = e0.y; out[i].emitted_old[2] = e0.z;
        out[i].emitted_new[0] = e1.x; out[i].emitted_new[1] = e1.y; out[i].emitted_new[2] = e1.z;
    }
}

void port_light_eval(void *h, int light, const rtb_light_query *q, uint64_t n, uint32_t seed, rtb_light_value *out) {
    const Scene &sc = *static_cast<Scene *>(h);
    Rng g(seed);
    for (uint64_t i = 0; i < n; ++i) {
        std::memset(&out[i], 0, sizeof(out[i]));
        const V p = mk(q[i].p), d = mk(q[i].d);
        const LS s = light_sample(sc, light, p, q[i].u[0], q[i].u[1], g);
        out[i].Li[0] = s.Li.x; out[i].Li[1] = s.Li.y; out[i].Li[2] = s.Li.z;
        out[i].wi[0] = s.wi.x; out[i].wi[1] = s.wi.y; out[i].wi[2] = s.wi.z;
        out[i].pdf = s.pdf;
        out[i].dist = s.dist;
        out[i].is_delta = s.delta;
        out[i].pdf_dir = light_pdf(sc, light, p, d);
        const V le = light_Le(sc, light, d);
        out[i].Le[0] = le.x; out[i].Le[1] = le.y; out[i].Le[2] = le.z;
    }
}

void port_texture_value(void *h, int texture, const double *uvp, uint64_t n, double *rgb) {
    const Scene &sc = *static_cast<Scene *>(h);
    for (uint64_t i = 0; i < n; ++i) {
        const double *q = uvp + 5 * i;
        const V c = tex_value(sc, texture, q[0], q[1], V{q[2], q[3], q[4]});
        rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z;
    }
}

// renderer.h:72-79 with the gamma/clamp of :126-140 bypassed: per-pixel sum and sum of squares
// of linear Li.  Returns wall seconds; counters = {closest-hit queries, shadow queries}.
double port_render_linear(void *h, int integrator, int width, int height, int spp, int max_depth, int n_threads,
                          uint32_t seed, double *sum, double *sumsq, uint64_t counters[2]) {
    const Scene &sc = *static_cast<Scene *>(h);
    if (n_threads <= 0)
        n_threads = int(std::thread::hardware_concurrency());
    std::atomic<int> next_row(0);
    std::atomic<uint64_t> nc(0), ns(0);
    const auto t0 = std::chrono::high_resolution_clock::now();
    auto worker = [&](int tid) {
        Counters cnt;
        while (true) {
            const int j = next_row.fetch_add(1);
            if (j >= height)
                break;
            Rng g(seed * 2654435761u + uint32_t(j) * 40503u + 977u + uint32_t(tid));
            for (int i = 0; i < width; ++i) {
                double s[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
                for (int k = 0; k < spp; ++k) {
                    const double u = (i + g.next()) / (width - 1), v = (j + g.next()) / (height - 1);
                    const Ray r = camera_ray(sc.cam, u, v, g);
                    const V c = Li(sc, integrator, r, max_depth, 3, g, cnt);
                    const double cc[3] = {c.x, c.y, c.z};
                    for (int ch = 0; ch < 3; ++ch) {
                        s[ch] += cc[ch];
                        s2[ch] += cc[ch] * cc[ch];
                    }
                }
                for (int ch = 0; ch < 3; ++ch) {
                    if (sum)
                        sum[(size_t(j) * width + i) * 3 + ch] = s[ch];
                    if (sumsq)
                        sumsq[(size_t(j) * width + i) * 3 + ch] = s2[ch];
                }
            }
        }
        nc += cnt.closest;
        ns += cnt.shadow;
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back(worker, t);
    for (auto &t : th)
        t.join();
    if (counters) {
        counters[0] = nc;
        counters[1] = ns;
    }
    return std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
}

int port_hardware_threads() { return int(std::thread::hardware_concurrency()); }
}
