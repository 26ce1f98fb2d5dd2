// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into, loaded by
// or shipped with the product library).
//
// A headless C-ABI around the UNMODIFIED reference renderer.  The reference
// sources are compiled from where they lie under /root/reference/src (see
// oracle/Makefile); nothing is copied.  The harness replaces only what
// src/main.cpp:61-151 does (scene selection, camera/integrator wiring, the SDL
// window) and adds the probes the parity tests need:
//
//   * ref_scene_create        select_scene(id) (scenes.cpp:1523) + a walk of the
//                             graph that emits the flat scene blob of
//                             include/rtb200_scene.h and tags every leaf with
//                             its flat primitive id (the reference's hit_record
//                             has no primitive id, hittable.h:10-23).
//   * ref_trace_batch         hittable::hit() of the reference on caller rays.
//   * ref_record_rays         runs Integrator::Li and logs every hit() query it
//                             issues (camera, bounce and shadow rays).
//   * ref_bsdf_* / ref_light_* / ref_texture_value
//                             material / Light / texture virtuals on caller grids.
//   * ref_render_linear       per-pixel sum and sum-of-squares of UNCLAMPED linear
//                             Li (bypasses renderer.h:126-140 gamma/clamp).
//   * ref_render_timed        the reference's own Renderer::render
//                             (renderer.h:30-102), untouched, timed.
//
// The reference keeps light / texture / camera members private; the harness
// reads them with the `#define private public` trick below.  That does not
// change any reference code path that is executed.
//
// Non-inline definitions live in the reference's headers (bvh.h, sphere.h, ...)
// so they may be included by one TU only; scenes.cpp is that TU and is
// #included here.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#define private public
#define protected public
#include "scenes.cpp" // the reference's one "geometry definitions" TU
#include "direct_light_integrator.h"
#include "mis_path_integrator.h"
#include "path_integrator.h"
#include "pbr_path_integrator.h"
#include "renderer.h"
#include "rr_path_integrator.h"
#include "camera.h"
#undef private
#undef protected

#include "rtb200_blob.hpp"
#include "rtb200_types.h"

namespace {

thread_local int g_last_prim = -1;

// Forwards to the wrapped leaf and remembers which flat primitive answered.
// In the reference's traversal every successful leaf hit shrinks t_max
// (bvh.h:46-47, hittable_list.h:39-45), so the LAST success is the closest.
struct id_tagger : public hittable {
    id_tagger(shared_ptr<hittable> in, int pid) : inner(std::move(in)), id(pid) {}
    bool hit(const ray &r, double t_min, double t_max, hit_record &rec) const override {
        if (!inner->hit(r, t_min, t_max, rec))
            return false;
        g_last_prim = id;
        return true;
    }
    bool bounding_box(double t0, double t1, aabb &out) const override {
        return inner->bounding_box(t0, t1, out);
    }
    shared_ptr<hittable> inner;
    int id;
};

struct RayLog {
    std::vector<rtb_ray> rays;
    std::vector<rtb_hit> hits;
    size_t limit = 0;
    // recent path vertices, to attribute an origin primitive to each ray
    double vp[2][3] = {{NAN, NAN, NAN}, {NAN, NAN, NAN}};
    int vprim[2] = {-1, -1};
    uint64_t n_closest = 0, n_shadow = 0;
    void new_path() {
        vprim[0] = vprim[1] = -1;
        vp[0][0] = vp[1][0] = NAN;
    }
};
thread_local RayLog *g_log = nullptr;

struct SceneHandle {
    SceneConfig config;
    shared_ptr<camera> cam;
    int width = 0, height = 0;
    std::vector<uint8_t> blob;
    std::vector<const material *> materials;
    std::vector<const texture *> textures;
    std::map<const material *, int> mat_ids;
    std::string error;
};

void fill_hit(const SceneHandle *h, bool ok, const hit_record &rec, int prim, rtb_hit &out) {
    std::memset(&out, 0, sizeof(out));
    out.prim = -1;
    out.material = -1;
    if (!ok)
        return;
    out.t = rec.t;
    for (int k = 0; k < 3; ++k) {
        out.p[k] = rec.p[k];
        out.normal[k] = rec.normal[k];
    }
    out.u = rec.u;
    out.v = rec.v;
    out.prim = prim;
    out.front_face = rec.front_face ? 1 : 0;
    auto it = h->mat_ids.find(rec.mat_ptr);
    out.material = it == h->mat_ids.end() ? -1 : it->second;
}

// Sits on top of the world while Integrator::Li runs and logs every query.
struct recorder : public hittable {
    recorder(shared_ptr<hittable> in, const SceneHandle *hh) : inner(std::move(in)), h(hh) {}
    bool hit(const ray &r, double t_min, double t_max, hit_record &rec) const override {
        g_last_prim = -1;
        // The state of the reference's xorshift32 (a function-local thread_local, rtweekend.h:24-34) is
        // read off one extra draw: random_double() returns state * 2^-32 exactly.  The draws
        // constant_medium::hit makes while answering THIS query are its successors, so a consumer that
        // walks the leaves in the reference's order can reproduce the query bit for bit, media included.
        uint32_t rng_state = 0;
        if (g_log)
            rng_state = uint32_t(random_double() * 4294967296.0);
        const bool ok = inner->hit(r, t_min, t_max, rec);
        RayLog *log = g_log;
        if (!log)
            return ok;
        const bool closest = !(t_max < infinity);
        if (closest)
            log->n_closest++;
        else
            log->n_shadow++;
        int origin_prim = -1;
        for (int s = 0; s < 2; ++s)
            if (log->vprim[s] >= 0 && log->vp[s][0] == r.origin()[0] &&
                log->vp[s][1] == r.origin()[1] && log->vp[s][2] == r.origin()[2]) {
                origin_prim = log->vprim[s];
                break;
            }
        if (log->rays.size() < log->limit) {
            rtb_ray q{};
            for (int k = 0; k < 3; ++k) {
                q.o[k] = r.origin()[k];
                q.d[k] = r.direction()[k];
            }
            q.time = r.time();
            q.t_min = t_min;
            q.t_max = t_max;
            q.origin_prim = origin_prim;
            q.reserved = int32_t(rng_state);
            rtb_hit out;
            fill_hit(h, ok, rec, g_last_prim, out);
            log->rays.push_back(q);
            log->hits.push_back(out);
        }
        if (ok && closest) {
            log->vprim[1] = log->vprim[0];
            std::memcpy(log->vp[1], log->vp[0], sizeof(log->vp[0]));
            log->vprim[0] = g_last_prim;
            for (int k = 0; k < 3; ++k)
                log->vp[0][k] = rec.p[k];
        }
        return ok;
    }
    bool bounding_box(double t0, double t1, aabb &out) const override {
        return inner->bounding_box(t0, t1, out);
    }
    shared_ptr<hittable> inner;
    const SceneHandle *h;
};

// ---------------------------------------------------------------------------
// graph walk -> flat tables
// ---------------------------------------------------------------------------
struct Walker {
    rtb::SceneTables T;
    SceneHandle *H;
    std::map<const texture *, int> tex_ids;
    std::map<const perlin *, int> perlin_ids;
    std::vector<const hittable *> wrappers; // current wrapper path, outermost first
    std::vector<rtb_xform_op> ops;          // their parameters
    std::map<std::vector<const hittable *>, int> chain_ids;

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
        const int id = int(T.chains.size());
        T.chains.push_back(c);
        chain_ids[wrappers] = id;
        return id;
    }

    int texture_id(const texture *t) {
        if (!t)
            return -1;
        auto it = tex_ids.find(t);
        if (it != tex_ids.end())
            return it->second;
        rtb_texture r{};
        r.even = r.odd = r.image = r.perlin = -1;
        if (auto s = dynamic_cast<const solid_color *>(t)) {
            r.type = RTB_TEX_SOLID;
            for (int k = 0; k < 3; ++k)
                r.color[k] = s->color_value[k];
        } else if (auto c = dynamic_cast<const checker_texture *>(t)) {
            r.type = RTB_TEX_CHECKER;
            r.even = texture_id(c->even.get());
            r.odd = texture_id(c->odd.get());
        } else if (auto im = dynamic_cast<const image_texture *>(t)) {
            r.type = RTB_TEX_IMAGE;
            rtb_image d{};
            d.offset = T.image_bytes.size();
            if (im->data) {
                d.width = im->width;
                d.height = im->height;
                T.image_bytes.insert(T.image_bytes.end(), im->data,
                                     im->data + size_t(im->width) * im->height * 3);
            }
            r.image = int(T.images.size());
            T.images.push_back(d);
        } else if (auto n = dynamic_cast<const noise_texture *>(t)) {
            r.type = RTB_TEX_NOISE;
            r.scale = n->scale;
            const perlin *pn = &n->noise;
            auto pit = perlin_ids.find(pn);
            if (pit == perlin_ids.end()) {
                rtb_perlin P{};
                for (int i = 0; i < 256; ++i) {
                    for (int k = 0; k < 3; ++k)
                        P.ranvec[i][k] = pn->ranvec[i][k];
                    P.perm_x[i] = pn->perm_x[i];
                    P.perm_y[i] = pn->perm_y[i];
                    P.perm_z[i] = pn->perm_z[i];
                }
                perlin_ids[pn] = int(T.perlins.size());
                r.perlin = int(T.perlins.size());
                T.perlins.push_back(P);
            } else
                r.perlin = pit->second;
        } else
            throw std::runtime_error("walker: unknown texture class");
        const int id = int(T.textures.size());
        T.textures.push_back(r);
        H->textures.push_back(t);
        tex_ids[t] = id;
        return id;
    }

    int material_id(const material *m) {
        auto it = H->mat_ids.find(m);
        if (it != H->mat_ids.end())
            return it->second;
        rtb_material r{};
        for (int k = 0; k < 4; ++k)
            r.tex[k] = -1;
        if (auto l = dynamic_cast<const lambertian *>(m)) {
            r.type = RTB_MAT_LAMBERTIAN;
            r.tex[0] = texture_id(l->albedo.get());
        } else if (auto me = dynamic_cast<const metal *>(m)) {
            r.type = RTB_MAT_METAL;
            for (int k = 0; k < 3; ++k)
                r.color[k] = me->albedo[k];
            r.fuzz = me->fuzz;
        } else if (auto d = dynamic_cast<const dielectric *>(m)) {
            r.type = RTB_MAT_DIELECTRIC;
            r.ir = d->ir;
        } else if (auto e = dynamic_cast<const diffuse_light *>(m)) {
            r.type = RTB_MAT_DIFFUSE_LIGHT;
            r.tex[0] = texture_id(e->emit.get());
        } else if (auto p = dynamic_cast<const PBRMaterial *>(m)) {
            r.type = RTB_MAT_PBR;
            r.tex[0] = texture_id(p->albedo.get());
            r.tex[1] = texture_id(p->roughness.get());
            r.tex[2] = texture_id(p->metallic.get());
            r.tex[3] = texture_id(p->normal_map.get());
        } else if (auto i = dynamic_cast<const isotropic *>(m)) {
            r.type = RTB_MAT_ISOTROPIC;
            r.tex[0] = texture_id(i->albedo.get());
        } else
            throw std::runtime_error("walker: unknown material class");
        const int id = int(T.materials.size());
        T.materials.push_back(r);
        H->materials.push_back(m);
        H->mat_ids[m] = id;
        return id;
    }

    int add_prim(int type, const material *m, int flags, const double *d, int nd) {
        rtb_prim p{};
        p.type = type;
        p.material = material_id(m);
        p.chain = chain_id();
        p.flags = flags;
        for (int k = 0; k < nd; ++k)
            p.d[k] = d[k];
        T.prims.push_back(p);
        return int(T.prims.size()) - 1;
    }

    void tag(shared_ptr<hittable> &slot, int id, bool boundary) {
        if (!boundary)
            slot = make_shared<id_tagger>(slot, id);
    }

    void push(const hittable *w, int kind, double a, double b, double c) {
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

    // `slot` is the owning pointer inside the parent so leaves can be re-pointed
    // at their tagger.  boundary = we are inside a constant_medium's boundary.
    void walk(shared_ptr<hittable> &slot, bool boundary, int flags) {
        hittable *o = slot.get();
        if (!o)
            throw std::runtime_error("walker: null hittable");
        if (auto n = dynamic_cast<bvh_node *>(o)) {
            // a negative-radius sphere directly below this node is only ever tested for rays that pass
            // n->box (its own reported box is inverted, sphere.h:62-66): recorded as the sphere's gate
            auto gated = [&](shared_ptr<hittable> &child) {
                auto s = dynamic_cast<sphere *>(child.get());
                return s && s->radius < 0;
            };
            auto add_gate = [&](int prim) {
                rtb_gate g{};
                g.prim = prim;
                for (int k = 0; k < 3; ++k) {
                    g.lo[k] = n->box.min()[k];
                    g.hi[k] = n->box.max()[k];
                }
                T.prims[prim].flags |= RTB_PRIM_GATED;
                T.gates.push_back(g);
            };
            const bool gl = gated(n->left), gr = gated(n->right);
            if (n->left.get() == n->right.get()) {
                walk(n->left, boundary, flags | RTB_PRIM_DUP_LEAF);
                n->right = n->left;
                if (gl)
                    add_gate(int(T.prims.size()) - 1);
            } else {
                walk(n->left, boundary, flags);
                if (gl)
                    add_gate(int(T.prims.size()) - 1);
                walk(n->right, boundary, flags);
                if (gr)
                    add_gate(int(T.prims.size()) - 1);
            }
        } else if (auto l = dynamic_cast<hittable_list *>(o)) {
            for (auto &c : l->objects)
                walk(c, boundary, flags);
        } else if (auto b = dynamic_cast<box *>(o)) {
            for (auto &c : b->sides.objects)
                walk(c, boundary, flags);
        } else if (auto t = dynamic_cast<translate *>(o)) {
            push(t, RTB_XF_TRANSLATE, t->offset[0], t->offset[1], t->offset[2]);
            walk(t->ptr, boundary, flags);
            pop();
        } else if (auto r = dynamic_cast<rotate_y *>(o)) {
            push(r, RTB_XF_ROTATE_Y, r->sin_theta, r->cos_theta, 0);
            walk(r->ptr, boundary, flags);
            pop();
        } else if (auto f = dynamic_cast<flip_face *>(o)) {
            push(f, RTB_XF_FLIP_FACE, 0, 0, 0);
            walk(f->ptr, boundary, flags);
            pop();
        } else if (auto s = dynamic_cast<sphere *>(o)) {
            const double d[4] = {s->center[0], s->center[1], s->center[2], s->radius};
            tag(slot, add_prim(RTB_PRIM_SPHERE, s->mat_ptr.get(), flags, d, 4), boundary);
        } else if (auto ms = dynamic_cast<moving_sphere *>(o)) {
            const double d[9] = {ms->center0[0], ms->center0[1], ms->center0[2],
                                 ms->center1[0], ms->center1[1], ms->center1[2],
                                 ms->time0,      ms->time1,      ms->radius};
            tag(slot, add_prim(RTB_PRIM_MOVING_SPHERE, ms->mat_ptr.get(), flags, d, 9), boundary);
        } else if (auto xy = dynamic_cast<xy_rect *>(o)) {
            const double d[5] = {xy->x0, xy->x1, xy->y0, xy->y1, xy->k};
            tag(slot, add_prim(RTB_PRIM_XY_RECT, xy->mp.get(), flags, d, 5), boundary);
        } else if (auto xz = dynamic_cast<xz_rect *>(o)) {
            const double d[5] = {xz->x0, xz->x1, xz->z0, xz->z1, xz->k};
            tag(slot, add_prim(RTB_PRIM_XZ_RECT, xz->mp.get(), flags, d, 5), boundary);
        } else if (auto yz = dynamic_cast<yz_rect *>(o)) {
            const double d[5] = {yz->y0, yz->y1, yz->z0, yz->z1, yz->k};
            tag(slot, add_prim(RTB_PRIM_YZ_RECT, yz->mp.get(), flags, d, 5), boundary);
        } else if (auto cm = dynamic_cast<constant_medium *>(o)) {
            if (boundary)
                throw std::runtime_error("walker: medium inside a medium boundary");
            const int first = int(T.prims.size());
            walk(cm->boundary, true, RTB_PRIM_BOUNDARY_ONLY);
            const int count = int(T.prims.size()) - first;
            const double d[1] = {cm->neg_inv_density};
            const int id = add_prim(RTB_PRIM_MEDIUM, cm->phase_function.get(), flags, d, 1);
            T.prims[id].aux0 = first;
            T.prims[id].aux1 = count;
            tag(slot, id, false);
        } else
            throw std::runtime_error("walker: unknown hittable class");
    }

    void add_light(const Light *l) {
        rtb_light r{};
        if (auto q = dynamic_cast<const QuadLight *>(l)) {
            r.type = RTB_LIGHT_QUAD;
            for (int k = 0; k < 3; ++k) {
                r.Q[k] = q->Q[k];
                r.u[k] = q->u[k];
                r.v[k] = q->v[k];
                r.intensity[k] = q->intensity[k];
            }
        } else if (auto p = dynamic_cast<const PointLight *>(l)) {
            r.type = RTB_LIGHT_POINT;
            for (int k = 0; k < 3; ++k) {
                r.Q[k] = p->m_position[k];
                r.intensity[k] = p->m_intensity[k];
            }
        } else if (auto s = dynamic_cast<const SpotLight *>(l)) {
            r.type = RTB_LIGHT_SPOT;
            for (int k = 0; k < 3; ++k) {
                r.Q[k] = s->position[k];
                r.u[k] = s->direction[k];
                r.intensity[k] = s->intensity[k];
            }
            r.cos_cutoff = s->cos_cutoff;
        } else if (auto d = dynamic_cast<const DirectionalLight *>(l)) {
            r.type = RTB_LIGHT_DIRECTIONAL;
            for (int k = 0; k < 3; ++k) {
                r.u[k] = d->direction[k];
                r.intensity[k] = d->L[k];
            }
        } else if (auto e = dynamic_cast<const EnvironmentLight *>(l)) {
            r.type = RTB_LIGHT_ENV;
            r.env_width = e->width;
            r.env_height = e->height;
            r.env_is_probe = e->is_light_probe ? 1 : 0;
            r.env_offset = T.env_texels.size();
            T.env_texels.insert(T.env_texels.end(), e->hdr_data.begin(), e->hdr_data.end());
        } else
            throw std::runtime_error("walker: unknown light class");
        T.lights.push_back(r);
    }
};

shared_ptr<Integrator> make_integrator(int id) {
    // the switch of src/main.cpp:81-100
    shared_ptr<Integrator> it;
    switch (id) {
    case 0: it = make_shared<PathIntegrator>(); break;
    case 1: it = make_shared<RRPathInterator>(); break;
    case 2: it = make_shared<PBRPathIntegrator>(); break;
    case 3: it = make_shared<DirectLightIntegrator>(); break;
    default: it = make_shared<MISPathIntegrator>(); break;
    }
    return it;
}

shared_ptr<camera> make_camera(const SceneConfig &c) {
    // src/main.cpp:63-66 with RenderConfig::kShutterOpen/Close = 0/1 (main.cpp:45-46)
    return make_shared<camera>(c.lookfrom, c.lookat, c.vup, c.vfov, c.aspect_ratio, c.aperture,
                               c.focus_dist, 0.0, 1.0);
}

hit_record make_rec(const rtb_bsdf_query &q, material *m) {
    hit_record rec;
    rec.p = point3(q.p[0], q.p[1], q.p[2]);
    rec.normal = vec3(q.normal[0], q.normal[1], q.normal[2]);
    rec.u = q.u;
    rec.v = q.v;
    rec.t = 1.0;
    rec.front_face = q.front_face != 0;
    rec.mat_ptr = m;
    return rec;
}

} // namespace

extern "C" {

// Builds select_scene(scene_id), walks + tags it.  env_hdr_path (may be NULL)
// is only used for scene ids whose builder loads an .hdr by a fixed name: the
// harness chdir()s nowhere and passes nothing — missing assets take the
// reference's own fallbacks.  Returns NULL on failure (message on stderr).
void *ref_scene_create(int scene_id) {
    auto *h = new SceneHandle();
    try {
        h->config = select_scene(scene_id);
        if (!h->config.world)
            throw std::runtime_error("select_scene returned no world");
        h->cam = make_camera(h->config);
        h->width = h->config.image_width;
        h->height = static_cast<int>(h->width / h->config.aspect_ratio); // main.cpp:68-69
        Walker w;
        w.H = h;
        w.walk(h->config.world, false, 0);
        for (const auto &l : h->config.lights)
            w.add_light(l.get());
        rtb_globals &g = w.T.globals;
        for (int k = 0; k < 3; ++k)
            g.background[k] = h->config.background[k];
        g.image_width = h->width;
        g.image_height = h->height;
        g.samples_per_pixel = h->config.samples_per_pixel;
        g.scene_id = scene_id;
        rtb_camera &c = w.T.camera;
        for (int k = 0; k < 3; ++k) {
            c.lookfrom[k] = h->config.lookfrom[k];
            c.lookat[k] = h->config.lookat[k];
            c.vup[k] = h->config.vup[k];
        }
        c.vfov = h->config.vfov;
        c.aspect_ratio = h->config.aspect_ratio;
        c.aperture = h->config.aperture;
        c.focus_dist = h->config.focus_dist;
        c.time0 = 0.0;
        c.time1 = 1.0;
        h->blob = w.T.serialise();
    } catch (const std::exception &e) {
        std::fprintf(stderr, "ref_scene_create(%d): %s\n", scene_id, e.what());
        delete h;
        return nullptr;
    }
    return h;
}

void ref_scene_destroy(void *hv) { delete static_cast<SceneHandle *>(hv); }

const void *ref_scene_blob(void *hv, uint64_t *nbytes) {
    auto *h = static_cast<SceneHandle *>(hv);
    *nbytes = h->blob.size();
    return h->blob.data();
}

// camera.h:44-50 derived members, in declaration order:
// origin, lower_left_corner, horizontal, vertical, u, v, w (7x3), lens_radius, time0, time1
void ref_camera_derived(void *hv, double out[24]) {
    auto *h = static_cast<SceneHandle *>(hv);
    const camera &c = *h->cam;
    const vec3 *vs[7] = {&c.origin, &c.lower_left_corner, &c.horizontal, &c.vertical,
                         &c.u,      &c.v,                 &c.w};
    for (int i = 0; i < 7; ++i)
        for (int k = 0; k < 3; ++k)
            out[3 * i + k] = (*vs[i])[k];
    out[21] = c.lens_radius;
    out[22] = c.time0;
    out[23] = c.time1;
}

// hittable::hit() of the (tagged) reference world on caller rays.
void ref_trace_batch(void *hv, const rtb_ray *rays, uint64_t n, rtb_hit *hits) {
    auto *h = static_cast<SceneHandle *>(hv);
    for (uint64_t i = 0; i < n; ++i) {
        const rtb_ray &q = rays[i];
        ray r(point3(q.o[0], q.o[1], q.o[2]), vec3(q.d[0], q.d[1], q.d[2]), q.time);
        hit_record rec;
        g_last_prim = -1;
        const bool ok = h->config.world->hit(r, q.t_min, q.t_max, rec);
        fill_hit(h, ok, rec, g_last_prim, hits[i]);
    }
}

// Runs Integrator::Li on n_paths camera samples (pixels drawn uniformly with the
// reference RNG; u,v as renderer.h:73-74) and logs up to max_rays hit() queries
// with their reference answers.  counters = {closest-hit queries, shadow queries}.
uint64_t ref_record_rays(void *hv, int integrator_id, uint64_t n_paths, uint64_t max_rays,
                         rtb_ray *rays, rtb_hit *hits, uint64_t counters[2]) {
    auto *h = static_cast<SceneHandle *>(hv);
    auto integ = make_integrator(integrator_id);
    integ->set_max_depth(50); // main.cpp:102
    recorder rec_world(h->config.world, h);
    RayLog log;
    log.limit = max_rays;
    log.rays.reserve(max_rays);
    log.hits.reserve(max_rays);
    g_log = &log;
    for (uint64_t s = 0; s < n_paths; ++s) {
        const int i = random_int(0, h->width - 1);
        const int j = random_int(0, h->height - 1);
        const double u = (i + random_double()) / (h->width - 1);
        const double v = (j + random_double()) / (h->height - 1);
        ray r = h->cam->get_ray(u, v);
        log.new_path();
        (void)integ->Li(r, rec_world, h->config.background, h->config.lights);
    }
    g_log = nullptr;
    const uint64_t n = log.rays.size();
    if (n) {
        std::memcpy(rays, log.rays.data(), n * sizeof(rtb_ray));
        std::memcpy(hits, log.hits.data(), n * sizeof(rtb_hit));
    }
    if (counters) {
        counters[0] = log.n_closest;
        counters[1] = log.n_shadow;
    }
    return n;
}

// material::eval / pdf / emitted on caller-supplied records (material.h:27-56).
void ref_bsdf_eval(void *hv, int material, const rtb_bsdf_query *q, uint64_t n,
                   rtb_bsdf_value *out) {
    auto *h = static_cast<SceneHandle *>(hv);
    if (material < 0 || size_t(material) >= h->materials.size())
        return;
    auto *m = const_cast<::material *>(h->materials[material]);
    for (uint64_t i = 0; i < n; ++i) {
        hit_record rec = make_rec(q[i], m);
        const vec3 wo(q[i].wo[0], q[i].wo[1], q[i].wo[2]);
        const vec3 wi(q[i].wi[0], q[i].wi[1], q[i].wi[2]);
        const color f = m->eval(rec, wo, wi);
        const color e0 = m->emitted(rec.u, rec.v, rec.p);
        const color e1 = m->emitted(rec, wo);
        for (int k = 0; k < 3; ++k) {
            out[i].f[k] = f[k];
            out[i].emitted_old[k] = e0[k];
            out[i].emitted_new[k] = e1[k];
        }
        out[i].pdf = m->pdf(rec, wo, wi);
    }
}

// material::sample (material.h:41-44) and legacy scatter (material.h:66-69) with the
// reference's own RNG; q[i].wi is ignored; the incoming ray direction for
// scatter() is -wo.
void ref_bsdf_sample(void *hv, int material, const rtb_bsdf_query *q, uint64_t n,
                     rtb_bsdf_sample *out) {
    auto *h = static_cast<SceneHandle *>(hv);
    if (material < 0 || size_t(material) >= h->materials.size())
        return;
    auto *m = const_cast<::material *>(h->materials[material]);
    for (uint64_t i = 0; i < n; ++i) {
        std::memset(&out[i], 0, sizeof(out[i]));
        hit_record rec = make_rec(q[i], m);
        const vec3 wo(q[i].wo[0], q[i].wo[1], q[i].wo[2]);
        BSDFSample bs;
        bs.pdf = 0;
        bs.is_specular = false;
        const bool ok = m->sample(rec, wo, bs);
        out[i].ok = ok;
        if (ok) {
            for (int k = 0; k < 3; ++k) {
                out[i].wi[k] = bs.wi[k];
                out[i].f[k] = bs.f[k];
            }
            out[i].pdf = bs.pdf;
            out[i].is_specular = bs.is_specular;
        }
        ray r_in(rec.p + wo, -wo, 0.0);
        ray scattered;
        color atten;
        const bool sok = m->scatter(r_in, rec, atten, scattered);
        out[i].scatter_ok = sok;
        if (sok)
            for (int k = 0; k < 3; ++k) {
                out[i].scatter_dir[k] = scattered.direction()[k];
                out[i].scatter_atten[k] = atten[k];
            }
    }
}

// texture::value (texture.h:13)
void ref_texture_value(void *hv, int texture, const double *uvp /* n x 5: u v px py pz */,
                       uint64_t n, double *rgb) {
    auto *h = static_cast<SceneHandle *>(hv);
    if (texture < 0 || size_t(texture) >= h->textures.size())
        return;
    const ::texture *t = h->textures[texture];
    for (uint64_t i = 0; i < n; ++i) {
        const double *q = uvp + 5 * i;
        const color c = t->value(q[0], q[1], point3(q[2], q[3], q[4]));
        for (int k = 0; k < 3; ++k)
            rgb[3 * i + k] = c[k];
    }
}

// Light::sample / pdf / Le (light.h:21-41)
void ref_light_eval(void *hv, int light, const rtb_light_query *q, uint64_t n,
                    rtb_light_value *out) {
    auto *h = static_cast<SceneHandle *>(hv);
    if (light < 0 || size_t(light) >= h->config.lights.size())
        return;
    const Light &L = *h->config.lights[light];
    for (uint64_t i = 0; i < n; ++i) {
        std::memset(&out[i], 0, sizeof(out[i]));
        const point3 p(q[i].p[0], q[i].p[1], q[i].p[2]);
        const vec3 d(q[i].d[0], q[i].d[1], q[i].d[2]);
        LightSample ls = L.sample(p, vec2(q[i].u[0], q[i].u[1]));
        for (int k = 0; k < 3; ++k) {
            out[i].Li[k] = ls.Li[k];
            out[i].wi[k] = ls.wi[k];
        }
        out[i].pdf = ls.pdf;
        out[i].dist = ls.dist;
        out[i].is_delta = ls.is_delta;
        out[i].pdf_dir = L.pdf(p, d);
        const color le = L.Le(ray(p, d));
        for (int k = 0; k < 3; ++k)
            out[i].Le[k] = le[k];
    }
}

int ref_light_flags(void *hv, int light) {
    auto *h = static_cast<SceneHandle *>(hv);
    if (light < 0 || size_t(light) >= h->config.lights.size())
        return -1;
    const Light &L = *h->config.lights[light];
    return (L.is_delta() ? 1 : 0) | (L.is_infinite() ? 2 : 0);
}

// Per-pixel sum and sum of squares of linear Li for spp samples per pixel at a
// caller-chosen resolution (the camera depends only on the aspect ratio, so any
// width/height with the scene's aspect samples the same image).  Li is called
// directly, exactly as renderer.h:72-79 does, but the sqrt/clamp of
// renderer.h:126-140 is bypassed.  sum/sumsq: height x width x 3 doubles, row 0 =
// bottom row (j = 0), as in RenderBuffer.  counters = {closest, shadow} queries.
void ref_render_linear(void *hv, int integrator_id, int width, int height, int spp, int max_depth,
                       int n_threads, double *sum, double *sumsq, uint64_t counters[2]) {
    auto *h = static_cast<SceneHandle *>(hv);
    if (n_threads <= 0)
        n_threads = int(std::thread::hardware_concurrency());
    std::atomic<int> next_row(0);
    std::atomic<uint64_t> n_closest(0), n_shadow(0);
    auto worker = [&]() {
        auto integ = make_integrator(integrator_id);
        integ->set_max_depth(max_depth);
        recorder rec_world(h->config.world, h);
        RayLog log; // count only
        g_log = &log;
        while (true) {
            const int j = next_row.fetch_add(1);
            if (j >= height)
                break;
            for (int i = 0; i < width; ++i) {
                double s[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
                for (int k = 0; k < spp; ++k) {
                    const double u = (i + random_double()) / (width - 1);
                    const double v = (j + random_double()) / (height - 1);
                    ray r = h->cam->get_ray(u, v);
                    const color c =
                        integ->Li(r, rec_world, h->config.background, h->config.lights);
                    for (int ch = 0; ch < 3; ++ch) {
                        s[ch] += c[ch];
                        s2[ch] += c[ch] * c[ch];
                    }
                }
                for (int ch = 0; ch < 3; ++ch) {
                    sum[(size_t(j) * width + i) * 3 + ch] = s[ch];
                    sumsq[(size_t(j) * width + i) * 3 + ch] = s2[ch];
                }
            }
        }
        g_log = nullptr;
        n_closest += log.n_closest;
        n_shadow += log.n_shadow;
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back(worker);
    for (auto &t : th)
        t.join();
    if (counters) {
        counters[0] = n_closest;
        counters[1] = n_shadow;
    }
}

// The reference's own Renderer::render (renderer.h:30-102), unmodified, on a
// FRESH untagged scene, all hardware threads (its own policy, renderer.h:48).
// width <= 0 keeps the scene's width; spp <= 0 keeps the scene's spp.  Returns
// wall seconds of the render() call; out (optional) receives the RenderBuffer
// contents (gamma-encoded, clamped, row 0 = bottom), height x width x 3 doubles.
double ref_render_timed(int scene_id, int integrator_id, int width, int spp, int max_depth,
                        int *out_w, int *out_h, double *out, uint64_t out_capacity) {
    SceneConfig config = select_scene(scene_id);
    auto cam = make_camera(config);
    const int w = width > 0 ? width : config.image_width;
    const int hgt = static_cast<int>(w / config.aspect_ratio);
    RenderBuffer buffer(w, hgt);
    Renderer renderer;
    renderer.set_samples(spp > 0 ? spp : config.samples_per_pixel);
    renderer.set_integrator(make_integrator(integrator_id));
    renderer.set_max_depth(max_depth);
    const auto t0 = std::chrono::high_resolution_clock::now();
    renderer.render(config.world, cam, config.background, buffer, config.lights);
    const auto t1 = std::chrono::high_resolution_clock::now();
    if (out_w)
        *out_w = w;
    if (out_h)
        *out_h = hgt;
    if (out && out_capacity >= uint64_t(w) * hgt * 3) {
        const auto &px = buffer.get_data();
        for (int j = 0; j < hgt; ++j)
            for (int i = 0; i < w; ++i)
                for (int k = 0; k < 3; ++k)
                    out[(size_t(j) * w + i) * 3 + k] = px[j][i][k];
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

// The reference's OUTPUT path on caller-supplied sums: Renderer::write_color_to_buffer (renderer.h:126-140)
// for every pixel, then RenderBuffer::save_to_png (render_buffer.h:35-55) to `png_path`.  sums: height x
// width x 3 linear sums, row 0 = the RenderBuffer's row 0 (bottom of the image).  buffer_out (optional):
// the RenderBuffer's contents, same layout.
int ref_output_path(int w, int h, const double *sums, int samples, const char *png_path, double *buffer_out) {
    RenderBuffer buffer(w, h);
    Renderer renderer;
    for (int j = 0; j < h; ++j)
        for (int i = 0; i < w; ++i) {
            const double *s = sums + (size_t(j) * w + i) * 3;
            renderer.write_color_to_buffer(buffer, i, j, color(s[0], s[1], s[2]), samples);
        }
    if (buffer_out) {
        const auto &px = buffer.get_data();
        for (int j = 0; j < h; ++j)
            for (int i = 0; i < w; ++i)
                for (int k = 0; k < 3; ++k)
                    buffer_out[(size_t(j) * w + i) * 3 + k] = px[j][i][k];
    }
    return png_path ? (buffer.save_to_png(png_path) ? 1 : 0) : 1;
}

int ref_hardware_threads() { return int(std::thread::hardware_concurrency()); }

} // extern "C"
