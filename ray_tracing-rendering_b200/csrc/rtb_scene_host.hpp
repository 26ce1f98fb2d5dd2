// rtb_scene_host.hpp — scene blob (include/rtb200_scene.h) -> host copies of the
// device tables: leaf-ordered primitive records in fp32 and fp64, the two-level
// BVH, wrapper chains, materials, textures, lights (+ env-map CDF tables) and the
// camera.  Pure host C++ (no CUDA calls) so it can be unit-tested without a GPU.
#ifndef RTB_SCENE_HOST_HPP
#define RTB_SCENE_HOST_HPP

#include "rtb200_blob.hpp"
#include "rtb_bvh.hpp"
#include "rtb_shading.cuh"
#include "rtb_wide.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace rtb {

template <class R> struct TypedTables {
    std::vector<PrimT<R>> prims;
    std::vector<MovingAux<R>> maux;
    std::vector<XfOp<R>> ops;
    std::vector<MatT<R>> mats;
    std::vector<TexT<R>> texs;
    std::vector<PerlinT<R>> perlins;
    std::vector<LightT<R>> lights;
    CameraT<R> camera;
};

struct HostScene {
    TypedTables<float> f32;
    TypedTables<double> f64;
    std::vector<Node32> nodes;
    WideTree wide; // the same trees collapsed into 128-byte 4-wide nodes (the production traversal's)
    std::vector<ChainRec> chains;
    std::vector<ChainAffine> affine;     // per chain, fp32 production path
    std::vector<int32_t> prim_chain;     // per sorted prim
    std::vector<int32_t> prim_orig;      // per sorted prim -> blob prim id (-1: instance)
    std::vector<int32_t> orig_to_sorted; // blob prim id -> sorted index
    std::vector<ImageRec> images;
    std::vector<uint8_t> image_bytes;
    std::vector<float> env_texels;           // EMPTY when env_on_device: the texels go from the caller's blob to the device
    const float *env_texels_src = nullptr;   // (env_on_device) the blob's texel pool: valid DURING the upload call only
    uint64_t n_env_texels = 0;
    std::vector<double> env_tables;          // Distribution2D tables of the env lights; EMPTY when env_on_device
    uint64_t env_table_doubles = 0;          // their size (the device builds them: rtb_api.cu build_env_tables_device)
    bool env_on_device = false;
    rtb_globals globals{};
    int n_infinite_lights = 0;
    uint32_t mat_type_mask = 0; // bit t set: some primitive uses a material of type t
    bool textured_lambertian = false; // some lambertian in use has a non-solid (noise / image / checker) albedo
    bool has_media = false;
    int n_instances = 0;
    int n_top_items = 0; // primitives + instance records of the top level = prims[0 .. n_top_items)
    uint32_t root_ref = kEmptyRef; // ref of the top-level root
    int n_boxes = 0;       // grouped `box` objects (PT_BOX records; their faces sit at the end of the world slots)
    int n_world_slots = 0; // sorted primitives from here on are medium boundaries only (GeomView::n_world)
    std::vector<uint32_t> gated; // sorted indices of the PT_GATED spheres, ascending blob index (GeomView::gated)
    bool flat_ok = false; // small enough for the lockstep / shared-memory traversal
    bool has_f64 = true;  // f64.prims / maux / mats / texs were built (see build_host_scene)
    std::vector<char> blob_copy; // the scene blob, kept only while has_f64 is false (to build those tables later)
};

namespace detail {

inline void xf_point_to_world(const rtb_xform_op *ops, int first, int count, double p[3]) {
    // inverse of the ray transform: what the wrappers do to rec.p on the way out
    // (hittable.h:58, :145-146), innermost wrapper first
    for (int i = count - 1; i >= 0; --i) {
        const rtb_xform_op &op = ops[first + i];
        if (op.kind == RTB_XF_TRANSLATE) {
            p[0] += op.a;
            p[1] += op.b;
            p[2] += op.c;
        } else if (op.kind == RTB_XF_ROTATE_Y) {
            const double x = op.b * p[0] + op.a * p[2];
            const double z = -op.a * p[0] + op.b * p[2];
            p[0] = x;
            p[2] = z;
        }
    }
}

inline Box box_to_world(const Box &b, const rtb_xform_op *ops, int first, int count) {
    Box w;
    for (int c = 0; c < 8; ++c) {
        double p[3] = {(c & 1) ? b.hi[0] : b.lo[0], (c & 2) ? b.hi[1] : b.lo[1],
                       (c & 4) ? b.hi[2] : b.lo[2]};
        xf_point_to_world(ops, first, count, p);
        w.grow(p);
    }
    return w;
}

// object-space bounds of a simple primitive (sphere.h:62-66, moving_sphere.h:64-71,
// aarect.h:24-28 with its 1e-4 padding)
inline Box prim_box(const rtb_prim &p, const rtb_camera &cam) {
    Box b;
    switch (p.type) {
    case RTB_PRIM_SPHERE: {
        const double r = std::fabs(p.d[3]);
        for (int k = 0; k < 3; ++k) {
            b.lo[k] = p.d[k] - r;
            b.hi[k] = p.d[k] + r;
        }
        break;
    }
    case RTB_PRIM_MOVING_SPHERE: {
        const double r = std::fabs(p.d[8]);
        // ray times lie in the shutter interval; shadow rays use time 0
        // (direct_light_integrator.h:115)
        const double times[3] = {cam.time0, cam.time1, 0.0};
        for (double tm : times) {
            const double s = (tm - p.d[6]) / (p.d[7] - p.d[6]);
            for (int k = 0; k < 3; ++k) {
                const double c = p.d[k] + s * (p.d[3 + k] - p.d[k]);
                b.lo[k] = std::min(b.lo[k], c - r);
                b.hi[k] = std::max(b.hi[k], c + r);
            }
        }
        break;
    }
    case RTB_PRIM_XY_RECT:
    case RTB_PRIM_XZ_RECT:
    case RTB_PRIM_YZ_RECT: {
        const int AX = p.type == RTB_PRIM_XY_RECT ? 2 : (p.type == RTB_PRIM_XZ_RECT ? 1 : 0);
        const int A = p.type == RTB_PRIM_YZ_RECT ? 1 : 0;
        const int B = p.type == RTB_PRIM_XY_RECT ? 1 : 2;
        b.lo[A] = std::min(p.d[0], p.d[1]);
        b.hi[A] = std::max(p.d[0], p.d[1]);
        b.lo[B] = std::min(p.d[2], p.d[3]);
        b.hi[B] = std::max(p.d[2], p.d[3]);
        b.lo[AX] = p.d[4] - 0.0001;
        b.hi[AX] = p.d[4] + 0.0001;
        break;
    }
    default:
        throw std::runtime_error("prim_box: not a simple primitive");
    }
    return b;
}

template <class R> PrimT<R> make_prim(const rtb_prim &p, uint32_t aux, uint32_t aux2) {
    PrimT<R> q;
    std::memset(&q, 0, sizeof(q));
    uint32_t type = 0;
    switch (p.type) {
    case RTB_PRIM_SPHERE:
        type = PT_SPHERE;
        for (int k = 0; k < 4; ++k)
            q.d[k] = R(p.d[k]);
        break;
    case RTB_PRIM_MOVING_SPHERE:
        type = PT_MSPHERE;
        for (int k = 0; k < 3; ++k)
            q.d[k] = R(p.d[k]);
        q.d[3] = R(p.d[8]);
        break;
    case RTB_PRIM_XY_RECT:
    case RTB_PRIM_XZ_RECT:
    case RTB_PRIM_YZ_RECT:
        type = p.type == RTB_PRIM_XY_RECT ? PT_XY : (p.type == RTB_PRIM_XZ_RECT ? PT_XZ : PT_YZ);
        for (int k = 0; k < 5; ++k)
            q.d[k] = R(p.d[k]);
        break;
    case RTB_PRIM_MEDIUM:
        type = PT_MEDIUM;
        q.d[0] = R(p.d[0]);
        break;
    }
    q.type_mat = type | ((p.flags & RTB_PRIM_DUP_LEAF) ? uint32_t(PT_DUP_LEAF) : 0u) |
                 ((p.flags & RTB_PRIM_GATED) ? uint32_t(PT_GATED) : 0u) |
                 (uint32_t(p.material) << PT_MAT_SHIFT);
    q.aux = aux;
    q.aux2 = aux2;
    return q;
}

template <class R, class V> void set3(R dst[3], const V &src) {
    for (int k = 0; k < 3; ++k)
        dst[k] = R(src[k]);
}

// camera::camera, camera.h:9-30, operation by operation in fp64
inline CameraT<double> derive_camera(const rtb_camera &c) {
    typedef V3<double> V;
    const double pi = Consts<double>::pi();
    const double theta = c.vfov * pi / 180.0; // rtweekend.h:20-22
    const double h = std::tan(theta / 2);
    const double viewport_height = 2.0 * h;
    const double viewport_width = c.aspect_ratio * viewport_height;
    const V lookfrom(c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]);
    const V lookat(c.lookat[0], c.lookat[1], c.lookat[2]);
    const V vup(c.vup[0], c.vup[1], c.vup[2]);
    CameraT<double> o;
    o.w = unit_vector(lookfrom - lookat);
    o.u = unit_vector(cross(vup, o.w));
    o.v = cross(o.w, o.u);
    o.origin = lookfrom;
    o.horizontal = (c.focus_dist * viewport_width) * o.u;
    o.vertical = (c.focus_dist * viewport_height) * o.v;
    o.lower_left_corner = o.origin - o.horizontal / 2.0 - o.vertical / 2.0 - c.focus_dist * o.w;
    o.lens_radius = c.aperture / 2;
    o.time0 = c.time0;
    o.time1 = c.time1;
    return o;
}

template <class R> CameraT<R> cast_camera(const CameraT<double> &c) {
    CameraT<R> o;
    auto cv = [](const V3<double> &v) { return V3<R>(R(v.x), R(v.y), R(v.z)); };
    o.origin = cv(c.origin);
    o.lower_left_corner = cv(c.lower_left_corner);
    o.horizontal = cv(c.horizontal);
    o.vertical = cv(c.vertical);
    o.u = cv(c.u);
    o.v = cv(c.v);
    o.w = cv(c.w);
    o.lens_radius = R(c.lens_radius);
    o.time0 = R(c.time0);
    o.time1 = R(c.time1);
    return o;
}

// Distribution2D over luminance*sin(theta), environmental_light.h:146-180 and :15-27
inline void build_env_tables(const float *tex, int W, int H, std::vector<double> &out) {
    const double pi = Consts<double>::pi();
    const size_t base = out.size();
    out.resize(base + EnvTables::doubles(W, H));
    double *cond_func = out.data() + base;
    double *cond_cdf = cond_func + size_t(W) * H;
    double *cond_int = cond_cdf + size_t(W + 1) * H;
    double *marg_cdf = cond_int + H;
    double *marg_int = marg_cdf + (H + 1);
    for (int v = 0; v < H; ++v) {
        const double sin_theta = std::sin(pi * (v + 0.5) / H);
        for (int u = 0; u < W; ++u) {
            const size_t idx = size_t(v) * W + u;
            const double r = tex[3 * idx], g = tex[3 * idx + 1], b = tex[3 * idx + 2];
            const double lum = 0.2126 * r + 0.7152 * g + 0.0722 * b;
            cond_func[idx] = lum * sin_theta;
        }
        double *cdf = cond_cdf + size_t(v) * (W + 1);
        cdf[0] = 0;
        for (int i = 1; i <= W; ++i)
            cdf[i] = cdf[i - 1] + cond_func[size_t(v) * W + i - 1];
        cond_int[v] = cdf[W];
        if (cond_int[v] > 0)
            for (int i = 0; i <= W; ++i)
                cdf[i] /= cond_int[v];
    }
    marg_cdf[0] = 0;
    for (int i = 1; i <= H; ++i)
        marg_cdf[i] = marg_cdf[i - 1] + cond_int[i - 1];
    *marg_int = marg_cdf[H];
    if (*marg_int > 0)
        for (int i = 0; i <= H; ++i)
            marg_cdf[i] /= *marg_int;
}

} // namespace detail

inline bool texture_reads_uv(const rtb::SceneView &S, int tex, int depth = 0) {
    if (tex < 0 || depth > 8)
        return false;
    const rtb_texture &t = S.textures()[tex];
    if (t.type == RTB_TEX_IMAGE)
        return true;
    if (t.type == RTB_TEX_CHECKER)
        return texture_reads_uv(S, t.even, depth + 1) || texture_reads_uv(S, t.odd, depth + 1);
    return false;
}

// max_leaf: primitives per BVH leaf.
// RTB200_TIMING=1 prints the phases of build_host_scene on stderr (upload-cost work)
struct PhaseTimer {
    bool on = std::getenv("RTB200_TIMING") != nullptr;
    std::chrono::high_resolution_clock::time_point t = std::chrono::high_resolution_clock::now();
    void mark(const char *what) {
        if (!on)
            return;
        const auto n = std::chrono::high_resolution_clock::now();
        std::fprintf(stderr, "[rtb200 build] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

// want_f64 = false: the large per-primitive / per-material fp64 tables (only the validation entry
// points read them) are left empty; H.has_f64 says so and the API layer builds them on first use.
// env_on_device = true: the env lights' Distribution2D tables are only sized here; the caller builds them
// on the device from the uploaded texels (the library does; the CPU test harness keeps the host build).
inline HostScene build_host_scene(const SceneView &S, int max_leaf = 4, double trav_cost = 1.0,
                                  bool layout_dfs = false, bool want_f64 = true, bool env_on_device = false,
                                  bool group_boxes = true) {
    using namespace detail;
    PhaseTimer timer;
    S.validate();
    timer.mark("validate");
    HostScene H;
    H.globals = S.globals();
    const rtb_prim *P = S.prims();
    const int np = int(S.n_prims());
    const rtb_chain *C = S.chains();
    const rtb_xform_op *X = S.xform_ops();
    const rtb_camera &cam = S.camera();

    for (uint64_t i = 0; i < S.n_chains(); ++i) {
        if (C[i].count > kMaxChainOps)
            throw std::runtime_error("scene: wrapper chain longer than kMaxChainOps");
        H.chains.push_back(ChainRec{C[i].first, C[i].count});
        // fold the chain: start from the identity (c=1, s=0, b=0) and push each op through
        double c = 1, s = 0, b[3] = {0, 0, 0};
        for (int k = 0; k < C[i].count; ++k) {
            const rtb_xform_op &op = X[C[i].first + k];
            if (op.kind == RTB_XF_TRANSLATE) { // o' = o - offset
                b[0] -= op.a;
                b[1] -= op.b;
                b[2] -= op.c;
            } else if (op.kind == RTB_XF_ROTATE_Y) { // o' = Ry(cos, sin) o  (hittable.h:132-138)
                const double cs = op.b, sn = op.a;
                const double nc = cs * c - sn * s, ns = sn * c + cs * s;
                const double bx = cs * b[0] - sn * b[2], bz = sn * b[0] + cs * b[2];
                c = nc;
                s = ns;
                b[0] = bx;
                b[2] = bz;
            }
        }
        H.affine.push_back(ChainAffine{float(c), float(s), float(b[0]), float(b[1]), float(b[2])});
    }
    for (uint64_t i = 0; i < S.n_xform_ops(); ++i) {
        XfOp<double> d{X[i].kind, X[i].a, X[i].b, X[i].c};
        XfOp<float> f{X[i].kind, float(X[i].a), float(X[i].b), float(X[i].c)};
        H.f64.ops.push_back(d);
        H.f32.ops.push_back(f);
    }
    auto chain_moves = [&](int chain) { // does the chain transform the ray at all?
        if (chain < 0)
            return false;
        for (int i = 0; i < C[chain].count; ++i)
            if (X[C[chain].first + i].kind != RTB_XF_FLIP_FACE)
                return true;
        return false;
    };
    auto world_box_of = [&](int i) -> Box { // world-space bounds of blob prim i
        const rtb_prim &p = P[i];
        if (p.type == RTB_PRIM_MEDIUM) {
            Box b;
            for (int k = p.aux0; k < p.aux0 + p.aux1; ++k) {
                Box ob = prim_box(P[k], cam);
                if (P[k].chain >= 0)
                    ob = box_to_world(ob, X, C[P[k].chain].first, C[P[k].chain].count);
                b.grow(ob);
            }
            return b; // the medium's own (outer) chain is applied by the caller
        }
        return prim_box(p, cam);
    };

    // ---- `box` objects (box.h: six rects in a fixed order — z hi, z lo, y hi, y lo, x hi, x lo — one material,
    // no wrapper) of scenes too large for the shared-memory kernels become ONE tree item each, a PT_BOX record
    // whose fp32 test is a slab test (the fp64 paths walk its six face records one by one); the faces keep their
    // own records behind the tree-ordered ones, so hits, records and shading name the rect as before.
    // scene09's ground is 400 such boxes = 2,400 rects whose side faces coincide pairwise.
    struct BoxGroup {
        int first; // blob index of its first face
        Box box;
    };
    std::vector<BoxGroup> boxes;
    std::vector<int> box_of(size_t(np), -1);
    if (group_boxes && np > kFlatMaxPrims) {
        for (int i = 0; i + 5 < np;) {
            const rtb_prim *q = P + i;
            bool ok = q[0].type == RTB_PRIM_XY_RECT && q[1].type == RTB_PRIM_XY_RECT && q[2].type == RTB_PRIM_XZ_RECT &&
                      q[3].type == RTB_PRIM_XZ_RECT && q[4].type == RTB_PRIM_YZ_RECT && q[5].type == RTB_PRIM_YZ_RECT;
            for (int k = 0; ok && k < 6; ++k)
                ok = (q[k].flags & ~uint32_t(RTB_PRIM_DUP_LEAF)) == 0 && q[k].chain < 0 && q[k].material == q[0].material; // (a rect tested twice answers twice the same)
            if (ok) {
                const double x0 = q[0].d[0], x1 = q[0].d[1], y0 = q[0].d[2], y1 = q[0].d[3], z1 = q[0].d[4], z0 = q[1].d[4];
                ok = x0 < x1 && y0 < y1 && z0 < z1 &&
                     q[1].d[0] == x0 && q[1].d[1] == x1 && q[1].d[2] == y0 && q[1].d[3] == y1 &&
                     q[2].d[0] == x0 && q[2].d[1] == x1 && q[2].d[2] == z0 && q[2].d[3] == z1 && q[2].d[4] == y1 &&
                     q[3].d[0] == x0 && q[3].d[1] == x1 && q[3].d[2] == z0 && q[3].d[3] == z1 && q[3].d[4] == y0 &&
                     q[4].d[0] == y0 && q[4].d[1] == y1 && q[4].d[2] == z0 && q[4].d[3] == z1 && q[4].d[4] == x1 &&
                     q[5].d[0] == y0 && q[5].d[1] == y1 && q[5].d[2] == z0 && q[5].d[3] == z1 && q[5].d[4] == x0;
            }
            if (!ok) {
                ++i;
                continue;
            }
            BoxGroup b;
            b.first = i;
            for (int k = 0; k < 6; ++k) {
                b.box.grow(prim_box(q[k], cam));
                box_of[size_t(i + k)] = int(boxes.size());
            }
            boxes.push_back(b);
            i += 6;
        }
    }
    constexpr uint32_t kBoxItem = 0x80000000u; // BuildItem ids: blob primitive < np <= instance < kBoxItem <= box

    // ---- group world primitives: top-level items and one BLAS per moving chain
    std::vector<BuildItem> top;
    std::map<int, std::vector<BuildItem>> groups; // chain id -> object-space items
    for (int i = 0; i < np; ++i) {
        if (P[i].flags & RTB_PRIM_BOUNDARY_ONLY)
            continue;
        H.mat_type_mask |= 1u << S.materials()[P[i].material].type;
        {
            const rtb_material &pm = S.materials()[P[i].material];
            if (pm.type == RTB_MAT_LAMBERTIAN && pm.tex[0] >= 0 && S.textures()[pm.tex[0]].type != RTB_TEX_SOLID)
                H.textured_lambertian = true;
        }
        if (P[i].type == RTB_PRIM_MEDIUM)
            H.has_media = true;
        if (P[i].type == RTB_PRIM_MEDIUM && chain_moves(P[i].chain))
            throw std::runtime_error("scene: a constant_medium under translate/rotate_y is not supported");
        BuildItem it;
        it.solitary = false;
        if (box_of[size_t(i)] >= 0) { // a face: its box joins the top level once
            if (boxes[size_t(box_of[size_t(i)])].first == i) {
                it.box = boxes[size_t(box_of[size_t(i)])].box;
                it.id = kBoxItem | uint32_t(box_of[size_t(i)]);
                top.push_back(it);
            }
            continue;
        }
        it.box = world_box_of(i);
        it.id = uint32_t(i);
        if (chain_moves(P[i].chain))
            groups[P[i].chain].push_back(it);
        else
            top.push_back(it);
    }
    // "Global" primitives: at most kMaxGlobalPrims top-level primitives whose boxes dwarf what is left
    // of the scene without them (a ground sphere, the boundary of a world-filling fog).  They stay in
    // the binary tree (alone in their leaves) but are left out of the 4-wide tree the production
    // kernels walk, which test them up front for every ray (rtb_wide.cuh build_wide).
    std::vector<uint32_t> global_ids;
    if (top.size() > 8) {
        Box all;
        for (const BuildItem &it : top)
            all.grow(it.box);
        std::vector<size_t> cand;
        for (size_t i = 0; i < top.size(); ++i)
            if (top[i].box.area() >= 0.25 * all.area() && top[i].id < kBoxItem)
                cand.push_back(i);
        if (!cand.empty() && cand.size() <= size_t(kMaxGlobalPrims)) {
            Box rest;
            for (size_t i = 0; i < top.size(); ++i)
                if (std::find(cand.begin(), cand.end(), i) == cand.end())
                    rest.grow(top[i].box);
            if (rest.valid() && rest.area() <= 0.5 * all.area())
                for (size_t i : cand) {
                    top[i].solitary = true;
                    global_ids.push_back(top[i].id);
                }
        }
    }
    // instances join the top level with their world-space bounds
    struct Inst {
        int chain;
        Box obj_box;
    };
    std::vector<Inst> insts;
    for (auto &kv : groups) {
        Inst in;
        in.chain = kv.first;
        for (const BuildItem &it : kv.second)
            in.obj_box.grow(it.box);
        BuildItem ti;
        ti.box = box_to_world(in.obj_box, X, C[in.chain].first, C[in.chain].count);
        ti.id = uint32_t(np + insts.size()); // ids >= np denote instances
        ti.solitary = true;
        ti.instance = true;
        top.push_back(ti);
        insts.push_back(in);
    }
    H.n_instances = int(insts.size());

    // ---- build trees; primitive array = [top items][BLAS 0 prims]...[boundary prims]
    struct Slot { // what sits at each sorted position
        int orig; // blob prim id, or -1 for an instance
        int inst; // instance index, or -1
        int box;  // box index (the record of a grouped box; orig = its first face), or -1
    };
    std::vector<Slot> slots;
    timer.mark("items + bounds");
    BuildResult tlas = build_bvh(top, max_leaf, 0, 0, trav_cost, layout_dfs);
    timer.mark("binary SAH build (top level)");
    for (uint32_t id : tlas.order) {
        if (id >= kBoxItem)
            slots.push_back(Slot{boxes[id - kBoxItem].first, -1, int(id - kBoxItem)});
        else
            slots.push_back(id < uint32_t(np) ? Slot{int(id), -1, -1} : Slot{-1, int(id) - np, -1});
    }
    H.nodes = tlas.nodes;
    H.root_ref = tlas.nodes[0].ref;
    H.n_top_items = int(slots.size());
    std::vector<uint32_t> blas_root(insts.size()), blas_first(insts.size()), blas_count(insts.size());
    {
        size_t gi = 0;
        for (auto &kv : groups) {
            const uint32_t node_off = uint32_t(H.nodes.size());
            BuildResult b = build_bvh(kv.second, max_leaf, uint32_t(slots.size()), node_off, trav_cost, layout_dfs);
            blas_first[gi] = uint32_t(slots.size());
            blas_count[gi] = uint32_t(b.order.size());
            blas_root[gi++] = b.nodes[0].ref; // the instance record carries the root REF
            for (uint32_t id : b.order)
                slots.push_back(Slot{int(id), -1, -1});
            H.nodes.insert(H.nodes.end(), b.nodes.begin(), b.nodes.end());
        }
    }
    H.orig_to_sorted.assign(np, -1);
    for (size_t s = 0; s < slots.size(); ++s)
        if (slots[s].orig >= 0 && slots[s].box < 0)
            H.orig_to_sorted[slots[s].orig] = int(s);
    // the faces of grouped boxes: six consecutive records per box, behind everything the trees point at
    std::vector<uint32_t> box_faces(boxes.size());
    for (size_t b = 0; b < boxes.size(); ++b) {
        box_faces[b] = uint32_t(slots.size());
        for (int k = 0; k < 6; ++k) {
            H.orig_to_sorted[boxes[b].first + k] = int(slots.size());
            slots.push_back(Slot{boxes[b].first + k, -1, -1});
        }
    }
    H.n_boxes = int(boxes.size());
    // boundary-only prims keep their blob order (media reference them as ranges)
    H.n_world_slots = int(slots.size());
    for (int i = 0; i < np; ++i)
        if (P[i].flags & RTB_PRIM_BOUNDARY_ONLY) {
            H.orig_to_sorted[i] = int(slots.size());
            slots.push_back(Slot{i, -1, -1});
        }

    // ---- emit typed primitive records
    auto emit = [&](auto &T) {
        typedef typename std::remove_reference<decltype(T.prims[0].d[0])>::type R;
        T.prims.clear();
        T.maux.clear();
        for (const Slot &s : slots) {
            if (s.inst >= 0) {
                PrimT<R> q;
                std::memset(&q, 0, sizeof(q));
                q.type_mat = PT_INSTANCE;
                q.aux = blas_root[s.inst];
                q.aux2 = uint32_t(insts[s.inst].chain);
                // the instance's primitives as a contiguous range (used by the lockstep
                // traversal of small scenes): d[0] = first, d[1] = count
                q.d[0] = R(blas_first[s.inst]);
                q.d[1] = R(blas_count[s.inst]);
                T.prims.push_back(q);
                continue;
            }
            if (s.box >= 0) { // lo, hi (the sixth number in aux, as float bits; the fp64 paths never read them), first face
                const rtb_prim *q = P + s.orig;
                PrimT<R> b;
                std::memset(&b, 0, sizeof(b));
                b.d[0] = R(q[5].d[4]);
                b.d[1] = R(q[3].d[4]);
                b.d[2] = R(q[1].d[4]);
                b.d[3] = R(q[4].d[4]);
                b.d[4] = R(q[2].d[4]);
                const float hz = float(q[0].d[4]);
                std::memcpy(&b.aux, &hz, 4);
                b.aux2 = box_faces[size_t(s.box)];
                b.type_mat = uint32_t(PT_BOX) | (uint32_t(q[0].material) << PT_MAT_SHIFT);
                T.prims.push_back(b);
                continue;
            }
            const rtb_prim &p = P[s.orig];
            uint32_t aux = 0, aux2 = 0;
            if (p.flags & RTB_PRIM_GATED) { // validate(): a sphere with exactly one gate
                MovingAux<R> m;
                for (uint64_t k = 0; k < S.n_gates(); ++k)
                    if (S.gates()[k].prim == s.orig) {
                        const rtb_gate &gt = S.gates()[k];
                        for (int a = 0; a < 3; ++a)
                            m.c1[a] = R(gt.lo[a]);
                        m.time0 = R(gt.hi[0]);
                        m.time1 = R(gt.hi[1]);
                        m.extra = R(gt.hi[2]);
                    }
                aux = uint32_t(T.maux.size());
                T.maux.push_back(m);
            } else if (p.type == RTB_PRIM_MOVING_SPHERE) {
                MovingAux<R> m;
                m.extra = R(0);
                for (int k = 0; k < 3; ++k)
                    m.c1[k] = R(p.d[3 + k]);
                m.time0 = R(p.d[6]);
                m.time1 = R(p.d[7]);
                aux = uint32_t(T.maux.size());
                T.maux.push_back(m);
            } else if (p.type == RTB_PRIM_MEDIUM) {
                aux = uint32_t(H.orig_to_sorted[p.aux0]);
                aux2 = uint32_t(p.aux1);
                for (int k = 0; k < p.aux1; ++k)
                    if (H.orig_to_sorted[p.aux0 + k] != int(aux) + k)
                        throw std::runtime_error("scene: medium boundary not contiguous");
            }
            T.prims.push_back(make_prim<R>(p, aux, aux2));
        }
    };
    timer.mark("bottom levels + slots");
    emit(H.f32);
    timer.mark("emit fp32 primitives");
    H.has_f64 = want_f64;
    if (want_f64)
        emit(H.f64);
    timer.mark("emit fp64 primitives");
    for (const Slot &s : slots) {
        H.prim_orig.push_back(s.orig);
        H.prim_chain.push_back(s.inst >= 0 ? insts[s.inst].chain : (s.box >= 0 ? -1 : P[s.orig].chain));
    }

    for (uint64_t i = 0; i < S.n_prims(); ++i) // (ascending blob index = the reference's traversal order)
        if (P[i].flags & RTB_PRIM_GATED)
            H.gated.push_back(uint32_t(H.orig_to_sorted[i]));
    if (H.gated.size() > size_t(kMaxGated))
        throw std::runtime_error("scene: more than " + std::to_string(kMaxGated) + " gated (negative-radius) spheres");

    // ---- materials / textures
    const unsigned host_threads = S.n_prims() >= 65536 ? builder_threads() : 1;
    H.f32.mats.resize(S.n_materials());
    if (want_f64)
        H.f64.mats.resize(S.n_materials());
    parallel_for(S.n_materials(), 4096, host_threads, [&](size_t i) {
        const rtb_material &m = S.materials()[i];
        int flags = 0;
        for (int k = 0; k < 4; ++k)
            if (texture_reads_uv(S, m.tex[k]))
                flags |= 1;
        MatT<double> d;
        MatT<float> f;
        d.type = f.type = m.type;
        for (int k = 0; k < 4; ++k)
            d.tex[k] = f.tex[k] = m.tex[k];
        d.flags = f.flags = flags;
        set3(d.color, m.color);
        set3(f.color, m.color);
        d.fuzz = m.fuzz;
        f.fuzz = float(m.fuzz);
        d.ir = m.ir;
        f.ir = float(m.ir);
        // bake solid-colour textures into the record (see mat_tex in rtb_shading.cuh); metal and
        // dielectric keep color / fuzz / ir for their own parameters
        auto solid = [&](int t) { return t >= 0 && S.textures()[t].type == RTB_TEX_SOLID; };
        if (m.type != RTB_MAT_METAL && m.type != RTB_MAT_DIELECTRIC && solid(m.tex[0])) {
            set3(d.color, S.textures()[m.tex[0]].color);
            set3(f.color, S.textures()[m.tex[0]].color);
            d.flags = f.flags = flags |= 2;
        }
        if (m.type == RTB_MAT_PBR && solid(m.tex[1]) && solid(m.tex[2])) {
            d.fuzz = S.textures()[m.tex[1]].color[0];
            d.ir = S.textures()[m.tex[2]].color[0];
            f.fuzz = float(d.fuzz);
            f.ir = float(d.ir);
            d.flags = f.flags = flags |= 4;
        }
        if (want_f64)
            H.f64.mats[i] = d;
        H.f32.mats[i] = f;
    });
    timer.mark("materials");
    // hit-queue key of every primitive (PT_KEY_SHIFT): material type, or 6 for textured lambertians
    parallel_for(H.f32.prims.size(), 8192, host_threads, [&](size_t i) {
        if ((H.f32.prims[i].type_mat & PT_TYPE_MASK) == PT_INSTANCE)
            return;
        const MatT<float> &m = H.f32.mats[H.f32.prims[i].type_mat >> PT_MAT_SHIFT];
        uint32_t key = uint32_t(m.type);
        if (m.type == RTB_MAT_LAMBERTIAN && !(m.flags & 2))
            key = uint32_t(RTB_MAT_TYPE_COUNT);
        H.f32.prims[i].type_mat |= key << PT_KEY_SHIFT;
        if (want_f64)
            H.f64.prims[i].type_mat |= key << PT_KEY_SHIFT;
    });
    H.f32.texs.resize(S.n_textures());
    if (want_f64)
        H.f64.texs.resize(S.n_textures());
    parallel_for(S.n_textures(), 4096, host_threads, [&](size_t i) {
        const rtb_texture &t = S.textures()[i];
        TexT<double> d;
        TexT<float> f;
        d.type = f.type = t.type;
        d.even = f.even = t.even;
        d.odd = f.odd = t.odd;
        d.image = f.image = t.image;
        d.perlin = f.perlin = t.perlin;
        set3(d.color, t.color);
        set3(f.color, t.color);
        d.scale = t.scale;
        f.scale = float(t.scale);
        if (want_f64)
            H.f64.texs[i] = d;
        H.f32.texs[i] = f;
    });
    for (uint64_t i = 0; i < S.n_images(); ++i)
        H.images.push_back(ImageRec{S.images()[i].width, S.images()[i].height, S.images()[i].offset});
    if (S.n_image_bytes())
        H.image_bytes.assign(S.image_bytes(), S.image_bytes() + S.n_image_bytes());
    for (uint64_t i = 0; i < S.n_perlins(); ++i) {
        const rtb_perlin &p = S.perlins()[i];
        PerlinT<double> d;
        PerlinT<float> f;
        for (int k = 0; k < 256; ++k) {
            for (int c = 0; c < 3; ++c) {
                d.ranvec[k][c] = p.ranvec[k][c];
                f.ranvec[k][c] = float(p.ranvec[k][c]);
            }
            d.perm_x[k] = f.perm_x[k] = p.perm_x[k];
            d.perm_y[k] = f.perm_y[k] = p.perm_y[k];
            d.perm_z[k] = f.perm_z[k] = p.perm_z[k];
        }
        H.f64.perlins.push_back(d);
        H.f32.perlins.push_back(f);
    }

    // ---- lights
    H.env_on_device = env_on_device;
    H.n_env_texels = S.n_env_texels();
    H.env_texels_src = S.env_texels();
    if (S.n_env_texels() && !env_on_device) // (a 2048 x 1024 map is 25 MB: no second host copy on the upload path)
        H.env_texels.assign(S.env_texels(), S.env_texels() + S.n_env_texels());
    for (uint64_t i = 0; i < S.n_lights(); ++i) {
        const rtb_light &l = S.lights()[i];
        LightT<double> d;
        std::memset(&d, 0, sizeof(d));
        d.type = l.type;
        d.env_w = l.env_width;
        d.env_h = l.env_height;
        d.env_probe = l.env_is_probe;
        d.env_texel_offset = l.env_offset;
        set3(d.Q, l.Q);
        set3(d.u, l.u);
        set3(d.v, l.v);
        set3(d.intensity, l.intensity);
        d.cos_cutoff = l.cos_cutoff;
        if (l.type == RTB_LIGHT_QUAD) { // quad_light.h:9-16
            const V3<double> n = cross(V3<double>(l.u[0], l.u[1], l.u[2]), V3<double>(l.v[0], l.v[1], l.v[2]));
            d.area = length(n);
            const V3<double> nn = unit_vector(n);
            d.normal[0] = nn.x;
            d.normal[1] = nn.y;
            d.normal[2] = nn.z;
        }
        if (l.type == RTB_LIGHT_ENV) {
            H.n_infinite_lights++;
            if (l.env_width > 0 && l.env_height > 0) {
                d.env_table_offset = H.env_table_doubles;
                H.env_table_doubles += EnvTables::doubles(l.env_width, l.env_height);
                if (!env_on_device)
                    build_env_tables(H.env_texels.data() + l.env_offset, l.env_width, l.env_height, H.env_tables);
            }
        }
        LightT<float> f;
        std::memset(&f, 0, sizeof(f));
        f.type = d.type;
        f.env_w = d.env_w;
        f.env_h = d.env_h;
        f.env_probe = d.env_probe;
        f.env_texel_offset = d.env_texel_offset;
        f.env_table_offset = d.env_table_offset;
        set3(f.Q, d.Q);
        set3(f.u, d.u);
        set3(f.v, d.v);
        set3(f.intensity, d.intensity);
        set3(f.normal, d.normal);
        f.area = float(d.area);
        f.cos_cutoff = float(d.cos_cutoff);
        H.f64.lights.push_back(d);
        H.f32.lights.push_back(f);
    }

    // the production traversal's tree (rtb_trace.cuh)
    timer.mark("textures, lights, env tables");
    std::vector<uint32_t> globals;
    for (uint32_t id : global_ids)
        globals.push_back(uint32_t(H.orig_to_sorted[id]));
    H.wide = build_wide(H.nodes, H.root_ref, H.f32.prims.data(), H.f32.prims.size(), H.chains.size(), globals);
    if (wide_stack_need(H.wide) > kWideStack)
        throw std::runtime_error("scene: BVH of depth " + std::to_string(H.wide.max_depth) +
                                 " exceeds the traversal stack (kWideStack)");

    // the binary traversals (rtb_geom.cuh traverse: validation kernels, round 1's extend) push at most one
    // ref per binary level plus the instance sentinel; a wide level spans at most two binary ones.  A
    // deeper tree would silently lose pushes (LocalStack / SmemStack drop them): refuse it here.
    if (2 * H.wide.max_depth + 2 > kStackDepth)
        throw std::runtime_error("scene: BVH of depth " + std::to_string(2 * H.wide.max_depth) +
                                 " exceeds the binary traversal stack (kStackDepth)");
    timer.mark("4-wide collapse + quantise");
    // (gated spheres are only known to hit_simple(): no typed shared-memory copy for their scenes)
    H.flat_ok = S.n_gates() == 0 && int(H.prim_orig.size()) <= kFlatMaxPrims && int(S.n_xform_ops()) <= kFlatMaxOps &&
                int(S.n_chains()) <= kFlatMaxChains;
    H.f64.camera = derive_camera(cam);
    H.f32.camera = cast_camera<float>(H.f64.camera);
    return H;
}

} // namespace rtb

#endif // RTB_SCENE_HOST_HPP
