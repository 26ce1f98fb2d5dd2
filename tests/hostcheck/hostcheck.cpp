// tests/hostcheck/hostcheck.cpp — DEVELOPMENT AID, NOT A PRODUCT PATH.
//
// The device headers under ray_tracing-rendering_b200/csrc are written as
// `__host__ __device__` templates.  This file instantiates them with g++ so that
// the flattener, the BVH builder, the traversal, the record reconstruction and the
// material / light math can be exercised against the oracle in the GPU-less build
// container (`pytest -m "not gpu"`), before GPU minutes are spent.  It is compiled
// into tests/hostcheck/libhostcheck.so, which only tests load; the product library
// (librtb200.so) contains no host execution path and fails loudly without a GPU.
#include "rtb200_types.h"
#include "rtb_scene_host.hpp"
#include "rtb_trace.cuh"
#include "rtb_wide.cuh"

#include <algorithm>
#include <map>
#include <string>

#include <cstdio>
#include <cstring>

using namespace rtb;

namespace {
template <class R> TypedTables<R> &typed(HostScene &H);
template <> TypedTables<float> &typed<float>(HostScene &H) { return H.f32; }
template <> TypedTables<double> &typed<double>(HostScene &H) { return H.f64; }

template <class R> GeomView<R> geom_view(HostScene &H) {
    TypedTables<R> &T = typed<R>(H);
    GeomView<R> g;
    g.nodes = H.nodes.data();
    g.prims = T.prims.data();
    g.maux = T.maux.data();
    g.ops = T.ops.data();
    g.chains = H.chains.data();
    g.affine = sizeof(R) == 4 ? H.affine.data() : nullptr;
    g.prim_chain = H.prim_chain.data();
    g.prim_orig = H.prim_orig.data();
    g.n_nodes = int(H.nodes.size());
    g.n_prims = int(T.prims.size());
    g.n_ops = int(T.ops.size());
    g.n_chains = int(H.chains.size());
    g.root_ref = H.root_ref;
    g.n_top = H.n_top_items;
    g.flat = H.flat_ok ? 1 : 0;
    g.n_world = H.n_world_slots;
    g.n_gated = int(H.gated.size());
    for (size_t k = 0; k < H.gated.size(); ++k)
        g.gated[k] = H.gated[k];
    return g;
}
template <class R> ShadeView<R> shade_view(HostScene &H) {
    TypedTables<R> &T = typed<R>(H);
    ShadeView<R> s;
    s.mats = T.mats.data();
    s.texs = T.texs.data();
    s.images = H.images.data();
    s.image_bytes = H.image_bytes.data();
    s.perlins = T.perlins.data();
    s.lights = T.lights.data();
    s.env_texels = H.env_texels.data();
    s.env_tables = H.env_tables.data();
    s.n_lights = int(T.lights.size());
    s.n_infinite_lights = H.n_infinite_lights;
    return s;
}

// the 4-wide tree of a scene (rtb_wide.cuh), built by build_host_scene
const WideTree &wide_of(HostScene &H) { return H.wide; }

template <class R, bool ROBUST>
void trace_batch(HostScene &H, const rtb_ray *rays, uint64_t n, rtb_hit *hits, uint64_t stats[2], bool use_flat,
                 bool inst_in_descent = false, bool plane_records = false, bool wide = false, bool reference_walk = false) {
    const GeomView<R> g = geom_view<R>(H);
    const WideTree *W = wide ? &wide_of(H) : nullptr;
    RngT<R> rng;
    rng.g = pcg_seed(1, 2);
    auto draw = [&]() { return rng.next_open(); };
    for (uint64_t i = 0; i < n; ++i) {
        const rtb_ray &q = rays[i];
        const V3<R> o{R(q.o[0]), R(q.o[1]), R(q.o[2])}, d{R(q.d[0]), R(q.d[1]), R(q.d[2])};
        uint32_t origin = kNoPrim;
        if (ROBUST && q.origin_prim >= 0 && q.origin_prim < int(H.orig_to_sorted.size()))
            origin = uint32_t(H.orig_to_sorted[q.origin_prim]);
        R t;
        uint32_t storage[kStackDepth];
        LocalStack stack(storage);
        auto run = [&](const GeomView<R> &gv, R t_max, R &t_out) -> uint32_t {
            stack.clear();
            return wide ? traverse_wide<R, false, ROBUST>(gv, W->nodes.data(), W->root_ref, W->chain_root.data(), o, d, R(q.time),
                                                          R(q.t_min), t_max, origin, draw, t_out, &stats[0], &stats[1], stack,
                                                          W->global_prims.data(), uint32_t(W->global_prims.size()))
                   : use_flat && gv.flat
                       ? traverse_flat<R, false, ROBUST>(gv, o, d, R(q.time), R(q.t_min), t_max, origin, draw, t_out,
                                                         &stats[0], &stats[1])
                   : inst_in_descent
                       ? traverse<R, false, ROBUST, decltype(draw), LocalStack, true, false>(
                             gv, o, d, R(q.time), R(q.t_min), t_max, origin, draw, t_out, &stats[0], &stats[1], stack)
                       : traverse<R, false, ROBUST>(gv, o, d, R(q.time), R(q.t_min), t_max, origin, draw, t_out,
                                                    &stats[0], &stats[1], stack);
        };
        // fp64: gated spheres exactly as the reference reaches them; fp32: inside the traversal, as the renderer does
        uint32_t pi;
        if (reference_walk) { // the library's precision 65: the reference's walk with its random stream
            XorShift32Draw xs{uint32_t(q.reserved)};
            pi = walk_reference_order<R>(g, H.orig_to_sorted.data(), int(H.orig_to_sorted.size()), o, d, R(q.time), R(q.t_min),
                                         R(q.t_max), xs, t);
        } else {
            pi = (!ROBUST && g.n_gated) ? trace_gated_exact<R>(g, o, d, R(q.t_min), R(q.t_max), t, run) : run(g, R(q.t_max), t);
        }
        rtb_hit &h = hits[i];
        std::memset(&h, 0, sizeof(h));
        h.prim = -1;
        h.material = -1;
        if (pi == kNoPrim)
            continue;
        RecT<R> rec;
        PlaneRec pl;
        pl.valid = 0;
        if (plane_records) // the fused kernel's digest of planar primitives (u,v are not produced)
            build_plane_rec(g, pi, pl);
        if (pl.valid)
            rec = plane_record<R>(pl, o, d, t);
        else
            rec = make_record<R, ROBUST, true>(g, pi, o, d, R(q.time), t);
        h.t = rec.t;
        h.p[0] = rec.p.x; h.p[1] = rec.p.y; h.p[2] = rec.p.z;
        h.normal[0] = rec.normal.x; h.normal[1] = rec.normal.y; h.normal[2] = rec.normal.z;
        h.u = rec.u;
        h.v = rec.v;
        h.prim = g.prim_orig[pi];
        h.front_face = rec.front_face;
        h.material = int(g.prims[pi].type_mat >> PT_MAT_SHIFT);
    }
}

// The production scheduler of rtb_trace.cuh (votes, window sort, refill) on emulated warps
// (rtb_warp.cuh): `warps` warps of 32 fibers pull windows off one cursor, as the persistent
// kernel's warps do.  Fills t / primitive only (records are tested through the scalar paths).
static uint32_t g_sim_warps = 24u;
void trace_batch_warp(HostScene &H, const rtb_ray *rays, uint64_t n, rtb_hit *hits, uint64_t stats[2], bool any_hit,
                      int warps) {
    const GeomView<float> g = geom_view<float>(H);
    const WideTree &W = wide_of(H);
    WideView wv;
    wv.nodes = reinterpret_cast<const Vec4f *>(W.qnodes.data());
    wv.chain_root = W.chain_root.data();
    wv.root_ref = W.root_ref;
    wv.n_nodes = uint32_t(W.qnodes.size());
    wv.n_global = uint32_t(W.global_prims.size());
    for (int i = 0; i < kMaxGlobalPrims; ++i)
        wv.global_prim[i] = size_t(i) < W.global_prims.size() ? W.global_prims[size_t(i)] : 0u;
    std::vector<Vec4f> a(n), b(n);
    std::vector<Vec2f> tt(n), out(n);
    for (uint64_t i = 0; i < n; ++i) {
        const rtb_ray &q = rays[i];
        uint32_t origin = kNoPrim;
        if (q.origin_prim >= 0 && q.origin_prim < int(H.orig_to_sorted.size()))
            origin = uint32_t(H.orig_to_sorted[q.origin_prim]);
        a[i] = Vec4f{float(q.o[0]), float(q.o[1]), float(q.o[2]), float(q.time)};
        b[i] = Vec4f{float(q.d[0]), float(q.d[1]), float(q.d[2]), u2f(origin)};
        tt[i] = Vec2f{float(q.t_min), float(q.t_max)};
        out[i] = Vec2f{-1.f, u2f(0x12345678u)};
    }
    uint32_t head = 0, overflow = 0;
    // the top of the tree from a separate copy, as the kernels read it from shared memory
    const uint32_t n_top = std::min<uint32_t>(wv.n_nodes, kTopNodesMax);
    std::vector<Vec4f> top(size_t(n_top) * kTopStride);
    for (uint32_t i = 0; i < n_top * kNodeRows; ++i)
        top[(i / kNodeRows) * kTopStride + (i % kNodeRows)] = wv.nodes[i];
    const bool media = H.has_media, inst = H.n_instances > 0;
    for (int wi = 0; wi < warps; ++wi) {
        TraceWarpSmem smem;
        HostWarp hw;
        head = uint32_t(n * uint64_t(wi) / uint64_t(warps));
        hw.run([&]() {
            BatchTraceJob job;
            job.ray_a = a.data();
            job.ray_b = b.data();
            job.ray_t = tt.data();
            job.out = out.data();
            job.n = uint32_t(n * uint64_t(wi + 1) / uint64_t(warps)); // this warp's share of the batch
            job.head = &head;
            job.base = 0;
            job.win = 0;
            job.warps = g_sim_warps; // (24: windows shrink quickly, the tail logic is exercised; 1: full-size windows, as in a long launch)
            job.seed = 0x51ed270b;
            uint64_t c[3] = {0, 0, 0};
            uint32_t ov = 0;
            if (any_hit)
                warp_trace<BatchTraceJob, true, true, true, true>(g, wv, top.data(), n_top, smem, job, c, ov);
            else if (media || inst)
                warp_trace<BatchTraceJob, false, true, true, true>(g, wv, top.data(), n_top, smem, job, c, ov);
            else
                warp_trace<BatchTraceJob, false, false, false, false>(g, wv, nullptr, 0, smem, job, c, ov);
            stats[0] += c[0];
            stats[1] += c[1];
            overflow |= ov;
        });
    }
    if (overflow)
        throw std::runtime_error("warp_trace: overflow flag " + std::to_string(overflow));
    for (uint64_t i = 0; i < n; ++i) {
        rtb_hit &h = hits[i];
        std::memset(&h, 0, sizeof(h));
        const uint32_t pi = f2u(out[i].y);
        h.prim = pi == kNoPrim ? -1 : g.prim_orig[pi];
        h.material = pi == kNoPrim ? -1 : int(g.prims[pi].type_mat >> PT_MAT_SHIFT);
        h.t = pi == kNoPrim ? 0.0 : double(out[i].x);
    }
}

template <class R> RecT<R> rec_of(const rtb_bsdf_query &q) {
    RecT<R> r;
    r.p = V3<R>(R(q.p[0]), R(q.p[1]), R(q.p[2]));
    r.normal = V3<R>(R(q.normal[0]), R(q.normal[1]), R(q.normal[2]));
    r.u = R(q.u);
    r.v = R(q.v);
    r.t = 1;
    r.front_face = q.front_face != 0;
    return r;
}

template <class R>
void bsdf_eval(HostScene &H, int mat, const rtb_bsdf_query *q, uint64_t n, rtb_bsdf_value *out) {
    const ShadeView<R> S = shade_view<R>(H);
    const MatT<R> m = S.mats[mat];
    for (uint64_t i = 0; i < n; ++i) {
        const RecT<R> rec = rec_of<R>(q[i]);
        const V3<R> wo(R(q[i].wo[0]), R(q[i].wo[1]), R(q[i].wo[2]));
        const V3<R> wi(R(q[i].wi[0]), R(q[i].wi[1]), R(q[i].wi[2]));
        const V3<R> f = mat_eval(S, m, rec, wo, wi);
        const V3<R> e0 = mat_emitted_old(S, m, rec), e1 = mat_emitted_new(S, m, rec);
        out[i].f[0] = f.x; out[i].f[1] = f.y; out[i].f[2] = f.z;
        out[i].pdf = mat_pdf(S, m, rec, wo, wi);
        out[i].emitted_old[0] = e0.x; out[i].emitted_old[1] = e0.y; out[i].emitted_old[2] = e0.z;
        out[i].emitted_new[0] = e1.x; out[i].emitted_new[1] = e1.y; out[i].emitted_new[2] = e1.z;
    }
}

template <class R>
void bsdf_sample(HostScene &H, int mat, const rtb_bsdf_query *q, uint64_t n, uint64_t seed,
                 rtb_bsdf_sample *out) {
    const ShadeView<R> S = shade_view<R>(H);
    const MatT<R> m = S.mats[mat];
    for (uint64_t i = 0; i < n; ++i) {
        RngT<R> g;
        g.g = pcg_seed(i, seed);
        std::memset(&out[i], 0, sizeof(out[i]));
        const RecT<R> rec = rec_of<R>(q[i]);
        const V3<R> wo(R(q[i].wo[0]), R(q[i].wo[1]), R(q[i].wo[2]));
        BsdfSampleT<R> bs;
        bs.pdf = 0;
        bs.is_specular = false;
        const bool ok = mat_sample(S, m, rec, wo, g, bs);
        out[i].ok = ok;
        if (ok) {
            out[i].wi[0] = bs.wi.x; out[i].wi[1] = bs.wi.y; out[i].wi[2] = bs.wi.z;
            out[i].f[0] = bs.f.x; out[i].f[1] = bs.f.y; out[i].f[2] = bs.f.z;
            out[i].pdf = bs.pdf;
            out[i].is_specular = bs.is_specular;
        }
        V3<R> atten, dout;
        const bool sok = mat_scatter(S, m, rec, -wo, g, atten, dout);
        out[i].scatter_ok = sok;
        if (sok) {
            out[i].scatter_dir[0] = dout.x; out[i].scatter_dir[1] = dout.y; out[i].scatter_dir[2] = dout.z;
            out[i].scatter_atten[0] = atten.x; out[i].scatter_atten[1] = atten.y; out[i].scatter_atten[2] = atten.z;
        }
    }
}

template <class R>
void light_eval(HostScene &H, int light, const rtb_light_query *q, uint64_t n, uint64_t seed,
                rtb_light_value *out) {
    const ShadeView<R> S = shade_view<R>(H);
    const LightT<R> l = S.lights[light];
    for (uint64_t i = 0; i < n; ++i) {
        RngT<R> g;
        g.g = pcg_seed(i, seed);
        std::memset(&out[i], 0, sizeof(out[i]));
        const V3<R> p(R(q[i].p[0]), R(q[i].p[1]), R(q[i].p[2]));
        const V3<R> d(R(q[i].d[0]), R(q[i].d[1]), R(q[i].d[2]));
        const LightSampleT<R> ls = light_sample(S, l, p, R(q[i].u[0]), R(q[i].u[1]), g);
        out[i].Li[0] = ls.Li.x; out[i].Li[1] = ls.Li.y; out[i].Li[2] = ls.Li.z;
        out[i].wi[0] = ls.wi.x; out[i].wi[1] = ls.wi.y; out[i].wi[2] = ls.wi.z;
        out[i].pdf = ls.pdf;
        out[i].dist = ls.dist;
        out[i].is_delta = ls.is_delta;
        out[i].pdf_dir = light_pdf(S, l, p, d);
        const V3<R> le = light_Le(S, l, d);
        out[i].Le[0] = le.x; out[i].Le[1] = le.y; out[i].Le[2] = le.z;
    }
}
} // namespace

extern "C" {

static bool g_group_boxes = true;
void hc_group_boxes(int on) { g_group_boxes = on != 0; } // RTB_OPT_GROUP_BOXES of the scenes created from now on
int hc_scene_boxes(void *h) { return static_cast<HostScene *>(h)->n_boxes; }
void *hc_scene_create(const void *blob, uint64_t nbytes, int max_leaf) {
    try {
        SceneView S(blob, nbytes);
        return new HostScene(build_host_scene(S, max_leaf, 1.0, false, true, false, g_group_boxes));
    } catch (const std::exception &e) {
        std::fprintf(stderr, "hc_scene_create: %s\n", e.what());
        return nullptr;
    }
}
void hc_scene_destroy(void *h) {
    delete static_cast<HostScene *>(h);
}

// nodes and occupied child slots of the scene's 4-wide tree
void hc_wide_info(void *h, uint64_t out[2]) {
    const WideTree &w = wide_of(*static_cast<HostScene *>(h));
    out[0] = w.nodes.size();
    out[1] = w.n_children;
}

// votes of the emulated scheduler since the last call: {node steps, lanes in them, leaf steps, lanes,
// refills, lanes refilled}; resets the counters
void hc_sched_stats(uint64_t out[6]) {
    TraceSchedStats &s = trace_sched_stats();
    out[0] = s.node_steps;
    out[1] = s.node_lanes;
    out[2] = s.leaf_steps;
    out[3] = s.leaf_lanes;
    out[4] = s.switches;
    out[5] = s.switch_lanes;
    s = TraceSchedStats();
}

// Quantised nodes (QNode64) against the fp32 nodes they were made from: out = {children whose
// quantised box does NOT contain the original one (must be 0), children, sum over children and axes
// of the relative growth of the box extent in 1e-6 units}
void hc_qnode_check(void *h, uint64_t out[3]) {
    const WideTree &w = wide_of(*static_cast<HostScene *>(h));
    out[0] = out[1] = out[2] = 0;
    for (size_t n = 0; n < w.nodes.size(); ++n) {
        const Node128 &a = w.nodes[n];
        const QNode64 &q = w.qnodes[n];
        const double o[3] = {q.ox, q.oy, q.oz}, s[3] = {q.sx, q.sy, q.sz};
        const uint32_t lo[3] = {q.lox, q.loy, q.loz}, hi[3] = {q.hix, q.hiy, q.hiz};
        for (int i = 0; i < 4; ++i) {
            if (a.ref[i] == kEmptyRef)
                continue;
            ++out[1];
            bool ok = q.ref[i] == a.ref[i];
            for (int k = 0; k < 3; ++k) {
                const double ql = o[k] + double((lo[k] >> (8 * i)) & 255u) * s[k], qh = o[k] + double((hi[k] >> (8 * i)) & 255u) * s[k];
                ok = ok && ql <= double(a.lo[k][i]) && qh >= double(a.hi[k][i]);
                const double ext = double(a.hi[k][i]) - double(a.lo[k][i]);
                if (ext > 0)
                    out[2] += uint64_t(1e6 * std::min(10.0, ((qh - ql) - ext) / ext));
            }
            if (!ok)
                ++out[0];
        }
    }
}

void hc_sched_tuning(int node_min, int switch_min) {
    trace_tuning().node_min = uint32_t(node_min);
    trace_tuning().switch_min = uint32_t(switch_min);
}
void hc_sched_leaf_min(int leaf_min) { trace_tuning().leaf_min = uint32_t(leaf_min); }
void hc_sched_sim_warps(int w) { g_sim_warps = uint32_t(w); }
void hc_sched_postpone(int on) { trace_tuning().postpone = on != 0; }

// sizes = {nodes, sorted prims, instances}
void hc_scene_info(void *h, int64_t sizes[3]) {
    auto *H = static_cast<HostScene *>(h);
    sizes[0] = int64_t(H->nodes.size());
    sizes[1] = int64_t(H->prim_orig.size());
    sizes[2] = H->n_instances;
}

// FNV-1a over the node array and the leaf order: equal iff the trees are identical
uint64_t hc_scene_tree_hash(void *h) {
    auto *H = static_cast<HostScene *>(h);
    uint64_t x = 1469598103934665603ull;
    auto mix = [&](const void *p, size_t n) {
        const unsigned char *b = static_cast<const unsigned char *>(p);
        for (size_t i = 0; i < n; ++i)
            x = (x ^ b[i]) * 1099511628211ull;
    };
    mix(H->nodes.data(), H->nodes.size() * sizeof(Node32));
    mix(H->prim_orig.data(), H->prim_orig.size() * sizeof(H->prim_orig[0]));
    return x;
}

// precision 64: the validation arithmetic (reference operation order, no self-hit
// logic); 32: the production arithmetic.  stats = {nodes visited, primitive tests}.
void hc_trace_batch(void *h, const rtb_ray *rays, uint64_t n, int precision, rtb_hit *hits,
                    uint64_t stats[2]) {
    auto *H = static_cast<HostScene *>(h);
    uint64_t local[2] = {0, 0};
    // 64: fp64 through the BVH; 65: fp64 lockstep; 66: fp64 through the BVH with instance entry /
    // exit inside the descent (the traversal shape of instance-free scenes, valid on any scene);
    // 32: fp32 as the renderer traces this scene (lockstep when it is small); 33: fp32 forced
    // through the BVH; 34: the same with instance entry / exit inside the descent; 35: as 32 with
    // the records of planar primitives from plane_record() (the fused kernel's shading input)
    // 67 / 37: fp64 / fp32 through the 4-wide tree of rtb_wide.cuh (stats[0] counts 128-byte nodes)
    // 38 / 39: fp32 through the production warp scheduler of rtb_trace.cuh on emulated warps
    // (closest hit / any hit)
    if (precision == 38 || precision == 39) {
        try {
            trace_batch_warp(*H, rays, n, hits, local, precision == 39, 3);
        } catch (const std::exception &e) {
            std::fprintf(stderr, "hc_trace_batch(%d): %s\n", precision, e.what());
            for (uint64_t i = 0; i < n; ++i)
                hits[i].prim = -2;
        }
    } else if (precision == 67)
        trace_batch<double, false>(*H, rays, n, hits, local, false, false, false, true);
    else if (precision == 37)
        trace_batch<float, true>(*H, rays, n, hits, local, false, false, false, true);
    else if (precision == 69)
        trace_batch<double, false>(*H, rays, n, hits, local, false, false, false, false, true);
    else if (precision >= 64)
        trace_batch<double, false>(*H, rays, n, hits, local, precision == 65, precision == 66);
    else
        trace_batch<float, true>(*H, rays, n, hits, local, precision == 32 || precision == 35, precision == 34,
                                 precision == 35);
    if (stats) {
        stats[0] = local[0];
        stats[1] = local[1];
    }
}

void hc_bsdf_eval(void *h, int mat, const rtb_bsdf_query *q, uint64_t n, int precision,
                  rtb_bsdf_value *out) {
    auto *H = static_cast<HostScene *>(h);
    if (precision == 64)
        bsdf_eval<double>(*H, mat, q, n, out);
    else
        bsdf_eval<float>(*H, mat, q, n, out);
}
void hc_bsdf_sample(void *h, int mat, const rtb_bsdf_query *q, uint64_t n, int precision, uint64_t seed,
                    rtb_bsdf_sample *out) {
    auto *H = static_cast<HostScene *>(h);
    if (precision == 64)
        bsdf_sample<double>(*H, mat, q, n, seed, out);
    else
        bsdf_sample<float>(*H, mat, q, n, seed, out);
}
void hc_light_eval(void *h, int light, const rtb_light_query *q, uint64_t n, int precision, uint64_t seed,
                   rtb_light_value *out) {
    auto *H = static_cast<HostScene *>(h);
    if (precision == 64)
        light_eval<double>(*H, light, q, n, seed, out);
    else
        light_eval<float>(*H, light, q, n, seed, out);
}
void hc_camera_derived(void *h, double out[24]) {
    auto *H = static_cast<HostScene *>(h);
    const CameraT<double> &c = H->f64.camera;
    const V3<double> *vs[7] = {&c.origin, &c.lower_left_corner, &c.horizontal, &c.vertical, &c.u, &c.v, &c.w};
    for (int i = 0; i < 7; ++i) {
        out[3 * i] = vs[i]->x;
        out[3 * i + 1] = vs[i]->y;
        out[3 * i + 2] = vs[i]->z;
    }
    out[21] = c.lens_radius;
    out[22] = c.time0;
    out[23] = c.time1;
}
// The renderer's per-sample random streams (rtb_shading.cuh): stream i = pcg_seed(first + i, seed); its first
// `draws` float draws (next_f: the top 24 bits of the LCG state) go to out[i * draws ...].
// the fp32 acos_ / atan2_ of csrc/rtb_math.cuh (polynomial forms) on n arguments
void hc_invtrig(const float *x, const float *y, uint64_t n, float *acos_out, float *atan2_out) {
    for (uint64_t i = 0; i < n; ++i) {
        acos_out[i] = acos_(x[i]);
        atan2_out[i] = atan2_(y[i], x[i]);
    }
}

void hc_rng_draws(uint64_t first, uint64_t n_streams, uint64_t draws, uint64_t seed, float *out) {
    for (uint64_t i = 0; i < n_streams; ++i) {
        Pcg g = pcg_seed(first + i, seed);
        for (uint64_t k = 0; k < draws; ++k)
            out[i * draws + k] = g.next_f();
    }
}
// the env lights' Distribution2D tables as the HOST builds them (detail::build_env_tables); returns their size
uint64_t hc_env_tables(void *h, double *out, uint64_t capacity) {
    auto *H = static_cast<HostScene *>(h);
    const uint64_t n = H->env_tables.size();
    for (uint64_t i = 0; i < n && i < capacity; ++i)
        out[i] = H->env_tables[i];
    return n;
}
}
