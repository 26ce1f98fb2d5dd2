// rtb_batch.cu — the parity-layer batch kernels behind rtb_trace_batch,
// rtb_bsdf_*_batch, rtb_light_eval_batch and rtb_texture_eval_batch.
//
// Compiled TWICE from this one source (see csrc/Makefile):
//   -DRTB_REAL=double -fmad=false  -> rtb_batch_f64.o : fp64 validation kernels.
//        No FMA contraction, reference operation order => t and primitive ids are
//        bit-exact against the reference's hit() (x86-64 baseline, no FMA either).
//   -DRTB_REAL=float               -> rtb_batch_f32.o : the SAME traversal / shading
//        device functions the wavefront kernels instantiate, with the production
//        self-intersection handling (ROBUST = true).
//
// One thread per query; queries are independent, inputs are read with coalesced
// struct loads, grids are sized in multiples of the SM count.
#include "rtb_internal.hpp"

#ifndef RTB_REAL
#error "compile with -DRTB_REAL=float or -DRTB_REAL=double"
#endif

namespace rtb {

typedef RTB_REAL Real;
constexpr bool kRobust = sizeof(Real) == 4;

namespace {

struct DrawOpen {
    RngT<Real> *g;
    __device__ Real operator()() { return g->next_open(); }
};

// the plain while-while traversal (the renderer's speculative one is warp-cooperative and is
// checked against this one ray by ray in tests/test_hostcheck_parity.py and, through whole
// renders, against the fused schedule and the reference images)
// INST: the traversal shape the renderer uses for this scene (instances handled in the leaf phase
// when the scene has any, inside the descent otherwise; see traverse() in rtb_geom.cuh)
template <bool INST, class Rng>
__device__ __forceinline__ uint32_t trace_bvh(const GeomView<Real> &g, V3<Real> o, V3<Real> d, Real time, Real t_min,
                                              Real t_max, uint32_t origin, Rng &rng, Real &t, uint64_t *nodes,
                                              uint64_t *tests) {
    uint32_t storage[kStackDepth];
    LocalStack stack(storage);
    return traverse<Real, false, kRobust, Rng, LocalStack, true, INST>(g, o, d, time, t_min, t_max, origin, rng, t, nodes,
                                                                        tests, stack);
}

template <bool INST>
__global__ void __launch_bounds__(128)
k_trace_batch(GeomView<Real> g, const int32_t *__restrict__ orig_to_sorted, int n_orig,
              const rtb_ray *__restrict__ rays, uint64_t n, rtb_hit *__restrict__ hits,
              unsigned long long *visits, int reference_walk) {
    uint64_t nodes = 0, tests = 0;
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n;
         i += uint64_t(gridDim.x) * blockDim.x) {
        const rtb_ray q = rays[i];
        const V3<Real> o(Real(q.o[0]), Real(q.o[1]), Real(q.o[2]));
        const V3<Real> d(Real(q.d[0]), Real(q.d[1]), Real(q.d[2]));
        uint32_t origin = kNoPrim;
        if (kRobust && q.origin_prim >= 0 && q.origin_prim < n_orig)
            origin = uint32_t(orig_to_sorted[q.origin_prim]);
        RngT<Real> rng;
        rng.g = pcg_seed(i, 0x51ed270b);
        DrawOpen draw{&rng};
        Real t;
        // the fp32 instantiation traces the way the renderer does: lockstep for small scenes
        auto run = [&](const GeomView<Real> &gv, Real t_max, Real &t_out) -> uint32_t {
            return (kRobust && gv.flat)
                       ? traverse_flat<Real, false, kRobust>(gv, o, d, Real(q.time), Real(q.t_min), t_max, origin, draw, t_out,
                                                             visits ? &nodes : nullptr, visits ? &tests : nullptr)
                       : trace_bvh<INST>(gv, o, d, Real(q.time), Real(q.t_min), t_max, origin, draw, t_out,
                                         visits ? &nodes : nullptr, visits ? &tests : nullptr);
        };
        // fp64: gated (negative-radius) spheres exactly as the reference reaches them (rtb_geom.cuh)
        uint32_t pi;
        if (!kRobust && reference_walk) {
            // precision 65: the reference's left-to-right walk with ITS random stream (state in rtb_ray.reserved)
            XorShift32Draw xs{uint32_t(q.reserved)};
            pi = walk_reference_order<Real>(g, orig_to_sorted, n_orig, o, d, Real(q.time), Real(q.t_min), Real(q.t_max), xs, t);
        } else {
            pi = (!kRobust && g.n_gated) ? trace_gated_exact<Real>(g, o, d, Real(q.t_min), Real(q.t_max), t, run)
                                         : run(g, Real(q.t_max), t);
        }
        rtb_hit h;
        h.t = 0;
        h.p[0] = h.p[1] = h.p[2] = 0;
        h.normal[0] = h.normal[1] = h.normal[2] = 0;
        h.u = h.v = 0;
        h.prim = -1;
        h.front_face = 0;
        h.material = -1;
        h.reserved = 0;
        if (pi != kNoPrim) {
            const RecT<Real> rec = make_record<Real, kRobust, true>(g, pi, o, d, Real(q.time), t);
            h.t = rec.t;
            h.p[0] = rec.p.x;
            h.p[1] = rec.p.y;
            h.p[2] = rec.p.z;
            h.normal[0] = rec.normal.x;
            h.normal[1] = rec.normal.y;
            h.normal[2] = rec.normal.z;
            h.u = rec.u;
            h.v = rec.v;
            h.prim = g.prim_orig[pi];
            h.front_face = rec.front_face ? 1 : 0;
            h.material = int32_t(g.prims[pi].type_mat >> PT_MAT_SHIFT);
        }
        hits[i] = h;
    }
    if (visits) {
        atomicAdd(&visits[0], (unsigned long long)nodes);
        atomicAdd(&visits[1], (unsigned long long)tests);
    }
}

__device__ RecT<Real> rec_of(const rtb_bsdf_query &q) {
    RecT<Real> r;
    r.p = V3<Real>(Real(q.p[0]), Real(q.p[1]), Real(q.p[2]));
    r.normal = V3<Real>(Real(q.normal[0]), Real(q.normal[1]), Real(q.normal[2]));
    r.u = Real(q.u);
    r.v = Real(q.v);
    r.t = 1;
    r.front_face = q.front_face != 0;
    return r;
}

__global__ void __launch_bounds__(128)
k_bsdf_eval(ShadeView<Real> S, int material, const rtb_bsdf_query *__restrict__ qs, uint64_t n,
            rtb_bsdf_value *__restrict__ out) {
    const MatT<Real> m = S.mats[material];
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n;
         i += uint64_t(gridDim.x) * blockDim.x) {
        const rtb_bsdf_query q = qs[i];
        const RecT<Real> rec = rec_of(q);
        const V3<Real> wo(Real(q.wo[0]), Real(q.wo[1]), Real(q.wo[2]));
        const V3<Real> wi(Real(q.wi[0]), Real(q.wi[1]), Real(q.wi[2]));
        const V3<Real> f = mat_eval(S, m, rec, wo, wi);
        const V3<Real> e0 = mat_emitted_old(S, m, rec), e1 = mat_emitted_new(S, m, rec);
        rtb_bsdf_value v;
        v.f[0] = f.x; v.f[1] = f.y; v.f[2] = f.z;
        v.pdf = mat_pdf(S, m, rec, wo, wi);
        v.emitted_old[0] = e0.x; v.emitted_old[1] = e0.y; v.emitted_old[2] = e0.z;
        v.emitted_new[0] = e1.x; v.emitted_new[1] = e1.y; v.emitted_new[2] = e1.z;
        out[i] = v;
    }
}

__global__ void __launch_bounds__(128)
k_bsdf_sample(ShadeView<Real> S, int material, const rtb_bsdf_query *__restrict__ qs, uint64_t n,
              uint64_t seed, rtb_bsdf_sample *__restrict__ out) {
    const MatT<Real> m = S.mats[material];
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n;
         i += uint64_t(gridDim.x) * blockDim.x) {
        const rtb_bsdf_query q = qs[i];
        RngT<Real> g;
        g.g = pcg_seed(i, seed);
        const RecT<Real> rec = rec_of(q);
        const V3<Real> wo(Real(q.wo[0]), Real(q.wo[1]), Real(q.wo[2]));
        rtb_bsdf_sample s;
        memset(&s, 0, sizeof(s));
        BsdfSampleT<Real> bs;
        bs.pdf = 0;
        bs.is_specular = false;
        bs.wi = V3<Real>(0, 0, 0);
        bs.f = V3<Real>(0, 0, 0);
        const bool ok = mat_sample(S, m, rec, wo, g, bs);
        s.ok = ok;
        if (ok) {
            s.wi[0] = bs.wi.x; s.wi[1] = bs.wi.y; s.wi[2] = bs.wi.z;
            s.f[0] = bs.f.x; s.f[1] = bs.f.y; s.f[2] = bs.f.z;
            s.pdf = bs.pdf;
            s.is_specular = bs.is_specular;
        }
        V3<Real> atten(0, 0, 0), dout(0, 0, 0);
        const bool sok = mat_scatter(S, m, rec, -wo, g, atten, dout);
        s.scatter_ok = sok;
        if (sok) {
            s.scatter_dir[0] = dout.x; s.scatter_dir[1] = dout.y; s.scatter_dir[2] = dout.z;
            s.scatter_atten[0] = atten.x; s.scatter_atten[1] = atten.y; s.scatter_atten[2] = atten.z;
        }
        out[i] = s;
    }
}

__global__ void __launch_bounds__(128)
k_light_eval(ShadeView<Real> S, int light, const rtb_light_query *__restrict__ qs, uint64_t n, uint64_t seed,
             rtb_light_value *__restrict__ out) {
    const LightT<Real> l = S.lights[light];
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n;
         i += uint64_t(gridDim.x) * blockDim.x) {
        const rtb_light_query q = qs[i];
        RngT<Real> g;
        g.g = pcg_seed(i, seed);
        const V3<Real> p(Real(q.p[0]), Real(q.p[1]), Real(q.p[2]));
        const V3<Real> d(Real(q.d[0]), Real(q.d[1]), Real(q.d[2]));
        const LightSampleT<Real> ls = light_sample(S, l, p, Real(q.u[0]), Real(q.u[1]), g);
        rtb_light_value v;
        memset(&v, 0, sizeof(v));
        v.Li[0] = ls.Li.x; v.Li[1] = ls.Li.y; v.Li[2] = ls.Li.z;
        v.wi[0] = ls.wi.x; v.wi[1] = ls.wi.y; v.wi[2] = ls.wi.z;
        v.pdf = ls.pdf;
        v.dist = ls.dist;
        v.is_delta = ls.is_delta ? 1 : 0;
        v.pdf_dir = light_pdf(S, l, p, d);
        const V3<Real> le = light_Le(S, l, d);
        v.Le[0] = le.x; v.Le[1] = le.y; v.Le[2] = le.z;
        out[i] = v;
    }
}

__global__ void __launch_bounds__(128)
k_texture_eval(ShadeView<Real> S, int texture, const double *__restrict__ uvp, uint64_t n,
               double *__restrict__ rgb) {
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n;
         i += uint64_t(gridDim.x) * blockDim.x) {
        const double *q = uvp + 5 * i;
        const V3<Real> c = tex_value(S, texture, Real(q[0]), Real(q[1]), V3<Real>(Real(q[2]), Real(q[3]), Real(q[4])));
        rgb[3 * i] = c.x;
        rgb[3 * i + 1] = c.y;
        rgb[3 * i + 2] = c.z;
    }
}

int grid_for(const rtb_context *ctx, uint64_t n, int block) {
    const uint64_t want = (n + block - 1) / block;
    const uint64_t cap = uint64_t(ctx->sm_count > 0 ? ctx->sm_count : 148) * 8;
    uint64_t g = want < cap ? want : cap;
    // whole multiples of the SM count once the batch is large enough to fill the chip
    if (ctx->sm_count > 0 && g > uint64_t(ctx->sm_count))
        g = (g / ctx->sm_count) * ctx->sm_count;
    return int(g ? g : 1);
}

} // namespace

template <>
void launch_trace_batch<Real>(rtb_context *ctx, const rtb_ray *d_rays, uint64_t n, rtb_hit *d_hits,
                              unsigned long long *d_visits, bool reference_walk) {
    if (!n)
        return;
    const DeviceScene &sc = *ctx->scene;
    GeomView<Real> gv = sc.geom<Real>();
    if (ctx->opt_flat == 0)
        gv.flat = 0;
    // fp32: the renderer's choice of traversal shape; fp64 validation: always the leaf-phase shape
    if (kRobust && sc.host.n_instances == 0)
        k_trace_batch<false><<<grid_for(ctx, n, 128), 128, 0, ctx->stream>>>(
            gv, sc.orig_to_sorted.as<int32_t>(), int(sc.host.orig_to_sorted.size()), d_rays, n, d_hits, d_visits, 0);
    else
        k_trace_batch<true><<<grid_for(ctx, n, 128), 128, 0, ctx->stream>>>(
            gv, sc.orig_to_sorted.as<int32_t>(), int(sc.host.orig_to_sorted.size()), d_rays, n, d_hits, d_visits,
            reference_walk ? 1 : 0);
    RTB_CUDA(cudaGetLastError());
}
template <>
void launch_bsdf_eval<Real>(rtb_context *ctx, int material, const rtb_bsdf_query *d_q, uint64_t n,
                            rtb_bsdf_value *d_out) {
    if (!n)
        return;
    k_bsdf_eval<<<grid_for(ctx, n, 128), 128, 0, ctx->stream>>>(ctx->scene->shade<Real>(), material, d_q, n,
                                                               d_out);
    RTB_CUDA(cudaGetLastError());
}
template <>
void launch_bsdf_sample<Real>(rtb_context *ctx, int material, const rtb_bsdf_query *d_q, uint64_t n,
                              uint64_t seed, rtb_bsdf_sample *d_out) {
    if (!n)
        return;
    k_bsdf_sample<<<grid_for(ctx, n, 128), 128, 0, ctx->stream>>>(ctx->scene->shade<Real>(), material, d_q, n,
                                                                 seed, d_out);
    RTB_CUDA(cudaGetLastError());
}
template <>
void launch_light_eval<Real>(rtb_context *ctx, int light, const rtb_light_query *d_q, uint64_t n, uint64_t seed,
                             rtb_light_value *d_out) {
    if (!n)
        return;
    k_light_eval<<<grid_for(ctx, n, 128), 128, 0, ctx->stream>>>(ctx->scene->shade<Real>(), light, d_q, n, seed,
                                                                d_out);
    RTB_CUDA(cudaGetLastError());
}
template <>
void launch_texture_eval<Real>(rtb_context *ctx, int texture, const double *d_uvp, uint64_t n, double *d_rgb) {
    if (!n)
        return;
    k_texture_eval<<<grid_for(ctx, n, 128), 128, 0, ctx->stream>>>(ctx->scene->shade<Real>(), texture, d_uvp, n,
                                                                  d_rgb);
    RTB_CUDA(cudaGetLastError());
}

} // namespace rtb
