// rtb_api.cu — the extern "C" boundary of librtb200.so (include/rtb200.h).
// Owns the context, copies scenes and query batches to the device, launches the
// kernels of rtb_batch.cu / rtb_wavefront.cu and translates every failure into a
// status code + message.  No exception leaves this file.  There is no host
// compute path: every entry point needs a live CUDA context.
#include "rtb_internal.hpp"

#include <cstring>
#include <new>

using namespace rtb;

namespace {

thread_local std::string g_create_error = "";

struct Cancelled {};

int fail(rtb_context *ctx, int code, const std::string &msg) {
    if (ctx)
        ctx->last_error = msg;
    else
        g_create_error = msg;
    return code;
}

template <class F> int guarded(rtb_context *ctx, F &&body) {
    try {
        cudaError_t e = cudaSetDevice(ctx->device);
        if (e != cudaSuccess)
            return fail(ctx, RTB_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
        body();
        return RTB_OK;
    } catch (const CudaError &e) {
        cudaGetLastError(); // clear the sticky-less error state
        const std::string m = e.what();
        return fail(ctx, m.find("out of memory") != std::string::npos ? RTB_ERR_OUT_OF_MEMORY : RTB_ERR_CUDA, m);
    } catch (const std::bad_alloc &) {
        return fail(ctx, RTB_ERR_OUT_OF_MEMORY, "host allocation failed");
    } catch (const std::exception &e) {
        const std::string m = e.what();
        if (m == "cancelled")
            return fail(ctx, RTB_ERR_CANCELLED, "render cancelled");
        if (m.rfind("scene", 0) == 0 || m.rfind("prim_box", 0) == 0)
            return fail(ctx, RTB_ERR_BAD_SCENE, m);
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, m);
    }
}

template <class R> void upload_typed(DeviceTyped<R> &D, const TypedTables<R> &T, cudaStream_t s, size_t &bytes) {
    D.prims.upload(T.prims, s);
    D.maux.upload(T.maux, s);
    D.ops.upload(T.ops, s);
    D.mats.upload(T.mats, s);
    D.texs.upload(T.texs, s);
    D.perlins.upload(T.perlins, s);
    D.lights.upload(T.lights, s);
    bytes += D.prims.bytes() + D.maux.bytes() + D.ops.bytes() + D.mats.bytes() + D.texs.bytes() +
             D.perlins.bytes() + D.lights.bytes();
}

int check_params(rtb_context *ctx, const rtb_render_params *p) {
    if (!p)
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "render: params is NULL");
    if (p->width < 2 || p->height < 2 || p->width > 65536 || p->height > 65536)
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "render: width/height must be in [2, 65536]");
    if (uint64_t(p->width) * uint64_t(p->height) >= (uint64_t(1) << 31))
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "render: width * height must be below 2^31 pixels");
    if (p->spp < 0 || p->max_depth < 0 || p->rr_start_depth < 0)
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "render: spp, max_depth, rr_start_depth must be >= 0");
    if (p->max_depth > 65535) // the path state packs the depth into 16 bits
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "render: max_depth must be <= 65535");
    if (p->integrator < 0 || p->integrator > RTB_INTEGRATOR_MIS)
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "render: integrator must be 0..4");
    if (p->sample_stride < 0 || p->sample_offset < 0 ||
        (p->sample_stride > 0 && p->sample_offset >= p->sample_stride))
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "render: need 0 <= sample_offset < sample_stride");
    if (p->row_stride < 0 || p->row_offset < 0 || (p->row_stride > 0 && p->row_offset >= p->row_stride) ||
        (p->row_stride == 0 && p->row_offset != 0))
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "render: need 0 <= row_offset < row_stride");
    if (p->pool_paths < 0)
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "render: pool_paths must be >= 0");
    if (!ctx->scene)
        return fail(ctx, RTB_ERR_NO_SCENE, "render: no scene uploaded");
    return RTB_OK;
}

// Stages a host batch on the device, runs `launch`, copies the result back.
template <class In, class Out, class L>
void run_batch(rtb_context *ctx, const In *in, uint64_t n, Out *out, L &&launch) {
    DeviceBuffer d_in, d_out;
    d_in.alloc(n * sizeof(In));
    d_out.alloc(n * sizeof(Out));
    RTB_CUDA(cudaMemcpyAsync(d_in.as<In>(), in, n * sizeof(In), cudaMemcpyHostToDevice, ctx->stream));
    launch(d_in.as<In>(), d_out.as<Out>());
    RTB_CUDA(cudaMemcpyAsync(out, d_out.as<Out>(), n * sizeof(Out), cudaMemcpyDeviceToHost, ctx->stream));
    RTB_CUDA(cudaStreamSynchronize(ctx->stream));
}

int check_batch(rtb_context *ctx, const void *in, const void *out, uint64_t n, int precision) {
    if (!ctx)
        return RTB_ERR_INVALID_ARGUMENT;
    if (!ctx->scene)
        return fail(ctx, RTB_ERR_NO_SCENE, "batch: no scene uploaded");
    if (n && (!in || !out))
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "batch: NULL buffer");
    if (precision != 32 && precision != 64)
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "batch: precision must be 32 or 64");
    return RTB_OK;
}

} // namespace

namespace rtb {

// ---- Distribution2D of an environment map, built on the device (environmental_light.h:146-180, :15-27) ----
// Same numbers as detail::build_env_tables (rtb_scene_host.hpp), bit for bit: every sum runs in the
// reference's order (one thread walks one row; rows are independent), products and sums are the explicit
// round-to-nearest intrinsics (no FMA contraction), and sin(theta) of the H rows comes from the host's libm.
// 2048 x 1024 texels: 69 ms on one host core + 33 MB of tables over PCIe -> two short kernels.
__global__ void k_env_func(const float *__restrict__ tex, const double *__restrict__ sin_theta, int W, int H,
                           double *__restrict__ cond_func) {
    const size_t n = size_t(W) * H;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const double r = tex[3 * i], g = tex[3 * i + 1], b = tex[3 * i + 2];
        const double lum = __dadd_rn(__dadd_rn(__dmul_rn(0.2126, r), __dmul_rn(0.7152, g)), __dmul_rn(0.0722, b));
        cond_func[i] = __dmul_rn(lum, sin_theta[i / W]);
    }
}
__global__ void k_env_rows(const double *__restrict__ cond_func, int W, int H, double *__restrict__ cond_cdf,
                           double *__restrict__ cond_int) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= H)
        return;
    const double *f = cond_func + size_t(v) * W;
    double *cdf = cond_cdf + size_t(v) * (W + 1);
    double acc = 0;
    cdf[0] = 0;
    for (int i = 1; i <= W; ++i) {
        acc = __dadd_rn(acc, f[i - 1]);
        cdf[i] = acc;
    }
    cond_int[v] = acc;
    if (acc > 0)
        for (int i = 0; i <= W; ++i)
            cdf[i] = __ddiv_rn(cdf[i], acc);
}
__global__ void k_env_marginal(const double *__restrict__ cond_int, int H, double *__restrict__ marg_cdf,
                               double *__restrict__ marg_int) {
    if (blockIdx.x || threadIdx.x)
        return;
    double acc = 0;
    marg_cdf[0] = 0;
    for (int i = 1; i <= H; ++i) {
        acc = __dadd_rn(acc, cond_int[i - 1]);
        marg_cdf[i] = acc;
    }
    *marg_int = acc;
    if (acc > 0)
        for (int i = 0; i <= H; ++i)
            marg_cdf[i] = __ddiv_rn(marg_cdf[i], acc);
}
void build_env_tables_device(rtb_context *ctx, DeviceScene &sc) {
    const HostScene &H = sc.host;
    sc.env_tables.alloc(H.env_table_doubles * sizeof(double));
    cudaStream_t st = ctx->stream;
    DeviceBuffer d_sin;
    for (const LightT<double> &l : H.f64.lights) {
        if (l.type != RTB_LIGHT_ENV || l.env_w <= 0 || l.env_h <= 0)
            continue;
        const int W = l.env_w, Hh = l.env_h;
        std::vector<double> sin_theta(Hh);
        for (int v = 0; v < Hh; ++v)
            sin_theta[v] = std::sin(Consts<double>::pi() * (v + 0.5) / Hh);
        d_sin.upload(sin_theta, st);
        double *cond_func = sc.env_tables.as<double>() + l.env_table_offset;
        double *cond_cdf = cond_func + size_t(W) * Hh;
        double *cond_int = cond_cdf + size_t(W + 1) * Hh;
        double *marg_cdf = cond_int + Hh;
        double *marg_int = marg_cdf + (Hh + 1);
        const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
        k_env_func<<<sms * 8, 256, 0, st>>>(sc.env_texels.as<float>() + l.env_texel_offset, d_sin.as<double>(), W, Hh, cond_func);
        k_env_rows<<<(Hh + 31) / 32, 32, 0, st>>>(cond_func, W, Hh, cond_cdf, cond_int);
        k_env_marginal<<<1, 32, 0, st>>>(cond_int, Hh, marg_cdf, marg_int);
        RTB_CUDA(cudaGetLastError());
        RTB_CUDA(cudaStreamSynchronize(st)); // (sin_theta and d_sin are reused by the next light)
    }
}

int check_render_params(rtb_context *ctx, const rtb_render_params *p) { return check_params(ctx, p); }

// blob -> host tables (BVH build included); throws std::runtime_error("scene: ...") on a bad blob
std::shared_ptr<const HostScene> build_scene_for(rtb_context *ctx, const void *blob, uint64_t nbytes) {
    try {
        SceneView view(blob, nbytes);
        // Large scenes: the fp64 twin of the per-primitive / per-material tables (read by the validation
        // entry points only) is built when a precision-64 call first asks for it (ensure_f64): it was a
        // quarter of the host time and a third of the bytes of every upload of the 1 M-sphere scene.
        const bool want_f64 = int64_t(view.n_prims()) <= ctx->opt_lazy_f64_prims;
        auto host = std::make_shared<HostScene>(
            build_host_scene(view, ctx->opt_max_leaf, 0.01 * ctx->opt_trav_cost_pct, ctx->opt_layout_dfs != 0, want_f64, true,
                             ctx->opt_group_boxes != 0));
        if (!want_f64)
            host->blob_copy.assign(static_cast<const char *>(blob), static_cast<const char *>(blob) + nbytes);
        return host;
    } catch (const CudaError &) {
        throw;
    } catch (const std::exception &e) {
        throw std::runtime_error(std::string("scene: ") + e.what());
    }
}

void upload_scene(rtb_context *ctx, std::shared_ptr<const HostScene> host) {
    PhaseTimer timer; // RTB200_TIMING=1
    std::unique_ptr<DeviceScene> sc(new DeviceScene(std::move(host)));
    const HostScene &H = sc->host;
    cudaStream_t s = ctx->stream;
    size_t bytes = 0;
    if (ctx->scene) { // re-upload: reuse the tables' allocations; from here on a failure leaves the context without a scene
        std::unique_ptr<DeviceScene> old = std::move(ctx->scene);
        sc->adopt_allocations(*old);
    }
    upload_typed(sc->f32, H.f32, s, bytes);
    upload_typed(sc->f64, H.f64, s, bytes); // (large tables empty when H.has_f64 is false)
    sc->f64_ready = H.has_f64;
    sc->build_max_leaf = ctx->opt_max_leaf;
    sc->build_trav_cost = 0.01 * ctx->opt_trav_cost_pct;
    sc->build_layout_dfs = ctx->opt_layout_dfs != 0;
    sc->build_group_boxes = ctx->opt_group_boxes != 0;
    sc->nodes.upload(H.nodes, s);
    sc->chains.upload(H.chains, s);
    sc->affine.upload(H.affine, s);
    sc->prim_chain.upload(H.prim_chain, s);
    sc->prim_orig.upload(H.prim_orig, s);
    sc->orig_to_sorted.upload(H.orig_to_sorted, s);
    sc->images.upload(H.images, s);
    sc->image_bytes.upload(H.image_bytes, s);
    if (timer.on)
        RTB_CUDA(cudaStreamSynchronize(s));
    timer.mark("device: tables alloc + H2D");
    if (H.env_on_device) { // straight from the caller's blob (alive until rtb_scene_upload / rtb_group_scene_upload returns)
        sc->env_texels.alloc(H.n_env_texels * sizeof(float));
        if (H.n_env_texels)
            RTB_CUDA(cudaMemcpyAsync(sc->env_texels.as<float>(), H.env_texels_src, H.n_env_texels * sizeof(float),
                                     cudaMemcpyHostToDevice, s));
    } else {
        sc->env_texels.upload(H.env_texels, s);
    }
    if (timer.on)
        RTB_CUDA(cudaStreamSynchronize(s));
    timer.mark("device: env texels alloc + H2D");
    if (H.env_on_device) {
        if (H.env_table_doubles)
            build_env_tables_device(ctx, *sc);
    } else {
        sc->env_tables.upload(H.env_tables, s);
    }
    sc->wide_nodes.upload(H.wide.qnodes, s);
    sc->wide_chain_root.upload(H.wide.chain_root, s);
    bytes += sc->wide_nodes.bytes() + sc->wide_chain_root.bytes();
    bytes += sc->nodes.bytes() + sc->chains.bytes() + sc->prim_chain.bytes() + sc->prim_orig.bytes() +
             sc->orig_to_sorted.bytes() + sc->images.bytes() + sc->image_bytes.bytes() +
             sc->env_texels.bytes() + sc->env_tables.bytes();
    sc->device_bytes = bytes;
    RTB_CUDA(cudaStreamSynchronize(s));
    timer.mark("device: env tables + wide nodes");
    ctx->scene = std::move(sc); // (frees the previous scene's tables)
    timer.mark("device: previous scene freed");
    ++ctx->scene_serial;
}

// The fp64 validation tables of a large scene, on first use: the scene is flattened again from the
// retained blob (same deterministic BVH, hence the same primitive order) with the fp64 tables on.
void ensure_f64(rtb_context *ctx) {
    DeviceScene &sc = *ctx->scene;
    if (sc.f64_ready)
        return;
    SceneView view(sc.host.blob_copy.data(), sc.host.blob_copy.size());
    const HostScene H = build_host_scene(view, sc.build_max_leaf, sc.build_trav_cost, sc.build_layout_dfs, true, false, sc.build_group_boxes);
    if (H.prim_orig != sc.host.prim_orig)
        throw std::runtime_error("internal error: the fp64 rebuild ordered the primitives differently");
    size_t bytes = 0;
    upload_typed(sc.f64, H.f64, ctx->stream, bytes);
    RTB_CUDA(cudaStreamSynchronize(ctx->stream));
    sc.device_bytes += bytes;
    sc.f64_ready = true;
}

} // namespace rtb

extern "C" {

const char *rtb_version(void) { return "rtb200 0.1.0 sm_100a"; }

int rtb_context_create(int device_id, rtb_context **out) {
    if (!out)
        return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "rtb_context_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(nullptr, RTB_ERR_NO_DEVICE,
                    std::string("no CUDA device available (") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                        "); librtb200 has no CPU path");
    }
    if (device_id < 0 || device_id >= count)
        return fail(nullptr, RTB_ERR_INVALID_ARGUMENT, "rtb_context_create: device_id out of range");
    rtb_context *ctx = new (std::nothrow) rtb_context();
    if (!ctx)
        return fail(nullptr, RTB_ERR_OUT_OF_MEMORY, "host allocation failed");
    ctx->device = device_id;
    const int rc = guarded(ctx, [&] {
        cudaDeviceProp prop;
        RTB_CUDA(cudaGetDeviceProperties(&prop, device_id));
        if (prop.major < 10)
            throw CudaError(std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                            std::to_string(prop.minor) + "; librtb200 is built for sm_100a only");
        ctx->sm_count = prop.multiProcessorCount;
        RTB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    });
    if (rc != RTB_OK) {
        g_create_error = ctx->last_error;
        delete ctx;
        return rc;
    }
    *out = ctx;
    return RTB_OK;
}

void rtb_context_destroy(rtb_context *ctx) {
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    if (ctx->stream)
        cudaStreamSynchronize(ctx->stream);
    rtb_comm_release(ctx);
    wavefront_release(ctx);
    ctx->scene.reset();
    ctx->accum.release();
    if (ctx->stream)
        cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *rtb_last_error(const rtb_context *ctx) {
    return ctx ? ctx->last_error.c_str() : g_create_error.c_str();
}

int rtb_set_option(rtb_context *ctx, int option, int64_t value) {
    if (!ctx)
        return RTB_ERR_INVALID_ARGUMENT;
    if (option == RTB_OPT_FLAT_TRAVERSAL)
        ctx->opt_flat = value != 0;
    else if (option == RTB_OPT_FUSED_SCHEDULE)
        ctx->opt_fused = value != 0;
    else if (option == RTB_OPT_BVH_MAX_LEAF)
        ctx->opt_max_leaf = int(value);
    else if (option == RTB_OPT_BVH_TRAVERSAL_COST_PCT)
        ctx->opt_trav_cost_pct = int(value);
    else if (option == RTB_OPT_BVH_LAYOUT_DFS)
        ctx->opt_layout_dfs = value != 0;
    else if (option == RTB_OPT_BINARY_TRAVERSAL)
        ctx->opt_binary_traversal = int(value);
    else if (option == RTB_OPT_LAZY_F64_PRIMS)
        ctx->opt_lazy_f64_prims = value < 0 ? 0 : value;
    else if (option == RTB_OPT_GROUP_BOXES)
        ctx->opt_group_boxes = value != 0;
    else
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_set_option: unknown option");
    return RTB_OK;
}

int rtb_scene_upload(rtb_context *ctx, const void *blob, uint64_t nbytes) {
    if (!ctx)
        return RTB_ERR_INVALID_ARGUMENT;
    if (!blob)
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "rtb_scene_upload: blob is NULL");
    return guarded(ctx, [&] { upload_scene(ctx, build_scene_for(ctx, blob, nbytes)); });
}

int rtb_scene_get_stats(rtb_context *ctx, rtb_scene_stats *out) {
    if (!ctx || !out)
        return RTB_ERR_INVALID_ARGUMENT;
    if (!ctx->scene)
        return fail(ctx, RTB_ERR_NO_SCENE, "no scene uploaded");
    const HostScene &H = ctx->scene->host;
    out->n_prims = int32_t(H.orig_to_sorted.size());
    out->n_nodes = int32_t(H.nodes.size());
    out->n_instances = H.n_instances;
    out->n_materials = int32_t(H.f32.mats.size());
    out->n_lights = int32_t(H.f32.lights.size());
    out->has_media = H.has_media ? 1 : 0;
    out->device_bytes = ctx->scene->device_bytes;
    return RTB_OK;
}

int rtb_scene_env_tables(rtb_context *ctx, double *out, uint64_t capacity, uint64_t *n_out) {
    if (!ctx)
        return RTB_ERR_INVALID_ARGUMENT;
    if (!ctx->scene)
        return fail(ctx, RTB_ERR_NO_SCENE, "rtb_scene_env_tables: no scene uploaded");
    return guarded(ctx, [&] {
        const uint64_t n = ctx->scene->env_tables.bytes() / sizeof(double);
        if (n_out)
            *n_out = n;
        const uint64_t m = n < capacity ? n : capacity;
        if (out && m) {
            RTB_CUDA(cudaMemcpyAsync(out, ctx->scene->env_tables.as<double>(), m * sizeof(double), cudaMemcpyDeviceToHost,
                                     ctx->stream));
            RTB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    });
}

int rtb_camera_derived(rtb_context *ctx, double out[24]) {
    if (!ctx || !out)
        return RTB_ERR_INVALID_ARGUMENT;
    if (!ctx->scene)
        return fail(ctx, RTB_ERR_NO_SCENE, "no scene uploaded");
    const CameraT<double> &c = ctx->scene->host.f64.camera;
    const V3<double> *vs[7] = {&c.origin, &c.lower_left_corner, &c.horizontal, &c.vertical, &c.u, &c.v, &c.w};
    for (int i = 0; i < 7; ++i) {
        out[3 * i] = vs[i]->x;
        out[3 * i + 1] = vs[i]->y;
        out[3 * i + 2] = vs[i]->z;
    }
    out[21] = c.lens_radius;
    out[22] = c.time0;
    out[23] = c.time1;
    return RTB_OK;
}

int rtb_render_device(rtb_context *ctx, const rtb_render_params *params, void *accum_rgba_device,
                      void *cuda_stream, rtb_render_stats *stats) {
    if (!ctx)
        return RTB_ERR_INVALID_ARGUMENT;
    const int rc = check_params(ctx, params);
    if (rc != RTB_OK)
        return rc;
    if (!accum_rgba_device)
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "render: accumulator buffer is NULL");
    ctx->cancel.store(0);
    return guarded(ctx, [&] {
        cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
        wavefront_render(ctx, *params, static_cast<float4 *>(accum_rgba_device), st, stats);
    });
}

int rtb_render(rtb_context *ctx, const rtb_render_params *params, float *accum_rgba_host,
               rtb_render_stats *stats) {
    if (!ctx)
        return RTB_ERR_INVALID_ARGUMENT;
    const int rc = check_params(ctx, params);
    if (rc != RTB_OK)
        return rc;
    if (!accum_rgba_host)
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "render: accumulator buffer is NULL");
    ctx->cancel.store(0);
    return guarded(ctx, [&] {
        const size_t bytes = size_t(params->width) * params->height * sizeof(float4);
        if (ctx->accum.bytes() != bytes)
            ctx->accum.alloc(bytes);
        ctx->accum_w = params->width;
        ctx->accum_h = params->height;
        wavefront_render(ctx, *params, ctx->accum.as<float4>(), ctx->stream, stats);
        RTB_CUDA(cudaMemcpyAsync(accum_rgba_host, ctx->accum.as<float4>(), bytes, cudaMemcpyDeviceToHost,
                                 ctx->stream));
        RTB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

int rtb_cancel(rtb_context *ctx) {
    if (!ctx)
        return RTB_ERR_INVALID_ARGUMENT;
    ctx->cancel.store(1);
    return RTB_OK;
}

int rtb_resolve_rgb8(rtb_context *ctx, int32_t spp, uint8_t *rgb8_host) {
    if (!ctx || !rgb8_host)
        return RTB_ERR_INVALID_ARGUMENT;
    if (spp <= 0 || ctx->accum_w == 0)
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "resolve: spp must be > 0 and a render must have run");
    return guarded(ctx, [&] {
        const size_t n = size_t(ctx->accum_w) * ctx->accum_h * 3;
        DeviceBuffer d;
        d.alloc(n);
        launch_resolve_rgb8(ctx, ctx->accum.as<float4>(), ctx->accum_w, ctx->accum_h, spp, d.as<uint8_t>(),
                            ctx->stream);
        RTB_CUDA(cudaMemcpyAsync(rgb8_host, d.as<uint8_t>(), n, cudaMemcpyDeviceToHost, ctx->stream));
        RTB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

int rtb_trace_batch(rtb_context *ctx, const rtb_ray *rays, uint64_t n, int precision, rtb_hit *hits,
                    uint64_t *visits) {
    const bool p32 = precision == 33 || precision == 34 || precision == 35 || precision == 36;
    const int rc = check_batch(ctx, rays, hits, n, p32 ? 32 : (precision == 65 ? 64 : precision));
    if (rc != RTB_OK)
        return rc;
    return guarded(ctx, [&] {
        if (precision == 64 || precision == 65)
            ensure_f64(ctx);
        DeviceBuffer d_vis;
        unsigned long long *dv = nullptr;
        if (visits) {
            d_vis.alloc(2 * sizeof(unsigned long long));
            RTB_CUDA(cudaMemsetAsync(d_vis.as<void>(), 0, 2 * sizeof(unsigned long long), ctx->stream));
            dv = d_vis.as<unsigned long long>();
        }
        if (n)
            run_batch(ctx, rays, n, hits, [&](const rtb_ray *di, rtb_hit *dout) {
                if (precision == 64 || precision == 65)
                    launch_trace_batch<double>(ctx, di, n, dout, dv, precision == 65);
                else if (precision == 33 || precision == 35)
                    launch_trace_fast_batch(ctx, di, n, dout, dv, precision == 35);
                else if (precision == 34 || precision == 36)
                    launch_trace_wide_batch(ctx, di, n, dout, dv, precision == 36);
                else
                    launch_trace_batch<float>(ctx, di, n, dout, dv);
            });
        if (visits) {
            unsigned long long h[2] = {0, 0};
            RTB_CUDA(cudaMemcpyAsync(h, dv, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
            RTB_CUDA(cudaStreamSynchronize(ctx->stream));
            visits[0] = h[0];
            visits[1] = h[1];
        }
    });
}

int rtb_bsdf_eval_batch(rtb_context *ctx, int material, const rtb_bsdf_query *queries, uint64_t n,
                        int precision, rtb_bsdf_value *out) {
    const int rc = check_batch(ctx, queries, out, n, precision);
    if (rc != RTB_OK)
        return rc;
    if (material < 0 || size_t(material) >= ctx->scene->host.f32.mats.size())
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "bsdf batch: material index out of range");
    if (!n)
        return RTB_OK;
    return guarded(ctx, [&] {
        if (precision == 64)
            ensure_f64(ctx);
        run_batch(ctx, queries, n, out, [&](const rtb_bsdf_query *di, rtb_bsdf_value *dout) {
            if (precision == 64)
                launch_bsdf_eval<double>(ctx, material, di, n, dout);
            else
                launch_bsdf_eval<float>(ctx, material, di, n, dout);
        });
    });
}

int rtb_bsdf_sample_batch(rtb_context *ctx, int material, const rtb_bsdf_query *queries, uint64_t n,
                          int precision, uint64_t seed, rtb_bsdf_sample *out) {
    const int rc = check_batch(ctx, queries, out, n, precision);
    if (rc != RTB_OK)
        return rc;
    if (material < 0 || size_t(material) >= ctx->scene->host.f32.mats.size())
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "bsdf batch: material index out of range");
    if (!n)
        return RTB_OK;
    return guarded(ctx, [&] {
        if (precision == 64)
            ensure_f64(ctx);
        run_batch(ctx, queries, n, out, [&](const rtb_bsdf_query *di, rtb_bsdf_sample *dout) {
            if (precision == 64)
                launch_bsdf_sample<double>(ctx, material, di, n, seed, dout);
            else
                launch_bsdf_sample<float>(ctx, material, di, n, seed, dout);
        });
    });
}

int rtb_light_eval_batch(rtb_context *ctx, int light, const rtb_light_query *queries, uint64_t n, int precision,
                         uint64_t seed, rtb_light_value *out) {
    const int rc = check_batch(ctx, queries, out, n, precision);
    if (rc != RTB_OK)
        return rc;
    if (light < 0 || size_t(light) >= ctx->scene->host.f32.lights.size())
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "light batch: light index out of range");
    if (!n)
        return RTB_OK;
    return guarded(ctx, [&] {
        if (precision == 64)
            ensure_f64(ctx);
        run_batch(ctx, queries, n, out, [&](const rtb_light_query *di, rtb_light_value *dout) {
            if (precision == 64)
                launch_light_eval<double>(ctx, light, di, n, seed, dout);
            else
                launch_light_eval<float>(ctx, light, di, n, seed, dout);
        });
    });
}

int rtb_texture_eval_batch(rtb_context *ctx, int texture, const double *uvp, uint64_t n, int precision,
                           double *rgb) {
    const int rc = check_batch(ctx, uvp, rgb, n, precision);
    if (rc != RTB_OK)
        return rc;
    if (texture < 0 || size_t(texture) >= ctx->scene->host.f32.texs.size())
        return fail(ctx, RTB_ERR_INVALID_ARGUMENT, "texture batch: texture index out of range");
    if (!n)
        return RTB_OK;
    return guarded(ctx, [&] {
        if (precision == 64)
            ensure_f64(ctx);
        DeviceBuffer d_in, d_out;
        d_in.alloc(n * 5 * sizeof(double));
        d_out.alloc(n * 3 * sizeof(double));
        RTB_CUDA(cudaMemcpyAsync(d_in.as<double>(), uvp, n * 5 * sizeof(double), cudaMemcpyHostToDevice,
                                 ctx->stream));
        if (precision == 64)
            launch_texture_eval<double>(ctx, texture, d_in.as<double>(), n, d_out.as<double>());
        else
            launch_texture_eval<float>(ctx, texture, d_in.as<double>(), n, d_out.as<double>());
        RTB_CUDA(cudaMemcpyAsync(rgb, d_out.as<double>(), n * 3 * sizeof(double), cudaMemcpyDeviceToHost,
                                 ctx->stream));
        RTB_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

} // extern "C"
