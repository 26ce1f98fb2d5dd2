// rtb_multi.cu — multi-GPU rendering behind the C-ABI: what replaces the reference's tile queue
// over CPU threads (src/renderer/renderer.h:40-94) when more than one B200 is available.
//
// The path shards trivially: the scene is replicated, every GPU renders a slice of the
// samples of every pixel (and of the rows when a job has fewer samples than GPUs), and ONE
// collective combines the accumulators.  Two ways to drive it, same kernels and same collective:
//   * rtb_group_*  — one process, one host thread + stream + NCCL communicator per GPU
//     (ncclCommInitAll): what a C++ caller of Renderer::render gets;
//   * rtb_comm_init + rtb_render_reduce — one process per GPU (torchrun): the library's own
//     communicator from a unique id the launcher hands to every rank.
// The exchange: k_stage_means folds the resolve (divide by the job's spp) into the kernel that
// packs the float4 accumulators into the float3 send buffer (12 instead of 16 bytes per pixel on
// the wire), ncclReduce(SUM) over NVLink puts the mean image on the root, and k_finish_root turns
// it into what the caller asked for: float4 sums in rtb_render's layout and / or the 8-bit image
// with the reference's sqrt / clamp / truncation / y flip (renderer.h:126-140, render_buffer.h:35-55).
// NCCL is bound at run time (dlopen of libnccl.so.2 — inside a torch process that is the copy torch
// already loaded), so the library itself carries no link-time dependency on it.
#include "rtb_internal.hpp"

#include <condition_variable>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>
#include <thread>
#include <vector>

using namespace rtb;

namespace {

// ---- NCCL, bound at run time -----------------------------------------------------------------
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    std::string error;
    bool ok = false;
};

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD); // the copy already in the process
        if (!h)
            h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h)
            h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            api.error = std::string("libnccl.so.2 not found: ") + dlerror();
            return;
        }
        auto sym = [&](const char *name) {
            void *p = dlsym(h, name);
            if (!p && api.error.empty())
                api.error = std::string("NCCL symbol missing: ") + name;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.Reduce = reinterpret_cast<decltype(api.Reduce)>(sym("ncclReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
        api.ok = api.error.empty();
    });
    return api;
}

void nccl_check(ncclResult_t r, const char *what) {
    if (r != ncclSuccess)
        throw CudaError(std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error"));
}
void need_nccl() {
    if (!nccl().ok)
        throw std::runtime_error("multi-GPU: " + nccl().error);
}

// ---- the split (Python twin: distributed.plan_split) ---------------------------------------------
// Samples of every pixel are split when there are enough of them (spp >= ranks): every rank touches
// every pixel and the ranks stay balanced whatever the image shows.  With fewer samples than ranks
// the image ROWS are interleaved across groups of ranks and the samples are split inside each group.
struct Split {
    int sample_offset, sample_stride, row_offset, row_stride;
};
Split plan_split(int spp, int height, int rank, int world) {
    int sample_ways = std::max(1, std::min(world, spp));
    while (world % sample_ways)
        --sample_ways;
    const int row_ways = world / sample_ways;
    if (row_ways > std::max(height, 1))
        throw std::runtime_error("multi-GPU: more ranks than image rows x samples");
    return Split{rank % sample_ways, sample_ways, rank / sample_ways, row_ways};
}

// ---- kernels ----------------------------------------------------------------------------------
// resolve, first half, fused into the staging of the send buffer: mean contribution of this rank
__global__ void k_stage_means(const float4 *__restrict__ accum, float *__restrict__ send3, size_t n, float inv_spp) {
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const float4 a = __ldcs(accum + i);
        send3[3 * i] = a.x * inv_spp;
        send3[3 * i + 1] = a.y * inv_spp;
        send3[3 * i + 2] = a.z * inv_spp;
    }
}
// resolve, second half, on the root: the reduced mean image -> float4 sums (rtb_render's layout)
// and / or 8-bit RGB (sqrt, clamp, (uchar)(x * 255), y flip)
__global__ void k_finish_root(const float *__restrict__ mean3, int w, int h, float spp, float4 *__restrict__ sums,
                              uint8_t *__restrict__ rgb8) {
    const size_t n = size_t(w) * h;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const float r = mean3[3 * i], g = mean3[3 * i + 1], b = mean3[3 * i + 2];
        if (sums)
            sums[i] = make_float4(r * spp, g * spp, b * spp, 0.f);
        if (rgb8) {
            const size_t row = i / w, col = i - row * w; // accumulator row 0 = bottom row of the image
            const size_t o = 3 * ((size_t(h) - 1 - row) * w + col);
            // (double, as renderer.h:131-139 computes it)
            const double vr = sqrt(double(r)), vg = sqrt(double(g)), vb = sqrt(double(b));
            rgb8[o] = uint8_t((vr > 1.0 ? 1.0 : vr) * 255);
            rgb8[o + 1] = uint8_t((vg > 1.0 ? 1.0 : vg) * 255);
            rgb8[o + 2] = uint8_t((vb > 1.0 ? 1.0 : vb) * 255);
        }
    }
}

// One rank's part of a multi-GPU render up to the collective: render its slice, stage the means.
void render_slice(rtb_context *ctx, const rtb_render_params &job, int rank, int world, cudaStream_t st,
                  rtb_render_stats *stats) {
    rtb_render_params p = job;
    const Split sp = plan_split(job.spp, job.height, rank, world);
    p.sample_offset = sp.sample_offset;
    p.sample_stride = sp.sample_stride;
    p.row_offset = sp.row_offset;
    p.row_stride = sp.row_stride;
    const size_t n = size_t(job.width) * job.height;
    if (ctx->accum.bytes() != n * sizeof(float4))
        ctx->accum.alloc(n * sizeof(float4));
    ctx->accum_w = job.width;
    ctx->accum_h = job.height;
    if (ctx->stage_send.bytes() != n * 3 * sizeof(float))
        ctx->stage_send.alloc(n * 3 * sizeof(float));
    wavefront_render(ctx, p, ctx->accum.as<float4>(), st, stats);
    const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    k_stage_means<<<sms * 8, 256, 0, st>>>(ctx->accum.as<float4>(), ctx->stage_send.as<float>(), n,
                                           1.0f / float(std::max(job.spp, 1)));
    RTB_CUDA(cudaGetLastError());
}

// The collective and, on the root, the finish + copies to the caller's host buffers.
void reduce_and_finish(rtb_context *ctx, const rtb_render_params &job, cudaStream_t st, float *accum_host,
                       uint8_t *rgb8_host) {
    const size_t n = size_t(job.width) * job.height;
    const bool root = ctx->comm_rank == 0;
    if (root && ctx->stage_recv.bytes() != n * 3 * sizeof(float))
        ctx->stage_recv.alloc(n * 3 * sizeof(float));
    nccl_check(nccl().Reduce(ctx->stage_send.as<float>(), root ? ctx->stage_recv.as<float>() : nullptr, n * 3, ncclFloat,
                             ncclSum, 0, static_cast<ncclComm_t>(ctx->comm), st),
               "ncclReduce");
    if (root) {
        // the float4 sums always land in ctx->accum (rtb_resolve_rgb8 and device-side callers read them there)
        if (rgb8_host && ctx->stage_rgb8.bytes() != n * 3)
            ctx->stage_rgb8.alloc(n * 3);
        const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
        k_finish_root<<<sms * 8, 256, 0, st>>>(ctx->stage_recv.as<float>(), job.width, job.height, float(std::max(job.spp, 1)),
                                               ctx->accum.as<float4>(), rgb8_host ? ctx->stage_rgb8.as<uint8_t>() : nullptr);
        RTB_CUDA(cudaGetLastError());
        if (accum_host)
            RTB_CUDA(cudaMemcpyAsync(accum_host, ctx->accum.as<float4>(), n * sizeof(float4), cudaMemcpyDeviceToHost, st));
        if (rgb8_host)
            RTB_CUDA(cudaMemcpyAsync(rgb8_host, ctx->stage_rgb8.as<uint8_t>(), n * 3, cudaMemcpyDeviceToHost, st));
    }
    RTB_CUDA(cudaStreamSynchronize(st));
}

int classify(const std::exception &e, std::string &msg) {
    msg = e.what();
    if (dynamic_cast<const CudaError *>(&e))
        return msg.find("out of memory") != std::string::npos ? RTB_ERR_OUT_OF_MEMORY : RTB_ERR_CUDA;
    if (msg == "cancelled") {
        msg = "render cancelled";
        return RTB_ERR_CANCELLED;
    }
    if (msg.rfind("scene", 0) == 0 || msg.rfind("prim_box", 0) == 0)
        return RTB_ERR_BAD_SCENE;
    return RTB_ERR_INVALID_ARGUMENT;
}

void add_stats(rtb_render_stats &total, const rtb_render_stats &s) {
    total.paths += s.paths;
    total.rays_closest += s.rays_closest;
    total.rays_shadow += s.rays_shadow;
    total.nodes_visited += s.nodes_visited;
    total.prim_tests += s.prim_tests;
    total.iterations = std::max(total.iterations, s.iterations);
    total.kernel_launches += s.kernel_launches;
    total.device_ms = std::max(total.device_ms, s.device_ms); // the slowest GPU sets the time
    total.extend_ms = std::max(total.extend_ms, s.extend_ms);
    total.extend_launches += s.extend_launches;
    total.schedule = s.schedule;
    for (int k = 0; k < 4; ++k)
        total.stage_ms[k] = std::max(total.stage_ms[k], s.stage_ms[k]);
}

} // namespace

// One process, several GPUs.
struct rtb_group {
    std::vector<rtb_context *> ctx;
    std::vector<ncclComm_t> comms;
    std::string last_error;
};

extern "C" {

int rtb_group_create(const int *device_ids, int n_devices, rtb_group **out) {
    if (!out)
        return RTB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!device_ids || n_devices < 1)
        return RTB_ERR_INVALID_ARGUMENT;
    rtb_group *g = new (std::nothrow) rtb_group();
    if (!g)
        return RTB_ERR_OUT_OF_MEMORY;
    int rc = RTB_OK;
    for (int i = 0; i < n_devices && rc == RTB_OK; ++i) {
        rtb_context *c = nullptr;
        rc = rtb_context_create(device_ids[i], &c);
        if (rc == RTB_OK) {
            c->comm_rank = i;
            c->comm_size = n_devices;
            g->ctx.push_back(c);
        }
    }
    if (rc == RTB_OK && n_devices > 1) {
        try {
            need_nccl();
            g->comms.resize(size_t(n_devices));
            nccl_check(nccl().CommInitAll(g->comms.data(), n_devices, device_ids), "ncclCommInitAll");
            for (int i = 0; i < n_devices; ++i)
                g->ctx[size_t(i)]->comm = g->comms[size_t(i)];
        } catch (const std::exception &e) {
            std::string m;
            rc = classify(e, m);
            g->comms.clear();
        }
    }
    if (rc != RTB_OK) {
        for (rtb_context *c : g->ctx)
            rtb_context_destroy(c);
        delete g;
        return rc;
    }
    *out = g;
    return RTB_OK;
}

void rtb_group_destroy(rtb_group *g) {
    if (!g)
        return;
    for (size_t i = 0; i < g->comms.size(); ++i) {
        cudaSetDevice(g->ctx[i]->device);
        nccl().CommDestroy(g->comms[i]);
        g->ctx[i]->comm = nullptr;
    }
    for (rtb_context *c : g->ctx)
        rtb_context_destroy(c);
    delete g;
}

const char *rtb_group_last_error(const rtb_group *g) { return g ? g->last_error.c_str() : "rtb_group is NULL"; }
int rtb_group_size(const rtb_group *g) { return g ? int(g->ctx.size()) : 0; }
rtb_context *rtb_group_context(rtb_group *g, int i) {
    return (g && i >= 0 && size_t(i) < g->ctx.size()) ? g->ctx[size_t(i)] : nullptr;
}

int rtb_group_scene_upload(rtb_group *g, const void *blob, uint64_t nbytes) {
    if (!g)
        return RTB_ERR_INVALID_ARGUMENT;
    if (!blob) {
        g->last_error = "rtb_group_scene_upload: blob is NULL";
        return RTB_ERR_INVALID_ARGUMENT;
    }
    try {
        // the BVH is built ONCE; every device gets its own copy of the tables
        const std::shared_ptr<const HostScene> host = build_scene_for(g->ctx[0], blob, nbytes);
        std::vector<std::string> err(g->ctx.size());
        std::vector<int> rcs(g->ctx.size(), RTB_OK);
        std::vector<std::thread> th;
        for (size_t i = 0; i < g->ctx.size(); ++i)
            th.emplace_back([&, i] {
                try {
                    RTB_CUDA(cudaSetDevice(g->ctx[i]->device));
                    upload_scene(g->ctx[i], host);
                } catch (const std::exception &e) {
                    rcs[i] = classify(e, err[i]);
                }
            });
        for (auto &t : th)
            t.join();
        for (size_t i = 0; i < rcs.size(); ++i)
            if (rcs[i] != RTB_OK) {
                g->last_error = "device " + std::to_string(g->ctx[i]->device) + ": " + err[i];
                return rcs[i];
            }
        return RTB_OK;
    } catch (const std::exception &e) {
        return classify(e, g->last_error);
    }
}

int rtb_group_cancel(rtb_group *g) {
    if (!g)
        return RTB_ERR_INVALID_ARGUMENT;
    for (rtb_context *c : g->ctx)
        rtb_cancel(c);
    return RTB_OK;
}

// Renders the WHOLE job described by `params` (its sample / row split fields are ignored: the
// group plans the split) on all GPUs of the group.  Either output may be NULL.
int rtb_group_render(rtb_group *g, const rtb_render_params *params, float *accum_rgba_host, uint8_t *rgb8_host,
                     rtb_render_stats *stats) {
    if (!g)
        return RTB_ERR_INVALID_ARGUMENT;
    const int n = int(g->ctx.size());
    for (rtb_context *c : g->ctx) {
        const int rc = check_render_params(c, params);
        if (rc != RTB_OK) {
            g->last_error = rtb_last_error(c);
            return rc;
        }
        c->cancel.store(0);
    }
    const size_t nn = size_t(n);
    std::vector<rtb_render_stats> st(nn);
    std::vector<int> rcs(nn, RTB_OK);
    std::vector<std::string> err(nn);
    // rendezvous between "every GPU has rendered" and the collective: a rank that failed must not
    // leave the others waiting inside ncclReduce
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    bool all_ok = true;
    auto worker = [&](int i) {
        rtb_context *c = g->ctx[size_t(i)];
        bool ok = true;
        try {
            RTB_CUDA(cudaSetDevice(c->device));
            std::memset(&st[size_t(i)], 0, sizeof(rtb_render_stats));
            render_slice(c, *params, i, n, c->stream, &st[size_t(i)]);
        } catch (const std::exception &e) {
            rcs[size_t(i)] = classify(e, err[size_t(i)]);
            cudaGetLastError();
            ok = false;
        }
        {
            std::unique_lock<std::mutex> lk(mu);
            all_ok = all_ok && ok;
            if (++arrived == n)
                cv.notify_all();
            else
                cv.wait(lk, [&] { return arrived == n; });
        }
        if (!all_ok)
            return;
        try {
            if (n > 1) {
                reduce_and_finish(c, *params, c->stream, i == 0 ? accum_rgba_host : nullptr, i == 0 ? rgb8_host : nullptr);
            } else { // a group of one: no collective, the same finish
                const size_t npx = size_t(params->width) * params->height;
                if (rgb8_host) {
                    if (c->stage_rgb8.bytes() != npx * 3)
                        c->stage_rgb8.alloc(npx * 3);
                    launch_resolve_rgb8(c, c->accum.as<float4>(), params->width, params->height, std::max(params->spp, 1),
                                        c->stage_rgb8.as<uint8_t>(), c->stream);
                    RTB_CUDA(cudaMemcpyAsync(rgb8_host, c->stage_rgb8.as<uint8_t>(), npx * 3, cudaMemcpyDeviceToHost, c->stream));
                }
                if (accum_rgba_host)
                    RTB_CUDA(cudaMemcpyAsync(accum_rgba_host, c->accum.as<float4>(), npx * sizeof(float4),
                                             cudaMemcpyDeviceToHost, c->stream));
                RTB_CUDA(cudaStreamSynchronize(c->stream));
            }
        } catch (const std::exception &e) {
            rcs[size_t(i)] = classify(e, err[size_t(i)]);
            cudaGetLastError();
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < n; ++i)
        th.emplace_back(worker, i);
    worker(0);
    for (auto &t : th)
        t.join();
    for (int i = 0; i < n; ++i)
        if (rcs[size_t(i)] != RTB_OK) {
            g->last_error = "device " + std::to_string(g->ctx[size_t(i)]->device) + ": " + err[size_t(i)];
            return rcs[size_t(i)];
        }
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        for (int i = 0; i < n; ++i)
            add_stats(*stats, st[size_t(i)]);
    }
    return RTB_OK;
}

// ---- one process per GPU ---------------------------------------------------------------------------

int rtb_comm_unique_id(void *id_out) {
    if (!id_out)
        return RTB_ERR_INVALID_ARGUMENT;
    try {
        need_nccl();
        static_assert(sizeof(ncclUniqueId) == RTB_COMM_ID_BYTES, "RTB_COMM_ID_BYTES must match ncclUniqueId");
        ncclUniqueId id;
        nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
        std::memcpy(id_out, &id, sizeof(id));
        return RTB_OK;
    } catch (const std::exception &) {
        return RTB_ERR_CUDA;
    }
}

int rtb_comm_init(rtb_context *ctx, int n_ranks, int rank, const void *id) {
    if (!ctx)
        return RTB_ERR_INVALID_ARGUMENT;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks || (n_ranks > 1 && !id)) {
        ctx->last_error = "rtb_comm_init: need 0 <= rank < n_ranks and a unique id";
        return RTB_ERR_INVALID_ARGUMENT;
    }
    try {
        if (cudaSetDevice(ctx->device) != cudaSuccess)
            throw CudaError("cudaSetDevice failed");
        if (ctx->comm && ctx->comm_owned) {
            nccl().CommDestroy(static_cast<ncclComm_t>(ctx->comm));
            ctx->comm = nullptr;
        }
        ctx->comm_rank = rank;
        ctx->comm_size = n_ranks;
        if (n_ranks > 1) {
            need_nccl();
            ncclUniqueId uid;
            std::memcpy(&uid, id, sizeof(uid));
            ncclComm_t comm;
            nccl_check(nccl().CommInitRank(&comm, n_ranks, uid, rank), "ncclCommInitRank");
            ctx->comm = comm;
            ctx->comm_owned = true;
        }
        return RTB_OK;
    } catch (const std::exception &e) {
        return classify(e, ctx->last_error);
    }
}

void rtb_comm_release(rtb_context *ctx) {
    if (ctx && ctx->comm && ctx->comm_owned) {
        cudaSetDevice(ctx->device);
        nccl().CommDestroy(static_cast<ncclComm_t>(ctx->comm));
        ctx->comm = nullptr;
        ctx->comm_owned = false;
        ctx->comm_size = 1;
        ctx->comm_rank = 0;
    }
}

// This rank's slice of the WHOLE job `params` + the reduce onto rank 0.  On rank 0 the float4 sums
// of the complete image land in the context's accumulators (and in accum_rgba_host / rgb8_host when
// given); other ranks pass NULL.  cuda_stream: NULL = the context's own stream.
int rtb_render_reduce(rtb_context *ctx, const rtb_render_params *params, float *accum_rgba_host, uint8_t *rgb8_host,
                      void *cuda_stream, rtb_render_stats *stats) {
    if (!ctx)
        return RTB_ERR_INVALID_ARGUMENT;
    const int rc = check_render_params(ctx, params);
    if (rc != RTB_OK)
        return rc;
    ctx->cancel.store(0);
    try {
        if (cudaSetDevice(ctx->device) != cudaSuccess)
            throw CudaError("cudaSetDevice failed");
        cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
        if (ctx->comm_size > 1 && !ctx->comm)
            throw std::runtime_error("rtb_render_reduce: rtb_comm_init has not been called on this context");
        render_slice(ctx, *params, ctx->comm_rank, ctx->comm_size, st, stats);
        if (ctx->comm_size > 1) {
            reduce_and_finish(ctx, *params, st, ctx->comm_rank == 0 ? accum_rgba_host : nullptr,
                              ctx->comm_rank == 0 ? rgb8_host : nullptr);
        } else {
            const size_t npx = size_t(params->width) * params->height;
            if (rgb8_host) {
                if (ctx->stage_rgb8.bytes() != npx * 3)
                    ctx->stage_rgb8.alloc(npx * 3);
                launch_resolve_rgb8(ctx, ctx->accum.as<float4>(), params->width, params->height, std::max(params->spp, 1),
                                    ctx->stage_rgb8.as<uint8_t>(), st);
                RTB_CUDA(cudaMemcpyAsync(rgb8_host, ctx->stage_rgb8.as<uint8_t>(), npx * 3, cudaMemcpyDeviceToHost, st));
            }
            if (accum_rgba_host)
                RTB_CUDA(cudaMemcpyAsync(accum_rgba_host, ctx->accum.as<float4>(), npx * sizeof(float4), cudaMemcpyDeviceToHost, st));
            RTB_CUDA(cudaStreamSynchronize(st));
        }
        return RTB_OK;
    } catch (const std::exception &e) {
        cudaGetLastError();
        return classify(e, ctx->last_error);
    }
}

// Copies the float4 accumulators of the last render on this context (after rtb_render_reduce on rank
// 0: the complete image) into a DEVICE buffer of the caller, on the given stream.
int rtb_accum_copy_device(rtb_context *ctx, void *dst_device, void *cuda_stream) {
    if (!ctx || !dst_device)
        return RTB_ERR_INVALID_ARGUMENT;
    if (ctx->accum_w == 0) {
        ctx->last_error = "rtb_accum_copy_device: no render has run on this context";
        return RTB_ERR_INVALID_ARGUMENT;
    }
    try {
        cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
        RTB_CUDA(cudaMemcpyAsync(dst_device, ctx->accum.as<float4>(), size_t(ctx->accum_w) * ctx->accum_h * sizeof(float4),
                                 cudaMemcpyDeviceToDevice, st));
        RTB_CUDA(cudaStreamSynchronize(st));
        return RTB_OK;
    } catch (const std::exception &e) {
        return classify(e, ctx->last_error);
    }
}

} // extern "C"
