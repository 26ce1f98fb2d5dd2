// rtb_internal.hpp — context, device-resident scene and the launch interfaces
// between the translation units of librtb200.so.
#ifndef RTB_INTERNAL_HPP
#define RTB_INTERNAL_HPP

#include "rtb200.h"
#include "rtb_scene_host.hpp"
#include "rtb_trace.cuh"

#include <atomic>
#include <cuda_runtime.h>
#include <memory>
#include <stdexcept>
#include <string>

namespace rtb {

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string &m) : std::runtime_error(m) {}
};

#define RTB_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            throw ::rtb::CudaError(std::string(#call) + ": " + cudaGetErrorString(e_));           \
    } while (0)

// Owning device allocation.
class DeviceBuffer {
  public:
    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    ~DeviceBuffer() { release(); }
    void release() {
        if (p_)
            cudaFree(p_);
        p_ = nullptr;
        bytes_ = cap_ = 0;
    }
    // Takes over another buffer's allocation (contents undefined) so that the next alloc() of a similar size
    // costs nothing: cudaMalloc / cudaFree of tens of megabytes map and unmap pages, measured at 20-100 ms per
    // re-upload of a 58 MB scene on B200 against 3 ms for the copy itself.
    void adopt(DeviceBuffer &o) {
        release();
        p_ = o.p_;
        cap_ = o.cap_;
        o.p_ = nullptr;
        o.bytes_ = o.cap_ = 0;
    }
    void alloc(size_t bytes) {
        if (bytes == 0)
            bytes = 16; // keep pointers non-null so views are always dereferenceable
        if (p_ && bytes <= cap_ && cap_ / 2 <= bytes + 4096) { // an adopted (or previous) allocation that fits
            bytes_ = bytes;
            return;
        }
        release();
        cudaError_t e = cudaMalloc(&p_, bytes);
        if (e != cudaSuccess) {
            p_ = nullptr;
            throw CudaError(std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
        }
        bytes_ = cap_ = bytes;
    }
    template <class T> void upload(const std::vector<T> &v, cudaStream_t s) {
        alloc(v.size() * sizeof(T));
        if (!v.empty())
            RTB_CUDA(cudaMemcpyAsync(p_, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    template <class T> T *as() const { return static_cast<T *>(p_); }
    size_t bytes() const { return bytes_; }

  private:
    void *p_ = nullptr;
    size_t bytes_ = 0, cap_ = 0;
};

template <class R> struct DeviceTyped {
    DeviceBuffer prims, maux, ops, mats, texs, perlins, lights;
    void adopt(DeviceTyped &o) {
        prims.adopt(o.prims);
        maux.adopt(o.maux);
        ops.adopt(o.ops);
        mats.adopt(o.mats);
        texs.adopt(o.texs);
        perlins.adopt(o.perlins);
        lights.adopt(o.lights);
    }
};

struct DeviceScene {
    // kept for ids, counts and the host-side maps; shared between the contexts of a multi-GPU group
    // (the BVH is built once, the tables are uploaded to every device)
    std::shared_ptr<const HostScene> host_ptr;
    const HostScene &host;
    explicit DeviceScene(std::shared_ptr<const HostScene> h) : host_ptr(std::move(h)), host(*host_ptr) {}
    DeviceTyped<float> f32;
    DeviceTyped<double> f64;
    DeviceBuffer nodes, chains, affine, prim_chain, prim_orig, orig_to_sorted, images, image_bytes, env_texels,
        env_tables, wide_nodes, wide_chain_root;
    size_t device_bytes = 0;
    // the allocations of the scene this one replaces (which must not be used afterwards)
    void adopt_allocations(DeviceScene &o) {
        f32.adopt(o.f32);
        f64.adopt(o.f64);
        DeviceBuffer *mine[] = {&nodes, &chains, &affine, &prim_chain, &prim_orig, &orig_to_sorted, &images, &image_bytes,
                                &env_texels, &env_tables, &wide_nodes, &wide_chain_root};
        DeviceBuffer *theirs[] = {&o.nodes, &o.chains, &o.affine, &o.prim_chain, &o.prim_orig, &o.orig_to_sorted, &o.images,
                                  &o.image_bytes, &o.env_texels, &o.env_tables, &o.wide_nodes, &o.wide_chain_root};
        for (size_t i = 0; i < sizeof(mine) / sizeof(mine[0]); ++i)
            mine[i]->adopt(*theirs[i]);
    }
    bool f64_ready = true;     // the large fp64 validation tables are resident (rtb_api.cu ensure_f64)
    int build_max_leaf = 4;    // builder options this scene was flattened with
    double build_trav_cost = 1.0;
    bool build_layout_dfs = false;
    bool build_group_boxes = true;

    template <class R> const DeviceTyped<R> &typed() const;
    template <class R> GeomView<R> geom() const {
        const DeviceTyped<R> &T = typed<R>();
        GeomView<R> g;
        g.nodes = nodes.as<Node32>();
        g.prims = T.prims.template as<PrimT<R>>();
        g.maux = T.maux.template as<MovingAux<R>>();
        g.ops = T.ops.template as<XfOp<R>>();
        g.chains = chains.as<ChainRec>();
        g.affine = sizeof(R) == 4 ? affine.as<ChainAffine>() : nullptr;
        g.prim_chain = prim_chain.as<int32_t>();
        g.prim_orig = prim_orig.as<int32_t>();
        g.n_nodes = int32_t(host.nodes.size());
        g.n_prims = int32_t(host.prim_orig.size());
        g.n_ops = int32_t(host.f32.ops.size());
        g.n_chains = int32_t(host.chains.size());
        g.root_ref = host.root_ref;
        g.n_top = host.n_top_items;
        g.flat = host.flat_ok ? 1 : 0;
        g.n_world = host.n_world_slots;
        g.n_gated = int32_t(host.gated.size());
        for (size_t k = 0; k < host.gated.size(); ++k)
            g.gated[k] = host.gated[k];
        return g;
    }
    WideView wide() const { // the 4-wide tree of the production traversal (rtb_trace.cuh)
        WideView w;
        w.nodes = wide_nodes.as<Vec4f>();
        w.chain_root = wide_chain_root.as<uint32_t>();
        w.root_ref = host.wide.root_ref;
        w.n_nodes = uint32_t(host.wide.qnodes.size());
        w.n_global = uint32_t(host.wide.global_prims.size());
        for (int i = 0; i < kMaxGlobalPrims; ++i)
            w.global_prim[i] = size_t(i) < host.wide.global_prims.size() ? host.wide.global_prims[size_t(i)] : 0u;
        return w;
    }
    template <class R> ShadeView<R> shade() const {
        const DeviceTyped<R> &T = typed<R>();
        ShadeView<R> s;
        s.mats = T.mats.template as<MatT<R>>();
        s.texs = T.texs.template as<TexT<R>>();
        s.images = images.as<ImageRec>();
        s.image_bytes = image_bytes.as<uint8_t>();
        s.perlins = T.perlins.template as<PerlinT<R>>();
        s.lights = T.lights.template as<LightT<R>>();
        s.env_texels = env_texels.as<float>();
        s.env_tables = env_tables.as<double>();
        s.n_lights = int32_t(host.f32.lights.size());
        s.n_infinite_lights = host.n_infinite_lights;
        return s;
    }
};
template <> inline const DeviceTyped<float> &DeviceScene::typed<float>() const { return f32; }
template <> inline const DeviceTyped<double> &DeviceScene::typed<double>() const { return f64; }

struct WavefrontPool; // rtb_wavefront.cu

} // namespace rtb

struct rtb_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string last_error;
    std::unique_ptr<rtb::DeviceScene> scene;
    uint64_t scene_serial = 0;          // bumped by every rtb_scene_upload (keys per-scene decisions)
    rtb::WavefrontPool *pool = nullptr; // owned; freed by wavefront_release
    rtb::DeviceBuffer accum;            // float4 accumulators of the last rtb_render
    int accum_w = 0, accum_h = 0;
    std::atomic<int> cancel{0};
    int sm_count = 0;
    int opt_flat = 1;  // RTB_OPT_FLAT_TRAVERSAL
    int opt_fused = 1; // RTB_OPT_FUSED_SCHEDULE
    int opt_max_leaf = 4;        // RTB_OPT_BVH_MAX_LEAF
    int opt_trav_cost_pct = 100; // RTB_OPT_BVH_TRAVERSAL_COST_PCT
    int opt_layout_dfs = 0;      // RTB_OPT_BVH_LAYOUT_DFS
    int opt_binary_traversal = 0; // RTB_OPT_BINARY_TRAVERSAL: 0 by scene, 1 binary, 2 wide
    int64_t opt_lazy_f64_prims = 100000; // RTB_OPT_LAZY_F64_PRIMS
    int opt_group_boxes = 1;             // RTB_OPT_GROUP_BOXES
    // multi-GPU (rtb_multi.cu): this context's rank in an NCCL communicator and its staging buffers
    void *comm = nullptr; // ncclComm_t
    int comm_rank = 0, comm_size = 1;
    bool comm_owned = false;           // created by rtb_comm_init (a group owns its communicators itself)
    rtb::DeviceBuffer stage_send, stage_recv, stage_out, stage_rgb8;
};

namespace rtb {

// rtb_batch_f32.cu / rtb_batch_f64.cu (same source, RTB_REAL = float / double)
template <class R>
void launch_trace_batch(rtb_context *ctx, const rtb_ray *d_rays, uint64_t n, rtb_hit *d_hits,
                        unsigned long long *d_visits, bool reference_walk = false);
template <class R>
void launch_bsdf_eval(rtb_context *ctx, int material, const rtb_bsdf_query *d_q, uint64_t n,
                      rtb_bsdf_value *d_out);
template <class R>
void launch_bsdf_sample(rtb_context *ctx, int material, const rtb_bsdf_query *d_q, uint64_t n, uint64_t seed,
                        rtb_bsdf_sample *d_out);
template <class R>
void launch_light_eval(rtb_context *ctx, int light, const rtb_light_query *d_q, uint64_t n, uint64_t seed,
                       rtb_light_value *d_out);
template <class R>
void launch_texture_eval(rtb_context *ctx, int texture, const double *d_uvp, uint64_t n, double *d_rgb);

// rtb_api.cu
std::shared_ptr<const HostScene> build_scene_for(rtb_context *ctx, const void *blob, uint64_t nbytes);
void upload_scene(rtb_context *ctx, std::shared_ptr<const HostScene> host); // copies the tables to ctx's device
int check_render_params(rtb_context *ctx, const rtb_render_params *p);

// rtb_wavefront.cu
void wavefront_render(rtb_context *ctx, const rtb_render_params &p, float4 *d_accum, cudaStream_t stream,
                      rtb_render_stats *stats);
void wavefront_release(rtb_context *ctx);
void launch_trace_fast_batch(rtb_context *ctx, const rtb_ray *d_rays, uint64_t n, rtb_hit *d_hits,
                             unsigned long long *d_visits, bool plane_records);
void launch_resolve_rgb8(rtb_context *ctx, const float4 *d_accum, int w, int h, int spp, uint8_t *d_rgb8,
                         cudaStream_t stream);
// rtb_trace_batch precision 34 / 36: the renderer's own traversal kernel (closest hit / any hit)
void launch_trace_wide_batch(rtb_context *ctx, const rtb_ray *d_rays, uint64_t n, rtb_hit *d_hits,
                             unsigned long long *d_visits, bool any_hit);

} // namespace rtb

#endif // RTB_INTERNAL_HPP
