/* rtb200.h — C-ABI of librtb200.so, the B200 (sm_100a) path-tracing core.
 *
 * The reference renderer (JiGuang283/Ray_Tracing-Rendering) has no plugin or FFI
 * boundary: its seam is the C++ method
 *
 *     void Renderer::render(shared_ptr<hittable> world, shared_ptr<camera> cam,
 *                           const color& background, RenderBuffer& target,
 *                           const vector<shared_ptr<Light>>& lights = {})
 *                                                   (src/renderer/renderer.h:30-32)
 *
 * plus set_integrator / set_samples / set_max_depth / cancel / is_rendering
 * (renderer.h:26,104-118) and the integrator selection by id 0..4 of
 * src/main.cpp:81-100.  The host layer (ray_tracing-rendering_b200/host/) keeps
 * those C++ signatures; their implementation flattens the scene graph into the
 * blob of rtb200_scene.h and calls the entry points below.  Everything here is
 * plain C: pointers, sizes and PODs, no C++ or torch types.
 *
 * Every function returns RTB_OK (0) or a negative rtb_status; the message of the
 * last failure on a context is available from rtb_last_error().  Nothing throws
 * across the boundary.  There is NO host fallback: without a CUDA device
 * rtb_context_create() fails with RTB_ERR_NO_DEVICE.
 */
#ifndef RTB200_H
#define RTB200_H

#include <stdint.h>

#include "rtb200_scene.h"
#include "rtb200_types.h"

#if defined(__GNUC__)
#define RTB_API __attribute__((visibility("default")))
#else
#define RTB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rtb_context rtb_context;

typedef enum rtb_status {
    RTB_OK = 0,
    RTB_ERR_INVALID_ARGUMENT = -1,
    RTB_ERR_NO_DEVICE = -2, /* no CUDA device / driver: there is no CPU path */
    RTB_ERR_CUDA = -3,      /* a CUDA runtime call failed; see rtb_last_error */
    RTB_ERR_BAD_SCENE = -4, /* malformed or unsupported scene blob */
    RTB_ERR_NO_SCENE = -5,  /* render/trace before rtb_scene_upload */
    RTB_ERR_CANCELLED = -6, /* rtb_cancel() interrupted the render */
    RTB_ERR_OUT_OF_MEMORY = -7
} rtb_status;

/* Integrator ids, exactly the switch of src/main.cpp:81-100. */
enum rtb_integrator {
    RTB_INTEGRATOR_PATH = 0,   /* PathIntegrator         path_integrator.h:22-44 */
    RTB_INTEGRATOR_RR = 1,     /* RRPathInterator        rr_path_integrator.h:21-59 */
    RTB_INTEGRATOR_PBR = 2,    /* PBRPathIntegrator      pbr_path_integrator.h:21-73 */
    RTB_INTEGRATOR_DIRECT = 3, /* DirectLightIntegrator  direct_light_integrator.h:25-142 */
    RTB_INTEGRATOR_MIS = 4     /* MISPathIntegrator      mis_path_integrator.h:25-234 */
};

/* What Renderer + Integrator hold as settings (renderer.h:19-21,104-111; the
 * m_max_depth / m_rr_start_depth members of every integrator), plus the work
 * split used for multi-GPU rendering. */
typedef struct rtb_render_params {
    int32_t width;          /* RenderBuffer::width()  */
    int32_t height;         /* RenderBuffer::height() */
    int32_t spp;            /* Renderer::set_samples: samples per pixel of the WHOLE job */
    int32_t max_depth;      /* Renderer::set_max_depth (50 in main.cpp:102) */
    int32_t rr_start_depth; /* set_rr_start_depth, default 3 */
    int32_t integrator;     /* enum rtb_integrator */
    /* Sample split: this call renders the samples s of every pixel with
     * s % sample_stride == sample_offset (stride 1, offset 0 = everything).  The
     * RNG stream of a sample depends only on (pixel, s, seed), so N calls with
     * stride N and offsets 0..N-1 sum to the image of one stride-1 call. */
    int32_t sample_offset;
    int32_t sample_stride;
    uint64_t seed;
    int32_t pool_paths; /* paths resident on the device at once; 0 = default */
    int32_t flags;      /* RTB_RENDER_* */
    /* Image split (the tile queue of renderer.h:40-94 across GPUs): this call renders the rows j
     * of the image with j % row_stride == row_offset (stride 0 or 1 = all rows); pixels of other
     * rows stay zero in the accumulators.  Combines freely with the sample split. */
    int32_t row_offset;
    int32_t row_stride;
} rtb_render_params;

enum rtb_render_flags {
    RTB_RENDER_COUNT_VISITS = 1, /* also count BVH nodes visited / primitive tests (slower) */
    RTB_RENDER_TIME_EXTEND = 2,  /* bracket every launch of the dominant kernel with CUDA events (extend_ms) */
    RTB_RENDER_FORCE_WAVEFRONT = 4, /* use the wavefront schedule even where the fused one applies */
    RTB_RENDER_FORCE_FUSED = 8      /* use the fused schedule wherever the scene fits shared memory */
};

typedef struct rtb_render_stats {
    uint64_t paths;         /* camera samples started */
    uint64_t rays_closest;  /* closest-hit rays traced (extend stage) */
    uint64_t rays_shadow;   /* shadow rays traced (connect stage) */
    uint64_t nodes_visited; /* only with RTB_RENDER_COUNT_VISITS */
    uint64_t prim_tests;    /* only with RTB_RENDER_COUNT_VISITS */
    uint64_t iterations;    /* wavefront iterations */
    uint64_t kernel_launches;
    double device_ms;        /* CUDA-event time of the whole render on its stream */
    double extend_ms;        /* CUDA-event time spent in the extend kernel (0 unless timed) */
    uint64_t extend_launches;
    int32_t schedule;        /* 0 = wavefront (queues in HBM), 1 = fused (small scene, state in registers) */
    int32_t traversal;       /* wavefront schedule: 1 = binary while-while kernels, 2 = warp-scheduled 4-wide kernels,
                              * 0 = lockstep walk of the shared-memory scene copy; fused schedule: 0 */
    /* with RTB_RENDER_TIME_EXTEND on the wavefront schedule: CUDA-event time per stage, summed
     * over the iterations: [0] extend, [1] shade (all material kernels), [2] miss, [3] connect */
    double stage_ms[4];
    /* only with RTB_RENDER_COUNT_VISITS on the wavefront schedule: the most BVH nodes any single
     * closest-hit ray visited (the longest traversal sets the tail of an extend launch) */
    uint64_t max_nodes_per_ray;
    /* only with RTB_RENDER_COUNT_VISITS on the wavefront schedule: nodes visited by closest-hit
     * rays, and the sum over 32-ray chunks of the LONGEST ray's node count.  Their ratio / 32 is
     * the lane utilisation a chunk-synchronous traversal can reach at best. */
    uint64_t extend_nodes;
    uint64_t extend_chunk_max_nodes;
} rtb_render_stats;

typedef struct rtb_scene_stats {
    int32_t n_prims;     /* flat primitives in the blob */
    int32_t n_nodes;     /* 32-byte BVH nodes (top level + all bottom levels) */
    int32_t n_instances; /* distinct moving wrapper chains = bottom-level trees */
    int32_t n_materials;
    int32_t n_lights;
    int32_t has_media;
    uint64_t device_bytes; /* scene tables resident in HBM */
} rtb_scene_stats;

/* ---- context ------------------------------------------------------------------------- */

/* Binds to CUDA device `device_id` (one context per GPU; a multi-GPU job is one
 * process per GPU, each with its own context).  Replaces nothing in the reference
 * (it has no device); owns every device allocation made below. */
RTB_API int rtb_context_create(int device_id, rtb_context **out);
RTB_API void rtb_context_destroy(rtb_context *ctx);
/* Message of the last failure on ctx; with ctx == NULL, of the last failed
 * rtb_context_create on this thread.  Never NULL. */
RTB_API const char *rtb_last_error(const rtb_context *ctx);

/* Tuning switches (all default to 1):
 *   RTB_OPT_FLAT_TRAVERSAL  scenes of <= 64 primitive records are traversed in lockstep from
 *                           shared memory instead of through the BVH;
 *   RTB_OPT_FUSED_SCHEDULE  such scenes are rendered by the fused persistent kernel. */
enum rtb_option {
    RTB_OPT_FLAT_TRAVERSAL = 1,
    RTB_OPT_FUSED_SCHEDULE = 2,
    /* BVH builder knobs, applied by the NEXT rtb_scene_upload: largest leaf (1..16, default 4) and
     * the SAH cost of one traversal step in percent of one primitive test (default 100) */
    RTB_OPT_BVH_MAX_LEAF = 3,
    RTB_OPT_BVH_TRAVERSAL_COST_PCT = 4,
    /* node order: 0 = level order, 1 = sibling pairs depth-first (left subtree right after its pair) */
    RTB_OPT_BVH_LAYOUT_DFS = 5,
    /* which traversal kernels the wavefront schedule uses on BVH scenes: 0 (default) = by scene — the
     * warp-scheduled 4-wide traversal (csrc/rtb_trace.cuh) unless the tree mixes media or instances
     * with its primitives, where round 1's binary while-while kernels measure faster; 1 = always the
     * binary kernels; 2 = always the 4-wide kernels (A/B measurements, tests) */
    RTB_OPT_BINARY_TRAVERSAL = 6,
    /* scenes with more primitives than this (default 100000) get their fp64 validation tables on the
     * first precision-64 call instead of at rtb_scene_upload (a third of the upload of the 1 M-sphere
     * scene); 0 = always on first use.  Applied by the NEXT rtb_scene_upload */
    RTB_OPT_LAZY_F64_PRIMS = 7,
    /* 1 (default): the `box` objects (box.h: six rects, one material, no wrapper) of a scene too large for the
     * shared-memory kernels enter the BVH as one item each, tested by one slab test in the production kernels
     * (hits, records and shading still name the face's own rect); 0: every rect is its own tree item (A/B
     * measurements, tests).  Applied by the NEXT rtb_scene_upload */
    RTB_OPT_GROUP_BOXES = 8
};
RTB_API int rtb_set_option(rtb_context *ctx, int option, int64_t value);

/* ---- scene ---------------------------------------------------------------------------- */

/* Copies the flattened scene (rtb200_scene.h) to the device: builds the two-level
 * SAH BVH that replaces bvh_node (src/geometry/bvh.h:52-94), the fp32 production
 * tables and the fp64 validation tables (large scenes: on their first use, RTB_OPT_LAZY_F64_PRIMS).
 * A re-upload reuses the device allocations of the scene it replaces.  A blob that fails validation
 * leaves the current scene in place; a CUDA failure during the copy leaves the context without one.  The blob may be freed on return.
 * Replaces: the shared_ptr graph handed to Renderer::render (renderer.h:30). */
RTB_API int rtb_scene_upload(rtb_context *ctx, const void *blob, uint64_t nbytes);
RTB_API int rtb_scene_get_stats(rtb_context *ctx, rtb_scene_stats *out);
/* camera.h:44-50 derived members as the device holds them (fp64), in declaration
 * order: origin, lower_left_corner, horizontal, vertical, u, v, w, lens_radius,
 * time0, time1 (24 doubles). */
RTB_API int rtb_camera_derived(rtb_context *ctx, double out[24]);
/* Validation: the Distribution2D tables of the scene's environment lights (environmental_light.h:146-180,
 * :15-27) as the device built them from the uploaded texels — per env light, in light order: cond_func
 * [H*W], cond_cdf [H*(W+1)], cond_int [H], marg_cdf [H+1], marg_int [1].  Copies at most `capacity` doubles;
 * *n_out (optional) = the number the scene holds. */
RTB_API int rtb_scene_env_tables(rtb_context *ctx, double *out, uint64_t capacity, uint64_t *n_out);

/* ---- render (replaces Renderer::render, renderer.h:30-102) ----------------------------- */

/* Renders into a HOST buffer of height*width float4 (r,g,b,unused): the linear SUM
 * over this call's samples of Integrator::Li, row 0 = bottom row as in
 * RenderBuffer (render_buffer.h:17-21).  The caller divides by spp and applies the
 * reference's sqrt/clamp (renderer.h:126-140).  Includes the device->host copy. */
RTB_API int rtb_render(rtb_context *ctx, const rtb_render_params *params, float *accum_rgba_host,
               rtb_render_stats *stats);
/* Same, into a DEVICE buffer (height*width float4, zeroed by the call) on the
 * given CUDA stream (cudaStream_t as void*, NULL = the context's own stream);
 * synchronises the stream before returning.  Used by the multi-GPU driver, which
 * reduces the buffers over NCCL. */
RTB_API int rtb_render_device(rtb_context *ctx, const rtb_render_params *params, void *accum_rgba_device,
                      void *cuda_stream, rtb_render_stats *stats);
/* Safe from any thread while a render is running (Renderer::cancel, renderer.h:113):
 * the render returns RTB_ERR_CANCELLED after the current wavefront batch. */
RTB_API int rtb_cancel(rtb_context *ctx);
/* sqrt(sum/spp) clamped to [0,1] -> 8-bit RGB with the reference's truncation and
 * y flip (renderer.h:126-140, render_buffer.h:35-55), computed on the device from
 * the accumulators of the LAST rtb_render on this context.  rgb8: height*width*3. */
RTB_API int rtb_resolve_rgb8(rtb_context *ctx, int32_t spp, uint8_t *rgb8_host);

/* ---- multi-GPU (replaces the tile queue over CPU threads, renderer.h:40-94) ----------------------
 *
 * The scene is replicated; every GPU renders a slice of the samples of every pixel (rows too when
 * the job has fewer samples than GPUs; the split is planned by the library) and ONE collective —
 * ncclReduce(SUM) over NVLink of float3 means, staged by a kernel that folds in the division by
 * spp — combines them on the first GPU.  The random stream of a sample depends only on
 * (pixel, sample, seed), so the result equals the one-GPU image up to float summation order.
 * NCCL (libnccl.so.2) is bound at run time, on first use. */

/* (1) One process, several GPUs: a group owns one context, host thread, stream and NCCL
 * communicator per device — what a C++ caller of Renderer::render uses (SURVEY 8b:
 * ctx_create(device_ids, n)). */
typedef struct rtb_group rtb_group;
RTB_API int rtb_group_create(const int *device_ids, int n_devices, rtb_group **out);
RTB_API void rtb_group_destroy(rtb_group *g);
RTB_API const char *rtb_group_last_error(const rtb_group *g);
RTB_API int rtb_group_size(const rtb_group *g);
/* the i-th device's context (options, scene stats, batch queries); owned by the group */
RTB_API rtb_context *rtb_group_context(rtb_group *g, int i);
/* builds the BVH once and copies the tables to every device of the group */
RTB_API int rtb_group_scene_upload(rtb_group *g, const void *blob, uint64_t nbytes);
/* Renders the WHOLE job `params` describes (its sample / row split fields are ignored) on all
 * devices.  accum_rgba_host: height*width float4 linear sums as rtb_render returns them;
 * rgb8_host: height*width*3 bytes as rtb_resolve_rgb8 returns them; either may be NULL.
 * stats: sums over the devices, times = the slowest device. */
RTB_API int rtb_group_render(rtb_group *g, const rtb_render_params *params, float *accum_rgba_host,
                             uint8_t *rgb8_host, rtb_render_stats *stats);
RTB_API int rtb_group_cancel(rtb_group *g);

/* (2) One process per GPU (e.g. under torchrun): rank 0 makes a unique id, the launcher hands its
 * RTB_COMM_ID_BYTES bytes to every rank, every rank joins with rtb_comm_init on its own context. */
#define RTB_COMM_ID_BYTES 128
RTB_API int rtb_comm_unique_id(void *id_out);
RTB_API int rtb_comm_init(rtb_context *ctx, int n_ranks, int rank, const void *id);
RTB_API void rtb_comm_release(rtb_context *ctx);
/* This rank's slice of the WHOLE job + the reduce onto rank 0 (collective: every rank calls it with
 * the same params).  On rank 0 the complete image's float4 sums are left in the context's
 * accumulators (rtb_resolve_rgb8 / rtb_accum_copy_device read them) and copied to accum_rgba_host /
 * rgb8_host when given; other ranks pass NULL.  cuda_stream: NULL = the context's stream. */
RTB_API int rtb_render_reduce(rtb_context *ctx, const rtb_render_params *params, float *accum_rgba_host,
                              uint8_t *rgb8_host, void *cuda_stream, rtb_render_stats *stats);
/* the accumulators of the last render on ctx -> a device buffer of the caller (height*width float4) */
RTB_API int rtb_accum_copy_device(rtb_context *ctx, void *dst_device, void *cuda_stream);

/* ---- parity-layer batch entry points ---------------------------------------------------- */

/* hittable::hit of the world on caller rays (bvh.h:40-50 and everything below it).
 * precision 64: fp64 validation kernels (reference operation order, no FMA):
 *               t / primitive id bit-exact against the reference.
 *               A constant_medium's free-flight draw comes from a per-ray stream of the library.
 * precision 65 (rtb_trace_batch only): the fp64 primitive tests driven by the reference's own
 *               left-to-right walk over the leaves (blob order, no tree) with the REFERENCE's random
 *               stream: rtb_ray.reserved holds the xorshift32 state (rtweekend.h:24-34) the reference
 *               had when it answered the same query, so constant_medium::hit (constant_medium.h:55-104)
 *               is compared bit for bit as well.  O(primitives) per ray: a validation tool.
 * precision 32: the production fp32 traversal the wavefront uses.
 * precision 34 (rtb_trace_batch only): the renderer's warp-scheduled traversal of the 4-wide BVH
 *   (windows sorted by octant, lanes refilled as they finish) — the kernel k_extend runs, fed with
 *   the caller's rays; precision 36: its any-hit form (k_connect): prim >= 0 iff the ray is blocked.
 * precision 33 (rtb_trace_batch only, scenes of <= 64 primitive records): the fused kernel's own
 *   typed lockstep traversal (rects grouped by axis, box instances as slab tests).
 * precision 35 (same restriction): as 33, with the hit records of planar primitives taken from the
 *   fused kernel's per-primitive plane digest (p, normal, front_face; u = v = 0) instead of the
 *   wrapper-chain replay.
 * visits (optional, 2 x uint64): BVH nodes visited, primitive tests. */
RTB_API int rtb_trace_batch(rtb_context *ctx, const rtb_ray *rays, uint64_t n, int precision, rtb_hit *hits,
                    uint64_t *visits);
/* material::eval / pdf / emitted x2 of material `material` (material.h:27-56). */
RTB_API int rtb_bsdf_eval_batch(rtb_context *ctx, int material, const rtb_bsdf_query *queries, uint64_t n,
                        int precision, rtb_bsdf_value *out);
/* material::sample and legacy scatter (material.h:41-44, 66-69) with the device RNG
 * seeded per query from (index, seed). */
RTB_API int rtb_bsdf_sample_batch(rtb_context *ctx, int material, const rtb_bsdf_query *queries, uint64_t n,
                          int precision, uint64_t seed, rtb_bsdf_sample *out);
/* Light::sample / pdf / Le of light `light` (light.h:21-41). */
RTB_API int rtb_light_eval_batch(rtb_context *ctx, int light, const rtb_light_query *queries, uint64_t n,
                         int precision, uint64_t seed, rtb_light_value *out);
/* texture::value (texture.h:13); uvp = n x 5 doubles (u, v, px, py, pz); rgb = n x 3. */
RTB_API int rtb_texture_eval_batch(rtb_context *ctx, int texture, const double *uvp, uint64_t n, int precision,
                           double *rgb);

/* Library identification: "rtb200 <version> sm_100a". */
RTB_API const char *rtb_version(void);

#ifdef __cplusplus
}
#endif

#endif /* RTB200_H */
