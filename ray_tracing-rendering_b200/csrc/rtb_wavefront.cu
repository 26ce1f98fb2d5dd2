// rtb_wavefront.cu — the production renderer, replacing Renderer::render + Integrator::Li
// of the reference (src/renderer/renderer.h:30-102 and the five *_integrator.h files).
//
// Two schedules over the SAME stage functions (trace, shade_surface<M>, miss_surface,
// shadow test):
//
// (A) WAVEFRONT — the general schedule (any scene size).
//   The path state TRAVELS with the queue entry: every queue is a structure of arrays of
//   16-byte vectors, read and written by consecutive lanes (one fully coalesced 512-byte
//   access per warp and array).  [Round-1 measurement behind this: queues of 32-bit indices
//   into a slot pool made every stage gather 4-5 random 16-byte vectors per path from a
//   0.3 GB pool, which B200 serves at 25-30 G sectors/s (tools/microbench: gather16 481 GB/s
//   against 6700 GB/s streaming), i.e. <= 6 G paths/s per stage.]
//     a  float4  origin.xyz, time
//     b  float4  direction.xyz (NOT normalised, as in the reference), origin primitive
//     c  float4  throughput.rgb, pixel index (0xffffffff = empty entry)
//     d  uint4   rng state (2 x u32), depth | specular_bounce << 16, prev_bsdf_pdf
//     e  float2  t, primitive index                        (hit queues only)
//   Queues: q_ext[2] (double buffered: paths that need a closest-hit ray), q_hit[8] (hits
//   sorted by material type, [6] = textured lambertians, [7] = rays that left the scene), the shadow-ray queue.
//   One iteration = k_extend -> k_shade<M> (x present material types) -> k_miss -> k_connect.
//     k_extend  persistent warps pull 32-entry chunks (next chunk's atomic issued before the
//               current trace); an EMPTY entry is refilled with the next camera sample from
//               the warp's private sample range (one global atomic per 128 samples); the hit
//               is pushed to its material queue with ONE multi-lane atomic per chunk: the
//               leaders of the __match_any groups add to the 8 counters of one 128-byte line
//               in a single instruction, which L2 serves as one transaction.
//     k_shade / k_miss  no atomics on the path: chunk c of material queue M maps to the fixed
//               position base_M + c*32 + lane of the next extend queue (base_M = entries of
//               the queues before M), where the lane writes its continued path or an empty
//               entry.  Only next-event-estimation requests are pushed (one atomic per chunk).
//   [Measured: same-line atomics serialise at 1.49 G/s on B200, so the former 6 atomics per
//   32 rays (queue head + 5 material counters in one line) alone cost 260 us per 2 Mi-ray
//   extend launch.]  Counters rotate over three sets (iteration i uses set i%3 and clears
//   set (i+2)%3): no reset kernel, no host round trip per iteration.
//
// (B) FUSED — scenes whose geometry fits in shared memory (<= 64 primitive records: the
//   Cornell-box class of BASELINE configs C1/C3/C4).  There the wavefront's costs are all
//   overhead: the scene is ~1 KB, so a bounce is a few hundred instructions, far less than
//   moving 144 B of path state through HBM and three queues.  One persistent kernel
//   (k_fused) keeps the path state in REGISTERS, traces in lockstep against a typed
//   shared-memory digest of the scene (FlatFast / traverse_flat_fast: rect records by axis, box
//   instances as slab tests, a plane digest per planar primitive for shading), shades through
//   the same shade_surface<M>, tests the shadow ray inline and regenerates finished lanes from
//   warp-private sample chunks.  The only global traffic is one vector reduction
//   (RED.ADD.v4.f32) per contribution.  The kernel is instruction-issue bound; its history in
//   executed instructions per 32-ray iteration is in DESIGN.md section 4.
//
// accum[H*W] float4: linear radiance SUMS.
#include "rtb_internal.hpp"

#include <algorithm>
#include <vector>

namespace rtb {

namespace {

constexpr int kMatTypes = RTB_MAT_TYPE_COUNT; // 6
constexpr uint32_t kFullMask = 0xffffffffu;

// Resident CTAs per SM k_fused is compiled for (register budget 65536 / (128 * N)): the kernel is
// bound by dependent-issue latency, so occupancy pays until spills take over.  Measured on B200
// with the generic lockstep traversal, C1 (legacy-API instantiation): 4 -> 30.4 ms, 6 -> 27.8,
// 7 -> 26.3, 8 -> 25.3, 10 -> 28.8, 12 -> 40.6; C3 (BSDF-API): 4 -> 68.7 ms, 6 -> 60.9, 7 -> 61.0,
// 8 -> 63.2, 10 -> 81.2.  With the typed traversal (fewer, leaner instructions, more live values):
// C1 6 -> 20.4 ms, 7 -> 19.7, 8 -> 20.3; C3 6 -> 43.0, 7 -> 47.1, 8 -> 51.3.  The all-planar
// instantiations need 55 (legacy API) / 64 (BSDF API) registers and run 9 / 8 CTAs per SM
// (C1 9.0 ms; forcing 8 or 10 CTAs: 9.14 / 11.8 ms at an earlier step).
#ifndef RTB_FLAT_UNROLL
#define RTB_FLAT_UNROLL 1
#endif
constexpr int kFlatUnroll = RTB_FLAT_UNROLL; // unroll factor of the typed rect loops of k_fused
#ifndef RTB_FUSED_MIN_BLOCKS_OLD
#define RTB_FUSED_MIN_BLOCKS_OLD 7
#endif
#ifndef RTB_FUSED_MIN_BLOCKS_NEW
#define RTB_FUSED_MIN_BLOCKS_NEW 4 // BSDF-API, lambertian + emitter scenes with spheres (SIMPLE == 1); was 6: catalogue scenes 2 / 10 / 18 / 20
#endif                             // under integrator 4: 0.96 -> 0.85, 1.91 -> 1.74, 1.56 -> 1.16, 1.01 -> 0.73 ms (400x400, 64 spp), scene 5 0.47 -> 0.50
// The general BSDF-API kernel (any material / texture / light mix; 13,700 instructions, instruction-fetch bound): FEWER
// resident warps run it faster — 6 CTAs (80 registers, 926 B spilled): C4-env 40.5 ms, scene23 fused 8.72 ms; 5 (96):
// 39.7 / 7.9; 4 (128, 110 B spilled): 38.5 / 6.78; 3 (166, none): 41.9 / 8.1; 2: 41.2 / 8.0 (B200, profiles/r02_late_knobs.txt).
#ifndef RTB_FUSED_MIN_BLOCKS_NEW_GENERAL
#define RTB_FUSED_MIN_BLOCKS_NEW_GENERAL 4
#endif
#ifndef RTB_FUSED_MIN_BLOCKS_OLD_GENERAL
#define RTB_FUSED_MIN_BLOCKS_OLD_GENERAL RTB_FUSED_MIN_BLOCKS_OLD // the legacy-API kernels that are not all-planar (SIMPLE 0 / 1): 5 and 4
                                                                 // measured 3-5 % slower over the catalogue (tools/fused_catalogue_probe.py)
#endif
#ifndef RTB_BOX_UNROLL
#define RTB_BOX_UNROLL 1 // box instances per trip of the box loop of k_fused (2: 66 registers, 7 CTAs/SM, C1 9.18 ms; 1: 55, 9 CTAs, 9.03 ms)
#endif
constexpr int kBoxUnroll = RTB_BOX_UNROLL;
#ifndef RTB_FUSED_MIN_BLOCKS_NEW_PLANAR
#define RTB_FUSED_MIN_BLOCKS_NEW_PLANAR 7 // the all-planar BSDF-API kernel (C3 at an earlier step: 6 -> 23.98 ms, 7 -> 22.95, 8 -> 22.68; with
                                          // the sample deck: 6 -> 16.76, 7 -> 16.35, 8 -> 16.62, 9 -> 17.87)
#endif
#ifndef RTB_SHADE_MIN_BLOCKS
#define RTB_SHADE_MIN_BLOCKS 1
#endif
#ifndef RTB_EXTEND_MIN_BLOCKS
#define RTB_EXTEND_MIN_BLOCKS 5 // resident CTAs per SM k_extend is compiled for (register budget 65536 / (128 * N))
#endif
// hit queues: one per material type, one for lambertians whose albedo is a procedural / image
// texture (Perlin turbulence is ~100x a solid colour: mixed into the plain lambertian queue it
// left 10 of 32 lanes active in that shade kernel on scene09), one for rays that left the scene
constexpr int kTexturedKey = kMatTypes;
constexpr int kMissKey = kMatTypes + 1;
constexpr int kKeys = kMatTypes + 2;
constexpr uint32_t kInvalidPix = 0xffffffffu; // c.w of an empty queue entry
// Resident paths of the wavefront schedule when the caller does not say: 8 Mi (6.5 GB of queues on
// a 180 GB part).  Measured (stage sums, 2 Mi / 4 Mi / 8 Mi): C5 87.0 / 81.6 / 78.5 ms, C2 52.0 /
// 47.7 / 45.3 ms — fewer, fuller launches; past 8 Mi the gain is within noise.
constexpr uint32_t kDefaultPool = 1u << 23;
constexpr uint32_t kWfChunk = 128;            // samples a warp of k_extend reserves per global atomic

// Every contended counter sits on its own 128-byte line (atomics to one line serialise in L2).
struct alignas(128) CounterLine {
    uint32_t v[32];
};
struct Counters {
    CounterLine key;         // v[k]: entries of hit queue k (8 counters, ONE line: pushed with one multi-lane atomic)
    CounterLine head_ext;    // work-distribution cursor of k_extend
    CounterLine n_shadow;    // entries of the shadow queue
    CounterLine head_shadow; // work-distribution cursor of k_connect
    CounterLine n_ext;       // entries of this iteration's extend queue (plain store by k_miss of the previous one)
};
static_assert(sizeof(Counters) == 5 * 128, "Counters layout");

struct Globals {
    unsigned long long next_sample; // next local sample index to hand out
    unsigned long long rays_closest;
    unsigned long long rays_shadow;
    unsigned long long nodes_visited;
    unsigned long long prim_tests;
    unsigned long long paths;
    unsigned long long max_nodes_per_ray;
    unsigned long long extend_nodes, extend_chunk_max_nodes;
    unsigned long long overflow; // != 0: a queue push fell outside its capacity (internal error, checked by the host)
};

struct WfParams {
    GeomView<float> geom;
    WideView wide; // the 4-wide tree k_extend_w / k_connect_w walk
    ShadeView<float> shade;
    CameraT<float> cam;
    // wavefront queues (SoA, see the file header)
    float4 *ext_a[2], *ext_b[2], *ext_c[2];
    uint4 *ext_d[2];
    float2 *ext_e[2];              // k_extend_w: (t, primitive) of an entry until its window is pushed
    float4 *hit_a, *hit_b, *hit_c; // kKeys queues of `cap` entries each, queue k at offset k * cap
    uint4 *hit_d;
    float2 *hit_e;
    float4 *sh_a, *sh_b, *sh_c;    // shadow queue
    unsigned long long *cursor;    // per warp of k_extend: private sample range [next, end)
    uint32_t cap;                  // entries per queue (P + slack for the final drain of the private ranges)
    Counters *ctr;              // 3 sets
    Globals *glob;
    float4 *accum;
    uint32_t P;
    int32_t width, height, spp, max_depth, rr_start, integrator;
    int32_t sample_offset, sample_stride;
    int32_t row_offset, row_stride; // this call's rows: row_offset + k * row_stride
    uint32_t npix;                  // pixels this call renders = width * (its rows)
    uint32_t inv_npix, inv_width;   // recip32(npix), recip32(width): see div_u32
    unsigned long long total_samples; // samples this call renders = npix * local spp
    unsigned long long window_end;    // fused schedule: samples [next_sample, window_end) this launch
    uint64_t seed;
    uint64_t seed_mixed; // mix64(seed)
    float bg[3];
    uint32_t mat_mask;
    int32_t has_media;
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// One atomic per warp: reserves a slot in a queue for every lane with want == true.
// Must be called by all 32 lanes.
__device__ __forceinline__ uint32_t warp_reserve(uint32_t *counter, bool want) {
    const uint32_t m = __ballot_sync(kFullMask, want);
    if (m == 0)
        return 0;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (int(lane_id()) == leader)
        base = atomicAdd(counter, uint32_t(__popc(m)));
    base = __shfl_sync(kFullMask, base, leader);
    return base + __popc(m & ((1u << lane_id()) - 1u));
}
// Persistent-thread work fetch: the warp takes the next 32 queue entries.
__device__ __forceinline__ uint32_t warp_fetch(uint32_t *head) {
    uint32_t base = 0;
    if (lane_id() == 0)
        base = atomicAdd(head, 32u);
    return __shfl_sync(kFullMask, base, 0);
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

// RED.ADD.v4.f32 — one vector reduction per contribution instead of three scalar ones.
__device__ __forceinline__ void accum_add(float4 *accum, uint32_t pix, V3<float> c) {
    if (c.x == 0.f && c.y == 0.f && c.z == 0.f)
        return;
    atomicAdd(accum + pix, make_float4(c.x, c.y, c.z, 0.f));
}

// ---- shared-memory copy of a small scene ---------------------------------------------------
struct FlatSmem {
    PrimT<float> prims[kFlatMaxPrims];
    XfOp<float> ops[kFlatMaxOps];
    ChainRec chains[kFlatMaxChains];
    ChainAffine affine[kFlatMaxChains];
    int32_t prim_chain[kFlatMaxPrims];
};

// Cooperative copy by the whole block of the fused kernel; the returned view's tables live in
// shared memory (and the compiler can prove it: LDS, not generic LD).
__device__ __forceinline__ GeomView<float> stage_scene_flat(const GeomView<float> &g, FlatSmem &sm) {
    const int n4 = g.n_prims * int(sizeof(PrimT<float>) / 16);
    const float4 *src = reinterpret_cast<const float4 *>(g.prims);
    float4 *dst = reinterpret_cast<float4 *>(sm.prims);
    for (int i = threadIdx.x; i < n4; i += blockDim.x)
        dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < g.n_prims; i += blockDim.x)
        sm.prim_chain[i] = g.prim_chain[i];
    for (int i = threadIdx.x; i < kFlatMaxOps; i += blockDim.x)
        if (i < g.n_ops)
            sm.ops[i] = g.ops[i];
    for (int i = threadIdx.x; i < kFlatMaxChains; i += blockDim.x)
        if (i < g.n_chains) {
            sm.chains[i] = g.chains[i];
            sm.affine[i] = g.affine[i];
        }
    __syncthreads();
    GeomView<float> s;
    s.nodes = g.nodes;
    s.prims = sm.prims;
    s.maux = g.maux;
    s.ops = sm.ops;
    s.chains = sm.chains;
    s.affine = sm.affine;
    s.prim_chain = sm.prim_chain;
    s.prim_orig = g.prim_orig;
    s.n_nodes = g.n_nodes;
    s.n_prims = g.n_prims;
    s.n_ops = g.n_ops;
    s.n_chains = g.n_chains;
    s.root_ref = g.root_ref;
    s.n_top = g.n_top;
    s.flat = 1;
    return s;
}

// FLAT_ONLY: the caller only ever runs on flat scenes (fused kernel) — the BVH code and its
// stack are not even compiled in.  Otherwise `stack_base` is the calling thread's column of
// the block's shared-memory traversal stack (kSmemStackDepth x blockDim.x words).
constexpr int kWfBlock = 128; // threads per block of every wavefront kernel
#ifndef RTB_SMEM_STACK
#define RTB_SMEM_STACK 0
#endif
// ---- typed lockstep traversal of the fused kernel ------------------------------------------------
// The generic traverse_flat() dispatches on the primitive type per primitive: measured on C1
// (ncu source view) that is ~42 instructions per rect test, 12 of them arithmetic — a jump-table
// switch, two type checks, the origin check and a BSSY/BSYNC pair around every test.  For the
// fused kernel the block digests the scene once into typed lists per SPACE (the world, then
// every instance): axis-aligned rects grouped by axis, each loop with compile-time axes and one
// branch-free predicate per rect; everything else (spheres, media) stays on the generic path.
// The set of primitives tested is the same, only the order inside a space changes (which could
// matter only for hits at bit-identical t).
struct FlatFast {
    // All fields are 32-bit and the records 16-byte aligned: one LDS.128 brings a whole record,
    // nothing is unpacked (the first version held int16 fields: the ncu source view showed ~30
    // instructions per space spent on loading and sign-extending them).
    struct alignas(16) Space {
        int32_t first[4]; // rects of this space: [first[a], first[a+1]) has constant axis a (0 = yz, 1 = xz, 2 = xy)
        int32_t chain;    // wrapper chain of the instance, -1 for the world
        int32_t reserved; // (box instances are not spaces: they live in FlatFast::box)
        int32_t sph_first, sph_end;     // indices into `sph`
        int32_t other_first, other_end; // indices into `other`
        int32_t pad[2];
    };
    // The six rects of a box (box.h:31-47) as one slab test.  face[s]: primitive index of the face
    // x = lo, x = hi, y = lo, y = hi, z = lo, z = hi; slot[r]: the same map backwards for the
    // primitives first .. first + 5.  The instance's folded transform travels with the box.
    struct alignas(16) Box {
        float xf[4];  // ChainAffine: c, s, bx, bz
        float lo[3];
        float by;     // ChainAffine: by
        float hi[3];
        uint32_t first;
        int32_t face[6];
        int32_t slot[6];
    };
    struct alignas(16) Rect {
        float k, a0, a1, b0; // plane, in-plane extents
        float b1;
        uint32_t id;         // primitive index
        float pad[2];
    };
    Box box[kFlatMaxChains];
    Rect rect[kFlatMaxPrims];
    float4 sph[kFlatMaxPrims]; // centre, radius
    int16_t sph_id[kFlatMaxPrims];
    int16_t other[kFlatMaxPrims];
    Space space[kFlatMaxChains + 1];
    int32_t n_spaces; // the world, then every instance that is not a plain box
    int32_t n_box;    // instances that are exactly one `box`: tested after the spaces
    PlaneRec plane[kFlatMaxPrims]; // shading input of planar primitives (build_plane_rec, one thread per primitive)
};

// Built by thread 0 from the staged (shared-memory) primitive table; the caller syncs afterwards.
__device__ inline void build_flat_fast(const GeomView<float> &g, FlatFast &ff) {
    int n_rect = 0, n_other = 0, n_sph = 0, n_spaces = 0, n_box = 0;
    // six rects, two per axis, that close up into one axis-aligned box?
    auto as_box = [&](uint32_t begin, uint32_t end, FlatFast::Box &bx) -> bool {
        if (end - begin != 6)
            return false;
        int cnt[3] = {0, 0, 0};
        bx.first = begin;
        for (uint32_t i = begin; i < end; ++i) {
            const PrimT<float> p = g.prims[i];
            const uint32_t type = p.type_mat & PT_TYPE_MASK;
            const int ax = type == PT_YZ ? 0 : type == PT_XZ ? 1 : type == PT_XY ? 2 : -1;
            if (ax < 0 || cnt[ax] >= 2)
                return false;
            // in-plane extents: YZ -> (y, z), XZ -> (x, z), XY -> (x, y)
            const int A = ax == 0 ? 1 : 0, B = ax == 2 ? 1 : 2;
            if (cnt[ax] == 0) {
                bx.lo[ax] = p.d[4];
                bx.face[2 * ax] = int32_t(i);
            } else {
                bx.hi[ax] = p.d[4];
                bx.face[2 * ax + 1] = int32_t(i);
                if (bx.hi[ax] < bx.lo[ax]) {
                    const float tf = bx.lo[ax];
                    bx.lo[ax] = bx.hi[ax];
                    bx.hi[ax] = tf;
                    const int32_t ti = bx.face[2 * ax];
                    bx.face[2 * ax] = bx.face[2 * ax + 1];
                    bx.face[2 * ax + 1] = ti;
                }
            }
            ++cnt[ax];
            (void)A;
            (void)B;
        }
        if (cnt[0] != 2 || cnt[1] != 2 || cnt[2] != 2 || !(bx.lo[0] < bx.hi[0]) || !(bx.lo[1] < bx.hi[1]) ||
            !(bx.lo[2] < bx.hi[2]))
            return false;
        for (uint32_t i = begin; i < end; ++i) { // every face must span exactly the other two extents
            const PrimT<float> p = g.prims[i];
            const uint32_t type = p.type_mat & PT_TYPE_MASK;
            const int ax = type == PT_YZ ? 0 : type == PT_XZ ? 1 : 2;
            const int A = ax == 0 ? 1 : 0, B = ax == 2 ? 1 : 2;
            if (p.d[0] != bx.lo[A] || p.d[1] != bx.hi[A] || p.d[2] != bx.lo[B] || p.d[3] != bx.hi[B])
                return false;
        }
        for (int s6 = 0; s6 < 6; ++s6)
            bx.slot[uint32_t(bx.face[s6]) - begin] = s6;
        return true;
    };
    auto add_space = [&](uint32_t begin, uint32_t end, int chain, bool top_level) {
        if (!top_level && n_box < kFlatMaxChains && as_box(begin, end, ff.box[n_box])) {
            // a box instance is not a space of its own: the traversal walks ff.box after the spaces
            const ChainAffine a = g.affine[chain];
            FlatFast::Box &bx = ff.box[n_box++];
            bx.xf[0] = a.c;
            bx.xf[1] = a.s;
            bx.xf[2] = a.bx;
            bx.xf[3] = a.bz;
            bx.by = a.by;
            return;
        }
        FlatFast::Space &sp = ff.space[n_spaces++];
        sp.chain = chain;
        sp.reserved = 0;
        const uint32_t axis_type[3] = {PT_YZ, PT_XZ, PT_XY}; // constant axis 0, 1, 2
        for (int a = 0; a < 3; ++a) {
            sp.first[a] = n_rect;
            for (uint32_t i = begin; i < end; ++i) {
                const PrimT<float> p = g.prims[i];
                if ((p.type_mat & PT_TYPE_MASK) != axis_type[a])
                    continue;
                FlatFast::Rect &r = ff.rect[n_rect++];
                r.k = p.d[4];
                r.a0 = p.d[0];
                r.a1 = p.d[1];
                r.b0 = p.d[2];
                r.b1 = p.d[3];
                r.id = i;
            }
        }
        sp.first[3] = n_rect;
        sp.other_first = n_other;
        sp.sph_first = n_sph;
        for (uint32_t i = begin; i < end; ++i) {
            const PrimT<float> p = g.prims[i];
            const uint32_t type = p.type_mat & PT_TYPE_MASK;
            if (type == PT_XY || type == PT_XZ || type == PT_YZ || (top_level && type == PT_INSTANCE))
                continue;
            if (type == PT_SPHERE) {
                ff.sph[n_sph] = make_float4(p.d[0], p.d[1], p.d[2], p.d[3]);
                ff.sph_id[n_sph++] = int16_t(i);
            } else {
                ff.other[n_other++] = int16_t(i);
            }
        }
        sp.other_end = n_other;
        sp.sph_end = n_sph;
    };
    const uint32_t n_top = uint32_t(g.n_top);
    add_space(0, n_top, -1, true);
    for (uint32_t i = 0; i < n_top && n_spaces <= kFlatMaxChains; ++i) {
        const PrimT<float> p = g.prims[i];
        if ((p.type_mat & PT_TYPE_MASK) != PT_INSTANCE)
            continue;
        const uint32_t first = uint32_t(p.d[0]);
        add_space(first, first + uint32_t(p.d[1]), int(p.aux2), false);
    }
    ff.n_spaces = n_spaces;
    ff.n_box = n_box;
}

// Closest hit of the rects [first, last) whose constant axis is AX (in-plane axes A, B).
template <int AX, int A, int B, bool COUNT>
__device__ __forceinline__ void flat_rects(const FlatFast &ff, int first, int last, V3<float> o, V3<float> d,
                                           float idir_ax, float t_min, float &t_max, uint32_t origin, uint32_t &best,
                                           uint64_t &tests) {
    const FlatFast::Rect *r = ff.rect + first, *const r_end = ff.rect + last;
#pragma unroll kFlatUnroll
    for (; r != r_end; ++r) {
        const float4 ra = *reinterpret_cast<const float4 *>(&r->k); // k, a0, a1, b0
        const float2 rb = *reinterpret_cast<const float2 *>(&r->b1); // b1, id
        const float t = (ra.x - o[AX]) * idir_ax; // the arithmetic of hit_rect(), bit for bit
        const float a = fmaf(t, d[A], o[A]);
        const float b = fmaf(t, d[B], o[B]);
        const uint32_t id = __float_as_uint(rb.y);
        // seven compares folded into ONE predicate chain (setp.and): the compiler's own choice was a
        // tree of six FSETP + three PLOP3, i.e. three instructions more per rect on an issue-bound
        // kernel.  Same comparisons, same NaN behaviour (every ordered compare fails).
        asm("{\n\t.reg .pred p;\n\t"
            "setp.ge.ftz.f32 p, %2, %4;\n\t"
            "setp.le.and.ftz.f32 p, %2, %1, p;\n\t"
            "setp.ge.and.ftz.f32 p, %5, %6, p;\n\t"
            "setp.le.and.ftz.f32 p, %5, %7, p;\n\t"
            "setp.ge.and.ftz.f32 p, %8, %9, p;\n\t"
            "setp.le.and.ftz.f32 p, %8, %10, p;\n\t"
            "setp.ne.and.u32 p, %11, %3, p;\n\t"
            "selp.u32 %0, %11, %0, p;\n\t"
            "selp.f32 %1, %2, %1, p;\n\t}"
            : "+r"(best), "+f"(t_max)
            : "f"(t), "r"(origin), "f"(t_min), "f"(a), "f"(ra.y), "f"(ra.z), "f"(b), "f"(ra.w), "f"(rb.x), "r"(id));
    }
    if (COUNT)
        tests += uint64_t(last - first);
}

// All six faces of a box at once.  The plane distances are the rect test's own
// (k - o) * idir, so a hit has the same t bit for bit; which FACE is hit is decided by the slab
// ordering instead of the in-rectangle tests, which can differ from the rect-by-rect answer only
// for rays within rounding distance of a box edge.
template <bool COUNT>
__device__ __forceinline__ void flat_box(const FlatFast::Box &bx, V3<float> wo, V3<float> wd, float t_min, float &t_max,
                                         uint32_t origin, uint32_t &best, uint64_t &tests) {
    // the ray in the box's object space: the arithmetic of enter_instance<float, true>
    const float4 xf = *reinterpret_cast<const float4 *>(bx.xf);
    const float4 lob = *reinterpret_cast<const float4 *>(bx.lo);
    const V3<float> o(xf.x * wo.x - xf.y * wo.z + xf.z, wo.y + lob.w, xf.y * wo.x + xf.x * wo.z + xf.w);
    const V3<float> d(xf.x * wd.x - xf.y * wd.z, wd.y, xf.y * wd.x + xf.x * wd.z);
    const V3<float> idir = safe_inv(d);
    const float4 hif = *reinterpret_cast<const float4 *>(bx.hi); // hi, first
    const float tx0 = (lob.x - o.x) * idir.x, tx1 = (hif.x - o.x) * idir.x;
    const float ty0 = (lob.y - o.y) * idir.y, ty1 = (hif.y - o.y) * idir.y;
    const float tz0 = (lob.z - o.z) * idir.z, tz1 = (hif.z - o.z) * idir.z;
    const float tn = fmaxf(fmaxf(fminf(tx0, tx1), fminf(ty0, ty1)), fminf(tz0, tz1));
    const float tf = fminf(fminf(fmaxf(tx0, tx1), fmaxf(ty0, ty1)), fmaxf(tz0, tz1));
    if (COUNT)
        tests += 6;
    // Branch-free (the ncu source view of the branchy version: its two divergent blocks — "the ray
    // starts on this box" and "which face" — ran at 3 of 32 lanes and made up 13 % of the kernel's
    // instructions).  A ray that starts ON this box cannot enter it: its only candidate is the far
    // side tf, and heading outward that far side is the very face it starts on (tf ~ 0), which the
    // face != origin test rejects; heading inward (a transmitted ray) it leaves through another face.
    const bool mine = (origin - __float_as_uint(hif.w)) < 6u;
    const bool enter = !mine & (tn >= t_min);
    const float cand = enter ? tn : tf;
    // which face: the first of tx0, tx1, ty0, ty1, tz0, tz1 that IS cand.  Written as five
    // setp / selp pairs over the six face ids (two vector loads): left to the compiler the ?: chain
    // came back as nested divergent branches plus a dependent shared-memory look-up.
    const uint4 f03 = *reinterpret_cast<const uint4 *>(bx.face);
    const uint2 f45 = *reinterpret_cast<const uint2 *>(bx.face + 4);
    uint32_t face;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.eq.ftz.f32 p, %1, %6;\n\tselp.u32 %0, %11, %12, p;\n\t"
        "setp.eq.ftz.f32 p, %1, %5;\n\tselp.u32 %0, %10, %0, p;\n\t"
        "setp.eq.ftz.f32 p, %1, %4;\n\tselp.u32 %0, %9, %0, p;\n\t"
        "setp.eq.ftz.f32 p, %1, %3;\n\tselp.u32 %0, %8, %0, p;\n\t"
        "setp.eq.ftz.f32 p, %1, %2;\n\tselp.u32 %0, %7, %0, p;\n\t}"
        : "=&r"(face)
        : "f"(cand), "f"(tx0), "f"(tx1), "f"(ty0), "f"(ty1), "f"(tz0), "r"(f03.x), "r"(f03.y), "r"(f03.z), "r"(f03.w),
          "r"(f45.x), "r"(f45.y));
    const bool ok = (tn <= tf) & (cand >= t_min) & (cand <= t_max) & (face != origin);
    best = ok ? face : best;
    t_max = ok ? cand : t_max;
}

// RECTS_ONLY: the host checked that the scene holds nothing but axis-aligned rects (all_planar),
// so the sphere / moving sphere / medium loops are not compiled into the kernel's hot loop.
template <bool ANY, bool COUNT, bool RECTS_ONLY = false, class Rng>
__device__ __forceinline__ uint32_t traverse_flat_fast(const GeomView<float> &g, const FlatFast &ff, V3<float> o,
                                                       V3<float> d, float time, float t_min, float t_max,
                                                       uint32_t origin, Rng &rng, float &t_hit, uint64_t &nodes,
                                                       uint64_t &tests) {
    uint32_t best = kNoPrim;
    const int n_spaces = ff.n_spaces;
    for (int s = 0; s < n_spaces; ++s) {
        const FlatFast::Space sp = ff.space[s];
        V3<float> lo = o, ld = d;
        if (sp.chain >= 0) {
            enter_instance<float, true>(g, sp.chain, lo, ld);
            if (COUNT)
                ++nodes;
        }
        const V3<float> lid = safe_inv(ld);
        flat_rects<0, 1, 2, COUNT>(ff, sp.first[0], sp.first[1], lo, ld, lid.x, t_min, t_max, origin, best, tests);
        flat_rects<1, 0, 2, COUNT>(ff, sp.first[1], sp.first[2], lo, ld, lid.y, t_min, t_max, origin, best, tests);
        flat_rects<2, 0, 1, COUNT>(ff, sp.first[2], sp.first[3], lo, ld, lid.z, t_min, t_max, origin, best, tests);
        if (RECTS_ONLY) {
            if (ANY && best != kNoPrim)
                break;
            continue;
        }
#pragma unroll 1
        for (int k = sp.sph_first; k < sp.sph_end; ++k) {
            const float4 c = ff.sph[k];
            const uint32_t i = uint32_t(ff.sph_id[k]);
            float t;
            if (hit_sphere<float, true>(V3<float>(c.x, c.y, c.z), c.w, lo, ld, t_min, t_max, i == origin, t)) {
                best = i;
                t_max = t;
            }
        }
        if (COUNT)
            tests += uint64_t(sp.sph_end - sp.sph_first);
#pragma unroll 1
        for (int k = sp.other_first; k < sp.other_end; ++k) { // moving spheres, media
            const uint32_t i = uint32_t(ff.other[k]);
            const PrimT<float> p = g.prims[i];
            const uint32_t type = p.type_mat & PT_TYPE_MASK;
            if (COUNT)
                ++tests;
            float t;
            bool h;
            if (type == PT_MEDIUM) {
                h = hit_medium<float, true>(g, p, lo, ld, time, t_min, t_max, rng(), t);
                if (p.type_mat & PT_DUP_LEAF) {
                    float t2;
                    if (hit_medium<float, true>(g, p, lo, ld, time, t_min, h ? t : t_max, rng(), t2)) {
                        h = true;
                        t = t2;
                    }
                }
            } else {
                h = hit_simple<float, true>(g, p, type, lo, ld, lid, time, t_min, t_max, i == origin, t);
            }
            if (h) {
                best = i;
                t_max = t;
            }
        }
        if (ANY && best != kNoPrim)
            break;
    }
    const int n_box = ff.n_box;
#pragma unroll kBoxUnroll
    for (int b = 0; b < n_box; ++b) {
        if (ANY && best != kNoPrim)
            break;
        if (COUNT)
            ++nodes; // one "node" = one instance entry (ray transform)
        flat_box<COUNT>(ff.box[b], o, d, t_min, t_max, origin, best, tests);
    }
    t_hit = t_max;
    return best;
}

// Lanes without a ray pass active = false.
template <bool ANY, bool COUNT, bool FLAT_ONLY = false, bool MEDIA = true, bool INST = true, class Rng>
__device__ __forceinline__ uint32_t trace(const GeomView<float> &g, bool active, V3<float> o, V3<float> d, float time,
                                          float t_min, float t_max, uint32_t origin, Rng &rng, float &t,
                                          uint64_t &nodes, uint64_t &tests, uint32_t *stack_base = nullptr) {
    if (FLAT_ONLY || g.flat) {
        t = t_max;
        if (!active)
            return kNoPrim;
        return traverse_flat<float, ANY, true>(g, o, d, time, t_min, t_max, origin, rng, t,
                                               COUNT ? &nodes : nullptr, COUNT ? &tests : nullptr);
    }
#if RTB_SMEM_STACK
    uint32_t spill[kStackDepth - kSmemStackDepth];
    SmemStack<kWfBlock> stack(stack_base, spill);
#else
    uint32_t storage[kStackDepth];
    LocalStack stack(storage);
#endif
    t = t_max;
    if (!active)
        return kNoPrim;
    return traverse<float, ANY, true, Rng, decltype(stack), MEDIA, INST>(g, o, d, time, t_min, t_max, origin, rng, t,
                                                                         COUNT ? &nodes : nullptr,
                                                                         COUNT ? &tests : nullptr, stack);
}

// ---- path state ----------------------------------------------------------------------------
struct PathState {
    V3<float> o, d, T;
    float time, prev_pdf;
    uint32_t origin_prim, pix, depth;
    bool spec;
    Pcg rng;
};

struct ShadowReq { // a next-event-estimation sample waiting for its visibility test
    bool want;
    V3<float> o, d, c; // origin, direction (segment or unit), weighted contribution
    float tmax;
    uint32_t origin;
};

__device__ __forceinline__ float4 pack_a(const PathState &s) { return make_float4(s.o.x, s.o.y, s.o.z, s.time); }
__device__ __forceinline__ float4 pack_b(const PathState &s) {
    return make_float4(s.d.x, s.d.y, s.d.z, __uint_as_float(s.origin_prim));
}
__device__ __forceinline__ float4 pack_c(const PathState &s) {
    return make_float4(s.T.x, s.T.y, s.T.z, __uint_as_float(s.pix));
}
__device__ __forceinline__ uint4 pack_d(const PathState &s) {
    return make_uint4(uint32_t(s.rng.s), uint32_t(s.rng.s >> 32), s.depth | (s.spec ? 0x10000u : 0u),
                      __float_as_uint(s.prev_pdf));
}
__device__ __forceinline__ PathState unpack(float4 a, float4 b, float4 c, uint4 x) {
    PathState s;
    s.o = V3<float>(a.x, a.y, a.z);
    s.time = a.w;
    s.d = V3<float>(b.x, b.y, b.z);
    s.origin_prim = __float_as_uint(b.w);
    s.T = V3<float>(c.x, c.y, c.z);
    s.pix = __float_as_uint(c.w);
    s.rng.s = uint64_t(x.x) | (uint64_t(x.y) << 32);
    s.depth = x.z & 0xffffu;
    s.spec = (x.z & 0x10000u) != 0;
    s.prev_pdf = __uint_as_float(x.w);
    return s;
}

// n / d and n % d for a divisor known at launch: inv = floor(2^32 / d) (0xffffffff for d = 1)
// puts umulhi(n, inv) at floor(n / d) or one below it, so one conditional step finishes the job:
// 5 instructions instead of the ~20 of the general 32-bit division, on a path every regenerated
// lane walks.
inline uint32_t recip32(uint32_t d) { return d <= 1 ? 0xffffffffu : uint32_t((1ull << 32) / d); }
__device__ __forceinline__ uint32_t div_u32(uint32_t n, uint32_t d, uint32_t inv, uint32_t &rem) {
    uint32_t q = __umulhi(n, inv);
    uint32_t r = n - q * d;
    const bool more = r >= d;
    rem = more ? r - d : r;
    return more ? q + 1 : q;
}

// Local sample index g -> (pixel, sample-in-pixel) in sample-major order, so concurrently
// resident paths belong to different pixels and consecutive indices are neighbouring pixels.
__device__ __forceinline__ void decode_sample(const WfParams &p, unsigned long long g, uint32_t &pix, uint32_t &smp) {
    uint32_t k;
    if (p.total_samples <= 0xffffffffull) { // 32-bit: ~5x cheaper than the 64-bit divide
        k = div_u32(uint32_t(g), p.npix, p.inv_npix, pix);
    } else {
        k = uint32_t(g / p.npix);
        pix = uint32_t(g - (unsigned long long)k * p.npix);
    }
    smp = uint32_t(p.sample_offset) + k * uint32_t(p.sample_stride);
    if (p.row_stride > 1) { // pix counts the pixels of this call's rows: map to the image
        uint32_t i;
        const uint32_t jl = div_u32(pix, uint32_t(p.width), p.inv_width, i);
        pix = (uint32_t(p.row_offset) + jl * uint32_t(p.row_stride)) * uint32_t(p.width) + i;
    }
}
__device__ __forceinline__ Pcg sample_stream(const WfParams &p, uint32_t pix, uint32_t smp) {
    return pcg_seed_premixed((unsigned long long)pix * uint32_t(p.spp) + smp, p.seed_mixed);
}

// renderer.h:72-75: one camera sample.
__device__ __forceinline__ void new_path(const WfParams &p, uint32_t pix, uint32_t smp, PathState &s) {
    s.rng = sample_stream(p, pix, smp);
    RngT<float> r;
    r.g = s.rng;
    uint32_t i;
    const uint32_t j = div_u32(pix, uint32_t(p.width), p.inv_width, i);
    const float u = (float(i) + r.next()) / float(p.width - 1);
    const float v = (float(j) + r.next()) / float(p.height - 1);
    camera_ray(p.cam, u, v, r, s.o, s.d, s.time);
    s.rng = r.g;
    s.T = V3<float>(1, 1, 1);
    s.pix = pix;
    s.depth = 0;
    s.spec = false;
    s.prev_pdf = 0.f;
    s.origin_prim = kNoPrim;
}
__device__ __forceinline__ void new_path(const WfParams &p, unsigned long long g, PathState &s) {
    uint32_t pix, smp;
    decode_sample(p, g, pix, smp);
    new_path(p, pix, smp, s);
}
// The generator state new_path() leaves behind, recomputed from the sample's identity: the
// camera consumes a fixed number of draws (2 jitter + 2 lens if the aperture is open + 1 time),
// so k_extend does not have to keep the state in registers across the traversal.
__device__ __forceinline__ Pcg rng_after_camera(const WfParams &p, uint32_t pix, uint32_t smp) {
    Pcg r = sample_stream(p, pix, smp);
    const int draws = p.cam.lens_radius != 0.f ? 5 : 3;
    for (int i = 0; i < draws; ++i)
        r.s = r.s * 6364136223846793005ULL + 1442695040888963407ULL;
    return r;
}

// ---- integrator stage functions (shared by both schedules) ----------------------------------

// mis_path_integrator.h:154-162
__device__ __forceinline__ V3<float> clamp_radiance(V3<float> L, float max_value) {
    if (L.x > max_value || L.y > max_value || L.z > max_value) {
        const float max_c = max3(L);
        if (max_c > max_value)
            return L * (max_value / max_c);
    }
    return L;
}
// mis_path_integrator.h:165-170
__device__ __forceinline__ float power_heuristic(float a, float b) {
    const float a2 = a * a, b2 = b * b, denom = a2 + b2;
    return denom > 0.f ? a2 / denom : 0.f;
}
// The light table in shared memory (north_star: "stage the top BVH levels and light tables in shared
// memory"): with RTB_SMEM_LIGHTS=1 every kernel that shades (k_shade, k_shade_minor, k_miss, k_fused) copies
// the first kSmemLights records (100 bytes each; no reference scene has more than four lights) at its start
// and light_at() serves them from there.  MEASURED on B200 (round 2, same box, same run order): C3 17.61 ->
// 17.73 ms, C4-env 42.57 -> 44.49 ms, C5 shade stage 486 -> 493 ms — slower everywhere: a light record is read
// with warp-uniform addresses, which the read-only path already serves as one L1 broadcast per field, while the
// shared-memory copy adds a block barrier per launch, a bounds select per access and generic-space loads.  Off
// by default; the top BVH levels, where lanes read DIFFERENT nodes, are staged (rtb_trace.cuh).
#ifndef RTB_SMEM_LIGHTS
#define RTB_SMEM_LIGHTS 0
#endif
constexpr int kSmemLights = 8;
#if RTB_SMEM_LIGHTS
__shared__ LightT<float> s_light_table[kSmemLights];
#endif
__device__ __forceinline__ void stage_lights(const WfParams &p) {
#if RTB_SMEM_LIGHTS
    const int n = p.shade.n_lights < kSmemLights ? p.shade.n_lights : kSmemLights;
    constexpr int kWords = int(sizeof(LightT<float>) / 4);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(p.shade.lights);
    uint32_t *dst = reinterpret_cast<uint32_t *>(s_light_table);
    for (int i = threadIdx.x; i < n * kWords; i += blockDim.x)
        dst[i] = __ldg(src + i);
    __syncthreads();
#else
    (void)p;
#endif
}
__device__ __forceinline__ const LightT<float> &light_at(const WfParams &p, int i) {
#if RTB_SMEM_LIGHTS
    return i < kSmemLights ? s_light_table[i] : p.shade.lights[i];
#else
    return p.shade.lights[i];
#endif
}
// mis_path_integrator.h:173-188 (also :53-60): sum over ALL lights of pdf(o,d)/N
// (out of line: two call sites, emitted-radiance MIS and miss MIS, each would inline light_pdf of every light type)
__device__ __noinline__ float all_lights_pdf(const WfParams &p, V3<float> o, V3<float> d) {
    float total = 0.f;
    const float sel = 1.0f / float(p.shade.n_lights);
    for (int i = 0; i < p.shade.n_lights; ++i)
        total += light_pdf(p.shade, light_at(p, i), o, d) * sel;
    return total;
}

// Everything Integrator::Li does at a surface hit on a material of type M.  OLD =
// integrators 0/1 (legacy scatter() + two-sided emitted(u,v,p)), otherwise integrators
// 2/3/4 (sample/eval/pdf + one-sided emitted(rec,wo), NEE, MIS).  Updates the path in
// place (next ray, throughput, depth, rng), adds emission to the image, and returns the
// NEE sample (if any) for the caller to test for visibility.
template <int M, bool OLD, bool ALL_PLANAR = false>
__device__ __forceinline__ void shade_surface(const WfParams &p, const GeomView<float> &g, PathState &s, float t,
                                              uint32_t pi, bool &alive, ShadowReq &sh, const PlaneRec *plane = nullptr) {
    sh.want = false;
    // `plane` (fused kernel): the per-primitive digest of planar primitives in shared memory.
    // ALL_PLANAR: the host checked that every primitive has one and that every albedo / emission
    // texture is a solid colour, which the kernel copied into the digest: neither make_record()
    // nor tex_value() is compiled in and shading reads no global table.
    MatT<float> m;
    if (ALL_PLANAR) {
        m = MatT<float>{};
        m.flags = 2;
        m.color[0] = plane[pi].color[0];
        m.color[1] = plane[pi].color[1];
        m.color[2] = plane[pi].color[2];
    } else {
        m = p.shade.mats[g.prims[pi].type_mat >> PT_MAT_SHIFT];
    }
    // M >= 0: the type is a compile-time constant, which prunes the per-type switches (the material-sorted
    // k_shade<M> kernels, the SIMPLE fused kernels).  M < 0: the record's own type, ONE copy of the code for all
    // types (the general fused kernel: six inlined copies were 25,600 instructions = 410 KB, and ncu showed that
    // kernel waiting for instructions — `no_instruction` 10 warps per issue, issue slots 30 % on C4-env).
    if (M >= 0)
        m.type = M;
    const int MT = M >= 0 ? M : m.type;
    RecT<float> rec;
    if (ALL_PLANAR || (plane != nullptr && !(m.flags & 1) && plane[pi].valid))
        rec = plane_record<float>(plane[pi], s.o, s.d, t);
    else if (M < 0) // (one copy of make_record in the one-copy kernel; the sphere's acos / atan2 only for textures that read u,v)
        rec = make_record<float, true, true>(g, pi, s.o, s.d, s.time, t, (m.flags & 1) != 0);
    else
        rec = (m.flags & 1) ? make_record<float, true, true>(g, pi, s.o, s.d, s.time, t)
                            : make_record<float, true, false>(g, pi, s.o, s.d, s.time, t);
    RngT<float> rng;
    rng.g = s.rng;
    alive = true;
    if (OLD) {
        // path_integrator.h:32-44, rr_path_integrator.h:35-57
        if (MT == RTB_MAT_DIFFUSE_LIGHT)
            accum_add(p.accum, s.pix, s.T * mat_emitted_old(p.shade, m, rec));
        V3<float> atten, dout;
        if (!mat_scatter(p.shade, m, rec, s.d, rng, atten, dout)) {
            alive = false;
        } else {
            s.T = s.T * atten;
            if (p.integrator == RTB_INTEGRATOR_RR && int(s.depth) >= p.rr_start) {
                const float ps = clamp_(max3(s.T), 0.005f, 0.95f);
                if (rng.next() > ps)
                    alive = false;
                else
                    s.T = s.T / ps;
            }
            s.o = rec.p;
            s.d = dout;
            s.origin_prim = pi;
        }
    } else {
        const V3<float> wo = -unit_vector(s.d);
        if (MT == RTB_MAT_DIFFUSE_LIGHT) {
            const V3<float> e = mat_emitted_new(p.shade, m, rec);
            if (p.integrator == RTB_INTEGRATOR_PBR) {
                accum_add(p.accum, s.pix, s.T * e); // pbr_path_integrator.h:38-39
            } else if (p.integrator == RTB_INTEGRATOR_DIRECT) {
                if (s.depth == 0 || s.spec) // direct_light_integrator.h:52-55
                    accum_add(p.accum, s.pix, s.T * e);
            } else if (length_squared(e) > 0.f) { // mis_path_integrator.h:72-94
                V3<float> Le;
                if (s.depth == 0 || s.spec)
                    Le = s.T * e;
                else if (p.shade.n_lights > 0)
                    Le = (s.T * e) * power_heuristic(s.prev_pdf, all_lights_pdf(p, s.o, s.d));
                else
                    Le = s.T * e;
                accum_add(p.accum, s.pix, s.depth == 0 ? Le : clamp_radiance(Le, 100.f));
            }
        }
        s.spec = false; // material::is_specular() is never overridden (material.h:37)
        // next-event estimation: direct_light_integrator.h:98-142, mis_path_integrator.h:192-234.
        // eval() is identically 0 for metal / dielectric / diffuse_light / isotropic, so only
        // lambertian and PBR can contribute; the others skip the (wasted) shadow ray.
        if ((MT == RTB_MAT_LAMBERTIAN || MT == RTB_MAT_PBR) && p.integrator >= RTB_INTEGRATOR_DIRECT &&
            p.shade.n_lights > 0) {
            const int nl = p.shade.n_lights;
            int li = int(float(nl) * rng.next()); // random_int(0, n-1), rtweekend.h:48-50
            li = li < nl ? li : nl - 1;
            const float sel = 1.0f / float(nl);
            const float u0 = rng.next(), u1 = rng.next();
            const LightT<float> &L = light_at(p, li);
            const LightSampleT<float> ls = light_sample(p.shade, L, rec.p, u0, u1, rng);
            if (ls.pdf > 0.f && length_squared(ls.Li) > 0.f) {
                const V3<float> f = mat_eval(p.shade, m, rec, wo, ls.wi);
                const float cos_theta = fabsf(dot(ls.wi, rec.normal));
                V3<float> Ld;
                if (p.integrator == RTB_INTEGRATOR_DIRECT) {
                    Ld = ls.is_delta ? (f * ls.Li) * (cos_theta / sel) : (f * ls.Li) * (cos_theta / (ls.pdf * sel));
                    // direct_light_integrator.h:133-139: sequential per-channel rescale
                    if (Ld.x > 100.f)
                        Ld = Ld * (100.f / Ld.x);
                    if (Ld.y > 100.f)
                        Ld = Ld * (100.f / Ld.y);
                    if (Ld.z > 100.f)
                        Ld = Ld * (100.f / Ld.z);
                    Ld = s.T * Ld;
                } else {
                    if (ls.is_delta) {
                        Ld = (f * ls.Li) * (cos_theta / sel);
                    } else {
                        const float bsdf_pdf = mat_pdf(p.shade, m, rec, wo, ls.wi);
                        const float light_p = ls.pdf * sel;
                        Ld = (f * ls.Li) * (cos_theta * power_heuristic(light_p, bsdf_pdf) / light_p);
                    }
                    Ld = clamp_radiance(s.T * Ld, 100.f);
                }
                if (Ld.x != 0.f || Ld.y != 0.f || Ld.z != 0.f) {
                    sh.want = true;
                    sh.o = rec.p;
                    sh.c = Ld;
                    sh.origin = pi;
                    if (isfinite(ls.dist)) {
                        // segment form: d = light_point - p, t in [0.001/dist, 1 - 0.001/dist]; the
                        // same interval as (wi, [0.001, dist - 0.001]) of the reference, but
                        // well-conditioned in fp32 (the far end is exactly t = 1)
                        sh.d = ls.to_light;
                        sh.tmax = 1.0f - 0.001f / ls.dist;
                    } else {
                        sh.d = ls.wi;
                        sh.tmax = Consts<float>::inf();
                    }
                }
            }
        }
        BsdfSampleT<float> bs;
        if (!mat_sample(p.shade, m, rec, wo, rng, bs)) {
            // mis_path_integrator.h:106-117: only integrator 4 falls back to scatter()
            V3<float> atten, dout;
            if (p.integrator == RTB_INTEGRATOR_MIS && mat_scatter(p.shade, m, rec, s.d, rng, atten, dout)) {
                s.T = s.T * atten;
                s.d = dout;
                s.spec = false;
                s.prev_pdf = 0.f;
            } else {
                alive = false;
            }
        } else if (bs.pdf < 1e-8f && !bs.is_specular) {
            alive = false;
        } else {
            s.spec = bs.is_specular;
            s.prev_pdf = bs.is_specular ? 0.f : bs.pdf;
            const float cos_theta = fabsf(dot(bs.wi, rec.normal));
            s.T = bs.is_specular ? s.T * bs.f : s.T * (bs.f * (cos_theta / bs.pdf));
            s.d = bs.wi;
        }
        if (alive) {
            s.o = rec.p;
            s.origin_prim = pi;
            if (int(s.depth) >= p.rr_start) { // e.g. mis_path_integrator.h:137-146
                const float ps = clamp_(max3(s.T), 0.05f, 0.95f);
                if (rng.next() > ps)
                    alive = false;
                else
                    s.T = s.T / ps;
            }
        }
    }
    s.rng = rng.g;
    s.depth += 1;
    if (int(s.depth) >= p.max_depth)
        alive = false;
    // fp32 guard: a degenerate continuation (zero or non-finite direction / origin) would make
    // every slab test pass and drag a NaN through up to max_depth full-tree traversals
    // (two compares that a NaN fails: no branch)
    const float sd = fabsf(s.d.x) + fabsf(s.d.y) + fabsf(s.d.z);
    const float chk = sd + fabsf(s.o.x) + fabsf(s.o.y) + fabsf(s.o.z);
    alive = alive & (chk < Consts<float>::inf()) & (sd > 0.f);
}

// The ray left the scene: background for integrators 0-2 (e.g. rr_path_integrator.h:30-33),
// environment lights for 3/4 (direct_light_integrator.h:35-48, mis_path_integrator.h:37-67).
// OLD: the caller is instantiated for the legacy-API integrators 0 / 1 only, which never look at
// environment lights: that code (acos / atan2, fp64 table look-ups) is left out of the kernel.
template <bool OLD = false>
__device__ __forceinline__ void miss_surface(const WfParams &p, const PathState &s) {
    const V3<float> bg(p.bg[0], p.bg[1], p.bg[2]);
    V3<float> L = s.T * bg;
    if (!OLD && p.integrator >= RTB_INTEGRATOR_DIRECT && p.shade.n_infinite_lights > 0) {
        // Le of every environment light and — where the MIS weight wants it — the sum over ALL lights of pdf / N
        // (all_lights_pdf), with ONE direction -> (u, v, theta) conversion per map shared by the two (acos + atan2 were
        // 9 % of the general fused kernel's instructions on C4-env, a third of them this duplicate and unread sphere uv)
        const bool want_pdf = !(p.integrator == RTB_INTEGRATOR_DIRECT || s.depth == 0 || s.spec);
        V3<float> env(0, 0, 0);
        float pdf_sum = 0.f;
        const V3<float> ud = unit_vector(s.d);
        for (int i = 0; i < p.shade.n_lights; ++i) {
            const LightT<float> &l = light_at(p, i);
            if (l.type == RTB_LIGHT_ENV) {
                if (l.env_w == 0) {
                    env = env + V3<float>(1, 1, 1);
                    pdf_sum += 1.0f / (4.0f * Consts<float>::pi());
                } else {
                    float u, v, theta;
                    env_dir_to_uv(l, ud, u, v, theta);
                    env = env + env_Le_uv(p.shade, l, u, v);
                    if (want_pdf)
                        pdf_sum += env_pdf_uv(p.shade, l, u, v, theta);
                }
            } else if (want_pdf) {
                pdf_sum += light_pdf(p.shade, l, s.o, s.d);
            }
        }
        L = want_pdf ? (s.T * env) * power_heuristic(s.prev_pdf, pdf_sum / float(p.shade.n_lights)) : s.T * env;
    }
    accum_add(p.accum, s.pix, L);
}

struct PathDraw { // RNG adaptor handed to the traversal for constant_medium tests
    Pcg *g;
    __device__ float operator()() { return g->next_open(); }
};

// Visibility of one NEE sample: scene.hit(shadow_ray, 0.001, dist - 0.001)
// (direct_light_integrator.h:115-130, mis_path_integrator.h:209-230).  Shadow rays carry
// time 0 regardless of the path's time (direct_light_integrator.h:115).
template <bool COUNT, bool FLAT_ONLY = false, bool MEDIA = true, bool INST = true>
__device__ __forceinline__ bool shadow_visible(const GeomView<float> &g, bool active, V3<float> o, V3<float> d,
                                               float tmax, uint32_t origin, Pcg &rng, uint64_t &nodes, uint64_t &tests,
                                               uint32_t *stack_base = nullptr) {
    // t_min is 0.001 along the UNIT direction; the stored direction may be the unnormalised segment
    const float len = isfinite(tmax) ? length(d) : 1.0f;
    PathDraw draw{&rng};
    float t;
    return trace<true, COUNT, FLAT_ONLY, MEDIA, INST>(g, active, o, d, 0.0f, 0.001f / len, tmax, origin, draw, t, nodes,
                                                      tests, stack_base) == kNoPrim;
}

// ---- (A) wavefront kernels -------------------------------------------------------------------
// Queue entries are touched exactly once per kernel: they are read with ld.global.cs and
// written with st.global.cs (evict-first), so that the streaming path state does not push
// the BVH nodes and primitive records out of the 126 MB L2.

// glob == nullptr: only the queue counters (start of a later segment of the same render)
__global__ void k_clear(Counters *ctr, Globals *glob) {
    for (uint32_t i = threadIdx.x; i < 3 * sizeof(Counters) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t *>(ctr)[i] = 0;
    if (threadIdx.x == 0 && glob) {
        glob->next_sample = 0;
        glob->rays_closest = glob->rays_shadow = glob->nodes_visited = glob->prim_tests = glob->paths = 0;
        glob->max_nodes_per_ray = 0;
        glob->extend_nodes = glob->extend_chunk_max_nodes = 0;
        glob->overflow = 0;
    }
}

// Initial state: n0 empty entries in extend queue 0 (k_extend fills them with camera samples),
// every warp's private sample range empty.
__global__ void __launch_bounds__(256) k_init(WfParams p, int it, uint32_t n0, uint32_t n_cursor) {
    const uint32_t stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t i = t0; i < n0; i += stride)
        p.ext_c[it & 1][i] = make_float4(0.f, 0.f, 0.f, __uint_as_float(kInvalidPix));
    for (uint32_t i = t0; i < 2 * n_cursor; i += stride)
        p.cursor[i] = 0ull;
    if (t0 == 0)
        p.ctr[it % 3].n_ext.v[0] = n0;
}

// extend: closest hit for every queued path; an empty entry is first refilled with the next
// camera sample; the result moves, with the path state, into the queue of the material it hit
// (or the miss queue).  Replaces scene.hit(current_ray, 0.001, infinity, rec) (e.g.
// rr_path_integrator.h:29), everything under bvh_node::hit and the pixel loop of
// renderer.h:66-80.
template <bool COUNT, bool MEDIA, bool INST>
__global__ void __launch_bounds__(kWfBlock, RTB_EXTEND_MIN_BLOCKS) k_extend(WfParams p, int it) {
#if RTB_SMEM_STACK
    __shared__ uint32_t s_stack[kSmemStackDepth * kWfBlock];
#else
    uint32_t *const s_stack = nullptr; // the traversal stack lives in local memory (measured faster)
#endif
    const GeomView<float> &g = p.geom;
    Counters &C = p.ctr[it % 3];
    if (blockIdx.x == 0)
        for (uint32_t i = threadIdx.x; i < sizeof(Counters) / 4; i += blockDim.x)
            reinterpret_cast<uint32_t *>(&p.ctr[(it + 2) % 3])[i] = 0;
    const uint32_t n = C.n_ext.v[0];
    const int buf = it & 1;
    const float4 *__restrict__ in_a = p.ext_a[buf], *__restrict__ in_b = p.ext_b[buf], *__restrict__ in_c = p.ext_c[buf];
    const uint4 *__restrict__ in_d = p.ext_d[buf];
    const uint32_t lane = lane_id();
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned long long chunk_next = p.cursor[2 * wid], chunk_end = p.cursor[2 * wid + 1]; // warp-uniform
    bool global_done = false;                                                            // warp-uniform
    uint64_t nodes = 0, tests = 0, max_nodes = 0, ext_nodes = 0, chunk_max = 0;
    uint32_t n_new = 0;
    uint32_t base = warp_fetch(&C.head_ext.v[0]);
    while (true) {
        const bool drain = base >= n;
        if (drain) {
            // The queue is done.  Once the global sample counter is exhausted the rest of this
            // warp's private range would never be asked for: start those paths now.
            const unsigned long long end = chunk_end < p.total_samples ? chunk_end : p.total_samples;
            if (chunk_next >= end)
                break;
            unsigned long long handed = 0;
            if (lane == 0)
                handed = *reinterpret_cast<volatile unsigned long long *>(&p.glob->next_sample);
            handed = __shfl_sync(kFullMask, handed, 0);
            if (handed < p.total_samples)
                break;
        }
        uint32_t next_base = 0xffffffffu;
        if (!drain && lane == 0) // the next chunk's atomic travels while this chunk is traced
            next_base = atomicAdd(&C.head_ext.v[0], 32u);
        const uint32_t idx = base + lane;
        const bool have = !drain && idx < n;
        uint32_t pix = kInvalidPix, smp = 0xffffffffu; // smp != ~0: this lane started a new path
        if (have)
            pix = __float_as_uint(__ldcs(in_c + idx).w);
        // ---- refill: every empty lane takes the next sample of the warp's private range
        const bool need = drain || (have && pix == kInvalidPix);
        const uint32_t m = __ballot_sync(kFullMask, need);
        bool fresh = false;
        PathState s;
        if (m) {
            const uint32_t cnt = __popc(m), rank = __popc(m & ((1u << lane) - 1u));
            unsigned long long avail = chunk_end - chunk_next;
            unsigned long long id;
            if (avail >= cnt || drain || global_done) {
                id = rank < avail ? chunk_next + rank : ~0ull;
                chunk_next += cnt < avail ? cnt : avail;
            } else {
                unsigned long long nb = 0;
                if (lane == 0)
                    nb = atomicAdd(&p.glob->next_sample, (unsigned long long)kWfChunk);
                nb = __shfl_sync(kFullMask, nb, 0);
                id = rank < avail ? chunk_next + rank : nb + (rank - avail);
                chunk_next = nb + (cnt - avail);
                chunk_end = nb + kWfChunk;
                if (nb >= p.total_samples) {
                    global_done = true;
                    chunk_next = chunk_end = p.total_samples;
                }
            }
            if (need && id < p.total_samples) {
                decode_sample(p, id, pix, smp);
                new_path(p, pix, smp, s);
                fresh = true;
            }
        }
        bool active = fresh || (have && pix != kInvalidPix);
        uint32_t key = kKeys; // kKeys = nothing to push
        uint32_t ray_nodes = 0;
        float4 a = make_float4(0, 0, 0, 0), b = a;
        float2 e = make_float2(0.f, 0.f);
        Pcg rg;
        rg.s = 0;
        if (active) {
            if (fresh) {
                a = pack_a(s);
                b = pack_b(s);
                if (MEDIA)
                    rg = s.rng;
            } else {
                a = __ldcs(in_a + idx);
                b = __ldcs(in_b + idx);
                if (MEDIA) { // constant_medium::hit draws from the path's stream
                    const uint4 x = __ldcs(in_d + idx);
                    rg.s = uint64_t(x.x) | (uint64_t(x.y) << 32);
                }
            }
            PathDraw draw{&rg};
            float t;
            const uint64_t nodes_before = nodes;
            const uint32_t pi = trace<false, COUNT, false, MEDIA, INST>(g, true, V3<float>(a.x, a.y, a.z),
                                                                  V3<float>(b.x, b.y, b.z), a.w, 0.001f,
                                                                  Consts<float>::inf(), __float_as_uint(b.w), draw, t,
                                                                  nodes, tests, s_stack + threadIdx.x);
            if (COUNT) {
                ray_nodes = uint32_t(nodes - nodes_before);
                if (ray_nodes > max_nodes)
                    max_nodes = ray_nodes;
            }
            e = make_float2(t, __uint_as_float(pi));
            key = kMissKey;
            if (pi != kNoPrim) {
                const MatT<float> &hm = p.shade.mats[g.prims[pi].type_mat >> PT_MAT_SHIFT];
                key = uint32_t(hm.type);
                if (key == RTB_MAT_LAMBERTIAN && !(hm.flags & 2))
                    key = kTexturedKey;
            }
        }
        if (COUNT) {
            ext_nodes += ray_nodes;
            chunk_max += __reduce_max_sync(kFullMask, ray_nodes);
        }
        // ---- push: the leaders of the equal-key groups reserve in ONE atomic instruction (the
        // counters share a 128-byte line, so L2 sees a single transaction)
        const uint32_t peers = __match_any_sync(kFullMask, key);
        const int leader = __ffs(peers) - 1;
        uint32_t pos = 0;
        if (int(lane) == leader && key < uint32_t(kKeys))
            pos = atomicAdd(&C.key.v[key], uint32_t(__popc(peers)));
        pos = __shfl_sync(kFullMask, pos, leader) + __popc(peers & ((1u << lane) - 1u));
        if (active && pos >= p.cap) { // cannot happen while the capacity bound of WavefrontPool::ensure holds
            atomicExch(&p.glob->overflow, 1ull);
            active = false;
        }
        if (active) {
            float4 c;
            uint4 d;
            if (fresh) {
                c = make_float4(1.f, 1.f, 1.f, __uint_as_float(pix));
                const Pcg r = MEDIA ? rg : rng_after_camera(p, pix, smp);
                d = make_uint4(uint32_t(r.s), uint32_t(r.s >> 32), 0u, 0u);
                ++n_new;
            } else {
                c = __ldcs(in_c + idx);
                d = __ldcs(in_d + idx);
                if (MEDIA) {
                    d.x = uint32_t(rg.s);
                    d.y = uint32_t(rg.s >> 32);
                }
            }
            const size_t o = size_t(key) * p.cap + pos;
            __stcs(p.hit_a + o, a);
            __stcs(p.hit_b + o, b);
            __stcs(p.hit_c + o, c);
            __stcs(p.hit_d + o, d);
            __stcs(p.hit_e + o, e);
        }
        if (drain)
            base = n; // stay in drain mode until the private range is empty
        else
            base = __shfl_sync(kFullMask, next_base, 0);
    }
    if (lane == 0) {
        p.cursor[2 * wid] = chunk_next;
        p.cursor[2 * wid + 1] = chunk_end;
    }
    const uint32_t started = __reduce_add_sync(kFullMask, n_new);
    if (lane == 0 && started)
        atomicAdd(&p.glob->paths, (unsigned long long)started);
    if (COUNT) {
        const unsigned long long x = warp_sum(nodes), y = warp_sum(tests);
        if (lane == 0) {
            atomicAdd(&p.glob->nodes_visited, x);
            atomicAdd(&p.glob->prim_tests, y);
        }
        atomicMax(&p.glob->max_nodes_per_ray, (unsigned long long)max_nodes);
        const unsigned long long z = warp_sum(ext_nodes);
        if (lane == 0) {
            atomicAdd(&p.glob->extend_nodes, z);
            atomicAdd(&p.glob->extend_chunk_max_nodes, (unsigned long long)chunk_max);
        }
    }
}

// Position of hit queue `q`'s entries in the next extend queue: after the entries of the
// queues before it.
__device__ __forceinline__ uint32_t out_base(const Counters &C, int q) {
    uint32_t b = 0;
    for (int k = 0; k < q; ++k)
        b += C.key.v[k];
    return b;
}

// shade_<material>: shade_surface<M> for the paths whose hit landed on material type M.  The
// continued path (or an empty entry where the path ended) goes to a FIXED position of the
// next extend queue, so nothing here contends: warps stride over the chunks statically.
template <int M, bool OLD>
__device__ __forceinline__ void shade_queue(const WfParams &p, int it, int q) {
    const GeomView<float> &g = p.geom;
    Counters &C = p.ctr[it % 3];
    const uint32_t n = C.key.v[q];
    const uint32_t base_out = out_base(C, q);
    const int nb = (it + 1) & 1;
    const size_t qoff = size_t(q) * p.cap;
    const uint32_t n_warps = gridDim.x * (blockDim.x >> 5), wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (uint32_t first = wid * 32u; first < n; first += n_warps * 32u) {
        const uint32_t idx = first + lane_id();
        const bool valid = idx < n;
        bool alive = false;
        PathState s;
        ShadowReq sh;
        sh.want = false;
        uint32_t sh_pix = 0;
        if (valid) {
            const size_t o = qoff + idx;
            const float4 a = __ldcs(p.hit_a + o), b = __ldcs(p.hit_b + o), c = __ldcs(p.hit_c + o);
            const uint4 d = __ldcs(p.hit_d + o);
            const float2 e = __ldcs(p.hit_e + o);
            s = unpack(a, b, c, d);
            sh_pix = s.pix;
            shade_surface<M, OLD>(p, g, s, e.x, __float_as_uint(e.y), alive, sh);
        }
        if (!OLD && (M == RTB_MAT_LAMBERTIAN || M == RTB_MAT_PBR)) { // the only materials with a non-zero eval()
            const uint32_t pos = warp_reserve(&C.n_shadow.v[0], sh.want);
            if (sh.want && pos >= p.cap) {
                atomicExch(&p.glob->overflow, 2ull);
                sh.want = false;
            }
            if (sh.want) {
                __stcs(p.sh_a + pos, make_float4(sh.o.x, sh.o.y, sh.o.z, sh.tmax));
                __stcs(p.sh_b + pos, make_float4(sh.d.x, sh.d.y, sh.d.z, __uint_as_float(sh_pix)));
                __stcs(p.sh_c + pos, make_float4(sh.c.x, sh.c.y, sh.c.z, __uint_as_float(sh.origin)));
            }
        }
        if (valid) {
            const uint32_t o = base_out + idx;
            if (alive) {
                __stcs(p.ext_a[nb] + o, pack_a(s));
                __stcs(p.ext_b[nb] + o, pack_b(s));
                __stcs(p.ext_c[nb] + o, pack_c(s));
                __stcs(p.ext_d[nb] + o, pack_d(s));
            } else {
                __stcs(p.ext_c[nb] + o, make_float4(0.f, 0.f, 0.f, __uint_as_float(kInvalidPix)));
            }
        }
    }
}
template <int M, bool OLD>
__global__ void __launch_bounds__(kWfBlock, RTB_SHADE_MIN_BLOCKS) k_shade(WfParams p, int it, int q) {
    if (!OLD)
        stage_lights(p);
    shade_queue<M, OLD>(p, it, q);
}
// The cheap, usually sparsely hit material types (metal, dielectric, emitter, isotropic) in ONE
// launch, queue after queue: each still runs its own shade_surface<M> over its own queue, so
// nothing diverges, but three launches (and their tails) per iteration are saved.
template <bool OLD> __global__ void __launch_bounds__(kWfBlock, RTB_SHADE_MIN_BLOCKS) k_shade_minor(WfParams p, int it) {
    if (!OLD)
        stage_lights(p);
    if ((p.mat_mask >> RTB_MAT_METAL) & 1u)
        shade_queue<RTB_MAT_METAL, OLD>(p, it, RTB_MAT_METAL);
    if ((p.mat_mask >> RTB_MAT_DIELECTRIC) & 1u)
        shade_queue<RTB_MAT_DIELECTRIC, OLD>(p, it, RTB_MAT_DIELECTRIC);
    if ((p.mat_mask >> RTB_MAT_DIFFUSE_LIGHT) & 1u)
        shade_queue<RTB_MAT_DIFFUSE_LIGHT, OLD>(p, it, RTB_MAT_DIFFUSE_LIGHT);
    if ((p.mat_mask >> RTB_MAT_ISOTROPIC) & 1u)
        shade_queue<RTB_MAT_ISOTROPIC, OLD>(p, it, RTB_MAT_ISOTROPIC);
}

// miss: the rays that left the scene add background / environment radiance; their entries of
// the next extend queue become empty.  SHADE = false when that radiance is identically zero
// (black background, no environment light): then nothing is read at all.
template <bool SHADE> __global__ void __launch_bounds__(128) k_miss(WfParams p, int it) {
    if (SHADE)
        stage_lights(p);
    Counters &C = p.ctr[it % 3];
    const uint32_t n = C.key.v[kMissKey];
    const uint32_t base_out = out_base(C, kMissKey);
    const int nb = (it + 1) & 1;
    const size_t qoff = size_t(kMissKey) * p.cap;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        p.ctr[(it + 1) % 3].n_ext.v[0] = base_out + n;
        atomicAdd(&p.glob->rays_closest, (unsigned long long)(base_out + n));
    }
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
        if (SHADE) {
            const size_t o = qoff + idx;
            const PathState s = unpack(__ldcs(p.hit_a + o), __ldcs(p.hit_b + o), __ldcs(p.hit_c + o), __ldcs(p.hit_d + o));
            miss_surface(p, s);
        }
        __stcs(p.ext_c[nb] + base_out + idx, make_float4(0.f, 0.f, 0.f, __uint_as_float(kInvalidPix)));
    }
}

// connect: any-hit test of the shadow rays queued by shade; unoccluded ones add their
// (already weighted) contribution.
template <bool COUNT, bool MEDIA, bool INST>
__global__ void __launch_bounds__(kWfBlock, RTB_EXTEND_MIN_BLOCKS) k_connect(WfParams p, int it) {
#if RTB_SMEM_STACK
    __shared__ uint32_t s_stack[kSmemStackDepth * kWfBlock];
#else
    uint32_t *const s_stack = nullptr; // the traversal stack lives in local memory (measured faster)
#endif
    const GeomView<float> &g = p.geom;
    Counters &C = p.ctr[it % 3];
    const uint32_t n = C.n_shadow.v[0];
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&p.glob->rays_shadow, (unsigned long long)n);
    uint64_t nodes = 0, tests = 0;
    uint32_t base = warp_fetch(&C.head_shadow.v[0]);
    while (base < n) {
        uint32_t next_base = 0;
        if (lane_id() == 0)
            next_base = atomicAdd(&C.head_shadow.v[0], 32u);
        const uint32_t idx = base + lane_id();
        if (idx < n) {
            const float4 a = __ldcs(p.sh_a + idx), b = __ldcs(p.sh_b + idx), c = __ldcs(p.sh_c + idx);
            // media on a shadow ray draw from a stream keyed by the queue entry
            Pcg rg = pcg_seed((uint64_t(__float_as_uint(a.x)) << 32) ^ __float_as_uint(b.y), p.seed ^ idx);
            if (shadow_visible<COUNT, false, MEDIA, INST>(g, true, V3<float>(a.x, a.y, a.z), V3<float>(b.x, b.y, b.z), a.w,
                                      __float_as_uint(c.w), rg, nodes, tests, s_stack + threadIdx.x))
                accum_add(p.accum, __float_as_uint(b.w), V3<float>(c.x, c.y, c.z));
        }
        base = __shfl_sync(kFullMask, next_base, 0);
    }
    if (COUNT) {
        const unsigned long long x = warp_sum(nodes), y = warp_sum(tests);
        if (lane_id() == 0) {
            atomicAdd(&p.glob->nodes_visited, x);
            atomicAdd(&p.glob->prim_tests, y);
        }
    }
}

// ---- (A') the warp-scheduled 4-wide traversal (rtb_trace.cuh) behind extend and connect ------------
// Same queues and the same stage contract as k_extend / k_connect above, which remain selectable
// (RTB_OPT_BINARY_TRAVERSAL) for A/B measurements.  A warp reserves a WINDOW of queue entries,
// starts camera samples in its empty entries (one global atomic per window), sorts the window by
// direction octant and walks it with lanes refilled as they finish; when the window's last ray is
// done the warp pushes all its hits into the material queues with ONE multi-lane atomic.
#ifndef RTB_TRACE_MIN_BLOCKS
#define RTB_TRACE_MIN_BLOCKS 3 // resident CTAs of 256 threads per SM the traversal kernels are compiled for
#endif
#ifndef RTB_TRACE_MIN_BLOCKS_LEAN
#define RTB_TRACE_MIN_BLOCKS_LEAN 4 // ... for scenes without media and instances (64 registers, no spills; C5 69.1 -> 67.0 ms)
#endif
#ifndef RTB_TRACE_TOP
#define RTB_TRACE_TOP 0 // stage the top levels of the tree in shared memory (measured: no gain)
#endif
#ifndef RTB_TRACE_BLOCK
#define RTB_TRACE_BLOCK 256
#endif
constexpr int kTraceBlock = RTB_TRACE_BLOCK;
constexpr bool kTraceTop = RTB_TRACE_TOP != 0;

// Cooperative copy of the first nodes of the (breadth-first) tree; returns how many were staged.
__device__ __forceinline__ uint32_t stage_top_nodes(const WideView &w, Vec4f *s_top) {
    const uint32_t n_top = kTraceTop ? (w.n_nodes < uint32_t(kTopNodesMax) ? w.n_nodes : uint32_t(kTopNodesMax)) : 0u;
    for (uint32_t i = threadIdx.x; i < n_top * kNodeRows; i += blockDim.x)
        s_top[(i / kNodeRows) * kTopStride + (i % kNodeRows)] = __ldg(w.nodes + i);
    return n_top;
}

__device__ __forceinline__ uint32_t total_warps() { return gridDim.x * (blockDim.x >> 5); }

template <bool MEDIA> struct ExtendJob {
    const WfParams &p;
    Counters &C;
    const int buf;
    const uint32_t n;   // entries of this iteration's extend queue
    const uint32_t warps;
    uint32_t win = 0;   // size of the window being walked
    uint32_t base = 0, cnt = 0, off8 = 0;
    unsigned long long sample_base = 0;
    uint32_t n_new = 0; // camera samples this warp started (kept by lane 0)
    __device__ ExtendJob(const WfParams &p_, Counters &C_, int buf_, uint32_t n_)
        : p(p_), C(C_), buf(buf_), n(n_), warps(total_warps()) {}

    __device__ __forceinline__ bool next_window(TraceWarpSmem &s, uint32_t &count) {
        const uint32_t lane = lane_id();
        win = guided_window(n, warps, base, win);
        uint32_t b = 0;
        if (lane == 0)
            b = atomicAdd(&C.head_ext.v[0], win);
        b = __shfl_sync(kFullMask, b, 0);
        if (b >= n)
            return false;
        base = b;
        cnt = n - b < win ? n - b : win;
        // sort keys: direction octant of a live path, 8 for an empty entry (a camera sample starts there).
        // Loads only in this loop: they overlap.
        const float4 *qc = p.ext_c[buf] + b, *qb = p.ext_b[buf] + b;
#pragma unroll 4
        for (uint32_t j = lane; j < cnt; j += 32u) {
            const uint32_t pix = __float_as_uint(reinterpret_cast<const float *>(qc + j)[3]);
            const float4 d = qb[j];
            s.key[j] = uint8_t(pix == kInvalidPix ? kFreshKey : (RTB_TRACE_SORT ? octant_of(d.x, d.y, d.z) : 0u));
        }
        __syncwarp();
        window_hist(s, cnt, 9u);
        // camera samples for the empty entries: ONE global atomic per window
        const uint32_t n_empty = s.hist[kFreshKey];
        unsigned long long sb = 0;
        if (n_empty && lane == 0)
            sb = atomicAdd(&p.glob->next_sample, (unsigned long long)n_empty);
        sb = __shfl_sync(kFullMask, sb, 0);
        uint32_t avail = 0;
        if (n_empty && sb < p.total_samples)
            avail = p.total_samples - sb < n_empty ? uint32_t(p.total_samples - sb) : n_empty;
        sample_base = sb;
        if (lane == 0)
            n_new += avail;
        count = window_place(s, cnt, 9u, kFreshKey, avail);
        off8 = s.hist[kFreshKey]; // window positions [off8, count) are the camera samples
        return true;
    }
    __device__ __forceinline__ void prefetch(TraceWarpSmem &s, uint32_t from, uint32_t to) {
        const uint32_t k = from + lane_id();
        if (k < to && k < off8) {
            const uint32_t idx = base + s.perm[k], slot = k % kDeck;
            async_copy<16>(&s.deck_a[slot], p.ext_a[buf] + idx);
            async_copy<16>(&s.deck_b[slot], p.ext_b[buf] + idx);
            if (MEDIA)
                async_copy<8>(&s.deck_c[slot], p.ext_d[buf] + idx);
        }
    }
    __device__ __forceinline__ void fetch(TraceWarpSmem &s, uint32_t k, TravLane &L, V3<float> &o, V3<float> &d,
                                          uint32_t &tag) {
        tag = s.perm[k];
        if (k >= off8) { // a camera sample starts in this (empty) entry: renderer.h:72-75
            const uint32_t idx = base + tag;
            PathState st;
            new_path(p, sample_base + (k - off8), st);
            p.ext_a[buf][idx] = pack_a(st);
            p.ext_b[buf][idx] = pack_b(st);
            p.ext_c[buf][idx] = pack_c(st);
            p.ext_d[buf][idx] = pack_d(st);
            o = st.o;
            d = st.d;
            L.time = st.time;
            L.origin = kNoPrim;
            if (MEDIA)
                L.rng = st.rng;
        } else {
            const uint32_t slot = k % kDeck;
            const float4 a = s.deck_a[slot], b = s.deck_b[slot];
            o = V3<float>(a.x, a.y, a.z);
            d = V3<float>(b.x, b.y, b.z);
            L.time = a.w;
            L.origin = __float_as_uint(b.w);
            if (MEDIA) { // constant_medium::hit draws from the path's stream
                const uint2 x = s.deck_c[slot];
                L.rng.s = uint64_t(x.x) | (uint64_t(x.y) << 32);
            }
        }
        L.t_min = 0.001f;
        L.t_max = Consts<float>::inf();
    }
    __device__ __forceinline__ void world_ray(uint32_t tag, V3<float> &o, V3<float> &d) {
        const float4 a = p.ext_a[buf][base + tag], b = p.ext_b[buf][base + tag];
        o = V3<float>(a.x, a.y, a.z);
        d = V3<float>(b.x, b.y, b.z);
    }
    __device__ __forceinline__ void commit(TraceWarpSmem &s, uint32_t tag, const TravLane &L) {
        const uint32_t idx = base + tag;
        p.ext_e[buf][idx] = make_float2(L.t_max, __uint_as_float(L.best));
        s.key[tag] = uint8_t(L.best == kNoPrim ? uint32_t(kMissKey) : L.best_key);
        if (MEDIA)
            *reinterpret_cast<uint2 *>(p.ext_d[buf] + idx) = make_uint2(uint32_t(L.rng.s), uint32_t(L.rng.s >> 32));
    }
    // Every ray of the window is traced: move the paths, with their hits, into the material queues.
    __device__ __forceinline__ void finish_window(TraceWarpSmem &s) {
        const uint32_t lane = lane_id();
        __syncwarp();
        window_hist(s, cnt, uint32_t(kKeys));
        // the window's hits enter the 8 queues with ONE atomic instruction (their counters share a
        // 128-byte line: L2 sees one transaction)
        uint32_t gbase = 0;
        if (lane < uint32_t(kKeys)) {
            const uint32_t h = s.hist[lane];
            gbase = h ? atomicAdd(&C.key.v[lane], h) : 0u;
        }
        const uint32_t total = window_place(s, cnt, uint32_t(kKeys), kSkipKey, 0u); // perm: entries by hit key
        if (lane < uint32_t(kKeys))
            s.cur[lane] = gbase - s.hist[lane]; // queue position of window position q with key k: cur[k] + q
        __syncwarp();
        // copies only in this loop (no collective): the loads of consecutive trips overlap
#pragma unroll 2
        for (uint32_t q = lane; q < total; q += 32u) {
            const uint32_t j = s.perm[q], key = s.key[j];
            const uint32_t pos = s.cur[key] + q;
            if (pos >= p.cap) { // cannot happen while the capacity bound of WavefrontPool::ensure holds
                atomicExch(&p.glob->overflow, 1ull);
                continue;
            }
            const uint32_t idx = base + j;
            const size_t o = size_t(key) * p.cap + pos;
            const float4 a = __ldcs(p.ext_a[buf] + idx), b = __ldcs(p.ext_b[buf] + idx), c = __ldcs(p.ext_c[buf] + idx);
            const uint4 d = __ldcs(p.ext_d[buf] + idx);
            const float2 e = __ldcs(p.ext_e[buf] + idx);
            __stcs(p.hit_a + o, a);
            __stcs(p.hit_b + o, b);
            __stcs(p.hit_c + o, c);
            __stcs(p.hit_d + o, d);
            __stcs(p.hit_e + o, e);
        }
        __syncwarp();
    }
};

template <bool COUNT, bool MEDIA, bool INST>
__global__ void __launch_bounds__(kTraceBlock, (MEDIA || INST) ? RTB_TRACE_MIN_BLOCKS : RTB_TRACE_MIN_BLOCKS_LEAN) k_extend_w(WfParams p, int it) {
    __shared__ Vec4f s_top[kTraceTop ? kTopNodesMax * kTopStride : 1];
    __shared__ TraceWarpSmem s_warp[kTraceBlock / 32];
    const uint32_t n_top = stage_top_nodes(p.wide, s_top);
    Counters &C = p.ctr[it % 3];
    if (blockIdx.x == 0)
        for (uint32_t i = threadIdx.x; i < sizeof(Counters) / 4; i += blockDim.x)
            reinterpret_cast<uint32_t *>(&p.ctr[(it + 2) % 3])[i] = 0;
    __syncthreads();
    ExtendJob<MEDIA> job(p, C, it & 1, C.n_ext.v[0]);
    uint64_t counters[3] = {0, 0, 0};
    uint32_t overflow = 0;
    warp_trace<ExtendJob<MEDIA>, false, MEDIA, INST, kTraceTop>(p.geom, p.wide, s_top, n_top, s_warp[threadIdx.x >> 5], job,
                                                               counters, overflow);
    if (overflow)
        atomicExch(&p.glob->overflow, 16ull + overflow);
    if (lane_id() == 0 && job.n_new)
        atomicAdd(&p.glob->paths, (unsigned long long)job.n_new);
    if (COUNT) {
        const unsigned long long x = warp_sum(counters[0]), y = warp_sum(counters[1]);
        if (lane_id() == 0) {
            atomicAdd(&p.glob->nodes_visited, x);
            atomicAdd(&p.glob->prim_tests, y);
            atomicAdd(&p.glob->extend_nodes, x);
        }
        atomicMax(&p.glob->max_nodes_per_ray, (unsigned long long)counters[2]);
    }
}

template <bool MEDIA> struct ConnectJob {
    const WfParams &p;
    Counters &C;
    const uint32_t n;
    const uint32_t warps;
    uint32_t win = 0;
    uint32_t base = 0;
    uint32_t pix = 0; // per lane: the pixel of the ray it walks
    __device__ ConnectJob(const WfParams &p_, Counters &C_, uint32_t n_)
        : p(p_), C(C_), n(n_), warps(total_warps()) {}
    __device__ __forceinline__ bool next_window(TraceWarpSmem &s, uint32_t &count) {
        const uint32_t lane = lane_id();
        win = guided_window(n, warps, base, win);
        uint32_t b = 0;
        if (lane == 0)
            b = atomicAdd(&C.head_shadow.v[0], win);
        b = __shfl_sync(kFullMask, b, 0);
        if (b >= n)
            return false;
        base = b;
        const uint32_t cnt = n - b < win ? n - b : win;
#pragma unroll 4
        for (uint32_t j = lane; j < cnt; j += 32u) {
            const float4 d = p.sh_b[b + j];
            s.key[j] = uint8_t(RTB_TRACE_SORT ? octant_of(d.x, d.y, d.z) : 0u);
        }
        __syncwarp();
        window_hist(s, cnt, 8u);
        count = window_place(s, cnt, 8u, kSkipKey, 0u);
        return true;
    }
    __device__ __forceinline__ void prefetch(TraceWarpSmem &s, uint32_t from, uint32_t to) {
        const uint32_t k = from + lane_id();
        if (k < to) {
            const uint32_t idx = base + s.perm[k], slot = k % kDeck;
            async_copy<16>(&s.deck_a[slot], p.sh_a + idx);
            async_copy<16>(&s.deck_b[slot], p.sh_b + idx);
            async_copy<4>(&s.deck_c[slot], reinterpret_cast<const float *>(p.sh_c + idx) + 3);
        }
    }
    __device__ __forceinline__ void fetch(TraceWarpSmem &s, uint32_t k, TravLane &L, V3<float> &o, V3<float> &d,
                                          uint32_t &tag) {
        const uint32_t idx = base + s.perm[k], slot = k % kDeck;
        tag = idx;
        const float4 a = s.deck_a[slot], b = s.deck_b[slot];
        o = V3<float>(a.x, a.y, a.z);
        d = V3<float>(b.x, b.y, b.z);
        pix = __float_as_uint(b.w);
        L.time = 0.0f; // shadow rays carry time 0 (direct_light_integrator.h:115)
        L.origin = s.deck_c[slot].x;
        // t_min is 0.001 along the UNIT direction; the stored direction may be the unnormalised segment
        const float len = isfinite(a.w) ? length(d) : 1.0f;
        L.t_min = 0.001f / len;
        L.t_max = a.w;
        if (MEDIA) // media on a shadow ray draw from a stream keyed by the queue entry
            L.rng = pcg_seed((uint64_t(__float_as_uint(a.x)) << 32) ^ __float_as_uint(b.y), p.seed ^ idx);
    }
    __device__ __forceinline__ void world_ray(uint32_t tag, V3<float> &o, V3<float> &d) {
        const float4 a = p.sh_a[tag], b = p.sh_b[tag];
        o = V3<float>(a.x, a.y, a.z);
        d = V3<float>(b.x, b.y, b.z);
    }
    __device__ __forceinline__ void commit(TraceWarpSmem &, uint32_t tag, const TravLane &L) {
        if (L.best == kNoPrim) { // unoccluded: the (already weighted) contribution counts
            const float4 c = __ldcs(p.sh_c + tag);
            accum_add(p.accum, pix, V3<float>(c.x, c.y, c.z));
        }
    }
    __device__ __forceinline__ void finish_window(TraceWarpSmem &) {}
};

template <bool COUNT, bool MEDIA, bool INST>
__global__ void __launch_bounds__(kTraceBlock, (MEDIA || INST) ? RTB_TRACE_MIN_BLOCKS : RTB_TRACE_MIN_BLOCKS_LEAN) k_connect_w(WfParams p, int it) {
    __shared__ Vec4f s_top[kTraceTop ? kTopNodesMax * kTopStride : 1];
    __shared__ TraceWarpSmem s_warp[kTraceBlock / 32];
    const uint32_t n_top = stage_top_nodes(p.wide, s_top);
    Counters &C = p.ctr[it % 3];
    const uint32_t n = C.n_shadow.v[0];
    if (blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd(&p.glob->rays_shadow, (unsigned long long)n);
    __syncthreads();
    ConnectJob<MEDIA> job(p, C, n);
    uint64_t counters[3] = {0, 0, 0};
    uint32_t overflow = 0;
    warp_trace<ConnectJob<MEDIA>, true, MEDIA, INST, kTraceTop>(p.geom, p.wide, s_top, n_top, s_warp[threadIdx.x >> 5], job,
                                                               counters, overflow);
    if (overflow)
        atomicExch(&p.glob->overflow, 32ull + overflow);
    if (COUNT) {
        const unsigned long long x = warp_sum(counters[0]), y = warp_sum(counters[1]);
        if (lane_id() == 0) {
            atomicAdd(&p.glob->nodes_visited, x);
            atomicAdd(&p.glob->prim_tests, y);
        }
    }
}

// ---- rtb_trace_batch precision 34 / 36: caller rays through the same scheduler ----------------------
__global__ void k_batch_rays_in(const rtb_ray *__restrict__ rays, uint32_t n, const int32_t *__restrict__ orig_to_sorted,
                                int n_orig, float4 *a, float4 *b, float2 *t) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const rtb_ray q = rays[i];
        uint32_t origin = kNoPrim;
        if (q.origin_prim >= 0 && q.origin_prim < n_orig)
            origin = uint32_t(orig_to_sorted[q.origin_prim]);
        a[i] = make_float4(float(q.o[0]), float(q.o[1]), float(q.o[2]), float(q.time));
        b[i] = make_float4(float(q.d[0]), float(q.d[1]), float(q.d[2]), __uint_as_float(origin));
        t[i] = make_float2(float(q.t_min), float(q.t_max));
    }
}
template <bool ANY>
__global__ void __launch_bounds__(kTraceBlock, RTB_TRACE_MIN_BLOCKS)
    k_trace_batch_w(GeomView<float> g, WideView w, BatchTraceJob job, unsigned long long *visits, uint32_t *flag) {
    __shared__ Vec4f s_top[kTraceTop ? kTopNodesMax * kTopStride : 1];
    __shared__ TraceWarpSmem s_warp[kTraceBlock / 32];
    const uint32_t n_top = stage_top_nodes(w, s_top);
    __syncthreads();
    uint64_t counters[3] = {0, 0, 0};
    uint32_t overflow = 0;
    warp_trace<BatchTraceJob, ANY, true, true, kTraceTop>(g, w, s_top, n_top, s_warp[threadIdx.x >> 5], job, counters, overflow);
    if (overflow)
        atomicExch(flag, overflow);
    if (visits) {
        const unsigned long long x = warp_sum(counters[0]), y = warp_sum(counters[1]);
        if (lane_id() == 0) {
            atomicAdd(&visits[0], x);
            atomicAdd(&visits[1], y);
        }
    }
}
__global__ void k_batch_hits_out(GeomView<float> g, const rtb_ray *__restrict__ rays, const float2 *__restrict__ res, uint32_t n,
                                 rtb_hit *__restrict__ hits) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const rtb_ray q = rays[i];
        const uint32_t pi = __float_as_uint(res[i].y);
        rtb_hit h;
        std::memset(&h, 0, sizeof(h));
        h.prim = -1;
        h.material = -1;
        if (pi != kNoPrim) {
            const V3<float> o{float(q.o[0]), float(q.o[1]), float(q.o[2])}, d{float(q.d[0]), float(q.d[1]), float(q.d[2])};
            const RecT<float> rec = make_record<float, true, true>(g, pi, o, d, float(q.time), res[i].x);
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
}

// ---- (B) fused kernel for shared-memory-resident scenes ----------------------------------------

// The set-up one thread runs once per block, kept out of the body of the small kernels (its 1,700 inlined
// instructions cost them registers around the main loop: C1 8.28 -> 8.06 ms, C3 17.17 -> 16.60 ms out of line; the
// general kernel measured 2 % slower that way and keeps the inlined copy).
__device__ __noinline__ void build_flat_fast_outlined(const GeomView<float> &g, FlatFast &ff) { build_flat_fast(g, ff); }

constexpr uint32_t kSampleChunk = 256; // samples a warp reserves per atomic
#ifndef RTB_FUSED_DECK
#define RTB_FUSED_DECK 1 // camera samples are generated 32 at a time by the whole warp into a shared-memory deck
#endif
struct FusedDeck { // one per warp: origin + time, direction + pixel, generator state after the camera's draws
    float4 a[32], b[32];
    uint2 r[32];
};
#ifndef RTB_FUSED_DYNAMIC_SHADE
#define RTB_FUSED_DYNAMIC_SHADE 1 // the general fused kernel shades through ONE copy of shade_surface (see there)
#endif

// SIMPLE >= 1: the scene only uses lambertian and diffuse_light materials (the Cornell boxes): the
// other four shade_surface instantiations are left out, which halves the kernel's code size
// (the profile of the general kernel showed instruction-cache misses).  SIMPLE == 2: in addition
// every primitive is an axis-aligned rect and every texture a solid colour: hit records come from
// the per-primitive plane digest only (ALL_PLANAR), the traversal holds the rect and box loops
// only (RECTS_ONLY) and no texture code is compiled in.
template <bool OLD, bool COUNT, int SIMPLE>
__global__ void __launch_bounds__(128, OLD ? (SIMPLE == 2 ? RTB_FUSED_MIN_BLOCKS_OLD : RTB_FUSED_MIN_BLOCKS_OLD_GENERAL)
                           : (SIMPLE == 2 ? RTB_FUSED_MIN_BLOCKS_NEW_PLANAR : (SIMPLE == 1 ? RTB_FUSED_MIN_BLOCKS_NEW : RTB_FUSED_MIN_BLOCKS_NEW_GENERAL)))
    k_fused(WfParams p) {
    __shared__ FlatSmem sm;
    __shared__ FlatFast ff;
    if (!OLD)
        stage_lights(p);
    const GeomView<float> g = stage_scene_flat(p.geom, sm);
    if (threadIdx.x == 0) {
        if (SIMPLE)
            build_flat_fast_outlined(g, ff);
        else
            build_flat_fast(g, ff);
    }
    if (int(threadIdx.x) < g.n_prims) {
        PlaneRec &pl = ff.plane[threadIdx.x];
        build_plane_rec(g, threadIdx.x, pl);
        if ((g.prims[threadIdx.x].type_mat & PT_TYPE_MASK) != PT_INSTANCE) {
            const MatT<float> m = p.shade.mats[g.prims[threadIdx.x].type_mat >> PT_MAT_SHIFT];
            pl.color[0] = m.color[0];
            pl.color[1] = m.color[1];
            pl.color[2] = m.color[2];
            pl.mat_type = m.type;
        }
    }
    __syncthreads();
    PathState s;
    bool alive = false, exhausted = false;
    unsigned long long chunk_next = 0, chunk_end = 0; // warp-uniform: this warp's private sample range
    uint32_t n_closest = 0, n_shadow = 0, n_paths = 0; // per thread and launch (a launch is <= 2^28 samples)
    uint64_t nodes = 0, tests = 0;
    // The deck pays where the kernel is small (C1 9.03 -> 8.31 ms, C3 17.81 -> 17.18 ms); the general kernel, which
    // already waits for instructions and spills, loses with it (C4-env 40.5 -> 46.2 ms) and keeps the lane-by-lane form.
    constexpr bool DECK = RTB_FUSED_DECK && SIMPLE != 0;
    __shared__ FusedDeck decks[DECK ? 4 : 1];
    FusedDeck &dk = decks[DECK ? (threadIdx.x >> 5) : 0];
    uint32_t deck_n = 0;   // warp-uniform: camera samples waiting in the deck
    bool gen_done = false; // warp-uniform: the launch has no samples left for this warp
    while (true) {
        // regeneration: every idle lane takes the next sample of the warp's chunk
        const bool need = !alive && !exhausted;
        const uint32_t m = __ballot_sync(kFullMask, need);
        if (DECK && m) {
            uint32_t want = __popc(m), r = __popc(m & ((1u << lane_id()) - 1u));
            bool pending = need;
            for (;;) {
                const uint32_t take = want < deck_n ? want : deck_n;
                if (pending) {
                    if (r < take) { // entry deck_n - 1 - r
                        const uint32_t e = deck_n - 1u - r;
                        const float4 a = dk.a[e], b = dk.b[e];
                        const uint2 g2 = dk.r[e];
                        s.o = V3<float>(a.x, a.y, a.z);
                        s.time = a.w;
                        s.d = V3<float>(b.x, b.y, b.z);
                        s.pix = __float_as_uint(b.w);
                        s.rng.s = uint64_t(g2.x) | (uint64_t(g2.y) << 32);
                        s.T = V3<float>(1, 1, 1);
                        s.depth = 0;
                        s.spec = false;
                        s.prev_pdf = 0.f;
                        s.origin_prim = kNoPrim;
                        alive = true;
                        pending = false;
                        ++n_paths;
                    } else {
                        r -= take;
                    }
                }
                deck_n -= take;
                want -= take;
                if (want == 0 || gen_done)
                    break;
                // the deck is empty: the WHOLE warp generates the next 32 camera samples (at 32 lanes; lane by
                // lane regeneration ran at 10 of 32 and was 17 % of the kernel's issued instructions)
                if (chunk_next == chunk_end) {
                    unsigned long long base = 0;
                    if (lane_id() == 0)
                        base = atomicAdd(&p.glob->next_sample, (unsigned long long)kSampleChunk);
                    chunk_next = __shfl_sync(kFullMask, base, 0);
                    chunk_end = chunk_next + kSampleChunk;
                }
                const unsigned long long id0 = chunk_next;
                chunk_next += 32u;
                if (id0 >= p.window_end) {
                    gen_done = true;
                    break;
                }
                const unsigned long long left = p.window_end - id0;
                deck_n = left < 32ull ? uint32_t(left) : 32u;
                __syncwarp();
                {
                    PathState q;
                    new_path(p, id0 + lane_id(), q); // (lanes past the end compute a sample nobody pops)
                    dk.a[lane_id()] = make_float4(q.o.x, q.o.y, q.o.z, q.time);
                    dk.b[lane_id()] = make_float4(q.d.x, q.d.y, q.d.z, __uint_as_float(q.pix));
                    dk.r[lane_id()] = make_uint2(uint32_t(q.rng.s), uint32_t(q.rng.s >> 32));
                }
                __syncwarp();
            }
            if (pending)
                exhausted = true;
        }
        if (!DECK && m) {
            const uint32_t n = __popc(m), rank = __popc(m & ((1u << lane_id()) - 1u));
            const unsigned long long avail = chunk_end - chunk_next;
            unsigned long long id;
            if (avail >= n) {
                id = chunk_next + rank;
                chunk_next += n;
            } else {
                unsigned long long base = 0;
                if (lane_id() == 0)
                    base = atomicAdd(&p.glob->next_sample, (unsigned long long)kSampleChunk);
                base = __shfl_sync(kFullMask, base, 0);
                id = rank < avail ? chunk_next + rank : base + (rank - avail);
                chunk_next = base + (n - avail);
                chunk_end = base + kSampleChunk;
            }
            if (need) {
                if (id < p.window_end) {
                    new_path(p, id, s);
                    alive = true;
                    ++n_paths;
                } else {
                    exhausted = true;
                }
            }
        }
        if (!__any_sync(kFullMask, alive))
            break;
        if (alive) {
            PathDraw draw{&s.rng};
            float t;
            const uint32_t pi = traverse_flat_fast<false, COUNT, SIMPLE == 2>(g, ff, s.o, s.d, s.time, 0.001f, Consts<float>::inf(),
                                                                 s.origin_prim, draw, t, nodes, tests);
            ++n_closest;
            if (pi == kNoPrim) {
                miss_surface<OLD>(p, s);
                alive = false;
            } else {
                ShadowReq sh;
                const uint32_t pix = s.pix;
                const int mtype = SIMPLE == 2 ? ff.plane[pi].mat_type : p.shade.mats[g.prims[pi].type_mat >> PT_MAT_SHIFT].type;
                if (SIMPLE) {
                    if (mtype == 0)
                        shade_surface<0, OLD, SIMPLE == 2>(p, g, s, t, pi, alive, sh, ff.plane);
                    else
                        shade_surface<3, OLD, SIMPLE == 2>(p, g, s, t, pi, alive, sh, ff.plane);
                } else {
#if RTB_FUSED_DYNAMIC_SHADE
                    shade_surface<-1, OLD>(p, g, s, t, pi, alive, sh, ff.plane); // one copy, the record's own type
#else
                    switch (mtype) {
                    case 0: shade_surface<0, OLD>(p, g, s, t, pi, alive, sh, ff.plane); break;
                    case 1: shade_surface<1, OLD>(p, g, s, t, pi, alive, sh, ff.plane); break;
                    case 2: shade_surface<2, OLD>(p, g, s, t, pi, alive, sh, ff.plane); break;
                    case 3: shade_surface<3, OLD>(p, g, s, t, pi, alive, sh, ff.plane); break;
                    case 4: shade_surface<4, OLD>(p, g, s, t, pi, alive, sh, ff.plane); break;
                    default: shade_surface<5, OLD>(p, g, s, t, pi, alive, sh, ff.plane); break;
                    }
#endif
                }
                if (!OLD && sh.want) {
                    ++n_shadow;
                    Pcg rg = s.rng; // a copy: the shadow test must not advance the path's stream
                    // t_min is 0.001 along the UNIT direction; sh.d may be the unnormalised segment
                    const float len = isfinite(sh.tmax) ? length(sh.d) : 1.0f;
                    PathDraw sdraw{&rg};
                    float st;
                    if (traverse_flat_fast<true, COUNT, SIMPLE == 2>(g, ff, sh.o, sh.d, 0.0f, 0.001f / len, sh.tmax, sh.origin, sdraw,
                                                        st, nodes, tests) == kNoPrim)
                        accum_add(p.accum, pix, sh.c);
                }
            }
        }
    }
    const unsigned long long a = warp_sum(n_closest), b = warp_sum(n_shadow), c = warp_sum(n_paths);
    if (lane_id() == 0) {
        atomicAdd(&p.glob->rays_closest, a);
        atomicAdd(&p.glob->rays_shadow, b);
        atomicAdd(&p.glob->paths, c);
    }
    if (COUNT) {
        const unsigned long long e = warp_sum(nodes), f = warp_sum(tests);
        if (lane_id() == 0) {
            atomicAdd(&p.glob->nodes_visited, e);
            atomicAdd(&p.glob->prim_tests, f);
        }
    }
}

__global__ void k_set_next_sample(Globals *glob, unsigned long long v) { glob->next_sample = v; }

// Parity layer 1 for the fused kernel's own traversal: the rays of rtb_trace_batch through
// traverse_flat_fast (typed lists, box slab tests), records through the same make_record.
struct FlatDraw {
    Pcg *g;
    __device__ float operator()() { return g->next_open(); }
};
__global__ void __launch_bounds__(128) k_trace_fast_batch(GeomView<float> geom, const int32_t *__restrict__ orig_to_sorted,
                                                          int n_orig, const rtb_ray *__restrict__ rays, uint64_t n,
                                                          rtb_hit *__restrict__ hits, unsigned long long *visits,
                                                          bool plane_records) {
    __shared__ FlatSmem sm;
    __shared__ FlatFast ff;
    const GeomView<float> g = stage_scene_flat(geom, sm);
    if (threadIdx.x == 0)
        build_flat_fast(g, ff);
    if (int(threadIdx.x) < g.n_prims)
        build_plane_rec(g, threadIdx.x, ff.plane[threadIdx.x]);
    __syncthreads();
    uint64_t nodes = 0, tests = 0;
    for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) {
        const rtb_ray q = rays[i];
        const V3<float> o{float(q.o[0]), float(q.o[1]), float(q.o[2])}, d{float(q.d[0]), float(q.d[1]), float(q.d[2])};
        uint32_t origin = kNoPrim;
        if (q.origin_prim >= 0 && q.origin_prim < n_orig)
            origin = uint32_t(orig_to_sorted[q.origin_prim]);
        Pcg rg = pcg_seed(i, 0x51ed270b);
        FlatDraw draw{&rg};
        float t;
        const uint32_t pi = traverse_flat_fast<false, true>(g, ff, o, d, float(q.time), float(q.t_min), float(q.t_max),
                                                            origin, draw, t, nodes, tests);
        rtb_hit h;
        std::memset(&h, 0, sizeof(h));
        h.prim = -1;
        h.material = -1;
        if (pi != kNoPrim) {
            // planar primitives: the record the fused kernel shades from (no u,v)
            const RecT<float> rec = plane_records && ff.plane[pi].valid
                                        ? plane_record<float>(ff.plane[pi], o, d, t)
                                        : make_record<float, true, true>(g, pi, o, d, float(q.time), t);
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

// renderer.h:126-140 + render_buffer.h:35-55: sqrt(sum/spp), clamp, (uchar)(x*255), y flip — in the
// reference's own double arithmetic (scale = 1.0 / samples; sqrt(scale * sum)), so that the bytes are the
// ones the reference's write_color_to_buffer + save_to_png produce from the same sums (3 DSQRT per pixel).
__device__ __forceinline__ uint8_t resolve_channel(float sum, double scale) {
    const double v = sqrt(scale * double(sum));
    return uint8_t((v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v)) * 255);
}
__global__ void k_resolve_rgb8(const float4 *__restrict__ accum, int w, int h, double scale,
                               uint8_t *__restrict__ rgb8) {
    const size_t n = size_t(w) * h;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
        const size_t row = i / w, col = i - row * w; // row in the PNG (top first)
        const float4 a = accum[(size_t(h) - 1 - row) * w + col];
        rgb8[3 * i] = resolve_channel(a.x, scale);
        rgb8[3 * i + 1] = resolve_channel(a.y, scale);
        rgb8[3 * i + 2] = resolve_channel(a.z, scale);
    }
}

// queue q holds hits on material type mtype (q == mtype, or kTexturedKey for textured lambertians)
template <bool OLD> void launch_shade(int mtype, int q, const WfParams &P, int it, int grid, cudaStream_t st) {
    switch (mtype) {
    case 0: k_shade<0, OLD><<<grid, 128, 0, st>>>(P, it, q); break;
    case 1: k_shade<1, OLD><<<grid, 128, 0, st>>>(P, it, q); break;
    case 2: k_shade<2, OLD><<<grid, 128, 0, st>>>(P, it, q); break;
    case 3: k_shade<3, OLD><<<grid, 128, 0, st>>>(P, it, q); break;
    case 4: k_shade<4, OLD><<<grid, 128, 0, st>>>(P, it, q); break;
    default: k_shade<5, OLD><<<grid, 128, 0, st>>>(P, it, q); break;
    }
}

template <class K> int blocks_per_sm(K kernel, int threads) {
    int n = 0;
    RTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, 0));
    return n > 0 ? n : 1;
}

} // namespace

// Device memory of the path pool; kept across renders on a context.
struct WavefrontPool {
    uint32_t P = 0;
    uint32_t cap = 0, n_cursor = 0;
    DeviceBuffer ext[2][4], ext_e[2], hit[5], sh_a, sh_b, sh_c, cursor, ctr, glob;
    uint32_t *h_live = nullptr; // pinned: extend-queue length probes
    Globals *h_glob = nullptr;  // pinned
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    std::vector<cudaEvent_t> ev_ext; // pairs bracketing the dominant kernel's launches (RTB_RENDER_TIME_EXTEND)
    cudaEvent_t ext_event(size_t i) {
        while (ev_ext.size() <= i) {
            cudaEvent_t e;
            RTB_CUDA(cudaEventCreate(&e));
            ev_ext.push_back(e);
        }
        return ev_ext[i];
    }
    void ensure_common() {
        if (h_live)
            return;
        RTB_CUDA(cudaMallocHost(&h_live, 64 * sizeof(uint32_t)));
        RTB_CUDA(cudaMallocHost(&h_glob, sizeof(Globals)));
        for (auto &e : ev)
            RTB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        RTB_CUDA(cudaEventCreate(&ev_begin));
        RTB_CUDA(cudaEventCreate(&ev_end));
        ctr.alloc(3 * sizeof(Counters));
        glob.alloc(sizeof(Globals));
    }
    // want = resident paths, warps = warps of one k_extend launch (each owns a private sample range)
    void ensure(uint32_t want, uint32_t warps) {
        ensure_common();
        if (want <= P && warps == n_cursor) // grow-only: reallocating ~2 GB costs tens of milliseconds
            return;
        P = 0;
        // the final drain of the private sample ranges can add up to warps * kWfChunk paths
        const size_t n = size_t(want) + size_t(warps) * kWfChunk + 32;
        for (auto &q : ext)
            for (auto &arr : q)
                arr.alloc(n * 16);
        for (auto &arr : ext_e)
            arr.alloc(n * 8);
        for (int k = 0; k < 4; ++k)
            hit[k].alloc(n * 16 * kKeys);
        hit[4].alloc(n * 8 * kKeys);
        sh_a.alloc(n * 16);
        sh_b.alloc(n * 16);
        sh_c.alloc(n * 16);
        cursor.alloc(size_t(warps) * 16);
        P = want;
        cap = uint32_t(n);
        n_cursor = warps;
    }
    ~WavefrontPool() {
        if (h_live)
            cudaFreeHost(h_live);
        if (h_glob)
            cudaFreeHost(h_glob);
        for (auto &e : ev)
            if (e)
                cudaEventDestroy(e);
        if (ev_begin)
            cudaEventDestroy(ev_begin);
        if (ev_end)
            cudaEventDestroy(ev_end);
        for (auto &e : ev_ext)
            cudaEventDestroy(e);
    }
};

void wavefront_release(rtb_context *ctx) {
    delete ctx->pool;
    ctx->pool = nullptr;
}

void wavefront_render(rtb_context *ctx, const rtb_render_params &rp, float4 *d_accum, cudaStream_t st,
                      rtb_render_stats *stats) {
    const DeviceScene &sc = *ctx->scene;
    if (!ctx->pool)
        ctx->pool = new WavefrontPool();
    WavefrontPool &pool = *ctx->pool;

    const uint32_t npix_image = uint32_t(rp.width) * uint32_t(rp.height);
    const int row_stride = rp.row_stride > 0 ? rp.row_stride : 1;
    const int my_rows = rp.row_offset < rp.height ? (rp.height - rp.row_offset + row_stride - 1) / row_stride : 0;
    const uint32_t npix = uint32_t(rp.width) * uint32_t(my_rows);
    const int stride = rp.sample_stride > 0 ? rp.sample_stride : 1;
    const int offset = rp.sample_offset;
    const int local_spp = offset < rp.spp ? (rp.spp - offset + stride - 1) / stride : 0;
    const unsigned long long n_samples = (unsigned long long)npix * (unsigned long long)local_spp;
    // max_depth 0: every Li() returns black without tracing anything (the depth loop of
    // e.g. rr_path_integrator.h:27 never runs)
    const unsigned long long total = rp.max_depth > 0 ? n_samples : 0ull;

    GeomView<float> geom = sc.geom<float>();
    if (ctx->opt_flat == 0)
        geom.flat = 0;

    WfParams W;
    std::memset(&W, 0, sizeof(W));
    W.geom = geom;
    W.wide = sc.wide();
    W.shade = sc.shade<float>();
    W.cam = sc.host.f32.camera;
    pool.ensure_common();
    const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    const int wf_grid = sms * 8; // 8 resident CTAs of 128 threads per SM
    W.ctr = pool.ctr.as<Counters>();
    W.glob = pool.glob.as<Globals>();
    W.accum = d_accum;
    W.width = rp.width;
    W.height = rp.height;
    W.spp = rp.spp;
    W.max_depth = rp.max_depth;
    W.rr_start = rp.rr_start_depth;
    W.integrator = rp.integrator;
    W.sample_offset = offset;
    W.sample_stride = stride;
    W.row_offset = rp.row_offset;
    W.row_stride = row_stride;
    W.npix = npix;
    W.inv_npix = recip32(npix);
    W.inv_width = recip32(uint32_t(rp.width));
    W.total_samples = total;
    W.window_end = total;
    W.seed = rp.seed;
    W.seed_mixed = mix64(rp.seed);
    for (int k = 0; k < 3; ++k)
        W.bg[k] = float(sc.host.globals.background[k]);
    W.mat_mask = sc.host.mat_type_mask;
    W.has_media = sc.host.has_media ? 1 : 0;

    const bool old_api = rp.integrator <= RTB_INTEGRATOR_RR;
    const bool nee = rp.integrator >= RTB_INTEGRATOR_DIRECT && !sc.host.f32.lights.empty();
    const bool count = (rp.flags & RTB_RENDER_COUNT_VISITS) != 0;
    const bool time_dom = (rp.flags & RTB_RENDER_TIME_EXTEND) != 0;
    const bool simple = (W.mat_mask & ~((1u << RTB_MAT_LAMBERTIAN) | (1u << RTB_MAT_DIFFUSE_LIGHT))) == 0;
    // every primitive has a plane digest (build_plane_rec) and every material a baked solid colour?
    bool all_planar = simple && sc.host.flat_ok;
    for (size_t i = 0; all_planar && i < sc.host.f32.prims.size(); ++i) {
        const uint32_t type = sc.host.f32.prims[i].type_mat & PT_TYPE_MASK;
        const int chain = sc.host.prim_chain[i];
        all_planar = type == PT_INSTANCE || ((type == PT_XY || type == PT_XZ || type == PT_YZ) &&
                                             (chain < 0 || sc.host.chains[size_t(chain)].count <= kMaxChainOps));
    }
    for (size_t i = 0; all_planar && i < sc.host.f32.mats.size(); ++i)
        all_planar = (sc.host.f32.mats[i].flags & 3) == 2; // solid colour, baked (mat_tex)
    size_t n_ext_events = 0;

    uint64_t launches = 0;
    int it = 0;
    bool cancelled = false;

    // ---- the two schedules, each over the local samples [begin, end) --------------------------
    auto run_fused = [&](unsigned long long begin, unsigned long long end) {
        // One persistent kernel per window of samples (windows keep rtb_cancel responsive).
        // pick the instantiation once: (legacy vs BSDF API) x (counting) x (simple material set)
        void (*kern)(WfParams) = nullptr;
        {
            void (*table[2][2][3])(WfParams) = {
                {{k_fused<false, false, 0>, k_fused<false, false, 1>, k_fused<false, false, 2>},
                 {k_fused<false, true, 0>, k_fused<false, true, 1>, k_fused<false, true, 2>}},
                {{k_fused<true, false, 0>, k_fused<true, false, 1>, k_fused<true, false, 2>},
                 {k_fused<true, true, 0>, k_fused<true, true, 1>, k_fused<true, true, 2>}}};
            kern = table[old_api ? 1 : 0][count ? 1 : 0][all_planar ? 2 : simple ? 1 : 0];
        }
        const int bps = blocks_per_sm(kern, 128);
        const int grid = sms * bps;
        const unsigned long long window = 1ull << 28; // ~30 ms of work on C1: bounds the latency of rtb_cancel
        int in_flight = 0;
        for (unsigned long long b0 = begin; b0 < end && !cancelled; b0 += window) {
            W.total_samples = total;
            W.window_end = std::min(end, b0 + window);
            k_set_next_sample<<<1, 1, 0, st>>>(W.glob, b0);
            if (time_dom)
                RTB_CUDA(cudaEventRecord(pool.ext_event(n_ext_events++), st));
            kern<<<grid, 128, 0, st>>>(W);
            if (time_dom)
                RTB_CUDA(cudaEventRecord(pool.ext_event(n_ext_events++), st));
            launches += 2;
            ++it;
            ++in_flight;
            RTB_CUDA(cudaGetLastError());
            if (b0 + window < end) { // keep at most two windows in flight
                const int slot = in_flight & 1;
                RTB_CUDA(cudaEventRecord(pool.ev[slot], st));
                if (in_flight >= 2)
                    RTB_CUDA(cudaEventSynchronize(pool.ev[slot ^ 1]));
            }
            if (ctx->cancel.load(std::memory_order_relaxed))
                cancelled = true;
        }
    };

    int traversal_used = 0; // rtb_render_stats.traversal
    auto run_wavefront = [&](unsigned long long begin, unsigned long long end, bool first_segment) {
        const unsigned long long seg = end - begin;
        uint32_t P = rp.pool_paths > 0 ? uint32_t(rp.pool_paths) : kDefaultPool;
        P = (P + 31u) & ~31u;
        if ((unsigned long long)P > seg)
            P = uint32_t((seg + 31ull) & ~31ull);
        if (P < 32)
            P = 32;
        // always at least the default pool, so that a small first render (a preview) is not
        // followed by a reallocation inside the next, full-size one
        pool.ensure(std::max(P, kDefaultPool), uint32_t(wf_grid) * 4u);
        for (int b = 0; b < 2; ++b) {
            W.ext_a[b] = pool.ext[b][0].as<float4>();
            W.ext_b[b] = pool.ext[b][1].as<float4>();
            W.ext_c[b] = pool.ext[b][2].as<float4>();
            W.ext_d[b] = pool.ext[b][3].as<uint4>();
            W.ext_e[b] = pool.ext_e[b].as<float2>();
        }
        W.hit_a = pool.hit[0].as<float4>();
        W.hit_b = pool.hit[1].as<float4>();
        W.hit_c = pool.hit[2].as<float4>();
        W.hit_d = pool.hit[3].as<uint4>();
        W.hit_e = pool.hit[4].as<float2>();
        W.sh_a = pool.sh_a.as<float4>();
        W.sh_b = pool.sh_b.as<float4>();
        W.sh_c = pool.sh_c.as<float4>();
        W.cursor = pool.cursor.as<unsigned long long>();
        W.cap = pool.cap;
        W.P = P;
        W.total_samples = end; // k_extend hands out samples [next_sample, total_samples)
        W.window_end = end;
        const int grid = wf_grid;
        if (!first_segment) {
            k_clear<<<1, 256, 0, st>>>(W.ctr, nullptr);
            ++launches;
        }
        k_set_next_sample<<<1, 1, 0, st>>>(W.glob, begin);
        k_init<<<sms * 4, 256, 0, st>>>(W, it, P, pool.n_cursor);
        launches += 2;
        RTB_CUDA(cudaGetLastError());
        // radiance of a ray that leaves the scene is identically zero: k_miss only empties entries
        const bool miss_shades = W.bg[0] != 0.f || W.bg[1] != 0.f || W.bg[2] != 0.f ||
                                 (rp.integrator >= RTB_INTEGRATOR_DIRECT && W.shade.n_infinite_lights > 0);
        const bool media = W.has_media != 0;
        const bool inst = sc.host.n_instances > 0; // instance code is compiled out of the kernels otherwise
        const uint32_t minor_mask = (1u << RTB_MAT_METAL) | (1u << RTB_MAT_DIELECTRIC) | (1u << RTB_MAT_DIFFUSE_LIGHT) |
                                    (1u << RTB_MAT_ISOTROPIC);
        // the warp-scheduled 4-wide traversal, unless round 1's kernels are asked for (or the scene is
        // small enough for the lockstep walk of the shared-memory copy, which those kernels hold)
        // Measured on B200 (profiles/r02_trace_knobs_and_upload_phases.txt, gpurun d_sweep of this round): on the 1 M-sphere field (one primitive type, no
        // wrappers) the 4-wide kernels are 16 % faster in extend; on scene09 (rects, spheres, an instance
        // and two media in one tree) the heterogeneous leaf steps run at 5-10 lanes and round 1's plain
        // kernels are 12 % faster.  opt_binary_traversal: 0 = by that rule, 1 = always binary, 2 = always wide.
        const bool mixed = media || inst;
        const bool wide = !geom.flat && (ctx->opt_binary_traversal == 2 || (ctx->opt_binary_traversal == 0 && !mixed));
        traversal_used = geom.flat ? 0 : (wide ? 2 : 1);
        int trace_grid = 0, connect_grid = 0;
        constexpr int kBatch = 4; // iterations between two host-side liveness probes
        int probe = 0, zero_probes = 0;
        const int it0 = it;
        bool done = seg == 0;
        int pending[2] = {-1, -1}; // probe slots in flight
        while (!done) {
            for (int b = 0; b < kBatch; ++b, ++it) {
                // with time_dom: 5 events per iteration bracket extend | shade | miss | connect
                auto mark = [&]() {
                    if (time_dom)
                        RTB_CUDA(cudaEventRecord(pool.ext_event(n_ext_events++), st));
                };
                mark();
                if (wide) {
                    typedef void (*ExtendKernel)(WfParams, int);
                    static const ExtendKernel table[2][2][2] = {
                        {{k_extend_w<false, false, false>, k_extend_w<false, false, true>},
                         {k_extend_w<false, true, false>, k_extend_w<false, true, true>}},
                        {{k_extend_w<true, false, false>, k_extend_w<true, false, true>},
                         {k_extend_w<true, true, false>, k_extend_w<true, true, true>}}};
                    const ExtendKernel k = table[count ? 1 : 0][media ? 1 : 0][inst ? 1 : 0];
                    if (!trace_grid)
                        trace_grid = sms * blocks_per_sm(k, kTraceBlock);
                    k<<<trace_grid, kTraceBlock, 0, st>>>(W, it);
                } else {
                    typedef void (*ExtendKernel)(WfParams, int);
                    static const ExtendKernel table[2][2][2] = {
                        {{k_extend<false, false, false>, k_extend<false, false, true>},
                         {k_extend<false, true, false>, k_extend<false, true, true>}},
                        {{k_extend<true, false, false>, k_extend<true, false, true>},
                         {k_extend<true, true, false>, k_extend<true, true, true>}}};
                    table[count ? 1 : 0][media ? 1 : 0][inst ? 1 : 0]<<<grid, 128, 0, st>>>(W, it);
                }
                mark();
                ++launches;
                for (int q = 0; q <= kTexturedKey; ++q) { // the two expensive types get their own launches
                    const int m = q == kTexturedKey ? int(RTB_MAT_LAMBERTIAN) : q;
                    if (m != RTB_MAT_LAMBERTIAN && m != RTB_MAT_PBR)
                        continue;
                    if (q == kTexturedKey ? !sc.host.textured_lambertian : !((W.mat_mask >> m) & 1u))
                        continue;
                    if (old_api)
                        launch_shade<true>(m, q, W, it, grid, st);
                    else
                        launch_shade<false>(m, q, W, it, grid, st);
                    ++launches;
                }
                if (W.mat_mask & minor_mask) {
                    if (old_api)
                        k_shade_minor<true><<<grid, 128, 0, st>>>(W, it);
                    else
                        k_shade_minor<false><<<grid, 128, 0, st>>>(W, it);
                    ++launches;
                }
                mark();
                if (miss_shades)
                    k_miss<true><<<grid, 128, 0, st>>>(W, it);
                else
                    k_miss<false><<<grid, 128, 0, st>>>(W, it);
                ++launches;
                mark();
                if (nee && wide) {
                    typedef void (*ConnectKernel)(WfParams, int);
                    static const ConnectKernel table[2][2][2] = {
                        {{k_connect_w<false, false, false>, k_connect_w<false, false, true>},
                         {k_connect_w<false, true, false>, k_connect_w<false, true, true>}},
                        {{k_connect_w<true, false, false>, k_connect_w<true, false, true>},
                         {k_connect_w<true, true, false>, k_connect_w<true, true, true>}}};
                    const ConnectKernel k = table[count ? 1 : 0][media ? 1 : 0][inst ? 1 : 0];
                    if (!connect_grid)
                        connect_grid = sms * blocks_per_sm(k, kTraceBlock);
                    k<<<connect_grid, kTraceBlock, 0, st>>>(W, it);
                    ++launches;
                } else if (nee) {
                    typedef void (*ConnectKernel)(WfParams, int);
                    static const ConnectKernel table[2][2][2] = {
                        {{k_connect<false, false, false>, k_connect<false, false, true>},
                         {k_connect<false, true, false>, k_connect<false, true, true>}},
                        {{k_connect<true, false, false>, k_connect<true, false, true>},
                         {k_connect<true, true, false>, k_connect<true, true, true>}}};
                    table[count ? 1 : 0][media ? 1 : 0][inst ? 1 : 0]<<<grid, 128, 0, st>>>(W, it);
                    ++launches;
                }
                mark();
            }
            RTB_CUDA(cudaGetLastError());
            // length of the NEXT extend queue, read back without stalling the pipeline: the
            // host only waits for the probe of the PREVIOUS batch.  An empty queue ends the
            // render once it has been seen twice in a row: the extend that found it empty
            // may still have started the last samples of the warps' private ranges.
            const int slot = probe & 1;
            RTB_CUDA(cudaMemcpyAsync(&pool.h_live[slot], &W.ctr[it % 3].n_ext.v[0], sizeof(uint32_t),
                                     cudaMemcpyDeviceToHost, st));
            RTB_CUDA(cudaEventRecord(pool.ev[slot], st));
            pending[slot] = it;
            const int prev = (probe + 1) & 1;
            if (pending[prev] >= 0) {
                RTB_CUDA(cudaEventSynchronize(pool.ev[prev]));
                zero_probes = pool.h_live[prev] == 0 ? zero_probes + 1 : 0;
                if (zero_probes >= 2)
                    done = true;
            }
            ++probe;
            if (ctx->cancel.load(std::memory_order_relaxed)) {
                cancelled = true;
                break;
            }
            // every entry is busy on every iteration until the samples run out, so the last
            // sample starts no later than iteration seg*max_depth/P; the tail adds max_depth
            if ((unsigned long long)(it - it0) >
                seg * (unsigned long long)rp.max_depth / P + 2ull * rp.max_depth + 16 * kBatch)
                throw std::runtime_error("wavefront: iteration bound exceeded (internal error)");
        }
    };

    // ---- which schedule ------------------------------------------------------------------
    // Not shared-memory sized: wavefront.  Small scenes: the fused kernel, except where every
    // bounce runs next-event estimation against area / delta lights on a mixed material set —
    // there the divergent Cook-Torrance + MIS shading and the partially filled inline shadow
    // rays cost more than the queues, and the material-sorted wavefront wins.  Measured on B200
    // (fused vs wavefront): C1 30 vs 55 ms, C3 69 vs 87 ms, C4-env 44 vs 55 ms, C4 17.4 vs 6.6 ms.
    const bool can_fuse = geom.flat && ctx->opt_fused != 0 && !(rp.flags & RTB_RENDER_FORCE_WAVEFRONT);
    const bool finite_lights = int(sc.host.f32.lights.size()) > sc.host.n_infinite_lights;
    const bool prefer_wavefront = !simple && nee && finite_lights;
    int schedule = 0;
    RTB_CUDA(cudaEventRecord(pool.ev_begin, st));
    RTB_CUDA(cudaMemsetAsync(d_accum, 0, size_t(npix_image) * sizeof(float4), st));
    k_clear<<<1, 256, 0, st>>>(W.ctr, W.glob);
    ++launches;
    if (can_fuse && (!prefer_wavefront || (rp.flags & RTB_RENDER_FORCE_FUSED))) {
        schedule = 1;
        run_fused(0, total);
    } else {
        run_wavefront(0, total, true);
    }
    const bool fused = schedule == 1;
    RTB_CUDA(cudaMemcpyAsync(pool.h_glob, W.glob, sizeof(Globals), cudaMemcpyDeviceToHost, st));
    RTB_CUDA(cudaEventRecord(pool.ev_end, st));
    RTB_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    RTB_CUDA(cudaEventElapsedTime(&ms, pool.ev_begin, pool.ev_end));
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        if (rp.max_depth <= 0)
            stats->paths = n_samples;
        else
            stats->paths = pool.h_glob->paths;
        stats->rays_closest = pool.h_glob->rays_closest;
        stats->rays_shadow = pool.h_glob->rays_shadow;
        stats->nodes_visited = pool.h_glob->nodes_visited;
        stats->prim_tests = pool.h_glob->prim_tests;
        stats->max_nodes_per_ray = pool.h_glob->max_nodes_per_ray;
        stats->extend_nodes = pool.h_glob->extend_nodes;
        stats->extend_chunk_max_nodes = pool.h_glob->extend_chunk_max_nodes;
        stats->iterations = uint64_t(it);
        stats->kernel_launches = launches;
        stats->device_ms = ms;
        if (fused) {
            for (size_t i = 0; i + 1 < n_ext_events; i += 2) {
                float e = 0.f;
                RTB_CUDA(cudaEventElapsedTime(&e, pool.ev_ext[i], pool.ev_ext[i + 1]));
                stats->extend_ms += e;
            }
            stats->extend_launches = n_ext_events / 2;
        } else {
            for (size_t i = 0; i + 4 < n_ext_events; i += 5)
                for (int k = 0; k < 4; ++k) {
                    float e = 0.f;
                    RTB_CUDA(cudaEventElapsedTime(&e, pool.ev_ext[i + k], pool.ev_ext[i + k + 1]));
                    stats->stage_ms[k] += e;
                }
            stats->extend_ms = stats->stage_ms[0];
            stats->extend_launches = n_ext_events / 5;
        }
        stats->schedule = fused ? 1 : 0;
        stats->traversal = fused ? 0 : traversal_used;
    }
    if (pool.h_glob->overflow)
        throw std::runtime_error("wavefront: queue overflow (internal error " + std::to_string(pool.h_glob->overflow) + ")");
    if (cancelled)
        throw std::runtime_error("cancelled");
}

void launch_trace_fast_batch(rtb_context *ctx, const rtb_ray *d_rays, uint64_t n, rtb_hit *d_hits,
                             unsigned long long *d_visits, bool plane_records) {
    const DeviceScene &sc = *ctx->scene;
    const GeomView<float> g = sc.geom<float>();
    if (!g.flat)
        throw std::runtime_error("precision 33 / 35 (the fused kernel's typed traversal) needs a scene of <= 64 primitive records");
    const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    k_trace_fast_batch<<<sms * 4, 128, 0, ctx->stream>>>(g, sc.orig_to_sorted.as<int32_t>(),
                                                         int(sc.host.orig_to_sorted.size()), d_rays, n, d_hits, d_visits,
                                                         plane_records);
    RTB_CUDA(cudaGetLastError());
}

void launch_trace_wide_batch(rtb_context *ctx, const rtb_ray *d_rays, uint64_t n, rtb_hit *d_hits,
                             unsigned long long *d_visits, bool any_hit) {
    const DeviceScene &sc = *ctx->scene;
    if (n >= (1ull << 31))
        throw std::runtime_error("precision 34 / 36: at most 2^31 - 1 rays per call");
    const GeomView<float> g = sc.geom<float>();
    const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    cudaStream_t st = ctx->stream;
    DeviceBuffer a, b, t, res, misc;
    a.alloc(n * 16);
    b.alloc(n * 16);
    t.alloc(n * 8);
    res.alloc(n * 8);
    misc.alloc(2 * sizeof(uint32_t)); // [0] work cursor, [1] overflow flag
    RTB_CUDA(cudaMemsetAsync(misc.as<void>(), 0, 2 * sizeof(uint32_t), st));
    k_batch_rays_in<<<sms * 4, 256, 0, st>>>(d_rays, uint32_t(n), sc.orig_to_sorted.as<int32_t>(),
                                             int(sc.host.orig_to_sorted.size()), a.as<float4>(), b.as<float4>(), t.as<float2>());
    BatchTraceJob job;
    job.ray_a = a.as<Vec4f>();
    job.ray_b = b.as<Vec4f>();
    job.ray_t = t.as<Vec2f>();
    job.out = res.as<Vec2f>();
    job.n = uint32_t(n);
    job.head = misc.as<uint32_t>();
    job.base = 0;
    job.seed = 0x51ed270b;
    const int grid = sms * (any_hit ? blocks_per_sm(k_trace_batch_w<true>, kTraceBlock) : blocks_per_sm(k_trace_batch_w<false>, kTraceBlock));
    job.win = 0;
    job.warps = uint32_t(grid) * (kTraceBlock / 32);
    if (any_hit)
        k_trace_batch_w<true><<<grid, kTraceBlock, 0, st>>>(g, sc.wide(), job, d_visits, misc.as<uint32_t>() + 1);
    else
        k_trace_batch_w<false><<<grid, kTraceBlock, 0, st>>>(g, sc.wide(), job, d_visits, misc.as<uint32_t>() + 1);
    k_batch_hits_out<<<sms * 4, 256, 0, st>>>(g, d_rays, res.as<float2>(), uint32_t(n), d_hits);
    RTB_CUDA(cudaGetLastError());
    uint32_t flag = 0;
    RTB_CUDA(cudaMemcpyAsync(&flag, misc.as<uint32_t>() + 1, sizeof(flag), cudaMemcpyDeviceToHost, st));
    RTB_CUDA(cudaStreamSynchronize(st));
    if (flag)
        throw std::runtime_error("wide traversal: internal error " + std::to_string(flag) + " (stack / scheduler guard)");
}

void launch_resolve_rgb8(rtb_context *ctx, const float4 *d_accum, int w, int h, int spp, uint8_t *d_rgb8,
                         cudaStream_t st) {
    const int sms = ctx->sm_count > 0 ? ctx->sm_count : 148;
    k_resolve_rgb8<<<sms * 4, 256, 0, st>>>(d_accum, w, h, 1.0 / spp, d_rgb8);
    RTB_CUDA(cudaGetLastError());
}

} // namespace rtb
