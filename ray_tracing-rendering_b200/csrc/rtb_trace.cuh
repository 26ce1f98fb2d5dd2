// rtb_trace.cuh — the production ray traversal of the wavefront renderer: a warp-scheduled walk
// of the 4-wide BVH (rtb_wide.cuh) by persistent warps that keep every lane supplied with rays.
// Replaces bvh_node::hit / aabb::hit (src/geometry/bvh.h:40-50, aabb.h:31-48) and everything the
// reference calls below them for one ray; closest hit for the extend stage, any hit for shadow
// rays (direct_light_integrator.h:115-130).
//
// Round 1's kernel handed each warp 32 rays and ran a per-lane while-while loop over a binary
// tree until the LAST ray of the 32 was done.  ncu on B200: 12 (scene09) to 18 (1 M spheres) of 32
// lanes active and 47-60 % of the issue slots, i.e. the SMs spent most of their time issuing
// instructions for a minority of lanes.  Three causes, three answers here:
//   * rays of one chunk differ in length (measured bound for any chunk-synchronous schedule:
//     0.56 / 0.74 of the lanes)            -> a lane that finishes takes the next ray of the
//     warp's WINDOW (up to 512 queue entries) at once; nobody waits for the longest ray of a chunk;
//   * lanes in the descent loop wait for lanes at a leaf and vice versa -> the warp VOTES on
//     what to do next: one 4-wide node step for the lanes that want one, or one primitive test
//     for the lanes parked at a leaf, whichever has the quorum; each step is straight-line code;
//   * neighbouring queue entries point in unrelated directions after the first bounce -> the
//     window is counting-sorted by direction octant in shared memory before its rays are handed
//     out, so the rays a warp walks at any moment share their traversal order.
// A node step is one 64-byte fetch (four LDG.128: the node's grid, its children's boxes as 8-bit
// planes, four refs — rtb_wide.cuh QNode64), one IDP4A per plane that adds the byte to the bits of 2^23
// (so the result IS the float 2^23 + q; a packed FADD2 takes the 2^23 off again, exactly — both on the FMA
// pipe: the first version used a byte-permute per plane, 15 % of all executed instructions on the half-rate
// ALU pipe that binds this kernel), twelve packed FFMA2 (two children per instruction, sm_100a) with the near / far words picked by the ray's signs, FMNMX3
// reductions, and a five-comparator sorting network over (entry distance, ref); children are
// pushed far to near with their entry distance, so a popped subtree that lies behind the closest
// hit found meanwhile is dropped without touching memory.
//
// Memory latency is kept off the warp's critical path (first GPU measurement of this design: the
// traversal itself was faster than round 1's, the kernel slower, because every refill, window
// set-up and queue push stalled all 32 lanes on dependent DRAM loads):
//   * the next 32 rays of the window are always ON DECK in shared memory, brought there by
//     cp.async (LDGSTS) issued one refill ahead, so a refill reads shared memory;
//   * the hit-queue key of a primitive (its material's type) is baked into the primitive record the
//     leaf test has just read, so sorting the window's hits by material needs no look-up;
//   * set-up and push are split into loops that only load / copy (unrolled, no warp collective in
//     them: their loads overlap) and loops that only vote on shared-memory bytes.
//
// The scheduler is written against WarpOps (rtb_warp.cuh): intrinsics on the device, a 32-fiber
// emulation in tests/hostcheck, where it is held ray by ray to traverse_wide().
#ifndef RTB_TRACE_CUH
#define RTB_TRACE_CUH

#include "rtb_shading.cuh"
#include "rtb_warp.cuh"
#include "rtb_wide.cuh"

namespace rtb {

#ifndef RTB_TRACE_NODE_MIN
#define RTB_TRACE_NODE_MIN 16 // a node step runs when at least this many lanes want one (or no lane waits at a leaf)
#endif
#ifndef RTB_TRACE_SWITCH_MIN
#define RTB_TRACE_SWITCH_MIN 8 // lanes without a ray that trigger a refill from the window
#endif
#ifndef RTB_TRACE_SORT
#define RTB_TRACE_SORT 0 // counting-sort each window by direction octant (measured on B200: no gain)
#endif
#ifndef RTB_TRACE_FULL_SORT
#define RTB_TRACE_FULL_SORT 1 // 1: children pushed far to near (five-comparator network); 0: nearest child next, the others pushed unsorted
#endif
#ifndef RTB_TRACE_TOP_CACHE
#define RTB_TRACE_TOP_CACHE 0 // 1: the top entry of a lane's stack lives in registers: a pop uses it at once and starts the load of the next one early
#endif
#ifndef RTB_TRACE_DP4A
#define RTB_TRACE_DP4A 1 // 1: plane bytes expanded by IDP4A + FADD2 (FMA pipe) instead of PRMT (ALU pipe, half rate): C5 -4 % (B200)
#endif
#ifndef RTB_TRACE_POSTPONE
#define RTB_TRACE_POSTPONE 0 // 1: a lane that reaches a leaf parks it and keeps descending (scenes without instances).  MEASURED on
                             // B200: the emulated warps gain lanes (0.65 -> 0.71 of them busy), the GPU loses 12 % (C5 68.3 -> 76.9 ms per
                             // 16 spp): the kernel is bound by the latency of a warp's dependent steps, not by idle lanes; off
#endif
#ifndef RTB_TRACE_LEAF_MIN
#define RTB_TRACE_LEAF_MIN 20 // with parked leaves: a primitive step runs when at least this many lanes have one (or no node quorum)
#endif
#ifndef RTB_TRACE_ANY_UNSORTED
#define RTB_TRACE_ANY_UNSORTED 0 // 1: any-hit queries walk a node's children in slot order (no sorting network)
#endif
#ifndef RTB_TRACE_PREFETCH
#define RTB_TRACE_PREFETCH 0 // 1: the second-nearest child's node is prefetched into L1 when it is pushed; 2: the nearest one's too
#endif
#ifndef RTB_TRACE_GUARD
#define RTB_TRACE_GUARD 1 // bound the scheduler loop (an internal error becomes a flag, not a hung GPU)
#endif
#ifndef RTB_TRACE_WINDOW
#define RTB_TRACE_WINDOW 512
#endif
#ifndef RTB_TRACE_TOP_NODES
#define RTB_TRACE_TOP_NODES 21 // three levels of the 4-wide tree
#endif
constexpr int kTraceWindow = RTB_TRACE_WINDOW;   // most queue entries a warp reserves, sorts and walks at a time
constexpr int kTopNodesMax = RTB_TRACE_TOP_NODES; // 128-byte nodes of the tree's top levels staged in shared memory
constexpr uint32_t kDeck = 32;                    // rays kept on deck in shared memory
constexpr uint32_t kDoneRef = 0xfffffffdu;        // lane state: its ray is finished, result not yet committed
constexpr uint32_t kIdleRef = 0xfffffffcu;        // lane state: no ray
constexpr uint32_t kSkipKey = 255;
constexpr uint32_t kFreshKey = 8; // window sort: 8 octants, then the empty entries where camera samples start

struct WideView {
    const Vec4f *nodes; // QNode64 array, as rows of 16 bytes (4 per node)
    const uint32_t *chain_root;
    uint32_t root_ref;
    uint32_t n_nodes;
    uint32_t n_global;                     // primitives outside the tree, tested for every ray
    uint32_t global_prim[kMaxGlobalPrims]; // (sorted primitive indices)
};

// Per-warp scratch in shared memory.
typedef uint16_t WindowPos; // an entry's offset inside its window
struct TraceWarpSmem {
    Vec4f deck_a[kDeck], deck_b[kDeck]; // rays of window positions [next, next + 32), slot = position % 32
    Vec2u deck_c[kDeck];                // their third word (job specific: rng state / t range / origin primitive)
    WindowPos perm[kTraceWindow];       // window position -> entry, sorted by key
    uint8_t key[kTraceWindow];          // per entry: sort key, later its hit-queue key
    uint32_t hist[16];
    uint32_t cur[16];
};

// window size for a queue of n entries walked by `warps` warps: large windows amortise the drain at a
// window's end, but a short queue (the tail of a render) must still spread over all warps
RTB_HD uint32_t window_size_for(uint32_t n, uint32_t warps) {
    uint32_t w = (n / (2u * (warps ? warps : 1u)) + 31u) & ~31u;
    w = w < 32u ? 32u : w;
    return w > uint32_t(kTraceWindow) ? uint32_t(kTraceWindow) : w;
}

// Guided self-scheduling: as the queue runs out the windows shrink, so the warps finish a launch
// together instead of the last ones each walking a full window alone.  b_prev / win_prev: where this
// warp's previous window started and how large it was (in the meantime every other warp has taken
// about one window of that size too).
RTB_HD uint32_t guided_window(uint32_t n, uint32_t warps, uint32_t b_prev, uint32_t win_prev) {
    const uint64_t gone = uint64_t(b_prev) + uint64_t(warps) * win_prev;
    const uint32_t rem = gone < n ? uint32_t(n - gone) : 0u;
    return window_size_for(rem, warps);
}

RTB_WD bool ref_is_node(uint32_t r) { return int32_t(r) >= 0; }
RTB_WD bool ref_is_leaf(uint32_t r) { return (r - kLeafFlag) < (kIdleRef - kLeafFlag); }

// ---- asynchronous global -> shared copies (LDGSTS); plain copies in the host emulation -------------
template <int BYTES> RTB_WD void async_copy(void *smem, const void *gmem) {
#ifdef __CUDA_ARCH__
    const uint32_t sa = uint32_t(__cvta_generic_to_shared(smem));
    if (BYTES == 16)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
    else if (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem) : "memory");
#else
    memcpy(smem, gmem, BYTES);
#endif
}
RTB_WD void async_wait_all() {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// The ray as the node step wants it: t = plane * idir + ood (one FMA per plane, |idir| capped so
// that ood stays finite for axis-parallel rays — the arithmetic of SlabRay<float, true>), and which
// of its direction components are negative (those axes enter a box through its hi plane).
struct TravRay {
    V3<float> o, d, idir, ood;
    uint32_t neg; // bit k: d[k] < 0
    RTB_WD void set(V3<float> o_, V3<float> d_) {
        o = o_;
        d = d_;
        const float eps = 1e-18f;
        const float dx = fabsf(d.x) < eps ? (d.x < 0 ? -eps : eps) : d.x;
        const float dy = fabsf(d.y) < eps ? (d.y < 0 ? -eps : eps) : d.y;
        const float dz = fabsf(d.z) < eps ? (d.z < 0 ? -eps : eps) : d.z;
        idir = V3<float>(1.0f / dx, 1.0f / dy, 1.0f / dz);
        ood = V3<float>(-(o.x * idir.x), -(o.y * idir.y), -(o.z * idir.z));
        neg = (dx < 0 ? 1u : 0u) | (dy < 0 ? 2u : 0u) | (dz < 0 ? 4u : 0u);
    }
};

// direction octant (the window sort key)
RTB_WD uint32_t octant_of(float dx, float dy, float dz) {
    return (dx < 0 ? 1u : 0u) | (dy < 0 ? 2u : 0u) | (dz < 0 ? 4u : 0u);
}

// The float 32768 + (byte i of word): the byte goes into mantissa bits 8..15 of 2^15 = 0x47000000.
RTB_WD float magic_byte(uint32_t word, int i) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(__byte_perm(word, 0x47000000u, 0x7504u | (uint32_t(i) << 4)));
#else
    return u2f(0x47000000u | (((word >> (8 * i)) & 0xffu) << 8));
#endif
}
// t of the four children's planes `word` (8-bit grid coordinates q): (32768 + q) * a + b, where the
// caller folded the grid into a = s * idir and b = (o * idir + ood) - 32768 * a.  Two packed FFMA2.
#if RTB_TRACE_DP4A && defined(__CUDACC__)
// one-hot byte selectors for IDP4A, in constant memory so that ptxas cannot fold the dot product back into PRMT
static __constant__ uint32_t g_onehot[4] = {0x00000001u, 0x00000100u, 0x00010000u, 0x01000000u};
#endif
RTB_WD void fma4q(uint32_t word, float a, float b, float out[4]) {
#ifdef __CUDA_ARCH__
#if RTB_TRACE_DP4A
    // 2^23 + q from one IDP4A (byte i of `word` times a one-hot byte, added to the bits of 2^23); minus 2^23
    // (exact) by a packed add; then the FFMA2.  `b` here is the plain o * idir + ood (see qnode_slabs).
    const float2 two23 = make_float2(-8388608.0f, -8388608.0f);
    uint32_t e0, e1, e2, e3;
    asm("dp4a.u32.u32 %0, %1, %2, 0x4B000000;" : "=r"(e0) : "r"(word), "r"(g_onehot[0]));
    asm("dp4a.u32.u32 %0, %1, %2, 0x4B000000;" : "=r"(e1) : "r"(word), "r"(g_onehot[1]));
    asm("dp4a.u32.u32 %0, %1, %2, 0x4B000000;" : "=r"(e2) : "r"(word), "r"(g_onehot[2]));
    asm("dp4a.u32.u32 %0, %1, %2, 0x4B000000;" : "=r"(e3) : "r"(word), "r"(g_onehot[3]));
    const float2 q01 = __fadd2_rn(make_float2(__uint_as_float(e0), __uint_as_float(e1)), two23);
    const float2 q23 = __fadd2_rn(make_float2(__uint_as_float(e2), __uint_as_float(e3)), two23);
    const float2 lo = __ffma2_rn(q01, make_float2(a, a), make_float2(b, b));
    const float2 hi = __ffma2_rn(q23, make_float2(a, a), make_float2(b, b));
#else
    const float2 lo = __ffma2_rn(make_float2(magic_byte(word, 0), magic_byte(word, 1)), make_float2(a, a), make_float2(b, b));
    const float2 hi = __ffma2_rn(make_float2(magic_byte(word, 2), magic_byte(word, 3)), make_float2(a, a), make_float2(b, b));
#endif
    out[0] = lo.x;
    out[1] = lo.y;
    out[2] = hi.x;
    out[3] = hi.y;
#else
    for (int i = 0; i < 4; ++i)
#if RTB_TRACE_DP4A
        out[i] = fmaf(float((word >> (8 * i)) & 0xffu), a, b);
#else
        out[i] = fmaf(magic_byte(word, i), a, b);
#endif
#endif
}

// Entry distances of the four children of a quantised node (rows r0..r2 of QNode64) for the ray,
// clipped to [t_min, t_max]: k[i] = +inf where the ray misses child i.  Shared by the kernels'
// node step and the scalar reference traversal, so both walk exactly the same nodes.
RTB_WD void qnode_slabs(const Vec4f &r0, const Vec4f &r1, const Vec4f &r2, const TravRay &r, float t_min, float t_max,
                        float k[4]) {
    const float ax = r0.w * r.idir.x, ay = r2.z * r.idir.y, az = r2.w * r.idir.z;
#if RTB_TRACE_DP4A
    const float bx = fmaf(r0.x, r.idir.x, r.ood.x), by = fmaf(r0.y, r.idir.y, r.ood.y), bz = fmaf(r0.z, r.idir.z, r.ood.z);
#else
    const float bx = fmaf(-32768.0f, ax, fmaf(r0.x, r.idir.x, r.ood.x));
    const float by = fmaf(-32768.0f, ay, fmaf(r0.y, r.idir.y, r.ood.y));
    const float bz = fmaf(-32768.0f, az, fmaf(r0.z, r.idir.z, r.ood.z));
#endif
    const uint32_t lox = f2u(r1.x), loy = f2u(r1.y), loz = f2u(r1.z), hix = f2u(r1.w), hiy = f2u(r2.x), hiz = f2u(r2.y);
    const bool nx = (r.neg & 1u) != 0, ny = (r.neg & 2u) != 0, nz = (r.neg & 4u) != 0;
    float tnx[4], tny[4], tnz[4], tfx[4], tfy[4], tfz[4];
    fma4q(nx ? hix : lox, ax, bx, tnx);
    fma4q(nx ? lox : hix, ax, bx, tfx);
    fma4q(ny ? hiy : loy, ay, by, tny);
    fma4q(ny ? loy : hiy, ay, by, tfy);
    fma4q(nz ? hiz : loz, az, bz, tnz);
    fma4q(nz ? loz : hiz, az, bz, tfz);
    const float inf = Consts<float>::inf();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float tn = fmaxf(fmaxf(tnx[i], tny[i]), fmaxf(tnz[i], t_min));
        const float tf = fminf(fminf(tfx[i], tfy[i]), fminf(tfz[i], t_max));
        k[i] = tn <= tf ? tn : inf;
    }
}

// Staged nodes sit 80 bytes apart (one 16-byte row of padding): lanes of a quarter warp that read
// the same row of DIFFERENT nodes then hit different banks (the first version staged 128-byte nodes
// back to back and lost more to bank conflicts — ncu: 40 M conflict wavefronts per launch on
// scene09 — than it saved).
constexpr uint32_t kNodeRows = 4;
constexpr uint32_t kTopStride = 5; // rows
template <bool TOP> RTB_WD const Vec4f *node_address(const WideView &w, const Vec4f *s_top, uint32_t n_top, uint32_t cur) {
    if (TOP && cur < n_top)
        return s_top + size_t(cur) * kTopStride;
    return w.nodes + size_t(cur) * kNodeRows;
}
template <bool TOP> RTB_WD Vec4f load_row(const Vec4f *p) {
#ifdef __CUDA_ARCH__
    if (!TOP) // global memory for certain: read-only path
        return __ldg(reinterpret_cast<const float4 *>(p));
#endif
    return *p;
}

RTB_WD void cmp_swap(float &ka, uint32_t &ra, float &kb, uint32_t &rb) {
    const bool s = kb < ka;
    const float k0 = s ? kb : ka, k1 = s ? ka : kb;
    const uint32_t r0 = s ? rb : ra, r1 = s ? ra : rb;
    ka = k0;
    kb = k1;
    ra = r0;
    rb = r1;
}

// The state of one lane's traversal.  The world-space ray is not kept: a lane that leaves an
// instance asks its job for it again (Job::world_ray), which is rare and saves six registers.
struct TravLane {
    TravRay r;
    float time, t_min, t_max;
    uint32_t best, best_key, origin, cur, sp;
    uint32_t pend; // a parked leaf (the rest of it), or kIdleRef
    Pcg rng;
#if RTB_TRACE_TOP_CACHE
    Vec2u top; // entry sp - 1 of the stack (the array holds the entries below it)
#endif
};
RTB_WD void stk_push(TravLane &L, Vec2u *stack, Vec2u e) {
#if RTB_TRACE_TOP_CACHE
    if (L.sp)
        stack[L.sp - 1] = L.top;
    L.top = e;
    ++L.sp;
#else
    stack[L.sp++] = e;
#endif
}
RTB_WD Vec2u stk_pop(TravLane &L, Vec2u *stack) { // L.sp > 0
#if RTB_TRACE_TOP_CACHE
    const Vec2u e = L.top;
    if (--L.sp)
        L.top = stack[L.sp - 1];
    return e;
#else
    return stack[--L.sp];
#endif
}

// Next subtree off the lane's stack (skipping those behind the closest hit so far), or kDoneRef.
template <bool INST, class Job> RTB_WD void trav_pop(TravLane &L, Vec2u *stack, Job &job, uint32_t tag) {
    for (;;) {
        if (L.sp == 0) {
            L.cur = kDoneRef;
            return;
        }
        const Vec2u e = stk_pop(L, stack);
        if (INST && e.x == kSentinelRef) { // leaving the instance: back to the world ray
            V3<float> wo, wd;
            job.world_ray(tag, wo, wd);
            L.r.set(wo, wd);
            continue;
        }
        if (u2f(e.y) <= L.t_max && e.x != kEmptyRef) { // (an unused slot of a degenerate node can pass the slab test)
            L.cur = e.x;
            return;
        }
    }
}

// One 4-wide node: L.cur is an interior node on entry, the next thing to do on exit.
template <bool ANY, bool TOP, bool INST, class Job>
RTB_WD void trav_node_step(const WideView &w, const Vec4f *s_top, uint32_t n_top, TravLane &L, Vec2u *stack,
                           uint32_t &overflow, Job &job, uint32_t tag) {
    const Vec4f *nb = node_address<TOP>(w, s_top, n_top, L.cur);
    const Vec4f q0 = load_row<TOP>(nb), q1 = load_row<TOP>(nb + 1), q2 = load_row<TOP>(nb + 2), rr = load_row<TOP>(nb + 3);
    const float inf = Consts<float>::inf();
    float k[4];
    qnode_slabs(q0, q1, q2, L.r, L.t_min, L.t_max, k);
    uint32_t r0 = f2u(rr.x), r1 = f2u(rr.y), r2 = f2u(rr.z), r3 = f2u(rr.w);
#if RTB_TRACE_ANY_UNSORTED
    if (ANY) {
        // any-hit queries (shadow rays) do not need the nearest child first: the children are walked in slot order,
        // no sorting network (25 ALU-pipe instructions of a node step)
        const bool h1 = k[1] < inf, h2 = k[2] < inf, h3 = k[3] < inf;
        if (h1 | h2 | h3) {
            if (L.sp + 3u > uint32_t(kWideStack)) {
                overflow = 1u;
            } else {
                if (h3)
                    stk_push(L, stack, Vec2u{r3, f2u(k[3])});
                if (h2)
                    stk_push(L, stack, Vec2u{r2, f2u(k[2])});
                if (h1)
                    stk_push(L, stack, Vec2u{r1, f2u(k[1])});
            }
        }
        if (k[0] < inf && r0 != kEmptyRef)
            L.cur = r0;
        else
            trav_pop<INST>(L, stack, job, tag);
        return;
    }
#endif
#if RTB_TRACE_FULL_SORT
    // sorting network, ascending entry distance: (0,1) (2,3) (0,2) (1,3) (1,2)
    cmp_swap(k[0], r0, k[1], r1);
    cmp_swap(k[2], r2, k[3], r3);
    cmp_swap(k[0], r0, k[2], r2);
    cmp_swap(k[1], r1, k[3], r3);
    cmp_swap(k[1], r1, k[2], r2);
#if RTB_TRACE_PREFETCH && defined(__CUDA_ARCH__)
    if (k[1] < inf && ref_is_node(r1) && !(TOP && r1 < n_top))
        asm volatile("prefetch.global.L1 [%0];" ::"l"(w.nodes + size_t(r1) * kNodeRows));
#if RTB_TRACE_PREFETCH >= 2
    if (k[0] < inf && ref_is_node(r0) && !(TOP && r0 < n_top))
        asm volatile("prefetch.global.L1 [%0];" ::"l"(w.nodes + size_t(r0) * kNodeRows));
#endif
#endif
    // far to near onto the stack; the nearest is next
    if (k[1] < inf) {
        if (L.sp + 3u > uint32_t(kWideStack)) {
            overflow = 1u;
        } else {
            if (k[3] < inf)
                stk_push(L, stack, Vec2u{r3, f2u(k[3])});
            if (k[2] < inf)
                stk_push(L, stack, Vec2u{r2, f2u(k[2])});
            stk_push(L, stack, Vec2u{r1, f2u(k[1])});
        }
    }
#else
    // The nearest child goes to slot 0 (three comparators), the others are pushed as they lie: a popped
    // entry carries its entry distance and is dropped when it lies behind the closest hit found meanwhile,
    // so their order only decides how soon that happens.  15 ALU-pipe instructions fewer per node step.
    cmp_swap(k[0], r0, k[1], r1);
    cmp_swap(k[0], r0, k[2], r2);
    cmp_swap(k[0], r0, k[3], r3);
    const bool h1 = k[1] < inf, h2 = k[2] < inf, h3 = k[3] < inf;
    if (h1 | h2 | h3) {
        if (L.sp + 3u > uint32_t(kWideStack)) {
            overflow = 1u;
        } else {
            if (h3)
                stk_push(L, stack, Vec2u{r3, f2u(k[3])});
            if (h2)
                stk_push(L, stack, Vec2u{r2, f2u(k[2])});
            if (h1)
                stk_push(L, stack, Vec2u{r1, f2u(k[1])});
        }
    }
#endif
    if (k[0] < inf && r0 != kEmptyRef)
        L.cur = r0;
    else
        trav_pop<INST>(L, stack, job, tag);
}

// One primitive (not an instance) against the lane's ray: updates the closest hit; true when an
// any-hit query is decided.
template <bool ANY, bool MEDIA>
RTB_WD bool trav_test_prim(const GeomView<float> &g, TravLane &L, const PrimT<float> &p, uint32_t type, uint32_t index,
                           uint32_t &tests) {
    if (type == PT_BOX) { // a grouped box: one slab test, the hit names the face's own rect record
        uint64_t one = 0;
        const uint32_t before = L.best;
        const bool done = box_slot<float, ANY, true>(g, p, L.r.o, L.r.d, L.r.idir, L.time, L.t_min, L.t_max, L.origin, L.best, &one);
        tests += uint32_t(one);
        if (L.best != before)
            L.best_key = (p.type_mat >> PT_KEY_SHIFT) & PT_KEY_MASK;
        return done;
    }
    ++tests;
    float t;
    bool h;
    if (MEDIA && type == PT_MEDIUM) {
        h = hit_medium<float, true>(g, p, L.r.o, L.r.d, L.time, L.t_min, L.t_max, L.rng.next_open(), t);
        if (p.type_mat & PT_DUP_LEAF) { // second visit of a one-object bvh_node (bvh.h:46-47)
            float t2;
            if (hit_medium<float, true>(g, p, L.r.o, L.r.d, L.time, L.t_min, h ? t : L.t_max, L.rng.next_open(), t2)) {
                h = true;
                t = t2;
            }
        }
    } else {
        // hit_simple() wants 1/d of the current-space ray for the rect tests: the capped one is the same
        // number wherever a rect can be hit at all (|d| >= 1e-18)
        h = hit_simple<float, true>(g, p, type, L.r.o, L.r.d, L.r.idir, L.time, L.t_min, L.t_max, index == L.origin, t);
    }
    if (h) {
        L.best = index;
        L.best_key = (p.type_mat >> PT_KEY_SHIFT) & PT_KEY_MASK;
        L.t_max = t;
    }
    return ANY && h;
}

// One primitive of the leaf L.cur (or the entry into an instance); the rest of the leaf stays in L.cur.
template <bool ANY, bool MEDIA, bool INST, class Job>
RTB_WD void trav_leaf_step(const GeomView<float> &g, const WideView &w, TravLane &L, Vec2u *stack, uint32_t &overflow,
                           uint32_t &tests, Job &job, uint32_t tag) {
    if (!INST && L.pend != kIdleRef) { // a parked leaf goes first: one primitive of it (no instances, no stack)
        const uint32_t first = L.pend & kLeafFirstMask, more = (L.pend >> 27) & 15u;
        const PrimT<float> p = g.prims[first];
        if (trav_test_prim<ANY, MEDIA>(g, L, p, p.type_mat & PT_TYPE_MASK, first, tests)) {
            L.cur = kDoneRef;
            L.pend = kIdleRef;
            return;
        }
        L.pend = more ? (kLeafFlag | ((more - 1u) << 27) | (first + 1u)) : kIdleRef;
        if (L.pend == kIdleRef) {
            // the closest hit may have moved: a node the lane is about to visit can now lie behind it, which the
            // node step's own slab test against t_max sorts out; a leaf waiting in cur is parked and replaced
            if (ref_is_leaf(L.cur)) {
                L.pend = L.cur;
                trav_pop<INST>(L, stack, job, tag);
            }
        }
        return;
    }
    const uint32_t first = L.cur & kLeafFirstMask, more = (L.cur >> 27) & 15u;
    const PrimT<float> p = g.prims[first];
    const uint32_t type = p.type_mat & PT_TYPE_MASK;
    if (INST && type == PT_INSTANCE) { // alone in its leaf (builder guarantee); only ever met by the world ray
        if (L.sp + 1u > uint32_t(kWideStack)) {
            overflow = 1u;
            trav_pop<INST>(L, stack, job, tag);
            return;
        }
        stk_push(L, stack, Vec2u{kSentinelRef, f2u(-Consts<float>::inf())});
        V3<float> co = L.r.o, cd = L.r.d;
        enter_instance<float, true>(g, int(p.aux2), co, cd);
        L.r.set(co, cd);
        L.cur = w.chain_root[p.aux2];
        if (L.cur == kEmptyRef)
            trav_pop<INST>(L, stack, job, tag);
        return;
    }
    if (trav_test_prim<ANY, MEDIA>(g, L, p, type, first, tests)) {
        L.cur = kDoneRef;
        return;
    }
    if (more)
        L.cur = kLeafFlag | ((more - 1u) << 27) | (first + 1u);
    else
        trav_pop<INST>(L, stack, job, tag);
}

// ---- window sort --------------------------------------------------------------------------------
// Counting sort of a window's entries by key, in two halves.  s.key[j] (j < cnt) is a key < n_keys or
// kSkipKey.  window_hist() leaves the count of every key in s.hist.  window_place() keeps at most
// `limit` entries of key `limit_key` (the others become kSkipKey), fills s.perm (window position ->
// entry, ascending key), leaves the first position of every key in s.hist and returns the number of
// entries placed.  Only shared-memory bytes are touched: no global load sits between two votes.
RTB_WD void window_hist(TraceWarpSmem &s, uint32_t cnt, uint32_t n_keys) {
    const uint32_t lane = WarpOps::lane();
    if (lane < 16)
        s.hist[lane] = 0;
    WarpOps::sync();
    for (uint32_t j0 = 0; j0 < cnt; j0 += 32u) {
        const uint32_t j = j0 + lane;
        const uint32_t key = j < cnt ? s.key[j] : kSkipKey;
        const uint32_t peers = WarpOps::match_any(key);
        if (key < n_keys && int(lane) == WarpOps::ffs(peers) - 1)
            s.hist[key] += WarpOps::popc(peers);
        WarpOps::sync();
    }
}
RTB_WD uint32_t window_place(TraceWarpSmem &s, uint32_t cnt, uint32_t n_keys, uint32_t limit_key, uint32_t limit) {
    const uint32_t lane = WarpOps::lane();
    uint32_t off = 0, mine = 0, off_limit = 0;
    for (uint32_t k = 0; k < n_keys; ++k) {
        uint32_t h = s.hist[k];
        if (k == limit_key) {
            off_limit = off;
            if (h > limit)
                h = limit;
        }
        if (k == lane)
            mine = off;
        off += h;
    }
    const uint32_t total = off;
    WarpOps::sync();
    if (lane < n_keys) {
        s.cur[lane] = mine;
        s.hist[lane] = mine;
    }
    WarpOps::sync();
    for (uint32_t j0 = 0; j0 < cnt; j0 += 32u) {
        const uint32_t j = j0 + lane;
        const uint32_t key = j < cnt ? s.key[j] : kSkipKey;
        const uint32_t peers = WarpOps::match_any(key);
        const int leader = WarpOps::ffs(peers) - 1;
        uint32_t base = 0;
        if (key < n_keys && int(lane) == leader) {
            base = s.cur[key];
            s.cur[key] = base + WarpOps::popc(peers);
        }
        base = WarpOps::shfl(base, leader);
        const uint32_t pos = base + WarpOps::popc(peers & ((1u << lane) - 1u));
        if (key < n_keys) {
            if (key == limit_key && pos - off_limit >= limit)
                s.key[j] = uint8_t(kSkipKey);
            else
                s.perm[pos] = WindowPos(j);
        }
        WarpOps::sync();
    }
    return total;
}

#if !defined(__CUDACC__)
// host emulation only: the scheduler's thresholds as run-time values (tools/sched_sim.py sweeps them)
struct TraceTuning {
    uint32_t node_min = RTB_TRACE_NODE_MIN, switch_min = RTB_TRACE_SWITCH_MIN, leaf_min = RTB_TRACE_LEAF_MIN;
    bool postpone = RTB_TRACE_POSTPONE != 0;
};
inline TraceTuning &trace_tuning() {
    static TraceTuning t;
    return t;
}
#define RTB_NODE_MIN_ (trace_tuning().node_min)
#define RTB_SWITCH_MIN_ (trace_tuning().switch_min)
#define RTB_LEAF_MIN_ (trace_tuning().leaf_min)
// what the votes decided (steps) and how many lanes took part (lanes)
struct TraceSchedStats {
    uint64_t node_steps = 0, node_lanes = 0, leaf_steps = 0, leaf_lanes = 0, switches = 0, switch_lanes = 0;
};
inline TraceSchedStats &trace_sched_stats() {
    static TraceSchedStats s;
    return s;
}
#else
#define RTB_NODE_MIN_ uint32_t(RTB_TRACE_NODE_MIN)
#define RTB_SWITCH_MIN_ uint32_t(RTB_TRACE_SWITCH_MIN)
#define RTB_LEAF_MIN_ uint32_t(RTB_TRACE_LEAF_MIN)
#endif

// ---- the scheduler ----------------------------------------------------------------------------
// Job (called by the whole warp unless noted):
//   bool next_window(TraceWarpSmem&, uint32_t &count)   reserve + sort the next window; false: no work left
//   void prefetch(TraceWarpSmem&, uint32_t from, uint32_t to)  start the async copies of window positions
//                                                       [from, to) into their deck slots (position % 32)
//   void fetch(TraceWarpSmem&, uint32_t k, TravLane&, V3 &o, V3 &d, uint32_t &tag)  [per lane] the ray of
//                                                       window position k, from its deck slot
//   void world_ray(uint32_t tag, V3 &o, V3 &d)         [per lane] that ray again (leaving an instance)
//   void commit(TraceWarpSmem&, uint32_t tag, const TravLane&)      [per lane] its result
//   void finish_window(TraceWarpSmem&)                  all rays of the window are committed
// counters[0] += node steps, [1] += primitive tests, [2] = max(node steps of one ray) (per lane; the
// caller reduces them).
template <class Job, bool ANY, bool MEDIA, bool INST, bool TOP>
RTB_WD void warp_trace(const GeomView<float> &g, const WideView &w, const Vec4f *s_top, uint32_t n_top, TraceWarpSmem &s,
                       Job &job, uint64_t counters[3], uint32_t &overflow) {
#if defined(__CUDACC__)
    constexpr bool POSTPONE = RTB_TRACE_POSTPONE && !INST; // (a parked leaf belongs to the space the ray was in)
#else
    const bool POSTPONE = trace_tuning().postpone && !INST;
#endif
    const uint32_t lane = WarpOps::lane();
    Vec2u stack[kWideStack];
    TravLane L;
    L.cur = kIdleRef;
    L.pend = kIdleRef;
    L.sp = 0;
    L.best = kNoPrim;
    L.best_key = 0;
    L.rng.s = 0;
    uint32_t tag = 0;
    uint32_t win_count = 0, win_next = 0; // warp-uniform
    uint32_t nodes = 0, tests = 0, ray_start = 0, ray_max = 0;
#if RTB_TRACE_GUARD
    uint32_t guard = 0;
#endif
    for (;;) {
        // ---- switch point: commit finished rays, hand out the window's next ones
        if (L.cur == kDoneRef && (!POSTPONE || L.pend == kIdleRef)) {
            job.commit(s, tag, L);
            L.cur = kIdleRef;
            ray_max = nodes - ray_start > ray_max ? nodes - ray_start : ray_max;
        }
        const uint32_t idle = WarpOps::ballot(L.cur == kIdleRef);
        if (idle && win_next < win_count) {
#if !defined(__CUDACC__)
            if (lane == 0) {
                trace_sched_stats().switches += 1;
                const uint32_t left = win_count - win_next;
                trace_sched_stats().switch_lanes += WarpOps::popc(idle) < left ? WarpOps::popc(idle) : left;
            }
#endif
            async_wait_all(); // the deck holds positions [win_next, win_next + 32)
            WarpOps::sync();
            const uint32_t k = win_next + WarpOps::popc(idle & ((1u << lane) - 1u));
            if (L.cur == kIdleRef && k < win_count) {
                V3<float> o, d;
                job.fetch(s, k, L, o, d, tag);
                ray_start = nodes;
                L.best = kNoPrim;
                L.pend = kIdleRef;
                L.sp = 0;
                L.r.set(o, d);
                L.cur = w.root_ref == kEmptyRef ? kDoneRef : w.root_ref;
                for (uint32_t gi = 0; gi < w.n_global; ++gi) { // the primitives kept out of the tree
                    const uint32_t pi = w.global_prim[gi];
                    const PrimT<float> p = g.prims[pi];
                    if (trav_test_prim<ANY, MEDIA>(g, L, p, p.type_mat & PT_TYPE_MASK, pi, tests)) {
                        L.cur = kDoneRef;
                        break;
                    }
                }
            }
            const uint32_t taken = WarpOps::popc(idle);
            const uint32_t next = win_next + taken < win_count ? win_next + taken : win_count;
            WarpOps::sync(); // the consumed deck slots are free: refill them one switch ahead
            const uint32_t from = win_next + kDeck, to = next + kDeck < win_count ? next + kDeck : win_count;
            if (from < to)
                job.prefetch(s, from, to);
            win_next = next;
        }
        if (WarpOps::ballot(L.cur != kIdleRef) == 0) { // nothing in flight, nothing left to hand out
            if (win_count)
                job.finish_window(s);
            win_count = 0;
            win_next = 0;
            if (!job.next_window(s, win_count))
                break;
            job.prefetch(s, 0u, win_count < kDeck ? win_count : kDeck);
            continue;
        }
        // ---- walk until enough lanes have finished.  The common case — a quorum of lanes wants a node
        // step — is decided by ONE vote; everything else (leaf steps, refill, the end of the window) is
        // looked at only when that quorum fails.
        const bool can_refill = win_next < win_count;
        for (;;) {
            const uint32_t m_node = WarpOps::ballot(ref_is_node(L.cur));
            bool node_step = WarpOps::popc(m_node) >= RTB_NODE_MIN_;
            if (POSTPONE) {
                // Parked leaves: a lane with a leaf in `pend` still takes node steps, so the lanes at a node and the
                // lanes with primitive work overlap.  Primitive step when enough lanes have one, or the node
                // quorum fails; refill when enough lanes have neither.
                const uint32_t m_leaf = WarpOps::ballot(L.pend != kIdleRef || ref_is_leaf(L.cur));
                if ((m_node | m_leaf) == 0)
                    break;
                if (can_refill && 32u - WarpOps::popc(m_node | m_leaf) >= RTB_SWITCH_MIN_)
                    break;
#if RTB_TRACE_GUARD
                if (++guard > (1u << 28)) {
                    overflow = 2u;
                    L.cur = kIdleRef;
                    L.pend = kIdleRef;
                    win_next = win_count;
                    break;
                }
#endif
                if (WarpOps::popc(m_leaf) >= RTB_LEAF_MIN_)
                    node_step = false;
                else if (!node_step)
                    node_step = m_leaf == 0 || (m_node != 0 && WarpOps::popc(m_node) >= WarpOps::popc(m_leaf));
#if !defined(__CUDACC__)
                if (lane == 0 && !node_step) {
                    trace_sched_stats().leaf_steps += 1;
                    trace_sched_stats().leaf_lanes += WarpOps::popc(m_leaf);
                }
#endif
            } else if (!node_step) {
                const uint32_t m_leaf = WarpOps::ballot(ref_is_leaf(L.cur));
                if ((m_node | m_leaf) == 0)
                    break;
                if (can_refill && 32u - WarpOps::popc(m_node | m_leaf) >= RTB_SWITCH_MIN_)
                    break;
#if RTB_TRACE_GUARD
                if (++guard > (1u << 26)) {
                    overflow = 2u;
                    L.cur = kIdleRef;
                    win_next = win_count;
                    break;
                }
#endif
                node_step = m_leaf == 0;
#if !defined(__CUDACC__)
                if (lane == 0 && !node_step) {
                    trace_sched_stats().leaf_steps += 1;
                    trace_sched_stats().leaf_lanes += WarpOps::popc(m_leaf);
                }
#endif
            }
#if !defined(__CUDACC__)
            if (lane == 0 && node_step) {
                trace_sched_stats().node_steps += 1;
                trace_sched_stats().node_lanes += WarpOps::popc(m_node);
            }
#endif
            if (node_step) {
                if (ref_is_node(L.cur)) {
                    ++nodes;
                    trav_node_step<ANY, TOP, INST>(w, s_top, n_top, L, stack, overflow, job, tag);
                    if (POSTPONE && L.pend == kIdleRef && ref_is_leaf(L.cur)) { // park the leaf, keep walking
                        L.pend = L.cur;
                        trav_pop<INST>(L, stack, job, tag);
                    }
                }
            } else {
                if ((POSTPONE && L.pend != kIdleRef) || ref_is_leaf(L.cur))
                    trav_leaf_step<ANY, MEDIA, INST>(g, w, L, stack, overflow, tests, job, tag);
            }
        }
    }
    counters[0] += nodes;
    counters[1] += tests;
    counters[2] = ray_max > counters[2] ? ray_max : counters[2];
}

// ---- batch job: rays in, (t, primitive) out — rtb_trace_batch precision 34 / 36 and the CPU suite ----
struct BatchTraceJob {
    const Vec4f *ray_a; // origin, time
    const Vec4f *ray_b; // direction, origin primitive (sorted index or kNoPrim)
    const Vec2f *ray_t; // t_min, t_max
    Vec2f *out;         // t, primitive (sorted index, as bits)
    uint32_t n;
    uint32_t *head; // work cursor (global)
    uint32_t base;  // first entry of the current window
    uint32_t win;   // size of the window being walked (0 before the first one)
    uint32_t warps; // warps walking this batch
    uint64_t seed;
    RTB_WD bool next_window(TraceWarpSmem &s, uint32_t &count) {
        const uint32_t lane = WarpOps::lane();
        win = guided_window(n, warps, base, win);
        uint32_t b = 0;
        if (lane == 0)
            b = WarpOps::atomic_add(head, win);
        b = WarpOps::shfl(b, 0);
        if (b >= n)
            return false;
        base = b;
        const uint32_t cnt = n - b < win ? n - b : win;
        for (uint32_t j = lane; j < cnt; j += 32u) {
            const Vec4f d = ray_b[b + j];
            s.key[j] = uint8_t(RTB_TRACE_SORT ? octant_of(d.x, d.y, d.z) : 0u);
        }
        WarpOps::sync();
        window_hist(s, cnt, 8u);
        count = window_place(s, cnt, 8u, kSkipKey, 0u);
        return true;
    }
    RTB_WD void prefetch(TraceWarpSmem &s, uint32_t from, uint32_t to) {
        const uint32_t k = from + WarpOps::lane();
        if (k < to) {
            const uint32_t idx = base + s.perm[k], slot = k % kDeck;
            async_copy<16>(&s.deck_a[slot], ray_a + idx);
            async_copy<16>(&s.deck_b[slot], ray_b + idx);
            async_copy<8>(&s.deck_c[slot], ray_t + idx);
        }
    }
    RTB_WD void fetch(TraceWarpSmem &s, uint32_t k, TravLane &L, V3<float> &o, V3<float> &d, uint32_t &tag) {
        tag = base + s.perm[k];
        const uint32_t slot = k % kDeck;
        const Vec4f a = s.deck_a[slot], b = s.deck_b[slot];
        const Vec2u t = s.deck_c[slot];
        o = V3<float>(a.x, a.y, a.z);
        d = V3<float>(b.x, b.y, b.z);
        L.time = a.w;
        L.origin = f2u(b.w);
        L.t_min = u2f(t.x);
        L.t_max = u2f(t.y);
        L.rng = pcg_seed(tag, seed);
    }
    RTB_WD void world_ray(uint32_t tag, V3<float> &o, V3<float> &d) {
        const Vec4f a = ray_a[tag], b = ray_b[tag];
        o = V3<float>(a.x, a.y, a.z);
        d = V3<float>(b.x, b.y, b.z);
    }
    RTB_WD void commit(TraceWarpSmem &, uint32_t tag, const TravLane &L) { out[tag] = Vec2f{L.t_max, u2f(L.best)}; }
    RTB_WD void finish_window(TraceWarpSmem &) {}
};

} // namespace rtb

#endif // RTB_TRACE_CUH
