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
//     warp's WINDOW (256 queue entries) at once; nobody waits for the longest ray of a chunk;
//   * lanes in the descent loop wait for lanes at a leaf and vice versa -> the warp VOTES on
//     what to do next: one 4-wide node step for the lanes that want one, or one primitive test
//     for the lanes parked at a leaf, whichever has the quorum; each step is straight-line code;
//   * neighbouring queue entries point in unrelated directions after the first bounce -> the
//     window is counting-sorted by direction octant in shared memory before its rays are handed
//     out, so the rays a warp walks at any moment share their traversal order.
// A node step is one 128-byte fetch (seven LDG.128, or shared memory for the top of the tree),
// twelve packed FFMA2 (two children per instruction, sm_100a) with the near / far rows picked by
// the ray's signs, FMNMX3 reductions, and a five-comparator sorting network over (entry distance,
// ref); children are pushed far to near with their entry distance, so a popped subtree that lies
// behind the closest hit found meanwhile is dropped without touching memory.
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
#define RTB_TRACE_NODE_MIN 14 // a node step runs when at least this many lanes want one (or no lane waits at a leaf)
#endif
#ifndef RTB_TRACE_SWITCH_MIN
#define RTB_TRACE_SWITCH_MIN 6 // lanes without a ray that trigger a refill from the window
#endif
#ifndef RTB_TRACE_SORT
#define RTB_TRACE_SORT 1 // counting-sort each window by direction octant
#endif
#ifndef RTB_TRACE_GUARD
#define RTB_TRACE_GUARD 1 // bound the scheduler loop (an internal error becomes a flag, not a hung GPU)
#endif

#ifndef RTB_TRACE_WINDOW
#define RTB_TRACE_WINDOW 512
#endif
constexpr int kTraceWindow = RTB_TRACE_WINDOW; // queue entries a warp reserves, sorts and walks at a time
constexpr int kTraceRounds = kTraceWindow / 32;
constexpr int kTopNodesMax = 64;       // 128-byte nodes of the tree's top levels staged in shared memory (8 KB)
constexpr uint32_t kDoneRef = 0xfffffffdu; // lane state: its ray is finished, result not yet committed
constexpr uint32_t kIdleRef = 0xfffffffcu; // lane state: no ray
constexpr uint32_t kSortKeys = 9;      // 8 octants + fresh camera samples
constexpr uint32_t kSkipKey = 255;

struct WideView {
    const Vec4f *nodes; // Node128 array, as rows of 16 bytes
    const uint32_t *chain_root;
    uint32_t root_ref;
    uint32_t n_nodes;
};

// Per-warp scratch in shared memory.
typedef uint16_t WindowPos; // an entry's offset inside its window
struct TraceWarpSmem {
    WindowPos perm[kTraceWindow]; // window position -> entry (offset inside the window), sorted by key
    uint8_t key[kTraceWindow];  // per entry: sort key, later the hit-queue key
    uint32_t hist[16];
    uint32_t cur[16];
};

RTB_WD bool ref_is_node(uint32_t r) { return int32_t(r) >= 0; }
RTB_WD bool ref_is_leaf(uint32_t r) { return (r - kLeafFlag) < (kIdleRef - kLeafFlag); }

// The ray as the node step wants it: t = plane * idir + ood (one FMA per plane, |idir| capped so
// that ood stays finite for axis-parallel rays — the arithmetic of SlabRay<float, true>), and the
// byte offsets of the near rows of a Node128 for this ray's signs.
struct TravRay {
    V3<float> o, d, idir, ood;
    uint32_t nx, ny, nz;
    RTB_WD void set(V3<float> o_, V3<float> d_) {
        o = o_;
        d = d_;
        const float eps = 1e-18f;
        const float dx = fabsf(d.x) < eps ? (d.x < 0 ? -eps : eps) : d.x;
        const float dy = fabsf(d.y) < eps ? (d.y < 0 ? -eps : eps) : d.y;
        const float dz = fabsf(d.z) < eps ? (d.z < 0 ? -eps : eps) : d.z;
        idir = V3<float>(1.0f / dx, 1.0f / dy, 1.0f / dz);
        ood = V3<float>(-(o.x * idir.x), -(o.y * idir.y), -(o.z * idir.z));
        nx = dx < 0 ? 48u : 0u;
        ny = dy < 0 ? 64u : 16u;
        nz = dz < 0 ? 80u : 32u;
    }
};

// direction octant (the window sort key)
RTB_WD uint32_t octant_of(float dx, float dy, float dz) {
    return (dx < 0 ? 1u : 0u) | (dy < 0 ? 2u : 0u) | (dz < 0 ? 4u : 0u);
}

// plane[0..3] * id + oo: two packed FFMA2 on sm_100a
RTB_WD void fma4(const Vec4f &p, float id, float oo, float out[4]) {
#ifdef __CUDA_ARCH__
    const float2 a = __ffma2_rn(make_float2(p.x, p.y), make_float2(id, id), make_float2(oo, oo));
    const float2 b = __ffma2_rn(make_float2(p.z, p.w), make_float2(id, id), make_float2(oo, oo));
    out[0] = a.x;
    out[1] = a.y;
    out[2] = b.x;
    out[3] = b.y;
#else
    out[0] = fmaf(p.x, id, oo);
    out[1] = fmaf(p.y, id, oo);
    out[2] = fmaf(p.z, id, oo);
    out[3] = fmaf(p.w, id, oo);
#endif
}

template <bool TOP> RTB_WD const char *node_address(const WideView &w, const Vec4f *s_top, uint32_t n_top, uint32_t cur) {
    if (TOP && cur < n_top)
        return reinterpret_cast<const char *>(s_top + size_t(cur) * 8);
    return reinterpret_cast<const char *>(w.nodes + size_t(cur) * 8);
}
template <bool TOP> RTB_WD Vec4f load_row(const char *p) {
#ifdef __CUDA_ARCH__
    if (!TOP) // global memory for certain: read-only path
        return __ldg(reinterpret_cast<const float4 *>(p));
#endif
    return *reinterpret_cast<const Vec4f *>(p);
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

// The state of one lane's traversal.
struct TravLane {
    TravRay r;
    V3<float> wo, wd; // the world-space ray (instances replace r by the object-space one)
    float time, t_min, t_max;
    uint32_t best, origin, cur, sp;
    Pcg rng;
};

// Next subtree off the lane's stack (skipping those behind the closest hit so far), or kDoneRef.
template <bool INST> RTB_WD void trav_pop(TravLane &L, Vec2u *stack) {
    for (;;) {
        if (L.sp == 0) {
            L.cur = kDoneRef;
            return;
        }
        const Vec2u e = stack[--L.sp];
        if (INST && e.x == kSentinelRef) { // leaving the instance: back to the world ray
            L.r.set(L.wo, L.wd);
            continue;
        }
        if (u2f(e.y) <= L.t_max) {
            L.cur = e.x;
            return;
        }
    }
}

// One 4-wide node: L.cur is an interior node on entry, the next thing to do on exit.
template <bool TOP, bool INST>
RTB_WD void trav_node_step(const WideView &w, const Vec4f *s_top, uint32_t n_top, TravLane &L, Vec2u *stack,
                           uint32_t &overflow) {
    const char *nb = node_address<TOP>(w, s_top, n_top, L.cur);
    const Vec4f nxr = load_row<TOP>(nb + L.r.nx), fxr = load_row<TOP>(nb + (48u - L.r.nx));
    const Vec4f nyr = load_row<TOP>(nb + L.r.ny), fyr = load_row<TOP>(nb + (80u - L.r.ny));
    const Vec4f nzr = load_row<TOP>(nb + L.r.nz), fzr = load_row<TOP>(nb + (112u - L.r.nz));
    const Vec4f rr = load_row<TOP>(nb + 96);
    float tnx[4], tny[4], tnz[4], tfx[4], tfy[4], tfz[4];
    fma4(nxr, L.r.idir.x, L.r.ood.x, tnx);
    fma4(fxr, L.r.idir.x, L.r.ood.x, tfx);
    fma4(nyr, L.r.idir.y, L.r.ood.y, tny);
    fma4(fyr, L.r.idir.y, L.r.ood.y, tfy);
    fma4(nzr, L.r.idir.z, L.r.ood.z, tnz);
    fma4(fzr, L.r.idir.z, L.r.ood.z, tfz);
    const float inf = Consts<float>::inf();
    float k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float tn = fmaxf(fmaxf(tnx[i], tny[i]), fmaxf(tnz[i], L.t_min));
        const float tf = fminf(fminf(tfx[i], tfy[i]), fminf(tfz[i], L.t_max));
        k[i] = tn <= tf ? tn : inf;
    }
    uint32_t r0 = f2u(rr.x), r1 = f2u(rr.y), r2 = f2u(rr.z), r3 = f2u(rr.w);
    // sorting network, ascending entry distance: (0,1) (2,3) (0,2) (1,3) (1,2)
    cmp_swap(k[0], r0, k[1], r1);
    cmp_swap(k[2], r2, k[3], r3);
    cmp_swap(k[0], r0, k[2], r2);
    cmp_swap(k[1], r1, k[3], r3);
    cmp_swap(k[1], r1, k[2], r2);
    // far to near onto the stack; the nearest is next
    if (k[1] < inf) {
        if (L.sp + 3u > uint32_t(kWideStack)) {
            overflow = 1u;
        } else {
            if (k[3] < inf)
                stack[L.sp++] = Vec2u{r3, f2u(k[3])};
            if (k[2] < inf)
                stack[L.sp++] = Vec2u{r2, f2u(k[2])};
            stack[L.sp++] = Vec2u{r1, f2u(k[1])};
        }
    }
    if (k[0] < inf)
        L.cur = r0;
    else
        trav_pop<INST>(L, stack);
}

// One primitive of the leaf L.cur (or the entry into an instance); the rest of the leaf stays in L.cur.
template <bool ANY, bool MEDIA, bool INST>
RTB_WD void trav_leaf_step(const GeomView<float> &g, const WideView &w, TravLane &L, Vec2u *stack, uint32_t &overflow,
                           uint64_t &tests) {
    const uint32_t first = L.cur & kLeafFirstMask, more = (L.cur >> 27) & 15u;
    const PrimT<float> p = g.prims[first];
    const uint32_t type = p.type_mat & PT_TYPE_MASK;
    if (INST && type == PT_INSTANCE) { // alone in its leaf (builder guarantee)
        if (L.sp + 1u > uint32_t(kWideStack)) {
            overflow = 1u;
            trav_pop<INST>(L, stack);
            return;
        }
        stack[L.sp++] = Vec2u{kSentinelRef, f2u(-Consts<float>::inf())};
        V3<float> co = L.wo, cd = L.wd;
        enter_instance<float, true>(g, int(p.aux2), co, cd);
        L.r.set(co, cd);
        L.cur = w.chain_root[p.aux2];
        if (L.cur == kEmptyRef)
            trav_pop<INST>(L, stack);
        return;
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
        h = hit_simple<float, true>(g, p, type, L.r.o, L.r.d, L.r.idir, L.time, L.t_min, L.t_max, first == L.origin, t);
    }
    if (h) {
        L.best = first;
        L.t_max = t;
        if (ANY) {
            L.cur = kDoneRef;
            return;
        }
    }
    if (more)
        L.cur = kLeafFlag | ((more - 1u) << 27) | (first + 1u);
    else
        trav_pop<INST>(L, stack);
}

// ---- window sort --------------------------------------------------------------------------------
// Counting sort of a window's entries by key: s.key[j] (written by the caller for j < cnt, all lanes
// converged) is 0..8 or kSkipKey.  At most `limit8` entries of key 8 are kept (camera samples still
// to be started); the others become kSkipKey.  Fills s.perm, returns the number of entries placed
// and, in off8, the position of the first key-8 entry.
RTB_WD uint32_t window_sort(TraceWarpSmem &s, uint32_t cnt, uint32_t limit8, uint32_t &off8) {
    const uint32_t lane = WarpOps::lane();
    if (lane < 16)
        s.hist[lane] = 0;
    WarpOps::sync();
    for (uint32_t r = 0; r < uint32_t(kTraceRounds); ++r) {
        const uint32_t j = r * 32u + lane;
        const uint32_t key = j < cnt ? s.key[j] : kSkipKey;
        const uint32_t peers = WarpOps::match_any(key);
        if (key < kSortKeys && int(lane) == WarpOps::ffs(peers) - 1)
            s.hist[key] += WarpOps::popc(peers);
        WarpOps::sync();
    }
    uint32_t off = 0, mine = 0, total = 0;
    for (uint32_t k = 0; k < kSortKeys; ++k) {
        uint32_t h = s.hist[k];
        if (k == 8 && h > limit8)
            h = limit8;
        if (k == lane)
            mine = off;
        if (k == 8)
            off8 = off;
        off += h;
    }
    total = off;
    WarpOps::sync();
    if (lane < kSortKeys)
        s.cur[lane] = mine;
    WarpOps::sync();
    for (uint32_t r = 0; r < uint32_t(kTraceRounds); ++r) {
        const uint32_t j = r * 32u + lane;
        const uint32_t key = j < cnt ? s.key[j] : kSkipKey;
        const uint32_t peers = WarpOps::match_any(key);
        const int leader = WarpOps::ffs(peers) - 1;
        uint32_t base = 0;
        if (key < kSortKeys && int(lane) == leader) {
            base = s.cur[key];
            s.cur[key] = base + WarpOps::popc(peers);
        }
        base = WarpOps::shfl(base, leader);
        const uint32_t pos = base + WarpOps::popc(peers & ((1u << lane) - 1u));
        if (key < kSortKeys) {
            if (key == 8 && pos - off8 >= limit8)
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
    uint32_t node_min = RTB_TRACE_NODE_MIN, switch_min = RTB_TRACE_SWITCH_MIN;
};
inline TraceTuning &trace_tuning() {
    static TraceTuning t;
    return t;
}
#define RTB_NODE_MIN_ (trace_tuning().node_min)
#define RTB_SWITCH_MIN_ (trace_tuning().switch_min)
#else
#define RTB_NODE_MIN_ uint32_t(RTB_TRACE_NODE_MIN)
#define RTB_SWITCH_MIN_ uint32_t(RTB_TRACE_SWITCH_MIN)
#endif
#if !defined(__CUDACC__)
// host emulation only: what the votes decided (steps) and how many lanes took part (lanes)
struct TraceSchedStats {
    uint64_t node_steps = 0, node_lanes = 0, leaf_steps = 0, leaf_lanes = 0, switches = 0, switch_lanes = 0;
};
inline TraceSchedStats &trace_sched_stats() {
    static TraceSchedStats s;
    return s;
}
#endif

// ---- the scheduler ----------------------------------------------------------------------------
// Job (all members called by the whole warp unless noted):
//   bool next_window(TraceWarpSmem&, uint32_t &count)   reserve + sort the next window; false: no work left
//   void fetch(TraceWarpSmem&, uint32_t k, TravLane&, uint32_t &tag)   [per lane] ray of window position k
//   void commit(uint32_t tag, const TravLane&)                          [per lane] its result
//   void finish_window(TraceWarpSmem&)                                   all rays of the window are committed
// counters[0] += node steps, [1] += primitive tests (per lane; the caller reduces them).
template <class Job, bool ANY, bool MEDIA, bool INST, bool TOP>
RTB_WD void warp_trace(const GeomView<float> &g, const WideView &w, const Vec4f *s_top, uint32_t n_top, TraceWarpSmem &s,
                       Job &job, uint64_t counters[2], uint32_t &overflow) {
    const uint32_t lane = WarpOps::lane();
    Vec2u stack[kWideStack];
    TravLane L;
    L.cur = kIdleRef;
    L.sp = 0;
    L.best = kNoPrim;
    L.rng.s = 0;
    uint32_t tag = 0;
    uint32_t win_count = 0, win_next = 0; // warp-uniform
    uint64_t nodes = 0, tests = 0;
#if RTB_TRACE_GUARD
    uint32_t guard = 0;
#endif
    for (;;) {
        // ---- switch point: commit finished rays, hand out the window's next ones
        if (L.cur == kDoneRef) {
            job.commit(tag, L);
            L.cur = kIdleRef;
        }
        const uint32_t idle = WarpOps::ballot(L.cur == kIdleRef);
#if !defined(__CUDACC__)
        if (lane == 0 && idle && win_next < win_count) {
            trace_sched_stats().switches += 1;
            const uint32_t left = win_count - win_next;
            trace_sched_stats().switch_lanes += WarpOps::popc(idle) < left ? WarpOps::popc(idle) : left;
        }
#endif
        if (idle && win_next < win_count) {
            const uint32_t k = win_next + WarpOps::popc(idle & ((1u << lane) - 1u));
            if (L.cur == kIdleRef && k < win_count) {
                job.fetch(s, k, L, tag);
                L.best = kNoPrim;
                L.sp = 0;
                L.r.set(L.wo, L.wd);
                L.cur = w.root_ref == kEmptyRef ? kDoneRef : w.root_ref;
            }
            const uint32_t taken = WarpOps::popc(idle);
            win_next = win_next + taken < win_count ? win_next + taken : win_count;
        }
        if (WarpOps::ballot(L.cur != kIdleRef) == 0) { // nothing in flight, nothing left to hand out
            if (win_count)
                job.finish_window(s);
            win_count = 0;
            win_next = 0;
            if (!job.next_window(s, win_count))
                break;
            continue;
        }
        // ---- walk until enough lanes have finished
        const bool can_refill = win_next < win_count;
        for (;;) {
            const uint32_t m_node = WarpOps::ballot(ref_is_node(L.cur));
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
#if !defined(__CUDACC__)
            if (lane == 0) {
                if (m_leaf == 0 || WarpOps::popc(m_node) >= RTB_NODE_MIN_) {
                    trace_sched_stats().node_steps += 1;
                    trace_sched_stats().node_lanes += WarpOps::popc(m_node);
                } else {
                    trace_sched_stats().leaf_steps += 1;
                    trace_sched_stats().leaf_lanes += WarpOps::popc(m_leaf);
                }
            }
#endif
            if (m_leaf == 0 || WarpOps::popc(m_node) >= RTB_NODE_MIN_) {
                if (ref_is_node(L.cur)) {
                    ++nodes;
                    trav_node_step<TOP, INST>(w, s_top, n_top, L, stack, overflow);
                }
            } else {
                if (ref_is_leaf(L.cur))
                    trav_leaf_step<ANY, MEDIA, INST>(g, w, L, stack, overflow, tests);
            }
        }
    }
    counters[0] += nodes;
    counters[1] += tests;
}

// ---- batch job: rays in, (t, primitive) out — rtb_trace_batch precision 34 and the CPU suite ----
struct BatchTraceJob {
    const Vec4f *ray_a; // origin, time
    const Vec4f *ray_b; // direction, origin primitive (sorted index or kNoPrim)
    const Vec2f *ray_t; // t_min, t_max
    Vec2f *out;         // t, primitive (sorted index, as bits)
    uint32_t n;
    uint32_t *head; // work cursor (global)
    uint32_t base;  // first entry of the current window
    uint64_t seed;
    RTB_WD bool next_window(TraceWarpSmem &s, uint32_t &count) {
        const uint32_t lane = WarpOps::lane();
        uint32_t b = 0;
        if (lane == 0)
            b = WarpOps::atomic_add(head, uint32_t(kTraceWindow));
        b = WarpOps::shfl(b, 0);
        if (b >= n)
            return false;
        base = b;
        const uint32_t cnt = n - b < uint32_t(kTraceWindow) ? n - b : uint32_t(kTraceWindow);
        for (uint32_t r = 0; r < uint32_t(kTraceRounds); ++r) {
            const uint32_t j = r * 32u + lane;
            if (j < cnt) {
                const Vec4f d = ray_b[b + j];
                s.key[j] = uint8_t(RTB_TRACE_SORT ? octant_of(d.x, d.y, d.z) : 0u);
            }
        }
        WarpOps::sync();
        uint32_t off8;
        count = window_sort(s, cnt, 0u, off8);
        return true;
    }
    RTB_WD void fetch(TraceWarpSmem &s, uint32_t k, TravLane &L, uint32_t &tag) {
        tag = base + s.perm[k];
        const Vec4f a = ray_a[tag], b = ray_b[tag];
        const Vec2f t = ray_t[tag];
        L.wo = V3<float>(a.x, a.y, a.z);
        L.wd = V3<float>(b.x, b.y, b.z);
        L.time = a.w;
        L.origin = f2u(b.w);
        L.t_min = t.x;
        L.t_max = t.y;
        L.rng = pcg_seed(tag, seed);
    }
    RTB_WD void commit(uint32_t tag, const TravLane &L) { out[tag] = Vec2f{L.t_max, u2f(L.best)}; }
    RTB_WD void finish_window(TraceWarpSmem &) {}
};

} // namespace rtb

#endif // RTB_TRACE_CUH
