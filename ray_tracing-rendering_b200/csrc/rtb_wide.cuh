// rtb_wide.cuh — the 4-wide BVH the production traversal kernels walk (rtb_trace.cuh): the binary
// SAH tree of rtb_bvh.hpp collapsed two levels at a time into 128-byte nodes, laid out breadth
// first, and a scalar reference traversal over it.
//
// Why wide: the binary while-while traversal of round 1 spent one dependent 64-byte fetch per two
// boxes and ran at 12-18 of 32 lanes (ncu).  A 128-byte node brings four
// boxes per dependent fetch (half the steps per ray: 7.1 instead of 15.2 on the scene09 rays,
// 10.4 instead of 21.7 on the 1 M-sphere field), its structure-of-arrays rows feed the packed
// FFMA2 slab test of sm_100a two children at a time, and the breadth-first order makes the top
// levels of the tree a contiguous prefix that the kernels stage in shared memory.
//
// Reference counterpart: bvh_node::hit (src/geometry/bvh.h:40-50), aabb::hit (aabb.h:31-48).
#ifndef RTB_WIDE_CUH
#define RTB_WIDE_CUH

#include "rtb_bvh.hpp"
#include "rtb_geom.cuh"

#include <cmath>
#include <limits>
#include <vector>

namespace rtb {

// 128 bytes = eight 16-byte rows: lo.x[4] lo.y[4] lo.z[4] hi.x[4] hi.y[4] hi.z[4] ref[4] pad[4].
// A ray reads the "near" row of an axis at byte offset axis*16 + (dir < 0 ? 48 : 0) and the "far"
// row at the other one, so no min/max is needed to order the two planes of a slab.
// ref[i] in the encoding of Node32::ref: an interior child is an index into the Node128 array, a
// leaf child is the binary tree's own leaf ref (kLeafFlag | count | [kLeafInstanceFlag] | first).
// Unused slots hold an inverted box (lo = +inf, hi = -inf), which no ray enters, and kEmptyRef.
struct alignas(128) Node128 {
    float lo[3][4];
    float hi[3][4];
    uint32_t ref[4];
    uint32_t pad[4];
};
static_assert(sizeof(Node128) == 128, "wide BVH node must be 128 bytes");

// Traversal stack of the kernels in rtb_trace.cuh: entries (ref, entry distance) per lane.  A
// descent pushes at most three siblings per level and one marker per instance entered; the scene
// upload refuses a tree that could need more (wide_stack_need).
constexpr int kWideStack = 64;
constexpr int kMaxGlobalPrims = 4; // primitives left out of the wide tree and tested for every ray

// The node the kernels fetch: the same four children in 64 bytes.  Child boxes are quantised to 8 bits
// per plane on a grid laid over the node's own box (origin o, cell size s per axis, plane = o + q s),
// rounded outwards, so a ray that enters a child's exact box always enters its quantised one; the
// primitive tests are unchanged, hence so is every hit.  Why: ncu showed the fp32 128-byte node
// binding on the L1 data pipe (seven 16-byte rows = seven wavefronts per lane and step, 78 % of
// the L1 cycles on the 1 M-sphere field); four rows cut the wavefronts, the tree's cache footprint
// and the registers a step holds in flight by 40-50 %.
//   row 0: o.x o.y o.z s.x          row 1: lo.x[4] lo.y[4] lo.z[4] hi.x[4]   (bytes, child 0..3)
//   row 2: hi.y[4] hi.z[4] s.y s.z  row 3: ref[4]
// Unused slots: lo = 255, hi = 0 (inverted: no ray enters) and kEmptyRef.
struct alignas(64) QNode64 {
    float ox, oy, oz, sx;
    uint32_t lox, loy, loz, hix;
    uint32_t hiy, hiz;
    float sy, sz;
    uint32_t ref[4];
};
static_assert(sizeof(QNode64) == 64, "quantised wide BVH node must be 64 bytes");

struct WideTree {
    std::vector<QNode64> qnodes;      // nodes[] quantised: what the device holds
    std::vector<Node128> nodes;       // [top-level tree, breadth first][bottom-level tree of chain 0]...
    uint32_t root_ref = kEmptyRef;    // of the top-level tree
    std::vector<uint32_t> chain_root; // per wrapper chain: root ref of that instance's bottom-level tree (kEmptyRef: none)
    uint64_t n_children = 0;          // occupied child slots (n_children / nodes.size() = mean arity)
    std::vector<uint32_t> global_prims; // sorted primitive indices left OUT of the tree (see build_wide): tested for every ray
    uint32_t n_top_nodes = 0;         // nodes of the top-level tree = nodes[0 .. n_top_nodes)
    int max_depth = 0;                // deepest wide node (root = 1), top level + bottom level
};

// Collapses the binary tree under `ref2` (Node32 refs are global indices into `nodes2`) and
// appends its nodes to `out` in breadth-first order.  A node keeps absorbing the children of its
// largest interior child (by surface area) until it has four children or only leaves.
// `skip`: sorted primitive indices whose (single-primitive) leaves are left out of the wide tree.
inline uint32_t collapse_wide(const std::vector<Node32> &nodes2, uint32_t ref2, WideTree &out, int *depth_out = nullptr,
                              const std::vector<uint32_t> *skip = nullptr) {
    if (depth_out)
        *depth_out = 0;
    auto skipped = [&](uint32_t ref) {
        if (!skip || !(ref & kLeafFlag) || ref == kEmptyRef || ((ref >> 27) & 15u) != 0)
            return false;
        const uint32_t first = ref & kLeafFirstMask;
        for (uint32_t g : *skip)
            if (g == first)
                return true;
        return false;
    };
    if (ref2 & kLeafFlag) // leaves (and kEmptyRef) keep their ref
        return skipped(ref2) ? kEmptyRef : ref2;
    struct Child {
        float lo[3], hi[3];
        uint32_t ref;
    };
    auto child_of = [&](uint32_t i) {
        Child c;
        for (int k = 0; k < 3; ++k) {
            c.lo[k] = nodes2[i].lo[k];
            c.hi[k] = nodes2[i].hi[k];
        }
        c.ref = nodes2[i].ref;
        return c;
    };
    auto area = [](const Child &c) {
        const double x = double(c.hi[0]) - c.lo[0], y = double(c.hi[1]) - c.lo[1], z = double(c.hi[2]) - c.lo[2];
        return x * y + y * z + z * x;
    };
    struct Pending {
        uint32_t ref2, self;
        int depth;
    };
    const uint32_t root = uint32_t(out.nodes.size());
    out.nodes.emplace_back();
    std::vector<Pending> queue{{ref2, root, 1}};
    const float inf = std::numeric_limits<float>::infinity();
    int deepest = 0;
    for (size_t head = 0; head < queue.size(); ++head) {
        const Pending cur = queue[head];
        deepest = cur.depth > deepest ? cur.depth : deepest;
        Child ch[5] = {child_of(cur.ref2), child_of(cur.ref2 + 1)};
        int n = 2;
        for (;;) {
            for (int i = 0; i < n;) // drop the leaves of left-out primitives
                if (skipped(ch[i].ref))
                    ch[i] = ch[--n];
                else
                    ++i;
            if (n >= 4)
                break;
            int pick = -1;
            for (int i = 0; i < n; ++i)
                if (!(ch[i].ref & kLeafFlag) && (pick < 0 || area(ch[i]) > area(ch[pick])))
                    pick = i;
            if (pick < 0)
                break;
            const uint32_t r = ch[pick].ref;
            ch[pick] = child_of(r);
            ch[n++] = child_of(r + 1);
        }
        Node128 node;
        for (int i = 0; i < 4; ++i) {
            for (int k = 0; k < 3; ++k) {
                node.lo[k][i] = i < n ? ch[i].lo[k] : inf;
                node.hi[k][i] = i < n ? ch[i].hi[k] : -inf;
            }
            node.ref[i] = kEmptyRef;
            node.pad[i] = 0;
        }
        for (int i = 0; i < n; ++i) {
            if (ch[i].ref & kLeafFlag) {
                node.ref[i] = ch[i].ref;
            } else { // interior: its node is appended now, i.e. level by level
                node.ref[i] = uint32_t(out.nodes.size());
                out.nodes.emplace_back();
                queue.push_back({ch[i].ref, node.ref[i], cur.depth + 1});
            }
        }
        out.nodes[cur.self] = node;
        out.n_children += uint64_t(n);
    }
    if (depth_out)
        *depth_out = deepest;
    return root;
}

inline void quantize_wide(WideTree &w);

// The whole scene: the top-level tree first, then the bottom-level tree of every instance record.
// `globals`: top-level primitives (sorted indices, each alone in its leaf) whose boxes dwarf the rest of
// the scene — a ground sphere of radius 1000 under a field of spheres of radius 0.2, the boundary of
// a fog that fills the world.  Every ray meets their box anyway, so the kernels test them once, up
// front, with all refilled lanes in step, and the tree (and its quantisation grids, which span a
// node's children) is built over the remaining primitives only.
template <class R>
inline WideTree build_wide(const std::vector<Node32> &nodes2, uint32_t root_ref2, const PrimT<R> *prims, size_t n_prims,
                           size_t n_chains, const std::vector<uint32_t> &globals = std::vector<uint32_t>()) {
    WideTree w;
    w.global_prims = globals;
    int top_depth = 0, deepest_bottom = 0;
    w.root_ref = collapse_wide(nodes2, root_ref2, w, &top_depth, globals.empty() ? nullptr : &globals);
    w.n_top_nodes = uint32_t(w.nodes.size());
    w.chain_root.assign(n_chains, kEmptyRef);
    for (size_t i = 0; i < n_prims; ++i)
        if ((prims[i].type_mat & PT_TYPE_MASK) == PT_INSTANCE && prims[i].aux2 < n_chains) {
            int d = 0;
            w.chain_root[prims[i].aux2] = collapse_wide(nodes2, prims[i].aux, w, &d);
            deepest_bottom = d > deepest_bottom ? d : deepest_bottom;
        }
    w.max_depth = top_depth + deepest_bottom;
    quantize_wide(w);
    return w;
}

inline int wide_stack_need(const WideTree &w) { return 3 * w.max_depth + 2; }

// nodes[] -> qnodes[].  The grid of a node spans the union of its children's (already padded) boxes;
// a quarter cell of slack on top of the outward rounding covers the rounding of the kernel's decode
// (qnode_slabs: four fp32 roundings on values of the size of t).
inline void quantize_wide(WideTree &w) {
    w.qnodes.resize(w.nodes.size());
    parallel_for(w.nodes.size(), 4096, w.nodes.size() >= 65536 ? builder_threads() : 1, [&](size_t n) {
        const Node128 &src = w.nodes[n];
        QNode64 q;
        float o[3], s[3];
        for (int k = 0; k < 3; ++k) {
            float lo = std::numeric_limits<float>::infinity(), hi = -lo;
            for (int i = 0; i < 4; ++i)
                if (src.ref[i] != kEmptyRef) {
                    lo = src.lo[k][i] < lo ? src.lo[k][i] : lo;
                    hi = src.hi[k][i] > hi ? src.hi[k][i] : hi;
                }
            if (!(lo <= hi))
                lo = hi = 0.f;
            o[k] = lo;
            const double cell = (double(hi) - double(lo)) / 254.0; // one cell of head room at the top
            float sf = float(cell);
            if (double(sf) < cell)
                sf = std::nextafterf(sf, std::numeric_limits<float>::infinity());
            if (!(sf > 0.f))
                sf = 1e-30f;
            s[k] = sf;
        }
        uint32_t lo_w[3] = {0, 0, 0}, hi_w[3] = {0, 0, 0};
        for (int i = 0; i < 4; ++i)
            for (int k = 0; k < 3; ++k) {
                uint32_t ql = 255, qh = 0;
                if (src.ref[i] != kEmptyRef) {
                    const double a = (double(src.lo[k][i]) - double(o[k])) / double(s[k]) - 0.25;
                    const double b = (double(src.hi[k][i]) - double(o[k])) / double(s[k]) + 0.25;
                    const double fl = std::floor(a), ce = std::ceil(b);
                    ql = uint32_t(fl < 0 ? 0 : (fl > 255 ? 255 : fl));
                    qh = uint32_t(ce < 0 ? 0 : (ce > 255 ? 255 : ce));
                }
                lo_w[k] |= ql << (8 * i);
                hi_w[k] |= qh << (8 * i);
            }
        q.ox = o[0];
        q.oy = o[1];
        q.oz = o[2];
        q.sx = s[0];
        q.sy = s[1];
        q.sz = s[2];
        q.lox = lo_w[0];
        q.loy = lo_w[1];
        q.loz = lo_w[2];
        q.hix = hi_w[0];
        q.hiy = hi_w[1];
        q.hiz = hi_w[2];
        for (int i = 0; i < 4; ++i)
            q.ref[i] = src.ref[i];
        w.qnodes[n] = q;
    });
}

// Scalar traversal over the 4-wide tree: what the warp-scheduled kernels of rtb_trace.cuh compute
// per ray, written as one loop (the CPU suite holds it to the reference's hits bit for bit in fp64
// and holds the kernels' scheduler to it).  The descent step tests four boxes, continues with the
// nearest child hit and pushes the others far to near; leaves are processed exactly as in the
// binary traversal, so the set of primitives tested against a ray's shrinking [t_min, t_max]
// gives the same closest hit (ties between coincident surfaces aside, as between any two orders).
template <class R, bool ANY, bool ROBUST, class Rng, class Stack, bool MEDIA = true>
RTB_HD uint32_t traverse_wide(const GeomView<R> &g, const Node128 *nodes4, uint32_t root_ref, const uint32_t *chain_root, V3<R> o,
                              V3<R> d, R time, R t_min, R t_max, uint32_t origin_prim, Rng &rng, R &t_hit,
                              uint64_t *n_nodes, uint64_t *n_tests, Stack &stack, const uint32_t *globals = nullptr,
                              uint32_t n_globals = 0) {
    uint32_t best = kNoPrim;
    V3<R> co = o, cd = d; // current-level ray
    SlabRay<R, ROBUST> sr;
    sr.set(o, d);
    // the primitives kept out of the tree (build_wide): tested first, for every ray
    for (uint32_t gi = 0; gi < n_globals; ++gi) {
        const uint32_t i = globals[gi];
        const PrimT<R> p = g.prims[i];
        const uint32_t type = p.type_mat & PT_TYPE_MASK;
        if (!ROBUST && g.prim_orig[i] >= g.orig_limit)
            continue;
        if (n_tests)
            ++*n_tests;
        R t;
        bool h;
        if (MEDIA && type == PT_MEDIUM) {
            h = hit_medium<R, ROBUST>(g, p, o, d, time, t_min, t_max, rng(), t);
            if (p.type_mat & PT_DUP_LEAF) {
                R t2;
                if (hit_medium<R, ROBUST>(g, p, o, d, time, t_min, h ? t : t_max, rng(), t2)) {
                    h = true;
                    t = t2;
                }
            }
        } else {
            h = hit_simple<R, ROBUST>(g, p, type, o, d, safe_inv(d), time, t_min, t_max, ROBUST && i == origin_prim, t);
        }
        if (!ROBUST && h && best != kNoPrim && t == t_max && g.prim_orig[i] < g.prim_orig[best])
            h = false;
        if (h) {
            best = i;
            t_max = t;
            if (ANY) {
                t_hit = t;
                return best;
            }
        }
    }
    uint32_t cur = root_ref;
    if (cur == kEmptyRef) {
        t_hit = t_max;
        return best;
    }
    while (true) {
        while (!(cur & kLeafFlag)) { // descend
            const Node128 &n = nodes4[cur];
            if (n_nodes)
                *n_nodes += 1;
            R tn[4];
            uint32_t ref[4];
            int hits = 0;
            for (int i = 0; i < 4; ++i) {
                const R x0 = sr.plane(R(n.lo[0][i]), sr.idir.x, sr.ood.x), x1 = sr.plane(R(n.hi[0][i]), sr.idir.x, sr.ood.x);
                const R y0 = sr.plane(R(n.lo[1][i]), sr.idir.y, sr.ood.y), y1 = sr.plane(R(n.hi[1][i]), sr.idir.y, sr.ood.y);
                const R z0 = sr.plane(R(n.lo[2][i]), sr.idir.z, sr.ood.z), z1 = sr.plane(R(n.hi[2][i]), sr.idir.z, sr.ood.z);
                const R near = fmax_(fmax_(fmin_(x0, x1), fmin_(y0, y1)), fmax_(fmin_(z0, z1), t_min));
                const R far = fmin_(fmin_(fmax_(x0, x1), fmax_(y0, y1)), fmin_(fmax_(z0, z1), t_max));
                if (near <= far && n.ref[i] != kEmptyRef) { // insertion by distance, farthest first
                    int k = hits++;
                    for (; k > 0 && tn[k - 1] < near; --k) {
                        tn[k] = tn[k - 1];
                        ref[k] = ref[k - 1];
                    }
                    tn[k] = near;
                    ref[k] = n.ref[i];
                }
            }
            if (hits == 0) {
                if (stack.empty()) {
                    t_hit = t_max;
                    return best;
                }
                cur = stack.pop();
                continue;
            }
            for (int k = 0; k + 1 < hits; ++k)
                stack.push(ref[k]);
            cur = ref[hits - 1]; // the nearest
        }
        bool entered = false;
        if (cur == kSentinelRef) { // leaving the instance: back to the world ray
            co = o;
            cd = d;
            sr.set(o, d);
        } else if (cur != kEmptyRef) {
            const uint32_t first = cur & kLeafFirstMask, last = first + ((cur >> 27) & 15u) + 1u;
            V3<R> cid = safe_inv(cd);
            for (uint32_t i = first; i < last; ++i) {
                const PrimT<R> p = g.prims[i];
                const uint32_t type = p.type_mat & PT_TYPE_MASK;
                if (type == PT_INSTANCE) { // builder guarantee: an instance is alone in its leaf, top level only
                    stack.push(kSentinelRef);
                    enter_instance<R, ROBUST>(g, int(p.aux2), co, cd);
                    sr.set(co, cd);
                    cur = chain_root[p.aux2];
                    entered = true;
                    break;
                }
                if (type == PT_BOX) {
                    if (box_slot<R, ANY, ROBUST>(g, p, co, cd, cid, time, t_min, t_max, origin_prim, best, n_tests)) {
                        t_hit = t_max;
                        return best;
                    }
                    continue;
                }
                if (!ROBUST && g.prim_orig[i] >= g.orig_limit)
                    continue;
                if (n_tests)
                    ++*n_tests;
                R t;
                bool h;
                if (MEDIA && type == PT_MEDIUM) {
                    h = hit_medium<R, ROBUST>(g, p, co, cd, time, t_min, t_max, rng(), t);
                    if (p.type_mat & PT_DUP_LEAF) { // second visit of a one-object bvh_node (bvh.h:46-47)
                        R t2;
                        if (hit_medium<R, ROBUST>(g, p, co, cd, time, t_min, h ? t : t_max, rng(), t2)) {
                            h = true;
                            t = t2;
                        }
                    }
                } else {
                    h = hit_simple<R, ROBUST>(g, p, type, co, cd, cid, time, t_min, t_max, ROBUST && i == origin_prim, t);
                }
                // a later primitive of the reference's walk wins a tie (its tests reject t > t_max only)
                if (!ROBUST && h && best != kNoPrim && t == t_max && g.prim_orig[i] < g.prim_orig[best])
                    h = false;
                if (h) {
                    best = i;
                    t_max = t;
                    if (ANY) {
                        t_hit = t;
                        return best;
                    }
                }
            }
        }
        if (entered)
            continue;
        if (stack.empty()) {
            t_hit = t_max;
            return best;
        }
        cur = stack.pop();
    }
}

} // namespace rtb

#endif // RTB_WIDE_CUH
