// rtb_wide.cuh — 4-wide BVH: the binary tree of rtb_bvh.hpp collapsed level pairs at a time, and a
// traversal over it.  GROUNDWORK for the next round: the binary
// while-while traversal of k_extend runs a lean loop at 9-18 of 32 lanes, i.e. what it loses is
// lane utilisation, and a wider node halves the number of dependent steps per ray (one 128-byte
// fetch and four slab tests per step instead of 64 bytes and two).  Nothing in librtb200.so includes
// this header yet: it is instantiated by tests/hostcheck only, where the 4-wide traversal is held
// to the same answers as the binary one (fp64: the reference's t and primitive bit for bit).
//
// Reference counterpart: bvh_node::hit (src/geometry/bvh.h:40-50).
#ifndef RTB_WIDE_CUH
#define RTB_WIDE_CUH

#include "rtb_geom.cuh"

#include <limits>
#include <vector>

namespace rtb {

// 128 bytes: up to four children, boxes as structure of arrays (one float4 per plane), and their
// refs in the encoding of Node32::ref — an interior child is an index into the Node128 array, a
// leaf child is the binary tree's own leaf ref (kLeafFlag | count | [kLeafInstanceFlag] | first).
// Unused slots hold an inverted box, which no ray enters.
struct alignas(16) Node128 {
    float lo[3][4];
    float hi[3][4];
    uint32_t ref[4];
    uint32_t pad[4];
};
static_assert(sizeof(Node128) == 128, "wide BVH node must be 128 bytes");

struct WideTree {
    std::vector<Node128> nodes;
    uint32_t root_ref = kEmptyRef;    // of the top-level tree
    std::vector<uint32_t> prim_root;  // per sorted primitive: root ref of an instance's bottom-level tree
    uint64_t n_children = 0;          // occupied child slots (n_children / nodes.size() = mean arity)
};

// Collapses the binary tree rooted at the child pair `ref2` (Node32 refs are global indices into
// `nodes2`, top level and bottom levels alike).  A node keeps absorbing the children of its
// largest interior child (by surface area) until it has four children or only leaves.
inline uint32_t collapse_wide(const std::vector<Node32> &nodes2, uint32_t ref2, WideTree &out) {
    if (ref2 & kLeafFlag) // leaves (and kEmptyRef) keep their ref
        return ref2;
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
    Child ch[4] = {child_of(ref2), child_of(ref2 + 1)};
    int n = 2;
    while (n < 4) {
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
    const uint32_t self = uint32_t(out.nodes.size());
    out.nodes.emplace_back();
    Node128 node;
    const float inf = std::numeric_limits<float>::infinity();
    for (int i = 0; i < 4; ++i) {
        for (int k = 0; k < 3; ++k) {
            node.lo[k][i] = i < n ? ch[i].lo[k] : inf;
            node.hi[k][i] = i < n ? ch[i].hi[k] : -inf;
        }
        node.ref[i] = kEmptyRef;
        node.pad[i] = 0;
    }
    for (int i = 0; i < n; ++i)
        node.ref[i] = collapse_wide(nodes2, ch[i].ref, out); // may grow out.nodes: `node` is a local copy
    out.nodes[self] = node;
    out.n_children += uint64_t(n);
    return self;
}

// The whole scene: the top-level tree and the bottom-level tree of every instance record.
template <class R> inline WideTree build_wide(const std::vector<Node32> &nodes2, uint32_t root_ref2, const PrimT<R> *prims, size_t n_prims) {
    WideTree w;
    w.root_ref = collapse_wide(nodes2, root_ref2, w);
    w.prim_root.assign(n_prims, kEmptyRef);
    for (size_t i = 0; i < n_prims; ++i)
        if ((prims[i].type_mat & PT_TYPE_MASK) == PT_INSTANCE)
            w.prim_root[i] = collapse_wide(nodes2, prims[i].aux, w);
    return w;
}

// traverse() of rtb_geom.cuh over the 4-wide tree (instance entry / exit in the leaf phase).  The
// descent step tests four boxes, continues with the nearest child hit and pushes the others far
// to near; leaves are processed exactly as in the binary traversal, so the set of primitives
// tested against a ray's shrinking [t_min, t_max] gives the same closest hit (ties between
// coincident surfaces aside, as between any two traversal orders).
template <class R, bool ANY, bool ROBUST, class Rng, class Stack, bool MEDIA = true>
RTB_HD uint32_t traverse_wide(const GeomView<R> &g, const Node128 *nodes4, uint32_t root_ref, const uint32_t *prim_root, V3<R> o,
                              V3<R> d, R time, R t_min, R t_max, uint32_t origin_prim, Rng &rng, R &t_hit,
                              uint64_t *n_nodes, uint64_t *n_tests, Stack &stack) {
    uint32_t best = kNoPrim;
    V3<R> co = o, cd = d; // current-level ray
    SlabRay<R, ROBUST> sr;
    sr.set(o, d);
    uint32_t cur = root_ref;
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
                    cur = prim_root[i];
                    entered = true;
                    break;
                }
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
