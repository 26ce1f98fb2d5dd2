// rtb_bvh.hpp — host-side construction of the flattened two-level BVH.
//
// Replaces the reference's bvh_node constructor (src/geometry/bvh.h:52-94: random
// split axis, std::sort with a virtual bounding_box() per comparison, a full copy
// of the object vector per node — super-quadratic).  Here: binned SAH, O(N log N),
// breadth-first node order (so the top levels are a contiguous prefix that the
// traversal kernels stage in shared memory), sibling pairs adjacent and 64-byte
// aligned, leaves referencing contiguous ranges of a leaf-ordered primitive array.
//
// Closest-hit results do not depend on tree topology, so the reference's random
// topology does not need to be reproduced.
#ifndef RTB_BVH_HPP
#define RTB_BVH_HPP

#include "rtb_geom.cuh"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace rtb {

struct Box {
    double lo[3] = {INFINITY, INFINITY, INFINITY};
    double hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    void grow(const double p[3]) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], p[k]);
            hi[k] = std::max(hi[k], p[k]);
        }
    }
    void grow(const Box &b) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], b.lo[k]);
            hi[k] = std::max(hi[k], b.hi[k]);
        }
    }
    bool valid() const { return lo[0] <= hi[0] && lo[1] <= hi[1] && lo[2] <= hi[2]; }
    double area() const {
        if (!valid())
            return 0;
        const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return 2 * (dx * dy + dy * dz + dz * dx);
    }
    double centre(int k) const { return 0.5 * (lo[k] + hi[k]); }
};

// A thing the builder places in a leaf: `id` is opaque to the builder.
struct BuildItem {
    Box box;
    uint32_t id;
    bool solitary; // must be alone in its leaf (instances)
};

struct BuildResult {
    std::vector<Node32> nodes;   // nodes[0] is the root; children pairs at even indices
    std::vector<uint32_t> order; // leaf-ordered item ids; leaves index into this
};

inline float round_down(double x) {
    float f = float(x);
    if (double(f) > x)
        f = std::nextafterf(f, -INFINITY);
    return f;
}
inline float round_up(double x) {
    float f = float(x);
    if (double(f) < x)
        f = std::nextafterf(f, INFINITY);
    return f;
}

// fp64 box -> conservative fp32 node bounds (outward rounding + a small pad that
// covers the fp32 slab-test rounding of the production path).
inline void store_bounds(Node32 &n, const Box &b) {
    for (int k = 0; k < 3; ++k) {
        const double ext = std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k]));
        const double pad = 1e-6 * ext + 1e-7;
        n.lo[k] = round_down(b.lo[k] - pad);
        n.hi[k] = round_up(b.hi[k] + pad);
    }
}

// Binned-SAH build.  `first_offset` is added to leaf `first` fields (the position
// of this tree's items in the global primitive array), `node_offset` to child
// indices (the position of this tree in the global node array; must be even).
inline BuildResult build_bvh(const std::vector<BuildItem> &items, int max_leaf, uint32_t first_offset,
                             uint32_t node_offset) {
    constexpr int kBins = 16;
    constexpr int kMaxDepth = 40;
    max_leaf = std::min(std::max(max_leaf, 1), kMaxLeafPrims);
    BuildResult out;
    const uint32_t n = uint32_t(items.size());
    std::vector<uint32_t> idx(n);
    for (uint32_t i = 0; i < n; ++i)
        idx[i] = i;
    struct Task {
        uint32_t node, begin, end;
        int depth;
    };
    out.nodes.resize(2); // slot 1 pads the root so that child pairs stay even-aligned
    std::memset(out.nodes.data(), 0, 2 * sizeof(Node32));
    std::vector<Task> level{{0, 0, n, 0}}, next;
    auto make_leaf = [&](const Task &t, const Box &b) {
        Node32 &nd = out.nodes[t.node];
        store_bounds(nd, b);
        const uint32_t first = first_offset + t.begin, count = t.end - t.begin;
        if (count < 1 || count > uint32_t(kMaxLeafPrims) || first > kLeafFirstMask - 2)
            throw std::runtime_error("bvh: leaf does not fit the node reference encoding");
        nd.ref = kLeafFlag | ((count - 1) << 27) | first;
        nd.count = count;
    };
    if (n == 0) { // empty tree: the root is a leaf over nothing
        Box b;
        for (int k = 0; k < 3; ++k)
            b.lo[k] = b.hi[k] = 0;
        store_bounds(out.nodes[0], b);
        out.nodes[0].ref = kEmptyRef;
        out.nodes[0].count = 0;
        return out;
    }
    while (!level.empty()) {
        next.clear();
        for (const Task &t : level) {
            Box b, cb;
            bool has_solitary = false;
            for (uint32_t i = t.begin; i < t.end; ++i) {
                const BuildItem &it = items[idx[i]];
                b.grow(it.box);
                double c[3] = {it.box.centre(0), it.box.centre(1), it.box.centre(2)};
                cb.grow(c);
                has_solitary |= it.solitary;
            }
            const uint32_t cnt = t.end - t.begin;
            if (cnt == 1) {
                make_leaf(t, b);
                continue;
            }
            // choose split: binned SAH over the 3 axes
            double best_cost = INFINITY;
            int best_axis = -1, best_bin = -1;
            for (int ax = 0; ax < 3; ++ax) {
                const double lo = cb.lo[ax], ext = cb.hi[ax] - cb.lo[ax];
                if (!(ext > 0))
                    continue;
                Box bb[kBins];
                uint32_t bc[kBins] = {0};
                for (uint32_t i = t.begin; i < t.end; ++i) {
                    const BuildItem &it = items[idx[i]];
                    int bi = int(kBins * ((it.box.centre(ax) - lo) / ext));
                    bi = std::min(std::max(bi, 0), kBins - 1);
                    bb[bi].grow(it.box);
                    bc[bi]++;
                }
                double la[kBins - 1], ra[kBins - 1];
                uint32_t lc[kBins - 1], rc[kBins - 1];
                Box acc;
                uint32_t c = 0;
                for (int i = 0; i < kBins - 1; ++i) {
                    acc.grow(bb[i]);
                    c += bc[i];
                    la[i] = acc.area();
                    lc[i] = c;
                }
                acc = Box();
                c = 0;
                for (int i = kBins - 1; i > 0; --i) {
                    acc.grow(bb[i]);
                    c += bc[i];
                    ra[i - 1] = acc.area();
                    rc[i - 1] = c;
                }
                for (int i = 0; i < kBins - 1; ++i) {
                    if (lc[i] == 0 || rc[i] == 0)
                        continue;
                    const double cost = la[i] * lc[i] + ra[i] * rc[i];
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_axis = ax;
                        best_bin = i;
                    }
                }
            }
            const double leaf_cost = b.area() * cnt;
            const bool can_leaf = cnt <= uint32_t(max_leaf) && !has_solitary;
            // traversal step ~ one primitive test
            if (can_leaf && (best_axis < 0 || best_cost + b.area() >= leaf_cost)) {
                make_leaf(t, b);
                continue;
            }
            uint32_t mid;
            if (best_axis >= 0 && t.depth < kMaxDepth) {
                const double lo = cb.lo[best_axis], ext = cb.hi[best_axis] - cb.lo[best_axis];
                auto it = std::partition(idx.begin() + t.begin, idx.begin() + t.end, [&](uint32_t id) {
                    int bi = int(kBins * ((items[id].box.centre(best_axis) - lo) / ext));
                    bi = std::min(std::max(bi, 0), kBins - 1);
                    return bi <= best_bin;
                });
                mid = uint32_t(it - idx.begin());
            } else {
                // coincident centroids or depth cap: median split in index order
                mid = t.begin + cnt / 2;
            }
            if (mid == t.begin || mid == t.end)
                mid = t.begin + cnt / 2;
            const uint32_t child = uint32_t(out.nodes.size());
            out.nodes.resize(child + 2);
            Node32 &nd = out.nodes[t.node];
            store_bounds(nd, b);
            nd.ref = node_offset + child;
            nd.count = 0;
            next.push_back({child, t.begin, mid, t.depth + 1});
            next.push_back({child + 1, mid, t.end, t.depth + 1});
        }
        level.swap(next);
    }
    out.order.resize(n);
    for (uint32_t i = 0; i < n; ++i)
        out.order[i] = items[idx[i]].id;
    return out;
}

} // namespace rtb

#endif // RTB_BVH_HPP
