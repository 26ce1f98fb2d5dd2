// rtb_bvh.hpp — host-side construction of the flattened two-level BVH.
//
// Replaces the reference's bvh_node constructor (src/geometry/bvh.h:52-94: random
// split axis, std::sort with a virtual bounding_box() per comparison, a full copy
// of the object vector per node — super-quadratic).  Here: binned SAH, O(N log N), multi-threaded,
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
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <mutex>
#include <stdexcept>
#include <thread>
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
    bool solitary;         // must be alone in its leaf (instances, primitives kept out of the wide tree)
    bool instance = false; // the item is an instance record: its leaf ref carries kLeafInstanceFlag
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

// ---- a minimal fork-join helper (the builder is the only multi-threaded host code) ------------
// Runs body(i) for i in [0, n) on up to `threads` std::threads, handing out indices in blocks of
// `grain` through one atomic counter.  Exceptions from body are rethrown on the caller.
template <class F> inline void parallel_for(size_t n, size_t grain, unsigned threads, F body) {
    if (n == 0)
        return;
    const size_t blocks = (n + grain - 1) / grain;
    threads = unsigned(std::min<size_t>(threads, blocks));
    if (threads <= 1) {
        for (size_t i = 0; i < n; ++i)
            body(i);
        return;
    }
    std::atomic<size_t> next{0};
    std::exception_ptr err;
    std::mutex err_mu;
    auto work = [&]() {
        try {
            for (;;) {
                const size_t b0 = next.fetch_add(1);
                if (b0 >= blocks)
                    return;
                const size_t lo = b0 * grain, hi = std::min(n, lo + grain);
                for (size_t i = lo; i < hi; ++i)
                    body(i);
            }
        } catch (...) {
            std::lock_guard<std::mutex> g(err_mu);
            if (!err)
                err = std::current_exception();
            next.store(blocks);
        }
    };
    std::vector<std::thread> pool;
    pool.reserve(threads - 1);
    for (unsigned t = 1; t < threads; ++t)
        pool.emplace_back(work);
    work();
    for (auto &t : pool)
        t.join();
    if (err)
        std::rethrow_exception(err);
}

// RTB200_BUILD_THREADS overrides the thread count (tests compare trees across counts).
inline unsigned builder_threads() {
    if (const char *e = std::getenv("RTB200_BUILD_THREADS")) {
        const int v = std::atoi(e);
        if (v >= 1)
            return unsigned(std::min(v, 64));
    }
    const unsigned hc = std::thread::hardware_concurrency();
    return std::min(std::max(hc, 1u), 32u);
}

// Binned-SAH build.  `first_offset` is added to leaf `first` fields (the position
// of this tree's items in the global primitive array), `node_offset` to child
// indices (the position of this tree in the global node array; must be even).
//
// Level-synchronous and multi-threaded: a level's tasks own disjoint ranges of the index
// array, so they are analysed (bounds, 3 x 16 bins, SAH sweep, partition) in parallel; near
// the root, where a level has fewer tasks than threads, the passes over one task's range are
// split across the threads instead.  Node numbering is assigned serially per level, so the
// tree is identical for any thread count.
// `trav_cost`: cost of one traversal step in units of one primitive test (SAH termination).
inline BuildResult build_bvh(const std::vector<BuildItem> &items, int max_leaf, uint32_t first_offset,
                             uint32_t node_offset, double trav_cost = 1.0, bool layout_dfs = false) {
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
    struct Decision { // of a task that is split: where (0 = the task became a leaf)
        uint32_t mid;
    };
    out.nodes.reserve(2 * size_t(n) + 2);
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
        if (count == 1 && items[idx[t.begin]].instance)
            nd.ref |= kLeafInstanceFlag;
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
    const unsigned threads = n >= 8192 ? builder_threads() : 1;

    struct Bounds {
        Box b, cb;
        bool solitary = false;
    };
    struct Bins {
        Box bb[3][kBins];
        uint32_t bc[3][kBins];
    };
    auto bin_of = [&](const BuildItem &it, int ax, double lo, double ext) {
        int bi = int(kBins * ((it.box.centre(ax) - lo) / ext));
        return std::min(std::max(bi, 0), kBins - 1);
    };
    // analyse one task (its node slot already exists: bounds and leaf references are written
    // here, by the thread that owns the task); `inner` = threads available for the passes
    auto analyse = [&](const Task &t, unsigned inner) -> Decision {
        const uint32_t cnt = t.end - t.begin;
        const unsigned parts = (inner > 1 && cnt >= 65536) ? inner : 1;
        const uint32_t per = (cnt + parts - 1) / parts;
        // pass 1: bounds of the boxes and of their centres
        std::vector<Bounds> pb(parts);
        parallel_for(parts, 1, parts, [&](size_t q) {
            Bounds &r = pb[q];
            const uint32_t lo = t.begin + uint32_t(q) * per, hi = std::min(t.end, lo + per);
            for (uint32_t i = lo; i < hi; ++i) {
                const BuildItem &it = items[idx[i]];
                r.b.grow(it.box);
                const double c[3] = {it.box.centre(0), it.box.centre(1), it.box.centre(2)};
                r.cb.grow(c);
                r.solitary |= it.solitary;
            }
        });
        Bounds B = pb[0];
        for (unsigned q = 1; q < parts; ++q) {
            B.b.grow(pb[q].b);
            B.cb.grow(pb[q].cb);
            B.solitary |= pb[q].solitary;
        }
        Decision d;
        d.mid = 0;
        if (cnt == 1) {
            make_leaf(t, B.b);
            return d;
        }
        // pass 2: the three axes' bins in one sweep
        double lo3[3], ext3[3];
        for (int ax = 0; ax < 3; ++ax) {
            lo3[ax] = B.cb.lo[ax];
            ext3[ax] = B.cb.hi[ax] - B.cb.lo[ax];
        }
        std::vector<Bins> bins(parts);
        parallel_for(parts, 1, parts, [&](size_t q) {
            Bins &r = bins[q];
            std::memset(r.bc, 0, sizeof(r.bc));
            const uint32_t lo = t.begin + uint32_t(q) * per, hi = std::min(t.end, lo + per);
            for (uint32_t i = lo; i < hi; ++i) {
                const BuildItem &it = items[idx[i]];
                for (int ax = 0; ax < 3; ++ax) {
                    if (!(ext3[ax] > 0))
                        continue;
                    const int bi = bin_of(it, ax, lo3[ax], ext3[ax]);
                    r.bb[ax][bi].grow(it.box);
                    r.bc[ax][bi]++;
                }
            }
        });
        for (unsigned q = 1; q < parts; ++q)
            for (int ax = 0; ax < 3; ++ax)
                for (int i = 0; i < kBins; ++i) {
                    bins[0].bb[ax][i].grow(bins[q].bb[ax][i]);
                    bins[0].bc[ax][i] += bins[q].bc[ax][i];
                }
        // SAH sweep
        double best_cost = INFINITY;
        int best_axis = -1, best_bin = -1;
        for (int ax = 0; ax < 3; ++ax) {
            if (!(ext3[ax] > 0))
                continue;
            const Box *bb = bins[0].bb[ax];
            const uint32_t *bc = bins[0].bc[ax];
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
        const double leaf_cost = B.b.area() * cnt;
        const bool can_leaf = cnt <= uint32_t(max_leaf) && !B.solitary;
        // traversal step ~ one primitive test
        if (can_leaf && (best_axis < 0 || best_cost + trav_cost * B.b.area() >= leaf_cost)) {
            make_leaf(t, B.b);
            return d;
        }
        store_bounds(out.nodes[t.node], B.b);
        uint32_t mid;
        if (best_axis >= 0 && t.depth < kMaxDepth) {
            const double lo = lo3[best_axis], ext = ext3[best_axis];
            auto it = std::partition(idx.begin() + t.begin, idx.begin() + t.end,
                                     [&](uint32_t id) { return bin_of(items[id], best_axis, lo, ext) <= best_bin; });
            mid = uint32_t(it - idx.begin());
        } else {
            // coincident centroids or depth cap: median split in index order
            mid = t.begin + cnt / 2;
        }
        if (mid == t.begin || mid == t.end)
            mid = t.begin + cnt / 2;
        d.mid = mid;
        return d;
    };

    std::vector<Decision> dec;
    while (!level.empty()) {
        next.clear();
        dec.resize(level.size());
        if (threads > 1 && level.size() < size_t(4 * threads)) {
            for (size_t k = 0; k < level.size(); ++k) // few, large tasks: parallel inside each
                dec[k] = analyse(level[k], threads);
        } else {
            parallel_for(level.size(), 8, threads, [&](size_t k) { dec[k] = analyse(level[k], 1); });
        }
        for (size_t k = 0; k < level.size(); ++k) { // serial: node numbering
            const Task &t = level[k];
            if (dec[k].mid == 0)
                continue;
            const uint32_t child = uint32_t(out.nodes.size());
            out.nodes.resize(child + 2); // within the reserved capacity: no reallocation
            Node32 &nd = out.nodes[t.node];
            nd.ref = node_offset + child;
            nd.count = 0;
            next.push_back({child, t.begin, dec[k].mid, t.depth + 1});
            next.push_back({child + 1, dec[k].mid, t.end, t.depth + 1});
        }
        level.swap(next);
    }
    out.order.resize(n);
    for (uint32_t i = 0; i < n; ++i)
        out.order[i] = items[idx[i]].id;
    if (layout_dfs && out.nodes.size() > 2) {
        // Re-number the sibling pairs depth-first (pair, left subtree, right subtree): the pair a
        // descent needs next is then usually the 64 bytes right after the current one, instead of
        // twice as far down the array as in level order.
        std::vector<Node32> dfs(out.nodes.size());
        dfs[0] = out.nodes[0];
        dfs[1] = out.nodes[1];
        uint32_t next_pair = 2;
        struct Fix {
            uint32_t at; // index in dfs[] of the interior node whose ref is to be set
            uint32_t old_pair;
        };
        std::vector<Fix> stack;
        if (!(dfs[0].ref & kLeafFlag))
            stack.push_back({0, dfs[0].ref - node_offset});
        while (!stack.empty()) {
            const Fix f = stack.back();
            stack.pop_back();
            const uint32_t at = next_pair;
            next_pair += 2;
            dfs[at] = out.nodes[f.old_pair];
            dfs[at + 1] = out.nodes[f.old_pair + 1];
            dfs[f.at].ref = node_offset + at;
            // push right first so that the left subtree is laid out immediately after its pair
            if (!(dfs[at + 1].ref & kLeafFlag))
                stack.push_back({at + 1, dfs[at + 1].ref - node_offset});
            if (!(dfs[at].ref & kLeafFlag))
                stack.push_back({at, dfs[at].ref - node_offset});
        }
        out.nodes.swap(dfs);
    }
    return out;
}

} // namespace rtb

#endif // RTB_BVH_HPP
