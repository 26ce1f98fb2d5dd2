// rtb_geom.cuh — device geometry tables, primitive tests, instance chains, the
// two-level BVH traversal and hit-record reconstruction.  Templated on R
// (float production / double validation).  In the double instantiation every
// expression that feeds `t` is written in the reference's operation order.
//
// Reference counterparts:
//   bvh_node::hit            src/geometry/bvh.h:40-50      -> traverse()
//   aabb::hit                src/geometry/aabb.h:31-48     -> slab()
//   sphere::hit              src/geometry/sphere.h:33-60   -> hit_sphere()
//   moving_sphere::hit       src/geometry/moving_sphere.h:36-62
//   xy/xz/yz_rect::hit       src/geometry/aarect.h:79-135  -> hit_rect()
//   translate/rotate_y/flip  src/geometry/hittable.h:51-62,127-156,163-170
//   constant_medium::hit     src/geometry/constant_medium.h:55-104
//
// The reference walks a pointer graph; here the world is a flat array of
// 32-byte primitive records ordered by BVH leaf, a flat array of 32-byte BVH
// nodes (top level over world-space items, one bottom level per distinct
// wrapper chain) and small tables of wrapper ops.
#ifndef RTB_GEOM_CUH
#define RTB_GEOM_CUH

#include "rtb_math.cuh"

namespace rtb {

enum : uint32_t {
    PT_SPHERE = 0,
    PT_MSPHERE = 1,
    PT_XY = 2,
    PT_XZ = 3,
    PT_YZ = 4,
    PT_MEDIUM = 5,
    PT_INSTANCE = 6,
    PT_BOX = 7, // a `box` object (box.h) of a large scene as ONE tree item: rtb_scene_host.hpp, box_slot() below
    PT_TYPE_MASK = 7,
    PT_DUP_LEAF = 8, // reference tests this object twice per ray (bvh.h:68-69)
    // bits 4-6: the hit queue a path that hits this primitive goes to (its material's type, or 6 for
    // a lambertian with a procedural / image albedo) — baked in at upload so that sorting hits by
    // material costs no look-up: the traversal already holds the record of the primitive it hit
    PT_KEY_SHIFT = 4,
    PT_KEY_MASK = 7,
    // a sphere of negative radius that the reference only reaches through one bvh_node's box
    // (rtb200_scene.h rtb_gate): aux -> the box, in the MovingAux table
    PT_GATED = 128,
    PT_MAT_SHIFT = 8
};

constexpr uint32_t kNoPrim = 0xffffffffu;
constexpr int kMaxChainOps = 8;
constexpr int kStackDepth = 48;

// One primitive.  float: 32 bytes exactly (one sector); double: 56 bytes.
//   SPHERE   d = cx cy cz r ; PT_GATED: aux -> its gate box (a MovingAux record)
//   MSPHERE  d = c0x c0y c0z r ; aux -> MovingAux
//   RECT     d = a0 a1 b0 b1 k   (XY: a=x b=y | XZ: a=x b=z | YZ: a=y b=z)
//   MEDIUM   d[0] = neg_inv_density ; aux = first boundary prim, aux2 = count
//   INSTANCE aux = BLAS root node, aux2 = chain id
//   BOX      d = lox loy loz hix hiy, aux = hiz (float bits), aux2 = first of its six face records
//            (z hi, z lo, y hi, y lo, x hi, x lo: the order of box.h); the fp64 record only carries aux2
template <class R> struct PrimT {
    R d[5];
    uint32_t type_mat;
    uint32_t aux;
    uint32_t aux2;
};
static_assert(sizeof(PrimT<float>) == 32, "production primitive record must be one 32-byte sector");

// Per-primitive auxiliary record: the second centre and the time range of a moving sphere, or the gate
// box of a PT_GATED sphere (lo = c1, hi = time0, time1, extra).
template <class R> struct MovingAux {
    R c1[3];
    R time0, time1;
    R extra;
};

template <class R> struct XfOp {
    int32_t kind; // rtb_xform_kind
    R a, b, c;
};

struct ChainRec {
    int32_t first, count;
};

// A whole wrapper chain folded into one map (fp32 production path only).  translate and
// rotate_y compose to "rotate about y, then shift":  o' = Ry(c,s) o + b,  d' = Ry(c,s) d.
// Built on the host in fp64 from the same ops the validation path applies one by one.
struct ChainAffine {
    float c, s, bx, by, bz;
};

// 32-byte BVH node: its own bounds plus a self-describing reference `ref`:
//   interior  ref = index of its first child; the two children are adjacent (a 64-byte
//                   aligned pair), so one step of the traversal is one 64-byte fetch;
//   leaf      ref = kLeafFlag | (count-1) << 27 | [kLeafInstanceFlag] | first   -> primitives [first, first+count).
// The traversal keeps only refs on its stack and never re-reads a node to learn what it is.
// `count` repeats the leaf size (0 for interior nodes) for tools that inspect the tree.
constexpr uint32_t kLeafFlag = 0x80000000u;
constexpr uint32_t kLeafFirstMask = 0x03ffffffu;    // 64 Mi primitive records
constexpr uint32_t kLeafInstanceFlag = 0x04000000u; // the leaf's single item is an instance record
constexpr int kMaxLeafPrims = 16;
constexpr uint32_t kEmptyRef = 0xffffffffu;    // a tree without primitives
constexpr uint32_t kSentinelRef = 0xfffffffeu; // stack marker: "leave the instance"
struct alignas(32) Node32 {
    float lo[3];
    uint32_t ref;
    float hi[3];
    uint32_t count;
};
static_assert(sizeof(Node32) == 32, "BVH node must be 32 bytes");

template <class R> struct GeomView {
    const Node32 *nodes;
    const PrimT<R> *prims;
    const MovingAux<R> *maux;
    const XfOp<R> *ops;
    const ChainRec *chains;
    const ChainAffine *affine; // per chain (fp32 view only; null in the fp64 view)
    const int32_t *prim_chain; // per sorted prim: chain id or -1
    const int32_t *prim_orig;  // per sorted prim: flat (blob) primitive id, -1 for instances
    int32_t n_nodes;
    int32_t n_prims;
    int32_t n_ops;
    int32_t n_chains;
    uint32_t root_ref; // ref of the top-level root (its box is never tested)
    int32_t n_top; // prims[0 .. n_top) are the top-level items (primitives and instance records)
    int32_t flat;  // != 0: the scene is small enough for the lockstep traversal (traverse_flat)
    // PT_GATED spheres (rtb_gate).  gate_mode 0: tested inside the traversal, the gate box against the
    // interval of that moment (production); 1: skipped — trace_gated_exact() adds them afterwards in the
    // reference's order.  orig_limit: primitives whose blob index is >= it are skipped (fp64 paths only).
    int32_t gate_mode = 0;
    int32_t orig_limit = 0x7fffffff;
    int32_t n_world = 0; // sorted primitives [n_world, n_prims) only exist as medium boundaries (RTB_PRIM_BOUNDARY_ONLY)
    int32_t n_gated = 0;
    uint32_t gated[4] = {0, 0, 0, 0}; // sorted indices of the gated spheres, ascending blob index
};
constexpr int kMaxGated = 4;

// Scenes with at most this many primitive records are traversed in lockstep from shared
// memory instead of through the BVH.
constexpr int kFlatMaxPrims = 64;
constexpr int kFlatMaxOps = 32;
constexpr int kFlatMaxChains = 16;

template <class R> struct RayT {
    V3<R> o, d;
    R time;
};

// ---- wrapper chains ---------------------------------------------------------------------
// hittable.h:53 (translate) and hittable.h:129-138 (rotate_y): the RAY is moved
// into object space; t is unchanged because directions are not normalised.
template <class R> RTB_HD void apply_op(const XfOp<R> &op, V3<R> &o, V3<R> &d) {
    if (op.kind == 0) { // translate
        o = V3<R>(o.x - op.a, o.y - op.b, o.z - op.c);
    } else if (op.kind == 1) { // rotate_y : a = sin, b = cos
        const R ox = op.b * o.x - op.a * o.z;
        const R oz = op.a * o.x + op.b * o.z;
        const R dx = op.b * d.x - op.a * d.z;
        const R dz = op.a * d.x + op.b * d.z;
        o.x = ox;
        o.z = oz;
        d.x = dx;
        d.z = dz;
    }
}

// (measured knobs, both off: the general fused kernel shrinks from 13,700 to 10,500 instructions with them and gets
// SLOWER — C4-env 40.8 -> 44.5 ms: the calls pin registers around the traversal loop)
#ifndef RTB_SLIM_CHAIN
#define RTB_SLIM_CHAIN 0
#endif
#ifndef RTB_SLIM_BOUNDARY
#define RTB_SLIM_BOUNDARY 0
#endif
template <class R> RTB_HD void apply_chain(const GeomView<R> &g, int chain, V3<R> &o, V3<R> &d) {
    if (chain < 0)
        return;
    const ChainRec c = g.chains[chain];
#if RTB_SLIM_CHAIN
#pragma unroll 1
#endif
    for (int i = 0; i < c.count; ++i)
        apply_op(g.ops[c.first + i], o, d);
}

// Moves the ray into the object space of an instance.  Validation (fp64): op by op, exactly
// as the reference.  Production (fp32): one fused rotate+shift, no loop, no branches.
template <class R, bool ROBUST>
RTB_HD void enter_instance(const GeomView<R> &g, int chain, V3<R> &o, V3<R> &d) {
    if (ROBUST && sizeof(R) == 4) {
        const ChainAffine a = g.affine[chain];
        const R ox = R(a.c) * o.x - R(a.s) * o.z + R(a.bx);
        const R oz = R(a.s) * o.x + R(a.c) * o.z + R(a.bz);
        const R dx = R(a.c) * d.x - R(a.s) * d.z;
        const R dz = R(a.s) * d.x + R(a.c) * d.z;
        o = V3<R>(ox, o.y + R(a.by), oz);
        d = V3<R>(dx, d.y, dz);
    } else {
        apply_chain(g, chain, o, d);
    }
}

// ---- primitive tests (return t or a negative "no hit" marker) --------------------------------

// sphere.h:33-50.  ROBUST (float production): `on_surface` says the ray starts on
// this very sphere, whose near root is then analytically 0 (rejected by t_min in
// the fp64 reference) and whose other root is -2*half_b/a.
template <class R, bool ROBUST>
RTB_HD bool hit_sphere(V3<R> c, R radius, V3<R> o, V3<R> d, R t_min, R t_max, bool on_surface,
                       R &t_out) {
    const V3<R> oc = o - c;
    const R a = length_squared(d);
    const R half_b = dot(oc, d);
    if (ROBUST) {
        if (on_surface) {
            const R t = R(-2) * half_b / a;
            if (t < t_min || t > t_max)
                return false;
            t_out = t;
            return true;
        }
        // Haines et al., "Precision improvements for ray/sphere intersection":
        // the discriminant from the perpendicular distance, the roots via q.
        const R inv_a = R(1) / a;
        const V3<R> l = oc - (half_b * inv_a) * d;
        const R disc = radius * radius - length_squared(l);
        if (disc < 0)
            return false;
        const R cc = length_squared(oc) - radius * radius;
        const R s = sqrt_(a * disc);
        const R q = -half_b + (half_b > 0 ? -s : s); // q = -(half_b + sign(half_b) s)
        R t0 = q * inv_a, t1 = (q != 0) ? cc / q : t0;
        if (t0 > t1) {
            const R tmp = t0;
            t0 = t1;
            t1 = tmp;
        }
        R root = t0;
        if (root < t_min || root > t_max) {
            root = t1;
            if (root < t_min || root > t_max)
                return false;
        }
        t_out = root;
        return true;
    } else {
        const R cc = length_squared(oc) - radius * radius;
        const R discriminant = half_b * half_b - a * cc;
        if (discriminant < 0)
            return false;
        const R sqrtd = sqrt_(discriminant);
        R root = (-half_b - sqrtd) / a;
        if (root < t_min || root > t_max) {
            root = (-half_b + sqrtd) / a;
            if (root < t_min || root > t_max)
                return false;
        }
        t_out = root;
        return true;
    }
}

// moving_sphere.h:32-34
template <class R> RTB_HD V3<R> moving_center(const PrimT<R> &p, const MovingAux<R> &m, R time) {
    const V3<R> c0(p.d[0], p.d[1], p.d[2]);
    const V3<R> c1(m.c1[0], m.c1[1], m.c1[2]);
    return c0 + ((time - m.time0) / (m.time1 - m.time0)) * (c1 - c0);
}

// aarect.h:79-135.  AX = index of the constant axis, A/B the two in-plane axes.
template <class R, bool ROBUST>
RTB_HD bool hit_rect(const PrimT<R> &p, int AX, int A, int B, V3<R> o, V3<R> d, V3<R> idir, R t_min,
                     R t_max, R &t_out) {
    if (ROBUST) {
        // production: one predicate, no early exits (in the lockstep traversal every lane of
        // the warp tests the same rect anyway, so branches only add reconvergence overhead)
        const R t = (p.d[4] - o[AX]) * idir[AX];
        const R a = o[A] + t * d[A];
        const R b = o[B] + t * d[B];
        const bool ok = (t >= t_min) & (t <= t_max) & (a >= p.d[0]) & (a <= p.d[1]) & (b >= p.d[2]) & (b <= p.d[3]);
        if (ok)
            t_out = t;
        return ok;
    }
    const R t = (p.d[4] - o[AX]) / d[AX];
    if (t < t_min || t > t_max)
        return false;
    const R a = o[A] + t * d[A];
    const R b = o[B] + t * d[B];
    if (a < p.d[0] || a > p.d[1] || b < p.d[2] || b > p.d[3])
        return false;
    t_out = t;
    return true;
}

// aabb::hit (aabb.h:31-48) of a gate box, with the interval the sphere itself is about to be tested with
template <class R> RTB_HD bool gate_pass(const MovingAux<R> &b, V3<R> o, V3<R> idir, R t_min, R t_max) {
    const R lo[3] = {b.c1[0], b.c1[1], b.c1[2]}, hi[3] = {b.time0, b.time1, b.extra};
    for (int a = 0; a < 3; ++a) {
        R t0 = (lo[a] - o[a]) * idir[a], t1 = (hi[a] - o[a]) * idir[a];
        if (idir[a] < 0) {
            const R tmp = t0;
            t0 = t1;
            t1 = tmp;
        }
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max <= t_min)
            return false;
    }
    return true;
}

// One non-instance, non-medium primitive in ITS OWN object space.
template <class R, bool ROBUST>
RTB_HD bool hit_simple(const GeomView<R> &g, const PrimT<R> &p, uint32_t type, V3<R> o, V3<R> d,
                       V3<R> idir, R time, R t_min, R t_max, bool is_origin, R &t_out) {
    switch (type) {
    case PT_SPHERE:
        if ((p.type_mat & PT_GATED) && (g.gate_mode != 0 || !gate_pass<R>(g.maux[p.aux], o, idir, t_min, t_max)))
            return false;
        return hit_sphere<R, ROBUST>(V3<R>(p.d[0], p.d[1], p.d[2]), p.d[3], o, d, t_min, t_max,
                                     is_origin, t_out);
    case PT_MSPHERE:
        return hit_sphere<R, ROBUST>(moving_center(p, g.maux[p.aux], time), p.d[3], o, d, t_min,
                                     t_max, false, t_out);
    case PT_XY:
        if (ROBUST && is_origin)
            return false;
        return hit_rect<R, ROBUST>(p, 2, 0, 1, o, d, idir, t_min, t_max, t_out);
    case PT_XZ:
        if (ROBUST && is_origin)
            return false;
        return hit_rect<R, ROBUST>(p, 1, 0, 2, o, d, idir, t_min, t_max, t_out);
    case PT_YZ:
        if (ROBUST && is_origin)
            return false;
        return hit_rect<R, ROBUST>(p, 0, 1, 2, o, d, idir, t_min, t_max, t_out);
    default:
        return false;
    }
}

template <class R> RTB_HD V3<R> safe_inv(V3<R> d) { return V3<R>(R(1) / d.x, R(1) / d.y, R(1) / d.z); }

// ---- a grouped box (PT_BOX) as one leaf slot ------------------------------------------------------
// Production (fp32, ROBUST): one slab test; the face the ray enters (or, for a ray that starts on this box or
// inside it, leaves) through is the hit, named by the index of that face's own rect record — the same plane
// arithmetic as hit_rect ((k - o) * idir), the rect's in-plane bounds replaced by the slab interval.
// Validation (fp64 / !ROBUST): the six face records one by one, exactly what the tree did before the faces
// were grouped (tie rule and orig_limit of the callers included) — grouping changes which tree node holds a
// rect, never which rect wins.  Returns true when an any-hit query is decided.
template <class R, bool ANY, bool ROBUST>
RTB_HD bool box_slot(const GeomView<R> &g, const PrimT<R> &p, V3<R> o, V3<R> d, V3<R> idir, R time, R t_min, R &t_max,
                     uint32_t origin_prim, uint32_t &best, uint64_t *n_tests) {
    const uint32_t first = p.aux2;
    if (ROBUST && sizeof(R) == 4) {
        if (n_tests)
            ++*n_tests;
        float hzf;
        memcpy(&hzf, &p.aux, 4);
        const R hz = R(hzf);
        const R tx0 = (p.d[0] - o.x) * idir.x, tx1 = (p.d[3] - o.x) * idir.x;
        const R ty0 = (p.d[1] - o.y) * idir.y, ty1 = (p.d[4] - o.y) * idir.y;
        const R tz0 = (p.d[2] - o.z) * idir.z, tz1 = (hz - o.z) * idir.z;
        const R tn = fmax_(fmax_(fmin_(tx0, tx1), fmin_(ty0, ty1)), fmin_(tz0, tz1));
        const R tf = fmin_(fmin_(fmax_(tx0, tx1), fmax_(ty0, ty1)), fmax_(tz0, tz1));
        const bool mine = (origin_prim - first) < 6u; // the ray starts on this box: it can only leave it
        const bool enter = !mine & (tn >= t_min);
        const R cand = enter ? tn : tf;
        const uint32_t face = cand == tz1 ? 0u : (cand == tz0 ? 1u : (cand == ty1 ? 2u : (cand == ty0 ? 3u : (cand == tx1 ? 4u : 5u))));
        const bool ok = (tn <= tf) & (cand >= t_min) & (cand <= t_max) & (first + face != origin_prim);
        if (ok) {
            best = first + face;
            t_max = cand;
        }
        return ANY && ok;
    }
    for (uint32_t f = 0; f < 6u; ++f) {
        const uint32_t i = first + f;
        if (!ROBUST && g.prim_orig[i] >= g.orig_limit)
            continue;
        if (n_tests)
            ++*n_tests;
        const PrimT<R> q = g.prims[i];
        R t;
        bool h = hit_simple<R, ROBUST>(g, q, q.type_mat & PT_TYPE_MASK, o, d, idir, time, t_min, t_max, ROBUST && i == origin_prim, t);
        if (!ROBUST && h && best != kNoPrim && t == t_max && g.prim_orig[i] < g.prim_orig[best])
            h = false;
        if (h) {
            best = i;
            t_max = t;
            if (ANY)
                return true;
        }
    }
    return false;
}

// Closest hit of the WORLD-space ray against boundary primitives [first, first+count)
// — constant_medium's `boundary->hit()` (constant_medium.h:64-68).  Each boundary
// primitive carries its full wrapper chain.
template <class R, bool ROBUST>
#if RTB_SLIM_BOUNDARY
RTB_HD_OUTLINE
#else
RTB_HD
#endif
bool hit_boundary(const GeomView<R> &g, uint32_t first, uint32_t count, V3<R> o, V3<R> d,
                         R time, R t_min, R t_max, R &t_out) {
    bool any = false;
    int cur_chain = -2;
    V3<R> lo = o, ld = d, lid = safe_inv(d);
    for (uint32_t i = first; i < first + count; ++i) {
        const PrimT<R> p = g.prims[i];
        const int ch = g.prim_chain[i];
        if (ch != cur_chain) {
            lo = o;
            ld = d;
            apply_chain(g, ch, lo, ld);
            lid = safe_inv(ld);
            cur_chain = ch;
        }
        R t;
        if (hit_simple<R, ROBUST>(g, p, p.type_mat & PT_TYPE_MASK, lo, ld, lid, time, t_min, t_max,
                                  false, t)) {
            any = true;
            t_max = t;
            t_out = t;
        }
    }
    return any;
}

// constant_medium.h:55-104.  draw() — uniform in (0,1) — is called exactly where the reference calls
// random_double() (constant_medium.h:85), i.e. only for a ray with a non-empty span inside the boundary.
template <class R, bool ROBUST, class Draw>
RTB_HD bool hit_medium_draw(const GeomView<R> &g, const PrimT<R> &p, V3<R> o, V3<R> d, R time, R t_min,
                            R t_max, Draw &draw, R &t_out) {
    const R inf = Consts<R>::inf();
    R t1, t2;
    bool have = false;
    if (ROBUST && p.aux2 == 1u && g.prim_chain[p.aux] < 0) {
        const PrimT<R> b = g.prims[p.aux];
        if ((b.type_mat & PT_TYPE_MASK) == PT_SPHERE) {
            // A single untransformed sphere as boundary (both media of scene09): the two
            // boundary->hit() calls are the two roots of ONE quadratic (the arithmetic of
            // hit_sphere's ROBUST branch), instead of two walks over the boundary list.
            const V3<R> oc = o - V3<R>(b.d[0], b.d[1], b.d[2]);
            const R a = length_squared(d), half_b = dot(oc, d), inv_a = R(1) / a;
            const V3<R> l = oc - (half_b * inv_a) * d;
            const R disc = b.d[3] * b.d[3] - length_squared(l);
            if (disc < 0)
                return false;
            const R cc = length_squared(oc) - b.d[3] * b.d[3];
            const R sq = sqrt_(a * disc);
            const R q = -half_b + (half_b > 0 ? -sq : sq);
            const R r0 = q * inv_a, r1 = (q != 0) ? cc / q : r0;
            t1 = fmin_(r0, r1);
            t2 = fmax_(r0, r1);
            const R step = fmax_(R(0.0001), fabs_(t1) * R(1e-6));
            if (!(t2 >= t1 + step)) // the second hit() starts at t1 + step
                return false;
            have = true;
        }
    }
    if (!have) {
        if (!hit_boundary<R, ROBUST>(g, p.aux, p.aux2, o, d, time, -inf, inf, t1))
            return false;
        // ROBUST: 1e-4 is below one fp32 ulp once |t1| > ~800 (a ray deep inside a large
        // boundary), which would return the entry point again; step by a few ulps instead
        const R step = ROBUST ? fmax_(R(0.0001), fabs_(t1) * R(1e-6)) : R(0.0001);
        if (!hit_boundary<R, ROBUST>(g, p.aux, p.aux2, o, d, time, t1 + step, inf, t2))
            return false;
    }
    if (t1 < t_min)
        t1 = t_min;
    if (t2 > t_max)
        t2 = t_max;
    if (t1 >= t2)
        return false;
    if (t1 < 0)
        t1 = 0;
    const R ray_length = length(d);
    const R distance_inside_boundary = (t2 - t1) * ray_length;
    const R hit_distance = p.d[0] * log_(draw());
    if (hit_distance > distance_inside_boundary)
        return false;
    t_out = t1 + hit_distance / ray_length;
    return true;
}
template <class R> struct FixedXi {
    R xi;
    RTB_HD R operator()() const { return xi; }
};
// with the draw made up front (the kernels: one draw per test keeps a ray's stream independent of the path
// the traversal takes through the tree)
template <class R, bool ROBUST>
RTB_HD bool hit_medium(const GeomView<R> &g, const PrimT<R> &p, V3<R> o, V3<R> d, R time, R t_min,
                       R t_max, R xi, R &t_out) {
    FixedXi<R> f{xi};
    return hit_medium_draw<R, ROBUST>(g, p, o, d, time, t_min, t_max, f, t_out);
}

// ---- BVH traversal -----------------------------------------------------------------------

// Slab test of one node against [t_min, t_max]; returns entry distance via t_near.
// Conservative (<=) where the reference rejects on equality (aabb.h:43): node
// boxes here are padded outward, so nothing the reference accepts is culled.
template <class R>
RTB_HD bool slab(const Node32 &n, V3<R> o, V3<R> idir, R t_min, R t_max, R &t_near) {
    const R x0 = (R(n.lo[0]) - o.x) * idir.x, x1 = (R(n.hi[0]) - o.x) * idir.x;
    const R y0 = (R(n.lo[1]) - o.y) * idir.y, y1 = (R(n.hi[1]) - o.y) * idir.y;
    const R z0 = (R(n.lo[2]) - o.z) * idir.z, z1 = (R(n.hi[2]) - o.z) * idir.z;
    const R tn = fmax_(fmax_(fmin_(x0, x1), fmin_(y0, y1)), fmax_(fmin_(z0, z1), t_min));
    const R tf = fmin_(fmin_(fmax_(x0, x1), fmax_(y0, y1)), fmin_(fmax_(z0, z1), t_max));
    t_near = tn;
    return tn <= tf;
}

RTB_HD Node32 load_node(const Node32 *nodes, uint32_t i) {
#ifdef __CUDA_ARCH__
    const float4 *q = reinterpret_cast<const float4 *>(nodes + i);
    const float4 a = __ldg(q), b = __ldg(q + 1);
    Node32 n;
    n.lo[0] = a.x; n.lo[1] = a.y; n.lo[2] = a.z; n.ref = __float_as_uint(a.w);
    n.hi[0] = b.x; n.hi[1] = b.y; n.hi[2] = b.z; n.count = __float_as_uint(b.w);
    return n;
#else
    return nodes[i];
#endif
}

// ---- traversal stacks ------------------------------------------------------------------------
// The traversal is written against a small stack interface so that kernels can keep the stack
// in shared memory while host instantiations (tests/hostcheck) and simple kernels use a plain
// array.  A push beyond the capacity is dropped (the builder caps the tree depth below it).
// The storage is a separate array and the object holds only two scalars (pointer, depth): kept
// together in one struct the dynamically indexed array drags the depth counter into local
// memory too, and every push / pop then pays a dependent local load for it.
struct LocalStack {
    uint32_t *s;
    int sp;
    RTB_HD explicit LocalStack(uint32_t *storage) : s(storage), sp(0) {}
    RTB_HD void push(uint32_t x) {
        if (sp < kStackDepth)
            s[sp++] = x;
    }
    RTB_HD uint32_t pop() { return s[--sp]; }
    RTB_HD bool empty() const { return sp == 0; }
    RTB_HD void clear() { sp = 0; }
};

#ifdef __CUDACC__
// Per-thread stack in shared memory, level-major: level L of thread T lives at word
// L * STRIDE + T of the block's stack array, so the 32 lanes of a warp always touch 32
// different banks whatever their stack depths are.  Local memory would route every push/pop
// through the L1 that the BVH node fetches are fighting over.  The first kSmemStackDepth
// levels are in shared memory; deeper ones (rare: a near-first descent consumes the top of a
// tall tree first) go to `spill`, a local array owned by the caller.  The members are three
// scalars (shared-space address, depth, pointer) so the object itself stays in registers.
constexpr int kSmemStackDepth = 24;
template <int STRIDE> struct SmemStack {
    uint32_t saddr; // shared-window address of this thread's level 0
    int sp;
    uint32_t *spill; // kStackDepth - kSmemStackDepth words
    __device__ __forceinline__ SmemStack(uint32_t *thread_base, uint32_t *spill_words)
        : sp(0), spill(spill_words) {
        // opaque move: keeps the address in a register instead of being re-derived from
        // %tid / the shared window base (two S2R) at every push and pop
        asm volatile("mov.u32 %0, %1;" : "=r"(saddr) : "r"(uint32_t(__cvta_generic_to_shared(thread_base))));
    }
    __device__ __forceinline__ void push(uint32_t x) {
        if (sp < kSmemStackDepth)
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr + uint32_t(sp) * (STRIDE * 4u)), "r"(x) : "memory");
        else if (sp < kStackDepth)
            spill[sp - kSmemStackDepth] = x;
        else
            return;
        ++sp;
    }
    __device__ __forceinline__ uint32_t pop() {
        --sp;
        uint32_t x;
        if (sp < kSmemStackDepth)
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(saddr + uint32_t(sp) * (STRIDE * 4u)) : "memory");
        else
            x = spill[sp - kSmemStackDepth];
        return x;
    }
    __device__ __forceinline__ bool empty() const { return sp == 0; }
    __device__ __forceinline__ void clear() { sp = 0; }
};
#endif

// The ray as the slab test wants it.  Production (float, ROBUST): t = lo * idir + ood with
// ood = -(o * idir), one FMA per plane, |idir| capped so that ood stays finite for axis-parallel rays.
// Validation: the reference's (lo - o) * idir (aabb.h:33-36).
#ifndef RTB_SLAB_FMA
#define RTB_SLAB_FMA 1
#endif
template <class R, bool ROBUST> struct SlabRay {
    V3<R> idir, ood;
    RTB_HD void set(V3<R> o, V3<R> d) {
        if (RTB_SLAB_FMA && ROBUST && sizeof(R) == 4) {
            const R eps = R(1e-18);
            const R dx = fabs_(d.x) < eps ? (d.x < 0 ? -eps : eps) : d.x;
            const R dy = fabs_(d.y) < eps ? (d.y < 0 ? -eps : eps) : d.y;
            const R dz = fabs_(d.z) < eps ? (d.z < 0 ? -eps : eps) : d.z;
            idir = V3<R>(R(1) / dx, R(1) / dy, R(1) / dz);
            ood = V3<R>(-(o.x * idir.x), -(o.y * idir.y), -(o.z * idir.z));
        } else {
            idir = safe_inv(d);
            ood = o;
        }
    }
    RTB_HD R plane(R v, R id, R oo) const {
        if (RTB_SLAB_FMA && ROBUST && sizeof(R) == 4)
            return fma_(v, id, oo);
        return (v - oo) * id;
    }
    // entry / exit distance of a box clipped to [t_min, t_max]; hit when near <= far
    RTB_HD bool box(const Node32 &n, R t_min, R t_max, R &t_near) const {
        const R x0 = plane(R(n.lo[0]), idir.x, ood.x), x1 = plane(R(n.hi[0]), idir.x, ood.x);
        const R y0 = plane(R(n.lo[1]), idir.y, ood.y), y1 = plane(R(n.hi[1]), idir.y, ood.y);
        const R z0 = plane(R(n.lo[2]), idir.z, ood.z), z1 = plane(R(n.hi[2]), idir.z, ood.z);
        const R tn = fmax_(fmax_(fmin_(x0, x1), fmin_(y0, y1)), fmax_(fmin_(z0, z1), t_min));
        const R tf = fmin_(fmin_(fmax_(x0, x1), fmax_(y0, y1)), fmin_(fmax_(z0, z1), t_max));
        t_near = tn;
        return tn <= tf;
    }
};

// Closest hit (ANY = false) or first hit (ANY = true) of the world-space ray through the
// two-level BVH.  Returns the SORTED primitive index or kNoPrim, and t.
//   origin_prim : sorted index of the primitive the ray leaves (ROBUST only).
//   rng         : callable returning R uniform in (0,1) — drawn once per medium test, as
//                 constant_medium::hit does (constant_medium.h:85).
//   stack       : LocalStack or SmemStack, empty on entry.
// "while-while" form: the inner loop only descends (one 64-byte child-pair fetch and two slab
// tests per step, nearer child first, no data-dependent branches besides the loop itself),
// leaves are processed between descents, so the lanes of a warp spend most of their time in
// the same loop.
// MEDIA = false: the caller knows the scene holds no constant_medium, so that code (two boundary
// walks, a log) is not compiled into its kernel at all.
// INST = true: instance entry / exit (a ray transform each) is handled in the leaf phase, with the
// warp converged.  INST = false: inside the descent loop, recognised from the ref alone
// (kLeafInstanceFlag, kSentinelRef), so the leaf phase only ever tests primitives.  Both shapes
// give the same hits on any scene; which one is faster was MEASURED on B200: scenes with
// instances prefer the leaf-phase shape (C2 extend 40.1 vs 41.9 ms), scenes without — where the
// extra branches never execute — run faster with the leaner leaf phase of the other (C5 extend +
// connect 72.3 vs 75.5 ms).  The renderer picks by whether the scene has instances.
template <class R, bool ANY, bool ROBUST, class Rng, class Stack, bool MEDIA = true, bool INST = true>
RTB_HD uint32_t traverse(const GeomView<R> &g, V3<R> o, V3<R> d, R time, R t_min, R t_max, uint32_t origin_prim,
                         Rng &rng, R &t_hit, uint64_t *n_nodes, uint64_t *n_tests, Stack &stack) {
    uint32_t best = kNoPrim;
    V3<R> co = o, cd = d; // current-level ray
    SlabRay<R, ROBUST> sr;
    sr.set(o, d);
    uint32_t cur = g.root_ref;
    while (true) {
        while (true) {
            if (!(cur & kLeafFlag)) { // descend
                const Node32 c0 = load_node(g.nodes, cur), c1 = load_node(g.nodes, cur + 1);
                if (n_nodes)
                    *n_nodes += 2;
                R e0, e1;
                const bool h0 = sr.box(c0, t_min, t_max, e0);
                const bool h1 = sr.box(c1, t_min, t_max, e1);
                const bool first1 = h1 && (!h0 || (e1 < e0)); // child 1 is the one to enter next
                const uint32_t near = first1 ? c1.ref : c0.ref, far = first1 ? c0.ref : c1.ref;
                if (h0 && h1)
                    stack.push(far);
                cur = near;
                if (!h0 && !h1) {
                    if (stack.empty()) {
                        t_hit = t_max;
                        return best;
                    }
                    cur = stack.pop();
                }
                continue;
            }
            if (!INST) { // instance entry / exit inside the descent
                if (cur == kSentinelRef) { // leaving the instance: back to the world ray
                    co = o;
                    cd = d;
                    sr.set(o, d);
                    if (stack.empty()) {
                        t_hit = t_max;
                        return best;
                    }
                    cur = stack.pop();
                    continue;
                }
                if ((cur & kLeafInstanceFlag) && cur != kEmptyRef) {
                    const PrimT<R> p = g.prims[cur & kLeafFirstMask];
                    stack.push(kSentinelRef);
                    enter_instance<R, ROBUST>(g, int(p.aux2), co, cd);
                    sr.set(co, cd);
                    cur = p.aux; // the bottom-level tree's root ref
                    continue;
                }
            }
            break; // a leaf
        }
        bool entered = false;
        if (INST && cur == kSentinelRef) { // leaving the instance: back to the world ray
            co = o;
            cd = d;
            sr.set(o, d);
        } else if (cur != kEmptyRef) {
            const uint32_t first = cur & kLeafFirstMask, last = first + ((cur >> 27) & 15u) + 1u;
            V3<R> cid = safe_inv(cd);
            for (uint32_t i = first; i < last; ++i) {
                const PrimT<R> p = g.prims[i];
                const uint32_t type = p.type_mat & PT_TYPE_MASK;
                if (INST && type == PT_INSTANCE) {
                    // builder guarantee: an instance is alone in its leaf, top level only
                    stack.push(kSentinelRef);
                    enter_instance<R, ROBUST>(g, int(p.aux2), co, cd);
                    sr.set(co, cd);
                    cur = p.aux; // the bottom-level tree's root ref
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
                    if (p.type_mat & PT_DUP_LEAF) {
                        // second visit of a one-object bvh_node (bvh.h:46-47): t_max has
                        // already shrunk to the first answer
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
// with a stack of its own in local memory
template <class R, bool ANY, bool ROBUST, class Rng>
RTB_HD uint32_t traverse(const GeomView<R> &g, V3<R> o, V3<R> d, R time, R t_min, R t_max, uint32_t origin_prim,
                         Rng &rng, R &t_hit, uint64_t *n_nodes, uint64_t *n_tests) {
    uint32_t storage[kStackDepth];
    LocalStack stack(storage);
    return traverse<R, ANY, ROBUST>(g, o, d, time, t_min, t_max, origin_prim, rng, t_hit, n_nodes, n_tests, stack);
}

// Closest hit of a scene with PT_GATED spheres, exactly as the reference finds it (fp64 validation
// paths; the production kernels test a gated sphere inside the traversal, see GeomView::gate_mode).
// The reference tests a gated sphere S only if the box of the bvh_node holding it overlaps
// [t_min, tA], tA = the closest hit among the primitives that PRECEDE S in its left-to-right walk
// (bvh.h:40-50) — the blob order.  trace(view, t_max, t_out) is any closest-hit traversal of `view`.
//   1. t* = closest hit over everything but the gated spheres (gate_mode 1);
//   2. per gated sphere, ascending blob index: the gate cannot pass with t_max -> next; the sphere has no
//      root below t* -> next (whatever the reference finds loses to t*); the gate passes with t* -> it
//      passes with tA >= t*, and sphere::hit picks the same root for both bounds whenever that root is
//      below t*; otherwise tA is measured by a second traversal restricted to the preceding primitives.
template <class R, class Trace>
RTB_HD uint32_t trace_gated_exact(const GeomView<R> &g, V3<R> o, V3<R> d, R t_min, R t_max, R &t_hit, Trace trace) {
    GeomView<R> gv = g;
    gv.gate_mode = 1;
    R t_cur;
    uint32_t best = trace(gv, t_max, t_cur);
    R t_gated = t_max; // closest accepted gated sphere so far (they precede the next one)
    for (int k = 0; k < g.n_gated; ++k) {
        const uint32_t gi = g.gated[k];
        const PrimT<R> p = g.prims[gi];
        V3<R> lo = o, ld = d;
        const int chain = g.prim_chain[gi];
        if (chain >= 0) {
            const ChainRec c = g.chains[chain];
            for (int i = 0; i < c.count && i < kMaxChainOps; ++i)
                apply_op(g.ops[c.first + i], lo, ld);
        }
        const V3<R> lid = safe_inv(ld);
        const MovingAux<R> box = g.maux[p.aux];
        if (!gate_pass<R>(box, lo, lid, t_min, t_max))
            continue;
        R r;
        if (!hit_sphere<R, false>(V3<R>(p.d[0], p.d[1], p.d[2]), p.d[3], lo, ld, t_min, t_cur, false, r))
            continue;
        if (!gate_pass<R>(box, lo, lid, t_min, t_cur)) {
            GeomView<R> gl = gv;
            gl.orig_limit = g.prim_orig[gi];
            R t_a;
            trace(gl, t_max, t_a);
            t_a = t_gated < t_a ? t_gated : t_a;
            if (!gate_pass<R>(box, lo, lid, t_min, t_a))
                continue;
        }
        best = gi;
        t_cur = r;
        t_gated = r;
    }
    t_hit = t_cur;
    return best;
}

// The reference's own walk, for the fp64 validation entry point that pins constant_medium: every world
// primitive in BLOB order — the order bvh_node::hit / hittable_list::hit reach the leaves in (bvh.h:40-50,
// the flattener emits them so) — each through its wrapper chain, closest hit carried along, media drawing
// from `draw` exactly when the reference draws.  With the reference's xorshift32 state (rtweekend.h:24-34)
// behind `draw`, t and the primitive are the reference's bit for bit, media included.  No tree: O(n) per ray.
struct XorShift32Draw { // random_double(), rtweekend.h:24-34
    uint32_t s;
    RTB_HD double operator()() {
        s ^= s << 13;
        s ^= s >> 17;
        s ^= s << 5;
        return s * 2.3283064365386963e-10;
    }
};
template <class R, class Draw>
RTB_HD uint32_t walk_reference_order(const GeomView<R> &g, const int32_t *orig_to_sorted, int n_orig, V3<R> o, V3<R> d,
                                     R time, R t_min, R t_max, Draw &draw, R &t_hit) {
    uint32_t best = kNoPrim;
    for (int i = 0; i < n_orig; ++i) {
        const uint32_t s = uint32_t(orig_to_sorted[i]);
        if (int32_t(s) >= g.n_world)
            continue;
        const PrimT<R> p = g.prims[s];
        R t;
        bool h;
        if ((p.type_mat & PT_TYPE_MASK) == PT_MEDIUM) {
            h = hit_medium_draw<R, false>(g, p, o, d, time, t_min, t_max, draw, t);
            if (p.type_mat & PT_DUP_LEAF) { // bvh.h:46-47 on a one-object node
                R t2;
                if (hit_medium_draw<R, false>(g, p, o, d, time, t_min, h ? t : t_max, draw, t2)) {
                    h = true;
                    t = t2;
                }
            }
        } else {
            h = hit_boundary<R, false>(g, s, 1u, o, d, time, t_min, t_max, t); // (one primitive through its chain)
        }
        if (h) {
            best = s;
            t_max = t;
        }
    }
    t_hit = t_max;
    return best;
}

// Lockstep traversal for small scenes (the Cornell-box class: a few dozen primitives).
// A BVH does not pay there: the 32 rays of a warp take different branches of a tiny tree
// and the warp serialises them (measured on scene07: 9 of 32 lanes active per instruction).
// Instead every lane walks the SAME item list in the SAME order — top-level items, and
// inside an instance its primitive range after one ray transform — so control flow and
// primitive fetches are warp-uniform (shared-memory broadcasts) and the only divergence
// left is inside the individual tests.  Formally this is the same two-level structure
// with each level collapsed into one wide leaf; closest-hit results are identical.
// `g` must describe the flat layout ([top items][instance ranges][boundary prims]); in the
// kernels its table pointers point at the shared-memory copy.
template <class R, bool ANY, bool ROBUST, class Rng>
RTB_HD uint32_t traverse_flat(const GeomView<R> &g, V3<R> o, V3<R> d, R time, R t_min, R t_max,
                              uint32_t origin_prim, Rng &rng, R &t_hit, uint64_t *n_nodes,
                              uint64_t *n_tests) {
    uint32_t best = kNoPrim;
    const V3<R> idir = safe_inv(d);
    const uint32_t n_top = uint32_t(g.n_top);
    for (uint32_t i = 0; i < n_top; ++i) {
        const PrimT<R> p = g.prims[i];
        const uint32_t type = p.type_mat & PT_TYPE_MASK;
        if (type == PT_INSTANCE) {
            if (n_nodes)
                ++*n_nodes; // one "node" = one instance entry (ray transform)
            V3<R> lo = o, ld = d;
            enter_instance<R, ROBUST>(g, int(p.aux2), lo, ld);
            const V3<R> lid = safe_inv(ld);
            const uint32_t first = uint32_t(p.d[0]), last = first + uint32_t(p.d[1]);
            for (uint32_t j = first; j < last; ++j) {
                const PrimT<R> q = g.prims[j];
                if (n_tests)
                    ++*n_tests;
                R t;
                if (hit_simple<R, ROBUST>(g, q, q.type_mat & PT_TYPE_MASK, lo, ld, lid, time, t_min, t_max,
                                          ROBUST && j == origin_prim, t)) {
                    best = j;
                    t_max = t;
                    if (ANY) {
                        t_hit = t;
                        return best;
                    }
                }
            }
            continue;
        }
        if (n_tests)
            ++*n_tests;
        R t;
        bool h;
        if (type == PT_MEDIUM) {
            h = hit_medium<R, ROBUST>(g, p, o, d, time, t_min, t_max, rng(), t);
            if (p.type_mat & PT_DUP_LEAF) {
                R t2;
                if (hit_medium<R, ROBUST>(g, p, o, d, time, t_min, h ? t : t_max, rng(), t2)) {
                    h = true;
                    t = t2;
                }
            }
        } else {
            h = hit_simple<R, ROBUST>(g, p, type, o, d, idir, time, t_min, t_max, ROBUST && i == origin_prim, t);
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
    t_hit = t_max;
    return best;
}

// ---- hit record ---------------------------------------------------------------------------

template <class R> struct RecT {
    V3<R> p, normal;
    R t, u, v;
    bool front_face;
};

// hittable.h:19-22
template <class R> RTB_HD void set_face_normal(RecT<R> &rec, V3<R> dir, V3<R> outward) {
    rec.front_face = dot(dir, outward) < 0;
    rec.normal = rec.front_face ? outward : -outward;
}

// sphere.h:25-31
template <class R> RTB_HD void sphere_uv(V3<R> p, R &u, R &v) {
    const R pi = Consts<R>::pi();
    const R theta = acos_(-p.y);
    const R phi = atan2_(-p.z, p.x) + pi;
    u = phi / (R(2) * pi);
    v = theta / pi;
}

// Rebuilds what the reference's nested hit() calls leave in hit_record for the
// closest primitive `pi` at parameter t: the leaf fills the record in object
// space (sphere.h:52-57, aarect.h:88-96), then every wrapper on the way out
// post-processes it (hittable.h:58-59, 142-153, 168).  WANT_UV = false skips the
// acos/atan2 of sphere uv (only textures that read u,v need it).
template <class R, bool ROBUST, bool WANT_UV>
RTB_HD RecT<R> make_record(const GeomView<R> &g, uint32_t pi, V3<R> o, V3<R> d, R time, R t, bool uv = true) {
    // uv: run-time form of WANT_UV for callers that keep ONE instantiation (the general fused kernel)
    RecT<R> rec;
    rec.t = t;
    rec.u = 0;
    rec.v = 0;
    const PrimT<R> p = g.prims[pi];
    const uint32_t type = p.type_mat & PT_TYPE_MASK;
    const int chain = g.prim_chain[pi];
    // forward pass: the ray in the leaf's object space
    int nops = 0, first = 0;
    V3<R> lo = o, ld = d;
    if (chain >= 0) {
        const ChainRec c = g.chains[chain];
        first = c.first;
        nops = c.count < kMaxChainOps ? c.count : kMaxChainOps;
        for (int i = 0; i < nops; ++i)
            apply_op(g.ops[first + i], lo, ld);
    }
    rec.p = lo + t * ld;
    if (type == PT_SPHERE || type == PT_MSPHERE) {
        const V3<R> c = type == PT_SPHERE ? V3<R>(p.d[0], p.d[1], p.d[2])
                                          : moving_center(p, g.maux[p.aux], time);
        V3<R> outward = (rec.p - c) / p.d[3]; // r < 0 => points inward (hollow glass)
        if (ROBUST) { // pull the point back onto the sphere
            outward = unit_vector(outward);
            rec.p = c + p.d[3] * outward;
        }
        set_face_normal(rec, ld, outward);
        if (WANT_UV && uv && type == PT_SPHERE)
            sphere_uv(outward, rec.u, rec.v);
    } else if (type == PT_MEDIUM) {
        rec.normal = V3<R>(1, 0, 0); // constant_medium.h:99-100, "arbitrary"
        rec.front_face = true;
    } else {
        const int AX = type == PT_XY ? 2 : (type == PT_XZ ? 1 : 0);
        const int A = type == PT_YZ ? 1 : 0;
        const int B = type == PT_XY ? 1 : 2;
        rec.u = (rec.p[A] - p.d[0]) / (p.d[1] - p.d[0]);
        rec.v = (rec.p[B] - p.d[2]) / (p.d[3] - p.d[2]);
        if (ROBUST)
            rec.p.set(AX, p.d[4]);
        V3<R> outward(0, 0, 0);
        outward.set(AX, R(1));
        set_face_normal(rec, ld, outward);
    }
    // reverse pass.  Each wrapper re-runs set_face_normal against ITS OWN ray (the ray after
    // its transform, hittable.h:59,153).  Only directions matter, and only rotate_y changes
    // them, so the direction after wrapper i is recovered by undoing the rotations of the
    // wrappers below it (chains are 1-3 ops long; nothing is kept in local memory).
    V3<R> dir_i = ld; // direction after the innermost wrapper
    for (int i = nops - 1; i >= 0; --i) {
        const XfOp<R> op = g.ops[first + i];
        if (i < nops - 1) {
            const XfOp<R> below = g.ops[first + i + 1];
            if (below.kind == 1) { // exact inverse of apply_op's rotation is not needed bit for
                // bit: recompute from the world direction to stay bit-exact in fp64
                V3<R> tmp_o(0, 0, 0);
                dir_i = d;
                for (int k = 0; k <= i; ++k)
                    apply_op(g.ops[first + k], tmp_o, dir_i);
            }
        }
        if (op.kind == 0) { // hittable.h:58-59
            rec.p = V3<R>(rec.p.x + op.a, rec.p.y + op.b, rec.p.z + op.c);
            set_face_normal(rec, dir_i, rec.normal);
        } else if (op.kind == 1) { // hittable.h:142-153 : a = sin, b = cos
            const R px = op.b * rec.p.x + op.a * rec.p.z;
            const R pz = -op.a * rec.p.x + op.b * rec.p.z;
            const R nx = op.b * rec.normal.x + op.a * rec.normal.z;
            const R nz = -op.a * rec.normal.x + op.b * rec.normal.z;
            rec.p.x = px;
            rec.p.z = pz;
            set_face_normal(rec, dir_i, V3<R>(nx, rec.normal.y, nz));
        } else { // hittable.h:168
            rec.front_face = !rec.front_face;
        }
    }
    return rec;
}

// ---- hit record of a planar primitive, digested once per scene ---------------------------------
// make_record() replays the wrapper chain op by op for every hit (two loops with a kind switch,
// a type switch, two divisions for u,v).  For an axis-aligned rect none of that depends on the
// ray: the chain maps the rect's plane to ONE world-space plane n.x = k, and what the reverse
// pass leaves in `normal` / `front_face` is a fixed function of the signs of d.v for at most two
// per-primitive vectors v.  Every translate / rotate_y wrapper re-runs set_face_normal with ITS
// ray against the normal as it stands (hittable.h:59,153), so
//   * the final normal is +-n with the sign the OUTERMOST such wrapper picks: it opposes
//     v_last = that wrapper's ray direction pulled back to the world.  For translate that is the
//     world direction itself (v_last = n); rotate_y tests its object-space ray against the
//     normal it has already rotated out (hittable.h:142-153), i.e. v_last = n turned once more;
//   * front_face says whether the normal that wrapper RECEIVED already opposed its ray: the signs
//     chosen by the outermost wrapper and by the one below it (or by the leaf: v_prev = n) agree;
//   * flip_face (hittable.h:168) only toggles front_face, so only flips above the outermost
//     translate / rotate_y survive.
// fp32 production only (the fused kernel); u,v are not produced, so materials whose textures read
// them stay on make_record().
struct PlaneRec {
    float n[3], k;  // world-space normal of the leaf's "outward" axis; plane n.x = k
    float vl[3];    // v_last
    uint32_t mode;  // bit 0: surviving flips (parity); bit 1: the chain has a translate / rotate_y
    float vp[3];    // v_prev
    uint32_t valid; // 0: not a planar primitive (or its chain is too long): use make_record()
    // filled by the fused kernel for scenes it shades from this record alone (solid colours only)
    float color[3];
    int32_t mat_type;
};
static_assert(sizeof(PlaneRec) == 64, "PlaneRec is four 16-byte vectors");

template <class R> RTB_HD void build_plane_rec(const GeomView<R> &g, uint32_t pi, PlaneRec &out) {
    out = PlaneRec{};
    const PrimT<R> p = g.prims[pi];
    const uint32_t type = p.type_mat & PT_TYPE_MASK;
    if (type != PT_XY && type != PT_XZ && type != PT_YZ)
        return;
    const int AX = type == PT_XY ? 2 : (type == PT_XZ ? 1 : 0);
    V3<R> n(0, 0, 0), q(0, 0, 0);
    n.set(AX, R(1));
    q.set(AX, p.d[4]);
    auto turn = [](const XfOp<R> &op, V3<R> v) { // the reverse pass of make_record() for rotate_y
        return V3<R>(op.b * v.x + op.a * v.z, v.y, -op.a * v.x + op.b * v.z);
    };
    uint32_t flips = 0;
    int last = -1, prev = -1, first = 0; // outermost and second-outermost translate / rotate_y
    const int chain = g.prim_chain[pi];
    if (chain >= 0) {
        const ChainRec c = g.chains[chain];
        if (c.count > kMaxChainOps)
            return;
        first = c.first;
        for (int i = c.count - 1; i >= 0; --i) {
            const XfOp<R> op = g.ops[first + i];
            if (op.kind == 0) {
                q = V3<R>(q.x + op.a, q.y + op.b, q.z + op.c);
            } else if (op.kind == 1) {
                q = turn(op, q);
                n = turn(op, n);
            }
            if (op.kind == 0 || op.kind == 1) {
                prev = last;
                last = i;
                flips = 0;
            } else {
                flips ^= 1u;
            }
        }
    }
    auto stage_vec = [&](int i) {
        if (i < 0)
            return n;
        const XfOp<R> op = g.ops[first + i];
        return op.kind == 1 ? turn(op, n) : n;
    };
    const V3<R> vl = stage_vec(last), vp = stage_vec(prev);
    for (int k = 0; k < 3; ++k) {
        out.n[k] = float(n[k]);
        out.vl[k] = float(vl[k]);
        out.vp[k] = float(vp[k]);
    }
    out.k = float(dot(n, q));
    out.mode = flips | (last >= 0 ? 2u : 0u);
    out.valid = 1;
}

template <class R> RTB_HD RecT<R> plane_record(const PlaneRec &pl, V3<R> o, V3<R> d, R t) {
    const V3<R> n(R(pl.n[0]), R(pl.n[1]), R(pl.n[2]));
    RecT<R> rec;
    rec.t = t;
    rec.u = 0;
    rec.v = 0;
    const V3<R> p = o + t * d;
    rec.p = p - (dot(n, p) - R(pl.k)) * n; // back onto the plane (exact for an axis-aligned normal)
    const bool a_last = dot(d, V3<R>(R(pl.vl[0]), R(pl.vl[1]), R(pl.vl[2]))) < 0;
    const bool a_prev = dot(d, V3<R>(R(pl.vp[0]), R(pl.vp[1]), R(pl.vp[2]))) < 0;
    rec.normal = a_last ? n : -n;
    rec.front_face = ((pl.mode & 2u) ? a_prev == a_last : a_last) != ((pl.mode & 1u) != 0);
    return rec;
}

} // namespace rtb

#endif // RTB_GEOM_CUH
