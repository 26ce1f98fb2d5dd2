/* rtb200_scene.h — the flat scene blob that crosses the C-ABI.
 *
 * The reference renderer (JiGuang283/Ray_Tracing-Rendering) keeps its scene as
 * a shared_ptr graph of virtual `hittable` / `material` / `texture` / `Light`
 * objects (src/geometry/hittable.h:25-32, src/materials/material.h:22-70,
 * src/materials/texture.h:11-33, src/lighting/light.h:15-47).  A GPU cannot
 * chase that graph, so the host side flattens it ONCE into the POD tables
 * declared here.  The tables are concatenated into one relocatable blob
 * (header + section directory + section payloads) so that the same bytes can
 * be (a) produced by the C++ host layer, (b) produced by the reference-walking
 * oracle harness (oracle/ref_harness.cpp), (c) committed as a test fixture and
 * (d) handed to rtb_scene_upload() unchanged.
 *
 * All real numbers are IEEE binary64, exactly the values the reference holds
 * (its vec3 is `double e[3]`, src/core/vec3.h:88).  The device library derives
 * its fp32 production tables and its fp64 validation tables from these.
 *
 * Everything in this header is plain C.
 */
#ifndef RTB200_SCENE_H
#define RTB200_SCENE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTB_SCENE_MAGIC 0x31424c46u /* "FLB1" little endian */
#define RTB_SCENE_VERSION 2u

/* ---- blob framing ------------------------------------------------------ */

typedef struct rtb_blob_header {
    uint32_t magic;       /* RTB_SCENE_MAGIC */
    uint32_t version;     /* RTB_SCENE_VERSION */
    uint64_t total_bytes; /* size of the whole blob */
    uint32_t n_sections;  /* number of rtb_blob_section records that follow */
    uint32_t reserved;
} rtb_blob_header;

typedef struct rtb_blob_section {
    uint32_t id;     /* enum rtb_section_id */
    uint32_t stride; /* bytes per record */
    uint64_t count;  /* number of records */
    uint64_t offset; /* byte offset of the payload from the blob start (8-byte aligned) */
} rtb_blob_section;

enum rtb_section_id {
    RTB_SEC_GLOBALS = 1,     /* 1 x rtb_globals */
    RTB_SEC_CAMERA = 2,      /* 1 x rtb_camera */
    RTB_SEC_PRIMS = 3,       /* rtb_prim[] */
    RTB_SEC_CHAINS = 4,      /* rtb_chain[] */
    RTB_SEC_XFORM_OPS = 5,   /* rtb_xform_op[] */
    RTB_SEC_MATERIALS = 6,   /* rtb_material[] */
    RTB_SEC_TEXTURES = 7,    /* rtb_texture[] */
    RTB_SEC_IMAGES = 8,      /* rtb_image[] (8-bit RGB image textures) */
    RTB_SEC_IMAGE_BYTES = 9, /* uint8_t[] pool the images index into */
    RTB_SEC_PERLIN = 10,     /* rtb_perlin[] */
    RTB_SEC_LIGHTS = 11,     /* rtb_light[] */
    RTB_SEC_ENV_TEXELS = 12, /* float[] pool (RGB32F) the env lights index into */
    RTB_SEC_GATES = 13       /* rtb_gate[] (optional: absent = none) */
};

/* ---- globals & camera -------------------------------------------------- */

/* SceneConfig of the reference (src/scene/scenes.h:11-24) minus the graph. */
typedef struct rtb_globals {
    double background[3];
    int32_t image_width;       /* SceneConfig::image_width */
    int32_t image_height;      /* int(width / aspect_ratio), src/main.cpp:69 */
    int32_t samples_per_pixel; /* SceneConfig::samples_per_pixel */
    int32_t scene_id;          /* id passed to select_scene(), or -1 */
} rtb_globals;

/* Constructor arguments of `camera` (src/renderer/camera.h:9-11); the derived
 * members (origin, lower_left_corner, ...) are recomputed from these in fp64
 * by every consumer, following camera.h:12-30 operation by operation. */
typedef struct rtb_camera {
    double lookfrom[3];
    double lookat[3];
    double vup[3];
    double vfov;
    double aspect_ratio;
    double aperture;
    double focus_dist;
    double time0;
    double time1;
} rtb_camera;

/* ---- geometry ---------------------------------------------------------- */

enum rtb_prim_type {
    RTB_PRIM_SPHERE = 0,        /* src/geometry/sphere.h */
    RTB_PRIM_MOVING_SPHERE = 1, /* src/geometry/moving_sphere.h */
    RTB_PRIM_XY_RECT = 2,       /* src/geometry/aarect.h:11-32  (normal +Z) */
    RTB_PRIM_XZ_RECT = 3,       /* src/geometry/aarect.h:34-54  (normal +Y) */
    RTB_PRIM_YZ_RECT = 4,       /* src/geometry/aarect.h:56-76  (normal +X) */
    RTB_PRIM_MEDIUM = 5         /* src/geometry/constant_medium.h:28-53 */
};

enum rtb_prim_flags {
    /* The primitive only exists as the boundary of a constant_medium; it is
     * not itself a member of the world (constant_medium.h:50). */
    RTB_PRIM_BOUNDARY_ONLY = 1,
    /* The reference stores the object as BOTH children of a one-object
     * bvh_node (src/geometry/bvh.h:68-69) and therefore tests it twice per
     * ray.  Only observable for stochastic primitives (media). */
    RTB_PRIM_DUP_LEAF = 2,
    /* The primitive has an rtb_gate: it is only tested for rays that pass the gate's box. */
    RTB_PRIM_GATED = 4
};

/* One leaf of the reference graph.  A `box` (src/geometry/box.h) contributes
 * its six rects; `hittable_list` and `bvh_node` contribute their members.
 *
 * d[] by type:
 *   SPHERE         cx cy cz radius
 *   MOVING_SPHERE  c0x c0y c0z c1x c1y c1z time0 time1 radius
 *   XY_RECT        x0 x1 y0 y1 k
 *   XZ_RECT        x0 x1 z0 z1 k
 *   YZ_RECT        y0 y1 z0 z1 k
 *   MEDIUM         neg_inv_density ; aux0 = first boundary prim, aux1 = count
 * `material` indexes rtb_material[] (MEDIUM: its isotropic phase function).
 * `chain` indexes rtb_chain[] or is -1: the wrapper objects (translate,
 * rotate_y, flip_face) between the world root and this leaf.
 * The index of a record in the PRIMS section is the primitive id reported by
 * rtb_trace_batch() and by the oracle. */
typedef struct rtb_prim {
    int32_t type;
    int32_t material;
    int32_t chain;
    int32_t flags;
    int32_t aux0;
    int32_t aux1;
    double d[9];
} rtb_prim;

/* A sphere of NEGATIVE radius (the hollow-glass idiom, e.g. scenes.cpp:903) reports an inverted
 * bounding box (sphere.h:62-66: centre -/+ radius), which surrounding_box() folds into its parent
 * bvh_node's box as two points well inside the sphere.  The reference therefore only tests such a
 * sphere for rays whose interval overlaps THAT node's box (bvh.h:40-50; every ancestor's box contains
 * it) — a property of the tree the reference happened to build.  A flattener that walks a reference
 * graph records that box here (same object space as the sphere); a sphere that is the only child of
 * its node gets the inverted box itself, i.e. is never hit (aabb.h:31-48 fails on every ray).
 * Scenes authored without a reference tree simply have no gates. */
typedef struct rtb_gate {
    int32_t prim; /* index into PRIMS: an RTB_PRIM_SPHERE carrying RTB_PRIM_GATED */
    int32_t reserved;
    double lo[3]; /* aabb::minimum of the bvh_node holding the sphere */
    double hi[3]; /* aabb::maximum */
} rtb_gate;

enum rtb_xform_kind {
    RTB_XF_TRANSLATE = 0, /* a,b,c = offset              (hittable.h:34-75) */
    RTB_XF_ROTATE_Y = 1,  /* a = sin_theta, b = cos_theta (hittable.h:77-156) */
    RTB_XF_FLIP_FACE = 2  /* no parameters                (hittable.h:158-179) */
};

typedef struct rtb_xform_op {
    int32_t kind;
    int32_t reserved;
    double a, b, c;
} rtb_xform_op;

/* ops [first, first+count) listed OUTERMOST wrapper first, i.e. in the order
 * the reference transforms the ray on its way down (hittable.h:51-62,127-156). */
typedef struct rtb_chain {
    int32_t first;
    int32_t count;
} rtb_chain;

/* ---- materials & textures ---------------------------------------------- */

enum rtb_material_type {
    RTB_MAT_LAMBERTIAN = 0,    /* material.h:72-116 */
    RTB_MAT_METAL = 1,         /* material.h:118-145 */
    RTB_MAT_DIELECTRIC = 2,    /* material.h:147-204 */
    RTB_MAT_DIFFUSE_LIGHT = 3, /* material.h:206-236 */
    RTB_MAT_PBR = 4,           /* material.h:238-439 */
    RTB_MAT_ISOTROPIC = 5,     /* constant_medium.h:12-26 */
    RTB_MAT_TYPE_COUNT = 6
};

/* tex[0] albedo / emit, tex[1] roughness, tex[2] metallic, tex[3] normal map
 * (-1 = absent).  `color`+`fuzz` are metal's albedo and fuzz, `ir` is the
 * dielectric index. */
typedef struct rtb_material {
    int32_t type;
    int32_t tex[4];
    int32_t reserved;
    double color[3];
    double fuzz;
    double ir;
} rtb_material;

enum rtb_texture_type {
    RTB_TEX_SOLID = 0,   /* texture.h:35-54 */
    RTB_TEX_CHECKER = 1, /* texture.h:56-80   even/odd index rtb_texture[] */
    RTB_TEX_IMAGE = 2,   /* texture.h:82-146  image indexes rtb_image[]   */
    RTB_TEX_NOISE = 3    /* texture.h:148-162 perlin indexes rtb_perlin[] */
};

typedef struct rtb_texture {
    int32_t type;
    int32_t even;   /* CHECKER */
    int32_t odd;    /* CHECKER */
    int32_t image;  /* IMAGE  */
    int32_t perlin; /* NOISE  */
    int32_t reserved;
    double color[3]; /* SOLID */
    double scale;    /* NOISE */
} rtb_texture;

/* width == 0 means "file missing": the reference then returns cyan
 * (texture.h:116-118). */
typedef struct rtb_image {
    int32_t width, height;
    uint64_t offset; /* into IMAGE_BYTES, 3 bytes per texel, row-major */
} rtb_image;

/* perlin.h:10-20 : 256 unit gradient vectors + three permutation tables. */
typedef struct rtb_perlin {
    double ranvec[256][3];
    int32_t perm_x[256];
    int32_t perm_y[256];
    int32_t perm_z[256];
} rtb_perlin;

/* ---- lights ------------------------------------------------------------ */

enum rtb_light_type {
    RTB_LIGHT_QUAD = 0,        /* lighting/quad_light.h */
    RTB_LIGHT_POINT = 1,       /* lighting/point_light.h */
    RTB_LIGHT_SPOT = 2,        /* lighting/spot_light.h */
    RTB_LIGHT_DIRECTIONAL = 3, /* lighting/directional_light.h */
    RTB_LIGHT_ENV = 4          /* lighting/environmental_light.h */
};

/* QUAD: Q,u,v,intensity as given to the constructor; normal and area are
 *       recomputed by the consumer (quad_light.h:9-16).
 * POINT: position = Q, intensity.
 * SPOT: position = Q, direction = u (already unit), cos_cutoff, intensity.
 * DIRECTIONAL: direction = u (already unit), radiance = intensity.
 * ENV: env_width/env_height texels at env_offset floats into ENV_TEXELS
 *      (width 0 = file missing => constant white, environmental_light.h:127-132);
 *      env_is_probe mirrors is_light_probe (environmental_light.h:138-140). */
typedef struct rtb_light {
    int32_t type;
    int32_t env_width;
    int32_t env_height;
    int32_t env_is_probe;
    uint64_t env_offset;
    double Q[3];
    double u[3];
    double v[3];
    double intensity[3];
    double cos_cutoff;
} rtb_light;

#ifdef __cplusplus
}
#endif

#endif /* RTB200_SCENE_H */
