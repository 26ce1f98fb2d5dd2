/* rtb200_types.h — POD records exchanged by the batch (parity-layer) entry
 * points of the C-ABI.  Shared by the product library (include/rtb200.h) and
 * by the test oracles so both sides speak the same layout.  Plain C. */
#ifndef RTB200_TYPES_H
#define RTB200_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One query against the world: the arguments of hittable::hit()
 * (src/geometry/hittable.h:28-29) plus the ray's time (src/core/ray.h:9).
 * origin_prim is the id of the primitive the ray leaves (-1 = none); the fp64
 * validation path ignores it, the fp32 production path uses it to suppress
 * self-intersection. */
typedef struct rtb_ray {
    double o[3];
    double d[3];
    double time;
    double t_min;
    double t_max;
    int32_t origin_prim;
    int32_t reserved; /* 0, or (rtb_trace_batch precision 65) the reference's xorshift32 state before the query */
} rtb_ray;

/* hit_record of the reference (src/geometry/hittable.h:10-23) + the flattened
 * primitive id and material index.  prim == -1 means "no hit" and every other
 * field is then zero. */
typedef struct rtb_hit {
    double t;
    double p[3];
    double normal[3];
    double u, v;
    int32_t prim;
    int32_t front_face;
    int32_t material;
    int32_t reserved;
} rtb_hit;

/* Inputs of material::{eval,pdf,sample,scatter,emitted}
 * (src/materials/material.h:27-69): the part of hit_record a material reads,
 * the outgoing direction wo (unit, pointing away from the surface) and the
 * incident direction wi. */
typedef struct rtb_bsdf_query {
    double p[3];
    double normal[3];
    double u, v;
    double wo[3];
    double wi[3];
    int32_t front_face;
    int32_t reserved;
} rtb_bsdf_query;

/* eval() and pdf() of one query, and both emitted() overloads. */
typedef struct rtb_bsdf_value {
    double f[3];
    double pdf;
    double emitted_old[3]; /* emitted(u,v,p)   material.h:27-29 */
    double emitted_new[3]; /* emitted(rec,wo)  material.h:32-34 */
} rtb_bsdf_value;

/* BSDFSample (material.h:13-20) + whether sample() returned true, and the
 * result of the legacy scatter() (material.h:66-69). */
typedef struct rtb_bsdf_sample {
    double wi[3];
    double f[3];
    double pdf;
    int32_t ok;
    int32_t is_specular;
    double scatter_dir[3];
    double scatter_atten[3];
    int32_t scatter_ok;
    int32_t reserved;
} rtb_bsdf_sample;

/* Light::sample() input (light.h:21) */
typedef struct rtb_light_query {
    double p[3];   /* shading point (sample) / ray origin (pdf, Le) */
    double d[3];   /* ray direction for pdf() and Le(); unused by sample() */
    double u[2];   /* the two uniform numbers handed to sample() */
} rtb_light_query;

/* LightSample (light.h:7-13) + pdf(origin,direction) + Le(ray) */
typedef struct rtb_light_value {
    double Li[3];
    double wi[3];
    double pdf;
    double dist;
    int32_t is_delta;
    int32_t reserved;
    double pdf_dir; /* Light::pdf(p, d) */
    double Le[3];   /* Light::Le(ray(p, d)) */
} rtb_light_value;

#ifdef __cplusplus
}
#endif

#endif /* RTB200_TYPES_H */
