// rtb_shading.cuh — RNG, samplers, textures, materials, lights and the camera,
// templated on R (float production / double validation).
//
// Reference counterparts:
//   random_double & samplers   src/core/rtweekend.h:24-50, src/core/vec3.h:226-269
//   textures                   src/materials/texture.h:35-162, src/materials/perlin.h:22-111
//   materials                  src/materials/material.h:72-439, src/geometry/constant_medium.h:12-26
//   lights                     src/lighting/*.h
//   camera                     src/renderer/camera.h:9-40
//
// The reference draws from a thread-local xorshift32 whose seed depends on the
// thread id (rtweekend.h:26-27) and is not reproducible; the GPU uses its own
// generator, so SAMPLING is matched in distribution (rejection loops become
// their closed-form equivalents) while eval / pdf / emitted / Light::sample
// are matched value for value.
#ifndef RTB_SHADING_CUH
#define RTB_SHADING_CUH

#include "rtb_geom.cuh"

namespace rtb {

// ---- RNG -------------------------------------------------------------------------------
// PCG-XSH-RR 64/32 (O'Neill 2014): 64-bit LCG state, 32-bit permuted output.  One
// state per path, seeded from (global sample index, seed), so a sample's random
// stream does not depend on which slot, SM or GPU runs it.
struct Pcg {
    uint64_t s;
    RTB_HD uint32_t next_u32() {
        const uint64_t old = s;
        s = old * 6364136223846793005ULL + 1442695040888963407ULL;
        const uint32_t x = uint32_t(((old >> 18u) ^ old) >> 27u);
        const uint32_t r = uint32_t(old >> 59u);
        return (x >> r) | (x << ((32u - r) & 31u));
    }
    // The float draws take the TOP 24 bits of the LCG state directly (the bits with the longest
    // period; the XSH-RR permutation exists to make the LOW bits usable, which a 24-bit float draw
    // never touches): 5 instructions per draw instead of 12 on a path that is instruction-issue
    // bound (ncu source view of k_fused on the Cornell box: 10 % of all issued instructions were
    // next_u32()).  Integer draws keep the permuted output.
    RTB_HD uint32_t next_top24() {
        const uint64_t old = s;
        s = old * 6364136223846793005ULL + 1442695040888963407ULL;
        return uint32_t(old >> 40u);
    }
    // uniform in [0,1), 24 bits
    RTB_HD float next_f() { return float(next_top24()) * (1.0f / 16777216.0f); }
    // uniform in (0,1): safe argument for log()
    RTB_HD float next_open() { return (float(next_top24()) + 0.5f) * (1.0f / 16777216.0f); }
};

RTB_HD uint64_t mix64(uint64_t z) { // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
// seed_mixed = mix64(seed): the same for every sample of a render, so the renderer mixes it once
// on the host
RTB_HD Pcg pcg_seed_premixed(uint64_t sample_index, uint64_t seed_mixed) {
    Pcg r;
    r.s = mix64(sample_index * 0x9e3779b97f4a7c15ULL + seed_mixed);
    r.next_u32();
    return r;
}
RTB_HD Pcg pcg_seed(uint64_t sample_index, uint64_t seed) { return pcg_seed_premixed(sample_index, mix64(seed)); }

template <class R> struct RngT {
    Pcg g;
    RTB_HD R next() { return R(g.next_f()); }
    RTB_HD R next_open() { return R(g.next_open()); }
};

// ---- samplers (vec3.h:226-269) -----------------------------------------------------------
// uniform on the unit sphere == unit_vector(random_in_unit_sphere())
template <class R, class G> RTB_HD V3<R> random_unit_vector(G &g) {
    const R z = R(1) - R(2) * g.next();
    const R r = sqrt_(fmax_(R(0), R(1) - z * z));
    R s, c;
    sincos_(R(2) * Consts<R>::pi() * g.next(), &s, &c);
    return V3<R>(r * c, r * s, z);
}
// uniform in the unit ball == the rejection loop of vec3.h:226-233
template <class R, class G> RTB_HD V3<R> random_in_unit_sphere(G &g) {
    const R rad = cbrt_(g.next_open()); // (0,1): the rejection loop never returns the centre itself
    return rad * random_unit_vector<R>(g);
}
// uniform in the unit disk == vec3.h:250-257
template <class R, class G> RTB_HD V3<R> random_in_unit_disk(G &g) {
    const R rad = sqrt_(g.next());
    R s, c;
    sincos_(R(2) * Consts<R>::pi() * g.next(), &s, &c);
    return V3<R>(rad * c, rad * s, 0);
}
// vec3.h:261-269
template <class R, class G> RTB_HD V3<R> random_cosine_direction(G &g) {
    const R r1 = g.next();
    const R r2 = g.next();
    const R z = sqrt_(R(1) - r2);
    R s, c;
    sincos_(R(2) * Consts<R>::pi() * r1, &s, &c);
    const R sr2 = sqrt_(r2);
    return V3<R>(c * sr2, s * sr2, z);
}

// ---- device tables -----------------------------------------------------------------------
template <class R> struct MatT {
    int32_t type;   // rtb_material_type
    int32_t tex[4]; // albedo/emit, roughness, metallic, normal map
    int32_t flags;  // bit0: some texture of this material reads u,v; bits 1,2: baked solid textures (mat_tex)
    R color[3];     // metal albedo
    R fuzz;
    R ir;
};
template <class R> struct TexT {
    int32_t type; // rtb_texture_type
    int32_t even, odd, image, perlin;
    R color[3];
    R scale;
};
struct ImageRec {
    int32_t width, height;
    uint64_t offset;
};
template <class R> struct PerlinT {
    R ranvec[256][3];
    int32_t perm_x[256], perm_y[256], perm_z[256];
};
template <class R> struct LightT {
    int32_t type; // rtb_light_type
    int32_t env_w, env_h, env_probe;
    uint64_t env_texel_offset; // into env_texels (floats)
    uint64_t env_table_offset; // into env_tables (doubles), see EnvTables
    R Q[3], u[3], v[3], intensity[3], normal[3];
    R area, cos_cutoff;
};
template <class R> struct CameraT {
    V3<R> origin, lower_left_corner, horizontal, vertical, u, v, w;
    R lens_radius, time0, time1;
};

template <class R> struct ShadeView {
    const MatT<R> *mats;
    const TexT<R> *texs;
    const ImageRec *images;
    const uint8_t *image_bytes;
    const PerlinT<R> *perlins;
    const LightT<R> *lights;
    const float *env_texels;
    const double *env_tables;
    int32_t n_lights;
    int32_t n_infinite_lights;
};

// ---- textures ----------------------------------------------------------------------------
// perlin.h:22-42, 96-111
template <class R> RTB_HD R perlin_noise(const PerlinT<R> &P, V3<R> p) {
    const R fx = floor_(p.x), fy = floor_(p.y), fz = floor_(p.z);
    const R u = p.x - fx, v = p.y - fy, w = p.z - fz;
    const int i = int(fx), j = int(fy), k = int(fz);
    const R uu = u * u * (R(3) - R(2) * u);
    const R vv = v * v * (R(3) - R(2) * v);
    const R ww = w * w * (R(3) - R(2) * w);
    R accum = 0;
    for (int di = 0; di < 2; di++)
        for (int dj = 0; dj < 2; dj++)
            for (int dk = 0; dk < 2; dk++) {
                const int idx = P.perm_x[(i + di) & 255] ^ P.perm_y[(j + dj) & 255] ^
                                P.perm_z[(k + dk) & 255];
                const V3<R> c(P.ranvec[idx][0], P.ranvec[idx][1], P.ranvec[idx][2]);
                const V3<R> weight_v(u - di, v - dj, w - dk);
                accum += (di * uu + (1 - di) * (R(1) - uu)) * (dj * vv + (1 - dj) * (R(1) - vv)) *
                         (dk * ww + (1 - dk) * (R(1) - ww)) * dot(c, weight_v);
            }
    return accum;
}
// perlin.h:44-56
template <class R> RTB_HD R perlin_turb(const PerlinT<R> &P, V3<R> p) {
    R accum = 0, weight = 1;
    V3<R> tp = p;
#pragma unroll 1
    for (int i = 0; i < 7; i++) {
        accum += weight * perlin_noise(P, tp);
        weight *= R(0.5);
        tp = R(2) * tp;
    }
    return fabs_(accum);
}

// Out of line, with the tables passed by pointer: mat_tex() reaches it from every BSDF entry point (albedo,
// roughness, metallic, normal map, emission), and the inlined copies — seven unrolled octaves of Perlin noise
// each — were 27 % of the general fused kernel's 20,000 instructions, a kernel ncu showed waiting for
// instructions (`no_instruction`, issue slots 30 % on C4-env).
template <class R>
RTB_HD_OUTLINE V3<R> tex_value_tables(const TexT<R> *texs, const ImageRec *images, const uint8_t *image_bytes,
                                      const PerlinT<R> *perlins, int tex, R u, R v, V3<R> p) {
    for (int hop = 0; hop < 8; ++hop) {
        const TexT<R> &t = texs[tex];
        switch (t.type) {
        case 0: // solid_color, texture.h:47-49
            return V3<R>(t.color[0], t.color[1], t.color[2]);
        case 1: { // checker_texture, texture.h:70-77
            const R sines = sin_(R(10) * p.x) * sin_(R(10) * p.y) * sin_(R(10) * p.z);
            tex = sines < 0 ? t.odd : t.even;
            break;
        }
        case 2: { // image_texture, texture.h:115-139
            const ImageRec im = images[t.image];
            if (im.width == 0)
                return V3<R>(0, 1, 1);
            const R uc = clamp_(u, R(0), R(1));
            const R vc = R(1) - clamp_(v, R(0), R(1));
            int i = int(uc * im.width), j = int(vc * im.height);
            if (i >= im.width)
                i = im.width - 1;
            if (j >= im.height)
                j = im.height - 1;
            const uint8_t *px = image_bytes + im.offset + (size_t(j) * im.width + i) * 3;
            const R s = R(1.0 / 255.0);
            return V3<R>(s * px[0], s * px[1], s * px[2]);
        }
        default: { // noise_texture, texture.h:155-158
            const R n = R(0.5) * (R(1) + sin_(t.scale * p.z + R(10) * perlin_turb(perlins[t.perlin], p)));
            return V3<R>(n, n, n);
        }
        }
    }
    return V3<R>(0, 0, 0);
}
template <class R> RTB_HD V3<R> tex_value(const ShadeView<R> &S, int tex, R u, R v, V3<R> p) {
    return tex_value_tables<R>(S.texs, S.images, S.image_bytes, S.perlins, tex, u, v, p);
}

// ---- materials ---------------------------------------------------------------------------
template <class R> struct BsdfSampleT {
    V3<R> wi, f;
    R pdf;
    bool is_specular;
};

// dielectric::reflectance, material.h:199-203
template <class R> RTB_HD R schlick_reflectance(R cosine, R ref_idx) {
    R r0 = (R(1) - ref_idx) / (R(1) + ref_idx);
    r0 = r0 * r0;
    return r0 + (R(1) - r0) * pow5_(R(1) - cosine);
}

// PBRMaterial helpers, material.h:397-432
RTB_HD double ggx_D(V3<double> N, V3<double> H, double roughness) {
    const double a = roughness * roughness;
    const double a2 = a * a;
    const double NdotH = fmax_(dot(N, H), 0.0);
    const double NdotH2 = NdotH * NdotH;
    double denom = (NdotH2 * (a2 - 1.0) + 1.0);
    denom = Consts<double>::pi() * denom * denom;
    return a2 / denom;
}
// fp32 production form of the same function.  NdotH2*(a2-1)+1 == (1-NdotH2) + a2*NdotH2,
// and with a2 as small as 1e-8 (roughness clamp 0.01) the reference's form cancels
// catastrophically in fp32 (a2-1 rounds to -1, denom to 0).  1-NdotH2 = |N x H|^2 is
// evaluated from the cross product, which keeps full relative precision near the peak.
RTB_HD float ggx_D(V3<float> N, V3<float> H, float roughness) {
    const float a = roughness * roughness;
    const float a2 = a * a;
    const float NdotH = fmax_(dot(N, H), 0.0f);
    const float sin2 = NdotH > 0.0f ? length_squared(cross(N, H)) : 1.0f;
    float denom = sin2 + a2 * (NdotH * NdotH);
    denom = Consts<float>::pi() * denom * denom;
    return a2 / denom;
}
template <class R> RTB_HD R ggx_G1(R NdotV, R roughness) {
    const R k = (roughness * roughness) / R(2);
    return NdotV / (NdotV * (R(1) - k) + k);
}
template <class R> RTB_HD R ggx_smith(V3<R> N, V3<R> V, V3<R> L, R roughness) {
    const R NdotV = fmax_(dot(N, V), R(0));
    const R NdotL = fmax_(dot(N, L), R(0));
    const R ggx2 = ggx_G1(NdotV, roughness);
    const R ggx1 = ggx_G1(NdotL, roughness);
    return ggx1 * ggx2;
}

// PBRMaterial's shading normal, material.h:247-262 (tangent frame from world up)
// Texture slot k of a material.  Solid-colour textures are baked into the material record at
// upload (flags bit 1: slot 0 lives in color[]; bit 2: the PBR roughness / metallic scalars live
// in fuzz / ir), which removes one dependent random gather per hit from the shade kernels (with
// one material and one texture record per sphere, C5's tables are far larger than any cache).
template <class R>
RTB_HD V3<R> mat_tex(const ShadeView<R> &S, const MatT<R> &m, int k, const RecT<R> &rec) {
    if (k == 0 && (m.flags & 2))
        return V3<R>(m.color[0], m.color[1], m.color[2]);
    if (k == 1 && (m.flags & 4))
        return V3<R>(m.fuzz, m.fuzz, m.fuzz);
    if (k == 2 && (m.flags & 4))
        return V3<R>(m.ir, m.ir, m.ir);
    return tex_value(S, m.tex[k], rec.u, rec.v, rec.p);
}

template <class R>
RTB_HD V3<R> pbr_normal(const ShadeView<R> &S, const MatT<R> &m, const RecT<R> &rec) {
    V3<R> N = rec.normal;
    if (m.tex[3] >= 0) {
        V3<R> ax0;
        if (fabs_(N.y) > R(0.999))
            ax0 = V3<R>(1, 0, 0);
        else
            ax0 = unit_vector(cross(N, V3<R>(0, 1, 0)));
        const V3<R> ax1 = cross(N, ax0);
        const V3<R> c = mat_tex(S, m, 3, rec);
        const V3<R> local_n = unit_vector(c * R(2) - V3<R>(1, 1, 1)); // texture.h:19-22
        N = unit_vector(local_n.x * ax0 + local_n.y * ax1 + local_n.z * N);
    }
    return N;
}

// material::pdf  (lambertian material.h:92-96, PBR material.h:310-344, others 0)
template <class R>
RTB_HD R mat_pdf(const ShadeView<R> &S, const MatT<R> &m, const RecT<R> &rec, V3<R> wo, V3<R> wi) {
    const R pi = Consts<R>::pi();
    if (m.type == 0) {
        const R cosine = dot(rec.normal, unit_vector(wi));
        return cosine < 0 ? R(0) : cosine / pi;
    }
    if (m.type == 4) {
        const V3<R> N = pbr_normal(S, m, rec);
        if (dot(N, wi) <= 0)
            return 0;
        R rough = mat_tex(S, m, 1, rec).x;
        rough = clamp_(rough, R(0.01), R(1));
        const R pdf_diff = dot(N, wi) / pi;
        const V3<R> H = unit_vector(wo + wi);
        const R D = ggx_D(N, H, rough);
        const R NdotH = fmax_(dot(N, H), R(0));
        const R HdotV = fmax_(dot(H, wo), R(0));
        const R pdf_spec = (D * NdotH) / (R(4) * HdotV + R(0.0001));
        return R(0.5) * pdf_diff + R(0.5) * pdf_spec;
    }
    return 0;
}

// material::eval  (lambertian material.h:98-101, PBR material.h:346-395, others 0)
template <class R>
RTB_HD V3<R> mat_eval(const ShadeView<R> &S, const MatT<R> &m, const RecT<R> &rec, V3<R> wo,
                      V3<R> wi) {
    const R pi = Consts<R>::pi();
    if (m.type == 0)
        return mat_tex(S, m, 0, rec) / pi;
    if (m.type == 4) {
        const V3<R> N = pbr_normal(S, m, rec);
        const R NdotL = dot(N, wi);
        const R NdotV = dot(N, wo);
        if (NdotL <= 0 || NdotV <= 0)
            return V3<R>(0, 0, 0);
        R rough = mat_tex(S, m, 1, rec).x;
        const R metal = mat_tex(S, m, 2, rec).x;
        const V3<R> base = mat_tex(S, m, 0, rec);
        rough = clamp_(rough, R(0.01), R(1));
        const V3<R> H = unit_vector(wo + wi);
        const V3<R> one(1, 1, 1);
        const V3<R> mv(metal, metal, metal);
        const V3<R> F0 = (one - mv) * V3<R>(R(0.04), R(0.04), R(0.04)) + mv * base;
        const R ct = fmax_(dot(H, wo), R(0));
        const V3<R> F = F0 + (one - F0) * pow5_(R(1) - ct);
        const R D = ggx_D(N, H, rough);
        const R G = ggx_smith(N, wo, wi, rough);
        const V3<R> numerator = (D * G) * F;
        const R denominator = R(4) * NdotV * NdotL + R(0.0001);
        const V3<R> specular = numerator / denominator;
        V3<R> kD = one - F;
        kD = kD * (R(1) - metal);
        const V3<R> diffuse = (kD * base) / pi;
        return diffuse + specular;
    }
    return V3<R>(0, 0, 0);
}

// material::emitted(u,v,p)  — the legacy, two-sided API (material.h:27-29, 220-222)
template <class R>
RTB_HD V3<R> mat_emitted_old(const ShadeView<R> &S, const MatT<R> &m, const RecT<R> &rec) {
    if (m.type == 3)
        return mat_tex(S, m, 0, rec);
    return V3<R>(0, 0, 0);
}
// material::emitted(rec,wo) — front face only (material.h:224-229)
template <class R>
RTB_HD V3<R> mat_emitted_new(const ShadeView<R> &S, const MatT<R> &m, const RecT<R> &rec) {
    if (m.type == 3 && rec.front_face)
        return mat_tex(S, m, 0, rec);
    return V3<R>(0, 0, 0);
}

// material::sample(rec, wo, sampled)  (material.h:41-44 and overrides)
template <class R, class G>
RTB_HD bool mat_sample(const ShadeView<R> &S, const MatT<R> &m, const RecT<R> &rec, V3<R> wo, G &g,
                       BsdfSampleT<R> &bs) {
    const R pi = Consts<R>::pi();
    switch (m.type) {
    case 0: { // lambertian, material.h:79-90
        V3<R> dir = rec.normal + random_unit_vector<R>(g);
        if (near_zero(dir))
            dir = rec.normal;
        bs.wi = unit_vector(dir);
        bs.pdf = dot(rec.normal, bs.wi) / pi;
        bs.f = mat_tex(S, m, 0, rec) / pi;
        bs.is_specular = false;
        return true;
    }
    case 1: { // metal, material.h:123-131
        const V3<R> reflected = reflect(unit_vector(-wo), rec.normal);
        bs.wi = unit_vector(reflected + m.fuzz * random_in_unit_sphere<R>(g));
        bs.f = V3<R>(m.color[0], m.color[1], m.color[2]);
        bs.pdf = 1;
        bs.is_specular = true;
        return dot(bs.wi, rec.normal) > 0;
    }
    case 2: { // dielectric, material.h:152-174
        bs.f = V3<R>(1, 1, 1);
        bs.is_specular = true;
        bs.pdf = 1;
        const R ratio = rec.front_face ? (R(1) / m.ir) : m.ir;
        const V3<R> ud = -wo;
        const R cos_theta = fmin_(dot(-ud, rec.normal), R(1));
        const R sin_theta = sqrt_(R(1) - cos_theta * cos_theta);
        const bool cannot_refract = ratio * sin_theta > R(1);
        if (cannot_refract || schlick_reflectance(cos_theta, ratio) > g.next())
            bs.wi = reflect(ud, rec.normal);
        else
            bs.wi = refract(ud, rec.normal, ratio);
        return true;
    }
    case 4: { // PBRMaterial, material.h:245-308
        const V3<R> N = pbr_normal(S, m, rec);
        R rough = mat_tex(S, m, 1, rec).x;
        rough = clamp_(rough, R(0.01), R(1));
        if (g.next() < R(0.5)) {
            Onb<R> uvw;
            uvw.build_from_w(N);
            const R r1 = g.next();
            const R r2 = g.next();
            const R a = rough * rough;
            const R phi = R(2) * pi * r1;
            R cos_theta, sin_theta;
            if (sizeof(R) == 8) { // material.h:275-276 as written
                cos_theta = sqrt_((R(1) - r2) / (R(1) + (a * a - R(1)) * r2));
                sin_theta = sqrt_(R(1) - cos_theta * cos_theta);
            } else { // same values, without the fp32 cancellation in 1+(a^2-1)r2 and 1-cos^2
                const R den = (R(1) - r2) + (a * a) * r2;
                cos_theta = sqrt_((R(1) - r2) / den);
                sin_theta = sqrt_(((a * a) * r2) / den);
            }
            R sp, cp;
            sincos_(phi, &sp, &cp);
            const V3<R> H = uvw.local(V3<R>(sin_theta * cp, sin_theta * sp, cos_theta));
            const V3<R> L = reflect(-wo, H);
            if (dot(N, L) <= 0)
                return false;
            bs.wi = L;
        } else {
            Onb<R> uvw;
            uvw.build_from_w(N);
            V3<R> L = uvw.local(random_cosine_direction<R>(g));
            if (dot(N, L) <= 0)
                L = N;
            bs.wi = unit_vector(L);
        }
        bs.is_specular = false;
        bs.pdf = mat_pdf(S, m, rec, wo, bs.wi);
        bs.f = mat_eval(S, m, rec, wo, bs.wi);
        return !(bs.pdf < R(1e-6));
    }
    default: // diffuse_light (material.h:213-216), isotropic (base class): false
        return false;
    }
}

// legacy material::scatter(r_in, rec, attenuation, scattered)  (material.h:66-69 and overrides);
// `din` is the incoming ray direction (NOT normalised), the result direction is not
// normalised either.
template <class R, class G>
RTB_HD bool mat_scatter(const ShadeView<R> &S, const MatT<R> &m, const RecT<R> &rec, V3<R> din, G &g,
                        V3<R> &atten, V3<R> &dout) {
    switch (m.type) {
    case 0: { // material.h:103-112
        V3<R> dir = rec.normal + random_unit_vector<R>(g);
        if (near_zero(dir))
            dir = rec.normal;
        dout = dir;
        atten = mat_tex(S, m, 0, rec);
        return true;
    }
    case 1: { // material.h:133-140
        const V3<R> reflected = reflect(unit_vector(din), rec.normal);
        dout = reflected + m.fuzz * random_in_unit_sphere<R>(g);
        atten = V3<R>(m.color[0], m.color[1], m.color[2]);
        return dot(dout, rec.normal) > 0;
    }
    case 2: { // material.h:176-193
        atten = V3<R>(1, 1, 1);
        const R ratio = rec.front_face ? (R(1) / m.ir) : m.ir;
        const V3<R> ud = unit_vector(din);
        const R cos_theta = fmin_(dot(-ud, rec.normal), R(1));
        const R sin_theta = sqrt_(R(1) - cos_theta * cos_theta);
        const bool cannot_refract = ratio * sin_theta > R(1);
        if (cannot_refract || schlick_reflectance(cos_theta, ratio) > g.next())
            dout = reflect(ud, rec.normal);
        else
            dout = refract(ud, rec.normal, ratio);
        return true;
    }
    case 5: // isotropic, constant_medium.h:19-24
        dout = random_in_unit_sphere<R>(g);
        atten = mat_tex(S, m, 0, rec);
        return true;
    default: // diffuse_light (material.h:231-234), PBRMaterial (no override)
        return false;
    }
}

// ---- lights ------------------------------------------------------------------------------
template <class R> struct LightSampleT {
    V3<R> Li, wi;
    V3<R> to_light; // light_point - p, NOT normalised (finite lights only; 0 otherwise)
    R pdf, dist;
    bool is_delta;
};

// Per env light, `env_tables + env_table_offset` holds (W = env_w, H = env_h):
//   cond_func [H*W]     luminance*sin(theta)            environmental_light.h:146-170
//   cond_cdf  [H*(W+1)] normalised running sums          environmental_light.h:15-27
//   cond_int  [H]       row integrals (un-normalised)
//   marg_cdf  [H+1]
//   marg_int  [1]
// (the marginal's func IS cond_int, environmental_light.h:78)
struct EnvTables {
    const double *cond_func, *cond_cdf, *cond_int, *marg_cdf;
    double marg_int;
    RTB_HD EnvTables(const double *base, int W, int H) {
        cond_func = base;
        cond_cdf = cond_func + size_t(W) * H;
        cond_int = cond_cdf + size_t(W + 1) * H;
        marg_cdf = cond_int + H;
        marg_int = marg_cdf[H + 1];
    }
    static size_t doubles(int W, int H) { return size_t(W) * H + size_t(W + 1) * H + H + (H + 1) + 1; }
};

// Distribution1D::sample, environmental_light.h:30-44
RTB_HD double dist1d_sample(const double *func, const double *cdf, double func_int, int n, double u,
                            double &pdf_out, int &offset) {
    // std::lower_bound(cdf, cdf + n + 1, u): first index with cdf[i] >= u
    int lo = 0, hi = n + 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] < u)
            lo = mid + 1;
        else
            hi = mid;
    }
    offset = lo - 1 > 0 ? lo - 1 : 0;
    offset = offset < n - 1 ? offset : n - 1;
    double du = u - cdf[offset];
    if (cdf[offset + 1] - cdf[offset] > 0)
        du /= (cdf[offset + 1] - cdf[offset]);
    pdf_out = (func_int > 0) ? func[offset] / func_int : 0;
    return (offset + du) / n;
}

// EnvironmentLight::get_pixel + Le, environmental_light.h:250-311
template <class R> RTB_HD V3<R> env_pixel(const float *tex, int W, int H, int i, int j) {
    if (i < 0)
        i += W;
    if (i >= W)
        i -= W;
    if (j < 0)
        j = 0;
    if (j >= H)
        j = H - 1;
    const float *q = tex + 3 * (size_t(j) * W + i);
    return V3<R>(R(q[0]), R(q[1]), R(q[2]));
}
template <class R> RTB_HD void env_dir_to_uv(const LightT<R> &l, V3<R> unit_dir, R &u, R &v, R &theta) {
    const R pi = Consts<R>::pi();
    if (l.env_probe) {
        const R d = sqrt_(unit_dir.x * unit_dir.x + unit_dir.y * unit_dir.y);
        const R r_coord = (d > 0) ? (R(1) / pi) * acos_(unit_dir.z) / d : R(0);
        u = (unit_dir.x * r_coord + R(1)) * R(0.5);
        v = (unit_dir.y * r_coord + R(1)) * R(0.5);
        v = R(1) - v;
        theta = acos_(unit_dir.z);
    } else {
        theta = acos_(unit_dir.y);
        const R phi = atan2_(-unit_dir.z, unit_dir.x) + pi;
        u = phi / (R(2) * pi);
        v = theta / pi;
    }
}
// (u, v) of the map -> bilinear radiance: the second half of EnvironmentLight::Le (environmental_light.h:250-311)
template <class R> RTB_HD V3<R> env_Le_uv(const ShadeView<R> &S, const LightT<R> &l, R u, R v) {
    const R u_img = u * l.env_w - R(0.5);
    const R v_img = v * l.env_h - R(0.5);
    const int i0 = int(floor_(u_img));
    const int j0 = int(floor_(v_img));
    const R du = u_img - i0;
    const R dv = v_img - j0;
    const float *tex = S.env_texels + l.env_texel_offset;
    const V3<R> c00 = env_pixel<R>(tex, l.env_w, l.env_h, i0, j0);
    const V3<R> c10 = env_pixel<R>(tex, l.env_w, l.env_h, i0 + 1, j0);
    const V3<R> c01 = env_pixel<R>(tex, l.env_w, l.env_h, i0, j0 + 1);
    const V3<R> c11 = env_pixel<R>(tex, l.env_w, l.env_h, i0 + 1, j0 + 1);
    const V3<R> c0 = c00 * (R(1) - du) + c10 * du;
    const V3<R> c1 = c01 * (R(1) - du) + c11 * du;
    return c0 * (R(1) - dv) + c1 * dv;
}
template <class R> RTB_HD V3<R> env_Le(const ShadeView<R> &S, const LightT<R> &l, V3<R> dir) {
    if (l.env_w == 0)
        return V3<R>(1, 1, 1);
    const V3<R> ud = unit_vector(dir);
    R u, v, theta;
    env_dir_to_uv(l, ud, u, v, theta);
    return env_Le_uv(S, l, u, v);
}
// (u, v, theta) of the map -> solid-angle density: the second half of EnvironmentLight::pdf (environmental_light.h:314-356)
template <class R> RTB_HD R env_pdf_uv(const ShadeView<R> &S, const LightT<R> &l, R u, R v, R theta) {
    const R pi = Consts<R>::pi();
    const R sin_theta = sin_(theta);
    if (sin_theta < R(1e-6))
        return 0;
    const int W = l.env_w, H = l.env_h;
    const int u_idx = int(clamp_(R(int(u * W)), R(0), R(W - 1)));
    const int v_idx = int(clamp_(R(int(v * H)), R(0), R(H - 1)));
    const EnvTables T(S.env_tables + l.env_table_offset, W, H);
    // Distribution1D::pdf(i) = func[i] / (func_int * n), environmental_light.h:46-48
    const double ci = T.cond_int[v_idx];
    const double pc = ci > 0 ? T.cond_func[size_t(v_idx) * W + u_idx] / (ci * W) : 0;
    const double pm = T.marg_int > 0 ? ci / (T.marg_int * H) : 0;
    const double map_pdf = pc * pm;
    return R(map_pdf * W * H / (2.0 * double(pi) * double(pi) * double(sin_theta)));
}

// Light::Le(ray)  (light.h:38-40; only EnvironmentLight overrides it)
template <class R> RTB_HD V3<R> light_Le(const ShadeView<R> &S, const LightT<R> &l, V3<R> dir) {
    if (l.type == 4)
        return env_Le(S, l, dir);
    return V3<R>(0, 0, 0);
}

// Light::pdf(origin, direction)  (quad_light.h:51-82, environmental_light.h:314-356, others 0)
template <class R> RTB_HD R light_pdf(const ShadeView<R> &S, const LightT<R> &l, V3<R> origin, V3<R> direction) {
    const R pi = Consts<R>::pi();
    if (l.type == 0) {
        const V3<R> normal(l.normal[0], l.normal[1], l.normal[2]);
        const V3<R> Q(l.Q[0], l.Q[1], l.Q[2]), uu(l.u[0], l.u[1], l.u[2]), vv(l.v[0], l.v[1], l.v[2]);
        const R denom = dot(direction, normal);
        if (denom >= R(-1e-6))
            return 0;
        const R t = dot(Q - origin, normal) / denom;
        if (t < R(0.001) || t > Consts<R>::inf())
            return 0;
        const V3<R> intersection = origin + t * direction;
        const V3<R> ph = intersection - Q;
        const R alpha = dot(ph, uu) / length_squared(uu);
        const R beta = dot(ph, vv) / length_squared(vv);
        if (alpha < 0 || alpha > 1 || beta < 0 || beta > 1)
            return 0;
        const R dist_sq = t * t * length_squared(direction);
        const R cos_theta = -denom / length(direction);
        return dist_sq / (l.area * cos_theta);
    }
    if (l.type == 4) {
        if (l.env_w == 0)
            return R(1) / (R(4) * pi);
        const V3<R> ud = unit_vector(direction);
        R u, v, theta;
        env_dir_to_uv(l, ud, u, v, theta);
        return env_pdf_uv(S, l, u, v, theta);
    }
    return 0;
}

// Light::sample(p, u)  (quad_light.h:18-49, point_light.h:13-26, spot_light.h:15-34,
// directional_light.h:14-22, environmental_light.h:182-248)
template <class R, class G>
RTB_HD LightSampleT<R> light_sample(const ShadeView<R> &S, const LightT<R> &l, V3<R> p, R u0, R u1, G &g) {
    const R pi = Consts<R>::pi();
    LightSampleT<R> s;
    s.Li = V3<R>(0, 0, 0);
    s.wi = V3<R>(0, 0, 0);
    s.to_light = V3<R>(0, 0, 0);
    s.pdf = 0;
    s.dist = 0;
    s.is_delta = false;
    const V3<R> Q(l.Q[0], l.Q[1], l.Q[2]);
    const V3<R> I(l.intensity[0], l.intensity[1], l.intensity[2]);
    switch (l.type) {
    case 0: {
        const V3<R> uu(l.u[0], l.u[1], l.u[2]), vv(l.v[0], l.v[1], l.v[2]);
        const V3<R> normal(l.normal[0], l.normal[1], l.normal[2]);
        const V3<R> light_point = Q + u0 * uu + u1 * vv;
        const V3<R> d = light_point - p;
        const R dist_sq = length_squared(d);
        s.dist = sqrt_(dist_sq);
        s.wi = d / s.dist;
        s.to_light = d;
        const R cos_theta = dot(-s.wi, normal);
        if (cos_theta <= 0)
            return s;
        s.Li = I;
        s.pdf = dist_sq / (l.area * cos_theta);
        return s;
    }
    case 1: {
        const V3<R> d = Q - p;
        const R d2 = length_squared(d);
        s.dist = sqrt_(d2);
        s.wi = d / s.dist;
        s.to_light = d;
        s.Li = I / d2;
        s.pdf = 1;
        s.is_delta = true;
        return s;
    }
    case 2: {
        const V3<R> dir(l.u[0], l.u[1], l.u[2]);
        const V3<R> d = Q - p;
        const R d2 = length_squared(d);
        s.dist = sqrt_(d2);
        s.wi = d / s.dist;
        s.to_light = d;
        s.is_delta = true;
        s.pdf = 1;
        const R cos_theta = dot(-s.wi, dir);
        if (!(cos_theta < l.cos_cutoff))
            s.Li = I / d2;
        return s;
    }
    case 3: {
        const V3<R> dir(l.u[0], l.u[1], l.u[2]);
        s.wi = -dir;
        s.dist = Consts<R>::inf();
        s.Li = I;
        s.is_delta = true;
        s.pdf = 1;
        return s;
    }
    default: {
        s.dist = Consts<R>::inf();
        if (l.env_w == 0) { // environmental_light.h:187-192
            s.wi = random_unit_vector<R>(g);
            s.pdf = R(1) / (R(4) * pi);
            s.Li = V3<R>(1, 1, 1);
            return s;
        }
        const int W = l.env_w, H = l.env_h;
        const EnvTables T(S.env_tables + l.env_table_offset, W, H);
        double pdfs[2];
        int v_idx, u_idx;
        const double v = dist1d_sample(T.cond_int, T.marg_cdf, T.marg_int, H, double(u1), pdfs[1], v_idx);
        const double u = dist1d_sample(T.cond_func + size_t(v_idx) * W, T.cond_cdf + size_t(v_idx) * (W + 1),
                                       T.cond_int[v_idx], W, double(u0), pdfs[0], u_idx);
        const double map_pdf = pdfs[0] * pdfs[1];
        if (map_pdf == 0)
            return s;
        R theta;
        if (l.env_probe) {
            const R uc = R(u) * R(2) - R(1);
            const R vc = (R(1) - R(v)) * R(2) - R(1);
            const R r = sqrt_(uc * uc + vc * vc);
            if (r > R(1))
                return s;
            theta = pi * r;
            const R phi = atan2_(vc, uc);
            const R st = sin_(theta);
            s.wi = V3<R>(st * cos_(phi), st * sin_(phi), cos_(theta));
        } else {
            const R phi = R(u) * R(2) * pi - pi;
            theta = R(v) * pi;
            const R st = sin_(theta);
            const R ct = cos_(theta);
            s.wi = V3<R>(st * cos_(phi), ct, -st * sin_(phi));
        }
        const R sin_theta = sin_(theta);
        if (sin_theta < R(1e-6))
            return s;
        s.pdf = R(map_pdf * W * H / (2.0 * double(pi) * double(pi) * double(sin_theta)));
        s.Li = env_Le(S, l, s.wi);
        return s;
    }
    }
}

// ---- camera ------------------------------------------------------------------------------
// camera::get_ray, camera.h:32-40 (direction NOT normalised)
template <class R, class G>
RTB_HD void camera_ray(const CameraT<R> &c, R s, R t, G &g, V3<R> &o, V3<R> &d, R &time) {
    V3<R> offset(0, 0, 0);
    if (c.lens_radius != 0) {
        const V3<R> rd = c.lens_radius * random_in_unit_disk<R>(g);
        offset = c.u * rd.x + c.v * rd.y;
    }
    o = c.origin + offset;
    d = c.lower_left_corner + s * c.horizontal + t * c.vertical - c.origin - offset;
    time = c.time0 + (c.time1 - c.time0) * g.next();
}

} // namespace rtb

#endif // RTB_SHADING_CUH
