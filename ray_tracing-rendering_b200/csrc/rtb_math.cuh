// rtb_math.cuh — scalar/vector math shared by every kernel, templated on the
// real type R: float for the production wavefront, double for the fp64
// validation entry points.
//
// The double instantiation follows the reference's operation ORDER exactly
// (src/core/vec3.h), because the validation kernels must reproduce its `t`
// bit for bit: e.g. v / t is (1/t) * v (vec3.h:208-210), unit_vector is
// v / length (vec3.h:222-224), dot is ((x*x + y*y) + z*z) (vec3.h:212-214).
// Those translation units are compiled with -fmad=false so nvcc does not
// contract a*b+c (the x86-64 reference build contains no FMA).
#ifndef RTB_MATH_CUH
#define RTB_MATH_CUH

#include <math.h>
#include <stdint.h>

// The headers are also consumed by plain host C++ (BVH builder, scene upload).
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define RTB_HD __host__ __device__ __forceinline__
#define RTB_D __device__ __forceinline__
#define RTB_HD_OUTLINE __host__ __device__ __noinline__ // one copy per kernel: for large bodies reached from many call sites
#else
#define RTB_HD inline
#define RTB_D inline
#define RTB_HD_OUTLINE inline
#endif

namespace rtb {

template <class R> struct Consts;
template <> struct Consts<float> {
    static RTB_HD float pi() { return 3.14159265358979323846f; }
    static RTB_HD float inf() { return INFINITY; }
};
template <> struct Consts<double> {
    // the reference's literal, src/core/rtweekend.h:18
    static RTB_HD double pi() { return 3.1415926535897932385; }
    static RTB_HD double inf() { return (double)INFINITY; }
};

template <class R> struct V3 {
    R x, y, z;
    RTB_HD V3() {}
    RTB_HD V3(R a, R b, R c) : x(a), y(b), z(c) {}
    RTB_HD R operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    RTB_HD void set(int i, R v) {
        if (i == 0)
            x = v;
        else if (i == 1)
            y = v;
        else
            z = v;
    }
};

template <class R> RTB_HD V3<R> operator+(V3<R> a, V3<R> b) { return V3<R>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class R> RTB_HD V3<R> operator-(V3<R> a, V3<R> b) { return V3<R>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class R> RTB_HD V3<R> operator-(V3<R> a) { return V3<R>(-a.x, -a.y, -a.z); }
template <class R> RTB_HD V3<R> operator*(V3<R> a, V3<R> b) { return V3<R>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <class R> RTB_HD V3<R> operator*(R t, V3<R> a) { return V3<R>(t * a.x, t * a.y, t * a.z); }
template <class R> RTB_HD V3<R> operator*(V3<R> a, R t) { return V3<R>(t * a.x, t * a.y, t * a.z); }
// vec3.h:208-210 : (1/t) * v, NOT a per-component divide
template <class R> RTB_HD V3<R> operator/(V3<R> a, R t) { return (R(1) / t) * a; }
template <class R> RTB_HD R dot(V3<R> a, V3<R> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class R> RTB_HD V3<R> cross(V3<R> u, V3<R> v) {
    return V3<R>(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
template <class R> RTB_HD R length_squared(V3<R> a) { return a.x * a.x + a.y * a.y + a.z * a.z; }

RTB_HD float rsqrt_(float x) {
#ifdef __CUDA_ARCH__
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}
RTB_HD float sqrt_(float x) { return sqrtf(x); }
RTB_HD double sqrt_(double x) { return sqrt(x); }
RTB_HD float fma_(float a, float b, float c) { return fmaf(a, b, c); }
RTB_HD double fma_(double a, double b, double c) { return fma(a, b, c); }
RTB_HD float fabs_(float x) { return fabsf(x); }
RTB_HD double fabs_(double x) { return fabs(x); }
RTB_HD float fmin_(float a, float b) { return fminf(a, b); }
RTB_HD double fmin_(double a, double b) { return fmin(a, b); }
RTB_HD float fmax_(float a, float b) { return fmaxf(a, b); }
RTB_HD double fmax_(double a, double b) { return fmax(a, b); }
RTB_HD float sin_(float x) { return sinf(x); }
RTB_HD double sin_(double x) { return sin(x); }
RTB_HD float cos_(float x) { return cosf(x); }
RTB_HD double cos_(double x) { return cos(x); }
// fp32 acos / atan2 (sphere u,v and every direction -> environment-map look-up): polynomial forms, a third of the
// instructions of acosf / atan2f, which were 9 % of the general fused kernel on C4-env.  acos(x) = sqrt(1 - |x|) P6(|x|)
// (Chebyshev fit on [0, 1]); atan(a) = a Q6(a^2) on [0, 1] with the octant folded in.  Absolute error against the exact
// functions of the same fp32 argument: acos <= 5e-7 rad, atan2 <= 8e-7 rad (tests/test_fast_invtrig.py, 2 x 10^5 arguments
// each): 3e-4 of an environment-map texel at 2048 x 1024.  The fp64 validation
// instantiations call the C library.
#ifndef RTB_FAST_INVTRIG
#define RTB_FAST_INVTRIG 1
#endif
RTB_HD float acos_(float x) {
#if RTB_FAST_INVTRIG
    const float ax = fminf(fabsf(x), 1.0f);
    float p = 0.00225136825f;
    p = fmaf(p, ax, -0.0110123861f);
    p = fmaf(p, ax, 0.0267493312f);
    p = fmaf(p, ax, -0.0487244023f);
    p = fmaf(p, ax, 0.0887373288f);
    p = fmaf(p, ax, -0.214583696f);
    p = fmaf(p, ax, 1.57079615f);
    const float r = sqrtf(1.0f - ax) * p;
    return x < 0.0f ? 3.14159265358979f - r : r;
#else
    return acosf(x);
#endif
}
RTB_HD double acos_(double x) { return acos(x); }
RTB_HD float atan2_(float y, float x) {
#if RTB_FAST_INVTRIG
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = mx > 0.0f ? mn / mx : 0.0f;
    const float s = a * a;
    float q = 0.00782548295f;
    q = fmaf(q, s, -0.0368986292f);
    q = fmaf(q, s, 0.0837415565f);
    q = fmaf(q, s, -0.134804056f);
    q = fmaf(q, s, 0.198798722f);
    q = fmaf(q, s, -0.333263745f);
    q = fmaf(q, s, 0.999999328f);
    float r = a * q;
    r = ay > ax ? 1.57079632679490f - r : r;
    r = x < 0.0f ? 3.14159265358979f - r : r;
    return y < 0.0f ? -r : r;
#else
    return atan2f(y, x);
#endif
}
RTB_HD double atan2_(double y, double x) { return atan2(y, x); }
RTB_HD float log_(float x) { return logf(x); }
RTB_HD double log_(double x) { return log(x); }
RTB_HD float floor_(float x) { return floorf(x); }
RTB_HD double floor_(double x) { return floor(x); }
RTB_HD float cbrt_(float x) { return cbrtf(x); }
RTB_HD double cbrt_(double x) { return cbrt(x); }
RTB_HD void sincos_(float x, float *s, float *c) {
#ifdef __CUDA_ARCH__
    sincosf(x, s, c);
#else
    *s = sinf(x);
    *c = cosf(x);
#endif
}
RTB_HD void sincos_(double x, double *s, double *c) {
    *s = sin(x);
    *c = cos(x);
}
// pow(x, 5) as the reference writes it (material.h:203, :435); the production
// float path multiplies it out.
RTB_HD float pow5_(float x) {
    const float x2 = x * x;
    return x2 * x2 * x;
}
RTB_HD double pow5_(double x) { return pow(x, 5.0); }

template <class R> RTB_HD R length(V3<R> a) { return sqrt_(length_squared(a)); }
// vec3.h:222-224
RTB_HD V3<double> unit_vector(V3<double> v) { return v / length(v); }
RTB_HD V3<float> unit_vector(V3<float> v) { return rsqrt_(length_squared(v)) * v; }

// rtweekend.h:40-46
template <class R> RTB_HD R clamp_(R x, R lo, R hi) { return x < lo ? lo : (x > hi ? hi : x); }

// vec3.h:239-241
template <class R> RTB_HD V3<R> reflect(V3<R> v, V3<R> n) { return v - (R(2) * dot(v, n)) * n; }
// vec3.h:243-248
template <class R> RTB_HD V3<R> refract(V3<R> uv, V3<R> n, R etai_over_etat) {
    const R cos_theta = fmin_(dot(-uv, n), R(1));
    const V3<R> r_out_perp = etai_over_etat * (uv + cos_theta * n);
    const V3<R> r_out_parallel = (-sqrt_(fabs_(R(1) - length_squared(r_out_perp)))) * n;
    return r_out_perp + r_out_parallel;
}
// vec3.h:80-84
template <class R> RTB_HD bool near_zero(V3<R> v) {
    const R s = R(1e-8);
    return fabs_(v.x) < s && fabs_(v.y) < s && fabs_(v.z) < s;
}
template <class R> RTB_HD R max3(V3<R> v) { return fmax_(v.x, fmax_(v.y, v.z)); }

// onb.h:29-34
template <class R> struct Onb {
    V3<R> u, v, w;
    RTB_HD void build_from_w(V3<R> n) {
        w = unit_vector(n);
        const V3<R> a = (fabs_(w.x) > R(0.9)) ? V3<R>(0, 1, 0) : V3<R>(1, 0, 0);
        v = unit_vector(cross(w, a));
        u = cross(w, v);
    }
    RTB_HD V3<R> local(V3<R> a) const { return a.x * u + a.y * v + a.z * w; }
};

} // namespace rtb

#endif // RTB_MATH_CUH
