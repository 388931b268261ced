// sphere_math.h -- pinned scalar math shared by the sm_100a kernels and the host code.
//
// Every function here is a fixed sequence of IEEE-754 correctly rounded operations
// (+ - * / sqrt fma, float or double) and nothing else, so the SAME bits come out on
// an x86 host and on the GPU.  Rules that make that true:
//   * device translation units are compiled with  --fmad=false  (nvcc never contracts
//     a*b+c on its own; --prec-div / --prec-sqrt keep their IEEE defaults),
//   * host translation units are compiled with    -ffp-contract=off -mfma
//     (gcc never contracts on its own; fmaf()/fma() become one vfmadd instruction),
//   * a fused multiply-add is used ONLY where fmaf()/fma() is written out below.
//
// These functions stand in for the libm calls of the reference's spherical path
//   asin / atan2 / round            include/RegisterPhotoICP.h:2676-2680 (== 2978-2981)
//   sin / cos (LUT tables)          include/RegisterPhotoICP.h:4558-4569
//   double sin / cos in the SE(3) pseudo-exponential   include/RegisterPhotoICP.h:4697
// so that index maps (r', c') are bit-exact between CPU and GPU.  Accuracy of each
// replacement against glibc is checked in tests/test_sphere_math.py (<= 2 ulp).
#pragma once
#include <math.h>
#include <stdint.h>
#include <limits.h>

#if defined(__CUDACC__)
#define R360_HD __host__ __device__ __forceinline__
#else
#define R360_HD inline
#endif

// The reference's PI macro is a truncated double literal (include/Miscellaneous.h:43-45).
#define R360_PI_D 3.14159265359

// ---------------------------------------------------------------- float sin / cos
// Cody-Waite reduction by pi/2 (three-part constant) + degree 7/8 polynomials.
R360_HD void r360_sincosf(float x, float* s_out, float* c_out) {
    float kf = rintf(x * 0.636619772367581343f);   // x * 2/pi
    int k = (int)kf;
    float r = fmaf(-kf, 1.5703125f, x);
    r = fmaf(-kf, 4.837512969970703125e-4f, r);
    r = fmaf(-kf, 7.54978995489188216e-8f, r);
    float z = r * r;
    // sin(r) on [-pi/4, pi/4]
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(z, ps, -1.6666654611e-1f);
    float sr = fmaf(r * z, ps, r);
    // cos(r) on [-pi/4, pi/4]
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(z, pc, 4.166664568298827e-2f);
    float cr = fmaf(z * z, pc, fmaf(z, -0.5f, 1.0f));
    float s, c;
    switch (k & 3) {
        case 0:  s = sr;  c = cr;  break;
        case 1:  s = cr;  c = -sr; break;
        case 2:  s = -sr; c = -cr; break;
        default: s = -cr; c = sr;  break;
    }
    *s_out = s;
    *c_out = c;
}
R360_HD float r360_sinf(float x) { float s, c; r360_sincosf(x, &s, &c); return s; }
R360_HD float r360_cosf(float x) { float s, c; r360_sincosf(x, &s, &c); return c; }

// ---------------------------------------------------------------- float asin
// |x| <= 0.5 : x + x*z*P(z), z = x^2 (degree-4 minimax, 1.6e-8 rel)
// |x| >  0.5 : pi/2 - 2*asin(sqrt((1-|x|)/2)) with a hi/lo split of pi/2
R360_HD float r360_asin_poly(float z) {
    float p = fmaf(z, 3.8206567683e-02f, 2.6494211752e-02f);
    p = fmaf(z, p, 4.5010712250e-02f);
    p = fmaf(z, p, 7.4988090911e-02f);
    p = fmaf(z, p, 1.6666672766e-01f);
    return p;
}
R360_HD float r360_asinf(float x) {
    float ax = fabsf(x);
    if (ax <= 0.5f) {
        float z = x * x;
        return fmaf(x * z, r360_asin_poly(z), x);
    }
    if (!(ax <= 1.0f)) return NAN;                   // |x| > 1 or NaN, as libm
    float z = (1.0f - ax) * 0.5f;
    float s = sqrtf(z);
    float t = fmaf(s * z, r360_asin_poly(z), s);     // asin(s)
    // pi/2 = 1.57079637050628662109375f - 4.37113900018624283e-8
    float r = fmaf(-2.0f, t, 1.57079637050628662109375f) - 4.37113900018624283e-8f;
    return x < 0.0f ? -r : r;
}

// ---------------------------------------------------------------- float atan2
// q = min/max in [0,1]; atan(q) = q + q*t*P(t), t = q^2 (degree-8 minimax, 1.3e-8 rel);
// octant fix-ups with hi/lo constants.
R360_HD float r360_atan2f(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float q = (mx == 0.0f) ? 0.0f : mn / mx;
    float t = q * q;
    float p = fmaf(t, -2.4470558835e-03f, 1.3750389111e-02f);
    p = fmaf(t, p, -3.6270357867e-02f);
    p = fmaf(t, p, 6.2843779659e-02f);
    p = fmaf(t, p, -8.6731798886e-02f);
    p = fmaf(t, p, 1.1037996988e-01f);
    p = fmaf(t, p, -1.4279111346e-01f);
    p = fmaf(t, p, 1.9999766029e-01f);
    p = fmaf(t, p, -3.3333331951e-01f);
    float a = fmaf(q * t, p, q);                      // atan(q) in [0, pi/4]
    if (ay > ax) a = (1.57079637050628662109375f - a) - 4.37113900018624283e-8f;
    if (x < 0.0f) a = (3.1415927410125732421875f - a) - 8.74227800037248566e-8f;
    return copysignf(a, y);                           // atan2(-0, x<0) = -pi, as libm
}

// ---------------------------------------------------------------- round to int
// (int)round(v) of the reference (RPI.h:2679-2680).  x86 cvttss2si returns INT_MIN
// for NaN / out-of-range; the GPU would saturate, so the case is made explicit.
R360_HD int r360_round_to_int(float v) {
    float r = roundf(v);
    if (!(r > -2147483648.0f && r < 2147483648.0f)) return INT_MIN;
    return (int)r;
}

// ---------------------------------------------------------------- double sin / cos
// fdlibm-style kernels on [-pi/4, pi/4] after a three-part Cody-Waite reduction
// (good to |x| ~ 1e5, far beyond any Gauss-Newton rotation update).
R360_HD void r360_sincos(double x, double* s_out, double* c_out) {
    double kf = rint(x * 6.36619772367581382433e-01);
    long long k = (long long)kf;
    double r = fma(-kf, 1.57079632673412561417e+00, x);
    r = fma(-kf, 6.07710050630396597660e-11, r);
    r = fma(-kf, 2.02226624879595063154e-21, r);
    double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    double sr = fma(r * z, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
    double s, c;
    switch ((int)(k & 3)) {
        case 0:  s = sr;  c = cr;  break;
        case 1:  s = cr;  c = -sr; break;
        case 2:  s = -sr; c = -cr; break;
        default: s = -cr; c = sr;  break;
    }
    *s_out = s;
    *c_out = c;
}

// ---------------------------------------------------------------- Huber weight
// weightHuber<float>, include/RegisterPhotoICP.h:544-554 (exact float form).
R360_HD float r360_huber(float e, float k) {
    float a = fabsf(e);
    if (a < k) return 1.0f;
    return sqrtf(2 * k * a - k * k) / a;
}
