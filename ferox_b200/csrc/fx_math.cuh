// fx_math.cuh -- scalar/vector arithmetic shared by host set-up code and the sm_100a kernels.
//
// Parity contract: every expression here keeps the operand order and the rounding points of the
// reference's inline math (reference include/ferox.h:597-686).  Device code is compiled with
// -fmad=false and host code with -ffp-contract=off, so no multiply-add is ever fused (the reference
// binary has no FMA, SURVEY.md finding 5); division and sqrt are IEEE-rounded on both sides.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define FX_HD __host__ __device__ __forceinline__
#define FX_D __device__ __forceinline__
#else
#define FX_HD inline
#define FX_D inline
#endif

struct V2 {
    float x, y;
};

FX_HD V2 v2(float x, float y) { V2 r; r.x = x; r.y = y; return r; }
FX_HD V2 vadd(V2 a, V2 b) { return v2(a.x + b.x, a.y + b.y); }
FX_HD V2 vsub(V2 a, V2 b) { return v2(a.x - b.x, a.y - b.y); }
FX_HD V2 vneg(V2 a) { return v2(-a.x, -a.y); }
FX_HD V2 vscale(V2 a, float k) { return v2(a.x * k, a.y * k); }
FX_HD float vdot(V2 a, V2 b) { return (a.x * b.x) + (a.y * b.y); }
FX_HD float vcross(V2 a, V2 b) { return (a.x * b.y) - (a.y * b.x); }
FX_HD float vlen2(V2 a) { return (a.x * a.x) + (a.y * a.y); }

// frVector2Normalize (ferox.h:648-653): scale by the reciprocal, zero vectors pass through
FX_HD V2 vunit(V2 a) {
    float m = sqrtf(vlen2(a));
    return (m > 0.0f) ? vscale(a, 1.0f / m) : a;
}
// frVector2LeftNormal / RightNormal (ferox.h:656-663): NORMALISED perpendiculars
FX_HD V2 vperp_left(V2 a) { return vunit(v2(-a.y, a.x)); }
FX_HD V2 vperp_right(V2 a) { return vunit(v2(a.y, -a.x)); }

// A body transform as the kernels see it: position + cached sin/cos of the (normalised) angle.
struct Xf {
    V2 p;
    float sn, cs;
};

// frVector2RotateTx / frVector2Transform (ferox.h:675-686)
FX_HD V2 xf_rotate(V2 v, Xf t) { return v2(v.x * t.cs - v.y * t.sn, v.x * t.sn + v.y * t.cs); }
FX_HD V2 xf_apply(V2 v, Xf t) {
    return v2(t.p.x + (v.x * t.cs - v.y * t.sn), t.p.y + (v.x * t.sn + v.y * t.cs));
}
// frVector2Rotate(v, -tx.angle) (ferox.h:666-672 as called from src/collision.c:328-330,646,715).
// The reference evaluates sinf(-a) and cosf(-a); glibc's sinf is odd and cosf even bit-for-bit on
// the stored angles (checked exhaustively in tests/test_sincos.py), and the cached sin_/cos_ are
// sinf(a)/cosf(a) (src/rigid_body.c:234-235), so with S = -sn, C = cs:
//   x' = v.x*C - v.y*S = (v.x*cs) + (v.y*sn)      y' = v.x*S + v.y*C = (v.y*cs) - (v.x*sn)
// which are the same IEEE operations (negation is exact and commutes with * and +/-).
FX_HD V2 xf_unrotate(V2 v, Xf t) { return v2((v.x * t.cs) + (v.y * t.sn), (v.y * t.cs) - (v.x * t.sn)); }

// fminf/fmaxf exactly as the COMPILED reference evaluates them: glibc 2.39 x86-64 is built around
// minss/maxss, so for ordered operands the SECOND argument wins ties (signed zeros) and a NaN
// operand yields the other one.  Call sites pass the operands in the order gcc 13 -O2 emitted for
// the reference (see oracle/step_oracle.c, pin_fminf).
FX_HD float pin_fminf(float x, float y) {
    if (y != y) return x;
    if (x != x) return y;
    return (x < y) ? x : y;
}
FX_HD float pin_fmaxf(float x, float y) {
    if (y != y) return x;
    if (x != x) return y;
    return (x > y) ? x : y;
}

// C float->int conversion (truncation toward zero) used for cell coordinates,
// reference src/broad_phase.c:104-108.
FX_HD int cell_coord(float v, float inv_cell) {
#if defined(__CUDA_ARCH__)
    return __float2int_rz(v * inv_cell);
#else
    return (int) (v * inv_cell);
#endif
}

// ---------------------------------------------------------------------------------------------
// sinf/cosf of glibc 2.39 (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h; the x86-64 build
// selects the FMA ifunc variant, so every a*b+c below is ONE fused operation), restated for the
// argument range the position integrator produces: |a| in [0.75, 120) uses the fast reduction,
// 2^-12 <= |a| < 0.75 the direct polynomial.  Angles are normalised to [pi, 3*pi) before this is called
// (src/rigid_body.c:228, 589-591).  Arguments outside [2^-12, 120) fall back to the CUDA library
// (never reached from the integrator).  tests/test_sincos.py compares every float in [pi, 3*pi]
// (and a wider sweep) against the host libm, bit for bit.
// ---------------------------------------------------------------------------------------------
struct SinCos {
    float s, c;
};

FX_HD float fx_sincos_poly(double x, double x2, bool negate_cos, int n) {
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5,
                 C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double s1 = fma(x2, S3, S2);
        double x7 = x3 * x2;
        double s = fma(x3, S1, x);
        return (float) fma(x7, s1, s);
    } else {
        double k = negate_cos ? -1.0 : 1.0;
        double x4 = x2 * x2;
        double c2 = fma(x2, k * C4, k * C3);
        double c1 = fma(x2, k * C1, k * C0);
        double x6 = x4 * x2;
        double c = fma(x4, k * C2, c1);
        return (float) fma(x6, c2, c);
    }
}

FX_HD SinCos fx_sincosf(float a) {
    SinCos r;
    float aa = fabsf(a);
    if (!(aa >= 0x1p-12f && aa < 120.0f)) {
        r.s = sinf(a);
        r.c = cosf(a);
        return r;
    }
    double x = (double) a;
    if (aa < 0.75f) {  // glibc tests abstop12(y) < abstop12(pi/4): top-12-bit granularity, i.e. |y| < 0.75
        double x2 = x * x;
        r.s = fx_sincos_poly(x, x2, false, 0);
        r.c = fx_sincos_poly(x, x2, false, 1);
        return r;
    }
    const double HPI_INV = 0x1.45F306DC9C883p+23, HPI = 0x1.921FB54442D18p0;
    double q = x * HPI_INV;
    int n = ((int32_t) q + 0x800000) >> 24;
    double y = fma(-(double) n, HPI, x);
    double y2 = y * y;
    double ss = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    double sc = (((n + 1) & 3) == 1 || ((n + 1) & 3) == 2) ? -1.0 : 1.0;
    r.s = fx_sincos_poly(y * ss, y2, (n & 2) != 0, n);
    r.c = fx_sincos_poly(y * sc, y2, ((n + 1) & 2) != 0, n ^ 1);
    return r;
}

// frNormalizeAngle (src/rigid_body.c:589-591): the sum and the product are double expressions
// (M_PI is a double constant), floorf first rounds its argument to float.
FX_HD float fx_normalize_angle(float angle) {
    const float TWO_PI = (float) (2.0 * 3.14159265358979323846);            // rigid_body.c:54
    const float INV_TWO_PI = (float) (1.0 / (2.0 * 3.14159265358979323846));
    float t = (float) (((double) angle + -3.14159265358979323846) * (double) INV_TWO_PI);
    return angle - (TWO_PI * floorf(t));
}
