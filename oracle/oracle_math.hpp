// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing in the product path may include, link or call this.
//
// Scalar math shared by the CPU oracle: a deterministic sin/cos (restating the published FDLIBM
// algorithm, which Julia 1.1's Base.sin/cos are a port of: base/special/trig.jl), forward-mode
// dual numbers restating ForwardDiff 0.10.3 (Manifest.toml:164-168; third-party, source not in
// /root/reference — rules from SURVEY.md Appendix B), and the canonical inner-product.
//
// Canonical arithmetic (DESIGN.md §"arithmetic contract"): inner products / matrix products are
// sequential fused-multiply-add chains over the inner index starting from 0.0; everything else is
// plain IEEE double arithmetic in the reference's expression order, with NO contraction
// (compile with -ffp-contract=off).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

static inline double fmad(double a, double b, double c) { return __builtin_fma(a, b, c); }

// ---------------------------------------------------------------------------------------------
// sin / cos: FDLIBM k_sin.c / k_cos.c / e_rem_pio2.c (medium-size branch), (C) Sun Microsystems,
// "Permission to use, copy, modify, and distribute this software is freely granted".
// ---------------------------------------------------------------------------------------------
static inline double k_sin(double x, double y, int iy) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = x * x;
    double v = z * x;
    double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    if (iy == 0) return x + v * (S1 + z * r);
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

static inline double k_cos(double x, double y) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = x * x;
    double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    // FreeBSD/msun form (also what Julia ports): w = 1 - z/2, result w + (((1-w)-z/2) + (z*r - x*y))
    double hz = 0.5 * z;
    double w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + (z * r - x * y));
}

// returns n mod 4 information in *n, reduced argument y0 + y1
static inline void rem_pio2_medium(double x, int* n, double* y0, double* y1) {
    const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00,
                 pio2_1t = 6.07710050650619224932e-11, pio2_2 = 6.07710050630396597660e-11,
                 pio2_2t = 2.02226624879595063154e-21, pio2_3 = 2.02226624871116645580e-21,
                 pio2_3t = 8.47842766036889956997e-32;
    double ax = std::fabs(x);
    double fn = std::nearbyint(ax * invpio2);  // round-to-nearest-even
    double r = ax - fn * pio2_1;
    double w = fn * pio2_1t;
    double y = r - w;
    // extra precision passes, driven by exponent loss as in e_rem_pio2.c
    uint64_t bx, by;
    std::memcpy(&bx, &ax, 8);
    std::memcpy(&by, &y, 8);
    int j = (int)((bx >> 52) & 0x7ff);
    int i = j - (int)((by >> 52) & 0x7ff);
    if (i > 16) {
        double t = r;
        w = fn * pio2_2;
        r = t - w;
        w = fn * pio2_2t - ((t - r) - w);
        y = r - w;
        std::memcpy(&by, &y, 8);
        i = j - (int)((by >> 52) & 0x7ff);
        if (i > 49) {
            t = r;
            w = fn * pio2_3;
            r = t - w;
            w = fn * pio2_3t - ((t - r) - w);
            y = r - w;
        }
    }
    double yl = (r - y) - w;
    long long nn = (long long)fn;
    if (x < 0) {
        *n = (int)((-nn) & 3);
        *y0 = -y;
        *y1 = -yl;
    } else {
        *n = (int)(nn & 3);
        *y0 = y;
        *y1 = yl;
    }
}

static inline void sincos_det(double x, double* s, double* c) {
    double ax = std::fabs(x);
    if (!(ax < 4503599627370496.0)) {  // NaN, Inf or |x| >= 2^52: no meaningful reduction -> NaN
        *s = std::nan("");
        *c = std::nan("");
        return;
    }
    if (ax <= 0.78539816339744827900) {  // pi/4
        if (ax < 7.450580596923828125e-9) {  // 2^-27
            *s = x;
            *c = 1.0;
            return;
        }
        *s = k_sin(x, 0.0, 0);
        *c = k_cos(x, 0.0);
        return;
    }
    int n;
    double y0, y1;
    rem_pio2_medium(x, &n, &y0, &y1);
    double sn = k_sin(y0, y1, 1), cs = k_cos(y0, y1);
    switch (n) {
        case 0: *s = sn; *c = cs; break;
        case 1: *s = cs; *c = -sn; break;
        case 2: *s = -sn; *c = -cs; break;
        default: *s = -cs; *c = sn; break;
    }
}
static inline double sin_det(double x) { double s, c; sincos_det(x, &s, &c); return s; }
static inline double cos_det(double x) { double s, c; sincos_det(x, &s, &c); return c; }

// ---------------------------------------------------------------------------------------------
// Dual numbers, ForwardDiff 0.10.3 rules.
// ---------------------------------------------------------------------------------------------
template <int P>
struct Dual {
    double v;
    double p[P];
    Dual() : v(0.0) { for (int i = 0; i < P; i++) p[i] = 0.0; }
    Dual(double x) : v(x) { for (int i = 0; i < P; i++) p[i] = 0.0; }  // NOLINT (implicit by design)
};

template <int P> inline Dual<P> operator+(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r; r.v = a.v + b.v; for (int i = 0; i < P; i++) r.p[i] = a.p[i] + b.p[i]; return r; }
template <int P> inline Dual<P> operator-(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r; r.v = a.v - b.v; for (int i = 0; i < P; i++) r.p[i] = a.p[i] - b.p[i]; return r; }
template <int P> inline Dual<P> operator-(const Dual<P>& a) {
    Dual<P> r; r.v = -a.v; for (int i = 0; i < P; i++) r.p[i] = -a.p[i]; return r; }
template <int P> inline Dual<P> operator+(const Dual<P>& a, double b) { Dual<P> r = a; r.v = a.v + b; return r; }
template <int P> inline Dual<P> operator+(double b, const Dual<P>& a) { Dual<P> r = a; r.v = b + a.v; return r; }
template <int P> inline Dual<P> operator-(const Dual<P>& a, double b) { Dual<P> r = a; r.v = a.v - b; return r; }
template <int P> inline Dual<P> operator-(double b, const Dual<P>& a) {
    Dual<P> r; r.v = b - a.v; for (int i = 0; i < P; i++) r.p[i] = -a.p[i]; return r; }
// Dual*Dual: value a*b, partial_i = (b.v * a.p_i) + (a.v * b.p_i)
template <int P> inline Dual<P> operator*(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r; r.v = a.v * b.v;
    for (int i = 0; i < P; i++) r.p[i] = (b.v * a.p[i]) + (a.v * b.p[i]);
    return r; }
template <int P> inline Dual<P> operator*(const Dual<P>& a, double b) {
    Dual<P> r; r.v = a.v * b; for (int i = 0; i < P; i++) r.p[i] = a.p[i] * b; return r; }
template <int P> inline Dual<P> operator*(double b, const Dual<P>& a) {
    Dual<P> r; r.v = b * a.v; for (int i = 0; i < P; i++) r.p[i] = a.p[i] * b; return r; }
// Dual/Dual: value a/b, partial_i = (inv(b) * a.p_i) + (-(a/(b*b)) * b.p_i)
template <int P> inline Dual<P> operator/(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r; r.v = a.v / b.v;
    double ib = 1.0 / b.v, cb = -(a.v / (b.v * b.v));
    for (int i = 0; i < P; i++) r.p[i] = (ib * a.p[i]) + (cb * b.p[i]);
    return r; }
template <int P> inline Dual<P> operator/(const Dual<P>& a, double b) {
    Dual<P> r; r.v = a.v / b; for (int i = 0; i < P; i++) r.p[i] = a.p[i] / b; return r; }
// Real/Dual: v = x/b ; partial_i = (-(v/b)) * b.p_i
template <int P> inline Dual<P> operator/(double x, const Dual<P>& b) {
    Dual<P> r; r.v = x / b.v; double c = -(r.v / b.v);
    for (int i = 0; i < P; i++) r.p[i] = c * b.p[i];
    return r; }

inline double value(double x) { return x; }
template <int P> inline double value(const Dual<P>& x) { return x.v; }

inline void sincos_t(double x, double* s, double* c) { sincos_det(x, s, c); }
template <int P> inline void sincos_t(const Dual<P>& x, Dual<P>* s, Dual<P>* c) {
    double sv, cv; sincos_det(x.v, &sv, &cv);
    s->v = sv; c->v = cv;
    double ms = -sv;
    for (int i = 0; i < P; i++) { s->p[i] = cv * x.p[i]; c->p[i] = ms * x.p[i]; }
}
inline double sqrt_t(double x) { return std::sqrt(x); }
template <int P> inline Dual<P> sqrt_t(const Dual<P>& x) {
    Dual<P> r; r.v = std::sqrt(x.v); double d = 1.0 / (2.0 * r.v);
    for (int i = 0; i < P; i++) r.p[i] = d * x.p[i];
    return r; }
inline double inv_t(double x) { return 1.0 / x; }
template <int P> inline Dual<P> inv_t(const Dual<P>& x) {
    Dual<P> r; r.v = 1.0 / x.v; double d = -(r.v * r.v);
    for (int i = 0; i < P; i++) r.p[i] = d * x.p[i];
    return r; }
// abs2 / literal ^2: value x*x, partials (x+x)*p  (== (2x)*p bit-for-bit)
inline double sq_t(double x) { return x * x; }
template <int P> inline Dual<P> sq_t(const Dual<P>& x) {
    Dual<P> r; r.v = x.v * x.v; double d = x.v + x.v;
    for (int i = 0; i < P; i++) r.p[i] = d * x.p[i];
    return r; }

}  // namespace orc
