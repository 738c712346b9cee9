// Device-side scalar math for the B200 engine: deterministic sin/cos, forward-mode dual numbers
// with ForwardDiff-0.10 partial rules, the reference's continuous dynamics models
// (dynamics/*.jl) and explicit integrators (src/integration.jl).
//
// Arithmetic contract: this translation unit is compiled with -fmad=false, so every
// `a*b+c` written here is a separate multiply and add exactly as Julia evaluates it; fused
// multiply-adds appear only where fma() is spelled out (the inner-product chains).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace tob {

#define TOB_DEV __device__ __forceinline__

// ------------------------------------------------------------------------------------------
// sin/cos: the FDLIBM kernels (k_sin/k_cos/e_rem_pio2 medium branch), < 1 ulp, and — unlike
// CUDA's sin()/cos() — bitwise reproducible against the host restatement.
// ------------------------------------------------------------------------------------------
TOB_DEV double ksin(double x, double y, int iy) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = x * x;
    double v = z * x;
    double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    if (iy == 0) return x + v * (S1 + z * r);
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
TOB_DEV double kcos(double x, double y) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = x * x;
    double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    double hz = 0.5 * z;
    double w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + (z * r - x * y));
}
TOB_DEV void sincos_det(double x, double* s, double* c) {
    double ax = fabs(x);
    if (!(ax < 4503599627370496.0)) {
        *s = __longlong_as_double(0x7ff8000000000000LL);
        *c = *s;
        return;
    }
    if (ax <= 0.78539816339744827900) {
        if (ax < 7.450580596923828125e-9) {
            *s = x;
            *c = 1.0;
            return;
        }
        *s = ksin(x, 0.0, 0);
        *c = kcos(x, 0.0);
        return;
    }
    const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00,
                 pio2_1t = 6.07710050650619224932e-11, pio2_2 = 6.07710050630396597660e-11,
                 pio2_2t = 2.02226624879595063154e-21, pio2_3 = 2.02226624871116645580e-21,
                 pio2_3t = 8.47842766036889956997e-32;
    double fn = rint(ax * invpio2);
    double r = ax - fn * pio2_1;
    double w = fn * pio2_1t;
    double y = r - w;
    int j = (int)((__double_as_longlong(ax) >> 52) & 0x7ff);
    int i = j - (int)((__double_as_longlong(y) >> 52) & 0x7ff);
    if (i > 16) {
        double t = r;
        w = fn * pio2_2;
        r = t - w;
        w = fn * pio2_2t - ((t - r) - w);
        y = r - w;
        i = j - (int)((__double_as_longlong(y) >> 52) & 0x7ff);
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
    int n;
    if (x < 0) {
        n = (int)((-nn) & 3);
        y = -y;
        yl = -yl;
    } else {
        n = (int)(nn & 3);
    }
    double sn = ksin(y, yl, 1), cs = kcos(y, yl);
    switch (n) {
        case 0: *s = sn; *c = cs; break;
        case 1: *s = cs; *c = -sn; break;
        case 2: *s = -sn; *c = -cs; break;
        default: *s = -cs; *c = sn; break;
    }
}

// ------------------------------------------------------------------------------------------
// dual numbers: value + P partials.  Rules follow ForwardDiff 0.10 (dual.jl / partials.jl).
// ------------------------------------------------------------------------------------------
template <int P>
struct Dual {
    double v;
    double p[P];
    TOB_DEV Dual() {}
    TOB_DEV Dual(double x) : v(x) {
#pragma unroll
        for (int i = 0; i < P; i++) p[i] = 0.0;
    }
};
#define TOB_FORP _Pragma("unroll") for (int i = 0; i < P; i++)

template <int P> TOB_DEV Dual<P> operator+(const Dual<P>& a, const Dual<P>& b) { Dual<P> r; r.v = a.v + b.v; TOB_FORP r.p[i] = a.p[i] + b.p[i]; return r; }
template <int P> TOB_DEV Dual<P> operator-(const Dual<P>& a, const Dual<P>& b) { Dual<P> r; r.v = a.v - b.v; TOB_FORP r.p[i] = a.p[i] - b.p[i]; return r; }
template <int P> TOB_DEV Dual<P> operator-(const Dual<P>& a) { Dual<P> r; r.v = -a.v; TOB_FORP r.p[i] = -a.p[i]; return r; }
template <int P> TOB_DEV Dual<P> operator+(const Dual<P>& a, double b) { Dual<P> r = a; r.v = a.v + b; return r; }
template <int P> TOB_DEV Dual<P> operator+(double b, const Dual<P>& a) { Dual<P> r = a; r.v = b + a.v; return r; }
template <int P> TOB_DEV Dual<P> operator-(const Dual<P>& a, double b) { Dual<P> r = a; r.v = a.v - b; return r; }
template <int P> TOB_DEV Dual<P> operator-(double b, const Dual<P>& a) { Dual<P> r; r.v = b - a.v; TOB_FORP r.p[i] = -a.p[i]; return r; }
template <int P> TOB_DEV Dual<P> operator*(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r; r.v = a.v * b.v; TOB_FORP r.p[i] = (b.v * a.p[i]) + (a.v * b.p[i]); return r; }
template <int P> TOB_DEV Dual<P> operator*(const Dual<P>& a, double b) { Dual<P> r; r.v = a.v * b; TOB_FORP r.p[i] = a.p[i] * b; return r; }
template <int P> TOB_DEV Dual<P> operator*(double b, const Dual<P>& a) { Dual<P> r; r.v = b * a.v; TOB_FORP r.p[i] = a.p[i] * b; return r; }
template <int P> TOB_DEV Dual<P> operator/(const Dual<P>& a, const Dual<P>& b) {
    Dual<P> r; r.v = a.v / b.v; double ib = 1.0 / b.v, cb = -(a.v / (b.v * b.v));
    TOB_FORP r.p[i] = (ib * a.p[i]) + (cb * b.p[i]); return r; }
template <int P> TOB_DEV Dual<P> operator/(const Dual<P>& a, double b) { Dual<P> r; r.v = a.v / b; TOB_FORP r.p[i] = a.p[i] / b; return r; }
template <int P> TOB_DEV Dual<P> operator/(double x, const Dual<P>& b) {
    Dual<P> r; r.v = x / b.v; double c = -(r.v / b.v); TOB_FORP r.p[i] = c * b.p[i]; return r; }

TOB_DEV double valof(double x) { return x; }
template <int P> TOB_DEV double valof(const Dual<P>& x) { return x.v; }
TOB_DEV void sincos_t(double x, double* s, double* c) { sincos_det(x, s, c); }
template <int P> TOB_DEV void sincos_t(const Dual<P>& x, Dual<P>* s, Dual<P>* c) {
    double sv, cv; sincos_det(x.v, &sv, &cv);
    s->v = sv; c->v = cv; double ms = -sv;
    TOB_FORP { s->p[i] = cv * x.p[i]; c->p[i] = ms * x.p[i]; }
}
TOB_DEV double sqrt_t(double x) { return sqrt(x); }
template <int P> TOB_DEV Dual<P> sqrt_t(const Dual<P>& x) { Dual<P> r; r.v = sqrt(x.v); double d = 1.0 / (2.0 * r.v); TOB_FORP r.p[i] = d * x.p[i]; return r; }
TOB_DEV double inv_t(double x) { return 1.0 / x; }
template <int P> TOB_DEV Dual<P> inv_t(const Dual<P>& x) { Dual<P> r; r.v = 1.0 / x.v; double d = -(r.v * r.v); TOB_FORP r.p[i] = d * x.p[i]; return r; }
TOB_DEV double sq_t(double x) { return x * x; }
template <int P> TOB_DEV Dual<P> sq_t(const Dual<P>& x) { Dual<P> r; r.v = x.v * x.v; double d = x.v + x.v; TOB_FORP r.p[i] = d * x.p[i]; return r; }
template <class T> TOB_DEV T make_t(double x) { return T(x); }

// x / 6.0, correctly rounded, without the division subroutine: q = x*RN(1/6); r = fma(-6,q,x) (exact);
// q' = fma(r, RN(1/6), q) is the IEEE quotient (Markstein's correction step; checked against x/6.0 on 3.2e9
// random and structured doubles, tools/check_div6.c).  Outside the safe exponent range (and for zeros, whose
// sign the correction step loses) the true division is used, so the result is ALWAYS bitwise x/6.0.
TOB_DEV double div6(double x) {
    const unsigned e = ((unsigned)__double2hiint(x) >> 20) & 0x7ffu;
    if (e - 64u < 1920u) {
        const double y = 1.0 / 6.0;
        const double q = x * y;
        const double r = fma(-6.0, q, x);
        return fma(r, y, q);
    }
    if (x == 0.0) return x;  // +-0 / 6 = +-0: most dual partials are exact zeros, keep them off the division subroutine
    return x / 6.0;
}
TOB_DEV double div6_t(double x) { return div6(x); }
template <int P> TOB_DEV Dual<P> div6_t(const Dual<P>& a) { Dual<P> r; r.v = div6(a.v); TOB_FORP r.p[i] = div6(a.p[i]); return r; }

// ------------------------------------------------------------------------------------------
// continuous dynamics  (model ids as in include/trajopt_b200.h)
// ------------------------------------------------------------------------------------------
template <int MODEL> struct ModelDims;
template <> struct ModelDims<0> { static constexpr int n = 2, m = 1; };
template <> struct ModelDims<1> { static constexpr int n = 2, m = 1; };
template <> struct ModelDims<2> { static constexpr int n = 3, m = 2; };
template <> struct ModelDims<3> { static constexpr int n = 4, m = 1; };
template <> struct ModelDims<4> { static constexpr int n = 13, m = 4; };
template <> struct ModelDims<5> { static constexpr int n = 4, m = 1; };
template <> struct ModelDims<6> { static constexpr int n = 4, m = 2; };

template <class T> TOB_DEV void quat_mul(T* out, const T* q2, const T* q1) {  // dynamics/quaternions.jl:23-27
    const T s1 = q1[0], s2 = q2[0];
    out[0] = s1 * s2 - ((q1[1] * q2[1] + q1[2] * q2[2]) + q1[3] * q2[3]);
    T cx = q2[2] * q1[3] - q2[3] * q1[2];
    T cy = q2[3] * q1[1] - q2[1] * q1[3];
    T cz = q2[1] * q1[2] - q2[2] * q1[1];
    out[1] = (s1 * q2[1] + s2 * q1[1]) + cx;
    out[2] = (s1 * q2[2] + s2 * q1[2]) + cy;
    out[3] = (s1 * q2[3] + s2 * q1[3]) + cz;
}

template <class T> TOB_DEV void f_dp_tau(T* xd, const T* x, const T& t1, const T& t2) {
    // URDF double pendulum (dynamics/urdf/doublependulum.urdf:15-63) in closed form; no joint damping
    const double m1 = 1.0, m2 = 1.0, l1 = 1.0, lc1 = 0.5, lc2 = 1.0, I1 = 0.083, I2 = 0.33, g = 9.81;
    T s1, c1, s2, c2, s12, c12;
    sincos_t(x[0], &s1, &c1);
    sincos_t(x[1], &s2, &c2);
    sincos_t(x[0] + x[1], &s12, &c12);
    const T qd1 = x[2], qd2 = x[3];
    T M11 = (((I1 + I2) + m1 * (lc1 * lc1)) + m2 * ((l1 * l1 + lc2 * lc2) + ((2.0 * l1) * lc2) * c2));
    T M12 = I2 + m2 * (lc2 * lc2 + (l1 * lc2) * c2);
    const double M22 = I2 + m2 * (lc2 * lc2);
    T h = ((m2 * l1) * lc2) * s2;
    T b1 = ((-h) * (((2.0 * qd1) * qd2) + qd2 * qd2) + ((m1 * lc1 + m2 * l1) * g) * s1) + ((m2 * lc2) * g) * s12;
    T b2 = (h * (qd1 * qd1)) + ((m2 * lc2) * g) * s12;
    T r1 = t1 - b1, r2 = t2 - b2;
    T det = M11 * M22 - M12 * M12;
    xd[0] = qd1;
    xd[1] = qd2;
    xd[2] = (M22 * r1 - M12 * r2) / det;
    xd[3] = (M11 * r2 - M12 * r1) / det;
}

template <int MODEL, class T> TOB_DEV void f_model(T* xd, const T* x, const T* u) {
    if constexpr (MODEL == 0) {  // dynamics/double_integrator.jl:1-4
        xd[0] = x[1];
        xd[1] = u[0];
    } else if constexpr (MODEL == 1) {  // dynamics/pendulum.jl:3-12
        const double m = 1.0, b = 0.1, lc = 0.5, I = 0.25, g = 9.81;
        T s, c;
        sincos_t(x[0], &s, &c);
        xd[0] = x[1];
        xd[1] = ((u[0] - ((m * g) * lc) * s) - b * x[1]) / I;
    } else if constexpr (MODEL == 2) {  // dynamics/car.jl:3-8
        T s, c;
        sincos_t(x[2], &s, &c);
        xd[0] = u[0] * c;
        xd[1] = u[0] * s;
        xd[2] = u[1];
    } else if constexpr (MODEL == 3) {  // dynamics/cartpole.jl:9-36 : qdd = (-H)\(C*qd + G - B*u), pivoted 2x2 LU
        const double mc = 1.0, mp = 0.2, l = 0.5, g = 9.81;
        T s, c;
        if (isfinite(valof(x[1]))) {
            sincos_t(x[1], &s, &c);
        } else {
            s = T(__longlong_as_double(0x7ff0000000000000LL));
            c = s;
        }
        const T qd0 = x[2], qd1 = x[3];
        T a00 = -T(mc + mp), a01 = -((mp * l) * c), a10 = -((mp * l) * c), a11 = -T(mp * (l * l));
        T c12 = (((-mp) * qd1) * l) * s;
        T r0 = ((T(0.0) * qd0 + c12 * qd1) + T(0.0)) - 1.0 * u[0];
        T r1 = ((T(0.0) * qd0 + T(0.0) * qd1) + ((mp * g) * l) * s) - 0.0 * u[0];
        if (fabs(valof(a10)) > fabs(valof(a00))) {
            T t = a00; a00 = a10; a10 = t;
            t = a01; a01 = a11; a11 = t;
            t = r0; r0 = r1; r1 = t;
        }
        T inv00 = inv_t(a00);
        a10 = a10 * inv00;
        a11 = a11 - a10 * a01;
        r1 = r1 - a10 * r0;
        T q1 = r1 / a11;
        r0 = r0 - a01 * q1;
        T q0 = r0 / a00;
        xd[0] = qd0;
        xd[1] = qd1;
        xd[2] = q0;
        xd[3] = q1;
    } else if constexpr (MODEL == 4) {  // dynamics/quadrotor.jl:10-71
        const double mass = 0.5, Jx = 0.0023, Jy = 0.0023, Jz = 0.004, L = 0.1750, kf = 1.0, km = 0.0245;
        const double Jinvx = 1.0 / 0.0023, Jinvy = 1.0 / 0.0023, Jinvz = 1.0 / 0.004;
        T nrm = sqrt_t(((sq_t(x[3]) + sq_t(x[4])) + sq_t(x[5])) + sq_t(x[6]));
        T in = inv_t(nrm);
        T q[4] = {in * x[3], in * x[4], in * x[5], in * x[6]};
        const T* v = x + 7;
        const T* om = x + 10;
        T F1 = kf * u[0], F2 = kf * u[1], F3 = kf * u[2], F4 = kf * u[3];
        T Fz = ((F1 + F2) + F3) + F4;
        T M1 = km * u[0], M2 = km * u[1], M3 = km * u[2], M4 = km * u[3];
        T tau0 = L * (F2 - F4), tau1 = L * (F3 - F1), tau2 = ((M1 - M2) + M3) - M4;
        xd[0] = v[0];
        xd[1] = v[1];
        xd[2] = v[2];
        T hq[4] = {0.5 * q[0], 0.5 * q[1], 0.5 * q[2], 0.5 * q[3]};
        T qo[4] = {T(0.0), om[0], om[1], om[2]};
        T qd[4];
        quat_mul(qd, hq, qo);
        xd[3] = qd[0]; xd[4] = qd[1]; xd[5] = qd[2]; xd[6] = qd[3];
        T qF[4] = {T(0.0), T(0.0), T(0.0), Fz};
        T t1[4], t2[4];
        quat_mul(t1, q, qF);
        T qi[4] = {q[0], -q[1], -q[2], -q[3]};
        quat_mul(t2, t1, qi);
        const double im = 1.0 / mass;
        xd[7] = 0.0 + im * t2[1];
        xd[8] = 0.0 + im * t2[2];
        xd[9] = -9.81 + im * t2[3];
        T Jo0 = Jx * om[0], Jo1 = Jy * om[1], Jo2 = Jz * om[2];
        T cr0 = om[1] * Jo2 - om[2] * Jo1, cr1 = om[2] * Jo0 - om[0] * Jo2, cr2 = om[0] * Jo1 - om[1] * Jo0;
        xd[10] = Jinvx * (tau0 - cr0);
        xd[11] = Jinvy * (tau1 - cr1);
        xd[12] = Jinvz * (tau2 - cr2);
    } else if constexpr (MODEL == 5) {  // dynamics/acrobot.jl:6  (torques = [0,1])
        f_dp_tau(xd, x, T(0.0), u[0]);
    } else {  // dynamics/doublependulum.jl:7
        f_dp_tau(xd, x, u[0], u[1]);
    }
}

// src/integration.jl:149-158 (rk3), :115-125 (rk4), :26-33 (midpoint)
template <int MODEL, int INTEG, class T> TOB_DEV void fd_model(T* xn, const T* x, const T* u, const T& dt) {
    constexpr int n = ModelDims<MODEL>::n;
    T k1[n], k2[n], tmp[n];
    if constexpr (INTEG == 0) {
        T k3[n];
        f_model<MODEL, T>(k1, x, u);
#pragma unroll
        for (int i = 0; i < n; i++) k1[i] = k1[i] * dt;
#pragma unroll
        for (int i = 0; i < n; i++) tmp[i] = x[i] + k1[i] / 2.0;
        f_model<MODEL, T>(k2, tmp, u);
#pragma unroll
        for (int i = 0; i < n; i++) k2[i] = k2[i] * dt;
#pragma unroll
        for (int i = 0; i < n; i++) tmp[i] = (x[i] - k1[i]) + 2.0 * k2[i];
        f_model<MODEL, T>(k3, tmp, u);
#pragma unroll
        for (int i = 0; i < n; i++) k3[i] = k3[i] * dt;
#pragma unroll
        for (int i = 0; i < n; i++) xn[i] = x[i] + div6_t((k1[i] + 4.0 * k2[i]) + k3[i]);
    } else if constexpr (INTEG == 1) {
        T k3[n], k4[n];
        f_model<MODEL, T>(k1, x, u);
#pragma unroll
        for (int i = 0; i < n; i++) k1[i] = k1[i] * dt;
#pragma unroll
        for (int i = 0; i < n; i++) tmp[i] = x[i] + k1[i] / 2.0;
        f_model<MODEL, T>(k2, tmp, u);
#pragma unroll
        for (int i = 0; i < n; i++) k2[i] = k2[i] * dt;
#pragma unroll
        for (int i = 0; i < n; i++) tmp[i] = x[i] + k2[i] / 2.0;
        f_model<MODEL, T>(k3, tmp, u);
#pragma unroll
        for (int i = 0; i < n; i++) k3[i] = k3[i] * dt;
#pragma unroll
        for (int i = 0; i < n; i++) tmp[i] = x[i] + k3[i];
        f_model<MODEL, T>(k4, tmp, u);
#pragma unroll
        for (int i = 0; i < n; i++) k4[i] = k4[i] * dt;
#pragma unroll
        for (int i = 0; i < n; i++) xn[i] = x[i] + div6_t(((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]);
    } else {
        f_model<MODEL, T>(k1, x, u);
        T hdt = dt / 2.0;
#pragma unroll
        for (int i = 0; i < n; i++) k1[i] = k1[i] * hdt;
#pragma unroll
        for (int i = 0; i < n; i++) tmp[i] = x[i] + k1[i];
        f_model<MODEL, T>(k2, tmp, u);
#pragma unroll
        for (int i = 0; i < n; i++) xn[i] = x[i] + k2[i] * dt;
    }
}

}  // namespace tob
