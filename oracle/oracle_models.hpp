// ORACLE — TEST INFRASTRUCTURE ONLY.
// Continuous dynamics of the reference's model zoo, templated on the scalar (double or Dual<P>),
// plus the explicit integrators.  Each function cites the reference lines it restates.
#pragma once
#include "oracle_math.hpp"

namespace orc {

// dynamics/double_integrator.jl:1-4
template <class T> inline void f_double_integrator(T* xd, const T* x, const T* u) {
    xd[0] = x[1];
    xd[1] = u[0];
}

// dynamics/pendulum.jl:3-12 :  xd2 = (u - m*g*lc*sin(x1) - b*x2)/I, products left to right
template <class T> inline void f_pendulum(T* xd, const T* x, const T* u) {
    const double m = 1.0, b = 0.1, lc = 0.5, I = 0.25, g = 9.81;
    T s, c;
    sincos_t(x[0], &s, &c);
    xd[0] = x[1];
    xd[1] = ((u[0] - ((m * g) * lc) * s) - b * x[1]) / I;
}

// dynamics/car.jl:3-8
template <class T> inline void f_car(T* xd, const T* x, const T* u) {
    T s, c;
    sincos_t(x[2], &s, &c);
    xd[0] = u[0] * c;
    xd[1] = u[0] * s;
    xd[2] = u[1];
}

// dynamics/cartpole.jl:9-36.  qdd = (-H) \ (C*qd + G - B*u): Julia's generic LU with partial
// pivoting on the VALUES (LinearAlgebra.generic_lufact!, reciprocal scaling of the column),
// then unit-lower forward and upper backward substitution.
template <class T> inline void f_cartpole(T* xd, const T* x, const T* u) {
    const double mc = 1.0, mp = 0.2, l = 0.5, g = 9.81;
    T s, c;
    if (std::isfinite(value(x[1]))) {
        sincos_t(x[1], &s, &c);
    } else {
        s = T(INFINITY);
        c = T(INFINITY);
    }
    const T qd0 = x[2], qd1 = x[3];
    // H = [mc+mp  mp*l*c ; mp*l*c  mp*l^2],  A = -H
    T A[2][2];
    A[0][0] = -T(mc + mp);
    A[0][1] = -((mp * l) * c);
    A[1][0] = -((mp * l) * c);
    A[1][1] = -T(mp * (l * l));
    // C*qd + G - B*u : C = [0 -mp*qd2*l*s; 0 0]  (generic matvec: zero + C[i,1]*qd1 + C[i,2]*qd2)
    T c12 = (((-mp) * qd1) * l) * s;
    T rhs[2];
    rhs[0] = ((T(0.0) * qd0 + c12 * qd1) + T(0.0)) - 1.0 * u[0];
    rhs[1] = ((T(0.0) * qd0 + T(0.0) * qd1) + ((mp * g) * l) * s) - 0.0 * u[0];
    // LU with partial pivoting, k = 1
    if (std::fabs(value(A[1][0])) > std::fabs(value(A[0][0]))) {
        for (int j = 0; j < 2; j++) { T t = A[0][j]; A[0][j] = A[1][j]; A[1][j] = t; }
        T t = rhs[0]; rhs[0] = rhs[1]; rhs[1] = t;
    }
    T inv00 = inv_t(A[0][0]);
    A[1][0] = A[1][0] * inv00;
    A[1][1] = A[1][1] - A[1][0] * A[0][1];
    // forward (unit lower), backward (upper)
    rhs[1] = rhs[1] - A[1][0] * rhs[0];
    T q1 = rhs[1] / A[1][1];
    rhs[0] = rhs[0] - A[0][1] * q1;
    T q0 = rhs[0] / A[0][0];
    xd[0] = qd0;
    xd[1] = qd1;
    xd[2] = q0;
    xd[3] = q1;
}

// dynamics/quaternions.jl:23-27 : q2*q1 -> w = s1*s2 - v1.v2 ; v = s1*v2 + s2*v1 + v2 x v1
template <class T> inline void quat_mul(T* out, const T* q2, const T* q1) {
    const T s1 = q1[0], s2 = q2[0];
    const T* v1 = q1 + 1;
    const T* v2 = q2 + 1;
    out[0] = s1 * s2 - ((v1[0] * v2[0] + v1[1] * v2[1]) + v1[2] * v2[2]);
    T cx = v2[1] * v1[2] - v2[2] * v1[1];
    T cy = v2[2] * v1[0] - v2[0] * v1[2];
    T cz = v2[0] * v1[1] - v2[1] * v1[0];
    out[1] = (s1 * v2[0] + s2 * v1[0]) + cx;
    out[2] = (s1 * v2[1] + s2 * v1[1]) + cy;
    out[3] = (s1 * v2[2] + s2 * v1[2]) + cz;
}

// dynamics/quadrotor.jl:10-71, parameters :1-7
template <class T> inline void f_quadrotor(T* xd, const T* x, const T* u) {
    const double mass = 0.5, Jx = 0.0023, Jy = 0.0023, Jz = 0.004, L = 0.1750, kf = 1.0, km = 0.0245;
    const double Jinvx = 1.0 / 0.0023, Jinvy = 1.0 / 0.0023, Jinvz = 1.0 / 0.004;
    // q = normalize(q): inv(norm)*q, norm = sqrt(abs2 sum) (StaticArrays 0.11)
    T nrm = sqrt_t(((sq_t(x[3]) + sq_t(x[4])) + sq_t(x[5])) + sq_t(x[6]));
    T in = inv_t(nrm);
    T q[4] = {in * x[3], in * x[4], in * x[5], in * x[6]};
    const T* v = x + 7;
    const T* om = x + 10;
    T F1 = kf * u[0], F2 = kf * u[1], F3 = kf * u[2], F4 = kf * u[3];
    T Fz = ((F1 + F2) + F3) + F4;
    T M1 = km * u[0], M2 = km * u[1], M3 = km * u[2], M4 = km * u[3];
    T tau[3] = {L * (F2 - F4), L * (F3 - F1), ((M1 - M2) + M3) - M4};
    xd[0] = v[0];
    xd[1] = v[1];
    xd[2] = v[2];
    // 0.5*q*Quaternion(0, omega)
    T hq[4] = {0.5 * q[0], 0.5 * q[1], 0.5 * q[2], 0.5 * q[3]};
    T qo[4] = {T(0.0), om[0], om[1], om[2]};
    T qd[4];
    quat_mul(qd, hq, qo);
    xd[3] = qd[0]; xd[4] = qd[1]; xd[5] = qd[2]; xd[6] = qd[3];
    // g + (1/m)*(q*F),  q*F = vec(q*Quaternion(0,F)*inv(q))
    T qF[4] = {T(0.0), T(0.0), T(0.0), Fz};
    T t1[4], t2[4];
    quat_mul(t1, q, qF);
    T qi[4] = {q[0], -q[1], -q[2], -q[3]};
    quat_mul(t2, t1, qi);
    const double im = 1.0 / mass;
    xd[7] = 0.0 + im * t2[1];
    xd[8] = 0.0 + im * t2[2];
    xd[9] = -9.81 + im * t2[3];
    // Jinv*(tau - cross(omega, J*omega))
    T Jo[3] = {Jx * om[0], Jy * om[1], Jz * om[2]};
    T cr[3] = {om[1] * Jo[2] - om[2] * Jo[1], om[2] * Jo[0] - om[0] * Jo[2], om[0] * Jo[1] - om[1] * Jo[0]};
    xd[10] = Jinvx * (tau[0] - cr[0]);
    xd[11] = Jinvy * (tau[1] - cr[1]);
    xd[12] = Jinvz * (tau[2] - cr[2]);
}

// URDF double pendulum through RigidBodyDynamics 2.1.0 (third-party, not in tree): closed-form
// manipulator equations, SURVEY.md Appendix A (dynamics/urdf/doublependulum.urdf:15-63,
// src/model.jl:394-431); joint damping NOT applied (Appendix F item 1).
// tau = (t1, t2); acrobot passes t1 = 0 (torques=[0,1], dynamics/acrobot.jl:6).
template <class T> inline void f_doublependulum_tau(T* xd, const T* x, const T& t1, const T& t2) {
    const double m1 = 1.0, m2 = 1.0, l1 = 1.0, lc1 = 0.5, lc2 = 1.0, I1 = 0.083, I2 = 0.33, g = 9.81;
    T s1, c1, s2, c2, s12, c12;
    sincos_t(x[0], &s1, &c1);
    sincos_t(x[1], &s2, &c2);
    sincos_t(x[0] + x[1], &s12, &c12);
    const T qd1 = x[2], qd2 = x[3];
    T M11 = (((I1 + I2) + m1 * (lc1 * lc1)) + m2 * ((l1 * l1 + lc2 * lc2) + ((2.0 * l1) * lc2) * c2));
    T M12 = I2 + m2 * (lc2 * lc2 + (l1 * lc2) * c2);
    double M22 = I2 + m2 * (lc2 * lc2);
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
template <class T> inline void f_acrobot(T* xd, const T* x, const T* u) { f_doublependulum_tau(xd, x, T(0.0), u[0]); }
template <class T> inline void f_doublependulum(T* xd, const T* x, const T* u) { f_doublependulum_tau(xd, x, u[0], u[1]); }

struct ModelInfo { int n, m; };
inline ModelInfo model_info(int model) {
    switch (model) {
        case 0: return {2, 1};
        case 1: return {2, 1};
        case 2: return {3, 2};
        case 3: return {4, 1};
        case 4: return {13, 4};
        case 5: return {4, 1};
        case 6: return {4, 2};
    }
    return {0, 0};
}

template <class T> inline void f_model(int model, T* xd, const T* x, const T* u) {
    switch (model) {
        case 0: f_double_integrator(xd, x, u); break;
        case 1: f_pendulum(xd, x, u); break;
        case 2: f_car(xd, x, u); break;
        case 3: f_cartpole(xd, x, u); break;
        case 4: f_quadrotor(xd, x, u); break;
        case 5: f_acrobot(xd, x, u); break;
        case 6: f_doublependulum(xd, x, u); break;
    }
}

// Discrete step x+ = fd(x,u,dt).  src/integration.jl:149-158 (rk3), :115-125 (rk4), :26-33 (midpoint)
template <class T> inline void fd_model(int model, int integ, int n, T* xn, const T* x, const T* u, const T& dt) {
    const int MAXN = 16;
    T k1[MAXN], k2[MAXN], k3[MAXN], k4[MAXN], tmp[MAXN];
    if (integ == 0) {  // rk3
        f_model(model, k1, x, u);
        for (int i = 0; i < n; i++) k1[i] = k1[i] * dt;
        for (int i = 0; i < n; i++) tmp[i] = x[i] + k1[i] / 2.0;
        f_model(model, k2, tmp, u);
        for (int i = 0; i < n; i++) k2[i] = k2[i] * dt;
        for (int i = 0; i < n; i++) tmp[i] = (x[i] - k1[i]) + 2.0 * k2[i];
        f_model(model, k3, tmp, u);
        for (int i = 0; i < n; i++) k3[i] = k3[i] * dt;
        for (int i = 0; i < n; i++) xn[i] = x[i] + ((k1[i] + 4.0 * k2[i]) + k3[i]) / 6.0;
    } else if (integ == 1) {  // rk4
        f_model(model, k1, x, u);
        for (int i = 0; i < n; i++) k1[i] = k1[i] * dt;
        for (int i = 0; i < n; i++) tmp[i] = x[i] + k1[i] / 2.0;
        f_model(model, k2, tmp, u);
        for (int i = 0; i < n; i++) k2[i] = k2[i] * dt;
        for (int i = 0; i < n; i++) tmp[i] = x[i] + k2[i] / 2.0;
        f_model(model, k3, tmp, u);
        for (int i = 0; i < n; i++) k3[i] = k3[i] * dt;
        for (int i = 0; i < n; i++) tmp[i] = x[i] + k3[i];
        f_model(model, k4, tmp, u);
        for (int i = 0; i < n; i++) k4[i] = k4[i] * dt;
        for (int i = 0; i < n; i++) xn[i] = x[i] + (((k1[i] + 2.0 * k2[i]) + 2.0 * k3[i]) + k4[i]) / 6.0;
    } else {  // midpoint: xdot = f(x,u); xdot .*= dt/2; xdot = f(x + xdot, u); x + xdot*dt
        f_model(model, k1, x, u);
        T hdt = dt / 2.0;
        for (int i = 0; i < n; i++) k1[i] = k1[i] * hdt;
        for (int i = 0; i < n; i++) tmp[i] = x[i] + k1[i];
        f_model(model, k2, tmp, u);
        for (int i = 0; i < n; i++) xn[i] = x[i] + k2[i] * dt;
    }
}

}  // namespace orc
