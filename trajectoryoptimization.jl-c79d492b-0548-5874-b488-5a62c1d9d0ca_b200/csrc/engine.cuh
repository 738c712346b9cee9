// B200 engine: the whole iLQR / AL-iLQR solve of one problem runs resident in ONE WARP of a
// persistent kernel; warps pull problems from a device-side queue (ragged iteration counts
// balance themselves), and the batch is the grid.  Phases of one iLQR iteration:
//
//   jacobians()     knot×partial-parallel: lanes own (knot, partial-chunk) work items and push
//                   dual numbers through the rk3 closure           (src/model.jl:491-512)
//   backwardpass()  time-sequential Riccati recursion with S, A, B, Q blocks and gains staged in
//                   shared memory; lanes own output elements, every inner product is a
//                   sequential FMA chain (bit-reproducible)        (backward_pass.jl:9-85)
//   forwardpass()   line search with the 21 step sizes α = 2^-lane evaluated in parallel, one
//                   closed-loop rk3 rollout + AL cost per lane     (forward_pass.jl:5-85)
//
// plus the augmented-Lagrangian outer loop (dual / penalty / active set) on the same warp.
// FP64 CUDA cores only: the contractions are ≤ 14×14, tensor cores do not apply.
#pragma once
#include <stdint.h>

#include "engine_host.h"
#include "models.cuh"

namespace tob {

template <int MODEL_, int INTEG_, bool INF_, bool MT_, int PC_>
struct Cfg {
    static constexpr int MODEL = MODEL_, INTEG = INTEG_, PC = PC_;
    static constexpr bool INF = INF_, MT = MT_;
    static constexpr int n0 = ModelDims<MODEL_>::n, m0 = ModelDims<MODEL_>::m;
    static constexpr int nq = n0, mq = m0 + (INF_ ? n0 : 0);
    static constexpr int n = n0 + (MT_ ? 1 : 0), m = mq + (MT_ ? 1 : 0);
    static constexpr int PT = n0 + m0 + (MT_ ? 1 : 0);  // partial directions of the inner model
    static constexpr int NCH = (PT + PC_ - 1) / PC_;
    static constexpr int ZS = n0 * PT;                  // inner Jacobian doubles per knot
    static constexpr int LDZ = (n + m + 1) & ~1;        // lockstep engine: row-major [A B] per knot, even row length
    static constexpr int ZA = n * LDZ;
    static constexpr int KDS = m * n + m;               // K and d per knot
    static constexpr int QS = n + m + n * n + m * m + m * n;
};

// workspace layout (doubles) of one warp
struct WsLayout {
    unsigned long long X, U, Z, KD, LAM, MU, XB, UB, QST, CB, total;
};
template <class C>
__host__ __device__ inline WsLayout ws_layout(int N, int Ptot, bool candidates = true) {
    WsLayout L;
    unsigned long long o = 0;
    L.X = o;   o += (unsigned long long)N * C::n;
    L.U = o;   o += (unsigned long long)(N - 1) * C::m;
    o = (o + 1ull) & ~1ull;  // Z blocks are 16-byte aligned (cp.async 16 B)
    L.Z = o;   o += (unsigned long long)(N - 1) * (candidates ? C::ZS : C::ZA);
    L.KD = o;  o += (unsigned long long)(N - 1) * C::KDS;
    L.LAM = o; o += Ptot;
    L.MU = o;  o += Ptot;
    // 32 candidate trajectories of the parallel line search (persistent engine only)
    L.XB = o;  o += candidates ? (unsigned long long)N * C::n * 32 : 0ull;
    L.UB = o;  o += candidates ? (unsigned long long)(N - 1) * C::m * 32 : 0ull;
    L.QST = o; o += (unsigned long long)N * C::QS;  // slot N-1: terminal expansion (lockstep CTA pass)
    L.CB = o;  o += 2ull * N;
    L.total = (o + 15ull) & ~15ull;
    return L;
}

// shared memory of one warp
template <class C>
struct Smem {
    double S[C::n * C::n], Sx[C::n];
    double A[C::n * C::n], B[C::n * C::m];
    double Qxx[C::n * C::n], Quu[C::m * C::m], Qux[C::m * C::n], Qx[C::n], Qu[C::m];
    double T[C::n * C::n], Tu[C::m * C::n];
    double K[C::m * C::n], d[C::m], KQ[C::n * C::m];
    double xk[C::n], uk[C::m], vQx[C::nq], vQu[C::mq];
    double stage[2][C::n + C::m + C::m * C::n + C::m];
};

#define FULL 0xffffffffu
TOB_DEV double bcast(double v, int src) { return __shfl_sync(FULL, v, src); }
TOB_DEV double dmax(double a, double b) { return (a < b) ? b : a; }  // std::max(a,b)
TOB_DEV double dmin(double a, double b) { return (b < a) ? b : a; }  // std::min(a,b)

// ------------------------------------------------------------------------------------------
// per-lane scalar pieces (used by the rollout lanes and by the knot-parallel phases)
// ------------------------------------------------------------------------------------------
template <class C>
TOB_DEV void dyn_eval(const DevProblem& P, double* xn, const double* x, const double* u) {
    double dt = P.dt;
    if constexpr (C::MT) {
        double h = u[C::m - 1];
        dt = h * h;
    }
    fd_model<C::MODEL, C::INTEG, double>(xn, x, u, dt);
    if constexpr (C::INF) {
#pragma unroll
        for (int i = 0; i < C::n0; i++) xn[i] = xn[i] + u[C::m0 + i];
    }
    if constexpr (C::MT) xn[C::n - 1] = u[C::m - 1];
}

// 0.5*x'Q*x + 0.5*u'R*u + q'x + r'u + c + u'H*x        (src/cost.jl:171-173)
template <class C>
TOB_DEV double quad_stage(const DevProblem& P, const double* x, const double* u) {
    constexpr int n = C::nq, m = C::mq;
    double t1 = 0.0, t2 = 0.0;
    if (P.q_diag) {
#pragma unroll
        for (int j = 0; j < n; j++) t1 = fma((0.5 * x[j]) * __ldg(&P.Q[j * n + j]), x[j], t1);
    } else {
#pragma unroll
        for (int j = 0; j < n; j++) {
            double w = 0.0;
#pragma unroll
            for (int i = 0; i < n; i++) w = fma(0.5 * x[i], __ldg(&P.Q[j * n + i]), w);
            t1 = fma(w, x[j], t1);
        }
    }
    if (P.r_diag) {
#pragma unroll
        for (int j = 0; j < m; j++) t2 = fma((0.5 * u[j]) * __ldg(&P.R[j * m + j]), u[j], t2);
    } else {
#pragma unroll
        for (int j = 0; j < m; j++) {
            double w = 0.0;
#pragma unroll
            for (int i = 0; i < m; i++) w = fma(0.5 * u[i], __ldg(&P.R[j * m + i]), w);
            t2 = fma(w, u[j], t2);
        }
    }
    double t3 = 0.0, t4 = 0.0;
#pragma unroll
    for (int i = 0; i < n; i++) t3 = fma(__ldg(&P.q[i]), x[i], t3);
#pragma unroll
    for (int i = 0; i < m; i++) t4 = fma(__ldg(&P.r[i]), u[i], t4);
    double s = (((t1 + t2) + t3) + t4) + P.c;
    if (!P.h_zero) {
        double t6 = 0.0;
#pragma unroll
        for (int j = 0; j < n; j++) {
            double w = 0.0;
#pragma unroll
            for (int i = 0; i < m; i++) w = fma(u[i], __ldg(&P.H[j * m + i]), w);
            t6 = fma(w, x[j], t6);
        }
        s = s + t6;
    } else {
        s = s + 0.0;
    }
    return s;
}
template <class C>
TOB_DEV double stage_cost(const DevProblem& P, const double* x, const double* u) {
    if constexpr (C::MT) {
        double tau = u[C::m - 1];
        double dt = tau * tau;
        return quad_stage<C>(P, x, u) * dt + P.R_mt * (tau * tau);
    } else {
        return quad_stage<C>(P, x, u) * P.dt;
    }
}
template <class C>
TOB_DEV double term_cost(const DevProblem& P, const double* x) {
    constexpr int n = C::nq;
    double acc = 0.0;
    if (P.qf_diag) {
#pragma unroll
        for (int j = 0; j < n; j++) acc = fma((0.5 * x[j]) * __ldg(&P.Qf[j * n + j]), x[j], acc);
    } else {
#pragma unroll
        for (int j = 0; j < n; j++) {
            double w = 0.0;
#pragma unroll
            for (int i = 0; i < n; i++) w = fma(0.5 * x[i], __ldg(&P.Qf[j * n + i]), w);
            acc = fma(w, x[j], acc);
        }
    }
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < n; i++) t = fma(__ldg(&P.qf[i]), x[i], t);
    return (acc + t) + P.cf;
}

// constraint row value (src/constraints.jl:212-227,299-314; src/utils.jl:140-156; minimum_time.jl:112-124)
template <class C>
TOB_DEV double row_value(const DevRow& r, const double* x, const double* u) {
    switch (r.kind) {
        case DR_LIN: {
            double z = 0.0;
            const int col = r.col;
#pragma unroll
            for (int i = 0; i < C::n; i++) if (col == i) z = x[i];
#pragma unroll
            for (int i = 0; i < C::m; i++) if (col == C::n + i) z = u[i];
            return (r.sign > 0) ? (z - r.a) : (r.a - z);
        }
        case DR_CIRCLE: {
            double dx = x[0] - r.a, dy = x[1] - r.b;
            return -(((dx * dx) + (dy * dy)) - (r.r * r.r));
        }
        case DR_SPHERE: {
            if constexpr (C::n >= 3) {
                double dx = x[0] - r.a, dy = x[1] - r.b, dz = x[2] - r.c;
                return -((((dx * dx) + (dy * dy)) + (dz * dz)) - (r.r * r.r));
            } else {
                return 0.0;
            }
        }
        default: return u[C::m - 1] - x[C::n - 1];
    }
}
// entry `col` (column of z̄ = [x̄;ū]) of the row's Jacobian
template <class C>
TOB_DEV double row_jac(const DevRow& r, const double* x, int col) {
    switch (r.kind) {
        case DR_LIN: return (col == r.col) ? ((r.sign > 0) ? 1.0 : -1.0) : 0.0;
        case DR_CIRCLE: return (col == 0) ? -(2.0 * (x[0] - r.a)) : ((col == 1) ? -(2.0 * (x[1] - r.b)) : 0.0);
        case DR_SPHERE:
            if constexpr (C::n >= 3)
                return (col == 0) ? -(2.0 * (x[0] - r.a)) : ((col == 1) ? -(2.0 * (x[1] - r.b)) : ((col == 2) ? -(2.0 * (x[2] - r.c)) : 0.0));
            else
                return 0.0;
        default: return (col == C::n + C::m - 1) ? 1.0 : ((col == C::n - 1) ? -1.0 : 0.0);
    }
}

// Σ over the rows of knot k of  λ'c + ((1/2 c')Diag(a∘μ))c      (augmented_lagrangian_methods.jl:284-286)
template <class C>
TOB_DEV double knot_al_cost(const DevProblem& P, int k, const double* lam, const double* mu, const double* x, const double* u) {
    const int rb = P.knot_row_begin[k], rc = P.knot_row_count[k], lo = P.knot_lam_off[k];
    double t1 = 0.0, t2 = 0.0;
    for (int i = 0; i < rc; i++) {
        const DevRow r = P.rows[rb + i];
        double c = row_value<C>(r, x, u);
        double l = lam[lo + i];
        bool act = r.eq ? true : ((c >= 0.0) || (l > 0.0));
        double am = act ? mu[lo + i] : 0.0;
        t1 = fma(l, c, t1);
        t2 = fma((0.5 * c) * am, c, t2);
    }
    return t1 + t2;
}

// ------------------------------------------------------------------------------------------
// the per-warp solver
// ------------------------------------------------------------------------------------------
template <class C>
struct Solver {
    const DevProblem& P;
    const DevBatch& Bt;
    const DevCtl& ctl;
    Smem<C>& sm;
    double* ws;
    WsLayout L;
    const int lane;
    int b;  // problem index
    bool al_on;
    TOiLQROptions io;  // live copy (tolerances are switched by the outer loop)
    // solver state (warp-uniform, replicated in every lane)
    double rho, drho;
    int iterations, dJ_zero, steps, status;
    double last_dJ, last_grad, last_cost;
    double fp_expected, fp_z, fp_alpha;
    int outer_idx;
    int n_inner_rec, n_outer_rec;
    unsigned long long ls_count;
    double x0[C::n];

    __device__ Solver(const DevProblem& P_, const DevBatch& B_, const DevCtl& c_, Smem<C>& s_, double* ws_, int lane_, bool candidates = true)
        : P(P_), Bt(B_), ctl(c_), sm(s_), ws(ws_), L(ws_layout<C>(P_.N, P_.Ptot, candidates)), lane(lane_) {}

    __device__ double* X(int k) { return ws + L.X + (size_t)k * C::n; }
    __device__ double* U(int k) { return ws + L.U + (size_t)k * C::m; }
    __device__ double* Z(int k) { return ws + L.Z + (size_t)k * C::ZS; }
    __device__ double* KD(int k) { return ws + L.KD + (size_t)k * C::KDS; }
    __device__ double* lam() { return ws + L.LAM; }
    __device__ double* mu() { return ws + L.MU; }
    __device__ double* CB() { return ws + L.CB; }

    // ---- regularisation schedule (ilqr_methods.jl:164-176) ----
    __device__ void reg_update(bool increase) {
        const double f = io.bp_reg_increase_factor;
        if (increase) {
            drho = dmax(drho * f, f);
            rho = dmax(rho * drho, io.bp_reg_min);
        } else {
            drho = dmin(drho / f, 1.0 / f);
            rho = rho * drho * ((rho * drho > io.bp_reg_min) ? 1.0 : 0.0);
        }
    }

    // ---- trajectory cost at the CURRENT X,U (objective.jl:40-48 + AL terms) ----
    __device__ double eval_cost() {
        const int N = P.N;
        double* cb = CB();
        for (int k = lane; k < N; k += 32) {
            double x[C::n], u[C::m];
#pragma unroll
            for (int i = 0; i < C::n; i++) x[i] = X(k)[i];
            if (k < N - 1) {
#pragma unroll
                for (int i = 0; i < C::m; i++) u[i] = U(k)[i];
            } else {
#pragma unroll
                for (int i = 0; i < C::m; i++) u[i] = 0.0;
            }
            cb[k] = (k < N - 1) ? stage_cost<C>(P, x, u) : term_cost<C>(P, x);
            cb[N + k] = al_on ? knot_al_cost<C>(P, k, lam(), mu(), x, u) : 0.0;
        }
        __syncwarp();
        double J = 0.0, Jc = 0.0;
        for (int k = 0; k < N; k++) J += cb[k];
        if (!al_on) { __syncwarp(); return J; }
        for (int k = 0; k < N; k++) Jc += cb[N + k];
        __syncwarp();
        return J + Jc;
    }

    // ---- max constraint violation at the current X,U (augmented_lagrangian_methods.jl:171-184) ----
    __device__ double max_violation() {
        const int N = P.N;
        double cmax = 0.0;
        for (int k = lane; k < N; k += 32) {
            const int rb = P.knot_row_begin[k], rc = P.knot_row_count[k];
            if (rc == 0) continue;
            double x[C::n], u[C::m];
#pragma unroll
            for (int i = 0; i < C::n; i++) x[i] = X(k)[i];
#pragma unroll
            for (int i = 0; i < C::m; i++) u[i] = (k < N - 1) ? U(k)[i] : 0.0;
            double e = 0.0, mi = -__longlong_as_double(0x7ff0000000000000LL);
            bool has_i = false;
            for (int i = 0; i < rc; i++) {
                const DevRow r = P.rows[rb + i];
                double c = row_value<C>(r, x, u);
                if (r.eq) e = dmax(e, fabs(c));
                else { has_i = true; mi = dmax(mi, c); }
            }
            cmax = dmax(e, cmax);
            if (has_i) cmax = dmax(dmax(0.0, mi), cmax);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cmax = dmax(cmax, __shfl_xor_sync(FULL, cmax, o));
        return cmax;
    }

    // ---- dual and penalty updates (augmented_lagrangian_methods.jl:107-126) ----
    __device__ void dual_penalty_update() {
        const int N = P.N;
        const TOALOptions& o = ctl.o;
        for (int k = 0; k < N; k++) {
            const int rb = P.knot_row_begin[k], rc = P.knot_row_count[k], lo = P.knot_lam_off[k];
            if (rc == 0) continue;
            double x[C::n], u[C::m];
#pragma unroll
            for (int i = 0; i < C::n; i++) x[i] = X(k)[i];
#pragma unroll
            for (int i = 0; i < C::m; i++) u[i] = (k < N - 1) ? U(k)[i] : 0.0;
            for (int i = lane; i < rc; i += 32) {
                const DevRow r = P.rows[rb + i];
                double c = row_value<C>(r, x, u);
                double v = lam()[lo + i] + mu()[lo + i] * c;
                v = dmax(o.dual_min, dmin(o.dual_max, v));
                if (!r.eq) v = dmax(0.0, v);
                lam()[lo + i] = v;
                mu()[lo + i] = dmax(0.0, dmin(o.penalty_max, o.penalty_scaling * mu()[lo + i]));
            }
        }
        __syncwarp();
    }

    // ---- calculate_gradient (ilqr_methods.jl:91-102) ----
    __device__ double gradient() {
        if (io.gradient_type == TO_GRAD_FEEDFORWARD) return gradient_feedforward();
        if (io.gradient_type == TO_GRAD_L2 || io.gradient_type == TO_GRAD_LINF) return gradient_cost(io.gradient_type == TO_GRAD_LINF);
        return gradient_todorov();
    }
    // norm(v) of a short Vector{Float64}: LinearAlgebra.generic_norm2 (Julia 1.1 generic.jl; BLAS nrm2 only from 32 elements on)
    __device__ static double julia_norm2(const double* v, int len) {
        double maxabs = 0.0;
        for (int i = 0; i < len; i++) { const double a = fabs(v[i]); maxabs = (maxabs != maxabs || maxabs > a) ? maxabs : a; }
        if (maxabs == 0.0 || isinf(maxabs)) return maxabs;
        if (isfinite((double)len * maxabs * maxabs) && maxabs * maxabs != 0.0) {
            double sum = 0.0;
            for (int i = 0; i < len; i++) sum += v[i] * v[i];
            return sqrt(sum);
        }
        const double inv = 1.0 / maxabs;
        double sum = 0.0;
        for (int i = 0; i < len; i++) { const double t = fabs(v[i]) * inv; sum += t * t; }
        return maxabs * sqrt(sum);
    }
    // gradient_feedforward (ilqr_methods.jl:133-137): norm(solver.d, Inf) over a vector of vectors = the largest 2-norm of a d[k]
    __device__ double gradient_feedforward() {
        const int N = P.N;
        double* cb = CB();
        for (int k = lane; k < N - 1; k += 32) cb[k] = julia_norm2(KD(k) + C::m * C::n, C::m);
        __syncwarp();
        double mx = 0.0;
        for (int k = 0; k < N - 1; k++) { const double a = cb[k]; mx = (k == 0) ? a : ((mx != mx || mx > a) ? mx : a); }
        __syncwarp();
        return mx;
    }
    // :ℓ2 / :ℓinf (ilqr_methods.jl:97-116): norm of [Q1.x; Q1.u; ...; QN.x] of the cost expansion at the current X, U
    __device__ double gradient_cost(bool inf_norm) {
        const int N = P.N;
        double mx = 0.0, sum = 0.0;
        bool first = true;
        for (int k = 0; k < N; k++) {
            expansion(k);  // ends with __syncwarp; lane 0 folds the knot's entries in order
            const int cnt = (k < N - 1) ? C::n + C::m : C::n;
            for (int e = 0; e < cnt; e++) {
                const double v = (e < C::n) ? sm.Qx[e] : sm.Qu[e - C::n];
                const double a = fabs(v);
                mx = first ? a : ((mx != mx || mx > a) ? mx : a);
                first = false;
                sum += v * v;
            }
            __syncwarp();
        }
        return inf_norm ? mx : sqrt(sum);
    }
    // ---- Todorov gradient (ilqr_methods.jl:122-129): mean over N of max_i |d_i|/(|u_i|+1) ----
    __device__ double gradient_todorov() {
        const int N = P.N;
        double* cb = CB();
        const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
        for (int k = lane; k < N - 1; k += 32) {
            const double* d = KD(k) + C::m * C::n;
            double mx = ninf;
            bool isnan_ = false;
#pragma unroll
            for (int i = 0; i < C::m; i++) {
                double v = fabs(d[i]) / (fabs(U(k)[i]) + 1.0);
                if (!isnan_) {
                    if (v != v) { mx = v; isnan_ = true; }
                    else mx = dmax(mx, v);
                }
            }
            cb[k] = mx;
        }
        __syncwarp();
        double s = 0.0;
        for (int k = 0; k < N - 1; k++) s += cb[k];
        s += 0.0;
        __syncwarp();
        return s / (double)N;
    }

    // ---- discrete dynamics Jacobians by dual numbers (src/model.jl:491-512) ----
    __device__ void jacobians() {
        constexpr int PC = C::PC;
        typedef Dual<PC> D;
        const int N = P.N;
        const int items = (N - 1) * C::NCH;
        for (int it = lane; it < items; it += 32) {
            const int k = it / C::NCH, ch = it - k * C::NCH;
            const int s0 = ch * PC;
            D xs[C::n0], us[C::m0], dts, xn[C::n0];
            const double* xk = X(k);
            const double* uk = U(k);
#pragma unroll
            for (int i = 0; i < C::n0; i++) {
                xs[i] = D(xk[i]);
#pragma unroll
                for (int j = 0; j < PC; j++) if (s0 + j == i) xs[i].p[j] = 1.0;
            }
#pragma unroll
            for (int i = 0; i < C::m0; i++) {
                us[i] = D(uk[i]);
#pragma unroll
                for (int j = 0; j < PC; j++) if (s0 + j == C::n0 + i) us[i].p[j] = 1.0;
            }
            double dt = P.dt;
            if constexpr (C::MT) {
                double h = uk[C::m - 1];
                dt = h * h;
            }
            dts = D(dt);
            if constexpr (C::MT) {
#pragma unroll
                for (int j = 0; j < PC; j++) if (s0 + j == C::n0 + C::m0) dts.p[j] = 1.0;
            }
            fd_model<C::MODEL, C::INTEG, D>(xn, xs, us, dts);
            double* z = Z(k);
#pragma unroll
            for (int j = 0; j < PC; j++) {
                if (s0 + j < C::PT) {
#pragma unroll
                    for (int i = 0; i < C::n0; i++) z[(s0 + j) * C::n0 + i] = xn[i].p[j];
                }
            }
        }
        __syncwarp();
    }

    // ---- load A_k, B_k of the augmented model into shared memory ----
    // (add_slack_controls: src/model.jl:761-779; add_min_time_controls: minimum_time.jl:85-104)
    __device__ void load_AB(int k) {
        constexpr int n = C::n, m = C::m, n0 = C::n0, m0 = C::m0;
        const double* z = Z(k);
        for (int e = lane; e < n * n; e += 32) {
            const int i = e % n, j = e / n;
            sm.A[e] = (i < n0 && j < n0) ? z[j * n0 + i] : 0.0;
        }
        double h2 = 0.0;
        if constexpr (C::MT) h2 = 2.0 * U(k)[m - 1];
        for (int e = lane; e < n * m; e += 32) {
            const int i = e % n, j = e / n;
            double v = 0.0;
            if (j < m0) {
                if (i < n0) v = z[(n0 + j) * n0 + i];
            } else if (C::INF && j < m0 + n0) {
                v = (i == j - m0) ? 1.0 : 0.0;
            } else if (C::MT && j == m - 1) {
                if (i < n0) v = z[(n0 + m0) * n0 + i] * h2;
                else if (i == n - 1) v = 1.0;
            }
            sm.B[e] = v;
        }
    }

    // ---- cost expansion of knot k (+ AL terms) into shared memory ----
    // (src/cost.jl:183-198, minimum_time.jl:161-204, augmented_lagrangian_methods.jl:186-229)
    __device__ void expansion(int k) {
        constexpr int n = C::n, m = C::m, nq = C::nq, mq = C::mq;
        const int N = P.N;
        const bool term = (k == N - 1);
        if (lane < n) sm.xk[lane] = X(k)[lane];
        if (lane < m) sm.uk[lane] = term ? 0.0 : U(k)[lane];
        __syncwarp();
        const double* x = sm.xk;
        const double* u = sm.uk;
        const double* Qm = term ? P.Qf : P.Q;
        const double* qv = term ? P.qf : P.q;
        const bool qd = term ? P.qf_diag : P.q_diag;
        // gradients of the wrapped quadratic cost (unscaled)
        for (int i = lane; i < nq; i += 32) {
            double a = 0.0, bq = 0.0;
            if (qd) a = fma(__ldg(&Qm[i * nq + i]), x[i], a);
            else for (int j = 0; j < nq; j++) a = fma(__ldg(&Qm[j * nq + i]), x[j], a);
            if (!term && !P.h_zero) for (int j = 0; j < mq; j++) bq = fma(__ldg(&P.H[i * mq + j]), u[j], bq);
            sm.vQx[i] = term ? (a + __ldg(&qv[i])) : ((a + __ldg(&qv[i])) + bq);
        }
        if (!term)
            for (int i = lane; i < mq; i += 32) {
                double a = 0.0, bq = 0.0;
                if (P.r_diag) a = fma(__ldg(&P.R[i * mq + i]), u[i], a);
                else for (int j = 0; j < mq; j++) a = fma(__ldg(&P.R[j * mq + i]), u[j], a);
                if (!P.h_zero) for (int j = 0; j < nq; j++) bq = fma(__ldg(&P.H[j * mq + i]), x[j], bq);
                sm.vQu[i] = (a + __ldg(&P.r[i])) + bq;
            }
        __syncwarp();
        double dt = P.dt, tau = 0.0, l1 = 0.0;
        if (C::MT && !term) {
            tau = u[m - 1];
            dt = tau * tau;
            double xr[n], ur[m];
#pragma unroll
            for (int i = 0; i < n; i++) xr[i] = x[i];
#pragma unroll
            for (int i = 0; i < m; i++) ur[i] = u[i];
            l1 = quad_stage<C>(P, xr, ur);
        }
        // base blocks
        for (int e = lane; e < n * n; e += 32) {
            const int i = e % n, j = e / n;
            double v = 0.0;
            if (i < nq && j < nq) v = term ? __ldg(&Qm[j * nq + i]) : __ldg(&Qm[j * nq + i]) * dt;
            if (C::MT && i == n - 1 && j == n - 1) v = P.R_mt;
            sm.Qxx[e] = v;
        }
        for (int i = lane; i < n; i += 32) {
            double v = 0.0;
            if (i < nq) v = term ? sm.vQx[i] : sm.vQx[i] * dt;
            if (C::MT && i == n - 1) v = P.R_mt * x[n - 1];
            sm.Qx[i] = v;
        }
        if (!term) {
            for (int e = lane; e < m * m; e += 32) {
                const int i = e % m, j = e / m;
                double v = 0.0;
                if (i < mq && j < mq) v = __ldg(&P.R[j * mq + i]) * dt;
                if (C::MT) {
                    if (j == m - 1 && i < mq) v = (2.0 * tau) * sm.vQu[i];
                    if (i == m - 1 && j < mq) v = (2.0 * tau) * sm.vQu[j];
                    if (i == m - 1 && j == m - 1) v = 2.0 * l1 + P.R_mt;
                }
                sm.Quu[e] = v;
            }
            for (int e = lane; e < m * n; e += 32) {
                const int i = e % m, j = e / m;  // ux is m×n
                double v = 0.0;
                if (i < mq && j < nq) v = P.h_zero ? (0.0 * dt) : __ldg(&P.H[j * mq + i]) * dt;
                if (C::MT && i == m - 1 && j < nq) v = (2.0 * tau) * sm.vQx[j];
                sm.Qux[e] = v;
            }
            for (int i = lane; i < m; i += 32) {
                double v = 0.0;
                if (i < mq) v = sm.vQu[i] * dt;
                if (C::MT && i == m - 1) v = tau * (2.0 * l1 + P.R_mt);
                sm.Qu[i] = v;
            }
        }
        __syncwarp();
        if (!al_on) return;
        const int rb = P.knot_row_begin[k], rc = P.knot_row_count[k], lo = P.knot_lam_off[k];
        if (rc == 0) return;
        const double* lamk = lam() + lo;
        const double* muk = mu() + lo;
        double xr[n], ur[m];
#pragma unroll
        for (int i = 0; i < n; i++) xr[i] = x[i];
#pragma unroll
        for (int i = 0; i < m; i++) ur[i] = u[i];
        // second-order pieces: lanes own elements of [xx | uu | ux]
        const int nel = term ? n * n : (n * n + m * m + m * n);
        for (int e = lane; e < nel; e += 32) {
            int ci, cj;
            double* dst;
            if (e < n * n) { ci = e % n; cj = e / n; dst = &sm.Qxx[e]; }
            else if (e < n * n + m * m) { const int f = e - n * n; ci = n + f % m; cj = n + f / m; dst = &sm.Quu[f]; }
            else { const int f = e - n * n - m * m; ci = n + f % m; cj = f / m; dst = &sm.Qux[f]; }
            double acc = 0.0;
            for (int i = 0; i < rc; i++) {
                const DevRow r = P.rows[rb + i];
                const double gi = row_jac<C>(r, xr, ci), gj = row_jac<C>(r, xr, cj);
                if (gi != 0.0 && gj != 0.0) {
                    const double c = row_value<C>(r, xr, ur);
                    const bool act = r.eq ? true : ((c >= 0.0) || (lamk[i] > 0.0));
                    const double im = act ? muk[i] : 0.0;
                    acc = fma(gi * im, gj, acc);
                }
            }
            *dst += acc;
        }
        // first-order pieces
        const int nv = term ? n : n + m;
        for (int e = lane; e < nv; e += 32) {
            double acc = 0.0;
            for (int i = 0; i < rc; i++) {
                const DevRow r = P.rows[rb + i];
                const double gi = row_jac<C>(r, xr, e);
                if (gi != 0.0) {
                    const double c = row_value<C>(r, xr, ur);
                    const bool act = r.eq ? true : ((c >= 0.0) || (lamk[i] > 0.0));
                    const double im = act ? muk[i] : 0.0;
                    const double g = im * c + lamk[i];
                    acc = fma(gi, g, acc);
                }
            }
            if (e < n) sm.Qx[e] += acc; else sm.Qu[e - n] += acc;
        }
        __syncwarp();
    }

    // ---- Quu_reg factorisations, replicated per lane in registers ----
    // isposdef(Hermitian(A)): upper Cholesky, left-looking dot-product form
    __device__ bool chol_pd(const double* Areg /* m*m col-major regs */) {
        constexpr int m = C::m;
        double Uc[m * m];
#pragma unroll
        for (int j = 0; j < m; j++) {
#pragma unroll
            for (int i = 0; i < j; i++) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < i; l++) acc = fma(Uc[i * m + l], Uc[j * m + l], acc);
                Uc[j * m + i] = (Areg[j * m + i] - acc) / Uc[i * m + i];
            }
            double acc = 0.0;
#pragma unroll
            for (int l = 0; l < j; l++) acc = fma(Uc[j * m + l], Uc[j * m + l], acc);
            double dd = Areg[j * m + j] - acc;
            if (!(dd > 0.0)) return false;
            Uc[j * m + j] = sqrt(dd);
        }
        return true;
    }

    // Julia's dense `\` (backward_pass.jl:66-67): triangular shortcuts, else LU with partial
    // pivoting (left-looking, reciprocal column scaling), forward + backward substitution.
    // A is factored in registers with static indexing; `solve` handles one right-hand side.
    struct LU {
        double a[C::m * C::m];
        int piv[C::m];
        bool tril_only, triu;
    };
    __device__ void lu_factor(LU& f) {
        constexpr int m = C::m;
        bool tril = true, triu = true;
#pragma unroll
        for (int j = 0; j < m; j++)
#pragma unroll
            for (int i = 0; i < m; i++) {
                if (i < j && f.a[j * m + i] != 0.0) tril = false;
                if (i > j && f.a[j * m + i] != 0.0) triu = false;
            }
        f.triu = triu;
        f.tril_only = tril && !triu;
#pragma unroll
        for (int j = 0; j < m; j++) f.piv[j] = j;
        if (triu || f.tril_only) return;
#pragma unroll
        for (int j = 0; j < m; j++) {
            // apply earlier interchanges to column j
#pragma unroll
            for (int i = 0; i < j; i++) {
#pragma unroll
                for (int q = i + 1; q < m; q++) {
                    const bool sw = (f.piv[i] == q);
                    const double t0 = f.a[j * m + i], t1 = f.a[j * m + q];
                    f.a[j * m + i] = sw ? t1 : t0;
                    f.a[j * m + q] = sw ? t0 : t1;
                }
            }
#pragma unroll
            for (int i = 1; i < j; i++) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < i; l++) acc = fma(f.a[l * m + i], f.a[j * m + l], acc);
                f.a[j * m + i] = f.a[j * m + i] - acc;
            }
#pragma unroll
            for (int i = j; i < m; i++) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < j; l++) acc = fma(f.a[l * m + i], f.a[j * m + l], acc);
                f.a[j * m + i] = f.a[j * m + i] - acc;
            }
            int p = j;
            double amax = fabs(f.a[j * m + j]);
#pragma unroll
            for (int i = j + 1; i < m; i++)
                if (fabs(f.a[j * m + i]) > amax) { amax = fabs(f.a[j * m + i]); p = i; }
            f.piv[j] = p;
#pragma unroll
            for (int q = j + 1; q < m; q++) {
                const bool sw = (p == q);
#pragma unroll
                for (int c = 0; c <= j; c++) {
                    const double t0 = f.a[c * m + j], t1 = f.a[c * m + q];
                    f.a[c * m + j] = sw ? t1 : t0;
                    f.a[c * m + q] = sw ? t0 : t1;
                }
            }
            const double rp = 1.0 / f.a[j * m + j];
#pragma unroll
            for (int i = j + 1; i < m; i++) f.a[j * m + i] = f.a[j * m + i] * rp;
        }
    }
    __device__ void lu_solve(const LU& f, double* bv /* m regs, in/out */) {
        constexpr int m = C::m;
        if (f.tril_only) {
#pragma unroll
            for (int i = 0; i < m; i++) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < i; l++) acc = fma(f.a[l * m + i], bv[l], acc);
                bv[i] = (bv[i] - acc) / f.a[i * m + i];
            }
            return;
        }
        if (!f.triu) {
#pragma unroll
            for (int i = 0; i < m; i++) {
#pragma unroll
                for (int q = i + 1; q < m; q++) {
                    const bool sw = (f.piv[i] == q);
                    const double t0 = bv[i], t1 = bv[q];
                    bv[i] = sw ? t1 : t0;
                    bv[q] = sw ? t0 : t1;
                }
            }
#pragma unroll
            for (int i = 1; i < m; i++) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < i; l++) acc = fma(f.a[l * m + i], bv[l], acc);
                bv[i] = bv[i] - acc;
            }
        }
#pragma unroll
        for (int i = m - 1; i >= 0; i--) {
            double acc = 0.0;
#pragma unroll
            for (int l = i + 1; l < m; l++) acc = fma(f.a[l * m + i], bv[l], acc);
            bv[i] = (bv[i] - acc) / f.a[i * m + i];
        }
    }

    // ---- backward pass (backward_pass.jl:9-85), with the in-place-accumulation restart quirk ----
    // false: PD test failed with a non-finite rho (TO_STATUS_REG_DIVERGED)
    __device__ bool backwardpass(double& dV0, double& dV1) {
        constexpr int n = C::n, m = C::m;
        const int N = P.N;
        bool store_mode = false;
        int stored_from = N - 1;
        double* qst = ws + L.QST;
        for (;;) {  // one execution of the reference's while-loop from k = N-1
            // terminal cost-to-go
            expansion(N - 1);
            for (int e = lane; e < n * n; e += 32) sm.S[e] = sm.Qxx[e];
            if (lane < n) sm.Sx[lane] = sm.Qx[lane];
            __syncwarp();
            dV0 = 0.0;
            dV1 = 0.0;
            bool failed = false;
            for (int k = N - 2; k >= 0; k--) {
                load_AB(k);
                if (store_mode && k >= stored_from) {
                    const double* q = qst + (size_t)k * C::QS;
                    for (int e = lane; e < C::QS; e += 32) {
                        const double v = q[e];
                        if (e < n) sm.Qx[e] = v;
                        else if (e < n + m) sm.Qu[e - n] = v;
                        else if (e < n + m + n * n) sm.Qxx[e - n - m] = v;
                        else if (e < n + m + n * n + m * m) sm.Quu[e - n - m - n * n] = v;
                        else sm.Qux[e - n - m - n * n - m * m] = v;
                    }
                    __syncwarp();
                } else {
                    expansion(k);  // ends with __syncwarp
                }
                // Qx += A'Sx ; Qu += B'Sx ; T = A'S ; Tu = B'S
                for (int e = lane; e < n + m; e += 32) {
                    double acc = 0.0;
                    if (e < n) {
#pragma unroll
                        for (int l = 0; l < n; l++) acc = fma(sm.A[e * n + l], sm.Sx[l], acc);
                        sm.Qx[e] += acc;
                    } else {
                        const int i = e - n;
#pragma unroll
                        for (int l = 0; l < n; l++) acc = fma(sm.B[i * n + l], sm.Sx[l], acc);
                        sm.Qu[i] += acc;
                    }
                }
                for (int e = lane; e < n * n + m * n; e += 32) {
                    double acc = 0.0;
                    if (e < n * n) {
                        const int i = e % n, j = e / n;
#pragma unroll
                        for (int l = 0; l < n; l++) acc = fma(sm.A[i * n + l], sm.S[j * n + l], acc);
                        sm.T[e] = acc;
                    } else {
                        const int f = e - n * n, i = f % m, j = f / m;
#pragma unroll
                        for (int l = 0; l < n; l++) acc = fma(sm.B[i * n + l], sm.S[j * n + l], acc);
                        sm.Tu[f] = acc;
                    }
                }
                __syncwarp();
                // Qxx += T*A ; Quu += Tu*B ; Qux += Tu*A
                for (int e = lane; e < n * n + m * m + m * n; e += 32) {
                    double acc = 0.0;
                    if (e < n * n) {
                        const int i = e % n, j = e / n;
#pragma unroll
                        for (int l = 0; l < n; l++) acc = fma(sm.T[l * n + i], sm.A[j * n + l], acc);
                        sm.Qxx[e] += acc;
                    } else if (e < n * n + m * m) {
                        const int f = e - n * n, i = f % m, j = f / m;
#pragma unroll
                        for (int l = 0; l < n; l++) acc = fma(sm.Tu[l * m + i], sm.B[j * n + l], acc);
                        sm.Quu[f] += acc;
                    } else {
                        const int f = e - n * n - m * m, i = f % m, j = f / m;
#pragma unroll
                        for (int l = 0; l < n; l++) acc = fma(sm.Tu[l * m + i], sm.A[j * n + l], acc);
                        sm.Qux[f] += acc;
                    }
                }
                __syncwarp();
                if (store_mode) {
                    double* q = qst + (size_t)k * C::QS;
                    for (int e = lane; e < C::QS; e += 32) {
                        double v;
                        if (e < n) v = sm.Qx[e];
                        else if (e < n + m) v = sm.Qu[e - n];
                        else if (e < n + m + n * n) v = sm.Qxx[e - n - m];
                        else if (e < n + m + n * n + m * m) v = sm.Quu[e - n - m - n * n];
                        else v = sm.Qux[e - n - m - n * n - m * m];
                        q[e] = v;
                    }
                    if (k < stored_from) stored_from = k;
                }
                // Quu_reg = Quu + ρI (:control) or Quu + (ρB')B with Qux_reg = Qux + (ρB')A (:state, backward_pass.jl:38-46),
                // replicated in registers
                const bool reg_state = (io.bp_reg_type == TO_REG_STATE);
                LU f;
                if (reg_state) {
#pragma unroll
                    for (int e = 0; e < m * m; e++) {
                        const int i = e % m, j = e / m;
                        double acc = 0.0;
#pragma unroll
                        for (int l = 0; l < n; l++) acc = fma(rho * sm.B[i * n + l], sm.B[j * n + l], acc);
                        f.a[e] = sm.Quu[e] + acc;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < m * m; e++) f.a[e] = sm.Quu[e];
#pragma unroll
                    for (int i = 0; i < m; i++) f.a[i * m + i] = sm.Quu[i * m + i] + rho * 1.0;
                }
                if (!chol_pd(f.a)) { failed = true; break; }
                lu_factor(f);
                // gains: lane c < n solves for column c of K, lane n for d
                if (lane <= n) {
                    double rhs[m];
#pragma unroll
                    for (int i = 0; i < m; i++) rhs[i] = (lane < n) ? sm.Qux[lane * m + i] : sm.Qu[i];
                    if (reg_state && lane < n) {
#pragma unroll
                        for (int i = 0; i < m; i++) {
                            double acc = 0.0;
#pragma unroll
                            for (int l = 0; l < n; l++) acc = fma(rho * sm.B[i * n + l], sm.A[lane * n + l], acc);
                            rhs[i] = sm.Qux[lane * m + i] + acc;
                        }
                    }
                    lu_solve(f, rhs);
#pragma unroll
                    for (int i = 0; i < m; i++) {
                        const double v = -1.0 * rhs[i];
                        if (lane < n) sm.K[lane * m + i] = v; else sm.d[i] = v;
                    }
                }
                __syncwarp();
                {   // publish K,d for the rollouts
                    double* kd = KD(k);
                    for (int e = lane; e < C::KDS; e += 32) kd[e] = (e < m * n) ? sm.K[e] : sm.d[e - m * n];
                }
                // KQ = K'Quu (n×m)
                for (int e = lane; e < n * m; e += 32) {
                    const int i = e % n, j = e / n;
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < m; l++) acc = fma(sm.K[i * m + l], sm.Quu[j * m + l], acc);
                    sm.KQ[e] = acc;
                }
                __syncwarp();
                // S.x = Qx + KQ d + K'Qu + Qux'd ; unsymmetrised S.xx into T
                for (int e = lane; e < n * n + n; e += 32) {
                    if (e < n * n) {
                        const int i = e % n, j = e / n;
                        double a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                        for (int l = 0; l < m; l++) a1 = fma(sm.KQ[l * n + i], sm.K[j * m + l], a1);
#pragma unroll
                        for (int l = 0; l < m; l++) a2 = fma(sm.K[i * m + l], sm.Qux[j * m + l], a2);
#pragma unroll
                        for (int l = 0; l < m; l++) a3 = fma(sm.Qux[i * m + l], sm.K[j * m + l], a3);
                        sm.T[e] = ((sm.Qxx[e] + a1) + a2) + a3;
                    } else {
                        const int i = e - n * n;
                        double a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                        for (int l = 0; l < m; l++) a1 = fma(sm.KQ[l * n + i], sm.d[l], a1);
#pragma unroll
                        for (int l = 0; l < m; l++) a2 = fma(sm.K[i * m + l], sm.Qu[l], a2);
#pragma unroll
                        for (int l = 0; l < m; l++) a3 = fma(sm.Qux[i * m + l], sm.d[l], a3);
                        sm.Sx[i] = ((sm.Qx[i] + a1) + a2) + a3;
                    }
                }
                __syncwarp();
                for (int e = lane; e < n * n; e += 32) {
                    const int i = e % n, j = e / n;
                    sm.S[e] = 0.5 * (sm.T[j * n + i] + sm.T[i * n + j]);
                }
                // ΔV (replicated)
                {
                    double a = 0.0;
#pragma unroll
                    for (int l = 0; l < m; l++) a = fma(sm.d[l], sm.Qu[l], a);
                    dV0 += a;
                    double acc = 0.0;
#pragma unroll
                    for (int j = 0; j < m; j++) {
                        double w = 0.0;
#pragma unroll
                        for (int l = 0; l < m; l++) w = fma(0.5 * sm.d[l], sm.Quu[j * m + l], w);
                        acc = fma(w, sm.d[j], acc);
                    }
                    dV1 += acc;
                }
                __syncwarp();
            }
            if (!failed) break;
            if (!store_mode) {
                // first failure of this backward pass: replay the pass while materialising the
                // accumulated Q trajectory, so that the restart sees the reference's in-place state (Q1)
                store_mode = true;
                stored_from = N - 1;
                continue;
            }
            if (!isfinite(rho)) return false;
            reg_update(true);
        }
        reg_update(false);
        return true;
    }

    // ---- forward pass: all step sizes in parallel (forward_pass.jl:5-85, rollout.jl:2-23) ----
    // returns J; sets `err` when the reference would throw "Cost increased during Forward Pass"
    __device__ double forwardpass(double dV0, double dV1, double J_prev, bool& err) {
        constexpr int n = C::n, m = C::m;
        constexpr int SS = n + m + m * n + m;  // staged doubles per knot: X[k], U[k], K[k], d[k]
        const int N = P.N;
        const int ntrial = io.iterations_linesearch + 1;  // trials iter = 0..iterations_linesearch
        const double alpha = __longlong_as_double((long long)(1023 - lane) << 52);  // 2^-lane
        double xb[n], ub[m];
#pragma unroll
        for (int i = 0; i < n; i++) xb[i] = x0[i];
        bool ok = true;
        double J = 0.0, Jc = 0.0;
        double* XB = ws + L.XB;
        double* UB = ws + L.UB;
        // prefetch registers for the staging buffer
        constexpr int NLD = (SS + 31) / 32;
        double pre[NLD];
        auto fetch = [&](int k) {
#pragma unroll
            for (int q = 0; q < NLD; q++) {
                const int e = lane + 32 * q;
                double v = 0.0;
                if (e < n) v = X(k)[e];
                else if (e < n + m) v = U(k)[e - n];
                else if (e < SS) v = KD(k)[e - n - m];
                pre[q] = v;
            }
        };
        fetch(0);
        for (int k = 0; k < N - 1; k++) {
            double* st = sm.stage[k & 1];
#pragma unroll
            for (int q = 0; q < NLD; q++) {
                const int e = lane + 32 * q;
                if (e < SS) st[e] = pre[q];
            }
            __syncwarp();
            if (k + 1 < N - 1) fetch(k + 1);
            const double* Xk = st;
            const double* Uk = st + n;
            const double* Kk = st + n + m;
            const double* dk = st + n + m + m * n;
            double dx[n];
#pragma unroll
            for (int i = 0; i < n; i++) dx[i] = xb[i] - Xk[i];
#pragma unroll
            for (int i = 0; i < m; i++) {
                double acc = 0.0;
#pragma unroll
                for (int j = 0; j < n; j++) acc = fma(Kk[j * m + i], dx[j], acc);
                ub[i] = (Uk[i] + acc) + alpha * dk[i];
            }
            // cost of knot k at (x̄_k, ū_k)
            J += stage_cost<C>(P, xb, ub);
            if (al_on) Jc += knot_al_cost<C>(P, k, lam(), mu(), xb, ub);
            if (lane < ntrial) {
#pragma unroll
                for (int i = 0; i < n; i++) XB[((size_t)k * n + i) * 32 + lane] = xb[i];
#pragma unroll
                for (int i = 0; i < m; i++) UB[((size_t)k * m + i) * 32 + lane] = ub[i];
            }
            double xn[n];
            dyn_eval<C>(P, xn, xb, ub);
            double mx = 0.0, mu_ = 0.0;
            bool bad = false;
#pragma unroll
            for (int i = 0; i < n; i++) { const double a = fabs(xn[i]); if (a != a) bad = true; mx = dmax(mx, a); }
#pragma unroll
            for (int i = 0; i < m; i++) { const double a = fabs(ub[i]); if (a != a) bad = true; mu_ = dmax(mu_, a); }
            if (bad || !(mx < io.max_state_value && mu_ < io.max_control_value)) ok = false;
#pragma unroll
            for (int i = 0; i < n; i++) xb[i] = xn[i];
        }
        {
            double uz[m];
#pragma unroll
            for (int i = 0; i < m; i++) uz[i] = 0.0;
            J += term_cost<C>(P, xb);
            if (al_on) Jc += knot_al_cost<C>(P, N - 1, lam(), mu(), xb, uz);
            if (lane < ntrial) {
#pragma unroll
                for (int i = 0; i < n; i++) XB[((size_t)(N - 1) * n + i) * 32 + lane] = xb[i];
            }
        }
        double Jt = al_on ? (J + Jc) : J;
        const double expected = -alpha * (dV0 + alpha * dV1);
        double z = (expected > 0) ? (J_prev - Jt) / expected : -1.0;
        const bool cont = (z <= io.line_search_lower_bound || z > io.line_search_upper_bound) && (Jt >= J_prev);
        const bool accept = ok && !cont && (lane < ntrial);
        const unsigned msk = __ballot_sync(FULL, accept);
        __syncwarp();
        double Jres;
        if (msk != 0) {
            const int w = __ffs(msk) - 1;
            ls_count += (unsigned long long)(w + 1);
            Jres = bcast(Jt, w);
            fp_expected = bcast(expected, w);
            fp_z = bcast(z, w);
            fp_alpha = bcast(alpha, w);
            err = (Jres > J_prev);
            if (!err && !(Jres > io.max_cost_value)) {
                // X ← X̄, U ← Ū of the accepted trial (ilqr_methods.jl:30-33)
                const int nx = N * n, nu = (N - 1) * m;
                for (int e = lane; e < nx; e += 32) ws[L.X + e] = XB[(size_t)e * 32 + w];
                for (int e = lane; e < nu; e += 32) ws[L.U + e] = UB[(size_t)e * 32 + w];
                __syncwarp();
            }
        } else {
            // line search failed (forward_pass.jl:22-37): X̄ ← X, Ū ← U, J recomputed, regularisation bumped
            ls_count += (unsigned long long)ntrial;
            Jres = eval_cost();
            fp_expected = 0.0;
            fp_z = 0.0;
            fp_alpha = 0.0;
            reg_update(true);
            rho += io.bp_reg_fp;
            err = (Jres > J_prev);
        }
        return Jres;
    }

    // ---- bookkeeping (ilqr_methods.jl:77-89, 139-162) ----
    __device__ void record_inner(double J, double dJ) {
        iterations += 1;
        last_cost = J;
        last_dJ = dJ;
        last_grad = gradient();
        if (dJ == 0.0) dJ_zero += 1; else dJ_zero = 0;
        if (Bt.inner_cap > 0) {
            if (n_inner_rec < Bt.inner_cap) {
                if (lane == 0) {
                    TOIterRecord r;
                    r.cost = J; r.dJ = dJ; r.gradient = last_grad; r.expected = fp_expected; r.z = fp_z;
                    r.alpha = fp_alpha; r.rho = rho; r.outer = outer_idx; r.iter = iterations;
                    Bt.inner[(size_t)b * Bt.inner_cap + n_inner_rec] = r;
                }
                n_inner_rec++;
            } else {
                status |= TO_STATUS_TRACE_TRUNC;
            }
        }
    }
    __device__ bool inner_converged() {
        if (0.0 < last_dJ && last_dJ < io.cost_tolerance) return true;
        if (last_grad < io.gradient_norm_tolerance) return true;
        if (iterations >= io.iterations) return true;
        if (dJ_zero > io.dJ_counter_limit) return true;
        return false;
    }

    __device__ bool all_finite_X() {
        const int nx = P.N * C::n;
        bool fin = true;
        for (int e = lane; e < nx; e += 32) if (!isfinite(ws[L.X + e])) fin = false;
        return __all_sync(FULL, fin);
    }
    // open-loop rollout from x0 (rollout.jl:25-38); optionally with the divergence guard of projection!
    __device__ void rollout_open(bool guard) {
        if (lane == 0) {
            double x[C::n], u[C::m], xn[C::n];
#pragma unroll
            for (int i = 0; i < C::n; i++) { x[i] = x0[i]; X(0)[i] = x0[i]; }
            bool stop = false;
            for (int k = 0; k < P.N - 1; k++) {
                if (stop) {
#pragma unroll
                    for (int i = 0; i < C::n; i++) X(k + 1)[i] = 0.0;
                    continue;
                }
#pragma unroll
                for (int i = 0; i < C::m; i++) u[i] = U(k)[i];
                dyn_eval<C>(P, xn, x, u);
                double mx = 0.0, mu_ = 0.0;
                bool bad = false;
#pragma unroll
                for (int i = 0; i < C::n; i++) { X(k + 1)[i] = xn[i]; x[i] = xn[i]; const double a = fabs(xn[i]); if (a != a) bad = true; mx = dmax(mx, a); }
#pragma unroll
                for (int i = 0; i < C::m; i++) { const double a = fabs(u[i]); if (a != a) bad = true; mu_ = dmax(mu_, a); }
                if (guard && (bad || !(mx < io.max_state_value && mu_ < io.max_control_value))) stop = true;
            }
        }
        __syncwarp();
    }

    // ---- iLQR solve (ilqr_methods.jl:3-45).  false = abort the whole solve ----
    __device__ bool ilqr_solve() {
        iterations = 0; dJ_zero = 0; rho = 0.0; drho = 0.0;
        fp_expected = 0.0; fp_z = 0.0; fp_alpha = 0.0;
        if (!all_finite_X()) rollout_open(false);
        double J_prev = eval_cost();
        record_inner(J_prev, __longlong_as_double(0x7ff0000000000000LL));
        for (int i = 1; i <= io.iterations; i++) {
            jacobians();
            double dV0, dV1;
            if (!backwardpass(dV0, dV1)) { status |= TO_STATUS_REG_DIVERGED; return false; }
            if (ctl.debug && b == 0 && ctl.debug_flag && *ctl.debug_flag == 0 && steps == 0) debug_dump(dV0, dV1, J_prev);
            bool err = false;
            const double J = forwardpass(dV0, dV1, J_prev, err);
            steps += 1;
            if (err) { status |= TO_STATUS_COST_INCREASED; return false; }
            if (J > io.max_cost_value) { status |= TO_STATUS_COST_BLOWUP; return true; }
            const double dJ = fabs(J - J_prev);
            J_prev = J;
            record_inner(J, dJ);
            if (inner_converged()) break;
        }
        return true;
    }

    __device__ void debug_dump(double dV0, double dV1, double J0) {
        // [0]=J0 [1]=dV0 [2]=dV1, then Z (all knots), then K,d (all knots)
        double* dbg = ctl.debug;
        const int N = P.N;
        if (lane == 0) { dbg[0] = J0; dbg[1] = dV0; dbg[2] = dV1; }
        const size_t nz = (size_t)(N - 1) * C::ZS, nk = (size_t)(N - 1) * C::KDS;
        for (size_t e = lane; e < nz; e += 32) dbg[3 + e] = ws[L.Z + e];
        for (size_t e = lane; e < nk; e += 32) dbg[3 + nz + e] = ws[L.KD + e];
        __syncwarp();
    }

    __device__ void record_outer(double J, double cmax, int& al_iterations, int& al_total) {
        al_iterations += 1;
        al_total += iterations;
        if (Bt.outer_cap > 0) {
            if (n_outer_rec < Bt.outer_cap) {
                // max_penalty: maximum(maximum(μ)) over a vector of vectors (lexicographic max, Q20)
                if (lane == 0) {
                    int best = -1;
                    for (int k = 0; k < P.N; k++) {
                        if (best < 0) { best = k; continue; }
                        const int ca = P.knot_row_count[best], cb_ = P.knot_row_count[k];
                        const double* a = mu() + P.knot_lam_off[best];
                        const double* c = mu() + P.knot_lam_off[k];
                        bool less = false, decided = false;
                        for (int i = 0; i < ca && i < cb_; i++) {
                            if (a[i] < c[i]) { less = true; decided = true; break; }
                            if (c[i] < a[i]) { decided = true; break; }
                        }
                        if (!decided) less = ca < cb_;
                        if (less) best = k;
                    }
                    double pm = 0.0;
                    if (best >= 0 && P.knot_row_count[best] > 0) {
                        const double* a = mu() + P.knot_lam_off[best];
                        pm = a[0];
                        for (int i = 1; i < P.knot_row_count[best]; i++) pm = dmax(pm, a[i]);
                    }
                    TOOuterRecord r;
                    r.cost = J; r.c_max = cmax; r.penalty_max = pm; r.iterations_inner = iterations; r.pad = 0;
                    Bt.outer[(size_t)b * Bt.outer_cap + n_outer_rec] = r;
                }
                n_outer_rec++;
            } else {
                status |= TO_STATUS_TRACE_TRUNC;
            }
        }
    }

    // ---- AL solve (augmented_lagrangian_methods.jl:2-31) ----
    __device__ bool al_solve(double& Jout, double& cmax_out, int& al_iterations, int& al_total) {
        const TOALOptions& o = ctl.o;
        al_on = true;
        for (int e = lane; e < P.Ptot; e += 32) { lam()[e] = 0.0; mu()[e] = o.penalty_initial; }
        // fresh inner solver: K = d = 0 (ilqr_solver.jl:125-140)
        {
            const size_t nk = (size_t)(P.N - 1) * C::KDS;
            for (size_t e = lane; e < nk; e += 32) ws[L.KD + e] = 0.0;
        }
        __syncwarp();
        iterations = 0; dJ_zero = 0; rho = 0.0; drho = 0.0;
        al_iterations = 0; al_total = 0;
        outer_idx = 0;
        if (!all_finite_X()) rollout_open(false);
        double J0 = eval_cost();
        double cmax = max_violation();
        record_outer(J0, cmax, al_iterations, al_total);
        Jout = J0; cmax_out = cmax;
        bool converged = false;
        for (int i = 1; i <= o.iterations; i++) {
            outer_idx = i - 1;
            if (i != o.iterations) {
                io.cost_tolerance = o.cost_tolerance_intermediate;
                io.gradient_norm_tolerance = o.gradient_norm_tolerance_intermediate;
            } else {
                io.cost_tolerance = o.cost_tolerance;
                io.gradient_norm_tolerance = o.gradient_norm_tolerance;
            }
            if (!ilqr_solve()) return false;
            const double J = eval_cost();
            cmax = max_violation();  // C is a function of (X,U) only: same values the dual update sees
            dual_penalty_update();
            record_outer(J, cmax, al_iterations, al_total);
            Jout = J; cmax_out = cmax;
            converged = false;
            if (o.kickout_max_penalty) {
                // same lexicographic maximum as record_outer; evaluated by lane 0 and broadcast
                double pm = 0.0;
                if (lane == 0) {
                    int best = -1;
                    for (int k = 0; k < P.N; k++) {
                        if (best < 0) { best = k; continue; }
                        const int ca = P.knot_row_count[best], cb_ = P.knot_row_count[k];
                        const double* a = mu() + P.knot_lam_off[best];
                        const double* c = mu() + P.knot_lam_off[k];
                        bool less = false, decided = false;
                        for (int q = 0; q < ca && q < cb_; q++) {
                            if (a[q] < c[q]) { less = true; decided = true; break; }
                            if (c[q] < a[q]) { decided = true; break; }
                        }
                        if (!decided) less = ca < cb_;
                        if (less) best = k;
                    }
                    if (best >= 0 && P.knot_row_count[best] > 0) {
                        const double* a = mu() + P.knot_lam_off[best];
                        pm = a[0];
                        for (int q = 1; q < P.knot_row_count[best]; q++) pm = dmax(pm, a[q]);
                    }
                }
                pm = bcast(pm, 0);
                converged = (pm == o.penalty_max);
            }
            converged = converged || (cmax < o.constraint_tolerance);
            if (converged) break;
            iterations = 0; dJ_zero = 0; rho = 0.0; drho = 0.0;
        }
        if (!converged) status |= TO_STATUS_MAX_OUTER;
        return true;
    }

    // ---- load one problem into the workspace (altro_methods.jl:98-124 initialisation included) ----
    __device__ void prologue(int b_) {
        constexpr int n = C::n, m = C::m, n0 = C::n0, m0 = C::m0;
        b = b_;
        const int N = P.N;
        io = ctl.o.opts_uncon;
        status = 0; steps = 0; n_inner_rec = 0; n_outer_rec = 0; ls_count = 0;
        last_cost = 0.0; last_dJ = 0.0; last_grad = 0.0;
        if (ctl.accumulate) {
            status = Bt.res[b].status;
            steps = Bt.res[b].steps;
            if (Bt.inner_cap > 0) n_inner_rec = Bt.n_inner[b];
            if (Bt.outer_cap > 0) n_outer_rec = Bt.n_outer[b];
        }
        const double nanv = __longlong_as_double(0x7ff8000000000000LL);
        // x0, X, U
        load_x0();
        const double sdt = sqrt(P.dt);
        for (int e = lane; e < N * n; e += 32) {
            const int k = e / n, i = e - k * n;
            double v;
            if (i < n0) v = Bt.X0 ? Bt.X0[((size_t)b * N + k) * n0 + i] : nanv;
            else v = sdt;  // minimum_time.jl:35
            ws[L.X + e] = v;
        }
        for (int e = lane; e < (N - 1) * m; e += 32) {
            const int k = e / m, i = e - k * m;
            double v = 0.0;
            if (i < m0) v = Bt.U0[((size_t)b * (N - 1) + k) * m0 + i];
            else if (C::MT && i == m - 1) v = sdt;  // minimum_time.jl:34
            ws[L.U + e] = v;
        }
        if constexpr (C::MT) {
            if (Bt.tau_in) {  // infeasible_to_feasible_problem with minimum time (infeasible.jl:43-51): sqrt(dt) of the previous solve
                for (int k = lane; k < N - 1; k += 32) ws[L.U + (size_t)k * m + (m - 1)] = Bt.tau_in[(size_t)b * (N - 1) + k];
                for (int k = lane; k < N; k += 32) ws[L.X + (size_t)k * n + (n - 1)] = (k == 0) ? 0.0 : Bt.xtau_in[(size_t)b * N + k];
            }
        }
        __syncwarp();
        if (C::INF && ctl.altro_init) {
            // slack_controls (infeasible.jl:62-80): u_s[k] = X0[k+1] - f(x̂[k],u[k]) on the original model
            if (lane == 0) {
                double x[n0], xn[n0], u[m0];
#pragma unroll
                for (int i = 0; i < n0; i++) x[i] = x0[i];
                for (int k = 0; k < N - 1; k++) {
#pragma unroll
                    for (int i = 0; i < m0; i++) u[i] = U(k)[i];
                    fd_model<C::MODEL, C::INTEG, double>(xn, x, u, P.dt);
#pragma unroll
                    for (int i = 0; i < n0; i++) {
                        const double us = X(k + 1)[i] - xn[i];
                        U(k)[m0 + i] = us;
                        x[i] = xn[i] + us;
                    }
                }
            }
            __syncwarp();
        }
        if (ctl.projection_first) rollout_open(true);
    }
    __device__ void load_x0() {
#pragma unroll
        for (int i = 0; i < C::n; i++) x0[i] = (i < C::n0) ? Bt.x0[(size_t)b * C::n0 + i] : 0.0;
    }

    // ---- write the result record, the solution and (AL) the multipliers / active set ----
    __device__ void epilogue(double Jout, double cmax, int al_it, int al_tot) {
        constexpr int n = C::n, m = C::m;
        const int N = P.N;
        if (lane == 0) {
            TOResult r;
            r.J = Jout; r.c_max = cmax; r.iterations_total = al_tot; r.iterations_outer = al_it;
            r.status = status; r.steps = steps;
            Bt.res[b] = r;
            if (Bt.inner_cap > 0) Bt.n_inner[b] = n_inner_rec;
            if (Bt.outer_cap > 0) Bt.n_outer[b] = n_outer_rec;
            if (ctl.ls_trials) atomicAdd(ctl.ls_trials, ls_count);
        }
        if (ctl.write_solution) {
            const int no = Bt.n_out, mo = Bt.m_out;
            for (int e = lane; e < N * no; e += 32) {
                const int k = e / no, i = e - k * no;
                Bt.X[(size_t)b * N * no + e] = ws[L.X + (size_t)k * n + i];
            }
            for (int e = lane; e < (N - 1) * mo; e += 32) {
                const int k = e / mo, i = e - k * mo;
                Bt.U[(size_t)b * (N - 1) * mo + e] = ws[L.U + (size_t)k * m + i];
            }
            if (Bt.dts)
                for (int k = lane; k < N - 1; k += 32) {
                    double dtk = P.dt;
                    if (C::MT) { const double h = U(k)[m - 1]; dtk = h * h; }
                    Bt.dts[(size_t)b * (N - 1) + k] = dtk;
                }
            if constexpr (C::MT) {
                if (Bt.tau_out) {
                    for (int k = lane; k < N - 1; k += 32) Bt.tau_out[(size_t)b * (N - 1) + k] = U(k)[m - 1];
                    for (int k = lane; k < N; k += 32) Bt.xtau_out[(size_t)b * N + k] = X(k)[n - 1];
                }
            }
        }
        if (al_on && Bt.lam_out) {
            for (int e = lane; e < P.Ptot; e += 32) {
                Bt.lam_out[(size_t)b * P.Ptot + e] = lam()[e];
                Bt.mu_out[(size_t)b * P.Ptot + e] = mu()[e];
            }
            // active set at the final (X,U,λ) (constraint_sets.jl:247-267)
            for (int k = 0; k < N; k++) {
                const int rb = P.knot_row_begin[k], rc = P.knot_row_count[k], lo = P.knot_lam_off[k];
                if (rc == 0) continue;
                double x[n], u[m];
#pragma unroll
                for (int i = 0; i < n; i++) x[i] = X(k)[i];
#pragma unroll
                for (int i = 0; i < m; i++) u[i] = (k < N - 1) ? U(k)[i] : 0.0;
                for (int i = lane; i < rc; i += 32) {
                    const DevRow r = P.rows[rb + i];
                    const double c = row_value<C>(r, x, u);
                    Bt.act_out[(size_t)b * P.Ptot + lo + i] = r.eq ? 1 : (((c >= 0.0) || (lam()[lo + i] > 0.0)) ? 1 : 0);
                }
            }
        }
        __syncwarp();
    }

    // ---- load one problem into the workspace, solve, write results ----
    __device__ void run(int b_) {
        prologue(b_);
        const int N = P.N;
        al_on = false;
        double Jout = 0.0, cmax = 0.0;
        int al_it = 0, al_tot = 0;
        bool ok;
        if (ctl.mode == 0) {
            const size_t nk = (size_t)(N - 1) * C::KDS;
            for (size_t e = lane; e < nk; e += 32) ws[L.KD + e] = 0.0;
            __syncwarp();
            outer_idx = 0;
            ok = ilqr_solve();
            Jout = last_cost;
            al_tot = iterations;
        } else {
            ok = al_solve(Jout, cmax, al_it, al_tot);
        }
        (void)ok;
        epilogue(Jout, cmax, al_it, al_tot);
    }
};

// ------------------------------------------------------------------------------------------
// the persistent kernel: one warp per CTA, warps pull problems from the queue
// ------------------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(32) solve_kernel(const DevProblem P, const DevBatch Bt, const DevCtl ctl) {
    __shared__ Smem<C> sm;
    const int lane = threadIdx.x;
    double* ws = ctl.ws + (size_t)blockIdx.x * ctl.ws_stride;
    Solver<C> s(P, Bt, ctl, sm, ws, lane);
    for (;;) {
        unsigned int b = 0;
        if (lane == 0) b = atomicAdd(ctl.queue, 1u);
        b = __shfl_sync(FULL, b, 0);
        if (b >= (unsigned)Bt.B) break;
        s.run((int)b);
    }
}

template <class C> unsigned long long ws_doubles_fn(int N, int Ptot) { return ws_layout<C>(N, Ptot).total; }
template <class C> size_t debug_doubles_fn(int N) { return 3 + (size_t)(N - 1) * (C::ZS + C::KDS); }
template <class C> int max_blocks_fn() {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, solve_kernel<C>, 32, 0);
    return nb;
}
template <class C> void launch_fn(int grid, cudaStream_t st, const DevProblem& P, const DevBatch& B, const DevCtl& c) {
    solve_kernel<C><<<grid, 32, 0, st>>>(P, B, c);
}
}  // namespace tob
