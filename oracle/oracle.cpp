// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing in the product path may include, link or call this;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// A CPU restatement (plain C++17, double precision) of the iLQR / AL-iLQR / ALTRO hot path of
// TrajectoryOptimization.jl v0.1.1.  The reference is pure Julia and cannot run in this image
// (no julia binary), so this restatement is the arbiter for the CUDA engine.  It is pinned
// against the reference's own known-answer vectors (test/constraint_tests.jl:91-108,190-202,
// test/test_utils.jl:3-5,82-94, test/cost_tests.jl:64-89) and against the two iteration traces
// embedded in examples/acrobot/Acrobot.ipynb cell 19 and examples/car/Car Escape.ipynb cell 23
// (tests/test_oracle_golden.py).  Quantities no reference artefact pins (regularisation
// restarts, square-root backward pass in a full solve, minimum time) are "parity unpinned".
//
// Every function cites the reference file:line it follows (paths relative to /root/reference).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../include/trajopt_b200.h"
#include "oracle_models.hpp"

namespace orc {

typedef std::vector<double> vec;

// ------------------------------------------------------------------------------------------
// internal (possibly augmented) problem
// ------------------------------------------------------------------------------------------
enum RowKind { R_LIN = 0, R_CIRCLE = 1, R_SPHERE = 2, R_MTEQ = 3 };
struct Row {
    int kind;
    bool eq;
    int col;  // R_LIN: column in z̄ = [x̄; ū] (terminal: in x̄)
    double sign, a, b, c, r;
};

struct Spec {
    int model, integ;
    int n0, m0;        // base model dims
    bool inf, mt;      // slack controls / minimum time
    int nq, mq;        // dims of the wrapped QuadraticCost: nq = n0, mq = m0 + (inf ? n0 : 0)
    int n, m, N;       // dims of the solved problem
    double dt;
    vec Q, R, H, q, r; // nq*nq, mq*mq, mq*nq, nq, mq
    double c;
    vec Qf, qf;
    double cf;
    double R_mt;
    std::vector<std::vector<Row>> rows;  // per knot (N entries); rows[N-1] is the terminal set
};

// ---- discrete dynamics of the augmented model --------------------------------------------
// src/model.jl:761-779 (add_slack_controls), src/solvers/altro/minimum_time.jl:85-104
static void dyn_eval(const Spec& S, double* xn, const double* x, const double* u) {
    double dt = S.dt;
    if (S.mt) {
        double h = u[S.m - 1];
        dt = h * h;
    }
    fd_model<double>(S.model, S.integ, S.n0, xn, x, u, dt);
    if (S.inf)
        for (int i = 0; i < S.n0; i++) xn[i] = xn[i] + u[S.m0 + i];
    if (S.mt) xn[S.n - 1] = u[S.m - 1];
}

// Jacobian of the inner discrete model w.r.t. [x;u;dt] via duals (src/model.jl:491-512),
// then scattered as add_slack_controls / add_min_time_controls do.  A: n×n, B: n×m col-major.
template <int P>
static void dyn_jac_inner(const Spec& S, const double* x, const double* u, double dt, double* Z /* n0×P col-major */) {
    const int n0 = S.n0, m0 = S.m0;
    Dual<P> xs[16], us[8], dts, xn[16];
    for (int i = 0; i < n0; i++) { xs[i] = Dual<P>(x[i]); xs[i].p[i] = 1.0; }
    for (int i = 0; i < m0; i++) { us[i] = Dual<P>(u[i]); us[i].p[n0 + i] = 1.0; }
    dts = Dual<P>(dt);
    dts.p[n0 + m0] = 1.0;
    fd_model<Dual<P>>(S.model, S.integ, n0, xn, xs, us, dts);
    for (int j = 0; j < P; j++)
        for (int i = 0; i < n0; i++) Z[j * n0 + i] = xn[i].p[j];
}

static void dyn_jac(const Spec& S, const double* x, const double* u, double* A, double* B) {
    const int n0 = S.n0, m0 = S.m0, n = S.n, m = S.m;
    double dt = S.dt, h = 0.0;
    if (S.mt) { h = u[m - 1]; dt = h * h; }
    double Z[16 * 24];
    const int P = n0 + m0 + 1;
    switch (P) {
        case 4: dyn_jac_inner<4>(S, x, u, dt, Z); break;
        case 6: dyn_jac_inner<6>(S, x, u, dt, Z); break;
        case 7: dyn_jac_inner<7>(S, x, u, dt, Z); break;
        case 18: dyn_jac_inner<18>(S, x, u, dt, Z); break;
        default: std::fprintf(stderr, "oracle: unsupported P=%d\n", P); std::abort();
    }
    for (int i = 0; i < n * n; i++) A[i] = 0.0;
    for (int i = 0; i < n * m; i++) B[i] = 0.0;
    for (int j = 0; j < n0; j++)
        for (int i = 0; i < n0; i++) A[j * n + i] = Z[j * n0 + i];
    for (int j = 0; j < m0; j++)
        for (int i = 0; i < n0; i++) B[j * n + i] = Z[(n0 + j) * n0 + i];
    if (S.inf)
        for (int i = 0; i < n0; i++) B[(m0 + i) * n + i] = 1.0;
    if (S.mt) {
        for (int i = 0; i < n0; i++) B[(m - 1) * n + i] = Z[(n0 + m0) * n0 + i] * (2.0 * h);
        B[(m - 1) * n + (n - 1)] = 1.0;
    }
}

// ---- cost --------------------------------------------------------------------------------
static inline double dotc(const double* a, const double* b, int len) {
    double acc = 0.0;
    for (int i = 0; i < len; i++) acc = fmad(a[i], b[i], acc);
    return acc;
}

// src/cost.jl:171-173 : 0.5*x'Q*x + 0.5*u'R*u + q'x + r'u + c + u'H*x   (row-vector*matrix, then dot)
static double quad_stage(const Spec& S, const double* x, const double* u) {
    const int n = S.nq, m = S.mq;
    double t1 = 0.0, t2 = 0.0, t6 = 0.0;
    {
        double acc = 0.0;
        for (int j = 0; j < n; j++) {
            double w = 0.0;
            for (int i = 0; i < n; i++) w = fmad(0.5 * x[i], S.Q[j * n + i], w);
            acc = fmad(w, x[j], acc);
        }
        t1 = acc;
    }
    {
        double acc = 0.0;
        for (int j = 0; j < m; j++) {
            double w = 0.0;
            for (int i = 0; i < m; i++) w = fmad(0.5 * u[i], S.R[j * m + i], w);
            acc = fmad(w, u[j], acc);
        }
        t2 = acc;
    }
    double t3 = dotc(S.q.data(), x, n);
    double t4 = dotc(S.r.data(), u, m);
    {
        double acc = 0.0;
        for (int j = 0; j < n; j++) {
            double w = 0.0;
            for (int i = 0; i < m; i++) w = fmad(u[i], S.H[j * m + i], w);
            acc = fmad(w, x[j], acc);
        }
        t6 = acc;
    }
    return ((((t1 + t2) + t3) + t4) + S.c) + t6;
}

// stage cost incl. dt scaling (src/cost.jl:175-177) and MinTimeCost (minimum_time.jl:155)
static double stage_cost(const Spec& S, const double* x, const double* u) {
    if (S.mt) {
        double tau = u[S.m - 1];
        double dt = tau * tau;
        return quad_stage(S, x, u) * dt + S.R_mt * (tau * tau);
    }
    return quad_stage(S, x, u) * S.dt;
}
// src/cost.jl:179-181 (terminal); MinTimeCost terminal ignores the extra state (minimum_time.jl:156)
static double term_cost(const Spec& S, const double* x) {
    const int n = S.nq;
    double acc = 0.0;
    for (int j = 0; j < n; j++) {
        double w = 0.0;
        for (int i = 0; i < n; i++) w = fmad(0.5 * x[i], S.Qf[j * n + i], w);
        acc = fmad(w, x[j], acc);
    }
    return (acc + dotc(S.qf.data(), x, n)) + S.cf;
}

struct Exp {  // Expansion (src/cost.jl:21-36), col-major; ux is m×n
    vec x, u, xx, uu, ux;
    void init(int n, int m) { x.assign(n, 0); u.assign(m, 0); xx.assign(n * n, 0); uu.assign(m * m, 0); ux.assign(m * n, 0); }
    void reset() { std::fill(x.begin(), x.end(), 0); std::fill(u.begin(), u.end(), 0); std::fill(xx.begin(), xx.end(), 0);
                   std::fill(uu.begin(), uu.end(), 0); std::fill(ux.begin(), ux.end(), 0); }
};

// src/cost.jl:183-192 (QuadraticCost) and minimum_time.jl:161-191 (MinTimeCost)
static void stage_expansion(const Spec& S, Exp& E, const double* x, const double* u) {
    const int nq = S.nq, mq = S.mq, n = S.n, m = S.m;
    vec Qx(nq), Qu(mq);
    for (int i = 0; i < nq; i++) {
        double a = 0.0, b = 0.0;
        for (int j = 0; j < nq; j++) a = fmad(S.Q[j * nq + i], x[j], a);
        for (int j = 0; j < mq; j++) b = fmad(S.H[i * mq + j], u[j], b);  // H'u
        Qx[i] = (a + S.q[i]) + b;
    }
    for (int i = 0; i < mq; i++) {
        double a = 0.0, b = 0.0;
        for (int j = 0; j < mq; j++) a = fmad(S.R[j * mq + i], u[j], a);
        for (int j = 0; j < nq; j++) b = fmad(S.H[j * mq + i], x[j], b);  // H x
        Qu[i] = (a + S.r[i]) + b;
    }
    double dt = S.dt, tau = 0.0;
    if (S.mt) { tau = u[m - 1]; dt = tau * tau; }
    for (int i = 0; i < nq; i++) E.x[i] = Qx[i] * dt;
    for (int i = 0; i < mq; i++) E.u[i] = Qu[i] * dt;
    for (int j = 0; j < nq; j++) for (int i = 0; i < nq; i++) E.xx[j * n + i] = S.Q[j * nq + i] * dt;
    for (int j = 0; j < mq; j++) for (int i = 0; i < mq; i++) E.uu[j * m + i] = S.R[j * mq + i] * dt;
    for (int j = 0; j < nq; j++) for (int i = 0; i < mq; i++) E.ux[j * m + i] = S.H[j * mq + i] * dt;
    if (S.mt) {
        double l1 = quad_stage(S, x, u);
        E.u[m - 1] = tau * (2.0 * l1 + S.R_mt);
        for (int i = 0; i < mq; i++) {
            double t = (2.0 * tau) * Qu[i];
            E.uu[(m - 1) * m + i] = t;
            E.uu[i * m + (m - 1)] = t;
        }
        E.uu[(m - 1) * m + (m - 1)] = 2.0 * l1 + S.R_mt;
        for (int i = 0; i < nq; i++) E.ux[i * m + (m - 1)] = (2.0 * tau) * Qx[i];
        E.x[n - 1] = S.R_mt * x[n - 1];
        E.xx[(n - 1) * n + (n - 1)] = S.R_mt;
    }
}
// src/cost.jl:194-198 ; minimum_time.jl:193-204
static void term_expansion(const Spec& S, Exp& E, const double* x) {
    const int nq = S.nq, n = S.n;
    for (int j = 0; j < nq; j++) for (int i = 0; i < nq; i++) E.xx[j * n + i] = S.Qf[j * nq + i];
    for (int i = 0; i < nq; i++) {
        double a = 0.0;
        for (int j = 0; j < nq; j++) a = fmad(S.Qf[j * nq + i], x[j], a);
        E.x[i] = a + S.qf[i];
    }
    if (S.mt) {
        E.xx[(n - 1) * n + (n - 1)] = S.R_mt;
        E.x[n - 1] = S.R_mt * x[n - 1];
    }
}

// ---- constraints -------------------------------------------------------------------------
// values: src/constraints.jl:212-227 (bounds), :299-304 (goal), :306-314 (infeasible),
// src/utils.jl:140-156 (circle/sphere), minimum_time.jl:112-124 (min-time equality)
static double row_value(const Spec& S, const Row& r, const double* x, const double* u) {
    switch (r.kind) {
        case R_LIN: {
            double z = (r.col < S.n) ? x[r.col] : u[r.col - S.n];
            return (r.sign > 0) ? (z - r.a) : (r.a - z);
        }
        case R_CIRCLE: {
            double dx = x[0] - r.a, dy = x[1] - r.b;
            return -(((dx * dx) + (dy * dy)) - (r.r * r.r));
        }
        case R_SPHERE: {
            double dx = x[0] - r.a, dy = x[1] - r.b, dz = x[2] - r.c;
            return -((((dx * dx) + (dy * dy)) + (dz * dz)) - (r.r * r.r));
        }
        case R_MTEQ: return u[S.m - 1] - x[S.n - 1];
    }
    return 0.0;
}
// Jacobian row over [x̄;ū] (terminal: over x̄).  Bounds/goal/slack: constant ±1
// (constraints.jl:229-237,302,311-312); obstacles: ForwardDiff of the value expression
// (constraints.jl:85-89): d/dx1 = -(2*(x1-a)) etc.
static void row_jac(const Spec& S, const Row& r, const double* x, double* g /* n+m zeros */) {
    switch (r.kind) {
        case R_LIN: g[r.col] = (r.sign > 0) ? 1.0 : -1.0; break;
        case R_CIRCLE: g[0] = -(2.0 * (x[0] - r.a)); g[1] = -(2.0 * (x[1] - r.b)); break;
        case R_SPHERE: g[0] = -(2.0 * (x[0] - r.a)); g[1] = -(2.0 * (x[1] - r.b)); g[2] = -(2.0 * (x[2] - r.c)); break;
        case R_MTEQ: g[S.n + S.m - 1] = 1.0; g[S.n - 1] = -1.0; break;
    }
}

// ------------------------------------------------------------------------------------------
// solver state
// ------------------------------------------------------------------------------------------
struct Trace {
    std::vector<TOIterRecord> inner;
    std::vector<TOOuterRecord> outer;
};

struct ILQR {  // src/solvers/ilqr/ilqr_solver.jl:93-112
    TOiLQROptions o;
    std::vector<vec> Xb, Ub, K, d, A, B;
    std::vector<Exp> Q;
    std::vector<vec> Sxx, Sx;
    double rho = 0, drho = 0;
    int iterations = 0, dJ_zero = 0, steps = 0;
    double last_dJ = 0, last_grad = 0, last_cost = 0;
    // forward-pass log
    double fp_expected = 0, fp_z = 0, fp_alpha = 0;
    bool rho_diverged = false;
};

struct AL {  // src/solvers/augmented_lagrangian/augmented_lagrangian_solver.jl:96-110
    std::vector<vec> C, lam, mu;
    std::vector<std::vector<uint8_t>> act;
    bool on = false;
};

struct Prob {
    const Spec* S;
    vec x0;
    std::vector<vec> X, U;
};

static bool all_finite(const std::vector<vec>& X) {
    for (auto& v : X) for (double a : v) if (!std::isfinite(a)) return false;
    return true;
}

// src/rollout.jl:25-38
static void rollout_open(Prob& p) {
    const Spec& S = *p.S;
    if (!all_finite(p.X)) {
        p.X[0] = p.x0;
        for (int k = 0; k < S.N - 1; k++) dyn_eval(S, p.X[k + 1].data(), p.X[k].data(), p.U[k].data());
    }
}

// src/constraint_sets.jl:221-228 + :247-267 (active set, tol = 0)
static void update_constraints(const Spec& S, AL& al, const std::vector<vec>& X, const std::vector<vec>& U) {
    for (int k = 0; k < S.N; k++) {
        const auto& rows = S.rows[k];
        const double* u = (k < S.N - 1) ? U[k].data() : nullptr;
        for (size_t i = 0; i < rows.size(); i++) al.C[k][i] = row_value(S, rows[i], X[k].data(), u);
        for (size_t i = 0; i < rows.size(); i++)
            al.act[k][i] = rows[i].eq ? 1 : ((al.C[k][i] >= 0.0) || (al.lam[k][i] > 0.0));
    }
}
static void update_active_set(const Spec& S, AL& al) {
    for (int k = 0; k < S.N; k++)
        for (size_t i = 0; i < S.rows[k].size(); i++)
            al.act[k][i] = S.rows[k][i].eq ? 1 : ((al.C[k][i] >= 0.0) || (al.lam[k][i] > 0.0));
}

// src/objective.jl:40-48 ; AL: augmented_lagrangian_methods.jl:284-313
static double cost_fn(const Spec& S, AL& al, const std::vector<vec>& X, const std::vector<vec>& U) {
    double J = 0.0;
    for (int k = 0; k < S.N - 1; k++) J += stage_cost(S, X[k].data(), U[k].data());
    J += term_cost(S, X[S.N - 1].data());
    if (!al.on) return J;
    update_constraints(S, al, X, U);
    double Jc = 0.0;
    for (int k = 0; k < S.N; k++) {
        const int p = (int)S.rows[k].size();
        // aula_cost: λ'c + ((1/2*c')*Diagonal(a.*μ))*c
        double t1 = dotc(al.lam[k].data(), al.C[k].data(), p);
        double t2 = 0.0;
        for (int i = 0; i < p; i++) {
            double am = al.act[k][i] ? al.mu[k][i] : 0.0;
            t2 = fmad((0.5 * al.C[k][i]) * am, al.C[k][i], t2);
        }
        Jc += t1 + t2;
    }
    return J + Jc;
}

// augmented_lagrangian_methods.jl:171-184
static double max_violation(const Spec& S, const AL& al) {
    double cmax = 0.0;
    for (int k = 0; k < S.N; k++) {
        const auto& rows = S.rows[k];
        if (rows.empty()) continue;
        double e = 0.0, mi = -std::numeric_limits<double>::infinity();
        bool has_i = false;
        for (size_t i = 0; i < rows.size(); i++) {
            if (rows[i].eq) e = std::max(e, std::fabs(al.C[k][i]));
            else { has_i = true; mi = std::max(mi, al.C[k][i]); }
        }
        cmax = std::max(e, cmax);
        if (has_i) cmax = std::max(std::max(0.0, mi), cmax);
    }
    return cmax;
}

// ilqr_methods.jl:55-62 -> objective.jl:56-63 / augmented_lagrangian_methods.jl:186-229
// returns false if the sqrt expansion met a non-PD Hessian (objective.jl:76-93)
static bool chol_upper_inplace(double* A, int n);  // fwd
static void chol_plus_inplace(double* Aup, int n, const double* Bm, int nb);  // fwd

static bool cost_expansion(const Spec& S, ILQR& s, AL& al, const Prob& p) {
    const int n = S.n, m = S.m, N = S.N;
    for (auto& e : s.Q) e.reset();
    for (int k = 0; k < N - 1; k++) {
        stage_expansion(S, s.Q[k], p.X[k].data(), p.U[k].data());
        if (s.o.square_root) {
            if (!chol_upper_inplace(s.Q[k].xx.data(), n)) return false;
            if (!chol_upper_inplace(s.Q[k].uu.data(), m)) return false;
        }
    }
    term_expansion(S, s.Q[N - 1], p.X[N - 1].data());
    if (s.o.square_root && !chol_upper_inplace(s.Q[N - 1].xx.data(), n)) return false;
    if (!al.on) return true;
    for (int k = 0; k < N; k++) {
        const auto& rows = S.rows[k];
        const int pk = (int)rows.size();
        if (pk == 0) continue;
        const bool term = (k == N - 1);
        const int nz = term ? n : n + m;
        vec G((size_t)pk * nz, 0.0);  // row-major p × nz
        vec Imu(pk), g(pk);
        for (int i = 0; i < pk; i++) {
            bool a = rows[i].eq ? true : ((al.C[k][i] >= 0.0) || (al.lam[k][i] > 0.0));
            Imu[i] = a ? al.mu[k][i] : 0.0;
            vec row(n + m, 0.0);
            row_jac(S, rows[i], p.X[k].data(), row.data());
            for (int j = 0; j < nz; j++) G[(size_t)i * nz + j] = row[j];
            g[i] = Imu[i] * al.C[k][i] + al.lam[k][i];
        }
        Exp& E = s.Q[k];
        if (!s.o.square_root) {
            // (cx'Iμ)*cx : element (i,j) = Σ_r (cx[r,i]*Iμ[r]) * cx[r,j]
            for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
                double acc = 0.0;
                for (int r = 0; r < pk; r++) acc = fmad(G[(size_t)r * nz + i] * Imu[r], G[(size_t)r * nz + j], acc);
                E.xx[j * n + i] += acc;
            }
            if (!term) {
                for (int j = 0; j < m; j++) for (int i = 0; i < m; i++) {
                    double acc = 0.0;
                    for (int r = 0; r < pk; r++) acc = fmad(G[(size_t)r * nz + n + i] * Imu[r], G[(size_t)r * nz + n + j], acc);
                    E.uu[j * m + i] += acc;
                }
                for (int j = 0; j < n; j++) for (int i = 0; i < m; i++) {
                    double acc = 0.0;
                    for (int r = 0; r < pk; r++) acc = fmad(G[(size_t)r * nz + n + i] * Imu[r], G[(size_t)r * nz + j], acc);
                    E.ux[j * m + i] += acc;
                }
            }
        } else {
            // augmented_lagrangian_methods.jl:231-276: chol_plus!(Q.xx, Iμ_sqrt*cx) ; no ux term (Q17)
            vec Bx((size_t)pk * n), Bu((size_t)pk * m);
            for (int r = 0; r < pk; r++) {
                double sq = (rows[r].eq || (al.C[k][r] >= 0.0) || (al.lam[k][r] > 0.0)) ? std::sqrt(al.mu[k][r]) : 0.0;
                for (int j = 0; j < n; j++) Bx[(size_t)j * pk + r] = sq * G[(size_t)r * nz + j];
                if (!term) for (int j = 0; j < m; j++) Bu[(size_t)j * pk + r] = sq * G[(size_t)r * nz + n + j];
            }
            chol_plus_inplace(E.xx.data(), n, Bx.data(), pk);
            if (!term) chol_plus_inplace(E.uu.data(), m, Bu.data(), pk);
        }
        for (int i = 0; i < n; i++) {
            double acc = 0.0;
            for (int r = 0; r < pk; r++) acc = fmad(G[(size_t)r * nz + i], g[r], acc);
            E.x[i] += acc;
        }
        if (!term)
            for (int i = 0; i < m; i++) {
                double acc = 0.0;
                for (int r = 0; r < pk; r++) acc = fmad(G[(size_t)r * nz + n + i], g[r], acc);
                E.u[i] += acc;
            }
    }
    return true;
}

// ---- small dense linear algebra ---------------------------------------------------------
// isposdef(Hermitian(A)) / cholesky(A).U : upper-triangle Cholesky, dot-product (left-looking) form.
static bool chol_upper(const double* A, int n, double* U) {
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < j; i++) {
            double acc = 0.0;
            for (int l = 0; l < i; l++) acc = fmad(U[i * n + l], U[j * n + l], acc);
            U[j * n + i] = (A[j * n + i] - acc) / U[i * n + i];
        }
        double acc = 0.0;
        for (int l = 0; l < j; l++) acc = fmad(U[j * n + l], U[j * n + l], acc);
        double dd = A[j * n + j] - acc;
        if (!(dd > 0.0)) return false;
        U[j * n + j] = std::sqrt(dd);
        for (int i = j + 1; i < n; i++) U[j * n + i] = 0.0;
    }
    return true;
}
static bool chol_upper_inplace(double* A, int n) {
    vec U((size_t)n * n, 0.0);
    if (!chol_upper(A, n, U.data())) return false;
    std::memcpy(A, U.data(), sizeof(double) * n * n);
    return true;
}

// Quu_reg \ RHS (backward_pass.jl:66-67): Julia's dense `\`: triangular shortcuts, else LU with
// partial pivoting (left-looking getf2 with reciprocal scaling, as OpenBLAS), then getrs.
static void solve_general(const double* Ain, int m, const double* Bin, int nrhs, double* Xout) {
    vec A(Ain, Ain + m * m), Bm(Bin, Bin + (size_t)m * nrhs);
    bool tril = true, triu = true;
    for (int j = 0; j < m; j++) for (int i = 0; i < m; i++) {
        if (i < j && A[j * m + i] != 0.0) tril = false;
        if (i > j && A[j * m + i] != 0.0) triu = false;
    }
    if (tril && !triu) {  // LowerTriangular forward substitution
        for (int c = 0; c < nrhs; c++) {
            double* b = &Bm[(size_t)c * m];
            for (int i = 0; i < m; i++) {
                double acc = 0.0;
                for (int l = 0; l < i; l++) acc = fmad(A[l * m + i], b[l], acc);
                b[i] = (b[i] - acc) / A[i * m + i];
            }
        }
        std::memcpy(Xout, Bm.data(), sizeof(double) * m * nrhs);
        return;
    }
    std::vector<int> piv(m);
    if (!triu) {
        for (int j = 0; j < m; j++) {
            // apply previous interchanges to column j
            for (int i = 0; i < j; i++) { int pi = piv[i]; if (pi != i) std::swap(A[j * m + i], A[j * m + pi]); }
            // U part: a[i,j] -= dot(L[i,0:i], a[0:i,j])
            for (int i = 1; i < j; i++) {
                double acc = 0.0;
                for (int l = 0; l < i; l++) acc = fmad(A[l * m + i], A[j * m + l], acc);
                A[j * m + i] = A[j * m + i] - acc;
            }
            // remaining rows
            for (int i = j; i < m; i++) {
                double acc = 0.0;
                for (int l = 0; l < j; l++) acc = fmad(A[l * m + i], A[j * m + l], acc);
                A[j * m + i] = A[j * m + i] - acc;
            }
            int p = j;
            double amax = std::fabs(A[j * m + j]);
            for (int i = j + 1; i < m; i++) if (std::fabs(A[j * m + i]) > amax) { amax = std::fabs(A[j * m + i]); p = i; }
            piv[j] = p;
            if (p != j) for (int c = 0; c <= j; c++) std::swap(A[c * m + j], A[c * m + p]);
            double rp = 1.0 / A[j * m + j];
            for (int i = j + 1; i < m; i++) A[j * m + i] = A[j * m + i] * rp;
        }
        // the interchanges of later columns still have to be applied to earlier L columns
        // (handled above by swapping columns 0..j at step j)
    } else {
        for (int j = 0; j < m; j++) piv[j] = j;
    }
    for (int c = 0; c < nrhs; c++) {
        double* b = &Bm[(size_t)c * m];
        if (!triu) {
            for (int i = 0; i < m; i++) if (piv[i] != i) std::swap(b[i], b[piv[i]]);
            for (int i = 1; i < m; i++) {
                double acc = 0.0;
                for (int l = 0; l < i; l++) acc = fmad(A[l * m + i], b[l], acc);
                b[i] = b[i] - acc;
            }
        }
        for (int i = m - 1; i >= 0; i--) {
            double acc = 0.0;
            for (int l = i + 1; l < m; l++) acc = fmad(A[l * m + i], b[l], acc);
            b[i] = (b[i] - acc) / A[i * m + i];
        }
    }
    std::memcpy(Xout, Bm.data(), sizeof(double) * m * nrhs);
}

// R factor of qr([A;B]) (backward_pass.jl:172-183), A: n×n upper (col-major, full storage), B: nb×n.
// Householder QR restating LAPACK dgeqr2/dlarfg; R may have negative diagonal entries, as LAPACK's.
static void qr_R(double* P, int rows, int n) {  // in place, P rows×n col-major; returns R in top n×n
    for (int j = 0; j < n && j < rows; j++) {
        // dlarfg on P[j:rows, j]
        double alpha = P[j * rows + j];
        double xn2 = 0.0;
        for (int i = j + 1; i < rows; i++) xn2 = fmad(P[j * rows + i], P[j * rows + i], xn2);
        double tau = 0.0;
        if (xn2 != 0.0) {
            double xnorm = std::sqrt(xn2);
            double beta = -std::copysign(std::sqrt(alpha * alpha + xnorm * xnorm), alpha);
            tau = (beta - alpha) / beta;
            double sc = 1.0 / (alpha - beta);
            for (int i = j + 1; i < rows; i++) P[j * rows + i] = P[j * rows + i] * sc;
            P[j * rows + j] = beta;
        }
        // apply H = I - tau v v' to the trailing columns, v = [1; P[j+1:rows, j]]
        if (tau != 0.0) {
            for (int c = j + 1; c < n; c++) {
                double w = P[c * rows + j];
                for (int i = j + 1; i < rows; i++) w = fmad(P[j * rows + i], P[c * rows + i], w);
                double tw = tau * w;
                P[c * rows + j] = P[c * rows + j] - tw;
                for (int i = j + 1; i < rows; i++) P[c * rows + i] = fmad(-tw, P[j * rows + i], P[c * rows + i]);
            }
        }
    }
}
static void chol_plus(const double* A, int na, const double* Bm, int nb, int n, double* Rout) {
    const int rows = na + nb;
    vec P((size_t)rows * n, 0.0);
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < na; i++) P[(size_t)j * rows + i] = A[j * na + i];
        for (int i = 0; i < nb; i++) P[(size_t)j * rows + na + i] = Bm[(size_t)j * nb + i];
    }
    qr_R(P.data(), rows, n);
    // qr(P).R is min(rows,n)×n; callers always have rows >= n
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) Rout[j * n + i] = (i <= j) ? P[(size_t)j * rows + i] : 0.0;
}
static void chol_plus_inplace(double* Aup, int n, const double* Bm, int nb) {
    vec R((size_t)n * n);
    chol_plus(Aup, n, Bm, nb, n, R.data());
    std::memcpy(Aup, R.data(), sizeof(double) * n * n);
}

// lowrankdowndate!(Cholesky(A,:U), v) (LinearAlgebra/cholesky.jl), used by chol_minus
// (backward_pass.jl:186-192).  Returns false on PosDefException.
static bool lowrank_downdate(double* U, int n, double* v) {
    for (int i = 0; i < n; i++) {
        double Aii = U[i * n + i];
        double s = v[i] / Aii;  // conj(v[i])/A[i,i]
        double s2 = s * s;
        if (s2 > 1.0) return false;
        double c = std::sqrt(1.0 - s2);
        U[i * n + i] = c * Aii;
        for (int j = i + 1; j < n; j++) {
            double vj = v[j];
            double Aij = (U[j * n + i] - s * vj) / c;
            U[j * n + i] = Aij;
            v[j] = -s * Aij + c * vj;
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------
// iLQR
// ------------------------------------------------------------------------------------------
// ilqr_methods.jl:164-176
static void reg_update(ILQR& s, bool increase) {
    const TOiLQROptions& o = s.o;
    if (increase) {
        s.drho = std::max(s.drho * o.bp_reg_increase_factor, o.bp_reg_increase_factor);
        s.rho = std::max(s.rho * s.drho, o.bp_reg_min);
    } else {
        s.drho = std::min(s.drho / o.bp_reg_increase_factor, 1.0 / o.bp_reg_increase_factor);
        s.rho = s.rho * s.drho * ((s.rho * s.drho > o.bp_reg_min) ? 1.0 : 0.0);
    }
}

// C = A' * B  (A: ra×ca, B: ra×cb) -> ca×cb ; all col-major
static void mul_AtB(const double* A, int ra, int ca, const double* Bm, int cb, double* C) {
    for (int j = 0; j < cb; j++) for (int i = 0; i < ca; i++) {
        double acc = 0.0;
        for (int l = 0; l < ra; l++) acc = fmad(A[i * ra + l], Bm[j * ra + l], acc);
        C[j * ca + i] = acc;
    }
}
// C = A * B (A: ra×ca, B: ca×cb)
static void mul_AB(const double* A, int ra, int ca, const double* Bm, int cb, double* C) {
    for (int j = 0; j < cb; j++) for (int i = 0; i < ra; i++) {
        double acc = 0.0;
        for (int l = 0; l < ca; l++) acc = fmad(A[l * ra + i], Bm[j * ca + l], acc);
        C[j * ra + i] = acc;
    }
}

// backward_pass.jl:9-85.  Returns ΔV.  false: the reference's restart loop can no longer make progress
// (PD test failed with a non-finite rho; the reference would spin forever, see TO_STATUS_REG_DIVERGED).
static bool backwardpass(const Spec& S, ILQR& s, double dV[2]) {
    const int n = S.n, m = S.m, N = S.N;
    s.Sxx[N - 1] = s.Q[N - 1].xx;
    s.Sx[N - 1] = s.Q[N - 1].x;
    dV[0] = dV[1] = 0.0;
    vec T(n * n), Tu(m * n), M(n * n), Mu(m * m), Mux(m * n), v(n), vu(m);
    vec Quu_reg(m * m), Kk(m * n), dk(m), KQ(n * m), M1(n * n), M2(n * n), M3(n * n), v1(n), v2(n), v3(n), Uc(m * m);
    vec RB(m * n), Qux_r(m * n);
    int k = N - 2;
    while (k >= 0) {
        const double* A = s.A[k].data();
        const double* B = s.B[k].data();
        Exp& Q = s.Q[k];
        const double* Sxx = s.Sxx[k + 1].data();
        const double* Sx = s.Sx[k + 1].data();
        mul_AtB(A, n, n, Sx, 1, v.data());
        for (int i = 0; i < n; i++) Q.x[i] += v[i];
        mul_AtB(B, n, m, Sx, 1, vu.data());
        for (int i = 0; i < m; i++) Q.u[i] += vu[i];
        mul_AtB(A, n, n, Sxx, n, T.data());       // A'S
        mul_AB(T.data(), n, n, A, n, M.data());   // (A'S)A
        for (int i = 0; i < n * n; i++) Q.xx[i] += M[i];
        mul_AtB(B, n, m, Sxx, n, Tu.data());      // B'S  (m×n)
        mul_AB(Tu.data(), m, n, B, m, Mu.data()); // (B'S)B
        for (int i = 0; i < m * m; i++) Q.uu[i] += Mu[i];
        mul_AB(Tu.data(), m, n, A, n, Mux.data());
        for (int i = 0; i < m * n; i++) Q.ux[i] += Mux[i];

        const double* Qux_reg = Q.ux.data();
        if (s.o.bp_reg_type == TO_REG_STATE) {
            // backward_pass.jl:38-41: Quu_reg = Q.uu + rho*fdu'*fdu, Qux_reg = Q.ux + rho*fdu'*fdx -- (rho*B') first, then the product
            for (int l = 0; l < n; l++) for (int i = 0; i < m; i++) RB[l * m + i] = s.rho * B[i * n + l];   // m×n, col-major
            mul_AB(RB.data(), m, n, B, m, Mu.data());
            for (int i = 0; i < m * m; i++) Quu_reg[i] = Q.uu[i] + Mu[i];
            mul_AB(RB.data(), m, n, A, n, Mux.data());
            for (int i = 0; i < m * n; i++) Qux_r[i] = Q.ux[i] + Mux[i];
            Qux_reg = Qux_r.data();
        } else {
            for (int i = 0; i < m * m; i++) Quu_reg[i] = Q.uu[i];
            for (int i = 0; i < m; i++) Quu_reg[i * m + i] = Q.uu[i * m + i] + s.rho * 1.0;
        }
        if (!chol_upper(Quu_reg.data(), m, Uc.data())) {
            if (!std::isfinite(s.rho)) return false;
            reg_update(s, true);
            k = N - 2;
            dV[0] = dV[1] = 0.0;
            continue;
        }
        solve_general(Quu_reg.data(), m, Qux_reg, n, Kk.data());
        for (int i = 0; i < m * n; i++) Kk[i] = -1.0 * Kk[i];
        solve_general(Quu_reg.data(), m, Q.u.data(), 1, dk.data());
        for (int i = 0; i < m; i++) dk[i] = -1.0 * dk[i];
        s.K[k] = Kk;
        s.d[k] = dk;
        // S.x = Q.x + (K'Quu)d + K'Qu + Qux'd
        mul_AtB(Kk.data(), m, n, Q.uu.data(), m, KQ.data());  // n×m
        mul_AB(KQ.data(), n, m, dk.data(), 1, v1.data());
        mul_AtB(Kk.data(), m, n, Q.u.data(), 1, v2.data());
        mul_AtB(Q.ux.data(), m, n, dk.data(), 1, v3.data());
        s.Sx[k].resize(n);
        for (int i = 0; i < n; i++) s.Sx[k][i] = ((Q.x[i] + v1[i]) + v2[i]) + v3[i];
        // S.xx = Q.xx + (K'Quu)K + K'Qux + Qux'K ; symmetrise
        mul_AB(KQ.data(), n, m, Kk.data(), n, M1.data());
        mul_AtB(Kk.data(), m, n, Q.ux.data(), n, M2.data());
        mul_AtB(Q.ux.data(), m, n, Kk.data(), n, M3.data());
        vec Sn(n * n);
        for (int i = 0; i < n * n; i++) Sn[i] = ((Q.xx[i] + M1[i]) + M2[i]) + M3[i];
        s.Sxx[k].resize(n * n);
        for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) s.Sxx[k][j * n + i] = 0.5 * (Sn[j * n + i] + Sn[i * n + j]);
        // ΔV
        dV[0] += dotc(dk.data(), Q.u.data(), m);
        {
            double acc = 0.0;
            for (int j = 0; j < m; j++) {
                double w = 0.0;
                for (int l = 0; l < m; l++) w = fmad(0.5 * dk[l], Q.uu[j * m + l], w);
                acc = fmad(w, dk[j], acc);
            }
            dV[1] += acc;
        }
        k--;
    }
    reg_update(s, false);
    return true;
}

// cond(A) for the sqrt restart test (backward_pass.jl:129): ratio of extreme singular values.
// One-sided Jacobi SVD (deterministic sweep order).  "parity unpinned" (LAPACK gesdd in Julia).
static double cond2(const double* Ain, int n) {
    vec A(Ain, Ain + n * n);
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int p = 0; p < n - 1; p++) for (int q = p + 1; q < n; q++) {
            double a = 0, b = 0, c = 0;
            for (int i = 0; i < n; i++) { a = fmad(A[p * n + i], A[p * n + i], a); b = fmad(A[q * n + i], A[q * n + i], b); c = fmad(A[p * n + i], A[q * n + i], c); }
            if (c == 0.0) continue;
            off = std::max(off, std::fabs(c) / std::sqrt(a * b));
            double zeta = (b - a) / (2.0 * c);
            double t = std::copysign(1.0, zeta) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
            double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
            for (int i = 0; i < n; i++) {
                double ap = A[p * n + i], aq = A[q * n + i];
                A[p * n + i] = cs * ap - sn * aq;
                A[q * n + i] = sn * ap + cs * aq;
            }
        }
        if (off < 1e-15) break;
    }
    double smax = 0.0, smin = std::numeric_limits<double>::infinity();
    for (int j = 0; j < n; j++) {
        double a = 0;
        for (int i = 0; i < n; i++) a = fmad(A[j * n + i], A[j * n + i], a);
        double sv = std::sqrt(a);
        smax = std::max(smax, sv);
        smin = std::min(smin, sv);
    }
    return smax / smin;
}

// upper-triangular solves used by the sqrt pass: X = U \ B and X = U' \ B
static void solve_upper(const double* U, int m, const double* Bin, int nrhs, double* X) {
    for (int c = 0; c < nrhs; c++) {
        for (int i = m - 1; i >= 0; i--) {
            double acc = 0.0;
            for (int l = i + 1; l < m; l++) acc = fmad(U[l * m + i], X[c * m + l], acc);
            X[c * m + i] = (Bin[c * m + i] - acc) / U[i * m + i];
        }
    }
}
static void solve_upper_t(const double* U, int m, const double* Bin, int nrhs, double* X) {
    for (int c = 0; c < nrhs; c++) {
        for (int i = 0; i < m; i++) {
            double acc = 0.0;
            for (int l = 0; l < i; l++) acc = fmad(U[i * m + l], X[c * m + l], acc);
            X[c * m + i] = (Bin[c * m + i] - acc) / U[i * m + i];
        }
    }
}

// backward_pass.jl:87-169.  Returns false on PosDefException in chol_minus.
static bool backwardpass_sqrt(const Spec& S, ILQR& s, double dV[2]) {
    const int n = S.n, m = S.m, N = S.N;
    s.Sxx[N - 1] = s.Q[N - 1].xx;
    s.Sx[N - 1] = s.Q[N - 1].x;
    dV[0] = dV[1] = 0.0;
    vec v(n), vu(m), tx(n * n), tu(n * m), Mux(m * n), Quu_reg(m * m), eye(m * m), t1(m * n), Kk(m * n), dk(m), t1d(m);
    int k = N - 2;
    while (k >= 0) {
        const double* A = s.A[k].data();
        const double* B = s.B[k].data();
        Exp& Q = s.Q[k];
        const double* Sxx = s.Sxx[k + 1].data();
        const double* Sx = s.Sx[k + 1].data();
        mul_AtB(A, n, n, Sx, 1, v.data());
        for (int i = 0; i < n; i++) Q.x[i] += v[i];
        mul_AtB(B, n, m, Sx, 1, vu.data());
        for (int i = 0; i < m; i++) Q.u[i] += vu[i];
        mul_AB(Sxx, n, n, A, n, tx.data());  // S*A (n×n)
        mul_AB(Sxx, n, n, B, m, tu.data());  // S*B (n×m)
        chol_plus_inplace(Q.xx.data(), n, tx.data(), n);
        chol_plus_inplace(Q.uu.data(), m, tu.data(), n);
        mul_AtB(tu.data(), n, m, tx.data(), n, Mux.data());
        for (int i = 0; i < m * n; i++) Q.ux[i] += Mux[i];
        std::fill(eye.begin(), eye.end(), 0.0);
        double sr = std::sqrt(s.rho);
        for (int i = 0; i < m; i++) eye[i * m + i] = sr * 1.0;
        chol_plus(Q.uu.data(), m, eye.data(), m, m, Quu_reg.data());
        if (cond2(Quu_reg.data(), m) > 1e8) {
            if (!std::isfinite(s.rho)) { s.rho_diverged = true; return false; }
            reg_update(s, true);
            k = N - 2;
            dV[0] = dV[1] = 0.0;
            continue;
        }
        // K = -Quu_reg \ (Quu_reg' \ Qux)
        solve_upper_t(Quu_reg.data(), m, Q.ux.data(), n, t1.data());
        solve_upper(Quu_reg.data(), m, t1.data(), n, Kk.data());
        for (int i = 0; i < m * n; i++) Kk[i] = -Kk[i];
        solve_upper_t(Quu_reg.data(), m, Q.u.data(), 1, t1d.data());
        solve_upper(Quu_reg.data(), m, t1d.data(), 1, dk.data());
        for (int i = 0; i < m; i++) dk[i] = -dk[i];
        s.K[k] = Kk;
        s.d[k] = dk;
        // S.x = Q.x + (K'Quu')(Quu d) + K'Qu + Qux'd
        vec KQt(n * m), Qd(m), v1(n), v2(n), v3(n);
        for (int j = 0; j < m; j++) for (int i = 0; i < n; i++) {  // (K' * Quu')[i,j] = Σ_l K[l,i]*Quu[j,l]
            double acc = 0.0;
            for (int l = 0; l < m; l++) acc = fmad(Kk[i * m + l], Q.uu[l * m + j], acc);
            KQt[j * n + i] = acc;
        }
        mul_AB(Q.uu.data(), m, m, dk.data(), 1, Qd.data());
        mul_AB(KQt.data(), n, m, Qd.data(), 1, v1.data());
        mul_AtB(Kk.data(), m, n, Q.u.data(), 1, v2.data());
        mul_AtB(Q.ux.data(), m, n, dk.data(), 1, v3.data());
        s.Sx[k].resize(n);
        for (int i = 0; i < n; i++) s.Sx[k][i] = ((Q.x[i] + v1[i]) + v2[i]) + v3[i];
        // tmp1 = (Q.xx') \ Q.ux'   (n×m)
        vec uxT(n * m), tmp1(n * m);
        for (int j = 0; j < m; j++) for (int i = 0; i < n; i++) uxT[j * n + i] = Q.ux[i * m + j];
        solve_upper_t(Q.xx.data(), n, uxT.data(), m, tmp1.data());
        // tmp2 = chol_minus(Q.uu, tmp1): downdate with each ROW of tmp1 (length m)
        vec U2(Q.uu);
        for (int i = 0; i < n; i++) {
            vec rowv(m);
            for (int j = 0; j < m; j++) rowv[j] = tmp1[j * n + i];
            if (!lowrank_downdate(U2.data(), m, rowv.data())) return false;
        }
        // S.xx = chol_plus(Q.xx + tmp1*K, tmp2*K)
        vec top(n * n), bot(m * n), Rn(n * n);
        mul_AB(tmp1.data(), n, m, Kk.data(), n, top.data());
        for (int i = 0; i < n * n; i++) top[i] = Q.xx[i] + top[i];
        mul_AB(U2.data(), m, m, Kk.data(), n, bot.data());
        chol_plus(top.data(), n, bot.data(), m, n, Rn.data());
        s.Sxx[k] = Rn;
        dV[0] += dotc(dk.data(), Q.u.data(), m);
        dV[1] += 0.5 * dotc(Qd.data(), Qd.data(), m);
        k--;
    }
    reg_update(s, false);
    return true;
}

// src/rollout.jl:2-23
static bool rollout_cl(const Spec& S, ILQR& s, const Prob& p, double alpha) {
    const int n = S.n, m = S.m, N = S.N;
    s.Xb[0] = p.x0;
    vec dx(n);
    for (int k = 1; k < N; k++) {
        for (int i = 0; i < n; i++) dx[i] = s.Xb[k - 1][i] - p.X[k - 1][i];
        const vec& K = s.K[k - 1];
        for (int i = 0; i < m; i++) {
            double acc = 0.0;
            for (int j = 0; j < n; j++) acc = fmad(K[j * m + i], dx[j], acc);
            s.Ub[k - 1][i] = (p.U[k - 1][i] + acc) + alpha * s.d[k - 1][i];
        }
        dyn_eval(S, s.Xb[k].data(), s.Xb[k - 1].data(), s.Ub[k - 1].data());
        double mx = 0.0, mu = 0.0;
        bool bad = false;
        for (int i = 0; i < n; i++) { double a = std::fabs(s.Xb[k][i]); if (a != a) bad = true; mx = std::max(mx, a); }
        for (int i = 0; i < m; i++) { double a = std::fabs(s.Ub[k - 1][i]); if (a != a) bad = true; mu = std::max(mu, a); }
        if (bad || !(mx < s.o.max_state_value && mu < s.o.max_control_value)) return false;
    }
    return true;
}

// ilqr_methods.jl:122-129
static double gradient_todorov(const Spec& S, const ILQR& s, const Prob& p) {
    double sum = 0.0;
    for (int k = 0; k < S.N - 1; k++) {
        double mx = -std::numeric_limits<double>::infinity();
        for (int i = 0; i < S.m; i++) {
            double v = std::fabs(s.d[k][i]) / (std::fabs(p.U[k][i]) + 1.0);
            if (v != v) { mx = v; break; }
            mx = std::max(mx, v);
        }
        sum += mx;
    }
    sum += 0.0;
    return sum / (double)S.N;
}

// norm(v) of a short Vector{Float64} (LinearAlgebra generic_norm2, Julia 1.1 generic.jl: BLAS nrm2 only from 32 elements on)
static double julia_norm2(const double* v, int len) {
    double maxabs = 0.0;  // generic_normInf
    for (int i = 0; i < len; i++) { double a = std::fabs(v[i]); maxabs = (maxabs != maxabs || maxabs > a) ? maxabs : a; }
    if (maxabs == 0.0 || std::isinf(maxabs)) return maxabs;
    if (std::isfinite((double)len * maxabs * maxabs) && maxabs * maxabs != 0.0) {
        double sum = 0.0;
        for (int i = 0; i < len; i++) sum += v[i] * v[i];
        return std::sqrt(sum);
    }
    double inv = 1.0 / maxabs, sum = 0.0;
    for (int i = 0; i < len; i++) { double t = std::fabs(v[i]) * inv; sum += t * t; }
    return maxabs * std::sqrt(sum);
}
// gradient_feedforward (ilqr_methods.jl:133-137): norm(solver.d, Inf) over a vector of vectors = the largest 2-norm of a d[k]
// (generic_normInf applies `norm` to each element)
static double gradient_feedforward(const Spec& S, const ILQR& s) {
    double maxabs = 0.0;
    for (int k = 0; k < S.N - 1; k++) {
        const double a = julia_norm2(s.d[k].data(), S.m);
        maxabs = (k == 0) ? a : ((maxabs != maxabs || maxabs > a) ? maxabs : a);
    }
    return maxabs;
}

struct Ctx {
    const Spec* S;
    Prob p;
    ILQR s;
    AL al;
    Trace* tr = nullptr;
    int outer = 0;
    int status = 0;
    int steps_total = 0;
};

// forward_pass.jl:5-85 ; returns J, sets err on "cost increased"
static double forwardpass(Ctx& c, const double dV[2], double J_prev, bool& err) {
    const Spec& S = *c.S;
    ILQR& s = c.s;
    double J = std::numeric_limits<double>::infinity();
    double alpha = 1.0, z = -1.0, expected = 0.0;
    int iter = 0;
    while ((z <= s.o.line_search_lower_bound || z > s.o.line_search_upper_bound) && J >= J_prev) {
        if (iter > s.o.iterations_linesearch) {
            s.Xb = c.p.X;
            s.Ub = c.p.U;
            J = cost_fn(S, c.al, s.Xb, s.Ub);
            z = 0.0;
            alpha = 0.0;
            expected = 0.0;
            reg_update(s, true);
            s.rho += s.o.bp_reg_fp;
            break;
        }
        bool flag = rollout_cl(S, s, c.p, alpha);
        if (!flag) {
            iter += 1;
            alpha /= 2.0;
            continue;
        }
        J = cost_fn(S, c.al, s.Xb, s.Ub);
        expected = -alpha * (dV[0] + alpha * dV[1]);
        if (expected > 0) z = (J_prev - J) / expected;
        else z = -1.0;
        iter += 1;
        alpha /= 2.0;
    }
    s.fp_expected = expected;
    s.fp_z = z;
    s.fp_alpha = 2 * alpha;
    err = (J > J_prev);
    return J;
}

// ilqr_methods.jl:77-89
static void record_inner(Ctx& c, double J, double dJ) {
    ILQR& s = c.s;
    s.iterations += 1;
    s.last_cost = J;
    s.last_dJ = dJ;
    // calculate_gradient (ilqr_methods.jl:91-102)
    if (s.o.gradient_type == TO_GRAD_FEEDFORWARD) {
        s.last_grad = gradient_feedforward(*c.S, s);
    } else if (s.o.gradient_type == TO_GRAD_L2 || s.o.gradient_type == TO_GRAD_LINF) {
        // compute_gradient (:104-116): the (non-square-root) cost expansion at the current X, U; its first-order parts stacked
        // [Q1.x; Q1.u; ...; QN.x].  norm(g, Inf) is exact; norm(g) is BLAS nrm2 in the reference (more than 32 elements), whose
        // internal accumulation is OpenBLAS's -- restated as the plain sequential sum of squares ("parity unpinned" at the last bit)
        const int sq = s.o.square_root;
        s.o.square_root = 0;
        cost_expansion(*c.S, s, c.al, c.p);
        s.o.square_root = sq;
        const Spec& S = *c.S;
        double mx = 0.0, sum = 0.0;
        bool first = true;
        auto take = [&](double v) {
            const double a = std::fabs(v);
            mx = first ? a : ((mx != mx || mx > a) ? mx : a);
            first = false;
            sum += v * v;
        };
        for (int k = 0; k < S.N; k++) {
            for (int i = 0; i < S.n; i++) take(s.Q[k].x[i]);
            if (k < S.N - 1) for (int i = 0; i < S.m; i++) take(s.Q[k].u[i]);
        }
        s.last_grad = (s.o.gradient_type == TO_GRAD_LINF) ? mx : std::sqrt(sum);
    } else {
        s.last_grad = gradient_todorov(*c.S, s, c.p);
    }
    if (dJ == 0.0) s.dJ_zero += 1; else s.dJ_zero = 0;
    if (c.tr) {
        TOIterRecord r;
        r.cost = J; r.dJ = dJ; r.gradient = s.last_grad;
        r.expected = s.fp_expected; r.z = s.fp_z; r.alpha = s.fp_alpha; r.rho = s.rho;
        r.outer = c.outer; r.iter = s.iterations;
        c.tr->inner.push_back(r);
    }
}

// ilqr_methods.jl:139-162
static bool inner_converged(const ILQR& s) {
    if (0.0 < s.last_dJ && s.last_dJ < s.o.cost_tolerance) return true;
    if (s.last_grad < s.o.gradient_norm_tolerance) return true;
    if (s.iterations >= s.o.iterations) return true;
    if (s.dJ_zero > s.o.dJ_counter_limit) return true;
    return false;
}

// ilqr_methods.jl:3-45.  Returns false if the solve must abort (error thrown in the reference).
static bool ilqr_solve(Ctx& c) {
    const Spec& S = *c.S;
    ILQR& s = c.s;
    // reset! (ilqr_solver.jl:146-154)
    s.iterations = 0; s.dJ_zero = 0; s.rho = 0.0; s.drho = 0.0;
    s.fp_expected = 0; s.fp_z = 0; s.fp_alpha = 0;
    rollout_open(c.p);
    double J_prev = cost_fn(S, c.al, c.p.X, c.p.U);
    record_inner(c, J_prev, std::numeric_limits<double>::infinity());
    for (int i = 1; i <= s.o.iterations; i++) {
        // step! (ilqr_methods.jl:47-53)
        for (int k = 0; k < S.N - 1; k++) dyn_jac(S, c.p.X[k].data(), c.p.U[k].data(), s.A[k].data(), s.B[k].data());
        if (!cost_expansion(S, s, c.al, c.p)) { c.status |= TO_STATUS_NOT_PD_SQRT; return false; }
        double dV[2];
        if (s.o.square_root) {
            if (!backwardpass_sqrt(S, s, dV)) { c.status |= s.rho_diverged ? TO_STATUS_REG_DIVERGED : TO_STATUS_NOT_PD_SQRT; return false; }
        } else {
            if (!backwardpass(S, s, dV)) { c.status |= TO_STATUS_REG_DIVERGED; return false; }
        }
        bool err = false;
        double J = forwardpass(c, dV, J_prev, err);
        c.steps_total += 1;
        if (err) { c.status |= TO_STATUS_COST_INCREASED; return false; }
        if (J > s.o.max_cost_value) { c.status |= TO_STATUS_COST_BLOWUP; return true; }
        c.p.X = s.Xb;
        c.p.U = s.Ub;
        double dJ = std::fabs(J - J_prev);
        J_prev = J;
        record_inner(c, J, dJ);
        if (inner_converged(s)) break;
    }
    return true;
}

static void ilqr_init(Ctx& c, const TOiLQROptions& o) {
    const Spec& S = *c.S;
    ILQR& s = c.s;
    s.o = o;
    s.Xb.assign(S.N, vec(S.n, 0.0));
    s.Ub.assign(S.N - 1, vec(S.m, 0.0));
    s.K.assign(S.N - 1, vec(S.m * S.n, 0.0));
    s.d.assign(S.N - 1, vec(S.m, 0.0));
    s.A.assign(S.N - 1, vec(S.n * S.n, 0.0));
    s.B.assign(S.N - 1, vec(S.n * S.m, 0.0));
    s.Q.resize(S.N);
    for (auto& e : s.Q) e.init(S.n, S.m);
    s.Sxx.assign(S.N, vec(S.n * S.n, 0.0));
    s.Sx.assign(S.N, vec(S.n, 0.0));
    s.rho = s.drho = 0.0;
    s.iterations = 0;
}

// ------------------------------------------------------------------------------------------
// augmented Lagrangian outer loop (augmented_lagrangian_methods.jl:2-126)
// ------------------------------------------------------------------------------------------
struct ALStats { int iterations = 0, iterations_total = 0; double cost = 0, c_max = 0; };

static bool is_constrained(const Spec& S) {
    for (auto& r : S.rows) if (!r.empty()) return true;
    return false;
}

static void record_outer(Ctx& c, ALStats& st, double J) {
    const Spec& S = *c.S;
    double cmax = max_violation(S, c.al);
    st.iterations += 1;
    st.iterations_total += c.s.iterations;
    st.cost = J;
    st.c_max = cmax;
    if (c.tr) {
        TOOuterRecord r;
        r.cost = J; r.c_max = cmax;
        // max_penalty = maximum(maximum(μ)) over a vector of vectors (Q20): lexicographic max vector, then its max
        const vec* best = nullptr;
        for (auto& v : c.al.mu) { if (!best || std::lexicographical_compare(best->begin(), best->end(), v.begin(), v.end())) best = &v; }
        double pm = 0.0;
        if (best && !best->empty()) pm = *std::max_element(best->begin(), best->end());
        r.penalty_max = pm;
        r.iterations_inner = c.s.iterations;
        r.pad = 0;
        c.tr->outer.push_back(r);
    }
}

static bool al_solve(Ctx& c, const TOALOptions& o, ALStats& st) {
    const Spec& S = *c.S;
    AL& al = c.al;
    al.on = true;
    al.C.assign(S.N, vec());
    al.lam.assign(S.N, vec());
    al.mu.assign(S.N, vec());
    al.act.assign(S.N, std::vector<uint8_t>());
    for (int k = 0; k < S.N; k++) {
        size_t p = S.rows[k].size();
        al.C[k].assign(p, 0.0);
        al.lam[k].assign(p, 0.0);
        al.mu[k].assign(p, o.penalty_initial);
        al.act[k].assign(p, 1);
    }
    ilqr_init(c, o.opts_uncon);
    st = ALStats();
    c.outer = 0;
    rollout_open(c.p);
    double J0 = cost_fn(S, al, c.p.X, c.p.U);
    record_outer(c, st, J0);
    bool converged = false;
    for (int i = 1; i <= o.iterations; i++) {
        c.outer = i - 1;
        // set_tolerances! (:39-50)
        if (i != o.iterations) {
            c.s.o.cost_tolerance = o.cost_tolerance_intermediate;
            c.s.o.gradient_norm_tolerance = o.gradient_norm_tolerance_intermediate;
        } else {
            c.s.o.cost_tolerance = o.cost_tolerance;
            c.s.o.gradient_norm_tolerance = o.gradient_norm_tolerance;
        }
        // step! (:53-67)
        if (!ilqr_solve(c)) return false;
        double J = cost_fn(S, al, c.p.X, c.p.U);
        // dual_update! (:107-118)
        for (int k = 0; k < S.N; k++) {
            for (size_t r = 0; r < S.rows[k].size(); r++) {
                double v = al.lam[k][r] + al.mu[k][r] * al.C[k][r];
                v = std::max(o.dual_min, std::min(o.dual_max, v));
                if (!S.rows[k][r].eq) v = std::max(0.0, v);
                al.lam[k][r] = v;
            }
        }
        update_active_set(S, al);
        // penalty_update! (:121-126)
        for (int k = 0; k < S.N; k++)
            for (size_t r = 0; r < S.rows[k].size(); r++)
                al.mu[k][r] = std::max(0.0, std::min(o.penalty_max, o.penalty_scaling * al.mu[k][r]));
        record_outer(c, st, J);
        // evaluate_convergence (:70-77)
        converged = false;
        if (o.kickout_max_penalty) {
            const vec* best = nullptr;
            for (auto& v : al.mu) { if (!best || std::lexicographical_compare(best->begin(), best->end(), v.begin(), v.end())) best = &v; }
            double pm = (best && !best->empty()) ? *std::max_element(best->begin(), best->end()) : 0.0;
            converged = (pm == o.penalty_max);
        }
        converged = converged || (st.c_max < o.constraint_tolerance);
        if (converged) break;
        // reset!(solver_uncon): stats and rho only (ilqr_solver.jl:146-154)
        c.s.iterations = 0; c.s.dJ_zero = 0; c.s.rho = 0; c.s.drho = 0;
    }
    if (!converged) c.status |= TO_STATUS_MAX_OUTER;
    return true;
}

#include "oracle_pn.hpp"

// ------------------------------------------------------------------------------------------
// building the (augmented) Spec from the public descriptor
// ------------------------------------------------------------------------------------------
static void build_spec(const TOProblemDesc& D, bool inf, bool mt, const TOALTROOptions* ao, Spec& S) {
    S.model = D.model; S.integ = D.integrator;
    S.n0 = D.n; S.m0 = D.m; S.N = D.N; S.dt = D.dt;
    S.inf = inf; S.mt = mt;
    S.nq = D.n;
    S.mq = D.m + (inf ? D.n : 0);
    S.n = D.n + (mt ? 1 : 0);
    S.m = S.mq + (mt ? 1 : 0);
    const int n0 = D.n, m0 = D.m, nq = S.nq, mq = S.mq;
    S.Q.assign(D.Q, D.Q + n0 * n0);
    S.q.assign(D.q, D.q + n0);
    S.Qf.assign(D.Qf, D.Qf + n0 * n0);
    S.qf.assign(D.qf, D.qf + n0);
    S.c = D.c; S.cf = D.cf;
    S.R.assign((size_t)mq * mq, 0.0);
    S.r.assign(mq, 0.0);
    S.H.assign((size_t)mq * nq, 0.0);
    for (int j = 0; j < m0; j++) for (int i = 0; i < m0; i++) S.R[j * mq + i] = D.R[j * m0 + i];
    for (int i = 0; i < m0; i++) S.r[i] = D.r[i];
    if (D.H) for (int j = 0; j < n0; j++) for (int i = 0; i < m0; i++) S.H[j * mq + i] = D.H[j * m0 + i];
    if (inf) {  // infeasible.jl:8-16 : R ⊕ (R_inf*I/dt)
        double rinf = (ao->R_inf * 1.0) / D.dt;
        for (int i = 0; i < n0; i++) S.R[(m0 + i) * mq + (m0 + i)] = rinf;
    }
    S.R_mt = mt ? ao->R_minimum_time : 0.0;
    // rows
    S.rows.assign(D.N, std::vector<Row>());
    bool constrained = false;
    for (int k = 0; k < D.N; k++) {
        int cls = (D.n_classes > 0 && D.class_of_knot) ? D.class_of_knot[k] : -1;
        if (cls >= 0 && D.class_row_start[cls + 1] > D.class_row_start[cls]) constrained = true;
    }
    (void)constrained;
    const int nbar = S.n;  // column offset of controls in z̄
    for (int k = 0; k < D.N; k++) {
        const bool term = (k == D.N - 1);
        int cls = (D.n_classes > 0 && D.class_of_knot) ? D.class_of_knot[k] : -1;
        std::vector<Row> nonb, bxmax, bumax, bxmin, bumin;
        if (cls >= 0) {
            for (int i = D.class_row_start[cls]; i < D.class_row_start[cls + 1]; i++) {
                const TOConstraintRow& cr = D.rows[i];
                Row r;
                r.kind = cr.kind; r.eq = cr.equality != 0; r.sign = cr.sign; r.a = cr.a; r.b = cr.b; r.c = cr.c; r.r = cr.r;
                r.col = cr.var;
                bool isx = true;
                if (cr.kind == TO_ROW_LINEAR) {
                    isx = cr.var < n0;
                    r.col = isx ? cr.var : (cr.var - n0) + nbar;
                }
                if (cr.is_bound && cr.kind == TO_ROW_LINEAR) {
                    if (cr.sign > 0) (isx ? bxmax : bumax).push_back(r);
                    else (isx ? bxmin : bumin).push_back(r);
                } else {
                    nonb.push_back(r);
                }
            }
        }
        std::vector<Row>& out = S.rows[k];
        bool has_bound = !(bxmax.empty() && bumax.empty() && bxmin.empty() && bumin.empty());
        if (!inf && !mt) {
            // keep the caller's order untouched
            if (cls >= 0)
                for (int i = D.class_row_start[cls]; i < D.class_row_start[cls + 1]; i++) {
                    const TOConstraintRow& cr = D.rows[i];
                    Row r;
                    r.kind = cr.kind; r.eq = cr.equality != 0; r.sign = cr.sign; r.a = cr.a; r.b = cr.b; r.c = cr.c; r.r = cr.r;
                    r.col = cr.var;
                    out.push_back(r);
                }
            continue;
        }
        // ALTRO transforms: non-bound rows first, bounds after (constraint_sets.jl:135-150)
        out = nonb;
        // Both transforms in one solve (altro_methods.jl:98-124): minimum_time_problem runs on the slack problem, so the :infeasible
        // rows come BEFORE the re-combined bound (minimum_time.jl:132-141), and combine(bnd, mt_bnd) appends the sqrt(dt) bounds
        // after the ORIGINAL m0 controls of the knot's BoundConstraint (constraints.jl:195-203): they land on the first slack
        // control.  A knot without a BoundConstraint gets bnd0 of the slack problem's size, whose extra entry is sqrt(dt) itself.
        const bool both = inf && mt;
        const int tau_col = (both && has_bound) ? nbar + m0 : nbar + S.m - 1;
        if (both && !term) {
            for (int i = 0; i < n0; i++) {
                Row r; r.kind = R_LIN; r.eq = true; r.col = nbar + m0 + i; r.sign = 1.0; r.a = 0.0; r.b = r.c = r.r = 0;
                out.push_back(r);
            }
        }
        if (mt) {
            // minimum_time.jl:126-147: combine(bnd, mt_bnd): [x_max; u_max; sqrt(dt_max); x_min; u_min; sqrt(dt_min)]
            out.insert(out.end(), bxmax.begin(), bxmax.end());
            if (!term) {
                out.insert(out.end(), bumax.begin(), bumax.end());
                Row r; r.kind = R_LIN; r.eq = false; r.col = tau_col; r.sign = 1.0; r.a = std::sqrt(ao->dt_max); r.b = r.c = r.r = 0;
                out.push_back(r);
            }
            out.insert(out.end(), bxmin.begin(), bxmin.end());
            if (!term) {
                out.insert(out.end(), bumin.begin(), bumin.end());
                Row r; r.kind = R_LIN; r.eq = false; r.col = tau_col; r.sign = -1.0; r.a = std::sqrt(ao->dt_min); r.b = r.c = r.r = 0;
                out.push_back(r);
            }
            (void)has_bound;
        } else {
            out.insert(out.end(), bxmax.begin(), bxmax.end());
            if (!term) out.insert(out.end(), bumax.begin(), bumax.end());
            out.insert(out.end(), bxmin.begin(), bxmin.end());
            if (!term) out.insert(out.end(), bumin.begin(), bumin.end());
        }
        if (inf && !both && !term) {  // infeasible.jl:19-29 : + infeasible_constraints (constraints.jl:306-314)
            for (int i = 0; i < n0; i++) {
                Row r; r.kind = R_LIN; r.eq = true; r.col = nbar + m0 + i; r.sign = 1.0; r.a = 0.0; r.b = r.c = r.r = 0;
                out.push_back(r);
            }
        }
        if (mt && k > 0 && !term) {  // minimum_time.jl:142-144
            Row r; r.kind = R_MTEQ; r.eq = true; r.col = 0; r.sign = 1.0; r.a = r.b = r.c = r.r = 0;
            out.push_back(r);
        }
    }
}

// ------------------------------------------------------------------------------------------
// one problem, full ALTRO (altro_methods.jl:2-124)
// ------------------------------------------------------------------------------------------
struct Out {
    std::vector<vec> X, U;  // original dims
    vec dts;
    TOResult res;
    Trace tr;
    vec lam, mu;
    std::vector<uint8_t> act;
};

static void finish(const Ctx& c, const ALStats* st, Out& o) {
    o.res.status = c.status;
    o.res.steps = c.steps_total;
    if (st) {
        o.res.J = st->cost; o.res.c_max = st->c_max;
        o.res.iterations_total = st->iterations_total; o.res.iterations_outer = st->iterations;
    } else {
        o.res.J = c.s.last_cost; o.res.c_max = 0.0;
        o.res.iterations_total = c.s.iterations; o.res.iterations_outer = 0;
    }
}

static void pack_duals(const Ctx& c, Out& o) {
    o.lam.clear(); o.mu.clear(); o.act.clear();
    if (!c.al.on) return;
    for (int k = 0; k < c.S->N; k++) {
        o.lam.insert(o.lam.end(), c.al.lam[k].begin(), c.al.lam[k].end());
        o.mu.insert(o.mu.end(), c.al.mu[k].begin(), c.al.mu[k].end());
        o.act.insert(o.act.end(), c.al.act[k].begin(), c.al.act[k].end());
    }
}

enum Mode { MODE_ILQR = 0, MODE_AL = 1, MODE_ALTRO = 2 };

static void solve_one(const TOProblemDesc& D, int mode, const TOALTROOptions& ao, const double* x0, const double* U0,
                      const double* X0, bool want_trace, Out& out) {
    const int n0 = D.n, m0 = D.m, N = D.N;
    TOALOptions alo = ao.opts_al;
    bool inf = false, mt = false;
    const bool pn = (mode == MODE_ALTRO) && ao.projected_newton;
    if (mode == MODE_ALTRO) {
        inf = (X0 != nullptr);     // altro_methods.jl:102 (X[1] not all NaN)
        mt = (D.tf == 0.0);        // :111
        if (pn) {                  // :6-14
            if (ao.projected_newton_tolerance >= 0) alo.constraint_tolerance = ao.projected_newton_tolerance;
            else { alo.constraint_tolerance = 0; alo.kickout_max_penalty = 1; }
            if (mt) { std::fprintf(stderr, "oracle: projected Newton with minimum time is not supported (MinTimeCost has no hessian!)\n"); std::abort(); }
        }
    }
    Spec S;
    build_spec(D, inf, mt, &ao, S);
    Ctx c;
    c.S = &S;
    c.tr = want_trace ? &out.tr : nullptr;
    Prob& p = c.p;
    p.S = &S;
    p.x0.assign(x0, x0 + n0);
    p.X.assign(N, vec(S.n, std::numeric_limits<double>::quiet_NaN()));
    p.U.assign(N - 1, vec(S.m, 0.0));
    for (int k = 0; k < N - 1; k++) for (int i = 0; i < m0; i++) p.U[k][i] = U0[k * m0 + i];
    if (X0) for (int k = 0; k < N; k++) for (int i = 0; i < n0; i++) p.X[k][i] = X0[k * n0 + i];
    if (inf) {
        // slack_controls (infeasible.jl:62-80) on the ORIGINAL model with prob.dt
        Spec S0;
        build_spec(D, false, false, &ao, S0);
        vec x(p.x0), xn(n0);
        for (int k = 0; k < N - 1; k++) {
            dyn_eval(S0, xn.data(), x.data(), p.U[k].data());
            for (int i = 0; i < n0; i++) {
                double us = X0[(k + 1) * n0 + i] - xn[i];
                p.U[k][m0 + i] = us;
                xn[i] += us;
            }
            x = xn;
        }
    }
    if (mt) {  // minimum_time.jl:33-36
        double sdt = std::sqrt(D.dt);
        for (int k = 0; k < N - 1; k++) p.U[k][S.m - 1] = sdt;
        for (int k = 0; k < N; k++) p.X[k][S.n - 1] = sdt;  // [X[k]; sqrt(dt)]: with X[k] NaN it stays non-finite -> rollout
        p.x0.push_back(0.0);
    }
    ALStats st;
    bool used_al = false;
    bool ok = true;
    if (mode == MODE_ILQR) {
        c.al.on = false;
        ilqr_init(c, alo.opts_uncon);
        ok = ilqr_solve(c);
    } else if (mode == MODE_AL && !is_constrained(S)) {  // augmented_lagrangian_methods.jl:33-36
        c.al.on = false;
        ilqr_init(c, alo.opts_uncon);
        ok = ilqr_solve(c);
    } else {
        used_al = true;
        ok = al_solve(c, alo, st);
    }
    // projected Newton on prob_altro (altro_methods.jl:31-39)
    bool pn_done = false;
    double pn_J = 0.0, pn_cmax = 0.0;
    if (ok && pn) {
        PNOpts po{ao.pn_n_steps, ao.pn_feasibility_tolerance, ao.pn_active_set_tolerance};
        if (pn_solve(S, p, po, pn_J, pn_cmax)) pn_done = true;
        else c.status |= TO_STATUS_PN_FAILED;
    }
    // process_results! (altro_methods.jl:56-95)
    if (ok && mode == MODE_ALTRO && inf) {
        // infeasible_to_feasible_problem (infeasible.jl:38-59): copy first n / m, projection! (Q9) = open-loop rollout
        // with minimum time the slack-free problem is the minimum-time problem again (infeasible.jl:43-51)
        Spec S0;
        build_spec(D, false, mt, &ao, S0);
        Ctx c2;
        c2.S = &S0;
        c2.tr = c.tr;
        c2.status = c.status;
        c2.steps_total = c.steps_total;
        c2.p.S = &S0;
        c2.p.x0.assign(x0, x0 + n0);
        if (mt) c2.p.x0.push_back(0.0);
        const int n2 = S0.n, m2 = S0.m;
        c2.p.X.assign(N, vec(n2));
        c2.p.U.assign(N - 1, vec(m2));
        for (int k = 0; k < N; k++) for (int i = 0; i < n0; i++) c2.p.X[k][i] = p.X[k][i];
        for (int k = 0; k < N - 1; k++) for (int i = 0; i < m0; i++) c2.p.U[k][i] = p.U[k][i];
        if (mt) {  // sqrt(dt) of the first solve: controls, and the extra state of every knot but the first (infeasible.jl:46-50)
            for (int k = 0; k < N - 1; k++) c2.p.U[k][m2 - 1] = p.U[k][S.m - 1];
            for (int k = 0; k < N; k++) c2.p.X[k][n2 - 1] = (k == 0) ? 0.0 : p.X[k][S.n - 1];
        }
        if (ao.dynamically_feasible_projection) {
            // projection! (ilqr_methods.jl:179-190): K = d = 0 -> Ū = U, X̄ = rollout from x0; abort leaves the tail of X̄ zero
            std::vector<vec> Xb(N, vec(n2, 0.0));
            Xb[0] = c2.p.x0;
            for (int k = 1; k < N; k++) {
                dyn_eval(S0, Xb[k].data(), Xb[k - 1].data(), c2.p.U[k - 1].data());
                double mx = 0, mu = 0; bool bad = false;
                for (int i = 0; i < n2; i++) { double a = std::fabs(Xb[k][i]); if (a != a) bad = true; mx = std::max(mx, a); }
                for (int i = 0; i < m2; i++) { double a = std::fabs(c2.p.U[k - 1][i]); if (a != a) bad = true; mu = std::max(mu, a); }
                if (bad || !(mx < alo.opts_uncon.max_state_value && mu < alo.opts_uncon.max_control_value)) break;
            }
            c2.p.X = Xb;
        }
        if (ao.resolve_feasible_problem) {
            if (!is_constrained(S0)) {
                c2.al.on = false;
                ilqr_init(c2, alo.opts_uncon);
                ok = ilqr_solve(c2);
                used_al = false;
            } else {
                ALStats st2;
                c2.outer = 0;
                ok = al_solve(c2, alo, st2);
                st = st2;
                used_al = true;
            }
            out.X.assign(N, vec(n0));
            out.U.assign(N - 1, vec(m0));
            for (int k = 0; k < N; k++) for (int i = 0; i < n0; i++) out.X[k][i] = c2.p.X[k][i];
            for (int k = 0; k < N - 1; k++) for (int i = 0; i < m0; i++) out.U[k][i] = c2.p.U[k][i];
            out.dts.assign(N - 1, D.dt);
            if (mt) for (int k = 0; k < N - 1; k++) { double h = c2.p.U[k][m2 - 1]; out.dts[k] = h * h; }
            finish(c2, used_al ? &st : nullptr, out);
            pack_duals(c2, out);
            return;
        } else {
            // without the re-solve the projected copy is discarded (altro_methods.jl:67-78): the
            // caller keeps the first n states / m controls of the infeasible solution
            out.X.assign(N, vec(n0));
            out.U.assign(N - 1, vec(m0));
            for (int k = 0; k < N; k++) for (int i = 0; i < n0; i++) out.X[k][i] = p.X[k][i];
            for (int k = 0; k < N - 1; k++) for (int i = 0; i < m0; i++) out.U[k][i] = p.U[k][i];
            finish(c, used_al ? &st : nullptr, out);
            if (pn_done) { out.res.J = pn_J; out.res.c_max = pn_cmax; }
            pack_duals(c, out);
        }
        out.dts.assign(N - 1, D.dt);
        if (mt) for (int k = 0; k < N - 1; k++) { double h = p.U[k][S.m - 1]; out.dts[k] = h * h; }
        return;
    }
    out.X.assign(N, vec(n0));
    out.U.assign(N - 1, vec(m0));
    for (int k = 0; k < N; k++) for (int i = 0; i < n0; i++) out.X[k][i] = p.X[k][i];
    for (int k = 0; k < N - 1; k++) for (int i = 0; i < m0; i++) out.U[k][i] = p.U[k][i];
    out.dts.assign(N - 1, D.dt);
    if (mt) for (int k = 0; k < N - 1; k++) { double h = p.U[k][S.m - 1]; out.dts[k] = h * h; }
    finish(c, used_al ? &st : nullptr, out);
    if (pn_done) { out.res.J = pn_J; out.res.c_max = pn_cmax; }
    pack_duals(c, out);
}

}  // namespace orc

// ==========================================================================================
// C API for the tests / bench baseline
// ==========================================================================================
extern "C" {

void oracle_default_ilqr_options(TOiLQROptions* o) {
    o->cost_tolerance = 1e-4; o->gradient_norm_tolerance = 1e-5; o->iterations = 300; o->dJ_counter_limit = 10;
    o->square_root = 0; o->iterations_linesearch = 20; o->line_search_lower_bound = 1e-8; o->line_search_upper_bound = 10.0;
    o->bp_reg_increase_factor = 1.6; o->bp_reg_max = 1e8; o->bp_reg_min = 1e-8; o->bp_reg_fp = 10.0;
    o->max_cost_value = 1e8; o->max_state_value = 1e8; o->max_control_value = 1e8;
    o->bp_reg_type = TO_REG_CONTROL; o->gradient_type = TO_GRAD_TODOROV;
}
void oracle_default_al_options(TOALOptions* o) {
    oracle_default_ilqr_options(&o->opts_uncon);
    o->cost_tolerance = 1e-4; o->cost_tolerance_intermediate = 1e-3; o->gradient_norm_tolerance = 1e-5;
    o->gradient_norm_tolerance_intermediate = 1e-5; o->constraint_tolerance = 1e-3; o->iterations = 30;
    o->kickout_max_penalty = 0; o->dual_min = -1e8; o->dual_max = 1e8; o->penalty_max = 1e8; o->penalty_initial = 1.0;
    o->penalty_scaling = 10.0;
}
void oracle_default_altro_options(TOALTROOptions* o) {
    oracle_default_al_options(&o->opts_al);
    o->projected_newton = 0; o->pn_n_steps = 1; o->projected_newton_tolerance = 1e-3;
    o->pn_feasibility_tolerance = 1e-6; o->pn_active_set_tolerance = 1e-3;
    o->R_inf = 1.0; o->dynamically_feasible_projection = 1; o->resolve_feasible_problem = 1;
    o->R_minimum_time = 1.0; o->dt_max = 1.0; o->dt_min = 1e-3;
}

// mode: 0 iLQR, 1 AL, 2 ALTRO.  All arrays problem-major.  Outputs may be NULL.
// inner/outer traces: [B][cap] ; n_inner/n_outer [B].  threads: number of std::threads (>=1).
int oracle_solve(const TOProblemDesc* D, int mode, const TOALTROOptions* ao, int B, const double* x0, const double* U0,
                 const double* X0, double* X, double* U, double* dts, TOResult* results, TOIterRecord* inner,
                 int32_t* n_inner, int inner_cap, TOOuterRecord* outer, int32_t* n_outer, int outer_cap,
                 double* lambda, double* mu, uint8_t* active, int threads) {
    const int n = D->n, m = D->m, N = D->N;
    const bool want_trace = (inner && inner_cap > 0) || (outer && outer_cap > 0);
    auto work = [&](int b0, int b1) {
        for (int b = b0; b < b1; b++) {
            orc::Out out;
            orc::solve_one(*D, mode, *ao, x0 + (size_t)b * n, U0 + (size_t)b * (N - 1) * m,
                           X0 ? X0 + (size_t)b * N * n : nullptr, want_trace, out);
            if (X) for (int k = 0; k < N; k++) for (int i = 0; i < n; i++) X[((size_t)b * N + k) * n + i] = out.X[k][i];
            if (U) for (int k = 0; k < N - 1; k++) for (int i = 0; i < m; i++) U[((size_t)b * (N - 1) + k) * m + i] = out.U[k][i];
            if (dts) for (int k = 0; k < N - 1; k++) dts[(size_t)b * (N - 1) + k] = out.dts[k];
            if (inner && inner_cap > 0) {
                int cnt = (int)out.tr.inner.size();
                if (cnt > inner_cap) { cnt = inner_cap; out.res.status |= TO_STATUS_TRACE_TRUNC; }
                for (int i = 0; i < cnt; i++) inner[(size_t)b * inner_cap + i] = out.tr.inner[i];
                if (n_inner) n_inner[b] = cnt;
            }
            if (outer && outer_cap > 0) {
                int cnt = (int)out.tr.outer.size();
                if (cnt > outer_cap) { cnt = outer_cap; out.res.status |= TO_STATUS_TRACE_TRUNC; }
                for (int i = 0; i < cnt; i++) outer[(size_t)b * outer_cap + i] = out.tr.outer[i];
                if (n_outer) n_outer[b] = cnt;
            }
            if (results) results[b] = out.res;
            size_t P = out.lam.size();
            if (lambda) for (size_t i = 0; i < P; i++) lambda[(size_t)b * P + i] = out.lam[i];
            if (mu) for (size_t i = 0; i < P; i++) mu[(size_t)b * P + i] = out.mu[i];
            if (active) for (size_t i = 0; i < P; i++) active[(size_t)b * P + i] = out.act[i];
        }
    };
    if (threads <= 1) {
        work(0, B);
    } else {
        std::vector<std::thread> th;
        // interleaved chunks keep the threads balanced
        int chunk = (B + threads - 1) / threads;
        for (int t = 0; t < threads; t++) {
            int b0 = t * chunk, b1 = std::min(B, b0 + chunk);
            if (b0 < b1) th.emplace_back(work, b0, b1);
        }
        for (auto& t : th) t.join();
    }
    return 0;
}

// ---- primitive entry points for unit/known-answer tests ---------------------------------
void oracle_sincos(double x, double* s, double* c) { orc::sincos_det(x, s, c); }

// continuous dynamics and discrete step of a base model
void oracle_dynamics(int model, const double* x, const double* u, double* xdot) { orc::f_model<double>(model, xdot, x, u); }
void oracle_discrete(int model, int integ, const double* x, const double* u, double dt, double* xn) {
    orc::fd_model<double>(model, integ, orc::model_info(model).n, xn, x, u, dt);
}
// Jacobian n×(n+m+1) col-major of the discrete base model w.r.t. [x;u;dt]
void oracle_discrete_jacobian(int model, int integ, const double* x, const double* u, double dt, double* Z) {
    orc::Spec S;
    S.model = model; S.integ = integ;
    auto mi = orc::model_info(model);
    S.n0 = mi.n; S.m0 = mi.m;
    switch (mi.n + mi.m + 1) {
        case 4: orc::dyn_jac_inner<4>(S, x, u, dt, Z); break;
        case 6: orc::dyn_jac_inner<6>(S, x, u, dt, Z); break;
        case 7: orc::dyn_jac_inner<7>(S, x, u, dt, Z); break;
        case 18: orc::dyn_jac_inner<18>(S, x, u, dt, Z); break;
    }
}

// Evaluate the constraint rows / Jacobians / costs / expansions / augmented dynamics of the
// (possibly ALTRO-transformed) problem at knot k.  Used by the known-answer tests.
struct OracleSpecHandle { orc::Spec S; };
void* oracle_spec_create(const TOProblemDesc* D, int infeasible, int min_time, const TOALTROOptions* ao) {
    auto* h = new OracleSpecHandle();
    orc::build_spec(*D, infeasible != 0, min_time != 0, ao, h->S);
    return h;
}
void oracle_spec_destroy(void* h) { delete (OracleSpecHandle*)h; }
void oracle_spec_dims(void* h, int* n, int* m) { auto& S = ((OracleSpecHandle*)h)->S; *n = S.n; *m = S.m; }
int oracle_spec_num_rows(void* h, int k) { return (int)((OracleSpecHandle*)h)->S.rows[k].size(); }
void oracle_spec_constraints(void* h, int k, const double* x, const double* u, double* c, double* jac /* p×(n+m) row-major */,
                             int32_t* is_eq) {
    auto& S = ((OracleSpecHandle*)h)->S;
    auto& rows = S.rows[k];
    int nz = S.n + S.m;
    for (size_t i = 0; i < rows.size(); i++) {
        if (c) c[i] = orc::row_value(S, rows[i], x, u);
        if (jac) {
            for (int j = 0; j < nz; j++) jac[i * nz + j] = 0.0;
            orc::row_jac(S, rows[i], x, jac + i * nz);
        }
        if (is_eq) is_eq[i] = rows[i].eq ? 1 : 0;
    }
}
double oracle_spec_stage_cost(void* h, const double* x, const double* u) { return orc::stage_cost(((OracleSpecHandle*)h)->S, x, u); }
double oracle_spec_term_cost(void* h, const double* x) { return orc::term_cost(((OracleSpecHandle*)h)->S, x); }
void oracle_spec_stage_expansion(void* h, const double* x, const double* u, double* Qx, double* Qu, double* Qxx, double* Quu, double* Qux) {
    auto& S = ((OracleSpecHandle*)h)->S;
    orc::Exp E; E.init(S.n, S.m);
    orc::stage_expansion(S, E, x, u);
    std::memcpy(Qx, E.x.data(), 8 * S.n); std::memcpy(Qu, E.u.data(), 8 * S.m);
    std::memcpy(Qxx, E.xx.data(), 8 * S.n * S.n); std::memcpy(Quu, E.uu.data(), 8 * S.m * S.m); std::memcpy(Qux, E.ux.data(), 8 * S.m * S.n);
}
void oracle_spec_term_expansion(void* h, const double* x, double* Qx, double* Qxx) {
    auto& S = ((OracleSpecHandle*)h)->S;
    orc::Exp E; E.init(S.n, S.m);
    orc::term_expansion(S, E, x);
    std::memcpy(Qx, E.x.data(), 8 * S.n); std::memcpy(Qxx, E.xx.data(), 8 * S.n * S.n);
}
void oracle_spec_dynamics(void* h, const double* x, const double* u, double* xn, double* A, double* B) {
    auto& S = ((OracleSpecHandle*)h)->S;
    if (xn) orc::dyn_eval(S, xn, x, u);
    if (A && B) orc::dyn_jac(S, x, u, A, B);
}

// One backward pass (standard or sqrt) on given A,B,Q trajectories: used for the sqrt≡standard
// equivalence test (test/sqrt_bp_tests.jl:30-44).  Layout: per knot contiguous, col-major blocks.
int oracle_backwardpass(int n, int m, int N, int square_root, double rho_in, const double* A, const double* Bm, const double* Qx,
                        const double* Qu, const double* Qxx, const double* Quu, const double* Qux, const double* QNx,
                        const double* QNxx, double* K, double* d, double* Sxx1, double* Sx1, double* dV) {
    orc::Spec S; S.n = n; S.m = m; S.N = N;
    orc::Ctx c; c.S = &S;
    TOiLQROptions o; oracle_default_ilqr_options(&o); o.square_root = square_root;
    orc::ilqr_init(c, o);
    c.s.rho = rho_in;
    for (int k = 0; k < N - 1; k++) {
        c.s.A[k].assign(A + (size_t)k * n * n, A + (size_t)(k + 1) * n * n);
        c.s.B[k].assign(Bm + (size_t)k * n * m, Bm + (size_t)(k + 1) * n * m);
        c.s.Q[k].x.assign(Qx + (size_t)k * n, Qx + (size_t)(k + 1) * n);
        c.s.Q[k].u.assign(Qu + (size_t)k * m, Qu + (size_t)(k + 1) * m);
        c.s.Q[k].xx.assign(Qxx + (size_t)k * n * n, Qxx + (size_t)(k + 1) * n * n);
        c.s.Q[k].uu.assign(Quu + (size_t)k * m * m, Quu + (size_t)(k + 1) * m * m);
        c.s.Q[k].ux.assign(Qux + (size_t)k * m * n, Qux + (size_t)(k + 1) * m * n);
    }
    c.s.Q[N - 1].x.assign(QNx, QNx + n);
    c.s.Q[N - 1].xx.assign(QNxx, QNxx + n * n);
    bool ok = true;
    if (square_root) {
        for (int k = 0; k < N - 1; k++) {
            if (!orc::chol_upper_inplace(c.s.Q[k].xx.data(), n)) return -1;
            if (!orc::chol_upper_inplace(c.s.Q[k].uu.data(), m)) return -1;
        }
        if (!orc::chol_upper_inplace(c.s.Q[N - 1].xx.data(), n)) return -1;
        ok = orc::backwardpass_sqrt(S, c.s, dV);
    } else {
        orc::backwardpass(S, c.s, dV);
    }
    for (int k = 0; k < N - 1; k++) {
        std::memcpy(K + (size_t)k * m * n, c.s.K[k].data(), 8 * m * n);
        std::memcpy(d + (size_t)k * m, c.s.d[k].data(), 8 * m);
    }
    std::memcpy(Sxx1, c.s.Sxx[0].data(), 8 * n * n);
    std::memcpy(Sx1, c.s.Sx[0].data(), 8 * n);
    return ok ? 0 : -2;
}

int oracle_hw_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
