// ORACLE — TEST INFRASTRUCTURE ONLY (included by oracle.cpp).
//
// Projected-Newton polish of ALTRO: the `solve_type = :feasible` path of src/solvers/direct/projected_newton.jl, i.e.
// newton_step! (:500-520) -> update! (:154-165) -> projection_solve! (:206-218) -> _projection_solve! (:221-270) ->
// _projection_linesearch! (:272-312) -> reg_solve (:314-333), hooked into ALTRO at altro/altro_methods.jl:6-14,31-39.
//
// The reference assembles the sparse constraint Jacobian Y of [x1 - x0; f(x_k,u_k) - x_{k+1}; active c_k ...] over
// Z = [x1,u1,...,xN] (direct_solvers.jl:62-110), forms S = Y H^-1 Y' with H = Diagonal(cost Hessian) and factorises S + 1e-2 I
// with CHOLMOD.  S is block tridiagonal (block s couples knot s-1 and the state of knot s); this restatement factorises it
// block by block with dense Cholesky -- the same matrix, another elimination order, so results agree with the reference to
// rounding (not bit for bit: "parity unpinned" at bit level, and no committed reference artefact pins the PN iterates; the
// tests pin the properties: feasibility below the tolerance, a cost within the AL solution's neighbourhood, same active rows).
#pragma once

struct PNOpts {
    int n_steps;        // direct_solvers.jl:19
    double feas_tol;    // feasibility_tolerance :28
    double active_tol;  // active_set_tolerance :25
};

struct PNWork {
    // per knot
    std::vector<vec> A, B;     // n*n, n*m col-major
    std::vector<vec> fv;       // N entries: fv[0] = X[0]-x0, fv[k+1] = f(x_k,u_k) - x_{k+1}
    std::vector<vec> C;        // constraint values
    std::vector<std::vector<uint8_t>> act;
    std::vector<vec> hx, hu;   // diagonal of the cost Hessian (projected_newton.jl:122-148 with cost.jl:214-228)
    // blocks s = 0..N
    std::vector<int> r;        // rows of block s
    std::vector<vec> G;        // r x w row-major, w = n (s = 0), n+m+n (stage blocks), n (s = N)
    std::vector<vec> Sd, So;   // S_ss (r x r), S_{s,s-1} (r_s x r_{s-1}), row-major
    std::vector<vec> Ld, Lo;   // Cholesky blocks of S + rho I
    std::vector<vec> y;        // active residual per block
};

// dynamics_constraints! (:34-42) + update_constraints! (:67-73): everything is evaluated, `y` keeps the active rows
static void pn_residuals(const Spec& S, const Prob& p, const std::vector<vec>& X, const std::vector<vec>& U, PNWork& w) {
    const int n = S.n, N = S.N;
    for (int i = 0; i < n; i++) w.fv[0][i] = X[0][i] - p.x0[i];
    for (int k = 0; k < N - 1; k++) {
        dyn_eval(S, w.fv[k + 1].data(), X[k].data(), U[k].data());
        for (int i = 0; i < n; i++) w.fv[k + 1][i] -= X[k + 1][i];
    }
    for (int k = 0; k < N; k++) {
        const double* u = (k < N - 1) ? U[k].data() : nullptr;
        for (size_t i = 0; i < S.rows[k].size(); i++) w.C[k][i] = row_value(S, S.rows[k][i], X[k].data(), u);
    }
}
// active_set! (:75-92): equality rows always, inequality rows with c >= -tol
static void pn_active_set(const Spec& S, PNWork& w, double tol) {
    for (int k = 0; k < S.N; k++)
        for (size_t i = 0; i < S.rows[k].size(); i++) w.act[k][i] = S.rows[k][i].eq ? 1 : (w.C[k][i] >= -tol);
}
// y = solver.y[a.duals] per block; returns norm(y, Inf)
static double pn_gather_y(const Spec& S, PNWork& w) {
    const int n = S.n, N = S.N;
    double viol = 0.0;
    for (int s = 0; s <= N; s++) {
        vec& y = w.y[s];
        y.clear();
        if (s < N) for (int i = 0; i < n; i++) y.push_back(w.fv[s][i]);  // x1 - x0, or f(x_{s-1},u_{s-1}) - x_s
        if (s >= 1) {                                                    // active rows of knot s-1 (s = N: the terminal set)
            const int k = s - 1;
            for (size_t i = 0; i < S.rows[k].size(); i++) if (w.act[k][i]) y.push_back(w.C[k][i]);
        }
        for (double v : y) viol = std::max(viol, std::fabs(v));
    }
    return viol;
}

// dense Cholesky (lower, row-major, in place); false if a pivot is not positive
static bool pn_chol(vec& A, int r) {
    for (int j = 0; j < r; j++) {
        double d = A[j * r + j];
        for (int l = 0; l < j; l++) d = fmad(-A[j * r + l], A[j * r + l], d);
        if (!(d > 0.0)) return false;
        d = std::sqrt(d);
        A[j * r + j] = d;
        for (int i = j + 1; i < r; i++) {
            double v = A[i * r + j];
            for (int l = 0; l < j; l++) v = fmad(-A[i * r + l], A[j * r + l], v);
            A[i * r + j] = v / d;
        }
    }
    return true;
}

// dynamics_jacobian! (:45-64) + constraint_jacobian! (:98-107) -> per-block rows of Y; S = Y H^-1 Y' (:246-247); Cholesky of S + rho I
static bool pn_factor(const Spec& S, const std::vector<vec>& X, const std::vector<vec>& U, PNWork& w, double rho) {
    const int n = S.n, m = S.m, N = S.N;
    for (int k = 0; k < N - 1; k++) dyn_jac(S, X[k].data(), U[k].data(), w.A[k].data(), w.B[k].data());
    // blocks
    for (int s = 0; s <= N; s++) {
        vec& G = w.G[s];
        if (s == 0) {
            w.r[s] = n;
            G.assign((size_t)n * n, 0.0);
            for (int i = 0; i < n; i++) G[i * n + i] = 1.0;
        } else if (s < N) {
            const int k = s - 1, wd = n + m + n;
            int pa = 0;
            for (size_t i = 0; i < S.rows[k].size(); i++) pa += w.act[k][i];
            w.r[s] = n + pa;
            G.assign((size_t)w.r[s] * wd, 0.0);
            for (int i = 0; i < n; i++) {
                for (int c = 0; c < n; c++) G[i * wd + c] = w.A[k][c * n + i];
                for (int c = 0; c < m; c++) G[i * wd + n + c] = w.B[k][c * n + i];
                G[i * wd + n + m + i] = -1.0;
            }
            int row = n;
            vec g(n + m);
            for (size_t i = 0; i < S.rows[k].size(); i++) {
                if (!w.act[k][i]) continue;
                std::fill(g.begin(), g.end(), 0.0);
                row_jac(S, S.rows[k][i], X[k].data(), g.data());
                for (int c = 0; c < n + m; c++) G[row * wd + c] = g[c];
                row++;
            }
        } else {
            const int k = N - 1;
            int pa = 0;
            for (size_t i = 0; i < S.rows[k].size(); i++) pa += w.act[k][i];
            w.r[s] = pa;
            G.assign((size_t)pa * n, 0.0);
            int row = 0;
            vec g(n + m);
            for (size_t i = 0; i < S.rows[k].size(); i++) {
                if (!w.act[k][i]) continue;
                std::fill(g.begin(), g.end(), 0.0);
                row_jac(S, S.rows[k][i], X[k].data(), g.data());
                for (int c = 0; c < n; c++) G[row * n + c] = g[c];
                row++;
            }
        }
    }
    // S blocks: S_ss = G_s D^-1 G_s', S_{s,s-1} couples through the state both blocks touch
    for (int s = 0; s <= N; s++) {
        const int r = w.r[s];
        const vec& G = w.G[s];
        w.Sd[s].assign((size_t)r * r, 0.0);
        const int wd = (s == 0 || s == N) ? n : n + m + n;
        auto hcc = [&](int c) {  // H_cc of column c of block s
            if (s == 0) return w.hx[0][c];
            if (s == N) return w.hx[N - 1][c];
            const int k = s - 1;
            return (c < n) ? w.hx[k][c] : ((c < n + m) ? w.hu[k][c - n] : w.hx[k + 1][c - n - m]);
        };
        for (int i = 0; i < r; i++)
            for (int j = 0; j < r; j++) {
                double acc = 0.0;
                for (int c = 0; c < wd; c++) acc = fmad(G[i * wd + c], G[j * wd + c] / hcc(c), acc);  // Y * (H \ Y')
                w.Sd[s][i * r + j] = acc;
            }
        if (s >= 1) {
            const int rp = w.r[s - 1];
            w.So[s].assign((size_t)r * rp, 0.0);
            // the shared state: x of knot s-1.  In block s it is columns 0..n-1 (z_k of a stage block, x_N of the last one);
            // in block s-1 it is the next-state columns (-I on the dynamics rows) or, for s-1 = 0, the identity block.
            for (int i = 0; i < r; i++)
                for (int j = 0; j < n && j < rp; j++) {
                    const double gprev = (s - 1 == 0) ? 1.0 : -1.0;
                    w.So[s][i * rp + j] = G[i * wd + j] * (gprev / w.hx[s - 1][j]);
                }
        }
    }
    // block Cholesky of S + rho I
    for (int s = 0; s <= N; s++) {
        const int r = w.r[s];
        vec& L = w.Ld[s];
        L = w.Sd[s];
        for (int i = 0; i < r; i++) L[i * r + i] += rho;
        if (s >= 1) {
            const int rp = w.r[s - 1];
            // L_{s,s-1} = S_{s,s-1} L_{s-1,s-1}^-T  (row by row: solve L_prev x' = So_row')
            vec& Lo = w.Lo[s];
            Lo.assign((size_t)r * rp, 0.0);
            const vec& Lp = w.Ld[s - 1];
            for (int i = 0; i < r; i++)
                for (int j = 0; j < rp; j++) {
                    double v = w.So[s][i * rp + j];
                    for (int l = 0; l < j; l++) v = fmad(-Lo[i * rp + l], Lp[j * rp + l], v);
                    Lo[i * rp + j] = v / Lp[j * rp + j];
                }
            for (int i = 0; i < r; i++)
                for (int j = 0; j <= i; j++) {
                    double acc = 0.0;
                    for (int l = 0; l < rp; l++) acc = fmad(Lo[i * rp + l], Lo[j * rp + l], acc);
                    L[i * r + j] -= acc;
                    if (j != i) L[j * r + i] = L[i * r + j];
                }
        }
        if (!pn_chol(L, r)) return false;
    }
    return true;
}

// x = (S + rho I)^-1 b with the block factor
static void pn_solve_factor(const PNWork& w, int N, const std::vector<vec>& b, std::vector<vec>& x) {
    std::vector<vec> t(N + 1);
    for (int s = 0; s <= N; s++) {
        const int r = w.r[s];
        t[s] = b[s];
        if (s >= 1) {
            const int rp = w.r[s - 1];
            for (int i = 0; i < r; i++) {
                double acc = 0.0;
                for (int l = 0; l < rp; l++) acc = fmad(w.Lo[s][i * rp + l], t[s - 1][l], acc);
                t[s][i] -= acc;
            }
        }
        const vec& L = w.Ld[s];
        for (int i = 0; i < r; i++) {
            double v = t[s][i];
            for (int l = 0; l < i; l++) v = fmad(-L[i * r + l], t[s][l], v);
            t[s][i] = v / L[i * r + i];
        }
    }
    x.assign(N + 1, vec());
    for (int s = N; s >= 0; s--) {
        const int r = w.r[s];
        x[s] = t[s];
        if (s < N) {
            const int rn = w.r[s + 1];
            for (int i = 0; i < r; i++) {
                double acc = 0.0;
                for (int l = 0; l < rn; l++) acc = fmad(w.Lo[s + 1][l * r + i], x[s + 1][l], acc);
                x[s][i] -= acc;
            }
        }
        const vec& L = w.Ld[s];
        for (int i = r - 1; i >= 0; i--) {  // sums run from the last unknown down (the order a parallel wavefront produces them)
            double v = x[s][i];
            for (int l = r - 1; l > i; l--) v = fmad(-L[l * r + i], x[s][l], v);
            x[s][i] = v / L[i * r + i];
        }
    }
}
// out = S x
static void pn_apply_S(const PNWork& w, int N, const std::vector<vec>& x, std::vector<vec>& out) {
    out.assign(N + 1, vec());
    for (int s = 0; s <= N; s++) {
        const int r = w.r[s];
        out[s].assign(r, 0.0);
        for (int i = 0; i < r; i++) {
            double acc = 0.0;
            if (s >= 1) {
                const int rp = w.r[s - 1];
                for (int l = 0; l < rp; l++) acc = fmad(w.So[s][i * rp + l], x[s - 1][l], acc);
            }
            for (int l = 0; l < r; l++) acc = fmad(w.Sd[s][i * r + l], x[s][l], acc);
            if (s < N) {
                const int rn = w.r[s + 1];
                for (int l = 0; l < rn; l++) acc = fmad(w.So[s + 1][l * r + i], x[s + 1][l], acc);
            }
            out[s][i] = acc;
        }
    }
}

// reg_solve(S, y, Sreg, 1e-8, 25) (:314-333): iterative refinement with the regularised factor
static void pn_reg_solve(const PNWork& w, int N, const std::vector<vec>& b, std::vector<vec>& x) {
    pn_solve_factor(w, N, b, x);
    std::vector<vec> Sx, rr(N + 1), dx;
    for (int count = 0; count < 25; count++) {
        pn_apply_S(w, N, x, Sx);
        double nrm = 0.0;
        for (int s = 0; s <= N; s++) {
            rr[s].assign(w.r[s], 0.0);
            for (int i = 0; i < w.r[s]; i++) { rr[s][i] = b[s][i] - Sx[s][i]; nrm += rr[s][i] * rr[s][i]; }
        }
        if (std::sqrt(nrm) < 1e-8) break;
        pn_solve_factor(w, N, rr, dx);
        for (int s = 0; s <= N; s++) for (int i = 0; i < w.r[s]; i++) x[s][i] += dx[s][i];
    }
}

// _projection_linesearch! (:272-312).  Returns the new violation; `threw` is set where the reference would raise
// (`count += a` adds a BitVector to an Int when the full step does not reduce the violation, :304)
static double pn_linesearch(const Spec& S, const Prob& p, std::vector<vec>& X, std::vector<vec>& U, PNWork& w, bool& threw) {
    const int n = S.n, m = S.m, N = S.N;
    const double viol0 = pn_gather_y(S, w);
    std::vector<vec> dl;
    pn_reg_solve(w, N, w.y, dl);
    // dZ = -H^-1 Y' dlambda
    std::vector<vec> Xn = X, Un = U;
    for (int k = 0; k < N; k++) {
        for (int c = 0; c < n; c++) {
            double acc = 0.0;
            // block k (as the NEXT state of the dynamics rows, or the identity rows of block 0)
            if (k == 0) acc = fmad(1.0, dl[0][c], acc);
            else acc = fmad(-1.0, dl[k][c], acc);
            // block k+1 (as x_k of z_k, or x_N of the terminal block)
            const int s = k + 1;
            const int wd = (s == N) ? n : n + m + n;
            for (int i = 0; i < w.r[s]; i++) acc = fmad(w.G[s][i * wd + c], dl[s][i], acc);
            Xn[k][c] = X[k][c] + 1.0 * (-(acc / w.hx[k][c]));
        }
        if (k < N - 1)
            for (int c = 0; c < m; c++) {
                double acc = 0.0;
                const int s = k + 1, wd = n + m + n;
                for (int i = 0; i < w.r[s]; i++) acc = fmad(w.G[s][i * wd + n + c], dl[s][i], acc);
                Un[k][c] = U[k][c] + 1.0 * (-(acc / w.hu[k][c]));
            }
    }
    pn_residuals(S, p, Xn, Un, w);
    const double viol = pn_gather_y(S, w);
    if (!(viol < viol0)) { threw = true; return viol; }
    X = Xn; U = Un;
    return viol;
}

// _projection_solve! (:221-270)
static double pn_projection_once(const Spec& S, const Prob& p, std::vector<vec>& X, std::vector<vec>& U, PNWork& w, const PNOpts& o,
                                 bool& threw) {
    pn_residuals(S, p, X, U, w);
    pn_active_set(S, w, o.active_tol);
    double viol_prev = pn_gather_y(S, w);
    if (!pn_factor(S, X, U, w, 1e-2)) { threw = true; return viol_prev; }  // PosDefException in cholesky(S + rho I)
    for (int count = 0; count < 10; count++) {
        const double viol = pn_linesearch(S, p, X, U, w, threw);
        if (threw) return viol;
        const double rate = std::log10(viol) / std::log10(viol_prev);
        viol_prev = viol;
        if (rate < 1.1 || viol < o.feas_tol) break;
    }
    return viol_prev;
}

// max_violation(prob) (src/problem.jl:242-267)
static double pn_max_violation(const Spec& S, const std::vector<vec>& X, const std::vector<vec>& U) {
    double cmax = 0.0;
    for (int k = 0; k < S.N; k++) {
        if (S.rows[k].empty()) continue;
        const double* u = (k < S.N - 1) ? U[k].data() : nullptr;
        double e = 0.0, mi = 0.0;
        for (const Row& r : S.rows[k]) {
            const double c = row_value(S, r, X[k].data(), u);
            if (r.eq) e = std::max(e, std::fabs(c));
            mi = std::max(mi, std::max(0.0, c));  // maximum(pos.(c)) runs over every row, equalities included
        }
        cmax = std::max(cmax, std::max(e, mi));
    }
    return cmax;
}

// solve!(prob, ::ProjectedNewtonSolver) (:4-19) with solve_type = :feasible.  Returns false where the reference throws
// (X, U untouched then: the copy back to prob happens after the Newton step, :8-11).
static bool pn_solve(const Spec& S, Prob& p, const PNOpts& o, double& J, double& cmax) {
    const int n = S.n, m = S.m, N = S.N;
    PNWork w;
    w.A.assign(N - 1, vec((size_t)n * n)); w.B.assign(N - 1, vec((size_t)n * m));
    w.fv.assign(N, vec(n)); w.C.resize(N); w.act.resize(N);
    for (int k = 0; k < N; k++) { w.C[k].assign(S.rows[k].size(), 0.0); w.act[k].assign(S.rows[k].size(), 1); }
    w.hx.assign(N, vec(n, 0.0)); w.hu.assign(N - 1, vec(m, 0.0));
    for (int k = 0; k < N - 1; k++) {  // hessian!(.., dt) = [Q H'; H R] * dt (cost.jl:214-223); only the diagonal is used (:236)
        for (int i = 0; i < S.nq; i++) w.hx[k][i] = S.Q[i * S.nq + i] * S.dt;
        for (int i = 0; i < S.mq; i++) w.hu[k][i] = S.R[i * S.mq + i] * S.dt;
    }
    for (int i = 0; i < S.nq; i++) w.hx[N - 1][i] = S.Qf[i * S.nq + i];
    w.r.assign(N + 1, 0); w.G.resize(N + 1); w.Sd.resize(N + 1); w.So.resize(N + 1); w.Ld.resize(N + 1); w.Lo.resize(N + 1);
    w.y.resize(N + 1);
    std::vector<vec> X = p.X, U = p.U;
    for (int step = 0; step < o.n_steps; step++) {
        // newton_step! -> update! (values, Jacobians, active set at the current point), then projection_solve! (:206-218)
        pn_residuals(S, p, X, U, w);
        pn_active_set(S, w, o.active_tol);
        double viol = pn_gather_y(S, w);
        bool threw = false;
        for (int count = 0; count < 10 && viol > o.feas_tol; count++) {
            viol = pn_projection_once(S, p, X, U, w, o, threw);
            if (threw) return false;
        }
        p.X = X; p.U = U;
        // record_iteration! (:21-29)
        J = 0.0;
        for (int k = 0; k < N - 1; k++) J += stage_cost(S, X[k].data(), U[k].data());
        J += term_cost(S, X[N - 1].data());
        cmax = pn_max_violation(S, X, U);
        if (cmax <= o.feas_tol) break;
    }
    return true;
}
