// Square-root backward pass of the lockstep engine (backward_pass.jl:87-192, objective.jl:70-94,
// augmented_lagrangian_methods.jl:231-276): the cost-to-go and the action-value Hessians are carried as
// upper-triangular factors, updated with QR "up-dates" (chol_plus) and rank-one down-dates (chol_minus).
//
// Two kernels with identical results.  One THREAD per problem (SqrtBp): the factor updates are chains of small Householder
// reflections with data-dependent branches (cond() by one-sided Jacobi, PosDefException exits), i.e. serial per problem; all
// matrices are column-major thread-local arrays of compile-time size, which the compiler keeps in registers for the small
// models (BASELINE config 5: n <= 5).  One WARP per problem (SqrtWarp, further down): matrices in shared memory, one lane per
// output element -- for the models whose matrices would be a local-memory frame (quadrotor: 22 KB) when few problems are live.
// Every output follows the operation order of the CPU restatement in both, so the results are bit-identical to it.
#pragma once
#include "engine.cuh"

namespace tob {

// `#pragma unroll (UNR)`: the helpers are size-generic; the recursion instantiates them with UNR = UF (full unrolling for the small
// models, whose matrices then live in registers instead of local memory), the expansion with UNR = 1 (runtime row counts)
#define SQ_UNROLL _Pragma("unroll (UNR)")

template <class C>
struct SqrtBp {
    static constexpr int n = C::n, m = C::m, nq = C::nq, mq = C::mq, nz = C::n + C::m;
    static constexpr int PMAX = 32;                       // constraint rows per knot supported by this path
    static constexpr int RMAX = n + ((n > PMAX) ? n : PMAX);  // rows of the tallest stacked matrix
    static constexpr int UF = (n <= 6 && m <= 6) ? 64 : 1;     // unroll factor of the recursion's helper loops

    // cholesky(A).U, dot-product (left-looking) form; false = not positive definite
    template <int D, int UNR = 1>
    static __device__ __forceinline__ bool chol_upper(const double* A, double* U) {
        SQ_UNROLL
        for (int j = 0; j < D; j++) {
            SQ_UNROLL
            for (int i = 0; i < j; i++) {
                double acc = 0.0;
                SQ_UNROLL
                for (int l = 0; l < i; l++) acc = fma(U[i * D + l], U[j * D + l], acc);
                U[j * D + i] = (A[j * D + i] - acc) / U[i * D + i];
            }
            double acc = 0.0;
            SQ_UNROLL
            for (int l = 0; l < j; l++) acc = fma(U[j * D + l], U[j * D + l], acc);
            const double dd = A[j * D + j] - acc;
            if (!(dd > 0.0)) return false;
            U[j * D + j] = sqrt(dd);
            SQ_UNROLL
            for (int i = j + 1; i < D; i++) U[j * D + i] = 0.0;
        }
        return true;
    }
    template <int D, int UNR = 1>
    static __device__ __forceinline__ bool chol_upper_inplace(double* A) {
        double U[D * D];
        SQ_UNROLL
        for (int e = 0; e < D * D; e++) U[e] = 0.0;
        if (!chol_upper<D, UNR>(A, U)) return false;
        SQ_UNROLL
        for (int e = 0; e < D * D; e++) A[e] = U[e];
        return true;
    }

    // Householder QR of the rows×D matrix P (column-major, leading dimension `rows`), in place: R ends up in the
    // top D×D block (LAPACK dgeqr2 / dlarfg; R may have negative diagonal entries)
    template <int D, int UNR = 1>
    static __device__ __forceinline__ void qr_R(double* P, int rows) {
        SQ_UNROLL
        for (int j = 0; j < D && j < rows; j++) {
            const double alpha = P[j * rows + j];
            double xn2 = 0.0;
            SQ_UNROLL
            for (int i = j + 1; i < rows; i++) xn2 = fma(P[j * rows + i], P[j * rows + i], xn2);
            double tau = 0.0;
            if (xn2 != 0.0) {
                const double xnorm = sqrt(xn2);
                const double beta = -copysign(sqrt(alpha * alpha + xnorm * xnorm), alpha);
                tau = (beta - alpha) / beta;
                const double sc = 1.0 / (alpha - beta);
                SQ_UNROLL
                for (int i = j + 1; i < rows; i++) P[j * rows + i] = P[j * rows + i] * sc;
                P[j * rows + j] = beta;
            }
            if (tau != 0.0) {
                SQ_UNROLL
                for (int c = j + 1; c < D; c++) {
                    double w = P[c * rows + j];
                    SQ_UNROLL
                    for (int i = j + 1; i < rows; i++) w = fma(P[j * rows + i], P[c * rows + i], w);
                    const double tw = tau * w;
                    P[c * rows + j] = P[c * rows + j] - tw;
                    SQ_UNROLL
                    for (int i = j + 1; i < rows; i++) P[c * rows + i] = fma(-tw, P[j * rows + i], P[c * rows + i]);
                }
            }
        }
    }
    // R = qr([A; B]).R with A: D×D (upper factor, full storage), B: nb×D (leading dimension nb).  `Pbuf`: (D+nb)*D doubles
    template <int D, int UNR = 1>
    static __device__ __forceinline__ void chol_plus(const double* A, const double* Bm, int nb, double* Rout, double* Pbuf) {
        const int rows = D + nb;
        SQ_UNROLL
        for (int j = 0; j < D; j++) {
            SQ_UNROLL
            for (int i = 0; i < D; i++) Pbuf[j * rows + i] = A[j * D + i];
            SQ_UNROLL
            for (int i = 0; i < nb; i++) Pbuf[j * rows + D + i] = Bm[j * nb + i];
        }
        qr_R<D, UNR>(Pbuf, rows);
        SQ_UNROLL
        for (int j = 0; j < D; j++)
            SQ_UNROLL
            for (int i = 0; i < D; i++) Rout[j * D + i] = (i <= j) ? Pbuf[j * rows + i] : 0.0;
    }
    // lowrankdowndate!(Cholesky(U,:U), v); false on PosDefException
    template <int D, int UNR = 1>
    static __device__ __forceinline__ bool lowrank_downdate(double* U, double* v) {
        SQ_UNROLL
        for (int i = 0; i < D; i++) {
            const double Aii = U[i * D + i];
            const double s = v[i] / Aii;
            const double s2 = s * s;
            if (s2 > 1.0) return false;
            const double c = sqrt(1.0 - s2);
            U[i * D + i] = c * Aii;
            SQ_UNROLL
            for (int j = i + 1; j < D; j++) {
                const double vj = v[j];
                const double Aij = (U[j * D + i] - s * vj) / c;
                U[j * D + i] = Aij;
                v[j] = -s * Aij + c * vj;
            }
        }
        return true;
    }
    // cond(A): ratio of the extreme singular values by one-sided Jacobi (fixed sweep order)
    template <int D, int UNR = 1>
    static __device__ __forceinline__ double cond2(const double* Ain) {
        double A[D * D];
        SQ_UNROLL
        for (int e = 0; e < D * D; e++) A[e] = Ain[e];
        for (int sweep = 0; sweep < 60; sweep++) {
            double off = 0.0;
            SQ_UNROLL
            for (int p = 0; p < D - 1; p++)
                SQ_UNROLL
                for (int q = p + 1; q < D; q++) {
                    double a = 0, b = 0, c = 0;
                    SQ_UNROLL
                    for (int i = 0; i < D; i++) {
                        a = fma(A[p * D + i], A[p * D + i], a);
                        b = fma(A[q * D + i], A[q * D + i], b);
                        c = fma(A[p * D + i], A[q * D + i], c);
                    }
                    if (c == 0.0) continue;
                    off = dmax(off, fabs(c) / sqrt(a * b));
                    const double zeta = (b - a) / (2.0 * c);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                    SQ_UNROLL
                    for (int i = 0; i < D; i++) {
                        const double ap = A[p * D + i], aq = A[q * D + i];
                        A[p * D + i] = cs * ap - sn * aq;
                        A[q * D + i] = sn * ap + cs * aq;
                    }
                }
            if (off < 1e-15) break;
        }
        double smax = 0.0, smin = __longlong_as_double(0x7ff0000000000000LL);
        SQ_UNROLL
        for (int j = 0; j < D; j++) {
            double a = 0;
            SQ_UNROLL
            for (int i = 0; i < D; i++) a = fma(A[j * D + i], A[j * D + i], a);
            const double sv = sqrt(a);
            smax = dmax(smax, sv);
            smin = dmin(smin, sv);
        }
        return smax / smin;
    }
    // X = U \ B and X = U' \ B for an upper-triangular D×D U, nrhs right-hand sides (column-major, ld D)
    template <int D, int UNR = 1>
    static __device__ __forceinline__ void solve_upper(const double* U, const double* Bin, int nrhs, double* X) {
        SQ_UNROLL
        for (int c = 0; c < nrhs; c++)
            SQ_UNROLL
            for (int i = D - 1; i >= 0; i--) {
                double acc = 0.0;
                SQ_UNROLL
                for (int l = i + 1; l < D; l++) acc = fma(U[l * D + i], X[c * D + l], acc);
                X[c * D + i] = (Bin[c * D + i] - acc) / U[i * D + i];
            }
    }
    template <int D, int UNR = 1>
    static __device__ __forceinline__ void solve_upper_t(const double* U, const double* Bin, int nrhs, double* X) {
        SQ_UNROLL
        for (int c = 0; c < nrhs; c++)
            SQ_UNROLL
            for (int i = 0; i < D; i++) {
                double acc = 0.0;
                SQ_UNROLL
                for (int l = 0; l < i; l++) acc = fma(U[i * D + l], X[c * D + l], acc);
                X[c * D + i] = (Bin[c * D + i] - acc) / U[i * D + i];
            }
    }
    // C = A'B (A: ra×ca, B: ra×cb) and C = A B (A: ra×ca, B: ca×cb), column-major
    template <int UNR = 1>
    static __device__ __forceinline__ void mul_AtB(const double* A, int ra, int ca, const double* Bm, int cb, double* Cm) {
        SQ_UNROLL
        for (int j = 0; j < cb; j++)
            SQ_UNROLL
            for (int i = 0; i < ca; i++) {
                double acc = 0.0;
                SQ_UNROLL
                for (int l = 0; l < ra; l++) acc = fma(A[i * ra + l], Bm[j * ra + l], acc);
                Cm[j * ca + i] = acc;
            }
    }
    template <int UNR = 1>
    static __device__ __forceinline__ void mul_AB(const double* A, int ra, int ca, const double* Bm, int cb, double* Cm) {
        SQ_UNROLL
        for (int j = 0; j < cb; j++)
            SQ_UNROLL
            for (int i = 0; i < ra; i++) {
                double acc = 0.0;
                SQ_UNROLL
                for (int l = 0; l < ca; l++) acc = fma(A[l * ra + i], Bm[j * ca + l], acc);
                Cm[j * ra + i] = acc;
            }
    }

    // cost expansion of knot k in square-root form -> E = [x(n) u(m) xx(n*n) uu(m*m) ux(m*n)] (the QST layout).
    // false: a stage Hessian is not positive definite (objective.jl:76-93)
    static __device__ bool expansion(const DevProblem& P, bool al_on, int k, const double* x, const double* u,
                                     const double* lam, const double* mu, double* E, double* Pbuf) {
        const int N = P.N;
        const bool term = (k == N - 1);
        double* Ex = E;
        double* Eu = E + n;
        double* Exx = E + n + m;
        double* Euu = Exx + n * n;
        double* Eux = Euu + m * m;
        for (int e = 0; e < C::QS; e++) E[e] = 0.0;
        if (!term) {
            // src/cost.jl:183-192 / minimum_time.jl:161-191
            double Qx[nq], Qu[mq];
            for (int i = 0; i < nq; i++) {
                double a = 0.0, b = 0.0;
                for (int j = 0; j < nq; j++) a = fma(P.Q[j * nq + i], x[j], a);
                for (int j = 0; j < mq; j++) b = fma(P.H[i * mq + j], u[j], b);
                Qx[i] = (a + P.q[i]) + b;
            }
            for (int i = 0; i < mq; i++) {
                double a = 0.0, b = 0.0;
                for (int j = 0; j < mq; j++) a = fma(P.R[j * mq + i], u[j], a);
                for (int j = 0; j < nq; j++) b = fma(P.H[j * mq + i], x[j], b);
                Qu[i] = (a + P.r[i]) + b;
            }
            double dt = P.dt, tau = 0.0;
            if (C::MT) { tau = u[m - 1]; dt = tau * tau; }
            for (int i = 0; i < nq; i++) Ex[i] = Qx[i] * dt;
            for (int i = 0; i < mq; i++) Eu[i] = Qu[i] * dt;
            for (int j = 0; j < nq; j++) for (int i = 0; i < nq; i++) Exx[j * n + i] = P.Q[j * nq + i] * dt;
            for (int j = 0; j < mq; j++) for (int i = 0; i < mq; i++) Euu[j * m + i] = P.R[j * mq + i] * dt;
            for (int j = 0; j < nq; j++) for (int i = 0; i < mq; i++) Eux[j * m + i] = P.H[j * mq + i] * dt;
            if (C::MT) {
                const double l1 = quad_stage<C>(P, x, u);
                Eu[m - 1] = tau * (2.0 * l1 + P.R_mt);
                for (int i = 0; i < mq; i++) {
                    const double t = (2.0 * tau) * Qu[i];
                    Euu[(m - 1) * m + i] = t;
                    Euu[i * m + (m - 1)] = t;
                }
                Euu[(m - 1) * m + (m - 1)] = 2.0 * l1 + P.R_mt;
                for (int i = 0; i < nq; i++) Eux[i * m + (m - 1)] = (2.0 * tau) * Qx[i];
                Ex[n - 1] = P.R_mt * x[n - 1];
                Exx[(n - 1) * n + (n - 1)] = P.R_mt;
            }
            if (!chol_upper_inplace<n>(Exx)) return false;
            if (!chol_upper_inplace<m>(Euu)) return false;
        } else {
            // src/cost.jl:194-198 / minimum_time.jl:193-204
            for (int j = 0; j < nq; j++) for (int i = 0; i < nq; i++) Exx[j * n + i] = P.Qf[j * nq + i];
            for (int i = 0; i < nq; i++) {
                double a = 0.0;
                for (int j = 0; j < nq; j++) a = fma(P.Qf[j * nq + i], x[j], a);
                Ex[i] = a + P.qf[i];
            }
            if (C::MT) {
                Exx[(n - 1) * n + (n - 1)] = P.R_mt;
                Ex[n - 1] = P.R_mt * x[n - 1];
            }
            if (!chol_upper_inplace<n>(Exx)) return false;
        }
        if (!al_on) return true;
        const int rb = P.knot_row_begin[k], pk = P.knot_row_count[k];
        if (pk == 0) return true;
        // augmented_lagrangian_methods.jl:231-276 (no ux term)
        double G[PMAX * nz], Bx[PMAX * n], Bu[PMAX * m], g[PMAX];
        const int cols = term ? n : nz;
        for (int r_ = 0; r_ < pk; r_++) {
            const DevRow r = P.rows[rb + r_];
            const double c = row_value<C>(r, x, u);
            const bool act = r.eq ? true : ((c >= 0.0) || (lam[r_] > 0.0));
            const double im = act ? mu[r_] : 0.0;
            for (int jj = 0; jj < nz; jj++) G[r_ * nz + jj] = (jj < cols) ? row_jac<C>(r, x, jj) : 0.0;
            g[r_] = im * c + lam[r_];
            const double sq = act ? sqrt(mu[r_]) : 0.0;
            for (int jj = 0; jj < n; jj++) Bx[jj * pk + r_] = sq * G[r_ * nz + jj];
            if (!term) for (int jj = 0; jj < m; jj++) Bu[jj * pk + r_] = sq * G[r_ * nz + n + jj];
        }
        {
            double R[n * n];
            chol_plus<n>(Exx, Bx, pk, R, Pbuf);
            for (int e = 0; e < n * n; e++) Exx[e] = R[e];
        }
        if (!term) {
            double R[m * m];
            chol_plus<m>(Euu, Bu, pk, R, Pbuf);
            for (int e = 0; e < m * m; e++) Euu[e] = R[e];
        }
        for (int i = 0; i < n; i++) {
            double acc = 0.0;
            for (int r_ = 0; r_ < pk; r_++) acc = fma(G[r_ * nz + i], g[r_], acc);
            Ex[i] += acc;
        }
        if (!term)
            for (int i = 0; i < m; i++) {
                double acc = 0.0;
                for (int r_ = 0; r_ < pk; r_++) acc = fma(G[r_ * nz + n + i], g[r_], acc);
                Eu[i] += acc;
            }
        return true;
    }

    // The square-root backward pass of one problem.  Returns 0 = ok, 1 = restart loop diverged (TO_STATUS_REG_DIVERGED),
    // 2 = PosDefException (TO_STATUS_NOT_PD_SQRT).
    // `pre_expanded`: ls_expand_sqrt_kernel has already written the expansion of every knot (slot N-1 = terminal knot) to the
    // Q trajectory, one thread per (problem, knot); the serial loop below is then skipped.
    static __device__ int run(const DevProblem& P, bool al_on, const TOiLQROptions& io, double* ws, const WsLayout& L,
                              double& rho, double& drho, double& dV0, double& dV1, bool pre_expanded = false) {
        const int N = P.N;
        double Pbuf[RMAX * ((n > m) ? n : m)];
        double Sxx[n * n], Sx[n];
        // cost_expansion_sqrt! of every knot first (ilqr_methods.jl:55-62): the Q trajectory lives in global memory and
        // is updated IN PLACE by the recursion, which reproduces the reference's restart behaviour as is
        double* qst = ws + L.QST;
        if (pre_expanded) {
            const double* E = qst + (size_t)(N - 1) * C::QS;
            for (int e = 0; e < n * n; e++) Sxx[e] = E[n + m + e];
            for (int i = 0; i < n; i++) Sx[i] = E[i];
        } else {
            double x[n], u[m], E[C::QS];
            for (int k = 0; k < N; k++) {
                for (int i = 0; i < n; i++) x[i] = ws[L.X + (size_t)k * n + i];
                for (int i = 0; i < m; i++) u[i] = (k < N - 1) ? ws[L.U + (size_t)k * m + i] : 0.0;
                const int lo = P.knot_lam_off[k];
                if (P.knot_row_count[k] > PMAX && al_on) return 2;
                if (!expansion(P, al_on, k, x, u, ws + L.LAM + lo, ws + L.MU + lo, E, Pbuf)) return 2;
                if (k < N - 1) {
                    for (int e = 0; e < C::QS; e++) qst[(size_t)k * C::QS + e] = E[e];
                } else {
                    for (int e = 0; e < n * n; e++) Sxx[e] = E[n + m + e];
                    for (int i = 0; i < n; i++) Sx[i] = E[i];
                }
            }
        }
        // the terminal factor is re-used by every restart
        double SxxN[n * n], SxN[n];
        for (int e = 0; e < n * n; e++) SxxN[e] = Sxx[e];
        for (int i = 0; i < n; i++) SxN[i] = Sx[i];
        dV0 = 0.0;
        dV1 = 0.0;
        int k = N - 2;
        while (k >= 0) {
            double A[n * n], Bm[n * m];
            {
                const double* ab = ws + L.Z + (size_t)k * C::ZA;
                for (int jj = 0; jj < n; jj++) for (int i = 0; i < n; i++) A[jj * n + i] = ab[i * C::LDZ + jj];
                for (int jj = 0; jj < m; jj++) for (int i = 0; i < n; i++) Bm[jj * n + i] = ab[i * C::LDZ + n + jj];
            }
            // the knot's blocks are worked on in a local copy (registers for the small models) and written back once they have
            // been updated -- before the regularisation test, so a restart finds the accumulated values (quirk Q1)
            double* Qg = qst + (size_t)k * C::QS;
            double Q[C::QS];
            for (int e = 0; e < C::QS; e++) Q[e] = Qg[e];
            double* Qx = Q;
            double* Qu = Q + n;
            double* Qxx = Q + n + m;
            double* Quu = Qxx + n * n;
            double* Qux = Quu + m * m;
            double v[n], vu[m], tx[n * n], tu[n * m], Mux[m * n];
            mul_AtB<UF>(A, n, n, Sx, 1, v);
            for (int i = 0; i < n; i++) Qx[i] += v[i];
            mul_AtB<UF>(Bm, n, m, Sx, 1, vu);
            for (int i = 0; i < m; i++) Qu[i] += vu[i];
            mul_AB<UF>(Sxx, n, n, A, n, tx);
            mul_AB<UF>(Sxx, n, n, Bm, m, tu);
            {
                double R[n * n];
                chol_plus<n, UF>(Qxx, tx, n, R, Pbuf);
                for (int e = 0; e < n * n; e++) Qxx[e] = R[e];
            }
            {
                double R[m * m];
                chol_plus<m, UF>(Quu, tu, n, R, Pbuf);
                for (int e = 0; e < m * m; e++) Quu[e] = R[e];
            }
            mul_AtB<UF>(tu, n, m, tx, n, Mux);
            for (int e = 0; e < m * n; e++) Qux[e] += Mux[e];
            for (int e = 0; e < C::QS; e++) Qg[e] = Q[e];
            double eye[m * m], Quu_reg[m * m];
            for (int e = 0; e < m * m; e++) eye[e] = 0.0;
            const double sr = sqrt(rho);
            for (int i = 0; i < m; i++) eye[i * m + i] = sr * 1.0;
            chol_plus<m, UF>(Quu, eye, m, Quu_reg, Pbuf);
            if (cond2<m, UF>(Quu_reg) > 1e8) {
                if (!isfinite(rho)) return 1;
                const double f = io.bp_reg_increase_factor;  // regularization_update!(:increase)
                drho = dmax(drho * f, f);
                rho = dmax(rho * drho, io.bp_reg_min);
                k = N - 2;
                for (int e = 0; e < n * n; e++) Sxx[e] = SxxN[e];
                for (int i = 0; i < n; i++) Sx[i] = SxN[i];
                dV0 = 0.0;
                dV1 = 0.0;
                continue;
            }
            // K = -Quu_reg \ (Quu_reg' \ Qux) ; d likewise
            double t1[m * n], Kk[m * n], dk[m], t1d[m];
            solve_upper_t<m, UF>(Quu_reg, Qux, n, t1);
            solve_upper<m, UF>(Quu_reg, t1, n, Kk);
            for (int e = 0; e < m * n; e++) Kk[e] = -Kk[e];
            solve_upper_t<m, UF>(Quu_reg, Qu, 1, t1d);
            solve_upper<m, UF>(Quu_reg, t1d, 1, dk);
            for (int i = 0; i < m; i++) dk[i] = -dk[i];
            {
                double* kd = ws + L.KD + (size_t)k * C::KDS;
                for (int e = 0; e < m * n; e++) kd[e] = Kk[e];
                for (int i = 0; i < m; i++) kd[m * n + i] = dk[i];
            }
            // S.x = Q.x + (K'Quu')(Quu d) + K'Qu + Qux'd
            double KQt[n * m], Qd[m], v1[n], v2[n], v3[n];
            for (int jj = 0; jj < m; jj++)
                for (int i = 0; i < n; i++) {
                    double acc = 0.0;
                    for (int l = 0; l < m; l++) acc = fma(Kk[i * m + l], Quu[l * m + jj], acc);
                    KQt[jj * n + i] = acc;
                }
            mul_AB<UF>(Quu, m, m, dk, 1, Qd);
            mul_AB<UF>(KQt, n, m, Qd, 1, v1);
            mul_AtB<UF>(Kk, m, n, Qu, 1, v2);
            mul_AtB<UF>(Qux, m, n, dk, 1, v3);
            double Sxk[n];
            for (int i = 0; i < n; i++) Sxk[i] = ((Qx[i] + v1[i]) + v2[i]) + v3[i];
            // tmp1 = (Q.xx') \ Q.ux'  (n×m) ; tmp2 = chol_minus(Q.uu, tmp1)
            double uxT[n * m], tmp1[n * m];
            for (int jj = 0; jj < m; jj++) for (int i = 0; i < n; i++) uxT[jj * n + i] = Qux[i * m + jj];
            solve_upper_t<n, UF>(Qxx, uxT, m, tmp1);
            double U2[m * m];
            for (int e = 0; e < m * m; e++) U2[e] = Quu[e];
            for (int i = 0; i < n; i++) {
                double rowv[m];
                for (int jj = 0; jj < m; jj++) rowv[jj] = tmp1[jj * n + i];
                if (!lowrank_downdate<m, UF>(U2, rowv)) return 2;
            }
            // S.xx = chol_plus(Q.xx + tmp1*K, tmp2*K)
            double top[n * n], bot[m * n];
            mul_AB<UF>(tmp1, n, m, Kk, n, top);
            for (int e = 0; e < n * n; e++) top[e] = Qxx[e] + top[e];
            mul_AB<UF>(U2, m, m, Kk, n, bot);
            chol_plus<n, UF>(top, bot, m, Sxx, Pbuf);
            for (int i = 0; i < n; i++) Sx[i] = Sxk[i];
            {
                double a = 0.0;
                for (int l = 0; l < m; l++) a = fma(dk[l], Qu[l], a);
                dV0 += a;
                double b = 0.0;
                for (int l = 0; l < m; l++) b = fma(Qd[l], Qd[l], b);
                dV1 += 0.5 * b;
            }
            k--;
        }
        {
            const double f = io.bp_reg_increase_factor;  // regularization_update!(:decrease)
            drho = dmin(drho / f, 1.0 / f);
            rho = rho * drho * ((rho * drho > io.bp_reg_min) ? 1.0 : 0.0);
        }
        return 0;
    }
};

// cost_expansion_sqrt! of every knot (objective.jl:70-94, augmented_lagrangian_methods.jl:231-276): thread per (problem, knot).
// The expansion does not depend on the cost-to-go; done here it is off the serial chain of ls_bp_sqrt_kernel (a lone problem
// paid N expansions in sequence before its recursion started).  A knot whose stage Hessian is not positive definite (or that has
// more rows than this path supports) flags the problem: st->bp_fail = 2, which the recursion kernel turns into NOT_PD_SQRT.
template <class C>
__global__ void __launch_bounds__(64) ls_expand_sqrt_kernel(const DevProblem P, const DevCtl ctl, const LsCtl lc, const int cur) {
    constexpr int n = C::n, m = C::m;
    const int N = P.N;
    const unsigned long long total = (unsigned long long)lc.counts[cur] * (unsigned long long)N;
    const WsLayout L = ws_layout<C>(N, P.Ptot, false);
    const bool al_on = (ctl.mode == 1);
    for (unsigned long long it = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; it < total;
         it += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned int a = (unsigned int)(it / N);
        const int k = (int)(it % N);
        const int b = lc.list[cur][a];
        double* ws = lc.ws + (size_t)b * lc.ws_stride;
        double x[n], u[m], E[C::QS];
        double Pbuf[SqrtBp<C>::RMAX * ((n > m) ? n : m)];
        for (int i = 0; i < n; i++) x[i] = ws[L.X + (size_t)k * n + i];
        for (int i = 0; i < m; i++) u[i] = (k < N - 1) ? ws[L.U + (size_t)k * m + i] : 0.0;
        const int lo = P.knot_lam_off[k];
        bool ok = !(P.knot_row_count[k] > SqrtBp<C>::PMAX && al_on);
        if (ok) ok = SqrtBp<C>::expansion(P, al_on, k, x, u, ws + L.LAM + lo, ws + L.MU + lo, E, Pbuf);
        if (ok) {
            double* q = ws + L.QST + (size_t)k * C::QS;
            for (int e = 0; e < C::QS; e++) q[e] = E[e];
        } else {
            lc.st[b].bp_fail = 2;
        }
    }
}

// thread per problem
template <class C>
__global__ void __launch_bounds__(64) ls_bp_sqrt_kernel(const DevProblem P, const DevCtl ctl, const LsCtl lc, const int cur) {
    const unsigned int na = lc.counts[cur];
    const WsLayout L = ws_layout<C>(P.N, P.Ptot, false);
    const bool al_on = (ctl.mode == 1);
    for (unsigned int a = blockIdx.x * blockDim.x + threadIdx.x; a < na; a += gridDim.x * blockDim.x) {
        const int b = lc.list[cur][a];
        LsState* st = &lc.st[b];
        TOiLQROptions io = ctl.o.opts_uncon;
        double rho = st->rho, drho = st->drho, dV0 = 0.0, dV1 = 0.0;
        // the expansions were written by ls_expand_sqrt_kernel, which also flags a failed knot
        const int rc = (st->bp_fail == 2) ? 2 : SqrtBp<C>::run(P, al_on, io, lc.ws + (size_t)b * lc.ws_stride, L, rho, drho, dV0, dV1, true);
        st->rho = rho; st->drho = drho; st->dV0 = dV0; st->dV1 = dV1;
        st->winner = -1;
        st->bp_fail = rc;
    }
}


// ---------------------------------------------------------------------------------------------------------------------------------
// Latency form of the same recursion: one WARP per problem, every matrix in shared memory, one lane per OUTPUT element.
// Each output is still produced by the sequential chain of SqrtBp::run (same operands, same order), so the two kernels agree bit
// for bit; what changes is that independent outputs -- the columns a Householder reflection updates, the right-hand sides of a
// triangular solve, the entries of a product, the two factor updates of a knot -- run side by side instead of one after the other,
// and that no matrix lives in a thread-local frame.  The rank-one down-dates (n rows against the m×m factor, backward_pass.jl:165-170)
// run as a pipeline: lane i owns pivot row i and hands the modified row vector to lane i+1.
template <class C>
struct SqrtWarpSmem {
    static constexpr int n = C::n, m = C::m;
    static constexpr int RB = (n > m) ? n : m;
    double Sxx[n * n], Sx[n], SxxN[n * n], SxN[n];
    double A[n * n], Bm[n * m];
    double Q[C::QS];
    double tx[n * n], tu[n * m];
    double P1[(n + RB) * n];      // stacked matrix of the n-column factor updates (rows: 2n, then n+m)
    double P2[(m + RB) * m];      // stacked matrix of the m-column factor updates (rows: m+n, then 2m)
    double Qreg[m * m];
    double K[m * n], d[m];
    double KQt[n * m], Qd[m], v2[n], v3[n], Sxk[n];
    double tmp1[n * m], U2[m * m];
};

template <class C>
struct SqrtWarp {
    static constexpr int n = C::n, m = C::m;

    // Householder QR (same reflections as SqrtBp::qr_R) of two stacked matrices at once: lanes 0-15 work on P1 (rows1 × D1),
    // lanes 16-31 on P2 (rows2 × D2).  Every lane of a half computes the column norm itself (a sequential chain that cannot be
    // split); the reflection is then applied to one column per lane.  The scaled Householder vector is never stored: only R is
    // read afterwards.
    static __device__ __forceinline__ void qr_pair(double* P1, int rows1, int D1, double* P2, int rows2, int D2, int lane) {
        const int half = lane >> 4, gl = lane & 15;
        double* Pm = half ? P2 : P1;
        const int rows = half ? rows2 : rows1, D = half ? D2 : D1;
        const int DM = (D1 > D2) ? D1 : D2;
        for (int j = 0; j < DM; j++) {
            const bool on = (j < D && j < rows);
            double tau = 0.0, sc = 0.0, beta = 0.0;
            bool refl = false;
            if (on) {
                const double* cj = Pm + j * rows;
                const double alpha = cj[j];
                double xn2 = 0.0;
                for (int i = j + 1; i < rows; i++) xn2 = fma(cj[i], cj[i], xn2);
                if (xn2 != 0.0) {
                    const double xnorm = sqrt(xn2);
                    beta = -copysign(sqrt(alpha * alpha + xnorm * xnorm), alpha);
                    tau = (beta - alpha) / beta;
                    sc = 1.0 / (alpha - beta);
                    refl = true;
                }
            }
            // (column j is only read from here on, columns c > j are each written by their own lane: no barrier needed yet)
            if (on && tau != 0.0) {
                const double* cj = Pm + j * rows;
                for (int c = j + 1 + gl; c < D; c += 16) {
                    double* cc = Pm + c * rows;
                    double w = cc[j];
                    for (int i = j + 1; i < rows; i++) w = fma(cj[i] * sc, cc[i], w);
                    const double tw = tau * w;
                    cc[j] = cc[j] - tw;
                    for (int i = j + 1; i < rows; i++) cc[i] = fma(-tw, cj[i] * sc, cc[i]);
                }
            }
            __syncwarp();
            // (the diagonal entry is written once every lane is past the reads of column j)
            if (refl && gl == 0) Pm[j * rows + j] = beta;
        }
        __syncwarp();
    }

    static __device__ int run(const DevProblem& P, const TOiLQROptions& io, double* ws, const WsLayout& L, SqrtWarpSmem<C>& sm,
                              const int lane, double& rho, double& drho, double& dV0, double& dV1) {
        const int N = P.N;
        double* qst = ws + L.QST;
        {
            const double* E = qst + (size_t)(N - 1) * C::QS;
            for (int e = lane; e < n * n; e += 32) { const double v = E[n + m + e]; sm.Sxx[e] = v; sm.SxxN[e] = v; }
            for (int i = lane; i < n; i += 32) { const double v = E[i]; sm.Sx[i] = v; sm.SxN[i] = v; }
        }
        dV0 = 0.0;
        dV1 = 0.0;
        double* Qx = sm.Q;
        double* Qu = sm.Q + n;
        double* Qxx = sm.Q + n + m;
        double* Quu = Qxx + n * n;
        double* Qux = Quu + m * m;
        int k = N - 2;
        while (k >= 0) {
            __syncwarp();
            {
                const double* ab = ws + L.Z + (size_t)k * C::ZA;
                for (int e = lane; e < n * n; e += 32) { const int jj = e / n, i = e - jj * n; sm.A[e] = ab[i * C::LDZ + jj]; }
                for (int e = lane; e < n * m; e += 32) { const int jj = e / n, i = e - jj * n; sm.Bm[e] = ab[i * C::LDZ + n + jj]; }
                const double* Qg = qst + (size_t)k * C::QS;
                for (int e = lane; e < C::QS; e += 32) sm.Q[e] = Qg[e];
            }
            __syncwarp();
            // Sxx*A, Sxx*B, A'Sx, B'Sx: one output per lane
            for (int e = lane; e < n * n + n * m + n + m; e += 32) {
                double acc = 0.0;
                if (e < n * n) {
                    const int j = e / n, i = e - j * n;
                    for (int l = 0; l < n; l++) acc = fma(sm.Sxx[l * n + i], sm.A[j * n + l], acc);
                    sm.tx[e] = acc;
                } else if (e < n * n + n * m) {
                    const int f = e - n * n, j = f / n, i = f - j * n;
                    for (int l = 0; l < n; l++) acc = fma(sm.Sxx[l * n + i], sm.Bm[j * n + l], acc);
                    sm.tu[f] = acc;
                } else if (e < n * n + n * m + n) {
                    const int i = e - n * n - n * m;
                    for (int l = 0; l < n; l++) acc = fma(sm.A[i * n + l], sm.Sx[l], acc);
                    Qx[i] += acc;
                } else {
                    const int i = e - n * n - n * m - n;
                    for (int l = 0; l < n; l++) acc = fma(sm.Bm[i * n + l], sm.Sx[l], acc);
                    Qu[i] += acc;
                }
            }
            __syncwarp();
            // [Qxx; Sxx A] and [Quu; Sxx B] stacked; Qux += (Sxx B)'(Sxx A)
            for (int e = lane; e < 2 * n * n; e += 32) {
                const int j = e / (2 * n), i = e - j * 2 * n;
                sm.P1[e] = (i < n) ? Qxx[j * n + i] : sm.tx[j * n + (i - n)];
            }
            for (int e = lane; e < (m + n) * m; e += 32) {
                const int j = e / (m + n), i = e - j * (m + n);
                sm.P2[e] = (i < m) ? Quu[j * m + i] : sm.tu[j * n + (i - m)];
            }
            for (int e = lane; e < m * n; e += 32) {
                const int j = e / m, i = e - j * m;
                double acc = 0.0;
                for (int l = 0; l < n; l++) acc = fma(sm.tu[i * n + l], sm.tx[j * n + l], acc);
                Qux[e] += acc;
            }
            __syncwarp();
            qr_pair(sm.P1, 2 * n, n, sm.P2, m + n, m, lane);
            {
                double* Qg = qst + (size_t)k * C::QS;
                for (int e = lane; e < n * n; e += 32) { const int j = e / n, i = e - j * n; Qxx[e] = (i <= j) ? sm.P1[j * 2 * n + i] : 0.0; }
                for (int e = lane; e < m * m; e += 32) { const int j = e / m, i = e - j * m; Quu[e] = (i <= j) ? sm.P2[j * (m + n) + i] : 0.0; }
                __syncwarp();
                // written back before the regularisation test: a restart finds the accumulated values (quirk Q1)
                for (int e = lane; e < C::QS; e += 32) Qg[e] = sm.Q[e];
            }
            // Quu_reg = chol_plus(Quu, sqrt(rho) I)
            const double sr = sqrt(rho);
            for (int e = lane; e < 2 * m * m; e += 32) {
                const int j = e / (2 * m), i = e - j * 2 * m;
                sm.P2[e] = (i < m) ? Quu[j * m + i] : ((i - m == j) ? sr * 1.0 : 0.0);
            }
            __syncwarp();
            qr_pair(sm.P2, 2 * m, m, sm.P2, 0, 0, lane);
            for (int e = lane; e < m * m; e += 32) { const int j = e / m, i = e - j * m; sm.Qreg[e] = (i <= j) ? sm.P2[j * 2 * m + i] : 0.0; }
            __syncwarp();
            // (every lane runs the Jacobi sweeps on its own register copy: a chain with nothing to share out)
            if (SqrtBp<C>::template cond2<m, 64>(sm.Qreg) > 1e8) {
                if (!isfinite(rho)) return 1;
                const double f = io.bp_reg_increase_factor;  // regularization_update!(:increase)
                drho = dmax(drho * f, f);
                rho = dmax(rho * drho, io.bp_reg_min);
                k = N - 2;
                __syncwarp();
                for (int e = lane; e < n * n; e += 32) sm.Sxx[e] = sm.SxxN[e];
                for (int i = lane; i < n; i += 32) sm.Sx[i] = sm.SxN[i];
                dV0 = 0.0;
                dV1 = 0.0;
                continue;
            }
            // K = -Quu_reg \ (Quu_reg' \ Qux), d likewise: one right-hand side per lane
            for (int c = lane; c < n + 1; c += 32) {
                double t1[m], x2[m];
#pragma unroll
                for (int i = 0; i < m; i++) {
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < i; l++) acc = fma(sm.Qreg[i * m + l], t1[l], acc);
                    const double rhs = (c < n) ? Qux[c * m + i] : Qu[i];
                    t1[i] = (rhs - acc) / sm.Qreg[i * m + i];
                }
#pragma unroll
                for (int i = m - 1; i >= 0; i--) {
                    double acc = 0.0;
#pragma unroll
                    for (int l = i + 1; l < m; l++) acc = fma(sm.Qreg[l * m + i], x2[l], acc);
                    x2[i] = (t1[i] - acc) / sm.Qreg[i * m + i];
                }
                double* kd = ws + L.KD + (size_t)k * C::KDS;
#pragma unroll
                for (int i = 0; i < m; i++) {
                    const double v = -x2[i];
                    if (c < n) { sm.K[c * m + i] = v; kd[c * m + i] = v; }
                    else { sm.d[i] = v; kd[m * n + i] = v; }
                }
            }
            __syncwarp();
            // K'Quu', Quu d, K'Qu, Qux'd, and (Qxx') \ Qux' (one column per lane, taken by the last lanes)
            for (int e = lane; e < n * m + m + n + n; e += 32) {
                double acc = 0.0;
                if (e < n * m) {
                    const int jj = e / n, i = e - jj * n;
                    for (int l = 0; l < m; l++) acc = fma(sm.K[i * m + l], Quu[l * m + jj], acc);
                    sm.KQt[e] = acc;
                } else if (e < n * m + m) {
                    const int i = e - n * m;
                    for (int l = 0; l < m; l++) acc = fma(Quu[l * m + i], sm.d[l], acc);
                    sm.Qd[i] = acc;
                } else if (e < n * m + m + n) {
                    const int i = e - n * m - m;
                    for (int l = 0; l < m; l++) acc = fma(sm.K[i * m + l], Qu[l], acc);
                    sm.v2[i] = acc;
                } else {
                    const int i = e - n * m - m - n;
                    for (int l = 0; l < m; l++) acc = fma(Qux[i * m + l], sm.d[l], acc);
                    sm.v3[i] = acc;
                }
            }
            for (int c = 31 - lane; c < m; c += 32) {
                for (int i = 0; i < n; i++) {
                    double acc = 0.0;
                    for (int l = 0; l < i; l++) acc = fma(Qxx[i * n + l], sm.tmp1[c * n + l], acc);
                    sm.tmp1[c * n + i] = (Qux[i * m + c] - acc) / Qxx[i * n + i];
                }
            }
            for (int e = lane; e < m * m; e += 32) sm.U2[e] = Quu[e];
            __syncwarp();
            // S.x = Q.x + (K'Quu')(Quu d) + K'Qu + Qux'd ; expected change
            for (int i = lane; i < n; i += 32) {
                double acc = 0.0;
                for (int l = 0; l < m; l++) acc = fma(sm.KQt[l * n + i], sm.Qd[l], acc);
                sm.Sxk[i] = ((Qx[i] + acc) + sm.v2[i]) + sm.v3[i];
            }
            // tmp2 = chol_minus(Q.uu, tmp1): n rank-one down-dates, pipelined over the pivot rows (lane i = pivot i)
            {
                double Urow[m], vv[m];
#pragma unroll
                for (int j = 0; j < m; j++) { Urow[j] = (lane < m && j >= lane) ? sm.U2[j * m + lane] : 0.0; vv[j] = 0.0; }
                bool bad = false;
                for (int t = 0; t < n + m - 1; t++) {
#pragma unroll
                    for (int j = 0; j < m; j++) {
                        const double up = __shfl_up_sync(FULL, vv[j], 1);
                        if (lane > 0) vv[j] = up;
                    }
                    if (lane == 0 && t < n) {
#pragma unroll
                        for (int j = 0; j < m; j++) vv[j] = sm.tmp1[j * n + t];
                    }
                    const int r = t - lane;
                    if (lane < m && r >= 0 && r < n) {
                        double Aii = 0.0, vi = 0.0;
#pragma unroll
                        for (int j = 0; j < m; j++) if (j == lane) { Aii = Urow[j]; vi = vv[j]; }
                        const double s = vi / Aii;
                        const double s2 = s * s;
                        if (s2 > 1.0) bad = true;
                        const double c = sqrt(1.0 - s2);
#pragma unroll
                        for (int j = 0; j < m; j++) {
                            if (j == lane) {
                                Urow[j] = c * Aii;
                            } else if (j > lane) {
                                const double vj = vv[j];
                                const double Aij = (Urow[j] - s * vj) / c;
                                Urow[j] = Aij;
                                vv[j] = -s * Aij + c * vj;
                            }
                        }
                    }
                }
                if (__any_sync(FULL, bad)) return 2;
                double a = 0.0;
                for (int l = 0; l < m; l++) a = fma(sm.d[l], Qu[l], a);
                dV0 += a;
                double b = 0.0;
                for (int l = 0; l < m; l++) b = fma(sm.Qd[l], sm.Qd[l], b);
                dV1 += 0.5 * b;
                if (lane < m) {
#pragma unroll
                    for (int j = 0; j < m; j++) if (j >= lane) sm.U2[j * m + lane] = Urow[j];
                }
            }
            __syncwarp();
            // S.xx = chol_plus(Q.xx + tmp1*K, tmp2*K)
            for (int e = lane; e < n * n + m * n; e += 32) {
                double acc = 0.0;
                if (e < n * n) {
                    const int j = e / n, i = e - j * n;
                    for (int l = 0; l < m; l++) acc = fma(sm.tmp1[l * n + i], sm.K[j * m + l], acc);
                    sm.P1[j * (n + m) + i] = Qxx[e] + acc;
                } else {
                    const int f = e - n * n, j = f / m, i = f - j * m;
                    for (int l = 0; l < m; l++) acc = fma(sm.U2[l * m + i], sm.K[j * m + l], acc);
                    sm.P1[j * (n + m) + n + i] = acc;
                }
            }
            __syncwarp();
            qr_pair(sm.P1, n + m, n, sm.P1, 0, 0, lane);
            for (int e = lane; e < n * n; e += 32) { const int j = e / n, i = e - j * n; sm.Sxx[e] = (i <= j) ? sm.P1[j * (n + m) + i] : 0.0; }
            for (int i = lane; i < n; i += 32) sm.Sx[i] = sm.Sxk[i];
            k--;
        }
        {
            const double f = io.bp_reg_increase_factor;  // regularization_update!(:decrease)
            drho = dmin(drho / f, 1.0 / f);
            rho = rho * drho * ((rho * drho > io.bp_reg_min) ? 1.0 : 0.0);
        }
        return 0;
    }
};

// warp per problem (few live problems: the pass is a latency chain)
template <class C>
__global__ void __launch_bounds__(32) ls_bp_sqrt_warp_kernel(const DevProblem P, const DevCtl ctl, const LsCtl lc, const int cur) {
    __shared__ SqrtWarpSmem<C> sm;
    const unsigned int na = lc.counts[cur];
    const WsLayout L = ws_layout<C>(P.N, P.Ptot, false);
    const int lane = threadIdx.x;
    for (unsigned int a = blockIdx.x; a < na; a += gridDim.x) {
        const int b = lc.list[cur][a];
        LsState* st = &lc.st[b];
        const TOiLQROptions& io = ctl.o.opts_uncon;
        double rho = st->rho, drho = st->drho, dV0 = 0.0, dV1 = 0.0;
        const int pre = st->bp_fail;
        __syncwarp();
        const int rc = (pre == 2) ? 2 : SqrtWarp<C>::run(P, io, lc.ws + (size_t)b * lc.ws_stride, L, sm, lane, rho, drho, dV0, dV1);
        __syncwarp();
        if (lane == 0) {
            st->rho = rho; st->drho = drho; st->dV0 = dV0; st->dV1 = dV1;
            st->winner = -1;
            st->bp_fail = rc;
        }
    }
}

#undef SQ_UNROLL

}  // namespace tob
