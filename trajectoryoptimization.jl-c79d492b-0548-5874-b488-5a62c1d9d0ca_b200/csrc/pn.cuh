// B200 engine: projected-Newton polish of ALTRO on the device (ls_pn_kernel), ONE CTA per problem.
//
// Reference: src/solvers/direct/projected_newton.jl, the `solve_type = :feasible` path ALTRO uses
//   solve! (:4-19) -> newton_step! (:500-520) -> update! (:154-165) -> projection_solve! (:206-218)
//   -> _projection_solve! (:221-270) -> _projection_linesearch! (:272-312) -> reg_solve (:314-333),
// hooked in at src/solvers/altro/altro_methods.jl:6-14,31-39.
//
// The reference stacks y = [x1 - x0; f(x_k,u_k) - x_{k+1}; active c_k; ...; active c_N] with its sparse Jacobian Y over
// Z = [x1,u1,...,xN] (direct_solvers.jl:62-110), forms S = Y H^-1 Y' (H = Diagonal of the cost Hessian) and lets CHOLMOD
// factorise S + 1e-2 I.  S is block tridiagonal: block s = 0 is x1 - x0, block s = 1..N-1 holds the dynamics rows into knot s and
// the active rows of knot s-1, block N the active terminal rows; neighbours couple through one state.  Here each problem's CTA
// builds the blocks (dense, at most n + LS_PN_PA rows), runs a block Cholesky in shared memory, and does the refinement /
// line-search loops of the reference with the block factor.  Every element is accumulated in the order oracle/oracle_pn.hpp
// uses (sequential FMA chains from index 0), so device and oracle agree to rounding of the final bits; neither is bit-identical
// to CHOLMOD's elimination order ("parity unpinned" for the PN iterates, properties pinned by tests).
#pragma once

namespace tob {

constexpr int LS_PN_PA = 24;  // active constraint rows of one knot the block factor holds (a knot with more -> TO_STATUS_PN_SKIPPED)
constexpr int LS_PN_THREADS = 128;
constexpr int LS_PN_MINB = 4;  // 128 registers, four CTAs per SM: the kernel is a latency chain per problem (profiles/r03f_pn_throughput.log: 2 / 4 / 5 / 6 CTAs per SM -> polish of 8,192 quadrotor problems 2,201 / 1,229 / 1,375 / 1,809 ms)

template <class C> struct PnDims {
    static constexpr int n = C::n, m = C::m;
    static constexpr int R = n + LS_PN_PA;   // rows of a block
    static constexpr int RR = R * R;
    static constexpr int W = n + m + n;      // columns a stage block touches: x_k, u_k, x_{k+1}
};

// per-slot scratch in global memory (doubles); the int tables sit at the end
template <class C>
struct PnLayout {
    unsigned long long Xb, Ub, Xn, Un, fv, Cv, Gc, y, dl, t, rr, sx, dx, rd, Sd, Ld, Lo, ints, total;
    // ints: act[Ptot], pa[N], aidx[N][LS_PN_PA], rblk[N+1]
};
template <class C>
__host__ __device__ inline PnLayout<C> pn_layout(int N, int Ptot) {
    typedef PnDims<C> D;
    PnLayout<C> L;
    unsigned long long o = 0;
    const unsigned long long nx = (unsigned long long)N * D::n, nu = (unsigned long long)(N - 1) * D::m;
    const unsigned long long nb = (unsigned long long)(N + 1) * D::R;
    L.Xb = o; o += nx;  L.Ub = o; o += nu;
    L.Xn = o; o += nx;  L.Un = o; o += nu;
    L.fv = o; o += nx;
    L.Cv = o; o += (unsigned long long)Ptot;
    L.Gc = o; o += (unsigned long long)N * LS_PN_PA * (D::n + D::m);
    L.y = o; o += nb;  L.dl = o; o += nb;  L.t = o; o += nb;  L.rr = o; o += nb;  L.sx = o; o += nb;  L.dx = o; o += nb;
    L.rd = o; o += nb;  // 1 / diag(L) of every block (exact reciprocals: the triangular solves divide through div_by)
    L.Sd = o; o += (unsigned long long)(N + 1) * D::RR;
    L.Ld = o; o += (unsigned long long)(N + 1) * D::RR;
    L.Lo = o; o += (unsigned long long)(N + 1) * D::RR;
    L.ints = o;
    const unsigned long long ni = (unsigned long long)Ptot + N + (unsigned long long)N * LS_PN_PA + (N + 1);
    o += (ni + 1) / 2;
    L.total = (o + 15ull) & ~15ull;
    return L;
}

template <class C>
struct PnSmem {
    typedef PnDims<C> D;
    double A[D::RR];        // the block being factorised (S_ss + rho I - Lo Lo'), then L_ss
    double Lp[D::RR];       // L_{s-1,s-1}
    double Lo[D::RR];       // L_{s,s-1}
    double G[D::R * D::W];  // rows of Y of the block
    double vec[2 * D::R];
    unsigned long long viol_bits;
    int flag;
};

template <class C>
struct PnCtx {
    typedef PnDims<C> D;
    static constexpr int n = D::n, m = D::m, R = D::R, RR = D::RR, W = D::W, NT = LS_PN_THREADS;
    const DevProblem& P;
    PnSmem<C>& sm;
    double* ws;       // the problem's workspace (X, U, [A B] of the AL solve)
    double* sc;       // this CTA's scratch
    int* ints;
    const WsLayout L;
    const PnLayout<C> PL;
    const double* x0;
    const int tid;
    const int N;

    __device__ PnCtx(const DevProblem& P_, PnSmem<C>& sm_, double* ws_, double* sc_, const double* x0_, int tid_)
        : P(P_), sm(sm_), ws(ws_), sc(sc_), L(ws_layout<C>(P_.N, P_.Ptot, false)), PL(pn_layout<C>(P_.N, P_.Ptot)), x0(x0_), tid(tid_),
          N(P_.N) {
        ints = reinterpret_cast<int*>(sc + PL.ints);
    }
    __device__ int* act() { return ints; }
    __device__ int* pa() { return ints + P.Ptot; }
    __device__ int* aidx() { return ints + P.Ptot + N; }
    __device__ int* rblk() { return ints + P.Ptot + N + N * LS_PN_PA; }

    // diagonal of the cost Hessian (projected_newton.jl:122-148 with cost.jl:214-228: [Q H'; H R] dt, terminal Qf)
    __device__ double hx(int k, int i) const { return (i < C::nq) ? ((k < N - 1) ? P.Q[i * C::nq + i] * P.dt : P.Qf[i * C::nq + i]) : 0.0; }
    __device__ double hu(int i) const { return (i < C::mq) ? P.R[i * C::mq + i] * P.dt : 0.0; }

    // dynamics_constraints! (:34-42) + update_constraints! (:67-73) at (X, U): fv, Cv
    __device__ void residuals(const double* X, const double* U) {
        for (int k = tid; k < N; k += NT) {
            double x[n], u[m];
#pragma unroll
            for (int i = 0; i < n; i++) x[i] = X[(size_t)k * n + i];
            if (k == 0) {
#pragma unroll
                for (int i = 0; i < n; i++) sc[PL.fv + i] = x[i] - ((i < C::n0) ? x0[i] : 0.0);
            }
            if (k < N - 1) {
#pragma unroll
                for (int i = 0; i < m; i++) u[i] = U[(size_t)k * m + i];
                double xn[n];
                fd_model<C::MODEL, C::INTEG, double>(xn, x, u, P.dt);
                if constexpr (C::INF) {
#pragma unroll
                    for (int i = 0; i < C::n0; i++) xn[i] = xn[i] + u[C::m0 + i];
                }
#pragma unroll
                for (int i = 0; i < n; i++) sc[PL.fv + (size_t)(k + 1) * n + i] = xn[i] - X[(size_t)(k + 1) * n + i];
            } else {
#pragma unroll
                for (int i = 0; i < m; i++) u[i] = 0.0;
            }
            const int rb = P.knot_row_begin[k], rc = P.knot_row_count[k], lo = P.knot_lam_off[k];
            for (int i = 0; i < rc; i++) sc[PL.Cv + lo + i] = row_value<C>(P.rows[rb + i], x, u);
        }
        __syncthreads();
    }
    // active_set! (:75-92): equality rows, and inequality rows with c >= -tol.  Also the per-knot list of active rows.
    // Returns false if a knot has more active rows than the block factor holds.
    __device__ bool active_set(double tol) {
        if (tid == 0) sm.flag = 1;
        __syncthreads();
        for (int k = tid; k < N; k += NT) {
            const int rb = P.knot_row_begin[k], rc = P.knot_row_count[k], lo = P.knot_lam_off[k];
            int cnt = 0;
            for (int i = 0; i < rc; i++) {
                const int a = P.rows[rb + i].eq ? 1 : (sc[PL.Cv + lo + i] >= -tol);
                act()[lo + i] = a;
                if (a) {
                    if (cnt < LS_PN_PA) aidx()[k * LS_PN_PA + cnt] = i;
                    cnt++;
                }
            }
            pa()[k] = cnt;
            if (cnt > LS_PN_PA) sm.flag = 0;
        }
        __syncthreads();
        return sm.flag != 0;
    }
    __device__ int rows_of(int s) { return (s == 0) ? n : ((s < N) ? n + pa()[s - 1] : pa()[N - 1]); }

    // y = solver.y[a.duals] per block (padded to R); returns norm(y, Inf)
    __device__ double gather_y(double* y) {
        if (tid == 0) sm.viol_bits = 0ull;
        __syncthreads();
        double v = 0.0;
        for (int e = tid; e < (N + 1) * R; e += NT) {
            const int s = e / R, i = e - s * R;
            const int r = rows_of(s);
            double val = 0.0;
            if (i < r) {
                if (s < N && i < n) val = sc[PL.fv + (size_t)s * n + i];
                else {
                    const int k = s - 1, q = (s < N) ? i - n : i;
                    val = sc[PL.Cv + P.knot_lam_off[k] + aidx()[k * LS_PN_PA + q]];
                }
                v = dmax(v, fabs(val));
            }
            y[e] = val;
        }
        atomicMax(&sm.viol_bits, (unsigned long long)__double_as_longlong(v));  // non-negative doubles order like their bit patterns
        __syncthreads();
        const double out = __longlong_as_double((long long)sm.viol_bits);
        __syncthreads();
        return out;
    }

    // rows of Y of block s into sm.G (r x wd, row-major; wd = n for s = 0 and s = N, else W)
    __device__ void build_G(int s, int r) {
        if (s == 0) {
            for (int e = tid; e < n * n; e += NT) sm.G[e] = ((e / n) == (e % n)) ? 1.0 : 0.0;
        } else if (s < N) {
            const int k = s - 1;
            const double* ab = ws + L.Z + (size_t)k * C::ZA;
            const double* gc = sc + PL.Gc + (size_t)k * LS_PN_PA * (n + m);
            for (int e = tid; e < r * W; e += NT) {
                const int i = e / W, c = e - i * W;
                double v;
                if (i < n) v = (c < n + m) ? ab[i * C::LDZ + c] : ((c - n - m == i) ? -1.0 : 0.0);
                else v = (c < n + m) ? gc[(i - n) * (n + m) + c] : 0.0;
                sm.G[e] = v;
            }
        } else {
            const double* gc = sc + PL.Gc + (size_t)(N - 1) * LS_PN_PA * (n + m);
            for (int e = tid; e < r * n; e += NT) sm.G[e] = gc[(e / n) * (n + m) + (e % n)];
        }
        __syncthreads();
    }
    __device__ double hcc(int s, int c) const {  // H_cc of column c of block s
        if (s == 0) return hx(0, c);
        if (s == N) return hx(N - 1, c);
        const int k = s - 1;
        return (c < n) ? hx(k, c) : ((c < n + m) ? hu(c - n) : hx(k + 1, c - n - m));
    }

    // dynamics_jacobian! (:45-64), constraint_jacobian! (:98-107), S = Y (H \ Y') (:246-247), block Cholesky of S + rho I.
    // Returns false if a pivot is not positive (the reference's cholesky throws).
    __device__ bool factor(double rho) {
        constexpr int JPC = (C::MODEL == 4) ? 1 : C::PC;
        constexpr int NCH = ls_jac_chunks<C, JPC>();
        for (int it = tid; it < (N - 1) * NCH; it += NT) {
            const int k = it / NCH;
            ls_jac_item<C, JPC>(P, ws, L, k, it - k * NCH);
        }
        // active constraint Jacobian rows, dense over [x;u]
        for (int k = tid; k < N; k += NT) {
            const int cnt = pa()[k];
            if (cnt == 0) continue;
            double z[n + m];
#pragma unroll
            for (int i = 0; i < n; i++) z[i] = ws[L.X + (size_t)k * n + i];
#pragma unroll
            for (int i = 0; i < m; i++) z[n + i] = (k < N - 1) ? ws[L.U + (size_t)k * m + i] : 0.0;
            const int rb = P.knot_row_begin[k];
            double* gc = sc + PL.Gc + (size_t)k * LS_PN_PA * (n + m);
            for (int q = 0; q < cnt; q++) {
                const DevRow r = P.rows[rb + aidx()[k * LS_PN_PA + q]];
                for (int c = 0; c < n + m; c++) gc[q * (n + m) + c] = (k < N - 1 || c < n) ? BpGroup<C>::row_jac_s(r, z, c) : 0.0;
            }
        }
        __syncthreads();
        if (tid == 0) sm.flag = 1;
        int rp = 0;
        for (int s = 0; s <= N; s++) {
            const int r = rows_of(s);
            if (tid == 0) rblk()[s] = r;
            const int wd = (s == 0 || s == N) ? n : W;
            build_G(s, r);
            double* Sd = sc + PL.Sd + (size_t)s * RR;
            for (int e = tid; e < r * r; e += NT) {
                const int i = e / r, j = e - i * r;
                double acc = 0.0;
                for (int c = 0; c < wd; c++) acc = fma(sm.G[i * wd + c], sm.G[j * wd + c] / hcc(s, c), acc);
                Sd[e] = acc;
                sm.A[e] = (i == j) ? acc + rho : acc;
            }
            __syncthreads();
            if (s >= 1) {
                // L_{s,s-1} = S_{s,s-1} L_{s-1,s-1}^-T, row by row; S_{s,s-1}(i,j) = G(i, x_j) * (-+1 / H) for the n dynamics rows j
                const double gprev = (s - 1 == 0) ? 1.0 : -1.0;
                for (int i = tid; i < r; i += NT) {
                    for (int j = 0; j < rp; j++) {
                        double v = (j < n) ? sm.G[i * wd + j] * (gprev / hx(s - 1, j)) : 0.0;
                        for (int l = 0; l < j; l++) v = fma(-sm.Lo[i * rp + l], sm.Lp[j * rp + l], v);
                        sm.Lo[i * rp + j] = v / sm.Lp[j * rp + j];
                    }
                }
                __syncthreads();
                double* Lo = sc + PL.Lo + (size_t)s * RR;
                for (int e = tid; e < r * rp; e += NT) Lo[e] = sm.Lo[e];
                for (int e = tid; e < r * r; e += NT) {
                    const int i = e / r, j = e - i * r;
                    if (j <= i) {
                        double acc = 0.0;
                        for (int l = 0; l < rp; l++) acc = fma(sm.Lo[i * rp + l], sm.Lo[j * rp + l], acc);
                        sm.A[i * r + j] -= acc;
                    }
                }
                __syncthreads();
            }
            // Cholesky of the block, column by column (oracle_pn.hpp pn_chol)
            for (int j = 0; j < r; j++) {
                if (tid == 0) {
                    double d = sm.A[j * r + j];
                    for (int l = 0; l < j; l++) d = fma(-sm.A[j * r + l], sm.A[j * r + l], d);
                    if (!(d > 0.0)) sm.flag = 0;
                    sm.A[j * r + j] = sqrt(d);
                }
                __syncthreads();
                const double d = sm.A[j * r + j];
                for (int i = j + 1 + tid; i < r; i += NT) {
                    double v = sm.A[i * r + j];
                    for (int l = 0; l < j; l++) v = fma(-sm.A[i * r + l], sm.A[j * r + l], v);
                    sm.A[i * r + j] = v / d;
                }
                __syncthreads();
            }
            if (!sm.flag) return false;
            double* Ld = sc + PL.Ld + (size_t)s * RR;
            for (int e = tid; e < r * r; e += NT) { Ld[e] = sm.A[e]; sm.Lp[e] = sm.A[e]; }
            for (int i = tid; i < r; i += NT) sc[PL.rd + (size_t)s * R + i] = 1.0 / sm.A[i * r + i];
            rp = r;
            __syncthreads();
        }
        return true;
    }

    // x = (S + rho I)^-1 b with the block factor (oracle_pn.hpp pn_solve_factor); b, x, t: padded block vectors in scratch
    __device__ void solve_factor(const double* b, double* x) {
        double* t = sc + PL.t;
        const int* rb_ = rblk();
        // forward
        for (int s = 0; s <= N; s++) {
            const int r = rb_[s];
            const double* Ld = sc + PL.Ld + (size_t)s * RR;
            if (s >= 1) {
                const int rp = rb_[s - 1];
                const double* Lo = sc + PL.Lo + (size_t)s * RR;
                for (int i = tid; i < r; i += NT) {
                    double acc = 0.0;
                    for (int l = 0; l < rp; l++) acc = fma(Lo[i * rp + l], t[(s - 1) * R + l], acc);
                    sm.vec[i] = b[s * R + i] - acc;
                }
            } else {
                for (int i = tid; i < r; i += NT) sm.vec[i] = b[i];
            }
            for (int e = tid; e < r * r; e += NT) sm.A[e] = Ld[e];   // the block's factor: shared memory for the serial solve
            for (int i = tid; i < r; i += NT) sm.vec[R + i] = sc[PL.rd + (size_t)s * R + i];
            __syncthreads();
            // triangular solve by the first warp: row i's sum runs over l = 0..i-1 in order, as in the oracle; the division by
            // the diagonal goes through its exact reciprocal (div_by: bitwise the IEEE quotient, a third of its latency)
            if (tid < 32) {
                double v0 = (tid < r) ? sm.vec[tid] : 0.0, v1 = (tid + 32 < r) ? sm.vec[tid + 32] : 0.0;
                for (int l = 0; l < r; l++) {
                    double tl = (l < 32) ? v0 : v1;
                    tl = div_by(__shfl_sync(0xffffffffu, tl, l & 31), sm.A[l * r + l], sm.vec[R + l]);
                    if (tid == (l & 31)) { if (l < 32) v0 = tl; else v1 = tl; }
                    if (tid > l && tid < r) v0 = fma(-sm.A[tid * r + l], tl, v0);
                    if (tid + 32 > l && tid + 32 < r) v1 = fma(-sm.A[(tid + 32) * r + l], tl, v1);
                }
                if (tid < r) t[s * R + tid] = v0;
                if (tid + 32 < r) t[s * R + tid + 32] = v1;
            }
            __syncthreads();
        }
        // backward
        for (int s = N; s >= 0; s--) {
            const int r = rb_[s];
            const double* Ld = sc + PL.Ld + (size_t)s * RR;
            if (s < N) {
                const int rn = rb_[s + 1];
                const double* Lo = sc + PL.Lo + (size_t)(s + 1) * RR;
                for (int i = tid; i < r; i += NT) {
                    double acc = 0.0;
                    for (int l = 0; l < rn; l++) acc = fma(Lo[l * r + i], x[(s + 1) * R + l], acc);
                    sm.vec[i] = t[s * R + i] - acc;
                }
            } else {
                for (int i = tid; i < r; i += NT) sm.vec[i] = t[s * R + i];
            }
            for (int e = tid; e < r * r; e += NT) sm.A[e] = Ld[e];
            for (int i = tid; i < r; i += NT) sm.vec[R + i] = sc[PL.rd + (size_t)s * R + i];
            __syncthreads();
            if (tid < 32) {
                double v0 = (tid < r) ? sm.vec[tid] : 0.0, v1 = (tid + 32 < r) ? sm.vec[tid + 32] : 0.0;
                for (int l = r - 1; l >= 0; l--) {
                    double xl = (l < 32) ? v0 : v1;
                    xl = div_by(__shfl_sync(0xffffffffu, xl, l & 31), sm.A[l * r + l], sm.vec[R + l]);
                    if (tid == (l & 31)) { if (l < 32) v0 = xl; else v1 = xl; }
                    if (tid < l) v0 = fma(-sm.A[l * r + tid], xl, v0);
                    if (tid + 32 < l) v1 = fma(-sm.A[l * r + tid + 32], xl, v1);
                }
                if (tid < r) x[s * R + tid] = v0;
                if (tid + 32 < r) x[s * R + tid + 32] = v1;
            }
            __syncthreads();
        }
    }
    // So(s)(i, j): S_{s,s-1}, only the n dynamics columns are non-zero; read from the stored rows of Y would need G again, so the
    // off-diagonal products use Lo Lp' = So exactly?  No: So is rebuilt from the Jacobians in scratch (AB, Gc), same expression.
    __device__ double So_entry(int s, int i, int j) {
        // G_s(i, x_j): block s >= 1, column j of its first n columns
        double g;
        if (s < N) {
            const int k = s - 1;
            if (i < n) g = ws[L.Z + (size_t)k * C::ZA + i * C::LDZ + j];
            else g = sc[PL.Gc + (size_t)k * LS_PN_PA * (n + m) + (i - n) * (n + m) + j];
        } else {
            g = sc[PL.Gc + (size_t)(N - 1) * LS_PN_PA * (n + m) + i * (n + m) + j];
        }
        const double gprev = (s - 1 == 0) ? 1.0 : -1.0;
        return g * (gprev / hx(s - 1, j));
    }
    // out = S x (oracle_pn.hpp pn_apply_S)
    __device__ void apply_S(const double* x, double* out) {
        const int* rb_ = rblk();
        for (int e = tid; e < (N + 1) * R; e += NT) {
            const int s = e / R, i = e - s * R;
            const int r = rb_[s];
            if (i >= r) continue;
            double acc = 0.0;
            if (s >= 1) {
                const int rp = rb_[s - 1];
                for (int l = 0; l < rp; l++) acc = fma((l < n) ? So_entry(s, i, l) : 0.0, x[(s - 1) * R + l], acc);
            }
            const double* Sd = sc + PL.Sd + (size_t)s * RR;
            for (int l = 0; l < r; l++) acc = fma(Sd[i * r + l], x[s * R + l], acc);
            if (s < N) {
                const int rn = rb_[s + 1];
                for (int l = 0; l < rn; l++) acc = fma((i < n) ? So_entry(s + 1, l, i) : 0.0, x[(s + 1) * R + l], acc);
            }
            out[e] = acc;
        }
        __syncthreads();
    }
    // reg_solve(S, y, Sreg, 1e-8, 25) (:314-333)
    __device__ void reg_solve(const double* b, double* x) {
        double* rr = sc + PL.rr;
        double* sx = sc + PL.sx;
        double* dx = sc + PL.dx;
        const int* rb_ = rblk();
        solve_factor(b, x);
        for (int count = 0; count < 25; count++) {
            apply_S(x, sx);
            // r = b - S x; norm(r) summed in element order by one thread (the order fixes the value; cheap: ~1.7k adds)
            for (int e = tid; e < (N + 1) * R; e += NT) {
                const int s = e / R, i = e - s * R;
                rr[e] = (i < rb_[s]) ? b[e] - sx[e] : 0.0;
            }
            __syncthreads();
            if (tid == 0) {
                double nrm = 0.0;
                for (int s = 0; s <= N; s++) for (int i = 0; i < rb_[s]; i++) nrm += rr[s * R + i] * rr[s * R + i];
                sm.flag = (sqrt(nrm) < 1e-8) ? 1 : 0;
            }
            __syncthreads();
            const int done = sm.flag;
            __syncthreads();
            if (done) break;
            solve_factor(rr, dx);
            for (int e = tid; e < (N + 1) * R; e += NT) {
                const int s = e / R, i = e - s * R;
                if (i < rb_[s]) x[e] += dx[e];
            }
            __syncthreads();
        }
    }

    // _projection_linesearch! (:272-312).  Returns the new violation; threw = the reference would raise (:304)
    __device__ double linesearch(bool& threw) {
        double* y = sc + PL.y;
        double* dl = sc + PL.dl;
        const int* rb_ = rblk();
        const double viol0 = gather_y(y);
        reg_solve(y, dl);
        // Z_ = Z + 1.0 * (-(H \ Y') dlambda)
        double* Xn = sc + PL.Xn;
        double* Un = sc + PL.Un;
        for (int e = tid; e < N * n; e += NT) {
            const int k = e / n, c = e - k * n;
            double acc = 0.0;
            if (k == 0) acc = fma(1.0, dl[c], acc);
            else acc = fma(-1.0, dl[k * R + c], acc);
            const int s = k + 1, r = rb_[s];
            for (int i = 0; i < r; i++) {
                double g;
                if (s < N) g = (i < n) ? ws[L.Z + (size_t)k * C::ZA + i * C::LDZ + c] : sc[PL.Gc + (size_t)k * LS_PN_PA * (n + m) + (i - n) * (n + m) + c];
                else g = sc[PL.Gc + (size_t)k * LS_PN_PA * (n + m) + i * (n + m) + c];
                acc = fma(g, dl[s * R + i], acc);
            }
            Xn[e] = ws[L.X + e] + 1.0 * (-(acc / hx(k, c)));
        }
        for (int e = tid; e < (N - 1) * m; e += NT) {
            const int k = e / m, c = e - k * m;
            double acc = 0.0;
            const int s = k + 1, r = rb_[s];
            for (int i = 0; i < r; i++) {
                const double g = (i < n) ? ws[L.Z + (size_t)k * C::ZA + i * C::LDZ + n + c] : sc[PL.Gc + (size_t)k * LS_PN_PA * (n + m) + (i - n) * (n + m) + n + c];
                acc = fma(g, dl[s * R + i], acc);
            }
            Un[e] = ws[L.U + e] + 1.0 * (-(acc / hu(c)));
        }
        __syncthreads();
        residuals(Xn, Un);
        const double viol = gather_y(y);
        if (!(viol < viol0)) { threw = true; return viol; }
        for (int e = tid; e < N * n; e += NT) ws[L.X + e] = Xn[e];
        for (int e = tid; e < (N - 1) * m; e += NT) ws[L.U + e] = Un[e];
        __syncthreads();
        return viol;
    }

    // _projection_solve! (:221-270); status: 0 ok, 1 the reference throws, 2 too many active rows for the block factor
    __device__ double projection_once(double feas_tol, double act_tol, int& status) {
        residuals(ws + L.X, ws + L.U);
        if (!active_set(act_tol)) { status = 2; return 0.0; }
        double viol_prev = gather_y(sc + PL.y);
        if (!factor(1e-2)) { status = 1; return viol_prev; }
        for (int count = 0; count < 10; count++) {
            bool threw = false;
            const double viol = linesearch(threw);
            if (threw) { status = 1; return viol; }
            const double rate = log10(viol) / log10(viol_prev);
            viol_prev = viol;
            if (rate < 1.1 || viol < feas_tol) break;
        }
        return viol_prev;
    }
};

struct PnOptsDev {
    int n_steps;
    double feas_tol, act_tol;
};

template <class C>
__global__ void __launch_bounds__(LS_PN_THREADS, LS_PN_MINB) ls_pn_kernel(const DevProblem Pg, const DevBatch Bt, const LsCtl lc, const PnOptsDev po,
                                                           double* scratch, const unsigned long long scratch_stride) {
    constexpr int n = C::n, m = C::m, NT = LS_PN_THREADS;
    extern __shared__ __align__(16) unsigned char pn_smem_raw[];
    DevProblem P = Pg;
    ls_stage_problem(P, Pg, pn_smem_raw);
    PnSmem<C>& sm = *reinterpret_cast<PnSmem<C>*>(pn_smem_raw + (((size_t)ls_tab_bytes(Pg.N, Pg.nrows) + 15) & ~(size_t)15));
    const int tid = threadIdx.x;
    const int N = P.N;
    double* sc = scratch + (size_t)blockIdx.x * scratch_stride;
    for (int b = blockIdx.x; b < Bt.B; b += gridDim.x) {
        TOResult res = Bt.res[b];
        // the AL solve of this problem ended the way the reference throws: no polish (altro_methods.jl:26 propagates the error)
        if (res.status & (TO_STATUS_COST_INCREASED | TO_STATUS_NOT_PD_SQRT | TO_STATUS_REG_DIVERGED)) continue;
        double* ws = lc.ws + (size_t)b * lc.ws_stride;
        PnCtx<C> cx(P, sm, ws, sc, Bt.x0 + (size_t)b * C::n0, tid);
        // keep the AL solution: on failure X, U stay what they were (the copy back happens after the Newton step, :8-11)
        for (int e = tid; e < N * n; e += NT) sc[cx.PL.Xb + e] = ws[cx.L.X + e];
        for (int e = tid; e < (N - 1) * m; e += NT) sc[cx.PL.Ub + e] = ws[cx.L.U + e];
        __syncthreads();
        int status = 0;
        double J = 0.0, cmax = 0.0;
        for (int step = 0; step < po.n_steps && status == 0; step++) {
            cx.residuals(ws + cx.L.X, ws + cx.L.U);
            if (!cx.active_set(po.act_tol)) { status = 2; break; }
            double viol = cx.gather_y(sc + cx.PL.y);
            for (int count = 0; count < 10 && viol > po.feas_tol && status == 0; count++) viol = cx.projection_once(po.feas_tol, po.act_tol, status);
            if (status != 0) break;
            // record_iteration! (:21-29): cost(prob), max_violation(prob) (src/problem.jl:242-267) -- sequential sums by one thread
            if (tid == 0) {
                double Jc = 0.0, cm = 0.0;
                for (int k = 0; k < N; k++) {
                    double x[n], u[m];
#pragma unroll
                    for (int i = 0; i < n; i++) x[i] = ws[cx.L.X + (size_t)k * n + i];
#pragma unroll
                    for (int i = 0; i < m; i++) u[i] = (k < N - 1) ? ws[cx.L.U + (size_t)k * m + i] : 0.0;
                    Jc += (k < N - 1) ? stage_cost<C>(P, x, u) : term_cost<C>(P, x);
                    const int rb = P.knot_row_begin[k], rc = P.knot_row_count[k];
                    double e_ = 0.0, mi = 0.0;
                    for (int i = 0; i < rc; i++) {
                        const DevRow r = P.rows[rb + i];
                        const double c = row_value<C>(r, x, u);
                        if (r.eq) e_ = dmax(e_, fabs(c));
                        mi = dmax(mi, dmax(0.0, c));
                    }
                    if (rc > 0) cm = dmax(cm, dmax(e_, mi));
                }
                sm.vec[0] = Jc;
                sm.vec[1] = cm;
            }
            __syncthreads();
            J = sm.vec[0];
            cmax = sm.vec[1];
            __syncthreads();
            if (cmax <= po.feas_tol) break;
        }
        if (status == 0) {
            // copy the polished trajectory out (process_results!, altro_methods.jl:60-61)
            const int no = Bt.n_out, mo = Bt.m_out;
            for (int e = tid; e < N * no; e += NT) {
                const int k = e / no, i = e - k * no;
                Bt.X[(size_t)b * N * no + e] = ws[cx.L.X + (size_t)k * n + i];
            }
            for (int e = tid; e < (N - 1) * mo; e += NT) {
                const int k = e / mo, i = e - k * mo;
                Bt.U[(size_t)b * (N - 1) * mo + e] = ws[cx.L.U + (size_t)k * m + i];
            }
            if (tid == 0) { res.J = J; res.c_max = cmax; Bt.res[b] = res; }
        } else {
            for (int e = tid; e < N * n; e += NT) ws[cx.L.X + e] = sc[cx.PL.Xb + e];
            for (int e = tid; e < (N - 1) * m; e += NT) ws[cx.L.U + e] = sc[cx.PL.Ub + e];
            if (tid == 0) { res.status |= (status == 2) ? TO_STATUS_PN_SKIPPED : TO_STATUS_PN_FAILED; Bt.res[b] = res; }
        }
        __syncthreads();
    }
}

}  // namespace tob
