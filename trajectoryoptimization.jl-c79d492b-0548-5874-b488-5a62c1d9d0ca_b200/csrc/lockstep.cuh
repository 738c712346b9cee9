// B200 engine, throughput path: the batch advances in LOCKSTEP, one iLQR iteration ("tick") at a
// time, through phase kernels that each use the parallelism natural to their phase:
//
//   ls_jac_kernel     thread per (problem, knot, partial-chunk): dual-number rk3 Jacobians        (src/model.jl:491-512)
//   ls_bp_kernel      a GROUP of GS lanes per problem (2 problems per warp for the quadrotor):
//                     lane j owns column j of S, A, Qxx, Qux, K in REGISTERS; the left operands
//                     (A, B, T, Tu, K, KQ) stream from shared memory as broadcast loads; every
//                     inner product is the same sequential FMA chain as in the oracle           (backward_pass.jl:9-85)
//   ls_trial_kernel   thread per (problem, step size): closed-loop rk3 rollout + AL cost, G step
//                     sizes per launch, first accepted wins; unaccepted problems go to a retry
//                     list for the next G step sizes                                             (forward_pass.jl:5-85)
//   ls_accept_kernel  thread per problem: re-rolls the accepted step in place (X <- X̄, U <- Ū),
//                     Todorov gradient, iteration record, convergence tests                      (ilqr_methods.jl:30-45,77-162)
//   ls_outer_kernel   warp per problem whose inner solve ended this tick: AL dual / penalty /
//                     active-set update, outer convergence, start of the next inner solve        (augmented_lagrangian_methods.jl:2-126)
//
// Problems leave the active list when they finish (ragged iteration counts), so every tick only
// touches live problems; the host just replays the tick until the list is empty.  Per-problem
// arithmetic is identical to the warp-persistent engine (engine.cuh) and to the CPU oracle: the
// assignment of work to lanes differs, the expression trees do not.
#pragma once
#include <cstdio>
#include <cstdlib>
#include "engine.cuh"
#include "sqrt_bp.cuh"

namespace tob {

template <class C> __host__ __device__ constexpr int ls_group_size() {
    return (C::n > 8 || C::m > 8) ? 16 : ((C::n > 4 || C::m > 4) ? 8 : 4);
}

// ------------------------------------------------------------------------------------------
// warp-level slow path: everything that is not the per-iteration hot loop reuses Solver<C>
// ------------------------------------------------------------------------------------------
template <class C>
struct LsSolver : Solver<C> {
    typedef Solver<C> S_;
    LsState* st;
    int outer_i, al_iterations, al_total, inner_i;
    double J_prev, Jout, cmax_out;

    __device__ LsSolver(const DevProblem& P_, const DevBatch& B_, const DevCtl& c_, Smem<C>& s_, double* ws_, int lane_)
        : S_(P_, B_, c_, s_, ws_, lane_, false) {}

    __device__ void load(LsState* s, int b_) {
        st = s;
        this->b = b_;
        this->load_x0();
        this->io = this->ctl.o.opts_uncon;
        this->io.cost_tolerance = s->cost_tol;
        this->io.gradient_norm_tolerance = s->grad_tol;
        this->al_on = (this->ctl.mode == 1);
        this->rho = s->rho; this->drho = s->drho;
        this->iterations = s->iterations; this->dJ_zero = s->dJ_zero; this->steps = s->steps; this->status = s->status;
        this->last_dJ = s->last_dJ; this->last_grad = s->last_grad; this->last_cost = s->last_cost;
        this->fp_expected = s->fp_expected; this->fp_z = s->fp_z; this->fp_alpha = s->fp_alpha;
        this->outer_idx = (s->outer_i > 0) ? s->outer_i - 1 : 0;
        this->n_inner_rec = s->n_inner_rec; this->n_outer_rec = s->n_outer_rec;
        this->ls_count = s->ls_count;
        outer_i = s->outer_i; al_iterations = s->al_iterations; al_total = s->al_total; inner_i = s->inner_i;
        J_prev = s->J_prev; Jout = s->Jout; cmax_out = s->cmax;
    }
    __device__ void store() {
        if (this->lane == 0) {
            LsState* s = st;
            s->cost_tol = this->io.cost_tolerance; s->grad_tol = this->io.gradient_norm_tolerance;
            s->rho = this->rho; s->drho = this->drho;
            s->iterations = this->iterations; s->dJ_zero = this->dJ_zero; s->steps = this->steps; s->status = this->status;
            s->last_dJ = this->last_dJ; s->last_grad = this->last_grad; s->last_cost = this->last_cost;
            s->fp_expected = this->fp_expected; s->fp_z = this->fp_z; s->fp_alpha = this->fp_alpha;
            s->n_inner_rec = this->n_inner_rec; s->n_outer_rec = this->n_outer_rec;
            s->ls_count = this->ls_count;
            s->outer_i = outer_i; s->al_iterations = al_iterations; s->al_total = al_total; s->inner_i = inner_i;
            s->J_prev = J_prev; s->Jout = Jout; s->cmax = cmax_out;
            s->winner = -1;
        }
        __syncwarp();
    }

    // first part of ilqr_solve (ilqr_methods.jl:3-20): reset, rollout, initial cost and record
    __device__ void begin_inner() {
        this->iterations = 0; this->dJ_zero = 0; this->rho = 0.0; this->drho = 0.0;
        this->fp_expected = 0.0; this->fp_z = 0.0; this->fp_alpha = 0.0;
        if (!this->all_finite_X()) this->rollout_open(false);
        J_prev = this->eval_cost();
        this->record_inner(J_prev, __longlong_as_double(0x7ff0000000000000LL));
        inner_i = 1;
    }
    __device__ void set_tolerances() {  // augmented_lagrangian_methods.jl:39-50
        const TOALOptions& o = this->ctl.o;
        if (outer_i != o.iterations) {
            this->io.cost_tolerance = o.cost_tolerance_intermediate;
            this->io.gradient_norm_tolerance = o.gradient_norm_tolerance_intermediate;
        } else {
            this->io.cost_tolerance = o.cost_tolerance;
            this->io.gradient_norm_tolerance = o.gradient_norm_tolerance;
        }
    }
    __device__ void finish() {
        if (this->ctl.mode == 0) this->epilogue(this->last_cost, 0.0, 0, this->iterations);
        else this->epilogue(Jout, cmax_out, al_iterations, al_total);
    }
    __device__ double max_penalty_lex() {  // maximum(maximum(mu)) over a vector of vectors (Q20)
        double pm = 0.0;
        const DevProblem& P = this->P;
        if (this->lane == 0) {
            int best = -1;
            for (int k = 0; k < P.N; k++) {
                if (best < 0) { best = k; continue; }
                const int ca = P.knot_row_count[best], cb_ = P.knot_row_count[k];
                const double* a = this->mu() + P.knot_lam_off[best];
                const double* c = this->mu() + P.knot_lam_off[k];
                bool less = false, decided = false;
                for (int q = 0; q < ca && q < cb_; q++) {
                    if (a[q] < c[q]) { less = true; decided = true; break; }
                    if (c[q] < a[q]) { decided = true; break; }
                }
                if (!decided) less = ca < cb_;
                if (less) best = k;
            }
            if (best >= 0 && P.knot_row_count[best] > 0) {
                const double* a = this->mu() + P.knot_lam_off[best];
                pm = a[0];
                for (int q = 1; q < P.knot_row_count[best]; q++) pm = dmax(pm, a[q]);
            }
        }
        return bcast(pm, 0);
    }

    // problem start.  Returns true if the problem needs iLQR steps (goes on the active list).
    __device__ bool start(int b_) {
        const DevProblem& P = this->P;
        const TOALOptions& o = this->ctl.o;
        this->prologue(b_);
        const size_t nk = (size_t)(P.N - 1) * C::KDS;
        for (size_t e = this->lane; e < nk; e += 32) this->ws[this->L.KD + e] = 0.0;
        __syncwarp();
        al_iterations = 0; al_total = 0; Jout = 0.0; cmax_out = 0.0; outer_i = 0; inner_i = 0; J_prev = 0.0;
        this->fp_expected = 0.0; this->fp_z = 0.0; this->fp_alpha = 0.0;
        this->rho = 0.0; this->drho = 0.0; this->iterations = 0; this->dJ_zero = 0;
        this->outer_idx = 0;
        if (this->ctl.mode == 0) {
            this->al_on = false;
            begin_inner();
            if (this->io.iterations >= 1) return true;
            finish();
            return false;
        }
        // al_solve prologue (augmented_lagrangian_methods.jl:2-15)
        this->al_on = true;
        for (int e = this->lane; e < P.Ptot; e += 32) { this->lam()[e] = 0.0; this->mu()[e] = o.penalty_initial; }
        __syncwarp();
        if (!this->all_finite_X()) this->rollout_open(false);
        const double J0 = this->eval_cost();
        const double cmax = this->max_violation();
        this->record_outer(J0, cmax, al_iterations, al_total);
        Jout = J0; cmax_out = cmax;
        if (o.iterations < 1) {
            this->status |= TO_STATUS_MAX_OUTER;
            finish();
            return false;
        }
        outer_i = 1;
        this->outer_idx = 0;
        set_tolerances();
        begin_inner();
        if (this->io.iterations >= 1) return true;
        return after_inner(true);
    }

    // the inner solve of this problem ended (ok = false: the reference would have thrown).
    // Returns true if another inner solve was started.
    __device__ bool after_inner(bool ok) {
        const TOALOptions& o = this->ctl.o;
        for (;;) {
            if (this->ctl.mode == 0 || !ok) { finish(); return false; }
            // rest of step! (augmented_lagrangian_methods.jl:53-67), record, convergence
            const double J = this->eval_cost();
            const double cmax = this->max_violation();
            this->dual_penalty_update();
            this->record_outer(J, cmax, al_iterations, al_total);
            Jout = J; cmax_out = cmax;
            bool converged = false;
            if (o.kickout_max_penalty) converged = (max_penalty_lex() == o.penalty_max);
            converged = converged || (cmax < o.constraint_tolerance);
            if (converged) { finish(); return false; }
            this->iterations = 0; this->dJ_zero = 0; this->rho = 0.0; this->drho = 0.0;
            outer_i += 1;
            if (outer_i > o.iterations) {
                this->status |= TO_STATUS_MAX_OUTER;
                finish();
                return false;
            }
            this->outer_idx = outer_i - 1;
            set_tolerances();
            begin_inner();
            if (this->io.iterations >= 1) return true;
            ok = true;  // an inner solve with zero allowed steps ends immediately
        }
    }
};

// Per-block shared-memory copy of the problem's knot tables (first row / row count / multiplier offset per knot) and,
// when they are few, of the constraint rows: every knot of every live problem looks these up, and a chain of dependent
// global loads per knot is what a latency-bound kernel can least afford.  The copy is reached through the same
// DevProblem fields (generic pointers into shared memory).
constexpr int LS_ROWCAP = 256;  // 14 KB of constraint rows (car_escape: ~190 distinct rows)
__host__ __device__ inline int ls_tab_bytes(int N, int nrows) {
    const int tab = ((4 * N + 1) * 4 + 15) & ~15;
    return tab + ((nrows <= LS_ROWCAP) ? nrows * (int)sizeof(DevRow) : 0);
}
__device__ inline void ls_stage_problem(DevProblem& Pl, const DevProblem& P, unsigned char* smem_tab) {
    const int N = P.N;
    int* tab = reinterpret_cast<int*>(smem_tab);
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        tab[i] = P.knot_row_begin[i];
        tab[N + i] = P.knot_row_count[i];
    }
    for (int i = threadIdx.x; i <= N; i += blockDim.x) tab[2 * N + i] = P.knot_lam_off[i];
    for (int i = threadIdx.x; i < N; i += blockDim.x) tab[3 * N + 1 + i] = P.knot_cols[i];
    Pl.knot_row_begin = tab;
    Pl.knot_row_count = tab + N;
    Pl.knot_lam_off = tab + 2 * N;
    Pl.knot_cols = tab + 3 * N + 1;
    if (P.nrows <= LS_ROWCAP) {
        DevRow* rows = reinterpret_cast<DevRow*>(smem_tab + (((4 * N + 1) * 4 + 15) & ~15));
        for (int i = threadIdx.x; i < P.nrows; i += blockDim.x) rows[i] = P.rows[i];
        Pl.rows = rows;
    }
    __syncthreads();
}

__device__ __forceinline__ void ls_append(int* list, unsigned int* count, int b) {
    const unsigned int pos = atomicAdd(count, 1u);
    list[pos] = b;
}

// ---- init: prologue + first inner-solve setup for every problem of the batch ----------------
template <class C>
__global__ void __launch_bounds__(32) ls_init_kernel(const DevProblem P, const DevBatch Bt, const DevCtl ctl, const LsCtl lc) {
    __shared__ Smem<C> sm;
    const int lane = threadIdx.x;
    for (int b = blockIdx.x; b < Bt.B; b += gridDim.x) {
        LsSolver<C> s(P, Bt, ctl, sm, lc.ws + (size_t)b * lc.ws_stride, lane);
        s.st = &lc.st[b];
        const bool active = s.start(b);
        s.store();
        if (active && lane == 0) ls_append(lc.list[0], &lc.counts[0], b);
    }
}

// ---- outer: problems whose inner solve ended in this tick ------------------------------------
template <class C>
__global__ void __launch_bounds__(32) ls_outer_kernel(const DevProblem Pg, const DevBatch Bt, const DevCtl ctl, const LsCtl lc, const int cur) {
    __shared__ Smem<C> sm;
    extern __shared__ __align__(16) unsigned char ls_tab_raw[];
    const int lane = threadIdx.x;
    const unsigned int n = lc.counts[4];
    if (blockIdx.x >= n) return;
    // the three passes over the constraint rows (cost, violation, dual update) walk the row table: from shared memory, not as
    // dependent global loads (car_escape, 17,900 rows per problem: 1.2 ms per tick for ONE finishing problem before)
    DevProblem P = Pg;
    ls_stage_problem(P, Pg, ls_tab_raw);
    for (unsigned int a = blockIdx.x; a < n; a += gridDim.x) {
        const int b = lc.outer_list[a];
        LsSolver<C> s(P, Bt, ctl, sm, lc.ws + (size_t)b * lc.ws_stride, lane);
        s.load(&lc.st[b], b);
        const bool active = s.after_inner(lc.st[b].inner_ok != 0);
        s.store();
        if (active && lane == 0) ls_append(lc.list[cur ^ 1], &lc.counts[cur ^ 1], b);
    }
}

// ---- Jacobians: thread per (problem, knot, chunk of partial directions) -----------------------
// Leading state directions the continuous dynamics do not depend on (quadrotor: the position r, dynamics/quadrotor.jl:
// only q, v, omega enter f).  For those the dual-number rk3 yields the unit column exactly (every partial is 0*finite,
// and x.p + (+-0) = x.p), so the Jacobian kernel writes e_c instead of pushing a dual through the integrator -- unless
// an input is not finite, in which case the directions are computed the long way to keep NaN patterns identical.
template <int MODEL> struct TrivialDirs { static constexpr int value = 0; };
template <> struct TrivialDirs<4> { static constexpr int value = 3; };

// one work item of the Jacobian phase: knot k of the problem with workspace ws, chunk `ch` of PC partial directions
template <class C, int PC> __host__ __device__ constexpr int ls_jac_chunks() {
    return (C::PT - ((!C::MT) ? TrivialDirs<C::MODEL>::value : 0) + PC - 1) / PC;
}
template <class C, int PC>
__device__ __forceinline__ void ls_jac_item(const DevProblem& P, double* ws, const WsLayout& L, const int k, const int ch) {
    constexpr int TZ = (!C::MT) ? TrivialDirs<C::MODEL>::value : 0;
    typedef Dual<PC> D;
    {
        const double* xk = ws + L.X + (size_t)k * C::n;
        const double* uk = ws + L.U + (size_t)k * C::m;
        double* ab = ws + L.Z + (size_t)k * C::ZA;
        int nrep = 1;
        // Finite check of the knot's [x;u] (only the lane with ch == 0 needs the verdict).  All loads are issued together and
        // combined without short-circuit: `fin = fin && isfinite(xk[i])` made that one lane walk 17 DEPENDENT DRAM loads while
        // its warp waited (36 % of the kernel's samples, profiles/r01y source view).  Every lane runs it, so the warp waits
        // once, together, and the values every lane needs below are in L1 by then.  They are not kept in registers: 17 more
        // live registers cost more in spills than the reloads (measured, profiles/r01u).
        bool fin = true;
        if (TZ > 0) {
#pragma unroll
            for (int i = 0; i < C::n0; i++) fin = fin & isfinite(xk[i]);
#pragma unroll
            for (int i = 0; i < C::m0; i++) fin = fin & isfinite(uk[i]);
        }
        if (TZ > 0 && ch == 0) {
            if (fin) {
#pragma unroll
                for (int c = 0; c < TZ; c++)
#pragma unroll
                    for (int i = 0; i < C::n0; i++) ab[i * C::LDZ + c] = (i == c) ? 1.0 : 0.0;
            } else {
                nrep = 1 + (TZ + PC - 1) / PC;
            }
        }
      for (int rep = 0; rep < nrep; rep++) {
        const int s0 = (rep == 0) ? (TZ + ch * PC) : (rep - 1) * PC;  // long way: directions 0..TZ-1, PC at a time (c < TZ guard below)
        D xs[C::n0], us[C::m0], dts, xn[C::n0];
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
            const double h = uk[C::m - 1];
            dt = h * h;
        }
        dts = D(dt);
        if constexpr (C::MT) {
#pragma unroll
            for (int j = 0; j < PC; j++) if (s0 + j == C::n0 + C::m0) dts.p[j] = 1.0;
        }
        fd_model<C::MODEL, C::INTEG, D>(xn, xs, us, dts);
        // row-major [A_k B_k] of the AUGMENTED model, rows of C::LDZ doubles
        // (add_slack_controls: src/model.jl:761-779; add_min_time_controls: minimum_time.jl:85-104)
        double h2 = 0.0;
        if constexpr (C::MT) h2 = 2.0 * uk[C::m - 1];
#pragma unroll
        for (int j = 0; j < PC; j++) {
            const int c = s0 + j;
            if (c < C::PT && (rep == 0 || c < TZ)) {
                const int col = (c < C::n0) ? c : ((c < C::n0 + C::m0) ? C::n + (c - C::n0) : C::n + C::m - 1);
                const bool is_dt = (C::MT && c == C::n0 + C::m0);
#pragma unroll
                for (int i = 0; i < C::n0; i++) ab[i * C::LDZ + col] = is_dt ? xn[i].p[j] * h2 : xn[i].p[j];
            }
        }
      }
        if (ch == 0 && (C::INF || C::MT)) {
            // the constant entries of the augmented Jacobian
            for (int i = 0; i < C::n; i++)
                for (int jj = 0; jj < C::n + C::m; jj++) {
                    const bool dyn = (i < C::n0) && (jj < C::n0 || (jj >= C::n && jj < C::n + C::m0) || (C::MT && jj == C::n + C::m - 1));
                    if (dyn) continue;
                    double v = 0.0;
                    if (jj >= C::n) {
                        const int bcol = jj - C::n;
                        if (C::INF && bcol >= C::m0 && bcol < C::m0 + C::n0) v = (i == bcol - C::m0) ? 1.0 : 0.0;
                        else if (C::MT && bcol == C::m - 1 && i == C::n - 1) v = 1.0;
                    }
                    ab[i * C::LDZ + jj] = v;
                }
        }
    }
}

template <class C, int PC, int MINB>
__global__ void __launch_bounds__(128, MINB) ls_jac_kernel(const DevProblem P, const LsCtl lc, const int cur) {
    constexpr int NCH = ls_jac_chunks<C, PC>();
    const unsigned int na = lc.counts[cur];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        lc.counts[cur ^ 1] = 0; lc.counts[2] = 0; lc.counts[3] = 0; lc.counts[4] = 0; lc.counts[5] = 0;
        *reinterpret_cast<unsigned long long*>(&lc.counts[8]) += na;  // iLQR iterations served by lockstep ticks (bench roofline)
    }
    const int N = P.N;
    const unsigned int per = (unsigned int)(N - 1) * NCH;
    const unsigned long long items = (unsigned long long)na * per;
    const WsLayout L = ws_layout<C>(N, P.Ptot, false);
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < items; t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned int a = (unsigned int)(t / per), it = (unsigned int)(t - (unsigned long long)a * per);
        const int b = lc.list[cur][a];
        const int k = it / NCH, ch = it - k * NCH;
        ls_jac_item<C, PC>(P, lc.ws + (size_t)b * lc.ws_stride, L, k, ch);
    }
}

// ------------------------------------------------------------------------------------------
// backward pass: GS lanes per problem, lane j owns column j
// ------------------------------------------------------------------------------------------
template <class C>
struct alignas(16) BpSmem {
    static constexpr int n = C::n, m = C::m;
    static constexpr int LDn = (n + 1) & ~1;   // even leading dimension: rows start 16-byte aligned
    static constexpr int LDm = (m + 1) & ~1;
    static constexpr int LDZ = C::LDZ;
    static constexpr int XU = (n + m + 1) & ~1;
    static constexpr int LAMCAP = 16;          // multipliers / penalties of one knot staged by the prefetch (larger sets read global)
    // per-knot inputs, filled by cp.async while the previous knot's factorisation / cost-to-go update runs
    double AB[n * LDZ];      // row l = [A(l,0..n-1) B(l,0..m-1)]
    double xu[XU];           // [x_k ; u_k]
    double lam[LAMCAP], mu[LAMCAP];
    // products streamed by the other lanes of the group (all "left operand, contiguous in the output row index")
    double T[n * LDn];       // T(i,l) at [l*LDn+i]; later the unsymmetrised S.xx
    double Tu[n * LDm];      // Tu(i,l) at [l*LDm+i]
    double KT[m * LDn];      // K(l,i)   at [l*LDn+i]
    double QuxT[m * LDn];    // Qux(l,i) at [l*LDn+i]
    double KQ[m * LDn];      // KQ(i,l)  at [l*LDn+i]
    double Quu[m * m], Sx[n], Qu[m], d[m], vQx[C::nq], vQu[C::mq], xN[n];
};
// byte stride between the shared-memory blocks of consecutive lane groups: ≡ 64 (mod 128), so the two groups of a
// warp never read different addresses that fall into the same banks
template <class C> __host__ __device__ constexpr int ls_bp_stride() {
    return (int)(((sizeof(BpSmem<C>) + 63) / 128) * 128 + 64);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
// L1 prefetch of the 128-byte lines that hold base[0..count), spread over the `nl` lanes that call it (lane index `lane`).
// Knots with more multipliers than the shared-memory staging holds (car_escape: 177 rows per knot) read lambda / mu from
// global memory inside a sequential row loop; without the prefetch every row pays an L2 round trip (355 cycles per row
// measured, profiles/r01e2), with it the loop runs out of L1.
__device__ __forceinline__ void prefetch_span(const double* base, int count, int lane, int nl) {
    const unsigned long long b = reinterpret_cast<unsigned long long>(base) & ~127ull;
    const unsigned long long e = reinterpret_cast<unsigned long long>(base + count);
    for (unsigned long long q = b + (unsigned long long)lane * 128ull; q < e; q += (unsigned long long)nl * 128ull)
        asm volatile("prefetch.global.L1 [%0];\n" ::"l"(q));
}

// shared memory of a lane group that only evaluates expansions (ls_resident_kernel): the minimum-time exchange vectors
template <class C>
struct alignas(16) BpExpSmem {
    static constexpr int XU = (C::n + C::m + 1) & ~1;
    double xu[XU], vQx[C::nq], vQu[C::mq];
};

// a / d from the correctly rounded reciprocal y = 1/d (an IEEE division done once, off the dependent chain) and two
// Newton-Markstein corrections: q1 is a faithful quotient, so r1 = a - d*q1 is exact and q2 = RN(a/d) by Markstein's theorem.
// BITWISE equal to a / d for operands in the safe exponent range (2^-300 .. 2^300, checked on 4.5e9 random and structured pairs
// by tools/check_divby.c); zeros keep their sign through q0; everything else (subnormal, huge, inf, nan, d = 0) takes the true
// division.  Five dependent FMAs (~45 cycles) instead of a DDIV (~126): the back-substitution of the LU solve divides by the
// pivots one after the other, and their reciprocals exist already (the factorisation scales the L column by them).
__device__ __forceinline__ double div_by(double a, double d, double y) {
    const unsigned ea = ((unsigned)__double2hiint(a) >> 20) & 0x7ffu, ed = ((unsigned)__double2hiint(d) >> 20) & 0x7ffu;
    const bool safe = (ed - 723u <= 600u) && ((ea - 723u <= 600u) || a == 0.0);
    if (!safe) return a / d;
    const double q0 = a * y;
    const double r0 = fma(-d, q0, a);
    const double q1 = fma(r0, y, q0);
    const double r1 = fma(-d, q1, a);
    const double q2 = fma(r1, y, q1);
    return (a == 0.0) ? q0 : q2;
}

template <class C, class SM = BpSmem<C>>
struct BpGroup {
    static constexpr int n = C::n, m = C::m, n0 = C::n0, m0 = C::m0, nq = C::nq, mq = C::mq;
    static constexpr int GS = ls_group_size<C>();
    static constexpr int DL = (n < GS) ? n : 0;  // lane that solves for the feed-forward term d
    static constexpr int LDn = BpSmem<C>::LDn, LDm = BpSmem<C>::LDm, LDZ = BpSmem<C>::LDZ, LAMCAP = BpSmem<C>::LAMCAP;
    const DevProblem& P;
    SM& sm;
    double* ws;
    const WsLayout L;
    const int j;            // lane within the group
    const unsigned gmask;   // lanes of this group
    const bool al_on;
    TOiLQROptions io;
    double rho, drho;
    // register-resident columns
    double Scol[n], Qxxc[n], Quxc[m], Quuc[m], Kcol[m];
    double Qx_j, Qu_j, Sx_j;
    // diagonal stage cost (every LQRObjective of the problem zoo): the lane's own diagonal / linear entries, loaded once
    bool fast_cost = false;
    double qjj = 0.0, qlin = 0.0, rjj = 0.0, rlin = 0.0;
    __device__ void load_cost_constants() {
        fast_cost = !C::MT && P.q_diag && P.r_diag && P.h_zero;
        if (fast_cost) {
            if (j < nq) { qjj = __ldg(&P.Q[j * nq + j]); qlin = __ldg(&P.q[j]); }
            if (j < mq) { rjj = __ldg(&P.R[j * mq + j]); rlin = __ldg(&P.r[j]); }
        }
    }
    // optional cycle profile of one group (diagnostics: to_debug_enable): cycles per section of the knot loop
    long long* prof = nullptr;
    long long pt0 = 0;
    __device__ void tick(int section) {
        if (prof) {
            const long long t = clock64();
            if (j == 0) prof[section] += t - pt0;
            pt0 = clock64();
        }
    }

    __device__ BpGroup(const DevProblem& P_, SM& s_, double* ws_, int j_, unsigned gmask_, bool al_on_, const TOiLQROptions& io_)
        : P(P_), sm(s_), ws(ws_), L(ws_layout<C>(P_.N, P_.Ptot, false)), j(j_), gmask(gmask_), al_on(al_on_), io(io_) {}

    __device__ void gsync() { __syncwarp(gmask); }

    __device__ void reg_update(bool increase) {  // ilqr_methods.jl:164-176
        const double f = io.bp_reg_increase_factor;
        if (increase) {
            drho = dmax(drho * f, f);
            rho = dmax(rho * drho, io.bp_reg_min);
        } else {
            drho = dmin(drho / f, 1.0 / f);
            rho = rho * drho * ((rho * drho > io.bp_reg_min) ? 1.0 : 0.0);
        }
    }

    // asynchronous copy of knot k's inputs ([A B], x, u, lambda, mu) into shared memory
    __device__ void prefetch(int k) {
        const double* ab = ws + L.Z + (size_t)k * C::ZA;  // 16-byte aligned, C::ZA even
        for (int e = 2 * j; e < C::ZA; e += 2 * GS) cp_async16(&sm.AB[e], ab + e);
        const double* xk = ws + L.X + (size_t)k * n;
        const double* uk = ws + L.U + (size_t)k * m;
        for (int e = j; e < n + m; e += GS) cp_async8(&sm.xu[e], (e < n) ? (xk + e) : (uk + (e - n)));
        if (al_on) {
            const int rc = P.knot_row_count[k], lo = P.knot_lam_off[k];
            if (rc <= LAMCAP) {
                for (int e = j; e < rc; e += GS) {
                    cp_async8(&sm.lam[e], ws + L.LAM + lo + e);
                    cp_async8(&sm.mu[e], ws + L.MU + lo + e);
                }
            }
        }
    }

    // constraint row value / Jacobian entry with z̄ = [x̄;ū] read from shared memory (same expressions as
    // row_value / row_jac in engine.cuh)
    __device__ static double row_value_s(const DevRow& r, const double* zs) {
        switch (r.kind) {
            case DR_LIN: {
                const double z = zs[r.col];
                return (r.sign > 0) ? (z - r.a) : (r.a - z);
            }
            case DR_CIRCLE: {
                const double dx = zs[0] - r.a, dy = zs[1] - r.b;
                return -(((dx * dx) + (dy * dy)) - (r.r * r.r));
            }
            case DR_SPHERE: {
                if constexpr (n >= 3) {
                    const double dx = zs[0] - r.a, dy = zs[1] - r.b, dz = zs[2] - r.c;
                    return -((((dx * dx) + (dy * dy)) + (dz * dz)) - (r.r * r.r));
                } else {
                    return 0.0;
                }
            }
            default: return zs[n + m - 1] - zs[n - 1];
        }
    }
    __device__ static double row_jac_s(const DevRow& r, const double* zs, int col) {
        switch (r.kind) {
            case DR_LIN: return (col == r.col) ? ((r.sign > 0) ? 1.0 : -1.0) : 0.0;
            case DR_CIRCLE: return (col == 0) ? -(2.0 * (zs[0] - r.a)) : ((col == 1) ? -(2.0 * (zs[1] - r.b)) : 0.0);
            case DR_SPHERE:
                if constexpr (n >= 3)
                    return (col == 0) ? -(2.0 * (zs[0] - r.a)) : ((col == 1) ? -(2.0 * (zs[1] - r.b)) : ((col == 2) ? -(2.0 * (zs[2] - r.c)) : 0.0));
                else
                    return 0.0;
            default: return (col == n + m - 1) ? 1.0 : ((col == n - 1) ? -1.0 : 0.0);
        }
    }

    // cost expansion of knot k (+ AL terms), column j -> registers
    // (src/cost.jl:183-198, minimum_time.jl:161-204, augmented_lagrangian_methods.jl:186-229)
    // `xs`: the knot's [x;u] in shared memory (u ignored at the terminal knot); `lams`/`mus`: its multipliers / penalties.
    // The AL sums are accumulated first (from zero, in row order) and the cost blocks added to them afterwards,
    // which is the same value as adding the finished sums to the blocks.  Bound / goal / slack rows (one ±1 entry)
    // only touch the diagonal element and the gradient entry of the lane that owns their column.
    __device__ void expansion(int k, const double* xs, const double* lams, const double* mus) {
        const int N = P.N;
        const bool term = (k == N - 1);
        const int rb = P.knot_row_begin[k], rc = al_on ? P.knot_row_count[k] : 0;
        const bool has_al = rc > 0;
        if (rc > LAMCAP) {  // multipliers / penalties come from global memory: fetch the knot's lines before the row loop
            prefetch_span(lams, rc, j, GS);
            prefetch_span(mus, rc, j, GS);
        }
        double dxx = 0.0, duu = 0.0, ax = 0.0, au = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++) Qxxc[i] = 0.0;
#pragma unroll
        for (int i = 0; i < m; i++) { Quxc[i] = 0.0; Quuc[i] = 0.0; }
        const int ct = has_al ? P.knot_cols[k] : -1;
        if (ct >= 0) {
            // every row of this knot has ONE +-1 entry (bounds, goal, slack equality): the lane walks the rows of its own
            // columns only (x_j, u_j), in row order -- the same sums as the row-by-row loop below, where each row is one
            // lane's work while the others wait (quadrotor: 4 of 505 warp instructions per knot were useful there)
            auto own = [&](const int r_, const double z, double& dd, double& aa) {
                if (r_ < 0) return;
                const DevRow* r = &P.rows[rb + r_];
                const double sg = r->sign, ra = r->a;
                const int req = r->eq;
                const double c = (sg > 0) ? (z - ra) : (ra - z);
                const double lam_r = lams[r_];
                const bool act = req ? true : ((c >= 0.0) || (lam_r > 0.0));
                const double im = act ? mus[r_] : 0.0;
                const double g = im * c + lam_r;
                const double gj = (sg > 0) ? 1.0 : -1.0;
                dd = fma(gj * im, gj, dd);
                aa = fma(gj, g, aa);
            };
            const int4* tab4 = reinterpret_cast<const int4*>(P.col_tab + ct);
            if (j < n) {
                const int4 sl = __ldg(tab4 + j);
                const double z = xs[j];
                own(sl.x, z, dxx, ax); own(sl.y, z, dxx, ax); own(sl.z, z, dxx, ax); own(sl.w, z, dxx, ax);
            }
            if (!term && j < m) {
                const int4 sl = __ldg(tab4 + n + j);
                const double z = xs[n + j];
                own(sl.x, z, duu, au); own(sl.y, z, duu, au); own(sl.z, z, duu, au); own(sl.w, z, duu, au);
            }
        } else {
            auto row_body = [&](const int r_) {
                const DevRow r = P.rows[rb + r_];
                if (r.kind == DR_LIN) {
                    const int col = r.col;
                    const bool ownx = (j < n) && (col == j);
                    const bool ownu = !term && (j < m) && (col == n + j);
                    if (ownx || ownu) {
                        const double z = xs[col];
                        const double c = (r.sign > 0) ? (z - r.a) : (r.a - z);
                        const double lam_r = lams[r_];
                        const bool act = r.eq ? true : ((c >= 0.0) || (lam_r > 0.0));
                        const double im = act ? mus[r_] : 0.0;
                        const double g = im * c + lam_r;
                        const double gj = (r.sign > 0) ? 1.0 : -1.0;
                        if (ownx) { dxx = fma(gj * im, gj, dxx); ax = fma(gj, g, ax); }
                        else { duu = fma(gj * im, gj, duu); au = fma(gj, g, au); }
                    }
                } else {
                    const double gx = (j < n) ? row_jac_s(r, xs, j) : 0.0;
                    const double gu = (!term && j < m) ? row_jac_s(r, xs, n + j) : 0.0;
                    if (gx != 0.0 || gu != 0.0) {
                        const double c = row_value_s(r, xs);
                        const double lam_r = lams[r_];
                        const bool act = r.eq ? true : ((c >= 0.0) || (lam_r > 0.0));
                        const double im = act ? mus[r_] : 0.0;
                        const double g = im * c + lam_r;
                        if (gx != 0.0) {
    #pragma unroll
                            for (int i = 0; i < n; i++) {
                                const double gi = row_jac_s(r, xs, i);
                                if (gi != 0.0) {
                                    if (i == j) dxx = fma(gi * im, gx, dxx);
                                    else Qxxc[i] = fma(gi * im, gx, Qxxc[i]);
                                }
                            }
                            if (!term) {
    #pragma unroll
                                for (int i = 0; i < m; i++) {
                                    const double gi = row_jac_s(r, xs, n + i);
                                    if (gi != 0.0) Quxc[i] = fma(gi * im, gx, Quxc[i]);
                                }
                            }
                            ax = fma(gx, g, ax);
                        }
                        if (gu != 0.0) {
    #pragma unroll
                            for (int i = 0; i < m; i++) {
                                const double gi = row_jac_s(r, xs, n + i);
                                if (gi != 0.0) {
                                    if (i == j) duu = fma(gi * im, gu, duu);
                                    else Quuc[i] = fma(gi * im, gu, Quuc[i]);
                                }
                            }
                            au = fma(gu, g, au);
                        }
                    }
                }
            };
            if (rc > LAMCAP) {
                // Large sets (car_escape: 170 circles per knot, most of them far away): a row that is inactive with a zero
                // multiplier and a finite value contributes im = 0, g = +0, i.e. only +-0 terms to sums that are never -0 -- an exact
                // no-op.  So the lanes first evaluate the rows in PARALLEL (one row per lane and round), and only the rows that do
                // contribute are walked in row order by the sequential body (the same sums, bit for bit).
                const int gshift = __ffs(gmask) - 1;
                for (int base = 0; base < rc; base += GS) {
                    const int rr = base + j;
                    bool contrib = false;
                    if (rr < rc) {
                        const DevRow* r = &P.rows[rb + rr];
                        const int kind = r->kind;
                        if (kind == DR_LIN && term && r->col >= n) {
                            contrib = false;  // a control row at the terminal knot: the body ignores it
                        } else {
                            const double c = row_value_s(*r, xs);
                            const double lam_r = lams[rr];
                            const bool act = r->eq ? true : ((c >= 0.0) || (lam_r > 0.0));
                            contrib = act || !(lam_r == 0.0) || !isfinite(c);
                        }
                    }
                    unsigned bits = (__ballot_sync(gmask, contrib) >> gshift) & ((GS == 32) ? 0xffffffffu : ((1u << GS) - 1u));
                    while (bits) {
                        const int q = __ffs(bits) - 1;
                        bits &= bits - 1u;
                        row_body(base + q);
                    }
                }
            } else {
                for (int r_ = 0; r_ < rc; r_++) row_body(r_);
            }
        }
        if (fast_cost && !term) {
            // diagonal Q, R and H = 0 (flags checked on the host: off-diagonal entries are exactly +0.0): column j of the
            // stage blocks is Q_jj*dt at the diagonal and +0.0 elsewhere -- the same values the general path computes
            const double dt = P.dt;
            if (j < n) {
                const double vqx = (j < nq) ? ((fma(qjj, xs[j], 0.0) + qlin) + 0.0) : 0.0;
                const double qd_ = qjj * dt;
#pragma unroll
                for (int i = 0; i < n; i++) {
                    const double v = (i == j && j < nq) ? qd_ : ((i < nq && j < nq) ? 0.0 * dt : 0.0);
                    const double acc = (i == j) ? dxx : Qxxc[i];
                    Qxxc[i] = has_al ? (v + acc) : v;
                }
                const double v = (j < nq) ? vqx * dt : 0.0;
                Qx_j = has_al ? (v + ax) : v;
#pragma unroll
                for (int i = 0; i < m; i++) {
                    const double w = (i < mq && j < nq) ? (0.0 * dt) : 0.0;
                    Quxc[i] = has_al ? (w + Quxc[i]) : w;
                }
            }
            if (j < m) {
                const double vqu = (j < mq) ? ((fma(rjj, xs[n + j], 0.0) + rlin) + 0.0) : 0.0;
                const double rd_ = rjj * dt;
#pragma unroll
                for (int i = 0; i < m; i++) {
                    const double v = (i == j && j < mq) ? rd_ : ((i < mq && j < mq) ? 0.0 * dt : 0.0);
                    const double acc = (i == j) ? duu : Quuc[i];
                    Quuc[i] = has_al ? (v + acc) : v;
                }
                const double v = (j < mq) ? vqu * dt : 0.0;
                Qu_j = has_al ? (v + au) : v;
            }
            return;
        }
        // the wrapped quadratic cost: own entries of the unscaled gradients
        const double* Qm = term ? P.Qf : P.Q;
        const double* qv = term ? P.qf : P.q;
        const bool qd = term ? P.qf_diag : P.q_diag;
        double vqx = 0.0, vqu = 0.0;
        if (j < nq) {
            double a = 0.0, bq = 0.0;
            if (qd) {
                a = fma(__ldg(&Qm[j * nq + j]), xs[j], a);
            } else {
#pragma unroll
                for (int i = 0; i < nq; i++) a = fma(__ldg(&Qm[i * nq + j]), xs[i], a);
            }
            if (!term && !P.h_zero) {
#pragma unroll
                for (int i = 0; i < mq; i++) bq = fma(__ldg(&P.H[j * mq + i]), xs[n + i], bq);
            }
            vqx = term ? (a + __ldg(&qv[j])) : ((a + __ldg(&qv[j])) + bq);
        }
        if (!term && j < mq) {
            double a = 0.0, bq = 0.0;
            if (P.r_diag) {
                a = fma(__ldg(&P.R[j * mq + j]), xs[n + j], a);
            } else {
#pragma unroll
                for (int i = 0; i < mq; i++) a = fma(__ldg(&P.R[i * mq + j]), xs[n + i], a);
            }
            if (!P.h_zero) {
#pragma unroll
                for (int i = 0; i < nq; i++) bq = fma(__ldg(&P.H[i * mq + j]), xs[i], bq);
            }
            vqu = (a + __ldg(&P.r[j])) + bq;
        }
        double dt = P.dt, tau = 0.0, l1 = 0.0;
        if constexpr (C::MT) {
            if (!term) {
                // the minimum-time blocks couple every lane's gradient entries: exchange them
                if (j < nq) sm.vQx[j] = vqx;
                if (j < mq) sm.vQu[j] = vqu;
                gsync();
                double x[n], u[m];
#pragma unroll
                for (int i = 0; i < n; i++) x[i] = xs[i];
#pragma unroll
                for (int i = 0; i < m; i++) u[i] = xs[n + i];
                tau = u[m - 1];
                dt = tau * tau;
                l1 = quad_stage<C>(P, x, u);
            }
        }
        // cost blocks, column j, added to the AL sums
        if (j < n) {
#pragma unroll
            for (int i = 0; i < n; i++) {
                double v = 0.0;
                if (i < nq && j < nq) v = term ? __ldg(&Qm[j * nq + i]) : __ldg(&Qm[j * nq + i]) * dt;
                if (C::MT && i == n - 1 && j == n - 1) v = P.R_mt;
                const double acc = (i == j) ? dxx : Qxxc[i];
                Qxxc[i] = has_al ? (v + acc) : v;
            }
            double v = 0.0;
            if (j < nq) v = term ? vqx : vqx * dt;
            if (C::MT && j == n - 1) v = P.R_mt * xs[n - 1];
            Qx_j = has_al ? (v + ax) : v;
            if (!term) {
#pragma unroll
                for (int i = 0; i < m; i++) {  // ux is m×n, element (i,j)
                    double w = 0.0;
                    if (i < mq && j < nq) w = P.h_zero ? (0.0 * dt) : __ldg(&P.H[j * mq + i]) * dt;
                    if (C::MT && i == m - 1 && j < nq) w = (2.0 * tau) * vqx;
                    Quxc[i] = has_al ? (w + Quxc[i]) : w;
                }
            }
        }
        if (!term && j < m) {
#pragma unroll
            for (int i = 0; i < m; i++) {
                double v = 0.0;
                if (i < mq && j < mq) v = __ldg(&P.R[j * mq + i]) * dt;
                if constexpr (C::MT) {
                    if (j == m - 1 && i < mq) v = (2.0 * tau) * sm.vQu[i];
                    if (i == m - 1 && j < mq) v = (2.0 * tau) * vqu;
                    if (i == m - 1 && j == m - 1) v = 2.0 * l1 + P.R_mt;
                }
                const double acc = (i == j) ? duu : Quuc[i];
                Quuc[i] = has_al ? (v + acc) : v;
            }
            double v = 0.0;
            if (j < mq) v = vqu * dt;
            if (C::MT && j == m - 1) v = tau * (2.0 * l1 + P.R_mt);
            Qu_j = has_al ? (v + au) : v;
        }
    }

    // Q trajectory (restart quirk Q1): layout per knot [Qx(n) Qu(m) Qxx(n*n) Quu(m*m) Qux(m*n)]
    __device__ void q_store(int k) {
        double* q = ws + L.QST + (size_t)k * C::QS;
        if (j < n) {
            q[j] = Qx_j;
#pragma unroll
            for (int i = 0; i < n; i++) q[n + m + j * n + i] = Qxxc[i];
#pragma unroll
            for (int i = 0; i < m; i++) q[n + m + n * n + m * m + j * m + i] = Quxc[i];
        }
        if (j < m) {
            q[n + j] = Qu_j;
#pragma unroll
            for (int i = 0; i < m; i++) q[n + m + n * n + j * m + i] = Quuc[i];
        }
    }
    // terminal knot (slot N-1 of the Q trajectory; written by ls_expand_kernel for the CTA-per-problem pass): Qx, Qxx only
    __device__ void q_store_term(int k) {
        double* q = ws + L.QST + (size_t)k * C::QS;
        if (j < n) {
            q[j] = Qx_j;
#pragma unroll
            for (int i = 0; i < n; i++) q[n + m + j * n + i] = Qxxc[i];
        }
    }
    __device__ void q_load(int k) {
        const double* q = ws + L.QST + (size_t)k * C::QS;
        if (j < n) {
            Qx_j = q[j];
#pragma unroll
            for (int i = 0; i < n; i++) Qxxc[i] = q[n + m + j * n + i];
#pragma unroll
            for (int i = 0; i < m; i++) Quxc[i] = q[n + m + n * n + m * m + j * m + i];
        }
        if (j < m) {
            Qu_j = q[n + j];
#pragma unroll
            for (int i = 0; i < m; i++) Quuc[i] = q[n + m + n * n + j * m + i];
        }
    }

    // isposdef(Hermitian(A)): the arithmetic of Solver::chol_pd without its early exit (a failed pivot makes the
    // later entries NaN, which nobody reads; only the verdict is used)
    __device__ static bool chol_pd(const double* Areg) {
        double Uc[m * m];
        bool pd = true;
#pragma unroll
        for (int c = 0; c < m; c++) {
#pragma unroll
            for (int i = 0; i < c; i++) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < i; l++) acc = fma(Uc[i * m + l], Uc[c * m + l], acc);
                Uc[c * m + i] = (Areg[c * m + i] - acc) / Uc[i * m + i];
            }
            double acc = 0.0;
#pragma unroll
            for (int l = 0; l < c; l++) acc = fma(Uc[c * m + l], Uc[c * m + l], acc);
            const double dd = Areg[c * m + c] - acc;
            pd = pd && (dd > 0.0);
            Uc[c * m + c] = sqrt(dd);
        }
        return pd;
    }
    struct LU {
        double a[m * m];
        double rp[m];   // 1 / pivot (general path only): correctly rounded reciprocals, reused by lu_solve through div_by
        int piv[m];
        bool tril_only, triu;
    };
    __device__ static void lu_factor(LU& f) {  // same code as Solver::lu_factor (Julia's dense `\`)
        bool tril = true, triu = true;
#pragma unroll
        for (int c = 0; c < m; c++)
#pragma unroll
            for (int i = 0; i < m; i++) {
                if (i < c && f.a[c * m + i] != 0.0) tril = false;
                if (i > c && f.a[c * m + i] != 0.0) triu = false;
            }
        f.triu = triu;
        f.tril_only = tril && !triu;
#pragma unroll
        for (int c = 0; c < m; c++) f.piv[c] = c;
        if (triu || f.tril_only) return;
#pragma unroll
        for (int c = 0; c < m; c++) {
#pragma unroll
            for (int i = 0; i < c; i++) {
#pragma unroll
                for (int q = i + 1; q < m; q++) {
                    const bool sw = (f.piv[i] == q);
                    const double t0 = f.a[c * m + i], t1 = f.a[c * m + q];
                    f.a[c * m + i] = sw ? t1 : t0;
                    f.a[c * m + q] = sw ? t0 : t1;
                }
            }
#pragma unroll
            for (int i = 1; i < c; i++) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < i; l++) acc = fma(f.a[l * m + i], f.a[c * m + l], acc);
                f.a[c * m + i] = f.a[c * m + i] - acc;
            }
#pragma unroll
            for (int i = c; i < m; i++) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < c; l++) acc = fma(f.a[l * m + i], f.a[c * m + l], acc);
                f.a[c * m + i] = f.a[c * m + i] - acc;
            }
            int p = c;
            double amax = fabs(f.a[c * m + c]);
#pragma unroll
            for (int i = c + 1; i < m; i++)
                if (fabs(f.a[c * m + i]) > amax) { amax = fabs(f.a[c * m + i]); p = i; }
            f.piv[c] = p;
#pragma unroll
            for (int q = c + 1; q < m; q++) {
                const bool sw = (p == q);
#pragma unroll
                for (int cc = 0; cc <= c; cc++) {
                    const double t0 = f.a[cc * m + c], t1 = f.a[cc * m + q];
                    f.a[cc * m + c] = sw ? t1 : t0;
                    f.a[cc * m + q] = sw ? t0 : t1;
                }
            }
            const double rp = 1.0 / f.a[c * m + c];
            f.rp[c] = rp;
#pragma unroll
            for (int i = c + 1; i < m; i++) f.a[c * m + i] = f.a[c * m + i] * rp;
        }
    }
    __device__ static void lu_solve(const LU& f, double* bv) {
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
        if (f.triu) {  // triangular input (always so for m = 1): no factorisation ran, plain divisions
#pragma unroll
            for (int i = m - 1; i >= 0; i--) {
                double acc = 0.0;
#pragma unroll
                for (int l = i + 1; l < m; l++) acc = fma(f.a[l * m + i], bv[l], acc);
                bv[i] = (bv[i] - acc) / f.a[i * m + i];
            }
            return;
        }
#pragma unroll
        for (int i = m - 1; i >= 0; i--) {
            double acc = 0.0;
#pragma unroll
            for (int l = i + 1; l < m; l++) acc = fma(f.a[l * m + i], bv[l], acc);
            bv[i] = div_by(bv[i] - acc, f.a[i * m + i], f.rp[i]);
        }
    }

    // the whole backward pass of one problem, including regularisation restarts.  Returns 1 = done, 0 = the regularisation
    // diverged (TO_STATUS_REG_DIVERGED), 2 = handed over: after `max_inline` regularisation increases (>= 0) the pass gives up
    // WITHOUT touching the solver state, and the caller queues the problem for the CTA-per-problem kernel, which replays the
    // pass from the start five times faster per attempt -- a 38-restart chain no longer holds a whole bulk launch for 50 ms.
    __device__ int run(double& dV0, double& dV1, int max_inline = -1) {
        const int N = P.N;
        int nreg = 0;
        bool store_mode = false;
        int stored_from = N - 1;
        for (;;) {
            cp_async_wait_all();
            gsync();
            prefetch(N - 2);
            // terminal cost-to-go: S = Qxx_N, Sx = Qx_N
            if (j < n) sm.xN[j] = ws[L.X + (size_t)(N - 1) * n + j];
            gsync();
            {
                const int lo = P.knot_lam_off[N - 1];
                expansion(N - 1, sm.xN, ws + L.LAM + lo, ws + L.MU + lo);
            }
            if (j < n) {
#pragma unroll
                for (int i = 0; i < n; i++) Scol[i] = Qxxc[i];
                sm.Sx[j] = Qx_j;
            }
            dV0 = 0.0;
            dV1 = 0.0;
            bool failed = false;
            for (int k = N - 2; k >= 0; k--) {
                if (prof) pt0 = clock64();
                cp_async_wait_all();
                gsync();  // knot k's inputs have landed; every lane is done with knot k+1
                const double* AB = sm.AB;
                tick(0);
                // T = A'S (column j) ; Tu = B'S (column j) ; the A'Sx, B'Sx sums
                double accA = 0.0, accB = 0.0;
                if (j < n) {
#pragma unroll
                    for (int l = 0; l < n; l++) accA = fma(AB[l * LDZ + j], sm.Sx[l], accA);
                    double t[n], tu[m];
#pragma unroll
                    for (int i = 0; i < n; i++) t[i] = 0.0;
#pragma unroll
                    for (int i = 0; i < m; i++) tu[i] = 0.0;
#pragma unroll
                    for (int l = 0; l < n; l++) {
                        const double s_l = Scol[l];
#pragma unroll
                        for (int i = 0; i < n; i++) t[i] = fma(AB[l * LDZ + i], s_l, t[i]);
#pragma unroll
                        for (int i = 0; i < m; i++) tu[i] = fma(AB[l * LDZ + n + i], s_l, tu[i]);
                    }
#pragma unroll
                    for (int i = 0; i < n; i++) sm.T[j * LDn + i] = t[i];
#pragma unroll
                    for (int i = 0; i < m; i++) sm.Tu[j * LDm + i] = tu[i];
                }
                if (j < m) {
#pragma unroll
                    for (int l = 0; l < n; l++) accB = fma(AB[l * LDZ + n + j], sm.Sx[l], accB);
                }
                gsync();
                tick(1);
                // T*A, Tu*A, Tu*B (column j), accumulated from zero
                double a1[n], a2[m], a3[m];
#pragma unroll
                for (int i = 0; i < n; i++) a1[i] = 0.0;
#pragma unroll
                for (int i = 0; i < m; i++) { a2[i] = 0.0; a3[i] = 0.0; }
                if (j < n) {
#pragma unroll
                    for (int l = 0; l < n; l++) {
                        const double a_l = AB[l * LDZ + j];
#pragma unroll
                        for (int i = 0; i < n; i++) a1[i] = fma(sm.T[l * LDn + i], a_l, a1[i]);
#pragma unroll
                        for (int i = 0; i < m; i++) a2[i] = fma(sm.Tu[l * LDm + i], a_l, a2[i]);
                    }
                }
                if (j < m) {
#pragma unroll
                    for (int l = 0; l < n; l++) {
                        const double b_l = AB[l * LDZ + n + j];
#pragma unroll
                        for (int i = 0; i < m; i++) a3[i] = fma(sm.Tu[l * LDm + i], b_l, a3[i]);
                    }
                }
                tick(2);
                // the knot's own cost / constraint expansion (independent of S: evaluated only now, so its registers are not
                // live during the two matrix phases), then Q += the S terms
                if (store_mode && k >= stored_from) {
                    q_load(k);
                } else {
                    const int rc = P.knot_row_count[k], lo = P.knot_lam_off[k];
                    const bool staged = (rc <= LAMCAP);
                    expansion(k, sm.xu, staged ? sm.lam : ws + L.LAM + lo, staged ? sm.mu : ws + L.MU + lo);
                }
                if (j < n) {
                    Qx_j += accA;
#pragma unroll
                    for (int i = 0; i < n; i++) Qxxc[i] += a1[i];
#pragma unroll
                    for (int i = 0; i < m; i++) Quxc[i] += a2[i];
                }
                if (j < m) {
                    Qu_j += accB;
#pragma unroll
                    for (int i = 0; i < m; i++) Quuc[i] += a3[i];
                }
                tick(3);
                if (store_mode) {
                    q_store(k);
                    if (k < stored_from) stored_from = k;
                }
                // publish Quu, Qu, Qux'
                if (j < m) {
#pragma unroll
                    for (int i = 0; i < m; i++) sm.Quu[j * m + i] = Quuc[i];
                    sm.Qu[j] = Qu_j;
                }
                if (j < n) {
#pragma unroll
                    for (int l = 0; l < m; l++) sm.QuxT[l * LDn + j] = Quxc[l];
                }
                gsync();
                // every lane is done with this knot's [A B], x, u, multipliers: fetch the next knot's while the
                // factorisation and the cost-to-go update run
                if (k > 0) prefetch(k - 1);
                tick(4);
                // Qxx is not needed until the cost-to-go update: park the column in shared memory (T is free now) so the
                // factorisation below has the registers
                if (j < n) {
#pragma unroll
                    for (int i = 0; i < n; i++) sm.T[j * LDn + i] = Qxxc[i];
                }
                // Quu_reg = Quu + rho*I, replicated per lane
                LU f;
#pragma unroll
                for (int e = 0; e < m * m; e++) f.a[e] = sm.Quu[e];
#pragma unroll
                for (int i = 0; i < m; i++) f.a[i * m + i] = sm.Quu[i * m + i] + rho * 1.0;
                // the positive-definiteness test and the LU factorisation are independent: no branch between them,
                // so their dependent chains (sqrt / divide latencies) overlap
                const bool pd = chol_pd(f.a);
                lu_factor(f);
                if (!pd) { failed = true; break; }
                tick(5);
                if constexpr (n < GS) {
                    // lanes 0..n-1 solve for their column of K, lane n for the feed-forward term d: ONE pass through the
                    // triangular solves (two divergent passes cost a second ~600-cycle dependent chain per knot)
                    if (j <= n) {
                        double rhs[m];
#pragma unroll
                        for (int i = 0; i < m; i++) rhs[i] = (j < n) ? Quxc[i] : sm.Qu[i];
                        lu_solve(f, rhs);
                        if (j < n) {
#pragma unroll
                            for (int i = 0; i < m; i++) {
                                Kcol[i] = -1.0 * rhs[i];
                                sm.KT[i * LDn + j] = Kcol[i];
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < m; i++) sm.d[i] = -1.0 * rhs[i];
                        }
                    }
                } else {
                    if (j < n) {
                        double rhs[m];
#pragma unroll
                        for (int i = 0; i < m; i++) rhs[i] = Quxc[i];
                        lu_solve(f, rhs);
#pragma unroll
                        for (int i = 0; i < m; i++) {
                            Kcol[i] = -1.0 * rhs[i];
                            sm.KT[i * LDn + j] = Kcol[i];
                        }
                    }
                    if (j == DL) {
                        double rhs[m];
#pragma unroll
                        for (int i = 0; i < m; i++) rhs[i] = sm.Qu[i];
                        lu_solve(f, rhs);
#pragma unroll
                        for (int i = 0; i < m; i++) sm.d[i] = -1.0 * rhs[i];
                    }
                }
                // KQ = K'Quu (row j)
                double KQr[m];
                if (j < n) {
#pragma unroll
                    for (int c = 0; c < m; c++) {
                        double acc = 0.0;
#pragma unroll
                        for (int l = 0; l < m; l++) acc = fma(Kcol[l], sm.Quu[c * m + l], acc);
                        KQr[c] = acc;
                        sm.KQ[c * LDn + j] = acc;
                    }
                }
                gsync();
                tick(6);
                double dk[m], Quv[m];
#pragma unroll
                for (int i = 0; i < m; i++) { dk[i] = sm.d[i]; Quv[i] = sm.Qu[i]; }
                {   // publish K, d for the rollouts
                    double* kd = ws + L.KD + (size_t)k * C::KDS;
                    if (j < n) {
#pragma unroll
                        for (int i = 0; i < m; i++) kd[j * m + i] = Kcol[i];
                    }
                    if (j == DL) {
#pragma unroll
                        for (int i = 0; i < m; i++) kd[m * n + i] = dk[i];
                    }
                }
                if (j < n) {
                    // S.x = Qx + KQ d + K'Qu + Qux'd
                    double a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                    for (int l = 0; l < m; l++) a1 = fma(KQr[l], dk[l], a1);
#pragma unroll
                    for (int l = 0; l < m; l++) a2 = fma(Kcol[l], Quv[l], a2);
#pragma unroll
                    for (int l = 0; l < m; l++) a3 = fma(Quxc[l], dk[l], a3);
                    Sx_j = ((Qx_j + a1) + a2) + a3;
                    // unsymmetrised S.xx, column j: Qxx + KQ*K + K'Qux + Qux'K
                    double b1[n], b2[n], b3[n];
#pragma unroll
                    for (int i = 0; i < n; i++) { b1[i] = 0.0; b2[i] = 0.0; b3[i] = 0.0; }
#pragma unroll
                    for (int l = 0; l < m; l++) {
                        const double k_l = Kcol[l], q_l = Quxc[l];
#pragma unroll
                        for (int i = 0; i < n; i++) {
                            b1[i] = fma(sm.KQ[l * LDn + i], k_l, b1[i]);
                            b2[i] = fma(sm.KT[l * LDn + i], q_l, b2[i]);
                            b3[i] = fma(sm.QuxT[l * LDn + i], k_l, b3[i]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < n; i++) {
                        Scol[i] = ((sm.T[j * LDn + i] + b1[i]) + b2[i]) + b3[i];
                        sm.T[j * LDn + i] = Scol[i];
                    }
                    sm.Sx[j] = Sx_j;
                }
                gsync();
                tick(7);
                if (j < n) {
#pragma unroll
                    for (int i = 0; i < n; i++) Scol[i] = 0.5 * (Scol[i] + sm.T[i * LDn + j]);
                }
                // dV (replicated)
                {
                    double a = 0.0;
#pragma unroll
                    for (int l = 0; l < m; l++) a = fma(dk[l], Quv[l], a);
                    dV0 += a;
                    double acc = 0.0;
#pragma unroll
                    for (int c = 0; c < m; c++) {
                        double w = 0.0;
#pragma unroll
                        for (int l = 0; l < m; l++) w = fma(0.5 * dk[l], sm.Quu[c * m + l], w);
                        acc = fma(w, dk[c], acc);
                    }
                    dV1 += acc;
                }
                tick(8);
            }
            if (!failed) break;
            if (!store_mode) {
                store_mode = true;
                stored_from = N - 1;
                continue;
            }
            if (!isfinite(rho)) {
                cp_async_wait_all();
                return 0;
            }
            if (max_inline >= 0 && ++nreg > max_inline) {
                cp_async_wait_all();
                return 2;
            }
            reg_update(true);
        }
        cp_async_wait_all();
        reg_update(false);
        return 1;
    }
};

template <class C, int WARPS>
__device__ __forceinline__ void ls_bp_kernel_body(const DevProblem& Pg, const DevCtl& ctl, const LsCtl& lc, const int cur);

template <class C, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB) ls_bp_kernel(const DevProblem Pg, const DevCtl ctl, const LsCtl lc, const int cur) {
    ls_bp_kernel_body<C, WARPS>(Pg, ctl, lc, cur);
}
// the same kernel under an explicit register cap (experiment variants: ptxas rounds a __launch_bounds__ cap of 136 down to 128)
template <class C, int WARPS, int MAXREG>
__global__ void __maxnreg__(MAXREG) ls_bp_kernel_mr(const DevProblem Pg, const DevCtl ctl, const LsCtl lc, const int cur) {
    ls_bp_kernel_body<C, WARPS>(Pg, ctl, lc, cur);
}

template <class C, int WARPS>
__device__ __forceinline__ void ls_bp_kernel_body(const DevProblem& Pg, const DevCtl& ctl, const LsCtl& lc, const int cur) {
    constexpr int GS = ls_group_size<C>();
    constexpr int GPB = (32 / GS) * WARPS;  // groups per block
    extern __shared__ __align__(16) unsigned char ls_smem_raw[];
    DevProblem P = Pg;
    ls_stage_problem(P, Pg, ls_smem_raw + (size_t)GPB * ls_bp_stride<C>());
    BpSmem<C>& smem_g = *reinterpret_cast<BpSmem<C>*>(ls_smem_raw + (size_t)(threadIdx.x / GS) * ls_bp_stride<C>());
    // cur: bit 0 = active list, bit 1 = hand long restart chains over to the restart list (bulk launches)
    const int max_inline = (cur & 2) ? (cur >> 4) : -1;  // bits 4.. = regularisation increases served inline before the hand-over
    const unsigned int na = lc.counts[cur & 1];
    const int g = threadIdx.x / GS, j = threadIdx.x % GS;
    const int lane = threadIdx.x & 31;
    const unsigned gmask = (GS == 32) ? 0xffffffffu : (((1u << GS) - 1u) << (lane - j));
    const bool al_on = (ctl.mode == 1);
    for (unsigned int a0 = blockIdx.x * GPB; a0 < na; a0 += gridDim.x * GPB) {
        const unsigned int a = a0 + g;
        if (a < na) {
            const int b = lc.list[cur & 1][a];
            LsState* st = &lc.st[b];
            TOiLQROptions io = ctl.o.opts_uncon;
            BpGroup<C> G(P, smem_g, lc.ws + (size_t)b * lc.ws_stride, j, gmask, al_on, io);
            G.rho = st->rho;
            G.drho = st->drho;
            G.load_cost_constants();
            if (ctl.debug && a == 0) G.prof = reinterpret_cast<long long*>(ctl.debug);
            double dV0, dV1;
            const int rc = G.run(dV0, dV1, max_inline);
            if (j == 0) {
                if (rc == 2) {
                    ls_append(lc.restart_list, &lc.counts[5], b);  // state untouched: the CTA kernel starts over from st->rho
                } else {
                    st->rho = G.rho; st->drho = G.drho; st->dV0 = dV0; st->dV1 = dV1;
                    st->winner = -1;
                    st->bp_fail = rc ? 0 : 1;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward pass, latency path: ONE CTA per problem (ls_expand_kernel + ls_bp_cta_kernel)
//
// With few live problems (the tail of a batch: a handful of problems that need > 1,500 iLQR
// iterations) the lane-owns-column pass above is a chain of ~3,000 dependent warp instructions per
// knot.  Here every OUTPUT ELEMENT of the knot's products gets its own thread, so the chain per
// knot is a few 13-long FMA chains plus the (inherently serial) 4x4 factorisation.  The arithmetic
// per element is the same expression as in BpGroup::run (sequential FMA chains from index 0, same
// order of the additions), so the result is bitwise the same; only the assignment to threads differs.
//
// The cost / constraint expansion does not depend on the cost-to-go, so it is knot-parallel:
// ls_expand_kernel evaluates it for every (problem, knot) with the lane-group code of BpGroup and
// writes it to the Q trajectory (slot N-1 = terminal knot).  The recursion then ALWAYS loads Q[k]
// from that trajectory -- which also reproduces the restart quirk Q1 (the reference accumulates
// into Q[k] in place, backward_pass.jl:30-36): a restarted pass finds the accumulated blocks of the
// knots it already visited and the fresh expansion everywhere else.
// ------------------------------------------------------------------------------------------
template <class C, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) ls_expand_kernel(const DevProblem Pg, const DevCtl ctl, const LsCtl lc, const int cur) {
    constexpr int GS = ls_group_size<C>();
    constexpr int GPB = (32 / GS) * WARPS;
    constexpr int n = C::n, m = C::m;
    extern __shared__ __align__(16) unsigned char ls_smem_raw[];
    DevProblem P = Pg;
    ls_stage_problem(P, Pg, ls_smem_raw + (size_t)GPB * ls_bp_stride<C>());
    BpSmem<C>& smem_g = *reinterpret_cast<BpSmem<C>*>(ls_smem_raw + (size_t)(threadIdx.x / GS) * ls_bp_stride<C>());
    const int N = P.N;
    // cur: bit 0 = active list; bit 2 = the restart list of this tick (problems a bulk backward pass handed over)
    const int* list = (cur & 4) ? lc.restart_list : lc.list[cur & 1];
    const unsigned long long total = (unsigned long long)((cur & 4) ? lc.counts[5] : lc.counts[cur & 1]) * (unsigned long long)N;
    const int g = threadIdx.x / GS, j = threadIdx.x % GS;
    const int lane = threadIdx.x & 31;
    const unsigned gmask = (GS == 32) ? 0xffffffffu : (((1u << GS) - 1u) << (lane - j));
    const bool al_on = (ctl.mode == 1);
    for (unsigned long long it0 = (unsigned long long)blockIdx.x * GPB; it0 < total; it0 += (unsigned long long)gridDim.x * GPB) {
        const unsigned long long it = it0 + g;
        if (it < total) {
            const unsigned int a = (unsigned int)(it / N);
            const int k = (int)(it % N);
            const int b = list[a];
            TOiLQROptions io = ctl.o.opts_uncon;
            BpGroup<C> G(P, smem_g, lc.ws + (size_t)b * lc.ws_stride, j, gmask, al_on, io);
            G.load_cost_constants();
            const double* xk = G.ws + G.L.X + (size_t)k * n;
            __syncwarp(gmask);  // the group is done with the previous item's [x;u]
            if (k < N - 1) {
                const double* uk = G.ws + G.L.U + (size_t)k * m;
                for (int e = j; e < n + m; e += GS) smem_g.xu[e] = (e < n) ? xk[e] : uk[e - n];
            } else {
                for (int e = j; e < n; e += GS) smem_g.xu[e] = xk[e];
            }
            __syncwarp(gmask);
            const int lo = P.knot_lam_off[k];
            G.expansion(k, smem_g.xu, G.ws + G.L.LAM + lo, G.ws + G.L.MU + lo);
            if (k < N - 1) G.q_store(k);
            else G.q_store_term(k);
        }
    }
}

template <class C>
struct alignas(16) BpCtaSmem {
    static constexpr int n = C::n, m = C::m;
    static constexpr int LDn = (n + 1) & ~1, LDm = (m + 1) & ~1;
    static constexpr int QSP = (C::QS + 1) & ~1;
    // per-knot inputs, double-buffered (cp.async of knot k-1 runs during knot k)
    double AB[2][C::ZA];   // row l = [A(l,:) B(l,:)], row length C::LDZ
    double Q[2][QSP];      // [Qx | Qu | Qxx (col-major) | Quu | Qux] of the knot (the QST layout)
    double S[n * LDn];     // cost-to-go S(l,j) at [j*LDn+l]
    double Sp[n * LDn];    // unsymmetrised S
    double T[n * LDn];     // T(i,l) = (A'S)(i,l) at [l*LDn+i]
    double Tu[n * LDm];    // Tu(i,l) = (B'S)(i,l) at [l*LDm+i]
    double Qxx[n * LDn];   // (i,j) at [j*LDn+i]
    double Qux[n * LDm];   // (i,j) at [j*LDm+i]
    double K[n * LDm];     // (i,j) at [j*LDm+i]
    double KQ[m * LDn];    // (K'Quu)(i,c) at [c*LDn+i]
    double Quu[m * m], Qx[n], Qu[m], d[m], Sx[n], accA[n], accB[m];
    double QuuC[m * m];    // overlap schedule: the PD-test warp's own copy of Quu
    int pd;
};

// Overlap schedule of the CTA pass (blocks of 256 threads, i.e. the quadrotor): the m x m factorisation chain (four dependent
// reciprocals, ~900 cycles) starts as soon as Quu = Quu_k + (B'S)B exists -- warp 0 computes those 16 elements itself -- and runs
// BESIDE the n x n products of the other warps instead of after them; the PD test does the same on warp 1; the symmetrisation is
// folded into the cost-to-go update (each thread forms S'(i,j) and S'(j,i)).  Three block barriers and one producer/consumer
// named barrier per knot instead of five block barriers with the factorisation alone between two of them.
template <class C, int NT> __host__ __device__ constexpr bool ls_bp_cta_overlap() {
    return NT == 256 && (C::n * C::n + C::m * C::n) <= (NT - 64) + 32 && C::m * C::m <= 16 && C::n < 31 && C::m > 1;
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void bar_sync_n(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// the backward pass of ONE problem by the whole CTA (NT threads; every thread of the block must call it)
template <class C, int NT>
__device__ __forceinline__ void ls_bp_cta_problem(const DevProblem& P, const TOiLQROptions& io, double* ws, const WsLayout& L,
                                                  BpCtaSmem<C>& sm, LsState* st, const int tid, long long* prof = nullptr) {
    constexpr int n = C::n, m = C::m, LDZ = C::LDZ;
    constexpr int LDn = BpCtaSmem<C>::LDn, LDm = BpCtaSmem<C>::LDm;
    typedef typename BpGroup<C>::LU LU;
    const int N = P.N;
    // optional cycle profile (thread 0 of the profiled CTA; diagnostics): cycles per section of the knot loop
    long long pt0 = prof ? clock64() : 0;
    auto ptick = [&](int section) {
        if (prof) {
            const long long t = clock64();
            prof[section] += t - pt0;
            pt0 = t;
        }
    };
    {
        double rho = st->rho, drho = st->drho;  // uniform over the CTA
        auto reg_update = [&](bool increase) {  // ilqr_methods.jl:164-176
            const double f = io.bp_reg_increase_factor;
            if (increase) {
                drho = dmax(drho * f, f);
                rho = dmax(rho * drho, io.bp_reg_min);
            } else {
                drho = dmin(drho / f, 1.0 / f);
                rho = rho * drho * ((rho * drho > io.bp_reg_min) ? 1.0 : 0.0);
            }
        };
        auto prefetch = [&](int k, int buf) {
            const double* ab = ws + L.Z + (size_t)k * C::ZA;  // 16-byte aligned, C::ZA even
            for (int e = 2 * tid; e < C::ZA; e += 2 * NT) cp_async16(&sm.AB[buf][e], ab + e);
            const double* q = ws + L.QST + (size_t)k * C::QS;
            for (int e = tid; e < C::QS; e += NT) cp_async8(&sm.Q[buf][e], q + e);
        };
        // bp_reg_type = :state (backward_pass.jl:38-46): Quu_reg = Quu + (rho B')B, Qux_reg = Qux + (rho B')A, each thread forms
        // what it needs from the knot's [A B] in shared memory (the same sums as the oracle: (rho*B(l,i)) first, l ascending)
        const bool reg_state = (io.bp_reg_type == TO_REG_STATE);
        auto quu_reg = [&](double* a, const double* AB, const double* Quu) {
            if (reg_state) {
#pragma unroll
                for (int e = 0; e < m * m; e++) {
                    const int i = e % m, j = e / m;
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < n; l++) acc = fma(rho * AB[l * LDZ + n + i], AB[l * LDZ + n + j], acc);
                    a[e] = Quu[e] + acc;
                }
            } else {
#pragma unroll
                for (int e = 0; e < m * m; e++) a[e] = Quu[e];
#pragma unroll
                for (int i = 0; i < m; i++) a[i * m + i] = Quu[i * m + i] + rho * 1.0;
            }
        };
        auto qux_reg_col = [&](double* rhs, const double* AB, const double* Quxcol, int col) {
#pragma unroll
            for (int i = 0; i < m; i++) {
                double v = Quxcol[i];
                if (reg_state) {
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < n; l++) acc = fma(rho * AB[l * LDZ + n + i], AB[l * LDZ + col], acc);
                    v = Quxcol[i] + acc;
                }
                rhs[i] = v;
            }
        };
        bool store_mode = false, ok = true;
        double dV0 = 0.0, dV1 = 0.0;  // kept by thread NT-1
        for (;;) {
            cp_async_wait_all();  // no copy of an abandoned attempt may land after the ones issued below
            __syncthreads();      // everybody is done with the previous attempt / problem
            prefetch(N - 2, (N - 2) & 1);
            {   // terminal cost-to-go: S = Qxx_N, Sx = Qx_N
                const double* qt = ws + L.QST + (size_t)(N - 1) * C::QS;
                for (int e = tid; e < n * n; e += NT) sm.S[(e / n) * LDn + (e % n)] = qt[n + m + e];
                for (int e = tid; e < n; e += NT) sm.Sx[e] = qt[e];
            }
            dV0 = 0.0;
            dV1 = 0.0;
            bool failed = false;
            for (int k = N - 2; k >= 0; k--) {
                const int buf = k & 1;
                ptick(0);
                cp_async_wait_all();
                __syncthreads();  // knot k's inputs have landed; S, Sx of knot k+1 are complete
                ptick(1);
                if (k > 0) prefetch(k - 1, buf ^ 1);
                const double* AB = sm.AB[buf];
                const double* Qk = sm.Q[buf];
                // ---- step 1: T = A'S, Tu = B'S, A'Sx, B'Sx (one output element per thread).  Every task is "column c of [A B] dot a
                // vector": one code path, selected by pointers, so the warps that hold two task kinds do not run two chains in turn
                for (int t = tid; t < n * n + m * n + n + m; t += NT) {
                    int c;
                    const double* V;
                    double* out;
                    if (t < n * n) {
                        const int i = t % n, j = t / n;
                        c = i; V = &sm.S[j * LDn]; out = &sm.T[j * LDn + i];
                    } else if (t < n * n + m * n) {
                        const int e = t - n * n, i = e % m, j = e / m;
                        c = n + i; V = &sm.S[j * LDn]; out = &sm.Tu[j * LDm + i];
                    } else if (t < n * n + m * n + n) {
                        const int j = t - n * n - m * n;
                        c = j; V = sm.Sx; out = &sm.accA[j];
                    } else {
                        const int j = t - n * n - m * n - n;
                        c = n + j; V = sm.Sx; out = &sm.accB[j];
                    }
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < n; l++) acc = fma(AB[l * LDZ + c], V[l], acc);
                    *out = acc;
                }
                ptick(2);
                __syncthreads();
                ptick(3);
                double* qg = ws + L.QST + (size_t)k * C::QS;
                if constexpr (ls_bp_cta_overlap<C, NT>()) {
                    const int warp = tid >> 5, lane = tid & 31;
                    // ---- step 2 (overlap schedule): one "row of T / Tu dot column of [A B]" task per thread
                    {
                        constexpr int NBIG = m * n + n * n;  // Qux first (the solves wait for it), then Qxx
                        int g = -1, e = -1;
                        if (warp >= 2) g = tid - 64;
                        else if (lane >= 16) g = (NT - 64) + warp * 16 + (lane - 16);
                        else e = lane;  // warps 0 and 1, lanes 0..15: Quu (both warps: each feeds its own factorisation)
                        const double* Lp = nullptr;
                        double* out = nullptr;
                        int ld = 0, c = 0, qi = 0;
                        bool glob = store_mode;
                        if (g >= 0 && g < m * n) {
                            const int i = g % m, j = g / m;
                            Lp = &sm.Tu[i]; ld = LDm; c = j; qi = n + m + n * n + m * m + g; out = &sm.Qux[j * LDm + i];
                        } else if (g >= m * n && g < NBIG) {
                            const int t = g - m * n, i = t % n, j = t / n;
                            Lp = &sm.T[i]; ld = LDn; c = j; qi = n + m + t; out = &sm.Qxx[j * LDn + i];
                        } else if (e >= 0 && e < m * m) {
                            const int i = e % m, j = e / m;
                            Lp = &sm.Tu[i]; ld = LDm; c = n + j; qi = n + m + n * n + e;
                            out = (warp == 0) ? &sm.Quu[j * m + i] : &sm.QuuC[j * m + i];
                            glob = store_mode && warp == 0;
                        }
                        if (out) {
                            double acc = 0.0;
#pragma unroll
                            for (int l = 0; l < n; l++) acc = fma(Lp[l * ld], AB[l * LDZ + c], acc);
                            const double v = Qk[qi] + acc;
                            *out = v;
                            if (glob) qg[qi] = v;
                        }
                    }
                    ptick(4);
                    if (warp >= 2) {
                        bar_arrive(2, NT - 32);  // Qux (and Qxx) of this warp are in shared memory
                    } else if (warp == 0) {
                        if (lane < n) {
                            const double v = Qk[lane] + sm.accA[lane];
                            sm.Qx[lane] = v;
                            if (store_mode) qg[lane] = v;
                        }
                        if (lane < m) {
                            const double v = Qk[n + lane] + sm.accB[lane];
                            sm.Qu[lane] = v;
                            if (store_mode) qg[n + lane] = v;
                        }
                        __syncwarp();
                        LU f;
                        if (lane <= n) {
                            quu_reg(f.a, AB, sm.Quu);
                            BpGroup<C>::lu_factor(f);
                        }
                        ptick(5);
                        bar_sync_n(2, NT - 32);  // Qux of the other warps has arrived
                        ptick(6);
                        if (lane <= n) {
                            double rhs[m];
                            if (lane < n) qux_reg_col(rhs, AB, &sm.Qux[lane * LDm], lane);
                            else {
#pragma unroll
                                for (int i = 0; i < m; i++) rhs[i] = sm.Qu[i];
                            }
                            BpGroup<C>::lu_solve(f, rhs);
                            if (lane < n) {
                                double Kc[m];
#pragma unroll
                                for (int i = 0; i < m; i++) {
                                    Kc[i] = -1.0 * rhs[i];
                                    sm.K[lane * LDm + i] = Kc[i];
                                }
#pragma unroll
                                for (int c = 0; c < m; c++) {  // KQ = K'Quu (row `lane`)
                                    double acc = 0.0;
#pragma unroll
                                    for (int l = 0; l < m; l++) acc = fma(Kc[l], sm.Quu[c * m + l], acc);
                                    sm.KQ[c * LDn + lane] = acc;
                                }
                            } else {
#pragma unroll
                                for (int i = 0; i < m; i++) sm.d[i] = -1.0 * rhs[i];
                            }
                        }
                    } else {  // warp 1: isposdef(Hermitian(Quu_reg))
                        __syncwarp();
                        if (lane == 0) {
                            double A_[m * m];
                            quu_reg(A_, AB, sm.QuuC);
                            sm.pd = BpGroup<C>::chol_pd(A_) ? 1 : 0;
                        }
                    }
                    ptick(7);
                    __syncthreads();
                    ptick(8);
                    if (!sm.pd) { failed = true; break; }
                    // ---- steps 4 + 5: gains out, symmetrised cost-to-go, dV
                    {
                        double* kd = ws + L.KD + (size_t)k * C::KDS;
                        for (int t = tid; t < n * n + n + m * n + m + 1; t += NT) {
                            if (t < n * n) {  // S.xx(i,j) = 0.5 (S'(i,j) + S'(j,i)),  S' = Qxx + KQ*K + K'Qux + Qux'K
                                const int i = t % n, j = t / n;
                                double b1 = 0.0, b2 = 0.0, b3 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
#pragma unroll
                                for (int l = 0; l < m; l++) {
                                    const double kj = sm.K[j * LDm + l], qj = sm.Qux[j * LDm + l];
                                    const double ki = sm.K[i * LDm + l], qi_ = sm.Qux[i * LDm + l];
                                    b1 = fma(sm.KQ[l * LDn + i], kj, b1);
                                    b2 = fma(ki, qj, b2);
                                    b3 = fma(qi_, kj, b3);
                                    c1 = fma(sm.KQ[l * LDn + j], ki, c1);
                                    c2 = fma(kj, qi_, c2);
                                    c3 = fma(qj, ki, c3);
                                }
                                const double sij = ((sm.Qxx[j * LDn + i] + b1) + b2) + b3;
                                const double sji = ((sm.Qxx[i * LDn + j] + c1) + c2) + c3;
                                sm.S[j * LDn + i] = 0.5 * (sij + sji);
                            } else if (t < n * n + n) {  // S.x = Qx + KQ d + K'Qu + Qux'd
                                const int j = t - n * n;
                                double a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                                for (int l = 0; l < m; l++) a1 = fma(sm.KQ[l * LDn + j], sm.d[l], a1);
#pragma unroll
                                for (int l = 0; l < m; l++) a2 = fma(sm.K[j * LDm + l], sm.Qu[l], a2);
#pragma unroll
                                for (int l = 0; l < m; l++) a3 = fma(sm.Qux[j * LDm + l], sm.d[l], a3);
                                sm.Sx[j] = ((sm.Qx[j] + a1) + a2) + a3;
                            } else if (t < n * n + n + m * n) {
                                const int e = t - n * n - n;  // kd[j*m+i] = K(i,j)
                                kd[e] = sm.K[(e / m) * LDm + (e % m)];
                            } else if (t < n * n + n + m * n + m) {
                                const int i = t - n * n - n - m * n;
                                kd[m * n + i] = sm.d[i];
                            }
                        }
                        if (tid == NT - 1) {  // dV
                            double a_ = 0.0;
#pragma unroll
                            for (int l = 0; l < m; l++) a_ = fma(sm.d[l], sm.Qu[l], a_);
                            dV0 += a_;
                            double acc = 0.0;
#pragma unroll
                            for (int c = 0; c < m; c++) {
                                double w = 0.0;
#pragma unroll
                                for (int l = 0; l < m; l++) w = fma(0.5 * sm.d[l], sm.Quu[c * m + l], w);
                                acc = fma(w, sm.d[c], acc);
                            }
                            dV1 += acc;
                        }
                    }
                    continue;  // the barrier at the top of the next knot separates this update from its readers
                }
                // ---- step 2: Q = Q[k] + (T*A, Tu*A, Tu*B, A'Sx, B'Sx); restart mode writes the sums back (quirk Q1)
                for (int t = tid; t < n * n + m * n + m * m + n + m; t += NT) {
                    if (t < n * n + m * n + m * m) {  // "row i of T or Tu dot column c of [A B]": one code path
                        const double* Lp;
                        int ld, c, qi;
                        double* out;
                        if (t < n * n) {
                            const int i = t % n, j = t / n;
                            Lp = &sm.T[i]; ld = LDn; c = j; qi = n + m + t; out = &sm.Qxx[j * LDn + i];
                        } else if (t < n * n + m * n) {
                            const int e = t - n * n, i = e % m, j = e / m;
                            Lp = &sm.Tu[i]; ld = LDm; c = j; qi = n + m + n * n + m * m + e; out = &sm.Qux[j * LDm + i];
                        } else {
                            const int e = t - n * n - m * n, i = e % m, j = e / m;
                            Lp = &sm.Tu[i]; ld = LDm; c = n + j; qi = n + m + n * n + e; out = &sm.Quu[j * m + i];
                        }
                        double acc = 0.0;
#pragma unroll
                        for (int l = 0; l < n; l++) acc = fma(Lp[l * ld], AB[l * LDZ + c], acc);
                        const double v = Qk[qi] + acc;
                        *out = v;
                        if (store_mode) qg[qi] = v;
                    } else if (t < n * n + m * n + m * m + n) {
                        const int j = t - n * n - m * n - m * m;
                        const double v = Qk[j] + sm.accA[j];
                        sm.Qx[j] = v;
                        if (store_mode) qg[j] = v;
                    } else {
                        const int j = t - n * n - m * n - m * m - n;
                        const double v = Qk[n + j] + sm.accB[j];
                        sm.Qu[j] = v;
                        if (store_mode) qg[n + j] = v;
                    }
                }
                __syncthreads();
                // ---- step 3: Quu_reg = Quu + rho*I: PD test (warp 1) beside the LU factorisation and the solves for K, d (warp 0)
                if (tid <= n) {
                    LU f;
                    quu_reg(f.a, AB, sm.Quu);
                    BpGroup<C>::lu_factor(f);
                    double rhs[m];
                    if (tid < n) qux_reg_col(rhs, AB, &sm.Qux[tid * LDm], tid);
                    else {
#pragma unroll
                        for (int i = 0; i < m; i++) rhs[i] = sm.Qu[i];
                    }
                    BpGroup<C>::lu_solve(f, rhs);
                    if (tid < n) {
                        double Kc[m];
#pragma unroll
                        for (int i = 0; i < m; i++) {
                            Kc[i] = -1.0 * rhs[i];
                            sm.K[tid * LDm + i] = Kc[i];
                        }
#pragma unroll
                        for (int c = 0; c < m; c++) {  // KQ = K'Quu (row tid)
                            double acc = 0.0;
#pragma unroll
                            for (int l = 0; l < m; l++) acc = fma(Kc[l], sm.Quu[c * m + l], acc);
                            sm.KQ[c * LDn + tid] = acc;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < m; i++) sm.d[i] = -1.0 * rhs[i];
                    }
                } else if (tid == 32) {
                    double A_[m * m];
                    quu_reg(A_, AB, sm.Quu);
                    sm.pd = BpGroup<C>::chol_pd(A_) ? 1 : 0;
                }
                __syncthreads();
                if (!sm.pd) { failed = true; break; }
                // ---- step 4: gains out, unsymmetrised cost-to-go, dV
                {
                    double* kd = ws + L.KD + (size_t)k * C::KDS;
                    for (int t = tid; t < n * n + n + m * n + m + 1; t += NT) {
                        if (t < n * n) {  // S.xx(i,j) = Qxx + KQ*K + K'Qux + Qux'K
                            const int i = t % n, j = t / n;
                            double b1 = 0.0, b2 = 0.0, b3 = 0.0;
#pragma unroll
                            for (int l = 0; l < m; l++) {
                                const double k_l = sm.K[j * LDm + l], q_l = sm.Qux[j * LDm + l];
                                b1 = fma(sm.KQ[l * LDn + i], k_l, b1);
                                b2 = fma(sm.K[i * LDm + l], q_l, b2);
                                b3 = fma(sm.Qux[i * LDm + l], k_l, b3);
                            }
                            sm.Sp[j * LDn + i] = ((sm.Qxx[j * LDn + i] + b1) + b2) + b3;
                        } else if (t < n * n + n) {  // S.x = Qx + KQ d + K'Qu + Qux'd
                            const int j = t - n * n;
                            double a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                            for (int l = 0; l < m; l++) a1 = fma(sm.KQ[l * LDn + j], sm.d[l], a1);
#pragma unroll
                            for (int l = 0; l < m; l++) a2 = fma(sm.K[j * LDm + l], sm.Qu[l], a2);
#pragma unroll
                            for (int l = 0; l < m; l++) a3 = fma(sm.Qux[j * LDm + l], sm.d[l], a3);
                            sm.Sx[j] = ((sm.Qx[j] + a1) + a2) + a3;
                        } else if (t < n * n + n + m * n) {
                            const int e = t - n * n - n;  // kd[j*m+i] = K(i,j)
                            kd[e] = sm.K[(e / m) * LDm + (e % m)];
                        } else if (t < n * n + n + m * n + m) {
                            const int i = t - n * n - n - m * n;
                            kd[m * n + i] = sm.d[i];
                        }
                    }
                    if (tid == NT - 1) {  // dV
                        double a_ = 0.0;
#pragma unroll
                        for (int l = 0; l < m; l++) a_ = fma(sm.d[l], sm.Qu[l], a_);
                        dV0 += a_;
                        double acc = 0.0;
#pragma unroll
                        for (int c = 0; c < m; c++) {
                            double w = 0.0;
#pragma unroll
                            for (int l = 0; l < m; l++) w = fma(0.5 * sm.d[l], sm.Quu[c * m + l], w);
                            acc = fma(w, sm.d[c], acc);
                        }
                        dV1 += acc;
                    }
                }
                __syncthreads();
                // ---- step 5: symmetrise
                for (int t = tid; t < n * n; t += NT) {
                    const int i = t % n, j = t / n;
                    sm.S[j * LDn + i] = 0.5 * (sm.Sp[j * LDn + i] + sm.Sp[i * LDn + j]);
                }
            }
            if (!failed) break;
            if (!store_mode) {  // first failure: replay the pass keeping the accumulated blocks
                store_mode = true;
                continue;
            }
            if (!isfinite(rho)) { ok = false; break; }
            reg_update(true);
        }
        cp_async_wait_all();
        if (ok) reg_update(false);
        if (tid == NT - 1) {
            st->rho = rho; st->drho = drho; st->dV0 = dV0; st->dV1 = dV1;
            st->winner = -1;
            st->bp_fail = ok ? 0 : 1;
        }
    }
}

template <class C, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) ls_bp_cta_kernel(const DevProblem P, const DevCtl ctl, const LsCtl lc, const int cur) {
    extern __shared__ __align__(16) unsigned char ls_smem_raw[];
    BpCtaSmem<C>& sm = *reinterpret_cast<BpCtaSmem<C>*>(ls_smem_raw);
    const WsLayout L = ws_layout<C>(P.N, P.Ptot, false);
    const int tid = threadIdx.x;
    const int* list = (cur & 4) ? lc.restart_list : lc.list[cur & 1];
    const unsigned int na = (cur & 4) ? lc.counts[5] : lc.counts[cur & 1];
    const TOiLQROptions io = ctl.o.opts_uncon;
    for (unsigned int a = blockIdx.x; a < na; a += gridDim.x) {
        const int b = list[a];
        ls_bp_cta_problem<C, NT>(P, io, lc.ws + (size_t)b * lc.ws_stride, L, sm, &lc.st[b], tid);
    }
}

// ------------------------------------------------------------------------------------------
// line search: thread per (problem, step size); G consecutive lanes serve one problem
// ------------------------------------------------------------------------------------------
template <class C>
struct RolloutStage {
    static constexpr int SS = C::n + C::m + C::KDS;  // x_k, u_k, K_k, d_k
    static constexpr int LC = 16;                    // multipliers / penalties of one knot (larger sets are read from global)
    static constexpr int SBUF = (SS + 2 * LC + 1) & ~1;
};

// knot_al_cost with the knot's own multiplier / penalty arrays (same arithmetic as knot_al_cost in engine.cuh)
template <class C>
TOB_DEV double knot_al_cost_at(const DevProblem& P, int k, const double* lamk, const double* muk, const double* x, const double* u) {
    const int rb = P.knot_row_begin[k], rc = P.knot_row_count[k];
    if (rc == 0) return 0.0 + 0.0;
    // z = [x;u] as a dynamically indexable (local-memory, L1-resident) array: a bound / goal / slack row reads z[col] with one
    // load instead of walking an (n+m)-way select chain over registers (that chain was 20 % of the instructions of the
    // quadrotor line search: 4 rows x 17 compares + selects per knot)
    double z[C::n + C::m];
#pragma unroll
    for (int i = 0; i < C::n; i++) z[i] = x[i];
#pragma unroll
    for (int i = 0; i < C::m; i++) z[C::n + i] = u[i];
    double t1 = 0.0, t2 = 0.0;
    int i = 0;
    if (rc > RolloutStage<C>::LC) {
        // large constraint sets (car_escape: 177 rows per knot; multipliers in global memory): four rows at a time with every
        // load issued before the first use -- the one-row-at-a-time loop paid a dependent load + branch chain of ~340 cycles
        // per row (profiles/r01e3).  Same per-row expressions, same accumulation order.
        // the multipliers / penalties of the NEXT four rows are requested before the current four are used (they come from L2;
        // the rows themselves sit in shared memory): one warp per SMSP has nothing else to hide that latency with
        double ln[4], mn[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            ln[q] = (q < rc) ? __ldg(lamk + q) : 0.0;
            mn[q] = (q < rc) ? __ldg(muk + q) : 0.0;
        }
        for (; i + 4 <= rc; i += 4) {
            double l4[4], m4[4], c4[4], ra[4], rb_[4], rr[4];
            int eq4[4], kind4[4];
#pragma unroll
            for (int q = 0; q < 4; q++) { l4[q] = ln[q]; m4[q] = mn[q]; }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int nx = i + 4 + q;
                ln[q] = (nx < rc) ? __ldg(lamk + nx) : 0.0;
                mn[q] = (nx < rc) ? __ldg(muk + nx) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {  // every load of the four rows first
                const DevRow* r = &P.rows[rb + i + q];
                kind4[q] = r->kind; eq4[q] = r->eq;
                ra[q] = r->a; rb_[q] = r->b; rr[q] = r->r;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {  // circle rows (the bulk of such sets) without a branch: four independent chains
                const double dx = x[0] - ra[q], dy = x[1] - rb_[q];
                c4[q] = -(((dx * dx) + (dy * dy)) - (rr[q] * rr[q]));
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (kind4[q] != DR_CIRCLE) c4[q] = BpGroup<C>::row_value_s(P.rows[rb + i + q], z);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const bool act = eq4[q] ? true : ((c4[q] >= 0.0) || (l4[q] > 0.0));
                const double am = act ? m4[q] : 0.0;
                t1 = fma(l4[q], c4[q], t1);
                t2 = fma((0.5 * c4[q]) * am, c4[q], t2);
            }
        }
    }
    for (; i < rc; i++) {
        const DevRow r = P.rows[rb + i];
        const double c = BpGroup<C>::row_value_s(r, z);
        const double l = lamk[i];
        const bool act = r.eq ? true : ((c >= 0.0) || (l > 0.0));
        const double am = act ? muk[i] : 0.0;
        t1 = fma(l, c, t1);
        t2 = fma((0.5 * c) * am, c, t2);
    }
    return t1 + t2;
}

// Candidate trajectories of the line search: CW step sizes of one problem share a buffer, interleaved element by element
// (element e of candidate `col` at e*CW + col), so the lanes of a problem write each element as one contiguous segment.
// (Measured alternative, profiles/r01e4: interleaving in 32-byte units makes the accept kernel's single-candidate read use whole
// sectors -- accept 0.68 -> 0.65 ms per 16,384 problems -- but turns every rollout store into partial-sector writes: line search
// 3.17 -> 3.54 ms.  Kept element-wise.)
__host__ __device__ __forceinline__ size_t cand_index(size_t e, int col, int CW) { return e * (size_t)CW + (size_t)col; }
__host__ __device__ __forceinline__ size_t cand_span(size_t count, int CW) { return count * (size_t)CW; }  // doubles for `count` elements
// Bulk candidates (CW = LS_TRIAL_G step sizes per problem): [knot][chunk of 4 components][step size][4 doubles].  A thread of the
// line search writes whole 32-byte sectors (its 4-component chunk), the 8 step sizes of a problem side by side (256 contiguous
// bytes), and the accept kernel reads exactly the winner's sectors.  The element-interleaved layout ([element][step size], kept
// for the 32-wide tail / resident buffers, which live in L2) made that read fetch 4x what it used: one 8-byte column out of every
// 32-byte sector, 1.07 GB for 0.27 GB per tick of 8,192 problems.
__host__ __device__ __forceinline__ int cand_chunks(int comps) { return (comps + 3) >> 2; }
__host__ __device__ __forceinline__ size_t cand_chunk_span(size_t knots, int comps, int CW) { return knots * (size_t)cand_chunks(comps) * (size_t)CW * 4; }
__host__ __device__ __forceinline__ size_t cand_chunk_index(size_t k, int chunk, int comps, int slot, int CW) {
    return ((k * (size_t)cand_chunks(comps) + (size_t)chunk) * (size_t)CW + (size_t)slot) * 4;
}
// store comps doubles of knot k as 4-component chunks (zero padded)
template <int COMPS>
__device__ __forceinline__ void cand_store_chunks(double* base, size_t k, int slot, int CW, const double* v) {
#pragma unroll
    for (int c = 0; c < (COMPS + 3) / 4; c++) {
        double2* dst = reinterpret_cast<double2*>(base + cand_chunk_index(k, c, COMPS, slot, CW));
        dst[0] = make_double2(v[4 * c], (4 * c + 1 < COMPS) ? v[4 * c + 1] : 0.0);
        dst[1] = make_double2((4 * c + 2 < COMPS) ? v[4 * c + 2] : 0.0, (4 * c + 3 < COMPS) ? v[4 * c + 3] : 0.0);
    }
}


template <class C>
struct Rollout {
    // Line-search rollout with the per-knot inputs (x_k, u_k, K_k, d_k, lambda_k, mu_k) staged in shared memory one knot
    // ahead by cp.async: the `nact` lanes that serve one problem (lane index t within them, warp mask amask) share one
    // double buffer `stg`.  Arithmetic identical to run<false, CAND, CW>.
    template <bool CAND, int CW, bool COST = true>
    static __device__ bool run_staged(const DevProblem& P, const TOiLQROptions& io, double* ws, const WsLayout& L, const double* x0,
                                      double alpha, bool al_on, double& Jt, double* XB, double* UB, int slot, double* stg, int t,
                                      int nact, unsigned amask) {
        constexpr int n = C::n, m = C::m;
        constexpr int SS = RolloutStage<C>::SS, LC = RolloutStage<C>::LC, SBUF = RolloutStage<C>::SBUF;
        const int N = P.N;
        const double* lam = ws + L.LAM;
        const double* mu = ws + L.MU;
        auto prefetch = [&](int k, int buf) {
            double* dst = stg + buf * SBUF;
            const double* xk = ws + L.X + (size_t)k * n;
            const double* uk = ws + L.U + (size_t)k * m;
            const double* kd = ws + L.KD + (size_t)k * C::KDS;
            for (int e = t; e < SS; e += nact) cp_async8(dst + e, (e < n) ? (xk + e) : ((e < n + m) ? (uk + (e - n)) : (kd + (e - n - m))));
            if (COST && al_on) {
                const int rc = P.knot_row_count[k], lo = P.knot_lam_off[k];
                if (rc <= LC) {
                    for (int e = t; e < rc; e += nact) {
                        cp_async8(dst + SS + e, lam + lo + e);
                        cp_async8(dst + SS + LC + e, mu + lo + e);
                    }
                } else {  // too many for the staging buffer: the row loop reads global memory, one knot ahead into L1
                    prefetch_span(lam + lo, rc, t, nact);
                    prefetch_span(mu + lo, rc, t, nact);
                }
            }
        };
        double xb[n], ub[m];
#pragma unroll
        for (int i = 0; i < n; i++) xb[i] = x0[i];
        bool ok = true;
        double J = 0.0, Jc = 0.0;
        prefetch(0, 0);
        for (int k = 0; k < N - 1; k++) {
            cp_async_wait_all();
            __syncwarp(amask);  // knot k has landed; every lane is done with the other buffer
            if (k + 1 < N - 1) prefetch(k + 1, (k + 1) & 1);
            // A rollout that left the state / control box (rollout.jl:14-18 returns false there) is rejected whatever comes
            // after: its lane only keeps serving the staging copies and barriers.  Without this the diverged lanes (about one
            // rollout in five at the large step sizes) drag their warp through the NaN slow paths of every division and
            // square root of the remaining knots (21 % of the kernel's samples, profiles/r01y source view).
            if (!ok) continue;
            const double* sk = stg + (k & 1) * SBUF;
            const double* Xk = sk;
            const double* Uk = sk + n;
            const double* Kk = sk + n + m;
            const double* dk = Kk + m * n;
            double dx[n];
#pragma unroll
            for (int i = 0; i < n; i++) dx[i] = xb[i] - Xk[i];
#pragma unroll
            for (int i = 0; i < m; i++) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < n; c++) acc = fma(Kk[c * m + i], dx[c], acc);
                ub[i] = (Uk[i] + acc) + alpha * dk[i];
            }
            if (COST) J += stage_cost<C>(P, xb, ub);
            if (COST && al_on) {
                const int rc = P.knot_row_count[k];
                const bool staged = (rc <= LC);
                const int lo = P.knot_lam_off[k];
                Jc += knot_al_cost_at<C>(P, k, staged ? (sk + SS) : (lam + lo), staged ? (sk + SS + LC) : (mu + lo), xb, ub);
            }
            if (CAND) {
                if constexpr (CW == 32) {
#pragma unroll
                    for (int i = 0; i < n; i++) XB[cand_index((size_t)k * n + i, slot, CW)] = xb[i];
#pragma unroll
                    for (int i = 0; i < m; i++) UB[cand_index((size_t)k * m + i, slot, CW)] = ub[i];
                } else {
                    cand_store_chunks<n>(XB, (size_t)k, slot, CW, xb);
                    cand_store_chunks<m>(UB, (size_t)k, slot, CW, ub);
                }
            }
            double xn[n];
            {
                double dt = P.dt;
                if constexpr (C::MT) {
                    const double h = ub[m - 1];
                    dt = h * h;
                }
                fd_model<C::MODEL, C::INTEG, double>(xn, xb, ub, dt);
                if constexpr (C::INF) {
#pragma unroll
                    for (int i = 0; i < C::n0; i++) xn[i] = xn[i] + ub[C::m0 + i];
                }
                if constexpr (C::MT) xn[n - 1] = ub[m - 1];
            }
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
        if (ok) {
            double uz[m];
#pragma unroll
            for (int i = 0; i < m; i++) uz[i] = 0.0;
            if (COST) J += term_cost<C>(P, xb);
            if (COST && al_on) Jc += knot_al_cost<C>(P, N - 1, lam, mu, xb, uz);
            if (CAND) {
                if constexpr (CW == 32) {
#pragma unroll
                    for (int i = 0; i < n; i++) XB[cand_index((size_t)(N - 1) * n + i, slot, CW)] = xb[i];
                } else {
                    cand_store_chunks<n>(XB, (size_t)(N - 1), slot, CW, xb);
                }
            }
        }
        __syncwarp(amask);  // nobody reads the staging buffers any more
        Jt = al_on ? (J + Jc) : J;
        return ok;
    }

    // one closed-loop rollout with step size alpha (rollout.jl:2-23) + its cost (objective.jl:40-48 + AL)
    // WRITE: also store X̄, Ū in place (X <- X̄, U <- Ū) and accumulate the Todorov gradient
    // CAND: also store the candidate trajectory X̄, Ū in the 32-way interleaved buffers XB / UB (column `slot`), so that the
    // tail-mode accept kernel copies the winner instead of re-rolling it
    template <bool WRITE, bool CAND = false, int CW = 32>
    static __device__ bool run(const DevProblem& P, const TOiLQROptions& io, double* ws, const WsLayout& L, const double* x0,
                               double alpha, bool al_on, double& Jt, double& grad_sum, double* XB = nullptr, double* UB = nullptr,
                               int slot = 0) {
        constexpr int n = C::n, m = C::m;
        const int N = P.N;
        double xb[n], ub[m];
#pragma unroll
        for (int i = 0; i < n; i++) xb[i] = x0[i];
        bool ok = true;
        double J = 0.0, Jc = 0.0, gs = 0.0;
        const double* lam = ws + L.LAM;
        const double* mu = ws + L.MU;
        const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
        for (int k = 0; k < N - 1; k++) {
            double* Xk = ws + L.X + (size_t)k * n;
            double* Uk = ws + L.U + (size_t)k * m;
            const double* Kk = ws + L.KD + (size_t)k * C::KDS;
            const double* dk = Kk + m * n;
            double dx[n];
#pragma unroll
            for (int i = 0; i < n; i++) dx[i] = xb[i] - Xk[i];
            double mxg = ninf;
            bool gnan = false;
#pragma unroll
            for (int i = 0; i < m; i++) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < n; c++) acc = fma(Kk[c * m + i], dx[c], acc);
                const double di = dk[i];
                ub[i] = (Uk[i] + acc) + alpha * di;
                if (WRITE) {
                    const double v = fabs(di) / (fabs(ub[i]) + 1.0);
                    if (!gnan) {
                        if (v != v) { mxg = v; gnan = true; }
                        else mxg = dmax(mxg, v);
                    }
                }
            }
            if (WRITE) gs += mxg;
            J += stage_cost<C>(P, xb, ub);
            if (al_on) Jc += knot_al_cost<C>(P, k, lam, mu, xb, ub);
            if (WRITE) {
#pragma unroll
                for (int i = 0; i < n; i++) Xk[i] = xb[i];
#pragma unroll
                for (int i = 0; i < m; i++) Uk[i] = ub[i];
            }
            if (CAND) {
#pragma unroll
                for (int i = 0; i < n; i++) XB[cand_index((size_t)k * n + i, slot, CW)] = xb[i];
#pragma unroll
                for (int i = 0; i < m; i++) UB[cand_index((size_t)k * m + i, slot, CW)] = ub[i];
            }
            double xn[n];
            // dyn_eval (augmented model)
            {
                double dt = P.dt;
                if constexpr (C::MT) {
                    const double h = ub[m - 1];
                    dt = h * h;
                }
                fd_model<C::MODEL, C::INTEG, double>(xn, xb, ub, dt);
                if constexpr (C::INF) {
#pragma unroll
                    for (int i = 0; i < C::n0; i++) xn[i] = xn[i] + ub[C::m0 + i];
                }
                if constexpr (C::MT) xn[n - 1] = ub[m - 1];
            }
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
            if (al_on) Jc += knot_al_cost<C>(P, N - 1, lam, mu, xb, uz);
            if (WRITE) {
                double* XN = ws + L.X + (size_t)(N - 1) * n;
#pragma unroll
                for (int i = 0; i < n; i++) XN[i] = xb[i];
            }
            if (CAND) {
#pragma unroll
                for (int i = 0; i < n; i++) XB[cand_index((size_t)(N - 1) * n + i, slot, CW)] = xb[i];
            }
        }
        Jt = al_on ? (J + Jc) : J;
        grad_sum = gs;
        return ok;
    }
};

// group `grp` of G step sizes: trials grp*G .. grp*G+G-1 (alpha = 2^-trial)
template <class C, int G, int MINB>
__global__ void __launch_bounds__(128, MINB) ls_trial_kernel(const DevProblem Pg, const DevBatch Bt, const DevCtl ctl, const LsCtl lc,
                                                       const int cur, const int grp) {
    extern __shared__ __align__(16) unsigned char ls_tab_raw[];
    DevProblem P = Pg;
    ls_stage_problem(P, Pg, ls_tab_raw);
    const int* list = (grp == 0) ? lc.list[cur] : lc.retry[(grp - 1) & 1];
    const unsigned int na = (grp == 0) ? lc.counts[cur] : lc.counts[2 + ((grp - 1) & 1)];
    // (the host zeroes counts[2 + (grp & 1)] before launching a group >= 2)
    __shared__ __align__(16) double stage_all[(128 / G) * 2 * RolloutStage<C>::SBUF];
    const WsLayout L = ws_layout<C>(P.N, P.Ptot, false);
    const bool al_on = (ctl.mode == 1);
    const int lane = threadIdx.x & 31;
    const int t = lane % G;
    const unsigned int per_block = blockDim.x / G;
    const unsigned int total = (na + per_block - 1) / per_block * per_block;  // keep warps convergent for the ballot
    for (unsigned int a0 = blockIdx.x * per_block + threadIdx.x / G; a0 < total; a0 += gridDim.x * per_block) {
        const bool valid = a0 < na;
        const int b = valid ? list[a0] : 0;
        LsState* st = &lc.st[b];
        TOiLQROptions io = ctl.o.opts_uncon;
        const int ntrial = io.iterations_linesearch + 1;
        const int trial = grp * G + t;
        bool accept = false;
        double Jt = 0.0, expected = 0.0, z = 0.0;
        const double alpha = __longlong_as_double((long long)(1023 - trial) << 52);  // 2^-trial
        const bool bp_fail = valid && (st->bp_fail != 0);  // backward pass aborted: no line search for this problem
        const bool runs = valid && !bp_fail && trial < ntrial;
        const unsigned amask = __ballot_sync(0xffffffffu, runs);
        if (runs) {
            double* ws = lc.ws + (size_t)b * lc.ws_stride;
            double x0[C::n];
#pragma unroll
            for (int i = 0; i < C::n; i++) x0[i] = (i < C::n0) ? Bt.x0[(size_t)b * C::n0 + i] : 0.0;
            const unsigned gm_ = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - t));
            const int nact = __popc(amask & gm_);  // lanes serving this problem (the first nact of its group)
            double* stg = stage_all + (size_t)(threadIdx.x / G) * 2 * RolloutStage<C>::SBUF;
            bool ok;
            if (lc.cand != nullptr) {
                // keep every candidate of this group: G-way interleaved, slot = problem id (bulk) or list position (tail),
                // column = step size within the group; the accept kernel copies the winner instead of re-rolling it
                const size_t xspan = (G == 32) ? cand_span((size_t)P.N * C::n, G) : cand_chunk_span((size_t)P.N, C::n, G);
                const size_t uspan = (G == 32) ? cand_span((size_t)(P.N - 1) * C::m, G) : cand_chunk_span((size_t)(P.N - 1), C::m, G);
                double* XB = lc.cand + (size_t)(lc.cand_by_problem ? (unsigned int)b : a0) * (xspan + uspan);
                double* UB = XB + xspan;
                ok = Rollout<C>::template run_staged<true, G>(P, io, ws, L, x0, alpha, al_on, Jt, XB, UB, t, stg, t, nact, amask);
            } else {
                ok = Rollout<C>::template run_staged<false, G>(P, io, ws, L, x0, alpha, al_on, Jt, nullptr, nullptr, 0, stg, t, nact, amask);
            }
            const double dV0 = st->dV0, dV1 = st->dV1, J_prev = st->J_prev;
            expected = -alpha * (dV0 + alpha * dV1);
            z = (expected > 0) ? (J_prev - Jt) / expected : -1.0;
            const bool cont = (z <= io.line_search_lower_bound || z > io.line_search_upper_bound) && (Jt >= J_prev);
            accept = ok && !cont;
        }
        const unsigned full = __ballot_sync(0xffffffffu, accept);
        const unsigned gm = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane - t));
        const unsigned msk = full & gm;
        if (valid && !bp_fail) {
            if (msk != 0) {
                const int wl = __ffs(msk) - 1;  // warp lane of the first accepted trial
                if (lane == wl) {
                    st->winner = trial;
                    st->Jres = Jt; st->exp_res = expected; st->z_res = z;
                }
            } else if (t == 0) {
                // nobody accepted in this group: next group, or a failed line search
                if ((grp + 1) * G < ntrial) ls_append(lc.retry[grp & 1], &lc.counts[2 + (grp & 1)], b);
            }
        }
    }
}

// accept with stored candidates: warp per problem; the accepted candidate is copied, not re-rolled (same bookkeeping as
// ls_accept_kernel, expressed with the warp-level Solver methods)
template <class C>
__global__ void __launch_bounds__(32) ls_accept_tail_kernel(const DevProblem P, const DevBatch Bt, const DevCtl ctl, const LsCtl lc, const int cur) {
    __shared__ Smem<C> sm;
    const int lane = threadIdx.x;
    const unsigned int na = lc.counts[cur];
    const int N = P.N;
    for (unsigned int a = blockIdx.x; a < na; a += gridDim.x) {
        const int b = lc.list[cur][a];
        LsState* st = &lc.st[b];
        LsSolver<C> s(P, Bt, ctl, sm, lc.ws + (size_t)b * lc.ws_stride, lane);
        s.load(st, b);
        const int bp_fail = st->bp_fail, w = st->winner;
        const double Jwin = st->Jres, ewin = st->exp_res, zwin = st->z_res;
        __syncwarp();
        bool inner_done = false;
        int inner_ok = 1;
        if (bp_fail) {
            s.status |= (bp_fail == 2) ? TO_STATUS_NOT_PD_SQRT : TO_STATUS_REG_DIVERGED;
            if (lane == 0) st->bp_fail = 0;
            inner_done = true;
            inner_ok = 0;
        } else {
            const int ntrial = s.io.iterations_linesearch + 1;
            const double J_prev = s.J_prev;
            double Jres;
            bool err;
            if (w >= 0) {
                s.ls_count += (unsigned long long)(w + 1);
                Jres = Jwin;
                s.fp_expected = ewin; s.fp_z = zwin;
                s.fp_alpha = __longlong_as_double((long long)(1023 - w) << 52);
                err = (Jres > J_prev);
                if (!err && !(Jres > s.io.max_cost_value)) {
                    const int W = lc.cand_width, col = w % W;
                    const int nx = N * C::n, nu = (N - 1) * C::m;
                    if (W == 32) {  // tail buffer: element-interleaved
                        const size_t per = cand_span((size_t)nx, W) + cand_span((size_t)nu, W);
                        const double* XB = lc.cand + (size_t)(lc.cand_by_problem ? (unsigned int)b : a) * per;
                        const double* UB = XB + cand_span((size_t)nx, W);
                        // (unrolled: 8 independent loads in flight per lane -- the copy is a DRAM-latency chain otherwise)
#pragma unroll 8
                        for (int e = lane; e < nx; e += 32) s.ws[s.L.X + e] = XB[cand_index((size_t)e, col, W)];
#pragma unroll 8
                        for (int e = lane; e < nu; e += 32) s.ws[s.L.U + e] = UB[cand_index((size_t)e, col, W)];
                    } else {        // bulk buffer: 4-component chunks, one 32-byte sector of the winner per lane and load pair
                        constexpr int XC = (C::n + 3) / 4, UC = (C::m + 3) / 4;
                        const size_t xspan = cand_chunk_span((size_t)N, C::n, W), uspan = cand_chunk_span((size_t)(N - 1), C::m, W);
                        const double* XB = lc.cand + (size_t)(lc.cand_by_problem ? (unsigned int)b : a) * (xspan + uspan);
                        const double* UB = XB + xspan;
#pragma unroll 4
                        for (int q = lane; q < N * XC; q += 32) {
                            const int k = q / XC, c = q - k * XC;
                            const double2* src = reinterpret_cast<const double2*>(XB + cand_chunk_index((size_t)k, c, C::n, col, W));
                            const double2 v0 = src[0], v1 = src[1];
                            double* dst = s.ws + s.L.X + (size_t)k * C::n + 4 * c;
                            dst[0] = v0.x;
                            if (4 * c + 1 < C::n) dst[1] = v0.y;
                            if (4 * c + 2 < C::n) dst[2] = v1.x;
                            if (4 * c + 3 < C::n) dst[3] = v1.y;
                        }
#pragma unroll 4
                        for (int q = lane; q < (N - 1) * UC; q += 32) {
                            const int k = q / UC, c = q - k * UC;
                            const double2* src = reinterpret_cast<const double2*>(UB + cand_chunk_index((size_t)k, c, C::m, col, W));
                            const double2 v0 = src[0], v1 = src[1];
                            double* dst = s.ws + s.L.U + (size_t)k * C::m + 4 * c;
                            dst[0] = v0.x;
                            if (4 * c + 1 < C::m) dst[1] = v0.y;
                            if (4 * c + 2 < C::m) dst[2] = v1.x;
                            if (4 * c + 3 < C::m) dst[3] = v1.y;
                        }
                    }
                    __syncwarp();
                }
            } else {
                s.ls_count += (unsigned long long)ntrial;
                Jres = s.eval_cost();
                s.fp_expected = 0.0; s.fp_z = 0.0; s.fp_alpha = 0.0;
                s.reg_update(true);
                s.rho += s.io.bp_reg_fp;
                err = (Jres > J_prev);
            }
            s.steps += 1;
            if (err) {
                s.status |= TO_STATUS_COST_INCREASED;
                inner_done = true;
                inner_ok = 0;
            } else if (Jres > s.io.max_cost_value) {
                s.status |= TO_STATUS_COST_BLOWUP;
                inner_done = true;
            } else {
                const double dJ = fabs(Jres - J_prev);
                s.J_prev = Jres;
                s.record_inner(Jres, dJ);
                if (s.inner_converged()) {
                    inner_done = true;
                } else {
                    s.inner_i += 1;
                    if (s.inner_i > s.io.iterations) inner_done = true;
                }
            }
        }
        s.store();
        if (lane == 0) {
            if (inner_done) {
                st->inner_ok = inner_ok;
                ls_append(lc.outer_list, &lc.counts[4], b);
            } else {
                ls_append(lc.list[cur ^ 1], &lc.counts[cur ^ 1], b);
            }
        }
    }
}

// accept the step of every active problem, record the iteration, decide who continues
template <class C>
__global__ void __launch_bounds__(64) ls_accept_kernel(const DevProblem P, const DevBatch Bt, const DevCtl ctl, const LsCtl lc, const int cur) {
    const unsigned int na = lc.counts[cur];
    const WsLayout L = ws_layout<C>(P.N, P.Ptot, false);
    const bool al_on = (ctl.mode == 1);
    const int N = P.N;
    for (unsigned int a = blockIdx.x * blockDim.x + threadIdx.x; a < na; a += gridDim.x * blockDim.x) {
        const int b = lc.list[cur][a];
        LsState* st = &lc.st[b];
        double* ws = lc.ws + (size_t)b * lc.ws_stride;
        TOiLQROptions io = ctl.o.opts_uncon;
        io.cost_tolerance = st->cost_tol;
        io.gradient_norm_tolerance = st->grad_tol;
        const int ntrial = io.iterations_linesearch + 1;
        const double J_prev = st->J_prev;
        const int w = st->winner;
        double x0[C::n];
#pragma unroll
        for (int i = 0; i < C::n; i++) x0[i] = (i < C::n0) ? Bt.x0[(size_t)b * C::n0 + i] : 0.0;
        double Jres, gsum = 0.0;
        bool err;
        bool copied = false;
        double rho = st->rho, drho = st->drho;
        if (st->bp_fail) {
            // 1: the reference would restart its backward pass forever (TO_STATUS_REG_DIVERGED); 2: PosDefException in the
            // square-root pass (TO_STATUS_NOT_PD_SQRT).  Either way the reference aborts the solve here.
            st->status |= (st->bp_fail == 2) ? TO_STATUS_NOT_PD_SQRT : TO_STATUS_REG_DIVERGED;
            st->bp_fail = 0;
            st->winner = -1;
            st->inner_ok = 0;
            ls_append(lc.outer_list, &lc.counts[4], b);
            continue;
        }
        if (w >= 0) {
            st->ls_count += (unsigned long long)(w + 1);
            Jres = st->Jres;
            st->fp_expected = st->exp_res; st->fp_z = st->z_res;
            st->fp_alpha = __longlong_as_double((long long)(1023 - w) << 52);
            err = (Jres > J_prev);
            if (!err && !(Jres > io.max_cost_value)) {
                double Jt;
                Rollout<C>::template run<true>(P, io, ws, L, x0, st->fp_alpha, al_on, Jt, gsum);
                copied = true;
            }
        } else {
            // line search failed (forward_pass.jl:22-37): X̄ <- X, Ū <- U, J recomputed, regularisation bumped
            st->ls_count += (unsigned long long)ntrial;
            double J = 0.0, Jc = 0.0;
            for (int k = 0; k < N; k++) {
                double x[C::n], u[C::m];
#pragma unroll
                for (int i = 0; i < C::n; i++) x[i] = ws[L.X + (size_t)k * C::n + i];
#pragma unroll
                for (int i = 0; i < C::m; i++) u[i] = (k < N - 1) ? ws[L.U + (size_t)k * C::m + i] : 0.0;
                J += (k < N - 1) ? stage_cost<C>(P, x, u) : term_cost<C>(P, x);
            }
            if (al_on) {
                for (int k = 0; k < N; k++) {
                    double x[C::n], u[C::m];
#pragma unroll
                    for (int i = 0; i < C::n; i++) x[i] = ws[L.X + (size_t)k * C::n + i];
#pragma unroll
                    for (int i = 0; i < C::m; i++) u[i] = (k < N - 1) ? ws[L.U + (size_t)k * C::m + i] : 0.0;
                    Jc += knot_al_cost<C>(P, k, ws + L.LAM, ws + L.MU, x, u);
                }
            }
            Jres = al_on ? (J + Jc) : J;
            st->fp_expected = 0.0; st->fp_z = 0.0; st->fp_alpha = 0.0;
            const double f = io.bp_reg_increase_factor;
            drho = dmax(drho * f, f);
            rho = dmax(rho * drho, io.bp_reg_min);
            rho += io.bp_reg_fp;
            err = (Jres > J_prev);
        }
        st->rho = rho; st->drho = drho;
        st->steps += 1;
        bool inner_done = false;
        int inner_ok = 1;
        if (err) {
            st->status |= TO_STATUS_COST_INCREASED;
            inner_done = true;
            inner_ok = 0;
        } else if (Jres > io.max_cost_value) {
            st->status |= TO_STATUS_COST_BLOWUP;
            inner_done = true;
        } else {
            if (!copied) {
                // gradient over the unchanged U and the new d (ilqr_methods.jl:122-129)
                const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
                for (int k = 0; k < N - 1; k++) {
                    const double* d = ws + L.KD + (size_t)k * C::KDS + C::m * C::n;
                    const double* Uk = ws + L.U + (size_t)k * C::m;
                    double mx = ninf;
                    bool isnan_ = false;
#pragma unroll
                    for (int i = 0; i < C::m; i++) {
                        const double v = fabs(d[i]) / (fabs(Uk[i]) + 1.0);
                        if (!isnan_) {
                            if (v != v) { mx = v; isnan_ = true; }
                            else mx = dmax(mx, v);
                        }
                    }
                    gsum += mx;
                }
            }
            const double dJ = fabs(Jres - J_prev);
            st->J_prev = Jres;
            // record_inner (ilqr_methods.jl:77-89)
            const int iterations = st->iterations + 1;
            st->iterations = iterations;
            st->last_cost = Jres;
            st->last_dJ = dJ;
            double s = gsum;
            s += 0.0;
            const double grad = s / (double)N;
            st->last_grad = grad;
            int dJ_zero = st->dJ_zero;
            if (dJ == 0.0) dJ_zero += 1; else dJ_zero = 0;
            st->dJ_zero = dJ_zero;
            if (Bt.inner_cap > 0) {
                int nrec = st->n_inner_rec;
                if (nrec < Bt.inner_cap) {
                    TOIterRecord r;
                    r.cost = Jres; r.dJ = dJ; r.gradient = grad; r.expected = st->fp_expected; r.z = st->fp_z;
                    r.alpha = st->fp_alpha; r.rho = rho; r.outer = (st->outer_i > 0) ? st->outer_i - 1 : 0; r.iter = iterations;
                    Bt.inner[(size_t)b * Bt.inner_cap + nrec] = r;
                    st->n_inner_rec = nrec + 1;
                } else {
                    st->status |= TO_STATUS_TRACE_TRUNC;
                }
            }
            // evaluate_convergence (ilqr_methods.jl:139-162)
            bool conv = false;
            if (0.0 < dJ && dJ < io.cost_tolerance) conv = true;
            else if (grad < io.gradient_norm_tolerance) conv = true;
            else if (iterations >= io.iterations) conv = true;
            else if (dJ_zero > io.dJ_counter_limit) conv = true;
            if (conv) {
                inner_done = true;
            } else {
                const int ii = st->inner_i + 1;
                st->inner_i = ii;
                if (ii > io.iterations) inner_done = true;
            }
        }
        st->winner = -1;
        if (inner_done) {
            st->inner_ok = inner_ok;
            ls_append(lc.outer_list, &lc.counts[4], b);
        } else {
            ls_append(lc.list[cur ^ 1], &lc.counts[cur ^ 1], b);
        }
    }
}


}  // namespace tob
#include "resident.cuh"
#include "pn.cuh"
namespace tob {

// ------------------------------------------------------------------------------------------
// host-side launch table of the lockstep kernels of one configuration
// ------------------------------------------------------------------------------------------
constexpr int LS_BP_WARPS = 4;
constexpr int LS_TRIAL_G = 8;

typedef void (*LsJacFn)(const DevProblem, const LsCtl, const int);
typedef void (*LsTrialFn)(const DevProblem, const DevBatch, const DevCtl, const LsCtl, const int, const int);
typedef void (*LsBpFn)(const DevProblem, const DevCtl, const LsCtl, const int);
// Block shape of the lane-group backward pass: 4 warps x 3 blocks per SM at 168 registers.  Measured alternatives on the 65,536
// quadrotor batch (profiles/r02q_bp_block_shapes.log; bp phase of a step, 3,634 ms for the default): more warps per SM under a
// register cap all lose to their spills -- 5 warps x 3 blocks at 136 registers 4,814 ms, 3 x 5 at 136: 4,230 ms, 2 x 7 at 144:
// 4,144 ms, 4 x 4 at 128 (kept as TRAJOPT_B200_BP_MINB=4): 3,700 ms at 16,384 problems vs 3,450.  The kernel wants ~190
// registers; its latency cannot be hidden by occupancy it does not have room for.
template <class C> LsBpFn ls_bp_variant(int minb, int warps = LS_BP_WARPS) {
    if constexpr (C::MODEL == 4) {
        if (minb == 4) return ls_bp_kernel<C, LS_BP_WARPS, 4>;
    }
    return ls_bp_kernel<C, LS_BP_WARPS, 3>;
}
template <class C> LsJacFn ls_jac_variant(int pc, int minb) {
    // quadrotor: one partial direction per thread at <= 168 registers (measured best, profiles/r01f; the other register-cap /
    // chunking variants of round 1 were removed after the sweep -- they tripled the compile time of the instance)
    if constexpr (C::MODEL == 4) return ls_jac_kernel<C, 1, 3>;
    return ls_jac_kernel<C, C::PC, 2>;
}
template <class C> LsTrialFn ls_trial_variant(int minb, bool all) {
    if constexpr (C::MODEL == 4) {
        if (all) return ls_trial_kernel<C, 32, 1>;  // tail mode: a latency chain of one warp per problem, no register cap (profiles/r01t1)
    }
    return all ? ls_trial_kernel<C, 32, 3> : ls_trial_kernel<C, LS_TRIAL_G, 3>;
}

template <class C> constexpr int ls_bp_cta_threads() {
    // one thread per element of the largest per-knot product set (n*n + m*n + m*m + n + m), rounded to warps, within 64..256
    constexpr int need = C::n * C::n + C::m * C::n + C::m * C::m + C::n + C::m;
    return need > 192 ? 256 : (need > 128 ? 192 : (need > 64 ? 128 : 64));
}

template <class C> LsBpFn ls_bp_cta_variant(int minb) {
    return ls_bp_cta_kernel<C, ls_bp_cta_threads<C>(), 2>;
}

template <class C> constexpr int ls_res_threads() {
    // 256 threads for every model: the Riccati recursion of a small model needs only 64, but the knot-parallel phases (Jacobians,
    // expansion, the cost of all step sizes) of a problem with a large constraint set (car_escape: 177 rows per knot) were
    // 2 warps walking 101 knots (profiles/r02k: T2 47 % of the iteration)
    return 256;
}
typedef void (*LsResFn)(const DevProblem, const DevBatch, const DevCtl, const LsCtl, const int);
template <class C> LsResFn ls_resident_variant(int minb) {
    constexpr int JPC = (C::MODEL == 4) ? 1 : C::PC;  // partial directions per Jacobian item (values do not depend on the chunking)
    if constexpr (C::MODEL == 4) {
        if (minb == 1) return ls_resident_kernel<C, ls_res_threads<C>(), 1, JPC>;  // no register cap: half the resident CTAs, no spills
    }
    return ls_resident_kernel<C, ls_res_threads<C>(), 2, JPC>;
}

template <class C> unsigned long long ls_ws_doubles_fn(int N, int Ptot) { return ws_layout<C>(N, Ptot, false).total; }

template <class C> int ls_setup_fn(int sm_count, int N, int nrows, LsGrids* g) {
    constexpr int GPB = (32 / ls_group_size<C>()) * LS_BP_WARPS;
    g->tab_bytes = ls_tab_bytes(N, nrows);
    g->bp_smem = ls_bp_stride<C>() * GPB + g->tab_bytes;   // ls_expand_kernel keeps the default block shape
    g->bp_groups_per_block = GPB;
    g->trial_group = LS_TRIAL_G;
    g->bp_minb = 3;
    g->bp_warps = LS_BP_WARPS;
    if constexpr (C::MODEL == 4) {
        if (const char* env = getenv("TRAJOPT_B200_BP_MINB")) { const int v = atoi(env); if (v == 3 || v == 4) g->bp_minb = v; }
    }
    g->bp_kernel_smem = ls_bp_stride<C>() * (32 / ls_group_size<C>()) * g->bp_warps + g->tab_bytes;
    if (cudaFuncSetAttribute(ls_bp_variant<C>(g->bp_minb, g->bp_warps), cudaFuncAttributeMaxDynamicSharedMemorySize, g->bp_kernel_smem) != cudaSuccess) return -1;
    cudaFuncSetAttribute(ls_bp_variant<C>(g->bp_minb, g->bp_warps), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ls_init_kernel<C>, 32, 0);
    g->init = sm_count * (nb > 0 ? nb : 1);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ls_outer_kernel<C>, 32, g->tab_bytes);
    g->outer = sm_count * (nb > 0 ? nb : 1);
    g->jac_pc = C::PC;
    g->jac_minb = 2;
    g->trial_minb = 3;
    g->trial_all_minb = 3;
    if constexpr (C::MODEL == 4) {  // quadrotor: kernel variants selectable at run time (tuning)
        g->jac_pc = 1;    // measured (profiles/r01f): one partial direction per thread at <=168 registers (12 warps/SM)
        g->jac_minb = 3;  // beats two directions at 255 registers (8 warps/SM): 4.4 vs 5.4 ms per 16,384 problems
        g->trial_all_minb = 1;  // tail mode is a latency chain of one warp per problem: no register cap (0.377 -> 0.364 ms, profiles/r01t1)
    }
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ls_jac_variant<C>(g->jac_pc, g->jac_minb), 128, 0);
    g->jac = sm_count * (nb > 0 ? nb : 1);
    g->occ_jac = nb;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ls_bp_variant<C>(g->bp_minb, g->bp_warps), 32 * g->bp_warps, g->bp_kernel_smem);
    if (nb < 1) return -2;
    g->bp = sm_count * nb;
    g->occ_bp = nb;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ls_trial_variant<C>(g->trial_minb, false), 128, g->tab_bytes);
    g->trial = sm_count * (nb > 0 ? nb : 1);
    g->occ_trial = nb;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ls_accept_kernel<C>, 64, 0);
    g->accept = sm_count * (nb > 0 ? nb : 1);
    // latency path of the backward pass: expansion kernel (lane groups, same block shape as ls_bp_kernel) + CTA per problem
    if (cudaFuncSetAttribute(ls_expand_kernel<C, LS_BP_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, g->bp_smem) != cudaSuccess) return -4;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ls_expand_kernel<C, LS_BP_WARPS>, 32 * LS_BP_WARPS, g->bp_smem);
    g->expand = sm_count * (nb > 0 ? nb : 1);
    g->bp_cta_smem = (int)sizeof(BpCtaSmem<C>);
    g->bp_cta_minb = 2;
    if (cudaFuncSetAttribute(ls_bp_cta_variant<C>(g->bp_cta_minb), cudaFuncAttributeMaxDynamicSharedMemorySize, g->bp_cta_smem) != cudaSuccess) return -5;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ls_bp_cta_variant<C>(g->bp_cta_minb), ls_bp_cta_threads<C>(), g->bp_cta_smem);
    g->occ_bp_cta = nb;
    g->bp_cta = sm_count * (nb > 0 ? nb : 1);
    // CTA-per-problem resident kernel
    g->res_threads = ls_res_threads<C>();
    g->res_smem = (int)ResLayout<C, ls_res_threads<C>()>::total(N, nrows);
    g->res_minb = 2;
    // quadrotor: one CTA per SM without a register cap (profiles/r02r: resident phase of the 65,536 batch 1,188 -> 1,046 ms;
    // half the slots, so the hand-over comes a few ticks later)
    if constexpr (C::MODEL == 4) g->res_minb = 1;
    if (const char* env = getenv("TRAJOPT_B200_RESIDENT_MINB")) { const int v = atoi(env); if (v == 1 || v == 2) g->res_minb = v; }
    g->res_capacity = 0;
    if (cudaFuncSetAttribute(ls_resident_variant<C>(g->res_minb), cudaFuncAttributeMaxDynamicSharedMemorySize, g->res_smem) == cudaSuccess) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ls_resident_variant<C>(g->res_minb), g->res_threads, g->res_smem);
        g->res_capacity = sm_count * (nb > 0 ? nb : 0);
    } else {
        cudaGetLastError();
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

template <class C> void ls_launch_fn(int phase, const LsGrids& g, cudaStream_t st, const DevProblem& P, const DevBatch& B, const DevCtl& c,
                                     const LsCtl& lc, int cur, int grp) {
    switch (phase) {
        case LS_PHASE_INIT: ls_init_kernel<C><<<g.init, 32, 0, st>>>(P, B, c, lc); break;
        case LS_PHASE_JAC: ls_jac_variant<C>(g.jac_pc, g.jac_minb)<<<g.jac, 128, 0, st>>>(P, lc, cur); break;
        case LS_PHASE_BP: ls_bp_variant<C>(g.bp_minb, g.bp_warps)<<<g.bp, 32 * g.bp_warps, g.bp_kernel_smem, st>>>(P, c, lc, cur); break;
        case LS_PHASE_EXPAND: ls_expand_kernel<C, LS_BP_WARPS><<<grp > 0 ? grp : g.expand, 32 * LS_BP_WARPS, g.bp_smem, st>>>(P, c, lc, cur); break;
        case LS_PHASE_BP_CTA: ls_bp_cta_variant<C>(g.bp_cta_minb)<<<grp > 0 ? grp : g.bp_cta, ls_bp_cta_threads<C>(), g.bp_cta_smem, st>>>(P, c, lc, cur); break;
        case LS_PHASE_BP_SQRT:
            ls_expand_sqrt_kernel<C><<<g.jac, 64, 0, st>>>(P, c, lc, cur);
            if (grp > 0) ls_bp_sqrt_warp_kernel<C><<<grp, 32, 0, st>>>(P, c, lc, cur);   // few live problems: warp per problem
            else ls_bp_sqrt_kernel<C><<<g.accept, 64, 0, st>>>(P, c, lc, cur);
            break;
        case LS_PHASE_TRIAL: ls_trial_variant<C>(g.trial_minb, false)<<<g.trial, 128, g.tab_bytes, st>>>(P, B, c, lc, cur, grp); break;
        case LS_PHASE_TRIAL_ALL: ls_trial_variant<C>(g.trial_all_minb, true)<<<g.trial, 128, g.tab_bytes, st>>>(P, B, c, lc, cur, 0); break;
        case LS_PHASE_ACCEPT: ls_accept_kernel<C><<<g.accept, 64, 0, st>>>(P, B, c, lc, cur); break;
        case LS_PHASE_OUTER: ls_outer_kernel<C><<<g.outer, 32, g.tab_bytes, st>>>(P, B, c, lc, cur); break;
        case LS_PHASE_ACCEPT_TAIL: ls_accept_tail_kernel<C><<<g.outer, 32, 0, st>>>(P, B, c, lc, cur); break;
        case LS_PHASE_RESIDENT: ls_resident_variant<C>(g.res_minb)<<<grp, g.res_threads, g.res_smem, st>>>(P, B, c, lc, cur); break;
        // split line search (tail mode, large constraint sets; resident.cuh): chains, then costs of all (knot, step size), then the pick
        case LS_PHASE_SPLIT_CHAIN: ls_split_chain_kernel<C><<<g.trial, 128, g.tab_bytes, st>>>(P, B, c, lc, cur); break;
        case LS_PHASE_SPLIT_COST: ls_split_cost_kernel<C><<<grp, 256, g.tab_bytes, st>>>(P, c, lc, cur); break;
        case LS_PHASE_SPLIT_PICK: ls_split_pick_kernel<C><<<g.outer, 128, 0, st>>>(P, c, lc, cur); break;
    }
}

template <class C> int ls_pn_setup_fn(int sm_count, int N, int nrows, int* slots, int* smem) {
    *slots = 0;
    *smem = 0;
    if constexpr (C::MT || PnDims<C>::R > 64) {
        return 0;  // MinTimeCost has no hessian! in the reference (projected_newton.jl:137-146 would throw): not available
    } else {
        const int bytes = (int)(((size_t)ls_tab_bytes(N, nrows) + 15) & ~(size_t)15) + (int)sizeof(PnSmem<C>);
        if (cudaFuncSetAttribute(ls_pn_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) { cudaGetLastError(); return -1; }
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ls_pn_kernel<C>, LS_PN_THREADS, bytes);
        if (nb < 1) return -2;
        *slots = sm_count * nb;
        *smem = bytes;
        return 0;
    }
}
template <class C> unsigned long long ls_pn_scratch_fn(int N, int Ptot) { return pn_layout<C>(N, Ptot).total; }
template <class C> void ls_pn_launch_fn(int grid, int smem, cudaStream_t st, const DevProblem& P, const DevBatch& B, const LsCtl& lc, int n_steps,
                                        double feas_tol, double act_tol, double* scratch, unsigned long long stride) {
    if constexpr (!(C::MT || PnDims<C>::R > 64)) {
        PnOptsDev po{n_steps, feas_tol, act_tol};
        ls_pn_kernel<C><<<grid, LS_PN_THREADS, smem, st>>>(P, B, lc, po, scratch, stride);
    }
}

template <class C> KernelInfo make_info() {
    KernelInfo k;
    k.model = C::MODEL; k.integ = C::INTEG; k.inf = C::INF; k.mt = C::MT; k.n = C::n; k.m = C::m;
    k.smem_bytes = sizeof(Smem<C>);
    k.ws_doubles = ws_doubles_fn<C>;
    k.debug_doubles = debug_doubles_fn<C>;
    k.max_blocks_per_sm = max_blocks_fn<C>;
    k.launch = launch_fn<C>;
    k.ls_ws_doubles = ls_ws_doubles_fn<C>;
    k.ls_setup = ls_setup_fn<C>;
    k.ls_launch = ls_launch_fn<C>;
    k.pn_setup = ls_pn_setup_fn<C>;
    k.pn_scratch_doubles = ls_pn_scratch_fn<C>;
    k.pn_launch = ls_pn_launch_fn<C>;
    return k;
}

}  // namespace tob
