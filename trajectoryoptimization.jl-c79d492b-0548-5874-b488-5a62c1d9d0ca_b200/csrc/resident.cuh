// B200 engine, latency path: ONE CTA per problem runs whole iLQR iterations -- and the augmented-Lagrangian outer loop --
// without returning to the host (ls_resident_kernel).
//
// When few problems are still live (the tail of a batch: a handful of problems that need > 1,500 iLQR iterations, or a
// small batch altogether) the lockstep tick is a pure latency chain: seven kernel launches per iteration, every problem
// waiting for the slowest one.  Here the problem is resident in a CTA until it is solved.  Per iteration:
//
//   jacobians    knot x partial-direction items over the threads of the CTA (ls_jac_item)        (src/model.jl:491-512)
//   expansion    knot-parallel, lane groups (BpGroup::expansion) -> Q trajectory                 (cost.jl:183-198, AL :186-229)
//   Riccati      thread per output element, the CTA recursion of ls_bp_cta_problem               (backward_pass.jl:9-85)
//   line search  T1 state chains of all step sizes (lane = step size).  The quadrotor's rk3 stage is split over three ROLE
//                WARPS that own disjoint state elements and exchange stage vectors through shared memory (named barrier),
//                so the dependent chain per knot is one third of the single-thread chain; the cost is NOT on that chain:
//                T2 evaluates stage + AL cost of every (step size, knot) in parallel from the stored candidates,
//                T3 sums them in knot order (bitwise the sequential sums of rollout.jl / objective.jl:40-48)
//                and picks the first accepted step size                                           (forward_pass.jl:5-85)
//   accept       copy of the winner by the whole CTA, bookkeeping / convergence / outer loop by warp 0 with the same
//                LsSolver methods the lockstep kernels use                                        (ilqr_methods.jl:30-45,77-162)
//
// Every per-element expression is the one of the lockstep kernels (same device functions), so results are bitwise the same.
#pragma once

namespace tob {

// role warps of the line-search state chain: how many, and which state elements [lo, hi) each one owns
template <class C> struct RollRoles {
    static constexpr bool split = (C::MODEL == 4 && C::INTEG == 0 && !C::INF && !C::MT);
    static constexpr int R = split ? 3 : 1;
};

template <class C>
struct alignas(16) ResTrialSmem {
    static constexpr int SS = C::n + C::m + C::KDS;   // x_k, u_k, K_k, d_k
    static constexpr int SSP = (SS + 1) & ~1;
    double kin[2][SSP];          // per-knot inputs, double-buffered (cp.async one knot ahead)
    double ex[2][C::n][32];      // stage vectors, [buffer][state element][step size]
    int bad[2][4][32];           // [knot parity][role][step size]: the rollout left the state / control box at this knot
};

struct ResFlags {
    int okf[32];                 // rollout of step size t stayed inside the box
    int copy, w, cont, pad;
};

__device__ __forceinline__ void res_role_barrier(int nthreads) {
    __syncwarp();
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// One role warp of the split rk3 state chain (integration.jl:149-158, rollout.jl:2-23): owns state elements LO..HI-1 of every
// step size (lane t = step size 2^-t).  All role warps call this together; `wtid` = thread index among the R*32 role threads.
template <class C, int LO, int HI, int R>
__device__ __forceinline__ void res_rollout_role(const DevProblem& P, const TOiLQROptions& io, const double* ws, const WsLayout& L,
                                                 ResTrialSmem<C>& ts, ResFlags& fl, double* XB, double* UB, const double* x0,
                                                 const double alpha, const int t, const bool runs, const int role, const int wtid,
                                                 long long* prof = nullptr) {
    constexpr int n = C::n, m = C::m, CNT = HI - LO, SS = ResTrialSmem<C>::SS;
    const int N = P.N;
    long long pt0 = prof ? clock64() : 0;
    auto ptick = [&](int section) {  // diagnostics: cycles per section of the knot loop (thread 0 of the profiled CTA)
        if (prof) {
            const long long tt = clock64();
            prof[section] += tt - pt0;
            pt0 = tt;
        }
    };
    auto prefetch = [&](int k, int buf) {
        const double* xk = ws + L.X + (size_t)k * n;
        const double* uk = ws + L.U + (size_t)k * m;
        const double* kd = ws + L.KD + (size_t)k * C::KDS;
        for (int e = wtid; e < SS; e += R * 32) cp_async8(&ts.kin[buf][e], (e < n) ? (xk + e) : ((e < n + m) ? (uk + (e - n)) : (kd + (e - n - m))));
    };
    double xo[CNT], k1[CNT], k2[CNT];
#pragma unroll
    for (int i = 0; i < CNT; i++) { xo[i] = x0[LO + i]; ts.ex[0][LO + i][t] = xo[i]; }
    prefetch(0, 0);
    cp_async_wait_all();
    res_role_barrier(R * 32);
    bool ok = runs;
    int xb = 0;
    const double dt = P.dt;
    for (int k = 0; k < N - 1; k++) {
        if (k > 0) {
            int bad = 0;
#pragma unroll
            for (int r = 0; r < R; r++) bad |= ts.bad[(k - 1) & 1][r][t];
            if (bad) ok = false;
        }
        if (k + 1 < N - 1) prefetch(k + 1, (k + 1) & 1);
        const double* Xk = ts.kin[k & 1];
        const double* Uk = Xk + n;
        const double* Kk = Uk + m;
        const double* dk = Kk + m * n;
        double u[m];
        int badl = 0;
        ptick(0);
        if (ok) {
            double x[n], dx[n], f[n];
#pragma unroll
            for (int i = 0; i < n; i++) x[i] = ts.ex[xb][i][t];
#pragma unroll
            for (int i = 0; i < n; i++) dx[i] = x[i] - Xk[i];
#pragma unroll
            for (int i = 0; i < m; i++) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < n; c++) acc = fma(Kk[c * m + i], dx[c], acc);
                u[i] = (Uk[i] + acc) + alpha * dk[i];
            }
            // candidate trajectory: every role its own state elements, role 0 the controls
#pragma unroll
            for (int i = 0; i < CNT; i++) XB[cand_index((size_t)k * n + LO + i, t, 32)] = xo[i];
            if (LO == 0) {
                double mu_ = 0.0;
                bool bad = false;
#pragma unroll
                for (int i = 0; i < m; i++) {
                    UB[cand_index((size_t)k * m + i, t, 32)] = u[i];
                    const double a = fabs(u[i]);
                    if (a != a) bad = true;
                    mu_ = dmax(mu_, a);
                }
                if (bad || !(mu_ < io.max_control_value)) badl = 1;
            }
            f_model<C::MODEL, double>(f, x, u);
#pragma unroll
            for (int i = 0; i < CNT; i++) {
                k1[i] = f[LO + i] * dt;
                ts.ex[xb ^ 1][LO + i][t] = xo[i] + k1[i] / 2.0;
            }
        }
        ptick(1);
        res_role_barrier(R * 32);
        ptick(2);
        if (ok) {
            double x[n], f[n];
#pragma unroll
            for (int i = 0; i < n; i++) x[i] = ts.ex[xb ^ 1][i][t];
            f_model<C::MODEL, double>(f, x, u);
#pragma unroll
            for (int i = 0; i < CNT; i++) {
                k2[i] = f[LO + i] * dt;
                ts.ex[xb][LO + i][t] = (xo[i] - k1[i]) + 2.0 * k2[i];
            }
        }
        ptick(3);
        res_role_barrier(R * 32);
        ptick(4);
        if (ok) {
            double x[n], f[n];
#pragma unroll
            for (int i = 0; i < n; i++) x[i] = ts.ex[xb][i][t];
            f_model<C::MODEL, double>(f, x, u);
            double mx = 0.0;
            bool bad = false;
#pragma unroll
            for (int i = 0; i < CNT; i++) {
                const double k3 = f[LO + i] * dt;
                const double xn = xo[i] + div6((k1[i] + 4.0 * k2[i]) + k3);
                ts.ex[xb ^ 1][LO + i][t] = xn;
                xo[i] = xn;
                const double a = fabs(xn);
                if (a != a) bad = true;
                mx = dmax(mx, a);
            }
            if (bad || !(mx < io.max_state_value)) badl = 1;
        }
        ts.bad[k & 1][role][t] = badl;
        ptick(5);
        cp_async_wait_all();
        res_role_barrier(R * 32);
        ptick(6);
        xb ^= 1;
    }
    {
        int bad = 0;
#pragma unroll
        for (int r = 0; r < R; r++) bad |= ts.bad[(N - 2) & 1][r][t];
        if (bad) ok = false;
    }
    if (ok) {
#pragma unroll
        for (int i = 0; i < CNT; i++) XB[cand_index((size_t)(N - 1) * n + LO + i, t, 32)] = xo[i];
    }
    if (LO == 0) fl.okf[t] = ok ? 1 : 0;
}

template <class C, int NT>
struct ResLayout {
    static constexpr int GS = ls_group_size<C>();
    static constexpr int NG = NT / GS;
    static __host__ __device__ constexpr size_t a16(size_t v) { return (v + 15) & ~(size_t)15; }
    static constexpr size_t bp_bytes = a16(sizeof(BpCtaSmem<C>));
    static constexpr size_t ex_bytes = a16(sizeof(BpExpSmem<C>)) * NG;
    static constexpr size_t tr_bytes = a16(sizeof(ResTrialSmem<C>)) + a16(2 * RolloutStage<C>::SBUF * sizeof(double));
    static constexpr size_t un_bytes = (bp_bytes > ex_bytes ? (bp_bytes > tr_bytes ? bp_bytes : tr_bytes) : (ex_bytes > tr_bytes ? ex_bytes : tr_bytes));
    static constexpr size_t fixed_bytes = un_bytes + a16(sizeof(ResFlags));
    static __host__ __device__ size_t total(int N, int nrows) { return a16((size_t)ls_tab_bytes(N, nrows)) + fixed_bytes; }
};

// doubles of per-CTA scratch in global memory: stage and AL cost of every (knot, step size) of the running line search
__host__ __device__ inline size_t res_scratch_doubles(int N) { return 2 * (size_t)N * 32; }

// ---- the phases of one iteration as separate (NOT inlined) functions: inlined into one kernel, the line search's live state
// (13-element stage vectors per role) and the loop-invariant problem description left the register allocator so little room
// that the 13-term products of the Riccati recursion were scheduled load -> use -> load -> use (profiles/r02e: 830 cycles for
// a chain that needs ~250); with their own register allocation the phases are scheduled like the stand-alone kernels ----
template <class C, int NT, int JPC>
__device__ __noinline__ void res_phase_jac(const DevProblem& P, double* ws, const WsLayout& L, const int tid) {
    constexpr int NCH = ls_jac_chunks<C, JPC>();
    const int N = P.N;
    for (int it = tid; it < (N - 1) * NCH; it += NT) {
        const int k = it / NCH;
        ls_jac_item<C, JPC>(P, ws, L, k, it - k * NCH);
    }
}

template <class C, int NT>
__device__ __noinline__ void res_phase_expand(const DevProblem& P, const TOiLQROptions& io, double* ws, const WsLayout& L,
                                              BpExpSmem<C>* exsm, const bool al_on, const int tid) {
    constexpr int n = C::n, m = C::m;
    constexpr int GS = ls_group_size<C>();
    constexpr int NG = NT / GS;
    const int N = P.N;
    const int lane = tid & 31;
    const int g = tid / GS, j = tid % GS;
    const unsigned gmask = (GS == 32) ? 0xffffffffu : (((1u << GS) - 1u) << (lane - j));
    BpExpSmem<C>& es = exsm[g];
    BpGroup<C, BpExpSmem<C>> G(P, es, ws, j, gmask, al_on, io);
    G.load_cost_constants();
    for (int k0 = 0; k0 < N; k0 += NG) {
        const int k = k0 + g;
        if (k < N) {
            const double* xk = ws + L.X + (size_t)k * n;
            __syncwarp(gmask);
            if (k < N - 1) {
                const double* uk = ws + L.U + (size_t)k * m;
                for (int e = j; e < n + m; e += GS) es.xu[e] = (e < n) ? xk[e] : uk[e - n];
            } else {
                for (int e = j; e < n; e += GS) es.xu[e] = xk[e];
            }
            __syncwarp(gmask);
            const int lo = P.knot_lam_off[k];
            G.expansion(k, es.xu, ws + L.LAM + lo, ws + L.MU + lo);
            if (k < N - 1) G.q_store(k);
            else G.q_store_term(k);
        }
    }
}

template <class C, int NT>
__device__ __noinline__ void res_phase_riccati(const DevProblem& P, const TOiLQROptions& io, double* ws, const WsLayout& L,
                                               BpCtaSmem<C>& sm, LsState* st, const int tid, long long* prof) {
    ls_bp_cta_problem<C, NT>(P, io, ws, L, sm, st, tid, prof);
}

// T1: state chains of all step sizes (lane = step size)
template <class C, int NT>
__device__ __noinline__ void res_phase_chains(const DevProblem& P, const TOiLQROptions& io, double* ws, const WsLayout& L,
                                              ResTrialSmem<C>& ts, ResFlags& fl, double* stg, double* XB, double* UB, const double* x0g,
                                              const int ntrial, const bool al_on, const int tid, long long* prof) {
    constexpr int n = C::n;
    constexpr int R = RollRoles<C>::R;
    const int lane = tid & 31, warp = tid >> 5;
    if (warp >= R) return;
    double x0[n];
#pragma unroll
    for (int i = 0; i < n; i++) x0[i] = (i < C::n0) ? x0g[i] : 0.0;
    const double alpha = __longlong_as_double((long long)(1023 - lane) << 52);  // 2^-lane
    const bool runs = lane < ntrial;
    if constexpr (RollRoles<C>::split) {
        if (warp == 0) res_rollout_role<C, 0, 7, R>(P, io, ws, L, ts, fl, XB, UB, x0, alpha, lane, runs, 0, tid, prof);
        else if (warp == 1) res_rollout_role<C, 7, 10, R>(P, io, ws, L, ts, fl, XB, UB, x0, alpha, lane, runs, 1, tid);
        else res_rollout_role<C, 10, 13, R>(P, io, ws, L, ts, fl, XB, UB, x0, alpha, lane, runs, 2, tid);
    } else {
        const unsigned amask = __ballot_sync(0xffffffffu, runs);
        bool ok = false;
        if (runs) {
            double Jt;
            ok = Rollout<C>::template run_staged<true, 32, false>(P, io, ws, L, x0, alpha, al_on, Jt, XB, UB, lane, stg, lane,
                                                                  __popc(amask), amask);
        }
        fl.okf[lane] = ok ? 1 : 0;
    }
}

// T2: stage + AL cost of every (step size, knot) of the rollouts that stayed in the box, from the candidates
template <class C, int NT>
__device__ __noinline__ void res_phase_costs(const DevProblem& P, double* ws, const WsLayout& L, const ResFlags& fl, const double* XB,
                                             const double* UB, double* cst, double* cal, const int ntrial, const bool al_on, const int tid) {
    constexpr int n = C::n, m = C::m;
    const int N = P.N;
    const double* lam = ws + L.LAM;
    const double* mu = ws + L.MU;
    // Large constraint sets (car_escape: 177 rows per knot): one WARP per knot, lane = step size -- the knot's rows / multipliers /
    // penalties are then the same for every lane (broadcast loads instead of 21 private walks through them), and the candidates
    // of a knot are interleaved by step size (one coalesced load per element)
    bool big = false;
    if (al_on) {
        for (int k = 0; k < N; k++) big = big || (P.knot_row_count[k] > RolloutStage<C>::LC);
    }
    if (big) {
        const int lane = tid & 31, warp = tid >> 5;
        const int t = lane;
        for (int k = warp; k < N; k += NT / 32) {
            if (t >= ntrial || !fl.okf[t]) continue;
            double x[n], u[m];
#pragma unroll
            for (int i = 0; i < n; i++) x[i] = XB[cand_index((size_t)k * n + i, t, 32)];
            double cs, ca = 0.0;
            if (k < N - 1) {
#pragma unroll
                for (int i = 0; i < m; i++) u[i] = UB[cand_index((size_t)k * m + i, t, 32)];
                cs = stage_cost<C>(P, x, u);
                if (al_on) {
                    const int lo = P.knot_lam_off[k];
                    ca = knot_al_cost_at<C>(P, k, lam + lo, mu + lo, x, u);
                }
            } else {
#pragma unroll
                for (int i = 0; i < m; i++) u[i] = 0.0;
                cs = term_cost<C>(P, x);
                if (al_on) ca = knot_al_cost<C>(P, N - 1, lam, mu, x, u);
            }
            cst[k * 32 + t] = cs;
            cal[k * 32 + t] = ca;
        }
        return;
    }
    for (int it = tid; it < N * ntrial; it += NT) {
        const int k = it / ntrial, t = it - k * ntrial;
        if (!fl.okf[t]) continue;
        double x[n], u[m];
#pragma unroll
        for (int i = 0; i < n; i++) x[i] = XB[cand_index((size_t)k * n + i, t, 32)];
        double cs, ca = 0.0;
        if (k < N - 1) {
#pragma unroll
            for (int i = 0; i < m; i++) u[i] = UB[cand_index((size_t)k * m + i, t, 32)];
            cs = stage_cost<C>(P, x, u);
            if (al_on) {
                const int lo = P.knot_lam_off[k];
                ca = knot_al_cost_at<C>(P, k, lam + lo, mu + lo, x, u);
            }
        } else {
#pragma unroll
            for (int i = 0; i < m; i++) u[i] = 0.0;
            cs = term_cost<C>(P, x);
            if (al_on) ca = knot_al_cost<C>(P, N - 1, lam, mu, x, u);
        }
        cst[k * 32 + t] = cs;
        cal[k * 32 + t] = ca;
    }
}

template <class C, int NT, int MINB, int JPC>
__global__ void __launch_bounds__(NT, MINB) ls_resident_kernel(const DevProblem Pg, const DevBatch Bt, const DevCtl ctl, const LsCtl lc,
                                                             const int cur) {
    constexpr int n = C::n, m = C::m;
    constexpr int R = RollRoles<C>::R;
    constexpr int GS = ls_group_size<C>();
    constexpr int NG = NT / GS;
    constexpr int NCH = ls_jac_chunks<C, JPC>();
    typedef ResLayout<C, NT> RL;
    static_assert(NT >= 32 * R && NT % 32 == 0, "role warps");
    static_assert(sizeof(Smem<C>) <= RL::un_bytes, "the warp-level expansion of the :l2 / :linf gradient borrows the phase union");
    extern __shared__ __align__(16) unsigned char res_smem[];
    DevProblem P = Pg;
    ls_stage_problem(P, Pg, res_smem);
    unsigned char* un = res_smem + RL::a16((size_t)ls_tab_bytes(Pg.N, Pg.nrows));
    BpCtaSmem<C>& bpsm = *reinterpret_cast<BpCtaSmem<C>*>(un);
    BpExpSmem<C>* exsm = reinterpret_cast<BpExpSmem<C>*>(un);
    ResTrialSmem<C>& ts = *reinterpret_cast<ResTrialSmem<C>*>(un);
    double* stg = reinterpret_cast<double*>(un + RL::a16(sizeof(ResTrialSmem<C>)));
    ResFlags& fl = *reinterpret_cast<ResFlags*>(un + RL::un_bytes);
    Smem<C>& unused_sm = *reinterpret_cast<Smem<C>*>(un);  // LsSolver's warp-level methods used here never touch it

    const int N = P.N;
    const WsLayout L = ws_layout<C>(N, P.Ptot, false);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool al_on = (ctl.mode == 1);
    const TOiLQROptions io = ctl.o.opts_uncon;
    const int ntrial = io.iterations_linesearch + 1;
    const unsigned int na = lc.counts[cur];
    const size_t cand_per = cand_span((size_t)N * n, 32) + cand_span((size_t)(N - 1) * m, 32);
    double* XB = lc.cand + (size_t)blockIdx.x * cand_per;
    double* UB = XB + cand_span((size_t)N * n, 32);
    double* cst = lc.res_scratch + (size_t)blockIdx.x * res_scratch_doubles(N);
    double* cal = cst + (size_t)N * 32;

    for (unsigned int a = blockIdx.x; a < na; a += gridDim.x) {
        const int b = lc.list[cur][a];
        LsState* st = &lc.st[b];
        double* ws = lc.ws + (size_t)b * lc.ws_stride;
        // optional cycle profile (diagnostics: to_debug_enable): thread 0 of CTA 0 accumulates the cycles of every phase
        long long* prof = (ctl.debug && blockIdx.x == 0 && tid == 0) ? reinterpret_cast<long long*>(ctl.debug) + 16 : nullptr;
        // the per-knot sub-profiles of the Riccati recursion / the state chain accumulate in SHARED memory (a global
        // read-modify-write per section would put an L2 round trip on the very chain that is being measured)
        __shared__ long long pacc[32];
        if (prof) {
            for (int i = 0; i < 32; i++) pacc[i] = 0;
        }
        long long pt0 = prof ? clock64() : 0;
        auto ptick = [&](int section) {
            if (prof) {
                const long long t = clock64();
                prof[section] += t - pt0;
                pt0 = t;
            }
        };
        for (;;) {  // one iLQR iteration of this problem
            if (prof) prof[15] += 1;
            // ---- Jacobians ----
            res_phase_jac<C, NT, JPC>(P, ws, L, tid);
            __syncthreads();
            ptick(0);
            // ---- cost / constraint expansion of every knot into the Q trajectory ----
            res_phase_expand<C, NT>(P, io, ws, L, exsm, al_on, tid);
            __syncthreads();
            ptick(1);
            // ---- Riccati recursion ----
            res_phase_riccati<C, NT>(P, io, ws, L, bpsm, st, tid, prof ? pacc : nullptr);
            __syncthreads();
            ptick(2);
            const int bp_fail = st->bp_fail;
            const double dV0 = st->dV0, dV1 = st->dV1;
            if (!bp_fail) {
                // ---- line search: state chains (T1), then costs (T2) ----
                res_phase_chains<C, NT>(P, io, ws, L, ts, fl, stg, XB, UB, Bt.x0 + (size_t)b * C::n0, ntrial, al_on, tid, prof ? pacc + 16 : nullptr);
                __syncthreads();
                ptick(3);
                res_phase_costs<C, NT>(P, ws, L, fl, XB, UB, cst, cal, ntrial, al_on, tid);
                __syncthreads();
                ptick(4);
            }
            // ---- T3 + accept, part 1 (warp 0): sums in knot order, first accepted step size, decision ----
            LsSolver<C> s(P, Bt, ctl, unused_sm, ws, lane);
            bool inner_done = false, err = false;
            int inner_ok = 1;
            double Jres = 0.0;
            if (warp == 0) {
                s.load(st, b);
                if (bp_fail) {
                    s.status |= (bp_fail == 2) ? TO_STATUS_NOT_PD_SQRT : TO_STATUS_REG_DIVERGED;
                    if (lane == 0) { st->bp_fail = 0; fl.copy = 0; fl.w = -1; }
                    inner_done = true;
                    inner_ok = 0;
                } else {
                    const double J_prev = s.J_prev;
                    const double alpha = __longlong_as_double((long long)(1023 - lane) << 52);
                    bool accept = false;
                    double Jt = 0.0, expected = 0.0, z = 0.0;
                    if (lane < ntrial && fl.okf[lane]) {
                        double J = 0.0, Jc = 0.0;
#pragma unroll 4
                        for (int k = 0; k < N; k++) J += cst[k * 32 + lane];
                        if (al_on) {
#pragma unroll 4
                            for (int k = 0; k < N; k++) Jc += cal[k * 32 + lane];
                        }
                        Jt = al_on ? (J + Jc) : J;
                        expected = -alpha * (dV0 + alpha * dV1);
                        z = (expected > 0) ? (J_prev - Jt) / expected : -1.0;
                        const bool cont = (z <= io.line_search_lower_bound || z > io.line_search_upper_bound) && (Jt >= J_prev);
                        accept = !cont;
                    }
                    const unsigned msk = __ballot_sync(0xffffffffu, accept);
                    int copy = 0, w = -1;
                    if (msk != 0) {
                        w = __ffs(msk) - 1;
                        s.ls_count += (unsigned long long)(w + 1);
                        Jres = bcast(Jt, w);
                        s.fp_expected = bcast(expected, w);
                        s.fp_z = bcast(z, w);
                        s.fp_alpha = __longlong_as_double((long long)(1023 - w) << 52);
                        err = (Jres > J_prev);
                        if (!err && !(Jres > s.io.max_cost_value)) copy = 1;
                    } else {
                        // line search failed (forward_pass.jl:22-37): X̄ <- X, Ū <- U, J recomputed, regularisation bumped
                        s.ls_count += (unsigned long long)ntrial;
                        Jres = s.eval_cost();
                        s.fp_expected = 0.0; s.fp_z = 0.0; s.fp_alpha = 0.0;
                        s.reg_update(true);
                        s.rho += s.io.bp_reg_fp;
                        err = (Jres > J_prev);
                    }
                    if (lane == 0) { fl.copy = copy; fl.w = w; }
                }
            }
            __syncthreads();
            ptick(5);
            // ---- X <- X̄, U <- Ū of the accepted step size (ilqr_methods.jl:30-33), by the whole CTA ----
            if (fl.copy) {
                const int w = fl.w;
                const int nx = N * n, nu = (N - 1) * m;
                for (int e = tid; e < nx; e += NT) ws[L.X + e] = XB[cand_index((size_t)e, w, 32)];
                for (int e = tid; e < nu; e += NT) ws[L.U + e] = UB[cand_index((size_t)e, w, 32)];
            }
            __syncthreads();
            ptick(6);
            // ---- accept, part 2 (warp 0): record, convergence, and -- when the inner solve ended -- the outer loop ----
            if (warp == 0) {
                if (!bp_fail) {
                    s.steps += 1;
                    if (err) {
                        s.status |= TO_STATUS_COST_INCREASED;
                        inner_done = true;
                        inner_ok = 0;
                    } else if (Jres > s.io.max_cost_value) {
                        s.status |= TO_STATUS_COST_BLOWUP;
                        inner_done = true;
                    } else {
                        const double dJ = fabs(Jres - s.J_prev);
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
                bool cont = true;
                if (inner_done) cont = s.after_inner(inner_ok != 0);
                s.store();
                if (lane == 0) fl.cont = cont ? 1 : 0;
            }
            __syncthreads();
            ptick(7);
            if (!fl.cont) break;
        }
        if (prof) {
            for (int i = 0; i < 32; i++) { prof[32 + i] += pacc[i]; pacc[i] = 0; }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Split line search of the lockstep tick in tail mode, for problems with large per-knot constraint sets (car_escape: 177 rows).
// ls_trial_kernel evaluates the cost ON the state chain: a thread walks 101 knots x 177 rows, 5-7 ms per tick whatever the number
// of live problems.  Here the chain only rolls the states out and stores the candidates (the accept kernel copies the winner from
// them anyway); the cost of every (problem, knot, step size) is then evaluated side by side from the candidates -- a warp per
// (problem, knot), lane = step size, so the knot's rows / multipliers are broadcast loads -- and a third kernel adds the knot
// costs up IN KNOT ORDER (the chain's own summation order: bit-identical) and picks the first accepted step size.  The same
// decomposition as the phases T1 / T2 / T3 of ls_resident_kernel.
template <class C>
__global__ void __launch_bounds__(128, 3) ls_split_chain_kernel(const DevProblem Pg, const DevBatch Bt, const DevCtl ctl, const LsCtl lc,
                                                                const int cur) {
    extern __shared__ __align__(16) unsigned char ls_tab_raw[];
    DevProblem P = Pg;
    ls_stage_problem(P, Pg, ls_tab_raw);
    __shared__ __align__(16) double stage_all[4 * 2 * RolloutStage<C>::SBUF];
    const int* list = lc.list[cur];
    const unsigned int na = lc.counts[cur];
    const WsLayout L = ws_layout<C>(P.N, P.Ptot, false);
    const bool al_on = (ctl.mode == 1);
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned int per_block = blockDim.x >> 5;
    const size_t per = cand_span((size_t)P.N * C::n, 32) + cand_span((size_t)(P.N - 1) * C::m, 32);
    for (unsigned int a = blockIdx.x * per_block + wib; a < na; a += gridDim.x * per_block) {  // one problem per warp, lane = step size
        const int b = list[a];
        const LsState* st = &lc.st[b];
        const TOiLQROptions io = ctl.o.opts_uncon;
        const int ntrial = io.iterations_linesearch + 1;
        const double alpha = __longlong_as_double((long long)(1023 - lane) << 52);  // 2^-lane
        const bool runs = (st->bp_fail == 0) && lane < ntrial;
        const unsigned amask = __ballot_sync(0xffffffffu, runs);
        bool ok = false;
        if (runs) {
            double* ws = lc.ws + (size_t)b * lc.ws_stride;
            double x0[C::n];
#pragma unroll
            for (int i = 0; i < C::n; i++) x0[i] = (i < C::n0) ? Bt.x0[(size_t)b * C::n0 + i] : 0.0;
            double* XB = lc.cand + (size_t)a * per;
            double* UB = XB + cand_span((size_t)P.N * C::n, 32);
            double Jt;
            ok = Rollout<C>::template run_staged<true, 32, false>(P, io, ws, L, x0, alpha, al_on, Jt, XB, UB, lane,
                                                                  stage_all + (size_t)wib * 2 * RolloutStage<C>::SBUF, lane, __popc(amask), amask);
        }
        lc.split_ok[(size_t)a * 32 + lane] = ok ? 1 : 0;
        __syncwarp();
    }
}

template <class C>
__global__ void __launch_bounds__(256) ls_split_cost_kernel(const DevProblem Pg, const DevCtl ctl, const LsCtl lc, const int cur) {
    constexpr int n = C::n, m = C::m;
    extern __shared__ __align__(16) unsigned char ls_tab_raw[];
    DevProblem P = Pg;
    ls_stage_problem(P, Pg, ls_tab_raw);
    const int N = P.N;
    const unsigned int na = lc.counts[cur];
    const WsLayout L = ws_layout<C>(N, P.Ptot, false);
    const bool al_on = (ctl.mode == 1);
    const int ntrial = ctl.o.opts_uncon.iterations_linesearch + 1;
    const int t = threadIdx.x & 31;
    const unsigned long long nwarps = (unsigned long long)gridDim.x * (blockDim.x >> 5);
    const unsigned long long total = (unsigned long long)na * (unsigned long long)N;
    const size_t per = cand_span((size_t)N * n, 32) + cand_span((size_t)(N - 1) * m, 32);
    for (unsigned long long it = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); it < total; it += nwarps) {
        const unsigned int a = (unsigned int)(it / N);
        const int k = (int)(it - (unsigned long long)a * N);
        if (t >= ntrial || !lc.split_ok[(size_t)a * 32 + t]) continue;
        const int b = lc.list[cur][a];
        const double* ws = lc.ws + (size_t)b * lc.ws_stride;
        const double* lam = ws + L.LAM;
        const double* mu = ws + L.MU;
        const double* XB = lc.cand + (size_t)a * per;
        const double* UB = XB + cand_span((size_t)N * n, 32);
        double x[n], u[m];
#pragma unroll
        for (int i = 0; i < n; i++) x[i] = XB[cand_index((size_t)k * n + i, t, 32)];
        double cs, ca = 0.0;
        if (k < N - 1) {
#pragma unroll
            for (int i = 0; i < m; i++) u[i] = UB[cand_index((size_t)k * m + i, t, 32)];
            cs = stage_cost<C>(P, x, u);
            if (al_on) {
                const int lo = P.knot_lam_off[k];
                ca = knot_al_cost_at<C>(P, k, lam + lo, mu + lo, x, u);
            }
        } else {
#pragma unroll
            for (int i = 0; i < m; i++) u[i] = 0.0;
            cs = term_cost<C>(P, x);
            if (al_on) ca = knot_al_cost<C>(P, N - 1, lam, mu, x, u);
        }
        double* out = lc.split_cost + ((size_t)a * N + k) * 64;
        out[t] = cs;
        out[32 + t] = ca;
    }
}

template <class C>
__global__ void __launch_bounds__(128) ls_split_pick_kernel(const DevProblem P, const DevCtl ctl, const LsCtl lc, const int cur) {
    const int N = P.N;
    const unsigned int na = lc.counts[cur];
    const bool al_on = (ctl.mode == 1);
    const int lane = threadIdx.x & 31;
    const unsigned int per_block = blockDim.x >> 5;
    for (unsigned int a = blockIdx.x * per_block + (threadIdx.x >> 5); a < na; a += gridDim.x * per_block) {
        const int b = lc.list[cur][a];
        LsState* st = &lc.st[b];
        if (st->bp_fail != 0) continue;  // backward pass aborted: no line search for this problem
        const TOiLQROptions io = ctl.o.opts_uncon;
        const int ntrial = io.iterations_linesearch + 1;
        const double alpha = __longlong_as_double((long long)(1023 - lane) << 52);
        bool accept = false;
        double Jt = 0.0, expected = 0.0, z = 0.0;
        if (lane < ntrial && lc.split_ok[(size_t)a * 32 + lane]) {
            const double* c = lc.split_cost + (size_t)a * N * 64;
            double J = 0.0, Jc = 0.0;
#pragma unroll 4
            for (int k = 0; k < N; k++) J += c[(size_t)k * 64 + lane];
            if (al_on) {
#pragma unroll 4
                for (int k = 0; k < N; k++) Jc += c[(size_t)k * 64 + 32 + lane];
            }
            Jt = al_on ? (J + Jc) : J;
            const double dV0 = st->dV0, dV1 = st->dV1, J_prev = st->J_prev;
            expected = -alpha * (dV0 + alpha * dV1);
            z = (expected > 0) ? (J_prev - Jt) / expected : -1.0;
            const bool cont = (z <= io.line_search_lower_bound || z > io.line_search_upper_bound) && (Jt >= J_prev);
            accept = !cont;
        }
        const unsigned msk = __ballot_sync(0xffffffffu, accept);
        if (msk != 0 && lane == __ffs(msk) - 1) {
            st->winner = lane;
            st->Jres = Jt; st->exp_res = expected; st->z_res = z;
        }
    }
}

}  // namespace tob
