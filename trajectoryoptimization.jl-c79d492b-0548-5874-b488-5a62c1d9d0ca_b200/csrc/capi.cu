// C ABI of include/trajopt_b200.h: host-side orchestration (descriptor validation, the ALTRO
// problem transforms of altro_methods.jl:98-124 / infeasible.jl:2-34 / minimum_time.jl:2-37 expressed
// as descriptor rewrites, device memory, kernel launches).  All numerics run in engine.cuh kernels;
// nothing here computes on the CPU and there is no fallback path.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/trajopt_b200.h"
#include "engine_host.h"

using namespace tob;

namespace {

thread_local std::string g_create_error;  // per thread: to_create may be called from several host threads (one handle each)

#define CK_RET(s, call)                                                                            \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf_[512];                                                                        \
            snprintf(buf_, sizeof buf_, "CUDA error '%s' at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return (s)->fail(TO_ERR_CUDA, buf_);                                                   \
        }                                                                                          \
    } while (0)

struct HostRow { int kind, eq, col; double sign, a, b, c, r; };

// device buffers of one (cfg) variant of the problem
struct Variant {
    bool built = false;
    int inf = 0, mt = 0;
    const KernelInfo* ki = nullptr;
    DevProblem P{};
    std::vector<void*> allocs;
    int grid = 0;
    unsigned long long ws_stride = 0;
    double* ws = nullptr;
    double key[6] = {0, 0, 0, 0, 0, 0};  // ALTRO parameters the variant was built with
    // lockstep engine buffers
    bool ls_ready = false;
    LsCtl lc{};
    LsGrids grids{};
    int* ls_lists = nullptr;
    double* cand_alloc = nullptr;  // tail-mode candidate trajectories (lockstep): 32 per list slot
    unsigned int cand_slots = 0;
    double* cand_bulk = nullptr;   // bulk candidate trajectories: trial_group per problem
    double* res_scratch = nullptr; // resident kernel: cost scratch of res_slots CTAs
    double* split_cost = nullptr;  // split line search (tail mode, large constraint sets): knot costs of cand_slots problems
    int* split_ok = nullptr;
    unsigned int res_slots = 0;
    double* pn_scratch = nullptr;  // projected Newton: block factors and vectors of pn_slots CTAs
    int pn_slots = 0, pn_smem = 0;
    unsigned long long pn_stride = 0;
};

}  // namespace

struct TOSolver {
    int device = 0;
    int B = 0;
    // descriptor copy
    TOProblemDesc d{};
    std::vector<double> Q, R, H, q, r, Qf, qf;
    std::vector<int32_t> class_of_knot, class_row_start;
    std::vector<TOConstraintRow> rows;
    // device batch
    double *x0 = nullptr, *U0 = nullptr, *X0 = nullptr, *X = nullptr, *U = nullptr, *dts = nullptr;
    bool has_X0 = false, batch_set = false;
    TOResult* res = nullptr;
    TOIterRecord* inner = nullptr;
    TOOuterRecord* outer = nullptr;
    int *n_inner = nullptr, *n_outer = nullptr;
    int inner_cap = 0, outer_cap = 0;
    double *lam_out = nullptr, *mu_out = nullptr;
    uint8_t* act_out = nullptr;
    int dual_P = 0;
    unsigned int* queue = nullptr;
    double* debug = nullptr;
    size_t debug_doubles = 0;
    Variant var[4];  // 0 plain, 1 infeasible, 2 minimum time, 3 infeasible + minimum time
    double *tau = nullptr, *xtau = nullptr;  // sqrt(dt) controls / extra state handed from variant 3 to the minimum-time re-solve
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int launches = 0;
    int sm_count = 0;
    int blocks_per_sm_override = 0;
    int engine = 1;              // 0 = warp-persistent kernel (engine.cuh), 1 = lockstep phase kernels (lockstep.cuh)
    int ticks = 0;               // lockstep ticks enqueued by the last solve
    int phase_timing = 0;        // 1: time every phase kernel of the lockstep engine with CUDA events
    double phase_ms[5] = {0, 0, 0, 0, 0};  // device ms of the last solve: jac, bp, trial, accept, outer
    double resident_ms = 0;      // device ms of the resident kernel of the last solve (phase timing / tick log only)
    int resident_problems = 0;   // upper bound of the problems the resident kernel of the last solve took over
    long long lockstep_passes = 0;  // iLQR iterations of the last solve that ran in lockstep ticks (the rest ran in the resident kernel)
    cudaEvent_t ev_res0 = nullptr, ev_res1 = nullptr;
    unsigned int* h_counts = nullptr;  // pinned ring of active-list sizes read back from the device
    cudaEvent_t ring_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::string err;

    int fail(int code, const std::string& msg) {
        err = msg;
        return code;
    }
};

namespace {

int fail_create(int code, const std::string& msg) {
    g_create_error = msg;
    return code;
}

// "diagonal" / "zero" in the bit-exact sense the kernels rely on: the other entries are +0.0 (a -0.0 would give -0.0*dt)
bool is_pzero(double v) { return v == 0.0 && !std::signbit(v); }
bool is_diag(const std::vector<double>& M, int n) {
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++)
            if (i != j && !is_pzero(M[(size_t)j * n + i])) return false;
    return true;
}
bool is_zero(const std::vector<double>& M) {
    for (double v : M) if (!is_pzero(v)) return false;
    return true;
}

template <class T>
int upload(TOSolver* s, Variant& v, const std::vector<T>& h, const T** out) {
    T* p = nullptr;
    size_t bytes = std::max<size_t>(1, h.size()) * sizeof(T);
    if (cudaMalloc(&p, bytes) != cudaSuccess) return s->fail(TO_ERR_NOMEM, "cudaMalloc failed (problem data)");
    if (!h.empty() && cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess)
        return s->fail(TO_ERR_CUDA, "cudaMemcpy failed (problem data)");
    v.allocs.push_back(p);
    *out = p;
    return 0;
}

void free_variant(Variant& v) {
    for (void* p : v.allocs) cudaFree(p);
    v.allocs.clear();
    if (v.ws) cudaFree(v.ws);
    v.ws = nullptr;
    if (v.lc.ws) cudaFree(v.lc.ws);
    if (v.lc.st) cudaFree(v.lc.st);
    if (v.lc.counts) cudaFree(v.lc.counts);
    if (v.cand_alloc) cudaFree(v.cand_alloc);
    if (v.cand_bulk) cudaFree(v.cand_bulk);
    if (v.res_scratch) cudaFree(v.res_scratch);
    v.res_scratch = nullptr;
    if (v.split_cost) cudaFree(v.split_cost);
    if (v.split_ok) cudaFree(v.split_ok);
    v.split_cost = nullptr;
    v.split_ok = nullptr;
    if (v.pn_scratch) cudaFree(v.pn_scratch);
    v.pn_scratch = nullptr;
    v.pn_slots = 0;
    v.res_slots = 0;
    v.cand_bulk = nullptr;
    v.cand_alloc = nullptr;
    v.cand_slots = 0;
    if (v.ls_lists) cudaFree(v.ls_lists);
    v.lc = LsCtl{};
    v.ls_lists = nullptr;
    v.ls_ready = false;
    v.built = false;
}

// Build the (possibly ALTRO-transformed) device problem.  Row ordering follows the reference:
// infeasible: (non-bound rows, bounds, :infeasible)           constraint_sets.jl:135-150, infeasible.jl:19-29
// min time:   (non-bound rows, combined bound incl. sqrt(dt) bounds, :min_time_eq for 1<k<N)   minimum_time.jl:126-147
int build_variant(TOSolver* s, int which, const TOALTROOptions* ao) {
    Variant& v = s->var[which];
    const bool inf = (which == 1 || which == 3), mt = (which == 2 || which == 3);
    double key[6] = {inf ? ao->R_inf : 0.0, mt ? ao->R_minimum_time : 0.0, mt ? ao->dt_max : 0.0, mt ? ao->dt_min : 0.0, 0, 0};
    if (v.built && memcmp(key, v.key, sizeof key) == 0) return 0;
    free_variant(v);
    memcpy(v.key, key, sizeof key);
    v.inf = inf; v.mt = mt;
    const TOProblemDesc& D = s->d;
    v.ki = find_kernel(D.model, D.integrator, inf, mt);
    if (!v.ki) {
        char buf[256];
        snprintf(buf, sizeof buf, "no sm_100a kernel instantiated for model=%d integrator=%d infeasible=%d min_time=%d", D.model,
                 D.integrator, (int)inf, (int)mt);
        return s->fail(TO_ERR_UNSUPPORTED, buf);
    }
    const int n0 = D.n, m0 = D.m, N = D.N;
    const int nq = n0, mq = m0 + (inf ? n0 : 0);
    const int nbar = n0 + (mt ? 1 : 0), mbar = mq + (mt ? 1 : 0);
    std::vector<double> R((size_t)mq * mq, 0.0), rr(mq, 0.0), H((size_t)mq * nq, 0.0);
    for (int j = 0; j < m0; j++) for (int i = 0; i < m0; i++) R[(size_t)j * mq + i] = s->R[(size_t)j * m0 + i];
    for (int i = 0; i < m0; i++) rr[i] = s->r[i];
    for (int j = 0; j < n0; j++) for (int i = 0; i < m0; i++) H[(size_t)j * mq + i] = s->H[(size_t)j * m0 + i];
    if (inf) {
        const double rinf = (ao->R_inf * 1.0) / D.dt;  // infeasible.jl:11
        for (int i = 0; i < n0; i++) R[(size_t)(m0 + i) * mq + (m0 + i)] = rinf;
    }
    DevProblem& P = v.P;
    P.N = N; P.dt = D.dt; P.R_mt = mt ? ao->R_minimum_time : 0.0; P.c = D.c; P.cf = D.cf;
    P.q_diag = is_diag(s->Q, n0); P.r_diag = is_diag(R, mq); P.h_zero = is_zero(H); P.qf_diag = is_diag(s->Qf, n0);
    int rc;
    if ((rc = upload(s, v, s->Q, &P.Q))) return rc;
    if ((rc = upload(s, v, R, &P.R))) return rc;
    if ((rc = upload(s, v, H, &P.H))) return rc;
    if ((rc = upload(s, v, s->q, &P.q))) return rc;
    if ((rc = upload(s, v, rr, &P.r))) return rc;
    if ((rc = upload(s, v, s->Qf, &P.Qf))) return rc;
    if ((rc = upload(s, v, s->qf, &P.qf))) return rc;
    // rows per knot (expanded; identical knots share storage through the class table)
    std::vector<DevRow> rows;
    std::vector<int> kb(N), kc(N), lo(N + 1), kcols(N, -1), col_tab;
    // cache per (class, position-kind) to keep the row table small: kind 0 = first knot, 1 = interior, 2 = terminal
    struct Key { int cls, pos; int begin, count; int cols; };
    std::vector<Key> cache;
    int off = 0;
    for (int k = 0; k < N; k++) {
        const bool term = (k == N - 1);
        const int cls = (D.n_classes > 0) ? s->class_of_knot[k] : -1;
        const int pos = term ? 2 : ((k == 0) ? 0 : 1);
        int found = -1;
        for (size_t i = 0; i < cache.size(); i++) if (cache[i].cls == cls && cache[i].pos == pos) found = (int)i;
        if (found < 0) {
            std::vector<DevRow> nonb, bxmax, bumax, bxmin, bumin, out;
            if (cls >= 0)
                for (int i = s->class_row_start[cls]; i < s->class_row_start[cls + 1]; i++) {
                    const TOConstraintRow& cr = s->rows[i];
                    DevRow r{};
                    r.kind = cr.kind; r.eq = cr.equality != 0; r.sign = cr.sign; r.a = cr.a; r.b = cr.b; r.c = cr.c; r.r = cr.r;
                    r.col = cr.var;
                    bool isx = true;
                    if (cr.kind == TO_ROW_LINEAR) {
                        isx = cr.var < n0;
                        r.col = isx ? cr.var : (cr.var - n0) + nbar;
                    }
                    if (!inf && !mt) { out.push_back(r); continue; }
                    if (cr.is_bound && cr.kind == TO_ROW_LINEAR) {
                        if (cr.sign > 0) (isx ? bxmax : bumax).push_back(r);
                        else (isx ? bxmin : bumin).push_back(r);
                    } else {
                        nonb.push_back(r);
                    }
                }
            if (inf || mt) {
                out = nonb;
                auto lin = [&](int col, double sign, double a, int eq) {
                    DevRow r{};
                    r.kind = DR_LIN; r.eq = eq; r.col = col; r.sign = sign; r.a = a;
                    return r;
                };
                // Both transforms in one solve (altro_methods.jl:98-124): minimum_time_problem runs on the slack problem, so the
                // :infeasible rows come BEFORE the re-combined bound (minimum_time.jl:132-141), and combine(bnd, mt_bnd) appends the
                // sqrt(dt) bounds after the ORIGINAL m0 controls of a knot's BoundConstraint (constraints.jl:195-203, u[1:m0+1]):
                // they land on the first slack control, not on sqrt(dt).  A knot without a BoundConstraint gets bnd0 of the slack
                // problem's size and there the bounds do reach sqrt(dt) (minimum_time.jl:131-137).  Reproduced as written.
                const bool both = inf && mt;
                const bool knot_has_bound = !(bxmax.empty() && bumax.empty() && bxmin.empty() && bumin.empty());
                const int tau_col = (both && knot_has_bound) ? nbar + m0 : nbar + mbar - 1;
                if (both && !term)
                    for (int i = 0; i < n0; i++) out.push_back(lin(nbar + m0 + i, 1.0, 0.0, 1));
                out.insert(out.end(), bxmax.begin(), bxmax.end());
                if (!term) {
                    out.insert(out.end(), bumax.begin(), bumax.end());
                    if (mt) out.push_back(lin(tau_col, 1.0, std::sqrt(ao->dt_max), 0));
                }
                out.insert(out.end(), bxmin.begin(), bxmin.end());
                if (!term) {
                    out.insert(out.end(), bumin.begin(), bumin.end());
                    if (mt) out.push_back(lin(tau_col, -1.0, std::sqrt(ao->dt_min), 0));
                }
                if (inf && !both && !term)
                    for (int i = 0; i < n0; i++) out.push_back(lin(nbar + m0 + i, 1.0, 0.0, 1));
                if (mt && k > 0 && !term) {
                    DevRow r{};
                    r.kind = DR_MTEQ; r.eq = 1; r.sign = 1.0;
                    out.push_back(r);
                }
            }
            // per-column row lists of a set made of one-entry rows only (at most 4 rows per column)
            int cols_off = -1;
            {
                const int nz = nbar + mbar;
                std::vector<int> tab((size_t)nz * 4, -1), cnt(nz, 0);
                bool ok = !out.empty();
                for (size_t i = 0; i < out.size() && ok; i++) {
                    const DevRow& r = out[i];
                    if (r.kind != DR_LIN || r.col < 0 || r.col >= nz || cnt[r.col] >= 4) { ok = false; break; }
                    tab[(size_t)r.col * 4 + cnt[r.col]++] = (int)i;
                }
                if (ok) {
                    cols_off = (int)col_tab.size();
                    col_tab.insert(col_tab.end(), tab.begin(), tab.end());
                }
            }
            Key key2{cls, pos, (int)rows.size(), (int)out.size(), cols_off};
            rows.insert(rows.end(), out.begin(), out.end());
            cache.push_back(key2);
            found = (int)cache.size() - 1;
        }
        kb[k] = cache[found].begin;
        kc[k] = cache[found].count;
        kcols[k] = cache[found].cols;
        lo[k] = off;
        off += kc[k];
    }
    lo[N] = off;
    P.Ptot = off;
    P.nrows = (int)rows.size();
    if ((rc = upload(s, v, rows, &P.rows))) return rc;
    if ((rc = upload(s, v, kb, &P.knot_row_begin))) return rc;
    if ((rc = upload(s, v, kc, &P.knot_row_count))) return rc;
    if ((rc = upload(s, v, lo, &P.knot_lam_off))) return rc;
    if ((rc = upload(s, v, kcols, &P.knot_cols))) return rc;
    if ((rc = upload(s, v, col_tab, &P.col_tab))) return rc;
    v.built = true;
    return 0;
}

// workspace of the selected engine (allocated on first use)
int ensure_engine_buffers(TOSolver* s, Variant& v) {
    const int N = s->d.N;
    char buf[160];
    if (s->engine == 0) {
        if (v.ws) return 0;
        int per_sm = s->blocks_per_sm_override > 0 ? s->blocks_per_sm_override : v.ki->max_blocks_per_sm();
        if (per_sm < 1) return s->fail(TO_ERR_CUDA, "kernel cannot be resident (occupancy 0)");
        long long grid = (long long)s->sm_count * per_sm;
        if (grid > s->B) grid = s->B;
        v.grid = (int)grid;
        v.ws_stride = v.ki->ws_doubles(N, v.P.Ptot);
        size_t bytes = (size_t)v.grid * v.ws_stride * sizeof(double);
        if (cudaMalloc(&v.ws, bytes) != cudaSuccess) {
            snprintf(buf, sizeof buf, "cudaMalloc of %.1f MB workspace failed", bytes / 1048576.0);
            return s->fail(TO_ERR_NOMEM, buf);
        }
        return 0;
    }
    if (v.ls_ready) return 0;
    const size_t B = (size_t)s->B;
    v.lc.ws_stride = v.ki->ls_ws_doubles(N, v.P.Ptot);
    const size_t bytes = B * v.lc.ws_stride * sizeof(double);
    if (cudaMalloc(&v.lc.ws, bytes) != cudaSuccess) {
        cudaGetLastError();
        snprintf(buf, sizeof buf, "cudaMalloc of %.1f MB per-problem workspaces failed", bytes / 1048576.0);
        return s->fail(TO_ERR_NOMEM, buf);
    }
    if (cudaMalloc(&v.lc.st, B * sizeof(LsState)) != cudaSuccess || cudaMalloc(&v.ls_lists, 6 * B * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&v.lc.counts, 64) != cudaSuccess) {
        cudaGetLastError();
        const bool built = v.built;  // keep the problem tables, drop the partly allocated workspaces
        std::vector<void*> keep;
        keep.swap(v.allocs);
        free_variant(v);
        v.allocs.swap(keep);
        v.built = built;
        return s->fail(TO_ERR_NOMEM, "cudaMalloc failed (lockstep state)");
    }
    v.lc.list[0] = v.ls_lists; v.lc.list[1] = v.ls_lists + B;
    v.lc.retry[0] = v.ls_lists + 2 * B; v.lc.retry[1] = v.ls_lists + 3 * B;
    v.lc.outer_list = v.ls_lists + 4 * B;
    v.lc.restart_list = v.ls_lists + 5 * B;
    int rc = v.ki->ls_setup(s->sm_count, N, v.P.nrows, &v.grids);
    if (rc != 0) {
        snprintf(buf, sizeof buf, "lockstep kernel setup failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError()));
        return s->fail(TO_ERR_CUDA, buf);
    }
    if (s->blocks_per_sm_override > 0) {
        v.grids.jac = s->sm_count * s->blocks_per_sm_override;
        v.grids.bp = s->sm_count * std::min(s->blocks_per_sm_override, std::max(1, v.grids.occ_bp));
        v.grids.trial = s->sm_count * s->blocks_per_sm_override;
    }
    // tail mode keeps all 32 candidate trajectories of up to `cand_slots` live problems
    {
        // 32 resident problem-warps per SM: below that a tick is latency-bound and one launch with all step sizes beats up to three
        // launches of 8 (measured, profiles/r02m: car_escape 16,384 problems 14.3 / 13.4 / 12.5 s at 8 / 16 / 32 per SM; the
        // quadrotor is indifferent) -- as long as the candidates fit in a sixteenth of the free device memory
        unsigned int slots = (unsigned int)s->sm_count * 32u;
        const size_t per = (size_t)(N * v.ki->n + (N - 1) * v.ki->m) * 32;  // doubles per problem for 32 candidates (cand_span)
        {
            size_t free_b = 0, total_b = 0;
            cudaMemGetInfo(&free_b, &total_b);
            const size_t fit = (free_b / 16) / (per * sizeof(double));
            if ((size_t)slots > fit) slots = (unsigned int)std::max<size_t>(fit, (size_t)s->sm_count * 2);
        }
        if (const char* env = getenv("TRAJOPT_B200_TAIL_THRESHOLD")) slots = (unsigned int)atoi(env);  // 0 disables tail mode
        if (slots > (unsigned int)s->B) slots = (unsigned int)s->B;
        if (slots > 0 && cudaMalloc(&v.cand_alloc, (size_t)slots * per * sizeof(double)) == cudaSuccess) v.cand_slots = slots;
        else { cudaGetLastError(); v.cand_alloc = nullptr; v.cand_slots = 0; }
        // bulk: the G candidates of every problem (skipped if it would not leave a quarter of the device memory free)
        // (4-component chunks per knot: cand_chunk_span in lockstep.cuh)
        const size_t bulk_per = ((size_t)N * (size_t)((v.ki->n + 3) / 4) + (size_t)(N - 1) * (size_t)((v.ki->m + 3) / 4)) * 4 * (size_t)v.grids.trial_group;
        const size_t bulk_bytes = B * bulk_per * sizeof(double);
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const char* env = getenv("TRAJOPT_B200_BULK_CANDIDATES");
        const bool want = !(env && env[0] == '0');
        if (want && bulk_bytes + total_b / 4 < free_b && cudaMalloc(&v.cand_bulk, bulk_bytes) != cudaSuccess) {
            cudaGetLastError();
            v.cand_bulk = nullptr;
        }
    }
    // resident kernel (one CTA per problem, runs to completion): as many slots as CTAs can be resident, at most the candidate slots
    {
        unsigned int slots = std::min<unsigned int>((unsigned int)std::max(0, v.grids.res_capacity), v.cand_slots);
        if (const char* env = getenv("TRAJOPT_B200_RESIDENT_THRESHOLD")) slots = std::min<unsigned int>(slots, (unsigned int)atoi(env));  // 0 disables
        const size_t per = 2 * (size_t)N * 32;  // res_scratch_doubles(N)
        // scratch for up to 4 launch waves of CTAs (a CTA beyond the resident capacity starts when an earlier one ends)
        const size_t scratch_slots = std::min<size_t>((size_t)slots * 4, (size_t)v.cand_slots);
        if (slots > 0 && cudaMalloc(&v.res_scratch, scratch_slots * per * sizeof(double)) == cudaSuccess) v.res_slots = slots;
        else { cudaGetLastError(); v.res_scratch = nullptr; v.res_slots = 0; }
        v.lc.res_scratch = v.res_scratch;
    }
    v.ls_ready = true;
    return 0;
}

// the lockstep engine: init kernel, then replay the tick (jac, bp, trial groups, accept, outer)
// until the active list is empty.  The host learns the list size from an asynchronous read-back
// that trails the enqueued work by LAG ticks, so the stream never drains between ticks.
int run_lockstep_impl(TOSolver* s, Variant& v, const DevBatch& Bt, const DevCtl& c);
int run_lockstep(TOSolver* s, Variant& v, const DevBatch& Bt, const DevCtl& c) {
    const int rc = run_lockstep_impl(s, v, Bt, c);
    if (rc == 0 && s->phase_timing) {  // diagnostics only: costs a stream synchronisation
        unsigned long long passes = 0;
        if (cudaStreamSynchronize(s->stream) == cudaSuccess && cudaMemcpy(&passes, v.lc.counts + 8, 8, cudaMemcpyDeviceToHost) == cudaSuccess)
            s->lockstep_passes += (long long)passes;
    }
    return rc;
}
int run_lockstep_impl(TOSolver* s, Variant& v, const DevBatch& Bt, const DevCtl& c) {
    const int LAG = 4, RING = 8;
    cudaStream_t st = s->stream;
    CK_RET(s, cudaMemsetAsync(v.lc.counts, 0, 64, st));
    CK_RET(s, cudaMemsetAsync(v.lc.st, 0, (size_t)s->B * sizeof(LsState), st));
    v.ki->ls_launch(LS_PHASE_INIT, v.grids, st, v.P, Bt, c, v.lc, 0, 0);
    CK_RET(s, cudaGetLastError());
    s->launches += 1;
    const int ntrial = c.o.opts_uncon.iterations_linesearch + 1;
    const int G = v.grids.trial_group;
    const int ngroups = (ntrial + G - 1) / G;
    // worst case number of ticks: every AL iteration runs every allowed inner step
    const long long max_ticks = (long long)std::max(1, c.o.iterations) * (long long)std::max(1, c.o.opts_uncon.iterations) + 8;
    // tail mode: with few live problems a tick is latency-bound, so every step size is tried in ONE launch
    // (a warp per problem) instead of G at a time
    const unsigned int tail_threshold = v.cand_slots;  // 0: tail mode off
    unsigned int known_active = (unsigned int)s->B;  // upper bound (the list only shrinks), refreshed LAG ticks late
    // backward pass: below this many live problems the CTA-per-problem kernel (latency path) replaces the lane-group kernel
    // (measured on the quadrotor, profiles/r01q, r01r: 4,096 beats 1,184 and 296 -- the regularisation-restart chains of single
    // problems, which stall a whole launch of the lane-group kernel, run 5x faster on the latency path)
    unsigned int cta_threshold = 4096u;
    bool defer_restarts = true;
    if (const char* env = getenv("TRAJOPT_B200_BP_DEFER_RESTARTS")) defer_restarts = (env[0] != '0');
    int inline_restarts = LS_BP_INLINE_RESTARTS;
    if (const char* env = getenv("TRAJOPT_B200_BP_INLINE_RESTARTS")) inline_restarts = std::max(0, std::min(64, atoi(env)));
    if (const char* env = getenv("TRAJOPT_B200_BP_CTA_THRESHOLD")) cta_threshold = (unsigned int)strtoul(env, nullptr, 10);
    // square-root pass: warp per problem once the live problems fit one wave of warps (sqrt_bp.cuh) -- for the models whose
    // matrices do not fit the registers of one thread (measured, profiles/r02x_sqrt_pass_timing.log: quadrotor 9.6x faster than
    // thread per problem, whose frame is local memory; the small models keep everything in registers and are 1.4-1.6x SLOWER
    // through shared memory, so they stay on the thread kernel)
    unsigned int sqrt_warp_threshold = (v.ki->n > 6 || v.ki->m > 6) ? (unsigned int)s->sm_count * 16u : 0u;
    if (const char* env = getenv("TRAJOPT_B200_SQRT_WARP_THRESHOLD")) sqrt_warp_threshold = (unsigned int)strtoul(env, nullptr, 10);
    // diagnostics: TRAJOPT_B200_TICK_LOG=<file> records one event per tick and writes "tick ms active" lines
    const char* tick_log = getenv("TRAJOPT_B200_TICK_LOG");
    const bool collect = tick_log || s->phase_timing;  // one CUDA event per phase (or per tick) on the solve stream
    const bool phase_log = (tick_log && getenv("TRAJOPT_B200_TICK_DETAIL")) || s->phase_timing;
    std::vector<cudaEvent_t> tick_ev;
    std::vector<unsigned int> tick_active;
    auto dump_ticks = [&]() {
        if (!collect || tick_ev.size() < 2) return;
        cudaStreamSynchronize(st);
        if (s->phase_timing) {
            // accumulate per-phase device time of this run: [jac, bp, trial, accept, outer]
            for (size_t i = 1; i + 4 < tick_ev.size(); i += 5)
                for (size_t q = 0; q < 5; q++) {
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, tick_ev[i + q - 1], tick_ev[i + q]);
                    s->phase_ms[q] += ms;
                }
        }
        FILE* f = tick_log ? fopen(tick_log, "a") : nullptr;
        if (f) {
            const size_t per = phase_log ? 5 : 1;  // events per tick: [jac, bp, trials, accept,] outer
            fprintf(f, "# solve B=%d ticks=%zu%s\n", s->B, (tick_ev.size() - 1) / per, phase_log ? " columns: tick jac bp trial accept outer active" : "");
            for (size_t i = 1; i + per - 1 < tick_ev.size(); i += per) {
                const size_t tk = (i - 1) / per;
                fprintf(f, "%zu", tk);
                for (size_t q = 0; q < per; q++) {
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, tick_ev[i + q - 1], tick_ev[i + q]);
                    fprintf(f, " %.4f", ms);
                }
                fprintf(f, " %u\n", tk < tick_active.size() ? tick_active[tk] : 0u);
            }
            fclose(f);
        }
        for (auto e : tick_ev) cudaEventDestroy(e);
    };
    if (collect) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        tick_ev.push_back(e);
    }
    auto mark = [&]() {
        if (!phase_log) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        tick_ev.push_back(e);
    };
    // the resident kernel takes over when the live problems fit its CTAs twice over: a CTA then serves two problems one after
    // the other, and nearly all of them end within a few iterations (the handful of stragglers sets the kernel's duration anyway)
    // (measured, profiles/r02l: car_escape 16,384 problems 17.2 -> 14.3 s with two waves -- its lockstep tick costs 8 ms while a
    // resident iteration costs 1 ms; the quadrotor's tail tick and resident iteration cost the same 0.65 ms, and two problems per
    // CTA only serialise them: 7.40 -> 7.58 s.  So: two waves for problems with large per-knot constraint sets, one otherwise.)
    int max_rows = 0;
    for (int k = 0; k < s->d.N; k++) {
        const int cls = (s->d.n_classes > 0) ? s->class_of_knot[k] : -1;
        if (cls >= 0) max_rows = std::max(max_rows, (int)(s->class_row_start[cls + 1] - s->class_row_start[cls]));
    }
    unsigned int res_waves = (max_rows > 32) ? 2u : 1u;
    if (const char* env = getenv("TRAJOPT_B200_RESIDENT_WAVES")) res_waves = (unsigned int)std::max(1, std::min(4, atoi(env)));
    if (c.o.opts_uncon.gradient_type != TO_GRAD_TODOROV && (!v.cand_bulk || !v.cand_alloc))
        return s->fail(TO_ERR_UNSUPPORTED, "gradient_type other than :todorov needs the candidate buffers of the line search "
                                           "(not enough device memory, or TRAJOPT_B200_BULK_CANDIDATES=0)");
    const unsigned int res_threshold = (!c.o.opts_uncon.square_root && ntrial <= 32 && v.cand_alloc) ? std::min(v.res_slots * res_waves, v.cand_slots) : 0u;
    // tail mode with large per-knot constraint sets: the cost comes off the state chain (three kernels instead of one, resident.cuh;
    // measured on car_escape, profiles/r03a: the line search of a tick 5-7 ms -> see DESIGN section 6)
    bool split_ls = (max_rows > 32) && v.cand_alloc && v.cand_slots > 0 && ntrial <= 32;
    if (const char* env = getenv("TRAJOPT_B200_SPLIT_LINESEARCH")) split_ls = split_ls && (env[0] != '0');
    if (split_ls && !v.split_cost) {
        const size_t cost_bytes = (size_t)v.cand_slots * (size_t)s->d.N * 64 * sizeof(double);
        if (cudaMalloc(&v.split_cost, cost_bytes) != cudaSuccess || cudaMalloc(&v.split_ok, (size_t)v.cand_slots * 32 * sizeof(int)) != cudaSuccess) {
            cudaGetLastError();
            if (v.split_cost) cudaFree(v.split_cost);
            v.split_cost = nullptr;
            v.split_ok = nullptr;
            split_ls = false;  // not enough memory: the one-kernel line search serves
        }
    }
    for (long long t = 0; t < max_ticks; t++) {
        const int cur = (int)(t & 1);
        if (known_active <= res_threshold) {
            // few live problems: each gets a CTA that runs it to completion (resident.cuh) -- the last launch of the solve
            LsCtl lcr = v.lc;
            lcr.cand = v.cand_alloc;
            lcr.cand_width = 32;
            lcr.cand_by_problem = 0;
            if (collect) { cudaEventRecord(s->ev_res0, st); }
            v.ki->ls_launch(LS_PHASE_RESIDENT, v.grids, st, v.P, Bt, c, lcr, cur, (int)known_active);
            CK_RET(s, cudaGetLastError());
            if (collect) { cudaEventRecord(s->ev_res1, st); }
            s->launches += 1;
            s->resident_problems = (int)known_active;
            dump_ticks();
            if (collect) {
                float ms = 0.f;
                cudaEventSynchronize(s->ev_res1);
                cudaEventElapsedTime(&ms, s->ev_res0, s->ev_res1);
                s->resident_ms = ms;
                if (tick_log) {
                    if (FILE* f = fopen(tick_log, "a")) {
                        fprintf(f, "# resident kernel: %.4f ms, took over <= %u problems after %d ticks\n", ms, known_active, s->ticks);
                        fclose(f);
                    }
                }
            }
            return 0;
        }
        v.ki->ls_launch(LS_PHASE_JAC, v.grids, st, v.P, Bt, c, v.lc, cur, 0);
        mark();
        // bp_reg_type = :state is implemented by the CTA-per-problem pass (and the resident kernel) only
        const bool bp_cta = !c.o.opts_uncon.square_root && (known_active <= cta_threshold || c.o.opts_uncon.bp_reg_type == TO_REG_STATE);
        if (bp_cta) {
            // latency path: knot-parallel expansion, then one CTA per problem for the recursion
            v.ki->ls_launch(LS_PHASE_EXPAND, v.grids, st, v.P, Bt, c, v.lc, cur, 0);
            v.ki->ls_launch(LS_PHASE_BP_CTA, v.grids, st, v.P, Bt, c, v.lc, cur, (int)std::min<unsigned int>(known_active, (unsigned int)v.grids.bp_cta));
            s->launches += 1;
        } else if (c.o.opts_uncon.square_root) {
            // thread per problem while the batch fills the machine, warp per problem (same bits, a shorter chain) below that
            v.ki->ls_launch(LS_PHASE_BP_SQRT, v.grids, st, v.P, Bt, c, v.lc, cur, (known_active <= sqrt_warp_threshold) ? (int)known_active : 0);
        } else if (defer_restarts) {
            // bulk: the lane-group kernel serves up to LS_BP_INLINE_RESTARTS regularisation increases per problem itself and
            // queues the rare long restart chains (restart list, zeroed by the Jacobian kernel) for the latency path
            v.ki->ls_launch(LS_PHASE_BP, v.grids, st, v.P, Bt, c, v.lc, cur | 2 | (inline_restarts << 4), 0);
            v.ki->ls_launch(LS_PHASE_EXPAND, v.grids, st, v.P, Bt, c, v.lc, cur | 4, 2 * s->sm_count);
            v.ki->ls_launch(LS_PHASE_BP_CTA, v.grids, st, v.P, Bt, c, v.lc, cur | 4, 0);
            s->launches += 2;
        } else {
            v.ki->ls_launch(LS_PHASE_BP, v.grids, st, v.P, Bt, c, v.lc, cur, 0);
        }
        mark();
        const bool tail = (known_active <= tail_threshold && ntrial <= 32);
        LsCtl lct = v.lc;
        lct.cand = tail ? v.cand_alloc : v.cand_bulk;
        lct.cand_width = tail ? 32 : v.grids.trial_group;
        lct.cand_by_problem = tail ? 0 : 1;
        const bool with_cand = (lct.cand != nullptr);
        if (tail && split_ls) {
            lct.split_cost = v.split_cost;
            lct.split_ok = v.split_ok;
            const unsigned long long items = (unsigned long long)known_active * (unsigned long long)s->d.N;  // one warp each, 8 per block
            v.ki->ls_launch(LS_PHASE_SPLIT_CHAIN, v.grids, st, v.P, Bt, c, lct, cur, 0);
            v.ki->ls_launch(LS_PHASE_SPLIT_COST, v.grids, st, v.P, Bt, c, lct, cur,
                            (int)std::max<unsigned long long>(1, std::min<unsigned long long>((items + 7) / 8, (unsigned long long)s->sm_count * 8)));
            v.ki->ls_launch(LS_PHASE_SPLIT_PICK, v.grids, st, v.P, Bt, c, lct, cur, 0);
            s->launches -= ngroups - 3;
        } else if (tail) {
            v.ki->ls_launch(LS_PHASE_TRIAL_ALL, v.grids, st, v.P, Bt, c, lct, cur, 0);
            s->launches -= ngroups - 1;
        } else {
            for (int g = 0; g < ngroups; g++) {
                if (g >= 2) CK_RET(s, cudaMemsetAsync(v.lc.counts + 2 + (g & 1), 0, sizeof(unsigned int), st));
                v.ki->ls_launch(LS_PHASE_TRIAL, v.grids, st, v.P, Bt, c, lct, cur, g);
            }
        }
        mark();
        v.ki->ls_launch(with_cand ? LS_PHASE_ACCEPT_TAIL : LS_PHASE_ACCEPT, v.grids, st, v.P, Bt, c, lct, cur, 0);
        mark();
        v.ki->ls_launch(LS_PHASE_OUTER, v.grids, st, v.P, Bt, c, v.lc, cur, 0);
        CK_RET(s, cudaGetLastError());
        s->launches += 4 + ngroups;
        s->ticks += 1;
        const int slot = (int)(t % RING);
        CK_RET(s, cudaMemcpyAsync(&s->h_counts[slot], v.lc.counts + (cur ^ 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
        CK_RET(s, cudaEventRecord(s->ring_ev[slot], st));
        if (collect) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, st);
            tick_ev.push_back(e);
        }
        if (t >= LAG) {
            const int old = (int)((t - LAG) % RING);
            CK_RET(s, cudaEventSynchronize(s->ring_ev[old]));
            known_active = s->h_counts[old];
            if (collect) tick_active.push_back(known_active);
            if (known_active == 0) { dump_ticks(); return 0; }
        }
    }
    dump_ticks();
    return 0;
}

bool constrained(const TOSolver* s) {
    const TOProblemDesc& D = s->d;
    if (D.n_classes <= 0) return false;
    for (int k = 0; k < D.N; k++) {
        int c = s->class_of_knot[k];
        if (c >= 0 && s->class_row_start[c + 1] > s->class_row_start[c]) return true;
    }
    return false;
}

// Multipliers / penalties / active set of the final iterate are exported only while a trace is enabled (to_set_trace with a
// non-zero capacity): they cost 17 bytes per constraint row and problem.  Whenever the buffers do not match the variant that is
// about to run they are released, so a kernel never sees a buffer sized for another row count.
int ensure_dual_buffers(TOSolver* s, int P) {
    if (s->inner_cap <= 0 && s->outer_cap <= 0) P = 0;
    if (P == s->dual_P && (P == 0 || s->lam_out)) return 0;
    if (s->lam_out) { cudaFree(s->lam_out); cudaFree(s->mu_out); cudaFree(s->act_out); }
    s->lam_out = s->mu_out = nullptr;
    s->act_out = nullptr;
    s->dual_P = P;
    if (P == 0) return 0;
    size_t cnt = (size_t)s->B * P;
    if (cudaMalloc(&s->lam_out, cnt * 8) != cudaSuccess || cudaMalloc(&s->mu_out, cnt * 8) != cudaSuccess ||
        cudaMalloc(&s->act_out, cnt) != cudaSuccess)
        return s->fail(TO_ERR_NOMEM, "cudaMalloc failed (dual buffers)");
    cudaMemsetAsync(s->lam_out, 0, cnt * 8, s->stream);
    cudaMemsetAsync(s->mu_out, 0, cnt * 8, s->stream);
    cudaMemsetAsync(s->act_out, 0, cnt, s->stream);
    return 0;
}

// enqueue one kernel over the whole batch
int launch(TOSolver* s, int which, int mode, const TOALOptions& alo, bool altro_init, bool projection_first, bool accumulate,
           const double* X0_in, const double* U0_in, bool tau_export = false, bool tau_import = false) {
    Variant& v = s->var[which];
    DevBatch Bt{};
    if (tau_export || tau_import) {
        if (!s->tau) {
            if (cudaMalloc(&s->tau, (size_t)s->B * (s->d.N - 1) * 8) != cudaSuccess || cudaMalloc(&s->xtau, (size_t)s->B * s->d.N * 8) != cudaSuccess)
                return s->fail(TO_ERR_NOMEM, "cudaMalloc failed (sqrt(dt) hand-over buffers)");
        }
        if (tau_export) { Bt.tau_out = s->tau; Bt.xtau_out = s->xtau; }
        if (tau_import) { Bt.tau_in = s->tau; Bt.xtau_in = s->xtau; }
    }
    Bt.B = s->B; Bt.n_out = s->d.n; Bt.m_out = s->d.m;
    Bt.x0 = s->x0; Bt.U0 = U0_in; Bt.X0 = X0_in;
    Bt.X = s->X; Bt.U = s->U; Bt.dts = s->dts; Bt.res = s->res;
    Bt.inner = s->inner; Bt.n_inner = s->n_inner; Bt.inner_cap = s->inner_cap;
    Bt.outer = s->outer; Bt.n_outer = s->n_outer; Bt.outer_cap = s->outer_cap;
    int rc = ensure_dual_buffers(s, mode == 1 ? v.P.Ptot : 0);
    if (rc) return rc;
    Bt.lam_out = (mode == 1) ? s->lam_out : nullptr; Bt.mu_out = Bt.lam_out ? s->mu_out : nullptr; Bt.act_out = Bt.lam_out ? s->act_out : nullptr;
    DevCtl c{};
    c.mode = mode; c.altro_init = altro_init; c.projection_first = projection_first; c.accumulate = accumulate;
    c.write_solution = 1;
    c.o = alo;
    c.queue = s->queue;
    c.ls_trials = (unsigned long long*)(s->queue + 2);
    c.ws = v.ws; c.ws_stride = v.ws_stride;
    c.debug_flag = nullptr;
    c.debug = nullptr;
    if (s->debug_doubles > 0 && !accumulate) {
        size_t need = v.ki->debug_doubles(s->d.N);
        if (need <= s->debug_doubles) c.debug = s->debug;
    }
    rc = ensure_engine_buffers(s, v);
    if (rc) return rc;
    c.ws = v.ws; c.ws_stride = v.ws_stride;
    if (s->engine == 1) {
        c.debug = (s->debug_doubles >= 16) ? s->debug : nullptr;  // lockstep: cycle profile of one backward-pass group
        return run_lockstep(s, v, Bt, c);
    }
    CK_RET(s, cudaMemsetAsync(s->queue, 0, sizeof(unsigned int), s->stream));
    v.ki->launch(v.grid, s->stream, v.P, Bt, c);
    CK_RET(s, cudaGetLastError());
    s->launches += 1;
    return 0;
}

int check_opts(TOSolver* s, const TOALOptions& o) {
    const TOiLQROptions& io = o.opts_uncon;
    if (io.iterations_linesearch < 0 || io.iterations_linesearch > 31)
        return s->fail(TO_ERR_UNSUPPORTED, "iterations_linesearch must be in [0,31] (one warp lane per step size)");
    if (io.square_root && s->engine != 1)
        return s->fail(TO_ERR_UNSUPPORTED, "the square-root backward pass runs on the lockstep engine only (unset TRAJOPT_B200_ENGINE)");
    if (!s->batch_set) return s->fail(TO_ERR_INVALID, "to_set_batch has not been called");
    if (io.bp_reg_type != TO_REG_CONTROL && io.bp_reg_type != TO_REG_STATE) return s->fail(TO_ERR_INVALID, "unknown bp_reg_type");
    if (io.gradient_type < TO_GRAD_TODOROV || io.gradient_type > TO_GRAD_LINF) return s->fail(TO_ERR_INVALID, "unknown gradient_type");
    if (io.bp_reg_type == TO_REG_STATE && io.square_root)
        return s->fail(TO_ERR_UNSUPPORTED, "bp_reg_type = :state with the square-root backward pass is not on the device path");
    return 0;
}

// projected-Newton polish of every problem of the batch (altro_methods.jl:31-39), after the AL solve of variant `which`
int launch_pn(TOSolver* s, int which, const TOALTROOptions& ao) {
    Variant& v = s->var[which];
    if (!v.pn_scratch) {
        int slots = 0, smem = 0;
        const int rc = v.ki->pn_setup(s->sm_count, s->d.N, v.P.nrows, &slots, &smem);
        if (rc != 0 || slots < 1) return s->fail(TO_ERR_UNSUPPORTED, "projected Newton is not available for this problem variant (minimum time)");
        slots = std::min(slots, s->B);
        v.pn_stride = v.ki->pn_scratch_doubles(s->d.N, v.P.Ptot);
        if (cudaMalloc(&v.pn_scratch, (size_t)slots * v.pn_stride * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            return s->fail(TO_ERR_NOMEM, "cudaMalloc failed (projected-Newton scratch)");
        }
        v.pn_slots = slots;
        v.pn_smem = smem;
    }
    DevBatch Bt{};
    Bt.B = s->B; Bt.n_out = s->d.n; Bt.m_out = s->d.m;
    Bt.x0 = s->x0; Bt.X = s->X; Bt.U = s->U; Bt.dts = s->dts; Bt.res = s->res;
    v.ki->pn_launch(v.pn_slots, v.pn_smem, s->stream, v.P, Bt, v.lc, ao.pn_n_steps, ao.pn_feasibility_tolerance, ao.pn_active_set_tolerance,
                    v.pn_scratch, v.pn_stride);
    CK_RET(s, cudaGetLastError());
    s->launches += 1;
    return 0;
}

int solve_common(TOSolver* s, int api_mode, const TOALTROOptions& ao, bool sync) {
    CK_RET(s, cudaSetDevice(s->device));
    int rc = check_opts(s, ao.opts_al);
    if (rc) return rc;
    TOALOptions alo = ao.opts_al;
    const bool pn = (api_mode == 2) && ao.projected_newton != 0;
    if (pn) {  // altro_methods.jl:6-14
        if (s->engine != 1) return s->fail(TO_ERR_UNSUPPORTED, "projected Newton runs on the lockstep engine only (unset TRAJOPT_B200_ENGINE)");
        if (s->d.tf == 0.0) return s->fail(TO_ERR_UNSUPPORTED, "projected Newton with minimum time: MinTimeCost has no hessian! in the reference");
        if (ao.pn_n_steps < 1) return s->fail(TO_ERR_INVALID, "pn_n_steps must be >= 1");
        if (ao.projected_newton_tolerance >= 0) alo.constraint_tolerance = ao.projected_newton_tolerance;
        else { alo.constraint_tolerance = 0; alo.kickout_max_penalty = 1; }
    }
    s->launches = 0;
    s->ticks = 0;
    for (double& v : s->phase_ms) v = 0.0;
    s->resident_ms = 0.0;
    s->resident_problems = 0;
    s->lockstep_passes = 0;
    const bool con = constrained(s);
    const double* X0_in = s->has_X0 ? s->X0 : nullptr;
    CK_RET(s, cudaMemsetAsync(s->queue, 0, 256, s->stream));
    CK_RET(s, cudaEventRecord(s->ev0, s->stream));
    if (api_mode == 0) {
        if ((rc = build_variant(s, 0, &ao))) return rc;
        if ((rc = launch(s, 0, 0, alo, false, false, false, X0_in, s->U0))) return rc;
    } else if (api_mode == 1) {
        if ((rc = build_variant(s, 0, &ao))) return rc;
        if ((rc = launch(s, 0, con ? 1 : 0, alo, false, false, false, X0_in, s->U0))) return rc;
    } else {
        const bool inf = s->has_X0;          // altro_methods.jl:102
        const bool mt = (s->d.tf == 0.0);    // altro_methods.jl:111
        const int which = inf ? (mt ? 3 : 1) : (mt ? 2 : 0);
        if ((rc = build_variant(s, which, &ao))) return rc;
        if ((rc = launch(s, which, 1, alo, true, false, false, X0_in, s->U0, inf && mt, false))) return rc;
        if (pn && (rc = launch_pn(s, which, ao))) return rc;
        if (inf && ao.resolve_feasible_problem) {
            // infeasible_to_feasible_problem + projection! + second AL solve (altro_methods.jl:67-78); with minimum time the
            // slack-free problem is the minimum-time problem again, started from the sqrt(dt) of the first solve (infeasible.jl:43-51)
            const int second = mt ? 2 : 0;
            if ((rc = build_variant(s, second, &ao))) return rc;
            if ((rc = launch(s, second, (con || mt) ? 1 : 0, alo, false, ao.dynamically_feasible_projection != 0, true, s->X, s->U, false, mt))) return rc;
        }
    }
    CK_RET(s, cudaEventRecord(s->ev1, s->stream));
    if (sync) CK_RET(s, cudaStreamSynchronize(s->stream));
    return 0;
}

}  // namespace

// ==========================================================================================
extern "C" {

void to_default_ilqr_options(TOiLQROptions* o) {
    o->cost_tolerance = 1e-4; o->gradient_norm_tolerance = 1e-5; o->iterations = 300; o->dJ_counter_limit = 10;
    o->square_root = 0; o->iterations_linesearch = 20; o->line_search_lower_bound = 1e-8; o->line_search_upper_bound = 10.0;
    o->bp_reg_increase_factor = 1.6; o->bp_reg_max = 1e8; o->bp_reg_min = 1e-8; o->bp_reg_fp = 10.0;
    o->max_cost_value = 1e8; o->max_state_value = 1e8; o->max_control_value = 1e8;
    o->bp_reg_type = TO_REG_CONTROL; o->gradient_type = TO_GRAD_TODOROV;
}
void to_default_al_options(TOALOptions* o) {
    to_default_ilqr_options(&o->opts_uncon);
    o->cost_tolerance = 1e-4; o->cost_tolerance_intermediate = 1e-3; o->gradient_norm_tolerance = 1e-5;
    o->gradient_norm_tolerance_intermediate = 1e-5; o->constraint_tolerance = 1e-3; o->iterations = 30;
    o->kickout_max_penalty = 0; o->dual_min = -1e8; o->dual_max = 1e8; o->penalty_max = 1e8; o->penalty_initial = 1.0;
    o->penalty_scaling = 10.0;
}
void to_default_altro_options(TOALTROOptions* o) {
    to_default_al_options(&o->opts_al);
    o->R_inf = 1.0; o->dynamically_feasible_projection = 1; o->resolve_feasible_problem = 1;
    o->R_minimum_time = 1.0; o->dt_max = 1.0; o->dt_min = 1e-3;
    o->projected_newton = 0; o->pn_n_steps = 1; o->projected_newton_tolerance = 1e-3;
    o->pn_feasibility_tolerance = 1e-6; o->pn_active_set_tolerance = 1e-3;
}

int to_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
const char* to_version(void) { return "trajopt_b200 0.1 (sm_100a, FP64, warp-resident iLQR/AL/ALTRO)"; }

const char* to_last_error(TOHandle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int to_create(const TOProblemDesc* desc, int32_t B, int32_t device, TOHandle* out) {
    if (!desc || !out || B <= 0) return fail_create(TO_ERR_INVALID, "to_create: bad arguments");
    if (desc->model < 0 || desc->model >= TO_NUM_MODELS) return fail_create(TO_ERR_UNSUPPORTED, "unknown model id");
    int mn, mm;
    model_dims(desc->model, &mn, &mm);
    if (desc->n != mn || desc->m != mm) return fail_create(TO_ERR_INVALID, "n/m do not match the model");
    if (desc->N < 2) return fail_create(TO_ERR_INVALID, "N must be >= 2");
    if (!(desc->dt > 0)) return fail_create(TO_ERR_INVALID, "dt must be strictly positive");  // problem.jl:66-68
    if (!desc->Q || !desc->R || !desc->q || !desc->r || !desc->Qf || !desc->qf)
        return fail_create(TO_ERR_INVALID, "to_create: Q, R, q, r, Qf, qf must not be NULL (only H may be)");
    if (desc->n_classes > 0) {
        if (!desc->class_of_knot || !desc->class_row_start) return fail_create(TO_ERR_INVALID, "to_create: class tables must not be NULL");
        if (desc->class_row_start[0] != 0) return fail_create(TO_ERR_INVALID, "class_row_start must begin at 0");
        for (int c = 0; c < desc->n_classes; c++)
            if (desc->class_row_start[c + 1] < desc->class_row_start[c]) return fail_create(TO_ERR_INVALID, "class_row_start must be non-decreasing");
        if (desc->class_row_start[desc->n_classes] > 0 && !desc->rows) return fail_create(TO_ERR_INVALID, "to_create: rows must not be NULL");
    }
    if (!find_kernel(desc->model, desc->integrator, 0, 0)) return fail_create(TO_ERR_UNSUPPORTED, "no kernel for this model/integrator");
    int ndev = to_device_count();
    if (ndev <= 0) return fail_create(TO_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail_create(TO_ERR_INVALID, "device index out of range");
    if (cudaSetDevice(device) != cudaSuccess) return fail_create(TO_ERR_CUDA, "cudaSetDevice failed");
    TOSolver* s = new TOSolver();
    s->device = device;
    s->B = B;
    s->d = *desc;
    const int n = desc->n, m = desc->m, N = desc->N;
    s->Q.assign(desc->Q, desc->Q + (size_t)n * n);
    s->R.assign(desc->R, desc->R + (size_t)m * m);
    if (desc->H) s->H.assign(desc->H, desc->H + (size_t)m * n); else s->H.assign((size_t)m * n, 0.0);
    s->q.assign(desc->q, desc->q + n);
    s->r.assign(desc->r, desc->r + m);
    s->Qf.assign(desc->Qf, desc->Qf + (size_t)n * n);
    s->qf.assign(desc->qf, desc->qf + n);
    if (desc->n_classes > 0) {
        s->class_of_knot.assign(desc->class_of_knot, desc->class_of_knot + N);
        s->class_row_start.assign(desc->class_row_start, desc->class_row_start + desc->n_classes + 1);
        s->rows.assign(desc->rows, desc->rows + s->class_row_start[desc->n_classes]);
        for (int k = 0; k < N; k++)
            if (s->class_of_knot[k] >= desc->n_classes) { delete s; return fail_create(TO_ERR_INVALID, "class_of_knot out of range"); }
        for (auto& r : s->rows) {
            if (r.kind < 0 || r.kind > TO_ROW_SPHERE) { delete s; return fail_create(TO_ERR_INVALID, "unknown constraint row kind"); }
            if (r.kind == TO_ROW_LINEAR && (r.var < 0 || r.var >= n + m)) { delete s; return fail_create(TO_ERR_INVALID, "constraint row var out of range"); }
            if (r.kind == TO_ROW_SPHERE && n < 3) { delete s; return fail_create(TO_ERR_INVALID, "sphere rows need n >= 3"); }
        }
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    s->sm_count = prop.multiProcessorCount;
    const char* env = getenv("TRAJOPT_B200_BLOCKS_PER_SM");
    if (env) s->blocks_per_sm_override = atoi(env);
    env = getenv("TRAJOPT_B200_ENGINE");  // "persistent" selects the single-kernel engine (cross-check / tiny batches)
    if (env && (env[0] == 'p' || env[0] == '0')) s->engine = 0;
    bool ok = true;
    ok &= cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok &= cudaEventCreate(&s->ev0) == cudaSuccess && cudaEventCreate(&s->ev1) == cudaSuccess;
    ok &= cudaEventCreate(&s->ev_res0) == cudaSuccess && cudaEventCreate(&s->ev_res1) == cudaSuccess;
    ok &= cudaMalloc(&s->x0, (size_t)B * n * 8) == cudaSuccess;
    ok &= cudaMalloc(&s->U0, (size_t)B * (N - 1) * m * 8) == cudaSuccess;
    ok &= cudaMalloc(&s->X, (size_t)B * N * n * 8) == cudaSuccess;
    ok &= cudaMalloc(&s->U, (size_t)B * (N - 1) * m * 8) == cudaSuccess;
    ok &= cudaMalloc(&s->dts, (size_t)B * (N - 1) * 8) == cudaSuccess;
    ok &= cudaMalloc(&s->res, (size_t)B * sizeof(TOResult)) == cudaSuccess;
    ok &= cudaMalloc(&s->queue, 256) == cudaSuccess;
    ok &= cudaMallocHost(&s->h_counts, 8 * sizeof(unsigned int)) == cudaSuccess;
    for (auto& e : s->ring_ev) ok &= cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { to_destroy(s); return fail_create(TO_ERR_NOMEM, "device allocation failed in to_create"); }
    cudaMemset(s->res, 0, (size_t)B * sizeof(TOResult));
    *out = s;
    return 0;
}

void to_destroy(TOHandle s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (auto& v : s->var) free_variant(v);
    void* ptrs[] = {s->x0, s->U0, s->X0, s->X, s->U, s->dts, s->res, s->inner, s->outer, s->n_inner, s->n_outer,
                    s->lam_out, s->mu_out, s->act_out, s->queue, s->debug, s->tau, s->xtau};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (s->h_counts) cudaFreeHost(s->h_counts);
    for (auto& e : s->ring_ev) if (e) cudaEventDestroy(e);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->ev_res0) cudaEventDestroy(s->ev_res0);
    if (s->ev_res1) cudaEventDestroy(s->ev_res1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

static int set_batch_impl(TOHandle s, const double* x0, const double* U0, const double* X0, cudaMemcpyKind kind) {
    if (!s || !x0 || !U0) return s ? s->fail(TO_ERR_INVALID, "to_set_batch: null pointer") : TO_ERR_INVALID;
    CK_RET(s, cudaSetDevice(s->device));
    const int n = s->d.n, m = s->d.m, N = s->d.N, B = s->B;
    CK_RET(s, cudaMemcpyAsync(s->x0, x0, (size_t)B * n * 8, kind, s->stream));
    CK_RET(s, cudaMemcpyAsync(s->U0, U0, (size_t)B * (N - 1) * m * 8, kind, s->stream));
    if (X0) {
        if (!s->X0) CK_RET(s, cudaMalloc(&s->X0, (size_t)B * N * n * 8));
        CK_RET(s, cudaMemcpyAsync(s->X0, X0, (size_t)B * N * n * 8, kind, s->stream));
        s->has_X0 = true;
    } else {
        s->has_X0 = false;
    }
    s->batch_set = true;
    return 0;
}
int to_set_batch(TOHandle s, const double* x0, const double* U0, const double* X0) {
    return set_batch_impl(s, x0, U0, X0, cudaMemcpyHostToDevice);
}
int to_set_batch_device(TOHandle s, const double* x0, const double* U0, const double* X0) {
    return set_batch_impl(s, x0, U0, X0, cudaMemcpyDeviceToDevice);
}

// MPC warm start on the device: the next solve's initial controls are the last solution shifted by `shift` knots (the last
// control repeated), its initial state is x0 (host, B x n) or, if x0 is NULL, the state the last plan predicts at knot `shift`.
__global__ void warm_start_shift_kernel(double* U0, double* x0, const double* U, const double* X, int B, int N, int n, int m, int shift,
                                        int take_x) {
    const size_t per = (size_t)(N - 1) * m;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)B * per; e += (size_t)gridDim.x * blockDim.x) {
        const size_t b = e / per, r = e % per;
        const int k = (int)(r / m), i = (int)(r % m);
        const int ks = (k + shift < N - 1) ? (k + shift) : (N - 2);
        U0[e] = U[b * per + (size_t)ks * m + i];
    }
    if (take_x)
        for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)B * n; e += (size_t)gridDim.x * blockDim.x)
            x0[e] = X[(e / n) * (size_t)N * n + (size_t)shift * n + (e % n)];
}

int to_warm_start_shift(TOHandle s, const double* x0, int32_t shift) {
    if (!s) return TO_ERR_INVALID;
    if (!s->batch_set || !s->X || !s->U) return s->fail(TO_ERR_INVALID, "to_warm_start_shift: no previous solve on this handle");
    if (shift < 0 || shift > s->d.N - 1) return s->fail(TO_ERR_INVALID, "to_warm_start_shift: shift outside [0, N-1]");
    CK_RET(s, cudaSetDevice(s->device));
    const int n = s->d.n, m = s->d.m, N = s->d.N, B = s->B;
    if (x0) CK_RET(s, cudaMemcpyAsync(s->x0, x0, (size_t)B * n * 8, cudaMemcpyHostToDevice, s->stream));
    const size_t total = (size_t)B * (N - 1) * m;
    const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)s->sm_count * 16);
    warm_start_shift_kernel<<<grid, 256, 0, s->stream>>>(s->U0, s->x0, s->U, s->X, B, N, n, m, shift, x0 ? 0 : 1);
    CK_RET(s, cudaGetLastError());
    s->has_X0 = false;  // a warm start is a feasible start (the rollout of the shifted controls)
    return 0;
}

int to_set_trace(TOHandle s, int32_t inner_capacity, int32_t outer_capacity) {
    if (!s || inner_capacity < 0 || outer_capacity < 0) return TO_ERR_INVALID;
    CK_RET(s, cudaSetDevice(s->device));
    if (s->inner) { cudaFree(s->inner); cudaFree(s->n_inner); s->inner = nullptr; s->n_inner = nullptr; }
    if (s->outer) { cudaFree(s->outer); cudaFree(s->n_outer); s->outer = nullptr; s->n_outer = nullptr; }
    s->inner_cap = inner_capacity;
    s->outer_cap = outer_capacity;
    if (inner_capacity > 0) {
        CK_RET(s, cudaMalloc(&s->inner, (size_t)s->B * inner_capacity * sizeof(TOIterRecord)));
        CK_RET(s, cudaMalloc(&s->n_inner, (size_t)s->B * 4));
        CK_RET(s, cudaMemset(s->n_inner, 0, (size_t)s->B * 4));
    }
    if (outer_capacity > 0) {
        CK_RET(s, cudaMalloc(&s->outer, (size_t)s->B * outer_capacity * sizeof(TOOuterRecord)));
        CK_RET(s, cudaMalloc(&s->n_outer, (size_t)s->B * 4));
        CK_RET(s, cudaMemset(s->n_outer, 0, (size_t)s->B * 4));
    }
    return 0;
}

int to_solve_ilqr(TOHandle s, const TOiLQROptions* o) {
    if (!s || !o) return TO_ERR_INVALID;
    TOALTROOptions ao;
    to_default_altro_options(&ao);
    ao.opts_al.opts_uncon = *o;
    return solve_common(s, 0, ao, true);
}
int to_solve_al(TOHandle s, const TOALOptions* o) {
    if (!s || !o) return TO_ERR_INVALID;
    TOALTROOptions ao;
    to_default_altro_options(&ao);
    ao.opts_al = *o;
    return solve_common(s, 1, ao, true);
}
int to_solve_altro(TOHandle s, const TOALTROOptions* o) {
    if (!s || !o) return TO_ERR_INVALID;
    return solve_common(s, 2, *o, true);
}
int to_solve_altro_async(TOHandle s, const TOALTROOptions* o) {
    if (!s || !o) return TO_ERR_INVALID;
    return solve_common(s, 2, *o, false);
}
int to_sync(TOHandle s) {
    if (!s) return TO_ERR_INVALID;
    CK_RET(s, cudaSetDevice(s->device));
    CK_RET(s, cudaStreamSynchronize(s->stream));
    return 0;
}
int to_last_kernel_ms(TOHandle s, float* ms) {
    if (!s || !ms) return TO_ERR_INVALID;
    CK_RET(s, cudaEventSynchronize(s->ev1));
    CK_RET(s, cudaEventElapsedTime(ms, s->ev0, s->ev1));
    return 0;
}
int to_last_launch_count(TOHandle s, int32_t* count) {
    if (!s || !count) return TO_ERR_INVALID;
    *count = s->launches;
    return 0;
}

int to_get_solution(TOHandle s, double* X, double* U, double* dts) {
    if (!s) return TO_ERR_INVALID;
    CK_RET(s, cudaSetDevice(s->device));
    const int n = s->d.n, m = s->d.m, N = s->d.N, B = s->B;
    if (X) CK_RET(s, cudaMemcpyAsync(X, s->X, (size_t)B * N * n * 8, cudaMemcpyDeviceToHost, s->stream));
    if (U) CK_RET(s, cudaMemcpyAsync(U, s->U, (size_t)B * (N - 1) * m * 8, cudaMemcpyDeviceToHost, s->stream));
    if (dts) CK_RET(s, cudaMemcpyAsync(dts, s->dts, (size_t)B * (N - 1) * 8, cudaMemcpyDeviceToHost, s->stream));
    CK_RET(s, cudaStreamSynchronize(s->stream));
    return 0;
}
int to_get_results(TOHandle s, TOResult* results) {
    if (!s || !results) return TO_ERR_INVALID;
    CK_RET(s, cudaSetDevice(s->device));
    CK_RET(s, cudaMemcpyAsync(results, s->res, (size_t)s->B * sizeof(TOResult), cudaMemcpyDeviceToHost, s->stream));
    CK_RET(s, cudaStreamSynchronize(s->stream));
    return 0;
}
int to_results_device_ptr(TOHandle s, void** ptr) {
    if (!s || !ptr) return TO_ERR_INVALID;
    *ptr = s->res;
    return 0;
}
int to_get_trace(TOHandle s, TOIterRecord* inner, int32_t* n_inner, TOOuterRecord* outer, int32_t* n_outer) {
    if (!s) return TO_ERR_INVALID;
    CK_RET(s, cudaSetDevice(s->device));
    CK_RET(s, cudaStreamSynchronize(s->stream));
    if (s->inner_cap > 0 && inner && n_inner) {
        CK_RET(s, cudaMemcpy(inner, s->inner, (size_t)s->B * s->inner_cap * sizeof(TOIterRecord), cudaMemcpyDeviceToHost));
        CK_RET(s, cudaMemcpy(n_inner, s->n_inner, (size_t)s->B * 4, cudaMemcpyDeviceToHost));
    } else if (n_inner) {
        memset(n_inner, 0, (size_t)s->B * 4);
    }
    if (s->outer_cap > 0 && outer && n_outer) {
        CK_RET(s, cudaMemcpy(outer, s->outer, (size_t)s->B * s->outer_cap * sizeof(TOOuterRecord), cudaMemcpyDeviceToHost));
        CK_RET(s, cudaMemcpy(n_outer, s->n_outer, (size_t)s->B * 4, cudaMemcpyDeviceToHost));
    } else if (n_outer) {
        memset(n_outer, 0, (size_t)s->B * 4);
    }
    return 0;
}
int to_num_constraint_rows(TOHandle s, int32_t* P) {
    if (!s || !P) return TO_ERR_INVALID;
    *P = s->dual_P;
    return 0;
}
int to_get_duals(TOHandle s, double* lambda, double* mu, uint8_t* active) {
    if (!s) return TO_ERR_INVALID;
    CK_RET(s, cudaSetDevice(s->device));
    CK_RET(s, cudaStreamSynchronize(s->stream));
    if (s->dual_P == 0 || !s->lam_out)
        return s->fail(TO_ERR_INVALID, "to_get_duals: no multipliers were recorded (enable a trace with to_set_trace before an AL / ALTRO solve)");
    size_t cnt = (size_t)s->B * s->dual_P;
    if (lambda) CK_RET(s, cudaMemcpy(lambda, s->lam_out, cnt * 8, cudaMemcpyDeviceToHost));
    if (mu) CK_RET(s, cudaMemcpy(mu, s->mu_out, cnt * 8, cudaMemcpyDeviceToHost));
    if (active) CK_RET(s, cudaMemcpy(active, s->act_out, cnt, cudaMemcpyDeviceToHost));
    return 0;
}

int to_stream(TOHandle s, void** stream) {
    if (!s || !stream) return TO_ERR_INVALID;
    *stream = (void*)s->stream;
    return 0;
}
int to_copy_results_device(TOHandle s, void* dst) {
    if (!s || !dst) return TO_ERR_INVALID;
    CK_RET(s, cudaSetDevice(s->device));
    CK_RET(s, cudaMemcpyAsync(dst, s->res, (size_t)s->B * sizeof(TOResult), cudaMemcpyDeviceToDevice, s->stream));
    CK_RET(s, cudaStreamSynchronize(s->stream));
    return 0;
}
int to_last_linesearch_trials(TOHandle s, int64_t* total) {
    if (!s || !total) return TO_ERR_INVALID;
    CK_RET(s, cudaSetDevice(s->device));
    CK_RET(s, cudaStreamSynchronize(s->stream));
    unsigned long long v = 0;
    CK_RET(s, cudaMemcpy(&v, s->queue + 2, 8, cudaMemcpyDeviceToHost));
    *total = (int64_t)v;
    return 0;
}
int to_measure_fp64_peak(int32_t device, double* tflops) {
    if (!tflops) return TO_ERR_INVALID;
    return tob::measure_fp64_peak(device, tflops);
}

// ---- private diagnostics (csrc/debug_api.h; not part of the public header) -----------------
// enable a first-iteration dump of problem 0: [J0, dV0, dV1, Z(all knots), K|d(all knots)]
int to_debug_enable(TOHandle s, int32_t doubles) {
    if (!s) return TO_ERR_INVALID;
    CK_RET(s, cudaSetDevice(s->device));
    if (s->debug) { cudaFree(s->debug); s->debug = nullptr; }
    s->debug_doubles = doubles;
    if (doubles > 0) {
        CK_RET(s, cudaMalloc(&s->debug, (size_t)doubles * 8));
        CK_RET(s, cudaMemset(s->debug, 0, (size_t)doubles * 8));
    }
    return 0;
}
int to_debug_read(TOHandle s, double* out, int32_t doubles) {
    if (!s || !s->debug) return TO_ERR_INVALID;
    CK_RET(s, cudaStreamSynchronize(s->stream));
    CK_RET(s, cudaMemcpy(out, s->debug, (size_t)std::min<size_t>(doubles, s->debug_doubles) * 8, cudaMemcpyDeviceToHost));
    return 0;
}
int to_debug_grid(TOHandle s, int which, int32_t* grid, int32_t* smem, uint64_t* ws_doubles) {
    if (!s || which < 0 || which > 3 || !s->var[which].built) return TO_ERR_INVALID;
    if (s->engine == 1) {
        *grid = s->var[which].grids.bp;
        *smem = s->var[which].grids.bp_smem;
        *ws_doubles = s->var[which].lc.ws_stride;
        return 0;
    }
    *grid = s->var[which].grid;
    *smem = (int32_t)s->var[which].ki->smem_bytes;
    *ws_doubles = s->var[which].ws_stride;
    return 0;
}
// engine selection (0 = warp-persistent single kernel, 1 = lockstep phase kernels) and tick count of the last solve
int to_debug_set_engine(TOHandle s, int32_t engine) {
    if (!s || engine < 0 || engine > 1) return TO_ERR_INVALID;
    s->engine = engine;
    return 0;
}
// per-phase device time of the lockstep engine (diagnostics / roofline): enable, then read after a solve.
// ms[0..4] = jacobian, backward pass, line-search trials, accept, outer-loop kernels (summed over all ticks)
int to_debug_phase_timing(TOHandle s, int32_t enable) {
    if (!s) return TO_ERR_INVALID;
    s->phase_timing = enable ? 1 : 0;
    return 0;
}
int to_debug_phase_ms(TOHandle s, double* ms) {
    if (!s || !ms) return TO_ERR_INVALID;
    for (int i = 0; i < 5; i++) ms[i] = s->phase_ms[i];
    return 0;
}
// resident kernel of the last solve: device ms (needs phase timing or the tick log) and the problems it took over
int to_debug_resident(TOHandle s, double* ms, int32_t* problems, int64_t* lockstep_passes) {
    if (!s || !ms || !problems || !lockstep_passes) return TO_ERR_INVALID;
    *ms = s->resident_ms;
    *problems = s->resident_problems;
    *lockstep_passes = s->lockstep_passes;
    return 0;
}
int to_debug_ticks(TOHandle s, int32_t* ticks) {
    if (!s || !ticks) return TO_ERR_INVALID;
    *ticks = s->ticks;
    return 0;
}

}  // extern "C"
