// Plain C++ view of the engine: the POD structures passed to the kernels and the kernel registry.
// Included by capi.cu (host orchestration) and by engine.cuh (device code).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/trajopt_b200.h"

namespace tob {

// ------------------------------------------------------------------------------------------
// device-side problem description (built by capi.cu from TOProblemDesc + ALTRO transforms)
// ------------------------------------------------------------------------------------------
enum { DR_LIN = 0, DR_CIRCLE = 1, DR_SPHERE = 2, DR_MTEQ = 3 };
struct DevRow {
    int kind, eq, col, pad;
    double sign, a, b, c, r;
};

struct DevProblem {
    int N, Ptot;
    int nrows, pad0;  // entries of rows[]
    double dt, R_mt;
    const double *Q, *R, *H, *q, *r, *Qf, *qf;  // dims nq / mq of the wrapped QuadraticCost
    double c, cf;
    int q_diag, r_diag, h_zero, qf_diag;
    const DevRow* rows;
    const int* knot_row_begin;  // N entries: first row of knot k in rows[]
    const int* knot_row_count;  // N entries
    const int* knot_lam_off;    // N+1 entries: offset of knot k in the per-problem λ/μ arrays
    // knots whose rows are all of the one-entry kind (bounds, goal, slack equality): per column of [x;u] the (at most 4) rows
    // that touch it, so a lane evaluates only its own rows.  knot_cols[k] = offset of the knot's table in col_tab ((n+m) x 4
    // row indices relative to the knot's first row, ascending, -1 = none), or -1: evaluate row by row (lockstep.cuh expansion)
    const int* knot_cols;       // N entries
    const int* col_tab;
};

struct DevBatch {
    int B, n_out, m_out;   // output dims (original problem)
    const double *x0, *U0, *X0;  // [B][n0], [B][N-1][m0], [B][N][n0] or null
    double *X, *U, *dts;         // outputs [B][N][n_out], [B][N-1][m_out], [B][N-1]
    TOResult* res;
    TOIterRecord* inner;
    int* n_inner;
    int inner_cap;
    TOOuterRecord* outer;
    int* n_outer;
    int outer_cap;
    double *lam_out, *mu_out;
    uint8_t* act_out;
    // infeasible start + minimum time in one ALTRO solve (altro_methods.jl:98-124): the first solve exports sqrt(dt) of every knot
    // (controls, [B][N-1]) and the extra state ([B][N]); the slack-free minimum-time re-solve starts from them
    // (infeasible.jl:43-51).  Null otherwise.
    double *tau_out, *xtau_out;
    const double *tau_in, *xtau_in;
};

struct DevCtl {
    int mode;              // 0 = iLQR, 1 = augmented Lagrangian
    int altro_init;        // 1: apply the ALTRO initialisation (slack controls / sqrt(dt) augmentation)
    int projection_first;  // 1: X ← open-loop rollout of U from x0 before solving (projection!, Q9)
    int accumulate;        // 1: keep status/steps/trace counters from a previous kernel on this batch
    int write_solution;    // 1: write X/U/dts
    TOALOptions o;
    unsigned int* queue;
    double* ws;
    unsigned long long ws_stride;  // doubles per warp
    unsigned long long* ls_trials;  // batch-wide sum of sequential-equivalent line-search trials
    int* debug_flag;
    double* debug;         // optional first-iteration dump of problem 0
};

// ------------------------------------------------------------------------------------------
// lockstep engine (lockstep.cuh): per-problem state carried between ticks, tick lists
// ------------------------------------------------------------------------------------------
struct LsState {
    double J_prev, rho, drho, last_dJ, last_grad, last_cost;
    double fp_expected, fp_z, fp_alpha;
    double dV0, dV1;
    double Jres, exp_res, z_res;  // accepted trial
    double cost_tol, grad_tol;    // tolerances of the running inner solve
    double Jout, cmax;            // last outer record
    unsigned long long ls_count;
    int iterations, dJ_zero, steps, status;
    int outer_i;                  // 1-based AL iteration (0: plain iLQR)
    int al_iterations, al_total;
    int n_inner_rec, n_outer_rec;
    int winner;                   // accepted trial index, -1 = none yet
    int inner_i;                  // 1-based step index of the running inner solve
    int inner_ok;                 // result of the inner solve that just ended (0 = aborted)
    int bp_fail;                  // the backward pass of this tick aborted (TO_STATUS_REG_DIVERGED)
};

// counters: [0],[1] = sizes of the two active lists, [2],[3] = retry lists, [4] = outer list, [5] = restart list
struct LsCtl {
    int* list[2];
    int* retry[2];
    int* outer_list;
    int* restart_list;             // problems whose backward pass a bulk launch handed over to the CTA kernel (counter: counts[5])
    unsigned int* counts;
    LsState* st;
    double* ws;                    // per-PROBLEM workspaces
    unsigned long long ws_stride;  // doubles per problem
    double* cand;                  // candidate trajectories of the line search, [slot][X|U][element][cand_width step sizes]; null = off
    int cand_width;                // step sizes per launch group (8 in the bulk, 32 in tail mode)
    int cand_by_problem;           // 1: slot = problem id (bulk buffer), 0: slot = position in the active list (tail buffer)
    double* res_scratch;           // ls_resident_kernel: per-CTA cost scratch (res_scratch_doubles(N) each)
    double* split_cost;            // split line search (resident.cuh): [list slot][knot][stage cost x 32 | AL cost x 32]; null = off
    int* split_ok;                 // split line search: [list slot][32] rollout stayed in the state / control box
};

constexpr int LS_BP_INLINE_RESTARTS = 2;  // regularisation increases a bulk backward-pass launch serves itself before handing the problem over
enum { LS_PHASE_INIT = 0, LS_PHASE_JAC, LS_PHASE_BP, LS_PHASE_TRIAL, LS_PHASE_ACCEPT, LS_PHASE_OUTER, LS_PHASE_TRIAL_ALL, LS_PHASE_BP_SQRT, LS_PHASE_ACCEPT_TAIL, LS_PHASE_EXPAND, LS_PHASE_BP_CTA, LS_PHASE_RESIDENT,
       LS_PHASE_SPLIT_CHAIN, LS_PHASE_SPLIT_COST, LS_PHASE_SPLIT_PICK };
struct LsGrids {
    int init, jac, bp, trial, accept, outer;  // grid sizes (persistent, grid-stride)
    int bp_smem, bp_groups_per_block, trial_group;
    int bp_warps, bp_kernel_smem;  // block shape of ls_bp_kernel itself (bp_smem: the default shape, shared with ls_expand_kernel)
    int occ_jac, occ_bp, occ_trial;
    int jac_pc;  // partial directions per thread in the Jacobian kernel
    int jac_minb, trial_minb, bp_minb, trial_all_minb;
    int expand, bp_cta, bp_cta_smem, occ_bp_cta, bp_cta_minb;  // latency path of the backward pass (ls_expand_kernel + ls_bp_cta_kernel)
    int res_threads, res_smem, res_capacity, res_minb;  // CTA-per-problem resident kernel (resident.cuh): block size, dynamic smem, resident CTAs
    int tab_bytes;  // dynamic shared memory of the per-block copy of the knot tables / constraint rows  // __launch_bounds__ min-blocks variants (register caps)
};

// host-visible launcher table entry
struct KernelInfo {
    int model, integ, inf, mt;
    int n, m;
    size_t smem_bytes;
    unsigned long long (*ws_doubles)(int N, int Ptot);
    size_t (*debug_doubles)(int N);
    int (*max_blocks_per_sm)();
    void (*launch)(int grid, cudaStream_t st, const DevProblem& P, const DevBatch& B, const DevCtl& c);
    // lockstep engine
    unsigned long long (*ls_ws_doubles)(int N, int Ptot);
    int (*ls_setup)(int sm_count, int N, int nrows, LsGrids* g);
    void (*ls_launch)(int phase, const LsGrids& g, cudaStream_t st, const DevProblem& P, const DevBatch& B, const DevCtl& c,
                      const LsCtl& lc, int cur, int grp);
    // projected-Newton polish (pn.cuh): CTAs that can be resident (0 = not available for this variant) and their dynamic smem;
    // scratch doubles per CTA; launch over the whole batch
    int (*pn_setup)(int sm_count, int N, int nrows, int* slots, int* smem);
    unsigned long long (*pn_scratch_doubles)(int N, int Ptot);
    void (*pn_launch)(int grid, int smem, cudaStream_t st, const DevProblem& P, const DevBatch& B, const LsCtl& lc, int n_steps,
                      double feas_tol, double act_tol, double* scratch, unsigned long long stride);
};


// registry (registry.cu, generated by build.py from the instantiation list)
const KernelInfo* find_kernel(int model, int integ, int inf, int mt);
void model_dims(int model, int* n, int* m);
int measure_fp64_peak(int device, double* tflops);  // peak.cu

}  // namespace tob
