/*
 * trajopt_b200.h — C ABI of the B200-native batched iLQR / AL-iLQR / ALTRO engine.
 *
 * This header is the drop-in boundary for the hot path of TrajectoryOptimization.jl v0.1.1
 * (reference paths below are relative to the reference tree).  The reference has no FFI of its
 * own: its extension point is the Julia solver interface
 *     solve!(prob::Problem, opts::AbstractSolverOptions)            src/solvers.jl:91-94
 *     solve!(prob, opts::ALTROSolverOptions)                        src/solvers/altro/altro_methods.jl:2-53
 *     solve!(prob, solver::AugmentedLagrangianSolver)               src/solvers/augmented_lagrangian/augmented_lagrangian_methods.jl:2-31
 *     solve!(prob, solver::iLQRSolver)                              src/solvers/ilqr/ilqr_methods.jl:3-45
 * A Julia shim (julia/TrajOptB200.jl) marshals a `Problem` into the POD
 * descriptors below and `ccall`s these entry points; Python binds the same symbols with ctypes.
 *
 * Conventions: plain pointers and sizes only; all matrices are COLUMN-MAJOR (Julia layout);
 * all batch arrays are problem-major on the host ([b][k][i], i fastest); every function returns
 * 0 on success or a negative TO_ERR_* code; a message is available from to_last_error().
 * No exception crosses the ABI.  A handle is not re-entrant; distinct handles are independent.
 */
#ifndef TRAJOPT_B200_H
#define TRAJOPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- continuous dynamics models (reference: the dynamics/ directory, Dynamics submodule) ------------- */
#define TO_MODEL_DOUBLE_INTEGRATOR 0 /* dynamics/double_integrator.jl:1-4      n=2  m=1 */
#define TO_MODEL_PENDULUM          1 /* dynamics/pendulum.jl:3-12              n=2  m=1 */
#define TO_MODEL_CAR               2 /* dynamics/car.jl:3-8                    n=3  m=2 */
#define TO_MODEL_CARTPOLE          3 /* dynamics/cartpole.jl:9-36              n=4  m=1 */
#define TO_MODEL_QUADROTOR         4 /* dynamics/quadrotor.jl:10-71            n=13 m=4 */
#define TO_MODEL_ACROBOT           5 /* dynamics/acrobot.jl:6 (URDF, torques [0,1])  n=4 m=1 */
#define TO_MODEL_DOUBLEPENDULUM    6 /* dynamics/doublependulum.jl:7 (URDF)          n=4 m=2 */
#define TO_NUM_MODELS              7

/* ---- integrators (reference: src/integration.jl) ------------------------------------------- */
#define TO_INTEG_RK3      0 /* :149-158 */
#define TO_INTEG_RK4      1 /* :115-125 */
#define TO_INTEG_MIDPOINT 2 /* :26-33   */

/* ---- constraint rows.  One row = one scalar constraint c(x,u) (<= 0, or == 0 if equality).
 * z = [x; u] of the ORIGINAL problem (0-based index `var`).
 *   LINEAR      c = sign * (z[var] - a)      BoundConstraint rows (src/constraints.jl:212-227:
 *                                            x - x_max, u - u_max, x_min - x, u_min - u) and
 *                                            goal_constraint rows x - xf (:299-304)
 *   CIRCLE      c = -((x[0]-a)^2 + (x[1]-b)^2 - r^2)            src/utils.jl:140-144
 *   SPHERE      c = -((x[0]-a)^2 + (x[1]-b)^2 + (x[2]-c)^2 - r^2)  src/utils.jl:150-156
 * Rows are listed per knot class in the reference's stacking order (bounds: x_max,u_max,x_min,
 * u_min with infinite bounds trimmed).  `is_bound` marks rows that came from a BoundConstraint:
 * the ALTRO transforms move those after the other rows (src/constraint_sets.jl:135-150). */
#define TO_ROW_LINEAR 0
#define TO_ROW_CIRCLE 1
#define TO_ROW_SPHERE 2

typedef struct TOConstraintRow {
    int32_t kind;      /* TO_ROW_* */
    int32_t equality;  /* 1 = equality row, 0 = inequality row */
    int32_t var;       /* LINEAR: index into z=[x;u] (terminal rows: into x) */
    int32_t is_bound;  /* 1 if the row belongs to a BoundConstraint */
    double  sign;      /* LINEAR: +1 (upper bound / goal) or -1 (lower bound) */
    double  a, b, c, r;
} TOConstraintRow;

/* ---- problem descriptor: one "shape class" shared by the whole batch.
 * Mirrors Problem (src/problem.jl:37-47) with an LQRObjective / QuadraticCost objective
 * (src/objective.jl:102-114, src/cost.jl:112-131) and Constraints (src/constraint_sets.jl:157-206). */
typedef struct TOProblemDesc {
    int32_t model;       /* TO_MODEL_* */
    int32_t integrator;  /* TO_INTEG_* */
    int32_t n, m, N;     /* states, controls, knot points */
    int32_t reserved0;
    double  dt;          /* time step (initial time step for minimum-time problems) */
    double  tf;          /* final time; 0.0 selects the minimum-time transform (altro_methods.jl:111) */
    /* stage cost  (1/2 x'Qx + 1/2 u'Ru + q'x + r'u + c + u'Hx) * dt ; terminal 1/2 x'Qf x + qf'x + cf */
    const double *Q;   /* n*n */
    const double *R;   /* m*m */
    const double *H;   /* m*n, may be NULL (zeros) */
    const double *q;   /* n */
    const double *r;   /* m */
    double        c;
    const double *Qf;  /* n*n */
    const double *qf;  /* n */
    double        cf;
    /* constraints: each knot k (0-based) has a class; class j owns rows[class_row_start[j] .. class_row_start[j+1]).
     * The class of knot N-1 is evaluated as a TERMINAL set (function of x only). n_classes may be 0. */
    int32_t n_classes;
    int32_t reserved1;
    const int32_t *class_of_knot;    /* N entries, each in [0,n_classes) or -1 for "no constraints" */
    const int32_t *class_row_start;  /* n_classes+1 entries */
    const TOConstraintRow *rows;
} TOProblemDesc;

/* ---- options: the LIVE fields of the reference option structs ------------------------------ */
typedef struct TOiLQROptions {          /* src/solvers/ilqr/ilqr_solver.jl:7-81 */
    double  cost_tolerance;             /* 1e-4 */
    double  gradient_norm_tolerance;    /* 1e-5 */
    int32_t iterations;                 /* 300 */
    int32_t dJ_counter_limit;           /* 10 */
    int32_t square_root;                /* 0 */
    int32_t iterations_linesearch;      /* 20 */
    double  line_search_lower_bound;    /* 1e-8 */
    double  line_search_upper_bound;    /* 10 */
    double  bp_reg_increase_factor;     /* 1.6 */
    double  bp_reg_max;                 /* 1e8 (only warns in the reference) */
    double  bp_reg_min;                 /* 1e-8 */
    double  bp_reg_fp;                  /* 10 */
    double  max_cost_value;             /* 1e8 */
    double  max_state_value;            /* 1e8 */
    double  max_control_value;          /* 1e8 */
    int32_t bp_reg_type;                /* TO_REG_CONTROL (:control): Quu + rho I; TO_REG_STATE (:state): Quu + rho B'B, Qux + rho B'A
                                           (backward_pass.jl:38-46).  :state runs on the CTA-per-problem backward pass. */
    int32_t gradient_type;              /* TO_GRAD_TODOROV (default), _FEEDFORWARD, _L2, _LINF (ilqr_methods.jl:91-137) */
} TOiLQROptions;
#define TO_REG_CONTROL 0
#define TO_REG_STATE   1
#define TO_GRAD_TODOROV     0
#define TO_GRAD_FEEDFORWARD 1
#define TO_GRAD_L2          2
#define TO_GRAD_LINF        3

typedef struct TOALOptions {            /* src/solvers/augmented_lagrangian/augmented_lagrangian_solver.jl:8-66 */
    TOiLQROptions opts_uncon;
    double  cost_tolerance;                       /* 1e-4 */
    double  cost_tolerance_intermediate;          /* 1e-3 */
    double  gradient_norm_tolerance;              /* 1e-5 */
    double  gradient_norm_tolerance_intermediate; /* 1e-5 */
    double  constraint_tolerance;                 /* 1e-3 */
    int32_t iterations;                           /* 30 */
    int32_t kickout_max_penalty;                  /* 0 */
    double  dual_min, dual_max;                   /* -1e8, 1e8 */
    double  penalty_max;                          /* 1e8 */
    double  penalty_initial;                      /* 1 */
    double  penalty_scaling;                      /* 10 */
} TOALOptions;

typedef struct TOALTROOptions {         /* src/solvers/altro/altro_solver.jl:6-65 */
    TOALOptions opts_al;
    double  R_inf;                            /* 1 */
    int32_t dynamically_feasible_projection;  /* 1 */
    int32_t resolve_feasible_problem;         /* 1 */
    double  R_minimum_time;                   /* 1 */
    double  dt_max;                           /* 1 */
    double  dt_min;                           /* 1e-3 */
    /* projected-Newton polish after the AL solve (altro_methods.jl:6-14,31-39; src/solvers/direct/projected_newton.jl,
     * solve_type = :feasible).  With projected_newton != 0 the AL phase stops at projected_newton_tolerance (or, if that is
     * negative, at the maximum penalty), then every problem is projected onto dynamics + active constraints. */
    int32_t projected_newton;                 /* 0 */
    int32_t pn_n_steps;                       /* 1     opts_pn.n_steps (direct_solvers.jl:19) */
    double  projected_newton_tolerance;       /* 1e-3 */
    double  pn_feasibility_tolerance;         /* 1e-6  opts_pn.feasibility_tolerance (direct_solvers.jl:28) */
    double  pn_active_set_tolerance;          /* 1e-3  opts_pn.active_set_tolerance (direct_solvers.jl:25) */
} TOALTROOptions;

/* ---- per-problem result record (32 bytes; this is what the multi-GPU allgather exchanges) -- */
#define TO_STATUS_OK             0
#define TO_STATUS_COST_BLOWUP    1  /* J > max_cost_value in an inner solve (ilqr_methods.jl:25-28) */
#define TO_STATUS_COST_INCREASED 2  /* "Cost increased during Forward Pass" (forward_pass.jl:80-82); solve aborted */
#define TO_STATUS_NOT_PD_SQRT    4  /* stage Hessian not PD in cost_expansion_sqrt! (objective.jl:76-93); aborted */
#define TO_STATUS_MAX_OUTER      8  /* AL iterations exhausted with c_max >= constraint_tolerance */
#define TO_STATUS_TRACE_TRUNC   16  /* history buffers were too small; records dropped */
#define TO_STATUS_PN_FAILED     64  /* projected Newton: the reference would have thrown (a full step that does not reduce the
                                       violation hits `count += a`, projected_newton.jl:304; or S + rho I not positive definite);
                                       X, U keep the AL solution */
#define TO_STATUS_PN_SKIPPED   128  /* projected Newton: a knot has more active constraint rows than the device block factor
                                       holds (24); X, U keep the AL solution */
#define TO_STATUS_REG_DIVERGED  32  /* backward pass: the PD test still fails with a non-finite regularisation rho
                                       (Quu is NaN).  The reference only warns on bp_reg_max (ilqr_methods.jl:169-171)
                                       and would restart forever (backward_pass.jl:52-63); the solve is aborted instead */

typedef struct TOResult {
    double  J;                 /* last recorded cost (AL: stats[:cost][end]) */
    double  c_max;             /* last recorded max violation (0 for iLQR) */
    int32_t iterations_total;  /* iLQR: stats[:iterations]; AL: stats[:iterations_total] (incl. initial records) */
    int32_t iterations_outer;  /* AL: stats[:iterations] (incl. the initial record); iLQR: 0 */
    int32_t status;            /* TO_STATUS_* bit mask */
    int32_t steps;             /* number of iLQR step! calls actually executed (the throughput unit) */
} TOResult;

/* one inner-iteration record: the parity observables of ilqr_methods.jl:77-89 / forward_pass.jl:75-78 */
typedef struct TOIterRecord {
    double cost, dJ, gradient, expected, z, alpha, rho;
    int32_t outer;      /* 0-based outer iteration the record belongs to (0 for plain iLQR) */
    int32_t iter;       /* stats[:iterations] after the record (1 = the initial record) */
} TOIterRecord;

/* one outer-iteration record: augmented_lagrangian_methods.jl:79-97 */
typedef struct TOOuterRecord {
    double cost, c_max, penalty_max;
    int32_t iterations_inner;  /* stats[:iterations] of the inner solver at the record */
    int32_t pad;
} TOOuterRecord;

#define TO_ERR_INVALID   -1
#define TO_ERR_UNSUPPORTED -2
#define TO_ERR_CUDA      -3
#define TO_ERR_NOMEM     -4

typedef struct TOSolver *TOHandle;

/* defaults == the reference's @with_kw defaults */
void to_default_ilqr_options(TOiLQROptions *o);
void to_default_al_options(TOALOptions *o);
void to_default_altro_options(TOALTROOptions *o);

/* Create a solver for a batch of B problems of one shape on CUDA device `device`.
 * The descriptor and everything it points to is copied before return.
 * Replaces: Problem(...) construction + AbstractSolver(prob, opts) (ilqr_solver.jl:118-144,
 * augmented_lagrangian_solver.jl:120-140, altro_solver.jl:85-94). */
int to_create(const TOProblemDesc *desc, int32_t B, int32_t device, TOHandle *out);
void to_destroy(TOHandle h);
const char *to_last_error(TOHandle h); /* h may be NULL: error of the last failed to_create */

/* Per-problem data, HOST pointers.  x0: B*n.  U0: B*(N-1)*m (initial_controls!, problem.jl:149-150).
 * X0: B*N*n or NULL; NULL mirrors an all-NaN prob.X (feasible start), non-NULL selects the
 * infeasible-start transform in to_solve_altro (altro_methods.jl:102) and is the initial state
 * trajectory otherwise. */
int to_set_batch(TOHandle h, const double *x0, const double *U0, const double *X0);
/* Same, but the pointers are DEVICE pointers already in problem-major layout (no H2D copy).
 * ORDERING: the device-to-device copies are enqueued on the handle's own (non-blocking) stream, see to_stream();
 * nothing orders them after work the caller still has in flight on another stream.  Synchronise the producer
 * (or make the handle's stream wait on an event of it) before calling. */
int to_set_batch_device(TOHandle h, const double *x0, const double *U0, const double *X0);
/* MPC re-solve without leaving the device (north_star "MPC initial conditions"; the reference's own tools for it are
 * initial_controls!(prob, U), src/problem.jl:149-150, and solver reset!): the initial controls of the next solve become the
 * last solution shifted by `shift` knots, U0[k] = U[min(k + shift, N-2)]; the initial states are x0 (host, B x n) or, when x0 is
 * NULL, the states the last plan predicts, X[shift].  Workspaces, kernels and streams of the handle are reused. */
int to_warm_start_shift(TOHandle h, const double *x0, int32_t shift);

/* Trace capacity per problem (0 = no histories).  Must be set before solving. */
int to_set_trace(TOHandle h, int32_t inner_capacity, int32_t outer_capacity);

/* The three solver entry points (blocking: return after the device work has finished). */
int to_solve_ilqr(TOHandle h, const TOiLQROptions *o);   /* ilqr_methods.jl:3-45 */
int to_solve_al(TOHandle h, const TOALOptions *o);       /* augmented_lagrangian_methods.jl:2-36 */
int to_solve_altro(TOHandle h, const TOALTROOptions *o); /* altro_methods.jl:2-53 */
/* Variant that does not wait for the END of the device work: the result copies of to_get_* and to_sync() wait.
 * It is not a pure enqueue: the lockstep engine replays its tick until the device-side list of live problems is
 * empty, so the host thread polls that (pinned) counter while the device works and returns once the last
 * kernel of the solve has been enqueued. */
int to_solve_altro_async(TOHandle h, const TOALTROOptions *o);
int to_sync(TOHandle h);
/* device time (ms) of the last solve's kernel(s), measured with CUDA events on the handle's stream */
int to_last_kernel_ms(TOHandle h, float *ms);
/* number of kernel launches issued by the last solve */
int to_last_launch_count(TOHandle h, int32_t *count);

/* Outputs, HOST pointers (may be NULL to skip).  X: B*N*n, U: B*(N-1)*m in the ORIGINAL dims
 * (process_results!, altro_methods.jl:60-61).  dts: B*(N-1) time steps actually used
 * (dt, or U[k][end]^2 of the minimum-time problem, minimum_time.jl:74-82). */
int to_get_solution(TOHandle h, double *X, double *U, double *dts);
int to_get_results(TOHandle h, TOResult *results /* B */);
/* device pointer to the B result records (for an NCCL allgather without a host round trip) */
int to_results_device_ptr(TOHandle h, void **ptr);
/* histories: inner [B][inner_capacity], outer [B][outer_capacity]; counts per problem */
int to_get_trace(TOHandle h, TOIterRecord *inner, int32_t *n_inner, TOOuterRecord *outer, int32_t *n_outer);
/* Recorded only while a trace is enabled (to_set_trace with a non-zero capacity BEFORE the solve); otherwise
 * to_num_constraint_rows() reports 0 and to_get_duals() returns TO_ERR_INVALID.
 * final multipliers/penalties/active set of the LAST AL solve, packed per problem as the
 * concatenation over knots of that knot's rows: lambda/mu: B*P doubles, active: B*P bytes,
 * P = to_num_constraint_rows().  (augmented_lagrangian_solver.jl:96-110) */
int to_num_constraint_rows(TOHandle h, int32_t *P);
int to_get_duals(TOHandle h, double *lambda, double *mu, uint8_t *active);

/* The CUDA stream (cudaStream_t) every copy and kernel of this handle is enqueued on: lets a caller
 * bracket the work with its own CUDA events or order other work after it. */
int to_stream(TOHandle h, void **stream);
/* device-to-device copy of the B result records into caller-owned device memory (e.g. the send
 * buffer of an NCCL allgather); returns after the copy has completed. */
int to_copy_results_device(TOHandle h, void *dst_device);
/* Sum over the batch of the line-search trials the sequential reference semantics evaluate
 * (index of the accepted step size + 1 per iLQR iteration; 21 on a failed search) in the last solve:
 * the L of the roofline formulas. */
int to_last_linesearch_trials(TOHandle h, int64_t *total);
/* Measured register-resident FP64 FMA throughput of `device` in TFLOP/s (FMA = 2 flops): the
 * FP64-pipe roofline denominator (MEASURED_PEAKS.json has no FP64 entry). */
int to_measure_fp64_peak(int32_t device, double *tflops);

/* library / device introspection */
int to_device_count(void);
const char *to_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TRAJOPT_B200_H */
