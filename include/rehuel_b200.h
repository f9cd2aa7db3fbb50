/* rehuel_b200 -- C ABI of the B200-native batched ensemble integrator.
 *
 * Drop-in boundary for ONE path of Pakketeretet2/rehuel: the adaptive implicit Runge-Kutta
 * solver  irk::odeint -> irk_guts -> newton_solve_stages -> error estimator -> step controller
 * (reference irk.hpp:884-898, 509-869, 398-493).  The reference has no FFI layer of its own
 * (C_interface_test/main.c is an empty file), so every entry point below cites the C++ interface
 * it stands in for.  Plain pointers and sizes only; no C++, CUDA or torch types.
 *
 * Arrays are structure-of-arrays over the M ensemble members: component k of member i of an
 * [n][M] array lives at ptr[k*M + i].  All arithmetic is IEEE FP64.
 *
 * Thread safety: every function may be called concurrently from several host threads as long as
 * the output objects are distinct.  The library never prints unless options.out_interval > 0.
 */
#ifndef REHUEL_B200_H
#define REHUEL_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library error codes (return values).  Numerical status is per member, see below. ---- */
#define REHUEL_OK            0
#define REHUEL_EINVAL       (-1) /* bad argument, unknown problem/method, unsupported combination */
#define REHUEL_ECUDA        (-2) /* CUDA runtime error, text in rehuel_last_error() */
#define REHUEL_ENCCL        (-3) /* reserved: collective error (the library itself uses none) */
#define REHUEL_ENOMEM       (-4) /* host or device allocation failed */

/* ---- per-member integration status: odeint_status_codes, reference enums.hpp:103-127 ---- */
#define REHUEL_SUCCESS                    0
#define REHUEL_DT_TOO_SMALL               1
#define REHUEL_INTERNAL_SOLVE_FEW_ITERS   2
#define REHUEL_GENERAL_ERROR              4
#define REHUEL_DT_TOO_LARGE               8
#define REHUEL_INTERNAL_SOLVE_FAILURE     16
#define REHUEL_ERROR_LARGER_THAN_RELTOL   32
#define REHUEL_ERROR_LARGER_THAN_ABSTOL   64
#define REHUEL_ERROR_MAX_STEPS_EXCEEDED   128

/* ---- method ids: irk::rk_methods, reference enums.hpp:38-58 (same integers) ---- */
#define REHUEL_IMPLICIT_EULER   200
#define REHUEL_RADAU_IIA_32     210
#define REHUEL_RADAU_IIA_53     211
#define REHUEL_RADAU_IIA_95     212
#define REHUEL_RADAU_IIA_137    213
#define REHUEL_LOBATTO_IIIA_43  220
#define REHUEL_LOBATTO_IIIC_43  230
#define REHUEL_LOBATTO_IIIC_64  231
#define REHUEL_LOBATTO_IIIC_85  232

/* ---- built-in device functors (the functor.hpp:35-73 concept compiled for the GPU).
 *      Names follow examples/example_equations.cpp:475-494. ---- */
#define REHUEL_PROBLEM_ROBERTSON      1 /* n=3, params (k1,k2,k3); test_equations.hpp:150-176 with the constants as parameters */
#define REHUEL_PROBLEM_VAN_DER_POL    2 /* n=2, params (mu);       test_equations.hpp:91-115 */
#define REHUEL_PROBLEM_BRUSSELATOR_1D 3 /* n=2*Ng, params (alpha,A,B); 1-D method-of-lines Brusselator */
#define REHUEL_PROBLEM_BRUSSELATOR_1D_DENSE 4 /* the same system with its Jacobian treated as dense (general CTA-per-system kernel;
                                                 REHUEL_EINVAL for RADAU_IIA_137 at n = 64: its four matrices exceed a block's
                                                 shared memory -- the banded form of the problem runs every method) */
#define REHUEL_PROBLEM_EXPONENTIAL    5 /* n=1, params (lambda);   test_equations.hpp:37-59 */
#define REHUEL_PROBLEM_STIFF_EQ       6 /* n=2, no params;         test_equations.hpp:207-250 */
#define REHUEL_PROBLEM_BRUSSELATOR    7 /* n=2, params (a,b);      test_equations.hpp:120-145 */
#define REHUEL_PROBLEM_LORENZ         8 /* n=3, params (sigma,rho,beta); test_equations.hpp:519-543 */
#define REHUEL_PROBLEM_HARMONIC       9 /* n=2, params (w);        test_equations.hpp:63-84 */
#define REHUEL_PROBLEM_DIMER         10 /* n=2, params (rate);     test_equations.hpp:180-203 */
#define REHUEL_PROBLEM_KINETIC_4     11 /* n=4, params (b2,b3,b4); test_equations.hpp:484-517 */
#define REHUEL_PROBLEM_THREE_BODY    12 /* n=12, params (m1,m2,m3); test_equations.hpp:254-421, fun and jac as written there */

/* Solver options.  Replaces irk::solver_options (irk.hpp:95-124) : common_solver_options
 * (options.hpp:40-85), the newton::options it points to (newton.hpp:57-78; only tol, dx_delta,
 * maxit, refresh_jac are read by the irk path) and output_options (options.hpp:93-133). */
typedef struct rehuel_options {
	double rel_tol;                 /* options.hpp:50   default 1e-4 */
	double abs_tol;                 /* options.hpp:51   default 10*rel_tol */
	double max_dt;                  /* options.hpp:52   0 = no cap */
	long long max_steps;            /* options.hpp:53   <0 = unlimited; compared against ACCEPTED steps (irk.hpp:612) */
	int adaptive_step_size;         /* irk.hpp:103      default 1 */
	int out_interval;               /* options.hpp:55   >0: the host-buffer entry points keep the reference's per-attempt log
	                                   "step t dt err iters" (irk.hpp:575-577, 805-810) of member `trace_member`: one row per
	                                   attempt that reaches the error estimate and whose accepted-step count is a multiple of
	                                   out_interval (rehuel_result.trace; text form: rehuel_result_format_log) */
	double newton_tol;              /* newton.hpp:58    tol      -> Rtol, default 1e-4 */
	double newton_dx_delta;         /* newton.hpp:58    dx_delta -> xtol, default 1e-4 */
	int newton_maxit;               /* newton.hpp:58    default 5000 */
	int newton_refresh_jac;         /* newton.hpp:59    default 10 */
	int output_mode;                /* options.hpp:96-103 bit 0: store samples (STORE_IN_VECTORS); bit 1: WRITE_TO_FILE, the
	                                   samples of member `trace_member` are also written to `output_file` as "t y0 y1 ...\n"
	                                   lines (irk.hpp:837-843); default 1 */
	unsigned long long output_interval; /* options.hpp:104 every k-th ACCEPTED step is a sample; default 1 */
	int keep_state;                 /* NEW: != 0 makes the host-buffer entry points return the controller state at exit
	                                   (rehuel_result.state) so that rehuel_irk_ensemble_continue can carry on where
	                                   this call stopped.  Fills in irk::restore_state, an empty stub in the reference
	                                   (irk.cpp:786-803).  Default 0. */
	unsigned long long dense_samples; /* NEW: dense output.  > 0: every member's solution is also
	                                   sampled on the uniform grid t_k = dense_t0 + k*dense_dt, k = 0..dense_samples-1,
	                                   from the collocation polynomial of the accepted step that covers t_k
	                                   (b_j(theta) of irk::project_b, irk.cpp:731-747); only for methods that carry
	                                   b_interp (the Radau IIA family and LOBATTO_IIIC_85).  Default 0. */
	double dense_t0, dense_dt;
	int handout_order;              /* NEW (scheduling of the batched overload, no reference counterpart): order in which
	                                   the host-buffer entry points hand members to the GPU lanes.  REHUEL_ORDER_INDEX,
	                                   REHUEL_ORDER_LOCALITY (default) or REHUEL_ORDER_LOCALITY_DESC.  Results do not
	                                   depend on it. */
	int shard_mode;                 /* NEW: how rehuel_irk_ensemble(..., n_gpus > 1) deals members to the GPUs.  One
	                                   issuing host thread per device in every mode.
	                                   REHUEL_SHARD_DYNAMIC (default): [0, M) is cut into small chunks (>= 16 Ki members, up to
	                                   64 per GPU) that the devices claim from a shared atomic cursor as they finish, so a GPU
	                                   that gets cheap members takes more of them;
	                                   REHUEL_SHARD_CONTIGUOUS: GPU g owns members [M g/G, M (g+1)/G), 8 chunks each;
	                                   REHUEL_SHARD_INTERLEAVED: chunks of >= 128 Ki members go to the GPUs in turn.
	                                   Results do not depend on it (a member's result does not depend on where it runs). */
	unsigned long long trace_member; /* NEW: the member out_interval and WRITE_TO_FILE report on (the reference integrates one
	                                   system per call, so it has no such choice).  Default 0. */
	const char *output_file;        /* output_options::output_stream (options.hpp:107,119-123) as a path; NULL = none.  Opened for
	                                   writing by the call when output_mode bit 1 is set. */
	unsigned result_mask;           /* NEW: which per-member arrays of rehuel_result the host-buffer entry points copy back
	                                   (REHUEL_OUT_* bits).  y, status and the ensemble sums always come back; an array that
	                                   is not asked for is NULL in the result.  On success t_final == t1 for every member and
	                                   the counters are summed on the device (rehuel_result.sum), so a caller that only
	                                   wants the states moves 8n bytes per member instead of 8n + 56 (the status row is filled on the
	                                   host and only fetched for chunks in which a member failed).
	                                   Default REHUEL_OUT_ALL (every field of rk_output, irk.hpp:140-163). */
	int time_internals;             /* irk.hpp:106, 328-360: != 0 fills rehuel_result.phase_ms (device time of the host-to-device
	                                   copies, the hand-out ordering, the integration kernels and the device-to-host copies,
	                                   CUDA events; the reference times its six host phases with a wall clock).  Default 0. */
} rehuel_options;

#define REHUEL_OUT_T_FINAL   1u   /* rehuel_result.t_final */
#define REHUEL_OUT_ERR_LAST  2u   /* rehuel_result.err_last */
#define REHUEL_OUT_COUNTERS  4u   /* the nine per-member counter arrays */
#define REHUEL_OUT_ALL       7u

#define REHUEL_SHARD_CONTIGUOUS  0
#define REHUEL_SHARD_INTERLEAVED 1
#define REHUEL_SHARD_DYNAMIC     2

#define REHUEL_ORDER_INDEX          0 /* member i is the i-th to be handed out */
#define REHUEL_ORDER_LOCALITY       1 /* space-filling-curve order of the parameter vectors: members with nearby
                                         parameters share a warp and stay in lockstep (see rehuel_locality_order) */
#define REHUEL_ORDER_LOCALITY_DESC  2 /* the same curve walked backwards (e.g. hand out the stiff end of a sweep first) */

/* Fills in the reference's constructor defaults (options.hpp:49-58, irk.hpp:103-107,
 * newton.hpp:58-60, options.hpp:103-104). */
void rehuel_default_options(rehuel_options *opts);

/* irk::name_to_method / irk::method_to_name (irk.cpp:709-718).  Unknown name -> 0;
 * unknown id -> "". */
int rehuel_name_to_method(const char *name);
const char *rehuel_method_to_name(int method);

/* Problem registry ("robertson", "van-der-pol", "brusselator-1d", "brusselator-1d-dense", "exponential",
 * "stiff-equation", "brusselator", "lorenz", "harmonic", "dimer", "kinetic-4", "three-body").  Unknown -> 0. */
int rehuel_name_to_problem(const char *name);
/* n = number of equations, p = number of parameters.  n_hint selects n for size-generic
 * problems (brusselator-1d: n = 2*Ng), ignored otherwise.  Returns REHUEL_EINVAL if unknown. */
int rehuel_problem_dims(int problem, int n_hint, int *n, int *p);

/* irk::get_coefficients (irk.cpp:215-689) + the derived weights of irk.hpp:599-601.
 * A row-major s*s; all output arrays must hold 7 (49 for A) doubles.  Returns s, 0 if the
 * method is unknown to this library. */
int rehuel_get_coefficients(int method, double *A, double *b, double *c, double *b2, double *gamma,
                            int *order, int *order2, double *d_weights, double *d2_weights);

/* Ensemble sums of the per-member counters (rk_output::counters, irk.hpp:142-154) plus
 * newton_iters (sum of newton::status::iters over successful solves) and steps (accepted). */
typedef struct rehuel_stats {
	unsigned long long attempt, reject_newton, reject_err, newton_success, newton_incr_diverge,
	    newton_iter_error_too_large, newton_maxit_exceed, fun_evals, jac_evals;
	unsigned long long newton_iters, steps, n_failed; /* n_failed = members with status != 0 */
	unsigned long long max_attempt;                   /* max over members */
} rehuel_stats;

/* Result of an ensemble solve; replaces irk::rk_output (irk.hpp:140-163) : basic_output
 * (output.hpp:30-35), one entry per member.  OWNERSHIP: the library allocates every array
 * (page-locked host memory); release with rehuel_result_free().  Inputs are only borrowed for
 * the duration of the call. */
typedef struct rehuel_result {
	size_t M;
	int n;
	double *t_final;        /* [M]    time reached (== t1 on success) */
	double *y;              /* [n][M] state at t_final */
	double *err_last;       /* [M]    scaled error norm of the last attempt that reached the error estimate (for a member
	                           that ran to t1 that is its last accepted step, rk_output::err.back()) */
	int *status;            /* [M]    REHUEL_SUCCESS or a status bit */
	unsigned *attempt, *reject_newton, *reject_err, *newton_success, *newton_maxit_exceed,
	    *fun_evals, *jac_evals, *newton_iters, *steps; /* [M] each */
	/* samples (only when output_mode bit 0 is set and n_samples_cap > 0) */
	size_t n_samples_cap;   /* rows allocated per member */
	unsigned *n_samples;    /* [M] samples produced (may exceed the cap; rows beyond it are dropped); NULL when n_samples_cap == 0 */
	double *t_samples;      /* [cap][M]    sample 0 is (t0, y0), irk.hpp:581-585 */
	double *y_samples;      /* [cap][n][M] */
	double *err_samples;    /* [cap][M] */
	double *state;          /* [REHUEL_NSTATE][M] controller state at exit (options.keep_state), else NULL */
	size_t n_dense;         /* options.dense_samples, 0 if none */
	double *y_dense;        /* [n_dense][n][M] dense-output samples; grid points the member never reached are NaN */
	double elapsed_ms;      /* host wall time of the call (rk_output::elapsed_time is in ms, my_timer.hpp:142-149) */
	double kernel_ms;       /* device time of the integration kernel(s), max over GPUs */
	double accept_frac;     /* sum(steps)/sum(attempt)  (irk.hpp:864) */
	rehuel_stats sum;
	size_t trace_len;       /* rows of `trace` (options.out_interval > 0), else 0 */
	double *trace;          /* [trace_len][5]: step, t, dt, err, newton iterations of member options.trace_member, one row
	                           per logged attempt in the order the reference prints them (irk.hpp:805-810) */
	double phase_ms[4];     /* options.time_internals: REHUEL_PHASE_* device times in ms, summed over chunks and devices */
	void *impl;             /* private */
} rehuel_result;

#define REHUEL_PHASE_H2D     0
#define REHUEL_PHASE_ORDER   1
#define REHUEL_PHASE_KERNEL  2
#define REHUEL_PHASE_D2H     3

/* The reference's log text for rehuel_result.trace: the header line "    Rehuel: step  t  dt   err   iters" (irk.hpp:576)
 * and one "    Rehuel: <step> <t> <dt> <err> <iters>" line per row (irk.hpp:807-809) at `precision` significant digits
 * (the reference streams at the stream's precision; 6 = the iostream default).  Writes at most cap-1 characters + NUL;
 * returns the length the full text needs (like snprintf). */
size_t rehuel_result_format_log(const rehuel_result *res, int precision, char *buf, size_t cap);
/* print_timing_breakdown (irk.hpp:328-360) for rehuel_result.phase_ms, same layout ("    Rehuel: <method> done solving,
 * timings (ms | %):" + one line per phase); snprintf-like. */
size_t rehuel_result_format_timings(const rehuel_result *res, const char *method_name, char *buf, size_t cap);

/* Batched irk::odeint (irk.hpp:884-898): every member is its own adaptive integration.
 *   y0      [n][M], or [n] when y0_broadcast != 0
 *   params  [p][M]
 *   n_hint  see rehuel_problem_dims
 *   n_samples_cap  rows of sample storage per member (0: final state only)
 *   n_gpus  devices 0..n_gpus-1 of this process share the members (options.shard_mode)
 * Blocking.  Returns REHUEL_OK or a negative library error; numerical failures are reported per
 * member in out->status[]. */
int rehuel_irk_ensemble(int problem, int method, double t0, double t1, double dt0, size_t M,
                        int n_hint, const double *y0, int y0_broadcast, const double *params,
                        const rehuel_options *opts, size_t n_samples_cap, int n_gpus,
                        rehuel_result *out);
void rehuel_result_free(rehuel_result *out);
/* The library recycles page-locked host and device blocks across calls (cudaMallocHost/cudaMalloc
 * cost more than a solve); this returns every block not in use to the driver. */
void rehuel_release_cached_memory(void);

/* Single system = the reference call itself: returns the whole trajectory like
 * rk_output::t_vals / y_vals / err (irk.hpp:581-585, 825-832).  *n_out = samples written into
 * t_vals[cap], y_vals[cap][n] (row-major, one state per row), err_vals[cap]; if the trajectory
 * is longer than cap the function still returns REHUEL_OK, *n_out holds the full length and only
 * the first cap rows are valid.  counters[9] in the order of rk_output::counters;
 * *accept_frac as irk.hpp:864; *status per-member status. */
int rehuel_irk_single(int problem, int method, double t0, double t1, double dt0, int n_hint,
                      const double *y0, const double *params, const rehuel_options *opts,
                      size_t cap, size_t *n_out, double *t_vals, double *y_vals, double *err_vals,
                      unsigned long long *counters, double *accept_frac, int *status);

/* Device-resident variant used when inputs already live in HBM (bench "value", pipelines that
 * generate parameters on the GPU).  Every pointer is a DEVICE pointer on the current device;
 * optional outputs may be NULL.  Enqueues on `cuda_stream` (a cudaStream_t, NULL = default
 * stream) and returns without synchronising.  `queue` is an 8-byte scratch word. */
typedef struct rehuel_device_io {
	const double *y0;      /* [n][M] or [n] */
	int y0_broadcast;
	const double *params;  /* [p][M] */
	double *y;             /* [n][M] */
	double *t_final;       /* [M] */
	double *err_last;      /* [M] or NULL */
	int *status;           /* [M] */
	unsigned *counters;    /* [REHUEL_NCOUNTERS][M] or NULL; rows in REHUEL_CNT_* order */
	size_t n_samples_cap;
	unsigned *n_samples;   /* [M] or NULL */
	double *t_samples, *y_samples, *err_samples; /* as in rehuel_result, or NULL */
	unsigned long long *queue; /* [1] work-queue cursor, zeroed by the call */
	double *trace;         /* [trace_cap][5] (step,t,dt,err,iters) of member trace_member, or NULL */
	size_t trace_cap;
	size_t trace_member;
	unsigned *trace_len;   /* [1] */
	const double *rtol, *atol, *newton_tol; /* [M] each or NULL: per-member rel_tol / abs_tol / newton tol overriding
	                          the options (mixed-tolerance ensembles); all three or none */
	const double *t_start;  /* [M] or NULL: per-member start time instead of t0 (continuing a stopped ensemble) */
	const double *state_in; /* [REHUEL_NSTATE][M] or NULL: controller state to continue from (rows REHUEL_STATE_*) */
	double *state_out;      /* [REHUEL_NSTATE][M] or NULL: controller state at exit */
	size_t n_dense;         /* dense output: number of grid times t_k = dense_t0 + k*dense_dt (0 = off) */
	double dense_t0, dense_dt;
	double *y_dense;        /* [n_dense][n][M]; the caller pre-fills it (e.g. with NaN): only reached grid points are written */
	unsigned long long *stats; /* [REHUEL_NSTATS] or NULL: ensemble sums accumulated by the kernel itself (the caller zeroes
	                          them): words 0..8 the counters in REHUEL_CNT_* order, 9 members with status != 0,
	                          10 the maximum number of attempts of any member */
	const unsigned *order; /* [M] or NULL: queue position q -> member index order[q].  Lets the caller hand out
	                          expensive members first (longest-processing-time-first) when cost correlates with a
	                          parameter; results do not depend on it. */
} rehuel_device_io;

/* rows of the controller state (loop variables of irk_guts at exit, irk.hpp:560-590) */
#define REHUEL_STATE_DT          0 /* step size the next attempt would use */
#define REHUEL_STATE_DTS0        1 /* dts[0], dts[1]: step-size history of the controller (irk.hpp:853-855) */
#define REHUEL_STATE_DTS1        2
#define REHUEL_STATE_ERRQ        3 /* errs[1]^(1/(1+min(order,order2))): error history as the kernels carry it */
#define REHUEL_STATE_ALT         4 /* alternative_error_formula flag (irk.hpp:587,763,845) */
#define REHUEL_STATE_MAXIT       5 /* current Newton iteration budget (irk.hpp:644-648) */
#define REHUEL_STATE_STEP        6 /* accepted steps so far */
#define REHUEL_STATE_LAST_RELAX  7 /* step at which the budget was last raised */
#define REHUEL_NSTATE            8

#define REHUEL_CNT_ATTEMPT             0
#define REHUEL_CNT_REJECT_NEWTON       1
#define REHUEL_CNT_REJECT_ERR          2
#define REHUEL_CNT_NEWTON_SUCCESS      3
#define REHUEL_CNT_NEWTON_MAXIT_EXCEED 4
#define REHUEL_CNT_FUN_EVALS           5
#define REHUEL_CNT_JAC_EVALS           6
#define REHUEL_CNT_NEWTON_ITERS        7
#define REHUEL_CNT_STEPS               8
#define REHUEL_NCOUNTERS               9
#define REHUEL_NSTATS                  11

int rehuel_irk_ensemble_device(int problem, int method, double t0, double t1, double dt0, size_t M,
                               int n_hint, const rehuel_options *opts, const rehuel_device_io *io,
                               void *cuda_stream);

/* Parameter-locality hand-out order for rehuel_device_io.order (device pointers, enqueued on
 * `cuda_stream`, no synchronisation): order[q] = index of the member handed out q-th, following a
 * Morton curve through the first min(p,3) parameters (strictly positive parameters on a log
 * scale).  Members with nearby parameters take nearly the same steps, so the 32 lanes of a warp
 * stay in lockstep; results are unchanged.  `workspace` must hold
 * rehuel_locality_order_workspace(M) bytes (0 = error, see rehuel_last_error()).
 * rehuel_irk_ensemble does this internally per options.handout_order. */
size_t rehuel_locality_order_workspace(size_t M);
int rehuel_locality_order(const double *params, int p, size_t M, int descending, unsigned *order,
                          void *workspace, size_t workspace_bytes, void *cuda_stream);

/* Continues a stopped ensemble to a later end time: every member starts from prev->t_final[i],
 * prev->y[.][i] and the controller state prev->state (the previous call must have run with
 * options.keep_state != 0), i.e. the time loop carries on as if its end time had been t1 all along
 * -- no restart transient from dt0.  What irk::restore_state (irk.cpp:786-803, an empty stub) was
 * meant for; combine the legs with rehuel_result_merge.  Counters in `out` are those of this leg. */
int rehuel_irk_ensemble_continue(int problem, int method, double t1, int n_hint, const double *params,
                                 const rehuel_options *opts, size_t n_samples_cap, int n_gpus,
                                 const rehuel_result *prev, rehuel_result *out);

/* Batched irk::merge_rk_output (irk.cpp:750-783): concatenates two legs of a restarted ensemble
 * into `a` (counters added -- fun_evals/jac_evals are NOT, exactly like the reference --, status
 * OR-ed, accept_frac weighted by the number of stored samples).  `b` is left untouched. */
int rehuel_result_merge(rehuel_result *a, const rehuel_result *b);

/* Work cursor shared by the processes of one node ("over-decompose into chunks pulled from a host-side
 * atomic"): a 64-bit counter in a named POSIX shared-memory object.  One process per GPU (the torchrun layout) cuts the
 * ensemble into chunks and every rank claims the next chunk index with rehuel_cursor_next until the indices run out, so
 * a rank whose members are cheap takes more of them; inside ONE process rehuel_irk_ensemble(..., n_gpus > 1) does the
 * same with REHUEL_SHARD_DYNAMIC.  `name` without slashes; create != 0 creates (or re-creates) the object and zeroes
 * the counter, create == 0 opens an existing one.  NULL on error (rehuel_last_error()). */
typedef struct rehuel_cursor rehuel_cursor;
rehuel_cursor *rehuel_cursor_open(const char *name, int create);
/* Atomically adds `count` and returns the value before the add (the first index claimed). */
unsigned long long rehuel_cursor_next(rehuel_cursor *c, unsigned long long count);
void rehuel_cursor_reset(rehuel_cursor *c);
/* Unmaps; unlink != 0 also removes the name (the creating rank does that). */
void rehuel_cursor_close(rehuel_cursor *c, int unlink);

/* Number of kernels this library has launched in this process (bench.py "gpu_launches"). */
unsigned long long rehuel_kernel_launches(void);
/* Theoretical occupancy / register use of the kernel a (problem, method) pair dispatches to:
 * fills regs per thread, local (spill) bytes per thread, static smem bytes, resident CTAs/SM,
 * block size.  For profiles/ and DESIGN.md. */
int rehuel_kernel_info(int problem, int method, int n_hint, int *regs, int *local_bytes,
                       int *smem_bytes, int *ctas_per_sm, int *block);

/* Measures the achievable FP64 FMA rate of the current device (register-resident DFMA chains,
 * best of 5) in TFLOP/s: the roofline denominator for the compute-bound integration kernels.
 * MEASURED_PEAKS.json carries no FP64 figure. */
int rehuel_fp64_peak(double *tflops, double *ms);

/* Diagnostic: evaluates a built-in device functor once on the GPU: f[n] = fun(t, y), J[n*n] (row-major) = jac(t, y).
 * HOST pointers.  The reference checks its analytic Jacobians against finite differences in test/test_test_equations.cpp
 * (vacuously: newton.hpp:235-242 never updates the maximum); tests/test_gpu_parity.py does it for real through this. */
int rehuel_eval_functor(int problem, int n_hint, double t, const double *y, const double *params, double *f, double *J);

/* Diagnostic: the warp-per-system kernel's band LU + band solve (n = 64, half bandwidth 2) on caller-given
 * matrices, 4 per problem, side by side in the lanes of one warp as in the integrator.  ab: [n_problems][4][2
 * (re,im)][66][7] band rows (slot 2 + j - i holds A(i,j), rows 64..65 zero); rhs, x: [n_problems][4][2][64]; is_cplx,
 * flags: [n_problems][4].  pivot != 0: the fallback, dgbtf2-style row pivoting, one lane per matrix; pivot == 0: the
 * twisted factorisation without interchanges the integrator normally runs, one lane pair per matrix, flags[] =
 * "partial pivoting would have exchanged rows".  HOST pointers; for the test-suite. */
int rehuel_selftest_band(int n_problems, int pivot, const double *ab, const double *rhs, const int *is_cplx,
                         double *x, int *flags);
/* Diagnostic: on != 0 sends every attempt of the warp-per-system kernels through their pivoted band fallback
 * (which real problems practically never take), so that the test-suite can compare the two paths.  Process-wide. */
void rehuel_debug_force_pivot(int on);

/* Last error text of the calling thread (the reference reports through free-text log_out). */
const char *rehuel_last_error(void);
const char *rehuel_version(void);

#ifdef __cplusplus
}
#endif
#endif /* REHUEL_B200_H */
