/* qpde_b200.h — C ABI of libqpde_b200.so
 *
 * A B200 (sm_100a, FP64) replacement for ONE path of parsiad/QuantPDE: the
 * backward time-stepping sweep over a 1-D RectilinearGrid, batched across
 * contracts.  One call does what one `stepper.solve(grid, payoff, root, solver)`
 * does in the reference (QuantPDE/src/Core/IterativeMethod.hpp:1065-1079) for
 * every contract of a batch, where the reference's iteration tree is
 *
 *   ReverseConstantStepper                   Core/Stepper.hpp:9-38
 *    └ [ToleranceIteration]                  Core/IterativeMethod.hpp:1211-1254
 *   root = [MinPenaltyMethodDifference1 ∘]   Core/PenaltyMethod.hpp:168-303
 *          ReverseRannacher | ReverseCrankNicolson | ReverseBDFOne | ReverseBDFTwo
 *                                            Core/Rannacher.hpp, CrankNicolson.hpp, BDF.hpp
 *          ∘ BlackScholes1(grid, r, v, q)    Modules/Operators/BlackScholes.hpp:114-237
 *   events: Bermudan exercise  V <- max(V, obstacle)     Core/Event.hpp:100-112
 *   solver: SparseLUSolver (replaced by the in-kernel tridiagonal solve)
 *                                            Bindings/Eigen.hpp:143-165
 *
 * Plain pointers and sizes only; no C++ or torch types.  Every function
 * returns QPDE_OK (0) or a negative error code; qpde_last_error() gives text.
 * All array arguments are HOST pointers unless a name ends in `_dev`.
 * Arrays over contracts and nodes are contract-major: element (c, i) lives at
 * base[c * stride + i]; a stride of 0 shares one row between all contracts.
 *
 * There is no CPU fallback: qpde_create() fails if no sm_100 device is present.
 */
#ifndef QPDE_B200_H
#define QPDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QPDE_ABI_VERSION 6

/* ---- status codes ----------------------------------------------------- */
enum {
	QPDE_OK             =  0,
	QPDE_ERR_INVALID    = -1, /* bad argument (message says which)          */
	QPDE_ERR_CUDA       = -2, /* CUDA runtime error                         */
	QPDE_ERR_NO_DEVICE  = -3, /* no sm_100 GPU: there is no CPU fallback    */
	QPDE_ERR_NOT_CONVERGED = -4, /* informational: some contract hit max_inner;
	                                results are still written, see status[] */
	QPDE_ERR_UNSUPPORTED = -5
};

/* ---- time discretisation (which reference IterationNode is the root) -- */
enum {
	QPDE_SCHEME_RANNACHER = 0, /* ReverseRannacher,      Core/Rannacher.hpp:23-135   */
	QPDE_SCHEME_CN        = 1, /* ReverseCrankNicolson,  Core/CrankNicolson.hpp:52-105 */
	QPDE_SCHEME_BDF1      = 2, /* ReverseBDFOne,         Core/BDF.hpp:484-517        */
	QPDE_SCHEME_BDF2      = 3, /* ReverseBDFTwo,         Core/BDF.hpp:522-590        */
	/* Higher-order BDF (variable-step formulas, start-up through the lower orders, restart after
	 * every event): INTERLEAVED family, linear and penalised sweeps with constant steps. */
	QPDE_SCHEME_BDF3      = 4, /* ReverseBDFThree,       Core/BDF.hpp:139-192,595-668 */
	QPDE_SCHEME_BDF4      = 5, /* ReverseBDFFour,        Core/BDF.hpp:213-279,674-753 */
	QPDE_SCHEME_BDF5      = 6, /* ReverseBDFFive,        Core/BDF.hpp:299-373,759-844 */
	QPDE_SCHEME_BDF6      = 7  /* ReverseBDFSix,         Core/BDF.hpp:391-470,850-942 */
};

/* ---- initial condition / obstacle generators (Modules/Lambdas/Payoffs.hpp:13-51) */
enum {
	QPDE_PAYOFF_ARRAY        = 0, /* use the array given                    */
	QPDE_PAYOFF_PUT          = 1, /* S < K ? K - S : 0          :38-40      */
	QPDE_PAYOFF_CALL         = 2, /* S < K ? 0 : S - K          :29-31      */
	QPDE_PAYOFF_DIGITAL_PUT  = 3, /* S > K ? 0 : 1              :20-22      */
	QPDE_PAYOFF_DIGITAL_CALL = 4, /* S < K ? 0 : 1              :13-15      */
	QPDE_PAYOFF_STRADDLE     = 5, /* S < K ? K - S : S - K      :47-49      */
	QPDE_PAYOFF_K_MINUS_S    = 6, /* K - S (exercise value of tutorial.cpp:103-105) */
	QPDE_EXERCISE_NONE       = -1 /* exercise_kind only: the event times carry no exercise transform */
};

/* ---- discrete dividend events (qpde_bs1d_batch.dividend_kind) ----------- */
enum { QPDE_DIVIDEND_NONE = 0, QPDE_DIVIDEND_CASH = 1, QPDE_DIVIDEND_PROPORTIONAL = 2 };

/* ---- volatility model of BlackScholes1's `v` argument ----------------- */
enum {
	QPDE_VOL_CONST     = 0, /* v = sigma[c]                                 */
	QPDE_VOL_INVERSE_S = 1, /* v(S) = sigma[c] / S  (tutorial.cpp:120-123)  */
	QPDE_VOL_NODES     = 2  /* v given per node in sigma_nodes              */
};

/* ---- kernel selection -------------------------------------------------- */
enum {
	QPDE_KERNEL_AUTO        = 0,
	QPDE_KERNEL_INTERLEAVED = 1, /* one thread per contract, [N][C] HBM layout */
	QPDE_KERNEL_RESIDENT    = 2  /* one CTA per contract, state in registers/SMEM */
};

/* ---- output selection (bit mask) -------------------------------------- */
enum {
	QPDE_OUT_V           = 1,
	QPDE_OUT_S           = 2,
	QPDE_OUT_VALUE       = 4,
	QPDE_OUT_STEPS       = 8,
	QPDE_OUT_INNER_TOTAL = 16,
	QPDE_OUT_INNER       = 32,
	QPDE_OUT_ACTIVE      = 64,
	QPDE_OUT_STATUS      = 128,
	QPDE_OUT_COMPUTED    = 256,
	QPDE_OUT_ALL         = 511
};

/* Op codes of an event program (qpde_bs1d_batch.event_program): the transform of a general
 * Event<1> (Core/Event.hpp:60-184) — what the reference takes as a lambda
 * [..] (const Interpolant1 &V, Real S) { ... } — written as a postfix expression over a small
 * stack, evaluated at every node S_i with V the piecewise-linear interpolant of the solution
 * before the event (Core/Event.hpp:100-112, Core/Interpolant.hpp:153-181,331-374).  Every
 * arithmetic op is one IEEE double operation on the two topmost entries (a below b), so a program
 * that spells the lambda's expression in its order reproduces it bit for bit.  CONST and PARAM are
 * followed by one operand word (an index). */
enum {
	QPDE_EV_END    = 0,  /* the result is the top of the stack                                  */
	QPDE_EV_S      = 1,  /* push S_i                                                              */
	QPDE_EV_CONST  = 2,  /* push event_consts[operand]                                            */
	QPDE_EV_PARAM  = 3,  /* push event_params[contract][event][operand] (per event time)          */
	QPDE_EV_ADD    = 4,  /* a + b                                                                 */
	QPDE_EV_SUB    = 5,  /* a - b                                                                 */
	QPDE_EV_MUL    = 6,  /* a * b                                                                 */
	QPDE_EV_DIV    = 7,  /* a / b                                                                 */
	QPDE_EV_MAX    = 8,  /* b > a ? b : a   ("if(tmp > best) best = tmp", QuantPDE::max)          */
	QPDE_EV_MIN    = 9,  /* b < a ? b : a                                                         */
	QPDE_EV_NEG    = 10, /* -a                                                                    */
	QPDE_EV_V      = 11, /* V(a): the interpolant of the pre-event solution, clamped at the ends  */
	QPDE_EV_GT     = 12, /* a > b ? 1 : 0                                                         */
	QPDE_EV_SELECT = 13  /* c, t, f on the stack (f on top): c != 0 ? t : f                       */
};
#define QPDE_EV_MAX_STACK 8
#define QPDE_EV_MAX_LENGTH 64

typedef struct qpde_handle qpde_handle;       /* device context + stream   */
typedef struct qpde_bs1d_plan qpde_bs1d_plan; /* device-resident batch     */

/* One entry of the replayed ReverseConstantStepper time loop
 * (Core/IterativeMethod.hpp:1259-1317, Core/Stepper.hpp:16-18). */
typedef struct {
	double  t;           /* implicit time of the step                      */
	double  dt;          /* stepper's dt after event clamping              */
	int32_t same_dt;     /* isTimestepTheSame(): dt == dtPrevious          */
	int32_t event_after; /* the event-loop body ends after this step       */
	int32_t user_events; /* user events applied after this step            */
	int32_t reserved;
} qpde_step;

/* Batch description: C contracts on grids of N nodes each. */
typedef struct {
	int32_t abi_version;     /* QPDE_ABI_VERSION                            */
	int32_t n_contracts;     /* C >= 1                                      */

	/* --- grid: RectilinearGrid1(axis).refined(refine)  (Core/Domain.hpp:1092-1151)
	 * axis: coarse ticks, strictly increasing, [C][n_axis] with axis_stride
	 * (0 = one axis shared by all contracts).  N = 2^refine (n_axis - 1) + 1. */
	const double *axis;
	int64_t axis_stride;
	int32_t n_axis;
	int32_t refine;

	/* --- operator BlackScholes1(grid, r, v, q)  (BlackScholes.hpp:114-226) */
	const double *r;         /* [C] interest rate                           */
	const double *sigma;     /* [C] volatility parameter (see vol_model)    */
	const double *q;         /* [C] dividend rate                           */
	int32_t vol_model;       /* QPDE_VOL_*                                  */
	int32_t forward;         /* 0: ReverseConstantStepper and the Reverse* discretisations (every example of
	                            the reference); 1: the forward-time classes — ForwardConstantStepper(0, T, dt)
	                            (TimeIteration<true>, IterativeMethod.hpp:1494-1495, Stepper.hpp:41) with
	                            ForwardRannacher / CrankNicolson / BDFOne .. BDFSix: time runs from 0 up to T,
	                            the payoff is the condition at t = 0, user events sit at T e / E for
	                            e = 1 .. n_exercise, and events at one time are applied in ascending add()
	                            order (dividend_first still names the transform applied first).  Constant
	                            steps only. */
	const double *sigma_nodes; /* [C][N] when vol_model == QPDE_VOL_NODES   */
	int64_t sigma_nodes_stride;

	/* --- stepper ReverseConstantStepper(0, T, dt)  (Stepper.hpp:27-36) */
	const double *T;         /* [C] expiry                                  */
	const double *dt;        /* [C] step size (e.g. T / 1000)               */
	int32_t scheme;          /* QPDE_SCHEME_*                               */
	int32_t max_steps;       /* upper bound on replayed steps per contract;
	                            0 = derive (the library replays the loop);
	                            required (> 0) with variable_target          */
	/* --- ReverseVariableStepper(0, T, dt, target, scale) instead of the constant
	 * stepper (Core/Stepper.hpp:60-126, examples/vanilla_options.cpp:52-58): dt[c]
	 * is the first step, afterwards dt *= target / (relativeError(V^{n+1}, V^n,
	 * scale) + epsilon) before the usual clamp at the terminal time.  NULL =
	 * constant steps.  No events (the reference's stepper reads an emptied history
	 * slot after one), no jump term.  A contract that would need more than
	 * max_steps steps stops with status 3. */
	const double *variable_target; /* [C] */
	double variable_scale;   /* relativeError scale of the stepper, 1 (EarlyInclude.hpp:63) */

	/* --- events: stepper.add(te, exerciseEvent, grid), te = T e / E for
	 * e = 0 .. n_exercise-1 (tutorial.cpp:108-114); 0 = none */
	int32_t n_exercise;
	int32_t exercise_kind;   /* QPDE_PAYOFF_* generator of the exercise value, or QPDE_EXERCISE_NONE */

	/* --- discrete dividends at the same event times: stepper.add(te, dividendEvent, grid)
	 * with  V(S) <- V(max(S - D, 0))  (cash; README.md:357-375, tutorial.cpp:82-99) or
	 * V(S) <- V((1 - D) S)  (proportional), V evaluated through the grid's piecewise-
	 * linear interpolant and mapped back pointwise (Core/Event.hpp:100-112,153-163,
	 * Core/Interpolant.hpp:153-181,331-374).  Events at one time are applied in
	 * descending add() order in a reverse sweep (IterativeMethod.hpp:1340-1352):
	 * dividend_first = 0: exercise transform, then dividend transform (the dividend
	 * events were add()ed before the exercise events — the dividend is paid before
	 * the holder's right to exercise); 1: the other way round. */
	int32_t dividend_kind;   /* QPDE_DIVIDEND_*                             */
	int32_t dividend_first;
	const double *dividend;  /* [C] D (cash amount >= 0, or proportion in [0, 1)) */

	/* --- initial condition and obstacle */
	const double *strike;    /* [C] K for the generators                    */
	int32_t payoff_kind;     /* initial condition V(T, S)                   */
	int32_t obstacle_kind;   /* penalty obstacle (american) — usually = payoff_kind */
	const double *payoff;    /* [C][N] when payoff_kind == QPDE_PAYOFF_ARRAY */
	int64_t payoff_stride;
	const double *obstacle;  /* [C][N] when obstacle_kind == QPDE_PAYOFF_ARRAY */
	int64_t obstacle_stride;

	/* --- American constraint: MinPenaltyMethodDifference1 + ToleranceIteration */
	int32_t american;        /* 0 = European/Bermudan (root = discretisation) */
	int32_t max_inner;       /* iteration guard per step (the reference has
	                            none and would loop forever); e.g. 100       */
	double  tolerance;       /* ToleranceIteration tolerance, 1e-6  (IterativeMethod.hpp:1243) */
	double  scale;           /* relativeError scale, 1              (EarlyInclude.hpp:63)     */
	double  penalty_tolerance; /* PenaltyMethod tolerance, 1e-6; penalty = 1/this (PenaltyMethod.hpp:85) */

	/* --- read-out point: value[c] = V(spot[c]) piecewise linear
	 * (Core/Interpolant.hpp:153-181,331-374); NULL = strike */
	const double *spot;

	/* --- HJB control: BlackScholes1(grid, Control1(grid), v, q) under
	 * Min/MaxPolicyIteration1_1(grid, controlGrid, bs) attached to the
	 * ToleranceIteration (Core/PolicyIteration.hpp:54-105,
	 * examples/unequal_borrowing_lending_rates.cpp:108-138).  The interest rate
	 * is the control; per inner iteration every node picks the control value
	 * that minimises (maximises) (A(q) x - b(q))_i.  Needs scheme BDF1 or BDF2,
	 * american = 0, n_exercise = 0; tolerance / scale / max_inner apply. */
	int32_t n_controls;      /* 0 = r is not a control; else 2 .. 8          */
	int32_t policy_max;      /* 1: MaxPolicyIteration1_1, 0: MinPolicyIteration1_1 */
	const double *controls;  /* [C][n_controls] ticks of the control grid, in order */
	int64_t controls_stride; /* 0 = one control grid shared by all contracts */

	/* --- jump diffusion: BlackScholesJumpDiffusion1(grid, r, v, q, lambda, density)
	 * (Modules/Operators/BlackScholes.hpp:271-461, examples/jump_diffusion.cpp:105-116).
	 * The operator gains lambda on the diagonal and the drift r - q - lambda kappa;
	 * its b() is the explicit correlation integral lambda h(log S_i), h = circular
	 * correlation of V(exp(x0 + k dx)) with the density weights (d'Halluin et al.).
	 * x0, dx, kappa and the weights are what the reference computes once at
	 * construction by adaptive quadrature; qpde_jump_setup() reproduces them on
	 * the host.  jump_lambda == NULL: no jumps.  Needs american = 0, no controls,
	 * scheme BDF1 or BDF2. */
	const double *jump_lambda;   /* [C] mean arrival rate                    */
	const double *jump_kappa;    /* [C] E[eta] - 1                           */
	const double *jump_x0;       /* [C] log-grid origin log S_{i0}           */
	const double *jump_dx;       /* [C] log-grid spacing                     */
	const double *jump_weights;  /* [C][n_fft] density weights f'_j (wrapped order) */
	int64_t jump_weights_stride; /* 0 = one weight vector shared by all contracts */
	int32_t n_fft;               /* power of two >= N (qpde_jump_nfft)       */
	int32_t event_program_length;/* words in event_program (<= QPDE_EV_MAX_LENGTH), 0 = none */

	/* --- general events: at each of the n_exercise event times the solution is replaced by
	 * program(V, S_i) at every node (QPDE_EV_* above; e.g. the three-way choice of
	 * examples/experimental/glwb.cpp:84-121).  Takes the place of the tagged exercise / dividend
	 * transforms (exercise_kind = QPDE_EXERCISE_NONE, dividend_kind = QPDE_DIVIDEND_NONE); linear
	 * and penalised sweeps, no policy iteration, no jump term. */
	const int32_t *event_program;
	const double *event_consts;  /* [n_event_consts] shared by the batch                    */
	const double *event_params;  /* [C][n_exercise][n_event_params]; row e belongs to the event at T e / E
	                                (forward: T (e + 1) / E)                                  */
	int64_t event_params_stride; /* doubles between contracts; 0 = one table shared by all contracts */
	int32_t n_event_consts;
	int32_t n_event_params;

	/* --- explicit event times: stepper.add(event_times[e], ...) instead of the uniform T e / E
	 * ([n_exercise], shared by the batch, any order, within [0, T]; NULL = uniform).  Row e of
	 * event_params belongs to the e-th SMALLEST time.  A time equal to the initial time of the sweep is
	 * taken as the reference takes it: a zero-length first step, then the event
	 * (examples/experimental/glwb.cpp:84-121 adds one at t = T). */
	const double *event_times;

	/* --- operator source: b(t) of the LinearSystem = source_table[floor(t)] * S_i instead of 0 — the
	 * GLWBOperator of examples/experimental/glwb.cpp:31-45, b(t) = -(R[n+1] - R[n]) S with n = floor(t)
	 * (the index is clamped to the table).  [n_source], shared by the batch; NULL = none.  BDF1 / BDF2
	 * only (b enters at one time), no penalty, no policy iteration, no jump term. */
	const double *source_table;
	int32_t n_source;
	int32_t reserved3;

	int32_t kernel;          /* QPDE_KERNEL_* (AUTO picks by N)             */
	int32_t outputs;         /* split form only: QPDE_OUT_* mask of the outputs
	                            to keep on the device; 0 = all.  The one-shot
	                            call derives it from the result pointers.   */
} qpde_bs1d_batch;

/* Output buffers (HOST).  Any pointer may be NULL to skip that output. */
typedef struct {
	double  *V;           /* [C][N] node values at t = 0                    */
	int64_t  V_stride;    /* >= N                                           */
	double  *S;           /* [C][N] the refined grid actually used          */
	int64_t  S_stride;
	double  *value;       /* [C] V(spot)                                    */
	int32_t *n_steps;     /* [C] time steps taken (Iteration::iterations()[0]) */
	int32_t *inner_total; /* [C] sum over steps of inner iterations         */
	uint8_t *inner;       /* [C][inner_stride] per-step inner iteration counts
	                         (ToleranceIteration::iterations(), saturated at 255) */
	int64_t  inner_stride;/* >= max steps                                   */
	uint8_t *active;      /* [C][N] final penalty active set, 0/1
	                         (PenaltyMethod::constraintMask, PenaltyMethod.hpp:127-139) */
	int64_t  active_stride;
	int32_t *status;      /* [C] 0 ok, 1 = some step hit max_inner, 3 = max_steps exhausted (variable steps) */
	int32_t *solves_computed; /* [C] linear solves the device actually ran.  inner_total is the reference's
	                         count (ToleranceIteration::iterations()); an inner iteration whose active set
	                         equals the previous one's would reproduce that iterate bit for bit and is
	                         counted, not run — the difference is work the reference does and the device skips */
} qpde_bs1d_result;

/* ---- context ---------------------------------------------------------- */
int qpde_abi_version(void);
int qpde_create(int device, qpde_handle **out);
/* A handle keeps device blocks of destroyed plans (up to 2 GB each, 6 GB in all) for its next plan,
 * so that back-to-back qpde_bs1d_sweep() calls do not pay cudaMalloc / cudaFree; qpde_destroy()
 * releases them. */
int qpde_destroy(qpde_handle *h);
const char *qpde_last_error(const qpde_handle *h); /* h may be NULL: global */
/* The CUDA stream (cudaStream_t) every launch of this handle goes to, for
 * callers that time with their own CUDA events. */
void *qpde_stream(qpde_handle *h);
int qpde_synchronize(qpde_handle *h);
/* Number of kernel launches issued by this handle so far. */
int64_t qpde_launch_count(const qpde_handle *h);
/* Multiprocessor count of the device (grid sizing / reporting). */
int qpde_sm_count(const qpde_handle *h);

/* Measured FP64 rate of the device in TFLOP/s (a chain of independent fused multiply-adds, 2 flops
 * each): the denominator of the "fp64" rooflines bench.py reports for the control and jump kernels. */
int qpde_measure_fp64_peak(qpde_handle *h, double *tflops);

/* Pinned host memory for callers that want asynchronous, full-rate copies. */
int qpde_host_alloc(size_t bytes, void **out);
int qpde_host_free(void *p);

/* ---- host-only helpers (no GPU needed) --------------------------------- */
/* Replays the reference time loop.  Returns the number of steps (>= 0) and
 * writes min(n, max_steps) entries; out may be NULL to count. */
int qpde_make_schedule(double T, double dt, const double *event_times,
		int n_events, qpde_step *out, int max_steps);
/* The same for ForwardConstantStepper(0, T, dt) (TimeIteration<true>, IterativeMethod.hpp:1494-1495):
 * time runs from 0 up to T, the smallest event time is popped first. */
int qpde_make_schedule_forward(double T, double dt, const double *event_times,
		int n_events, qpde_step *out, int max_steps);
/* N after refinement. */
int qpde_refined_size(int n_axis, int refine);

/* Jump-diffusion set-up on the host, as BlackScholesJumpDiffusion does at
 * construction (BlackScholes.hpp:53-62 kappa, :275-297 log-grid, :299-329
 * density weights; quadrature Core/Integral.hpp:122-212,367-456).
 * density_kind: QPDE_DENSITY_LOGNORMAL (params = mu, sigma) or
 * QPDE_DENSITY_DOUBLE_EXPONENTIAL (params = p, eta_1, eta_2).
 * S: the refined grid [n_nodes]; weights: [qpde_jump_nfft(n_nodes)] out. */
enum { QPDE_DENSITY_LOGNORMAL = 0, QPDE_DENSITY_DOUBLE_EXPONENTIAL = 1 };
int qpde_jump_nfft(int n_nodes);
int qpde_jump_setup(int n_nodes, const double *S, int density_kind,
		const double *params, double *x0, double *dx, double *kappa, double *weights);
/* The refined grid RectilinearGrid1(axis).refined(refine) on the host. */
int qpde_refine_axis(const double *axis, int n_axis, int refine, double *out);

/* ---- the sweep -------------------------------------------------------- */
/* One-shot: upload (H2D), run, download (D2H).  This is the call the
 * reference-facing host API makes (quantpde_b200/include/QuantPDE_B200). */
int qpde_bs1d_sweep(qpde_handle *h, const qpde_bs1d_batch *batch,
		qpde_bs1d_result *result);

/* Split form, for callers that keep a batch resident in HBM. */
int qpde_bs1d_plan_create(qpde_handle *h, const qpde_bs1d_batch *batch,
		qpde_bs1d_plan **out);      /* validates, allocates, uploads      */
int qpde_bs1d_plan_run(qpde_bs1d_plan *plan);    /* async on qpde_stream() */
int qpde_bs1d_plan_fetch(qpde_bs1d_plan *plan, qpde_bs1d_result *result); /* syncs */
int qpde_bs1d_plan_destroy(qpde_bs1d_plan *plan);
/* Which kernel the plan resolved to (QPDE_KERNEL_*), and N. */
int qpde_bs1d_plan_kernel(const qpde_bs1d_plan *plan);
int qpde_bs1d_plan_nodes(const qpde_bs1d_plan *plan);
/* 1 if the plan's sweep stages its HBM streams through TMA (the INTERLEAVED family at
 * N >= 32 without policy iteration, up to BDF2), else 0. */
int qpde_bs1d_plan_tma_staged(const qpde_bs1d_plan *plan);
/* Row length of the per-step `inner` output (upper bound on steps). */
int qpde_bs1d_plan_max_steps(const qpde_bs1d_plan *plan);

/* ---- GMWB: 2-D (W, A) impulse-control HJBQVI, penalised scheme ------------- */
/* One call does what HJBQVI<2,1,1>::solve(refinement) does for the model of
 * examples/hjbqvi/gmwb.cpp:35-208 (Modules/HJBQVI/HJBQVI.hpp:590-811 with
 * usePenalizedScheme(), useBiCGSTABSolver(), HJBQVILinearBoundary on W and
 * HJBQVIZeroDiffusionRightBoundary on A):
 *   state   w = account value, a = guarantee balance; volatility sigma w in W, none in A
 *   drift   ((r - eta) w - q 1{a>0}, -q 1{a>0}),  flow q 1{a>0},  discount r
 *   control q in `controls` (the stochastic control grid, not refined)
 *   impulse z = z_frac a: (w, a) -> (max(w - z, 0), a - z), reward (1 - kappa) z - c,
 *           z_frac on `impulse_axis` refined `refine` times
 *   exit    V(T, w, a) = max(w, (1 - kappa) a - c)
 *   time    ReverseConstantStepper(0, T, T / timesteps), BDF1, ToleranceIteration(tolerance)
 *   penalty 1 / (scaling_factor dt)
 * `timesteps` is the count at this refinement (the reference doubles it per level).
 * Grid: RectilinearGrid2(w_axis, a_axis).refined(refine); node (iw, ia) is element
 * ia * n_w + iw of V (W fastest, Core/Domain.hpp:1408-1428). */
typedef struct {
	int32_t abi_version;
	int32_t refine;
	const double *w_axis; int32_t n_w_axis;
	const double *a_axis; int32_t n_a_axis;
	const double *controls; int32_t n_controls;      /* e.g. {0, G}                    */
	const double *impulse_axis; int32_t n_impulse_axis; /* e.g. uniform(0, 1, 3)       */
	double T; int32_t timesteps; int32_t max_inner;  /* inner-iteration guard per step  */
	double r, eta, sigma, kappa, c;
	double scaling_factor;   /* HJBQVI default 1e-2                                     */
	double tolerance;        /* iteration_tolerance, 1e-6                               */
	int32_t max_sweeps;      /* guard of the linear solve's Gauss-Seidel sweeps, e.g. 500 */
	int32_t reserved;
} qpde_gmwb2d;

typedef struct {
	int32_t nodes_w, nodes_a; /* refined grid sizes                                     */
	int32_t n_steps;          /* time steps taken                                       */
	int32_t status;           /* 0 ok, 1 max_inner hit, 2 the linear sweeps did not settle */
	double  mean_inner;       /* "Mean Policy Iterations" of HJBQVI_main                */
	int32_t *inner;           /* [inner_capacity] per-step inner iterations, or NULL    */
	int32_t inner_capacity;
	int32_t reserved;
} qpde_gmwb2d_stats;

/* Refined grid sizes for the buffers the caller allocates. */
int qpde_gmwb2d_nodes(const qpde_gmwb2d *problem, int *nodes_w, int *nodes_a);
/* V: HOST buffer [nodes_a][nodes_w] out; W_out / A_out: optional refined axes. */
int qpde_gmwb2d_solve(qpde_handle *h, const qpde_gmwb2d *problem, double *V,
		double *W_out, double *A_out, qpde_gmwb2d_stats *stats);

/* ---- 1-D impulse-control HJBQVI, penalised scheme: a batch of problems in one launch --------
 * One call does what HJBQVI<1,1,1>::solve(refinement) does (Modules/HJBQVI/HJBQVI.hpp:590-811 with
 * usePenalizedScheme(), useBiCGSTABSolver(), no boundary routines) for every problem of a batch; one CTA
 * per problem, the whole sweep on chip (qpde_hjbqvi1d.cu).  The coefficient functions are tagged forms:
 *   QPDE_HJBQVI1D_EXCHANGE_RATE   examples/hjbqvi/exchange_rate.cpp:118-152
 *     params[7] = { rho, sigma, a, b, m, lambda, c }:  discount rho, volatility sigma, drift -a q,
 *     flow -(x - m)^2 - b q^2, impulse x -> z with reward -lambda |z - x| - c (candidates with
 *     |z - m| >= |x - m| are pruned), exit value 0
 *   time    ReverseConstantStepper(0, T, T / timesteps), BDF1, ToleranceIteration(tolerance)
 *   penalty 1 / (scaling_factor dt)
 * Grids: the coarse axes refined `refine` times (RectilinearGrid::refined); `timesteps` is the count at
 * this refinement (the reference doubles it per level).  An axis stride of 0 shares the axis. */
#define QPDE_HJBQVI1D_EXCHANGE_RATE 1
#define QPDE_HJBQVI1D_MAX_NODES 2304
typedef struct {
	int32_t abi_version;
	int32_t model;            /* QPDE_HJBQVI1D_*                                         */
	int32_t n_problems;
	int32_t refine;
	const double *x_axis;       int32_t n_x_axis;       int32_t reserved0; int64_t x_axis_stride;       /* state grid            */
	const double *control_axis; int32_t n_control_axis; int32_t reserved1; int64_t control_axis_stride; /* stochastic controls q */
	const double *impulse_axis; int32_t n_impulse_axis; int32_t reserved2; int64_t impulse_axis_stride; /* impulse controls z    */
	const double *params;       int32_t n_params;       int32_t reserved3; int64_t params_stride;       /* [n_problems][n_params] */
	double T; int32_t timesteps; int32_t max_inner;  /* inner-iteration guard per step       */
	double scaling_factor;    /* HJBQVI default 1e-2                                     */
	double tolerance;         /* iteration_tolerance, 1e-6                               */
	int32_t max_sweeps;       /* guard of the linear solve's lagged sweeps, e.g. 500     */
	int32_t reserved4;
} qpde_hjbqvi1d_batch;

/* HOST buffers, [n_problems][nodes] (any but V may be NULL); Q / Z are NaN where the other control
 * acts (HJBQVI.hpp:977-990), mask is PenaltyMethod::constraintMask (PenaltyMethod.hpp:127-139). */
typedef struct {
	double *V, *X, *Q, *Z;
	unsigned char *mask;
	double *mean_inner;       /* [n_problems] "Mean Policy Iterations" of HJBQVI_main      */
	double *mean_sweeps;      /* [n_problems] tridiagonal solves per policy iteration (the reference reports BiCGSTAB's "Mean Solver Iterations" there) */
	int32_t *n_steps;         /* [n_problems]                                             */
	int32_t *status;          /* [n_problems] 0 ok, 1 max_inner hit, 2 the lagged sweeps did not settle */
} qpde_hjbqvi1d_result;

/* Refined sizes for the buffers the caller allocates. */
int qpde_hjbqvi1d_nodes(const qpde_hjbqvi1d_batch *batch, int *nodes, int *n_controls, int *n_impulses);
int qpde_hjbqvi1d_solve(qpde_handle *h, const qpde_hjbqvi1d_batch *batch, qpde_hjbqvi1d_result *result);

/* ---- 2-D impulse-control HJBQVI, penalised scheme, impulses in both directions ----------------
 * One call does what HJBQVI<2,1,1>::solve(refinement) does (Modules/HJBQVI/HJBQVI.hpp:590-811 with
 * usePenalizedScheme(), useBiCGSTABSolver(), no boundary routines) for a model whose impulses point both
 * ways in the state, so that the matrix of an inner iteration has no block-triangular order (the GMWB
 * entry points above exploit one).  Tagged coefficient forms:
 *   QPDE_HJBQVI2D_OPTIMAL_CONSUMPTION   examples/hjbqvi/optimal_consumption.cpp:85-176
 *     params[8] = { rho, r, mu, sigma, gamma, lambda, c, boundary }: state (w, b) = (stock, bank) on
 *     [0, boundary]^2, volatility (sigma w, 0), drift (mu w, r b - q 1{0 < b < boundary}), flow
 *     q^gamma / gamma 1{..}, discount rho; impulse z = z_frac (hi - lo) + lo within the admissible
 *     transfers: (w, b) -> (w + z, b - z - lambda |z| - c), reward 0 (-inf if z = 0); exit value
 *     max(b + (1 - lambda) w - c, 0)^gamma / gamma
 *   time    ReverseConstantStepper(0, T, T / timesteps), BDF1, ToleranceIteration(tolerance)
 *   penalty 1 / (scaling_factor dt)
 * All four axes are refined `refine` times; `timesteps` is the count at this refinement.  Node (iw, ib)
 * is element ib * n_w + iw of the outputs (w fastest, Core/Domain.hpp:1408-1428). */
#define QPDE_HJBQVI2D_OPTIMAL_CONSUMPTION 1
typedef struct {
	int32_t abi_version;
	int32_t model;            /* QPDE_HJBQVI2D_*                                         */
	int32_t refine;
	int32_t n_params;
	const double *w_axis;       int32_t n_w_axis;       int32_t reserved0;
	const double *b_axis;       int32_t n_b_axis;       int32_t reserved1;
	const double *control_axis; int32_t n_control_axis; int32_t reserved2;   /* consumption rates q */
	const double *impulse_axis; int32_t n_impulse_axis; int32_t reserved3;   /* z_frac in [0, 1]    */
	const double *params;
	double T; int32_t timesteps; int32_t max_inner;
	double scaling_factor;    /* HJBQVI default 1e-2                                     */
	double tolerance;         /* iteration_tolerance, 1e-6                               */
	int32_t max_sweeps;       /* guard of the linear solve's line sweeps, e.g. 20000     */
	int32_t reserved4;
} qpde_hjbqvi2d;

typedef struct {
	int32_t nodes_w, nodes_b; /* refined grid sizes                                     */
	int32_t n_steps;          /* time steps taken                                       */
	int32_t status;           /* 0 ok, 1 max_inner hit, 2 the line sweeps did not settle */
	double  mean_inner;       /* "Mean Policy Iterations" of HJBQVI_main                */
	double  mean_sweeps;      /* line sweeps per policy iteration (the reference reports BiCGSTAB's "Mean Solver Iterations") */
	int32_t *inner;           /* [inner_capacity] per-step inner iterations, or NULL    */
	int32_t inner_capacity;
	int32_t reserved;
} qpde_hjbqvi2d_stats;

int qpde_hjbqvi2d_nodes(const qpde_hjbqvi2d *problem, int *nodes_w, int *nodes_b);
/* V: HOST buffer [nodes_b][nodes_w] out; Q, Z (controls, NaN where the other control acts), mask
 * (PenaltyMethod::constraintMask), W_out, B_out: optional. */
int qpde_hjbqvi2d_solve(qpde_handle *h, const qpde_hjbqvi2d *problem, double *V, double *Q, double *Z,
		unsigned char *mask, double *W_out, double *B_out, qpde_hjbqvi2d_stats *stats);

/* ---- GMWB sharded by the A-axis over the GPUs of a node: the library's own driver --------
 * One process per GPU.  A handle gets a communicator (NCCL over NVLink /
 * NVSwitch, loaded at run time): rank 0 makes an id with qpde_comm_unique_id, the caller ships the
 * 128 bytes to the other ranks by whatever it has (MPI, torch.distributed, a file), every rank calls
 * qpde_comm_init on its own handle.  qpde_gmwb2d_solve_sharded is then a COLLECTIVE call with the
 * arguments of qpde_gmwb2d_solve: the control search (HJBQVI.hpp:590-754 — the impulse arg-max over
 * every node) is cut along the A-axis over the ranks, the assembled rows of each block reach all
 * ranks by NCCL broadcasts on the library's streams (the all-gather in front of the impulse step),
 * and every rank returns the full V, bit-identical to the single-GPU solve.  No host round trip
 * besides the stopping flag; DESIGN.md §5. */
#define QPDE_COMM_ID_BYTES 128
int qpde_comm_unique_id(unsigned char id[QPDE_COMM_ID_BYTES]);
int qpde_comm_init(qpde_handle *h, const unsigned char id[QPDE_COMM_ID_BYTES], int rank, int world);
int qpde_comm_destroy(qpde_handle *h);
int qpde_gmwb2d_solve_sharded(qpde_handle *h, const qpde_gmwb2d *problem, double *V,
		double *W_out, double *A_out, qpde_gmwb2d_stats *stats);

/* ---- GMWB sharded by the A-axis, piecewise (one shard per rank / GPU; the caller drives) ---- */
/* The pieces of one inner iteration restricted to the refined A-lines
 * [a_begin, a_end) a rank owns.  x, v0, xprev are DEVICE pointers to full-grid
 * vectors [nodes_a][nodes_w] replicated on every rank (e.g. torch tensors); the
 * caller moves the owned lines between ranks (NCCL broadcast / all-gather of the
 * iterate: the impulse targets (w - z, a - z) live on arbitrary lower A) and
 * reduces the flags.  The system is block lower-triangular by A-line, so the
 * lines are solved in ascending A across the ranks (v0_dev: V^n of the running time step): rank k runs its pass after
 * rank k-1 has published its lines, and one round is a direct solve.
 * quantpde_b200/gmwb_sharded.py is the driver; see DESIGN.md §5. */
typedef struct qpde_gmwb2d_shard qpde_gmwb2d_shard;
int qpde_gmwb2d_shard_create(qpde_handle *h, const qpde_gmwb2d *problem, int a_begin, int a_end,
		qpde_gmwb2d_shard **out);
int qpde_gmwb2d_shard_destroy(qpde_gmwb2d_shard *shard);
/* V(T) on the owned lines */
int qpde_gmwb2d_shard_exit(qpde_gmwb2d_shard *shard, double *x_dev);
/* control searches + row assembly of the owned lines from the full previous iterate (the rows do not
 * depend on V^n: the line pass adds it) */
int qpde_gmwb2d_shard_policy(qpde_gmwb2d_shard *shard, const double *x_dev, double h0);
/* the line solves of the owned lines in ascending A, exact once the lines below
 * a_begin are final in x_dev; *status_dev (device int) |= 2 if a line that is
 * iterated in place (same-line impulse targets) did not settle in max_sweeps */
int qpde_gmwb2d_shard_sweep(qpde_gmwb2d_shard *shard, double *x_dev, const double *v0_dev, int *status_dev);
/* relativeError(x, xprev) > tolerance on the owned lines; *notdone_dev (device int) |= 1 */
int qpde_gmwb2d_shard_relerr(qpde_gmwb2d_shard *shard, const double *x_dev, const double *xprev_dev,
		int *notdone_dev);

#ifdef __cplusplus
}
#endif

#endif /* QPDE_B200_H */
