/* TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C) of the reference's 1-D
 * time-stepping hot path, used as the parity checker by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline leg.  The product
 * path (quantpde_b200/, libqpde_b200.so) never includes, links or calls this.
 *
 * Parity pin: every function here is checked bit-for-bit against programs
 * built from the UNMODIFIED reference headers (oracle/_ref/qpde_ref, see
 * oracle/Makefile) and against the committed fixtures in tests/golden/.
 * The linear algebra underneath those programs is oracle/shim's (Eigen 3 is
 * not available offline), so the LU arithmetic is "Eigen-like", not Eigen's.
 *
 * All citations are file:line under /root/reference/QuantPDE/src.
 */
#ifndef QPDE_ORACLE_H
#define QPDE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum {
	QPO_SCHEME_RANNACHER = 0, /* Core/Rannacher.hpp:23-121 */
	QPO_SCHEME_CN        = 1, /* Core/CrankNicolson.hpp:52-78, theta = 1/2 */
	QPO_SCHEME_BDF1      = 2, /* Core/BDF.hpp:32-46,484-517 */
	QPO_SCHEME_BDF2      = 3, /* Core/BDF.hpp:68-117,522-590 */
	QPO_SCHEME_BDF3      = 4, /* Core/BDF.hpp:139-192,595-668: BDF1, BDF2, then BDF3 steps */
	QPO_SCHEME_BDF4      = 5, /* Core/BDF.hpp:213-279,674-753 */
	QPO_SCHEME_BDF5      = 6, /* Core/BDF.hpp:299-373,759-844 */
	QPO_SCHEME_BDF6      = 7  /* Core/BDF.hpp:391-470,850-942 */
};

/* One entry of the replayed time loop (Core/IterativeMethod.hpp:1259-1317). */
typedef struct {
	double t;        /* implicit time of the step                         */
	double dt;       /* the stepper's dt variable after clamping          */
	int    same_dt;  /* TimeIteration::isTimestepTheSame(): dt==dtPrevious */
	int    event_after; /* 1 if the event loop body ends after this step
	                       (afterEvent() + history clear follow)           */
	int    user_events; /* user events applied after this step (0 when only
	                       the terminating NullEvent fires)                */
} qpo_step;

/* Axis::special, 33 ticks (Core/Axis.hpp:445-459). Returns 33. */
int qpo_axis_special(double *out);
/* Axis union (Core/Axis.hpp:270-313). out must hold na+nb. Returns size. */
int qpo_axis_union(const double *a, int na, const double *b, int nb, double *out);
/* RectilinearGrid::refined (Core/Domain.hpp:1092-1151). Returns new size
 * 2^times (n-1) + 1; out may be NULL to query the size. */
int qpo_axis_refine(const double *in, int n, int times, double *out);

/* Replays ReverseConstantStepper over [0, T] with events at event_times
 * (any order).  Returns the number of steps (writes at most max_steps). */
int qpo_schedule(double T, double dt, const double *event_times, int n_events,
		qpo_step *out, int max_steps);
/* The same for ForwardConstantStepper (time runs from 0 up to T; events may not sit at 0). */
int qpo_schedule_forward(double T, double dt, const double *event_times, int n_events,
		qpo_step *out, int max_steps);

/* BlackScholes<1,0>::A rows (Modules/Operators/BlackScholes.hpp:136-226).
 * r, v, q, lambda are per-node arrays of length n. */
void qpo_bs_coeffs(int n, const double *S, const double *r, const double *v,
		const double *q, const double *lambda, double kappa,
		double *lo, double *di, double *up);

/* Jump-diffusion set-up and explicit jump term
 * (Modules/Operators/BlackScholes.hpp:53-62, 275-329, 389-459). */
enum { QPO_DENSITY_LOGNORMAL = 0, QPO_DENSITY_DOUBLE_EXPONENTIAL = 1 };
int qpo_jump_nfft(int n);
/* weights must hold qpo_jump_nfft(n) doubles.  lognormal: p0 = mu, p1 = sigma;
 * double exponential: p0 = p, p1 = eta_1, p2 = eta_2. */
void qpo_jump_setup(int n, const double *S, int density_kind, double p0, double p1,
		double p2, double *x0, double *dx, double *weights, double *kappa);
void qpo_jump_term(int n, const double *S, const double *v, double lambda, int N,
		double x0, double dx, const double *weights, double *b);

/* Piecewise-linear read-out (Core/Interpolant.hpp:153-181,331-374). */
double qpo_interp(int n, const double *x, const double *v, double c);

typedef struct {
	int n;                 /* nodes */
	double T;              /* initial (expiry) time of the reverse sweep   */
	const double *S;       /* [n] */
	const double *lo, *di, *up; /* operator rows from qpo_bs_coeffs        */
	const double *v0;      /* [n] initial condition at t = T               */
	int scheme;            /* QPO_SCHEME_*                                 */
	int n_steps;
	const qpo_step *steps; /* from qpo_schedule                            */
	int american;          /* 1: MinPenaltyMethodDifference + ToleranceIteration */
	const double *obstacle;/* [n] payoff for the penalty / exercise events */
	double tolerance;      /* ToleranceIteration tolerance (1e-6)          */
	double scale;          /* relativeError scale (1)                      */
	double penalty_tol;    /* PenaltyMethod tolerance; scale = 1/this      */
	int max_inner;         /* guard (the reference has none)               */
	const double *event_obstacle; /* [n] or NULL: at each user event
	                          V = max(V, event_obstacle) (tutorial.cpp:103-105) */
	int n_controls;        /* > 0: PolicyIteration over this many control values */
	int policy_max;        /* 1: MaxPolicyIteration, 0: MinPolicyIteration */
	const double *plo, *pdi, *pup; /* [n_controls][n] operator rows per control */
	/* jump diffusion (NULL weights = none): b() of the operator is the explicit
	 * correlation integral; lo/di/up must already carry lambda and kappa */
	const double *jump_weights; int jump_nfft;
	double jump_lambda, jump_x0, jump_dx;
	int solver;            /* 0: pivoted LU (the reference's policy);
	                          1: arithmetic model of the INTERLEAVED kernel   */
	/* discrete dividends at the user event times (README.md:357-375):
	 * V(S) <- V(max(S - D, 0)) [cash] or V((1 - D) S) [proportional] through the
	 * grid's piecewise-linear interpolant (Event.hpp:100-112).  dividend_first = 0:
	 * at one event time the reverse sweep applies the exercise transform, then the
	 * dividend transform (the dividend events were add()ed first); 1: the reverse. */
	int dividend_kind;     /* QPO_DIVIDEND_* */
	int dividend_first;
	double dividend;
	/* ReverseVariableStepper(0, T, variable_dt0, variable_target, variable_scale)
	 * (Core/Stepper.hpp:60-126, examples/vanilla_options.cpp:52-64) when
	 * variable_target > 0: `steps` is ignored, n_steps is the capacity of `inner`,
	 * *steps_taken the number of steps taken.  No user events (after an event the
	 * reference's stepper reads a history slot that clear() has emptied). */
	double variable_target, variable_scale, variable_dt0;
	int *steps_taken;
	/* general Event<1> transforms (Core/Event.hpp:60-184) as a postfix program over a stack, the op codes of
	 * include/qpde_b200.h (QPDE_EV_*) restated here: 0 END, 1 S, 2 CONST i, 3 PARAM i, 4 ADD, 5 SUB, 6 MUL, 7 DIV,
	 * 8 MAX (b > a ? b : a), 9 MIN, 10 NEG, 11 V (interpolant of the pre-event solution), 12 GT, 13 SELECT.
	 * NULL = the tagged transforms above.  event_params: [n_events][n_event_params], row e for the event at
	 * T e / E (reverse) or T (e + 1) / E (forward). */
	const int *event_program;
	const double *event_consts, *event_params;
	int n_events, n_event_params;
	/* b(t) of the operator = source_table[floor(t)] * S (NULL = 0): GLWBOperator::b, glwb.cpp:40-44 */
	const double *source_table; int n_source;
	/* 1: the forward-time classes (TimeIteration<true>, ForwardConstantStepper, ForwardRannacher ...):
	 * v0 is the condition at t = 0, `steps` comes from qpo_schedule_forward.  Constant steps only. */
	int forward;
} qpo_problem;
enum { QPO_DIVIDEND_NONE = 0, QPO_DIVIDEND_CASH = 1, QPO_DIVIDEND_PROPORTIONAL = 2 };

/* Runs the sweep.  V [n] out; inner [n_steps] out (may be NULL); mask [n] out
 * (may be NULL).  Returns 0, 1 if some step hit max_inner, 2 if the
 * combination is not restated (policy iteration under a theta scheme). */
int qpo_sweep_1d(const qpo_problem *p, double *V, int *inner,
		unsigned char *mask);

/* ---- GMWB 2-D impulse-control HJBQVI (qpde_oracle_gmwb.c) ----------------- */
typedef struct {
	int n_w, n_a;
	const double *W, *A;        /* refined axes; node (iw, ia) is row iw + n_w ia   */
	int n_q; const double *q;   /* stochastic control grid, e.g. {0, G}             */
	int n_z; const double *zfrac; /* impulse control grid (fractions of a)          */
	double T; int timesteps;
	double r, eta, sigma, kfee, c;
	double scaling_factor;      /* HJBQVI default 1e-2: penalty = 1 / (this * dt)    */
	double tolerance;           /* iteration_tolerance, 1e-6                         */
	int max_inner;
} qpo_gmwb;
/* V [n_w n_a] out.  Returns 0, 1 (max_inner hit) or 2 (linear sweeps did not settle). */
int qpo_gmwb_solve(const qpo_gmwb *p, double *V, double *mean_inner, int *n_steps);
/* The same, cut into the pieces an A-axis shard owns: lines [lo, hi). */
typedef struct qpo_gmwb_state qpo_gmwb_state;
qpo_gmwb_state *qpo_gmwb_state_new(const qpo_gmwb *p);
void qpo_gmwb_state_free(qpo_gmwb_state *s);
void qpo_gmwb_exit(const qpo_gmwb *p, int lo, int hi, double *x);
void qpo_gmwb_policy(qpo_gmwb_state *s, const double *xprev, int lo, int hi);
double qpo_gmwb_sweep(qpo_gmwb_state *s, double *x, const double *v0, double h0, int lo, int hi);
double qpo_gmwb_relerr(const qpo_gmwb *p, const double *x, const double *xprev, int lo, int hi);

/* ---- 1-D impulse-control HJBQVI of examples/hjbqvi/exchange_rate.cpp (qpde_oracle_hjbqvi1d.c) ---- */
typedef struct {
	int n_x; const double *x;   /* refined spatial grid                              */
	int n_q; const double *q;   /* stochastic control grid (interest-rate differential) */
	int n_z; const double *z;   /* impulse control grid (the new state)              */
	double T; int timesteps;
	double rho, sigma, a, b, m, lambda, c;   /* exchange_rate.cpp:52-69               */
	double scaling_factor;      /* HJBQVI default 1e-2: penalty = 1 / (this * dt)    */
	double tolerance;           /* iteration_tolerance, 1e-6                         */
	int max_inner;
} qpo_hjbqvi1d;
/* V, Q, Z [n_x] out (controls NaN where the other control acts), mask [n_x] out; any of Q, Z, mask
 * may be NULL.  Returns 0, 1 (max_inner hit) or 2 (the lagged entries did not settle). */
int qpo_hjbqvi1d_solve(const qpo_hjbqvi1d *p, double *V, double *Q, double *Z, unsigned char *mask,
		double *mean_inner, int *n_steps);

/* ---- 2-D impulse-control HJBQVI of examples/hjbqvi/optimal_consumption.cpp (qpde_oracle_hjbqvi2d.c) ---- */
typedef struct {
	int n_w, n_b;
	const double *W, *B;        /* refined axes; node (iw, ib) is row iw + n_w ib      */
	int n_q; const double *q;   /* consumption-rate grid                             */
	int n_z; const double *zfrac; /* impulse control grid (fractions of the admissible transfer) */
	double T; int timesteps;
	double rho, r, mu, sigma, gamma, lambda, c, boundary;   /* optimal_consumption.cpp:36-80 */
	double scaling_factor, tolerance;
	int max_inner, max_sweeps;
} qpo_hjbqvi2d;
/* V, Q, Z [n_w n_b] out (controls NaN where the other control acts), mask out; Q, Z, mask may be NULL.
 * Returns 0, 1 (max_inner hit) or 2 (the line sweeps did not settle). */
int qpo_hjbqvi2d_solve(const qpo_hjbqvi2d *p, double *V, double *Q, double *Z, unsigned char *mask,
		double *mean_inner, int *n_steps);

#ifdef __cplusplus
}
#endif

#endif
