/* TEST INFRASTRUCTURE ONLY — see qpde_oracle.h.  Plain-C restatement of the
 * reference's 1-D impulse-control HJBQVI path: examples/hjbqvi/exchange_rate.cpp
 * (Mundaca & Oksendal's exchange-rate intervention) on Modules/HJBQVI/HJBQVI.hpp
 * with usePenalizedScheme().
 *
 *   HJBQVI::solve wiring                  Modules/HJBQVI/HJBQVI.hpp:590-811
 *   ControlledOperator::A / b             :158-335   (no boundary routines: rows 0 and n-1 are [rho])
 *   Impulse::A / b                        Core/Impulse.hpp:87-216
 *   PolicyIteration::onIterationStart     Core/PolicyIteration.hpp:54-105
 *   PenaltyMethod (non-direct)            Core/PenaltyMethod.hpp:37-119,127-139
 *   BDFOne                                Core/BDF.hpp:32-46
 *   ToleranceIteration / relativeError    Core/IterativeMethod.hpp:1125-1254
 *   model lambdas                         examples/hjbqvi/exchange_rate.cpp:118-152
 *
 * Parity pin: NOT bit-for-bit.  The reference solves each linear system with
 * BiCGSTAB + ILUT to machine-epsilon residual (Bindings/Eigen.hpp:170-204); this
 * file solves the same system — a tridiagonal matrix plus the two interpolation
 * entries of every penalised row — by lagging the entries that are not on the
 * three diagonals and repeating the Thomas solve to a 1e-15 relative update.
 * Checked against the reference's own outputs (tests/golden/hjbqvi1d_reference.npz,
 * made by qpde_ref exchange) to 1e-9 relative on every node.
 */
#include "qpde_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* linearInterpolationData: Core/Interpolant.hpp:153-181 */
static void interp_data(const double *x, int length, double ci, int *idx, double *w) {
	int lo = 0, hi = length - 2, mid = 0;
	if(ci <= x[0]) { *idx = 0; *w = 1.; return; }
	if(ci >= x[length - 1]) { *idx = length - 2; *w = 0.; return; }
	*w = 0.;
	while(lo <= hi) {
		mid = (lo + hi) / 2;
		if(ci < x[mid]) hi = mid - 1;
		else if(ci >= x[mid + 1]) lo = mid + 1;
		else { *w = (x[mid + 1] - ci) / (x[mid + 1] - x[mid]); break; }
	}
	*idx = mid;
}

typedef struct { double lo, up, diag, b; } op_row;   /* one row of L^q and the flow b^q */

/* ControlledOperator::A / ::b, row i under the control q (HJBQVI.hpp:181-283, 303-332) */
static op_row controlled_row(const qpo_hjbqvi1d *p, int i, double q) {
	op_row o;
	const double xv = p->x[i];
	double total = 0.;
	o.lo = o.up = 0.;
	if(i > 0 && i < p->n_x - 1) {
		const double dxb = p->x[i] - p->x[i - 1], dxc = p->x[i + 1] - p->x[i - 1], dxf = p->x[i + 1] - p->x[i];
		const double v = p->sigma, vv = v * v;
		const double mu = 0. + -p->a * q;                                 /* exchange_rate.cpp:127 */
		const double alpha_common = vv / dxb / dxc, beta_common = vv / dxf / dxc;
		double alpha = alpha_common - mu / dxc, beta = beta_common + mu / dxc;
		if(alpha < 0.) { alpha = alpha_common; beta = beta_common + mu / dxf; }
		else if(beta < 0.) { alpha = alpha_common - mu / dxb; beta = beta_common; }
		o.lo = -alpha; o.up = -beta;
		total += alpha + beta;
	}
	o.diag = total + p->rho;
	{
		const double tmp = xv - p->m;                                     /* exchange_rate.cpp:130-133 */
		o.b = - tmp * tmp - q * q * p->b;
	}
	return o;
}

typedef struct { int idx; double w; double reward; } impulse_row;   /* M row: (idx, w), (idx + 1, -w + 1) */

/* Impulse::A / ::b, row i under the control x_new (Impulse.hpp:87-216, exchange_rate.cpp:136-149) */
static impulse_row impulse_of(const qpo_hjbqvi1d *p, int i, double x_new) {
	impulse_row m;
	const double xv = p->x[i];
	interp_data(p->x, p->n_x, x_new, &m.idx, &m.w);
	if(fabs(x_new - p->m) >= fabs(xv - p->m)) m.reward = -INFINITY;
	else m.reward = - p->lambda * fabs(x_new - xv) - p->c;
	return m;
}

/* ((I - M) x - reward)_row, the row of I - M accumulated in column order */
static double impulse_candidate(const impulse_row *m, int row, const double *x) {
	const int t[2] = { m->idx, m->idx + 1 };
	const double f[2] = { m->w, -m->w + 1. };
	double acc = 0.;
	int k, diag_done = 0;
	for(k = 0; k < 2; ++k) {
		if(!diag_done && row < t[k]) { acc += 1. * x[row]; diag_done = 1; }
		if(t[k] == row) { acc += (1. - f[k]) * x[row]; diag_done = 1; }
		else acc += (0. - f[k]) * x[t[k]];
	}
	if(!diag_done) acc += 1. * x[row];
	return acc - m->reward;
}

static double relative_error1(int n, const double *a, const double *b, double scale) {
	double m = 0.; int i;
	for(i = 0; i < n; ++i) {
		const double fa = fabs(a[i]), fb = fabs(b[i]);
		double den = fa < fb ? fb : fa, quo;
		den = scale < den ? den : scale;
		quo = fabs(a[i] - b[i]) / den;
		if(i == 0 || quo > m) m = quo;
	}
	return m;
}

int qpo_hjbqvi1d_solve(const qpo_hjbqvi1d *p, double *V, double *Q, double *Z, unsigned char *mask,
		double *mean_inner, int *n_steps_out) {
	const int n = p->n_x;
	const double dt_nominal = p->T / p->timesteps;
	const double pen = 1. / (p->scaling_factor * dt_nominal);        /* HJBQVI.hpp:635, PenaltyMethod.hpp:85 */
	double *x = (double *)malloc(sizeof(double) * n), *xprev = (double *)malloc(sizeof(double) * n);
	double *v0 = (double *)malloc(sizeof(double) * n), *xn = (double *)malloc(sizeof(double) * n);
	double *la = (double *)malloc(sizeof(double) * n), *lb = (double *)malloc(sizeof(double) * n);
	double *lc = (double *)malloc(sizeof(double) * n), *lr = (double *)malloc(sizeof(double) * n);
	double *cp = (double *)malloc(sizeof(double) * n), *dp = (double *)malloc(sizeof(double) * n);
	op_row *rows = (op_row *)malloc(sizeof(op_row) * n);
	impulse_row *imp = (impulse_row *)malloc(sizeof(impulse_row) * n);
	double *qsel = (double *)calloc(n, sizeof(double)), *zsel = (double *)calloc(n, sizeof(double));
	unsigned char *active = (unsigned char *)calloc(n, 1);
	qpo_step *steps = (qpo_step *)malloc(sizeof(qpo_step) * (size_t)(p->timesteps + 8));
	const int n_steps = qpo_schedule(p->T, dt_nominal, NULL, 0, steps, p->timesteps + 8);
	int s, i, status = 0; long inner_sum = 0;
	double time1 = p->T;

	for(i = 0; i < n; ++i) v0[i] = 0.;                                /* exit function, exchange_rate.cpp:152 */
	for(s = 0; s < n_steps; ++s) {
		const double t = steps[s].t, h0 = time1 - t;
		int its = 0, notdone;
		memcpy(x, v0, sizeof(double) * n);
		do {
			int sweep, k;
			memcpy(xprev, x, sizeof(double) * n);
			for(i = 0; i < n; ++i) {
				double best = INFINITY;
				/* 1. stochastic policy: argmin_q (A(q) x - b(q))_i, strict <; row entries in column order */
				for(k = 0; k < p->n_q; ++k) {
					const op_row o = controlled_row(p, i, p->q[k]);
					double acc = 0., cand;
					if(i > 0 && i < n - 1) acc += o.lo * xprev[i - 1];
					acc += o.diag * xprev[i];
					if(i > 0 && i < n - 1) acc += o.up * xprev[i + 1];
					cand = acc - o.b;
					if(cand < best) { best = cand; rows[i] = o; qsel[i] = p->q[k]; }
				}
				/* 2. impulse policy: argmin_z ((I - M_z) x - reward_z)_i, strict <.  A node where every
				 * candidate is pruned keeps its control of the previous iteration (zero at first:
				 * the Vector the reference hands to setInputs is the shim's zero-filled one) */
				best = INFINITY;
				for(k = 0; k < p->n_z; ++k) {
					const impulse_row m = impulse_of(p, i, p->z[k]);
					const double cand = impulse_candidate(&m, i, xprev);
					if(cand < best) { best = cand; zsel[i] = p->z[k]; }
				}
				if(!(best < INFINITY)) zsel[i] = 0.;
				imp[i] = impulse_of(p, i, zsel[i]);
				/* 3. penalty: active iff scale * ((I - M*) x - reward*) < 0 */
				active[i] = (pen * impulse_candidate(&imp[i], i, xprev)) < 0.;
			}
			/* 4. (I + h0 L^q + P (I - M*)) x = v0 + h0 b^q + P reward*: entries off the three diagonals lagged */
			for(sweep = 0; sweep < 500; ++sweep) {
				double worst = 0., cprev = 0., dprev = 0., xv = 0.;
				for(i = 0; i < n; ++i) {
					const op_row *o = &rows[i];
					double a_ = o->lo * h0, b_ = 1. + o->diag * h0, c_ = o->up * h0;
					double r_ = v0[i] + h0 * o->b;
					if(active[i]) {
						const int t[2] = { imp[i].idx, imp[i].idx + 1 };
						const double f[2] = { pen * imp[i].w, pen * (-imp[i].w + 1.) };
						b_ += pen;
						r_ += pen * imp[i].reward;
						for(k = 0; k < 2; ++k) {
							if(t[k] == i) b_ -= f[k];
							else if(t[k] == i - 1) a_ -= f[k];
							else if(t[k] == i + 1) c_ -= f[k];
							else r_ += f[k] * x[t[k]];
						}
					}
					la[i] = a_; lb[i] = b_; lc[i] = c_; lr[i] = r_;
				}
				for(i = 0; i < n; ++i) {
					const double inv = 1. / (lb[i] - la[i] * cprev);
					cprev = lc[i] * inv; dprev = (lr[i] - la[i] * dprev) * inv;
					cp[i] = cprev; dp[i] = dprev;
				}
				for(i = n - 1; i >= 0; --i) { xv = dp[i] - cp[i] * xv; xn[i] = xv; }
				for(i = 0; i < n; ++i) {
					const double d = fabs(xn[i] - x[i]);
					const double den = fabs(xn[i]) > 1. ? fabs(xn[i]) : 1.;
					if(d / den > worst) worst = d / den;
					x[i] = xn[i];
				}
				if(worst <= 1e-15) break;
			}
			if(sweep >= 500) status = 2;
			++its;
			notdone = relative_error1(n, x, xprev, 1.) > p->tolerance;
			if(notdone && its >= p->max_inner) { status = status ? status : 1; break; }
		} while(notdone);
		inner_sum += its;
		memcpy(v0, x, sizeof(double) * n);
		time1 = t;
	}
	memcpy(V, v0, sizeof(double) * n);
	/* the controls of the last iteration; PenaltyMethod::constraintMask on the final iterand (:127-139);
	 * NaN where the other control acts (HJBQVI.hpp:977-990) */
	for(i = 0; i < n; ++i) {
		const int act = impulse_candidate(&imp[i], i, v0) < 0.;
		if(mask) mask[i] = (unsigned char)act;
		if(Q) Q[i] = act ? NAN : qsel[i];
		if(Z) Z[i] = act ? zsel[i] : NAN;
	}
	if(mean_inner) *mean_inner = (double)inner_sum / n_steps;
	if(n_steps_out) *n_steps_out = n_steps;
	free(x); free(xprev); free(v0); free(xn); free(la); free(lb); free(lc); free(lr); free(cp); free(dp);
	free(rows); free(imp); free(qsel); free(zsel); free(active); free(steps);
	return status;
}
