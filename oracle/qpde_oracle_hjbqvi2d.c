/* TEST INFRASTRUCTURE ONLY — see qpde_oracle.h.  Plain-C restatement of the reference's 2-D
 * impulse-control HJBQVI path with impulses that point both ways in the state:
 * examples/hjbqvi/optimal_consumption.cpp (optimal consumption with fixed and proportional
 * transaction costs; w = stock, b = bank) on Modules/HJBQVI/HJBQVI.hpp with usePenalizedScheme().
 *
 *   HJBQVI::solve wiring                  Modules/HJBQVI/HJBQVI.hpp:590-811
 *   ControlledOperator::A / b             :158-335   (no boundary routines: boundary rows carry no
 *                                                     derivative in that direction)
 *   Impulse::A / b                        Core/Impulse.hpp:87-216
 *   PolicyIteration::onIterationStart     Core/PolicyIteration.hpp:54-105
 *   PenaltyMethod (non-direct)            Core/PenaltyMethod.hpp:37-119,127-139
 *   BDFOne                                Core/BDF.hpp:32-46
 *   ToleranceIteration / relativeError    Core/IterativeMethod.hpp:1125-1254
 *   model lambdas                         examples/hjbqvi/optimal_consumption.cpp:85-176
 *
 * Parity pin: NOT bit-for-bit.  The reference solves each linear system with BiCGSTAB + ILUT to a
 * machine-epsilon residual (Bindings/Eigen.hpp:170-204); this file solves the same system by
 * symmetric block Gauss-Seidel over the w-lines (tridiagonal solves along w, everything else
 * taken from the running iterate) repeated to a 1e-15 relative update.  Checked against the
 * reference's own outputs (tests/golden/hjbqvi2d_reference.npz, made by qpde_ref consumption) to
 * 1e-9 relative on every node.
 */
#include "qpde_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

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

static double qmax(double x, double y) { return x > y ? x : y; }   /* EarlyInclude.hpp:48-49 */
static double qmin(double x, double y) { return x < y ? x : y; }

typedef struct { double aw, cw, ab, cb, diag, b; } op_row;

/* one dimension of the stencil (HJBQVI.hpp:224-274) */
static void stencil(const double *x, int i, int n, double v, double mu, double *lo, double *up, double *total) {
	*lo = 0.; *up = 0.;
	if(i == 0 || i == n - 1) return;                  /* no boundary routine: nothing in this direction */
	{
		const double dxb = x[i] - x[i - 1], dxc = x[i + 1] - x[i - 1], dxf = x[i + 1] - x[i];
		const double vv = v * v;
		const double alpha_common = vv / dxb / dxc, beta_common = vv / dxf / dxc;
		double alpha = alpha_common - mu / dxc, beta = beta_common + mu / dxc;
		if(alpha < 0.) { alpha = alpha_common; beta = beta_common + mu / dxf; }
		else if(beta < 0.) { alpha = alpha_common - mu / dxb; beta = beta_common; }
		*lo = -alpha; *up = -beta;
		*total += alpha + beta;
	}
}

/* ControlledOperator::A / ::b under the consumption rate q (optimal_consumption.cpp:118-141) */
static op_row controlled_row(const qpo_hjbqvi2d *p, int iw, int ib, double q, double flow_q) {
	op_row o;
	const double w = p->W[iw], b = p->B[ib];
	const int inside = 0 < b && b < p->boundary;
	double total = 0.;
	stencil(p->W, iw, p->n_w, p->sigma * w, 0. + p->mu * w, &o.aw, &o.cw, &total);
	stencil(p->B, ib, p->n_b, 0., 0. + (p->r * b + (inside ? -q : 0.)), &o.ab, &o.cb, &total);
	o.diag = total + p->rho;
	o.b = inside ? flow_q : 0.;
	return o;
}

typedef struct { int t[4]; double f[4]; double reward; } impulse_row;

static double z_bounded(const qpo_hjbqvi2d *p, double w, double b, double z_frac) {   /* :85-99 */
	const double tmp1 = b - p->c;
	const double tmp2 = tmp1 - p->boundary;
	const double lo = qmax(-w, (tmp2 > 0) ? (tmp2 / (1 - p->lambda)) : (tmp2 / (1 + p->lambda)));
	const double hi = qmin(p->boundary - w, (tmp1 > 0) ? (tmp1 / (1 + p->lambda)) : (tmp1 / (1 - p->lambda)));
	const double z = z_frac * (hi - lo) + lo;
	if(fabs(z) < DBL_EPSILON) return 0.;
	return z;
}

/* Impulse::A / ::b under z_frac (Impulse.hpp:87-216, optimal_consumption.cpp:144-162) */
static impulse_row impulse_of(const qpo_hjbqvi2d *p, int iw, int ib, double z_frac) {
	impulse_row m;
	const double w = p->W[iw], b = p->B[ib];
	const double z = z_bounded(p, w, b, z_frac);
	const double w_plus = w + z, b_plus = b - z - p->lambda * fabs(z) - p->c;
	int i0, i1; double f0, f1;
	interp_data(p->W, p->n_w, w_plus, &i0, &f0);
	interp_data(p->B, p->n_b, b_plus, &i1, &f1);
	m.t[0] = i0 + p->n_w * i1;           m.f[0] = 1. * f0 * f1;
	m.t[1] = i0 + 1 + p->n_w * i1;       m.f[1] = 1. * (-f0 + 1.) * f1;
	m.t[2] = i0 + p->n_w * (i1 + 1);     m.f[2] = 1. * f0 * (-f1 + 1.);
	m.t[3] = i0 + 1 + p->n_w * (i1 + 1); m.f[3] = 1. * (-f0 + 1.) * (-f1 + 1.);
	m.reward = z == 0. ? -INFINITY : 0.;
	return m;
}

static double impulse_candidate(const impulse_row *m, int row, const double *x) {
	double acc = 0.;
	int k, diag_done = 0;
	for(k = 0; k < 4; ++k) {
		if(!diag_done && row < m->t[k]) { acc += 1. * x[row]; diag_done = 1; }
		if(m->t[k] == row) { acc += (1. - m->f[k]) * x[row]; diag_done = 1; }
		else acc += (0. - m->f[k]) * x[m->t[k]];
	}
	if(!diag_done) acc += 1. * x[row];
	return acc - m->reward;
}

static double relative_error2(int n, const double *a, const double *b, double scale) {
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

static long last_sweeps, last_solves;
/* symmetric sweeps per linear solve of the last call (diagnostic) */
double qpo_hjbqvi2d_mean_sweeps(void) { return last_solves ? (double)last_sweeps / (double)last_solves : 0.; }

int qpo_hjbqvi2d_solve(const qpo_hjbqvi2d *p, double *V, double *Q, double *Z, unsigned char *mask,
		double *mean_inner, int *n_steps_out) {
	const int nw = p->n_w, nb = p->n_b, n = nw * nb;
	const double dt_nominal = p->T / p->timesteps;
	const double pen = 1. / (p->scaling_factor * dt_nominal);
	double *x = (double *)malloc(sizeof(double) * n), *xprev = (double *)malloc(sizeof(double) * n);
	double *v0 = (double *)malloc(sizeof(double) * n);
	double *la = (double *)malloc(sizeof(double) * nw), *lb = (double *)malloc(sizeof(double) * nw);
	double *lc = (double *)malloc(sizeof(double) * nw), *lr = (double *)malloc(sizeof(double) * nw);
	double *cp = (double *)malloc(sizeof(double) * nw), *dp = (double *)malloc(sizeof(double) * nw);
	double *flow_q = (double *)malloc(sizeof(double) * p->n_q);
	op_row *rows = (op_row *)malloc(sizeof(op_row) * n);
	impulse_row *imp = (impulse_row *)malloc(sizeof(impulse_row) * n);
	double *qsel = (double *)calloc(n, sizeof(double)), *zsel = (double *)calloc(n, sizeof(double));
	unsigned char *active = (unsigned char *)calloc(n, 1);
	qpo_step *steps = (qpo_step *)malloc(sizeof(qpo_step) * (size_t)(p->timesteps + 8));
	const int n_steps = qpo_schedule(p->T, dt_nominal, NULL, 0, steps, p->timesteps + 8);
	int s, i, k, status = 0; long inner_sum = 0;
	double time1 = p->T;

	last_sweeps = 0; last_solves = 0;
	for(k = 0; k < p->n_q; ++k) flow_q[k] = pow(p->q[k], p->gamma) / p->gamma;          /* :138-141 */
	for(i = 0; i < n; ++i) {                                                            /* :165-168 */
		const double w = p->W[i % nw], b = p->B[i / nw];
		const double liquidate = qmax(b + (1 - p->lambda) * w - p->c, 0.);
		v0[i] = pow(liquidate, p->gamma) / p->gamma;
	}
	for(s = 0; s < n_steps; ++s) {
		const double t = steps[s].t, h0 = time1 - t;
		int its = 0, notdone;
		memcpy(x, v0, sizeof(double) * n);
		do {
			int sweep;
			memcpy(xprev, x, sizeof(double) * n);
			for(i = 0; i < n; ++i) {
				const int iw = i % nw, ib = i / nw;
				double best = INFINITY;
				for(k = 0; k < p->n_q; ++k) {
					const op_row o = controlled_row(p, iw, ib, p->q[k], flow_q[k]);
					double acc = 0., cand;
					/* row entries in column order; an interior direction carries both of its entries */
					if(ib > 0 && ib < nb - 1) acc += o.ab * xprev[i - nw];
					if(iw > 0 && iw < nw - 1) acc += o.aw * xprev[i - 1];
					acc += o.diag * xprev[i];
					if(iw > 0 && iw < nw - 1) acc += o.cw * xprev[i + 1];
					if(ib > 0 && ib < nb - 1) acc += o.cb * xprev[i + nw];
					cand = acc - o.b;
					if(cand < best) { best = cand; rows[i] = o; qsel[i] = p->q[k]; }
				}
				best = INFINITY;
				for(k = 0; k < p->n_z; ++k) {
					const impulse_row m = impulse_of(p, iw, ib, p->zfrac[k]);
					const double cand = impulse_candidate(&m, i, xprev);
					if(cand < best) { best = cand; zsel[i] = p->zfrac[k]; }
				}
				if(!(best < INFINITY)) zsel[i] = 0.;     /* the zero-filled vector handed to setInputs */
				imp[i] = impulse_of(p, iw, ib, zsel[i]);
				active[i] = (pen * impulse_candidate(&imp[i], i, xprev)) < 0.;
			}
			/* (I + h0 L^q + P (I - M*)) x = v0 + h0 b^q + P reward*: symmetric line Gauss-Seidel */
			for(sweep = 0; sweep < p->max_sweeps; ++sweep) {
				double worst = 0.;
				int pass;
				for(pass = 0; pass < 2; ++pass) {
					int ll;
					for(ll = 0; ll < nb; ++ll) {
						const int ib = pass == 0 ? ll : nb - 1 - ll;
						int iw;
						double cprev = 0., dprev = 0., xv = 0.;
						for(iw = 0; iw < nw; ++iw) {
							const int row = iw + nw * ib;
							const op_row *o = &rows[row];
							double a_ = o->aw * h0, b_ = 1. + o->diag * h0, c_ = o->cw * h0;
							double r_ = v0[row] + h0 * o->b;
							if(ib > 0) r_ -= (o->ab * h0) * x[row - nw];
							if(ib < nb - 1) r_ -= (o->cb * h0) * x[row + nw];
							if(active[row]) {
								b_ += pen;
								r_ += pen * imp[row].reward;
								for(k = 0; k < 4; ++k) {
									const int tk = imp[row].t[k];
									const double f = pen * imp[row].f[k];
									if(tk == row) b_ -= f;
									else if(tk == row - 1 && iw > 0) a_ -= f;
									else if(tk == row + 1 && iw < nw - 1) c_ -= f;
									else r_ += f * x[tk];
								}
							}
							la[iw] = a_; lb[iw] = b_; lc[iw] = c_; lr[iw] = r_;
						}
						for(iw = 0; iw < nw; ++iw) {
							const double inv = 1. / (lb[iw] - la[iw] * cprev);
							cprev = lc[iw] * inv; dprev = (lr[iw] - la[iw] * dprev) * inv;
							cp[iw] = cprev; dp[iw] = dprev;
						}
						for(iw = nw - 1; iw >= 0; --iw) {
							const int row = iw + nw * ib;
							double d, den;
							xv = dp[iw] - cp[iw] * xv;
							d = fabs(xv - x[row]); den = fabs(xv) > 1. ? fabs(xv) : 1.;
							if(d / den > worst) worst = d / den;
							x[row] = xv;
						}
					}
				}
				if(worst <= 1e-15) break;
			}
			if(sweep >= p->max_sweeps) status = 2;
			last_sweeps += sweep + 1; ++last_solves;
			++its;
			notdone = relative_error2(n, x, xprev, 1.) > p->tolerance;
			if(notdone && its >= p->max_inner) { status = status ? status : 1; break; }
		} while(notdone);
		inner_sum += its;
		memcpy(v0, x, sizeof(double) * n);
		time1 = t;
	}
	memcpy(V, v0, sizeof(double) * n);
	for(i = 0; i < n; ++i) {
		const int act = impulse_candidate(&imp[i], i, v0) < 0.;
		if(mask) mask[i] = (unsigned char)act;
		if(Q) Q[i] = act ? NAN : qsel[i];
		if(Z) Z[i] = act ? zsel[i] : NAN;
	}
	if(mean_inner) *mean_inner = (double)inner_sum / n_steps;
	if(n_steps_out) *n_steps_out = n_steps;
	free(x); free(xprev); free(v0); free(la); free(lb); free(lc); free(lr); free(cp); free(dp); free(flow_q);
	free(rows); free(imp); free(qsel); free(zsel); free(active); free(steps);
	return status;
}
