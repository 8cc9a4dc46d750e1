/* TEST INFRASTRUCTURE ONLY — see qpde_oracle.h.  Plain-C restatement of the
 * reference's GMWB path: examples/hjbqvi/gmwb.cpp on
 * Modules/HJBQVI/HJBQVI.hpp with usePenalizedScheme().
 *
 *   HJBQVI::solve wiring                  Modules/HJBQVI/HJBQVI.hpp:590-811
 *   ControlledOperator::A / b             :158-335
 *   HJBQVILinearBoundary                  :1285-1306
 *   HJBQVIZeroDiffusionRightBoundary      :1313-1340
 *   Impulse::A / b                        Core/Impulse.hpp:87-216
 *   PolicyIteration::onIterationStart     Core/PolicyIteration.hpp:54-105
 *   PenaltyMethod (non-direct)            Core/PenaltyMethod.hpp:37-119
 *   BDFOne                                Core/BDF.hpp:32-46
 *   ToleranceIteration / relativeError    Core/IterativeMethod.hpp:1125-1254
 *   model lambdas                         examples/hjbqvi/gmwb.cpp:92-176
 *
 * Parity pin: NOT bit-for-bit.  The reference solves each linear system with
 * BiCGSTAB + ILUT to machine-epsilon residual (Bindings/Eigen.hpp:170-204);
 * this file solves the same system with block Gauss-Seidel over the A-lines
 * (tridiagonal W-solves, couplings to lower A-lines are exact because every
 * other entry of the matrix points to an equal-or-lower A) iterated to a
 * 1e-15 relative update.  Checked against the reference's own outputs
 * (tests/golden/gmwb_reference.npz) to 1e-9 relative on every node.
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

static double qmax(double x, double y) { return x > y ? x : y; }   /* EarlyInclude.hpp:48 */

typedef struct {
	double aw, cw, aa, ca, diag, b;   /* W sub/super, A sub/super, diagonal of L^q, flow b^q */
} op_row;

/* One row of ControlledOperator::A and ::b for control q (HJBQVI.hpp:158-335) */
static op_row controlled_row(const qpo_gmwb *p, int iw, int ia, double q) {
	op_row o;
	const double w = p->W[iw], a = p->A[ia];
	double total = 0.;
	o.aw = o.cw = o.aa = o.ca = 0.;
	/* d = 0: W */
	{
		const double mu = (p->r - p->eta) * w + ((0. < a) ? -q : 0.);     /* gmwb.cpp:122-124 */
		if(iw == 0) {
			/* no left boundary routine */
		} else if(iw == p->n_w - 1) {
			total += - mu / w;                                            /* LinearBoundary */
		} else {
			const double dxb = p->W[iw] - p->W[iw - 1], dxc = p->W[iw + 1] - p->W[iw - 1];
			const double dxf = p->W[iw + 1] - p->W[iw];
			const double v = p->sigma * w, vv = v * v;
			const double alpha_common = vv / dxb / dxc, beta_common = vv / dxf / dxc;
			double alpha = alpha_common - mu / dxc, beta = beta_common + mu / dxc;
			if(alpha < 0.) { alpha = alpha_common; beta = beta_common + mu / dxf; }
			else if(beta < 0.) { alpha = alpha_common - mu / dxb; beta = beta_common; }
			o.aw = -alpha; o.cw = -beta;
			total += alpha + beta;
		}
	}
	/* d = 1: A (zero volatility) */
	{
		const double mu = (0. < a) ? -q : 0.;                             /* gmwb.cpp:127-129 */
		if(ia == 0) {
		} else if(ia == p->n_a - 1) {
			const double dxb = p->A[ia] - p->A[ia - 1];
			const double alpha = - mu / dxb;                              /* ZeroDiffusionRightBoundary */
			o.aa = -alpha;
			total += alpha;
		} else {
			const double dxb = p->A[ia] - p->A[ia - 1], dxc = p->A[ia + 1] - p->A[ia - 1];
			const double dxf = p->A[ia + 1] - p->A[ia];
			const double vv = 0. * 0.;
			const double alpha_common = vv / dxb / dxc, beta_common = vv / dxf / dxc;
			double alpha = alpha_common - mu / dxc, beta = beta_common + mu / dxc;
			if(alpha < 0.) { alpha = alpha_common; beta = beta_common + mu / dxf; }
			else if(beta < 0.) { alpha = alpha_common - mu / dxb; beta = beta_common; }
			o.aa = -alpha; o.ca = -beta;
			total += alpha + beta;
		}
	}
	o.diag = total + p->r;
	o.b = (0. < a) ? q : 0.;                                              /* gmwb.cpp:134-136 */
	return o;
}

typedef struct {
	int t[4];        /* target rows, ascending column order */
	double f[4];     /* bilinear factors */
	double reward;   /* impulse flow; -inf when the impulse does not move the state */
} impulse_row;

/* One row of Impulse::A / ::b for the control z_frac (Impulse.hpp:87-216, gmwb.cpp:139-165) */
static impulse_row impulse_of(const qpo_gmwb *p, int iw, int ia, double z_frac) {
	impulse_row m;
	const double w = p->W[iw], a = p->A[ia];
	const double z = z_frac * a;
	const double w_plus = qmax(w - z, 0.), a_plus = a - z;
	int i0, i1; double f0, f1;
	interp_data(p->W, p->n_w, w_plus, &i0, &f0);
	interp_data(p->A, p->n_a, a_plus, &i1, &f1);
	/* corners in ascending row index: (i0,i1), (i0+1,i1), (i0,i1+1), (i0+1,i1+1) */
	m.t[0] = i0 + p->n_w * i1;           m.f[0] = 1. * f0 * f1;
	m.t[1] = i0 + 1 + p->n_w * i1;       m.f[1] = 1. * (-f0 + 1.) * f1;
	m.t[2] = i0 + p->n_w * (i1 + 1);     m.f[2] = 1. * f0 * (-f1 + 1.);
	m.t[3] = i0 + 1 + p->n_w * (i1 + 1); m.f[3] = 1. * (-f0 + 1.) * (-f1 + 1.);
	if((w - w_plus) + (a - a_plus) < DBL_EPSILON) m.reward = -INFINITY;
	else m.reward = (1 - p->kfee) * z - p->c;
	return m;
}

/* ((I - M) x - reward)_row with the row of I - M accumulated in column order */
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

/* Thomas solve of one W-line (no pivoting; diagonally dominant M-matrix rows) */
static void line_solve(int n, const double *a, const double *b, const double *c,
		const double *r, double *x, double *cp, double *dp) {
	int i; double cprev = 0., dprev = 0., xn = 0.;
	for(i = 0; i < n; ++i) {
		const double inv = 1. / (b[i] - a[i] * cprev);
		cprev = c[i] * inv; dprev = (r[i] - a[i] * dprev) * inv;
		cp[i] = cprev; dp[i] = dprev;
	}
	for(i = n - 1; i >= 0; --i) { xn = dp[i] - cp[i] * xn; x[i] = xn; }
}

/* ------------------------------------------------------------------------ */
/* The pieces of one inner iteration, restricted to the A-lines [lo, hi): the  */
/* unit an A-axis shard owns.  The monolithic solve below    */
/* calls them with [0, n_a).                                                   */
/* ------------------------------------------------------------------------ */

struct qpo_gmwb_state {
	const qpo_gmwb *p;
	op_row *rows; impulse_row *imp; unsigned char *active;
	double *la, *lb, *lc, *lr, *cp, *dp, *xl;
	double pen;
};

qpo_gmwb_state *qpo_gmwb_state_new(const qpo_gmwb *p) {
	const int n = p->n_w * p->n_a, nw = p->n_w;
	qpo_gmwb_state *s = (qpo_gmwb_state *)malloc(sizeof(qpo_gmwb_state));
	s->p = p;
	s->rows = (op_row *)malloc(sizeof(op_row) * n);
	s->imp = (impulse_row *)malloc(sizeof(impulse_row) * n);
	s->active = (unsigned char *)calloc(n, 1);
	s->la = (double *)malloc(sizeof(double) * nw); s->lb = (double *)malloc(sizeof(double) * nw);
	s->lc = (double *)malloc(sizeof(double) * nw); s->lr = (double *)malloc(sizeof(double) * nw);
	s->cp = (double *)malloc(sizeof(double) * nw); s->dp = (double *)malloc(sizeof(double) * nw);
	s->xl = (double *)malloc(sizeof(double) * nw);
	s->pen = 1. / (p->scaling_factor * (p->T / p->timesteps));   /* HJBQVI.hpp:635, PenaltyMethod.hpp:85 */
	return s;
}

void qpo_gmwb_state_free(qpo_gmwb_state *s) {
	free(s->rows); free(s->imp); free(s->active);
	free(s->la); free(s->lb); free(s->lc); free(s->lr); free(s->cp); free(s->dp); free(s->xl);
	free(s);
}

/* exit values V(T) on the lines [lo, hi) (gmwb.cpp:168) */
void qpo_gmwb_exit(const qpo_gmwb *p, int lo, int hi, double *x) {
	int i;
	for(i = lo * p->n_w; i < hi * p->n_w; ++i) {
		const double w = p->W[i % p->n_w], a = p->A[i / p->n_w];
		x[i] = qmax(w, (1 - p->kfee) * a - p->c);
	}
}

/* steps 1-3: control searches and penalty flags for the rows of lines [lo, hi),
 * reading the whole previous iterate */
void qpo_gmwb_policy(qpo_gmwb_state *s, const double *xprev, int lo, int hi) {
	const qpo_gmwb *p = s->p;
	const int nw = p->n_w, na = p->n_a;
	int i;
	for(i = lo * nw; i < hi * nw; ++i) {
		const int iw = i % nw, ia = i / nw;
		double best = INFINITY; int k, has = 0;
		/* 1. stochastic policy: argmin_q (A(q) x - b(q))_i, strict < */
		for(k = 0; k < p->n_q; ++k) {
			const op_row o = controlled_row(p, iw, ia, p->q[k]);
			double acc = 0., cand;
			/* row entries in column order: A-sub, W-sub, diag, W-super, A-super;
			 * interior A rows carry both A entries even when they are zero */
			if(ia > 0) acc += o.aa * xprev[i - nw];
			if(iw > 0 && iw < nw - 1) acc += o.aw * xprev[i - 1];
			acc += o.diag * xprev[i];
			if(iw > 0 && iw < nw - 1) acc += o.cw * xprev[i + 1];
			if(ia > 0 && ia < na - 1) acc += o.ca * xprev[i + nw];
			cand = acc - o.b;
			if(cand < best) { best = cand; s->rows[i] = o; }
		}
		/* 2. impulse policy: argmin_z ((I - M_z) x - reward_z)_i, strict < */
		best = INFINITY;
		for(k = 0; k < p->n_z; ++k) {
			const impulse_row m = impulse_of(p, iw, ia, p->zfrac[k]);
			const double cand = impulse_candidate(&m, i, xprev);
			if(cand < best) { best = cand; s->imp[i] = m; has = 1; }
		}
		/* 3. penalty: active iff scale * ((I - M*) x - reward*) < 0 */
		s->active[i] = has && (s->pen * best) < 0.;
	}
}

/* step 4, one block Gauss-Seidel sweep over the lines [lo, hi) of
 * (I + h0 L^q + P (I - M*)) x = v0 + h0 b^q + P reward*; lines below lo must
 * already hold this sweep's values.  Returns the largest relative update. */
double qpo_gmwb_sweep(qpo_gmwb_state *s, double *x, const double *v0, double h0, int lo, int hi) {
	const qpo_gmwb *p = s->p;
	const int nw = p->n_w, na = p->n_a;
	const double pen = s->pen;
	double worst = 0.;
	int ia;
	for(ia = lo; ia < hi; ++ia) {
		int iw;
		for(iw = 0; iw < nw; ++iw) {
			const int row = iw + nw * ia;
			const op_row *o = &s->rows[row];
			double a_ = o->aw * h0, b_ = 1. + o->diag * h0, c_ = o->cw * h0;
			double r_ = v0[row] + h0 * o->b;
			if(ia > 0) r_ -= (o->aa * h0) * x[row - nw];
			if(ia < na - 1) r_ -= (o->ca * h0) * x[row + nw];
			if(s->active[row]) {
				int k;
				b_ += pen;
				r_ += pen * s->imp[row].reward;
				for(k = 0; k < 4; ++k) {
					const int tk = s->imp[row].t[k];
					const double f = pen * s->imp[row].f[k];
					if(tk == row) b_ -= f;
					else if(tk == row - 1) a_ -= f;
					else if(tk == row + 1) c_ -= f;
					else r_ += f * x[tk];
				}
			}
			s->la[iw] = a_; s->lb[iw] = b_; s->lc[iw] = c_; s->lr[iw] = r_;
		}
		line_solve(nw, s->la, s->lb, s->lc, s->lr, s->xl, s->cp, s->dp);
		for(iw = 0; iw < nw; ++iw) {
			const int row = iw + nw * ia;
			const double d = fabs(s->xl[iw] - x[row]);
			const double den = fabs(s->xl[iw]) > 1. ? fabs(s->xl[iw]) : 1.;
			if(d / den > worst) worst = d / den;
			x[row] = s->xl[iw];
		}
	}
	return worst;
}

/* relativeError over the lines [lo, hi) (Core/IterativeMethod.hpp:1125-1137) */
double qpo_gmwb_relerr(const qpo_gmwb *p, const double *x, const double *xprev, int lo, int hi) {
	const int first = lo * p->n_w, count = (hi - lo) * p->n_w;
	return count > 0 ? relative_error2(count, x + first, xprev + first, 1.) : 0.;
}

int qpo_gmwb_solve(const qpo_gmwb *p, double *V, double *mean_inner, int *n_steps_out) {
	const int nw = p->n_w, na = p->n_a, n = nw * na;
	const double dt_nominal = p->T / p->timesteps;
	double *x = (double *)malloc(sizeof(double) * n), *xprev = (double *)malloc(sizeof(double) * n);
	double *v0 = (double *)malloc(sizeof(double) * n);
	qpo_gmwb_state *st = qpo_gmwb_state_new(p);
	qpo_step *steps = (qpo_step *)malloc(sizeof(qpo_step) * (size_t)(p->timesteps + 8));
	const int n_steps = qpo_schedule(p->T, dt_nominal, NULL, 0, steps, p->timesteps + 8);
	int s, status = 0; long inner_sum = 0;
	double time1 = p->T;

	qpo_gmwb_exit(p, 0, na, v0);
	for(s = 0; s < n_steps; ++s) {
		const double t = steps[s].t, h0 = time1 - t;
		int its = 0, notdone;
		memcpy(x, v0, sizeof(double) * n);
		do {
			int sweep;
			memcpy(xprev, x, sizeof(double) * n);
			qpo_gmwb_policy(st, xprev, 0, na);
			for(sweep = 0; sweep < 500; ++sweep) {
				if(qpo_gmwb_sweep(st, x, v0, h0, 0, na) <= 1e-15) break;
			}
			if(sweep >= 500) status = 2;
			++its;
			notdone = qpo_gmwb_relerr(p, x, xprev, 0, na) > p->tolerance;
			if(notdone && its >= p->max_inner) { status = status ? status : 1; break; }
		} while(notdone);
		inner_sum += its;
		memcpy(v0, x, sizeof(double) * n);
		time1 = t;
	}
	memcpy(V, v0, sizeof(double) * n);
	if(mean_inner) *mean_inner = (double)inner_sum / n_steps;
	if(n_steps_out) *n_steps_out = n_steps;
	free(x); free(xprev); free(v0); free(steps);
	qpo_gmwb_state_free(st);
	return status;
}
