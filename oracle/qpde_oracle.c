/* TEST INFRASTRUCTURE ONLY — see qpde_oracle.h.  Plain-C restatement of the
 * reference's 1-D sweep.  Compile WITHOUT -ffast-math / -march=native and with
 * -ffp-contract=off: the reference's Release build is plain -O3 x86-64
 * (CMakeLists.txt:9,29), i.e. no fused multiply-add.
 *
 * Citations are file:line under /root/reference/QuantPDE/src.
 */
#include "qpde_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* Grid generation                                                           */
/* ------------------------------------------------------------------------ */

/* Core/Axis.hpp:445-459 */
int qpo_axis_special(double *out) {
	static const double ticks[33] = {
		0., .1, .2, .3, .4, .5, .6, .7, .80, .84, .88, .92, .94, .96, .98,
		1., 1.02, 1.04, 1.06, 1.08, 1.10, 1.14, 1.18, 1.23, 1.30, 1.40,
		1.50, 1.75, 2.25, 3.00, 7.50, 20., 100.
	};
	memcpy(out, ticks, sizeof(ticks));
	return 33;
}

/* Core/Axis.hpp:270-313: merge of two strictly increasing tick arrays,
 * exact-equality duplicates dropped. */
int qpo_axis_union(const double *a, int na, const double *b, int nb,
		double *out) {
	int i = 0, j = 0, k = 0;
	while(i < na && j < nb) {
		if(a[i] == b[j])      { out[k++] = a[i++]; ++j; }
		else if(a[i] < b[j])  { out[k++] = a[i++]; }
		else                  { out[k++] = b[j++]; }
	}
	while(i < na) out[k++] = a[i++];
	while(j < nb) out[k++] = b[j++];
	return k;
}

/* Core/Domain.hpp:1092-1151: tmp = 2^times sub-intervals per interval,
 * new ticks k * dx + left with dx = (right - left) / tmp. */
int qpo_axis_refine(const double *in, int n, int times, double *out) {
	int tmp = 1, i, j, k;
	if(times <= 0 || n < 2) {
		if(out) memcpy(out, in, sizeof(double) * (size_t)n);
		return n;
	}
	for(i = 0; i < times; ++i) tmp *= 2;
	if(!out) return tmp * (n - 1) + 1;
	out[0] = in[0];
	j = 1;
	for(i = 1; i < n; ++i) {
		const double dx = (in[i] - in[i - 1]) / tmp;
		for(k = 1; k < tmp; ++k) out[j++] = k * dx + in[i - 1];
		out[j++] = in[i];
	}
	return j;
}

/* ------------------------------------------------------------------------ */
/* Time loop replay: Core/IterativeMethod.hpp:1259-1317, Core/Stepper.hpp:16 */
/* ------------------------------------------------------------------------ */

static int cmp_desc(const void *a, const void *b) {
	const double x = *(const double *)a, y = *(const double *)b;
	return x < y ? 1 : (x > y ? -1 : 0);
}

int qpo_schedule(double T, double dt_const, const double *event_times,
		int n_events, qpo_step *out, int max_steps) {
	/* Reverse time: the priority queue pops the largest time first
	 * (IterativeMethod.hpp:1340-1358); a NullEvent sits at t = 0 (:1268-1276). */
	double *ev = (double *)malloc(sizeof(double) * (size_t)(n_events + 1));
	int n = 0, e = 0, i;
	double t = T, dt = -1., dtPrevious;
	for(i = 0; i < n_events; ++i) ev[i] = event_times[i];
	qsort(ev, (size_t)n_events, sizeof(double), cmp_desc);

	do {
		/* top of the queue: the largest remaining user event time, or the
		 * NullEvent at 0 */
		const double nextEventTime = e < n_events ? ev[e] : 0.;
		int users = 0;
		do {
			double tmp;
			dtPrevious = dt;
			dt = dt_const;
			tmp = t + -1. * dt;
			if(fabs(tmp - nextEventTime) < DBL_EPSILON) {
				tmp = nextEventTime;
			} else if(tmp < nextEventTime) {
				tmp = nextEventTime;
				dt = -1. * (nextEventTime - t);
			}
			t = tmp;
			if(n < max_steps) {
				out[n].t = t;
				out[n].dt = dt;
				out[n].same_dt = (dt == dtPrevious);
				out[n].event_after = 0;
				out[n].user_events = 0;
			}
			++n;
		} while(nextEventTime < t + -1. * DBL_EPSILON);
		t = nextEventTime;
		/* pop every event at exactly this time (:1283-1288) */
		while(e < n_events && ev[e] == t) { ++e; ++users; }
		if(n - 1 < max_steps) {
			out[n - 1].event_after = 1;
			out[n - 1].user_events = users;
		}
	} while(0. < t);

	free(ev);
	return n;
}

static int cmp_asc(const void *a, const void *b) {
	const double x = *(const double *)a, y = *(const double *)b;
	return x < y ? -1 : (x > y ? 1 : 0);
}

/* TimeIteration<true> (IterativeMethod.hpp:1494-1495) driven by ForwardConstantStepper
 * (Stepper.hpp:41): direction = +1, Order = std::greater, so the queue pops the SMALLEST time
 * first and, among events at one time, the one add()ed first (TimeOrder, :1340-1352); the
 * NullEvent (id UINT_MAX) sits at the terminal time T and goes last there. */
int qpo_schedule_forward(double T, double dt_const, const double *event_times,
		int n_events, qpo_step *out, int max_steps) {
	double *ev = (double *)malloc(sizeof(double) * (size_t)(n_events + 1));
	int n = 0, e = 0, i;
	double t = 0., dt = -1., dtPrevious;
	for(i = 0; i < n_events; ++i) ev[i] = event_times[i];
	qsort(ev, (size_t)n_events, sizeof(double), cmp_asc);

	do {
		/* top of the queue: the smallest remaining user event time, or the NullEvent at T if
		 * that comes first */
		const double nextEventTime = (e < n_events && ev[e] <= T) ? ev[e] : T;
		int users = 0;
		do {
			double tmp;
			dtPrevious = dt;
			dt = dt_const;
			tmp = t + 1. * dt;
			if(fabs(tmp - nextEventTime) < DBL_EPSILON) {
				tmp = nextEventTime;
			} else if(tmp > nextEventTime) {
				tmp = nextEventTime;
				dt = 1. * (nextEventTime - t);
			}
			t = tmp;
			if(n < max_steps) {
				out[n].t = t;
				out[n].dt = dt;
				out[n].same_dt = (dt == dtPrevious);
				out[n].event_after = 0;
				out[n].user_events = 0;
			}
			++n;
		} while(nextEventTime > t + 1. * DBL_EPSILON);
		t = nextEventTime;
		while(e < n_events && ev[e] == t) { ++e; ++users; }
		if(n - 1 < max_steps) {
			out[n - 1].event_after = 1;
			out[n - 1].user_events = users;
		}
	} while(T > t);

	free(ev);
	return n;
}

/* ------------------------------------------------------------------------ */
/* Operator rows: Modules/Operators/BlackScholes.hpp:136-226                 */
/* ------------------------------------------------------------------------ */

void qpo_bs_coeffs(int n, const double *S, const double *r, const double *v,
		const double *q, const double *lambda, double kappa,
		double *lo, double *di, double *up) {
	int i;
	lo[0] = 0.; di[0] = r[0]; up[0] = 0.;                 /* :173-176 */
	lo[n - 1] = 0.; di[n - 1] = q[n - 1]; up[n - 1] = 0.; /* :177-180 */
	for(i = 1; i < n - 1; ++i) {
		const double r_i = r[i], v_i = v[i], q_i = q[i];
		const double l_i = lambda ? lambda[i] : 0.;
		const double dSb = S[i] - S[i - 1];
		const double dSc = S[i + 1] - S[i - 1];
		const double dSf = S[i + 1] - S[i];
		const double tmp1 = v_i * v_i * S[i] * S[i];
		const double tmp2 = (r_i - q_i - l_i * kappa) * S[i];
		const double alpha_common = tmp1 / dSb / dSc;
		const double beta_common = tmp1 / dSf / dSc;
		double alpha_i = alpha_common - tmp2 / dSc;
		double beta_i = beta_common + tmp2 / dSc;
		if(alpha_i < 0.) {
			alpha_i = alpha_common;
			beta_i = beta_common + tmp2 / dSf;
		} else if(beta_i < 0.) {
			alpha_i = alpha_common - tmp2 / dSb;
			beta_i = beta_common;
		}
		lo[i] = -alpha_i;
		di[i] = alpha_i + beta_i + r_i + l_i;
		up[i] = -beta_i;
	}
}

/* ------------------------------------------------------------------------ */
/* Interpolation: Core/Interpolant.hpp:153-181 and :331-374                  */
/* ------------------------------------------------------------------------ */

double qpo_interp(int n, const double *x, const double *v, double c) {
	int idx = 0;
	double w = 0., sum = 0.;
	if(c <= x[0]) { idx = 0; w = 1.; }
	else if(c >= x[n - 1]) { idx = n - 2; w = 0.; }
	else {
		int lo = 0, hi = n - 2, mid = 0;
		while(lo <= hi) {
			mid = (lo + hi) / 2;
			if(c < x[mid]) hi = mid - 1;
			else if(c >= x[mid + 1]) lo = mid + 1;
			else { w = (x[mid + 1] - c) / (x[mid + 1] - x[mid]); break; }
		}
		idx = mid;
	}
	/* corners with factor <= epsilon are skipped (:369-371) */
	if(w > DBL_EPSILON) sum += w * v[idx];
	if(1. - w > DBL_EPSILON) sum += (1. - w) * v[idx + 1];
	return sum;
}

/* ------------------------------------------------------------------------ */
/* Quadrature: AdaptiveQuadrature1<TrapezoidalRule1<1>> with the expanding   */
/* window for infinite bounds (Core/Integral.hpp:122-212, 243-256, 367-456)  */
/* ------------------------------------------------------------------------ */

typedef double (*qpo_fn)(double, const void *);
static const double QPO_TOL = 1e-6;   /* QuantPDE::tolerance, EarlyInclude.hpp:62 */

static double trap1(qpo_fn f, const void *ctx, double a, double x) {
	/* TrapezoidalRule<1,1>::compute (:280-296) */
	double scale = 1. / 2, sum;
	scale *= (x - a) / 1;
	sum = f(a, ctx) * 1.;
	sum += f(x, ctx) * 1.;
	return scale * sum;
}

static double quad_refine(qpo_fn f, const void *ctx, double previous, double a,
		double x, int n) {
	/* AdaptiveQuadrature::refine for Dimension = 1 (:386-449) */
	const double mid = (a + x) / 2.;
	const double left = trap1(f, ctx, a, mid);
	const double right = trap1(f, ctx, mid, x);
	double sum = 0., error;
	sum += left;
	sum += right;
	error = fabs((sum - previous) / sum);
	if(n > 0 && error > QPO_TOL) {
		sum = 0.;
		sum += quad_refine(f, ctx, left, a, mid, n - 1);
		sum += quad_refine(f, ctx, right, mid, x, n - 1);
	}
	return sum;
}

static double quad_compute(qpo_fn f, const void *ctx, double a, double x) {
	return quad_refine(f, ctx, trap1(f, ctx, a, x), a, x, 64);
}

static double quad_integral(qpo_fn f, const void *ctx, double a, double x) {
	/* Integral<1>::operator() (:122-212) */
	const int neginf = a == -INFINITY, posinf = x == INFINITY;
	double aa = a, xx = x, m = 0., previous, integral;
	if(neginf) {
		if(posinf) { aa = -1.; xx = 1.; m = 0.; }
		else       { aa = xx - 1.; m = xx; }
	} else if(posinf) { xx = aa + 1.; m = aa; }
	if(!neginf && !posinf) return quad_compute(f, ctx, aa, xx);
	previous = quad_compute(f, ctx, aa, xx);
	for(;;) {
		if(neginf) aa += aa - m;
		if(posinf) xx += xx - m;
		integral = quad_compute(f, ctx, aa, xx);
		if(fabs((integral - previous) / integral) < QPO_TOL) break;
		previous = integral;
	}
	return integral;
}

/* Jump amplitude densities (Modules/Lambdas/Densities.hpp:21-66) */
typedef struct { int kind; double p0, p1, p2; } qpo_density;

static double density_eval(double x, const qpo_density *d) {
	if(x == 0.) return 0.;
	if(d->kind == QPO_DENSITY_LOGNORMAL) {
		const double mu = d->p0, sigma = d->p1;
		return 1. / (x * sigma * sqrt(2. * M_PI))
			* exp( -((log(x) - mu) * (log(x) - mu)) / (2. * sigma * sigma) );
	} else {
		const double p = d->p0, eta_1 = d->p1, eta_2 = d->p2;
		return (x >= 1) ? (p * eta_1 * pow(x, -eta_1 - 1))
		                : ((1 - p) * eta_2 * pow(x, eta_2 - 1));
	}
}

/* exp(2y) g(exp(y)): BlackScholes::computeKappa (BlackScholes.hpp:53-62) */
static double kappa_integrand(double y, const void *ctx) {
	return exp(2 * y) * density_eval(exp(y), (const qpo_density *)ctx);
}

/* g(exp(x)) exp(x): computeDensityFFT's fbar (BlackScholes.hpp:306-308) */
static double fbar_integrand(double x, const void *ctx) {
	return density_eval(exp(x), (const qpo_density *)ctx) * exp(x);
}

int qpo_jump_nfft(int n) {
	int N = 2;
	while(N < n) N *= 2;   /* BlackScholes.hpp:281-284 */
	return N;
}

void qpo_jump_setup(int n, const double *S, int density_kind, double p0, double p1,
		double p2, double *x0_out, double *dx_out, double *weights, double *kappa_out) {
	qpo_density d;
	const int N = qpo_jump_nfft(n);
	const int i0 = S[0] < DBL_EPSILON ? 1 : 0;                /* :288-291 */
	const double x0 = log(S[i0]), xf = log(S[n - 2]);
	const double dx = (xf - x0) / (N - 1);
	int i;
	d.kind = density_kind; d.p0 = p0; d.p1 = p1; d.p2 = p2;
	*x0_out = x0; *dx_out = dx;
	*kappa_out = quad_integral(kappa_integrand, &d, -INFINITY, INFINITY) - 1.;
	for(i = 0; i <= N / 2; ++i) {                             /* :313-318 */
		const double a = dx * (-.5 + i), b = dx * (.5 + i);
		weights[i] = quad_integral(fbar_integrand, &d, a, b);
	}
	for(i = N / 2 + 1; i < N; ++i) {                          /* :319-324 */
		const double a = dx * (-.5 + i - N), b = dx * (.5 + i - N);
		weights[i] = quad_integral(fbar_integrand, &d, a, b);
	}
}

/* Radix-2 complex FFT, same butterflies and twiddles as the binding's
 * Eigen::FFT stand-in (oracle/shim/qpde_linalg_binding.hpp). */
static void fft_inplace(int n, double *re, double *im, int inverse) {
	int i, j = 0, len, k;
	double *wr, *wi;
	const double sgn = inverse ? 1. : -1.;
	for(i = 1; i < n; ++i) {
		int bit = n >> 1;
		for(; j & bit; bit >>= 1) j ^= bit;
		j ^= bit;
		if(i < j) {
			double t = re[i]; re[i] = re[j]; re[j] = t;
			t = im[i]; im[i] = im[j]; im[j] = t;
		}
	}
	wr = (double *)malloc(sizeof(double) * (size_t)(n / 2));
	wi = (double *)malloc(sizeof(double) * (size_t)(n / 2));
	for(k = 0; k < n / 2; ++k) {
		const double ang = sgn * 2. * M_PI * (double)k / (double)n;
		wr[k] = cos(ang); wi[k] = sin(ang);
	}
	for(len = 2; len <= n; len <<= 1) {
		const int half = len >> 1, stride = n / len;
		for(i = 0; i < n; i += len) {
			for(k = 0; k < half; ++k) {
				/* std::complex multiplication: (a+bi)(c+di) = (ac - bd) + (ad + bc)i */
				const double a = re[i + k + half], b = im[i + k + half];
				const double c = wr[k * stride], d = wi[k * stride];
				const double tr = a * c - b * d, ti = a * d + b * c;
				const double ur = re[i + k], ui = im[i + k];
				re[i + k] = ur + tr; im[i + k] = ui + ti;
				re[i + k + half] = ur - tr; im[i + k + half] = ui - ti;
			}
		}
	}
	free(wr); free(wi);
}

/* BlackScholesJumpDiffusion::b (BlackScholes.hpp:389-459): b_i = lambda h(log S_i),
 * h = circular correlation of V(exp(x0 + k dx)) with the density weights. */
void qpo_jump_term(int n, const double *S, const double *v, double lambda, int N,
		double x0, double dx, const double *weights, double *b) {
	double *re = (double *)malloc(sizeof(double) * (size_t)N);
	double *im = (double *)calloc((size_t)N, sizeof(double));
	double *wr = (double *)malloc(sizeof(double) * (size_t)N);
	double *wi = (double *)calloc((size_t)N, sizeof(double));
	double *F = (double *)malloc(sizeof(double) * (size_t)N);
	const double xf = log(S[n - 2]);
	const double dxu = (xf - x0) / (N - 1);     /* Axis::uniform(x0, xf, N), Axis.hpp:219-226 */
	int i;
	(void)dx;
	for(i = 0; i < N; ++i) {
		re[i] = qpo_interp(n, S, v, exp(x0 + i * dx));
		wr[i] = weights[i];
		F[i] = x0 + dxu * i;
	}
	fft_inplace(N, re, im, 0);
	fft_inplace(N, wr, wi, 0);
	for(i = 0; i < N; ++i) {
		/* bufferFFT[i] *= conj(fprimeFFT[i]) */
		const double a = re[i], bb = im[i], c = wr[i], d = -wi[i];
		re[i] = a * c - bb * d; im[i] = a * d + bb * c;
	}
	fft_inplace(N, re, im, 1);
	{
		const double s = 1. / (double)N;
		for(i = 0; i < N; ++i) re[i] = re[i] * s;
	}
	b[0] = 0.;
	for(i = 1; i < n - 1; ++i) b[i] = lambda * qpo_interp(N, F, re, log(S[i]));
	b[n - 1] = 0.;
	free(re); free(im); free(wr); free(wi); free(F);
}

/* ------------------------------------------------------------------------ */
/* Tridiagonal LU with row partial pivoting, natural order.  Same operation  */
/* order as oracle/shim's banded SparseLU stand-in with kl = ku = 1 (the     */
/* pivot policy Eigen::SparseLU defaults to: keep the diagonal unless a      */
/* strictly larger entry lies below it).                                     */
/* ------------------------------------------------------------------------ */

typedef struct {
	int n;
	double *dl;  /* multipliers                */
	double *d;   /* U diagonal                 */
	double *du;  /* U first superdiagonal      */
	double *du2; /* U second superdiagonal     */
	int *swap;   /* 1 if rows j, j+1 swapped   */
} tri_lu;

static void tri_alloc(tri_lu *f, int n) {
	f->n = n;
	f->dl = (double *)malloc(sizeof(double) * (size_t)n);
	f->d = (double *)malloc(sizeof(double) * (size_t)n);
	f->du = (double *)malloc(sizeof(double) * (size_t)n);
	f->du2 = (double *)malloc(sizeof(double) * (size_t)n);
	f->swap = (int *)malloc(sizeof(int) * (size_t)n);
}

static void tri_free(tri_lu *f) {
	free(f->dl); free(f->d); free(f->du); free(f->du2); free(f->swap);
}

/* a: sub (a[0] unused), b: diag, c: super (c[n-1] unused) */
static void tri_factor(tri_lu *f, const double *a, const double *b,
		const double *c) {
	const int n = f->n;
	int j;
	/* working copy of rows j, j+1 */
	for(j = 0; j < n; ++j) {
		f->d[j] = b[j];
		f->du[j] = j < n - 1 ? c[j] : 0.;
		f->du2[j] = 0.;
		f->dl[j] = j < n - 1 ? a[j + 1] : 0.;
	}
	for(j = 0; j < n - 1; ++j) {
		/* pivot search among rows j, j+1 of column j */
		if(fabs(f->dl[j]) > fabs(f->d[j])) {
			/* swap rows j and j+1 (columns j .. j+2) */
			double t;
			t = f->d[j];  f->d[j] = f->dl[j];      f->dl[j] = t;
			t = f->du[j]; f->du[j] = f->d[j + 1];  f->d[j + 1] = t;
			t = f->du2[j]; f->du2[j] = f->du[j + 1]; f->du[j + 1] = t;
			f->swap[j] = 1;
		} else {
			f->swap[j] = 0;
		}
		f->dl[j] /= f->d[j];
		if(f->du[j] != 0.)  f->d[j + 1]  -= f->dl[j] * f->du[j];
		if(f->du2[j] != 0.) f->du[j + 1] -= f->dl[j] * f->du2[j];
	}
	f->swap[n - 1] = 0;
}

static void tri_solve(const tri_lu *f, double *x) {
	const int n = f->n;
	int j;
	for(j = 0; j < n - 1; ++j) {
		if(f->swap[j]) { const double t = x[j]; x[j] = x[j + 1]; x[j + 1] = t; }
		x[j + 1] -= f->dl[j] * x[j];
	}
	/* column-oriented back substitution, same update order as the banded
	 * stand-in: x[j] /= U_jj, then x[i] -= U_ij x[j] for i = j-2, j-1 */
	for(j = n - 1; j >= 0; --j) {
		x[j] /= f->d[j];
		if(j >= 2) x[j - 2] -= f->du2[j - 2] * x[j];
		if(j >= 1) x[j - 1] -= f->du[j - 1] * x[j];
	}
}

/* Arithmetic model of the product's INTERLEAVED kernel (Thomas elimination
 * without pivoting, one reciprocal per row, no FMA).  Selected with
 * qpo_problem.solver = 1; used by the tests to tell apart "the kernel differs
 * from its own arithmetic model" (a bug) from "a different but valid
 * elimination order moved a value by one ulp" (a tie). */
static void thomas_model(int n, const double *a, const double *b,
		const double *c, double *x, double *cp, double *dp) {
	double cprev = 0., dprev = 0., xn = 0.;
	int i;
	for(i = 0; i < n; ++i) {
		const double denom = b[i] - a[i] * cprev;
		const double inv = 1. / denom;
		cprev = c[i] * inv;
		dprev = (x[i] - a[i] * dprev) * inv;
		cp[i] = cprev; dp[i] = dprev;
	}
	for(i = n - 1; i >= 0; --i) {
		xn = dp[i] - cp[i] * xn;
		x[i] = xn;
	}
}

/* ------------------------------------------------------------------------ */
/* relativeError: Core/IterativeMethod.hpp:1125-1137                         */
/* ------------------------------------------------------------------------ */

static double relative_error(int n, const double *a, const double *b,
		double scale) {
	double m = 0.;
	int i;
	for(i = 0; i < n; ++i) {
		const double fa = fabs(a[i]), fb = fabs(b[i]);
		double den = fa < fb ? fb : fa;      /* a.cwiseAbs().cwiseMax(b.cwiseAbs()) */
		double quo;
		den = (scale * 1.) < den ? den : (scale * 1.); /* (scale*Ones).cwiseMax(..) */
		quo = fabs(a[i] - b[i]) / den;
		if(i == 0 || quo > m) m = quo;
	}
	return m;
}

/* ------------------------------------------------------------------------ */
/* The sweep                                                                 */
/* ------------------------------------------------------------------------ */

/* row-major SpMV of the tridiagonal (lo, di, up): accumulate from zero in
 * column order (the binding's operator*, standing in for Eigen's). */
static void tri_spmv(int n, const double *lo, const double *di,
		const double *up, const double *x, double *y) {
	int i;
	for(i = 0; i < n; ++i) {
		double acc = 0.;
		if(i > 0) acc += lo[i] * x[i - 1];
		acc += di[i] * x[i];
		if(i < n - 1) acc += up[i] * x[i + 1];
		y[i] = acc;
	}
}

/* Variable-step BDF formulas of order k = 3 .. 6 (Core/BDF.hpp:139-470): g[m] = time(m) - t
 * (reverse time, m = 0 the newest iterand); the reference names them h_{k-1} .. h_0.  On
 * return c[m] multiplies iterand(m) in b (signs alternate, + for m = 0) and *hh scales A and b.
 * Every expression is written as the reference writes it, so that it rounds the same way. */
static void bdf_coefficients(int k, const double *g, double *c, double *hh) {
	if(k == 3) {
		const double h2 = g[0], h1 = g[1], h0 = g[2];
		*hh = (h0*h1*h2)/(h0*h1 + h0*h2 + h1*h2);
		c[0] = (h0*h1)/(h2*(h0 - h2)*(h1 - h2));
		c[1] = (h0*h2)/(h1*(h0 - h1)*(h1 - h2));
		c[2] = (h1*h2)/(h0*(h0 - h1)*(h0 - h2));
	} else if(k == 4) {
		const double h3 = g[0], h2 = g[1], h1 = g[2], h0 = g[3];
		*hh = (h0*h1*h2*h3)/(h0*h1*h2 + h0*h1*h3 + h0*h2*h3 + h1*h2*h3);
		c[0] = (h0*h1*h2)/(h3*(h0 - h3)*(h1 - h3)*(h2 - h3));
		c[1] = (h0*h1*h3)/(h2*(h0 - h2)*(h1 - h2)*(h2 - h3));
		c[2] = (h0*h2*h3)/(h1*(h0 - h1)*(h1 - h2)*(h1 - h3));
		c[3] = (h1*h2*h3)/(h0*(h0 - h1)*(h0 - h2)*(h0 - h3));
	} else if(k == 5) {
		const double h4 = g[0], h3 = g[1], h2 = g[2], h1 = g[3], h0 = g[4];
		*hh = (h0*h1*h2*h3*h4)/(h0*h1*h2*h3 + h0*h1*h2*h4 + h0*h1*h3*h4 + h0*h2*h3*h4 + h1*h2*h3*h4);
		c[0] = (h0*h1*h2*h3)/(h4*(h0 - h4)*(h1 - h4)*(h2 - h4)*(h3 - h4));
		c[1] = (h0*h1*h2*h4)/(h3*(h0 - h3)*(h1 - h3)*(h2 - h3)*(h3 - h4));
		c[2] = (h0*h1*h3*h4)/(h2*(h0 - h2)*(h1 - h2)*(h2 - h3)*(h2 - h4));
		c[3] = (h0*h2*h3*h4)/(h1*(h0 - h1)*(h1 - h2)*(h1 - h3)*(h1 - h4));
		c[4] = (h1*h2*h3*h4)/(h0*(h0 - h1)*(h0 - h2)*(h0 - h3)*(h0 - h4));
	} else {
		const double h5 = g[0], h4 = g[1], h3 = g[2], h2 = g[3], h1 = g[4], h0 = g[5];
		*hh = (h0*h1*h2*h3*h4*h5)/(h0*h1*h2*h3*h4 + h0*h1*h2*h3*h5 + h0*h1*h2*h4*h5 + h0*h1*h3*h4*h5 + h0*h2*h3*h4*h5 + h1*h2*h3*h4*h5);
		c[0] = (h0*h1*h2*h3*h4)/(h5*(h0 - h5)*(h1 - h5)*(h2 - h5)*(h3 - h5)*(h4 - h5));
		c[1] = (h0*h1*h2*h3*h5)/(h4*(h0 - h4)*(h1 - h4)*(h2 - h4)*(h3 - h4)*(h4 - h5));
		c[2] = (h0*h1*h2*h4*h5)/(h3*(h0 - h3)*(h1 - h3)*(h2 - h3)*(h3 - h4)*(h3 - h5));
		c[3] = (h0*h1*h3*h4*h5)/(h2*(h0 - h2)*(h1 - h2)*(h2 - h3)*(h2 - h4)*(h2 - h5));
		c[4] = (h0*h2*h3*h4*h5)/(h1*(h0 - h1)*(h1 - h2)*(h1 - h3)*(h1 - h4)*(h1 - h5));
		c[5] = (h1*h2*h3*h4*h5)/(h0*(h0 - h1)*(h0 - h2)*(h0 - h3)*(h0 - h4)*(h0 - h5));
	}
}

/* diagnostics: number of inner solves of the last sweep and how many of them
 * saw an active set different from the one of the previous solve */
long qpo_diag_solves = 0, qpo_diag_mask_changes = 0;

/* One evaluation of an event program at price S (Event<1>::_doEvent, Core/Event.hpp:100-112: the user's
 * transform called with the interpolant of the solution before the event). */
static double event_program_eval(const qpo_problem *p, const double *params, int n, const double *vold, double S) {
	double st[16];
	int sp = 0, pc = 0;
	for(;;) {
		const int op = p->event_program[pc++];
		double a, b;
		if(op == 0) break;
		switch(op) {
		case 1: st[sp++] = S; break;
		case 2: st[sp++] = p->event_consts[p->event_program[pc++]]; break;
		case 3: st[sp++] = params[p->event_program[pc++]]; break;
		case 10: st[sp - 1] = -st[sp - 1]; break;
		case 11: st[sp - 1] = qpo_interp(n, p->S, vold, st[sp - 1]); break;
		case 13: sp -= 2; st[sp - 1] = st[sp - 1] != 0. ? st[sp] : st[sp + 1]; break;
		default:
			b = st[--sp]; a = st[sp - 1];
			if(op == 4) a = a + b;
			else if(op == 5) a = a - b;
			else if(op == 6) a = a * b;
			else if(op == 7) a = a / b;
			else if(op == 8) a = b > a ? b : a;
			else if(op == 9) a = b < a ? b : a;
			else a = a > b ? 1. : 0.;
			st[sp - 1] = a;
		}
	}
	return st[sp - 1];
}

/* time elapsed between a history time and the new implicit time (Rannacher.hpp:17-21) */
#define LAPSE(then, now) (p->forward ? (now) - (then) : (then) - (now))

int qpo_sweep_1d(const qpo_problem *p, double *V, int *inner,
		unsigned char *mask) {
	const int n = p->n;
	const double pen = 1. / p->penalty_tol;   /* PenaltyMethod.hpp:85 */
	double *A_lo = (double *)malloc(sizeof(double) * (size_t)n);
	double *A_di = (double *)malloc(sizeof(double) * (size_t)n);
	double *A_up = (double *)malloc(sizeof(double) * (size_t)n);
	double *M_lo = (double *)malloc(sizeof(double) * (size_t)n);
	double *M_di = (double *)malloc(sizeof(double) * (size_t)n);
	double *M_up = (double *)malloc(sizeof(double) * (size_t)n);
	double *P_di = (double *)malloc(sizeof(double) * (size_t)n);
	double *lb = (double *)malloc(sizeof(double) * (size_t)n);
	double *rhs = (double *)malloc(sizeof(double) * (size_t)n);
	double *x = (double *)malloc(sizeof(double) * (size_t)n);
	double *xprev = (double *)malloc(sizeof(double) * (size_t)n);
	double *xin = (double *)calloc((size_t)n, sizeof(double));
	double *v1 = (double *)malloc(sizeof(double) * (size_t)n); /* iterand(0) */
	double *v0 = (double *)malloc(sizeof(double) * (size_t)n); /* iterand(1) */
	double *bj = (double *)calloc((size_t)n, sizeof(double)); /* op.b(): explicit jump term or 0 */
	/* BDF3 .. BDF6: iterand(2) .. iterand(5) and their times */
	const int order = p->scheme >= QPO_SCHEME_BDF2 ? p->scheme - QPO_SCHEME_BDF2 + 2 : 1;
	double *vh[4] = { NULL, NULL, NULL, NULL };
	double th[4] = { 0., 0., 0., 0. };
	tri_lu lu;
	int s, i, k, status = 0;
	const int variable = p->variable_target > 0.;
	qpo_step vstep;
	double vdt = p->variable_dt0, vt_t = p->T, vt_dt = -1., vt_dtPrevious = -1.;
	int vt_done = 0;
	/* stepper history: times of iterand(0), iterand(1) */
	double time1 = 0., time0 = 0.;
	int hist = 0;          /* number of iterands in the stepper's ring */
	int events_popped = 0; /* user events applied so far */
	int initialized = 0;   /* IterativeMethod.hpp:1150-1157 */
	/* node state machines */
	int ran_phase = 1;     /* Rannacher: 1,2 = implicit steps, 3 = first CN, 4 = CN */
	int bdf_phase = 1;     /* BDFTwo: 1 = BDF1 step, 2 = first BDF2, 3 = BDF2; BDFk: up to k + 1 */

	for(i = 0; i + 2 < order; ++i) vh[i] = (double *)calloc((size_t)n, sizeof(double));
	tri_alloc(&lu, n);
	qpo_diag_solves = 0; qpo_diag_mask_changes = 0;
	for(i = 0; i < n; ++i) P_di[i] = 0.;
	memcpy(v1, p->v0, sizeof(double) * (size_t)n);
	/* the stepper pushes (T, V^0) (:1171-1174) */
	time1 = p->forward ? 0. : p->T;   /* initialTime() */
	hist = 1;

	for(s = 0; ; ++s) {
		const qpo_step *st;
		double t;
		int a_same, its = 0, notdone = 0;
		if(variable) {
			/* VariableStepper::timestep (Stepper.hpp:63-83) inside
			 * QUANT_PDE_TMP_TIMESTEP (IterativeMethod.hpp:1296-1308); the only event
			 * is the NullEvent at 0 */
			double tmp, dt;
			if(vt_done) break;
			if(s >= p->n_steps) { status = 3; break; }   /* capacity exceeded */
			if(s > 0) {
				double err = 0.;
				for(i = 0; i < n; ++i) {
					/* relativeError(v1, v0, scale), IterativeMethod.hpp:1125-1137 */
					const double fa = fabs(v1[i]), fb = fabs(v0[i]);
					double den = fa < fb ? fb : fa, e;
					den = p->variable_scale < den ? den : p->variable_scale;
					e = fabs(v1[i] - v0[i]) / den;
					if(e > err) err = e;
				}
				vdt *= p->variable_target / (err + DBL_EPSILON);
			}
			vt_dtPrevious = vt_dt;
			dt = vdt;
			tmp = vt_t + -1. * dt;
			if(fabs(tmp - 0.) < DBL_EPSILON) tmp = 0.;
			else if(tmp < 0.) { tmp = 0.; dt = -1. * (0. - vt_t); }
			vt_t = tmp; vt_dt = dt;
			vstep.t = tmp; vstep.dt = dt; vstep.same_dt = (dt == vt_dtPrevious);
			vt_done = !(0. < tmp + -1. * DBL_EPSILON);
			vstep.event_after = vt_done; vstep.user_events = 0;
			st = &vstep;
		} else {
			if(s >= p->n_steps) break;
			st = &p->steps[s];
		}
		t = st->t;
		int use_cn = 0, use_bdf2 = 0, use_bdfk = 0;
		double hA = 0.;

		/* --- which A / b formulas apply at this step ------------------- */
		switch(p->scheme) {
		case QPO_SCHEME_RANNACHER:
			/* Rannacher.hpp:23-30, 75-92 */
			use_cn = ran_phase >= 3;
			a_same = (ran_phase == 3) ? 0 : st->same_dt;
			break;
		case QPO_SCHEME_CN:
			use_cn = 2; /* CrankNicolson.hpp:52-78 rounding: (L*theta)*dt */
			a_same = st->same_dt; /* :82-90 */
			break;
		case QPO_SCHEME_BDF1:
			a_same = st->same_dt; /* BDF.hpp:511-513 */
			break;
		default: {
			/* BDFk: steps 1 .. k - 1 after clear() use the lower orders (BDF.hpp:601-627,856-900) */
			const int use = bdf_phase < order ? bdf_phase : order;
			use_bdf2 = use == 2;
			use_bdfk = use >= 3 ? use : 0;
			a_same = (bdf_phase <= order) ? 0 : st->same_dt; /* BDF.hpp:20-26,531-550,615-619 */
			break;
		}
		}
		/* PenaltyMethod keeps LinearSystem's default isATheSame() == false
		 * (IterativeMethod.hpp:143-146) */
		if(p->american) a_same = 0;

		/* --- explicit jump term: BlackScholesJumpDiffusion::b at iterand(0) -- */
		if(p->jump_weights) {
			qpo_jump_term(n, p->S, v1, p->jump_lambda, p->jump_nfft, p->jump_x0,
					p->jump_dx, p->jump_weights, bj);
		} else if(p->source_table) {
			/* an operator whose b(t) is a scalar of time times the grid, as GLWBOperator::b
			 * (examples/experimental/glwb.cpp:40-44):  b(t) = g(floor(t)) S,  g(n) = -(R[n+1] - R[n]) */
			int nidx = (int)floor(t);
			double g;
			if(use_cn) { status = 2; break; }   /* restated for the BDF schemes only (b at one time) */
			if(nidx < 0) nidx = 0;
			if(nidx > p->n_source - 1) nidx = p->n_source - 1;
			g = p->source_table[nidx];
			for(i = 0; i < n; ++i) bj[i] = g * p->S[i];
		}

		/* --- left-hand side  lA  and right-hand side  lb ----------------- */
		{
			const double h0 = LAPSE(time1, t); /* difference(t1, t0): t0 - t1 in reverse time, t1 - t0 forward */
			if(use_bdfk) {
				/* BDF.hpp:139-470: A = I + hh op.A, b = (c v_new - c' v' + ... + op.b) hh */
				double g[6], c[6], hh;
				g[0] = LAPSE(time1, t); g[1] = LAPSE(time0, t);
				for(k = 2; k < use_bdfk; ++k) g[k] = LAPSE(th[k - 2], t);
				bdf_coefficients(use_bdfk, g, c, &hh);
				hA = hh;
				for(i = 0; i < n; ++i) {
					double acc = c[0] * v1[i];
					acc = acc - c[1] * v0[i];
					for(k = 2; k < use_bdfk; ++k) {
						if(k & 1) acc = acc - c[k] * vh[k - 2][i];
						else      acc = acc + c[k] * vh[k - 2][i];
					}
					M_lo[i] = hh * p->lo[i];
					M_di[i] = 1. + hh * p->di[i];
					M_up[i] = hh * p->up[i];
					lb[i] = (acc + bj[i]) * hh;
				}
			} else if(use_bdf2) {
				/* BDF.hpp:68-117 */
				const double h1 = LAPSE(time1, t);
				const double hh0 = LAPSE(time0, t);
				const double hh = (hh0 * h1) / (hh0 + h1);
				const double c1 = hh0 / h1, c0 = h1 / hh0, den = hh0 - h1;
				hA = hh;
				for(i = 0; i < n; ++i) {
					M_lo[i] = hh * p->lo[i];
					M_di[i] = 1. + hh * p->di[i];
					M_up[i] = hh * p->up[i];
					lb[i] = ((c1 * v1[i] - c0 * v0[i]) / den + bj[i]) * hh;
				}
			} else if(use_cn == 1) {
				/* Rannacher.hpp:50-73 */
				hA = h0;
				for(i = 0; i < n; ++i) {
					M_lo[i] = p->lo[i] * h0 / 2.;
					M_di[i] = 1. + p->di[i] * h0 / 2.;
					M_up[i] = p->up[i] * h0 / 2.;
				}
				for(i = 0; i < n; ++i) {
					/* (I - A h0/2) v0 + (b(t1)+b(t0))/2, SpMV in column order */
					double acc = 0.;
					if(i > 0) acc += (0. - p->lo[i] * h0 / 2.) * v1[i - 1];
					acc += (1. - p->di[i] * h0 / 2.) * v1[i];
					if(i < n - 1) acc += (0. - p->up[i] * h0 / 2.) * v1[i + 1];
					lb[i] = acc + (bj[i] + bj[i]) / 2.;
				}
			} else if(use_cn == 2) {
				/* CrankNicolson.hpp:52-78 with theta = 1/2 */
				const double theta = 0.5;
				hA = h0;
				for(i = 0; i < n; ++i) {
					M_lo[i] = p->lo[i] * theta * h0;
					M_di[i] = 1. + p->di[i] * theta * h0;
					M_up[i] = p->up[i] * theta * h0;
				}
				for(i = 0; i < n; ++i) {
					double acc = 0.;
					if(i > 0) acc += (0. - p->lo[i] * (1 - theta) * h0) * v1[i - 1];
					acc += (1. - p->di[i] * (1 - theta) * h0) * v1[i];
					if(i < n - 1) acc += (0. - p->up[i] * (1 - theta) * h0) * v1[i + 1];
					lb[i] = acc + (theta * bj[i] + (1 - theta) * bj[i]);
				}
			} else {
				/* implicit Euler: Rannacher.hpp:32-48 / BDF.hpp:32-46 */
				hA = h0;
				for(i = 0; i < n; ++i) {
					M_lo[i] = p->lo[i] * h0;
					M_di[i] = 1. + p->di[i] * h0;
					M_up[i] = p->up[i] * h0;
					lb[i] = p->scheme == QPO_SCHEME_RANNACHER ? v1[i] + bj[i] * h0 : v1[i] + h0 * bj[i];
				}
			}
		}

		if(p->n_controls > 0) {
			/* ToleranceIteration over PolicyIteration<1,1,Max>
			 * (PolicyIteration.hpp:54-105) under BDFOne / BDFTwo: the root's
			 * A is I + hA * bs.A(control), b is the BDF right-hand side.  The
			 * discretisation node's isATheSame() is false because its operator
			 * (the policy node) keeps LinearSystem's default. */
			/* Under the theta schemes the explicit half is the policy node's A as well: Rannacher::_b2 /
			 * CrankNicolson::b call system.A(t0) (Rannacher.hpp:60-73, CrankNicolson.hpp:64-78), and
			 * PolicyIteration::A hands back the controlled operator under the controls of the CURRENT inner
			 * iteration — so the right-hand side is re-formed with every control update. */
			memcpy(x, v1, sizeof(double) * (size_t)n);
			do {
				memcpy(xprev, x, sizeof(double) * (size_t)n);
				for(i = 0; i < n; ++i) {
					int k, best_k = 0;
					double best = p->policy_max ? -INFINITY : INFINITY;
					for(k = 0; k < p->n_controls; ++k) {
						const double *lo_k = p->plo + (size_t)k * n;
						const double *di_k = p->pdi + (size_t)k * n;
						const double *up_k = p->pup + (size_t)k * n;
						/* candidate = A(q) x - b(q): SpMV row from zero in column
						 * order over the stored entries, then minus b = 0 */
						double acc = 0., cand;
						if(i > 0 && i < n - 1) acc += lo_k[i] * xprev[i - 1];
						acc += di_k[i] * xprev[i];
						if(i > 0 && i < n - 1) acc += up_k[i] * xprev[i + 1];
						cand = acc - 0.;
						if(p->policy_max ? cand > best : cand < best) { best = cand; best_k = k; }
					}
					{
						const double l = p->plo[(size_t)best_k * n + i];
						const double d = p->pdi[(size_t)best_k * n + i];
						const double u = p->pup[(size_t)best_k * n + i];
						if(use_bdf2 || use_bdfk) { A_lo[i] = hA * l; A_di[i] = 1. + hA * d; A_up[i] = hA * u; }
						else if(use_cn == 1) { A_lo[i] = l * hA / 2.; A_di[i] = 1. + d * hA / 2.; A_up[i] = u * hA / 2.; }
						else if(use_cn == 2) { A_lo[i] = l * 0.5 * hA; A_di[i] = 1. + d * 0.5 * hA; A_up[i] = u * 0.5 * hA; }
						else         { A_lo[i] = l * hA; A_di[i] = 1. + d * hA; A_up[i] = u * hA; }
						if(use_cn) {
							/* row i of (I - (1 - theta) h A(q*)) V^n, SpMV in column order (the rows of A(q*) carry the
							 * control of their own node) */
							double acc = 0.;
							if(use_cn == 1) {
								if(i > 0) acc += (0. - l * hA / 2.) * v1[i - 1];
								acc += (1. - d * hA / 2.) * v1[i];
								if(i < n - 1) acc += (0. - u * hA / 2.) * v1[i + 1];
								lb[i] = acc + (bj[i] + bj[i]) / 2.;
							} else {
								if(i > 0) acc += (0. - l * (1 - 0.5) * hA) * v1[i - 1];
								acc += (1. - d * (1 - 0.5) * hA) * v1[i];
								if(i < n - 1) acc += (0. - u * (1 - 0.5) * hA) * v1[i + 1];
								lb[i] = acc + (0.5 * bj[i] + (1 - 0.5) * bj[i]);
							}
						}
					}
				}
				memcpy(x, lb, sizeof(double) * (size_t)n);
				if(p->solver) {
					thomas_model(n, A_lo, A_di, A_up, x, P_di, rhs);
				} else {
					tri_factor(&lu, A_lo, A_di, A_up);
					tri_solve(&lu, x);
				}
				++its;
				notdone = relative_error(n, x, xprev, p->scale) > p->tolerance;
				if(notdone && its >= p->max_inner) { status = 1; break; }
			} while(notdone);
		} else if(!p->american) {
			/* IterativeMethod.hpp:899-917 */
			if(!initialized || !a_same) {
				memcpy(A_lo, M_lo, sizeof(double) * (size_t)n);
				memcpy(A_di, M_di, sizeof(double) * (size_t)n);
				memcpy(A_up, M_up, sizeof(double) * (size_t)n);
				if(!p->solver) tri_factor(&lu, A_lo, A_di, A_up);
			}
			memcpy(x, lb, sizeof(double) * (size_t)n);
			if(p->solver) thomas_model(n, A_lo, A_di, A_up, x, xprev, rhs);
			else tri_solve(&lu, x);
			its = 1;
		} else {
			/* ToleranceIteration (:1211-1254) over PenaltyMethodDifference
			 * (PenaltyMethod.hpp:37-69,96-119,259-273) */
			memcpy(x, v1, sizeof(double) * (size_t)n);
			do {
				memcpy(xprev, x, sizeof(double) * (size_t)n);
				{ int changed = 0;
				for(i = 0; i < n; ++i) {
					const int c = (pen * ((0. + 1. * xprev[i]) - p->obstacle[i])) < 0.;
					if((P_di[i] != 0.) != c) changed = 1;
				}
				++qpo_diag_solves; qpo_diag_mask_changes += changed; }
				for(i = 0; i < n; ++i) {
					/* predicate = I x - payoff ; c = (scale*pred) < 0 */
					const double acc = 0. + 1. * xprev[i];
					const double pred = acc - p->obstacle[i];
					const int c = (pen * pred) < 0.;
					P_di[i] = c ? pen : 0.;
					/* A = Q lA + P rA ; b = Q lb + P rb   (Q = I, rA = I) */
					A_lo[i] = M_lo[i];
					A_di[i] = c ? (M_di[i] + pen * 1.) : M_di[i];
					A_up[i] = M_up[i];
					rhs[i] = c ? (lb[i] + pen * p->obstacle[i]) : lb[i];
				}
				memcpy(x, rhs, sizeof(double) * (size_t)n);
				if(p->solver) {
					thomas_model(n, A_lo, A_di, A_up, x, P_di, rhs);
				} else {
					tri_factor(&lu, A_lo, A_di, A_up);
					tri_solve(&lu, x);
				}
				++its;
				notdone = relative_error(n, x, xprev, p->scale) > p->tolerance;
				/* guard only; the reference would keep iterating */
				if(notdone && its >= p->max_inner) { status = 1; break; }
			} while(notdone);
		}
		initialized = 1;
		if(inner) inner[s] = its;
		/* the tolerance iteration's iterand(0): what constraintMask() reads, untouched by the events
		 * below (they transform the stepper's iterand) */
		memcpy(xin, x, sizeof(double) * (size_t)n);

		/* stepper history push (lookback = the BDF order, else >= 1) */
		for(k = order - 3; k >= 0; --k) {
			memcpy(vh[k], k > 0 ? vh[k - 1] : v0, sizeof(double) * (size_t)n);
			th[k] = k > 0 ? th[k - 1] : time0;
		}
		memcpy(v0, v1, sizeof(double) * (size_t)n);
		time0 = time1;
		memcpy(v1, x, sizeof(double) * (size_t)n);
		time1 = t;
		if(hist < 2) ++hist;

		/* node onIterationEnd (called by the stepper's endNodes) */
		if(ran_phase < 4) ++ran_phase;  /* Rannacher.hpp:75-92 */
		if(bdf_phase < order + 1) ++bdf_phase;  /* BDF.hpp:531-545,601-621 */

		if(st->event_after) {
			/* IterativeMethod.hpp:1279-1295 */
			/* Events at one time are popped in descending id (TimeOrder,
			 * IterativeMethod.hpp:1340-1352): the event add()ed last goes first. */
			for(k = 0; k < st->user_events; ++k) {
				int op;
				if(p->event_program) {
					/* events pop in descending time (forward: ascending): the e-th row of the parameter table */
					const int e = p->forward ? events_popped : p->n_events - 1 - events_popped;
					++events_popped;
					memcpy(x, v1, sizeof(double) * (size_t)n);
					for(i = 0; i < n; ++i) {
						v1[i] = event_program_eval(p, p->event_params + (size_t)e * p->n_event_params, n, x, p->S[i]);
					}
					continue;
				}
				for(op = 0; op < 2; ++op) {
					const int is_dividend = (op == 0) == (p->dividend_first != 0);
					if(is_dividend && p->dividend_kind != QPO_DIVIDEND_NONE) {
						/* Event.hpp:100-112 with V(max(S - D, 0)) (README.md:361-363)
						 * or V((1 - D) S); PointwiseMap evaluates it at every node */
						for(i = 0; i < n; ++i) {
							const double Si = p->S[i];
							const double at = p->dividend_kind == QPO_DIVIDEND_CASH
								? (Si - p->dividend > 0. ? Si - p->dividend : 0.)
								: (1. - p->dividend) * Si;
							x[i] = qpo_interp(n, p->S, v1, at);
						}
						memcpy(v1, x, sizeof(double) * (size_t)n);
					}
					if(!is_dividend && p->event_obstacle) {
						/* the exercise transform max(V(S), K - S)
						 * (QuantPDE::max, EarlyInclude.hpp:48) */
						for(i = 0; i < n; ++i) {
							const double ex = p->event_obstacle[i];
							v1[i] = v1[i] > ex ? v1[i] : ex;
						}
					}
				}
			}
			/* afterEvent(): default IterationNode::onAfterEvent -> clear()
			 * (IterativeMethod.hpp:780-783); Rannacher overrides it with a
			 * no-op (Rannacher.hpp:115-117) */
			bdf_phase = 1;
			hist = 1;
		}
	}

	if(p->steps_taken) *p->steps_taken = s;
	if(mask) {
		/* PenaltyMethod::constraintMask (PenaltyMethod.hpp:127-139) on the
		 * tolerance iteration's last iterand */
		for(i = 0; i < n; ++i) {
			const double acc = 0. + 1. * xin[i];
			mask[i] = (acc - (p->obstacle ? p->obstacle[i] : 0.)) < 0.;
		}
	}
	memcpy(V, v1, sizeof(double) * (size_t)n);

	tri_free(&lu);
	free(A_lo); free(A_di); free(A_up); free(M_lo); free(M_di); free(M_up);
	free(bj); free(P_di); free(lb); free(rhs); free(x); free(xprev); free(v1); free(v0); free(xin);
	for(i = 0; i < 4; ++i) free(vh[i]);
	return status;
}
