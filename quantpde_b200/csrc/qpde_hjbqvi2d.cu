// qpde_hjbqvi2d.cu — 2-D impulse-control HJBQVI, penalised scheme, with impulses that point both
// ways in the state (no block-triangular order as in the GMWB problem, qpde_gmwb.cu).
//
// Reference call stack replaced (examples/hjbqvi/optimal_consumption.cpp:29-203):
//   HJBQVI<2,1,1>::solve(refinement)          Modules/HJBQVI/HJBQVI.hpp:590-811
//     ReverseConstantStepper / TimeIteration  Core/IterativeMethod.hpp:1259-1317
//     ToleranceIteration + relativeError      :1125-1137,1211-1254
//     MinPolicyIteration (stochastic control) Core/PolicyIteration.hpp:54-105 over
//       ControlledOperator::A / b             HJBQVI.hpp:158-335 (no boundary routines)
//     MinPolicyIteration (impulse control)    over Impulse::A / b, Core/Impulse.hpp:87-216
//     MinPenaltyMethod (non-direct)           Core/PenaltyMethod.hpp:37-119,127-139
//     ReverseBDFOne                           Core/BDF.hpp:32-46
//     BiCGSTABSolver (+ ILUT)                 Bindings/Eigen.hpp:170-204
//
// Per policy iteration:
//   k_policy   one thread per node: both control searches against the previous iterate, the penalty flag,
//              and the node's row of (I + h L^q + P (I - M*)) split into its w-line part (tridiagonal:
//              the diffusion acts along w only) and the rest (the two b-neighbours of the upwinded drift,
//              the impulse entries that are off the line's three diagonals)
//   k_factor   one thread per w-line: Thomas factors of the line's tridiagonal matrix
//   the linear solve: block Jacobi over the w-lines, repeated to a 1e-15 relative update —
//     k_rhs    one thread per node: right-hand side with the off-line entries from the last sweep
//     k_lines  one warp per w-line: forward / back substitution with the cached factors, both first-order
//              linear recurrences, as warp scans over tiles of 32 rows (composition of affine maps)
//     (folding k_rhs into k_lines — one launch per sweep — was slower: 1.7 -> 2.3 s at 153 x 153 nodes; the
//     gathers of the right-hand side want one thread per node, not one warp per line walking its tiles)
//   Every sweep is parallel over all lines (a sequential symmetric line Gauss-Seidel on the CPU needs as
//   many line solves: measured 105 / 128 / 154 Jacobi sweeps per solve at refinements 0 / 1 / 2 against
//   47 / 54 / 65 symmetric Gauss-Seidel sweeps of two passes each).  The reference drives BiCGSTAB +
//   ILUT to a machine-epsilon residual, so both reach the solution of the same system.
//   k_relerr   relativeError(x, xprev) > tolerance
// Storage is the reference's (row = iw + n_w ib, w fastest): a warp of k_lines reads 32 consecutive rows of
// its line at a time.  Sums over a row of the reference's matrices are formed in its column order.
#include <math_constants.h>

#include <cmath>
#include <string>
#include <vector>

#include "qpde_kernels.cuh"

namespace qpde {

namespace {

struct Model2 {     // optimal_consumption.cpp:36-80
	double rho, r, mu, sigma, gamma, lambda, c, boundary;
};

struct Grid2 {
	int nw, nb, nq, nz;
	const double *W, *B, *q, *zfrac, *flow_q;   // device
};

// per-node arrays
struct Rows2 {
	double *la, *lb, *lc;        // the line's tridiagonal row
	double *base;                // v0-independent part of the right-hand side: h0 b^q + P reward*
	double *abh, *cbh;           // couplings to the b-neighbours (h0 * entries of L^q)
	double *ff;                  // [4][n] P * bilinear factor of the impulse targets that are off the line's diagonals (0 if folded / inactive)
	int *ft;                     // [4][n] their storage indices
	double *inv, *cp;            // Thomas factors of the line
	double *qsel, *zsel;         // chosen controls
	int *it;                     // [4][n] impulse targets (reference row indices) of the chosen impulse
	double *iw4;                 // [4][n] their bilinear factors
	double *rew;                 // reward of the chosen impulse
	unsigned char *active;
};

// linearInterpolationData (Core/Interpolant.hpp:153-181): the bracket x[idx] <= ci < x[idx + 1] is unique, so it is
// found from a guess on the mean spacing and a short walk instead of a bisection (the axes of this model are
// refined uniform axes: the guess is right or one off; any strictly increasing axis still ends at the same bracket
// and the same weight).  ncu: the bisections were 45 % of k_policy's stall samples.
__device__ __forceinline__ void interp_dev(const double *x, int length, double inv_mean, double ci, int &idx, double &w) {
	if(ci <= x[0]) { idx = 0; w = 1.; return; }
	if(ci >= x[length - 1]) { idx = length - 2; w = 0.; return; }
	int i = (int)((ci - x[0]) * inv_mean);
	i = i < 0 ? 0 : (i > length - 2 ? length - 2 : i);
	while(x[i] > ci) --i;                 // terminates: x[0] < ci
	while(x[i + 1] <= ci) ++i;            // terminates: ci < x[length - 1]
	idx = i;
	w = (x[i + 1] - ci) / (x[i + 1] - x[i]);
}

// one dimension of the stencil (HJBQVI.hpp:224-274)
__device__ __forceinline__ void stencil(const double *x, int i, int n, double v, double mu, double &lo, double &up, double &total) {
	lo = 0.; up = 0.;
	if(i == 0 || i == n - 1) return;
	const double dxb = x[i] - x[i - 1], dxc = x[i + 1] - x[i - 1], dxf = x[i + 1] - x[i];
	const double vv = v * v;
	const double alpha_common = vv / dxb / dxc, beta_common = vv / dxf / dxc;
	double alpha = alpha_common - mu / dxc, beta = beta_common + mu / dxc;
	if(alpha < 0.) { alpha = alpha_common; beta = beta_common + mu / dxf; }
	else if(beta < 0.) { alpha = alpha_common - mu / dxb; beta = beta_common; }
	lo = -alpha; up = -beta;
	total += alpha + beta;
}

// ((I - M) x)_row in the reference's column order; t[] ascending reference rows, x indexed by storage
struct Imp4 { int t[4]; double f[4]; };
__device__ __forceinline__ double impulse_lhs(const Imp4 &m, int row, const double *x, int nw, int nb) {
	auto at = [&](int r) { return r; };
	double acc = 0.;
	bool diag_done = false;
#pragma unroll
	for(int k = 0; k < 4; ++k) {
		if(!diag_done && row < m.t[k]) { acc += 1. * x[at(row)]; diag_done = true; }
		if(m.t[k] == row) { acc += (1. - m.f[k]) * x[at(row)]; diag_done = true; }
		else acc += (0. - m.f[k]) * x[at(m.t[k])];
	}
	if(!diag_done) acc += 1. * x[at(row)];
	return acc;
}

__device__ __forceinline__ Imp4 impulse_of(const Model2 &md, const Grid2 &g, const double *sW, const double *sB, double invW, double invB,
		double w, double b, double lo, double hi, double z_frac, double &reward) {
	double z = z_frac * (hi - lo) + lo;                         // z_bounded, optimal_consumption.cpp:85-99
	if(fabs(z) < DBL_EPSILON) z = 0.;
	const double w_plus = w + z, b_plus = b - z - md.lambda * fabs(z) - md.c;
	int i0, i1; double f0, f1;
	interp_dev(sW, g.nw, invW, w_plus, i0, f0);
	interp_dev(sB, g.nb, invB, b_plus, i1, f1);
	Imp4 m;
	m.t[0] = i0 + g.nw * i1;           m.f[0] = 1. * f0 * f1;
	m.t[1] = i0 + 1 + g.nw * i1;       m.f[1] = 1. * (-f0 + 1.) * f1;
	m.t[2] = i0 + g.nw * (i1 + 1);     m.f[2] = 1. * f0 * (-f1 + 1.);
	m.t[3] = i0 + 1 + g.nw * (i1 + 1); m.f[3] = 1. * (-f0 + 1.) * (-f1 + 1.);
	reward = z == 0. ? -CUDART_INF : 0.;
	return m;
}

__global__ void __launch_bounds__(128)
k_policy(Model2 md, Grid2 g, Rows2 R, const double *xprev, const double *v0, double h0, double pen) {
	extern __shared__ double sAxes[];
	double *sW = sAxes, *sB = sAxes + g.nw;
	for(int i = threadIdx.x; i < g.nw; i += blockDim.x) sW[i] = g.W[i];
	for(int i = threadIdx.x; i < g.nb; i += blockDim.x) sB[i] = g.B[i];
	__syncthreads();
	const int n = g.nw * g.nb;
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if(s >= n) return;
	const int nw = g.nw, nb = g.nb;
	const int ib = s / nw, iw = s % nw;
	const int row = s;
	const double w = sW[iw], b = sB[ib];
	const double invW = (double)(nw - 1) / (sW[nw - 1] - sW[0]), invB = (double)(nb - 1) / (sB[nb - 1] - sB[0]);
	const bool inside = 0 < b && b < md.boundary;
	const bool wint = iw > 0 && iw < nw - 1, bint = ib > 0 && ib < nb - 1;
	const double xc = xprev[s];
	const double xwm = wint ? xprev[s - 1] : 0., xwp = wint ? xprev[s + 1] : 0.;
	const double xbm = bint ? xprev[s - nw] : 0., xbp = bint ? xprev[s + nw] : 0.;
	// 1. stochastic policy (PolicyIteration.hpp:88-96): the w-direction does not depend on the control
	double aw, cw, total_w = 0.;
	stencil(sW, iw, nw, md.sigma * w, 0. + md.mu * w, aw, cw, total_w);
	double best = CUDART_INF, bab = 0., bcb = 0., bdi = 0., bfl = 0., bq = 0.;
	for(int k = 0; k < g.nq; ++k) {
		const double q = g.q[k];
		double ab, cb, total = total_w;
		stencil(sB, ib, nb, 0., 0. + (md.r * b + (inside ? -q : 0.)), ab, cb, total);
		const double di = total + md.rho, fl = inside ? g.flow_q[k] : 0.;
		double acc = 0.;
		if(bint) acc += ab * xbm;
		if(wint) acc += aw * xwm;
		acc += di * xc;
		if(wint) acc += cw * xwp;
		if(bint) acc += cb * xbp;
		const double cand = acc - fl;
		if(cand < best) { best = cand; bab = ab; bcb = cb; bdi = di; bfl = fl; bq = q; }
	}
	// 2. impulse policy
	const double tmp1 = b - md.c, tmp2 = tmp1 - md.boundary;
	const double t2 = (tmp2 > 0) ? (tmp2 / (1 - md.lambda)) : (tmp2 / (1 + md.lambda));
	const double t1 = (tmp1 > 0) ? (tmp1 / (1 + md.lambda)) : (tmp1 / (1 - md.lambda));
	const double lo = -w > t2 ? -w : t2;                              // QuantPDE::max / min, EarlyInclude.hpp:48-49
	const double hi = md.boundary - w < t1 ? md.boundary - w : t1;
	best = CUDART_INF;
	double zbest = 0.;
	for(int k = 0; k < g.nz; ++k) {
		double rew;
		const Imp4 m = impulse_of(md, g, sW, sB, invW, invB, w, b, lo, hi, g.zfrac[k], rew);
		if(rew != 0.) continue;                                      // -inf: the impulse does not move the state
		const double cand = impulse_lhs(m, row, xprev, nw, nb) - rew;
		if(cand < best) { best = cand; zbest = g.zfrac[k]; }
	}
	// a node without an admissible candidate keeps the control of a zero-filled vector
	double rew;
	const Imp4 m = impulse_of(md, g, sW, sB, invW, invB, w, b, lo, hi, zbest, rew);
	// 3. penalty (PenaltyMethod.hpp:45-68)
	const bool act = (pen * (impulse_lhs(m, row, xprev, nw, nb) - rew)) < 0.;
	// 4. the row, split
	double a_ = aw * h0, b_ = 1. + bdi * h0, c_ = cw * h0;
	double base = h0 * bfl;
	const int n4 = n;
#pragma unroll
	for(int k = 0; k < 4; ++k) {
		R.it[k * n4 + s] = m.t[k]; R.iw4[k * n4 + s] = m.f[k];
		double f = 0.; int tgt = s;
		if(act) {
			const int tk = m.t[k];
			const double pf = pen * m.f[k];
			if(tk == row) b_ -= pf;
			else if(tk == row - 1 && iw > 0) a_ -= pf;
			else if(tk == row + 1 && iw < nw - 1) c_ -= pf;
			else { f = pf; tgt = tk; }
		}
		R.ff[k * n4 + s] = f; R.ft[k * n4 + s] = tgt;
	}
	if(act) { b_ += pen; base += pen * rew; }
	R.la[s] = a_; R.lb[s] = b_; R.lc[s] = c_;
	R.base[s] = base;
	R.abh[s] = ib > 0 ? bab * h0 : 0.; R.cbh[s] = ib < nb - 1 ? bcb * h0 : 0.;
	R.qsel[s] = bq; R.zsel[s] = zbest; R.rew[s] = rew; R.active[s] = act ? 1 : 0;
}

// Thomas factors of every w-line (one thread per line; once per policy iteration)
__global__ void k_factor(int nw, int nb, Rows2 R) {
	const int ib = blockIdx.x * blockDim.x + threadIdx.x;
	if(ib >= nb) return;
	double cprev = 0.;
	for(int iw = 0; iw < nw; ++iw) {
		const int s = iw + nw * ib;
		const double inv = 1. / (R.lb[s] - R.la[s] * cprev);
		cprev = R.lc[s] * inv;
		R.inv[s] = inv; R.cp[s] = cprev;
	}
}

// right-hand side of every node with the off-line entries taken from the last sweep
__global__ void k_rhs(int n, int nw, int nb, Rows2 R, const double *v0, const double *xo, double *rhs) {
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if(s >= n) return;
	const int ib = s / nw;
	double r = v0[s] + R.base[s];
	if(ib > 0) r -= R.abh[s] * xo[s - nw];
	if(ib < nb - 1) r -= R.cbh[s] * xo[s + nw];
	if(R.active[s]) {
#pragma unroll
		for(int k = 0; k < 4; ++k) {
			const double f = R.ff[k * n + s];
			if(f != 0.) r += f * xo[R.ft[k * n + s]];
		}
	}
	rhs[s] = r;
}

// Forward / back substitution of every line with the cached Thomas factors, one warp per line.  Both are
// first-order linear recurrences  y_i = P_i y_(i-1) + Q_i  (forward: P = -la inv, Q = rhs inv; backward, in
// descending rows: P = -cp, Q = d): a tile of 32 consecutive rows is scanned across the lanes by composing
// the affine maps (5 shuffle steps), the value at the end of a tile is carried into the next one.
// *unsettled is raised if some node moved by more than 1e-15 relative.
__global__ void __launch_bounds__(128)
k_lines(int nw, int nb, Rows2 R, double *rhs, const double *xo, double *xn, int *unsettled) {
	const int lane = threadIdx.x & 31;
	const int ib = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if(ib >= nb) return;                                  // whole warps leave together
	const size_t l0 = (size_t)nw * ib;
	const int tiles = (nw + 31) / 32;
	double carry = 0.;
	for(int t = 0; t < tiles; ++t) {
		const int iw = 32 * t + lane;
		double P = 1., Q = 0.;
		if(iw < nw) { const double inv = R.inv[l0 + iw]; P = -R.la[l0 + iw] * inv; Q = rhs[l0 + iw] * inv; }
#pragma unroll
		for(int off = 1; off < 32; off <<= 1) {
			const double Pp = __shfl_up_sync(0xffffffffu, P, off), Qp = __shfl_up_sync(0xffffffffu, Q, off);
			if(lane >= off) { Q = fma(P, Qp, Q); P = P * Pp; }
		}
		const double d = fma(P, carry, Q);
		if(iw < nw) rhs[l0 + iw] = d;
		carry = __shfl_sync(0xffffffffu, d, 31);
	}
	bool big = false;
	carry = 0.;
	for(int t = tiles - 1; t >= 0; --t) {
		const int iw = 32 * t + lane;
		double P = 1., Q = 0.;
		if(iw < nw) { P = -R.cp[l0 + iw]; Q = rhs[l0 + iw]; }
#pragma unroll
		for(int off = 1; off < 32; off <<= 1) {
			const double Pp = __shfl_down_sync(0xffffffffu, P, off), Qp = __shfl_down_sync(0xffffffffu, Q, off);
			if(lane + off < 32) { Q = fma(P, Qp, Q); P = P * Pp; }
		}
		const double xv = fma(P, carry, Q);
		if(iw < nw) {
			const double dd = fabs(xv - xo[l0 + iw]);
			const double den = fabs(xv) > 1. ? fabs(xv) : 1.;
			big = big || dd > 1e-15 * den;
			xn[l0 + iw] = xv;
		}
		carry = __shfl_sync(0xffffffffu, xv, 0);
	}
	if(__any_sync(0xffffffffu, big) && lane == 0) atomicOr(unsettled, 1);
}

// relativeError(x, xprev) > tolerance (IterativeMethod.hpp:1125-1137, scale 1)
__global__ void k_relerr(int n, const double *x, const double *xprev, double tol, int *notdone) {
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	bool nd = false;
	if(s < n) {
		const double fa = fabs(x[s]), fb = fabs(xprev[s]);
		double den = fa < fb ? fb : fa;
		den = 1. < den ? den : 1.;
		nd = fabs(x[s] - xprev[s]) / den > tol;
	}
	if(__any_sync(0xffffffffu, nd) && (threadIdx.x & 31) == 0) atomicOr(notdone, 1);
}

// outputs in the reference's order (w fastest): value, controls (NaN where the other control acts,
// HJBQVI.hpp:977-990), PenaltyMethod::constraintMask on the final iterate (PenaltyMethod.hpp:127-139)
__global__ void k_outputs(int nw, int nb, Rows2 R, const double *x, double *V, double *Q, double *Z, unsigned char *mask) {
	const int n = nw * nb;
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if(s >= n) return;
	const int row = s;
	Imp4 m;
#pragma unroll
	for(int k = 0; k < 4; ++k) { m.t[k] = R.it[k * n + s]; m.f[k] = R.iw4[k * n + s]; }
	const bool act = impulse_lhs(m, row, x, nw, nb) - R.rew[s] < 0.;
	const double qnan = __longlong_as_double(0x7ff8000000000000LL);
	V[row] = x[s];
	if(Q) Q[row] = act ? qnan : R.qsel[s];
	if(Z) Z[row] = act ? R.zsel[s] : qnan;
	if(mask) mask[row] = act ? 1 : 0;
}

} // namespace

int hjbqvi2d_run(const qpde_hjbqvi2d *p, const double *W, int nw, const double *B, int nb, const double *q, int nq,
		const double *z, int nz, double *V_host, double *Q_host, double *Z_host, unsigned char *mask_host,
		qpde_hjbqvi2d_stats *stats, cudaStream_t stream, long long *launches, std::string *error) {
	const int n = nw * nb;
	Model2 md;
	md.rho = p->params[0]; md.r = p->params[1]; md.mu = p->params[2]; md.sigma = p->params[3]; md.gamma = p->params[4];
	md.lambda = p->params[5]; md.c = p->params[6]; md.boundary = p->params[7];
	// the transcendental parts on the host, through the same libm as the reference: the consumption
	// utility per control (optimal_consumption.cpp:138-141) and the exit values (:165-168)
	std::vector<double> flow((size_t)nq), exitv((size_t)n);
	for(int k = 0; k < nq; ++k) flow[k] = std::pow(q[k], md.gamma) / md.gamma;
	for(int i = 0; i < n; ++i) {
		const double w = W[i % nw], b = B[i / nw];
		const double t = b + (1 - md.lambda) * w - md.c;
		const double liquidate = t > 0. ? t : 0.;
		exitv[i] = std::pow(liquidate, md.gamma) / md.gamma;
	}
	// one device block: doubles first, then ints, then bytes
	const size_t nd = (size_t)nw + nb + 2 * (size_t)nq + nz + (size_t)n * (4 /* x, xn, xprev, v0 */ + 1 /* rhs */ + 1 /* exit / scratch */
		+ 3 + 1 + 2 + 4 + 2 + 2 + 4 + 1 + 3 /* V Q Z out */);
	const size_t ni = (size_t)n * 8 + 4;
	const size_t bytes = nd * sizeof(double) + ni * sizeof(int) + 2 * (size_t)n + 64;
	unsigned char *block = nullptr;
	if(cudaMalloc((void **)&block, bytes) != cudaSuccess) { cudaGetLastError(); *error = "hjbqvi2d: out of device memory"; return QPDE_ERR_CUDA; }
	double *d = reinterpret_cast<double *>(block);
	auto take = [&](size_t count) { double *ptr = d; d += count; return ptr; };
	double *dW = take(nw), *dB = take(nb), *dq = take(nq), *dflow = take(nq), *dz = take(nz);
	double *x = take(n), *xn = take(n), *xprev = take(n), *v0 = take(n), *rhs = take(n); take(n);   // (one spare array)
	Rows2 R;
	R.la = take(n); R.lb = take(n); R.lc = take(n); R.base = take(n); R.abh = take(n); R.cbh = take(n);
	R.ff = take(4 * (size_t)n); R.inv = take(n); R.cp = take(n); R.qsel = take(n); R.zsel = take(n);
	R.iw4 = take(4 * (size_t)n); R.rew = take(n);
	double *dV = take(n), *dQ = take(n), *dZ = take(n);
	int *di = reinterpret_cast<int *>(d);
	R.ft = di; R.it = di + 4 * (size_t)n;
	int *dflag = di + 8 * (size_t)n;                   // [0] notdone (policy iteration); [1] unsettled (line sweeps)
	int *dunsettled = dflag + 1;
	R.active = reinterpret_cast<unsigned char *>(dflag + 4);
	unsigned char *dmask = R.active + n;

	Grid2 g; g.nw = nw; g.nb = nb; g.nq = nq; g.nz = nz; g.W = dW; g.B = dB; g.q = dq; g.zfrac = dz; g.flow_q = dflow;
	int rc = QPDE_OK;
	int *hflag = nullptr;
	cudaHostAlloc((void **)&hflag, 16, cudaHostAllocDefault);
	cudaGraph_t graph = nullptr; cudaGraphExec_t group = nullptr;
	const int TB = 128, nblk = (n + TB - 1) / TB, lblk = (nb + 63) / 64, wblk = (nb + 3) / 4;   // k_lines: 4 warps = 4 lines per block
	const size_t axes_smem = sizeof(double) * ((size_t)nw + nb);
	long long solves = 0, sweeps_total = 0;
	int status = 0, n_steps = 0; long long inner_sum = 0;
	do {
		if(!hflag) { rc = QPDE_ERR_CUDA; break; }
		if(axes_smem > 48 * 1024 && cudaFuncSetAttribute(k_policy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)axes_smem) != cudaSuccess) { rc = QPDE_ERR_CUDA; break; }
		cudaMemcpyAsync(dW, W, sizeof(double) * nw, cudaMemcpyHostToDevice, stream);
		cudaMemcpyAsync(dB, B, sizeof(double) * nb, cudaMemcpyHostToDevice, stream);
		cudaMemcpyAsync(dq, q, sizeof(double) * nq, cudaMemcpyHostToDevice, stream);
		cudaMemcpyAsync(dflow, flow.data(), sizeof(double) * nq, cudaMemcpyHostToDevice, stream);
		cudaMemcpyAsync(dz, z, sizeof(double) * nz, cudaMemcpyHostToDevice, stream);
		cudaMemcpyAsync(x, exitv.data(), sizeof(double) * n, cudaMemcpyHostToDevice, stream);
		const double dt = p->T / (double)p->timesteps;
		const double pen = 1. / (p->scaling_factor * dt);          // HJBQVI.hpp:634-635, PenaltyMethod.hpp:85
		Replay rp; rp.start(p->T, dt, 0, 0);
		double time1 = rp.initial();
		bool more = true;
		const int CHECK = 16;                                      // sweeps between two looks at the flag (even)
		bool use_graph = true;
		while(more && rc == QPDE_OK) {
			qpde_step st;
			more = rp.next(st);
			const double t = st.t, h0 = rp.lapse(time1, t);
			cudaMemcpyAsync(v0, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream);
			int its = 0; bool notdone = true;
			while(notdone) {
				cudaMemcpyAsync(xprev, x, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream);
				k_policy<<<nblk, TB, axes_smem, stream>>>(md, g, R, xprev, v0, h0, pen);
				k_factor<<<lblk, 64, 0, stream>>>(nw, nb, R);
				*launches += 2;
				// CHECK sweeps, the flag of the last one, and its copy to the host: one graph, captured once
				// (the pointers of a group never change: an even number of sweeps ends where it began)
				auto sweeps = [&]() {
					double *xo = x, *xw = xn;
					for(int k = 0; k < CHECK; ++k) {
						if(k == CHECK - 1) cudaMemsetAsync(dunsettled, 0, sizeof(int), stream);
						k_rhs<<<nblk, TB, 0, stream>>>(n, nw, nb, R, v0, xo, rhs);
						k_lines<<<wblk, 128, 0, stream>>>(nw, nb, R, rhs, xo, xw, dunsettled);
						double *tmp = xo; xo = xw; xw = tmp;
					}
					cudaMemcpyAsync(hflag + 1, dunsettled, sizeof(int), cudaMemcpyDeviceToHost, stream);
				};
				if(!group && use_graph) {
					if(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
						sweeps();
						if(cudaStreamEndCapture(stream, &graph) != cudaSuccess || cudaGraphInstantiate(&group, graph, 0) != cudaSuccess) {
							group = nullptr; use_graph = false; cudaGetLastError();
						}
					} else { use_graph = false; cudaGetLastError(); }
				}
				int sweep = 0; bool settled = false;
				while(!settled && sweep < p->max_sweeps) {
					if(group) cudaGraphLaunch(group, stream); else sweeps();
					*launches += 2 * CHECK; sweep += CHECK;
					if(cudaStreamSynchronize(stream) != cudaSuccess) { rc = QPDE_ERR_CUDA; break; }
					settled = hflag[1] == 0;
				}
				if(rc != QPDE_OK) break;
				if(!settled) status = 2;
				++solves; sweeps_total += sweep;
				++its;
				cudaMemsetAsync(dflag, 0, sizeof(int), stream);
				k_relerr<<<nblk, TB, 0, stream>>>(n, x, xprev, p->tolerance, dflag); ++*launches;
				cudaMemcpyAsync(hflag, dflag, sizeof(int), cudaMemcpyDeviceToHost, stream);
				if(cudaStreamSynchronize(stream) != cudaSuccess) { rc = QPDE_ERR_CUDA; break; }
				notdone = hflag[0] != 0;
				if(notdone && its >= p->max_inner) { if(status == 0) status = 1; notdone = false; }
			}
			if(stats && stats->inner && n_steps < stats->inner_capacity) stats->inner[n_steps] = its;
			inner_sum += its; ++n_steps;
			time1 = t;
		}
		if(rc != QPDE_OK) break;
		k_outputs<<<nblk, TB, 0, stream>>>(nw, nb, R, x, dV, dQ, dZ, dmask); ++*launches;
		cudaMemcpyAsync(V_host, dV, sizeof(double) * n, cudaMemcpyDeviceToHost, stream);
		if(Q_host) cudaMemcpyAsync(Q_host, dQ, sizeof(double) * n, cudaMemcpyDeviceToHost, stream);
		if(Z_host) cudaMemcpyAsync(Z_host, dZ, sizeof(double) * n, cudaMemcpyDeviceToHost, stream);
		if(mask_host) cudaMemcpyAsync(mask_host, dmask, n, cudaMemcpyDeviceToHost, stream);
		if(cudaStreamSynchronize(stream) != cudaSuccess) rc = QPDE_ERR_CUDA;
	} while(0);
	const cudaError_t last = cudaGetLastError();
	if(rc == QPDE_OK && last != cudaSuccess) rc = QPDE_ERR_CUDA;
	if(rc != QPDE_OK) *error = std::string("hjbqvi2d: CUDA error: ") + cudaGetErrorString(last);
	if(group) cudaGraphExecDestroy(group);
	if(graph) cudaGraphDestroy(graph);
	if(hflag) cudaFreeHost(hflag);
	cudaFree(block);
	if(stats && rc == QPDE_OK) {
		stats->nodes_w = nw; stats->nodes_b = nb; stats->n_steps = n_steps; stats->status = status;
		stats->mean_inner = n_steps ? (double)inner_sum / (double)n_steps : 0.;
		stats->mean_sweeps = solves ? (double)sweeps_total / (double)solves : 0.;
	}
	return rc;
}

} // namespace qpde
