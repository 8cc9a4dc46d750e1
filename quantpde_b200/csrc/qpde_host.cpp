// qpde_host.cpp — host-only helpers of the C ABI (no GPU needed): the set-up
// work the reference does once when a BlackScholesJumpDiffusion operator is
// constructed, in the reference's arithmetic so that the device sweep starts
// from the same numbers.
//
//   kappa            Modules/Operators/BlackScholes.hpp:53-62
//   log-grid x0, dx  Modules/Operators/BlackScholes.hpp:275-297
//   density weights  Modules/Operators/BlackScholes.hpp:299-329
//   quadrature       Core/Integral.hpp:122-212 (expanding window for infinite
//                    bounds), :243-296 (trapezoid), :367-456 (adaptive refinement)
//   densities        Modules/Lambdas/Densities.hpp:21-66
//   grid refinement  Core/Domain.hpp:1092-1151
//
// Compiled with -ffp-contract=off: no fused multiply-adds, as in the
// reference's plain -O3 x86-64 build (CMakeLists.txt:9,29).
#include <cfloat>
#include <cmath>
#include <limits>

#include "../../include/qpde_b200.h"

namespace {

const double kTolerance = 1e-6;   // QuantPDE::tolerance (Core/EarlyInclude.hpp:62)
const int kMaxDepth = 64;         // AdaptiveQuadrature::maxDepth (Core/Integral.hpp:370)

struct Density {
	int kind; double a, b, c;
	double operator()(double x) const {
		if(x == 0.) return 0.;
		if(kind == QPDE_DENSITY_LOGNORMAL) {
			const double mu = a, sigma = b;
			return 1. / (x * sigma * std::sqrt(2. * M_PI))
				* std::exp( -((std::log(x) - mu) * (std::log(x) - mu)) / (2. * sigma * sigma) );
		}
		const double p = a, eta_1 = b, eta_2 = c;
		return (x >= 1) ? (p * eta_1 * std::pow(x, -eta_1 - 1))
		                : ((1 - p) * eta_2 * std::pow(x, eta_2 - 1));
	}
};

// One-interval trapezoid on [lo, hi]
template <typename F>
double trapezoid(const F &f, double lo, double hi) {
	double scale = 1. / 2;
	scale *= (hi - lo) / 1;
	double sum = f(lo) * 1.;
	sum += f(hi) * 1.;
	return scale * sum;
}

// Bisect while the two halves disagree with the parent by more than the tolerance
template <typename F>
double bisect(const F &f, double parent, double lo, double hi, int depth) {
	const double mid = (lo + hi) / 2.;
	const double first = trapezoid(f, lo, mid), second = trapezoid(f, mid, hi);
	double total = 0.;
	total += first;
	total += second;
	if(depth > 0 && std::abs((total - parent) / total) > kTolerance) {
		total = 0.;
		total += bisect(f, first, lo, mid, depth - 1);
		total += bisect(f, second, mid, hi, depth - 1);
	}
	return total;
}

template <typename F>
double adaptive(const F &f, double lo, double hi) {
	return bisect(f, trapezoid(f, lo, hi), lo, hi, kMaxDepth);
}

// Integral over [lo, hi] where either bound may be infinite: the finite window
// doubles away from its anchor until the value settles.
template <typename F>
double integrate(const F &f, double lo, double hi) {
	const double inf = std::numeric_limits<double>::infinity();
	const bool openLeft = lo == -inf, openRight = hi == inf;
	if(!openLeft && !openRight) return adaptive(f, lo, hi);
	double anchor, a = lo, b = hi;
	if(openLeft && openRight) { a = -1.; b = 1.; anchor = 0.; }
	else if(openLeft)         { a = b - 1.; anchor = b; }
	else                      { b = a + 1.; anchor = a; }
	double previous = adaptive(f, a, b);
	for(;;) {
		if(openLeft)  a += a - anchor;
		if(openRight) b += b - anchor;
		const double value = adaptive(f, a, b);
		if(std::abs((value - previous) / value) < kTolerance) return value;
		previous = value;
	}
}

} // namespace

extern "C" {

int qpde_jump_nfft(int n_nodes) {
	if(n_nodes < 3) return QPDE_ERR_INVALID;
	int N = 2;
	while(N < n_nodes) N *= 2;
	return N;
}

int qpde_jump_setup(int n, const double *S, int density_kind, const double *params,
		double *x0_out, double *dx_out, double *kappa_out, double *weights) {
	if(n < 3 || !S || !params || !x0_out || !dx_out || !kappa_out || !weights) return QPDE_ERR_INVALID;
	if(density_kind != QPDE_DENSITY_LOGNORMAL && density_kind != QPDE_DENSITY_DOUBLE_EXPONENTIAL) {
		return QPDE_ERR_INVALID;
	}
	const Density g = { density_kind, params[0], params[1], params[2] };
	const int N = qpde_jump_nfft(n);
	const int first = S[0] < DBL_EPSILON ? 1 : 0;
	const double x0 = std::log(S[first]);
	const double xf = std::log(S[n - 2]);
	const double dx = (xf - x0) / (N - 1);
	*x0_out = x0; *dx_out = dx;
	// E[eta] - 1 with eta = e^y
	*kappa_out = integrate([&] (double y) { return std::exp(2 * y) * g(std::exp(y)); },
			-std::numeric_limits<double>::infinity(), std::numeric_limits<double>::infinity()) - 1.;
	// mass of the log-jump density around each log-grid offset, wrapped order
	auto fbar = [&] (double x) { return g(std::exp(x)) * std::exp(x); };
	for(int i = 0; i < N; ++i) {
		const int k = i <= N / 2 ? i : i - N;
		weights[i] = integrate(fbar, dx * (-.5 + k), dx * (.5 + k));
	}
	return QPDE_OK;
}

int qpde_refine_axis(const double *axis, int n_axis, int refine, double *out) {
	if(!axis || !out || n_axis < 2 || refine < 0 || refine > 20) return QPDE_ERR_INVALID;
	const int parts = 1 << refine;
	int j = 0;
	out[j++] = axis[0];
	for(int i = 1; i < n_axis; ++i) {
		const double width = (axis[i] - axis[i - 1]) / parts;
		for(int k = 1; k < parts; ++k) out[j++] = k * width + axis[i - 1];
		out[j++] = axis[i];
	}
	return j;
}

} // extern "C"
