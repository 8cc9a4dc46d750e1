// QuantPDE_B200/Core.hpp — header-only host API over the C ABI (qpde_b200.h).
//
// Same call shape as the classes of parsiad/QuantPDE that sit on the 1-D
// time-stepping path, so that an example written against the reference reads
// the same against this header (see examples/vanilla_options_b200.cpp next to
// the reference's examples/vanilla_options.cpp:34-130):
//
//   reference (QuantPDE/src/...)                          here (namespace QuantPDE_B200)
//   Core/Axis.hpp:98,205-257,270-358,445-459  Axis         Axis  {ticks, special, uniform, range, *, +}
//   Core/Domain.hpp:1092-1172   RectilinearGrid1           RectilinearGrid1 (+ refined(k))
//   Modules/Operators/BlackScholes.hpp:114-134             BlackScholes1(grid, r, v, q)
//   Core/Stepper.hpp:27-41      Reverse/ForwardConstantStepper   ReverseConstantStepper(t0, T, dt), ForwardConstantStepper
//   Core/IterativeMethod.hpp:1243 ToleranceIteration       ToleranceIteration(tolerance, scale)
//   Core/Rannacher.hpp:123  CrankNicolson.hpp:97  BDF.hpp:496,570
//                                                          ReverseRannacher / ReverseCrankNicolson /
//                                                          ReverseBDFOne / ReverseBDFTwo / ... / ReverseBDFSix (grid, op)
//   Core/PenaltyMethod.hpp:281-301                         MinPenaltyMethodDifference1(grid, node, payoff, tol)
//   Bindings/Eigen.hpp:143-165  SparseLUSolver             B200Solver (owns the device context)
//   Core/IterativeMethod.hpp:1065-1079 Iteration::solve    stepper.solve(grid, payoff, root, solver)
//
// What differs, and why: the reference assembles and factorises one sparse
// matrix per inner iteration on the host; here solve() only *describes* the
// iteration tree and hands the whole backward sweep of every contract to one
// call of qpde_bs1d_sweep().  The batched entry point is Batch::solve(), which
// sweeps many such descriptions (same grid size, scheme and constraint type)
// in one launch.  Errors: every C-ABI failure is thrown as std::runtime_error
// carrying qpde_last_error(); there is no CPU fallback.
//
// Multi-TU safe (everything is inline); no global state.
#ifndef QUANT_PDE_B200_CORE_HPP
#define QUANT_PDE_B200_CORE_HPP

#include <cfloat>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "qpde_b200.h"

namespace QuantPDE_B200 {

typedef double Real;
typedef int Index;

constexpr Real epsilon = DBL_EPSILON;   // Core/EarlyInclude.hpp:61
constexpr Real tolerance = 1e-6;        // :62
constexpr Real scale = 1.;              // :63

////////////////////////////////////////////////////////////////////////////////
// Axis — Core/Axis.hpp
////////////////////////////////////////////////////////////////////////////////

class Axis {
	std::vector<Real> n;
public:
	Axis() {}
	Axis(std::initializer_list<Real> ticks) : n(ticks) {}
	explicit Axis(std::vector<Real> ticks) : n(std::move(ticks)) {}

	Index size() const { return (Index)n.size(); }
	Real operator[](Index i) const { return n[i]; }
	const Real *data() const { return n.data(); }

	// Axis::special (Core/Axis.hpp:445-459)
	static Axis special() {
		return Axis {
			0., .1, .2, .3, .4, .5, .6, .7, .80, .84, .88, .92, .94, .96, .98,
			1., 1.02, 1.04, 1.06, 1.08, 1.10, 1.14, 1.18, 1.23, 1.30, 1.40,
			1.50, 1.75, 2.25, 3.00, 7.50, 20., 100.
		};
	}
	// Core/Axis.hpp:205-211
	static Axis range(Real begin, Real step, Real end) {
		std::vector<Real> v((size_t)std::floor((end - begin) / step) + 1);
		for(size_t i = 0; i < v.size(); ++i) v[i] = begin + step * (Real)i;
		return Axis(std::move(v));
	}
	// Core/Axis.hpp:219-226
	static Axis uniform(Real begin, Real end, Index points) {
		const Real dx = (end - begin) / (points - 1);
		std::vector<Real> v((size_t)points);
		for(Index i = 0; i < points; ++i) v[i] = begin + dx * i;
		return Axis(std::move(v));
	}
	// points clustered around one feature by a sinh map (Core/Axis.hpp:51-90); the ticks go through libm's
	// asinh / sinh on the host exactly as the reference's do
	static Axis nonsymmetricCluster(Real begin, Real feature, Real end, Index points, Real intensity = 1.) {
		Real K = (feature - begin) / (end - begin);
		bool flipped = false;
		if(K < 0.5) { flipped = true; K = 1. - K; }
		const Real c = K / intensity;
		const Real xi_0 = std::asinh((0. - K) / c);
		const Real dxi = (std::asinh((1. - K) / c) - xi_0) / (points - 1);
		std::vector<Real> v((size_t)points);
		v[0] = begin;
		for(Index i = 1; i < points - 1; ++i) {
			if(flipped) v[points - 1 - i] = (1. - (K + c * std::sinh(xi_0 + i * dxi))) * (end - begin) + begin;
			else v[i] = ((K + c * std::sinh(xi_0 + i * dxi))) * (end - begin) + begin;
		}
		v[points - 1] = end;
		return Axis(std::move(v));
	}
	// Core/Axis.hpp:238-257 (an even number of points gives points + 1)
	static Axis cluster(Real begin, Real feature, Real end, Index points, Real intensity = 1.) {
		const Index p = (points + 1 + (points % 2 == 0 ? 1 : 0)) / 2;
		return nonsymmetricCluster(begin, feature, feature, p, intensity) + nonsymmetricCluster(feature, feature, end, p, intensity);
	}
	// union, Core/Axis.hpp:270-313
	friend Axis operator+(const Axis &a, const Axis &b) {
		std::vector<Real> c;
		size_t i = 0, j = 0;
		while(i < a.n.size() && j < b.n.size()) {
			if(a.n[i] == b.n[j])     { c.push_back(a.n[i++]); ++j; }
			else if(a.n[i] < b.n[j]) { c.push_back(a.n[i++]); }
			else                     { c.push_back(b.n[j++]); }
		}
		while(i < a.n.size()) c.push_back(a.n[i++]);
		while(j < b.n.size()) c.push_back(b.n[j++]);
		return Axis(std::move(c));
	}
	// scaling, Core/Axis.hpp:320-358
	friend Axis operator*(const Axis &a, Real c) {
		std::vector<Real> v(a.n);
		for(Real &x : v) x = x * c;
		return Axis(std::move(v));
	}
	friend Axis operator*(Real c, const Axis &a) { return a * c; }
};

////////////////////////////////////////////////////////////////////////////////
// RectilinearGrid1 — Core/Domain.hpp:1092-1172.  Keeps the coarse axis and the
// refinement count; the refined ticks are produced on the device with the
// reference's arithmetic (and on the host by nodes(), identically).
////////////////////////////////////////////////////////////////////////////////

class RectilinearGrid1 {
	Axis axis;
	int times;
public:
	RectilinearGrid1() : times(0) {}
	RectilinearGrid1(Axis a, int refine = 0) : axis(std::move(a)), times(refine) {}

	RectilinearGrid1 refined(int k = 1) const { return RectilinearGrid1(axis, times + k); }
	const Axis &coarse() const { return axis; }
	int refinement() const { return times; }
	Index size() const { return (Index)qpde_refined_size(axis.size(), times); }

	// refined ticks (Core/Domain.hpp:1126-1145)
	std::vector<Real> nodes() const {
		std::vector<Real> m((size_t)size());
		if(times <= 0) { for(Index i = 0; i < axis.size(); ++i) m[i] = axis[i]; return m; }
		const int tmp = 1 << times;
		size_t j = 0;
		m[j++] = axis[0];
		for(Index i = 1; i < axis.size(); ++i) {
			const Real dx = (axis[i] - axis[i - 1]) / tmp;
			for(int k = 1; k < tmp; ++k) m[j++] = k * dx + axis[i - 1];
			m[j++] = axis[i];
		}
		return m;
	}
};

////////////////////////////////////////////////////////////////////////////////
// Payoffs — Modules/Lambdas/Payoffs.hpp.  A Payoff is either one of the tagged
// closed forms (evaluated on the device) or any Real(Real) callable (evaluated
// on the host at the grid nodes and shipped as an array).
////////////////////////////////////////////////////////////////////////////////

class Payoff {
	int kind_;
	Real K_;
	std::function<Real(Real)> f_;
public:
	Payoff() : kind_(QPDE_PAYOFF_ARRAY), K_(0.) {}
	Payoff(int kind, Real K) : kind_(kind), K_(K) {}
	template <typename F, typename = decltype(std::declval<F>()(Real()))>
	Payoff(F f) : kind_(QPDE_PAYOFF_ARRAY), K_(0.), f_(std::move(f)) {}

	int kind() const { return kind_; }
	Real strike() const { return K_; }
	Real operator()(Real S) const {
		switch(kind_) {
		case QPDE_PAYOFF_PUT:          return S < K_ ? K_ - S : 0.;
		case QPDE_PAYOFF_CALL:         return S < K_ ? 0. : S - K_;
		case QPDE_PAYOFF_DIGITAL_PUT:  return S > K_ ? 0. : 1.;
		case QPDE_PAYOFF_DIGITAL_CALL: return S < K_ ? 0. : 1.;
		case QPDE_PAYOFF_STRADDLE:     return S < K_ ? K_ - S : S - K_;
		case QPDE_PAYOFF_K_MINUS_S:    return K_ - S;
		default:                       return f_(S);
		}
	}
};

inline Payoff putPayoff(Real K)         { return Payoff(QPDE_PAYOFF_PUT, K); }
inline Payoff callPayoff(Real K)        { return Payoff(QPDE_PAYOFF_CALL, K); }
inline Payoff digitalPutPayoff(Real K)  { return Payoff(QPDE_PAYOFF_DIGITAL_PUT, K); }
inline Payoff digitalCallPayoff(Real K) { return Payoff(QPDE_PAYOFF_DIGITAL_CALL, K); }
inline Payoff straddlePayoff(Real K)    { return Payoff(QPDE_PAYOFF_STRADDLE, K); }
inline Payoff exerciseValue(Real K)     { return Payoff(QPDE_PAYOFF_K_MINUS_S, K); }

////////////////////////////////////////////////////////////////////////////////
// Volatility argument of BlackScholes1: a constant, alpha / S, or any f(S)
// (the constant / Function1 cases of Controllable<1>, IterativeMethod.hpp:449-499)
////////////////////////////////////////////////////////////////////////////////

class Volatility {
	int model_;
	Real param_;
	std::function<Real(Real)> f_;
public:
	Volatility(Real sigma) : model_(QPDE_VOL_CONST), param_(sigma) {}
	Volatility(int model, Real param) : model_(model), param_(param) {}
	template <typename F, typename = decltype(std::declval<F>()(Real()))>
	Volatility(F f) : model_(QPDE_VOL_NODES), param_(0.), f_(std::move(f)) {}
	int model() const { return model_; }
	Real param() const { return param_; }
	Real operator()(Real S) const {
		return model_ == QPDE_VOL_CONST ? param_ : (model_ == QPDE_VOL_INVERSE_S ? param_ / S : f_(S));
	}
};
inline Volatility inverseSVolatility(Real alpha) { return Volatility(QPDE_VOL_INVERSE_S, alpha); }

////////////////////////////////////////////////////////////////////////////////
// The iteration tree
////////////////////////////////////////////////////////////////////////////////

class LinearSystem {
public:
	virtual ~LinearSystem() {}
};

// Control1(grid): marks a coefficient as a control (Core/IterativeMethod.hpp:449-499,
// examples/unequal_borrowing_lending_rates.cpp:108-118)
struct Control1 {
	explicit Control1(const RectilinearGrid1 &) {}
};

// BlackScholes1(grid, r, v, q) — Modules/Operators/BlackScholes.hpp:114-134
class BlackScholes1 : public LinearSystem {
public:
	const RectilinearGrid1 &grid;
	Real r; Volatility v; Real q;
	bool rIsControl;
	// b(t) of the operator: 0 (BlackScholes.hpp:228-230), or sourceTable[floor(t)] * S — what a subclass of the
	// reference's BlackScholes1 that overrides b(t) with a scalar of time times the grid amounts to
	// (GLWBOperator, examples/experimental/glwb.cpp:31-45)
	std::vector<Real> sourceTable;
	void setSource(std::vector<Real> table) { sourceTable = std::move(table); }
	BlackScholes1(const RectilinearGrid1 &grid, Real r, Volatility v, Real q)
			: grid(grid), r(r), v(std::move(v)), q(q), rIsControl(false) {}
	BlackScholes1(const RectilinearGrid1 &grid, Control1, Volatility v, Real q)
			: grid(grid), r(0.), v(std::move(v)), q(q), rIsControl(true) {}
};

// Jump amplitude densities — Modules/Lambdas/Densities (lognormal, doubleExponential) as
// examples/jump_diffusion.cpp:88-99 builds them.  Tagged forms: the density weights and kappa are
// produced by qpde_jump_setup(), which reproduces the reference's adaptive quadrature.
struct JumpDensity {
	int kind;        // 0 lognormal (Merton), 1 double exponential (Kou)
	Real params[3];
};
inline JumpDensity lognormal(Real mean, Real deviation) { return JumpDensity{0, {mean, deviation, 0.}}; }
inline JumpDensity doubleExponential(Real upProbability, Real upMeanReciprocal, Real downMeanReciprocal) {
	return JumpDensity{1, {upProbability, upMeanReciprocal, downMeanReciprocal}};
}

// BlackScholesJumpDiffusion1(grid, r, v, q, lambda, density) — Modules/Operators/BlackScholes.hpp:271-461:
// the jump term is explicit (the d'Halluin correlation integral of the previous iterand); runs in the
// jump-diffusion sweep (BDF1 / BDF2, no penalty, no controls).
class BlackScholesJumpDiffusion1 : public BlackScholes1 {
public:
	Real lambda; JumpDensity density;
	BlackScholesJumpDiffusion1(const RectilinearGrid1 &grid, Real r, Volatility v, Real q, Real lambda,
			JumpDensity density) : BlackScholes1(grid, r, std::move(v), q), lambda(lambda), density(density) {}
	// the reference's operator reads the stepper's iterand to form its right-hand side
	// (examples/jump_diffusion.cpp:113); here the sweep owns the iterand, so this only keeps the call shape
	template <typename I> void setIteration(I &) {}
};

class Iteration {
protected:
	std::vector<size_t> its;
public:
	virtual ~Iteration() {}
	const std::vector<size_t> &iterations() const { return its; }     // IterativeMethod.hpp:997
	Real meanIterations() const {
		Real s = 0.; for(size_t k : its) s += (Real)k;
		return its.empty() ? 0. : s / (Real)its.size();
	}
	friend class ReverseConstantStepper;
	friend class Batch;
};

// ToleranceIteration(tolerance, scale) — IterativeMethod.hpp:1243-1252
class ToleranceIteration : public Iteration {
public:
	Real tol, scl;
	int maxIterations;   // guard; the reference has none and would loop forever
	ToleranceIteration(Real tolerance = QuantPDE_B200::tolerance, Real scale = QuantPDE_B200::scale,
			int maxIterations = 100) : tol(tolerance), scl(scale), maxIterations(maxIterations) {}
};

class IterationNode : public LinearSystem {
public:
	Iteration *iteration;
	IterationNode() : iteration(nullptr) {}
	void setIteration(Iteration &it) { iteration = &it; }             // IterativeMethod.hpp:848
};

// Min/MaxPolicyIteration1_1(grid, controlGrid, system) — Core/PolicyIteration.hpp:109-174
class PolicyIteration1_1 : public IterationNode {
public:
	const RectilinearGrid1 &grid;
	std::vector<Real> controls;   // ticks of the control grid, in order
	const BlackScholes1 &system;
	bool maximise;
	PolicyIteration1_1(const RectilinearGrid1 &g, const RectilinearGrid1 &controlGrid,
			const BlackScholes1 &bs, bool maximise)
			: grid(g), controls(controlGrid.nodes()), system(bs), maximise(maximise) {
		if(!bs.rIsControl) throw std::invalid_argument("QuantPDE_B200: the operator has no Control1 coefficient");
	}
};
struct MinPolicyIteration1_1 : PolicyIteration1_1 {
	MinPolicyIteration1_1(const RectilinearGrid1 &g, const RectilinearGrid1 &c, const BlackScholes1 &bs)
			: PolicyIteration1_1(g, c, bs, false) {}
};
struct MaxPolicyIteration1_1 : PolicyIteration1_1 {
	MaxPolicyIteration1_1(const RectilinearGrid1 &g, const RectilinearGrid1 &c, const BlackScholes1 &bs)
			: PolicyIteration1_1(g, c, bs, true) {}
};

class Discretization : public IterationNode {
public:
	const RectilinearGrid1 &grid;
	const BlackScholes1 &op;
	const PolicyIteration1_1 *policy;   // non-null when the operator is wrapped in a policy iteration
	int scheme;
	bool forward;                       // a Forward* class: goes with a ForwardConstantStepper
	Discretization(const RectilinearGrid1 &g, const BlackScholes1 &o, int s, bool forward = false)
			: grid(g), op(o), policy(nullptr), scheme(s), forward(forward) {}
	Discretization(const RectilinearGrid1 &g, const PolicyIteration1_1 &p, int s, bool forward = false)
			: grid(g), op(p.system), policy(&p), scheme(s), forward(forward) {}
};
struct ReverseRannacher : Discretization {
	ReverseRannacher(const RectilinearGrid1 &g, const BlackScholes1 &o) : Discretization(g, o, QPDE_SCHEME_RANNACHER) {}
	ReverseRannacher(const RectilinearGrid1 &g, const PolicyIteration1_1 &p) : Discretization(g, p, QPDE_SCHEME_RANNACHER) {}
};
struct ReverseCrankNicolson : Discretization {
	ReverseCrankNicolson(const RectilinearGrid1 &g, const BlackScholes1 &o) : Discretization(g, o, QPDE_SCHEME_CN) {}
	ReverseCrankNicolson(const RectilinearGrid1 &g, const PolicyIteration1_1 &p) : Discretization(g, p, QPDE_SCHEME_CN) {}
};
struct ReverseBDFOne : Discretization {
	ReverseBDFOne(const RectilinearGrid1 &g, const BlackScholes1 &o) : Discretization(g, o, QPDE_SCHEME_BDF1) {}
	ReverseBDFOne(const RectilinearGrid1 &g, const PolicyIteration1_1 &p) : Discretization(g, p, QPDE_SCHEME_BDF1) {}
};
typedef ReverseBDFOne ReverseImplicitEuler;
struct ReverseBDFTwo : Discretization {
	ReverseBDFTwo(const RectilinearGrid1 &g, const BlackScholes1 &o) : Discretization(g, o, QPDE_SCHEME_BDF2) {}
	ReverseBDFTwo(const RectilinearGrid1 &g, const PolicyIteration1_1 &p) : Discretization(g, p, QPDE_SCHEME_BDF2) {}
};
// Core/BDF.hpp:595-942 (start-up through the lower orders, restart after every event)
#define QPDE_B200_BDF_CLASS(NAME, SCHEME) \
struct NAME : Discretization { \
	NAME(const RectilinearGrid1 &g, const BlackScholes1 &o) : Discretization(g, o, SCHEME) {} \
	NAME(const RectilinearGrid1 &g, const PolicyIteration1_1 &p) : Discretization(g, p, SCHEME) {} \
};
QPDE_B200_BDF_CLASS(ReverseBDFThree, QPDE_SCHEME_BDF3)
QPDE_B200_BDF_CLASS(ReverseBDFFour,  QPDE_SCHEME_BDF4)
QPDE_B200_BDF_CLASS(ReverseBDFFive,  QPDE_SCHEME_BDF5)
QPDE_B200_BDF_CLASS(ReverseBDFSix,   QPDE_SCHEME_BDF6)
#undef QPDE_B200_BDF_CLASS
// The forward-time classes (Rannacher<true>, CrankNicolson<true>, BDFOne<true> .. BDFSix<true>:
// Core/Rannacher.hpp:140, CrankNicolson.hpp:110, BDF.hpp:512-942): same formulas with
// h = t_new - t_old; they go with a ForwardConstantStepper.
#define QPDE_B200_FORWARD_CLASS(NAME, SCHEME) \
struct NAME : Discretization { \
	NAME(const RectilinearGrid1 &g, const BlackScholes1 &o) : Discretization(g, o, SCHEME, true) {} \
	NAME(const RectilinearGrid1 &g, const PolicyIteration1_1 &p) : Discretization(g, p, SCHEME, true) {} \
};
QPDE_B200_FORWARD_CLASS(ForwardRannacher,     QPDE_SCHEME_RANNACHER)
QPDE_B200_FORWARD_CLASS(ForwardCrankNicolson, QPDE_SCHEME_CN)
QPDE_B200_FORWARD_CLASS(ForwardBDFOne,        QPDE_SCHEME_BDF1)
QPDE_B200_FORWARD_CLASS(ForwardBDFTwo,        QPDE_SCHEME_BDF2)
QPDE_B200_FORWARD_CLASS(ForwardBDFThree,      QPDE_SCHEME_BDF3)
QPDE_B200_FORWARD_CLASS(ForwardBDFFour,       QPDE_SCHEME_BDF4)
QPDE_B200_FORWARD_CLASS(ForwardBDFFive,       QPDE_SCHEME_BDF5)
QPDE_B200_FORWARD_CLASS(ForwardBDFSix,        QPDE_SCHEME_BDF6)
#undef QPDE_B200_FORWARD_CLASS
typedef ForwardBDFOne ForwardImplicitEuler;

// MinPenaltyMethodDifference1(grid, constraintNode, payoff, tolerance) —
// Core/PenaltyMethod.hpp:281-301
class MinPenaltyMethodDifference1 : public IterationNode {
	std::vector<bool> mask;
public:
	const RectilinearGrid1 &grid;
	const Discretization &constraint;
	Payoff obstacle;
	Real tol;
	MinPenaltyMethodDifference1(const RectilinearGrid1 &g, const Discretization &d, Payoff payoff,
			Real tolerance = QuantPDE_B200::tolerance)
			: grid(g), constraint(d), obstacle(std::move(payoff)), tol(tolerance) {}
	// PenaltyMethod::constraintMask — PenaltyMethod.hpp:127-139
	const std::vector<bool> &constraintMask() const { return mask; }
	friend class ReverseConstantStepper;
	friend class Batch;
};

// The result of solve(): node values + piecewise-linear read-out
// (InterpolantWrapper1, Core/Interpolant.hpp:59-80,153-181,331-374)
class Solution {
	std::vector<Real> S, V;
public:
	Solution() {}
	Solution(std::vector<Real> S, std::vector<Real> V) : S(std::move(S)), V(std::move(V)) {}
	const std::vector<Real> &nodes() const { return S; }
	const std::vector<Real> &values() const { return V; }
	Real operator()(Real c) const {
		const Index n = (Index)S.size();
		Index idx = 0; Real w = 0.;
		if(c <= S[0]) { idx = 0; w = 1.; }
		else if(c >= S[n - 1]) { idx = n - 2; w = 0.; }
		else {
			Index lo = 0, hi = n - 2, mid = 0;
			while(lo <= hi) {
				mid = (lo + hi) / 2;
				if(c < S[mid]) hi = mid - 1;
				else if(c >= S[mid + 1]) lo = mid + 1;
				else { w = (S[mid + 1] - c) / (S[mid + 1] - S[mid]); break; }
			}
			idx = mid;
		}
		Real sum = 0.;
		if(w > epsilon) sum += w * V[idx];
		if(1. - w > epsilon) sum += (1. - w) * V[idx + 1];
		return sum;
	}
};

// The device context; stands where the reference passes a LinearSolver
// (SparseLUSolver, Bindings/Eigen.hpp:143-165).
class B200Solver {
	qpde_handle *h;
public:
	explicit B200Solver(int device = 0) : h(nullptr) {
		if(qpde_create(device, &h) != QPDE_OK) {
			throw std::runtime_error(std::string("QuantPDE_B200: ") + qpde_last_error(nullptr));
		}
	}
	~B200Solver() { qpde_destroy(h); }
	B200Solver(const B200Solver &) = delete;
	B200Solver &operator=(const B200Solver &) = delete;
	qpde_handle *handle() const { return h; }
	void check(int rc) const {
		if(rc != QPDE_OK) throw std::runtime_error(std::string("QuantPDE_B200: ") + qpde_last_error(h));
	}
};

// The tagged forms of the README's dividend event  [D] (const Interpolant1 &V, Real S) { return
// V( max( S - D, 0. ) ); }  (README.md:357-375, tutorial.cpp:82-99) and its proportional sibling
// V( (1. - D) * S ): an arbitrary Event<1> lambda cannot cross the C ABI.
struct Dividend {
	int kind; Real amount;
	Dividend() : kind(QPDE_DIVIDEND_NONE), amount(0.) {}
	Dividend(int kind, Real amount) : kind(kind), amount(amount) {}
	static Dividend cash(Real D) { return Dividend(QPDE_DIVIDEND_CASH, D); }
	static Dividend proportional(Real D) { return Dividend(QPDE_DIVIDEND_PROPORTIONAL, D); }
};

// A general event transform — what the reference takes as a lambda
//     [..] (const Interpolant1 &V, Real S) { return ...; }            (Core/Event.hpp:60-184)
// — written as an expression over S, V(.), constants and per-event parameters.  A lambda cannot cross the
// C ABI; an expression can: it is lowered to the postfix program of include/qpde_b200.h (QPDE_EV_*) and
// evaluated at every node inside the sweep.  Every operator is one IEEE operation, so an expression that
// spells the lambda in its order reproduces it bit for bit.  Example (examples/experimental/glwb.cpp:84-121):
//     EventExpression S = EventExpression::S(), R = EventExpression::param(0), pen = EventExpression::param(1);
//     auto best = max(max(c(1. + bonus) * V(S / c(1. + bonus)), V(max(S - c(GQ), c(0.))) + R * c(GQ)),
//                     select(S > c(GQ), R * (c(GQ) + (c(1.) - pen) * (S - c(GQ))), c(-infinity)));
//     stepper.addEvents(E, best, table);          // table[e] = { R_e, pen_e } for the event at T e / E
class EventExpression {
	std::vector<int32_t> code;          // postfix words; constants are interned into consts at lowering
	std::vector<Real> consts;
	static EventExpression binary(const EventExpression &a, const EventExpression &b, int op) {
		EventExpression r;
		r.consts = a.consts;
		r.code = a.code;
		for(size_t k = 0; k < b.code.size(); ++k) {
			r.code.push_back(b.code[k]);
			if(b.code[k] == QPDE_EV_CONST) { r.consts.push_back(b.consts[(size_t)b.code[k + 1]]); r.code.push_back((int32_t)r.consts.size() - 1); ++k; }
			else if(b.code[k] == QPDE_EV_PARAM) { r.code.push_back(b.code[k + 1]); ++k; }
		}
		r.code.push_back(op);
		return r;
	}
	static EventExpression unary(const EventExpression &a, int op) { EventExpression r = a; r.code.push_back(op); return r; }
public:
	static EventExpression S() { EventExpression r; r.code.push_back(QPDE_EV_S); return r; }
	static EventExpression constant(Real v) { EventExpression r; r.consts.push_back(v); r.code.push_back(QPDE_EV_CONST); r.code.push_back(0); return r; }
	static EventExpression param(int index) { EventExpression r; r.code.push_back(QPDE_EV_PARAM); r.code.push_back(index); return r; }
	friend EventExpression operator+(const EventExpression &a, const EventExpression &b) { return binary(a, b, QPDE_EV_ADD); }
	friend EventExpression operator-(const EventExpression &a, const EventExpression &b) { return binary(a, b, QPDE_EV_SUB); }
	friend EventExpression operator*(const EventExpression &a, const EventExpression &b) { return binary(a, b, QPDE_EV_MUL); }
	friend EventExpression operator/(const EventExpression &a, const EventExpression &b) { return binary(a, b, QPDE_EV_DIV); }
	friend EventExpression operator-(const EventExpression &a) { return unary(a, QPDE_EV_NEG); }
	friend EventExpression operator>(const EventExpression &a, const EventExpression &b) { return binary(a, b, QPDE_EV_GT); }
	// max(a, b) = b > a ? b : a, the "if(tmp > best) best = tmp" of the reference's examples
	friend EventExpression max(const EventExpression &a, const EventExpression &b) { return binary(a, b, QPDE_EV_MAX); }
	friend EventExpression min(const EventExpression &a, const EventExpression &b) { return binary(a, b, QPDE_EV_MIN); }
	// the interpolant of the solution before the event, at any price
	friend EventExpression V(const EventExpression &at) { return unary(at, QPDE_EV_V); }
	friend EventExpression select(const EventExpression &condition, const EventExpression &yes, const EventExpression &no) {
		return binary(binary(condition, yes, -1), no, QPDE_EV_SELECT);      // -1: a join without an op word (removed below)
	}
	// the program words (END appended) and the constants they index
	void lower(std::vector<int32_t> &words, std::vector<Real> &constants) const {
		words.clear();
		for(size_t k = 0; k < code.size(); ++k) {
			if(code[k] == -1) continue;
			words.push_back(code[k]);
			if(code[k] == QPDE_EV_CONST || code[k] == QPDE_EV_PARAM) { words.push_back(code[k + 1]); ++k; }
		}
		words.push_back(QPDE_EV_END);
		constants = consts;
	}
	bool empty() const { return code.empty(); }
};
inline EventExpression c(Real v) { return EventExpression::constant(v); }

// ReverseConstantStepper(t0, T, dt) — Core/Stepper.hpp:27-36
class ReverseConstantStepper : public Iteration {
	Real t0, T, dt;
	ToleranceIteration *inner;
	int nExercise, nDividend;
	Payoff exercise;
	Dividend dividend_;
	bool dividendAddedFirst;
	EventExpression eventExpression_;
	std::vector<std::vector<Real>> eventTable_;
	std::vector<Real> eventTimes_;      // explicit event times (else uniform T e / E)
protected:
	Real vTarget, vScale;      // > 0: ReverseVariableStepper
	int stepLimit_;
	bool forward_;             // ForwardConstantStepper
public:
	ReverseConstantStepper(Real startTime, Real endTime, Real dt)
			: t0(startTime), T(endTime), dt(dt), inner(nullptr), nExercise(0), nDividend(0),
			  dividendAddedFirst(false), vTarget(0.), vScale(1.), stepLimit_(0), forward_(false) {
		if(startTime != 0.) throw std::invalid_argument("QuantPDE_B200: startTime must be 0");
	}
	void setInnerIteration(ToleranceIteration &it) { inner = &it; }   // IterativeMethod.hpp:977
	// stepper.add(T / E * e, exerciseEvent, grid) for e = 0 .. E-1 (tutorial.cpp:103-114)
	void addExerciseDates(int E, Payoff exerciseValue) {
		if(nDividend > 0 && nDividend != E) throw std::invalid_argument("QuantPDE_B200: exercise and dividend dates must coincide");
		nExercise = E; exercise = std::move(exerciseValue);
	}
	// stepper.add(T / E * e, dividendEvent, grid) for e = 0 .. E-1 (README.md:357-375).  As in the
	// reference the order of the add() calls matters: events at one time are applied in descending
	// add() order by the reverse sweep (IterativeMethod.hpp:1340-1352), so calling this BEFORE
	// addExerciseDates models a dividend paid before the holder's right to exercise.
	void addDividendDates(int E, Dividend d) {
		if(nExercise > 0 && nExercise != E) throw std::invalid_argument("QuantPDE_B200: exercise and dividend dates must coincide");
		nDividend = E; dividend_ = d; dividendAddedFirst = nExercise == 0;
	}

	// stepper.add(T / E * e, event, grid) for e = 0 .. E-1 (forward: e = 1 .. E) with a general transform:
	// table[e] holds the parameters param(0), param(1), ... of the event at that time (may be empty rows)
	void addEvents(int E, EventExpression event, std::vector<std::vector<Real>> table = {}) {
		if(nExercise > 0 || nDividend > 0) throw std::invalid_argument("QuantPDE_B200: a general event takes the place of the tagged exercise / dividend events");
		if(!table.empty() && (int)table.size() != E) throw std::invalid_argument("QuantPDE_B200: one parameter row per event time");
		nExercise = E; exercise = Payoff(QPDE_EXERCISE_NONE, 0.);
		eventExpression_ = std::move(event); eventTable_ = std::move(table);
	}
	// the same at explicit times: stepper.add(times[e], event, grid); table[e] goes with the e-th SMALLEST time.
	// A time equal to the initial time of the sweep is allowed (the reference then takes a zero-length first step).
	void addEvents(std::vector<Real> times, EventExpression event, std::vector<std::vector<Real>> table = {}) {
		addEvents((int)times.size(), std::move(event), std::move(table));
		eventTimes_ = std::move(times);
	}
	const std::vector<Real> &eventTimes() const { return eventTimes_; }
	const EventExpression &eventExpression() const { return eventExpression_; }
	const std::vector<std::vector<Real>> &eventTable() const { return eventTable_; }

	Real expiry() const { return T; }
	Real timestep() const { return dt; }
	int exerciseDates() const { return nExercise; }
	const Payoff &exercisePayoff() const { return exercise; }
	Real variableTarget() const { return vTarget; }
	Real variableScale() const { return vScale; }
	int stepLimit() const { return stepLimit_ > 0 ? stepLimit_ : 16 * (int)std::ceil(T / dt) + 64; }
	int eventDates() const { return nExercise > 0 ? nExercise : nDividend; }
	const Dividend &dividend() const { return dividend_; }
	bool hasDividends() const { return nDividend > 0 && dividend_.kind != QPDE_DIVIDEND_NONE; }
	// the dividend transform goes first in the reverse sweep iff its events were add()ed last; a forward
	// sweep applies events at one time in ascending add() order (TimeOrder, IterativeMethod.hpp:1340-1352)
	bool dividendFirstInSweep() const {
		return hasDividends() && nExercise > 0 && (forward_ ? dividendAddedFirst : !dividendAddedFirst);
	}
	bool isForward() const { return forward_; }
	ToleranceIteration *innerIteration() const { return inner; }

	// Iteration::solve(domain, initialCondition, root, solver) — IterativeMethod.hpp:1065-1079
	inline Solution solve(const RectilinearGrid1 &grid, const Payoff &payoff, IterationNode &root,
			B200Solver &solver);
	friend class Batch;
};

// ForwardConstantStepper(t0, T, dt) — Core/Stepper.hpp:41 (TimeIteration<true>,
// IterativeMethod.hpp:1494-1495): time runs from 0 up to T, solve() takes the condition at t = 0;
// addExerciseDates / addDividendDates(E, ...) place their events at T e / E for e = 1 .. E (an
// event may not sit at the initial time).  Goes with the Forward* discretisations.
class ForwardConstantStepper : public ReverseConstantStepper {
public:
	ForwardConstantStepper(Real startTime, Real endTime, Real dt) : ReverseConstantStepper(startTime, endTime, dt) {
		forward_ = true;
	}
};

// ReverseVariableStepper(t0, T, dt, target, scale) — Core/Stepper.hpp:60-126: dt is the first
// step; afterwards dt *= target / (relativeError(V^{n+1}, V^n, scale) + epsilon).  The number
// of steps is data dependent: setStepLimit bounds it (iterations()[0] reports what was taken;
// solve() throws if the limit was hit).
class ReverseVariableStepper : public ReverseConstantStepper {
public:
	ReverseVariableStepper(Real startTime, Real endTime, Real dt, Real target, Real scale_ = scale)
			: ReverseConstantStepper(startTime, endTime, dt) {
		if(!(target > 0.) || !(scale_ > 0.)) throw std::invalid_argument("QuantPDE_B200: target and scale must be > 0");
		vTarget = target; vScale = scale_;
	}
	void setStepLimit(int limit) { stepLimit_ = limit; }
};

////////////////////////////////////////////////////////////////////////////////
// Batch — many solve() descriptions in one device sweep
////////////////////////////////////////////////////////////////////////////////

class Batch {
	struct Item {
		const RectilinearGrid1 *grid; Payoff payoff; IterationNode *root; ReverseConstantStepper *stepper;
	};
	std::vector<Item> items;
public:
	void add(ReverseConstantStepper &stepper, const RectilinearGrid1 &grid, Payoff payoff, IterationNode &root) {
		items.push_back(Item{&grid, std::move(payoff), &root, &stepper});
	}
	size_t size() const { return items.size(); }

	// Runs every added problem; results come back in order.
	std::vector<Solution> solve(B200Solver &solver, int kernel = QPDE_KERNEL_AUTO) {
		const int C = (int)items.size();
		if(C == 0) return {};
		// --- decode the first tree; the rest must have the same shape
		auto decode = [] (IterationNode *root, const Discretization *&disc,
				MinPenaltyMethodDifference1 *&pen) {
			pen = dynamic_cast<MinPenaltyMethodDifference1 *>(root);
			disc = pen ? &pen->constraint : dynamic_cast<const Discretization *>(root);
			if(!disc) throw std::invalid_argument("QuantPDE_B200: unsupported root node");
		};
		const Discretization *d0; MinPenaltyMethodDifference1 *p0;
		decode(items[0].root, d0, p0);
		const int N = items[0].grid->size(), nAxis = items[0].grid->coarse().size();
		const int refine = items[0].grid->refinement();
		const bool american = p0 != nullptr;
		const int nEx = items[0].stepper->exerciseDates(), nEv = items[0].stepper->eventDates();
		const bool variableSteps = items[0].stepper->variableTarget() > 0.;
		std::vector<Real> vTargets;
		const bool dividends = items[0].stepper->hasDividends();
		const int divKind = items[0].stepper->dividend().kind;
		const bool divFirst = items[0].stepper->dividendFirstInSweep();
		std::vector<Real> divAmount;
		ToleranceIteration *tol0 = items[0].stepper->innerIteration();
		if(american && !tol0) throw std::invalid_argument("QuantPDE_B200: penalty root needs setInnerIteration()");

		std::vector<Real> axis((size_t)C * nAxis), r(C), sig(C), q(C), T(C), dt(C), K(C), spot(C);
		std::vector<Real> payArr, obsArr, sigArr;
		int payKind = items[0].payoff.kind(), obsKind = american ? p0->obstacle.kind() : QPDE_PAYOFF_PUT;
		int volModel = d0->op.v.model();
		const bool policy = d0->policy != nullptr;
		const int nControls = policy ? (int)d0->policy->controls.size() : 0;
		std::vector<Real> controls;
		if(policy && !tol0) throw std::invalid_argument("QuantPDE_B200: policy iteration needs setInnerIteration()");
		// jump diffusion: density weights / kappa / log-grid per distinct (grid, density)
		const BlackScholesJumpDiffusion1 *j0 = dynamic_cast<const BlackScholesJumpDiffusion1 *>(&d0->op);
		const bool jump = j0 != nullptr;
		const int nfft = jump ? qpde_jump_nfft(N) : 0;
		std::vector<Real> jLambda, jKappa, jX0, jDx, jWeights;
		struct JumpKey { std::vector<Real> axis; JumpDensity density; size_t at; Real x0, dx, kappa; };
		std::vector<JumpKey> jumpKeys;
		std::vector<int> jIndex;        // weight vector of contract c
		for(int c = 0; c < C; ++c) {
			const Item &it = items[c];
			const Discretization *d; MinPenaltyMethodDifference1 *p;
			decode(it.root, d, p);
			if(d->forward != it.stepper->isForward()) {
				throw std::invalid_argument("QuantPDE_B200: a Forward* discretisation goes with a ForwardConstantStepper "
					"(and a Reverse* one with a ReverseConstantStepper)");
			}
			if(it.grid->size() != N || it.grid->coarse().size() != nAxis || it.grid->refinement() != refine
					|| d->scheme != d0->scheme || (p != nullptr) != american
					|| it.stepper->isForward() != items[0].stepper->isForward()
					|| it.stepper->exerciseDates() != nEx || it.stepper->eventDates() != nEv
					|| it.stepper->hasDividends() != dividends
					|| (it.stepper->variableTarget() > 0.) != variableSteps
					|| (variableSteps && it.stepper->variableScale() != items[0].stepper->variableScale())
					|| (dividends && (it.stepper->dividend().kind != divKind || it.stepper->dividendFirstInSweep() != divFirst))
					|| d->op.v.model() != volModel
					|| it.payoff.kind() != payKind || (american && p->obstacle.kind() != obsKind)
					|| (dynamic_cast<const BlackScholesJumpDiffusion1 *>(&d->op) != nullptr) != jump
					|| (d->policy != nullptr) != policy
					|| (policy && ((int)d->policy->controls.size() != nControls
						|| d->policy->maximise != d0->policy->maximise))) {
				throw std::invalid_argument("QuantPDE_B200: all problems of a batch must share grid size, "
					"scheme, constraint type, payoff kind and volatility model");
			}
			for(int k = 0; k < nAxis; ++k) axis[(size_t)c * nAxis + k] = it.grid->coarse()[k];
			r[c] = d->op.r; sig[c] = d->op.v.param(); q[c] = d->op.q;
			if(policy) controls.insert(controls.end(), d->policy->controls.begin(), d->policy->controls.end());
			if(jump) {
				const BlackScholesJumpDiffusion1 &j = static_cast<const BlackScholesJumpDiffusion1 &>(d->op);
				const Axis &ax = it.grid->coarse();
				const std::vector<Real> coarse(ax.data(), ax.data() + ax.size());
				size_t k = 0;
				for(; k < jumpKeys.size(); ++k) {
					const JumpKey &key = jumpKeys[k];
					if(key.axis == coarse && key.density.kind == j.density.kind
							&& key.density.params[0] == j.density.params[0] && key.density.params[1] == j.density.params[1]
							&& key.density.params[2] == j.density.params[2]) break;
				}
				if(k == jumpKeys.size()) {
					JumpKey key; key.axis = coarse; key.density = j.density; key.at = jumpKeys.size();
					const std::vector<Real> S = it.grid->nodes();
					std::vector<Real> wts((size_t)nfft);
					const int rc = qpde_jump_setup(N, S.data(), j.density.kind, j.density.params,
						&key.x0, &key.dx, &key.kappa, wts.data());
					if(rc != QPDE_OK) throw std::runtime_error(std::string("QuantPDE_B200: qpde_jump_setup: ") + qpde_last_error(nullptr));
					jumpKeys.push_back(key);
					jWeights.insert(jWeights.end(), wts.begin(), wts.end());
				}
				jLambda.push_back(j.lambda); jKappa.push_back(jumpKeys[k].kappa);
				jX0.push_back(jumpKeys[k].x0); jDx.push_back(jumpKeys[k].dx);
				jIndex.push_back((int)k);
			}
			T[c] = it.stepper->expiry(); dt[c] = it.stepper->timestep();
			if(dividends) divAmount.push_back(it.stepper->dividend().amount);
			if(variableSteps) vTargets.push_back(it.stepper->variableTarget());
			K[c] = it.payoff.kind() != QPDE_PAYOFF_ARRAY ? it.payoff.strike()
				: (american && p->obstacle.kind() != QPDE_PAYOFF_ARRAY ? p->obstacle.strike()
				: (nEx > 0 && it.stepper->exercisePayoff().kind() > QPDE_PAYOFF_ARRAY ? it.stepper->exercisePayoff().strike() : 0.));
			// one strike per contract crosses the C ABI: every generator of the tree must use it
			if(it.payoff.kind() != QPDE_PAYOFF_ARRAY) {
				if((american && p->obstacle.kind() != QPDE_PAYOFF_ARRAY && p->obstacle.strike() != K[c])
						|| (nEx > 0 && it.stepper->exercisePayoff().kind() > QPDE_PAYOFF_ARRAY
							&& it.stepper->exercisePayoff().strike() != K[c])) {
					throw std::invalid_argument("QuantPDE_B200: payoff, obstacle and exercise generators of one "
						"problem must share their strike (pass the obstacle as an array otherwise)");
				}
			} else if(american && p->obstacle.kind() != QPDE_PAYOFF_ARRAY && nEx > 0
					&& it.stepper->exercisePayoff().kind() > QPDE_PAYOFF_ARRAY
					&& it.stepper->exercisePayoff().strike() != K[c]) {
				throw std::invalid_argument("QuantPDE_B200: obstacle and exercise generators of one problem must share their strike");
			}
			spot[c] = K[c];
			const bool needNodes = payKind == QPDE_PAYOFF_ARRAY || (american && obsKind == QPDE_PAYOFF_ARRAY)
				|| volModel == QPDE_VOL_NODES;
			if(needNodes) {
				const std::vector<Real> S = it.grid->nodes();
				if(payKind == QPDE_PAYOFF_ARRAY) for(Real s : S) payArr.push_back(it.payoff(s));
				if(american && obsKind == QPDE_PAYOFF_ARRAY) for(Real s : S) obsArr.push_back(p->obstacle(s));
				if(volModel == QPDE_VOL_NODES) for(Real s : S) sigArr.push_back(d->op.v(s));
			}
		}

		qpde_bs1d_batch b = qpde_bs1d_batch();
		b.abi_version = QPDE_ABI_VERSION;
		b.n_contracts = C;
		b.axis = axis.data(); b.axis_stride = nAxis; b.n_axis = nAxis; b.refine = refine;
		b.r = r.data(); b.sigma = sig.data(); b.q = q.data();
		b.vol_model = volModel;
		if(volModel == QPDE_VOL_NODES) { b.sigma_nodes = sigArr.data(); b.sigma_nodes_stride = N; }
		b.T = T.data(); b.dt = dt.data(); b.scheme = d0->scheme; b.max_steps = 0;
		b.forward = items[0].stepper->isForward() ? 1 : 0;
		b.n_exercise = nEv;
		b.exercise_kind = nEx > 0 ? items[0].stepper->exercisePayoff().kind() : QPDE_EXERCISE_NONE;
		if(dividends) { b.dividend_kind = divKind; b.dividend_first = divFirst ? 1 : 0; b.dividend = divAmount.data(); }
		for(int cc = 0; cc < C; ++cc) {
			const Discretization *dd; MinPenaltyMethodDifference1 *pp;
			decode(items[cc].root, dd, pp);
			if(dd->op.sourceTable != d0->op.sourceTable || items[cc].stepper->eventTimes() != items[0].stepper->eventTimes()) {
				throw std::invalid_argument("QuantPDE_B200: all problems of a batch must share the operator's source table and the event times");
			}
		}
		if(!d0->op.sourceTable.empty()) { b.source_table = d0->op.sourceTable.data(); b.n_source = (int32_t)d0->op.sourceTable.size(); }
		if(!items[0].stepper->eventTimes().empty()) b.event_times = items[0].stepper->eventTimes().data();
		std::vector<int32_t> evWords; std::vector<Real> evConsts, evParams;
		if(!items[0].stepper->eventExpression().empty()) {
			items[0].stepper->eventExpression().lower(evWords, evConsts);
			const size_t width = items[0].stepper->eventTable().empty() ? 0 : items[0].stepper->eventTable()[0].size();
			for(int cc = 0; cc < C; ++cc) {
				std::vector<int32_t> w; std::vector<Real> k;
				items[cc].stepper->eventExpression().lower(w, k);
				const auto &tab = items[cc].stepper->eventTable();
				if(w != evWords || k != evConsts || (tab.empty() ? 0 : tab[0].size()) != width) {
					throw std::invalid_argument("QuantPDE_B200: all problems of a batch must share the event expression (parameters may differ)");
				}
				for(const auto &row : tab) {
					if(row.size() != width) throw std::invalid_argument("QuantPDE_B200: ragged event parameter table");
					evParams.insert(evParams.end(), row.begin(), row.end());
				}
			}
			b.event_program = evWords.data(); b.event_program_length = (int32_t)evWords.size();
			b.event_consts = evConsts.data(); b.n_event_consts = (int32_t)evConsts.size();
			b.event_params = evParams.data(); b.n_event_params = (int32_t)width;
			b.event_params_stride = (int64_t)nEv * (int64_t)width;
		}
		b.strike = K.data();
		b.payoff_kind = payKind; b.obstacle_kind = obsKind;
		if(payKind == QPDE_PAYOFF_ARRAY) { b.payoff = payArr.data(); b.payoff_stride = N; }
		if(american && obsKind == QPDE_PAYOFF_ARRAY) { b.obstacle = obsArr.data(); b.obstacle_stride = N; }
		b.american = american ? 1 : 0;
		const bool iterate = american || policy;
		b.max_inner = iterate ? tol0->maxIterations : 1;
		b.tolerance = iterate ? tol0->tol : tolerance;
		b.scale = iterate ? tol0->scl : scale;
		if(policy) {
			b.n_controls = nControls; b.policy_max = d0->policy->maximise ? 1 : 0;
			b.controls = controls.data(); b.controls_stride = nControls;
		}
		std::vector<Real> jPerContract;
		if(jump) {
			if(american || policy) throw std::invalid_argument("QuantPDE_B200: jump diffusion takes no penalty or policy node");
			b.jump_lambda = jLambda.data(); b.jump_kappa = jKappa.data(); b.jump_x0 = jX0.data(); b.jump_dx = jDx.data();
			b.n_fft = nfft;
			if(jumpKeys.size() == 1) {
				b.jump_weights = jWeights.data(); b.jump_weights_stride = 0;      // one vector shared by the batch
			} else {
				jPerContract.resize((size_t)C * nfft);
				for(int c = 0; c < C; ++c) {
					std::copy(jWeights.begin() + (size_t)jIndex[c] * nfft, jWeights.begin() + (size_t)(jIndex[c] + 1) * nfft,
						jPerContract.begin() + (size_t)c * nfft);
				}
				b.jump_weights = jPerContract.data(); b.jump_weights_stride = nfft;
			}
		}
		b.penalty_tolerance = american ? p0->tol : tolerance;
		b.spot = spot.data();
		b.kernel = kernel;

		int maxSteps = 0;
		for(int c = 0; c < C; ++c) {
			const int bound = variableSteps ? items[c].stepper->stepLimit() : (int)std::floor(T[c] / dt[c]) + 3 + 2 * nEv;
			if(bound > maxSteps) maxSteps = bound;
		}
		if(variableSteps) {
			b.variable_target = vTargets.data(); b.variable_scale = items[0].stepper->variableScale();
			b.max_steps = maxSteps;
		}
		std::vector<Real> V((size_t)C * N), S((size_t)C * N);
		std::vector<int32_t> nSteps(C), status(C);
		std::vector<uint8_t> inner((size_t)C * maxSteps), active((size_t)C * N);
		qpde_bs1d_result res = qpde_bs1d_result();
		res.V = V.data(); res.V_stride = N; res.S = S.data(); res.S_stride = N;
		res.n_steps = nSteps.data(); res.status = status.data();
		res.inner = inner.data(); res.inner_stride = maxSteps;
		res.active = active.data(); res.active_stride = N;
		solver.check(qpde_bs1d_sweep(solver.handle(), &b, &res));

		std::vector<Solution> out;
		for(int c = 0; c < C; ++c) {
			const Item &it = items[c];
			if(status[c] == 3) {
				throw std::runtime_error("QuantPDE_B200: contract " + std::to_string(c)
					+ " needs more time steps than ReverseVariableStepper::setStepLimit allows");
			}
			if(status[c] != 0) {
				throw std::runtime_error("QuantPDE_B200: the penalty iteration of contract "
					+ std::to_string(c) + " hit ToleranceIteration::maxIterations (the reference would not terminate)");
			}
			out.emplace_back(std::vector<Real>(S.begin() + (size_t)c * N, S.begin() + (size_t)(c + 1) * N),
				std::vector<Real>(V.begin() + (size_t)c * N, V.begin() + (size_t)(c + 1) * N));
			// iteration counts, as Iteration::iterations() reports them
			it.stepper->its.assign(1, (size_t)nSteps[c]);
			const Discretization *d; MinPenaltyMethodDifference1 *p;
			decode(it.root, d, p);
			if(p || d->policy) {
				ToleranceIteration *ti = it.stepper->innerIteration();
				ti->its.clear();
				for(int s = 0; s < nSteps[c]; ++s) ti->its.push_back(inner[(size_t)c * maxSteps + s]);
			}
			if(p) {
				p->mask.assign(active.begin() + (size_t)c * N, active.begin() + (size_t)(c + 1) * N);
			}
		}
		return out;
	}
};

////////////////////////////////////////////////////////////////////////////////
// GMWB — the 2-D (W, A) impulse-control HJBQVI of examples/hjbqvi/gmwb.cpp.  The reference's
// HJBQVI<2,1,1> takes the model as lambdas (Modules/HJBQVI/HJBQVI.hpp:590-754); a C ABI cannot, so
// the model's functional forms are fixed (include/qpde_b200.h, qpde_gmwb2d) and its parameters are
// the members below, with the example's values as defaults.  solve(refinement) does what
// HJBQVI::solve(refinement) does with usePenalizedScheme() and HJBQVILinearBoundary /
// HJBQVIZeroDiffusionRightBoundary: `timesteps`, the W / A axes and the impulse grid are refined
// `refinement` times (gmwb.cpp:61-64, HJBQVI.hpp:1285-1340).
////////////////////////////////////////////////////////////////////////////////
class GMWB {
public:
	Real T, r, eta, sigma, G, kappa, c;      // gmwb.cpp:41-59 (kappa: "penalty rate k")
	Real w_0;
	int timesteps, controlPoints;
	Axis wAxis, aAxis;
	Real scalingFactor, iterationTolerance;  // HJBQVI defaults: 1e-2, 1e-6
	int maxIterations, maxSweeps;            // guards (the reference has none)

	GMWB() : T(10.), r(.05), eta(0.), sigma(.2), G(10.), kappa(.1), c(0.), w_0(100.), timesteps(32),
			controlPoints(3),
			wAxis({0., 5., 10., 15., 20., 25., 30., 35., 40., 45., 50., 55., 60., 65., 70., 72.5, 75., 77.5, 80.,
				82., 84., 86., 88., 90., 91., 92., 93., 94., 95., 96., 97., 98., 99., 100., 101., 102., 103., 104.,
				105., 106., 107., 108., 109., 110., 112., 114., 116., 118., 120., 123., 126., 130., 135., 140.,
				145., 150., 160., 175., 200., 225., 250., 300., 500., 750., 1000.}),    // gmwb.cpp:92-105
			aAxis(Axis::uniform(0., 100., 51)),                                          // gmwb.cpp:108
			scalingFactor(1e-2), iterationTolerance(1e-6), maxIterations(200), maxSweeps(500) {}

	struct Result {
		std::vector<Real> W, A, V;   // V[ia * W.size() + iw] (W fastest, Core/Domain.hpp:1416-1425)
		int steps; Real meanPolicyIterations;
		// PiecewiseLinear<2>::interpolate (Core/Interpolant.hpp:153-181,331-374): clamped, terms with
		// weight <= epsilon skipped
		Real operator()(Real w, Real a) const {
			auto bracket = [] (const std::vector<Real> &x, Real c, int &i, Real &wt) {
				const int n = (int)x.size();
				if(c <= x[0]) { i = 0; wt = 1.; return; }
				if(c >= x[n - 1]) { i = n - 2; wt = 0.; return; }
				int lo = 0, hi = n - 2, mid = 0; wt = 0.;
				while(lo <= hi) {
					mid = (lo + hi) / 2;
					if(c < x[mid]) hi = mid - 1;
					else if(c >= x[mid + 1]) lo = mid + 1;
					else { wt = (x[mid + 1] - c) / (x[mid + 1] - x[mid]); break; }
				}
				i = mid;
			};
			int iw, ia; Real fw, fa;
			bracket(W, w, iw, fw); bracket(A, a, ia, fa);
			const size_t nw = W.size();
			const Real wt[4] = { fw * fa, (1. - fw) * fa, fw * (1. - fa), (1. - fw) * (1. - fa) };
			const size_t at[4] = { ia * nw + iw, ia * nw + iw + 1, (ia + 1) * nw + iw, (ia + 1) * nw + iw + 1 };
			Real sum = 0.;
			for(int m = 0; m < 4; ++m) if(wt[m] > epsilon) sum += wt[m] * V[at[m]];
			return sum;
		}
	};

	Result solve(B200Solver &solver, int refinement) const {
		const Real controls[2] = { 0., G };                     // gmwb.cpp:111-114
		const Axis impulse = Axis::uniform(0., 1., controlPoints);
		qpde_gmwb2d g = qpde_gmwb2d();
		g.abi_version = QPDE_ABI_VERSION; g.refine = refinement;
		g.w_axis = wAxis.data(); g.n_w_axis = wAxis.size();
		g.a_axis = aAxis.data(); g.n_a_axis = aAxis.size();
		g.controls = controls; g.n_controls = 2;
		g.impulse_axis = impulse.data(); g.n_impulse_axis = impulse.size();
		g.T = T;
		g.timesteps = timesteps;
		for(int k = 0; k < refinement; ++k) g.timesteps *= 2;
		g.max_inner = maxIterations;
		g.r = r; g.eta = eta; g.sigma = sigma; g.kappa = kappa; g.c = c;
		g.scaling_factor = scalingFactor; g.tolerance = iterationTolerance; g.max_sweeps = maxSweeps;
		int nw = 0, na = 0;
		solver.check(qpde_gmwb2d_nodes(&g, &nw, &na));
		Result res;
		res.W.resize(nw); res.A.resize(na); res.V.resize((size_t)nw * na);
		qpde_gmwb2d_stats st = qpde_gmwb2d_stats();
		solver.check(qpde_gmwb2d_solve(solver.handle(), &g, res.V.data(), res.W.data(), res.A.data(), &st));
		if(st.status != 0) {
			throw std::runtime_error(st.status == 1 ? "QuantPDE_B200: GMWB policy iteration hit maxIterations"
				: "QuantPDE_B200: GMWB line sweeps did not settle");
		}
		res.steps = st.n_steps; res.meanPolicyIterations = st.mean_inner;
		return res;
	}
};

////////////////////////////////////////////////////////////////////////////////
// ExchangeRate — examples/hjbqvi/exchange_rate.cpp on Modules/HJBQVI/HJBQVI.hpp: the combined stochastic /
// impulse control of an exchange rate (Mundaca & Oksendal), a 1-D HJBQVI under the penalised scheme.  The
// reference takes the coefficient functions as lambdas; a lambda cannot cross the C ABI, so the model is a
// tagged form of qpde_hjbqvi1d_batch with the example's parameters as members.  solve(solver, problems, k)
// sweeps a batch of problems in one launch (one CTA per problem).
////////////////////////////////////////////////////////////////////////////////

class ExchangeRate {
public:
	Real x_min, x_max, m;                    // exchange_rate.cpp:36-47: state domain and optimal parity
	Real q_min, q_max;                       // interest rate differential bounds
	Real T, rho, sigma, a, b, lambda, c;     // :52-69
	Real intensity;                          // clustering of the nodes around the parity
	int points, stochasticControlPoints, impulseControlPoints, timesteps;   // :40-42, :72
	Real scalingFactor, iterationTolerance;  // HJBQVI defaults: 1e-2, 1e-6
	int maxIterations, maxSweeps;            // guards (the reference has none)

	ExchangeRate() : x_min(-2.), x_max(2.), m(0.), q_min(-.07), q_max(.07), T(10.), rho(.02), sigma(.3), a(.25),
			b(3.), lambda(1.), c(.1), intensity(10.), points(32), stochasticControlPoints(9), impulseControlPoints(16),
			timesteps(16), scalingFactor(1e-2), iterationTolerance(1e-6), maxIterations(200), maxSweeps(500) {}

	Axis xAxis() const { return Axis::cluster(x_min, m, x_max, points, intensity); }                      // :79-85
	Axis controlAxis() const { return Axis::uniform(q_min, q_max, stochasticControlPoints); }             // :88
	Axis impulseAxis() const { return Axis::cluster(x_min, m, x_max, impulseControlPoints, intensity); }  // :91-97

	struct Result {
		std::vector<Real> X, V, Q, Z;        // nodes, value, controls (NaN where the other control acts)
		int stochasticControlNodes, impulseControlNodes, steps;
		Real meanPolicyIterations, meanSolverIterations;
		// PiecewiseLinear<1>::interpolate (Core/Interpolant.hpp:153-181,331-374)
		Real operator()(Real x) const {
			const int n = (int)X.size();
			int i; Real w;
			if(x <= X[0]) { i = 0; w = 1.; }
			else if(x >= X[n - 1]) { i = n - 2; w = 0.; }
			else {
				int lo = 0, hi = n - 2, mid = 0; w = 0.;
				while(lo <= hi) {
					mid = (lo + hi) / 2;
					if(x < X[mid]) hi = mid - 1;
					else if(x >= X[mid + 1]) lo = mid + 1;
					else { w = (X[mid + 1] - x) / (X[mid + 1] - X[mid]); break; }
				}
				i = mid;
			}
			Real sum = 0.;
			if(w > epsilon) sum += w * V[i];
			if(1. - w > epsilon) sum += (1. - w) * V[i + 1];
			return sum;
		}
	};

	Result solve(B200Solver &solver, int refinement) const {
		return solve(solver, std::vector<ExchangeRate>(1, *this), refinement)[0];
	}

	// Problems of one batch share the discretisation counts (points, timesteps, T, tolerances); the model
	// parameters and the axes are per problem.
	static std::vector<Result> solve(B200Solver &solver, const std::vector<ExchangeRate> &problems, int refinement) {
		if(problems.empty()) return std::vector<Result>();
		const ExchangeRate &first = problems[0];
		const size_t P = problems.size();
		std::vector<Real> xa, qa, za, par;
		int nx = 0, nq = 0, nz = 0;
		for(const ExchangeRate &e : problems) {
			const Axis x = e.xAxis(), q = e.controlAxis(), z = e.impulseAxis();
			if(&e == &first) { nx = x.size(); nq = q.size(); nz = z.size(); }
			if(x.size() != nx || q.size() != nq || z.size() != nz || e.timesteps != first.timesteps || e.T != first.T
					|| e.scalingFactor != first.scalingFactor || e.iterationTolerance != first.iterationTolerance) {
				throw std::invalid_argument("QuantPDE_B200: the problems of one ExchangeRate batch share the grid sizes, T, timesteps and tolerances");
			}
			xa.insert(xa.end(), x.data(), x.data() + nx); qa.insert(qa.end(), q.data(), q.data() + nq);
			za.insert(za.end(), z.data(), z.data() + nz);
			const Real pp[7] = { e.rho, e.sigma, e.a, e.b, e.m, e.lambda, e.c };
			par.insert(par.end(), pp, pp + 7);
		}
		qpde_hjbqvi1d_batch g = qpde_hjbqvi1d_batch();
		g.abi_version = QPDE_ABI_VERSION; g.model = QPDE_HJBQVI1D_EXCHANGE_RATE;
		g.n_problems = (int32_t)P; g.refine = refinement;
		g.x_axis = xa.data(); g.n_x_axis = nx; g.x_axis_stride = nx;
		g.control_axis = qa.data(); g.n_control_axis = nq; g.control_axis_stride = nq;
		g.impulse_axis = za.data(); g.n_impulse_axis = nz; g.impulse_axis_stride = nz;
		g.params = par.data(); g.n_params = 7; g.params_stride = 7;
		g.T = first.T; g.timesteps = first.timesteps;
		for(int k = 0; k < refinement; ++k) g.timesteps *= 2;       // HJBQVI.hpp:629-632
		g.max_inner = first.maxIterations; g.max_sweeps = first.maxSweeps;
		g.scaling_factor = first.scalingFactor; g.tolerance = first.iterationTolerance;
		int N = 0, cq = 0, cz = 0;
		solver.check(qpde_hjbqvi1d_nodes(&g, &N, &cq, &cz));
		std::vector<Real> V(P * N), X(P * N), Q(P * N), Z(P * N), mi(P), ms(P);
		std::vector<int32_t> steps(P), status(P);
		qpde_hjbqvi1d_result r = qpde_hjbqvi1d_result();
		r.V = V.data(); r.X = X.data(); r.Q = Q.data(); r.Z = Z.data();
		r.mean_inner = mi.data(); r.mean_sweeps = ms.data(); r.n_steps = steps.data(); r.status = status.data();
		solver.check(qpde_hjbqvi1d_solve(solver.handle(), &g, &r));
		std::vector<Result> out(P);
		for(size_t p = 0; p < P; ++p) {
			if(status[p] != 0) {
				throw std::runtime_error(status[p] == 1 ? "QuantPDE_B200: ExchangeRate policy iteration hit maxIterations"
					: "QuantPDE_B200: ExchangeRate lagged sweeps did not settle");
			}
			Result &res = out[p];
			res.X.assign(X.begin() + p * N, X.begin() + (p + 1) * N); res.V.assign(V.begin() + p * N, V.begin() + (p + 1) * N);
			res.Q.assign(Q.begin() + p * N, Q.begin() + (p + 1) * N); res.Z.assign(Z.begin() + p * N, Z.begin() + (p + 1) * N);
			res.stochasticControlNodes = cq; res.impulseControlNodes = cz; res.steps = steps[p];
			res.meanPolicyIterations = mi[p]; res.meanSolverIterations = ms[p];
		}
		return out;
	}
};

////////////////////////////////////////////////////////////////////////////////
// OptimalConsumption — examples/hjbqvi/optimal_consumption.cpp on Modules/HJBQVI/HJBQVI.hpp: optimal
// consumption with fixed and proportional transaction costs, a 2-D (stock w, bank b) HJBQVI under the
// penalised scheme whose impulses (buying and selling stock) point both ways in the state.  The model is a
// tagged form of qpde_hjbqvi2d with the example's parameters as members.
////////////////////////////////////////////////////////////////////////////////

class OptimalConsumption {
public:
	Real T, rho, r, mu, sigma, gamma, lambda, c, q_max;   // optimal_consumption.cpp:36-63
	Real w_0, b_0;                                        // :66-67: the convergence test point
	Real boundary;                                        // :79
	int timesteps, controlPoints, spatialPoints;          // :70, :73, :80
	Real scalingFactor, iterationTolerance;               // HJBQVI defaults: 1e-2, 1e-6
	int maxIterations, maxSweeps;                         // guards (the reference has none)

	OptimalConsumption() : T(40.), rho(.1), r(.07), mu(.11), sigma(.3), gamma(.3), lambda(.1), c(.05), q_max(100.),
			w_0(45.2), b_0(45.2), boundary(200.), timesteps(32), controlPoints(16), spatialPoints(20),
			scalingFactor(1e-2), iterationTolerance(1e-6), maxIterations(200), maxSweeps(20000) {}

	struct Result {
		std::vector<Real> W, B, V, Q, Z;     // V[ib * W.size() + iw] (w fastest, Core/Domain.hpp:1416-1425)
		int stochasticControlNodes, impulseControlNodes, steps;
		Real meanPolicyIterations, meanSolverIterations;
		// PiecewiseLinear<2>::interpolate (Core/Interpolant.hpp:153-181,331-374)
		Real operator()(Real w, Real b) const {
			auto bracket = [] (const std::vector<Real> &x, Real c, int &i, Real &wt) {
				const int n = (int)x.size();
				if(c <= x[0]) { i = 0; wt = 1.; return; }
				if(c >= x[n - 1]) { i = n - 2; wt = 0.; return; }
				int lo = 0, hi = n - 2, mid = 0; wt = 0.;
				while(lo <= hi) {
					mid = (lo + hi) / 2;
					if(c < x[mid]) hi = mid - 1;
					else if(c >= x[mid + 1]) lo = mid + 1;
					else { wt = (x[mid + 1] - c) / (x[mid + 1] - x[mid]); break; }
				}
				i = mid;
			};
			int iw, ib; Real fw, fb;
			bracket(W, w, iw, fw); bracket(B, b, ib, fb);
			const size_t nw = W.size();
			// corner order of Interpolant.hpp:343-368: bit j of the corner index set <=> lower node in dimension j
			const Real wt[4] = { (1. - fw) * (1. - fb), fw * (1. - fb), (1. - fw) * fb, fw * fb };
			const size_t at[4] = { (ib + 1) * nw + iw + 1, (ib + 1) * nw + iw, ib * nw + iw + 1, ib * nw + iw };
			Real sum = 0.;
			for(int m = 0; m < 4; ++m) if(wt[m] > epsilon) sum += wt[m] * V[at[m]];
			return sum;
		}
	};

	Result solve(B200Solver &solver, int refinement) const {
		const Axis axis = Axis::uniform(0., boundary, spatialPoints);                                     // :80-81
		const Axis controls = Axis::uniform(0., q_max, controlPoints), impulses = Axis::uniform(0., 1., controlPoints);   // :110-113
		const Real par[8] = { rho, r, mu, sigma, gamma, lambda, c, boundary };
		qpde_hjbqvi2d g = qpde_hjbqvi2d();
		g.abi_version = QPDE_ABI_VERSION; g.model = QPDE_HJBQVI2D_OPTIMAL_CONSUMPTION; g.refine = refinement;
		g.w_axis = axis.data(); g.n_w_axis = axis.size(); g.b_axis = axis.data(); g.n_b_axis = axis.size();
		g.control_axis = controls.data(); g.n_control_axis = controls.size();
		g.impulse_axis = impulses.data(); g.n_impulse_axis = impulses.size();
		g.params = par; g.n_params = 8;
		g.T = T; g.timesteps = timesteps;
		for(int k = 0; k < refinement; ++k) g.timesteps *= 2;       // HJBQVI.hpp:629-632
		g.max_inner = maxIterations; g.max_sweeps = maxSweeps;
		g.scaling_factor = scalingFactor; g.tolerance = iterationTolerance;
		int nw = 0, nb = 0;
		solver.check(qpde_hjbqvi2d_nodes(&g, &nw, &nb));
		Result res;
		const size_t n = (size_t)nw * nb;
		res.W.resize(nw); res.B.resize(nb); res.V.resize(n); res.Q.resize(n); res.Z.resize(n);
		qpde_hjbqvi2d_stats st = qpde_hjbqvi2d_stats();
		solver.check(qpde_hjbqvi2d_solve(solver.handle(), &g, res.V.data(), res.Q.data(), res.Z.data(), nullptr,
			res.W.data(), res.B.data(), &st));
		if(st.status != 0) {
			throw std::runtime_error(st.status == 1 ? "QuantPDE_B200: OptimalConsumption policy iteration hit maxIterations"
				: "QuantPDE_B200: OptimalConsumption line sweeps did not settle");
		}
		int cn = controlPoints - 1;
		for(int k = 0; k < refinement; ++k) cn *= 2;
		res.stochasticControlNodes = cn + 1; res.impulseControlNodes = cn + 1;
		res.steps = st.n_steps; res.meanPolicyIterations = st.mean_inner; res.meanSolverIterations = st.mean_sweeps;
		return res;
	}
};

inline Solution ReverseConstantStepper::solve(const RectilinearGrid1 &grid, const Payoff &payoff,
		IterationNode &root, B200Solver &solver) {
	Batch batch;
	batch.add(*this, grid, payoff, root);
	return batch.solve(solver)[0];
}

} // namespace QuantPDE_B200

#endif
