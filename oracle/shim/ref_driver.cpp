// TEST INFRASTRUCTURE ONLY (oracle/_ref).  Never linked into the product path.
//
// A client of the UNMODIFIED reference headers under /root/reference, compiled
// against oracle/shim/qpde_linalg_binding.hpp through the reference's own
// QUANT_PDE_BOUND seam (QuantPDE/Core:6-9).  It wires the reference classes the
// way the reference's example programs do
//   american / european : examples/vanilla_options.cpp:34-130
//   bermudan            : examples/tutorial.cpp:58-141
//   hjb                 : examples/unequal_borrowing_lending_rates.cpp:90-170
//   jump                : examples/jump_diffusion.cpp:60-160
// and dumps every node value (hex floats), the per-step inner iteration counts,
// the final penalty active set and the wall-clock of stepper.solve (timed as
// Modules/Utilities/Results.hpp:103-108 does).
//
// Usage: qpde_ref <mode> key=value ...        (see usage() below)

#include <QuantPDE/src/Core/EarlyInclude.hpp>
#include "qpde_linalg_binding.hpp"
#include <QuantPDE/Core>
#include <QuantPDE/Modules/Lambdas>
#include <QuantPDE/Modules/Operators>
#include <QuantPDE/Modules/HJBQVI>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>

using namespace QuantPDE;
using namespace QuantPDE::Modules;

namespace {

typedef std::map<std::string, std::string> Args;

double num(const Args &a, const char *key, double dflt) {
	auto it = a.find(key);
	return it == a.end() ? dflt : std::strtod(it->second.c_str(), nullptr);
}

long integer(const Args &a, const char *key, long dflt) {
	auto it = a.find(key);
	return it == a.end() ? dflt : std::strtol(it->second.c_str(), nullptr, 10);
}

std::string str(const Args &a, const char *key, const char *dflt) {
	auto it = a.find(key);
	return it == a.end() ? std::string(dflt) : it->second;
}

void dumpReals(const char *name, const std::vector<double> &v) {
	std::printf("%s %zu\n", name, v.size());
	for(size_t i = 0; i < v.size(); ++i) {
		std::printf("%a%c", v[i], (i + 1) % 8 == 0 ? '\n' : ' ');
	}
	if(v.size() % 8 != 0) std::printf("\n");
}

void dumpInts(const char *name, const std::vector<long> &v) {
	std::printf("%s %zu\n", name, v.size());
	for(size_t i = 0; i < v.size(); ++i) {
		std::printf("%ld%c", v[i], (i + 1) % 32 == 0 ? '\n' : ' ');
	}
	if(v.size() % 32 != 0) std::printf("\n");
}

// Base axis selection: "special" = K/100-scaled Axis::special times 100 (what
// vanilla_options builds with S_0 = K), "tutorial" = the 34-tick hand-picked
// axis of examples/tutorial.cpp:41-56 scaled by K/100, "jump" = the
// jump_diffusion.cpp:166 axis.
Axis baseAxis(const std::string &kind, Real S0, Real K) {
	if(kind == "special") {
		return (S0 * Axis::special) + (K * Axis::special);
	}
	if(kind == "tutorial") {
		Axis a {
			0., 10., 20., 30., 40., 50., 60., 70., 75., 80., 84., 88.,
			92., 94., 96., 98., 100., 102., 104., 106., 108., 110., 114.,
			118., 123., 130., 140., 150., 175., 225., 300., 750., 2000.,
			10000.
		};
		return a * (K / 100.);
	}
	std::fprintf(stderr, "unknown axis kind %s\n", kind.c_str());
	std::exit(2);
}

template <typename G, typename I>
std::vector<double> nodeValues(const G &grid, const I &solution) {
	const Axis &S = grid[0];
	std::vector<double> v((size_t)S.size());
	for(Index i = 0; i < S.size(); ++i) v[i] = solution(S[i]);
	return v;
}

template <typename G>
std::vector<double> nodes(const G &grid) {
	const Axis &S = grid[0];
	std::vector<double> v((size_t)S.size());
	for(Index i = 0; i < S.size(); ++i) v[i] = S[i];
	return v;
}

////////////////////////////////////////////////////////////////////////////////
// vanilla: European / American put or call, Rannacher (default), plain
// Crank-Nicolson, BDF1 or BDF2 time stepping, penalty method for American
////////////////////////////////////////////////////////////////////////////////

// per-event parameters of the GLWB-like event (tests hand the same numbers to the device as event_params)
static Real glwbSurvival(int e) { return 1. - .01 * (e + 1); }
static Real glwbPenalty(int e) { const Real p = .03 - .01 * e; return p > 0. ? p : 0.; }

int runVanilla(const Args &a) {
	const Real K = num(a, "K", 100.), S0 = num(a, "S0", K);
	const Real r = num(a, "r", .04), vol = num(a, "vol", .2);
	const Real q = num(a, "q", 0.), T = num(a, "T", 1.);
	const long steps = integer(a, "steps", 12);
	const int refine = (int) integer(a, "refine", 0);
	const bool american = integer(a, "american", 1) != 0;
	const bool call = integer(a, "call", 0) != 0;
	const std::string scheme = str(a, "scheme", "rannacher");
	const int repeat = (int) integer(a, "repeat", 1);

	RectilinearGrid1 coarse( baseAxis(str(a, "axis", "special"), S0, K) );
	auto grid = coarse.refined(refine);

	Function1 payoff = call ? callPayoff(K) : putPayoff(K);

	double seconds = 0.;
	std::vector<double> values;
	std::vector<long> inner, mask;
	long timesteps = 0;

	// examples/vanilla_options.cpp:52-64: ReverseVariableStepper(0, T, T / N2k, 1 / factor)
	const Real variableTarget = num(a, "variable_target", 0.);

	// optional events (none by default)
	const int E = (int) integer(a, "E", 0);
	const Real D = num(a, "dividend", 0.);
	const std::string dividendKind = str(a, "dividend_kind", "none");
	const bool dividendFirst = integer(a, "dividend_first", 0) != 0;  // first in the reverse sweep = add()ed last
	const bool withExercise = integer(a, "exercise", 0) != 0;
	const std::string eventKind = str(a, "event", "tagged");
	const Real bonus = num(a, "bonus", .06), GQ = num(a, "GQ", 5.);
	auto exercise = [K] (const Interpolant1 &V, Real S) {
		return max(V(S), K - S);
	};
	auto cashDividend = [D] (const Interpolant1 &V, Real S) {
		return V( max( S - D, 0. ) );
	};
	auto proportionalDividend = [D] (const Interpolant1 &V, Real S) {
		return V( (1. - D) * S );
	};

	for(int rep = 0; rep < repeat; ++rep) {
		BlackScholes1 bs(grid, r, vol, q);
		std::unique_ptr<ReverseConstantStepper> constantStepper;
		std::unique_ptr<ReverseVariableStepper> variableStepper;
		if(variableTarget > 0.) variableStepper.reset(new ReverseVariableStepper(0., T, T / steps, variableTarget));
		else constantStepper.reset(new ReverseConstantStepper(0., T, T / steps));
		Iteration &stepper = variableTarget > 0. ? (Iteration &) *variableStepper : (Iteration &) *constantStepper;

		// README.md:357-375 on top of the penalty method: E event times T e / E with a dividend transform
		// and / or the exercise transform, add()ed in the order the arguments ask for
		if(E > 0 && constantStepper && eventKind == "glwb") {
			// the three-way choice of examples/experimental/glwb.cpp:84-121 (non-withdrawal with bonus, withdrawal,
			// surrender) as the event at every T e / E, with per-event survival and penalty rates
			for(int e = 0; e < E; ++e) {
				const Real Rn = glwbSurvival(e), pen = glwbPenalty(e);
				auto event = [=] (const Interpolant1 &V, Real S) {
					Real best = (1. + bonus) * V(S / (1. + bonus));
					{
						const Real S_plus = max(S - GQ, 0.);
						const Real tmp = V(S_plus) + Rn * GQ;
						if(tmp > best) best = tmp;
					}
					if(S > GQ) {
						const Real tmp = Rn * (GQ + (1. - pen) * (S - GQ));
						if(tmp > best) best = tmp;
					}
					return best;
				};
				constantStepper->add(T / E * e, event, grid);
			}
		} else if(E > 0 && constantStepper) {
			auto addDividends = [&] () {
				for(int e = 0; e < E; ++e) {
					if(dividendKind == "cash") constantStepper->add(T / E * e, cashDividend, grid);
					else if(dividendKind == "proportional") constantStepper->add(T / E * e, proportionalDividend, grid);
				}
			};
			if(!dividendFirst) addDividends();
			if(withExercise) for(int e = 0; e < E; ++e) constantStepper->add(T / E * e, exercise, grid);
			if(dividendFirst) addDividends();
		}

		std::unique_ptr<IterationNode> discretization;
		if(scheme == "rannacher") {
			discretization.reset(new ReverseRannacher(grid, bs));
		} else if(scheme == "cn") {
			discretization.reset(new ReverseCrankNicolson(grid, bs));
		} else if(scheme == "bdf1") {
			discretization.reset(new ReverseBDFOne(grid, bs));
		} else if(scheme == "bdf2") {
			discretization.reset(new ReverseBDFTwo(grid, bs));
		} else if(scheme == "bdf3") {
			discretization.reset(new ReverseBDFThree(grid, bs));
		} else if(scheme == "bdf4") {
			discretization.reset(new ReverseBDFFour(grid, bs));
		} else if(scheme == "bdf5") {
			discretization.reset(new ReverseBDFFive(grid, bs));
		} else if(scheme == "bdf6") {
			discretization.reset(new ReverseBDFSix(grid, bs));
		} else {
			std::fprintf(stderr, "unknown scheme\n");
			return 2;
		}
		discretization->setIteration(stepper);

		ToleranceIteration tolerance;
		std::unique_ptr<MinPenaltyMethodDifference1> penalty;
		IterationNode *root = discretization.get();
		if(american) {
			penalty.reset(new MinPenaltyMethodDifference1(
					grid, *discretization, payoff));
			penalty->setIteration(tolerance);
			stepper.setInnerIteration(tolerance);
			root = penalty.get();
		}

		SparseLUSolver solver;

		auto begin = std::chrono::steady_clock::now();
		auto solution = stepper.solve(grid, payoff, *root, solver);
		auto end = std::chrono::steady_clock::now();
		seconds += std::chrono::duration<double>(end - begin).count();

		if(rep == 0) {
			values = nodeValues(grid, solution);
			timesteps = (long) stepper.iterations()[0];
			if(american) {
				for(auto k : tolerance.iterations()) inner.push_back((long)k);
				for(bool b : penalty->constraintMask()) mask.push_back(b);
			}
			std::printf("value %a\n", (double) solution(S0));
		}
	}

	std::printf("seconds %.9g\n", seconds / repeat);
	std::printf("steps %ld\n", timesteps);
	dumpReals("S", nodes(grid));
	dumpReals("V", values);
	dumpInts("inner", inner);
	dumpInts("mask", mask);
	return 0;
}

////////////////////////////////////////////////////////////////////////////////
// forward: the forward-time classes (TimeIteration<true>, IterativeMethod.hpp:1494-1495;
// ForwardConstantStepper, Stepper.hpp:41; ForwardRannacher / CrankNicolson / BDFOne .. BDFSix):
// V_t + L V = 0 marched from t = 0 to T from the payoff, constant or alpha / S volatility,
// optional penalty method, optional events at T e / E, e = 1 .. E (an event at the initial
// time is not allowed, IterativeMethod.hpp:1411)
////////////////////////////////////////////////////////////////////////////////

int runForward(const Args &a) {
	const Real K = num(a, "K", 100.);
	const Real r = num(a, "r", .04), vol = num(a, "vol", .2), alpha = num(a, "alpha", 0.);
	const Real q = num(a, "q", 0.), T = num(a, "T", 1.);
	const long steps = integer(a, "steps", 100);
	const int E = (int) integer(a, "E", 0);
	const int refine = (int) integer(a, "refine", 0);
	const bool american = integer(a, "american", 0) != 0;
	const bool call = integer(a, "call", 0) != 0;
	const std::string scheme = str(a, "scheme", "rannacher");
	const Real D = num(a, "dividend", 0.);
	const std::string dividendKind = str(a, "dividend_kind", "none");
	const bool dividendFirst = integer(a, "dividend_first", 0) != 0;  // applied first = add()ed first (forward: ascending id)
	const bool withExercise = integer(a, "exercise", 0) != 0;

	RectilinearGrid1 coarse( baseAxis(str(a, "axis", "special"), K, K) );
	auto grid = coarse.refined(refine);

	Function1 payoff = call ? callPayoff(K) : putPayoff(K);
	auto exercise = [K] (const Interpolant1 &V, Real S) { return max(V(S), K - S); };
	auto cashDividend = [D] (const Interpolant1 &V, Real S) { return V( max( S - D, 0. ) ); };
	auto proportionalDividend = [D] (const Interpolant1 &V, Real S) { return V( (1. - D) * S ); };
	auto localVol = [alpha] (Real S) { return alpha / S; };

	ForwardConstantStepper stepper(0., T, T / steps);
	auto addDividends = [&] () {
		for(int e = 1; e <= E; ++e) {
			if(dividendKind == "cash") stepper.add(T / E * e, cashDividend, grid);
			else if(dividendKind == "proportional") stepper.add(T / E * e, proportionalDividend, grid);
		}
	};
	if(dividendFirst) addDividends();
	if(withExercise) for(int e = 1; e <= E; ++e) stepper.add(T / E * e, exercise, grid);
	if(!dividendFirst) addDividends();

	std::unique_ptr<BlackScholes1> bs(alpha > 0. ? new BlackScholes1(grid, r, localVol, q) : new BlackScholes1(grid, r, vol, q));

	std::unique_ptr<IterationNode> discretization;
	if(scheme == "rannacher") discretization.reset(new ForwardRannacher(grid, *bs));
	else if(scheme == "cn") discretization.reset(new ForwardCrankNicolson(grid, *bs));
	else if(scheme == "bdf1") discretization.reset(new ForwardBDFOne(grid, *bs));
	else if(scheme == "bdf2") discretization.reset(new ForwardBDFTwo(grid, *bs));
	else if(scheme == "bdf3") discretization.reset(new ForwardBDFThree(grid, *bs));
	else if(scheme == "bdf4") discretization.reset(new ForwardBDFFour(grid, *bs));
	else if(scheme == "bdf5") discretization.reset(new ForwardBDFFive(grid, *bs));
	else if(scheme == "bdf6") discretization.reset(new ForwardBDFSix(grid, *bs));
	else { std::fprintf(stderr, "unknown scheme\n"); return 2; }
	discretization->setIteration(stepper);

	ToleranceIteration tolerance;
	std::unique_ptr<MinPenaltyMethodDifference1> penalty;
	IterationNode *root = discretization.get();
	if(american) {
		penalty.reset(new MinPenaltyMethodDifference1(grid, *discretization, payoff));
		penalty->setIteration(tolerance);
		stepper.setInnerIteration(tolerance);
		root = penalty.get();
	}

	SparseLUSolver solver;
	auto begin = std::chrono::steady_clock::now();
	auto solution = stepper.solve(grid, payoff, *root, solver);
	auto end = std::chrono::steady_clock::now();

	std::vector<long> inner, mask;
	if(american) {
		for(auto k : tolerance.iterations()) inner.push_back((long)k);
		for(bool b : penalty->constraintMask()) mask.push_back(b);
	}
	std::printf("value %a\n", (double) solution(K));
	std::printf("seconds %.9g\n", std::chrono::duration<double>(end - begin).count());
	std::printf("steps %ld\n", (long) stepper.iterations()[0]);
	dumpReals("S", nodes(grid));
	dumpReals("V", nodeValues(grid, solution));
	dumpInts("inner", inner);
	dumpInts("mask", mask);
	return 0;
}

////////////////////////////////////////////////////////////////////////////////
// glwb: examples/experimental/glwb.cpp run(k) — GLWBOperator (a BlackScholes1 whose b(t) is
// -(R[n+1] - R[n]) S, n = floor(t)), ReverseBDFTwo (or BDFOne), an event at every year
// n = 1 .. years (the last one AT the initial time of the reverse sweep) with the three-way choice,
// zero initial condition.  Mortality and penalty tables are synthetic and short:
// mortality[n] = .01 (n + 1) (the last year 1), penalty[n] = max(.03 - .01 n, 0).
////////////////////////////////////////////////////////////////////////////////

namespace glwbmode {
Real *R;
class GLWBOperator : public BlackScholes1 {
public:
	GLWBOperator(RectilinearGrid1 &grid, Real r, Real vol, Real management_rate) noexcept
			: BlackScholes1(grid, r, vol, management_rate) {}
	virtual Vector b(Real t) {
		Vector S_vec = G.image([] (Real S) { return S; });
		int n = floor(t);
		return -(R[n+1] - R[n]) * S_vec;
	}
};
}

int runGlwb(const Args &a) {
	const Real r = num(a, "r", .04), vol = num(a, "vol", .2), S_0 = num(a, "S0", 100.);
	const Real management_rate = num(a, "management_rate", .015), withdrawal_rate = num(a, "withdrawal_rate", .05);
	const Real bonus_rate = num(a, "bonus_rate", .06);
	const int T = (int) integer(a, "years", 6);
	const long steps = integer(a, "steps", 12 * T);
	const int refine = (int) integer(a, "refine", 0);
	const std::string scheme = str(a, "scheme", "bdf2");

	std::vector<Real> mortality((size_t) T), penalty((size_t) T), R((size_t) T + 2);
	for(int n = 0; n < T; ++n) { mortality[n] = n == T - 1 ? 1. : .01 * (n + 1); penalty[n] = max(.03 - .01 * n, 0.); }
	R[0] = 1.;
	for(int n = 1; n <= T; ++n) R[n] = R[n-1] * (1. - mortality[n-1]);
	R[T + 1] = R[T];            // the reference reads R[T+1] in its zero-length first step (multiplied by h = 0)
	glwbmode::R = R.data();

	RectilinearGrid1 coarse( S_0 * Axis::special );
	auto grid = coarse.refined(refine);

	glwbmode::GLWBOperator op(grid, r, vol, management_rate);
	ReverseConstantStepper stepper(0., (Real) T, (Real) T / steps);
	std::unique_ptr<IterationNode> discretization;
	if(scheme == "bdf2") discretization.reset(new ReverseBDFTwo(grid, op));
	else discretization.reset(new ReverseBDFOne(grid, op));
	discretization->setIteration(stepper);

	const Real Q = S_0;
	const Real *Rp = R.data();
	for(int n = 1; n <= T; ++n) {
		const Real pen = penalty[n-1];
		auto event = [=] (const Interpolant1 &V, Real S) {
			const Real GQ = withdrawal_rate * Q;
			Real best = (1. + bonus_rate) * V(S / (1. + bonus_rate));
			{
				const Real S_plus = max(S - GQ, 0.);
				const Real tmp = V(S_plus) + Rp[n] * GQ;
				if(tmp > best) best = tmp;
			}
			if(S > GQ) {
				const Real tmp = Rp[n] * ( GQ + ( 1. - pen ) * (S - GQ) );
				if(tmp > best) best = tmp;
			}
			return best;
		};
		stepper.add((Real) n, event, grid);
	}

	SparseLUSolver solver;
	auto begin = std::chrono::steady_clock::now();
	auto solution = stepper.solve(grid, [] (Real) { return 0.; }, *discretization, solver);
	auto end = std::chrono::steady_clock::now();

	std::printf("value %a\n", (double) solution(S_0));
	std::printf("seconds %.9g\n", std::chrono::duration<double>(end - begin).count());
	std::printf("steps %ld\n", (long) stepper.iterations()[0]);
	dumpReals("S", nodes(grid));
	dumpReals("V", nodeValues(grid, solution));
	dumpReals("R", R);
	return 0;
}

////////////////////////////////////////////////////////////////////////////////
// bermudan: local volatility alpha / S, E exercise dates at T e / E,
// BDF2 (tutorial.cpp) or Rannacher, no inner iteration
////////////////////////////////////////////////////////////////////////////////

int runBermudan(const Args &a) {
	const Real K = num(a, "K", 100.);
	const Real r = num(a, "r", .04), alpha = num(a, "alpha", 20.);
	const Real q = num(a, "q", 0.), T = num(a, "T", 1.);
	const long steps = integer(a, "steps", 100);
	const int E = (int) integer(a, "E", 10);
	const int refine = (int) integer(a, "refine", 0);
	const std::string scheme = str(a, "scheme", "bdf2");
	const int repeat = (int) integer(a, "repeat", 1);
	// README.md:357-375: discrete dividends at the event times
	const Real D = num(a, "dividend", 0.);
	const std::string dividendKind = str(a, "dividend_kind", "none");
	const bool dividendFirst = integer(a, "dividend_first", 0) != 0;  // first in the reverse sweep = add()ed last
	const bool withExercise = integer(a, "exercise", 1) != 0;

	RectilinearGrid1 coarse( baseAxis(str(a, "axis", "tutorial"), K, K) );
	auto grid = coarse.refined(refine);

	auto payoff = [K] (Real S) { return max(K - S, 0.); };
	auto exercise = [K] (const Interpolant1 &V, Real S) {
		return max(V(S), K - S);
	};
	auto cashDividend = [D] (const Interpolant1 &V, Real S) {
		return V( max( S - D, 0. ) );
	};
	auto proportionalDividend = [D] (const Interpolant1 &V, Real S) {
		return V( (1. - D) * S );
	};
	auto localVol = [alpha] (Real S) { return alpha / S; };

	double seconds = 0.;
	std::vector<double> values;
	long timesteps = 0;

	for(int rep = 0; rep < repeat; ++rep) {
		ReverseConstantStepper stepper(0., T, T / steps);
		auto addDividends = [&] () {
			for(int e = 0; e < E; ++e) {
				if(dividendKind == "cash") stepper.add(T / E * e, cashDividend, grid);
				else if(dividendKind == "proportional") stepper.add(T / E * e, proportionalDividend, grid);
			}
		};
		if(!dividendFirst) addDividends();
		if(withExercise) for(int e = 0; e < E; ++e) {
			stepper.add(T / E * e, exercise, grid);
		}
		if(dividendFirst) addDividends();

		BlackScholes1 bs(grid, r, localVol, q);

		std::unique_ptr<IterationNode> discretization;
		if(scheme == "bdf2") {
			discretization.reset(new ReverseBDFTwo(grid, bs));
		} else if(scheme == "bdf3") {
			discretization.reset(new ReverseBDFThree(grid, bs));
		} else if(scheme == "bdf4") {
			discretization.reset(new ReverseBDFFour(grid, bs));
		} else if(scheme == "bdf5") {
			discretization.reset(new ReverseBDFFive(grid, bs));
		} else if(scheme == "bdf6") {
			discretization.reset(new ReverseBDFSix(grid, bs));
		} else if(scheme == "bdf1") {
			discretization.reset(new ReverseBDFOne(grid, bs));
		} else if(scheme == "rannacher") {
			discretization.reset(new ReverseRannacher(grid, bs));
		} else {
			discretization.reset(new ReverseCrankNicolson(grid, bs));
		}
		discretization->setIteration(stepper);

		SparseLUSolver solver;

		auto begin = std::chrono::steady_clock::now();
		auto solution = stepper.solve(grid, payoff, *discretization, solver);
		auto end = std::chrono::steady_clock::now();
		seconds += std::chrono::duration<double>(end - begin).count();

		if(rep == 0) {
			values = nodeValues(grid, solution);
			timesteps = (long) stepper.iterations()[0];
			std::printf("value %a\n", (double) solution(K));
		}
	}

	std::printf("seconds %.9g\n", seconds / repeat);
	std::printf("steps %ld\n", timesteps);
	dumpReals("S", nodes(grid));
	dumpReals("V", values);
	return 0;
}

////////////////////////////////////////////////////////////////////////////////
// hjb: unequal borrowing / lending rates.  The interest rate is a control with
// values {r_l, r_b}; Min/MaxPolicyIteration1_1 inside a ToleranceIteration,
// BDF2 (or BDF1) time stepping, straddle payoff.
////////////////////////////////////////////////////////////////////////////////

int runHjb(const Args &a) {
	const Real K = num(a, "K", 100.), S0 = num(a, "S0", K);
	const Real r_l = num(a, "r_l", .03), r_b = num(a, "r_b", .05);
	const Real vol = num(a, "vol", .3), q = num(a, "q", 0.), T = num(a, "T", 1.);
	const long steps = integer(a, "steps", 12);
	const int refine = (int) integer(a, "refine", 0);
	const bool long_position = integer(a, "long", 0) != 0;
	const std::string scheme = str(a, "scheme", "bdf2");
	const int repeat = (int) integer(a, "repeat", 1);

	RectilinearGrid1 coarse( baseAxis(str(a, "axis", "special"), S0, K) );
	auto grid = coarse.refined(refine);
	RectilinearGrid1 controls( Axis { r_l, r_b } );
	auto payoff = straddlePayoff(K);

	double seconds = 0.;
	std::vector<double> values;
	std::vector<long> inner;
	long timesteps = 0;

	for(int rep = 0; rep < repeat; ++rep) {
		ReverseConstantStepper stepper(0., T, T / steps);
		ToleranceIteration tolerance;
		stepper.setInnerIteration(tolerance);

		BlackScholes1 bs(grid, Control1(grid), vol, q);

		std::unique_ptr<IterationNode> policy(long_position
			? (IterationNode *) new MaxPolicyIteration1_1(grid, controls, bs)
			: (IterationNode *) new MinPolicyIteration1_1(grid, controls, bs));
		policy->setIteration(tolerance);

		std::unique_ptr<IterationNode> discretization;
		if(scheme == "bdf2")      discretization.reset(new ReverseBDFTwo(grid, *policy));
		else if(scheme == "bdf3") discretization.reset(new ReverseBDFThree(grid, *policy));
		else if(scheme == "bdf4") discretization.reset(new ReverseBDFFour(grid, *policy));
		else if(scheme == "bdf5") discretization.reset(new ReverseBDFFive(grid, *policy));
		else if(scheme == "bdf6") discretization.reset(new ReverseBDFSix(grid, *policy));
		else if(scheme == "rannacher") discretization.reset(new ReverseRannacher(grid, *policy));
		else if(scheme == "cn")   discretization.reset(new ReverseCrankNicolson(grid, *policy));
		else                      discretization.reset(new ReverseBDFOne(grid, *policy));
		discretization->setIteration(stepper);

		SparseLUSolver solver;

		auto begin = std::chrono::steady_clock::now();
		auto solution = stepper.solve(grid, payoff, *discretization, solver);
		auto end = std::chrono::steady_clock::now();
		seconds += std::chrono::duration<double>(end - begin).count();

		if(rep == 0) {
			values = nodeValues(grid, solution);
			timesteps = (long) stepper.iterations()[0];
			for(auto k : tolerance.iterations()) inner.push_back((long)k);
			std::printf("value %a\n", (double) solution(S0));
		}
	}

	std::printf("seconds %.9g\n", seconds / repeat);
	std::printf("steps %ld\n", timesteps);
	dumpReals("S", nodes(grid));
	dumpReals("V", values);
	dumpInts("inner", inner);
	return 0;
}

////////////////////////////////////////////////////////////////////////////////
// jump: European call under Merton (lognormal) or Kou (double-exponential) jump
// diffusion, BDF1 with the explicit correlation-integral jump term
// (examples/jump_diffusion.cpp:43-140).  Also dumps what the operator computed
// at construction: the log-grid (N, x0, dx), the density weights and kappa.
////////////////////////////////////////////////////////////////////////////////

int runJump(const Args &a) {
	const Real K = num(a, "K", 100.), S0 = num(a, "S0", K);
	const Real r = num(a, "r", .05), vol = num(a, "vol", .15);
	const Real q = num(a, "q", 0.), T = num(a, "T", .25);
	const Real lambda = num(a, "lambda", .1);
	const long steps = integer(a, "steps", 12);
	const int refine = (int) integer(a, "refine", 0);
	const bool kou = str(a, "density", "lognormal") == "kou";
	const Real mu = num(a, "mu", -.1), sd = num(a, "sd", .45);
	const Real p_up = num(a, "p_up", .3445), eta1 = num(a, "eta1", 3.0465), eta2 = num(a, "eta2", 3.0775);
	const std::string scheme = str(a, "scheme", "bdf1");
	const int repeat = (int) integer(a, "repeat", 1);

	RectilinearGrid1 coarse( (S0 * Axis::special) + (K * Axis::special)
			+ (S0 * Axis { 1000. }) + (K * Axis { 1000. }) );   // jump_diffusion.cpp:166
	auto grid = coarse.refined(refine);
	auto payoff = callPayoff(K);

	double seconds = 0.;
	std::vector<double> values;
	long timesteps = 0;

	for(int rep = 0; rep < repeat; ++rep) {
		ReverseConstantStepper stepper(0., T, T / steps);
		auto density = kou ? doubleExponential(p_up, eta1, eta2) : lognormal(mu, sd);
		BlackScholesJumpDiffusion1 bs(grid, r, vol, q, lambda, density);
		bs.setIteration(stepper);

		std::unique_ptr<IterationNode> discretization;
		if(scheme == "bdf2") discretization.reset(new ReverseBDFTwo(grid, bs));
		else                 discretization.reset(new ReverseBDFOne(grid, bs));
		discretization->setIteration(stepper);

		SparseLUSolver solver;
		auto begin = std::chrono::steady_clock::now();
		auto solution = stepper.solve(grid, payoff, *discretization, solver);
		auto end = std::chrono::steady_clock::now();
		seconds += std::chrono::duration<double>(end - begin).count();

		if(rep == 0) {
			values = nodeValues(grid, solution);
			timesteps = (long) stepper.iterations()[0];
			std::printf("value %a\n", (double) solution(S0));
		}
	}

	std::printf("seconds %.9g\n", seconds / repeat);
	std::printf("steps %ld\n", timesteps);
	dumpReals("S", nodes(grid));
	dumpReals("V", values);
	return 0;
}

////////////////////////////////////////////////////////////////////////////////
// gmwb: the 2-D (W, A) guaranteed-minimum-withdrawal-benefit HJBQVI with the
// penalised scheme and BiCGSTAB, wired as examples/hjbqvi/gmwb.cpp:35-208.
// Fixture generator for the row of SURVEY.md §8(a) that the device path does
// not cover yet: dumps the solution, both control vectors and the iteration
// statistics of HJBQVI::solve(refinement).
////////////////////////////////////////////////////////////////////////////////

int runGmwb(const Args &a) {
	constexpr int Dimension = 2, SC = 1, IC = 1;
	const Real T = num(a, "T", 10.), r = num(a, "r", .05), eta = num(a, "eta", 0.);
	const Real sigma = num(a, "sigma", .2), G = num(a, "G", 10.), k = num(a, "k", .1);
	const Real c = num(a, "c", 0.), w_0 = num(a, "w0", 100.);
	const int timesteps = (int) integer(a, "timesteps", 32);
	const int a_points = (int) integer(a, "a_points", 51);
	const int control_points = (int) integer(a, "control_points", 3);
	const int refinement = (int) integer(a, "refine", 0);

	HJBQVI<Dimension, SC, IC> hjbqvi(timesteps,
		{ (w_0 / 100.) * Axis { 0., 5., 10., 15., 20., 25., 30., 35., 40., 45., 50., 55., 60., 65., 70.,
			72.5, 75., 77.5, 80., 82., 84., 86., 88., 90., 91., 92., 93., 94., 95., 96., 97., 98., 99.,
			100., 101., 102., 103., 104., 105., 106., 107., 108., 109., 110., 112., 114., 116., 118., 120.,
			123., 126., 130., 135., 140., 145., 150., 160., 175., 200., 225., 250., 300., 500., 750., 1000. },
		  Axis::uniform(0., w_0, a_points) },
		{ Axis { 0., G } }, { Axis::uniform(0., 1., control_points) }, T,
		[=] (Real, Real, Real) { return r; },
		{ [=] (Real, Real w, Real) { return sigma * w; }, [=] (Real, Real, Real) { return 0.; } },
		{ [=] (Real, Real w, Real av, Real q) { return (r - eta) * w + ((0. < av) ? -q : 0.); },
		  [=] (Real, Real, Real av, Real q) { return (0. < av) ? -q : 0.; } },
		[=] (Real, Real, Real av, Real q) { return (0. < av) ? q : 0.; },
		{ [=] (Real, Real w, Real av, Real z_frac) { const Real z = z_frac * av; return max(w - z, 0.); },
		  [=] (Real, Real, Real av, Real z_frac) { const Real z = z_frac * av; return av - z; } },
		[=] (Real, Real w, Real av, Real z_frac) {
			const Real z = z_frac * av;
			const Real w_plus = max(w - z, 0.), a_plus = av - z;
			if( (w - w_plus) + (av - a_plus) < QuantPDE::epsilon ) {
				return -std::numeric_limits<Real>::infinity();
			}
			return (1 - k) * z - c; },
		[=] (Real, Real w, Real av) { return max(w, (1 - k) * av - c); });
	hjbqvi.usePenalizedScheme();
	hjbqvi.useBiCGSTABSolver();
	hjbqvi.disableStochasticControlRefinement();
	hjbqvi.coefficientsAreTimeIndependent();
	hjbqvi.right_boundary(0, HJBQVILinearBoundary<Dimension, SC, IC>);
	hjbqvi.right_boundary(1, HJBQVIZeroDiffusionRightBoundary<Dimension, SC, IC>);

	auto result = hjbqvi.solve(refinement);
	PiecewiseLinear<Dimension> u(result.spatial_grid, result.solution_vector);
	std::printf("value %a\n", (double) u.interpolate({{ w_0, w_0 }}));
	std::printf("seconds %.9g\n", (double) result.execution_time_seconds);
	std::printf("steps %d\n", (int) result.timesteps);
	std::printf("mean_inner %a\n", (double) result.mean_inner_iterations);
	std::printf("mean_solver %a\n", (double) result.mean_solver_iterations);
	std::vector<double> W, A, V, Q, Z;
	for(Index i = 0; i < result.spatial_grid[0].size(); ++i) W.push_back(result.spatial_grid[0][i]);
	for(Index i = 0; i < result.spatial_grid[1].size(); ++i) A.push_back(result.spatial_grid[1][i]);
	for(Index i = 0; i < result.solution_vector.size(); ++i) {
		V.push_back(result.solution_vector(i));
		Q.push_back(result.stochastic_control_vector[0](i));
		Z.push_back(result.impulse_control_vector[0](i));
	}
	dumpReals("W", W); dumpReals("A", A); dumpReals("V", V); dumpReals("Q", Q); dumpReals("Z", Z);
	return 0;
}

////////////////////////////////////////////////////////////////////////////////
// exchange: the 1-D combined stochastic / impulse control of the exchange rate
// (Mundaca & Oksendal), penalised scheme + BiCGSTAB, wired as
// examples/hjbqvi/exchange_rate.cpp:29-175.  Dumps the coarse axes, the refined
// grids, the solution, both control vectors (NaN where the other control acts,
// HJBQVI.hpp:977-990) and the iteration statistics of HJBQVI::solve(refinement).
////////////////////////////////////////////////////////////////////////////////

int runExchange(const Args &a) {
	constexpr int Dimension = 1, SC = 1, IC = 1;
	const Real x_min = num(a, "x_min", -2.), x_max = num(a, "x_max", 2.), m = num(a, "m", 0.);
	const Real q_min = num(a, "q_min", -.07), q_max = num(a, "q_max", .07);
	const Real T = num(a, "T", 10.), rho = num(a, "rho", .02), sigma = num(a, "sigma", .3);
	const Real aa = num(a, "a", .25), b = num(a, "b", 3.), lambda = num(a, "lambda", 1.), c = num(a, "c", .1);
	const Real intensity = num(a, "intensity", 10.);
	const int points = (int) integer(a, "points", 32);
	const int q_points = (int) integer(a, "q_points", 9);
	const int z_points = (int) integer(a, "z_points", 16);
	const int timesteps = (int) integer(a, "timesteps", 16);
	const int refinement = (int) integer(a, "refine", 0);

	const Axis x = Axis::cluster(x_min, m, x_max, points, intensity);
	const Axis w = Axis::uniform(q_min, q_max, q_points);
	const Axis x_impulse = Axis::cluster(x_min, m, x_max, z_points, intensity);

	HJBQVI<Dimension, SC, IC> hjbqvi(timesteps, { x }, { w }, { x_impulse }, T,
		[=] (Real, Real) { return rho; },
		{ [=] (Real, Real) { return sigma; } },
		{ [=] (Real, Real, Real q) { return -aa * q; } },
		[=] (Real, Real xv, Real q) { const Real tmp = xv - m; return - tmp * tmp - q * q * b; },
		{ [=] (Real, Real, Real x_new) { return x_new; } },
		[=] (Real, Real xv, Real x_new) {
			if(std::abs(x_new - m) >= std::abs(xv - m)) return -std::numeric_limits<Real>::infinity();
			return - lambda * std::fabs(x_new - xv) - c; },
		[=] (Real, Real) { return 0.; });
	hjbqvi.usePenalizedScheme();
	hjbqvi.useBiCGSTABSolver();
	hjbqvi.coefficientsAreTimeIndependent();

	auto result = hjbqvi.solve(refinement);
	PiecewiseLinear<Dimension> u(result.spatial_grid, result.solution_vector);
	std::printf("value %a\n", (double) u.interpolate({{ m }}));
	std::printf("seconds %.9g\n", (double) result.execution_time_seconds);
	std::printf("steps %d\n", (int) result.timesteps);
	std::printf("mean_inner %a\n", (double) result.mean_inner_iterations);
	std::printf("mean_solver %a\n", (double) result.mean_solver_iterations);
	std::vector<double> X0, Q0, Z0, X, QA, ZA, V, Q, Z;
	for(Index i = 0; i < x.size(); ++i) X0.push_back(x[i]);
	for(Index i = 0; i < w.size(); ++i) Q0.push_back(w[i]);
	for(Index i = 0; i < x_impulse.size(); ++i) Z0.push_back(x_impulse[i]);
	for(Index i = 0; i < result.spatial_grid[0].size(); ++i) X.push_back(result.spatial_grid[0][i]);
	for(Index i = 0; i < result.stochastic_control_grid[0].size(); ++i) QA.push_back(result.stochastic_control_grid[0][i]);
	for(Index i = 0; i < result.impulse_control_grid[0].size(); ++i) ZA.push_back(result.impulse_control_grid[0][i]);
	for(Index i = 0; i < result.solution_vector.size(); ++i) {
		V.push_back(result.solution_vector(i));
		Q.push_back(result.stochastic_control_vector[0](i));
		Z.push_back(result.impulse_control_vector[0](i));
	}
	dumpReals("X0", X0); dumpReals("Q0", Q0); dumpReals("Z0", Z0);
	dumpReals("X", X); dumpReals("QA", QA); dumpReals("ZA", ZA);
	dumpReals("V", V); dumpReals("Q", Q); dumpReals("Z", Z);
	return 0;
}

////////////////////////////////////////////////////////////////////////////////
// consumption: optimal consumption with fixed and proportional transaction costs, a 2-D (w, b)
// HJBQVI with impulses pointing both ways in the state, penalised scheme + BiCGSTAB, wired as
// examples/hjbqvi/optimal_consumption.cpp:29-203.
////////////////////////////////////////////////////////////////////////////////

int runConsumption(const Args &a) {
	constexpr int Dimension = 2, SC = 1, IC = 1;
	const Real T = num(a, "T", 40.), rho = num(a, "rho", .1), r = num(a, "r", .07), mu = num(a, "mu", .11);
	const Real sigma = num(a, "sigma", .3), gamma = num(a, "gamma", .3), lambda = num(a, "lambda", .1), c = num(a, "c", .05);
	const Real q_max = num(a, "q_max", 100.), w_0 = num(a, "w0", 45.2), b_0 = num(a, "b0", 45.2);
	const Real boundary = num(a, "boundary", 200.);
	const int timesteps = (int) integer(a, "timesteps", 32);
	const int control_points = (int) integer(a, "control_points", 16);
	const int spatial_points = (int) integer(a, "spatial_points", 20);
	const int refinement = (int) integer(a, "refine", 0);
	const Axis axis = Axis::uniform(0., boundary, spatial_points);

	auto z_bounded = [=] (Real, Real w, Real b, Real z_frac) {
		const Real tmp1 = b - c;
		const Real tmp2 = tmp1 - boundary;
		const Real lo = max(-w, (tmp2 > 0) ? (tmp2 / (1 - lambda)) : (tmp2 / (1 + lambda)));
		const Real hi = min(boundary - w, (tmp1 > 0) ? (tmp1 / (1 + lambda)) : (tmp1 / (1 - lambda)));
		const Real z = z_frac * (hi - lo) + lo;
		if(std::fabs(z) < QuantPDE::epsilon) { return 0.; }
		return z;
	};

	HJBQVI<Dimension, SC, IC> hjbqvi(timesteps, { axis, axis },
		{ Axis::uniform(0., q_max, control_points) }, { Axis::uniform(0., 1., control_points) }, T,
		[=] (Real, Real, Real) { return rho; },
		{ [=] (Real, Real w, Real) { return sigma * w; }, [=] (Real, Real, Real) { return 0.; } },
		{ [=] (Real, Real w, Real, Real) { return mu * w; },
		  [=] (Real, Real, Real b, Real q) { return r * b + ((0 < b && b < boundary) ? -q : 0.); } },
		[=] (Real, Real, Real b, Real q) { return (0 < b && b < boundary) ? std::pow(q, gamma) / gamma : 0.; },
		{ [=] (Real t, Real w, Real b, Real z_frac) { const Real z = z_bounded(t, w, b, z_frac); return w + z; },
		  [=] (Real t, Real w, Real b, Real z_frac) { const Real z = z_bounded(t, w, b, z_frac); return b - z - lambda * std::fabs(z) - c; } },
		[=] (Real t, Real w, Real b, Real z_frac) {
			const Real z = z_bounded(t, w, b, z_frac);
			return z == 0. ? -std::numeric_limits<Real>::infinity() : 0.; },
		[=] (Real, Real w, Real b) { const Real liquidate = max(b + (1 - lambda) * w - c, 0.); return std::pow(liquidate, gamma) / gamma; });
	hjbqvi.usePenalizedScheme();
	hjbqvi.useBiCGSTABSolver();
	hjbqvi.coefficientsAreTimeIndependent();
	hjbqvi.ignoreExtrapolatoryControls();

	auto result = hjbqvi.solve(refinement);
	PiecewiseLinear<Dimension> u(result.spatial_grid, result.solution_vector);
	std::printf("value %a\n", (double) u.interpolate({{ w_0, b_0 }}));
	std::printf("seconds %.9g\n", (double) result.execution_time_seconds);
	std::printf("steps %d\n", (int) result.timesteps);
	std::printf("mean_inner %a\n", (double) result.mean_inner_iterations);
	std::printf("mean_solver %a\n", (double) result.mean_solver_iterations);
	std::vector<double> W, B, QA, ZA, V, Q, Z;
	for(Index i = 0; i < result.spatial_grid[0].size(); ++i) W.push_back(result.spatial_grid[0][i]);
	for(Index i = 0; i < result.spatial_grid[1].size(); ++i) B.push_back(result.spatial_grid[1][i]);
	for(Index i = 0; i < result.stochastic_control_grid[0].size(); ++i) QA.push_back(result.stochastic_control_grid[0][i]);
	for(Index i = 0; i < result.impulse_control_grid[0].size(); ++i) ZA.push_back(result.impulse_control_grid[0][i]);
	for(Index i = 0; i < result.solution_vector.size(); ++i) {
		V.push_back(result.solution_vector(i));
		Q.push_back(result.stochastic_control_vector[0](i));
		Z.push_back(result.impulse_control_vector[0](i));
	}
	dumpReals("W", W); dumpReals("B", B); dumpReals("QA", QA); dumpReals("ZA", ZA);
	dumpReals("V", V); dumpReals("Q", Q); dumpReals("Z", Z);
	return 0;
}

void usage() {
	std::fprintf(stderr,
		"qpde_ref vanilla  K= S0= r= vol= q= T= steps= refine= american=0|1 [E= dividend= dividend_kind= dividend_first= exercise=] "
		"call=0|1 scheme=rannacher|cn|bdf1|bdf2|...|bdf6 axis=special|tutorial repeat=\n"
		"qpde_ref bermudan K= r= alpha= q= T= steps= E= refine= "
		"scheme=bdf2|...|bdf6|bdf1|rannacher|cn axis= repeat=\n"
		"qpde_ref hjb      K= S0= r_l= r_b= vol= q= T= steps= refine= long=0|1 "
		"scheme=bdf2|bdf1 repeat=\n"
		"qpde_ref jump     K= S0= r= vol= q= T= lambda= steps= refine= density=lognormal|kou "
		"mu= sd= p_up= eta1= eta2= scheme=bdf1|bdf2 repeat=\n"
		"qpde_ref gmwb     T= r= eta= sigma= G= k= c= w0= timesteps= a_points= control_points= refine=\n"
		"qpde_ref consumption T= rho= r= mu= sigma= gamma= lambda= c= q_max= w0= b0= boundary= timesteps= control_points= spatial_points= refine=\n"
		"qpde_ref exchange x_min= x_max= m= q_min= q_max= T= rho= sigma= a= b= lambda= c= intensity= points= q_points= z_points= timesteps= refine=\n");
}

} // namespace

int main(int argc, char **argv) {
	if(argc < 2) { usage(); return 2; }
	Args a;
	for(int i = 2; i < argc; ++i) {
		std::string s(argv[i]);
		const size_t eq = s.find('=');
		if(eq == std::string::npos) { usage(); return 2; }
		a[s.substr(0, eq)] = s.substr(eq + 1);
	}
	const std::string mode(argv[1]);
	if(mode == "vanilla")  return runVanilla(a);
	if(mode == "bermudan") return runBermudan(a);
	if(mode == "forward") return runForward(a);
	if(mode == "glwb") return runGlwb(a);
	if(mode == "hjb")      return runHjb(a);
	if(mode == "jump")     return runJump(a);
	if(mode == "gmwb")     return runGmwb(a);
	if(mode == "exchange") return runExchange(a);
	if(mode == "consumption") return runConsumption(a);
	usage();
	return 2;
}
