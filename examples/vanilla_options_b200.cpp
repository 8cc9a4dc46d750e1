// vanilla_options_b200.cpp — the reference's examples/vanilla_options.cpp
// convergence table (European/American put or call, Rannacher time stepping,
// penalty method) computed on a B200 through QuantPDE_B200/Core.hpp.
//
// The wiring below mirrors examples/vanilla_options.cpp:34-130 of the
// reference; the table layout mirrors Modules/Utilities/Results.hpp:79-146
// (headers, Value / Change / Ratio / Timing, precision 6, column width 23).
// All refinement levels are put into ONE batch and swept in one device call.
//
//   vanilla_options_b200 [american=1] [call=0] [kn=5] [k0=0] [K=100] [S0=100]
//                        [r=.04] [vol=.2] [q=0] [T=1] [N=12] [variable=0] [device=0]
// variable=1: use_variable_timestepping of vanilla_options.cpp:52-58 (ReverseVariableStepper,
// target 1 / factor)
#include <QuantPDE_B200/Core.hpp>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <string>

using namespace QuantPDE_B200;
using namespace std;

static double arg(const map<string, string> &a, const char *k, double d) {
	auto it = a.find(k);
	return it == a.end() ? d : atof(it->second.c_str());
}

int main(int argc, char **argv) {
	map<string, string> a;
	for(int i = 1; i < argc; ++i) {
		const char *eq = strchr(argv[i], '=');
		if(!eq) { cerr << "expected key=value, got " << argv[i] << endl; return 2; }
		a[string(argv[i], eq - argv[i])] = string(eq + 1);
	}
	const bool american = arg(a, "american", 1) != 0, call = arg(a, "call", 0) != 0;
	const int kn = (int) arg(a, "kn", 5), k0 = (int) arg(a, "k0", 0);
	const Real K = arg(a, "K", 100.), S_0 = arg(a, "S0", 100.);
	const Real r = arg(a, "r", .04), vol = arg(a, "vol", .2), divs = arg(a, "q", 0.);
	const Real T = arg(a, "T", 1.);
	const int N = (int) arg(a, "N", 12);
	const bool variable = arg(a, "variable", 0) != 0;

	RectilinearGrid1 grid( (S_0 * Axis::special()) + (K * Axis::special()) );
	const Payoff payoff = call ? callPayoff(K) : putPayoff(K);

	struct Level {
		RectilinearGrid1 refined_grid;
		unique_ptr<BlackScholes1> bs;
		unique_ptr<ReverseConstantStepper> stepper;
		unique_ptr<ReverseRannacher> discretization;
		unique_ptr<ToleranceIteration> tolerance;
		unique_ptr<MinPenaltyMethodDifference1> penalty;
		IterationNode *root;
	};
	vector<unique_ptr<Level>> levels;

	try {
		B200Solver solver((int) arg(a, "device", 0));
		const int spacing = 23;
		cout << setw(spacing) << "Nodes" << setw(spacing) << "Steps"
		     << setw(spacing) << "Mean Policy Iterations" << setw(spacing) << "Value"
		     << setw(spacing) << "Change" << setw(spacing) << "Ratio"
		     << setw(spacing) << "Timing (Seconds)" << endl;
		cout.precision(6);
		Real previousValue = nan(""), previousChange = nan("");

		for(int k = k0; k <= kn; ++k) {
			// 2^k or 2^(2k) (vanilla_options.cpp:36-42)
			int factor = 1;
			for(int i = 0; i < k; ++i) factor *= american ? 4 : 2;
			const int N2k = N * factor;

			unique_ptr<Level> L(new Level());
			L->refined_grid = grid.refined(k);
			L->bs.reset(new BlackScholes1(L->refined_grid, r, vol, divs));
			if(variable) L->stepper.reset(new ReverseVariableStepper(0., T, T / N2k, 1. / factor));
			else L->stepper.reset(new ReverseConstantStepper(0., T, T / N2k));
			L->discretization.reset(new ReverseRannacher(L->refined_grid, *L->bs));
			L->discretization->setIteration(*L->stepper);
			L->root = L->discretization.get();
			if(american) {
				L->penalty.reset(new MinPenaltyMethodDifference1(L->refined_grid, *L->discretization, payoff));
				L->tolerance.reset(new ToleranceIteration());
				L->penalty->setIteration(*L->tolerance);
				L->stepper->setInnerIteration(*L->tolerance);
				L->root = L->penalty.get();
			}

			auto start = chrono::steady_clock::now();
			Solution solution = L->stepper->solve(L->refined_grid, payoff, *L->root, solver);
			const Real seconds = chrono::duration<Real>(chrono::steady_clock::now() - start).count();

			const Real value = solution(S_0);
			const Real change = value - previousValue, ratio = previousChange / change;
			cout << fixed << setw(spacing) << (Real) L->refined_grid.size()
			     << setw(spacing) << (Real) L->stepper->iterations()[0];
			const Real its = american ? L->tolerance->meanIterations() : nan("");
			Real intpart;
			if(abs(modf(its, &intpart)) < epsilon) cout << fixed; else cout << scientific;
			cout << setw(spacing) << its;
			cout << scientific << setw(spacing) << value << setw(spacing) << change
			     << setw(spacing) << ratio << setw(spacing) << seconds << endl;
			previousChange = change; previousValue = value;
			levels.push_back(std::move(L));
		}
	} catch(const exception &e) {
		cerr << e.what() << endl;
		return 1;
	}
	return 0;
}
