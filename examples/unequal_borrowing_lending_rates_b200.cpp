// unequal_borrowing_lending_rates_b200.cpp — the reference's
// examples/unequal_borrowing_lending_rates.cpp convergence table (straddle under unequal
// borrowing / lending rates: an HJB equation with the interest rate as a control, policy
// iteration under BDF2) computed on a B200 through QuantPDE_B200/Core.hpp.
//
// The wiring mirrors examples/unequal_borrowing_lending_rates.cpp:40-176; the table layout
// mirrors Modules/Utilities/Results.hpp:79-146.
//
//   unequal_borrowing_lending_rates_b200 [long_position=0] [kn=5] [k0=0] [K=100] [S0=100] [r_b=.05]
//                                        [r_l=.03] [vol=.3] [q=0] [T=1] [N=12] [device=0]
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
	const int kn = (int) arg(a, "kn", 5), k0 = (int) arg(a, "k0", 0);
	const Real K = arg(a, "K", 100.), S_0 = arg(a, "S0", 100.);
	const Real r_b = arg(a, "r_b", .05), r_l = arg(a, "r_l", .03), vol = arg(a, "vol", .3), divs = arg(a, "q", 0.);
	const Real T = arg(a, "T", 1.);
	const int N = (int) arg(a, "N", 12);
	const bool long_position = arg(a, "long_position", 0) != 0;

	RectilinearGrid1 grid( (S_0 * Axis::special()) + (K * Axis::special()) );
	// Control can be two interest rates: lending or borrowing
	RectilinearGrid1 controls( Axis { r_l, r_b } );
	const Payoff payoff = straddlePayoff(K);

	try {
		B200Solver solver((int) arg(a, "device", 0));
		const int spacing = 23;
		cout << setw(spacing) << "Nodes" << setw(spacing) << "Steps" << setw(spacing) << "Mean Policy Iterations"
		     << setw(spacing) << "Value" << setw(spacing) << "Change" << setw(spacing) << "Ratio"
		     << setw(spacing) << "Timing (Seconds)" << endl;
		cout.precision(6);
		Real previousValue = nan(""), previousChange = nan("");
		for(int k = k0; k <= kn; ++k) {
			int factor = 1;
			for(int i = 0; i < k; ++i) factor *= 2;
			RectilinearGrid1 refined_grid = grid.refined(k);
			ReverseConstantStepper stepper(0., T, T / (N * factor));
			ToleranceIteration tolerance;
			stepper.setInnerIteration(tolerance);
			// Interest rate passed as a control
			BlackScholes1 bs(refined_grid, Control1(refined_grid), vol, divs);
			// sup[ V_tau - LV ] for the long position, inf[ V_tau - LV ] for the short one
			unique_ptr<PolicyIteration1_1> policy(long_position
				? (PolicyIteration1_1 *) new MaxPolicyIteration1_1(refined_grid, controls, bs)
				: (PolicyIteration1_1 *) new MinPolicyIteration1_1(refined_grid, controls, bs));
			policy->setIteration(tolerance);
			ReverseBDFTwo discretization(refined_grid, *policy);
			discretization.setIteration(stepper);

			auto start = chrono::steady_clock::now();
			Solution solution = stepper.solve(refined_grid, payoff, discretization, solver);
			const Real seconds = chrono::duration<Real>(chrono::steady_clock::now() - start).count();

			const Real value = solution(S_0);
			const Real change = value - previousValue, ratio = previousChange / change;
			cout << fixed << setw(spacing) << (Real) refined_grid.size() << setw(spacing) << (Real) stepper.iterations()[0]
			     << scientific << setw(spacing) << tolerance.meanIterations()
			     << setw(spacing) << value << setw(spacing) << change << setw(spacing) << ratio
			     << setw(spacing) << seconds << endl;
			previousChange = change; previousValue = value;
		}
	} catch(const exception &e) {
		cerr << e.what() << endl;
		return 1;
	}
	return 0;
}
