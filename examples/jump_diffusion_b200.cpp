// jump_diffusion_b200.cpp — the reference's examples/jump_diffusion.cpp convergence table
// (European call under Merton / Kou jump diffusion, BDF1, explicit correlation-integral jump
// term) computed on a B200 through QuantPDE_B200/Core.hpp.
//
// The wiring mirrors examples/jump_diffusion.cpp:34-126; the table layout mirrors
// Modules/Utilities/Results.hpp:79-146.  With jump_type=double_exponential and the default
// parameters the rows reproduce the table recorded in examples/jump_diffusion.cpp:179-190
// (3.970527 at 265 nodes, 3.971451 at 529, ...).
//
//   jump_diffusion_b200 [jump_type=lognormal|double_exponential] [kn=6] [k0=3] [K=100] [S0=100]
//                       [r=.05] [vol=.15] [q=0] [T=.25] [N=12] [lambda=.1] [device=0]
#include <QuantPDE_B200/Core.hpp>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <map>
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
	const int kn = (int) arg(a, "kn", 6), k0 = (int) arg(a, "k0", 3);
	const Real K = arg(a, "K", 100.), S_0 = arg(a, "S0", 100.);
	const Real r = arg(a, "r", .05), vol = arg(a, "vol", .15), divs = arg(a, "q", 0.);
	const Real T = arg(a, "T", .25), jump_arrival_rate = arg(a, "lambda", .1);
	const int N = (int) arg(a, "N", 12);
	const string jump_type = a.count("jump_type") ? a["jump_type"] : "lognormal";
	if(jump_type != "lognormal" && jump_type != "double_exponential") {
		cerr << "error: jump_type can either be \"lognormal\" or \"double_exponential\"" << endl;
		return 2;
	}
	const JumpDensity jump_density = jump_type == "lognormal"
		? lognormal(arg(a, "jump_amplitude_mean", -.1), arg(a, "jump_amplitude_deviation", .45))
		: doubleExponential(arg(a, "jump_up_probability", .3445), arg(a, "jump_up_mean_reciprocal", 3.0465),
			arg(a, "jump_down_mean_reciprocal", 3.0775));

	// examples/jump_diffusion.cpp:166
	RectilinearGrid1 grid( (S_0 * Axis::special()) + (K * Axis::special()) + (S_0 * Axis { 1000. }) + (K * Axis { 1000. }) );
	const Payoff payoff = callPayoff(K);

	try {
		B200Solver solver((int) arg(a, "device", 0));
		const int spacing = 23;
		cout << setw(spacing) << "Nodes" << setw(spacing) << "Steps" << setw(spacing) << "Value"
		     << setw(spacing) << "Change" << setw(spacing) << "Ratio" << setw(spacing) << "Timing (Seconds)" << endl;
		cout.precision(6);
		Real previousValue = nan(""), previousChange = nan("");
		for(int k = k0; k <= kn; ++k) {
			int factor = 1;
			for(int i = 0; i < k; ++i) factor *= 2;
			RectilinearGrid1 refined_grid = grid.refined(k);
			ReverseConstantStepper stepper(0., T, T / (N * factor));
			BlackScholesJumpDiffusion1 bs(refined_grid, r, vol, divs, jump_arrival_rate, jump_density);
			bs.setIteration(stepper);
			ReverseBDFOne discretization(refined_grid, bs);
			discretization.setIteration(stepper);

			auto start = chrono::steady_clock::now();
			Solution solution = stepper.solve(refined_grid, payoff, discretization, solver);
			const Real seconds = chrono::duration<Real>(chrono::steady_clock::now() - start).count();

			const Real value = solution(S_0);
			const Real change = value - previousValue, ratio = previousChange / change;
			cout << fixed << setw(spacing) << (Real) refined_grid.size() << setw(spacing) << (Real) stepper.iterations()[0]
			     << scientific << setw(spacing) << value << setw(spacing) << change << setw(spacing) << ratio
			     << setw(spacing) << seconds << endl;
			previousChange = change; previousValue = value;
		}
	} catch(const exception &e) {
		cerr << e.what() << endl;
		return 1;
	}
	return 0;
}
