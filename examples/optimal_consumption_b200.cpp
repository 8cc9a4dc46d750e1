// optimal_consumption_b200.cpp — the reference's examples/hjbqvi/optimal_consumption.cpp (optimal
// consumption with fixed and proportional transaction costs: a 2-D impulse-control HJBQVI under the
// penalised scheme whose impulses point both ways in the state) computed on a B200 through
// QuantPDE_B200::OptimalConsumption.
//
// Prints what HJBQVI_main prints (Modules/HJBQVI/HJBQVI.hpp:1353-1476: precision 12, column width 23): the
// convergence table with the value at (w_0, b_0) = (45.2, 45.2), then the solution and both controls at
// every node of the finest level.  The reference program is fixed at max_refinement = 6 (1217 x 1217 nodes,
// 961 + 961 controls, 2048 steps), which its CPU path cannot finish in days (refinement 3 takes it 11 minutes,
// and a level costs ~13 times the one before); here the level is an argument.  Equal to the reference's
// output apart from "Mean Solver Iterations" (BiCGSTAB iterations there, line sweeps per policy iteration
// here: DESIGN.md 4.8) and the execution time.
//
//   optimal_consumption_b200 [max_refinement=4] [min_refinement=0] [device=0] [dump=1]
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
	const int max_refinement = (int) arg(a, "max_refinement", 4), min_refinement = (int) arg(a, "min_refinement", 0);
	const bool dump = arg(a, "dump", 1.) != 0.;
	OptimalConsumption problem;                    // parameters of optimal_consumption.cpp:36-80

	try {
		B200Solver solver((int) arg(a, "device", 0));
		ostream &out = cout;
		out.precision(12);
		const int spacing = 23;
		auto space = [=] () { return setw(spacing); };
		const char *headers[] = { "Spatial Nodes", "Stochastic Ctrl Nodes", "Impulse Ctrl Nodes", "Timesteps",
			"Scaling Factor", "Policy Tolerance", "Mean Policy Iterations", "Mean Solver Iterations", "Value",
			"Change", "Ratio", "Execution Time (sec)" };
		for(const char *h : headers) out << space() << h;
		out << endl;
		Real previousValue = nan(""), previousChange = nan("");
		for(int refinement = min_refinement; ; ++refinement) {
			auto start = chrono::steady_clock::now();
			const OptimalConsumption::Result result = problem.solve(solver, refinement);
			const Real seconds = chrono::duration<Real>(chrono::steady_clock::now() - start).count();
			const Real value = result(problem.w_0, problem.b_0);
			const Real change = value - previousValue, ratio = previousChange / change;
			previousValue = value; previousChange = change;
			const Real dt = problem.T / result.steps;
			out << space() << result.W.size() * result.B.size() << space() << result.stochasticControlNodes
			    << space() << result.impulseControlNodes << space() << result.steps
			    << space() << problem.scalingFactor * dt << space() << problem.iterationTolerance
			    << space() << result.meanPolicyIterations << space() << result.meanSolverIterations
			    << space() << value << space() << change << space() << ratio
			    << space() << seconds << endl;
			if(refinement == max_refinement) {
				if(dump) {
					out << endl;
					out << space() << "x_1" << space() << "x_2" << space() << "Value u(t=0, x)"
					    << space() << "Stochastic Control w_1" << space() << "Impulse Control z_1" << endl;
					const size_t nw = result.W.size();
					for(size_t k = 0; k < result.V.size(); ++k) {
						out << space() << result.W[k % nw] << space() << result.B[k / nw] << space() << result.V[k]
						    << space() << result.Q[k] << space() << result.Z[k] << endl;
					}
				}
				break;
			}
		}
	} catch(const exception &e) {
		cerr << e.what() << endl;
		return 1;
	}
	return 0;
}
