// exchange_rate_b200.cpp — the reference's examples/hjbqvi/exchange_rate.cpp (optimal combined control
// of the exchange rate, Mundaca & Oksendal; a 1-D impulse-control HJBQVI under the penalised scheme)
// computed on a B200 through QuantPDE_B200::ExchangeRate.
//
// Like the reference program it takes no configuration by default and prints what HJBQVI_main prints
// (Modules/HJBQVI/HJBQVI.hpp:1353-1476: precision 12, column width 23): the convergence table over the
// refinement levels 0 .. 6 with the value at the parity m, then the solution and both controls at every
// node of the finest level.  The output equals the reference's byte for byte apart from two columns:
// "Mean Solver Iterations" (the reference counts BiCGSTAB iterations; here: tridiagonal solves per policy
// iteration) and the execution time.
//
//   exchange_rate_b200 [max_refinement=6] [min_refinement=0] [device=0] [sigma=..] [lambda=..] [c=..] [a=..] [b=..]
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
	const int max_refinement = (int) arg(a, "max_refinement", 6), min_refinement = (int) arg(a, "min_refinement", 0);
	ExchangeRate problem;                          // parameters of exchange_rate.cpp:36-76
	problem.sigma = arg(a, "sigma", problem.sigma);
	problem.lambda = arg(a, "lambda", problem.lambda);
	problem.c = arg(a, "c", problem.c);
	problem.a = arg(a, "a", problem.a);
	problem.b = arg(a, "b", problem.b);

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
			const ExchangeRate::Result result = problem.solve(solver, refinement);
			const Real seconds = chrono::duration<Real>(chrono::steady_clock::now() - start).count();
			const Real value = result(problem.m);            // convergence test at the optimal parity
			const Real change = value - previousValue, ratio = previousChange / change;
			previousValue = value; previousChange = change;
			const Real dt = problem.T / result.steps;
			out << space() << result.X.size() << space() << result.stochasticControlNodes
			    << space() << result.impulseControlNodes << space() << result.steps
			    << space() << problem.scalingFactor * dt << space() << problem.iterationTolerance
			    << space() << result.meanPolicyIterations << space() << result.meanSolverIterations
			    << space() << value << space() << change << space() << ratio
			    << space() << seconds << endl;
			if(refinement == max_refinement) {
				out << endl;
				out << space() << "x_1" << space() << "Value u(t=0, x)" << space() << "Stochastic Control w_1"
				    << space() << "Impulse Control z_1" << endl;
				for(size_t k = 0; k < result.X.size(); ++k) {
					out << space() << result.X[k] << space() << result.V[k] << space() << result.Q[k]
					    << space() << result.Z[k] << endl;
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
