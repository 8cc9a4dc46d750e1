// gmwb_b200.cpp — the convergence table of the reference's examples/hjbqvi/gmwb.cpp
// (guaranteed minimum withdrawal benefit: a 2-D impulse-control HJBQVI, penalised scheme)
// computed on a B200 through QuantPDE_B200::GMWB.
//
// The table layout mirrors HJBQVI_main (Modules/HJBQVI/HJBQVI.hpp:1353-1420: precision 12,
// column width 23, value at the test point (w_0, w_0), Change / Ratio between refinement
// levels).  "Mean Solver Iterations" is 1: the block-lower-triangular system of one policy
// iteration is solved directly by an ascending pass of line solves, where the
// reference iterates BiCGSTAB + ILUT.
//
//   gmwb_b200 [max_refinement=3] [min_refinement=0] [eta=0] [c=0] [a_points=51]
//             [control_points=3] [timesteps=32] [device=0]
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
	const int max_refinement = (int) arg(a, "max_refinement", 3), min_refinement = (int) arg(a, "min_refinement", 0);
	GMWB gmwb;                                    // parameters of gmwb.cpp:41-59
	gmwb.eta = arg(a, "eta", gmwb.eta);
	gmwb.c = arg(a, "c", gmwb.c);
	gmwb.timesteps = (int) arg(a, "timesteps", gmwb.timesteps);
	gmwb.controlPoints = (int) arg(a, "control_points", gmwb.controlPoints);
	gmwb.aAxis = Axis::uniform(0., gmwb.w_0, (Index) arg(a, "a_points", 51));

	try {
		B200Solver solver((int) arg(a, "device", 0));
		cout.precision(12);
		const int spacing = 23;
		const char *headers[] = { "Spatial Nodes", "Stochastic Ctrl Nodes", "Impulse Ctrl Nodes", "Timesteps",
			"Scaling Factor", "Policy Tolerance", "Mean Policy Iterations", "Mean Solver Iterations", "Value",
			"Change", "Ratio", "Execution Time (sec)" };
		for(const char *h : headers) cout << setw(spacing) << h;
		cout << endl;
		Real previousValue = nan(""), previousChange = nan("");
		for(int refinement = min_refinement; refinement <= max_refinement; ++refinement) {
			auto start = chrono::steady_clock::now();
			const GMWB::Result result = gmwb.solve(solver, refinement);
			const Real seconds = chrono::duration<Real>(chrono::steady_clock::now() - start).count();
			const Real value = result(gmwb.w_0, gmwb.w_0);
			const Real change = value - previousValue, ratio = previousChange / change;
			previousValue = value; previousChange = change;
			int impulse_nodes = gmwb.controlPoints - 1;
			for(int k = 0; k < refinement; ++k) impulse_nodes *= 2;
			cout << setw(spacing) << result.W.size() * result.A.size() << setw(spacing) << 2
			     << setw(spacing) << impulse_nodes + 1 << setw(spacing) << result.steps
			     << setw(spacing) << gmwb.scalingFactor << setw(spacing) << gmwb.iterationTolerance
			     << setw(spacing) << result.meanPolicyIterations << setw(spacing) << 1
			     << setw(spacing) << value << setw(spacing) << change << setw(spacing) << ratio
			     << setw(spacing) << seconds << endl;
		}
	} catch(const exception &e) {
		cerr << e.what() << endl;
		return 1;
	}
	return 0;
}
