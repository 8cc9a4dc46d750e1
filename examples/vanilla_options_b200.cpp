// vanilla_options_b200.cpp — the reference's examples/vanilla_options.cpp (convergence table for
// a European / American, digital / non-digital call / put; Rannacher time stepping, penalty
// method) computed on a B200 through QuantPDE_B200/Core.hpp.
//
// Same invocation as the reference program — `vanilla_options_b200 [CONF_FILE | -i]` with the
// JSON keys of examples/vanilla_options.cpp:139-156 — and the same stdout: run(k) below mirrors
// vanilla_options.cpp:34-130 line by line, the table is Modules/Utilities/Results.hpp:79-168
// (QuantPDE_B200/Utilities.hpp).  Extra: key=value arguments override single entries, and
// "device" selects the GPU.
#include <QuantPDE_B200/Core.hpp>
#include <QuantPDE_B200/Utilities.hpp>

#include <iostream>
#include <memory>
#include <numeric>

using namespace QuantPDE_B200;
using namespace std;

Real T, r, vol, divs, S_0, K;
int N;
bool call, digital, american, variable;
RectilinearGrid1 *grid;
B200Solver *solver;

ResultsTuple1 run(int k) {

	// 2^k or 2^(2k)
	int factor = 1;
	for(int i = 0; i < k; ++i) {
		if(american) { factor *= 4; }
		else         { factor *= 2; }
	}
	const int N2k = N * factor;

	Payoff payoff = digital ?
		  ( call ? digitalCallPayoff(K) : digitalPutPayoff(K) )
		: (call ? callPayoff(K) : putPayoff(K))
	;

	auto refined_grid = grid->refined(k);

	// Black-Scholes operator (L in V_t = LV)
	BlackScholes1 bs(
		refined_grid,
		r, vol, divs
	);

	// Timestepping method
	unique_ptr<ReverseConstantStepper> stepper(variable
		? (ReverseConstantStepper *) new ReverseVariableStepper(
			0.,         // startTime
			T,          // endTime
			T / N2k,    // dt
			1. / factor // target
		)
		: new ReverseConstantStepper(
			0.,     // startTime
			T,      // endTime
			T / N2k // dt
		)
	);

	// Time discretization method
	ReverseRannacher discretization(refined_grid, bs);
	discretization.setIteration(*stepper);

	// American-specific components; penalty method or not?
	IterationNode *root;
	unique_ptr<ToleranceIteration> tolerance;
	unique_ptr<MinPenaltyMethodDifference1> penalty;
	if(american) {
		// American case
		penalty = unique_ptr<MinPenaltyMethodDifference1>(
			new MinPenaltyMethodDifference1(
				refined_grid,
				discretization,
				payoff
			)
		);

		tolerance = unique_ptr<ToleranceIteration>(
			new ToleranceIteration()
		);

		penalty->setIteration(*tolerance);
		stepper->setInnerIteration(*tolerance);

		root = penalty.get();
	} else {
		// European case
		root = &discretization;
	}

	// Compute solution: the whole backward sweep is one device call
	auto solution = stepper->solve(
		refined_grid,
		payoff,
		*root,
		*solver
	);

	// Timesteps
	unsigned timesteps = stepper->iterations()[0];

	// Average number of policy iterations
	Real policy_its = nan("");
	if(american) {
		auto its = tolerance->iterations();
		policy_its = accumulate(its.begin(), its.end(), 0.)/its.size();
	}

	return ResultsTuple1(
		{(Real) refined_grid.size(), (Real) timesteps, policy_its },
		solution, S_0
	);

}

int main(int argc, char **argv) {
	try {
		// Parse configuration file
		Configuration configuration = getConfiguration(argc, argv);

		// Get options
		int kn, k0;
		Real S_max, S_min, dS;
		kn = getInt(configuration, "maximum_refinement", 5);
		k0 = getInt(configuration, "minimum_refinement", 0);
		T = getReal(configuration, "time_to_expiry", 1.);
		r = getReal(configuration, "interest_rate", .04);
		vol = getReal(configuration, "volatility", .2);
		divs = getReal(configuration, "dividend_rate", 0.);
		S_0 = getReal(configuration, "asset_price", 100.);
		K = getReal(configuration, "strike_price", 100.);
		S_min = getReal(configuration, "print_asset_price_minimum", 0.);
		S_max = getReal(configuration, "print_asset_price_maximum", S_0 * 2.);
		dS = getReal(configuration, "print_asset_price_step_size", S_0 / 10.);
		N = getInt(configuration, "initial_number_of_timesteps", 12);
		call = getBool(configuration, "is_call", true);
		digital = getBool(configuration, "is_digital", false);
		american = getBool(configuration, "is_american", false);
		variable = getBool(configuration, "use_variable_timestepping", false);
		RectilinearGrid1 default_grid( (S_0 * Axis::special()) + (K * Axis::special()) );
		RectilinearGrid1 tmp = getGrid(configuration, "initial_grid", default_grid);
		grid = &tmp;
		B200Solver device(getInt(configuration, "device", 0));
		solver = &device;

		// Print configuration file
		cerr << configuration << endl << endl;

		// Run and print results
		ResultsBuffer1 buffer(
			run,
			{ "Nodes", "Steps", "Mean Policy Iterations" },
			kn, k0
		);
		buffer.setPrintGrid( RectilinearGrid1(Axis::range(S_min, dS, S_max)) );
		buffer.stream();
	} catch(const exception &e) {
		cerr << e.what() << endl;
		return 1;
	}

	return 0;
}
