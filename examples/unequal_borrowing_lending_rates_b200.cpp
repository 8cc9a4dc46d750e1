// unequal_borrowing_lending_rates_b200.cpp — the reference's
// examples/unequal_borrowing_lending_rates.cpp convergence table (straddle under unequal
// borrowing / lending rates: an HJB equation with the interest rate as a control, policy
// iteration under BDF2) computed on a B200 through QuantPDE_B200/Core.hpp.
//
// Same invocation as the reference program — `unequal_borrowing_lending_rates_b200 [CONF_FILE | -i]`
// with the JSON keys of examples/unequal_borrowing_lending_rates.cpp:175-197 — and the same stdout:
// run(k) mirrors unequal_borrowing_lending_rates.cpp:40-171, the table is
// Modules/Utilities/Results.hpp:79-168.  Extra: key=value arguments override single entries,
// "device" selects the GPU.
#include <QuantPDE_B200/Core.hpp>
#include <QuantPDE_B200/Utilities.hpp>

#include <iostream>
#include <memory>
#include <numeric>

using namespace QuantPDE_B200;
using namespace std;

Real T, r_l, r_b, vol, divs, S_0, K;
int N;
bool long_position;
RectilinearGrid1 *grid;
B200Solver *solver;

ResultsTuple1 run(int k) {

	// 2^k
	int factor = 1;
	for(int i = 0; i < k; ++i) { factor *= 2; }

	auto refined_grid = grid->refined(k);

	// Control can be two interest rates: lending or borrowing
	RectilinearGrid1 controls( Axis { r_l, r_b } );

	auto payoff = straddlePayoff(K);

	ReverseConstantStepper stepper(
		0.,              // Initial time
		T,               // Expiry time
		T / (N * factor) // Timestep size
	);
	ToleranceIteration tolerance;
	stepper.setInnerIteration(tolerance);

	BlackScholes1 bs(
		refined_grid,

		// Interest rate (passed as a control)
		Control1(refined_grid),

		vol, // Volatility
		divs  // Dividend rate
	);

	// Pick min/max policy iteration depending on whether we are considering
	// the long or short position problem.
	unique_ptr<PolicyIteration1_1> policy(long_position
		? (PolicyIteration1_1 *)
			// sup[ V_tau - LV ]
			new MaxPolicyIteration1_1(refined_grid, controls, bs)
		: (PolicyIteration1_1 *)
			// inf[ V_tau - LV ]
			new MinPolicyIteration1_1(refined_grid, controls, bs)
	);
	policy->setIteration(tolerance); // Associate with k-iteration

	// Discretization method
	typedef ReverseBDFTwo Discretization;
	Discretization discretization(refined_grid, *policy);
	discretization.setIteration(stepper); // Associate with n-iteration

	// Calculate the solution at time zero
	auto solution = stepper.solve(
		refined_grid,   // Domain
		payoff,         // Initial condition
		discretization, // Root of linear system tree
		*solver         // The device
	);

	// Timesteps
	unsigned timesteps = stepper.iterations()[0];

	// Average number of policy iterations
	Real policy_its = nan("");
	auto its = tolerance.iterations();
	policy_its = accumulate(its.begin(), its.end(), 0.) / its.size();

	return ResultsTuple1(
		{(Real) refined_grid.size(), (Real) timesteps, policy_its},
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
		r_b = getReal(configuration, "interest_rate_long", .05);
		r_l = getReal(configuration, "interest_rate_short", .03);
		vol = getReal(configuration, "volatility", .3);
		divs = getReal(configuration, "dividend_rate", 0.);
		S_0 = getReal(configuration, "asset_price", 100.);
		K = getReal(configuration, "strike_price", 100.);
		S_min = getReal(configuration, "print_asset_price_minimum", 0.);
		S_max = getReal(configuration, "print_asset_price_maximum", S_0 * 2.);
		dS = getReal(configuration, "print_asset_price_step_size", S_0 / 10.);
		N = getInt(configuration, "initial_number_of_timesteps", 12);
		long_position = getBool(configuration, "long_position", false);
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
