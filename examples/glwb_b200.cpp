// glwb_b200.cpp — the reference's examples/experimental/glwb.cpp (guaranteed lifelong withdrawal benefit:
// a Black-Scholes operator with a time-dependent source, BDF2, and at every policy year an event that picks
// the best of non-withdrawal with bonus / withdrawal / surrender) computed on a B200 through
// QuantPDE_B200/Core.hpp.
//
// Same invocation as the reference program — `glwb_b200 [CONF_FILE | -i]` with the JSON keys of
// glwb.cpp:140-206 — and the same stdout.  What differs from the reference's source, and why:
//   * GLWBOperator overrides b(t) with -(R[n+1] - R[n]) S, n = floor(t) (glwb.cpp:31-45); a virtual function
//     cannot cross the C ABI, a table of the scalar per year can: BlackScholes1::setSource;
//   * the event lambda (glwb.cpp:84-121) is written as an EventExpression with the per-year numbers
//     R[n] and the penalty rate as its parameters.
// The reference adds its last event AT the initial time of the reverse sweep and therefore takes a
// zero-length first step; so does the device.
#include <QuantPDE_B200/Core.hpp>
#include <QuantPDE_B200/Utilities.hpp>

#include <iostream>
#include <limits>
#include <vector>

using namespace QuantPDE_B200;
using namespace std;

std::vector<Real> R;

Real r, vol, S_0, management_rate, withdrawal_rate, bonus_rate;
int N, T, kn, k0;
RectilinearGrid1 *grid;
std::vector<Real> mortality, penalty_rates;
B200Solver *solver;

ResultsTuple1 run(int k) {

	// Refinement
	int factor = 1;
	for(int i = 0; i < k; ++i) {
		factor *= 2;
	}
	const int N2k = N * factor;
	auto refined_grid = grid->refined(k);

	typedef ReverseBDFTwo Discretization;

	////////////////////////////////////////////////////////////////////////

	// GLWBOperator: b(t) = -(R[n+1] - R[n]) S with n = floor(t)
	BlackScholes1 op(refined_grid, r, vol, management_rate);
	{
		std::vector<Real> source((size_t) T + 1, 0.);
		for(int n = 0; n < T; ++n) source[n] = -(R[n+1] - R[n]);
		op.setSource(source);
	}

	ReverseConstantStepper stepper(0., (Real) T, (Real) T / N2k);

	Discretization discretization(refined_grid, op);
	discretization.setIteration(stepper);

	////////////////////////////////////////////////////////////////////////

	const Real Q = S_0;
	{
		const Real GQ = withdrawal_rate * Q;
		EventExpression S = EventExpression::S(), Rn = EventExpression::param(0), pen = EventExpression::param(1);

		// Nonwithdrawal
		EventExpression best = c(1. + bonus_rate) * V(S / c(1. + bonus_rate));

		// Withdraw
		best = max(best, V(max(S - c(GQ), c(0.))) + Rn * c(GQ));

		// Surrender
		best = max(best, select(S > c(GQ),
			Rn * ( c(GQ) + ( c(1.) - pen ) * (S - c(GQ)) ),
			c(-numeric_limits<Real>::infinity())
		));

		std::vector<Real> times;
		std::vector<std::vector<Real>> table;
		for(int n = 1; n <= T; ++n) {
			times.push_back((Real) n);
			table.push_back({ R[n], penalty_rates[n-1] });
		}
		stepper.addEvents(times, best, table);
	}

	////////////////////////////////////////////////////////////////////////

	auto solution = stepper.solve(
		refined_grid,
		[] (Real) { return 0.; },
		discretization,
		*solver
	);

	return ResultsTuple1(
		{(Real) refined_grid.size(), (Real) N2k },
		solution, S_0
	);

}

int main(int argc, char **argv) {
	try {
		// Parse configuration file
		Configuration configuration = getConfiguration(argc, argv);

		// Get options
		Real S_max, S_min, dS;
		kn = getInt(configuration, "maximum_refinement", 8);
		k0 = getInt(configuration, "minimum_refinement", 0);
		r = getReal(configuration, "interest_rate", .04);
		vol = getReal(configuration, "volatility", .2);
		S_0 = getReal(configuration, "asset_price", 100.);
		management_rate = getReal(configuration, "management_rate", .015);
		withdrawal_rate = getReal(configuration, "withdrawal_rate", .05);
		bonus_rate = getReal(configuration, "bonus_rate", .06);
		S_min = getReal(configuration, "print_asset_price_minimum", 0.);
		S_max = getReal(configuration, "print_asset_price_maximum", S_0 * 2.);
		dS = getReal(configuration, "print_asset_price_step_size", S_0 / 10.);
		N = getInt(configuration, "initial_number_of_timesteps", 12);

		// PDE grid
		RectilinearGrid1 default_grid(S_0 * Axis::special());
		RectilinearGrid1 tmp1 = getGrid(configuration, "initial_grid", default_grid);
		grid = &tmp1;

		// Mortality data (default is DAV2004R)
		RectilinearGrid1 default_mortality( Axis {
			0.008886, 0.009938, 0.011253, 0.012687, 0.014231,
			0.015887, 0.017663, 0.019598, 0.021698, 0.023990,
			0.026610, 0.029533, 0.032873, 0.036696, 0.041106,
			0.046239, 0.052094, 0.058742, 0.066209, 0.074583,
			0.083899, 0.094103, 0.105171, 0.116929, 0.129206,
			0.141850, 0.154860, 0.168157, 0.181737, 0.195567,
			0.209614, 0.223854, 0.238280, 0.252858, 0.267526,
			0.278816, 0.293701, 0.308850, 0.324261, 0.339936,
			0.355873, 0.372069, 0.388523, 0.405229, 0.422180,
			0.439368, 0.456782, 0.474411, 0.492237, 0.510241,
			0.528401, 0.546689, 0.565074, 0.583517, 0.601976,
			0.620400, 1.000000
		} );
		mortality = getGrid(configuration, "mortality_data", default_mortality).nodes();
		T = (int) mortality.size();

		// Determine survival rate
		R.assign((size_t) T + 1, 0.); // R[0], ..., R[T]
		R[0] = 1.;
		for(int n = 1; n <= T; ++n) {
			R[n] = R[n-1] * (1. - mortality[n-1]);
		}

		// Default penalty rates
		std::vector<Real> default_penalty((size_t) 57, 0.);
		default_penalty[0] = 0.03; default_penalty[1] = 0.02; default_penalty[2] = 0.01;
		penalty_rates = getGrid(configuration, "penalty_data", RectilinearGrid1(Axis(default_penalty))).nodes();

		B200Solver device(getInt(configuration, "device", 0));
		solver = &device;

		// Print configuration file
		cerr << configuration << endl << endl;

		// Run and print results
		ResultsBuffer1 buffer(
			run,
			{ "Nodes", "Steps" },
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
