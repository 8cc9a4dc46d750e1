// tutorial_b200.cpp — the reference's examples/tutorial.cpp (Bermudan put under a
// local volatility alpha / S, ten exercise dates, BDF2) computed on a B200 through
// QuantPDE_B200/Core.hpp, with the README's optional discrete dividends
// (README.md:357-375, the block that tutorial.cpp:82-99 keeps commented out).
//
// The wiring mirrors examples/tutorial.cpp:30-141 line by line; what differs is
// that the event lambdas are tagged forms (exerciseValue(K), Dividend::cash(D)):
// an arbitrary C++ lambda cannot cross the C ABI.
//
//   tutorial_b200 [K=100] [r=.04] [alpha=20] [T=1] [dt=.01] [E=10] [D=0] [refine=0]
//                 [dividends_after_exercise=0] [device=0]
//                 [print_min=0] [print_step=10] [print_max=200] [precision=6]
// Without arguments the output is byte for byte that of the reference program.
#include <QuantPDE_B200/Core.hpp>

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
	const Real K = arg(a, "K", 100.), r = arg(a, "r", .04), alpha = arg(a, "alpha", 20.);
	const Real T = arg(a, "T", 1.), dt = arg(a, "dt", .01), D = arg(a, "D", 0.);
	const int E = (int) arg(a, "E", 10), refine = (int) arg(a, "refine", 0);
	const bool dividendsAfterExercise = arg(a, "dividends_after_exercise", 0) != 0;

	try {
		// Hand-picked grid (tutorial.cpp:41-57); K / 100 scales it with the strike
		RectilinearGrid1 grid( ((K / 100.) * Axis {
			0., 10., 20., 30., 40., 50., 60., 70., 75., 80., 84., 88., 92., 94., 96., 98., 100.,
			102., 104., 106., 108., 110., 114., 118., 123., 130., 140., 150., 175., 225., 300.,
			750., 2000., 10000.
		}) );
		RectilinearGrid1 refined = grid.refined(refine);

		const Payoff payoff = putPayoff(K);                       // max( K - S, 0. )
		ReverseConstantStepper stepper(0., T, dt);

		// "If we are modelling [dividends paid before the holder's right to exercise], the
		// above code should be placed before the for-loop responsible for adding exercise
		// events" (README.md:375)
		if(D > 0. && !dividendsAfterExercise) stepper.addDividendDates(E, Dividend::cash(D));
		stepper.addExerciseDates(E, exerciseValue(K));            // max( V(S), K - S )
		if(D > 0. && dividendsAfterExercise) stepper.addDividendDates(E, Dividend::cash(D));

		BlackScholes1 bs(refined, r, inverseSVolatility(alpha), 0.);   // v(S) = alpha / S
		ReverseBDFTwo bdf2(refined, bs);
		bdf2.setIteration(stepper);

		B200Solver solver((int) arg(a, "device", 0));
		Solution V = stepper.solve(refined, payoff, bdf2, solver);

		// Print the solution as the reference does (tutorial.cpp:150-154):
		//   RectilinearGrid1 printGrid( Axis::range(0., 10., 200.) );  cout << accessor( printGrid, V );
		// i.e. two columns of width 23 in the stream's default notation (Core/Domain.hpp:417-431)
		RectilinearGrid1 printGrid( Axis::range(arg(a, "print_min", 0.), arg(a, "print_step", 10.), arg(a, "print_max", 200.)) );
		cout.precision((int) arg(a, "precision", 6));
		for(Real S : printGrid.nodes()) cout << setw(23) << S << setw(23) << V(S) << endl;
	} catch(const exception &e) {
		cerr << e.what() << endl;
		return 1;
	}
	return 0;
}
