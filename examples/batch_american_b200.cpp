// batch_american_b200.cpp — many American puts in ONE device sweep through the header-only host
// API (QuantPDE_B200::Batch): what the reference does as a host loop over run(k)
// (Modules/Utilities/Results.hpp:101-105), as BASELINE configs[1] shapes it: varied strike /
// volatility / expiry / rate, grid K * Axis::special refined `refine` times, Rannacher + penalty.
//
//   batch_american_b200 [contracts=4096] [refine=6] [steps=1000] [seed=1234] [device=0]
// prints: contracts, nodes, seconds, contracts per second, mean penalty iterations per step, and
// the first three contracts (K, vol, T, r) with their values V(S_0 = K).
#include <QuantPDE_B200/Core.hpp>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <vector>

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
	const int C = (int) arg(a, "contracts", 4096), refine = (int) arg(a, "refine", 6), steps = (int) arg(a, "steps", 1000);
	mt19937_64 rng((unsigned long long) arg(a, "seed", 1234));
	uniform_real_distribution<double> uK(50., 150.), uVol(.10, .60), uT(.25, 2.), uR(.01, .08);

	struct Problem {
		RectilinearGrid1 grid; Real K, vol, T, r;
		unique_ptr<BlackScholes1> bs;
		unique_ptr<ReverseConstantStepper> stepper;
		unique_ptr<ReverseRannacher> discretization;
		unique_ptr<ToleranceIteration> tolerance;
		unique_ptr<MinPenaltyMethodDifference1> penalty;
	};
	vector<unique_ptr<Problem>> problems;
	try {
		B200Solver solver((int) arg(a, "device", 0));
		Batch batch;
		for(int c = 0; c < C; ++c) {
			const Real K = uK(rng), vol = uVol(rng), T = uT(rng), r = uR(rng);
			unique_ptr<Problem> p(new Problem());
			p->K = K; p->vol = vol; p->T = T; p->r = r;
			p->grid = RectilinearGrid1(K * Axis::special()).refined(refine);
			p->bs.reset(new BlackScholes1(p->grid, r, vol, 0.));
			p->stepper.reset(new ReverseConstantStepper(0., T, T / steps));
			p->discretization.reset(new ReverseRannacher(p->grid, *p->bs));
			p->discretization->setIteration(*p->stepper);
			p->penalty.reset(new MinPenaltyMethodDifference1(p->grid, *p->discretization, putPayoff(K)));
			p->tolerance.reset(new ToleranceIteration());
			p->penalty->setIteration(*p->tolerance);
			p->stepper->setInnerIteration(*p->tolerance);
			batch.add(*p->stepper, p->grid, putPayoff(K), *p->penalty);
			problems.push_back(std::move(p));
		}
		auto start = chrono::steady_clock::now();
		vector<Solution> V = batch.solve(solver);            // one qpde_bs1d_sweep
		const Real seconds = chrono::duration<Real>(chrono::steady_clock::now() - start).count();
		Real its = 0.;
		for(auto &p : problems) its += p->tolerance->meanIterations();
		cout.precision(17);
		cout << "contracts " << C << " nodes " << problems[0]->grid.size() << " seconds " << seconds
		     << " contracts_per_second " << C / seconds << " mean_penalty_iterations " << its / C << endl;
		for(int c = 0; c < 3 && c < C; ++c) {
			const Problem &p = *problems[c];
			cout << "value " << c << " K " << p.K << " vol " << p.vol << " T " << p.T << " r " << p.r << " V " << V[c](p.K) << endl;
		}
	} catch(const exception &e) {
		cerr << e.what() << endl;
		return 1;
	}
	return 0;
}
