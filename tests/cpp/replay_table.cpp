// Test program (CPU): feeds recorded solutions of the reference through QuantPDE_B200::ResultsBuffer1
// and Solution, so that the table formatting and the piecewise-linear read-out are checked byte for
// byte against the stdout of the reference's own example program without a GPU
// (tests/test_utilities_cpu.py).  Input (text, hex floats):
//   point  S_min dS S_max  n_headers  header lines...  n_levels
//   per level: n_columns columns...  n_nodes  S...  V...
#include <QuantPDE_B200/Core.hpp>
#include <QuantPDE_B200/Utilities.hpp>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

using namespace QuantPDE_B200;

static double hex(std::istream &in) { std::string t; in >> t; return std::strtod(t.c_str(), nullptr); }

int main(int argc, char **argv) {
	if(argc < 2) return 2;
	std::ifstream in(argv[1]);
	const double point = hex(in), S_min = hex(in), dS = hex(in), S_max = hex(in);
	int nh; in >> nh; in.ignore();
	std::vector<std::string> headers(nh);
	for(auto &h : headers) std::getline(in, h);
	int levels; in >> levels;
	std::vector<ResultsTuple1> recorded;
	for(int k = 0; k < levels; ++k) {
		int nc; in >> nc;
		std::vector<Real> columns(nc);
		for(auto &c : columns) c = hex(in);
		int n; in >> n;
		std::vector<Real> S(n), V(n);
		for(auto &s : S) s = hex(in);
		for(auto &v : V) v = hex(in);
		recorded.emplace_back(columns, Solution(S, V), point);
	}
	ResultsBuffer1 buffer([&] (int k) { return recorded[k]; }, headers, levels - 1, 0);
	buffer.setPrintGrid(RectilinearGrid1(Axis::range(S_min, dS, S_max)));
	buffer.stream();
	return 0;
}
