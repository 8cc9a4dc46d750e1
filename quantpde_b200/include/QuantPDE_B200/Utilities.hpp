// QuantPDE_B200/Utilities.hpp — the two utilities every example of parsiad/QuantPDE is driven by,
// so that an example ported to QuantPDE_B200/Core.hpp takes the same invocation and prints the
// same table as the reference program:
//
//   reference (QuantPDE/src/Modules/Utilities/...)          here (namespace QuantPDE_B200)
//   Configuration.hpp:15      typedef Json::Value           Configuration (a small JSON document tree; no JsonCpp)
//   Configuration.hpp:17-34   getInt/getReal/getBool/...    getInt, getReal, getBool, getString
//   Configuration.hpp:36-76   getConfiguration(argc, argv)  getConfiguration: [CONF_FILE | -i], -h
//   Configuration.hpp:78-134  getGrid                       getGrid (1-D)
//   Results.hpp:13-17         ResultsTuple1                 ResultsTuple1 = (columns, Solution, point)
//   Results.hpp:22-171        ResultsBuffer1                ResultsBuffer1: headers, Value / Change / Ratio /
//                                                           Timing (Seconds), precision 6, width 23, fixed
//                                                           notation for integral columns, the solution on a
//                                                           print grid after the last level
//
// Byte-for-byte the reference's stdout apart from the timing column
// (tests/test_tables_gpu.py compares against the recorded output of the reference's own example
// programs).  On top of the reference's invocation, arguments of the form key=value override
// single configuration entries (vanilla_options_b200 conf.json maximum_refinement=3).
#ifndef QUANT_PDE_B200_UTILITIES_HPP
#define QUANT_PDE_B200_UTILITIES_HPP

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <map>
#include <sstream>
#include <tuple>

#include "Core.hpp"

namespace QuantPDE_B200 {

////////////////////////////////////////////////////////////////////////////////
// Configuration
////////////////////////////////////////////////////////////////////////////////

class Configuration {
public:
	enum Type { NONE, INTEGER, REAL, BOOLEAN, TEXT, LIST, DICT };

	Configuration() : type_(NONE), integer_(0), real_(0.), flag_(false) {}
	static Configuration ofInt(long long v) { Configuration c; c.type_ = INTEGER; c.integer_ = v; return c; }
	static Configuration ofReal(Real v) { Configuration c; c.type_ = REAL; c.real_ = v; return c; }
	static Configuration ofBool(bool v) { Configuration c; c.type_ = BOOLEAN; c.flag_ = v; return c; }
	static Configuration ofString(const std::string &v) { Configuration c; c.type_ = TEXT; c.text_ = v; return c; }

	Type type() const { return type_; }
	bool has(const std::string &key) const { return type_ == DICT && dict_.find(key) != dict_.end(); }
	size_t size() const { return type_ == LIST ? list_.size() : (type_ == DICT ? dict_.size() : 0); }

	// entry of a dictionary / element of a list, created on demand (as Json::Value::operator[] does)
	Configuration &at(const std::string &key) {
		if(type_ == NONE) type_ = DICT;
		if(type_ != DICT) throw std::invalid_argument("QuantPDE_B200: configuration entry is not an object");
		return dict_[key];
	}
	Configuration &at(size_t index) {
		if(type_ == NONE) type_ = LIST;
		if(type_ != LIST) throw std::invalid_argument("QuantPDE_B200: configuration entry is not an array");
		if(index >= list_.size()) list_.resize(index + 1);
		return list_[index];
	}
	const Configuration *find(const std::string &key) const {
		if(type_ != DICT) return nullptr;
		std::map<std::string, Configuration>::const_iterator it = dict_.find(key);
		return it == dict_.end() ? nullptr : &it->second;
	}

	// conversions with JsonCpp's leniency between the numeric types and bool
	long long asInt() const {
		switch(type_) {
			case INTEGER: return integer_;
			case REAL: return (long long) real_;
			case BOOLEAN: return flag_ ? 1 : 0;
			case NONE: return 0;
			default: throw std::invalid_argument("QuantPDE_B200: configuration entry is not a number");
		}
	}
	Real asReal() const {
		switch(type_) {
			case INTEGER: return (Real) integer_;
			case REAL: return real_;
			case BOOLEAN: return flag_ ? 1. : 0.;
			case NONE: return 0.;
			default: throw std::invalid_argument("QuantPDE_B200: configuration entry is not a number");
		}
	}
	bool asBool() const {
		switch(type_) {
			case INTEGER: return integer_ != 0;
			case REAL: return real_ != 0.;
			case BOOLEAN: return flag_;
			case NONE: return false;
			default: throw std::invalid_argument("QuantPDE_B200: configuration entry is not a boolean");
		}
	}
	std::string asString() const {
		switch(type_) {
			case TEXT: return text_;
			case NONE: return "";
			case BOOLEAN: return flag_ ? "true" : "false";
			case INTEGER: return std::to_string(integer_);
			case REAL: { char buf[40]; std::snprintf(buf, sizeof buf, "%.17g", real_); return buf; }
			default: throw std::invalid_argument("QuantPDE_B200: configuration entry is not a string");
		}
	}

	// ---- JSON text in
	static Configuration parse(const std::string &text) {
		Reader rd(text);
		Configuration c = rd.value();
		rd.blanks();
		if(!rd.done()) rd.fail("trailing characters");
		return c;
	}

	// ---- JSON text out (objects sorted by key, one entry per line, tabs)
	void print(std::ostream &os, int depth = 0) const {
		const std::string pad((size_t) depth + 1, '\t'), close((size_t) depth, '\t');
		switch(type_) {
			case NONE: os << "null"; break;
			case INTEGER: os << integer_; break;
			case BOOLEAN: os << (flag_ ? "true" : "false"); break;
			case REAL: {
				char buf[40];
				std::snprintf(buf, sizeof buf, "%.17g", real_);
				os << buf;
				if(!std::strpbrk(buf, ".eEn")) os << ".0";
				break;
			}
			case TEXT: {
				os << '"';
				for(char ch : text_) {
					if(ch == '"' || ch == '\\') os << '\\' << ch;
					else if(ch == '\n') os << "\\n";
					else if(ch == '\t') os << "\\t";
					else os << ch;
				}
				os << '"';
				break;
			}
			case LIST: {
				if(list_.empty()) { os << "[]"; break; }
				os << "[\n";
				for(size_t k = 0; k < list_.size(); ++k) {
					os << pad; list_[k].print(os, depth + 1);
					os << (k + 1 < list_.size() ? ",\n" : "\n");
				}
				os << close << "]";
				break;
			}
			case DICT: {
				if(dict_.empty()) { os << "{}"; break; }
				os << "{\n";
				size_t k = 0;
				for(const auto &kv : dict_) {
					os << pad << '"' << kv.first << "\" : "; kv.second.print(os, depth + 1);
					os << (++k < dict_.size() ? ",\n" : "\n");
				}
				os << close << "}";
				break;
			}
		}
	}

private:
	Type type_;
	long long integer_;
	Real real_;
	bool flag_;
	std::string text_;
	std::vector<Configuration> list_;
	std::map<std::string, Configuration> dict_;

	// recursive-descent reader: RFC 8259 plus the // and /* */ comments JsonCpp tolerates
	class Reader {
		const std::string &s; size_t p;
	public:
		explicit Reader(const std::string &text) : s(text), p(0) {}
		bool done() const { return p >= s.size(); }
		void fail(const char *what) const {
			throw std::invalid_argument("QuantPDE_B200: configuration, offset " + std::to_string(p) + ": " + what);
		}
		void blanks() {
			while(p < s.size()) {
				const char ch = s[p];
				if(ch == ' ' || ch == '\n' || ch == '\t' || ch == '\r') { ++p; continue; }
				if(ch == '/' && p + 1 < s.size() && s[p + 1] == '/') { while(p < s.size() && s[p] != '\n') ++p; continue; }
				if(ch == '/' && p + 1 < s.size() && s[p + 1] == '*') {
					const size_t e = s.find("*/", p + 2);
					p = e == std::string::npos ? s.size() : e + 2;
					continue;
				}
				break;
			}
		}
		bool eat(char ch) { blanks(); if(p < s.size() && s[p] == ch) { ++p; return true; } return false; }
		bool word(const char *w) {
			const size_t n = std::strlen(w);
			if(s.compare(p, n, w) == 0) { p += n; return true; }
			return false;
		}
		std::string quoted() {
			std::string out;
			++p;                                  // opening quote
			for(;;) {
				if(p >= s.size()) fail("unterminated string");
				char ch = s[p++];
				if(ch == '"') return out;
				if(ch == '\\') {
					if(p >= s.size()) fail("unterminated escape");
					const char e = s[p++];
					switch(e) {
						case 'n': ch = '\n'; break; case 't': ch = '\t'; break; case 'r': ch = '\r'; break;
						case 'b': ch = '\b'; break; case 'f': ch = '\f'; break;
						case 'u': {
							if(p + 4 > s.size()) fail("short \\u escape");
							const unsigned code = (unsigned) std::strtoul(s.substr(p, 4).c_str(), nullptr, 16);
							p += 4;
							if(code < 0x80) { ch = (char) code; }
							else if(code < 0x800) { out += (char) (0xC0 | (code >> 6)); ch = (char) (0x80 | (code & 0x3F)); }
							else {
								out += (char) (0xE0 | (code >> 12)); out += (char) (0x80 | ((code >> 6) & 0x3F));
								ch = (char) (0x80 | (code & 0x3F));
							}
							break;
						}
						default: ch = e;
					}
				}
				out += ch;
			}
		}
		Configuration value() {
			blanks();
			if(done()) fail("unexpected end of input");
			const char ch = s[p];
			if(ch == '{') {
				++p;
				Configuration c; c.type_ = DICT;
				if(eat('}')) return c;
				do {
					blanks();
					if(done() || s[p] != '"') fail("expected a member name");
					const std::string key = quoted();
					if(!eat(':')) fail("expected ':'");
					c.dict_[key] = value();
				} while(eat(','));
				if(!eat('}')) fail("expected ',' or '}'");
				return c;
			}
			if(ch == '[') {
				++p;
				Configuration c; c.type_ = LIST;
				if(eat(']')) return c;
				do { c.list_.push_back(value()); } while(eat(','));
				if(!eat(']')) fail("expected ',' or ']'");
				return c;
			}
			if(ch == '"') return ofString(quoted());
			if(word("true")) return ofBool(true);
			if(word("false")) return ofBool(false);
			if(word("null")) return Configuration();
			const size_t begin = p;
			bool real = false;
			if(p < s.size() && (s[p] == '-' || s[p] == '+')) ++p;
			while(p < s.size() && (std::isdigit((unsigned char) s[p]) || std::strchr(".eE+-", s[p]))) {
				if(s[p] == '.' || s[p] == 'e' || s[p] == 'E') real = true;
				++p;
			}
			if(p == begin) fail("unexpected character");
			const std::string num = s.substr(begin, p - begin);
			return real ? ofReal(std::strtod(num.c_str(), nullptr)) : ofInt(std::strtoll(num.c_str(), nullptr, 10));
		}
	};
};

inline std::ostream &operator<<(std::ostream &os, const Configuration &c) {
	os << "\n"; c.print(os); os << "\n";
	return os;
}
inline std::istream &operator>>(std::istream &is, Configuration &c) {
	const std::string text((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
	c = Configuration::parse(text);
	return is;
}

// Configuration.hpp:17-34: read `key` (default if absent) and write the value used back into the
// configuration, so that printing it afterwards shows every parameter of the run
inline int getInt(Configuration &configuration, const std::string &key, int defaultValue) {
	const Configuration *e = configuration.find(key);
	const int v = e ? (int) e->asInt() : defaultValue;
	configuration.at(key) = Configuration::ofInt(v);
	return v;
}
inline Real getReal(Configuration &configuration, const std::string &key, Real defaultValue) {
	const Configuration *e = configuration.find(key);
	const Real v = e ? e->asReal() : defaultValue;
	configuration.at(key) = Configuration::ofReal(v);
	return v;
}
inline bool getBool(Configuration &configuration, const std::string &key, bool defaultValue) {
	const Configuration *e = configuration.find(key);
	const bool v = e ? e->asBool() : defaultValue;
	configuration.at(key) = Configuration::ofBool(v);
	return v;
}
inline std::string getString(Configuration &configuration, const std::string &key, const std::string &defaultValue) {
	const Configuration *e = configuration.find(key);
	const std::string v = e ? e->asString() : defaultValue;
	configuration.at(key) = Configuration::ofString(v);
	return v;
}

// Configuration.hpp:78-134 for Dimension = 1: "key" : [ [ tick, tick, ... ] ]
inline RectilinearGrid1 getGrid(Configuration &configuration, const std::string &key, const RectilinearGrid1 &defaultValue) {
	if(!configuration.has(key)) {
		const std::vector<Real> ticks = defaultValue.nodes();
		for(size_t j = 0; j < ticks.size(); ++j) configuration.at(key).at(0).at(j) = Configuration::ofReal(ticks[j]);
		return defaultValue;
	}
	Configuration &axis = configuration.at(key).at(0);
	std::vector<Real> ticks(axis.size());
	for(size_t j = 0; j < ticks.size(); ++j) ticks[j] = axis.at(j).asReal();
	return RectilinearGrid1(Axis(std::move(ticks)));
}

// Configuration.hpp:36-76: `prog [CONF_FILE | -i]`; -h prints the usage and exits with 2.
inline Configuration getConfiguration(int argc, char **argv) {
	bool input = false;
	const char *file = nullptr;
	std::vector<std::pair<std::string, std::string>> overrides;
	for(int k = 1; k < argc; ++k) {
		const char *a = argv[k];
		if(std::strcmp(a, "-h") == 0) {
			std::cerr << argv[0] << " [CONF_FILE | -i] [key=value ...]" << std::endl << std::endl
				<< "CONF_FILE" << std::endl
				<< "    Path to a JSON file used to specify parameters." << std::endl << std::endl
				<< "-i" << std::endl
				<< "    Places the binary in interactive mode so that the contents of a " << std::endl
				<< "    configuration file can be piped in or inputted with a terminating EOF char." << std::endl << std::endl
				<< "key=value" << std::endl
				<< "    Overrides one entry of the configuration (value in JSON syntax)." << std::endl;
			std::exit(2);
		} else if(std::strcmp(a, "-i") == 0) {
			input = true;
		} else if(const char *eq = std::strchr(a, '=')) {
			overrides.emplace_back(std::string(a, eq), std::string(eq + 1));
		} else if(!file) {
			file = a;
		}
	}
	Configuration configuration;
	if(input) {
		std::cin >> configuration;
	} else if(file) {
		std::ifstream in(file, std::ifstream::binary);
		if(!in) throw std::invalid_argument(std::string("QuantPDE_B200: cannot open ") + file);
		in >> configuration;
	}
	for(const auto &kv : overrides) {
		Configuration v;
		try { v = Configuration::parse(kv.second); }
		catch(const std::invalid_argument &) { v = Configuration::ofString(kv.second); }   // bare words are strings
		configuration.at(kv.first) = v;
	}
	return configuration;
}

////////////////////////////////////////////////////////////////////////////////
// ResultsBuffer1 — Results.hpp:22-171
////////////////////////////////////////////////////////////////////////////////

typedef std::tuple<std::vector<Real>, Solution, Real> ResultsTuple1;

class ResultsBuffer1 {
	typedef std::vector<std::string> Headers;
	std::function<ResultsTuple1 (int)> run;
	const Headers headers;
	const int kn, k0, precision, spacing;
	const bool ratio;
	std::unique_ptr<RectilinearGrid1> grid;
public:
	template <typename F>
	ResultsBuffer1(F &&run, const Headers &headers = {}, int kn = 5, int k0 = 0, int precision = 6, int spacing = 23,
			bool ratio = true)
			: run(std::forward<F>(run)), headers(headers), kn(kn), k0(k0), precision(precision), spacing(spacing),
			  ratio(ratio) {}
	ResultsBuffer1(const ResultsBuffer1 &) = delete;
	ResultsBuffer1 &operator=(const ResultsBuffer1 &) = delete;

	void setPrintGrid(const RectilinearGrid1 &g) { grid.reset(new RectilinearGrid1(g)); }

	void stream(std::ostream &os = std::cout) {
		const std::ios::fmtflags saved(os.flags());
		for(const std::string &h : headers) os << std::setw(spacing) << h;
		os << std::setw(spacing) << "Value";
		if(ratio) os << std::setw(spacing) << "Change" << std::setw(spacing) << "Ratio";
		os << std::setw(spacing) << "Timing (Seconds)" << std::endl;
		os.precision(precision);

		Real previousValue = std::nan(""), previousChange = std::nan("");
		for(int k = k0; k <= kn; ++k) {
			const auto start = std::chrono::steady_clock::now();
			ResultsTuple1 ret = run(k);
			const Real seconds = std::chrono::duration<Real>(std::chrono::steady_clock::now() - start).count();
			const std::vector<Real> &columns = std::get<0>(ret);
			const Solution &solution = std::get<1>(ret);
			const Real point = std::get<2>(ret);

			// integral columns in fixed notation, the rest in scientific (Results.hpp:115-125)
			for(Real column : columns) {
				Real intpart;
				if(std::abs(std::modf(column, &intpart)) < epsilon) os << std::fixed; else os << std::scientific;
				os << std::setw(spacing) << column;
			}
			const Real value = solution(point);
			os << std::scientific << std::setw(spacing) << value;
			if(ratio) {
				const Real change = value - previousValue;
				os << std::setw(spacing) << change << std::setw(spacing) << previousChange / change;
				previousChange = change;
				previousValue = value;
			}
			os << std::setw(spacing) << seconds << std::endl;
			os.flags(saved);

			// the solution on the print grid after the last level (Results.hpp:148-164, Domain.hpp:417-431)
			if(k == kn && grid) {
				os << std::endl << std::setw(spacing) << "x_1" << std::setw(spacing) << "Solution at x" << std::endl;
				for(Real x : grid->nodes()) os << std::setw(spacing) << x << std::setw(spacing) << solution(x) << std::endl;
			}
		}
	}
};

} // namespace QuantPDE_B200

#endif
