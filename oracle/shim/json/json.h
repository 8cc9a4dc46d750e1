// TEST INFRASTRUCTURE ONLY (oracle). Never included by the product path.
//
// A stand-in for <json/json.h> (JsonCpp, not installed in this image) covering exactly what the
// reference's Modules/Utilities/Configuration.hpp and its example programs use, so that the
// UNMODIFIED examples under /root/reference/examples compile here (oracle/Makefile: `examples`)
// and their stdout — the convergence tables of Modules/Utilities/Results.hpp:79-168 — can be
// recorded as fixtures:
//
//   Json::Value: default ctor, ctor / assignment from int, double, bool, string;
//                get(key, default), operator[](key), operator[](index), isMember(key), size(),
//                asInt / asDouble / asBool / asString;  operator>> (parse), operator<< (styled).
//
// Objects keep their keys sorted (JsonCpp stores members in a std::map).  The writer's layout
// follows JsonCpp's StyledStreamWriter closely but is not claimed to be byte-identical: the
// examples send it to stderr, the tables go to stdout.
#ifndef QPDE_ORACLE_JSON_SHIM_H
#define QPDE_ORACLE_JSON_SHIM_H

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <iterator>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace Json {

class Value {
public:
	enum Kind { Null, Int, Real, Bool, String, Array, Object };

	Value() : kind_(Null), i_(0), d_(0.), b_(false) {}
	Value(int v) : kind_(Int), i_(v), d_(v), b_(v != 0) {}
	Value(unsigned v) : kind_(Int), i_((long long) v), d_(v), b_(v != 0) {}
	Value(long long v) : kind_(Int), i_(v), d_((double) v), b_(v != 0) {}
	Value(double v) : kind_(Real), i_((long long) v), d_(v), b_(v != 0.) {}
	Value(bool v) : kind_(Bool), i_(v), d_(v), b_(v) {}
	Value(const char *v) : kind_(String), i_(0), d_(0.), b_(false), s_(v) {}
	Value(const std::string &v) : kind_(String), i_(0), d_(0.), b_(false), s_(v) {}

	Kind kind() const { return kind_; }
	bool isNull() const { return kind_ == Null; }

	int asInt() const {
		switch(kind_) {
			case Int: return (int) i_;
			case Real: return (int) d_;
			case Bool: return b_ ? 1 : 0;
			case Null: return 0;
			default: throw std::runtime_error("Json shim: value is not convertible to Int");
		}
	}
	double asDouble() const {
		switch(kind_) {
			case Int: return (double) i_;
			case Real: return d_;
			case Bool: return b_ ? 1. : 0.;
			case Null: return 0.;
			default: throw std::runtime_error("Json shim: value is not convertible to double");
		}
	}
	bool asBool() const {
		switch(kind_) {
			case Int: return i_ != 0;
			case Real: return d_ != 0.;
			case Bool: return b_;
			case Null: return false;
			default: throw std::runtime_error("Json shim: value is not convertible to bool");
		}
	}
	std::string asString() const {
		switch(kind_) {
			case String: return s_;
			case Null: return "";
			case Bool: return b_ ? "true" : "false";
			case Int: return std::to_string(i_);
			case Real: { char buf[64]; std::snprintf(buf, sizeof buf, "%.17g", d_); return buf; }
			default: throw std::runtime_error("Json shim: value is not convertible to string");
		}
	}

	bool isMember(const std::string &key) const {
		return kind_ == Object && members_.count(key) != 0;
	}
	Value get(const std::string &key, const Value &fallback) const {
		if(kind_ != Object) return fallback;
		std::map<std::string, Value>::const_iterator it = members_.find(key);
		return it == members_.end() ? fallback : it->second;
	}
	Value &operator[](const std::string &key) {
		if(kind_ == Null) kind_ = Object;
		if(kind_ != Object) throw std::runtime_error("Json shim: operator[](key) on a non-object");
		return members_[key];
	}
	Value &operator[](const char *key) { return (*this)[std::string(key)]; }
	Value &operator[](int index) {
		if(kind_ == Null) kind_ = Array;
		if(kind_ != Array) throw std::runtime_error("Json shim: operator[](index) on a non-array");
		if((size_t) index >= items_.size()) items_.resize((size_t) index + 1);
		return items_[(size_t) index];
	}
	unsigned size() const {
		if(kind_ == Array) return (unsigned) items_.size();
		if(kind_ == Object) return (unsigned) members_.size();
		return 0;
	}

	// ---- writer
	void write(std::ostream &os, int depth) const {
		const std::string in((size_t) depth, '\t'), in1((size_t) depth + 1, '\t');
		switch(kind_) {
			case Null: os << "null"; break;
			case Int: os << i_; break;
			case Bool: os << (b_ ? "true" : "false"); break;
			case Real: {
				char buf[64];
				std::snprintf(buf, sizeof buf, "%.17g", d_);
				std::string t(buf);
				if(t.find_first_of(".eEn") == std::string::npos) t += ".0";
				os << t;
				break;
			}
			case String: {
				os << '"';
				for(size_t k = 0; k < s_.size(); ++k) {
					const char c = s_[k];
					if(c == '"' || c == '\\') os << '\\' << c;
					else if(c == '\n') os << "\\n";
					else if(c == '\t') os << "\\t";
					else os << c;
				}
				os << '"';
				break;
			}
			case Array: {
				if(items_.empty()) { os << "[]"; break; }
				bool flat = true;
				for(size_t k = 0; k < items_.size(); ++k)
					if(items_[k].kind_ == Array || items_[k].kind_ == Object) flat = false;
				if(flat && items_.size() * 3 < 74) {
					os << "[ ";
					for(size_t k = 0; k < items_.size(); ++k) { if(k) os << ", "; items_[k].write(os, depth + 1); }
					os << " ]";
				} else {
					os << "[\n";
					for(size_t k = 0; k < items_.size(); ++k) {
						os << in1; items_[k].write(os, depth + 1);
						os << (k + 1 < items_.size() ? ",\n" : "\n");
					}
					os << in << "]";
				}
				break;
			}
			case Object: {
				if(members_.empty()) { os << "{}"; break; }
				os << "{\n";
				size_t k = 0;
				for(std::map<std::string, Value>::const_iterator it = members_.begin(); it != members_.end(); ++it, ++k) {
					os << in1 << '"' << it->first << "\" : ";
					it->second.write(os, depth + 1);
					os << (k + 1 < members_.size() ? ",\n" : "\n");
				}
				os << in << "}";
				break;
			}
		}
	}

	// ---- reader (RFC 8259 plus // and /* */ comments, which JsonCpp accepts by default)
	static Value parse(const std::string &text) {
		size_t at = 0;
		Value v = parseValue(text, at);
		skip(text, at);
		if(at != text.size()) throw std::runtime_error("Json shim: trailing characters");
		return v;
	}

private:
	Kind kind_;
	long long i_;
	double d_;
	bool b_;
	std::string s_;
	std::vector<Value> items_;
	std::map<std::string, Value> members_;

	static void skip(const std::string &t, size_t &at) {
		for(;;) {
			while(at < t.size() && (t[at] == ' ' || t[at] == '\t' || t[at] == '\n' || t[at] == '\r')) ++at;
			if(at + 1 < t.size() && t[at] == '/' && t[at + 1] == '/') {
				while(at < t.size() && t[at] != '\n') ++at;
			} else if(at + 1 < t.size() && t[at] == '/' && t[at + 1] == '*') {
				at += 2;
				while(at + 1 < t.size() && !(t[at] == '*' && t[at + 1] == '/')) ++at;
				at += 2;
			} else break;
		}
	}
	static std::string parseString(const std::string &t, size_t &at) {
		std::string out;
		++at;
		while(at < t.size() && t[at] != '"') {
			char c = t[at++];
			if(c == '\\' && at < t.size()) {
				const char e = t[at++];
				switch(e) {
					case 'n': c = '\n'; break; case 't': c = '\t'; break; case 'r': c = '\r'; break;
					case 'b': c = '\b'; break; case 'f': c = '\f'; break;
					case 'u': {
						const unsigned code = (unsigned) std::strtoul(t.substr(at, 4).c_str(), 0, 16);
						at += 4;
						if(code < 0x80) c = (char) code;
						else if(code < 0x800) { out += (char) (0xC0 | (code >> 6)); c = (char) (0x80 | (code & 0x3F)); }
						else { out += (char) (0xE0 | (code >> 12)); out += (char) (0x80 | ((code >> 6) & 0x3F)); c = (char) (0x80 | (code & 0x3F)); }
						break;
					}
					default: c = e;
				}
			}
			out += c;
		}
		if(at >= t.size()) throw std::runtime_error("Json shim: unterminated string");
		++at;
		return out;
	}
	static Value parseValue(const std::string &t, size_t &at) {
		skip(t, at);
		if(at >= t.size()) throw std::runtime_error("Json shim: unexpected end of input");
		const char c = t[at];
		if(c == '{') {
			Value v; v.kind_ = Object;
			++at; skip(t, at);
			if(at < t.size() && t[at] == '}') { ++at; return v; }
			for(;;) {
				skip(t, at);
				if(at >= t.size() || t[at] != '"') throw std::runtime_error("Json shim: expected a member name");
				const std::string key = parseString(t, at);
				skip(t, at);
				if(at >= t.size() || t[at] != ':') throw std::runtime_error("Json shim: expected ':'");
				++at;
				v.members_[key] = parseValue(t, at);
				skip(t, at);
				if(at < t.size() && t[at] == ',') { ++at; continue; }
				if(at < t.size() && t[at] == '}') { ++at; return v; }
				throw std::runtime_error("Json shim: expected ',' or '}'");
			}
		}
		if(c == '[') {
			Value v; v.kind_ = Array;
			++at; skip(t, at);
			if(at < t.size() && t[at] == ']') { ++at; return v; }
			for(;;) {
				v.items_.push_back(parseValue(t, at));
				skip(t, at);
				if(at < t.size() && t[at] == ',') { ++at; continue; }
				if(at < t.size() && t[at] == ']') { ++at; return v; }
				throw std::runtime_error("Json shim: expected ',' or ']'");
			}
		}
		if(c == '"') return Value(parseString(t, at));
		if(t.compare(at, 4, "true") == 0) { at += 4; return Value(true); }
		if(t.compare(at, 5, "false") == 0) { at += 5; return Value(false); }
		if(t.compare(at, 4, "null") == 0) { at += 4; return Value(); }
		// number: an integer literal stays an Int, anything with '.', 'e' or 'E' is a Real
		size_t end = at;
		bool real = false;
		if(end < t.size() && (t[end] == '-' || t[end] == '+')) ++end;
		while(end < t.size() && ((t[end] >= '0' && t[end] <= '9') || t[end] == '.' || t[end] == 'e' || t[end] == 'E'
				|| t[end] == '-' || t[end] == '+')) {
			if(t[end] == '.' || t[end] == 'e' || t[end] == 'E') real = true;
			++end;
		}
		if(end == at) throw std::runtime_error("Json shim: unexpected character");
		const std::string num = t.substr(at, end - at);
		at = end;
		if(real) return Value(std::strtod(num.c_str(), 0));
		return Value((long long) std::strtoll(num.c_str(), 0, 10));
	}
};

inline std::istream &operator>>(std::istream &is, Value &v) {
	const std::string text((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
	v = Value::parse(text);
	return is;
}

inline std::ostream &operator<<(std::ostream &os, const Value &v) {
	os << "\n";
	v.write(os, 0);
	os << "\n";
	return os;
}

} // namespace Json

#endif
