/* test-infrastructure shim: boost::split / boost::is_any_of as used by tools/useful.cpp:81,97 */
#pragma once
#include <string>
#include <vector>
namespace boost {
struct is_any_of_t { std::string set; };
inline is_any_of_t is_any_of(const std::string& s) { return is_any_of_t{s}; }
inline is_any_of_t is_any_of(const char* s) { return is_any_of_t{std::string(s)}; }
inline is_any_of_t is_any_of(char c) { return is_any_of_t{std::string(1, c)}; }
template <class Seq>
inline Seq& split(Seq& out, const std::string& in, const is_any_of_t& pred)
{
	out.clear();
	std::string cur;
	for (char c : in) {
		if (pred.set.find(c) != std::string::npos) { out.push_back(cur); cur.clear(); }
		else cur.push_back(c);
	}
	out.push_back(cur);
	return out;
}
}
