/* test-infrastructure shim (oracle/_ref build): only the TYPES tools/global.h:92-134 names in the
 * declaration of class Global. tools/global.cpp (the option parser) is not part of the hot path and is
 * not compiled; oracle/ref_glue.cpp supplies Global's members (verbose level 0, no options set). */
#pragma once
#include <map>
#include <string>
#include <vector>
#include <stdexcept>
namespace boost { namespace program_options {
class variable_value {
public:
	bool empty() const { return true; }
	template <class T> const T& as() const { throw std::runtime_error("program_options shim: no options are ever set"); }
};
class variables_map : public std::map<std::string, variable_value> {};
class options_description {
public:
	options_description() {}
	explicit options_description(const std::string&) {}
	options_description& add(const options_description&) { return *this; }
};
}}
