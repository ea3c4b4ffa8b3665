/* TEST INFRASTRUCTURE (oracle/_ref build only): in-memory registry of what the reference's file writers
 * were asked to write (see ref_glue.cpp). Keys are "<file>:<group>/<dataset>". */
#pragma once
#include <string>
#include <vector>
struct ref_dataset {
	std::vector<size_t> dims;
	std::vector<double> data;
};
void ref_recorder_clear();
void ref_recorder_put(const std::string& key, const ref_dataset& ds);
const ref_dataset* ref_recorder_get(const std::string& key);
std::vector<std::string> ref_recorder_keys();
