// entry_set.h -- de-duplication of operator coefficient tuples (memcmp equality, like the
// SSE_coeff key of Operator_SSE_Compressed, FDTD/operator_sse_compressed.cpp:182-225).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/openems_b200.h"

namespace {
struct EntrySet {
	std::vector<oems_coeff_entry> items;
	std::vector<int64_t> slots;
	size_t mask;
	EntrySet() : slots(1 << 12, -1), mask((1 << 12) - 1) {}
	static uint64_t hash(const oems_coeff_entry& e)
	{
		const uint64_t* w = reinterpret_cast<const uint64_t*>(&e);
		uint64_t h = 0x9E3779B97F4A7C15ull;
		for (size_t i = 0; i < sizeof(oems_coeff_entry) / 8; ++i) {
			h ^= w[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
			h *= 0xff51afd7ed558ccdull;
		}
		return h ^ (h >> 29);
	}
	void grow()
	{
		std::vector<int64_t> ns(slots.size() * 4, -1);
		const size_t nm = ns.size() - 1;
		for (size_t u = 0; u < items.size(); ++u) {
			size_t s = hash(items[u]) & nm;
			while (ns[s] >= 0) s = (s + 1) & nm;
			ns[s] = (int64_t)u;
		}
		slots.swap(ns);
		mask = nm;
	}
	uint32_t insert(const oems_coeff_entry& e)
	{
		size_t s = hash(e) & mask;
		while (slots[s] >= 0) {
			if (memcmp(&items[slots[s]], &e, sizeof(e)) == 0) return (uint32_t)slots[s];
			s = (s + 1) & mask;
		}
		slots[s] = (int64_t)items.size();
		items.push_back(e);
		if (items.size() * 2 > slots.size()) grow();
		return (uint32_t)(items.size() - 1);
	}
};
} // namespace
