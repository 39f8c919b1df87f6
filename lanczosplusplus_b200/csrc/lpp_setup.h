// lpp_setup.h -- host-side bookkeeping shared by the engine and the CPU-side unit tests of lpp_device.cuh:
// binomial table and the FeAs partition layout (Partitions.h:32-77, BasisOneSpinFeAs.h:56-84).
#pragma once
#include <cstdint>
#include <vector>
#include "lpp_device.cuh"

inline std::vector<uint64_t> lpp_make_binom()
{
	std::vector<uint64_t> binom((size_t)LPP_BINOM_N * LPP_BINOM_N, 0);
	for (int n = 0; n < LPP_BINOM_N; n++) {
		binom[n * LPP_BINOM_N] = 1;
		for (int k = 1; k <= n; k++)
			binom[n * LPP_BINOM_N + k] = binom[(n - 1) * LPP_BINOM_N + k - 1] + (k <= n - 1 ? binom[(n - 1) * LPP_BINOM_N + k] : 0);
	}
	return binom;
}

// Partitions.h:32-77 restated: odometer with orbital 0 fastest; only tuples summing to npart are kept
inline void lpp_make_partitions(int npart, int no, std::vector<std::vector<int>>& out)
{
	std::vector<int> v(no, 0);
	auto sum = [&]() { int s = 0; for (int x : v) s += x; return s; };
	while (true) {
		if (sum() == npart) out.push_back(v);
		v[0]++;
		if (sum() > npart) {
			if (no == 1) break;
			v[0] = 0;
			int i = 1;
			bool done = false;
			while (true) {
				v[i]++;
				if (sum() <= npart) break;
				if (i == no - 1) { done = true; break; }
				v[i] = 0;
				i++;
			}
			if (done) break;
		}
	}
}

struct LppFeasLayout {
	std::vector<uint64_t> off;    // block offset by key = sum_{o>=1} n_o*(npart+1)^(o-1)
	std::vector<uint64_t> start;  // non-empty partition p starts at start[p]; start.back() = total
	std::vector<int> pn;          // occupations, LPP_MAX_ORB per non-empty partition
	uint64_t total = 0;
};

inline LppFeasLayout lpp_feas_layout(const std::vector<uint64_t>& binom, int nsite, int no, int npart)
{
	LppFeasLayout L;
	std::vector<std::vector<int>> parts;
	uint64_t keysz = 1;
	for (int o = 1; o < no; o++) keysz *= (uint64_t)(npart + 1);
	L.off.assign(keysz, 0);
	if (npart == 0) parts.push_back(std::vector<int>(no, 0));  // BasisOneSpinFeAs.h:57-62
	else lpp_make_partitions(npart, no, parts);
	for (auto& p : parts) {
		uint64_t sz = 1;
		for (int o = 0; o < no; o++) sz *= (p[o] <= nsite) ? binom[nsite * LPP_BINOM_N + p[o]] : 0;
		uint64_t key = 0, mul = 1;
		for (int o = 1; o < no; o++) { key += (uint64_t)p[o] * mul; mul *= (uint64_t)(npart + 1); }
		L.off[key] = L.total;
		if (sz == 0) continue;
		L.start.push_back(L.total);
		for (int o = 0; o < LPP_MAX_ORB; o++) L.pn.push_back(o < no ? p[o] : 0);
		L.total += sz;
	}
	L.start.push_back(L.total);
	return L;
}
