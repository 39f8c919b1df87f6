// tests/dblock_plan_check.cu -- TEST-ONLY, runs on the CPU (no device call).  Builds the two-pass block plan of
// lpp_dblock_kernel.cuh for small Hubbard-type bases and walks its tables on the host exactly the way k_dblock does
// (tile slot 0 = zero line, "+" quads/pair then "-" quads/pair per step, 4 states per step), then compares
//   x = beta x + alpha (U0 popc(up & dn) + dv2 + T_dn) y
// with the plain ELL application.  Also checks the structural invariants: every hop of the table appears in exactly one
// pass, no operand leaves its block, every bond misses the fixed sites of at least one pass, the plan fits its CTAs per SM.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../lanczosplusplus_b200/csrc/lpp_dblock_kernel.cuh"

static int run_case(int nx, int ny, int npart, bool periodic, bool potential, int layout = 0, int passes = 0)
{
	const int nsite = nx * ny;
	std::vector<uint32_t> words;
	for (uint32_t s = 0; s < (1u << nsite); s++)
		if (__builtin_popcount(s) == npart) words.push_back(s);
	const uint64_t n = words.size();
	std::vector<double> hop((size_t)nsite * nsite, 0.0);
	auto site = [&](int x, int y) { return ((x + nx) % nx) + nx * ((y + ny) % ny); };
	for (int y = 0; y < ny; y++)
		for (int x = 0; x < nx; x++) {
			const int i = site(x, y);
			const int nb[2] = {site(x + 1, y), site(x, y + 1)};
			for (int b = 0; b < 2; b++) {
				if (b == 1 && ny == 1) continue;
				if (!periodic && ((b == 0 && x == nx - 1) || (b == 1 && y == ny - 1))) continue;
				const int j = nb[b];
				if (j == i) continue;
				hop[i * nsite + j] = -1.0;
				hop[j * nsite + i] = -1.0;
			}
		}
	std::vector<uint32_t> lut(1u << nsite, 0xffffffffu);
	for (uint64_t s = 0; s < n; s++) lut[words[s]] = (uint32_t)s;
	std::vector<std::vector<std::pair<uint32_t, double>>> hl(n);
	int width = 1;
	for (uint64_t s = 0; s < n; s++) {
		const uint32_t w = words[s];
		for (int i = 0; i < nsite; i++)
			for (int j = 0; j < nsite; j++) {
				const double h = hop[i * nsite + j];
				if (h == 0 || !((w >> i) & 1) || ((w >> j) & 1)) continue;
				const int lo = std::min(i, j), hi = std::max(i, j);
				const uint32_t between = ((1u << hi) - 1) & ~((1u << (lo + 1)) - 1);
				const double sg = (__builtin_popcount(w & between) & 1) ? -1.0 : 1.0;
				hl[s].push_back({lut[w ^ (1u << i) ^ (1u << j)], h * sg});
			}
		width = std::max<int>(width, (int)hl[s].size());
	}
	std::vector<uint32_t> idx((size_t)width * n), cnt(n);
	std::vector<double> val((size_t)width * n, 0.0), dv2(n, 0.0);
	uint64_t nhops = 0;
	for (uint64_t s = 0; s < n; s++) {
		cnt[s] = (uint32_t)hl[s].size();
		nhops += cnt[s];
		for (int k = 0; k < width; k++) {
			idx[(size_t)k * n + s] = k < (int)cnt[s] ? hl[s][k].first : (uint32_t)s;
			val[(size_t)k * n + s] = k < (int)cnt[s] ? hl[s][k].second : 0.0;
		}
		if (potential) dv2[s] = 0.01 * (double)(words[s] % 17) - 0.05;
	}
	DbHostPlan hp;
	std::string err;
	if (!db_build_host_plan(words.data(), n, nsite, idx.data(), val.data(), cnt.data(), width, dv2.data(), (size_t)232448, (size_t)233472, layout, passes, &hp, &err)) {
		std::printf("%dx%d N=%d %s: no plan (%s)\n", nx, ny, npart, periodic ? "pbc" : "open", err.c_str());
		return nhops == 0 ? 0 : 2;
	}
	// invariants
	for (int i = 0; i < nsite; i++)
		for (int j = 0; j < nsite; j++) {
			if (hop[i * nsite + j] == 0) continue;
			bool free_pass = false;                                   // every bond misses the fixed sites of some pass
			for (int k = 0; k < hp.npass; k++) free_pass = free_pass || !(((hp.fmask[k] >> i) | (hp.fmask[k] >> j)) & 1u);
			if (!free_pass) { std::printf("FAIL a bond touches the fixed sites of every pass\n"); return 1; }
		}
	if (passes && hp.npass != passes) { std::printf("FAIL %d passes asked for, %d planned\n", passes, hp.npass); return 1; }
	if (hp.smem_bytes + (size_t)hp.max_pos * 8 + 2048 > (size_t)233472 / hp.ctas_per_sm) { std::printf("FAIL plan does not fit %d CTA(s) per SM\n", hp.ctas_per_sm); return 1; }
	// host walk of the tables (two columns are enough: the kernel treats columns independently)
	const int ncol = 2;
	std::vector<double> y(n * ncol), x(n * ncol), xr(n * ncol);
	for (size_t i = 0; i < y.size(); i++) { y[i] = std::sin(0.37 * (double)i + 0.1); x[i] = std::cos(0.11 * (double)i); }
	xr = x;
	const double alpha = 0.37, beta = -0.81, U0 = 4.0, tmag = hp.tmag;
	const uint32_t upw[2] = {words[n / 3], words[(2 * n) / 3]};
	for (uint64_t s = 0; s < n; s++)
		for (int c = 0; c < ncol; c++) {
			double acc = (U0 * (double)__builtin_popcount(upw[c] & words[s]) + dv2[s]) * y[s * ncol + c];
			for (uint32_t k = 0; k < cnt[s]; k++) acc += val[(size_t)k * n + s] * y[(size_t)idx[(size_t)k * n + s] * ncol + c];
			xr[s * ncol + c] = beta * xr[s * ncol + c] + alpha * acc;
		}
	uint64_t walked = 0;
	for (int pass = 0; pass < hp.npass; pass++) {
		const DbHostPass& P = hp.pass[pass];
		for (const DbBlock& b : P.blocks) {
			const uint4* blob = P.blob.data() + b.blob_off;
			const uint32_t npos = b.nsteps * 4;
			const uint4* meta = blob;
			const uint32_t* info = reinterpret_cast<const uint32_t*>(blob + npos);
			const uint32_t* tab = reinterpret_cast<const uint32_t*>(blob + npos + ((b.nsteps + 3) >> 2));
			std::vector<double> tile(((size_t)npos + 1) * ncol, 0.0);          // slot 0 = zero line
			for (uint32_t p = 0; p < b.nstates; p++) {
				const uint32_t row = P.rows[b.rows_off + p];
				if (row != meta[p].x) { std::printf("FAIL row list and meta disagree\n"); return 1; }
				for (int c = 0; c < ncol; c++) tile[((size_t)p + 1) * ncol + c] = y[(size_t)row * ncol + c];
			}
			for (uint32_t st = 0; st < b.nsteps; st++) {
				const uint32_t pp = (info[st] >> 20) & 63u, pm = info[st] >> 26;
				for (uint32_t q = 0; q < 4; q++) {
					const uint32_t pos = st * 4 + q;
					const uint4 m = meta[pos];
					if (m.x == DB_ROW_NONE) continue;
					double plus[2] = {0, 0}, minus[2] = {0, 0};
					uint32_t unit = info[st] & 0x000fffffu;                     // 32-byte units = 8 words
					for (int sgn = 0; sgn < 2; sgn++) {
						const uint32_t np = sgn ? pm : pp;
						double* acc = sgn ? minus : plus;
						for (uint32_t g = 0; g < np / 2; g++, unit += 2)
							for (int e = 0; e < 4; e++) {
								const uint32_t off = tab[(size_t)unit * 8 + q * 4 + e];
								if (off % DB_LINE || off / DB_LINE > npos) { std::printf("FAIL operand outside the tile\n"); return 1; }
								if (off) walked++;
								for (int c = 0; c < ncol; c++) acc[c] += tile[(size_t)(off / DB_LINE) * ncol + c];
							}
						if (np & 1) {
							for (int e = 0; e < 2; e++) {
								const uint32_t off = tab[(size_t)unit * 8 + q * 2 + e];
								if (off % DB_LINE || off / DB_LINE > npos) { std::printf("FAIL operand outside the tile\n"); return 1; }
								if (off) walked++;
								for (int c = 0; c < ncol; c++) acc[c] += tile[(size_t)(off / DB_LINE) * ncol + c];
							}
							unit += 1;
						}
					}
					double d2;
					const unsigned long long bits = ((unsigned long long)m.w << 32) | m.z;
					memcpy(&d2, &bits, 8);
					for (int c = 0; c < ncol; c++) {
						double h = tmag * (plus[c] - minus[c]);
						double& xe = x[(size_t)m.x * ncol + c];
						if (pass == 0) {
							h += (U0 * (double)__builtin_popcount(upw[c] & m.y) + d2) * tile[((size_t)pos + 1) * ncol + c];
							xe = alpha * h + beta * xe;
						} else {
							xe += alpha * h;
						}
					}
				}
			}
		}
	}
	double maxd = 0, maxv = 0;
	for (size_t i = 0; i < x.size(); i++) { maxd = std::max(maxd, std::fabs(x[i] - xr[i])); maxv = std::max(maxv, std::fabs(xr[i])); }
	const bool ok = walked == nhops && maxd <= 1e-13 * std::max(1.0, maxv);
	uint64_t slots = 0;
	for (int k = 0; k < hp.npass; k++) slots += hp.pass[k].exec_slots;
	std::printf("%dx%d N=%d %s%s: %llu states, %d passes x %d CTA/SM, F %#x %#x %#x, max block %u, slots/state %.2f, hops %llu walked %llu, max diff %.2e %s\n", nx, ny,
	            npart, periodic ? "pbc" : "open", potential ? " +V" : "", (unsigned long long)n, hp.npass, hp.ctas_per_sm, hp.fmask[0], hp.fmask[1], hp.fmask[2],
	            hp.max_pos, (double)slots / (double)n, (unsigned long long)nhops, (unsigned long long)walked, maxd, ok ? "ok" : "MISMATCH");
	return ok ? 0 : 1;
}

int main()
{
	int bad = 0;
	bad += run_case(4, 3, 6, true, false);
	bad += run_case(4, 3, 5, true, true);
	bad += run_case(12, 1, 6, false, true);
	bad += run_case(14, 1, 3, false, false);
	bad += run_case(8, 1, 4, true, false);
	bad += run_case(3, 3, 4, true, true);       // the 3x3 torus is not bipartite
	bad += run_case(4, 4, 3, true, false);
	bad += run_case(4, 4, 8, true, true);        // config 3: three passes, two CTAs per SM
	bad += run_case(4, 4, 8, true, false, 1);    // the same basis in the one-CTA layout (two passes)
	bad += run_case(4, 4, 8, true, true, 0, 3);  // three passes over three disjoint site sets
	bad += run_case(4, 3, 6, true, true, 1, 3);  // three passes, one CTA per SM
	bad += run_case(3, 3, 4, true, false, 0, 3);
	std::printf(bad ? "FAIL\n" : "OK\n");
	return bad ? 1 : 0;
}
