// lpp_dblock_kernel.cuh -- two-pass BLOCK sweep over one spin species of a product basis (HubbardHelper.h:119-133,191-243:
// the spin-down hopping terms and the diagonal of x += H y):   x = beta x + alpha (D + 1 (x) T_dn) y.
//
// Pick two disjoint sets of sites F1, F2 with no hopping amplitude between them.  Pass 1 groups the down states by their
// occupation of F1: every hop that does not touch F1 stays inside its group ("block").  Pass 2 groups them by F2 and applies
// the hops that touch F1; none of those touches F2, so they stay inside the F2 blocks.  A half-tile = (block, 8 columns of the
// Ndn x Nup matrix) of y is staged in shared memory; no operand lies outside it, so the global traffic is y once and x
// read + write per pass.
//
// Shared-memory layout of a half-tile: a state owns 64 bytes (8 columns).  States are split in two classes by the parity of
// their occupation of a site set Z chosen as a maximum cut of the hopping graph (a sublattice on bipartite lattices); class 0
// lives in banks 0-15, class 1 in banks 16-31 of a 128-byte line.  A quarter-warp serves one class-0 and one class-1 state
// (4 lanes x 16 bytes each); every hop across the cut flips the class, so the two operands of a quarter-warp again sit in
// opposite bank halves: one conflict-free wavefront per 128 bytes by construction.
//
// Two half-tile buffers: the next half-tile is filled by cp.async while the current one is computed (the L2->SM fabric gives an
// SM about 45 bytes/clock, a third of what the LSU reads from shared memory, so an exposed fill costs 25 % of the sweep).
// Tables (2-byte entries, "+" and "-" operands in separate lists) arrive by one bulk copy (TMA) per tile on an mbarrier, double
// buffered as well.  A persistent grid takes tiles from a ticket counter in panel-major order, pass 2 of a panel `lag` panels
// behind pass 1, so a panel's x and y stay L2 resident between its two passes.
//
// This header holds the host-side plan builder, the kernel and the launcher; it is included by lpp_dblock.cu (engine) and by
// tools/proto_dblock.cu (stand-alone timing harness).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#define DB_THREADS 1024
#define DB_NW (DB_THREADS / 32)
#define DB_HCOLS 8                     // columns of a half-tile
#define DB_COLS 16                     // columns of a ticket (two half-tiles share the tables)
#define DB_ROW_NONE 0xffffffffu

struct DbBlock {
	uint32_t nstates;                  // real states
	uint32_t npos;                     // positions = 8 * nsteps (padding included)
	uint32_t nsteps;
	uint32_t blob_off, blob_len;       // in 16-byte units, inside the pass blob
	uint32_t fill_off;                 // first entry of the block in the pass fill list
	uint32_t off_words, off_dv2, off_info, off_tab;   // byte offsets inside the block's blob (meta at 0)
	uint32_t pad0, pad1;
};

struct DbHostPass {
	std::vector<DbBlock> blocks;
	std::vector<uint4> blob;           // per block: meta {row, address code}[npos] | words[npos] | dv2[npos] | info[nsteps] | table
	std::vector<uint2> fill;           // per block: {row, address code} of every real state
	uint32_t max_lines = 0, max_blob = 0, max_states = 0;
	double mean_hops = 0;
	uint64_t exec_slots = 0;           // state-slots executed (padding included)
	uint64_t conflicts = 0;            // quarter-warp slots whose two operands share a bank half
};

struct DbHostPlan {
	uint32_t f1 = 0, f2 = 0, zmask = 0;
	DbHostPass pass[2];
	size_t smem_bytes = 0;
	uint32_t tile_bytes = 0;           // bytes of one half-tile buffer: (max_lines + 1) * 128, line 0 is the zero line
	uint32_t blob_bytes = 0;           // bytes of one table buffer
	uint32_t fill_bytes = 0;           // bytes of one fill-list buffer
	int has_dv2 = 0;
	double tmag = 1.0;                 // the one hop magnitude (entries carry no amplitude)
};

// ---------------------------------------------------------------------------------------------------------------
// host: plan
// ---------------------------------------------------------------------------------------------------------------
template <class W>
static bool db_build_pass(const W* words, uint64_t n, const uint32_t* idx, const double* val, const uint32_t* cnt, const double* dv2,
                          uint64_t fmask_group, uint64_t f1mask, uint64_t zmask, int which, bool has_dv2, DbHostPass* out, std::string* err)
{
	std::vector<std::pair<uint64_t, uint32_t>> key(n);
	for (uint64_t s = 0; s < n; s++) key[s] = {(uint64_t)words[s] & fmask_group, (uint32_t)s};
	std::stable_sort(key.begin(), key.end(), [](const std::pair<uint64_t, uint32_t>& a, const std::pair<uint64_t, uint32_t>& b) { return a.first < b.first; });
	std::vector<uint32_t> acode_of(n, 0);
	auto cls_of = [&](uint32_t s) { return (uint32_t)(__builtin_popcountll((uint64_t)words[s] & zmask) & 1); };
	uint64_t total_hops = 0;
	size_t b0 = 0;
	while (b0 < n) {
		size_t b1 = b0;
		while (b1 < n && key[b1].first == key[b0].first) b1++;
		const uint32_t ns = (uint32_t)(b1 - b0);
		// hops of this pass, split by sign
		std::vector<std::vector<uint32_t>> plus(ns), minus(ns);
		for (uint32_t i = 0; i < ns; i++) {
			const uint32_t s = key[b0 + i].second;
			for (uint32_t k = 0; k < cnt[s]; k++) {
				const uint32_t t = idx[(uint64_t)k * n + s];
				const double v = val[(uint64_t)k * n + s];
				if (v == 0.0) continue;
				const uint64_t diff = (uint64_t)words[s] ^ (uint64_t)words[t];
				const bool touches_f1 = (diff & f1mask) != 0;
				if ((which == 0) == touches_f1) continue;
				if (((uint64_t)words[t] & fmask_group) != key[b0].first) { *err = "a hop leaves its block"; return false; }
				(v < 0 ? minus[i] : plus[i]).push_back(t);
				total_hops++;
			}
		}
		auto np2 = [](size_t c) { return (uint32_t)((c + 1) / 2); };
		// the two classes, each sorted by (pairs of "+" hops, pairs of "-" hops): zipped into quarter-warp pairs
		std::vector<uint32_t> cl[2];
		for (uint32_t i = 0; i < ns; i++) cl[cls_of(key[b0 + i].second)].push_back(i);
		for (int c = 0; c < 2; c++)
			std::stable_sort(cl[c].begin(), cl[c].end(), [&](uint32_t a, uint32_t b) {
				const uint32_t pa = np2(plus[a].size()), pb = np2(plus[b].size());
				if (pa != pb) return pa > pb;
				return np2(minus[a].size()) > np2(minus[b].size());
			});
		const uint32_t npairs = (uint32_t)std::max(cl[0].size(), cl[1].size());
		const uint32_t nsteps = (npairs + 3) / 4, npos = nsteps * 8;
		// storage: the i-th state of class c sits in line i + 1, half c  ->  address code 2 * (i + 1) + c  (x 64 bytes)
		for (int c = 0; c < 2; c++)
			for (uint32_t i = 0; i < cl[c].size(); i++) acode_of[key[b0 + cl[c][i]].second] = 2u * (i + 1u) + (uint32_t)c;
		if (2u * (npairs + 1u) + 1u > 0xffffu) { *err = "block too large for 2-byte table entries"; return false; }
		// state of position p = step * 8 + 2 * quarter + class
		auto state_at = [&](uint32_t p) -> int {
			const uint32_t pair = (p >> 3) * 4 + ((p & 7) >> 1), c = p & 1;
			return pair < cl[c].size() ? (int)cl[c][pair] : -1;
		};
		DbBlock blk;
		memset(&blk, 0, sizeof(blk));
		blk.nstates = ns;
		blk.npos = npos;
		blk.nsteps = nsteps;
		blk.blob_off = (uint32_t)out->blob.size();
		blk.fill_off = (uint32_t)out->fill.size();
		std::vector<uint32_t> bw;                                      // the block's blob as 32-bit words
		for (uint32_t p = 0; p < npos; p++) {
			const int i = state_at(p);
			if (i >= 0) {
				const uint32_t s = key[b0 + i].second;
				bw.push_back(s);
				bw.push_back(acode_of[s]);
				out->fill.push_back(make_uint2(s, acode_of[s]));
			} else {
				bw.push_back(DB_ROW_NONE);
				bw.push_back(0u);
			}
		}
		auto align16 = [&]() { while (bw.size() % 4) bw.push_back(0u); };
		align16();
		if (which == 0) {
			blk.off_words = (uint32_t)bw.size() * 4;
			for (uint32_t p = 0; p < npos; p++) { const int i = state_at(p); bw.push_back(i >= 0 ? (uint32_t)words[key[b0 + i].second] : 0u); }
			align16();
			if (has_dv2) {
				blk.off_dv2 = (uint32_t)bw.size() * 4;
				for (uint32_t p = 0; p < npos; p++) {
					const int i = state_at(p);
					const double d = i >= 0 ? dv2[key[b0 + i].second] : 0.0;
					unsigned long long bits;
					memcpy(&bits, &d, 8);
					bw.push_back((uint32_t)bits);
					bw.push_back((uint32_t)(bits >> 32));
				}
				align16();
			}
		}
		// table of a step, in 32-byte units: for each sign, quad rows ([8 states] x 4 entries of 2 bytes = 64 bytes) followed
		// by one pair row ([8 states] x 2 entries = 32 bytes) when the pair count is odd.  Entry = address code, 0 = zero line.
		std::vector<uint32_t> info(nsteps, 0u);
		std::vector<uint16_t> tab;
		for (uint32_t st = 0; st < nsteps; st++) {
			uint32_t pp = 0, pm = 0;
			for (uint32_t j = 0; j < 8; j++) {
				const int i = state_at(st * 8 + j);
				if (i >= 0) { pp = std::max(pp, np2(plus[i].size())); pm = std::max(pm, np2(minus[i].size())); }
			}
			if (pp > 63 || pm > 63 || tab.size() / 16 >= (1u << 20)) { *err = "block table too large"; return false; }
			info[st] = (uint32_t)(tab.size() / 16) | (pp << 20) | (pm << 26);
			for (int sgn = 0; sgn < 2; sgn++) {
				const uint32_t npair = sgn ? pm : pp;
				auto entry = [&](uint32_t j, uint32_t k) -> uint16_t {
					const int i = state_at(st * 8 + j);
					if (i < 0) return 0;
					const std::vector<uint32_t>& l = sgn ? minus[i] : plus[i];
					if (k >= l.size()) return 0;
					return (uint16_t)acode_of[l[k]];
				};
				for (uint32_t g = 0; g < npair / 2; g++)
					for (uint32_t j = 0; j < 8; j++)
						for (uint32_t e = 0; e < 4; e++) tab.push_back(entry(j, g * 4 + e));
				if (npair & 1)
					for (uint32_t j = 0; j < 8; j++)
						for (uint32_t e = 0; e < 2; e++) tab.push_back(entry(j, (npair - 1) * 2 + e));
				for (uint32_t k = 0; k < npair * 2; k++)
					for (uint32_t qd = 0; qd < 4; qd++) {
						const uint16_t e0 = entry(2 * qd, k), e1 = entry(2 * qd + 1, k);
						if (e0 > 1 && e1 > 1 && (e0 & 1) == (e1 & 1)) out->conflicts++;
					}
			}
			out->exec_slots += 16ull * (pp + pm);
		}
		blk.off_info = (uint32_t)bw.size() * 4;
		for (uint32_t st = 0; st < nsteps; st++) bw.push_back(info[st]);
		align16();
		blk.off_tab = (uint32_t)bw.size() * 4;
		for (size_t i = 0; i < tab.size(); i += 2) bw.push_back((uint32_t)tab[i] | ((uint32_t)(i + 1 < tab.size() ? tab[i + 1] : 0) << 16));
		align16();
		for (size_t i = 0; i < bw.size(); i += 4) out->blob.push_back(make_uint4(bw[i], bw[i + 1], bw[i + 2], bw[i + 3]));
		blk.blob_len = (uint32_t)out->blob.size() - blk.blob_off;
		out->blocks.push_back(blk);
		out->max_lines = std::max(out->max_lines, npairs);
		out->max_blob = std::max(out->max_blob, blk.blob_len);
		out->max_states = std::max(out->max_states, ns);
		b0 = b1;
	}
	out->mean_hops = (double)total_hops / (double)n;
	return true;
}

// words: one-spin basis (any order), nbits sites; ELL hop table (column-major idx/val, cnt) on the host.
// Returns false (with *err) when the two-pass block scheme does not apply (then the caller keeps the streaming sweep).
template <class W>
static bool db_build_host_plan(const W* words, uint64_t n, int nbits, const uint32_t* idx, const double* val, const uint32_t* cnt, int width,
                               const double* dv2, size_t max_smem, DbHostPlan* hp, std::string* err)
{
	(void)width;
	if (n == 0 || n >= (1ull << 24)) { *err = "basis size out of range"; return false; }
	if (nbits > 32) { *err = "more than 32 sites"; return false; }
	// site adjacency and the hop magnitude
	std::vector<uint64_t> adj(nbits, 0);
	double mag = 0;
	for (uint64_t s = 0; s < n; s++)
		for (uint32_t k = 0; k < cnt[s]; k++) {
			const double v = val[(uint64_t)k * n + s];
			if (v == 0.0) continue;
			if (mag == 0) mag = fabs(v);
			if (fabs(v) != mag) { *err = "more than one hop magnitude"; return false; }
			const uint64_t diff = (uint64_t)words[s] ^ (uint64_t)words[idx[(uint64_t)k * n + s]];
			if (__builtin_popcountll(diff) != 2) { *err = "a table entry is not a single hop"; return false; }
			const int i = __builtin_ctzll(diff), j = 63 - __builtin_clzll(diff);
			adj[i] |= 1ull << j;
			adj[j] |= 1ull << i;
		}
	if (mag == 0) { *err = "no hops"; return false; }
	bool has_dv2 = false;
	if (dv2)
		for (uint64_t s = 0; s < n; s++) has_dv2 = has_dv2 || dv2[s] != 0.0;
	const uint64_t all = (nbits == 64) ? ~0ull : ((1ull << nbits) - 1);
	// Z: greedy local search for a maximum cut of the hopping graph (exact on bipartite lattices)
	uint64_t Z = 0;
	{
		std::vector<int> col(nbits, -1);
		for (int r = 0; r < nbits; r++) {
			if (col[r] >= 0) continue;
			col[r] = 0;
			std::vector<int> queue(1, r);
			for (size_t h = 0; h < queue.size(); h++)
				for (int j = 0; j < nbits; j++)
					if (((adj[queue[h]] >> j) & 1) && col[j] < 0) { col[j] = col[queue[h]] ^ 1; queue.push_back(j); }
		}
		for (int i = 0; i < nbits; i++) if (col[i]) Z |= 1ull << i;
		bool moved = true;
		while (moved) {
			moved = false;
			for (int i = 0; i < nbits; i++) {
				const int same = __builtin_popcountll(adj[i] & (((Z >> i) & 1) ? Z : ~Z & all));
				const int other = __builtin_popcountll(adj[i]) - same;
				if (same > other) { Z ^= 1ull << i; moved = true; }
			}
		}
	}
	// smallest number of fixed sites whose largest block fits; F1 = highest sites possible
	for (int f = 1; f <= nbits / 2; f++) {
		std::vector<int> c(f);
		for (int i = 0; i < f; i++) c[i] = nbits - 1 - i;      // descending positions
		bool more = true;
		while (more) {
			uint64_t F1 = 0, nb = 0;
			for (int i = 0; i < f; i++) { F1 |= 1ull << c[i]; nb |= adj[c[i]]; }
			const uint64_t allowed = all & ~(F1 | nb);
			if (__builtin_popcountll(allowed) >= f) {
				uint64_t F2 = 0, a = allowed;
				for (int i = 0; i < f; i++) { const int b = 63 - __builtin_clzll(a); F2 |= 1ull << b; a &= ~(1ull << b); }
				// block sizes
				uint32_t mx = 0;
				for (int p = 0; p < 2; p++) {
					const uint64_t F = p ? F2 : F1;
					std::vector<uint64_t> keys(n);
					for (uint64_t s = 0; s < n; s++) keys[s] = (uint64_t)words[s] & F;
					std::sort(keys.begin(), keys.end());
					uint32_t run = 0;
					for (uint64_t s = 0; s < n; s++) {
						run = (s && keys[s] == keys[s - 1]) ? run + 1 : 1;
						mx = std::max(mx, run);
					}
				}
				if ((size_t)mx * 64 * 2 + 8 * 1024 <= max_smem) {           // two half-tile buffers; the exact check follows
					DbHostPlan cand;
					cand.f1 = (uint32_t)F1;
					cand.f2 = (uint32_t)F2;
					cand.zmask = (uint32_t)Z;
					cand.tmag = mag;
					cand.has_dv2 = has_dv2 ? 1 : 0;
					std::string e2;
					if (db_build_pass(words, n, idx, val, cnt, dv2, F1, F1, Z, 0, has_dv2, &cand.pass[0], &e2) &&
					    db_build_pass(words, n, idx, val, cnt, dv2, F2, F1, Z, 1, has_dv2, &cand.pass[1], &e2)) {
						const uint32_t ml = std::max(cand.pass[0].max_lines, cand.pass[1].max_lines);
						const uint32_t mb = std::max(cand.pass[0].max_blob, cand.pass[1].max_blob);
						const uint32_t ms = std::max(cand.pass[0].max_states, cand.pass[1].max_states);
						cand.tile_bytes = (ml + 1) * 128u;
						cand.blob_bytes = mb * 16u;
						cand.fill_bytes = ((ms * 8u) + 15u) & ~15u;
						cand.smem_bytes = 2 * ((size_t)cand.tile_bytes + cand.blob_bytes + cand.fill_bytes);
						if (cand.smem_bytes + 1024 <= max_smem) { *hp = std::move(cand); return true; }
					}
				}
				// all f-subsets give the same block sizes when the basis is every word of a fixed particle number: next f
				break;
			}
			// next combination (descending)
			int i = f - 1;
			while (i >= 0 && c[i] == f - 1 - i) i--;
			if (i < 0) more = false;
			else {
				c[i]--;
				for (int j = i + 1; j < f; j++) c[j] = c[j - 1] - 1;
			}
		}
	}
	*err = "no pair of separated site sets whose blocks fit shared memory";
	return false;
}

// ---------------------------------------------------------------------------------------------------------------
// device plan
// ---------------------------------------------------------------------------------------------------------------
struct DbDevPass {
	DbBlock* blocks = nullptr;
	uint4* blob = nullptr;
	uint2* fill = nullptr;
	uint32_t nblocks = 0;
};
struct DbDevPlan {
	DbDevPass pass[2];
	unsigned long long* ctrl = nullptr;   // [0] ticket counter, then uint32 done counters per panel
	uint32_t ctrl_panels = 0;
	size_t smem_bytes = 0;
	uint32_t tile_bytes = 0, blob_bytes = 0, fill_bytes = 0;
	int has_dv2 = 0;
	int lag = 12;                         // pass 2 runs this many panels behind pass 1
	double tmag = 1.0;
	long long* profile = nullptr;         // DB_PROFILE builds: 8 cycle counters per CTA
	bool attr_set = false;
};

struct DbArgs {
	double* x;
	const double* y;
	uint64_t pitch, ncols;
	double alpha, beta, U0, tmag;
	const uint32_t* w1;                // up word of every column (32-bit copy)
	const double* dv1;                 // up potential of every column
	double* dot_partials;              // optional: per pass-2 tile partial sums of y . x_new  [npanels * nblocks2]
};

static bool db_upload_plan(const DbHostPlan& hp, DbDevPlan* dp, std::string* err)
{
	for (int p = 0; p < 2; p++) {
		const DbHostPass& h = hp.pass[p];
		DbDevPass& d = dp->pass[p];
		d.nblocks = (uint32_t)h.blocks.size();
		if (cudaMalloc(&d.blocks, h.blocks.size() * sizeof(DbBlock)) != cudaSuccess || cudaMalloc(&d.blob, h.blob.size() * sizeof(uint4)) != cudaSuccess ||
		    cudaMalloc(&d.fill, h.fill.size() * sizeof(uint2)) != cudaSuccess) { *err = "cudaMalloc failed"; return false; }
		cudaMemcpy(d.blocks, h.blocks.data(), h.blocks.size() * sizeof(DbBlock), cudaMemcpyHostToDevice);
		cudaMemcpy(d.blob, h.blob.data(), h.blob.size() * sizeof(uint4), cudaMemcpyHostToDevice);
		cudaMemcpy(d.fill, h.fill.data(), h.fill.size() * sizeof(uint2), cudaMemcpyHostToDevice);
	}
	dp->smem_bytes = hp.smem_bytes;
	dp->tile_bytes = hp.tile_bytes;
	dp->blob_bytes = hp.blob_bytes;
	dp->fill_bytes = hp.fill_bytes;
	dp->has_dv2 = hp.has_dv2;
	dp->tmag = hp.tmag;
	if (cudaGetLastError() != cudaSuccess) { *err = "plan upload failed"; return false; }
	return true;
}

static void db_free_plan(DbDevPlan* dp)
{
	for (int p = 0; p < 2; p++) {
		cudaFree(dp->pass[p].blocks);
		cudaFree(dp->pass[p].blob);
		cudaFree(dp->pass[p].fill);
	}
	cudaFree(dp->ctrl);
	*dp = DbDevPlan();
}

// ---------------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------------
struct DbKernelArgs {
	DbArgs a;
	const DbBlock* blocks[2];
	const uint4* blob[2];
	const uint2* fill[2];
	uint32_t nb[2];
	uint32_t npanels, lag;
	uint32_t tile_bytes, blob_bytes, fill_bytes;   // shared memory: 2 half-tile buffers | 2 table buffers | 2 fill lists
	int has_dv2;
	long long* profile;
	unsigned long long* ticket;
	uint32_t* done1;
};

__device__ __forceinline__ void db_ld2(uint32_t addr, double& vx, double& vy)
{
	asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(addr));
}
// one quad row: 4 operands of this lane's state (2-byte address codes x 64 bytes) added into two accumulator pairs
__device__ __forceinline__ void db_quad(uint32_t ta, uint32_t lane_base, double& a0, double& a1, double& b0, double& b1)
{
	uint32_t w0, w1;
	asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(ta));
	double v0, v1, v2, v3, v4, v5, v6, v7;
	db_ld2(lane_base + ((w0 & 0xffffu) << 6), v0, v1);
	db_ld2(lane_base + ((w0 >> 16) << 6), v2, v3);
	db_ld2(lane_base + ((w1 & 0xffffu) << 6), v4, v5);
	db_ld2(lane_base + ((w1 >> 16) << 6), v6, v7);
	a0 += v0; a1 += v1;
	b0 += v2; b1 += v3;
	a0 += v4; a1 += v5;
	b0 += v6; b1 += v7;
}
__device__ __forceinline__ void db_pair(uint32_t ta, uint32_t lane_base, double& a0, double& a1, double& b0, double& b1)
{
	uint32_t w0;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(ta));
	double v0, v1, v2, v3;
	db_ld2(lane_base + ((w0 & 0xffffu) << 6), v0, v1);
	db_ld2(lane_base + ((w0 >> 16) << 6), v2, v3);
	a0 += v0; a1 += v1;
	b0 += v2; b1 += v3;
}

// ticket t -> (pass, panel, block): panel-major, pass 2 of a panel `L` panels behind its pass 1
__device__ __forceinline__ void db_decode(const DbKernelArgs& ka, unsigned long long t, uint32_t L, uint32_t nbt, uint32_t& pass, uint32_t& panel,
                                          uint32_t& blk)
{
	if (t < (unsigned long long)L * ka.nb[0]) {
		pass = 0; panel = (uint32_t)(t / ka.nb[0]); blk = (uint32_t)(t % ka.nb[0]);
	} else {
		const unsigned long long u = t - (unsigned long long)L * ka.nb[0];
		const unsigned long long mid = (unsigned long long)(ka.npanels - L) * nbt;
		if (u < mid) {
			const uint32_t i = (uint32_t)(u / nbt), r = (uint32_t)(u % nbt);
			if (r < ka.nb[0]) { pass = 0; panel = L + i; blk = r; }
			else { pass = 1; panel = i; blk = r - ka.nb[0]; }
		} else {
			const unsigned long long v = u - mid;
			pass = 1; panel = (ka.npanels - L) + (uint32_t)(v / ka.nb[1]); blk = (uint32_t)(v % ka.nb[1]);
		}
	}
}

__device__ __forceinline__ void db_mbar_wait(uint32_t bar, uint32_t phase)
{
	asm volatile("{\n.reg .pred p;\nDBW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DBD_%=;\nbra DBW_%=;\nDBD_%=:\n}" ::"r"(bar), "r"(phase) : "memory");
}

struct DbTileRef {
	uint32_t pass, panel, blk;
	bool valid;
};

template <bool DOT>
__global__ void __launch_bounds__(DB_THREADS, 1) k_dblock(const DbKernelArgs ka)
{
	extern __shared__ __align__(128) unsigned char db_smem[];
	const uint32_t smem_s = (uint32_t)__cvta_generic_to_shared(db_smem);
	const uint32_t tiles_sa = smem_s;                                            // 2 half-tile buffers
	const uint32_t blobs_sa = smem_s + 2u * ka.tile_bytes;                       // 2 table buffers
	const uint32_t fills_sa = blobs_sa + 2u * ka.blob_bytes;                     // 2 fill lists
	const unsigned char* blobs = db_smem + 2u * ka.tile_bytes;
	const uint2* fills = reinterpret_cast<const uint2*>(db_smem + 2u * ka.tile_bytes + 2u * ka.blob_bytes);
	__shared__ unsigned long long s_ticket[2];
	__shared__ DbBlock s_bd[2];
	__shared__ __align__(8) unsigned long long s_bar[2];
	__shared__ double s_red[DB_NW];
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int j = lane >> 2, c = lane & 3;                                       // state of the step, column pair
	const DbArgs& a = ka.a;
	const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&s_bar[0]);
	// line 0 of both half-tile buffers is the zero line (padding entries point at it)
	if (tid < 64) reinterpret_cast<float*>(db_smem + (tid >> 5) * ka.tile_bytes)[tid & 31] = 0.0f;
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s + 8u));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	const uint32_t nbt = ka.nb[0] + ka.nb[1];
	const unsigned long long total = (unsigned long long)ka.npanels * nbt;
	const uint32_t L = min(ka.lag, ka.npanels);
	const bool need_x1 = a.beta != 0.0;
	auto tile_of = [&](unsigned long long t) {
		DbTileRef T;
		T.valid = t < total;
		T.pass = T.panel = T.blk = 0;
		if (T.valid) db_decode(ka, t, L, nbt, T.pass, T.panel, T.blk);
		return T;
	};
	// descriptor + fill list of ticket T into buffer `buf` (cp.async, caller commits)
	auto stage_ticket = [&](const DbTileRef& T, uint32_t buf) {
		if (!T.valid) return;
		const DbBlock* bdp = ka.blocks[T.pass] + T.blk;
		if (tid == 0) s_bd[buf] = *bdp;
		const uint32_t ns = __ldg(&bdp->nstates), fo = __ldg(&bdp->fill_off);
		const uint2* __restrict__ fl = ka.fill[T.pass] + fo;
		for (uint32_t p = tid; p < ns; p += DB_THREADS)
			asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(fills_sa + buf * ka.fill_bytes + p * 8u), "l"(fl + p) : "memory");
	};
	// tables of the ticket in buffer `buf`: one bulk copy completing on the buffer's mbarrier
	auto load_blob = [&](const DbTileRef& T, uint32_t buf) {
		if (tid == 0 && T.valid) {
			const uint32_t bytes = s_bd[buf].blob_len * 16u;
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s + buf * 8u), "r"(bytes) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(blobs_sa + buf * ka.blob_bytes),
			             "l"(ka.blob[T.pass] + s_bd[buf].blob_off), "r"(bytes), "r"(bar_s + buf * 8u)
			             : "memory");
		}
	};
	// 64-byte chunks of half-tile `h` of ticket T (fill list in buffer fbuf) into half-tile buffer tbuf
	auto load_half = [&](const DbTileRef& T, uint32_t fbuf, uint32_t h, uint32_t tbuf) {
		if (!T.valid) return;
		const uint32_t ns = s_bd[fbuf].nstates;
		const uint64_t col = (uint64_t)T.panel * DB_COLS + h * DB_HCOLS + 2u * (tid & 3);
		const bool ok = col < a.ncols;
		const double* ycol = a.y + (ok ? col : 0);
		const uint32_t nbytes = ok ? 16u : 0u;
		const uint2* fl = fills + (size_t)fbuf * (ka.fill_bytes / 8u);
		const uint32_t dst0 = tiles_sa + tbuf * ka.tile_bytes + (uint32_t)(tid & 3) * 16u;
		for (uint32_t p = tid >> 2; p < ns; p += DB_THREADS / 4) {
			const uint2 f = fl[p];
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + f.y * 64u), "l"(ycol + (uint64_t)f.x * a.pitch), "r"(nbytes) : "memory");
		}
	};

	// prologue: two tickets, the first one staged and its first half-tile in flight
	if (tid == 0) { s_ticket[0] = atomicAdd(ka.ticket, 1ull); s_ticket[1] = atomicAdd(ka.ticket, 1ull); }
	__syncthreads();
	DbTileRef cur = tile_of(s_ticket[0]), nxt = tile_of(s_ticket[1]);
	stage_ticket(cur, 0);
	asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
	__syncthreads();
	load_blob(cur, 0);
	load_half(cur, 0, 0, 0);
	asm volatile("cp.async.commit_group;" ::: "memory");
	uint32_t bphase0 = 0u, bphase1 = 0u;
	uint32_t hcount = 0;                                                         // half-tiles done: buffer parity
#ifdef DB_PROFILE
	long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	long long tk = clock64();
#define DB_TICK(i_) do { if (tid == 0) { const long long now_ = clock64(); pf[i_] += now_ - tk; tk = now_; } } while (0)
#else
#define DB_TICK(i_) do { } while (0)
#endif
	for (uint32_t tno = 0; cur.valid; tno++) {
		const uint32_t tb = tno & 1u;                                            // table / fill-list / descriptor buffer of this ticket
		const uint32_t pass = cur.pass, panel = cur.panel, blk = cur.blk;
		double contrib = 0.0;
#pragma unroll 1
		for (uint32_t h = 0; h < 2; h++, hcount++) {
			const uint32_t hb = hcount & 1u;
			// ---- this half-tile (and, on the first half, the tables) has landed
			asm volatile("cp.async.wait_group 0;" ::: "memory");
			if (h == 0) {
				if (tb == 0) { db_mbar_wait(bar_s, bphase0); bphase0 ^= 1u; }
				else { db_mbar_wait(bar_s + 8u, bphase1); bphase1 ^= 1u; }
				if (pass == 1 && tid == 0) {
					// wait until every pass-1 tile of this panel has written its x
					const uint32_t want = ka.nb[0];
					uint32_t seen;
					do {
						asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ka.done1 + panel) : "memory");
						if (seen < want) __nanosleep(100);
					} while (seen < want);
				}
			}
			DB_TICK(0);
			__syncthreads();
			DB_TICK(1);
			// ---- loads for the next half-tile go out now and land while this one is computed
			if (h == 0) {
				load_half(cur, tb, 1, hb ^ 1u);
				stage_ticket(nxt, tb ^ 1u);                                      // descriptor + fill list of the next ticket
				if (tid == 0) s_ticket[tb] = atomicAdd(ka.ticket, 1ull);         // the ticket after next
			} else {
				load_blob(nxt, tb ^ 1u);
				load_half(nxt, tb ^ 1u, 0, hb ^ 1u);
			}
			asm volatile("cp.async.commit_group;" ::: "memory");

			// ---- compute: a warp takes 8 states per step (4 quarter-warps = 4 pairs of a class-0 and a class-1 state)
			const DbBlock& bd = s_bd[tb];
			const unsigned char* blob = blobs + (size_t)tb * ka.blob_bytes;
			const uint2* meta_s = reinterpret_cast<const uint2*>(blob);
			const uint32_t* words_s = reinterpret_cast<const uint32_t*>(blob + bd.off_words);
			const double* dv2_s = reinterpret_cast<const double*>(blob + bd.off_dv2);
			const uint32_t* info_s = reinterpret_cast<const uint32_t*>(blob + bd.off_info);
			const uint32_t tab_sa = blobs_sa + tb * ka.blob_bytes + bd.off_tab;
			const uint32_t nsteps = bd.nsteps;
			const uint64_t mycol = (uint64_t)panel * DB_COLS + h * DB_HCOLS + 2u * c;
			const bool colok = mycol < a.ncols;
			const uint32_t lane_base = tiles_sa + hb * ka.tile_bytes + (uint32_t)c * 16u;
			const bool read_x = pass == 1 || need_x1;
			uint32_t k1[2] = {0u, 0u};
			double dv1[2] = {0.0, 0.0};
			if (pass == 0 && colok) {
				k1[0] = __ldg(a.w1 + mycol); k1[1] = __ldg(a.w1 + mycol + 1);
				dv1[0] = __ldg(a.dv1 + mycol); dv1[1] = __ldg(a.dv1 + mycol + 1);
			}
			for (uint32_t st = wid; st < nsteps; st += DB_NW) {
				const uint32_t pos = st * 8u + j;
				const uint32_t info = info_s[st];
				const uint2 m = meta_s[pos];
				const bool valid = colok && m.x != DB_ROW_NONE;
				double* xp = a.x + (uint64_t)m.x * a.pitch + mycol;
				double2 xo = make_double2(0.0, 0.0);
				if (valid && read_x) xo = __ldcg(reinterpret_cast<const double2*>(xp));
				double yo0 = 0.0, yo1 = 0.0;
				if (pass == 0 || DOT) db_ld2(lane_base + m.y * 64u, yo0, yo1);
				double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0, c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
				uint32_t ta = tab_sa + (info & 0x000fffffu) * 32u;
				const uint32_t pp = (info >> 20) & 63u, pm = info >> 26;
#pragma unroll 1
				for (uint32_t g = 0; g < (pp >> 1); g++, ta += 64u) db_quad(ta + (uint32_t)j * 8u, lane_base, a0, a1, b0, b1);
				if (pp & 1u) { db_pair(ta + (uint32_t)j * 4u, lane_base, a0, a1, b0, b1); ta += 32u; }
#pragma unroll 1
				for (uint32_t g = 0; g < (pm >> 1); g++, ta += 64u) db_quad(ta + (uint32_t)j * 8u, lane_base, c0, c1, d0, d1);
				if (pm & 1u) db_pair(ta + (uint32_t)j * 4u, lane_base, c0, c1, d0, d1);
				double h0 = a.tmag * ((a0 + b0) - (c0 + d0)), h1 = a.tmag * ((a1 + b1) - (c1 + d1));
				double xn0, xn1;
				if (pass == 0) {
					const uint32_t wd = words_s[pos];
					const double dv2 = ka.has_dv2 ? dv2_s[pos] : 0.0;
					h0 += (a.U0 * (double)__popc(k1[0] & wd) + dv1[0] + dv2) * yo0;
					h1 += (a.U0 * (double)__popc(k1[1] & wd) + dv1[1] + dv2) * yo1;
					xn0 = a.alpha * h0;
					xn1 = a.alpha * h1;
					if (need_x1) { xn0 += a.beta * xo.x; xn1 += a.beta * xo.y; }
				} else {
					xn0 = xo.x + a.alpha * h0;
					xn1 = xo.y + a.alpha * h1;
				}
				if (valid) {
					__stcg(reinterpret_cast<double2*>(xp), make_double2(xn0, xn1));
					if (DOT && pass == 1) contrib += yo0 * xn0 + yo1 * xn1;
				}
			}
			DB_TICK(2 + pass);
			__syncthreads();                               // all x of this half-tile are written, its buffer may be overwritten
			DB_TICK(4);
		}
		if (pass == 0) {
			if (tid == 0) {
				__threadfence();
				atomicAdd(ka.done1 + panel, 1u);
			}
		} else if (DOT) {
			// deterministic per-tile partial sum
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) contrib += __shfl_down_sync(0xffffffffu, contrib, o);
			if (lane == 0) s_red[wid] = contrib;
			__syncthreads();
			if (wid == 0) {
				double v = lane < DB_NW ? s_red[lane] : 0.0;
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
				if (lane == 0) a.dot_partials[(uint64_t)panel * ka.nb[1] + blk] = v;
			}
		}
		cur = nxt;
		nxt = tile_of(s_ticket[tb]);
	}
	asm volatile("cp.async.wait_group 0;" ::: "memory");
#ifdef DB_PROFILE
	if (tid == 0 && ka.profile)
		for (int i = 0; i < 8; i++) ka.profile[blockIdx.x * 8 + i] = pf[i];
#endif
}

// returns 0 on success
static int db_launch(DbDevPlan& dp, const DbArgs& a, int nsm, cudaStream_t s)
{
	const uint32_t npanels = (uint32_t)((a.ncols + DB_COLS - 1) / DB_COLS);
	if (npanels == 0) return 0;
	if (dp.ctrl_panels < npanels) {
		cudaFree(dp.ctrl);
		dp.ctrl = nullptr;
		if (cudaMalloc(&dp.ctrl, 8 + (size_t)npanels * 4) != cudaSuccess) return -1;
		dp.ctrl_panels = npanels;
	}
	if (!dp.attr_set) {
		if (cudaFuncSetAttribute(k_dblock<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dp.smem_bytes) != cudaSuccess) return -1;
		if (cudaFuncSetAttribute(k_dblock<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dp.smem_bytes) != cudaSuccess) return -1;
		dp.attr_set = true;
	}
	if (cudaMemsetAsync(dp.ctrl, 0, 8 + (size_t)npanels * 4, s) != cudaSuccess) return -1;
	DbKernelArgs ka;
	ka.a = a;
	for (int p = 0; p < 2; p++) {
		ka.blocks[p] = dp.pass[p].blocks;
		ka.blob[p] = dp.pass[p].blob;
		ka.fill[p] = dp.pass[p].fill;
		ka.nb[p] = dp.pass[p].nblocks;
	}
	ka.npanels = npanels;
	ka.lag = (uint32_t)std::max(dp.lag, 1);
	ka.tile_bytes = dp.tile_bytes;
	ka.blob_bytes = dp.blob_bytes;
	ka.fill_bytes = dp.fill_bytes;
	ka.has_dv2 = dp.has_dv2;
	ka.profile = dp.profile;
	ka.ticket = dp.ctrl;
	ka.done1 = reinterpret_cast<uint32_t*>(dp.ctrl + 1);
	const unsigned long long total = (unsigned long long)npanels * (ka.nb[0] + ka.nb[1]);
	const unsigned grid = (unsigned)std::min<unsigned long long>((unsigned long long)nsm, total);
	if (a.dot_partials) k_dblock<true><<<grid, DB_THREADS, dp.smem_bytes, s>>>(ka);
	else k_dblock<false><<<grid, DB_THREADS, dp.smem_bytes, s>>>(ka);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
