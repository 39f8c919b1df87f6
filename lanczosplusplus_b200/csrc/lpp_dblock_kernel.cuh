// lpp_dblock_kernel.cuh -- multi-pass BLOCK sweep over one spin species of a product basis (HubbardHelper.h:119-133,191-243:
// the spin-down hopping terms and the diagonal of x += H y):   x = beta x + alpha (D + 1 (x) T_dn) y.
//
// Pick K sets of sites F_1 .. F_K such that every bond misses at least one of them.  Pass k groups the down states by their
// occupation of F_k ("blocks") and applies the hops that touch F_1 .. F_{k-1} but not F_k: such a hop cannot leave its block.
// Two sets do when no bond joins them; three pairwise disjoint sets always do (a bond has two ends; built and tested, but
// slower than the streaming sweep, so only on request).  A tile = (block, 16
// columns of the Ndn x Nup matrix) of y is staged in shared memory; each hop operand is a conflict-free 16-byte shared-memory
// load (the 8 lanes of a state read one 128-byte line).  No operand lies outside the tile: the global traffic is y once and x
// read + write per pass.  A persistent grid takes tiles from a ticket counter in panel-major order, pass k of a panel LAG
// panels behind pass k-1, so a panel's x and y stay L2 resident between its passes (DRAM sees 24 bytes per element).
//
// The plan prefers blocks small enough for TWO resident CTAs (512 threads each) per SM, so that one CTA's tile fill and
// barriers hide behind the other's gathers; when only the one-CTA layout fits, the CTA has 1024 threads.
//
// This header holds the host-side plan builder, the kernel and the launcher; it is included by lpp_dblock.cu (engine), by
// tools/proto_dblock.cu (stand-alone timing harness) and by tests/dblock_plan_check.cu (CPU walk of the tables).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "lpp_smem_attr.cuh"

#define DB_MAX_PASS 3
#define DB_COLS 16
#define DB_LINE 128u                   // bytes of one state's 16 columns
#define DB_ROW_NONE 0xffffffffu

struct DbBlock {
	uint32_t nstates;                  // real states
	uint32_t nsteps;                   // groups of 4 positions
	uint32_t blob_off, blob_len;       // in 16-byte units, inside the pass blob
	uint32_t rows_off;                 // first entry of the block in the pass row list
	uint32_t pad;
};

struct DbHostPass {
	std::vector<DbBlock> blocks;
	std::vector<uint4> blob;           // per block: meta[npos] | stepinfo[ceil(nsteps/4)] | table[groups][4 states]
	std::vector<uint32_t> rows;        // per block: row (down state) of every position
	uint32_t max_pos = 0, max_blob = 0;
	double mean_hops = 0;
	uint64_t exec_slots = 0;          // state-slots executed (padding included)
};

struct DbHostPlan {
	int npass = 0;
	uint32_t fmask[DB_MAX_PASS] = {0, 0, 0};
	DbHostPass pass[DB_MAX_PASS];
	size_t smem_bytes = 0;             // tile + largest table blob (the two row lists come on top: max_pos * 8)
	uint32_t tile_bytes = 0;           // (max_pos + 1) * 128: slot 0 is the zero line
	uint32_t max_pos = 0;
	int ctas_per_sm = 1, threads = 1024;
	double tmag = 1.0;                 // the one hop magnitude (entries carry signs only)
};

// ---------------------------------------------------------------------------------------------------------------
// host: plan
// ---------------------------------------------------------------------------------------------------------------
template <class W>
static bool db_build_pass(const W* words, uint64_t n, const uint32_t* idx, const double* val, const uint32_t* cnt, const double* dv2,
                          const uint64_t* masks, int npass, int which, int nwarps, DbHostPass* out, std::string* err)
{
	const uint64_t fmask_group = masks[which];
	// group states by their occupation of the fixed sites
	std::vector<std::pair<uint64_t, uint32_t>> key(n);
	for (uint64_t s = 0; s < n; s++) key[s] = {(uint64_t)words[s] & fmask_group, (uint32_t)s};
	std::stable_sort(key.begin(), key.end(), [](const std::pair<uint64_t, uint32_t>& a, const std::pair<uint64_t, uint32_t>& b) { return a.first < b.first; });
	std::vector<uint32_t> pos_of(n, 0);
	uint64_t total_hops = 0;
	size_t b0 = 0;
	while (b0 < n) {
		size_t b1 = b0;
		while (b1 < n && key[b1].first == key[b0].first) b1++;
		const uint32_t ns = (uint32_t)(b1 - b0);
		// hops of this pass per state
		std::vector<std::vector<std::pair<uint32_t, double>>> hl(ns);
		for (uint32_t i = 0; i < ns; i++) {
			const uint32_t s = key[b0 + i].second;
			for (uint32_t k = 0; k < cnt[s]; k++) {
				const uint32_t t = idx[(uint64_t)k * n + s];
				const double v = val[(uint64_t)k * n + s];
				if (v == 0.0) continue;
				const uint64_t diff = (uint64_t)words[s] ^ (uint64_t)words[t];
				int owner = 0;                                       // the first pass whose fixed sites the hop does not touch
				while (owner < npass && (diff & masks[owner]) != 0) owner++;
				if (owner == npass) { *err = "a hop touches every set of fixed sites"; return false; }
				if (owner != which) continue;
				if (((uint64_t)words[t] & fmask_group) != key[b0].first) { *err = "a hop leaves its block"; return false; }
				hl[i].push_back({t, v});
			}
		}
		// every state's hops split by sign: the kernel adds the "+" operands and subtracts the "-" ones, so a hop costs an
		// address add, the 16-byte load and two DADDs.  States are sorted by (groups of 4 "+" hops, groups of 4 "-" hops) so
		// the 4 states of a step execute (nearly) no padding.
		std::vector<std::vector<uint32_t>> plus(ns), minus(ns);
		for (uint32_t i = 0; i < ns; i++)
			for (auto& h : hl[i]) (h.second < 0 ? minus[i] : plus[i]).push_back(h.first);
		auto np2 = [](size_t c) { return (uint32_t)((c + 1) / 2); };
		std::vector<uint32_t> order(ns);
		for (uint32_t i = 0; i < ns; i++) order[i] = i;
		std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
			const uint32_t pa = np2(plus[a].size()), pb = np2(plus[b].size());
			if (pa != pb) return pa > pb;
			return np2(minus[a].size()) > np2(minus[b].size());
		});
		const uint32_t npos = (ns + 3) & ~3u, nsteps = npos / 4;
		// warp w walks the steps w, w + nwarps, ...: spread the steps over the warps by longest-processing-time-first so that they
		// reach the end of the tile together (in sorted order warp 0 would get the longest step of every round).  A partial last
		// step keeps its place, so that the padding positions stay at the end of the tile.
		if (nwarps > 1 && nsteps > 1) {
			auto step_cost = [&](uint32_t st) {
				uint32_t pp = 0, pm = 0;
				for (uint32_t p = st * 4; p < st * 4 + 4 && p < ns; p++) { pp = std::max(pp, np2(plus[order[p]].size())); pm = std::max(pm, np2(minus[order[p]].size())); }
				return 3u * (pp + pm) + 4u;
			};
			std::vector<uint32_t> cap(nwarps, 0), load(nwarps, 0), taken(nwarps, 0);
			for (uint32_t st = 0; st < nsteps; st++) cap[st % nwarps]++;
			const bool pinned = (ns & 3u) != 0;
			const uint32_t nfree = pinned ? nsteps - 1 : nsteps;
			if (pinned) { cap[(nsteps - 1) % nwarps]--; load[(nsteps - 1) % nwarps] += step_cost(nsteps - 1); }
			std::vector<uint32_t> sorted_steps(nfree);
			for (uint32_t st = 0; st < nfree; st++) sorted_steps[st] = st;
			std::stable_sort(sorted_steps.begin(), sorted_steps.end(), [&](uint32_t a, uint32_t b) { return step_cost(a) > step_cost(b); });
			std::vector<uint32_t> order2(order);
			for (uint32_t old_st : sorted_steps) {
				uint32_t best = nwarps;
				for (uint32_t w = 0; w < (uint32_t)nwarps; w++)
					if (taken[w] < cap[w] && (best == (uint32_t)nwarps || load[w] < load[best])) best = w;
				const uint32_t new_st = best + taken[best] * (uint32_t)nwarps;
				taken[best]++;
				load[best] += step_cost(old_st);
				for (int q = 0; q < 4; q++) order2[new_st * 4 + q] = order[old_st * 4 + q];
			}
			order.swap(order2);
		}
		for (uint32_t p = 0; p < ns; p++) pos_of[key[b0 + order[p]].second] = p;
		DbBlock blk;
		blk.nstates = ns;
		blk.nsteps = nsteps;
		blk.blob_off = (uint32_t)out->blob.size();
		blk.rows_off = (uint32_t)out->rows.size();
		blk.pad = 0;
		for (uint32_t p = 0; p < npos; p++) {
			uint4 m;
			if (p < ns) {
				const uint32_t s = key[b0 + order[p]].second;
				m.x = s;
				m.y = (uint32_t)words[s];
				const double d = dv2 ? dv2[s] : 0.0;
				unsigned long long bits;
				memcpy(&bits, &d, 8);
				m.z = (uint32_t)bits;
				m.w = (uint32_t)(bits >> 32);
				out->rows.push_back(s);
			} else {
				m.x = DB_ROW_NONE; m.y = 0; m.z = 0; m.w = 0;
				out->rows.push_back(DB_ROW_NONE);
			}
			out->blob.push_back(m);
		}
		// table of a step, in 32-byte units: for each sign, quad rows ([4 states] x 4 entries, 2 units each) followed by one
		// pair row ([4 states] x 2 entries) when the pair count is odd.  Entry = byte offset of the source line, 0 = zero line.
		std::vector<uint32_t> info(((size_t)nsteps + 3) & ~(size_t)3, 0u);
		std::vector<uint32_t> tab;                                   // 8 words per unit
		for (uint32_t st = 0; st < nsteps; st++) {
			uint32_t pp = 0, pm = 0;
			for (int q = 0; q < 4; q++) {
				const uint32_t p = st * 4 + q;
				if (p < ns) { pp = std::max(pp, np2(plus[order[p]].size())); pm = std::max(pm, np2(minus[order[p]].size())); }
			}
			if (pp > 63 || pm > 63 || tab.size() / 8 >= (1u << 20)) { *err = "block table too large"; return false; }
			info[st] = (uint32_t)(tab.size() / 8) | (pp << 20) | (pm << 26);
			for (int sgn = 0; sgn < 2; sgn++) {
				const uint32_t npair = sgn ? pm : pp;
				auto entry = [&](int q, uint32_t k) -> uint32_t {
					const uint32_t p = st * 4 + q;
					if (p >= ns) return 0u;
					const std::vector<uint32_t>& l = sgn ? minus[order[p]] : plus[order[p]];
					if (k >= l.size()) return 0u;
					total_hops++;
					return (pos_of[l[k]] + 1u) * DB_LINE;
				};
				for (uint32_t g = 0; g < npair / 2; g++)
					for (int q = 0; q < 4; q++)
						for (int j = 0; j < 4; j++) tab.push_back(entry(q, g * 4 + j));
				if (npair & 1)
					for (int q = 0; q < 4; q++)
						for (int j = 0; j < 2; j++) tab.push_back(entry(q, (npair - 1) * 2 + j));
			}
			out->exec_slots += 8ull * (pp + pm);
		}
		while (tab.size() % 4) tab.push_back(0u);
		for (size_t i = 0; i < info.size(); i += 4) out->blob.push_back(make_uint4(info[i], info[i + 1], info[i + 2], info[i + 3]));
		for (size_t i = 0; i < tab.size(); i += 4) out->blob.push_back(make_uint4(tab[i], tab[i + 1], tab[i + 2], tab[i + 3]));
		blk.blob_len = (uint32_t)out->blob.size() - blk.blob_off;
		out->blocks.push_back(blk);
		out->max_pos = std::max(out->max_pos, npos);
		out->max_blob = std::max(out->max_blob, blk.blob_len);
		b0 = b1;
	}
	out->mean_hops = (double)total_hops / (double)n;
	return true;
}

// largest group of states with equal occupation of the sites in F
template <class W>
static uint32_t db_max_block(const W* words, uint64_t n, uint64_t F)
{
	std::vector<uint64_t> keys(n);
	for (uint64_t s = 0; s < n; s++) keys[s] = (uint64_t)words[s] & F;
	std::sort(keys.begin(), keys.end());
	uint32_t run = 0, mx = 0;
	for (uint64_t s = 0; s < n; s++) {
		run = (s && keys[s] == keys[s - 1]) ? run + 1 : 1;
		mx = std::max(mx, run);
	}
	return mx;
}

// site sets for `npass` passes with f fixed sites each.  F_1 takes the highest sites (its blocks are then runs of rows).
// Two passes: F_2 = the highest f sites that no bond joins to F_1 (tried over all F_1 in descending order).
// Three passes: the next f sites and the f after them (fewer if the lattice runs out): pairwise disjoint is all it takes.
static bool db_pick_sets(int nbits, const std::vector<uint64_t>& adj, int npass, int f, uint64_t* masks)
{
	const uint64_t all = (nbits == 64) ? ~0ull : ((1ull << nbits) - 1);
	if (npass == 3) {
		if (2 * f >= nbits) return false;
		const int f3 = std::min(f, nbits - 2 * f);
		masks[0] = (((1ull << f) - 1) << (nbits - f)) & all;
		masks[1] = (((1ull << f) - 1) << (nbits - 2 * f)) & all;
		masks[2] = (((1ull << f3) - 1) << (nbits - 2 * f - f3)) & all;
		return true;
	}
	std::vector<int> c(f);
	for (int i = 0; i < f; i++) c[i] = nbits - 1 - i;          // f-subsets in descending order of their mask
	for (;;) {
		uint64_t F1 = 0, nb = 0;
		for (int i = 0; i < f; i++) { F1 |= 1ull << c[i]; nb |= adj[c[i]]; }
		uint64_t a = all & ~(F1 | nb);
		if (__builtin_popcountll(a) >= f) {
			uint64_t F2 = 0;
			for (int i = 0; i < f; i++) { const int b = 63 - __builtin_clzll(a); F2 |= 1ull << b; a &= ~(1ull << b); }
			masks[0] = F1;
			masks[1] = F2;
			return true;
		}
		int i = f - 1;
		while (i >= 0 && c[i] == f - 1 - i) i--;
		if (i < 0) return false;
		c[i]--;
		for (int j = i + 1; j < f; j++) c[j] = c[j - 1] - 1;
	}
}

// words: one-spin basis (any order), nbits sites; ELL hop table (column-major idx/val, cnt) on the host.
// smem_block = the opt-in shared-memory limit of one CTA, smem_sm = shared memory of one SM.  layout: 0 = best (two CTAs per
// SM when some plan fits, else one), 1 = one CTA per SM only (the round-2 first version; kept for A/B timing).  passes: 0 or 2 = two
// passes; 3 = three passes, on request only: a pass costs about 0.5 ms on config 3 whatever its hop count (y in, x in and out), and
// three of them (2.52 ms) lose to the streaming sweep (2.09 ms), so a lattice without two separated site sets keeps that sweep.
// Returns false (with *err) when the block scheme does not apply (then the caller keeps the streaming sweep).
template <class W>
static bool db_build_host_plan(const W* words, uint64_t n, int nbits, const uint32_t* idx, const double* val, const uint32_t* cnt, int width,
                               const double* dv2, size_t smem_block, size_t smem_sm, int layout, int passes, DbHostPlan* hp, std::string* err)
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
	hp->tmag = mag;
	const size_t fixed = 2048;                                       // static shared memory of the kernel + the 1 KB the system keeps per CTA
	for (int cps = (layout == 1 ? 1 : 2); cps >= 1; cps--) {
		const size_t per_cta = std::min(smem_block, smem_sm / (size_t)cps);
		if (per_cta <= fixed + 4096) continue;
		const size_t budget = per_cta - fixed;
		for (int npass = (passes ? passes : 2); npass <= (passes ? passes : 2); npass++)
			// the smallest number of fixed sites whose largest block fits
			for (int f = 1; f <= nbits / 2; f++) {
				uint64_t masks[DB_MAX_PASS] = {0, 0, 0};
				if (!db_pick_sets(nbits, adj, npass, f, masks)) continue;
				uint32_t mx = 0;
				for (int k = 0; k < npass; k++) mx = std::max(mx, db_max_block(words, n, masks[k]));
				const size_t tile = ((size_t)((mx + 3) & ~3u) + 1) * DB_LINE;
				if (tile + 4096 > budget) continue;                  // room for the tables is checked exactly below
				DbHostPlan cand;
				cand.npass = npass;
				cand.tmag = mag;
				cand.ctas_per_sm = cps;
				cand.threads = cps == 2 ? 512 : 1024;
				std::string e2;
				bool ok = true;
				uint32_t mp = 0, mb = 0;
				for (int k = 0; k < npass && ok; k++) {
					cand.fmask[k] = (uint32_t)masks[k];
					ok = db_build_pass(words, n, idx, val, cnt, dv2, masks, npass, k, cand.threads / 32, &cand.pass[k], &e2);
					mp = std::max(mp, cand.pass[k].max_pos);
					mb = std::max(mb, cand.pass[k].max_blob);
				}
				if (!ok) continue;
				cand.max_pos = mp;
				cand.tile_bytes = (mp + 1) * DB_LINE;
				cand.smem_bytes = (size_t)cand.tile_bytes + (size_t)mb * 16;
				if (cand.smem_bytes + (size_t)mp * 8 > budget) continue;
				*hp = std::move(cand);
				return true;
			}
	}
	*err = "no sets of fixed sites whose blocks fit shared memory";
	return false;
}

// ---------------------------------------------------------------------------------------------------------------
// device plan
// ---------------------------------------------------------------------------------------------------------------
struct DbDevPass {
	DbBlock* blocks = nullptr;
	uint4* blob = nullptr;
	uint32_t* rows = nullptr;
	uint32_t nblocks = 0;
};
struct DbDevPlan {
	int npass = 0;
	DbDevPass pass[DB_MAX_PASS];
	unsigned long long* ctrl = nullptr;   // [0] ticket counter, then uint32 done counters [npass - 1][npanels]
	uint32_t ctrl_panels = 0;
	size_t smem_bytes = 0;
	uint32_t tile_bytes = 0;
	uint32_t max_pos = 0;
	int ctas_per_sm = 1, threads = 1024;
	int lag = 8;                          // pass k of a panel runs this many panels behind pass k-1
	double tmag = 1.0;
	long long* profile = nullptr;         // DB_PROFILE builds: 8 cycle counters per CTA
	bool attr_set = false;
};

struct DbArgs {
	double* x;
	const double* y;
	uint64_t pitch, ncols;
	double alpha, beta, U0, tmag;
	const double* alpha_dev;           // optional: alpha / beta in device memory (device-resident Lanczos scalars)
	const double* beta_dev;
	const uint32_t* w1;                // up word of every column (32-bit copy)
	const double* dv1;                 // up potential of every column
	double* dot_partials;              // optional: per last-pass tile partial sums of y . x_new  [npanels * nblocks of the last pass]
};

static bool db_upload_plan(const DbHostPlan& hp, DbDevPlan* dp, std::string* err)
{
	dp->npass = hp.npass;
	for (int p = 0; p < hp.npass; p++) {
		const DbHostPass& h = hp.pass[p];
		DbDevPass& d = dp->pass[p];
		d.nblocks = (uint32_t)h.blocks.size();
		if (cudaMalloc(&d.blocks, h.blocks.size() * sizeof(DbBlock)) != cudaSuccess || cudaMalloc(&d.blob, h.blob.size() * sizeof(uint4)) != cudaSuccess ||
		    cudaMalloc(&d.rows, h.rows.size() * sizeof(uint32_t)) != cudaSuccess) { *err = "cudaMalloc failed"; return false; }
		cudaMemcpy(d.blocks, h.blocks.data(), h.blocks.size() * sizeof(DbBlock), cudaMemcpyHostToDevice);
		cudaMemcpy(d.blob, h.blob.data(), h.blob.size() * sizeof(uint4), cudaMemcpyHostToDevice);
		cudaMemcpy(d.rows, h.rows.data(), h.rows.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
	}
	dp->smem_bytes = hp.smem_bytes;
	dp->tile_bytes = hp.tile_bytes;
	dp->max_pos = hp.max_pos;
	dp->ctas_per_sm = hp.ctas_per_sm;
	dp->threads = hp.threads;
	dp->tmag = hp.tmag;
	if (cudaGetLastError() != cudaSuccess) { *err = "plan upload failed"; return false; }
	return true;
}

static void db_free_plan(DbDevPlan* dp)
{
	for (int p = 0; p < DB_MAX_PASS; p++) {
		cudaFree(dp->pass[p].blocks);
		cudaFree(dp->pass[p].blob);
		cudaFree(dp->pass[p].rows);
	}
	cudaFree(dp->ctrl);
	*dp = DbDevPlan();
}

// ---------------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------------
#ifndef DB_SKIP_PADDING
#define DB_SKIP_PADDING 0                // 1: predicate padded operands off (measured 1.92 ms against 1.88 ms without)
#endif
#ifndef DB_RED_LATER_PASSES
#define DB_RED_LATER_PASSES 1           // the passes after the first add into x with red.global.add.f64 instead of load + store (1.741 -> 1.711 ms); 0: load + store
#endif
#ifndef DB_FILL_MODE
#define DB_FILL_MODE 0                  // 0: 16-byte cp.async per thread; 2: one 128-byte bulk copy (TMA) per state on the tile's mbarrier
#endif

struct DbKernelArgs {
	DbArgs a;
	const DbBlock* blocks[DB_MAX_PASS];
	const uint4* blob[DB_MAX_PASS];
	const uint32_t* rows[DB_MAX_PASS];
	uint32_t nb[DB_MAX_PASS];
	uint32_t npass, npanels, lag;
	uint32_t tile_bytes, blob_bytes, max_pos;   // shared-memory layout: tile | tables | 2 row lists of max_pos words
	long long* profile;
	unsigned long long* ticket;
	uint32_t* done;                             // [npass - 1][npanels]: tiles of pass k that have written their x
};

__device__ __forceinline__ void db_ld2(uint32_t addr, double& vx, double& vy)
{
	asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(addr));
}
// operand of a table entry: offset 0 is padding (the state has fewer operands than the longest of its step).  The 8 lanes of
// a state are one quarter-warp, i.e. one shared-memory wavefront of the 16-byte load: a predicated-off state costs none.
__device__ __forceinline__ void db_ldop(uint32_t e, uint32_t lane_off, double& vx, double& vy)
{
#if DB_SKIP_PADDING
	asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\nmov.f64 %0, 0d0000000000000000;\nmov.f64 %1, 0d0000000000000000;\n"
	             "@p ld.shared.v2.f64 {%0, %1}, [%3];\n}"
	             : "=&d"(vx), "=&d"(vy)
	             : "r"(e), "r"(lane_off + e));
#else
	db_ld2(lane_off + e, vx, vy);
#endif
}
// one quad row: 4 operands of this lane's state added into two accumulator pairs
__device__ __forceinline__ void db_quad(uint32_t ta, uint32_t lane_off, double& a0, double& a1, double& b0, double& b1)
{
	uint32_t e0, e1, e2, e3;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e0), "=r"(e1), "=r"(e2), "=r"(e3) : "r"(ta));
	double v0, v1, v2, v3, v4, v5, v6, v7;
	db_ldop(e0, lane_off, v0, v1);
	db_ldop(e1, lane_off, v2, v3);
	db_ldop(e2, lane_off, v4, v5);
	db_ldop(e3, lane_off, v6, v7);
	a0 += v0; a1 += v1;
	b0 += v2; b1 += v3;
	a0 += v4; a1 += v5;
	b0 += v6; b1 += v7;
}
__device__ __forceinline__ void db_pair(uint32_t ta, uint32_t lane_off, double& a0, double& a1, double& b0, double& b1)
{
	uint32_t e0, e1;
	asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e0), "=r"(e1) : "r"(ta));
	double v0, v1, v2, v3;
	db_ldop(e0, lane_off, v0, v1);
	db_ldop(e1, lane_off, v2, v3);
	a0 += v0; a1 += v1;
	b0 += v2; b1 += v3;
}

// ticket t -> (pass, panel, block).  Time step ts = t / nbt holds the tiles of pass 0 of panel ts, pass 1 of panel ts - L,
// pass 2 of panel ts - 2 L, in that order; slots whose panel is out of range are empty (the first and last (npass-1) L steps).
__device__ __forceinline__ bool db_decode(const DbKernelArgs& ka, unsigned long long t, uint32_t L, uint32_t nbt, uint32_t& pass, uint32_t& panel,
                                          uint32_t& blk)
{
	const unsigned long long ts = t / nbt;
	uint32_t r = (uint32_t)(t % nbt), k = 0;
	while (k + 1 < ka.npass && r >= ka.nb[k]) { r -= ka.nb[k]; k++; }
	pass = k;
	blk = r;
	const unsigned long long back = (unsigned long long)k * L;
	if (ts < back || ts - back >= ka.npanels) return false;
	panel = (uint32_t)(ts - back);
	return true;
}

__device__ __forceinline__ void db_mbar_wait(uint32_t bar, uint32_t phase)
{
	asm volatile("{\n.reg .pred p;\nDBW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DBD_%=;\nbra DBW_%=;\nDBD_%=:\n}" ::"r"(bar), "r"(phase) : "memory");
}

struct DbStepCtx {
	const uint4* meta_s;
	const uint32_t* info_s;
	uint32_t tab_sa, lane_off, nsteps;
	double* x;                          // x + this lane's first column
	uint64_t pitch;
	double alpha, beta, tmag;
};

// the steps of one tile.  FIRST: the first pass (x = beta x + alpha (diag y + hops)), otherwise x += alpha hops.  READX: the old
// x is read (always but for the first pass with beta == 0).  DOTNOW: returns this thread's share of y . x_new.
template <bool FIRST, bool READX, bool DOTNOW, uint32_t NW>
__device__ __forceinline__ double db_steps(const DbStepCtx& sc, int wid, int q, bool colok, double U0, const uint32_t* k1, const double* dv1)
{
	double contrib = 0.0;
	for (uint32_t st = (uint32_t)wid; st < sc.nsteps; st += NW) {
		const uint32_t pos = st * 4u + (uint32_t)q;
		const uint32_t info = sc.info_s[st];
		const uint32_t row = sc.meta_s[pos].x;
		const bool valid = colok && row != DB_ROW_NONE;
		double* xp = sc.x + (uint64_t)row * sc.pitch;
		double2 xo = make_double2(0.0, 0.0);
		if (READX && valid) xo = __ldcg(reinterpret_cast<const double2*>(xp));
		double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0, c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
		uint32_t ta = sc.tab_sa + (info & 0x000fffffu) * 32u;
		const uint32_t pp = (info >> 20) & 63u, pm = info >> 26;
#pragma unroll 1
		for (uint32_t g = 0; g < (pp >> 1); g++, ta += 64u) db_quad(ta + (uint32_t)q * 16u, sc.lane_off, a0, a1, b0, b1);
		if (pp & 1u) { db_pair(ta + (uint32_t)q * 8u, sc.lane_off, a0, a1, b0, b1); ta += 32u; }
#pragma unroll 1
		for (uint32_t g = 0; g < (pm >> 1); g++, ta += 64u) db_quad(ta + (uint32_t)q * 16u, sc.lane_off, c0, c1, d0, d1);
		if (pm & 1u) db_pair(ta + (uint32_t)q * 8u, sc.lane_off, c0, c1, d0, d1);
		double h0 = sc.tmag * ((a0 + b0) - (c0 + d0)), h1 = sc.tmag * ((a1 + b1) - (c1 + d1));
		double yo0 = 0.0, yo1 = 0.0;                           // the state's own element: loaded late, not live across the gathers
		if (FIRST || DOTNOW) db_ld2(sc.lane_off + (pos + 1u) * DB_LINE, yo0, yo1);
		double xn0, xn1;
		if (FIRST) {
			const uint4 m = sc.meta_s[pos];
			const double dv2 = __hiloint2double((int)m.w, (int)m.z);
			h0 += (U0 * (double)__popc(k1[0] & m.y) + dv1[0] + dv2) * yo0;
			h1 += (U0 * (double)__popc(k1[1] & m.y) + dv1[1] + dv2) * yo1;
			xn0 = sc.alpha * h0;
			xn1 = sc.alpha * h1;
			if (READX) { xn0 += sc.beta * xo.x; xn1 += sc.beta * xo.y; }
		} else {
			xn0 = xo.x + sc.alpha * h0;
			xn1 = xo.y + sc.alpha * h1;
		}
		if (valid) {
#if DB_RED_LATER_PASSES
			if (!FIRST && !READX) {
				// x += alpha h as two reductions at L2 (one add per element, so the result equals load + add + store bit for bit):
				// the pass does not read x at all
				asm volatile("red.global.add.f64 [%0], %1;" ::"l"(xp), "d"(xn0) : "memory");
				asm volatile("red.global.add.f64 [%0], %1;" ::"l"(xp + 1), "d"(xn1) : "memory");
			} else
#endif
			__stcg(reinterpret_cast<double2*>(xp), make_double2(xn0, xn1));
			if (DOTNOW) contrib += yo0 * xn0 + yo1 * xn1;
		}
	}
	return contrib;
}

struct DbTileRef {
	uint32_t pass, panel, blk;
	bool valid;
};

// NT threads per CTA, 2048 / NT ... 1 or 2 CTAs per SM.  Per tile: the tables arrive by one bulk copy (TMA) on an mbarrier, the
// 128-byte lines of y by 16-byte cp.async (rows taken from a shared-memory copy made while the previous tile was computed, so
// the fill issues no dependent global load), then the warps walk the steps.  The ticket, block descriptor and row list of the
// NEXT tile are fetched during the compute phase.
template <bool DOT, int NT>
__global__ void __launch_bounds__(NT, (NT <= 512 ? 2 : 1)) k_dblock(const DbKernelArgs ka)
{
	constexpr uint32_t NW = NT / 32;            // warps
	constexpr uint32_t SPR = NT / 8;            // states staged per round of the CTA (8 lanes per state)
	extern __shared__ __align__(128) unsigned char db_smem[];
	const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(db_smem);
	uint4* blob_s = reinterpret_cast<uint4*>(db_smem + ka.tile_bytes);
	const uint32_t blob_sa = tile_s + ka.tile_bytes;
	const uint32_t rows_sa = blob_sa + ka.blob_bytes;                             // [2][max_pos] row lists (current / next tile)
	const uint32_t* rows_sm = reinterpret_cast<const uint32_t*>(db_smem + ka.tile_bytes + ka.blob_bytes);
	__shared__ unsigned long long s_ticket[2];
	__shared__ DbBlock s_bd[2];
	__shared__ __align__(8) unsigned long long s_bar;
	__shared__ double s_red[NW];
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	const int q = lane >> 3, c = lane & 7;
	DbArgs a = ka.a;
	if (a.alpha_dev) a.alpha = *a.alpha_dev;
	if (a.beta_dev) a.beta = *a.beta_dev;
	const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&s_bar);
	if (tid < 32) reinterpret_cast<float*>(db_smem)[tid] = 0.0f;                 // slot 0 of the tile is the zero line
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	uint32_t nbt = 0;
	for (uint32_t k = 0; k < ka.npass; k++) nbt += ka.nb[k];
	const uint32_t L = max(min(ka.lag, ka.npanels), 1u);
	const unsigned long long total = ((unsigned long long)ka.npanels + (unsigned long long)(ka.npass - 1) * L) * nbt;
	const uint32_t last = ka.npass - 1;
	const bool need_x1 = a.beta != 0.0;
	const uint32_t lane_off = tile_s + (uint32_t)c * 16u;
	const uint32_t max_pos = ka.max_pos;
	uint32_t bphase = 0;
	auto tile_of = [&](unsigned long long t) {
		DbTileRef T;
		T.pass = T.panel = T.blk = 0;
		T.valid = t < total && db_decode(ka, t, L, nbt, T.pass, T.panel, T.blk);
		return T;
	};
	// thread 0: the next ticket that holds a tile (>= total when the sweep is over)
	auto take_ticket = [&]() {
		unsigned long long t;
		uint32_t p0, p1, p2;
		do t = atomicAdd(ka.ticket, 1ull); while (t < total && !db_decode(ka, t, L, nbt, p0, p1, p2));
		return t;
	};
	// stage descriptor + row list of tile T into buffer `buf` (asynchronously)
	auto stage_next = [&](const DbTileRef& T, uint32_t buf) {
		if (!T.valid) return;
		const DbBlock* bdp = ka.blocks[T.pass] + T.blk;
		if (tid == 0) s_bd[buf] = *bdp;
		const uint32_t ns = __ldg(&bdp->nstates), ro = __ldg(&bdp->rows_off);
		const uint32_t* __restrict__ rows = ka.rows[T.pass] + ro;
		for (uint32_t p = tid; p < ns; p += NT)
			asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(rows_sa + (buf * max_pos + p) * 4u), "l"(rows + p) : "memory");
	};
	if (tid == 0) s_ticket[0] = take_ticket();
	__syncthreads();
	DbTileRef cur = tile_of(s_ticket[0]);
	stage_next(cur, 0);
	asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
	__syncthreads();
#ifdef DB_PROFILE
	long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	long long tk = clock64();
#define DB_TICK(i_) do { if (tid == 0) { const long long now_ = clock64(); pf[i_] += now_ - tk; tk = now_; } } while (0)
#else
#define DB_TICK(i_) do { } while (0)
#endif
	for (uint32_t it = 0; cur.valid; it++) {
		const uint32_t buf = it & 1u;
		const uint32_t pass = cur.pass, panel = cur.panel, blk = cur.blk;
		const DbBlock bd = s_bd[buf];
		const uint64_t col0 = (uint64_t)panel * DB_COLS;
		const uint64_t mycol = col0 + 2u * c;
		const bool colok = mycol < a.ncols;
		// ---- tables: one bulk copy; tile: one state's 128-byte line per 8 lanes (cp.async) or per thread (bulk copy)
		if (tid == 0) {
			uint32_t bytes = bd.blob_len * 16u;
#if DB_FILL_MODE == 2
			bytes += bd.nstates * (uint32_t)(min((uint64_t)DB_COLS, a.ncols - col0) * 8u);
#endif
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(blob_sa),
			             "l"(ka.blob[pass] + bd.blob_off), "r"(bd.blob_len * 16u), "r"(bar_s)
			             : "memory");
			s_ticket[buf ^ 1u] = take_ticket();                             // next tile's ticket, read after the fill barrier
		}
		{
			const uint32_t* rows = rows_sm + buf * max_pos;
#if DB_FILL_MODE == 0
			const double* ycol = a.y + (colok ? mycol : 0);
			const uint32_t nbytes = colok ? 16u : 0u;
			uint32_t dst = tile_s + ((uint32_t)(tid >> 3) + 1u) * DB_LINE + (uint32_t)c * 16u;
			for (uint32_t p = tid >> 3; p < bd.nstates; p += SPR, dst += SPR * DB_LINE)
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(ycol + (uint64_t)rows[p] * a.pitch), "r"(nbytes) : "memory");
			asm volatile("cp.async.commit_group;" ::: "memory");
#else
			// the tx-count of an mbarrier may run negative inside a phase, so these copies need not wait for thread 0's expect_tx
			const uint32_t line = (uint32_t)(min((uint64_t)DB_COLS, a.ncols - col0) * 8u);
			const double* ycol = a.y + col0;
			for (uint32_t p = tid; p < bd.nstates; p += NT)
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tile_s + (p + 1u) * DB_LINE),
				             "l"(ycol + (uint64_t)rows[p] * a.pitch), "r"(line), "r"(bar_s)
				             : "memory");
#endif
		}
		DB_TICK(0);
		if (pass != 0 && tid == 0) {
			// wait until every tile of the previous pass of this panel has written its x
			const uint32_t want = ka.nb[pass - 1];
			const uint32_t* flag = ka.done + (uint64_t)(pass - 1) * ka.npanels + panel;
			uint32_t seen;
			do {
				asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
				if (seen < want) __nanosleep(100);
			} while (seen < want);
		}
		DB_TICK(1);
		db_mbar_wait(bar_s, bphase);
		bphase ^= 1u;
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncthreads();
		DB_TICK(2);
		// next tile: descriptor and rows into the other buffer while this tile is computed
		const DbTileRef nxt = tile_of(s_ticket[buf ^ 1u]);
		stage_next(nxt, buf ^ 1u);
		asm volatile("cp.async.commit_group;" ::: "memory");

		// ---- compute: a warp takes 4 states per step, 8 lanes (2 columns each) per state.  One instantiation of the step loop
		// per kind of pass, so that the first pass's diagonal operands and the last pass's dot product cost the others no registers
		DbStepCtx sc;
		sc.meta_s = blob_s;
		sc.info_s = reinterpret_cast<const uint32_t*>(blob_s + bd.nsteps * 4u);
		sc.tab_sa = blob_sa + (bd.nsteps * 4u + ((bd.nsteps + 3u) >> 2)) * 16u;
		sc.lane_off = lane_off;
		sc.nsteps = bd.nsteps;
		sc.x = a.x + mycol;
		sc.pitch = a.pitch;
		sc.alpha = a.alpha;
		sc.beta = a.beta;
		sc.tmag = a.tmag;
		double contrib = 0.0;
		if (pass == 0) {
			uint32_t k1[2] = {0u, 0u};
			double dv1[2] = {0.0, 0.0};
			if (colok) {
				k1[0] = __ldg(a.w1 + mycol); k1[1] = __ldg(a.w1 + mycol + 1);
				dv1[0] = __ldg(a.dv1 + mycol); dv1[1] = __ldg(a.dv1 + mycol + 1);
			}
			if (need_x1) db_steps<true, true, false, NW>(sc, wid, q, colok, a.U0, k1, dv1);
			else db_steps<true, false, false, NW>(sc, wid, q, colok, a.U0, k1, dv1);
		} else if (DOT && pass == last) {
			contrib = db_steps<false, true, true, NW>(sc, wid, q, colok, 0.0, nullptr, nullptr);
		} else {
#if DB_RED_LATER_PASSES
			db_steps<false, false, false, NW>(sc, wid, q, colok, 0.0, nullptr, nullptr);
#else
			db_steps<false, true, false, NW>(sc, wid, q, colok, 0.0, nullptr, nullptr);
#endif
		}
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		DB_TICK(pass == 0 ? 3 : 4);
		__syncthreads();                                   // all x of this tile are written, the tile may be overwritten
		DB_TICK(5);
		if (pass != last) {
			if (tid == 0) {
				__threadfence();
				atomicAdd(ka.done + (uint64_t)pass * ka.npanels + panel, 1u);
			}
		} else if (DOT) {
			// deterministic per-tile partial sum
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) contrib += __shfl_down_sync(0xffffffffu, contrib, o);
			if (lane == 0) s_red[wid] = contrib;
			__syncthreads();
			if (wid == 0) {
				double v = lane < (int)NW ? s_red[lane] : 0.0;
#pragma unroll
				for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
				if (lane == 0) a.dot_partials[(uint64_t)panel * ka.nb[last] + blk] = v;
			}
		}
		cur = nxt;
	}
#ifdef DB_PROFILE
	if (tid == 0 && ka.profile)
		for (int i = 0; i < 8; i++) ka.profile[blockIdx.x * 8 + i] = pf[i];
#endif
}

template <int NT>
static int db_launch_nt(DbDevPlan& dp, const DbKernelArgs& ka, unsigned grid, size_t smem, bool dot, cudaStream_t s)
{
	if (!dp.attr_set) {
		if (lpp_raise_smem(k_dblock<false, NT>, (size_t)(smem)) != cudaSuccess) return -1;
		if (lpp_raise_smem(k_dblock<true, NT>, (size_t)(smem)) != cudaSuccess) return -1;
		dp.attr_set = true;
	}
	if (dot) k_dblock<true, NT><<<grid, NT, smem, s>>>(ka);
	else k_dblock<false, NT><<<grid, NT, smem, s>>>(ka);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// returns 0 on success
static int db_launch(DbDevPlan& dp, const DbArgs& a, int nsm, cudaStream_t s)
{
	const uint32_t npanels = (uint32_t)((a.ncols + DB_COLS - 1) / DB_COLS);
	if (npanels == 0) return 0;
	const size_t ctrl_bytes = 8 + (size_t)npanels * 4 * (size_t)(dp.npass - 1);
	if (dp.ctrl_panels < npanels) {
		cudaFree(dp.ctrl);
		dp.ctrl = nullptr;
		if (cudaMalloc(&dp.ctrl, ctrl_bytes) != cudaSuccess) return -1;
		dp.ctrl_panels = npanels;
	}
	const size_t smem = dp.smem_bytes + (size_t)dp.max_pos * 8;
	if (cudaMemsetAsync(dp.ctrl, 0, ctrl_bytes, s) != cudaSuccess) return -1;
	DbKernelArgs ka;
	ka.a = a;
	unsigned long long tiles = 0;
	for (int p = 0; p < DB_MAX_PASS; p++) {
		const bool on = p < dp.npass;
		ka.blocks[p] = on ? dp.pass[p].blocks : nullptr;
		ka.blob[p] = on ? dp.pass[p].blob : nullptr;
		ka.rows[p] = on ? dp.pass[p].rows : nullptr;
		ka.nb[p] = on ? dp.pass[p].nblocks : 0u;
		tiles += (unsigned long long)npanels * ka.nb[p];
	}
	ka.npass = (uint32_t)dp.npass;
	ka.npanels = npanels;
	ka.lag = (uint32_t)std::max(dp.lag, 1);
	ka.tile_bytes = dp.tile_bytes;
	ka.blob_bytes = (uint32_t)(dp.smem_bytes - dp.tile_bytes);
	ka.max_pos = dp.max_pos;
	ka.profile = dp.profile;
	ka.ticket = dp.ctrl;
	ka.done = reinterpret_cast<uint32_t*>(dp.ctrl + 1);
	const unsigned grid = (unsigned)std::min<unsigned long long>((unsigned long long)nsm * (unsigned)dp.ctas_per_sm, tiles);
	return dp.threads == 512 ? db_launch_nt<512>(dp, ka, grid, smem, a.dot_partials != nullptr, s)
	                         : db_launch_nt<1024>(dp, ka, grid, smem, a.dot_partials != nullptr, s);
}
