// lpp_dblock_kernel.cuh -- two-pass BLOCK sweep over one spin species of a product basis (HubbardHelper.h:119-133,191-243:
// the spin-down hopping terms and the diagonal of x += H y):   x = beta x + alpha (D + 1 (x) T_dn) y.
//
// Pick two disjoint sets of sites F1, F2 with no hopping amplitude between them.  Pass 1 groups the down states by their
// occupation of F1: every hop that does not touch F1 stays inside its group ("block").  Pass 2 groups them by F2 and applies
// the hops that touch F1; none of those touches F2, so they stay inside the F2 blocks.  A tile = (block, 16 columns of the
// Ndn x Nup matrix) of y is staged in shared memory; each hop operand is a conflict-free 16-byte shared-memory load (the 8
// lanes of a state read one 128-byte line).  No operand lies outside the tile: the global traffic is y once and x read + write
// per pass.  A persistent grid takes tiles from a ticket counter in panel-major order, pass 2 of a panel LAG panels behind
// pass 1, so a panel's x and y stay L2 resident between its two passes (DRAM sees 24 bytes per element).
//
// This header holds the host-side plan builder, the kernel and the launcher; it is included by lpp_dblock.cu (engine) and by
// tools/proto_dblock.cu (stand-alone timing harness).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#define DB_THREADS 1024
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
	uint32_t f1 = 0, f2 = 0;
	DbHostPass pass[2];
	size_t smem_bytes = 0;
	uint32_t tile_bytes = 0;           // (max_pos + 1) * 128: slot 0 is the zero line
	double tmag = 1.0;                 // the one hop magnitude (entries carry signs only)
};

// ---------------------------------------------------------------------------------------------------------------
// host: plan
// ---------------------------------------------------------------------------------------------------------------
template <class W>
static bool db_build_pass(const W* words, uint64_t n, const uint32_t* idx, const double* val, const uint32_t* cnt, const double* dv2,
                          uint64_t fmask_group, uint64_t f1mask, int which, DbHostPass* out, std::string* err)
{
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
				const bool touches_f1 = (diff & f1mask) != 0;
				if ((which == 0) == touches_f1) continue;
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
		for (uint32_t p = 0; p < ns; p++) pos_of[key[b0 + order[p]].second] = p;
		const uint32_t npos = (ns + 3) & ~3u, nsteps = npos / 4;
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
	hp->tmag = mag;
	const uint64_t all = (nbits == 64) ? ~0ull : ((1ull << nbits) - 1);
	// smallest number of fixed sites whose largest block fits; F1 = highest sites possible (its blocks are then runs of rows)
	for (int f = 1; f <= nbits / 2; f++) {
		// enumerate f-subsets in descending order of their mask
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
				for (int p = 0; p < 2 && mx != 0xffffffffu; p++) {
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
				const size_t tile = ((size_t)((mx + 3) & ~3u) + 1) * DB_LINE;
				if (tile + 16 * 1024 <= max_smem) {           // room for the tables is checked exactly below
					DbHostPlan cand;
					cand.f1 = (uint32_t)F1;
					cand.f2 = (uint32_t)F2;
					cand.tmag = mag;
					std::string e2;
					if (db_build_pass(words, n, idx, val, cnt, dv2, F1, F1, 0, &cand.pass[0], &e2) &&
					    db_build_pass(words, n, idx, val, cnt, dv2, F2, F1, 1, &cand.pass[1], &e2)) {
						const uint32_t mp = std::max(cand.pass[0].max_pos, cand.pass[1].max_pos);
						const uint32_t mb = std::max(cand.pass[0].max_blob, cand.pass[1].max_blob);
						cand.tile_bytes = (mp + 1) * DB_LINE;
						cand.smem_bytes = (size_t)cand.tile_bytes + (size_t)mb * 16;
						if (cand.smem_bytes + (size_t)mp * 8 + 512 <= max_smem) { *hp = std::move(cand); return true; }
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
	uint32_t* rows = nullptr;
	uint32_t nblocks = 0;
};
struct DbDevPlan {
	DbDevPass pass[2];
	unsigned long long* ctrl = nullptr;   // [0] ticket counter, then uint32 done counters per panel group
	uint32_t ctrl_groups = 0;
	size_t smem_bytes = 0;
	uint32_t tile_bytes = 0;
	uint32_t max_pos = 0;
	int lag = 8;                          // pass 2 runs this many panels behind pass 1
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
	double* dot_partials;              // optional: per pass-2 tile partial sums of y . x_new  [npanels * nblocks2]
};

static bool db_upload_plan(const DbHostPlan& hp, DbDevPlan* dp, std::string* err)
{
	for (int p = 0; p < 2; p++) {
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
	dp->max_pos = std::max(hp.pass[0].max_pos, hp.pass[1].max_pos);
	dp->tmag = hp.tmag;
	if (cudaGetLastError() != cudaSuccess) { *err = "plan upload failed"; return false; }
	return true;
}

static void db_free_plan(DbDevPlan* dp)
{
	for (int p = 0; p < 2; p++) {
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
#define DB_NW (DB_THREADS / 32)
#define DB_SPR (DB_THREADS / 8)         // states staged per round of the CTA (8 lanes per state)
#ifndef DB_DYNAMIC_STEPS
#define DB_DYNAMIC_STEPS 0               // 1: warps take steps from a shared-memory counter (measured 2.00 ms against 1.88 ms)
#endif
#ifndef DB_SKIP_PADDING
#define DB_SKIP_PADDING 0                // 1: predicate padded operands off (measured 1.92 ms against 1.88 ms without)
#endif
#ifndef DB_FILL_MODE
#define DB_FILL_MODE 0                  // 0: 16-byte cp.async per thread (4.2 k cycles per tile), 1: LDG.128 batches + STS.128 (9 k)
#endif

struct DbKernelArgs {
	DbArgs a;
	const DbBlock* blocks[2];
	const uint4* blob[2];
	const uint32_t* rows[2];
	uint32_t nb[2];
	uint32_t npanels, lag;
	uint32_t tile_bytes, blob_bytes, max_pos;   // shared-memory layout: tile | tables | 2 row lists of max_pos words
	long long* profile;
	unsigned long long* ticket;
	uint32_t* done1;
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

// One CTA per SM.  Per tile: the tables arrive by one bulk copy (TMA) on an mbarrier, the 128-byte lines of y by 16-byte
// cp.async (rows taken from a shared-memory copy made while the previous tile was computed, so the fill issues no dependent
// global load), then 32 warps walk the steps.  The ticket, block descriptor and row list of the NEXT tile are fetched during
// the compute phase.
template <bool DOT>
__global__ void __launch_bounds__(DB_THREADS, 1) k_dblock(const DbKernelArgs ka)
{
	extern __shared__ __align__(128) unsigned char db_smem[];
	const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(db_smem);
	uint4* blob_s = reinterpret_cast<uint4*>(db_smem + ka.tile_bytes);
	const uint32_t blob_sa = tile_s + ka.tile_bytes;
	const uint32_t rows_sa = blob_sa + ka.blob_bytes;                             // [2][max_pos] row lists (current / next tile)
	const uint32_t* rows_sm = reinterpret_cast<const uint32_t*>(db_smem + ka.tile_bytes + ka.blob_bytes);
	__shared__ unsigned long long s_ticket[2];
	__shared__ DbBlock s_bd[2];
	__shared__ __align__(8) unsigned long long s_bar;
	__shared__ double s_red[DB_NW];
	__shared__ uint32_t s_step;
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
	const uint32_t nbt = ka.nb[0] + ka.nb[1];
	const unsigned long long total = (unsigned long long)ka.npanels * nbt;
	const uint32_t L = min(ka.lag, ka.npanels);
	const bool need_x1 = a.beta != 0.0;
	const uint32_t lane_off = tile_s + (uint32_t)c * 16u;
	const uint32_t max_pos = ka.max_pos;
	uint32_t bphase = 0;
	auto tile_of = [&](unsigned long long t) {
		DbTileRef T;
		T.valid = t < total;
		T.pass = T.panel = T.blk = 0;
		if (T.valid) db_decode(ka, t, L, nbt, T.pass, T.panel, T.blk);
		return T;
	};
	// stage descriptor + row list of tile T into buffer `buf` (asynchronously)
	auto stage_next = [&](const DbTileRef& T, uint32_t buf) {
		if (!T.valid) return;
		const DbBlock* bdp = ka.blocks[T.pass] + T.blk;
		if (tid == 0) s_bd[buf] = *bdp;
		const uint32_t ns = __ldg(&bdp->nstates), ro = __ldg(&bdp->rows_off);
		const uint32_t* __restrict__ rows = ka.rows[T.pass] + ro;
		for (uint32_t p = tid; p < ns; p += DB_THREADS)
			asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(rows_sa + (buf * max_pos + p) * 4u), "l"(rows + p) : "memory");
	};
	if (tid == 0) s_ticket[0] = atomicAdd(ka.ticket, 1ull);
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
		const uint64_t mycol = (uint64_t)panel * DB_COLS + 2u * c;
		const bool colok = mycol < a.ncols;
		// ---- tables: one bulk copy; tile: 16 bytes per thread and copy (8 lanes = one state's 128-byte line)
		if (tid == 0) {
			const uint32_t bytes = bd.blob_len * 16u;
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(blob_sa),
			             "l"(ka.blob[pass] + bd.blob_off), "r"(bytes), "r"(bar_s)
			             : "memory");
			s_ticket[buf ^ 1u] = atomicAdd(ka.ticket, 1ull);                 // next tile's ticket, read after the fill barrier
			s_step = DB_NW;                                                  // steps 0 .. DB_NW-1 are the warps' first steps
		}
		{
			const uint32_t* rows = rows_sm + buf * max_pos;
			const double* ycol = a.y + (colok ? mycol : 0);
#if DB_FILL_MODE == 0
			const uint32_t nbytes = colok ? 16u : 0u;
			uint32_t dst = tile_s + ((uint32_t)(tid >> 3) + 1u) * DB_LINE + (uint32_t)c * 16u;
			for (uint32_t p = tid >> 3; p < bd.nstates; p += DB_SPR, dst += DB_SPR * DB_LINE)
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(ycol + (uint64_t)rows[p] * a.pitch), "r"(nbytes) : "memory");
			asm volatile("cp.async.commit_group;" ::: "memory");
#else
			// batches of 8 independent 16-byte loads per thread, then 8 shared-memory stores
			for (uint32_t p0 = tid >> 3; p0 < bd.nstates; p0 += DB_SPR * 8u) {
				uint4 v[8];
#pragma unroll
				for (int k = 0; k < 8; k++) {
					const uint32_t p = p0 + (uint32_t)k * DB_SPR;
					v[k] = make_uint4(0u, 0u, 0u, 0u);
					if (colok && p < bd.nstates) v[k] = __ldcs(reinterpret_cast<const uint4*>(ycol + (uint64_t)rows[p] * a.pitch));
				}
#pragma unroll
				for (int k = 0; k < 8; k++) {
					const uint32_t p = p0 + (uint32_t)k * DB_SPR;
					if (p < bd.nstates)
						asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(tile_s + (p + 1u) * DB_LINE + (uint32_t)c * 16u), "r"(v[k].x), "r"(v[k].y),
						             "r"(v[k].z), "r"(v[k].w)
						             : "memory");
				}
			}
#endif
		}
		uint32_t k1[2] = {0u, 0u};
		double dv1[2] = {0.0, 0.0};
		if (pass == 0 && colok) {
			k1[0] = __ldg(a.w1 + mycol); k1[1] = __ldg(a.w1 + mycol + 1);
			dv1[0] = __ldg(a.dv1 + mycol); dv1[1] = __ldg(a.dv1 + mycol + 1);
		}
		DB_TICK(0);
		if (pass == 1 && tid == 0) {
			// wait until every pass-1 tile of this panel has written its x
			const uint32_t want = ka.nb[0];
			uint32_t seen;
			do {
				asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ka.done1 + panel) : "memory");
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

		// ---- compute: a warp takes 4 states per step, 8 lanes (2 columns each) per state
		const uint4* meta_s = blob_s;
		const uint32_t* info_s = reinterpret_cast<const uint32_t*>(blob_s + bd.nsteps * 4u);
		const uint32_t tab_sa = blob_sa + (bd.nsteps * 4u + ((bd.nsteps + 3u) >> 2)) * 16u;
		const bool read_x = pass == 1 || need_x1;
		double contrib = 0.0;
#if DB_DYNAMIC_STEPS
		// steps are handed out from a shared-memory counter (they are sorted longest first): with a fixed stride the last round
		// has 231 - 7 * 32 = 7 steps for 32 warps and the tile waits for them.  The next index is fetched one step ahead.
		uint32_t st_next = 0;
		if (lane == 0) st_next = atomicAdd(&s_step, 1u);
		for (uint32_t st = wid; st < bd.nsteps;) {
#else
		for (uint32_t st = wid; st < bd.nsteps; st += DB_NW) {
#endif
			const uint32_t pos = st * 4u + q;
			const uint32_t info = info_s[st];
			const uint4 m = meta_s[pos];
			const bool valid = colok && m.x != DB_ROW_NONE;
			double* xp = a.x + (uint64_t)m.x * a.pitch + mycol;
			double2 xo = make_double2(0.0, 0.0);
			if (valid && read_x) xo = __ldcg(reinterpret_cast<const double2*>(xp));
			double yo0 = 0.0, yo1 = 0.0;
			if (pass == 0 || DOT) db_ld2(lane_off + (pos + 1u) * DB_LINE, yo0, yo1);
			double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0, c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
			uint32_t ta = tab_sa + (info & 0x000fffffu) * 32u;
			const uint32_t pp = (info >> 20) & 63u, pm = info >> 26;
#pragma unroll 1
			for (uint32_t g = 0; g < (pp >> 1); g++, ta += 64u) db_quad(ta + (uint32_t)q * 16u, lane_off, a0, a1, b0, b1);
			if (pp & 1u) { db_pair(ta + (uint32_t)q * 8u, lane_off, a0, a1, b0, b1); ta += 32u; }
#pragma unroll 1
			for (uint32_t g = 0; g < (pm >> 1); g++, ta += 64u) db_quad(ta + (uint32_t)q * 16u, lane_off, c0, c1, d0, d1);
			if (pm & 1u) db_pair(ta + (uint32_t)q * 8u, lane_off, c0, c1, d0, d1);
			double h0 = a.tmag * ((a0 + b0) - (c0 + d0)), h1 = a.tmag * ((a1 + b1) - (c1 + d1));
			double xn0, xn1;
			if (pass == 0) {
				const double dv2 = __hiloint2double((int)m.w, (int)m.z);
				h0 += (a.U0 * (double)__popc(k1[0] & m.y) + dv1[0] + dv2) * yo0;
				h1 += (a.U0 * (double)__popc(k1[1] & m.y) + dv1[1] + dv2) * yo1;
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
#if DB_DYNAMIC_STEPS
			st = __shfl_sync(0xffffffffu, st_next, 0);
			if (lane == 0 && st < bd.nsteps) st_next = atomicAdd(&s_step, 1u);
#endif
		}
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		DB_TICK(3 + pass);
		__syncthreads();                                   // all x of this tile are written, the tile may be overwritten
		DB_TICK(5);
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
	}
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
	if (dp.ctrl_groups < npanels) {
		cudaFree(dp.ctrl);
		dp.ctrl = nullptr;
		if (cudaMalloc(&dp.ctrl, 8 + (size_t)npanels * 4) != cudaSuccess) return -1;
		dp.ctrl_groups = npanels;
	}
	const size_t smem = dp.smem_bytes + (size_t)dp.max_pos * 8;
	if (!dp.attr_set) {
		if (cudaFuncSetAttribute(k_dblock<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
		if (cudaFuncSetAttribute(k_dblock<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
		dp.attr_set = true;
	}
	if (cudaMemsetAsync(dp.ctrl, 0, 8 + (size_t)npanels * 4, s) != cudaSuccess) return -1;
	DbKernelArgs ka;
	ka.a = a;
	for (int p = 0; p < 2; p++) {
		ka.blocks[p] = dp.pass[p].blocks;
		ka.blob[p] = dp.pass[p].blob;
		ka.rows[p] = dp.pass[p].rows;
		ka.nb[p] = dp.pass[p].nblocks;
	}
	ka.npanels = npanels;
	ka.lag = (uint32_t)std::max(dp.lag, 1);
	ka.tile_bytes = dp.tile_bytes;
	ka.blob_bytes = (uint32_t)(dp.smem_bytes - dp.tile_bytes);
	ka.max_pos = dp.max_pos;
	ka.profile = dp.profile;
	ka.ticket = dp.ctrl;
	ka.done1 = reinterpret_cast<uint32_t*>(dp.ctrl + 1);
	const unsigned long long total = (unsigned long long)npanels * (ka.nb[0] + ka.nb[1]);
	const unsigned grid = (unsigned)std::min<unsigned long long>((unsigned long long)nsm, total);
	if (a.dot_partials) k_dblock<true><<<grid, DB_THREADS, smem, s>>>(ka);
	else k_dblock<false><<<grid, DB_THREADS, smem, s>>>(ka);
	return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
