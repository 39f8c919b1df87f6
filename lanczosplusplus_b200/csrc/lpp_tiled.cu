// lpp_tiled.cu -- product-basis fast path (HubbardOneBand, FeAsBasedSc hopping part).
//
// The vector is the matrix Y[idn][iup] (index = iup + idn*Nup, BasisHubbardLanczos.h:59-63) and
// H = D + 1 (x) T_dn + T_up (x) 1 (+ on-site two-spin terms for FeAs), applied as
//   sweep A  x = beta x + alpha (D + 1 (x) T_dn) y : tile = (hop-closed block of down states) x (W contiguous columns).
//            The block (all arrangements of the particles on the lower sites for one fixed pattern of the top F sites)
//            is staged in shared memory; hops inside the block are conflict-free shared-memory reads with a
//            sub-warp-uniform table entry, hops that leave the block are coalesced W*8-byte reads served by L2.
//   sweep B  x += alpha (T_up (x) 1) y             : tile = (R rows) x (contiguous block of up states), rows interleaved in
//            shared memory so one 16-byte gather serves both rows; table entries are 4 bytes.
//   sweep C  x += alpha (two-spin terms) y         : FeAs U2/U3 only.
// Hop tables are per spin species (O(N_spin * z) entries), built on device; the Hamiltonian itself is never stored.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <vector>
#include "lpp_tiled.cuh"
#include "lpp_sweep_common.cuh"
#include "lpp_dtile.cuh"
#include "lpp_smem_attr.cuh"
#include "lpp_dblock.cuh"

static thread_local std::string g_terr;
const char* lpp_tiled_error() { return g_terr.c_str(); }

#define TCK(call)                                                                  \
	do {                                                                           \
		cudaError_t e_ = (call);                                                   \
		if (e_ != cudaSuccess) { g_terr = std::string(#call) + ": " + cudaGetErrorString(e_); return -1; } \
	} while (0)

// compressed table entry: [31] sign, [30] leaves the block, [24..29] magnitude index, [0..23] index
// (index = position inside the block for internal hops, one-spin state index for external hops)
#define TE_SIGN 0x80000000u
#define TE_EXT 0x40000000u
#define TE_IDX 0x00ffffffu
#define TE_HOLE 0xffffffffu

struct SpinPlan {
	uint32_t* tab = nullptr;      // down: row-major [s][width]; up: column-major [k][n]
	uint32_t* meta = nullptr;     // per state: cnt | (next << 8)   (externals are sorted first)
	uint32_t* rowlist = nullptr;  // states grouped by block
	uint32_t* blk_off = nullptr;  // nblocks+1
	uint32_t* local = nullptr;    // position of a state inside its block
	uint32_t* blk_of = nullptr;   // block of a state
	int width = 0, nblocks = 0;
	uint32_t max_block = 0;
	uint64_t n = 0;
};

struct TiledPlan {
	uint64_t d0 = 0, dcount = 0, nloc = 0;
	int v2 = 0;                   // 1: block kernels usable, 0: fall back to the v1 sweeps
	MagTable mt;
	SpinPlan dn, up;
	int W = 16;                   // columns per sweep-A tile
	int R = 2;                    // rows per sweep-B tile
	size_t smemA = 0, smemB = 0;
	std::vector<uint32_t> tilesA_host;
	uint32_t* tilesA = nullptr;   // (block, panel) pairs intersecting the local rows: block ids only, panel-major grid
	uint32_t ntilesA_blocks = 0, npanels = 0;
	int has_twospin = 0, tsRows = 0;
	uint32_t* ts_up = nullptr;    // two-spin tables [site][state] (two orbitals), see k_sweep_twospin_tab
	uint32_t* ts_dn = nullptr;
	double u2half = 0, u3 = 0;
	int dot_blocks = 0;
	std::vector<void*> allocs;
	size_t sched_holes = 0, sched_real = 0;
	int blocksA = 0;              // sweep A: 0 = streaming panels, 1 = shared-memory blocks (v2), 2 = pipelined blocks (v3)
	int threadsA = 1024;
	int pipeB = 1, NE = 24;       // sweep B: software-pipelined kernel with NE prefetched table slots
	uint32_t* tabL = nullptr;     // lean (branch-free) up table, [k][n], widthL slots per state
	int widthL = 0;
	uint8_t* wcntL = nullptr;     // per 32 consecutive up states: table slots needed (even)
	void* tabP = nullptr;         // packed up table [chunk][group of 4 slots][lane] (k_sweep_up_packed)
	uint32_t* choffP = nullptr;   // first group of every chunk
	uint8_t* wcnt4P = nullptr;    // slots per chunk, multiple of 4
	int packedE16 = 0, packedB = 0, packedNG = 8;
	size_t smemBP = 0;
	double packed_mean_slots = 0;
	int leanA = 1, leanB = 0;
	// staged down sweep (k_sweep_down_staged): runs of consecutive down states that share their high sites
	uint32_t* stTab = nullptr;    // [n2][stWidth] entries, in-run sources first (run-local row), then the others (global row)
	uint16_t* stCnt = nullptr;    // [n2] in-run count | other count << 8
	uint32_t* stOff = nullptr;    // [stNblk + 1] first down state of every run
	std::vector<uint32_t> stOff_host;
	uint32_t stNblk = 0, stMaxRows = 0;
	int stWidth = 0, stagedA = 0, stPC = 64, stSplit = 0, stNR = 2;
	size_t smemST = 0;
	double stInternal = 0;
	DownRowsPlan* drows = nullptr;   // row-walking sweep A (k_sweep_down_rows, lpp_dtile.cu); nullptr: streaming kernel
	DownTilePlan* dtile = nullptr;   // shared-memory tile kernel for sweep A (lpp_dtile.cu, opt-in); nullptr: streaming kernel
	DownBlockPlan* dblock = nullptr; // two-pass block sweep A (lpp_dblock.cu): default for HubbardOneBand when every down state is local
	size_t smemAL = 0, smemBL = 0;
	size_t smemA3 = 0;
	// v1 fallback
	uint32_t nrowchunks = 0, npanels_v1 = 0;
	int up_in_smem = 0;
	size_t up_smem_bytes = 0;
};

#define PA_COLS 256
#define PA_ROWS 8
#define PB_THREADS 1024
#define TA_THREADS 1024

__device__ __forceinline__ double te_amp(const MagTable& mt, uint32_t e)
{
	double a = mt.mag[(e >> 24) & 63u];
	return (e & TE_SIGN) ? -a : a;
}

// =====================================================================================================
// table compression: ELL (idx, val) -> 4-byte entries, externals first
// =====================================================================================================
__global__ void k_compress(HopTable t, MagTable mt, const uint32_t* __restrict__ blk_of, const uint32_t* __restrict__ local,
                           uint32_t* __restrict__ tab, uint32_t* __restrict__ meta, int row_major, int* bad)
{
	uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= t.n) return;
	const int cnt = (int)t.cnt[s];
	const uint32_t myblk = blk_of[s];
	int pos = 0, next = 0;
	for (int pass = 0; pass < 2; pass++) {   // pass 0: hops leaving the block, pass 1: hops inside it
		for (int k = 0; k < cnt; k++) {
			uint32_t tgt = t.idx[(uint64_t)k * t.n + s];
			double v = t.val[(uint64_t)k * t.n + s];
			bool ext = blk_of[tgt] != myblk;
			if (ext != (pass == 0)) continue;
			double av = fabs(v);
			int mi = -1;
			for (int q = 0; q < mt.nmag; q++)
				if (mt.mag[q] == av) mi = q;
			if (mi < 0) { *bad = 1; mi = 0; }
			uint32_t e = (ext ? (tgt | TE_EXT) : local[tgt]) | ((uint32_t)mi << 24) | (v < 0 ? TE_SIGN : 0u);
			if (row_major) tab[s * (uint64_t)t.width + pos] = e;
			else tab[(uint64_t)pos * t.n + s] = e;
			pos++;
		}
		if (pass == 0) next = pos;
	}
	for (int k = pos; k < t.width; k++) {
		if (row_major) tab[s * (uint64_t)t.width + k] = TE_HOLE;
		else tab[(uint64_t)k * t.n + s] = TE_HOLE;
	}
	meta[s] = (uint32_t)cnt | ((uint32_t)next << 8);
}

// =====================================================================================================
// sweep A v2: hop-closed block of down states in shared memory, W contiguous columns
// =====================================================================================================
template <int W>
__global__ void __launch_bounds__(TA_THREADS, 1)
k_sweep_down_blocks(ModelDev m, SpinPlan dn, MagTable mt, DiagTables dt, SpmvArgs a, uint64_t d0, uint64_t dcount,
                    const uint32_t* __restrict__ tiles, uint32_t ntile_blocks)
{
	extern __shared__ double ys[];                       // [block row][W]
	constexpr int RPW = 32 / W;                          // rows handled by one warp at a time
	const uint32_t panel = blockIdx.x / ntile_blocks;
	const uint32_t blk = tiles[blockIdx.x % ntile_blocks];
	const uint32_t boff = dn.blk_off[blk], bsize = dn.blk_off[blk + 1] - boff;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = TA_THREADS / 32;
	const int sub = lane / W, col = lane % W;
	const unsigned submask = (W == 32) ? 0xffffffffu : (((1u << W) - 1u) << (sub * W));
	const uint64_t n1 = m.n1;
	const uint64_t u = (uint64_t)panel * W + col;
	const bool ucol = u < n1;
	const double* __restrict__ y = a.y;

	// stage the whole block (gather sources) : W*8-byte row segments
	for (uint32_t r = warp * RPW + sub; r < bsize; r += nwarps * RPW) {
		const uint64_t d = dn.rowlist[boff + r];
		ys[r * W + col] = ucol ? y[d * n1 + u] : 0.0;
	}
	__syncthreads();
	// lanes past the last column stay alive (they take part in the shuffles) but neither load nor store
	const uint64_t uc = ucol ? u : 0;
	const word_t k1 = m.b1[uc];
	const int width = dn.width;
	for (uint32_t r = warp * RPW + sub; r < bsize; r += nwarps * RPW) {
		const uint64_t d = dn.rowlist[boff + r];
		if (d < d0 || d >= d0 + dcount) continue;        // row sharding: only local output rows (uniform per sub-warp)
		const uint32_t meta = dn.meta[d];
		const int cnt = (int)(meta & 0xffu), next = (int)((meta >> 8) & 0xffu);
		// the sub-warp reads the row's (<= 4W) table entries with coalesced loads and broadcasts them by shuffle
		const uint32_t* __restrict__ trow = dn.tab + d * (uint64_t)width;
		const uint32_t e0 = (col < width) ? trow[col] : 0u;
		const uint32_t e1 = (col + W < width) ? trow[col + W] : 0u;
		const uint32_t e2 = (col + 2 * W < width) ? trow[col + 2 * W] : 0u;
		const uint32_t e3 = (col + 3 * W < width) ? trow[col + 3 * W] : 0u;
#define TA_ENTRY(kk) __shfl_sync(submask, ((kk) < W ? e0 : (kk) < 2 * W ? e1 : (kk) < 3 * W ? e2 : e3), ((kk) % W) + sub * W)
		double acc = tiled_diag(m, dt, k1, m.b2[d], uc, d) * ys[r * W + col];
		double acc2 = 0.0;
		int k = 0;
		for (; k + 1 < next; k += 2) {                   // hops leaving the block: coalesced reads served by L2
			const uint32_t ea = TA_ENTRY(k);
			const uint32_t eb = TA_ENTRY(k + 1);
			const double va = y[(uint64_t)(ea & TE_IDX) * n1 + uc];
			const double vb = y[(uint64_t)(eb & TE_IDX) * n1 + uc];
			acc += te_amp(mt, ea) * va;
			acc2 += te_amp(mt, eb) * vb;
		}
		for (; k < next; k++) {
			const uint32_t ea = TA_ENTRY(k);
			acc += te_amp(mt, ea) * y[(uint64_t)(ea & TE_IDX) * n1 + uc];
		}
		for (; k < cnt; k++) {                           // hops inside the block: shared memory, conflict free
			const uint32_t ea = TA_ENTRY(k);
			acc2 += te_amp(mt, ea) * ys[(ea & TE_IDX) * W + col];
		}
#undef TA_ENTRY
		acc += acc2;
		if (ucol) {
			const uint64_t t = (d - d0) * n1 + u;
			double xn = a.alpha * acc;
			if (a.beta != 0.0) xn += a.beta * a.x[t];
			a.x[t] = xn;
		}
	}
}

// =====================================================================================================
// sweep A v3: same tile as v2, engineered against latency: cp.async staging of the block, row list / row metadata in
// shared memory, and a software pipeline that prefetches the next row's table entries and x value while the current
// row is being reduced.
// =====================================================================================================
struct RowCtx {
	uint32_t d, meta, e0, e1, e2, e3;
	double xold, dv2;
	word_t k2;
};

template <int W, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_sweep_down_blocks3(ModelDev m, SpinPlan dn, MagTable mt, DiagTables dt, SpmvArgs a, uint64_t d0, uint64_t dcount,
                     const uint32_t* __restrict__ tiles, uint32_t ntile_blocks, uint32_t max_block)
{
	extern __shared__ double ys[];                       // [block row][W] | rows_s[max_block] | meta_s[max_block]
	uint32_t* rows_s = reinterpret_cast<uint32_t*>(ys + (size_t)max_block * W);
	uint32_t* meta_s = rows_s + max_block;
	constexpr int NSW = THREADS / W;                     // sub-warps (row slots) per CTA
	const uint32_t panel = blockIdx.x / ntile_blocks;
	const uint32_t blk = tiles[blockIdx.x % ntile_blocks];
	const uint32_t boff = dn.blk_off[blk], bsize = dn.blk_off[blk + 1] - boff;
	const int lane = threadIdx.x & 31;
	const int sub = lane / W, col = lane % W;
	const uint32_t sw = threadIdx.x / W;
	const unsigned submask = (W == 32) ? 0xffffffffu : (((1u << W) - 1u) << (sub * W));
	const uint64_t n1 = m.n1;
	const uint64_t u = (uint64_t)panel * W + col;
	const bool ucol = u < n1;
	const uint64_t uc = ucol ? u : 0;
	const double* __restrict__ y = a.y;
	const uint32_t ys_s = (uint32_t)__cvta_generic_to_shared(ys);

	for (uint32_t r = threadIdx.x; r < bsize; r += THREADS) {
		const uint32_t d = dn.rowlist[boff + r];
		rows_s[r] = d;
		const bool local = d >= d0 && d < d0 + dcount;
		meta_s[r] = local ? dn.meta[d] : 0xffffffffu;   // 0xffffffff: not an output row of this shard
	}
	for (uint32_t r = sw; r < bsize; r += NSW) {
		const uint64_t d = dn.rowlist[boff + r];
		if (ucol) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ys_s + (r * W + col) * 8u), "l"(y + d * n1 + u));
		else ys[r * W + col] = 0.0;
	}
	asm volatile("cp.async.commit_group;");
	asm volatile("cp.async.wait_group 0;");
	__syncthreads();

	const word_t k1 = m.b1[uc];
	const double dv1 = dt.dv1[uc];
	const int width = dn.width;
	const bool need_x = a.beta != 0.0;

	auto load_row = [&](uint32_t r, RowCtx& c) {
		c.d = rows_s[r];
		c.meta = meta_s[r];
		c.e0 = c.e1 = c.e2 = c.e3 = 0u;
		c.xold = 0.0;
		if (c.meta == 0xffffffffu) return;
		const int cnt = (int)(c.meta & 0xffu);
		const uint32_t* __restrict__ trow = dn.tab + (uint64_t)c.d * width;
		if (col < cnt) c.e0 = trow[col];
		if (col + W < cnt) c.e1 = trow[col + W];
		if (col + 2 * W < cnt) c.e2 = trow[col + 2 * W];
		if (col + 3 * W < cnt) c.e3 = trow[col + 3 * W];
		c.k2 = m.b2[c.d];
		c.dv2 = dt.dv2[c.d];
		if (need_x && ucol) c.xold = a.x[((uint64_t)c.d - d0) * n1 + u];
	};

	RowCtx cur, nxt;
	uint32_t r = sw;
	if (r < bsize) load_row(r, cur);
	for (; r < bsize; r += NSW) {
		const uint32_t rn = r + NSW;
		if (rn < bsize) load_row(rn, nxt);
		if (cur.meta != 0xffffffffu) {
			const int cnt = (int)(cur.meta & 0xffu), next = (int)((cur.meta >> 8) & 0xffu);
#define TA_ENTRY(kk) __shfl_sync(submask, ((kk) < W ? cur.e0 : (kk) < 2 * W ? cur.e1 : (kk) < 3 * W ? cur.e2 : cur.e3), ((kk) % W) + sub * W)
			double diag;
			if (m.model == LPP_MODEL_HUBBARD && dt.uniformU) diag = dt.U0 * (double)lpp_popc(k1 & cur.k2) + dv1 + cur.dv2;
			else diag = tiled_diag(m, dt, k1, cur.k2, uc, cur.d);
			double acc = diag * ys[r * W + col];
			double acc2 = 0.0;
			int k = 0;
			for (; k + 1 < next; k += 2) {               // hops leaving the block: coalesced W*8-byte reads (L2)
				const uint32_t ea = TA_ENTRY(k);
				const uint32_t eb = TA_ENTRY(k + 1);
				const double va = y[(uint64_t)(ea & TE_IDX) * n1 + uc];
				const double vb = y[(uint64_t)(eb & TE_IDX) * n1 + uc];
				acc += te_amp(mt, ea) * va;
				acc2 += te_amp(mt, eb) * vb;
			}
			if (k < next) {
				const uint32_t ea = TA_ENTRY(k);
				acc += te_amp(mt, ea) * y[(uint64_t)(ea & TE_IDX) * n1 + uc];
				k++;
			}
			for (; k + 1 < cnt; k += 2) {                // hops inside the block: shared memory, conflict free
				const uint32_t ea = TA_ENTRY(k);
				const uint32_t eb = TA_ENTRY(k + 1);
				acc += te_amp(mt, ea) * ys[(ea & TE_IDX) * W + col];
				acc2 += te_amp(mt, eb) * ys[(eb & TE_IDX) * W + col];
			}
			if (k < cnt) {
				const uint32_t ea = TA_ENTRY(k);
				acc += te_amp(mt, ea) * ys[(ea & TE_IDX) * W + col];
			}
#undef TA_ENTRY
			acc += acc2;
			if (ucol) {
				double xn = a.alpha * acc;
				if (need_x) xn += a.beta * cur.xold;
				a.x[((uint64_t)cur.d - d0) * n1 + u] = xn;
			}
		}
		cur = nxt;
	}
}

// =====================================================================================================
// sweep B v2: R rows x one contiguous block of up states, rows interleaved in shared memory
// =====================================================================================================
template <int R>
__global__ void __launch_bounds__(PB_THREADS, 1)
k_sweep_up_blocks(ModelDev m, SpinPlan up, MagTable mt, SpmvArgs a, uint64_t d0, uint64_t dcount, int want_dot)
{
	extern __shared__ double ys[];                       // [block position][R]
	const uint32_t blk = blockIdx.x % up.nblocks;
	const uint64_t dl0 = (uint64_t)(blockIdx.x / up.nblocks) * R;
	const uint32_t boff = up.blk_off[blk], bsize = up.blk_off[blk + 1] - boff;  // up blocks are contiguous: state = boff + i
	const uint64_t n1 = m.n1;
	const double* __restrict__ y = a.y;
	const double* yrow[R];
	double* xrow[R];
	bool live[R];
#pragma unroll
	for (int r = 0; r < R; r++) {
		live[r] = dl0 + r < dcount;
		const uint64_t dl = live[r] ? dl0 + r : dl0;
		yrow[r] = y + (d0 + dl) * n1;
		xrow[r] = a.x + dl * n1;
	}
	// stage the rows with cp.async (no register round trip, every element in flight at once)
	const uint32_t ys_s = (uint32_t)__cvta_generic_to_shared(ys);
	for (uint32_t i = threadIdx.x; i < bsize; i += PB_THREADS) {
#pragma unroll
		for (int r = 0; r < R; r++) {
			if (live[r])
				asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ys_s + (i * R + r) * 8u), "l"(yrow[r] + boff + i));
			else
				ys[i * R + r] = 0.0;
		}
	}
	asm volatile("cp.async.commit_group;");
	asm volatile("cp.async.wait_group 0;");
	__syncthreads();
	double contrib = 0.0;
	for (uint32_t i = threadIdx.x; i < bsize; i += PB_THREADS) {
		const uint64_t u = boff + i;
		const uint32_t meta = up.meta[u];
		const int cnt = (int)(meta & 0xffu), next = (int)((meta >> 8) & 0xffu);
		const uint32_t* __restrict__ tcol = up.tab + u;
		double acc[R], xold[R];
#pragma unroll
		for (int r = 0; r < R; r++) {
			acc[r] = 0.0;
			xold[r] = (live[r] && a.beta != 0.0) ? xrow[r][u] : 0.0;   // issued early, consumed after the gathers
		}
		int k = 0;
		for (; k < next; k++) {                          // hops leaving the block (only when the up basis is split)
			const uint32_t e = tcol[(uint64_t)k * n1];
			const double amp = te_amp(mt, e);
#pragma unroll
			for (int r = 0; r < R; r++) acc[r] += amp * yrow[r][e & TE_IDX];
		}
		// hops inside the block: 4 table entries in flight, then 4 shared-memory gathers.  TE_HOLE entries are padding
		// inserted by the bank-conflict-free slot scheduling (the lane sits the slot out).
		for (; k < cnt; k += 4) {
			uint32_t e[4];
#pragma unroll
			for (int j = 0; j < 4; j++) e[j] = (k + j < cnt) ? tcol[(uint64_t)(k + j) * n1] : TE_HOLE;
#pragma unroll
			for (int j = 0; j < 4; j++) {
				if (e[j] == TE_HOLE) continue;
				const double amp = te_amp(mt, e[j]);
				if (R == 2) {
					const double2 v = reinterpret_cast<const double2*>(ys)[e[j] & TE_IDX];
					acc[0] += amp * v.x;
					acc[1] += amp * v.y;
				} else {
					acc[0] += amp * ys[e[j] & TE_IDX];
				}
			}
		}
#pragma unroll
		for (int r = 0; r < R; r++) {
			if (!live[r]) continue;
			double xn = a.beta * xold[r] + a.alpha * acc[r];
			xrow[r][u] = xn;
			contrib += ys[i * R + r] * xn;
		}
	}
	if (want_dot && a.dot_partials) {
		double s = tiled_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = s;
	}
}

// =====================================================================================================
// sweep B v4: same tile as v2, software pipelined against the L2 latency of the table: the NE table entries of the
// thread's NEXT up state are loaded (unconditionally; the table is padded with TE_HOLE) while the current state's
// gathers are reduced, so the shared-memory gathers of one state issue back to back.
// =====================================================================================================
#define PBP_THREADS 512
template <int R, int NE>
__global__ void __launch_bounds__(PBP_THREADS, 1)
k_sweep_up_pipe(ModelDev m, SpinPlan up, MagTable mt, SpmvArgs a, uint64_t d0, uint64_t dcount, int want_dot)
{
	extern __shared__ double ys[];                       // [block position][R]
	const uint32_t blk = blockIdx.x % up.nblocks;
	const uint64_t dl0 = (uint64_t)(blockIdx.x / up.nblocks) * R;
	const uint32_t boff = up.blk_off[blk], bsize = up.blk_off[blk + 1] - boff;
	const uint64_t n1 = m.n1;
	const int width = up.width;
	const double* __restrict__ y = a.y;
	const double* yrow[R];
	double* xrow[R];
	bool live[R];
#pragma unroll
	for (int r = 0; r < R; r++) {
		live[r] = dl0 + r < dcount;
		const uint64_t dl = live[r] ? dl0 + r : dl0;
		yrow[r] = y + (d0 + dl) * n1;
		xrow[r] = a.x + dl * n1;
	}
	const uint32_t ys_s = (uint32_t)__cvta_generic_to_shared(ys);
	for (uint32_t i = threadIdx.x; i < bsize; i += PBP_THREADS) {
#pragma unroll
		for (int r = 0; r < R; r++) {
			if (live[r])
				asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ys_s + (i * R + r) * 8u), "l"(yrow[r] + boff + i));
			else
				ys[i * R + r] = 0.0;
		}
	}
	asm volatile("cp.async.commit_group;");

	uint32_t en[NE], mnext = 0;
	double xn_old[R];
	auto prefetch = [&](uint32_t ii) {
		const uint64_t u = boff + ii;
		const uint32_t* __restrict__ tcol = up.tab + u;
#pragma unroll
		for (int j = 0; j < NE; j++) en[j] = (j < width) ? tcol[(uint64_t)j * n1] : TE_HOLE;
		mnext = up.meta[u];
#pragma unroll
		for (int r = 0; r < R; r++) xn_old[r] = (live[r] && a.beta != 0.0) ? xrow[r][u] : 0.0;
	};
	uint32_t i = threadIdx.x;
	if (i < bsize) prefetch(i);                          // overlaps with the cp.async staging
	asm volatile("cp.async.wait_group 0;");
	__syncthreads();

	double contrib = 0.0;
	for (; i < bsize; i += PBP_THREADS) {
		uint32_t ec[NE];
		double xold[R];
#pragma unroll
		for (int j = 0; j < NE; j++) ec[j] = en[j];
#pragma unroll
		for (int r = 0; r < R; r++) xold[r] = xn_old[r];
		const uint32_t meta = mnext;
		const uint64_t u = boff + i;
		if (i + PBP_THREADS < bsize) prefetch(i + PBP_THREADS);
		double acc[R];
#pragma unroll
		for (int r = 0; r < R; r++) acc[r] = 0.0;
#pragma unroll
		for (int j = 0; j < NE; j++) {
			const uint32_t e = ec[j];
			if (e == TE_HOLE) continue;
			const double amp = te_amp(mt, e);
			if (e & TE_EXT) {
#pragma unroll
				for (int r = 0; r < R; r++) acc[r] += amp * yrow[r][e & TE_IDX];
			} else if (R == 2) {
				const double2 v = reinterpret_cast<const double2*>(ys)[e & TE_IDX];
				acc[0] += amp * v.x;
				acc[1] += amp * v.y;
			} else {
				acc[0] += amp * ys[e & TE_IDX];
			}
		}
		const int cnt = (int)(meta & 0xffu);
		for (int k = NE; k < cnt; k++) {                 // rare: states with more than NE table slots
			const uint32_t e = up.tab[(uint64_t)k * n1 + u];
			if (e == TE_HOLE) continue;
			const double amp = te_amp(mt, e);
			if (e & TE_EXT) {
#pragma unroll
				for (int r = 0; r < R; r++) acc[r] += amp * yrow[r][e & TE_IDX];
			} else if (R == 2) {
				const double2 v = reinterpret_cast<const double2*>(ys)[e & TE_IDX];
				acc[0] += amp * v.x;
				acc[1] += amp * v.y;
			} else {
				acc[0] += amp * ys[e & TE_IDX];
			}
		}
#pragma unroll
		for (int r = 0; r < R; r++) {
			if (!live[r]) continue;
			double xn = a.beta * xold[r] + a.alpha * acc[r];
			xrow[r][u] = xn;
			contrib += ys[i * R + r] * xn;
		}
	}
	if (want_dot && a.dot_partials) {
		double s = tiled_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = s;
	}
}

// =====================================================================================================
// lean sweeps (round-1 profile: the sweeps above are instruction-issue bound, ~50 warp instructions per table slot).
// sweep B lean: every thread executes exactly NE branch-free slots per up state.  A slot is a 4-byte word
// [31] sign | [24..29] magnitude | [0..23] BYTE offset into the staged tile; padding and scheduling holes point at one of
// G zero slots behind the tile, chosen on a bank group that is free in that slot, so they cost no conflict and no branch.
// =====================================================================================================
// N branch-free gathers of one up state: ec[j] = [31] sign | [24..29] magnitude | [0..23] byte offset into the staged tile
template <int R, int N, int NE, bool UNI>
__device__ __forceinline__ void up_lean_gather(const uint32_t (&ec)[NE], const MagTable& mt, uint32_t ys_s, double (&acc)[R])
{
#pragma unroll
	for (int j = 0; j < N; j++) {
		const uint32_t e = ec[j];
		double amp;
		if (UNI) amp = __hiloint2double(0x3ff00000 | (int)(e & TE_SIGN), 0);          // +-1.0, scaled by |t| once by the caller
		else {
			const double mg = mt.mag[(e >> 24) & 63u];
			amp = __hiloint2double(__double2hiint(mg) ^ (int)(e & TE_SIGN), __double2loint(mg));
		}
		const uint32_t addr = ys_s + (e & TE_IDX);
		if (R == 2) {
			double vx, vy;
			asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(addr));
			acc[0] = fma(amp, vx, acc[0]);
			acc[1] = fma(amp, vy, acc[1]);
		} else {
			double vx;
			asm volatile("ld.shared.f64 %0, [%1];" : "=d"(vx) : "r"(addr));
			acc[0] = fma(amp, vx, acc[0]);
		}
	}
}

template <int R, int NE, bool UNI>
__global__ void __launch_bounds__(PBP_THREADS, 1)
k_sweep_up_lean(ModelDev m, const uint32_t* __restrict__ tabL, const uint8_t* __restrict__ wcnt, uint32_t boff, uint32_t bsize,
                MagTable mt, SpmvArgs a, uint64_t d0, uint64_t dcount, int want_dot)
{
	extern __shared__ double ys[];                       // [block position][R] + G zero slots
	constexpr int G = 16 / R;
	const uint64_t dl0 = (uint64_t)blockIdx.x * R;
	const uint64_t n1 = m.n1;
	const double* yrow[R];
	double* xrow[R];
	bool live[R];
#pragma unroll
	for (int r = 0; r < R; r++) {
		live[r] = dl0 + r < dcount;
		const uint64_t dl = live[r] ? dl0 + r : dl0;
		yrow[r] = a.y + (d0 + dl) * n1;
		xrow[r] = a.x + dl * n1;
	}
	const uint32_t ys_s = (uint32_t)__cvta_generic_to_shared(ys);
	for (uint32_t i = threadIdx.x; i < bsize; i += PBP_THREADS) {
#pragma unroll
		for (int r = 0; r < R; r++) {
			if (live[r])
				asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ys_s + (i * R + r) * 8u), "l"(yrow[r] + boff + i));
			else
				ys[i * R + r] = 0.0;
		}
	}
	if (threadIdx.x < G * R) ys[(size_t)bsize * R + threadIdx.x] = 0.0;
	asm volatile("cp.async.commit_group;");

	uint32_t en[NE];
	double xn_old[R];
	const bool need_x = a.beta != 0.0;
	const uint64_t stride = n1;
	// wcnt: slots the 32 consecutive states of a warp need (even, warp-uniform); only those are fetched and executed
	auto prefetch = [&](uint32_t ii) {
		const uint32_t* __restrict__ tcol = tabL + boff + ii;
		const int pc = (int)wcnt[ii >> 5];
#pragma unroll
		for (int j = 0; j < NE; j += 2)
			if (j < pc) {
				en[j] = tcol[(uint64_t)j * stride];
				en[j + 1] = tcol[(uint64_t)(j + 1) * stride];
			}
#pragma unroll
		for (int r = 0; r < R; r++) xn_old[r] = (live[r] && need_x) ? xrow[r][boff + ii] : 0.0;
	};
	uint32_t i = threadIdx.x;
	if (i < bsize) prefetch(i);                          // overlaps with the cp.async staging
	asm volatile("cp.async.wait_group 0;");
	__syncthreads();

	const double t0 = mt.mag[0];
	double contrib = 0.0;
	for (; i < bsize; i += PBP_THREADS) {
		uint32_t ec[NE];
		double xold[R];
		const int cnt = (int)wcnt[i >> 5];
#pragma unroll
		for (int j = 0; j < NE; j++) ec[j] = en[j];
#pragma unroll
		for (int r = 0; r < R; r++) xold[r] = xn_old[r];
		if (i + PBP_THREADS < bsize) prefetch(i + PBP_THREADS);
		double acc[R];
#pragma unroll
		for (int r = 0; r < R; r++) acc[r] = 0.0;
		// branch-free runs of N independent gathers; N = the warp's slot count (states with fewer hops read zero slots)
		switch (cnt) {
#define LPP_UP_CASE(N_) case N_: up_lean_gather<R, N_, NE, UNI>(ec, mt, ys_s, acc); break;
			LPP_UP_CASE(2) LPP_UP_CASE(4) LPP_UP_CASE(6) LPP_UP_CASE(8) LPP_UP_CASE(10) LPP_UP_CASE(12) LPP_UP_CASE(14)
			LPP_UP_CASE(16) LPP_UP_CASE(18) LPP_UP_CASE(20) LPP_UP_CASE(22) LPP_UP_CASE(24) LPP_UP_CASE(26) LPP_UP_CASE(28)
			LPP_UP_CASE(30)
#undef LPP_UP_CASE
		case 0: break;
		default: up_lean_gather<R, NE, NE, UNI>(ec, mt, ys_s, acc); break;
		}
#pragma unroll
		for (int r = 0; r < R; r++) {
			if (!live[r]) continue;
			const double hv = UNI ? t0 * acc[r] : acc[r];
			double xn = a.beta * xold[r] + a.alpha * hv;
			xrow[r][boff + i] = xn;
			contrib += ys[i * R + r] * xn;
		}
	}
	if (want_dot && a.dot_partials) {
		double s = tiled_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = s;
	}
}

#ifndef UPP_THREADS
#define UPP_THREADS 512
#endif
// sweep B packed: the lean kernel with (1) a per-warp slot count (a warp's 32 consecutive states execute only the slots
// they need, rounded to 4: 24 on average instead of 32 on the 4x4 lattice) and (2) the table stored per warp chunk as
// [chunk][group of 4 slots][lane], so the 4 entries of a group are one coalesced 8-byte (E16: 2-byte entries
// [15] sign | [14] magnitude index | [13:0] tile position, at most two hop magnitudes) or 16-byte (4-byte lean entries)
// load per lane.  NG = groups a state can have: 8 (32 slots) or, with 2-byte entries, 12 (48 slots: FeAs with orbital hoppings).
template <int R, bool UNI, bool E16, int NG, int N4>
__device__ __forceinline__ void up_packed_gather(const typename std::conditional<E16, uint2, uint4>::type (&ec)[NG], const MagTable& mt,
                                                 uint32_t ys_s, uint32_t hole_off, double (&acc)[R])
{
#pragma unroll
	for (int g = 0; g < N4; g++) {
		uint32_t off[4], sg[4], mi[4];
		if constexpr (E16) {
			const uint32_t w0 = ec[g].x, w1 = ec[g].y;
			off[0] = (w0 & 0x3fffu) * (8u * R); sg[0] = w0 << 16;
			off[1] = ((w0 >> 16) & 0x3fffu) * (8u * R); sg[1] = w0;
			off[2] = (w1 & 0x3fffu) * (8u * R); sg[2] = w1 << 16;
			off[3] = ((w1 >> 16) & 0x3fffu) * (8u * R); sg[3] = w1;
			mi[0] = (w0 >> 14) & 1u; mi[1] = (w0 >> 30) & 1u; mi[2] = (w1 >> 14) & 1u; mi[3] = (w1 >> 30) & 1u;
		} else {
			const uint32_t w[4] = {ec[g].x, ec[g].y, ec[g].z, ec[g].w};
#pragma unroll
			for (int i = 0; i < 4; i++) { off[i] = w[i] & TE_IDX; sg[i] = w[i]; mi[i] = (w[i] >> 24) & 63u; }
		}
#pragma unroll
		for (int i = 0; i < 4; i++) {
			double amp;
			if (UNI) amp = __hiloint2double((int)(0x3ff00000u | (sg[i] & TE_SIGN)), 0);   // +-1.0, scaled by |t| once by the caller
			else {
				const double mg = mt.mag[mi[i]];
				amp = __hiloint2double(__double2hiint(mg) ^ (int)(sg[i] & TE_SIGN), __double2loint(mg));
			}
			const uint32_t addr = ys_s + off[i];
			if (R == 2) {
				// Predicating the padding / hole entries off (a quarter-warp without an operand in a slot then costs no shared-memory
				// wavefront: 24.0 executed slots against 20.9 needed per quarter-warp) was measured and is NOT faster: x += H y
				// 3.593 ms against 3.568 ms, the extra compare and zero moves cost what the 13 % fewer wavefronts save.
				// LPP_UP_SKIP_HOLES (compile time) keeps the variant.
				double vx = 0.0, vy = 0.0;
#ifdef LPP_UP_SKIP_HOLES
				if (off[i] < hole_off)
#else
				(void)hole_off;
#endif
				asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(vx), "=d"(vy) : "r"(addr));
				acc[0] = fma(amp, vx, acc[0]);
				acc[1] = fma(amp, vy, acc[1]);
			} else {
				double vx;
				asm volatile("ld.shared.f64 %0, [%1];" : "=d"(vx) : "r"(addr));
				acc[0] = fma(amp, vx, acc[0]);
			}
		}
	}
}

template <int R, bool UNI, bool E16, int NG>
__global__ void __launch_bounds__(UPP_THREADS, 1)
k_sweep_up_packed(ModelDev m, const void* __restrict__ tabP, const uint32_t* __restrict__ choff, const uint8_t* __restrict__ wcnt4,
                  uint32_t bsize, MagTable mt, SpmvArgs a, uint64_t d0, uint64_t dcount, int want_dot)
{
	using VT = typename std::conditional<E16, uint2, uint4>::type;
	extern __shared__ double ys[];                       // [position][R] + G zero slots
	const double c_alpha = a.alpha, c_beta = a.beta;     // read once (they may live in device memory: LppCoef)
	constexpr int G = 16 / R;
	const uint64_t dl0 = (uint64_t)blockIdx.x * R;
	const uint64_t n1 = m.n1;
	const double* yrow[R];
	double* xrow[R];
	bool live[R];
#pragma unroll
	for (int r = 0; r < R; r++) {
		live[r] = dl0 + r < dcount;
		const uint64_t dl = live[r] ? dl0 + r : dl0;
		yrow[r] = a.y + (d0 + dl) * n1;
		xrow[r] = a.x + dl * n1;
	}
	const uint32_t ys_s = (uint32_t)__cvta_generic_to_shared(ys);
	// tile fill: a warp copies 256 contiguous bytes of one row per instruction into every other 8-byte slot of the interleaved
	// tile (2-way bank conflict on the shared-memory side).  Alternating the rows across lanes instead removes that conflict
	// but splits every quarter-warp over two global lines and was measured slower (2.03 vs 1.69 ms).
	for (uint32_t i = threadIdx.x; i < bsize; i += UPP_THREADS) {
#pragma unroll
		for (int r = 0; r < R; r++) {
			if (live[r])
				asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ys_s + (i * R + r) * 8u), "l"(yrow[r] + i));
			else
				ys[i * R + r] = 0.0;
		}
	}
	if (threadIdx.x < G * R) ys[(size_t)bsize * R + threadIdx.x] = 0.0;
	asm volatile("cp.async.commit_group;");
	// chunk directory (first group and slot count of every 32-state chunk) in shared memory behind the tile: the table
	// prefetch of the next state must not wait for a dependent global load
	const uint32_t nch = (bsize + 31) >> 5;
	uint32_t* s_choff = reinterpret_cast<uint32_t*>(ys + (size_t)bsize * R + G * R);
	uint8_t* s_wcnt = reinterpret_cast<uint8_t*>(s_choff + nch);
	for (uint32_t c = threadIdx.x; c < nch; c += UPP_THREADS) {
		s_choff[c] = choff[c];
		s_wcnt[c] = wcnt4[c];
	}
	__syncthreads();

	VT en[NG];
	double xn_old[R];
	const bool need_x = c_beta != 0.0;
	const uint32_t lane = threadIdx.x & 31;
	auto prefetch = [&](uint32_t ii) {
		const uint32_t c = ii >> 5;
		const int pc = (int)s_wcnt[c];
		const VT* __restrict__ base = reinterpret_cast<const VT*>(tabP) + (size_t)s_choff[c] * 32 + lane;
#pragma unroll
		for (int g = 0; g < NG; g++)
			if (4 * g < pc) en[g] = __ldg(base + g * 32);
#pragma unroll
		for (int r = 0; r < R; r++) xn_old[r] = (live[r] && need_x) ? xrow[r][ii] : 0.0;
	};
	uint32_t i = threadIdx.x;
	if (i < bsize) prefetch(i);                          // overlaps with the cp.async staging
	asm volatile("cp.async.wait_group 0;");
	__syncthreads();

	const double t0 = mt.mag[0];
	const uint32_t hole_off = bsize * (uint32_t)(8 * R);     // byte offset of the zero slots = first offset that is no operand
	double contrib = 0.0;
	for (; i < bsize; i += UPP_THREADS) {
		VT ec[NG];
		double xold[R];
		const int cnt = (int)s_wcnt[i >> 5];
#pragma unroll
		for (int g = 0; g < NG; g++) ec[g] = en[g];
#pragma unroll
		for (int r = 0; r < R; r++) xold[r] = xn_old[r];
		if (i + UPP_THREADS < bsize) prefetch(i + UPP_THREADS);
		double acc[R], yown[R];
		if (R == 2) {                                      // own elements for the dot: one 16-byte load (no 2-way conflict), issued early
			const double2 yo = reinterpret_cast<const double2*>(ys)[i];
			yown[0] = yo.x; yown[R - 1] = yo.y;
		} else yown[0] = ys[i];
#pragma unroll
		for (int r = 0; r < R; r++) acc[r] = 0.0;
		switch (cnt >> 2) {                                // warp-uniform: branch-free runs of 4*N independent gathers
#define LPP_UPP_CASE(N_) case N_: up_packed_gather<R, UNI, E16, NG, (N_ <= NG ? N_ : NG)>(ec, mt, ys_s, hole_off, acc); break;
		case 0: break;
			LPP_UPP_CASE(1) LPP_UPP_CASE(2) LPP_UPP_CASE(3) LPP_UPP_CASE(4) LPP_UPP_CASE(5) LPP_UPP_CASE(6) LPP_UPP_CASE(7)
			LPP_UPP_CASE(8) LPP_UPP_CASE(9) LPP_UPP_CASE(10) LPP_UPP_CASE(11)
#undef LPP_UPP_CASE
		default: up_packed_gather<R, UNI, E16, NG, NG>(ec, mt, ys_s, hole_off, acc); break;
		}
#pragma unroll
		for (int r = 0; r < R; r++) {
			if (!live[r]) continue;
			const double hv = UNI ? t0 * acc[r] : acc[r];
			double xn = c_beta * xold[r] + c_alpha * hv;
			xrow[r][i] = xn;
			contrib += yown[r] * xn;
		}
	}
	if (want_dot && a.dot_partials) {
		double s = tiled_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = s;
	}
}

// sweep A lean: streaming panels as in k_sweep_down, but the CTA-uniform table entries of its rows are staged once in
// shared memory as (byte offset of the source row, amplitude) pairs, so a hop costs a broadcast 16-byte shared load,
// one 64-bit add, the coalesced global load and the FMA.
#define PAL_COLS 256
#define PAL_ROWS 16
struct DownEntry {
	unsigned long long off8;   // byte offset of row idx in y
	double amp;
};
// ColView: the kernel sees an (nrows x ncols) row-major matrix with row pitch `pitch`; column c is up state u0 + c.
// Single GPU: pitch = ncols = Nup, u0 = 0.  Two-layout multi-GPU: the rank's column shard, all Ndn rows.
template <int VEC>
__global__ void __launch_bounds__(PAL_COLS, (VEC == 2) ? 3 : 4) k_sweep_down_lean(ModelDev m, HopTable dn, DiagTables dt, SpmvArgs a, uint64_t d0,
                                                               uint64_t dcount, uint32_t nrowchunks, ColView cv)
{
	extern __shared__ double ys[];
	const double c_alpha = a.alpha, c_beta = a.beta;     // read once (they may live in device memory: LppCoef)
	DownEntry* ent = reinterpret_cast<DownEntry*>(ys);                                  // [PAL_ROWS][width]
	const uint32_t panel = blockIdx.x / nrowchunks, chunk = blockIdx.x % nrowchunks;
	const uint64_t pitch = cv.pitch;
	const int width = dn.width;
	const uint64_t dl_first = (uint64_t)chunk * PAL_ROWS;
	const int nrows = (int)min((uint64_t)PAL_ROWS, dcount - dl_first);
	for (int q = threadIdx.x; q < nrows * width; q += PAL_COLS) {
		const int r = q / width, k = q % width;
		const uint64_t d = d0 + dl_first + r;
		DownEntry e;
		e.off8 = (unsigned long long)dn.idx[(uint64_t)k * dn.n + d] * pitch * 8ull;
		e.amp = dn.val[(uint64_t)k * dn.n + d];                                         // padded entries carry amp = 0, idx = d
		ent[r * width + k] = e;
	}
	__syncthreads();
	// VEC adjacent columns per thread (VEC = 2: 16-byte accesses; needs an even pitch and an even column count)
	const uint64_t c = ((uint64_t)panel * PAL_COLS + threadIdx.x) * VEC;
	double contrib = 0.0;
	if (c < cv.ncols) {
		word_t k1[VEC];
		double dv1[VEC];
#pragma unroll
		for (int v = 0; v < VEC; v++) {
			k1[v] = m.b1[cv.u0 + c + v];
			dv1[v] = dt.dv1[cv.u0 + c + v];
		}
		const char* __restrict__ ycol = reinterpret_cast<const char*>(a.y + c);
		const bool need_x = c_beta != 0.0;
		constexpr int NB = (VEC == 2) ? 4 : 8;            // independent row reads in flight per thread
#pragma unroll 1
		for (int r = 0; r < nrows; r++) {
			const uint64_t dl = dl_first + r, d = d0 + dl;
			const int cd = (int)dn.cnt[d];
			const uint64_t t = dl * pitch + c;
			double xold[VEC], yr[VEC], acc[VEC], acc2[VEC];
			if (VEC == 2) {
				const double2 yv = *reinterpret_cast<const double2*>(ycol + d * pitch * 8ull);
				yr[0] = yv.x; yr[VEC - 1] = yv.y;
				if (need_x) {
					const double2 xv = *reinterpret_cast<const double2*>(a.x + t);
					xold[0] = xv.x; xold[VEC - 1] = xv.y;
				}
			} else {
				yr[0] = *reinterpret_cast<const double*>(ycol + d * pitch * 8ull);
				if (need_x) xold[0] = a.x[t];
			}
			const word_t k2 = m.b2[d];
			const double dv2 = dt.dv2[d];
#pragma unroll
			for (int v = 0; v < VEC; v++) {
				double diag;
				if (m.model == LPP_MODEL_HUBBARD && dt.uniformU) diag = dt.U0 * (double)lpp_popc(k1[v] & k2) + dv1[v] + dv2;
				else diag = tiled_diag(m, dt, k1[v], k2, cv.u0 + c + v, d);
				acc[v] = diag * yr[v];
				acc2[v] = 0.0;
			}
			const DownEntry* __restrict__ er = ent + r * width;
			// entries past cnt are padding (amp = 0, source row = d itself), so whole batches can be issued blindly
			const int cdr = min((cd + NB - 1) / NB * NB, width);
			int k = 0;
			for (; k + NB <= cdr; k += NB) {
				DownEntry e[NB];
#pragma unroll
				for (int j = 0; j < NB; j++) e[j] = er[k + j];
				if (VEC == 2) {
					double2 v[NB];
#pragma unroll
					for (int j = 0; j < NB; j++) v[j] = *reinterpret_cast<const double2*>(ycol + e[j].off8);
#pragma unroll
					for (int j = 0; j < NB; j += 2) {
						acc[0] = fma(e[j].amp, v[j].x, acc[0]);
						acc[VEC - 1] = fma(e[j].amp, v[j].y, acc[VEC - 1]);
						acc2[0] = fma(e[j + 1].amp, v[j + 1].x, acc2[0]);
						acc2[VEC - 1] = fma(e[j + 1].amp, v[j + 1].y, acc2[VEC - 1]);
					}
				} else {
					double v[NB];
#pragma unroll
					for (int j = 0; j < NB; j++) v[j] = *reinterpret_cast<const double*>(ycol + e[j].off8);
#pragma unroll
					for (int j = 0; j < NB; j += 2) {
						acc[0] = fma(e[j].amp, v[j], acc[0]);
						acc2[0] = fma(e[j + 1].amp, v[j + 1], acc2[0]);
					}
				}
			}
			for (; k < cd; k++) {
				const DownEntry ea = er[k];
				if (VEC == 2) {
					const double2 v = *reinterpret_cast<const double2*>(ycol + ea.off8);
					acc[0] = fma(ea.amp, v.x, acc[0]);
					acc[VEC - 1] = fma(ea.amp, v.y, acc[VEC - 1]);
				} else {
					acc[0] = fma(ea.amp, *reinterpret_cast<const double*>(ycol + ea.off8), acc[0]);
				}
			}
			double xn[VEC];
#pragma unroll
			for (int v = 0; v < VEC; v++) {
				xn[v] = c_alpha * (acc[v] + acc2[v]);
				if (need_x) xn[v] += c_beta * xold[v];
				contrib += yr[v] * xn[v];
			}
			if (VEC == 2) *reinterpret_cast<double2*>(a.x + t) = make_double2(xn[0], xn[VEC - 1]);
			else a.x[t] = xn[0];
		}
	}
	if (a.dot_partials) {
		double sum = tiled_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = sum;
	}
}

static inline int down_lean_vec(const ColView& cv) { return (cv.pitch % 2 == 0 && cv.ncols % 2 == 0) ? 2 : 1; }
static inline uint32_t down_lean_panels(const ColView& cv) { return (uint32_t)((cv.ncols + (uint64_t)PAL_COLS * down_lean_vec(cv) - 1) / ((uint64_t)PAL_COLS * down_lean_vec(cv))); }


// sweep A staged: the streaming down sweep with the in-run operands taken from shared memory.  Down states are ordered by
// their word (colex), so the states that share the occupation of the high sites [split, nsite) are one consecutive run
// (<= C(split, k) rows); every hop between two of the low sites stays inside its run (4x4 lattice, split 8: 12 of the 32
// bonds, 37.5 % of the operands).  A CTA owns (run, panel of PC columns): it stages the run's rows of y once (coalesced
// 16-byte cp.async), then every row takes its in-run sources from that tile (conflict-free 16-byte LDS, consecutive lanes =
// consecutive columns) and only the other sources as coalesced row reads from L2.  The kernel is bound by the L2->SM
// fabric like k_sweep_down_lean, with fewer bytes to move.  Panel-major grid: a panel's slab of y (Ndn x PC doubles) stays
// L2 resident while its runs are processed.
#define DST_THREADS 256
struct StEntry {
	unsigned long long off;    // in-run source: byte offset of its row in the tile; other source: byte offset of its row in y
	double amp;
};
template <int PC, int NB>
__global__ void __launch_bounds__(DST_THREADS, (NB <= 4) ? 3 : 2)
k_sweep_down_staged(ModelDev m, const uint32_t* __restrict__ stab, const uint16_t* __restrict__ scnt, const uint32_t* __restrict__ sboff,
                    uint32_t blk0, uint32_t nblk_loc, int width, uint32_t maxrows, MagTable mt, DiagTables dt, SpmvArgs a, uint64_t d0,
                    uint64_t dcount, ColView cv)
{
	extern __shared__ double ys[];                                   // [run row][PC]
	StEntry* ent = reinterpret_cast<StEntry*>(ys + (size_t)maxrows * PC);     // [run row][width]
	double* s_dv2 = reinterpret_cast<double*>(ent + (size_t)maxrows * width); // [run row] one-spin diagonal part
	word_t* s_k2 = reinterpret_cast<word_t*>(s_dv2 + maxrows);       // [run row] down word
	uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_k2 + maxrows);   // [run row] in-run | others << 8
	constexpr int HP = PC / 2;                                       // 16-byte pieces (column pairs) per row
	constexpr int RPI = DST_THREADS / HP;                            // row slots per CTA
	const uint32_t panel = blockIdx.x / nblk_loc, blk = blk0 + blockIdx.x % nblk_loc;
	const uint32_t b0 = sboff[blk], nb = sboff[blk + 1] - b0;
	const uint64_t pitch = cv.pitch, pitch8 = cv.pitch * 8ull;
	const uint64_t c0 = (uint64_t)panel * PC;
	const uint32_t ys_s = (uint32_t)__cvta_generic_to_shared(ys);
	for (uint32_t q = threadIdx.x; q < nb * HP; q += DST_THREADS) {
		const uint32_t r = q / HP, cp = q % HP;
		const uint64_t c = c0 + 2ull * cp;
		if (c < cv.ncols)
			asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ys_s + (r * PC + 2u * cp) * 8u), "l"(a.y + (uint64_t)(b0 + r) * pitch + c));
	}
	asm volatile("cp.async.commit_group;");
	for (uint32_t q = threadIdx.x; q < nb * (uint32_t)width; q += DST_THREADS) {
		const uint32_t r = q / width, k = q % width;
		const uint32_t en = stab[(size_t)b0 * width + q];
		const uint32_t cnt = scnt[b0 + r];
		const uint32_t ci = cnt & 0xffu, ct = ci + (cnt >> 8);
		StEntry e;
		const double mg = mt.mag[(en >> 24) & 63u];
		e.amp = (k < ct) ? ((en & TE_SIGN) ? -mg : mg) : 0.0;
		e.off = (k < ci) ? (unsigned long long)(en & TE_IDX) * (PC * 8ull) : (unsigned long long)(en & TE_IDX) * pitch8;
		ent[q] = e;
	}
	for (uint32_t q = threadIdx.x; q < nb; q += DST_THREADS) {
		s_cnt[q] = scnt[b0 + q];
		s_dv2[q] = dt.dv2[b0 + q];
		s_k2[q] = m.b2[b0 + q];
	}
	asm volatile("cp.async.wait_group 0;");
	__syncthreads();

	const uint32_t rs = threadIdx.x / HP, cp = threadIdx.x % HP;
	const uint64_t c = c0 + 2ull * cp;
	double contrib = 0.0;
	if (c < cv.ncols) {
		const word_t k1a = m.b1[cv.u0 + c], k1b = m.b1[cv.u0 + c + 1];
		const double dv1a = dt.dv1[cv.u0 + c], dv1b = dt.dv1[cv.u0 + c + 1];
		const char* __restrict__ ycol = reinterpret_cast<const char*>(a.y + c);
		const bool need_x = a.beta != 0.0;
		const uint32_t tile_s = ys_s + 2u * cp * 8u;                 // this thread's column pair in row 0 of the tile
#pragma unroll 1
		for (uint32_t r = rs; r < nb; r += RPI) {
			const uint64_t d = (uint64_t)b0 + r;
			if (d < d0 || d >= d0 + dcount) continue;
			const uint64_t t = (d - d0) * pitch + c;
			const uint32_t cnt = s_cnt[r];
			const int ci = (int)(cnt & 0xffu), kend = ci + (int)(cnt >> 8);
			const StEntry* __restrict__ er = ent + r * width;
			double2 xv = make_double2(0.0, 0.0);
			if (need_x) xv = *reinterpret_cast<const double2*>(a.x + t);
			double2 h0 = make_double2(0.0, 0.0), h1 = make_double2(0.0, 0.0);
			// straight-line batches of N independent L2 row reads (N is warp-uniform: no per-slot predicates)
#define DST_BATCH(N_)                                                                                                   \
	{                                                                                                                   \
		StEntry e_[N_];                                                                                                 \
		double2 v_[N_];                                                                                                 \
		_Pragma("unroll") for (int j = 0; j < N_; j++) { e_[j] = er[k + j]; v_[j] = *reinterpret_cast<const double2*>(ycol + e_[j].off); } \
		_Pragma("unroll") for (int j = 0; j < N_; j++) {                                                                \
			if (j & 1) { h1.x = fma(e_[j].amp, v_[j].x, h1.x); h1.y = fma(e_[j].amp, v_[j].y, h1.y); }                  \
			else { h0.x = fma(e_[j].amp, v_[j].x, h0.x); h0.y = fma(e_[j].amp, v_[j].y, h0.y); }                        \
		}                                                                                                               \
	}
			int k = 0;
			for (; k + 2 <= ci; k += 2) {
				const StEntry ea = er[k], eb = er[k + 1];
				double2 va, vb;
				asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(va.x), "=d"(va.y) : "r"(tile_s + (uint32_t)ea.off));
				asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(vb.x), "=d"(vb.y) : "r"(tile_s + (uint32_t)eb.off));
				h0.x = fma(ea.amp, va.x, h0.x); h0.y = fma(ea.amp, va.y, h0.y);
				h1.x = fma(eb.amp, vb.x, h1.x); h1.y = fma(eb.amp, vb.y, h1.y);
			}
			if (k < ci) {
				const StEntry ea = er[k];
				double2 va;
				asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(va.x), "=d"(va.y) : "r"(tile_s + (uint32_t)ea.off));
				h0.x = fma(ea.amp, va.x, h0.x); h0.y = fma(ea.amp, va.y, h0.y);
			}
			for (k = ci; k + NB <= kend; k += NB) DST_BATCH(NB)
			switch (kend - k) {
			case 7: DST_BATCH(7) break;
			case 6: DST_BATCH(6) break;
			case 5: DST_BATCH(5) break;
			case 4: DST_BATCH(4) break;
			case 3: DST_BATCH(3) break;
			case 2: DST_BATCH(2) break;
			case 1: DST_BATCH(1) break;
			default: break;
			}
#undef DST_BATCH
			double2 yv;
			asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(yv.x), "=d"(yv.y) : "r"(tile_s + r * (PC * 8u)));
			const word_t k2 = s_k2[r];
			const double dv2 = s_dv2[r];
			double dga, dgb;
			if (m.model == LPP_MODEL_HUBBARD && dt.uniformU) {
				dga = dt.U0 * (double)lpp_popc(k1a & k2) + dv1a + dv2;
				dgb = dt.U0 * (double)lpp_popc(k1b & k2) + dv1b + dv2;
			} else {
				dga = tiled_diag(m, dt, k1a, k2, cv.u0 + c, d);
				dgb = tiled_diag(m, dt, k1b, k2, cv.u0 + c + 1, d);
			}
			double xa = a.alpha * fma(dga, yv.x, h0.x + h1.x), xb = a.alpha * fma(dgb, yv.y, h0.y + h1.y);
			if (need_x) { xa = fma(a.beta, xv.x, xa); xb = fma(a.beta, xv.y, xb); }
			contrib += yv.x * xa + yv.y * xb;
			*reinterpret_cast<double2*>(a.x + t) = make_double2(xa, xb);
		}
	}
	if (a.dot_partials) {
		double sum = tiled_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = sum;
	}
}

// =====================================================================================================
// v1 sweeps (fallbacks: too many distinct amplitudes, or one-spin bases that cannot be blocked)
// =====================================================================================================
__global__ void __launch_bounds__(PA_COLS) k_sweep_down(ModelDev m, HopTable dn, DiagTables dt, SpmvArgs a, uint64_t d0,
                                                       uint64_t dcount, uint32_t nrowchunks)
{
	const uint32_t panel = blockIdx.x / nrowchunks, chunk = blockIdx.x % nrowchunks;
	const uint64_t u = (uint64_t)panel * PA_COLS + threadIdx.x;
	if (u >= m.n1) return;
	const double* __restrict__ y = a.y;
	const word_t k1 = m.b1[u];
	const uint64_t n1 = m.n1;
#pragma unroll 1
	for (int r = 0; r < PA_ROWS; r++) {
		const uint64_t dl = (uint64_t)chunk * PA_ROWS + r;
		if (dl >= dcount) break;
		const uint64_t d = d0 + dl;
		const int cd = (int)dn.cnt[d];
		double acc = tiled_diag(m, dt, k1, m.b2[d], u, d) * y[d * n1 + u];
#pragma unroll 4
		for (int k = 0; k < cd; k++)
			acc += dn.val[(uint64_t)k * dn.n + d] * y[(uint64_t)dn.idx[(uint64_t)k * dn.n + d] * n1 + u];
		const uint64_t t = dl * n1 + u;
		double xn = a.alpha * acc;
		if (a.beta != 0.0) xn += a.beta * a.x[t];
		a.x[t] = xn;
	}
}

__global__ void __launch_bounds__(PB_THREADS, 1) k_sweep_up_smem(ModelDev m, HopTable up, SpmvArgs a, uint64_t d0, int want_dot)
{
	extern __shared__ double ys[];
	const uint64_t dl = blockIdx.x, d = d0 + dl, n1 = m.n1;
	const double* __restrict__ yrow = a.y + d * n1;
	double* __restrict__ xrow = a.x + dl * n1;
	for (uint64_t u = threadIdx.x; u < n1; u += PB_THREADS) ys[u] = yrow[u];
	__syncthreads();
	double contrib = 0.0;
	for (uint64_t u = threadIdx.x; u < n1; u += PB_THREADS) {
		const int cu = (int)up.cnt[u];
		double acc = 0.0;
#pragma unroll 4
		for (int k = 0; k < cu; k++) acc += up.val[(uint64_t)k * n1 + u] * ys[up.idx[(uint64_t)k * n1 + u]];
		double xn = xrow[u] + a.alpha * acc;
		xrow[u] = xn;
		contrib += ys[u] * xn;
	}
	if (want_dot && a.dot_partials) {
		double s = tiled_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = s;
	}
}

__global__ void __launch_bounds__(256) k_sweep_up_global(ModelDev m, HopTable up, SpmvArgs a, uint64_t d0, uint32_t nbx,
                                                        int want_dot)
{
	const uint64_t dl = blockIdx.x / nbx, n1 = m.n1;
	const uint64_t u = (uint64_t)(blockIdx.x % nbx) * 256 + threadIdx.x;
	double contrib = 0.0;
	if (u < n1) {
		const double* __restrict__ yrow = a.y + (d0 + dl) * n1;
		const int cu = (int)up.cnt[u];
		double acc = 0.0;
#pragma unroll 4
		for (int k = 0; k < cu; k++) acc += up.val[(uint64_t)k * n1 + u] * yrow[up.idx[(uint64_t)k * n1 + u]];
		const uint64_t t = dl * n1 + u;
		double xn = a.x[t] + a.alpha * acc;
		a.x[t] = xn;
		contrib = yrow[u] * xn;
	}
	if (want_dot && a.dot_partials) {
		double s = tiled_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = s;
	}
}

struct TwoEmit {
	const double* __restrict__ y;
	uint64_t n1;
	double acc;
	__device__ void operator()(uint64_t a, uint64_t b, double v) { acc += v * y[a + b * n1]; }
};

// sweep C (FeAs only): on-site inter-orbital spin exchange and pair hopping change both spin words at once
__global__ void __launch_bounds__(256) k_sweep_twospin(ModelDev m, SpmvArgs a)
{
	const uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	double contrib = 0.0;
	if (t < a.nloc) {
		const uint64_t r = a.row0 + t;
		const uint64_t i1 = r % m.n1, i2 = r / m.n1;
		TwoEmit e{a.y, m.n1, 0.0};
		lpp_feas_twospin(m, m.b1[i1], m.b2[i2], m.u3_all_pairs, e);
		double xn = a.x[t] + a.alpha * e.acc;
		a.x[t] = xn;
		contrib = a.y[r] * xn;
	}
	if (a.dot_partials) {
		double s = tiled_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = s;
	}
}

// sweep C from per-spin tables (two orbitals): entry of (site i, one-spin state s) =
//   [27] sign of doSign(s, i, 1, i, 0) | [26] sign of doSign(s, i, 0, i, 1) | [25] orbital holding the electron | [24] site i
//   holds exactly one electron of this species | [23:0] index of the state with that electron in the other orbital.
// FeBasedSc.h:376-432,678-713 (setU2OffDiagonalTerm / setU3Term): with the up electron in orbital orb2 and orb1 = 1 - orb2,
// the down electron in orb1 gives the U2 spin exchange (U[2]/2), in orb2 the U3 pair hop (-U[3]; OTF literal: orb2 > orb1 only).
#define TS_VALID 0x01000000u
#define TS_ROWS 8              // down states per CTA: the up-table entries of a thread are loaded once for all of them
__global__ void __launch_bounds__(256) k_sweep_twospin_tab(ModelDev m, const uint32_t* __restrict__ tu, const uint32_t* __restrict__ td,
                                                          double u2half, double u3, int all_pairs, SpmvArgs a, uint64_t dcount)
{
	__shared__ uint32_t tds[TS_ROWS][32];
	const uint64_t n1 = m.n1;
	const uint64_t dl0 = (uint64_t)blockIdx.y * TS_ROWS, d0 = a.row0 / n1;
	for (int q = threadIdx.x; q < TS_ROWS * m.nsite; q += 256) {
		const int r = q / m.nsite, i = q % m.nsite;
		tds[r][i] = (dl0 + r < dcount) ? td[(uint64_t)i * m.n2 + d0 + dl0 + r] : 0u;
	}
	__syncthreads();
	const uint64_t u = (uint64_t)blockIdx.x * 256 + threadIdx.x;
	double contrib = 0.0;
	if (u < n1) {
		double acc[TS_ROWS];
#pragma unroll
		for (int r = 0; r < TS_ROWS; r++) acc[r] = 0.0;
		for (int i = 0; i < m.nsite; i++) {
			const uint32_t eu = __ldg(tu + (uint64_t)i * n1 + u);
			if (!(eu & TS_VALID)) continue;
			const uint32_t ou = (eu >> 25) & 1u;
			// orb1 = 1 - ou: doSign(., i, orb1, i, orb2) is the (0,1) sign when orb1 = 0, the (1,0) sign otherwise
			const uint32_t su = (ou ? (eu >> 26) : (eu >> 27)) & 1u;
			const double* __restrict__ ycol = a.y + (eu & 0xffffffu);
#pragma unroll
			for (int r = 0; r < TS_ROWS; r++) {
				const uint32_t ed = tds[r][i];
				if (!(ed & TS_VALID)) continue;
				const uint32_t od = (ed >> 25) & 1u;
				double coef;
				if (od != ou) coef = u2half;
				else if (all_pairs || ou == 1u) coef = -u3;
				else continue;
				const uint32_t neg = (su ^ (ou ? (ed >> 26) : (ed >> 27))) & 1u;
				acc[r] += (neg ? -coef : coef) * ycol[(uint64_t)(ed & 0xffffffu) * n1];
			}
		}
#pragma unroll
		for (int r = 0; r < TS_ROWS; r++) {
			if (dl0 + r >= dcount) continue;
			const uint64_t t = (dl0 + r) * n1 + u;
			const double xn = a.x[t] + a.alpha * acc[r];
			a.x[t] = xn;
			contrib += a.y[a.row0 + t] * xn;
		}
	}
	if (a.dot_partials) {
		double s = tiled_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[(uint64_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
	}
}

// sweep C by rows: one CTA per down state d, site loop outside, so that all gathers of one (d, site) pair fall into ONE source
// row of y (the row of d with the two orbitals of that site exchanged): 64 KB for 8 008 up states, L1 resident while the CTA
// walks it.  k_sweep_twospin_tab (256 up states x 8 down states per CTA) touches 8 different source rows per site and pulls a
// 32-byte sector from L2 for every 8-byte operand.  Thread t owns up states t, t + TSR_THREADS, ... (TSR_KC per pass).
// Experiment, off by default: parity-tested (LPP_TWOSPIN_ROWS=1) but slower on config 4 (1.6 ms against 1.2 ms).
#define TSR_THREADS 512
#define TSR_KC 16
__global__ void __launch_bounds__(TSR_THREADS) k_sweep_twospin_rows(ModelDev m, const uint32_t* __restrict__ tu, const uint32_t* __restrict__ td,
                                                                  double u2half, double u3, int all_pairs, SpmvArgs a, uint64_t dcount)
{
	__shared__ uint32_t s_ed[64];
	const uint64_t n1 = m.n1;
	const uint64_t dl = blockIdx.x, d = a.row0 / n1 + dl;
	if (threadIdx.x < (unsigned)m.nsite) s_ed[threadIdx.x] = td[(uint64_t)threadIdx.x * m.n2 + d];
	__syncthreads();
	double contrib = 0.0;
	for (uint64_t base = 0; base < n1; base += (uint64_t)TSR_THREADS * TSR_KC) {
		double acc[TSR_KC];
#pragma unroll
		for (int k = 0; k < TSR_KC; k++) acc[k] = 0.0;
		for (int i = 0; i < m.nsite; i++) {
			const uint32_t ed = s_ed[i];
			if (!(ed & TS_VALID)) continue;                                 // CTA-uniform
			const uint32_t od = (ed >> 25) & 1u;
			const double* __restrict__ yrow = a.y + (uint64_t)(ed & 0xffffffu) * n1;
			const uint32_t* __restrict__ tui = tu + (uint64_t)i * n1;
#pragma unroll
			for (int k = 0; k < TSR_KC; k++) {
				const uint64_t u = base + (uint64_t)k * TSR_THREADS + threadIdx.x;
				if (u >= n1) continue;
				const uint32_t eu = __ldg(tui + u);
				if (!(eu & TS_VALID)) continue;
				const uint32_t ou = (eu >> 25) & 1u;
				double coef;
				if (od != ou) coef = u2half;
				else if (all_pairs || ou == 1u) coef = -u3;
				else continue;
				// orb1 = 1 - ou: doSign(., i, orb1, i, orb2) is the (0,1) sign when orb1 = 0, the (1,0) sign otherwise
				const uint32_t neg = ((ou ? (eu >> 26) : (eu >> 27)) ^ (ou ? (ed >> 26) : (ed >> 27))) & 1u;
				acc[k] += (neg ? -coef : coef) * yrow[eu & 0xffffffu];
			}
		}
#pragma unroll
		for (int k = 0; k < TSR_KC; k++) {
			const uint64_t u = base + (uint64_t)k * TSR_THREADS + threadIdx.x;
			if (u >= n1) continue;
			const uint64_t t = dl * n1 + u;
			const double xn = a.x[t] + a.alpha * acc[k];
			a.x[t] = xn;
			contrib += a.y[a.row0 + t] * xn;
		}
	}
	if (a.dot_partials) {
		double sum = tiled_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = sum;
	}
}

// =====================================================================================================
// plan construction
// =====================================================================================================
template <class T>
static int plan_upload(TiledPlan* p, T** out, const std::vector<T>& v)
{
	void* q = nullptr;
	TCK(cudaMalloc(&q, std::max<size_t>(v.size(), 1) * sizeof(T)));
	p->allocs.push_back(q);
	if (!v.empty()) TCK(cudaMemcpy(q, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
	*out = (T*)q;
	return 0;
}
template <class T>
static int plan_alloc(TiledPlan* p, T** out, size_t n)
{
	void* q = nullptr;
	TCK(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
	p->allocs.push_back(q);
	*out = (T*)q;
	return 0;
}

// group the one-spin states by the occupation pattern of the top `fsites` sites; keep basis order inside a group
static void make_blocks(const std::vector<word_t>& words, int nbits_per_site, int nsite, int fsites,
                        std::vector<uint32_t>& rowlist, std::vector<uint32_t>& blk_off, std::vector<uint32_t>& local,
                        std::vector<uint32_t>& blk_of, uint32_t* max_block)
{
	const size_t n = words.size();
	const int shift = (nsite - fsites) * nbits_per_site;
	std::vector<uint64_t> key(n);
	std::vector<uint64_t> keys;
	for (size_t i = 0; i < n; i++) { key[i] = fsites ? (words[i] >> shift) : 0; keys.push_back(key[i]); }
	std::sort(keys.begin(), keys.end());
	keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
	const size_t nb = keys.size();
	std::vector<uint32_t> count(nb, 0);
	blk_of.resize(n);
	for (size_t i = 0; i < n; i++) {
		blk_of[i] = (uint32_t)(std::lower_bound(keys.begin(), keys.end(), key[i]) - keys.begin());
		count[blk_of[i]]++;
	}
	blk_off.assign(nb + 1, 0);
	for (size_t b = 0; b < nb; b++) blk_off[b + 1] = blk_off[b] + count[b];
	std::vector<uint32_t> fill(nb, 0);
	rowlist.resize(n);
	local.resize(n);
	*max_block = 0;
	for (size_t i = 0; i < n; i++) {
		uint32_t b = blk_of[i];
		local[i] = fill[b];
		rowlist[blk_off[b] + fill[b]++] = (uint32_t)i;
	}
	for (size_t b = 0; b < nb; b++) *max_block = std::max(*max_block, count[b]);
}

// Bank-conflict-free slot scheduling for the sweep-B gathers.  The G = 16/R lanes that the hardware serves together
// (quarter-warp for 16-byte, half-warp for 8-byte shared-memory loads) handle G consecutive up states; at table slot k
// they gather from G positions of the staged block.  Re-order every state's list (and pad with TE_HOLE) so that in each
// slot the G targets fall into G different bank groups (position mod G).  Amplitudes travel with their entries, so the
// result is the same sum in a different order.
static int schedule_conflict_free(TiledPlan* p, SpinPlan* sp, int G, const std::vector<uint32_t>& blk_off, cudaStream_t s)
{
	const uint64_t n = sp->n;
	const int width = sp->width;
	std::vector<uint32_t> tab((size_t)std::max(width, 1) * n), meta(n);
	TCK(cudaMemcpy(tab.data(), sp->tab, tab.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	TCK(cudaMemcpy(meta.data(), sp->meta, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	std::vector<std::vector<uint32_t>> out(n);
	int newwidth = 0;
	for (size_t b = 0; b + 1 < blk_off.size(); b++) {
		const uint32_t boff = blk_off[b], bsize = blk_off[b + 1] - boff;
		for (uint32_t base = 0; base < bsize; base += G) {
			const int nl = (int)std::min<uint32_t>(G, bsize - base);
			std::vector<std::vector<uint32_t>> rem(nl);
			for (int l = 0; l < nl; l++) {
				const uint64_t u = boff + base + l;
				const int cnt = meta[u] & 0xff, next = (meta[u] >> 8) & 0xff;
				for (int k = next; k < cnt; k++) rem[l].push_back(tab[(size_t)k * n + u]);
			}
			std::vector<int> order(nl);
			for (;;) {
				bool any = false;
				for (int l = 0; l < nl; l++) any = any || !rem[l].empty();
				if (!any) break;
				for (int l = 0; l < nl; l++) order[l] = l;
				std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return rem[x].size() > rem[y].size(); });
				uint32_t used = 0;
				for (int oi = 0; oi < nl; oi++) {
					const int l = order[oi];
					if (rem[l].empty()) continue;
					const uint64_t u = boff + base + l;
					size_t pick = rem[l].size();
					for (size_t j = 0; j < rem[l].size(); j++)
						if (!(used & (1u << ((rem[l][j] & TE_IDX) % G)))) { pick = j; break; }
					if (pick == rem[l].size()) { out[u].push_back(TE_HOLE); continue; }
					used |= 1u << ((rem[l][pick] & TE_IDX) % G);
					out[u].push_back(rem[l][pick]);
					rem[l].erase(rem[l].begin() + pick);
				}
			}
			for (int l = 0; l < nl; l++) {
				const uint64_t u = boff + base + l;
				const int next = (meta[u] >> 8) & 0xff;
				newwidth = std::max(newwidth, next + (int)out[u].size());
			}
		}
	}
	if (newwidth > 255) { g_terr = "scheduled table too wide"; return 1; }
	std::vector<uint32_t> ntab((size_t)std::max(newwidth, 1) * n, TE_HOLE), nmeta(n);
	size_t holes = 0, real = 0;
	for (uint64_t u = 0; u < n; u++) {
		const int next = (meta[u] >> 8) & 0xff;
		for (int k = 0; k < next; k++) ntab[(size_t)k * n + u] = tab[(size_t)k * n + u];
		for (size_t j = 0; j < out[u].size(); j++) {
			ntab[(size_t)(next + j) * n + u] = out[u][j];
			if (out[u][j] == TE_HOLE) holes++; else real++;
		}
		nmeta[u] = (uint32_t)(next + out[u].size()) | ((uint32_t)next << 8);
	}
	p->sched_holes = holes;
	p->sched_real = real;
	// lean table for k_sweep_up_lean: branch-free, byte offsets, holes/padding on free bank groups of the zero slots.
	// Only when no hop leaves its block (single-block up basis), which is the case whenever an up-segment fits.
	bool has_ext = false;
	for (uint64_t u = 0; u < n; u++) has_ext = has_ext || ((meta[u] >> 8) & 0xff) != 0;
	if (!has_ext && blk_off.size() == 2) {
		const int R = 16 / G;
		const uint32_t bsize = blk_off[1] - blk_off[0];
		const int WL = std::max(newwidth, 32);
		std::vector<uint32_t> lean((size_t)WL * n);
		for (uint32_t base = 0; base < bsize; base += G) {
			const int nl = (int)std::min<uint32_t>(G, bsize - base);
			for (int sidx = 0; sidx < WL; sidx++) {
				uint32_t used = 0;
				for (int l = 0; l < nl; l++) {
					const auto& o = out[base + l];
					if (sidx < (int)o.size() && o[sidx] != TE_HOLE) used |= 1u << ((o[sidx] & TE_IDX) % G);
				}
				for (int l = 0; l < nl; l++) {
					const auto& o = out[base + l];
					uint32_t e;
					if (sidx < (int)o.size() && o[sidx] != TE_HOLE) {
						e = ((o[sidx] & TE_IDX) * (uint32_t)(8 * R)) | (o[sidx] & (TE_SIGN | 0x3f000000u));
					} else {
						int g = 0;
						while (used & (1u << g)) g++;      // a free bank group always exists: at most nl-1 <= G-1 are taken
						used |= 1u << g;
						const uint32_t zslot = bsize + (uint32_t)(((g - (int)(bsize % G)) % G + G) % G);
						e = zslot * (uint32_t)(8 * R);
					}
					lean[(size_t)sidx * n + base + l] = e;
				}
			}
		}
		std::vector<uint8_t> wc((bsize + 31) / 32, 0);
		for (uint32_t uu = 0; uu < bsize; uu++) {
			int c = (int)out[uu].size();
			while (c > 0 && out[uu][c - 1] == TE_HOLE) c--;
			c = (c + 1) & ~1;
			wc[uu >> 5] = (uint8_t)std::max<int>(wc[uu >> 5], c);
		}
		if (plan_upload(p, &p->tabL, lean)) return -1;
		if (plan_upload(p, &p->wcntL, wc)) return -1;
		p->widthL = WL;
		// packed layout for k_sweep_up_packed: [chunk][group of 4 slots][lane]
		const bool e16ok = p->mt.nmag <= 2 && (size_t)bsize + G < (1u << 14);
		if (WL <= 32 || (WL <= 48 && e16ok)) {
			const uint32_t nch = (bsize + 31) / 32;
			std::vector<uint8_t> wc4(nch);
			std::vector<uint32_t> choff(nch + 1, 0);
			for (uint32_t c = 0; c < nch; c++) {
				wc4[c] = (uint8_t)((wc[c] + 3) & ~3);
				choff[c + 1] = choff[c] + wc4[c] / 4;
			}
			const bool e16 = e16ok;
			const uint32_t hole0 = bsize * (uint32_t)(8 * R);      // zero slot 0 (states past the end of the last chunk)
			std::vector<uint32_t> pk32;
			std::vector<uint16_t> pk16;
			if (e16) pk16.resize((size_t)choff[nch] * 32 * 4);
			else pk32.resize((size_t)choff[nch] * 32 * 4);
			for (uint32_t c = 0; c < nch; c++)
				for (uint32_t g = 0; g < (uint32_t)wc4[c] / 4; g++)
					for (uint32_t l = 0; l < 32; l++)
						for (uint32_t k = 0; k < 4; k++) {
							const uint32_t u = c * 32 + l, sidx = 4 * g + k;
							const uint32_t e = (u < bsize && sidx < (uint32_t)WL) ? lean[(size_t)sidx * n + u] : hole0;
							const size_t at = (((size_t)choff[c] + g) * 32 + l) * 4 + k;
							if (e16) pk16[at] = (uint16_t)(((e & TE_IDX) / (uint32_t)(8 * R)) | ((e & TE_SIGN) ? 0x8000u : 0u) | (((e >> 24) & 1u) << 14));
							else pk32[at] = e;
						}
			if (e16) { uint16_t* d = nullptr; if (plan_upload(p, &d, pk16)) return -1; p->tabP = d; }
			else { uint32_t* d = nullptr; if (plan_upload(p, &d, pk32)) return -1; p->tabP = d; }
			if (plan_upload(p, &p->choffP, choff)) return -1;
			if (plan_upload(p, &p->wcnt4P, wc4)) return -1;
			p->packedE16 = e16 ? 1 : 0;
			p->packed_mean_slots = 4.0 * choff[nch] / std::max<uint32_t>(nch, 1);
			p->packedNG = (WL <= 32) ? 8 : 12;
		}
	}
	uint32_t* dtab = nullptr;
	if (plan_upload(p, &dtab, ntab)) return -1;
	TCK(cudaMemcpyAsync(sp->meta, nmeta.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
	TCK(cudaStreamSynchronize(s));
	sp->tab = dtab;
	sp->width = newwidth;
	return 0;
}

static int build_spin_plan(TiledPlan* p, SpinPlan* sp, const HopTable& t, const std::vector<word_t>& words, int bits_per_site,
                           int nsite, uint32_t cap, bool need_contiguous, bool row_major, int sched_group, cudaStream_t s)
{
	std::vector<uint32_t> rowlist, blk_off, local, blk_of;
	uint32_t mx = 0;
	int f = 0;
	for (; f <= nsite; f++) {
		make_blocks(words, bits_per_site, nsite, f, rowlist, blk_off, local, blk_of, &mx);
		if (mx <= cap) break;
	}
	if (mx > cap) { g_terr = "cannot block the one-spin basis"; return 1; }
	if (need_contiguous) {
		for (size_t i = 0; i < rowlist.size(); i++)
			if (rowlist[i] != i) { g_terr = "up blocks are not contiguous"; return 1; }
	}
	sp->n = t.n;
	sp->width = t.width;
	sp->nblocks = (int)blk_off.size() - 1;
	sp->max_block = mx;
	if (plan_upload(p, &sp->rowlist, rowlist)) return -1;
	if (plan_upload(p, &sp->blk_off, blk_off)) return -1;
	if (plan_upload(p, &sp->local, local)) return -1;
	if (plan_upload(p, &sp->blk_of, blk_of)) return -1;
	if (plan_alloc(p, &sp->tab, (size_t)std::max(t.width, 1) * t.n)) return -1;
	if (plan_alloc(p, &sp->meta, (size_t)t.n)) return -1;
	int* bad = nullptr;
	if (plan_alloc(p, &bad, 1)) return -1;
	TCK(cudaMemsetAsync(bad, 0, sizeof(int), s));
	k_compress<<<(unsigned)((t.n + 255) / 256), 256, 0, s>>>(t, p->mt, sp->blk_of, sp->local, sp->tab, sp->meta, row_major ? 1 : 0, bad);
	int hbad = 0;
	TCK(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
	TCK(cudaStreamSynchronize(s));
	TCK(cudaGetLastError());
	if (hbad) { g_terr = "hop amplitude not in the magnitude table"; return 1; }
	if (sched_group > 0 && !row_major) return schedule_conflict_free(p, sp, sched_group, blk_off, s);
	return 0;
}

int lpp_tiled_create(const ModelDev& m, const double* hop_host, const HopTable& up, const HopTable& dn, const DiagTables& dt,
                     uint64_t row0,
                     uint64_t nloc, cudaStream_t s, TiledPlan** out)
{
	if (m.model == LPP_MODEL_HEISENBERG) { g_terr = "tiled path is for product bases"; return -1; }
	TiledPlan* p = new TiledPlan();
	p->d0 = row0 / m.n1;
	p->dcount = nloc / m.n1;
	p->nloc = nloc;
	p->has_twospin = (m.model == LPP_MODEL_FEAS) ? 1 : 0;
	int dev = 0, maxsm = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);

	// distinct hop magnitudes
	p->mt.nmag = 0;
	bool mags_ok = true;
	for (int i = 0; i < m.nbits * m.nbits; i++) {
		double av = fabs(hop_host[i]);
		if (av == 0) continue;
		bool found = false;
		for (int q = 0; q < p->mt.nmag; q++) found = found || p->mt.mag[q] == av;
		if (found) continue;
		if (p->mt.nmag == LPP_MAXMAG) { mags_ok = false; break; }
		p->mt.mag[p->mt.nmag++] = av;
	}
	const char* env = getenv("LPP_TILED_V1");
	bool try_v2 = mags_ok && up.width <= 255 && dn.width <= 255 && up.n < (1u << 24) && dn.n < (1u << 24) && !(env && env[0] == '1');
	const char* envw = getenv("LPP_TILED_W");
	p->W = envw ? atoi(envw) : 16;
	if (p->W != 8 && p->W != 16 && p->W != 32) p->W = 16;
	if (dn.width > 4 * p->W) try_v2 = false;
	const char* envr = getenv("LPP_TILED_R");
	p->R = envr ? atoi(envr) : 2;
	if (p->R != 1 && p->R != 2) p->R = 2;
	const char* envs = getenv("LPP_TILED_SCHED");
	const bool sched = !(envs && envs[0] == '0');
	const char* enva = getenv("LPP_TILED_A");
	p->blocksA = enva ? atoi(enva) : 0;
	const char* envb = getenv("LPP_TILED_B");
	p->pipeB = envb ? atoi(envb) : 1;
	const char* envn = getenv("LPP_TILED_NE");
	p->NE = envn ? atoi(envn) : 24;
	if (p->NE != 16 && p->NE != 24 && p->NE != 32) p->NE = 24;
	const char* envt = getenv("LPP_TILED_TA");
	p->threadsA = (envt && atoi(envt) == 512) ? 512 : 1024;
	const char* envk = getenv("LPP_TILED_SMEMA_KB");
	const size_t smemA_budget = envk ? (size_t)atoi(envk) * 1024 : (size_t)maxsm - 2048;
	if (try_v2) {
		std::vector<word_t> w1(m.n1), w2(m.n2);
		if (cudaMemcpy(w1.data(), m.b1, sizeof(word_t) * m.n1, cudaMemcpyDeviceToHost) != cudaSuccess ||
		    cudaMemcpy(w2.data(), m.b2, sizeof(word_t) * m.n2, cudaMemcpyDeviceToHost) != cudaSuccess) {
			g_terr = "basis download failed";
			delete p;
			return -1;
		}
		const size_t budget = (size_t)maxsm - 2048;
		uint32_t capA = (uint32_t)(std::min(budget, smemA_budget) / ((size_t)p->W * 8 + 8));
		uint32_t capB = (uint32_t)(budget / ((size_t)p->R * 8));
		int ra = build_spin_plan(p, &p->dn, dn, w2, m.orbitals, m.nsite, capA, false, true, 0, s);
		int rb = ra == 0 ? build_spin_plan(p, &p->up, up, w1, m.orbitals, m.nsite, capB, true, false, sched ? 16 / p->R : 0, s) : ra;
		if (ra < 0 || rb < 0) { delete p; return -1; }
		if (ra == 0 && rb == 0) {
			p->v2 = 1;
			p->smemA = (size_t)p->dn.max_block * p->W * 8;
			p->smemB = (size_t)p->up.max_block * p->R * 8;
			p->smemA3 = p->smemA + (size_t)p->dn.max_block * 8;
			// blocks of down states that contain at least one local row
			std::vector<uint32_t> rowlist(m.n2), blk_off(p->dn.nblocks + 1);
			cudaMemcpy(rowlist.data(), p->dn.rowlist, sizeof(uint32_t) * m.n2, cudaMemcpyDeviceToHost);
			cudaMemcpy(blk_off.data(), p->dn.blk_off, sizeof(uint32_t) * blk_off.size(), cudaMemcpyDeviceToHost);
			for (int b = 0; b < p->dn.nblocks; b++) {
				bool hit = false;
				for (uint32_t i = blk_off[b]; i < blk_off[b + 1] && !hit; i++)
					hit = rowlist[i] >= p->d0 && rowlist[i] < p->d0 + p->dcount;
				if (hit) p->tilesA_host.push_back((uint32_t)b);
			}
			if (plan_upload(p, &p->tilesA, p->tilesA_host)) { delete p; return -1; }
			p->ntilesA_blocks = (uint32_t)p->tilesA_host.size();
			p->npanels = (uint32_t)((m.n1 + p->W - 1) / p->W);
			cudaError_t e1 = cudaSuccess;
			if (p->W == 8) e1 = lpp_raise_smem(k_sweep_down_blocks<8>, (size_t)(p->smemA));
			if (p->W == 16) e1 = lpp_raise_smem(k_sweep_down_blocks<16>, (size_t)(p->smemA));
			if (p->W == 32) e1 = lpp_raise_smem(k_sweep_down_blocks<32>, (size_t)(p->smemA));
			cudaError_t e2 = p->R == 2
			                     ? lpp_raise_smem(k_sweep_up_blocks<2>, (size_t)(p->smemB))
			                     : lpp_raise_smem(k_sweep_up_blocks<1>, (size_t)(p->smemB));
			cudaError_t e3 = cudaSuccess;
#define SETB(R_, N_) if (e3 == cudaSuccess) e3 = lpp_raise_smem(k_sweep_up_pipe<R_, N_>, (size_t)(p->smemB))
			SETB(1, 16); SETB(1, 24); SETB(1, 32); SETB(2, 16); SETB(2, 24); SETB(2, 32);
#undef SETB
#define SETA3(W_, T_, B_) if (e3 == cudaSuccess) e3 = lpp_raise_smem(k_sweep_down_blocks3<W_, T_, B_>, (size_t)(p->smemA3))
			SETA3(8, 1024, 1); SETA3(16, 1024, 1); SETA3(32, 1024, 1); SETA3(8, 512, 2); SETA3(16, 512, 2); SETA3(32, 512, 2);
#undef SETA3
			if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { g_terr = "cudaFuncSetAttribute(smem) failed"; delete p; return -1; }
		}
	}
	{
		p->nrowchunks = (uint32_t)((p->dcount + PA_ROWS - 1) / PA_ROWS);
		p->npanels_v1 = (uint32_t)((m.n1 + PA_COLS - 1) / PA_COLS);
		p->up_smem_bytes = (size_t)m.n1 * sizeof(double);
		p->up_in_smem = (p->up_smem_bytes + 1024 <= (size_t)maxsm) ? 1 : 0;
		if (p->up_in_smem && !p->v2) {
			cudaError_t e = lpp_raise_smem(k_sweep_up_smem, (size_t)(p->up_smem_bytes));
			if (e != cudaSuccess) { g_terr = cudaGetErrorString(e); delete p; return -1; }
		}
	}
	{
		const char* envl = getenv("LPP_TILED_LEAN");      // bit 0: lean sweep A, bit 1: lean sweep B (default both)
		const int lean = envl ? atoi(envl) : 3;
		p->leanA = (lean & 1) && (p->blocksA == 0);
		p->smemAL = (size_t)PAL_ROWS * std::max(dn.width, 1) * sizeof(DownEntry);
		p->leanB = (lean & 2) && p->v2 && p->tabL != nullptr && p->widthL <= 32 && p->up.nblocks == 1;
		if (p->leanB) {
			p->smemBL = ((size_t)m.n1 + 16 / p->R) * p->R * 8;
			if (p->smemBL + 1024 > (size_t)maxsm) p->leanB = 0;
		}
		cudaError_t e = cudaSuccess;
		if (p->leanA && p->smemAL > 48 * 1024) {
			e = lpp_raise_smem(k_sweep_down_lean<1>, (size_t)(p->smemAL));
			if (e == cudaSuccess) e = lpp_raise_smem(k_sweep_down_lean<2>, (size_t)(p->smemAL));
		}
#define SETL(R_, U_) if (e == cudaSuccess) e = lpp_raise_smem(k_sweep_up_lean<R_, 32, U_>, (size_t)(p->smemBL))
		if (p->leanB) { SETL(1, true); SETL(1, false); SETL(2, true); SETL(2, false); }
#undef SETL
		const char* envp = getenv("LPP_TILED_PACKED");
		p->smemBP = ((size_t)m.n1 + 16 / p->R) * p->R * 8 + ((size_t)(m.n1 + 31) / 32) * 5 + 16;
		p->packedB = p->v2 && p->R == 2 && p->up.nblocks == 1 && p->tabP != nullptr && !(envp && envp[0] == '0') && (lean & 2) &&
		             p->smemBP + 1024 <= (size_t)maxsm;
#define SETP(U_, E_, NG_) if (e == cudaSuccess) e = lpp_raise_smem(k_sweep_up_packed<2, U_, E_, NG_>, (size_t)(p->smemBP))
		if (p->packedB) { SETP(true, true, 8); SETP(false, true, 8); SETP(true, true, 12); SETP(false, true, 12); SETP(true, false, 8); SETP(false, false, 8); }
#undef SETP
		if (e != cudaSuccess) { g_terr = cudaGetErrorString(e); delete p; return -1; }
	}
	// staged down sweep: runs of consecutive down states sharing the occupation of the sites >= split
	if (mags_ok && p->leanA && dn.n < (1u << 24) && dn.width >= 1 && dn.width <= 255 && (getenv("LPP_DSTAGE") && getenv("LPP_DSTAGE")[0] == '1')) {
		const uint64_t n2 = dn.n;
		const int W = dn.width;
		std::vector<uint32_t> hidx((size_t)W * n2), hcnt(n2);
		std::vector<double> hval((size_t)W * n2);
		std::vector<word_t> w2(n2);
		if (cudaMemcpy(hidx.data(), dn.idx, hidx.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess ||
		    cudaMemcpy(hval.data(), dn.val, hval.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
		    cudaMemcpy(hcnt.data(), dn.cnt, n2 * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess ||
		    cudaMemcpy(w2.data(), m.b2, n2 * sizeof(word_t), cudaMemcpyDeviceToHost) != cudaSuccess) {
			g_terr = "down table download failed";
			delete p;
			return -1;
		}
		const char* envr = getenv("LPP_DSTAGE_ROWS");
		const char* envc = getenv("LPP_DSTAGE_PC");
		const uint32_t cap = envr ? (uint32_t)atoi(envr) : 72u;
		p->stPC = (envc && atoi(envc) == 128) ? 128 : 64;
		// largest split whose longest run fits the cap
		int split = -1;
		std::vector<uint32_t> off;
		for (int sp = m.nbits; sp >= 1 && split < 0; sp--) {
			std::vector<uint32_t> o(1, 0u);
			uint32_t mx = 0;
			for (uint64_t d = 1; d <= n2; d++)
				if (d == n2 || (w2[d] >> sp) != (w2[d - 1] >> sp)) { mx = std::max<uint32_t>(mx, (uint32_t)d - o.back()); o.push_back((uint32_t)d); }
			if (mx <= cap) { split = sp; off.swap(o); p->stMaxRows = mx; }
		}
		if (split >= 1) {
			const int WP = (W + 3) & ~3;
			std::vector<uint32_t> tab((size_t)n2 * WP);
			std::vector<uint16_t> cnt(n2);
			size_t nint = 0, ntot = 0;
			bool ok = true;
			for (size_t b = 0; b + 1 < off.size() && ok; b++)
				for (uint32_t d = off[b]; d < off[b + 1] && ok; d++) {
					uint32_t* row = tab.data() + (size_t)d * WP;
					int ci = 0, ce = 0;
					uint32_t ext[256];
					for (uint32_t k = 0; k < hcnt[d]; k++) {
						const uint32_t src = hidx[(size_t)k * n2 + d];
						const double v = hval[(size_t)k * n2 + d];
						int mi = -1;
						for (int q = 0; q < p->mt.nmag; q++) if (p->mt.mag[q] == fabs(v)) mi = q;
						if (mi < 0) { ok = false; break; }
						const uint32_t e = ((uint32_t)mi << 24) | (v < 0 ? TE_SIGN : 0u);
						if (src >= off[b] && src < off[b + 1]) row[ci++] = e | (src - off[b]);
						else ext[ce++] = e | src;
					}
					for (int k = 0; k < ce; k++) row[ci + k] = ext[k];
					for (int k = ci + ce; k < WP; k++) row[k] = d;            // padding: never dereferenced past the counts
					cnt[d] = (uint16_t)(ci | (ce << 8));
					nint += ci;
					ntot += ci + ce;
				}
			p->stInternal = ntot ? (double)nint / (double)ntot : 0.0;
			p->smemST = (size_t)p->stMaxRows * p->stPC * 8 + (size_t)p->stMaxRows * WP * 16 + (size_t)p->stMaxRows * 20 + 16;
			if (ok && p->stInternal >= 0.15 && p->smemST + 1024 <= (size_t)maxsm) {
				if (plan_upload(p, &p->stTab, tab) || plan_upload(p, &p->stCnt, cnt) || plan_upload(p, &p->stOff, off)) { delete p; return -1; }
				p->stOff_host = off;
				p->stNblk = (uint32_t)off.size() - 1;
				p->stWidth = WP;
				p->stSplit = split;
				cudaError_t e = cudaSuccess;
#define SETS(PC_, NB_) if (e == cudaSuccess) e = lpp_raise_smem(k_sweep_down_staged<PC_, NB_>, (size_t)(p->smemST))
				SETS(64, 4); SETS(64, 8); SETS(128, 4); SETS(128, 8);
#undef SETS
				{ const char* envn = getenv("LPP_DSTAGE_NB"); p->stNR = (envn && atoi(envn) == 4) ? 4 : 8; }
				if (e != cudaSuccess) { g_terr = cudaGetErrorString(e); delete p; return -1; }
				p->stagedA = 1;
			}
		}
		if (getenv("LPP_VERBOSE"))
			fprintf(stderr, "[lpp tiled] staged down sweep: on=%d split=%d runs=%u max rows=%u in-run operands=%.3f smem=%zu PC=%d\n", p->stagedA,
			        split, p->stNblk, p->stMaxRows, p->stInternal, p->smemST, p->stPC);
	}
	if (mags_ok) {
		int rd = lpp_dtile_create(m, dn, dt.dv2, p->mt, s, &p->dtile);
		if (rd < 0) { g_terr = std::string("down tile plan: ") + lpp_dtile_error(); delete p; return -1; }
		if (rd > 0 && getenv("LPP_VERBOSE")) fprintf(stderr, "[lpp tiled] down tile kernel not used: %s\n", lpp_dtile_error());
		rd = lpp_drows_create(m, dn, p->mt, s, &p->drows);
		if (rd < 0) { g_terr = std::string("down rows plan: ") + lpp_dtile_error(); delete p; return -1; }
		if (rd > 0 && getenv("LPP_VERBOSE")) fprintf(stderr, "[lpp tiled] row-walking down kernel not used: %s\n", lpp_dtile_error());
	}
	if (mags_ok && p->leanA && !p->dtile && !p->drows && !p->stagedA) {
		int rb = lpp_dblock_create(m, dn, dt, s, &p->dblock);
		if (rb < 0) { g_terr = std::string("down block plan: ") + lpp_dblock_error(); delete p; return -1; }
		if (getenv("LPP_VERBOSE")) {
			char buf[640] = "";
			if (p->dblock) lpp_dblock_describe(p->dblock, buf, sizeof(buf));
			fprintf(stderr, "[lpp tiled] block down sweep: %s\n", p->dblock ? buf : lpp_dblock_error());
		}
	}
	if (getenv("LPP_VERBOSE"))
		fprintf(stderr, "[lpp tiled] leanA=%d leanB=%d widthL=%d dtile=%d packedB=%d e16=%d mean slots/warp=%.2f\n", p->leanA, p->leanB, p->widthL,
		        p->dtile ? 1 : 0, p->packedB, p->packedE16, p->packed_mean_slots);
	if (getenv("LPP_VERBOSE"))
		fprintf(stderr, "[lpp tiled] v2=%d blocksA=%d TA=%d W=%d R=%d pipeB=%d NE=%d dn: blocks=%d max=%u width=%d | up: blocks=%d max=%u width=%d sched real=%zu holes=%zu\n",
		        p->v2, p->blocksA, p->threadsA, p->W, p->R, p->pipeB, p->NE, p->dn.nblocks, p->dn.max_block, p->dn.width, p->up.nblocks, p->up.max_block,
		        p->up.width, p->sched_real, p->sched_holes);
	if (p->has_twospin && m.orbitals == 2 && m.nsite <= 32 && m.n1 < (1u << 24) && m.n2 < (1u << 24) && p->dcount <= 65535ull * TS_ROWS && !(getenv("LPP_TWOSPIN_TAB") && getenv("LPP_TWOSPIN_TAB")[0] == '0')) {
		for (int spin = 0; spin < 2; spin++) {
			const uint64_t n = spin ? m.n2 : m.n1;
			std::vector<word_t> w(n);
			if (cudaMemcpy(w.data(), spin ? m.b2 : m.b1, sizeof(word_t) * n, cudaMemcpyDeviceToHost) != cudaSuccess) { g_terr = "basis download failed"; delete p; return -1; }
			std::vector<std::pair<word_t, uint32_t>> byword(n);
			for (uint64_t i = 0; i < n; i++) byword[i] = {w[i], (uint32_t)i};
			std::sort(byword.begin(), byword.end());
			std::vector<uint32_t> tab((size_t)m.nsite * n, 0u);
			for (int site = 0; site < m.nsite; site++)
				for (uint64_t i = 0; i < n; i++) {
					const int o0 = lpp_feas_occ(w[i], site, 0, 2), o1 = lpp_feas_occ(w[i], site, 1, 2);
					if (o0 + o1 != 1) continue;
					const word_t partner = w[i] ^ (lpp_bit(site * 2) | lpp_bit(site * 2 + 1));
					auto it = std::lower_bound(byword.begin(), byword.end(), std::make_pair(partner, (uint32_t)0));
					if (it == byword.end() || it->first != partner) { g_terr = "two-spin table: partner state not in the basis"; delete p; return -1; }
					uint32_t e = it->second | TS_VALID | (o1 ? (1u << 25) : 0u);
					if (lpp_feas_dosign(w[i], site, 0, site, 1, 2) < 0) e |= 1u << 26;
					if (lpp_feas_dosign(w[i], site, 1, site, 0, 2) < 0) e |= 1u << 27;
					tab[(size_t)site * n + i] = e;
				}
			if (plan_upload(p, spin ? &p->ts_dn : &p->ts_up, tab)) { delete p; return -1; }
		}
		double Uh[4] = {0, 0, 0, 0};
		if (cudaMemcpy(Uh, m.U, sizeof(double) * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { g_terr = "U download failed"; delete p; return -1; }
		p->u2half = 0.5 * Uh[2];
		p->u3 = Uh[3];
	}
	// the last sweep owns the dot-product partial sums
	// opt-in (LPP_TWOSPIN_ROWS=1): measured 1.6 ms against 1.2 ms for k_sweep_twospin_tab on config 4
	{ const char* envt = getenv("LPP_TWOSPIN_ROWS"); p->tsRows = (envt && envt[0] == '1') && m.nsite <= 64; }
	if (p->has_twospin && p->ts_up && p->tsRows) p->dot_blocks = (int)p->dcount;
	else if (p->has_twospin && p->ts_up) p->dot_blocks = (int)(((m.n1 + 255) / 256) * ((p->dcount + TS_ROWS - 1) / TS_ROWS));
	else if (p->has_twospin) p->dot_blocks = (int)((nloc + 255) / 256);
	else if (p->v2 && (p->leanB || p->packedB)) p->dot_blocks = (int)((p->dcount + p->R - 1) / p->R);
	else if (p->v2) p->dot_blocks = (int)(((p->dcount + p->R - 1) / p->R) * p->up.nblocks);
	else if (p->up_in_smem) p->dot_blocks = (int)p->dcount;
	else p->dot_blocks = (int)(((m.n1 + 255) / 256) * p->dcount);
	*out = p;
	return 0;
}

void lpp_tiled_destroy(TiledPlan* p)
{
	if (!p) return;
	for (void* q : p->allocs) cudaFree(q);
	lpp_dtile_destroy(p->dtile);
	lpp_drows_destroy(p->drows);
	lpp_dblock_destroy(p->dblock);
	delete p;
}


// staged down sweep launcher: runs that intersect the local rows [d0, d0 + dcount)
static inline bool staged_accepts(const TiledPlan* p, const ColView& cv) { return p->stagedA && cv.pitch % 2 == 0 && cv.ncols % 2 == 0; }
static inline void staged_range(const TiledPlan* p, uint64_t d0, uint64_t dcount, uint32_t* blk0, uint32_t* nblk)
{
	const std::vector<uint32_t>& off = p->stOff_host;
	uint32_t a = (uint32_t)(std::upper_bound(off.begin(), off.end(), (uint32_t)d0) - off.begin()) - 1;
	uint32_t b = (uint32_t)(std::lower_bound(off.begin(), off.end(), (uint32_t)(d0 + dcount)) - off.begin());
	*blk0 = a;
	*nblk = (dcount == 0 || b <= a) ? 0u : b - a;
}
static inline uint32_t staged_grid(const TiledPlan* p, uint64_t d0, uint64_t dcount, const ColView& cv)
{
	uint32_t blk0, nblk;
	staged_range(p, d0, dcount, &blk0, &nblk);
	return (uint32_t)((cv.ncols + p->stPC - 1) / p->stPC) * nblk;
}
static int staged_launch(TiledPlan* p, const ModelDev& m, const DiagTables& dt, const SpmvArgs& a, uint64_t d0, uint64_t dcount, const ColView& cv,
                         cudaStream_t s)
{
	uint32_t blk0, nblk;
	staged_range(p, d0, dcount, &blk0, &nblk);
	const uint32_t grid = (uint32_t)((cv.ncols + p->stPC - 1) / p->stPC) * nblk;
	if (grid == 0) return 0;
#define RUNS(PC_, NB_) k_sweep_down_staged<PC_, NB_><<<grid, DST_THREADS, p->smemST, s>>>(m, p->stTab, p->stCnt, p->stOff, blk0, nblk, p->stWidth, p->stMaxRows, p->mt, dt, a, d0, dcount, cv)
	if (p->stNR == 4) { if (p->stPC == 128) RUNS(128, 4); else RUNS(64, 4); }
	else { if (p->stPC == 128) RUNS(128, 8); else RUNS(64, 8); }
#undef RUNS
	return 0;
}

int lpp_tiled_dot_blocks(const TiledPlan* p) { return p->dot_blocks; }

int lpp_tiled_spmv(TiledPlan* p, const ModelDev& m, const HopTable& up, const HopTable& dn, const DiagTables& dt,
                   const SpmvArgs& a, cudaStream_t s)
{
	int launches = 0;
	const int dot_in_b = p->has_twospin ? 0 : 1;
	const ColView cvfull{m.n1, m.n1, 0};
	if (p->dblock && p->blocksA == 0 && lpp_dblock_accepts(p->dblock, m, dt, p->d0, p->dcount, cvfull)) {
		SpmvArgs aa = a;
		aa.dot_partials = nullptr;
		if (lpp_dblock_sweep(p->dblock, m, dt, aa, cvfull, s) < 0) { g_terr = lpp_dblock_error(); return -1; }
	} else if (p->drows && !p->dtile && p->blocksA == 0 && lpp_drows_accepts(p->drows, cvfull)) {
		SpmvArgs aa = a;
		aa.dot_partials = nullptr;
		if (lpp_drows_sweep(p->drows, m, dt, aa, p->d0, p->dcount, cvfull, s) < 0) { g_terr = lpp_dtile_error(); return -1; }
	} else if (p->dtile && p->blocksA == 0 && lpp_dtile_accepts(p->dtile, cvfull)) {
		SpmvArgs aa = a;
		aa.dot_partials = nullptr;
		if (lpp_dtile_sweep(p->dtile, m, dt, aa, p->d0, p->dcount, cvfull, s) < 0) { g_terr = lpp_dtile_error(); return -1; }
	} else if (p->v2 && p->blocksA == 2) {
		const unsigned gridA = p->ntilesA_blocks * p->npanels;
#define RUNA3(W_, T_, B_) k_sweep_down_blocks3<W_, T_, B_><<<gridA, T_, p->smemA3, s>>>(m, p->dn, p->mt, dt, a, p->d0, p->dcount, p->tilesA, p->ntilesA_blocks, p->dn.max_block)
		if (p->threadsA == 512) {
			if (p->W == 8) RUNA3(8, 512, 2); else if (p->W == 16) RUNA3(16, 512, 2); else RUNA3(32, 512, 2);
		} else {
			if (p->W == 8) RUNA3(8, 1024, 1); else if (p->W == 16) RUNA3(16, 1024, 1); else RUNA3(32, 1024, 1);
		}
#undef RUNA3
	} else if (p->v2 && p->blocksA == 1) {
		const unsigned gridA = p->ntilesA_blocks * p->npanels;
		if (p->W == 8)
			k_sweep_down_blocks<8><<<gridA, TA_THREADS, p->smemA, s>>>(m, p->dn, p->mt, dt, a, p->d0, p->dcount, p->tilesA, p->ntilesA_blocks);
		else if (p->W == 16)
			k_sweep_down_blocks<16><<<gridA, TA_THREADS, p->smemA, s>>>(m, p->dn, p->mt, dt, a, p->d0, p->dcount, p->tilesA, p->ntilesA_blocks);
		else
			k_sweep_down_blocks<32><<<gridA, TA_THREADS, p->smemA, s>>>(m, p->dn, p->mt, dt, a, p->d0, p->dcount, p->tilesA, p->ntilesA_blocks);
	} else if (p->leanA && staged_accepts(p, cvfull)) {
		SpmvArgs aa = a;
		aa.dot_partials = nullptr;
		staged_launch(p, m, dt, aa, p->d0, p->dcount, cvfull, s);
	} else if (p->leanA) {
		const uint32_t nchunks = (uint32_t)((p->dcount + PAL_ROWS - 1) / PAL_ROWS);
		SpmvArgs aa = a;
		aa.dot_partials = nullptr;
		ColView cv{m.n1, m.n1, 0};
		const uint32_t npan = down_lean_panels(cv);
		if (down_lean_vec(cv) == 2) k_sweep_down_lean<2><<<npan * nchunks, PAL_COLS, p->smemAL, s>>>(m, dn, dt, aa, p->d0, p->dcount, nchunks, cv);
		else k_sweep_down_lean<1><<<npan * nchunks, PAL_COLS, p->smemAL, s>>>(m, dn, dt, aa, p->d0, p->dcount, nchunks, cv);
	} else {
		uint64_t nblkA = (uint64_t)p->npanels_v1 * p->nrowchunks;
		k_sweep_down<<<(unsigned)nblkA, PA_COLS, 0, s>>>(m, dn, dt, a, p->d0, p->dcount, p->nrowchunks);
	}
	launches++;
	if (p->v2 && (p->leanB || p->packedB)) {
		const unsigned gridL = (unsigned)((p->dcount + p->R - 1) / p->R);
		const uint32_t bsz = (uint32_t)m.n1;
		SpmvArgs ab = a;
		ab.beta = 1.0;                                  // sweep A already applied beta
#define RUNL(R_, U_) k_sweep_up_lean<R_, 32, U_><<<gridL, PBP_THREADS, p->smemBL, s>>>(m, p->tabL, p->wcntL, 0u, bsz, p->mt, ab, p->d0, p->dcount, dot_in_b)
#define RUNP(U_, E_, NG_) k_sweep_up_packed<2, U_, E_, NG_><<<gridL, UPP_THREADS, p->smemBP, s>>>(m, p->tabP, p->choffP, p->wcnt4P, bsz, p->mt, ab, p->d0, p->dcount, dot_in_b)
		if (p->packedB) {
			const bool uni = p->mt.nmag == 1;
			if (p->packedE16 && p->packedNG == 12) { if (uni) RUNP(true, true, 12); else RUNP(false, true, 12); }
			else if (p->packedE16) { if (uni) RUNP(true, true, 8); else RUNP(false, true, 8); }
			else { if (uni) RUNP(true, false, 8); else RUNP(false, false, 8); }
		} else if (p->R == 2) { if (p->mt.nmag == 1) RUNL(2, true); else RUNL(2, false); }
		else { if (p->mt.nmag == 1) RUNL(1, true); else RUNL(1, false); }
#undef RUNP
#undef RUNL
	} else if (p->v2) {
		const unsigned gridB = (unsigned)(((p->dcount + p->R - 1) / p->R) * p->up.nblocks);
		SpmvArgs ab = a;
		ab.beta = 1.0;                                  // sweep A already applied beta
#define RUNB(R_, N_) k_sweep_up_pipe<R_, N_><<<gridB, PBP_THREADS, p->smemB, s>>>(m, p->up, p->mt, ab, p->d0, p->dcount, dot_in_b)
		if (p->pipeB) {
			if (p->R == 2) { if (p->NE == 16) RUNB(2, 16); else if (p->NE == 24) RUNB(2, 24); else RUNB(2, 32); }
			else { if (p->NE == 16) RUNB(1, 16); else if (p->NE == 24) RUNB(1, 24); else RUNB(1, 32); }
		} else if (p->R == 2) k_sweep_up_blocks<2><<<gridB, PB_THREADS, p->smemB, s>>>(m, p->up, p->mt, ab, p->d0, p->dcount, dot_in_b);
		else k_sweep_up_blocks<1><<<gridB, PB_THREADS, p->smemB, s>>>(m, p->up, p->mt, ab, p->d0, p->dcount, dot_in_b);
#undef RUNB
	} else if (p->up_in_smem) {
		k_sweep_up_smem<<<(unsigned)p->dcount, PB_THREADS, p->up_smem_bytes, s>>>(m, up, a, p->d0, dot_in_b);
	} else {
		uint32_t nbx = (uint32_t)((m.n1 + 255) / 256);
		k_sweep_up_global<<<(unsigned)(nbx * p->dcount), 256, 0, s>>>(m, up, a, p->d0, nbx, dot_in_b);
	}
	launches++;
	if (p->has_twospin && p->ts_up && p->tsRows) {
		k_sweep_twospin_rows<<<(unsigned)p->dcount, TSR_THREADS, 0, s>>>(m, p->ts_up, p->ts_dn, p->u2half, p->u3, m.u3_all_pairs, a, p->dcount);
		launches++;
	} else if (p->has_twospin && p->ts_up) {
		// m.U lives on the device; the two couplings were read once at plan creation
		const dim3 g((unsigned)((m.n1 + 255) / 256), (unsigned)((p->dcount + TS_ROWS - 1) / TS_ROWS), 1);
		k_sweep_twospin_tab<<<g, 256, 0, s>>>(m, p->ts_up, p->ts_dn, p->u2half, p->u3, m.u3_all_pairs, a, p->dcount);
		launches++;
	} else if (p->has_twospin) {
		k_sweep_twospin<<<(unsigned)((a.nloc + 255) / 256), 256, 0, s>>>(m, a);
		launches++;
	}
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { g_terr = cudaGetErrorString(e); return -1; }
	return launches;
}

// ---------------------------------------------------------------------------------------------------------------
// two-layout multi-GPU entry points: the up sweep runs on the rank's ROW shard (all up states, local down range),
// the down sweep (+ diagonal) on the rank's COLUMN shard (all down states, local up range).
// ---------------------------------------------------------------------------------------------------------------
// any v2 up sweep works on a row shard: the one-block kernels (lean / packed) and the blocked ones (an up segment larger than
// shared memory, e.g. 48 620 up states of the 18-site chain)
// returns 1 for the one-block kernels, 2 for the blocked ones (the caller decides: with peer memory the blocked two-layout
// run of config 5 takes 17.6 ms per iteration against 41.9 ms for the gather scheme; over NCCL send/recv it is slower, 54.7 ms
// against 34.8 ms), 0 when the plan cannot be sharded that way
int lpp_tiled_two_layout_ok(const TiledPlan* p)
{
	static const bool blocked_ok = !(getenv("LPP_TWO_LAYOUT_BLOCKED") && getenv("LPP_TWO_LAYOUT_BLOCKED")[0] == '0');
	if (!p->v2 || p->has_twospin) return 0;
	if (p->leanB || p->packedB) return 1;
	return blocked_ok ? 2 : 0;
}

int lpp_tiled_up_rows_blocks(const TiledPlan* p, uint64_t nrows)
{
	const uint64_t pairs = (nrows + p->R - 1) / p->R;
	return (int)((p->leanB || p->packedB) ? pairs : pairs * p->up.nblocks);
}

// x = beta x + alpha (T_up (x) 1) y on `nrows` local rows; x, y are the local row blocks (row r at r*Nup)
int lpp_tiled_sweep_up_rows(TiledPlan* p, const ModelDev& m, const SpmvArgs& a, uint64_t nrows, cudaStream_t s)
{
	const unsigned gridL = (unsigned)lpp_tiled_up_rows_blocks(p, nrows);
	const uint32_t bsz = (uint32_t)m.n1;
	const int want_dot = a.dot_partials ? 1 : 0;
#define RUNL(R_, U_) k_sweep_up_lean<R_, 32, U_><<<gridL, PBP_THREADS, p->smemBL, s>>>(m, p->tabL, p->wcntL, 0u, bsz, p->mt, a, 0, nrows, want_dot)
#define RUNP(U_, E_, NG_) k_sweep_up_packed<2, U_, E_, NG_><<<gridL, UPP_THREADS, p->smemBP, s>>>(m, p->tabP, p->choffP, p->wcnt4P, bsz, p->mt, a, 0, nrows, want_dot)
	if (p->packedB) {
		const bool uni = p->mt.nmag == 1;
		if (p->packedE16 && p->packedNG == 12) { if (uni) RUNP(true, true, 12); else RUNP(false, true, 12); }
		else if (p->packedE16) { if (uni) RUNP(true, true, 8); else RUNP(false, true, 8); }
		else { if (uni) RUNP(true, false, 8); else RUNP(false, false, 8); }
	} else if (p->leanB) {
		if (p->R == 2) { if (p->mt.nmag == 1) RUNL(2, true); else RUNL(2, false); }
		else { if (p->mt.nmag == 1) RUNL(1, true); else RUNL(1, false); }
	} else {
		// blocked up sweep on the row shard (same kernels as lpp_tiled_spmv, local rows 0 .. nrows)
#define RUNB(R_, N_) k_sweep_up_pipe<R_, N_><<<gridL, PBP_THREADS, p->smemB, s>>>(m, p->up, p->mt, a, 0, nrows, want_dot)
		if (p->pipeB) {
			if (p->R == 2) { if (p->NE == 16) RUNB(2, 16); else if (p->NE == 24) RUNB(2, 24); else RUNB(2, 32); }
			else { if (p->NE == 16) RUNB(1, 16); else if (p->NE == 24) RUNB(1, 24); else RUNB(1, 32); }
		} else if (p->R == 2) k_sweep_up_blocks<2><<<gridL, PB_THREADS, p->smemB, s>>>(m, p->up, p->mt, a, 0, nrows, want_dot);
		else k_sweep_up_blocks<1><<<gridL, PB_THREADS, p->smemB, s>>>(m, p->up, p->mt, a, 0, nrows, want_dot);
#undef RUNB
	}
#undef RUNP
#undef RUNL
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { g_terr = cudaGetErrorString(e); return -1; }
	return 1;
}

int lpp_tiled_down_cols_blocks(const TiledPlan* p, const ModelDev& m, uint64_t ncols)
{
	const ColView cvt{ncols, ncols, 0};
	{
		DiagTables dtu{};
		dtu.uniformU = 1;
		if (p->dblock && lpp_dblock_accepts(p->dblock, m, dtu, 0, m.n2, cvt)) return lpp_dblock_partials(p->dblock, cvt);
	}
	if (p->drows && !p->dtile && lpp_drows_accepts(p->drows, cvt)) return lpp_drows_grid(p->drows, cvt);
	if (p->dtile && lpp_dtile_accepts(p->dtile, cvt)) return lpp_dtile_grid(p->dtile, cvt);
	if (staged_accepts(p, cvt)) return (int)staged_grid(p, 0, m.n2, cvt);
	const uint32_t nchunks = (uint32_t)((m.n2 + PAL_ROWS - 1) / PAL_ROWS);
	ColView cv{ncols, ncols, 0};
	return (int)(down_lean_panels(cv) * nchunks);
}

// xcol = beta xcol + alpha (D + 1 (x) T_dn) ycol on the column shard [u0, u0+ncols), all Ndn rows, pitch = ncols
int lpp_tiled_sweep_down_cols(TiledPlan* p, const ModelDev& m, const HopTable& dn, const DiagTables& dt, const SpmvArgs& a,
                              uint64_t u0, uint64_t ncols, cudaStream_t s)
{
	const uint32_t nchunks = (uint32_t)((m.n2 + PAL_ROWS - 1) / PAL_ROWS);
	ColView cv{ncols, ncols, u0};
	if (p->dblock && lpp_dblock_accepts(p->dblock, m, dt, 0, m.n2, cv)) {
		if (lpp_dblock_sweep(p->dblock, m, dt, a, cv, s) < 0) { g_terr = lpp_dblock_error(); return -1; }
		return 2;
	}
	if (p->drows && !p->dtile && lpp_drows_accepts(p->drows, cv)) {
		if (lpp_drows_sweep(p->drows, m, dt, a, 0, m.n2, cv, s) < 0) { g_terr = lpp_dtile_error(); return -1; }
		return 1;
	}
	if (p->dtile && lpp_dtile_accepts(p->dtile, cv)) {
		if (lpp_dtile_sweep(p->dtile, m, dt, a, 0, m.n2, cv, s) < 0) { g_terr = lpp_dtile_error(); return -1; }
		return 1;
	}
	if (staged_accepts(p, cv)) {
		staged_launch(p, m, dt, a, 0, m.n2, cv, s);
		cudaError_t es = cudaGetLastError();
		if (es != cudaSuccess) { g_terr = cudaGetErrorString(es); return -1; }
		return 1;
	}
	if (down_lean_vec(cv) == 2) k_sweep_down_lean<2><<<lpp_tiled_down_cols_blocks(p, m, ncols), PAL_COLS, p->smemAL, s>>>(m, dn, dt, a, 0, m.n2, nchunks, cv);
	else k_sweep_down_lean<1><<<lpp_tiled_down_cols_blocks(p, m, ncols), PAL_COLS, p->smemAL, s>>>(m, dn, dt, a, 0, m.n2, nchunks, cv);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { g_terr = cudaGetErrorString(e); return -1; }
	return 1;
}
