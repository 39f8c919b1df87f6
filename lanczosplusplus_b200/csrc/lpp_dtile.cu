// lpp_dtile.cu -- spin-down sweep of the product-basis mat-vec on shared-memory tiles.
//
//   x[d][u] = beta x[d][u] + alpha ( diag(d,u) y[d][u] + sum_{d'} T_dn[d][d'] y[d'][u] )
//
// (the spin-down hopping part + diagonal of HubbardHelper.h:105-134 / FeBasedSc.h:228-245; index = iup + idn*Nup,
// BasisHubbardLanczos.h:59-63, so a down hop moves a whole row of the matrix Y[idn][iup]).
//
// Tile = (block of down states) x (16 contiguous columns), 128 bytes per tile row, staged once with cp.async.
// A block is the set of down states that share the occupation of the top F sites; F is the smallest value for which
// the largest block fits in shared memory (4x4, 8 particles: F = 3, 8 blocks of <= 1716 states, 214.5 KB).  Hops that do
// not touch the top F sites stay inside the block (22 of 32 bonds on the 4x4 torus): their operand is a conflict-free
// 16-byte shared-memory read (a quarter-warp reads one whole 128-byte tile row).  Hops that leave the block read the
// same 128-byte row segment from global memory; the grid is ordered column-group-major so those rows are L2 hits
// (a sibling CTA staged them moments ago).
//
// Work assignment: a quarter-warp (8 lanes x 2 columns) owns one tile row, a warp owns a "quad" of four rows with similar
// hop counts (rows are sorted by (external, internal) count inside the block, so padding is small and the trip counts
// are warp-uniform).  Table entries are per (row, hop), identical for every column group: external entries are 4 bytes
// (state index | magnitude | sign), internal entries 2 bytes (tile row | magnitude | sign); a quad's entries for the NEXT
// iteration are prefetched into registers while the current quad is reduced (the table lives in L2).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "lpp_dtile.cuh"
#include "lpp_smem_attr.cuh"

static thread_local std::string g_derr;
const char* lpp_dtile_error() { return g_derr.c_str(); }

#define DCK(call)                                                                  \
	do {                                                                           \
		cudaError_t e_ = (call);                                                   \
		if (e_ != cudaSuccess) { g_derr = std::string(#call) + ": " + cudaGetErrorString(e_); return -1; } \
	} while (0)

#ifndef DT_THREADS
#define DT_THREADS 384
#endif
#define DT_WARPS (DT_THREADS / 32)
#define DT_COLS 16
#define DT_ROWB 128            // bytes per tile row
#define DT_RPW 8               // rows per warp step: 4 lanes per row, every lane owns two 16-byte chunks (4 columns)
#define DT_NX 12               // <= 12 external hops per row (4-byte entries)
#define DT_NW 12               // internal entries of a row: <= 12 words of two entries, '+' list first, then the '-' list
#ifndef DT_XFIRST
#define DT_XFIRST 4            // external operands requested before the internal loop
#endif
#define DT_REC 1024            // bytes per record (8 rows) = two 16-byte cp.async per lane
#define DT_NOROW 0xffffffffu

// Record of a warp step (8 rows, 1024 bytes; fetched into the warp's slot while the previous one is reduced):
//   [  0] rowinfo[8]  { down-state index or DT_NOROW, tile row | external count << 16 }
//   [ 64] k2[8]       down word of the row (diagonal term)
//   [128] dv2[8]      one-spin potential sum of the row
//   [192] counts      { nx4 = steps of 4 external entries, np2 | nm2 << 8 | nbt << 16 } : words of the '+' / '-' internal lists
//                     (max over the 8 rows) and nbt = ceil((np2 + nm2) / 2) batches of two words
//   [256] ext[8][12]  4-byte entries: [31] sign | [27:24] magnitude | [23:0] down-state index
//   [640] int[8][12]  words of two internal entries: [10:0] tile row A, [14:11] magnitude A, [18:15] magnitude B, [31:21] tile row B;
//                     words 0..np2-1 carry hops with positive amplitude, words np2..np2+nm2-1 hops with negative amplitude;
//                     padding entries point at the zero row behind the block.  Splitting by sign makes the sign a warp-uniform
//                     multiplier, so an entry costs address + loads + FMAs.
// Bank conflicts: a quarter-warp is two rows x 4 lanes.  Lanes of the even row slot read chunk l4 first and chunk l4+4 second,
// lanes of the odd row slot read them in the opposite order, so every 16-byte load instruction touches bytes 0..63 of one
// source row and bytes 64..127 of the other: conflict free whatever rows the table points at.
struct DTileDev {
	const uint32_t* blk_off;    // nblocks+1: tile rows of block b are rowlist[blk_off[b] .. blk_off[b+1])
	const uint32_t* rowlist;    // down-state index of every tile row (basis order inside a block)
	const uint32_t* quad_off;   // nblocks+1: records of block b
	const uint4* rec;           // DT_REC/16 uint4 per record
	uint32_t ring_off;          // byte offset of the per-warp record slots behind the tile
	int nblocks;
};

struct DownTilePlan {
	DTileDev dev;
	int fsites = 0;
	uint32_t max_block = 0;
	size_t smem = 0;
	MagTable mt;
	size_t n_int = 0, n_ext = 0, n_int_padded = 0, n_ext_padded = 0, nquads = 0, nrows = 0;
	std::vector<void*> allocs;
};

__device__ __forceinline__ void dt_cp16(uint32_t dst, const void* src)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ double2 dt_lds16(uint32_t addr)
{
	double2 v;
	asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
	return v;
}
__device__ __forceinline__ uint4 dt_lds16u(uint32_t addr)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
	return v;
}
__device__ __forceinline__ uint2 dt_lds8u(uint32_t addr)
{
	uint2 v;
	asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
	return v;
}
// predicated 16-byte read-only global load (no branch): zero when the predicate is off
__device__ __forceinline__ double2 dt_ldg16_if(const void* ptr, bool pred)
{
	double2 v;
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
	             "@p ld.global.nc.v2.f64 {%0, %1}, [%2];\n\t}"
	             : "=d"(v.x), "=d"(v.y) : "l"(ptr), "r"((uint32_t)pred));
	return v;
}
// flip the sign of v when bit 31 of s is set
__device__ __forceinline__ double dt_sgn(double v, uint32_t s)
{
	return __hiloint2double(__double2hiint(v) ^ (int)(s & 0x80000000u), __double2loint(v));
}

template <bool UNI>
__global__ void __launch_bounds__(DT_THREADS, 1)
k_down_tile(ModelDev m, DTileDev P, MagTable mt, DiagTables dt, SpmvArgs a, uint64_t d0, uint64_t dcount, ColView cv)
{
	extern __shared__ __align__(128) unsigned char dt_smem[];
	const uint32_t b = blockIdx.x % (uint32_t)P.nblocks, g = blockIdx.x / (uint32_t)P.nblocks;
	const uint32_t boff = P.blk_off[b], bsize = P.blk_off[b + 1] - boff;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(dt_smem);
	const uint32_t pitch8 = (uint32_t)(cv.pitch * 8ull);        // < 4 GB per row (checked on the host)
	const uint32_t slot = sbase + P.ring_off + warp * DT_REC;   // this warp's record slot
	const uint32_t q0 = P.quad_off[b], nq = P.quad_off[b + 1] - q0;
	const uint64_t cg0 = (uint64_t)g * DT_COLS;                 // first column of the group

	// ---- stage the tile: row r of the block, 16 columns (8 lanes x 16 bytes); and the first record of every warp
	{
		const uint32_t l8 = threadIdx.x & 7;
		const uint64_t c = cg0 + 2 * l8;
		const bool ok = c < cv.ncols;
		const char* __restrict__ src = reinterpret_cast<const char*>(a.y + (ok ? c : 0));
		for (uint32_t r = threadIdx.x >> 3; r < bsize; r += DT_THREADS / 8) {
			const uint64_t d = P.rowlist[boff + r];
			if (ok) dt_cp16(sbase + r * DT_ROWB + l8 * 16, src + d * pitch8);
			else *reinterpret_cast<double2*>(dt_smem + (size_t)r * DT_ROWB + l8 * 16) = make_double2(0.0, 0.0);
		}
		if (threadIdx.x < 8) *reinterpret_cast<double2*>(dt_smem + (size_t)bsize * DT_ROWB + l8 * 16) = make_double2(0.0, 0.0);
	}
	if (warp < nq) {
		const uint4* __restrict__ rsrc = P.rec + (size_t)(q0 + warp) * (DT_REC / 16);
		dt_cp16(slot + lane * 16, rsrc + lane);
		dt_cp16(slot + 512 + lane * 16, rsrc + 32 + lane);
	}
	asm volatile("cp.async.commit_group;");

	// ---- compute mapping: 4 lanes per row; chunk order depends on the parity of the row slot
	const uint32_t rs = lane >> 2, l4 = lane & 3;
	const uint32_t ch[2] = {l4 + 4 * (rs & 1), l4 + 4 * ((rs & 1) ^ 1)};   // first / second 16-byte chunk of this lane
	uint32_t lb[2];
	bool colok[2];
	const char* __restrict__ ycol[2];
	char* __restrict__ xcol[2];
	word_t k1[2][2];
	double dv1[2][2];
#pragma unroll
	for (int h = 0; h < 2; h++) {
		lb[h] = sbase + ch[h] * 16;
		const uint64_t c = cg0 + 2 * ch[h];
		colok[h] = c < cv.ncols;
		ycol[h] = reinterpret_cast<const char*>(a.y + (colok[h] ? c : 0));
		xcol[h] = reinterpret_cast<char*>(a.x + (colok[h] ? c : 0));
#pragma unroll
		for (int v = 0; v < 2; v++) {
			k1[h][v] = colok[h] ? m.b1[cv.u0 + c + v] : 0;
			dv1[h][v] = colok[h] ? dt.dv1[cv.u0 + c + v] : 0.0;
		}
	}
	const bool need_x = a.beta != 0.0;
	const bool fastdiag = (m.model == LPP_MODEL_HUBBARD) && dt.uniformU;
	const double t0 = mt.mag[0];
	const uint32_t d0w = (uint32_t)d0, dcw = (uint32_t)dcount;   // down-state indices fit 24 bits
	double contrib = 0.0;

	asm volatile("cp.async.wait_group 0;");
	__syncthreads();
	for (uint32_t q = warp; q < nq; q += DT_WARPS) {
		if (q != warp) {
			asm volatile("cp.async.wait_group 0;");
			__syncwarp();
		}
		// ---- the record goes to registers, then its slot is refilled with the next record
		const uint2 ri = dt_lds8u(slot + rs * 8);
		const uint2 k2w = dt_lds8u(slot + 64 + rs * 8);
		const uint2 dvw = dt_lds8u(slot + 128 + rs * 8);
		const uint2 cn = dt_lds8u(slot + 192);
		uint32_t ex[DT_NX], iw[DT_NW];
#pragma unroll
		for (int k = 0; k < DT_NX / 4; k++) {
			const uint4 v = dt_lds16u(slot + 256 + rs * (DT_NX * 4) + k * 16);
			ex[4 * k + 0] = v.x; ex[4 * k + 1] = v.y; ex[4 * k + 2] = v.z; ex[4 * k + 3] = v.w;
		}
#pragma unroll
		for (int k = 0; k < DT_NW / 4; k++) {
			const uint4 v = dt_lds16u(slot + 256 + DT_RPW * DT_NX * 4 + rs * (DT_NW * 4) + k * 16);
			iw[4 * k + 0] = v.x; iw[4 * k + 1] = v.y; iw[4 * k + 2] = v.z; iw[4 * k + 3] = v.w;
		}
		__syncwarp();
		if (q + DT_WARPS < nq) {
			const uint4* __restrict__ rsrc = P.rec + (size_t)(q0 + q + DT_WARPS) * (DT_REC / 16);
			dt_cp16(slot + lane * 16, rsrc + lane);
			dt_cp16(slot + 512 + lane * 16, rsrc + 32 + lane);
		}
		asm volatile("cp.async.commit_group;");

		const uint32_t d = ri.x, lpos = ri.y & 0xffffu, nx = ri.y >> 16;
		const uint32_t nx4 = cn.x, np2 = cn.y & 0xffu, nbt = (cn.y >> 16) & 0xffu;   // nbt: batches of 2 words
		const uint32_t dl = d - d0w;
		const bool rowok = d != DT_NOROW && dl < dcw;
		const uint64_t xoff = (uint64_t)(rowok ? dl : 0u) * pitch8;
		double2 xo[2];
#pragma unroll
		for (int h = 0; h < 2; h++) xo[h] = dt_ldg16_if(xcol[h] + xoff, rowok && colok[h] && need_x);
		// external operands: 128-byte row segments served by L2; the first DT_XFIRST are requested before the internal loop
		double2 xv[DT_XFIRST][2];
#pragma unroll
		for (int k = 0; k < DT_XFIRST / 4; k++)
			if (k < (int)nx4) {                                 // warp-uniform
#pragma unroll
				for (int i = 0; i < 4; i++) {
					const int j = 4 * k + i;
					const uint64_t off = (uint64_t)(ex[j] & 0xffffffu) * pitch8;
#pragma unroll
					for (int h = 0; h < 2; h++) xv[j][h] = dt_ldg16_if(ycol[h] + off, rowok && colok[h] && j < (int)nx);
				}
			}
		// internal hops: batches of 2 words = 8 independent conflict-free 16-byte shared-memory reads, then 16 FMAs with a
		// warp-uniform amplitude (+|t| for the words of the '+' list, -|t| behind it); padding words read the zero row
		double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, acd[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
		for (int bt = 0; bt < DT_NW / 2; bt++)
			if (bt < (int)nbt) {                                // warp-uniform
				double2 va[2][2], vb[2][2];
#pragma unroll
				for (int i = 0; i < 2; i++) {
					const uint32_t w = iw[2 * bt + i];
					const uint32_t offa = (w & 0x7ffu) * DT_ROWB;
					const uint32_t offb = UNI ? (w >> 14) : ((w >> 14) & ~0x7fu);   // UNI: magnitude fields are zero
#pragma unroll
					for (int h = 0; h < 2; h++) {
						va[i][h] = dt_lds16(lb[h] + offa);
						vb[i][h] = dt_lds16(lb[h] + offb);
					}
				}
#pragma unroll
				for (int i = 0; i < 2; i++) {
					const uint32_t w = iw[2 * bt + i];
					const uint32_t sg = (2 * bt + i < (int)np2) ? 0u : 0x80000000u;
					const double ma = dt_sgn(UNI ? t0 : mt.mag[(w >> 11) & 15u], sg);
					const double mb = UNI ? ma : dt_sgn(mt.mag[(w >> 15) & 15u], sg);
#pragma unroll
					for (int h = 0; h < 2; h++) {
						acc[h][0] = fma(ma, va[i][h].x, acc[h][0]);
						acc[h][1] = fma(ma, va[i][h].y, acc[h][1]);
						acd[h][0] = fma(mb, vb[i][h].x, acd[h][0]);
						acd[h][1] = fma(mb, vb[i][h].y, acd[h][1]);
					}
				}
			}
		double2 ys[2];
#pragma unroll
		for (int h = 0; h < 2; h++) ys[h] = dt_lds16(lb[h] + lpos * DT_ROWB);
		// external contributions (sign by xor, amplitude per entry)
#pragma unroll
		for (int k = 0; k < DT_XFIRST / 4; k++)
			if (k < (int)nx4) {
#pragma unroll
				for (int i = 0; i < 4; i++) {
					const uint32_t e = ex[4 * k + i];
					const double mg = dt_sgn(UNI ? t0 : mt.mag[(e >> 24) & 15u], e);
#pragma unroll
					for (int h = 0; h < 2; h++) {
						double (&pp)[2] = (i & 1) ? acd[h] : acc[h];
						pp[0] = fma(mg, xv[4 * k + i][h].x, pp[0]);
						pp[1] = fma(mg, xv[4 * k + i][h].y, pp[1]);
					}
				}
			}
#pragma unroll
		for (int k = DT_XFIRST / 4; k < DT_NX / 4; k++)
			if (k < (int)nx4) {                                 // later steps of 4 external hops: request, then reduce
				double2 xw[4][2];
#pragma unroll
				for (int i = 0; i < 4; i++) {
					const int j = 4 * k + i;
					const uint64_t off = (uint64_t)(ex[j] & 0xffffffu) * pitch8;
#pragma unroll
					for (int h = 0; h < 2; h++) xw[i][h] = dt_ldg16_if(ycol[h] + off, rowok && colok[h] && j < (int)nx);
				}
#pragma unroll
				for (int i = 0; i < 4; i++) {
					const uint32_t e = ex[4 * k + i];
					const double mg = dt_sgn(UNI ? t0 : mt.mag[(e >> 24) & 15u], e);
#pragma unroll
					for (int h = 0; h < 2; h++) {
						acc[h][0] = fma(mg, xw[i][h].x, acc[h][0]);
						acc[h][1] = fma(mg, xw[i][h].y, acc[h][1]);
					}
				}
			}
		if (rowok) {
			const word_t k2 = (word_t)k2w.x | ((word_t)k2w.y << 32);
			const double dv2 = __hiloint2double((int)dvw.y, (int)dvw.x);
#pragma unroll
			for (int h = 0; h < 2; h++) {
				if (!colok[h]) continue;
				double xn[2];
				const double yv[2] = {ys[h].x, ys[h].y};
				const double xov[2] = {xo[h].x, xo[h].y};
#pragma unroll
				for (int v = 0; v < 2; v++) {
					const double dg = fastdiag ? dt.U0 * (double)lpp_popc(k1[h][v] & k2) + dv1[h][v] + dv2
					                           : tiled_diag(m, dt, k1[h][v], k2, cv.u0 + cg0 + 2 * ch[h] + v, d);
					xn[v] = a.alpha * fma(dg, yv[v], acc[h][v] + acd[h][v]);
					if (need_x) xn[v] = fma(a.beta, xov[v], xn[v]);
					contrib = fma(yv[v], xn[v], contrib);
				}
				*reinterpret_cast<double2*>(xcol[h] + xoff) = make_double2(xn[0], xn[1]);
			}
		}
	}
	if (a.dot_partials) {
		const double s = tiled_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = s;
	}
}

// =====================================================================================================
// plan construction (host)
// =====================================================================================================
template <class T>
static int dt_upload(DownTilePlan* p, const T** out, const std::vector<T>& v)
{
	void* q = nullptr;
	DCK(cudaMalloc(&q, std::max<size_t>(v.size(), 1) * sizeof(T)));
	p->allocs.push_back(q);
	if (!v.empty()) DCK(cudaMemcpy(q, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
	*out = (const T*)q;
	return 0;
}

int lpp_dtile_create(const ModelDev& m, const HopTable& dn, const double* dv2_dev, const MagTable& mt, cudaStream_t s,
                     DownTilePlan** out)
{
	*out = nullptr;
	// Opt-in (LPP_DTILE=1): on B200 this kernel measures 2.6-3.1 ms per sweep on the 4x4 lattice against 2.05 ms for the
	// streaming kernel (profiles/README.md, "down tile kernel"): it is latency bound on the L2 round trip of the external
	// operands of every row step at the 12-20 warps the 214 KB tile leaves room for.
	const char* env = getenv("LPP_DTILE");
	if (!(env && env[0] == '1')) { g_derr = "not enabled (set LPP_DTILE=1)"; return 1; }
	if (m.model == LPP_MODEL_HEISENBERG) { g_derr = "one-spin basis"; return 1; }
	const uint64_t n = dn.n;
	if (n >= (1ull << 24)) { g_derr = "down basis too large for 24-bit entries"; return 1; }
	if (mt.nmag > 16) { g_derr = "too many hop magnitudes"; return 1; }
	int dev = 0, maxsm = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
	// shared memory: tile rows + the zero row + one record slot per warp + 256 static bytes of the block reduction
	uint32_t cap = std::min<uint32_t>(2046u, (uint32_t)(((size_t)maxsm - 256 - DT_WARPS * DT_REC) / DT_ROWB - 1));
	const char* ecap = getenv("LPP_DTILE_CAP");               // tests: force several blocks on small bases
	if (ecap && atoi(ecap) > 0) cap = std::min<uint32_t>(cap, (uint32_t)atoi(ecap));

	DCK(cudaStreamSynchronize(s));
	const int width = dn.width;
	std::vector<word_t> words(n);
	std::vector<uint32_t> idx((size_t)std::max(width, 1) * n), cnt(n);
	std::vector<double> val((size_t)std::max(width, 1) * n);
	DCK(cudaMemcpy(words.data(), m.b2, sizeof(word_t) * n, cudaMemcpyDeviceToHost));
	DCK(cudaMemcpy(cnt.data(), dn.cnt, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
	if (width > 0) {
		DCK(cudaMemcpy(idx.data(), dn.idx, sizeof(uint32_t) * idx.size(), cudaMemcpyDeviceToHost));
		DCK(cudaMemcpy(val.data(), dn.val, sizeof(double) * val.size(), cudaMemcpyDeviceToHost));
	}

	// blocks: occupation pattern of the top f sites; smallest f whose largest block fits
	std::vector<uint32_t> blk_of(n), blk_off, rowlist(n), local(n);
	int f = 0;
	uint32_t mx = 0;
	for (;; f++) {
		if (f > m.nsite) { g_derr = "cannot block the down basis"; return 1; }
		const int shift = (m.nsite - f) * m.orbitals;
		std::vector<uint64_t> keys(n);
		for (uint64_t i = 0; i < n; i++) keys[i] = f ? (uint64_t)(words[i] >> shift) : 0;
		std::vector<uint64_t> uniq(keys);
		std::sort(uniq.begin(), uniq.end());
		uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
		std::vector<uint32_t> count(uniq.size(), 0);
		for (uint64_t i = 0; i < n; i++) {
			blk_of[i] = (uint32_t)(std::lower_bound(uniq.begin(), uniq.end(), keys[i]) - uniq.begin());
			count[blk_of[i]]++;
		}
		mx = *std::max_element(count.begin(), count.end());
		if (mx > cap) continue;
		blk_off.assign(uniq.size() + 1, 0);
		for (size_t bb = 0; bb < uniq.size(); bb++) blk_off[bb + 1] = blk_off[bb] + count[bb];
		std::vector<uint32_t> fill(uniq.size(), 0);
		for (uint64_t i = 0; i < n; i++) {
			const uint32_t bb = blk_of[i];
			local[i] = fill[bb];
			rowlist[blk_off[bb] + fill[bb]++] = (uint32_t)i;
		}
		break;
	}
	const int nblocks = (int)blk_off.size() - 1;

	DownTilePlan* p = new DownTilePlan();
	p->fsites = f;
	p->max_block = mx;
	p->dev.ring_off = (mx + 1) * DT_ROWB;
	p->smem = (size_t)p->dev.ring_off + (size_t)DT_WARPS * DT_REC;
	p->mt = mt;

	struct RowEnt { std::vector<uint32_t> ex; std::vector<uint16_t> inp, inm; };   // internal: tile row | magnitude << 11
	std::vector<RowEnt> re(n);
	for (uint64_t d = 0; d < n; d++) {
		for (uint32_t k = 0; k < cnt[d]; k++) {
			const uint32_t tgt = idx[(size_t)k * n + d];
			const double v = val[(size_t)k * n + d];
			if (v == 0.0) continue;
			int mi = -1;
			for (int q = 0; q < mt.nmag; q++)
				if (mt.mag[q] == fabs(v)) mi = q;
			if (mi < 0) { g_derr = "hop amplitude not in the magnitude table"; delete p; return 1; }
			if (blk_of[tgt] == blk_of[d]) (v < 0 ? re[d].inm : re[d].inp).push_back((uint16_t)(local[tgt] | ((uint32_t)mi << 11)));
			else re[d].ex.push_back(tgt | ((uint32_t)mi << 24) | (v < 0 ? 0x80000000u : 0u));
		}
		if (re[d].ex.size() > DT_NX || (re[d].inp.size() + 1) / 2 + (re[d].inm.size() + 1) / 2 > DT_NW) {
			g_derr = "too many hops per state for the tile kernel";
			delete p;
			return 1;
		}
		p->n_int += re[d].inp.size() + re[d].inm.size();
		p->n_ext += re[d].ex.size();
	}

	std::vector<uint32_t> quad_off(nblocks + 1, 0);
	std::vector<uint32_t> rec;                                  // DT_REC/4 words per quad
	std::vector<double> dv2(n, 0.0);
	if (dv2_dev) DCK(cudaMemcpy(dv2.data(), dv2_dev, sizeof(double) * n, cudaMemcpyDeviceToHost));
	for (int bb = 0; bb < nblocks; bb++) {
		const uint32_t bsize = blk_off[bb + 1] - blk_off[bb];
		std::vector<uint32_t> rows(rowlist.begin() + blk_off[bb], rowlist.begin() + blk_off[bb + 1]);
		// similar rows share a quad: trip counts are the maxima over the quad, padding reads the zero row
		std::stable_sort(rows.begin(), rows.end(), [&](uint32_t x, uint32_t y) {
			if (re[x].ex.size() != re[y].ex.size()) return re[x].ex.size() > re[y].ex.size();
			const size_t tx = re[x].inp.size() + re[x].inm.size(), ty = re[y].inp.size() + re[y].inm.size();
			if (tx != ty) return tx > ty;
			return re[x].inp.size() > re[y].inp.size();
		});
		const uint32_t zero_a = bsize, zero_b = bsize << 21;    // the zero row behind the block, magnitude index 0
		size_t q = 0;
		while (q < rows.size()) {
			// greedy: up to eight rows whose padded '+' and '-' lists fit the 12 words together
			int nr = 0;
			uint32_t np2 = 0, nm2 = 0;
			size_t mxx = 0;
			while (nr < DT_RPW && q + nr < rows.size()) {
				const RowEnt& r = re[rows[q + nr]];
				const uint32_t a2 = std::max<uint32_t>(np2, (uint32_t)((r.inp.size() + 1) / 2));
				const uint32_t b2 = std::max<uint32_t>(nm2, (uint32_t)((r.inm.size() + 1) / 2));
				if (a2 + b2 > DT_NW) break;
				np2 = a2;
				nm2 = b2;
				mxx = std::max(mxx, r.ex.size());
				nr++;
			}
			const uint32_t nx4 = (uint32_t)((mxx + 3) / 4);
			const size_t base = rec.size();
			rec.resize(base + DT_REC / 4, 0u);
			uint32_t* R = rec.data() + base;
			uint32_t* W = R + 64 + DT_RPW * DT_NX;              // [8][DT_NW] words
			for (int r = 0; r < DT_RPW; r++) {
				for (int j = 0; j < DT_NW; j++) W[r * DT_NW + j] = zero_a | zero_b;
				if (r < nr) {
					const uint32_t d = rows[q + r];
					R[2 * r] = d;
					R[2 * r + 1] = local[d] | ((uint32_t)re[d].ex.size() << 16);
					const uint64_t w = (uint64_t)words[d];
					R[16 + 2 * r] = (uint32_t)w;
					R[16 + 2 * r + 1] = (uint32_t)(w >> 32);
					uint64_t bits;
					memcpy(&bits, &dv2[d], 8);
					R[32 + 2 * r] = (uint32_t)bits;
					R[32 + 2 * r + 1] = (uint32_t)(bits >> 32);
					for (size_t j = 0; j < re[d].ex.size(); j++) R[64 + r * DT_NX + j] = re[d].ex[j];
					auto pack = [&](const std::vector<uint16_t>& v, size_t j) {
						const uint32_t ea = v[2 * j];
						const uint32_t eb = (2 * j + 1 < v.size()) ? v[2 * j + 1] : (uint32_t)bsize;
						return (ea & 0x7ffu) | ((ea >> 11) << 11) | (((eb >> 11) & 15u) << 15) | ((eb & 0x7ffu) << 21);
					};
					for (size_t j = 0; j < (re[d].inp.size() + 1) / 2; j++) W[r * DT_NW + j] = pack(re[d].inp, j);
					for (size_t j = 0; j < (re[d].inm.size() + 1) / 2; j++) W[r * DT_NW + np2 + j] = pack(re[d].inm, j);
				} else {
					R[2 * r] = DT_NOROW;
					R[2 * r + 1] = bsize;                       // tile row = the zero row, no external hops
				}
			}
			R[48] = nx4;
			R[49] = np2 | (nm2 << 8) | (((np2 + nm2 + 1) / 2) << 16);
			p->n_int_padded += (size_t)((np2 + nm2 + 1) / 2) * 4 * nr;
			p->n_ext_padded += (size_t)nx4 * 4 * nr;
			p->nrows += nr;
			q += nr;
		}
		quad_off[bb + 1] = (uint32_t)(rec.size() / (DT_REC / 4));
	}
	p->nquads = rec.size() / (DT_REC / 4);
	p->dev.nblocks = nblocks;
	const uint32_t* recdev = nullptr;
	if (dt_upload(p, &p->dev.blk_off, blk_off) || dt_upload(p, &p->dev.rowlist, rowlist) || dt_upload(p, &p->dev.quad_off, quad_off) ||
	    dt_upload(p, &recdev, rec)) {
		lpp_dtile_destroy(p);
		return -1;
	}
	p->dev.rec = reinterpret_cast<const uint4*>(recdev);
	cudaError_t e1 = lpp_raise_smem(k_down_tile<true>, (size_t)(p->smem));
	cudaError_t e2 = lpp_raise_smem(k_down_tile<false>, (size_t)(p->smem));
	if (e1 != cudaSuccess || e2 != cudaSuccess) {
		g_derr = std::string("cudaFuncSetAttribute(k_down_tile): ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2);
		lpp_dtile_destroy(p);
		return -1;
	}
	if (getenv("LPP_VERBOSE")) {
		char buf[256];
		lpp_dtile_describe(p, buf, sizeof buf);
		fprintf(stderr, "[lpp dtile] %s\n", buf);
	}
	*out = p;
	return 0;
}

void lpp_dtile_describe(const DownTilePlan* p, char* buf, size_t n)
{
	snprintf(buf, n, "top sites=%d blocks=%d max rows=%u smem=%zu quads=%zu hops: internal %zu (padded %zu) external %zu (slots %zu)",
	         p->fsites, p->dev.nblocks, p->max_block, p->smem, p->nquads, p->n_int, p->n_int_padded, p->n_ext, p->n_ext_padded);
}

void lpp_dtile_destroy(DownTilePlan* p)
{
	if (!p) return;
	for (void* q : p->allocs) cudaFree(q);
	delete p;
}

int lpp_dtile_accepts(const DownTilePlan* p, const ColView& cv)
{
	return (p && cv.pitch % 2 == 0 && cv.ncols % 2 == 0 && cv.ncols > 0 && cv.pitch * 8ull < (1ull << 32)) ? 1 : 0;
}

int lpp_dtile_grid(const DownTilePlan* p, const ColView& cv)
{
	return (int)(((cv.ncols + DT_COLS - 1) / DT_COLS) * (uint64_t)p->dev.nblocks);
}

int lpp_dtile_sweep(DownTilePlan* p, const ModelDev& m, const DiagTables& dt, const SpmvArgs& a, uint64_t d0, uint64_t dcount,
                    const ColView& cv, cudaStream_t s)
{
	if (!lpp_dtile_accepts(p, cv)) { g_derr = "column view needs an even pitch and an even column count"; return -1; }
	const unsigned grid = (unsigned)lpp_dtile_grid(p, cv);
	if (p->mt.nmag == 1) k_down_tile<true><<<grid, DT_THREADS, p->smem, s>>>(m, p->dev, p->mt, dt, a, d0, dcount, cv);
	else k_down_tile<false><<<grid, DT_THREADS, p->smem, s>>>(m, p->dev, p->mt, dt, a, d0, dcount, cv);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { g_derr = cudaGetErrorString(e); return -1; }
	return 1;
}

// =====================================================================================================
// k_sweep_down_rows: the down sweep without a staged tile (see lpp_dtile.cuh)
// =====================================================================================================
#ifndef DR_THREADS
#define DR_THREADS 768
#endif
#define DR_WARPS (DR_THREADS / 32)
#define DR_MAXHOPS 32
#define DR_REC 272             // bytes per record of 4 consecutive rows: 16-byte header + 4 x 32 two-byte entries
#ifndef DR_SPLIT
#define DR_SPLIT 8             // row ranges per column group (grid = column groups x DR_SPLIT)
#endif

struct DownRowsPlan {
	const uint4* rec = nullptr;   // record q covers down states 4q .. 4q+3: header bytes {cnt0, cnt1, cnt2, cnt3, max, ...},
	                              // then entries[4][32]: [15] sign | [13:0] source down state
	uint32_t nquads = 0;
	uint64_t n2 = 0;
	MagTable mt;
	std::vector<void*> allocs;
};

__device__ __forceinline__ double2 dr_ldg16_if(const void* ptr, bool pred)
{
	double2 v;
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
	             "@p ld.global.v2.f64 {%0, %1}, [%2];\n\t}"
	             : "=d"(v.x), "=d"(v.y) : "l"(ptr), "r"((uint32_t)pred));
	return v;
}

__global__ void __launch_bounds__(DR_THREADS, 1)
k_sweep_down_rows(ModelDev m, const uint4* __restrict__ rec, uint32_t nquads, double t0, DiagTables dt, SpmvArgs a, uint64_t d0,
                  uint64_t dcount, ColView cv)
{
	extern __shared__ __align__(16) unsigned char dr_smem[];   // [warp][2 stages][DR_REC]
	const uint32_t g = blockIdx.x / DR_SPLIT, sp = blockIdx.x % DR_SPLIT;
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l8 = lane & 7, grp = lane >> 3;
	const uint64_t c = (uint64_t)g * 16 + 2 * l8;
	const bool colok = c < cv.ncols;
	const uint32_t pitch8 = (uint32_t)(cv.pitch * 8ull);
	const char* __restrict__ ycol = reinterpret_cast<const char*>(a.y + (colok ? c : 0));
	char* __restrict__ xcol = reinterpret_cast<char*>(a.x + (colok ? c : 0));
	// this CTA's quads: an equal share of the quad list, walked in rounds of DR_WARPS consecutive quads
	const uint32_t per = (nquads + DR_SPLIT - 1) / DR_SPLIT;
	const uint32_t qa = sp * per, qb = min(nquads, qa + per);
	const uint32_t ring = (uint32_t)__cvta_generic_to_shared(dr_smem) + warp * (2 * DR_REC);
	word_t k1[2] = {0, 0};
	double dv1[2] = {0.0, 0.0};
	if (colok) {
#pragma unroll
		for (int v = 0; v < 2; v++) {
			k1[v] = m.b1[cv.u0 + c + v];
			dv1[v] = dt.dv1[cv.u0 + c + v];
		}
	}
	const bool need_x = a.beta != 0.0;
	const bool fastdiag = (m.model == LPP_MODEL_HUBBARD) && dt.uniformU;
	const uint32_t d0w = (uint32_t)d0, dcw = (uint32_t)dcount, n2w = (uint32_t)m.n2;
	double contrib = 0.0;

	auto fetch = [&](uint32_t q, uint32_t stage) {
		if (q < qb && lane < DR_REC / 16) dt_cp16(ring + stage * DR_REC + lane * 16, rec + (size_t)q * (DR_REC / 16) + lane);
		asm volatile("cp.async.commit_group;");
	};
	uint32_t q = qa + warp;
	fetch(q, 0);
	fetch(q + DR_WARPS, 1);
	uint32_t stage = 0;
	for (; q < qb; q += DR_WARPS, stage ^= 1) {
		asm volatile("cp.async.wait_group 1;");
		__syncwarp();
		const uint32_t rb = ring + stage * DR_REC;
		const uint2 hd = dt_lds8u(rb);
		const int padnet = (int)(int8_t)((hd.x >> (8 * grp)) & 0xffu);      // ('+' pads) - ('-' pads) of my row
		const uint32_t pg = hd.y & 0xffu, ng = (hd.y >> 8) & 0xffu;          // groups of 4 slots: '+' groups, all groups
		const uint32_t d = 4 * q + grp;
		const uint32_t dl = d - d0w;
		const bool act = colok && d < n2w && dl < dcw;
		const uint64_t xoff = (uint64_t)(act ? dl : 0u) * pitch8;
		const double2 ys = __ldg(reinterpret_cast<const double2*>(ycol + (uint64_t)(d < n2w ? d : 0u) * pitch8));
		double2 xo = make_double2(0.0, 0.0);
		if (act && need_x) xo = __ldcs(reinterpret_cast<const double2*>(xcol + xoff));
		word_t k2 = 0;
		double dv2 = 0.0;
		if (act) { k2 = m.b2[d]; dv2 = dt.dv2[d]; }
		// every slot is an unconditional 16-byte load (padding slots point at the row itself and are subtracted below);
		// the sign is warp-uniform per group of 4 slots: '+' groups first, then '-' groups
		double acc0[2] = {0.0, 0.0}, acc1[2] = {0.0, 0.0};
#pragma unroll
		for (int bt = 0; bt < DR_MAXHOPS / 8; bt++)
			if (8 * bt < (int)(4 * ng)) {                       // warp-uniform
				const uint4 e4 = dt_lds16u(rb + 16 + grp * (DR_MAXHOPS * 2) + bt * 16);
				const uint32_t w[4] = {e4.x, e4.y, e4.z, e4.w};
				double2 v[8];
#pragma unroll
				for (int i = 0; i < 8; i++) {
					const uint32_t e = (i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xffffu);
					v[i] = __ldg(reinterpret_cast<const double2*>(ycol + (uint64_t)e * pitch8));
				}
				const double sa = (2 * bt < (int)pg) ? t0 : -t0;      // slots 8bt .. 8bt+3
				const double sb = (2 * bt + 1 < (int)pg) ? t0 : ((2 * bt + 1 < (int)ng) ? -t0 : 0.0);
#pragma unroll
				for (int i = 0; i < 4; i++) {
					double (&pp)[2] = (i & 1) ? acc1 : acc0;
					pp[0] = fma(sa, v[i].x, pp[0]);
					pp[1] = fma(sa, v[i].y, pp[1]);
				}
#pragma unroll
				for (int i = 4; i < 8; i++) {
					double (&pp)[2] = (i & 1) ? acc1 : acc0;
					pp[0] = fma(sb, v[i].x, pp[0]);
					pp[1] = fma(sb, v[i].y, pp[1]);
				}
			}
		const double corr = -t0 * (double)padnet;
		__syncwarp();
		fetch(q + 2 * DR_WARPS, stage);
		if (act) {
			double xn[2];
			const double yv[2] = {ys.x, ys.y};
			const double xov[2] = {xo.x, xo.y};
#pragma unroll
			for (int v = 0; v < 2; v++) {
				const double dg = fastdiag ? dt.U0 * (double)lpp_popc(k1[v] & k2) + dv1[v] + dv2
				                           : tiled_diag(m, dt, k1[v], k2, cv.u0 + c + v, d);
				xn[v] = a.alpha * fma(dg + corr, yv[v], acc0[v] + acc1[v]);
				if (need_x) xn[v] = fma(a.beta, xov[v], xn[v]);
				contrib = fma(yv[v], xn[v], contrib);
			}
			__stcs(reinterpret_cast<double2*>(xcol + xoff), make_double2(xn[0], xn[1]));
		}
	}
	asm volatile("cp.async.wait_group 0;");
	if (a.dot_partials) {
		const double s = tiled_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = s;
	}
}

int lpp_drows_create(const ModelDev& m, const HopTable& dn, const MagTable& mt, cudaStream_t s, DownRowsPlan** out)
{
	*out = nullptr;
	// Opt-in (LPP_DROWS=1): measured 2.3 ms per sweep on the 4x4 lattice against 2.1 ms for the streaming kernel.  It does
	// what it was built for (L1 hit rate 55 %, L2->SM traffic 18 GB instead of 26.9 GB) but gathers served by L1 hits and L2
	// together top out near 45 B/clk/SM, so the sweep stays latency bound (profiles/README.md, "row-walking kernel").
	const char* env = getenv("LPP_DROWS");
	if (!(env && env[0] == '1')) { g_derr = "not enabled (set LPP_DROWS=1)"; return 1; }
	if (m.model == LPP_MODEL_HEISENBERG) { g_derr = "one-spin basis"; return 1; }
	if (mt.nmag != 1) { g_derr = "more than one hop magnitude"; return 1; }
	const uint64_t n = dn.n;
	if (n > (1u << 14)) { g_derr = "down basis too large for 14-bit entries"; return 1; }
	if (dn.width > DR_MAXHOPS) { g_derr = "too many hops per state"; return 1; }
	DCK(cudaStreamSynchronize(s));
	const int width = dn.width;
	std::vector<uint32_t> idx((size_t)std::max(width, 1) * n), cnt(n);
	std::vector<double> val((size_t)std::max(width, 1) * n);
	DCK(cudaMemcpy(cnt.data(), dn.cnt, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
	if (width > 0) {
		DCK(cudaMemcpy(idx.data(), dn.idx, sizeof(uint32_t) * idx.size(), cudaMemcpyDeviceToHost));
		DCK(cudaMemcpy(val.data(), dn.val, sizeof(double) * val.size(), cudaMemcpyDeviceToHost));
	}
	DownRowsPlan* p = new DownRowsPlan();
	p->mt = mt;
	p->n2 = n;
	p->nquads = (uint32_t)((n + 3) / 4);
	std::vector<uint8_t> rec((size_t)p->nquads * DR_REC, 0);
	unsigned long long nslots_total = 0;
	for (uint32_t q = 0; q < p->nquads; q++) {
		uint8_t* R = rec.data() + (size_t)q * DR_REC;
		uint16_t* E = reinterpret_cast<uint16_t*>(R + 16);
		std::vector<uint16_t> plus[4], minus[4];
		uint32_t mp = 0, mm = 0;
		for (int r = 0; r < 4; r++) {
			const uint64_t d = 4ull * q + r;
			if (d >= n) continue;
			for (uint32_t k = 0; k < cnt[d]; k++) {
				const double v = val[(size_t)k * n + d];
				if (v == 0.0) continue;
				if (fabs(v) != mt.mag[0]) { g_derr = "hop amplitude not in the magnitude table"; delete p; return 1; }
				(v < 0 ? minus[r] : plus[r]).push_back((uint16_t)idx[(size_t)k * n + d]);
			}
			mp = std::max<uint32_t>(mp, (uint32_t)plus[r].size());
			mm = std::max<uint32_t>(mm, (uint32_t)minus[r].size());
		}
		const uint32_t pg = (mp + 3) / 4, mg = (mm + 3) / 4;     // groups of 4 slots
		if (4 * (pg + mg) > DR_MAXHOPS) { g_derr = "too many hop slots per quad"; delete p; return 1; }
		for (int r = 0; r < 4; r++) {
			const uint64_t d = std::min<uint64_t>(4ull * q + r, n - 1);   // padding slots: the row itself (always a valid address)
			for (uint32_t j = 0; j < 4 * (pg + mg); j++) E[r * DR_MAXHOPS + j] = (uint16_t)d;
			for (size_t j = 0; j < plus[r].size(); j++) E[r * DR_MAXHOPS + j] = plus[r][j];
			for (size_t j = 0; j < minus[r].size(); j++) E[r * DR_MAXHOPS + 4 * pg + j] = minus[r][j];
			const int padp = (int)(4 * pg - plus[r].size()), padm = (int)(4 * mg - minus[r].size());
			R[r] = (uint8_t)(int8_t)(padp - padm);
		}
		R[4] = (uint8_t)pg;
		R[5] = (uint8_t)(pg + mg);
		nslots_total += 4ull * (pg + mg);
	}
	if (getenv("LPP_VERBOSE")) fprintf(stderr, "[lpp drows] quads=%u mean slots per row %.2f\n", p->nquads, (double)nslots_total / std::max<uint32_t>(p->nquads, 1));
	void* dptr = nullptr;
	DCK(cudaMalloc(&dptr, rec.size()));
	p->allocs.push_back(dptr);
	DCK(cudaMemcpy(dptr, rec.data(), rec.size(), cudaMemcpyHostToDevice));
	p->rec = reinterpret_cast<const uint4*>(dptr);
	*out = p;
	return 0;
}

void lpp_drows_destroy(DownRowsPlan* p)
{
	if (!p) return;
	for (void* q : p->allocs) cudaFree(q);
	delete p;
}

int lpp_drows_accepts(const DownRowsPlan* p, const ColView& cv)
{
	return (p && cv.pitch % 2 == 0 && cv.ncols % 2 == 0 && cv.ncols > 0 && cv.pitch * 8ull < (1ull << 32)) ? 1 : 0;
}

int lpp_drows_grid(const DownRowsPlan* p, const ColView& cv) { return (int)(((cv.ncols + 15) / 16) * DR_SPLIT); }

int lpp_drows_sweep(DownRowsPlan* p, const ModelDev& m, const DiagTables& dt, const SpmvArgs& a, uint64_t d0, uint64_t dcount,
                    const ColView& cv, cudaStream_t s)
{
	if (!lpp_drows_accepts(p, cv)) { g_derr = "column view needs an even pitch and an even column count"; return -1; }
	const size_t smem = (size_t)DR_WARPS * 2 * DR_REC;
	k_sweep_down_rows<<<(unsigned)lpp_drows_grid(p, cv), DR_THREADS, smem, s>>>(m, p->rec, p->nquads, p->mt.mag[0], dt, a, d0, dcount, cv);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { g_derr = cudaGetErrorString(e); return -1; }
	return 1;
}
