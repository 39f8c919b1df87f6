// lpp_tiled.cu -- product-basis fast path (HubbardOneBand, FeAsBasedSc hopping part):
//   x = beta x + alpha (D + 1 (x) T_dn) y      sweep A: column panels, down-hops are whole coalesced row segments
//   x += alpha (T_up (x) 1) y                  sweep B: one up-segment (row of the Ndn x Nup matrix) staged in shared memory
//   x += alpha (two-spin on-site terms) y      sweep C: FeAs U2/U3 only
// The vector is viewed as the matrix Y[idn][iup] (index = iup + idn*Nup, BasisHubbardLanczos.h:59-63).
#include <string>
#include "lpp_tiled.cuh"

static thread_local std::string g_terr;
const char* lpp_tiled_error() { return g_terr.c_str(); }

struct TiledPlan {
	uint64_t d0, dcount;
	uint32_t nrowchunks, npanels;
	int up_in_smem;
	size_t up_smem_bytes;
	int has_twospin;
	int sm_count;
	int dot_blocks;
};

#define PA_COLS 256
#define PA_ROWS 8
#define PB_THREADS 1024

__device__ __forceinline__ double tiled_warp_sum(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ double tiled_block_sum(double v)
{
	__shared__ double red[32];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	v = tiled_warp_sum(v);
	if (lane == 0) red[wid] = v;
	__syncthreads();
	const int nw = (blockDim.x + 31) >> 5;
	v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
	if (wid == 0) v = tiled_warp_sum(v);
	__syncthreads();
	return v;
}

__device__ __forceinline__ double tiled_feas_diag(const ModelDev& m, word_t k1, word_t k2)
{
	const int no = m.orbitals;
	double s = m.U[0] * (double)lpp_popc(k1 & k2);
	word_t m0 = 0;
	for (int i = 0; i < m.nsite; i++) m0 |= lpp_bit(i * no);
	for (int a = 0; a < no; a++) {
		word_t A1 = (k1 >> a) & m0, A2 = (k2 >> a) & m0;
		for (int b = a + 1; b < no; b++) {
			word_t B1 = (k1 >> b) & m0, B2 = (k2 >> b) & m0;
			int uu = lpp_popc(A1 & B1), ud = lpp_popc(A1 & B2), du = lpp_popc(A2 & B1), dd = lpp_popc(A2 & B2);
			s += m.U[1] * (double)(uu + ud + du + dd);
			s += m.U[4] * 0.25 * (double)(uu - ud - du + dd);
			s += m.U[5] * (double)(uu + dd);
		}
	}
	if (m.D[0] != 0.0) {
		for (int i = 0; i < m.nsite; i++) {
			word_t sm = lpp_below(no) << (i * no);
			double sz = 0.5 * (double)(lpp_popc(k1 & sm) - lpp_popc(k2 & sm));
			s += m.D[0] * sz * sz;
		}
	}
	return s;
}

__device__ __forceinline__ double tiled_diag(const ModelDev& m, const DiagTables& dt, word_t k1, word_t k2, uint64_t i1,
                                             uint64_t i2)
{
	double s;
	if (m.model == LPP_MODEL_HUBBARD) {
		if (dt.uniformU) s = dt.U0 * (double)lpp_popc(k1 & k2);
		else {
			s = 0;
			word_t b = k1 & k2;
			while (b) { s += m.U[lpp_ctz(b)]; b &= b - 1; }
		}
	} else {
		s = tiled_feas_diag(m, k1, k2);
	}
	return s + dt.dv1[i1] + dt.dv2[i2];
}

// sweep A: thread = one column u of a 256-column panel, CTA walks PA_ROWS rows.  Every down-hop of row d is a
// contiguous 2 KB read of another row of the same panel (CTA-uniform table entry, broadcast through L1).
// blockIdx is panel-major so the panels in flight (a few tens of MB) stay L2 resident.
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

// sweep B: CTA = one up-segment Y[d][0..Nup) staged in shared memory; up-hops are shared-memory gathers.
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

// sweep B fallback when one up-segment does not fit in shared memory: gathers from global memory (L1/L2)
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

int lpp_tiled_create(const ModelDev& m, const double* hop_host, const HopTable& up, const HopTable& dn, uint64_t row0,
                     uint64_t nloc, cudaStream_t s, TiledPlan** out)
{
	(void)hop_host; (void)up; (void)dn; (void)s;
	if (m.model == LPP_MODEL_HEISENBERG) { g_terr = "tiled path is for product bases"; return -1; }
	TiledPlan* p = new TiledPlan();
	p->d0 = row0 / m.n1;
	p->dcount = nloc / m.n1;
	p->nrowchunks = (uint32_t)((p->dcount + PA_ROWS - 1) / PA_ROWS);
	p->npanels = (uint32_t)((m.n1 + PA_COLS - 1) / PA_COLS);
	p->up_smem_bytes = (size_t)m.n1 * sizeof(double);
	int dev = 0, maxsm = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
	cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, dev);
	p->up_in_smem = (p->up_smem_bytes + 1024 <= (size_t)maxsm) ? 1 : 0;
	if (p->up_in_smem) {
		cudaError_t e = cudaFuncSetAttribute(k_sweep_up_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->up_smem_bytes);
		if (e != cudaSuccess) { g_terr = cudaGetErrorString(e); delete p; return -1; }
	}
	p->has_twospin = (m.model == LPP_MODEL_FEAS) ? 1 : 0;
	// the last sweep owns the dot-product partial sums
	if (p->has_twospin) p->dot_blocks = (int)((nloc + 255) / 256);
	else if (p->up_in_smem) p->dot_blocks = (int)p->dcount;
	else p->dot_blocks = (int)(((m.n1 + 255) / 256) * p->dcount);
	*out = p;
	return 0;
}

void lpp_tiled_destroy(TiledPlan* p) { delete p; }

int lpp_tiled_dot_blocks(const TiledPlan* p) { return p->dot_blocks; }

int lpp_tiled_spmv(TiledPlan* p, const ModelDev& m, const HopTable& up, const HopTable& dn, const DiagTables& dt,
                   const SpmvArgs& a, cudaStream_t s)
{
	int launches = 0;
	uint64_t nblkA = (uint64_t)p->npanels * p->nrowchunks;
	k_sweep_down<<<(unsigned)nblkA, PA_COLS, 0, s>>>(m, dn, dt, a, p->d0, p->dcount, p->nrowchunks);
	launches++;
	const int dot_in_b = p->has_twospin ? 0 : 1;
	if (p->up_in_smem) {
		k_sweep_up_smem<<<(unsigned)p->dcount, PB_THREADS, p->up_smem_bytes, s>>>(m, up, a, p->d0, dot_in_b);
	} else {
		uint32_t nbx = (uint32_t)((m.n1 + 255) / 256);
		k_sweep_up_global<<<(unsigned)(nbx * p->dcount), 256, 0, s>>>(m, up, a, p->d0, nbx, dot_in_b);
	}
	launches++;
	if (p->has_twospin) {
		k_sweep_twospin<<<(unsigned)((a.nloc + 255) / 256), 256, 0, s>>>(m, a);
		launches++;
	}
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { g_terr = cudaGetErrorString(e); return -1; }
	return launches;
}
