// lpp_kernels.cu -- sm_100a kernels: device basis construction (K1), per-spin hop tables, diagonal tables (K2),
// generic / table-driven on-the-fly x = beta x + alpha H y (K3, K4), stored CRS build + SpMV (K5), the fused Lanczos
// vector sweeps (K6, K7, K8) and operator application (K9).  The shared-memory tiled fast path is in lpp_tiled.cu.
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include "lpp_kernels.cuh"

#define LPP_TPB 256

// ------------------------------------------------------------------ reductions (fixed order => deterministic)
__device__ __forceinline__ double lpp_warp_sum(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	return v;
}

__device__ __forceinline__ double lpp_block_sum(double v)
{
	__shared__ double red[32];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	v = lpp_warp_sum(v);
	if (lane == 0) red[wid] = v;
	__syncthreads();
	const int nw = (blockDim.x + 31) >> 5;
	v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
	if (wid == 0) v = lpp_warp_sum(v);
	__syncthreads();
	return v;  // valid in thread 0
}

// ------------------------------------------------------------------ K1: bases, ranks, tables
__global__ void k_build_colex(const uint64_t* __restrict__ binom, int nsite, int npart, uint64_t n, word_t* __restrict__ out)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = lpp_unrank_colex(binom, nsite, npart, i);
}

__global__ void k_build_feas(ModelDev m, int spin, uint64_t n, word_t* __restrict__ out)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = lpp_unrank_feas(m, spin, i);
}

__global__ void k_rank(ModelDev m, int spin, const word_t* __restrict__ w, uint64_t n, uint64_t* __restrict__ out)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = lpp_rank_onespin(m, spin, w[i]);
}

// split colex rank tables: rank(w) = rlo[lo] + rhi[popc(lo)][hi]
__global__ void k_split_tables(const uint64_t* __restrict__ binom, int nbits, int lobits, uint32_t* __restrict__ rlo,
                               uint32_t* __restrict__ rhi)
{
	const int hibits = nbits - lobits;
	const uint64_t nlo = 1ull << lobits, nhi = 1ull << hibits;
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < nlo) rlo[i] = lpp_split_lo_entry(binom, i);
	if (i < (uint64_t)(lobits + 1) * nhi) rhi[i] = lpp_split_hi_entry(binom, lobits, hibits, i);
}

__global__ void k_lut(const word_t* __restrict__ b, uint64_t n, uint32_t* __restrict__ lut)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) lut[b[i]] = (uint32_t)i;
}

struct CountEmit {
	int n;
	__device__ void operator()(uint64_t, double) { n++; }
};

template <class Emit>
__device__ __forceinline__ void lpp_spin_hops(const ModelDev& m, int spin, word_t ket, Emit& e)
{
	if (m.model == LPP_MODEL_HUBBARD) lpp_hubbard_spin_hops(m, spin, ket, e);
	else lpp_feas_spin_hops(m, spin, ket, e);
}

__global__ void k_hop_count(ModelDev m, int spin, uint64_t n, uint32_t* __restrict__ cnt, uint32_t* __restrict__ maxcnt)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	CountEmit e{0};
	lpp_spin_hops(m, spin, (spin ? m.b2 : m.b1)[i], e);
	cnt[i] = (uint32_t)e.n;
	atomicMax(maxcnt, (uint32_t)e.n);
}

struct EllEmit {
	uint32_t* idx;
	double* val;
	uint64_t n, s;
	int k;
	__device__ void operator()(uint64_t t, double v)
	{
		idx[(uint64_t)k * n + s] = (uint32_t)t;
		val[(uint64_t)k * n + s] = v;
		k++;
	}
};

__global__ void k_hop_fill(ModelDev m, int spin, HopTable t)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= t.n) return;
	EllEmit e{t.idx, t.val, t.n, i, 0};
	lpp_spin_hops(m, spin, (spin ? m.b2 : m.b1)[i], e);
	for (int k = e.k; k < t.width; k++) {
		t.idx[(uint64_t)k * t.n + i] = (uint32_t)i;
		t.val[(uint64_t)k * t.n + i] = 0.0;
	}
}

// K2: per-spin potential sums (Hubbard: HubbardHelper.h:180-183 ; FeAs: FeBasedSc.h:557-561)
__global__ void k_spin_diag(ModelDev m, int spin, uint64_t n, double* __restrict__ dv)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	word_t w = (spin ? m.b2 : m.b1)[i];
	double s = 0;
	if (m.model == LPP_MODEL_HUBBARD) {
		for (int k = 0; k < m.nsite; k++) s += m.V[k] * (double)((w >> k) & 1);
	} else {
		for (int k = 0; k < m.nsite; k++)
			for (int o = 0; o < m.orbitals; o++)
				s += m.V[k + (o + m.orbitals * spin) * m.nsite] * (double)lpp_feas_occ(w, k, o, m.orbitals);
	}
	dv[i] = s;
}

void lpp_launch_build_colex(const uint64_t* binom, int nsite, int npart, uint64_t n, word_t* out, cudaStream_t s)
{
	k_build_colex<<<(unsigned)((n + LPP_TPB - 1) / LPP_TPB), LPP_TPB, 0, s>>>(binom, nsite, npart, n, out);
}
void lpp_launch_build_feas(const ModelDev& m, int spin, uint64_t n, word_t* out, cudaStream_t s)
{
	k_build_feas<<<(unsigned)((n + LPP_TPB - 1) / LPP_TPB), LPP_TPB, 0, s>>>(m, spin, n, out);
}
__global__ void __launch_bounds__(LPP_TPB) k_row_words(ModelDev m, uint64_t first, uint64_t count, word_t* up, word_t* dn)
{
	uint64_t t = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x;
	if (t >= count) return;
	const LppRowKets k = lpp_row_kets(m, first + t);
	if (up) up[t] = k.k1;
	if (dn) dn[t] = k.k2;
}
__global__ void __launch_bounds__(LPP_TPB) k_rank_pairs(ModelDev m, const word_t* __restrict__ up, const word_t* __restrict__ dn, uint64_t n,
                                                       uint64_t* __restrict__ out)
{
	uint64_t t = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x;
	if (t < n) out[t] = lpp_rank_pair(m, up[t], dn[t]);
}
void lpp_launch_row_words(const ModelDev& m, uint64_t first, uint64_t count, word_t* up, word_t* dn, cudaStream_t s)
{
	if (count) k_row_words<<<(unsigned)((count + LPP_TPB - 1) / LPP_TPB), LPP_TPB, 0, s>>>(m, first, count, up, dn);
}
void lpp_launch_rank_pairs(const ModelDev& m, const word_t* up, const word_t* dn, uint64_t n, uint64_t* out, cudaStream_t s)
{
	if (n) k_rank_pairs<<<(unsigned)((n + LPP_TPB - 1) / LPP_TPB), LPP_TPB, 0, s>>>(m, up, dn, n, out);
}
void lpp_launch_rank(const ModelDev& m, int spin, const word_t* w, uint64_t n, uint64_t* out, cudaStream_t s)
{
	k_rank<<<(unsigned)((n + LPP_TPB - 1) / LPP_TPB), LPP_TPB, 0, s>>>(m, spin, w, n, out);
}
void lpp_launch_split_tables(const uint64_t* binom, int nbits, int lobits, uint32_t* rlo, uint32_t* rhi, cudaStream_t s)
{
	uint64_t total = (uint64_t)(lobits + 1) << (nbits - lobits);
	uint64_t nlo = 1ull << lobits;
	uint64_t n = total > nlo ? total : nlo;
	k_split_tables<<<(unsigned)((n + LPP_TPB - 1) / LPP_TPB), LPP_TPB, 0, s>>>(binom, nbits, lobits, rlo, rhi);
}
void lpp_launch_lut(const word_t* b, uint64_t n, uint32_t* lut, cudaStream_t s)
{
	k_lut<<<(unsigned)((n + LPP_TPB - 1) / LPP_TPB), LPP_TPB, 0, s>>>(b, n, lut);
}
void lpp_launch_hop_count(const ModelDev& m, int spin, uint64_t n, uint32_t* cnt, uint32_t* maxcnt, cudaStream_t s)
{
	k_hop_count<<<(unsigned)((n + LPP_TPB - 1) / LPP_TPB), LPP_TPB, 0, s>>>(m, spin, n, cnt, maxcnt);
}
void lpp_launch_hop_fill(const ModelDev& m, int spin, HopTable t, cudaStream_t s)
{
	k_hop_fill<<<(unsigned)((t.n + LPP_TPB - 1) / LPP_TPB), LPP_TPB, 0, s>>>(m, spin, t);
}
void lpp_launch_spin_diag(const ModelDev& m, int spin, uint64_t n, double* dv, cudaStream_t s)
{
	k_spin_diag<<<(unsigned)((n + LPP_TPB - 1) / LPP_TPB), LPP_TPB, 0, s>>>(m, spin, n, dv);
}

// ------------------------------------------------------------------ K3/K4 generic on-the-fly SpMV
struct AccEmit {
	const double* __restrict__ y;
	double acc;
	__device__ void operator()(uint64_t c, double v) { acc += v * y[c]; }
};

// One thread per row: hops enumerated with ffs/popc, signs from masked popcounts, target states ranked on the fly
// (HubbardHelper.h:119-129 / FeBasedSc.h:66-105 / Heisenberg.h:94-106 as a gather).
__global__ void __launch_bounds__(LPP_TPB) k_spmv_generic(ModelDev m, SpmvArgs a)
{
	uint64_t t = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x;
	double contrib = 0.0;
	if (t < a.nloc) {
		uint64_t r = a.row0 + t;
		LppRowKets k = lpp_row_kets(m, r);
		double yr = a.y[r];
		AccEmit e{a.y, lpp_row_diag(m, k) * yr};
		lpp_row_offdiag(m, k, 0, e);
		double xn = a.alpha * e.acc;
		if (a.beta != 0.0) xn += a.beta * a.x[t];
		a.x[t] = xn;
		contrib = yr * xn;
	}
	if (a.dot_partials) {
		double s = lpp_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = s;
	}
}

// K4 fast path: Heisenberg S=1/2 with the couplings flattened into a bond list (p<q, jpm(p,q), jpm(q,p), jzz(p,q)).
// Same terms as Heisenberg.h:242-307: diagonal sum_{p<q} jzz m_p m_q (+ field, + D m^2), off-diagonal 1/2 jpm(i,j) with
// i the down site and j the up site; the target state is ranked with the split colex tables (2 look-ups).
__global__ void __launch_bounds__(LPP_TPB) k_spmv_heis(ModelDev m, const HeisBond* __restrict__ bonds, int nbonds, int has_field,
                                                      SpmvArgs a)
{
	uint64_t t = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x;
	double contrib = 0.0;
	if (t < a.nloc) {
		const uint64_t r = a.row0 + t;
		const word_t w = m.b1[r];
		const double* __restrict__ y = a.y;
		double diag = 0.0, acc = 0.0, acc2 = 0.0;
		if (has_field) {
			for (int i = 0; i < m.nsite; i++) {
				const double mi = (double)((w >> i) & 1) - 0.5;
				diag += m.V[i] * mi + m.D[i] * (mi * mi);
			}
		}
		for (int b = 0; b < nbonds; b++) {
			const HeisBond bd = bonds[b];
			const int bp = (int)((w >> bd.p) & 1), bq = (int)((w >> bd.q) & 1);
			diag += (bp == bq) ? 0.25 * bd.jzz : -0.25 * bd.jzz;
			if (bp != bq) {
				const double amp = 0.5 * (bp ? bd.jpm_qp : bd.jpm_pq);
				if (amp != 0.0) {
					const uint64_t c = lpp_rank_onespin(m, 0, w ^ (lpp_bit(bd.p) | lpp_bit(bd.q)));
					if (b & 1) acc2 += amp * y[c]; else acc += amp * y[c];
				}
			}
		}
		const double yr = y[r];
		double xn = a.alpha * (diag * yr + acc + acc2);
		if (a.beta != 0.0) xn += a.beta * a.x[t];
		a.x[t] = xn;
		contrib = yr * xn;
	}
	if (a.dot_partials) {
		double s = lpp_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = s;
	}
}

void lpp_launch_spmv_heis(const ModelDev& m, const HeisBond* bonds, int nbonds, int has_field, const SpmvArgs& a, cudaStream_t s)
{
	k_spmv_heis<<<lpp_spmv_generic_blocks(a.nloc), LPP_TPB, 0, s>>>(m, bonds, nbonds, has_field, a);
}

int lpp_spmv_generic_blocks(uint64_t nloc) { return (int)((nloc + LPP_TPB - 1) / LPP_TPB); }
void lpp_launch_spmv_generic(const ModelDev& m, const SpmvArgs& a, cudaStream_t s)
{
	k_spmv_generic<<<lpp_spmv_generic_blocks(a.nloc), LPP_TPB, 0, s>>>(m, a);
}

// ------------------------------------------------------------------ K3 table-driven SpMV for product bases
// bit-parallel form of findSnoDecay (FeBasedSc.h:573-623) without the potential part
__device__ __forceinline__ double lpp_feas_diag_fast(const ModelDev& m, word_t k1, word_t k2)
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

__device__ __forceinline__ double lpp_product_diag(const ModelDev& m, const DiagTables& dt, word_t k1, word_t k2,
                                                   uint64_t i1, uint64_t i2)
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
		s = lpp_feas_diag_fast(m, k1, k2);
	}
	return s + dt.dv1[i1] + dt.dv2[i2];
}

struct TwoAccEmit {
	const double* __restrict__ y;
	uint64_t n1;
	double acc;
	__device__ void operator()(uint64_t a, uint64_t b, double v) { acc += v * y[a + b * n1]; }
};

// thread (u, d): x[d,u] = beta x + alpha ( diag*y + sum_k up[k][u] y[d, upidx] + sum_k dn[k][d] y[dnidx, u] (+ two-spin) )
__global__ void __launch_bounds__(LPP_TPB) k_spmv_table(ModelDev m, HopTable up, HopTable dn, DiagTables dt, SpmvArgs a,
                                                       uint32_t nbx)
{
	const uint64_t dl = blockIdx.x / nbx;                 // local slow index
	const uint64_t u = (uint64_t)(blockIdx.x % nbx) * LPP_TPB + threadIdx.x;
	const uint64_t d = a.row0 / m.n1 + dl;
	double contrib = 0.0;
	if (u < m.n1) {
		const double* __restrict__ y = a.y;
		const uint64_t base = d * m.n1;
		word_t k1 = m.b1[u], k2 = m.b2[d];
		double yr = y[base + u];
		double acc = lpp_product_diag(m, dt, k1, k2, u, d) * yr;
		const int cd = (int)dn.cnt[d];
		for (int k = 0; k < cd; k++)
			acc += dn.val[(uint64_t)k * dn.n + d] * y[(uint64_t)dn.idx[(uint64_t)k * dn.n + d] * m.n1 + u];
		const int cu = (int)up.cnt[u];
		for (int k = 0; k < cu; k++)
			acc += up.val[(uint64_t)k * up.n + u] * y[base + up.idx[(uint64_t)k * up.n + u]];
		if (m.model == LPP_MODEL_FEAS) {
			TwoAccEmit e{y, m.n1, 0.0};
			lpp_feas_twospin(m, k1, k2, m.u3_all_pairs, e);
			acc += e.acc;
		}
		const uint64_t t = dl * m.n1 + u;
		double xn = a.alpha * acc;
		if (a.beta != 0.0) xn += a.beta * a.x[t];
		a.x[t] = xn;
		contrib = yr * xn;
	}
	if (a.dot_partials) {
		double s = lpp_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = s;
	}
}

int lpp_spmv_table_blocks(const ModelDev& m, uint64_t nloc)
{
	uint64_t nbx = (m.n1 + LPP_TPB - 1) / LPP_TPB;
	return (int)(nbx * (nloc / m.n1));
}
void lpp_launch_spmv_table(const ModelDev& m, const HopTable& up, const HopTable& dn, const DiagTables& dt,
                           const SpmvArgs& a, cudaStream_t s)
{
	uint32_t nbx = (uint32_t)((m.n1 + LPP_TPB - 1) / LPP_TPB);
	k_spmv_table<<<lpp_spmv_table_blocks(m, a.nloc), LPP_TPB, 0, s>>>(m, up, dn, dt, a, nbx);
}

// ------------------------------------------------------------------ K5 stored CRS
__global__ void __launch_bounds__(128) k_crs_count(ModelDev m, uint64_t row0, uint64_t nloc, int64_t* __restrict__ counts,
                                                   int* overflow)
{
	uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nloc) return;
	uint64_t c[LPP_ROW_CAP];
	double v[LPP_ROW_CAP];
	int n = lpp_stored_row(m, row0 + t, c, v);
	if (n < 0) { *overflow = 1; n = 0; }
	counts[t] = n;
}

__global__ void __launch_bounds__(128) k_crs_fill(ModelDev m, uint64_t row0, uint64_t nloc, const int64_t* __restrict__ rowptr,
                                                  int64_t* __restrict__ colind, double* __restrict__ values)
{
	uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nloc) return;
	uint64_t c[LPP_ROW_CAP];
	double v[LPP_ROW_CAP];
	int n = lpp_stored_row(m, row0 + t, c, v);
	int64_t p = rowptr[t];
	for (int i = 0; i < n; i++) { colind[p + i] = (int64_t)c[i]; values[p + i] = v[i]; }
}

void lpp_launch_crs_count(const ModelDev& m, uint64_t row0, uint64_t nloc, int64_t* counts, int* overflow, cudaStream_t s)
{
	k_crs_count<<<(unsigned)((nloc + 127) / 128), 128, 0, s>>>(m, row0, nloc, counts, overflow);
}
void lpp_launch_crs_fill(const ModelDev& m, uint64_t row0, uint64_t nloc, const int64_t* rowptr, int64_t* colind,
                         double* values, cudaStream_t s)
{
	k_crs_fill<<<(unsigned)((nloc + 127) / 128), 128, 0, s>>>(m, row0, nloc, rowptr, colind, values);
}

void lpp_exclusive_scan(int64_t* data, uint64_t n, int64_t* total_dev, cudaStream_t s)
{
	// data[0..n) counts, data[n] = 0 on entry; exclusive scan over n+1 slots puts the total in data[n]
	void* tmp = nullptr;
	size_t bytes = 0;
	cub::DeviceScan::ExclusiveSum(tmp, bytes, data, data, (int64_t)(n + 1), s);
	cudaMallocAsync(&tmp, bytes, s);
	cub::DeviceScan::ExclusiveSum(tmp, bytes, data, data, (int64_t)(n + 1), s);
	cudaFreeAsync(tmp, s);
	if (total_dev) cudaMemcpyAsync(total_dev, data + n, sizeof(int64_t), cudaMemcpyDeviceToDevice, s);
}

// PsimagLite::CrsMatrix::matrixVectorProduct: x += A y, LANES lanes per row
template <int LANES>
__global__ void __launch_bounds__(LPP_TPB) k_spmv_crs(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ colind,
                                                     const double* __restrict__ values, SpmvArgs a)
{
	const uint64_t t = ((uint64_t)blockIdx.x * LPP_TPB + threadIdx.x) / LANES;
	const int lane = threadIdx.x % LANES;
	double contrib = 0.0;
	double sum = 0.0;
	if (t < a.nloc) {
		for (int64_t k = rowptr[t] + lane; k < rowptr[t + 1]; k += LANES) sum += values[k] * a.y[colind[k]];
	}
#pragma unroll
	for (int o = LANES / 2; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o, LANES);
	if (t < a.nloc && lane == 0) {
		double xn = a.alpha * sum;
		if (a.beta != 0.0) xn += a.beta * a.x[t];
		a.x[t] = xn;
		contrib = a.y[a.row0 + t] * xn;
	}
	if (a.dot_partials) {
		double s = lpp_block_sum(contrib);
		if (threadIdx.x == 0) a.dot_partials[blockIdx.x] = s;
	}
}

#define LPP_CRS_LANES 8
int lpp_spmv_crs_blocks(uint64_t nloc) { return (int)((nloc * LPP_CRS_LANES + LPP_TPB - 1) / LPP_TPB); }
void lpp_launch_spmv_crs(const int64_t* rowptr, const int64_t* colind, const double* values, const SpmvArgs& a,
                         cudaStream_t s)
{
	k_spmv_crs<LPP_CRS_LANES><<<lpp_spmv_crs_blocks(a.nloc), LPP_TPB, 0, s>>>(rowptr, colind, values, a);
}

// ------------------------------------------------------------------ K6/K7/K8 Lanczos vector sweeps
#define LPP_VEC_MAXBLOCKS (148 * 8)
int lpp_vec_blocks(uint64_t n)
{
	uint64_t b = (n / 2 + LPP_TPB - 1) / LPP_TPB;
	if (b < 1) b = 1;
	return (int)(b > LPP_VEC_MAXBLOCKS ? LPP_VEC_MAXBLOCKS : b);
}

__global__ void __launch_bounds__(LPP_TPB) k_fill_random(double* __restrict__ v, uint64_t row0, uint64_t n, uint64_t seed)
{
	for (uint64_t i = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x; i < n; i += (uint64_t)gridDim.x * LPP_TPB)
		v[i] = lpp_splitmix_uniform(seed, row0 + i);
}

__global__ void __launch_bounds__(LPP_TPB) k_dot(const double* __restrict__ a, const double* __restrict__ b, uint64_t n,
                                                double* __restrict__ partials)
{
	double s = 0.0;
	const uint64_t n2 = n / 2;
	const double2* a2 = reinterpret_cast<const double2*>(a);
	const double2* b2 = reinterpret_cast<const double2*>(b);
	for (uint64_t i = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x; i < n2; i += (uint64_t)gridDim.x * LPP_TPB) {
		double2 p = a2[i], q = b2[i];
		s += p.x * q.x + p.y * q.y;
	}
	if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) s += a[n - 1] * b[n - 1];
	s = lpp_block_sum(s);
	if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// PsimagLite oneStepDecomposition sweep 2: x -= a y ; b2 += |x|^2
__global__ void __launch_bounds__(LPP_TPB) k_axpy_norm(double* __restrict__ x, const double* __restrict__ y, double coef,
                                                      const double* __restrict__ coef_dev, uint64_t n, double* __restrict__ partials)
{
	if (coef_dev) coef = *coef_dev;
	double s = 0.0;
	const uint64_t n2 = n / 2;
	double2* x2 = reinterpret_cast<double2*>(x);
	const double2* y2 = reinterpret_cast<const double2*>(y);
	for (uint64_t i = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x; i < n2; i += (uint64_t)gridDim.x * LPP_TPB) {
		double2 p = x2[i], q = y2[i];
		p.x -= coef * q.x;
		p.y -= coef * q.y;
		x2[i] = p;
		s += p.x * p.x + p.y * p.y;
	}
	if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
		double p = x[n - 1] - coef * y[n - 1];
		x[n - 1] = p;
		s += p * p;
	}
	s = lpp_block_sum(s);
	if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(LPP_TPB) k_axpy(double* __restrict__ z, const double* __restrict__ v, double coef, uint64_t n)
{
	for (uint64_t i = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x; i < n; i += (uint64_t)gridDim.x * LPP_TPB)
		z[i] += coef * v[i];
}

__global__ void __launch_bounds__(LPP_TPB) k_scale(double* __restrict__ v, double coef, uint64_t n)
{
	for (uint64_t i = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x; i < n; i += (uint64_t)gridDim.x * LPP_TPB)
		v[i] *= coef;
}

__global__ void __launch_bounds__(1024) k_finalize_sum(const double* __restrict__ partials, int n, double* __restrict__ out)
{
	double s = 0.0;
	for (int i = threadIdx.x; i < n; i += 1024) s += partials[i];
	s = lpp_block_sum(s);
	if (threadIdx.x == 0) out[0] = s;
}

// full reorthogonalisation (<prefix>Options=reortho), classical Gram-Schmidt against a block of saved Lanczos vectors:
// one pass computes the block's dot products with x, one pass subtracts the projections and accumulates |x|^2
__global__ void __launch_bounds__(LPP_TPB) k_reortho_dots(const double* __restrict__ x, RoVecs r, uint64_t n, double* __restrict__ partials)
{
	double s[LPP_RO_NV];
#pragma unroll
	for (int k = 0; k < LPP_RO_NV; k++) s[k] = 0.0;
	for (uint64_t i = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x; i < n; i += (uint64_t)gridDim.x * LPP_TPB) {
		const double xi = x[i];
#pragma unroll
		for (int k = 0; k < LPP_RO_NV; k++)
			if (k < r.nv) s[k] += r.v[k][i] * xi;
	}
#pragma unroll
	for (int k = 0; k < LPP_RO_NV; k++) {
		const double t = lpp_block_sum(s[k]);
		if (threadIdx.x == 0) partials[(uint64_t)k * gridDim.x + blockIdx.x] = t;
	}
}
__global__ void __launch_bounds__(LPP_TPB) k_reortho_axpy_norm(double* __restrict__ x, RoVecs r, uint64_t n, double* __restrict__ partials)
{
	double s = 0.0;
	for (uint64_t i = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x; i < n; i += (uint64_t)gridDim.x * LPP_TPB) {
		double t = x[i];
#pragma unroll
		for (int k = 0; k < LPP_RO_NV; k++)
			if (k < r.nv) t -= r.coef[k] * r.v[k][i];
		x[i] = t;
		s += t * t;
	}
	s = lpp_block_sum(s);
	if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
// out[k] = sum of the k-th row of partials[nv][n]
__global__ void __launch_bounds__(1024) k_finalize_sums(const double* __restrict__ partials, int n, double* __restrict__ out)
{
	double s = 0.0;
	for (int i = threadIdx.x; i < n; i += 1024) s += partials[(uint64_t)blockIdx.x * n + i];
	s = lpp_block_sum(s);
	if (threadIdx.x == 0) out[blockIdx.x] = s;
}
// Engine::twoPoint (Engine.h:262-331): result(i, j) = <m_j | m_i> for the modified states m_i = O_i |gs>.  One CTA column
// accumulates a 4 x 4 tile of the Gram matrix over a grid-stride range of elements (8 vectors read per pass).
__global__ void __launch_bounds__(LPP_TPB) k_gram_tile(const double* __restrict__ veci, const double* __restrict__ vecj, uint64_t stride,
                                                      uint64_t n, int nvec, int ti, int tj, double* __restrict__ partials)
{
	double acc[4][4];
#pragma unroll
	for (int a = 0; a < 4; a++)
#pragma unroll
		for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
	for (uint64_t e = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x; e < n; e += (uint64_t)gridDim.x * LPP_TPB) {
		double vi[4], vj[4];
#pragma unroll
		for (int a = 0; a < 4; a++) {
			vi[a] = (ti * 4 + a < nvec) ? veci[(uint64_t)(ti * 4 + a) * stride + e] : 0.0;
			vj[a] = (tj * 4 + a < nvec) ? vecj[(uint64_t)(tj * 4 + a) * stride + e] : 0.0;
		}
#pragma unroll
		for (int a = 0; a < 4; a++)
#pragma unroll
			for (int b = 0; b < 4; b++) acc[a][b] = fma(vi[a], vj[b], acc[a][b]);
	}
#pragma unroll
	for (int a = 0; a < 4; a++)
#pragma unroll
		for (int b = 0; b < 4; b++) {
			const double t = lpp_block_sum(acc[a][b]);
			if (threadIdx.x == 0) partials[(uint64_t)(a * 4 + b) * gridDim.x + blockIdx.x] = t;
		}
}
// <bra| prod ops |ket> without materialising the intermediate vector: every source row contributes bra[target] * factor * ket[row]
__global__ void __launch_bounds__(LPP_TPB) k_measure(ModelDev m, LppMeasureOps ops, const double* __restrict__ bra, const double* __restrict__ ket,
                                                    uint64_t row0, uint64_t nloc, double* __restrict__ partials)
{
	double s = 0.0;
	for (uint64_t t = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x; t < nloc; t += (uint64_t)gridDim.x * LPP_TPB) {
		const LppRowKets k = lpp_row_kets(m, row0 + t);
		word_t o1, o2;
		double v;
		if (!lpp_rahul_apply(ops, k.k1, k.k2, &o1, &o2, &v)) continue;
		if (m.model == LPP_MODEL_TJ && (o1 & o2)) continue;            // left the no-double-occupancy space
		const uint64_t target = (o1 == k.k1 && o2 == k.k2) ? row0 + t : lpp_rank_pair(m, o1, o2);
		s += bra[target - row0] * v * ket[t];
	}
	s = lpp_block_sum(s);
	if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
void lpp_launch_measure(const ModelDev& m, const LppMeasureOps& ops, const double* bra, const double* ket, uint64_t row0, uint64_t nloc,
                        double* partials, cudaStream_t s)
{
	k_measure<<<lpp_vec_blocks(nloc * 2), LPP_TPB, 0, s>>>(m, ops, bra, ket, row0, nloc, partials);
}

void lpp_launch_gram_tile(const double* veci, const double* vecj, uint64_t stride, uint64_t n, int nvec, int ti, int tj, double* partials,
                          cudaStream_t s)
{
	k_gram_tile<<<lpp_vec_blocks(n * 2), LPP_TPB, 0, s>>>(veci, vecj, stride, n, nvec, ti, tj, partials);
}

void lpp_launch_reortho_dots(const double* x, const RoVecs& r, uint64_t n, double* partials, cudaStream_t s)
{
	k_reortho_dots<<<lpp_vec_blocks(n * 2), LPP_TPB, 0, s>>>(x, r, n, partials);
}
void lpp_launch_reortho_axpy_norm(double* x, const RoVecs& r, uint64_t n, double* partials, cudaStream_t s)
{
	k_reortho_axpy_norm<<<lpp_vec_blocks(n * 2), LPP_TPB, 0, s>>>(x, r, n, partials);
}
void lpp_launch_finalize_sums(const double* partials, int n, int nv, double* out, cudaStream_t s)
{
	k_finalize_sums<<<nv, 1024, 0, s>>>(partials, n, out);
}

void lpp_launch_fill_random(double* v, uint64_t row0, uint64_t n, uint64_t seed, cudaStream_t s)
{
	k_fill_random<<<lpp_vec_blocks(n * 2), LPP_TPB, 0, s>>>(v, row0, n, seed);
}
void lpp_launch_dot(const double* a, const double* b, uint64_t n, double* partials, cudaStream_t s)
{
	k_dot<<<lpp_vec_blocks(n), LPP_TPB, 0, s>>>(a, b, n, partials);
}
void lpp_launch_axpy_norm(double* x, const double* y, double coef, uint64_t n, double* partials, cudaStream_t s, const double* coef_dev)
{
	k_axpy_norm<<<lpp_vec_blocks(n), LPP_TPB, 0, s>>>(x, y, coef, coef_dev, n, partials);
}

// device-resident Lanczos scalars (one thread): the same operations, in the same order, as the host loop performs
__global__ void k_lz_init(double nj, double* coefs)
{
	coefs[LPP_LZ_NORM] = nj;
	coefs[LPP_LZ_ALPHA] = 1.0 / nj;
	coefs[LPP_LZ_BETA] = 0.0;
	coefs[LPP_LZ_AXPY] = 0.0;
	coefs[LPP_LZ_STEP] = 0.0;
}
__global__ void k_lz_after_dot(const double* __restrict__ dot_parts, int nparts, double* coefs, double* __restrict__ a_out)
{
	double dot = dot_parts[0];
	for (int i = 1; i < nparts; i++) dot += dot_parts[i];
	const double nj = coefs[LPP_LZ_NORM];
	const double aj = dot / nj;
	a_out[(int)coefs[LPP_LZ_STEP]] = aj;
	coefs[LPP_LZ_AXPY] = aj / nj;
}
__global__ void k_lz_after_norm(const double* __restrict__ b2, double* coefs, double* __restrict__ b_out)
{
	const double bj = sqrt(*b2);
	const int j = (int)coefs[LPP_LZ_STEP];
	b_out[j] = bj;
	const double nprev = coefs[LPP_LZ_NORM];
	const double nj = (bj < 1e-10) ? 1.0 : bj;
	coefs[LPP_LZ_NORM] = nj;
	coefs[LPP_LZ_ALPHA] = 1.0 / nj;
	coefs[LPP_LZ_BETA] = -(bj / nprev);
	coefs[LPP_LZ_STEP] = (double)(j + 1);
}
__global__ void k_lzp_init(double nj, double* coefs)
{
	coefs[LPP_LZ_NORM] = nj;
	coefs[LPP_LZP_C1] = 1.0 / nj;
	coefs[LPP_LZP_C2] = 0.0;
	coefs[LPP_LZP_C3] = 0.0;
	coefs[LPP_LZP_STEPA] = 0.0;
	coefs[LPP_LZP_STEPB] = 0.0;
}
// dot_parts[0] + dot_parts[1] = <y, H y> with y = U_j = n_j v_j:  a_j = dot / n_j^2
__global__ void k_lzp_after_dot(const double* __restrict__ dot_parts, double* coefs, double* __restrict__ a_out)
{
	const double nj = coefs[LPP_LZ_NORM];
	const double aj = (dot_parts[0] + dot_parts[1]) / nj / nj;
	const int j = (int)coefs[LPP_LZP_STEPA];
	a_out[j] = aj;
	coefs[LPP_LZP_C1] = 1.0 / nj;
	coefs[LPP_LZP_C2] = aj / nj;
	coefs[LPP_LZP_STEPA] = (double)(j + 1);
}
// *b2 = |U_{j+1}|^2:  b_j = its root, the next n, and the weight b_j / n_j of U_j in the step after
__global__ void k_lzp_after_norm(const double* __restrict__ b2, double* coefs, double* __restrict__ b_out)
{
	const double bj = sqrt(*b2);
	const int j = (int)coefs[LPP_LZP_STEPB];
	b_out[j] = bj;
	const double nprev = coefs[LPP_LZ_NORM];
	const double nj = (bj < 1e-10) ? 1.0 : bj;
	coefs[LPP_LZ_NORM] = nj;
	coefs[LPP_LZP_C3] = bj / nprev;
	coefs[LPP_LZP_STEPB] = (double)(j + 1);
}
void lpp_launch_lzp_init(double nj, double* coefs, cudaStream_t s) { k_lzp_init<<<1, 1, 0, s>>>(nj, coefs); }
void lpp_launch_lzp_after_dot(const double* dot_parts, double* coefs, double* a_out, cudaStream_t s) { k_lzp_after_dot<<<1, 1, 0, s>>>(dot_parts, coefs, a_out); }
void lpp_launch_lzp_after_norm(const double* b2, double* coefs, double* b_out, cudaStream_t s) { k_lzp_after_norm<<<1, 1, 0, s>>>(b2, coefs, b_out); }
void lpp_launch_lz_init(double nj, double* coefs, cudaStream_t s) { k_lz_init<<<1, 1, 0, s>>>(nj, coefs); }
void lpp_launch_lz_after_dot(const double* dot_parts, int nparts, double* coefs, double* a_out, cudaStream_t s)
{
	k_lz_after_dot<<<1, 1, 0, s>>>(dot_parts, nparts, coefs, a_out);
}
void lpp_launch_lz_after_norm(const double* b2, double* coefs, double* b_out, cudaStream_t s) { k_lz_after_norm<<<1, 1, 0, s>>>(b2, coefs, b_out); }
void lpp_launch_axpy(double* z, const double* v, double coef, uint64_t n, cudaStream_t s)
{
	k_axpy<<<lpp_vec_blocks(n * 2), LPP_TPB, 0, s>>>(z, v, coef, n);
}
void lpp_launch_scale(double* v, double coef, uint64_t n, cudaStream_t s)
{
	k_scale<<<lpp_vec_blocks(n * 2), LPP_TPB, 0, s>>>(v, coef, n);
}
void lpp_launch_finalize_sum(const double* partials, int n, double* out, cudaStream_t s)
{
	k_finalize_sum<<<1, 1024, 0, s>>>(partials, n, out);
}

// ------------------------------------------------------------------ K9 operator application
// Engine.h:416-458 as a gather on the destination basis: every destination state has at most one source.
// c:       dst word has the site empty, source = dst | bit      (BasisOneSpin.h:127-134)
// cdagger: dst word has the site occupied, source = dst ^ bit   (:135-142)
// n:       same sector, site occupied                            (:143-147)
__global__ void __launch_bounds__(LPP_TPB) k_apply_op(ModelDev src, ModelDev dst, int op, int site, int spin, int orb, double factor,
                                                     const double* __restrict__ srcv, double* __restrict__ z,
                                                     uint64_t dst_row0, uint64_t dst_nloc)
{
	uint64_t t = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x;
	if (t >= dst_nloc) return;
	uint64_t srow;
	double sg;
	if (!lpp_apply_op_source(src, dst, op, site, spin, orb, dst_row0 + t, &srow, &sg)) return;
	z[t] += factor * sg * srcv[srow];
}

void lpp_launch_apply_op(const ModelDev& src, const ModelDev& dst, int op, int site, int spin, int orb, double factor,
                         const double* srcv, double* z, uint64_t dst_row0, uint64_t dst_nloc, cudaStream_t s)
{
	k_apply_op<<<(unsigned)((dst_nloc + LPP_TPB - 1) / LPP_TPB), LPP_TPB, 0, s>>>(src, dst, op, site, spin, orb, factor, srcv, z,
	                                                                            dst_row0, dst_nloc);
}

// ------------------------------------------------------------------ two-layout exchange helpers (multi-GPU)
// ROW shard (nrows x n1, local down range) <-> COLUMN shards: peer q owns columns [cs[q], cs[q+1]).
// Buffers hold one block per peer, block q = row-major nrows x ncols_q at element offset nrows*cs[q].
__device__ __forceinline__ int lpp_col_owner(const ColSplit& c, uint64_t u)
{
	int q = 0;
	while (q + 1 < c.nranks && u >= c.cs[q + 1]) q++;
	return q;
}

// y_row -> sendbuf blocks; the rank's own block goes straight into its column shard ycol (rows d0loc.., pitch ncols_me)
__global__ void __launch_bounds__(LPP_TPB) k_pack_cols(const double* __restrict__ src, double* __restrict__ sendbuf,
                                                      double* __restrict__ ycol, uint64_t nrows, uint64_t n1, ColSplit c,
                                                      uint64_t d0loc)
{
	const uint64_t total = nrows * n1;
	for (uint64_t i = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x; i < total; i += (uint64_t)gridDim.x * LPP_TPB) {
		const uint64_t r = i / n1, u = i - r * n1;
		const int q = lpp_col_owner(c, u);
		const uint64_t nc = c.cs[q + 1] - c.cs[q], cu = u - c.cs[q];
		const double v = src[i];
		if (q == c.me) ycol[(d0loc + r) * nc + cu] = v;
		else sendbuf[nrows * c.cs[q] + r * nc + cu] = v;
	}
}

// x_row += blocks received from the column shards (own block read from xcol)
__global__ void __launch_bounds__(LPP_TPB) k_unpack_add(double* __restrict__ x, const double* __restrict__ recvbuf,
                                                       const double* __restrict__ xcol, uint64_t nrows, uint64_t n1, ColSplit c,
                                                       uint64_t d0loc)
{
	const uint64_t total = nrows * n1;
	for (uint64_t i = (uint64_t)blockIdx.x * LPP_TPB + threadIdx.x; i < total; i += (uint64_t)gridDim.x * LPP_TPB) {
		const uint64_t r = i / n1, u = i - r * n1;
		const int q = lpp_col_owner(c, u);
		const uint64_t nc = c.cs[q + 1] - c.cs[q], cu = u - c.cs[q];
		const double v = (q == c.me) ? xcol[(d0loc + r) * nc + cu] : recvbuf[nrows * c.cs[q] + r * nc + cu];
		x[i] += v;
	}
}

// Peer-memory variants (NVLink, CUDA IPC mappings): the pack kernel stores each element straight into the owning rank's
// column shard, the unpack kernel loads the column results from the owning rank -- transfer and re-layout are one kernel,
// no staging buffers and no library collective on the data path.
// one block row per matrix row (blockIdx.x; no integer division), each thread walks columns with stride blockDim * gridDim.y
__global__ void __launch_bounds__(LPP_TPB) k_pack_cols_p2p(const double* __restrict__ src, const __grid_constant__ PeerPtrs ycols, uint64_t nrows,
                                                          uint64_t n1, const __grid_constant__ ColSplit c, uint64_t d0loc)
{
	const uint64_t r = blockIdx.x;
	const double* __restrict__ srow = src + r * n1;
	for (uint64_t u = (uint64_t)blockIdx.y * LPP_TPB + threadIdx.x; u < n1; u += (uint64_t)gridDim.y * LPP_TPB) {
		const int q = lpp_col_owner(c, u);
		const uint64_t nc = c.cs[q + 1] - c.cs[q], cu = u - c.cs[q];
		ycols.p[q][(d0loc + r) * nc + cu] = srow[u];
	}
}

__global__ void __launch_bounds__(LPP_TPB) k_unpack_add_p2p(double* __restrict__ x, const __grid_constant__ PeerPtrs xcols, uint64_t nrows, uint64_t n1,
                                                           const __grid_constant__ ColSplit c, uint64_t d0loc)
{
	const uint64_t r = blockIdx.x;
	double* __restrict__ xrow = x + r * n1;
	for (uint64_t u = (uint64_t)blockIdx.y * LPP_TPB + threadIdx.x; u < n1; u += (uint64_t)gridDim.y * LPP_TPB) {
		const int q = lpp_col_owner(c, u);
		const uint64_t nc = c.cs[q + 1] - c.cs[q], cu = u - c.cs[q];
		xrow[u] += xcols.p[q][(d0loc + r) * nc + cu];
	}
}

// x <- x + (peers' column-shard results) - coef * y ;  partials <- block sums of x_new^2.
// The re-layout of the down-sweep result and the Lanczos "x -= a y, b^2 = |x|^2" sweep in one pass over the row shard.
// PACK also stores the new vector into the owners' column shards (the pack of the NEXT mat-vec; opt-in, it measured slower).
// (A variant with one owner per block and 16-byte accesses measured 0.84-0.93 ms against 0.70 ms for this one on 2 x B200.)
template <bool PACK>
__global__ void __launch_bounds__(LPP_TPB) k_unpack_axpy_norm_p2p(double* __restrict__ x, const double* __restrict__ y, double coef,
                                                                 const double* __restrict__ coef_dev, const __grid_constant__ PeerPtrs xcols,
                                                                 const __grid_constant__ PeerPtrs ycols, uint64_t nrows, uint64_t n1,
                                                                 const __grid_constant__ ColSplit c, uint64_t d0loc,
                                                                 double* __restrict__ partials)
{
	if (coef_dev) coef = *coef_dev;
	const uint64_t r = blockIdx.x;
	double* __restrict__ xrow = x + r * n1;
	const double* __restrict__ yrow = y + r * n1;
	double s = 0.0;
	for (uint64_t u = (uint64_t)blockIdx.y * LPP_TPB + threadIdx.x; u < n1; u += (uint64_t)gridDim.y * LPP_TPB) {
		const int q = lpp_col_owner(c, u);
		const uint64_t nc = c.cs[q + 1] - c.cs[q], cu = u - c.cs[q];
		const double v = xrow[u] + xcols.p[q][(d0loc + r) * nc + cu] - coef * yrow[u];
		xrow[u] = v;
		if (PACK) ycols.p[q][(d0loc + r) * nc + cu] = v;
		s += v * v;
	}
	s = lpp_block_sum(s);
	if (threadIdx.x == 0) partials[(uint64_t)blockIdx.x * gridDim.y + blockIdx.y] = s;
}

// The same sweep with 16-byte accesses (n1 even: every shard boundary and every row start is a multiple of two columns).
// Peer loads are request bound over NVLink: 8-byte loads measured 354 GB/s per GPU on 8 x B200 (0.41 ms for 145 MB), the
// 16-byte form moves twice the bytes per request; two independent pairs per thread keep more requests in flight.
// coef is read from `coef_dev` when non-null (device-resident recurrence: no host round trip before this sweep).
// PACK: the new vector is also stored into the owners' column shards (the pack of the NEXT mat-vec): NVLink carries the loads
// of this sweep in one direction and the stores in the other at the same time, and the next up sweep runs without the copy
// engines next to it.
template <bool PACK>
__global__ void __launch_bounds__(LPP_TPB) k_unpack_axpy_norm_p2p_v2(double* __restrict__ x, const double* __restrict__ y, double coef,
                                                                    const double* __restrict__ coef_dev, const __grid_constant__ PeerPtrs xcols,
                                                                    const __grid_constant__ PeerPtrs ycols, uint64_t nrows, uint64_t n1,
                                                                    const __grid_constant__ ColSplit c, uint64_t d0loc,
                                                                    double* __restrict__ partials)
{
	const uint64_t r = blockIdx.x;
	if (coef_dev) coef = *coef_dev;
	double2* __restrict__ xrow = reinterpret_cast<double2*>(x + r * n1);
	const double2* __restrict__ yrow = reinterpret_cast<const double2*>(y + r * n1);
	const uint64_t npair = n1 >> 1;
	const uint64_t stride = (uint64_t)gridDim.y * LPP_TPB;
	// Every rank walks the columns starting behind its OWN shard (rank r reads peer r+1 first, then r+2, ...): with all ranks
	// starting at column 0 the eight GPUs read the same peer at the same time and share one NVLink egress (measured on
	// 8 x B200: 0.45-0.59 ms per sweep against 0.28 ms for the one rank whose first shard is local).
	const uint64_t rot = c.cs[(c.me + 1) % c.nranks] >> 1;
	double s = 0.0;
	for (uint64_t l0 = (uint64_t)blockIdx.y * LPP_TPB + threadIdx.x; l0 < npair; l0 += 2 * stride) {
		const uint64_t l1 = l0 + stride;
		const bool two = l1 < npair;
		uint64_t p0 = l0 + rot, p1 = (two ? l1 : l0) + rot;
		if (p0 >= npair) p0 -= npair;
		if (p1 >= npair) p1 -= npair;
		const uint64_t u0 = 2 * p0, u1 = 2 * p1;
		const int q0 = lpp_col_owner(c, u0), q1 = lpp_col_owner(c, u1);
		const uint64_t nc0 = c.cs[q0 + 1] - c.cs[q0], nc1 = c.cs[q1 + 1] - c.cs[q1];
		const double2 r0 = *reinterpret_cast<const double2*>(xcols.p[q0] + (d0loc + r) * nc0 + (u0 - c.cs[q0]));
		const double2 r1 = *reinterpret_cast<const double2*>(xcols.p[q1] + (d0loc + r) * nc1 + (u1 - c.cs[q1]));
		const double2 xa = xrow[p0], ya = yrow[p0];
		double2 va = make_double2(xa.x + r0.x - coef * ya.x, xa.y + r0.y - coef * ya.y);
		xrow[p0] = va;
		if (PACK) *reinterpret_cast<double2*>(ycols.p[q0] + (d0loc + r) * nc0 + (u0 - c.cs[q0])) = va;
		s += va.x * va.x + va.y * va.y;
		if (two) {
			const double2 xb = xrow[p1], yb = yrow[p1];
			double2 vb = make_double2(xb.x + r1.x - coef * yb.x, xb.y + r1.y - coef * yb.y);
			xrow[p1] = vb;
			if (PACK) *reinterpret_cast<double2*>(ycols.p[q1] + (d0loc + r) * nc1 + (u1 - c.cs[q1])) = vb;
			s += vb.x * vb.x + vb.y * vb.y;
		}
	}
	s = lpp_block_sum(s);
	if (threadIdx.x == 0) partials[(uint64_t)blockIdx.x * gridDim.y + blockIdx.y] = s;
}

// Pipelined recurrence: x = C1 (z + pieces) - C2 y - C3 x on a block of rows, 16-byte accesses, column walk rotated per rank as in
// k_unpack_axpy_norm_p2p_v2; partial sums of squares per (row, column chunk).
__global__ void __launch_bounds__(LPP_TPB) k_unpack3_norm_p2p(double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                                             const double* __restrict__ coefs, const __grid_constant__ PeerPtrs xcols,
                                                             uint64_t nrows, uint64_t n1, const __grid_constant__ ColSplit c, uint64_t d0loc,
                                                             double* __restrict__ partials)
{
	const double c1 = coefs[LPP_LZP_C1], c2 = coefs[LPP_LZP_C2], c3 = coefs[LPP_LZP_C3];
	const bool use_x = c3 != 0.0;                     // first step: x is uninitialised memory, not a vector
	const uint64_t npair = n1 >> 1;
	const uint64_t stride = (uint64_t)gridDim.y * LPP_TPB;
	const uint64_t rot = c.cs[(c.me + 1) % c.nranks] >> 1;
	// the grid may hold fewer CTA rows than the block has matrix rows (a small footprint beside the up sweep): stride over the rows
	for (uint64_t r = blockIdx.x; r < nrows; r += gridDim.x) {
		double2* __restrict__ xrow = reinterpret_cast<double2*>(x + r * n1);
		const double2* __restrict__ yrow = reinterpret_cast<const double2*>(y + r * n1);
		const double2* __restrict__ zrow = reinterpret_cast<const double2*>(z + r * n1);
		double s = 0.0;
		for (uint64_t l0 = (uint64_t)blockIdx.y * LPP_TPB + threadIdx.x; l0 < npair; l0 += 2 * stride) {
			const uint64_t l1 = l0 + stride;
			const bool two = l1 < npair;
			uint64_t p0 = l0 + rot, p1 = (two ? l1 : l0) + rot;
			if (p0 >= npair) p0 -= npair;
			if (p1 >= npair) p1 -= npair;
			const uint64_t u0 = 2 * p0, u1 = 2 * p1;
			const int q0 = lpp_col_owner(c, u0), q1 = lpp_col_owner(c, u1);
			const uint64_t nc0 = c.cs[q0 + 1] - c.cs[q0], nc1 = c.cs[q1 + 1] - c.cs[q1];
			const double2 r0 = *reinterpret_cast<const double2*>(xcols.p[q0] + (d0loc + r) * nc0 + (u0 - c.cs[q0]));
			const double2 r1 = *reinterpret_cast<const double2*>(xcols.p[q1] + (d0loc + r) * nc1 + (u1 - c.cs[q1]));
			const double2 ya = yrow[p0], za = zrow[p0];
			const double2 xa = use_x ? xrow[p0] : make_double2(0.0, 0.0);
			const double2 va = make_double2(c1 * (za.x + r0.x) - c2 * ya.x - c3 * xa.x, c1 * (za.y + r0.y) - c2 * ya.y - c3 * xa.y);
			xrow[p0] = va;
			s += va.x * va.x + va.y * va.y;
			if (two) {
				const double2 yb = yrow[p1], zb = zrow[p1];
				const double2 xb = use_x ? xrow[p1] : make_double2(0.0, 0.0);
				const double2 vb = make_double2(c1 * (zb.x + r1.x) - c2 * yb.x - c3 * xb.x, c1 * (zb.y + r1.y) - c2 * yb.y - c3 * xb.y);
				xrow[p1] = vb;
				s += vb.x * vb.x + vb.y * vb.y;
			}
		}
		s = lpp_block_sum(s);
		if (threadIdx.x == 0) partials[r * gridDim.y + blockIdx.y] = s;
		__syncthreads();                               // lpp_block_sum's shared scratch is reused by the next row
	}
}

__global__ void __launch_bounds__(32) k_psx_allreduce(double* __restrict__ vals, int nvals, const __grid_constant__ PeerPtrs areas, int me,
                                                      int nranks, unsigned long long seq, int* __restrict__ err)
{
	__shared__ double sv[LPP_MAX_RANKS][4];
	const int q = threadIdx.x;
	const unsigned par = (unsigned)(seq & 1ull);
	if (q < nranks) {
		volatile double* dst = areas.p[q] + ((size_t)par * LPP_MAX_RANKS + me) * LPP_PSX_SLOT_DOUBLES;
		for (int i = 0; i < nvals; i++) dst[i] = vals[i];
		__threadfence_system();
		*reinterpret_cast<volatile unsigned long long*>(dst + 4) = seq;
		volatile double* src = areas.p[me] + ((size_t)par * LPP_MAX_RANKS + q) * LPP_PSX_SLOT_DOUBLES;
		const long long t0 = clock64();
		bool ok = true;
		while (*reinterpret_cast<volatile unsigned long long*>(src + 4) != seq) {
			if (clock64() - t0 > (1ll << 32)) { ok = false; break; }
		}
		__threadfence_system();
		if (!ok) {
			*err = 1;
			printf("[lpp psx] rank %d timed out waiting for rank %d: want seq %llu, slot holds %llu (nvals %d)\n", me, q, seq,
			       *reinterpret_cast<volatile unsigned long long*>(src + 4), nvals);
		}
		for (int i = 0; i < nvals; i++) sv[q][i] = src[i];
	}
	__syncthreads();
	if (threadIdx.x < nvals) {
		double s = 0.0;
		for (int r = 0; r < nranks; r++) s += sv[r][threadIdx.x];
		vals[threadIdx.x] = s;
	}
}
void lpp_launch_psx_allreduce(double* vals, int nvals, const PeerPtrs& areas, int me, int nranks, unsigned long long seq, int* err,
                              cudaStream_t s)
{
	k_psx_allreduce<<<1, 32, 0, s>>>(vals, nvals, areas, me, nranks, seq, err);
}

static dim3 lpp_rowwise_grid(uint64_t nrows, uint64_t n1)
{
	// rows in grid.x (up to 2^31-1), column chunks in grid.y: a rank's row shard can exceed the 65 535 limit of grid.y
	unsigned gy = (unsigned)((n1 + (uint64_t)LPP_TPB * 4 - 1) / ((uint64_t)LPP_TPB * 4));
	return dim3((unsigned)nrows, gy ? gy : 1, 1);
}

void lpp_launch_pack_cols_p2p(const double* src, const PeerPtrs& ycols, uint64_t nrows, uint64_t n1, const ColSplit& c,
                              uint64_t d0loc, cudaStream_t s)
{
	k_pack_cols_p2p<<<lpp_rowwise_grid(nrows, n1), LPP_TPB, 0, s>>>(src, ycols, nrows, n1, c, d0loc);
}
void lpp_launch_unpack_add_p2p(double* x, const PeerPtrs& xcols, uint64_t nrows, uint64_t n1, const ColSplit& c, uint64_t d0loc,
                               cudaStream_t s)
{
	k_unpack_add_p2p<<<lpp_rowwise_grid(nrows, n1), LPP_TPB, 0, s>>>(x, xcols, nrows, n1, c, d0loc);
}

int lpp_unpack_axpy_norm_blocks(uint64_t nrows, uint64_t n1, int nranks)
{
	(void)nranks;
	dim3 g = lpp_rowwise_grid(nrows, n1);
	return (int)(g.x * g.y);
}
void lpp_launch_unpack_axpy_norm_p2p(double* x, const double* y, double coef, const PeerPtrs& xcols, const PeerPtrs* ycols_or_null,
                                     uint64_t nrows, uint64_t n1, const ColSplit& c, uint64_t d0loc, double* partials, cudaStream_t s,
                                     const double* coef_dev)
{
	PeerPtrs yc;
	for (int i = 0; i < LPP_MAX_RANKS; i++) yc.p[i] = ycols_or_null ? ycols_or_null->p[i] : nullptr;
	const dim3 g = lpp_rowwise_grid(nrows, n1);
	static const bool v2 = !(getenv("LPP_UNPACK_V2") && getenv("LPP_UNPACK_V2")[0] == '0');
	bool even = (n1 % 2 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0);
	for (int q = 0; q <= c.nranks; q++) even = even && (c.cs[q] % 2 == 0);
	if (ycols_or_null && v2 && even) k_unpack_axpy_norm_p2p_v2<true><<<g, LPP_TPB, 0, s>>>(x, y, coef, coef_dev, xcols, yc, nrows, n1, c, d0loc, partials);
	else if (ycols_or_null) k_unpack_axpy_norm_p2p<true><<<g, LPP_TPB, 0, s>>>(x, y, coef, coef_dev, xcols, yc, nrows, n1, c, d0loc, partials);
	else if (v2 && even) k_unpack_axpy_norm_p2p_v2<false><<<g, LPP_TPB, 0, s>>>(x, y, coef, coef_dev, xcols, yc, nrows, n1, c, d0loc, partials);
	else k_unpack_axpy_norm_p2p<false><<<g, LPP_TPB, 0, s>>>(x, y, coef, coef_dev, xcols, yc, nrows, n1, c, d0loc, partials);
}

int lpp_unpack3_partials_per_row(uint64_t n1) { return (int)lpp_rowwise_grid(1, n1).y; }
void lpp_launch_unpack3_norm_p2p(double* x, const double* y, const double* z, const double* coefs, const PeerPtrs& xcols, uint64_t nrows,
                                 uint64_t n1, const ColSplit& c, uint64_t d0loc, double* partials, int max_cta_rows, cudaStream_t s)
{
	if (nrows == 0) return;
	dim3 g = lpp_rowwise_grid(nrows, n1);
	if (max_cta_rows > 0 && g.x > (unsigned)max_cta_rows) g.x = (unsigned)max_cta_rows;
	k_unpack3_norm_p2p<<<g, LPP_TPB, 0, s>>>(x, y, z, coefs, xcols, nrows, n1, c, d0loc, partials);
}
void lpp_launch_pack_cols(const double* src, double* sendbuf, double* ycol, uint64_t nrows, uint64_t n1, const ColSplit& c,
                          uint64_t d0loc, cudaStream_t s)
{
	k_pack_cols<<<lpp_vec_blocks(nrows * n1 * 2), LPP_TPB, 0, s>>>(src, sendbuf, ycol, nrows, n1, c, d0loc);
}
void lpp_launch_unpack_add(double* x, const double* recvbuf, const double* xcol, uint64_t nrows, uint64_t n1, const ColSplit& c,
                           uint64_t d0loc, cudaStream_t s)
{
	k_unpack_add<<<lpp_vec_blocks(nrows * n1 * 2), LPP_TPB, 0, s>>>(x, recvbuf, xcol, nrows, n1, c, d0loc);
}
