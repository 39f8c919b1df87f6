// lpp_sweep_common.cuh -- pieces shared by the product-basis sweep kernels (lpp_tiled.cu, lpp_dtile.cu, lpp_urow.cu).
#pragma once
#include "lpp_kernels.cuh"

#define LPP_MAXMAG 16

// distinct hop magnitudes |h(i,j)| of the model; table entries carry an index into it plus a sign bit
struct MagTable {
	double mag[LPP_MAXMAG];
	int nmag;
};

// ColView: a kernel sees an (nrows x ncols) row-major matrix with row pitch `pitch`; column c is up state u0 + c.
// Single GPU: pitch = ncols = Nup, u0 = 0.  Two-layout multi-GPU: the rank's column shard, all Ndn rows.
struct ColView {
	uint64_t pitch, ncols, u0;
};

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

// FeAs INT_PAPER33 diagonal (FeBasedSc.h:573-623) from the two spin words
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

// diagonal element of the product-basis models: two-spin part from the words, one-spin potentials from tables
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
