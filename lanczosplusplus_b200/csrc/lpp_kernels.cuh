// lpp_kernels.cuh -- launch wrappers of the sm_100a kernels (definitions in lpp_kernels.cu, lpp_tiled.cu).
#pragma once
#include <cuda_runtime.h>
#include "lpp_device.cuh"

// per-spin hop table in ELL form, column-major: entry k of one-spin state s is idx[k*n+s], val[k*n+s];
// padded entries have val = 0 and idx = s.
struct HopTable {
	uint32_t* idx;
	double* val;
	uint32_t* cnt;
	int width;
	uint64_t n;
};

// per-spin diagonal pieces of product-basis models
struct DiagTables {
	const double* dv1;   // sum_i V_i n_i for each up state
	const double* dv2;   // same, down
	int uniformU;        // Hubbard: all U equal -> U0 * popc(up & dn)
	double U0;
};

// A coefficient that is either a host value or lives in device memory (the Lanczos loop keeps 1/n_j, -b_{j-1}/n_{j-1} and a_j/n_j
// on the device so that an iteration needs no host round trip).  Kernels read it like a double.
struct LppCoef {
	double v;
	const double* p;
	__host__ __device__ LppCoef& operator=(double x) { v = x; p = nullptr; return *this; }
	__host__ __device__ operator double() const
	{
#ifdef __CUDA_ARCH__
		return p ? *p : v;
#else
		return v;                // the host never dereferences the device copy
#endif
	}
};

struct SpmvArgs {
	LppCoef alpha, beta;     // x = beta*x + alpha*(H y)
	double* x;               // local rows
	const double* y;         // full vector (global indexing)
	double* dot_partials;    // optional: per-block partial sums of y_local . x_new
	uint64_t row0, nloc;     // first global row and number of local rows
};

#define LPP_RO_NV 4
struct RoVecs {              // a block of saved Lanczos vectors and the projections to subtract
	const double* v[LPP_RO_NV];
	double coef[LPP_RO_NV];
	int nv;
};
void lpp_launch_reortho_dots(const double* x, const RoVecs& r, uint64_t n, double* partials, cudaStream_t s);
void lpp_launch_reortho_axpy_norm(double* x, const RoVecs& r, uint64_t n, double* partials, cudaStream_t s);
void lpp_launch_measure(const ModelDev& m, const LppMeasureOps& ops, const double* bra, const double* ket, uint64_t row0, uint64_t nloc,
                        double* partials, cudaStream_t s);
void lpp_launch_gram_tile(const double* veci, const double* vecj, uint64_t stride, uint64_t n, int nvec, int ti, int tj, double* partials,
                          cudaStream_t s);
void lpp_launch_finalize_sums(const double* partials, int n, int nv, double* out, cudaStream_t s);
void lpp_launch_build_colex(const uint64_t* binom, int nsite, int npart, uint64_t n, word_t* out, cudaStream_t s);
void lpp_launch_build_feas(const ModelDev& m, int spin, uint64_t n, word_t* out, cudaStream_t s);
void lpp_launch_rank(const ModelDev& m, int spin, const word_t* w, uint64_t n, uint64_t* out, cudaStream_t s);
void lpp_launch_row_words(const ModelDev& m, uint64_t first, uint64_t count, word_t* up, word_t* dn, cudaStream_t s);
void lpp_launch_rank_pairs(const ModelDev& m, const word_t* up, const word_t* dn, uint64_t n, uint64_t* out, cudaStream_t s);
void lpp_launch_split_tables(const uint64_t* binom, int nbits, int lobits, uint32_t* rlo, uint32_t* rhi, cudaStream_t s);
void lpp_launch_lut(const word_t* b, uint64_t n, uint32_t* lut, cudaStream_t s);
void lpp_launch_hop_count(const ModelDev& m, int spin, uint64_t n, uint32_t* cnt, uint32_t* maxcnt, cudaStream_t s);
void lpp_launch_hop_fill(const ModelDev& m, int spin, HopTable t, cudaStream_t s);
void lpp_launch_spin_diag(const ModelDev& m, int spin, uint64_t n, double* dv, cudaStream_t s);

struct HeisBond {
	int p, q;
	double jpm_pq, jpm_qp, jzz;
};
void lpp_launch_spmv_heis(const ModelDev& m, const HeisBond* bonds, int nbonds, int has_field, const SpmvArgs& a, cudaStream_t s);
int lpp_spmv_generic_blocks(uint64_t nloc);
void lpp_launch_spmv_generic(const ModelDev& m, const SpmvArgs& a, cudaStream_t s);
int lpp_spmv_table_blocks(const ModelDev& m, uint64_t nloc);
void lpp_launch_spmv_table(const ModelDev& m, const HopTable& up, const HopTable& dn, const DiagTables& dt,
                           const SpmvArgs& a, cudaStream_t s);

// stored CRS
void lpp_launch_crs_count(const ModelDev& m, uint64_t row0, uint64_t nloc, int64_t* counts, int* overflow, cudaStream_t s);
void lpp_launch_crs_fill(const ModelDev& m, uint64_t row0, uint64_t nloc, const int64_t* rowptr, int64_t* colind,
                         double* values, cudaStream_t s);
void lpp_exclusive_scan(int64_t* data, uint64_t n, int64_t* total_dev, cudaStream_t s);  // in place; data has n+1 slots
int lpp_spmv_crs_blocks(uint64_t nloc);
void lpp_launch_spmv_crs(const int64_t* rowptr, const int64_t* colind, const double* values, const SpmvArgs& a,
                         cudaStream_t s);

// Lanczos vector kernels
int lpp_vec_blocks(uint64_t n);
void lpp_launch_fill_random(double* v, uint64_t row0, uint64_t n, uint64_t seed, cudaStream_t s);
void lpp_launch_dot(const double* a, const double* b, uint64_t n, double* partials, cudaStream_t s);
// x -= coef*y ; partials <- block sums of x_new^2
void lpp_launch_axpy_norm(double* x, const double* y, double coef, uint64_t n, double* partials, cudaStream_t s,
                          const double* coef_dev = nullptr);
void lpp_launch_axpy(double* z, const double* v, double coef, uint64_t n, cudaStream_t s);
void lpp_launch_scale(double* v, double coef, uint64_t n, cudaStream_t s);
// out[0] = sum of partials[0..n) in a fixed order
void lpp_launch_finalize_sum(const double* partials, int n, double* out, cudaStream_t s);

// operator application (Engine.h:416-458), gather form on the destination basis
void lpp_launch_apply_op(const ModelDev& src, const ModelDev& dst, int op, int site, int spin, int orb, double factor,
                         const double* srcv, double* z, uint64_t dst_row0, uint64_t dst_nloc, cudaStream_t s);

// two-layout exchange helpers (multi-GPU)
#define LPP_MAX_RANKS 16
struct ColSplit {
	uint64_t cs[LPP_MAX_RANKS + 1];   // column range of peer q: [cs[q], cs[q+1])
	int nranks, me;
};
struct PeerPtrs {
	double* p[LPP_MAX_RANKS];
};
void lpp_launch_pack_cols_p2p(const double* src, const PeerPtrs& ycols, uint64_t nrows, uint64_t n1, const ColSplit& c,
                              uint64_t d0loc, cudaStream_t s);
void lpp_launch_unpack_add_p2p(double* x, const PeerPtrs& xcols, uint64_t nrows, uint64_t n1, const ColSplit& c, uint64_t d0loc,
                               cudaStream_t s);
int lpp_unpack_axpy_norm_blocks(uint64_t nrows, uint64_t n1, int nranks);
void lpp_launch_unpack_axpy_norm_p2p(double* x, const double* y, double coef, const PeerPtrs& xcols, const PeerPtrs* ycols_or_null,
                                     uint64_t nrows, uint64_t n1, const ColSplit& c, uint64_t d0loc, double* partials, cudaStream_t s,
                                     const double* coef_dev = nullptr);
// device-resident Lanczos scalars: coefs = {alpha = 1/n_j, beta = -b_{j-1}/n_{j-1}, a_j/n_j, n_j, j}
#define LPP_LZ_ALPHA 0
#define LPP_LZ_BETA 1
#define LPP_LZ_AXPY 2
#define LPP_LZ_NORM 3
#define LPP_LZ_STEP 4
void lpp_launch_lz_init(double nj, double* coefs, cudaStream_t s);
// Pipelined two-layout recurrence (lanczos_loop, LPP_PIPELINE): the sweeps run with alpha = 1, beta = 0 into a third buffer z and
// every scalar enters in the fused unpack, U_{j+1} = C1 (z + column-shard pieces) - C2 y - C3 x with C1 = 1/n_j, C2 = a_j/n_j,
// C3 = b_j/n_{j-1}.  The norm of U_{j+1} is only needed by the NEXT unpack, so the next up sweep and pack start chunk by chunk
// behind the unpack instead of after its all-reduce.  coefs slots (the array has 16 doubles):
#define LPP_LZP_C1 8
#define LPP_LZP_C2 9
#define LPP_LZP_C3 10
#define LPP_LZP_STEPA 11      // next index of a[]
#define LPP_LZP_STEPB 12      // next index of b[]
void lpp_launch_lzp_init(double nj, double* coefs, cudaStream_t s);
void lpp_launch_lzp_after_dot(const double* dot_parts, double* coefs, double* a_out, cudaStream_t s);
void lpp_launch_lzp_after_norm(const double* b2, double* coefs, double* b_out, cudaStream_t s);
// rows [0, nrows) of the blocks x, y, z handed in (the caller offsets the pointers and d0loc for a chunk of rows):
// x = C1 (z + pieces of the peers' column-shard results) - C2 y - C3 x, partials[r * gy + chunk] = sum of squares
int lpp_unpack3_partials_per_row(uint64_t n1);
void lpp_launch_unpack3_norm_p2p(double* x, const double* y, const double* z, const double* coefs, const PeerPtrs& xcols, uint64_t nrows,
                                 uint64_t n1, const ColSplit& c, uint64_t d0loc, double* partials, int max_cta_rows, cudaStream_t s);
void lpp_launch_lz_after_dot(const double* dot_parts, int nparts, double* coefs, double* a_out, cudaStream_t s);
void lpp_launch_lz_after_norm(const double* b2, double* coefs, double* b_out, cudaStream_t s);
void lpp_launch_pack_cols(const double* src, double* sendbuf, double* ycol, uint64_t nrows, uint64_t n1, const ColSplit& c,
                          uint64_t d0loc, cudaStream_t s);
void lpp_launch_unpack_add(double* x, const double* recvbuf, const double* xcol, uint64_t nrows, uint64_t n1, const ColSplit& c,
                           uint64_t d0loc, cudaStream_t s);

// All-reduce (sum) of up to 4 doubles over peer memory (NVLink): every rank stores its values into a slot of every peer's exchange
// area and a sequence number behind a system-scope fence, waits for the G slots of its own area, and adds them in rank order
// (identical result on every rank).  Slots alternate with the parity of `seq`.  A rank that waits longer than ~2 s sets *err.
#define LPP_PSX_SLOT_DOUBLES 8          // 4 values, sequence number, padding: 64 bytes
#define LPP_PSX_DOUBLES (2 * LPP_MAX_RANKS * LPP_PSX_SLOT_DOUBLES)
void lpp_launch_psx_allreduce(double* vals, int nvals, const PeerPtrs& areas, int me, int nranks, unsigned long long seq, int* err,
                              cudaStream_t s);

// tiled two-sweep kernels (lpp_tiled.cu)
struct TiledPlan;
