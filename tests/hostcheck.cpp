// tests/hostcheck.cpp -- TEST-ONLY: compiles the product's __host__ __device__ basis / rank / row-generator code
// (lanczosplusplus_b200/csrc/lpp_device.cuh) for the CPU so that pytest can compare it with the oracle without a GPU.
// It is never part of liblpp_b200.so and is not a CPU fallback: it exists to catch index/sign bugs before a GPU run.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../lanczosplusplus_b200/csrc/lpp_device.cuh"
#include "../lanczosplusplus_b200/csrc/lpp_setup.h"

struct HostModel {
	ModelDev m;
	std::vector<uint64_t> binom;
	std::vector<double> hop, jzz, U, V, D, jpm, w;
	std::vector<word_t> b1, b2;
	LppFeasLayout L1, L2;
	std::vector<uint32_t> rlo, rhi, lut1, lut2;
};

extern "C" {

HostModel* hc_create(int model, int nsite, int orbitals, int nup, int ndown, const double* hop, const double* jzz,
                     const double* U, int nU, const double* V, int nV, const double* D, int nD, int u3_all_pairs,
                     int use_tables)
{
	HostModel* h = new HostModel();
	const int no = model == LPP_MODEL_FEAS ? orbitals : 1, nb = nsite * no;
	h->binom = lpp_make_binom();
	h->hop.assign(hop, hop + (size_t)nb * nb);
	h->jzz.assign((size_t)nb * nb, 0.0);
	if (jzz) h->jzz.assign(jzz, jzz + (size_t)nb * nb);
	const int needU = model == LPP_MODEL_FEAS ? 6 : nsite;
	h->U.assign(needU, 0.0);
	for (int i = 0; i < needU && i < nU; i++) h->U[i] = U[i];
	if (model == LPP_MODEL_FEAS && (nU == 4 || nU == 5)) { h->U[4] = h->U[2]; h->U[5] = 0; }
	const int needV = model == LPP_MODEL_FEAS ? 2 * no * nsite : (model == LPP_MODEL_TJ ? 2 * nsite : nsite);
	h->jpm.assign((size_t)nb * nb, 0.0);
	h->w.assign((size_t)nb * nb, 0.0);
	h->V.assign(needV, 0.0);
	for (int i = 0; i < needV && i < nV; i++) h->V[i] = V[i];
	h->D.assign(nsite, 0.0);
	for (int i = 0; i < nsite && i < nD; i++) h->D[i] = D[i];
	ModelDev& m = h->m;
	memset(&m, 0, sizeof(m));
	m.model = model; m.nsite = nsite; m.orbitals = no; m.nbits = nb; m.nup = nup; m.ndn = ndown;
	m.u3_all_pairs = u3_all_pairs;
	m.binom = h->binom.data();
	m.hop = h->hop.data(); m.jzz = h->jzz.data(); m.U = h->U.data(); m.V = h->V.data(); m.D = h->D.data();
	m.jpm = h->jpm.data(); m.w = h->w.data();
	if (model == LPP_MODEL_FEAS) {
		h->L1 = lpp_feas_layout(h->binom, nsite, no, nup);
		h->L2 = lpp_feas_layout(h->binom, nsite, no, ndown);
		m.part_off1 = h->L1.off.data(); m.part_start1 = h->L1.start.data(); m.part_n1 = h->L1.pn.data();
		m.nparts1 = (int)h->L1.start.size() - 1;
		m.part_off2 = h->L2.off.data(); m.part_start2 = h->L2.start.data(); m.part_n2 = h->L2.pn.data();
		m.nparts2 = (int)h->L2.start.size() - 1;
		m.n1 = h->L1.total; m.n2 = h->L2.total;
		h->b1.resize(m.n1); h->b2.resize(m.n2);
		for (uint64_t i = 0; i < m.n1; i++) h->b1[i] = lpp_unrank_feas(m, 0, i);
		for (uint64_t i = 0; i < m.n2; i++) h->b2[i] = lpp_unrank_feas(m, 1, i);
	} else if (model == LPP_MODEL_TJ) {      // same set-up as create_impl in lpp_engine.cu
		m.n1 = h->binom[(nsite - ndown) * LPP_BINOM_N + nup];
		m.n2 = h->binom[nsite * LPP_BINOM_N + ndown];
		h->b1.resize(m.n1); h->b2.resize(m.n2);
		for (uint64_t i = 0; i < m.n1; i++) h->b1[i] = lpp_unrank_colex(m.binom, nsite - ndown, nup, i);
		for (uint64_t i = 0; i < m.n2; i++) h->b2[i] = lpp_unrank_colex(m.binom, nsite, ndown, i);
	} else {
		m.n1 = h->binom[nsite * LPP_BINOM_N + nup];
		m.n2 = model == LPP_MODEL_HUBBARD ? h->binom[nsite * LPP_BINOM_N + ndown] : 1;
		h->b1.resize(m.n1); h->b2.assign(m.n2, 0);
		for (uint64_t i = 0; i < m.n1; i++) h->b1[i] = lpp_unrank_colex(m.binom, nsite, nup, i);
		if (model == LPP_MODEL_HUBBARD)
			for (uint64_t i = 0; i < m.n2; i++) h->b2[i] = lpp_unrank_colex(m.binom, nsite, ndown, i);
	}
	m.b1 = h->b1.data(); m.b2 = h->b2.data();
	m.rows = m.n1 * m.n2;
	if (use_tables) {
		if (model != LPP_MODEL_FEAS) {
			int lobits = (nb + 1) / 2, hibits = nb - lobits;
			h->rlo.resize((size_t)1 << lobits);
			h->rhi.resize((size_t)(lobits + 1) << hibits);
			for (uint64_t i = 0; i < h->rlo.size(); i++) h->rlo[i] = lpp_split_lo_entry(m.binom, i);
			for (uint64_t i = 0; i < h->rhi.size(); i++) h->rhi[i] = lpp_split_hi_entry(m.binom, lobits, hibits, i);
			m.rlo = h->rlo.data(); m.rhi = h->rhi.data(); m.lobits = lobits;
		} else {
			h->lut1.assign((size_t)1 << nb, 0xffffffffu);
			h->lut2.assign((size_t)1 << nb, 0xffffffffu);
			for (uint64_t i = 0; i < m.n1; i++) h->lut1[h->b1[i]] = (uint32_t)i;
			for (uint64_t i = 0; i < m.n2; i++) h->lut2[h->b2[i]] = (uint32_t)i;
			m.lut1 = h->lut1.data(); m.lut2 = h->lut2.data();
		}
	}
	return h;
}

// t-J couplings of geometry terms 1 and 3 (TjMultiOrb.h:68-79)
void hc_set_tj(HostModel* h, const double* jpm, const double* w)
{
	const size_t nn = (size_t)h->m.nbits * h->m.nbits;
	if (jpm) h->jpm.assign(jpm, jpm + nn);
	if (w) h->w.assign(w, w + nn);
	h->m.jpm = h->jpm.data();
	h->m.w = h->w.data();
}
void hc_row_words(const HostModel* h, int spin, uint64_t* out)
{
	for (uint64_t r = 0; r < h->m.rows; r++) {
		const LppRowKets k = lpp_row_kets(h->m, r);
		out[r] = spin ? k.k2 : k.k1;
	}
}
uint64_t hc_rank_pair(const HostModel* h, uint64_t k1, uint64_t k2) { return lpp_rank_pair(h->m, k1, k2); }

// Engine::measure pieces on the host: psiNew = prod ops |psi> (scatter, as ModelBase::rahulMethod) with the product's lpp_rahul_apply
void hc_rahul(const HostModel* h, int nops, const int* labels, const int* dofs, const int* transposes, const int* sites, const double* psi,
              double* psiNew)
{
	LppMeasureOps ops;
	memset(&ops, 0, sizeof(ops));
	ops.n = nops;
	for (int i = 0; i < nops; i++) { ops.label[i] = labels[i]; ops.dof[i] = dofs[i]; ops.transpose[i] = transposes[i]; ops.site[i] = sites[i]; }
	for (uint64_t r = 0; r < h->m.rows; r++) {
		const LppRowKets k = lpp_row_kets(h->m, r);
		word_t o1, o2;
		double v;
		if (!lpp_rahul_apply(ops, k.k1, k.k2, &o1, &o2, &v)) continue;
		if (h->m.model == LPP_MODEL_TJ && (o1 & o2)) continue;
		psiNew[lpp_rank_pair(h->m, o1, o2)] += v * psi[r];
	}
}

void hc_destroy(HostModel* h) { delete h; }
uint64_t hc_rows(const HostModel* h) { return h->m.rows; }
uint64_t hc_basis_size(const HostModel* h, int spin) { return spin ? h->m.n2 : h->m.n1; }
void hc_basis(const HostModel* h, int spin, uint64_t* out)
{
	const auto& b = spin ? h->b2 : h->b1;
	memcpy(out, b.data(), sizeof(uint64_t) * b.size());
}
uint64_t hc_rank(const HostModel* h, int spin, uint64_t w) { return lpp_rank_onespin(h->m, spin, w); }

// stored CRS; pass rowptr == NULL to count. returns nnz or -1 on row overflow
int64_t hc_crs(const HostModel* h, int64_t* rowptr, int64_t* colind, double* values)
{
	int64_t nnz = 0;
	std::vector<uint64_t> c(LPP_ROW_CAP);
	std::vector<double> v(LPP_ROW_CAP);
	for (uint64_t r = 0; r < h->m.rows; r++) {
		int n = lpp_stored_row(h->m, r, c.data(), v.data());
		if (n < 0) return -1;
		if (rowptr) rowptr[r] = nnz;
		if (colind)
			for (int i = 0; i < n; i++) { colind[nnz + i] = (int64_t)c[i]; values[nnz + i] = v[i]; }
		nnz += n;
	}
	if (rowptr) rowptr[h->m.rows] = nnz;
	return nnz;
}

struct HcAcc {
	const double* y;
	double acc;
	void operator()(uint64_t c, double v) { acc += v * y[c]; }
};

// x += H y with the on-the-fly row generator
void hc_matvec(const HostModel* h, double* x, const double* y)
{
	for (uint64_t r = 0; r < h->m.rows; r++) {
		LppRowKets k = lpp_row_kets(h->m, r);
		HcAcc e{y, lpp_row_diag(h->m, k) * y[r]};
		lpp_row_offdiag(h->m, k, 0, e);
		x[r] += e.acc;
	}
}

// z (dst basis) += factor * O |srcv> using the gather form
void hc_apply_op(const HostModel* src, const HostModel* dst, int op, int site, int spin, int orb, double factor, const double* srcv,
                 double* z)
{
	for (uint64_t r = 0; r < dst->m.rows; r++) {
		uint64_t srow;
		double sg;
		if (!lpp_apply_op_source(src->m, dst->m, op, site, spin, orb, r, &srow, &sg)) continue;
		z[r] += factor * sg * srcv[srow];
	}
}

double hc_splitmix(uint64_t seed, uint64_t idx) { return lpp_splitmix_uniform(seed, idx); }
}
