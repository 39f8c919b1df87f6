// oracle/ref_bridge.cpp -- TEST INFRASTRUCTURE, never linked into or called by the product.
//
// Compiles the REFERENCE'S OWN model headers, unmodified and from where they lie under /root/reference/src
// (HubbardOneOrbital.h + HubbardHelper.h + BasisHubbardLanczos.h + BasisOneSpin.h, FeBasedSc.h + BasisFeAsBasedSc.h +
// BasisOneSpinFeAs.h + Partitions.h, Heisenberg.h + BasisHeisenberg.h, TjMultiOrb.h + BasisTjMultiOrbLanczos.h, ModelBase.h, BasisBase.h, ProgramGlobals.h,
// LabeledOperator.h, RahulOperator.h and the three Parameters*.h), against oracle/psimag_shim/ -- a minimal stand-in for the
// un-vendored PsimagLite containers those headers include (Vector, Matrix, CrsMatrix, SparseRow, BitManip::count,
// Parallelizer).  Output: oracle/_ref/liblpp_ref.so (git-ignored; built by `make -C oracle _ref` when /root/reference exists).
//
// What this pins: basis words and their order, perfectIndex, fermion signs, hop / exchange / pair-hop enumeration, diagonal
// terms, the order in which rows are assembled, getBraIndex / doSignGf of the continued-fraction path -- all executed by the
// reference's code.  What it does not pin: PsimagLite itself (SparseRow's sort-and-merge, CrsMatrix, LanczosSolver,
// ContinuedFraction), which is restated in the shim / in lanczos_oracle.c from SURVEY App. B.
//
// The geometry and the input reader are small local classes with the call signatures the reference uses
// (geometry(i, orb_i, j, orb_j, term), numberOfSites(), terms(); io.readline(x, "Label="), io.read(vector, "Label")).
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <type_traits>
#include <sstream>
#include <string>
#include <vector>

#include "Vector.h"
#include "Matrix.h"
#include "CrsMatrix.h"

#include "ProgramGlobals.h"
#include "HubbardOneOrbital.h"
#include "BasisFeAsBasedSc.h"
#include "FeBasedSc.h"
#include "Heisenberg.h"
#include "TjMultiOrb.h"

// static members the reference defines in LanczosDriver0.cpp:53-58 and ProgramGlobals.cpp:5
SizeType LanczosPlusPlus::BasisOneSpin::nsite_ = 0;
PsimagLite::Matrix<SizeType> LanczosPlusPlus::BasisOneSpin::comb_;
SizeType LanczosPlusPlus::BasisOneSpinFeAs::orbitals_ = 2;
SizeType LanczosPlusPlus::BasisOneSpinFeAs::nsite_ = 0;
PsimagLite::Matrix<SizeType> LanczosPlusPlus::BasisOneSpinFeAs::comb_;
PsimagLite::Vector<LanczosPlusPlus::ProgramGlobals::WordType>::Type LanczosPlusPlus::ProgramGlobals::bitmask_;

namespace {

class BridgeGeometry {
public:
	BridgeGeometry(SizeType nsite, SizeType orbitals, SizeType terms, const double* t0, const double* t1, const double* t2 = nullptr,
	               const double* t3 = nullptr)
	    : nsite_(nsite), orbitals_(orbitals), terms_(terms)
	{
		const SizeType nb = nsite * orbitals;
		const double* t[4] = {t0, t1, t2, t3};
		for (int q = 0; q < 4; q++) {
			term_[q].assign(nb * nb, 0.0);
			if (t[q]) std::memcpy(term_[q].data(), t[q], sizeof(double) * nb * nb);
		}
	}
	SizeType numberOfSites() const { return nsite_; }
	SizeType terms() const { return terms_; }
	double operator()(SizeType i, SizeType orb1, SizeType j, SizeType orb2, SizeType term) const
	{
		const SizeType nb = nsite_ * orbitals_;
		const SizeType a = i * orbitals_ + orb1, b = j * orbitals_ + orb2;
		if (term >= terms_) throw PsimagLite::RuntimeError("BridgeGeometry: term out of range\n");
		return term_[term][a * nb + b];
	}
private:
	SizeType nsite_, orbitals_, terms_;
	std::vector<double> term_[4];
};

class BridgeInput {
public:
	std::map<std::string, std::string> lines;
	std::map<std::string, std::vector<double> > vectors;
	template <typename T> void readline(T& x, const std::string& label)
	{
		auto it = lines.find(label);
		if (it == lines.end()) throw std::runtime_error("BridgeInput: no " + label);
		std::istringstream ss(it->second);
		ss >> x;
	}
	template <typename T> typename std::enable_if<std::is_arithmetic<T>::value, void>::type read(T& x, const std::string& label)
	{
		readline(x, label);                    // ParametersTjMultiOrb.h:99 reads the scalar Orbitals= with read()
	}
	template <typename T> void read(std::vector<T>& v, const std::string& label)
	{
		auto it = vectors.find(label);
		if (it == vectors.end()) throw std::runtime_error("BridgeInput: no " + label);
		v.assign(it->second.begin(), it->second.end());
	}
	template <typename T> void read(PsimagLite::Matrix<T>&, const std::string& label)
	{
		throw std::runtime_error("BridgeInput: no " + label);
	}
};

typedef LanczosPlusPlus::ModelBase<double, BridgeGeometry, BridgeInput> ModelBaseType;
typedef ModelBaseType::BasisBaseType BasisBaseType;
typedef LanczosPlusPlus::HubbardOneOrbital<double, BridgeGeometry, BridgeInput> HubbardType;
typedef LanczosPlusPlus::BasisFeAsBasedSc<BridgeGeometry> BasisFeAsType;
typedef LanczosPlusPlus::FeBasedSc<double, BasisFeAsType, BridgeInput> FeAsType;
typedef LanczosPlusPlus::Heisenberg<double, BridgeGeometry, BridgeInput> HeisenbergType;
typedef LanczosPlusPlus::TjMultiOrb<double, BridgeGeometry, BridgeInput> TjType;
typedef LanczosPlusPlus::LabeledOperator LabeledOperatorType;

struct RefModel {
	int kind = 0;
	std::unique_ptr<BridgeGeometry> geometry;
	std::unique_ptr<ModelBaseType> model;
	const BasisBaseType* basis = nullptr;          // the model's own basis, or a new-sector basis owned by `parent`
	RefModel* parent = nullptr;
	ModelBaseType::SparseMatrixType matrix;
	bool have_matrix = false;
};

thread_local std::string g_err;

struct CoutSilencer {                               // FeBasedSc.h:242-243 prints on every product
	std::streambuf* old;
	std::ostringstream sink;
	CoutSilencer() : old(std::cout.rdbuf(sink.rdbuf())) {}
	~CoutSilencer() { std::cout.rdbuf(old); }
};

LabeledOperatorType::Label label_of(int op)
{
	switch (op) {                                   // numbering of include/lpp_b200.h (LPP_OP_*) = LabeledOperator.h:10-17
	case 1: return LabeledOperatorType::Label::OPERATOR_C;
	case 2: return LabeledOperatorType::Label::OPERATOR_SZ;
	case 3: return LabeledOperatorType::Label::OPERATOR_CDAGGER;
	case 4: return LabeledOperatorType::Label::OPERATOR_N;
	case 5: return LabeledOperatorType::Label::OPERATOR_SPLUS;
	case 6: return LabeledOperatorType::Label::OPERATOR_SMINUS;
	default: return LabeledOperatorType::Label::OPERATOR_NIL;
	}
}

} // namespace

#define REF_TRY try {
#define REF_CATCH(rv) } catch (std::exception& e) { g_err = e.what(); return rv; }

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void ref_set_threads(int n) { PsimagLite::Concurrency::codeSectionParams.npthreads = n > 0 ? (SizeType)n : 1; }

// model: 0 HubbardOneOrbital, 1 FeBasedSc (FeAsMode=INT_PAPER33), 2 Heisenberg (HeisenbergTwiceS=1; nup = TargetSzPlusConst)
void* ref_create(int model, int nsite, int orbitals, int nup, int ndown, const double* hop, const double* jzz, const double* U,
                 int nU, const double* V, int nV, const double* D, int nD)
{
	REF_TRY
	CoutSilencer quiet;
	std::unique_ptr<RefModel> r(new RefModel());
	r->kind = model;
	BridgeInput io;
	if (U) io.vectors["hubbardU"].assign(U, U + nU);
	// the bases keep the site count in static members and refuse a different one (BasisOneSpin.h:28-29)
	LanczosPlusPlus::BasisOneSpin::nsite_ = 0;
	LanczosPlusPlus::BasisOneSpinFeAs::nsite_ = 0;
	if (model == 0) {
		r->geometry.reset(new BridgeGeometry(nsite, 1, 1, hop, nullptr));
		io.lines["Model="] = "HubbardOneBand";
		io.vectors["potentialV"].assign(2 * (size_t)nsite, 0.0);          // input0.inp:14-16 carries 2*nsite values
		for (int i = 0; i < nsite && V && i < nV; i++) io.vectors["potentialV"][i] = io.vectors["potentialV"][i + nsite] = V[i];
		HubbardType* m = new HubbardType(nup, ndown, io, *r->geometry);
		r->model.reset(m);
	} else if (model == 1) {
		r->geometry.reset(new BridgeGeometry(nsite, orbitals, 1, hop, nullptr));
		io.lines["Orbitals="] = std::to_string(orbitals);
		io.lines["FeAsMode="] = "INT_PAPER33";
		io.vectors["potentialV"].assign(2 * (size_t)orbitals * nsite, 0.0);
		for (int i = 0; V && i < nV && i < 2 * orbitals * nsite; i++) io.vectors["potentialV"][i] = V[i];
		if (D && nD > 0) { std::ostringstream ss; ss.precision(17); ss << D[0]; io.lines["AnisotropyD="] = ss.str(); }
		FeAsType* m = new FeAsType(nup, ndown, io, *r->geometry);
		r->model.reset(m);
	} else if (model == 2) {
		r->geometry.reset(new BridgeGeometry(nsite, 1, 2, hop, jzz));
		io.lines["HeisenbergTwiceS="] = "1";
		if (V && nV > 0) io.vectors["MagneticField"].assign(V, V + nV);
		if (D && nD > 0) io.vectors["AnisotropyD"].assign(D, D + nD);
		HeisenbergType* m = new HeisenbergType(nup, io, *r->geometry);
		r->model.reset(m);
	} else {
		throw std::runtime_error("ref_create: unknown model");
	}
	r->basis = &r->model->basis();
	return r.release();
	REF_CATCH(nullptr)
}

// Tj1Orbital: TjMultiOrb with Orbitals=1; geometry terms 0..3 = hopping, S+S- coupling, SzSz coupling, n n coupling
// (TjMultiOrb.h:68-79); potentialV holds 2*nsite values or none (ParametersTjMultiOrb.h:91-93)
void* ref_create_tj(int nsite, int nup, int ndown, const double* hop, const double* jpm, const double* jzz, const double* w,
                    const double* V, int nV)
{
	REF_TRY
	CoutSilencer quiet;
	std::unique_ptr<RefModel> r(new RefModel());
	r->kind = 3;
	BridgeInput io;
	r->geometry.reset(new BridgeGeometry(nsite, 1, 4, hop, jpm, jzz, w));
	io.lines["Orbitals="] = "1";
	if (V && nV > 0) io.vectors["potentialV"].assign(V, V + nV);
	r->model.reset(new TjType(nup, ndown, io, *r->geometry));
	r->basis = &r->model->basis();
	return r.release();
	REF_CATCH(nullptr)
}

// a new-sector basis of the same model (Engine.h:395-414 -> model.createBasis): owned by the parent model's garbage list
void* ref_new_sector(void* parent, int nup, int ndown)
{
	REF_TRY
	RefModel* p = static_cast<RefModel*>(parent);
	std::unique_ptr<RefModel> r(new RefModel());
	r->kind = p->kind;
	r->parent = p;
	r->basis = p->model->createBasis(nup, ndown);
	return r.release();
	REF_CATCH(nullptr)
}

void ref_destroy(void* h) { delete static_cast<RefModel*>(h); }

uint64_t ref_rows(void* h) { return static_cast<RefModel*>(h)->basis->size(); }

// out[i] = basis(i, spin) for every row i (BasisHubbardLanczos.h:77-84 and siblings)
int ref_basis_words(void* h, int spin, uint64_t* out)
{
	REF_TRY
	const BasisBaseType& b = *static_cast<RefModel*>(h)->basis;
	const SizeType n = b.size();
	for (SizeType i = 0; i < n; ++i) out[i] = b(i, spin);
	return 0;
	REF_CATCH(-1)
}

int64_t ref_perfect_index(void* h, uint64_t ket1, uint64_t ket2)
{
	REF_TRY
	return (int64_t) static_cast<RefModel*>(h)->basis->perfectIndex(ket1, ket2);
	REF_CATCH(-1)
}

// stored Hamiltonian through model.setupHamiltonian(matrix, basis); pass nulls to get nnz first
int64_t ref_crs(void* h, int64_t* rowptr, int64_t* colind, double* values)
{
	REF_TRY
	RefModel* r = static_cast<RefModel*>(h);
	if (!r->have_matrix) {
		CoutSilencer quiet;
		ModelBaseType* m = r->parent ? r->parent->model.get() : r->model.get();
		m->setupHamiltonian(r->matrix, *r->basis);
		r->have_matrix = true;
	}
	const SizeType n = r->matrix.rows(), nnz = r->matrix.getRowPtr(n);
	if (rowptr) for (SizeType i = 0; i <= n; ++i) rowptr[i] = (int64_t)r->matrix.getRowPtr(i);
	if (colind) for (SizeType k = 0; k < nnz; ++k) colind[k] = (int64_t)r->matrix.getCol(k);
	if (values) for (SizeType k = 0; k < nnz; ++k) values[k] = r->matrix.getValue(k);
	return (int64_t)nnz;
	REF_CATCH(-1)
}

// x += H y through model.matrixVectorProduct(x, y, basis): the on-the-fly path (HubbardHelper.h:105-134, FeBasedSc.h:228-245).
// Heisenberg has none in the reference (ModelBase.h:65-71 throws): returns -1.
int ref_matvec(void* h, double* x, const double* y)
{
	REF_TRY
	RefModel* r = static_cast<RefModel*>(h);
	CoutSilencer quiet;
	const SizeType n = r->basis->size();
	std::vector<double> xv(x, x + n), yv(y, y + n);
	ModelBaseType* m = r->parent ? r->parent->model.get() : r->model.get();
	m->matrixVectorProduct(xv, yv, *r->basis);
	std::memcpy(x, xv.data(), sizeof(double) * n);
	return 0;
	REF_CATCH(-1)
}

// z += factor * O(site, spin, orb) |src>: the loop of Engine::accModifiedState_ (Engine.h:416-458) restated around the
// reference's own newBasis.getBraIndex / srcBasis.doSignGf / doSignSpSm
int ref_apply_op(void* hsrc, void* hdst, int op, int site, int spin, int orb, double factor, const double* src, double* z)
{
	REF_TRY
	const BasisBaseType& sb = *static_cast<RefModel*>(hsrc)->basis;
	const BasisBaseType& nb = *static_cast<RefModel*>(hdst)->basis;
	const LabeledOperatorType lop(label_of(op));
	const SizeType nz = nb.size();
	for (SizeType ispace = 0; ispace < sb.size(); ++ispace) {
		const LanczosPlusPlus::ProgramGlobals::WordType ket1 = sb(ispace, 0), ket2 = sb(ispace, 1);
		const LanczosPlusPlus::ProgramGlobals::PairIntType t = nb.getBraIndex(ket1, ket2, lop, site, spin, orb);
		const int temp = t.first;
		const double value = t.second;
		if (temp >= 0 && (SizeType)temp >= nz) throw std::runtime_error("ref_apply_op: index outside the new basis");
		if (temp < 0) continue;
		double mysign = lop.isFermionic() ? sb.doSignGf(ket1, ket2, site, spin, orb) : 1;
		if (lop.id() == LabeledOperatorType::Label::OPERATOR_SPLUS || lop.id() == LabeledOperatorType::Label::OPERATOR_SMINUS)
			mysign *= sb.doSignSpSm(ket1, ket2, site, spin, orb);
		z[temp] += factor * mysign * value * src[ispace];
	}
	return 0;
	REF_CATCH(-1)
}

// psiNew = prod ops |psi> through the reference's own ModelBase::rahulMethod (ModelBase.h:89-141); labels 0 identity, 1 n, 2 sz,
// 3 c (RahulOperator.h:56-63), the operator list in the order Engine::measure builds it (Engine.h:217-234)
int ref_rahul(void* h, int nops, const int* labels, const int* dofs, const int* transposes, const int* sites, const double* psi,
              double* psiNew)
{
	REF_TRY
	RefModel* r = static_cast<RefModel*>(h);
	ModelBaseType* m = r->parent ? r->parent->model.get() : r->model.get();
	static const char* names[4] = {"identity", "n", "sz", "c"};
	ModelBaseType::VectorRahulOperatorType vops;
	ModelBaseType::VectorSizeType vsites(nops);
	for (int i = 0; i < nops; i++) {
		if (labels[i] < 0 || labels[i] > 3) throw std::runtime_error("ref_rahul: bad label");
		vops.push_back(ModelBaseType::RahulOperatorType(names[labels[i]], dofs[i], transposes[i] != 0));
		vsites[i] = sites[i];
	}
	const SizeType n = r->basis->size();
	std::vector<double> in(psi, psi + n), out(n, 0.0);
	m->rahulMethod(out, vops, vsites, in, *r->basis);
	std::memcpy(psiNew, out.data(), sizeof(double) * n);
	return 0;
	REF_CATCH(-1)
}

// hasNewParts (HubbardOneOrbital.h:88-107 and siblings): returns 1 and the new sector, 0 when the operator keeps the sector
int ref_has_new_parts(void* h, int op, int spin, int orb, int nup, int ndown, int* new_up, int* new_down)
{
	REF_TRY
	RefModel* r = static_cast<RefModel*>(h);
	std::pair<SizeType, SizeType> np(0, 0), old((SizeType)nup, (SizeType)ndown);
	const bool b = r->model->hasNewParts(np, old, LabeledOperatorType(label_of(op)), spin, orb);
	*new_up = (int)np.first;
	*new_down = (int)np.second;
	return b ? 1 : 0;
	REF_CATCH(-1)
}

} // extern "C"
