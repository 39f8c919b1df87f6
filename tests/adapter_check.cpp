// tests/adapter_check.cpp -- TEST-ONLY.  Compiles include/InternalProductCuda.h at the template slot it is written for:
// Engine's `InternalProductTemplate<ModelType, SpecialSymmetryType>` (Engine.h:37-52), next to the reference's own
// InternalProductOnTheFly and InternalProductStored, with the reference's ModelBase / DefaultSymmetry / model headers taken
// unmodified from /root/reference/src and PsimagLite replaced by oracle/psimag_shim.  It then runs x += H y through all
// three on the same vectors (the CUDA one needs a B200) and prints the largest differences.
//
// The reference's ModelBase has no accessor for the (private) model parameters; INTEGRATION.md adds two virtual functions
// to it.  Here the same arrangement is a wrapper around the reference's model object (ModelWithCuda below), so that no
// reference source is modified or copied.  Built by `make -C oracle _ref` into oracle/_ref/adapter_check.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <type_traits>
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
#include "DefaultSymmetry.h"
#include "InternalProductOnTheFly.h"
#include "InternalProductStored.h"
#include "../include/InternalProductCuda.h"

SizeType LanczosPlusPlus::BasisOneSpin::nsite_ = 0;
PsimagLite::Matrix<SizeType> LanczosPlusPlus::BasisOneSpin::comb_;
SizeType LanczosPlusPlus::BasisOneSpinFeAs::orbitals_ = 2;
SizeType LanczosPlusPlus::BasisOneSpinFeAs::nsite_ = 0;
PsimagLite::Matrix<SizeType> LanczosPlusPlus::BasisOneSpinFeAs::comb_;
PsimagLite::Vector<LanczosPlusPlus::ProgramGlobals::WordType>::Type LanczosPlusPlus::ProgramGlobals::bitmask_;

namespace {

// geometry(i, orb_i, j, orb_j, term): nearest-neighbour chain, one value per term (orbital-diagonal)
class ChainGeometry {
public:
	typedef double ComplexOrRealType;
	ChainGeometry(SizeType nsite, bool periodic, const std::vector<double>& termValues) : nsite_(nsite), periodic_(periodic), v_(termValues) {}
	SizeType numberOfSites() const { return nsite_; }
	SizeType terms() const { return v_.size(); }
	double operator()(SizeType i, SizeType orb1, SizeType j, SizeType orb2, SizeType term) const
	{
		if (orb1 != orb2) return 0.0;
		const SizeType d = i > j ? i - j : j - i;
		const bool bond = d == 1 || (periodic_ && nsite_ > 2 && d == nsite_ - 1);
		return bond ? v_[term] : 0.0;
	}
private:
	SizeType nsite_;
	bool periodic_;
	std::vector<double> v_;
};

class MapInput {
public:
	std::map<std::string, std::string> lines;
	std::map<std::string, std::vector<double> > vectors;
	template <typename T> void readline(T& x, const std::string& label)
	{
		auto it = lines.find(label);
		if (it == lines.end()) throw std::runtime_error("no " + label);
		std::istringstream ss(it->second);
		ss >> x;
	}
	template <typename T> typename std::enable_if<std::is_arithmetic<T>::value, void>::type read(T& x, const std::string& label) { readline(x, label); }
	template <typename T> void read(std::vector<T>& v, const std::string& label)
	{
		auto it = vectors.find(label);
		if (it == vectors.end()) throw std::runtime_error("no " + label);
		v.assign(it->second.begin(), it->second.end());
	}
	template <typename T> void read(PsimagLite::Matrix<T>&, const std::string& label) { throw std::runtime_error("no " + label); }
};

typedef LanczosPlusPlus::ModelBase<double, ChainGeometry, MapInput> ModelBaseType;

// "ModelBase with the two virtual functions of INTEGRATION.md": forwards everything the InternalProduct classes use
class ModelWithCuda {
public:
	typedef ModelBaseType::BasisBaseType BasisBaseType;
	typedef ModelBaseType::RealType RealType;
	typedef ModelBaseType::GeometryType GeometryType;
	typedef ModelBaseType::SparseMatrixType SparseMatrixType;
	typedef ModelBaseType::VectorType VectorType;
	ModelWithCuda(const ModelBaseType& m, int id, const MapInput& io) : m_(m), id_(id)
	{
		auto get = [&io](const char* k) { auto it = io.vectors.find(k); return it == io.vectors.end() ? std::vector<double>() : it->second; };
		U_ = get("hubbardU");
		V_ = get(id == LPP_MODEL_HEISENBERG ? "MagneticField" : "potentialV");
		if (id == LPP_MODEL_HEISENBERG) D_ = get("AnisotropyD");
		if (id == LPP_MODEL_FEAS) D_.assign(1, 0.0);
	}
	const BasisBaseType& basis() const { return m_.basis(); }
	const GeometryType& geometry() const { return m_.geometry(); }
	SizeType orbitals(SizeType i) const { return m_.orbitals(i); }
	void setupHamiltonian(SparseMatrixType& matrix, const BasisBaseType& b) const { m_.setupHamiltonian(matrix, b); }
	void matrixVectorProduct(VectorType& x, const VectorType& y, const BasisBaseType& b) const { m_.matrixVectorProduct(x, y, b); }
	void printOperators(std::ostream& os) const { m_.printOperators(os); }
	int cudaModelId() const { return id_; }
	void exportForCuda(lpp_desc& d) const
	{
		d.U = U_.empty() ? 0 : &U_[0]; d.nU = U_.size();
		d.V = V_.empty() ? 0 : &V_[0]; d.nV = V_.size();
		d.D = D_.empty() ? 0 : &D_[0]; d.nD = D_.size();
	}
private:
	const ModelBaseType& m_;
	int id_;
	std::vector<double> U_, V_, D_;
};

typedef LanczosPlusPlus::DefaultSymmetry<ChainGeometry, ModelWithCuda::BasisBaseType> SymmetryType;
typedef LanczosPlusPlus::InternalProductOnTheFly<ModelWithCuda, SymmetryType> OnTheFlyType;
typedef LanczosPlusPlus::InternalProductStored<ModelWithCuda, SymmetryType> StoredType;
typedef LanczosPlusPlus::InternalProductCuda<ModelWithCuda, SymmetryType> CudaType;

double splitmix(uint64_t seed, uint64_t idx)
{
	uint64_t z = idx * 0x9E3779B97F4A7C15ull + seed * 0xD1B54A32D192ED03ull + 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z = z ^ (z >> 31);
	return (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

double maxdiff(const std::vector<double>& a, const std::vector<double>& b)
{
	double d = 0;
	for (size_t i = 0; i < a.size(); i++) d = std::max(d, std::abs(a[i] - b[i]));
	return d;
}

// ---- what PsimagLite::LanczosSolver does at Engine.h:474-478 / :626, restated on the host for any product
struct Tridiagonal {
	std::vector<double> a_, b_;
	void resize(SizeType n) { a_.assign(n, 0.0); b_.assign(n, 0.0); }
	SizeType size() const { return a_.size(); }
	double& a(SizeType i) { return a_[i]; }
	double& b(SizeType i) { return b_[i]; }
};

template <class ProductType>
void hostDecomposition(const ProductType& m, const std::vector<double>& init, Tridiagonal& ab, SizeType steps)
{
	const SizeType n = m.rows();
	steps = std::min(steps, n);
	std::vector<double> x(n, 0.0), y(init);
	double nrm = 0;
	for (SizeType i = 0; i < n; i++) nrm += y[i]*y[i];
	nrm = std::sqrt(nrm);
	for (SizeType i = 0; i < n; i++) y[i] /= nrm;
	ab.resize(steps);
	for (SizeType j = 0; j < steps; j++) {
		m.matrixVectorProduct(x, y);
		double a = 0, b = 0;
		for (SizeType i = 0; i < n; i++) a += y[i]*x[i];
		for (SizeType i = 0; i < n; i++) { x[i] -= a*y[i]; b += x[i]*x[i]; }
		b = std::sqrt(b);
		ab.a(j) = a;
		ab.b(j) = b;
		for (SizeType i = 0; i < n; i++) { const double t = y[i]; y[i] = (b < 1e-10) ? x[i] : x[i]/b; x[i] = -b*t; }
	}
}

// lowest eigenvalue of the tridiagonal matrix by bisection on the Sturm count
double lowestEigenvalue(const Tridiagonal& ab)
{
	const SizeType n = ab.size();
	double lo = 1e300, hi = -1e300;
	for (SizeType i = 0; i < n; i++) {
		const double r = (i ? std::abs(ab.b_[i - 1]) : 0.0) + (i + 1 < n ? std::abs(ab.b_[i]) : 0.0);
		lo = std::min(lo, ab.a_[i] - r);
		hi = std::max(hi, ab.a_[i] + r);
	}
	for (int it = 0; it < 200; it++) {
		const double mid = 0.5*(lo + hi);
		int below = 0;
		double q = 1.0;
		for (SizeType i = 0; i < n; i++) {
			const double b2 = i ? ab.b_[i - 1]*ab.b_[i - 1] : 0.0;
			q = ab.a_[i] - mid - (i ? b2/q : 0.0);
			if (q == 0.0) q = 1e-300;
			if (q < 0) below++;
		}
		if (below >= 1) hi = mid; else lo = mid;
	}
	return 0.5*(lo + hi);
}

// The two call sites of integration/engine_cuda.patch, in the form the patch gives them: tag dispatch on KrylovPlacement
template <class ProductType>
void engineDecomposition(const ProductType& m, const std::vector<double>& init, Tridiagonal& ab, SizeType steps, LanczosPlusPlus::KrylovOnHostTag)
{
	hostDecomposition(m, init, ab, steps);
}
template <class ProductType>
void engineDecomposition(const ProductType& m, const std::vector<double>& init, Tridiagonal& ab, SizeType steps, LanczosPlusPlus::KrylovOnDeviceTag)
{
	m.decomposition(init, ab, steps, 0.0, 4);
}

// decomposition() and groundState() of the adapter against the host recurrence through the reference's own product
template <class HostProductType>
int checkKrylov(const char* name, const HostProductType& host, const CudaType& cuda, const std::vector<double>& init)
{
	const SizeType n = host.rows();
	Tridiagonal abh, abc;
	const SizeType steps = std::min<SizeType>(n, 24);
	engineDecomposition(host, init, abh, steps, typename LanczosPlusPlus::KrylovPlacement<HostProductType>::Tag());
	engineDecomposition(cuda, init, abc, steps, typename LanczosPlusPlus::KrylovPlacement<CudaType>::Tag());
	double dab = 0;
	const SizeType ncmp = std::min<SizeType>(std::min(abh.size(), abc.size()), 12);
	for (SizeType i = 0; i < ncmp; i++) {
		dab = std::max(dab, std::abs(abh.a_[i] - abc.a_[i])/std::max(1.0, std::abs(abh.a_[i])));
		if (i + 1 < ncmp) dab = std::max(dab, std::abs(abh.b_[i] - abc.b_[i])/std::max(1.0, std::abs(abh.b_[i])));
	}
	Tridiagonal full;
	hostDecomposition(host, init, full, std::min<SizeType>(n, 200));
	const double eh = lowestEigenvalue(full);
	double ec = 0;
	std::vector<double> z;
	cuda.groundState(ec, z, init, 200, 1e-12, 4);
	std::vector<double> hz(n, 0.0);
	host.matrixVectorProduct(hz, z);
	double res = 0, zz = 0;
	for (SizeType i = 0; i < n; i++) { res += (hz[i] - ec*z[i])*(hz[i] - ec*z[i]); zz += z[i]*z[i]; }
	res = std::sqrt(res);
	// the form Engine::computeAllStatesBelow calls after integration/engine_cuda.patch
	std::vector<double> eigs;
	std::vector<std::vector<double> > zs;
	cuda.statesBelow(eigs, zs, init, 1, 200, 1e-12, 4);
	const bool sb = eigs.size() == 1 && zs.size() == 1 && zs[0].size() == n && std::abs(eigs[0] - ec) <= 1e-12*std::max(1.0, std::abs(ec)) && maxdiff(zs[0], z) <= 1e-12;
	const bool ok = sb && abc.size() == abh.size() && dab <= 1e-10 && std::abs(eh - ec) <= 1e-9*std::max(1.0, std::abs(eh)) && res <= 1e-5 && std::abs(zz - 1.0) <= 1e-9;
	std::printf("%s krylov steps=%zu ab_rel_diff=%.3e energy_host=%.12f energy_cuda=%.12f residual=%.3e %s\n", name, (size_t)abc.size(), dab, eh, ec, res,
	            ok ? "ok" : "MISMATCH");
	return ok ? 0 : 1;
}

int check(const char* name, const ModelBaseType& model, int id, const MapInput& io, bool hasOnTheFly, bool run_cuda)
{
	ModelWithCuda m(model, id, io);
	SymmetryType rs(m.basis(), m.geometry(), "");
	StoredType stored(m, rs);
	const SizeType n = stored.rows();
	std::vector<double> y(n), x0(n), xs, xo, xc;
	for (SizeType i = 0; i < n; i++) { y[i] = splitmix(42, i); x0[i] = splitmix(7, i); }
	xs = x0;
	stored.matrixVectorProduct(xs, y);
	double d_otf = -1, d_cuda = -1;
	if (hasOnTheFly) {
		std::streambuf* old = std::cout.rdbuf();
		std::ostringstream sink;
		std::cout.rdbuf(sink.rdbuf());                       // FeBasedSc.h:242-243 prints on every product
		OnTheFlyType otf(m, rs);
		xo = x0;
		otf.matrixVectorProduct(xo, y);
		std::cout.rdbuf(old);
		d_otf = maxdiff(xo, xs);
	}
	if (run_cuda) {
		CudaType cuda(m, rs);
		if (cuda.rows() != n) { std::printf("FAIL %s rows %zu vs %zu\n", name, (size_t)cuda.rows(), (size_t)n); return 1; }
		xc = x0;
		cuda.matrixVectorProduct(xc, y);
		d_cuda = maxdiff(xc, xs);
		std::vector<double> init(n);
		for (SizeType i = 0; i < n; i++) init[i] = splitmix(1234, i);
		if (checkKrylov(name, stored, cuda, init)) return 1;
	}
	std::printf("%s rows=%zu otf_vs_stored=%.3e cuda_vs_stored=%.3e\n", name, (size_t)n, d_otf, d_cuda);
	return (run_cuda && !(d_cuda <= 1e-12)) ? 1 : 0;
}

} // namespace

int main(int argc, char** argv)
{
	const bool run_cuda = !(argc > 1 && !std::strcmp(argv[1], "--no-cuda"));
	int bad = 0;
	try {
		{   // HubbardOneBand, 6-site open chain, 3 up 3 down, U = 4, site potentials
			LanczosPlusPlus::BasisOneSpin::nsite_ = 0;
			ChainGeometry g(6, false, {-1.0});
			MapInput io;
			io.lines["Model="] = "HubbardOneBand";
			io.vectors["hubbardU"] = std::vector<double>(6, 4.0);
			io.vectors["potentialV"] = {0.3, -0.2, 0.1, 0.0, 0.5, -0.4, 0.3, -0.2, 0.1, 0.0, 0.5, -0.4};
			LanczosPlusPlus::HubbardOneOrbital<double, ChainGeometry, MapInput> model(3, 3, io, g);
			bad += check("HubbardOneBand", model, LPP_MODEL_HUBBARD, io, true, run_cuda);
		}
		{   // FeAsBasedSc INT_PAPER33, TestSuite/inputs/input100.inp couplings on 3 sites.  U[3] = 0 here: with U[3] != 0 the
			// reference's own on-the-fly and stored products differ and there is no single answer to compare with
			LanczosPlusPlus::BasisOneSpinFeAs::nsite_ = 0;
			ChainGeometry g(3, false, {-1.0});
			MapInput io;
			io.lines["Orbitals="] = "2";
			io.lines["FeAsMode="] = "INT_PAPER33";
			io.vectors["hubbardU"] = {4.0, 3.0, -0.8, 0.0};
			io.vectors["potentialV"] = std::vector<double>(12, 0.0);
			std::streambuf* old = std::cout.rdbuf();
			std::ostringstream sink;
			std::cout.rdbuf(sink.rdbuf());
			LanczosPlusPlus::FeBasedSc<double, LanczosPlusPlus::BasisFeAsBasedSc<ChainGeometry>, MapInput> model(2, 1, io, g);
			std::cout.rdbuf(old);
			bad += check("FeAsBasedSc", model, LPP_MODEL_FEAS, io, true, run_cuda);
		}
		{   // Heisenberg S=1/2, 10-site ring, Sz = 0 sector + 1
			ChainGeometry g(10, true, {1.0, 0.7});
			MapInput io;
			io.lines["HeisenbergTwiceS="] = "1";
			LanczosPlusPlus::Heisenberg<double, ChainGeometry, MapInput> model(4, io, g);
			bad += check("Heisenberg", model, LPP_MODEL_HEISENBERG, io, false, run_cuda);
		}
		{   // Tj1Orbital, 7-site chain, 3 up 2 down: t = -1, J = 0.4 (jpm = jzz = J, w = -J/4)
			ChainGeometry g(7, false, {-1.0, 0.4, 0.4, -0.1});
			MapInput io;
			io.lines["Orbitals="] = "1";
			LanczosPlusPlus::TjMultiOrb<double, ChainGeometry, MapInput> model(3, 2, io, g);
			bad += check("Tj1Orbital", model, LPP_MODEL_TJ, io, false, run_cuda);
		}
	} catch (std::exception& e) {
		std::printf("FAIL exception: %s\n", e.what());
		return 2;
	}
	std::printf(bad ? "FAIL\n" : "OK\n");
	return bad ? 1 : 0;
}
