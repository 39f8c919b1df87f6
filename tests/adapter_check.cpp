// tests/adapter_check.cpp -- TEST-ONLY.  Compiles include/InternalProductCuda.h at the template slot it is written for:
// Engine's `InternalProductTemplate<ModelType, SpecialSymmetryType>` (Engine.h:37-52), next to the reference's own
// InternalProductOnTheFly and InternalProductStored, with the reference's ModelBase / DefaultSymmetry / model headers taken
// unmodified from /root/reference/src and PsimagLite replaced by oracle/psimag_shim.  It then runs x += H y through all
// three on the same vectors (the CUDA one needs a B200) and prints the largest differences.
//
// The reference's ModelBase has no accessor for the (private) model parameters; INTEGRATION.md adds two virtual functions
// to it.  Here the same arrangement is a wrapper around the reference's model object (ModelWithCuda below), so that no
// reference source is modified or copied.  Built by `make -C oracle _ref` into oracle/_ref/adapter_check.
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
