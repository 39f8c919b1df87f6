/*
 * InternalProductCuda.h -- the drop-in a LanczosPlusPlus maintainer adds next to
 * src/Engine/InternalProductOnTheFly.h and src/Engine/InternalProductStored.h.
 *
 * It satisfies the template-template slot `InternalProductTemplate` of Engine (Engine.h:37-52,
 * LanczosDriver1.h:47-56): same typedefs, same two constructors, rows(), matrixVectorProduct(x,y) with
 * x += H y semantics (InternalProductOnTheFly.h:76,115-123), specialSymmetrySector(), reflectionSector(),
 * fullDiag().  All arithmetic happens in liblpp_b200.so (include/lpp_b200.h); this header only marshals the
 * model description and converts status codes into the reference's err() exceptions.
 *
 * Compiles inside the reference tree (needs PsimagLite for SizeType, err(), Vector.h, Matrix.h).  In this repository it is
 * compiled against the reference's own Engine/Model headers and the PsimagLite stand-in of oracle/psimag_shim by
 * tests/adapter_check.cpp (built into oracle/_ref/adapter_check), which runs it side by side with the reference's
 * InternalProductOnTheFly and InternalProductStored on the same models.
 *
 * Requirements on ModelType beyond the reference's ModelBase: two virtual functions added to Engine/ModelBase.h and
 * overridden by the four models on the path:
 *   virtual int cudaModelId() const;                  // LPP_MODEL_HUBBARD | LPP_MODEL_FEAS | LPP_MODEL_HEISENBERG | LPP_MODEL_TJ
 *   virtual void exportForCuda(lpp_desc& d) const;    // U/nU, V/nV, D/nD from hubbardU, potentialV, anisotropy (mp_ is private)
 */
#ifndef INTERNALPRODUCT_CUDA_H
#define INTERNALPRODUCT_CUDA_H

#include <vector>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>
#include "Vector.h"
#include "Matrix.h"
#include "lpp_b200.h"

namespace LanczosPlusPlus {

// Which GPU this process drives and, for a row-sharded run (one process per GPU, all started with the same input file), its
// rank.  Filled by the driver from the input label `Gpus=` (integration/engine_cuda.patch, LanczosDriver1.h) or, when the
// label is absent, from the launcher's environment: LPP_DEVICE | LOCAL_RANK, LPP_RANK | RANK, LPP_NRANKS | WORLD_SIZE,
// LPP_NCCL_ID_FILE (a path every rank can read: rank 0 writes the 128-byte NCCL id there).
struct CudaTopology {
	int device, rank, nranks;
	std::string idFile;
	CudaTopology() : device(0), rank(0), nranks(1) {}
	static CudaTopology& global()
	{
		static CudaTopology t = fromEnvironment();
		return t;
	}
	// `Gpus=` value: a comma-separated list of device ordinals, one per rank; this process takes entry `rank`
	static void setFromGpusLabel(const std::string& gpus)
	{
		CudaTopology& t = global();
		std::vector<int> devs;
		size_t pos = 0;
		while (pos < gpus.size()) {
			size_t q = gpus.find(',', pos);
			if (q == std::string::npos) q = gpus.size();
			if (q > pos) devs.push_back(std::atoi(gpus.substr(pos, q - pos).c_str()));
			pos = q + 1;
		}
		if (devs.empty()) return;
		t.nranks = devs.size();
		if (t.rank >= t.nranks) err("InternalProductCuda: rank >= number of entries of Gpus=\n");
		t.device = devs[t.rank];
	}
private:
	static int envInt(const char* a, const char* b, int dflt)
	{
		const char* v = std::getenv(a);
		if (!v) v = std::getenv(b);
		return v ? std::atoi(v) : dflt;
	}
	static CudaTopology fromEnvironment()
	{
		CudaTopology t;
		t.device = envInt("LPP_DEVICE", "LOCAL_RANK", 0);
		t.rank = envInt("LPP_RANK", "RANK", 0);
		t.nranks = envInt("LPP_NRANKS", "WORLD_SIZE", 1);
		const char* f = std::getenv("LPP_NCCL_ID_FILE");
		if (f) t.idFile = f;
		return t;
	}
};

// Tag for Engine's two solver call sites (integration/engine_cuda.patch): products that keep the Krylov loop on the GPU
struct KrylovOnHostTag {};
struct KrylovOnDeviceTag {};
template<typename InternalProductType>
struct KrylovPlacement {
	typedef KrylovOnHostTag Tag;
};

template<typename ModelType_, typename SpecialSymmetryType_>
class InternalProductCuda {

public:

	typedef ModelType_ ModelType;
	typedef SpecialSymmetryType_ SpecialSymmetryType;
	typedef typename ModelType::BasisBaseType BasisType;
	typedef typename SpecialSymmetryType::SparseMatrixType SparseMatrixType;
	typedef typename ModelType::RealType RealType;
	typedef typename ModelType::GeometryType GeometryType;
	typedef typename GeometryType::ComplexOrRealType ComplexOrRealType;
	typedef PsimagLite::Matrix<ComplexOrRealType> MatrixType;
	typedef typename PsimagLite::Vector<RealType>::Type VectorRealType;
	typedef typename PsimagLite::Vector<ComplexOrRealType>::Type VectorType;

	// Engine::spectralFunction path: a new-sector basis (Engine.h:186-187)
	InternalProductCuda(const ModelType& model,
	                    const BasisType& basis,
	                    SpecialSymmetryType&)
	    : model_(model), basis_(basis), handle_(0)
	{
		init();
	}

	// Engine::computeAllStatesBelow path (Engine.h:608)
	InternalProductCuda(const ModelType& model,
	                    SpecialSymmetryType&)
	    : model_(model), basis_(model.basis()), handle_(0)
	{
		init();
	}

	~InternalProductCuda()
	{
		if (handle_) lpp_destroy(handle_);
	}

	SizeType rows() const
	{
		return basis_.size();
	}

	// x += H y.  Host vectors cross PCIe on every call: kept for drop-in completeness and parity tests.
	// The production path is decomposition()/groundState() below, which keep the Krylov loop on the GPU.
	void matrixVectorProduct(VectorType& x, const VectorType& y) const
	{
		assert(x.size() == rows() && y.size() == rows());
		check(lpp_matvec_host(handle_, LPP_KERNEL_AUTO, &(x[0]), &(y[0])));
	}

	SizeType reflectionSector() const { return 0; }

	void specialSymmetrySector(SizeType) { }

	void fullDiag(VectorRealType&,
	              MatrixType&)
	{
		err("no fullDiag possible when on the GPU\n");
	}

	// Device-resident replacement of LanczosSolver::decomposition(init, ab) (Engine.h:474-478).
	template<typename TridiagonalMatrixType>
	void decomposition(const VectorType& init,
	                   TridiagonalMatrixType& ab,
	                   SizeType steps,
	                   RealType eps,
	                   SizeType minSteps) const
	{
		lpp_solver_params p;
		p.steps = steps; p.minsteps = minSteps; p.eps = eps; p.kernel = LPP_KERNEL_AUTO; p.reortho = 0; p.seed = 0;
		typename PsimagLite::Vector<RealType>::Type a(steps + 1), b(steps + 1);
		int32_t n = 0;
		double nrm2 = 0;
		check(lpp_lanczos_decomposition(handle_, &p, &(init[0]), 0, &(a[0]), &(b[0]), &n, &nrm2));
		ab.resize(n);
		for (SizeType i = 0; i < SizeType(n); ++i) {
			ab.a(i) = a[i];
			ab.b(i) = b[i];
		}
	}

	// Device-resident replacement of LanczosSolver::computeOneState (Engine.h:626 with excited = 0).
	void groundState(RealType& energy,
	                 VectorType& z,
	                 const VectorType& init,
	                 SizeType steps,
	                 RealType eps,
	                 SizeType minSteps) const
	{
		lpp_solver_params p;
		p.steps = steps; p.minsteps = minSteps; p.eps = eps; p.kernel = LPP_KERNEL_AUTO; p.reortho = 0; p.seed = 0;
		z.resize(rows());
		int32_t n = 0;
		check(lpp_ground_state(handle_, &p, &(init[0]), 1, &energy, &(z[0]), 0, 0, &n));
	}

	// Device-resident replacement of lanczosSolver.computeAllStatesBelow(eigs, zs, initial, excitedPlusOne) (Engine.h:626):
	// the form integration/engine_cuda.patch calls.  zs[k] is the k-th Ritz vector (length rows()).
	template<typename VectorVectorType>
	void statesBelow(VectorRealType& eigs,
	                 VectorVectorType& zs,
	                 const VectorType& init,
	                 SizeType excitedPlusOne,
	                 SizeType steps,
	                 RealType eps,
	                 SizeType minSteps) const
	{
		lpp_solver_params p;
		p.steps = steps; p.minsteps = minSteps; p.eps = eps; p.kernel = LPP_KERNEL_AUTO; p.reortho = 0; p.seed = 0;
		const SizeType n = rows();
		eigs.resize(excitedPlusOne);
		zs.resize(excitedPlusOne);
		int32_t ns = 0;
		if (excitedPlusOne == 1) {
			zs[0].resize(n);
			check(lpp_ground_state(handle_, &p, &(init[0]), 1, &(eigs[0]), &(zs[0][0]), 0, 0, &ns));
			return;
		}
		std::vector<double> all(excitedPlusOne*n);
		check(lpp_states_below(handle_, &p, &(init[0]), excitedPlusOne, &(eigs[0]), &(all[0]), &ns));
		for (SizeType k = 0; k < excitedPlusOne; ++k) zs[k].assign(all.begin() + k*n, all.begin() + (k + 1)*n);
	}

private:

	static void check(int status)
	{
		if (status == 0) return;
		err(PsimagLite::String("InternalProductCuda: ") + lpp_last_error() + "\n");
	}

	void init()
	{
		if (!PsimagLite::IsSame<ComplexOrRealType, double>::True)
			err("InternalProductCuda: real double precision only (do not use useComplex)\n");

		const GeometryType& geometry = model_.geometry();
		const SizeType nsite = geometry.numberOfSites();
		const int modelId = model_.cudaModelId();         // the model refuses variants the engine does not implement
		// geometry terms the engine reads: 1 (HubbardOneBand, FeAsBasedSc), 2 (Heisenberg: J+-, Jzz), 4 (Tj1Orbital).  More terms
		// mean an Extended / Super / KaneMele variant (HubbardHelper.h:39-65, FeBasedSc.h ctor) whose couplings would be dropped.
		const SizeType wantTerms = (modelId == LPP_MODEL_HEISENBERG) ? 2 : (modelId == LPP_MODEL_TJ) ? 4 : 1;
		if (geometry.terms() != wantTerms)
			err("InternalProductCuda: this model variant has geometry terms the CUDA engine does not implement\n");
		const SizeType orbitals = (modelId == LPP_MODEL_FEAS) ? model_.orbitals(0) : 1;
		const SizeType nb = nsite*orbitals;

		// term 0 (and term 1 for Heisenberg): the values the models read through geometry_(i,orb,j,orb2,term)
		// HubbardHelper.h:60-71 ; FeBasedSc.h:320-323 ; Heisenberg.h:54-58
		std::vector<double> hop(nb*nb, 0.0), jzz, jpm, w;
		for (SizeType i = 0; i < nsite; ++i)
			for (SizeType o1 = 0; o1 < orbitals; ++o1)
				for (SizeType j = 0; j < nsite; ++j)
					for (SizeType o2 = 0; o2 < orbitals; ++o2)
						hop[(i*orbitals + o1)*nb + j*orbitals + o2] = geometry(i, o1, j, o2, 0);
		if (modelId == LPP_MODEL_HEISENBERG) {
			jzz.resize(nb*nb, 0.0);
			for (SizeType i = 0; i < nsite; ++i)
				for (SizeType j = 0; j < nsite; ++j)
					jzz[i*nb + j] = geometry(i, 0, j, 0, 1);
		}

		if (modelId == LPP_MODEL_TJ) {       // TjMultiOrb.h:68-79: terms 1, 2, 3 = S+S-, SzSz, n n couplings
			jpm.resize(nb*nb, 0.0); jzz.resize(nb*nb, 0.0); w.resize(nb*nb, 0.0);
			for (SizeType i = 0; i < nsite; ++i)
				for (SizeType j = 0; j < nsite; ++j) {
					jpm[i*nb + j] = geometry(i, 0, j, 0, 1);
					jzz[i*nb + j] = geometry(i, 0, j, 0, 2);
					w[i*nb + j] = geometry(i, 0, j, 0, 3);
				}
		}

		typename ProgramGlobals::PairIntType parts = basis_.parts();   // (nup, ndown) or (twiceS, szPlusConst)

		lpp_desc d;
		std::memset(&d, 0, sizeof(d));
		d.model = modelId;
		d.nsite = nsite;
		d.orbitals = orbitals;
		d.nup = (modelId == LPP_MODEL_HEISENBERG) ? parts.second : parts.first;
		d.ndown = (modelId == LPP_MODEL_HEISENBERG) ? 0 : parts.second;
		d.feas_u3_all_pairs = 1;
		d.hop = &(hop[0]);
		d.jzz = jzz.size() ? &(jzz[0]) : 0;
		d.jpm = jpm.size() ? &(jpm[0]) : 0;
		d.w = w.size() ? &(w[0]) : 0;
		model_.exportForCuda(d);            // fills U/nU, V/nV, D/nD from hubbardU, potentialV, anisotropy
		const CudaTopology& topo = CudaTopology::global();
		d.device = topo.device;
		d.rank = topo.rank;
		d.nranks = topo.nranks;
		check(lpp_create(&d, &handle_));
		if (topo.nranks > 1) joinRanks(topo);
	}

	// nranks > 1: rank 0 creates the NCCL id and publishes it in a file, every rank joins the communicator; the CUDA IPC
	// handles of the column shards (peer-memory exchange) travel the same way.  Handles are created in the same order on every
	// rank (ground state first, then one per new sector of Engine::spectralFunction), so a per-process counter names the files.
	void joinRanks(const CudaTopology& topo)
	{
		if (topo.idFile.empty()) err("InternalProductCuda: nranks > 1 needs LPP_NCCL_ID_FILE\n");
		static int serial = 0;
		char tag[32];
		std::snprintf(tag, sizeof(tag), ".%d", serial++);
		const std::string base = topo.idFile + tag;
		uint8_t id[128];
		if (topo.rank == 0) {
			check(lpp_comm_unique_id(id));
			publish(base + ".id", id, 128);
		} else {
			await(base + ".id", id, 128);
		}
		check(lpp_comm_init(handle_, id));
		uint8_t mine[128];
		if (lpp_p2p_export(handle_, LPP_KERNEL_AUTO, mine) != 0) return;     // sharding without peer memory: NCCL paths
		char r[16];
		std::snprintf(r, sizeof(r), ".p2p.%d", topo.rank);
		publish(base + r, mine, 128);
		std::vector<uint8_t> all(128*topo.nranks);
		for (int q = 0; q < topo.nranks; ++q) {
			std::snprintf(r, sizeof(r), ".p2p.%d", q);
			await(base + r, &(all[128*q]), 128);
		}
		check(lpp_p2p_import(handle_, &(all[0])));
	}

	static void publish(const std::string& name, const uint8_t* data, size_t n)
	{
		const std::string tmp = name + ".tmp";
		FILE* f = std::fopen(tmp.c_str(), "wb");
		if (!f || std::fwrite(data, 1, n, f) != n) err("InternalProductCuda: cannot write " + tmp + "\n");
		std::fclose(f);
		if (std::rename(tmp.c_str(), name.c_str()) != 0) err("InternalProductCuda: cannot rename " + tmp + "\n");
	}

	static void await(const std::string& name, uint8_t* data, size_t n)
	{
		for (int tries = 0; tries < 6000; ++tries) {                             // 60 s
			FILE* f = std::fopen(name.c_str(), "rb");
			if (f) {
				const size_t got = std::fread(data, 1, n, f);
				std::fclose(f);
				if (got == n) return;
			}
			usleep(10000);
		}
		err("InternalProductCuda: timed out waiting for " + name + "\n");
	}

	const ModelType& model_;
	const BasisType& basis_;
	lpp_handle* handle_;
}; // class InternalProductCuda

template<typename ModelType, typename SpecialSymmetryType>
struct KrylovPlacement<InternalProductCuda<ModelType, SpecialSymmetryType> > {
	typedef KrylovOnDeviceTag Tag;
};
} // namespace LanczosPlusPlus

#endif // INTERNALPRODUCT_CUDA_H
