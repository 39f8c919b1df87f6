// host/engine_b200.h -- C++ host layer above the C-ABI (include/lpp_b200.h): the part of Engine (Engine.h:84-98, 133-206,
// 460-533) that sits on the hot path.  Ground state on construction; spectralFunction() runs the reference's type loop
// (operator or its conjugate, sum or difference of the two sites), applies the operator on the device
// (accModifiedState_), runs the "Spectral" Lanczos decomposition on the new sector and records the continued fraction
// (a, b, Eg, weight, isign) exactly as cf.set(ab, Eg, weight*s2, -s) does at Engine.h:481-489.
// Header only, no PsimagLite: used by host/lanczos_b200.cpp; inside the reference tree the same calls live behind
// include/InternalProductCuda.h.
#ifndef LPP_ENGINE_B200_H
#define LPP_ENGINE_B200_H

#include <complex>
#include <stdexcept>
#include <string>
#include <vector>
#include "lpp_b200.h"

namespace lppb200 {

// PsimagLite::ContinuedFraction as Engine uses it: G(z) = weight * sum_l I_l / (z - isign (eps_l - Eg))
struct ContinuedFraction {
	int type;
	std::vector<double> a, b;
	double Eg, weight;
	int isign;
	std::vector<std::complex<double> > operator()(const std::vector<double>& omega, double delta) const
	{
		std::vector<double> out(2 * omega.size());
		if (lpp_cf_eval((int32_t)a.size(), a.data(), b.data(), Eg, weight, isign, (int32_t)omega.size(), omega.data(), delta, out.data()) != 0)
			throw std::runtime_error(lpp_last_error());
		std::vector<std::complex<double> > g(omega.size());
		for (size_t i = 0; i < omega.size(); i++) g[i] = std::complex<double>(out[2 * i], out[2 * i + 1]);
		return g;
	}
};

class Engine {
public:
	// Engine.h:84-98: the constructor computes the ground state (computeAllStatesBelow(0), Engine.h:601-657)
	Engine(const lpp_desc& desc, const lpp_solver_params& lanczos, const lpp_solver_params& spectral)
	    : desc_(desc), lanczos_(lanczos), spectral_(spectral), h_(nullptr), energy_(0), steps_(0)
	{
		check(lpp_create(&desc_, &h_));
		int32_t ns = 0;
		check(lpp_ground_state(h_, &lanczos_, nullptr, 1, &energy_, nullptr, nullptr, nullptr, &ns));
		steps_ = ns;
	}
	~Engine() { if (h_) lpp_destroy(h_); }
	Engine(const Engine&) = delete;
	Engine& operator=(const Engine&) = delete;

	double energies(size_t) const { return energy_; }
	int lanczosSteps() const { return steps_; }
	uint64_t rows() const { uint64_t r = 0; lpp_rows(h_, &r); return r; }

	// Engine.h:133-206: c / cdagger of HubbardOneBand, FeAsBasedSc (orbital pair orb0, orb1) and Tj1Orbital; sz / splus / sminus
	// of HubbardOneBand and Heisenberg (hasNewParts: HubbardOneOrbital.h:212-257, BasisFeAsBasedSc.h:305-326, TjMultiOrb.h:538-557,
	// Heisenberg.h:218-240)
	void spectralFunction(std::vector<ContinuedFraction>& cfCollection, int what, int isite, int jsite, int spin, int orb0 = 0,
	                      int orb1 = 0) const
	{
		const bool fermionic = what == LPP_OP_C || what == LPP_OP_CDAGGER;       // LabeledOperator::isFermionic
		int conj;                                                                // LabeledOperator::transposeConjugate
		switch (what) {
		case LPP_OP_C: conj = LPP_OP_CDAGGER; break;
		case LPP_OP_CDAGGER: conj = LPP_OP_C; break;
		case LPP_OP_SPLUS: conj = LPP_OP_SMINUS; break;
		case LPP_OP_SMINUS: conj = LPP_OP_SPLUS; break;
		case LPP_OP_SZ: conj = LPP_OP_SZ; break;
		default: throw std::runtime_error("spectralFunction: operator must be c, cdagger, sz, splus or sminus");
		}
		const bool isDiagonal = isite == jsite && orb0 == orb1;
		const int nmax = desc_.nsite * (desc_.model == LPP_MODEL_FEAS ? desc_.orbitals : 1);
		for (int type = 0; type < 4; type++) {
			if (isDiagonal && type > 1) continue;
			const int op = (type & 1) ? what : conj;                       // Engine.h:163
			lpp_desc d = desc_;
			if (op == LPP_OP_C || op == LPP_OP_CDAGGER) {
				const int c = (op == LPP_OP_C) ? -1 : 1;
				d.nup += (spin == 0) ? c : 0;
				d.ndown += (spin == 1) ? c : 0;
				if (d.nup == 0 && d.ndown == 0) continue;
			} else if (op == LPP_OP_SPLUS || op == LPP_OP_SMINUS) {
				const int c = (op == LPP_OP_SPLUS) ? 1 : -1;
				d.nup += c;
				if (d.model != LPP_MODEL_HEISENBERG) d.ndown -= c;
			}
			if (d.nup < 0 || d.ndown < 0 || d.nup > nmax || d.ndown > nmax) continue;
			if (d.model == LPP_MODEL_TJ && d.nup + d.ndown > d.nsite) continue;           // no double occupancy
			const bool same = op == LPP_OP_SZ;                             // needsNewBasis() is false
			lpp_handle* dst = h_;
			if (!same) check(lpp_create(&d, &dst));
			try {
				const double isign = (type > 1) ? -1.0 : 1.0;
				check(lpp_apply_op(h_, dst, op, isite, spin, orb0, 1.0, 0));   // Engine.h:509-517
				check(lpp_apply_op(h_, dst, op, jsite, spin, orb1, isign, 1)); // Engine.h:523-531
				ContinuedFraction cf;
				cf.a.resize((size_t)spectral_.steps + 1);
				cf.b.resize((size_t)spectral_.steps + 1);
				int32_t n = 0;
				double weight = 0;
				check(lpp_lanczos_decomposition(dst, &spectral_, nullptr, 1, cf.a.data(), cf.b.data(), &n, &weight));   // Engine.h:474-479
				cf.a.resize((size_t)n);
				cf.b.resize((size_t)n);
				const int s = (type & 1) ? -1 : 1;
				double s2 = (type > 1) ? -1.0 : 1.0;
				if (!fermionic) s2 *= s;                                      // Engine.h:482
				if (!isDiagonal) s2 *= 0.5;                                   // Engine.h:481-485
				cf.type = type;
				cf.Eg = energy_;
				cf.weight = weight * s2;
				cf.isign = -s;                                                // cf.set(ab, Eg, weight*s2, -s), Engine.h:489
				cfCollection.push_back(cf);
			} catch (...) {
				if (!same) lpp_destroy(dst);
				throw;
			}
			if (!same) lpp_destroy(dst);
		}
	}

	// Engine.h:262-331 with bra = ket = ground state: result[i * nsite + j] = <O_j gs | O_i gs> (c: <cdagger_j c_i>)
	std::vector<double> twoPoint(int what, int spin, int orb0 = 0, int orb1 = 0) const
	{
		std::vector<double> result((size_t)desc_.nsite * desc_.nsite, 0.0);
		if (what == LPP_OP_N) {
			check(lpp_two_point(h_, h_, what, spin, orb0, orb1, result.data()));
			return result;
		}
		const int c = (what == LPP_OP_C) ? -1 : 1;
		lpp_desc d = desc_;
		d.nup += (spin == 0) ? c : 0;
		d.ndown += (spin == 1) ? c : 0;
		lpp_handle* dst = nullptr;
		check(lpp_create(&d, &dst));
		const int rc = lpp_two_point(h_, dst, what, spin, orb0, orb1, result.data());
		lpp_destroy(dst);
		check(rc);
		return result;
	}

	// Engine.h:341-389 with bra = ket = ground state: <gs| O_n ... O_1 |gs>, O_k = what[k] at (sites[k], spins[k], orbs[k]), what[0]
	// applied first; 0 when the string leaves the allowed particle numbers or does not come back to the ground state's sector
	double manyPoint(const std::vector<int>& sites, const std::vector<int>& what, const std::vector<int>& spins, const std::vector<int>& orbs) const
	{
		const int n = (int)sites.size();
		std::vector<lpp_handle*> chain(1, h_), own;
		std::vector<std::pair<int, int> > parts(1, std::make_pair((int)desc_.nup, (int)desc_.ndown));
		int nup = desc_.nup, ndn = desc_.ndown;
		const bool heis = desc_.model == LPP_MODEL_HEISENBERG;
		const int nmax = desc_.nsite * (desc_.model == LPP_MODEL_FEAS ? desc_.orbitals : 1);
		double r = 0;
		int rc = 0;
		bool zero = false;
		for (int k = 0; k < n && !zero && rc == 0; k++) {
			const int op = what[k], spin = spins[k];
			if (op == LPP_OP_C || op == LPP_OP_CDAGGER) { const int d = (op == LPP_OP_C) ? -1 : 1; if (spin == 0) nup += d; else ndn += d; }
			else if (op == LPP_OP_SPLUS) { nup += 1; if (!heis) ndn -= 1; }
			else if (op == LPP_OP_SMINUS) { nup -= 1; if (!heis) ndn += 1; }
			if (nup < 0 || ndn < 0 || nup > nmax || ndn > nmax) { zero = true; break; }
			if (nup == parts.back().first && ndn == parts.back().second) chain.push_back(chain.back());
			else if (k == n - 1 && nup == (int)desc_.nup && ndn == (int)desc_.ndown) chain.push_back(h_);
			else {
				lpp_desc d = desc_;
				d.nup = nup;
				d.ndown = ndn;
				lpp_handle* s = nullptr;
				rc = lpp_create(&d, &s);
				if (rc == 0) { own.push_back(s); chain.push_back(s); }
			}
			parts.push_back(std::make_pair(nup, ndn));
		}
		if (!zero && rc == 0 && nup == (int)desc_.nup && ndn == (int)desc_.ndown) {
			std::vector<int32_t> o(what.begin(), what.end()), s(sites.begin(), sites.end()), sp(spins.begin(), spins.end()), ob(orbs.begin(), orbs.end());
			rc = lpp_many_point(chain.data(), n, o.data(), s.data(), sp.data(), ob.data(), &r);
		}
		for (lpp_handle* q : own) lpp_destroy(q);
		check(rc);
		return r;
	}

	// Engine.h:208-249 with bra = ket = ground state: <gs| op_0[site_0]; ...; op_{n-1}[site_{n-1}] |gs>, ModelBase::rahulMethod
	// semantics; labels 0 identity, 1 n, 2 sz, 3 c (cdagger when transpose), dof 0 up / 1 down, site = bit position
	struct MeasureOp { int label, dof, site, transpose; };
	double measure(const std::vector<MeasureOp>& ops) const
	{
		std::vector<int32_t> l, d, t, s;
		for (const MeasureOp& o : ops) { l.push_back(o.label); d.push_back(o.dof); t.push_back(o.transpose); s.push_back(o.site); }
		double r = 0;
		check(lpp_measure(h_, (int32_t)ops.size(), l.data(), d.data(), t.data(), s.data(), &r));
		return r;
	}

private:
	static void check(int status)
	{
		if (status != 0) throw std::runtime_error(lpp_last_error());
	}
	lpp_desc desc_;
	lpp_solver_params lanczos_, spectral_;
	lpp_handle* h_;
	double energy_;
	int steps_;
};

} // namespace lppb200
#endif
