// lpp_device.cuh -- bit-string bases, ranks, fermion signs and Hamiltonian row generators.
// __host__ __device__ so that tests/ can run the identical code on the CPU against the oracle; the product
// only ever calls them from kernels.  Reference citations are file:line under /root/reference/src.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define LPP_HD __host__ __device__ __forceinline__
#else
#define LPP_HD inline
#endif

typedef uint64_t word_t;

#define LPP_BINOM_N 65        // binom[n*65+k] = C(n,k), 0 when k>n
#define LPP_MAX_ORB 4
#define LPP_MODEL_HUBBARD 0
#define LPP_MODEL_FEAS 1
#define LPP_MODEL_HEISENBERG 2
#define LPP_MODEL_TJ 3        // TjMultiOrb with Orbitals=1 (Tj1Orbital): no double occupancy

LPP_HD int lpp_popc(word_t a)
{
#ifdef __CUDA_ARCH__
	return __popcll(a);
#else
	return __builtin_popcountll(a);
#endif
}

LPP_HD int lpp_ctz(word_t a)
{
#ifdef __CUDA_ARCH__
	return __ffsll((long long)a) - 1;
#else
	return __builtin_ctzll(a);
#endif
}

LPP_HD word_t lpp_bit(int i) { return ((word_t)1) << i; }
LPP_HD word_t lpp_below(int i) { return (((word_t)1) << i) - 1; }  // bits [0,i)

// Everything a kernel needs to know about one (model, sector). Pointers are device pointers in kernels
// (host pointers in the CPU-side unit tests).
struct ModelDev {
	int model, nsite, orbitals, nbits;
	int nup, ndn;
	int u3_all_pairs;
	uint64_t n1, n2;        // one-spin basis sizes (Heisenberg: n2 = 1)
	uint64_t rows;          // n1*n2
	const word_t* b1;       // basis words, spin up (or Heisenberg words)
	const word_t* b2;       // spin down
	const uint64_t* binom;  // LPP_BINOM_N^2
	const double* hop;      // nbits x nbits
	const double* jzz;      // Heisenberg term 1 / t-J term 2
	const double* jpm;      // t-J term 1 (S+S- couplings); unused otherwise
	const double* w;        // t-J term 3 (n_i n_j couplings); unused otherwise
	const double* U;
	const double* V;
	const double* D;
	// FeAs partition bookkeeping (BasisOneSpinFeAs.h:56-84): block offset by key = sum_{o>=1} n_o*(npart+1)^(o-1)
	const uint64_t* part_off1;
	const uint64_t* part_off2;
	// FeAs enumeration: partition p starts at part_start[p]; occupations part_n[p*LPP_MAX_ORB+o]
	const uint64_t* part_start1;
	const int* part_n1;
	int nparts1;
	const uint64_t* part_start2;
	const int* part_n2;
	int nparts2;
	// split rank tables for colex ranks (optional): rank = rlo[w & lomask] + rhi[popc(lo)*hisize + (w >> lobits)]
	const uint32_t* rlo;
	const uint32_t* rhi;
	int lobits;
	// word -> index look-up tables for FeAs (optional)
	const uint32_t* lut1;
	const uint32_t* lut2;
};

// ---------------------------------------------------------------- colex (combinatorial number system)
// rank: BasisOneSpin.h:73-81  n = sum_c C(b_c, c)
LPP_HD uint64_t lpp_rank_colex(const uint64_t* binom, word_t w)
{
	uint64_t r = 0;
	int c = 1;
	while (w) {
		int b = lpp_ctz(w);
		r += binom[b * LPP_BINOM_N + c];
		c++;
		w &= w - 1;
	}
	return r;
}

// inverse of the enumeration BasisOneSpin.h:52-62 (ascending numeric order of n-bit words with k bits set)
LPP_HD word_t lpp_unrank_colex(const uint64_t* binom, int nsite, int npart, uint64_t idx)
{
	word_t w = 0;
	int b = nsite - 1;
	for (int c = npart; c >= 1; c--) {
		while (binom[b * LPP_BINOM_N + c] > idx) b--;
		w |= lpp_bit(b);
		idx -= binom[b * LPP_BINOM_N + c];
		b--;
	}
	return w;
}

// ---------------------------------------------------------------- FeAs one-spin basis
// word -> per-orbital words (BasisOneSpinFeAs.h:384-398)
LPP_HD void lpp_feas_uncollate(word_t w, int no, word_t* kets)
{
	for (int o = 0; o < LPP_MAX_ORB; o++) kets[o] = 0;
	while (w) {
		int b = lpp_ctz(w);
		kets[b % no] |= lpp_bit(b / no);
		w &= w - 1;
	}
}

// closed-form replacement of the linear search BasisOneSpinFeAs.h:96-101, following the enumeration order of
// the constructor :62-84 with collateBasis/getKets :297-331 (orbital 0 is the fastest digit).
LPP_HD uint64_t lpp_rank_feas(const ModelDev& m, int spin, word_t w)
{
	const int no = m.orbitals;
	const int npart = spin ? m.ndn : m.nup;
	word_t kets[LPP_MAX_ORB];
	lpp_feas_uncollate(w, no, kets);
	uint64_t key = 0, mul = 1;
	for (int o = 1; o < no; o++) {
		key += (uint64_t)lpp_popc(kets[o]) * mul;
		mul *= (uint64_t)(npart + 1);
	}
	uint64_t idx = (spin ? m.part_off2 : m.part_off1)[key];
	uint64_t stride = 1;
	for (int o = 0; o < no; o++) {
		idx += lpp_rank_colex(m.binom, kets[o]) * stride;
		stride *= m.binom[m.nsite * LPP_BINOM_N + lpp_popc(kets[o])];
	}
	return idx;
}

LPP_HD word_t lpp_unrank_feas(const ModelDev& m, int spin, uint64_t idx)
{
	const int no = m.orbitals;
	const uint64_t* start = spin ? m.part_start2 : m.part_start1;
	const int* pn = spin ? m.part_n2 : m.part_n1;
	const int np = spin ? m.nparts2 : m.nparts1;
	int p = 0;
	while (p + 1 < np && start[p + 1] <= idx) p++;
	uint64_t local = idx - start[p];
	word_t w = 0;
	for (int o = 0; o < no; o++) {
		int n_o = pn[p * LPP_MAX_ORB + o];
		uint64_t size_o = m.binom[m.nsite * LPP_BINOM_N + n_o];
		uint64_t digit = local % size_o;
		local /= size_o;
		word_t k = lpp_unrank_colex(m.binom, m.nsite, n_o, digit);
		while (k) {  // getCollatedKet, BasisOneSpinFeAs.h:357-374: bit position = site*orbitals + orb
			int s = lpp_ctz(k);
			w |= lpp_bit(s * no + o);
			k &= k - 1;
		}
	}
	return w;
}

// perfectIndex of a one-spin word
LPP_HD uint64_t lpp_rank_onespin(const ModelDev& m, int spin, word_t w)
{
	if (m.model == LPP_MODEL_FEAS) {
		const uint32_t* lut = spin ? m.lut2 : m.lut1;
		if (lut) return lut[w];
		return lpp_rank_feas(m, spin, w);
	}
	if (m.rlo) {
		word_t lo = w & lpp_below(m.lobits);
		return (uint64_t)m.rlo[lo] + (uint64_t)m.rhi[((uint64_t)lpp_popc(lo) << (m.nbits - m.lobits)) + (w >> m.lobits)];
	}
	return lpp_rank_colex(m.binom, w);
}

// ---------------------------------------------------------------- signs
// ProgramGlobals.h:109-114 with a 64-bit mask
LPP_HD int lpp_sign_below(word_t a, int i) { return (lpp_popc(a & lpp_below(i)) & 1) ? -1 : 1; }

// parity of the occupied bits in [lo, hi)
LPP_HD int lpp_sign_range(word_t a, int lo, int hi) { return (lpp_popc(a & (lpp_below(hi) ^ lpp_below(lo))) & 1) ? -1 : 1; }

// BasisOneSpinFeAs.h:150-181 (i<j: counted bits are exactly [ii, jj)) and :252-263 (i==j)
LPP_HD int lpp_feas_dosign(word_t ket, int i, int orb1, int j, int orb2, int no)
{
	int ii = i * no + orb1, jj = j * no + orb2;
	if (i == j) return (orb1 > orb2) ? -lpp_sign_range(ket, jj, ii) : lpp_sign_range(ket, ii, jj);
	return lpp_sign_range(ket, ii, jj);
}

// ---------------------------------------------------------------- diagonals
// HubbardHelper.h:138-189 (HubbardOneBand: U n_up n_dn + V (n_up + n_dn), V indexed [i] only)
LPP_HD double lpp_hubbard_diag(const ModelDev& m, word_t k1, word_t k2)
{
	double s = 0;
	for (int i = 0; i < m.nsite; i++) {
		int nu = (int)((k1 >> i) & 1), nd = (int)((k2 >> i) & 1);
		s += m.U[i] * (double)(nu * nd);
		double v = m.V[i];
		if (v != 0) s += v * (double)(nu + nd);
	}
	return s;
}

LPP_HD int lpp_feas_occ(word_t k, int site, int orb, int no) { return (int)((k >> (site * no + orb)) & 1); }

// FeBasedSc.h:534-571 + findSnoDecay :573-623 (INT_PAPER33, one geometry term, no SpinOrbit)
LPP_HD double lpp_feas_diag(const ModelDev& m, word_t k1, word_t k2)
{
	const int no = m.orbitals, nsite = m.nsite;
	double s = 0;
	for (int i = 0; i < nsite; i++) {
		double szOrb = 0;
		for (int orb = 0; orb < no; orb++) {
			int u1 = lpp_feas_occ(k1, i, orb, no), d1 = lpp_feas_occ(k2, i, orb, no);
			double t = m.U[0] * (double)(u1 * d1);
			for (int orb2 = orb + 1; orb2 < no; orb2++) {
				int u2 = lpp_feas_occ(k1, i, orb2, no), d2 = lpp_feas_occ(k2, i, orb2, no);
				t += m.U[1] * (double)((u1 + d1) * (u2 + d2));
				t += m.U[4] * (0.5 * (u1 - d1)) * (0.5 * (u2 - d2));
				t += m.U[5] * (double)(u1 * u2);
				t += m.U[5] * (double)(d1 * d2);
			}
			s += t;
			s += m.V[i + (orb + no * 0) * nsite] * (double)u1 + m.V[i + (orb + no * 1) * nsite] * (double)d1;
			szOrb += 0.5 * (u1 - d1);
		}
		s += m.D[0] * szOrb * szOrb;
	}
	return s;
}

// Heisenberg.h:242-276 for S=1/2 (m_i = bit - 1/2)
LPP_HD double lpp_heis_diag(const ModelDev& m, word_t ket)
{
	const int nsite = m.nsite;
	double s = 0;
	for (int i = 0; i < nsite; i++) {
		double t1 = (double)((ket >> i) & 1) - 0.5;
		s += m.V[i] * t1;
		s += m.D[i] * (t1 * t1);
		for (int j = i + 1; j < nsite; j++) {
			double jz = m.jzz[i * nsite + j];
			double t2 = (double)((ket >> j) & 1) - 0.5;
			s += t1 * t2 * jz;
		}
	}
	return s;
}

// ---------------------------------------------------------------- one-spin hop generators
// emit(target one-spin index, amplitude)
// Hubbard: HubbardHelper.h:204-241: for (i,j) with hoppings_(i,j)!=0, bit_i=1, bit_j=0:
//   bra = ket^m_i^m_j, value = h(i,j) * doSign(ket,i) * doSign(ket^m_i, j)
template <class Emit>
LPP_HD void lpp_hubbard_spin_hops(const ModelDev& m, int spin, word_t ket, Emit& emit)
{
	const int nsite = m.nsite;
	word_t occ = ket;
	while (occ) {
		int i = lpp_ctz(occ);
		occ &= occ - 1;
		word_t bra0 = ket ^ lpp_bit(i);
		int si = lpp_sign_below(ket, i);
		for (int j = 0; j < nsite; j++) {
			double h = m.hop[i * nsite + j];
			if (h == 0 || ((ket >> j) & 1)) continue;
			int sj = lpp_sign_below(bra0, j);
			emit(lpp_rank_onespin(m, spin, bra0 ^ lpp_bit(j)), h * (double)(si * sj));
		}
	}
}

// FeAs: FeBasedSc.h:325-374: site pairs j>=i, all (orb,orb2), h = -geometry(ii,jj); exactly one of (ii,jj) occupied
template <class Emit>
LPP_HD void lpp_feas_spin_hops(const ModelDev& m, int spin, word_t ket, Emit& emit)
{
	const int no = m.orbitals, nsite = m.nsite, nb = m.nbits;
	for (int i = 0; i < nsite; i++) {
		for (int orb = 0; orb < no; orb++) {
			int ii = i * no + orb;
			int si = (int)((ket >> ii) & 1);
			for (int j = i; j < nsite; j++) {
				for (int orb2 = 0; orb2 < no; orb2++) {
					int jj = j * no + orb2;
					double h = -m.hop[ii * nb + jj];
					if (h == 0) continue;
					int sj = (int)((ket >> jj) & 1);
					if (si + sj != 1) continue;
					word_t bra = ket ^ (lpp_bit(ii) | lpp_bit(jj));
					double extra = si ? -1.0 : 1.0;
					double sg = (double)lpp_feas_dosign(ket, i, orb, j, orb2, no);
					emit(lpp_rank_onespin(m, spin, bra), h * extra * sg);
				}
			}
		}
	}
}

// ---------------------------------------------------------------- FeAs two-spin terms
// emit(up index, down index, amplitude): on-site inter-orbital spin exchange (U[2], FeBasedSc.h:376-411,678-695)
// and pair hopping (U[3], :414-432,697-713).  all_pairs: stored definition loops orb2 != orb (:192-197),
// the literal on-the-fly doTask loops orb2 > orb only (:85-88).
template <class Emit>
LPP_HD void lpp_feas_twospin(const ModelDev& m, word_t k1, word_t k2, int all_pairs, Emit& emit)
{
	const int no = m.orbitals, nsite = m.nsite;
	const double u2 = 0.5 * m.U[2], u3 = m.U[3];
	// two orbitals: a term exists only on sites holding exactly one up and exactly one down electron
	word_t cand = ~(word_t)0;
	if (no == 2) cand = (k1 ^ (k1 >> 1)) & (k2 ^ (k2 >> 1));
	for (int i = 0; i < nsite; i++) {
		if (no == 2 && !((cand >> (2 * i)) & 1)) continue;
		for (int orb1 = 0; orb1 < no; orb1++) {
			for (int orb2 = 0; orb2 < no; orb2++) {
				if (orb1 == orb2) continue;
				int up1 = lpp_feas_occ(k1, i, orb1, no), up2 = lpp_feas_occ(k1, i, orb2, no);
				int dn1 = lpp_feas_occ(k2, i, orb1, no), dn2 = lpp_feas_occ(k2, i, orb2, no);
				if (!(up2 == 1 && up1 == 0)) continue;
				word_t mk = lpp_bit(i * no + orb1) | lpp_bit(i * no + orb2);
				double sg = (double)(lpp_feas_dosign(k1, i, orb1, i, orb2, no) * lpp_feas_dosign(k2, i, orb1, i, orb2, no));
				if (dn1 == 1 && dn2 == 0 && u2 != 0)
					emit(lpp_rank_onespin(m, 0, k1 ^ mk), lpp_rank_onespin(m, 1, k2 ^ mk), u2 * sg);
				if (dn1 == 0 && dn2 == 1 && (all_pairs || orb2 > orb1) && u3 != 0)
					emit(lpp_rank_onespin(m, 0, k1 ^ mk), lpp_rank_onespin(m, 1, k2 ^ mk), -u3 * sg);
			}
		}
	}
}

// ---------------------------------------------------------------- Heisenberg off-diagonal
// Heisenberg.h:94-106 + :278-307, S=1/2: site i down, site j up, jpm(i,j)!=0 -> flip both, value jpm(i,j)/2
template <class Emit>
LPP_HD void lpp_heis_offdiag(const ModelDev& m, word_t ket, Emit& emit)
{
	const int nsite = m.nsite;
	word_t up = ket;
	while (up) {
		int j = lpp_ctz(up);
		up &= up - 1;
		for (int i = 0; i < nsite; i++) {
			if ((ket >> i) & 1) continue;
			double jpm = m.hop[i * nsite + j];
			if (jpm == 0) continue;
			word_t bra = (ket | lpp_bit(i)) & ~lpp_bit(j);
			emit(lpp_rank_onespin(m, 0, bra), 0.5 * jpm);
		}
	}
}

// ---------------------------------------------------------------- t-J (TjMultiOrb.h with Orbitals=1)
// Basis (BasisTjMultiOrbLanczos.h:30-42,351-366): all (up, down) pairs without doubly occupied sites, as combined words
// (down << nsite) | up sorted ascending.  Ascending order = down words in colex order (slow index i2), and for a fixed down
// word the up words that avoid its sites in ascending order = colex order of the up word compressed onto the nsite - ndn
// free sites (fast index i1).  rows = C(nsite, ndn) * C(nsite - ndn, nup); b2 = colex(nsite, ndn), b1 = colex(nsite - ndn, nup).
LPP_HD word_t lpp_deposit(word_t compressed, word_t freemask)      // k-th bit of `compressed` -> k-th set bit of `freemask`
{
	word_t out = 0;
	while (compressed) {
		const int b = lpp_ctz(freemask);
		if (compressed & 1) out |= lpp_bit(b);
		compressed >>= 1;
		freemask &= freemask - 1;
	}
	return out;
}
LPP_HD word_t lpp_extract(word_t w, word_t freemask)               // inverse of lpp_deposit
{
	word_t out = 0;
	int k = 0;
	while (freemask) {
		const int b = lpp_ctz(freemask);
		if ((w >> b) & 1) out |= lpp_bit(k);
		k++;
		freemask &= freemask - 1;
	}
	return out;
}
LPP_HD uint64_t lpp_rank_colex_any(const ModelDev& m, word_t w)
{
	if (m.rlo) {
		word_t lo = w & lpp_below(m.lobits);
		return (uint64_t)m.rlo[lo] + (uint64_t)m.rhi[((uint64_t)lpp_popc(lo) << (m.nbits - m.lobits)) + (w >> m.lobits)];
	}
	return lpp_rank_colex(m.binom, w);
}
// closed form of the search BasisTjMultiOrbLanczos.h:71-112
LPP_HD uint64_t lpp_tj_rank(const ModelDev& m, word_t k1, word_t k2)
{
	const word_t freemask = ~k2 & lpp_below(m.nsite);
	return lpp_rank_colex_any(m, lpp_extract(k1, freemask)) + lpp_rank_colex_any(m, k2) * m.n1;
}
// TjMultiOrb.h:586-647 for one orbital (proij = 1): potentialV[i] n_up + potentialV[i+nsite] n_dn,
// sum_{i<j} (n_iu - n_id)(n_ju - n_jd) jzz(i,j)/4 + (n_iu + n_id)(n_ju + n_jd) w(i,j)
LPP_HD double lpp_tj_diag(const ModelDev& m, word_t k1, word_t k2)
{
	const int nsite = m.nsite;
	double s = 0;
	for (int i = 0; i < nsite; i++) {
		const int niu = (int)((k1 >> i) & 1), nid = (int)((k2 >> i) & 1);
		s += m.V[i] * (double)niu;
		s += m.V[i + nsite] * (double)nid;
		for (int j = i + 1; j < nsite; j++) {
			const int nju = (int)((k1 >> j) & 1), njd = (int)((k2 >> j) & 1);
			s += (double)((niu - nid) * (nju - njd)) * m.jzz[i * nsite + j] * 0.25;
			s += (double)((niu + nid) * (nju + njd)) * m.w[i * nsite + j];
		}
	}
	return s;
}
// TjMultiOrb.h:649-695 (projected hopping, j >= i; value h(i,j) (-1)^{bits of the moving species strictly between i and j})
// and :697-771 with :773-801 (S+S-: h = jpm(i,j)/2 times the parities of bra1 and bra2 over [i, j))
template <class Emit>
LPP_HD void lpp_tj_offdiag(const ModelDev& m, word_t k1, word_t k2, Emit& emit)
{
	const int nsite = m.nsite;
	for (int i = 0; i < nsite; i++) {
		const int s1i = (int)((k1 >> i) & 1), s2i = (int)((k2 >> i) & 1);
		for (int j = i; j < nsite; j++) {
			const double h = m.hop[i * nsite + j];
			if (h == 0) continue;
			const int s1j = (int)((k1 >> j) & 1), s2j = (int)((k2 >> j) & 1);
			if (s1i + s1j == 1 && !(s1j == 0 && s2j > 0) && !(s1j > 0 && s2i > 0)) {
				const word_t bra1 = k1 ^ (lpp_bit(i) | lpp_bit(j));
				const double extra = (s1i == 1) ? -1.0 : 1.0;
				emit(lpp_tj_rank(m, bra1, k2), h * extra * (double)lpp_sign_range(k1, i, j));
			}
			if (s2i + s2j == 1 && !(s2j == 0 && s1j > 0) && !(s2j > 0 && s1i > 0)) {
				const word_t bra2 = k2 ^ (lpp_bit(i) | lpp_bit(j));
				const double extra = (s2i == 1) ? -1.0 : 1.0;
				emit(lpp_tj_rank(m, k1, bra2), h * extra * (double)lpp_sign_range(k2, i, j));
			}
		}
		for (int j = i; j < nsite; j++) {
			const double h = m.jpm[i * nsite + j] * 0.5;
			if (h == 0) continue;
			const int s1j = (int)((k1 >> j) & 1), s2j = (int)((k2 >> j) & 1);
			if (s1i == 1 && s1j == 0 && s2i == 0 && s2j == 1) {
				const word_t bra1 = (k1 ^ lpp_bit(i)) | lpp_bit(j);
				const word_t bra2 = (k2 | lpp_bit(i)) ^ lpp_bit(j);
				emit(lpp_tj_rank(m, bra1, bra2), h * (double)(lpp_sign_range(bra1, i, j) * lpp_sign_range(bra2, i, j)));
			}
			if (s1i == 0 && s1j == 1 && s2i == 1 && s2j == 0) {
				const word_t bra1 = (k1 | lpp_bit(i)) ^ lpp_bit(j);
				const word_t bra2 = (k2 ^ lpp_bit(i)) | lpp_bit(j);
				emit(lpp_tj_rank(m, bra1, bra2), h * (double)(lpp_sign_range(bra1, i, j) * lpp_sign_range(bra2, i, j)));
			}
		}
	}
}

// ---------------------------------------------------------------- full rows
struct LppRowKets {
	word_t k1, k2;
	uint64_t i1, i2;
};

// BasisHubbardLanczos.h:77-84 / BasisFeAsBasedSc.h:84-89: up index fast, down index slow
LPP_HD LppRowKets lpp_row_kets(const ModelDev& m, uint64_t row)
{
	LppRowKets k;
	if (m.model == LPP_MODEL_HEISENBERG) {
		k.i1 = row; k.i2 = 0; k.k1 = m.b1[row]; k.k2 = 0;
	} else if (m.model == LPP_MODEL_TJ) {
		k.i1 = row % m.n1; k.i2 = row / m.n1; k.k2 = m.b2[k.i2];
		k.k1 = lpp_deposit(m.b1[k.i1], ~k.k2 & lpp_below(m.nsite));
	} else {
		k.i1 = row % m.n1; k.i2 = row / m.n1; k.k1 = m.b1[k.i1]; k.k2 = m.b2[k.i2];
	}
	return k;
}

LPP_HD double lpp_row_diag(const ModelDev& m, const LppRowKets& k)
{
	if (m.model == LPP_MODEL_HUBBARD) return lpp_hubbard_diag(m, k.k1, k.k2);
	if (m.model == LPP_MODEL_FEAS) return lpp_feas_diag(m, k.k1, k.k2);
	if (m.model == LPP_MODEL_TJ) return lpp_tj_diag(m, k.k1, k.k2);
	return lpp_heis_diag(m, k.k1);
}

template <class Emit>
struct LppUpAdapter {
	Emit& e; uint64_t i2, n1;
	LPP_HD void operator()(uint64_t t, double v) { e(t + i2 * n1, v); }
};
template <class Emit>
struct LppDnAdapter {
	Emit& e; uint64_t i1, n1;
	LPP_HD void operator()(uint64_t t, double v) { e(i1 + t * n1, v); }
};
template <class Emit>
struct LppTwoAdapter {
	Emit& e; uint64_t n1;
	LPP_HD void operator()(uint64_t a, uint64_t b, double v) { e(a + b * n1, v); }
};

// all off-diagonal entries of one row, emit(global column, value); stored selects the U3 definition for FeAs
template <class Emit>
LPP_HD void lpp_row_offdiag(const ModelDev& m, const LppRowKets& k, int stored, Emit& emit)
{
	if (m.model == LPP_MODEL_HEISENBERG) {
		lpp_heis_offdiag(m, k.k1, emit);
		return;
	}
	if (m.model == LPP_MODEL_TJ) {
		lpp_tj_offdiag(m, k.k1, k.k2, emit);
		return;
	}
	LppUpAdapter<Emit> up{emit, k.i2, m.n1};
	LppDnAdapter<Emit> dn{emit, k.i1, m.n1};
	if (m.model == LPP_MODEL_HUBBARD) {
		lpp_hubbard_spin_hops(m, 0, k.k1, up);
		lpp_hubbard_spin_hops(m, 1, k.k2, dn);
	} else {
		lpp_feas_spin_hops(m, 0, k.k1, up);
		lpp_feas_spin_hops(m, 1, k.k2, dn);
		LppTwoAdapter<Emit> two{emit, m.n1};
		lpp_feas_twospin(m, k.k1, k.k2, stored ? 1 : m.u3_all_pairs, two);
	}
}

// perfectIndex(ket1, ket2) of the full basis (BasisHubbardLanczos.h:59-63, BasisFeAsBasedSc.h:91-100, BasisHeisenberg.h:73-80,
// BasisTjMultiOrbLanczos.h:71-112)
LPP_HD uint64_t lpp_rank_pair(const ModelDev& m, word_t k1, word_t k2)
{
	if (m.model == LPP_MODEL_HEISENBERG) return lpp_rank_onespin(m, 0, k1);
	if (m.model == LPP_MODEL_TJ) return lpp_tj_rank(m, k1, k2);
	return lpp_rank_onespin(m, 0, k1) + lpp_rank_onespin(m, 1, k2) * m.n1;
}

// ---------------------------------------------------------------- Green-function sign and operator targets
// BasisHubbardLanczos.h:106-137 (literal: for SPIN_DOWN and ind>0 the up-parity is overwritten)
LPP_HD int lpp_hubbard_sign_gf(word_t a, word_t b, int ind, int sector)
{
	if (sector == 0) {
		if (ind == 0) return 1;
		int s = lpp_sign_range(a, 1, ind);
		if (a & 1) s = -s;
		return s;
	}
	int s = (lpp_popc(a) & 1) ? -1 : 1;
	if (ind == 0) return s;
	s = lpp_sign_range(b, 1, ind);
	if (b & 1) s = -s;
	return s;
}

// ---------------------------------------------------------------- split colex rank tables
// rank(w) = rlo[lo] + rhi[popc(lo) << hibits | hi]   (lo = low `lobits` bits of w)
LPP_HD uint32_t lpp_split_lo_entry(const uint64_t* binom, uint64_t i) { return (uint32_t)lpp_rank_colex(binom, (word_t)i); }
LPP_HD uint32_t lpp_split_hi_entry(const uint64_t* binom, int lobits, int hibits, uint64_t i)
{
	int p = (int)(i >> hibits);
	word_t h = i & ((((word_t)1) << hibits) - 1);
	uint64_t r = 0;
	int c = p + 1;
	while (h) {
		int b = lpp_ctz(h);
		r += binom[(b + lobits) * LPP_BINOM_N + c];
		c++;
		h &= h - 1;
	}
	return (uint32_t)r;
}

// ---------------------------------------------------------------- stored rows (PsimagLite::SparseRow semantics)
#define LPP_ROW_CAP 256

struct LppBufEmit {
	uint64_t* c;
	double* v;
	int n;
	LPP_HD void operator()(uint64_t col, double val)
	{
		if (n < LPP_ROW_CAP) { c[n] = col; v[n] = val; }
		n++;
	}
};

// SparseRow::finalize(CrsMatrix&): stable sort by column, duplicates summed, explicit zeros kept
LPP_HD int lpp_sort_merge(uint64_t* c, double* v, int n)
{
	for (int i = 1; i < n; i++) {
		uint64_t cc = c[i];
		double vv = v[i];
		int j = i - 1;
		while (j >= 0 && c[j] > cc) { c[j + 1] = c[j]; v[j + 1] = v[j]; j--; }
		c[j + 1] = cc;
		v[j + 1] = vv;
	}
	if (n == 0) return 0;
	int k = 0;
	for (int i = 1; i < n; i++) {
		if (c[i] == c[k]) v[k] += v[i];
		else { k++; c[k] = c[i]; v[k] = v[i]; }
	}
	return k + 1;
}

// one row of setupHamiltonian: diagonal always stored first (HubbardHelper.h:93, FeBasedSc.h:184, Heisenberg.h:100).
// returns the merged entry count, or -1 when the row has more than LPP_ROW_CAP raw entries
LPP_HD int lpp_stored_row(const ModelDev& m, uint64_t r, uint64_t* c, double* v)
{
	LppRowKets k = lpp_row_kets(m, r);
	LppBufEmit e{c, v, 0};
	e(r, lpp_row_diag(m, k));
	lpp_row_offdiag(m, k, 1, e);
	if (e.n > LPP_ROW_CAP) return -1;
	return lpp_sort_merge(c, v, e.n);
}

// ---------------------------------------------------------------- operator application (Engine.h:416-458) as a gather
// For destination row r of `dst`: the unique source row of `src` (if any) and the Green-function sign doSignGf.
// c:       dst word has the orbital empty, source = dst | bit      (BasisOneSpin.h:127-134, BasisOneSpinFeAs.h:127-143,
//          BasisTjMultiOrbLanczos.h:413-433); t-J: the source must not be doubly occupied (:400-411)
// cdagger: dst word has the orbital occupied, source = dst ^ bit
// n:       same sector, orbital occupied (HubbardOneBand only)
// Signs: Hubbard literal BasisHubbardLanczos.h:106-137 (quirk C.3); FeAs BasisFeAsBasedSc.h:170-178 with
// BasisOneSpinFeAs.h:227-239; t-J BasisTjMultiOrbLanczos.h:163-192: parity of the same-species electrons below the orbital,
// times the parity of all up electrons for a down operator.
// Spin operators (not fermionic: mysign = 1 except doSignSpSm for S+-), gather form of
//   HubbardOneBand: getBraIndexSz / getBraIndexSplusSminus (BasisHubbardLanczos.h:221-257) with doSignSpSm (:151-160)
//   Heisenberg S=1/2: getBraIndex_ / getBraIndexSplusSminus (BasisHeisenberg.h:230-280): sz value 1 - 2 bit, n value bit (spin 0) or 1 - bit
// op: 2 sz (Hubbard, Heisenberg), 4 n (Heisenberg), 5 splus, 6 sminus (all four models)
LPP_HD bool lpp_apply_spin_op_source(const ModelDev& src, const ModelDev& dst, int op, int site, int spin, int orb, uint64_t r, word_t b1,
                                     word_t b2, uint64_t* src_row, double* sign)
{
	if (dst.model == LPP_MODEL_HEISENBERG) {
		const word_t ms = lpp_bit(site);
		const bool up = (b1 & ms) != 0;
		if (op == 2) { *src_row = r; *sign = up ? -1.0 : 1.0; return true; }
		if (op == 4) { if ((spin == 0) != up) return false; *src_row = r; *sign = 1.0; return true; }
		word_t ket;
		if (op == 5) { if (!up) return false; ket = b1 ^ ms; }         // S+ |ket>: the site was down in the source
		else { if (up) return false; ket = b1 | ms; }
		*src_row = lpp_rank_onespin(src, 0, ket);
		*sign = 1.0;
		return true;
	}
	// two-species bases: HubbardOneBand (BasisHubbardLanczos.h:221-257), FeAsBasedSc per orbital (BasisFeAsBasedSc.h:291-303,
	// 356-379 with doSignSpSm :202-211), Tj1Orbital (BasisTjMultiOrbLanczos.h:213-242; doSignSpSm is the BasisBase default 1)
	const int pos = site * dst.orbitals + orb;
	const word_t ms = lpp_bit(pos);
	const bool nu = (b1 & ms) != 0, nd = (b2 & ms) != 0;
	if (op == 2) {
		if (dst.model != LPP_MODEL_HUBBARD || nu == nd) return false;
		*src_row = r; *sign = nu ? 1.0 : -1.0;
		return true;
	}
	if (op != 5 && op != 6) return false;
	word_t k1, k2;
	if (op == 5) { if (!nu || nd) return false; k1 = b1 ^ ms; k2 = b2 | ms; }   // source: up empty, down occupied
	else { if (nu || !nd) return false; k1 = b1 | ms; k2 = b2 ^ ms; }
	*sign = (dst.model == LPP_MODEL_TJ) ? 1.0 : (double)(lpp_sign_below(k1, pos) * lpp_sign_below(k2, pos));   // doSignSpSm of the source words
	*src_row = lpp_rank_pair(src, k1, k2);
	return true;
}

LPP_HD bool lpp_apply_op_source(const ModelDev& src, const ModelDev& dst, int op, int site, int spin, int orb, uint64_t r,
                                uint64_t* src_row, double* sign)
{
	const LppRowKets k = lpp_row_kets(dst, r);
	const word_t b1 = k.k1, b2 = k.k2;
	if (op == 2 || op == 5 || op == 6 || dst.model == LPP_MODEL_HEISENBERG)
		return lpp_apply_spin_op_source(src, dst, op, site, spin, orb, r, b1, b2, src_row, sign);
	const word_t bra = spin == 0 ? b1 : b2;
	const int pos = site * dst.orbitals + orb;
	const word_t ms = lpp_bit(pos);
	word_t ket;
	if (op == 1) { if (bra & ms) return false; ket = bra | ms; }
	else if (op == 3) { if (!(bra & ms)) return false; ket = bra ^ ms; }
	else { if (!(bra & ms)) return false; ket = bra; }
	const word_t k1 = spin == 0 ? ket : b1, k2 = spin == 0 ? b2 : ket;
	if (src.model == LPP_MODEL_TJ && (k1 & k2)) return false;
	if (src.model == LPP_MODEL_HUBBARD) {
		*sign = (op == 1 || op == 3) ? (double)lpp_hubbard_sign_gf(k1, k2, site, spin) : 1.0;
	} else {
		int sg = lpp_sign_below(spin == 0 ? k1 : k2, pos);
		if (spin == 1 && (lpp_popc(k1) & 1)) sg = -sg;
		*sign = (double)sg;
	}
	*src_row = lpp_rank_pair(src, k1, k2);
	return true;
}

// ---------------------------------------------------------------- Engine::measure (Engine.h:208-249) -> ModelBase::rahulMethod
// A product of one-site operators applied to a basis state, right to left (ModelBase.h:89-141 with RahulOperator.h:26-50):
// label 0 identity, 1 n, 2 sz (value -1/2 when the orbital is occupied, +1/2 otherwise), 3 c (cdagger when transpose != 0);
// dof 0 acts on the up word, 1 on the down word; fermionic operators carry the parity of the orbitals below `site` in their own
// (updated) word and, for dof 1, the parity of all up electrons.  `site` is the bit position in the word.
#define LPP_MAX_MEASURE_OPS 8
struct LppMeasureOps {
	int n;
	int label[LPP_MAX_MEASURE_OPS], dof[LPP_MAX_MEASURE_OPS], transpose[LPP_MAX_MEASURE_OPS], site[LPP_MAX_MEASURE_OPS];
};
// returns false when the product annihilates the state; otherwise the target words and the accumulated factor
LPP_HD bool lpp_rahul_apply(const LppMeasureOps& ops, word_t k1, word_t k2, word_t* o1, word_t* o2, double* value)
{
	double v = 1.0;
	bool nonzero = false;
	for (int jj = 0; jj < ops.n; jj++) {
		const int j = ops.n - jj - 1;
		const word_t mask = lpp_bit(ops.site[j]);
		word_t& ketp = (ops.dof[j] == 0) ? k1 : k2;
		const bool bit = (ketp & mask) != 0;
		double result = 1.0;
		bool newbit = bit;
		switch (ops.label[j]) {
		case 0: nonzero = true; break;
		case 1: nonzero = bit; break;
		case 2: result = bit ? -0.5 : 0.5; nonzero = true; break;
		default: newbit = !bit; nonzero = (bit && !ops.transpose[j]) || (!bit && ops.transpose[j]); break;
		}
		if (!nonzero) return false;
		if (newbit != bit) ketp ^= mask;
		if (ops.label[j] == 3) {
			if (ops.dof[j] != 0 && (lpp_popc(k1) & 1)) result = -result;
			result *= (double)lpp_sign_below(ketp, ops.site[j]);
		}
		v *= result;
	}
	if (!nonzero) return false;
	*o1 = k1; *o2 = k2; *value = v;
	return true;
}

// counter-based initial vector, identical to lanczosplusplus_b200/geometry.py::splitmix64_vector
LPP_HD double lpp_splitmix_uniform(uint64_t seed, uint64_t idx)
{
	uint64_t z = idx * 0x9E3779B97F4A7C15ull + seed * 0xD1B54A32D192ED03ull + 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z = z ^ (z >> 31);
	return (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}
