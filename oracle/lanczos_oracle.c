/*
 * oracle/lanczos_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99 + OpenMP) of the hot path of g1257/LanczosPlusPlus:
 * bases, ranks, fermion signs, Hamiltonian row generators, stored-CRS builder,
 * x += H y, the PsimagLite Lanczos recurrence and the continued fraction.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (lanczosplusplus_b200/)
 * never links, imports or calls it.
 *
 * PINNING.  The reference ships no golden vectors (TestSuite/ holds inputs only) and its executables cannot be built here
 * (PsimagLite -- g1257/PsimagLite, un-vendored, un-pinned; README.md:80-82 -- and LAPACK are absent).  Two pins:
 *  (1) MODEL CODE: PINNED.  The reference's own model headers (bases, ranks, signs, row generators, diagonal, stored
 *      assembly, on-the-fly product, getBraIndex/doSignGf) are compiled unmodified against oracle/psimag_shim/ into
 *      oracle/_ref/liblpp_ref.so (ref_bridge.cpp).  This restatement reproduces them exactly: basis words, ranks, rowptr,
 *      colind bit-exact, CRS values and operator applications identical, x += H y to 1e-14 (tests/test_reference_pin.py,
 *      and tests/test_golden.py against the committed outputs under tests/golden/).
 *  (2) PSIMAGLITE PARTS: PARITY UNPINNED.  SparseRow's sort-and-merge, CrsMatrix, LanczosSolver and ContinuedFraction are
 *      external and absent; they are restated from SURVEY App. B and pinned only on analytic answers, numpy/scipy
 *      eigen-solvers applied to the exported CRS and the survey-time cross-check values (tests/test_oracle.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src).  PsimagLite pieces (SparseRow, CrsMatrix, LanczosSolver,
 * ContinuedFraction) are restated from their published behaviour as used at the
 * reference's own call sites (Engine.h:460-490,601-657; HubbardHelper.h:84-133).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t word_t;

enum { ORC_HUBBARD = 0, ORC_FEAS = 1, ORC_HEISENBERG = 2, ORC_TJ = 3 };
enum { ORC_OP_C = 1, ORC_OP_SZ = 2, ORC_OP_CDAGGER = 3, ORC_OP_N = 4, ORC_OP_SPLUS = 5, ORC_OP_SMINUS = 6 }; /* LabeledOperator.h:10-17 */

typedef struct {
	int model;
	int nsite;
	int orbitals;      /* 1 unless FeAs */
	int nup, ndown;    /* Heisenberg: nup = TargetSzPlusConst */
	int u3_all_pairs;  /* FeAs: 1 = stored (Hermitian) definition FeBasedSc.h:192-197, 0 = OTF doTask :85-88 */
	int fast_rank;     /* 0 = faithful rank (loops / linear search), 1 = table lookup (tuned baseline) */
	/* matrices, row-major [a*nb+b] with nb = nsite*orbitals */
	double* hop;       /* Hubbard: hoppings_(i,j); FeAs: geometry(i,o1,j,o2,0); Heisenberg: jpm */
	double* jzz;       /* Heisenberg only */
	double* U;         /* Hubbard: nsite; FeAs: 6 */
	double* V;         /* Hubbard: nsite; FeAs: 2*orbitals*nsite; Heisenberg: magnetic field (nsite) */
	double* Dani;      /* Heisenberg: anisotropy (nsite) ; FeAs: Dani[0]=anisotropyD */
	/* bases */
	size_t n1, n2;     /* sizes of spin-up / spin-down (or single) bases */
	word_t* b1;
	word_t* b2;
	uint64_t* comb;    /* comb table */
	int combn;
	int32_t* lut1;     /* word -> index tables when fast_rank */
	int32_t* lut2;
	/* t-J (TjMultiOrb, Orbitals=1): combined words (down << nsite) | up, sorted; couplings of geometry terms 1 and 3 */
	word_t* tj;
	size_t ntj;
	double* jpm;
	double* w;
} orc_model;

/* ---------------------------------------------------------------- bit utils */
static inline int popc(word_t a) { return __builtin_popcountll(a); }

/* ProgramGlobals.h:109-114 (mask widened to 64 bit; quirk C.2 of SURVEY) */
static inline int do_sign(word_t a, int i)
{
	word_t mask = (((word_t)1) << i) - 1;
	return (popc(a & mask) & 1) ? -1 : 1;
}

/* BasisOneSpin.h:178-191 / BasisOneSpinFeAs.h:333-346 */
static uint64_t* comb_fill(int rows)
{
	uint64_t* c = (uint64_t*)calloc((size_t)rows * rows, sizeof(uint64_t));
	for (int n = 0; n < rows; n++) {
		int m = 0;
		int j = n;
		uint64_t i = 1;
		uint64_t cnm = 1;
		for (; m <= n / 2; m++, cnm = cnm * (uint64_t)j / i, i++, j--)
			c[(size_t)n * rows + m] = c[(size_t)n * rows + (n - m)] = cnm;
	}
	return c;
}

/* BasisOneSpin.h:34-62 and BasisOneSpinFeAs.h:267-295 (same enumeration rule) */
static size_t onespin_size(int nsite, int npart)
{
	size_t hilbert = 1;
	int n = nsite;
	size_t m = 1;
	for (; m <= (size_t)npart; n--, m++)
		hilbert = hilbert * (size_t)n / m;
	return hilbert;
}

static size_t onespin_fill(int nsite, int npart, word_t* data)
{
	size_t hilbert = onespin_size(nsite, npart);
	if (npart == 0) {
		data[0] = 0;
		return 1;
	}
	word_t ket = (((word_t)1) << npart) - 1;
	for (size_t i = 0; i < hilbert; i++) {
		data[i] = ket;
		int n = 0, m = 0;
		for (; (ket & 3) != 1; n++, ket >>= 1)
			m += (int)(ket & 1);
		ket = ((ket + 1) << n) ^ ((((word_t)1) << m) - 1);
	}
	return hilbert;
}

/* BasisOneSpin.h:73-81 */
static size_t onespin_rank(const uint64_t* comb, int combn, word_t state)
{
	size_t n = 0;
	for (size_t b = 0, c = 1; state > 0; b++, state >>= 1)
		if (state & 1) n += comb[b * combn + (c++)];
	return n;
}

size_t orc_onespin_basis(int nsite, int npart, word_t* out)
{
	if (!out) return onespin_size(nsite, npart);
	return onespin_fill(nsite, npart, out);
}

size_t orc_onespin_rank(int nsite, word_t state)
{
	int rows = 2 * nsite + 2;
	uint64_t* c = comb_fill(rows);
	size_t r = onespin_rank(c, rows, state);
	free(c);
	return r;
}

/* ------------------------------------------------------------- FeAs basis */
/* Partitions.h:32-77 : odometer, orbital 0 fastest, keep tuples with sum == length */
static int partitions_make(int length, int parts, int** out)
{
	int cap = 64, count = 0;
	int* res = (int*)malloc(sizeof(int) * cap * parts);
	int* values = (int*)calloc(parts, sizeof(int));
	for (;;) {
		int s = 0;
		for (int i = 0; i < parts; i++) s += values[i];
		if (s == length) {
			if (count == cap) { cap *= 2; res = (int*)realloc(res, sizeof(int) * cap * parts); }
			memcpy(res + (size_t)count * parts, values, sizeof(int) * parts);
			count++;
		}
		values[0]++;
		s = 0;
		for (int i = 0; i < parts; i++) s += values[i];
		if (s > length) {
			/* increaseNextIndices, Partitions.h:62-75 */
			int x = -1;
			if (parts - 1 != 0) {
				values[0] = 0;
				int i = 1;
				for (;;) {
					values[i]++;
					int s2 = 0;
					for (int k = 0; k < parts; k++) s2 += values[k];
					if (s2 <= length) { x = i; break; }
					if (i == parts - 1) { x = -1; break; }
					values[i] = 0;
					i++;
				}
			}
			if (x < 0) break;
		}
	}
	free(values);
	*out = res;
	return count;
}

/* BasisOneSpinFeAs.h:357-374 */
static word_t feas_collate(const word_t* kets, int orbitals)
{
	word_t rem[8];
	for (int o = 0; o < orbitals; o++) rem[o] = kets[o];
	int counter = 0;
	word_t ket = 0;
	for (;;) {
		word_t any = 0;
		for (int o = 0; o < orbitals; o++) any |= rem[o];
		if (!any) break;
		for (int o = 0; o < orbitals; o++) {
			if (rem[o] & 1) ket |= ((word_t)1) << counter;
			counter++;
			if (rem[o]) rem[o] >>= 1;
		}
	}
	return ket;
}

/* BasisOneSpinFeAs.h:45-84,297-331 */
static size_t feas_onespin_fill(int nsite, int npart, int orbitals, const uint64_t* comb, int combn, word_t* data)
{
	if (npart == 0) {
		if (data) data[0] = 0;
		return 1;
	}
	int* parts = NULL;
	int np = partitions_make(npart, orbitals, &parts);
	size_t size = 0;
	for (int p = 0; p < np; p++) {
		size_t tmp = 1;
		for (int o = 0; o < orbitals; o++) tmp *= comb[(size_t)nsite * combn + parts[p * orbitals + o]];
		size += tmp;
	}
	if (!data) { free(parts); return size; }
	size_t counter = 0;
	for (int p = 0; p < np; p++) {
		const int* na = parts + (size_t)p * orbitals;
		word_t* basisA[8];
		size_t sizesA[8];
		int skip = 0;
		for (int o = 0; o < orbitals; o++) {
			if (na[o] > nsite) skip = 1;
		}
		if (skip) continue; /* comb_(nsite,n)=0 => fillPartialBasis makes an empty list, total = 0 */
		for (int o = 0; o < orbitals; o++) {
			sizesA[o] = onespin_size(nsite, na[o]);
			basisA[o] = (word_t*)malloc(sizeof(word_t) * sizesA[o]);
			onespin_fill(nsite, na[o], basisA[o]);
		}
		size_t total = 1;
		for (int o = 0; o < orbitals; o++) total *= sizesA[o];
		for (size_t i = 0; i < total; i++) {
			/* getKets, BasisOneSpinFeAs.h:313-331 (literal, including `tmp = ind % sizes`) */
			word_t kets[8];
			size_t tmp = i;
			size_t sizes = 1;
			for (int o = 0; o < orbitals - 1; o++) sizes *= sizesA[o];
			for (int o = 1; o < orbitals; o++) {
				size_t ix = tmp / sizes;
				tmp = i % sizes;
				kets[orbitals - o] = basisA[orbitals - o][ix];
				sizes /= sizesA[orbitals - o - 1];
			}
			kets[0] = basisA[0][tmp];
			data[counter++] = feas_collate(kets, orbitals);
		}
		for (int o = 0; o < orbitals; o++) free(basisA[o]);
	}
	free(parts);
	return counter;
}

/* BasisOneSpinFeAs.h:431-442 */
static int feas_nbyket(word_t ket, int from, int upto)
{
	int sum = 0;
	for (int c = from; c < upto; c++)
		if (ket & (((word_t)1) << c)) sum++;
	return sum;
}

/* BasisOneSpinFeAs.h:252-263 */
static int feas_dosign_onsite(word_t ket, int i, int orb1, int orb2, int orbitals)
{
	if (orb1 > orb2) return -feas_dosign_onsite(ket, i, orb2, orb1, orbitals);
	int x0 = i * orbitals + orb1;
	int x1 = i * orbitals + orb2;
	int sum = feas_nbyket(ket, x0, x1);
	return (sum & 1) ? -1 : 1;
}

/* BasisOneSpinFeAs.h:150-181 */
static int feas_dosign(word_t ket, int i, int orb1, int j, int orb2, int orbitals)
{
	if (i == j) return feas_dosign_onsite(ket, i, orb1, orb2, orbitals);
	int x0 = (i + 1) * orbitals;
	int x1 = j * orbitals;
	int sum = feas_nbyket(ket, x0, x1);
	x0 = i * orbitals + orb1;
	x1 = (i + 1) * orbitals;
	sum += feas_nbyket(ket, x0, x1);
	x0 = j * orbitals;
	x1 = j * orbitals + orb2;
	sum += feas_nbyket(ket, x0, x1);
	return (sum & 1) ? -1 : 1;
}

/* ------------------------------------------------------------ model object */
static int32_t* make_lut(const word_t* b, size_t n, int nbits)
{
	size_t sz = ((size_t)1) << nbits;
	int32_t* lut = (int32_t*)malloc(sizeof(int32_t) * sz);
	for (size_t i = 0; i < sz; i++) lut[i] = -1;
	for (size_t i = 0; i < n; i++) lut[b[i]] = (int32_t)i;
	return lut;
}

orc_model* orc_create(int model, int nsite, int orbitals, int nup, int ndown,
                      const double* hop, const double* jzz, const double* U, int nU,
                      const double* V, int nV, const double* Dani, int nD,
                      int u3_all_pairs, int fast_rank)
{
	orc_model* m = (orc_model*)calloc(1, sizeof(orc_model));
	m->model = model;
	m->nsite = nsite;
	m->orbitals = (model == ORC_FEAS) ? orbitals : 1;
	m->nup = nup;
	m->ndown = ndown;
	m->u3_all_pairs = u3_all_pairs;
	m->fast_rank = fast_rank;
	int nb = nsite * m->orbitals;
	m->hop = (double*)calloc((size_t)nb * nb, sizeof(double));
	if (hop) memcpy(m->hop, hop, sizeof(double) * nb * nb);
	m->jzz = (double*)calloc((size_t)nb * nb, sizeof(double));
	if (jzz) memcpy(m->jzz, jzz, sizeof(double) * nb * nb);
	int needU = (model == ORC_FEAS) ? 6 : nsite;
	m->U = (double*)calloc(needU, sizeof(double));
	if (U) memcpy(m->U, U, sizeof(double) * (nU < needU ? nU : needU));
	if (model == ORC_FEAS && (nU == 4 || nU == 5)) { /* ParametersModelFeAs.h:147-151 */
		m->U[4] = m->U[2];
		m->U[5] = 0.0;
	}
	int needV = (model == ORC_FEAS) ? 2 * m->orbitals * nsite : nsite;
	m->V = (double*)calloc(needV, sizeof(double));
	if (V) memcpy(m->V, V, sizeof(double) * (nV < needV ? nV : needV));
	m->Dani = (double*)calloc(nsite > 1 ? nsite : 1, sizeof(double));
	if (Dani) memcpy(m->Dani, Dani, sizeof(double) * (nD < nsite ? nD : nsite));

	if (model == ORC_HUBBARD) {
		m->combn = 2 * nsite + 2;
		m->comb = comb_fill(m->combn);
		m->n1 = onespin_size(nsite, nup);
		m->n2 = onespin_size(nsite, ndown);
		m->b1 = (word_t*)malloc(sizeof(word_t) * m->n1);
		m->b2 = (word_t*)malloc(sizeof(word_t) * m->n2);
		onespin_fill(nsite, nup, m->b1);
		onespin_fill(nsite, ndown, m->b2);
	} else if (model == ORC_FEAS) {
		m->combn = m->orbitals * nsite + 1;
		m->comb = comb_fill(m->combn);
		m->n1 = feas_onespin_fill(nsite, nup, m->orbitals, m->comb, m->combn, NULL);
		m->n2 = feas_onespin_fill(nsite, ndown, m->orbitals, m->comb, m->combn, NULL);
		m->b1 = (word_t*)malloc(sizeof(word_t) * m->n1);
		m->b2 = (word_t*)malloc(sizeof(word_t) * m->n2);
		feas_onespin_fill(nsite, nup, m->orbitals, m->comb, m->combn, m->b1);
		feas_onespin_fill(nsite, ndown, m->orbitals, m->comb, m->combn, m->b2);
	} else { /* BasisHeisenberg.h:24-47, S=1/2: bits_=1, ascending scan of all words */
		m->combn = 2 * nsite + 2;
		m->comb = comb_fill(m->combn);
		size_t cap = 1024, cnt = 0;
		m->b1 = (word_t*)malloc(sizeof(word_t) * cap);
		word_t total = ((word_t)1) << nsite;
		for (word_t lui = 0; lui < total; ++lui) {
			if (popc(lui) != nup) continue;
			if (cnt == cap) { cap *= 2; m->b1 = (word_t*)realloc(m->b1, sizeof(word_t) * cap); }
			m->b1[cnt++] = lui;
		}
		m->n1 = cnt;
		m->n2 = 1;
		m->b2 = (word_t*)calloc(1, sizeof(word_t));
	}
	if (fast_rank) {
		m->lut1 = make_lut(m->b1, m->n1, nb);
		if (model != ORC_HEISENBERG) m->lut2 = make_lut(m->b2, m->n2, nb);
	}
	return m;
}

void orc_destroy(orc_model* m)
{
	if (!m) return;
	free(m->hop); free(m->jzz); free(m->U); free(m->V); free(m->Dani);
	free(m->b1); free(m->b2); free(m->comb); free(m->lut1); free(m->lut2);
	free(m->tj); free(m->jpm); free(m->w);
	free(m);
}

size_t orc_rows(const orc_model* m)
{
	if (m->model == ORC_TJ) return m->ntj;
	return (m->model == ORC_HEISENBERG) ? m->n1 : m->n1 * m->n2;
}

/* ---------------------------------------------------------------- t-J basis
 * BasisTjMultiOrbLanczos.h:30-42: fillOneSector for both species (:321-349, the BasisOneSpin ripple), combineAndFilter
 * (:351-366: every pair without a doubly occupied site, combined as (down << n) | up), std::sort. */
static int cmp_word(const void* a, const void* b)
{
	word_t x = *(const word_t*)a, y = *(const word_t*)b;
	return (x > y) - (x < y);
}

orc_model* orc_create_tj(int nsite, int nup, int ndown, const double* hop, const double* jpm, const double* jzz,
                         const double* w, const double* V, int nV)
{
	orc_model* m = (orc_model*)calloc(1, sizeof(orc_model));
	m->model = ORC_TJ;
	m->nsite = nsite;
	m->orbitals = 1;
	m->nup = nup;
	m->ndown = ndown;
	size_t nn = (size_t)nsite * nsite;
	m->hop = (double*)calloc(nn, sizeof(double));
	m->jpm = (double*)calloc(nn, sizeof(double));
	m->jzz = (double*)calloc(nn, sizeof(double));
	m->w = (double*)calloc(nn, sizeof(double));
	if (hop) memcpy(m->hop, hop, sizeof(double) * nn);
	if (jpm) memcpy(m->jpm, jpm, sizeof(double) * nn);
	if (jzz) memcpy(m->jzz, jzz, sizeof(double) * nn);
	if (w) memcpy(m->w, w, sizeof(double) * nn);
	m->V = (double*)calloc(2 * (size_t)nsite, sizeof(double));
	if (V) memcpy(m->V, V, sizeof(double) * (size_t)(nV < 2 * nsite ? nV : 2 * nsite));
	m->n1 = onespin_size(nsite, nup);
	m->n2 = onespin_size(nsite, ndown);
	m->b1 = (word_t*)malloc(sizeof(word_t) * m->n1);
	m->b2 = (word_t*)malloc(sizeof(word_t) * m->n2);
	onespin_fill(nsite, nup, m->b1);
	onespin_fill(nsite, ndown, m->b2);
	size_t cap = 1024, cnt = 0;
	m->tj = (word_t*)malloc(sizeof(word_t) * cap);
	for (size_t i = 0; i < m->n1; i++)
		for (size_t j = 0; j < m->n2; j++) {
			if (m->b1[i] & m->b2[j]) continue;               /* one or more doubly occupied sites */
			if (cnt == cap) { cap *= 2; m->tj = (word_t*)realloc(m->tj, sizeof(word_t) * cap); }
			m->tj[cnt++] = (m->b2[j] << nsite) | m->b1[i];
		}
	qsort(m->tj, cnt, sizeof(word_t), cmp_word);
	m->ntj = cnt;
	return m;
}

/* BasisTjMultiOrbLanczos.h:71-112: bisection (with a linear fallback) over the sorted combined words */
static size_t tj_perfect_index(const orc_model* m, word_t k1, word_t k2)
{
	word_t key = (k2 << m->nsite) | k1;
	size_t lo = 0, hi = m->ntj;
	while (lo < hi) {
		size_t mid = lo + (hi - lo) / 2;
		if (m->tj[mid] < key) lo = mid + 1;
		else hi = mid;
	}
	if (lo >= m->ntj || m->tj[lo] != key) {
		fprintf(stderr, "orc: t-J perfectIndex: state not in the basis\n");
		abort();
	}
	return lo;
}

/* basis(i, spin) for every row (BasisBase::operator(); BasisTjMultiOrbLanczos.h:127-141) */
static inline void row_kets(const orc_model* m, size_t r, word_t* k1, word_t* k2);
void orc_row_words(const orc_model* m, int spin, word_t* out)
{
	size_t n = orc_rows(m);
	for (size_t r = 0; r < n; r++) {
		word_t k1, k2;
		row_kets(m, r, &k1, &k2);
		out[r] = spin == 0 ? k1 : k2;
	}
}

size_t orc_basis_size(const orc_model* m, int spin) { return spin == 0 ? m->n1 : m->n2; }

void orc_basis_words(const orc_model* m, int spin, word_t* out)
{
	if (spin == 0) memcpy(out, m->b1, sizeof(word_t) * m->n1);
	else memcpy(out, m->b2, sizeof(word_t) * m->n2);
}

/* one-spin rank: BasisOneSpin.h:73-81 ; BasisOneSpinFeAs.h:96-101 ; BasisHeisenberg.h:73-80 */
static size_t rank1(const orc_model* m, int spin, word_t w)
{
	const word_t* b = spin == 0 ? m->b1 : m->b2;
	size_t n = spin == 0 ? m->n1 : m->n2;
	if (m->fast_rank) return (size_t)((spin == 0 ? m->lut1 : m->lut2)[w]);
	if (m->model == ORC_HUBBARD) return onespin_rank(m->comb, m->combn, w);
	for (size_t i = 0; i < n; i++) /* linear search, as the reference does */
		if (b[i] == w) return i;
	fprintf(stderr, "orc: perfectIndex: no index found\n");
	abort();
}

size_t orc_rank(const orc_model* m, int spin, word_t w) { return rank1(m, spin, w); }

/* BasisHubbardLanczos.h:59-63 ; BasisFeAsBasedSc.h:97-100 */
static inline size_t perfect_index(const orc_model* m, word_t k1, word_t k2)
{
	if (m->model == ORC_TJ) return tj_perfect_index(m, k1, k2);
	if (m->model == ORC_HEISENBERG) return rank1(m, 0, k1);
	return rank1(m, 0, k1) + rank1(m, 1, k2) * m->n1;
}

size_t orc_perfect_index(const orc_model* m, word_t k1, word_t k2) { return perfect_index(m, k1, k2); }

/* --------------------------------------------------------------- sparse row */
typedef struct {
	size_t* cols;
	double* vals;
	int n, cap;
} srow;

static void srow_add(srow* r, size_t col, double v)
{
	if (r->n == r->cap) {
		r->cap = r->cap ? 2 * r->cap : 16;
		r->cols = (size_t*)realloc(r->cols, sizeof(size_t) * r->cap);
		r->vals = (double*)realloc(r->vals, sizeof(double) * r->cap);
	}
	r->cols[r->n] = col;
	r->vals[r->n] = v;
	r->n++;
}

/* PsimagLite SparseRow::finalize(y): plain sum in insertion order */
static double srow_dot(const srow* r, const double* y)
{
	double s = 0;
	for (int i = 0; i < r->n; i++) s += r->vals[i] * y[r->cols[i]];
	return s;
}

/* PsimagLite SparseRow::finalize(CrsMatrix&): stable sort by column, merge duplicates, keep zeros */
static int srow_compress(srow* r)
{
	for (int i = 1; i < r->n; i++) {
		size_t c = r->cols[i];
		double v = r->vals[i];
		int j = i - 1;
		while (j >= 0 && r->cols[j] > c) {
			r->cols[j + 1] = r->cols[j];
			r->vals[j + 1] = r->vals[j];
			j--;
		}
		r->cols[j + 1] = c;
		r->vals[j + 1] = v;
	}
	if (r->n == 0) return 0;
	int k = 0;
	for (int i = 1; i < r->n; i++) {
		if (r->cols[i] == r->cols[k]) r->vals[k] += r->vals[i];
		else { k++; r->cols[k] = r->cols[i]; r->vals[k] = r->vals[i]; }
	}
	r->n = k + 1;
	return r->n;
}

/* ------------------------------------------------------------ Hubbard rows */
/* HubbardHelper.h:138-189 (plain HubbardOneBand: no J, no Coulomb, no potentialT) */
static double hubbard_diag(const orc_model* m, word_t ket1, word_t ket2)
{
	double s = 0;
	for (int i = 0; i < m->nsite; i++) {
		word_t mask = ((word_t)1) << i;
		int nu = (ket1 & mask) ? 1 : 0;
		int nd = (ket2 & mask) ? 1 : 0;
		s += m->U[i] * nu * nd;
		double ne = nu + nd;
		double tmp = m->V[i];
		if (tmp != 0) s += tmp * ne;
	}
	return s;
}

/* HubbardHelper.h:191-243 */
static void hubbard_hops(const orc_model* m, srow* row, word_t ket1, word_t ket2, int i)
{
	const int nsite = m->nsite;
	word_t mi = ((word_t)1) << i;
	int s1i = (ket1 & mi) ? 1 : 0;
	int s2i = (ket2 & mi) ? 1 : 0;
	for (int j = 0; j < nsite; ++j) {
		double h = m->hop[(size_t)i * nsite + j];
		int hasHop = (h != 0);
		word_t mj = ((word_t)1) << j;
		int s1j = (ket1 & mj) ? 1 : 0;
		int s2j = (ket2 & mj) ? 1 : 0;
		if (hasHop && s1i == 1 && s1j == 0) {
			word_t bra1 = ket1 ^ mi;
			double tmp2 = do_sign(ket1, i) * do_sign(bra1, j);
			bra1 = bra1 ^ mj;
			size_t temp = perfect_index(m, bra1, ket2);
			srow_add(row, temp, h * tmp2);
		}
		if (hasHop && s2i == 1 && s2j == 0) {
			word_t bra2 = ket2 ^ mi;
			double tmp2 = do_sign(ket2, i) * do_sign(bra2, j);
			bra2 = bra2 ^ mj;
			size_t temp = perfect_index(m, ket1, bra2);
			srow_add(row, temp, h * tmp2);
		}
	}
}

/* --------------------------------------------------------------- FeAs rows */
static inline int feas_occ(word_t ket, int site, int orb, int orbitals)
{
	return (ket & (((word_t)1) << (site * orbitals + orb))) ? 1 : 0;
}

/* FeBasedSc.h:725-736 */
static inline double feas_sz(word_t k1, word_t k2, int i, int orb, int orbitals)
{
	double sz = feas_occ(k1, i, orb, orbitals);
	sz -= feas_occ(k2, i, orb, orbitals);
	return 0.5 * sz;
}

/* FeBasedSc.h:534-571 + findSnoDecay :573-623 (INT_PAPER33; one geometry term => no J_zz; no SpinOrbit) */
static double feas_diag(const orc_model* m, word_t ket1, word_t ket2)
{
	const int nsite = m->nsite, no = m->orbitals;
	double s = 0;
	for (int i = 0; i < nsite; i++) {
		double szOrb = 0;
		for (int orb = 0; orb < no; orb++) {
			double t = m->U[0] * feas_occ(ket1, i, orb, no) * feas_occ(ket2, i, orb, no);
			for (int orb2 = orb + 1; orb2 < no; orb2++) {
				int nix1 = feas_occ(ket1, i, orb, no) + feas_occ(ket2, i, orb, no);
				int nix2 = feas_occ(ket1, i, orb2, no) + feas_occ(ket2, i, orb2, no);
				t += m->U[1] * nix1 * nix2;
				t += m->U[4] * feas_sz(ket1, ket2, i, orb, no) * feas_sz(ket1, ket2, i, orb2, no);
				t += m->U[5] * feas_occ(ket1, i, orb, no) * feas_occ(ket1, i, orb2, no);
				t += m->U[5] * feas_occ(ket2, i, orb, no) * feas_occ(ket2, i, orb2, no);
			}
			s += t;
			s += m->V[i + (orb + no * 0) * nsite] * feas_occ(ket1, i, orb, no) +
			     m->V[i + (orb + no * 1) * nsite] * feas_occ(ket2, i, orb, no);
			szOrb += feas_sz(ket1, ket2, i, orb, no);
		}
		s += m->Dani[0] * szOrb * szOrb;
	}
	return s;
}

/* FeBasedSc.h:320-374 */
static void feas_hops(const orc_model* m, srow* row, word_t ket1, word_t ket2, int i, int orb)
{
	const int nsite = m->nsite, no = m->orbitals, nb = nsite * no;
	int ii = i * no + orb;
	word_t mii = ((word_t)1) << ii;
	int s1i = (ket1 & mii) ? 1 : 0;
	int s2i = (ket2 & mii) ? 1 : 0;
	for (int j = 0; j < nsite; j++) {
		if (j < i) continue;
		for (int orb2 = 0; orb2 < no; orb2++) {
			int jj = j * no + orb2;
			double h = -m->hop[(size_t)ii * nb + jj]; /* hoppings(), :320-323 */
			if (h == 0) continue;
			word_t mjj = ((word_t)1) << jj;
			int s1j = (ket1 & mjj) ? 1 : 0;
			int s2j = (ket2 & mjj) ? 1 : 0;
			if (s1i + s1j == 1) {
				word_t bra1 = ket1 ^ (mii | mjj);
				size_t temp = perfect_index(m, bra1, ket2);
				double extraSign = (s1i == 1) ? -1 : 1;
				double tmp2 = feas_dosign(ket1, i, orb, j, orb2, no);
				srow_add(row, temp, h * extraSign * tmp2);
			}
			if (s2i + s2j == 1) {
				word_t bra2 = ket2 ^ (mii | mjj);
				size_t temp = perfect_index(m, ket1, bra2);
				double extraSign = (s2i == 1) ? -1 : 1;
				double tmp2 = feas_dosign(ket2, i, orb, j, orb2, no);
				srow_add(row, temp, h * extraSign * tmp2);
			}
		}
	}
}

/* FeBasedSc.h:503-518 (i<=j callers only) */
static int feas_jterm_sign(word_t k1, word_t k2, int i, int orb1, int j, int orb2, int no)
{
	if (i > j) return feas_jterm_sign(k1, k2, j, orb2, i, orb1, no);
	int x = feas_dosign(k1, i, orb1, j, orb2, no);
	x *= feas_dosign(k2, i, orb1, j, orb2, no);
	return x;
}

/* FeBasedSc.h:376-411 with splusSminusNonZero :678-695 */
static void feas_u2(const orc_model* m, srow* row, word_t ket1, word_t ket2, int i, int orb1)
{
	const int no = m->orbitals;
	double val = m->U[2] * 0.5;
	for (int orb2 = 0; orb2 < no; orb2++) {
		if (orb1 == orb2) continue;
		double sign = feas_jterm_sign(ket1, ket2, i, orb1, i, orb2, no);
		/* setSplusSminus(i,orb1,j=i,orb2) */
		if (feas_occ(ket1, i, orb2, no) == 0) continue;
		if (feas_occ(ket1, i, orb1, no) == 1) continue;
		if (feas_occ(ket2, i, orb1, no) == 0) continue;
		if (feas_occ(ket2, i, orb2, no) == 1) continue;
		word_t mk = (((word_t)1) << (i * no + orb1)) | (((word_t)1) << (i * no + orb2));
		size_t temp = perfect_index(m, ket1 ^ mk, ket2 ^ mk);
		srow_add(row, temp, val * sign);
	}
}

/* FeBasedSc.h:414-432 with u3TermNonZero :697-713 */
static void feas_u3(const orc_model* m, srow* row, word_t ket1, word_t ket2, int i, int orb1, int orb2)
{
	const int no = m->orbitals;
	if (feas_occ(ket1, i, orb2, no) == 0) return;
	if (feas_occ(ket1, i, orb1, no) == 1) return;
	if (feas_occ(ket2, i, orb1, no) == 1) return;
	if (feas_occ(ket2, i, orb2, no) == 0) return;
	word_t mk = (((word_t)1) << (i * no + orb1)) | (((word_t)1) << (i * no + orb2));
	size_t temp = perfect_index(m, ket1 ^ mk, ket2 ^ mk);
	double sign = feas_jterm_sign(ket1, ket2, i, orb1, i, orb2, no);
	srow_add(row, temp, -1.0 * m->U[3] * sign);
}

/* --------------------------------------------------------- Heisenberg rows */
/* Heisenberg.h:242-276, S=1/2 */
static double heis_diag(const orc_model* m, word_t ket)
{
	const int nsite = m->nsite;
	double s = 0;
	for (int i = 0; i < nsite; i++) {
		int val1 = (int)((ket >> i) & 1);
		double tmp1 = val1 - 0.5;
		double tmp1d = tmp1 * tmp1;
		s += m->V[i] * tmp1;      /* MagneticField */
		s += m->Dani[i] * tmp1d;  /* AnisotropyD */
		for (int j = i + 1; j < nsite; j++) {
			int val2 = (int)((ket >> j) & 1);
			double tmp2 = val2 - 0.5;
			s += tmp1 * tmp2 * m->jzz[(size_t)i * nsite + j];
		}
	}
	return s;
}

/* Heisenberg.h:94-106 + setSplusSminus :278-307 (S=1/2: both sqrt factors are 1) */
static void heis_offdiag(const orc_model* m, srow* row, word_t ket)
{
	const int nsite = m->nsite;
	const double spin = 0.5;
	for (int i = 0; i < nsite; i++) {
		int val1 = (int)((ket >> i) & 1);
		if (val1 == 1) continue; /* val1 == twiceTheSpin */
		for (int j = 0; j < nsite; j++) {
			if (i == j) continue;
			double jpm = m->hop[(size_t)i * nsite + j];
			if (jpm == 0) continue;
			int val2 = (int)((ket >> j) & 1);
			if (val2 == 0) continue;
			double m2 = val2 - spin;
			double m1 = (val2 - 1) - spin;
			word_t bra = (ket | (((word_t)1) << i)) & ~(((word_t)1) << j);
			size_t temp = rank1(m, 0, bra);
			double tmp = sqrt(spin * (spin + 1.0) - m1 * (m1 + 1.0));
			tmp *= sqrt(spin * (spin + 1.0) - m2 * (m2 - 1.0));
			srow_add(row, temp, 0.5 * tmp * jpm);
		}
	}
}

/* ------------------------------------------------------------ generic rows */
/* ---------------------------------------------------------------- t-J rows (TjMultiOrb.h, Orbitals=1) */
/* BasisTjMultiOrbLanczos.h:378-396: parity of the occupied sites in [i, j) */
static int tj_dosign(word_t ket, int i, int j)
{
	int sum = 0;
	for (int c = i + 1; c < j; c++) if (ket & (((word_t)1) << c)) sum++;
	for (int c = i; c < i + 1; c++) if (ket & (((word_t)1) << c)) sum++;
	return (sum & 1) ? -1 : 1;
}

/* TjMultiOrb.h:784-801: parity of the occupied sites in [i, j], both ends included */
static int tj_parity_from(int i, int j, word_t ket)
{
	if (i == j) return (ket & (((word_t)1) << j)) ? -1 : 1;
	word_t mask = ket & (((((word_t)1) << (i + 1)) - 1) ^ ((((word_t)1) << j) - 1));
	int s = (popc(mask) & 1) ? -1 : 1;
	if (ket & (((word_t)1) << i)) s = -s;
	if (ket & (((word_t)1) << j)) s = -s;
	return s;
}

/* TjMultiOrb.h:773-782 */
static double tj_sign_spsm(int i, int j, word_t bra1, word_t bra2)
{
	int s = 1;
	if (j > 0) s *= tj_parity_from(0, j - 1, bra2);
	if (i > 0) s *= tj_parity_from(0, i - 1, bra2);
	if (i > 0) s *= tj_parity_from(0, i - 1, bra1);
	if (j > 0) s *= tj_parity_from(0, j - 1, bra1);
	return (double)s;
}

/* TjMultiOrb.h:586-647, one orbital (proij = 1) */
static double tj_diag(const orc_model* m, word_t ket1, word_t ket2)
{
	int nsite = m->nsite;
	double s = 0;
	for (int i = 0; i < nsite; i++) {
		int niup = (int)((ket1 >> i) & 1), nidown = (int)((ket2 >> i) & 1);
		s += m->V[i] * niup;
		s += m->V[i + nsite] * nidown;
		for (int j = i + 1; j < nsite; j++) {
			int njup = (int)((ket1 >> j) & 1), njdown = (int)((ket2 >> j) & 1);
			s += (niup - nidown) * (njup - njdown) * m->jzz[i * nsite + j] * 0.25;
			s += (niup + nidown) * (njup + njdown) * m->w[i * nsite + j];
		}
	}
	return s;
}

/* TjMultiOrb.h:649-695 */
static void tj_hops(const orc_model* m, srow* row, word_t ket1, word_t ket2, int i)
{
	int nsite = m->nsite;
	word_t mi = ((word_t)1) << i;
	int s1i = (ket1 & mi) ? 1 : 0, s2i = (ket2 & mi) ? 1 : 0;
	for (int j = 0; j < nsite; j++) {
		if (j < i) continue;
		double h = m->hop[i * nsite + j];
		if (h == 0) continue;
		word_t mj = ((word_t)1) << j;
		int s1j = (ket1 & mj) ? 1 : 0, s2j = (ket2 & mj) ? 1 : 0;
		if (s1i + s1j == 1 && !(s1j == 0 && s2j > 0) && !(s1j > 0 && s2i > 0)) {
			word_t bra1 = ket1 ^ (mi | mj);
			double extraSign = (s1i == 1) ? -1 : 1;
			srow_add(row, perfect_index(m, bra1, ket2), h * extraSign * tj_dosign(ket1, i, j));
		}
		if (s2i + s2j == 1 && !(s2j == 0 && s1j > 0) && !(s2j > 0 && s1i > 0)) {
			word_t bra2 = ket2 ^ (mi | mj);
			double extraSign = (s2i == 1) ? -1 : 1;
			srow_add(row, perfect_index(m, ket1, bra2), h * extraSign * tj_dosign(ket2, i, j));
		}
	}
}

/* TjMultiOrb.h:697-771 */
static void tj_spsm(const orc_model* m, srow* row, word_t ket1, word_t ket2, int i)
{
	int nsite = m->nsite;
	word_t mi = ((word_t)1) << i;
	int s1i = (ket1 & mi) ? 1 : 0, s2i = (ket2 & mi) ? 1 : 0;
	for (int j = 0; j < nsite; j++) {
		if (j < i) continue;
		double h = m->jpm[i * nsite + j] * 0.5;
		if (h == 0) continue;
		word_t mj = ((word_t)1) << j;
		int s1j = (ket1 & mj) ? 1 : 0, s2j = (ket2 & mj) ? 1 : 0;
		if (s1i == 1 && s1j == 0 && s2i == 0 && s2j == 1) {
			word_t bra1 = (ket1 ^ mi) | mj;
			word_t bra2 = (ket2 | mi) ^ mj;
			srow_add(row, perfect_index(m, bra1, bra2), h * tj_sign_spsm(i, j, bra1, bra2));
		}
		if (s1i == 0 && s1j == 1 && s2i == 1 && s2j == 0) {
			word_t bra1 = (ket1 | mi) ^ mj;
			word_t bra2 = (ket2 ^ mi) | mj;
			srow_add(row, perfect_index(m, bra1, bra2), h * tj_sign_spsm(i, j, bra1, bra2));
		}
	}
}

static inline void row_kets(const orc_model* m, size_t r, word_t* k1, word_t* k2)
{
	if (m->model == ORC_HEISENBERG) { *k1 = m->b1[r]; *k2 = 0; return; }
	if (m->model == ORC_TJ) { /* BasisTjMultiOrbLanczos.h:127-141 */
		*k1 = m->tj[r] & ((((word_t)1) << m->nsite) - 1);
		*k2 = m->tj[r] >> m->nsite;
		return;
	}
	/* BasisHubbardLanczos.h:77-84 ; BasisFeAsBasedSc.h:84-89 */
	*k1 = m->b1[r % m->n1];
	*k2 = m->b2[r / m->n1];
}

static double row_diag(const orc_model* m, word_t k1, word_t k2)
{
	if (m->model == ORC_HUBBARD) return hubbard_diag(m, k1, k2);
	if (m->model == ORC_FEAS) return feas_diag(m, k1, k2);
	if (m->model == ORC_TJ) return tj_diag(m, k1, k2);
	return heis_diag(m, k1);
}

/* off-diagonal emission order of setupHamiltonian (stored=1) or matrixVectorProduct (stored=0):
 * HubbardHelper.h:95-98,122-126 ; FeBasedSc.h:185-215 vs :76-103 ; Heisenberg.h:101-106 */
static void row_offdiag(const orc_model* m, srow* row, word_t k1, word_t k2, int stored)
{
	if (m->model == ORC_HUBBARD) {
		for (int i = 0; i < m->nsite; i++) hubbard_hops(m, row, k1, k2, i);
	} else if (m->model == ORC_FEAS) {
		int allpairs = stored ? 1 : m->u3_all_pairs;
		for (int i = 0; i < m->nsite; i++) {
			for (int orb = 0; orb < m->orbitals; orb++) {
				feas_hops(m, row, k1, k2, i, orb);
				feas_u2(m, row, k1, k2, i, orb);
				for (int orb2 = allpairs ? 0 : orb + 1; orb2 < m->orbitals; orb2++) {
					if (orb == orb2) continue;
					feas_u3(m, row, k1, k2, i, orb, orb2);
				}
			}
		}
	} else if (m->model == ORC_TJ) { /* TjMultiOrb.h:112-117 (one orbital) */
		for (int i = 0; i < m->nsite; i++) {
			tj_hops(m, row, k1, k2, i);
			tj_spsm(m, row, k1, k2, i);
		}
	} else {
		heis_offdiag(m, row, k1);
	}
}

/* raw (uncompressed) row in reference emission order; diag first when stored */
int orc_row(const orc_model* m, size_t r, int stored, size_t* cols, double* vals, int cap)
{
	srow row = {0};
	word_t k1, k2;
	row_kets(m, r, &k1, &k2);
	if (stored) srow_add(&row, r, row_diag(m, k1, k2));
	row_offdiag(m, &row, k1, k2, stored);
	int n = row.n;
	for (int i = 0; i < n && i < cap; i++) { cols[i] = row.cols[i]; vals[i] = row.vals[i]; }
	free(row.cols); free(row.vals);
	return n;
}

double orc_diag(const orc_model* m, size_t r)
{
	word_t k1, k2;
	row_kets(m, r, &k1, &k2);
	return row_diag(m, k1, k2);
}

/* setupHamiltonian: HubbardHelper.h:75-103 ; FeBasedSc.h:163-221 ; Heisenberg.h:80-114.
 * pass rowptr==NULL to get nnz only. rowptr has rows+1 entries (int64), colind int64. */
int64_t orc_crs_build(const orc_model* m, int64_t* rowptr, int64_t* colind, double* values)
{
	size_t hilbert = orc_rows(m);
	int64_t nCounter = 0;
	for (size_t r = 0; r < hilbert; r++) {
		srow row = {0};
		word_t k1, k2;
		row_kets(m, r, &k1, &k2);
		if (rowptr) rowptr[r] = nCounter;
		srow_add(&row, r, row_diag(m, k1, k2));
		row_offdiag(m, &row, k1, k2, 1);
		int n = srow_compress(&row);
		if (colind)
			for (int i = 0; i < n; i++) { colind[nCounter + i] = (int64_t)row.cols[i]; values[nCounter + i] = row.vals[i]; }
		nCounter += n;
		free(row.cols); free(row.vals);
	}
	if (rowptr) rowptr[hilbert] = nCounter;
	return nCounter;
}

/* PsimagLite CrsMatrix::matrixVectorProduct: x += A y */
void orc_crs_matvec(size_t rows, const int64_t* rowptr, const int64_t* colind, const double* values,
                    double* x, const double* y)
{
#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t)rows; i++) {
		double s = 0;
		for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++) s += values[k] * y[colind[k]];
		x[i] += s;
	}
}

/* x += H y on the fly, rows [r0, r1); x is indexed from r0 (x[r - r0]), y is the full vector.
 * faithful=1: HubbardHelper.h:105-134 (serial diagonal pass recomputed per call, per-row heap SparseRow),
 *             FeBasedSc.h:66-105 (diag inline per row).  Heisenberg has no OTF in the reference
 *             (ModelBase.h:73-79 throws); its semantics are those of the stored builder.
 * faithful=0: same arithmetic, diagonal inline and row buffer reused (tuned CPU baseline). */
void orc_matvec_range(const orc_model* m, double* x, const double* y, int faithful, int64_t r0, int64_t r1)
{
	if (faithful && m->model == ORC_HUBBARD) {
		double* diag = (double*)malloc(sizeof(double) * (size_t)(r1 - r0));
		for (int64_t r = r0; r < r1; r++) {
			word_t k1, k2;
			row_kets(m, r, &k1, &k2);
			diag[r - r0] = hubbard_diag(m, k1, k2);
		}
		for (int64_t r = r0; r < r1; r++) x[r - r0] += diag[r - r0] * y[r];
		free(diag);
#pragma omp parallel for schedule(static)
		for (int64_t r = r0; r < r1; r++) {
			srow row = {0};
			word_t k1, k2;
			row_kets(m, r, &k1, &k2);
			row_offdiag(m, &row, k1, k2, 0);
			x[r - r0] += srow_dot(&row, y);
			free(row.cols); free(row.vals);
		}
		return;
	}
#pragma omp parallel
	{
		srow row = {0};
#pragma omp for schedule(static)
		for (int64_t r = r0; r < r1; r++) {
			word_t k1, k2;
			row_kets(m, r, &k1, &k2);
			row.n = 0;
			x[r - r0] += row_diag(m, k1, k2) * y[r];
			row_offdiag(m, &row, k1, k2, 0);
			x[r - r0] += srow_dot(&row, y);
			if (faithful) { free(row.cols); free(row.vals); row.cols = NULL; row.vals = NULL; row.cap = 0; }
		}
		free(row.cols); free(row.vals);
	}
}

void orc_matvec(const orc_model* m, double* x, const double* y, int faithful)
{
	orc_matvec_range(m, x, y, faithful, 0, (int64_t)orc_rows(m));
}

/* the three PsimagLite vector sweeps of one Lanczos step on n elements; used by the timed
 * CPU baseline on a bounded sample. returns b. */
double orc_lanczos_sweeps(double* x, double* y, int64_t n)
{
	double a = 0;
#pragma omp parallel for reduction(+ : a) schedule(static)
	for (int64_t i = 0; i < n; i++) a += y[i] * x[i];
	double b = 0;
#pragma omp parallel for reduction(+ : b) schedule(static)
	for (int64_t i = 0; i < n; i++) {
		x[i] -= a * y[i];
		b += x[i] * x[i];
	}
	b = sqrt(b);
	double bb = b < 1e-10 ? 1.0 : b;
#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < n; i++) {
		double tmp = y[i];
		y[i] = x[i] / bb;
		x[i] = -b * tmp;
	}
	return b;
}

/* ------------------------------------------------- tridiagonal eigen-solver */
/* symmetric tridiagonal QL with implicit shifts (EISPACK tql2 restated). d[n] diag, e[n] with
 * e[i] coupling i and i+1 (e[n-1] unused).  z (n*n, row-major z[i*n+k] = component i of vector k)
 * may be NULL.  Eigenvalues returned ascending in d. Returns 0 on success. */
int orc_tridiag_eig(int n, double* d, double* e_in, double* z)
{
	double* e = (double*)malloc(sizeof(double) * (n + 1));
	for (int i = 0; i < n - 1; i++) e[i] = e_in[i];
	e[n - 1] = 0.0;
	if (z) {
		for (int i = 0; i < n * n; i++) z[i] = 0.0;
		for (int i = 0; i < n; i++) z[i * n + i] = 1.0;
	}
	for (int l = 0; l < n; l++) {
		int iter = 0, mm;
		do {
			for (mm = l; mm < n - 1; mm++) {
				double dd = fabs(d[mm]) + fabs(d[mm + 1]);
				if (fabs(e[mm]) <= 2.3e-16 * dd) break;
			}
			if (mm != l) {
				if (iter++ == 200) { free(e); return 1; }
				double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
				double r = hypot(g, 1.0);
				g = d[mm] - d[l] + e[l] / (g + (g >= 0 ? fabs(r) : -fabs(r)));
				double s = 1.0, c = 1.0, p = 0.0;
				int i;
				for (i = mm - 1; i >= l; i--) {
					double f = s * e[i];
					double b = c * e[i];
					e[i + 1] = (r = hypot(f, g));
					if (r == 0.0) {
						d[i + 1] -= p;
						e[mm] = 0.0;
						break;
					}
					s = f / r;
					c = g / r;
					g = d[i + 1] - p;
					r = (d[i] - g) * s + 2.0 * c * b;
					d[i + 1] = g + (p = s * r);
					g = c * r - b;
					if (z) {
						for (int k = 0; k < n; k++) {
							f = z[k * n + i + 1];
							z[k * n + i + 1] = s * z[k * n + i] + c * f;
							z[k * n + i] = c * z[k * n + i] - s * f;
						}
					}
				}
				if (r == 0.0 && i >= l) continue;
				d[l] -= p;
				e[l] = g;
				e[mm] = 0.0;
			}
		} while (mm != l);
	}
	/* sort ascending */
	for (int i = 0; i < n - 1; i++) {
		int k = i;
		double p = d[i];
		for (int j = i + 1; j < n; j++)
			if (d[j] < p) { k = j; p = d[j]; }
		if (k != i) {
			d[k] = d[i];
			d[i] = p;
			if (z)
				for (int j = 0; j < n; j++) {
					double t = z[j * n + i];
					z[j * n + i] = z[j * n + k];
					z[j * n + k] = t;
				}
		}
	}
	free(e);
	return 0;
}

static double tridiag_lowest(int n, const double* a, const double* b)
{
	double* d = (double*)malloc(sizeof(double) * n);
	double* e = (double*)malloc(sizeof(double) * (n > 0 ? n : 1));
	memcpy(d, a, sizeof(double) * n);
	memcpy(e, b, sizeof(double) * n);
	orc_tridiag_eig(n, d, e, NULL);
	double r = d[0];
	free(d); free(e);
	return r;
}

/* ---------------------------------------------------------------- Lanczos */
typedef void (*orc_mv_fn)(void* ctx, double* x, const double* y);

typedef struct { const orc_model* m; int faithful; } mv_model_ctx;
static void mv_model(void* ctx, double* x, const double* y)
{
	mv_model_ctx* c = (mv_model_ctx*)ctx;
	orc_matvec(c->m, x, y, c->faithful);
}

typedef struct { size_t rows; const int64_t* rowptr; const int64_t* colind; const double* values; } mv_crs_ctx;
static void mv_crs(void* ctx, double* x, const double* y)
{
	mv_crs_ctx* c = (mv_crs_ctx*)ctx;
	orc_crs_matvec(c->rows, c->rowptr, c->colind, c->values, x, y);
}

/* PsimagLite LanczosCore::oneStepDecomposition: three separate sweeps */
static void one_step(orc_mv_fn mv, void* ctx, int64_t n, double* x, double* y, double* a_out, double* b_out)
{
	mv(ctx, x, y);
	double a = 0;
#pragma omp parallel for reduction(+ : a) schedule(static)
	for (int64_t i = 0; i < n; i++) a += y[i] * x[i];
	double b = 0;
#pragma omp parallel for reduction(+ : b) schedule(static)
	for (int64_t i = 0; i < n; i++) {
		x[i] -= a * y[i];
		b += x[i] * x[i];
	}
	b = sqrt(b);
	if (b < 1e-10) {
#pragma omp parallel for schedule(static)
		for (int64_t i = 0; i < n; i++) {
			double tmp = y[i];
			y[i] = x[i];
			x[i] = -b * tmp;
		}
	} else {
#pragma omp parallel for schedule(static)
		for (int64_t i = 0; i < n; i++) {
			double tmp = y[i];
			y[i] = x[i] / b;
			x[i] = -b * tmp;
		}
	}
	*a_out = a;
	*b_out = b;
}

/* <prefix>Options=reortho (PARITY UNPINNED: PsimagLite's own routine is absent): after x -= a y the new
 * vector is orthogonalised against every saved Lanczos vector v_0..v_j, classical Gram-Schmidt in blocks of 4 in basis order,
 * and b is the norm of the result.  Same step with reorthogonalisation as one_step() above. */
#define ORC_RO_NV 4
static void one_step_reortho(orc_mv_fn mv, void* ctx, int64_t n, double* x, double* y, double** saved, int nsaved,
                             double* a_out, double* b_out)
{
	mv(ctx, x, y);
	double a = 0;
#pragma omp parallel for reduction(+ : a) schedule(static)
	for (int64_t i = 0; i < n; i++) a += y[i] * x[i];
#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < n; i++) x[i] -= a * y[i];
	for (int k0 = 0; k0 < nsaved; k0 += ORC_RO_NV) {
		int nv = nsaved - k0 < ORC_RO_NV ? nsaved - k0 : ORC_RO_NV;
		double c[ORC_RO_NV] = {0, 0, 0, 0};
		for (int k = 0; k < nv; k++) {
			double s = 0;
			const double* v = saved[k0 + k];
#pragma omp parallel for reduction(+ : s) schedule(static)
			for (int64_t i = 0; i < n; i++) s += v[i] * x[i];
			c[k] = s;
		}
#pragma omp parallel for schedule(static)
		for (int64_t i = 0; i < n; i++) {
			double t = x[i];
			for (int k = 0; k < nv; k++) t -= c[k] * saved[k0 + k][i];
			x[i] = t;
		}
	}
	double b = 0;
#pragma omp parallel for reduction(+ : b) schedule(static)
	for (int64_t i = 0; i < n; i++) b += x[i] * x[i];
	b = sqrt(b);
	const double div = (b < 1e-10) ? 1.0 : b;
#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < n; i++) {
		double tmp = y[i];
		y[i] = x[i] / div;
		x[i] = -b * tmp;
	}
	*a_out = a;
	*b_out = b;
}

static double tridiag_kth(int n, const double* a, const double* b, int k)
{
	double* d = (double*)malloc(sizeof(double) * n);
	double* e = (double*)malloc(sizeof(double) * (n > 0 ? n : 1));
	memcpy(d, a, sizeof(double) * n);
	memcpy(e, b, sizeof(double) * n);
	orc_tridiag_eig(n, d, e, NULL);
	double r = d[k];
	free(d); free(e);
	return r;
}

/* LanczosSolver::computeAllStatesBelow (Engine.h:626; PARITY UNPINNED, PsimagLite absent): one decomposition with every
 * vector saved and reorthogonalised, convergence watched on Ritz value nstates-1, then the lowest nstates Ritz pairs;
 * z (nstates x rows) may be NULL.  Returns the step count. */
int orc_states_below(const orc_model* m, const double* init, int steps, double eps, int minsteps, int nstates,
                     double* energies, double* z)
{
	mv_model_ctx c = {m, 0};
	int64_t n = (int64_t)orc_rows(m);
	double* x = (double*)calloc(n, sizeof(double));
	double* y = (double*)malloc(sizeof(double) * n);
	double nrm = 0;
	for (int64_t i = 0; i < n; i++) nrm += init[i] * init[i];
	nrm = sqrt(nrm);
	for (int64_t i = 0; i < n; i++) y[i] = init[i] / nrm;
	if (steps > n) steps = (int)n;
	double** saved = (double**)calloc(steps > 0 ? steps : 1, sizeof(double*));
	double* a = (double*)calloc(steps + 1, sizeof(double));
	double* b = (double*)calloc(steps + 1, sizeof(double));
	double eold = 100.0;
	int j = 0;
	for (; j < steps; j++) {
		saved[j] = (double*)malloc(sizeof(double) * n);
		memcpy(saved[j], y, sizeof(double) * n);
		one_step_reortho(mv_model, &c, n, x, y, saved, j + 1, &a[j], &b[j]);
		if (eps > 0 && j >= nstates - 1) {
			double enew = tridiag_kth(j + 1, a, b, nstates - 1);
			if (fabs(enew - eold) < eps && (j >= minsteps || n <= 4)) { j++; break; }
			eold = enew;
		}
	}
	int ns = j;
	double* d = (double*)malloc(sizeof(double) * ns);
	double* e = (double*)calloc(ns > 0 ? ns : 1, sizeof(double));
	double* zz = (double*)malloc(sizeof(double) * (size_t)ns * ns);
	memcpy(d, a, sizeof(double) * ns);
	for (int i = 0; i + 1 < ns; i++) e[i] = b[i];
	orc_tridiag_eig(ns, d, e, zz);
	for (int k = 0; k < nstates && k < ns; k++) {
		energies[k] = d[k];
		if (!z) continue;
		for (int64_t i = 0; i < n; i++) z[(size_t)k * n + i] = 0;
		for (int q = 0; q < ns; q++) {
			double cq = zz[(size_t)q * ns + k];
			for (int64_t i = 0; i < n; i++) z[(size_t)k * n + i] += cq * saved[q][i];
		}
	}
	for (int k = 0; k < steps; k++) free(saved[k]);
	free(saved); free(x); free(y); free(a); free(b); free(d); free(e); free(zz);
	return ns;
}

/* decomposition with every Lanczos vector saved and full reorthogonalisation; returns the step count */
int orc_lanczos_decomposition_reortho(const orc_model* m, const double* init, int steps, double eps, int minsteps,
                                      double* a, double* b)
{
	mv_model_ctx c = {m, 0};
	int64_t n = (int64_t)orc_rows(m);
	double* x = (double*)calloc(n, sizeof(double));
	double* y = (double*)malloc(sizeof(double) * n);
	double nrm = 0;
	for (int64_t i = 0; i < n; i++) nrm += init[i] * init[i];
	nrm = sqrt(nrm);
	for (int64_t i = 0; i < n; i++) y[i] = init[i] / nrm;
	if (steps > n) steps = (int)n;
	double** saved = (double**)calloc(steps > 0 ? steps : 1, sizeof(double*));
	double eold = 100.0;
	int j = 0;
	for (; j < steps; j++) {
		saved[j] = (double*)malloc(sizeof(double) * n);
		memcpy(saved[j], y, sizeof(double) * n);
		one_step_reortho(mv_model, &c, n, x, y, saved, j + 1, &a[j], &b[j]);
		if (eps > 0) {
			double enew = tridiag_lowest(j + 1, a, b);
			if (fabs(enew - eold) < eps && (j >= minsteps || n <= 4)) { j++; break; }
			eold = enew;
		}
	}
	for (int k = 0; k < steps; k++) free(saved[k]);
	free(saved); free(x); free(y);
	return j;
}

/* PsimagLite LanczosSolver::decomposition.  If zcoef!=NULL (second pass, App. B.4)
 * accumulates z += zcoef[j] * v_j while replaying exactly nfixed steps. Returns the step count. */
static int decomposition(orc_mv_fn mv, void* ctx, int64_t n, const double* init, int steps, double eps,
                         int minsteps, double* a, double* b, const double* zcoef, int nfixed, double* z)
{
	double* x = (double*)calloc(n, sizeof(double));
	double* y = (double*)malloc(sizeof(double) * n);
	double nrm = 0;
	for (int64_t i = 0; i < n; i++) nrm += init[i] * init[i];
	nrm = sqrt(nrm);
	for (int64_t i = 0; i < n; i++) y[i] = init[i] / nrm;
	if (steps > n) steps = (int)n;
	if (zcoef) steps = nfixed;
	double eold = 100.0;
	int j = 0, done = 0;
	for (; j < steps; j++) {
		if (zcoef) {
			double c = zcoef[j];
#pragma omp parallel for schedule(static)
			for (int64_t i = 0; i < n; i++) z[i] += c * y[i];
		}
		double aj, bj;
		one_step(mv, ctx, n, x, y, &aj, &bj);
		a[j] = aj;
		b[j] = bj;
		if (!zcoef && eps > 0) {
			double enew = tridiag_lowest(j + 1, a, b);
			if (fabs(enew - eold) < eps && (j >= minsteps || n <= 4)) { done = 1; j++; break; }
			eold = enew;
		}
	}
	(void)done;
	free(x); free(y);
	return j;
}

int orc_lanczos_decomposition(const orc_model* m, int faithful, const double* init, int steps, double eps,
                              int minsteps, double* a, double* b)
{
	mv_model_ctx c = {m, faithful};
	return decomposition(mv_model, &c, (int64_t)orc_rows(m), init, steps, eps, minsteps, a, b, NULL, 0, NULL);
}

int orc_lanczos_decomposition_crs(size_t rows, const int64_t* rowptr, const int64_t* colind, const double* values,
                                  const double* init, int steps, double eps, int minsteps, double* a, double* b)
{
	mv_crs_ctx c = {rows, rowptr, colind, values};
	return decomposition(mv_crs, &c, (int64_t)rows, init, steps, eps, minsteps, a, b, NULL, 0, NULL);
}

/* PsimagLite LanczosSolver::computeOneState: energy = lowest Ritz value of T,
 * z = sum_j c_j v_j by replaying the recurrence (vectors not saved). z may be NULL. */
int orc_ground_state(const orc_model* m, int faithful, const double* init, int steps, double eps, int minsteps,
                     double* energy, double* z, double* a_out, double* b_out)
{
	int64_t n = (int64_t)orc_rows(m);
	int cap = steps > n ? (int)n : steps;
	double* a = (double*)calloc(cap + 1, sizeof(double));
	double* b = (double*)calloc(cap + 1, sizeof(double));
	mv_model_ctx c = {m, faithful};
	int ns = decomposition(mv_model, &c, n, init, steps, eps, minsteps, a, b, NULL, 0, NULL);
	double* d = (double*)malloc(sizeof(double) * ns);
	double* e = (double*)malloc(sizeof(double) * ns);
	double* zz = (double*)malloc(sizeof(double) * ns * ns);
	memcpy(d, a, sizeof(double) * ns);
	memcpy(e, b, sizeof(double) * ns);
	orc_tridiag_eig(ns, d, e, zz);
	*energy = d[0];
	if (z) {
		double* coef = (double*)malloc(sizeof(double) * ns);
		for (int j = 0; j < ns; j++) coef[j] = zz[j * ns + 0];
		for (int64_t i = 0; i < n; i++) z[i] = 0;
		double* a2 = (double*)calloc(ns + 1, sizeof(double));
		double* b2 = (double*)calloc(ns + 1, sizeof(double));
		decomposition(mv_model, &c, n, init, steps, eps, minsteps, a2, b2, coef, ns, z);
		free(coef); free(a2); free(b2);
	}
	if (a_out) memcpy(a_out, a, sizeof(double) * ns);
	if (b_out) memcpy(b_out, b, sizeof(double) * ns);
	free(a); free(b); free(d); free(e); free(zz);
	return ns;
}

/* ------------------------------------------------------- continued fraction */
/* PsimagLite ContinuedFraction: diagonalise T(a,b); intensity_l = (first component)^2;
 * G(z) = weight * sum_l I_l / (z - isign*(eps_l - Eg)), z = omega + i delta. out = (re,im) pairs. */
void orc_cf_eval(int n, const double* a, const double* b, double Eg, double weight, int isign,
                 int nomega, const double* omega, double delta, double* out)
{
	double* d = (double*)malloc(sizeof(double) * n);
	double* e = (double*)malloc(sizeof(double) * n);
	double* zz = (double*)malloc(sizeof(double) * n * n);
	memcpy(d, a, sizeof(double) * n);
	memcpy(e, b, sizeof(double) * n);
	orc_tridiag_eig(n, d, e, zz);
	for (int w = 0; w < nomega; w++) {
		double re = 0, im = 0;
		for (int l = 0; l < n; l++) {
			double I = zz[0 * n + l] * zz[0 * n + l];
			double xr = omega[w] - isign * (d[l] - Eg);
			double den = xr * xr + delta * delta;
			re += weight * I * xr / den;
			im += -weight * I * delta / den;
		}
		out[2 * w] = re;
		out[2 * w + 1] = im;
	}
	free(d); free(e); free(zz);
}

/* -------------------------------------------- operator application (CF path) */
/* BasisHubbardLanczos.h:106-137 (literal, including the SPIN_DOWN overwrite of the up parity, quirk C.3) */
static int hubbard_dosign_gf(word_t a, word_t b, int ind, int sector)
{
	if (sector == 0) {
		if (ind == 0) return 1;
		word_t mask = a;
		mask &= ((((word_t)1) << 1) - 1) ^ ((((word_t)1) << ind) - 1);
		int s = (popc(mask) & 1) ? -1 : 1;
		if (a & 1) s = -s;
		return s;
	}
	int s = (popc(a) & 1) ? -1 : 1;
	if (ind == 0) return s;
	word_t mask = b;
	mask &= ((((word_t)1) << 1) - 1) ^ ((((word_t)1) << ind) - 1);
	s = (popc(mask) & 1) ? -1 : 1;
	if (b & 1) s = -s;
	return s;
}

/* BasisOneSpinFeAs.h:227-239: parity of the occupied orbitals below site*orbitals + orb */
static int feas_dosign_gf1(word_t a, int ind, int orb, int orbitals)
{
	int sum = feas_nbyket(a, 0, ind * orbitals);
	sum += feas_nbyket(a, ind * orbitals, ind * orbitals + orb);
	return (sum & 1) ? -1 : 1;
}

/* BasisTjMultiOrbLanczos.h:163-192 */
static int tj_dosign_gf(word_t a, word_t b, int ind, int sector)
{
	if (sector == 0) {
		if (ind == 0) return 1;
		word_t mask = a & ((((word_t)1 << 1) - 1) ^ (((word_t)1 << ind) - 1));
		int s = (popc(mask) & 1) ? -1 : 1;
		if (a & 1) s = -s;
		return s;
	}
	int s = (popc(a) & 1) ? -1 : 1;
	if (ind == 0) return s;
	word_t mask = b & ((((word_t)1 << 1) - 1) ^ (((word_t)1 << ind) - 1));
	s *= (popc(mask) & 1) ? -1 : 1;
	if (b & 1) s = -s;
	return s;
}

/* Engine.h:416-458 accModifiedState_ with c / cdagger (/ n for Hubbard): getBraIndex of the new basis
 * (BasisHubbardLanczos.h:162-182 with BasisOneSpin.h:121-151; BasisFeAsBasedSc.h:276-289 with BasisOneSpinFeAs.h:127-143;
 * BasisTjMultiOrbLanczos.h:302-318,398-433), doSignGf of the source basis, 64-bit indices instead of int.
 * z (dst basis) += factor*sign*src. */
/* accModifiedState_ for the spin operators: getBraIndex of BasisHubbardLanczos.h:162-257 (sz -> getBraIndexSz, splus/sminus ->
 * getBraIndexSplusSminus, sign doSignSpSm :151-160) and of BasisHeisenberg.h:123-139,230-280 (sz value 1 - 2 n_up, n value
 * n_up or 1 - n_up, splus/sminus flip the site; doSignSpSm is the BasisBase default 1) */
static void apply_spin_op(const orc_model* src, const orc_model* dst, int op, int site, int spin, int orb, double factor,
                          const double* srcv, double* z)
{
	size_t n = orc_rows(src);
	int pos = site * src->orbitals + orb;
	word_t ms = ((word_t)1) << pos;
	for (size_t r = 0; r < n; r++) {
		word_t k1, k2;
		row_kets(src, r, &k1, &k2);
		long idx = -1;
		double value = 1, mysign = 1;
		if (src->model == ORC_HEISENBERG) {
			int nup = (k1 & ms) ? 1 : 0;
			if (op == ORC_OP_SZ) { idx = (long)r; value = 1 - 2 * nup; }
			else if (op == ORC_OP_N) { idx = (long)r; value = (spin == 0) ? nup : 1 - nup; }
			else {
				if (nup) { if (op == ORC_OP_SPLUS) continue; }
				else { if (op == ORC_OP_SMINUS) continue; }
				idx = (long)perfect_index(dst, k1 ^ ms, 0);
			}
		} else { /* Hubbard; FeAs per orbital (BasisFeAsBasedSc.h:291-303,356-379, doSignSpSm :202-211); t-J (BasisTjMultiOrbLanczos.h:213-242) */
			int b1 = (k1 & ms) ? 1 : 0, b2 = (k2 & ms) ? 1 : 0;
			if (op == ORC_OP_SZ) {
				if (src->model != ORC_HUBBARD) continue;
				if (!b1 && !b2) continue;
				if (b1 && b2) continue;
				value = b1 ? 1 : -1;
				idx = (long)perfect_index(dst, k1, k2);
			} else {
				int up_first = (op == ORC_OP_SPLUS);           /* spin = SPIN_UP for S+, SPIN_DOWN for S- */
				word_t brar1, brar2;
				if (up_first) {
					if (b1) continue;                          /* cdagger up: needs the up orbital empty */
					brar1 = k1 ^ ms;
					if (!b2) continue;                         /* c down: needs the down orbital occupied */
					brar2 = k2 ^ ms;
					idx = (long)perfect_index(dst, brar1, brar2);
				} else {
					if (b2) continue;                          /* cdagger down */
					brar1 = k2 ^ ms;
					if (!b1) continue;                         /* c up */
					brar2 = k1 ^ ms;
					idx = (long)perfect_index(dst, brar2, brar1);
				}
				if (src->model != ORC_TJ) mysign = do_sign(k1, pos) * do_sign(k2, pos);   /* doSignSpSm (BasisBase default 1 for t-J) */
			}
		}
		if (idx < 0) continue;
		z[idx] += factor * mysign * value * srcv[r];
	}
}

void orc_apply_op_orb(const orc_model* src, const orc_model* dst, int op, int site, int spin, int orb, double factor,
                      const double* srcv, double* z)
{
	if (op == ORC_OP_SZ || op == ORC_OP_SPLUS || op == ORC_OP_SMINUS || src->model == ORC_HEISENBERG) {
		apply_spin_op(src, dst, op, site, spin, orb, factor, srcv, z);
		return;
	}
	size_t n = orc_rows(src);
	int pos = site * src->orbitals + orb;
	word_t ms = ((word_t)1) << pos;
	for (size_t r = 0; r < n; r++) {
		word_t k1, k2;
		row_kets(src, r, &k1, &k2);
		word_t ket = spin == 0 ? k1 : k2;
		word_t bra;
		word_t si = ket & ms;
		if (op == ORC_OP_C) { if (!si) continue; bra = ket ^ ms; }
		else if (op == ORC_OP_CDAGGER) { if (si) continue; bra = ket ^ ms; }
		else { if (!si) continue; bra = ket; }
		if (src->model == ORC_TJ && (spin == 0 ? (bra & k2) : (k1 & bra))) continue;   /* isDoublyOccupied */
		size_t idx = spin == 0 ? perfect_index(dst, bra, k2) : perfect_index(dst, k1, bra);
		double mysign = 1;
		if (op == ORC_OP_C || op == ORC_OP_CDAGGER) {
			if (src->model == ORC_HUBBARD) mysign = hubbard_dosign_gf(k1, k2, site, spin);
			else if (src->model == ORC_FEAS) {
				if (spin == 0) mysign = feas_dosign_gf1(k1, site, orb, src->orbitals);
				else mysign = ((popc(k1) & 1) ? -1 : 1) * feas_dosign_gf1(k2, site, orb, src->orbitals);
			} else mysign = tj_dosign_gf(k1, k2, site, spin);
		}
		z[idx] += factor * mysign * 1.0 * srcv[r];
	}
}

/* Engine::twoPoint (Engine.h:262-331) with bra = ket = gs: result(isite, jsite) = modifVector2 * modifVector1 */
void orc_two_point(const orc_model* src, const orc_model* dst, int op, int spin, int orb_i, int orb_j, const double* gs,
                   double* result)
{
	int nsite = src->nsite;
	size_t nd = orc_rows(dst);
	double* m1 = (double*)malloc(sizeof(double) * (nd ? nd : 1));
	double* m2 = (double*)malloc(sizeof(double) * (nd ? nd : 1));
	for (int isite = 0; isite < nsite; isite++) {
		memset(m1, 0, sizeof(double) * nd);
		orc_apply_op_orb(src, dst, op, isite, spin, orb_i, 1.0, gs, m1);
		for (int jsite = 0; jsite < nsite; jsite++) {
			memset(m2, 0, sizeof(double) * nd);
			orc_apply_op_orb(src, dst, op, jsite, spin, orb_j, 1.0, gs, m2);
			double s = 0;
			for (size_t k = 0; k < nd; k++) s += m2[k] * m1[k];
			result[isite * nsite + jsite] = s;
		}
	}
	free(m1); free(m2);
}

void orc_apply_op(const orc_model* src, const orc_model* dst, int op, int site, int spin, double factor,
                  const double* srcv, double* z)
{
	orc_apply_op_orb(src, dst, op, site, spin, 0, factor, srcv, z);
}

int orc_num_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/* bench.py --impl reference: torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm uses all host cores anyway */
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
	if (n > 0) omp_set_num_threads(n);
#else
	(void)n;
#endif
}
