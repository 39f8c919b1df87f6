/*
 * lpp_b200.h -- C-ABI of the B200-native Lanczos engine (liblpp_b200.so).
 *
 * Drop-in boundary for the ONE hot path of g1257/LanczosPlusPlus: the Hamiltonian matrix-vector product
 * and Lanczos recurrence PsimagLite::LanczosSolver drives through MatrixType::rows() /
 * MatrixType::matrixVectorProduct(x, y).  Citations are file:line under the reference's src/.
 * The C++ adapter a maintainer adds to the reference (InternalProductCuda, a sibling of
 * InternalProductOnTheFly / InternalProductStored) is include/InternalProductCuda.h; INTEGRATION.md
 * shows the wiring.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every entry point returns 0 on success or a
 * negative lpp_status, the message is available from lpp_last_error() (thread-local).  The reference signals
 * errors with exceptions (PsimagLite::RuntimeError / err(), e.g. InternalProductOnTheFly.h:129-133); the C++
 * shim converts a non-zero status into err(lpp_last_error()).
 * There is NO CPU fallback: every compute entry point fails with LPP_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef LPP_B200_H
#define LPP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lpp_handle lpp_handle;

typedef enum {
	LPP_OK = 0,
	LPP_ERR_ARG = -1,      /* bad argument / unsupported model option */
	LPP_ERR_CUDA = -2,     /* CUDA runtime failure (includes "no device") */
	LPP_ERR_STATE = -3,    /* call order (e.g. CRS export before build) */
	LPP_ERR_NCCL = -4,
	LPP_ERR_OVERFLOW = -5  /* a row produced more entries than the builder supports */
} lpp_status;

/* Models on the path (ModelSelector.h:45-96 names): */
typedef enum {
	LPP_MODEL_HUBBARD = 0,    /* HubbardOneBand        : src/Models/HubbardOneOrbital  */
	LPP_MODEL_FEAS = 1,       /* FeAsBasedSc INT_PAPER33: src/Models/FeBasedSc          */
	LPP_MODEL_HEISENBERG = 2, /* Heisenberg, TwiceS=1  : src/Models/Heisenberg         */
	LPP_MODEL_TJ = 3          /* Tj1Orbital (TjMultiOrb, Orbitals=1, no JHundInfinity): src/Models/TjMultiOrb;
	                             stored CRS (the reference's only path for it) and the generic on-the-fly kernel */
} lpp_model;

/* How x += H y is evaluated (LanczosDriver1.h:222-238 chooses Stored vs OnTheFly from SolverOptions=). */
typedef enum {
	LPP_KERNEL_AUTO = 0,     /* fastest on-the-fly variant for the model */
	LPP_KERNEL_GENERIC = 1,  /* one thread per row, hops enumerated with bit operations, states ranked on the fly */
	LPP_KERNEL_TABLE = 2,    /* product bases: per-spin hop tables built on device, gathers from global memory */
	LPP_KERNEL_TILED = 3,    /* product bases: shared-memory tiled two-sweep kernel (the B200 fast path) */
	LPP_KERNEL_STORED = 4    /* stored CRS built on device (InternalProductStored) */
} lpp_kernel;

/* Operator ids of LabeledOperator::Label (LabeledOperator.h:10-17). */
typedef enum { LPP_OP_C = 1, LPP_OP_SZ = 2, LPP_OP_CDAGGER = 3, LPP_OP_N = 4, LPP_OP_SPLUS = 5, LPP_OP_SMINUS = 6 } lpp_op;

/* One (model, symmetry sector).  Replaces the model constructor + createBasis():
 * HubbardOneOrbital.h:41-49,117-122 ; FeBasedSc.h:132-140,249-254 ; Heisenberg.h:38-60.
 * Matrices are dense row-major nb x nb with nb = nsite*orbitals and index site*orbitals+orb
 * (BasisOneSpinFeAs.h:195-199): the values geometry(i,orb_i,j,orb_j,term) of PsimagLite::Geometry. */
typedef struct {
	int32_t model;        /* lpp_model */
	int32_t nsite;
	int32_t orbitals;     /* 1 for Hubbard / Heisenberg */
	int32_t nup;          /* TargetElectronsUp   (Heisenberg: TargetSzPlusConst) */
	int32_t ndown;        /* TargetElectronsDown (Heisenberg: ignored) */
	int32_t feas_u3_all_pairs; /* 1: Hermitian/stored definition FeBasedSc.h:192-197 (default); 0: literal OTF doTask :85-88 */
	const double* hop;    /* term 0: hoppings_(i,j) HubbardHelper.h:60-71 | geometry(i,o,j,o2,0) FeBasedSc.h:320-323 | jpm Heisenberg.h:56 */
	const double* jzz;    /* term 1: Heisenberg jzz (Heisenberg.h:57); NULL otherwise */
	const double* U;      /* hubbardU: Hubbard nsite values (ParametersModelHubbard.h:93); FeAs 4..6 values (ParametersModelFeAs.h:100-151) */
	int32_t nU;
	const double* V;      /* potentialV: Hubbard uses [i] only (HubbardHelper.h:180); FeAs i+(orb+orbitals*spin)*nsite (FeBasedSc.h:558-561); Heisenberg: MagneticField */
	int32_t nV;
	const double* D;      /* Heisenberg AnisotropyD vector (ParametersHeisenberg.h:100) ; FeAs: D[0] = AnisotropyD= scalar */
	int32_t nD;
	int32_t device;       /* CUDA device ordinal */
	int32_t rank;         /* row sharding over the slow (spin-down) index: this shard ... */
	int32_t nranks;       /* ... of nranks (1 = whole Hilbert space on this GPU) */
	/* t-J only (NULL otherwise), TjMultiOrb.h:68-79: hop = term 0, jpm = term 1, jzz = term 2, w = term 3 */
	const double* jpm;
	const double* w;
} lpp_desc;

/* PsimagLite::ParametersForSolver: <prefix>Steps, <prefix>Eps, <prefix>MinSteps. */
typedef struct {
	int32_t steps;     /* default 200 */
	int32_t minsteps;  /* default 4 */
	double eps;        /* default 1e-12; <=0: run exactly `steps` */
	int32_t kernel;    /* lpp_kernel */
	int32_t reortho;   /* <prefix>Options=reortho : full reorthogonalisation against saved vectors */
	uint64_t seed;     /* used when no initial vector is given */
} lpp_solver_params;

const char* lpp_last_error(void);
int lpp_version(void);
/* 0 if an sm_100 device is present and usable, LPP_ERR_CUDA otherwise (never falls back to the CPU). */
int lpp_device_check(int32_t device);

int lpp_create(const lpp_desc* desc, lpp_handle** out);
int lpp_destroy(lpp_handle* h);

/* InternalProductOnTheFly::rows() (InternalProductOnTheFly.h:115-118) = basis.size() (BasisHubbardLanczos.h:43). */
int lpp_rows(const lpp_handle* h, uint64_t* rows);
/* rows owned by this shard: [first, first+count) */
int lpp_local_rows(const lpp_handle* h, uint64_t* first, uint64_t* count);

/* Basis built on device, bit-exact with BasisOneSpin.h:25-63 / BasisOneSpinFeAs.h:45-84 / BasisHeisenberg.h:24-47.
 * spin: 0 = up (or the single Heisenberg word), 1 = down. */
int lpp_basis_size(const lpp_handle* h, int32_t spin, uint64_t* n);
int lpp_basis_export(const lpp_handle* h, int32_t spin, uint64_t* words);
/* perfectIndex of n one-spin words computed on device (BasisOneSpin.h:73-81; closed form replacing the linear
 * searches of BasisOneSpinFeAs.h:96-101 and BasisHeisenberg.h:73-80). */
int lpp_rank(const lpp_handle* h, int32_t spin, const uint64_t* words, uint64_t n, uint64_t* index);
/* basis(i, SPIN_UP) / basis(i, SPIN_DOWN) for rows [first, first+count) (BasisBase::operator(), e.g.
 * BasisHubbardLanczos.h:77-84, BasisTjMultiOrbLanczos.h:127-141), computed on the device; either output may be NULL. */
int lpp_row_words(const lpp_handle* h, uint64_t first, uint64_t count, uint64_t* up_words, uint64_t* down_words);
/* perfectIndex(ket1, ket2) of the full basis for n word pairs (BasisHubbardLanczos.h:59-63, BasisFeAsBasedSc.h:91-100,
 * BasisTjMultiOrbLanczos.h:71-112; Heisenberg ignores ket2), computed on the device. */
int lpp_rank_pairs(const lpp_handle* h, const uint64_t* up_words, const uint64_t* down_words, uint64_t n, uint64_t* index);

/* MatrixType::matrixVectorProduct(x, y): x += H y with caller-owned HOST vectors of rows() doubles
 * (InternalProductOnTheFly.h:120-123 -> HubbardHelper.h:105-134 / FeBasedSc.h:228-245).  Single shard only. */
int lpp_matvec_host(lpp_handle* h, int32_t kernel, double* x, const double* y);
/* Same with DEVICE pointers on h's device (local rows of x; y is the full vector when nranks>1). */
int lpp_matvec_device(lpp_handle* h, int32_t kernel, double* x_dev, const double* y_dev);

/* InternalProductStored: model.setupHamiltonian(matrix, basis) on device (HubbardHelper.h:75-103, FeBasedSc.h:163-221,
 * Heisenberg.h:80-114) with PsimagLite::SparseRow::finalize semantics (diagonal always stored, columns sorted,
 * duplicates merged, explicit zeros kept). */
int lpp_crs_build(lpp_handle* h, int64_t* nnz);
int lpp_crs_export(const lpp_handle* h, int64_t* rowptr, int64_t* colind, double* values);

/* PsimagLite::LanczosSolver::decomposition(init, ab) (call site Engine.h:474-478), device resident.
 * init_host: rows() doubles or NULL (splitmix64(seed) vector generated on device, or the vector left in the handle by
 * lpp_apply_op when use_modified != 0).  a,b: capacity >= min(steps, rows).  */
int lpp_lanczos_decomposition(lpp_handle* h, const lpp_solver_params* p, const double* init_host, int32_t use_modified,
                              double* a, double* b, int32_t* nsteps, double* init_norm2);
/* PsimagLite::LanczosSolver::computeOneState / computeAllStatesBelow(eigs, zs, initial, 1) (Engine.h:609-626):
 * lowest Ritz value and, when want_vector != 0, the Ritz vector by a second replay pass (vectors are never saved);
 * the vector stays on device inside the handle (and is copied to z_host when non-NULL). */
int lpp_ground_state(lpp_handle* h, const lpp_solver_params* p, const double* init_host, int32_t want_vector,
                     double* energy, double* z_host, double* a, double* b, int32_t* nsteps);

/* lanczosSolver.computeAllStatesBelow(eigs, zs, initial, excitedPlusOne) (Engine.h:601-657 with excited > 0): the lowest nstates
 * Ritz values of one decomposition (convergence watched on state nstates-1) and, when z_host != NULL, their vectors
 * (nstates x rows, state k at z_host + k*rows).  State 0 stays in the handle as the ground state. */
int lpp_states_below(lpp_handle* h, const lpp_solver_params* p, const double* init_host, int32_t nstates, double* energies,
                     double* z_host, int32_t* nsteps);

/* Engine::accModifiedState_ (Engine.h:416-458): dst.modified (+)= factor * O_{site,spin,orb} |src.groundstate>, 64-bit indices.
 * HubbardOneBand: c, cdagger, n, sz, splus, sminus (BasisHubbardLanczos.h:106-257); FeAsBasedSc (per orbital) and Tj1Orbital: c,
 * cdagger, splus, sminus; Heisenberg S=1/2: sz, n, splus, sminus (BasisHeisenberg.h:123-139,230-280).  dst is a handle on the sector hasNewParts
 * gives (src itself for sz / n).
 * accumulate == 0 zeroes dst.modified first. */
int lpp_apply_op(lpp_handle* src, lpp_handle* dst, int32_t op, int32_t site, int32_t spin, int32_t orb, double factor,
                 int32_t accumulate);
/* Engine::twoPoint (Engine.h:262-331) with bra = ket = ground state: result[i*nsite + j] = <O_{j,spin,orb_j} gs | O_{i,spin,orb_i} gs>
 * for the operators of lpp_apply_op (c: the one-body density matrix <cdagger_j c_i>; n: <n_j n_i>); dst is a handle on the
 * sector the operator leads to (src itself for n).  Modified states and the Gram matrix stay on the device. */
int lpp_two_point(lpp_handle* src, lpp_handle* dst, int32_t op, int32_t spin, int32_t orb_i, int32_t orb_j, double* result);
/* Engine::measure (Engine.h:208-249) for <gs| op_0[site_0]; ...; op_{n-1}[site_{n-1}] |gs> with ModelBase::rahulMethod semantics
 * (ModelBase.h:89-141, RahulOperator.h:26-50): labels 0 identity, 1 n, 2 sz, 3 c (cdagger when transposes[i] != 0); dofs 0 up,
 * 1 down; sites are bit positions in the one-spin word (site*orbitals + orb).  The operators must conserve the sector. */
int lpp_measure(lpp_handle* h, int32_t nops, const int32_t* labels, const int32_t* dofs, const int32_t* transposes,
                const int32_t* sites, double* result);
/* Engine::manyPoint (Engine.h:341-389), bra = ket = ground state of chain[0]: tmp_0 = |gs>, tmp_k = O_k tmp_{k-1} with
 * O_k = ops[k-1] at (sites[k-1], spins[k-1], orbs[k-1]) applied like lpp_apply_op (accModifiedState_, isign = 1); chain[k] is a handle on
 * the sector the k-th operator leads to (the basis Engine::getNeededBasis creates, Engine.h:391-413; chain[k] may equal chain[k-1] for
 * sz / n); chain[nops] must be on the ground state's sector (it may be chain[0]).  result = <gs | tmp_nops>. */
int lpp_many_point(lpp_handle* const* chain, int32_t nops, const int32_t* ops, const int32_t* sites, const int32_t* spins,
                   const int32_t* orbs, double* result);
/* copy the handle's ground-state / modified vector to the host (parity tests) */
int lpp_get_vector(lpp_handle* h, int32_t which /*0 = ground state, 1 = modified*/, double* out_host);
int lpp_set_groundstate(lpp_handle* h, const double* z_host);

/* PsimagLite::ContinuedFraction::set + plot (Engine.h:487-489), host side:
 * G(omega + i delta) = weight * sum_l I_l / (z - isign (eps_l - Eg)); out = nomega (re, im) pairs. */
int lpp_cf_eval(int32_t n, const double* a, const double* b, double eg, double weight, int32_t isign, int32_t nomega,
                const double* omega, double delta, double* out);
/* symmetric tridiagonal eigen-solver used by the Krylov loops (PsimagLite::TridiagonalMatrix role): eigenvalues
 * ascending into eigs[n]; if z != NULL, z[i*n+k] = component i of eigenvector k. */
int lpp_tridiag_eig(int32_t n, const double* a, const double* b, double* eigs, double* z);

/* Multi-GPU: one process per GPU.  The caller (torch.distributed or MPI-less launcher) creates the id on rank 0 with
 * lpp_comm_unique_id, broadcasts the 128 bytes, and every rank calls lpp_comm_init. */
int lpp_comm_unique_id(uint8_t id[128]);
int lpp_comm_init(lpp_handle* h, const uint8_t id[128]);
/* A new-sector handle created for the continued-fraction path (Engine.h:165-187 creates the (nup-1, ndown) basis while the
 * ground-state handle is alive) borrows the communicator of the handle it was derived from: same ranks, same device.  The
 * borrowed communicator is not destroyed with `h`; `parent` must outlive it and must be idle while `h` runs. */
int lpp_comm_share(lpp_handle* h, const lpp_handle* parent);

/* Peer-memory exchange for the two-layout sharding (Hubbard-type product bases, nranks > 1): every rank exports the CUDA IPC
 * handles of its column-shard buffers (128 bytes), the launcher all-gathers them, every rank imports the nranks*128 bytes.
 * After that the row<->column re-layout of each mat-vec is done by the engine's own kernels with stores/loads on the
 * peers' memory over NVLink; NCCL only carries the scalar all-reduces.  lpp_p2p_export returns LPP_ERR_STATE when the
 * sharding does not apply (the engine then uses the NCCL paths). */
int lpp_p2p_export(lpp_handle* h, int32_t kernel, uint8_t handles[128]);
int lpp_p2p_import(lpp_handle* h, const uint8_t* all_handles);

/* test hook: all-reduce (sum) of n <= 4 doubles over the handle's ranks through the scalar path of the sharded Krylov loop */
int lpp_allreduce_selftest(lpp_handle* h, double* inout, int32_t n);

/* contiguous near-equal split of n items over nranks (the row-sharding rule used for the slow index) */
int lpp_shard_range(uint64_t n, int32_t rank, int32_t nranks, uint64_t* first, uint64_t* count);

/* Measurement hooks used by bench.py (not part of the reference interface). */
typedef struct {
	double spmv_ms;      /* CUDA-event average of one x += H y */
	double iter_ms;      /* CUDA-event average of one full Lanczos iteration */
	int64_t launches;    /* kernels launched inside the timed region */
} lpp_timing;
int lpp_bench_spmv(lpp_handle* h, int32_t kernel, int32_t iters, int32_t warmup, lpp_timing* t);
int lpp_bench_lanczos(lpp_handle* h, const lpp_solver_params* p, int32_t iters, int32_t warmup, lpp_timing* t);

#ifdef __cplusplus
}
#endif
#endif
