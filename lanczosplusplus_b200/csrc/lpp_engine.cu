// lpp_engine.cu -- host side of liblpp_b200.so: handle management, device basis/table construction, kernel
// dispatch, the device-resident Krylov loops (PsimagLite::LanczosSolver role) and the C-ABI of include/lpp_b200.h.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lpp_b200.h"
#include "lpp_kernels.cuh"
#include "lpp_tiled.cuh"
#include "lpp_setup.h"

// NVTX ranges around the host phases ("lpp:spmv", "lpp:lanczos", ...): no cost without a tool attached; under ncu they select
// kernels by phase (ncu --nvtx --nvtx-include "lpp:spmv/")
struct NvtxRange {
	explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
	~NvtxRange() { nvtxRangePop(); }
	NvtxRange(const NvtxRange&) = delete;
	NvtxRange& operator=(const NvtxRange&) = delete;
};

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const std::string& msg)
{
	g_err = msg;
	return code;
}
#define CK(call)                                                                                               \
	do {                                                                                                       \
		cudaError_t e_ = (call);                                                                               \
		if (e_ != cudaSuccess)                                                                                 \
			return fail(LPP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
			                          std::to_string(__LINE__) + ")");                                         \
	} while (0)
#define CKR(expr)               \
	do {                        \
		int r_ = (expr);        \
		if (r_ != 0) return r_; \
	} while (0)

extern "C" const char* lpp_last_error(void) { return g_err.c_str(); }
extern "C" int lpp_version(void) { return 100; }

// ------------------------------------------------------------------ NCCL (resolved lazily; only needed for nranks>1)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
	void* lib = nullptr;
	int (*GetUniqueId)(ncclUniqueId*) = nullptr;
	int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	int (*CommDestroy)(ncclComm_t) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static const int kNcclFloat64 = 8, kNcclSum = 0;

static int nccl_load()
{
	if (g_nccl.lib) return 0;
	void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!lib) return fail(LPP_ERR_NCCL, std::string("dlopen libnccl.so.2: ") + dlerror());
#define SYM(field, name)                                                    \
	*(void**)(&g_nccl.field) = dlsym(lib, name);                            \
	if (!g_nccl.field) return fail(LPP_ERR_NCCL, std::string("dlsym ") + name)
	SYM(GetUniqueId, "ncclGetUniqueId");
	SYM(CommInitRank, "ncclCommInitRank");
	SYM(CommDestroy, "ncclCommDestroy");
	SYM(AllReduce, "ncclAllReduce");
	SYM(Broadcast, "ncclBroadcast");
	SYM(Send, "ncclSend");
	SYM(Recv, "ncclRecv");
	SYM(GroupStart, "ncclGroupStart");
	SYM(GroupEnd, "ncclGroupEnd");
	SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
	g_nccl.lib = lib;
	return 0;
}
#define CKN(call)                                                                                         \
	do {                                                                                                  \
		int e_ = (call);                                                                                  \
		if (e_ != 0) return fail(LPP_ERR_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(e_));   \
	} while (0)

// ------------------------------------------------------------------ handle
struct lpp_handle {
	lpp_desc desc;
	std::vector<double> hop, jzz, U, V, D, jpm, w;
	int device = 0;
	cudaStream_t stream = nullptr;
	ModelDev md;
	std::vector<void*> allocs;
	uint64_t rows = 0, row0 = 0, nloc = 0;
	std::vector<uint64_t> shard_row0, shard_nloc;
	// product tables
	bool tables_ready = false;
	HopTable up{}, dn{};
	DiagTables dt{};
	TiledPlan* tiled = nullptr;
	HeisBond* heis_bonds = nullptr;
	int heis_nbonds = 0, heis_field = 0;
	// CRS
	bool crs_ready = false;
	int64_t* rowptr = nullptr;
	int64_t* colind = nullptr;
	double* values = nullptr;
	int64_t nnz = 0;
	// vectors
	double* vx = nullptr;
	double* vy = nullptr;
	double* yfull = nullptr;
	double* gs = nullptr;
	double* modified = nullptr;
	double* partials = nullptr;
	int partials_cap = 0;
	int conv_index = 0;           // which Ritz value the convergence test of the Krylov loop watches (0 = lowest)
	double* scal_dev = nullptr;
	double* lz_coefs = nullptr;   // device-resident Lanczos scalars (LPP_LZ_*), then a[steps], b[steps]
	double* lz_ab = nullptr;
	int lz_ab_cap = 0;
	double* lz_ab_host = nullptr; // pinned
	double* scal_host = nullptr;
	// comm
	ncclComm_t comm = nullptr;
	bool comm_borrowed = false;    // lpp_comm_share: the communicator belongs to another handle
	bool p2p_requested = false;    // lpp_p2p_export was called: the launcher is setting up the peer-memory exchange
	// LPP_PHASES=1: CUDA-event breakdown of the sharded iteration (printed by lpp_destroy)
	cudaStream_t copy_stream2 = nullptr;
	cudaEvent_t ev_copy2 = nullptr;
	cudaStream_t copy_streams[4] = {};    // extra copy streams of the pack: the remote blocks go out on several copy engines
	cudaEvent_t ev_copies[4] = {};
	const double* packed_vec = nullptr;   // vector whose column-shard copy (ycol on every rank) is already in place
	// pipelined two-layout recurrence (lanczos_pipelined): third work vector, norm partials of the unpack in flight, streams
	double* vz = nullptr;
	double* partials3 = nullptr;
	int partials3_cap = 0;
	cudaStream_t pipe_up[2] = {};          // the up-sweep chunks alternate between two high-priority streams (the second fills the tail waves of the first)
	cudaStream_t pipe_unpack = nullptr;    // the unpack chunks (low priority: they fill in beside the up sweep)
	cudaEvent_t ev_pipe_chunk[8] = {};     // unpack of chunk k has written its rows of the new vector
	cudaEvent_t ev_pipe_dot = nullptr, ev_pipe_up[2] = {}, ev_pipe_y = nullptr;
	int phases = -1;
	cudaEvent_t pev[8] = {};
	cudaEvent_t cev[2] = {};      // around the pack on the second stream
	double pack_ms = 0;
	double phase_ms[8] = {};
	long phase_n = 0;
	// two-layout sharding (product bases without two-spin terms): the up sweep runs on the ROW shard, the down sweep on the
	// COLUMN shard, linked by two all-to-all transposes per mat-vec (SURVEY §8e alternative A)
	int two_layout = -1;          // -1 undecided, 0 no, 1 yes
	ColSplit cols{};
	std::vector<uint64_t> dstart; // down-index range of every rank (nranks+1)
	uint64_t ucol0 = 0, ncols = 0;
	double* ycol = nullptr;
	double* xcol = nullptr;
	double* sendbuf = nullptr;
	double* recvbuf = nullptr;
	double* partials2 = nullptr;
	int partials2_cap = 0;
	int p2p = 0;                  // 1: peers' column shards are mapped (CUDA IPC): exchange by our own kernels over NVLink
	bool pull_pack = false;       // every rank's work vectors are mapped as well: the pack is a pull, all-reduce #1 is not needed
	PeerPtrs peer_ycol{}, peer_xcol{};
	PeerPtrs peer_vx{}, peer_vy{};  // peers' Lanczos work vectors (same slab): the pack PULLS the columns it owns from them
	double* slab = nullptr;       // one allocation [ycol | xcol | vx | vy]: one CUDA IPC handle covers all four
	uint64_t slab_off_xcol = 0, slab_off_vx = 0, slab_off_vy = 0;   // offsets in doubles; vx/vy = 0: not in the slab
	uint64_t slab_off_psx = 0;    // scalar exchange area (lpp_launch_psx_allreduce)
	PeerPtrs peer_psx{};
	bool psx = false;             // scalar all-reduces of the peer-memory path go over NVLink stores + flags instead of NCCL
	unsigned long long psx_seq = 0;
	int* psx_err = nullptr;
	std::vector<void*> ipc_opened;
	cudaStream_t comm_stream = nullptr;
	cudaEvent_t ev_pack = nullptr, ev_ycol = nullptr, ev_xcol = nullptr, ev_recv = nullptr, ev_scal = nullptr;
	// stats
	int64_t launches = 0;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

template <class T>
static int dev_alloc(lpp_handle* h, T** p, size_t count)
{
	void* q = nullptr;
	CK(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
	h->allocs.push_back(q);
	*p = (T*)q;
	return 0;
}

template <class T>
static int dev_upload(lpp_handle* h, T** p, const T* src, size_t count)
{
	CKR(dev_alloc(h, p, count));
	if (count) CK(cudaMemcpy(*p, src, count * sizeof(T), cudaMemcpyHostToDevice));
	return 0;
}

static void dev_free(lpp_handle* h, void* p)
{
	if (!p) return;
	auto it = std::find(h->allocs.begin(), h->allocs.end(), p);
	if (it != h->allocs.end()) h->allocs.erase(it);
	cudaFree(p);
}

// call-scoped device buffer: freed on every way out of the entry point, the CK() error returns included
template <class T>
struct ScopedDev {
	T* p = nullptr;
	ScopedDev() = default;
	ScopedDev(const ScopedDev&) = delete;
	ScopedDev& operator=(const ScopedDev&) = delete;
	~ScopedDev() { if (p) cudaFree(p); }
	cudaError_t alloc(size_t count) { return cudaMalloc((void**)&p, sizeof(T) * std::max<size_t>(count, 1)); }
	operator T*() const { return p; }
};

extern "C" int lpp_device_check(int32_t device)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0)
		return fail(LPP_ERR_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "count=0") +
		                              " (liblpp_b200 has no CPU fallback)");
	if (device < 0 || device >= n) return fail(LPP_ERR_ARG, "device ordinal out of range");
	cudaDeviceProp p;
	CK(cudaGetDeviceProperties(&p, device));
	if (p.major != 10) return fail(LPP_ERR_CUDA, std::string("device is sm_") + std::to_string(p.major * 10 + p.minor) +
	                                                 ", kernels are built for sm_100a only");
	return 0;
}

static int setup_feas_spin(lpp_handle* h, int spin, const std::vector<uint64_t>& binom, uint64_t* nout)
{
	const int npart = spin ? h->desc.ndown : h->desc.nup;
	LppFeasLayout L = lpp_feas_layout(binom, h->desc.nsite, h->desc.orbitals, npart);
	uint64_t* d_off; uint64_t* d_start; int* d_pn;
	CKR(dev_upload(h, &d_off, L.off.data(), L.off.size()));
	CKR(dev_upload(h, &d_start, L.start.data(), L.start.size()));
	CKR(dev_upload(h, &d_pn, L.pn.data(), L.pn.size()));
	if (spin == 0) { h->md.part_off1 = d_off; h->md.part_start1 = d_start; h->md.part_n1 = d_pn; h->md.nparts1 = (int)L.start.size() - 1; }
	else { h->md.part_off2 = d_off; h->md.part_start2 = d_start; h->md.part_n2 = d_pn; h->md.nparts2 = (int)L.start.size() - 1; }
	*nout = L.total;
	return 0;
}

static void shard_range(uint64_t n, int rank, int nranks, uint64_t* first, uint64_t* count)
{
	uint64_t base = n / nranks, rem = n % nranks;
	*first = base * rank + std::min<uint64_t>(rank, rem);
	*count = base + ((uint64_t)rank < rem ? 1 : 0);
}

extern "C" int lpp_shard_range(uint64_t n, int32_t rank, int32_t nranks, uint64_t* first, uint64_t* count)
{
	if (nranks < 1 || rank < 0 || rank >= nranks) return fail(LPP_ERR_ARG, "bad rank/nranks");
	shard_range(n, rank, nranks, first, count);
	return 0;
}

extern "C" int lpp_destroy(lpp_handle* h)
{
	if (!h) return 0;
	cudaSetDevice(h->device);
	if (h->phases == 1 && h->phase_n > 0) {
		static const char* nm[6] = {"up sweep", "wait pack + allreduce1", "down sweep", "finalize + allreduce2",
		                            "host trip + unpack/axpy/norm", "reduce + allreduce3"};
		fprintf(stderr, "[lpp phases] rank %d, %ld iterations:", h->desc.rank, h->phase_n);
		for (int k = 0; k < 6; k++) fprintf(stderr, " %s %.3f ms |", nm[k], h->phase_ms[k] / h->phase_n);
		fprintf(stderr, " [pack on the second stream %.3f ms]\n", h->pack_ms / h->phase_n);
		for (int i = 0; i < 8; i++) if (h->pev[i]) cudaEventDestroy(h->pev[i]);
	}
	for (void* q : h->ipc_opened) cudaIpcCloseMemHandle(q);
	if (h->comm && !h->comm_borrowed && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
	if (h->tiled) lpp_tiled_destroy(h->tiled);
	for (void* p : h->allocs) cudaFree(p);
	if (h->scal_host) cudaFreeHost(h->scal_host);
	if (h->lz_ab_host) cudaFreeHost(h->lz_ab_host);
	if (h->psx_err) cudaFreeHost(h->psx_err);
	if (h->ev0) cudaEventDestroy(h->ev0);
	if (h->ev1) cudaEventDestroy(h->ev1);
	for (cudaEvent_t e : {h->ev_pack, h->ev_ycol, h->ev_xcol, h->ev_recv, h->ev_scal}) if (e) cudaEventDestroy(e);
	if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
	if (h->copy_stream2) cudaStreamDestroy(h->copy_stream2);
	if (h->ev_copy2) cudaEventDestroy(h->ev_copy2);
	for (int k = 0; k < 4; k++) {
		if (h->copy_streams[k]) cudaStreamDestroy(h->copy_streams[k]);
		if (h->ev_copies[k]) cudaEventDestroy(h->ev_copies[k]);
	}
	for (cudaStream_t st : h->pipe_up) if (st) cudaStreamDestroy(st);
	if (h->pipe_unpack) cudaStreamDestroy(h->pipe_unpack);
	for (cudaEvent_t e : h->ev_pipe_chunk) if (e) cudaEventDestroy(e);
	for (cudaEvent_t e : {h->ev_pipe_dot, h->ev_pipe_up[0], h->ev_pipe_up[1], h->ev_pipe_y}) if (e) cudaEventDestroy(e);
	if (h->stream) cudaStreamDestroy(h->stream);
	delete h;
	return 0;
}

static int create_impl(const lpp_desc* d, lpp_handle* h)
{
	h->desc = *d;
	const int model = d->model, nsite = d->nsite;
	const int no = (model == LPP_MODEL_FEAS) ? d->orbitals : 1;
	h->desc.orbitals = no;
	const int nb = nsite * no;
	if (model < 0 || model > 3) return fail(LPP_ERR_ARG, "unknown model");
	if (model == LPP_MODEL_TJ) {
		if (d->orbitals > 1) return fail(LPP_ERR_ARG, "t-J: Orbitals=1 only (Tj1Orbital)");
		if (nsite > 31) return fail(LPP_ERR_ARG, "t-J: at most 31 sites (combined word down << nsite | up)");
		if (d->nup < 0 || d->ndown < 0 || d->nup + d->ndown > nsite) return fail(LPP_ERR_ARG, "t-J: nup + ndown must not exceed the number of sites");
		if (d->nV != 0 && d->nV != 2 * nsite) return fail(LPP_ERR_ARG, "t-J: potentialV holds 2*nsite values (or none)");
		if (d->nranks != 1) return fail(LPP_ERR_ARG, "t-J: single GPU only");
	}
	if (nsite < 1 || nb > 62) return fail(LPP_ERR_ARG, "nsite*orbitals must be in [1,62]");
	if (no < 1 || no > LPP_MAX_ORB) return fail(LPP_ERR_ARG, "orbitals must be in [1,4]");
	if (d->nup < 0 || d->nup > nb || (model != LPP_MODEL_HEISENBERG && (d->ndown < 0 || d->ndown > nb)))
		return fail(LPP_ERR_ARG, "particle numbers out of range");
	if (!d->hop) return fail(LPP_ERR_ARG, "hop matrix is required");
	if (d->nranks < 1 || d->rank < 0 || d->rank >= d->nranks) return fail(LPP_ERR_ARG, "bad rank/nranks");
	CKR(lpp_device_check(d->device));
	h->device = d->device;
	CK(cudaSetDevice(h->device));
	CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
	CK(cudaEventCreate(&h->ev0));
	CK(cudaEventCreate(&h->ev1));

	h->hop.assign(d->hop, d->hop + (size_t)nb * nb);
	h->jzz.assign((size_t)nb * nb, 0.0);
	if (d->jzz) h->jzz.assign(d->jzz, d->jzz + (size_t)nb * nb);
	const int needU = (model == LPP_MODEL_FEAS) ? 6 : nsite;
	h->U.assign(needU, 0.0);
	if (d->U) for (int i = 0; i < std::min(needU, d->nU); i++) h->U[i] = d->U[i];
	if (model == LPP_MODEL_FEAS) {
		if (d->nU < 4 || d->nU > 6) return fail(LPP_ERR_ARG, "FeAsMode INT_PAPER33 expects 4, 5 or 6 U values");
		if (d->nU == 4 || d->nU == 5) { h->U[4] = h->U[2]; h->U[5] = 0.0; }  // ParametersModelFeAs.h:147-151
	}
	h->jpm.assign((size_t)nb * nb, 0.0);
	h->w.assign((size_t)nb * nb, 0.0);
	if (model == LPP_MODEL_TJ && d->jpm) h->jpm.assign(d->jpm, d->jpm + (size_t)nb * nb);
	if (model == LPP_MODEL_TJ && d->w) h->w.assign(d->w, d->w + (size_t)nb * nb);
	const int needV = (model == LPP_MODEL_FEAS) ? 2 * no * nsite : (model == LPP_MODEL_TJ ? 2 * nsite : nsite);
	h->V.assign(needV, 0.0);
	if (d->V) for (int i = 0; i < std::min(needV, d->nV); i++) h->V[i] = d->V[i];
	h->D.assign(std::max(nsite, 1), 0.0);
	if (d->D) for (int i = 0; i < std::min(nsite, d->nD); i++) h->D[i] = d->D[i];

	std::vector<uint64_t> binom = lpp_make_binom();
	ModelDev& m = h->md;
	memset(&m, 0, sizeof(m));
	m.model = model; m.nsite = nsite; m.orbitals = no; m.nbits = nb;
	m.nup = d->nup; m.ndn = d->ndown; m.u3_all_pairs = d->feas_u3_all_pairs;
	uint64_t* d_binom;
	CKR(dev_upload(h, &d_binom, binom.data(), binom.size()));
	m.binom = d_binom;
	double* p;
	CKR(dev_upload(h, &p, h->hop.data(), h->hop.size())); m.hop = p;
	CKR(dev_upload(h, &p, h->jzz.data(), h->jzz.size())); m.jzz = p;
	CKR(dev_upload(h, &p, h->jpm.data(), h->jpm.size())); m.jpm = p;
	CKR(dev_upload(h, &p, h->w.data(), h->w.size())); m.w = p;
	CKR(dev_upload(h, &p, h->U.data(), h->U.size())); m.U = p;
	CKR(dev_upload(h, &p, h->V.data(), h->V.size())); m.V = p;
	CKR(dev_upload(h, &p, h->D.data(), h->D.size())); m.D = p;

	// bases on device
	word_t* b1 = nullptr; word_t* b2 = nullptr;
	if (model == LPP_MODEL_FEAS) {
		CKR(setup_feas_spin(h, 0, binom, &m.n1));
		CKR(setup_feas_spin(h, 1, binom, &m.n2));
		CKR(dev_alloc(h, &b1, m.n1));
		CKR(dev_alloc(h, &b2, m.n2));
		lpp_launch_build_feas(m, 0, m.n1, b1, h->stream);
		lpp_launch_build_feas(m, 1, m.n2, b2, h->stream);
	} else if (model == LPP_MODEL_TJ) {
		// fast index: up words compressed onto the nsite - ndown sites the down word leaves free; slow index: down words
		m.n1 = binom[(nsite - d->ndown) * LPP_BINOM_N + d->nup];
		m.n2 = binom[nsite * LPP_BINOM_N + d->ndown];
		CKR(dev_alloc(h, &b1, m.n1));
		CKR(dev_alloc(h, &b2, m.n2));
		lpp_launch_build_colex(m.binom, nsite - d->ndown, d->nup, m.n1, b1, h->stream);
		lpp_launch_build_colex(m.binom, nsite, d->ndown, m.n2, b2, h->stream);
	} else {
		m.n1 = binom[nsite * LPP_BINOM_N + d->nup];
		m.n2 = (model == LPP_MODEL_HUBBARD) ? binom[nsite * LPP_BINOM_N + d->ndown] : 1;
		CKR(dev_alloc(h, &b1, m.n1));
		CKR(dev_alloc(h, &b2, m.n2));
		lpp_launch_build_colex(m.binom, nsite, d->nup, m.n1, b1, h->stream);
		if (model == LPP_MODEL_HUBBARD) lpp_launch_build_colex(m.binom, nsite, d->ndown, m.n2, b2, h->stream);
		else CK(cudaMemsetAsync(b2, 0, sizeof(word_t), h->stream));
	}
	m.b1 = b1; m.b2 = b2;
	m.rows = m.n1 * m.n2;
	h->rows = m.rows;
	CK(cudaGetLastError());

	// rank accelerators
	if (model != LPP_MODEL_FEAS && nb <= 40 && m.n1 < (1ull << 32) && m.n2 < (1ull << 32)) {
		int lobits = (nb + 1) / 2;
		uint32_t* rlo; uint32_t* rhi;
		CKR(dev_alloc(h, &rlo, (size_t)1 << lobits));
		CKR(dev_alloc(h, &rhi, (size_t)(lobits + 1) << (nb - lobits)));
		lpp_launch_split_tables(m.binom, nb, lobits, rlo, rhi, h->stream);
		CK(cudaStreamSynchronize(h->stream));
		m.rlo = rlo; m.rhi = rhi; m.lobits = lobits;
	}
	if (model == LPP_MODEL_FEAS && nb <= 26) {
		uint32_t* l1; uint32_t* l2;
		CKR(dev_alloc(h, &l1, (size_t)1 << nb));
		CKR(dev_alloc(h, &l2, (size_t)1 << nb));
		CK(cudaMemsetAsync(l1, 0xff, sizeof(uint32_t) << nb, h->stream));
		CK(cudaMemsetAsync(l2, 0xff, sizeof(uint32_t) << nb, h->stream));
		lpp_launch_lut(m.b1, m.n1, l1, h->stream);
		lpp_launch_lut(m.b2, m.n2, l2, h->stream);
		CK(cudaStreamSynchronize(h->stream));
		m.lut1 = l1; m.lut2 = l2;
	}

	// sharding over the slow index (product bases: whole up-segments; Heisenberg: contiguous rows)
	h->shard_row0.resize(d->nranks);
	h->shard_nloc.resize(d->nranks);
	for (int r = 0; r < d->nranks; r++) {
		uint64_t f, c;
		if (model == LPP_MODEL_HEISENBERG) shard_range(m.rows, r, d->nranks, &f, &c);
		else { shard_range(m.n2, r, d->nranks, &f, &c); f *= m.n1; c *= m.n1; }
		h->shard_row0[r] = f;
		h->shard_nloc[r] = c;
	}
	h->row0 = h->shard_row0[d->rank];
	h->nloc = h->shard_nloc[d->rank];

	CKR(dev_alloc(h, &h->scal_dev, 8));
	CK(cudaMallocHost((void**)&h->scal_host, 8 * sizeof(double)));
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaGetLastError());
	return 0;
}

extern "C" int lpp_create(const lpp_desc* d, lpp_handle** out)
{
	if (!d || !out) return fail(LPP_ERR_ARG, "null argument");
	lpp_handle* h = new lpp_handle();
	int rc = create_impl(d, h);
	if (rc != 0) {
		std::string keep = g_err;
		lpp_destroy(h);
		g_err = keep;
		*out = nullptr;
		return rc;
	}
	*out = h;
	return 0;
}

extern "C" int lpp_rows(const lpp_handle* h, uint64_t* rows)
{
	if (!h || !rows) return fail(LPP_ERR_ARG, "null argument");
	*rows = h->rows;
	return 0;
}

extern "C" int lpp_local_rows(const lpp_handle* h, uint64_t* first, uint64_t* count)
{
	if (!h) return fail(LPP_ERR_ARG, "null argument");
	if (first) *first = h->row0;
	if (count) *count = h->nloc;
	return 0;
}

extern "C" int lpp_basis_size(const lpp_handle* h, int32_t spin, uint64_t* n)
{
	if (!h || !n) return fail(LPP_ERR_ARG, "null argument");
	*n = spin ? h->md.n2 : h->md.n1;
	return 0;
}

extern "C" int lpp_basis_export(const lpp_handle* h, int32_t spin, uint64_t* words)
{
	if (!h || !words) return fail(LPP_ERR_ARG, "null argument");
	CK(cudaSetDevice(h->device));
	CK(cudaMemcpy(words, spin ? h->md.b2 : h->md.b1, sizeof(word_t) * (spin ? h->md.n2 : h->md.n1), cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int lpp_rank(const lpp_handle* hc, int32_t spin, const uint64_t* words, uint64_t n, uint64_t* index)
{
	lpp_handle* h = const_cast<lpp_handle*>(hc);
	if (!h || !words || !index) return fail(LPP_ERR_ARG, "null argument");
	CK(cudaSetDevice(h->device));
	ScopedDev<word_t> dw; ScopedDev<uint64_t> di;
	CK(dw.alloc(n));
	CK(di.alloc(n));
	CK(cudaMemcpy(dw, words, sizeof(word_t) * n, cudaMemcpyHostToDevice));
	lpp_launch_rank(h->md, spin, dw, n, di, h->stream);
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaMemcpy(index, di, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int lpp_row_words(const lpp_handle* hc, uint64_t first, uint64_t count, uint64_t* up_words, uint64_t* down_words)
{
	lpp_handle* h = const_cast<lpp_handle*>(hc);
	if (!h) return fail(LPP_ERR_ARG, "null argument");
	if (first > h->rows || count > h->rows - first) return fail(LPP_ERR_ARG, "row range outside the basis");
	CK(cudaSetDevice(h->device));
	ScopedDev<word_t> du, dd;
	if (up_words) CK(du.alloc(count));
	if (down_words) CK(dd.alloc(count));
	lpp_launch_row_words(h->md, first, count, du, dd, h->stream);
	CK(cudaStreamSynchronize(h->stream));
	if (du) CK(cudaMemcpy(up_words, du, sizeof(word_t) * count, cudaMemcpyDeviceToHost));
	if (dd) CK(cudaMemcpy(down_words, dd, sizeof(word_t) * count, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int lpp_rank_pairs(const lpp_handle* hc, const uint64_t* up_words, const uint64_t* down_words, uint64_t n, uint64_t* index)
{
	lpp_handle* h = const_cast<lpp_handle*>(hc);
	if (!h || !up_words || !down_words || !index) return fail(LPP_ERR_ARG, "null argument");
	CK(cudaSetDevice(h->device));
	ScopedDev<word_t> du, dd; ScopedDev<uint64_t> di;
	CK(du.alloc(n));
	CK(dd.alloc(n));
	CK(di.alloc(n));
	CK(cudaMemcpy(du, up_words, sizeof(word_t) * n, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dd, down_words, sizeof(word_t) * n, cudaMemcpyHostToDevice));
	lpp_launch_rank_pairs(h->md, du, dd, n, di, h->stream);
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaMemcpy(index, di, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost));
	return 0;
}

// ------------------------------------------------------------------ tables / CRS
static int ensure_tables(lpp_handle* h)
{
	if (h->tables_ready) return 0;
	if (h->md.model == LPP_MODEL_HEISENBERG || h->md.model == LPP_MODEL_TJ) return fail(LPP_ERR_ARG, "hop tables exist for product bases only");
	const ModelDev& m = h->md;
	for (int spin = 0; spin < 2; spin++) {
		HopTable& t = spin ? h->dn : h->up;
		t.n = spin ? m.n2 : m.n1;
		CKR(dev_alloc(h, &t.cnt, t.n));
		uint32_t* dmax = (uint32_t*)h->scal_dev;
		CK(cudaMemsetAsync(dmax, 0, sizeof(uint32_t), h->stream));
		lpp_launch_hop_count(m, spin, t.n, t.cnt, dmax, h->stream);
		uint32_t hmax = 0;
		CK(cudaMemcpyAsync(&hmax, dmax, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
		CK(cudaStreamSynchronize(h->stream));
		t.width = (int)hmax;
		CKR(dev_alloc(h, &t.idx, (size_t)t.width * t.n));
		CKR(dev_alloc(h, &t.val, (size_t)t.width * t.n));
		lpp_launch_hop_fill(m, spin, t, h->stream);
		double* dv;
		CKR(dev_alloc(h, &dv, t.n));
		lpp_launch_spin_diag(m, spin, t.n, dv, h->stream);
		if (spin) h->dt.dv2 = dv; else h->dt.dv1 = dv;
	}
	h->dt.uniformU = 1;
	h->dt.U0 = h->U[0];
	if (m.model == LPP_MODEL_HUBBARD)
		for (int i = 1; i < m.nsite; i++) if (h->U[i] != h->U[0]) h->dt.uniformU = 0;
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaGetLastError());
	h->tables_ready = true;
	return 0;
}

extern "C" int lpp_crs_build(lpp_handle* h, int64_t* nnz)
{
	if (!h) return fail(LPP_ERR_ARG, "null argument");
	CK(cudaSetDevice(h->device));
	if (!h->crs_ready) {
		CKR(dev_alloc(h, &h->rowptr, h->nloc + 1));
		int* dflag = (int*)h->scal_dev;
		CK(cudaMemsetAsync(dflag, 0, sizeof(int), h->stream));
		CK(cudaMemsetAsync(h->rowptr, 0, sizeof(int64_t) * (h->nloc + 1), h->stream));
		lpp_launch_crs_count(h->md, h->row0, h->nloc, h->rowptr, dflag, h->stream);
		int flag = 0;
		CK(cudaMemcpyAsync(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
		lpp_exclusive_scan(h->rowptr, h->nloc, nullptr, h->stream);
		CK(cudaMemcpyAsync(&h->nnz, h->rowptr + h->nloc, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
		CK(cudaStreamSynchronize(h->stream));
		if (flag) return fail(LPP_ERR_OVERFLOW, "a Hamiltonian row has more than 256 entries");
		CKR(dev_alloc(h, &h->colind, (size_t)h->nnz));
		CKR(dev_alloc(h, &h->values, (size_t)h->nnz));
		lpp_launch_crs_fill(h->md, h->row0, h->nloc, h->rowptr, h->colind, h->values, h->stream);
		CK(cudaStreamSynchronize(h->stream));
		CK(cudaGetLastError());
		h->crs_ready = true;
	}
	if (nnz) *nnz = h->nnz;
	return 0;
}

extern "C" int lpp_crs_export(const lpp_handle* h, int64_t* rowptr, int64_t* colind, double* values)
{
	if (!h) return fail(LPP_ERR_ARG, "null argument");
	if (!h->crs_ready) return fail(LPP_ERR_STATE, "lpp_crs_build has not been called");
	CK(cudaSetDevice(h->device));
	if (rowptr) CK(cudaMemcpy(rowptr, h->rowptr, sizeof(int64_t) * (h->nloc + 1), cudaMemcpyDeviceToHost));
	if (colind) CK(cudaMemcpy(colind, h->colind, sizeof(int64_t) * h->nnz, cudaMemcpyDeviceToHost));
	if (values) CK(cudaMemcpy(values, h->values, sizeof(double) * h->nnz, cudaMemcpyDeviceToHost));
	return 0;
}

// ------------------------------------------------------------------ SpMV dispatch
static int resolve_kernel(const lpp_handle* h, int kernel)
{
	if (h->md.model == LPP_MODEL_TJ) return (kernel == LPP_KERNEL_STORED) ? LPP_KERNEL_STORED : LPP_KERNEL_GENERIC;   // not a product basis
	if (kernel == LPP_KERNEL_AUTO) return (h->md.model == LPP_MODEL_HEISENBERG) ? LPP_KERNEL_TABLE : LPP_KERNEL_TILED;
	if (h->md.model == LPP_MODEL_HEISENBERG && kernel == LPP_KERNEL_TILED) return LPP_KERNEL_TABLE;
	return kernel;
}

static int ensure_partials(lpp_handle* h, int n)
{
	if (n <= h->partials_cap) return 0;
	if (h->partials) dev_free(h, h->partials);
	h->partials = nullptr;
	CKR(dev_alloc(h, &h->partials, (size_t)n));
	h->partials_cap = n;
	return 0;
}

static int ensure_tiled(lpp_handle* h)
{
	CKR(ensure_tables(h));
	if (!h->tiled) {
		int rc = lpp_tiled_create(h->md, h->hop.data(), h->up, h->dn, h->dt, h->row0, h->nloc, h->stream, &h->tiled);
		if (rc != 0) return fail(LPP_ERR_CUDA, std::string("tiled plan: ") + lpp_tiled_error());
	}
	return 0;
}

// x = beta x + alpha H y ; if want_dot, returns in *npartials the number of block partial sums left in h->partials
static int do_spmv(lpp_handle* h, int kernel, double alpha, double beta, double* x, const double* y, bool want_dot,
                   int* npartials, const double* coefs_dev = nullptr)
{
	NvtxRange nvtx("lpp:spmv");
	kernel = resolve_kernel(h, kernel);
	SpmvArgs a;
	a.alpha = alpha; a.beta = beta; a.x = x; a.y = y; a.row0 = h->row0; a.nloc = h->nloc;
	if (coefs_dev) { a.alpha.p = coefs_dev + LPP_LZ_ALPHA; a.beta.p = coefs_dev + LPP_LZ_BETA; }
	a.dot_partials = nullptr;
	int nb = 0;
	if (kernel == LPP_KERNEL_GENERIC) {
		nb = lpp_spmv_generic_blocks(h->nloc);
		if (want_dot) { CKR(ensure_partials(h, nb)); a.dot_partials = h->partials; }
		lpp_launch_spmv_generic(h->md, a, h->stream);
		h->launches += 1;
	} else if (kernel == LPP_KERNEL_TABLE && h->md.model == LPP_MODEL_HEISENBERG) {
		if (!h->heis_bonds) {
			std::vector<HeisBond> bl;
			const int n = h->md.nsite;
			for (int p = 0; p < n; p++)
				for (int q = p + 1; q < n; q++) {
					HeisBond b{p, q, h->hop[(size_t)p * n + q], h->hop[(size_t)q * n + p], h->jzz[(size_t)p * n + q]};
					if (b.jpm_pq != 0 || b.jpm_qp != 0 || b.jzz != 0) bl.push_back(b);
				}
			h->heis_nbonds = (int)bl.size();
			CKR(dev_upload(h, &h->heis_bonds, bl.data(), bl.size()));
			h->heis_field = 0;
			for (int i = 0; i < n; i++) if (h->V[i] != 0 || h->D[i] != 0) h->heis_field = 1;
		}
		nb = lpp_spmv_generic_blocks(h->nloc);
		if (want_dot) { CKR(ensure_partials(h, nb)); a.dot_partials = h->partials; }
		lpp_launch_spmv_heis(h->md, h->heis_bonds, h->heis_nbonds, h->heis_field, a, h->stream);
		h->launches += 1;
	} else if (kernel == LPP_KERNEL_TABLE) {
		CKR(ensure_tables(h));
		nb = lpp_spmv_table_blocks(h->md, h->nloc);
		if (want_dot) { CKR(ensure_partials(h, nb)); a.dot_partials = h->partials; }
		lpp_launch_spmv_table(h->md, h->up, h->dn, h->dt, a, h->stream);
		h->launches += 1;
	} else if (kernel == LPP_KERNEL_TILED) {
		CKR(ensure_tiled(h));
		nb = lpp_tiled_dot_blocks(h->tiled);
		if (want_dot) { CKR(ensure_partials(h, nb)); a.dot_partials = h->partials; }
		int nl = lpp_tiled_spmv(h->tiled, h->md, h->up, h->dn, h->dt, a, h->stream);
		if (nl < 0) return fail(LPP_ERR_CUDA, std::string("tiled spmv: ") + lpp_tiled_error());
		h->launches += nl;
	} else if (kernel == LPP_KERNEL_STORED) {
		CKR(lpp_crs_build(h, nullptr));
		nb = lpp_spmv_crs_blocks(h->nloc);
		if (want_dot) { CKR(ensure_partials(h, nb)); a.dot_partials = h->partials; }
		lpp_launch_spmv_crs(h->rowptr, h->colind, h->values, a, h->stream);
		h->launches += 1;
	} else {
		return fail(LPP_ERR_ARG, "unknown kernel id");
	}
	CK(cudaGetLastError());
	if (npartials) *npartials = nb;
	return 0;
}

extern "C" int lpp_matvec_device(lpp_handle* h, int32_t kernel, double* x_dev, const double* y_dev)
{
	if (!h || !x_dev || !y_dev) return fail(LPP_ERR_ARG, "null argument");
	CK(cudaSetDevice(h->device));
	CKR(do_spmv(h, kernel, 1.0, 1.0, x_dev, y_dev, false, nullptr));
	CK(cudaStreamSynchronize(h->stream));
	return 0;
}

extern "C" int lpp_matvec_host(lpp_handle* h, int32_t kernel, double* x, const double* y)
{
	if (!h || !x || !y) return fail(LPP_ERR_ARG, "null argument");
	if (h->desc.nranks != 1) return fail(LPP_ERR_STATE, "lpp_matvec_host needs the whole Hilbert space on one GPU");
	CK(cudaSetDevice(h->device));
	ScopedDev<double> dx, dy;
	CK(dx.alloc(h->rows));
	CK(dy.alloc(h->rows));
	CK(cudaMemcpyAsync(dx, x, sizeof(double) * h->rows, cudaMemcpyHostToDevice, h->stream));
	CK(cudaMemcpyAsync(dy, y, sizeof(double) * h->rows, cudaMemcpyHostToDevice, h->stream));
	int rc = do_spmv(h, kernel, 1.0, 1.0, dx, dy, false, nullptr);
	if (rc == 0) {
		cudaError_t e = cudaMemcpyAsync(x, dx, sizeof(double) * h->rows, cudaMemcpyDeviceToHost, h->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
		if (e != cudaSuccess) rc = fail(LPP_ERR_CUDA, std::string("matvec_host: ") + cudaGetErrorString(e));
	} else {
		cudaStreamSynchronize(h->stream);      // nothing of this call is left in flight when the buffers go
	}
	return rc;
}

// ------------------------------------------------------------------ tridiagonal eigen-solvers (host)
// lowest eigenvalue by Sturm-sequence bisection (used every Lanczos step for the convergence test)
static double tridiag_kth(int n, const double* a, const double* b, int k);
static double tridiag_lowest(int n, const double* a, const double* b) { return tridiag_kth(n, a, b, 0); }

// k-th lowest eigenvalue (k = 0 .. n-1) of the symmetric tridiagonal (a, b) by Sturm-sequence bisection
static double tridiag_kth(int n, const double* a, const double* b, int k)
{
	if (n == 1) return a[0];
	double lo = a[0], hi = a[0];
	for (int i = 0; i < n; i++) {
		double r = (i > 0 ? fabs(b[i - 1]) : 0.0) + (i < n - 1 ? fabs(b[i]) : 0.0);
		lo = std::min(lo, a[i] - r);
		hi = std::max(hi, a[i] + r);
	}
	auto count_below = [&](double x) {  // number of eigenvalues < x
		int c = 0;
		double q = a[0] - x;
		if (q < 0) c++;
		for (int i = 1; i < n; i++) {
			double bb = b[i - 1] * b[i - 1];
			if (q == 0.0) q = 1e-300;
			q = a[i] - x - bb / q;
			if (q < 0) c++;
		}
		return c;
	};
	for (int it = 0; it < 200; it++) {
		double mid = 0.5 * (lo + hi);
		if (mid <= lo || mid >= hi) break;
		if (count_below(mid) >= k + 1) hi = mid; else lo = mid;
	}
	return 0.5 * (lo + hi);
}

// full decomposition: implicit-shift QL on (d, e); z (n*n, z[i*n+k] = component i of vector k) optional
static int tridiag_full(int n, std::vector<double>& d, std::vector<double>& e, double* z)
{
	e.resize(n + 1);
	e[n - 1] = 0.0;
	if (z) {
		std::fill(z, z + (size_t)n * n, 0.0);
		for (int i = 0; i < n; i++) z[(size_t)i * n + i] = 1.0;
	}
	const double eps = 2.220446049250313e-16;
	for (int l = 0; l < n; l++) {
		for (int iter = 0;; iter++) {
			int m = l;
			for (; m < n - 1; m++) {
				double dd = fabs(d[m]) + fabs(d[m + 1]);
				if (fabs(e[m]) <= eps * dd) break;
			}
			if (m == l) break;
			if (iter >= 300) return -1;
			// Wilkinson shift from the leading 2x2 of the unreduced block
			double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
			double r = std::hypot(g, 1.0);
			g = d[m] - d[l] + e[l] / (g + std::copysign(r, g));
			double s = 1.0, c = 1.0, p = 0.0;
			bool underflow = false;
			for (int i = m - 1; i >= l; i--) {
				double f = s * e[i], bq = c * e[i];
				r = std::hypot(f, g);
				e[i + 1] = r;
				if (r == 0.0) {
					d[i + 1] -= p;
					e[m] = 0.0;
					underflow = true;
					break;
				}
				s = f / r;
				c = g / r;
				g = d[i + 1] - p;
				r = (d[i] - g) * s + 2.0 * c * bq;
				p = s * r;
				d[i + 1] = g + p;
				g = c * r - bq;
				if (z) {
					for (int k = 0; k < n; k++) {
						double zk1 = z[(size_t)k * n + i + 1], zk0 = z[(size_t)k * n + i];
						z[(size_t)k * n + i + 1] = s * zk0 + c * zk1;
						z[(size_t)k * n + i] = c * zk0 - s * zk1;
					}
				}
			}
			if (underflow) continue;
			d[l] -= p;
			e[l] = g;
			e[m] = 0.0;
		}
	}
	// ascending order
	std::vector<int> perm(n);
	for (int i = 0; i < n; i++) perm[i] = i;
	std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) { return d[x] < d[y]; });
	std::vector<double> d2(n);
	for (int i = 0; i < n; i++) d2[i] = d[perm[i]];
	if (z) {
		std::vector<double> z2((size_t)n * n);
		for (int i = 0; i < n; i++)
			for (int k = 0; k < n; k++) z2[(size_t)i * n + k] = z[(size_t)i * n + perm[k]];
		std::copy(z2.begin(), z2.end(), z);
	}
	d = d2;
	return 0;
}

extern "C" int lpp_tridiag_eig(int32_t n, const double* a, const double* b, double* eigs, double* z)
{
	if (n < 1 || !a || !eigs) return fail(LPP_ERR_ARG, "bad argument");
	std::vector<double> d(a, a + n), e(n, 0.0);
	for (int i = 0; i + 1 < n; i++) e[i] = b[i];
	if (tridiag_full(n, d, e, z) != 0) return fail(LPP_ERR_STATE, "tridiagonal QL did not converge");
	std::copy(d.begin(), d.end(), eigs);
	return 0;
}

extern "C" int lpp_cf_eval(int32_t n, const double* a, const double* b, double eg, double weight, int32_t isign,
                           int32_t nomega, const double* omega, double delta, double* out)
{
	if (n < 1 || !a || !b || !omega || !out) return fail(LPP_ERR_ARG, "bad argument");
	std::vector<double> d(a, a + n), e(n, 0.0), z((size_t)n * n);
	for (int i = 0; i + 1 < n; i++) e[i] = b[i];
	if (tridiag_full(n, d, e, z.data()) != 0) return fail(LPP_ERR_STATE, "tridiagonal QL did not converge");
	for (int w = 0; w < nomega; w++) {
		double re = 0, im = 0;
		for (int l = 0; l < n; l++) {
			double inten = z[l] * z[l];  // first component of eigenvector l
			double xr = omega[w] - isign * (d[l] - eg);
			double den = xr * xr + delta * delta;
			re += weight * inten * xr / den;
			im -= weight * inten * delta / den;
		}
		out[2 * w] = re;
		out[2 * w + 1] = im;
	}
	return 0;
}

// ------------------------------------------------------------------ Krylov loop
static int ensure_tiled(lpp_handle* h);
static int ensure_two_layout(lpp_handle* h, int kernel);

// sum of `n` doubles at `vals` (device) over the ranks, result in place: peer-memory exchange when the handle has it, else NCCL
static int allreduce_small(lpp_handle* h, double* vals, int n, cudaStream_t s);

static int ensure_vectors(lpp_handle* h, int kernel)
{
	if (!h->vx) CKR(dev_alloc(h, &h->vx, h->nloc));
	if (!h->vy) CKR(dev_alloc(h, &h->vy, h->nloc));
	CKR(ensure_two_layout(h, kernel));
	if (h->desc.nranks > 1 && h->two_layout != 1 && !h->yfull) CKR(dev_alloc(h, &h->yfull, h->rows));
	CKR(ensure_partials(h, lpp_vec_blocks(h->nloc)));
	return 0;
}

// sum of n block partials (+ allreduce over ranks) -> host double
static int reduce_scalar(lpp_handle* h, int npartials, double* out)
{
	lpp_launch_finalize_sum(h->partials, npartials, h->scal_dev, h->stream);
	h->launches += 1;
	cudaStream_t cs = h->stream;
	if (h->desc.nranks > 1) {
		if (!h->comm) return fail(LPP_ERR_STATE, "nranks>1 but lpp_comm_init has not been called");
		if (h->comm_stream && !h->p2p) {   // every NCCL call of this handle goes through one stream
			cs = h->comm_stream;
			CK(cudaEventRecord(h->ev_scal, h->stream));
			CK(cudaStreamWaitEvent(cs, h->ev_scal, 0));
		}
		CKN(g_nccl.AllReduce(h->scal_dev, h->scal_dev, 1, kNcclFloat64, kNcclSum, h->comm, cs));
	}
	CK(cudaMemcpyAsync(h->scal_host, h->scal_dev, sizeof(double), cudaMemcpyDeviceToHost, cs));
	CK(cudaStreamSynchronize(cs));
	*out = h->scal_host[0];
	return 0;
}

// nv sums at once: partials laid out [nv][npartials]
static int reduce_scalars(lpp_handle* h, int npartials, int nv, double* out)
{
	lpp_launch_finalize_sums(h->partials, npartials, nv, h->scal_dev, h->stream);
	h->launches += 1;
	cudaStream_t cs = h->stream;
	if (h->desc.nranks > 1) {
		if (!h->comm) return fail(LPP_ERR_STATE, "nranks>1 but lpp_comm_init has not been called");
		if (h->comm_stream && !h->p2p) {
			cs = h->comm_stream;
			CK(cudaEventRecord(h->ev_scal, h->stream));
			CK(cudaStreamWaitEvent(cs, h->ev_scal, 0));
		}
		CKN(g_nccl.AllReduce(h->scal_dev, h->scal_dev, nv, kNcclFloat64, kNcclSum, h->comm, cs));
	}
	CK(cudaMemcpyAsync(h->scal_host, h->scal_dev, sizeof(double) * nv, cudaMemcpyDeviceToHost, cs));
	CK(cudaStreamSynchronize(cs));
	for (int k = 0; k < nv; k++) out[k] = h->scal_host[k];
	return 0;
}

// ------------------------------------------------------------------ two-layout sharded SpMV
static int ensure_two_layout(lpp_handle* h, int kernel)
{
	if (h->two_layout >= 0) return 0;
	h->two_layout = 0;
	const char* env = getenv("LPP_TWO_LAYOUT");
	if (h->desc.nranks == 1 || h->desc.nranks > LPP_MAX_RANKS || (env && env[0] == '0')) return 0;
	if (h->md.model != LPP_MODEL_HUBBARD) return 0;          // FeAs two-spin terms need arbitrary remote elements
	if (resolve_kernel(h, kernel) != LPP_KERNEL_TILED) return 0;
	CKR(ensure_tiled(h));
	const int tl = lpp_tiled_two_layout_ok(h->tiled);
	if (tl == 0) return 0;
	// blocked up sweep (tl == 2): two-layout pays off with the peer-memory exchange only; handles that borrowed a communicator
	// (new sectors of the continued-fraction path) and were not given peer mappings exchange over NCCL send/recv and keep the
	// gather scheme
	if (tl == 2 && h->comm_borrowed && !h->p2p_requested) return 0;
	const int G = h->desc.nranks, me = h->desc.rank;
	const uint64_t n1 = h->md.n1, n2 = h->md.n2;
	// every rank needs a non-empty column shard (pairs of columns) and row shard: otherwise the gather scheme (all ranks
	// compute the same split, so the decision is consistent)
	if (n1 / 2 < (uint64_t)G || n2 < (uint64_t)G) return 0;
	h->cols.nranks = G;
	h->cols.me = me;
	h->dstart.assign(G + 1, 0);
	for (int r = 0; r < G; r++) {
		uint64_t f, c;
		shard_range(n1 / 2, r, G, &f, &c);                 // split pairs of columns: even shard widths allow 16-byte accesses
		h->cols.cs[r] = 2 * f;
		h->cols.cs[r + 1] = (r == G - 1) ? n1 : 2 * (f + c);
		shard_range(n2, r, G, &f, &c);
		h->dstart[r] = f;
		h->dstart[r + 1] = f + c;
	}
	h->ucol0 = h->cols.cs[me];
	h->ncols = h->cols.cs[me + 1] - h->cols.cs[me];
	{
		// [ycol | xcol | vx | vy] in one allocation (each part 256-byte aligned).  vx / vy join the slab when they do not exist
		// yet, so that peers can read this rank's Lanczos vector and PULL the columns they own (no "pack has landed" barrier).
		auto up32 = [](uint64_t v) { return (v + 31) & ~(uint64_t)31; };
		const uint64_t ncol = up32(n2 * h->ncols);
		const bool own_vecs = !h->vx && !h->vy;
		const uint64_t nv = own_vecs ? up32(h->nloc) : 0;
		CKR(dev_alloc(h, &h->slab, 2 * ncol + 2 * nv + LPP_PSX_DOUBLES));
		h->slab_off_psx = 2 * ncol + 2 * nv;
		CK(cudaMemset(h->slab + h->slab_off_psx, 0, sizeof(double) * LPP_PSX_DOUBLES));
		h->ycol = h->slab;
		h->xcol = h->slab + ncol;
		h->slab_off_xcol = ncol;
		if (own_vecs) {
			h->slab_off_vx = 2 * ncol;
			h->slab_off_vy = 2 * ncol + nv;
			h->vx = h->slab + h->slab_off_vx;
			h->vy = h->slab + h->slab_off_vy;
		}
	}
	h->partials2_cap = lpp_tiled_down_cols_blocks(h->tiled, h->md, h->ncols);
	CKR(dev_alloc(h, &h->partials2, (size_t)h->partials2_cap));
	CK(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
	CK(cudaEventCreateWithFlags(&h->ev_pack, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&h->ev_ycol, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&h->ev_xcol, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&h->ev_recv, cudaEventDisableTiming));
	CK(cudaEventCreateWithFlags(&h->ev_scal, cudaEventDisableTiming));
	h->two_layout = 1;
	return 0;
}

static void phase_mark(lpp_handle* h, int k, cudaStream_t s)
{
	if (h->phases < 0) {
		const char* e = getenv("LPP_PHASES");
		h->phases = (e && e[0] == '1') ? 1 : 0;
		if (h->phases)
			for (int i = 0; i < 8; i++) cudaEventCreate(&h->pev[i]);
		if (h->phases) { cudaEventCreate(&h->cev[0]); cudaEventCreate(&h->cev[1]); }
	}
	if (h->phases) cudaEventRecord(h->pev[k], s);
}
static void phase_collect(lpp_handle* h, int last)
{
	if (h->phases != 1) return;
	cudaEventSynchronize(h->pev[last]);
	for (int k = 0; k < last; k++) {
		float ms = 0;
		if (cudaEventElapsedTime(&ms, h->pev[k], h->pev[k + 1]) == cudaSuccess) h->phase_ms[k] += ms;
	}
	float pm = 0;
	if (cudaEventElapsedTime(&pm, h->cev[0], h->cev[1]) == cudaSuccess) h->pack_ms += pm;
	h->phase_n++;
}

// x = beta x + alpha H y on the ROW shard (x, y: local rows).  Flow (S = compute stream, C = comm stream):
//   S pack y -> per-peer column blocks | C all-to-all (y: ROW -> COLUMN layout)  ||  S up sweep on the ROW shard
//   S down sweep + diagonal on the COLUMN shard | C all-to-all (x: COLUMN -> ROW) | S x_row += received blocks
// The Lanczos dot <y, x> is the sum of the two sweeps' partial sums (it is layout independent).
static int spmv_two_layout(lpp_handle* h, double alpha, double beta, double* x, const double* y, bool want_dot, double* dot_out,
                           bool defer_unpack = false, const double* coefs_dev = nullptr)
{
	NvtxRange nvtx("lpp:spmv two-layout");
	if (!h->comm) return fail(LPP_ERR_STATE, "nranks>1 but lpp_comm_init has not been called");
	cudaStream_t S = h->stream, C = h->comm_stream;
	const int G = h->desc.nranks, me = h->desc.rank;
	const uint64_t n1 = h->md.n1, nrows = h->nloc / n1, d0loc = h->row0 / n1, ncme = h->ncols;
	if (h->p2p) {
		// Peer-memory path, one stream.  The two tiny all-reduces are also the cross-GPU barriers:
		//   pack (remote stores into every rank's ycol) | up sweep (local) | all-reduce #1: everybody's pack has landed
		//   down sweep on ycol -> xcol | all-reduce #2 (the Lanczos dot): everybody's xcol is final | unpack (remote loads)
		// The next pack / down sweep cannot overwrite a buffer a peer still reads: each is issued after an all-reduce
		// that the peer only enters once it has finished reading (stream order).
		// The pack (NVLink stores, no shared memory) runs on the second stream next to the up sweep (L1/shared bound).
		C = S;
		phase_mark(h, 0, S);
		const bool prepacked = h->packed_vec == y;              // the previous fused sweep already stored y into every ycol
		h->packed_vec = nullptr;
		CK(cudaEventRecord(h->ev_pack, S));
		CK(cudaStreamWaitEvent(h->comm_stream, h->ev_pack, 0));
		// The re-layout is G strided 2-D copies on the copy engines (rows x my-columns-for-peer-q blocks straight into the
		// peers' column shards), so it costs the up sweep no SM time; LPP_PACK_DMA=0 selects the store kernel instead.
		static const bool pack_dma = !(getenv("LPP_PACK_DMA") && getenv("LPP_PACK_DMA")[0] == '0');
		if (h->phases == 1) cudaEventRecord(h->cev[0], h->comm_stream);
		// pull form: y is this handle's vx or vy on every rank (all ranks swap them in step), and every rank's y is final once the
		// previous all-reduce has completed, so a rank can copy the columns it owns out of its peers' rows without any further
		// handshake; the copy is complete when the rank's own copy streams say so -- all-reduce #1 disappears.
		const PeerPtrs* ysrc = (y == h->vx) ? &h->peer_vx : (y == h->vy) ? &h->peer_vy : nullptr;
		const bool pulled = h->pull_pack && ysrc && !prepacked;
		if (prepacked) {
			// nothing to move
		} else if (pulled || pack_dma) {
			if (!h->copy_stream2) {
				CK(cudaStreamCreateWithFlags(&h->copy_stream2, cudaStreamNonBlocking));
				CK(cudaEventCreateWithFlags(&h->ev_copy2, cudaEventDisableTiming));
			}
			CK(cudaStreamWaitEvent(h->copy_stream2, h->ev_pack, 0));
			static const int nextra = []() { const char* e = getenv("LPP_PACK_STREAMS"); int v = e ? atoi(e) : 3; return v < 0 ? 0 : v > 4 ? 4 : v; }();
			for (int k = 0; k < nextra; k++) {
				if (!h->copy_streams[k]) {
					CK(cudaStreamCreateWithFlags(&h->copy_streams[k], cudaStreamNonBlocking));
					CK(cudaEventCreateWithFlags(&h->ev_copies[k], cudaEventDisableTiming));
				}
				CK(cudaStreamWaitEvent(h->copy_streams[k], h->ev_pack, 0));
			}
			for (int q = 0; q < G; q++) {
				const int qq = (me + q) % G;                     // the local block goes to its own stream (another copy engine)
				// remote blocks round-robin over comm_stream and the extra copy streams
				cudaStream_t cs = (q == 0) ? h->copy_stream2 : ((q - 1) % (nextra + 1) == 0) ? h->comm_stream : h->copy_streams[(q - 1) % (nextra + 1) - 1];
				if (pulled) {
					// rows of rank qq (its whole row shard), my columns  ->  rows dstart[qq].. of my column shard
					const uint64_t nrq = h->dstart[qq + 1] - h->dstart[qq];
					CK(cudaMemcpy2DAsync(h->ycol + h->dstart[qq] * ncme, ncme * sizeof(double), ysrc->p[qq] + h->cols.cs[me], n1 * sizeof(double),
					                     ncme * sizeof(double), nrq, cudaMemcpyDefault, cs));
				} else {
					const uint64_t ncq = h->cols.cs[qq + 1] - h->cols.cs[qq];
					CK(cudaMemcpy2DAsync(h->peer_ycol.p[qq] + d0loc * ncq, ncq * sizeof(double), y + h->cols.cs[qq], n1 * sizeof(double),
					                     ncq * sizeof(double), nrows, cudaMemcpyDefault, cs));
				}
			}
			CK(cudaEventRecord(h->ev_copy2, h->copy_stream2));
			CK(cudaStreamWaitEvent(h->comm_stream, h->ev_copy2, 0));
			for (int k = 0; k < nextra; k++) {
				CK(cudaEventRecord(h->ev_copies[k], h->copy_streams[k]));
				CK(cudaStreamWaitEvent(h->comm_stream, h->ev_copies[k], 0));
			}
		} else {
			lpp_launch_pack_cols_p2p(y, h->peer_ycol, nrows, n1, h->cols, d0loc, h->comm_stream);
		}
		if (h->phases == 1) cudaEventRecord(h->cev[1], h->comm_stream);
		CK(cudaEventRecord(h->ev_ycol, h->comm_stream));
		const int nbB = lpp_tiled_up_rows_blocks(h->tiled, nrows);
		CKR(ensure_partials(h, std::max(nbB, lpp_vec_blocks(h->nloc))));
		SpmvArgs ab;
		ab.alpha = alpha; ab.beta = beta; ab.x = x; ab.y = y; ab.row0 = 0; ab.nloc = h->nloc;
		if (coefs_dev) { ab.alpha.p = coefs_dev + LPP_LZ_ALPHA; ab.beta.p = coefs_dev + LPP_LZ_BETA; }
		ab.dot_partials = want_dot ? h->partials : nullptr;
		if (lpp_tiled_sweep_up_rows(h->tiled, h->md, ab, nrows, S) < 0) return fail(LPP_ERR_CUDA, lpp_tiled_error());
		phase_mark(h, 1, S);                                    // 0->1 up sweep
		CK(cudaStreamWaitEvent(S, h->ev_ycol, 0));
		// prepacked: the stores were issued before the previous all-reduce, which every rank enters after its fused sweep
		if (!pulled && !prepacked) CKR(allreduce_small(h, h->scal_dev + 4, 0, S));
		phase_mark(h, 2, S);                                    // 1->2 wait for the pack (+ all-reduce #1 when it was pushed)
		SpmvArgs aa;
		aa.alpha = alpha; aa.beta = 0.0; aa.x = h->xcol; aa.y = h->ycol; aa.row0 = 0; aa.nloc = h->md.n2 * ncme;
		if (coefs_dev) aa.alpha.p = coefs_dev + LPP_LZ_ALPHA;
		aa.dot_partials = want_dot ? h->partials2 : nullptr;
		if (lpp_tiled_sweep_down_cols(h->tiled, h->md, h->dn, h->dt, aa, h->ucol0, ncme, S) < 0)
			return fail(LPP_ERR_CUDA, lpp_tiled_error());
		phase_mark(h, 3, S);                                    // 2->3 down sweep
		if (want_dot) {
			lpp_launch_finalize_sum(h->partials, nbB, h->scal_dev, S);
			lpp_launch_finalize_sum(h->partials2, h->partials2_cap, h->scal_dev + 1, S);
		}
		CKR(allreduce_small(h, h->scal_dev, 2, S));
		phase_mark(h, 4, S);                                    // 3->4 finalize + all-reduce #2
		if (want_dot && !coefs_dev) CK(cudaMemcpyAsync(h->scal_host, h->scal_dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, S));
		// defer_unpack: the caller folds the re-layout of the column-shard result into its next pass over x
		if (!defer_unpack) lpp_launch_unpack_add_p2p(x, h->peer_xcol, nrows, n1, h->cols, d0loc, S);
		h->launches += (want_dot ? 6 : 4) - (defer_unpack ? 1 : 0);
		if (want_dot && !coefs_dev) {
			CK(cudaStreamSynchronize(S));
			*dot_out = h->scal_host[0] + h->scal_host[1];
		}
		CK(cudaGetLastError());
		return 0;
	}
	if (!h->sendbuf) CKR(dev_alloc(h, &h->sendbuf, h->nloc));
	if (!h->recvbuf) CKR(dev_alloc(h, &h->recvbuf, h->nloc));
	lpp_launch_pack_cols(y, h->sendbuf, h->ycol, nrows, n1, h->cols, d0loc, S);
	CK(cudaEventRecord(h->ev_pack, S));
	CK(cudaStreamWaitEvent(C, h->ev_pack, 0));
	CKN(g_nccl.GroupStart());
	for (int q = 0; q < G; q++) {
		if (q == me) continue;
		const uint64_t ncq = h->cols.cs[q + 1] - h->cols.cs[q], ndq = h->dstart[q + 1] - h->dstart[q];
		CKN(g_nccl.Send(h->sendbuf + nrows * h->cols.cs[q], nrows * ncq, kNcclFloat64, q, h->comm, C));
		CKN(g_nccl.Recv(h->ycol + h->dstart[q] * ncme, ndq * ncme, kNcclFloat64, q, h->comm, C));
	}
	CKN(g_nccl.GroupEnd());
	CK(cudaEventRecord(h->ev_ycol, C));
	const int nbB = lpp_tiled_up_rows_blocks(h->tiled, nrows);
	CKR(ensure_partials(h, std::max(nbB, lpp_vec_blocks(h->nloc))));
	SpmvArgs ab;
	ab.alpha = alpha; ab.beta = beta; ab.x = x; ab.y = y; ab.row0 = 0; ab.nloc = h->nloc;
	ab.dot_partials = want_dot ? h->partials : nullptr;
	if (lpp_tiled_sweep_up_rows(h->tiled, h->md, ab, nrows, S) < 0) return fail(LPP_ERR_CUDA, lpp_tiled_error());
	CK(cudaStreamWaitEvent(S, h->ev_ycol, 0));
	SpmvArgs aa;
	aa.alpha = alpha; aa.beta = 0.0; aa.x = h->xcol; aa.y = h->ycol; aa.row0 = 0; aa.nloc = h->md.n2 * ncme;
	aa.dot_partials = want_dot ? h->partials2 : nullptr;
	if (lpp_tiled_sweep_down_cols(h->tiled, h->md, h->dn, h->dt, aa, h->ucol0, ncme, S) < 0)
		return fail(LPP_ERR_CUDA, lpp_tiled_error());
	CK(cudaEventRecord(h->ev_xcol, S));
	h->launches += 3;
	if (want_dot) {
		lpp_launch_finalize_sum(h->partials, nbB, h->scal_dev, S);
		lpp_launch_finalize_sum(h->partials2, h->partials2_cap, h->scal_dev + 1, S);
		CK(cudaEventRecord(h->ev_scal, S));
		h->launches += 2;
	}
	CK(cudaStreamWaitEvent(C, h->ev_xcol, 0));
	CKN(g_nccl.GroupStart());
	for (int q = 0; q < G; q++) {
		if (q == me) continue;
		const uint64_t ncq = h->cols.cs[q + 1] - h->cols.cs[q], ndq = h->dstart[q + 1] - h->dstart[q];
		CKN(g_nccl.Send(h->xcol + h->dstart[q] * ncme, ndq * ncme, kNcclFloat64, q, h->comm, C));
		CKN(g_nccl.Recv(h->recvbuf + nrows * h->cols.cs[q], nrows * ncq, kNcclFloat64, q, h->comm, C));
	}
	CKN(g_nccl.GroupEnd());
	CK(cudaEventRecord(h->ev_recv, C));
	if (want_dot) {
		CK(cudaStreamWaitEvent(C, h->ev_scal, 0));
		CKN(g_nccl.AllReduce(h->scal_dev, h->scal_dev, 2, kNcclFloat64, kNcclSum, h->comm, C));
		CK(cudaMemcpyAsync(h->scal_host, h->scal_dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, C));
	}
	CK(cudaStreamWaitEvent(S, h->ev_recv, 0));
	lpp_launch_unpack_add(x, h->recvbuf, h->xcol, nrows, n1, h->cols, d0loc, S);
	h->launches += 2;
	if (want_dot) {
		CK(cudaStreamSynchronize(C));
		*dot_out = h->scal_host[0] + h->scal_host[1];
	}
	CK(cudaGetLastError());
	return 0;
}

static int allreduce_small(lpp_handle* h, double* vals, int n, cudaStream_t s)
{
	if (h->psx && h->p2p && n <= 4) {
		lpp_launch_psx_allreduce(vals, n, h->peer_psx, h->desc.rank, h->desc.nranks, ++h->psx_seq, h->psx_err, s);
		h->launches += 1;
		return 0;
	}
	CKN(g_nccl.AllReduce(vals, vals, std::max(n, 1), kNcclFloat64, kNcclSum, h->comm, s));
	return 0;
}

// gather the row-sharded vector into the full-length buffer every rank needs as gather source (halo = whole vector)
static int gather_full(lpp_handle* h, const double* local, const double** src)
{
	if (h->desc.nranks == 1) { *src = local; return 0; }
	if (!h->comm) return fail(LPP_ERR_STATE, "nranks>1 but lpp_comm_init has not been called");
	CKN(g_nccl.GroupStart());
	for (int r = 0; r < h->desc.nranks; r++)
		CKN(g_nccl.Broadcast(local, h->yfull + h->shard_row0[r], h->shard_nloc[r], kNcclFloat64, r, h->comm, h->stream));
	CKN(g_nccl.GroupEnd());
	*src = h->yfull;
	return 0;
}

struct LoopTiming {
	int from = -1, to = -1;  // record ev0 before iteration `from`, ev1 after iteration `to-1`
	int64_t launches_at_from = 0, launches_at_to = 0;
};

// The sharded recurrence with the exchange pipelined behind the sweeps (peer-memory two-layout handles, device-resident scalars).
// In lanczos_loop's plain form an iteration is a chain  [pack || up sweep] -> down sweep -> unpack -> norm all-reduce -> next,
// and the unpack (remote loads over NVLink) overlaps with nothing.  Here the sweeps run with alpha = 1, beta = 0 (the up sweep into
// a third vector z), and every scalar enters in the unpack:  U_{j+1} = (z + pieces)/n_j - (a_j/n_j) U_j - (b_j/n_{j-1}) U_{j-1}.
// |U_{j+1}| is then needed by the NEXT unpack only, so the next up sweep and the next pack start chunk by chunk of rows as soon
// as the unpack has written them, and the norm rides on the all-reduce that doubles as the "packs have landed" barrier:
//   S   : up(c0) up(c2) ..          | all-reduce A {norm of the previous unpack} | down sweep | all-reduce B {dot}
//   S2  :    up(c1) up(c3) ..       |
//   DMA :  pack(c0) pack(c1) ..     |                                             (copy engines, into the peers' column shards)
//   U   :                                                                          unpack(c0) unpack(c1) ..   -> events per chunk
// Buffer safety: a peer overwrites my ycol (its next pack) only after all-reduce B of this iteration, which I enter after my
// down sweep; I overwrite xcol (my next down sweep) only after all-reduce A of the next iteration, which every peer enters
// after its up sweeps, i.e. after all of its unpack chunks.  All all-reduces are issued on S, in the same order on every rank.
static int lanczos_pipelined(lpp_handle* h, const lpp_solver_params* p, int steps, bool check_convergence, double nj0, double* x, double* y,
                             double* a, double* b, int* nsteps, LoopTiming* tm, int nchunks)
{
	NvtxRange nvtx("lpp:lanczos pipelined");
	cudaStream_t S = h->stream;
	const int G = h->desc.nranks, me = h->desc.rank;
	const uint64_t n1 = h->md.n1, nrows = h->nloc / n1, d0loc = h->row0 / n1, ncme = h->ncols;
	if (!h->vz) CKR(dev_alloc(h, &h->vz, h->nloc));
	const int gy = lpp_unpack3_partials_per_row(n1);
	const int npb = (int)nrows * gy;
	if (h->partials3_cap < npb) {
		if (h->partials3) dev_free(h, h->partials3);
		h->partials3 = nullptr;
		CKR(dev_alloc(h, &h->partials3, (size_t)npb));
		h->partials3_cap = npb;
	}
	if (!h->pipe_unpack) {
		int least = 0, greatest = 0;
		CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
		for (int k = 0; k < 2; k++) {
			CK(cudaStreamCreateWithPriority(&h->pipe_up[k], cudaStreamNonBlocking, greatest));
			CK(cudaEventCreateWithFlags(&h->ev_pipe_up[k], cudaEventDisableTiming));
		}
		CK(cudaStreamCreateWithPriority(&h->pipe_unpack, cudaStreamNonBlocking, least));
		for (cudaEvent_t& e : h->ev_pipe_chunk) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&h->ev_pipe_dot, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&h->ev_pipe_y, cudaEventDisableTiming));
	}
	// the unpack runs beside the next up sweep; LPP_PIPE_UNPACK_CPS = n limits its grid to n CTAs per SM striding over the rows.
	// Measured on 2 x B200 (config 3, 4 chunks): 1 / 2 / 4 CTAs per SM 5.12 / 3.75 / 3.08 ms per iteration, no limit 2.88 ms
	static const int unpack_cps = getenv("LPP_PIPE_UNPACK_CPS") ? atoi(getenv("LPP_PIPE_UNPACK_CPS")) : 0;
	int nsm = 148;
	CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device));
	const int unpack_rows = unpack_cps > 0 ? std::max(1, unpack_cps * nsm / std::max(1, lpp_unpack3_partials_per_row(n1))) : 0;
	const int ncopy = 4;
	for (int k = 0; k < ncopy; k++)
		if (!h->copy_streams[k]) {
			CK(cudaStreamCreateWithFlags(&h->copy_streams[k], cudaStreamNonBlocking));
			CK(cudaEventCreateWithFlags(&h->ev_copies[k], cudaEventDisableTiming));
		}
	// chunks of rows (even sizes: the up sweep takes two rows per CTA)
	nchunks = std::max(1, std::min(nchunks, 8));
	uint64_t cb[9];
	for (int k = 0; k <= nchunks; k++) cb[k] = std::min(nrows, ((nrows * (uint64_t)k / (uint64_t)nchunks) + 1) & ~(uint64_t)1);
	cb[nchunks] = nrows;
	const int nbB = lpp_tiled_up_rows_blocks(h->tiled, nrows);
	int up_blocks_before[9];
	up_blocks_before[0] = 0;
	for (int k = 0; k < nchunks; k++) up_blocks_before[k + 1] = up_blocks_before[k] + lpp_tiled_up_rows_blocks(h->tiled, cb[k + 1] - cb[k]);
	CKR(ensure_partials(h, std::max(std::max(nbB, up_blocks_before[nchunks]), lpp_vec_blocks(h->nloc))));
	double* coefs = h->lz_coefs;
	double* a_dev = h->lz_ab;
	double* b_dev = h->lz_ab + h->lz_ab_cap;
	double* z = h->vz;
	lpp_launch_lzp_init(nj0, coefs, S);
	h->launches += 1;
	const bool watch = check_convergence && p->eps > 0;
	static const int sync_env = getenv("LPP_SYNC_EVERY") ? atoi(getenv("LPP_SYNC_EVERY")) : 0;
	const int sync_every = std::max(1, sync_env > 0 ? sync_env : (watch ? 8 : 32));
	bool in_flight = false;          // unpack chunks of the previous iteration are (possibly) still running on pipe_unpack
	bool norm_pending = false;       // their norm partials have not been reduced yet
	double eold = 100.0;
	int j = 0;
	bool stop = false;
	// S waits for every unpack chunk, reduces the pending norm over the ranks (n = 1) or just meets them (n = 0)
	auto close_norm = [&](bool barrier_anyway) -> int {
		if (in_flight)
			for (int k = 0; k < nchunks; k++) CK(cudaStreamWaitEvent(S, h->ev_pipe_chunk[k], 0));
		in_flight = false;
		if (norm_pending) {
			lpp_launch_finalize_sum(h->partials3, npb, h->scal_dev + 2, S);
			CKR(allreduce_small(h, h->scal_dev + 2, 1, S));
			lpp_launch_lzp_after_norm(h->scal_dev + 2, coefs, b_dev, S);
			h->launches += 2;
			norm_pending = false;
		} else if (barrier_anyway) {
			CKR(allreduce_small(h, h->scal_dev + 4, 0, S));
		}
		return 0;
	};
	while (j < steps && !stop) {
		const int j0 = j, j1 = std::min(steps, j + sync_every);
		for (int jj = j0; jj < j1; jj++) {
			if (tm && jj == tm->from) {
				CKR(close_norm(false));                           // the timed region starts with nothing in flight
				CK(cudaEventRecord(h->ev0, S));
				tm->launches_at_from = h->launches;
			}
			// ---- up sweep (z = T_up y) and pack (y -> the peers' column shards), chunk by chunk behind the previous unpack
			CK(cudaEventRecord(h->ev_pipe_y, S));                  // everything S has done so far (y final when nothing is in flight)
			for (int k = 0; k < 2; k++) CK(cudaStreamWaitEvent(h->pipe_up[k], h->ev_pipe_y, 0));
			if (!in_flight)
				for (int k = 0; k < ncopy; k++) CK(cudaStreamWaitEvent(h->copy_streams[k], h->ev_pipe_y, 0));
			int ncp = 0;
			for (int k = 0; k < nchunks; k++) {
				const uint64_t r0 = cb[k], nr = cb[k + 1] - cb[k];
				if (nr == 0) continue;
				cudaStream_t us = h->pipe_up[k & 1];
				if (in_flight) {
					CK(cudaStreamWaitEvent(us, h->ev_pipe_chunk[k], 0));
					for (int c = 0; c < ncopy; c++) CK(cudaStreamWaitEvent(h->copy_streams[c], h->ev_pipe_chunk[k], 0));
				}
				SpmvArgs ab;
				ab.alpha = 1.0; ab.beta = 0.0; ab.x = z + r0 * n1; ab.y = y + r0 * n1; ab.row0 = 0; ab.nloc = nr * n1;
				ab.dot_partials = h->partials + up_blocks_before[k];
				if (lpp_tiled_sweep_up_rows(h->tiled, h->md, ab, nr, us) < 0) return fail(LPP_ERR_CUDA, lpp_tiled_error());
				for (int q = 0; q < G; q++) {
					const int qq = (me + q) % G;
					const uint64_t ncq = h->cols.cs[qq + 1] - h->cols.cs[qq];
					if (ncq == 0) continue;
					CK(cudaMemcpy2DAsync(h->peer_ycol.p[qq] + (d0loc + r0) * ncq, ncq * sizeof(double), y + r0 * n1 + h->cols.cs[qq], n1 * sizeof(double),
					                     ncq * sizeof(double), nr, cudaMemcpyDefault, h->copy_streams[ncp % ncopy]));
					ncp++;
				}
			}
			for (int k = 0; k < 2; k++) {
				CK(cudaEventRecord(h->ev_pipe_up[k], h->pipe_up[k]));
				CK(cudaStreamWaitEvent(S, h->ev_pipe_up[k], 0));
			}
			for (int k = 0; k < ncopy; k++) {
				CK(cudaEventRecord(h->ev_copies[k], h->copy_streams[k]));
				CK(cudaStreamWaitEvent(S, h->ev_copies[k], 0));
			}
			h->launches += nchunks;
			// ---- all-reduce A: the norm of the vector the previous unpack built; every rank's pack has landed behind it
			in_flight = false;                                   // S has waited for every chunk through its up sweeps
			CKR(close_norm(true));
			// ---- down sweep on the column shard, dot = <y, z> + <ycol, xcol>
			SpmvArgs aa;
			aa.alpha = 1.0; aa.beta = 0.0; aa.x = h->xcol; aa.y = h->ycol; aa.row0 = 0; aa.nloc = h->md.n2 * ncme;
			aa.dot_partials = h->partials2;
			if (lpp_tiled_sweep_down_cols(h->tiled, h->md, h->dn, h->dt, aa, h->ucol0, ncme, S) < 0) return fail(LPP_ERR_CUDA, lpp_tiled_error());
			lpp_launch_finalize_sum(h->partials, up_blocks_before[nchunks], h->scal_dev, S);
			lpp_launch_finalize_sum(h->partials2, h->partials2_cap, h->scal_dev + 1, S);
			CKR(allreduce_small(h, h->scal_dev, 2, S));            // also: every rank's xcol is final
			lpp_launch_lzp_after_dot(h->scal_dev, coefs, a_dev, S);
			h->launches += 4;
			// ---- unpack: x = C1 (z + pieces) - C2 y - C3 x, chunk by chunk on its own stream
			CK(cudaEventRecord(h->ev_pipe_dot, S));
			CK(cudaStreamWaitEvent(h->pipe_unpack, h->ev_pipe_dot, 0));
			for (int k = 0; k < nchunks; k++) {
				const uint64_t r0 = cb[k], nr = cb[k + 1] - cb[k];
				lpp_launch_unpack3_norm_p2p(x + r0 * n1, y + r0 * n1, z + r0 * n1, coefs, h->peer_xcol, nr, n1, h->cols, d0loc + r0,
				                            h->partials3 + r0 * (uint64_t)gy, unpack_rows, h->pipe_unpack);
				CK(cudaEventRecord(h->ev_pipe_chunk[k], h->pipe_unpack));
			}
			h->launches += nchunks;
			in_flight = true;
			norm_pending = true;
			if (tm && jj + 1 == tm->to) {
				CKR(close_norm(false));                           // the timed region ends with the iteration complete, b_j included
				CK(cudaEventRecord(h->ev1, S));
				tm->launches_at_to = h->launches;
			}
			std::swap(x, y);
		}
		CKR(close_norm(false));
		CK(cudaMemcpyAsync(h->lz_ab_host + j0, a_dev + j0, sizeof(double) * (j1 - j0), cudaMemcpyDeviceToHost, S));
		CK(cudaMemcpyAsync(h->lz_ab_host + h->lz_ab_cap + j0, b_dev + j0, sizeof(double) * (j1 - j0), cudaMemcpyDeviceToHost, S));
		CK(cudaStreamSynchronize(S));
		if (h->psx_err && *h->psx_err) return fail(LPP_ERR_STATE, "peer-memory scalar exchange timed out (a rank did not arrive)");
		for (j = j0; j < j1; j++) {
			a[j] = h->lz_ab_host[j];
			b[j] = h->lz_ab_host[h->lz_ab_cap + j];
			if (watch && j >= h->conv_index) {
				double enew = tridiag_kth(j + 1, a, b, h->conv_index);
				if (fabs(enew - eold) < p->eps && (j >= p->minsteps || h->rows <= 4)) { j++; stop = true; break; }
				eold = enew;
			}
		}
	}
	*nsteps = j;
	h->packed_vec = nullptr;
	CK(cudaGetLastError());
	return 0;
}

// PsimagLite::LanczosSolver::decomposition with the three vector sweeps fused into two:
// the dot <y,x> rides on the SpMV epilogue, the swap/scale sweep is folded into scalar coefficients.
// State: y = U_j (unnormalised Lanczos vector, v_j = U_j/n_j), x = U_{j-1}.
static int lanczos_loop(lpp_handle* h, const lpp_solver_params* p, int steps, bool check_convergence,
                        const double* zcoef, double* z, double* a, double* b, int* nsteps, double* init_norm2,
                        LoopTiming* tm)
{
	NvtxRange nvtx("lpp:lanczos");
	const uint64_t n = h->nloc;
	double* x = h->vx;
	double* y = h->vy;
	h->packed_vec = nullptr;                                 // y was just (re)loaded: no column-shard copy of it exists
	int np = lpp_vec_blocks(n);
	lpp_launch_dot(y, y, n, h->partials, h->stream);
	h->launches += 1;
	double nrm2 = 0;
	CKR(reduce_scalar(h, np, &nrm2));
	if (init_norm2) *init_norm2 = nrm2;
	if (!(nrm2 > 0)) return fail(LPP_ERR_ARG, "initial Lanczos vector has zero norm");
	if (steps < 1) return fail(LPP_ERR_ARG, "the number of Lanczos steps must be at least 1");
	// the sharding of a handle is decided once (by the kernel of the first call / of lpp_p2p_export); a two-layout handle only runs
	// the tiled sweeps, so asking it for another kernel later is an error rather than a silent switch
	if (h->two_layout == 1 && resolve_kernel(h, p->kernel) != LPP_KERNEL_TILED)
		return fail(LPP_ERR_ARG, "this sharded handle is laid out for the tiled kernel (two-layout exchange); create another handle for a different kernel");
	double nj = sqrt(nrm2), nprev = 1.0, bprev = 0.0, eold = 100.0;
	if ((uint64_t)steps > h->rows) steps = (int)h->rows;
	// <prefix>Options=reortho: every Lanczos vector is kept on the device (un-normalised U_k with its squared norm) and the new
	// vector is orthogonalised against all of them after x -= a y; freed when the loop ends
	struct SavedVectors {
		std::vector<double*> v;
		std::vector<double> n2;
		std::vector<double*> slabs;       // the vectors are carved out of slabs of 8: one cudaMalloc (an implicit device sync) per 8 steps
		~SavedVectors() { for (double* q : slabs) cudaFree(q); }
	} saved;
	const bool reortho = p->reortho != 0;
	const int npro = lpp_vec_blocks(n * 2);
	if (reortho) CKR(ensure_partials(h, std::max(np, npro * LPP_RO_NV)));
	int j = 0;
	// Device-resident recurrence: 1/n_j, -b_{j-1}/n_{j-1} and a_j/n_j live in device memory (LPP_LZ_*), updated by one-thread
	// kernels right after the reductions, so an iteration is enqueued without a host round trip; the host reads (a_j, b_j) back
	// every `sync_every` steps and applies LanczosSolver's convergence test to them in order.  When the test fires inside a
	// batch the device has run at most sync_every - 1 steps too many; (a, b) are truncated exactly where the reference stops.
	static const bool dev_scalars_on = !(getenv("LPP_DEV_SCALARS") && getenv("LPP_DEV_SCALARS")[0] == '0');
	const bool nccl_two_layout = h->two_layout == 1 && !h->p2p;      // send/recv exchange keeps its own host synchronisation
	if (dev_scalars_on && !reortho && !zcoef && !nccl_two_layout && (h->desc.nranks == 1 || h->comm)) {
		if (!h->lz_coefs) CKR(dev_alloc(h, &h->lz_coefs, 16));
		if (h->lz_ab_cap < steps) {
			if (h->lz_ab) dev_free(h, h->lz_ab);
			if (h->lz_ab_host) cudaFreeHost(h->lz_ab_host);
			h->lz_ab = nullptr;
			h->lz_ab_host = nullptr;
			CKR(dev_alloc(h, &h->lz_ab, 2 * (size_t)steps));
			CK(cudaMallocHost((void**)&h->lz_ab_host, 2 * (size_t)steps * sizeof(double)));
			h->lz_ab_cap = steps;
		}
		// LPP_PIPELINE=<chunks>: the exchange pipelined behind the sweeps (lanczos_pipelined); peer-memory two-layout handles whose
		// column split is even (16-byte accesses).  Opt-in: on 2 x B200 it is slower than the plain chain (2.85-2.94 ms against
		// 2.63 ms per iteration of config 3) -- the unpack needs the whole GPU to run at NVLink speed, beside the up sweep both slow down
		{
			static const int pipe_chunks = getenv("LPP_PIPELINE") ? atoi(getenv("LPP_PIPELINE")) : 0;
			bool even = (h->md.n1 % 2 == 0) && reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0;
			if (h->two_layout == 1) for (int q = 0; q <= h->desc.nranks; q++) even = even && (h->cols.cs[q] % 2 == 0);
			if (pipe_chunks > 0 && h->two_layout == 1 && h->p2p && even && h->nloc / h->md.n1 >= (uint64_t)(4 * pipe_chunks))
				return lanczos_pipelined(h, p, steps, check_convergence, nj, x, y, a, b, nsteps, tm, pipe_chunks);
		}
		double* coefs = h->lz_coefs;
		double* a_dev = h->lz_ab;
		double* b_dev = h->lz_ab + h->lz_ab_cap;
		lpp_launch_lz_init(nj, coefs, h->stream);
		h->launches += 1;
		const bool watch = check_convergence && p->eps > 0;
		static const int sync_env = getenv("LPP_SYNC_EVERY") ? atoi(getenv("LPP_SYNC_EVERY")) : 0;
		const int sync_every = std::max(1, sync_env > 0 ? sync_env : (watch ? 4 : 32));
		const bool fuse_unpack = h->two_layout == 1 && h->p2p;
		auto reduce_dev = [&](int npartials, double* out_dev) -> int {
			lpp_launch_finalize_sum(h->partials, npartials, out_dev, h->stream);
			h->launches += 1;
			if (h->desc.nranks > 1) CKR(allreduce_small(h, out_dev, 1, h->stream));
			return 0;
		};
		bool stop = false;
		while (j < steps && !stop) {
			const int j0 = j, j1 = std::min(steps, j + sync_every);
			for (int jj = j0; jj < j1; jj++) {
				if (tm && jj == tm->from) { CK(cudaEventRecord(h->ev0, h->stream)); tm->launches_at_from = h->launches; }
				if (h->two_layout == 1) {
					CKR(spmv_two_layout(h, 0.0, 0.0, x, y, true, nullptr, fuse_unpack, coefs));   // dot parts in scal_dev[0..1], reduced
					lpp_launch_lz_after_dot(h->scal_dev, 2, coefs, a_dev, h->stream);
				} else {
					const double* src = nullptr;
					CKR(gather_full(h, y, &src));
					int nparts = 0;
					CKR(do_spmv(h, p->kernel, 0.0, 0.0, x, src, true, &nparts, coefs));
					CKR(reduce_dev(nparts, h->scal_dev));
					lpp_launch_lz_after_dot(h->scal_dev, 1, coefs, a_dev, h->stream);
				}
				int npb = np;
				if (fuse_unpack) {
					const uint64_t n1 = h->md.n1, nrows = h->nloc / n1;
					npb = lpp_unpack_axpy_norm_blocks(nrows, n1, h->desc.nranks);
					CKR(ensure_partials(h, npb));
					// LPP_FUSE_PACK=1: the fused sweep also stores the new vector into the owners' column shards (the next pack), so the
					// next up sweep runs without the copy engines beside it and all-reduce #1 disappears.  Measured on 8 x B200: up sweep
					// 0.272 -> 0.247 ms, pack wait 0.07 -> 0.007 ms, but the sweep itself 0.255 -> 0.48 ms (1.03 ms per iteration against
					// 0.92 ms); 4 GPUs 1.84 against 1.49 ms; 2 GPUs 3.23 against 2.97 ms.  Off by default.
					static const bool fuse_pack = getenv("LPP_FUSE_PACK") && getenv("LPP_FUSE_PACK")[0] == '1';
					lpp_launch_unpack_axpy_norm_p2p(x, y, 0.0, h->peer_xcol, fuse_pack ? &h->peer_ycol : nullptr, nrows, n1, h->cols, h->row0 / n1,
					                                h->partials, h->stream, coefs + LPP_LZ_AXPY);
					if (fuse_pack) h->packed_vec = x;               // x is the next Lanczos vector (becomes y after the swap)
					phase_mark(h, 5, h->stream);
				} else {
					lpp_launch_axpy_norm(x, y, 0.0, n, h->partials, h->stream, coefs + LPP_LZ_AXPY);
				}
				CKR(reduce_dev(npb, h->scal_dev + 2));
				lpp_launch_lz_after_norm(h->scal_dev + 2, coefs, b_dev, h->stream);
				h->launches += 3;
				if (fuse_unpack) { phase_mark(h, 6, h->stream); phase_collect(h, 6); }
				if (tm && jj + 1 == tm->to) { CK(cudaEventRecord(h->ev1, h->stream)); tm->launches_at_to = h->launches; }
				std::swap(x, y);
			}
			CK(cudaMemcpyAsync(h->lz_ab_host + j0, a_dev + j0, sizeof(double) * (j1 - j0), cudaMemcpyDeviceToHost, h->stream));
			CK(cudaMemcpyAsync(h->lz_ab_host + h->lz_ab_cap + j0, b_dev + j0, sizeof(double) * (j1 - j0), cudaMemcpyDeviceToHost, h->stream));
			CK(cudaStreamSynchronize(h->stream));
			if (h->psx_err && *h->psx_err) return fail(LPP_ERR_STATE, "peer-memory scalar exchange timed out (a rank did not arrive)");
			for (j = j0; j < j1; j++) {
				a[j] = h->lz_ab_host[j];
				b[j] = h->lz_ab_host[h->lz_ab_cap + j];
				if (watch && j >= h->conv_index) {
					double enew = tridiag_kth(j + 1, a, b, h->conv_index);
					if (fabs(enew - eold) < p->eps && (j >= p->minsteps || h->rows <= 4)) { j++; stop = true; break; }
					eold = enew;
				}
			}
		}
		*nsteps = j;
		h->packed_vec = nullptr;
		return 0;
	}
	for (; j < steps; j++) {
		if (tm && j == tm->from) { CK(cudaEventRecord(h->ev0, h->stream)); tm->launches_at_from = h->launches; }
		if (reortho) {
			const uint64_t stride = (std::max<uint64_t>(n, 1) + 31) & ~(uint64_t)31;
			if (saved.v.size() % 8 == 0) {
				double* slab = nullptr;
				if (cudaMalloc((void**)&slab, sizeof(double) * stride * 8) != cudaSuccess) {
					cudaGetLastError();
					return fail(LPP_ERR_CUDA, "reortho: out of device memory for the saved Lanczos vectors (steps x rows x 8 bytes)");
				}
				saved.slabs.push_back(slab);
			}
			double* keep = saved.slabs.back() + (saved.v.size() % 8) * stride;
			saved.v.push_back(keep);
			saved.n2.push_back(nj * nj);
			CK(cudaMemcpyAsync(keep, y, sizeof(double) * n, cudaMemcpyDeviceToDevice, h->stream));
		}
		if (zcoef) { lpp_launch_axpy(z, y, zcoef[j] / nj, n, h->stream); h->launches += 1; }
		double dot = 0;
		const bool fuse_unpack = h->two_layout == 1 && h->p2p;
		if (h->two_layout == 1) {
			CKR(spmv_two_layout(h, 1.0 / nj, j == 0 ? 0.0 : -(bprev / nprev), x, y, true, &dot, fuse_unpack));
		} else {
			const double* src = nullptr;
			CKR(gather_full(h, y, &src));
			int nparts = 0;
			CKR(do_spmv(h, p->kernel, 1.0 / nj, j == 0 ? 0.0 : -(bprev / nprev), x, src, true, &nparts));
			CKR(reduce_scalar(h, nparts, &dot));
		}
		double aj = dot / nj;
		int npb = np;
		if (fuse_unpack) {
			const uint64_t n1 = h->md.n1, nrows = h->nloc / n1;
			npb = lpp_unpack_axpy_norm_blocks(nrows, n1, h->desc.nranks);
			CKR(ensure_partials(h, npb));
			// LPP_FUSE_PACK=1 also stores the new vector into the peers' column shards from this kernel (the next pack);
			// measured slower on 2 x B200 (1.71 ms against 0.70 + 0.59 ms for the separate kernels), so it is off by default.
			static const bool fuse_pack = getenv("LPP_FUSE_PACK") && getenv("LPP_FUSE_PACK")[0] == '1';
			lpp_launch_unpack_axpy_norm_p2p(x, y, aj / nj, h->peer_xcol, fuse_pack ? &h->peer_ycol : nullptr, nrows, n1, h->cols,
			                                h->row0 / n1, h->partials, h->stream);
			if (fuse_pack) h->packed_vec = x;                   // x is the next Lanczos vector (becomes y after the swap)
		} else {
			lpp_launch_axpy_norm(x, y, aj / nj, n, h->partials, h->stream);
		}
		h->launches += 1;
		if (fuse_unpack) phase_mark(h, 5, h->stream);            // 4->5 host round trip + fused unpack/axpy/norm
		double b2 = 0;
		CKR(reduce_scalar(h, npb, &b2));
		if (fuse_unpack) { phase_mark(h, 6, h->stream); phase_collect(h, 6); }   // 5->6 reduce + all-reduce #3
		if (reortho) {
			for (size_t k0 = 0; k0 < saved.v.size(); k0 += LPP_RO_NV) {
				RoVecs r;
				r.nv = (int)std::min<size_t>(LPP_RO_NV, saved.v.size() - k0);
				for (int k = 0; k < LPP_RO_NV; k++) { r.v[k] = saved.v[k0 + std::min(k, r.nv - 1)]; r.coef[k] = 0.0; }
				double dots[LPP_RO_NV];
				lpp_launch_reortho_dots(x, r, n, h->partials, h->stream);
				h->launches += 1;
				CKR(reduce_scalars(h, npro, r.nv, dots));
				for (int k = 0; k < r.nv; k++) r.coef[k] = dots[k] / saved.n2[k0 + k];
				lpp_launch_reortho_axpy_norm(x, r, n, h->partials, h->stream);
				h->launches += 1;
			}
			CKR(reduce_scalar(h, npro, &b2));
		}
		double bj = sqrt(b2);
		a[j] = aj;
		b[j] = bj;
		if (tm && j + 1 == tm->to) { CK(cudaEventRecord(h->ev1, h->stream)); tm->launches_at_to = h->launches; }
		nprev = nj;
		bprev = bj;
		nj = (bj < 1e-10) ? 1.0 : bj;
		std::swap(x, y);
		if (check_convergence && p->eps > 0) {
			// LanczosSolver::computeAllStatesBelow watches the highest requested state; until the tridiagonal has that many
			// rows there is nothing to compare
			if (j >= h->conv_index) {
				double enew = tridiag_kth(j + 1, a, b, h->conv_index);
				if (fabs(enew - eold) < p->eps && (j >= p->minsteps || h->rows <= 4)) { j++; break; }
				eold = enew;
			}
		}
	}
	*nsteps = j;
	return 0;
}

static int load_init(lpp_handle* h, const lpp_solver_params* p, const double* init_host, int use_modified)
{
	CKR(ensure_vectors(h, p->kernel));
	if (init_host) {
		CK(cudaMemcpyAsync(h->vy, init_host + h->row0, sizeof(double) * h->nloc, cudaMemcpyHostToDevice, h->stream));
	} else if (use_modified) {
		if (!h->modified) return fail(LPP_ERR_STATE, "no modified vector in this handle (call lpp_apply_op first)");
		CK(cudaMemcpyAsync(h->vy, h->modified, sizeof(double) * h->nloc, cudaMemcpyDeviceToDevice, h->stream));
	} else {
		lpp_launch_fill_random(h->vy, h->row0, h->nloc, p->seed, h->stream);
		h->launches += 1;
	}
	return 0;
}

extern "C" int lpp_lanczos_decomposition(lpp_handle* h, const lpp_solver_params* p, const double* init_host,
                                         int32_t use_modified, double* a, double* b, int32_t* nsteps, double* init_norm2)
{
	if (!h || !p || !a || !b || !nsteps) return fail(LPP_ERR_ARG, "null argument");
	CK(cudaSetDevice(h->device));
	CKR(load_init(h, p, init_host, use_modified));
	int ns = 0;
	CKR(lanczos_loop(h, p, p->steps, true, nullptr, nullptr, a, b, &ns, init_norm2, nullptr));
	*nsteps = ns;
	return 0;
}

extern "C" int lpp_ground_state(lpp_handle* h, const lpp_solver_params* p, const double* init_host, int32_t want_vector,
                                double* energy, double* z_host, double* a, double* b, int32_t* nsteps)
{
	if (!h || !p || !energy) return fail(LPP_ERR_ARG, "null argument");
	CK(cudaSetDevice(h->device));
	int cap = (int)std::min<uint64_t>((uint64_t)p->steps, h->rows);
	std::vector<double> aa(cap + 1), bb(cap + 1);
	CKR(load_init(h, p, init_host, 0));
	int ns = 0;
	CKR(lanczos_loop(h, p, p->steps, true, nullptr, nullptr, aa.data(), bb.data(), &ns, nullptr, nullptr));
	std::vector<double> d(aa.begin(), aa.begin() + ns), e(ns, 0.0), zz((size_t)ns * ns);
	for (int i = 0; i + 1 < ns; i++) e[i] = bb[i];
	if (tridiag_full(ns, d, e, zz.data()) != 0) return fail(LPP_ERR_STATE, "tridiagonal QL did not converge");
	*energy = d[0];
	if (want_vector || z_host) {
		// second pass: replay the identical recurrence, accumulating z = sum_j c_j v_j
		std::vector<double> coef(ns);
		for (int j = 0; j < ns; j++) coef[j] = zz[(size_t)j * ns + 0];
		if (!h->gs) CKR(dev_alloc(h, &h->gs, h->nloc));
		CK(cudaMemsetAsync(h->gs, 0, sizeof(double) * h->nloc, h->stream));
		CKR(load_init(h, p, init_host, 0));
		std::vector<double> a2(ns + 1), b2(ns + 1);
		int ns2 = 0;
		CKR(lanczos_loop(h, p, ns, false, coef.data(), h->gs, a2.data(), b2.data(), &ns2, nullptr, nullptr));
		CK(cudaStreamSynchronize(h->stream));
		if (z_host) CK(cudaMemcpy(z_host + h->row0, h->gs, sizeof(double) * h->nloc, cudaMemcpyDeviceToHost));
	}
	if (a) std::copy(aa.begin(), aa.begin() + ns, a);
	if (b) std::copy(bb.begin(), bb.begin() + ns, b);
	if (nsteps) *nsteps = ns;
	return 0;
}

// LanczosSolver::computeAllStatesBelow(eigs, zs, initial, excitedPlusOne) (Engine.h:626): the lowest `nstates` Ritz pairs of one
// decomposition whose convergence test watches state nstates-1.  Vectors are rebuilt by replaying the recurrence once per state
// (nothing is saved unless p->reortho); state 0 stays in the handle as the ground state.  Without reorthogonalisation the higher
// Ritz values of a long run can be ghost copies of converged ones (the same holds for the reference): use Options=reortho.
extern "C" int lpp_states_below(lpp_handle* h, const lpp_solver_params* p, const double* init_host, int32_t nstates, double* energies,
                                double* z_host, int32_t* nsteps)
{
	if (!h || !p || !energies || nstates < 1) return fail(LPP_ERR_ARG, "bad argument");
	if ((uint64_t)nstates > h->rows) return fail(LPP_ERR_ARG, "more states requested than the sector has rows");
	CK(cudaSetDevice(h->device));
	int cap = (int)std::min<uint64_t>((uint64_t)p->steps, h->rows);
	if (cap < nstates) return fail(LPP_ERR_ARG, "LanczosSteps smaller than the number of states requested");
	std::vector<double> aa(cap + 1), bb(cap + 1);
	CKR(load_init(h, p, init_host, 0));
	int ns = 0;
	h->conv_index = nstates - 1;
	int rc = lanczos_loop(h, p, p->steps, true, nullptr, nullptr, aa.data(), bb.data(), &ns, nullptr, nullptr);
	h->conv_index = 0;
	if (rc != 0) return rc;
	if (ns < nstates) return fail(LPP_ERR_STATE, "the Krylov space closed before the requested number of states");
	std::vector<double> d(aa.begin(), aa.begin() + ns), e(ns, 0.0), zz((size_t)ns * ns);
	for (int i = 0; i + 1 < ns; i++) e[i] = bb[i];
	if (tridiag_full(ns, d, e, zz.data()) != 0) return fail(LPP_ERR_STATE, "tridiagonal QL did not converge");
	for (int k = 0; k < nstates; k++) energies[k] = d[k];
	if (z_host) {
		if (!h->gs) CKR(dev_alloc(h, &h->gs, h->nloc));
		std::vector<double> coef(ns), a2(ns + 1), b2(ns + 1);
		for (int k = nstates - 1; k >= 0; k--) {                       // state 0 last: it stays in h->gs
			for (int j = 0; j < ns; j++) coef[j] = zz[(size_t)j * ns + k];
			CK(cudaMemsetAsync(h->gs, 0, sizeof(double) * h->nloc, h->stream));
			CKR(load_init(h, p, init_host, 0));
			int ns2 = 0;
			CKR(lanczos_loop(h, p, ns, false, coef.data(), h->gs, a2.data(), b2.data(), &ns2, nullptr, nullptr));
			CK(cudaStreamSynchronize(h->stream));
			CK(cudaMemcpy(z_host + (uint64_t)k * h->rows + h->row0, h->gs, sizeof(double) * h->nloc, cudaMemcpyDeviceToHost));
		}
	}
	if (nsteps) *nsteps = ns;
	return 0;
}

// ------------------------------------------------------------------ operator application / vectors
// dst.modified (+)= factor * O |srcvec>, srcvec = a vector on src's sector (local rows)
static int apply_op_vec(lpp_handle* src, const double* srcvec_local, lpp_handle* dst, int32_t op, int32_t site, int32_t spin, int32_t orb,
                        double factor, int32_t accumulate)
{
	NvtxRange nvtx("lpp:apply_op");
	if (!src || !dst) return fail(LPP_ERR_ARG, "null argument");
	const int model = src->md.model;
	if (model != dst->md.model) return fail(LPP_ERR_ARG, "source and destination models differ");
	if (orb < 0 || orb >= src->md.orbitals) return fail(LPP_ERR_ARG, "bad orbital");
	if (site < 0 || site >= src->md.nsite || spin < 0 || spin > 1) return fail(LPP_ERR_ARG, "bad site/spin");
	const bool fermion = op == LPP_OP_C || op == LPP_OP_CDAGGER;
	const bool spinop = op == LPP_OP_SZ || op == LPP_OP_SPLUS || op == LPP_OP_SMINUS;
	if (model == LPP_MODEL_HUBBARD) { if (!fermion && !spinop && op != LPP_OP_N) return fail(LPP_ERR_ARG, "unsupported operator"); }
	else if (model == LPP_MODEL_HEISENBERG) { if (!spinop && op != LPP_OP_N) return fail(LPP_ERR_ARG, "Heisenberg: sz, splus, sminus, n"); }
	else if (!fermion && op != LPP_OP_SPLUS && op != LPP_OP_SMINUS) return fail(LPP_ERR_ARG, "FeAsBasedSc / Tj1Orbital: c, cdagger, splus, sminus");
	if (!srcvec_local) return fail(LPP_ERR_STATE, "source handle holds no ground-state vector");
	if (src->desc.nranks != dst->desc.nranks || src->desc.rank != dst->desc.rank || src->device != dst->device)
		return fail(LPP_ERR_ARG, "source and destination must share device and sharding");
	if (src->desc.nranks > 1 && (spin != 0 || op == LPP_OP_SPLUS || op == LPP_OP_SMINUS || model == LPP_MODEL_HEISENBERG))
		return fail(LPP_ERR_ARG, "row-sharded operator application is local for spin-up c/cdagger and for sz/n only");
	// hasNewParts: HubbardOneOrbital.h:212-257, BasisFeAsBasedSc.h:305-326, TjMultiOrb.h:538-557, Heisenberg.h:218-240
	int eu = src->md.nup, ed = src->md.ndn;
	if (fermion) { const int dup = (op == LPP_OP_C) ? -1 : 1; if (spin == 0) eu += dup; else ed += dup; }
	else if (op == LPP_OP_SPLUS) { eu += 1; if (model != LPP_MODEL_HEISENBERG) ed -= 1; }
	else if (op == LPP_OP_SMINUS) { eu -= 1; if (model != LPP_MODEL_HEISENBERG) ed += 1; }
	if (dst->md.nup != eu || (model != LPP_MODEL_HEISENBERG && dst->md.ndn != ed) || dst->md.nsite != src->md.nsite)
		return fail(LPP_ERR_ARG, "destination sector does not match operator (hasNewParts)");
	if (model == LPP_MODEL_FEAS && dst->md.orbitals != src->md.orbitals) return fail(LPP_ERR_ARG, "orbital count differs");
	CK(cudaSetDevice(dst->device));
	if (!dst->modified) {
		CKR(dev_alloc(dst, &dst->modified, dst->nloc));
		accumulate = 0;
	}
	if (!accumulate) CK(cudaMemsetAsync(dst->modified, 0, sizeof(double) * dst->nloc, dst->stream));
	CK(cudaStreamSynchronize(src->stream));
	// source vector is indexed globally inside the kernel: shift the local pointer by the shard's first row
	const double* srcv = srcvec_local - src->row0;
	lpp_launch_apply_op(src->md, dst->md, op, site, spin, orb, factor, srcv, dst->modified, dst->row0, dst->nloc, dst->stream);
	dst->launches += 1;
	CK(cudaStreamSynchronize(dst->stream));
	CK(cudaGetLastError());
	return 0;
}

extern "C" int lpp_apply_op(lpp_handle* src, lpp_handle* dst, int32_t op, int32_t site, int32_t spin, int32_t orb,
                            double factor, int32_t accumulate)
{
	if (!src || !dst) return fail(LPP_ERR_ARG, "null argument");
	return apply_op_vec(src, src->gs, dst, op, site, spin, orb, factor, accumulate);
}

// Engine::manyPoint (Engine.h:341-389) with bra = ket = ground state: the operators are applied one after the other,
// tmp_0 = |gs>, tmp_k = O_k tmp_{k-1} on the sector hasNewParts gives (chain[k], created by the caller like Engine's getNeededBasis),
// result = <gs | tmp_n> when the last sector is the first one again.  Nothing leaves the device.
extern "C" int lpp_many_point(lpp_handle* const* chain, int32_t nops, const int32_t* ops, const int32_t* sites, const int32_t* spins,
                              const int32_t* orbs, double* result)
{
	if (!chain || nops < 1 || !ops || !sites || !spins || !orbs || !result) return fail(LPP_ERR_ARG, "bad argument");
	for (int k = 0; k <= nops; k++)
		if (!chain[k]) return fail(LPP_ERR_ARG, "null handle in the sector chain");
	lpp_handle* h0 = chain[0];
	lpp_handle* hn = chain[nops];
	if (!h0->gs) return fail(LPP_ERR_STATE, "the first handle holds no ground-state vector");
	if (hn->md.nup != h0->md.nup || hn->md.ndn != h0->md.ndn || hn->rows != h0->rows)
		return fail(LPP_ERR_ARG, "the operator string does not return to the sector of the ground state");
	const double* cur = h0->gs;
	ScopedDev<double> tmp;                                           // copy of the source when an operator maps a handle onto itself
	for (int k = 1; k <= nops; k++) {
		lpp_handle* src = chain[k - 1];
		lpp_handle* dst = chain[k];
		CK(cudaSetDevice(dst->device));
		if (dst->modified && cur == dst->modified) {
			if (!tmp && tmp.alloc(dst->nloc) != cudaSuccess) { cudaGetLastError(); return fail(LPP_ERR_CUDA, "many_point: out of device memory"); }
			CK(cudaMemcpyAsync(tmp, cur, sizeof(double) * dst->nloc, cudaMemcpyDeviceToDevice, dst->stream));
			CK(cudaStreamSynchronize(dst->stream));
			cur = tmp;
		}
		const int rc = apply_op_vec(src, cur, dst, ops[k - 1], sites[k - 1], spins[k - 1], orbs[k - 1], 1.0, 0);
		if (rc != 0) return rc;
		cur = dst->modified;
	}
	CK(cudaSetDevice(h0->device));
	CKR(ensure_partials(h0, lpp_vec_blocks(h0->nloc)));
	lpp_launch_dot(h0->gs, cur, h0->nloc, h0->partials, h0->stream);
	h0->launches += 1;
	double v = 0;
	const int rc = reduce_scalar(h0, lpp_vec_blocks(h0->nloc), &v);
	if (rc != 0) return rc;
	*result = v;
	return 0;
}

// Engine::twoPoint (Engine.h:262-331) for the operators lpp_apply_op knows: m_i = O_{i, spin, orb_i} |src.groundstate> on the
// sector of `dst` for every site i, result[i * nsite + j] = <m_j(orb_j) | m_i(orb_i)> (bra = ket = ground state)
extern "C" int lpp_two_point(lpp_handle* src, lpp_handle* dst, int32_t op, int32_t spin, int32_t orb_i, int32_t orb_j, double* result)
{
	if (!src || !dst || !result) return fail(LPP_ERR_ARG, "null argument");
	const int nsite = src->md.nsite;
	CK(cudaSetDevice(dst->device));
	const uint64_t n = dst->nloc, stride = std::max<uint64_t>(n, 1);
	const int nsets = (orb_i == orb_j) ? 1 : 2;                     // second set of modified states when the orbitals differ
	double* vecs = nullptr;
	if (cudaMalloc((void**)&vecs, sizeof(double) * stride * nsite * nsets) != cudaSuccess) {
		cudaGetLastError();
		return fail(LPP_ERR_CUDA, "two_point: out of device memory for nsite modified states");
	}
	double* keep = dst->modified;
	int rc = 0;
	for (int set = 0; set < nsets && rc == 0; set++)
		for (int i = 0; i < nsite && rc == 0; i++) {
			dst->modified = vecs + (uint64_t)(set * nsite + i) * stride;   // lpp_apply_op writes into the handle's modified vector
			rc = lpp_apply_op(src, dst, op, i, spin, set == 0 ? orb_i : orb_j, 1.0, 0);
		}
	dst->modified = keep;
	if (rc != 0) { cudaFree(vecs); return rc; }
	const int nt = (nsite + 3) / 4, npb = lpp_vec_blocks(n * 2);
	int rcp = ensure_partials(dst, npb * 16);
	if (rcp != 0) { cudaFree(vecs); return rcp; }
	const double* vi = vecs;
	const double* vj = vecs + (uint64_t)(nsets - 1) * nsite * stride;
	for (int ti = 0; ti < nt && rc == 0; ti++)
		for (int tj = 0; tj < nt && rc == 0; tj++) {
			// tile (ti, tj): rows from the orb_i set of modified states, columns from the orb_j set
			lpp_launch_gram_tile(vi, vj, stride, n, nsite, ti, tj, dst->partials, dst->stream);
			dst->launches += 1;
			double sums[16];
			double* const all_partials = dst->partials;
			for (int q = 0; q < 16 && rc == 0; q += 4) {                  // scal_dev holds 8 values: reduce the 16 sums in four rounds
				dst->partials = all_partials + (uint64_t)q * npb;         // reduce_scalars sums rows [0, 4) of h->partials (and all-reduces)
				rc = reduce_scalars(dst, npb, 4, sums + q);
			}
			dst->partials = all_partials;
			for (int a = 0; a < 4; a++)
				for (int b = 0; b < 4; b++) {
					const int i = ti * 4 + a, j = tj * 4 + b;
					if (i < nsite && j < nsite) result[i * nsite + j] = sums[a * 4 + b];
				}
		}
	cudaFree(vecs);
	return rc;
}

// Engine::measure (Engine.h:208-249) with bra = ket = ground state: <gs| op_0[site_0] ... op_{n-1}[site_{n-1}] |gs> through
// ModelBase::rahulMethod semantics (the rightmost operator acts first)
extern "C" int lpp_measure(lpp_handle* h, int32_t nops, const int32_t* labels, const int32_t* dofs, const int32_t* transposes,
                           const int32_t* sites, double* result)
{
	if (!h || !labels || !dofs || !transposes || !sites || !result) return fail(LPP_ERR_ARG, "null argument");
	if (nops < 1 || nops > LPP_MAX_MEASURE_OPS) return fail(LPP_ERR_ARG, "between 1 and 8 operators");
	if (h->md.model == LPP_MODEL_HEISENBERG) return fail(LPP_ERR_ARG, "lpp_measure: fermionic models (up / down words)");
	if (!h->gs) return fail(LPP_ERR_STATE, "handle holds no ground-state vector");
	LppMeasureOps ops;
	memset(&ops, 0, sizeof(ops));
	ops.n = nops;
	bool diagonal = true;
	for (int i = 0; i < nops; i++) {
		if (labels[i] < 0 || labels[i] > 3 || dofs[i] < 0 || dofs[i] > 1 || sites[i] < 0 || sites[i] >= h->md.nbits)
			return fail(LPP_ERR_ARG, "bad operator label / dof / site");
		ops.label[i] = labels[i]; ops.dof[i] = dofs[i]; ops.transpose[i] = transposes[i] ? 1 : 0; ops.site[i] = sites[i];
		diagonal = diagonal && labels[i] != 3;
	}
	if (h->desc.nranks > 1 && !diagonal) return fail(LPP_ERR_ARG, "row-sharded measure supports identity / n / sz products");
	CK(cudaSetDevice(h->device));
	const int npb = lpp_vec_blocks(h->nloc * 2);
	CKR(ensure_partials(h, npb));
	lpp_launch_measure(h->md, ops, h->gs, h->gs, h->row0, h->nloc, h->partials, h->stream);
	h->launches += 1;
	CKR(reduce_scalar(h, npb, result));
	return 0;
}

extern "C" int lpp_get_vector(lpp_handle* h, int32_t which, double* out_host)
{
	if (!h || !out_host) return fail(LPP_ERR_ARG, "null argument");
	const double* v = which == 0 ? h->gs : h->modified;
	if (!v) return fail(LPP_ERR_STATE, "vector not available");
	CK(cudaSetDevice(h->device));
	CK(cudaStreamSynchronize(h->stream));
	CK(cudaMemcpy(out_host + h->row0, v, sizeof(double) * h->nloc, cudaMemcpyDeviceToHost));
	return 0;
}

extern "C" int lpp_set_groundstate(lpp_handle* h, const double* z_host)
{
	if (!h || !z_host) return fail(LPP_ERR_ARG, "null argument");
	CK(cudaSetDevice(h->device));
	if (!h->gs) CKR(dev_alloc(h, &h->gs, h->nloc));
	CK(cudaMemcpy(h->gs, z_host + h->row0, sizeof(double) * h->nloc, cudaMemcpyHostToDevice));
	return 0;
}

// ------------------------------------------------------------------ communicator
extern "C" int lpp_comm_unique_id(uint8_t id[128])
{
	CKR(nccl_load());
	ncclUniqueId u;
	CKN(g_nccl.GetUniqueId(&u));
	memcpy(id, u.internal, 128);
	return 0;
}

extern "C" int lpp_comm_init(lpp_handle* h, const uint8_t id[128])
{
	if (!h || !id) return fail(LPP_ERR_ARG, "null argument");
	CKR(nccl_load());
	CK(cudaSetDevice(h->device));
	ncclUniqueId u;
	memcpy(u.internal, id, 128);
	CKN(g_nccl.CommInitRank(&h->comm, h->desc.nranks, u, h->desc.rank));
	return 0;
}

extern "C" int lpp_comm_share(lpp_handle* h, const lpp_handle* parent)
{
	if (!h || !parent) return fail(LPP_ERR_ARG, "null argument");
	if (!parent->comm) return fail(LPP_ERR_STATE, "parent handle has no communicator");
	if (h->desc.nranks != parent->desc.nranks || h->desc.rank != parent->desc.rank || h->device != parent->device)
		return fail(LPP_ERR_ARG, "handles must share ranks and device");
	if (h->comm && !h->comm_borrowed) return fail(LPP_ERR_STATE, "handle already owns a communicator");
	h->comm = parent->comm;
	h->comm_borrowed = true;
	return 0;
}

// peer-memory exchange: export the handles of this rank's column-shard buffers / map everybody's
extern "C" int lpp_p2p_export(lpp_handle* h, int32_t kernel, uint8_t handles[128])
{
	if (!h || !handles) return fail(LPP_ERR_ARG, "null argument");
	CK(cudaSetDevice(h->device));
	h->p2p_requested = true;
	CKR(ensure_two_layout(h, kernel));
	if (h->two_layout != 1) return fail(LPP_ERR_STATE, "two-layout sharding does not apply to this handle");
	// 128 bytes: the IPC handle of the slab [ycol | xcol | vx | vy] and the offsets of its parts (in doubles)
	cudaIpcMemHandle_t a;
	CK(cudaIpcGetMemHandle(&a, h->slab));
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
	memset(handles, 0, 128);
	memcpy(handles, &a, 64);
	const uint64_t hdr[5] = {0x3142414c5350504cull /* "LPPSLAB1" */, h->slab_off_xcol, h->slab_off_vx, h->slab_off_vy, h->slab_off_psx};
	memcpy(handles + 64, hdr, sizeof(hdr));
	return 0;
}

extern "C" int lpp_p2p_import(lpp_handle* h, const uint8_t* all_handles)
{
	if (!h || !all_handles) return fail(LPP_ERR_ARG, "null argument");
	if (h->two_layout != 1) return fail(LPP_ERR_STATE, "lpp_p2p_export must succeed first");
	CK(cudaSetDevice(h->device));
	bool pull = h->slab_off_vx != 0;
	for (int q = 0; q < h->desc.nranks; q++) {
		if (q == h->desc.rank) {
			h->peer_ycol.p[q] = h->ycol;
			h->peer_xcol.p[q] = h->xcol;
			h->peer_vx.p[q] = h->vx;
			h->peer_vy.p[q] = h->vy;
			h->peer_psx.p[q] = h->slab + h->slab_off_psx;
			continue;
		}
		cudaIpcMemHandle_t a;
		uint64_t hdr[5];
		memcpy(&a, all_handles + (size_t)q * 128, 64);
		memcpy(hdr, all_handles + (size_t)q * 128 + 64, sizeof(hdr));
		if (hdr[0] != 0x3142414c5350504cull) return fail(LPP_ERR_ARG, "peer handle block has the wrong format");
		void* pa = nullptr;
		CK(cudaIpcOpenMemHandle(&pa, a, cudaIpcMemLazyEnablePeerAccess));
		h->ipc_opened.push_back(pa);
		double* base = (double*)pa;
		h->peer_ycol.p[q] = base;
		h->peer_xcol.p[q] = base + hdr[1];
		h->peer_vx.p[q] = hdr[2] ? base + hdr[2] : nullptr;
		h->peer_vy.p[q] = hdr[3] ? base + hdr[3] : nullptr;
		h->peer_psx.p[q] = base + hdr[4];
		pull = pull && hdr[2] && hdr[3];
	}
	h->p2p = 1;
	// scalar all-reduces over peer memory (default; LPP_PSX=0 keeps NCCL for them)
	h->psx = !(getenv("LPP_PSX") && getenv("LPP_PSX")[0] == '0');
	h->psx_seq = 0;
	if (h->psx && !h->psx_err) {
		// pinned, device-visible: the exchange kernel raises it when a peer did not show up within ~2 s
		CK(cudaHostAlloc((void**)&h->psx_err, sizeof(int), cudaHostAllocMapped));
		*h->psx_err = 0;
	}
	// opt-in: on 8 x B200 the copy engines PULL 145 MB per rank in 0.42 ms against 0.29 ms for the push form, and without the
	// all-reduce the ranks drift apart (1.44 ms per iteration against 1.15 ms); on 2 GPUs the two forms are equal
	h->pull_pack = pull && (getenv("LPP_PULL_PACK") && getenv("LPP_PULL_PACK")[0] == '1');
	return 0;
}

// diagnostic / test hook: all-reduce (sum) of n <= 4 host doubles over the ranks of a handle through the path the sharded Krylov loop
// uses for its scalars (peer-memory exchange when the handle has it, NCCL otherwise)
extern "C" int lpp_allreduce_selftest(lpp_handle* h, double* inout, int32_t n)
{
	if (!h || !inout || n < 0 || n > 4) return fail(LPP_ERR_ARG, "bad argument");
	if (h->desc.nranks > 1 && !h->comm) return fail(LPP_ERR_STATE, "nranks>1 but lpp_comm_init has not been called");
	CK(cudaSetDevice(h->device));
	// eight all-reduces of the same input back to back (no host synchronisation in between, like the Krylov loop), interleaved
	// with empty ones (the barrier form); every result must be the same
	double in[4] = {0, 0, 0, 0}, out[8][4];
	for (int i = 0; i < n; i++) in[i] = inout[i];
	double* dev = nullptr;
	CKR(dev_alloc(h, &dev, 8 * 4));
	for (int r = 0; r < 8; r++) {
		CK(cudaMemcpyAsync(dev + 4 * r, in, sizeof(double) * 4, cudaMemcpyHostToDevice, h->stream));
		if (h->desc.nranks > 1) {
			CKR(allreduce_small(h, dev + 4 * r, n, h->stream));
			if (r % 3 == 1) CKR(allreduce_small(h, h->scal_dev + 4, 0, h->stream));
		}
	}
	CK(cudaMemcpyAsync(&out[0][0], dev, sizeof(double) * 32, cudaMemcpyDeviceToHost, h->stream));
	CK(cudaStreamSynchronize(h->stream));
	dev_free(h, dev);
	for (int r = 1; r < 8; r++)
		for (int i = 0; i < n; i++)
			if (out[r][i] != out[0][i]) return fail(LPP_ERR_STATE, "back-to-back all-reduces disagree");
	for (int i = 0; i < n; i++) inout[i] = out[0][i];
	if (h->psx_err && *h->psx_err) return fail(LPP_ERR_STATE, "peer-memory scalar exchange timed out (a rank did not arrive)");
	return 0;
}

// ------------------------------------------------------------------ measurement hooks
extern "C" int lpp_bench_spmv(lpp_handle* h, int32_t kernel, int32_t iters, int32_t warmup, lpp_timing* t)
{
	if (!h || !t || iters < 1) return fail(LPP_ERR_ARG, "bad argument");
	CK(cudaSetDevice(h->device));
	CKR(ensure_vectors(h, kernel));
	lpp_launch_fill_random(h->vy, h->row0, h->nloc, 42, h->stream);
	CK(cudaMemsetAsync(h->vx, 0, sizeof(double) * h->nloc, h->stream));
	// peers read this rank's vy (pull pack): nobody starts before every rank's vector is filled
	if (h->desc.nranks > 1 && h->comm) CKN(g_nccl.AllReduce(h->scal_dev + 4, h->scal_dev + 4, 1, kNcclFloat64, kNcclSum, h->comm, h->stream));
	const double* src = nullptr;
	if (h->two_layout != 1) CKR(gather_full(h, h->vy, &src));
	auto one = [&]() -> int {
		if (h->two_layout == 1) return spmv_two_layout(h, 1.0, 1.0, h->vx, h->vy, false, nullptr);
		return do_spmv(h, kernel, 1.0, 1.0, h->vx, src, false, nullptr);
	};
	for (int i = 0; i < warmup; i++) CKR(one());
	CK(cudaStreamSynchronize(h->stream));
	int64_t l0 = h->launches;
	CK(cudaEventRecord(h->ev0, h->stream));
	for (int i = 0; i < iters; i++) CKR(one());
	CK(cudaEventRecord(h->ev1, h->stream));
	CK(cudaEventSynchronize(h->ev1));
	float ms = 0;
	CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
	t->spmv_ms = ms / iters;
	t->iter_ms = 0;
	t->launches = h->launches - l0;
	return 0;
}

extern "C" int lpp_bench_lanczos(lpp_handle* h, const lpp_solver_params* p, int32_t iters, int32_t warmup, lpp_timing* t)
{
	if (!h || !p || !t || iters < 1) return fail(LPP_ERR_ARG, "bad argument");
	CK(cudaSetDevice(h->device));
	lpp_solver_params q = *p;
	q.eps = 0;
	CKR(load_init(h, &q, nullptr, 0));
	int total = warmup + iters;
	if ((uint64_t)total > h->rows) return fail(LPP_ERR_ARG, "warmup+iters exceeds the Hilbert-space dimension");
	std::vector<double> a(total + 1), b(total + 1);
	LoopTiming tm;
	tm.from = warmup;
	tm.to = total;
	int ns = 0;
	CKR(lanczos_loop(h, &q, total, false, nullptr, nullptr, a.data(), b.data(), &ns, nullptr, &tm));
	CK(cudaEventSynchronize(h->ev1));
	float ms = 0;
	CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
	t->iter_ms = ms / iters;
	t->spmv_ms = 0;
	t->launches = tm.launches_at_to - tm.launches_at_from;
	return 0;
}
