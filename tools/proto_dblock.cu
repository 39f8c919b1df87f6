// tools/proto_dblock.cu -- standalone timing harness of the BLOCK down sweep (csrc/lpp_dblock_kernel.cuh)  x = beta x + alpha (D + 1 (x) T_dn) y.
//
// Idea: pick two disjoint site sets F1, F2 with no hopping between them.  Pass 1 groups the down states by their occupation of
// F1 ("blocks": every hop that does not touch F1 stays inside its block); pass 2 groups them by F2 and applies the hops that
// touch F1 (none of them touches F2, so they stay inside the F2 blocks).  A tile = (block, 16 columns) of y sits in shared
// memory; every hop operand is a conflict-free 16-byte shared-memory load (8 lanes per state read one 128-byte line).  There
// are no operands outside the tile, so the only global traffic is y once and x read+write per pass.
// A persistent grid takes tiles from a ticket counter in panel-major order, pass 2 of a panel LAG panels behind pass 1, so
// the panel's x and y stay L2 resident between the passes.
//
// Options: --layout 1 (one CTA of 1024 threads per SM), --passes 3 (three disjoint site sets), --lag <panels>, --dot (fused
// dot product in the last pass), --cols <n> (fewer columns), --chain <sites> (open chain instead of the 4 x 4 torus), --iters <n>.
// Compile-time variants: -DDB_FILL_MODE=2 (bulk-copy tile lines), -DDB_RED_LATER_PASSES=0, -DDB_SKIP_PADDING=1, -DDB_PROFILE
// (cycle counters per phase).  The numbers of profiles/README.md ("Two CTAs per SM for the block down sweep") come from here.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/proto_dblock tools/proto_dblock.cu
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(1); } } while (0)

#include "../lanczosplusplus_b200/csrc/lpp_dblock_kernel.cuh"

// ---------------------------------------------------------------------------------------------------------------
// host: basis, hops
// ---------------------------------------------------------------------------------------------------------------
static std::vector<uint32_t> colex_basis(int nsite, int npart)
{
	std::vector<uint32_t> w;
	for (uint32_t s = 0; s < (1u << nsite); s++)
		if (__builtin_popcount(s) == npart) w.push_back(s);
	return w;
}

__global__ void k_ref(const double* __restrict__ y, double* __restrict__ x, uint64_t pitch, uint64_t ncols, uint64_t n2, const uint32_t* __restrict__ idx,
                      const double* __restrict__ val, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ w1, const uint32_t* __restrict__ w2,
                      double U0, double alpha, double beta)
{
	const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, d = blockIdx.y;
	if (u >= ncols) return;
	double acc = U0 * (double)__popc(w1[u] & w2[d]) * y[d * pitch + u];
	for (uint32_t k = 0; k < cnt[d]; k++) acc += val[(uint64_t)k * n2 + d] * y[(uint64_t)idx[(uint64_t)k * n2 + d] * pitch + u];
	x[d * pitch + u] = beta * x[d * pitch + u] + alpha * acc;
}

__global__ void k_fill(double* v, uint64_t n, uint64_t seed)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint64_t z = (i + seed * 0x9E3779B97F4A7C15ull) + 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	z ^= z >> 31;
	v[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

int main(int argc, char** argv)
{
	int nx = 4, ny = 4, npart = 8, iters = 10;
	uint64_t ncols_arg = 0;
	int lag = 8, layout = 0, want_dot = 0, passes = 0;
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "--cols")) ncols_arg = strtoull(argv[++i], 0, 10);
		else if (!strcmp(argv[i], "--iters")) iters = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--lag")) lag = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--layout")) layout = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--dot")) want_dot = 1;
		else if (!strcmp(argv[i], "--passes")) passes = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--chain")) { nx = atoi(argv[++i]); ny = 1; npart = nx / 2; }
	}
	const int nsite = nx * ny;
	std::vector<uint32_t> words = colex_basis(nsite, npart);
	const uint64_t n2 = words.size();
	const uint64_t ncols = ncols_arg ? ncols_arg : n2;
	std::vector<double> hop((size_t)nsite * nsite, 0.0);
	auto site = [&](int x, int y) { return ((x + nx) % nx) + nx * ((y + ny) % ny); };
	for (int y = 0; y < ny; y++)
		for (int x = 0; x < nx; x++) {
			int i = site(x, y);
			int nb[2] = {site(x + 1, y), site(x, y + 1)};
			for (int b = 0; b < 2; b++) {
				if (ny == 1 && b == 1) continue;
				if (ny == 1 && x == nx - 1) continue;   // open chain
				int j = nb[b];
				if (j == i) continue;
				hop[i * nsite + j] = -1.0;
				hop[j * nsite + i] = -1.0;
			}
		}
	// rank lookup
	std::vector<uint32_t> lut(1u << nsite, 0xffffffffu);
	for (uint64_t s = 0; s < n2; s++) lut[words[s]] = (uint32_t)s;
	// ELL hop table
	std::vector<std::vector<std::pair<uint32_t, double>>> hl(n2);
	int width = 0;
	for (uint64_t s = 0; s < n2; s++) {
		uint32_t w = words[s];
		for (int i = 0; i < nsite; i++)
			for (int j = 0; j < nsite; j++) {
				double h = hop[i * nsite + j];
				if (h == 0 || !((w >> i) & 1) || ((w >> j) & 1)) continue;
				int lo = std::min(i, j), hi = std::max(i, j);
				uint32_t between = ((1u << hi) - 1) & ~((1u << (lo + 1)) - 1);
				double sg = (__builtin_popcount(w & between) & 1) ? -1.0 : 1.0;
				hl[s].push_back({lut[w ^ (1u << i) ^ (1u << j)], h * sg});
			}
		width = std::max<int>(width, (int)hl[s].size());
	}
	std::vector<uint32_t> idx((size_t)width * n2), cnt(n2);
	std::vector<double> val((size_t)width * n2, 0.0);
	double meanh = 0;
	for (uint64_t s = 0; s < n2; s++) {
		cnt[s] = (uint32_t)hl[s].size();
		meanh += cnt[s];
		for (int k = 0; k < width; k++) {
			idx[(size_t)k * n2 + s] = k < (int)cnt[s] ? hl[s][k].first : (uint32_t)s;
			val[(size_t)k * n2 + s] = k < (int)cnt[s] ? hl[s][k].second : 0.0;
		}
	}
	printf("basis %llu states, width %d, mean hops %.3f, cols %llu\n", (unsigned long long)n2, width, meanh / n2, (unsigned long long)ncols);

	int dev = 0, maxsm = 0, maxblk = 0, nsm = 0;
	CK(cudaGetDevice(&dev));
	CK(cudaDeviceGetAttribute(&maxblk, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
	CK(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
	CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
	std::vector<double> dv2(n2, 0.0);
	DbHostPlan hp;
	std::string err;
	if (!db_build_host_plan(words.data(), n2, nsite, idx.data(), val.data(), cnt.data(), width, dv2.data(), (size_t)maxblk, (size_t)maxsm, layout, passes, &hp, &err)) {
		printf("plan failed: %s\n", err.c_str());
		return 1;
	}
	printf("%d passes, %d CTA(s) x %d threads per SM, smem %zu, lag %d, fill mode %d\n", hp.npass, hp.ctas_per_sm, hp.threads,
	       hp.smem_bytes + (size_t)hp.max_pos * 8, lag, DB_FILL_MODE);
	double slots = 0;
	for (int k = 0; k < hp.npass; k++) {
		printf("  pass %d: F %#x, %zu blocks, max %u pos, hops %.3f/state, executed slots %.3f/state\n", k + 1, hp.fmask[k], hp.pass[k].blocks.size(),
		       hp.pass[k].max_pos, hp.pass[k].mean_hops, (double)hp.pass[k].exec_slots / n2);
		slots += (double)hp.pass[k].exec_slots / n2;
	}
	printf("executed state-slots per state (padding included): %.3f\n", slots);

	// device data
	const uint64_t pitch = ncols;
	const uint64_t nel = n2 * pitch;
	double *dy, *dx, *dxr;
	CK(cudaMalloc(&dy, nel * 8));
	CK(cudaMalloc(&dx, nel * 8));
	CK(cudaMalloc(&dxr, nel * 8));
	k_fill<<<(unsigned)((nel + 255) / 256), 256>>>(dy, nel, 42);
	k_fill<<<(unsigned)((nel + 255) / 256), 256>>>(dx, nel, 7);
	CK(cudaMemcpy(dxr, dx, nel * 8, cudaMemcpyDeviceToDevice));
	uint32_t *didx, *dcnt, *dw1, *dw2;
	double* dval;
	CK(cudaMalloc(&didx, idx.size() * 4));
	CK(cudaMalloc(&dval, val.size() * 8));
	CK(cudaMalloc(&dcnt, n2 * 4));
	CK(cudaMalloc(&dw2, n2 * 4));
	CK(cudaMalloc(&dw1, ncols * 4));
	CK(cudaMemcpy(didx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dval, val.data(), val.size() * 8, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dcnt, cnt.data(), n2 * 4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(dw2, words.data(), n2 * 4, cudaMemcpyHostToDevice));
	std::vector<uint32_t> w1(ncols);
	for (uint64_t u = 0; u < ncols; u++) w1[u] = words[u % n2];
	CK(cudaMemcpy(dw1, w1.data(), ncols * 4, cudaMemcpyHostToDevice));
	std::vector<double> dv1h(ncols, 0.0);
	double* ddv1;
	CK(cudaMalloc(&ddv1, ncols * 8));
	CK(cudaMemcpy(ddv1, dv1h.data(), ncols * 8, cudaMemcpyHostToDevice));

	DbDevPlan dp;
	if (!db_upload_plan(hp, &dp, &err)) { printf("upload failed: %s\n", err.c_str()); return 1; }
	dp.lag = lag;

	const double alpha = 0.37, beta = -0.81, U0 = 4.0;
	// reference
	{
		dim3 g((unsigned)((ncols + 255) / 256), (unsigned)n2);
		k_ref<<<g, 256>>>(dy, dxr, pitch, ncols, n2, didx, dval, dcnt, dw1, dw2, U0, alpha, beta);
		CK(cudaGetLastError());
	}
	DbArgs a;
	a.x = dx; a.y = dy; a.pitch = pitch; a.ncols = ncols; a.alpha = alpha; a.beta = beta; a.U0 = U0; a.w1 = dw1; a.dv1 = ddv1; a.tmag = 1.0;
	a.dot_partials = nullptr;
	if (want_dot) {
		const size_t np = ((ncols + DB_COLS - 1) / DB_COLS) * hp.pass[hp.npass - 1].blocks.size();
		CK(cudaMalloc(&a.dot_partials, np * 8));
		printf("fused dot: %zu partial sums\n", np);
	}
	a.alpha_dev = nullptr;
	a.beta_dev = nullptr;
	if (db_launch(dp, a, nsm, 0)) { printf("launch failed\n"); return 1; }
	CK(cudaDeviceSynchronize());
	// compare
	{
		std::vector<double> hx(nel > (1ull << 27) ? (1ull << 27) : nel), hr(hx.size());
		double maxd = 0, maxv = 0;
		for (int part = 0; part < 3; part++) {
			uint64_t off = part == 0 ? 0 : part == 1 ? (nel - hx.size()) / 2 : nel - hx.size();
			CK(cudaMemcpy(hx.data(), dx + off, hx.size() * 8, cudaMemcpyDeviceToHost));
			CK(cudaMemcpy(hr.data(), dxr + off, hx.size() * 8, cudaMemcpyDeviceToHost));
			for (size_t i = 0; i < hx.size(); i++) { maxd = std::max(maxd, fabs(hx[i] - hr[i])); maxv = std::max(maxv, fabs(hr[i])); }
		}
		printf("check: max |x - ref| = %.3e (max |ref| %.3f)\n", maxd, maxv);
	}
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	for (int i = 0; i < 3; i++) db_launch(dp, a, nsm, 0);
	CK(cudaEventRecord(e0));
	for (int i = 0; i < iters; i++) db_launch(dp, a, nsm, 0);
	CK(cudaEventRecord(e1));
	CK(cudaEventSynchronize(e1));
	float ms = 0;
	CK(cudaEventElapsedTime(&ms, e0, e1));
	ms /= iters;
#ifdef DB_PROFILE
	{
		long long* dprof;
		const int nctas = nsm * hp.ctas_per_sm;
		CK(cudaMalloc(&dprof, nctas * 8 * sizeof(long long)));
		CK(cudaMemset(dprof, 0, nctas * 8 * sizeof(long long)));
		dp.profile = dprof;
		db_launch(dp, a, nsm, 0);
		CK(cudaDeviceSynchronize());
		std::vector<long long> hpf(nctas * 8);
		CK(cudaMemcpy(hpf.data(), dprof, hpf.size() * sizeof(long long), cudaMemcpyDeviceToHost));
		const char* names[8] = {"ticket + fill issue", "wait for previous pass", "fill in flight", "compute pass 1 (warp 0)", "compute later passes", "end barrier", "-", "-"};
		for (int i = 0; i < 6; i++) {
			double s = 0, mx = 0;
			for (int b = 0; b < nctas; b++) { s += hpf[b * 8 + i]; mx = std::max<double>(mx, (double)hpf[b * 8 + i]); }
			printf("  phase %-26s mean %10.0f cycles per CTA, max %10.0f\n", names[i], s / nctas, mx);
		}
		dp.profile = nullptr;
	}
#endif
	printf("dblock: %.4f ms per sweep, %.1f GB/s of 24 B/elem\n", ms, 24.0 * nel / ms * 1e-6);
	{
		dim3 g((unsigned)((ncols + 255) / 256), (unsigned)n2);
		CK(cudaEventRecord(e0));
		for (int i = 0; i < 3; i++) k_ref<<<g, 256>>>(dy, dxr, pitch, ncols, n2, didx, dval, dcnt, dw1, dw2, U0, alpha, beta);
		CK(cudaEventRecord(e1));
		CK(cudaEventSynchronize(e1));
		CK(cudaEventElapsedTime(&ms, e0, e1));
		printf("naive ref kernel: %.4f ms per sweep\n", ms / 3);
	}
	return 0;
}
