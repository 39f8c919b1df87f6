// tools/microbench_bulk.cu -- can a CTA pull a COLUMN PAIR of a row-major fp64 matrix (12870 rows x 16 bytes, row stride
// 102 960 bytes) into shared memory without paying one LSU wavefront per row?  Times three ways of filling a 206 KB tile and
// one way of adding it back, per SM, on all 148 SMs at once:
//   ldg      : every thread LDG.128 + STS.128 (the baseline: uncoalesced 16-byte accesses)
//   ldgsts   : cp.async 16 bytes per thread
//   bulk     : cp.async.bulk.shared::cluster.global.mbarrier 16 bytes per thread (TMA engine, no LSU data path)
//   red      : cp.reduce.async.bulk.global.shared::cta add.f64 16 bytes per row (tile added back to a second matrix)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_bulk tools/microbench_bulk.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
#define ROWS 12870
#define THREADS 512

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase)
{
	asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(phase) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) k_fill(const double* __restrict__ y, double* __restrict__ xout, uint64_t pitch, int tiles, double* sink)
{
	extern __shared__ __align__(128) double tile[];                 // [ROWS][2]
	__shared__ __align__(8) unsigned long long bar;
	const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
	const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&bar);
	if (threadIdx.x == 0) { mbar_init(bar_s, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
	__syncthreads();
	double acc = 0;
	uint32_t phase = 0;
	for (int t = 0; t < tiles; t++) {
		const uint64_t col = ((uint64_t)blockIdx.x * tiles + t) * 2 % (pitch - 2);
		const double* src = y + (col & ~1ull);
		if (MODE == 0) {
			for (int r = threadIdx.x; r < ROWS; r += THREADS) {
				const double2 v = *reinterpret_cast<const double2*>(src + (uint64_t)r * pitch);
				reinterpret_cast<double2*>(tile)[r] = v;
			}
			__syncthreads();
		} else if (MODE == 1) {
			for (int r = threadIdx.x; r < ROWS; r += THREADS)
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tile_s + r * 16u), "l"(src + (uint64_t)r * pitch));
			asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
			__syncthreads();
		} else {
			if (threadIdx.x == 0) mbar_expect(bar_s, ROWS * 16u);
			__syncthreads();
			for (int r = threadIdx.x; r < ROWS; r += THREADS)
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(tile_s + r * 16u),
				             "l"(src + (uint64_t)r * pitch), "r"(bar_s)
				             : "memory");
			mbar_wait(bar_s, phase);
			phase ^= 1;
		}
		acc += tile[(threadIdx.x * 37 + t) % (2 * ROWS)];
		if (MODE == 3) {
			// add the tile back: one 16-byte bulk reduction per row
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			__syncthreads();
			double* dst = xout + (col & ~1ull);
			for (int r = threadIdx.x; r < ROWS; r += THREADS)
				asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 16;" ::"l"(dst + (uint64_t)r * pitch),
				             "r"(tile_s + r * 16u)
				             : "memory");
			asm volatile("cp.async.bulk.commit_group;\ncp.async.bulk.wait_group.read 0;" ::: "memory");
		}
		__syncthreads();
	}
	if (acc == 1.2345) sink[0] = acc;
}

// Row SEGMENTS instead of column pairs: what the streamed version of the staged down sweep would do.  Every CTA pulls `nseg`
// segments of SEGB bytes (one per source row, row stride = pitch) per batch into a shared-memory ring, 400 batches.
//   MODE 0: one elected lane per segment issues cp.async.bulk (SEGB bytes each)      MODE 1: cp.async 16 bytes per thread
template <int MODE, int SEGB>
__global__ void __launch_bounds__(256, 1) k_segments(const double* __restrict__ y, uint64_t pitch, int nseg, int batches, double* sink)
{
	extern __shared__ __align__(128) double ring[];                 // [nseg][SEGB / 8]
	__shared__ __align__(8) unsigned long long bar;
	const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
	const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(&bar);
	if (threadIdx.x == 0) { mbar_init(bar_s, 1); asm volatile("fence.mbarrier_init.release.cluster;"); }
	__syncthreads();
	double acc = 0;
	uint32_t phase = 0;
	const uint64_t col0 = ((uint64_t)blockIdx.x * 64) % (pitch - SEGB / 8);
	for (int b = 0; b < batches; b++) {
		const uint64_t row0 = ((uint64_t)b * 131 + blockIdx.x * 17) % (ROWS - nseg * 7);
		if (MODE == 0) {
			if (threadIdx.x == 0) mbar_expect(bar_s, (uint32_t)nseg * SEGB);
			__syncthreads();
			for (int q = threadIdx.x; q < nseg; q += 256)
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring_s + q * SEGB),
				             "l"(y + (row0 + (uint64_t)q * 7) * pitch + (col0 & ~1ull)), "r"(SEGB), "r"(bar_s)
				             : "memory");
			mbar_wait(bar_s, phase);
			phase ^= 1;
		} else {
			constexpr int LPS = SEGB / 16;                           // lanes per segment
			for (int q = threadIdx.x; q < nseg * LPS; q += 256) {
				const int sgm = q / LPS, piece = q % LPS;
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_s + sgm * SEGB + piece * 16),
				             "l"(y + (row0 + (uint64_t)sgm * 7) * pitch + (col0 & ~1ull) + piece * 2));
			}
			asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
			__syncthreads();
		}
		acc += ring[(threadIdx.x * 13 + b) % (nseg * SEGB / 8)];
		__syncthreads();
	}
	if (acc == 1.2345) sink[0] = acc;
}

template <int MODE, int SEGB>
static int run_segments(const double* y, uint64_t pitch, double* sink, int khz)
{
	const int nseg = 96, batches = 400;
	const size_t smem = (size_t)nseg * SEGB;
	CK(cudaFuncSetAttribute(k_segments<MODE, SEGB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	float ms = 0;
	for (int rep = 0; rep < 3; rep++) {
		CK(cudaEventRecord(e0));
		k_segments<MODE, SEGB><<<148, 256, smem>>>(y, pitch, nseg, batches, sink);
		CK(cudaEventRecord(e1));
		CK(cudaEventSynchronize(e1));
		CK(cudaGetLastError());
		cudaEventElapsedTime(&ms, e0, e1);
	}
	const double bytes = 148.0 * batches * nseg * SEGB;
	printf("%-22s %4d-byte segments, %d per batch (no overlap between batches): %.3f ms, %.1f clk/segment/SM, %.0f GB/s chip-wide\n",
	       MODE == 0 ? "cp.async.bulk" : "cp.async 16B/thread", SEGB, nseg, ms, 1e-3 * ms / batches / nseg * khz * 1e3, bytes / (ms * 1e6));
	return 0;
}

int main()
{
	const uint64_t pitch = 12870, n = (uint64_t)ROWS * pitch;
	double *y, *x, *sink;
	CK(cudaMalloc(&y, n * 8));
	CK(cudaMalloc(&x, n * 8));
	CK(cudaMalloc(&sink, 8));
	CK(cudaMemset(y, 0, n * 8));
	CK(cudaMemset(x, 0, n * 8));
	const size_t smem = (size_t)ROWS * 16;
	CK(cudaFuncSetAttribute(k_fill<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	CK(cudaFuncSetAttribute(k_fill<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	CK(cudaFuncSetAttribute(k_fill<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	CK(cudaFuncSetAttribute(k_fill<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	int dev = 0, khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
	const int tiles = 43;                                            // what one SM processes per sweep of the 4x4 problem
	const char* names[4] = {"ldg+sts 16B", "cp.async 16B", "cp.async.bulk 16B", "bulk load + bulk reduce-add f64"};
	for (int mode = 0; mode < 4; mode++) {
		for (int rep = 0; rep < 3; rep++) {
			CK(cudaEventRecord(e0));
			if (mode == 0) k_fill<0><<<148, THREADS, smem>>>(y, x, pitch, tiles, sink);
			if (mode == 1) k_fill<1><<<148, THREADS, smem>>>(y, x, pitch, tiles, sink);
			if (mode == 2) k_fill<2><<<148, THREADS, smem>>>(y, x, pitch, tiles, sink);
			if (mode == 3) k_fill<3><<<148, THREADS, smem>>>(y, x, pitch, tiles, sink);
			CK(cudaEventRecord(e1));
			CK(cudaEventSynchronize(e1));
			CK(cudaGetLastError());
			float ms = 0;
			cudaEventElapsedTime(&ms, e0, e1);
			if (rep == 2)
				printf("%-34s %.3f ms for %d tiles/SM: %.2f us/tile, %.2f clk/row at %d MHz, %.1f GB/s useful\n", names[mode], ms, tiles,
				       1e3 * ms / tiles, 1e-3 * ms / tiles / ROWS * khz * 1e3, khz / 1000, 148.0 * tiles * ROWS * 16 / (ms * 1e6));
		}
	}
	if (run_segments<0, 512>(y, pitch, sink, khz) || run_segments<1, 512>(y, pitch, sink, khz) || run_segments<0, 1024>(y, pitch, sink, khz) ||
	    run_segments<1, 1024>(y, pitch, sink, khz) || run_segments<0, 2048>(y, pitch, sink, khz))
		return 1;
	return 0;
}
