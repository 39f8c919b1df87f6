// tools/microbench.cu -- measures the three bandwidths the SpMV design trades against each other on a B200:
// HBM streaming, L2-resident coalesced reads, and shared-memory gathers (random vs conflict-free, 8 B and 16 B).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_stream_triad(double* __restrict__ x, const double* __restrict__ y, size_t n)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
		x[i] = x[i] + 1.0000001 * y[i];
}

// every thread block streams over the same `win` doubles `reps` times (coalesced) -> L2 (and L1) resident reads
__global__ void k_l2_read(const double* __restrict__ y, size_t win, int reps, double* out, int bypass_l1)
{
	double acc = 0;
	size_t off = ((size_t)blockIdx.x * 7919u * 256u) % win;
	for (int r = 0; r < reps; r++) {
		for (size_t i = threadIdx.x; i < win; i += blockDim.x * 4) {
			size_t a = (off + i) % win, b = (off + i + blockDim.x) % win, c = (off + i + 2 * blockDim.x) % win,
			       d = (off + i + 3 * blockDim.x) % win;
			if (bypass_l1) {
				acc += __ldcg(y + a) + __ldcg(y + b) + __ldcg(y + c) + __ldcg(y + d);
			} else {
				acc += y[a] + y[b] + y[c] + y[d];
			}
		}
	}
	if (acc == 1.2345) out[0] = acc;
}

// shared-memory gather: each thread does `iters` dependent-free gathers from a 12870-double table
template <int VEC>
__global__ void k_smem_gather(const uint32_t* __restrict__ idx, int nidx, int iters, int tablen, double* out)
{
	extern __shared__ double tab[];
	for (int i = threadIdx.x; i < tablen * VEC; i += blockDim.x) tab[i] = i;
	__syncthreads();
	double acc = 0, acc2 = 0;
	int base = (blockIdx.x * blockDim.x + threadIdx.x) % nidx;
	for (int it = 0; it < iters; it++) {
		uint32_t j = idx[(base + it * 1024) % nidx];
		if (VEC == 1) acc += tab[j];
		else {
			double2 v = reinterpret_cast<double2*>(tab)[j];
			acc += v.x;
			acc2 += v.y;
		}
	}
	if (acc + acc2 == 1.2345) out[0] = acc;
}

int main()
{
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	float ms;
	double* out;
	CK(cudaMalloc(&out, 64));
	int sms = 0;
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	printf("SMs %d\n", sms);
	{   // HBM triad: 24 B per element
		size_t n = (size_t)1 << 28;  // 2 GiB per vector
		double *x, *y;
		CK(cudaMalloc(&x, n * 8));
		CK(cudaMalloc(&y, n * 8));
		CK(cudaMemset(x, 0, n * 8));
		CK(cudaMemset(y, 0, n * 8));
		for (int rep = 0; rep < 3; rep++) {
			CK(cudaEventRecord(e0));
			k_stream_triad<<<sms * 16, 512>>>(x, y, n);
			CK(cudaEventRecord(e1));
			CK(cudaEventSynchronize(e1));
			CK(cudaEventElapsedTime(&ms, e0, e1));
			printf("hbm triad (24 B/elem): %.3f ms  %.1f GB/s\n", ms, 24.0 * n / ms / 1e6);
		}
		cudaFree(x);
		// L2-resident reads
		for (size_t mb : {8, 32, 64, 100}) {
			size_t win = mb * 1024 * 1024 / 8;
			for (int bypass = 0; bypass < 2; bypass++) {
				int reps = (int)(2048 / mb);
				k_l2_read<<<sms * 4, 512>>>(y, win, 1, out, bypass);
				CK(cudaEventRecord(e0));
				k_l2_read<<<sms * 4, 512>>>(y, win, reps, out, bypass);
				CK(cudaEventRecord(e1));
				CK(cudaEventSynchronize(e1));
				CK(cudaEventElapsedTime(&ms, e0, e1));
				double bytes = (double)sms * 4 * reps * win * 8;
				printf("L2 read window %3zu MB %s: %.3f ms  %.1f GB/s\n", mb, bypass ? "ld.cg" : "ld   ", ms, bytes / ms / 1e6);
			}
		}
		cudaFree(y);
	}
	{   // shared-memory gathers
		const int tablen = 12870, nidx = 1 << 20;
		uint32_t* h = new uint32_t[nidx];
		uint32_t* d;
		CK(cudaMalloc(&d, nidx * 4));
		for (int mode = 0; mode < 2; mode++) {
			uint64_t s = 88172645463325252ull;
			for (int i = 0; i < nidx; i++) {
				s ^= s << 13; s ^= s >> 7; s ^= s << 17;
				h[i] = mode == 0 ? (uint32_t)(s % tablen) : (uint32_t)((i * 1 + (i / 1024) * 37) % tablen);  // random | contiguous per warp
			}
			CK(cudaMemcpy(d, h, nidx * 4, cudaMemcpyHostToDevice));
			const int iters = 2000;
			CK(cudaFuncSetAttribute(k_smem_gather<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tablen * 8));
			CK(cudaFuncSetAttribute(k_smem_gather<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tablen * 16));
			k_smem_gather<1><<<sms, 1024, tablen * 8>>>(d, nidx, 10, tablen, out);
			CK(cudaEventRecord(e0));
			k_smem_gather<1><<<sms, 1024, tablen * 8>>>(d, nidx, iters, tablen, out);
			CK(cudaEventRecord(e1));
			CK(cudaEventSynchronize(e1));
			CK(cudaEventElapsedTime(&ms, e0, e1));
			double g = (double)sms * 1024 * iters;
			printf("smem gather  8B %s: %.3f ms  %.2f Ggather/s  %.1f GB/s\n", mode ? "contig" : "random", ms, g / ms / 1e6, g * 8 / ms / 1e6);
			k_smem_gather<2><<<sms, 1024, tablen * 16>>>(d, nidx, 10, tablen, out);
			CK(cudaEventRecord(e0));
			k_smem_gather<2><<<sms, 1024, tablen * 16>>>(d, nidx, iters, tablen, out);
			CK(cudaEventRecord(e1));
			CK(cudaEventSynchronize(e1));
			CK(cudaEventElapsedTime(&ms, e0, e1));
			printf("smem gather 16B %s: %.3f ms  %.2f Ggather/s  %.1f GB/s\n", mode ? "contig" : "random", ms, g / ms / 1e6, g * 16 / ms / 1e6);
		}
	}
	return 0;
}
