// tools/microbench_lds.cu -- the shared-memory gather ceilings the two sweeps are measured against.
//
// Replaces the shared-memory part of tools/microbench.cu, whose loop fetched every index with a global load and an integer
// `%` and therefore measured issue rate, not the shared-memory data stage (round-1 review, item 6).  Here every address comes
// from registers (one multiply-add of a per-thread linear congruential state, one multiply-high to bring it into range), the
// loop is unrolled eight times over independent accumulators, and nothing but LDS touches the memory pipe.
//
// Patterns (all on a 12870-entry table, the length of one row of config 3):
//   seq      lane i reads element base+i          (conflict-free by construction)
//   random   every lane reads its own random element
//   line     the eight lanes of a quarter-warp read the eight 16-byte pieces of one random 128-byte line
//            (the access of k_dblock: LDS.128, one wavefront per quarter-warp)
//            (LDS.64: the sixteen lanes of a half-warp read the sixteen 8-byte pieces of one line)
// for LDS.64 (8 B per lane) and LDS.128 (16 B per lane).  Printed: bytes per clock per SM against the 128 B/clk data stage.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_lds tools/microbench_lds.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

enum { SEQ = 0, RANDOM = 1, LINE = 2 };

template <int VEC, int MODE>
__global__ void __launch_bounds__(1024, 1) k_lds(int iters, uint32_t nelem, double* out, long long* cycles)
{
	extern __shared__ __align__(16) double tab[];
	for (uint32_t i = threadIdx.x; i < nelem * VEC; i += blockDim.x) tab[i] = (double)i;
	__syncthreads();
	const uint32_t lane = threadIdx.x & 31u;
	// LINE: the quarter-warp (LDS.128) or half-warp (LDS.64) shares one generator; RANDOM: one per lane; SEQ: one per warp
	uint32_t s = MODE == RANDOM ? threadIdx.x * 2654435761u + blockIdx.x : MODE == LINE ? (threadIdx.x >> (VEC == 2 ? 3 : 4)) * 2654435761u + blockIdx.x
	                                                                                     : (threadIdx.x >> 5) * 2654435761u + blockIdx.x;
	const uint32_t span = MODE == LINE ? (VEC == 2 ? nelem / 8 : nelem / 16) : MODE == SEQ ? nelem - 32 : nelem;   // units the generator picks from
	const uint32_t esz = VEC * 8;
	const uint32_t unit = MODE == LINE ? 128u : esz;                                // bytes per generated unit
	const uint32_t self = MODE == LINE ? (VEC == 2 ? (lane & 7u) * 16u : (lane & 15u) * 8u) : MODE == SEQ ? lane * esz : 0u;
	const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab) + self;
	double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0;
	const long long t0 = clock64();
	for (int it = 0; it < iters; it++) {
#pragma unroll
		for (int u = 0; u < 8; u++) {
			s = s * 1664525u + 1013904223u;
			const uint32_t addr = base + __umulhi(s, span) * unit;
			if (VEC == 1) {
				double v;
				asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
				if (u == 0) a0 += v; else if (u == 1) a1 += v; else if (u == 2) a2 += v; else if (u == 3) a3 += v;
				else if (u == 4) a4 += v; else if (u == 5) a5 += v; else if (u == 6) a6 += v; else a7 += v;
			} else {
				double v, w;
				asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v), "=d"(w) : "r"(addr));
				if (u == 0) a0 += v; else if (u == 1) a1 += v; else if (u == 2) a2 += v; else if (u == 3) a3 += v;
				else if (u == 4) a4 += w; else if (u == 5) a5 += w; else if (u == 6) a6 += w; else a7 += w;
				if (u < 4) a7 += w; else a0 += v;
			}
		}
	}
	const long long t1 = clock64();
	const double acc = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
	if (acc == 1.2345) out[0] = acc;
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int VEC, int MODE>
static int run(const char* name, int sms, double* out, long long* cyc)
{
	const uint32_t nelem = 12870;
	const size_t smem = (size_t)nelem * VEC * 8;
	const int iters = 4000;
	CK(cudaFuncSetAttribute(k_lds<VEC, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	k_lds<VEC, MODE><<<sms, 1024, smem>>>(10, nelem, out, cyc);
	CK(cudaEventRecord(e0));
	k_lds<VEC, MODE><<<sms, 1024, smem>>>(iters, nelem, out, cyc);
	CK(cudaEventRecord(e1));
	CK(cudaEventSynchronize(e1));
	float ms;
	CK(cudaEventElapsedTime(&ms, e0, e1));
	long long h[256];
	CK(cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
	double mean = 0;
	for (int i = 0; i < sms; i++) mean += (double)h[i] / sms;
	const double bytes_per_sm = 1024.0 * iters * 8 * VEC * 8;
	printf("LDS.%-3d %-6s: %.3f ms  %6.1f B/clk/SM (loop clocks)  %7.1f GB/s chip (events)\n", VEC * 64, name, ms, bytes_per_sm / mean,
	       bytes_per_sm * sms / ms / 1e6);
	return 0;
}

int main()
{
	int sms = 0;
	CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
	double* out;
	long long* cyc;
	CK(cudaMalloc(&out, 64));
	CK(cudaMalloc(&cyc, sizeof(long long) * 256));
	printf("SMs %d; data stage 128 B/clk/SM\n", sms);
	if (run<1, SEQ>("seq", sms, out, cyc)) return 1;
	if (run<1, RANDOM>("random", sms, out, cyc)) return 1;
	if (run<1, LINE>("line", sms, out, cyc)) return 1;
	if (run<2, SEQ>("seq", sms, out, cyc)) return 1;
	if (run<2, RANDOM>("random", sms, out, cyc)) return 1;
	if (run<2, LINE>("line", sms, out, cyc)) return 1;
	return 0;
}
