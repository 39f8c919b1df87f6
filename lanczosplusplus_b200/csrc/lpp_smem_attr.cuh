// lpp_smem_attr.cuh -- cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the FUNCTION (per device), not of a launch:
// a plan that sets it to its own need lowers it for every earlier plan that runs the same kernel with a larger tile (the ground
// state sector and the N-1 sector of a continued fraction live side by side).  lpp_raise_smem only ever raises the limit.
#pragma once
#include <cstddef>
#include <map>
#include <mutex>
#include <utility>
#include <cuda_runtime.h>

static inline cudaError_t lpp_raise_smem_ptr(const void* func, size_t bytes)
{
	static std::mutex mu;
	static std::map<std::pair<int, const void*>, size_t> limit;       // (device, kernel) -> the value last set
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	std::lock_guard<std::mutex> guard(mu);
	size_t& cur = limit[std::make_pair(dev, func)];
	if (bytes <= cur) return cudaSuccess;
	e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
	if (e == cudaSuccess) cur = bytes;
	return e;
}
template <class K>
static inline cudaError_t lpp_raise_smem(K* kernel, size_t bytes)
{
	return lpp_raise_smem_ptr(reinterpret_cast<const void*>(kernel), bytes);
}
