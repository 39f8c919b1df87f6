// oracle/psimag_shim/Parallelizer2.h -- test infrastructure (see Vector.h).  parallelFor(begin, end, f(i, thread)) over OpenMP.
#ifndef LPP_SHIM_PARALLELIZER2_H
#define LPP_SHIM_PARALLELIZER2_H
#include "Concurrency.h"
#ifdef _OPENMP
#include <omp.h>
#endif
namespace PsimagLite {
template <typename Unused = int>
class Parallelizer2 {
public:
	explicit Parallelizer2(const CodeSectionParams& c) : nthreads_(c.npthreads ? c.npthreads : 1) {}
	template <typename F> void parallelFor(SizeType begin, SizeType end, const F& f)
	{
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads((int)nthreads_)
		for (long long i = (long long)begin; i < (long long)end; ++i) f((SizeType)i, (SizeType)omp_get_thread_num());
#else
		for (SizeType i = begin; i < end; ++i) f(i, 0);
#endif
	}
private:
	SizeType nthreads_;
};
} // namespace PsimagLite
#endif
