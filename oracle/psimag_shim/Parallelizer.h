// oracle/psimag_shim/Parallelizer.h -- test infrastructure (see Vector.h).  loopCreate(helper): helper.doTask(task, thread) for
// every task in [0, helper.tasks()), over OpenMP.
#ifndef LPP_SHIM_PARALLELIZER_H
#define LPP_SHIM_PARALLELIZER_H
#include "Concurrency.h"
#ifdef _OPENMP
#include <omp.h>
#endif
namespace PsimagLite {
template <typename HelperType>
class Parallelizer {
public:
	explicit Parallelizer(const CodeSectionParams& c) : nthreads_(c.npthreads ? c.npthreads : 1) {}
	String name() const { return "openmp-shim"; }
	void loopCreate(HelperType& helper)
	{
		const long long n = (long long)helper.tasks();
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads((int)nthreads_)
		for (long long i = 0; i < n; ++i) helper.doTask((SizeType)i, (SizeType)omp_get_thread_num());
#else
		for (long long i = 0; i < n; ++i) helper.doTask((SizeType)i, 0);
#endif
	}
private:
	SizeType nthreads_;
};
} // namespace PsimagLite
#endif
