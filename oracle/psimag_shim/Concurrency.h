// oracle/psimag_shim/Concurrency.h -- test infrastructure (see Vector.h)
#ifndef LPP_SHIM_CONCURRENCY_H
#define LPP_SHIM_CONCURRENCY_H
#include "Vector.h"
namespace PsimagLite {
struct CodeSectionParams {
	CodeSectionParams(SizeType n = 1) : npthreads(n) {}
	SizeType npthreads;
};
class Concurrency {
public:
	static inline CodeSectionParams codeSectionParams{1};
	static void setOptions(const CodeSectionParams& c) { codeSectionParams = c; }
};
} // namespace PsimagLite
#endif
