// oracle/psimag_shim/ProgressIndicator.h -- test infrastructure (see Vector.h).  DefaultSymmetry.h includes it and uses nothing of it.
#ifndef LPP_SHIM_PROGRESS_H
#define LPP_SHIM_PROGRESS_H
#include "Vector.h"
namespace PsimagLite {
class ProgressIndicator {
public:
	explicit ProgressIndicator(const String&) {}
};
} // namespace PsimagLite
#endif
