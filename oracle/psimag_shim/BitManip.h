// oracle/psimag_shim/BitManip.h -- test infrastructure (see Vector.h).  Only count() is used on the path.
#ifndef LPP_SHIM_BITMANIP_H
#define LPP_SHIM_BITMANIP_H
#include "Vector.h"
namespace PsimagLite {
namespace BitManip {
inline int count(unsigned long w)
{
	int c = 0;
	for (; w; w &= w - 1) c++;
	return c;
}
} // namespace BitManip
} // namespace PsimagLite
#endif
