// oracle/psimag_shim/TypeToString.h -- test infrastructure (see Vector.h)
#ifndef LPP_SHIM_TTOS_H
#define LPP_SHIM_TTOS_H
#include "Vector.h"
template <typename T> PsimagLite::String ttos(const T& t)
{
	std::ostringstream ss;
	ss << t;
	return ss.str();
}
#endif
