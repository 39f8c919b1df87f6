// oracle/psimag_shim/PsimagLite.h -- test infrastructure (see Vector.h)
#ifndef LPP_SHIM_PSIMAGLITE_H
#define LPP_SHIM_PSIMAGLITE_H
#include "Vector.h"
#include "TypeToString.h"
#include "Matrix.h"
#endif
