// oracle/psimag_shim/Geometry/GeometryDca.h -- test infrastructure (see ../Vector.h).
// Only FeAsMode=INT_KSPACE (out of scope) calls into this class; the members exist so FeBasedSc.h compiles.
#ifndef LPP_SHIM_GEOMETRYDCA_H
#define LPP_SHIM_GEOMETRYDCA_H
#include "../Vector.h"
namespace PsimagLite {
template <typename RealType, typename GeometryType>
class GeometryDca {
public:
	GeometryDca(const GeometryType&, SizeType) {}
	SizeType kSum(SizeType, SizeType) const { throw RuntimeError("GeometryDca: not available in the shim\n"); }
	SizeType kSustract(SizeType, SizeType) const { throw RuntimeError("GeometryDca: not available in the shim\n"); }
};
} // namespace PsimagLite
#endif
