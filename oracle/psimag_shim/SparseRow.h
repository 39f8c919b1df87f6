// oracle/psimag_shim/SparseRow.h -- test infrastructure (see Vector.h).
// One matrix row under construction: add(col, value) in any order; finalize(matrix) sorts by
// column, sums duplicates into one entry, keeps explicit zeros, appends to the CRS arrays and returns the entry count;
// finalize(y) returns sum_k value_k * y[col_k].
#ifndef LPP_SHIM_SPARSEROW_H
#define LPP_SHIM_SPARSEROW_H
#include "Vector.h"
namespace PsimagLite {
template <typename CrsMatrixType>
class SparseRow {
public:
	typedef typename CrsMatrixType::value_type ValueType;
	void add(SizeType col, const ValueType& value)
	{
		cols_.push_back(col);
		values_.push_back(value);
	}
	SizeType finalize(CrsMatrixType& matrix)
	{
		std::vector<SizeType> perm(cols_.size());
		for (SizeType i = 0; i < perm.size(); ++i) perm[i] = i;
		std::stable_sort(perm.begin(), perm.end(), [this](SizeType a, SizeType b) { return cols_[a] < cols_[b]; });
		SizeType counter = 0;
		for (SizeType i = 0; i < perm.size();) {
			const SizeType col = cols_[perm[i]];
			ValueType sum = values_[perm[i]];
			SizeType j = i + 1;
			for (; j < perm.size() && cols_[perm[j]] == col; ++j) sum += values_[perm[j]];
			matrix.pushCol(col);
			matrix.pushValue(sum);
			counter++;
			i = j;
		}
		return counter;
	}
	template <typename VectorLike> ValueType finalize(const VectorLike& y)
	{
		ValueType sum = 0;
		for (SizeType i = 0; i < cols_.size(); ++i) sum += values_[i] * y[cols_[i]];
		return sum;
	}
	template <typename VectorLike> ValueType matrixVectorProduct(const VectorLike& y) { return finalize(y); }
private:
	std::vector<SizeType> cols_;
	std::vector<ValueType> values_;
};
} // namespace PsimagLite
#endif
