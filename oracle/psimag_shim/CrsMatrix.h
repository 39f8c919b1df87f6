// oracle/psimag_shim/CrsMatrix.h -- test infrastructure (see Vector.h).
// Compressed-row matrix with the builder interface the reference drives (resize / setRow / pushCol / pushValue) and the
// accumulating product x += A y.
#ifndef LPP_SHIM_CRS_H
#define LPP_SHIM_CRS_H
#include "Vector.h"
#include "Matrix.h"
namespace PsimagLite {
template <typename T>
class CrsMatrix {
public:
	typedef T value_type;
	CrsMatrix() : nrow_(0), ncol_(0) {}
	void resize(SizeType nrow, SizeType ncol)
	{
		nrow_ = nrow; ncol_ = ncol;
		rowptr_.assign(nrow + 1, 0);
		colind_.clear();
		values_.clear();
	}
	void clear() { resize(0, 0); }
	void setRow(SizeType i, SizeType n) { assert(i < rowptr_.size()); rowptr_[i] = n; }
	void pushCol(SizeType c) { colind_.push_back(c); }
	void pushValue(const T& v) { values_.push_back(v); }
	SizeType rows() const { return nrow_; }
	SizeType cols() const { return ncol_; }
	SizeType nonZeros() const { return colind_.size(); }
	SizeType getRowPtr(SizeType i) const { return rowptr_[i]; }
	SizeType getCol(SizeType k) const { return colind_[k]; }
	const T& getValue(SizeType k) const { return values_[k]; }
	void checkValidity() const
	{
		if (rowptr_.size() != nrow_ + 1 || rowptr_[nrow_] != colind_.size() || colind_.size() != values_.size())
			throw RuntimeError("CrsMatrix::checkValidity\n");
	}
	template <typename V> void matrixVectorProduct(V& x, const V& y) const
	{
		for (SizeType i = 0; i < nrow_; ++i)
			for (SizeType k = rowptr_[i]; k < rowptr_[i + 1]; ++k) x[i] += values_[k] * y[colind_[k]];
	}
	Matrix<T> toDense() const
	{
		Matrix<T> m(nrow_, ncol_);
		for (SizeType i = 0; i < nrow_; ++i)
			for (SizeType k = rowptr_[i]; k < rowptr_[i + 1]; ++k) m(i, colind_[k]) += values_[k];
		return m;
	}
private:
	SizeType nrow_, ncol_;
	std::vector<SizeType> rowptr_, colind_;
	std::vector<T> values_;
};
} // namespace PsimagLite
#endif
