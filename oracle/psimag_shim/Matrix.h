// oracle/psimag_shim/Matrix.h -- test infrastructure (see Vector.h).  Dense matrix with the members the path uses.
#ifndef LPP_SHIM_MATRIX_H
#define LPP_SHIM_MATRIX_H
#include "Vector.h"
namespace PsimagLite {
template <typename T>
class Matrix {
public:
	typedef T value_type;
	Matrix() : nrow_(0), ncol_(0) {}
	Matrix(SizeType nrow, SizeType ncol) : nrow_(nrow), ncol_(ncol), data_(nrow * ncol, T()) {}
	void resize(SizeType nrow, SizeType ncol) { nrow_ = nrow; ncol_ = ncol; data_.assign(nrow * ncol, T()); }
	void resize(SizeType nrow, SizeType ncol, const T& v) { nrow_ = nrow; ncol_ = ncol; data_.assign(nrow * ncol, v); }
	void setTo(const T& v) { std::fill(data_.begin(), data_.end(), v); }
	void clear() { nrow_ = ncol_ = 0; data_.clear(); }
	SizeType n_row() const { return nrow_; }
	SizeType n_col() const { return ncol_; }
	SizeType rows() const { return nrow_; }
	SizeType cols() const { return ncol_; }
	const T& operator()(SizeType i, SizeType j) const { assert(i < nrow_ && j < ncol_); return data_[i + j * nrow_]; }
	T& operator()(SizeType i, SizeType j) { assert(i < nrow_ && j < ncol_); return data_[i + j * nrow_]; }
private:
	SizeType nrow_, ncol_;
	std::vector<T> data_;
};
template <typename T> std::ostream& operator<<(std::ostream& os, const Matrix<T>& m)
{
	os << m.n_row() << " " << m.n_col() << "\n";
	for (SizeType i = 0; i < m.n_row(); ++i) {
		for (SizeType j = 0; j < m.n_col(); ++j) os << m(i, j) << " ";
		os << "\n";
	}
	return os;
}
} // namespace PsimagLite
#endif
