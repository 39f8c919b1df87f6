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
// diag(m, eigs, 'V'): eigenvalues ascending, eigenvectors in the columns of m.  PsimagLite calls LAPACK; the stand-in is a
// cyclic Jacobi iteration (only DefaultSymmetry::fullDiag reaches it, for matrices up to 4900 rows).
template <typename T> void diag(Matrix<T>& m, std::vector<T>& eigs, char)
{
	const SizeType n = m.n_row();
	Matrix<T> v(n, n);
	for (SizeType i = 0; i < n; ++i) v(i, i) = 1;
	for (int sweep = 0; sweep < 100; ++sweep) {
		T off = 0;
		for (SizeType p = 0; p < n; ++p)
			for (SizeType q = p + 1; q < n; ++q) off += m(p, q) * m(p, q);
		if (off < 1e-26) break;
		for (SizeType p = 0; p < n; ++p)
			for (SizeType q = p + 1; q < n; ++q) {
				if (std::abs(m(p, q)) < 1e-300) continue;
				const T theta = (m(q, q) - m(p, p)) / (2 * m(p, q));
				const T t = (theta >= 0 ? 1 : -1) / (std::abs(theta) + std::sqrt(theta * theta + 1));
				const T c = 1 / std::sqrt(t * t + 1), s = t * c;
				for (SizeType k = 0; k < n; ++k) {
					const T a = m(k, p), b = m(k, q);
					m(k, p) = c * a - s * b;
					m(k, q) = s * a + c * b;
				}
				for (SizeType k = 0; k < n; ++k) {
					const T a = m(p, k), b = m(q, k);
					m(p, k) = c * a - s * b;
					m(q, k) = s * a + c * b;
				}
				for (SizeType k = 0; k < n; ++k) {
					const T a = v(k, p), b = v(k, q);
					v(k, p) = c * a - s * b;
					v(k, q) = s * a + c * b;
				}
			}
	}
	std::vector<SizeType> order(n);
	for (SizeType i = 0; i < n; ++i) order[i] = i;
	std::sort(order.begin(), order.end(), [&m](SizeType a, SizeType b) { return m(a, a) < m(b, b); });
	eigs.resize(n);
	Matrix<T> out(n, n);
	for (SizeType j = 0; j < n; ++j) {
		eigs[j] = m(order[j], order[j]);
		for (SizeType i = 0; i < n; ++i) out(i, j) = v(i, order[j]);
	}
	m = out;
}
} // namespace PsimagLite
#endif
