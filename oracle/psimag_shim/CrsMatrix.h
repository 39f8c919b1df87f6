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
	CrsMatrix(SizeType nrow, SizeType ncol) { resize(nrow, ncol); }
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
// isHermitian(A [, verbose]) as the models assert it after setupHamiltonian
template <typename T> bool isHermitian(const CrsMatrix<T>& a, bool = false)
{
	if (a.rows() != a.cols()) return false;
	const Matrix<T> d = a.toDense();
	for (SizeType i = 0; i < d.n_row(); ++i)
		for (SizeType j = i + 1; j < d.n_col(); ++j)
			if (std::abs(d(i, j) - conj(d(j, i))) > 1e-12) return false;
	return true;
}

// C = A^dagger and C = A B: only the JHundInfinity=1 branch of TjMultiOrb.h (two orbitals, out of scope) and
// ProgramGlobals::transform reach these; they exist so that the reference headers compile
template <typename T> void transposeConjugate(CrsMatrix<T>& c, const CrsMatrix<T>& a)
{
	std::vector<std::vector<std::pair<SizeType, T> > > rows(a.cols());
	for (SizeType i = 0; i < a.rows(); ++i)
		for (SizeType k = a.getRowPtr(i); k < a.getRowPtr(i + 1); ++k) rows[a.getCol(k)].push_back(std::make_pair(i, conj(a.getValue(k))));
	c.resize(a.cols(), a.rows());
	SizeType counter = 0;
	for (SizeType i = 0; i < rows.size(); ++i) {
		c.setRow(i, counter);
		for (SizeType k = 0; k < rows[i].size(); ++k) { c.pushCol(rows[i][k].first); c.pushValue(rows[i][k].second); counter++; }
	}
	c.setRow(rows.size(), counter);
}
template <typename T> void multiply(CrsMatrix<T>& c, const CrsMatrix<T>& a, const CrsMatrix<T>& b)
{
	c.resize(a.rows(), b.cols());
	SizeType counter = 0;
	std::vector<T> acc(b.cols());
	std::vector<char> used(b.cols());
	for (SizeType i = 0; i < a.rows(); ++i) {
		c.setRow(i, counter);
		std::fill(acc.begin(), acc.end(), T(0));
		std::fill(used.begin(), used.end(), 0);
		for (SizeType k = a.getRowPtr(i); k < a.getRowPtr(i + 1); ++k)
			for (SizeType l = b.getRowPtr(a.getCol(k)); l < b.getRowPtr(a.getCol(k) + 1); ++l) {
				acc[b.getCol(l)] += a.getValue(k) * b.getValue(l);
				used[b.getCol(l)] = 1;
			}
		for (SizeType j = 0; j < b.cols(); ++j)
			if (used[j]) { c.pushCol(j); c.pushValue(acc[j]); counter++; }
	}
	c.setRow(a.rows(), counter);
}
template <typename V, typename T> void multiply(V& x, const CrsMatrix<T>& a, const V& y)
{
	for (SizeType i = 0; i < a.rows(); ++i) {
		x[i] = 0;
		for (SizeType k = a.getRowPtr(i); k < a.getRowPtr(i + 1); ++k) x[i] += a.getValue(k) * y[a.getCol(k)];
	}
}
} // namespace PsimagLite
#endif
