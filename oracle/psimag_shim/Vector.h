// oracle/psimag_shim/Vector.h -- TEST INFRASTRUCTURE, not product code.
// Minimal stand-in for the parts of PsimagLite (github.com/g1257/PsimagLite, un-vendored and absent from /root/reference)
// that the reference's hot-path headers use, so that those headers compile unmodified from where they lie
// (oracle/Makefile, target _ref).  Written from the usage visible in the reference; nothing is copied.
#ifndef LPP_SHIM_VECTOR_H
#define LPP_SHIM_VECTOR_H
#include <vector>
#include <string>
#include <complex>
#include <stdexcept>
#include <iostream>
#include <sstream>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <algorithm>

typedef std::size_t SizeType;       // PsimagLite puts SizeType in the global namespace

namespace PsimagLite {

typedef std::string String;

class RuntimeError : public std::runtime_error {
public:
	explicit RuntimeError(const String& s) : std::runtime_error(s) {}
};

template <typename T> struct Vector { typedef std::vector<T> Type; };

template <typename T> struct IsVectorLike { enum { True = false }; };
template <typename T> struct IsVectorLike<std::vector<T> > { enum { True = true }; };

template <typename A, typename B> struct IsSame { enum { True = false }; };
template <typename A> struct IsSame<A, A> { enum { True = true }; };

template <bool B, typename T> struct EnableIf {};
template <typename T> struct EnableIf<true, T> { typedef T Type; };

template <typename T> struct Real { typedef T Type; };
template <typename T> struct Real<std::complex<T> > { typedef T Type; };

inline double real(double x) { return x; }
inline double imag(double) { return 0.0; }
inline double conj(double x) { return x; }
inline double real(const std::complex<double>& x) { return x.real(); }
inline double imag(const std::complex<double>& x) { return x.imag(); }
inline std::complex<double> conj(const std::complex<double>& x) { return std::conj(x); }

// Sort<V>::sort(v, permutation): ascending sort that also returns where every element came from (used by the
// JHundInfinity branch of TjMultiOrb.h only)
template <typename V>
class Sort {
public:
	void sort(V& v, std::vector<SizeType>& perm)
	{
		perm.resize(v.size());
		for (SizeType i = 0; i < perm.size(); ++i) perm[i] = i;
		std::stable_sort(perm.begin(), perm.end(), [&v](SizeType a, SizeType b) { return v[a] < v[b]; });
		V w(v.size());
		for (SizeType i = 0; i < perm.size(); ++i) w[i] = v[perm[i]];
		v.swap(w);
	}
};

template <typename V> void vectorPrint(const V& v, const char* label, std::ostream& os)
{
	os << label << " " << v.size() << "\n";
	for (SizeType i = 0; i < v.size(); ++i) os << v[i] << " ";
	os << "\n";
}

} // namespace PsimagLite

namespace std {
// PsimagLite prints vectors with operator<< (used by the Parameters*.h printers); it has to live where argument-dependent
// lookup on std::vector finds it
template <typename T, typename A> ostream& operator<<(ostream& os, const vector<T, A>& v)
{
	os << v.size() << "\n";
	for (size_t i = 0; i < v.size(); ++i) os << v[i] << " ";
	return os << "\n";
}
} // namespace std

inline void err(const PsimagLite::String& s) { throw PsimagLite::RuntimeError(s); }

#endif
