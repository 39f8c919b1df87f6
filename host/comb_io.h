// host/comb_io.h -- the `.comb` file the reference writes after `lanczos -g` (LanczosDriver1.h:147-181) and the evaluator side of
// PsimagLite's `continuedFractionCollection -f file -b begin -e end -s step -d delta` (scripts/sqomega.pl:24-27).
// PsimagLite's IoSimple / ContinuedFraction::write are not in the reference repository, so the layout below is the one its
// in-repo consumers rely on (scripts/extractOrbitals.pl: a `#INDEXTOCF spin,type,orb1,orb2 ...` line, a
// `#CONTINUEDFRACTIONCOLLECTION=<n>` line, then one block per continued fraction that BEGINS with `#Avector`); the remaining
// labels of a block (#Bvector, #CFWeight=, #CFEnergy=, #CFIsign=) are this repository's choice.  The reference's
// extractOrbitals.pl runs unchanged on these files (tests/test_comb.py).
#ifndef LPP_COMB_IO_H
#define LPP_COMB_IO_H

#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace lppb200 {

struct CombFraction {
	std::vector<double> a, b;
	double weight = 0, Eg = 0;
	int isign = 1;
};

struct CombFile {
	int site0 = 0, site1 = 0;
	std::vector<std::string> indexToCf;      // "spin,type,orb1,orb2" per fraction (Engine.h:199-202)
	std::vector<CombFraction> cfs;
};

inline void writeComb(const std::string& path, const CombFile& c)
{
	std::ofstream os(path);
	if (!os) throw std::runtime_error("cannot write " + path);
	os << std::setprecision(17);
	os << "#Site0=" << c.site0 << "\n#Site1=" << c.site1 << "\n#INDEXTOCF ";
	for (const std::string& s : c.indexToCf) os << s << " ";
	os << "\n#CONTINUEDFRACTIONCOLLECTION=" << c.cfs.size() << "\n";
	for (const CombFraction& f : c.cfs) {
		os << "#Avector\n" << f.a.size() << "\n";
		for (double v : f.a) os << v << "\n";
		os << "#Bvector\n" << f.b.size() << "\n";
		for (double v : f.b) os << v << "\n";
		os << "#CFWeight=" << f.weight << "\n#CFEnergy=" << f.Eg << "\n#CFIsign=" << f.isign << "\n";
	}
}

inline CombFile readComb(const std::string& path)
{
	std::ifstream is(path);
	if (!is) throw std::runtime_error("cannot open " + path);
	CombFile c;
	std::string line;
	size_t expected = 0;
	bool haveCount = false;
	auto readVector = [&is](std::vector<double>& v) {
		size_t n = 0;
		is >> n;
		v.resize(n);
		for (size_t i = 0; i < n; i++) is >> v[i];
		std::string rest;
		std::getline(is, rest);
	};
	while (std::getline(is, line)) {
		if (line.rfind("#Site0=", 0) == 0) c.site0 = std::stoi(line.substr(7));
		else if (line.rfind("#Site1=", 0) == 0) c.site1 = std::stoi(line.substr(7));
		else if (line.rfind("#INDEXTOCF", 0) == 0) {
			std::istringstream ss(line.substr(10));
			std::string t;
			while (ss >> t) c.indexToCf.push_back(t);
		} else if (line.rfind("#CONTINUEDFRACTIONCOLLECTION=", 0) == 0) {
			expected = std::stoul(line.substr(29));
			haveCount = true;
		} else if (line.rfind("#Avector", 0) == 0) {
			c.cfs.push_back(CombFraction());
			readVector(c.cfs.back().a);
		} else if (line.rfind("#Bvector", 0) == 0 && !c.cfs.empty()) readVector(c.cfs.back().b);
		else if (line.rfind("#CFWeight=", 0) == 0 && !c.cfs.empty()) c.cfs.back().weight = std::stod(line.substr(10));
		else if (line.rfind("#CFEnergy=", 0) == 0 && !c.cfs.empty()) c.cfs.back().Eg = std::stod(line.substr(10));
		else if (line.rfind("#CFIsign=", 0) == 0 && !c.cfs.empty()) c.cfs.back().isign = std::stoi(line.substr(9));
	}
	if (!haveCount) throw std::runtime_error(path + ": no #CONTINUEDFRACTIONCOLLECTION= line");
	if (c.cfs.size() != expected) throw std::runtime_error(path + ": fraction count does not match #CONTINUEDFRACTIONCOLLECTION=");
	return c;
}

} // namespace lppb200
#endif
