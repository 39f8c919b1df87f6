// host/cf_collection.cpp -- the role of PsimagLite's `continuedFractionCollection` driver for the files host/comb_io.h writes:
//   cf_collection -f file.comb[2] -b begin -e end -s step -d delta
// prints one line per frequency: omega  Im G  Re G  (the column order scripts/sqomega.pl:readData expects), G = sum of the
// collection's continued fractions evaluated by lpp_cf_eval (host-side entry point of liblpp_b200.so, no GPU needed).
#include <cstdio>
#include <cstring>
#include <iostream>
#include "../include/lpp_b200.h"
#include "comb_io.h"

int main(int argc, char** argv)
{
	std::string file;
	double wb = 0, we = 0, ws = 0, delta = 0.1;
	for (int i = 1; i + 1 < argc; i += 2) {
		if (!strcmp(argv[i], "-f")) file = argv[i + 1];
		else if (!strcmp(argv[i], "-b")) wb = atof(argv[i + 1]);
		else if (!strcmp(argv[i], "-e")) we = atof(argv[i + 1]);
		else if (!strcmp(argv[i], "-s")) ws = atof(argv[i + 1]);
		else if (!strcmp(argv[i], "-d")) delta = atof(argv[i + 1]);
	}
	if (file.empty() || ws <= 0) {
		std::cerr << "USAGE: " << argv[0] << " -f file -b omegaBegin -e omegaEnd -s omegaStep -d delta\n";
		return 1;
	}
	try {
		const lppb200::CombFile c = lppb200::readComb(file);
		std::vector<double> omega;
		for (double w = wb; w < we + 0.5 * ws; w += ws) omega.push_back(w);
		std::vector<double> re(omega.size(), 0.0), im(omega.size(), 0.0), out(2 * omega.size());
		for (const lppb200::CombFraction& f : c.cfs) {
			if (lpp_cf_eval((int32_t)f.a.size(), f.a.data(), f.b.data(), f.Eg, f.weight, f.isign, (int32_t)omega.size(), omega.data(), delta,
			                out.data()) != 0)
				throw std::runtime_error(lpp_last_error());
			for (size_t k = 0; k < omega.size(); k++) { re[k] += out[2 * k]; im[k] += out[2 * k + 1]; }
		}
		std::cout.precision(15);
		for (size_t k = 0; k < omega.size(); k++) std::cout << omega[k] << " " << im[k] << " " << re[k] << "\n";
	} catch (std::exception& e) {
		std::cerr << "cf_collection: " << e.what() << "\n";
		return 2;
	}
	return 0;
}
