// host/lanczos_b200.cpp -- minimal stand-alone driver over the C-ABI (include/lpp_b200.h).
//
// Mirrors what `lanczos -f input.inp` does on the hot path (src/lanczos.cpp:99-227 -> LanczosDriver1.h:47-66):
// read the input file, build the geometry's connection matrix, construct the model sector, run the Lanczos ground
// state, print "Energy=".  Only the subset of the PsimagLite InputNg format used by the configs is understood:
//   TotalNumberOfSites= NumberOfTerms= GeometryKind=chain|ladder GeometryOptions=ConstantValues IsPeriodicX= LadderLeg=
//   Connectors ... (one block per geometry direction/term, in file order)   Model=HubbardOneBand|FeAsBasedSc|Heisenberg|Tj1Orbital
//   hubbardU / potentialV / MagneticField / AnisotropyD vectors   Orbitals= FeAsMode=INT_PAPER33
//   TargetElectronsUp= TargetElectronsDown= TargetSzPlusConst= HeisenbergTwiceS=1
//   SolverOptions= (InternalProductCuda | InternalProductStored)  LanczosSteps= LanczosEps= LanczosMinSteps= LanczosOptions=reortho Threads=(ignored)
// Usage: lanczos_b200 -f input.inp [-p precision] [--parse-only] [-g c|cdagger [--omega begin,end,step,delta]] [-c c|cdagger|n]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include "../include/lpp_b200.h"
#include "engine_b200.h"
#include "comb_io.h"

struct Input {
	std::map<std::string, std::string> scalars;               // "Label=" lines
	std::map<std::string, std::vector<std::vector<double>>> vectors;  // "label n v..." blocks, in file order per label
};

static bool is_number(const std::string& s)
{
	char* end = nullptr;
	std::strtod(s.c_str(), &end);
	return end != s.c_str() && *end == 0;
}

static Input parse(const std::string& path)
{
	std::ifstream f(path.c_str());
	if (!f) throw std::runtime_error("cannot open " + path);
	std::vector<std::string> tok;
	std::string line;
	Input in;
	while (std::getline(f, line)) {
		size_t h = line.find('#');
		if (h != std::string::npos) line = line.substr(0, h);
		size_t eq = line.find('=');
		if (eq != std::string::npos) {
			std::string k = line.substr(0, eq), v = line.substr(eq + 1);
			k.erase(0, k.find_first_not_of(" \t"));
			k.erase(k.find_last_not_of(" \t;") + 1);
			v.erase(0, v.find_first_not_of(" \t"));
			v.erase(v.find_last_not_of(" \t\r;") + 1);
			in.scalars[k] = v;
			continue;
		}
		std::istringstream ss(line);
		std::string t;
		while (ss >> t) tok.push_back(t);
	}
	// "label n v1 .. vn"  or matrix form "label r c v..." (Connectors 2 2 ...): a label followed by numbers
	for (size_t i = 0; i < tok.size();) {
		if (is_number(tok[i])) { i++; continue; }
		std::string label = tok[i++];
		std::vector<double> nums;
		while (i < tok.size() && is_number(tok[i])) nums.push_back(std::atof(tok[i++].c_str()));
		if (nums.empty()) continue;
		std::vector<double> vals;
		size_t n = (size_t)nums[0];
		if (nums.size() == n + 1) vals.assign(nums.begin() + 1, nums.end());
		else if (nums.size() >= 2 && nums.size() == (size_t)(nums[0] * nums[1]) + 2) vals.assign(nums.begin() + 2, nums.end());
		else vals.assign(nums.begin() + 1, nums.end());
		in.vectors[label].push_back(vals);
	}
	return in;
}

static int geti(const Input& in, const std::string& k, int def, bool required = false)
{
	auto it = in.scalars.find(k);
	if (it == in.scalars.end()) {
		if (required) throw std::runtime_error("missing " + k + "=");
		return def;
	}
	return std::atoi(it->second.c_str());
}
static double getd(const Input& in, const std::string& k, double def)
{
	auto it = in.scalars.find(k);
	return it == in.scalars.end() ? def : std::atof(it->second.c_str());
}
static std::string gets(const Input& in, const std::string& k, const std::string& def)
{
	auto it = in.scalars.find(k);
	return it == in.scalars.end() ? def : it->second;
}

// connection matrix of term `term`: chain (one Connectors block per term) or ladder (two blocks per term: along x, along y).
// With orbitals > 1 a Connectors block holds an orbitals x orbitals matrix (TestSuite/inputs/input100.inp:17-19).
static std::vector<double> connection_matrix(const Input& in, int nsite, int orbitals, int term)
{
	const int nb = nsite * orbitals;
	std::vector<double> m((size_t)nb * nb, 0.0);
	std::string kind = gets(in, "GeometryKind", "chain");
	bool periodic = geti(in, "IsPeriodicX", 0) != 0;
	auto it = in.vectors.find("Connectors");
	if (it == in.vectors.end()) throw std::runtime_error("missing Connectors");
	const auto& blocks = it->second;
	auto orbmat = [&](const std::vector<double>& v, int a, int b) {
		if ((int)v.size() == orbitals * orbitals) return v[a * orbitals + b];
		return a == b ? v[0] : 0.0;   // ConstantValues: one number
	};
	auto bond = [&](int i, int j, const std::vector<double>& v) {
		for (int a = 0; a < orbitals; a++)
			for (int b = 0; b < orbitals; b++) {
				m[(size_t)(i * orbitals + a) * nb + j * orbitals + b] = orbmat(v, a, b);
				m[(size_t)(j * orbitals + b) * nb + i * orbitals + a] = orbmat(v, a, b);
			}
	};
	if (kind == "chain") {
		if ((int)blocks.size() <= term) throw std::runtime_error("not enough Connectors blocks");
		for (int i = 0; i + 1 < nsite; i++) bond(i, i + 1, blocks[term]);
		if (periodic && nsite > 2) bond(0, nsite - 1, blocks[term]);
	} else if (kind == "ladder") {
		int leg = geti(in, "LadderLeg", 2);
		int lx = nsite / leg;
		if ((int)blocks.size() < 2 * (term + 1)) throw std::runtime_error("ladder needs two Connectors blocks per term");
		for (int x = 0; x < lx; x++)
			for (int y = 0; y < leg; y++) {
				int s = y + leg * x;
				if (x + 1 < lx) bond(s, s + leg, blocks[2 * term]);
				else if (periodic && lx > 2) bond(s, y, blocks[2 * term]);
				if (y + 1 < leg) bond(s, s + 1, blocks[2 * term + 1]);
			}
	} else {
		throw std::runtime_error("GeometryKind " + kind + " is not supported by this driver");
	}
	return m;
}

int main(int argc, char** argv)
{
	std::string file;
	int precision = 12;
	bool parse_only = false;
	std::string gf, omega_spec, cicj;
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "-f") && i + 1 < argc) file = argv[++i];
		else if (!strcmp(argv[i], "-p") && i + 1 < argc) precision = std::atoi(argv[++i]);
		else if (!strcmp(argv[i], "--parse-only")) parse_only = true;
		else if (!strcmp(argv[i], "-g") && i + 1 < argc) gf = argv[++i];                 // lanczos -g c | -g cdagger (LanczosOptions.h)
		else if (!strcmp(argv[i], "--omega") && i + 1 < argc) omega_spec = argv[++i];    // begin,end,step,delta
		else if (!strcmp(argv[i], "-c") && i + 1 < argc) cicj = argv[++i];               // lanczos -c c | -c n: two-point matrix
	}
	if (file.empty()) {
		std::cerr << "USAGE: " << argv[0] << " -f input.inp [-p precision] [--parse-only]\n";
		return 1;
	}
	try {
		Input in = parse(file);
		std::string model = gets(in, "Model", "");
		lpp_desc d;
		memset(&d, 0, sizeof(d));
		d.nsite = geti(in, "TotalNumberOfSites", 0, true);
		d.orbitals = 1;
		std::vector<double> U, V, D, jzz, jpm, w;
		if (in.vectors.count("hubbardU")) U = in.vectors["hubbardU"][0];
		if (in.vectors.count("potentialV")) V = in.vectors["potentialV"][0];
		if (model == "HubbardOneBand") {
			d.model = LPP_MODEL_HUBBARD;
			d.nup = geti(in, "TargetElectronsUp", 0, true);
			d.ndown = geti(in, "TargetElectronsDown", 0, true);
		} else if (model == "FeAsBasedSc") {
			d.model = LPP_MODEL_FEAS;
			d.orbitals = geti(in, "Orbitals", 2, true);
			if (gets(in, "FeAsMode", "INT_PAPER33") != "INT_PAPER33") throw std::runtime_error("only FeAsMode=INT_PAPER33 is on the path");
			d.nup = geti(in, "TargetElectronsUp", 0, true);
			d.ndown = geti(in, "TargetElectronsDown", 0, true);
			D.assign(1, getd(in, "AnisotropyD", 0.0));
		} else if (model == "Heisenberg") {
			d.model = LPP_MODEL_HEISENBERG;
			if (geti(in, "HeisenbergTwiceS", 1) != 1) throw std::runtime_error("only HeisenbergTwiceS=1 is on the path");
			d.nup = geti(in, "TargetSzPlusConst", 0, true);
			if (in.vectors.count("MagneticField")) V = in.vectors["MagneticField"][0];
			if (in.vectors.count("AnisotropyD")) D = in.vectors["AnisotropyD"][0];
			jzz = connection_matrix(in, d.nsite, 1, 1);
		} else if (model == "Tj1Orbital" || model == "TjMultiOrb") {
			// TjMultiOrb.h:52-80: four geometry terms (hopping, S+S-, SzSz, n n); Orbitals=1 only
			d.model = LPP_MODEL_TJ;
			if (geti(in, "Orbitals", 1) != 1) throw std::runtime_error("t-J: only Orbitals=1 is on the path");
			d.nup = geti(in, "TargetElectronsUp", 0, true);
			d.ndown = geti(in, "TargetElectronsDown", 0, true);
			jpm = connection_matrix(in, d.nsite, 1, 1);
			jzz = connection_matrix(in, d.nsite, 1, 2);
			w = connection_matrix(in, d.nsite, 1, 3);
		} else {
			throw std::runtime_error("Model=" + model + " is not on the accelerated path");
		}
		std::vector<double> hop = connection_matrix(in, d.nsite, d.orbitals, 0);
		d.feas_u3_all_pairs = 1;
		d.hop = hop.data();
		d.jzz = jzz.empty() ? nullptr : jzz.data();
		d.jpm = jpm.empty() ? nullptr : jpm.data();
		d.w = w.empty() ? nullptr : w.data();
		d.U = U.empty() ? nullptr : U.data(); d.nU = (int)U.size();
		d.V = V.empty() ? nullptr : V.data(); d.nV = (int)V.size();
		d.D = D.empty() ? nullptr : D.data(); d.nD = (int)D.size();
		d.device = 0; d.rank = 0; d.nranks = 1;
		std::string opts = gets(in, "SolverOptions", "none");
		lpp_solver_params p;
		p.steps = geti(in, "LanczosSteps", 200);
		p.minsteps = geti(in, "LanczosMinSteps", 4);
		p.eps = getd(in, "LanczosEps", 1e-12);
		p.kernel = opts.find("InternalProductStored") != std::string::npos ? LPP_KERNEL_STORED : LPP_KERNEL_AUTO;
		p.reortho = gets(in, "LanczosOptions", "none").find("reortho") != std::string::npos ? 1 : 0;
		p.seed = 1234;
		if (parse_only) {
			std::cout << "model=" << d.model << " nsite=" << d.nsite << " orbitals=" << d.orbitals << " nup=" << d.nup
			          << " ndown=" << d.ndown << " nU=" << d.nU << " nV=" << d.nV << " kernel=" << p.kernel << " hop01=" << hop[1]
			          << " steps=" << p.steps << "\n";
			return 0;
		}
		lpp_solver_params ps = p;                       // ParametersForSolver(io, "Spectral"), Engine.h:472
		ps.steps = geti(in, "SpectralSteps", 200);
		ps.minsteps = geti(in, "SpectralMinSteps", 4);
		ps.eps = getd(in, "SpectralEps", 1e-12);
		ps.reortho = gets(in, "SpectralOptions", "none").find("reortho") != std::string::npos ? 1 : 0;
		lppb200::Engine engine(d, p, ps);
		std::cout.precision(precision);
		std::cout << "#Hilbert=" << engine.rows() << " LanczosSteps=" << engine.lanczosSteps() << "\n";
		std::cout << "Energy=" << engine.energies(0) << "\n";   // LanczosDriver1.h:64-66
		if (!cicj.empty()) {
			// LanczosDriver1.h:183-199 -> Engine::twoPoint (Engine.h:262-331): result(i, j) = <O_j gs | O_i gs>, spin TSPSpin
			if (d.model == LPP_MODEL_HEISENBERG) throw std::runtime_error("-c is available for the fermionic models");
			const int what = cicj == "c" ? LPP_OP_C : (cicj == "cdagger" ? LPP_OP_CDAGGER : (cicj == "n" ? LPP_OP_N : 0));
			if (!what) throw std::runtime_error("-c expects c, cdagger or n");
			if (what == LPP_OP_N && d.model != LPP_MODEL_HUBBARD) throw std::runtime_error("-c n is available for Model=HubbardOneBand");
			const int spin = geti(in, "TSPSpin", 0);
			const std::vector<double> m = engine.twoPoint(what, spin);
			std::cout << "spins=" << spin << " " << spin << "\norbs=0 0\n" << d.nsite << " " << d.nsite << "\n";
			double sum = 0;
			for (int i = 0; i < d.nsite; i++) {
				for (int j = 0; j < d.nsite; j++) std::cout << m[(size_t)i * d.nsite + j] << " ";
				std::cout << "\n";
				sum += m[(size_t)i * d.nsite + i];
			}
			std::cout << "MatrixDiagonal = " << sum << "\n";          // Engine.h:330
		}
		if (!gf.empty()) {
			// LanczosDriver1.h:96-181: TSPSites (one site = diagonal), one continued-fraction collection per pair of sites
			const int what = gf == "c" ? LPP_OP_C : gf == "cdagger" ? LPP_OP_CDAGGER : gf == "sz" ? LPP_OP_SZ : gf == "splus" ? LPP_OP_SPLUS
			               : gf == "sminus" ? LPP_OP_SMINUS : 0;
			if (!what) throw std::runtime_error("-g expects c, cdagger, sz, splus or sminus");
			if (!in.vectors.count("TSPSites") || in.vectors["TSPSites"][0].empty()) throw std::runtime_error("TSPSites must have at least one site");
			std::vector<double> sites = in.vectors["TSPSites"][0];
			if (sites.size() == 1) sites.push_back(sites[0]);
			const int site0 = (int)sites[0], site1 = (int)sites[1];
			const int spin = geti(in, "TSPSpin", 0);
			std::vector<lppb200::ContinuedFraction> cfs;
			engine.spectralFunction(cfs, what, site0, site1, spin);
			std::cout << "#gf(i=" << site0 << ", j=" << site1 << ")\n";
			{   // <basename of the input>0.comb, LanczosDriver1.h:147-181
				lppb200::CombFile comb;
				comb.site0 = site0;
				comb.site1 = site1;
				for (const auto& cf : cfs) {
					std::ostringstream key;
					key << spin << "," << cf.type << ",0,0";                 // Engine.h:199-202
					comb.indexToCf.push_back(key.str());
					lppb200::CombFraction f;
					f.a = cf.a; f.b = cf.b; f.weight = cf.weight; f.Eg = cf.Eg; f.isign = cf.isign;
					comb.cfs.push_back(f);
				}
				std::string base = file.substr(file.find_last_of('/') == std::string::npos ? 0 : file.find_last_of('/') + 1);
				const std::string out = base + "0.comb";
				lppb200::writeComb(out, comb);
				std::cerr << "lanczos_b200: Written to " << out << "\n";
			}
			for (const auto& cf : cfs) {
				std::cout << "#CF type=" << cf.type << " isign=" << cf.isign << " weight=" << cf.weight << " Eg=" << cf.Eg << " steps=" << cf.a.size() << "\n";
				for (size_t k = 0; k < cf.a.size(); k++) std::cout << cf.a[k] << " " << cf.b[k] << "\n";
			}
			if (!omega_spec.empty()) {
				double ob = 0, oe = 0, os = 1, delta = 0.1;
				if (sscanf(omega_spec.c_str(), "%lf,%lf,%lf,%lf", &ob, &oe, &os, &delta) != 4 || os <= 0) throw std::runtime_error("--omega expects begin,end,step,delta");
				std::vector<double> omega;
				for (double w = ob; w < oe + 0.5 * os; w += os) omega.push_back(w);
				std::vector<std::complex<double> > g(omega.size());
				for (const auto& cf : cfs) {
					const std::vector<std::complex<double> > gi = cf(omega, delta);
					for (size_t k = 0; k < omega.size(); k++) g[k] += gi[k];
				}
				std::cout << "#omega ReG ImG\n";       // continuedFractionCollection -b -e -s -d (scripts/sqomega.pl:24-27)
				for (size_t k = 0; k < omega.size(); k++) std::cout << omega[k] << " " << g[k].real() << " " << g[k].imag() << "\n";
			}
		}
	} catch (std::exception& e) {
		std::cerr << "lanczos_b200: " << e.what() << "\n";
		return 2;
	}
	return 0;
}
