"""The stand-alone C++ driver (host/lanczos_b200.cpp): input parsing on the CPU, energies on the GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(lpp, args, cwd=None):
    exe = lpp._lib.build_driver()
    return subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=cwd)


def test_parses_reference_style_inputs(lpp):
    r = run(lpp, ["-f", os.path.join(ROOT, "tests/inputs/input0.inp"), "--parse-only"])
    assert r.returncode == 0 and "model=0 nsite=4 orbitals=1 nup=2 ndown=2 nU=4 nV=8 kernel=0 hop01=-1" in r.stdout
    r = run(lpp, ["-f", os.path.join(ROOT, "tests/inputs/c1_hubbard8.inp"), "--parse-only"])
    assert r.returncode == 0 and "nsite=8" in r.stdout and "kernel=4" in r.stdout      # InternalProductStored
    r = run(lpp, ["-f", os.path.join(ROOT, "tests/inputs/feas6.inp"), "--parse-only"])
    assert r.returncode == 0 and "model=1 nsite=6 orbitals=2 nup=3 ndown=3 nU=4 nV=24" in r.stdout
    r = run(lpp, ["-f", os.path.join(ROOT, "tests/inputs/tj8.inp"), "--parse-only"])
    assert r.returncode == 0 and "model=3 nsite=8 orbitals=1 nup=3 ndown=3" in r.stdout and "kernel=4" in r.stdout
    r = run(lpp, ["-f", "/nonexistent.inp"])
    assert r.returncode == 2 and "cannot open" in r.stderr


@pytest.mark.gpu
def test_driver_energies(lpp):
    for name, ref in (("input0.inp", -2 * np.sqrt(5.0)), ("c1_hubbard8.inp", -4.235806999130),
                      ("feas6.inp", -3.099464014219), ("tj8.inp", -4.430663564423)):   # t-J: dense eigvalsh of the oracle's CRS
        r = run(lpp, ["-f", os.path.join(ROOT, "tests/inputs", name), "-p", "14"])
        assert r.returncode == 0, r.stderr
        e = float(re.search(r"Energy=(\S+)", r.stdout).group(1))
        assert abs(e - ref) < 1e-9, (name, e)


@pytest.mark.gpu
def test_driver_spectral_function_matches_python_engine(lpp, tmp_path):
    """`lanczos_b200 -g c`: the C++ Engine mirror (host/engine_b200.h, Engine.h:133-206) against the ctypes mirror and its
    oracle-checked continued fractions (tests/test_gpu_parity.py::test_continued_fraction_parity)."""
    from tests import cases
    r = run(lpp, ["-f", os.path.join(ROOT, "tests/inputs/hubbard6_gf.inp"), "-p", "15", "-g", "c", "--omega", "-4,4,0.5,0.1"],
            cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert "#gf(i=1, j=3)" in r.stdout
    # the .comb file of LanczosDriver1.h:147-181 next to the run, evaluated by the continuedFractionCollection stand-in
    comb_path = str(tmp_path / "hubbard6_gf.inp0.comb")
    assert os.path.exists(comb_path)
    ev = subprocess.run([lpp._lib.build_cf_collection(), "-f", comb_path, "-b", "-4", "-e", "4", "-s", "0.5", "-d", "0.1"],
                        stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert ev.returncode == 0, ev.stderr
    comb_rows = np.array([[float(x) for x in line.split()] for line in ev.stdout.strip().splitlines()])
    heads = re.findall(r"#CF type=(\d) isign=(-?\d+) weight=(\S+) Eg=(\S+) steps=(\d+)", r.stdout)
    assert [int(h[0]) for h in heads] == [0, 1, 2, 3]
    table = r.stdout.split("#omega ReG ImG\n")[1].strip().splitlines()
    om = np.array([float(l.split()[0]) for l in table])
    g = np.array([float(l.split()[1]) + 1j * float(l.split()[2]) for l in table])
    eng = cases.make_engine(lpp, cases.hubbard_chain(6, 3, 3))
    en = lpp.Engine(eng, {"LanczosSteps": 300, "LanczosEps": 1e-13, "SpectralSteps": 40, "SpectralEps": 0.0})
    assert abs(en.energies(0) - float(heads[0][3])) < 1e-10
    cfs = en.spectralFunction(lpp.OP_C, 1, 3, spin=0)
    gref = sum(cf(om, 0.1) for _, cf in cfs)
    for (typ, cf), h in zip(cfs, heads):
        assert cf.isign == int(h[1]) and abs(cf.weight - float(h[2])) <= 1e-9 * max(1.0, abs(cf.weight))
    assert np.abs(g - gref).max() <= 1e-8 * max(1.0, np.abs(gref).max())
    assert np.abs(comb_rows[:, 1] + 1j * 0 - g.imag).max() <= 1e-9 and np.abs(comb_rows[:, 2] - g.real).max() <= 1e-9
    eng.close()


@pytest.mark.gpu
def test_driver_two_point_matrix(lpp):
    """`lanczos_b200 -c c`: Engine::twoPoint through the C++ mirror; the trace of <cdagger_j c_i> is the number of up electrons."""
    r = run(lpp, ["-f", os.path.join(ROOT, "tests/inputs/hubbard6_gf.inp"), "-p", "14", "-c", "c"])
    assert r.returncode == 0, r.stderr
    assert abs(float(re.search(r"MatrixDiagonal = (\S+)", r.stdout).group(1)) - 3.0) < 1e-9
    rows = r.stdout.split("6 6\n")[1].strip().splitlines()[:6]
    m = np.array([[float(x) for x in line.split()] for line in rows])
    assert np.abs(m - m.T).max() < 1e-10 and np.linalg.eigvalsh(m).min() > -1e-10
