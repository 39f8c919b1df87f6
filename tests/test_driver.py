"""The stand-alone C++ driver (host/lanczos_b200.cpp): input parsing on the CPU, energies on the GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(lpp, args):
    exe = lpp._lib.build_driver()
    return subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


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
