""".comb files (LanczosDriver1.h:147-181) and their evaluator (the role of PsimagLite's continuedFractionCollection): the layout
is written by lanczosplusplus_b200/comb.py and host/comb_io.h, read back by both, accepted by the reference's own
scripts/extractOrbitals.pl, and evaluated by host/cf_collection with the column order scripts/sqomega.pl reads (omega, Im, Re)."""
import os
import subprocess

import numpy as np
import pytest

from lanczosplusplus_b200 import comb

REF_SCRIPT = "/root/reference/scripts/extractOrbitals.pl"


def _fractions(lpp):
    rng = np.random.default_rng(4)
    cfs, keys = [], []
    for typ in range(4):
        n = 12 + typ
        a, b = rng.uniform(-3, 3, n), rng.uniform(0.5, 2.0, n)
        s = -1 if (typ & 1) else 1
        cfs.append(lpp.ContinuedFraction(a, b, -3.25, (0.5 + 0.1 * typ) * (-1.0 if typ > 1 else 1.0), -s))
        keys.append("0,%d,0,0" % typ)
    return cfs, keys


def test_comb_roundtrip_and_evaluator(lpp, tmp_path):
    cfs, keys = _fractions(lpp)
    path = str(tmp_path / "input0.comb")
    comb.write_comb(path, 1, 3, keys, cfs)
    back = comb.read_comb(path)
    assert back["site0"] == 1 and back["site1"] == 3 and back["index_to_cf"] == keys and len(back["cfs"]) == 4
    for cf, d in zip(cfs, back["cfs"]):
        assert np.array_equal(cf.a, d["a"]) and np.array_equal(cf.b, d["b"])
        assert cf.weight == d["weight"] and cf.eg == d["eg"] and cf.isign == d["isign"]
    exe = lpp._lib.build_cf_collection()
    r = subprocess.run([exe, "-f", path, "-b", "-4", "-e", "4", "-s", "0.25", "-d", "0.1"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    rows = np.array([[float(x) for x in line.split()] for line in r.stdout.strip().splitlines()])
    omega = rows[:, 0]
    g = sum(cf(omega, 0.1) for cf in cfs)
    assert len(omega) == 33
    assert np.abs(rows[:, 1] - g.imag).max() <= 1e-12 * max(1.0, np.abs(g).max())      # column 1 = Im, column 2 = Re
    assert np.abs(rows[:, 2] - g.real).max() <= 1e-12 * max(1.0, np.abs(g).max())


@pytest.mark.skipif(not os.path.exists(REF_SCRIPT), reason="the reference's scripts are not on this machine")
def test_reference_extract_orbitals_accepts_the_file(lpp, tmp_path):
    """scripts/extractOrbitals.pl (unchanged, run from /root/reference) selects the (orb1, orb2) fractions of a .comb file;
    its output must still be a collection the evaluator reads."""
    cfs, keys = _fractions(lpp)
    path = str(tmp_path / "input0.comb")
    comb.write_comb(path, 0, 0, keys, cfs)
    r = subprocess.run(["perl", REF_SCRIPT, "0", "0", "1"], stdin=open(path), stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True)
    assert r.returncode == 0, r.stderr
    assert "Offset= 0    total= 4" in r.stderr
    out = str(tmp_path / "input0.comb2")
    open(out, "w").write(r.stdout)
    assert r.stdout.startswith("#CONTINUEDFRACTIONCOLLECTION=4\n#Avector\n")
    back = comb.read_comb(out)
    assert len(back["cfs"]) == 4
    for cf, d in zip(cfs, back["cfs"]):
        assert np.array_equal(cf.a, d["a"]) and cf.weight == d["weight"]
    exe = lpp._lib.build_cf_collection()
    r2 = subprocess.run([exe, "-f", out, "-b", "0", "-e", "1", "-s", "0.5", "-d", "0.1"], stdout=subprocess.PIPE,
                        stderr=subprocess.PIPE, text=True)
    assert r2.returncode == 0 and len(r2.stdout.strip().splitlines()) == 3, r2.stderr
