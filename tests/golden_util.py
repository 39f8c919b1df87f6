"""Loader for tests/golden/*.npz: outputs of the reference's own model code (tools/make_golden.py, oracle/ref_bridge.cpp)."""
import hashlib
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
Y_SEED, SRC_SEED = 7, 11


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def inputs_digest(case):
    h = hashlib.sha256()
    for k in sorted(case):
        v = case[k]
        h.update(k.encode())
        h.update(b"none" if v is None else np.ascontiguousarray(v, dtype=np.float64).tobytes())
    return h.hexdigest()


def load(name, case):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    assert str(g["inputs"]) == inputs_digest(case), "tests/cases.py changed: regenerate with tools/make_golden.py"
    return g


def ops(g):
    return json.loads(str(g["ops"]))


def measures(g):
    return json.loads(str(g["measure"])) if "measure" in g.files else {}


def check_crs(g, rowptr, colind, values, tol=1e-14):
    """CRS structure bit-exact against the reference's stored Hamiltonian; values to `tol`."""
    assert int(g["nnz"]) == colind.size
    rowptr = np.asarray(rowptr, dtype=np.int64)
    colind = np.asarray(colind, dtype=np.int64)
    assert sha(rowptr) == str(g["rowptr_sha"]) and sha(colind) == str(g["colind_sha"])
    scale = max(1.0, float(g["values_abs_sum"]))
    assert abs(values.sum() - float(g["values_sum"])) <= 1e-13 * scale
    assert abs(np.abs(values).sum() - float(g["values_abs_sum"])) <= 1e-13 * scale
    if "values" in g.files:
        assert np.array_equal(g["rowptr"], rowptr) and np.array_equal(g["colind"], colind)
        assert np.abs(g["values"] - values).max() <= tol * max(1.0, np.abs(g["values"]).max())
