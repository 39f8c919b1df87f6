"""Generate tests/golden/*.npz from the reference's own model code (oracle/_ref, see oracle/ref_bridge.cpp).

Run here (needs /root/reference to build oracle/_ref):  python tools/make_golden.py
Every fixture holds, for one case of tests/cases.py::SMALL_CASES:
  up_words, dn_words        one-spin bases in the reference's order (BasisOneSpin / BasisOneSpinFeAs / BasisHeisenberg);
                            for Tj1Orbital (not a product basis): basis(i, SPIN_UP), basis(i, SPIN_DOWN) of every row
  nnz, rowptr, colind, values   the stored Hamiltonian of model.setupHamiltonian (full arrays when nnz <= 50 000, else
                            sha256 digests of the int64 rowptr/colind arrays plus sum / abs-sum of the values)
  x_otf                     x = 0 + H y through model.matrixVectorProduct (on-the-fly path; absent for Heisenberg)
  x_stored                  the same product through the stored matrix (numpy on the reference's CRS)
  op_*                      Engine::accModifiedState_ results for c / cdagger on the (nup-1 | nup+1, ndown) sectors
y = geometry.splitmix64_vector(rows, 7); operator sources = splitmix64_vector(rows, 11).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lanczosplusplus_b200 import geometry as geo  # noqa: E402
from oracle import reference as ref  # noqa: E402
from tests import cases  # noqa: E402

Y_SEED, SRC_SEED, FULL_CRS_NNZ = 7, 11, 50000


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def inputs_digest(case):
    h = hashlib.sha256()
    for k in sorted(case):
        v = case[k]
        h.update(k.encode())
        h.update(b"none" if v is None else np.ascontiguousarray(v, dtype=np.float64).tobytes())
    return h.hexdigest()


def make_reference(case):
    return ref.ReferenceModel(case["model"], case["nsite"], case["nup"], case["ndown"], case["orbitals"],
                              hop=case.get("hop"), jzz=case.get("jzz"), U=case.get("U"), V=case.get("V"),
                              D=case.get("D"), jpm=case.get("jpm"), w=case.get("w"))


def op_list(case):
    """(op, site, spin, orb) tuples exercised per model: fermionic c / cdagger on both spins (quirk C.3 lives in spin 1) and,
    for HubbardOneBand and Heisenberg, the spin operators sz / splus / sminus (/ n)."""
    n = case["nsite"]
    sites = sorted({0, n // 2, n - 1})
    out = []
    if case["model"] != cases.HEISENBERG:
        for op in (ref.OP_C, ref.OP_CDAGGER):
            for spin in (0, 1):
                for site in sites:
                    for orb in range(case["orbitals"]):
                        out.append((op, site, spin, orb))
    if case["model"] in (cases.HUBBARD, cases.HEISENBERG):
        for op in (ref.OP_SZ, ref.OP_SPLUS, ref.OP_SMINUS):
            for site in sites[:2]:
                out.append((op, site, 0, 0))
    if case["model"] == cases.HEISENBERG:
        for spin in (0, 1):
            out.append((ref.OP_N, sites[-1], spin, 0))
    if case["model"] in (cases.FEAS, cases.TJ):
        for op in (ref.OP_SPLUS, ref.OP_SMINUS):
            for site in sites[:2]:
                for orb in range(case["orbitals"]):
                    out.append((op, site, 0, orb))
    return out


def new_sector_of(r, case, op, spin, orb):
    """(reference handle of the destination sector, nup, ndown) or None when hasNewParts says there is none."""
    if op in (ref.OP_SZ, ref.OP_N):
        return r, case["nup"], case["ndown"]
    if case["model"] == cases.HEISENBERG:          # parts() = (twiceS, szPlusConst): Heisenberg.h:218-240
        sz = case["nup"] + (1 if op == ref.OP_SPLUS else -1)
        if sz < 0 or sz > case["nsite"]:
            return None
        return r.new_sector(1, sz), sz, 0
    has, (nu, nd) = r.has_new_parts(op, spin, orb)
    if not has or max(nu, nd) > case["nsite"] * case["orbitals"]:
        return None
    return r.new_sector(nu, nd), nu, nd


def measure_list(case):
    """operator products for Engine::measure / ModelBase::rahulMethod: (label 0 identity 1 n 2 sz 3 c, dof, bit position, transpose).
    t-J gets the diagonal ones only: the reference asserts when a product leaves the no-double-occupancy space."""
    if case["model"] == cases.HEISENBERG:
        return {}
    nb = case["nsite"] * case["orbitals"]
    s0, s1 = 0, nb - 1
    out = {"n_up": [(1, 0, s0, 0)], "nn": [(1, 0, s0, 0), (1, 1, s1, 0)], "szsz": [(2, 0, s0, 0), (2, 1, s1, 0)],
           "id_n": [(0, 0, s0, 0), (1, 1, s0, 0)]}
    if case["model"] != cases.TJ:
        out.update({"hop_up": [(3, 0, s0, 1), (3, 0, s1, 0)], "hop_dn": [(3, 1, s1, 1), (3, 1, s0, 0)],
                    "pair": [(3, 0, s0, 1), (3, 0, s1, 0), (3, 1, s1, 1), (3, 1, s0, 0)],
                    "same_site": [(3, 0, s0, 1), (3, 0, s0, 0)]})
    return out


def generate(name, case):
    r = make_reference(case)
    n = r.rows()
    w1, w2 = r.row_words(0), r.row_words(1)
    if case["model"] == cases.HEISENBERG:
        up, dn = w1, np.zeros(0, dtype=np.uint64)
    elif case["model"] == cases.TJ:            # not a product basis: the fixture keeps basis(i, spin) for every row
        up, dn = w1, w2
    else:
        n1 = int(np.argmax(w2 != w2[0])) if np.any(w2 != w2[0]) else n
        up, dn = w1[:n1].copy(), w2[::n1].copy()
    rowptr, colind, values = r.crs()
    y = geo.splitmix64_vector(n, Y_SEED)
    out = dict(inputs=np.array(inputs_digest(case)), rows=np.int64(n), up_words=up, dn_words=dn, nnz=np.int64(colind.size),
               rowptr_sha=np.array(sha(rowptr)), colind_sha=np.array(sha(colind)), values_sum=np.float64(values.sum()),
               values_abs_sum=np.float64(np.abs(values).sum()))
    if colind.size <= FULL_CRS_NNZ:
        out.update(rowptr=rowptr, colind=colind, values=values)
    xs = np.zeros(n)
    np.add.at(xs, np.repeat(np.arange(n), np.diff(rowptr)), values * y[colind])
    out["x_stored"] = xs
    if case["model"] not in (cases.HEISENBERG, cases.TJ):     # these two have no on-the-fly product in the reference
        x = np.zeros(n)
        r.matvec(x, y)
        out["x_otf"] = x
    src = geo.splitmix64_vector(n, SRC_SEED)
    ops = []
    for (op, site, spin, orb) in op_list(case):
        sec = new_sector_of(r, case, op, spin, orb)
        if sec is None:
            continue
        dst, nu, nd = sec
        if dst.rows() == 0:
            continue
        z = np.zeros(dst.rows())
        r.apply_op(dst, op, site, spin, 1.0, src, z, orb=orb)
        key = "op_%d_%d_%d_%d" % (op, site, spin, orb)
        out[key] = z
        ops.append(dict(key=key, op=op, site=site, spin=spin, orb=orb, nup=nu, ndown=nd))
    out["ops"] = np.array(json.dumps(ops))
    meas = {}
    for key, oplist in measure_list(case).items():
        if any(o[0] == 3 for o in oplist) and (case["nup"] == 0 or case["ndown"] == 0):
            continue
        psi_new = r.rahul(oplist, src)
        meas[key] = dict(ops=[list(map(int, o)) for o in oplist], value=float(src @ psi_new))
        if key in ("hop_up", "pair", "nn"):
            out["rahul_" + key] = psi_new
    out["measure"] = np.array(json.dumps(meas))
    return out


def main():
    gdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gdir, exist_ok=True)
    total = 0
    for name, case in list(cases.SMALL_CASES.items()) + list(cases.TJ_CASES.items()):
        data = generate(name, case)
        path = os.path.join(gdir, name + ".npz")
        np.savez_compressed(path, **data)
        total += os.path.getsize(path)
        print("%-28s rows %6d nnz %7d ops %2d -> %7d bytes" % (name, data["rows"], data["nnz"],
                                                              len(json.loads(str(data["ops"]))), os.path.getsize(path)))
    print("total %.1f KB" % (total / 1024))


if __name__ == "__main__":
    main()
