"""The CPU oracle against the reference's own model code, live (oracle/_ref/liblpp_ref.so, built from /root/reference/src
with oracle/psimag_shim).  Skipped where neither the prebuilt library nor the reference sources exist; tests/test_golden.py
holds the same comparison against committed fixtures.
"""
import numpy as np
import pytest

from lanczosplusplus_b200 import geometry as geo
from tests import cases

reference = pytest.importorskip("oracle.reference")
if reference.build() is None:
    pytest.skip("oracle/_ref is not built and /root/reference is absent", allow_module_level=True)

MEDIUM = {
    "hub10": cases.hubbard_chain(10, 5, 5),
    "hub_4x3": cases.hubbard_square(4, 3, 4, 3),
    "hub_rand9": cases.hubbard_random(9, 4, 3, 21),
    "feas_input100_small": cases.feas_chain(4, 4, 4),
    "feas_2x2_quirk": dict(cases.feas_cluster(2, 2, 2, 3), feas_u3_all_pairs=0),
    "heis14": cases.heisenberg_ring(14, 7),
    "tj10": cases.tj_chain(10, 4, 4, periodic=True),
    "tj_4x3": cases.tj_square(4, 3, 5, 4),
    "tj9_rand": cases.tj_chain(9, 3, 4, seed=11),
}


def make_reference(case):
    return reference.ReferenceModel(case["model"], case["nsite"], case["nup"], case["ndown"], case["orbitals"],
                                    hop=case.get("hop"), jzz=case.get("jzz"), U=case.get("U"), V=case.get("V"),
                                    D=case.get("D"), jpm=case.get("jpm"), w=case.get("w"))


@pytest.mark.parametrize("name", sorted(MEDIUM))
def test_oracle_vs_live_reference(oracle, name):
    case = MEDIUM[name]
    r = make_reference(case)
    o = cases.make_oracle(oracle, dict(case, feas_u3_all_pairs=0), fast_rank=1)
    n = r.rows()
    assert o.rows() == n
    w1, w2 = r.row_words(0), r.row_words(1)
    b1 = o.basis(0)
    if case["model"] == cases.TJ:
        assert np.array_equal(w1, o.row_words(0)) and np.array_equal(w2, o.row_words(1))
    elif case["model"] == cases.HEISENBERG:
        assert np.array_equal(w1, b1)
    else:
        b2 = o.basis(1)
        assert np.array_equal(w1, np.tile(b1, len(b2))) and np.array_equal(w2, np.repeat(b2, len(b1)))
    for i in range(0, n, max(1, n // 50)):
        assert r.perfect_index(int(w1[i]), int(w2[i])) == i
    rp, ci, v = r.crs()
    rp0, ci0, v0 = o.crs()
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0) and np.array_equal(v, v0)
    if case["model"] not in (cases.HEISENBERG, cases.TJ):
        y = geo.splitmix64_vector(n, 3)
        x = geo.splitmix64_vector(n, 4)
        x0 = x.copy()
        r.matvec(x, y)                                                     # accumulates into x (x += H y)
        o.matvec(x0, y, faithful=True)
        assert np.abs(x - x0).max() <= 1e-14 * max(1.0, np.abs(x0).max())


def test_apply_op_and_new_parts_vs_live_reference(oracle):
    case = cases.hubbard_chain(7, 3, 4, U=3.0, periodic=True)
    r = make_reference(case)
    o = cases.make_oracle(oracle, case)
    src = geo.splitmix64_vector(r.rows(), 5)
    for op, d in ((reference.OP_C, -1), (reference.OP_CDAGGER, 1)):
        for spin in (0, 1):
            has, (nu, nd) = r.has_new_parts(op, spin)
            assert has and (nu, nd) == (3 + (d if spin == 0 else 0), 4 + (d if spin == 1 else 0))
            rd = r.new_sector(nu, nd)
            od = cases.make_oracle(oracle, dict(case, nup=nu, ndown=nd))
            for site in range(7):
                z, z0 = np.zeros(rd.rows()), np.zeros(od.rows())
                r.apply_op(rd, op, site, spin, 0.5, src, z)
                o.apply_op(od, op, site, spin, 0.5, src, z0)
                assert np.array_equal(z, z0), (op, spin, site)
