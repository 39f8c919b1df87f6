"""CPU-side parity of the product's host/device code (lpp_device.cuh compiled by g++) against the oracle:
basis words, ranks and CRS structure bit-exact; values and x += H y to 1e-13 relative."""
import numpy as np
import pytest

from lanczosplusplus_b200 import geometry as geo
from tests import cases
from tests.hostcheck import HostModel, lib as hclib


@pytest.mark.parametrize("name", sorted(cases.SMALL_CASES))
@pytest.mark.parametrize("tables", [0, 1])
def test_basis_rank_crs_matvec(oracle, name, tables):
    case = cases.SMALL_CASES[name]
    o = cases.make_oracle(oracle, case, fast_rank=0)
    h = HostModel(case, use_tables=tables)
    assert h.rows() == o.rows()
    nspin = 1 if case["model"] == cases.HEISENBERG else 2
    for spin in range(nspin):
        bo, bh = o.basis(spin), h.basis(spin)
        assert np.array_equal(bo, bh)                       # bit-exact ordering
        step = max(1, len(bo) // 97)
        for i in range(0, len(bo), step):
            assert h.rank(spin, bo[i]) == i == o.rank(spin, bo[i])
    rp0, ci0, v0 = o.crs()
    rp1, ci1, v1 = h.crs()
    assert np.array_equal(rp0, rp1) and np.array_equal(ci0, ci1)   # bit-exact structure
    assert np.abs(v0 - v1).max() <= 1e-14 * max(1.0, np.abs(v0).max())
    y = geo.splitmix64_vector(o.rows(), 42)
    x0 = geo.splitmix64_vector(o.rows(), 7)
    x1 = x0.copy()
    o.matvec(x0, y, faithful=True)
    h.matvec(x1, y)
    assert np.abs(x0 - x1).max() <= 1e-13 * max(1.0, np.abs(x0).max())


@pytest.mark.parametrize("name", sorted(cases.TJ_CASES))
@pytest.mark.parametrize("tables", [0, 1])
def test_tj_basis_rank_crs_matvec(oracle, name, tables):
    """Tj1Orbital: closed-form basis / rank of the product's code against the literal restatement (sorted combined words)."""
    case = cases.TJ_CASES[name]
    o = cases.make_oracle(oracle, case)
    h = HostModel(case, use_tables=tables)
    assert h.rows() == o.rows()
    w1, w2 = o.row_words(0), o.row_words(1)
    assert np.array_equal(h.row_words(0), w1) and np.array_equal(h.row_words(1), w2)   # bit-exact ordering
    for i in range(0, len(w1), max(1, len(w1) // 97)):
        assert h.rank_pair(w1[i], w2[i]) == i == o.perfect_index(w1[i], w2[i])
    rp0, ci0, v0 = o.crs()
    rp1, ci1, v1 = h.crs()
    assert np.array_equal(rp0, rp1) and np.array_equal(ci0, ci1)
    assert np.abs(v0 - v1).max() <= 1e-14 * max(1.0, np.abs(v0).max())
    y = geo.splitmix64_vector(o.rows(), 42)
    x0 = geo.splitmix64_vector(o.rows(), 7)
    x1 = x0.copy()
    o.matvec(x0, y, faithful=True)
    h.matvec(x1, y)
    assert np.abs(x0 - x1).max() <= 1e-13 * max(1.0, np.abs(x0).max())


def test_tj_larger_basis_bit_exact(oracle):
    case = cases.tj_square(4, 3, 5, 4)          # 12 sites, 3 holes: dim 27 720
    o = cases.make_oracle(oracle, case)
    h = HostModel(case, use_tables=1)
    assert h.rows() == o.rows() == 27720
    assert np.array_equal(h.row_words(0), o.row_words(0)) and np.array_equal(h.row_words(1), o.row_words(1))


def test_bigger_bases_bit_exact(oracle):
    for case in (cases.hubbard_square(4, 4, 8, 8), cases.feas_cluster(2, 4, 6, 6), cases.heisenberg_ring(16, 8),
                 cases.hubbard_chain(18, 9, 9)):
        o = cases.make_oracle(oracle, case, fast_rank=1)
        h = HostModel(case, use_tables=1)
        nspin = 1 if case["model"] == cases.HEISENBERG else 2
        for spin in range(nspin):
            b = o.basis(spin)
            assert np.array_equal(b, h.basis(spin))
            idx = np.random.default_rng(3).integers(0, len(b), 200)
            for i in idx:
                assert h.rank(spin, b[i]) == i


def test_apply_op_matches_oracle(oracle):
    src_case = cases.hubbard_chain(6, 3, 3)
    for op, spin, dn in ((oracle.OP_C, 0, -1), (oracle.OP_CDAGGER, 0, 1), (oracle.OP_C, 1, -1), (oracle.OP_CDAGGER, 1, 1),
                         (oracle.OP_N, 0, 0), (oracle.OP_N, 1, 0)):
        dst_case = cases.hubbard_chain(6, 3 + (dn if spin == 0 else 0), 3 + (dn if spin == 1 else 0))
        os_, od = cases.make_oracle(oracle, src_case), cases.make_oracle(oracle, dst_case)
        hs, hd = HostModel(src_case), HostModel(dst_case)
        v = geo.splitmix64_vector(os_.rows(), 5)
        for site in range(6):
            z0, z1 = np.zeros(od.rows()), np.zeros(od.rows())
            os_.apply_op(od, op, site, spin, 0.7, v, z0)
            hs.apply_op(hd, op, site, spin, 0.7, v, z1)
            assert np.array_equal(z0, z1)


@pytest.mark.parametrize("name", ["feas3", "feas4_interorb_otfquirk", "tj6_pbc", "tj8_V", "tj6_full"])
def test_apply_op_other_models_match_oracle(oracle, name):
    """c / cdagger for FeAsBasedSc (every orbital) and Tj1Orbital through the product's gather form vs the oracle's scatter."""
    src_case = (cases.TJ_CASES if name.startswith("tj") else cases.SMALL_CASES)[name]
    os_, hs = cases.make_oracle(oracle, src_case), HostModel(src_case)
    v = geo.splitmix64_vector(os_.rows(), 5)
    n, no = src_case["nsite"], src_case["orbitals"]
    done = 0
    for op, dn in ((oracle.OP_C, -1), (oracle.OP_CDAGGER, 1)):
        for spin in (0, 1):
            nu, nd = src_case["nup"] + (dn if spin == 0 else 0), src_case["ndown"] + (dn if spin == 1 else 0)
            if min(nu, nd) < 0 or max(nu, nd) > n * no or (nu == 0 and nd == 0):
                continue
            if src_case["model"] == cases.TJ and nu + nd > n:
                continue
            dst_case = dict(src_case, nup=nu, ndown=nd)
            od, hd = cases.make_oracle(oracle, dst_case), HostModel(dst_case)
            for site in range(n):
                for orb in range(no):
                    z0, z1 = np.zeros(od.rows()), np.zeros(od.rows())
                    os_.apply_op(od, op, site, spin, 0.7, v, z0, orb=orb)
                    hs.apply_op(hd, op, site, spin, 0.7, v, z1, orb=orb)
                    assert np.array_equal(z0, z1), (op, spin, site, orb)
                    done += 1
    assert done > 0


@pytest.mark.parametrize("name", ["c1_hub8", "hub_rand7", "hub_empty_dn", "heis12", "heis10_field", "feas3", "feas_2x2", "tj8_V", "tj_3x3"])
def test_spin_operators_match_reference_fixture(name):
    """sz / splus / sminus (/ n): the product's gather form (lpp_apply_spin_op_source, compiled for the host) against the
    fixtures produced by the reference's getBraIndex / doSignSpSm."""
    from tests import golden_util as gu
    case = (cases.TJ_CASES if name.startswith("tj") else cases.SMALL_CASES)[name]
    g = gu.load(name, case)
    hs = HostModel(case)
    src = geo.splitmix64_vector(hs.rows(), gu.SRC_SEED)
    seen = 0
    for rec in gu.ops(g):
        if rec["op"] not in (2, 4, 5, 6):
            continue
        hd = HostModel(dict(case, nup=rec["nup"], ndown=rec["ndown"]))
        z = np.zeros(hd.rows())
        hs.apply_op(hd, rec["op"], rec["site"], rec["spin"], 1.0, src, z, orb=rec["orb"])
        assert np.array_equal(z, g[rec["key"]]), rec
        seen += 1
    assert seen >= 3


@pytest.mark.parametrize("name", ["c1_hub8", "hub_rand7", "hub6_pbc_V", "feas4", "feas_2x2", "tj8_V", "tj_3x3"])
def test_measure_products_match_reference_rahul_method(name):
    """Engine::measure: the product's lpp_rahul_apply (compiled for the host) against the reference's own ModelBase::rahulMethod
    (fixtures): the modified vectors element by element, the expectation values to rounding."""
    from tests import golden_util as gu
    case = (cases.TJ_CASES if name.startswith("tj") else cases.SMALL_CASES)[name]
    g = gu.load(name, case)
    h = HostModel(case)
    src = geo.splitmix64_vector(h.rows(), gu.SRC_SEED)
    meas = gu.measures(g)
    assert len(meas) >= 4
    for key, rec in meas.items():
        psi_new = h.rahul(rec["ops"], src)
        assert abs(src @ psi_new - rec["value"]) <= 1e-12 * max(1.0, abs(rec["value"])), key
        if "rahul_" + key in g.files:
            assert np.array_equal(psi_new, g["rahul_" + key]), key


def test_splitmix_matches_numpy():
    v = geo.splitmix64_vector(64, 1234, offset=10**12)
    for i in range(64):
        assert hclib().hc_splitmix(1234, 10**12 + i) == v[i]
