"""GPU parity tests (run on a real B200 with `pytest -m gpu`): the CUDA path, called through the C-ABI, against the
CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): basis words, ranks, CRS rowptr/colind bit-exact; x += H y <= 1e-13 relative;
tridiagonal alpha/beta and ground-state energies <= 1e-10 relative; continued-fraction spectra <= 1e-8.
"""
import numpy as np
import pytest

from lanczosplusplus_b200 import geometry as geo
from tests import cases

pytestmark = pytest.mark.gpu


def kernels_for(lpp, case):
    if case["model"] == cases.HEISENBERG:
        return [lpp.KERNEL_GENERIC, lpp.KERNEL_TABLE, lpp.KERNEL_STORED, lpp.KERNEL_AUTO]
    return [lpp.KERNEL_GENERIC, lpp.KERNEL_TABLE, lpp.KERNEL_TILED, lpp.KERNEL_STORED, lpp.KERNEL_AUTO]


def relerr(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


@pytest.mark.parametrize("name", sorted(cases.SMALL_CASES))
def test_small_cases_full_parity(lpp, oracle, name):
    case = cases.SMALL_CASES[name]
    o = cases.make_oracle(oracle, case, fast_rank=0)
    e = cases.make_engine(lpp, case)
    assert e.rows() == o.rows()
    nspin = 1 if case["model"] == cases.HEISENBERG else 2
    for spin in range(nspin):
        b = o.basis(spin)
        assert np.array_equal(b, e.basis(spin))                          # bit-exact, built on device
        assert np.array_equal(e.perfectIndex(spin, b), np.arange(len(b), dtype=np.uint64))
    rp0, ci0, v0 = o.crs()
    rp1, ci1, v1 = e.setupHamiltonian()
    assert np.array_equal(rp0, rp1) and np.array_equal(ci0, ci1)        # bit-exact CRS structure
    assert np.abs(v0 - v1).max() <= 1e-14 * max(1.0, np.abs(v0).max())
    y = geo.splitmix64_vector(o.rows(), 42)
    x0 = geo.splitmix64_vector(o.rows(), 7)
    xref = x0.copy()
    o.matvec(xref, y, faithful=True)                                      # x += H y, reference semantics
    xref_stored = x0.copy()
    oracle.crs_matvec(rp0, ci0, v0, xref_stored, y)                       # InternalProductStored semantics
    for k in kernels_for(lpp, case):
        x = x0.copy()
        e.matrixVectorProduct(x, y, kernel=k)
        # (FeAs: the literal on-the-fly doTask and the stored builder differ when U[3] != 0)
        assert relerr(x, xref_stored if k == lpp.KERNEL_STORED else xref) <= 1e-13, (name, k)
    e.close()


@pytest.mark.parametrize("name", ["c1_hub8", "hub_rand7", "feas4", "feas_2x2", "heis12", "hub6_pbc_V"])
def test_lanczos_tridiagonal_and_energy(lpp, oracle, name):
    case = cases.SMALL_CASES[name]
    o = cases.make_oracle(oracle, case, fast_rank=1)
    init = geo.splitmix64_vector(o.rows(), 1234)
    e0, z0, a0, b0 = o.ground_state(init, 200, 1e-12, 4)
    for k in kernels_for(lpp, case):
        e = cases.make_engine(lpp, case, kernel=k)
        solver = lpp.LanczosSolver(e, lpp.ParametersForSolver(steps=200, eps=1e-12, minsteps=4))
        e1, z1, a1, b1 = solver.computeOneState(init, want_vector=True, copy_vector=True)
        assert abs(e1 - e0) <= 1e-10 * max(1.0, abs(e0)), (name, k)
        n = min(len(a0), len(a1), 25)                                     # before rounding noise is amplified
        assert relerr(a1[:n], a0[:n]) <= 1e-10 and relerr(b1[:n], b0[:n]) <= 1e-10, (name, k)
        assert abs(len(a1) - len(a0)) <= 2
        assert abs(np.linalg.norm(z1) - 1.0) < 1e-9
        r = -e1 * z1
        o.matvec(r, z1, faithful=False)
        assert np.linalg.norm(r) < 1e-5                                   # Ritz residual ||H z - E z||
        assert abs(abs(z1 @ z0) - 1.0) < 1e-8
        # device-generated initial vector == geometry.splitmix64_vector
        e2, _, a2, b2 = solver.computeOneState(None, want_vector=False)
        assert abs(e2 - e1) <= 1e-12 * max(1.0, abs(e1)) and np.array_equal(a2, a1)
        e.close()


def test_known_answers(lpp):
    pins = [("input0", -2.0 * np.sqrt(5.0)), ("hub2", 2.0 - np.sqrt(8.0)), ("c1_hub8", -4.235806999130),
            ("heis4", -2.0), ("heis12", -5.387390917445), ("feas4", -1.957490136573)]
    for name, ref in pins:
        for k in (lpp.KERNEL_AUTO, lpp.KERNEL_STORED):
            e = cases.make_engine(lpp, cases.SMALL_CASES[name], kernel=k)
            en = lpp.Engine(e, {"LanczosSteps": 300, "LanczosEps": 1e-13})
            assert abs(en.energies(0) - ref) <= 1e-10 * max(1.0, abs(ref)), (name, k)
            e.close()


def test_medium_sizes(lpp, oracle):
    # input100.inp sector (dim 48400): survey-time value; Heisenberg 16-ring; Hubbard 12 sites (dim 853776) mat-vec
    e = cases.make_engine(lpp, cases.feas_chain(6, 3, 3))
    assert abs(lpp.Engine(e, {"LanczosSteps": 400}).energies(0) - (-3.099464014219)) < 1e-9
    e.close()
    e = cases.make_engine(lpp, cases.heisenberg_ring(16, 8))
    assert abs(lpp.Engine(e, {"LanczosSteps": 400}).energies(0) - (-7.142296360617)) < 1e-9
    e.close()
    case = cases.hubbard_chain(12, 6, 6, periodic=True)
    o = cases.make_oracle(oracle, case, fast_rank=1)
    e = cases.make_engine(lpp, case)
    y = geo.splitmix64_vector(o.rows(), 42)
    xref = np.zeros(o.rows())
    o.matvec(xref, y, faithful=False)
    for k in (lpp.KERNEL_GENERIC, lpp.KERNEL_TABLE, lpp.KERNEL_TILED):
        x = np.zeros(o.rows())
        e.matrixVectorProduct(x, y, kernel=k)
        assert relerr(x, xref) <= 1e-13, k
    e.close()
    case = cases.feas_cluster(2, 3, 4, 4)                                  # dim 245025, orbital hoppings + U2/U3
    o = cases.make_oracle(oracle, case, fast_rank=1)
    e = cases.make_engine(lpp, case)
    y = geo.splitmix64_vector(o.rows(), 42)
    xref = np.zeros(o.rows())
    o.matvec(xref, y, faithful=False)
    for k in (lpp.KERNEL_GENERIC, lpp.KERNEL_TABLE, lpp.KERNEL_TILED):
        x = np.zeros(o.rows())
        e.matrixVectorProduct(x, y, kernel=k)
        assert relerr(x, xref) <= 1e-13, k
    e.close()


def test_continued_fraction_parity(lpp, oracle):
    """Engine::spectralFunction (Engine.h:133-206) for c at sites (1,1) and (1,3), spin up, 6-site chain.
    Both sides start from the SAME ground-state vector (the oracle's), so the comparison isolates the operator
    application, the Krylov recurrence and the continued fraction; bar: spectra within 1e-8."""
    case = cases.hubbard_chain(6, 3, 3)
    o = cases.make_oracle(oracle, case)
    init = geo.splitmix64_vector(o.rows(), 1234)
    e0, z0, _, _ = o.ground_state(init, 300, 1e-13, 4)
    omega = np.linspace(-8, 8, 161)
    for steps, tol in ((40, 1e-8), (150, 1e-4)):     # beyond ~50 steps rounding noise (ghost states) is amplified
        eng = cases.make_engine(lpp, case)
        en = lpp.Engine(eng, {"LanczosSteps": 300, "LanczosEps": 1e-13, "SpectralSteps": steps, "SpectralEps": 0.0},
                        init=init)
        assert abs(en.energies(0) - e0) < 1e-10
        zg = eng.get_vector(0)
        assert abs(abs(zg @ z0) - 1.0) < 1e-9
        eng.set_groundstate(z0)
        en.energy = e0
        for (isite, jsite) in ((1, 1), (1, 3)):
            cfs = en.spectralFunction(lpp.OP_C, isite, jsite, spin=0)
            assert len(cfs) == (2 if isite == jsite else 4)
            for typ, cf in cfs:
                lop = oracle.OP_C if (typ & 1) else oracle.OP_CDAGGER
                dn = -1 if lop == oracle.OP_C else 1
                od = cases.make_oracle(oracle, cases.hubbard_chain(6, 3 + dn, 3))
                phi = np.zeros(od.rows())
                o.apply_op(od, lop, isite, 0, 1.0, z0, phi)
                o.apply_op(od, lop, jsite, 0, -1.0 if typ > 1 else 1.0, z0, phi)
                a, b = od.decomposition(phi, steps=steps, eps=0.0)
                weight = (phi @ phi) * (-1.0 if typ > 1 else 1.0) * (1.0 if isite == jsite else 0.5)
                assert abs(cf.weight - weight) <= 1e-12 * max(1.0, abs(weight))
                n = min(len(a), 10)   # dim is only 225/400 here: rounding noise grows fast with the step count
                assert relerr(cf.a[:n], a[:n]) <= 1e-10 and relerr(cf.b[:n], b[:n]) <= 1e-10
                s = -1 if (typ & 1) else 1
                gref = oracle.cf_eval(a, b, e0, weight, -s, omega, 0.1)
                g = cf(omega, 0.1)
                assert np.abs(g - gref).max() <= tol * max(1.0, np.abs(gref).max()), (steps, isite, jsite, typ)
        eng.close()


def test_full_size_properties_c3(lpp):
    """Config 3 (4x4 Hubbard, 8 up 8 down, dim 165 636 900): size-independent checks at the benchmark size."""
    case = cases.hubbard_square(4, 4, 8, 8, U=4.0)
    e = cases.make_engine(lpp, case)
    n = e.rows()
    assert n == 165636900
    y1 = geo.splitmix64_vector(n, 42)
    y2 = geo.splitmix64_vector(n, 43)
    h1 = np.zeros(n)
    e.matrixVectorProduct(h1, y1, kernel=lpp.KERNEL_TILED)
    h1t = np.zeros(n)
    e.matrixVectorProduct(h1t, y1, kernel=lpp.KERNEL_TABLE)
    assert relerr(h1, h1t) <= 1e-13                                       # independent kernels agree
    h2 = np.zeros(n)
    e.matrixVectorProduct(h2, y2, kernel=lpp.KERNEL_TILED)
    assert abs(y2 @ h1 - y1 @ h2) <= 1e-10 * abs(y2 @ h1)                 # Hermiticity <y2|H y1> = <y1|H y2>
    h12 = h1.copy()
    e.matrixVectorProduct(h12, y2, kernel=lpp.KERNEL_TILED)              # accumulate semantics + linearity
    assert relerr(h12, h1 + h2) <= 1e-13
    del h1, h1t, h2, h12, y1, y2
    e.close()
    # U = 0: free fermions, analytic ground state = 2 * sum of the 8 lowest single-particle levels
    levels = np.sort(np.linalg.eigvalsh(geo.square(4, 4, -1.0)))
    # the half-filled shell is degenerate: any of the degenerate ground states has this energy
    eref = 2.0 * levels[:8].sum()
    e = cases.make_engine(lpp, cases.hubbard_square(4, 4, 8, 8, U=0.0))
    solver = lpp.LanczosSolver(e, lpp.ParametersForSolver(steps=120, eps=1e-11))
    en, _, a, b = solver.computeOneState(None, want_vector=False)
    assert abs(en - eref) <= 1e-8 * abs(eref)
    e.close()


@pytest.mark.parametrize("cap", ["3", "16", "200"])
def test_down_tile_kernel_with_small_blocks(lpp, oracle, cap, monkeypatch):
    """The shared-memory tile kernel of the down sweep (lpp_dtile.cu) with the block size capped, so that small bases are
    split into several blocks and hops leave their block (the external-operand path), against the oracle."""
    monkeypatch.setenv("LPP_DTILE", "1")
    monkeypatch.setenv("LPP_DTILE_CAP", cap)
    for name in ("c1_hub8", "hub6_pbc_V", "hub_3x3", "feas4", "feas_2x2"):
        case = cases.SMALL_CASES[name]
        o = cases.make_oracle(oracle, case, fast_rank=1)
        e = cases.make_engine(lpp, case)
        y = geo.splitmix64_vector(o.rows(), 42)
        x0 = geo.splitmix64_vector(o.rows(), 7)
        xref = x0.copy()
        o.matvec(xref, y, faithful=False)
        x = x0.copy()
        e.matrixVectorProduct(x, y, kernel=lpp.KERNEL_TILED)
        assert relerr(x, xref) <= 1e-13, (name, cap)
        e.close()
    case = cases.hubbard_chain(12, 6, 6, periodic=True)                    # dim 853776, 924 down states
    o = cases.make_oracle(oracle, case, fast_rank=1)
    e = cases.make_engine(lpp, case)
    y = geo.splitmix64_vector(o.rows(), 42)
    xref = np.zeros(o.rows())
    o.matvec(xref, y, faithful=False)
    x = np.zeros(o.rows())
    e.matrixVectorProduct(x, y, kernel=lpp.KERNEL_TILED)
    assert relerr(x, xref) <= 1e-13, cap
    e.close()


def test_row_walking_down_kernel(lpp, oracle, monkeypatch):
    """k_sweep_down_rows (opt-in, lpp_dtile.cu): unconditional gathers with sign-split slot groups and self-padding."""
    monkeypatch.setenv("LPP_DROWS", "1")
    for case in (cases.SMALL_CASES["c1_hub8"], cases.SMALL_CASES["hub6_pbc_V"], cases.hubbard_square(4, 3, 6, 6),
                 cases.hubbard_chain(12, 6, 5, periodic=True)):
        o = cases.make_oracle(oracle, case, fast_rank=1)
        e = cases.make_engine(lpp, case)
        y = geo.splitmix64_vector(o.rows(), 42)
        x0 = geo.splitmix64_vector(o.rows(), 7)
        xref = x0.copy()
        o.matvec(xref, y, faithful=False)
        x = x0.copy()
        e.matrixVectorProduct(x, y, kernel=lpp.KERNEL_TILED)
        assert relerr(x, xref) <= 1e-13
        e.close()


@pytest.mark.parametrize("rows,pc,nb", [("3", "64", "4"), ("20", "64", "8"), ("72", "128", "4")])
def test_staged_down_kernel(lpp, oracle, rows, pc, nb, monkeypatch):
    """k_sweep_down_staged (opt-in, LPP_DSTAGE=1): runs of consecutive down states staged in shared memory; the cap on the
    run length moves the split between in-run (shared memory) and other (L2) sources, down to runs of 1-3 rows."""
    monkeypatch.setenv("LPP_DSTAGE", "1")
    monkeypatch.setenv("LPP_DSTAGE_ROWS", rows)
    monkeypatch.setenv("LPP_DSTAGE_PC", pc)
    monkeypatch.setenv("LPP_DSTAGE_NB", nb)
    for case in (cases.SMALL_CASES["c1_hub8"], cases.SMALL_CASES["hub6_pbc_V"], cases.SMALL_CASES["hub_rand7"],
                 cases.SMALL_CASES["hub_3x3"], cases.hubbard_square(4, 3, 6, 6), cases.hubbard_chain(12, 6, 5, periodic=True)):
        o = cases.make_oracle(oracle, case, fast_rank=1)
        e = cases.make_engine(lpp, case)
        y = geo.splitmix64_vector(o.rows(), 42)
        x0 = geo.splitmix64_vector(o.rows(), 7)
        xref = x0.copy()
        o.matvec(xref, y, faithful=False)
        x = x0.copy()
        e.matrixVectorProduct(x, y, kernel=lpp.KERNEL_TILED)
        assert relerr(x, xref) <= 1e-13
        e.close()


def test_twospin_rows_kernel(lpp, oracle, monkeypatch):
    """k_sweep_twospin_rows (opt-in, LPP_TWOSPIN_ROWS=1): FeAs on-site two-spin terms with one CTA per down state."""
    monkeypatch.setenv("LPP_TWOSPIN_ROWS", "1")
    for name in ("feas3", "feas4", "feas4_interorb_otfquirk", "feas_2x2"):
        case = cases.SMALL_CASES[name]
        o = cases.make_oracle(oracle, case, fast_rank=1)
        e = cases.make_engine(lpp, case)
        y = geo.splitmix64_vector(o.rows(), 42)
        x0 = geo.splitmix64_vector(o.rows(), 7)
        xref = x0.copy()
        o.matvec(xref, y, faithful=False)
        x = x0.copy()
        e.matrixVectorProduct(x, y, kernel=lpp.KERNEL_TILED)
        assert relerr(x, xref) <= 1e-13, name
        e.close()


def test_large_up_basis_c5_structure(lpp, oracle):
    """Config-5 structure at a size one GPU and the oracle handle quickly: 18-site open chain, 9 up electrons (48 620 up
    states: an up segment no longer fits in shared memory, so the blocked up-sweep path runs), 2 down; dim 7 438 860."""
    case = cases.hubbard_chain(18, 9, 2)
    o = cases.make_oracle(oracle, case, fast_rank=1)
    e = cases.make_engine(lpp, case)
    assert e.rows() == o.rows() == 48620 * 153
    y = geo.splitmix64_vector(o.rows(), 42)
    x0 = geo.splitmix64_vector(o.rows(), 7)
    xref = x0.copy()
    o.matvec(xref, y, faithful=False)
    for k in (lpp.KERNEL_TABLE, lpp.KERNEL_TILED):
        x = x0.copy()
        e.matrixVectorProduct(x, y, kernel=k)
        assert relerr(x, xref) <= 1e-13, k
    # U = 0: ground state energy = sum of the lowest single-particle levels of each species
    e.close()
    levels = np.sort(np.linalg.eigvalsh(geo.chain(18, -1.0, False)))
    eref = levels[:9].sum() + levels[:2].sum()
    e = cases.make_engine(lpp, cases.hubbard_chain(18, 9, 2, U=0.0))
    en, _, a, b = lpp.LanczosSolver(e, lpp.ParametersForSolver(steps=300, eps=1e-12)).computeOneState(None, want_vector=False)
    assert abs(en - eref) <= 1e-9 * abs(eref)
    e.close()


# ---------------------------------------------------------------------------------------------------------------------
# against the golden fixtures of the reference's own model code (tests/golden/*.npz, tools/make_golden.py)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(cases.SMALL_CASES))
def test_engine_matches_reference_fixture(lpp, name):
    from tests import golden_util as gu
    case = cases.SMALL_CASES[name]
    g = gu.load(name, case)
    # feas_u3_all_pairs=0: the literal on-the-fly loop of FeBasedSc.h:85-88, which is what the fixture's x_otf holds
    e = cases.make_engine(lpp, dict(case, feas_u3_all_pairs=0))
    assert e.rows() == int(g["rows"])
    assert np.array_equal(e.basis(0), g["up_words"])                      # bit-exact, built on device
    assert np.array_equal(e.perfectIndex(0, g["up_words"]), np.arange(len(g["up_words"]), dtype=np.uint64))
    if case["model"] != cases.HEISENBERG:
        assert np.array_equal(e.basis(1), g["dn_words"])
        assert np.array_equal(e.perfectIndex(1, g["dn_words"]), np.arange(len(g["dn_words"]), dtype=np.uint64))
    rp, ci, v = e.setupHamiltonian()
    gu.check_crs(g, rp, ci, v)                                            # rowptr / colind bit-exact, values <= 1e-14
    y = geo.splitmix64_vector(e.rows(), gu.Y_SEED)
    for k in kernels_for(lpp, case):
        x = np.zeros(e.rows())
        e.matrixVectorProduct(x, y, kernel=k)
        ref = g["x_stored"] if (k == lpp.KERNEL_STORED or "x_otf" not in g.files) else g["x_otf"]
        assert relerr(x, ref) <= 1e-13, (name, k)
    e.close()


_ALL = dict(cases.SMALL_CASES, **cases.TJ_CASES)


@pytest.mark.parametrize("name", sorted(_ALL))
def test_apply_op_matches_reference_fixture(lpp, name):
    """accModifiedState_ on the device against the reference's getBraIndex / doSignGf / doSignSpSm results: c, cdagger on both
    spins (HubbardOneBand, FeAsBasedSc per orbital, Tj1Orbital), sz / splus / sminus (HubbardOneBand, Heisenberg), n (Heisenberg)."""
    from tests import golden_util as gu
    case = _ALL[name]
    g = gu.load(name, case)
    e = cases.make_engine(lpp, case)
    e.set_groundstate(geo.splitmix64_vector(e.rows(), gu.SRC_SEED))
    seen = 0
    for rec in gu.ops(g):
        same = rec["op"] in (2, 4)                                        # sz / n stay in the sector: the handle is its own destination
        dst = e if same else e.sector(rec["nup"], rec["ndown"])
        e.apply_op(dst, rec["op"], rec["site"], rec["spin"], 1.0, accumulate=False, orb=rec["orb"])
        assert np.array_equal(dst.get_vector(1), g[rec["key"]]), rec      # +-source elements: exact
        if not same:
            dst.close()
        seen += 1
    assert seen >= (4 if name in ("tj6_full", "tj5_no_dn", "hub_empty_dn") else 8)
    e.close()


# ---------------------------------------------------------------------------------------------------------------------
# Tj1Orbital (TjMultiOrb.h with Orbitals=1): basis built on the device in closed form, stored CRS (the reference's path for
# this model) and the generic on-the-fly kernel
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(cases.TJ_CASES))
def test_tj_matches_reference_fixture(lpp, oracle, name):
    from tests import golden_util as gu
    case = cases.TJ_CASES[name]
    g = gu.load(name, case)
    e = cases.make_engine(lpp, case)
    assert e.rows() == int(g["rows"])
    up, dn = e.row_words()
    assert np.array_equal(up, g["up_words"]) and np.array_equal(dn, g["dn_words"])          # bit-exact order
    assert np.array_equal(e.perfectIndexPairs(up, dn), np.arange(e.rows(), dtype=np.uint64))
    rp, ci, v = e.setupHamiltonian()
    gu.check_crs(g, rp, ci, v)                                                               # rowptr / colind bit-exact
    y = geo.splitmix64_vector(e.rows(), gu.Y_SEED)
    for k in (lpp.KERNEL_GENERIC, lpp.KERNEL_STORED, lpp.KERNEL_AUTO):
        x = np.zeros(e.rows())
        e.matrixVectorProduct(x, y, kernel=k)
        assert relerr(x, g["x_stored"]) <= 1e-13, (name, k)
    e.close()


@pytest.mark.parametrize("name", ["tj8_V", "tj_3x3", "tj7_rand"])
def test_tj_lanczos_energy(lpp, oracle, name):
    case = cases.TJ_CASES[name]
    o = cases.make_oracle(oracle, case)
    init = geo.splitmix64_vector(o.rows(), 1234)
    e0, _, a0, b0 = o.ground_state(init, 200, 1e-12, 4)
    for k in (lpp.KERNEL_GENERIC, lpp.KERNEL_STORED):
        e = cases.make_engine(lpp, case, kernel=k)
        e1, _, a1, b1 = lpp.LanczosSolver(e, lpp.ParametersForSolver(steps=200, eps=1e-12, minsteps=4)).computeOneState(
            init, want_vector=False)
        assert abs(e1 - e0) <= 1e-10 * max(1.0, abs(e0)), (name, k)
        n = min(len(a0), len(a1), 20)
        assert relerr(a1[:n], a0[:n]) <= 1e-10 and relerr(b1[:n], b0[:n]) <= 1e-10
        e.close()


def test_tj_medium_size_vs_oracle(lpp, oracle):
    """4x4 lattice with two holes (7 up, 7 down): dim 16!/(7! 7! 2!) = 411 840; product's closed-form rank vs the oracle's search."""
    case = cases.tj_square(4, 4, 7, 7)
    o = cases.make_oracle(oracle, case)
    e = cases.make_engine(lpp, case)
    assert e.rows() == o.rows() == 411840
    up, dn = e.row_words()
    assert np.array_equal(up, o.row_words(0)) and np.array_equal(dn, o.row_words(1))
    y = geo.splitmix64_vector(o.rows(), 42)
    xref = np.zeros(o.rows())
    o.matvec(xref, y, faithful=False)
    for k in (lpp.KERNEL_GENERIC, lpp.KERNEL_STORED):
        x = np.zeros(o.rows())
        e.matrixVectorProduct(x, y, kernel=k)
        assert relerr(x, xref) <= 1e-13, k
    e.close()


# ---------------------------------------------------------------------------------------------------------------------
# <prefix>Options=reortho: every Lanczos vector saved on the device, full reorthogonalisation after x -= a y
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["c1_hub8", "feas4", "heis12", "tj8_V"])
def test_reortho_matches_oracle_and_resolves_the_spectrum(lpp, oracle, name):
    case = (cases.TJ_CASES if name.startswith("tj") else cases.SMALL_CASES)[name]
    o = cases.make_oracle(oracle, case)
    n = o.rows()
    init = geo.splitmix64_vector(n, 99)
    steps = min(n, 120)
    a0, b0 = o.decomposition_reortho(init, steps=steps, eps=0.0)
    e = cases.make_engine(lpp, case)
    a1, b1, _ = lpp.LanczosSolver(e, lpp.ParametersForSolver(steps=steps, eps=0.0, options="reortho")).decomposition(init)
    assert len(a1) == len(a0) == steps
    # the map (H, v0) -> (a, b) is itself ill conditioned at late steps for clustered spectra (heis12): the first 40 steps are
    # compared coefficient by coefficient, the whole run through its converged Ritz values
    m = 40
    assert relerr(a1[:m], a0[:m]) <= 1e-10 and relerr(b1[:m], b0[:m]) <= 1e-10
    assert np.abs(lpp.tridiag_eig(a1, b1)[:5] - oracle.tridiag_eig(a0, b0)[:5]).max() <= 1e-9
    # without it the same run drifts once the lowest Ritz values converge; with it the Ritz values are eigenvalues of H, once each
    rp, ci, v = o.crs()
    import scipy.sparse as sp
    H = sp.csr_matrix((v, ci, rp), shape=(n, n)).toarray()
    exact = np.linalg.eigvalsh(H) if n <= 1600 else None
    ritz = lpp.tridiag_eig(a1, b1)
    if exact is not None:
        assert abs(ritz[0] - exact[0]) <= 1e-10 * max(1.0, abs(exact[0]))
        conv = ritz[:5]
        for r in conv:                                                # the lowest Ritz values sit on distinct eigenvalues: no ghosts
            assert np.abs(exact - r).min() <= 1e-8
        assert np.all(np.diff(conv) > 1e-9) or len(np.unique(np.round(exact[:8], 9))) < 8
    e.close()


def test_reortho_full_space_reproduces_all_eigenvalues(lpp, oracle):
    """steps = dim on a small sector: only a recurrence that stays orthogonal yields the complete spectrum."""
    case = cases.hubbard_chain(5, 2, 2, U=3.0)                        # dim 100
    o = cases.make_oracle(oracle, case)
    n = o.rows()
    rp, ci, v = o.crs()
    import scipy.sparse as sp
    exact = np.linalg.eigvalsh(sp.csr_matrix((v, ci, rp), shape=(n, n)).toarray())
    e = cases.make_engine(lpp, case)
    init = geo.splitmix64_vector(n, 3)
    a, b, _ = lpp.LanczosSolver(e, lpp.ParametersForSolver(steps=n, eps=0.0, options="reortho")).decomposition(init)
    # the Krylov space of a random vector misses symmetry-degenerate copies: compare as sets of distinct values
    k = int(np.argmax(b < 1e-8)) + 1 if np.any(b < 1e-8) else len(a)     # the Krylov space may close before dim steps
    ritz = lpp.tridiag_eig(a[:k], b[:k])
    distinct = np.unique(np.round(exact, 8))
    for r in ritz:
        assert np.abs(distinct - r).min() <= 1e-8
    assert len(ritz) >= 0.5 * len(distinct)
    e.close()


@pytest.mark.parametrize("name,orbs,steps", [("feas3", (0, 1), 8), ("feas4", (1, 1), 30), ("tj8_V", (0, 0), 30), ("tj6_pbc", (0, 0), 10)])
def test_continued_fraction_other_models(lpp, oracle, name, orbs, steps):
    """spectralFunction for FeAsBasedSc (orbital pair) and Tj1Orbital: modified state, weights, tridiagonal coefficients and
    G(omega) against the oracle pipeline (apply_op -> decomposition -> continued fraction), same ground-state vector on both sides."""
    case = (cases.TJ_CASES if name.startswith("tj") else cases.SMALL_CASES)[name]
    o = cases.make_oracle(oracle, case)
    init = geo.splitmix64_vector(o.rows(), 77)
    e0, z0, _, _ = o.ground_state(init, 300, 1e-13, 4)
    eng = cases.make_engine(lpp, case, kernel=lpp.KERNEL_AUTO)
    # `steps` stays well below the dimension of the smallest new sector: once the Krylov space closes the late coefficients are noise
    en = lpp.Engine(eng, {"LanczosSteps": 300, "LanczosEps": 1e-13, "SpectralSteps": steps, "SpectralEps": 0.0}, init=init)
    assert abs(en.energies(0) - e0) < 1e-10
    eng.set_groundstate(z0)
    en.energy = e0
    omega = np.linspace(-6, 6, 49)
    isite, jsite = 0, case["nsite"] - 1
    for spin in (0, 1):
        cfs = en.spectralFunction(lpp.OP_C, isite, jsite, spin=spin, orbs=orbs)
        assert len(cfs) >= 2
        for typ, cf in cfs:
            lop = oracle.OP_C if (typ & 1) else oracle.OP_CDAGGER
            dn = -1 if lop == oracle.OP_C else 1
            od = cases.make_oracle(oracle, dict(case, nup=case["nup"] + (dn if spin == 0 else 0),
                                                ndown=case["ndown"] + (dn if spin == 1 else 0)))
            phi = np.zeros(od.rows())
            o.apply_op(od, lop, isite, spin, 1.0, z0, phi, orb=orbs[0])
            o.apply_op(od, lop, jsite, spin, -1.0 if typ > 1 else 1.0, z0, phi, orb=orbs[1])
            a, b = od.decomposition(phi, steps=steps, eps=0.0)
            weight = (phi @ phi) * (-1.0 if typ > 1 else 1.0) * 0.5
            assert abs(cf.weight - weight) <= 1e-12 * max(1.0, abs(weight))
            n = min(len(a), 8)
            assert relerr(cf.a[:n], a[:n]) <= 1e-10 and relerr(cf.b[:n], b[:n]) <= 1e-10
            s = -1 if (typ & 1) else 1
            gref = oracle.cf_eval(a, b, e0, weight, -s, omega, 0.1)
            assert np.abs(cf(omega, 0.1) - gref).max() <= 1e-8 * max(1.0, np.abs(gref).max()), (name, spin, typ)
    eng.close()


@pytest.mark.parametrize("name,orbs", [("c1_hub8", (0, 0)), ("hub_rand7", (0, 0)), ("feas4", (0, 1)), ("tj8_V", (0, 0))])
def test_two_point_matches_oracle(lpp, oracle, name, orbs):
    """Engine::twoPoint (Engine.h:262-331) on the device: <cdagger_j c_i> for both spins and <n_j n_i> (Hubbard) against the
    oracle; the trace of the one-body density matrix is the electron number of that species and orbital."""
    case = (cases.TJ_CASES if name.startswith("tj") else cases.SMALL_CASES)[name]
    o = cases.make_oracle(oracle, case)
    init = geo.splitmix64_vector(o.rows(), 5)
    e0, z0, _, _ = o.ground_state(init, 300, 1e-13, 4)
    eng = cases.make_engine(lpp, case)
    en = lpp.Engine(eng, {"LanczosSteps": 300, "LanczosEps": 1e-13}, init=init)
    eng.set_groundstate(z0)
    for spin in (0, 1):
        od = cases.make_oracle(oracle, dict(case, nup=case["nup"] - (1 if spin == 0 else 0), ndown=case["ndown"] - (1 if spin == 1 else 0)))
        ref = oracle.two_point(o, od, oracle.OP_C, spin, z0, orbs)
        got = en.twoPoint(lpp.OP_C, spin=spin, orbs=orbs)
        assert np.abs(got - ref).max() <= 1e-12
        if orbs[0] == orbs[1] and case["orbitals"] == 1:
            assert abs(np.trace(got) - (case["nup"] if spin == 0 else case["ndown"])) <= 1e-10
    if case["model"] == cases.HUBBARD:
        ref = oracle.two_point(o, o, oracle.OP_N, 0, z0)
        assert np.abs(en.twoPoint(lpp.OP_N, spin=0) - ref).max() <= 1e-12
    eng.close()


@pytest.mark.parametrize("name", ["hub_rand7", "feas4", "tj7_rand"])
def test_states_below_excited_states(lpp, oracle, name):
    """computeAllStatesBelow with excited > 0 (Engine.h:601-657): the lowest four Ritz pairs with Options=reortho against the
    oracle and, for the energies, against a dense diagonalisation of the stored Hamiltonian (distinct eigenvalues)."""
    case = (cases.TJ_CASES if name.startswith("tj") else cases.SMALL_CASES)[name]
    o = cases.make_oracle(oracle, case)
    n = o.rows()
    init = geo.splitmix64_vector(n, 31)
    e0, z0, _ = o.states_below(init, 4, steps=300, eps=1e-12)
    eng = cases.make_engine(lpp, case)
    solver = lpp.LanczosSolver(eng, lpp.ParametersForSolver(steps=300, eps=1e-12, options="reortho"))
    e1, z1, steps = solver.computeAllStatesBelow(init, 4)
    assert np.abs(e1 - e0).max() <= 1e-9 * max(1.0, np.abs(e0).max())
    import scipy.sparse as sp
    rp, ci, v = o.crs()
    H = sp.csr_matrix((v, ci, rp), shape=(n, n))
    distinct = np.unique(np.round(np.linalg.eigvalsh(H.toarray()), 8))
    assert np.abs(e1 - distinct[:4]).max() <= 1e-7
    for k in range(4):
        assert abs(np.linalg.norm(z1[k]) - 1.0) <= 1e-8
        r = H @ z1[k] - e1[k] * z1[k]
        assert np.linalg.norm(r) <= 1e-5                                  # Ritz residual of a converged pair
        for q in range(k):
            assert abs(z1[k] @ z1[q]) <= 1e-7
    assert np.abs(eng.get_vector(0) - z1[0]).max() == 0.0                  # state 0 stays in the handle
    eng.close()


@pytest.mark.parametrize("name,op", [("c1_hub8", "sz"), ("c1_hub8", "splus"), ("heis12", "sz"), ("heis12", "sminus")])
def test_spin_operator_continued_fractions(lpp, oracle, name, op):
    """spectralFunction with the spin operators (the S(q, omega) building blocks): type loop with transposeConjugate, the
    non-fermionic sign s2 *= s (Engine.h:482), sz without a new basis; against the oracle pipeline."""
    case = cases.SMALL_CASES[name]
    OP = {"sz": lpp.OP_SZ, "splus": lpp.OP_SPLUS, "sminus": lpp.OP_SMINUS}[op]
    conj = {lpp.OP_SZ: lpp.OP_SZ, lpp.OP_SPLUS: lpp.OP_SMINUS, lpp.OP_SMINUS: lpp.OP_SPLUS}
    o = cases.make_oracle(oracle, case)
    init = geo.splitmix64_vector(o.rows(), 13)
    e0, z0, _, _ = o.ground_state(init, 300, 1e-13, 4)
    eng = cases.make_engine(lpp, case)
    en = lpp.Engine(eng, {"LanczosSteps": 300, "LanczosEps": 1e-13, "SpectralSteps": 25, "SpectralEps": 0.0}, init=init)
    eng.set_groundstate(z0)
    en.energy = e0
    omega = np.linspace(-5, 5, 41)
    isite, jsite = 1, 4
    cfs = en.spectralFunction(OP, isite, jsite, spin=0)
    assert len(cfs) == 4
    for typ, cf in cfs:
        lop = OP if (typ & 1) else conj[OP]
        nup, ndown = case["nup"], case["ndown"]
        if lop in (lpp.OP_SPLUS, lpp.OP_SMINUS):
            c = 1 if lop == lpp.OP_SPLUS else -1
            nup += c
            if case["model"] == cases.HUBBARD:
                ndown -= c
        od = o if lop == lpp.OP_SZ else cases.make_oracle(oracle, dict(case, nup=nup, ndown=ndown))
        phi = np.zeros(od.rows())
        o.apply_op(od, lop, isite, 0, 1.0, z0, phi)
        o.apply_op(od, lop, jsite, 0, -1.0 if typ > 1 else 1.0, z0, phi)
        a, b = od.decomposition(phi, steps=25, eps=0.0)
        s = -1 if (typ & 1) else 1
        weight = (phi @ phi) * (-1.0 if typ > 1 else 1.0) * s * 0.5
        assert abs(cf.weight - weight) <= 1e-12 * max(1.0, abs(weight)), (name, op, typ)
        n = min(len(a), 8)
        assert relerr(cf.a[:n], a[:n]) <= 1e-10 and relerr(cf.b[:n], b[:n]) <= 1e-10
        gref = oracle.cf_eval(a, b, e0, weight, -s, omega, 0.1)
        assert np.abs(cf(omega, 0.1) - gref).max() <= 1e-8 * max(1.0, np.abs(gref).max()), (name, op, typ)
    eng.close()


@pytest.mark.parametrize("name", ["c1_hub8", "hub_rand7", "feas4", "tj8_V"])
def test_measure_matches_reference_fixture(lpp, name):
    """Engine::measure on the device (lpp_measure): <v| prod ops |v> for the fixture's vector v against the values the reference's
    ModelBase::rahulMethod gave; physical checks on a ground state (densities sum to the electron numbers)."""
    from tests import golden_util as gu
    case = (cases.TJ_CASES if name.startswith("tj") else cases.SMALL_CASES)[name]
    g = gu.load(name, case)
    eng = cases.make_engine(lpp, case)
    eng.set_groundstate(geo.splitmix64_vector(eng.rows(), gu.SRC_SEED))
    en = lpp.Engine.__new__(lpp.Engine)
    en.mat, en.io = eng, {}
    label = {0: "identity", 1: "n", 2: "sz", 3: "c"}
    for key, rec in gu.measures(g).items():
        got = en.measure([(label[o[0]], o[1], o[2], o[3]) for o in rec["ops"]])
        assert abs(got - rec["value"]) <= 1e-12 * max(1.0, abs(rec["value"])), (name, key)
    en2 = lpp.Engine(eng, {"LanczosSteps": 300, "LanczosEps": 1e-13}, init=geo.splitmix64_vector(eng.rows(), 1))
    nb = case["nsite"] * case["orbitals"]
    assert abs(sum(en2.measure([("n", 0, s)]) for s in range(nb)) - case["nup"]) <= 1e-9
    assert abs(sum(en2.measure([("n", 1, s)]) for s in range(nb)) - case["ndown"]) <= 1e-9
    eng.close()


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs at their NAMED size against the oracle
# ---------------------------------------------------------------------------------------------------------------------
def _fullsize_case(name):
    import bench
    return bench.workload(name)[0]


def _pins():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fullsize_pins.json")))


@pytest.mark.parametrize("name", ["c2", "c3", "c4"])
def test_full_size_sampled_parity(lpp, oracle, name):
    """x += H y with the AUTO kernel at the named size of configs 2, 3 and 4, compared with the oracle
    (HubbardHelper.h:105-134, FeBasedSc.h:66-105, Heisenberg.h:80-114 restated) on three windows of 200 000 rows: the first
    rows, the middle and the LAST rows (the last panel / chunk edges of the big-size code paths).  Config 2 is compared whole.
    Bar: 1e-13 relative to the largest element of the window."""
    case = _fullsize_case(name)
    o = cases.make_oracle(oracle, case, fast_rank=1)
    e = cases.make_engine(lpp, case)
    n = e.rows()
    assert n == o.rows() == {"c2": 2704156, "c3": 165636900, "c4": 64128064}[name]
    y = geo.splitmix64_vector(n, 42)
    x0 = geo.splitmix64_vector(n, 7)
    x = x0.copy()
    e.matrixVectorProduct(x, y)                                            # AUTO kernel, accumulate semantics
    w = n if name == "c2" else 200000
    for r0 in sorted({0, (n - w) // 2, n - w}):
        xr = x0[r0:r0 + w].copy()
        o.matvec_range(xr, y, r0, r0 + w, faithful=False)
        assert relerr(x[r0:r0 + w], xr) <= 1e-13, (name, r0)
    e.close()


@pytest.mark.parametrize("name", ["c2", "c3", "c4"])
def test_full_size_pins(lpp, name):
    """The first Lanczos coefficients of the seeded initial vector at the named size against the committed oracle pins
    (tests/golden/fullsize_pins.json, tools/make_fullsize_pins.py: one full-size oracle mat-vec): alpha_0, beta_0 to 1e-10
    relative (north star), and for configs 3 / 4 the per-chunk sums of H v0 over 64 row chunks, so that EVERY panel of the
    full-size mat-vec is covered by the oracle, not only the sampled windows."""
    pin = _pins()[name]
    case = _fullsize_case(name)
    e = cases.make_engine(lpp, case)
    n = e.rows()
    assert n == pin["rows"]
    v = geo.splitmix64_vector(n, pin["seed"])
    solver = lpp.LanczosSolver(e, lpp.ParametersForSolver(steps=10 if name == "c2" else 2, eps=0.0))
    a, b, _ = solver.decomposition(v)
    if name == "c2":
        k = len(pin["alpha"])
        assert relerr(np.asarray(a[:k]), np.asarray(pin["alpha"])) <= 1e-10
        assert relerr(np.asarray(b[:k - 1]), np.asarray(pin["beta"][:k - 1])) <= 1e-10
    else:
        assert abs(a[0] - pin["alpha0"]) <= 1e-10 * abs(pin["alpha0"])
        assert abs(b[0] - pin["beta0"]) <= 1e-10 * abs(pin["beta0"])
        v /= np.sqrt(np.dot(v, v))
        hv = np.zeros(n)
        e.matrixVectorProduct(hv, v)
        nc = pin["chunks"]
        for i in range(nc):
            r0, r1 = (n * i) // nc, (n * (i + 1)) // nc
            c = hv[r0:r1]
            scale = np.sqrt(pin["chunk_sumsq"][i] * (r1 - r0))
            assert abs(c.sum() - pin["chunk_sum"][i]) <= 1e-11 * scale, (name, i)
            assert abs(np.dot(c, c) - pin["chunk_sumsq"][i]) <= 1e-11 * pin["chunk_sumsq"][i], (name, i)
    e.close()


def test_multi_gpu_matches_single(lpp):
    """tools/check_multi_gpu.py (row-sharded engine on every visible GPU against the unsharded engine: energy, alpha/beta,
    x += H y) as a test; skipped on a one-GPU box."""
    import os
    import subprocess
    import sys
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # two ranks here (every exchange path is exercised); `torchrun --nproc-per-node 8 tools/check_multi_gpu.py` is the 8-rank form
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(root, "tools", "check_multi_gpu.py")]
    r = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ALL OK" in r.stdout
    # the same cases through the opt-in pipelined recurrence (LPP_PIPELINE: unpack of step j beside the up sweep of step j+1)
    env = dict(os.environ, LPP_PIPELINE="3", LPP_CHECK_FULL="0")
    cmd[cmd.index("29731")] = "29733"
    r = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=1500, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ALL OK" in r.stdout


def test_block_down_sweep_cases(lpp, oracle, monkeypatch):
    """The two-pass block down sweep (k_dblock, default for HubbardOneBand with one U) against the oracle on cases that exercise
    its corners: site potentials (the per-state dv2 path), a column count that is not a multiple of 16 (last panel partly
    empty), odd block sizes, an open chain (different F1 / F2 choice), and the streaming sweep (LPP_DBLOCK=0) as a second
    opinion on the same inputs."""
    V12 = np.linspace(-0.4, 0.6, 12)
    todo = (cases.hubbard_square(4, 3, 6, 6), cases.hubbard_chain(12, 6, 5, periodic=True, V=V12),
            cases.hubbard_chain(12, 5, 6, V=V12), cases.hubbard_chain(14, 3, 7), cases.hubbard_square(4, 3, 5, 7))
    for case in todo:
        o = cases.make_oracle(oracle, case, fast_rank=1)
        n = o.rows()
        y = geo.splitmix64_vector(n, 42)
        x0 = geo.splitmix64_vector(n, 7)
        xref = x0.copy()
        o.matvec(xref, y, faithful=False)
        for off in ("1", "0"):
            monkeypatch.setenv("LPP_DBLOCK", off)
            e = cases.make_engine(lpp, case)
            x = x0.copy()
            e.matrixVectorProduct(x, y, kernel=lpp.KERNEL_TILED)
            assert relerr(x, xref) <= 1e-13, (case["nsite"], case["nup"], case["ndown"], off)
            a, b, _ = lpp.LanczosSolver(e, lpp.ParametersForSolver(steps=12, eps=0.0)).decomposition(y)
            a0, b0 = o.decomposition(y, steps=12, eps=0.0)
            assert relerr(a, a0) <= 1e-10 and relerr(b[:-1], b0[:-1]) <= 1e-10
            e.close()


@pytest.mark.parametrize("layout,passes", [("1", "0"), ("0", "3"), ("1", "3")])
def test_block_down_sweep_layouts(lpp, oracle, monkeypatch, capfd, layout, passes):
    """The code paths of k_dblock that config 3 does not take by default: one CTA of 1024 threads per SM (LPP_DBLOCK_LAYOUT=1, what a
    lattice gets whose blocks do not fit twice into an SM) and three passes over three disjoint site sets (LPP_DBLOCK_PASSES=3,
    what a lattice gets that has no two separated site sets), against the oracle: mat-vec with the old x (beta != 0), the fused
    Lanczos recurrence, and the last panel of a column count that is not a multiple of 16."""
    monkeypatch.setenv("LPP_DBLOCK_LAYOUT", layout)
    monkeypatch.setenv("LPP_DBLOCK_PASSES", passes)
    monkeypatch.setenv("LPP_VERBOSE", "1")                    # the plan is described on stderr: the forced path must be the one that ran
    want = ("3 passes" if passes == "3" else "2 passes", "1 CTA(s) of 1024 threads" if layout == "1" else "2 CTA(s) of 512 threads")
    V12 = np.linspace(-0.4, 0.6, 12)
    for case in (cases.hubbard_square(4, 3, 6, 6), cases.hubbard_chain(12, 5, 6, V=V12), cases.hubbard_square(4, 3, 5, 7)):
        capfd.readouterr()
        o = cases.make_oracle(oracle, case, fast_rank=1)
        n = o.rows()
        y = geo.splitmix64_vector(n, 42)
        x0 = geo.splitmix64_vector(n, 7)
        xref = x0.copy()
        o.matvec(xref, y, faithful=False)
        e = cases.make_engine(lpp, case)
        x = x0.copy()
        e.matrixVectorProduct(x, y, kernel=lpp.KERNEL_TILED)
        said = [l for l in capfd.readouterr().err.splitlines() if "block down sweep:" in l]
        assert said and all(w in said[-1] for w in want), said
        assert relerr(x, xref) <= 1e-13, (case["nsite"], case["nup"], case["ndown"], layout, passes)
        a, b, _ = lpp.LanczosSolver(e, lpp.ParametersForSolver(steps=12, eps=0.0)).decomposition(y)
        a0, b0 = o.decomposition(y, steps=12, eps=0.0)
        assert relerr(a, a0) <= 1e-10 and relerr(b[:-1], b0[:-1]) <= 1e-10
        e.close()


def test_many_point_matches_oracle_chain(lpp, oracle):
    """Engine::manyPoint (Engine.h:341-389): <gs| O_n ... O_1 |gs> through the chain of sectors, on the device, against the
    oracle's accModifiedState_ restatement applied step by step (same ground-state vector on both sides); also against
    twoPoint for the two-operator string and against the reference's answer 0 for strings that do not return."""
    case = cases.hubbard_chain(6, 3, 3, periodic=True, V=[0.3, -0.2, 0.1, 0.0, 0.5, -0.4])
    o = cases.make_oracle(oracle, case)
    init = geo.splitmix64_vector(o.rows(), 1234)
    e0, z0, _, _ = o.ground_state(init, 300, 1e-13, 4)
    eng = cases.make_engine(lpp, case)
    en = lpp.Engine(eng, {"LanczosSteps": 300, "LanczosEps": 1e-13}, init=init)
    eng.set_groundstate(z0)
    OPS = {lpp.OP_C: oracle.OP_C, lpp.OP_CDAGGER: oracle.OP_CDAGGER, lpp.OP_N: oracle.OP_N, lpp.OP_SZ: oracle.OP_SZ}

    def chain_oracle(sites, what, spins):
        cur, vec, nu, nd = o, z0, 3, 3
        for s, w, sp in zip(sites, what, spins):
            if w in (lpp.OP_C, lpp.OP_CDAGGER):
                d = -1 if w == lpp.OP_C else 1
                nu, nd = (nu + d, nd) if sp == 0 else (nu, nd + d)
            nxt = cases.make_oracle(oracle, dict(case, nup=nu, ndown=nd))
            out = np.zeros(nxt.rows())
            cur.apply_op(nxt, OPS[w], s, sp, 1.0, vec, out)
            cur, vec = nxt, out
        return float(z0 @ vec) if (nu, nd) == (3, 3) else 0.0

    strings = [([1, 4], [lpp.OP_C, lpp.OP_CDAGGER], [0, 0]),
               ([2, 2], [lpp.OP_C, lpp.OP_CDAGGER], [1, 1]),
               ([0, 3, 5, 2], [lpp.OP_C, lpp.OP_C, lpp.OP_CDAGGER, lpp.OP_CDAGGER], [0, 1, 1, 0]),
               ([1, 1, 4], [lpp.OP_N, lpp.OP_C, lpp.OP_CDAGGER], [0, 0, 0]),
               ([3, 2], [lpp.OP_N, lpp.OP_N], [0, 1]),
               ([1, 4], [lpp.OP_C, lpp.OP_C], [0, 0])]          # does not return to the sector: 0
    for sites, what, spins in strings:
        got = en.manyPoint(sites, what, spins)
        ref = chain_oracle(sites, what, spins)
        assert abs(got - ref) <= 1e-12 * max(1.0, abs(ref)), (sites, what, spins, got, ref)
    tp = en.twoPoint(lpp.OP_C, spin=0)                         # result(i, j) = <c_j gs | c_i gs> = <gs| cdagger_j c_i |gs>
    assert abs(en.manyPoint([1, 4], [lpp.OP_C, lpp.OP_CDAGGER], [0, 0]) - tp[1, 4]) <= 1e-12
    eng.close()
