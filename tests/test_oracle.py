"""Independent pins for the CPU oracle: energies and spectra (the PsimagLite solver parts have no reference fixture -- PARITY
UNPINNED there; the model code is pinned by tests/test_golden.py and tests/test_reference_pin.py).

Pins used instead: analytic energies, dense/sparse eigensolvers on the exported CRS, and the survey-time
independent restatement values of SURVEY App. E / BASELINE.md §2.
"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from lanczosplusplus_b200 import geometry as geo
from tests import cases


def crs_matrix(m):
    rp, ci, v = m.crs()
    n = m.rows()
    return sp.csr_matrix((v, ci, rp), shape=(n, n))


def lowest(S):
    if S.shape[0] <= 1500:
        return np.linalg.eigvalsh(S.toarray())[0]
    return sla.eigsh(S, k=1, which="SA", tol=1e-13)[0][0]


PINS = [
    ("input0", -2.0 * np.sqrt(5.0), 1e-12),          # TestSuite/inputs/input0.inp: free fermions, analytic
    ("hub2", 2.0 - np.sqrt(8.0), 1e-13),             # U/2 - sqrt(U^2/4 + 4t^2)
    ("c1_hub8", -4.235806999130, 2e-12),             # SURVEY App. E (independent restatement + eigvalsh)
    ("feas2", -0.830292210803, 2e-12),
    ("feas3", -1.238806375931, 2e-12),
    ("feas4", -1.957490136573, 2e-12),
    ("heis4", -2.0, 1e-13),                          # analytic
    ("heis12", -5.387390917445, 2e-12),
]


@pytest.mark.parametrize("name,ref,tol", PINS)
def test_energy_pins(oracle, name, ref, tol):
    m = cases.make_oracle(oracle, cases.SMALL_CASES[name])
    S = crs_matrix(m)
    assert abs(S - S.T).max() == 0.0
    assert abs(lowest(S) - ref) < tol


def test_free_fermion_8chain(oracle):
    m = cases.make_oracle(oracle, cases.hubbard_chain(8, 4, 4, U=0.0))
    levels = np.sort(np.linalg.eigvalsh(geo.chain(8, -1.0)))
    assert abs(lowest(crs_matrix(m)) - 2 * levels[:4].sum()) < 1e-12   # -9.517540966287


def test_heis16_and_feas6_survey_values(oracle):
    m = cases.make_oracle(oracle, cases.heisenberg_ring(16, 8))
    assert abs(lowest(crs_matrix(m)) - (-7.142296360617)) < 2e-11
    m = cases.make_oracle(oracle, cases.feas_chain(6, 3, 3))           # input100.inp sector, dim 48400
    S = crs_matrix(m)
    assert m.rows() == 48400 and S.nnz == 493000
    assert abs(lowest(S) - (-3.099464014219)) < 2e-11


def test_c1_structure(oracle):
    m = cases.make_oracle(oracle, cases.SMALL_CASES["c1_hub8"])
    rp, ci, v = m.crs()
    assert m.rows() == 4900 and len(ci) == 44100
    # diagonal always present, columns strictly ascending within a row
    for r in range(0, 4900, 97):
        cols = ci[rp[r]:rp[r + 1]]
        assert r in cols and np.all(np.diff(cols) > 0)


def test_onespin_basis_and_rank(oracle):
    for nsite, npart in ((4, 2), (8, 4), (16, 8), (18, 9), (7, 0), (5, 5)):
        b = oracle.onespin_basis(nsite, npart)
        ref = np.array([w for w in range(1 << nsite) if bin(w).count("1") == npart], dtype=np.uint64)
        assert np.array_equal(b, ref)
        for i in (0, len(b) // 3, len(b) - 1):
            assert oracle.onespin_rank(nsite, b[i]) == i


def test_feas_basis_order(oracle):
    m = cases.make_oracle(oracle, cases.feas_cluster(2, 4, 6, 6), fast_rank=1)
    b = m.basis(0)
    assert len(b) == 8008 and len(set(b.tolist())) == 8008
    assert not np.all(np.diff(b.astype(np.int64)) > 0)   # not numerically sorted
    # first partition block is (6,0): all six electrons in orbital 0 => only even bit positions occupied
    assert all((int(w) & 0xAAAA) == 0 for w in b[:28])


def test_faithful_and_tuned_matvec_agree(oracle):
    for name in ("c1_hub8", "feas4", "heis12", "hub_rand7"):
        case = cases.SMALL_CASES[name]
        m0 = cases.make_oracle(oracle, case, fast_rank=0)
        m1 = cases.make_oracle(oracle, case, fast_rank=1)
        S = crs_matrix(m0)
        y = geo.splitmix64_vector(m0.rows(), 42)
        x0, x1 = np.zeros_like(y), np.ones_like(y)
        m0.matvec(x0, y, faithful=True)
        m1.matvec(x1, y, faithful=False)
        ref = S @ y
        assert np.abs(x0 - ref).max() < 1e-13 * max(1.0, np.abs(ref).max())
        assert np.abs(x1 - 1.0 - ref).max() < 1e-13 * max(1.0, np.abs(ref).max())


def test_feas_otf_quirk_is_nonhermitian(oracle):
    """literal doTask (FeBasedSc.h:85-88) drops half of the pair-hopping terms."""
    m = cases.make_oracle(oracle, cases.feas_chain(3, 2, 1, u3_all_pairs=0))
    n = m.rows()
    B = np.zeros((n, n))
    for r in range(n):
        c, v = m.row(r, stored=False)
        np.add.at(B[r], c, v)
    assert abs(np.abs(B - B.T).max() - 0.4) < 1e-15


def test_lanczos_ground_state(oracle):
    m = cases.make_oracle(oracle, cases.SMALL_CASES["c1_hub8"])
    S = crs_matrix(m)
    init = geo.splitmix64_vector(m.rows(), 1234)
    e, z, a, b = m.ground_state(init, 200, 1e-12, 4)
    assert abs(e - (-4.235806999130)) < 1e-10
    assert abs(np.linalg.norm(z) - 1.0) < 1e-10
    assert np.linalg.norm(S @ z - e * z) < 1e-5
    # tridiagonal solver against numpy
    T = np.diag(a) + np.diag(b[:-1], 1) + np.diag(b[:-1], -1)
    assert np.abs(oracle.tridiag_eig(a, b) - np.linalg.eigvalsh(T)).max() < 1e-12


def test_continued_fraction_matches_resolvent(oracle):
    """G(z) from (a,b) equals <phi|(z - (H - Eg))^-1|phi> computed densely (full Krylov space, isign=+1)."""
    src = cases.make_oracle(oracle, cases.hubbard_chain(4, 2, 2, U=4.0))
    dst = cases.make_oracle(oracle, cases.hubbard_chain(4, 3, 2, U=4.0))
    Ssrc, Sdst = crs_matrix(src).toarray(), crs_matrix(dst).toarray()
    w, vec = np.linalg.eigh(Ssrc)
    gs, eg = vec[:, 0], w[0]
    phi = np.zeros(dst.rows())
    src.apply_op(dst, oracle.OP_CDAGGER, 1, 0, 1.0, gs, phi)
    a, b = dst.decomposition(phi, steps=dst.rows(), eps=0.0)
    weight = phi @ phi
    omega = np.linspace(-6, 6, 41)
    g = oracle.cf_eval(a, b, eg, weight, 1, omega, 0.1)
    wd, vd = np.linalg.eigh(Sdst)
    amp = (vd.T @ phi) ** 2
    ref = np.array([(amp / (o + 0.1j - (wd - eg))).sum() for o in omega])
    assert np.abs(g - ref).max() < 1e-8


def test_reortho_removes_ghost_ritz_values(oracle):
    """<prefix>Options=reortho in the oracle (one_step_reortho): the plain recurrence on the Heisenberg 12-ring repeats its lowest
    Ritz value after 120 steps, the reorthogonalised one lands on distinct eigenvalues of the stored Hamiltonian."""
    m = cases.make_oracle(oracle, cases.SMALL_CASES["heis12"])
    n = m.rows()
    init = geo.splitmix64_vector(n, 99)
    a0, b0 = m.decomposition(init, steps=120, eps=0.0)
    a1, b1 = m.decomposition_reortho(init, steps=120, eps=0.0)
    assert np.abs(a0[:20] - a1[:20]).max() < 1e-12 and np.abs(b0[:20] - b1[:20]).max() < 1e-12
    plain, ro = oracle.tridiag_eig(a0, b0), oracle.tridiag_eig(a1, b1)
    exact = np.unique(np.round(np.linalg.eigvalsh(crs_matrix(m).toarray()), 9))
    assert abs(plain[1] - plain[0]) < 1e-8                       # ghost copy of the ground state
    assert np.all(np.diff(ro[:4]) > 1e-6)
    for r in ro[:4]:
        assert np.abs(exact - r).min() < 1e-9


def test_states_below_and_two_point_in_the_oracle(oracle):
    """computeAllStatesBelow (excited states) and Engine::twoPoint restated in the oracle, against dense linear algebra."""
    case = cases.SMALL_CASES["hub_rand7"]
    m = cases.make_oracle(oracle, case)
    n = m.rows()
    H = crs_matrix(m).toarray()
    w, v = np.linalg.eigh(H)
    eigs, zs, _ = m.states_below(geo.splitmix64_vector(n, 31), 3, steps=300, eps=1e-12)
    assert np.abs(eigs - w[:3]).max() < 1e-9
    for k in range(3):
        assert abs(abs(zs[k] @ v[:, k]) - 1.0) < 1e-7
    # one-body density matrix of the ground state: Hermitian, trace = number of up electrons, eigenvalues in [0, 1]
    dst = cases.make_oracle(oracle, dict(case, nup=case["nup"] - 1))
    rho = oracle.two_point(m, dst, oracle.OP_C, 0, v[:, 0])
    assert np.abs(rho - rho.T).max() < 1e-12 and abs(np.trace(rho) - case["nup"]) < 1e-12
    occ = np.linalg.eigvalsh(rho)
    assert occ.min() > -1e-12 and occ.max() < 1 + 1e-12


def test_feas_two_spin_terms_conserve_up_xor_down(oracle):
    """The entry classes a multi-GPU layout for FeAsBasedSc has to serve (tests/feas_three_layouts.py): every
    off-diagonal entry of the oracle's matrix changes the down word only (column layout), the up word only (row layout) or both --
    and the ones that change both (FeBasedSc.h:376-432, spin flip and pair hop on one site) flip the same two bits in both words,
    so up XOR down is the same on both sides: a layout sharded by it keeps them local."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("feas_three_layouts", os.path.join(os.path.dirname(os.path.abspath(__file__)), "feas_three_layouts.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    for case in (cases.feas_cluster(2, 2, 3, 3), cases.feas_chain(3, 3, 2, inter_orbital=0.5)):
        o = cases.make_oracle(oracle, case, fast_rank=1)
        cl = tool.classify(o)
        two = cl["kind"] == 2
        assert two.sum() > 0 and (cl["kind"] == 0).sum() > 0 and (cl["kind"] == 1).sum() > 0
        assert np.array_equal(cl["dup"][two], cl["ddn"][two])
        w = cl["up"] ^ cl["dn"]
        assert np.array_equal(w[cl["rows"][two]], w[cl["cols"][two]])
        assert all(bin(int(d)).count("1") == 2 for d in np.unique(cl["dup"][two]))
        # the three classes together are the matrix: applying them one after the other reproduces the oracle's mat-vec
        y = geo.splitmix64_vector(cl["n"], 5)
        x = np.zeros(cl["n"])
        rowptr = cl["rowptr"]
        rows_all = np.repeat(np.arange(cl["n"]), np.diff(rowptr))
        diag = rows_all == cl["colind"]
        np.add.at(x, rows_all[diag], cl["allvals"][diag] * y[cl["colind"][diag]])
        for k in range(3):
            sel = cl["kind"] == k
            np.add.at(x, cl["rows"][sel], cl["vals"][sel] * y[cl["cols"][sel]])
        xref = np.zeros(cl["n"])
        o.matvec(xref, y, faithful=False)
        assert np.abs(x - xref).max() <= 1e-12 * max(1.0, np.abs(xref).max())
        # the mat-vec as 2 and 3 ranks would do it: no entry joins two ranks in the layout of its class, the shards are balanced
        for nranks in (2, 3):
            x3, moved = tool.three_layout_matvec(cl, len(o.basis(0)), nranks, y)
            assert np.abs(x3 - xref).max() <= 1e-12 * max(1.0, np.abs(xref).max())
            load = tool.owners(cl, len(o.basis(0)), nranks)[3]
            assert load.sum() == cl["n"] and load.max() - load.min() <= 0.1 * cl["n"] / nranks


def test_pipelined_recurrence_formula_matches_the_oracle(oracle):
    """The algebra of lanczos_pipelined (csrc/lpp_engine.cu, opt-in LPP_PIPELINE): with un-normalised vectors U_j = n_j v_j the sweeps
    can run without any scalar (w = H U_j) when every scalar enters afterwards,
        a_j = <U_j, w> / n_j^2,   U_{j+1} = w / n_j - (a_j / n_j) U_j - (b_j / n_{j-1}) U_{j-1},   b_{j+1} = n_{j+1} = |U_{j+1}|,
    which is what k_lzp_after_dot / k_lzp_after_norm / k_unpack3_norm_p2p compute.  Checked against the oracle's decomposition
    (LanczosSolver semantics) on a small Hubbard case."""
    case = cases.hubbard_chain(8, 4, 4, periodic=True, V=np.linspace(-0.3, 0.4, 8))
    o = cases.make_oracle(oracle, case, fast_rank=1)
    n = o.rows()
    init = geo.splitmix64_vector(n, 1234)
    steps = 25
    a0, b0 = o.decomposition(init, steps=steps, eps=0.0)
    u, uprev = init.copy(), np.zeros(n)
    nj, nprev, bj = float(np.sqrt(init @ init)), 1.0, 0.0
    a, b = [], []
    for _ in range(steps):
        w = np.zeros(n)
        o.matvec(w, u, faithful=False)
        aj = float(u @ w) / nj / nj
        unext = w / nj - (aj / nj) * u - (bj / nprev) * uprev
        a.append(aj)
        bj = float(np.sqrt(unext @ unext))
        b.append(bj)
        uprev, u, nprev, nj = u, unext, nj, bj
    assert np.abs(np.array(a) - a0[:steps]).max() <= 1e-10 * np.abs(a0).max()
    assert np.abs(np.array(b[:-1]) - b0[:steps - 1]).max() <= 1e-10 * np.abs(b0).max()
