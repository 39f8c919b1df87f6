"""Model cases shared by the CPU and GPU parity tests (same seeded inputs for the oracle and the engine)."""
import numpy as np

from lanczosplusplus_b200 import geometry as geo

HUBBARD, FEAS, HEISENBERG, TJ = 0, 1, 2, 3


def hubbard_chain(n, nup, ndn, U=4.0, t=-1.0, periodic=False, V=None):
    return dict(model=HUBBARD, nsite=n, nup=nup, ndown=ndn, orbitals=1, hop=geo.chain(n, t, periodic),
                U=np.full(n, U), V=np.zeros(n) if V is None else np.asarray(V, dtype=float))


def hubbard_square(lx, ly, nup, ndn, U=4.0, t=-1.0):
    n = lx * ly
    return dict(model=HUBBARD, nsite=n, nup=nup, ndown=ndn, orbitals=1, hop=geo.square(lx, ly, t),
                U=np.full(n, U), V=np.zeros(n))


def hubbard_random(n, nup, ndn, seed):
    """dense symmetric random hoppings, site-dependent U and V: exercises every sign/rank path"""
    rng = np.random.default_rng(seed)
    h = rng.uniform(-1, 1, (n, n))
    h = np.triu(h, 1)
    h = h + h.T
    return dict(model=HUBBARD, nsite=n, nup=nup, ndown=ndn, orbitals=1, hop=h, U=rng.uniform(0, 8, n),
                V=rng.uniform(-1, 1, n))


def feas_chain(n, nup, ndn, inter_orbital=0.0, with_potential=True, u3_all_pairs=1, D=0.0, periodic=False):
    """input100.inp parameters (TestSuite/inputs/input100.inp): U = 4 3 -0.8 -0.4, connectors -1 orbital-diagonal"""
    hop = geo.with_orbitals(geo.chain(n, -1.0, periodic), 2, 1.0, inter_orbital)
    V = np.zeros(4 * n)
    if with_potential:
        V[0:n] = 4.1
        V[2 * n:3 * n] = 4.1
    return dict(model=FEAS, nsite=n, nup=nup, ndown=ndn, orbitals=2, hop=hop, U=np.array([4.0, 3.0, -0.8, -0.4]), V=V,
                D=np.array([D]), feas_u3_all_pairs=u3_all_pairs)


def feas_cluster(lx, ly, nup, ndn, inter_orbital=0.5):
    n = lx * ly
    hop = geo.with_orbitals(geo.square(lx, ly, -1.0, periodic_x=lx > 2, periodic_y=ly > 2), 2, 1.0, inter_orbital)
    return dict(model=FEAS, nsite=n, nup=nup, ndown=ndn, orbitals=2, hop=hop, U=np.array([4.0, 3.0, -0.8, -0.4]),
                V=np.zeros(4 * n), D=np.array([0.0]), feas_u3_all_pairs=1)


def heisenberg_ring(n, szplus, jz=1.0, field=None):
    J = geo.chain(n, 1.0, True)
    return dict(model=HEISENBERG, nsite=n, nup=szplus, ndown=0, orbitals=1, hop=J, jzz=jz * J,
                V=None if field is None else np.asarray(field, dtype=float))


def tj_chain(n, nup, ndn, t=-1.0, J=0.4, periodic=False, V=None, seed=None):
    """Tj1Orbital: H = sum t c+c (projected) + J (S.S - n n / 4): geometry terms hop = t, jpm = jzz = J, w = -J/4
    (TjMultiOrb.h:68-79).  With `seed`, bond-dependent random couplings exercise every term separately."""
    bonds = geo.chain(n, 1.0, periodic)
    if seed is None:
        hop, jpm, jzz, w = t * bonds, J * bonds, J * bonds, -0.25 * J * bonds
    else:
        rng = np.random.default_rng(seed)

        def sym():
            a = np.triu(rng.uniform(-1, 1, (n, n)), 1)
            return (a + a.T) * (bonds + np.roll(bonds, 1, axis=0) * 0 + (rng.uniform(0, 1, (n, n)) < 0.3) * (1 - np.eye(n)) > 0)
        hop, jpm, jzz, w = sym(), sym(), sym(), sym()
        for m in (hop, jpm, jzz, w):
            m[:] = np.triu(m, 1) + np.triu(m, 1).T
    return dict(model=TJ, nsite=n, nup=nup, ndown=ndn, orbitals=1, hop=hop, jpm=jpm, jzz=jzz, w=w,
                V=None if V is None else np.asarray(V, dtype=float))


def tj_square(lx, ly, nup, ndn, t=-1.0, J=0.4):
    bonds = geo.square(lx, ly, 1.0)
    return dict(model=TJ, nsite=lx * ly, nup=nup, ndown=ndn, orbitals=1, hop=t * bonds, jpm=J * bonds, jzz=J * bonds,
                w=-0.25 * J * bonds, V=None)


def make_oracle(orc, case, fast_rank=1):
    c = dict(case)
    u3 = c.pop("feas_u3_all_pairs", 1)
    return orc.OracleModel(c.pop("model"), c.pop("nsite"), c.pop("nup"), c.pop("ndown"), c.pop("orbitals"),
                           hop=c.get("hop"), jzz=c.get("jzz"), U=c.get("U"), V=c.get("V"), D=c.get("D"),
                           u3_all_pairs=u3, fast_rank=fast_rank, jpm=c.get("jpm"), w=c.get("w"))


def make_engine(lpp, case, **kw):
    c = dict(case)
    return lpp.InternalProductCuda(c.pop("model"), c.pop("nsite"), c.pop("nup"), c.pop("ndown"), c.pop("orbitals"),
                                   hop=c.get("hop"), jzz=c.get("jzz"), U=c.get("U"), V=c.get("V"), D=c.get("D"),
                                   feas_u3_all_pairs=c.get("feas_u3_all_pairs", 1), jpm=c.get("jpm"), w=c.get("w"), **kw)


SMALL_CASES = {
    "input0": hubbard_chain(4, 2, 2, U=0.0),
    "hub2": hubbard_chain(2, 1, 1),
    "c1_hub8": hubbard_chain(8, 4, 4),
    "hub6_pbc_V": hubbard_chain(6, 3, 2, periodic=True, V=[0.3, -0.2, 0.1, 0.0, 0.5, -0.4]),
    "hub_rand7": hubbard_random(7, 3, 4, 7),
    "hub_3x3": hubbard_square(3, 3, 4, 5),
    "hub_empty_dn": hubbard_chain(5, 2, 0),
    "feas2": feas_chain(2, 1, 1),
    "feas3": feas_chain(3, 2, 1),
    "feas4": feas_chain(4, 2, 2),
    "feas4_interorb_otfquirk": feas_chain(4, 2, 3, inter_orbital=0.5, u3_all_pairs=0, D=0.7, periodic=True),
    "feas_2x2": feas_cluster(2, 2, 3, 2),
    "heis4": heisenberg_ring(4, 2),
    "heis12": heisenberg_ring(12, 6),
    "heis10_field": heisenberg_ring(10, 4, jz=0.7, field=np.linspace(-0.5, 0.5, 10)),
}

# Tj1Orbital (stored path + generic on-the-fly kernel; not a product basis)
TJ_CASES = {
    "tj4": tj_chain(4, 1, 2),
    "tj6_pbc": tj_chain(6, 2, 2, periodic=True),
    "tj8_V": tj_chain(8, 3, 3, V=np.linspace(-0.3, 0.4, 16)),
    "tj7_rand": tj_chain(7, 3, 2, seed=5),
    "tj_3x3": tj_square(3, 3, 4, 3),
    "tj6_full": tj_chain(6, 3, 3),          # no holes: pure Heisenberg limit
    "tj5_no_dn": tj_chain(5, 2, 0),
}
