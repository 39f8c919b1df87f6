"""Connection matrices for the lattices the configs use (host side, numpy only).

Mirrors what PsimagLite::Geometry hands the reference models: ``geometry(i, orb_i, j, orb_j, term)``
(call sites HubbardHelper.h:60-71, FeBasedSc.h:320-323, Heisenberg.h:56-57).
The engine's boundary takes these dense (nsite*orbitals)^2 matrices, so any PsimagLite geometry can be
fed by evaluating it on the host once.
"""
import numpy as np

_MASK64 = (1 << 64) - 1


def chain(nsite, value=-1.0, periodic=False):
    """GeometryKind=chain, GeometryOptions=ConstantValues, `Connectors 1 value` (TestSuite/inputs/input0.inp:4-7)."""
    m = np.zeros((nsite, nsite))
    for i in range(nsite - 1):
        m[i, i + 1] = m[i + 1, i] = value
    if periodic and nsite > 2:
        m[0, nsite - 1] = m[nsite - 1, 0] = value
    return m


def square(lx, ly, value=-1.0, periodic_x=True, periodic_y=True):
    """Nearest-neighbour lx*ly cluster, site = x + lx*y (config 3: 4x4 with both directions periodic => 32 bonds)."""
    n = lx * ly
    m = np.zeros((n, n))

    def bond(a, b):
        if a != b:
            m[a, b] = m[b, a] = value

    for y in range(ly):
        for x in range(lx):
            s = x + lx * y
            if x + 1 < lx:
                bond(s, s + 1)
            elif periodic_x and lx > 2:
                bond(s, lx * y)
            if y + 1 < ly:
                bond(s, s + lx)
            elif periodic_y and ly > 2:
                bond(s, x)
    return m


def with_orbitals(site_matrix, orbitals, diag_value_scale=1.0, offdiag_value_scale=0.0):
    """Expand a site-level matrix to (site*orbitals) indices, bit position = site*orbitals + orb
    (BasisOneSpinFeAs.h:195-199).  Orbital-diagonal connections get `diag_value_scale`, inter-orbital
    connections on the same bond get `offdiag_value_scale` (config 4: "orbital hoppings")."""
    n = site_matrix.shape[0]
    nb = n * orbitals
    m = np.zeros((nb, nb))
    for i in range(n):
        for j in range(n):
            if site_matrix[i, j] == 0:
                continue
            for a in range(orbitals):
                for b in range(orbitals):
                    s = diag_value_scale if a == b else offdiag_value_scale
                    m[i * orbitals + a, j * orbitals + b] = site_matrix[i, j] * s
    return m


def splitmix64_vector(n, seed, offset=0, normalize=False):
    """Counter-based initial vector: element i = uniform(-0.5, 0.5) from splitmix64(seed, offset+i).
    Identical on the oracle, on one GPU and on any row sharding; the reference itself uses an
    unseeded PsimagLite::fillRandom (Engine.h:621)."""
    idx = np.arange(offset, offset + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed) * np.uint64(0xD1B54A32D192ED03) + np.uint64(
            0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    v = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) - 0.5
    if normalize:
        v /= np.linalg.norm(v)
    return v
