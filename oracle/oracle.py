"""ctypes binding of the CPU oracle (oracle/lanczos_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
Model code pinned by the reference's own headers (oracle/reference.py, tests/golden/); the PsimagLite solver parts
stay PARITY UNPINNED (see the lanczos_oracle.c header).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liblpp_oracle.so")

HUBBARD, FEAS, HEISENBERG, TJ = 0, 1, 2, 3
OP_C, OP_SZ, OP_CDAGGER, OP_N, OP_SPLUS, OP_SMINUS = 1, 2, 3, 4, 5, 6


def build(force=False):
    src = os.path.join(_HERE, "lanczos_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liblpp_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        dp = C.POINTER(C.c_double)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int] * 5 + [dp, dp, dp, C.c_int, dp, C.c_int, dp, C.c_int, C.c_int, C.c_int]
        L.orc_create_tj.restype = C.c_void_p
        L.orc_create_tj.argtypes = [C.c_int] * 3 + [dp, dp, dp, dp, dp, C.c_int]
        L.orc_row_words.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_perfect_index.restype = C.c_size_t
        L.orc_perfect_index.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_rows.restype = C.c_size_t
        L.orc_rows.argtypes = [C.c_void_p]
        L.orc_basis_size.restype = C.c_size_t
        L.orc_basis_size.argtypes = [C.c_void_p, C.c_int]
        L.orc_basis_words.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_rank.restype = C.c_size_t
        L.orc_rank.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.orc_onespin_basis.restype = C.c_size_t
        L.orc_onespin_basis.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_onespin_rank.restype = C.c_size_t
        L.orc_onespin_rank.argtypes = [C.c_int, C.c_uint64]
        L.orc_row.restype = C.c_int
        L.orc_row.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_diag.restype = C.c_double
        L.orc_diag.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_crs_build.restype = C.c_int64
        L.orc_crs_build.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_crs_matvec.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_matvec.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_matvec_range.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64]
        L.orc_lanczos_sweeps.restype = C.c_double
        L.orc_lanczos_sweeps.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_tridiag_eig.restype = C.c_int
        L.orc_tridiag_eig.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_lanczos_decomposition.restype = C.c_int
        L.orc_lanczos_decomposition.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int,
                                                C.c_void_p, C.c_void_p]
        L.orc_lanczos_decomposition_reortho.restype = C.c_int
        L.orc_lanczos_decomposition_reortho.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p,
                                                        C.c_void_p]
        L.orc_states_below.restype = C.c_int
        L.orc_states_below.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_lanczos_decomposition_crs.restype = C.c_int
        L.orc_lanczos_decomposition_crs.argtypes = [C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_ground_state.restype = C.c_int
        L.orc_ground_state.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int,
                                       C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_cf_eval.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int,
                                  C.c_void_p, C.c_double, C.c_void_p]
        L.orc_apply_op.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                   C.c_void_p]
        L.orc_apply_op_orb.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                       C.c_void_p]
        L.orc_two_point.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def onespin_basis(nsite, npart):
    n = lib().orc_onespin_basis(nsite, npart, None)
    out = np.zeros(n, dtype=np.uint64)
    lib().orc_onespin_basis(nsite, npart, out.ctypes.data)
    return out


def onespin_rank(nsite, word):
    return lib().orc_onespin_rank(nsite, int(word))


def tridiag_eig(a, b, vectors=False):
    n = len(a)
    d = np.array(a, dtype=np.float64)
    e = np.zeros(n, dtype=np.float64)
    e[: n - 1] = np.asarray(b, dtype=np.float64)[: n - 1]
    z = np.zeros((n, n), dtype=np.float64) if vectors else None
    rc = lib().orc_tridiag_eig(n, d.ctypes.data, e.ctypes.data, z.ctypes.data if vectors else None)
    assert rc == 0
    return (d, z) if vectors else d


def cf_eval(a, b, eg, weight, isign, omega, delta):
    a = _f64(a)
    b = _f64(b)
    omega = _f64(omega)
    out = np.zeros(2 * len(omega))
    lib().orc_cf_eval(len(a), a.ctypes.data, b.ctypes.data, eg, weight, isign, len(omega), omega.ctypes.data, delta,
                      out.ctypes.data)
    return out[0::2] + 1j * out[1::2]


class OracleModel:
    """CPU restatement of one (model, sector): HubbardOneOrbital / FeBasedSc(INT_PAPER33) / Heisenberg S=1/2."""

    def __init__(self, model, nsite, nup, ndown=0, orbitals=1, hop=None, jzz=None, U=None, V=None, D=None,
                 u3_all_pairs=1, fast_rank=0, jpm=None, w=None):
        self.model, self.nsite, self.orbitals = model, nsite, (orbitals if model == FEAS else 1)
        self.nup, self.ndown = nup, ndown
        hop, jzz, U, V, D, jpm, w = map(_f64, (hop, jzz, U, V, D, jpm, w))
        self._keep = (hop, jzz, U, V, D, jpm, w)
        if model == TJ:      # Tj1Orbital: geometry terms hop, jpm, jzz, w (TjMultiOrb.h:68-79)
            self.h = lib().orc_create_tj(nsite, nup, ndown, _dptr(hop), _dptr(jpm), _dptr(jzz), _dptr(w), _dptr(V),
                                         0 if V is None else V.size)
            return
        self.h = lib().orc_create(model, nsite, orbitals, nup, ndown, _dptr(hop), _dptr(jzz), _dptr(U),
                                  0 if U is None else U.size, _dptr(V), 0 if V is None else V.size, _dptr(D),
                                  0 if D is None else D.size, u3_all_pairs, fast_rank)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def rows(self):
        return lib().orc_rows(self.h)

    def basis(self, spin):
        n = lib().orc_basis_size(self.h, spin)
        out = np.zeros(n, dtype=np.uint64)
        lib().orc_basis_words(self.h, spin, out.ctypes.data)
        return out

    def row_words(self, spin):
        """basis(i, spin) for every row."""
        out = np.zeros(self.rows(), dtype=np.uint64)
        lib().orc_row_words(self.h, spin, out.ctypes.data)
        return out

    def perfect_index(self, ket1, ket2):
        return lib().orc_perfect_index(self.h, int(ket1), int(ket2))

    def rank(self, spin, word):
        return lib().orc_rank(self.h, spin, int(word))

    def row(self, r, stored=True, cap=4096):
        cols = np.zeros(cap, dtype=np.uint64)
        vals = np.zeros(cap, dtype=np.float64)
        n = lib().orc_row(self.h, r, 1 if stored else 0, cols.ctypes.data, vals.ctypes.data, cap)
        return cols[:n].astype(np.int64), vals[:n]

    def diag(self, r):
        return lib().orc_diag(self.h, r)

    def crs(self):
        n = self.rows()
        nnz = lib().orc_crs_build(self.h, None, None, None)
        rowptr = np.zeros(n + 1, dtype=np.int64)
        colind = np.zeros(nnz, dtype=np.int64)
        vals = np.zeros(nnz, dtype=np.float64)
        lib().orc_crs_build(self.h, rowptr.ctypes.data, colind.ctypes.data, vals.ctypes.data)
        return rowptr, colind, vals

    def matvec(self, x, y, faithful=True):
        """x += H y (in place on x)."""
        assert x.dtype == np.float64 and y.dtype == np.float64 and x.flags.c_contiguous and y.flags.c_contiguous
        lib().orc_matvec(self.h, x.ctypes.data, y.ctypes.data, 1 if faithful else 0)
        return x

    def matvec_range(self, x_local, y, r0, r1, faithful=True):
        """x_local[r - r0] += (H y)[r] for r in [r0, r1): bounded-sample CPU baseline."""
        lib().orc_matvec_range(self.h, x_local.ctypes.data, y.ctypes.data, 1 if faithful else 0, r0, r1)
        return x_local

    def decomposition(self, init, steps=200, eps=1e-12, minsteps=4, faithful=False):
        n = self.rows()
        cap = min(steps, n) + 1
        a = np.zeros(cap)
        b = np.zeros(cap)
        init = _f64(init)
        ns = lib().orc_lanczos_decomposition(self.h, 1 if faithful else 0, init.ctypes.data, steps, eps, minsteps,
                                             a.ctypes.data, b.ctypes.data)
        return a[:ns].copy(), b[:ns].copy()

    def decomposition_reortho(self, init, steps=200, eps=1e-12, minsteps=4):
        """Lanczos decomposition with every vector saved and full reorthogonalisation (<prefix>Options=reortho)."""
        n = self.rows()
        cap = min(steps, n) + 1
        a = np.zeros(cap)
        b = np.zeros(cap)
        init = _f64(init)
        ns = lib().orc_lanczos_decomposition_reortho(self.h, init.ctypes.data, steps, eps, minsteps, a.ctypes.data,
                                                     b.ctypes.data)
        return a[:ns].copy(), b[:ns].copy()

    def states_below(self, init, nstates, steps=200, eps=1e-12, minsteps=4):
        """Lowest nstates Ritz pairs with full reorthogonalisation (computeAllStatesBelow, Engine.h:626)."""
        n = self.rows()
        init = _f64(init)
        eigs = np.zeros(nstates)
        zs = np.zeros((nstates, n))
        ns = lib().orc_states_below(self.h, init.ctypes.data, steps, eps, minsteps, nstates, eigs.ctypes.data, zs.ctypes.data)
        return eigs, zs, ns

    def ground_state(self, init, steps=200, eps=1e-12, minsteps=4, want_vector=True, faithful=False):
        n = self.rows()
        cap = min(steps, n) + 1
        a = np.zeros(cap)
        b = np.zeros(cap)
        init = _f64(init)
        z = np.zeros(n) if want_vector else None
        e = C.c_double(0)
        ns = lib().orc_ground_state(self.h, 1 if faithful else 0, init.ctypes.data, steps, eps, minsteps, C.byref(e),
                                    z.ctypes.data if want_vector else None, a.ctypes.data, b.ctypes.data)
        return e.value, z, a[:ns].copy(), b[:ns].copy()

    def apply_op(self, dst, op, site, spin, factor, srcv, z, orb=0):
        srcv = _f64(srcv)
        lib().orc_apply_op_orb(self.h, dst.h, op, site, spin, orb, factor, srcv.ctypes.data, z.ctypes.data)
        return z


def two_point(src, dst, op, spin, gs, orbs=(0, 0)):
    """Engine::twoPoint with bra = ket = gs: nsite x nsite matrix <O_j gs | O_i gs>."""
    gs = _f64(gs)
    out = np.zeros((src.nsite, src.nsite))
    lib().orc_two_point(src.h, dst.h, op, spin, orbs[0], orbs[1], gs.ctypes.data, out.ctypes.data)
    return out


def crs_matvec(rowptr, colind, vals, x, y):
    lib().orc_crs_matvec(len(rowptr) - 1, rowptr.ctypes.data, colind.ctypes.data, vals.ctypes.data, x.ctypes.data,
                         y.ctypes.data)
    return x


def decomposition_crs(rowptr, colind, vals, init, steps=200, eps=1e-12, minsteps=4):
    n = len(rowptr) - 1
    cap = min(steps, n) + 1
    a = np.zeros(cap)
    b = np.zeros(cap)
    init = _f64(init)
    ns = lib().orc_lanczos_decomposition_crs(n, rowptr.ctypes.data, colind.ctypes.data, vals.ctypes.data,
                                             init.ctypes.data, steps, eps, minsteps, a.ctypes.data, b.ctypes.data)
    return a[:ns].copy(), b[:ns].copy()


def lanczos_sweeps(x, y):
    return lib().orc_lanczos_sweeps(x.ctypes.data, y.ctypes.data, x.size)


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    """OpenMP thread count of the oracle (overrides OMP_NUM_THREADS, which torchrun sets to 1)."""
    lib().orc_set_num_threads(int(n))
