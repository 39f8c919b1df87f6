"""ctypes binding of oracle/_ref/liblpp_ref.so: the REFERENCE'S OWN model headers (bases, signs, row generators, stored
Hamiltonian assembly, on-the-fly product) compiled unmodified from /root/reference/src against oracle/psimag_shim/
(see oracle/ref_bridge.cpp for what that does and does not pin).

TEST INFRASTRUCTURE ONLY.  It exists where /root/reference exists (this container); on the GPU box only the prebuilt
.so travels.  Used to (1) pin oracle/lanczos_oracle.c (tests/test_reference_pin.py) and (2) generate the golden
fixtures under tests/golden/ (tools/make_golden.py).  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "liblpp_ref.so")
ADAPTER_CHECK = os.path.join(_HERE, "_ref", "adapter_check")
REFERENCE_SRC = "/root/reference/src"

HUBBARD, FEAS, HEISENBERG, TJ = 0, 1, 2, 3
OP_C, OP_SZ, OP_CDAGGER, OP_N, OP_SPLUS, OP_SMINUS = 1, 2, 3, 4, 5, 6


def available():
    return os.path.exists(_LIB) or os.path.isdir(REFERENCE_SRC)


def build(force=False):
    """Compile the bridge when the reference sources are present; otherwise use the prebuilt library if there is one."""
    if os.path.isdir(REFERENCE_SRC):
        srcs = [os.path.join(_HERE, "ref_bridge.cpp")] + [os.path.join(_HERE, "psimag_shim", f)
                                                          for f in os.listdir(os.path.join(_HERE, "psimag_shim"))
                                                          if f.endswith(".h")]
        stale = not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs)
        adapter_srcs = srcs[1:] + [os.path.join(_HERE, "..", "tests", "adapter_check.cpp"),
                                   os.path.join(_HERE, "..", "include", "InternalProductCuda.h"),
                                   os.path.join(_HERE, "..", "include", "lpp_b200.h")]
        stale = stale or not os.path.exists(ADAPTER_CHECK) or any(os.path.getmtime(s) > os.path.getmtime(ADAPTER_CHECK)
                                                                  for s in adapter_srcs)
        if force or stale:
            subprocess.check_call(["make", "-C", _HERE, "_ref"], stdout=subprocess.DEVNULL)
    return _LIB if os.path.exists(_LIB) else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        if build() is None:
            raise RuntimeError("oracle/_ref/liblpp_ref.so is not built and /root/reference is absent")
        L = C.CDLL(_LIB)
        dp = C.POINTER(C.c_double)
        L.ref_last_error.restype = C.c_char_p
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int] * 5 + [dp, dp, dp, C.c_int, dp, C.c_int, dp, C.c_int]
        L.ref_create_tj.restype = C.c_void_p
        L.ref_create_tj.argtypes = [C.c_int] * 3 + [dp, dp, dp, dp, dp, C.c_int]
        L.ref_new_sector.restype = C.c_void_p
        L.ref_new_sector.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_rows.restype = C.c_uint64
        L.ref_rows.argtypes = [C.c_void_p]
        L.ref_basis_words.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_perfect_index.restype = C.c_int64
        L.ref_perfect_index.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.ref_crs.restype = C.c_int64
        L.ref_crs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_matvec.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_apply_op.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                   C.c_void_p]
        L.ref_has_new_parts.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                        C.POINTER(C.c_int)]
        L.ref_rahul.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def set_threads(n):
    lib().ref_set_threads(int(n))


class ReferenceModel:
    """One (model, sector) of the reference: HubbardOneOrbital / FeBasedSc INT_PAPER33 / Heisenberg S=1/2."""

    def __init__(self, model, nsite, nup, ndown=0, orbitals=1, hop=None, jzz=None, U=None, V=None, D=None, _handle=None,
                 _parent=None, jpm=None, w=None):
        self.model, self.nsite, self.nup, self.ndown = model, nsite, nup, ndown
        self.orbitals = orbitals if model == FEAS else 1
        self._parent = _parent                     # keeps the owning model alive for new-sector handles
        if _handle is not None:
            self.h = _handle
            return
        hop, jzz, U, V, D, jpm, w = map(_f64, (hop, jzz, U, V, D, jpm, w))
        if model == TJ:
            self.h = lib().ref_create_tj(nsite, nup, ndown, _dptr(hop), _dptr(jpm), _dptr(jzz), _dptr(w), _dptr(V),
                                         0 if V is None else V.size)
            if not self.h:
                raise RuntimeError("reference: " + lib().ref_last_error().decode())
            return
        self.h = lib().ref_create(model, nsite, self.orbitals, nup, ndown, _dptr(hop), _dptr(jzz), _dptr(U),
                                  0 if U is None else U.size, _dptr(V), 0 if V is None else V.size, _dptr(D),
                                  0 if D is None else D.size)
        if not self.h:
            raise RuntimeError("reference: " + lib().ref_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_destroy(self.h)
            self.h = None

    def new_sector(self, nup, ndown):
        h = lib().ref_new_sector(self.h, nup, ndown)
        if not h:
            raise RuntimeError("reference: " + lib().ref_last_error().decode())
        return ReferenceModel(self.model, self.nsite, nup, ndown, self.orbitals, _handle=h, _parent=self)

    def rows(self):
        return int(lib().ref_rows(self.h))

    def row_words(self, spin):
        """basis(i, spin) for every row i."""
        out = np.zeros(self.rows(), dtype=np.uint64)
        assert lib().ref_basis_words(self.h, spin, out.ctypes.data) == 0, lib().ref_last_error()
        return out

    def perfect_index(self, ket1, ket2):
        r = lib().ref_perfect_index(self.h, int(ket1), int(ket2))
        if r < 0:
            raise RuntimeError("reference: " + lib().ref_last_error().decode())
        return r

    def crs(self):
        nnz = lib().ref_crs(self.h, None, None, None)
        if nnz < 0:
            raise RuntimeError("reference: " + lib().ref_last_error().decode())
        rowptr = np.zeros(self.rows() + 1, dtype=np.int64)
        colind = np.zeros(nnz, dtype=np.int64)
        vals = np.zeros(nnz, dtype=np.float64)
        lib().ref_crs(self.h, rowptr.ctypes.data, colind.ctypes.data, vals.ctypes.data)
        return rowptr, colind, vals

    def matvec(self, x, y):
        """x += H y through the model's on-the-fly product (raises for Heisenberg: the reference has none)."""
        assert x.dtype == np.float64 and y.dtype == np.float64 and x.flags.c_contiguous and y.flags.c_contiguous
        if lib().ref_matvec(self.h, x.ctypes.data, y.ctypes.data) != 0:
            raise RuntimeError("reference: " + lib().ref_last_error().decode())
        return x

    def apply_op(self, dst, op, site, spin, factor, srcv, z, orb=0):
        srcv = _f64(srcv)
        if lib().ref_apply_op(self.h, dst.h, op, site, spin, orb, factor, srcv.ctypes.data, z.ctypes.data) != 0:
            raise RuntimeError("reference: " + lib().ref_last_error().decode())
        return z

    def rahul(self, ops, psi):
        """psiNew = prod ops |psi> by ModelBase::rahulMethod; ops = [(label 0..3, dof, site, transpose), ...]."""
        a = np.array(ops, dtype=np.int32).reshape(-1, 4)
        lab, dof, site, tr = (np.ascontiguousarray(a[:, k]) for k in (0, 1, 2, 3))
        psi = _f64(psi)
        out = np.zeros(self.rows())
        if lib().ref_rahul(self.h, len(a), lab.ctypes.data, dof.ctypes.data, tr.ctypes.data, site.ctypes.data, psi.ctypes.data,
                           out.ctypes.data) != 0:
            raise RuntimeError("reference: " + lib().ref_last_error().decode())
        return out

    def has_new_parts(self, op, spin, orb=0):
        a, b = C.c_int(0), C.c_int(0)
        r = lib().ref_has_new_parts(self.h, op, spin, orb, self.nup, self.ndown, C.byref(a), C.byref(b))
        if r < 0:
            raise RuntimeError("reference: " + lib().ref_last_error().decode())
        return bool(r), (a.value, b.value)
