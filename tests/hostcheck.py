"""TEST-ONLY ctypes binding of tests/hostcheck.cpp (the product's host/device basis+row code compiled for the CPU)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_hostcheck.so")
_DEPS = [os.path.join(_HERE, "hostcheck.cpp"),
         os.path.join(_HERE, "..", "lanczosplusplus_b200", "csrc", "lpp_device.cuh"),
         os.path.join(_HERE, "..", "lanczosplusplus_b200", "csrc", "lpp_setup.h")]
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in _DEPS):
            subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", _DEPS[0], "-o", _SO])
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        L.hc_create.restype = C.c_void_p
        L.hc_create.argtypes = [C.c_int] * 5 + [dp, dp, dp, C.c_int, dp, C.c_int, dp, C.c_int, C.c_int, C.c_int]
        L.hc_set_tj.argtypes = [C.c_void_p, dp, dp]
        L.hc_row_words.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hc_rank_pair.restype = C.c_uint64
        L.hc_rank_pair.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.hc_rahul.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hc_destroy.argtypes = [C.c_void_p]
        L.hc_rows.restype = C.c_uint64
        L.hc_rows.argtypes = [C.c_void_p]
        L.hc_basis_size.restype = C.c_uint64
        L.hc_basis_size.argtypes = [C.c_void_p, C.c_int]
        L.hc_basis.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.hc_rank.restype = C.c_uint64
        L.hc_rank.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.hc_crs.restype = C.c_int64
        L.hc_crs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.hc_matvec.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hc_apply_op.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.hc_splitmix.restype = C.c_double
        L.hc_splitmix.argtypes = [C.c_uint64, C.c_uint64]
        _lib = L
    return _lib


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class HostModel:
    def __init__(self, case, use_tables=1):
        c = dict(case)
        arrs = [_f(c.get(k)) for k in ("hop", "jzz", "U", "V", "D")]
        self._keep = arrs
        hop, jzz, U, V, D = arrs
        self.h = lib().hc_create(c["model"], c["nsite"], c["orbitals"], c["nup"], c["ndown"], _p(hop), _p(jzz), _p(U),
                                 0 if U is None else U.size, _p(V), 0 if V is None else V.size, _p(D),
                                 0 if D is None else D.size, c.get("feas_u3_all_pairs", 1), use_tables)
        if c["model"] == 3:
            self._keep_tj = [_f(c.get("jpm")), _f(c.get("w"))]
            lib().hc_set_tj(self.h, _p(self._keep_tj[0]), _p(self._keep_tj[1]))

    def __del__(self):
        if getattr(self, "h", None):
            lib().hc_destroy(self.h)
            self.h = None

    def rows(self):
        return lib().hc_rows(self.h)

    def basis(self, spin):
        out = np.zeros(lib().hc_basis_size(self.h, spin), dtype=np.uint64)
        lib().hc_basis(self.h, spin, out.ctypes.data)
        return out

    def rank(self, spin, w):
        return lib().hc_rank(self.h, spin, int(w))

    def row_words(self, spin):
        out = np.zeros(self.rows(), dtype=np.uint64)
        lib().hc_row_words(self.h, spin, out.ctypes.data)
        return out

    def rahul(self, ops, psi):
        a = np.array(ops, dtype=np.int32).reshape(-1, 4)
        lab, dof, site, tr = (np.ascontiguousarray(a[:, k]) for k in (0, 1, 2, 3))
        psi = _f(psi)
        out = np.zeros(self.rows())
        lib().hc_rahul(self.h, len(a), lab.ctypes.data, dof.ctypes.data, tr.ctypes.data, site.ctypes.data, psi.ctypes.data,
                       out.ctypes.data)
        return out

    def rank_pair(self, k1, k2):
        return lib().hc_rank_pair(self.h, int(k1), int(k2))

    def crs(self):
        nnz = lib().hc_crs(self.h, None, None, None)
        assert nnz >= 0
        rp = np.zeros(self.rows() + 1, dtype=np.int64)
        ci = np.zeros(nnz, dtype=np.int64)
        v = np.zeros(nnz)
        lib().hc_crs(self.h, rp.ctypes.data, ci.ctypes.data, v.ctypes.data)
        return rp, ci, v

    def matvec(self, x, y):
        lib().hc_matvec(self.h, x.ctypes.data, y.ctypes.data)
        return x

    def apply_op(self, dst, op, site, spin, factor, srcv, z, orb=0):
        srcv = _f(srcv)
        lib().hc_apply_op(self.h, dst.h, op, site, spin, orb, factor, srcv.ctypes.data, z.ctypes.data)
        return z
