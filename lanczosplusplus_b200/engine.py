"""Host-side mirror of the reference interface for the hot path, bound to liblpp_b200.so through ctypes.

Names follow the reference (file:line under /root/reference/src):
  InternalProductCuda   sibling of InternalProductOnTheFly / InternalProductStored (Engine/InternalProductOnTheFly.h:88-139)
  LanczosSolver         PsimagLite::LanczosSolver as used at Engine.h:609-626 and :472-478
  ContinuedFraction     PsimagLite::ContinuedFraction as used at Engine.h:487-489
  Engine                Engine.h:84-98 (ground state) and :133-206 (spectralFunction type loop)
All arithmetic runs on the GPU inside the library; nothing here falls back to the CPU.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Desc, LppError, SolverParams, Timing, check

HUBBARD, FEAS, HEISENBERG, TJ = 0, 1, 2, 3
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_TABLE, KERNEL_TILED, KERNEL_STORED = 0, 1, 2, 3, 4
OP_C, OP_SZ, OP_CDAGGER, OP_N, OP_SPLUS, OP_SMINUS = 1, 2, 3, 4, 5, 6
MODEL_NAMES = {"HubbardOneBand": HUBBARD, "FeAsBasedSc": FEAS, "Heisenberg": HEISENBERG, "Tj1Orbital": TJ, "TjMultiOrb": TJ}


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def kernel_from_solver_options(options):
    """SolverOptions= substring dispatch, LanczosDriver1.h:222-238 with the new InternalProductCuda token."""
    if "InternalProductStored" in options:
        return KERNEL_STORED
    if "InternalProductCudaGeneric" in options:
        return KERNEL_GENERIC
    if "InternalProductCudaTable" in options:
        return KERNEL_TABLE
    return KERNEL_AUTO


class InternalProductCuda:
    """x += H y for one (model, sector) on one B200 (or one row shard of it).

    Mirrors InternalProductOnTheFly: rows(), matrixVectorProduct(x, y) (InternalProductOnTheFly.h:115-123)."""

    def __init__(self, model, nsite, nup, ndown=0, orbitals=1, hop=None, jzz=None, U=None, V=None, D=None,
                 feas_u3_all_pairs=1, device=0, rank=0, nranks=1, kernel=KERNEL_AUTO, jpm=None, w=None):
        if isinstance(model, str):
            model = MODEL_NAMES[model]
        self.model, self.nsite, self.nup, self.ndown = model, nsite, nup, ndown
        self.orbitals = orbitals if model == FEAS else 1
        self.kernel = kernel
        self._keep = tuple(map(_f64, (hop, jzz, U, V, D)))
        self._keep_tj = tuple(map(_f64, (jpm, w)))          # t-J: geometry terms 1 and 3 (TjMultiOrb.h:68-79)
        hop, jzz, U, V, D = self._keep
        jpm, w = self._keep_tj
        if hop is None:
            raise LppError("hop matrix is required")
        d = Desc(model, nsite, self.orbitals, nup, ndown, feas_u3_all_pairs, _dp(hop), _dp(jzz), _dp(U),
                 0 if U is None else U.size, _dp(V), 0 if V is None else V.size, _dp(D), 0 if D is None else D.size,
                 device, rank, nranks, _dp(jpm), _dp(w))
        self._desc = d
        self.h = C.c_void_p()
        check(_lib.lib().lpp_create(C.byref(d), C.byref(self.h)))
        self.rank, self.nranks = rank, nranks

    def close(self):
        if getattr(self, "h", None) and self.h:
            _lib.lib().lpp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sector(self, nup, ndown, **kw):
        """model.createBasis(nup', ndown') + InternalProduct on it (Engine.h:165-187)."""
        hop, jzz, U, V, D = self._keep
        s = InternalProductCuda(self.model, self.nsite, nup, ndown, self.orbitals, hop, jzz, U, V, D,
                                self._desc.feas_u3_all_pairs, self._desc.device, self.rank, self.nranks,
                                kw.get("kernel", self.kernel), jpm=self._keep_tj[0], w=self._keep_tj[1])
        if self.nranks > 1 and getattr(self, "_has_comm", False):
            check(_lib.lib().lpp_comm_share(s.h, self.h))      # same ranks, same device: borrow the communicator
            s._has_comm = True
            dist = getattr(self, "_dist", None)
            if dist is not None and getattr(self, "_has_p2p", False):
                # peer-memory exchange for the new sector as well (every rank creates the same sector at the same point of
                # Engine::spectralFunction, so this is a collective): 128 bytes of IPC handles per rank
                mine = s.p2p_export()
                parts = [None] * dist.get_world_size()
                dist.all_gather_object(parts, mine)
                if all(q is not None for q in parts):
                    s.p2p_import(b"".join(parts))
                    s._has_p2p = True
                s._dist = dist
        return s

    # --- InternalProductOnTheFly interface
    def rows(self):
        n = C.c_uint64()
        check(_lib.lib().lpp_rows(self.h, C.byref(n)))
        return n.value

    def local_rows(self):
        f, c = C.c_uint64(), C.c_uint64()
        check(_lib.lib().lpp_local_rows(self.h, C.byref(f), C.byref(c)))
        return f.value, c.value

    def matrixVectorProduct(self, x, y, kernel=None):
        """x += H y, host numpy vectors (the literal drop-in call; vectors cross PCIe both ways)."""
        if x.dtype != np.float64 or y.dtype != np.float64 or not x.flags.c_contiguous or not y.flags.c_contiguous:
            raise LppError("matrixVectorProduct expects contiguous float64 vectors")
        if x.size != self.rows() or y.size != self.rows():
            raise LppError("vector length must equal rows()")
        check(_lib.lib().lpp_matvec_host(self.h, self.kernel if kernel is None else kernel, x.ctypes.data,
                                         y.ctypes.data))
        return x

    def matvec_device(self, x_ptr, y_ptr, kernel=None):
        check(_lib.lib().lpp_matvec_device(self.h, self.kernel if kernel is None else kernel, x_ptr, y_ptr))

    # --- basis
    def basis(self, spin):
        n = C.c_uint64()
        check(_lib.lib().lpp_basis_size(self.h, spin, C.byref(n)))
        out = np.zeros(n.value, dtype=np.uint64)
        check(_lib.lib().lpp_basis_export(self.h, spin, out.ctypes.data))
        return out

    def perfectIndex(self, spin, words):
        w = np.ascontiguousarray(words, dtype=np.uint64)
        out = np.zeros(w.size, dtype=np.uint64)
        check(_lib.lib().lpp_rank(self.h, spin, w.ctypes.data, w.size, out.ctypes.data))
        return out

    def row_words(self, first=0, count=None):
        """basis(i, SPIN_UP), basis(i, SPIN_DOWN) for rows [first, first+count) (BasisBase::operator())."""
        count = self.rows() - first if count is None else count
        up, dn = np.zeros(count, dtype=np.uint64), np.zeros(count, dtype=np.uint64)
        check(_lib.lib().lpp_row_words(self.h, first, count, up.ctypes.data, dn.ctypes.data))
        return up, dn

    def perfectIndexPairs(self, up_words, down_words):
        """perfectIndex(ket1, ket2) of the full basis."""
        u = np.ascontiguousarray(up_words, dtype=np.uint64)
        d = np.ascontiguousarray(down_words, dtype=np.uint64)
        out = np.zeros(u.size, dtype=np.uint64)
        check(_lib.lib().lpp_rank_pairs(self.h, u.ctypes.data, d.ctypes.data, u.size, out.ctypes.data))
        return out

    # --- stored CRS (InternalProductStored)
    def setupHamiltonian(self):
        nnz = C.c_int64()
        check(_lib.lib().lpp_crs_build(self.h, C.byref(nnz)))
        _, nloc = self.local_rows()
        rowptr = np.zeros(nloc + 1, dtype=np.int64)
        colind = np.zeros(nnz.value, dtype=np.int64)
        values = np.zeros(nnz.value, dtype=np.float64)
        check(_lib.lib().lpp_crs_export(self.h, rowptr.ctypes.data, colind.ctypes.data, values.ctypes.data))
        return rowptr, colind, values

    # --- vectors kept on device
    def get_vector(self, which=0):
        out = np.zeros(self.rows())
        check(_lib.lib().lpp_get_vector(self.h, which, out.ctypes.data))
        return out

    def set_groundstate(self, z):
        z = _f64(z)
        check(_lib.lib().lpp_set_groundstate(self.h, z.ctypes.data))

    def apply_op(self, dst, op, site, spin, factor=1.0, accumulate=False, orb=0):
        """Engine::accModifiedState_ (Engine.h:416-458): dst.modified (+)= factor * O_{site,spin} |self.groundstate>."""
        check(_lib.lib().lpp_apply_op(self.h, dst.h, op, site, spin, orb, factor, 1 if accumulate else 0))

    def comm_init(self, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        check(_lib.lib().lpp_comm_init(self.h, buf))
        self._has_comm = True

    def p2p_export(self, kernel=None):
        """128 bytes of CUDA IPC handles of this rank's column-shard buffers, or None when two-layout sharding does not apply."""
        buf = (C.c_uint8 * 128)()
        rc = _lib.lib().lpp_p2p_export(self.h, self.kernel if kernel is None else kernel, buf)
        if rc == -3:
            return None
        check(rc)
        return bytes(buf)

    def p2p_import(self, all_handles):
        buf = (C.c_uint8 * len(all_handles)).from_buffer_copy(all_handles)
        check(_lib.lib().lpp_p2p_import(self.h, buf))

    # --- measurement hooks
    def bench_spmv(self, iters, warmup, kernel=None):
        t = Timing()
        check(_lib.lib().lpp_bench_spmv(self.h, self.kernel if kernel is None else kernel, iters, warmup, C.byref(t)))
        return t.spmv_ms, t.launches

    def bench_lanczos(self, iters, warmup, kernel=None, seed=1234):
        p = SolverParams(iters + warmup, 4, 0.0, self.kernel if kernel is None else kernel, 0, seed)
        t = Timing()
        check(_lib.lib().lpp_bench_lanczos(self.h, C.byref(p), iters, warmup, C.byref(t)))
        return t.iter_ms, t.launches


def comm_unique_id():
    buf = (C.c_uint8 * 128)()
    check(_lib.lib().lpp_comm_unique_id(buf))
    return bytes(buf)


class ParametersForSolver:
    """PsimagLite::ParametersForSolver(io, prefix): <prefix>Steps=, <prefix>Eps=, <prefix>MinSteps=, <prefix>Options=."""

    def __init__(self, io=None, prefix="Lanczos", steps=200, eps=1e-12, minsteps=4, seed=1234, options=""):
        io = io or {}
        self.steps = int(io.get(prefix + "Steps", steps))
        self.eps = float(io.get(prefix + "Eps", eps))
        self.minsteps = int(io.get(prefix + "MinSteps", minsteps))
        self.options = str(io.get(prefix + "Options", options))     # "reortho": full reorthogonalisation, vectors saved on device
        self.seed = seed


class LanczosSolver:
    """PsimagLite::LanczosSolver<Params, MatrixType, Vector> over an InternalProductCuda, device resident."""

    def __init__(self, matrix, params=None):
        self.mat = matrix
        self.params = params or ParametersForSolver()

    def _p(self):
        p = self.params
        return SolverParams(p.steps, p.minsteps, p.eps, self.mat.kernel, 1 if "reortho" in p.options else 0, p.seed)

    def decomposition(self, init=None, use_modified=False):
        """-> (a, b, <init|init>) ; Engine.h:474-478."""
        p = self._p()
        cap = min(p.steps, self.mat.rows()) + 1
        a, b = np.zeros(cap), np.zeros(cap)
        ns, nrm = C.c_int32(), C.c_double()
        init = _f64(init)
        check(_lib.lib().lpp_lanczos_decomposition(self.mat.h, C.byref(p), None if init is None else init.ctypes.data,
                                                   1 if use_modified else 0, a.ctypes.data, b.ctypes.data,
                                                   C.byref(ns), C.byref(nrm)))
        return a[:ns.value].copy(), b[:ns.value].copy(), nrm.value

    def computeOneState(self, init=None, want_vector=True, copy_vector=False):
        """-> (energy, z or None, a, b) ; Engine.h:626 computeAllStatesBelow(eigs, zs, initial, 1)."""
        p = self._p()
        cap = min(p.steps, self.mat.rows()) + 1
        a, b = np.zeros(cap), np.zeros(cap)
        ns, e = C.c_int32(), C.c_double()
        init = _f64(init)
        z = np.zeros(self.mat.rows()) if copy_vector else None
        check(_lib.lib().lpp_ground_state(self.mat.h, C.byref(p), None if init is None else init.ctypes.data,
                                          1 if want_vector else 0, C.byref(e), None if z is None else z.ctypes.data,
                                          a.ctypes.data, b.ctypes.data, C.byref(ns)))
        return e.value, z, a[:ns.value].copy(), b[:ns.value].copy()


def _states_below(solver, init, excited_plus_one, want_vectors):
    p = solver._p()
    eigs = np.zeros(excited_plus_one)
    init = _f64(init)
    zs = np.zeros((excited_plus_one, solver.mat.rows())) if want_vectors else None
    ns = C.c_int32()
    check(_lib.lib().lpp_states_below(solver.mat.h, C.byref(p), None if init is None else init.ctypes.data, excited_plus_one,
                                      eigs.ctypes.data, None if zs is None else zs.ctypes.data, C.byref(ns)))
    return eigs, zs, ns.value


def _computeAllStatesBelow(self, init=None, excited_plus_one=1, want_vectors=True):
    """-> (eigs, zs, steps): lanczosSolver.computeAllStatesBelow(eigs, zs, initial, excitedPlusOne), Engine.h:626."""
    return _states_below(self, init, excited_plus_one, want_vectors)


LanczosSolver.computeAllStatesBelow = _computeAllStatesBelow


def tridiag_eig(a, b, vectors=False):
    n = len(a)
    a = _f64(a)
    bb = np.zeros(max(n, 1))
    bb[: max(n - 1, 0)] = np.asarray(b, dtype=np.float64)[: max(n - 1, 0)]
    eigs = np.zeros(n)
    z = np.zeros((n, n)) if vectors else None
    check(_lib.lib().lpp_tridiag_eig(n, a.ctypes.data, bb.ctypes.data, eigs.ctypes.data,
                                     None if z is None else z.ctypes.data))
    return (eigs, z) if vectors else eigs


class ContinuedFraction:
    """PsimagLite::ContinuedFraction: set(ab, Eg, weight, isign) + evaluation (Engine.h:487-489)."""

    def __init__(self, a, b, eg, weight, isign):
        self.a, self.b, self.eg, self.weight, self.isign = _f64(a), _f64(b), float(eg), float(weight), int(isign)

    def __call__(self, omega, delta):
        omega = _f64(np.atleast_1d(omega))
        out = np.zeros(2 * omega.size)
        check(_lib.lib().lpp_cf_eval(self.a.size, self.a.ctypes.data, self.b.ctypes.data, self.eg, self.weight,
                                     self.isign, omega.size, omega.ctypes.data, delta, out.ctypes.data))
        return out[0::2] + 1j * out[1::2]


class Engine:
    """Engine.h:84-98: ground state on construction; spectralFunction() for the continued-fraction path."""

    def __init__(self, matrix, io=None, init=None):
        self.mat = matrix
        self.io = io or {}
        solver = LanczosSolver(matrix, ParametersForSolver(self.io, "Lanczos"))
        self.energy, _, self.a, self.b = solver.computeOneState(init, want_vector=True)

    def energies(self, ind=0):
        return self.energy

    RAHUL_LABELS = {"identity": 0, "n": 1, "sz": 2, "c": 3}

    def measure(self, ops):
        """Engine.h:208-249 for <gs|op_0[site_0];...|gs>: ops = [(label, dof, site[, transpose]), ...] with the labels of
        RahulOperator.h:56-63 ("c", "identity", "sz", "n"); the rightmost operator acts first (ModelBase::rahulMethod)."""
        n = len(ops)
        lab = np.array([self.RAHUL_LABELS[o[0]] for o in ops], dtype=np.int32)
        dof = np.array([o[1] for o in ops], dtype=np.int32)
        site = np.array([o[2] for o in ops], dtype=np.int32)
        tr = np.array([1 if (len(o) > 3 and o[3]) else 0 for o in ops], dtype=np.int32)
        out = C.c_double()
        check(_lib.lib().lpp_measure(self.mat.h, n, lab.ctypes.data, dof.ctypes.data, tr.ctypes.data, site.ctypes.data,
                                     C.byref(out)))
        return out.value

    def twoPoint(self, op, spin=0, orbs=(0, 0)):
        """Engine.h:262-331 with bra = ket = ground state: matrix result(i, j) = <O_j gs | O_i gs> (c: <cdagger_j c_i>)."""
        if op == OP_N:
            dst, own = self.mat, False
        else:
            dn = -1 if op == OP_C else 1
            dst, own = self.mat.sector(self.mat.nup + (dn if spin == 0 else 0), self.mat.ndown + (dn if spin == 1 else 0)), True
        n = self.mat.nsite
        out = np.zeros((n, n))
        check(_lib.lib().lpp_two_point(self.mat.h, dst.h, op, spin, orbs[0], orbs[1], out.ctypes.data))
        if own:
            dst.close()
        return out

    def manyPoint(self, sites, what, spins, orbs=None):
        """Engine.h:341-389 with bra = ket = ground state: <gs| O_n ... O_2 O_1 |gs>, O_k = what[k] at (sites[k], spins[k], orbs[k]),
        applied in the order given (what[0] first).  Every operator leads to the sector hasNewParts names (getNeededBasis,
        Engine.h:391-413); a string that leaves the allowed particle numbers or does not return to the ground state's sector gives 0
        like the reference.  The chain of modified vectors stays on the device (lpp_many_point)."""
        n = len(sites)
        orbs = [0] * n if orbs is None else list(orbs)
        chain, own = [self.mat], []
        nup, ndn = self.mat.nup, self.mat.ndown
        heis = self.mat.model == HEISENBERG
        nmax = self.mat.nsite * self.mat.orbitals
        try:
            for k in range(n):
                op, spin = what[k], spins[k]
                if op in (OP_C, OP_CDAGGER):
                    d = -1 if op == OP_C else 1
                    nup, ndn = (nup + d, ndn) if spin == 0 else (nup, ndn + d)
                elif op == OP_SPLUS:
                    nup, ndn = nup + 1, (ndn if heis else ndn - 1)
                elif op == OP_SMINUS:
                    nup, ndn = nup - 1, (ndn if heis else ndn + 1)
                if min(nup, ndn) < 0 or max(nup, ndn) > nmax:
                    return 0.0
                last = k == n - 1
                if (nup, ndn) == (chain[-1].nup, chain[-1].ndown):
                    chain.append(chain[-1])                      # sz, n: same sector, same handle
                elif last and (nup, ndn) == (self.mat.nup, self.mat.ndown):
                    chain.append(self.mat)
                else:
                    s = self.mat.sector(nup, ndn)
                    own.append(s)
                    chain.append(s)
            if (nup, ndn) != (self.mat.nup, self.mat.ndown):
                return 0.0
            arr = (C.c_void_p * (n + 1))(*[c.h for c in chain])
            i32 = lambda v: np.ascontiguousarray(v, dtype=np.int32)   # noqa: E731
            o, si, sp, ob = i32(what), i32(sites), i32(spins), i32(orbs)
            out = C.c_double()
            check(_lib.lib().lpp_many_point(arr, n, o.ctypes.data, si.ctypes.data, sp.ctypes.data, ob.ctypes.data, C.byref(out)))
            return out.value
        finally:
            for s in own:
                s.close()

    def spectralFunction(self, op, isite, jsite, spin=0, orbs=(0, 0)):
        """Engine.h:133-206: list of (type, ContinuedFraction).  Fermionic c/cdagger for HubbardOneBand, FeAsBasedSc (orbital
        pair) and Tj1Orbital; sz / splus / sminus for HubbardOneBand and Heisenberg (S(q, omega) building blocks)."""
        if spin != 0 and self.mat.nranks > 1:
            raise LppError("row-sharded spectral functions support spin 0")
        out = []
        is_diag = isite == jsite and orbs[0] == orbs[1]
        fermionic = op in (OP_C, OP_CDAGGER)                                 # LabeledOperator::isFermionic
        op2 = {OP_C: OP_CDAGGER, OP_CDAGGER: OP_C, OP_SPLUS: OP_SMINUS, OP_SMINUS: OP_SPLUS, OP_SZ: OP_SZ}[op]   # transposeConjugate
        nmax = self.mat.nsite * self.mat.orbitals
        for typ in range(4):
            if is_diag and typ > 1:
                continue
            lop = op if (typ & 1) else op2  # Engine.h:163
            nup, ndown = self.mat.nup, self.mat.ndown
            if lop in (OP_C, OP_CDAGGER):
                dn = -1 if lop == OP_C else 1
                nup, ndown = nup + (dn if spin == 0 else 0), ndown + (dn if spin == 1 else 0)
            elif lop in (OP_SPLUS, OP_SMINUS):      # HubbardOneOrbital.h:232-257 ; Heisenberg.h:218-240 (parts = (2S, Sz + const))
                c = 1 if lop == OP_SPLUS else -1
                nup += c
                if self.mat.model != HEISENBERG:
                    ndown -= c
            if nup < 0 or ndown < 0 or nup > nmax or ndown > nmax:
                continue
            if lop in (OP_C, OP_CDAGGER) and nup == 0 and ndown == 0:
                continue  # hasNewPartsCorCdagger: HubbardOneOrbital.h:212-230, BasisFeAsBasedSc.h:305-326, TjMultiOrb.h:538-557
            if self.mat.model == TJ and nup + ndown > self.mat.nsite:
                continue  # no double occupancy, TjMultiOrb.h:553
            if lop == OP_SZ:                        # needsNewBasis() is false: the sector is its own destination
                self._spectral_same_sector(out, typ, lop, isite, jsite, spin, orbs, is_diag)
                continue
            dst = self.mat.sector(nup, ndown)
            self.mat.apply_op(dst, lop, isite, spin, 1.0, accumulate=False, orb=orbs[0])   # Engine.h:509-517
            isign = -1.0 if typ > 1 else 1.0
            self.mat.apply_op(dst, lop, jsite, spin, isign, accumulate=True, orb=orbs[1])   # Engine.h:523-531
            solver = LanczosSolver(dst, ParametersForSolver(self.io, "Spectral"))
            a, b, weight = solver.decomposition(use_modified=True)                     # Engine.h:474-479
            s = -1 if (typ & 1) else 1
            s2 = -1.0 if typ > 1 else 1.0
            if not fermionic:
                s2 *= s                                                                # Engine.h:482
            s2 *= 1.0 if is_diag else 0.5                                              # Engine.h:481-485
            out.append((typ, ContinuedFraction(a, b, self.energy, weight * s2, -s)))   # Engine.h:489
            dst.close()
        return out

    def _spectral_same_sector(self, out, typ, lop, isite, jsite, spin, orbs, is_diag):
        self.mat.apply_op(self.mat, lop, isite, spin, 1.0, accumulate=False, orb=orbs[0])
        isign = -1.0 if typ > 1 else 1.0
        self.mat.apply_op(self.mat, lop, jsite, spin, isign, accumulate=True, orb=orbs[1])
        solver = LanczosSolver(self.mat, ParametersForSolver(self.io, "Spectral"))
        a, b, weight = solver.decomposition(use_modified=True)
        s = -1 if (typ & 1) else 1
        s2 = (-1.0 if typ > 1 else 1.0) * s * (1.0 if is_diag else 0.5)               # not fermionic: s2 *= s
        out.append((typ, ContinuedFraction(a, b, self.energy, weight * s2, -s)))
