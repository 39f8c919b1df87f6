"""Loader / builder of liblpp_b200.so (the C-ABI of include/lpp_b200.h).

The library is compiled IN-TREE with nvcc for sm_100a only; there is no CPU fallback: without the shared
library, or without a B200, every compute entry point raises.
"""
import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "liblpp_b200.so")
SOURCES = ["lpp_kernels.cu", "lpp_tiled.cu", "lpp_dtile.cu", "lpp_dblock.cu", "lpp_engine.cu"]
HEADERS = ["lpp_device.cuh", "lpp_kernels.cuh", "lpp_tiled.cuh", "lpp_dtile.cuh", "lpp_dblock.cuh", "lpp_dblock_kernel.cuh",
           "lpp_sweep_common.cuh", "lpp_smem_attr.cuh", "lpp_setup.h"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-shared"]


class LppError(RuntimeError):
    pass


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise LppError("nvcc not found: liblpp_b200.so cannot be built (there is no CPU fallback)")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(_HERE, "..", "include", "lpp_b200.h")]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into lanczosplusplus_b200/liblpp_b200.so."""
    if not force and not needs_build():
        return LIB_PATH
    extra = os.environ.get("LPP_NVCC_EXTRA", "").split()     # tuning experiments only (e.g. -DDT_THREADS=768)
    cmd = [_nvcc()] + NVCC_FLAGS + extra + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES] + ["-ldl"]
    if verbose:
        print(" ".join(cmd))
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if r.returncode != 0:
        raise LppError("nvcc failed:\n" + r.stdout)
    return LIB_PATH


DRIVER_SRC = os.path.join(_HERE, "..", "host", "lanczos_b200.cpp")
DRIVER_BIN = os.path.join(_HERE, "..", "host", "lanczos_b200")


def build_driver(force=False):
    """Compile the stand-alone C++ driver (host/lanczos_b200.cpp) against the in-tree shared library."""
    build()
    hdr = os.path.join(os.path.dirname(DRIVER_SRC), "engine_b200.h")
    if not force and os.path.exists(DRIVER_BIN) and os.path.getmtime(DRIVER_BIN) >= max(
            os.path.getmtime(DRIVER_SRC), os.path.getmtime(hdr), os.path.getmtime(LIB_PATH)):
        return DRIVER_BIN
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-I" + os.path.join(os.path.dirname(_HERE), "include"), "-o", DRIVER_BIN,
           DRIVER_SRC, "-L" + _HERE, "-llpp_b200",
           "-Wl,-rpath,$ORIGIN/../lanczosplusplus_b200"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise LppError("g++ failed:\n" + r.stdout)
    return DRIVER_BIN


def build_cf_collection(force=False):
    """Compile host/cf_collection.cpp (the evaluator for .comb files) against the in-tree shared library."""
    build()
    src = os.path.join(os.path.dirname(DRIVER_SRC), "cf_collection.cpp")
    hdr = os.path.join(os.path.dirname(DRIVER_SRC), "comb_io.h")
    exe = os.path.join(os.path.dirname(DRIVER_SRC), "cf_collection")
    if not force and os.path.exists(exe) and os.path.getmtime(exe) >= max(os.path.getmtime(src), os.path.getmtime(hdr),
                                                                         os.path.getmtime(LIB_PATH)):
        return exe
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-o", exe, src, "-L" + _HERE, "-llpp_b200",
           "-Wl,-rpath,$ORIGIN/../lanczosplusplus_b200"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise LppError("g++ failed:\n" + r.stdout)
    return exe


class Desc(C.Structure):
    _fields_ = [("model", C.c_int32), ("nsite", C.c_int32), ("orbitals", C.c_int32), ("nup", C.c_int32),
                ("ndown", C.c_int32), ("feas_u3_all_pairs", C.c_int32),
                ("hop", C.POINTER(C.c_double)), ("jzz", C.POINTER(C.c_double)),
                ("U", C.POINTER(C.c_double)), ("nU", C.c_int32),
                ("V", C.POINTER(C.c_double)), ("nV", C.c_int32),
                ("D", C.POINTER(C.c_double)), ("nD", C.c_int32),
                ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32),
                ("jpm", C.POINTER(C.c_double)), ("w", C.POINTER(C.c_double))]


class SolverParams(C.Structure):
    _fields_ = [("steps", C.c_int32), ("minsteps", C.c_int32), ("eps", C.c_double), ("kernel", C.c_int32),
                ("reortho", C.c_int32), ("seed", C.c_uint64)]


class Timing(C.Structure):
    _fields_ = [("spmv_ms", C.c_double), ("iter_ms", C.c_double), ("launches", C.c_int64)]


# every symbol include/lpp_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
_DP = C.POINTER(C.c_double)
SYMBOLS = {
    "lpp_last_error": (C.c_char_p, []),
    "lpp_version": (C.c_int, []),
    "lpp_device_check": (C.c_int, [C.c_int32]),
    "lpp_create": (C.c_int, [C.POINTER(Desc), C.POINTER(_VP)]),
    "lpp_destroy": (C.c_int, [_VP]),
    "lpp_rows": (C.c_int, [_VP, C.POINTER(C.c_uint64)]),
    "lpp_local_rows": (C.c_int, [_VP, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "lpp_basis_size": (C.c_int, [_VP, C.c_int32, C.POINTER(C.c_uint64)]),
    "lpp_basis_export": (C.c_int, [_VP, C.c_int32, _VP]),
    "lpp_rank": (C.c_int, [_VP, C.c_int32, _VP, C.c_uint64, _VP]),
    "lpp_states_below": (C.c_int, [_VP, _VP, _VP, C.c_int32, _VP, _VP, _VP]),
    "lpp_measure": (C.c_int, [_VP, C.c_int32, _VP, _VP, _VP, _VP, _VP]),
    "lpp_two_point": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _VP]),
    "lpp_many_point": (C.c_int, [_VP, C.c_int32, _VP, _VP, _VP, _VP, _VP]),
    "lpp_allreduce_selftest": (C.c_int, [_VP, _VP, C.c_int32]),
    "lpp_row_words": (C.c_int, [_VP, C.c_uint64, C.c_uint64, _VP, _VP]),
    "lpp_rank_pairs": (C.c_int, [_VP, _VP, _VP, C.c_uint64, _VP]),
    "lpp_matvec_host": (C.c_int, [_VP, C.c_int32, _VP, _VP]),
    "lpp_matvec_device": (C.c_int, [_VP, C.c_int32, _VP, _VP]),
    "lpp_crs_build": (C.c_int, [_VP, C.POINTER(C.c_int64)]),
    "lpp_crs_export": (C.c_int, [_VP, _VP, _VP, _VP]),
    "lpp_lanczos_decomposition": (C.c_int, [_VP, C.POINTER(SolverParams), _VP, C.c_int32, _VP, _VP,
                                            C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
    "lpp_ground_state": (C.c_int, [_VP, C.POINTER(SolverParams), _VP, C.c_int32, C.POINTER(C.c_double), _VP, _VP, _VP,
                                   C.POINTER(C.c_int32)]),
    "lpp_apply_op": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32]),
    "lpp_get_vector": (C.c_int, [_VP, C.c_int32, _VP]),
    "lpp_set_groundstate": (C.c_int, [_VP, _VP]),
    "lpp_cf_eval": (C.c_int, [C.c_int32, _VP, _VP, C.c_double, C.c_double, C.c_int32, C.c_int32, _VP, C.c_double, _VP]),
    "lpp_tridiag_eig": (C.c_int, [C.c_int32, _VP, _VP, _VP, _VP]),
    "lpp_comm_unique_id": (C.c_int, [_VP]),
    "lpp_comm_init": (C.c_int, [_VP, _VP]),
    "lpp_comm_share": (C.c_int, [_VP, _VP]),
    "lpp_p2p_export": (C.c_int, [_VP, C.c_int32, _VP]),
    "lpp_p2p_import": (C.c_int, [_VP, _VP]),
    "lpp_bench_spmv": (C.c_int, [_VP, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Timing)]),
    "lpp_bench_lanczos": (C.c_int, [_VP, C.POINTER(SolverParams), C.c_int32, C.c_int32, C.POINTER(Timing)]),
    "lpp_shard_range": (C.c_int, [C.c_uint64, C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
}

_lib = None


def lib():
    """The loaded C-ABI library. Raises LppError when it is missing (the product never falls back to the CPU)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LppError("%s is missing; run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(no CPU fallback exists)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise LppError("liblpp_b200 status %d: %s" % (rc, lib().lpp_last_error().decode()))
