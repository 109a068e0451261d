"""ctypes front-end of the CPU parity oracle (TEST INFRASTRUCTURE ONLY).

Loads ``oracle/_build/liboracle.so`` (built by ``oracle/Makefile``) and exposes the restated
reference algorithm (see ``oracle/oracle.h``) as numpy/scipy objects.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module; the product library never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
MAXL = 32


class OCSR(C.Structure):
    _fields_ = [("nrows", C.c_int), ("ncols", C.c_int), ("ia", C.POINTER(C.c_int)),
                ("ja", C.POINTER(C.c_int)), ("a", C.POINTER(C.c_double))]


class Params(C.Structure):
    _fields_ = [("coarsen_type", C.c_int), ("strong_th", C.c_double), ("max_row_sum", C.c_double),
                ("max_coarse_size", C.c_int), ("min_coarse_size", C.c_int), ("max_levels", C.c_int),
                ("interp_type", C.c_int), ("max_nnz_row", C.c_int), ("trunc_factor", C.c_double),
                ("relax_down", C.c_int), ("relax_up", C.c_int), ("relax_coarse", C.c_int),
                ("sweeps_down", C.c_int), ("sweeps_up", C.c_int), ("sweeps_coarse", C.c_int),
                ("relax_weight", C.c_double), ("outer_weight", C.c_double), ("rand_seed", C.c_int)]


PO = C.POINTER(OCSR)
PD = C.POINTER(C.c_double)
PI = C.POINTER(C.c_int)


class OAMG(C.Structure):
    _fields_ = [("prm", Params), ("nlev", C.c_int),
                ("A", PO * MAXL), ("P", PO * MAXL), ("R", PO * MAXL), ("S", PO * MAXL),
                ("cf", PI * MAXL), ("cf_raw", PI * MAXL),
                ("l1_down", PD * MAXL), ("l1_up", PD * MAXL), ("measure", PD * MAXL),
                ("ge", PD), ("ge_n", C.c_int),
                ("u", PD * MAXL), ("f", PD * MAXL), ("t", PD * MAXL), ("t2", PD * MAXL)]


class Krylov(C.Structure):
    _fields_ = [("max_iter", C.c_int), ("rel_tol", C.c_double), ("abs_tol", C.c_double),
                ("krylov_dim", C.c_int), ("skip_real_res_check", C.c_int),
                ("iters", C.c_int), ("converged", C.c_int), ("rel_res_norm", C.c_double),
                ("hist", PD)]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (``make -C oracle``); returns the library path."""
    out = os.path.join(_HERE, "_build", "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(out)) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs)
    if force or stale:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE], check=True, env=env, stdout=subprocess.DEVNULL)
    return out


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.ocsr_from_arrays.restype = PO
        L.ocsr_from_arrays.argtypes = [C.c_int, C.c_int, PI, PI, PD]
        L.ocsr_free.argtypes = [PO]
        L.ocsr_diag_first.argtypes = [PO]
        L.ocsr_transpose.restype = PO
        L.ocsr_transpose.argtypes = [PO]
        L.ocsr_matvec.argtypes = [C.c_double, PO, PD, C.c_double, PD]
        L.ovec_dot.restype = C.c_double
        L.ovec_dot.argtypes = [C.c_int, PD, PD]
        L.oracle_rand_stream.argtypes = [C.c_int, C.c_int, PD]
        L.ovec_set_random.argtypes = [C.c_int, C.c_int, PD]
        L.oracle_set_threads.restype = C.c_int
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oamg_default_params.argtypes = [C.POINTER(Params), C.c_int]
        L.oamg_strength.restype = PO
        L.oamg_strength.argtypes = [PO, C.c_double, C.c_double]
        L.oamg_pmis.argtypes = [PO, C.c_int, C.c_int, PI, PD]
        L.oamg_rs_first_pass.argtypes = [PO, PI]
        L.oamg_extpi_interp.restype = PO
        L.oamg_extpi_interp.argtypes = [PO, PO, PI, C.c_int, C.c_double, PI]
        L.oamg_rap.restype = PO
        L.oamg_rap.argtypes = [PO, PO, PO]
        L.oamg_l1_norms.argtypes = [PO, C.c_int, PD]
        L.oamg_setup.restype = C.POINTER(OAMG)
        L.oamg_setup.argtypes = [PO, C.POINTER(Params)]
        L.oamg_destroy.argtypes = [C.POINTER(OAMG)]
        L.oamg_relax.argtypes = [PO, PD, PD, C.c_int, C.c_double, PD, PD, PD]
        L.oamg_vcycle.argtypes = [C.POINTER(OAMG), PD, PD]
        L.oamg_precond.argtypes = [C.POINTER(OAMG), PD, PD]
        L.opcg.argtypes = [PO, C.POINTER(OAMG), PD, PD, C.POINTER(Krylov)]
        L.ogmres.argtypes = [PO, C.POINTER(OAMG), PD, PD, C.POINTER(Krylov)]
        L.ofgmres.argtypes = [PO, C.POINTER(OAMG), PD, PD, C.POINTER(Krylov)]
        L.obicgstab.argtypes = [PO, C.POINTER(OAMG), PD, PD, C.POINTER(Krylov)]
        for g in ("ogen_laplace7", "ogen_laplace27"):
            getattr(L, g).restype = PO
            getattr(L, g).argtypes = [C.c_int] * 3 + [C.c_double] * 3
        L.ogen_convdif7.restype = PO
        L.ogen_convdif7.argtypes = [C.c_int] * 3 + [C.c_double] * 3
        L.ogen_rhs_yplane.argtypes = [C.c_int] * 3 + [PD]
        L.ogen_convdif7_rhs.argtypes = [C.c_int] * 3 + [C.c_double] * 3 + [PD]
    return _LIB


def _pd(x):
    return x.ctypes.data_as(PD)


def _pi(x):
    return x.ctypes.data_as(PI)


def to_ocsr(A: sp.csr_matrix):
    """scipy CSR -> oracle CSR (copies; column order preserved exactly as stored)."""
    ia = np.ascontiguousarray(A.indptr, dtype=np.int32)
    ja = np.ascontiguousarray(A.indices, dtype=np.int32)
    a = np.ascontiguousarray(A.data, dtype=np.float64)
    return lib().ocsr_from_arrays(A.shape[0], A.shape[1], _pi(ia), _pi(ja), _pd(a))


def from_ocsr(p, pattern_only: bool = False) -> sp.csr_matrix:
    """oracle CSR -> scipy CSR without re-sorting (storage order is part of the contract)."""
    o = p.contents
    n = o.nrows
    ia = np.ctypeslib.as_array(o.ia, shape=(n + 1,)).copy()
    nnz = int(ia[n])
    ja = np.ctypeslib.as_array(o.ja, shape=(max(nnz, 1),))[:nnz].copy()
    if pattern_only or not o.a:
        a = np.ones(nnz)
    else:
        a = np.ctypeslib.as_array(o.a, shape=(max(nnz, 1),))[:nnz].copy()
    M = sp.csr_matrix((n, o.ncols))
    M.indptr, M.indices, M.data = ia, ja, a
    M._shape = (n, o.ncols)
    return M


def default_params(gpu_defaults: bool = True, **kw) -> Params:
    p = Params()
    lib().oamg_default_params(C.byref(p), 1 if gpu_defaults else 0)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def gen(kind: str, nx: int, ny: int, nz: int, c=(1.0, 1.0, 1.0), diag_first: bool = True):
    """Reference example generators.  Returns (A scipy CSR in hypre storage order, b)."""
    L = lib()
    n = nx * ny * nz
    b = np.zeros(n)
    if kind == "lap7":
        p = L.ogen_laplace7(nx, ny, nz, *map(float, c))
        L.ogen_rhs_yplane(nx, ny, nz, _pd(b))
    elif kind == "lap27":
        p = L.ogen_laplace27(nx, ny, nz, *map(float, c))
        L.ogen_rhs_yplane(nx, ny, nz, _pd(b))
    elif kind == "convdif":
        kappa, umax, dt = c
        p = L.ogen_convdif7(nx, ny, nz, float(kappa), float(umax), float(dt))
        L.ogen_convdif7_rhs(nx, ny, nz, float(kappa), float(umax), float(dt), _pd(b))
    else:
        raise ValueError(kind)
    if diag_first:
        L.ocsr_diag_first(p)
    A = from_ocsr(p)
    L.ocsr_free(p)
    return A, b


def set_threads(n: int = 0) -> int:
    """Set (n > 0) and return the OpenMP team size of the oracle's parallel loops."""
    return int(lib().oracle_set_threads(int(n)))


def set_random(seed: int, n: int) -> np.ndarray:
    """HYPRE_ParVectorSetRandomValues(v, seed) on one rank."""
    out = np.zeros(n)
    lib().ovec_set_random(seed, n, _pd(out))
    return out


def rand_stream(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n)
    lib().oracle_rand_stream(seed, n, _pd(out))
    return out


class Hierarchy:
    """Owns an oracle AMG hierarchy; exposes per-level copies as scipy/numpy objects."""

    def __init__(self, A: sp.csr_matrix, params: Params | None = None):
        self.params = params if params is not None else default_params(True)
        self._A = to_ocsr(A)
        self._h = lib().oamg_setup(self._A, C.byref(self.params))
        self.n = A.shape[0]

    def __del__(self):
        try:
            if self._h:
                lib().oamg_destroy(self._h)
                self._h = None
            if self._A:
                lib().ocsr_free(self._A)
                self._A = None
        except Exception:
            pass

    @property
    def nlev(self) -> int:
        return self._h.contents.nlev

    def A(self, l):
        return from_ocsr(self._h.contents.A[l])

    def P(self, l):
        return from_ocsr(self._h.contents.P[l])

    def S(self, l):
        return from_ocsr(self._h.contents.S[l], pattern_only=True)

    def cf(self, l, raw=False):
        n = self._h.contents.A[l].contents.nrows
        arr = self._h.contents.cf_raw[l] if raw else self._h.contents.cf[l]
        return np.ctypeslib.as_array(arr, shape=(n,)).copy()

    def measure(self, l):
        n = self._h.contents.A[l].contents.nrows
        return np.ctypeslib.as_array(self._h.contents.measure[l], shape=(n,)).copy()

    def l1(self, l, up=False):
        n = self._h.contents.A[l].contents.nrows
        arr = self._h.contents.l1_up[l] if up else self._h.contents.l1_down[l]
        return np.ctypeslib.as_array(arr, shape=(n,)).copy()

    def sizes(self):
        out = []
        for l in range(self.nlev):
            o = self._h.contents.A[l].contents
            out.append((o.nrows, int(o.ia[o.nrows])))
        return out

    def precond(self, r: np.ndarray) -> np.ndarray:
        r = np.ascontiguousarray(r, dtype=np.float64)
        z = np.zeros_like(r)
        lib().oamg_precond(self._h, _pd(r), _pd(z))
        return z

    def vcycle(self, f: np.ndarray, u0: np.ndarray) -> np.ndarray:
        f = np.ascontiguousarray(f, dtype=np.float64)
        u = np.array(u0, dtype=np.float64, copy=True)
        lib().oamg_vcycle(self._h, _pd(f), _pd(u))
        return u


def _krylov(fn, A, b, x0, M, max_iter, rel_tol, abs_tol, krylov_dim, skip_real_res_check):
    pA = to_ocsr(A)
    lib().ocsr_diag_first(pA) if False else None
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
    hist = np.zeros(max_iter + 2)
    k = Krylov(max_iter=max_iter, rel_tol=rel_tol, abs_tol=abs_tol, krylov_dim=krylov_dim,
               skip_real_res_check=skip_real_res_check, hist=_pd(hist))
    fn(pA, M._h if M is not None else None, _pd(b), _pd(x), C.byref(k))
    lib().ocsr_free(pA)
    return x, dict(iters=k.iters, converged=bool(k.converged), rel_res_norm=k.rel_res_norm,
                   hist=hist[:k.iters + 1].copy())


def pcg(A, b, x0=None, M: Hierarchy | None = None, max_iter=100, rel_tol=1e-6, abs_tol=0.0):
    return _krylov(lib().opcg, A, b, x0, M, max_iter, rel_tol, abs_tol, 0, 0)


def gmres(A, b, x0=None, M: Hierarchy | None = None, max_iter=300, rel_tol=1e-6, abs_tol=0.0,
          krylov_dim=30, skip_real_res_check=0):
    return _krylov(lib().ogmres, A, b, x0, M, max_iter, rel_tol, abs_tol, krylov_dim,
                   skip_real_res_check)


def fgmres(A, b, x0=None, M: Hierarchy | None = None, max_iter=300, rel_tol=1e-6, abs_tol=0.0, krylov_dim=30):
    return _krylov(lib().ofgmres, A, b, x0, M, max_iter, rel_tol, abs_tol, krylov_dim, 0)


def bicgstab(A, b, x0=None, M: Hierarchy | None = None, max_iter=100, rel_tol=1e-6, abs_tol=0.0):
    return _krylov(lib().obicgstab, A, b, x0, M, max_iter, rel_tol, abs_tol, 0, 0)


def strength(A, theta=0.25, max_row_sum=0.9):
    pA = to_ocsr(A)
    pS = lib().oamg_strength(pA, theta, max_row_sum)
    S = from_ocsr(pS, pattern_only=True)
    lib().ocsr_free(pS)
    lib().ocsr_free(pA)
    return S


def matvec(A, x):
    pA = to_ocsr(A)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros(A.shape[0])
    lib().ocsr_matvec(1.0, pA, _pd(x), 0.0, _pd(y))
    lib().ocsr_free(pA)
    return y
