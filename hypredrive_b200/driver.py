"""Python host-side mirror of the reference's driver interface over the HYPREDRV_* C-ABI.

Same names and lifecycle as the reference's ``interfaces/python/src/driver.py`` (HypreDrive,
solve, set_matrix_from_csr, set_rhs, last_* properties, solution_norm) so that its tests read
the same here; everything goes through ctypes into ``libHYPREDRV.so`` -- the C entry points a
reference maintainer would bind (``interfaces/python/src/_core.pxd:19-58``).  There is no
Python arithmetic on the solve path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Any, Mapping, Optional, Sequence

import numpy as np

from . import hdk

BIGINT_DTYPE = np.dtype(np.int64)
REAL_DTYPE = np.dtype(np.float64)

ERROR_INVALID_VAL = 0x200


class HypreDriveError(RuntimeError):
    def __init__(self, code: int, where: str, messages: str = ""):
        self.code = int(code)
        super().__init__(f"{where} failed with error code 0x{self.code:08x}{(': ' + messages) if messages else ''}")


_api = None


def api():
    """HYPREDRV_* entry points of libHYPREDRV.so with argument types declared."""
    global _api
    if _api is None:
        L = hdk.lib()
        vp, u32 = C.c_void_p, C.c_uint32
        for name, args in {
            "HYPREDRV_Initialize": [], "HYPREDRV_Finalize": [],
            "HYPREDRV_Create": [C.c_int, C.POINTER(vp)], "HYPREDRV_Destroy": [C.POINTER(vp)],
            "HYPREDRV_SetLibraryMode": [vp],
            "HYPREDRV_InputArgsParse": [C.c_int, C.POINTER(C.c_char_p), vp],
            "HYPREDRV_InputArgsSetPreconPreset": [vp, C.c_char_p],
            "HYPREDRV_InputArgsSetSolverPreset": [vp, C.c_char_p],
            "HYPREDRV_LinearSystemSetMatrixFromCSR": [vp, C.c_longlong, C.c_longlong, vp, vp, vp],
            "HYPREDRV_LinearSystemSetRHSFromArray": [vp, C.c_longlong, C.c_longlong, vp],
            "HYPREDRV_LinearSystemSetStencil": [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                                C.c_longlong, C.c_longlong],
            "HYPREDRV_LinearSystemSetRHS": [vp, vp],
            "HYPREDRV_LinearSystemSetInitialGuess": [vp, vp],
            "HYPREDRV_LinearSystemResetInitialGuess": [vp],
            "HYPREDRV_LinearSystemSetPrecMatrix": [vp, vp],
            "HYPREDRV_LinearSystemGetSolutionValues": [vp, C.POINTER(C.POINTER(C.c_double))],
            "HYPREDRV_LinearSystemGetSolutionLength": [vp, C.POINTER(C.c_longlong)],
            "HYPREDRV_LinearSystemGetSolutionNorm": [vp, C.c_char_p, C.POINTER(C.c_double)],
            "HYPREDRV_LinearSystemGetDevicePointers": [vp, C.POINTER(vp), C.POINTER(vp)],
            "HYPREDRV_GetDeviceHandles": [vp, C.POINTER(vp), C.POINTER(vp)],
            "HYPREDRV_PreconCreate": [vp], "HYPREDRV_PreconSetup": [vp], "HYPREDRV_PreconDestroy": [vp],
            "HYPREDRV_LinearSolverCreate": [vp], "HYPREDRV_LinearSolverSetup": [vp],
            "HYPREDRV_LinearSolverApply": [vp], "HYPREDRV_LinearSolverDestroy": [vp],
            "HYPREDRV_LinearSolverGetNumIter": [vp, C.POINTER(C.c_int)],
            "HYPREDRV_LinearSolverGetConverged": [vp, C.POINTER(C.c_int)],
            "HYPREDRV_LinearSolverGetFinalRelativeResidualNorm": [vp, C.POINTER(C.c_double)],
            "HYPREDRV_LinearSolverGetSetupTime": [vp, C.POINTER(C.c_double)],
            "HYPREDRV_LinearSolverGetSolveTime": [vp, C.POINTER(C.c_double)],
            "HYPREDRV_StatsPrint": [vp], "HYPREDRV_AnnotateBegin": [vp, C.c_char_p, C.c_int],
            "HYPREDRV_AnnotateEnd": [vp, C.c_char_p, C.c_int],
        }.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = u32
        L.HYPREDRV_ErrorCodeClear.restype = None
        L.HYPREDRV_ErrorCodeDescribe.restype = None
        L.HYPREDRV_ErrorCodeDescribe.argtypes = [u32]
        _api = L
    return _api


_initialized = False


def initialize():
    global _initialized
    if not _initialized:
        _check(api().HYPREDRV_Initialize(), "HYPREDRV_Initialize")
        _initialized = True


def _check(code: int, where: str):
    if code:
        L = api()
        L.HYPREDRV_ErrorCodeDescribe(code)   # prints queued messages to stderr
        L.HYPREDRV_ErrorCodeClear()
        msg = hdk.lib().hdk_last_error().decode(errors="replace")
        raise HypreDriveError(code, where, msg)


def _emit_yaml(obj: Any, indent: int = 0) -> str:
    out = []
    pad = "  " * indent
    for k, v in obj.items():
        if isinstance(v, Mapping):
            out.append(f"{pad}{k}:")
            out.append(_emit_yaml(v, indent + 1))
        else:
            if isinstance(v, bool):
                v = "on" if v else "off"
            out.append(f"{pad}{k}: {v}")
    return "\n".join(x for x in out if x)


def normalize_options(options: Any) -> str:
    """dict / YAML text / path -> YAML text (reference interfaces/python/src/options.py)."""
    if options is None:
        return "solver: pcg\npreconditioner: amg\n"
    if isinstance(options, Mapping):
        opts = dict(options)
        opts.setdefault("preconditioner", "amg")
        return _emit_yaml(opts) + "\n"
    if isinstance(options, os.PathLike) or (isinstance(options, str) and "\n" not in options
                                            and ":" not in options and os.path.exists(options)):
        with open(options, "r", encoding="utf-8") as fh:
            return fh.read()
    if isinstance(options, str):
        return options if options.endswith("\n") else options + "\n"
    raise TypeError(f"unsupported options type {type(options).__name__}")


@dataclass
class SolveResult:
    x: np.ndarray
    solution_norm: float
    iterations: Optional[int] = None
    converged: Optional[bool] = None
    final_res_norm: Optional[float] = None
    setup_time: Optional[float] = None
    solve_time: Optional[float] = None


class HypreDrive:
    """Stateful driver wrapping one ``HYPREDRV_t`` handle in library mode."""

    def __init__(self, options: Any = None, comm: Any = None, input_args: Optional[Sequence[str]] = None):
        initialize()
        self._h = C.c_void_p()
        _check(api().HYPREDRV_Create(1, C.byref(self._h)), "HYPREDRV_Create")
        _check(api().HYPREDRV_SetLibraryMode(self._h), "HYPREDRV_SetLibraryMode")
        self._row_start = None
        self._row_end = None
        self._rhs_set = False
        self._reset_last()
        argv = [normalize_options(options).encode()] + [str(a).encode() for a in (input_args or [])]
        arr = (C.c_char_p * len(argv))(*argv)
        try:
            _check(api().HYPREDRV_InputArgsParse(len(argv), arr, self._h), "HYPREDRV_InputArgsParse")
        except Exception:
            self.close()
            raise

    def _reset_last(self):
        self._last_iterations = None
        self._last_converged = None
        self._last_final_res_norm = None
        self._last_setup_time = None
        self._last_solve_time = None

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        self.close()

    def close(self):
        if getattr(self, "_h", None):
            api().HYPREDRV_Destroy(C.byref(self._h))
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _require_open(self):
        if not self._h:
            raise RuntimeError("HypreDrive handle is closed")

    # ---- data ingest -----------------------------------------------------------------
    def set_matrix_from_csr(self, matrix_or_indptr, cols=None, data=None, *, row_start=None, row_end=None):
        self._require_open()
        if (row_start is None) != (row_end is None):
            raise TypeError("row_start and row_end must be provided together")
        if cols is None and data is None:
            m = matrix_or_indptr
            if isinstance(m, (tuple, list)) and len(m) == 3:
                indptr, cols, data = m
            else:
                m = m.tocsr()
                indptr, cols, data = m.indptr, m.indices, m.data
        else:
            indptr = matrix_or_indptr
        ip = np.ascontiguousarray(indptr, dtype=BIGINT_DTYPE)
        cj = np.ascontiguousarray(cols, dtype=BIGINT_DTYPE)
        va = np.ascontiguousarray(data, dtype=REAL_DTYPE)
        n = ip.size - 1
        if row_start is None:
            row_start, row_end = 0, n - 1
        if row_end - row_start + 1 != n:
            raise ValueError("indptr length does not match the row range")
        _check(api().HYPREDRV_LinearSystemSetMatrixFromCSR(self._h, row_start, row_end, ip.ctypes.data,
                                                           cj.ctypes.data, va.ctypes.data),
               "HYPREDRV_LinearSystemSetMatrixFromCSR")
        self._row_start, self._row_end = int(row_start), int(row_end)
        self._rhs_set = False

    def set_stencil(self, kind: int, nx: int, ny: int, nz: int, c=(1.0, 1.0, 1.0), row_start=0, row_end=None):
        """B200 extension: assemble a reference example stencil (7 / 27 / 107) and its RHS in HBM."""
        self._require_open()
        if row_end is None:
            row_end = nx * ny * nz - 1
        cc = (C.c_double * 3)(*map(float, c))
        _check(api().HYPREDRV_LinearSystemSetStencil(self._h, kind, nx, ny, nz, cc, row_start, row_end),
               "HYPREDRV_LinearSystemSetStencil")
        self._row_start, self._row_end = int(row_start), int(row_end)
        self._rhs_set = True

    def set_rhs(self, values, *, row_start=None, row_end=None):
        self._require_open()
        if self._row_start is None:
            raise RuntimeError("set_rhs(): no matrix set; call set_matrix_from_csr first")
        b = np.ascontiguousarray(values, dtype=REAL_DTYPE)
        rs = self._row_start if row_start is None else row_start
        re_ = self._row_end if row_end is None else row_end
        if b.size != re_ - rs + 1:
            raise ValueError("RHS length does not match the row range")
        _check(api().HYPREDRV_LinearSystemSetRHSFromArray(self._h, rs, re_, b.ctypes.data),
               "HYPREDRV_LinearSystemSetRHSFromArray")
        self._rhs_set = True

    # ---- solve -----------------------------------------------------------------------
    def solve(self):
        self._require_open()
        if self._row_start is None:
            raise RuntimeError("solve(): no matrix set; call set_matrix_from_csr")
        if not self._rhs_set:
            raise RuntimeError("solve(): no RHS set; call set_rhs")
        self._reset_last()
        L = api()
        _check(L.HYPREDRV_LinearSystemSetInitialGuess(self._h, None), "HYPREDRV_LinearSystemSetInitialGuess")
        _check(L.HYPREDRV_LinearSystemResetInitialGuess(self._h), "HYPREDRV_LinearSystemResetInitialGuess")
        _check(L.HYPREDRV_LinearSolverCreate(self._h), "HYPREDRV_LinearSolverCreate")
        try:
            _check(L.HYPREDRV_LinearSolverSetup(self._h), "HYPREDRV_LinearSolverSetup")
            _check(L.HYPREDRV_LinearSolverApply(self._h), "HYPREDRV_LinearSolverApply")
            it, cv, d = C.c_int(), C.c_int(), C.c_double()
            _check(L.HYPREDRV_LinearSolverGetNumIter(self._h, C.byref(it)), "GetNumIter")
            self._last_iterations = it.value
            _check(L.HYPREDRV_LinearSolverGetConverged(self._h, C.byref(cv)), "GetConverged")
            self._last_converged = bool(cv.value)
            _check(L.HYPREDRV_LinearSolverGetFinalRelativeResidualNorm(self._h, C.byref(d)), "GetFinalRelRes")
            self._last_final_res_norm = d.value
            _check(L.HYPREDRV_LinearSolverGetSetupTime(self._h, C.byref(d)), "GetSetupTime")
            self._last_setup_time = d.value
            _check(L.HYPREDRV_LinearSolverGetSolveTime(self._h, C.byref(d)), "GetSolveTime")
            self._last_solve_time = d.value
        finally:
            L.HYPREDRV_LinearSolverDestroy(self._h)
            L.HYPREDRV_ErrorCodeClear() if False else None

    last_iterations = property(lambda self: self._last_iterations)
    last_converged = property(lambda self: self._last_converged)
    last_final_res_norm = property(lambda self: self._last_final_res_norm)
    last_setup_time = property(lambda self: self._last_setup_time)
    last_solve_time = property(lambda self: self._last_solve_time)

    def get_solution(self) -> np.ndarray:
        self._require_open()
        p = C.POINTER(C.c_double)()
        n = C.c_longlong()
        _check(api().HYPREDRV_LinearSystemGetSolutionValues(self._h, C.byref(p)), "GetSolutionValues")
        _check(api().HYPREDRV_LinearSystemGetSolutionLength(self._h, C.byref(n)), "GetSolutionLength")
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def solution_norm(self, kind: str = "l2") -> float:
        self._require_open()
        key = {"l1": b"L1", "l2": b"L2", "inf": b"inf", "linf": b"inf"}.get(str(kind).lower())
        if key is None:
            raise ValueError("kind must be one of 'l1', 'l2', 'inf'")
        d = C.c_double()
        _check(api().HYPREDRV_LinearSystemGetSolutionNorm(self._h, key, C.byref(d)), "GetSolutionNorm")
        return d.value

    def stats_print(self):
        _check(api().HYPREDRV_StatsPrint(self._h), "HYPREDRV_StatsPrint")

    def device_handles(self):
        """(hdk_csr*, hdk_amg*) of the installed system -- roofline instrumentation only."""
        a, m = C.c_void_p(), C.c_void_p()
        _check(api().HYPREDRV_GetDeviceHandles(self._h, C.byref(a), C.byref(m)), "GetDeviceHandles")
        return a, m


def solve(A, b, options=None, comm=None, *, row_start=None, row_end=None, input_args=None) -> SolveResult:
    if (row_start is None) != (row_end is None):
        raise TypeError("row_start and row_end must be provided together")
    with HypreDrive(options=options, comm=comm, input_args=input_args) as drv:
        if row_start is None:
            drv.set_matrix_from_csr(A)
        else:
            drv.set_matrix_from_csr(A, row_start=row_start, row_end=row_end)
        drv.set_rhs(b, row_start=row_start, row_end=row_end)
        drv.solve()
        x = drv.get_solution()
        return SolveResult(x=x, solution_norm=drv.solution_norm("l2"), iterations=drv.last_iterations,
                           converged=drv.last_converged, final_res_norm=drv.last_final_res_norm,
                           setup_time=drv.last_setup_time, solve_time=drv.last_solve_time)
