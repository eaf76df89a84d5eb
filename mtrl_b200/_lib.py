"""ctypes binding of libmtrl_b200.so (the C-ABI declared in include/mtrl_b200.h).

There is no CPU fallback: if the library is missing or a call fails, the error is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_LIB_PATH = _PKG / "libmtrl_b200.so"
_lib = None


class MtrlError(RuntimeError):
    pass


class GemmProblem(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_longlong), ("a_major", C.c_int),
        ("B", C.c_void_p), ("ldb", C.c_longlong), ("b_major", C.c_int),
        ("D", C.c_void_p), ("ldd", C.c_longlong),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("block_n", C.c_int), ("k_splits", C.c_int), ("epilogue", C.c_int),
        ("bias", C.c_void_p), ("mask", C.c_void_p), ("ldmask", C.c_longlong),
        ("mask_bits", C.c_void_p), ("relu_bits_out", C.c_void_p), ("ldbits", C.c_longlong), ("colsum_partial", C.c_void_p),
        ("schedule_first", C.c_int), ("phase", C.c_int),
        ("A_lo", C.c_void_p), ("B_lo", C.c_void_p), ("D_lo", C.c_void_p),
        ("head_w", C.c_void_p), ("head_out", C.c_void_p), ("head_tile_task", C.c_void_p), ("head_dim", C.c_int),
    ]


EPI_STORE, EPI_BIAS_RELU, EPI_RELU_MASK, EPI_ATOMIC_ADD, EPI_STORE_TF32 = range(5)
GEMM_STREAMK = 16   # OR-ed into GemmPlan(ctas=...): stream-K schedule (MTRL_GEMM_STREAMK)
GEMM_ROWDEPS = 32   # ... phases ordered by per-row-tile dependencies instead of grid barriers (MTRL_GEMM_ROWDEPS)


ABI_VERSION = 6   # mtrl_abi_version() of the library these ctypes structures were written for


def lib() -> C.CDLL:
    """Load and return the library.  The build step runs first whenever nvcc is there: it is a digest comparison
    (sources + header + flags against build/stamp.txt) when the binary is current and a rebuild when it is stale, so a
    shipped .so can never run against newer struct layouts; ranks importing at once serialise on a file lock."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build

    if _build.have_nvcc():
        _build.build_library(force=bool(os.environ.get("MTRL_B200_REBUILD")))
    elif _LIB_PATH.exists() and not _build.is_current():
        raise MtrlError(f"{_LIB_PATH} was built from different sources than the ones in csrc/ and nvcc is not available to rebuild it")
    if not _LIB_PATH.exists():
        raise MtrlError(f"{_LIB_PATH} is missing: the CUDA extension was not built; there is no CPU fallback")
    l = C.CDLL(str(_LIB_PATH), mode=C.RTLD_GLOBAL)
    l.mtrl_last_error.restype = C.c_char_p
    l.mtrl_abi_version.restype = C.c_int
    if l.mtrl_abi_version() != ABI_VERSION:
        raise MtrlError(f"{_LIB_PATH} has ABI version {l.mtrl_abi_version()}, this package expects {ABI_VERSION}")
    _declare(l)
    _lib = l
    return l


def check(rc: int) -> None:
    if rc != 0:
        raise MtrlError(f"mtrl_b200 error {rc}: {lib().mtrl_last_error().decode()}")


def _declare(l: C.CDLL) -> None:
    vp, i, ll = C.c_void_p, C.c_int, C.c_longlong
    l.mtrl_gemm_plan_create.argtypes = [C.POINTER(vp), C.POINTER(GemmProblem), i]
    l.mtrl_gemm_plan_create_ex.argtypes = [C.POINTER(vp), C.POINTER(GemmProblem), i, i]
    l.mtrl_gemm_plan_run.argtypes = [vp, vp]
    l.mtrl_gemm_plan_units.argtypes = [vp]
    l.mtrl_gemm_plan_ctas.argtypes = [vp]
    l.mtrl_gemm_plan_set_debug.argtypes = [vp, vp]
    l.mtrl_gemm_plan_destroy.argtypes = [vp]
    l.mtrl_gemm_plan_destroy.restype = None
    for name, spec in _EXTRA_DECLS.items():
        _apply(l, name, spec)


def _apply(l: C.CDLL, name: str, spec) -> None:
    fn = getattr(l, name)
    fn.argtypes = spec[0]
    fn.restype = spec[1] if len(spec) > 1 else C.c_int


class _Decls(dict):
    """name -> (argtypes[, restype]).  Modules that own a group of C entry points register them here;
    a registration made after the library is already loaded is applied immediately (without argtypes
    ctypes would truncate 64-bit device pointers to int)."""

    def update(self, other=(), **kw):  # noqa: A003
        super().update(other, **kw)
        if _lib is not None:
            for name, spec in dict(other, **kw).items():
                _apply(_lib, name, spec)


_EXTRA_DECLS = _Decls()


def current_stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


class GemmPlan:
    """One persistent grouped launch; holds the encoded TMA descriptors."""

    def __init__(self, problems: list[GemmProblem], ctas: int = 0):
        arr = (GemmProblem * len(problems))(*problems)
        h = C.c_void_p()
        check(lib().mtrl_gemm_plan_create_ex(C.byref(h), arr, len(problems), ctas))
        self._h = h
        self.units = lib().mtrl_gemm_plan_units(h)
        self.ctas = lib().mtrl_gemm_plan_ctas(h)

    def set_debug(self, ptr: int | None) -> None:
        check(lib().mtrl_gemm_plan_set_debug(self._h, C.c_void_p(ptr)))

    def run(self, stream: int | None = None) -> None:
        check(lib().mtrl_gemm_plan_run(self._h, C.c_void_p(stream if stream is not None else current_stream_ptr())))

    def __del__(self):
        if getattr(self, "_h", None) is not None and _lib is not None:
            _lib.mtrl_gemm_plan_destroy(self._h)
            self._h = None
