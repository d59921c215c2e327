"""ctypes binding of slim_b200/lib/libslim.so -- the same library the reference python-package
loads as site-packages/SLIM/libslim.so (python-package/SLIM/core.py:31-43), plus the SLIMB200_*
extension of include/slim_b200.h.  Loading fails loudly when the library has not been built;
nothing here falls back to a CPU implementation."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

LIB_PATH = Path(os.environ.get("SLIMB200_LIBRARY") or (Path(__file__).resolve().parent / "lib" / "libslim.so"))

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
c_ssp = C.POINTER(C.c_ssize_t)

_SIGNATURES = {
    # include/slim.h
    "SLIM_iSetDefaults": (C.c_int32, [c_i32p]),
    "SLIM_dSetDefaults": (C.c_int32, [c_f64p]),
    "SLIM_Learn": (C.c_void_p, [C.c_int32, c_ssp, c_i32p, c_f32p, c_i32p, c_f64p, C.c_void_p, c_i32p]),
    "SLIM_GetTopN": (C.c_int32, [C.c_void_p, C.c_int32, c_i32p, c_f32p, c_i32p, C.c_int32, c_i32p, c_f32p]),
    "SLIM_WriteModel": (C.c_int32, [C.c_void_p, C.c_char_p]),
    "SLIM_ReadModel": (C.c_void_p, [C.c_char_p]),
    "SLIM_FreeModel": (None, [C.POINTER(C.c_void_p)]),
    "SLIM_DetermineHeadAndTail": (c_i32p, [C.c_int32, C.c_int32, c_ssp, c_i32p]),
    "Py_csr_wrapper": (C.c_int32, [C.c_int32, c_ssp, c_i32p, c_f32p, C.POINTER(C.c_void_p)]),
    "Py_csr_save": (C.c_int32, [C.c_void_p, C.c_char_p]),
    "Py_csr_load": (C.c_int32, [C.POINTER(C.c_void_p), C.c_char_p]),
    "Py_csr_free": (C.c_int32, [C.c_void_p]),
    "Py_csr_stat": (C.c_int32, [C.c_void_p, c_i32p]),
    "Py_csr_export": (C.c_int32, [C.c_void_p, c_i32p, c_i32p, c_f32p]),
    "Py_SLIM_Learn": (C.c_int32, [C.c_void_p, c_i32p, c_f64p, C.POINTER(C.c_void_p)]),
    "Py_SLIM_Mselect": (C.c_int32, [C.c_void_p, C.c_void_p, c_i32p, c_f64p, c_f64p, c_f64p, C.c_int32,
                                    C.c_int32] + [c_f64p] * 8),
    "Py_SLIM_GetTopN": (C.c_int32, [C.c_void_p, C.c_int32, c_i32p, c_f32p, C.c_int32, c_i32p, c_f32p, C.c_int32]),
    "Py_SLIM_GetTopN_1vsk": (C.c_int32, [C.c_void_p, C.c_int32, c_i32p, c_f32p, C.c_int32, c_i32p, c_f32p,
                                         C.c_int32, c_i32p, C.c_int32]),
    "Py_SLIM_Predict_1vsk": (C.c_int32, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, c_i32p, c_i32p, c_f32p]),
    "Py_SLIM_Predict": (C.c_int32, [C.c_int32, C.c_void_p, C.c_void_p, c_i32p, c_f32p]),
    # include/slim_b200.h
    "SLIMB200_DeviceCount": (C.c_int32, []),
    "SLIMB200_LastError": (C.c_char_p, []),
    "SLIMB200_Stage": (C.c_void_p, [C.c_int32, C.c_int32, c_ssp, c_i32p, c_f32p, c_i32p]),
    "SLIMB200_StageDevice": (C.c_void_p, [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, c_i32p]),
    "SLIMB200_FreeMatrix": (None, [C.POINTER(C.c_void_p)]),
    "SLIMB200_MatrixInfo": (C.c_int32, [C.c_void_p, c_i32p, c_i32p, c_i64p, c_i32p, c_f64p, c_i32p]),
    "SLIMB200_MatrixCSC": (C.c_int32, [C.c_void_p, c_i64p, c_i32p, c_f32p, c_f32p]),
    "SLIMB200_MatrixItemOrder": (C.c_int32, [C.c_void_p, c_i32p]),
    "SLIMB200_MatrixWindowGram": (C.c_int32, [C.c_void_p, c_f64p]),
    "SLIMB200_MatrixGramInfo": (C.c_int32, [C.c_void_p, c_i32p, c_f64p]),
    "SLIMB200_MatrixGram": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "SLIMB200_MatrixGramLayout": (C.c_int32, [C.c_void_p, c_i64p, c_i32p, c_i32p]),
    "SLIMB200_MatrixGramStair": (C.c_int32, [C.c_void_p, c_i32p, c_i32p]),
    "SLIMB200_LearnColumns": (C.c_void_p, [C.c_void_p, c_i32p, c_f64p, c_i32p, C.c_int32, C.c_void_p, c_i32p]),
    "SLIMB200_FreeResult": (None, [C.POINTER(C.c_void_p)]),
    "SLIMB200_ResultInfo": (C.c_int32, [C.c_void_p, c_i32p, c_i64p, c_f64p, c_f64p, c_i32p]),
    "SLIMB200_ResultStats": (C.c_int32, [C.c_void_p, c_i32p, c_i32p, c_i64p, c_i64p, c_f64p, c_f64p]),
    "SLIMB200_ResultPhases": (C.c_int32, [C.c_void_p, c_f32p, c_i32p]),
    "SLIMB200_ResultToHost": (C.c_int32, [C.c_void_p, c_i64p, c_i32p, c_f32p]),
    "SLIMB200_ResultToDevice": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "SLIMB200_CommUniqueId": (C.c_int32, [C.c_void_p]),
    "SLIMB200_CommInitRank": (C.c_void_p, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, c_i32p]),
    "SLIMB200_CommFree": (None, [C.POINTER(C.c_void_p)]),
    "SLIMB200_AllGatherColumns": (C.c_void_p, [C.c_void_p, C.c_void_p, c_i32p, C.c_int32, c_i32p]),
    "SLIMB200_ResultToModel": (C.c_void_p, [C.c_void_p, c_i32p]),
    "SLIMB200_AssembleModel": (C.c_void_p, [C.c_int32, c_i64p, c_i32p, c_f32p, c_i32p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def _prefer_bundled_nccl():
    """libslim.so binds NCCL at run time (dlopen "libnccl.so.2").  In a python environment that also holds torch,
    point it at the NCCL wheel torch itself loads (site-packages/nvidia/nccl/lib/libnccl.so.2), so that the two
    never end up with different NCCL builds under one soname -- whichever of them is imported first."""
    if os.environ.get("SLIMB200_NCCL_LIBRARY"):
        return
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            cand = Path(base) / "lib" / "libnccl.so.2"
            if cand.exists():
                os.environ["SLIMB200_NCCL_LIBRARY"] = str(cand)
                return
    except Exception:
        pass


def load():
    """Load libslim.so (once) and attach the argument types of every exported entry point."""
    global _lib
    if _lib is None:
        _prefer_bundled_nccl()
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m slim_b200.build` "
                "(slim_b200 has no CPU fallback)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
