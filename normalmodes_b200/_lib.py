"""ctypes binding of libnm_b200.so (include/nm_b200.h).  There is no CPU fallback: if the library is
missing it is built (nvcc), and if that fails the import raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnm_b200.so")
_lib = None

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
MATVEC_FN = C.CFUNCTYPE(None, c_double_p, c_double_p, C.c_void_p)


class NmError(RuntimeError):
    pass


def _preload_nccl():
    # the library links libnccl.so.2; make sure the copy torch uses (if any) is the one resolved
    try:
        import torch  # noqa: F401  (loads nvidia/nccl/lib/libnccl.so.2)
    except Exception:
        pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _b
        _b.build()
    _preload_nccl()
    try:
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    except OSError as e:
        raise NmError("cannot load %s: %s (normalmodes_b200 has no CPU fallback)" % (LIB_PATH, e))
    _lib.nm_last_error_message.restype = C.c_char_p
    _lib.nm_launch_count.restype = C.c_longlong
    _lib.nm_stream.restype = C.c_void_p
    return _lib


def check(rc):
    if rc != 0:
        raise NmError(lib().nm_last_error_message().decode())


def dptr(a):
    """double* of a C-contiguous float64 numpy array (or a raw device address given as int)."""
    if isinstance(a, (int, np.integer)):
        return C.cast(C.c_void_p(int(a)), c_double_p)
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# symbols declared in include/*.h (checked by the CPU test-suite against the built library)
def declared_symbols():
    import re
    inc = os.path.join(_HERE, "..", "include")
    names = []
    for f in sorted(os.listdir(inc)):
        txt = open(os.path.join(inc, f)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        for m in re.finditer(r"\b(nm_[a-z0-9_]+|pevsl_[a-z0-9_]+_)\s*\(", txt):
            if m.group(1) not in names:
                names.append(m.group(1))
    return names
