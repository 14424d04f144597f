"""ctypes binding of libos2s_b200.so (the C ABI declared in include/os2s.h).

There is deliberately no CPU fallback: if the shared library is missing or a call fails, an
exception is raised.  Only device pointers (torch tensors' data_ptr()) cross this boundary.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# OS2S_LIB_PATH: load another build of the same library (same-box A/B measurements of two builds)
LIB_PATH = os.environ.get("OS2S_LIB_PATH") or os.path.join(_HERE, "lib", "libos2s_b200.so")

_lib = None


class Os2sError(RuntimeError):
    pass


def load():
    """Load (building first if the in-tree .so is absent) and return the ctypes library."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    lib = ctypes.CDLL(LIB_PATH)
    lib.os2s_last_error.restype = ctypes.c_char_p
    lib.os2s_ctc_workspace_bytes.restype = ctypes.c_size_t
    _lib = lib
    return lib


def ptr(t):
    """Device/host pointer of a torch tensor (or None) as c_void_p."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(status, what=""):
    if status != 0:
        msg = load().os2s_last_error()
        raise Os2sError("%s failed (%d): %s" % (what, status, msg.decode() if msg else "?"))


def call(name, *args):
    lib = load()
    fn = getattr(lib, name)
    check(fn(*args), name)
