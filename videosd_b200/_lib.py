"""ctypes loader for libvideosd.so (the C ABI declared in include/videosd.h).

There is deliberately no fallback: if the shared library is missing or an entry point fails, we raise.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvideosd.so")

_lib = None


class VsdError(RuntimeError):
    pass


def lib():
    """Returns the loaded CDLL, loading it on first use."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VsdError(
                f"{LIB_PATH} not found. Build it with `python -m videosd_b200.build` "
                "(or __graft_entry__.build()). There is no CPU / PyTorch fallback."
            )
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.vsd_last_error.restype = ctypes.c_char_p
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().vsd_last_error()
        raise VsdError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def _p(t):
    """Device/host pointer of a torch tensor (or None) as c_void_p."""
    if t is None:
        return ctypes.c_void_p(0)
    if isinstance(t, int):
        return ctypes.c_void_p(t)
    return ctypes.c_void_p(t.data_ptr())


def cur_stream():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def check_fault():
    check(lib().vsd_check_pipeline_fault(), "pipeline fault check")
