"""ctypes binding of libvqacl_b200.so (the C-ABI declared in include/vqacl_b200.h).

There is no CPU fallback: if the library is missing or a call fails, we raise.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvqacl_b200.so")

_lib = None


class VqaclError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VqaclError(
                f"{LIB_PATH} not found: build it with `python -m vqacl_b200.build` "
                "(there is no CPU/eager fallback for the VQACL hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.vqacl_last_error.restype = ctypes.c_char_p
    return _lib


def check(rc):
    if rc != 0:
        raise VqaclError(lib().vqacl_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device (or host) pointer of a torch tensor as c_void_p; None -> NULL."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
