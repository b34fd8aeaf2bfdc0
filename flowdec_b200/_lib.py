"""ctypes binding of the C-ABI library `libflowdec_b200.so` (see include/flowdec_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, the
product path raises.  PyTorch only supplies device memory and the current stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libflowdec_b200.so")
_lib = None


class FlowDecNativeError(RuntimeError):
    pass


class ConvSrc(ctypes.Structure):
    """mirror of `struct fd_conv_src` (include/flowdec_b200.h)"""
    _fields_ = [
        ("ptr", ctypes.c_void_p),
        ("C", ctypes.c_int),
        ("c_begin", ctypes.c_int),
        ("c_count", ctypes.c_int),
        ("taps", ctypes.c_int),
    ]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise FlowDecNativeError(
                f"{_LIB_PATH} not found: build it with `python -m flowdec_b200.build` "
                "(flowdec_b200 has no CPU or PyTorch fallback path)")
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.fd_last_error.restype = ctypes.c_char_p
        _lib.fd_abi_version.restype = ctypes.c_int
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().fd_last_error().decode("utf-8", "replace")
        raise FlowDecNativeError(f"{what} failed (status {rc}): {msg}")


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())
