"""ctypes binding of the C-ABI library `libflowdec_b200.so` (see include/flowdec_b200.h).

There is deliberately no fallback: if the library is missing or a call fails, the
product path raises.  PyTorch only supplies device memory and the current stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("FD_LIB_PATH") or os.path.join(_HERE, "libflowdec_b200.so")   # FD_LIB_PATH: A/B builds
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
        ("scale_shift", ctypes.c_void_p),
        ("ss_pitch", ctypes.c_int),
    ]


_P = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float
_D = ctypes.c_double
_Z = ctypes.c_size_t

# One entry per `extern "C"` function declared in include/flowdec_b200.h (same order).
SIGNATURES = {
    "fd_abi_version": [],
    "fd_conv2d_igemm": [ctypes.POINTER(ConvSrc), _I, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P],
    "fd_conv_cluster4": [_I],
    "fd_chan_stats": [_P, _I, _I, _I, _P, _I, _P],
    "fd_slab_reduce": [_P, _I, _I, _I, _P, _I, _P],
    "fd_gn_finalize": [_P, _I, _I, _P, _I, _I, _I, _D, _P, _P, _I, _F, _P, _P],
    "fd_gn_act_resample": [_P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _I, _P],
    "fd_pack4": [_P, _P, _P, _Z, _P],
    "fd_fir_down4": [_P, _P, _I, _I, _I, _P],
    "fd_pyramid_up_add": [_P, _P, _P, _I, _I, _I, _P],
    "fd_pyramid_gather": [_P, _I, _P, _P, _P, _I, _I, _I, _P],
    "fd_conv_in": [_P, _P, _P, _P, _I, _I, _I, _P],
    "fd_combine": [_P, _P, _P, _P, _P, _Z, _I, _P],
    "fd_output_axpy": [_P, ctypes.POINTER(ctypes.c_float), _P, _F, _P, _F, _P, _F, _F, _P, _P, _Z, _P],
    "fd_x0": [_P, _P, _P, _F, _P, _I, _I, _I, _P],
    "fd_fourier_embed": [_F, _P, _I, _P, _P],
    "fd_matvec": [_P, _I, _I, _P, _P, _P, _F, _P, _I, _P],
    "fd_twiddles1534": [_P, _P],
    "fd_normfac": [_P, _I, _I, _I, _P, _P],
    "fd_stft1534_compress": [_P, _I, _I, _P, _P, _P, _F, _F, _I, _P, _P],
    "fd_istft1534_decompress": [_P, _I, _I, _I, _P, _P, _P, _F, _F, _P, _P],
    "fd_normfac_ragged": [_P, _I, _I, _P, _I, _P, _P],
    "fd_stft1534_compress_ragged": [_P, _I, _I, _P, _P, _P, _P, _F, _F, _I, _P, _P],
    "fd_istft1534_decompress_ragged": [_P, _I, _I, _I, _P, _P, _P, _P, _F, _F, _P, _P],
    "fd_rvq_from_codes": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "fd_dac_conv1d": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "fd_dac_conv_transpose1d": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "fd_dac_conv1d_strided": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "fd_rvq_encode": [_P] * 13 + [_I] * 6 + [_P],
    "fd_fir_tiles_enable": [_I],
    "fd_chan_stats_f32": [_P, _I, _I, _I, _P, _I, _P],
    "fd_gn_act_resample_f32": [_P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _I, _P],
    "fd_conv_in_f32": [_P, _P, _P, _P, _I, _I, _I, _P],
    "fd_combine_f32": [_P, _P, _P, _P, _P, _Z, _I, _P],
    "fd_dac_conv_tc": [_P, _I, _I, ctypes.c_longlong, _I, _P, _I, _I, ctypes.POINTER(ctypes.c_int), _P, _P,
                       ctypes.c_longlong, _P, _I, _P, _P, _I, ctypes.c_longlong, _P],
    "fd_dac_final_conv": [_P, ctypes.c_longlong, _P, _P, _P, _I, _I, _I, _P],
    "fd_dac_nct_to_ntc": [_P, _P, _I, _I, _I, _I, _P],
    "fd_dac_first_conv": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "fd_stft_use_pfa": [_I],
    "fd_istft1534_decompress_pfa": [_P, _I, _I, _I, _P, _P, _P, _P, _F, _F, _P, _P, _P],
    "fd_upfirdn2d_f32": [_P, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "fd_conv2d_direct": [ctypes.POINTER(ConvSrc), _I, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "fd_attention": [_P, _I, _I, _I, _F, _P, _P],
    "fd_gn_act_down_any": [_P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _P],
    "fd_conv_in_any": [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    "fd_output_conv3_axpy": [_P, _P, _P, _F, _P, _F, _P, _F, _F, _P, _P, _I, _I, _I, _P],
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise FlowDecNativeError(
                f"{_LIB_PATH} not found: build it with `python -m flowdec_b200.build` "
                "(flowdec_b200 has no CPU or PyTorch fallback path)")
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.fd_last_error.restype = ctypes.c_char_p
        for name, argtypes in SIGNATURES.items():
            fn = getattr(_lib, name)          # AttributeError here = header / library mismatch
            fn.restype = ctypes.c_int
            fn.argtypes = argtypes
        if os.environ.get("FD_CONV_CLUSTER4") is not None:      # A/B switch for tools / bench
            _lib.fd_conv_cluster4(int(os.environ["FD_CONV_CLUSTER4"]))
    return _lib


LAUNCHES = 0   # number of successful fd_* kernel-launching calls (bench.py's gpu_launches)


def check(rc, what):
    global LAUNCHES
    LAUNCHES += 1
    if rc != 0:
        msg = lib().fd_last_error().decode("utf-8", "replace")
        raise FlowDecNativeError(f"{what} failed (status {rc}): {msg}")


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return ctypes.c_void_p(0)
    if not t.is_cuda:
        # a host address must never reach a kernel (it faults, or on HMM boxes silently streams over PCIe)
        raise FlowDecNativeError(f"expected a CUDA tensor, got a {t.device} tensor of shape {tuple(t.shape)}")
    return ctypes.c_void_p(t.data_ptr())
