"""Python-side wrappers over the C-ABI entry points (one function per `fd_*` symbol).

Every wrapper takes torch CUDA tensors only to obtain device pointers; all arithmetic
happens in the hand-written sm_100a kernels of libflowdec_b200.so.
"""
import ctypes

import torch

from . import _lib


def pack_conv_weight(segments, npad):
    """Pack conv weights for fd_conv2d_igemm.

    segments: list of (w[Cout, Cin_seg, kh, kw], taps) in the K order the kernel walks:
    for segment: for tap (kh-major): for channel.  Returns bf16 [npad, Ktot] (K-major),
    rows >= Cout zero-filled.
    """
    cols = []
    cout = segments[0][0].shape[0]
    for w, taps in segments:
        assert w.shape[0] == cout and w.shape[2] * w.shape[3] == taps
        # [Cout, Cin, kh, kw] -> [Cout, kh, kw, Cin] -> [Cout, taps*Cin]
        cols.append(w.permute(0, 2, 3, 1).reshape(cout, -1))
    wp = torch.cat(cols, dim=1).to(torch.bfloat16)
    if npad > cout:
        wp = torch.cat([wp, torch.zeros(npad - cout, wp.shape[1], dtype=wp.dtype, device=wp.device)], 0)
    return wp.contiguous()


def conv_igemm(srcs, wpacked, bias, out, max_ctas=0):
    """srcs: list of (tensor NHWC bf16, c_begin, c_count, taps). out: NHWC bf16 [.., npad] or fp32 [.., cout<=16]."""
    L = _lib.lib()
    n = len(srcs)
    arr = (_lib.ConvSrc * n)()
    B, H, W = srcs[0][0].shape[:3]
    for i, (t, c0, cc, taps) in enumerate(srcs):
        assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.shape[:3] == (B, H, W)
        arr[i].ptr = t.data_ptr()
        arr[i].C = t.shape[3]
        arr[i].c_begin = c0
        arr[i].c_count = cc
        arr[i].taps = taps
    npad, ktot = wpacked.shape
    out_f32 = out.dtype == torch.float32
    rc = L.fd_conv2d_igemm(arr, n, _lib.ptr(wpacked), ktot, _lib.ptr(bias), _lib.ptr(out),
                           int(out_f32), out.shape[3], npad, B, H, W, max_ctas, _lib.stream_ptr())
    _lib.check(rc, "fd_conv2d_igemm")
    return out
