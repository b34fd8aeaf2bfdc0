"""Python-side wrappers over the C-ABI entry points (one function per `fd_*` symbol).

Every wrapper takes torch CUDA tensors only to obtain device pointers; all arithmetic
happens in the hand-written sm_100a kernels of libflowdec_b200.so.
"""
import ctypes

import torch

from . import _lib

# CTA-pair (cta_group::2) convolution tiles; tests flip this to compare both MMA variants
CTA_PAIRS = True
# 16x8 "halo" tiles (one A box per k-slice serves all nine taps; needed for fused GroupNorm+SiLU operands).
# False selects the per-tap kernel (one A box per tap and k-slice).
HALO_TILES = True


def halo_eligible(B, H, W, npad):
    """mirror of the dispatch rule in fd_conv2d_igemm (npad 16 = the 4-channel fp32 pyramid form)"""
    return (CTA_PAIRS and HALO_TILES and npad in (16, 48, 128, 256) and W % 8 == 0 and H % 16 == 0
            and (B * (H // 16) * (W // 8)) % 2 == 0)


def round_tf32(w):
    """fp32 -> nearest tf32 (10-bit mantissa) kept in fp32 words; weights are rounded once at pack time, the
    tensor pipe would otherwise truncate them"""
    i = w.float().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)

# when set to a list, conv_igemm appends (start_event, end_event, algorithmic_flops) per launch
# (bench.py's roofline leg; events are recorded on the launching stream)
PROFILE = None


def pack_conv_weight(segments, npad, tf32=False):
    """Pack conv weights for fd_conv2d_igemm.

    segments: list of (w[Cout, Cin_seg, kh, kw], taps) in the K order the kernel walks:
    for segment: for tap (kh-major): for channel.  Returns bf16 (or, tf32=True, tf32-rounded fp32)
    [npad, Ktot] (K-major), rows >= Cout zero-filled.
    """
    cols = []
    cout = segments[0][0].shape[0]
    for w, taps in segments:
        assert w.shape[0] == cout and w.shape[2] * w.shape[3] == taps
        # [Cout, Cin, kh, kw] -> [Cout, kh, kw, Cin] -> [Cout, taps*Cin]
        cols.append(w.permute(0, 2, 3, 1).reshape(cout, -1))
    wp = torch.cat(cols, dim=1)
    wp = round_tf32(wp.float()) if tf32 else wp.to(torch.bfloat16)
    if npad > cout:
        wp = torch.cat([wp, torch.zeros(npad - cout, wp.shape[1], dtype=wp.dtype, device=wp.device)], 0)
    return wp.contiguous()


def conv_stats_slabs(H, W):
    """number of partial-sum slabs the conv epilogue writes per sample: one per 128-pixel tile"""
    return H * W // 128


def conv_igemm(srcs, wpacked, bias, out, max_ctas=0, algo_k=None, stats=None, algo_cout=None):
    """srcs: list of (tensor NHWC bf16, c_begin, c_count, taps[, scale_shift, channel_offset]).
    With scale_shift (fp32 [B, Cvirt, 2]) the kernel consumes SiLU(x*scale+shift) instead of x.
    out: NHWC bf16 [.., npad] or fp32 [.., cout]."""
    L = _lib.lib()
    arr, n, B, H, W = _fill_srcs(srcs)
    npad, ktot = wpacked.shape
    out_f32 = out.dtype == torch.float32
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    tf32 = srcs[0][0].dtype == torch.float32          # fp32 activations + tf32-rounded fp32 weights
    assert (wpacked.dtype == torch.float32) == tf32
    rc = L.fd_conv2d_igemm(arr, n, _lib.ptr(wpacked), ktot, _lib.ptr(bias), _lib.ptr(out),
                           int(out_f32), out.shape[3], npad, B, H, W, _lib.ptr(stats), max_ctas,
                           int(CTA_PAIRS) | (2 if HALO_TILES else 0) | (4 if tf32 else 0), _lib.stream_ptr())
    _lib.check(rc, "fd_conv2d_igemm")
    if PROFILE is not None:
        e1.record()
        # algorithmic FLOPs: real output channels, and K without the identity-skip segment that
        # stands in for the res-block's "+ x" (an implementation device, not reference arithmetic)
        k_algo = sum(src[2] * src[3] for src in srcs)
        if algo_k is not None:
            k_algo = algo_k
        PROFILE.append((e0, e1, 2.0 * B * H * W * (algo_cout or out.shape[3]) * k_algo))
    return out


def conv_cluster4(on):
    """4-CTA clusters with weight multicast for the bf16 halo convs on / off; returns the previous setting"""
    return bool(_lib.lib().fd_conv_cluster4(int(bool(on))))


def tensor_conv_ok(B, H, W, npad, seg_channels, out_f32=False):
    """can fd_conv2d_igemm (tcgen05 tiles) take this conv?  Otherwise conv_direct (CUDA cores) runs it."""
    if W % 8 or W < 8 or any(c % 64 for c in seg_channels):
        return False
    bw = 8
    while bw < 128 and W % (bw * 2) == 0:
        bw *= 2
    if H % (128 // bw):
        return False
    return npad in ((16, 48) if out_f32 else (128, 256))


def _fill_srcs(srcs):
    n = len(srcs)
    arr = (_lib.ConvSrc * n)()
    B, H, W = srcs[0][0].shape[:3]
    for i, src in enumerate(srcs):
        t, c0, cc, taps = src[:4]
        assert t.is_cuda and t.dtype == srcs[0][0].dtype and t.dtype in (torch.bfloat16, torch.float32)
        assert t.is_contiguous() and t.shape[:3] == (B, H, W)
        arr[i].ptr = t.data_ptr()
        arr[i].C = t.shape[3]
        arr[i].c_begin = c0
        arr[i].c_count = cc
        arr[i].taps = taps
        if len(src) > 4 and src[4] is not None:
            ss, ch_off = src[4], src[5]          # fp32 [B, Cvirt, 2], channel offset of this source in it
            assert ss.is_cuda and ss.dtype == torch.float32 and ss.is_contiguous() and ss.shape[0] == B
            arr[i].scale_shift = ss.data_ptr() + ch_off * 8
            arr[i].ss_pitch = ss.shape[1]
        else:
            arr[i].scale_shift = None
            arr[i].ss_pitch = 0
    return arr, n, B, H, W


def conv_direct(srcs, wpacked, bias, out, cout=None, affine_only=False):
    """Shape-generic conv (csrc/fd_generic.cu): same source tuples / packed weights as conv_igemm; any H, W,
    channel counts multiples of 8.  out: bf16 or fp32 NHWC [B,H,W,pitch >= cout]."""
    arr, n, B, H, W = _fill_srcs(srcs)
    rows, ktot = wpacked.shape
    cout = out.shape[3] if cout is None else cout
    assert rows >= cout and out.shape[:3] == (B, H, W)
    rc = _lib.lib().fd_conv2d_direct(arr, n, _lib.ptr(wpacked), ktot, _lib.ptr(bias), _lib.ptr(out),
                                     int(out.dtype == torch.float32), cout, out.shape[3], B, H, W,
                                     int(bool(affine_only)), _lib.stream_ptr())
    _lib.check(rc, "fd_conv2d_direct")
    return out


def attention(qkv, out):
    """qkv fp32 [B,H,W,3C] (q | k | v), out bf16 [B,H,W,C]: softmax over all H*W positions (AttnBlockpp)"""
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    rc = _lib.lib().fd_attention(_lib.ptr(qkv), B, H * W, C, ctypes.c_float(float(C) ** -0.5), _lib.ptr(out),
                                 _lib.stream_ptr())
    _lib.check(rc, "fd_attention")
    return out


def upfirdn2d(x, kernel, up=1, down=1, pad=(0, 0)):
    """The reference op `upfirdn2d(input, kernel, up, down, pad)` (op/upfirdn2d.py:169-180) on fp32 NCHW CUDA
    tensors through fd_upfirdn2d_f32; allocates and returns the output like the reference."""
    up_x = up_y = int(up)
    down_x = down_y = int(down)
    px0, px1, py0, py1 = int(pad[0]), int(pad[1]), int(pad[0]), int(pad[1])
    N, C, H, W = x.shape
    kh, kw = kernel.shape
    xc = x.float().contiguous()
    kc = kernel.to(x.device, torch.float32).contiguous()
    oh = (H * up_y + py0 + py1 - kh) // down_y + 1
    ow = (W * up_x + px0 + px1 - kw) // down_x + 1
    out = torch.empty(N, C, oh, ow, device=x.device, dtype=torch.float32)
    rc = _lib.lib().fd_upfirdn2d_f32(_lib.ptr(xc), N * C, H, W, _lib.ptr(kc), kh, kw, up_x, up_y, down_x, down_y,
                                     px0, px1, py0, py1, _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "fd_upfirdn2d_f32")
    return out


# ---------------------------------------------------------------------------------------------
# HBM-bound side kernels (csrc/fd_elementwise.cu)
# ---------------------------------------------------------------------------------------------
def chan_stats(x, slabs, out=None):
    """x: bf16 NHWC [B,H,W,C] -> fp32 partial sums [B,S,C,2] (sum, sum of squares per slab)."""
    B, H, W, C = x.shape
    if out is None:
        out = torch.empty(B, slabs, C, 2, device=x.device, dtype=torch.float32)
    fn = _lib.lib().fd_chan_stats_f32 if x.dtype == torch.float32 else _lib.lib().fd_chan_stats
    _lib.check(fn(_lib.ptr(x), B, H * W, C, _lib.ptr(out), slabs, _lib.stream_ptr()), "fd_chan_stats")
    return out


def slab_reduce(part, chunks, out):
    """part fp32 [B,S,C,2] -> out fp32 [B,chunks,C,2]"""
    B, S, C, _ = part.shape
    _lib.check(_lib.lib().fd_slab_reduce(_lib.ptr(part), B, S, C, _lib.ptr(out), chunks, _lib.stream_ptr()),
               "fd_slab_reduce")
    return out


def gn_finalize(parts, chans, count, gamma, beta, groups, eps, out):
    """parts: 1 or 2 partial-sum tensors [B,Si,Ci,2] forming the virtual concat. out: fp32 [B,C,2]."""
    p1 = parts[0]
    p2 = parts[1] if len(parts) > 1 else None
    c2 = chans[1] if len(parts) > 1 else 0
    s2 = p2.shape[1] if p2 is not None else 0
    B = p1.shape[0]
    rc = _lib.lib().fd_gn_finalize(_lib.ptr(p1), chans[0], p1.shape[1], _lib.ptr(p2), c2, s2, B,
                                   ctypes.c_double(count), _lib.ptr(gamma), _lib.ptr(beta), groups,
                                   ctypes.c_float(eps), _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "fd_gn_finalize")
    return out


def gn_act_resample(srcs, scale_shift, out, mode, out_raw=None):
    """srcs: 1 or 2 bf16 NHWC tensors (virtual concat). mode 0/1/2 = same/down/up.
    out: SiLU(GroupNorm) output (or None); out_raw: FIR of the raw input (or None)."""
    s1 = srcs[0]
    s2 = srcs[1] if len(srcs) > 1 else None
    B, H, W, C1 = s1.shape
    C2 = s2.shape[3] if s2 is not None else 0
    if s1.dtype == torch.float32:
        rc = _lib.lib().fd_gn_act_resample_f32(_lib.ptr(s1), C1, _lib.ptr(s2), C2, _lib.ptr(scale_shift),
                                               _lib.ptr(out), _lib.ptr(out_raw), B, H, W, mode, _lib.stream_ptr())
        _lib.check(rc, "fd_gn_act_resample_f32")
        return out
    if mode == 1 and (H % 4 or W % 4):
        rc = _lib.lib().fd_gn_act_down_any(_lib.ptr(s1), C1, _lib.ptr(s2), C2, _lib.ptr(scale_shift), _lib.ptr(out),
                                           _lib.ptr(out_raw), B, H, W, _lib.stream_ptr())
        _lib.check(rc, "fd_gn_act_down_any")
        return out
    rc = _lib.lib().fd_gn_act_resample(_lib.ptr(s1), C1, _lib.ptr(s2), C2, _lib.ptr(scale_shift),
                                       _lib.ptr(out), _lib.ptr(out_raw), B, H, W, mode, _lib.stream_ptr())
    _lib.check(rc, "fd_gn_act_resample")
    return out


def fir_tiles_enable(on):
    """TMA-tiled up / down kernels (fd_fir_tiles.cu) on/off; returns the previous setting"""
    return bool(_lib.lib().fd_fir_tiles_enable(int(bool(on))))


def pack4(x, y, out):
    n = x.numel() // 2
    _lib.check(_lib.lib().fd_pack4(_lib.ptr(x), _lib.ptr(y), _lib.ptr(out), ctypes.c_size_t(n),
                                   _lib.stream_ptr()), "fd_pack4")
    return out


def fir_down4(x, out):
    B, H, W, _ = x.shape
    _lib.check(_lib.lib().fd_fir_down4(_lib.ptr(x), _lib.ptr(out), B, H, W, _lib.stream_ptr()), "fd_fir_down4")
    return out


def pyramid_up_add(lo, add, out):
    B, H, W, _ = lo.shape
    _lib.check(_lib.lib().fd_pyramid_up_add(_lib.ptr(lo), _lib.ptr(add), _lib.ptr(out), B, H, W,
                                            _lib.stream_ptr()), "fd_pyramid_up_add")
    return out


def pack_tap_weight(w, npad=48, tf32=False):
    """[Cout(4), Cin, 3, 3] -> bf16 (or tf32-rounded fp32) [npad, Cin]: row tap*4+co = w[co, :, kh, kw] (tap = kh*3+kw)"""
    cout, cin = w.shape[:2]
    wp = w.permute(2, 3, 0, 1).reshape(9 * cout, cin)
    wp = round_tf32(wp.float().contiguous()) if tf32 else wp.to(torch.bfloat16)
    pad = torch.zeros(npad - 9 * cout, cin, dtype=wp.dtype, device=wp.device)
    return torch.cat([wp, pad], 0).contiguous()


def pyramid_gather(part, bias4, lo, out):
    B, H, W, pc = part.shape
    _lib.check(_lib.lib().fd_pyramid_gather(_lib.ptr(part), pc, _lib.ptr(bias4), _lib.ptr(lo), _lib.ptr(out),
                                            B, H, W, _lib.stream_ptr()), "fd_pyramid_gather")
    return out


def conv_in(x4, w, b, out):
    B, H, W, _ = x4.shape
    if out.dtype == torch.float32:
        assert out.shape[3] == 64, "fp32 activations: the 4 -> 64 input conv of the FlowDec configuration"
        _lib.check(_lib.lib().fd_conv_in_f32(_lib.ptr(x4), _lib.ptr(w), _lib.ptr(b), _lib.ptr(out), B, H, W,
                                             _lib.stream_ptr()), "fd_conv_in_f32")
        return out
    if out.shape[3] != 64:
        _lib.check(_lib.lib().fd_conv_in_any(_lib.ptr(x4), _lib.ptr(w), _lib.ptr(b), _lib.ptr(out), B, H, W,
                                             out.shape[3], _lib.stream_ptr()), "fd_conv_in_any")
        return out
    _lib.check(_lib.lib().fd_conv_in(_lib.ptr(x4), _lib.ptr(w), _lib.ptr(b), _lib.ptr(out), B, H, W,
                                     _lib.stream_ptr()), "fd_conv_in")
    return out


def combine(pyr4, w, b, h, out):
    B, H, W, C = h.shape
    fn = _lib.lib().fd_combine_f32 if h.dtype == torch.float32 else _lib.lib().fd_combine
    _lib.check(fn(_lib.ptr(pyr4), _lib.ptr(w), _lib.ptr(b), _lib.ptr(h), _lib.ptr(out),
                  ctypes.c_size_t(B * H * W), C, _lib.stream_ptr()), "fd_combine")
    return out


def output_axpy(pyr4, w_out8, base1, c1, base2, c2, coef, out, v_out=None, base3=None, c3=0.0):
    """w_out8: host ctypes float[8] (2x4 output-layer weights)."""
    npix = pyr4.numel() // 4
    rc = _lib.lib().fd_output_axpy(_lib.ptr(pyr4), w_out8, _lib.ptr(base1), ctypes.c_float(c1),
                                   _lib.ptr(base2), ctypes.c_float(c2), _lib.ptr(base3), ctypes.c_float(c3),
                                   ctypes.c_float(coef), _lib.ptr(out), _lib.ptr(v_out),
                                   ctypes.c_size_t(npix), _lib.stream_ptr())
    _lib.check(rc, "fd_output_axpy")
    return out


def output_conv3_axpy(pyr4, w72, base1, c1, base2, c2, coef, out, v_out=None, base3=None, c3=0.0):
    """3x3 output layer (w72: device fp32 [2,4,3,3]) fused with the sampler stage"""
    B, H, W, _ = pyr4.shape
    rc = _lib.lib().fd_output_conv3_axpy(_lib.ptr(pyr4), _lib.ptr(w72), _lib.ptr(base1), ctypes.c_float(c1),
                                         _lib.ptr(base2), ctypes.c_float(c2), _lib.ptr(base3), ctypes.c_float(c3),
                                         ctypes.c_float(coef), _lib.ptr(out), _lib.ptr(v_out), B, H, W,
                                         _lib.stream_ptr())
    _lib.check(rc, "fd_output_conv3_axpy")
    return out


def x0_noise(Y, sigma_f64, eps, fac, out):
    """Y, eps, out: fp32 [B,F,T,2] (= complex64 [B,F,T]); sigma_f64: f64 [F]."""
    B, Fq, T = Y.shape[0], Y.shape[1], Y.shape[2]
    rc = _lib.lib().fd_x0(_lib.ptr(Y), _lib.ptr(sigma_f64), _lib.ptr(eps), ctypes.c_float(fac), _lib.ptr(out),
                          B, Fq, T, _lib.stream_ptr())
    _lib.check(rc, "fd_x0")
    return out


def fourier_embed(t, Wf, out):
    _lib.check(_lib.lib().fd_fourier_embed(ctypes.c_float(t), _lib.ptr(Wf), Wf.numel(), _lib.ptr(out),
                                           _lib.stream_ptr()), "fd_fourier_embed")
    return out


def matvec(inp, Wm, b, out, silu_in=False, add=None, out_scale=1.0):
    M, K = Wm.shape
    rc = _lib.lib().fd_matvec(_lib.ptr(inp), K, int(silu_in), _lib.ptr(Wm), _lib.ptr(b), _lib.ptr(add),
                              ctypes.c_float(out_scale), _lib.ptr(out), M, _lib.stream_ptr())
    _lib.check(rc, "fd_matvec")
    return out


# ---------------------------------------------------------------------------------------------
# STFT side (csrc/fd_stft.cu)
# ---------------------------------------------------------------------------------------------
def twiddles1534(device):
    tw = torch.empty(1534, 2, device=device, dtype=torch.float32)
    _lib.check(_lib.lib().fd_twiddles1534(_lib.ptr(tw), _lib.stream_ptr()), "fd_twiddles1534")
    return tw


def normfac(y2d, mode, out, lengths=None):
    """lengths (int32 [B], device) selects the ragged form: row b holds lengths[b] <= L valid samples"""
    B, L = y2d.shape
    if lengths is None:
        _lib.check(_lib.lib().fd_normfac(_lib.ptr(y2d), B, L, int(mode), _lib.ptr(out), _lib.stream_ptr()),
                   "fd_normfac")
    else:
        _lib.check(_lib.lib().fd_normfac_ragged(_lib.ptr(y2d), B, L, _lib.ptr(lengths), int(mode), _lib.ptr(out),
                                                _lib.stream_ptr()), "fd_normfac_ragged")
    return out


def stft_compress(y2d, nf, window, tw, alpha, beta, out, lengths=None):
    """y2d fp32 [B,L]; out float2 [B,768,Tp] (as fp32 [B,768,Tp,2])."""
    B, L = y2d.shape
    Tp = out.shape[2]
    if lengths is None:
        rc = _lib.lib().fd_stft1534_compress(_lib.ptr(y2d), B, L, _lib.ptr(nf), _lib.ptr(window), _lib.ptr(tw),
                                             ctypes.c_float(alpha), ctypes.c_float(beta), Tp, _lib.ptr(out),
                                             _lib.stream_ptr())
        _lib.check(rc, "fd_stft1534_compress")
    else:
        rc = _lib.lib().fd_stft1534_compress_ragged(_lib.ptr(y2d), B, L, _lib.ptr(lengths), _lib.ptr(nf),
                                                    _lib.ptr(window), _lib.ptr(tw), ctypes.c_float(alpha),
                                                    ctypes.c_float(beta), Tp, _lib.ptr(out), _lib.stream_ptr())
        _lib.check(rc, "fd_stft1534_compress_ragged")
    return out


STFT_PFA = True     # prime-factor FFT kernels (csrc/fd_stft_pfa.cu); False = the direct-DFT kernels (A/B, tests)


def stft_use_pfa(on):
    """selects the STFT / iSTFT algorithm in the library AND in these wrappers; returns the previous setting"""
    global STFT_PFA
    prev = STFT_PFA
    STFT_PFA = bool(on)
    _lib.lib().fd_stft_use_pfa(int(STFT_PFA))
    return prev


def istft_decompress(X, L, window, tw, nf, alpha, beta, out, lengths=None, ws=None):
    """ws: optional fp32 workspace [B, Tp, 1536] (static buffers of a captured graph pass theirs)"""
    B, Tp = X.shape[0], X.shape[2]
    if STFT_PFA and Tp % 4 == 0:
        if ws is None:
            ws = torch.empty(B, Tp, 1536, device=X.device, dtype=torch.float32)
        assert ws.is_contiguous() and ws.numel() >= B * Tp * 1536
        rc = _lib.lib().fd_istft1534_decompress_pfa(_lib.ptr(X), B, Tp, L, _lib.ptr(lengths), _lib.ptr(window),
                                                    _lib.ptr(tw), _lib.ptr(nf), ctypes.c_float(alpha),
                                                    ctypes.c_float(beta), _lib.ptr(ws), _lib.ptr(out), _lib.stream_ptr())
        _lib.check(rc, "fd_istft1534_decompress_pfa")
        return out
    if lengths is None:
        rc = _lib.lib().fd_istft1534_decompress(_lib.ptr(X), B, Tp, L, _lib.ptr(window), _lib.ptr(tw), _lib.ptr(nf),
                                                ctypes.c_float(alpha), ctypes.c_float(beta), _lib.ptr(out),
                                                _lib.stream_ptr())
        _lib.check(rc, "fd_istft1534_decompress")
    else:
        rc = _lib.lib().fd_istft1534_decompress_ragged(_lib.ptr(X), B, Tp, L, _lib.ptr(lengths), _lib.ptr(window),
                                                       _lib.ptr(tw), _lib.ptr(nf), ctypes.c_float(alpha),
                                                       ctypes.c_float(beta), _lib.ptr(out), _lib.stream_ptr())
        _lib.check(rc, "fd_istft1534_decompress_ragged")
    return out
