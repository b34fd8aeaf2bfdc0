"""GPU parity of the side kernels (fd_elementwise.cu, fd_stft.cu) against the CPU oracle
(oracle/flowdec_oracle.py) on seeded inputs.  fp32-exact ops: rel-L2 <= 1e-5; bf16-output ops:
max error <= 2^-8 relative to the tensor scale (one bf16 rounding)."""

import pytest
import torch
import torch.nn.functional as F

from flowdec_b200 import ops
from flowdec_b200.data.feature_extractors import AmplitudeCompressedComplexSTFT
from flowdec_b200.util.synth import synth_waveforms
from oracle import flowdec_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def nhwc_bf16(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw_f32(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("C1,C2,mode", [(64, 0, 0), (256, 0, 1), (128, 0, 2), (256, 64, 0), (128, 256, 0), (256, 256, 0),
                                        (64, 0, 1), (256, 0, 2)])
def test_groupnorm_silu_resample(C1, C2, mode):
    torch.manual_seed(0)
    B, H, W = 2, 16, 24
    x1 = (torch.randn(B, C1, H, W) * 1.5 + 0.3)
    x2 = torch.randn(B, C2, H, W) * 0.7 if C2 else None
    C = C1 + C2
    gamma, beta = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
    s1 = nhwc_bf16(x1).to(DEV)
    s2 = nhwc_bf16(x2).to(DEV) if C2 else None
    srcs = [s1] + ([s2] if C2 else [])
    parts = [ops.chan_stats(s, 8) for s in srcs]
    ss = torch.empty(B, C, 2, device=DEV)
    ops.gn_finalize(parts, [s.shape[3] for s in srcs], H * W, gamma.to(DEV), beta.to(DEV), min(C // 4, 32), 1e-6, ss)
    Ho, Wo = (H // 2, W // 2) if mode == 1 else ((2 * H, 2 * W) if mode == 2 else (H, W))
    out = torch.empty(B, Ho, Wo, C, device=DEV, dtype=torch.bfloat16)
    raw = torch.empty_like(out)
    if mode == 0:
        ops.gn_act_resample(srcs, ss, out, mode)
        raw = None
    else:
        ops.gn_act_resample(srcs, ss, out, mode, out_raw=raw)       # fused: one read, two outputs
        raw2 = torch.empty_like(out)
        ops.gn_act_resample(srcs, None, None, mode, out_raw=raw2)   # raw only
        out2 = torch.empty_like(out)
        ops.gn_act_resample(srcs, ss, out2, mode)                   # activated only
        torch.cuda.synchronize()
        assert torch.equal(raw2, raw) and torch.equal(out2, out)
    torch.cuda.synchronize()
    # oracle on the same bf16-rounded inputs
    xc = torch.cat([nchw_f32(s.cpu()) for s in srcs], 1)
    h = F.silu(F.group_norm(xc, min(C // 4, 32), gamma, beta, eps=1e-6))
    r = xc
    if mode == 1:
        h, r = O.fir_down2(h), O.fir_down2(r)
    elif mode == 2:
        h, r = O.fir_up2(h), O.fir_up2(r)
    e1 = (nchw_f32(out.cpu()) - h).abs().max().item()
    assert e1 <= 2 ** -8 * h.abs().max().item() + 1e-3, e1
    if raw is not None:
        e2 = (nchw_f32(raw.cpu()) - r).abs().max().item()
        assert e2 <= 2 ** -8 * r.abs().max().item() + 1e-3, e2


@pytest.mark.parametrize("B,H,W,C1,C2,mode", [(2, 16, 32, 64, 0, 2), (2, 16, 32, 64, 0, 1), (1, 24, 48, 128, 64, 2),
                                              (1, 24, 48, 128, 64, 1), (4, 128, 64, 256, 0, 1), (2, 64, 64, 256, 0, 2),
                                              (1, 8, 16, 64, 64, 2), (1, 8, 16, 64, 64, 1)])
def test_fir_tile_kernels(B, H, W, C1, C2, mode):
    """TMA-tiled GroupNorm+SiLU+FIR kernels (fd_fir_tiles.cu) vs the oracle and vs the register kernels;
    the largest cases give every persistent block several tiles (both ring stages wrap)."""
    torch.manual_seed(2)
    C = C1 + C2
    x1 = torch.randn(B, C1, H, W) * 1.5 + 0.3
    x2 = torch.randn(B, C2, H, W) * 0.7 if C2 else None
    gamma, beta = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
    srcs = [nhwc_bf16(x1).to(DEV)] + ([nhwc_bf16(x2).to(DEV)] if C2 else [])
    parts = [ops.chan_stats(s, 8) for s in srcs]
    ss = torch.empty(B, C, 2, device=DEV)
    ops.gn_finalize(parts, [s.shape[3] for s in srcs], H * W, gamma.to(DEV), beta.to(DEV), min(C // 4, 32), 1e-6, ss)
    Ho, Wo = (H // 2, W // 2) if mode == 1 else (2 * H, 2 * W)
    out = torch.full((B, Ho, Wo, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    raw = torch.full_like(out, float("nan"))
    assert ops.fir_tiles_enable(True) in (True, False)
    ops.gn_act_resample(srcs, ss, out, mode, out_raw=raw)
    prev = ops.fir_tiles_enable(False)
    assert prev is True
    try:
        out_r, raw_r = torch.empty_like(out), torch.empty_like(raw)
        ops.gn_act_resample(srcs, ss, out_r, mode, out_raw=raw_r)
    finally:
        ops.fir_tiles_enable(True)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all() and torch.isfinite(raw.float()).all()
    xc = torch.cat([nchw_f32(s.cpu()) for s in srcs], 1)
    h = F.silu(F.group_norm(xc, min(C // 4, 32), gamma, beta, eps=1e-6))
    h, r = (O.fir_down2(h), O.fir_down2(xc)) if mode == 1 else (O.fir_up2(h), O.fir_up2(xc))
    e1 = (nchw_f32(out.cpu()) - h).abs().max().item()
    e2 = (nchw_f32(raw.cpu()) - r).abs().max().item()
    assert e1 <= 2 ** -8 * h.abs().max().item() + 1e-3, e1
    assert e2 <= 2 ** -8 * r.abs().max().item() + 1e-3, e2
    # same fp32 arithmetic up to summation order: at most one bf16 ulp apart from the register kernels
    assert (out.float() - out_r.float()).abs().max().item() <= 2 ** -7 * h.abs().max().item()
    assert (raw.float() - raw_r.float()).abs().max().item() <= 2 ** -7 * r.abs().max().item()
    assert (out != out_r).float().mean().item() < 0.05


def test_conv_in_combine_pyramid_output():
    torch.manual_seed(1)
    B, H, W = 2, 32, 16
    x = torch.randn(B, 1, H, W, dtype=torch.complex64)
    y = torch.randn(B, 1, H, W, dtype=torch.complex64)
    xr = torch.view_as_real(x).squeeze(1).contiguous().to(DEV)
    yr = torch.view_as_real(y).squeeze(1).contiguous().to(DEV)
    p4 = torch.empty(B, H, W, 4, device=DEV)
    ops.pack4(xr, yr, p4)
    h0 = torch.cat([x.real, x.imag, y.real, y.imag], 1)
    assert torch.equal(p4.cpu().permute(0, 3, 1, 2), h0)
    # conv_in
    w, b = torch.randn(64, 4, 3, 3) * 0.2, torch.randn(64) * 0.1
    out = torch.empty(B, H, W, 64, device=DEV, dtype=torch.bfloat16)
    ops.conv_in(p4, w.to(DEV), b.to(DEV), out)
    ref = F.conv2d(h0, w, b, padding=1)
    assert (nchw_f32(out.cpu()) - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item() + 1e-3
    # fir_down4
    d4 = torch.empty(B, H // 2, W // 2, 4, device=DEV)
    ops.fir_down4(p4, d4)
    assert rel_l2(d4.cpu().permute(0, 3, 1, 2), O.fir_down2(h0)) < 1e-6
    # combine
    C = 256
    hh = torch.randn(B, C, H // 2, W // 2)
    hb = nhwc_bf16(hh).to(DEV)
    wc, bc = torch.randn(C, 4) * 0.3, torch.randn(C) * 0.1
    oc = torch.empty_like(hb)
    ops.combine(d4, wc.to(DEV), bc.to(DEV), hb, oc)
    refc = F.conv2d(O.fir_down2(h0), wc.reshape(C, 4, 1, 1), bc) + nchw_f32(hb.cpu())
    assert (nchw_f32(oc.cpu()) - refc).abs().max().item() <= 2 ** -8 * refc.abs().max().item() + 1e-3
    # pyramid up + add
    lo = torch.randn(B, 4, H // 2, W // 2)
    add = torch.randn(B, 4, H, W)
    lo_d = lo.permute(0, 2, 3, 1).contiguous().to(DEV)
    add_d = add.permute(0, 2, 3, 1).contiguous().to(DEV)
    ops.pyramid_up_add(lo_d, add_d, add_d)
    assert rel_l2(add_d.cpu().permute(0, 3, 1, 2), O.fir_up2(lo) + add) < 1e-6
    # output layer + ODE stage
    import ctypes
    wo = torch.randn(2, 4)
    w8 = (ctypes.c_float * 8)(*wo.flatten().tolist())
    base = torch.randn(B, H, W, 2)
    base2 = torch.randn(B, H, W, 2)
    od = torch.empty(B, H, W, 2, device=DEV)
    vd = torch.empty(B, H, W, 2, device=DEV)
    base3 = torch.randn(B, H, W, 2)
    ops.output_axpy(add_d, w8, base.to(DEV), 0.5, base2.to(DEV), 0.25, 0.125, od, vd, base3=base3.to(DEV), c3=-2.0)
    v = torch.einsum("oc,bhwc->bhwo", wo, add_d.cpu())
    assert rel_l2(vd.cpu(), v) < 1e-6
    assert rel_l2(od.cpu(), 0.5 * base + 0.25 * base2 - 2.0 * base3 + 0.125 * v) < 1e-6


def test_time_embedding_matvec():
    torch.manual_seed(2)
    nf = 64
    sd = {"all_modules.0.W": torch.randn(nf) * 16,
          "all_modules.1.weight": torch.randn(256, 128) * 0.1, "all_modules.1.bias": torch.randn(256) * 0.1,
          "all_modules.2.weight": torch.randn(256, 256) * 0.1, "all_modules.2.bias": torch.randn(256) * 0.1}
    for t in (0.0, 0.3, 5.0 / 6.0):
        ref = O.time_embedding(sd, torch.tensor([t], dtype=torch.float32))[0]
        four = torch.empty(128, device=DEV)
        h1 = torch.empty(256, device=DEV)
        te = torch.empty(256, device=DEV)
        ops.fourier_embed(t, sd["all_modules.0.W"].to(DEV), four)
        ops.matvec(four, sd["all_modules.1.weight"].to(DEV), sd["all_modules.1.bias"].to(DEV), h1)
        ops.matvec(h1, sd["all_modules.2.weight"].to(DEV), sd["all_modules.2.bias"].to(DEV), te, silu_in=True)
        assert rel_l2(te.cpu(), ref) < 2e-5, (t, rel_l2(te.cpu(), ref))


@pytest.fixture(params=[True, False], ids=["pfa_fft", "direct_dft"])
def stft_algo(request):
    """both STFT / iSTFT implementations: the prime-factor FFT (default) and the direct DFT"""
    old = ops.stft_use_pfa(request.param)
    yield request.param
    ops.stft_use_pfa(old)


@pytest.mark.parametrize("L,kind", [(24000, "tones"), (48000, "gauss"), (30011, "tones"), (768, "gauss"), (24000, "zeros"),
                                    (96000, "tones"), (61001, "gauss")])
def test_stft_istft_vs_oracle(L, kind, stft_algo):
    B = 2
    y = synth_waveforms(B, L, seed=99, kind=kind)
    fe = AmplitudeCompressedComplexSTFT("hann", 1534, 48000, alpha=0.3, beta=0.33, n_hops=4).to(DEV)
    window = O.hann_window()
    Yref, info = O.preprocess(y, window)                 # normalise + stft + compress + pad
    Tp = Yref.shape[-1]
    y2 = y.reshape(B, L).to(DEV)
    nf = torch.empty(B, device=DEV)
    ops.normfac(y2, 1, nf)
    assert torch.allclose(nf.cpu(), info["normfac"].reshape(B), rtol=0, atol=0)
    Yd = torch.empty(B, 768, Tp, 2, device=DEV)
    fe.stft_compress(y2, nf, Yd)
    Yg = torch.view_as_complex(Yd.cpu()).unsqueeze(1)
    if kind != "zeros":
        assert rel_l2(torch.view_as_real(Yg), torch.view_as_real(Yref)) < 1e-5
    else:
        assert Yg.abs().max().item() == 0.0
    # inverse of the oracle's spectrogram
    out = torch.empty(B, L, device=DEV)
    fe.istft_decompress(torch.view_as_real(Yref.squeeze(1).contiguous()).to(DEV), L, nf, out)
    xref = O.postprocess(Yref, info, window).reshape(B, L)
    if kind != "zeros":
        assert rel_l2(out.cpu(), xref) < 2e-5
        # invertibility contract of the reference feature extractor (feature_extractors.py:20-23)
        assert (out.cpu() - y.reshape(B, L)).abs().max().item() < 2e-5
    else:
        assert out.abs().max().item() == 0.0


def test_x0_noise():
    torch.manual_seed(3)
    B, Fq, T = 2, 768, 64
    Y = torch.randn(B, 1, Fq, T, dtype=torch.complex64)
    eps = torch.randn(B, 1, Fq, T, dtype=torch.complex64)
    sigma = torch.rand(Fq, 1, dtype=torch.float64) * 0.5 + 0.1
    ref = O.initial_noise(Y, sigma, eps)
    out = torch.empty(B, Fq, T, 2, device=DEV)
    ops.x0_noise(torch.view_as_real(Y.squeeze(1).contiguous()).to(DEV), sigma.reshape(-1).to(DEV),
                 torch.view_as_real(eps.squeeze(1).contiguous()).to(DEV), 1.0, out)
    assert torch.equal(torch.view_as_complex(out.cpu()).unsqueeze(1), ref)
