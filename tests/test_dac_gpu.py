"""GPU parity of the NDAC decode path (fd_dac.cu).

Per-op: each kernel against torch functional fp32 on CPU (rel-L2 <= 1e-5).
End to end: against oracle/dac_oracle.py (pinned to the transformers port of
descript-audio-codec) evaluated in fp64; the synthetic decoder is chaotic enough that the
oracle's own fp32 result sits ~1e-4 from fp64, so the gate is "as accurate as the fp32 CPU
reference": rel(gpu, fp64) <= 3 * rel(cpu_fp32, fp64) + 1e-5."""
import math

import pytest
import torch
import torch.nn.functional as F

from flowdec_b200 import _lib
from flowdec_b200.ndac import DAC
from oracle import dac_oracle as D

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


@pytest.mark.parametrize("Cin,Cout,K,dil,T", [(96, 96, 7, 1, 300), (192, 192, 7, 9, 257), (64, 130, 7, 3, 77),
                                               (48, 48, 1, 1, 500), (96, 1, 7, 1, 1000)])
def test_conv1d_snake_residual(Cin, Cout, K, dil, T):
    torch.manual_seed(0)
    B = 2
    x = torch.randn(B, Cin, T)
    w = torch.randn(Cout, Cin, K) / math.sqrt(Cin * K)
    b = torch.randn(Cout) * 0.1
    alpha = torch.rand(Cin) + 0.5
    pad = (K - 1) * dil // 2
    res = torch.randn(B, Cout, T)
    ref = torch.tanh(F.conv1d(D.snake(x, alpha.reshape(1, -1, 1)), w, b, dilation=dil, padding=pad) + res)
    out = torch.empty(B, Cout, T, device="cuda")
    c = lambda t: t.cuda().contiguous()
    xd, wd, bd, ad, rd = c(x), c(w), c(b), c(alpha), c(res)
    rc = _lib.lib().fd_dac_conv1d(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(ad), _lib.ptr(rd),
                                  _lib.ptr(out), B, Cin, Cout, T, K, dil, pad, 1, _lib.stream_ptr())
    _lib.check(rc, "fd_dac_conv1d")
    assert rel(out.cpu(), ref) < 1e-5
    # plain conv (no snake / residual / tanh)
    ref2 = F.conv1d(x, w, b, dilation=dil, padding=pad)
    rc = _lib.lib().fd_dac_conv1d(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), None, None, _lib.ptr(out), B, Cin,
                                  Cout, T, K, dil, pad, 0, _lib.stream_ptr())
    _lib.check(rc, "fd_dac_conv1d")
    assert rel(out.cpu(), ref2) < 1e-5


@pytest.mark.parametrize("Cin,Cout,s,T", [(128, 64, 2, 100), (96, 48, 3, 61), (256, 128, 4, 150), (192, 96, 5, 40),
                                          (256, 128, 8, 33)])
def test_conv_transpose1d(Cin, Cout, s, T):
    torch.manual_seed(1)
    B = 2
    x = torch.randn(B, Cin, T)
    w = torch.randn(Cin, Cout, 2 * s) / math.sqrt(Cin * 2)
    b = torch.randn(Cout) * 0.1
    alpha = torch.rand(Cin) + 0.5
    pad = math.ceil(s / 2)
    ref = F.conv_transpose1d(D.snake(x, alpha.reshape(1, -1, 1)), w, b, stride=s, padding=pad)
    out = torch.empty(B, Cout, ref.shape[-1], device="cuda")
    c = lambda t: t.cuda().contiguous()
    xd, wd, bd, ad = c(x), c(w), c(b), c(alpha)
    rc = _lib.lib().fd_dac_conv_transpose1d(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(ad), _lib.ptr(out),
                                            B, Cin, Cout, T, s, pad, _lib.stream_ptr())
    _lib.check(rc, "fd_dac_conv_transpose1d")
    assert rel(out.cpu(), ref) < 1e-5


@pytest.mark.parametrize("latent,dim,rates,nq,T", [(64, 96, (4, 3, 2), 5, 37), (128, 256, (8, 5, 4, 4), 10, 19)])
def test_from_codes_and_decode(latent, dim, rates, nq, T):
    sd = D.synth_dac_state_dict(latent, dim, rates, nq, seed=1)
    model = DAC(sd, decoder_dim=dim, decoder_rates=rates, n_codebooks=nq, latent_dim=latent,
                sample_rate=48000).to("cuda").eval()
    # built on the CPU, moved with .to("cuda"): every tensor the op lists resolve to must have followed (round-1 bug)
    assert all(t.is_cuda for t in model.op_tensors().values())
    g = torch.Generator().manual_seed(2)
    codes = torch.randint(0, 1024, (2, nq, T), generator=g)
    sd64 = {k: v.double() for k, v in sd.items()}
    with torch.no_grad():
        z_ref = D.from_codes(sd, codes)
        x32 = D.decode(sd, z_ref, rates)
        x64 = D.decode(sd64, D.from_codes(sd64, codes), rates)
    zq, _, c = model.quantizer.from_codes(codes)
    assert rel(zq.cpu(), z_ref) < 1e-5
    x = model.decode(zq)
    assert x.shape == x64.shape
    gate = 3 * rel(x32, x64) + 1e-5
    r = rel(x.cpu(), x64)
    print(f"\nNDAC decode rel-L2 vs fp64 oracle: gpu {r:.3e}, cpu fp32 {rel(x32, x64):.3e}")
    assert r <= gate, (r, gate)
    # fewer codebooks than the model has (bitrate scalability, demo.ipynb:85-88)
    zq2, _, _ = model.quantizer.from_codes(codes[:, :3])
    assert rel(zq2.cpu(), D.from_codes(sd, codes[:, :3])) < 1e-5


# ------------------------------------------------------------------------------------------------
# encode half: strided conv, RVQ encode, DAC.preprocess / encode / forward
@pytest.mark.parametrize("Cin,Cout,s,T", [(8, 16, 2, 1000), (64, 128, 4, 517), (96, 192, 8, 2048), (32, 64, 3, 301),
                                          (128, 256, 8, 130)])
def test_conv1d_strided_snake(Cin, Cout, s, T):
    torch.manual_seed(3)
    B, K, pad = 2, 2 * s, math.ceil(s / 2)
    x = torch.randn(B, Cin, T)
    w = torch.randn(Cout, Cin, K) / math.sqrt(Cin * K)
    b = torch.randn(Cout) * 0.1
    alpha = torch.rand(Cin) + 0.5
    ref = F.conv1d(D.snake(x, alpha.reshape(1, -1, 1)), w, b, stride=s, padding=pad)
    out = torch.empty(B, Cout, ref.shape[-1], device="cuda")
    c = lambda t: t.cuda().contiguous()
    xd, wd, bd, ad = c(x), c(w), c(b), c(alpha)
    rc = _lib.lib().fd_dac_conv1d_strided(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(ad), _lib.ptr(out),
                                          B, Cin, Cout, T, K, s, pad, _lib.stream_ptr())
    _lib.check(rc, "fd_dac_conv1d_strided")
    assert rel(out.cpu(), ref) < 1e-5


def _check_codes(codes_gpu, z_in, sd, n_q):
    """index parity with the oracle on the SAME latent; a differing index is only accepted where the oracle's
    own best/second-best margin is below fp32 resolution of the distance (and then later stages of that
    column legitimately diverge, so the comparison stops at the first such quantizer)."""
    zq_o, codes_o, lat_o, commit_o, _, margin = D.rvq_encode(sd, z_in, n_q)
    cg = codes_gpu.cpu()
    assert cg.shape == codes_o.shape and cg.dtype == torch.int64
    B, nq, T = cg.shape
    diverged = torch.zeros(B, T, dtype=torch.bool)
    n_bad = 0
    for i in range(nq):
        diff = (cg[:, i] != codes_o[:, i]) & ~diverged
        assert (margin[:, i][diff] < 2e-6).all(), f"quantizer {i}: index differs with margin {margin[:, i][diff].max()}"
        n_bad += int(diff.sum())
        diverged |= diff
    return zq_o, codes_o, lat_o, commit_o, diverged, n_bad


@pytest.mark.parametrize("B,D_,T,nq_total,n_q", [(2, 64, 42, 5, None), (3, 1024, 173, 9, 9), (1, 128, 1, 4, 2),
                                                  (2, 256, 90, 16, 16)])
def test_rvq_encode(B, D_, T, nq_total, n_q):
    sd = D.synth_dac_state_dict(D_, 32, (2,), nq_total, seed=4)
    model = DAC(sd, decoder_dim=32, decoder_rates=(2,), n_codebooks=nq_total, latent_dim=D_).to("cuda").eval()
    g = torch.Generator().manual_seed(5)
    z = torch.randn(B, D_, T, generator=g)
    zq, codes, latents, commit, cbl = model.quantizer(z.cuda(), n_q)
    zq_o, codes_o, lat_o, commit_o, diverged, n_bad = _check_codes(codes, z, sd, n_q)
    ok = ~diverged
    print(f"\nRVQ encode: {int(ok.sum())}/{ok.numel()} columns index-identical through all quantizers "
          f"({n_bad} near-tie flips)")
    assert ok.float().mean() > 0.98
    assert rel(zq.cpu().transpose(1, 2)[ok], zq_o.transpose(1, 2)[ok]) < 1e-5
    assert rel(latents.cpu().transpose(1, 2)[ok], lat_o.transpose(1, 2)[ok]) < 1e-4
    if ok.all():
        assert abs(commit.item() - commit_o.item()) <= 1e-4 * abs(commit_o.item())
        assert cbl.item() == commit.item()
        # from_codes(encode(z).codes) reproduces z_q: the decode-side entry point and the encoder agree
        assert rel(model.quantizer.from_codes(codes)[0], zq) < 1e-6


def test_dac_preprocess_encode_decode():
    latent, dim, rates, nq, edim, erates = 64, 96, (4, 3, 2), 6, 8, (2, 3, 4)
    sd = D.synth_dac_state_dict(latent, dim, rates, nq, seed=1, encoder_dim=edim, encoder_rates=erates)
    model = DAC(sd, decoder_dim=dim, decoder_rates=rates, n_codebooks=nq, latent_dim=latent, encoder_dim=edim,
                encoder_rates=erates, sample_rate=48000).to("cuda").eval()
    assert model.hop_length == 24
    g = torch.Generator().manual_seed(6)
    audio = 0.3 * torch.randn(3, 1, 4001, generator=g)
    x = model.preprocess(audio.cuda(), 48000)
    assert x.shape[-1] == 4008 and torch.equal(x[..., :4001].cpu(), audio) and (x[..., 4001:] == 0).all()
    with pytest.raises(AssertionError):
        model.preprocess(audio, 44100)
    sd64 = {k: v.double() for k, v in sd.items()}
    with torch.no_grad():
        z32 = D.encode(sd, x.cpu(), erates)
        z64 = D.encode(sd64, x.cpu().double(), erates)
    z_gpu = model._run(model._enc_ops, x)
    r, r32 = rel(z_gpu.cpu(), z64), rel(z32, z64)
    print(f"\nNDAC encoder rel-L2 vs fp64 oracle: gpu {r:.3e}, cpu fp32 {r32:.3e}")
    assert z_gpu.shape == z64.shape and r <= 3 * r32 + 1e-5
    # encode(): codes are checked against the oracle quantizer run on the GPU's own encoder output
    zq, codes, latents, commit, _ = model.encode(x, n_quantizers=4)
    assert codes.shape == (3, 4, 167) and latents.shape == (3, 32, 167)
    zq_o, codes_o, _, _, diverged, _ = _check_codes(codes, z_gpu.cpu(), sd, 4)
    assert (~diverged).float().mean() > 0.98
    ok = ~diverged
    assert rel(zq.cpu().transpose(1, 2)[ok], zq_o.transpose(1, 2)[ok]) < 1e-5
    out = model(audio.cuda(), 48000, n_quantizers=4)
    assert out["audio"].shape == audio.shape and torch.equal(out["codes"], codes)
    assert rel(out["audio"], model.decode(zq)[..., :4001]) == 0.0


# ------------------------------------------------------------------------------------------------
# tensor-core decoder (fd_dac_tc.cu): tf32 operands, fp32 accumulation, time-major activations
def _rt(x):
    """fp64 tensor with values rounded to tf32 (10-bit mantissa)"""
    i = x.float().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32).double()


def _tc_layer(x_ntc, w_packed, offs, bias, residual=None, alpha=None, alpha_mod=0, want_raw=True, Tout=None):
    B, T, C = x_ntc.shape
    raw, act = DAC._tc_conv((x_ntc, 0), T, T * C, C, w_packed, offs, bias, (residual, 0) if residual is not None else None,
                            (residual.shape[1] * residual.shape[2]) if residual is not None else 0, alpha, alpha_mod,
                            want_raw, alpha is not None, Tout or T)
    return raw, act


@pytest.mark.parametrize("C,Cout,dil,T", [(96, 96, 1, 300), (192, 192, 9, 257), (64, 384, 3, 130), (1024, 1536, 1, 75)])
def test_tc_conv1d_layer(C, Cout, dil, T):
    from flowdec_b200.ops import round_tf32
    torch.manual_seed(0)
    B = 2
    x = torch.randn(B, C, T)
    w = torch.randn(Cout, C, 7) / math.sqrt(7 * C)
    b = torch.randn(Cout) * 0.1
    alpha = torch.rand(Cout) + 0.5
    res = torch.randn(B, Cout, T)
    ref_raw = F.conv1d(x.double(), w.double(), b.double(), dilation=dil, padding=3 * dil) + res.double()
    ref_act = D.snake(ref_raw, alpha.double().reshape(1, -1, 1))
    wp = round_tf32(w.permute(0, 2, 1).reshape(Cout, -1).contiguous()).cuda()
    raw, act = _tc_layer(x.permute(0, 2, 1).contiguous().cuda(), wp, [(j - 3) * dil for j in range(7)], b.cuda(),
                         residual=res.permute(0, 2, 1).contiguous().cuda(), alpha=alpha.cuda(), alpha_mod=Cout)
    r1, r2 = rel(raw.cpu().permute(0, 2, 1), ref_raw), rel(act.cpu().permute(0, 2, 1), ref_act)
    print(f"\ntc conv1d C={C}->{Cout} dil={dil}: rel-L2 raw {r1:.3e} act {r2:.3e}")
    assert r1 < 1.5e-3 and r2 < 1.5e-3


@pytest.mark.parametrize("Cin,Cout,s,T", [(128, 64, 2, 100), (256, 128, 4, 150), (192, 96, 5, 40), (1536, 768, 8, 33)])
def test_tc_conv_transpose_layer(Cin, Cout, s, T):
    from flowdec_b200.ops import round_tf32
    torch.manual_seed(1)
    B = 2
    x = torch.randn(B, Cin, T)
    w = torch.randn(Cin, Cout, 2 * s) / math.sqrt(Cin * 2)
    b = torch.randn(Cout) * 0.1
    pad = math.ceil(s / 2)
    ref = F.conv_transpose1d(x.double(), w.double(), b.double(), stride=s, padding=pad)
    wp = round_tf32(w.reshape(Cin, Cout, 2, s).permute(3, 1, 2, 0).reshape(s * Cout, 2 * Cin).contiguous()).cuda()
    raw, _ = _tc_layer(x.permute(0, 2, 1).contiguous().cuda(), wp, [0, -1], b.repeat(s).cuda(), Tout=T + 1)
    Tout = ref.shape[-1]
    got = raw.cpu().reshape(B, (T + 1) * s, Cout)[:, pad:pad + Tout].permute(0, 2, 1)
    r = rel(got, ref)
    print(f"\ntc conv_transpose1d {Cin}->{Cout} s={s}: rel-L2 {r:.3e}")
    assert r < 1.5e-3


@pytest.mark.parametrize("latent,dim,rates,nq,T", [(64, 256, (4, 2), 3, 37), (128, 512, (8, 5, 2), 6, 150)])
def test_tc_decode_end_to_end(latent, dim, rates, nq, T):
    """whole decoder on tensor cores vs the fp64 oracle; the fp32 CUDA-core decoder on the same model is the A/B"""
    sd = D.synth_dac_state_dict(latent, dim, rates, nq, seed=2)
    model = DAC(sd, decoder_dim=dim, decoder_rates=rates, n_codebooks=nq, latent_dim=latent, sample_rate=48000)
    assert model.tc_eligible and model.precision == "tf32"
    model = model.to("cuda").eval()
    assert all(t.is_cuda for t in model.op_tensors().values())
    codes = torch.randint(0, 1024, (3, nq, T), generator=torch.Generator().manual_seed(3))
    sd64 = {k: v.double() for k, v in sd.items()}
    with torch.no_grad():
        x64 = D.decode(sd64, D.from_codes(sd64, codes), rates)
        # the tf32 floor of THIS (synthetic, error-amplifying) decoder: the fp64 oracle with only the conv operands
        # rounded to tf32 — what any tf32 execution (incl. the reference's cuDNN default) would show
        c1, ct = F.conv1d, F.conv_transpose1d
        try:
            F.conv1d = lambda x, w, b=None, **k: c1(_rt(x), _rt(w), b, **k)
            F.conv_transpose1d = lambda x, w, b=None, **k: ct(_rt(x), _rt(w), b, **k)
            x_em = D.decode(sd64, D.from_codes(sd64, codes), rates)
        finally:
            F.conv1d, F.conv_transpose1d = c1, ct
    zq, _, _ = model.quantizer.from_codes(codes)
    x_tc = model.decode(zq)
    model.precision = "fp32"
    x_32 = model.decode(zq)
    assert x_tc.shape == x64.shape == x_32.shape
    r_tc, r_32, r_em = rel(x_tc.cpu(), x64), rel(x_32.cpu(), x64), rel(x_em, x64)
    print(f"\nNDAC decode rel-L2 vs fp64 oracle: tensor-core tf32 {r_tc:.3e} (tf32-operand emulation of the oracle "
          f"{r_em:.3e}), CUDA-core fp32 {r_32:.3e}")
    assert r_tc <= 2.0 * r_em + 1e-4 and r_32 <= 1e-3


@pytest.mark.parametrize("latent,edim,erates,T", [(64, 32, (2, 4), 4000), (128, 32, (2, 4, 5), 12000)])
def test_tc_encode_end_to_end(latent, edim, erates, T):
    """encoder on tensor cores (strided convs as 3-tap GEMMs over the [T/s, s*C] view) vs the fp64 oracle, gated on the
    tf32-operand emulation of the oracle; the fp32 CUDA-core encoder on the same model is the A/B"""
    drates = tuple(reversed(erates))
    sd = D.synth_dac_state_dict(latent, 256, drates[:2], 4, seed=5, encoder_dim=edim, encoder_rates=erates)
    model = DAC(sd, decoder_dim=256, decoder_rates=drates[:2], n_codebooks=4, latent_dim=latent, encoder_dim=edim,
                encoder_rates=erates, sample_rate=48000).to("cuda").eval()
    assert model.enc_tc_eligible and model.precision == "tf32" and model.encoder_precision == "fp32"
    g = torch.Generator().manual_seed(7)
    audio = 0.3 * torch.randn(2, 1, T, generator=g)
    x = model.preprocess(audio, 48000)
    sd64 = {k: v.double() for k, v in sd.items()}
    with torch.no_grad():
        z64 = D.encode(sd64, x.double(), erates)
        c1 = F.conv1d
        try:
            F.conv1d = lambda a, w, b=None, **k: c1(_rt(a), _rt(w), b, **k)
            z_em = D.encode(sd64, x.double(), erates)
        finally:
            F.conv1d = c1
    z_tc = model._encode_tc(x.cuda().float().contiguous())
    z_32 = model._run(model._enc_ops, x.cuda().float().contiguous())
    assert z_tc.shape == z64.shape == z_32.shape
    r_tc, r_32, r_em = rel(z_tc.cpu(), z64), rel(z_32.cpu(), z64), rel(z_em, z64)
    print(f"\nNDAC encoder latent rel-L2 vs fp64 oracle: tensor-core tf32 {r_tc:.3e} (tf32-operand emulation {r_em:.3e}), "
          f"CUDA-core fp32 {r_32:.3e}")
    assert r_tc <= 2.0 * r_em + 1e-4 and r_32 <= 1e-3
    # opted in, the public entry point runs the tensor-core encoder and the RVQ on its latent
    model.encoder_precision = "tf32"
    zq, codes, latents, _, _ = model.encode(x.cuda(), n_quantizers=4)
    assert codes.shape == (2, 4, z64.shape[-1]) and torch.isfinite(zq).all()
    assert rel(model.quantizer.from_codes(codes)[0], zq) < 1e-6
