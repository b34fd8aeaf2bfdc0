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
