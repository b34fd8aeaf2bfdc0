"""GPU parity: tcgen05 implicit-GEMM conv (fd_conv2d_igemm) vs torch F.conv2d in fp32 on
the same bf16-rounded operands.  Tolerance: fp32-accumulation-order noise + one bf16
output rounding (rel 2^-8), i.e. |err| <= 1e-2 * max|ref| elementwise is generous while any
descriptor/layout bug produces O(1) errors."""

import pytest
import torch
import torch.nn.functional as F

from flowdec_b200.ops import conv_igemm, pack_conv_weight

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[(True, True, True), (True, True, False), (True, False, False), (False, False, False)],
                ids=["halo_cluster4", "halo_pair", "cta_pair", "single_cta"])
def _mma_variant(request):
    """every conv test runs with 16x8 halo tiles in 4-CTA clusters (two MMA pairs sharing a multicast weight stream),
    halo tiles in CTA pairs, plain CTA pairs and single-CTA MMAs"""
    from flowdec_b200 import ops
    old = (ops.CTA_PAIRS, ops.HALO_TILES)
    ops.CTA_PAIRS, ops.HALO_TILES, cl4 = request.param
    old_cl4 = ops.conv_cluster4(cl4)
    yield
    ops.CTA_PAIRS, ops.HALO_TILES = old
    ops.conv_cluster4(old_cl4)


def _ref_conv(x_nhwc_bf16, w_oihw_bf16, bias):
    x = x_nhwc_bf16.float().permute(0, 3, 1, 2)
    y = F.conv2d(x, w_oihw_bf16.float(), bias, padding=w_oihw_bf16.shape[-1] // 2)
    return y.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [
    (1, 16, 16, 64, 256, 1),
    (1, 16, 16, 64, 256, 3),
    (2, 96, 8, 256, 128, 3),
    (1, 32, 64, 128, 256, 3),
    (2, 24, 128, 256, 256, 3),
    (1, 16, 256, 512, 256, 3),
    (3, 48, 24, 384, 128, 3),
])
def test_conv_single_segment(B, H, W, Cin, Cout, k):
    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5).to(torch.bfloat16)
    b = torch.randn(Cout, device=dev)
    wp = pack_conv_weight([(w, k * k)], npad=Cout)
    out = torch.empty(B, H, W, Cout, device=dev, dtype=torch.bfloat16)
    conv_igemm([(x, 0, Cin, k * k)], wp, b, out)
    torch.cuda.synchronize()
    ref = _ref_conv(x, w, b)
    err = (out.float() - ref).abs().max().item()
    assert err <= 1e-2 * ref.abs().max().item() + 1e-3, err


def test_conv_two_sources_plus_skip():
    """conv1 (3x3 on a) + 1x1 skip on the virtual concat [xa, xb] accumulated in one TMEM tile"""
    torch.manual_seed(1)
    dev = "cuda"
    B, H, W = 2, 32, 32
    a = torch.randn(B, H, W, 256, device=dev).to(torch.bfloat16)
    xa = torch.randn(B, H, W, 256, device=dev).to(torch.bfloat16)
    xb = torch.randn(B, H, W, 64, device=dev).to(torch.bfloat16)
    w1 = (torch.randn(256, 256, 3, 3, device=dev) / 48).to(torch.bfloat16)
    w2 = (torch.randn(256, 320, 1, 1, device=dev) / 18).to(torch.bfloat16)
    b = torch.randn(256, device=dev)
    wp = pack_conv_weight([(w1, 9), (w2[:, :256], 1), (w2[:, 256:], 1)], npad=256)
    out = torch.empty(B, H, W, 256, device=dev, dtype=torch.bfloat16)
    conv_igemm([(a, 0, 256, 9), (xa, 0, 256, 1), (xb, 0, 64, 1)], wp, b, out)
    torch.cuda.synchronize()
    ref = _ref_conv(a, w1, b) + _ref_conv(torch.cat([xa, xb], -1), w2, None)
    err = (out.float() - ref).abs().max().item()
    assert err <= 1e-2 * ref.abs().max().item() + 1e-3, err


def test_conv_f32_out_4ch():
    torch.manual_seed(2)
    dev = "cuda"
    B, H, W, Cin = 2, 48, 16, 256
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(4, Cin, 3, 3, device=dev) / 48).to(torch.bfloat16)
    b = torch.randn(4, device=dev)
    wp = pack_conv_weight([(w, 9)], npad=16)
    bias16 = torch.zeros(16, device=dev)
    bias16[:4] = b
    out = torch.empty(B, H, W, 4, device=dev, dtype=torch.float32)
    conv_igemm([(x, 0, Cin, 9)], wp, bias16, out)
    torch.cuda.synchronize()
    ref = _ref_conv(x, w, b)
    err = (out - ref).abs().max().item()
    assert err <= 1e-4 * ref.abs().max().item() + 1e-4, err


def test_conv_epilogue_stats():
    """GroupNorm partial sums written by the conv epilogue == column sums of the output"""
    from flowdec_b200 import ops
    torch.manual_seed(4)
    dev = "cuda"
    B, H, W, C = 2, 32, 64, 256
    x = torch.randn(B, H, W, 128, device=dev).to(torch.bfloat16)
    w = (torch.randn(C, 128, 3, 3, device=dev) / 34).to(torch.bfloat16)
    b = torch.randn(C, device=dev)
    wp = pack_conv_weight([(w, 9)], npad=C)
    out = torch.empty(B, H, W, C, device=dev, dtype=torch.bfloat16)
    S = ops.conv_stats_slabs(H, W)
    st = torch.full((B, S, C, 2), float("nan"), device=dev)
    conv_igemm([(x, 0, 128, 9)], wp, b, out, stats=st)
    torch.cuda.synchronize()
    ref = _ref_conv(x, w, b).double()
    tot = st.double().sum(1)
    assert torch.allclose(tot[..., 0], ref.sum((1, 2)), rtol=1e-5, atol=1e-2)
    assert torch.allclose(tot[..., 1], (ref * ref).sum((1, 2)), rtol=1e-5, atol=1e-2)
    red = ops.slab_reduce(st, 16, torch.empty(B, 16, C, 2, device=dev))
    assert torch.allclose(red.double().sum(1), tot, rtol=1e-6, atol=1e-3)


def test_conv_many_tiles_persistent():
    """more tiles than SMs -> exercises the TMEM double buffer and the smem ring wrap"""
    torch.manual_seed(3)
    dev = "cuda"
    B, H, W, C = 4, 96, 128, 256
    x = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
    w = (torch.randn(C, C, 3, 3, device=dev) / 48).to(torch.bfloat16)
    b = torch.randn(C, device=dev)
    wp = pack_conv_weight([(w, 9)], npad=C)
    out = torch.empty(B, H, W, C, device=dev, dtype=torch.bfloat16)
    conv_igemm([(x, 0, C, 9)], wp, b, out)
    torch.cuda.synchronize()
    ref = _ref_conv(x, w, b)
    err = (out.float() - ref).abs().max().item()
    assert err <= 1e-2 * ref.abs().max().item() + 1e-3, err


def test_pyramid_conv_gemm_first_shift_after():
    """3x3 conv to 4 channels as 36 per-tap tensor-core products + gather-sum (+ FIR-up pyramid)"""
    from flowdec_b200 import ops
    from oracle import flowdec_oracle as O
    torch.manual_seed(5)
    dev = "cuda"
    B, H, W, Cin = 2, 32, 24, 256
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    w = (torch.randn(4, Cin, 3, 3, device=dev) / 48).to(torch.bfloat16)
    b = torch.randn(4, device=dev)
    lo = torch.randn(B, H // 2, W // 2, 4, device=dev)
    part = torch.empty(B, H, W, 36, device=dev)
    conv_igemm([(x, 0, Cin, 1)], ops.pack_tap_weight(w.float()), None, part)
    out = torch.empty(B, H, W, 4, device=dev)
    ops.pyramid_gather(part, b, None, out)
    out2 = torch.empty(B, H, W, 4, device=dev)
    ops.pyramid_gather(part, b, lo, out2)
    torch.cuda.synchronize()
    ref = _ref_conv(x, w, b)
    assert (out - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-4
    ref2 = ref + O.fir_up2(lo.cpu().permute(0, 3, 1, 2)).permute(0, 2, 3, 1).to(dev)
    assert (out2 - ref2).abs().max().item() <= 1e-4 * ref2.abs().max().item() + 1e-4


@pytest.mark.parametrize("B,H,W,C1,C2,Cout", [(1, 16, 16, 64, 0, 256), (2, 32, 48, 256, 0, 256), (2, 48, 8, 256, 64, 256),
                                              (1, 96, 32, 128, 256, 128), (3, 16, 64, 256, 256, 256)])
def test_conv_fused_groupnorm_silu_operand(B, H, W, C1, C2, Cout):
    """halo kernel with scale_shift: conv(SiLU(x*scale+shift)) incl. zero padding of the ACTIVATED
    tensor, over a virtual concat of two raw sources, plus a raw 1x1 skip segment."""
    from flowdec_b200 import ops
    if not (ops.CTA_PAIRS and ops.HALO_TILES):
        pytest.skip("fused operand transform exists in the halo kernel only")
    torch.manual_seed(6)
    dev = "cuda"
    C = C1 + C2
    x1 = (torch.randn(B, H, W, C1, device=dev) * 1.3 + 0.2).to(torch.bfloat16)
    x2 = (torch.randn(B, H, W, C2, device=dev) * 0.7).to(torch.bfloat16) if C2 else None
    ss = torch.empty(B, C, 2, device=dev)
    ss[..., 0] = 0.5 + torch.rand(B, C, device=dev)
    ss[..., 1] = 0.3 * torch.randn(B, C, device=dev)
    w = (torch.randn(Cout, C, 3, 3, device=dev) / (3 * C ** 0.5)).to(torch.bfloat16)
    w2 = (torch.randn(Cout, C1, 1, 1, device=dev) / C1 ** 0.5).to(torch.bfloat16)
    b = torch.randn(Cout, device=dev)
    segs = [(w[:, :C1], 9)] + ([(w[:, C1:], 9)] if C2 else []) + [(w2, 1)]
    wp = pack_conv_weight(segs, npad=Cout)
    srcs = [(x1, 0, C1, 9, ss, 0)] + ([(x2, 0, C2, 9, ss, C1)] if C2 else []) + [(x1, 0, C1, 1)]
    out = torch.empty(B, H, W, Cout, device=dev, dtype=torch.bfloat16)
    conv_igemm(srcs, wp, b, out)
    torch.cuda.synchronize()
    xc = torch.cat([x1, x2], -1).float() if C2 else x1.float()
    a = F.silu(xc * ss[:, None, None, :, 0] + ss[:, None, None, :, 1]).to(torch.bfloat16)
    ref = _ref_conv(a, w, b) + _ref_conv(x1, w2, None)
    err = (out.float() - ref).abs().max().item()
    assert err <= 1.5e-2 * ref.abs().max().item() + 1e-3, err
