"""GPU: the shape-generic kernels (csrc/fd_generic.cu) and the 7-level / bottleneck-attention / 3x3-output NCSN++
(config/model/backbone/ncsnpp_default_ycond.yaml) against torch fp32 on CPU and the golden vector
produced by the unmodified reference (tests/golden/ncsnpp_attn_seed4.npz)."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from flowdec_b200 import ops
from flowdec_b200.backbones.ncsnpp import NCSNpp
from flowdec_b200.util.synth import synth_state_dict
from oracle import flowdec_oracle as O
from oracle.make_golden import ATTN_KW, attn_golden_inputs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ncsnpp_attn_seed4.npz")


def rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def nhwc_bf16(x):
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


@pytest.mark.parametrize("B,H,W,segs,cout,taps,f32", [
    (2, 12, 4, [256], 256, 9, False), (1, 24, 8, [128, 256], 128, 9, False), (2, 5, 7, [16], 32, 9, False),
    (1, 3, 1, [64, 8], 24, 1, True), (2, 16, 16, [32], 4, 9, True), (1, 6, 10, [72], 130, 9, False)])
def test_conv_direct(B, H, W, segs, cout, taps, f32):
    g = torch.Generator().manual_seed(3)
    xs = [torch.randn(B, c, H, W, generator=g) for c in segs]
    cin = sum(segs)
    w = torch.randn(cout, cin, 3 if taps == 9 else 1, 3 if taps == 9 else 1, generator=g) / math.sqrt(cin * taps)
    bias = torch.randn(cout, generator=g) * 0.1
    srcs = [nhwc_bf16(x).cuda() for x in xs]
    c0, wsegs = 0, []
    for c in segs:
        wsegs.append((w[:, c0:c0 + c].contiguous(), taps))
        c0 += c
    wp = ops.pack_conv_weight(wsegs, cout).cuda()
    out = torch.empty(B, H, W, cout, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    ops.conv_direct([(s, 0, s.shape[3], taps) for s in srcs], wp, bias.cuda(), out)
    xr = torch.cat([s.float().cpu().permute(0, 3, 1, 2) for s in srcs], 1)
    ref = F.conv2d(xr, w.to(torch.bfloat16).float(), bias, padding=1 if taps == 9 else 0)
    got = out.float().cpu().permute(0, 3, 1, 2)
    assert rel_l2(got, ref) < (1e-5 if f32 else 4e-3)


def test_conv_direct_fused_groupnorm():
    """sources with scale/shift: SiLU(GroupNorm(x)) (or the affine alone) applied on load, channel sub-ranges"""
    g = torch.Generator().manual_seed(4)
    B, H, W, C, cout = 2, 6, 5, 48, 40
    x = torch.randn(B, C, H, W, generator=g) * 1.3 + 0.2
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    w = torch.randn(cout, C, 3, 3, generator=g) / math.sqrt(9 * C)
    s = nhwc_bf16(x).cuda()
    part = ops.chan_stats(s, 4)
    ss = torch.empty(B, C, 2, device="cuda")
    ops.gn_finalize([part], [C], H * W, gamma.cuda(), beta.cuda(), min(C // 4, 32), 1e-6, ss)
    wp = ops.pack_conv_weight([(w, 9)], cout).cuda()
    xr = s.float().cpu().permute(0, 3, 1, 2)
    gn = F.group_norm(xr, min(C // 4, 32), gamma, beta, eps=1e-6)
    for affine_only, act in ((False, F.silu(gn)), (True, gn)):
        out = torch.empty(B, H, W, cout, device="cuda", dtype=torch.float32)
        ops.conv_direct([(s, 0, C, 9, ss, 0)], wp, None, out, affine_only=affine_only)
        ref = F.conv2d(act, w.to(torch.bfloat16).float(), None, padding=1)
        assert rel_l2(out.cpu().permute(0, 3, 1, 2), ref) < 5e-3      # operand rounded to bf16 after the transform


@pytest.mark.parametrize("B,H,W,C", [(2, 12, 4, 256), (1, 2, 2, 32), (1, 12, 59, 256), (3, 1, 1, 64)])
def test_attention_core(B, H, W, C):
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(B, H, W, 3 * C, generator=g)
    out = ops.attention(qkv.cuda(), torch.empty(B, H, W, C, device="cuda", dtype=torch.bfloat16))
    q, k, v = (t.reshape(B, H * W, C) for t in qkv.split(C, dim=-1))
    w = torch.softmax(torch.einsum("bic,bjc->bij", q, k) * C ** -0.5, dim=-1)
    ref = torch.einsum("bij,bjc->bic", w, v).reshape(B, H, W, C)
    assert (out.float().cpu() - ref).abs().max() <= 2 ** -8 * ref.abs().max() + 1e-4


@pytest.mark.parametrize("H,W,C1,C2", [(2, 2, 32, 0), (6, 2, 16, 16), (10, 14, 64, 8)])
def test_gn_act_down_any(H, W, C1, C2):
    g = torch.Generator().manual_seed(6)
    B, C = 2, C1 + C2
    xs = [torch.randn(B, c, H, W, generator=g) for c in (C1, C2) if c]
    srcs = [nhwc_bf16(x).cuda() for x in xs]
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    parts = [ops.chan_stats(s, 2) for s in srcs]
    ss = torch.empty(B, C, 2, device="cuda")
    ops.gn_finalize(parts, [s.shape[3] for s in srcs], H * W, gamma.cuda(), beta.cuda(), min(C // 4, 32), 1e-6, ss)
    out = torch.empty(B, H // 2, W // 2, C, device="cuda", dtype=torch.bfloat16)
    raw = torch.empty_like(out)
    ops.gn_act_resample(srcs, ss, out, 1, out_raw=raw)
    xc = torch.cat([s.float().cpu().permute(0, 3, 1, 2) for s in srcs], 1)
    h = O.fir_down2(F.silu(F.group_norm(xc, min(C // 4, 32), gamma, beta, eps=1e-6)))
    r = O.fir_down2(xc)
    assert (out.float().cpu().permute(0, 3, 1, 2) - h).abs().max() <= 2 ** -8 * h.abs().max() + 1e-3
    assert (raw.float().cpu().permute(0, 3, 1, 2) - r).abs().max() <= 2 ** -8 * r.abs().max() + 1e-3


@pytest.mark.parametrize("cout", [16, 128])
def test_conv_in_any_and_output_conv3(cout):
    g = torch.Generator().manual_seed(7)
    B, H, W = 2, 12, 20
    x4 = torch.randn(B, H, W, 4, generator=g)
    w = torch.randn(cout, 4, 3, 3, generator=g) / 6
    b = torch.randn(cout, generator=g) * 0.1
    out = ops.conv_in(x4.cuda(), w.cuda(), b.cuda(), torch.empty(B, H, W, cout, device="cuda", dtype=torch.bfloat16))
    ref = F.conv2d(x4.permute(0, 3, 1, 2), w, b, padding=1)
    assert (out.float().cpu().permute(0, 3, 1, 2) - ref).abs().max() <= 2 ** -8 * ref.abs().max() + 1e-4
    # 3x3 output layer fused with an axpy stage
    wo = torch.randn(2, 4, 3, 3, generator=g) / 6
    base = torch.randn(B, H, W, 2, generator=g)
    o = torch.empty(B, H, W, 2, device="cuda")
    v = torch.empty(B, H, W, 2, device="cuda")
    ops.output_conv3_axpy(x4.cuda(), wo.cuda(), base.cuda(), 0.5, None, 0.0, -0.25, o, v_out=v)
    vr = F.conv2d(x4.permute(0, 3, 1, 2), wo, None, padding=1).permute(0, 2, 3, 1)
    assert rel_l2(v.cpu(), vr) < 1e-5 and rel_l2(o.cpu(), 0.5 * base - 0.25 * vr) < 1e-5


def _attn_model():
    G = np.load(GOLD)
    tmpl = {str(k): torch.empty([int(d) for d in str(s).split(",") if d], dtype=torch.float32)
            for k, s in zip(G["keys"], G["shapes"])}
    sd = synth_state_dict(tmpl, seed=4)
    net = NCSNpp(nonlinearity="swish", attn_resolutions=[], num_channels=4, **ATTN_KW)
    assert {"backbone." + k for k in net.state_dict()} == set(sd)
    net.load_state_dict({k[len("backbone."):]: v for k, v in sd.items()})
    return net.cuda().eval(), sd, torch.from_numpy(G["backbone_v"])


def test_attention_backbone_vs_reference_golden():
    net, sd, gold = _attn_model()
    I = attn_golden_inputs()
    v = net(I["X"].cuda(), I["Y"].cuda(), I["t"].cuda())
    r = rel_l2(torch.view_as_real(v.cpu()), gold)
    print(f"\n7-level / attention backbone rel-L2 vs reference golden: {r:.4e}")
    assert r <= 2e-2
    # batch of 2 with a different width (bottleneck 2 x 3 tokens), against the CPU oracle
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 1, 128, 192, dtype=torch.complex64, generator=g)
    y = torch.randn(2, 1, 128, 192, dtype=torch.complex64, generator=g)
    with torch.no_grad():
        ref = O.ncsnpp_forward(sd, x, y, torch.tensor([0.3]), num_resolutions=7, num_res_blocks=2, bottleneck_attn=True)
    v2 = net(x.cuda(), y.cuda(), torch.tensor([0.3]).cuda())
    r2 = rel_l2(torch.view_as_real(v2.cpu()), torch.view_as_real(ref))
    print(f"7-level / attention backbone (B=2, 128x192) rel-L2 vs oracle: {r2:.4e}")
    assert r2 <= 2e-2


def test_attention_backbone_tensor_core_levels():
    """same architecture at a width the tcgen05 tiles take (nf=64 -> 64/128-channel levels; 256x128 image):
    the large levels run on tensor cores, the 8x4 / 4x2 levels and the attention on the generic kernels.
    Checked against the CPU oracle with the same synthetic weights."""
    kw = dict(ATTN_KW)
    kw.update(image_size=256, nf=128, ch_mult=[1, 1, 2, 2, 2, 2, 2])
    net = NCSNpp(nonlinearity="swish", attn_resolutions=[], num_channels=4, **kw)
    sd = synth_state_dict({"backbone." + k: v for k, v in net.state_dict().items()}, seed=5)
    net.load_state_dict({k[len("backbone."):]: v for k, v in sd.items()})
    net = net.cuda().eval()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 1, 256, 128, dtype=torch.complex64, generator=g)
    y = torch.randn(1, 1, 256, 128, dtype=torch.complex64, generator=g)
    with torch.no_grad():
        ref = O.ncsnpp_forward(sd, x, y, torch.tensor([0.7]), num_resolutions=7, num_res_blocks=2, bottleneck_attn=True)
    launches = []
    old = ops.conv_igemm

    def spy(*a, **k):
        launches.append(1)
        return old(*a, **k)
    ops.conv_igemm = spy
    try:
        v = net(x.cuda(), y.cuda(), torch.tensor([0.7]).cuda())
    finally:
        ops.conv_igemm = old
    r = rel_l2(torch.view_as_real(v.cpu()), torch.view_as_real(ref))
    print(f"\n7-level nf=128 backbone: {len(launches)} tcgen05 conv launches, rel-L2 vs oracle {r:.4e}")
    assert len(launches) >= 20 and r <= 2e-2
