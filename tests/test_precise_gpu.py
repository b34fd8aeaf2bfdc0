"""GPU: the tf32 ("precise") form of the tcgen05 halo conv and the 4-channel fp32-output form (pyramid convs), against
torch F.conv2d in fp64 on the same inputs.

Tolerances: tf32 operands (10-bit mantissa, weights rounded to nearest at pack time, activations truncated by the
tensor pipe or rounded by the fused transform), fp32 accumulation: rel-L2 <= 1.5e-3 (bf16 gives ~4e-3 on the same
data); the end-to-end gates of the precise mode are in test_precise_model_* below."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from flowdec_b200 import ops

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def _gn_ss(xs_nhwc, gamma, beta):
    B, H, W = xs_nhwc[0].shape[:3]
    C = sum(x.shape[3] for x in xs_nhwc)
    xc = torch.cat([x.double() for x in xs_nhwc], 3).permute(0, 3, 1, 2)
    gn = F.group_norm(xc, min(C // 4, 32), gamma.double(), beta.double(), eps=1e-6)
    G = min(C // 4, 32)
    xg = xc.reshape(B, G, -1)
    mean, var = xg.mean(-1), xg.var(-1, unbiased=False)
    rstd = (var + 1e-6).rsqrt().repeat_interleave(C // G, 1)
    mean = mean.repeat_interleave(C // G, 1)
    scale = gamma.double()[None] * rstd
    shift = beta.double()[None] - mean * scale
    return torch.stack([scale, shift], -1).float().contiguous(), F.silu(gn)


@pytest.mark.parametrize("B,H,W,segs,cout,xf", [
    (2, 16, 16, [64], 256, False), (1, 32, 24, [128, 64], 128, False), (2, 48, 8, [256], 256, True),
    (1, 16, 32, [256, 64], 256, True), (2, 96, 16, [128], 128, True)])
def test_conv_tf32(B, H, W, segs, cout, xf):
    g = torch.Generator().manual_seed(1)
    xs = [torch.randn(B, H, W, c, generator=g) * 1.2 + 0.1 for c in segs]
    cin = sum(segs)
    w = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)
    skip_w = torch.randn(cout, cin, 1, 1, generator=g) / math.sqrt(cin)
    bias = torch.randn(cout, generator=g) * 0.1
    gamma, beta = 1 + 0.1 * torch.randn(cin, generator=g), 0.1 * torch.randn(cin, generator=g)
    ss, act = _gn_ss(xs, gamma, beta)
    xc = torch.cat(xs, 3).double().permute(0, 3, 1, 2)
    main_in = act if xf else xc
    ref = F.conv2d(main_in, w.double(), bias.double(), padding=1) + F.conv2d(xc, skip_w.double())
    # K order: 3x3 segments over the sources, then 1x1 skip segments over the same sources (raw)
    wsegs, c0 = [], 0
    for c in segs:
        wsegs.append((w[:, c0:c0 + c].contiguous(), 9))
        c0 += c
    c0 = 0
    for c in segs:
        wsegs.append((skip_w[:, c0:c0 + c].contiguous(), 1))
        c0 += c
    if len(wsegs) > 4:
        pytest.skip("more than 4 K segments")
    wp = ops.pack_conv_weight(wsegs, cout, tf32=True).cuda()
    xd = [x.cuda().contiguous() for x in xs]
    ssd = ss.cuda()
    srcs, off = [], 0
    for x in xd:
        srcs.append((x, 0, x.shape[3], 9, ssd, off) if xf else (x, 0, x.shape[3], 9))
        off += x.shape[3]
    srcs += [(x, 0, x.shape[3], 1) for x in xd]
    out = torch.empty(B, H, W, cout, device="cuda", dtype=torch.float32)
    S = ops.conv_stats_slabs(H, W)
    stats = torch.empty(B, S, cout, 2, device="cuda")
    ops.conv_igemm(srcs, wp, bias.cuda(), out, stats=stats)
    torch.cuda.synchronize()
    got = out.cpu().permute(0, 3, 1, 2)
    r = rel_l2(got, ref)
    print(f"\ntf32 conv {segs}->{cout} xf={xf}: rel-L2 {r:.3e}")
    assert r <= 1.5e-3
    # fused GroupNorm partial sums of the fp32 output
    st = stats.sum(1).cpu()
    assert torch.allclose(st[..., 0], out.cpu().sum((1, 2)), rtol=1e-3, atol=1e-2)
    assert torch.allclose(st[..., 1], out.cpu().pow(2).sum((1, 2)), rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("tf32", [False, True])
@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 128), (1, 96, 32, 256), (2, 32, 8, 64)])
def test_conv_out4_pyramid_form(B, H, W, C, tf32):
    """3x3 conv C -> 4 with fp32 [B,H,W,4] output and GroupNorm+SiLU fused into the operand (halo kernel, N = 16)"""
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, H, W, C, generator=g) * 1.5 - 0.2
    if not tf32:
        x = x.to(torch.bfloat16)
    w = torch.randn(4, C, 3, 3, generator=g) / math.sqrt(9 * C)
    bias = torch.randn(4, generator=g) * 0.1
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    ss, act = _gn_ss([x.float()], gamma, beta)
    ref = F.conv2d(act, w.double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    wp = ops.pack_conv_weight([(w, 9)], 16, tf32=tf32).cuda()
    b16 = torch.zeros(16)
    b16[:4] = bias
    out = torch.empty(B, H, W, 4, device="cuda", dtype=torch.float32)
    ops.conv_igemm([(x.cuda(), 0, C, 9, ss.cuda(), 0)], wp, b16.cuda(), out)
    torch.cuda.synchronize()
    r = rel_l2(out.cpu(), ref)
    print(f"\nout4 conv C={C} tf32={tf32}: rel-L2 {r:.3e}")
    assert r <= (1.5e-3 if tf32 else 8e-3)


# ------------------------------------------------------------------------------------------------
# model level: the tf32 mode explains the bf16 numbers (SURVEY.md §7 hard part 3 / §8c)
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def model():
    from flowdec_b200.model import build_flowdec
    from flowdec_b200.util.synth import synth_state_dict
    m = build_flowdec("75m")
    m.load_state_dict(synth_state_dict(m.state_dict(), seed=0))
    return m.cuda()


def test_precise_backbone_vs_golden(model):
    from oracle.make_golden import golden_inputs
    I = golden_inputs()
    gold = torch.from_numpy(np.load(os.path.join(GOLD, "flowdec_75m_seed0.npz"))["backbone_v"])
    x, y, t = I["X"].cuda(), I["Y"].cuda(), I["t"].cuda()
    try:
        model.set_precision("tf32")
        v32 = torch.view_as_real(model.backbone(x, y, t).cpu())
    finally:
        model.set_precision("bf16")
    v16 = torch.view_as_real(model.backbone(x, y, t).cpu())
    r32, r16 = rel_l2(v32, gold), rel_l2(v16, gold)
    print(f"\nbackbone rel-L2 vs reference golden: tf32 {r32:.3e}, bf16 {r16:.3e}")
    assert r32 <= 3e-3 and r16 <= 2e-2 and r32 < r16 / 3


@pytest.mark.parametrize("N,solver,key,file", [(1, "euler", "enhance_euler_N1", "flowdec_75m_seed0.npz"),
                                               (3, "midpoint", "enhance_midpoint_N3", "flowdec_75m_headline.npz")])
def test_precise_enhance_vs_golden(model, N, solver, key, file):
    from oracle import metrics as M
    from oracle.make_golden import golden_inputs
    I = golden_inputs()
    gold = torch.from_numpy(np.load(os.path.join(GOLD, file))[key])
    try:
        model.set_precision("tf32")
        x32 = model.enhance(I["y"], N=N, solver=solver, noise=I["eps"])
        x32b = model.enhance(I["y"], N=N, solver=solver, noise=I["eps"])      # CUDA-graph replay path
    finally:
        model.set_precision("bf16")
    x16 = model.enhance(I["y"], N=N, solver=solver, noise=I["eps"])
    s32, s16 = M.snr_db(x32, gold), M.snr_db(x16, gold)
    print(f"\nenhance {solver} N={N}: SNR vs reference golden tf32 {s32:.2f} dB, bf16 {s16:.2f} dB")
    assert torch.equal(x32, x32b)
    assert s32 >= 50.0 and s16 >= 30.0 and s32 >= s16 + 10.0


@pytest.mark.parametrize("tf32", [False, True])
@pytest.mark.parametrize("B,H,W,C", [(2, 16, 16, 128), (1, 96, 32, 256)])
def test_conv_out36_gemm_first_pyramid_form(B, H, W, C, tf32):
    """the production pyramid path: 1-tap halo GEMM with N = 48 (36 per-tap products per pixel, GroupNorm+SiLU fused
    into the operand) followed by fd_pyramid_gather (shifted sum + bias + FIR-up of the coarser pyramid)"""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, H, W, C, generator=g) * 1.5 - 0.2
    if not tf32:
        x = x.to(torch.bfloat16)
    w = torch.randn(4, C, 3, 3, generator=g) / math.sqrt(9 * C)
    bias = torch.randn(4, generator=g) * 0.1
    lo = torch.randn(B, H // 2, W // 2, 4, generator=g)
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    ss, act = _gn_ss([x.float()], gamma, beta)
    from oracle import flowdec_oracle as O
    ref = F.conv2d(act, w.double(), bias.double(), padding=1) + O.fir_up2(lo.permute(0, 3, 1, 2)).double()
    wt = ops.pack_tap_weight(w, tf32=tf32).cuda()
    part = torch.empty(B, H, W, 36, device="cuda", dtype=torch.float32)
    ops.conv_igemm([(x.cuda(), 0, C, 1, ss.cuda(), 0)], wt, None, part)
    out = ops.pyramid_gather(part, bias.cuda(), lo.cuda(), torch.empty(B, H, W, 4, device="cuda"))
    torch.cuda.synchronize()
    r = rel_l2(out.cpu().permute(0, 3, 1, 2), ref)
    print(f"\nout36 + gather C={C} tf32={tf32}: rel-L2 {r:.3e}")
    assert r <= (1.5e-3 if tf32 else 8e-3)


def test_precise_mode_microbatches_and_lanes(model):
    """tf32 mode through the micro-batch / two-lane machinery: a batch split over lanes equals the clips enhanced alone"""
    from flowdec_b200.util.synth import synth_waveforms
    y = synth_waveforms(3, 24000, seed=42)
    eps = torch.randn(3, 1, 768, 64, dtype=torch.complex64, generator=torch.Generator().manual_seed(11))
    old = model.max_batch
    try:
        model.set_precision("tf32")
        model.max_batch = 2
        full = model.enhance(y, N=1, solver="midpoint", noise=eps)
        one = model.enhance(y[2:3], N=1, solver="midpoint", noise=eps[2:3])
        assert torch.isfinite(full).all() and torch.equal(full[2:3], one)
    finally:
        model.max_batch = old
        model.set_precision("bf16")
