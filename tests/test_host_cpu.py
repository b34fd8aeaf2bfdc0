"""CPU tests of the host-side mirror: state_dict contract, solver schedules, C-ABI symbols."""
import os
import re

import numpy as np
import pytest
import torch

from flowdec_b200 import _lib
from flowdec_b200.model import build_flowdec
from flowdec_b200.sampling import solvers
from flowdec_b200.util.synth import synth_state_dict, synth_waveforms
from oracle import flowdec_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "flowdec_b200.h")).read()
    declared = set(re.findall(r"\b(fd_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared - {"fd_last_error"} == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert L.fd_abi_version() == 1


def test_state_dict_contract():
    m = build_flowdec("75m")
    sd = m.state_dict()
    assert len(sd) == 265
    assert sd["sigma_y"].dtype == torch.float64 and tuple(sd["sigma_y"].shape) == (768, 1)
    assert tuple(sd["backbone.output_layer.weight"].shape) == (2, 4, 1, 1)
    assert tuple(sd["feature_extractor.complex_stft.window"].shape) == (1534,)
    assert tuple(sd["backbone.all_modules.32.Conv_0.weight"].shape) == (256, 320, 3, 3)
    assert "backbone.all_modules.34.bias" in sd and "backbone.all_modules.14.Conv_2.weight" not in sd
    n = sum(v.numel() for k, v in sd.items() if k.startswith("backbone."))
    assert n == 23_703_704
    m2 = build_flowdec("25s")
    assert not torch.equal(m2.state_dict()["sigma_y"], sd["sigma_y"])
    # loading a reference-style EMA dict round-trips
    syn = synth_state_dict(sd, seed=0)
    m.load_state_dict(syn)
    assert all(torch.equal(m.state_dict()[k], syn[k]) for k in syn)


def test_synth_is_order_independent():
    m = build_flowdec("75m")
    sd = m.state_dict()
    a = synth_state_dict(sd, seed=0)
    b = synth_state_dict(dict(reversed(list(sd.items()))), seed=0)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(synth_waveforms(3, 1000, seed=5)[2], synth_waveforms(1, 1000, seed=7)[0])


@pytest.mark.parametrize("solver", ["euler", "midpoint", "heun2", "heun2_eulerlast"])
@pytest.mark.parametrize("N", [1, 3, 4])
def test_solver_schedule_matches_oracle(solver, N):
    """run the fused-stage schedule on a scalar ODE and compare with the oracle's stepper"""
    f = lambda t, x: torch.sin(3 * t) * x + t
    x0 = torch.tensor([0.7], dtype=torch.float32)
    ref = O.ode_solve(f, x0, N, solver)[-1]
    bufs = {"x": x0.clone(), "xn": None, "tmp": None}
    nfe = 0
    for (t, dt) in solvers.t_grid(N):
        for (te, src, dst, b1, c1, b2, c2, coef) in solvers.stages(solver, t, dt):
            v = f(torch.tensor(te), bufs[src])
            nfe += 1
            r = np.float32(coef) * v
            if b1 is not None:
                r = r + np.float32(c1) * bufs[b1]
            if b2 is not None:
                r = r + np.float32(c2) * bufs[b2]
            bufs[dst] = r
        bufs["x"] = bufs["xn"]
    assert torch.allclose(bufs["x"], ref, rtol=2e-6, atol=1e-7)
    if solver in ("euler", "midpoint", "heun2"):
        assert nfe == N * solvers.nfe_per_step(solver)


def test_unsupported_configs_fail_loudly():
    from flowdec_b200.backbones.ncsnpp import NCSNpp
    with pytest.raises(NotImplementedError):
        NCSNpp()                       # reference defaults include attention
    with pytest.raises(ValueError):
        solvers.get_solver("rk4")
    m = build_flowdec("75m")
    with pytest.raises(RuntimeError):  # CPU model: there is no CPU fallback
        m.enhance(torch.zeros(1, 1, 24000), N=1)


def test_length_bucketed_batching_host_logic():
    """flowdec_b200/batching.py: buckets = padded STFT frame counts (pad_spec, util/other.py:25-52)"""
    import torch
    from flowdec_b200.batching import bucket_by_frames, frames_bucket, pad_batch
    assert [frames_bucket(n) for n in (768, 24191, 24192, 48000, 96000, 192000)] == [64, 64, 64, 128, 256, 512]
    assert frames_bucket(64 * 384 - 1) == 64 and frames_bucket(64 * 384) == 128
    lens = [96000, 48000, 95000, 100000, 24960, 96001, 97000]
    batches = bucket_by_frames(lens, 2)
    assert batches == [[3], [0, 2], [5, 6], [1, 4]]
    assert sorted(i for b in batches for i in b) == list(range(len(lens)))
    for b in batches:
        assert len({frames_bucket(lens[i]) for i in b}) == 1 and len(b) <= 2
    with pytest.raises(ValueError):
        bucket_by_frames([96000, 767], 4)
    with pytest.raises(ValueError):
        bucket_by_frames([96000], 0)
    y, l = pad_batch([torch.arange(5.0), torch.ones(1, 3)])
    assert y.shape == (2, 1, 5) and l == [5, 3] and y[1, 0].tolist() == [1, 1, 1, 0, 0]


def test_ragged_enhance_argument_checks():
    """FlowModel.enhance(lengths=): shape / length / bucket validation happens before any device work;
    without a CUDA device the call then fails loudly instead of falling back"""
    import torch
    m = build_flowdec("75m")
    y = torch.zeros(2, 1, 96000)
    with pytest.raises(ValueError, match="lengths for a batch"):
        m.enhance(y, N=1, lengths=[96000])
    with pytest.raises(ValueError, match="clip lengths must be in"):
        m.enhance(y, N=1, lengths=[96000, 700])
    with pytest.raises(ValueError, match="clip lengths must be in"):
        m.enhance(y, N=1, lengths=[96000, 96001])
    with pytest.raises(ValueError, match="padded-frame bucket"):
        m.enhance(y, N=1, lengths=[96000, 48000])
    with pytest.raises(ValueError, match="ragged batches are"):
        m.enhance(torch.zeros(2, 2, 96000), N=1, lengths=[96000, 96000])
    with pytest.raises(NotImplementedError):
        m.enhance(y, N=1, lengths=[96000, 96000], return_traj=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.enhance(y, N=1, lengths=[96000, 95000])


# ------------------------------------------------------------------------------------------------
# round-2 host logic
def test_backbone_plan_and_attention_keys():
    """the static channel walk used to pack every fused weight up front, and the parameter names of the
    7-level / bottleneck-attention configuration (must equal the reference's state_dict keys stored with the golden)"""
    import numpy as np
    from flowdec_b200.backbones.ncsnpp import NCSNpp
    from flowdec_b200.model import build_flowdec
    from oracle.make_golden import ATTN_KW
    plan = build_flowdec("75m").backbone._plan()
    assert len(plan) == 20 and plan[4] == [64] and plan[17] == [128, 256] and plan[31] == [256, 256] and plan[32] == [256, 64]
    net = NCSNpp(nonlinearity="swish", attn_resolutions=[], num_channels=4, **ATTN_KW)
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ncsnpp_attn_seed4.npz"))
    ref = {str(k): tuple(int(d) for d in str(s).split(",") if d) for k, s in zip(G["keys"], G["shapes"])}
    mine = {"backbone." + k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert mine == ref
    assert len(net._plan()) == 49
    with pytest.raises(NotImplementedError):
        NCSNpp(resblock_type="ddpm")
    with pytest.raises(NotImplementedError):
        NCSNpp(image_size=256, attn_resolutions=(64,))


def test_round_tf32_and_ndac_gemm_packing():
    """tf32 rounding helper and the GEMM forms of the NDAC layers (what fd_dac_conv_tc contracts), emulated with
    torch on CPU: k7 dilated conv, transposed conv as the 2-tap conv with s*Cout columns, strided conv as the 3-tap
    conv over the [T/s, s*C] view"""
    import math
    import torch.nn.functional as F
    from flowdec_b200.ops import round_tf32
    from flowdec_b200.ndac import DAC
    from flowdec_b200.util.synth import synth_dac_state_dict
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -10, -3.14159265, 1e-30, 65504.0])
    r = round_tf32(x)
    assert torch.equal(r.view(torch.int32) & 0x1FFF, torch.zeros(6, dtype=torch.int32))
    assert ((r - x).abs() <= x.abs() * 2 ** -11).all()

    def gemm(xt, wp, offs, Tout):        # xt [B,T,C]; out[b,t,n] = sum_tap sum_c xt[b, t+off, c] wp[n, tap*C + c]
        B, T, C = xt.shape
        out = torch.zeros(B, Tout, wp.shape[0], dtype=xt.dtype)
        for i, o in enumerate(offs):
            idx = torch.arange(Tout) + o
            ok = (idx >= 0) & (idx < T)
            sl = torch.zeros(B, Tout, C, dtype=xt.dtype)
            sl[:, ok] = xt[:, idx[ok]]
            out += sl @ wp[:, i * C:(i + 1) * C].t()
        return out

    sd = synth_dac_state_dict(64, 128, (4, 2), 2, seed=3, encoder_dim=32, encoder_rates=(2, 4))
    m = DAC(sd, decoder_dim=128, decoder_rates=(4, 2), n_codebooks=2, latent_dim=64, encoder_dim=32, encoder_rates=(2, 4))
    g = torch.Generator().manual_seed(1)
    dec = {}
    for lay in m._tc_prepare():
        dec.setdefault(lay[0], lay)        # first layer of each kind = m._ops[1] (convtr), m._ops[2] (res, dilation 1)
    # transposed conv (first decoder block: 128 -> 64, s = 4, pad 2)
    _, wp, b, a, s, pad, cout = dec["convtr"]
    w = m._t(m._ops[1][1]).double()
    xin = torch.randn(2, w.shape[0], 9, generator=g, dtype=torch.float64)
    ref = F.conv_transpose1d(xin, w, None, stride=s, padding=pad)
    z = gemm(xin.permute(0, 2, 1), wp.double(), [0, -1], xin.shape[-1] + 1).reshape(2, -1, cout)
    got = z[:, pad:pad + ref.shape[-1]].permute(0, 2, 1)
    assert (got - ref).abs().max() <= 2e-3 * ref.abs().max()
    # k7 dilated conv of a residual unit
    _, w1p, b1, a1, offs, w2p, b2, a2 = dec["res"]
    w1 = m._t(m._ops[2][1][0]).double()
    d = m._ops[2][1][3]
    xin = torch.randn(2, w1.shape[1], 40, generator=g, dtype=torch.float64)
    ref = F.conv1d(xin, w1, None, dilation=d, padding=3 * d)
    got = gemm(xin.permute(0, 2, 1), w1p.double(), offs, 40).permute(0, 2, 1)
    assert offs == [(j - 3) * d for j in range(7)] and (got - ref).abs().max() <= 2e-3 * ref.abs().max()
    # strided encoder conv (32 -> 64, s = 2, pad 1) over the [T/s, s*C] view
    enc = {}
    for lay in m._tc_enc_prepare():
        enc.setdefault(lay[0], lay)
    _, wdp, bd, ad, s = enc["down"]
    wd = m._t(next(op for op in m._enc_ops if op[0] == "down")[1]).double()
    xin = torch.randn(2, wd.shape[1], 24, generator=g, dtype=torch.float64)
    ref = F.conv1d(xin, wd, None, stride=s, padding=math.ceil(s / 2))
    view = xin.permute(0, 2, 1).reshape(2, 24 // s, s * wd.shape[1])
    got = gemm(view, wdp.double(), [-1, 0, 1], 24 // s).permute(0, 2, 1)
    assert got.shape == ref.shape and (got - ref).abs().max() <= 2e-3 * ref.abs().max()


def test_score_model_argument_contract():
    from flowdec_b200.model import build_scoredec
    sm = build_scoredec()
    with pytest.raises(NotImplementedError):
        sm.enhance(torch.zeros(1, 1, 24000), sampler_type="ode")
    with pytest.raises(NotImplementedError):
        sm.enhance(torch.zeros(1, 1, 24000), predictor="none")
    with pytest.raises(RuntimeError):                      # CUDA only, and says so
        sm.enhance(torch.zeros(1, 1, 24000), predictor="euler_maruyama")


def test_checkpoint_loaders(tmp_path):
    """DAC.load reads audiotools-style {"state_dict", "metadata": {"kwargs"}} files; load_from_checkpoint refuses a
    checkpoint whose backbone keys do not match instead of silently keeping random weights"""
    from flowdec_b200.model import EnhancementModel, build_flowdec
    from flowdec_b200.ndac import DAC
    from flowdec_b200.util.synth import synth_dac_state_dict, synth_state_dict
    sd = synth_dac_state_dict(64, 96, (4, 3, 2), 5, seed=1)
    p = tmp_path / "weights.pth"
    torch.save({"state_dict": sd, "metadata": {"kwargs": dict(decoder_dim=96, decoder_rates=[4, 3, 2], n_codebooks=5,
                                                                latent_dim=64, sample_rate=48000)}}, p)
    m = DAC.load(str(p))
    assert m.decoder_rates == [4, 3, 2] and m.n_codebooks == 5 and m.sample_rate == 48000 and not m.tc_eligible
    assert m.precision == "fp32"                      # 96 / 48 / 24 / 12 channels: the tensor-core decoder needs multiples of 32
    fm = build_flowdec("75m")
    good = synth_state_dict(fm.state_dict(), seed=0)
    ck = tmp_path / "model.ckpt"
    torch.save({"_pl_ema_state_dict": good, "state_dict": good}, ck)
    loaded = EnhancementModel.load_from_checkpoint(str(ck), map_location="cpu")
    assert torch.equal(loaded.state_dict()["backbone.all_modules.3.weight"], good["backbone.all_modules.3.weight"])
    bad = {("model." + k): v for k, v in good.items()}          # wrong prefix: nothing would match
    torch.save({"_pl_ema_state_dict": bad}, ck)
    with pytest.raises(RuntimeError):
        EnhancementModel.load_from_checkpoint(str(ck), map_location="cpu")
    torch.save({"other": 1}, ck)
    with pytest.raises(KeyError):
        EnhancementModel.load_from_checkpoint(str(ck), map_location="cpu")
