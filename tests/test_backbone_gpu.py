"""GPU parity of the full NCSN++ evaluation and of enhance() against the CPU oracle and the
golden vectors produced by the unmodified reference (tests/golden, oracle/make_golden.py).

Stated tolerances (bf16 tensor-core operands, fp32 accumulation, bf16
activation storage):  one backbone evaluation rel-L2 <= 3e-2 vs the fp32 reference;
enhance() waveform SNR >= 25 dB at NFE <= 2 on the synthetic (untrained, amplitude-expanding)
network.  Measured values are printed (-s) and recorded in DESIGN.md."""
import os

import numpy as np
import pytest
import torch

from flowdec_b200.model import build_flowdec
from flowdec_b200.util.synth import synth_state_dict
from oracle.make_golden import golden_inputs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "flowdec_75m_seed0.npz")


@pytest.fixture(scope="module")
def model():
    m = build_flowdec("75m")
    m.load_state_dict(synth_state_dict(m.state_dict(), seed=0))
    return m.cuda()


def rel_l2(a, b):
    return ((a - b).norm() / b.norm()).item()


def snr_db(x, ref):
    return (10 * torch.log10(ref.pow(2).sum() / (x - ref).pow(2).sum())).item()


def test_backbone_forward_vs_golden_and_oracle(model):
    I = golden_inputs()
    gold = torch.from_numpy(np.load(GOLD)["backbone_v"])
    with torch.no_grad():
        v = model.backbone(I["X"].cuda(), I["Y"].cuda(), I["t"].cuda())
    v = torch.view_as_real(v.cpu())
    r = rel_l2(v, gold)
    print(f"\nbackbone rel-L2 vs reference golden: {r:.4e}")
    assert r <= 3e-2
    # module API: FlowModel.forward takes a 0-dim t as well (model.py:470-474)
    with torch.no_grad():
        v2 = model(I["X"].cuda(), I["Y"].cuda(), torch.tensor(0.3).cuda())
    assert torch.equal(torch.view_as_real(v2.cpu()), v)


def test_backbone_per_tap_kernel_path(model):
    """non-default path: per-tap conv kernel + materialised GroupNorm/SiLU activations
    (ops.HALO_TILES = False) must meet the same parity bar as the default halo / fused path"""
    from flowdec_b200 import ops
    I = golden_inputs()
    gold = torch.from_numpy(np.load(GOLD)["backbone_v"])
    old = ops.HALO_TILES
    ops.HALO_TILES = False
    try:
        with torch.no_grad():
            v = model.backbone(I["X"].cuda(), I["Y"].cuda(), I["t"].cuda())
    finally:
        ops.HALO_TILES = old
    r = rel_l2(torch.view_as_real(v.cpu()), gold)
    print(f"\nbackbone (per-tap kernel, separate GroupNorm/SiLU pass) rel-L2 vs reference golden: {r:.4e}")
    assert r <= 3e-2


@pytest.mark.parametrize("N,solver", [(1, "euler"), (1, "midpoint"), (2, "heun2_eulerlast")])
def test_enhance_vs_golden(model, N, solver):
    I = golden_inputs()
    gold = torch.from_numpy(np.load(GOLD)[f"enhance_{solver}_N{N}"])
    x = model.enhance(I["y"], N=N, solver=solver, noise=I["eps"])
    assert x.shape == gold.shape and x.device == I["y"].device
    s = snr_db(x, gold)
    print(f"\nenhance {solver} N={N}: waveform SNR vs reference golden = {s:.2f} dB")
    assert s >= 25.0
    # second and third call go through CUDA-graph capture / replay: identical result
    x2 = model.enhance(I["y"], N=N, solver=solver, noise=I["eps"])
    x3 = model.enhance(I["y"], N=N, solver=solver, noise=I["eps"])
    assert torch.equal(x2, x) and torch.equal(x3, x)


def test_scoredec_pc_sampler_vs_golden(model):
    """ScoreDec baseline path (ScoreModel + OUVE SDE + PC sampler) around the same backbone.
    The untrained score network diverges (|x| ~ 1e5), so the gate is relative: SNR >= 20 dB."""
    from flowdec_b200.model import ScoreModel
    from flowdec_b200.sdes import OUVESDE
    I = golden_inputs()
    gold = torch.from_numpy(np.load(GOLD)["scoredec_pc_N2"])
    sm = ScoreModel(OUVESDE(theta=1.5, sigma_min=0.05, sigma_max=0.82, N=30), 3e-2, backbone=model.backbone,
                    feature_extractor=model.feature_extractor, sampling_rate=48000, lr=1e-4).cuda()
    x = sm.enhance(I["y"], N=2, snr=0.5, noise=I["score_draws"])
    # second / third call: the whole sampler loop as one captured CUDA graph, then its replay
    assert torch.equal(sm.enhance(I["y"], N=2, snr=0.5, noise=I["score_draws"]), x)
    assert torch.equal(sm.enhance(I["y"], N=2, snr=0.5, noise=I["score_draws"]), x)
    # Euler-Maruyama predictor = the same update for the OUVE SDE; probability flow runs and differs
    assert torch.equal(sm.enhance(I["y"], N=2, snr=0.5, noise=I["score_draws"], predictor="euler_maruyama"), x)
    xpf = sm.enhance(I["y"], N=2, snr=0.5, noise=I["score_draws"], probability_flow=True)
    assert torch.isfinite(xpf).all() and not torch.equal(xpf, x)
    s = snr_db(x, gold)
    print(f"\nScoreDec PC sampler N=2: waveform SNR vs reference golden = {s:.2f} dB")
    assert x.shape == gold.shape and s >= 20.0
    # score = -backbone / sigma_t (model.py:613-628)
    t = torch.tensor([0.5]).cuda()
    sc = sm(I["X"].cuda(), I["Y"].cuda(), t)
    v = model.backbone(I["X"].cuda(), I["Y"].cuda(), t)
    assert torch.allclose(sc, -v / float(sm.sde._std(0.5)))


def test_enhance_api_shapes_and_info(model):
    I = golden_inputs()
    y = I["y"]
    out, info = model.enhance(y[0], N=1, solver="euler", noise=I["eps"], return_preprocess_info=True,
                              predictor="x", corrector="y", snr=0.5)       # CLI extras are ignored
    assert out.shape == y[0].shape
    assert {"orig_length", "normfac", "undo_pad_fn", "squeeze_dims"} <= info.keys()
    assert info["orig_length"] == y.shape[-1] and info["squeeze_dims"] == 1
    out1 = model.enhance(y[0, 0], N=1, solver="euler", noise=I["eps"])
    assert out1.shape == y[0, 0].shape and torch.equal(out1, out[0])
    Xs, xs = model.enhance(y, N=2, solver="euler", noise=I["eps"], return_traj=True)
    assert Xs.shape[0] == 3 and len(xs) == 3 and xs[-1].shape == y.shape


def test_batch_independence_and_microbatching(model):
    """clips are independent end to end (per-sample normfac / GroupNorm): any batch split gives
    bit-identical waveforms -> the basis of multi-GPU sharding."""
    from flowdec_b200.util.synth import synth_waveforms
    y = synth_waveforms(3, 24000, seed=42)
    g = torch.Generator().manual_seed(11)
    eps = torch.randn(3, 1, 768, 64, dtype=torch.complex64, generator=g)
    full = model.enhance(y, N=1, solver="midpoint", noise=eps)
    model.max_batch = 2
    model.reset_cache()         # micro-batch size is baked into the cached graphs
    split = model.enhance(y, N=1, solver="midpoint", noise=eps)
    model.max_batch = 16
    one = model.enhance(y[1:2], N=1, solver="midpoint", noise=eps[1:2])
    assert torch.equal(full, split)
    assert torch.equal(full[1:2], one)


def test_long_and_ragged_clips(model):
    """a 12.3 s clip (Tp = 1536 frames, forces micro-batch 2) and a length that is not a hop multiple:
    finite output of the right shape, and the per-clip result does not depend on its batch neighbours"""
    from flowdec_b200.util.synth import synth_waveforms
    L = 590_411
    y = synth_waveforms(3, L, seed=77, kind="gauss")
    g = torch.Generator().manual_seed(3)
    Tp = 64 * -(-(1 + L // 384) // 64)
    eps = torch.randn(3, 1, 768, Tp, dtype=torch.complex64, generator=g)
    out = model.enhance(y, N=1, solver="euler", noise=eps)
    assert out.shape == y.shape and torch.isfinite(out).all()
    one = model.enhance(y[2:3], N=1, solver="euler", noise=eps[2:3])
    assert torch.equal(one, out[2:3])


def test_ragged_batch_matches_single_clips(model):
    """length-bucketed batching across files: clips of different lengths inside one
    padded-frame bucket share a batch, and each comes out bit-identical to enhancing it alone"""
    from flowdec_b200.batching import enhance_list, frames_bucket
    from flowdec_b200.util.synth import synth_waveforms
    lens = [48000, 45001, 41233, 48383, 24960 + 767]          # frames 126, 118, 108, 126, 67 -> Tp 128
    assert {frames_bucket(n) for n in lens} == {128}
    waves = [synth_waveforms(1, n, seed=100 + i)[0, 0] for i, n in enumerate(lens)]
    g = torch.Generator().manual_seed(5)
    eps = torch.randn(len(lens), 1, 768, 128, dtype=torch.complex64, generator=g)
    y = torch.zeros(len(lens), 1, max(lens))
    for i, w in enumerate(waves):
        y[i, 0, :lens[i]] = w
    out, info = model.enhance(y, N=1, solver="midpoint", noise=eps, lengths=lens, return_preprocess_info=True)
    assert out.shape == (len(lens), 1, max(lens)) and torch.isfinite(out).all()
    assert info["orig_length"] == lens
    for i, n in enumerate(lens):
        alone, info1 = model.enhance(waves[i].reshape(1, 1, n), N=1, solver="midpoint", noise=eps[i:i + 1],
                                     return_preprocess_info=True)
        assert torch.equal(out[i, :, :n], alone[0]), f"clip {i} (L={n}) differs from its single-clip result"
        assert (out[i, :, n:] == 0).all()
        assert torch.equal(info["normfac"][i].cpu(), info1["normfac"][0].cpu())
    # second call replays the captured graph with other lengths of the same bucket
    lens2 = [47000, 48000, 30000, 46001, 44444]
    y2 = torch.zeros(len(lens2), 1, 48000)
    for i, n in enumerate(lens2):
        y2[i, 0, :n] = waves[i][:n] if n <= lens[i] else synth_waveforms(1, n, seed=200 + i)[0, 0]
    for _ in range(2):
        out2 = model.enhance(y2, N=1, solver="midpoint", noise=eps, lengths=lens2)
    alone = model.enhance(y2[2:3, :, :30000], N=1, solver="midpoint", noise=eps[2:3])
    assert torch.equal(out2[2, :, :30000], alone[0])
    # mixed buckets are rejected by enhance(lengths=) and handled by enhance_list
    with pytest.raises(ValueError):
        model.enhance(torch.zeros(2, 1, 96000), N=1, lengths=[96000, 48000])
    clips = [waves[0], synth_waveforms(1, 96000, seed=300)[0, 0], waves[2].reshape(1, -1), waves[4]]
    outs = enhance_list(model, clips, max_batch=2, N=1, solver="euler")
    assert [o.shape for o in outs] == [c.shape for c in clips]
    assert all(torch.isfinite(o).all() for o in outs)


def test_config2_full_size_properties(model):
    """BASELINE.json config 2 at full size (32 clips x 2 s, midpoint N=3 = NFE 6), checked through
    size-independent properties: finiteness incl. an all-zero clip (normfac guard, util/other.py:77), per-clip
    independence (a clip enhanced alone is bit-identical to its row of the batch) and exact power-of-two scale
    equivariance of the normalise -> enhance -> de-normalise wrapper."""
    from flowdec_b200.util.synth import synth_waveforms
    B, L = 32, 96000
    y = synth_waveforms(B, L, seed=7)
    y[5] = 0.0
    g = torch.Generator().manual_seed(9)
    eps = torch.randn(B, 1, 768, 256, dtype=torch.complex64, generator=g)
    out = model.enhance(y, N=3, solver="midpoint", noise=eps)
    assert out.shape == y.shape and torch.isfinite(out).all()
    # (the all-zero clip is NOT silent after the untrained synthetic network: the CPU oracle gives max|x| = 460 for
    #  this clip and noise draw; what must hold is finiteness and independence from its neighbours)
    for i in (0, 5, 17, 31):
        one = model.enhance(y[i:i + 1], N=3, solver="midpoint", noise=eps[i:i + 1])
        assert torch.equal(one, out[i:i + 1]), f"clip {i} depends on its batch neighbours"
    half = model.enhance(0.5 * y, N=3, solver="midpoint", noise=eps)
    keep = [i for i in range(B) if i != 5]
    assert torch.equal(half[keep], 0.5 * out[keep])
    assert torch.equal(half[5], out[5])          # the all-zero clip is normalised by 1 either way (util/other.py:77)


def test_caches_are_bounded_over_many_lengths(model):
    """enhance.py's one-file-per-call loop over files of different lengths must not grow GPU memory without bound:
    clips of one padded-frame bucket share a cache entry, entries and backbone workspaces are evicted LRU"""
    from flowdec_b200.util.synth import synth_waveforms
    model.reset_cache()
    outs = {}
    for L in (24000, 24100, 30000, 50000, 50500, 75000, 99000, 125000, 24000):
        y = synth_waveforms(1, L, seed=L)
        g = torch.Generator().manual_seed(L)
        from flowdec_b200.util.other import padded_frames
        eps = torch.randn(1, 1, 768, padded_frames(1 + L // 384), dtype=torch.complex64, generator=g)
        x = model.enhance(y, N=1, solver="euler", noise=eps)
        assert x.shape == y.shape and torch.isfinite(x).all()
        if L in outs:                                    # evicted and rebuilt: same result
            assert torch.equal(outs[L], x)
        outs[L] = x
        assert len(model._graphs) <= model.graph_cache_size
        assert len(model.backbone._workspaces) <= max(model.backbone.max_workspaces, 2 * model.graph_cache_size)
    # 24000 and 24100 fall into the same 64-frame bucket -> one entry served both
    keys = {k[1] for k in model._graphs}
    assert len(keys) == len(model._graphs)


def test_per_sample_times(model):
    """NCSNpp.forward with one t per sample (reference ncsnpp.py:254): equal to evaluating every sample with its own
    scalar t, bit for bit"""
    g = torch.Generator().manual_seed(21)
    x = torch.randn(3, 1, 768, 64, dtype=torch.complex64, generator=g).cuda()
    y = torch.randn(3, 1, 768, 64, dtype=torch.complex64, generator=g).cuda()
    t = torch.tensor([0.3, 0.8, 0.3]).cuda()
    v = model.backbone(x, y, t)
    for i in range(3):
        vi = model.backbone(x[i:i + 1], y[i:i + 1], t[i:i + 1])
        assert torch.equal(v[i:i + 1], vi)
    with pytest.raises(ValueError):
        model.backbone(x, y, torch.tensor([0.1, 0.2]).cuda())
