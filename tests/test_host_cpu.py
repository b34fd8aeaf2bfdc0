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
