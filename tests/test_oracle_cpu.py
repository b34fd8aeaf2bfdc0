"""CPU tests (no GPU): the oracle against the golden vectors produced by the unmodified
reference, and — where /root/reference is present — against the reference itself."""
import os

import numpy as np
import pytest
import torch

from flowdec_b200.util.synth import synth_state_dict
from oracle import flowdec_oracle as O
from oracle import ref_shim
from oracle.make_golden import golden_inputs

GOLD = os.path.join(os.path.dirname(__file__), "golden", "flowdec_75m_seed0.npz")


@pytest.fixture(scope="module")
def sd():
    from flowdec_b200.model import build_flowdec
    m = build_flowdec("75m")
    return synth_state_dict(m.state_dict(), seed=0)


@pytest.fixture(scope="module")
def gold():
    return {k: torch.from_numpy(v) for k, v in np.load(GOLD).items()}


def rel_l2(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_feature_path_vs_golden(sd, gold):
    I = golden_inputs()
    w = sd["feature_extractor.complex_stft.window"]
    Y, info = O.preprocess(I["y"], w)
    assert rel_l2(torch.view_as_real(Y), gold["preprocess_Y"]) < 1e-6
    assert torch.equal(info["normfac"], gold["normfac"])
    assert rel_l2(O.postprocess(Y, info, w), gold["postprocess_of_Y"]) < 1e-6


def test_fir_vs_golden(gold):
    xf = torch.randn(2, 8, 12, 16, generator=torch.Generator().manual_seed(5))
    assert rel_l2(O.fir_down2(xf), gold["fir_down"]) < 1e-6
    assert rel_l2(O.fir_up2(xf), gold["fir_up"]) < 1e-6


def test_backbone_vs_golden(sd, gold):
    I = golden_inputs()
    with torch.no_grad():
        v = O.ncsnpp_forward(sd, I["X"], I["Y"], I["t"])
    assert rel_l2(torch.view_as_real(v), gold["backbone_v"]) < 1e-5


@pytest.mark.parametrize("N,solver", [(1, "euler"), (1, "midpoint")])
def test_enhance_vs_golden(sd, gold, N, solver):
    I = golden_inputs()
    with torch.no_grad():
        x = O.enhance(sd, I["y"], N=N, solver=solver, eps=I["eps"])
    assert rel_l2(x, gold[f"enhance_{solver}_N{N}"]) < 1e-4


def test_enhance_headline_nfe6_vs_golden(sd):
    """the headline solver setting (midpoint N=3 = NFE 6) on the 0.5 s clip: oracle vs the reference's own output
    (tests/golden/flowdec_75m_headline.npz, oracle/make_golden.py --headline); also ties the metric helpers"""
    from oracle import metrics as M
    H = np.load(os.path.join(os.path.dirname(__file__), "golden", "flowdec_75m_headline.npz"))
    I = golden_inputs()
    with torch.no_grad():
        x = O.enhance(sd, I["y"], N=3, solver="midpoint", eps=I["eps"])
    g = torch.from_numpy(H["enhance_midpoint_N3"])
    assert rel_l2(x, g) < 1e-3
    assert M.snr_db(x, g) > 60 and M.si_sdr_db(x, g) > 60 and M.logspec_mse(x, g) < 1e-3
    assert abs(M.snr_db(0.9 * g, g) - 20.0) < 1e-4 and M.si_sdr_db(0.9 * g, g) > 100


def test_scoredec_pc_sampler_vs_golden(sd, gold):
    I = golden_inputs()
    with torch.no_grad():
        x = O.score_enhance(sd, I["y"], N=2, snr=0.5, noise=I["score_draws"])
    assert rel_l2(x, gold["scoredec_pc_N2"]) < 1e-4


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present on this machine")
def test_oracle_vs_live_reference_small():
    """cheap live pin: a 2-level, nf=16 backbone of the same family, reference vs oracle"""
    R = ref_shim.load_reference()
    kw = dict(ref_shim.BACKBONE_KW)
    kw.update(image_size=64, nf=16, ch_mult=[2, 2])
    net = R.ncsnpp.NCSNpp(**kw)
    sd = synth_state_dict({"backbone." + k: v for k, v in net.state_dict().items()}, seed=3)
    net.load_state_dict({k[len("backbone."):]: v for k, v in sd.items()})
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 1, 64, 16, dtype=torch.complex64, generator=g)
    y = torch.randn(2, 1, 64, 16, dtype=torch.complex64, generator=g)
    t = torch.tensor([0.7])
    with torch.no_grad():
        ref = net(x, y, t)
        mine = O.ncsnpp_forward(sd, x, y, t, num_resolutions=2)
    assert rel_l2(torch.view_as_real(mine), torch.view_as_real(ref)) < 1e-5


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present on this machine")
def test_oracle_attention_backbone_vs_live_reference():
    """SURVEY.md §8f-3 (next row): the SGMSE-style baseline backbone — config/model/backbone/
    ncsnpp_default_ycond.yaml: 7 levels, 2 res-blocks per level, bottleneck attention, 3x3 output layer — at a
    reduced width, reference vs oracle"""
    R = ref_shim.load_reference()
    kw = dict(ref_shim.BACKBONE_KW)
    kw.update(image_size=128, nf=16, ch_mult=[1, 1, 2, 2, 2, 2, 2], num_res_blocks=2, bottleneck_attn=True,
              output_layer_kwargs=dict(kernel_size=3, bias=False, padding="same", padding_mode="zeros"))
    net = R.ncsnpp.NCSNpp(**kw)
    sd = synth_state_dict({"backbone." + k: v for k, v in net.state_dict().items()}, seed=4)
    net.load_state_dict({k[len("backbone."):]: v for k, v in sd.items()})
    assert any(".NIN_0.W" in k for k in sd)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 1, 128, 192, dtype=torch.complex64, generator=g)   # bottleneck: 2 x 3 tokens
    y = torch.randn(2, 1, 128, 192, dtype=torch.complex64, generator=g)
    t = torch.tensor([0.3])
    with torch.no_grad():
        ref = net(x, y, t)
        mine = O.ncsnpp_forward(sd, x, y, t, num_resolutions=7, num_res_blocks=2, bottleneck_attn=True)
    assert rel_l2(torch.view_as_real(mine), torch.view_as_real(ref)) < 1e-5


def test_oracle_attention_backbone_vs_golden():
    """same pin without the reference tree: tests/golden/ncsnpp_attn_seed4.npz was produced by the reference itself
    (oracle/make_golden.py:attn_backbone_golden); the state_dict is rebuilt from the stored key names / shapes"""
    from oracle.make_golden import attn_golden_inputs
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ncsnpp_attn_seed4.npz"))
    tmpl = {str(k): torch.empty([int(d) for d in str(s).split(",") if d], dtype=torch.float32)
            for k, s in zip(G["keys"], G["shapes"])}
    sd = synth_state_dict(tmpl, seed=4)
    I = attn_golden_inputs()
    with torch.no_grad():
        v = O.ncsnpp_forward(sd, I["X"], I["Y"], I["t"], num_resolutions=7, num_res_blocks=2, bottleneck_attn=True)
    assert rel_l2(torch.view_as_real(v), torch.from_numpy(G["backbone_v"])) < 1e-5
