"""CPU: pins oracle/dac_oracle.py (restatement of descript-audio-codec's decoder + from_codes)
against the architecturally identical port shipped in `transformers.models.dac`."""
import pytest
import torch

from oracle import dac_oracle as D

tdac = pytest.importorskip("transformers.models.dac.modeling_dac")


def _hf_modules(sd, latent, dim, rates, nq):
    from transformers.models.dac.configuration_dac import DacConfig
    cfg = DacConfig(hidden_size=latent, decoder_hidden_size=dim, upsampling_ratios=list(rates), n_codebooks=nq,
                    codebook_size=1024, codebook_dim=8)
    dec = tdac.DacDecoder(cfg).eval()
    rvq = tdac.DacResidualVectorQuantizer(cfg).eval()
    with torch.no_grad():
        m = "decoder.model."
        dec.conv1.weight.copy_(D.wn(sd, m + "0")); dec.conv1.bias.copy_(sd[m + "0.bias"])
        for i in range(len(rates)):
            b = f"{m}{i + 1}.block."
            blk = dec.block[i]
            blk.snake1.alpha.copy_(sd[b + "0.alpha"])
            blk.conv_t1.weight.copy_(D.wn(sd, b + "1")); blk.conv_t1.bias.copy_(sd[b + "1.bias"])
            for j, ru in enumerate((blk.res_unit1, blk.res_unit2, blk.res_unit3)):
                r = f"{b}{j + 2}.block."
                ru.snake1.alpha.copy_(sd[r + "0.alpha"])
                ru.conv1.weight.copy_(D.wn(sd, r + "1")); ru.conv1.bias.copy_(sd[r + "1.bias"])
                ru.snake2.alpha.copy_(sd[r + "2.alpha"])
                ru.conv2.weight.copy_(D.wn(sd, r + "3")); ru.conv2.bias.copy_(sd[r + "3.bias"])
        n = len(rates)
        dec.snake1.alpha.copy_(sd[f"{m}{n + 1}.alpha"])
        dec.conv2.weight.copy_(D.wn(sd, f"{m}{n + 2}")); dec.conv2.bias.copy_(sd[f"{m}{n + 2}.bias"])
        for i in range(nq):
            q = f"quantizer.quantizers.{i}."
            rvq.quantizers[i].codebook.weight.copy_(sd[q + "codebook.weight"])
            rvq.quantizers[i].out_proj.weight.copy_(D.wn(sd, q + "out_proj"))
            rvq.quantizers[i].out_proj.bias.copy_(sd[q + "out_proj.bias"])
    return dec, rvq


def test_dac_oracle_vs_transformers_port():
    latent, dim, rates, nq = 64, 96, (4, 3, 2), 5
    sd = D.synth_dac_state_dict(latent, dim, rates, nq, seed=1)
    dec, rvq = _hf_modules(sd, latent, dim, rates, nq)
    g = torch.Generator().manual_seed(2)
    codes = torch.randint(0, 1024, (2, nq, 21), generator=g)
    with torch.no_grad():
        z_ref = rvq.from_codes(codes)[0]
        z = D.from_codes(sd, codes)
        assert torch.allclose(z, z_ref, rtol=1e-5, atol=1e-5)
        x_ref = dec(z_ref)
        x = D.decode(sd, z, rates)
    assert x.shape == x_ref.shape
    assert torch.allclose(x, x_ref, rtol=1e-4, atol=1e-5), (x - x_ref).abs().max()


def _hf_encoder(sd, latent, edim, erates, nq):
    from transformers.models.dac.configuration_dac import DacConfig
    cfg = DacConfig(encoder_hidden_size=edim, downsampling_ratios=list(erates), hidden_size=latent, n_codebooks=nq,
                    codebook_size=1024, codebook_dim=8)
    enc = tdac.DacEncoder(cfg).eval()
    rvq = tdac.DacResidualVectorQuantizer(cfg).eval()

    def cp(mod, p):
        mod.weight.copy_(D.wn(sd, p)); mod.bias.copy_(sd[p + ".bias"])

    with torch.no_grad():
        e = "encoder.block."
        cp(enc.conv1, e + "0")
        for i in range(len(erates)):
            b = f"{e}{i + 1}.block."
            blk = enc.block[i]
            for j, ru in enumerate((blk.res_unit1, blk.res_unit2, blk.res_unit3)):
                r = f"{b}{j}.block."
                ru.snake1.alpha.copy_(sd[r + "0.alpha"]); cp(ru.conv1, r + "1")
                ru.snake2.alpha.copy_(sd[r + "2.alpha"]); cp(ru.conv2, r + "3")
            blk.snake1.alpha.copy_(sd[b + "3.alpha"]); cp(blk.conv1, b + "4")
        n = len(erates)
        enc.snake1.alpha.copy_(sd[f"{e}{n + 1}.alpha"]); cp(enc.conv2, f"{e}{n + 2}")
        for i in range(nq):
            q = f"quantizer.quantizers.{i}."
            rvq.quantizers[i].codebook.weight.copy_(sd[q + "codebook.weight"])
            cp(rvq.quantizers[i].in_proj, q + "in_proj")
            cp(rvq.quantizers[i].out_proj, q + "out_proj")
    return enc, rvq


def test_dac_encode_oracle_vs_transformers_port():
    """§8f-1: Encoder + ResidualVectorQuantize.forward (eval) restatement vs the transformers port"""
    latent, dim, rates, nq, edim, erates = 64, 96, (4, 3, 2), 5, 8, (2, 3, 4)
    sd = D.synth_dac_state_dict(latent, dim, rates, nq, seed=1, encoder_dim=edim, encoder_rates=erates)
    enc, rvq = _hf_encoder(sd, latent, edim, erates, nq)
    g = torch.Generator().manual_seed(3)
    x = D.preprocess(0.3 * torch.randn(2, 1, 1000, generator=g), 24)
    assert x.shape[-1] == 1008
    with torch.no_grad():
        z_ref = enc(x)
        z = D.encode(sd, x, erates)
        assert z.shape == z_ref.shape == (2, latent, 42)
        assert torch.allclose(z, z_ref, rtol=1e-4, atol=1e-5), (z - z_ref).abs().max()
        for n_q in (None, 3):
            zq_ref, codes_ref, lat_ref, commit_ref, cb_ref = rvq(z_ref, n_q)
            zq, codes, lat, commit, cbl, margin = D.rvq_encode(sd, z_ref, n_q)
            assert codes.shape == codes_ref.shape and torch.equal(codes, codes_ref)
            assert torch.allclose(zq, zq_ref, rtol=1e-4, atol=1e-5)
            assert torch.allclose(lat, lat_ref, rtol=1e-4, atol=1e-5)
            assert torch.allclose(commit, commit_ref.mean(), rtol=1e-4) and torch.allclose(cbl, cb_ref.mean(), rtol=1e-4)
            assert (margin >= 0).all()


def test_op_lists_follow_module_to():
    """nn.Module._apply replaces buffers: the decoder / encoder op lists must resolve tensors at call time
    (round-1 bug: they kept the construction-time CPU tensors and handed host pointers to the kernels)"""
    from flowdec_b200.ndac import DAC
    from flowdec_b200 import _lib
    sd = D.synth_dac_state_dict(64, 96, (4, 3, 2), 5, seed=1, encoder_dim=8, encoder_rates=(2, 3, 4))
    m = DAC(sd, decoder_dim=96, decoder_rates=(4, 3, 2), n_codebooks=5, latent_dim=64, sample_rate=48000,
            encoder_dim=8, encoder_rates=(2, 3, 4))
    assert m._enc_ops is not None
    before = m.op_tensors()
    assert before and all(t.dtype == torch.float32 for t in before.values())
    m = m.to(torch.float64)
    after = m.op_tensors()
    assert set(after) == set(before)
    for n, t in after.items():
        assert t.dtype == torch.float64 and t is getattr(m, n), n
    # and a host tensor can never be turned into a kernel argument
    with pytest.raises(_lib.FlowDecNativeError):
        _lib.ptr(torch.zeros(4))
