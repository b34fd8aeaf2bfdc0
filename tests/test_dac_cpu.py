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
