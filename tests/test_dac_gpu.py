"""GPU parity of the NDAC decode path (fd_dac.cu) against oracle/dac_oracle.py (itself pinned to
the transformers port of descript-audio-codec).  fp32 kernels: relative L2 <= 1e-4."""
import pytest
import torch

from flowdec_b200.ndac import DAC
from oracle import dac_oracle as D

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("latent,dim,rates,nq,T", [(64, 96, (4, 3, 2), 5, 37), (128, 256, (8, 5, 4, 4), 10, 19)])
def test_from_codes_and_decode(latent, dim, rates, nq, T):
    sd = D.synth_dac_state_dict(latent, dim, rates, nq, seed=1)
    model = DAC(sd, decoder_dim=dim, decoder_rates=rates, n_codebooks=nq, latent_dim=latent,
                sample_rate=48000).to("cuda").eval()
    g = torch.Generator().manual_seed(2)
    codes = torch.randint(0, 1024, (2, nq, T), generator=g)
    with torch.no_grad():
        z_ref = D.from_codes(sd, codes)
        x_ref = D.decode(sd, z_ref, rates)
    zq, _, c = model.quantizer.from_codes(codes)
    assert ((zq.cpu() - z_ref).norm() / z_ref.norm()).item() < 1e-5
    x = model.decode(zq)
    assert x.shape == x_ref.shape
    rel = ((x.cpu() - x_ref).norm() / x_ref.norm()).item()
    assert rel < 1e-4, rel
    # fewer codebooks than the model has (bitrate scalability, demo.ipynb:85-88)
    zq2, _, _ = model.quantizer.from_codes(codes[:, :3])
    assert ((zq2.cpu() - D.from_codes(sd, codes[:, :3])).norm() / z_ref.norm()).item() < 1e-5
