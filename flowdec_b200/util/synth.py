"""Deterministic synthetic weights and waveforms (SURVEY.md §8d "Synthetic inputs").

Real checkpoints are not available offline, and a freshly initialised reference backbone is
numerically a skip-path-only network (every Conv_1 / pyramid conv is scaled by 1e-10:
reference layers.py:98-101, layerspp.py:243, ncsnpp.py:218,230).  Parity fixtures and the
benchmark therefore overwrite every learnable tensor with seeded, non-degenerate values.
Each tensor depends only on (seed, key, shape), never on iteration order, so the same
state_dict can be regenerated on any machine with the same torch build.
"""
import math
import zlib

import torch

_KEEP = ("sigma_x", "sigma_y", "feature_extractor.complex_stft.window")


def _gen(seed, key):
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63 - 1))
    return g


def synth_tensor(seed, key, like):
    """like: tensor giving shape/dtype. Returns the synthetic value for state_dict entry `key`."""
    shape = tuple(like.shape)
    g = _gen(seed, key)
    if key in _KEEP or key.endswith("sigma_y") or key.endswith("sigma_x") or key.endswith(".window"):
        return like.clone()
    if key.endswith(".W") and len(shape) == 1:  # GaussianFourierProjection, frozen, scale 16 (layerspp.py:47)
        return torch.randn(shape, generator=g) * 16.0
    if key.endswith(".b"):  # NIN bias of the attention block (layers.py NIN)
        return 0.05 * torch.randn(shape, generator=g)
    if len(shape) >= 2:  # conv / linear weight: U(+-sqrt(3/fan_in)) -> unit gain
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        lim = math.sqrt(3.0 / fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * lim
    if key.endswith(".weight"):  # GroupNorm gamma
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if key.endswith(".bias"):
        return 0.05 * torch.randn(shape, generator=g)
    return like.clone()


def synth_state_dict(template_sd, seed=0):
    """template_sd: mapping key -> tensor (shapes/dtypes). Returns a new dict with synthetic values."""
    out = {}
    for k, v in template_sd.items():
        out[k] = synth_tensor(seed, k, v).to(v.dtype)
    return out


def synth_waveforms(batch, length, seed=1234, kind="tones", sr=48000):
    """[batch, 1, length] float32 test clips (clip i depends only on seed+i).

    tones : sum of 8 sinusoids (log-uniform 50 Hz..20 kHz, 1/f-ish amplitudes, random phase)
            + white noise at -30 dB, peak-normalised to 0.5
    gauss : white Gaussian noise, sigma 0.1
    zeros : all-zero clip (exercises the normfac guard, reference util/other.py:77)
    """
    out = torch.zeros(batch, 1, length, dtype=torch.float32)
    t = torch.arange(length, dtype=torch.float64) / sr
    for i in range(batch):
        g = torch.Generator(device="cpu")
        g.manual_seed(seed + i)
        if kind == "zeros":
            continue
        if kind == "gauss":
            out[i, 0] = 0.1 * torch.randn(length, generator=g)
            continue
        u = torch.rand(8, generator=g, dtype=torch.float64)
        freqs = 50.0 * (20000.0 / 50.0) ** u
        phases = 2 * math.pi * torch.rand(8, generator=g, dtype=torch.float64)
        amps = (200.0 / freqs) ** 0.5
        x = (amps[:, None] * torch.sin(2 * math.pi * freqs[:, None] * t[None, :] + phases[:, None])).sum(0)
        x = x / x.abs().max()
        x = x + 10 ** (-30 / 20) * torch.randn(length, generator=g, dtype=torch.float64)
        x = 0.5 * x / x.abs().max()
        out[i, 0] = x.float()
    return out


def synth_dac_state_dict(latent_dim, decoder_dim, rates, n_codebooks, codebook_size=1024, codebook_dim=8,
                         seed=0, encoder_dim=None, encoder_rates=None):
    """descript-style decoder + quantizer state_dict with seeded non-degenerate values"""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(p, cout, cin, k, transpose=False):
        shape = (cin, cout, k) if transpose else (cout, cin, k)
        v = torch.randn(shape, generator=g)
        fan = cin * k if not transpose else cin * k / 2
        sd[p + ".weight_v"] = v
        sd[p + ".weight_g"] = (torch.rand(shape[0], 1, 1, generator=g) + 0.5) * (
            v.norm(dim=(1, 2), keepdim=True) / math.sqrt(fan) * (math.sqrt(cin / cout) if transpose else 1.0))
        sd[p + ".bias"] = 0.05 * torch.randn(cout, generator=g)

    def alpha(p, c):
        sd[p] = 0.5 + torch.rand(1, c, 1, generator=g)

    for i in range(n_codebooks):
        q = f"quantizer.quantizers.{i}."
        sd[q + "codebook.weight"] = torch.randn(codebook_size, codebook_dim, generator=g)
        conv(q + "out_proj", latent_dim, codebook_dim, 1)
        conv(q + "in_proj", codebook_dim, latent_dim, 1)
    m = "decoder.model."
    conv(m + "0", decoder_dim, latent_dim, 7)
    ch = decoder_dim
    for i, s in enumerate(rates):
        b = f"{m}{i + 1}.block."
        alpha(b + "0.alpha", ch)
        conv(b + "1", ch // 2, ch, 2 * s, transpose=True)
        ch //= 2
        for j in range(3):
            r = f"{b}{j + 2}.block."
            alpha(r + "0.alpha", ch)
            conv(r + "1", ch, ch, 7)
            alpha(r + "2.alpha", ch)
            conv(r + "3", ch, ch, 1)
    n = len(rates)
    alpha(f"{m}{n + 1}.alpha", ch)
    conv(f"{m}{n + 2}", 1, ch, 7)
    if encoder_dim is not None:          # drawn last so the decoder / quantizer values above do not change
        e = "encoder.block."
        conv(e + "0", encoder_dim, 1, 7)
        ch = encoder_dim
        for i, st in enumerate(encoder_rates):
            b = f"{e}{i + 1}.block."
            for j in range(3):
                r = f"{b}{j}.block."
                alpha(r + "0.alpha", ch)
                conv(r + "1", ch, ch, 7)
                alpha(r + "2.alpha", ch)
                conv(r + "3", ch, ch, 1)
            alpha(b + "3.alpha", ch)
            conv(b + "4", 2 * ch, ch, 2 * st)
            ch *= 2
        n = len(encoder_rates)
        alpha(f"{e}{n + 1}.alpha", ch)
        conv(f"{e}{n + 2}", latent_dim, ch, 3)
        sd[f"{e}{n + 2}.weight_g"] *= 0.05        # residual stacks grow the activations; keep z = O(1)
    return sd
