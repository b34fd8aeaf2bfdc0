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
