"""TEST INFRASTRUCTURE — CPU oracle for the upstream NDAC (DAC) codec (SURVEY.md §8 a11 decode, §8f-1 encode).

The codec is NOT in the reference tree: it is the un-vendored dependency
`descript-audio-codec==1.0.0` (requirements.txt:4; call sites demo.ipynb:56,101-105:
`dac.quantizer.from_codes(codes)` -> `dac.decode(zq)`).  This file restates its published
algorithm (dac/model/dac.py `Decoder`, `DecoderBlock`, `ResidualUnit`; dac/nn/layers.py
`Snake1d`, `WNConv1d`, `WNConvTranspose1d`; dac/nn/quantize.py `ResidualVectorQuantize.from_codes`;
and for the encode half (demo.ipynb:101-102) dac/model/dac.py `Encoder`, `EncoderBlock`, `DAC.preprocess`,
`DAC.encode`, dac/nn/quantize.py `VectorQuantize.forward/decode_latents`, `ResidualVectorQuantize.forward`)
on a descript-style state_dict (weight-normalised convs stored as weight_g / weight_v).

Pinning: descript-audio-codec itself is not installed, so this oracle is pinned against the
architecturally identical port in `transformers.models.dac` (tests/test_dac_cpu.py maps the
weights and compares); vs descript's own code it is "parity unpinned" (no copy available offline).
"""
import math

import torch
import torch.nn.functional as F


def wn(sd, p):
    """torch.nn.utils.weight_norm (dim=0): w = g * v / ||v||, norm over all dims but the first."""
    v, g = sd[p + ".weight_v"], sd[p + ".weight_g"]
    return g * v / v.norm(dim=tuple(range(1, v.ndim)), keepdim=True)


def snake(x, alpha):
    """dac/nn/layers.py snake(): x + (alpha + 1e-9)^-1 * sin(alpha x)^2, alpha [1,C,1]"""
    return x + (alpha + 1e-9).reciprocal() * torch.sin(alpha * x).pow(2)


def from_codes(sd, codes, prefix="quantizer."):
    """ResidualVectorQuantize.from_codes: z_q = sum_i out_proj_i(codebook_i[codes[:, i]]^T).
    codes int64 [B, n_q, T] -> z_q [B, D, T]"""
    z = 0.0
    for i in range(codes.shape[1]):
        q = f"{prefix}quantizers.{i}."
        e = F.embedding(codes[:, i, :], sd[q + "codebook.weight"]).transpose(1, 2)      # [B, 8, T]
        z = z + F.conv1d(e, wn(sd, q + "out_proj"), sd[q + "out_proj.bias"])
    return z


def residual_unit(sd, p, x, dilation):
    """ResidualUnit: Snake -> WNConv1d(k7, dilation, pad 3*dilation) -> Snake -> WNConv1d(k1); + x"""
    y = snake(x, sd[p + ".block.0.alpha"])
    y = F.conv1d(y, wn(sd, p + ".block.1"), sd[p + ".block.1.bias"], dilation=dilation, padding=3 * dilation)
    y = snake(y, sd[p + ".block.2.alpha"])
    y = F.conv1d(y, wn(sd, p + ".block.3"), sd[p + ".block.3.bias"])
    return x + y


def decode(sd, z, rates, prefix="decoder."):
    """Decoder: WNConv1d(k7) -> [Snake -> WNConvTranspose1d(k=2s, stride s, pad ceil(s/2)) ->
    ResidualUnit x3 (dilation 1, 3, 9)] per rate -> Snake -> WNConv1d(k7 -> 1) -> tanh"""
    m = prefix + "model."
    x = F.conv1d(z, wn(sd, m + "0"), sd[m + "0.bias"], padding=3)
    for i, s in enumerate(rates):
        b = f"{m}{i + 1}.block."
        x = snake(x, sd[b + "0.alpha"])
        x = F.conv_transpose1d(x, wn(sd, b + "1"), sd[b + "1.bias"], stride=s, padding=math.ceil(s / 2))
        for j, d in enumerate((1, 3, 9)):
            x = residual_unit(sd, f"{b}{j + 2}", x, d)
    n = len(rates)
    x = snake(x, sd[f"{m}{n + 1}.alpha"])
    x = F.conv1d(x, wn(sd, f"{m}{n + 2}"), sd[f"{m}{n + 2}.bias"], padding=3)
    return torch.tanh(x)


def preprocess(x, hop_length):
    """DAC.preprocess: zero-pad on the right to a multiple of hop_length = prod(encoder_rates)"""
    L = x.shape[-1]
    return F.pad(x, (0, math.ceil(L / hop_length) * hop_length - L))


def encode(sd, x, rates, prefix="encoder."):
    """Encoder: WNConv1d(1->d, k7) -> [ResidualUnit x3 (dilation 1, 3, 9) -> Snake ->
    WNConv1d(k=2s, stride s, pad ceil(s/2), ch -> 2ch)] per rate -> Snake -> WNConv1d(k3 -> latent).
    x [B, 1, L] -> z [B, D, L / prod(rates)]"""
    m = prefix + "block."
    x = F.conv1d(x, wn(sd, m + "0"), sd[m + "0.bias"], padding=3)
    for i, s in enumerate(rates):
        b = f"{m}{i + 1}.block."
        for j, d in enumerate((1, 3, 9)):
            x = residual_unit(sd, f"{b}{j}", x, d)
        x = snake(x, sd[b + "3.alpha"])
        x = F.conv1d(x, wn(sd, b + "4"), sd[b + "4.bias"], stride=s, padding=math.ceil(s / 2))
    n = len(rates)
    x = snake(x, sd[f"{m}{n + 1}.alpha"])
    return F.conv1d(x, wn(sd, f"{m}{n + 2}"), sd[f"{m}{n + 2}.bias"], padding=1)


def vq_lookup(e, codebook):
    """VectorQuantize.decode_latents: nearest code by euclidean distance between the L2-normalised
    latent and the L2-normalised codebook; e [N, d] -> (indices [N], margin [N] = best - second-best of -dist)"""
    en, cn = F.normalize(e), F.normalize(codebook)
    dist = en.pow(2).sum(1, keepdim=True) - 2 * en @ cn.t() + cn.pow(2).sum(1, keepdim=True).t()
    top = (-dist).topk(2, dim=1)
    return (-dist).max(1)[1], top.values[:, 0] - top.values[:, 1]


def rvq_encode(sd, z, n_quantizers=None, prefix="quantizer."):
    """ResidualVectorQuantize.forward in eval mode: per quantizer i < n_quantizers on the running residual
    z_e = in_proj_i(residual); idx = lookup(z_e); z_q_i = out_proj_i(codebook_i[idx]); residual -= z_q_i.
    Returns (z_q [B,D,T], codes [B,nq,T], latents [B,nq*d,T], commitment_loss, codebook_loss, margins [B,nq,T])"""
    nq_total = len({k.split(".")[2] for k in sd if k.startswith(prefix + "quantizers.")})
    nq = nq_total if n_quantizers is None else min(int(n_quantizers), nq_total)
    B, D, T = z.shape
    zq, res = torch.zeros_like(z), z
    codes, lat, margins, commit = [], [], [], z.new_zeros(())
    for i in range(nq):
        q = f"{prefix}quantizers.{i}."
        ze = F.conv1d(res, wn(sd, q + "in_proj"), sd[q + "in_proj.bias"])                # [B, d, T]
        idx, mg = vq_lookup(ze.transpose(1, 2).reshape(B * T, -1), sd[q + "codebook.weight"])
        idx, mg = idx.reshape(B, T), mg.reshape(B, T)
        zc = F.embedding(idx, sd[q + "codebook.weight"]).transpose(1, 2)
        commit = commit + F.mse_loss(ze, zc, reduction="none").mean([1, 2]).mean()
        zqi = F.conv1d(zc, wn(sd, q + "out_proj"), sd[q + "out_proj.bias"])
        zq, res = zq + zqi, res - zqi
        codes.append(idx); lat.append(ze); margins.append(mg)
    return zq, torch.stack(codes, 1), torch.cat(lat, 1), commit, commit.clone(), torch.stack(margins, 1)


from flowdec_b200.util.synth import synth_dac_state_dict  # noqa: E402,F401  (seeded synthetic NDAC weights)
