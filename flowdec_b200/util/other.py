"""Waveform / spectrogram helpers with the reference's signatures
(/root/reference/flowdec/util/other.py:25-82).  FlowModel.enhance does not call these — the
same arithmetic is fused into the STFT kernels — they exist for API parity and are plain
tensor bookkeeping."""
import torch


def pad_spec(Y, mode="zero"):
    """Right-pad the time axis to a multiple of 64; returns (padded, undo_fn)."""
    if mode != "zero":
        raise NotImplementedError("FlowDec uses pad_spec(mode='zero') (reference model.py:152)")
    T = Y.size(-1)
    num_pad = (64 - T % 64) % 64
    return torch.nn.functional.pad(Y, (0, num_pad)), lambda Y_: Y_[..., :T]


def padded_frames(T):
    return T + (64 - T % 64) % 64


def normalize_noisy(y, mode, x=None):
    if mode == "noisy":
        normfac = y.abs().amax(dim=tuple(range(1, y.ndim)), keepdim=True)
    elif mode == "none":
        normfac = torch.ones((), device=y.device, dtype=y.dtype)
    else:
        raise ValueError(f"Unknown normalize mode: {mode}!")
    normfac = torch.where(torch.isclose(normfac, torch.zeros_like(normfac)), torch.ones_like(normfac), normfac)
    return y / normfac, (x / normfac if x is not None else None), normfac
