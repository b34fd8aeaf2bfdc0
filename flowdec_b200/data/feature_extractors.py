"""Invertible feature extractors of FlowDec on B200.

API mirror of /root/reference/flowdec/data/feature_extractors.py (classes, constructor
keywords, `forward(x)` / `invert(X, orig_length=...)`), backed by the fused sm_100a kernels
fd_stft1534_compress / fd_istft1534_decompress (csrc/fd_stft.cu).  The window is an
nn.Parameter with the reference's state_dict key `complex_stft.window`.
"""
import abc
import math

import torch
from torch import nn

from .. import ops


class InvertibleFeatureExtractor(nn.Module, abc.ABC):
    """one-to-one mapping: extractor.invert(extractor(x)) == x up to numerical error"""

    @abc.abstractmethod
    def invert(self, x, **kwargs):
        pass


class ComplexSTFT(nn.Module):
    """parameter/config holder for reference feature_extractors.py:62-109"""

    def __init__(self, window_fn, n_fft, sampling_rate, hop_length=None, n_hops=None, learnable_window=False):
        super().__init__()
        assert (hop_length is not None) ^ (n_hops is not None), \
            "Exactly one of {hop_length, n_hops} must be specified!"
        if hop_length is None:
            hop_length = int(math.ceil(n_fft / n_hops))
        self.window = nn.Parameter(getattr(torch.signal.windows, window_fn)(n_fft), requires_grad=learnable_window)
        self.n_fft, self.hop_length, self.sampling_rate, self.center = n_fft, hop_length, sampling_rate, True


class CompressAmplitudesAndScale(nn.Module):
    """config holder for reference feature_extractors.py:112-139"""

    def __init__(self, compression_exponent: float, scale_factor: float):
        super().__init__()
        self.compression_exponent, self.scale_factor = compression_exponent, scale_factor


class AmplitudeCompressedComplexSTFT(InvertibleFeatureExtractor):
    """STFT (n_fft 1534, hop 384) + |X|^alpha e^{j angle X} * beta, fused in one kernel each way."""

    def __init__(self, window_fn, n_fft, sampling_rate, alpha, beta, hop_length=None, n_hops=None,
                 learnable_window=False, *args, **kwargs):
        super().__init__()
        self.complex_stft = ComplexSTFT(window_fn, n_fft, sampling_rate, hop_length=hop_length,
                                        n_hops=n_hops, learnable_window=learnable_window)
        self.compress = CompressAmplitudesAndScale(compression_exponent=alpha, scale_factor=beta)
        if n_fft != 1534 or self.complex_stft.hop_length != 384:
            raise NotImplementedError("the sm_100a STFT kernels are specialised for n_fft=1534, hop=384 "
                                      "(config/model/feature_extractor/compressed_complex_stft_final.yaml)")
        self._tw = None

    @property
    def alpha(self):
        return float(self.compress.compression_exponent)

    @property
    def beta(self):
        return float(self.compress.scale_factor)

    def twiddles(self):
        dev = self.complex_stft.window.device
        if self._tw is None or self._tw.device != dev:
            self._tw = ops.twiddles1534(dev)
        return self._tw

    @staticmethod
    def num_frames(L):
        return 1 + L // 384

    # fused entry points used by FlowModel.enhance ------------------------------------------
    def stft_compress(self, y2d, normfac, out, lengths=None):
        """y2d fp32 [B,L] (un-normalised), normfac fp32 [B] -> out fp32 [B,768,Tp,2]; frames beyond
        1+L//384 are zero (= pad_spec 'zero').  `lengths` (int32 [B]): ragged batch, per-clip L."""
        return ops.stft_compress(y2d, normfac, self.complex_stft.window, self.twiddles(), self.alpha, self.beta, out,
                                 lengths=lengths)

    def istft_decompress(self, X, L, normfac, out, lengths=None, ws=None):
        return ops.istft_decompress(X, L, self.complex_stft.window, self.twiddles(), normfac, self.alpha,
                                    self.beta, out, lengths=lengths, ws=ws)

    # reference-shaped API ---------------------------------------------------------------------
    def forward(self, x, comp_eps=None, **kwargs):
        """x: [B,C,T] or [B,T] waveform -> complex64 [B,C,768,frames]"""
        if comp_eps is not None:
            raise NotImplementedError("comp_eps is a training-time option (reference feature_extractors.py:124)")
        shp = x.shape
        y2d = x.reshape(-1, shp[-1]).float().contiguous()
        B, L = y2d.shape
        T = self.num_frames(L)
        ones = torch.ones(B, device=x.device, dtype=torch.float32)
        out = torch.empty(B, 768, T, 2, device=x.device, dtype=torch.float32)
        self.stft_compress(y2d, ones, out)
        return torch.view_as_complex(out).reshape(*shp[:-1], 768, T)

    def invert(self, X, orig_length=None, **kwargs):
        shp = X.shape
        T = shp[-1]
        L = orig_length if orig_length is not None else 384 * (T - 1)
        Xr = torch.view_as_real(X.to(torch.complex64).reshape(-1, 768, T).contiguous())
        B = Xr.shape[0]
        frames = self.num_frames(L)
        if frames > T:
            raise ValueError(f"orig_length={L} needs {frames} frames, spectrogram has {T}")
        ones = torch.ones(B, device=X.device, dtype=torch.float32)
        out = torch.empty(B, L, device=X.device, dtype=torch.float32)
        self.istft_decompress(Xr, L, ones, out)
        return out.reshape(*shp[:-2], L)


class NoOp(InvertibleFeatureExtractor):
    def forward(self, x, **kwargs):
        return x

    def invert(self, x, **kwargs):
        return x
