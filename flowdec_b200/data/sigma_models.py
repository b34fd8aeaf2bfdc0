"""Frequency-dependent sigma_y curves (reference: flowdec/data/sigma_models/__init__.py:21-47).

`from_file` keeps the reference signature.  The two curves FlowDec ships
(`data/flowdec_autoparams_{75m,25s}.npy`, 768 x f64 model parameters) are packaged in
`sigma_y_curves.npz`; a filename whose stem matches one of them resolves to the packaged copy
when the path itself does not exist, so the reference YAMLs work unchanged.
"""
import os
from typing import Optional

import numpy as np
import torch
from scipy.ndimage import gaussian_filter

_PACKAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sigma_y_curves.npz")


def _load_curve(filename):
    cands = [filename]
    if not os.path.isabs(filename):
        cands.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), filename))
    for c in cands:
        if os.path.exists(c):
            return np.load(c)
    stem = os.path.splitext(os.path.basename(filename))[0]
    with np.load(_PACKAGED) as z:
        if stem in z.files:
            return z[stem]
    raise FileNotFoundError(filename)


def from_file(filename: str, factor: float = 1.0, kernel_bandwidth: Optional[float] = None):
    """Load a 1-D sigma_y(f) curve, optionally Gaussian-smooth it (mode='nearest'), and return
    `factor * curve` as a float64 tensor of shape [F, 1] (broadcasts along time)."""
    curve = _load_curve(filename)
    if kernel_bandwidth is not None:
        curve = gaussian_filter(curve, sigma=kernel_bandwidth, mode="nearest")
    return factor * torch.from_numpy(np.asarray(curve)).unsqueeze(-1)


__all__ = ["from_file"]
