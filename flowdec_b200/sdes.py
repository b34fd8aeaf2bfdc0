"""SDEs of the ScoreDec / SGMSE+ baseline path (reference: flowdec/sdes.py:132-206, OUVESDE).

Inference with a shared scalar time needs only the scalar coefficient schedules; tensors never
pass through this module — the per-step affine updates are fused into the backbone's last kernel
(fd_output_axpy).  Scalars are evaluated in float32 like the reference's torch expressions."""
import numpy as np


class OUVESDE:
    """Ornstein-Uhlenbeck variance-exploding SDE:  dx = theta (y - x) dt + sigma(t) dw,
    sigma(t) = sigma_min (sigma_max/sigma_min)^t sqrt(2 log(sigma_max/sigma_min))."""

    def __init__(self, theta, sigma_min, sigma_max, N=1000, **ignored_kwargs):
        self.theta, self.sigma_min, self.sigma_max, self.N = theta, sigma_min, sigma_max, N
        self.logsig = np.log(self.sigma_max / self.sigma_min)

    def copy(self):
        return OUVESDE(self.theta, self.sigma_min, self.sigma_max, N=self.N)

    @property
    def T(self):
        return 1

    def diffusion(self, t):
        """sdes.py:176-184"""
        f32 = np.float32
        sigma = f32(self.sigma_min) * f32(self.sigma_max / self.sigma_min) ** f32(t)
        return f32(sigma * np.sqrt(2 * self.logsig))

    def _std(self, t):
        """sdes.py:191-204 (scalar t)"""
        f32 = np.float32
        t = f32(t)
        smin, theta, logsig = self.sigma_min, self.theta, self.logsig
        num = f32(smin ** 2) * np.exp(f32(-2 * theta) * t) * (np.exp(f32(2 * (theta + logsig)) * t) - f32(1)) * f32(logsig)
        return f32(np.sqrt(f32(num) / f32(theta + logsig)))

    def discretize(self, t):
        """sdes.py:68-76: f = drift * dt (drift = theta (y - x)), G = diffusion * sqrt(dt), dt = 1/N.
        Returns (theta*dt, G) — the drift is applied as an affine combination of x and y."""
        dt = 1 / self.N
        G = np.float32(self.diffusion(t) * np.sqrt(np.float32(dt)))
        return np.float32(self.theta * dt), G
