"""Fixed-step ODE steppers of the FlowDec sampler, expressed as fused-stage schedules.

Reference: torchdyn==1.0.6 `Euler` / `Midpoint` (driven from flowdec/model.py:511-514) and
flowdec/sampling/solvers.py:15-68 (`Heun2`, `Heun2_EulerLast`, `get_solver`).

A stepper here does not touch tensors: it yields, for one step (t, dt), the list of
backbone evaluations and for each the affine combination that the last kernel of the
backbone (fd_output_axpy) writes:   dst = c1*base1 + c2*base2 + coef * v(t_eval, src).
"""
import numpy as np

SOLVERS = ("euler", "midpoint", "heun2", "heun2_eulerlast")


def get_solver(name, *args, **kwargs):
    """reference solvers.py:64-68 — returns the (validated) solver name."""
    if not isinstance(name, str) or name.lower() not in SOLVERS:
        raise ValueError(f"unknown solver {name!r}; flowdec_b200 implements {SOLVERS}")
    return name.lower()


def nfe_per_step(solver):
    return 1 if solver == "euler" else 2


def t_grid(N):
    """float32 replica of torch.linspace(0, 1, N+1) and of torchdyn's t/dt bookkeeping:
    t <- t + dt ; dt <- t_span[k+1] - t.  Returns [(t, dt)] per step as np.float32."""
    import torch
    ts = torch.linspace(0, 1, N + 1).numpy().astype(np.float32)
    out = []
    t = ts[0]
    dt = np.float32(ts[1] - ts[0])
    for step in range(1, N + 1):
        out.append((np.float32(t), np.float32(dt)))
        t = np.float32(t + dt)
        if step < N:
            dt = np.float32(ts[step + 1] - t)
    return out


def stages(solver, t, dt):
    """-> list of (t_eval, src, dst, base1, c1, base2, c2, coef) with symbolic buffer names
    'x' (state at step start), 'tmp' (intermediate), 'xn' (state at step end)."""
    f32 = np.float32
    if solver == "euler":
        return [(t, "x", "xn", "x", 1.0, None, 0.0, float(dt))]
    if solver == "midpoint":
        half = f32(0.5) * dt
        return [(t, "x", "tmp", "x", 1.0, None, 0.0, float(half)),
                (f32(t + half), "tmp", "xn", "x", 1.0, None, 0.0, float(dt))]
    if solver in ("heun2", "heun2_eulerlast"):
        if solver == "heun2_eulerlast" and np.isclose(f32(t + dt), f32(1.0), rtol=1e-5, atol=1e-8):
            return [(t, "x", "xn", "x", 1.0, None, 0.0, float(dt))]
        # x_pred = x + dt k1 ; x_sol = x + dt/2 (k1 + k2) = (x + x_pred)/2 + dt/2 k2
        return [(t, "x", "tmp", "x", 1.0, None, 0.0, float(dt)),
                (f32(t + dt), "tmp", "xn", "x", 0.5, "tmp", 0.5, float(dt * f32(0.5)))]
    raise ValueError(solver)
