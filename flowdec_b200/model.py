"""FlowDec enhancement model API on B200.

Mirror of /root/reference/flowdec/model.py `EnhancementModel` / `FlowModel` for the inference
half: constructor keywords, `enhance()` signature and return conventions,
`forward(xt, y, t)`, `_preprocess/_postprocess`, `.device`, `.sampling_rate`, and the 265
state_dict keys.  No lightning / hydra / torchdyn / torchcfm dependency: the ODE loop
(torchdyn.NeuralODE.trajectory in the reference, model.py:511-514) is a host-side schedule of
fused kernel launches (flowdec_b200/sampling/solvers.py), optionally captured as one CUDA graph
per (batch, length, N, solver).
"""
import functools
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .sampling.solvers import get_solver, stages, t_grid
from .util.other import padded_frames


def _on_model_device(fn):
    """Run a method with the model's GPU as the current CUDA device: the ctypes launches, the tensor-map
    encoder and torch.cuda.current_stream() all act on the *current* device, which need not be the model's
    (reference CLI: `enhance.py --device cuda:1`)."""
    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        dev = self.device
        if dev.type != "cuda":
            return fn(self, *a, **k)
        with torch.cuda.device(dev):
            return fn(self, *a, **k)
    return wrapped


class EnhancementModel(nn.Module):
    """inference-only stand-in for the reference's LightningModule base (model.py:37-190)"""

    strict_loading = False

    def __init__(self, backbone: nn.Module, feature_extractor, sampling_rate: int, lr: float = 1e-4,
                 normalize_mode: str = "noisy", optimizer_init=None, datamodule=None, eval_metrics=None,
                 eval_variants=None, num_eval_files: int = 20, full_config: Optional[dict] = None,
                 evaluation_seed: Optional[int] = None):
        super().__init__()
        self.sampling_rate = sampling_rate
        self.normalize_mode = normalize_mode
        assert self.normalize_mode in ("noisy", "none")
        self.lr = lr
        self.backbone = backbone
        self.feature_extractor = feature_extractor
        self.full_config = full_config
        self.hparams = full_config
        self.eval_metrics, self.eval_variants = eval_metrics, eval_variants
        self.num_eval_files, self.evaluation_seed = num_eval_files, evaluation_seed
        self.datamodule, self.optimizer_init = datamodule, optimizer_init

    @property
    def device(self):
        return next(self.parameters()).device

    def set_precision(self, precision):
        """"bf16" (default, the benchmarked mode) or "tf32" (fp32 activations, tf32 tensor-core operands: the
        precision class of the reference's own GPU path); see NCSNpp.set_precision"""
        if hasattr(self, "reset_cache"):
            self.reset_cache()
        self.backbone.set_precision(precision)
        return self

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, ema=True, build_fn=None, **kwargs):
        """What the reference intends (model.py:352-385, commented out there): build the model, then
        load `_pl_ema_state_dict` (ema=True) or `state_dict`.  Hydra is not a dependency here, so the
        model comes from `build_fn()` (default: the shipped flowdec_75m configuration, or the ScoreDec
        baseline when called on ScoreModel).  A checkpoint whose backbone / feature-extractor keys do not
        match the model raises instead of silently leaving random weights in place."""
        try:
            ckpt = torch.load(checkpoint_path, map_location=map_location, weights_only=True)
        except Exception:    # Lightning checkpoints pickle hyper-parameter objects next to the tensors
            ckpt = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
        if build_fn is not None:
            model = build_fn()
        else:
            model = build_scoredec() if issubclass(cls, ScoreModel) else build_flowdec("75m")
        key = "_pl_ema_state_dict" if ema else "state_dict"
        if key not in ckpt:
            raise KeyError(f"{checkpoint_path}: no '{key}' entry (keys: {sorted(ckpt)[:8]})")
        res = model.load_state_dict(ckpt[key], strict=False)
        tolerated = ("sigma_x", "sigma_y")     # absent from ScoreDec / older checkpoints
        missing = [k for k in res.missing_keys if k not in tolerated]
        unexpected = [k for k in res.unexpected_keys if k not in tolerated]
        if missing or unexpected:
            raise RuntimeError(f"{checkpoint_path} does not match {type(model).__name__}: "
                               f"{len(missing)} missing keys (e.g. {missing[:3]}), "
                               f"{len(unexpected)} unexpected keys (e.g. {unexpected[:3]})")
        return model


class FlowModel(EnhancementModel):
    """Flow-matching postfilter (reference model.py:391-536), inference path."""

    def __init__(self, flow_matcher, sigma_x, sigma_y, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.flow_matcher = flow_matcher
        as_t = lambda v: v if isinstance(v, torch.Tensor) else torch.tensor(float(v))
        if callable(sigma_x) or callable(sigma_y):
            raise NotImplementedError("callable sigma_x / sigma_y (reference model.py:407-419) is a training-time option")
        self.sigma_x = nn.Parameter(as_t(sigma_x), requires_grad=False)
        self.sigma_y = nn.Parameter(as_t(sigma_y), requires_grad=False)
        # static buffers + CUDA graph per (batch, padded-frame bucket, N, solver, sigma_fac); least recently
        # used entries (and the backbone workspaces only they used) are dropped beyond `graph_cache_size`
        self._graphs = OrderedDict()
        self.graph_cache_size = 4
        self.use_cuda_graph = True
        self.max_batch = 16         # clips per backbone pass (micro-batch); 16 x 2 s = 8 x 4 s = 4096 frames
        self.max_frames_per_pass = 4096   # padded STFT frames per backbone pass (8 clips x 4 s)
        self.overlap_streams = 2    # micro-batches in flight on separate CUDA streams
        self._streams = []
        self._sig_cache = None

    def reset_cache(self):
        self._graphs = OrderedDict()
        self._sig_cache = None
        if hasattr(self.backbone, "pin_workspaces"):
            self.backbone.pin_workspaces(set(), owner=id(self))

    def _apply(self, fn, *a, **k):
        before = (self.sigma_y.data_ptr(), self.sigma_y.device)
        r = super()._apply(fn, *a, **k)
        if (self.sigma_y.data_ptr(), self.sigma_y.device) != before:
            self.reset_cache()
            self._streams = []
        return r

    def load_state_dict(self, state_dict, strict=None, **kw):
        self.reset_cache()
        return super().load_state_dict(state_dict, strict=self.strict_loading if strict is None else strict, **kw)

    # ------------------------------------------------------------------------------------
    def forward(self, xt, y, t):
        if torch.is_tensor(t) and t.ndim == 0:
            t = t.unsqueeze(0)
        return self.backbone(xt, y, t)

    def _micro_batch(self, Tp):
        """clips per backbone pass: bounded by `max_batch` and by `max_frames_per_pass` STFT frames, so
        that long clips (the reference CLI accepts up to 30 s, enhance.py:115) keep the activation
        workspace at the size it has for 8 x 4 s"""
        return max(1, min(self.max_batch, self.max_frames_per_pass // max(Tp, 1)))

    def _chunks(self, B, Tp):
        mb = self._micro_batch(Tp)
        return [(lo, min(B, lo + mb)) for lo in range(0, B, mb)]

    def _side_streams(self, n):
        while len(self._streams) < n:
            self._streams.append(torch.cuda.Stream(device=self.device))
        return self._streams[:n]

    def _sigma_vec(self, Fq):
        if self._sig_cache is None:
            s = self.sigma_y.detach().to(torch.float64).reshape(-1)
            if s.numel() == 1:
                s = s.expand(Fq)
            self._sig_cache = s.contiguous()
        if self._sig_cache.numel() != Fq:
            raise ValueError(f"sigma_y has {self._sig_cache.numel()} entries, spectrogram has {Fq} bins")
        return self._sig_cache

    def _get_noise(self, x, sigma):
        """reference model.py:530-536 (used by enhance unless `noise=` is injected)"""
        return (sigma * torch.randn_like(x)).type(x.dtype)

    # ------------------------------------------------------------------------------------
    def _run(self, st, N, solver, sigma_fac, want_traj):
        """All device work of enhance() on static buffers `st` (graph-capturable)."""
        fe, bb = self.feature_extractor, self.backbone
        B, L, Tp = st["B"], st["L"], st["Tp"]
        lens = st["len"]              # int32 [B]: per-clip sample counts (rows of pitch L = Tp*384)
        ops.normfac(st["y"], 1 if self.normalize_mode == "noisy" else 0, st["nf"], lengths=lens)
        fe.stft_compress(st["y"], st["nf"], st["Y"], lengths=lens)
        ops.x0_noise(st["Y"], self._sigma_vec(768), st["eps"], sigma_fac, st["x"][0])
        cur = 0
        traj = [st["x"][0].clone()] if want_traj else None
        # packed weights (every fused-skip / multi-source variant) and the time-embedding biases are produced
        # on this stream BEFORE any lane forks: side lanes only ever read them
        bb.prepare()
        for (t, dt) in t_grid(N):
            for stage in stages(solver, t, dt):
                bb.temb_biases(float(stage[0]))
        chunks = self._chunks(B, Tp)
        nlanes = max(1, min(self.overlap_streams, len(chunks)))
        for (t, dt) in t_grid(N):
            bufs = {"x": st["x"][cur], "xn": st["x"][cur ^ 1], "tmp": st["tmp"]}
            for (te, src, dst, b1, c1, b2, c2, coef) in stages(solver, t, dt):
                # micro-batches are independent: alternate them over `overlap_streams` CUDA streams so
                # the HBM-bound GroupNorm/FIR passes of one overlap the tensor-bound convs of another
                main = torch.cuda.current_stream()
                lanes = [main] + self._side_streams(nlanes - 1)
                for s_ in lanes[1:]:
                    s_.wait_stream(main)
                for i, (lo, hi) in enumerate(chunks):
                    sl = lambda name: bufs[name][lo:hi] if name is not None else None
                    with torch.cuda.stream(lanes[i % nlanes]):
                        bb.velocity(sl(src), st["Y"][lo:hi], float(te), out=sl(dst), base1=sl(b1), c1=c1,
                                    base2=sl(b2), c2=c2, coef=coef, lane=i % nlanes)
                for s_ in lanes[1:]:
                    main.wait_stream(s_)
            cur ^= 1
            if want_traj:
                traj.append(st["x"][cur].clone())
        fe.istft_decompress(st["x"][cur], L, st["nf"], st["out"], lengths=lens, ws=st["fr"])
        return traj

    def _static(self, B, Tp, dev):
        """static buffers of one (batch, padded-frame bucket): waveform rows have pitch Tp*384 >= any clip length
        of the bucket, so every L that pads to Tp frames shares the entry (and its captured graph)"""
        L = Tp * 384
        f32 = dict(device=dev, dtype=torch.float32)
        return dict(B=B, L=L, Tp=Tp, y=torch.zeros(B, L, **f32), nf=torch.empty(B, **f32),
                    len=torch.empty(B, device=dev, dtype=torch.int32),
                    Y=torch.empty(B, 768, Tp, 2, **f32), eps=torch.empty(B, 768, Tp, 2, **f32),
                    x=[torch.empty(B, 768, Tp, 2, **f32) for _ in range(2)],
                    tmp=torch.empty(B, 768, Tp, 2, **f32), out=torch.empty(B, L, **f32),
                    fr=torch.empty(B, Tp, 1536, **f32))      # iSTFT workspace: windowed frames before overlap-add

    def _entry(self, B, Tp, N, solver, sigma_fac, dev):
        # the micro-batch split and the lane count are baked into a captured graph
        key = (B, Tp, int(N), solver, float(sigma_fac), self._micro_batch(Tp), int(self.overlap_streams))
        gen = getattr(self.backbone, "generation", 0)
        if self._graphs and next(iter(self._graphs.values()))["gen"] != gen:
            self.reset_cache()       # the (possibly shared) backbone repacked its weights / dropped its workspaces
        entry = self._graphs.get(key)
        if entry is None:
            while len(self._graphs) >= max(1, self.graph_cache_size):
                self._graphs.popitem(last=False)
            entry = dict(st=self._static(B, Tp, dev), graph=None, warm=0, gen=gen,
                         sigs={(hi - lo, 768, Tp) for lo, hi in self._chunks(B, Tp)})
            self._graphs[key] = entry
            if hasattr(self.backbone, "pin_workspaces"):
                # workspaces of evicted entries become collectable; live graphs keep theirs (stable addresses)
                self.backbone.pin_workspaces(set().union(*[e["sigs"] for e in self._graphs.values()]), owner=id(self))
        else:
            self._graphs.move_to_end(key)
        return entry

    def _enhance_bucket(self, y2d, lens, N, solver, sigma_fac, noise, want_traj):
        """y2d: [B, Lin] (host or device) whose row b holds lens[b] valid samples; all clips share one
        padded-frame bucket.  Returns the static-buffer dict (st["out"][b, :lens[b]] is the enhanced clip)
        and the trajectory (or None)."""
        dev = self.device
        B, Lin = y2d.shape
        Tp = padded_frames(1 + max(lens) // 384)
        entry = self._entry(B, Tp, N, solver, sigma_fac, dev)
        st = entry["st"]
        n_copy = min(Lin, st["L"])
        if entry.get("filled", 0) > n_copy:
            st["y"][:, n_copy:entry["filled"]].zero_()      # a longer clip of the bucket was here before
        entry["filled"] = n_copy
        st["y"][:, :n_copy].copy_(y2d[:, :n_copy], non_blocking=True)
        st["len"].copy_(torch.tensor(lens, dtype=torch.int32), non_blocking=True)
        if noise is None:
            eps = torch.randn(B, 768, Tp, dtype=torch.complex64, device=dev)
        else:
            eps = noise.to(dev, torch.complex64).reshape(B, 768, Tp)
        st["eps"].copy_(torch.view_as_real(eps))
        traj = None
        if want_traj or not self.use_cuda_graph:
            traj = self._run(st, N, solver, sigma_fac, want_traj)
        elif entry["graph"] is None:
            # first call: eager (also builds packed weights / time-embedding caches); second: capture
            if entry["warm"] == 0:
                self._run(st, N, solver, sigma_fac, False)
                entry["warm"] = 1
            else:
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    self._run(st, N, solver, sigma_fac, False)
                entry["graph"] = g
                g.replay()
        else:
            entry["graph"].replay()
        return st, traj

    @torch.no_grad()
    @_on_model_device
    def enhance(self, y, return_preprocess_info: bool = False, N: int = 50, solver: str = "euler",
                with_grad: bool = False, sigma_fac: float = 1.0, return_traj: bool = False,
                noise: Optional[torch.Tensor] = None, lengths=None, **kwargs):
        """Enhance a coded waveform `y` ([B,1,L], [1,L] or [L]); reference model.py:476-528.

        Extra keyword `noise`: complex64 [B,1,768,Tp] standing in for torch.randn_like(Y) (parity
        tests inject the oracle's draw).  Extra keyword `lengths` (B ints): `y` is a zero-padded ragged
        batch and clip b has lengths[b] valid samples; all clips must fall into the same padded-frame
        bucket (64*ceil((1 + L_b//384)/64)), and each then gets exactly the computation it would get alone
        (own normfac, frames, reflect padding and istft length; samples beyond L_b come back as 0) — see
        flowdec_b200/batching.py.  Unknown kwargs (predictor/corrector/snr from the reference CLI,
        enhance.py:55-61) are accepted and ignored like the reference does."""
        if with_grad:
            raise NotImplementedError("with_grad=True (backprop through the solver) is a training feature")
        if lengths is not None:
            return self._enhance_ragged(y, lengths, N, solver, sigma_fac, noise, return_preprocess_info, return_traj)
        solver = get_solver(solver)
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("flowdec_b200 runs on CUDA (sm_100a) only; call model.cuda()")
        squeeze_dims = 0
        y_in = y
        while y.ndim < 3:
            y = y.unsqueeze(0)
            squeeze_dims += 1
        B, C, L = y.shape
        if L <= 767:
            raise ValueError(f"waveform length {L} must exceed the STFT reflect pad (767)")
        st, traj = self._enhance_bucket(y.reshape(B * C, L), [L] * (B * C), N, solver, sigma_fac, noise, return_traj)
        Tp = st["Tp"]
        info = dict(orig_length=L, normfac=st["nf"].clone().reshape(B, C, 1) if C == 1 else st["nf"].clone(),
                    undo_pad_fn=(lambda Y_, T=1 + L // 384: Y_[..., :T]), squeeze_dims=squeeze_dims)
        if return_traj:
            fe = self.feature_extractor
            X_hats, x_hats = [], []
            for X in traj:
                w = torch.empty(B * C, L, device=dev, dtype=torch.float32)
                fe.istft_decompress(X, L, st["nf"], w)
                X_hats.append(torch.view_as_complex(X).reshape(B, C, 768, Tp))
                xw = w.reshape(B, C, L)
                for _ in range(squeeze_dims):
                    xw = xw.squeeze(0)
                x_hats.append(xw)
            return torch.stack(X_hats), x_hats
        x_hat = st["out"][:, :L].clone().reshape(B, C, L)
        for _ in range(squeeze_dims):
            x_hat = x_hat.squeeze(0)
        x_hat = x_hat.to(y_in.device)
        return (x_hat, info) if return_preprocess_info else x_hat

    def _enhance_ragged(self, y, lengths, N, solver, sigma_fac, noise, return_preprocess_info, return_traj):
        """length-bucketed batch: see enhance(lengths=)"""
        if return_traj:
            raise NotImplementedError("return_traj is not available for ragged batches")
        solver = get_solver(solver)
        dev = self.device
        if y.ndim == 2:
            y = y.unsqueeze(1)
        if y.ndim != 3 or y.shape[1] != 1:
            raise ValueError(f"ragged batches are [B,1,Lmax] (or [B,Lmax]), got {tuple(y.shape)}")
        lens = [int(v) for v in (lengths.tolist() if torch.is_tensor(lengths) else lengths)]
        B, Lin = y.shape[0], y.shape[2]
        if len(lens) != B:
            raise ValueError(f"{len(lens)} lengths for a batch of {B}")
        if min(lens) <= 767 or max(lens) > Lin:
            raise ValueError(f"clip lengths must be in (767, {Lin}], got min {min(lens)} max {max(lens)}")
        buckets = {padded_frames(1 + n // 384) for n in lens}
        if len(buckets) != 1:
            raise ValueError(f"clips of one ragged batch must share the padded-frame bucket, got {sorted(buckets)}; "
                             "use flowdec_b200.batching.enhance_list to bucket a list of clips")
        if dev.type != "cuda":
            raise RuntimeError("flowdec_b200 runs on CUDA (sm_100a) only; call model.cuda()")
        st, _ = self._enhance_bucket(y.reshape(B, Lin), lens, N, solver, sigma_fac, noise, False)
        Lmax = max(lens)
        x_hat = st["out"][:, :Lmax].clone().reshape(B, 1, Lmax).to(y.device)
        if not return_preprocess_info:
            return x_hat
        info = dict(orig_length=lens, normfac=st["nf"].clone().reshape(B, 1, 1),
                    undo_pad_fn=(lambda Y_, T=[1 + n // 384 for n in lens]: [Y_[i, ..., :t] for i, t in enumerate(T)]),
                    squeeze_dims=0)
        return x_hat, info


class ScoreModel(EnhancementModel):
    """Score-based baseline (ScoreDec / SGMSE+) around the same backbone; reference model.py:583-688.
    Inference path: predictor-corrector sampler with the reverse-diffusion predictor and the
    annealed-Langevin corrector (sampling/__init__.py:32-72, predictors.py:61-71,
    correctors.py:43-66).  Each update is affine in (x, y, noise, backbone output) and is fused
    into the backbone's last kernel, so one sampler step = 1 + corrector_steps kernel chains."""

    def __init__(self, sde, t_eps, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.sde = sde
        self.t_eps = t_eps
        self.max_batch = 8
        self.use_cuda_graph = True
        self._sgraphs = OrderedDict()     # static buffers + CUDA graph per (batch, length, N, corrector_steps, ...), LRU of 2

    def sde_std(self, t_batch):
        t = float(t_batch.flatten()[0]) if torch.is_tensor(t_batch) else float(t_batch)
        return self.sde._std(t)

    def forward(self, xt, y, t_batch):
        """score = -backbone(xt, y, t) / sigma_t (model.py:613-628); shared scalar t."""
        std = float(self.sde_std(t_batch))
        return -self.backbone(xt, y, t_batch) / std

    def reset_cache(self):
        self._sgraphs = OrderedDict()
        if hasattr(self.backbone, "pin_workspaces"):
            self.backbone.pin_workspaces(set(), owner=id(self))

    def _apply(self, fn, *a, **k):
        gen = getattr(self.backbone, "generation", 0)
        r = super()._apply(fn, *a, **k)
        if getattr(self.backbone, "generation", 0) != gen:
            self.reset_cache()
        return r

    def load_state_dict(self, state_dict, strict=None, **kw):
        self.reset_cache()
        return super().load_state_dict(state_dict, strict=self.strict_loading if strict is None else strict, **kw)

    def _sampler(self, st, N, corrector_steps, snr, denoise, probability_flow):
        """the whole PC sampler on static buffers (graph-capturable): draws come from st["z"][i] in call order"""
        fe, bb, sde = self.feature_extractor, self.backbone, self.sde.copy()
        sde.N = N
        B, L, Tp = st["B"], st["L"], st["Tp"]
        Y, x, xn, x_mean, z = st["Y"], st["x"], st["xn"], st["x_mean"], st["z"]
        pf = 0.5 if probability_flow else 1.0
        ops.normfac(st["y"], 1 if self.normalize_mode == "noisy" else 0, st["nf"])
        fe.stft_compress(st["y"], st["nf"], Y)
        mb = max(1, min(self.max_batch, 4096 // Tp))
        nz = iter(range(z.shape[0]))

        def backbone_stage(x, t, out, c1, c2, base3, c3, coef):
            for lo in range(0, B, mb):
                hi = min(B, lo + mb)
                bb.velocity(x[lo:hi], Y[lo:hi], float(t), out=out[lo:hi], base1=x[lo:hi], c1=c1,
                            base2=Y[lo:hi], c2=c2, base3=base3[lo:hi] if base3 is not None else None,
                            c3=c3, coef=coef)

        # prior: x_T = y + z * std(T)     (sdes.py:201-206); x0_kernel with a unit sigma vector is out = Y + fac*z
        ops.x0_noise(Y, st["ones64"], z[next(nz)], float(sde._std(1.0)), x)
        timesteps = torch.linspace(sde.T, self.t_eps, N).numpy()
        for i in range(N):
            t = timesteps[i]
            std = float(sde._std(t))
            for _ in range(corrector_steps):
                # ALD (correctors.py:54-66): x <- x + step*score + sqrt(2 step) z, score = -v/std
                step = (snr * std) ** 2 * 2
                backbone_stage(x, t, xn, 1.0, 0.0, z[next(nz)], float(np.sqrt(np.float32(step * 2))), -step / std)
                x, xn = xn, x
            # reverse diffusion (predictors.py:66-71, sdes.py:118-123):
            #   rev_f = theta dt (y - x) - G^2 score ; x_mean = x - rev_f ; x = x_mean + G z
            th_dt, G = sde.discretize(t)
            th_dt, G = float(th_dt), float(G)
            last = (i == N - 1)
            zi = None if (last and denoise) else z[next(nz)]
            if last:
                backbone_stage(x, t, x_mean, 1.0 + th_dt, -th_dt, None, 0.0, -pf * G * G / std)
                if not denoise and not probability_flow:
                    ops.x0_noise(x_mean, st["ones64"], zi, G, x_mean)
            else:
                backbone_stage(x, t, xn, 1.0 + th_dt, -th_dt, zi if not probability_flow else None,
                               0.0 if probability_flow else G, -pf * G * G / std)
                x, xn = xn, x
        fe.istft_decompress(x_mean, L, st["nf"], st["out"], ws=st["fr"])

    @torch.no_grad()
    @_on_model_device
    def enhance(self, y, sampler_type="pc", predictor="reverse_diffusion", corrector="ald", N=30,
                corrector_steps=1, snr=0.5, return_preprocess_info=False, denoise=True,
                probability_flow=False, noise=None, **kwargs):
        """reference model.py:630-657.  `noise`: optional list of complex64 [B,1,768,Tp] tensors standing
        in for the sampler's draws, in call order: prior, then per step (corrector draws..., predictor draw).
        All draws of a call are made up front into a static buffer, so the sampler loop (N x (1 + corrector_steps)
        backbone evaluations with their fused updates) replays as ONE CUDA graph per (batch, length, N, ...)."""
        if sampler_type != "pc" or predictor not in ("reverse_diffusion", "euler_maruyama") or \
                corrector not in ("ald", "none"):
            raise NotImplementedError("flowdec_b200.ScoreModel implements the PC sampler with the reverse_diffusion / "
                                      "euler_maruyama predictors and the ald / none correctors (the scipy RK45 "
                                      "'ode' sampler of sampling/__init__.py:75-147 is host-driven and out of scope)")
        # For the OUVE SDE the two predictors are the same update: Euler-Maruyama (predictors.py:47-59) steps
        # x + [theta (y - x) - g^2 score] (-1/N) + g sqrt(1/N) z, reverse diffusion (predictors.py:61-71) steps
        # x - [theta (y - x)/N - G^2 score] + G z with G = g sqrt(1/N) (sdes.py:68-76): identical coefficients.
        # probability_flow (sdes.py:107-123): half the score term, no predictor noise.
        if corrector == "none":
            corrector_steps = 0
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("flowdec_b200 runs on CUDA (sm_100a) only; call model.cuda()")
        y_in = y
        squeeze_dims = 0
        while y.ndim < 3:
            y = y.unsqueeze(0)
            squeeze_dims += 1
        B, C, L = y.shape
        if L <= 767:
            raise ValueError(f"waveform length {L} must exceed the STFT reflect pad (767)")
        Tp = padded_frames(1 + L // 384)
        n_draws = 1 + N * corrector_steps + N - (1 if denoise else 0)
        key = (B * C, L, int(N), int(corrector_steps), float(snr), bool(denoise), bool(probability_flow))
        gen = getattr(self.backbone, "generation", 0)
        if self._sgraphs and next(iter(self._sgraphs.values()))["gen"] != gen:
            self.reset_cache()
        entry = self._sgraphs.get(key)
        if entry is None:
            while len(self._sgraphs) >= 2:
                self._sgraphs.popitem(last=False)
            f32 = dict(device=dev, dtype=torch.float32)
            mb = max(1, min(self.max_batch, 4096 // Tp))
            spec = lambda: torch.empty(B * C, 768, Tp, 2, **f32)
            st = dict(B=B * C, L=L, Tp=Tp, y=torch.empty(B * C, L, **f32), nf=torch.empty(B * C, **f32), Y=spec(),
                      x=spec(), xn=spec(), x_mean=spec(), z=torch.empty(n_draws, B * C, 768, Tp, 2, **f32),
                      out=torch.empty(B * C, L, **f32), fr=torch.empty(B * C, Tp, 1536, **f32),
                      ones64=torch.ones(768, device=dev, dtype=torch.float64))
            entry = dict(st=st, graph=None, warm=0, gen=gen,
                         sigs={(min(B * C, lo + mb) - lo, 768, Tp) for lo in range(0, B * C, mb)})
            self._sgraphs[key] = entry
            self.backbone.pin_workspaces(set().union(*[e["sigs"] for e in self._sgraphs.values()]), owner=id(self))
        else:
            self._sgraphs.move_to_end(key)
        st = entry["st"]
        st["y"].copy_(y.reshape(B * C, L), non_blocking=True)
        if noise is None:
            st["z"].normal_(0.0, 0.5 ** 0.5)       # complex standard normal: variance 1/2 per component (randn_like)
        else:
            noise = list(noise)
            if len(noise) < n_draws:
                raise ValueError(f"the sampler makes {n_draws} draws, got {len(noise)} noise tensors")
            for i in range(n_draws):
                st["z"][i].copy_(torch.view_as_real(noise[i].to(dev, torch.complex64).reshape(B * C, 768, Tp)))
        args = (st, N, corrector_steps, snr, denoise, probability_flow)
        if entry["graph"] is not None:
            entry["graph"].replay()
        elif entry["warm"] == 0 or not getattr(self, "use_cuda_graph", True):
            self._sampler(*args)                  # first call: eager (builds packed weights / temb caches)
            entry["warm"] = 1
        else:
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                self._sampler(*args)
            entry["graph"] = g
            g.replay()
        x_hat = st["out"].clone().reshape(B, C, L)
        for _ in range(squeeze_dims):
            x_hat = x_hat.squeeze(0)
        x_hat = x_hat.to(y_in.device)
        info = dict(orig_length=L, normfac=st["nf"].clone().reshape(B, C, 1),
                    undo_pad_fn=(lambda Y_, T=1 + L // 384: Y_[..., :T]), squeeze_dims=squeeze_dims)
        return (x_hat, info) if return_preprocess_info else x_hat


def build_scoredec(device=None):
    """the model `config/baseline_scoredec_75s.yaml` instantiates (score_model_final.yaml + ouve_final.yaml)"""
    from .sdes import OUVESDE
    fm = build_flowdec("75m")
    m = ScoreModel(OUVESDE(theta=1.5, sigma_min=0.05, sigma_max=0.82, N=30), 3e-2, backbone=fm.backbone,
                   feature_extractor=fm.feature_extractor, sampling_rate=48000, lr=1e-4)
    m.eval()
    return m.to(device) if device is not None else m


def build_flowdec(variant="75m", device=None):
    """The model `config/flowdec_{75m,25s}.yaml` instantiates (hydra replaced by direct calls)."""
    from .backbones.ncsnpp import NCSNpp
    from .data.feature_extractors import AmplitudeCompressedComplexSTFT
    from .data.sigma_models import from_file
    backbone = NCSNpp(image_size=768, nonlinearity="swish", nf=64, ch_mult=[4, 4, 4, 2], num_res_blocks=1,
                      attn_resolutions=[], bottleneck_attn=False, resamp_with_conv=True, conditional=True,
                      fir=True, fir_kernel=[1, 3, 3, 1], skip_rescale=True, resblock_type="biggan",
                      progressive="output_skip", progressive_input="input_skip", progressive_combine="sum",
                      init_scale=0.0, embedding_type="fourier", fourier_scale=16, dropout=0.0, num_channels=4,
                      output_layer_kwargs=dict(kernel_size=1, bias=False, padding="same", padding_mode="zeros"))
    fe = AmplitudeCompressedComplexSTFT(window_fn="hann", n_fft=1534, sampling_rate=48000, alpha=0.3, beta=0.33,
                                        n_hops=4)
    sigma_y = from_file(f"flowdec_autoparams_{variant}.npy", factor=1, kernel_bandwidth=3)
    m = FlowModel(None, 0.0, sigma_y, backbone=backbone, feature_extractor=fe, sampling_rate=48000, lr=1e-4)
    m.eval()
    return m.to(device) if device is not None else m
