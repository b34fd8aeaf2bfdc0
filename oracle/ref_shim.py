"""TEST INFRASTRUCTURE — not part of the product path.

Imports the *unmodified* reference (`/root/reference/flowdec`) on CPU so that the oracle
restatement (oracle/flowdec_oracle.py) can be pinned against it and golden vectors can be
generated (oracle/make_golden.py).  Only usable where /root/reference exists (the build
container); nothing on the GPU box imports this module.

The reference's un-vendored pip dependencies are absent here and are stubbed:
  * pytorch_lightning, omegaconf, hydra, wandb(optional), torchcfm  -> inert stand-ins
  * torchdyn==1.0.6 (requirements.txt:53): `NeuralODE.trajectory` + "euler"/"midpoint"
    restated from its published fixed-step algorithm (call sites: flowdec/model.py:511-514,
    flowdec/sampling/solvers.py:12,24-33)
  * torch.utils.cpp_extension.load -> dummy (the op's CUDA JIT build is skipped; on CPU
    tensors the reference itself routes to its pure-torch `upfirdn2d_native`,
    op/upfirdn2d.py:170-173)
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REF_ROOT = os.environ.get("FLOWDEC_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "flowdec"))


# --------------------------------------------------------------------------- torchdyn restatement
class DiffEqSolver(nn.Module):
    def __init__(self, order=1, stepping_class="fixed", **kw):
        super().__init__()
        self.order = order
        self.stepping_class = stepping_class


class _Euler(DiffEqSolver):
    def step(self, f, x, t, dt, k1=None, args=None):
        if k1 is None:
            k1 = f(t, x)
        return None, x + dt * k1, None


class _Midpoint(DiffEqSolver):
    def step(self, f, x, t, dt, k1=None, args=None):
        if k1 is None:
            k1 = f(t, x)
        x_mid = x + 0.5 * dt * k1
        return None, x + dt * f(t + 0.5 * dt, x_mid), None


class NeuralODE(nn.Module):
    """fixed-step subset of torchdyn.core.NeuralODE used by FlowModel.enhance"""

    def __init__(self, vector_field, solver="euler", sensitivity="adjoint", **kw):
        super().__init__()
        self.vf = vector_field
        if isinstance(solver, str):
            solver = {"euler": _Euler, "midpoint": _Midpoint}[solver]()
        self.solver = solver
        self.nfe = 0

    def trajectory(self, x, t_span):
        def f(t, x_):
            self.nfe += 1
            return self.vf(t, x_)

        t = t_span[0]
        dt = t_span[1] - t_span[0]
        sol = [x]
        steps = 1
        while steps <= len(t_span) - 1:
            _, x, _ = self.solver.step(f, x, t, dt)
            sol.append(x)
            t = t + dt
            if steps < len(t_span) - 1:
                dt = t_span[steps + 1] - t
            steps += 1
        return torch.stack(sol)


# --------------------------------------------------------------------------- stubs
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _LightningModule(nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def log(self, *a, **k):
        pass


def _install_stubs():
    if "pytorch_lightning" not in sys.modules:
        pl = _mod("pytorch_lightning", LightningModule=_LightningModule,
                  LightningDataModule=object, Callback=object, Trainer=object)
        _mod("pytorch_lightning.loggers", WandbLogger=object, TensorBoardLogger=object)
        _mod("pytorch_lightning.plugins")
        _mod("pytorch_lightning.plugins.environments", SLURMEnvironment=object)
        _mod("pytorch_lightning.callbacks", ModelCheckpoint=object, Callback=object)
        _mod("pytorch_lightning.utilities")
        _mod("pytorch_lightning.utilities.rank_zero", rank_zero_info=print, rank_zero_only=lambda f: f)
        _mod("pytorch_lightning.utilities.types", STEP_OUTPUT=object)
        _mod("pytorch_lightning.utilities.exceptions", MisconfigurationException=Exception)
        pl.loggers = sys.modules["pytorch_lightning.loggers"]
    if "omegaconf" not in sys.modules:
        class _OC:
            @staticmethod
            def create(x):
                return x

            @staticmethod
            def to_yaml(x):
                return str(x)
        _mod("omegaconf", OmegaConf=_OC, DictConfig=dict)
    if "hydra" not in sys.modules:
        _mod("hydra")
        _mod("hydra.utils", instantiate=lambda *a, **k: None)
    if "torchcfm" not in sys.modules:
        _mod("torchcfm", ConditionalFlowMatcher=object)
    if "torchdyn" not in sys.modules:
        _mod("torchdyn")
        _mod("torchdyn.core", NeuralODE=NeuralODE)
        _mod("torchdyn.numerics")
        _mod("torchdyn.numerics.solvers")
        _mod("torchdyn.numerics.solvers.templates", DiffEqSolver=DiffEqSolver)
    try:
        import wandb  # noqa: F401
    except Exception:
        _mod("wandb")


_loaded = None


def load_reference():
    """Returns a namespace with the reference's FlowModel, NCSNpp, feature extractor, from_file, ops."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    _install_stubs()
    # synthetic top-level package so flowdec/__init__.py (imports everything) is bypassed
    pkg = types.ModuleType("flowdec")
    pkg.__path__ = [os.path.join(REF_ROOT, "flowdec")]
    saved = sys.modules.get("flowdec")
    sys.modules["flowdec"] = pkg
    # metrics module pulls librosa/pesq/...: stub the two names model.py imports
    _mod("flowdec.eval", )
    sys.modules["flowdec.eval"].__path__ = []
    _mod("flowdec.eval.metrics", Metric=object, get_metrics_row=lambda *a, **k: {})
    import torch.utils.cpp_extension as cpp
    real_load = cpp.load
    cpp.load = lambda *a, **k: types.SimpleNamespace()
    try:
        model = importlib.import_module("flowdec.model")
        ncsnpp = importlib.import_module("flowdec.backbones.ncsnpp")
        fe = importlib.import_module("flowdec.data.feature_extractors")
        sig = importlib.import_module("flowdec.data.sigma_models")
        other = importlib.import_module("flowdec.util.other")
        updown = importlib.import_module("flowdec.backbones.ncsnpp_utils.up_or_down_sampling")
        layerspp = importlib.import_module("flowdec.backbones.ncsnpp_utils.layerspp")
        solvers = importlib.import_module("flowdec.sampling.solvers")
    finally:
        cpp.load = real_load
    _loaded = types.SimpleNamespace(model=model, ncsnpp=ncsnpp, fe=fe, sigma_models=sig, other=other,
                                    updown=updown, layerspp=layerspp, solvers=solvers, saved=saved)
    return _loaded


BACKBONE_KW = dict(  # config/model/backbone/ncsnpp_final_no_attn.yaml:8-33
    image_size=768, nonlinearity="swish", nf=64, ch_mult=[4, 4, 4, 2], num_res_blocks=1,
    attn_resolutions=[], bottleneck_attn=False, resamp_with_conv=True, conditional=True, fir=True,
    fir_kernel=[1, 3, 3, 1], skip_rescale=True, resblock_type="biggan", progressive="output_skip",
    progressive_input="input_skip", progressive_combine="sum", init_scale=0.0,
    embedding_type="fourier", fourier_scale=16, dropout=0.0, num_channels=4,
    output_layer_kwargs=dict(kernel_size=1, bias=False, padding="same", padding_mode="zeros"),
)


def build_reference_model(variant="75m", backbone_kw=None):
    """FlowModel exactly as config/flowdec_{75m,25s}.yaml instantiates it (hydra replaced by hand)."""
    R = load_reference()
    import warnings
    kw = dict(BACKBONE_KW)
    kw.update(backbone_kw or {})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        backbone = R.ncsnpp.NCSNpp(**kw)
        fe = R.fe.AmplitudeCompressedComplexSTFT(window_fn="hann", n_fft=1534, sampling_rate=48000,
                                                 alpha=0.3, beta=0.33, n_hops=4)
        sigma_y = R.sigma_models.from_file(
            os.path.join(REF_ROOT, "data", f"flowdec_autoparams_{variant}.npy"), factor=1, kernel_bandwidth=3)
        m = R.model.FlowModel(None, 0.0, sigma_y, backbone=backbone, feature_extractor=fe,
                              sampling_rate=48000, lr=1e-4)
    m.eval()
    return m
