"""Upstream NDAC (DAC) codec on B200 (SURVEY.md §8 a11 decode, §8f-1 encode).

API mirror of the part of descript-audio-codec 1.0.0 the reference's demo uses
(/root/reference/demo.ipynb:56,101-105):

    dac_model = DAC.load(".../weights.pth"); dac_model.to("cuda"); dac_model.eval()
    x = dac_model.preprocess(signal.audio_data, signal.sample_rate)
    z, codes, latents, _, _ = dac_model.encode(x, n_quantizers=nq)
    zq, _, _ = dac_model.quantizer.from_codes(codes)
    xhat_ndac = dac_model.decode(zq)

`weights.pth` is audiotools' `{"state_dict": ..., "metadata": {"kwargs": {...}}}`; every
hyper-parameter (decoder_dim, decoder_rates, n_codebooks, codebook_size, codebook_dim,
latent_dim / encoder_dim + encoder_rates, sample_rate) comes from that metadata.  Weight-norm
(`weight_g`, `weight_v`) is folded and the codebooks are L2-normalised once at load.

All arithmetic of encode / from_codes / decode runs in csrc/fd_dac.cu / csrc/fd_dac_tc.cu; there is no fallback.

Encoder precision (`DAC.encoder_precision`): "fp32" by default (it decides discrete codes), "tf32" = the same tensor-core
GEMM kernel as the decoder.  Decoder precision (`DAC.precision`): "tf32" (default when every channel count is a multiple of 32, as in the
NDAC checkpoints: 1024 / 1536 / 768 / 384 / 192 / 96) runs each layer as a tcgen05 implicit GEMM on fp32
time-major activations with tf32 operands and fp32 accumulation (residual adds in fp32); "fp32" runs the
register-tiled CUDA-core kernels (bit-for-bit fp32 FMAs, ~50x slower).  Stated tolerance of the tf32 decoder:
waveform rel-L2 <= 5e-3 vs the fp64 oracle (tests/test_dac_gpu.py prints the measured value).
"""
import ctypes
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib


def _fold_weight_norm(sd, p):
    v, g = sd[p + ".weight_v"].float(), sd[p + ".weight_g"].float()
    return (g * v / v.norm(dim=tuple(range(1, v.ndim)), keepdim=True)).contiguous()


class _Quantizer:
    def __init__(self, owner):
        self._o = owner

    def from_codes(self, codes):
        """codes int64 [B, n_q, T] -> (z_q [B, D, T], None, codes); reference call site demo.ipynb:104."""
        o = self._o
        o._require_cuda()
        torch.cuda.set_device(o.device)
        codes = codes.to(o.device, torch.int64).contiguous()
        B, nq, T = codes.shape
        if nq > o.n_codebooks:
            raise ValueError(f"codes use {nq} codebooks, model has {o.n_codebooks}")
        z = torch.empty(B, o.latent_dim, T, device=o.device, dtype=torch.float32)
        rc = _lib.lib().fd_rvq_from_codes(_lib.ptr(codes), _lib.ptr(o.codebooks), _lib.ptr(o.out_proj_w),
                                          _lib.ptr(o.out_proj_b), _lib.ptr(z), B, nq, T, o.latent_dim,
                                          o.codebook_dim, o.codebook_size, _lib.stream_ptr())
        _lib.check(rc, "fd_rvq_from_codes")
        return z, None, codes

    def __call__(self, z, n_quantizers=None):
        """ResidualVectorQuantize.forward, eval mode (dac/nn/quantize.py): z [B, D, T] ->
        (z_q [B,D,T], codes int64 [B,n_q,T], latents [B,n_q*cdim,T], commitment_loss, codebook_loss)"""
        o = self._o
        o._require_cuda()
        if o.in_proj_w is None:
            raise RuntimeError("this checkpoint has no quantizer in_proj weights (decode-only state_dict)")
        torch.cuda.set_device(o.device)
        z = z.to(o.device, torch.float32).contiguous()
        B, D, T = z.shape
        if D != o.latent_dim:
            raise ValueError(f"latent has {D} channels, model expects {o.latent_dim}")
        nq = o.n_codebooks if n_quantizers is None else max(1, min(int(n_quantizers), o.n_codebooks))
        codes = torch.empty(B, nq, T, device=o.device, dtype=torch.int64)
        zq = torch.empty_like(z)
        latents = torch.empty(B, nq * o.codebook_dim, T, device=o.device, dtype=torch.float32)
        loss = torch.empty(1, device=o.device, dtype=torch.float32)
        ws = torch.empty(nq * B * T, device=o.device, dtype=torch.float32)
        rc = _lib.lib().fd_rvq_encode(_lib.ptr(z), _lib.ptr(o.in_proj_w), _lib.ptr(o.in_proj_b),
                                      _lib.ptr(o.codebooks), _lib.ptr(o.codebooks_l2n), _lib.ptr(o.codebooks_l2n_sq),
                                      _lib.ptr(o.out_proj_w), _lib.ptr(o.out_proj_b), _lib.ptr(codes), _lib.ptr(zq),
                                      _lib.ptr(latents), _lib.ptr(loss), _lib.ptr(ws), B, nq, T, D, o.codebook_dim,
                                      o.codebook_size, _lib.stream_ptr())
        _lib.check(rc, "fd_rvq_encode")
        return zq, codes, latents, loss[0], loss[0].clone()


class DAC(nn.Module):
    """dac.DAC (dac/model/dac.py) inference — preprocess / encode / quantizer / decode — on hand-written CUDA kernels"""

    def __init__(self, state_dict, decoder_dim=1536, decoder_rates=(8, 8, 4, 2), n_codebooks=9,
                 codebook_size=1024, codebook_dim=8, latent_dim=None, encoder_dim=64,
                 encoder_rates=(2, 4, 8, 8), sample_rate=44100, **unused):
        super().__init__()
        self.decoder_dim, self.decoder_rates = int(decoder_dim), [int(r) for r in decoder_rates]
        self.n_codebooks, self.codebook_size, self.codebook_dim = int(n_codebooks), int(codebook_size), int(codebook_dim)
        self.latent_dim = int(latent_dim) if latent_dim is not None else int(encoder_dim * 2 ** len(encoder_rates))
        self.sample_rate = int(sample_rate)
        self.encoder_dim, self.encoder_rates = int(encoder_dim), [int(r) for r in encoder_rates]
        self.hop_length = int(np.prod(self.encoder_rates))
        sd = state_dict
        # --- RVQ tables
        self.register_buffer("codebooks", torch.stack(
            [sd[f"quantizer.quantizers.{i}.codebook.weight"].float() for i in range(self.n_codebooks)]).contiguous())
        self.register_buffer("out_proj_w", torch.stack(
            [_fold_weight_norm(sd, f"quantizer.quantizers.{i}.out_proj").squeeze(-1) for i in range(self.n_codebooks)]).contiguous())
        self.register_buffer("out_proj_b", torch.stack(
            [sd[f"quantizer.quantizers.{i}.out_proj.bias"].float() for i in range(self.n_codebooks)]).contiguous())
        has_in = "quantizer.quantizers.0.in_proj.weight_v" in sd
        self.register_buffer("in_proj_w", torch.stack(
            [_fold_weight_norm(sd, f"quantizer.quantizers.{i}.in_proj").squeeze(-1) for i in range(self.n_codebooks)]
        ).contiguous() if has_in else None)
        self.register_buffer("in_proj_b", torch.stack(
            [sd[f"quantizer.quantizers.{i}.in_proj.bias"].float() for i in range(self.n_codebooks)]
        ).contiguous() if has_in else None)
        # VectorQuantize.decode_latents compares L2-normalised rows: normalise the tables once (F.normalize, eps 1e-12)
        cbn = self.codebooks / self.codebooks.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        self.register_buffer("codebooks_l2n", cbn.contiguous())
        self.register_buffer("codebooks_l2n_sq", cbn.pow(2).sum(-1).contiguous())
        # --- decoder layers as a flat op list
        self._ops = []
        m = "decoder.model."

        def reg(name, t):
            # the op lists hold buffer NAMES: nn.Module._apply (.to / .cuda) replaces the buffer objects, so
            # a tensor reference taken here would stay on the CPU; `_t` resolves the name at call time
            name = name.replace(".", "_")
            self.register_buffer(name, t.float().contiguous())
            return name

        def conv(p):
            return reg(p + ".w", _fold_weight_norm(sd, p)), reg(p + ".b", sd[p + ".bias"])

        w, b = conv(m + "0")
        self._ops.append(("conv", w, b, None, 1, 3, False, False))
        for i, s in enumerate(self.decoder_rates):
            blk = f"{m}{i + 1}.block."
            a = reg(blk + "0.alpha", sd[blk + "0.alpha"].reshape(-1))
            w, b = conv(blk + "1")
            self._ops.append(("convtr", w, b, a, s, math.ceil(s / 2)))
            for j, d in enumerate((1, 3, 9)):
                r = f"{blk}{j + 2}.block."
                a1 = reg(r + "0.alpha", sd[r + "0.alpha"].reshape(-1))
                w1, b1 = conv(r + "1")
                a2 = reg(r + "2.alpha", sd[r + "2.alpha"].reshape(-1))
                w2, b2 = conv(r + "3")
                self._ops.append(("res", (w1, b1, a1, d), (w2, b2, a2)))
        n = len(self.decoder_rates)
        a = reg(f"{m}{n + 1}.alpha", sd[f"{m}{n + 1}.alpha"].reshape(-1))
        w, b = conv(f"{m}{n + 2}")
        self._ops.append(("conv", w, b, a, 1, 3, False, True))
        # --- encoder layers (dac/model/dac.py Encoder / EncoderBlock), present in full checkpoints
        self._enc_ops = None
        e = "encoder.block."
        if e + "0.weight_v" in sd:
            self._enc_ops = []
            w, b = conv(e + "0")
            self._enc_ops.append(("conv", w, b, None, 1, 3, False, False))
            for i, s in enumerate(self.encoder_rates):
                blk = f"{e}{i + 1}.block."
                for j, d in enumerate((1, 3, 9)):
                    r = f"{blk}{j}.block."
                    a1 = reg(r + "0.alpha", sd[r + "0.alpha"].reshape(-1))
                    w1, b1 = conv(r + "1")
                    a2 = reg(r + "2.alpha", sd[r + "2.alpha"].reshape(-1))
                    w2, b2 = conv(r + "3")
                    self._enc_ops.append(("res", (w1, b1, a1, d), (w2, b2, a2)))
                a = reg(blk + "3.alpha", sd[blk + "3.alpha"].reshape(-1))
                w, b = conv(blk + "4")
                self._enc_ops.append(("down", w, b, a, s, math.ceil(s / 2)))
            n = len(self.encoder_rates)
            a = reg(f"{e}{n + 1}.alpha", sd[f"{e}{n + 1}.alpha"].reshape(-1))
            w, b = conv(f"{e}{n + 2}")
            self._enc_ops.append(("conv", w, b, a, 1, 1, False, False))
        self.quantizer = _Quantizer(self)
        chans = [self.latent_dim] + [self.decoder_dim // 2 ** i for i in range(len(self.decoder_rates) + 1)]
        self.tc_eligible = all(c % 32 == 0 for c in chans)
        self.precision = "tf32" if self.tc_eligible else "fp32"
        self._tc = None       # packed tf32 weights of the tensor-core decoder (built lazily on the device)
        self._tc_enc = None   # ... and of the tensor-core encoder
        ech = [self.encoder_dim * 2 ** i for i in range(len(self.encoder_rates) + 1)] + [self.latent_dim]
        self.enc_tc_eligible = self._enc_ops is not None and all(c % 32 == 0 for c in ech)
        # the encoder decides discrete codes: its default stays fp32 (codes index-identical to the fp32 oracle except at
        # near-ties, tests/test_dac_gpu.py::_check_codes); "tf32" opts into the tensor-core encoder (latent rel-L2 2e-3 ...
        # 7e-3 on the synthetic encoder = its tf32 floor, so a few per cent of the codes may land on a neighbouring entry)
        self.encoder_precision = "fp32"

    def _apply(self, fn, *a, **k):
        self._tc = self._tc_enc = None
        return super()._apply(fn, *a, **k)

    @property
    def device(self):
        return self.codebooks.device

    def _require_cuda(self):
        if self.device.type != "cuda":
            raise RuntimeError("flowdec_b200.ndac runs on CUDA (sm_100a) only; call .to('cuda')")
        torch.cuda.set_device(self.device)     # ctypes launches go to the current device

    @classmethod
    def load(cls, location, *args, **kwargs):
        """audiotools.ml.BaseModel.load for package-less checkpoints: {"state_dict", "metadata": {"kwargs"}}"""
        try:
            ckpt = torch.load(str(location), map_location="cpu", weights_only=True)
        except Exception:      # checkpoints that pickle non-tensor objects next to the state_dict
            ckpt = torch.load(str(location), map_location="cpu", weights_only=False)
        kw = dict(ckpt["metadata"]["kwargs"])
        kw.update(kwargs)
        return cls(ckpt["state_dict"], **kw)

    # ---------------------------------------------------------------------------------------
    def _t(self, name):
        """buffer name (or None) -> the tensor currently registered under it (follows .to() / .cuda())"""
        return None if name is None else getattr(self, name)

    def op_tensors(self):
        """every weight / bias / alpha tensor the decoder and encoder op lists refer to, resolved now"""
        names = []
        for ops_list in (self._ops, self._enc_ops or []):
            for op in ops_list:
                for f in op[1:]:
                    for g in (f if isinstance(f, tuple) else (f,)):
                        if isinstance(g, str):
                            names.append(g)
        return {n: getattr(self, n) for n in names}

    def _conv(self, x, w, b, alpha, dil, pad, res=None, tanh=False):
        w, b, alpha = self._t(w), self._t(b), self._t(alpha)
        B, Cin, Tin = x.shape
        Cout, _, K = w.shape
        Tout = Tin + 2 * pad - dil * (K - 1)
        out = torch.empty(B, Cout, Tout, device=x.device, dtype=torch.float32)
        rc = _lib.lib().fd_dac_conv1d(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(alpha), _lib.ptr(res),
                                      _lib.ptr(out), B, Cin, Cout, Tin, K, dil, pad, int(tanh), _lib.stream_ptr())
        _lib.check(rc, "fd_dac_conv1d")
        return out

    def _run(self, ops_list, x):
        for op in ops_list:
            if op[0] == "conv":
                _, w, b, a, dil, pad, _, tanh = op
                x = self._conv(x, w, b, a, dil, pad, tanh=tanh)
            elif op[0] == "convtr":
                _, w, b, a, s, pad = op
                w, b, a = self._t(w), self._t(b), self._t(a)
                B, Cin, Tin = x.shape
                Cout = w.shape[1]
                Tout = (Tin - 1) * s - 2 * pad + 2 * s
                out = torch.empty(B, Cout, Tout, device=x.device, dtype=torch.float32)
                rc = _lib.lib().fd_dac_conv_transpose1d(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(a),
                                                        _lib.ptr(out), B, Cin, Cout, Tin, s, pad, _lib.stream_ptr())
                _lib.check(rc, "fd_dac_conv_transpose1d")
                x = out
            elif op[0] == "down":
                _, w, b, a, s, pad = op
                w, b, a = self._t(w), self._t(b), self._t(a)
                B, Cin, Tin = x.shape
                Cout, _, K = w.shape
                Tout = (Tin + 2 * pad - K) // s + 1
                out = torch.empty(B, Cout, Tout, device=x.device, dtype=torch.float32)
                rc = _lib.lib().fd_dac_conv1d_strided(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(a),
                                                      _lib.ptr(out), B, Cin, Cout, Tin, K, s, pad, _lib.stream_ptr())
                _lib.check(rc, "fd_dac_conv1d_strided")
                x = out
            else:
                _, (w1, b1, a1, d), (w2, b2, a2) = op
                y = self._conv(x, w1, b1, a1, d, 3 * d)
                x = self._conv(y, w2, b2, a2, 1, 0, res=x)
        return x

    # --------------------------------------------------------------------------------------- tensor-core decoder
    def _tc_prepare(self):
        """tf32-rounded GEMM weights of every decoder layer (see fd_dac_conv_tc in include/flowdec_b200.h)"""
        if self._tc is not None:
            return self._tc
        from .ops import round_tf32
        L = []

        def conv_w(w):            # [Cout, Cin, K] -> [Cout, K*Cin]
            return round_tf32(w.permute(0, 2, 1).reshape(w.shape[0], -1).contiguous())

        def convtr_w(w, s):       # [Cin, Cout, 2s] -> [s*Cout, 2*Cin]: row r*Cout+co, col tap*Cin+ci = w[ci,co,r+tap*s]
            cin, cout, _ = w.shape
            return round_tf32(w.reshape(cin, cout, 2, s).permute(3, 1, 2, 0).reshape(s * cout, 2 * cin).contiguous())

        for op in self._ops:
            if op[0] == "conv":
                _, w, b, a, dil, pad, _, tanh = op
                w, b, a = self._t(w), self._t(b), self._t(a)
                if tanh:
                    L.append(("final", w.reshape(w.shape[1], w.shape[2]).contiguous(), b, a))
                else:
                    L.append(("conv", conv_w(w), b, [(j - 3) * dil for j in range(w.shape[2])]))
            elif op[0] == "convtr":
                _, w, b, a, s, pad = op
                w, b, a = self._t(w), self._t(b), self._t(a)
                L.append(("convtr", convtr_w(w, s), b.repeat(s).contiguous(), a, s, pad, w.shape[1]))
            else:
                _, (w1, b1, a1, d), (w2, b2, a2) = op
                w1, b1, a1, w2, b2, a2 = (self._t(t) for t in (w1, b1, a1, w2, b2, a2))
                L.append(("res", conv_w(w1), b1, a1, [(j - 3) * d for j in range(7)], conv_w(w2), b2, a2))
        self._tc = L
        return L

    @staticmethod
    def _tc_conv(x, Tin, x_bs, Cin, wp, offs, bias, residual, res_bs, alpha, alpha_mod, want_raw, want_act, Tout):
        """one fd_dac_conv_tc launch; x / residual are (tensor, element offset) views with batch stride *_bs"""
        xt, xo = x
        B = xt.shape[0]
        Ntot = wp.shape[0]
        dev = xt.device
        raw = torch.empty(B, Tout, Ntot, device=dev, dtype=torch.float32) if want_raw else None
        act = torch.empty(B, Tout, Ntot, device=dev, dtype=torch.float32) if want_act else None
        offs_c = (ctypes.c_int * len(offs))(*offs)
        rp = ctypes.c_void_p(residual[0].data_ptr() + 4 * residual[1]) if residual is not None else ctypes.c_void_p(0)
        rc = _lib.lib().fd_dac_conv_tc(ctypes.c_void_p(xt.data_ptr() + 4 * xo), B, Tin, x_bs, Cin, _lib.ptr(wp), Ntot,
                                       len(offs), offs_c, _lib.ptr(bias), rp, res_bs, _lib.ptr(alpha), alpha_mod,
                                       _lib.ptr(raw), _lib.ptr(act), Tout, Tout * Ntot, _lib.stream_ptr())
        _lib.check(rc, "fd_dac_conv_tc")
        return raw, act

    def _decode_tc(self, z):
        """the decoder as a chain of tf32 implicit GEMMs on time-major activations; every epilogue applies the NEXT
        layer's Snake, so each layer reads an already activated tensor"""
        L = self._tc_prepare()
        B, D, T = z.shape
        dev = z.device
        zt = torch.empty(B, T, D, device=dev, dtype=torch.float32)
        _lib.check(_lib.lib().fd_dac_nct_to_ntc(_lib.ptr(z), _lib.ptr(zt), B, D, T, 1, _lib.stream_ptr()),
                   "fd_dac_nct_to_ntc")

        def next_alpha(i):
            """Snake alpha applied to the output of layer i = the first activation of layer i + 1"""
            nxt = L[i + 1]
            return nxt[3] if nxt[0] in ("convtr", "res", "final") else None

        assert L[0][0] == "conv" and L[-1][0] == "final"
        _, wp, b, offs = L[0]
        a = next_alpha(0)
        _, act = self._tc_conv((zt, 0), T, T * D, D, wp, offs, b, None, 0, a, a.numel(), False, True, T)
        cur_act, cur_raw = (act, 0), None          # (tensor, element offset) views
        Tcur, C, bs = T, wp.shape[0], T * wp.shape[0]
        for i in range(1, len(L) - 1):
            lay = L[i]
            a = next_alpha(i)
            last_of_block = L[i + 1][0] != "res"
            if lay[0] == "convtr":
                _, wp, b, _, s, pad, cout = lay
                raw, act = self._tc_conv(cur_act, Tcur, bs, C, wp, [0, -1], b, None, 0, a, cout, True, True, Tcur + 1)
                Tout = (Tcur - 1) * s - 2 * pad + 2 * s
                bs = (Tcur + 1) * s * cout
                cur_raw, cur_act = (raw, pad * cout), (act, pad * cout)
                Tcur, C = Tout, cout
            else:
                _, w1, b1, _, offs, w2, b2, a2 = lay
                _, h = self._tc_conv(cur_act, Tcur, bs, C, w1, offs, b1, None, 0, a2, C, False, True, Tcur)
                raw, act = self._tc_conv((h, 0), Tcur, Tcur * C, C, w2, [0], b2, cur_raw, bs, a, C,
                                         not last_of_block, True, Tcur)
                bs = Tcur * C
                cur_raw, cur_act = ((raw, 0) if raw is not None else None), (act, 0)
        _, wf, bf, _ = L[-1]
        out = torch.empty(B, 1, Tcur, device=dev, dtype=torch.float32)
        xt, xo = cur_act
        rc = _lib.lib().fd_dac_final_conv(ctypes.c_void_p(xt.data_ptr() + 4 * xo), bs, _lib.ptr(wf), _lib.ptr(bf),
                                          _lib.ptr(out), B, Tcur, C, _lib.stream_ptr())
        _lib.check(rc, "fd_dac_final_conv")
        return out

    # --------------------------------------------------------------------------------------- tensor-core encoder
    def _tc_enc_prepare(self):
        if self._tc_enc is not None:
            return self._tc_enc
        from .ops import round_tf32

        def conv_w(w):
            return round_tf32(w.permute(0, 2, 1).reshape(w.shape[0], -1).contiguous())

        def down_w(w, s, pad):    # [Cout, C, 2s] -> [Cout, 3*s*C]: col tap*(s*C) + j*C + ci = w[co, ci, (tap-1)*s + j + pad]
            cout, c, k2 = w.shape
            wp = torch.zeros(cout, 3, s, c, device=w.device, dtype=torch.float32)
            for tap in range(3):
                for j in range(s):
                    k = (tap - 1) * s + j + pad
                    if 0 <= k < k2:
                        wp[:, tap, j, :] = w[:, :, k]
            return round_tf32(wp.reshape(cout, 3 * s * c).contiguous())

        L = []
        for i, op in enumerate(self._enc_ops):
            if op[0] == "conv":
                _, w, b, a, dil, pad, _, _ = op
                w, b, a = self._t(w), self._t(b), self._t(a)
                if i == 0:
                    L.append(("first", w.reshape(w.shape[0], w.shape[2]).contiguous(), b, None))
                else:
                    L.append(("last", conv_w(w), b, a, [j - pad for j in range(w.shape[2])]))
            elif op[0] == "down":
                _, w, b, a, s, pad = op
                w, b, a = self._t(w), self._t(b), self._t(a)
                L.append(("down", down_w(w, s, pad), b, a, s))
            else:
                _, (w1, b1, a1, d), (w2, b2, a2) = op
                w1, b1, a1, w2, b2, a2 = (self._t(t) for t in (w1, b1, a1, w2, b2, a2))
                L.append(("res", conv_w(w1), b1, a1, [(j - 3) * d for j in range(7)], conv_w(w2), b2, a2))
        self._tc_enc = L
        return L

    def _encode_tc(self, x):
        """Encoder on tensor cores: x [B, 1, T] -> latent z [B, D, T / hop] (same GEMM kernel as the decoder; the
        strided convs read the activated tensor as [B, T/s, s*C], a free view of the time-major layout)"""
        L = self._tc_enc_prepare()
        B, _, T = x.shape
        dev = x.device
        f32 = dict(device=dev, dtype=torch.float32)
        _, w0, b0, _ = L[0]
        C = w0.shape[0]
        raw = torch.empty(B, T, C, **f32)
        act = torch.empty(B, T, C, **f32)
        _lib.check(_lib.lib().fd_dac_first_conv(_lib.ptr(x), _lib.ptr(w0), _lib.ptr(b0), _lib.ptr(L[1][3]), _lib.ptr(raw),
                                                _lib.ptr(act), B, T, C, _lib.stream_ptr()), "fd_dac_first_conv")
        for i in range(1, len(L)):
            lay = L[i]
            nxt = L[i + 1] if i + 1 < len(L) else None
            a_next = nxt[3] if nxt is not None else None
            need_raw = nxt is not None and nxt[0] == "res"
            if lay[0] == "res":
                _, w1, b1, _, offs, w2, b2, a2 = lay
                _, h = self._tc_conv((act, 0), T, T * C, C, w1, offs, b1, None, 0, a2, C, False, True, T)
                raw, act = self._tc_conv((h, 0), T, T * C, C, w2, [0], b2, (raw, 0), T * C, a_next, C, need_raw, True, T)
            elif lay[0] == "down":
                _, wp, b, _, s = lay
                if T % s:
                    raise ValueError(f"encoder input length must be a multiple of the hop ({self.hop_length}); call preprocess()")
                cout = wp.shape[0]
                raw, act = self._tc_conv((act, 0), T // s, T * C, s * C, wp, [-1, 0, 1], b, None, 0, a_next, cout,
                                         need_raw, True, T // s)
                T, C = T // s, cout
            else:
                _, wp, b, _, offs = lay
                zt, _ = self._tc_conv((act, 0), T, T * C, C, wp, offs, b, None, 0, None, 0, True, False, T)
                D = wp.shape[0]
                z = torch.empty(B, D, T, **f32)
                # [B, T, D] -> [B, D, T]: the same transpose kernel with the roles of C and T swapped, no rounding
                _lib.check(_lib.lib().fd_dac_nct_to_ntc(_lib.ptr(zt), _lib.ptr(z), B, T, D, 0, _lib.stream_ptr()),
                           "fd_dac_nct_to_ntc")
                return z
        raise RuntimeError("encoder layer list has no final conv")

    @torch.no_grad()
    def decode(self, z):
        """z [B, latent_dim, T] -> waveform [B, 1, ~T*hop] (dac.DAC.decode, demo.ipynb:105)"""
        self._require_cuda()
        z = z.to(self.device, torch.float32).contiguous()
        if self.precision == "tf32":
            if not self.tc_eligible:
                raise RuntimeError("precision 'tf32' needs decoder channel counts that are multiples of 32")
            return self._decode_tc(z)
        return self._run(self._ops, z)

    def preprocess(self, audio_data, sample_rate=None):
        """dac.DAC.preprocess (demo.ipynb:101): zero-pad on the right to a multiple of hop_length.
        Pure data movement (no arithmetic)."""
        if sample_rate is None:
            sample_rate = self.sample_rate
        assert sample_rate == self.sample_rate, f"expected {self.sample_rate} Hz audio, got {sample_rate}"
        L = audio_data.shape[-1]
        right = math.ceil(L / self.hop_length) * self.hop_length - L
        if right == 0:
            return audio_data
        out = audio_data.new_zeros(*audio_data.shape[:-1], L + right)
        out[..., :L].copy_(audio_data)
        return out

    @torch.no_grad()
    def encode(self, audio_data, n_quantizers=None):
        """dac.DAC.encode (demo.ipynb:102): audio [B, 1, L] (L a multiple of hop_length) ->
        (z_q [B,D,T], codes [B,n_q,T], latents [B,n_q*cdim,T], commitment_loss, codebook_loss)"""
        self._require_cuda()
        if self._enc_ops is None:
            raise RuntimeError("this checkpoint has no encoder weights (decode-only state_dict)")
        x = audio_data.to(self.device, torch.float32).contiguous()
        if x.ndim != 3 or x.shape[1] != 1:
            raise ValueError(f"audio_data must be [B, 1, L], got {tuple(x.shape)}")
        if self.encoder_precision == "tf32":
            if not self.enc_tc_eligible:
                raise RuntimeError("encoder_precision 'tf32' needs encoder channel counts that are multiples of 32")
            return self.quantizer(self._encode_tc(x), n_quantizers)
        return self.quantizer(self._run(self._enc_ops, x), n_quantizers)

    @torch.no_grad()
    def forward(self, audio_data, sample_rate=None, n_quantizers=None):
        """dac.DAC.forward: preprocess -> encode -> decode, audio cropped back to the input length"""
        L = audio_data.shape[-1]
        z, codes, latents, commit, cbl = self.encode(self.preprocess(audio_data, sample_rate), n_quantizers)
        return {"audio": self.decode(z)[..., :L], "z": z, "codes": codes, "latents": latents,
                "vq/commitment_loss": commit, "vq/codebook_loss": cbl}
