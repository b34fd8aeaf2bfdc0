"""NCSN++ score/velocity backbone of FlowDec, B200-native execution.

Drop-in for the reference class `flowdec.backbones.ncsnpp.NCSNpp`
(/root/reference/flowdec/backbones/ncsnpp.py:52-411): same constructor keywords, same
`all_modules.{i}.*` / `output_layer.weight` parameter names and shapes (so
`load_state_dict(ckpt['_pl_ema_state_dict'])` works unchanged), same
`forward(x, y, t) -> complex64 [B,1,F,T]` contract.

The nn.Modules below only *hold parameters*.  `forward` never calls their `forward`: it walks
the network issuing the hand-written sm_100a kernels of libflowdec_b200.so through
`flowdec_b200.ops` (tcgen05 implicit-GEMM convolutions + HBM-bound GroupNorm/SiLU/FIR
passes).  Supported configuration family = what the reference's configs instantiate:
biggan res-blocks, FIR resampling, progressive output_skip / input_skip with 'sum' combine,
Fourier embedding, any nf / ch_mult / num_res_blocks, optional bottleneck attention and a 1x1 or 3x3
bias-free output layer — i.e. config/model/backbone/ncsnpp_final_no_attn.yaml (FlowDec / ScoreDec) and
ncsnpp_default_ycond.yaml (the SGMSE-style 7-level baseline).  Convolutions whose shape
fits the tcgen05 tiles (channel segments % 64, Cout 128/256, W % 8, H % tile height) run on tensor cores;
the rest (the 24x8 ... 12x1 levels of the 7-level net, narrow test networks) run on the shape-generic
CUDA-core kernels of csrc/fd_generic.cu.  Anything else raises NotImplementedError rather than silently
running a different path.
"""
import ctypes
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from .. import ops

DEFAULT_OUTPUTLAYER_KWARGS = dict(kernel_size=3, bias=False, padding="same", padding_mode="zeros")


def _ddpm_uniform_(w, scale=1.0):
    """variance_scaling(scale, 'fan_avg', 'uniform') of reference layers.py:64-101."""
    scale = 1e-10 if scale == 0 else scale
    shape = w.shape
    rf = float(np.prod(shape)) / shape[1] / shape[0]
    fan_in, fan_out = shape[1] * rf, shape[0] * rf
    lim = math.sqrt(3 * scale / ((fan_in + fan_out) / 2))
    with torch.no_grad():
        w.uniform_(-lim, lim)
    return w


def _conv(cin, cout, k, init_scale=1.0, bias=True):
    c = nn.Conv2d(cin, cout, kernel_size=k, padding=k // 2, bias=bias)
    _ddpm_uniform_(c.weight, init_scale)
    if bias:
        nn.init.zeros_(c.bias)
    return c


def _dense(cin, cout):
    d = nn.Linear(cin, cout)
    _ddpm_uniform_(d.weight)
    nn.init.zeros_(d.bias)
    return d


class GaussianFourierProjection(nn.Module):
    """parameter holder for reference layerspp.py:42-51"""

    def __init__(self, embedding_size=256, scale=1.0):
        super().__init__()
        self.W = nn.Parameter(torch.randn(embedding_size) * scale, requires_grad=False)


class ResnetBlockBigGANpp(nn.Module):
    """parameter holder for reference layerspp.py:222-284"""

    def __init__(self, in_ch, out_ch=None, temb_dim=None, up=False, down=False, init_scale=0.0):
        super().__init__()
        out_ch = out_ch if out_ch else in_ch
        self.GroupNorm_0 = nn.GroupNorm(min(in_ch // 4, 32), in_ch, eps=1e-6)
        self.Conv_0 = _conv(in_ch, out_ch, 3)
        if temb_dim is not None:
            self.Dense_0 = _dense(temb_dim, out_ch)
        self.GroupNorm_1 = nn.GroupNorm(min(out_ch // 4, 32), out_ch, eps=1e-6)
        self.Conv_1 = _conv(out_ch, out_ch, 3, init_scale=init_scale)
        if in_ch != out_ch or up or down:
            self.Conv_2 = _conv(in_ch, out_ch, 1)
        self.up, self.down, self.in_ch, self.out_ch = up, down, in_ch, out_ch


class Combine(nn.Module):
    """parameter holder for reference layerspp.py:54-69 (method='sum')"""

    def __init__(self, dim1, dim2):
        super().__init__()
        self.Conv_0 = _conv(dim1, dim2, 1)


class NIN(nn.Module):
    """parameter holder for reference layers.py:566-575 (y = x . W + b over channels)"""

    def __init__(self, in_dim, num_units, init_scale=0.1):
        super().__init__()
        self.W = nn.Parameter(_ddpm_uniform_(torch.empty(in_dim, num_units), init_scale))
        self.b = nn.Parameter(torch.zeros(num_units))


class AttnBlockpp(nn.Module):
    """parameter holder for reference layerspp.py:72-101 (skip_rescale=True)"""

    def __init__(self, channels, init_scale=0.0):
        super().__init__()
        self.GroupNorm_0 = nn.GroupNorm(min(channels // 4, 32), channels, eps=1e-6)
        self.NIN_0 = NIN(channels, channels)
        self.NIN_1 = NIN(channels, channels)
        self.NIN_2 = NIN(channels, channels)
        self.NIN_3 = NIN(channels, channels, init_scale=init_scale)
        self.channels = channels


class _Workspace:
    """Named device buffers with stable addresses (CUDA-graph friendly, no per-call allocation)."""

    def __init__(self):
        self.bufs = {}

    def get(self, name, shape, dtype, device):
        key = (name, tuple(shape), dtype)
        t = self.bufs.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=device)
            self.bufs[key] = t
        return t

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


class NCSNpp(nn.Module):
    """NCSN++ model (FlowDec configuration family) running on hand-written sm_100a kernels."""

    def __init__(self,
                 nonlinearity="swish", nf=128, ch_mult=(1, 1, 2, 2, 2, 2, 2), num_res_blocks=2,
                 attn_resolutions=(64, 32, 16, 8), resamp_with_conv=True, conditional=True, fir=True,
                 fir_kernel=(1, 3, 3, 1), skip_rescale=True, resblock_type="biggan",
                 progressive="output_skip", progressive_input="input_skip", progressive_combine="sum",
                 init_scale=0.0, fourier_scale=16, image_size=256, embedding_type="fourier", dropout=0.0,
                 num_channels=4, output_layer_kwargs: dict = DEFAULT_OUTPUTLAYER_KWARGS,
                 bottleneck_attn: bool = True):
        super().__init__()
        ch_mult = list(ch_mult)
        all_res = [image_size // (2 ** i) for i in range(len(ch_mult))]
        unsupported = []
        if nonlinearity != "swish": unsupported.append(f"nonlinearity={nonlinearity}")
        if any(r in list(attn_resolutions) for r in all_res): unsupported.append("attn_resolutions")
        if not conditional: unsupported.append("conditional=False")
        if not fir or list(fir_kernel) != [1, 3, 3, 1]: unsupported.append("fir/fir_kernel")
        if not skip_rescale: unsupported.append("skip_rescale=False")
        if resblock_type.lower() != "biggan": unsupported.append(f"resblock_type={resblock_type}")
        if progressive.lower() != "output_skip": unsupported.append(f"progressive={progressive}")
        if progressive_input.lower() != "input_skip": unsupported.append(f"progressive_input={progressive_input}")
        if progressive_combine.lower() != "sum": unsupported.append(f"progressive_combine={progressive_combine}")
        if embedding_type.lower() != "fourier": unsupported.append(f"embedding_type={embedding_type}")
        if dropout != 0.0: unsupported.append("dropout")
        if num_channels != 4: unsupported.append("num_channels")
        olk = dict(output_layer_kwargs)
        if (olk.get("kernel_size", 3) not in (1, 3) or olk.get("bias", False)
                or olk.get("padding", "same") != "same" or olk.get("padding_mode", "zeros") != "zeros"):
            unsupported.append("output_layer_kwargs (need kernel_size 1 or 3, no bias, zero 'same' padding)")
        if nf % 8 or any((nf * m) % 8 for m in ch_mult): unsupported.append("channel counts must be multiples of 8")
        if unsupported:
            raise NotImplementedError(
                "flowdec_b200.NCSNpp implements the configuration family of the reference's configs "
                "(ncsnpp_final_no_attn.yaml, ncsnpp_default_ycond.yaml); unsupported: " + ", ".join(unsupported))

        self.nf, self.ch_mult, self.num_res_blocks = nf, ch_mult, num_res_blocks
        self.num_resolutions = len(ch_mult)
        self.image_size = image_size
        self.bottleneck_attn = bool(bottleneck_attn)
        self.out_k = int(olk.get("kernel_size", 3))
        self.output_layer = nn.Conv2d(num_channels, 2, kernel_size=self.out_k, padding=self.out_k // 2, bias=False)

        # --- module list in the reference's construction order (ncsnpp.py:102-252) ---
        mods = [GaussianFourierProjection(embedding_size=nf, scale=fourier_scale),
                _dense(2 * nf, 4 * nf), _dense(4 * nf, 4 * nf), _conv(num_channels, nf, 3)]
        RB = lambda **kw: ResnetBlockBigGANpp(temb_dim=4 * nf, init_scale=init_scale, **kw)
        hs_c = [nf]
        in_ch = nf
        for lvl in range(self.num_resolutions):
            for _ in range(num_res_blocks):
                out_ch = nf * ch_mult[lvl]
                mods.append(RB(in_ch=in_ch, out_ch=out_ch))
                in_ch = out_ch
                hs_c.append(in_ch)
            if lvl != self.num_resolutions - 1:
                mods.append(RB(down=True, in_ch=in_ch))
                mods.append(Combine(dim1=num_channels, dim2=in_ch))
                hs_c.append(in_ch)
        in_ch = hs_c[-1]
        mods.append(RB(in_ch=in_ch))
        if self.bottleneck_attn:
            mods.append(AttnBlockpp(in_ch, init_scale=init_scale))
        mods.append(RB(in_ch=in_ch))
        for lvl in reversed(range(self.num_resolutions)):
            for _ in range(num_res_blocks + 1):
                out_ch = nf * ch_mult[lvl]
                mods.append(RB(in_ch=in_ch + hs_c.pop(), out_ch=out_ch))
                in_ch = out_ch
            mods.append(nn.GroupNorm(min(in_ch // 4, 32), in_ch, eps=1e-6))
            mods.append(_conv(in_ch, num_channels, 3, init_scale=init_scale))
            if lvl != 0:
                mods.append(RB(in_ch=in_ch, up=True))
        assert not hs_c
        self.all_modules = nn.ModuleList(mods)

        self._prepared = None      # packed weights (built lazily, dropped on load_state_dict / .to())
        self._temb_cache = {}
        self.generation = 0        # bumped whenever packed weights / workspaces are invalidated
        # (lane, B, F, T) -> _Workspace, least recently used first; signatures in `_pinned` (those of live CUDA
        # graphs, set by FlowModel) are never evicted, others beyond `max_workspaces` are
        self._workspaces = OrderedDict()
        self._pins, self._pinned = {}, set()
        self.max_workspaces = 4
        self._ws_cur = None
        self.stats_slabs = 64
        self.fuse_stats = True     # GroupNorm partial sums produced by the conv epilogue
        self.pyramid_shift_after_gemm = True
        # pyramid convs on the halo kernel (1-tap N = 48 form with the fused transform): measured SLOWER than the
        # materialised pass + per-tap GEMM (r2b: 1.52 + 0.23 ms vs 0.75 + 0.45 + 0.23 ms per forward of 8 clips) because the
        # exposed transform chain, not the tensor pipe, paces a kernel with so few MMAs; used by the tf32 mode only
        self.pyramid_halo = False
        self.fuse_gn_into_conv = True   # GroupNorm+SiLU applied inside the conv kernel (needs ops.HALO_TILES)
        self.max_ctas = 0
        self.precision = "bf16"

    def set_precision(self, precision):
        """"bf16" (default): bf16 operands / activation storage, fp32 accumulation.  "tf32": fp32 activation
        storage, tf32 tensor-core operands (kind::tf32) — the precision class of the reference's own GPU convs
        (cuDNN with allow_tf32, layers.py:110-134); about half the conv throughput and twice the activation bytes.
        The tf32 kernels exist for shapes the halo tiles take (the FlowDec configuration family)."""
        if precision not in ("bf16", "tf32"):
            raise ValueError(f"precision must be 'bf16' or 'tf32', got {precision!r}")
        if precision != self.precision:
            self.precision = precision
            self._invalidate()
            self._workspaces = OrderedDict()
        return self

    @property
    def _adt(self):
        return torch.float32 if self.precision == "tf32" else torch.bfloat16

    def _pack(self, segs, rows):
        return ops.pack_conv_weight(segs, rows, tf32=self.precision == "tf32")

    # ------------------------------------------------------------------ parameter management
    def _invalidate(self):
        """packed weights / caches / workspaces are stale: owners of captured graphs (FlowModel, ScoreModel — several
        may share this backbone) compare `generation` and drop their graphs"""
        self._prepared = None
        self._temb_cache = {}
        self.generation = getattr(self, "generation", 0) + 1

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._invalidate()
        return r

    def _load_from_state_dict(self, *a, **k):
        self._invalidate()
        return super()._load_from_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        # .cuda() / .to() on a module that is already there (e.g. a second model built around this backbone) must
        # not invalidate buffers that captured graphs replay into: only a real move / cast does
        w = self.output_layer.weight
        before = (w.data_ptr(), w.device, w.dtype)
        r = super()._apply(fn, *a, **k)
        w = self.output_layer.weight
        if (w.data_ptr(), w.device, w.dtype) != before:
            self._invalidate()
            self._workspaces = OrderedDict()
        return r

    # ------------------------------------------------------------------ workspaces
    def pin_workspaces(self, sigs, owner="default"):
        """sigs: set of (micro-batch, F, T) whose workspaces must keep their addresses (captured CUDA graphs of
        `owner` replay into them; several models may share one backbone); everything unpinned becomes evictable"""
        self._pins[owner] = set(sigs)
        self._pinned = set().union(*self._pins.values())
        self._evict()

    def _evict(self):
        free = [k for k in self._workspaces if k[1:] not in self._pinned and k != self._ws_cur]
        while len(self._workspaces) > self.max_workspaces and free:
            del self._workspaces[free.pop(0)]

    def _select_ws(self, lane, B, Fq, T):
        key = (lane, B, Fq, T)
        self._ws_cur = key
        w = self._workspaces.get(key)
        if w is None:
            w = self._workspaces[key] = _Workspace()
            self._evict()
        else:
            self._workspaces.move_to_end(key)
        self._ws = w
        return w

    def workspace_bytes(self):
        return sum(w.nbytes() for w in self._workspaces.values())

    @staticmethod
    def _npad(cout):
        """rows of the packed weight: the tcgen05 tiles take 128 / 256 output channels, every other width runs
        on the shape-generic kernel, which needs no padding"""
        return cout

    # ------------------------------------------------------------------ structure walk
    def _plan(self):
        """Static walk of the network on channel counts only: for every res-block the channel split of its
        (virtual concat) input.  Lets prepare() pack every fused conv1+skip and multi-source conv0 weight up
        front, so no packing kernel runs inside velocity() (where lanes run on separate streams)."""
        mods = self.all_modules
        plan, idx = {}, 4
        hs = [self.nf]
        for lvl in range(self.num_resolutions):
            for _ in range(self.num_res_blocks):
                plan[idx] = [hs[-1]]
                hs.append(mods[idx].out_ch)
                idx += 1
            if lvl != self.num_resolutions - 1:
                plan[idx] = [hs[-1]]
                idx += 2                      # down res-block + Combine
                hs.append(mods[idx - 2].out_ch)
        h = hs[-1]
        plan[idx] = [h]
        h = mods[idx].out_ch
        idx += 1
        if self.bottleneck_attn:
            idx += 1
        plan[idx] = [h]
        h = mods[idx].out_ch
        idx += 1
        for lvl in reversed(range(self.num_resolutions)):
            for _ in range(self.num_res_blocks + 1):
                plan[idx] = [h, hs.pop()]
                h = mods[idx].out_ch
                idx += 1
            idx += 2                          # GroupNorm + pyramid conv
            if lvl != 0:
                plan[idx] = [h]
                h = mods[idx].out_ch
                idx += 1
        assert not hs and idx == len(mods)
        return plan

    def prepare(self):
        """Repack weights for the kernels (OIHW fp32 -> K-major bf16 [rows, Ktot], fused skips)."""
        if self._prepared is not None:
            return self._prepared
        dev = self.output_layer.weight.device
        if dev.type != "cuda":
            raise RuntimeError("flowdec_b200.NCSNpp runs on CUDA (sm_100a) only; call .cuda() first")
        P = {}
        inv = 1.0 / math.sqrt(2.0)
        plan = self._plan()
        with torch.no_grad():
            for i, m in enumerate(self.all_modules):
                if isinstance(m, ResnetBlockBigGANpp):
                    e = {}
                    cin, cout = m.in_ch, m.out_ch
                    e["w0"] = self._pack([(m.Conv_0.weight.float(), 9)], self._npad(cout))
                    e["w0_full"] = m.Conv_0.weight.float()
                    e["b0"] = m.Conv_0.bias.float().contiguous()
                    w1 = m.Conv_1.weight.float() * inv
                    if hasattr(m, "Conv_2"):
                        w2 = m.Conv_2.weight.float() * inv       # [cout, cin, 1, 1]
                        e["w2_full"] = w2
                        e["b1"] = ((m.Conv_1.bias + m.Conv_2.bias).float() * inv).contiguous()
                    else:
                        w2 = (torch.eye(cout, device=dev) * inv).reshape(cout, cout, 1, 1)
                        e["w2_full"] = w2
                        e["b1"] = (m.Conv_1.bias.float() * inv).contiguous()
                    e["w1_conv"] = w1
                    e["w1_cache"] = {}
                    e["g0"], e["be0"] = m.GroupNorm_0.weight.float().contiguous(), m.GroupNorm_0.bias.float().contiguous()
                    e["g1"], e["be1"] = m.GroupNorm_1.weight.float().contiguous(), m.GroupNorm_1.bias.float().contiguous()
                    e["dw"], e["db"] = m.Dense_0.weight.float().contiguous(), m.Dense_0.bias.float().contiguous()
                    P[i] = e
                    # every variant velocity() will ask for (the splits are static)
                    split = plan[i]
                    self._skip_weight(e, [cin] if (m.up or m.down) else split, cout)
                    if not (m.up or m.down):
                        self._conv0_weight(e, split, cout)
                elif isinstance(m, AttnBlockpp):
                    C = m.channels
                    wq = torch.cat([n.W.float().t() for n in (m.NIN_0, m.NIN_1, m.NIN_2)], 0).reshape(3 * C, C, 1, 1)
                    eye = (torch.eye(C, device=dev) * inv).reshape(C, C, 1, 1)
                    P[i] = dict(g=m.GroupNorm_0.weight.float().contiguous(), be=m.GroupNorm_0.bias.float().contiguous(),
                                wqkv=ops.pack_conv_weight([(wq, 1)], 3 * C),
                                bqkv=torch.cat([n.b.float() for n in (m.NIN_0, m.NIN_1, m.NIN_2)]).contiguous(),
                                wo=ops.pack_conv_weight([((m.NIN_3.W.float().t() * inv).reshape(C, C, 1, 1), 1),
                                                         (eye, 1)], C),
                                bo=(m.NIN_3.b.float() * inv).contiguous())
                elif isinstance(m, Combine):
                    P[i] = dict(w=m.Conv_0.weight.float().reshape(m.Conv_0.weight.shape[0], 4).contiguous(),
                                b=m.Conv_0.bias.float().contiguous())
                elif isinstance(m, nn.GroupNorm):
                    P[i] = dict(g=m.weight.float().contiguous(), b=m.bias.float().contiguous())
                elif isinstance(m, nn.Conv2d) and i > 3:     # pyramid conv C -> 4
                    b16 = torch.zeros(16, device=dev)
                    b16[:m.bias.shape[0]] = m.bias.float()
                    P[i] = dict(w=self._pack([(m.weight.float(), 9)], 16), b=b16,
                                wt=ops.pack_tap_weight(m.weight.float(), tf32=self.precision == "tf32"),
                                b4=m.bias.float().contiguous())
            P["conv_in_w"] = self.all_modules[3].weight.float().contiguous()
            P["conv_in_b"] = self.all_modules[3].bias.float().contiguous()
            P["Wf"] = self.all_modules[0].W.float().contiguous()
            P["l1w"], P["l1b"] = self.all_modules[1].weight.float().contiguous(), self.all_modules[1].bias.float().contiguous()
            P["l2w"], P["l2b"] = self.all_modules[2].weight.float().contiguous(), self.all_modules[2].bias.float().contiguous()
            if self.out_k == 1:
                wo = self.output_layer.weight.float().reshape(2, 4).cpu().contiguous()
                P["w_out8"] = (ctypes.c_float * 8)(*wo.flatten().tolist())
            else:
                P["w_out72"] = self.output_layer.weight.float().contiguous()      # [2,4,3,3] on the device
        self._prepared = P
        return P

    def _skip_weight(self, e, seg_channels, cout):
        """packed [rows, 9*cout + sum(seg)] weight of conv1 + 1x1 skip over the given source split"""
        key = tuple(seg_channels)
        wp = e["w1_cache"].get(key)
        if wp is None:
            segs = [(e["w1_conv"], 9)]
            c0 = 0
            for c in seg_channels:
                segs.append((e["w2_full"][:, c0:c0 + c].contiguous(), 1))
                c0 += c
            assert c0 == e["w2_full"].shape[1]
            wp = self._pack(segs, self._npad(cout))
            e["w1_cache"][key] = wp
        return wp

    def _conv0_weight(self, e, seg_channels, cout):
        """Conv_0 packed for a multi-source (virtual concat) operand: K = segment > tap > channel"""
        if len(seg_channels) == 1:
            return e["w0"]
        key = ("w0",) + tuple(seg_channels)
        wp = e["w1_cache"].get(key)
        if wp is None:
            segs, c0 = [], 0
            for c in seg_channels:
                segs.append((e["w0_full"][:, c0:c0 + c].contiguous(), 9))
                c0 += c
            wp = self._pack(segs, self._npad(cout))
            e["w1_cache"][key] = wp
        return wp

    # ------------------------------------------------------------------ time embedding
    def temb_biases(self, t):
        """Per-res-block conv0 bias vectors b0 + Dense_0(SiLU(temb(t))) (ncsnpp.py:263-274,
        layerspp.py:270-272).  temb has batch 1 at inference (scalar t) -> depends only on t."""
        key = float(t)
        hit = self._temb_cache.get(key)
        if hit is not None:
            return hit
        P = self.prepare()
        dev = P["Wf"].device
        nf = self.nf
        four = torch.empty(2 * nf, device=dev)
        h1 = torch.empty(4 * nf, device=dev)
        temb = torch.empty(4 * nf, device=dev)
        ops.fourier_embed(key, P["Wf"], four)
        ops.matvec(four, P["l1w"], P["l1b"], h1)
        ops.matvec(h1, P["l2w"], P["l2b"], temb, silu_in=True)
        out = {}
        for i, m in enumerate(self.all_modules):
            if isinstance(m, ResnetBlockBigGANpp):
                e = P[i]
                b = torch.zeros(self._npad(m.out_ch), device=dev)
                ops.matvec(temb, e["dw"], e["db"], b, silu_in=True, add=e["b0"])
                out[i] = b
        if len(self._temb_cache) >= 256:      # e.g. a long run of N=50 Euler calls with changing grids
            self._temb_cache.pop(next(iter(self._temb_cache)))
        self._temb_cache[key] = out
        return out

    # ------------------------------------------------------------------ building blocks
    def _partials(self, x, cache):
        k = x.data_ptr()
        p = cache.get(k)
        if p is None:
            B = x.shape[0]
            self._stat_counter += 1
            p = self._ws.get(f"stats{self._stat_counter}", (B, self.stats_slabs, x.shape[3], 2),
                             torch.float32, x.device)
            ops.chan_stats(x, self.stats_slabs, out=p)
            cache[k] = p
        return p

    def _gn_scale_shift(self, srcs, gamma, beta, scache, name):
        """GroupNorm statistics of the virtual concat -> per-(sample, channel) scale/shift [B, C, 2]."""
        B, H, W = srcs[0].shape[:3]
        C = sum(s.shape[3] for s in srcs)
        parts = [self._partials(s, scache) for s in srcs]
        ss = self._ws.get(name, (B, C, 2), torch.float32, srcs[0].device)
        return ops.gn_finalize(parts, [s.shape[3] for s in srcs], H * W, gamma, beta, min(C // 4, 32), 1e-6, ss)

    def _gn_act(self, srcs, gamma, beta, mode, scache, name, raw_name=None):
        """a = FIR?(SiLU(GroupNorm(cat(srcs)))) as one bf16 NHWC tensor (+ FIR(cat(srcs)) if raw_name)."""
        B, H, W = srcs[0].shape[:3]
        C = sum(s.shape[3] for s in srcs)
        ss = self._gn_scale_shift(srcs, gamma, beta, scache, "gn_ss")
        Ho, Wo = (H // 2, W // 2) if mode == 1 else ((H * 2, W * 2) if mode == 2 else (H, W))
        a = self._ws.get(name, (B, Ho, Wo, C), self._adt, srcs[0].device)
        raw = self._ws.get(raw_name, (B, Ho, Wo, C), self._adt, srcs[0].device) if raw_name else None
        ops.gn_act_resample(srcs, ss, a, mode, out_raw=raw)
        return (a, raw) if raw_name else a

    def _conv(self, srcs, wp, bias, out, stats_name=None, algo_k=None, scache=None):
        """One convolution over the virtual concat `srcs` -> `out` (bf16 NHWC): tcgen05 tiles when the shape
        fits, the shape-generic kernel otherwise.  With `stats_name` the tensor-core epilogue also emits the
        GroupNorm partial sums of the output (registered in `scache`); the generic path leaves them to an
        on-demand fd_chan_stats pass."""
        B, H, W, cout = out.shape
        if scache is not None:
            scache.pop(out.data_ptr(), None)
        if self.precision == "tf32" and not ops.halo_eligible(B, H, W, wp.shape[0]):
            raise NotImplementedError(f"tf32 precision: conv at {B}x{H}x{W} -> {cout} does not fit the halo tiles")
        if ops.tensor_conv_ok(B, H, W, wp.shape[0], [s[2] for s in srcs]) and \
                (all(len(s) <= 4 or s[4] is None for s in srcs) or ops.halo_eligible(B, H, W, wp.shape[0])):
            st = None
            if self.fuse_stats and stats_name is not None:
                # one slab per conv tile; fd_gn_finalize reduces them directly (no compaction pass)
                st = self._ws.get(stats_name, (B, ops.conv_stats_slabs(H, W), cout, 2), torch.float32, out.device)
            ops.conv_igemm(srcs, wp, bias, out, self.max_ctas, algo_k=algo_k, stats=st)
            if st is not None:
                scache[out.data_ptr()] = st
        else:
            ops.conv_direct(srcs, wp, bias, out)
        return out

    def _resblock(self, i, srcs, tb, scache, want_stats=True):
        """reference layerspp.py:252-284; srcs = virtual concat of bf16 NHWC tensors."""
        m = self.all_modules[i]
        e = self._prepared[i]
        dev = srcs[0].device
        B, H, W = srcs[0].shape[:3]
        mode = 1 if m.down else (2 if m.up else 0)
        Ho, Wo = (H // 2, W // 2) if mode == 1 else ((H * 2, W * 2) if mode == 2 else (H, W))
        cin, cout = m.in_ch, m.out_ch
        seg = [s.shape[3] for s in srcs]
        assert sum(seg) == cin
        # GroupNorm+SiLU applied inside the conv (halo kernel's operand transform, or on load in the generic
        # kernel) when possible; otherwise (and for the FIR-resampled conv0 input) a materialised bf16 tensor
        tensor0 = ops.tensor_conv_ok(B, Ho, Wo, self._npad(cout), seg if mode == 0 else [cin])
        tensor1 = ops.tensor_conv_ok(B, Ho, Wo, self._npad(cout), [cout] + (seg if mode == 0 else [cin]))
        halo = self.fuse_gn_into_conv and ops.halo_eligible(B, Ho, Wo, self._npad(cout))
        if mode != 0:
            a0, xr = self._gn_act(srcs, e["g0"], e["be0"], mode, scache, "act", raw_name="xr")
            conv0_srcs, w0 = [(a0, 0, cin, 9)], e["w0"]
        elif halo or not tensor0:
            ss0 = self._gn_scale_shift(srcs, e["g0"], e["be0"], scache, "gn_ss0")
            conv0_srcs, off = [], 0
            for sx in srcs:
                conv0_srcs.append((sx, 0, sx.shape[3], 9, ss0, off))
                off += sx.shape[3]
            w0 = self._conv0_weight(e, seg, cout)
        else:
            a0 = self._gn_act(srcs, e["g0"], e["be0"], mode, scache, "act")
            conv0_srcs, w0 = [(a0, 0, cin, 9)], e["w0"]
        h1 = self._ws.get("h1", (B, Ho, Wo, cout), self._adt, dev)
        self._conv(conv0_srcs, w0, tb[i], h1, stats_name="h1_stats_c", scache=scache)
        if halo or not tensor1:
            ss1 = self._gn_scale_shift([h1], e["g1"], e["be1"], scache, "gn_ss1")
            conv1_main = (h1, 0, cout, 9, ss1, 0)
        else:
            conv1_main = (self._gn_act([h1], e["g1"], e["be1"], 0, scache, "act"), 0, cout, 9)
        scache.pop(h1.data_ptr(), None)
        skip = [xr] if mode != 0 else list(srcs)
        wp = self._skip_weight(e, [s.shape[3] for s in skip], cout)
        out = self._ws.get(f"rb{i}", (B, Ho, Wo, cout), self._adt, dev)
        algo_k = 9 * cout + (cin if hasattr(m, "Conv_2") else 0)
        self._conv([conv1_main] + [(s, 0, s.shape[3], 1) for s in skip], wp, e["b1"], out,
                   stats_name=f"rb{i}_stats_c" if want_stats else None, algo_k=algo_k, scache=scache)
        return out

    def _attn(self, i, h, scache):
        """reference layerspp.py:72-101 (AttnBlockpp, skip_rescale): GroupNorm -> q, k, v (NIN) -> softmax over
        all H*W positions -> NIN -> (x + h)/sqrt(2); the residual rides as an identity K segment of the last conv"""
        e = self._prepared[i]
        B, H, W, C = h.shape
        ss = self._gn_scale_shift([h], e["g"], e["be"], scache, "attn_ss")
        qkv = self._ws.get("attn_qkv", (B, H, W, 3 * C), torch.float32, h.device)
        ops.conv_direct([(h, 0, C, 1, ss, 0)], e["wqkv"], e["bqkv"], qkv, affine_only=True)
        a = ops.attention(qkv, self._ws.get("attn_a", (B, H, W, C), self._adt, h.device))
        out = self._ws.get(f"attn{i}", (B, H, W, C), self._adt, h.device)
        scache.pop(out.data_ptr(), None)
        return ops.conv_direct([(a, 0, C, 1), (h, 0, C, 1)], e["wo"], e["bo"], out)

    # ------------------------------------------------------------------ forward
    def velocity(self, x, y, t, out=None, base1=None, c1=0.0, base2=None, c2=0.0, coef=1.0, v_out=None,
                 lane=0, base3=None, c3=0.0):
        """v = backbone(x, y, t) with the ODE stage fused into the last kernel:
        out = c1*base1 + c2*base2 + c3*base3 + coef*v.  x, y, bases, out: fp32 [B,F,T,2] (= complex64 [B,F,T])."""
        P = self.prepare()
        tb = self.temb_biases(t)
        B, Fq, T = x.shape[0], x.shape[1], x.shape[2]
        nres = self.num_resolutions
        if Fq % (1 << (nres - 1)) or T % (1 << (nres - 1)):
            raise ValueError(f"spectrogram {Fq}x{T} cannot be halved {nres - 1} times "
                             "(pad_spec pads T to a multiple of 64)")
        # lanes may run concurrently on different streams: one workspace per (lane, shape)
        ws, dev = self._select_ws(lane, B, Fq, T), x.device
        scache = {}
        self._stat_counter = 0
        mods = self.all_modules
        pyr_in = ops.pack4(x, y, ws.get("pyr_in0", (B, Fq, T, 4), torch.float32, dev))
        h = ops.conv_in(pyr_in, P["conv_in_w"], P["conv_in_b"], ws.get("h_in", (B, Fq, T, self.nf), self._adt, dev))
        hs = [h]
        idx = 4
        H, W = Fq, T
        for lvl in range(nres):
            for _ in range(self.num_res_blocks):
                hs.append(self._resblock(idx, [hs[-1]], tb, scache))
                idx += 1
            if lvl != nres - 1:
                hd = self._resblock(idx, [hs[-1]], tb, scache)
                idx += 1
                H, W = H // 2, W // 2
                pyr_in = ops.fir_down4(pyr_in, ws.get(f"pyr_in{lvl + 1}", (B, H, W, 4), torch.float32, dev))
                c = P[idx]
                hc = ops.combine(pyr_in, c["w"], c["b"], hd, ws.get(f"comb{idx}", hd.shape, self._adt, dev))
                scache.pop(hc.data_ptr(), None)
                idx += 1
                hs.append(hc)
        h = hs[-1]
        h = self._resblock(idx, [h], tb, scache)
        idx += 1
        if self.bottleneck_attn:
            h = self._attn(idx, h, scache)
            idx += 1
        h = self._resblock(idx, [h], tb, scache)
        idx += 1
        pyramid = None
        for lvl in reversed(range(nres)):
            for _ in range(self.num_res_blocks + 1):
                h = self._resblock(idx, [h, hs.pop()], tb, scache)
                idx += 1
            g = P[idx]
            idx += 1
            pc = P[idx]
            C = h.shape[3]
            ph = ws.get(f"pyr_out{lvl}", (B, H, W, 4), torch.float32, dev)
            if (self.pyramid_halo or self.precision == "tf32") and ops.halo_eligible(B, H, W, 48) and C % 64 == 0:
                # "GEMM first, shift after" on the halo kernel (N = 48, 1 tap): GroupNorm+SiLU in the operand
                # transform (nothing is materialised), 36 per-tap products per pixel out, then the gather-sum
                ss = self._gn_scale_shift([h], g["g"], g["b"], scache, "gn_ss")
                part = ws.get("pyr_part", (B, H, W, 36), torch.float32, dev)
                ops.conv_igemm([(h, 0, C, 1, ss, 0)], pc["wt"], None, part, self.max_ctas, algo_k=9 * C, algo_cout=4)
                ops.pyramid_gather(part, pc["b4"], pyramid, ph)
                pyramid = ph
            elif self.pyramid_shift_after_gemm and ops.tensor_conv_ok(B, H, W, 48, [C], out_f32=True):
                # one pass over `a`: 36 per-tap products per pixel on tensor cores, then a gather-sum
                a = self._gn_act([h], g["g"], g["b"], 0, scache, "act")
                part = ws.get("pyr_part", (B, H, W, 36), torch.float32, dev)
                ops.conv_igemm([(a, 0, C, 1)], pc["wt"], None, part, self.max_ctas, algo_k=9 * C, algo_cout=4)
                ops.pyramid_gather(part, pc["b4"], pyramid, ph)
                pyramid = ph
            else:
                if ops.tensor_conv_ok(B, H, W, 16, [C], out_f32=True):
                    a = self._gn_act([h], g["g"], g["b"], 0, scache, "act")
                    ops.conv_igemm([(a, 0, C, 9)], pc["w"], pc["b"], ph, self.max_ctas)
                else:
                    ss = self._gn_scale_shift([h], g["g"], g["b"], scache, "gn_ss")
                    ops.conv_direct([(h, 0, C, 9, ss, 0)], pc["w"], pc["b"], ph, cout=4)
                pyramid = ph if pyramid is None else ops.pyramid_up_add(pyramid, ph, ph)
            idx += 1
            if lvl != 0:
                h = self._resblock(idx, [h], tb, scache)
                idx += 1
                H, W = H * 2, W * 2
        assert not hs and idx == len(mods)
        if out is None and v_out is None:
            v_out = torch.empty(B, Fq, T, 2, device=dev, dtype=torch.float32)
        if self.out_k == 1:
            ops.output_axpy(pyramid, P["w_out8"], base1, c1, base2, c2, coef, out, v_out, base3=base3, c3=c3)
        else:
            ops.output_conv3_axpy(pyramid, P["w_out72"], base1, c1, base2, c2, coef, out, v_out, base3=base3, c3=c3)
        return out if out is not None else v_out

    @torch.no_grad()
    def forward(self, x, y, t):
        with torch.cuda.device(x.device):
            return self._forward(x, y, t)

    def _forward(self, x, y, t):
        """x, y: complex64 [B,1,F,T]; t: float tensor with one element (the reference passes a
        shared scalar time, model.py:470-474).  Returns complex64 [B,1,F,T]."""
        xr = torch.view_as_real(x.to(torch.complex64).contiguous()).squeeze(1).contiguous()
        yr = torch.view_as_real(y.to(torch.complex64).contiguous()).squeeze(1).contiguous()
        v = torch.empty_like(xr)
        if torch.is_tensor(t) and t.numel() != 1 and not bool((t == t.flatten()[0]).all()):
            # per-sample times (ncsnpp.py:254 accepts t of shape [B]): samples are independent, so the batch is run in
            # groups of equal t — the time embedding (and with it every conv0 bias) is per group, nothing else changes
            tv = t.flatten().to(xr.device)
            if tv.numel() != xr.shape[0]:
                raise ValueError(f"t has {tv.numel()} entries for a batch of {xr.shape[0]}")
            for val in torch.unique(tv).tolist():
                idx = (tv == val).nonzero().flatten()
                vg = torch.empty(idx.numel(), *xr.shape[1:], device=xr.device, dtype=xr.dtype)
                self.velocity(xr.index_select(0, idx).contiguous(), yr.index_select(0, idx).contiguous(), float(val), v_out=vg)
                v.index_copy_(0, idx, vg)
            return torch.view_as_complex(v).unsqueeze(1)
        if torch.is_tensor(t):
            t = float(t.flatten()[0])
        self.velocity(xr, yr, t, v_out=v)
        return torch.view_as_complex(v).unsqueeze(1)
