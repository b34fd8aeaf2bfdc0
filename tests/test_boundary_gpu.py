"""GPU: the drop-in boundary.  (1) fd_upfirdn2d_f32 — the C-ABI form of the reference's native op
`upfirdn2d(input, kernel, up, down, pad)` (op/upfirdn2d.cpp:38-48) — against the oracle's restatement of
op/upfirdn2d.py:182-224 and the reference-generated golden vectors; (2) the ctypes stub printed in INTEGRATION.md §2,
executed verbatim."""
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from flowdec_b200 import _lib, ops
from oracle import flowdec_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "flowdec_75m_seed0.npz")


def upfirdn2d_ref(x, k, up, down, pad):
    """CPU restatement of upfirdn2d_native (op/upfirdn2d.py:182-224) with up_x = up_y, pad_x = pad_y"""
    n, c, h, w = x.shape
    kh, kw = k.shape
    u = torch.zeros(n * c, 1, h * up, w * up)
    u[:, :, ::up, ::up] = x.reshape(n * c, 1, h, w)
    u = F.pad(u, [max(pad[0], 0), max(pad[1], 0), max(pad[0], 0), max(pad[1], 0)])
    u = u[:, :, max(-pad[0], 0):u.shape[2] - max(-pad[1], 0), max(-pad[0], 0):u.shape[3] - max(-pad[1], 0)]
    o = F.conv2d(u, torch.flip(k, [0, 1]).reshape(1, 1, kh, kw))[:, :, ::down, ::down]
    return o.reshape(n, c, o.shape[2], o.shape[3])


def fir_kernel(factor_gain):
    k = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k = torch.outer(k, k)
    return k / k.sum() * factor_gain


def test_upfirdn2d_vs_reference_golden():
    """the exact calls upsample_2d / downsample_2d make (up_or_down_sampling.py:246-248,279-282) on the golden input"""
    G = np.load(GOLD)
    xf = torch.randn(2, 8, 12, 16, generator=torch.Generator().manual_seed(5))
    up = ops.upfirdn2d(xf.cuda(), fir_kernel(4.0).cuda(), up=2, pad=(2, 1))
    down = ops.upfirdn2d(xf.cuda(), fir_kernel(1.0).cuda(), down=2, pad=(1, 1))
    assert up.shape == (2, 8, 24, 32) and down.shape == (2, 8, 6, 8)
    assert torch.allclose(up.cpu(), torch.from_numpy(G["fir_up"]), rtol=1e-6, atol=1e-6)
    assert torch.allclose(down.cpu(), torch.from_numpy(G["fir_down"]), rtol=1e-6, atol=1e-6)
    assert torch.allclose(up.cpu(), O.fir_up2(xf), rtol=1e-6, atol=1e-6)
    assert torch.allclose(down.cpu(), O.fir_down2(xf), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("shape,kshape,up,down,pad", [
    ((1, 3, 7, 9), (4, 4), 1, 1, (0, 0)), ((2, 2, 8, 8), (4, 4), 2, 1, (2, 1)), ((2, 2, 9, 11), (4, 4), 1, 2, (1, 1)),
    ((1, 1, 5, 6), (3, 5), 3, 2, (4, 2)), ((1, 2, 16, 16), (2, 2), 2, 1, (1, 0)), ((1, 1, 12, 10), (4, 4), 1, 1, (-1, -2)),
    ((3, 5, 33, 17), (6, 6), 2, 3, (5, 3))])
def test_upfirdn2d_general(shape, kshape, up, down, pad):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(*shape, generator=g)
    k = torch.randn(*kshape, generator=g)
    ref = upfirdn2d_ref(x, k, up, down, pad)
    out = ops.upfirdn2d(x.cuda(), k.cuda(), up=up, down=down, pad=pad)
    assert out.shape == ref.shape
    assert torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-5)


def test_upfirdn2d_errors():
    x = torch.zeros(1, 1, 2, 2, device="cuda")
    k = torch.zeros(4, 4, device="cuda")
    rc = _lib.lib().fd_upfirdn2d_f32(_lib.ptr(x), 1, 2, 2, _lib.ptr(k), 4, 4, 1, 1, 1, 1, 0, 0, 0, 0, _lib.ptr(x),
                                     _lib.stream_ptr())
    assert rc != 0 and b"empty output" in _lib.lib().fd_last_error()
    rc = _lib.lib().fd_upfirdn2d_f32(_lib.ptr(x), 1, 2, 2, _lib.ptr(k), 4, 4, 0, 1, 1, 1, 2, 2, 2, 2, _lib.ptr(x),
                                     _lib.stream_ptr())
    assert rc != 0


def test_integration_stub_verbatim():
    """INTEGRATION.md §2 is executable documentation: run the printed stub as is"""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"<!-- integration-stub:begin -->\s*```python\n(.*?)```\s*<!-- integration-stub:end -->", text, re.S)
    assert m, "INTEGRATION.md lost its integration-stub block"
    os.environ["FLOWDEC_B200_LIB"] = _lib._LIB_PATH
    ns = {}
    exec(compile(m.group(1), "INTEGRATION.md#stub", "exec"), ns)
    G = np.load(GOLD)
    xf = torch.randn(2, 8, 12, 16, generator=torch.Generator().manual_seed(5))
    up = ns["upfirdn2d"](xf.cuda(), fir_kernel(4.0).cuda(), up=2, pad=(2, 1))
    down = ns["upfirdn2d"](xf.cuda(), fir_kernel(1.0).cuda(), down=2, pad=(1, 1))
    assert torch.allclose(up.cpu(), torch.from_numpy(G["fir_up"]), rtol=1e-6, atol=1e-6)
    assert torch.allclose(down.cpu(), torch.from_numpy(G["fir_down"]), rtol=1e-6, atol=1e-6)
    with pytest.raises(RuntimeError):
        ns["upfirdn2d"](xf, fir_kernel(1.0).cuda())
    # fused-path form: bf16 NHWC, raw FIR only
    xb = torch.randn(2, 16, 24, 64, generator=torch.Generator().manual_seed(6)).to(torch.bfloat16)
    xc = xb.float().permute(0, 3, 1, 2)
    for is_up, ref in ((True, O.fir_up2(xc)), (False, O.fir_down2(xc))):
        o = ns["upfirdn2d_2x"](xb.cuda(), is_up).float().cpu().permute(0, 3, 1, 2)
        assert (o - ref).abs().max() <= 2 ** -8 * ref.abs().max() + 1e-3
