"""Probe: UMMA SW128 K-major descriptors with row-shifted start addresses / non-1024 SBO."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
from tools import probe_lib

L = probe_lib.load()
L.fd_umma_probe.restype = ctypes.c_int
L.fd_umma_probe.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
torch.manual_seed(0)
A = torch.randn(256, 64, device="cuda").to(torch.bfloat16)
Bm = torch.randn(16, 64, device="cuda").to(torch.bfloat16)
full = A.float() @ Bm.float().t()            # [256,16]: row r of A against all 16 B rows


def expect(row_off, sbo, phase_fix=None):
    rows = [row_off + (m // 8) * (sbo // 128) + (m % 8) for m in range(128)]
    return full[torch.tensor(rows, device="cuda")]


def best_row_match(out):
    """for each output row find the A row that explains it (or -1)"""
    d = (out[:, None, :] - full[None, :, :]).abs().amax(-1)      # [128,256]
    v, i = d.min(1)
    return [int(ii) if float(vv) < 1e-2 else -1 for vv, ii in zip(v, i)]


for row_off in (0, 1, 3, 8, 10, 17):
    for sbo in (1024, 1280, 2048):
        if row_off + 15 * (sbo // 128) + 8 > 256:
            continue
        for bo in sorted({0, row_off & 7}):
            out = torch.full((128, 16), float("nan"), device="cuda")
            rc = L.fd_umma_probe(A.data_ptr(), Bm.data_ptr(), out.data_ptr(), row_off, sbo, bo,
                                 torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            exp = expect(row_off, sbo)
            ok = bool(((out - exp).abs().amax() < 1e-2).item())
            m = best_row_match(out)
            print(f"row_off={row_off:2d} sbo={sbo:4d} base_offset={bo}: match_expected={ok}  "
                  f"rows[0:10]={m[:10]} rows[8:12]={m[8:12]} unexplained={sum(1 for x in m if x < 0)}")
