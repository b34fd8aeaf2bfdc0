"""Probe: does a row-misaligned / non-1024-SBO A descriptor slow the tensor pipe down?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
from tools import probe_lib

L = probe_lib.load()
L.fd_umma_rate.restype = ctypes.c_int
L.fd_umma_rate.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p]
A = torch.randn(256, 64, device="cuda").to(torch.bfloat16)
Bm = torch.randn(256, 64, device="cuda").to(torch.bfloat16)
cyc = torch.zeros(1, dtype=torch.int64, device="cuda")
iters = 4000
for row_off, sbo in [(0, 1024), (8, 1024), (1, 1024), (3, 1024), (0, 1280), (1, 1280), (11, 1280), (21, 1280), (0, 2048)]:
    res = []
    for _ in range(3):
        rc = L.fd_umma_rate(A.data_ptr(), Bm.data_ptr(), cyc.data_ptr(), row_off, sbo, iters,
                            torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        res.append(cyc.item() / iters)
    print(f"row_off={row_off:2d} sbo={sbo}: cycles per 128x256x16 MMA = {min(res):.1f}")
