"""Probe: barrier routing of cta_group::2 multicast TMA loads in a 4-CTA cluster (see tools/csrc/fd_probe.cu)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from tools import probe_lib  # noqa: E402

L = probe_lib.load()
L.fd_mcast_probe.restype = ctypes.c_int
L.fd_mcast_probe.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
L.fd_last_error.restype = ctypes.c_char_p
src = torch.arange(16, dtype=torch.float32, device="cuda").reshape(16, 1).repeat(1, 256).contiguous()
for mode in (0, 1):
    res = torch.full((4,), -7, dtype=torch.int32, device="cuda")
    dump = torch.zeros(4, 256, device="cuda")
    rc = L.fd_mcast_probe(src.data_ptr(), res.data_ptr(), dump.data_ptr(), mode, None)
    torch.cuda.synchronize()
    print(f"mode {mode} (barrier operand {'peer bit cleared' if mode == 0 else 'own address'}): rc={rc} "
          f"completed per rank={res.tolist()}")
    for r in range(4):
        vals = sorted(set(dump[r].tolist()))
        print(f"   rank {r} received rows {vals[:6]}{'...' if len(vals) > 6 else ''}")
