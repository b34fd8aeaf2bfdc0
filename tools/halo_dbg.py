"""Where does the halo kernel's MMA thread wait?  (cycle counters of block 0, ablation modes of the transform)

The counters exist only in a debug build:  python tools/ab_bench.py --build "-DFD_HALO_DEBUG=1"   (CPU box)
then on the GPU box:                       python tools/halo_dbg.py    (picks up flowdec_b200/lib_variant.so)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if not os.environ.get("FD_LIB_PATH") and os.path.exists(os.path.join(ROOT, "flowdec_b200", "lib_variant.so")):
    os.environ["FD_LIB_PATH"] = os.path.join(ROOT, "flowdec_b200", "lib_variant.so")
import torch
dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
os.environ["FD_HALO_DBG"] = str(dbg.data_ptr())
from flowdec_b200 import ops
from flowdec_b200.ops import conv_igemm, pack_conv_weight

torch.manual_seed(0)
B, H, W, C = 8, 384, 128, 256
x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
w = (torch.randn(C, C, 3, 3, device="cuda") / 48).to(torch.bfloat16)
b = torch.randn(C, device="cuda")
wp = pack_conv_weight([(w, 9)], npad=C)
out = torch.empty(B, H, W, C, device="cuda", dtype=torch.bfloat16)
ss = torch.ones(B, C, 2, device="cuda")
for name, srcs, halo, mode in [("per-tap pair kernel", [(x, 0, C, 9)], False, 0), ("halo raw", [(x, 0, C, 9)], True, 0),
                               ("halo fused GN+SiLU", [(x, 0, C, 9, ss, 0)], True, 0),
                               ("  fused, barriers only", [(x, 0, C, 9, ss, 0)], True, 1),
                               ("  fused, loads only", [(x, 0, C, 9, ss, 0)], True, 2),
                               ("  fused, no stores", [(x, 0, C, 9, ss, 0)], True, 3)]:
    ops.HALO_TILES = halo
    os.environ["FD_HALO_XF_MODE"] = str(mode)
    for _ in range(3):
        conv_igemm(srcs, wp, b, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        conv_igemm(srcs, wp, b, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * B * H * W * C * C * 9
    d = dbg.tolist()
    line = f"{name:22s} {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s"
    if halo and d[0]:
        line += (f" | block0 MMA thread: total {d[0]} cyc over {d[4]} tiles ({d[0]/max(d[4],1):.0f}/tile; ideal {36*512}), "
                 f"wait tmem-empty {d[1]/d[0]:.1%}, wait A-ready {d[2]/d[0]:.1%}, wait B-full {d[3]/d[0]:.1%}")
    print(line)
    if halo and d[15]:
        print(f"      transform warp 7: waits for TMA {d[14]/max(d[4],1):.0f} cyc/tile, transform+publish {d[15]/max(d[4],1):.0f} cyc/tile "
              f"(of which loads {d[12]/max(d[4],1):.0f}, proxy fence {d[13]/max(d[4],1):.0f})")
