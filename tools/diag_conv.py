"""Diagnostic dump for the tcgen05 conv: structured error report for a few tiny cases."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from flowdec_b200.ops import conv_igemm, pack_conv_weight


def ref_conv(x, w, b):
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=w.shape[-1] // 2)
    return y.permute(0, 2, 3, 1).contiguous()


def case(B, H, W, Cin, Cout, k, seed=0, ident=False):
    torch.manual_seed(seed)
    x = torch.randn(B, H, W, Cin, device="cuda").to(torch.bfloat16)
    if ident:
        w = torch.zeros(Cout, Cin, k, k, device="cuda")
        for i in range(min(Cin, Cout)):
            w[i, i, k // 2, k // 2] = 1.0
        w = w.to(torch.bfloat16)
        b = torch.zeros(Cout, device="cuda")
    else:
        w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5).to(torch.bfloat16)
        b = torch.randn(Cout, device="cuda")
    wp = pack_conv_weight([(w, k * k)], npad=Cout)
    out = torch.full((B, H, W, Cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    try:
        conv_igemm([(x, 0, Cin, k * k)], wp, b, out)
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        print(f"case {(B,H,W,Cin,Cout,k,ident)} EXC {e}")
        return
    ref = ref_conv(x, w, b)
    d = (out.float() - ref).abs()
    nan = torch.isnan(out.float()).sum().item()
    print(f"case B{B} H{H} W{W} Cin{Cin} Cout{Cout} k{k} ident={ident}: max_err={d.nan_to_num(9e9).max().item():.4g} "
          f"ref_max={ref.abs().max().item():.3g} nan={nan}")
    if d.nan_to_num(9e9).max().item() > 0.05:
        dd = d.nan_to_num(9e9).reshape(-1, Cout)
        rows_bad = (dd.max(1).values > 0.05).nonzero().flatten()[:16].tolist()
        cols_bad = (dd.max(0).values > 0.05).nonzero().flatten()[:16].tolist()
        print("   first bad pixel rows:", rows_bad)
        print("   first bad channels  :", cols_bad)
        print("   out[0,0,0,:8]", out[0, 0, 0, :8].float().tolist())
        print("   ref[0,0,0,:8]", ref[0, 0, 0, :8].tolist())


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    case(1, 16, 8, 64, 128, 1, ident=True)
    case(1, 16, 8, 64, 128, 1)
    case(1, 16, 8, 128, 128, 1)
    case(1, 16, 8, 64, 256, 1)
    case(1, 16, 8, 64, 128, 3, ident=True)
    case(1, 16, 8, 64, 128, 3)
    case(1, 16, 16, 64, 256, 3)
    case(1, 8, 128, 256, 256, 3)
    case(2, 96, 16, 256, 256, 3)
