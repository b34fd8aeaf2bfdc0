import torch, sys
sys.path.insert(0, "/root/repo")
from flowdec_b200.model import build_flowdec
from flowdec_b200.util.synth import synth_state_dict
m = build_flowdec("75m"); m.load_state_dict(synth_state_dict(m.state_dict(), seed=0)); m = m.cuda()
g = torch.Generator().manual_seed(9)
eps = torch.randn(2, 1, 768, 64, dtype=torch.complex64, generator=g)
y = torch.zeros(2, 1, 24000); y[1] = 0.1 * torch.randn(1, 24000, generator=g)
for halo in (True, False):
    m.backbone.pyramid_halo = halo
    m.reset_cache()
    for N, solver in ((1, "euler"), (3, "midpoint")):
        out = m.enhance(y, N=N, solver=solver, noise=eps)
        print("pyramid_halo", halo, solver, N, "zero clip max", out[0].abs().max().item(), "other", out[1].abs().max().item())
torch.save(dict(eps=eps, y=y), "/root/repo/gpurun_out/zero_in.pt")
