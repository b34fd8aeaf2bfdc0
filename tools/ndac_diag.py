import sys, time, torch
sys.path.insert(0, "/root/repo")
from flowdec_b200.ndac import DAC
from flowdec_b200.util.synth import synth_dac_state_dict, synth_state_dict
from flowdec_b200.model import build_flowdec
rates, nq, latent, dim = (8, 5, 4, 4), 10, 1024, 1536
dac = DAC(synth_dac_state_dict(latent, dim, rates, nq, seed=7), decoder_dim=dim, decoder_rates=rates, n_codebooks=nq, latent_dim=latent, sample_rate=48000).to("cuda").eval()
codes = torch.randint(0, 1024, (32, nq, 150), generator=torch.Generator().manual_seed(3)).cuda()
def tdec(tag):
    zq = dac.quantizer.from_codes(codes)[0]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); x = dac.decode(zq); e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"{tag}: gpu {e0.elapsed_time(e1):.2f} ms, host enqueue {1e3*(t1-t0):.2f} ms, reserved {torch.cuda.memory_reserved()/2**30:.1f} GiB", flush=True)
    return x
for i in range(3): x = tdec(f"cold {i}")
m = build_flowdec("75m"); m.load_state_dict(synth_state_dict(m.state_dict(), seed=0)); m = m.cuda()
for i in range(3):
    out = m.enhance(x, N=3, solver="midpoint"); torch.cuda.synchronize()
    x = tdec(f"after enhance {i}")
m.reset_cache()
for i in range(2): x = tdec(f"after reset {i}")
