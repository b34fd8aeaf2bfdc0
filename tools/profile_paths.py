"""Small drivers for ncu captures of the non-headline paths (the headline step is profiled through bench.py):

    python tools/profile_paths.py ndac     # one tensor-core NDAC decode of 32 x 2 s (codes -> from_codes -> decode)
    python tools/profile_paths.py tf32     # one enhance() of 8 x 2 s, Euler N=1, backbone in tf32 precision

The profiled region is bracketed by cudaProfilerStart/Stop (use `ncu --profile-from-start off`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def ndac():
    from flowdec_b200.ndac import DAC
    from flowdec_b200.util.synth import synth_dac_state_dict
    rates, nq, latent, dim = (8, 5, 4, 4), 10, 1024, 1536
    dac = DAC(synth_dac_state_dict(latent, dim, rates, nq, seed=7), decoder_dim=dim, decoder_rates=rates,
              n_codebooks=nq, latent_dim=latent, sample_rate=48000).to("cuda").eval()
    codes = torch.randint(0, 1024, (32, nq, 150), generator=torch.Generator().manual_seed(3)).cuda()
    for _ in range(2):
        x = dac.decode(dac.quantizer.from_codes(codes)[0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    x = dac.decode(dac.quantizer.from_codes(codes)[0])
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(f"ndac decode 32 x 2 s: {e0.elapsed_time(e1):.2f} ms, out {tuple(x.shape)}")


def tf32():
    from flowdec_b200.model import build_flowdec
    from flowdec_b200.util.synth import synth_state_dict, synth_waveforms
    m = build_flowdec("75m")
    m.load_state_dict(synth_state_dict(m.state_dict(), seed=0))
    m = m.cuda().set_precision("tf32")
    m.use_cuda_graph = False
    y = synth_waveforms(8, 96000, seed=1).cuda()
    for _ in range(2):
        m.enhance(y, N=1, solver="euler")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    m.enhance(y, N=1, solver="euler")
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(f"tf32 enhance 8 x 2 s, NFE 1: {e0.elapsed_time(e1):.2f} ms = {16 / e0.elapsed_time(e1) * 1e3:.1f} audio-s/s per NFE")


if __name__ == "__main__":
    {"ndac": ndac, "tf32": tf32}[sys.argv[1]]()
