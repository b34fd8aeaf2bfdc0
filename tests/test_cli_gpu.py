"""GPU: the file-level CLI (enhance.py) on a synthetic checkpoint and wav files."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_enhance_cli_roundtrip(tmp_path):
    import scipy.io.wavfile as wavfile

    import enhance as cli
    from flowdec_b200.model import build_flowdec
    from flowdec_b200.util.synth import synth_state_dict, synth_waveforms
    m = build_flowdec("75m")
    sd = synth_state_dict(m.state_dict(), seed=0)
    ckpt = tmp_path / "model.ckpt"
    torch.save({"_pl_ema_state_dict": sd, "state_dict": sd, "hyper_parameters": {}}, ckpt)
    indir, outdir = tmp_path / "in", tmp_path / "out"
    indir.mkdir()
    y = synth_waveforms(2, 24000, seed=5)
    wavfile.write(indir / "a.wav", 48000, y[0, 0].numpy())
    wavfile.write(indir / "b.wav", 24000, y[1, 0, ::2].numpy().copy())      # resampled on load
    cli.main(["--ckpt", str(ckpt), "--files", str(indir), "--outdir", str(outdir), "--N", "1", "--solver", "euler", "--rtf"])
    for name in ("a.wav", "b.wav"):
        sr, d = wavfile.read(outdir / name)
        assert sr == 48000 and np.isfinite(d).all() and d.shape[0] >= 23990
    lines = open(outdir / "rtfs.csv").read().strip().splitlines()
    assert lines[0] == "path,runtime,filetime,rtf" and len(lines) == 3
    # length-bucketed batching across files: same outputs' shapes, one rtf line per file
    out2 = tmp_path / "out_batched"
    wavfile.write(indir / "c.wav", 48000, synth_waveforms(1, 30000, seed=6)[0, 0].numpy())
    cli.main(["--ckpt", str(ckpt), "--files", str(indir), "--outdir", str(out2), "--N", "1", "--solver", "euler",
              "--rtf", "--batch-files", "2"])
    for name, n in (("a.wav", 24000), ("b.wav", 23990), ("c.wav", 30000)):
        sr, d = wavfile.read(out2 / name)
        assert sr == 48000 and np.isfinite(d).all() and d.shape[0] >= n
    assert len(open(out2 / "rtfs.csv").read().strip().splitlines()) == 4
