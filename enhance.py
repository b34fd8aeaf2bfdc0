#!/usr/bin/env python
"""File-level enhancement CLI with the reference's flags (/root/reference/enhance.py:24-49) on the
B200-native model.  RTF = runtime / filetime around `model.enhance` with CUDA events, as the
reference defines it (enhance.py:120-136; includes H2D + STFT + ODE + iSTFT + D2H, excludes file I/O).

The reference CLI is broken at its pinned commit (commented-out loader, tuple-unpack bug
with --single-file, --rtf crash without a pairs list); this implements the intended behaviour.

    python enhance.py --ckpt step=800000.ckpt --files in_dir --outdir out_dir --N 3 --solver midpoint [--rtf]
"""
import contextlib
import glob
import os
from argparse import ArgumentParser

import numpy as np
import torch


def load_wav(path):
    """-> (float32 tensor [C, L], sample_rate)"""
    try:
        import torchaudio
        y, sr = torchaudio.load(path)
        return y, sr
    except Exception:
        import scipy.io.wavfile as wavfile
        sr, d = wavfile.read(path)
        if d.dtype.kind in "iu":
            d = d.astype(np.float32) / float(np.iinfo(d.dtype).max + 1)
        d = d.astype(np.float32)
        d = d[None, :] if d.ndim == 1 else d.T
        return torch.from_numpy(np.ascontiguousarray(d)), sr


def save_wav(path, x, sr):
    try:
        import torchaudio
        torchaudio.save(path, x, sr)
    except Exception:
        import scipy.io.wavfile as wavfile
        wavfile.write(path, sr, x.detach().cpu().numpy().T.astype(np.float32))


def read_list(listfile):
    """reference enhance.py:146-164: plain list, or 'clean ---> noisy' / 'clean,noisy' pairs"""
    filenames, from_pairs = [], False
    with open(listfile) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if " ---> " in line:
                from_pairs = True
                filenames.append(line.split(" ---> "))
            elif "," in line:
                from_pairs = True
                filenames.append(line.split(","))
            else:
                assert not from_pairs, "Inconsistent file list format with and without pairs detected!"
                filenames.append(line)
    return filenames, from_pairs


def build_parser():
    p = ArgumentParser()
    p.add_argument("--ckpt", type=str, required=True, help="Lightning checkpoint (.ckpt) to load the model from.")
    p.add_argument("--files", type=str, required=True, help="Input directory or filelist containing *.wav files.")
    p.add_argument("--outdir", type=str, required=True, help="Output directory (created if needed).")
    p.add_argument("--N", type=int, required=True, help="Solver steps (NFE = N for euler, 2N for midpoint).")
    p.add_argument("--single-file", action="store_true", help="treat --files as one wav file")
    p.add_argument("--exclude-files-matching", type=str, required=False)
    p.add_argument("--predictor", type=str, default="reverse_diffusion", choices=["euler_maruyama", "reverse_diffusion"])
    p.add_argument("--corrector", type=str, default="ald", choices=["ald", "none"])
    p.add_argument("--snr", type=float, default=0.5)
    p.add_argument("--solver", type=str, default="midpoint")
    p.add_argument("--device", type=str, default="cuda:0")
    p.add_argument("--ema", type=bool, default=True)
    p.add_argument("--skip-existing", type=bool, default=True)
    p.add_argument("--i-min", type=int, default=None)
    p.add_argument("--i-max", type=int, default=None)
    p.add_argument("--rtf", action="store_true", help="time each file and write rtfs.csv")
    p.add_argument("--variant", type=str, default="75m", help="flowdec_{75m,25s} architecture of the checkpoint")
    p.add_argument("--precision", type=str, default="bf16", choices=["bf16", "tf32"],
                   help="backbone arithmetic: bf16 operands (default, fastest) or fp32 activations with tf32 tensor-core "
                        "operands (the precision class of the reference's own GPU convolutions; ~0.55x the speed)")
    p.add_argument("--batch-files", type=int, default=1,
                   help="enhance up to this many files per model call (length-bucketed, flowdec_b200/batching.py); "
                        "1 = one file per call like the reference")
    return p


def flush_batch(model, pending, enhance_kwargs, batch_files, rtf_f):
    """pending: list of (out_path, y [C,L], sr).  Multi-channel files contribute one clip per channel."""
    from flowdec_b200.batching import enhance_list
    if not pending:
        return
    clips, owner = [], []
    for k, (_, y, _) in enumerate(pending):
        for c in range(y.shape[0]):
            clips.append(y[c])
            owner.append((k, c))
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    outs = enhance_list(model, clips, max_batch=batch_files, **enhance_kwargs)
    end.record()
    torch.cuda.synchronize()
    runtime = start.elapsed_time(end) / 1000.0
    total = sum(y.shape[-1] / sr for _, y, sr in pending)
    print(f"batch of {len(pending)} files: {runtime:.3f} s for {total:.2f} s of audio -> rtf = {runtime / total:.5f}")
    for k, (out_path, y, sr) in enumerate(pending):
        x_hat = torch.stack([outs[j] for j, (kk, _) in enumerate(owner) if kk == k])
        if rtf_f is not None:       # the batch's runtime is attributed in proportion to duration
            filetime = y.shape[-1] / sr
            print(f"{out_path},{runtime * filetime / total:.5f},{filetime:.5f},{runtime / total:.5f}", file=rtf_f)
        save_wav(out_path, x_hat.cpu(), sr)
    pending.clear()


def main(argv=None):
    args = build_parser().parse_args(argv)
    from flowdec_b200.model import EnhancementModel, build_flowdec
    enhance_kwargs = dict(N=args.N, solver=args.solver, predictor=args.predictor, corrector=args.corrector, snr=args.snr)
    print(f"Enhance kwargs: {enhance_kwargs}")
    os.makedirs(args.outdir, exist_ok=True)
    model = EnhancementModel.load_from_checkpoint(args.ckpt, map_location="cpu", ema=args.ema,
                                                  build_fn=lambda: build_flowdec(args.variant))
    model = model.to(args.device).eval()
    if args.precision != "bf16":
        model.set_precision(args.precision)

    clean, trf_path = None, None
    if args.single_file:
        noisy = [args.files]
    elif os.path.isfile(args.files):
        entries, from_pairs = read_list(args.files)
        if from_pairs:
            clean, noisy = [e[0] for e in entries], [e[1] for e in entries]
            suffix = f"_{args.i_min}-{args.i_max}" if args.i_max else ""
            trf_path = os.path.join(args.outdir, f"triples_list{suffix}.txt")
        else:
            noisy = entries
    else:
        noisy = sorted(glob.glob(f"{args.files}/*.wav"))
    if args.exclude_files_matching is not None:
        keep = [i for i, f in enumerate(noisy) if args.exclude_files_matching not in f]
        noisy = [noisy[i] for i in keep]
        clean = [clean[i] for i in keep] if clean else None

    trf_cm = open(trf_path, "w") if trf_path else contextlib.nullcontext()
    rtf_cm = open(os.path.join(args.outdir, "rtfs.csv"), "w") if args.rtf else contextlib.nullcontext()
    with torch.no_grad(), trf_cm as trf, rtf_cm as rtf_f:
        if rtf_f is not None:
            print("path,runtime,filetime,rtf", file=rtf_f)
        pending = []
        for i, path in enumerate(noisy):
            if (args.i_min is not None and i < args.i_min) or (args.i_max is not None and i > args.i_max):
                continue
            out_path = os.path.join(args.outdir, os.path.basename(path))
            if not os.path.exists(out_path) or not args.skip_existing:
                y, sr = load_wav(path)
                if y.shape[-1] / sr > 30.0:                    # reference enhance.py:115,138-139
                    print("Skipping file due to length:", path)
                    if trf is not None:                        # the reference still lists the (unwritten) triple
                        print(f"{clean[i]} ---> {noisy[i]} ---> {out_path}", file=trf)
                    continue
                if sr != model.sampling_rate:
                    import torchaudio
                    print("RESAMPLING from", sr, "to", model.sampling_rate)
                    y = torchaudio.functional.resample(y, sr, model.sampling_rate, lowpass_filter_width=64)
                    sr = model.sampling_rate
                if args.batch_files > 1:
                    pending.append((out_path, y, sr))
                    if len(pending) >= 4 * args.batch_files:      # enough clips to fill the length buckets
                        flush_batch(model, pending, enhance_kwargs, args.batch_files, rtf_f)
                    if trf is not None:
                        print(f"{clean[i]} ---> {noisy[i]} ---> {out_path}", file=trf)
                    continue
                if args.rtf:
                    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    start.record()
                x_hat = model.enhance(y, **enhance_kwargs)
                if args.rtf:
                    end.record()
                    torch.cuda.synchronize()
                    runtime, filetime = start.elapsed_time(end) / 1000.0, y.shape[-1] / sr
                    print(runtime, filetime, "-> rtf =", runtime / filetime)
                    print(f"{out_path},{runtime:.5f},{filetime:.5f},{runtime / filetime:.5f}", file=rtf_f)
                save_wav(out_path, x_hat.cpu(), sr)
            if trf is not None:
                print(f"{clean[i]} ---> {noisy[i]} ---> {out_path}", file=trf)
        flush_batch(model, pending, enhance_kwargs, args.batch_files, rtf_f)


if __name__ == "__main__":
    main()
