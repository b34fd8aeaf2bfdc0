"""Length-bucketed batching across files.

The reference CLI enhances one file per `model.enhance` call (/root/reference/enhance.py:113-131), i.e.
batch 1 — the worst case for a B200 (one 2 s clip keeps < 1/8 of the tensor pipes busy).  Clips whose
STFTs pad to the same number of frames (`pad_spec`, util/other.py:25-52: 64*ceil(frames/64)) run through
an identically shaped backbone, so they can share one batch *exactly*: `FlowModel.enhance(lengths=)`
gives every clip its own normfac, frame count, reflect padding and istft length, and the per-sample
GroupNorm makes the backbone independent across the batch.  With the same x0 noise the result for a clip
is bit-identical to enhancing it alone (tests/test_backbone_gpu.py::test_ragged_batch_matches_single_clips).
"""
from collections import defaultdict

import torch

from .util.other import padded_frames


def frames_bucket(length):
    """padded STFT frame count of a clip of `length` samples (n_fft 1534, hop 384, centred)"""
    return padded_frames(1 + int(length) // 384)


def bucket_by_frames(lengths, max_batch):
    """-> list of index lists; every list holds <= max_batch clips of one padded-frame bucket.
    Buckets are emitted longest first, clips inside a bucket in input order (deterministic)."""
    if max_batch < 1:
        raise ValueError("max_batch must be >= 1")
    groups = defaultdict(list)
    for i, n in enumerate(lengths):
        if int(n) <= 767:
            raise ValueError(f"clip {i}: length {n} must exceed the STFT reflect pad (767)")
        groups[frames_bucket(n)].append(i)
    batches = []
    for tp in sorted(groups, reverse=True):
        idx = groups[tp]
        batches += [idx[k:k + max_batch] for k in range(0, len(idx), max_batch)]
    return batches


def pad_batch(waves):
    """list of 1-D (or [1, L]) float tensors -> (zero-padded [B, 1, Lmax] tensor, lengths list)"""
    flat = [w.reshape(-1).float() for w in waves]
    lens = [int(w.numel()) for w in flat]
    out = torch.zeros(len(flat), 1, max(lens), dtype=torch.float32, device=flat[0].device)
    for i, w in enumerate(flat):
        out[i, 0, :lens[i]] = w
    return out, lens


@torch.no_grad()
def enhance_list(model, waves, max_batch=32, **enhance_kwargs):
    """Enhance a list of mono clips of arbitrary lengths; returns the enhanced clips in input order,
    each with the shape of its input.  `enhance_kwargs` go to FlowModel.enhance (N, solver, sigma_fac)."""
    outs = [None] * len(waves)
    for idx in bucket_by_frames([w.numel() for w in waves], max_batch):
        y, lens = pad_batch([waves[i] for i in idx])
        x = model.enhance(y, lengths=lens, **enhance_kwargs)
        for k, i in enumerate(idx):
            outs[i] = x[k, 0, :lens[k]].reshape(waves[i].shape).to(waves[i].device)
    return outs
