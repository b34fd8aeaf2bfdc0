"""Batch sharding of enhance() over the GPUs of one box.

Every clip is independent end to end (per-sample normfac, per-sample GroupNorm, shared t), so
the path shards over the clip batch with NO data-path collective: each rank enhances a
contiguous shard; results are gathered only if the caller asks for the full batch.  Noise is
drawn per clip from a generator seeded by the GLOBAL clip index, so the sharded result is
bit-identical to the single-GPU run of the same clips for any world size.
"""
import torch


def shard_bounds(n_clips, world, rank):
    """contiguous shards, sizes differ by at most one"""
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def clip_noise(global_index, Tp, seed=4321, n_bins=768):
    """complex64 [1, n_bins, Tp] standard complex normal (real/imag var 1/2 each, like
    torch.randn_like on a complex tensor), a function of (seed, global clip index) only."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed * 1_000_003 + global_index)
    return torch.randn(1, n_bins, Tp, dtype=torch.complex64, generator=g)


def enhance_sharded(model_fn, y, N, solver, rank, world, seed=4321, gather=None):
    """y: full batch [B,1,L] (every rank holds or can load it; only its shard is touched).
    model_fn(y_shard, noise) -> enhanced shard.  gather: None -> returns the local shard and
    (lo, hi); or a callable(list_of_tensors_per_rank <- local) implementing all_gather."""
    from .util.other import padded_frames
    B, _, L = y.shape
    lo, hi = shard_bounds(B, world, rank)
    Tp = padded_frames(1 + L // 384)
    if hi > lo:
        noise = torch.stack([clip_noise(i, Tp, seed) for i in range(lo, hi)], 0)  # [b,1,768,Tp]
        out = model_fn(y[lo:hi], noise)
    else:
        out = y[lo:hi].clone()
    if gather is None:
        return out, (lo, hi)
    return gather(out, [shard_bounds(B, world, r) for r in range(world)])
