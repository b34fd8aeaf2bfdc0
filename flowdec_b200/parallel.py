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


def nccl_scatter_enhance_gather(model, dist, y_full, N, solver, seed=4321):
    """The sharded form of enhance() with its two collectives on the wire: rank 0 holds the full
    batch y_full [B,1,L] on its GPU; `dist.scatter` hands every rank its contiguous shard (NCCL over NVLink),
    each rank enhances it with noise seeded by GLOBAL clip index, `dist.all_gather_into_tensor` returns the full
    enhanced batch to every rank.  B must be divisible by the world size (equal shards, as NCCL scatter needs);
    y_full is only read on rank 0 (other ranks may pass an empty tensor of the same shape/dtype on their device)."""
    from .util.other import padded_frames
    world, rank = dist.get_world_size(), dist.get_rank()
    B, _, L = y_full.shape
    if B % world:
        raise ValueError(f"batch {B} is not divisible by the world size {world}")
    per = B // world
    dev = y_full.device
    shard = torch.empty(per, 1, L, device=dev, dtype=torch.float32)
    dist.scatter(shard, [c.contiguous() for c in y_full.split(per)] if rank == 0 else None, src=0)
    Tp = padded_frames(1 + L // 384)
    noise = torch.stack([clip_noise(i, Tp, seed) for i in range(rank * per, (rank + 1) * per)], 0).to(dev)
    out = model.enhance(shard, N=N, solver=solver, noise=noise).contiguous()
    full = torch.empty(B, 1, L, device=dev, dtype=torch.float32)
    dist.all_gather_into_tensor(full, out)
    return full


def verify_sharding(model, dist, L, N, solver, clips_per_rank=2, seed=4321, wave_seed=777):
    """Hardware check of the sharding contract: the batch enhanced in shards on all ranks (scatter -> enhance ->
    all_gather) must equal, BIT FOR BIT, the same batch enhanced on rank 0 alone.  Returns True/False on rank 0
    (None elsewhere)."""
    from .util.other import padded_frames
    from .util.synth import synth_waveforms
    world, rank = dist.get_world_size(), dist.get_rank()
    B = clips_per_rank * world
    dev = model.device
    y = synth_waveforms(B, L, seed=wave_seed).to(dev) if rank == 0 else torch.empty(B, 1, L, device=dev)
    full = nccl_scatter_enhance_gather(model, dist, y, N, solver, seed)
    if rank != 0:
        return None
    Tp = padded_frames(1 + L // 384)
    noise = torch.stack([clip_noise(i, Tp, seed) for i in range(B)], 0).to(dev)
    alone = model.enhance(y, N=N, solver=solver, noise=noise)
    return bool(torch.equal(alone, full))
