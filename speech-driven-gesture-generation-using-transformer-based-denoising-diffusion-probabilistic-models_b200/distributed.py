"""Clip-sharded sampling across the GPUs of one box (one process per GPU, torch.distributed).

Every clip is an independent chain, so the batch is split into contiguous slices, each rank samples its own clips
with its slice of the speech / x_T / noise tape, and the generated poses are collected with ONE all-gather after the
chain (NCCL over NVLink on GPUs, gloo in the CPU tests).  There is no per-step communication.  The kernels are
batch-invariant (fixed K order, no split-K, per-clip attention/LayerNorm), so a clip's result does not depend on
which rank or which batch it was sampled in.
"""
import torch as th
import torch.distributed as dist


def shard_bounds(n_clips, world_size, rank):
    """Contiguous slice [lo, hi) of rank `rank`; sizes differ by at most one clip, earlier ranks take the extras."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_clips, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_clips(local, n_clips, group=None):
    """all-gather per-rank results (n_local, ...) into the full (n_clips, ...) tensor, present on every rank.
    Shards may be ragged (or empty): they are padded to the largest shard for the collective and trimmed after."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    biggest = shard_bounds(n_clips, world, 0)[1]
    padded = local.new_zeros((biggest,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    out = local.new_empty((world * biggest,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_clips, world, r)
        parts.append(out[r * biggest:r * biggest + (hi - lo)])
    return th.cat(parts, dim=0)


def sample_sharded(sample_fn, wavs, noise=None, noise_tape=None, group=None):
    """Run `sample_fn(wavs, noise, noise_tape) -> (n, T, C)` on this rank's clip slice and gather all clips.
    `wavs` (N, T_wav), `noise` (N, C, T) and `noise_tape` (steps, N, C, T) are the FULL-batch tensors (or None)."""
    n = wavs.shape[0]
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    lo, hi = shard_bounds(n, world, rank)
    local = sample_fn(wavs[lo:hi], None if noise is None else noise[lo:hi],
                      None if noise_tape is None else noise_tape[:, lo:hi])
    return gather_clips(local, n, group)


def generate_sample_sharded(generator, shape, wavs, noise=None, noise_tape=None, group=None, **kw):
    """`Generator.generate_sample` over a clip-sharded batch; returns all N clips' poses (N, T, C) on every rank."""
    _, C, T = shape

    def fn(w, x, tape):
        if w.shape[0] == 0:
            return th.empty(0, T, C, device=kw.get("device", "cpu"))
        return generator.generate_sample((w.shape[0], C, T), w, noise=x, noise_tape=tape, **kw)

    return sample_sharded(fn, wavs, noise, noise_tape, group)
