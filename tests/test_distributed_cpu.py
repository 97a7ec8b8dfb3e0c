"""world_size-2 (and 3, ragged) gloo tests of the clip-sharding host logic; the sampler is a deterministic stand-in
because the CUDA chain itself cannot run on the CPU box."""
import os
import socket

import pytest
import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp

import util  # noqa: F401  (sys.path)
from gesture_b200.distributed import gather_clips, sample_sharded, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 5, 8, 1024, 1023):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_bounds(1024, 8, 3) == (384, 512)
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _fake_sampler(w, x, tape):
    # per-clip deterministic function of the rank's slice: (n, T=3, C=2)
    n = w.shape[0]
    base = w.sum(dim=1).view(n, 1, 1) + (0 if x is None else x.sum(dim=(1, 2)).view(n, 1, 1))
    if tape is not None:
        base = base + tape.sum(dim=(0, 2, 3)).view(n, 1, 1)
    return base + th.arange(6, dtype=th.float32).view(1, 3, 2)


def _worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = th.Generator().manual_seed(0)
        wavs = th.randn(n_clips, 16, generator=g)
        noise = th.randn(n_clips, 2, 3, generator=g)
        tape = th.randn(4, n_clips, 2, 3, generator=g)
        out = sample_sharded(_fake_sampler, wavs, noise, tape)
        ref = _fake_sampler(wavs, noise, tape)
        ok = out.shape == ref.shape and th.equal(out, ref)
        # ragged / empty shards through the raw gather
        lo, hi = shard_bounds(n_clips, world, rank)
        got = gather_clips(ref[lo:hi].clone(), n_clips)
        ok = ok and th.equal(got, ref)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n_clips", [(2, 8), (2, 7), (3, 4), (2, 1)])
def test_sharded_sampling_gloo(world, n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(r, True) for r in range(world)]


def test_single_process_passthrough():
    wavs = th.randn(5, 16)
    assert th.equal(sample_sharded(_fake_sampler, wavs), _fake_sampler(wavs, None, None))
