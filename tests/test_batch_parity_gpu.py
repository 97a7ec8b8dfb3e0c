"""Parity at the BENCHMARKED batch sizes and across processes.

* every clip's chain is independent of the batch it runs in: clips [0:2] of a tedexp 256-clip / beat 1024-clip run
  (whole-chain CUDA graph, concurrent branches, CTA-pair GEMMs, persistent multi-item attention) are BIT-identical to
  the 2-clip run that tests/test_parity_golden_gpu.py pins against the reference's golden vectors;
* the result does not depend on how the kernels are scheduled: graph vs eager launches, programmatic dependent launch
  on/off, in fresh processes, give the same bytes (the reference is deterministic given the noise tape,
  gaussian_diffusion.py:393-412).
"""
import json
import os
import subprocess
import sys

import pytest
import torch as th

from util import ROOT, build, load_golden, rel_l2, synthetic_wav

pytestmark = pytest.mark.gpu

BENCH_BATCH = {"tedexp": 256, "beat": 1024}


def _device_tape(shape, n_steps, seed):
    g = th.Generator(device="cuda").manual_seed(seed)
    x_T = th.randn(shape, device="cuda", generator=g)
    tape = th.randn((n_steps,) + tuple(shape), device="cuda", generator=g)
    return x_T, tape


def _run(model, diffusion, shape, wav, x_T, tape, alg="ddpm"):
    from gesture_b200.generator import Generator
    kw = {"noise_tape": tape} if alg == "ddpm" else {}
    out = Generator(model, diffusion).generate_sample(shape, wav, noise=x_T, sample_alg=alg, device="cuda", progress=False, **kw)
    th.cuda.synchronize()
    return out


@pytest.mark.parametrize("name", ["tedexp", "beat"])
@pytest.mark.parametrize("precision,respacing", [("bf16", "ddim20"), ("fp32act", "ddim20"), ("bf16", "")])
def test_benchmark_batch_is_bit_identical_to_two_clip_run(name, precision, respacing):
    """ddim20-spaced ancestral chain and (bf16) the full 1000-step chain at the bench batch vs the same clips run 2 at a time."""
    N = BENCH_BATCH[name]
    model, diffusion, C, T, L, params = build(name, "boost", respacing=respacing, device="cuda")
    model.precision = precision
    n_steps = diffusion.num_timesteps
    wav = synthetic_wav(N, L, seed=123)
    x_T, tape = _device_tape((N, C, T), n_steps, seed=7)
    big = _run(model, diffusion, (N, C, T), wav, x_T, tape).clone()
    assert th.isfinite(big).all()
    for lo in (0, N - 2):  # first and last shard of the batch
        small = _run(model, diffusion, (2, C, T), wav[lo:lo + 2], x_T[lo:lo + 2].contiguous(),
                     tape[:, lo:lo + 2].contiguous())
        assert th.equal(big[lo:lo + 2], small), (
            f"{name}/{precision}/{respacing or 'full'}: clips [{lo}:{lo + 2}] of the {N}-clip run differ from the 2-clip run "
            f"(rel-L2 {rel_l2(big[lo:lo + 2], small):.3e})")


@pytest.mark.parametrize("name", ["tedexp", "beat"])
def test_benchmark_batch_matches_reference_golden(name):
    """Clips [0:2] of the bench-size batch carry the golden inputs of the reference run (same wav / x_T / tape as
    tests/golden/make_golden.py), the rest of the batch is other clips: the full 1000-step chain at N = 256 / 1024 must
    land on the reference's final poses."""
    from util import noise_tape
    g = load_golden(name)
    N = BENCH_BATCH[name]
    model, diffusion, C, T, L, params = build(name, "boost", device="cuda")
    wav = th.cat([synthetic_wav(2, L, seed=123), synthetic_wav(N - 2, L, seed=124)])
    x2, tape2 = noise_tape((2, C, T), 1000, seed=99)
    x_T, tape = _device_tape((N, C, T), 1000, seed=8)
    x_T[:2] = x2.cuda()
    tape[:, :2] = tape2.cuda()
    out = _run(model, diffusion, (N, C, T), wav, x_T, tape)
    err = rel_l2(out[:2].transpose(1, 2), g["boost.ddpm.final"])
    print(f"[{name}] N={N} full chain, golden clips: final pose rel-L2 vs reference {err:.3e}")
    assert err < 2e-2, err


VARIANTS = [("default", {}), ("pdl0", {"GD_PDL": "0"}), ("eager", {"GD_GRAPH": "0"}), ("eager_pdl0", {"GD_GRAPH": "0", "GD_PDL": "0"}),
            ("noconc", {"GD_CONCURRENT": "0"})]


@pytest.mark.parametrize("workload,n_clips", [("beat-ours", 2), ("beat-ours", 300), ("tedexp-ours", 2), ("tedexp-ours", 64)])
def test_cross_process_bit_identity(workload, n_clips):
    """The smoke chain (10-step process, the spacing that exposed the stale step-counter read of round 1) in fresh
    processes: graph / eager, PDL on / off, concurrent branches on / off must all print the same sha256."""
    script = os.path.join(ROOT, "profiles", "determinism_bisect.py")
    seen = {}
    for name, env in (VARIANTS if n_clips <= 2 else VARIANTS[:3]):
        e = dict(os.environ)
        e.update(env)
        p = subprocess.run([sys.executable, script, "--child", workload, str(n_clips), "ddim10"], env=e, capture_output=True,
                           text=True, timeout=900)
        lines = [l for l in p.stdout.splitlines() if l.startswith("HASH ")]
        assert p.returncode == 0 and lines, p.stderr[-800:]
        seen[name] = json.loads(lines[0][5:])
    ref = seen["default"]
    assert ref["x"] == ref["x_again"], "graph replay differs from the first run in the same process"
    for name, h in seen.items():
        assert h == ref, f"{workload} N={n_clips}: variant {name} differs from default: {h} vs {ref}"


@pytest.mark.parametrize("name,N,parts", [("beat", 128, 4), ("beat", 50, 3), ("tedexp", 32, 2)])
@pytest.mark.parametrize("alg", ["ddpm", "ddim"])
def test_parallel_sub_chains_do_not_change_any_clip(name, N, parts, alg):
    """SplitChain (engine.py): the batch sampled as `parts` parallel sub-chains inside one CUDA graph gives the same bytes as
    the single chain (ragged split included), with in-painting on, for both samplers."""
    from gesture_b200.generator import Generator
    model, diffusion, C, T, L, params = build(name, "boost", respacing="ddim20", device="cuda")
    wav = synthetic_wav(N, L, seed=41)
    x_T, tape = _device_tape((N, C, T), 20, seed=42)
    seedp = th.randn(N, T, C, generator=th.Generator().manual_seed(43))
    masks = th.ones(N, T, 1)
    masks[:, 6:] = 0
    outs = []
    for k in (1, parts):
        model.sub_chains = k
        kw = {"noise_tape": tape} if alg == "ddpm" else {}
        outs.append(Generator(model, diffusion).generate_sample((N, C, T), wav, noise=x_T, sample_alg=alg, device="cuda",
                                                                progress=False, inpaint_poses=seedp, inpaint_masks=masks,
                                                                trans_factor=0.5, pose_seed_len=6, **kw).clone())
    assert th.isfinite(outs[0]).all() and th.equal(outs[0], outs[1]), f"rel-L2 {rel_l2(outs[1], outs[0]):.3e}"
