#!/usr/bin/env python
"""Benchmark of the DDPM reverse-sampling hot path (BASELINE.json metric: generated gesture frames/sec for the full
1000-step chain; ms per denoise step).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload tedexp-ours|beat-ours]
                    [--clips B] [--precision bf16|fp32act]

One bench "step" = one full 1000-step ancestral chain over the rank's batch of clips.  Workload at N=1 is
BASELINE.json configs[1]: tedexp-ours, 256 clips, bf16, CUDA-graph replays.  N>1 (torchrun) shards clips: every rank
samples its own 256 clips (weak scaling) and the generated poses are all-gathered over NCCL inside the timed region.
`value` times the chain with inputs resident in HBM; `e2e` goes through Generator.generate_sample with pinned HOST
wav/noise buffers and a device->host read of the poses.  `--impl reference` times the CPU oracle port of the reference's
sampler (the reference is pure Python and cannot travel to the GPU box) on a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def workload_preset(name):
    """-> (params, d_pose, frames, wav_len, default clips per GPU) of a bench workload."""
    from gesture_b200.presets import preset
    if name == "beat-ours-4x":
        params, C, T, L = preset("beat-ours")
        return params, C, 4 * T, 4 * L, 64
    params, C, T, L = preset(name)
    return params, C, T, L, (256 if name == "tedexp-ours" else 1024)

import torch as th  # noqa: E402

METRIC = "generated gesture frames/sec, full 1000-step DDPM chain"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="tedexp-ours", choices=["tedexp-ours", "beat-ours", "beat-ours-4x"],
                    help="beat-ours-4x = BASELINE config 5: the beat model at 4x its window (160 frames, 8 s of speech)")
    ap.add_argument("--clips", type=int, default=None, help="clips per GPU (default 256 tedexp / 1024 beat / 64 beat-4x)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32act"])
    ap.add_argument("--graph-steps", type=int, default=0,
                    help="denoise steps captured per CUDA graph (0 = the whole chain as one graph, the default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------------------------- CPU reference arm
class CpuReference:
    """The oracle port of the reference's sampler AS SHIPPED (speech encoder re-run inside every denoiser call,
    models/model.py:54-56), timed for a few denoise steps of a `clips`-clip batch and extrapolated to the chain."""

    def __init__(self, workload, clips, threads=None):
        import gesture_b200  # noqa: F401
        from gesture_b200.model_creation import create_model
        from gesture_b200.presets import preset
        from gesture_b200.synthetic import noise_tape, synthetic_wav
        from oracle import ddpm_oracle as orc
        # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it can
        th.set_num_threads(threads or os.cpu_count() or 1)
        self.orc = orc
        self.params, self.C, self.T, L, _ = workload_preset(workload)
        th.manual_seed(0)
        model, _, *_ = create_model(self.C, self.params)  # parameter container only; the oracle does the arithmetic
        self.sd = dict(model.state_dict())
        self.tabs = orc.spaced_diffusion_tables("linear", 1000, "")
        self.clips = clips
        self.wav = synthetic_wav(clips, L, seed=123)
        self.x_T, self.tape = noise_tape((clips, self.C, self.T), 8, seed=99)

    def sample(self, denoise_steps):
        t0 = time.perf_counter()
        self.orc.sample_chain(self.sd, self.params.type, self.params.Decoder.heads, self.tabs, self.x_T, self.wav,
                              self.tape.repeat((denoise_steps + 7) // 8, 1, 1, 1), steps=denoise_steps, reencode_every_step=True)
        dt = time.perf_counter() - t0
        per_step = dt / denoise_steps
        return {"frames_per_s": self.clips * self.T / (per_step * 1000), "ms_per_denoise_step": per_step * 1e3, "seconds": dt,
                "cores": th.get_num_threads(), "clips": self.clips, "denoise_steps": denoise_steps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    clips, per_step = 16, 2
    ref = CpuReference(args.workload, clips)
    vals, secs = [], []
    for it in range(args.warmup + args.steps):
        r = ref.sample(per_step)
        if it >= args.warmup:
            vals.append(r["frames_per_s"])
            secs.append(r["seconds"])
    v = sum(vals) / len(vals)
    sample = (f"{clips} clips x {per_step} denoise steps per bench step of the {args.workload} chain, as shipped (speech encoder "
              f"re-run every step), fp32 torch CPU, extrapolated x(1000/{per_step}) to the full chain")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs) * (1000 / per_step), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload} full 1000-step DDPM chain", "clips_per_sample": clips},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": th.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------------------------- B200 arm
def kernel_breakdown(chain):
    """One eager denoise step with a CUDA event between consecutive launches (queued behind a device-side sleep so
    the host never starves the stream): per-kernel-class device time, algorithmic FLOPs and bytes."""
    chain.set_state(chain.x.clone(), chain.n_steps - 1)
    th.cuda.synchronize()
    reps = 3
    agg = {}
    for rep in range(reps):
        chain.step.fill_(chain.n_steps - 1)
        evs = [th.cuda.Event(enable_timing=True) for _ in range(len(chain.plan) + 1)]
        th.cuda._sleep(int(60e6))  # ~30 ms: lets the host enqueue the whole step first
        evs[0].record()
        for k, op in enumerate(chain.plan):
            op()
            evs[k + 1].record()
        th.cuda.synchronize()
        if rep == 0:
            continue  # first pass warms caches / clocks
        for k, op in enumerate(chain.plan):
            a = agg.setdefault(op.kind, {"ms": 0.0, "flops": 0, "bytes": 0, "launches": 0})
            a["ms"] += evs[k].elapsed_time(evs[k + 1]) / (reps - 1)
            a["flops"] += op.flops / (reps - 1)
            a["bytes"] += op.bytes / (reps - 1)
            a["launches"] += 1 / (reps - 1)
    return agg


def ncu_gemm_traffic(workload):
    """DRAM bytes per GEMM launch (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu launch list of one
    denoise step of this workload (profiles/r01_launch_summary_v9_*.txt); None if there is no capture for it."""
    name = {"tedexp-ours": "tedexp256", "beat-ours": "beat1024x40"}.get(workload)
    if name is None:
        return None
    try:
        for line in open(os.path.join(ROOT, "profiles", f"r01_launch_summary_v9_{name}.txt")):
            f = [v.strip() for v in line.split("|")]
            if len(f) >= 5 and f[0] == "gemm":
                return float(f[4]) * 1e6 / float(f[1])
    except Exception:
        pass
    return None


def run_b200(args):
    import gesture_b200  # noqa: F401
    from gesture_b200 import _lib
    from gesture_b200.engine import chain_for
    from gesture_b200.generator import Generator
    from gesture_b200.model_creation import create_model
    from gesture_b200.presets import preset
    from gesture_b200.synthetic import synthetic_wav
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    params, C, T, L, default_clips = workload_preset(args.workload)
    clips = args.clips or default_clips
    th.manual_seed(0)
    model, diffusion, *_ = create_model(C, params)  # random-init weights (seed 0), as BASELINE.json prescribes
    model.eval().to(dev)
    model.precision, model.graph_steps = args.precision, args.graph_steps
    n_steps = diffusion.num_timesteps
    shape = (clips, C, T)

    # rank r owns clips [r*clips, (r+1)*clips): its own speech, x_T and noise tape
    wav_host = synthetic_wav(clips, L, seed=123 + rank).pin_memory()
    g = th.Generator(device=dev).manual_seed(99 + rank)
    x_T = th.randn(shape, device=dev, generator=g)
    x_host = x_T.cpu().pin_memory()
    wav_dev = wav_host.to(dev)
    chain = chain_for(model, diffusion, shape, "ddpm", dev)
    gathered = th.empty(world * clips, T, C, device=dev) if world > 1 else None

    def one_chain():
        chain.begin(x_T, wav_dev)  # conditioning (speech encoder once per clip) + on-device noise tape + reset
        out = chain.run()["sample"].transpose(1, 2).contiguous()
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        th.cuda.synchronize()

    for _ in range(args.warmup):
        one_chain()
    barrier()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            one_chain()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = th.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    ms_per_chain = ms / args.steps
    value = world * clips * T * args.steps / (ms / 1e3)

    # chain-only share: graph replays alone (no conditioning, no tape generation)
    chain.begin(x_T, wav_dev)
    th.cuda.synchronize()
    c0, c1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    c0.record()
    chain.run()
    c1.record()
    th.cuda.synchronize()
    ms_replay = c0.elapsed_time(c1)

    # end-to-end through the public API with HOST buffers
    e2e = None
    if not args.no_e2e:
        gen = Generator(model, diffusion)

        def one_e2e():
            poses = gen.generate_sample(shape, wav_host, noise=x_host, sample_alg="ddpm", device=dev, progress=False)
            if world > 1:
                dist.all_gather_into_tensor(gathered, poses.contiguous())
            return poses.cpu()  # device -> host read of this rank's result
        one_e2e()
        barrier()
        s0, s1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        s0.record()
        n_e2e = max(1, min(args.steps, 2))
        for _ in range(n_e2e):
            one_e2e()
        s1.record()
        barrier()
        ems = s0.elapsed_time(s1)
        if world > 1:
            t = th.tensor([ems], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = t.item()
        e2e = {"value": world * clips * T * n_e2e / (ems / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": wav_host.numel() * 4 + x_host.numel() * 4, "d2h_bytes_per_step": clips * T * C * 4}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel class (the tcgen05 GEMM), measured live with CUDA events
    agg = kernel_breakdown(chain)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
    gm = {k: sum(agg[c][k] for c in ("gemm", "gemm_ln") if c in agg) for k in ("ms", "flops", "bytes", "launches")}
    ach = gm["flops"] / (gm["ms"] * 1e-3) / 1e12
    step_ms = sum(a["ms"] for a in agg.values())
    total_flops = sum(a["flops"] for a in agg.values())
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_tn_kernel" + (" + gemm_resid_ln_kernel" if "gemm_ln" in agg else "") + " (tcgen05+TMA)", "achieved": ach, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": ncu_gemm_traffic(args.workload),
                "traffic_note": "average DRAM bytes per GEMM launch, ncu launch list of one step (profiles/r01_launch_summary_v9_*.txt); "
                                "algorithmic bytes per launch = %.0f" % (gm["bytes"] / gm["launches"]),
                "peak_source": peak_src,
                "launches_per_step": gm["launches"], "share_of_step": gm["ms"] / step_ms,
                "flops_per_step": gm["flops"], "avg_launch_us": 1e3 * gm["ms"] / gm["launches"]}
    breakdown = {k: {"ms_per_step": round(a["ms"], 4), "launches": round(a["launches"]),
                     "tflops": round(a["flops"] / (a["ms"] * 1e-3) / 1e12, 2) if a["ms"] else 0,
                     "gbs": round(a["bytes"] / (a["ms"] * 1e-3) / 1e9, 1) if a["ms"] else 0} for k, a in agg.items()}

    cpu = None
    if not args.no_cpu_baseline:
        ref = CpuReference(args.workload, 1)
        ref.sample(2)  # warm-up
        r = ref.sample(40 if args.workload == "tedexp-ours" else 150)
        cpu = {"value": r["frames_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"1 clip, first {r['denoise_steps']} of 1000 denoise steps as shipped (speech encoder re-run every step), "
                         f"{r['seconds']:.1f} s CPU, extrapolated to the full chain; {r['ms_per_denoise_step']:.1f} ms/denoise-step",
               "host_cpus": os.cpu_count()}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_chain, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "bf16 operands / fp32 activations", "data": "synthetic",
            "config": {"workload": f"{args.workload} full {n_steps}-step DDPM chain, {clips} clips/GPU x {T} frames, random-init weights "
                                   "(seed 0), synthetic speech, CUDA-graphed chain", "clips_per_gpu": clips, "frames": T,
                       "d_pose": C, "denoise_steps": n_steps, "graph_steps": args.graph_steps or n_steps,
                       "l2_policy": "inputs_larger_than_L2 (per-step activations + 1000-step noise tape >> 126 MB)",
                       "parallelism": f"clip-sharded x{world}, all_gather of poses"},
            "ms_per_denoise_step": ms_replay / n_steps, "chain_replay_ms": ms_replay,
            # per chain, outside the 1000 replays: speech encoder + conditioning GEMMs + on-device noise tape
            "chain_begin_ms": ms_per_chain - ms_replay, "speech_encoder": chain.speech_impl,
            "step_flops_executed": total_flops,
            "model_tflops_chain": total_flops * n_steps / (ms_replay * 1e-3) / 1e12,
            "clocks": clk.summary(), "e2e": e2e,
            "gpu_launches": len(chain.plan) * n_steps * args.steps, "kernels_per_denoise_step": len(chain.plan),
            "lib_launch_counter": int(lib.gd_launch_count()),
            "roofline": roofline, "kernel_breakdown": breakdown, "cpu_baseline": cpu}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def emit(line):
    """The ONE JSON line of the contract, on the real stdout."""
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


if __name__ == "__main__":
    # Libraries print banners on fd 1 (NCCL: "NCCL version ..." when NCCL_DEBUG is set in the environment); keep stdout for
    # the JSON line alone by pointing fd 1 at stderr for everything else.
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
