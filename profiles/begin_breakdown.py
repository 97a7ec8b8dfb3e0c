"""Where chain.begin() spends its time: speech encoder (once per clip), noise tape, the rest.

    python profiles/begin_breakdown.py [--workload beat-ours|tedexp-ours]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402
import bench  # noqa: E402
from gesture_b200.engine import chain_for  # noqa: E402
from gesture_b200.model_creation import create_model  # noqa: E402
from gesture_b200.synthetic import synthetic_wav  # noqa: E402


def timed(fn, reps=3):
    fn()
    th.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    th.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="beat-ours")
    ap.add_argument("--encoder-only", type=int, default=0, help="N: run the speech encoder on N clips twice and exit (ncu target)")
    args = ap.parse_args()
    params, C, T, L, clips = bench.workload_preset(args.workload)
    th.manual_seed(0)
    model, diffusion, *_ = create_model(C, params)
    model.eval().to("cuda")
    if args.encoder_only:
        clips = args.encoder_only
    shape = (clips, C, T)
    chain = chain_for(model, diffusion, shape, "ddpm", "cuda", use_graph=False)
    x_T = th.randn(shape, device="cuda")
    wav = synthetic_wav(clips, L).cuda()
    if args.encoder_only:
        for _ in range(2):
            chain._speech_features(wav)
        th.cuda.synchronize()
        print("ok")
        return
    row = {"workload": args.workload, "clips": clips, "speech_impl": chain.speech_impl,
           "encoder_chunk": chain.encoder_chunk if chain.speech_impl == "torch" else chain.native_encoder_chunk}
    row["begin_ms"] = timed(lambda: chain.begin(x_T, wav))
    row["begin_no_tape_ms"] = timed(lambda: chain.begin(x_T, wav, need_tape=False))
    row["speech_encoder_ms"] = timed(lambda: chain._speech_features(wav))
    row["conditioning_ms"] = timed(lambda: chain._conditioning(wav))
    enc = model.speech_encoder
    with th.no_grad():
        w = wav[:chain.encoder_chunk].float()
        row["mel_chunk_ms"] = timed(lambda: enc.mel_spec_norm(enc.wav2spec(w) + 1e-6))
        mel = enc.mel_spec_norm(enc.wav2spec(w) + 1e-6)
        row["resnet_chunk_ms"] = timed(lambda: enc.wav_encoder(mel))
        th.backends.cudnn.allow_tf32 = False
        row["resnet_chunk_fp32_ms"] = timed(lambda: enc.wav_encoder(mel))
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in row.items()}))


if __name__ == "__main__":
    main()
