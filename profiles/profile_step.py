"""Profiling target: tedexp-ours, 256 clips, three eager denoise steps of the sampling chain (no graph), so that
ncu sees every kernel of a step as its own launch.  Usage (on the GPU box):
    python profiles/profile_step.py [workload] [clips]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th  # noqa: E402
import gesture_b200  # noqa: E402,F401
from gesture_b200.engine import chain_for  # noqa: E402
from gesture_b200.model_creation import create_model  # noqa: E402
from gesture_b200.presets import preset  # noqa: E402
from gesture_b200.synthetic import synthetic_wav  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "tedexp-ours"
clips = int(sys.argv[2]) if len(sys.argv) > 2 else 256
params, C, T, L = preset(workload)
th.manual_seed(0)
model, diffusion, *_ = create_model(C, params)
model.eval().to("cuda")
chain = chain_for(model, diffusion, (clips, C, T), "ddpm", "cuda", use_graph=False)
x_T = th.randn(clips, C, T, device="cuda")
chain.begin(x_T, synthetic_wav(clips, L).cuda())
for _ in range(3):
    chain.step_eager()
th.cuda.synchronize()
print("ok", len(chain.plan), "kernels per step")
