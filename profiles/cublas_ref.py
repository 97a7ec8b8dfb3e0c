"""cuBLAS (torch.matmul, bf16) on the denoiser's GEMM shapes: the library ceiling our tcgen05 kernel is compared with."""
import json
import torch as th
from torch.profiler import profile, ProfilerActivity

shapes = [(35328, 1536, 512), (35328, 2048, 512), (35328, 512, 2048), (35328, 512, 512), (8704, 1536, 512),
          (40960, 768, 256), (40960, 1024, 256), (8192, 8192, 8192)]
for M, N, K in shapes:
    sets = [(th.randn(M, K, device="cuda").bfloat16(), th.randn(N, K, device="cuda").bfloat16(),
             th.empty(M, N, device="cuda", dtype=th.bfloat16)) for _ in range(3)]
    for a, w, o in sets:
        th.matmul(a, w.t(), out=o)
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    iters = 20
    e0.record()
    for i in range(iters * 3):
        a, w, o = sets[i % 3]
        th.matmul(a, w.t(), out=o)
    e1.record()
    th.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (iters * 3)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        th.matmul(sets[0][0], sets[0][1].t(), out=sets[0][2])
        th.cuda.synchronize()
    names = [e.key for e in prof.key_averages() if "memcpy" not in e.key.lower()]
    print(json.dumps({"M": M, "N": N, "K": K, "us": round(us, 2), "TFLOPs": round(2.0 * M * N * K / us * 1e-6, 1),
                      "kernel": names[:2]}))
